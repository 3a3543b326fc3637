/* Plain C use of the drop-in boundary (include/oflib_b200.h): no Python, no torch.
 *
 *   gcc -O2 -Iinclude examples/abi_example.c -Loflibnumpy_b200/lib -loflib_b200 -Wl,-rpath,'$ORIGIN/../oflibnumpy_b200/lib' \
 *       -lm -o examples/abi_example && ./examples/abi_example
 *
 * Host-buffer entry points (ofh_*): what `apply_flow(flow, img, 't')` / `Flow.apply(..., return_valid_area=True)` and
 * `combine_flows(a, b, 3, 't')` look like from C. Checks two properties that need no reference: an integer translation
 * reproduces the shifted image exactly (the reference's own test, tests/test_utils.py:277-283), and composing a
 * translation with its inverse gives the zero flow where both are valid.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "oflib_b200.h"

#define CHECK(call)                                                   \
    do {                                                              \
        int rc_ = (call);                                             \
        if (rc_ != OFK_OK) {                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ofk_last_error()); \
            return 1;                                                 \
        }                                                             \
    } while (0)

int main(void) {
    const int N = 3, H = 270, W = 480, C = 3, dx = 7, dy = -4; /* W % 16 == 0: takes the TMA kernels */
    const size_t px = (size_t)H * W;
    int ndev = 0;
    if (ofk_rt_device_count(&ndev) != OFK_OK || ndev < 1) {
        fprintf(stderr, "no CUDA device: %s\n", ofk_last_error());
        return 2;
    }
    float* flow = malloc(N * px * 2 * sizeof(float));
    float* back = malloc(N * px * 2 * sizeof(float));
    float* comb = malloc(N * px * 2 * sizeof(float));
    uint8_t* img = malloc(N * px * C);
    uint8_t* out = malloc(N * px * C);
    uint8_t* valid = malloc(N * px);
    uint8_t* cmask = malloc(N * px);
    int flags[2 * 3];
    for (size_t i = 0; i < N * px; ++i) {
        flow[2 * i] = (float)dx; flow[2 * i + 1] = (float)dy;         /* target-referenced: out[p] = img[p - flow] */
        back[2 * i] = (float)-dx; back[2 * i + 1] = (float)-dy;
    }
    for (size_t i = 0; i < N * px * C; ++i) img[i] = (uint8_t)((i * 2654435761u) >> 24);

    /* Flow.apply(img, return_valid_area=True) for a uint8 image without target mask: int16 arithmetic, rule S > 1/2 */
    CHECK(ofh_warp_t(img, OFK_U8, C, OFK_ARITH_RINT, flow, -1.0f, NULL, NULL, out, valid, OFK_RULE_GT_HALF, N, H, W, 0));
    size_t bad = 0;
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const int sx = x - dx, sy = y - dy;
                const int inside = sx >= 0 && sx < W && sy >= 0 && sy < H;
                const size_t o = (size_t)n * px + (size_t)y * W + x;
                if (valid[o] != (uint8_t)inside) { if (bad < 5) printf("  valid mismatch n=%d y=%d x=%d got %d want %d\n", n, y, x, valid[o], inside); ++bad; }
                for (int c = 0; c < C; ++c) {
                    const uint8_t want = inside ? img[((size_t)n * px + (size_t)sy * W + sx) * C + c] : 0;
                    if (out[o * C + c] != want) { if (bad < 5) printf("  value mismatch n=%d y=%d x=%d c=%d got %d want %d\n", n, y, x, c, out[o * C + c], want); ++bad; }
                }
            }
    printf("integer translation: %zu mismatches\n", bad);

    /* combine_flows(flow, back, 3, 't'): back + flow sampled at p - back  ==  0 where the sample is inside */
    CHECK(ofh_combine3(flow, NULL, back, NULL, 't', 0.0f, comb, cmask, flags, N, H, W, 0));
    size_t bad2 = 0;
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t o = (size_t)n * px + (size_t)y * W + x;
                const int sx = x + dx, sy = y + dy;
                const int inside = sx >= 0 && sx < W && sy >= 0 && sy < H;
                if (cmask[o] != (uint8_t)inside) ++bad2;
                if (inside && (comb[2 * o] != 0.0f || comb[2 * o + 1] != 0.0f)) ++bad2;
            }
    printf("translation o inverse: %zu mismatches, flags A %d B %d, kernel launches %llu\n", bad2, flags[0], flags[1],
           ofk_rt_launch_count());
    ofh_release();
    free(flow); free(back); free(comb); free(img); free(out); free(valid); free(cmask);
    if (bad || bad2) return 1;
    printf("abi example ok\n");
    return 0;
}
