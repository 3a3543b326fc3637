// Device helpers shared by the backward-warp kernels (warp_t.cu, warp_t_ws.cu). Static: every translation unit gets
// its own copy (the library is built without relocatable device code).
#pragma once
#include "ofk_common.cuh"

namespace ofk {

// Pixels whose taps touch the border (or whose coordinates leave the fast-quantisation range): exact but slow, out
// of line, by-value in / out (no stack traffic). Returns the 3 result bytes in bits 0..23 and the validity in bit 24.
static __device__ __noinline__ uint32_t border_px_u8x3(const uint8_t* __restrict__ p, const uint8_t* __restrict__ pm, float X,
                                                float Y, int H, int W, int half_even, int rule) {
    const QCoord qx = quantise(X), qy = quantise(Y);
    const int ix = qx.i, iy = qy.i;
    const QWeights w = qweights(qx.f, qy.f);
    const bool in[4] = {(unsigned)ix < (unsigned)W && (unsigned)iy < (unsigned)H,
                        (unsigned)(ix + 1) < (unsigned)W && (unsigned)iy < (unsigned)H,
                        (unsigned)ix < (unsigned)W && (unsigned)(iy + 1) < (unsigned)H,
                        (unsigned)(ix + 1) < (unsigned)W && (unsigned)(iy + 1) < (unsigned)H};
    const long long o = (long long)iy * W + ix;
    const long long off[4] = {o, o + 1, o + W, o + W + 1};
    const int wi[4] = {w.w00, w.w01, w.w10, w.w11};
    int acc[3] = {0, 0, 0}, S = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!in[k]) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += (int)p[off[k] * 3 + c] * wi[k];
        if (pm == nullptr || pm[off[k]]) S += wi[k];
    }
    uint32_t v = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int a = acc[c];
        const uint32_t r = half_even ? (uint32_t)(a + 511 + ((a >> 10) & 1)) >> 10 : (uint32_t)(a + 512) >> 10;
        v |= r << (8 * c);
    }
    return v | (mask_rule_pass(S, rule) ? (1u << 24) : 0u);
}

}  // namespace ofk
