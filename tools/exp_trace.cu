// Timeline of the composition pipeline of one CTA (global-timer stamps per tile): where producer and consumers wait.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DOFK_TRACE tools/exp_trace.cu -o tools/exp_trace
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../oflibnumpy_b200/csrc/combine3_ws.cu"
#include "../oflibnumpy_b200/csrc/warp_t_ws.cu"
namespace ofk {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); }
std::atomic<unsigned long long> g_launches{0};
std::atomic<unsigned long long> g_paths[4];
int sm_count() { return 148; }
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)
struct Affine { float a, b, c, d, e, f; };
__global__ void gen_flow(float2* out, uint8_t* mask, const Affine* aff, int H, int W, unsigned seed) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const Affine A = aff[n];
    const size_t i = ((size_t)n * H + y) * W + x;
    out[i] = make_float2(A.a * x + A.b * y + A.c, A.d * x + A.e * y + A.f);
    unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    mask[i] = (h % 100u) >= 2u;
}
int main(int argc, char** argv) {
    const int N = 64, H = 1080, W = 1920;
    const size_t px = (size_t)N * H * W;
    float2 *A, *B, *out; uint8_t *Am, *Bm, *outm;
    CK(cudaMalloc(&A, px * 8)); CK(cudaMalloc(&B, px * 8)); CK(cudaMalloc(&out, px * 8));
    CK(cudaMalloc(&Am, px)); CK(cudaMalloc(&Bm, px)); CK(cudaMalloc(&outm, px));
    std::vector<Affine> ha(N), hb(N);
    srand(1);
    auto rnd = [](float lo, float hi) { return lo + (hi - lo) * (rand() / (float)RAND_MAX); };
    for (int n = 0; n < N; ++n) for (int k = 0; k < 2; ++k) {
        float th = rnd(-10.f, 10.f) * 3.14159265f / 180.f, s = rnd(0.9f, 1.1f), tx = rnd(-20, 20), ty = rnd(-20, 20);
        const float cx = W / 2.f, cy = H / 2.f, c = cosf(th) * s, sn = sinf(th) * s;
        Affine F; F.a = 1.f - c; F.b = sn; F.c = -(1.f - c) * cx - sn * cy - tx; F.d = -sn; F.e = 1.f - c; F.f = sn * cx - (1.f - c) * cy - ty;
        (k ? hb : ha)[n] = F;
    }
    Affine *da, *db;
    CK(cudaMalloc(&da, sizeof(Affine) * N)); CK(cudaMalloc(&db, sizeof(Affine) * N));
    CK(cudaMemcpy(da, ha.data(), sizeof(Affine) * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), sizeof(Affine) * N, cudaMemcpyHostToDevice));
    dim3 gg((W + 255) / 256, H, N);
    gen_flow<<<gg, 256>>>(A, Am, da, H, W, 1u);
    gen_flow<<<gg, 256>>>(B, Bm, db, H, W, 2u);
    CK(cudaDeviceSynchronize());
    unsigned long long* dtr;
    CK(cudaMalloc(&dtr, 512 * 16 * 8)); CK(cudaMemset(dtr, 0, 512 * 16 * 8));
    const int mode = argc > 1 ? atoi(argv[1]) : 0;   // 0 composition, 1 image warp (uint8 x3, half-even, geometry mask, flow mask)
    uint8_t *img, *oimg;
    CK(cudaMalloc(&img, px * 3)); CK(cudaMalloc(&oimg, px * 3)); CK(cudaMemset(img, 77, px * 3));
    auto run = [&]() {
        if (mode == 0) ofk::launch_combine3_ws((float*)B, Bm, (float*)A, Am, -1.0f, true, (float*)out, outm, N, H, W, 0);
        else ofk::launch_warp_u8_ws(3, true, img, (float*)A, -1.0f, nullptr, Am, oimg, outm, OFK_RULE_GT_HALF, N, H, W, 0);
    };
    for (int w = 0; w < 3; ++w) run();
    if (false) ofk::launch_combine3_ws((float*)B, Bm, (float*)A, Am, -1.0f, true, (float*)out, outm, N, H, W, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyToSymbol(ofk::ws::g_trace, &dtr, sizeof(dtr)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    run();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d kernel %.3f ms (%.0f GB/s)\n", mode, ms, px * (mode ? 16 : 27) / ms / 1e6);
    std::vector<unsigned long long> tr(512 * 16);
    CK(cudaMemcpy(tr.data(), dtr, tr.size() * 8, cudaMemcpyDeviceToHost));
    const unsigned long long t0 = tr[20 * 16 + 0];
    printf("tile | producer: pfull_wait  bbox  bempty_wait  box_issue  pempty+P_issue | cons w0: wait_start box_arrived(wait) released tile_end | w7: wait gather end   (us rel.)\n");
    double sum_wait0 = 0, sum_tile = 0, sum_bempty = 0, sum_pempty = 0, sum_lat = 0;
    int cnt = 0;
    for (int i = 20; i < 110; ++i) {
        auto T = [&](int s) { return (double)(tr[i * 16 + s] - t0) / 1000.0; };
        if (i < 36) printf("%4d | %7.2f %7.2f %7.2f %7.2f %7.2f | %7.2f %7.2f(%5.2f) %7.2f %7.2f | %7.2f(%5.2f) %7.2f %7.2f | box latency %5.2f\n", i, T(0), T(1), T(2), T(3), T(5),
                           T(8), T(9), T(9) - T(8), T(10), T(11), T(13), T(13) - T(12), T(14), T(15), T(9) - T(4));
        sum_wait0 += T(9) - T(8); sum_tile += T(11) - T(8); sum_bempty += T(3) - T(2); sum_pempty += T(5) - T(4); sum_lat += T(9) - T(4); ++cnt;
    }
    printf("avg per tile: consumer w0 wait %.2f us of %.2f us tile; producer bempty wait %.2f, pempty+P issue %.2f; box issue->consumer saw it %.2f us\n",
           sum_wait0 / cnt, sum_tile / cnt, sum_bempty / cnt, sum_pempty / cnt, sum_lat / cnt);
    return 0;
}
