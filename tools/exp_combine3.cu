// Experiment harness for the mode-3 composition kernel (not part of the library): times candidate kernels against
// the library's ofk_combine3 on the headline workload shape and checks them bit for bit against it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/exp_combine3.cu \
//        -Loflibnumpy_b200/lib -loflib_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../oflibnumpy_b200/lib' -o tools/exp_combine3
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "../include/oflib_b200.h"
#include "c3_ws.cuh"

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);  \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

struct Affine {
    float a, b, c, d, e, f;  // u = a*x + b*y + c ; v = d*x + e*y + f
};

__global__ void gen_flow(float2* out, uint8_t* mask, const Affine* aff, int H, int W, unsigned seed) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const Affine A = aff[n];
    const size_t i = ((size_t)n * H + y) * W + x;
    out[i] = make_float2(A.a * x + A.b * y + A.c, A.d * x + A.e * y + A.f);
    unsigned h = (unsigned)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    mask[i] = (h % 100u) >= 2u;   // 2 % invalid
}

// ------------------------------------------------------------------------------------------------ TEX variant
__device__ __forceinline__ void quant(float X, int& i, int& f) {
    const int bits = __float_as_int(__fmaf_rn(X, 32.0f, 12582912.0f)) - 0x4B400000;
    i = bits >> 5;
    f = bits & 31;
}

template <int ROWS>
__global__ void __launch_bounds__(256) c3_tex(const float2* __restrict__ P, const uint8_t* __restrict__ Pm,
                                              const cudaTextureObject_t* __restrict__ texG,
                                              const cudaTextureObject_t* __restrict__ texGm, float2* __restrict__ out,
                                              uint8_t* __restrict__ om, int H, int W, float sign) {
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + lane;
    const int y0 = (blockIdx.y * 8 + wrp) * ROWS;
    const int n = blockIdx.z;
    if (x >= W) return;
    const size_t fbase = (size_t)n * H * W;
    const cudaTextureObject_t tg = texG[n], tm = texGm[n];
    float2 p[ROWS];
    unsigned pm[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const int y = min(y0 + j, H - 1);
        p[j] = P[fbase + (size_t)y * W + x];
        pm[j] = Pm[fbase + (size_t)y * W + x];
    }
    float4 gu[ROWS], gv[ROWS];
    uchar4 gm[ROWS];
    int fa[ROWS], fb[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        float X = __fmaf_rn(sign, p[j].x, (float)x), Y = __fmaf_rn(sign, p[j].y, (float)(y0 + j));
        X = fminf(fmaxf(X, -2.0f), (float)(W + 1));
        Y = fminf(fmaxf(Y, -2.0f), (float)(H + 1));
        int ix, iy;
        quant(X, ix, fa[j]);
        quant(Y, iy, fb[j]);
        const float tx = (float)(ix + 1), ty = (float)(iy + 1);
        gu[j] = tex2Dgather<float4>(tg, tx, ty, 0);
        gv[j] = tex2Dgather<float4>(tg, tx, ty, 1);
        gm[j] = tex2Dgather<uchar4>(tm, tx, ty, 0);
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const int y = y0 + j;
        if (y >= H) break;
        const int a = fa[j], b = fb[j];
        // gather order: x = (x0,y1), y = (x1,y1), z = (x1,y0), w = (x0,y0)
        const float t00u = gu[j].w, t01u = gu[j].z, t10u = gu[j].x, t11u = gu[j].y;
        const float t00v = gv[j].w, t01v = gv[j].z, t10v = gv[j].x, t11v = gv[j].y;
        const unsigned za = a == 0, zb = b == 0;
        const unsigned strict = gm[j].w & (gm[j].z | za) & (gm[j].x | zb) & (gm[j].y | za | zb);
        const float ffa = (float)a * (1.0f / 32.0f), ffb = (float)b * (1.0f / 32.0f);
        const float na = 1.0f - ffa, nb = 1.0f - ffb;
        const float f00 = __fmul_rn(na, nb), f01 = __fmul_rn(ffa, nb), f10 = __fmul_rn(na, ffb), f11 = __fmul_rn(ffa, ffb);
        const float su = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00u, f00), __fmul_rn(t01u, f01)), __fmul_rn(t10u, f10)),
                                   __fmul_rn(t11u, f11));
        const float sv = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00v, f00), __fmul_rn(t01v, f01)), __fmul_rn(t10v, f10)),
                                   __fmul_rn(t11v, f11));
        const size_t idx = fbase + (size_t)y * W + x;
        out[idx] = make_float2(__fadd_rn(p[j].x, su), __fadd_rn(p[j].y, sv));
        om[idx] = (uint8_t)(pm[j] & strict & 1u);
    }
}


// ------------------------------------------------------------------------------------------------ streaming probes
// same byte mix as the composition (read 8+1+8+1, write 8+1 per pixel) without the gather
template <int ROWS>
__global__ void __launch_bounds__(256) stream32(const float2* __restrict__ P, const uint8_t* __restrict__ Pm,
                                                const float2* __restrict__ G, const uint8_t* __restrict__ Gm,
                                                float2* __restrict__ out, uint8_t* __restrict__ om, int H, int W) {
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + lane;
    const int y0 = (blockIdx.y * 8 + wrp) * ROWS;
    const size_t fbase = (size_t)blockIdx.z * H * W;
    if (x >= W) return;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const int y = y0 + j;
        if (y >= H) break;
        const size_t i = fbase + (size_t)y * W + x;
        const float2 p = P[i], g = G[i];
        out[i] = make_float2(p.x + g.x, p.y + g.y);
        om[i] = Pm[i] & Gm[i];
    }
}
// 4 consecutive pixels per thread: 2x float4 + one 32-bit mask word per operand
__global__ void __launch_bounds__(256) stream128(const float4* __restrict__ P, const uint32_t* __restrict__ Pm,
                                                 const float4* __restrict__ G, const uint32_t* __restrict__ Gm,
                                                 float4* __restrict__ out, uint32_t* __restrict__ om, size_t quads) {
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (size_t)gridDim.x * blockDim.x) {
        const float4 p0 = P[2 * q], p1 = P[2 * q + 1], g0 = G[2 * q], g1 = G[2 * q + 1];
        out[2 * q] = make_float4(p0.x + g0.x, p0.y + g0.y, p0.z + g0.z, p0.w + g0.w);
        out[2 * q + 1] = make_float4(p1.x + g1.x, p1.y + g1.y, p1.z + g1.z, p1.w + g1.w);
        om[q] = Pm[q] & Gm[q];
    }
}
__global__ void __launch_bounds__(256) copy16(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) b[q] = a[q];
}


// ------------------------------------------------------------------------------------------------ pipe-mix probe
// VM / MM: how the vector / mask taps are fetched: 0 = not at all, 1 = TLD4, 2 = LDG (indices clamped; speed only)
template <int ROWS, int VM, int MM>
__global__ void __launch_bounds__(256) c3_mix(const float2* __restrict__ P, const uint8_t* __restrict__ Pm,
                                              const float2* __restrict__ G, const uint8_t* __restrict__ Gm,
                                              const cudaTextureObject_t* __restrict__ texG,
                                              const cudaTextureObject_t* __restrict__ texGm, float2* __restrict__ out,
                                              uint8_t* __restrict__ om, int H, int W, float sign) {
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + lane;
    const int y0 = (blockIdx.y * 8 + wrp) * ROWS;
    const int n = blockIdx.z;
    if (x >= W) return;
    const size_t fbase = (size_t)n * H * W;
    const cudaTextureObject_t tg = texG[n], tm = texGm[n];
    const float2* Gf = G + fbase;
    const uint8_t* Gmf = Gm + fbase;
    float2 p[ROWS];
    unsigned pm[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const int y = min(y0 + j, H - 1);
        p[j] = P[fbase + (size_t)y * W + x];
        pm[j] = Pm[fbase + (size_t)y * W + x];
    }
    float2 t[ROWS][4];
    unsigned m[ROWS][4];
    int fa[ROWS], fb[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        float X = __fmaf_rn(sign, p[j].x, (float)x), Y = __fmaf_rn(sign, p[j].y, (float)(y0 + j));
        X = fminf(fmaxf(X, -2.0f), (float)(W + 1));
        Y = fminf(fmaxf(Y, -2.0f), (float)(H + 1));
        int ix, iy;
        quant(X, ix, fa[j]);
        quant(Y, iy, fb[j]);
        const float tx = (float)(ix + 1), ty = (float)(iy + 1);
        const int cx = min(max(ix, 0), W - 2), cy = min(max(iy, 0), H - 2);
        const unsigned o = (unsigned)(cy * W + cx);
#pragma unroll
        for (int k = 0; k < 4; ++k) { t[j][k] = make_float2(1.f, 2.f); m[j][k] = 1; }
        if (VM == 1) {
            const float4 gu = tex2Dgather<float4>(tg, tx, ty, 0), gv = tex2Dgather<float4>(tg, tx, ty, 1);
            t[j][0] = make_float2(gu.w, gv.w); t[j][1] = make_float2(gu.z, gv.z);
            t[j][2] = make_float2(gu.x, gv.x); t[j][3] = make_float2(gu.y, gv.y);
        } else if (VM == 2) {
            t[j][0] = __ldg(Gf + o); t[j][1] = __ldg(Gf + o + 1); t[j][2] = __ldg(Gf + o + W); t[j][3] = __ldg(Gf + o + W + 1);
        }
        if (MM == 1) {
            const uchar4 g = tex2Dgather<uchar4>(tm, tx, ty, 0);
            m[j][0] = g.w; m[j][1] = g.z; m[j][2] = g.x; m[j][3] = g.y;
        } else if (MM == 2) {
            m[j][0] = __ldg(Gmf + o); m[j][1] = __ldg(Gmf + o + 1); m[j][2] = __ldg(Gmf + o + W); m[j][3] = __ldg(Gmf + o + W + 1);
        }
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const int y = y0 + j;
        if (y >= H) break;
        const int a = fa[j], b = fb[j];
        const unsigned za = a == 0, zb = b == 0;
        const unsigned strict = m[j][0] & (m[j][1] | za) & (m[j][2] | zb) & (m[j][3] | za | zb);
        const float ffa = (float)a * (1.0f / 32.0f), ffb = (float)b * (1.0f / 32.0f);
        const float na = 1.0f - ffa, nb = 1.0f - ffb;
        const float f00 = __fmul_rn(na, nb), f01 = __fmul_rn(ffa, nb), f10 = __fmul_rn(na, ffb), f11 = __fmul_rn(ffa, ffb);
        const float su = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[j][0].x, f00), __fmul_rn(t[j][1].x, f01)), __fmul_rn(t[j][2].x, f10)),
                                   __fmul_rn(t[j][3].x, f11));
        const float sv = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[j][0].y, f00), __fmul_rn(t[j][1].y, f01)), __fmul_rn(t[j][2].y, f10)),
                                   __fmul_rn(t[j][3].y, f11));
        const size_t idx = fbase + (size_t)y * W + x;
        out[idx] = make_float2(__fadd_rn(p[j].x, su), __fadd_rn(p[j].y, sv));
        om[idx] = (uint8_t)(pm[j] & strict & 1u);
    }
}

static cudaTextureObject_t make_tex(const void* base, int H, int W, int esize) {
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = const_cast<void*>(base);
    rd.res.pitch2D.width = W;
    rd.res.pitch2D.height = H;
    rd.res.pitch2D.pitchInBytes = (size_t)W * esize;
    rd.res.pitch2D.desc = esize == 8 ? cudaCreateChannelDesc<float2>() : cudaCreateChannelDesc<unsigned char>();
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t t = 0;
    CK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
    return t;
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 64, H = 1080, W = 1920;
    const int iters = argc > 2 ? atoi(argv[2]) : 10;
    const int mode = argc > 3 ? atoi(argv[3]) : 0;   // 0 cfg4-like, 1 zero flow, 2 translation, 3 rotation 10 deg, 4 rotation 3 deg
    printf("== mode %d\n", mode);
    const size_t px = (size_t)N * H * W;
    float2 *A, *B, *ref, *out;
    uint8_t *Am, *Bm, *refm, *outm;
    int* flags;
    CK(cudaMalloc(&A, px * 8)); CK(cudaMalloc(&B, px * 8)); CK(cudaMalloc(&ref, px * 8)); CK(cudaMalloc(&out, px * 8));
    CK(cudaMalloc(&Am, px)); CK(cudaMalloc(&Bm, px)); CK(cudaMalloc(&refm, px)); CK(cudaMalloc(&outm, px));
    CK(cudaMalloc(&flags, sizeof(int) * 2 * N));
    std::vector<Affine> ha(N), hb(N);
    srand(1);
    auto rnd = [](float lo, float hi) { return lo + (hi - lo) * (rand() / (float)RAND_MAX); };
    for (int n = 0; n < N; ++n) {
        for (int k = 0; k < 2; ++k) {
            float th = rnd(-10.f, 10.f) * 3.14159265f / 180.f, s = rnd(0.9f, 1.1f), tx = rnd(-20, 20), ty = rnd(-20, 20);
            if (mode == 1) { th = 0; s = 1; tx = ty = 0; }
            if (mode == 2) { th = 0; s = 1; }
            if (mode == 3) { th = 10.f * 3.14159265f / 180.f; s = 1; tx = ty = 0; }
            if (mode == 4) { th = 3.f * 3.14159265f / 180.f; s = 1; tx = ty = 0; }
            if (mode == 5) { th = 0; s = 1.1f; tx = ty = 0; }
            const float cx = W / 2.f, cy = H / 2.f, c = cosf(th) * s, sn = sinf(th) * s;
            Affine F;   // F = (p - ctr) - R (p - ctr) - t
            F.a = 1.f - c; F.b = sn; F.c = -(1.f - c) * cx - sn * cy - tx;
            F.d = -sn; F.e = 1.f - c; F.f = sn * cx - (1.f - c) * cy - ty;
            (k ? hb : ha)[n] = F;
        }
    }
    Affine *da, *db;
    CK(cudaMalloc(&da, sizeof(Affine) * N)); CK(cudaMalloc(&db, sizeof(Affine) * N));
    CK(cudaMemcpy(da, ha.data(), sizeof(Affine) * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), sizeof(Affine) * N, cudaMemcpyHostToDevice));
    dim3 gg((W + 255) / 256, H, N);
    gen_flow<<<gg, 256>>>(A, Am, da, H, W, 1u);
    gen_flow<<<gg, 256>>>(B, Bm, db, H, W, 2u);
    CK(cudaDeviceSynchronize());

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double gb = (double)px * 27 / 1e9;
    auto report = [&](const char* name, float ms) {
        printf("%-28s %8.3f ms  %8.1f GB/s (27 B/px)  %8.1f Mpx/s\n", name, ms, gb / (ms * 1e-3), px / (ms * 1e-3) / 1e6);
    };

    // ---- library kernel (reference result), ref 't': pointwise = B, gathered = A
    for (int w = 0; w < 2; ++w)
        if (ofk_combine3((float*)A, Am, (float*)B, Bm, 't', 0.f, (float*)ref, refm, flags, N, H, W, 0) != 0) {
            printf("ofk_combine3: %s\n", ofk_last_error());
            return 1;
        }
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) ofk_combine3((float*)A, Am, (float*)B, Bm, 't', 0.f, (float*)ref, refm, flags, N, H, W, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    report("library ofk_combine3", ms / iters);
    for (int i = 0; i < iters; ++i) ofk_combine3((float*)A, Am, (float*)B, Bm, 't', 0.f, (float*)ref, refm, nullptr, N, H, W, 0);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) ofk_combine3((float*)A, Am, (float*)B, Bm, 't', 0.f, (float*)ref, refm, nullptr, N, H, W, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    report("library, no flags", ms / iters);


#define TIME(name, bytes_px, ...)                                                              \
    do {                                                                                       \
        for (int w = 0; w < 2; ++w) { __VA_ARGS__; }                                           \
        CK(cudaDeviceSynchronize());                                                           \
        CK(cudaEventRecord(e0));                                                               \
        for (int i = 0; i < iters; ++i) { __VA_ARGS__; }                                       \
        CK(cudaEventRecord(e1));                                                               \
        CK(cudaEventSynchronize(e1));                                                          \
        CK(cudaEventElapsedTime(&ms, e0, e1));                                                 \
        printf("%-28s %8.3f ms  %8.1f GB/s (%d B/px)\n", name, ms / iters, (double)px * bytes_px / 1e9 / (ms / iters * 1e-3), bytes_px); \
    } while (0)
    {
        dim3 g1((W + 31) / 32, (H + 7) / 8, N), g4((W + 31) / 32, (H + 31) / 32, N);
        TIME("stream32 rows=1", 27, (stream32<1><<<g1, 256>>>(B, Bm, A, Am, out, outm, H, W)));
        TIME("stream32 rows=4", 27, (stream32<4><<<g4, 256>>>(B, Bm, A, Am, out, outm, H, W)));
        TIME("stream128 grid=148*8", 27, (stream128<<<148 * 8, 256>>>((float4*)B, (uint32_t*)Bm, (float4*)A, (uint32_t*)Am, (float4*)out, (uint32_t*)outm, px / 4)));
        TIME("stream128 grid=full", 27, (stream128<<<(unsigned)(px / 4 / 256), 256>>>((float4*)B, (uint32_t*)Bm, (float4*)A, (uint32_t*)Am, (float4*)out, (uint32_t*)outm, px / 4)));
        TIME("copy16 grid=148*8", 16, (copy16<<<148 * 8, 256>>>((float4*)A, (float4*)out, px / 2)));
        TIME("copy16 grid=full", 16, (copy16<<<(unsigned)(px / 2 / 256), 256>>>((float4*)A, (float4*)out, px / 2)));
        TIME("cudaMemcpy d2d", 16, CK(cudaMemcpyAsync(out, A, px * 8, cudaMemcpyDeviceToDevice)));
    }
    // ---- texture objects
    std::vector<cudaTextureObject_t> tg(N), tm(N);
    auto t0 = std::chrono::steady_clock::now();
    for (int n = 0; n < N; ++n) {
        tg[n] = make_tex(A + (size_t)n * H * W, H, W, 8);
        tm[n] = make_tex(Am + (size_t)n * H * W, H, W, 1);
    }
    auto t1 = std::chrono::steady_clock::now();
    printf("texture object creation: %.1f us per object\n",
           std::chrono::duration<double, std::micro>(t1 - t0).count() / (2 * N));
    cudaTextureObject_t *dtg, *dtm;
    CK(cudaMalloc(&dtg, 8 * N)); CK(cudaMalloc(&dtm, 8 * N));
    CK(cudaMemcpy(dtg, tg.data(), 8 * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dtm, tm.data(), 8 * N, cudaMemcpyHostToDevice));

    std::vector<float2> h_ref(H * W), h_out(H * W);
    std::vector<uint8_t> hm_ref(H * W), hm_out(H * W);
    auto check = [&](const char* name) {
        size_t bad = 0, badm = 0;
        for (int n : {0, N / 2, N - 1}) {
            CK(cudaMemcpy(h_ref.data(), ref + (size_t)n * H * W, (size_t)H * W * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h_out.data(), out + (size_t)n * H * W, (size_t)H * W * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hm_ref.data(), refm + (size_t)n * H * W, (size_t)H * W, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hm_out.data(), outm + (size_t)n * H * W, (size_t)H * W, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < (size_t)H * W; ++i) {
                if (memcmp(&h_ref[i], &h_out[i], 8) != 0) {
                    if (bad < 5) printf("  frame %d px %zu (x=%zu y=%zu): ref (%.9g,%.9g) got (%.9g,%.9g)\n", n, i, i % W, i / W, h_ref[i].x, h_ref[i].y, h_out[i].x, h_out[i].y);
                    ++bad;
                }
                if (hm_ref[i] != hm_out[i]) ++badm;
            }
        }
        printf("%-28s mismatches: vecs %zu, mask %zu (3 frames)\n", name, bad, badm);
    };

#define RUN_TEX(R)                                                                                       \
    do {                                                                                                 \
        dim3 grid((W + 31) / 32, (H + 8 * R - 1) / (8 * R), N);                                          \
        CK(cudaMemset(out, 0xff, px * 8));                                                               \
        CK(cudaMemset(outm, 0xff, px));                                                                  \
        for (int w = 0; w < 2; ++w) c3_tex<R><<<grid, 256>>>(B, Bm, dtg, dtm, out, outm, H, W, -1.0f);   \
        CK(cudaDeviceSynchronize());                                                                     \
        CK(cudaEventRecord(e0));                                                                         \
        for (int i = 0; i < iters; ++i) c3_tex<R><<<grid, 256>>>(B, Bm, dtg, dtm, out, outm, H, W, -1.0f); \
        CK(cudaEventRecord(e1));                                                                         \
        CK(cudaEventSynchronize(e1));                                                                    \
        CK(cudaEventElapsedTime(&ms, e0, e1));                                                           \
        report("tex tld4 rows=" #R, ms / iters);                                                         \
        check("tex tld4 rows=" #R);                                                                      \
    } while (0)
    RUN_TEX(1);
    RUN_TEX(2);
    RUN_TEX(4);

#define RUN_MIX(R, VM, MM)                                                                               \
    do {                                                                                                 \
        dim3 grid((W + 31) / 32, (H + 8 * R - 1) / (8 * R), N);                                          \
        TIME("mix rows=" #R " vec=" #VM " mask=" #MM, 27, (c3_mix<R, VM, MM><<<grid, 256>>>(B, Bm, A, Am, dtg, dtm, out, outm, H, W, -1.0f))); \
    } while (0)

#define RUN_WS(NP, NB, LA, CPS)                                                                          \
    do {                                                                                                 \
        CK(cudaMemset(out, 0xff, px * 8));                                                               \
        CK(cudaMemset(outm, 0xff, px));                                                                  \
        int rc_ = c3ws::launch<NP, NB, LA>((float*)B, Bm, (float*)A, Am, -1.0f, (float*)out, outm, N, H, W, CPS, 148, 0); \
        if (rc_ != 1) { printf("ws launch rc=%d\n", rc_); break; }                                      \
        cudaError_t e_ = cudaDeviceSynchronize();                                                        \
        if (e_ != cudaSuccess) { printf("ws<%d,%d,%d> x%d: %s\n", NP, NB, LA, CPS, cudaGetErrorString(e_)); return 1; } \
        TIME("ws NP=" #NP " NB=" #NB " LA=" #LA " cps=" #CPS, 27, (c3ws::launch<NP, NB, LA>((float*)B, Bm, (float*)A, Am, -1.0f, (float*)out, outm, N, H, W, CPS, 148, 0))); \
        check("ws NP=" #NP " NB=" #NB " LA=" #LA " cps=" #CPS);                                          \
    } while (0)
    RUN_WS(5, 2, 2, 2);
    if (argc > 5) return 0;
    RUN_WS(6, 2, 2, 2);
    RUN_WS(6, 2, 3, 2);
    if (argc > 4) return 0;
    RUN_MIX(4, 0, 0); RUN_MIX(4, 1, 0); RUN_MIX(4, 0, 1); RUN_MIX(4, 1, 1);
    RUN_MIX(4, 2, 0); RUN_MIX(4, 0, 2); RUN_MIX(4, 2, 2); RUN_MIX(4, 2, 1); RUN_MIX(4, 1, 2);
    RUN_MIX(2, 2, 1); RUN_MIX(2, 2, 2); RUN_MIX(1, 2, 1); RUN_MIX(1, 2, 2);
    return 0;
}
