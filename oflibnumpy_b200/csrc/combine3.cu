// Fused flow composition, mode 3 (flow_class.py:1411-1422 of the reference) for sm_100a.
//
//   ref 't' : out[p] = B[p] + Q(A, p - B[p]);   out_mask[p] = Bm[p] & strict(Am at the taps)
//   ref 's' : out[p] = A[p] + Q(B, p + A[p]);   out_mask[p] = Am[p] & strict(Bm at the taps)
//
// Q is the cv2.remap float32 bilinear sample (1/32-px coordinates, zero border). The reference evaluates this as
// warp (remap of vecs||mask) -> Flow construction -> add -> mask AND, i.e. five full-frame passes plus two masked
// zero tests; here one kernel reads both operands once (27 B/px algorithmic) and also produces the zero-test flags
// that gate the reference's early exits (flow_class.py:1338-1354), which a second tiny kernel applies on the device.
#include "ofk_common.cuh"

namespace ofk {

// "P" is the operand read at p (pointwise), "G" the operand gathered at p + sign*P[p].
struct SampleResult {
    float u, v;
    int strict;
};

// exact but slow: any coordinates, taps may leave the frame (out of line, by-value in / out: no stack traffic)
__device__ __noinline__ SampleResult sample_flow_border(const float2* __restrict__ G, const uint8_t* __restrict__ Gm,
                                                        int H, int W, float X, float Y) {
    const QCoord qx = quantise(X), qy = quantise(Y);
    const int ix = qx.i, iy = qy.i;
    const QWeights w = qweights(qx.f, qy.f);
    const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
    const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
    const long long o = (long long)iy * W + ix;
    const float2 z = make_float2(0.f, 0.f);
    const float2 t00 = (x0 && y0) ? __ldg(G + o) : z;
    const float2 t01 = (x1 && y0) ? __ldg(G + o + 1) : z;
    const float2 t10 = (x0 && y1) ? __ldg(G + o + W) : z;
    const float2 t11 = (x1 && y1) ? __ldg(G + o + W + 1) : z;
    int S = 0;
    if (x0 && y0 && (!Gm || __ldg(Gm + o))) S += w.w00;
    if (x1 && y0 && (!Gm || __ldg(Gm + o + 1))) S += w.w01;
    if (x0 && y1 && (!Gm || __ldg(Gm + o + W))) S += w.w10;
    if (x1 && y1 && (!Gm || __ldg(Gm + o + W + 1))) S += w.w11;
    const float s = 1.0f / 1024.0f;
    const float f00 = float(w.w00) * s, f01 = float(w.w01) * s, f10 = float(w.w10) * s, f11 = float(w.w11) * s;
    SampleResult r;
    r.u = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00.x, f00), __fmul_rn(t01.x, f01)), __fmul_rn(t10.x, f10)),
                    __fmul_rn(t11.x, f11));
    r.v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00.y, f00), __fmul_rn(t01.y, f01)), __fmul_rn(t10.y, f10)),
                    __fmul_rn(t11.y, f11));
    r.strict = (S == 1024);
    return r;
}

// cv2.remap float32 sample of the flow field G (and strict validity of its mask) at (X, Y). Interior pixels take the
// branch-free path: 4 unconditional 64-bit loads, weights from the 5-bit fractions, sums left to right without FMA
// contraction (bit-exact with OpenCV).
template <bool MASKS>
__device__ __forceinline__ SampleResult sample_flow(const float2* __restrict__ G, const uint8_t* __restrict__ Gm,
                                                    int H, int W, float X, float Y) {
    const QCoord qx = quantise_fast(X), qy = quantise_fast(Y);
    const bool interior = (unsigned)qx.i < (unsigned)(W - 1) && (unsigned)qy.i < (unsigned)(H - 1) &&
                          fabsf(X) < OFK_FAST_COORD_LIMIT && fabsf(Y) < OFK_FAST_COORD_LIMIT;
    if (!interior) return sample_flow_border(G, MASKS ? Gm : nullptr, H, W, X, Y);
    const unsigned o = (unsigned)(qy.i * W + qx.i);
    const float2* g0 = G + o;
    const float2* g1 = g0 + W;
    const float2 t00 = __ldg(g0), t01 = __ldg(g0 + 1), t10 = __ldg(g1), t11 = __ldg(g1 + 1);
    const int a = qx.f, b = qy.f;
    int strict = 1;
    if (MASKS) {
        // a tap only matters when its weight is non-zero: (32-a)(32-b), a(32-b), (32-a)b, ab
        const uint8_t* m0 = Gm + o;
        const uint8_t* m1 = m0 + W;
        const unsigned i00 = __ldg(m0) ^ 1u, i01 = __ldg(m0 + 1) ^ 1u, i10 = __ldg(m1) ^ 1u, i11 = __ldg(m1 + 1) ^ 1u;
        const unsigned ha = a != 0, hb = b != 0;
        strict = ((i00 | (i01 & ha) | (i10 & hb) | (i11 & ha & hb)) & 1u) ^ 1u;
    }
    const float fa = (float)a * (1.0f / 32.0f), fb = (float)b * (1.0f / 32.0f);
    const float na = 1.0f - fa, nb = 1.0f - fb;                      // exact: multiples of 1/32
    const float f00 = __fmul_rn(na, nb), f01 = __fmul_rn(fa, nb), f10 = __fmul_rn(na, fb), f11 = __fmul_rn(fa, fb);
    SampleResult r;
    r.u = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00.x, f00), __fmul_rn(t01.x, f01)), __fmul_rn(t10.x, f10)),
                    __fmul_rn(t11.x, f11));
    r.v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00.y, f00), __fmul_rn(t01.y, f01)), __fmul_rn(t10.y, f10)),
                    __fmul_rn(t11.y, f11));
    r.strict = strict;
    return r;
}

__device__ __forceinline__ bool nonzero(float c, float thr) { return thr > 0.f ? !(c < thr && c > -thr) : (c != 0.f); }

// flags[n*2 + 0] = A has a non-zero vector on a valid pixel, flags[n*2 + 1] = same for B. Warps record what they saw
// in shared memory; the last warp of the CTA to finish publishes (no CTA-wide barrier at the tail, so finished warps
// release their slots immediately).
struct FlagScratch {
    int nzA, nzB, done;
};
__device__ __forceinline__ void publish_flags(bool nzA, bool nzB, int* __restrict__ flags, int n, FlagScratch* sc) {
    const bool a = __any_sync(0xffffffffu, nzA), b = __any_sync(0xffffffffu, nzB);
    if ((threadIdx.x & 31) == 0) {
        if (a) sc->nzA = 1;   // benign race: every writer stores the same value
        if (b) sc->nzB = 1;
        __threadfence_block();
        if (atomicAdd(&sc->done, 1) == (int)(blockDim.x >> 5) - 1) {
            __threadfence_block();
            if (*(volatile int*)&sc->nzA) flags[n * 2 + 0] = 1;
            if (*(volatile int*)&sc->nzB) flags[n * 2 + 1] = 1;
        }
    }
}

// A warp owns 32 consecutive pixels of a row (lane = x) and walks 4 rows; a CTA owns a 32x32 tile. All accesses of a
// warp instruction are contiguous along x: the pointwise operand, the zero-test read of the gathered operand and the
// output are fully coalesced 256-byte requests, and each gather instruction touches the 2-3 cache lines of a rotated
// row segment. Loads use clamped indices (no branches), only the stores are predicated.
// REF_T: pointwise operand is B, gathered is A (sign -1); otherwise pointwise A, gathered B (sign +1).
template <bool REF_T, bool MASKS>
__global__ void __launch_bounds__(256) combine3_rows(const float* __restrict__ A, const uint8_t* __restrict__ Am,
                                                     const float* __restrict__ B, const uint8_t* __restrict__ Bm,
                                                     float thr, float* __restrict__ out, uint8_t* __restrict__ omask,
                                                     int* __restrict__ flags, int H, int W) {
    __shared__ FlagScratch sc;
    if (threadIdx.x == 0) sc.nzA = sc.nzB = sc.done = 0;
    __syncthreads();   // every warp is here at the start of the CTA anyway
    const unsigned lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const unsigned x = blockIdx.x * 32 + lane, xc = min(x, (unsigned)W - 1);
    const unsigned y0 = blockIdx.y * 32 + wrp * 4;
    const size_t fbase = (size_t)blockIdx.z * ((size_t)H * W);
    const float2* P = reinterpret_cast<const float2*>(REF_T ? B : A) + fbase;
    const float2* G = reinterpret_cast<const float2*>(REF_T ? A : B) + fbase;
    const uint8_t* Pm = MASKS ? (REF_T ? Bm : Am) + fbase : nullptr;
    const uint8_t* Gm = MASKS ? (REF_T ? Am : Bm) + fbase : nullptr;
    float2* O = reinterpret_cast<float2*>(out) + fbase;
    uint8_t* Om = omask + fbase;
    const float sign = REF_T ? -1.0f : 1.0f;
    const float Xg = static_cast<float>(x);
    // zero test as one compare per component: |c| >= t with t = thr, or the smallest denormal for the exact test
    const float tz = thr > 0.f ? thr : 1.401298464e-45f;
    float2 p[4], g[4];
    unsigned pmv[4], gmv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned idx = min(y0 + j, (unsigned)H - 1) * (unsigned)W + xc;
        p[j] = ld_stream_f2(P + idx);
        g[j] = __ldg(G + idx);            // zero test of the gathered operand AT p; these lines are gathered anyway
        pmv[j] = MASKS ? Pm[idx] : 1u;
        gmv[j] = MASKS ? __ldg(Gm + idx) : 1u;
    }
    unsigned nzP = 0, nzG = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned y = y0 + j;
        nzP |= pmv[j] & (unsigned)(fabsf(p[j].x) >= tz || fabsf(p[j].y) >= tz);
        nzG |= gmv[j] & (unsigned)(fabsf(g[j].x) >= tz || fabsf(g[j].y) >= tz);
        const SampleResult r = sample_flow<MASKS>(G, Gm, H, W, __fmaf_rn(sign, p[j].x, Xg),
                                                  __fmaf_rn(sign, p[j].y, static_cast<float>(y)));
        if (x < (unsigned)W && y < (unsigned)H) {
            const unsigned idx = y * (unsigned)W + x;
            float2 o;
            o.x = __fadd_rn(p[j].x, r.u);
            o.y = __fadd_rn(p[j].y, r.v);
            asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(O + idx), "f"(o.x), "f"(o.y) : "memory");
            Om[idx] = (uint8_t)(pmv[j] & (unsigned)r.strict);
        }
    }
    // clamped duplicates only repeat pixels of the same frame, so they cannot create false positives
    if (flags != nullptr) publish_flags(REF_T ? nzG : nzP, REF_T ? nzP : nzG, flags, blockIdx.z, &sc);
}

__device__ __forceinline__ void publish_flags_barrier(bool nzA, bool nzB, int* __restrict__ flags, int n) {
    const int a = __syncthreads_or(nzA);
    const int b = __syncthreads_or(nzB);
    if (threadIdx.x == 0) {
        if (a) flags[n * 2 + 0] = 1;
        if (b) flags[n * 2 + 1] = 1;
    }
}

// scalar fallback: any W, any alignment
template <bool REF_T>
__global__ void __launch_bounds__(256) combine3_scalar(const float* __restrict__ A, const uint8_t* __restrict__ Am,
                                                       const float* __restrict__ B, const uint8_t* __restrict__ Bm,
                                                       float thr, float* __restrict__ out,
                                                       uint8_t* __restrict__ omask, int* __restrict__ flags, int H,
                                                       int W) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int n = blockIdx.z;
    const size_t frame = (size_t)H * W;
    bool nzP = false, nzG = false;
    if (x < W && y < H) {
        const size_t pix = (size_t)n * frame + (size_t)y * W + x;
        const float* P = REF_T ? B : A;
        const uint8_t* Pm = REF_T ? Bm : Am;
        const float* G = REF_T ? A : B;
        const uint8_t* Gm = REF_T ? Am : Bm;
        const float sign = REF_T ? -1.0f : 1.0f;
        const float pu = P[pix * 2], pv = P[pix * 2 + 1];
        const float gu = G[pix * 2], gv = G[pix * 2 + 1];
        const bool pvalid = Pm ? Pm[pix] != 0 : true, gvalid = Gm ? Gm[pix] != 0 : true;
        nzP = pvalid && (nonzero(pu, thr) || nonzero(pv, thr));
        nzG = gvalid && (nonzero(gu, thr) || nonzero(gv, thr));
        const float X = __fadd_rn(sign * pu, static_cast<float>(x));
        const float Y = __fadd_rn(sign * pv, static_cast<float>(y));
        // float2 gathers need 8-byte alignment of the frame base; the scalar path cannot assume it
        const QCoord qx = quantise(X), qy = quantise(Y);
        const int ix = qx.i, iy = qy.i;
        const QWeights w = qweights(qx.f, qy.f);
        const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
        const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
        const float* Gf = G + (size_t)n * frame * 2;
        const uint8_t* Gmf = Gm ? Gm + (size_t)n * frame : nullptr;
        const long long o = (long long)iy * W + ix;
        const bool in[4] = {x0 && y0, x1 && y0, x0 && y1, x1 && y1};
        const long long off[4] = {o, o + 1, o + W, o + W + 1};
        const int wi[4] = {w.w00, w.w01, w.w10, w.w11};
        float au = 0.f, av = 0.f;
        int S = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float tu = in[k] ? Gf[off[k] * 2] : 0.f, tv = in[k] ? Gf[off[k] * 2 + 1] : 0.f;
            const float f = float(wi[k]) * (1.0f / 1024.0f);
            au = (k == 0) ? __fmul_rn(tu, f) : __fadd_rn(au, __fmul_rn(tu, f));
            av = (k == 0) ? __fmul_rn(tv, f) : __fadd_rn(av, __fmul_rn(tv, f));
            if (in[k] && (Gmf == nullptr || Gmf[off[k]])) S += wi[k];
        }
        out[pix * 2] = __fadd_rn(pu, au);
        out[pix * 2 + 1] = __fadd_rn(pv, av);
        omask[pix] = (pvalid && S == 1024) ? 1 : 0;
    }
    if (flags != nullptr) publish_flags_barrier(REF_T ? nzG : nzP, REF_T ? nzP : nzG, flags, n);
}

// Early exits of combine_with (flow_class.py:1338-1354), applied on the device: frame n becomes a copy of B when A
// is zero on its valid pixels, else a copy of A when B is. A NULL input mask means all valid (copied as ones).
__global__ void __launch_bounds__(256) combine3_fixup(const float* __restrict__ A, const uint8_t* __restrict__ Am,
                                                      const float* __restrict__ B, const uint8_t* __restrict__ Bm,
                                                      const int* __restrict__ flags, float* __restrict__ out,
                                                      uint8_t* __restrict__ omask, size_t frame) {
    const int n = blockIdx.y;
    const int a_nz = flags[n * 2], b_nz = flags[n * 2 + 1];
    if (a_nz && b_nz) return;
    const float* src = a_nz ? A : B;  // A zero -> B wins (tested first, like the reference); else B zero -> A
    const uint8_t* srcm = a_nz ? Am : Bm;
    const size_t base = (size_t)n * frame;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < frame; i += (size_t)gridDim.x * blockDim.x) {
        out[(base + i) * 2] = src[(base + i) * 2];
        out[(base + i) * 2 + 1] = src[(base + i) * 2 + 1];
        omask[base + i] = srcm ? srcm[base + i] : 1;
    }
}

}  // namespace ofk

using namespace ofk;

extern "C" int ofk_combine3(const float* A, const uint8_t* Am, const float* B, const uint8_t* Bm, int ref, float thr,
                            float* out, uint8_t* out_mask, int* flags, int N, int H, int W, ofk_stream_t stream) {
    OFK_CHECK_ARG(A && B && out && out_mask, "ofk_combine3: NULL operand (A=%p B=%p out=%p out_mask=%p)", (void*)A,
                  (void*)B, (void*)out, (void*)out_mask);
    OFK_CHECK_ARG(ref == 's' || ref == 't', "ofk_combine3: ref must be 's' or 't', got %d", ref);
    OFK_CHECK_ARG(N >= 0 && H > 0 && W > 0, "ofk_combine3: bad shape N=%d H=%d W=%d", N, H, W);
    OFK_CHECK_ARG(thr >= 0.f, "ofk_combine3: negative threshold");
    if (N == 0) return OFK_OK;
    OFK_CHECK_ARG(N <= 65535, "ofk_combine3: N=%d exceeds 65535 frames per call", N);
    cudaStream_t st = as_stream(stream);
    if (flags != nullptr) OFK_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * 2 * (size_t)N, st));
    const bool fast = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
                        reinterpret_cast<uintptr_t>(out)) & 7) == 0 && (size_t)H * W < ((size_t)1 << 30) && H < 32768 &&
                      W < 32768 && ((Am == nullptr) == (Bm == nullptr));
    if (fast) {
        dim3 grid((W + 31) / 32, (H + 31) / 32, N);
        if (Am && Bm) {
            if (ref == 't') combine3_rows<true, true><<<grid, 256, 0, st>>>(A, Am, B, Bm, thr, out, out_mask, flags, H, W);
            else combine3_rows<false, true><<<grid, 256, 0, st>>>(A, Am, B, Bm, thr, out, out_mask, flags, H, W);
        } else {
            if (ref == 't') combine3_rows<true, false><<<grid, 256, 0, st>>>(A, Am, B, Bm, thr, out, out_mask, flags, H, W);
            else combine3_rows<false, false><<<grid, 256, 0, st>>>(A, Am, B, Bm, thr, out, out_mask, flags, H, W);
        }
    } else {
        dim3 grid((W + 31) / 32, (H + 7) / 8, N);
        if (ref == 't') combine3_scalar<true><<<grid, 256, 0, st>>>(A, Am, B, Bm, thr, out, out_mask, flags, H, W);
        else combine3_scalar<false><<<grid, 256, 0, st>>>(A, Am, B, Bm, thr, out, out_mask, flags, H, W);
    }
    OFK_LAUNCHED();
    if (flags != nullptr) {
        const size_t frame = (size_t)H * W;
        int bx = (int)((frame + 256 * 8 - 1) / (256 * 8));
        if (bx > 64) bx = 64;
        combine3_fixup<<<dim3(bx, N), 256, 0, st>>>(A, Am, B, Bm, flags, out, out_mask, frame);
        OFK_LAUNCHED();
    }
    return OFK_OK;
}
