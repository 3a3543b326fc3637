// Target-referenced (backward) warp for sm_100a: bilinear gather with OpenCV's 1/32-px fixed-point coordinates,
// payload and validity mask resampled in the same pass. Replaces cv2.remap at utils.py:236 of the reference plus the
// mask plumbing of Flow.apply (flow_class.py:631-680). HBM-bound gather: no tensor cores.
//
// Two kernels:
//   warp_t_vec4    4 output pixels per thread along x, 128-bit flow loads, packed stores, 2-D CTA tiles so the
//                  gathered neighbourhood of a tile stays in L1; interior pixels of uint8x3 images read both
//                  horizontal taps of a row with one (two when straddling) aligned 64-bit load.
//   warp_t_generic 1 pixel per thread, any channel count / dtype / padding offsets / unaligned pointers.
#include "ofk_common.cuh"

namespace ofk {

enum : int { AR_U8_FIXED = 0, AR_RINT = 1, AR_F32 = 2, AR_F64 = 3 };

template <typename T>
__device__ __forceinline__ T saturate_rint(float v);
template <>
__device__ __forceinline__ uint8_t saturate_rint<uint8_t>(float v) {
    return static_cast<uint8_t>(max(0, min(255, __float2int_rn(v))));
}
template <>
__device__ __forceinline__ int16_t saturate_rint<int16_t>(float v) {
    return static_cast<int16_t>(max(-32768, min(32767, __float2int_rn(v))));
}
template <>
__device__ __forceinline__ uint16_t saturate_rint<uint16_t>(float v) {
    return static_cast<uint16_t>(max(0, min(65535, __float2int_rn(v))));
}
template <>
__device__ __forceinline__ float saturate_rint<float>(float v) {
    return v;
}
template <>
__device__ __forceinline__ double saturate_rint<double>(float v) {
    return v;
}

// One interpolated value from four taps, in OpenCV's arithmetic for the given mode (sums left to right, no FMA
// contraction: cv2.remap's results are reproduced bit for bit, see tests/test_oracle_remap.py for the CPU statement).
template <typename T, int AR>
__device__ __forceinline__ T blend(T t00, T t01, T t10, T t11, const QWeights& w) {
    if (AR == AR_U8_FIXED) {
        // (sum t*w*32 + 2^14) >> 15 with 15-bit weights w*32  ==  (sum t*w + 512) >> 10
        int acc = int(t00) * w.w00 + int(t01) * w.w01 + int(t10) * w.w10 + int(t11) * w.w11;
        return static_cast<T>((acc + 512) >> 10);
    } else if (AR == AR_F64) {
        const double s = 1.0 / 1024.0;
        double acc = __dmul_rn(double(t00), double(w.w00) * s);
        acc = __dadd_rn(acc, __dmul_rn(double(t01), double(w.w01) * s));
        acc = __dadd_rn(acc, __dmul_rn(double(t10), double(w.w10) * s));
        acc = __dadd_rn(acc, __dmul_rn(double(t11), double(w.w11) * s));
        return static_cast<T>(acc);
    } else {
        const float s = 1.0f / 1024.0f;  // weights are exact multiples of 2^-10
        float acc = __fmul_rn(float(t00), float(w.w00) * s);
        acc = __fadd_rn(acc, __fmul_rn(float(t01), float(w.w01) * s));
        acc = __fadd_rn(acc, __fmul_rn(float(t10), float(w.w10) * s));
        acc = __fadd_rn(acc, __fmul_rn(float(t11), float(w.w11) * s));
        if (AR == AR_RINT) return saturate_rint<T>(acc);
        return static_cast<T>(acc);
    }
}

struct Taps {
    int ix, iy;
    QWeights w;
    bool in00, in01, in10, in11;
    bool interior;
};

__device__ __forceinline__ Taps make_taps(float X, float Y, int Hs, int Ws) {
    QCoord qx = quantise(X), qy = quantise(Y);
    Taps t;
    t.ix = qx.i;
    t.iy = qy.i;
    t.w = qweights(qx.f, qy.f);
    bool x0 = (unsigned)t.ix < (unsigned)Ws, x1 = (unsigned)(t.ix + 1) < (unsigned)Ws;
    bool y0 = (unsigned)t.iy < (unsigned)Hs, y1 = (unsigned)(t.iy + 1) < (unsigned)Hs;
    t.in00 = x0 && y0;
    t.in01 = x1 && y0;
    t.in10 = x0 && y1;
    t.in11 = x1 && y1;
    t.interior = t.in00 && t.in11;
    return t;
}

// valid-weight sum in 1/1024 units; pm == nullptr means "payload everywhere valid" (only bounds matter)
__device__ __forceinline__ int valid_weight_sum(const Taps& t, const uint8_t* __restrict__ pm, int Ws) {
    int S = 0;
    if (pm == nullptr) {
        S = (t.in00 ? t.w.w00 : 0) + (t.in01 ? t.w.w01 : 0) + (t.in10 ? t.w.w10 : 0) + (t.in11 ? t.w.w11 : 0);
    } else {
        const long long o = (long long)t.iy * Ws + t.ix;
        if (t.in00 && pm[o]) S += t.w.w00;
        if (t.in01 && pm[o + 1]) S += t.w.w01;
        if (t.in10 && pm[o + Ws]) S += t.w.w10;
        if (t.in11 && pm[o + Ws + 1]) S += t.w.w11;
    }
    return S;
}

// ----------------------------------------------------------------------------------------------- generic kernel
template <typename T, int AR>
__global__ void __launch_bounds__(256) warp_t_generic(const T* __restrict__ payload, int C,
                                                      const float* __restrict__ flow, float sign,
                                                      const uint8_t* __restrict__ pmask,
                                                      const uint8_t* __restrict__ fmask, T* __restrict__ out,
                                                      uint8_t* __restrict__ omask, int rule, int H, int W, int Hs,
                                                      int Ws, int Ho, int Wo, int oy, int ox, int fy, int fx) {
    // output frame (Ho, Wo); output pixel (y, x) sits at (y + oy, x + ox) of the payload frame and at
    // (y - fy, x - fx) of the flow frame (zero flow / invalid outside it).
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int n = blockIdx.z;
    if (x >= Wo || y >= Ho) return;
    const int yf = y - fy, xf = x - fx;
    const bool in_flow = (unsigned)yf < (unsigned)H && (unsigned)xf < (unsigned)W;
    float u = 0.f, v = 0.f;
    if (in_flow) {
        const float* f = flow + (((size_t)n * H + yf) * W + xf) * 2;
        u = f[0];
        v = f[1];
    }
    const float X = sample_coord(u, sign, x + ox);
    const float Y = sample_coord(v, sign, y + oy);
    const Taps t = make_taps(X, Y, Hs, Ws);
    const size_t opix = ((size_t)n * Ho + y) * Wo + x;
    if (omask != nullptr) {
        const uint8_t* pm = pmask ? pmask + (size_t)n * Hs * Ws : nullptr;
        bool ok = mask_rule_pass(valid_weight_sum(t, pm, Ws), rule);
        if (!in_flow) ok = false;
        else if (fmask) ok = ok && fmask[((size_t)n * H + yf) * W + xf];
        omask[opix] = ok ? 1 : 0;
    }
    if (C > 0) {
        const T* p = payload + (size_t)n * Hs * Ws * C;
        const long long o = ((long long)t.iy * Ws + t.ix) * C;
        const long long row = (long long)Ws * C;
        for (int c = 0; c < C; ++c) {
            T t00 = t.in00 ? p[o + c] : T(0);
            T t01 = t.in01 ? p[o + C + c] : T(0);
            T t10 = t.in10 ? p[o + row + c] : T(0);
            T t11 = t.in11 ? p[o + row + C + c] : T(0);
            out[opix * C + c] = blend<T, AR>(t00, t01, t10, t11, t.w);
        }
    }
}

// ----------------------------------------------------------------------------------------------- vec4 kernel
// Per-(T,C) tap fetch for one pixel: fills v[4][C] (tap order 00,01,10,11) with zero for out-of-bounds taps.
template <typename T, int C>
struct Fetch {
    __device__ __forceinline__ static void run(const T* __restrict__ p, int Ws, const Taps& t, T (&v)[4][C],
                                               const void* /*buf_end*/) {
        const long long o = ((long long)t.iy * Ws + t.ix) * C;
        const long long row = (long long)Ws * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            v[0][c] = t.in00 ? p[o + c] : T(0);
            v[1][c] = t.in01 ? p[o + C + c] : T(0);
            v[2][c] = t.in10 ? p[o + row + c] : T(0);
            v[3][c] = t.in11 ? p[o + row + C + c] : T(0);
        }
    }
};

// float32 x2 (flow fields): one 64-bit load per tap
template <>
struct Fetch<float, 2> {
    __device__ __forceinline__ static void run(const float* __restrict__ p, int Ws, const Taps& t, float (&v)[4][2],
                                               const void*) {
        const float2* q = reinterpret_cast<const float2*>(p);
        const long long o = (long long)t.iy * Ws + t.ix;
        const float2 z = make_float2(0.f, 0.f);
        float2 a = t.in00 ? __ldg(q + o) : z;
        float2 b = t.in01 ? __ldg(q + o + 1) : z;
        float2 c = t.in10 ? __ldg(q + o + Ws) : z;
        float2 d = t.in11 ? __ldg(q + o + Ws + 1) : z;
        v[0][0] = a.x; v[0][1] = a.y;
        v[1][0] = b.x; v[1][1] = b.y;
        v[2][0] = c.x; v[2][1] = c.y;
        v[3][0] = d.x; v[3][1] = d.y;
    }
};

// uint8 x4: one 32-bit load per tap
template <>
struct Fetch<uint8_t, 4> {
    __device__ __forceinline__ static void run(const uint8_t* __restrict__ p, int Ws, const Taps& t,
                                               uint8_t (&v)[4][4], const void*) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
        const long long o = (long long)t.iy * Ws + t.ix;
        uint32_t w[4];
        w[0] = t.in00 ? __ldg(q + o) : 0u;
        w[1] = t.in01 ? __ldg(q + o + 1) : 0u;
        w[2] = t.in10 ? __ldg(q + o + Ws) : 0u;
        w[3] = t.in11 ? __ldg(q + o + Ws + 1) : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k][0] = w[k] & 0xff;
            v[k][1] = (w[k] >> 8) & 0xff;
            v[k][2] = (w[k] >> 16) & 0xff;
            v[k][3] = w[k] >> 24;
        }
    }
};

// 6-byte window [b0..b5] starting at address a: the two horizontal taps of a uint8x3 row.
// One aligned 64-bit load, a second one only when the window straddles the 8-byte boundary.
__device__ __forceinline__ uint64_t window6(const uint8_t* __restrict__ a) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(a);
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(addr & ~uintptr_t(7));
    const unsigned sh = (unsigned)(addr & 7) * 8;
    uint64_t lo = __ldg(q);
    if (sh > 16) {  // bytes 0..5 of the window cross into the next word
        uint64_t hi = __ldg(q + 1);
        lo = (lo >> sh) | (hi << (64 - sh));
    } else {
        lo >>= sh;
    }
    return lo;
}

template <>
struct Fetch<uint8_t, 3> {
    __device__ __forceinline__ static void run(const uint8_t* __restrict__ p, int Ws, const Taps& t,
                                               uint8_t (&v)[4][3], const void* buf_end) {
        const long long o = ((long long)t.iy * Ws + t.ix) * 3;
        const long long row = (long long)Ws * 3;
        // fast path: all four taps inside, and the aligned word covering the end of the lower window inside the buffer
        if (t.interior &&
            ((reinterpret_cast<uintptr_t>(p + o + row + 6) + 7) & ~uintptr_t(7)) <= reinterpret_cast<uintptr_t>(buf_end)) {
            uint64_t r0 = window6(p + o);
            uint64_t r1 = window6(p + o + row);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[0][c] = (uint8_t)(r0 >> (8 * c));
                v[1][c] = (uint8_t)(r0 >> (8 * (c + 3)));
                v[2][c] = (uint8_t)(r1 >> (8 * c));
                v[3][c] = (uint8_t)(r1 >> (8 * (c + 3)));
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[0][c] = t.in00 ? p[o + c] : 0;
                v[1][c] = t.in01 ? p[o + 3 + c] : 0;
                v[2][c] = t.in10 ? p[o + row + c] : 0;
                v[3][c] = t.in11 ? p[o + row + 3 + c] : 0;
            }
        }
    }
};

// Packed store of 4 pixels x C elements of T (contiguous, 4*C*sizeof(T) bytes, a multiple of 4 bytes; the address is
// 4-byte aligned because W % 4 == 0 and x0 % 4 == 0).
template <typename T, int C>
__device__ __forceinline__ void store4(T* __restrict__ dst, const T (&r)[4][C]) {
    constexpr int BYTES = 4 * C * sizeof(T);
    union {
        T e[4 * C];
        uint32_t w[BYTES / 4];
        uint4 q[(BYTES + 15) / 16];
    } u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < C; ++c) u.e[j * C + c] = r[j][c];
    if constexpr (BYTES % 16 == 0) {
        // 16-byte aligned when the per-4-pixel footprint is a multiple of 16 bytes
        uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int k = 0; k < BYTES / 16; ++k) d[k] = u.q[k];
    } else {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
        for (int k = 0; k < BYTES / 4; ++k) d[k] = u.w[k];
    }
}

// TXT x-threads per tile row (tile width = 4*TXT pixels), 256/TXT rows per tile.
template <typename T, int C, int AR, int TXT>
__global__ void __launch_bounds__(256) warp_t_vec4(const T* __restrict__ payload, const float* __restrict__ flow,
                                                   float sign, const uint8_t* __restrict__ pmask,
                                                   const uint8_t* __restrict__ fmask, T* __restrict__ out,
                                                   uint8_t* __restrict__ omask, int rule, int H, int W) {
    constexpr int ROWS = 256 / TXT;
    const int tx = threadIdx.x % TXT, ty = threadIdx.x / TXT;
    const int x0 = (blockIdx.x * TXT + tx) * 4;
    const int y = blockIdx.y * ROWS + ty;
    const int n = blockIdx.z;
    if (x0 >= W || y >= H) return;
    const size_t pix0 = ((size_t)n * H + y) * W + x0;

    const float4* f4 = reinterpret_cast<const float4*>(flow + pix0 * 2);
    const float4 fa = ld_stream_f4(f4), fb = ld_stream_f4(f4 + 1);
    const float u[4] = {fa.x, fa.z, fb.x, fb.z};
    const float v[4] = {fa.y, fa.w, fb.y, fb.w};
    uint32_t fm = 0x01010101u;
    if (omask != nullptr && fmask != nullptr) fm = ld_stream_u32(reinterpret_cast<const uint32_t*>(fmask + pix0));

    const size_t frame_elems = (size_t)H * W * C;
    const T* p = payload + (size_t)n * frame_elems;
    const uint8_t* pm = pmask ? pmask + (size_t)n * H * W : nullptr;
    const float Y0 = static_cast<float>(y);

    const void* buf_end = payload + (size_t)gridDim.z * frame_elems;
    T res[4][C > 0 ? C : 1];
    uint32_t om = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float X = __fadd_rn(sign * u[j], static_cast<float>(x0 + j));
        const float Y = __fadd_rn(sign * v[j], Y0);
        const Taps t = make_taps(X, Y, H, W);
        if constexpr (C > 0) {
            T taps[4][C];
            Fetch<T, C>::run(p, W, t, taps, buf_end);
#pragma unroll
            for (int c = 0; c < C; ++c) res[j][c] = blend<T, AR>(taps[0][c], taps[1][c], taps[2][c], taps[3][c], t.w);
        }
        if (omask != nullptr) {
            const bool ok = mask_rule_pass(valid_weight_sum(t, pm, W), rule) && ((fm >> (8 * j)) & 1u);
            om |= (ok ? 1u : 0u) << (8 * j);
        }
    }
    if constexpr (C > 0) store4<T, C>(out + pix0 * C, res);
    if (omask != nullptr) *reinterpret_cast<uint32_t*>(omask + pix0) = om;
}

template <typename T, int C, int AR>
static int launch_vec4(const void* payload, const float* flow, float sign, const uint8_t* pmask, const uint8_t* fmask,
                       void* out, uint8_t* omask, int rule, int N, int H, int W, cudaStream_t st) {
    constexpr int TXT = 8;  // 32 x 32 pixel tiles
    dim3 grid((W / 4 + TXT - 1) / TXT, (H + (256 / TXT) - 1) / (256 / TXT), N);
    warp_t_vec4<T, C, AR, TXT><<<grid, 256, 0, st>>>(static_cast<const T*>(payload), flow, sign, pmask, fmask,
                                                     static_cast<T*>(out), omask, rule, H, W);
    OFK_LAUNCHED();
    return OFK_OK;
}

template <typename T, int AR>
static int launch_generic(const void* payload, int C, const float* flow, float sign, const uint8_t* pmask,
                          const uint8_t* fmask, void* out, uint8_t* omask, int rule, int N, int H, int W, int Hs,
                          int Ws, int Ho, int Wo, int oy, int ox, int fy, int fx, cudaStream_t st) {
    dim3 grid((Wo + 31) / 32, (Ho + 7) / 8, N);
    warp_t_generic<T, AR><<<grid, 256, 0, st>>>(static_cast<const T*>(payload), C, flow, sign, pmask, fmask,
                                                static_cast<T*>(out), omask, rule, H, W, Hs, Ws, Ho, Wo, oy, ox, fy,
                                                fx);
    OFK_LAUNCHED();
    return OFK_OK;
}

}  // namespace ofk

using namespace ofk;

extern "C" int ofk_warp_t(const void* payload, int dtype, int C, int arith, const float* flow, float flow_sign,
                          const uint8_t* payload_mask, const uint8_t* flow_mask, void* out, uint8_t* out_mask,
                          int mask_rule, int N, int H, int W, int Hs, int Ws, int top, int left, int cut,
                          ofk_stream_t stream) {
    OFK_CHECK_ARG(flow != nullptr, "ofk_warp_t: flow is NULL");
    OFK_CHECK_ARG(N >= 0 && H > 0 && W > 0 && Hs > 0 && Ws > 0, "ofk_warp_t: bad shape N=%d H=%d W=%d Hs=%d Ws=%d", N,
                  H, W, Hs, Ws);
    OFK_CHECK_ARG(C >= 0 && (C == 0 || (payload != nullptr && out != nullptr)),
                  "ofk_warp_t: payload/out NULL with C=%d", C);
    OFK_CHECK_ARG(C > 0 || out_mask != nullptr, "ofk_warp_t: nothing to compute (C=0 and out_mask NULL)");
    OFK_CHECK_ARG(dtype >= OFK_U8 && dtype <= OFK_F64, "ofk_warp_t: unknown dtype %d", dtype);
    OFK_CHECK_ARG(mask_rule >= OFK_RULE_STRICT && mask_rule <= OFK_RULE_GE_HALF, "ofk_warp_t: unknown mask rule %d",
                  mask_rule);
    OFK_CHECK_ARG(top >= 0 && left >= 0 && top + H <= Hs && left + W <= Ws,
                  "ofk_warp_t: flow frame (%d,%d)+(%d,%d) does not fit the payload frame (%d,%d)", top, left, H, W, Hs,
                  Ws);
    OFK_CHECK_ARG(flow_sign == 1.0f || flow_sign == -1.0f, "ofk_warp_t: flow_sign must be +1 or -1");
    if (N == 0) return OFK_OK;
    OFK_CHECK_ARG(N <= 65535, "ofk_warp_t: N=%d exceeds 65535 frames per call", N);
    cudaStream_t st = as_stream(stream);

    const bool same_frame = (Hs == H && Ws == W);
    int ar;
    switch (dtype) {
        case OFK_U8: ar = (arith == OFK_ARITH_RINT) ? AR_RINT : AR_U8_FIXED; break;
        case OFK_I16:
        case OFK_U16: ar = AR_RINT; break;
        case OFK_F32: ar = AR_F32; break;
        default: ar = AR_F64; break;
    }
    const bool fast = same_frame && (W % 4 == 0) && aligned16(flow) && (C == 0 || (aligned16(payload) && aligned16(out))) &&
                      (payload_mask == nullptr || aligned16(payload_mask)) &&
                      (flow_mask == nullptr || aligned16(flow_mask)) && (out_mask == nullptr || aligned16(out_mask));
    if (fast) {
#define OFK_V4(T, CC, AR)                                                                                         \
    return launch_vec4<T, CC, AR>(payload, flow, flow_sign, payload_mask, flow_mask, out, out_mask, mask_rule, N, H, \
                                  W, st)
        if (C == 0) OFK_V4(uint8_t, 0, AR_U8_FIXED);
        if (dtype == OFK_U8 && ar == AR_U8_FIXED) {
            if (C == 3) OFK_V4(uint8_t, 3, AR_U8_FIXED);
            if (C == 1) OFK_V4(uint8_t, 1, AR_U8_FIXED);
            if (C == 4) OFK_V4(uint8_t, 4, AR_U8_FIXED);
        } else if (dtype == OFK_U8 && ar == AR_RINT) {
            if (C == 3) OFK_V4(uint8_t, 3, AR_RINT);
            if (C == 1) OFK_V4(uint8_t, 1, AR_RINT);
            if (C == 4) OFK_V4(uint8_t, 4, AR_RINT);
        } else if (dtype == OFK_F32) {
            if (C == 2) OFK_V4(float, 2, AR_F32);
            if (C == 3) OFK_V4(float, 3, AR_F32);
            if (C == 1) OFK_V4(float, 1, AR_F32);
            if (C == 4) OFK_V4(float, 4, AR_F32);
        }
#undef OFK_V4
    }
    // generic path: output frame is the flow frame (cut) or the payload frame (no cut)
    const int Ho = cut ? H : Hs, Wo = cut ? W : Ws;
    const int oy = cut ? top : 0, ox = cut ? left : 0;
    const int fy = cut ? 0 : top, fx = cut ? 0 : left;
#define OFK_GEN(T, AR)                                                                                              \
    return launch_generic<T, AR>(payload, C, flow, flow_sign, payload_mask, flow_mask, out, out_mask, mask_rule, N, H, \
                                 W, Hs, Ws, Ho, Wo, oy, ox, fy, fx, st)
    switch (dtype) {
        case OFK_U8:
            if (ar == AR_RINT) OFK_GEN(uint8_t, AR_RINT);
            OFK_GEN(uint8_t, AR_U8_FIXED);
        case OFK_I16: OFK_GEN(int16_t, AR_RINT);
        case OFK_U16: OFK_GEN(uint16_t, AR_RINT);
        case OFK_F32: OFK_GEN(float, AR_F32);
        default: OFK_GEN(double, AR_F64);
    }
#undef OFK_GEN
}

extern "C" int ofk_valid_geom_t(const float* flow, float flow_sign, const uint8_t* flow_mask, uint8_t* out, int N,
                                int H, int W, ofk_stream_t stream) {
    OFK_CHECK_ARG(out != nullptr, "ofk_valid_geom_t: out is NULL");
    return ofk_warp_t(nullptr, OFK_U8, 0, OFK_ARITH_NATIVE, flow, flow_sign, nullptr, flow_mask, nullptr, out,
                      OFK_RULE_STRICT, N, H, W, H, W, 0, 0, 1, stream);
}
