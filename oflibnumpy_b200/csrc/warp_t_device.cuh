// Device helpers shared by the backward-warp kernels (warp_t.cu, warp_t_ws.cu). Static: every translation unit gets
// its own copy (the library is built without relocatable device code).
#pragma once
#include "ofk_common.cuh"

namespace ofk {

// Pixels whose taps touch the border (or whose coordinates leave the fast-quantisation range): exact but slow, out
// of line, by-value in / out (no stack traffic). C interleaved uint8 channels (C <= 4): returns the C result bytes in
// bits 0..8C-1 and the validity in bit 32.
template <int C>
static __device__ __noinline__ unsigned long long border_px_u8q(const uint8_t* __restrict__ p,
                                                                const uint8_t* __restrict__ pm, int ix, int iy, int a,
                                                                int b, int H, int W, int half_even, int rule) {
    const QWeights w = qweights(a, b);
    const bool in[4] = {(unsigned)ix < (unsigned)W && (unsigned)iy < (unsigned)H,
                        (unsigned)(ix + 1) < (unsigned)W && (unsigned)iy < (unsigned)H,
                        (unsigned)ix < (unsigned)W && (unsigned)(iy + 1) < (unsigned)H,
                        (unsigned)(ix + 1) < (unsigned)W && (unsigned)(iy + 1) < (unsigned)H};
    const long long o = (long long)iy * W + ix;
    const long long off[4] = {o, o + 1, o + W, o + W + 1};
    const int wi[4] = {w.w00, w.w01, w.w10, w.w11};
    int acc[C], S = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!in[k]) continue;
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += (int)p[off[k] * C + c] * wi[k];
        if (pm == nullptr || pm[off[k]]) S += wi[k];
    }
    unsigned long long v = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int a = acc[c];
        const uint32_t r = half_even ? (uint32_t)(a + 511 + ((a >> 10) & 1)) >> 10 : (uint32_t)(a + 512) >> 10;
        v |= (unsigned long long)r << (8 * c);
    }
    return v | (mask_rule_pass(S, rule) ? (1ull << 32) : 0ull);
}
// the same from float coordinates (exact quantiser, any value)
template <int C>
static __device__ __forceinline__ unsigned long long border_px_u8c(const uint8_t* __restrict__ p,
                                                                   const uint8_t* __restrict__ pm, float X, float Y,
                                                                   int H, int W, int half_even, int rule) {
    const QCoord qx = quantise(X), qy = quantise(Y);
    return border_px_u8q<C>(p, pm, qx.i, qy.i, qx.f, qy.f, H, W, half_even, rule);
}

// 3 channels, packed for warp_t.cu: result bytes in bits 0..23, validity in bit 24
static __device__ __forceinline__ uint32_t border_px_u8x3(const uint8_t* __restrict__ p, const uint8_t* __restrict__ pm,
                                                         float X, float Y, int H, int W, int half_even, int rule) {
    const unsigned long long r = border_px_u8c<3>(p, pm, X, Y, H, W, half_even, rule);
    return (uint32_t)r | ((uint32_t)(r >> 32) << 24);
}

}  // namespace ofk
