// Target-referenced (backward) warp of uint8 x3 images for sm_100a: warp-specialised, TMA-staged kernel (the default
// path of ofk_warp_t for images). Replaces cv2.remap at utils.py:236 of the reference plus the mask plumbing of
// Flow.apply (flow_class.py:631-680).
//
// Same pipeline as combine3_ws.cu: CTA = 1 producer warp + 8 consumer warps, persistent over 32x32 output tiles.
//   producer : TMA-loads the flow tile (and the flow-mask tile) ahead of time, estimates the bounding box of the
//              tile's sample positions from the tile perimeter and TMA-loads that box of the image -- as rows of
//              3*W bytes, 192 bytes x 48 rows, start aligned down to 16 bytes -- and, if a target mask is resampled, the
//              matching box of the mask. Hardware zero fill outside the frame == cv2.remap's BORDER_CONSTANT(0).
//   consumers: both horizontal taps of a row (6 bytes at an arbitrary byte offset) come out of three aligned shared-
//              memory words; the blend is pure integer (cv2.remap's uint8 fixed-point path and its int16 path are both
//              exact integers for 8-bit taps, see warp_t.cu): two 16-bit x 8-bit dot products (dp2a) per channel against
//              the packed weights {(32-a)(32-b), a(32-b)} and {(32-a)b, ab}. The three result bytes and the validity
//              byte are written over the flow tile in place; the warp's four rows leave through TMA stores (clipped at
//              the frame border by the hardware).
// Pixels whose taps are not covered by the box (discontinuous or noisy flows) read the same words from global memory.
// HBM traffic is the algorithmic 8 (flow) + 3 + 3 (image in / out) + 1 + 1 (flow mask, validity) bytes per pixel.
#include <type_traits>
#include <stdlib.h>

#include "warp_t_device.cuh"
#include "ws_common.cuh"

namespace ofk {
namespace wtws {
using namespace ws;

constexpr int TS = 32;
constexpr int BH = 48;            // box rows
constexpr int NCW = 8;            // consumer warps
#ifndef OFK_BOX3
#define OFK_BOX3 192
#endif
// image box width in bytes for C interleaved uint8 channels: 48 pixels x C + up to 15 bytes of alignment slack, rounded
// up to the 16 bytes TMA needs (C = 1: 64, C = 4: 208). C = 3 would be 160 bytes = 40 words per box row: under a rotation
// the 32 pixels of a warp row sample 2-3 consecutive box rows, each segment 8 words further along, and with a row pitch
// of 8 banks (mod 32) the first and third segment land on the same banks (either sense of rotation: 8k +- 8k) -- measured
// 2.15 wavefronts per tap load. 192 bytes = 48 words = 16 banks (mod 32) keeps the segments apart: 24k, 8k.
__host__ __device__ constexpr int box_bytes(int C) { return C == 3 ? OFK_BOX3 : (48 * C + 15 + 15) / 16 * 16; }
// pixels of a box row that are usable whatever the 16-byte alignment of its start
__host__ __device__ constexpr int box_px(int C) { return (box_bytes(C) - 15) / C; }
// mask box width (bytes = pixels): its start is the first needed pixel aligned down to 16, the image box reaches at most
// (box_bytes - 2C) / C + 15 <= 77 pixels (+ 1 tap) beyond that
#ifndef OFK_BMW
#define OFK_BMW 80
#endif
constexpr int BMW = OFK_BMW;
// BH rows, rounded up to 128 with room for the word loads that run a few bytes past the last tap
__host__ __device__ constexpr int img_stage(int C) { return (BH * box_bytes(C) + 16 + 127) / 128 * 128; }

enum : int { MM_NONE = 0, MM_GEOM = 1, MM_PMASK = 2 };

template <int NP, int NB, bool PMBOX, int C>
struct Smem {
    struct PStage {
        float2 f[TS * TS];        // flow tile; rows [4w, 4w+4) double as warp w's output staging (4 x 32*C bytes)
        uint8_t fm[TS * TS];      // flow-mask tile; overwritten by the validity bytes
    };
    struct BStage {
        uint8_t img[img_stage(C)];
        uint8_t m[PMBOX ? BH * BMW : 128];
    };
    alignas(128) PStage ps[NP];
    alignas(128) BStage bs[NB];
    alignas(16) int4 binfo[NB][2];   // {bx0 (bytes), mx0 (pixels), by0, -}, {tx0, ty0, n, -}
    uint64_t pfull[NP], pempty[NP], bfull[NB], bempty[NB];
    uint32_t sink[NCW];              // scratch words of mbar_arrive_after, one per consumer warp
};

struct Maps {
    CUtensorMap f, fm, ib, pmb, oi, om;
};

// Rounding of acc = sum(t * w) + RBIAS to a multiple of 1024. Round half up: RBIAS = 512, acc >> 10. cvRound (round half
// to even) without a tie test: RBIAS = 511 and the bit above the fraction is added back before the shift -- a carry out
// of the low 10 bits then needs either a fraction > 1/2 (the added bit makes no difference) or an exact tie with an odd
// integer part, which is the one case that has to round up.
template <bool HALF_EVEN>
struct Rnd {
    static constexpr uint32_t BIAS = HALF_EVEN ? 511u : 512u;
    static __device__ __forceinline__ uint32_t fin(uint32_t acc) {
        // (acc & 1024) >> 10 written as a high multiply: LOP3 + LEA.HI + SHF (the plain form costs a fourth instruction)
        return HALF_EVEN ? (acc + __umulhi(acc & 1024u, 1u << 22)) >> 10 : acc >> 10;
    }
};
__device__ __forceinline__ void weights_u8(unsigned a, unsigned b, uint32_t& W0, uint32_t& W1) {
    const uint32_t pa = a * 65535u + 32u;              // (32 - a) | a << 16
    W1 = pa * b;                                       // {(32-a)b, ab}
    W0 = pa * 32u - W1;                                // {(32-a)(32-b), a(32-b)}: exact, no borrow between the halves
}
// the three result bytes of one pixel (v0, v1, v2: values 0..255) from the two 12-byte windows around its taps
template <bool HALF_EVEN>
__device__ __forceinline__ void blend_u8x3(uint32_t r0w0, uint32_t r0w1, uint32_t r0w2, uint32_t r1w0, uint32_t r1w1,
                                           uint32_t r1w2, unsigned sh8, unsigned a, unsigned b, uint32_t& W0,
                                           uint32_t& W1, uint32_t& v0, uint32_t& v1, uint32_t& v2) {
    using R = Rnd<HALF_EVEN>;
    const uint32_t lo0 = __funnelshift_r(r0w0, r0w1, sh8), hi0 = __funnelshift_r(r0w1, r0w2, sh8);
    const uint32_t lo1 = __funnelshift_r(r1w0, r1w1, sh8), hi1 = __funnelshift_r(r1w1, r1w2, sh8);
    // lo = [c0t0 c1t0 c2t0 c0t1], hi = [c1t1 c2t1 . .]  ->  X = [c0t0 c0t1 c1t0 c1t1], Y = [c2t0 c2t1 . .]
    const uint32_t X0 = __byte_perm(lo0, hi0, 0x4130), Y0 = __byte_perm(lo0, hi0, 0x0052);
    const uint32_t X1 = __byte_perm(lo1, hi1, 0x4130), Y1 = __byte_perm(lo1, hi1, 0x0052);
    weights_u8(a, b, W0, W1);
    v0 = R::fin(__dp2a_lo(W1, X1, __dp2a_lo(W0, X0, R::BIAS)));
    v1 = R::fin(__dp2a_hi(W1, X1, __dp2a_hi(W0, X0, R::BIAS)));
    v2 = R::fin(__dp2a_lo(W1, Y1, __dp2a_lo(W0, Y0, R::BIAS)));
}

// 1 channel: the two taps of a row are adjacent bytes (2 aligned words per row cover them)
template <bool HALF_EVEN>
__device__ __forceinline__ uint32_t blend_u8x1(uint32_t r0w0, uint32_t r0w1, uint32_t r1w0, uint32_t r1w1, unsigned sh8,
                                               unsigned a, unsigned b, uint32_t& W0, uint32_t& W1) {
    using R = Rnd<HALF_EVEN>;
    const uint32_t lo0 = __funnelshift_r(r0w0, r0w1, sh8), lo1 = __funnelshift_r(r1w0, r1w1, sh8);
    weights_u8(a, b, W0, W1);
    return R::fin(__dp2a_lo(W1, lo1, __dp2a_lo(W0, lo0, R::BIAS)));
}
// 4 channels: a pixel is one aligned word; returns the four result bytes packed
template <bool HALF_EVEN>
__device__ __forceinline__ uint32_t blend_u8x4(uint32_t t00, uint32_t t01, uint32_t t10, uint32_t t11, unsigned a,
                                               unsigned b, uint32_t& W0, uint32_t& W1) {
    using R = Rnd<HALF_EVEN>;
    // [c0t0 c0t1 c1t0 c1t1] and [c2t0 c2t1 c3t0 c3t1] per row
    const uint32_t X0 = __byte_perm(t00, t01, 0x5140), Y0 = __byte_perm(t00, t01, 0x7362);
    const uint32_t X1 = __byte_perm(t10, t11, 0x5140), Y1 = __byte_perm(t10, t11, 0x7362);
    weights_u8(a, b, W0, W1);
    const uint32_t v0 = R::fin(__dp2a_lo(W1, X1, __dp2a_lo(W0, X0, R::BIAS)));
    const uint32_t v1 = R::fin(__dp2a_hi(W1, X1, __dp2a_hi(W0, X0, R::BIAS)));
    const uint32_t v2 = R::fin(__dp2a_lo(W1, Y1, __dp2a_lo(W0, Y0, R::BIAS)));
    const uint32_t v3 = R::fin(__dp2a_hi(W1, Y1, __dp2a_hi(W0, Y0, R::BIAS)));
    return v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
}

// words per pixel that hold its taps (two rows): C = 3 needs 3 aligned words per row, C = 1 and C = 4 need 2
template <int C>
struct TapWords {
    static constexpr int N = (C == 3) ? 6 : 4;
};
// phase 1: the aligned words around the taps of one pixel (o = byte offset of tap 00 in the box)
template <int C>
__device__ __forceinline__ void load_taps(const uint8_t* __restrict__ box, unsigned o, uint32_t (&w)[TapWords<C>::N]) {
    constexpr int PITCH = box_bytes(C) / 4;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(box + (o & ~3u));
    if constexpr (C == 3) {
        w[0] = q[0]; w[1] = q[1]; w[2] = q[2];
        w[3] = q[PITCH]; w[4] = q[PITCH + 1]; w[5] = q[PITCH + 2];
    } else {
        w[0] = q[0]; w[1] = q[1];
        w[2] = q[PITCH]; w[3] = q[PITCH + 1];
    }
}
// the same words from global memory: g = address of tap 00, all four taps inside the frame. Only words that overlap a
// tap byte are read (the frame is a multiple of 16 bytes, so every such word lies inside the buffer).
template <int C>
__device__ __forceinline__ void load_taps_global(const uint8_t* __restrict__ g, int W, uint32_t (&w)[TapWords<C>::N]) {
    const unsigned ph = (unsigned)(uintptr_t)g & 3u;
    const uint32_t* q0 = reinterpret_cast<const uint32_t*>(g - ph);
    const uint32_t* q1 = reinterpret_cast<const uint32_t*>(g - ph + (size_t)W * C);   // W % 16 == 0: same phase
    if constexpr (C == 3) {
        w[0] = __ldg(q0); w[1] = __ldg(q0 + 1); w[3] = __ldg(q1); w[4] = __ldg(q1 + 1);
        if (ph == 3u) { w[2] = __ldg(q0 + 2); w[5] = __ldg(q1 + 2); }
    } else if constexpr (C == 1) {
        w[0] = __ldg(q0); w[2] = __ldg(q1);
        if (ph == 3u) { w[1] = __ldg(q0 + 1); w[3] = __ldg(q1 + 1); }
    } else {
        w[0] = __ldg(q0); w[1] = __ldg(q0 + 1); w[2] = __ldg(q1); w[3] = __ldg(q1 + 1);
    }
}
// A value that depends on every word loaded for one pixel, at no extra cost: the first intermediate results of the
// blend (the compiler shares them with blend_store). Feeds the release of the box stage.
template <int C>
__device__ __forceinline__ uint32_t tap_digest(const uint32_t (&w)[TapWords<C>::N], unsigned o) {
    const unsigned sh8 = o << 3;
    if constexpr (C == 1) {
        return __funnelshift_r(w[0], w[1], sh8) | __funnelshift_r(w[2], w[3], sh8);
    } else if constexpr (C == 3) {
        const uint32_t lo0 = __funnelshift_r(w[0], w[1], sh8), hi0 = __funnelshift_r(w[1], w[2], sh8);
        const uint32_t lo1 = __funnelshift_r(w[3], w[4], sh8), hi1 = __funnelshift_r(w[4], w[5], sh8);
        return __byte_perm(lo0, hi0, 0x4130) | __byte_perm(lo1, hi1, 0x4130);
    } else {
        return __byte_perm(w[0], w[1], 0x5140) | __byte_perm(w[2], w[3], 0x5140);
    }
}
// phase 2: blend and store the C result bytes of the pixel at dst
template <bool HALF_EVEN, int C>
__device__ __forceinline__ void blend_store(const uint32_t (&w)[TapWords<C>::N], unsigned o, unsigned a, unsigned b,
                                            uint8_t* __restrict__ dst, uint32_t& W0, uint32_t& W1) {
    if constexpr (C == 1) {
        dst[0] = (uint8_t)blend_u8x1<HALF_EVEN>(w[0], w[1], w[2], w[3], o << 3, a, b, W0, W1);
    } else if constexpr (C == 3) {
        uint32_t v0, v1, v2;
        blend_u8x3<HALF_EVEN>(w[0], w[1], w[2], w[3], w[4], w[5], o << 3, a, b, W0, W1, v0, v1, v2);
        dst[0] = (uint8_t)v0; dst[1] = (uint8_t)v1; dst[2] = (uint8_t)v2;
    } else {
        *reinterpret_cast<uint32_t*>(dst) = blend_u8x4<HALF_EVEN>(w[0], w[1], w[2], w[3], a, b, W0, W1);
    }
}

// which taps of (ix, iy) lie inside the frame, one byte each (tap order 00, 01, 10, 11)
__device__ __forceinline__ uint32_t taps_in_frame(int ix, int iy, int H, int W) {
    const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
    const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
    return (x0 && y0 ? 1u : 0u) | (x1 && y0 ? 1u << 8 : 0u) | (x0 && y1 ? 1u << 16 : 0u) | (x1 && y1 ? 1u << 24 : 0u);
}

// this thread's four pixels of a tile: flow-mask bytes, sample positions relative to the box (bytes / rows), integer
// taps and 1/32-px fractions; returns whether all of them are covered by the box
template <int C, bool READ_FM>
__device__ __forceinline__ bool prepare_rows(const float2* frow, const uint8_t* mrow, float sign, float xg, float yg,
                                             int4 info, unsigned (&fmv)[4], int (&dxb)[4], int (&dy)[4], int (&ixs)[4],
                                             int (&iys)[4], unsigned (&fa)[4], unsigned (&fb)[4]) {
    constexpr int BWB = box_bytes(C);
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 fj = frow[j * TS];
        fmv[j] = READ_FM ? mrow[j * TS] : 1u;
        const float X = __fmaf_rn(sign, fj.x, xg), Y = __fmaf_rn(sign, fj.y, yg + (float)j);
        const QCoord qx = quantise_fast(X), qy = quantise_fast(Y);
        ixs[j] = qx.i; iys[j] = qy.i;
        fa[j] = (unsigned)qx.f; fb[j] = (unsigned)qy.f;
        dxb[j] = C * qx.i - info.x;
        dy[j] = qy.i - info.z;
        // Covered by the box? Everything else is decided per pixel. No range test is needed for the fast quantiser:
        // it is exact for |X| < 2^17, and beyond that (or for NaN / Inf) the integer it produces is far outside
        // [-2^16, 2^16], so such a pixel can never pass the box test (frames are smaller than 32768).
        ok = ok && (unsigned)dxb[j] <= (unsigned)(BWB - 2 * C) && (unsigned)dy[j] <= (unsigned)(BH - 2);
    }
    return ok;
}

// The global-tap path (out of line, see mixed_rows): taps of rows [J0, J1) of a warp's 4 x 32 pixels, all requested
// before the first use; then release of the box (with the last rows), blend, validity. A pixel covered by the box reads
// it; a pixel outside the box whose four taps lie inside the frame reads the same aligned words from global memory (same
// arithmetic); a pixel that samples nothing but the zero border is zero / invalid; only the pixels that straddle the
// frame border outside the box go through the out-of-line border routine.
template <int C, bool HALF_EVEN, int MM, int J0, int J1, class BStage>
__device__ __forceinline__ void sample_rows_mixed(const BStage& bs, uint8_t* orow, uint8_t* mrow, const unsigned (&fmv)[4],
                                            const int (&dxb)[4], const int (&dy)[4], const int (&ixs)[4],
                                            const int (&iys)[4], const unsigned (&fa)[4], const unsigned (&fb)[4],
                                            int4 info, int n, const uint8_t* __restrict__ img,
                                            const uint8_t* __restrict__ pmask, int H, int W, int rule, unsigned s_pass,
                                            uint64_t* bempty, uint32_t* sink, unsigned lane) {
    constexpr int BWB = box_bytes(C);
    constexpr int NW = TapWords<C>::N;
    uint32_t w[4][NW];
    uint32_t mt[4];
    auto in_box = [&](int j) {
        return (unsigned)dxb[j] <= (unsigned)(BWB - 2 * C) && (unsigned)dy[j] <= (unsigned)(BH - 2);
    };
    auto interior = [&](int j) {   // all four taps inside the frame (the fast quantiser's garbage never is)
        return (unsigned)ixs[j] < (unsigned)(W - 1) && (unsigned)iys[j] < (unsigned)(H - 1);
    };
#pragma unroll
    for (int j = J0; j < J1; ++j) {
        if (in_box(j)) {
            load_taps<C>(bs.img, (unsigned)(dy[j] * BWB + dxb[j]), w[j]);
            if (MM == MM_PMASK) {
                const uint8_t* mm = bs.m + (dy[j] * BMW + (ixs[j] - info.y));
                mt[j] = (uint32_t)mm[0] | ((uint32_t)mm[1] << 8) | ((uint32_t)mm[BMW] << 16) |
                        ((uint32_t)mm[BMW + 1] << 24);
            }
        } else {
#pragma unroll
            for (int k = 0; k < NW; ++k) w[j][k] = 0u;
            mt[j] = 0u;
            if (interior(j)) {
                const long long px = (long long)n * ((long long)H * W) + ((long long)iys[j] * W + ixs[j]);
                load_taps_global<C>(img + px * C, W, w[j]);
                if (MM == MM_PMASK) {
                    const uint8_t* mm = pmask + px;
                    mt[j] = (uint32_t)__ldg(mm) | ((uint32_t)__ldg(mm + 1) << 8) | ((uint32_t)__ldg(mm + W) << 16) |
                            ((uint32_t)__ldg(mm + W + 1) << 24);
                }
            }
        }
    }
    if (J1 == 4) {
        // everything loaded from the box is consumed by the reduction below before the stage is handed back
        unsigned dep = 0;
#pragma unroll
        for (int j = J0; j < J1; ++j)         // every word: the loads may be issued in any order
            dep |= tap_digest<C>(w[j], (unsigned)dxb[j]) | (MM == MM_PMASK ? mt[j] : 0u);
        dep = __reduce_or_sync(0xffffffffu, dep);
        if (lane == 0) mbar_arrive_after(bempty, dep, sink);
    }
#pragma unroll
    for (int j = J0; j < J1; ++j) {
        uint8_t* dst = orow + j * (TS * C);
        const bool boxed = in_box(j);
        if (boxed || interior(j)) {
            // the byte phase of tap 00 inside its aligned word is the same in the box and in the frame: box starts and
            // row pitches are multiples of 16 bytes
            uint32_t W0, W1;
            blend_store<HALF_EVEN, C>(w[j], (unsigned)dxb[j], fa[j], fb[j], dst, W0, W1);
            if (MM == MM_PMASK) {
                const unsigned valid = __dp2a_hi(W1, mt[j], __dp2a_lo(W0, mt[j], 0u)) >= s_pass;
                mrow[j * TS] = (uint8_t)(valid & fmv[j]);
            } else if (MM == MM_GEOM) {
                unsigned valid = 1u;
                if (boxed && !info.w) {         // tile-uniform: the box reaches over the frame border
                    const uint32_t in = taps_in_frame(ixs[j], iys[j], H, W);
                    valid = __dp2a_hi(W1, in, __dp2a_lo(W0, in, 0u)) >= s_pass;
                }
                mrow[j * TS] = (uint8_t)(valid & fmv[j]);
            }
        } else {
            // (quantised coordinates beyond the fast quantiser's range are far outside the frame)
            unsigned valid = 0u;
            if (ixs[j] < -1 || iys[j] < -1 || ixs[j] >= W || iys[j] >= H) {
#pragma unroll
                for (int c = 0; c < C; ++c) dst[c] = 0;      // every tap lies outside the frame
            } else {
                const size_t fbase = (size_t)n * ((size_t)H * W);
                const unsigned long long r = border_px_u8q<C>(img + fbase * C, MM == MM_PMASK ? pmask + fbase : nullptr,
                                                              ixs[j], iys[j], (int)fa[j], (int)fb[j], H, W, HALF_EVEN,
                                                              rule);
#pragma unroll
                for (int c = 0; c < C; ++c) dst[c] = (uint8_t)(r >> (8 * c));
                valid = (unsigned)(r >> 32);
            }
            if (MM != MM_NONE) mrow[j * TS] = (uint8_t)(valid & fmv[j] & 1u);
        }
    }
}

// The rare path, out of line so that it does not weigh on the register allocation of the kernel's main loop: redoes the
// preparation from shared memory and samples with sample_rows_mixed, two rows at a time.
static __device__ unsigned long long g_mixed_warp_tiles;   // test hook: how often the out-of-line path ran (per warp and tile)

template <int C, bool HALF_EVEN, int MM, bool FM, class SM>
__device__ __noinline__ void mixed_rows(SM* sm, unsigned s, unsigned b, const uint8_t* __restrict__ img,
                                        const uint8_t* __restrict__ pmask, float sign, int H, int W, int rule) {
    typename SM::PStage* ps = &sm->ps[s];
    const typename SM::BStage* bs = &sm->bs[b];
    uint64_t* bempty = &sm->bempty[b];
    const int4 info = sm->binfo[b][0], tile = sm->binfo[b][1];
    const unsigned lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    if (lane == 0) atomicAdd(&g_mixed_warp_tiles, 1ull);
    const unsigned s_pass = rule == OFK_RULE_STRICT ? 1024u : (rule == OFK_RULE_GT_HALF ? 513u : 512u);
    const unsigned own = wrp * 4 * TS + lane;
    uint8_t* mrow = ps->fm + own;
    uint8_t* orow = reinterpret_cast<uint8_t*>(ps->f) + (wrp * 4 * TS * 8 + lane * C);
    unsigned fmv[4], fa[4], fb[4];
    int dxb[4], dy[4], ixs[4], iys[4];
    prepare_rows<C, FM && MM != MM_NONE>(ps->f + own, mrow, sign, (float)(tile.x + (int)lane),
                                         (float)(tile.y + (int)wrp * 4), info, fmv, dxb, dy, ixs, iys, fa, fb);
    __syncwarp();   // every lane holds its flow values before the first result bytes overwrite the tile rows
    sample_rows_mixed<C, HALF_EVEN, MM, 0, 2>(*bs, orow, mrow, fmv, dxb, dy, ixs, iys, fa, fb, info, tile.z, img, pmask, H,
                                              W, rule, s_pass, bempty, &sm->sink[wrp], lane);
    sample_rows_mixed<C, HALF_EVEN, MM, 2, 4>(*bs, orow, mrow, fmv, dxb, dy, ixs, iys, fa, fb, info, tile.z, img, pmask, H,
                                              W, rule, s_pass, bempty, &sm->sink[wrp], lane);
}

template <int C, bool HALF_EVEN, int MM, bool FM, int NP, int NB, int LA, int PW>
__global__ void __launch_bounds__((NCW + PW) * 32, 3) warp_u8_ws_kernel(const __grid_constant__ Maps maps,
                                                                      const uint8_t* __restrict__ img,
                                                                      const uint8_t* __restrict__ pmask, float sign,
                                                                      int rule, int H, int W, unsigned tiles_x,
                                                                      unsigned tiles_per_frame, unsigned total_tiles) {
    static_assert(NP >= NB + LA + 1, "P stages must outlive the box pipeline and the output store");
    using SM = Smem<NP, NB, MM == MM_PMASK, C>;
    constexpr int BWB = box_bytes(C);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);   // no static shared memory in this kernel: the window starts aligned
    if (smem_u32(smem_raw) & 127u) __trap();
    const unsigned tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const unsigned first = blockIdx.x, stride = gridDim.x;
    if (first >= total_tiles) return;
    const unsigned T = (total_tiles - first + stride - 1) / stride;
    constexpr uint32_t P_BYTES = TS * TS * 8 + (FM ? TS * TS : 0);
    constexpr uint32_t B_BYTES = BH * BWB + (MM == MM_PMASK ? BH * BMW : 0);

    if (tid == 0) {
        for (int k = 0; k < NP; ++k) { mbar_init(&sm.pfull[k], 1); mbar_init(&sm.pempty[k], NCW); }
        for (int k = 0; k < NB; ++k) { mbar_init(&sm.bfull[k], 1); mbar_init(&sm.bempty[k], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (PW == 2 && wrp == NCW + 1) {
        // -------------------------------------------------------------------------------------------- P loader warp
        // Streams the flow tiles into the ring of P stages as fast as stages are released, independent of the box
        // pipeline.
        if (lane == 0) {
            const int tiles_y = (int)(tiles_per_frame / tiles_x);
            TileIter pit;
            pit.init(first, stride, (int)tiles_x, tiles_y);
            unsigned ps_i = 0, ps_ph = 0;
            for (unsigned i = 0; i < T; ++i) {
                const int tx0 = pit.tx * TS, ty0 = pit.ty * TS, n = pit.n;
                pit.advance((int)tiles_x, tiles_y);
                if (i >= NP) mbar_wait(&sm.pempty[ps_i], ps_ph ^ 1);
                mbar_expect_tx(&sm.pfull[ps_i], P_BYTES);
                tma_load_3d(sm.ps[ps_i].f, &maps.f, &sm.pfull[ps_i], tx0, ty0, n);
                if (FM) tma_load_3d(sm.ps[ps_i].fm, &maps.fm, &sm.pfull[ps_i], tx0, ty0, n);
                if (++ps_i == NP) { ps_i = 0; ps_ph ^= 1; }
            }
        }
        return;
    }
    if (wrp == NCW) {
        // ------------------------------------------------------------------------------------------ producer warp
        const int tiles_y = (int)(tiles_per_frame / tiles_x);
        TileIter pit, bit;   // cursors of the flow-tile loads and of the box preparation
        pit.init(first, stride, (int)tiles_x, tiles_y);
        bit = pit;
        unsigned ps_i = 0, ps_ph = 0;
        auto issue_p = [&](bool wait_empty) {   // lane 0 only; tiles are issued in order
            const int tx0 = pit.tx * TS, ty0 = pit.ty * TS, n = pit.n;
            pit.advance((int)tiles_x, tiles_y);
            if (wait_empty) mbar_wait(&sm.pempty[ps_i], ps_ph ^ 1);
            mbar_expect_tx(&sm.pfull[ps_i], P_BYTES);
            tma_load_3d(sm.ps[ps_i].f, &maps.f, &sm.pfull[ps_i], tx0, ty0, n);
            if (FM) tma_load_3d(sm.ps[ps_i].fm, &maps.fm, &sm.pfull[ps_i], tx0, ty0, n);
            if (++ps_i == NP) { ps_i = 0; ps_ph ^= 1; }
        };
        if (PW == 1 && lane == 0)
            for (unsigned k = 0; k < (unsigned)LA && k < T; ++k) issue_p(false);
        unsigned s = 0, s_ph = 0, b = 0, b_ph = 0;
        for (unsigned i = 0; i < T; ++i) {
            const int tx0 = bit.tx * TS, ty0 = bit.ty * TS, n = bit.n;
            bit.advance((int)tiles_x, tiles_y);
            if (lane == 0) OFK_TR(i, 0);
            mbar_wait(&sm.pfull[s], s_ph);
            if (lane == 0) OFK_TR(i, 1);
            // Sample positions on a 4 x 4 grid over the tile (corners included): exact bounding box for affine fields, an
            // estimate otherwise (consumers verify per pixel). One shared-memory load per lane, four integer warp
            // reductions (REDUX): the producer's serial time per tile is what paces the whole pipeline.
            int x0, x1, y0, y1;
            {
                const int r = (int)(((lane >> 2) & 3) * 31 + 1) / 3, c = (int)((lane & 3) * 31 + 1) / 3;   // 0, 10, 21, 31
                const int x = tx0 + c, y = ty0 + r;
                int fx0 = 0x7fffffff, fx1 = -0x7fffffff, fy0 = 0x7fffffff, fy1 = -0x7fffffff;
                if (x < W && y < H) {
                    const float2 v = sm.ps[s].f[r * TS + c];
                    const float lim = 60000.f;
                    const float X = fminf(fmaxf(__fmaf_rn(sign, v.x, (float)x), -lim), lim);
                    const float Y = fminf(fmaxf(__fmaf_rn(sign, v.y, (float)y), -lim), lim);
                    fx0 = fx1 = __float2int_rd(X);
                    fy0 = fy1 = __float2int_rd(Y);
                }
                x0 = __reduce_min_sync(0xffffffffu, fx0); x1 = __reduce_max_sync(0xffffffffu, fx1);
                y0 = __reduce_min_sync(0xffffffffu, fy0); y1 = __reduce_max_sync(0xffffffffu, fy1);
            }
            if (lane == 0) {
                // integer tap range, clamped to the taps that can contribute: ix in [-1, W-1], iy in [-1, H-1]
                x0 = max(-1, min(W - 1, x0)); x1 = max(-1, min(W - 1, x1));
                y0 = max(-1, min(H - 1, y0)); y1 = max(-1, min(H - 1, y1));
                // centre the needed range [x0, x1 + 1] in the box (spare margin on both sides for curved flows)
                const int needw = x1 + 2 - x0, needh = y1 + 2 - y0;
                const int vx0 = x0 - max(0, (box_px(C) - needw) / 2);
                const int bx0 = (C * vx0) & ~15;             // 16-byte aligned box start, in bytes of the C*W-byte row
                const int mx0 = vx0 & ~15;
                const int by0 = y0 - max(0, (BH - needh) / 2);
                OFK_TR(i, 2);
                if (i >= NB) mbar_wait(&sm.bempty[b], b_ph ^ 1);
                OFK_TR(i, 3);
                // every tap of a pixel covered by a box that lies inside the frame is inside the frame
                const int inframe = (bx0 >= 0 && bx0 + BWB <= C * W && by0 >= 0 && by0 + BH <= H) ? 1 : 0;
                sm.binfo[b][0] = make_int4(bx0, mx0, by0, inframe);
                sm.binfo[b][1] = make_int4(tx0, ty0, n, 0);
                mbar_expect_tx(&sm.bfull[b], B_BYTES);
                tma_load_3d(sm.bs[b].img, &maps.ib, &sm.bfull[b], bx0, by0, n);
                if (MM == MM_PMASK) tma_load_3d(sm.bs[b].m, &maps.pmb, &sm.bfull[b], mx0, by0, n);
                OFK_TR(i, 4);
                if (PW == 1 && i + LA < T) issue_p(i + LA >= NP);
                OFK_TR(i, 5);
            }
            __syncwarp();
            if (++s == NP) { s = 0; s_ph ^= 1; }
            if (++b == NB) { b = 0; b_ph ^= 1; }
        }
        return;
    }

    // ---------------------------------------------------------------------------------------------- consumer warps
    const unsigned s_pass = rule == OFK_RULE_STRICT ? 1024u : (rule == OFK_RULE_GT_HALF ? 513u : 512u);   // S >= s_pass
    const unsigned own = wrp * 4 * TS + lane;          // this thread's pixel in row 4w of a tile (+ j * TS)
    const unsigned own3 = wrp * 4 * TS * 8 + lane * C; // byte offset of its first output byte in the flow tile (+ j * 32 * C)
    unsigned s = 0, s_ph = 0, b = 0, b_ph = 0;
    int prev_s = -1;
    for (unsigned i = 0; i < T; ++i) {
        if (lane == 0 && wrp == 0) OFK_TR(i, 8);
        mbar_wait(&sm.bfull[b], b_ph);
        mbar_wait(&sm.pfull[s], s_ph);
        if (lane == 0 && wrp == 0) OFK_TR(i, 9);
        const int4 info = sm.binfo[b][0], tile = sm.binfo[b][1];
        typename SM::PStage& ps = sm.ps[s];
        const typename SM::BStage& bs = sm.bs[b];
        const int tx0 = tile.x, ty0 = tile.y, n = tile.z;
        const float xg = (float)(tx0 + (int)lane), yg = (float)(ty0 + (int)wrp * 4);
        const float2* frow = ps.f + own;
        uint8_t* mrow = ps.fm + own;                                        // validity bytes, in place
        uint8_t* orow = reinterpret_cast<uint8_t*>(ps.f) + own3;            // image bytes, in place

        // flow-mask bytes of the thread's pixels: read only where the validity byte is not already in place (see below)
        auto load_fmv = [&](unsigned (&fmv)[4]) {
#pragma unroll
            for (int j = 0; j < 4; ++j) fmv[j] = (FM && MM != MM_NONE) ? mrow[j * TS] : 1u;
        };
        // Integer taps stay biased by the quantiser's magic number in the main path (rx = ix + KB): the bias folds into
        // the per-tile box origin, which saves the subtraction per coordinate; ix / iy proper are formed where a slow
        // path needs them. (Coordinates beyond the fast quantiser's range, NaN and Inf give integers far outside
        // [-2^16, 2^16] either way and can never pass the box test; frames are smaller than 32768.)
        constexpr int KB = 0x4B400000 >> 5;
        const int cx = info.x + C * KB, cz = info.z + KB;
        int dxb[4], dy[4], rx[4], ry[4];
        unsigned fa[4], fb[4];
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fj = frow[j * TS];
            const float X = __fmaf_rn(sign, fj.x, xg), Y = __fmaf_rn(sign, fj.y, yg + (float)j);
            const int bx = __float_as_int(__fmaf_rn(X, 32.0f, 12582912.0f)), by = __float_as_int(__fmaf_rn(Y, 32.0f, 12582912.0f));
            rx[j] = bx >> 5; ry[j] = by >> 5;
            fa[j] = (unsigned)bx & 31u; fb[j] = (unsigned)by & 31u;
            dxb[j] = C * rx[j] - cx;
            dy[j] = ry[j] - cz;
            ok = ok && (unsigned)dxb[j] <= (unsigned)(BWB - 2 * C) && (unsigned)dy[j] <= (unsigned)(BH - 2);
        }
        if (__all_sync(0xffffffffu, ok)) {
            uint32_t w[4][TapWords<C>::N];
            uint32_t mt[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                load_taps<C>(bs.img, (unsigned)(dy[j] * BWB + dxb[j]), w[j]);
                if (MM == MM_PMASK) {
                    const uint8_t* mm = bs.m + (dy[j] * BMW + (rx[j] - KB - info.y));
                    mt[j] = (uint32_t)mm[0] | ((uint32_t)mm[1] << 8) | ((uint32_t)mm[BMW] << 16) |
                            ((uint32_t)mm[BMW + 1] << 24);
                }
            }
            // everything loaded from the box is consumed by the reduction below before the stage is handed back
            unsigned dep = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)       // every word: the loads may be issued in any order
                dep |= tap_digest<C>(w[j], (unsigned)(dy[j] * BWB + dxb[j])) | (MM == MM_PMASK ? mt[j] : 0u);
            dep = __reduce_or_sync(0xffffffffu, dep);
            if (lane == 0) mbar_arrive_after(&sm.bempty[b], dep, &sm.sink[wrp]);
            if (lane == 0 && wrp == 0) OFK_TR(i, 10);
            if (MM == MM_GEOM && info.w) {
                // The box lies inside the frame (tile-uniform; all tiles but those along the frame border): every tap is
                // inside, so the validity byte is the flow-mask byte -- which already sits in the tile, in place.
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t W0, W1;
                    blend_store<HALF_EVEN, C>(w[j], (unsigned)(dy[j] * BWB + dxb[j]), fa[j], fb[j], orow + j * (TS * C), W0, W1);
                    if (!FM) mrow[j * TS] = (uint8_t)1;
                }
            } else {
                unsigned fmv[4];
                load_fmv(fmv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t W0, W1;
                    blend_store<HALF_EVEN, C>(w[j], (unsigned)(dy[j] * BWB + dxb[j]), fa[j], fb[j], orow + j * (TS * C), W0, W1);
                    if (MM == MM_PMASK) {
                        const unsigned valid = __dp2a_hi(W1, mt[j], __dp2a_lo(W0, mt[j], 0u)) >= s_pass;
                        mrow[j * TS] = (uint8_t)(valid & fmv[j]);
                    } else if (MM == MM_GEOM) {         // the box reaches over the frame border
                        const uint32_t in = taps_in_frame(rx[j] - KB, ry[j] - KB, H, W);
                        const unsigned valid = __dp2a_hi(W1, in, __dp2a_lo(W0, in, 0u)) >= s_pass;
                        mrow[j * TS] = (uint8_t)(valid & fmv[j]);
                    }
                }
            }
        } else {
            // Some pixel of this warp is not covered by the box. Along the frame border that is the rule (the box is
            // clamped to the frame): those pixels sample nothing but the zero border. Pixels that need taps from inside
            // the frame (discontinuous or noisy flow, estimate too small) send the warp to the out-of-line path.
            bool need_global = false;
            const int fw = cold_value(W), fh = cold_value(H);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool inbox = (unsigned)dxb[j] <= (unsigned)(BWB - 2 * C) && (unsigned)dy[j] <= (unsigned)(BH - 2);
                need_global = need_global || !(inbox || rx[j] < KB - 1 || ry[j] < KB - 1 || rx[j] - KB >= fw || ry[j] - KB >= fh);
            }
            if (__any_sync(0xffffffffu, need_global)) {
                mixed_rows<C, HALF_EVEN, MM, FM, SM>(&sm, s, b, img, pmask, sign, H, W, rule);
            } else {
                unsigned fmv[4];
                load_fmv(fmv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool inbox = (unsigned)dxb[j] <= (unsigned)(BWB - 2 * C) && (unsigned)dy[j] <= (unsigned)(BH - 2);
                    uint8_t* dst = orow + j * (TS * C);
                    unsigned valid = 0u;
                    if (inbox) {
                        const unsigned o = (unsigned)(dy[j] * BWB + dxb[j]);
                        uint32_t wj[TapWords<C>::N], W0, W1;
                        load_taps<C>(bs.img, o, wj);
                        blend_store<HALF_EVEN, C>(wj, o, fa[j], fb[j], dst, W0, W1);
                        uint32_t in;
                        if (MM == MM_PMASK) {
                            const uint8_t* mm = bs.m + (dy[j] * BMW + (rx[j] - KB - info.y));
                            in = (uint32_t)mm[0] | ((uint32_t)mm[1] << 8) | ((uint32_t)mm[BMW] << 16) |
                                 ((uint32_t)mm[BMW + 1] << 24);
                        } else {
                            in = taps_in_frame(rx[j] - KB, ry[j] - KB, H, W);
                        }
                        valid = __dp2a_hi(W1, in, __dp2a_lo(W0, in, 0u)) >= s_pass;
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c) dst[c] = 0;      // every tap lies outside the frame
                    }
                    if (MM != MM_NONE) mrow[j * TS] = (uint8_t)(valid & fmv[j] & 1u);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.bempty[b]);
            }
        }
        // the warp's 4 result rows go out as bulk tensor stores (clipped at the frame border by the hardware)
        fence_async_smem();
        __syncwarp();
        // (issued under elect.sync: the compiler moves the operands to uniform registers without a waterfall loop)
        const int u_tx = tx0, u_ty = ty0 + (int)wrp * 4, u_n = n;
        const uint32_t u_src = smem_u32(ps.f + (int)wrp * 4 * TS), u_srcm = smem_u32(ps.fm + (int)wrp * 4 * TS);
        if (elect_one()) {
            tma_store_3d_u(&maps.oi, u_src, C * u_tx, u_ty, u_n);
            if (MM != MM_NONE) tma_store_3d_u(&maps.om, u_srcm, u_tx, u_ty, u_n);
            bulk_commit();
            if (prev_s >= 0) {
                bulk_wait_read<1>();                 // the previous tile's rows have been read out of shared memory
                mbar_arrive(&sm.pempty[prev_s]);
            }
        }
        if (lane == 0 && wrp == 0) OFK_TR(i, 11);
        prev_s = (int)s;
        if (++s == NP) { s = 0; s_ph ^= 1; }
        if (++b == NB) { b = 0; b_ph ^= 1; }
    }
    if (elect_one()) bulk_wait_all();   // the lane that committed the groups
}

#ifndef OFK_WS_NP
#define OFK_WS_NP 6
#define OFK_WS_NB 2
#define OFK_WS_LA 3
#endif
constexpr int WS_NP = OFK_WS_NP, WS_NB = OFK_WS_NB, WS_LA = OFK_WS_LA, WS_PW = 1;   // measured best: one producer warp (a P loader warp costs the 3rd CTA its registers)

template <int C, bool HALF_EVEN, int MM, bool FM>
static int launch_variant(const Maps& maps, const uint8_t* img, const uint8_t* pmask, float sign, int rule, int H, int W,
                          unsigned tx, unsigned ty, unsigned total, int ctas_per_sm, cudaStream_t st) {
    using SM = Smem<WS_NP, WS_NB, MM == MM_PMASK, C>;
    static std::atomic<bool> attr_done_dev[64];   // the attribute is per device (racing threads set it twice)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    auto kernel = warp_u8_ws_kernel<C, HALF_EVEN, MM, FM, WS_NP, WS_NB, WS_LA, WS_PW>;
    if (!attr_done_dev[dev]) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM)) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        attr_done_dev[dev] = true;
    }
    static std::atomic<int> resident[64];       // co-resident CTAs per SM (shared memory / registers), per device
    if (resident[dev] == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, (NCW + WS_PW) * 32, sizeof(SM)) != cudaSuccess || nb < 1) {
            cudaGetLastError();
            nb = 1;
        }
        resident[dev] = nb;
    }
    const int res = resident[dev].load();
    unsigned grid = (unsigned)(sm_count() * (ctas_per_sm < res ? ctas_per_sm : res));
    if (grid > total) grid = total;
    grid = cap_ctas(grid);
    kernel<<<grid, (NCW + WS_PW) * 32, sizeof(SM), st>>>(maps, img, pmask, sign, rule, H, W, tx, tx * ty, total);
    return 1;
}

template <int C>
static int launch_channels(bool half_even, const void* payload, const float* flow, float sign, const uint8_t* pmask,
                           const uint8_t* fmask, void* out, uint8_t* omask, int rule, int N, int H, int W,
                           cudaStream_t st) {
    const int mm = omask == nullptr ? MM_NONE : (pmask != nullptr ? MM_PMASK : MM_GEOM);
    const bool fm = omask != nullptr && fmask != nullptr;
    Maps maps;
    if (!make_map3(&maps.f, flow, 8, W, H, N, TS, TS) ||
        !make_map3(&maps.ib, payload, 1, (size_t)W * C, H, N, box_bytes(C), BH) ||
        !make_map3(&maps.oi, out, 1, (size_t)W * C, H, N, TS * C, 4))
        return 0;
    maps.fm = maps.pmb = maps.om = maps.f;
    if (fm && !make_map3(&maps.fm, fmask, 1, W, H, N, TS, TS)) return 0;
    if (mm == MM_PMASK && !make_map3(&maps.pmb, pmask, 1, W, H, N, BMW, BH)) return 0;
    if (mm != MM_NONE && !make_map3(&maps.om, omask, 1, W, H, N, TS, 4)) return 0;
    const unsigned tx = (W + TS - 1) / TS, ty = (H + TS - 1) / TS;
    if ((double)tx * ty * N >= 4.0e9) return 0;
    const unsigned total = tx * ty * (unsigned)N;
    static int cps = -1;
    if (cps < 0) {
        const char* e = getenv("OFK_WARP_WS_CPS");
        cps = e ? atoi(e) : 3;
        if (cps < 1 || cps > 4) cps = 3;
    }
    int rc;
#define OFK_WV(HE, MMV, FMV) \
    rc = launch_variant<C, HE, MMV, FMV>(maps, (const uint8_t*)payload, pmask, sign, rule, H, W, tx, ty, total, cps, st)
#define OFK_WV_HE(MMV, FMV)            \
    do {                               \
        if (half_even) OFK_WV(true, MMV, FMV);  \
        else OFK_WV(false, MMV, FMV);  \
    } while (0)
    if (mm == MM_NONE) OFK_WV_HE(MM_NONE, false);
    else if (mm == MM_GEOM) { if (fm) OFK_WV_HE(MM_GEOM, true); else OFK_WV_HE(MM_GEOM, false); }
    else { if (fm) OFK_WV_HE(MM_PMASK, true); else OFK_WV_HE(MM_PMASK, false); }
#undef OFK_WV_HE
#undef OFK_WV
    return rc;
}

}  // namespace wtws

unsigned long long warp_ws_mixed_count() {
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, wtws::g_mixed_warp_tiles, sizeof(v)) != cudaSuccess) cudaGetLastError();
    return v;
}

bool warp_ws_enabled() {
    static std::atomic<int> state{-1};
    if (state.load() < 0) {
        const char* e = getenv("OFK_WARP_WS");
        state.store((e != nullptr && e[0] == '0') ? 0 : 1);
    }
    return state.load() == 1;
}

// uint8 images with C = 1, 3 or 4 interleaved channels. Returns 1 if the kernel was launched, 0 if the configuration
// is not eligible (caller uses the gather kernels), negative OFK_E* on error.
int launch_warp_u8_ws(int C, bool half_even, const void* payload, const float* flow, float sign, const uint8_t* pmask,
                      const uint8_t* fmask, void* out, uint8_t* omask, int rule, int N, int H, int W, cudaStream_t st) {
    using namespace wtws;
    if (W % 16 != 0 || H >= 32768 || W >= 32768) return 0;    // 16-byte row pitch of the uint8 tensors
    const uintptr_t align = reinterpret_cast<uintptr_t>(payload) | reinterpret_cast<uintptr_t>(flow) |
                            reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(omask) |
                            reinterpret_cast<uintptr_t>(pmask) | reinterpret_cast<uintptr_t>(fmask);
    if (align & 15) return 0;
    int rc;
    switch (C) {
        case 1: rc = launch_channels<1>(half_even, payload, flow, sign, pmask, fmask, out, omask, rule, N, H, W, st); break;
        case 3: rc = launch_channels<3>(half_even, payload, flow, sign, pmask, fmask, out, omask, rule, N, H, W, st); break;
        case 4: rc = launch_channels<4>(half_even, payload, flow, sign, pmask, fmask, out, omask, rule, N, H, W, st); break;
        default: return 0;
    }
    if (rc != 1) return rc;
    OFK_LAUNCHED();
    return 1;
}

}  // namespace ofk
