// Warp-specialised, TMA-staged composition kernel (experiment; becomes oflibnumpy_b200/csrc/combine3_ws.cu).
//
// CTA = 1 producer warp + NCW consumer warps, persistent over 32x32 output tiles.
//   producer : TMA-loads the pointwise operand tile (P, Pm) LA tiles ahead; when a P tile has landed it estimates the
//              bounding box of the tile's sample positions from the tile perimeter, and TMA-loads that box of the
//              gathered operand (G, Gm) into a box stage (hardware zero fill outside the frame == cv2 constant border
//              and "invalid outside").
//   consumers: wait for the box, gather the four taps of every pixel from shared memory (32-bit offsets, no bounds
//              tests), blend with packed f32x2 arithmetic in cv2.remap's rounding sequence, store.
// Pixels whose taps are not covered by the box (box estimate too small, wild flows) take an exact global-memory path
// individually, so the result never depends on the quality of the estimate.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace c3ws {

constexpr int TS = 32;
constexpr int BW = 48, BH = 48;   // vector box (pixels)
constexpr int BMW = 64;           // mask box width (bytes): start is aligned down to 16
constexpr int NCW = 8;            // consumer warps

template <int NP, int NB>
struct Smem {
    struct PStage {
        float2 p[TS * TS];
        uint8_t pm[TS * TS];
    };
    struct BStage {
        float2 v[BH * BW];
        uint8_t m[BH * BMW];
    };
    alignas(128) PStage ps[NP];
    alignas(128) BStage bs[NB];
    alignas(16) int4 binfo[NB][2];   // {vx0, vx0 - mx0, by0, -}, {tx0, ty0, n, -}
    uint64_t pfull[NP], pempty[NP], bfull[NB], bempty[NB];
};

struct Maps {
    CUtensorMap p, pm, gb, gmb, ov, om;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive once `dep` has been computed: ties the release of a buffer to the registers loaded from it
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, unsigned dep) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];  // after %1" ::"r"(smem_u32(bar)), "r"(dep) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// packed f32x2 arithmetic (sm_100): both halves are IEEE round-to-nearest, no contraction
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(uint64_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// a * b rounded once: fma with a (-0, -0) addend the compiler cannot see through (kernel parameter). A plain
// mul.rn.f32x2 feeding add.rn.f32x2 is contracted into FFMA2 by ptxas 12.9 in spite of the .rn qualifiers, which would
// change the rounding sequence of cv2.remap.
__device__ __forceinline__ uint64_t mul2_nofuse(uint64_t a, uint64_t b, uint64_t negzero2) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(negzero2));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __forceinline__ void quant(float X, int& i, int& f) {
    const int bits = __float_as_int(__fmaf_rn(X, 32.0f, 12582912.0f)) - 0x4B400000;
    i = bits >> 5;
    f = bits & 31;
}

// exact per-pixel path from global memory (any coordinates); returns sample in (u, v) and strict validity
__device__ __noinline__ float4 slow_sample(const float2* __restrict__ G, const uint8_t* __restrict__ Gm, int H, int W,
                                           float X, float Y) {
    int sx = __float2int_rn(X * 32.0f), sy = __float2int_rn(Y * 32.0f);
    const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
    const int a = sx & 31, b = sy & 31;
    const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
    const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
    const long long o = (long long)iy * W + ix;
    const float2 z = make_float2(0.f, 0.f);
    const float2 t00 = (x0 && y0) ? __ldg(G + o) : z, t01 = (x1 && y0) ? __ldg(G + o + 1) : z;
    const float2 t10 = (x0 && y1) ? __ldg(G + o + W) : z, t11 = (x1 && y1) ? __ldg(G + o + W + 1) : z;
    const int w00 = (32 - a) * (32 - b), w01 = a * (32 - b), w10 = (32 - a) * b, w11 = a * b;
    int S = 0;
    if (x0 && y0 && (!Gm || __ldg(Gm + o))) S += w00;
    if (x1 && y0 && (!Gm || __ldg(Gm + o + 1))) S += w01;
    if (x0 && y1 && (!Gm || __ldg(Gm + o + W))) S += w10;
    if (x1 && y1 && (!Gm || __ldg(Gm + o + W + 1))) S += w11;
    const float s = 1.0f / 1024.0f;
    const float f00 = float(w00) * s, f01 = float(w01) * s, f10 = float(w10) * s, f11 = float(w11) * s;
    float4 r;
    r.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00.x, f00), __fmul_rn(t01.x, f01)), __fmul_rn(t10.x, f10)),
                    __fmul_rn(t11.x, f11));
    r.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00.y, f00), __fmul_rn(t01.y, f01)), __fmul_rn(t10.y, f10)),
                    __fmul_rn(t11.y, f11));
    r.z = (S == 1024) ? 1.f : 0.f;
    r.w = 0.f;
    return r;
}

// Linear tile index -> (frame, tile row, tile column), advanced by the grid stride without divisions.
struct TileIter {
    int n, ty, tx, dn, dy, dx;
    __device__ __forceinline__ void init(unsigned t, unsigned stride, int tiles_x, int tiles_y) {
        const unsigned per_frame = (unsigned)tiles_x * tiles_y;
        n = t / per_frame;
        unsigned r = t - n * per_frame;
        ty = r / tiles_x;
        tx = r - ty * tiles_x;
        dn = stride / per_frame;
        r = stride - dn * per_frame;
        dy = r / tiles_x;
        dx = r - dy * tiles_x;
    }
    __device__ __forceinline__ void advance(int tiles_x, int tiles_y) {
        tx += dx;
        if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
        ty += dy;
        if (ty >= tiles_y) { ty -= tiles_y; ++n; }
        n += dn;
    }
};

// bulk tensor store shared -> global (clipped at the tensor bounds), tracked by the issuing thread's bulk groups
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool MASKS, int NP, int NB, int LA>
__global__ void __launch_bounds__((NCW + 1) * 32) c3_ws_kernel(const __grid_constant__ Maps maps,
                                                               const float2* __restrict__ G,
                                                               const uint8_t* __restrict__ Gm, float sign,
                                                               int H, int W, unsigned tiles_x, unsigned tiles_per_frame,
                                                               unsigned total_tiles, uint64_t negzero2) {
    // a P stage is released one tile late (its rows double as the staging buffer of the output store)
    static_assert(NP >= NB + LA + 1, "P stages must outlive the box pipeline and the output store");
    using SM = Smem<NP, NB>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);   // no static shared memory in this kernel: the window starts aligned
    if (smem_u32(smem_raw) & 127u) __trap();
    const unsigned tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const unsigned first = blockIdx.x, stride = gridDim.x;
    if (first >= total_tiles) return;
    const unsigned T = (total_tiles - first + stride - 1) / stride;
    constexpr uint32_t P_BYTES = TS * TS * 8 + (MASKS ? TS * TS : 0);
    constexpr uint32_t B_BYTES = BH * BW * 8 + (MASKS ? BH * BMW : 0);

    if (tid == 0) {
        for (int k = 0; k < NP; ++k) { mbar_init(&sm.pfull[k], 1); mbar_init(&sm.pempty[k], NCW); }
        for (int k = 0; k < NB; ++k) { mbar_init(&sm.bfull[k], 1); mbar_init(&sm.bempty[k], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (wrp == NCW) {
        // ------------------------------------------------------------------------------------------ producer warp
        const int tiles_y = (int)(tiles_per_frame / tiles_x);
        TileIter pit, bit;   // cursors of the P loads and of the box preparation
        pit.init(first, stride, (int)tiles_x, tiles_y);
        bit = pit;
        unsigned ps_i = 0, ps_ph = 0;      // P stage / phase of the next P load
        auto issue_p = [&](bool wait_empty) {   // lane 0 only; tiles are issued in order
            const int tx0 = pit.tx * TS, ty0 = pit.ty * TS, n = pit.n;
            pit.advance((int)tiles_x, tiles_y);
            if (wait_empty) mbar_wait(&sm.pempty[ps_i], ps_ph ^ 1);
            mbar_expect_tx(&sm.pfull[ps_i], P_BYTES);
            tma_load_3d(sm.ps[ps_i].p, &maps.p, &sm.pfull[ps_i], tx0, ty0, n);
            if (MASKS) tma_load_3d(sm.ps[ps_i].pm, &maps.pm, &sm.pfull[ps_i], tx0, ty0, n);
            if (++ps_i == NP) { ps_i = 0; ps_ph ^= 1; }
        };
        if (lane == 0)
            for (unsigned k = 0; k < (unsigned)LA && k < T; ++k) issue_p(false);
        unsigned s = 0, s_ph = 0, b = 0, b_ph = 0;
        for (unsigned i = 0; i < T; ++i) {
            const int tx0 = bit.tx * TS, ty0 = bit.ty * TS, n = bit.n;
            bit.advance((int)tiles_x, tiles_y);
            mbar_wait(&sm.pfull[s], s_ph);
            // sample positions along the tile perimeter and two interior rows (lane = column for the rows, lane = row
            // for the two columns): exact for affine fields, an estimate otherwise (consumers verify per pixel)
            const float2* p = sm.ps[s].p;
            float mnx = 1e30f, mxx = -1e30f, mny = 1e30f, mxy = -1e30f;
            auto acc = [&](int r, int c) {
                const int x = tx0 + c, y = ty0 + r;
                if (x < W && y < H) {
                    const float2 v = p[r * TS + c];
                    const float X = __fmaf_rn(sign, v.x, (float)x), Y = __fmaf_rn(sign, v.y, (float)y);
                    mnx = fminf(mnx, X); mxx = fmaxf(mxx, X);
                    mny = fminf(mny, Y); mxy = fmaxf(mxy, Y);
                }
            };
            acc(0, lane); acc(10, lane); acc(21, lane); acc(31, lane);
            acc(lane, 0); acc(lane, 31);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
                mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
                mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
                mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
            }
            if (lane == 0) {
                // integer tap range, clamped to the taps that can contribute: ix in [-1, W-1], iy in [-1, H-1]
                const float lim = 60000.f;
                int x0 = (int)floorf(fminf(fmaxf(mnx, -lim), lim)), x1 = (int)floorf(fminf(fmaxf(mxx, -lim), lim));
                int y0 = (int)floorf(fminf(fmaxf(mny, -lim), lim)), y1 = (int)floorf(fminf(fmaxf(mxy, -lim), lim));
                x0 = max(-1, min(W - 1, x0)); x1 = max(-1, min(W - 1, x1));
                y0 = max(-1, min(H - 1, y0)); y1 = max(-1, min(H - 1, y1));
                // centre the needed range [x0, x1 + 1] in the box (spare margin on both sides for curved flows)
                const int needw = x1 + 2 - x0, needh = y1 + 2 - y0;
                int vx0 = x0 - max(0, (BW - needw) / 2);
                vx0 &= ~1;                                   // 16-byte aligned box start (float2 elements)
                const int by0 = y0 - max(0, (BH - needh) / 2);
                const int mx0 = vx0 & ~15;
                if (i >= NB) mbar_wait(&sm.bempty[b], b_ph ^ 1);
                sm.binfo[b][0] = make_int4(vx0, vx0 - mx0, by0, 0);
                sm.binfo[b][1] = make_int4(tx0, ty0, n, 0);
                mbar_expect_tx(&sm.bfull[b], B_BYTES);
                tma_load_3d(sm.bs[b].v, &maps.gb, &sm.bfull[b], vx0, by0, n);
                if (MASKS) tma_load_3d(sm.bs[b].m, &maps.gmb, &sm.bfull[b], mx0, by0, n);
                if (i + LA < T) issue_p(i + LA >= NP);
            }
            __syncwarp();
            if (++s == NP) { s = 0; s_ph ^= 1; }
            if (++b == NB) { b = 0; b_ph ^= 1; }
        }
        return;
    }

    // ---------------------------------------------------------------------------------------------- consumer warps
    const uint64_t one2 = pack2(1.0f, 1.0f);
    const float lim = 60000.f;
    unsigned s = 0, s_ph = 0, b = 0, b_ph = 0;
    int prev_s = -1;
    for (unsigned i = 0; i < T; ++i) {
        mbar_wait(&sm.bfull[b], b_ph);
        mbar_wait(&sm.pfull[s], s_ph);
        const int4 info = sm.binfo[b][0], tile = sm.binfo[b][1];
        typename SM::PStage& ps = sm.ps[s];
        const typename SM::BStage& bs = sm.bs[b];
        const int tx0 = tile.x, ty0 = tile.y, n = tile.z;
        const float xg = (float)(tx0 + (int)lane);
        float2* prow = ps.p + ((int)wrp * 4 * TS + lane);      // this thread's pixels: prow[j * TS], in place
        uint8_t* mrow = ps.pm + ((int)wrp * 4 * TS + lane);

        uint64_t p[4];
        unsigned pm[4];
        int dx[4], dy[4], fa[4], fb[4];
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            p[j] = *reinterpret_cast<const uint64_t*>(prow + j * TS);
            pm[j] = MASKS ? mrow[j * TS] : 1u;
            const float2 pv = unpack2(p[j]);
            const float X = __fmaf_rn(sign, pv.x, xg), Y = __fmaf_rn(sign, pv.y, (float)(ty0 + (int)wrp * 4 + j));
            int ix, iy;
            quant(X, ix, fa[j]);
            quant(Y, iy, fb[j]);
            dx[j] = ix - info.x;
            dy[j] = iy - info.z;
            // covered by the box and inside the range of the fast quantiser; everything else is decided per pixel
            ok = ok && (unsigned)dx[j] < (unsigned)(BW - 1) && (unsigned)dy[j] < (unsigned)(BH - 1) &&
                 fabsf(X) < lim && fabsf(Y) < lim;
        }
        if (__all_sync(0xffffffffu, ok)) {
            uint64_t t[4][4];
            unsigned m[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t* v = reinterpret_cast<const uint64_t*>(bs.v) + (dy[j] * BW + dx[j]);
                t[j][0] = v[0]; t[j][1] = v[1]; t[j][2] = v[BW]; t[j][3] = v[BW + 1];
                if (MASKS) {
                    const uint8_t* mm = bs.m + (dy[j] * BMW + dx[j] + info.y);
                    m[j][0] = mm[0]; m[j][1] = mm[1]; m[j][2] = mm[BMW]; m[j][3] = mm[BMW + 1];
                }
            }
            // validity first: it consumes the values loaded last, so once it is known every tap of this warp has
            // left the box and the stage can be handed back to the producer while the blend is still running
            unsigned strict[4], dep = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned a = fa[j], bb = fb[j];
                const unsigned ha = min(a, 1u), hb = min(bb, 1u);      // 1 where the right / lower taps have weight
                if (MASKS) {
                    strict[j] = m[j][0] & (m[j][1] | ~ha) & (m[j][2] | ~hb) & (m[j][3] | ~(ha & hb)) & pm[j];
                } else {
                    const int ix = dx[j] + info.x, iy = dy[j] + info.z;
                    strict[j] = (ix >= 0 && iy >= 0 && (ix + 1 < W || a == 0) && (iy + 1 < H || bb == 0)) ? 1u : 0u;
                    strict[j] |= (unsigned)(t[j][3] >> 63) << 8;    // data dependency on the tap loaded last
                }
                dep |= strict[j];
            }
            dep = __reduce_or_sync(0xffffffffu, dep);
            if (lane == 0) mbar_arrive_after(&sm.bempty[b], dep);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned a = fa[j], bb = fb[j];
                const float ffa = (float)a * (1.0f / 32.0f), ffb = (float)bb * (1.0f / 32.0f);
                const uint64_t fa2 = pack2(ffa, ffa), fb2 = pack2(ffb, ffb);
                const uint64_t na2 = add2(one2, fa2 ^ 0x8000000080000000ull), nb2 = add2(one2, fb2 ^ 0x8000000080000000ull);
                uint64_t accv = mul2_nofuse(t[j][0], mul2(na2, nb2), negzero2);
                accv = add2(accv, mul2_nofuse(t[j][1], mul2(fa2, nb2), negzero2));
                accv = add2(accv, mul2_nofuse(t[j][2], mul2(na2, fb2), negzero2));
                accv = add2(accv, mul2_nofuse(t[j][3], mul2(fa2, fb2), negzero2));
                *reinterpret_cast<uint64_t*>(prow + j * TS) = add2(p[j], accv);
                mrow[j * TS] = (uint8_t)(strict[j] & 1u);
            }
        } else {
            // some pixel of this warp is not covered by the box: per-pixel decision
            const size_t fbase = (size_t)n * ((size_t)H * W);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 pv = unpack2(p[j]);
                const float X = __fmaf_rn(sign, pv.x, xg), Y = __fmaf_rn(sign, pv.y, (float)(ty0 + (int)wrp * 4 + j));
                const bool inbox = (unsigned)dx[j] < (unsigned)(BW - 1) && (unsigned)dy[j] < (unsigned)(BH - 1) &&
                                   fabsf(X) < lim && fabsf(Y) < lim;
                float su, sv;
                unsigned strict;
                if (inbox) {
                    const float2* v = bs.v + (dy[j] * BW + dx[j]);
                    const float2 t00 = v[0], t01 = v[1], t10 = v[BW], t11 = v[BW + 1];
                    const int a = fa[j], bb = fb[j];
                    const unsigned za = a == 0, zb = bb == 0;
                    if (MASKS) {
                        const uint8_t* mm = bs.m + (dy[j] * BMW + dx[j] + info.y);
                        strict = mm[0] & (mm[1] | za) & (mm[BMW] | zb) & (mm[BMW + 1] | za | zb);
                    } else {
                        const int ix = dx[j] + info.x, iy = dy[j] + info.z;
                        strict = (ix >= 0 && iy >= 0 && (ix + 1 < W || za) && (iy + 1 < H || zb)) ? 1u : 0u;
                    }
                    const float ffa = (float)a * (1.0f / 32.0f), ffb = (float)bb * (1.0f / 32.0f);
                    const float na = 1.0f - ffa, nb = 1.0f - ffb;
                    const float f00 = __fmul_rn(na, nb), f01 = __fmul_rn(ffa, nb), f10 = __fmul_rn(na, ffb),
                                f11 = __fmul_rn(ffa, ffb);
                    su = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00.x, f00), __fmul_rn(t01.x, f01)),
                                             __fmul_rn(t10.x, f10)), __fmul_rn(t11.x, f11));
                    sv = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00.y, f00), __fmul_rn(t01.y, f01)),
                                             __fmul_rn(t10.y, f10)), __fmul_rn(t11.y, f11));
                } else if (X <= -1.0f || Y <= -1.0f || X >= (float)W || Y >= (float)H) {
                    su = 0.f; sv = 0.f; strict = 0;      // every tap lies outside the frame
                } else {
                    const float4 r = slow_sample(G + fbase, MASKS ? Gm + fbase : nullptr, H, W, X, Y);
                    su = r.x; sv = r.y; strict = r.z != 0.f;
                }
                prow[j * TS] = make_float2(__fadd_rn(pv.x, su), __fadd_rn(pv.y, sv));
                mrow[j * TS] = (uint8_t)(pm[j] & strict & 1u);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.bempty[b]);
        }
        // the warp's 4 result rows go out as two bulk tensor stores (clipped at the frame border by the hardware)
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
#ifndef C3_NO_STORE
            tma_store_3d(&maps.ov, ps.p + (int)wrp * 4 * TS, tx0, ty0 + (int)wrp * 4, n);
            tma_store_3d(&maps.om, ps.pm + (int)wrp * 4 * TS, tx0, ty0 + (int)wrp * 4, n);
#endif
            bulk_commit();
            if (prev_s >= 0) {
                bulk_wait_read<1>();                 // the previous tile's rows have been read out of shared memory
                mbar_arrive(&sm.pempty[prev_s]);
            }
        }
        prev_s = (int)s;
        if (++s == NP) { s = 0; s_ph ^= 1; }
        if (++b == NB) { b = 0; b_ph ^= 1; }
    }
    if (lane == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}
static inline bool make_map3(CUtensorMap* map, const void* base, int esize, size_t row_elems, size_t H, size_t N,
                             int box_w, int box_h) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    CUtensorMapDataType dt = esize == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    cuuint64_t dims[3] = {row_elems, H, N};
    cuuint64_t strides[2] = {row_elems * esize, row_elems * esize * H};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15)) return false;
    return fn(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns 1 launched, 0 not eligible, <0 CUDA error code (negated)
template <int NP, int NB, int LA>
static inline int launch(const float* P, const uint8_t* Pm, const float* G, const uint8_t* Gm, float sign, float* out,
                         uint8_t* omask, int N, int H, int W, int ctas_per_sm, int sms, cudaStream_t st) {
    const bool masks = Pm != nullptr && Gm != nullptr;
    if ((Pm == nullptr) != (Gm == nullptr)) return 0;
    if (W % 16 != 0) return 0;   // 16-byte row pitch of the uint8 tensors (output mask always, input masks if given)
    Maps maps;
    if (!make_map3(&maps.p, P, 8, W, H, N, TS, TS) || !make_map3(&maps.gb, G, 8, W, H, N, BW, BH)) return 0;
    if (masks) {
        if (!make_map3(&maps.pm, Pm, 1, W, H, N, TS, TS) || !make_map3(&maps.gmb, Gm, 1, W, H, N, BMW, BH)) return 0;
    } else {
        maps.pm = maps.gmb = maps.p;
    }
    if (!make_map3(&maps.ov, out, 8, W, H, N, TS, 4) || !make_map3(&maps.om, omask, 1, W, H, N, TS, 4)) return 0;
    const unsigned tx = (W + TS - 1) / TS, ty = (H + TS - 1) / TS;
    if ((double)tx * ty * N >= 4.0e9) return 0;
    const unsigned total = tx * ty * (unsigned)N;
    unsigned grid = (unsigned)(sms * ctas_per_sm);
    if (grid > total) grid = total;
    const size_t smem = sizeof(Smem<NP, NB>);
    cudaError_t e;
    if (masks) {
        e = cudaFuncSetAttribute(c3_ws_kernel<true, NP, NB, LA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
        c3_ws_kernel<true, NP, NB, LA><<<grid, (NCW + 1) * 32, smem, st>>>(maps, (const float2*)G, Gm, sign, H, W, tx, tx * ty,
                                                                          total, 0x8000000080000000ull);
    } else {
        e = cudaFuncSetAttribute(c3_ws_kernel<false, NP, NB, LA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
        c3_ws_kernel<false, NP, NB, LA><<<grid, (NCW + 1) * 32, smem, st>>>(maps, (const float2*)G, Gm, sign, H, W, tx, tx * ty,
                                                                           total, 0x8000000080000000ull);
    }
    e = cudaPeekAtLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace c3ws
