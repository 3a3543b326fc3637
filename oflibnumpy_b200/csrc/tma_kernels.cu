// TMA-staged variants of the two headline kernels for sm_100a.
//
// The dependent part of a backward warp is the gather: its addresses come from the flow just loaded, so a plain
// gather kernel pays two DRAM round trips in sequence per pixel and spends a large share of its instructions on
// 64-bit addresses and border tests. Here a CTA (32x32 output tile) first reduces the bounding box of its taps, then
// ONE elected thread asks the Tensor Memory Accelerator for that box of the source (cp.async.bulk.tensor into shared
// memory, completion on an mbarrier). Out-of-bounds elements of the box are zero-filled by the hardware, which is
// exactly cv2.remap's BORDER_CONSTANT(0) and "mask invalid outside the frame": the gather then runs from shared
// memory with 32-bit addresses, no bounds tests and no border path. Tiles whose box does not fit (very large flow
// gradients) fall back to the global-memory gather of the *_rows kernels, CTA-uniformly.
#include <cuda.h>
#include <stdlib.h>

#include "combine3_device.cuh"

namespace ofk {

// The TMA-pipelined composition kernel is correct (same parity tests) but, as measured on B200 in round 1, ~10 %
// slower than the register-pipelined gather kernel on the headline workload (profiles/), so it is opt-in:
// OFK_TMA=1 in the environment selects it (A/B measurements).
bool tma_enabled() {
    static int state = -1;
    if (state < 0) {
        const char* e = getenv("OFK_TMA");
        state = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return state == 1;
}

// ------------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// rank-3 map over [N][H][row_elems] of `esize`-byte elements, box [1][box_h][box_w]
bool make_map3(CUtensorMap* map, const void* base, int esize, size_t row_elems, size_t H, size_t N, int box_w,
               int box_h) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    CUtensorMapDataType dt = esize == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64
                                        : (esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
    cuuint64_t dims[3] = {row_elems, H, N};
    cuuint64_t strides[2] = {row_elems * esize, row_elems * esize * H};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15)) return false;
    CUresult r = fn(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ device: PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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

constexpr int BOX = 48;   // box width and height in pixels (tile 32x32 plus the reach of a ~20 degree rotation)
constexpr int BOXM = 64;  // width of byte-element boxes: TMA wants the first box byte 16-byte aligned in its row, so
                          // byte boxes start at a multiple of 16 pixels and need up to 15 extra columns

struct TileBox {
    int x0, x1, y0, y1;  // min / max integer tap (left / top) over the tile, inclusive
};

// CTA-wide reduction of the tap bounding box through warp REDUX + 4 shared atomics per warp
__device__ __forceinline__ void reduce_box(int mnx, int mxx, int mny, int mxy, int* sbox) {
    mnx = __reduce_min_sync(0xffffffffu, mnx);
    mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny);
    mxy = __reduce_max_sync(0xffffffffu, mxy);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&sbox[0], mnx);
        atomicMax(&sbox[1], mxx);
        atomicMin(&sbox[2], mny);
        atomicMax(&sbox[3], mxy);
    }
}

// ------------------------------------------------------------------------------------------------ combine mode 3
// Persistent, TMA-pipelined composition kernel. A CTA walks over 32x32 tiles (grid-stride); per tile there are two
// dependent fetches -- the pointwise operands, then the box of the gathered operand their values point at -- and
// both are bulk tensor copies into shared memory that are issued one tile (box) resp. two to three tiles (pointwise
// stage) ahead of their use:
//
//   iteration i:   wait P-stage(i+1) -> sample coordinates of tile i+1 -> tap bounding box -> TMA box(i+1)
//                  wait box(i)       -> gather + blend + store tile i from shared memory
//                  TMA P-stage(i+3) into the stage tile i just released
//
// so the HBM latency of both fetches overlaps the arithmetic of the previous tiles without holding registers, and the
// inner loop works on 32-bit shared-memory offsets only (hardware zero fill = constant border, invalid mask outside).
constexpr int TS = 32;        // tile edge
constexpr int NSTAGE = 3;     // pointwise stages in flight

struct __align__(128) PStage {
    float2 p[TS * TS];        // pointwise operand
    float2 g[TS * TS];        // gathered operand AT p (zero test only)
    uint8_t pm[TS * TS];
    uint8_t gm[TS * TS];
};
struct __align__(128) GBox {
    float2 v[BOX * BOX];
    uint8_t m[BOX * BOXM];
};
struct PipeSmem {
    PStage st[NSTAGE];
    GBox box[2];
    uint64_t pbar[NSTAGE];
    uint64_t gbar[2];
    int sbox[2][4];
};

struct PipeMaps {
    CUtensorMap p, pm, gt, gmt, gb, gmb;   // pointwise tiles (p, pm, g@p, gm@p) and gather boxes (vecs, mask)
};

template <bool REF_T, bool MASKS>
__global__ void __launch_bounds__(256, 2) combine3_pipe(const __grid_constant__ PipeMaps maps,
                                                        const float* __restrict__ A, const uint8_t* __restrict__ Am,
                                                        const float* __restrict__ B, const uint8_t* __restrict__ Bm,
                                                        float thr, float* __restrict__ out,
                                                        uint8_t* __restrict__ omask, int* __restrict__ flags, int H,
                                                        int W, unsigned tiles_x, unsigned tiles_per_frame,
                                                        unsigned total_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PipeSmem& sm = *reinterpret_cast<PipeSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const unsigned tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const unsigned first = blockIdx.x, stride = gridDim.x;
    if (first >= total_tiles) return;
    const unsigned my_tiles = (total_tiles - first + stride - 1) / stride;
    const float sign = REF_T ? -1.0f : 1.0f;
    const float tz = thr > 0.f ? thr : 1.401298464e-45f;
    constexpr uint32_t P_BYTES = TS * TS * 16 + (MASKS ? TS * TS * 2 : 0);
    constexpr uint32_t G_BYTES = BOX * BOX * 8 + (MASKS ? BOX * BOXM : 0);

    auto tile_origin = [&](unsigned i, int& tx0, int& ty0, int& n) {
        const unsigned t = first + i * stride;
        const unsigned nn = t / tiles_per_frame, r = t - nn * tiles_per_frame;
        const unsigned ty = r / tiles_x, tx = r - ty * tiles_x;
        tx0 = (int)tx * TS; ty0 = (int)ty * TS; n = (int)nn;
    };
    auto issue_stage = [&](unsigned i) {   // one thread
        int tx0, ty0, n;
        tile_origin(i, tx0, ty0, n);
        PStage& s = sm.st[i % NSTAGE];
        uint64_t* bar = &sm.pbar[i % NSTAGE];
        mbar_expect_tx(bar, P_BYTES);
        tma_load_3d(s.p, &maps.p, bar, tx0, ty0, n);
        tma_load_3d(s.g, &maps.gt, bar, tx0, ty0, n);
        if (MASKS) {
            tma_load_3d(s.pm, &maps.pm, bar, tx0, ty0, n);
            tma_load_3d(s.gm, &maps.gmt, bar, tx0, ty0, n);
        }
    };

    if (tid == 0) {
        for (int k = 0; k < NSTAGE; ++k) mbar_init(&sm.pbar[k], 1);
        mbar_init(&sm.gbar[0], 1);
        mbar_init(&sm.gbar[1], 1);
        for (int k = 0; k < 2; ++k) {
            sm.sbox[k][0] = sm.sbox[k][2] = 0x7fffffff;
            sm.sbox[k][1] = sm.sbox[k][3] = -0x7fffffff;
        }
    }
    __syncthreads();
    if (tid == 0)
        for (unsigned k = 0; k < NSTAGE && k < my_tiles; ++k) issue_stage(k);

    // per-thread state of the tile whose box is in flight: packed taps (ix+32768 | iy+32768 << 16), fractions, use bits
    unsigned cur_xy[4], cur_ab[4], nxt_xy[4], nxt_ab[4];
    unsigned cur_use = 0, nxt_use = 0;
    int cur_box[4] = {0, 0, 0, 0}, nxt_box[4] = {0, 0, 0, 0};   // vx0, mx0, by0, fits
    unsigned nz_p = 0, nz_g = 0;
    int flag_n = -1;

    // coordinates + bounding box + box TMA of tile i (reads stage i % NSTAGE); results into nxt_*
    auto prepare = [&](unsigned i) {
        int tx0, ty0, n;
        tile_origin(i, tx0, ty0, n);
        const PStage& s = sm.st[i % NSTAGE];
        mbar_wait(&sm.pbar[i % NSTAGE], (i / NSTAGE) & 1);
        int* sb = sm.sbox[i & 1];
        const int x = tx0 + (int)lane;
        int mnx = 0x7fffffff, mxx = -0x7fffffff, mny = 0x7fffffff, mxy = -0x7fffffff;
        nxt_use = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ly = (int)wrp * 4 + j, y = ty0 + ly;
            const float2 pv = s.p[ly * TS + lane];
            const float X = __fmaf_rn(sign, pv.x, (float)x), Y = __fmaf_rn(sign, pv.y, (float)y);
            const QCoord qx = quantise_fast(X), qy = quantise_fast(Y);
            // taps entirely outside the frame (or beyond the fast quantiser) sample zero / take the exact slow path
            const bool inrange = fabsf(X) < OFK_FAST_COORD_LIMIT && fabsf(Y) < OFK_FAST_COORD_LIMIT;
            const bool use = inrange && qx.i >= -1 && qx.i < W && qy.i >= -1 && qy.i < H;
            nxt_xy[j] = (unsigned)(qx.i + 32768) | ((unsigned)(qy.i + 32768) << 16);
            nxt_ab[j] = (unsigned)qx.f | ((unsigned)qy.f << 8);
            nxt_use |= (use ? 1u : 0u) << j;
            nxt_use |= (inrange ? 0u : 1u) << (4 + j);             // bit 4+j: needs the exact slow path
            if (use && x < W && y < H) {
                mnx = min(mnx, qx.i); mxx = max(mxx, qx.i);
                mny = min(mny, qy.i); mxy = max(mxy, qy.i);
            }
        }
        reduce_box(mnx, mxx, mny, mxy, sb);
        __syncthreads();
        const int bx0 = sb[0], bx1 = sb[1], by0 = sb[2], by1 = sb[3];
        const bool nonempty = bx1 >= bx0;
        const int vx0 = bx0 - (bx0 & 1), mx0 = bx0 - (bx0 & 15);   // TMA: first box byte 16-byte aligned in its row
        const bool fits = nonempty && (bx1 + 2 - vx0 <= BOX) && (bx1 + 2 - mx0 <= BOXM) && (by1 - by0 + 2 <= BOX);
        nxt_box[0] = vx0; nxt_box[1] = mx0; nxt_box[2] = by0; nxt_box[3] = fits ? 1 : (nonempty ? 0 : 2);
        if (tid == 0 && fits) {
            GBox& gb = sm.box[i & 1];
            uint64_t* bar = &sm.gbar[i & 1];
            mbar_expect_tx(bar, G_BYTES);
            tma_load_3d(gb.v, &maps.gb, bar, vx0, by0, n);
            if (MASKS) tma_load_3d(gb.m, &maps.gmb, bar, mx0, by0, n);
        }
    };

    prepare(0);
#pragma unroll
    for (int j = 0; j < 4; ++j) { cur_xy[j] = nxt_xy[j]; cur_ab[j] = nxt_ab[j]; }
    cur_use = nxt_use;
#pragma unroll
    for (int k = 0; k < 4; ++k) cur_box[k] = nxt_box[k];
    __syncthreads();                 // everyone has read bounding-box slot 0
    if (tid == 0) {
        sm.sbox[0][0] = sm.sbox[0][2] = 0x7fffffff;
        sm.sbox[0][1] = sm.sbox[0][3] = -0x7fffffff;
    }
    unsigned uses0 = 0, uses1 = 0;   // completed phases of the two box barriers (a tile without a box skips its phase)

    for (unsigned i = 0; i < my_tiles; ++i) {
        if (i + 1 < my_tiles) prepare(i + 1);      // its __syncthreads also orders the reuse of box (i+1)&1
        // ---------------------------------------------------------------------------------- compute tile i
        int tx0, ty0, n;
        tile_origin(i, tx0, ty0, n);
        const PStage& s = sm.st[i % NSTAGE];
        const GBox& gb = sm.box[i & 1];
        const int fits = cur_box[3];
        if (fits == 1) {
            if (i & 1) { mbar_wait(&sm.gbar[1], uses1 & 1); ++uses1; }
            else { mbar_wait(&sm.gbar[0], uses0 & 1); ++uses0; }
        }
        const size_t fbase = (size_t)n * ((size_t)H * W);
        const int x = tx0 + (int)lane;
        if (flag_n != n) {   // frame changed: publish what was seen for the previous frame
            if (flags != nullptr && flag_n >= 0) {
                const bool fa_ = __any_sync(0xffffffffu, (REF_T ? nz_g : nz_p) != 0);
                const bool fb_ = __any_sync(0xffffffffu, (REF_T ? nz_p : nz_g) != 0);
                if (lane == 0) {
                    if (fa_) flags[flag_n * 2 + 0] = 1;
                    if (fb_) flags[flag_n * 2 + 1] = 1;
                }
            }
            flag_n = n; nz_p = 0; nz_g = 0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ly = (int)wrp * 4 + j, y = ty0 + ly;
            const float2 pv = s.p[ly * TS + lane];
            const float2 gv = s.g[ly * TS + lane];
            const unsigned pmv = MASKS ? s.pm[ly * TS + lane] : ((x < W && y < H) ? 1u : 0u);
            const unsigned gmv = MASKS ? s.gm[ly * TS + lane] : ((x < W && y < H) ? 1u : 0u);
            nz_p |= pmv & (unsigned)(fabsf(pv.x) >= tz || fabsf(pv.y) >= tz);
            nz_g |= gmv & (unsigned)(fabsf(gv.x) >= tz || fabsf(gv.y) >= tz);
            const int ix = (int)(cur_xy[j] & 0xffffu) - 32768, iy = (int)(cur_xy[j] >> 16) - 32768;
            const int a = cur_ab[j] & 0xff, b = cur_ab[j] >> 8;
            float su = 0.f, sv = 0.f;
            unsigned strict = 0;
            if ((cur_use >> j) & 1u) {
                if (fits == 1) {
                    const int o = (iy - cur_box[2]) * BOX + (ix - cur_box[0]);
                    const float2 t00 = gb.v[o], t01 = gb.v[o + 1], t10 = gb.v[o + BOX], t11 = gb.v[o + BOX + 1];
                    const unsigned ha = a != 0, hb = b != 0;
                    if (MASKS) {
                        const int om = (iy - cur_box[2]) * BOXM + (ix - cur_box[1]);
                        const unsigned i00 = gb.m[om] ^ 1u, i01 = gb.m[om + 1] ^ 1u, i10 = gb.m[om + BOXM] ^ 1u,
                                       i11 = gb.m[om + BOXM + 1] ^ 1u;
                        strict = ((i00 | (i01 & ha) | (i10 & hb) | (i11 & ha & hb)) & 1u) ^ 1u;
                    } else {
                        strict = (ix >= 0 && iy >= 0 && (ix + 1 < W || !ha) && (iy + 1 < H || !hb)) ? 1u : 0u;
                    }
                    const float ffa = (float)a * (1.0f / 32.0f), ffb = (float)b * (1.0f / 32.0f);
                    const float na = 1.0f - ffa, nb = 1.0f - ffb;
                    const float f00 = __fmul_rn(na, nb), f01 = __fmul_rn(ffa, nb), f10 = __fmul_rn(na, ffb),
                                f11 = __fmul_rn(ffa, ffb);
                    su = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00.x, f00), __fmul_rn(t01.x, f01)),
                                             __fmul_rn(t10.x, f10)), __fmul_rn(t11.x, f11));
                    sv = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00.y, f00), __fmul_rn(t01.y, f01)),
                                             __fmul_rn(t10.y, f10)), __fmul_rn(t11.y, f11));
                } else {   // the box of this tile did not fit: exact global-memory path
                    const float2* G = reinterpret_cast<const float2*>(REF_T ? A : B) + fbase;
                    const uint8_t* Gm = MASKS ? (REF_T ? Am : Bm) + fbase : nullptr;
                    const SampleResult r = sample_flow_border(G, Gm, H, W, __fmaf_rn(sign, pv.x, (float)x),
                                                              __fmaf_rn(sign, pv.y, (float)y));
                    su = r.u; sv = r.v; strict = (unsigned)r.strict;
                }
            } else if ((cur_use >> (4 + j)) & 1u) {   // beyond the fast quantiser's range: exact slow path
                const float2* G = reinterpret_cast<const float2*>(REF_T ? A : B) + fbase;
                const uint8_t* Gm = MASKS ? (REF_T ? Am : Bm) + fbase : nullptr;
                const SampleResult r = sample_flow_border(G, Gm, H, W, __fmaf_rn(sign, pv.x, (float)x),
                                                          __fmaf_rn(sign, pv.y, (float)y));
                su = r.u; sv = r.v; strict = (unsigned)r.strict;
            }
            if (x < W && y < H) {
                const size_t idx = fbase + (size_t)y * W + x;
                float2 o;
                o.x = __fadd_rn(pv.x, su);
                o.y = __fadd_rn(pv.y, sv);
                asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(reinterpret_cast<float2*>(out) + idx),
                             "f"(o.x), "f"(o.y) : "memory");
                omask[idx] = (uint8_t)((MASKS ? pmv : 1u) & strict);
            }
        }
        __syncthreads();   // every thread is done with stage i % NSTAGE, box i & 1 and sbox slot i & 1
        if (tid == 0) {
            // slot (i+1)&1 was used by prepare(i+1) above and is next used by prepare(i+3), two barriers from here
            sm.sbox[(i + 1) & 1][0] = sm.sbox[(i + 1) & 1][2] = 0x7fffffff;
            sm.sbox[(i + 1) & 1][1] = sm.sbox[(i + 1) & 1][3] = -0x7fffffff;
            if (i + NSTAGE < my_tiles) issue_stage(i + NSTAGE);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { cur_xy[j] = nxt_xy[j]; cur_ab[j] = nxt_ab[j]; }
        cur_use = nxt_use;
#pragma unroll
        for (int k = 0; k < 4; ++k) cur_box[k] = nxt_box[k];
    }
    if (flags != nullptr && flag_n >= 0) {
        const bool fa_ = __any_sync(0xffffffffu, (REF_T ? nz_g : nz_p) != 0);
        const bool fb_ = __any_sync(0xffffffffu, (REF_T ? nz_p : nz_g) != 0);
        if (lane == 0) {
            if (fa_) flags[flag_n * 2 + 0] = 1;
            if (fb_) flags[flag_n * 2 + 1] = 1;
        }
    }
}

// Returns 1 if the TMA kernel was launched, 0 if the configuration is not eligible (caller uses the rows kernel),
// negative on error.
int launch_combine3_tma(const float* A, const uint8_t* Am, const float* B, const uint8_t* Bm, int ref, float thr,
                        float* out, uint8_t* out_mask, int* flags, int N, int H, int W, cudaStream_t st) {
    const bool masks = Am != nullptr && Bm != nullptr;
    if ((Am == nullptr) != (Bm == nullptr)) return 0;
    if (W % 2 != 0 || (masks && (W % 16 != 0 || ((size_t)H * W) % 16 != 0))) return 0;
    if (H >= 32768 || W >= 32768 || (size_t)H * W >= ((size_t)1 << 30)) return 0;
    const float* P = ref == 't' ? B : A;
    const uint8_t* Pm = ref == 't' ? Bm : Am;
    const float* G = ref == 't' ? A : B;
    const uint8_t* Gm = ref == 't' ? Am : Bm;
    PipeMaps maps;
    if (!make_map3(&maps.p, P, 8, W, H, N, TS, TS) || !make_map3(&maps.gt, G, 8, W, H, N, TS, TS) ||
        !make_map3(&maps.gb, G, 8, W, H, N, BOX, BOX))
        return 0;
    if (masks) {
        if (!make_map3(&maps.pm, Pm, 1, W, H, N, TS, TS) || !make_map3(&maps.gmt, Gm, 1, W, H, N, TS, TS) ||
            !make_map3(&maps.gmb, Gm, 1, W, H, N, BOXM, BOX))
            return 0;
    } else {
        maps.pm = maps.gmt = maps.gmb = maps.p;
    }
    const unsigned tx = (W + TS - 1) / TS, ty = (H + TS - 1) / TS;
    const double total_d = (double)tx * ty * N;
    if (total_d >= 4.0e9) return 0;
    const unsigned total = tx * ty * (unsigned)N;
    unsigned grid = (unsigned)sm_count() * 2;
    if (grid > total) grid = total;
    const size_t smem = sizeof(PipeSmem) + 128;
#define OFK_C3(RT, MK)                                                                                              \
    do {                                                                                                            \
        static bool attr_done = false;                                                                              \
        if (!attr_done) {                                                                                           \
            if (cudaFuncSetAttribute(combine3_pipe<RT, MK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != \
                cudaSuccess) {                                                                                      \
                cudaGetLastError();                                                                                 \
                return 0;                                                                                           \
            }                                                                                                       \
            attr_done = true;                                                                                       \
        }                                                                                                           \
        combine3_pipe<RT, MK><<<grid, 256, smem, st>>>(maps, A, Am, B, Bm, thr, out, out_mask, flags, H, W, tx,     \
                                                       tx * ty, total);                                             \
    } while (0)
    if (ref == 't') {
        if (masks) OFK_C3(true, true);
        else OFK_C3(true, false);
    } else {
        if (masks) OFK_C3(false, true);
        else OFK_C3(false, false);
    }
#undef OFK_C3
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error("combine3_pipe launch failed: %s", cudaGetErrorString(e));
        return OFK_ECUDA;
    }
    return 1;
}

}  // namespace ofk
