// Plumbing shared by the warp-specialised TMA kernels (combine3_ws.cu, warp_t_ws.cu): mbarrier / bulk-tensor PTX
// wrappers, the persistent tile iterator and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "ofk_common.cuh"

namespace ofk {
namespace ws {

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Arrive once `dep` has been computed: ties the release of a buffer to the registers loaded from it. The dependency
// has to be real for ptxas, not just for the PTX text: `dep` is stored to a scratch word in shared memory first, and the
// arrive (release semantics) is ordered after that store -- which cannot issue before every load feeding `dep` has
// returned. (A version that only mentioned `dep` in the asm string let ptxas schedule the arrive ahead of the last
// shared-memory loads in the out-of-line path: the producer's next TMA load then overwrote taps still being read.)
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, unsigned dep, uint32_t* sink) {
    asm volatile(
        "st.shared.u32 [%2], %1;\n\t"
        "mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)), "r"(dep), "r"(smem_u32(sink))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    // try_wait suspends the warp until the phase completes or the hint (ns) expires, so a long hint means few
    // wake-ups (issue slots) while waiting and no extra latency when the data arrives
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase), "r"(4000u)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}


// bulk tensor store shared -> global (clipped at the tensor bounds), tracked by the issuing thread's bulk groups
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d_u(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// identity the compiler cannot move: what is computed from the result stays in the (rare) branch that asks for it
// instead of being hoisted into the common path
__device__ __forceinline__ int cold_value(int v) {
    int r;
    asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


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

#ifdef OFK_TRACE   // timeline instrumentation for tools/exp_trace.cu (never defined in the library build)
static __device__ unsigned long long* g_trace = nullptr;   // [tile][16] global-timer stamps of CTA 0
__device__ __forceinline__ void trace(unsigned tile, int slot) {
    if (blockIdx.x == 0 && g_trace != nullptr && tile < 512) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_trace[tile * 16 + slot] = t;
    }
}
#define OFK_TR(tile, slot) trace(tile, slot)
#else
#define OFK_TR(tile, slot)
#endif

#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------ host side
// Test hook: OFK_WS_MAX_CTAS=n caps the persistent grids, so that small frames already give every CTA many tiles (stage
// rings wrap, barrier phases flip) -- tests/test_gpu_warp_t.py runs the parity suite that way in a subprocess.
static inline unsigned cap_ctas(unsigned grid) {
    static int cap = -1;
    if (cap < 0) {
        const char* e = getenv("OFK_WS_MAX_CTAS");
        cap = e ? atoi(e) : 0;
    }
    return (cap > 0 && grid > (unsigned)cap) ? (unsigned)cap : grid;
}

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
// rank-3 map over [N][H][row_elems] of esize-byte elements, box [1][box_h][box_w]
static inline bool make_map3(CUtensorMap* map, const void* base, int esize, size_t row_elems, size_t H, size_t N, int box_w,
                      int box_h) {
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


}  // namespace ws
}  // namespace ofk
