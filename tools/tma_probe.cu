// Standalone bisect of the TMA box load used by tma_kernels.cu. Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void probe(const __grid_constant__ CUtensorMap map, float2* out, int x0, int y0, int n, int BOX) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    float2* s = (float2*)smem;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        if (MODE >= 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (MODE >= 2) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BOX * BOX * 8) : "memory");
            if (MODE >= 3)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(s)), "l"(&map), "r"(smem_u32(&bar)), "r"(x0), "r"(y0), "r"(n) : "memory");
        }
        if (MODE >= 3) {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            }
        }
    }
    if (MODE >= 4) {
        int v = threadIdx.x;
        v = __reduce_min_sync(0xffffffffu, v);
        if (threadIdx.x == 0) out[BOX * BOX].x = (float)v;
    }
    for (int i = threadIdx.x; i < BOX * BOX; i += blockDim.x) out[i] = (MODE >= 3) ? s[i] : make_float2(1.f, 2.f);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE>
int run(const CUtensorMap& map, float2* dout, int BOX) {
    probe<MODE><<<1, 256, BOX * BOX * 8>>>(map, dout, -1, -2, 0, BOX);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", MODE, cudaGetErrorString(e));
    return e != cudaSuccess;
}

int test(int use_f32, int x0, int y0) {
    const int H = 64, W = 96, N = 2, BOX = 48;
    std::vector<float2> h((size_t)N * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = make_float2((float)i, -(float)i);
    float2 *d, *dout;
    CK(cudaMalloc(&d, h.size() * 8));
    CK(cudaMalloc(&dout, (BOX * BOX + 1) * 8));
    CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)(use_f32 ? 2 * W : W), H, N};
    cuuint64_t strides[2] = {(cuuint64_t)W * 8, (cuuint64_t)W * 8 * H};
    cuuint32_t box[3] = {(cuuint32_t)(use_f32 ? 2 * BOX : BOX), BOX, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&map, use_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, d, dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("f32=%d origin (%d,%d): encode -> %d\n", use_f32, x0, y0, (int)r);
    probe<3><<<1, 256, BOX * BOX * 8>>>(map, dout, use_f32 ? 2 * x0 : x0, y0, 1, BOX);
    cudaError_t e = cudaDeviceSynchronize();
    printf("   run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float2> o(BOX * BOX + 1);
    CK(cudaMemcpy(o.data(), dout, o.size() * 8, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r2 = 0; r2 < BOX; ++r2)
        for (int c = 0; c < BOX; ++c) {
            int sr = r2 + y0, sc = c + x0;
            float want = (sr >= 0 && sr < H && sc >= 0 && sc < W) ? (float)(H * W + sr * W + sc) : 0.f;
            if (o[r2 * BOX + c].x != want) bad++;
        }
    printf("   box check: %d mismatches\n", bad);
    return 0;
}

int main(int argc, char** argv) {
    int use_f32 = argc > 1 ? atoi(argv[1]) : 0;
    int x0 = argc > 2 ? atoi(argv[2]) : 0, y0 = argc > 3 ? atoi(argv[3]) : 0;
    return test(use_f32, x0, y0);
}
