// Affine / projective flow-field generator for sm_100a (utils.py:91-111,319-344 of the reference).
//
// out[y,x] = sign * float32( proj(M [x,y,1]) - [x,y] ) in float64, reproducing the rounding sequence of the
// reference's `np.matmul(M, grid[..., None])` (OpenBLAS dgemv: t = m1*y; t = fma(m0, x, t); r = t + m2), then the
// perspective divide, the subtraction of the float32 grid and the cast to float32. Write-only: 8 B/px.
#include "ofk_common.cuh"

namespace ofk {

struct Mat3 {
    double m[9];
};
constexpr int kMatsPerLaunch = 32;   // host matrices travel as kernel parameters: 32 x 72 B
constexpr int kRowsPerBlock = 4;     // rows per CTA: amortises the matrix fetch and the CTA launch over 4 x W pixels
struct MatBatch {
    Mat3 mats[kMatsPerLaunch];
};

// AFFINE: last matrix row is exactly (0, 0, 1), so tz == 1.0 exactly and the two float64 divisions (the bulk of the
// work: the kernel is otherwise 8 B/px write-only) are the identity.
template <bool AFFINE>
__device__ __forceinline__ void eval_pixel(const double* __restrict__ m, int x, int y, float sign, float& u, float& v) {
    const double gx = static_cast<double>(x), gy = static_cast<double>(y);
    const double tx = __dadd_rn(__fma_rn(m[0], gx, __dmul_rn(m[1], gy)), m[2]);
    const double ty = __dadd_rn(__fma_rn(m[3], gx, __dmul_rn(m[4], gy)), m[5]);
    if (AFFINE) {
        u = sign * __double2float_rn(__dsub_rn(tx, gx));
        v = sign * __double2float_rn(__dsub_rn(ty, gy));
    } else {
        const double tz = __dadd_rn(__fma_rn(m[6], gx, __dmul_rn(m[7], gy)), m[8]);
        u = sign * __double2float_rn(__dsub_rn(__ddiv_rn(tx, tz), gx));
        v = sign * __double2float_rn(__dsub_rn(__ddiv_rn(ty, tz), gy));
    }
}

template <bool AFFINE>
__device__ __forceinline__ void fill_row(const double* __restrict__ m, float* __restrict__ row, int y, float sign, int W,
                                         int vec_ok) {
    if (vec_ok) {
        for (int x = (blockIdx.x * blockDim.x + threadIdx.x) * 2; x < W; x += gridDim.x * blockDim.x * 2) {
            float u0, v0, u1, v1;
            eval_pixel<AFFINE>(m, x, y, sign, u0, v0);
            eval_pixel<AFFINE>(m, x + 1, y, sign, u1, v1);
            st_stream_f4(reinterpret_cast<float4*>(row + (size_t)x * 2), make_float4(u0, v0, u1, v1));
        }
    } else {
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
            float u, v;
            eval_pixel<AFFINE>(m, x, y, sign, u, v);
            row[(size_t)x * 2] = u;
            row[(size_t)x * 2 + 1] = v;
        }
    }
}

// 2 pixels per thread -> one 16-byte store; a CTA walks kRowsPerBlock rows with a grid-stride loop over "pixel pairs".
__global__ void __launch_bounds__(256) from_matrix_kernel(const double* __restrict__ mats_dev, MatBatch host_mats,
                                                          int use_host, int n0, float sign, float* __restrict__ out,
                                                          int H, int W, int vec_ok) {
    const int n = blockIdx.z;
    double m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = use_host ? host_mats.mats[n].m[k] : mats_dev[(size_t)(n0 + n) * 9 + k];
    const bool affine = m[6] == 0.0 && m[7] == 0.0 && m[8] == 1.0;                                  // block-uniform
    for (int r = 0; r < kRowsPerBlock; ++r) {
        const int y = blockIdx.y * kRowsPerBlock + r;
        if (y >= H) break;
        float* row = out + (((size_t)(n0 + n) * H + y) * W) * 2;
        if (affine) fill_row<true>(m, row, y, sign, W, vec_ok);
        else fill_row<false>(m, row, y, sign, W, vec_ok);
    }
}

}  // namespace ofk

using namespace ofk;

extern "C" int ofk_from_matrix(const double* mats, int mats_on_host, float sign, float* out, int N, int H, int W,
                               ofk_stream_t stream) {
    OFK_CHECK_ARG(mats && out, "ofk_from_matrix: NULL argument");
    OFK_CHECK_ARG(N >= 0 && H > 0 && W > 0, "ofk_from_matrix: bad shape N=%d H=%d W=%d", N, H, W);
    OFK_CHECK_ARG(sign == 1.0f || sign == -1.0f, "ofk_from_matrix: sign must be +1 or -1");
    OFK_CHECK_ARG(H <= 65535 * kRowsPerBlock, "ofk_from_matrix: H=%d too large", H);
    OFK_CHECK_ARG(!mats_on_host || N <= 64, "ofk_from_matrix: host matrices limited to N <= 64 (got %d)", N);
    if (N == 0) return OFK_OK;
    cudaStream_t st = as_stream(stream);
    const int vec_ok = (W % 2 == 0) && aligned16(out);
    const int per_thread = vec_ok ? 2 : 1;
    int bx = (W + 256 * per_thread - 1) / (256 * per_thread);
    if (bx < 1) bx = 1;
    const int by = (H + kRowsPerBlock - 1) / kRowsPerBlock;
    if (mats_on_host) {
        for (int n0 = 0; n0 < N; n0 += kMatsPerLaunch) {
            MatBatch mb;
            const int cnt = (N - n0 < kMatsPerLaunch) ? (N - n0) : kMatsPerLaunch;
            for (int i = 0; i < cnt; ++i)
                for (int k = 0; k < 9; ++k) mb.mats[i].m[k] = mats[(size_t)(n0 + i) * 9 + k];
            from_matrix_kernel<<<dim3(bx, by, cnt), 256, 0, st>>>(nullptr, mb, 1, n0, sign, out, H, W, vec_ok);
            OFK_LAUNCHED();
        }
    } else {
        OFK_CHECK_ARG(N <= 65535, "ofk_from_matrix: N=%d exceeds 65535", N);
        MatBatch mb = {};
        from_matrix_kernel<<<dim3(bx, by, N), 256, 0, st>>>(mats, mb, 0, 0, sign, out, H, W, vec_ok);
        OFK_LAUNCHED();
    }
    return OFK_OK;
}
