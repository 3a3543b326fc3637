// Flow visualisation on the device: Flow.visualise of the reference (flow_class.py:869-951).
//
//   ofk_vis_magnitude : mag[p] = |threshold(flow[p])| exactly as cv2.cartToPolar computes it (float32,
//                       sqrt(fma(x, x, y*y))), plus the maximum of the frame.
//   ofk_kth_smallest  : exact order statistics of a non-negative float32 array by radix selection on the bit patterns
//                       (4 passes of 8 bits; histogram in shared memory, the bin is chosen on the device) -- the two
//                       values numpy.percentile(mag, 99) interpolates between. No sort, no host round trip per pass.
//   ofk_visualise     : hue = cv2's fastAtan2 polynomial (FMA Horner scheme of the AVX2 / AVX-512 code path of the
//                       OpenCV wheels), saturation = clip(mag * 255 / range_max), value 255 (180 on invalid pixels with
//                       show_mask), mask borders = valid pixels with an invalid or out-of-frame 4-neighbour (what
//                       cv2.findContours + drawContours(thickness 1) paint); 'hsv' rounds half-even to uint8, 'rgb' /
//                       'bgr' convert in float64 in numpy's operation order (no contraction: explicit intrinsics).
#include "ofk_common.cuh"

namespace ofk {
namespace vis {

__device__ __forceinline__ float thresholded(float c, float thr) { return (c < thr && c > -thr) ? 0.f : c; }

__device__ __forceinline__ float magnitude(float x, float y) { return __fsqrt_rn(__fmaf_rn(x, x, __fmul_rn(y, y))); }

// cv::fastAtan2 in degrees, vector code path (mathfuncs_core.simd.hpp, v_atan_f32): FMA Horner scheme
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s,
                p7 = -0.04432655554792128f * s;
    const float ax = fabsf(x), ay = fabsf(y);
    const float c = __fdiv_rn(fminf(ax, ay), __fadd_rn(fmaxf(ax, ay), 2.220446049250313e-16f));
    const float cc = __fmul_rn(c, c);
    float a = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmaf_rn(cc, p7, p5), cc, p3), cc, p1), c);
    if (!(ax >= ay)) a = __fsub_rn(90.f, a);
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

__global__ void __launch_bounds__(256) magnitude_kernel(const float2* __restrict__ flow, float thr,
                                                        float* __restrict__ mag, unsigned int* __restrict__ max_bits,
                                                        size_t n) {
    unsigned int m = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 f = __ldg(flow + i);
        const float v = magnitude(thresholded(f.x, thr), thresholded(f.y, thr));
        mag[i] = v;
        m = max(m, __float_as_uint(v));   // non-negative floats order like their bit patterns
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(max_bits, m);
}

// state[q] = {prefix bits found so far, rank still to resolve inside the prefix}
struct SelectState {
    unsigned int prefix;
    unsigned int pad;
    unsigned long long rank;
};

__global__ void __launch_bounds__(256) select_hist(const unsigned int* __restrict__ bits, size_t n,
                                                   const SelectState* __restrict__ state, int nq, int shift,
                                                   unsigned int* __restrict__ hist /* [nq][256] */) {
    __shared__ unsigned int sh[4][256];
    for (int q = 0; q < nq; ++q) sh[q][threadIdx.x] = 0;
    __syncthreads();
    unsigned int prefix[4];
    for (int q = 0; q < nq; ++q) prefix[q] = state[q].prefix;
    const unsigned int himask = shift == 24 ? 0u : ~0u << (shift + 8);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned int b = bits[i];
        for (int q = 0; q < nq; ++q)
            if ((b & himask) == prefix[q]) atomicAdd(&sh[q][(b >> shift) & 255u], 1u);
    }
    __syncthreads();
    for (int q = 0; q < nq; ++q)
        if (sh[q][threadIdx.x]) atomicAdd(&hist[q * 256 + threadIdx.x], sh[q][threadIdx.x]);
}

__global__ void select_pick(SelectState* __restrict__ state, int nq, int shift, unsigned int* __restrict__ hist,
                            float* __restrict__ out) {
    const int q = threadIdx.x;
    if (q >= nq) return;
    unsigned long long rank = state[q].rank;
    unsigned int bin = 255;
    for (unsigned int k = 0; k < 256; ++k) {
        const unsigned int c = hist[q * 256 + k];
        if (rank < c) { bin = k; break; }
        rank -= c;
    }
    state[q].prefix |= bin << shift;
    state[q].rank = rank;
    for (unsigned int k = 0; k < 256; ++k) hist[q * 256 + k] = 0;
    if (shift == 0) out[q] = __uint_as_float(state[q].prefix);
}

__global__ void select_init(SelectState* __restrict__ state, const unsigned long long* __restrict__ ranks, int nq,
                            unsigned int* __restrict__ hist) {
    for (int k = threadIdx.x; k < nq * 256; k += blockDim.x) hist[k] = 0;
    if ((int)threadIdx.x < nq) {
        state[threadIdx.x].prefix = 0;
        state[threadIdx.x].pad = 0;
        state[threadIdx.x].rank = ranks[threadIdx.x];
    }
}

__device__ __forceinline__ uint8_t round_u8(float v) { return (uint8_t)__float2int_rn(v); }   // half-even; 0 <= v <= 255
__device__ __forceinline__ uint8_t round_u8(double v) { return (uint8_t)__double2int_rn(v); }

__global__ void __launch_bounds__(256) colourise_kernel(const float2* __restrict__ flow,
                                                        const uint8_t* __restrict__ mask, float thr, int mode,
                                                        int show_mask, int show_borders, float range_max,
                                                        uint8_t* __restrict__ out, int H, int W) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t p = (size_t)y * W + x;
    const float2 f = __ldg(flow + p);
    const float fx = thresholded(f.x, thr), fy = thresholded(f.y, thr);
    const float mag = magnitude(fx, fy);
    const float ang = fast_atan2_deg(fy, fx);
    float h = __fmul_rn(ang >= 360.f ? 0.f : ang, 0.5f);                            // np.mod(ang, 360) / 2
    float s = fminf(fmaxf(__fdiv_rn(__fmul_rn(mag, 255.f), range_max), 0.f), 255.f);   // np.clip(mag * 255 / range_max, 0, 255)
    float v = 255.f;
    const bool valid = mask == nullptr || mask[p] != 0;
    if (show_mask && !valid) v = 180.f;
    if (show_borders && valid && mask != nullptr) {
        const bool inner = y > 0 && y < H - 1 && x > 0 && x < W - 1 && mask[p - W] && mask[p + W] && mask[p - 1] && mask[p + 1];
        if (!inner) h = s = v = 0.f;
    } else if (show_borders && mask == nullptr) {
        if (y == 0 || y == H - 1 || x == 0 || x == W - 1) h = s = v = 0.f;
    }
    uint8_t* o = out + p * 3;
    if (mode == 0) {
        o[0] = round_u8(h); o[1] = round_u8(s); o[2] = round_u8(v);
        return;
    }
    // hsv -> rgb in numpy's order of operations (flow_class.py:930-945); float32 until `f = h * 6. - i`, float64 after
    const float hn = __fdiv_rn(h, 180.f), sn = __fdiv_rn(s, 255.f), vn = __fdiv_rn(v, 255.f);
    const float h6 = __fmul_rn(hn, 6.f);
    const int i = (int)h6;                                    // np.int_ truncates
    const double fr = __dsub_rn((double)h6, (double)i);
    const double t = __dsub_rn(1.0, fr);
    const double sd = (double)sn, vd = (double)vn;
    const double c0 = __dmul_rn(__dsub_rn(1.0, __dmul_rn(sd, 0.0)), vd);          // v
    const double c1 = __dmul_rn(__dsub_rn(1.0, sd), vd);                          // p  (s * 1)
    const double c2 = __dmul_rn(__dsub_rn(1.0, __dmul_rn(sd, fr)), vd);           // q
    const double c3 = __dmul_rn(__dsub_rn(1.0, __dmul_rn(sd, t)), vd);            // t
    double r, g, b;
    switch (i % 6) {
        case 0: r = c0; g = c3; b = c1; break;
        case 1: r = c2; g = c0; b = c1; break;
        case 2: r = c1; g = c0; b = c3; break;
        case 3: r = c1; g = c2; b = c0; break;
        case 4: r = c3; g = c1; b = c0; break;
        default: r = c0; g = c1; b = c2; break;
    }
    const uint8_t R = round_u8(__dmul_rn(r, 255.0)), G = round_u8(__dmul_rn(g, 255.0)), B = round_u8(__dmul_rn(b, 255.0));
    if (mode == 1) { o[0] = R; o[1] = G; o[2] = B; }
    else { o[0] = B; o[1] = G; o[2] = R; }
}

}  // namespace vis
}  // namespace ofk

using namespace ofk;

extern "C" int ofk_vis_magnitude(const float* flow, float thr, float* mag, float* max_out, size_t n_pixels,
                                 ofk_stream_t stream) {
    OFK_CHECK_ARG(flow != nullptr && mag != nullptr && max_out != nullptr, "ofk_vis_magnitude: NULL pointer");
    OFK_CHECK_ARG((reinterpret_cast<uintptr_t>(flow) & 7) == 0, "ofk_vis_magnitude: flow must be 8-byte aligned");
    cudaStream_t st = as_stream(stream);
    OFK_CUDA(cudaMemsetAsync(max_out, 0, sizeof(float), st));
    if (n_pixels == 0) return OFK_OK;
    int blocks = (int)((n_pixels + 255) / 256);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    vis::magnitude_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float2*>(flow), thr, mag,
                                                   reinterpret_cast<unsigned int*>(max_out), n_pixels);
    OFK_LAUNCHED();
    return OFK_OK;
}

extern "C" size_t ofk_kth_smallest_workspace(int n_ranks) {
    if (n_ranks < 1 || n_ranks > 4) return 0;
    return sizeof(vis::SelectState) * 4 + sizeof(unsigned int) * 4 * 256 + sizeof(unsigned long long) * 4;
}

extern "C" int ofk_kth_smallest(const float* values, size_t n, const unsigned long long* ranks_host, int n_ranks,
                                float* out, void* ws, size_t ws_bytes, ofk_stream_t stream) {
    OFK_CHECK_ARG(values != nullptr && out != nullptr && ranks_host != nullptr, "ofk_kth_smallest: NULL pointer");
    OFK_CHECK_ARG(n_ranks >= 1 && n_ranks <= 4, "ofk_kth_smallest: 1 to 4 ranks per call, got %d", n_ranks);
    OFK_CHECK_ARG(n > 0, "ofk_kth_smallest: empty array");
    for (int q = 0; q < n_ranks; ++q)
        OFK_CHECK_ARG(ranks_host[q] < n, "ofk_kth_smallest: rank %llu out of range (n = %zu)", ranks_host[q], n);
    OFK_CHECK_ARG(ws != nullptr && ws_bytes >= ofk_kth_smallest_workspace(n_ranks), "ofk_kth_smallest: workspace too small");
    cudaStream_t st = as_stream(stream);
    vis::SelectState* state = static_cast<vis::SelectState*>(ws);
    unsigned int* hist = reinterpret_cast<unsigned int*>(state + 4);
    unsigned long long* d_ranks = reinterpret_cast<unsigned long long*>(hist + 4 * 256);
    OFK_CUDA(cudaMemcpyAsync(d_ranks, ranks_host, sizeof(unsigned long long) * n_ranks, cudaMemcpyHostToDevice, st));
    vis::select_init<<<1, 256, 0, st>>>(state, d_ranks, n_ranks, hist);
    OFK_LAUNCHED();
    int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    if (blocks < 1) blocks = 1;
    for (int shift = 24; shift >= 0; shift -= 8) {
        vis::select_hist<<<blocks, 256, 0, st>>>(reinterpret_cast<const unsigned int*>(values), n, state, n_ranks, shift,
                                                  hist);
        OFK_LAUNCHED();
        vis::select_pick<<<1, 32, 0, st>>>(state, n_ranks, shift, hist, out);
        OFK_LAUNCHED();
    }
    return OFK_OK;
}

extern "C" int ofk_visualise(const float* flow, const uint8_t* mask, float thr, int mode, int show_mask,
                             int show_mask_borders, float range_max, uint8_t* out, int H, int W, ofk_stream_t stream) {
    OFK_CHECK_ARG(flow != nullptr && out != nullptr, "ofk_visualise: NULL pointer");
    OFK_CHECK_ARG(H > 0 && W > 0, "ofk_visualise: bad shape H=%d W=%d", H, W);
    OFK_CHECK_ARG(mode >= 0 && mode <= 2, "ofk_visualise: mode must be 0 (hsv), 1 (rgb) or 2 (bgr)");
    OFK_CHECK_ARG(range_max > 0.f, "ofk_visualise: range_max must be positive");
    OFK_CHECK_ARG((reinterpret_cast<uintptr_t>(flow) & 7) == 0, "ofk_visualise: flow must be 8-byte aligned");
    dim3 grid((W + 31) / 32, (H + 7) / 8);
    vis::colourise_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(flow), mask, thr, mode,
                                                             show_mask, show_mask_borders, range_max, out, H, W);
    OFK_LAUNCHED();
    return OFK_OK;
}
