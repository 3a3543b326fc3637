// Flow composition, modes 1 and 2 (flow_class.py:1357-1410 of the reference), as device-resident chains for a batch of
// N frame pairs behind ONE entry point: no host round trip between the warps, additions and mask-ANDs the reference
// spells out as Flow methods, every intermediate in the caller's workspace.
//
//   mode 2 's':  A.apply(B - A)                                                   1 forward resampling
//   mode 2 't':  B - resample(A from grid - A to grid - B), mask > .99             1 launch (ofk_combine2_t)
//   mode 1 's':  Bi = -B as 't';  B - (Bi + Bi.apply(A.switch_ref())).apply(A)     1 forward, 1 fused backward + add,
//                                                                                 1 backward warp of A
//   mode 1 't':  As = A.switch_ref(); (B.switch_ref() - (As + (-As as 't').apply(-B as 's')).apply(As)).switch_ref()
//                                                                                 4 forward, 1 fused backward + add
//
// `flow + flow.apply(other)` / `self + self.invert('t').apply(other)` inside mode 1 are exactly the fused mode-3
// kernel (ofk_combine3 without its zero tests). The forward resamplings are bound by float64 geometry, not by HBM
// (DESIGN.md), so the intermediates between them live in HBM; what is fused is everything that is free: negations into
// the sample sign where the kernels take one, subtraction + threshold of mode 2 't' into the resampler's epilogue.
//
// Zero-flow semantics of the reference inside the chains are kept per frame on the device: `switch_ref` relabels a
// flow that is exactly zero on its valid pixels (flow_class.py:717-718) and `apply_flow` returns its target for a flow
// that is zero below the threshold (utils.py:215-216); both end in "payload passes through", decided by two zero
// tests whose flags feed ofk_forward_s_ex. The early exits of combine_with itself (:1338-1354) return an OPERAND
// OBJECT and are the caller's business (the Python layer reads ofk_nonzero_flags first).
#include "ofk_common.cuh"

namespace ofk {
namespace c12 {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Ws {
    unsigned char* base;
    size_t used, cap;
    void* take(size_t bytes) {
        void* p = base + used;
        used = align_up(used + bytes, 256);
        return p;
    }
};

struct Dims {
    int N, H, W;
    size_t px() const { return (size_t)N * H * W; }
};

__global__ void and_flags_kernel(const int* a, const int* b, int* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (a[i] != 0 && b[i] != 0) ? 1 : 0;
}

// Flow.apply of an 's' flow F to a flow T (flow_class.py:600-670 with consider_mask = True): the payload mask is
// T.mask & F.mask, the points of F with a 0 mask are removed. `relabel_exit`: the call comes from switch_ref, which
// relabels instead of resampling when F is exactly zero on its valid pixels.
static int apply_s(const float* F, const uint8_t* Fm, float sign, const float* T, const uint8_t* Tm, bool relabel_exit,
                   float* out, uint8_t* out_mask, const Dims& d, Ws fws, int* flags /* 3 * N ints */, cudaStream_t st) {
    ofk_stream_t s = reinterpret_cast<ofk_stream_t>(st);
    uint8_t* pm = static_cast<uint8_t*>(fws.take(d.px()));
    int rc = ofk_mask_and(Tm, Fm, pm, d.px(), s);
    if (rc != OFK_OK) return rc;
    int* nz_thr = flags;
    rc = ofk_nonzero_flags(F, nullptr, 1e-3f, nz_thr, d.N, d.H, d.W, s);      // apply_flow: thresholded, unmasked
    if (rc != OFK_OK) return rc;
    const int* active = nz_thr;
    if (relabel_exit) {
        int* nz_exact = flags + d.N;
        rc = ofk_nonzero_flags(F, Fm, 0.0f, nz_exact, d.N, d.H, d.W, s);      // switch_ref: exact, masked
        if (rc != OFK_OK) return rc;
        and_flags_kernel<<<(d.N + 255) / 256, 256, 0, st>>>(nz_thr, nz_exact, flags + 2 * d.N, d.N);
        OFK_LAUNCHED();
        active = flags + 2 * d.N;
    }
    const size_t fbytes = ofk_forward_s_workspace(d.N, d.H, d.W);
    void* w = fws.take(fbytes);
    return ofk_forward_s_ex(T, 2, F, sign, pm, Fm, active, out, out_mask, OFK_RULE_STRICT, d.N, d.H, d.W, w, fbytes, s);
}

}  // namespace c12
}  // namespace ofk

using namespace ofk;
using namespace ofk::c12;

extern "C" size_t ofk_combine12_workspace(int mode, int ref, int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    if (mode == 2 && ref == 't') return 256;
    const size_t px = (size_t)N * H * W;
    const size_t vec = align_up(px * 8, 256), msk = align_up(px, 256);
    const size_t fwd = align_up(ofk_forward_s_workspace(N, H, W), 256) + msk + 256;   // + payload-mask plane
    const size_t flags = align_up(sizeof(int) * 3 * (size_t)N, 256);
    // ones masks (2), up to 6 flow intermediates with masks
    return 2 * msk + 6 * (vec + msk) + fwd + flags + 4096;
}

extern "C" int ofk_combine12(int mode, int ref, const float* A, const uint8_t* Am, const float* B, const uint8_t* Bm,
                             float* out, uint8_t* out_mask, int N, int H, int W, void* ws, size_t ws_bytes,
                             ofk_stream_t stream) {
    OFK_CHECK_ARG(mode == 1 || mode == 2, "ofk_combine12: mode must be 1 or 2, got %d", mode);
    OFK_CHECK_ARG(ref == 's' || ref == 't', "ofk_combine12: ref must be 's' or 't', got %d", ref);
    OFK_CHECK_ARG(A && B && out && out_mask, "ofk_combine12: NULL operand");
    OFK_CHECK_ARG(N >= 0 && H > 1 && W > 1, "ofk_combine12: bad shape N=%d H=%d W=%d", N, H, W);
    if (N == 0) return OFK_OK;
    if (mode == 2 && ref == 't') return ofk_combine2_t(A, Am, B, Bm, out, out_mask, N, H, W, stream);
    const size_t need = ofk_combine12_workspace(mode, ref, N, H, W);
    OFK_CHECK_ARG(ws != nullptr && ws_bytes >= need, "ofk_combine12: workspace of %zu bytes needed, got %zu", need,
                  ws_bytes);
    cudaStream_t st = as_stream(stream);
    const Dims d{N, H, W};
    const size_t px = d.px();
    Ws w{static_cast<unsigned char*>(ws), 0, ws_bytes};
    if (Am == nullptr) {
        uint8_t* m = static_cast<uint8_t*>(w.take(px));
        OFK_CUDA(cudaMemsetAsync(m, 1, px, st));
        Am = m;
    }
    if (Bm == nullptr) {
        uint8_t* m = static_cast<uint8_t*>(w.take(px));
        OFK_CUDA(cudaMemsetAsync(m, 1, px, st));
        Bm = m;
    }
    int* flags = static_cast<int*>(w.take(sizeof(int) * 3 * (size_t)N));
    auto vecs = [&]() { return static_cast<float*>(w.take(px * 8)); };
    auto mask = [&]() { return static_cast<uint8_t*>(w.take(px)); };
    int rc;
#define OFK_TRY(call)             \
    do {                          \
        rc = (call);              \
        if (rc != OFK_OK) return rc; \
    } while (0)

    if (mode == 2) {   // ref 's': A.apply(B - A)
        float* D = vecs();
        uint8_t* Dm = mask();
        OFK_TRY(ofk_addsub(OFK_OP_SUB, B, Bm, A, Am, D, Dm, N, H, W, stream));
        return apply_s(A, Am, 1.0f, D, Dm, false, out, out_mask, d, w, flags, st);
    }
    if (ref == 's') {
        // flow_inv_t = flow.invert('t');  flow - (flow_inv_t + flow_inv_t.apply(self.switch_ref())).apply(self)
        float* Bi = vecs();
        OFK_TRY(ofk_scale(OFK_OP_MUL, B, -1.0, -1.0, 0, Bi, px, stream));
        float* As = vecs();
        uint8_t* Asm = mask();
        OFK_TRY(apply_s(A, Am, 1.0f, A, Am, true, As, Asm, d, w, flags, st));                 // A.switch_ref(): s -> t
        float* U = vecs();
        uint8_t* Um = mask();
        OFK_TRY(ofk_combine3(As, Asm, Bi, Bm, 't', 0.0f, U, Um, nullptr, N, H, W, stream));   // Bi + Bi.apply(As)
        float* V = vecs();
        uint8_t* Vm = mask();
        // (...).apply(self): the 't' flow U warps self (vectors and mask) backwards
        OFK_TRY(ofk_warp_t(A, OFK_F32, 2, OFK_ARITH_NATIVE, U, -1.0f, Am, Um, V, Vm, OFK_RULE_STRICT, N, H, W, H, W, 0, 0,
                           1, stream));
        return ofk_addsub(OFK_OP_SUB, B, Bm, V, Vm, out, out_mask, N, H, W, stream);
    }
    // ref 't':  self_s = self.switch_ref()
    //           result = flow.switch_ref() - (self_s + self_s.invert(ref='t').apply(flow.invert('s'))).apply(self_s)
    //           result.switch_ref()
    float* As = vecs();
    uint8_t* Asm = mask();
    OFK_TRY(apply_s(A, Am, -1.0f, A, Am, true, As, Asm, d, w, flags, st));      // t -> s: (-f).apply(f), f = A as 's'
    float* Bs = vecs();
    uint8_t* Bsm = mask();
    OFK_TRY(apply_s(B, Bm, -1.0f, B, Bm, true, Bs, Bsm, d, w, flags, st));
    float* Y = vecs();                                                           // flow.invert('s') = -B
    OFK_TRY(ofk_scale(OFK_OP_MUL, B, -1.0, -1.0, 0, Y, px, stream));
    float* Wv = vecs();
    uint8_t* Wm = mask();
    OFK_TRY(ofk_combine3(As, Asm, Y, Bm, 's', 0.0f, Wv, Wm, nullptr, N, H, W, stream));   // As + As.invert('t').apply(Y)
    float* V = vecs();
    uint8_t* Vm = mask();
    OFK_TRY(apply_s(Wv, Wm, 1.0f, As, Asm, false, V, Vm, d, w, flags, st));               // (...).apply(self_s): 's' flow W
    // R = Bs - V may overwrite Y / W: both are dead
    float* R = Y;
    uint8_t* Rm = Wm;
    OFK_TRY(ofk_addsub(OFK_OP_SUB, Bs, Bsm, V, Vm, R, Rm, N, H, W, stream));
    return apply_s(R, Rm, 1.0f, R, Rm, true, out, out_mask, d, w, flags, st);             // s -> t
#undef OFK_TRY
}
