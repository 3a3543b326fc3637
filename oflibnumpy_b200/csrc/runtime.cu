// Runtime plumbing of oflib_b200: error reporting, device/stream/memory wrappers for the ctypes shim, and the
// host-buffer entry points (ofh_*) that stream frames through a small ring of device buffers so host<->device copies
// overlap the kernels.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "ofk_common.cuh"

namespace ofk {

static thread_local char t_error[512] = "";
std::atomic<unsigned long long> g_launches{0};
std::atomic<unsigned long long> g_paths[4];

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

}  // namespace ofk

using namespace ofk;

extern "C" const char* ofk_last_error(void) { return t_error; }
extern "C" int ofk_version(void) { return OFK_VERSION; }
extern "C" unsigned long long ofk_rt_launch_count(void) { return g_launches.load(); }
extern "C" unsigned long long ofk_rt_path_count(int which) {
    if (which >= 0 && which < 4) return g_paths[which].load();
    if (which == 4) return c3_ws_mixed_count();      // synchronous reads of device counters (current device)
    if (which == 5) return warp_ws_mixed_count();
    if (which >= 6 && which <= 21) return forward_s_stat(which - 6);
    return 0ull;
}

// ------------------------------------------------------------------------------------------------------- runtime
extern "C" int ofk_rt_device_count(int* count) {
    OFK_CHECK_ARG(count, "ofk_rt_device_count: NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return OFK_ECUDA;
    }
    return OFK_OK;
}
extern "C" int ofk_rt_set_device(int device) {
    OFK_CUDA(cudaSetDevice(device));
    return OFK_OK;
}
extern "C" int ofk_rt_get_device(int* device) {
    OFK_CHECK_ARG(device, "ofk_rt_get_device: NULL");
    OFK_CUDA(cudaGetDevice(device));
    return OFK_OK;
}
extern "C" int ofk_rt_device_info(int device, int* sm, int* cc_major, int* cc_minor, size_t* l2_bytes,
                                  size_t* total_mem) {
    cudaDeviceProp p;
    OFK_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm) *sm = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return OFK_OK;
}

// stream-ordered pool allocation: freed blocks stay cached in the device's default mempool (release threshold
// raised to "never"), so the per-call allocations of the Python shim cost microseconds, not cudaMalloc round trips.
static int ensure_pool(int dev) {
    static std::mutex mu;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64 || done[dev]) return OFK_OK;
    cudaMemPool_t pool;
    OFK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long thr = ~0ull;
    OFK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    done[dev] = true;
    return OFK_OK;
}
extern "C" int ofk_rt_malloc(void** dptr, size_t bytes, ofk_stream_t stream) {
    OFK_CHECK_ARG(dptr, "ofk_rt_malloc: NULL");
    int dev = 0;
    OFK_CUDA(cudaGetDevice(&dev));
    int rc = ensure_pool(dev);
    if (rc != OFK_OK) return rc;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(dptr, bytes, as_stream(stream));
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        set_error("device allocation of %zu bytes failed", bytes);
        return OFK_ENOMEM;
    }
    OFK_CUDA(e);
    return OFK_OK;
}
extern "C" int ofk_rt_free(void* dptr, ofk_stream_t stream) {
    if (dptr == nullptr) return OFK_OK;
    OFK_CUDA(cudaFreeAsync(dptr, as_stream(stream)));
    return OFK_OK;
}
extern "C" int ofk_rt_host_alloc(void** hptr, size_t bytes) {
    OFK_CHECK_ARG(hptr, "ofk_rt_host_alloc: NULL");
    OFK_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocPortable));
    return OFK_OK;
}
extern "C" int ofk_rt_host_free(void* hptr) {
    if (hptr) OFK_CUDA(cudaFreeHost(hptr));
    return OFK_OK;
}
extern "C" int ofk_rt_host_register(void* hptr, size_t bytes) {
    OFK_CUDA(cudaHostRegister(hptr, bytes, cudaHostRegisterPortable));
    return OFK_OK;
}
extern "C" int ofk_rt_host_unregister(void* hptr) {
    OFK_CUDA(cudaHostUnregister(hptr));
    return OFK_OK;
}
// Large copies from / to pageable memory go through staging.cu (worker threads + pinned pieces); everything else is a
// plain cudaMemcpyAsync.
extern "C" int ofk_rt_memcpy_h2d(void* dst, const void* src, size_t bytes, ofk_stream_t s) {
    if (bytes == 0) return OFK_OK;
    const int staged = staged_copy(dst, src, bytes, true, as_stream(s));
    if (staged < 0) return staged;
    if (staged == 0) OFK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(s)));
    return OFK_OK;
}
extern "C" int ofk_rt_memcpy_d2h(void* dst, const void* src, size_t bytes, ofk_stream_t s) {
    if (bytes == 0) return OFK_OK;
    const int staged = staged_copy(dst, src, bytes, false, as_stream(s));
    if (staged < 0) return staged;
    if (staged == 0) OFK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(s)));
    return OFK_OK;
}
extern "C" int ofk_rt_memcpy_d2d(void* dst, const void* src, size_t bytes, ofk_stream_t s) {
    if (bytes) OFK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(s)));
    return OFK_OK;
}
extern "C" int ofk_rt_memset(void* dst, int value, size_t bytes, ofk_stream_t s) {
    if (bytes) OFK_CUDA(cudaMemsetAsync(dst, value, bytes, as_stream(s)));
    return OFK_OK;
}
extern "C" int ofk_rt_stream_create(ofk_stream_t* s) {
    OFK_CHECK_ARG(s, "ofk_rt_stream_create: NULL");
    cudaStream_t st;
    OFK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *s = (ofk_stream_t)st;
    return OFK_OK;
}
extern "C" int ofk_rt_stream_destroy(ofk_stream_t s) {
    if (s) OFK_CUDA(cudaStreamDestroy(as_stream(s)));
    return OFK_OK;
}
extern "C" int ofk_rt_stream_sync(ofk_stream_t s) {
    OFK_CUDA(cudaStreamSynchronize(as_stream(s)));
    return OFK_OK;
}
extern "C" int ofk_rt_device_sync(void) {
    OFK_CUDA(cudaDeviceSynchronize());
    return OFK_OK;
}
extern "C" int ofk_rt_event_create(void** ev) {
    OFK_CHECK_ARG(ev, "ofk_rt_event_create: NULL");
    cudaEvent_t e;
    OFK_CUDA(cudaEventCreate(&e));
    *ev = (void*)e;
    return OFK_OK;
}
extern "C" int ofk_rt_event_destroy(void* ev) {
    if (ev) OFK_CUDA(cudaEventDestroy((cudaEvent_t)ev));
    return OFK_OK;
}
extern "C" int ofk_rt_event_record(void* ev, ofk_stream_t s) {
    OFK_CUDA(cudaEventRecord((cudaEvent_t)ev, as_stream(s)));
    return OFK_OK;
}
extern "C" int ofk_rt_stream_wait_event(ofk_stream_t s, void* ev) {
    OFK_CUDA(cudaStreamWaitEvent(as_stream(s), (cudaEvent_t)ev, 0));
    return OFK_OK;
}
extern "C" int ofk_rt_event_sync(void* ev) {
    OFK_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
    return OFK_OK;
}
extern "C" int ofk_rt_event_elapsed_ms(void* a, void* b, float* ms) {
    OFK_CHECK_ARG(ms, "ofk_rt_event_elapsed_ms: NULL");
    OFK_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
    return OFK_OK;
}

// ------------------------------------------------------------------------------------------------ host-buffer API
namespace {

constexpr int kSlots = 3;
constexpr size_t kChunkTarget = 48u << 20;  // ~48 MiB of traffic per chunk keeps PCIe busy without hoarding HBM

struct Slot {
    cudaStream_t st = nullptr;
    char* dev = nullptr;
    size_t cap = 0;
    size_t used = 0;
    // Small results (the zero flags of ofh_combine3) come back through a pinned staging buffer: a device-to-host copy
    // into the caller's pageable array would block the host until the chunk has finished and serialise the ring.
    int* hflags = nullptr;
    size_t hflags_cap = 0;       // ints
    int* flags_dst = nullptr;    // where the staged flags of the chunk in flight belong
    size_t flags_n = 0;
    void deliver() {             // call after the slot's stream has drained
        if (flags_dst != nullptr && flags_n) memcpy(flags_dst, hflags, flags_n * sizeof(int));
        flags_dst = nullptr;
        flags_n = 0;
    }
    void* take(size_t bytes) {
        size_t off = (used + 255) & ~size_t(255);
        used = off + bytes;
        return dev + off;
    }
};
struct Ring {
    int device = -1;
    Slot slot[kSlots];
    std::mutex mu;
};
Ring g_ring;

int ring_prepare(int device, size_t bytes_per_slot) {
    if (g_ring.device != device) {
        if (g_ring.device >= 0) {
            cudaSetDevice(g_ring.device);
            for (auto& s : g_ring.slot) {
                if (s.dev) cudaFree(s.dev);
                if (s.hflags) cudaFreeHost(s.hflags);
                if (s.st) cudaStreamDestroy(s.st);
                s = Slot();
            }
        }
        g_ring.device = device;
    }
    OFK_CUDA(cudaSetDevice(device));
    for (auto& s : g_ring.slot) {
        s.flags_dst = nullptr;   // nothing is pending between calls (a call that failed half-way drops its flags)
        s.flags_n = 0;
        if (!s.st) OFK_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
        if (s.cap < bytes_per_slot) {
            if (s.dev) {
                OFK_CUDA(cudaStreamSynchronize(s.st));
                OFK_CUDA(cudaFree(s.dev));
                s.dev = nullptr;
                s.cap = 0;
            }
            cudaError_t e = cudaMalloc((void**)&s.dev, bytes_per_slot);
            if (e != cudaSuccess) {
                cudaGetLastError();
                set_error("ring allocation of %zu bytes failed: %s", bytes_per_slot, cudaGetErrorString(e));
                return OFK_ENOMEM;
            }
            s.cap = bytes_per_slot;
        }
    }
    return OFK_OK;
}

// ofh_* calls run on the device they are given and leave the caller's current device as they found it.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Leaves no transfer in flight when an ofh_* call returns, on the error paths too: the copies read and write the
// caller's buffers, which may be freed as soon as the call is over.
struct DrainRing {
    ~DrainRing() {
        for (auto& s : g_ring.slot)
            if (s.st) cudaStreamSynchronize(s.st);
    }
};

size_t esize(int dtype) {
    switch (dtype) {
        case OFK_U8: return 1;
        case OFK_I16:
        case OFK_U16: return 2;
        case OFK_F32: return 4;
        default: return 8;
    }
}
size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

// Copies of the ring: pinned / registered user buffers -> plain asynchronous copies on the slot's stream (they overlap
// with the other slots); pageable buffers (plain numpy arrays) -> staging.cu (worker threads + pinned pieces), where the
// upload returns once the source has been read and the download once the destination is complete.
int ring_copy(void* dst, const void* src, size_t bytes, bool to_device, cudaStream_t st) {
    if (bytes == 0) return OFK_OK;
    const int staged = staged_copy(dst, src, bytes, to_device, st);
    if (staged < 0) return staged;
    if (staged == 0)
        OFK_CUDA(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st));
    return OFK_OK;
}
#define OFH_COPY_IN(dst, src, bytes)                                                   \
    do {                                                                               \
        const int rc_ = ring_copy(dst, src, bytes, true, s.st);                        \
        if (rc_ != OFK_OK) return rc_;                                                 \
    } while (0)
#define OFH_COPY_OUT(dst, src, bytes)                                                  \
    do {                                                                               \
        const int rc_ = ring_copy(dst, src, bytes, false, s.st);                       \
        if (rc_ != OFK_OK) return rc_;                                                 \
    } while (0)

}  // namespace

extern "C" int ofh_warp_t(const void* payload, int dtype, int C, int arith, const float* flow, float flow_sign,
                          const uint8_t* payload_mask, const uint8_t* flow_mask, void* out, uint8_t* out_mask,
                          int mask_rule, int N, int H, int W, int device) {
    OFK_CHECK_ARG(flow != nullptr && N >= 0 && H > 0 && W > 0 && C >= 0, "ofh_warp_t: bad arguments");
    OFK_CHECK_ARG(dtype >= OFK_U8 && dtype <= OFK_F64, "ofh_warp_t: unknown dtype %d", dtype);
    if (N == 0) return OFK_OK;
    std::lock_guard<std::mutex> lk(g_ring.mu);
    DeviceGuard restore_device;
    const size_t px = (size_t)H * W, es = esize(dtype);
    const size_t b_flow = px * 8, b_pay = px * C * es, b_m = px;
    const size_t per_frame = pad256(b_flow) + 2 * pad256(b_pay) + 3 * pad256(b_m);
    int cf = (int)(kChunkTarget / per_frame);
    if (cf < 1) cf = 1;
    if (cf > N) cf = N;
    int rc = ring_prepare(device, per_frame * cf + 4096);
    if (rc != OFK_OK) return rc;
    DrainRing drain_on_exit;
    struct Pending {
        Slot* slot;
        int n0, cn;
        void* d_out;
        uint8_t* d_om;
    };
    Pending prev = {};
    bool have_prev = false;
    auto download = [&](const Pending& p) -> int {
        Slot& s = *p.slot;
        if (C) OFH_COPY_OUT((char*)out + (size_t)p.n0 * b_pay, p.d_out, b_pay * p.cn);
        if (p.d_om) OFH_COPY_OUT(out_mask + (size_t)p.n0 * px, p.d_om, b_m * p.cn);
        return OFK_OK;
    };
    int chunk = 0;
    for (int n0 = 0; n0 < N; n0 += cf, ++chunk) {
        Slot& s = g_ring.slot[chunk % kSlots];
        const int cn = (N - n0 < cf) ? (N - n0) : cf;
        // a slot is reused only after everything queued on its stream has drained (keeps device buffers private)
        OFK_CUDA(cudaStreamSynchronize(s.st));
        s.used = 0;
        float* d_flow = (float*)s.take(b_flow * cn);
        void* d_pay = C ? s.take(b_pay * cn) : nullptr;
        void* d_out = C ? s.take(b_pay * cn) : nullptr;
        uint8_t* d_pm = payload_mask ? (uint8_t*)s.take(b_m * cn) : nullptr;
        uint8_t* d_fm = flow_mask ? (uint8_t*)s.take(b_m * cn) : nullptr;
        uint8_t* d_om = out_mask ? (uint8_t*)s.take(b_m * cn) : nullptr;
        OFH_COPY_IN(d_flow, flow + (size_t)n0 * px * 2, b_flow * cn);
        if (C) OFH_COPY_IN(d_pay, (const char*)payload + (size_t)n0 * b_pay, b_pay * cn);
        if (d_pm) OFH_COPY_IN(d_pm, payload_mask + (size_t)n0 * px, b_m * cn);
        if (d_fm) OFH_COPY_IN(d_fm, flow_mask + (size_t)n0 * px, b_m * cn);
        rc = ofk_warp_t(d_pay, dtype, C, arith, d_flow, flow_sign, d_pm, d_fm, d_out, d_om, mask_rule, cn, H, W, H, W,
                        0, 0, 1, (ofk_stream_t)s.st);
        if (rc != OFK_OK) return rc;
        // The download of a chunk is issued after the NEXT chunk has been uploaded and launched: with pageable user
        // buffers a download blocks the host, and this order keeps the device busy meanwhile (with pinned buffers the
        // order is irrelevant, every copy is asynchronous).
        if (have_prev) {
            rc = download(prev);
            if (rc != OFK_OK) return rc;
        }
        prev = {&s, n0, cn, d_out, d_om};
        have_prev = true;
    }
    if (have_prev) {
        rc = download(prev);
        if (rc != OFK_OK) return rc;
    }
    for (auto& s : g_ring.slot) OFK_CUDA(cudaStreamSynchronize(s.st));
    return OFK_OK;
}

extern "C" int ofh_combine3(const float* A, const uint8_t* Am, const float* B, const uint8_t* Bm, int ref, float thr,
                            float* out, uint8_t* out_mask, int* flags, int N, int H, int W, int device) {
    OFK_CHECK_ARG(A && B && out && out_mask && N >= 0 && H > 0 && W > 0, "ofh_combine3: bad arguments");
    if (N == 0) return OFK_OK;
    std::lock_guard<std::mutex> lk(g_ring.mu);
    DeviceGuard restore_device;
    const size_t px = (size_t)H * W;
    const size_t b_flow = px * 8, b_m = px;
    const size_t per_frame = 3 * pad256(b_flow) + 3 * pad256(b_m) + 256;
    int cf = (int)(kChunkTarget / per_frame);
    if (cf < 1) cf = 1;
    if (cf > N) cf = N;
    int rc = ring_prepare(device, per_frame * cf + 4096);
    if (rc != OFK_OK) return rc;
    DrainRing drain_on_exit;
    struct Pending {
        Slot* slot;
        int n0, cn;
        float* dO;
        uint8_t* dOm;
        int* dF;
    };
    Pending prev = {};
    bool have_prev = false;
    auto download = [&](const Pending& p) -> int {
        Slot& s = *p.slot;
        OFH_COPY_OUT(out + (size_t)p.n0 * px * 2, p.dO, b_flow * p.cn);
        OFH_COPY_OUT(out_mask + (size_t)p.n0 * px, p.dOm, b_m * p.cn);
        if (flags) {                                     // small: always through the slot's pinned staging words
            OFK_CUDA(cudaMemcpyAsync(s.hflags, p.dF, sizeof(int) * 2 * p.cn, cudaMemcpyDeviceToHost, s.st));
            s.flags_dst = flags + (size_t)p.n0 * 2;
            s.flags_n = (size_t)2 * p.cn;
        }
        return OFK_OK;
    };
    int chunk = 0;
    for (int n0 = 0; n0 < N; n0 += cf, ++chunk) {
        Slot& s = g_ring.slot[chunk % kSlots];
        const int cn = (N - n0 < cf) ? (N - n0) : cf;
        OFK_CUDA(cudaStreamSynchronize(s.st));
        s.deliver();
        s.used = 0;
        if (flags && s.hflags_cap < (size_t)2 * cf) {
            if (s.hflags) OFK_CUDA(cudaFreeHost(s.hflags));
            s.hflags = nullptr;
            s.hflags_cap = 0;
            OFK_CUDA(cudaHostAlloc((void**)&s.hflags, sizeof(int) * 2 * cf, cudaHostAllocDefault));
            s.hflags_cap = (size_t)2 * cf;
        }
        float* dA = (float*)s.take(b_flow * cn);
        float* dB = (float*)s.take(b_flow * cn);
        float* dO = (float*)s.take(b_flow * cn);
        uint8_t* dAm = Am ? (uint8_t*)s.take(b_m * cn) : nullptr;
        uint8_t* dBm = Bm ? (uint8_t*)s.take(b_m * cn) : nullptr;
        uint8_t* dOm = (uint8_t*)s.take(b_m * cn);
        int* dF = (int*)s.take(sizeof(int) * 2 * cn);
        OFH_COPY_IN(dA, A + (size_t)n0 * px * 2, b_flow * cn);
        OFH_COPY_IN(dB, B + (size_t)n0 * px * 2, b_flow * cn);
        if (dAm) OFH_COPY_IN(dAm, Am + (size_t)n0 * px, b_m * cn);
        if (dBm) OFH_COPY_IN(dBm, Bm + (size_t)n0 * px, b_m * cn);
        rc = ofk_combine3(dA, dAm, dB, dBm, ref, thr, dO, dOm, dF, cn, H, W, (ofk_stream_t)s.st);
        if (rc != OFK_OK) return rc;
        if (have_prev) {                                 // see ofh_warp_t: downloads trail the launches by one chunk
            rc = download(prev);
            if (rc != OFK_OK) return rc;
        }
        prev = {&s, n0, cn, dO, dOm, dF};
        have_prev = true;
    }
    if (have_prev) {
        rc = download(prev);
        if (rc != OFK_OK) return rc;
    }
    for (auto& s : g_ring.slot) {
        OFK_CUDA(cudaStreamSynchronize(s.st));
        s.deliver();
    }
    return OFK_OK;
}

// The pair of calls a frame pipeline makes per batch -- `A.apply(image, return_valid_area=True)` and
// `A.combine_with(B, 3)` -- on host buffers in one pass of the ring: A and its mask are uploaded ONCE and feed both
// kernels of a chunk (two separate calls upload them twice: 30 bytes per pixel in, this 21).
extern "C" int ofh_apply_combine3(const void* image, int dtype, int C, int arith, int mask_rule, const float* A,
                                  const uint8_t* Am, const float* B, const uint8_t* Bm, int ref, float thr,
                                  void* out_image, uint8_t* out_valid, float* out, uint8_t* out_mask, int* flags, int N,
                                  int H, int W, int device) {
    OFK_CHECK_ARG(image && A && B && out_image && out && out_mask && N >= 0 && H > 0 && W > 0 && C > 0,
                  "ofh_apply_combine3: bad arguments");
    OFK_CHECK_ARG(dtype >= OFK_U8 && dtype <= OFK_F64, "ofh_apply_combine3: unknown dtype %d", dtype);
    OFK_CHECK_ARG(ref == 's' || ref == 't', "ofh_apply_combine3: ref must be 's' or 't'");
    if (N == 0) return OFK_OK;
    std::lock_guard<std::mutex> lk(g_ring.mu);
    DeviceGuard restore_device;
    const size_t px = (size_t)H * W, es = esize(dtype);
    const size_t b_flow = px * 8, b_m = px, b_img = px * C * es;
    const size_t per_frame = 3 * pad256(b_flow) + 4 * pad256(b_m) + 2 * pad256(b_img) + 256;
    int cf = (int)(kChunkTarget / per_frame);
    if (cf < 1) cf = 1;
    if (cf > N) cf = N;
    int rc = ring_prepare(device, per_frame * cf + 8192);
    if (rc != OFK_OK) return rc;
    DrainRing drain_on_exit;
    struct Pending {
        Slot* slot;
        int n0, cn;
        void* dOi;
        uint8_t* dOv;
        float* dO;
        uint8_t* dOm;
        int* dF;
    };
    Pending prev = {};
    bool have_prev = false;
    auto download = [&](const Pending& p) -> int {
        Slot& s = *p.slot;
        OFH_COPY_OUT((char*)out_image + (size_t)p.n0 * b_img, p.dOi, b_img * p.cn);
        if (p.dOv) OFH_COPY_OUT(out_valid + (size_t)p.n0 * px, p.dOv, b_m * p.cn);
        OFH_COPY_OUT(out + (size_t)p.n0 * px * 2, p.dO, b_flow * p.cn);
        OFH_COPY_OUT(out_mask + (size_t)p.n0 * px, p.dOm, b_m * p.cn);
        if (flags) {
            OFK_CUDA(cudaMemcpyAsync(s.hflags, p.dF, sizeof(int) * 2 * p.cn, cudaMemcpyDeviceToHost, s.st));
            s.flags_dst = flags + (size_t)p.n0 * 2;
            s.flags_n = (size_t)2 * p.cn;
        }
        return OFK_OK;
    };
    int chunk = 0;
    for (int n0 = 0; n0 < N; n0 += cf, ++chunk) {
        Slot& s = g_ring.slot[chunk % kSlots];
        const int cn = (N - n0 < cf) ? (N - n0) : cf;
        OFK_CUDA(cudaStreamSynchronize(s.st));
        s.deliver();
        s.used = 0;
        if (flags && s.hflags_cap < (size_t)2 * cf) {
            if (s.hflags) OFK_CUDA(cudaFreeHost(s.hflags));
            s.hflags = nullptr;
            s.hflags_cap = 0;
            OFK_CUDA(cudaHostAlloc((void**)&s.hflags, sizeof(int) * 2 * cf, cudaHostAllocDefault));
            s.hflags_cap = (size_t)2 * cf;
        }
        float* dA = (float*)s.take(b_flow * cn);
        float* dB = (float*)s.take(b_flow * cn);
        float* dO = (float*)s.take(b_flow * cn);
        void* dI = s.take(b_img * cn);
        void* dOi = s.take(b_img * cn);
        uint8_t* dAm = Am ? (uint8_t*)s.take(b_m * cn) : nullptr;
        uint8_t* dBm = Bm ? (uint8_t*)s.take(b_m * cn) : nullptr;
        uint8_t* dOm = (uint8_t*)s.take(b_m * cn);
        uint8_t* dOv = out_valid ? (uint8_t*)s.take(b_m * cn) : nullptr;
        int* dF = (int*)s.take(sizeof(int) * 2 * cn);
        OFH_COPY_IN(dA, A + (size_t)n0 * px * 2, b_flow * cn);
        if (dAm) OFH_COPY_IN(dAm, Am + (size_t)n0 * px, b_m * cn);
        OFH_COPY_IN(dI, (const char*)image + (size_t)n0 * b_img, b_img * cn);
        // the image warp starts while B is still on its way
        rc = ofk_warp_t(dI, dtype, C, arith, dA, ref == 't' ? -1.0f : 1.0f, nullptr, dOv ? dAm : nullptr, dOi, dOv,
                        mask_rule, cn, H, W, H, W, 0, 0, 1, (ofk_stream_t)s.st);
        if (rc != OFK_OK) return rc;
        OFH_COPY_IN(dB, B + (size_t)n0 * px * 2, b_flow * cn);
        if (dBm) OFH_COPY_IN(dBm, Bm + (size_t)n0 * px, b_m * cn);
        rc = ofk_combine3(dA, dAm, dB, dBm, ref, thr, dO, dOm, dF, cn, H, W, (ofk_stream_t)s.st);
        if (rc != OFK_OK) return rc;
        if (have_prev) {                                 // see ofh_warp_t: downloads trail the launches by one chunk
            rc = download(prev);
            if (rc != OFK_OK) return rc;
        }
        prev = {&s, n0, cn, dOi, dOv, dO, dOm, dF};
        have_prev = true;
    }
    if (have_prev) {
        rc = download(prev);
        if (rc != OFK_OK) return rc;
    }
    for (auto& s : g_ring.slot) {
        OFK_CUDA(cudaStreamSynchronize(s.st));
        s.deliver();
    }
    return OFK_OK;
}

extern "C" int ofh_release(void) {
    std::lock_guard<std::mutex> lk(g_ring.mu);
    DeviceGuard restore_device;
    if (g_ring.device >= 0) {
        cudaSetDevice(g_ring.device);
        for (auto& s : g_ring.slot) {
            if (s.st) cudaStreamSynchronize(s.st);
            if (s.dev) cudaFree(s.dev);
            if (s.hflags) cudaFreeHost(s.hflags);
            if (s.st) cudaStreamDestroy(s.st);
            s = Slot();
        }
        g_ring.device = -1;
    }
    return OFK_OK;
}
