// Copies between PAGEABLE host memory (ordinary numpy arrays) and the device at close to PCIe rate.
//
// cudaMemcpyAsync on pageable memory is staged by the driver through one thread (measured here: 7-10 GB/s against
// 55 GB/s for pinned memory), and that is what the drop-in functions pay when they are handed plain numpy arrays
// (apply_flow, combine_flows, Flow(...).vecs). Large pageable copies are therefore split into chunks that a few worker
// threads move through their own pinned buffers: thread t copies chunk k into pinned memory while the DMA engine
// transfers chunk k-1 (and the other threads' chunks). Small copies, pinned or registered memory and device pointers
// take the plain cudaMemcpyAsync path.
//
// Ordering: the staged transfer runs on an internal copy stream that first waits for the work already queued on the
// caller's stream; afterwards the caller's stream waits for the transfer (H2D), or the call returns only when the data
// is in the destination (D2H: the destination is pageable memory, the call is synchronous like the plain path).
// H2D returns after the source has been read completely, as cudaMemcpyAsync does for pageable memory.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <thread>
#include <vector>

#include "ofk_common.cuh"

namespace ofk {
namespace {

constexpr size_t kChunk = 2u << 20;        // bytes per pinned buffer = largest staged piece
constexpr int kThreads = 4;                // worker threads per transfer (8 measure the same: host memory bound)
constexpr int kBufsPerThread = 3;          // pinned pieces in flight per thread
constexpr size_t kMinChunk = 256u << 10;   // smallest piece (small transfers still get every lane busy)
constexpr size_t kMinStaged = 4u << 20;    // below this the plain path wins (thread start-up, event traffic: measured)

struct Lane {
    char* pinned[kBufsPerThread] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[kBufsPerThread] = {nullptr, nullptr, nullptr};
};

struct Stager {
    std::mutex mu;                // one staged transfer at a time per process
    int device = -1;
    cudaStream_t copy = nullptr;
    cudaEvent_t before = nullptr, after = nullptr;
    Lane lane[kThreads];
    bool broken = false;          // a set-up failure disables staging for the process (plain path instead)
};
Stager g_stager;

int staging_mode() {              // OFK_STAGED_COPIES=0 disables
    static std::atomic<int> mode{-1};
    if (mode.load() < 0) {
        const char* e = getenv("OFK_STAGED_COPIES");
        mode.store((e != nullptr && e[0] == '0') ? 0 : 1);
    }
    return mode.load();
}

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

bool prepare(Stager& s, int device) {
    if (s.broken) return false;
    if (s.device == device && s.copy != nullptr) return true;
    if (s.device >= 0 && s.device != device) return false;   // buffers belong to another device: plain path
    bool ok = cudaStreamCreateWithFlags(&s.copy, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&s.before, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&s.after, cudaEventDisableTiming) == cudaSuccess;
    for (int t = 0; ok && t < kThreads; ++t)
        for (int b = 0; ok && b < kBufsPerThread; ++b)
            ok = cudaHostAlloc((void**)&s.lane[t].pinned[b], kChunk, cudaHostAllocPortable) == cudaSuccess &&
                 cudaEventCreateWithFlags(&s.lane[t].ev[b], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        s.broken = true;
        return false;
    }
    s.device = device;
    return true;
}

// One worker: chunks t, t + kThreads, ... of the transfer. Returns the first CUDA error it met.
cudaError_t run_lane(Stager& s, int t, int device, bool to_device, char* dst, const char* src, size_t bytes,
                     size_t chunk) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    Lane& ln = s.lane[t];
    const size_t n_chunks = (bytes + chunk - 1) / chunk;
    size_t issued = 0;                                   // chunks of this lane handed to the DMA engine so far
    size_t pending_off[kBufsPerThread], pending_len[kBufsPerThread];
    for (size_t k = (size_t)t; k < n_chunks; k += kThreads, ++issued) {
        const int b = (int)(issued % kBufsPerThread);
        const size_t off = k * chunk, len = bytes - off < chunk ? bytes - off : chunk;
        // The piece that used this buffer last must have left it: an earlier piece of this transfer, or -- for the
        // first pieces -- the tail of the PREVIOUS upload, whose DMA may still be running when this call starts
        // (an event that was never recorded counts as complete).
        if ((e = cudaEventSynchronize(ln.ev[b])) != cudaSuccess) return e;
        if (issued >= (size_t)kBufsPerThread && !to_device) memcpy(dst + pending_off[b], ln.pinned[b], pending_len[b]);
        if (to_device) {
            memcpy(ln.pinned[b], src + off, len);
            if ((e = cudaMemcpyAsync(dst + off, ln.pinned[b], len, cudaMemcpyHostToDevice, s.copy)) != cudaSuccess) return e;
        } else {
            if ((e = cudaMemcpyAsync(ln.pinned[b], src + off, len, cudaMemcpyDeviceToHost, s.copy)) != cudaSuccess) return e;
            pending_off[b] = off;
            pending_len[b] = len;
        }
        if ((e = cudaEventRecord(ln.ev[b], s.copy)) != cudaSuccess) return e;
    }
    // drain: the last pieces of this lane, oldest first
    const size_t tail = issued < (size_t)kBufsPerThread ? issued : (size_t)kBufsPerThread;
    for (size_t i = issued - tail; i < issued; ++i) {
        const int b = (int)(i % kBufsPerThread);
        if ((e = cudaEventSynchronize(ln.ev[b])) != cudaSuccess) return e;
        if (!to_device) memcpy(dst + pending_off[b], ln.pinned[b], pending_len[b]);
    }
    return cudaSuccess;
}

}  // namespace

// Returns 1 if the copy was done here, 0 if the caller should use the plain path, negative OFK_E* on error.
int staged_copy(void* dst, const void* src, size_t bytes, bool to_device, cudaStream_t user) {
    if (bytes < kMinStaged || staging_mode() == 0) return 0;
    if (!is_pageable(to_device ? src : dst)) return 0;
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    Stager& s = g_stager;
    std::lock_guard<std::mutex> lk(s.mu);
    if (!prepare(s, device)) return 0;
    // the transfer starts after everything already queued on the caller's stream
    OFK_CUDA(cudaEventRecord(s.before, user));
    OFK_CUDA(cudaStreamWaitEvent(s.copy, s.before, 0));
    // piece size: two pieces per lane at least, between kMinChunk and the pinned buffer size, a multiple of 4 KiB
    size_t chunk = bytes / (2 * kThreads);
    chunk = chunk < kMinChunk ? kMinChunk : (chunk > kChunk ? kChunk : chunk);
    chunk = (chunk + 4095) & ~size_t(4095);
    if (chunk > kChunk) chunk = kChunk;
    cudaError_t errs[kThreads];
    std::vector<std::thread> workers;
    workers.reserve(kThreads - 1);
    for (int t = 1; t < kThreads; ++t)
        workers.emplace_back([&, t]() {
            errs[t] = run_lane(s, t, device, to_device, static_cast<char*>(dst), static_cast<const char*>(src), bytes, chunk);
        });
    errs[0] = run_lane(s, 0, device, to_device, static_cast<char*>(dst), static_cast<const char*>(src), bytes, chunk);
    for (auto& w : workers) w.join();
    for (int t = 0; t < kThreads; ++t)
        if (errs[t] != cudaSuccess) {
            set_error("staged copy failed: %s", cudaGetErrorString(errs[t]));
            cudaGetLastError();
            return OFK_ECUDA;
        }
    // later work on the caller's stream sees the data (H2D); D2H has been drained by the lanes already
    OFK_CUDA(cudaEventRecord(s.after, s.copy));
    OFK_CUDA(cudaStreamWaitEvent(user, s.after, 0));
    return 1;
}

}  // namespace ofk
