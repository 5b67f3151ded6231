// common.cuh -- engine context, device workspace, launch instrumentation (host side) for libimcvt_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include <mutex>

namespace imcvt {

void set_error(const char *fmt, ...);

#define IMCVT_CUDA(call)                                                                        \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            imcvt::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

// One growable device allocation carved into aligned sub-buffers per call (bump allocator).
// Kept across calls so steady-state encoding does no cudaMalloc.
struct Workspace {
    char  *base = nullptr;
    size_t cap = 0, used = 0;
    int reserve(size_t bytes);                      // grow (frees + reallocates) if needed; resets the bump pointer
    void reset() { used = 0; }
    template <class T> T *take(size_t count) {      // 256-byte aligned slice; nullptr on overflow
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        if (used + bytes > cap) return nullptr;
        T *p = reinterpret_cast<T *>(base + used);
        used += bytes;
        return p;
    }
    void release();
    // A workspace is carved anew by every call, while the kernels of the call before may still be running on ANOTHER stream
    // (the device-buffer entry points do not synchronise): the new call's stream first waits for the previous user.
    cudaEvent_t busy = nullptr;
    int wait_previous(cudaStream_t st);             // after reserve(), before anything is queued that writes the workspace
    int mark_busy(cudaStream_t st);                 // after the last kernel of the call has been queued
};

struct KernelStat { int launches = 0; double ms = 0.0; };

struct Engine {
    int device = -1;
    bool ready = false;
    std::recursive_mutex mu;    // the extern "C" entry points serialise on the engine they run on
    Workspace ws;           // scratch for the device pipelines
    Workspace io;           // staging for the host-buffer entry points (frames in, streams out)
    cudaStream_t stream = nullptr;      // engine-owned stream for the host-buffer entry points
    static constexpr int NCHUNK = 4;    // host-buffer JPEG-LS batches are cut into chunks whose copies and kernels overlap
    Workspace ws_chunk[NCHUNK];
    cudaStream_t chunk_stream[NCHUNK] = {nullptr, nullptr, nullptr, nullptr};
    static constexpr int NFLAG = 64;
    int *d_flag = nullptr;              // device overflow flags of the device-buffer entry points: NFLAG of them, one per call in
    unsigned flag_next = 0;             // turn, so that calls queued on different streams do not share one
    bool attr_hevc = false, attr_jls_tile = false, attr_jls_rowscan = false, attr_jls_nearscan = false;   // dynamic shared-memory opt-in done on this device
    uint64_t *pinned_lens = nullptr;    // pinned host buffer the chunks' stream lengths come back into
    int pinned_lens_cap = 0;
    bool profile = false;
    long launches = 0;
    std::map<std::string, KernelStat> stats;
    struct Pending { std::string name; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> event_pool;
    bool jls_force_serial = false;
    int jls_path = 0;       // 0 = default (tile kernel when the plane fits), 1 = force the multi-kernel pipeline, 3 = tile kernel on CTA pairs
    int jls_host_mode = -1; // host-buffer JPEG-LS batches with page-locked buffers: -1 = in place when the batch runs on the cluster-pair kernel, 0 = always staged through HBM, 1 = frames read in place over PCIe, 2 = streams written in place too
    int hevc_path = 0;      // 0 = default (CTA pair per frame when the batch has at most SMs/2 frames), 1 = one CTA per frame, 2 = CTA pair
    int sm_count = 148;
};

// The engine the calling thread works on: the process-wide default engine (bound by imcvtInit, else to the current device),
// or -- on the worker threads of the multi-device entry points -- the engine of that worker's device.
Engine &engine();
Engine *device_engine(int device);      // the engine of one device (created on first use; the default engine when it sits on that device)
void select_engine(Engine *e);          // this thread's engine from now on (nullptr = the default engine)
void shutdown_all_engines();
void for_each_engine(void (*fn)(Engine &, int), int v);
int ensure_engine();
// The extern "C" entry points serialise on their engine's lock, so calls from several host threads are safe (calls on ONE
// engine do not overlap each other -- overlap comes from the streams inside a call; calls on different engines, i.e.
// different devices, run side by side).  The scope also puts the calling thread's current CUDA device back when the entry
// point returns: the engine's device is only current inside the call.
std::recursive_mutex &api_mutex();
struct DeviceScope {
    int saved = -1;
    DeviceScope() { if (cudaGetDevice(&saved) != cudaSuccess) { (void)cudaGetLastError(); saved = -1; } }
    ~DeviceScope() { if (saved >= 0) { int cur = -1; if (cudaGetDevice(&cur) == cudaSuccess && cur != saved) cudaSetDevice(saved); } }
};
#define IMCVT_API_LOCK() std::lock_guard<std::recursive_mutex> imcvt_api_lock__(imcvt::api_mutex()); imcvt::DeviceScope imcvt_dev_scope__

// RAII bracket around one kernel launch: counts it and (when profiling) times it with events on `s`.
struct LaunchScope {
    cudaStream_t s; const char *name; cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(const char *name_, cudaStream_t s_);
    ~LaunchScope();
};
void profile_collect();     // resolves pending event pairs (synchronises on them)

}  // namespace imcvt
