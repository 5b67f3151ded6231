// engine.cu -- context, workspace, error reporting and launch instrumentation of libimcvt_b200.so
#include "common.cuh"
#include "../../include/imcvt_b200.h"
#include <stdarg.h>
#include <string.h>
#include <mutex>

namespace imcvt {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

Engine &engine()
{
    static Engine e;
    return e;
}

int Workspace::reserve(size_t bytes)
{
    used = 0;
    if (bytes <= cap) return 0;
    if (base) { cudaFree(base); base = nullptr; cap = 0; }
    size_t want = bytes + (bytes >> 3) + (1u << 20);            // head-room so that small growth does not realloc
    cudaError_t e = cudaMalloc(&base, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        e = cudaMalloc(&base, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        base = nullptr;
        set_error("device workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return 1;
    }
    cap = want;
    return 0;
}

void Workspace::release()
{
    if (base) cudaFree(base);
    base = nullptr; cap = used = 0;
    if (busy) cudaEventDestroy(busy);
    busy = nullptr;
}

int Workspace::wait_previous(cudaStream_t st)
{
    if (busy) IMCVT_CUDA(cudaStreamWaitEvent(st, busy, 0));
    return 0;
}

int Workspace::mark_busy(cudaStream_t st)
{
    if (!busy) IMCVT_CUDA(cudaEventCreateWithFlags(&busy, cudaEventDisableTiming));
    IMCVT_CUDA(cudaEventRecord(busy, st));
    return 0;
}

std::recursive_mutex &api_mutex()
{
    static std::recursive_mutex m;
    return m;
}

int ensure_engine()
{
    Engine &E = engine();
    if (E.ready) {
        IMCVT_CUDA(cudaSetDevice(E.device));
        return 0;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (%s): libimcvt_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return 1;
    }
    int dev = 0;
    if (E.device >= 0) dev = E.device; else IMCVT_CUDA(cudaGetDevice(&dev));
    IMCVT_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    IMCVT_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) {
        set_error("device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; this library only carries sm_100a code",
                  dev, prop.name, prop.major, prop.minor);
        return 1;
    }
    E.device = dev;
    E.sm_count = prop.multiProcessorCount;
    IMCVT_CUDA(cudaStreamCreateWithFlags(&E.stream, cudaStreamNonBlocking));
    E.ready = true;
    return 0;
}

LaunchScope::LaunchScope(const char *name_, cudaStream_t s_) : s(s_), name(name_)
{
    Engine &E = engine();
    E.launches++;
    if (!E.profile) return;
    auto get = [&]() {
        cudaEvent_t ev;
        if (!E.event_pool.empty()) { ev = E.event_pool.back(); E.event_pool.pop_back(); }
        else cudaEventCreate(&ev);
        return ev;
    };
    a = get(); b = get();
    cudaEventRecord(a, s);
}

LaunchScope::~LaunchScope()
{
    if (!a) return;
    cudaEventRecord(b, s);
    engine().pending.push_back({name, a, b});
}

void profile_collect()
{
    Engine &E = engine();
    for (auto &p : E.pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            KernelStat &k = E.stats[p.name];
            k.launches++;
            k.ms += ms;
        } else {
            (void)cudaGetLastError();
        }
        E.event_pool.push_back(p.a);
        E.event_pool.push_back(p.b);
    }
    E.pending.clear();
}

}  // namespace imcvt

using namespace imcvt;

extern "C" {

int imcvtInit(int device)
{
    IMCVT_API_LOCK();
    Engine &E = engine();
    if (E.ready && E.device != device) imcvtShutdown();
    E.device = device;
    return ensure_engine();
}

void imcvtShutdown(void)
{
    IMCVT_API_LOCK();
    Engine &E = engine();
    if (!E.ready) return;
    cudaSetDevice(E.device);
    profile_collect();
    for (auto ev : E.event_pool) cudaEventDestroy(ev);
    E.event_pool.clear();
    E.ws.release();
    E.io.release();
    if (E.d_flag) cudaFree(E.d_flag);
    E.d_flag = nullptr;
    E.attr_hevc = E.attr_jls_tile = E.attr_jls_rowscan = false;
    if (E.pinned_lens) cudaFreeHost(E.pinned_lens);
    E.pinned_lens = nullptr; E.pinned_lens_cap = 0;
    for (int c = 0; c < Engine::NCHUNK; c++) {
        E.ws_chunk[c].release();
        if (E.chunk_stream[c]) cudaStreamDestroy(E.chunk_stream[c]);
        E.chunk_stream[c] = nullptr;
    }
    if (E.stream) cudaStreamDestroy(E.stream);
    E.stream = nullptr;
    E.ready = false;
    E.device = -1;
}

const char *imcvtLastError(void) { return g_err; }
int imcvtVersion(void) { return 100; }

void imcvtProfileEnable(int on) { engine().profile = on != 0; }

void imcvtProfileReset(void)
{
    IMCVT_API_LOCK();
    profile_collect();
    engine().stats.clear();
    engine().launches = 0;
}

int imcvtProfileGet(const char *kernel, int *launches, double *total_ms)
{
    IMCVT_API_LOCK();
    profile_collect();
    auto &st = engine().stats;
    auto it = st.find(kernel);
    if (it == st.end()) { if (launches) *launches = 0; if (total_ms) *total_ms = 0; return 1; }
    if (launches) *launches = it->second.launches;
    if (total_ms) *total_ms = it->second.ms;
    return 0;
}

int imcvtProfileList(char *buf, size_t cap)
{
    IMCVT_API_LOCK();
    profile_collect();
    std::string s;
    for (auto &kv : engine().stats) {
        char line[160];
        snprintf(line, sizeof line, "%s:%d:%.6f\n", kv.first.c_str(), kv.second.launches, kv.second.ms);
        s += line;
    }
    if (s.size() + 1 > cap) return 1;
    memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}

long imcvtLaunchCount(void) { return engine().launches; }

void imcvtDebugJLSForceSerial(int on) { engine().jls_force_serial = on != 0; }
void imcvtDebugJLSPath(int path) { engine().jls_path = path; }
void imcvtDebugHEVCPath(int path) { engine().hevc_path = path; }
void imcvtJLSHostMode(int mode) { engine().jls_host_mode = mode; }

}  // extern "C"
