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

// The default engine plus one engine per device for the multi-device entry points (created on first use, never moved:
// worker threads keep pointers to them).
static constexpr int MAX_DEVICES = 64;
static Engine g_default;
static Engine *g_dev[MAX_DEVICES];
static std::mutex g_pool_mu;
static thread_local Engine *tl_engine = nullptr;

Engine &engine() { return tl_engine ? *tl_engine : g_default; }

void select_engine(Engine *e) { tl_engine = e; }

Engine *device_engine(int device)
{
    if (device < 0 || device >= MAX_DEVICES) return nullptr;
    std::lock_guard<std::mutex> g(g_pool_mu);
    if (g_default.ready && g_default.device == device) return &g_default;     // one context, one workspace per device
    if (!g_dev[device]) {
        g_dev[device] = new Engine();
        g_dev[device]->device = device;
        // the debug / mode switches of the default engine apply to every device
        g_dev[device]->jls_force_serial = g_default.jls_force_serial;
        g_dev[device]->jls_path = g_default.jls_path;
        g_dev[device]->jls_host_mode = g_default.jls_host_mode;
        g_dev[device]->hevc_path = g_default.hevc_path;
    }
    return g_dev[device];
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

std::recursive_mutex &api_mutex() { return engine().mu; }

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
    if (prop.major != 10 || prop.minor != 0) {       // sm_100a code is architecture-specific: it runs on sm_100 only (not sm_103 / sm_12x)
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

static void shutdown_engine(Engine &E)
{
    std::lock_guard<std::recursive_mutex> lock(E.mu);
    if (!E.ready) return;
    Engine *prev = tl_engine;
    tl_engine = &E;                                   // profile_collect() and friends work on the calling thread's engine
    cudaSetDevice(E.device);
    profile_collect();
    for (auto ev : E.event_pool) cudaEventDestroy(ev);
    E.event_pool.clear();
    E.ws.release();
    E.io.release();
    if (E.d_flag) cudaFree(E.d_flag);
    E.d_flag = nullptr;
    E.attr_hevc = E.attr_jls_tile = E.attr_jls_rowscan = E.attr_jls_nearscan = false;
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
    tl_engine = prev;
}

void for_each_engine(void (*fn)(Engine &, int), int v)
{
    fn(g_default, v);
    std::lock_guard<std::mutex> g(g_pool_mu);
    for (int d = 0; d < MAX_DEVICES; d++)
        if (g_dev[d]) fn(*g_dev[d], v);
}

void shutdown_all_engines()
{
    shutdown_engine(g_default);
    g_default.device = -1;
    std::lock_guard<std::mutex> g(g_pool_mu);
    for (int d = 0; d < MAX_DEVICES; d++)
        if (g_dev[d]) shutdown_engine(*g_dev[d]);     // the objects stay: a later multi-device call starts them again
}

}  // namespace imcvt

using namespace imcvt;

extern "C" {

int imcvtInit(int device)
{
    IMCVT_API_LOCK();
    Engine &E = engine();
    if (E.ready && E.device != device) shutdown_engine(E);
    E.device = device;
    return ensure_engine();
}


void imcvtShutdown(void)
{
    imcvt::DeviceScope scope;
    imcvt::shutdown_all_engines();     // the objects stay: a later multi-device call starts them again
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

// switches apply to the default engine and to every device engine of the multi-device entry points
void imcvtDebugJLSForceSerial(int on) { for_each_engine([](Engine &E, int v) { E.jls_force_serial = v != 0; }, on); }
void imcvtDebugJLSPath(int path) { for_each_engine([](Engine &E, int v) { E.jls_path = v; }, path); }
void imcvtDebugHEVCPath(int path) { for_each_engine([](Engine &E, int v) { E.hevc_path = v; }, path); }
void imcvtJLSHostMode(int mode) { for_each_engine([](Engine &E, int v) { E.jls_host_mode = v; }, mode); }

}  // extern "C"
