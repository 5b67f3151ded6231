// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY (never part of libimcvt_b200.so).
//
// A minimal CPU SIMT emulator: runs the *source* of a CUDA kernel on the host so that the kernel's logic can be
// checked bit-for-bit against the oracle in the CPU test tier (and debugged without a GPU).  Every CUDA thread
// of a block is a fiber (own stack, hand-rolled context switch); __syncthreads / __syncwarp / shuffles / ballots are cooperative yield
// points resolved by a round-robin scheduler.  A barrier that can never complete (divergent __syncthreads,
// a shuffle executed by part of a warp only) is reported as a deadlock instead of hanging -- which is exactly
// the class of bug that would hang or corrupt a real GPU run.
//
// Blocks run one after the other, so there is no inter-block concurrency (kernels that need it cannot be
// emulated).  Dynamic shared memory is filled with 0xCD to surface reads of uninitialised shared memory.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __constant__ static
#define __launch_bounds__(...)
#define IMCVT_HOST_EMU 1

namespace emu {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

enum { RUNNABLE = 0, AT_BLOCK_BARRIER = 1, AT_WARP_BARRIER = 2, DONE = 3, AT_NAMED_BARRIER = 4 };

// Hand-rolled x86-64 context switch (callee-saved registers + stack pointer): ~50x cheaper than swapcontext,
// which makes a system call per switch.
extern "C" void emu_switch(void **save_sp, void *new_sp);
#if defined(__x86_64__)
asm(".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_switch,.-emu_switch\n");
#else
#error "cuda_emu.h: only x86-64 hosts are supported"
#endif

struct Fiber {
    void *sp = nullptr;
    char *stack = nullptr;
    int state = RUNNABLE;
    int bar_id = 0, bar_count = 0;          // named barrier (bar.sync id, count) this fiber waits at
};

struct State {
    Dim3 blockIdx, blockDim, gridDim;
    int tid = 0, nthreads = 0;
    void *sched_sp = nullptr;
    std::vector<Fiber> fibers;
    std::vector<uint64_t> slots;            // [warp][32] exchange slots for shuffles / ballots
    unsigned char *smem = nullptr;
    int bar_arrived[24] = {0};              // bar.arrive: threads that announced themselves at named barrier id without waiting
    int cluster = 1;                        // CTAs per cluster: the fibers of one launch step are cluster x cta_threads
    int cta_threads = 0;                    // threads of one CTA
    size_t smem_stride = 0;                 // distance of the CTAs' shared-memory images
    unsigned cluster_id = 0;
    std::function<void()> body;
    long n_block_barriers = 0, n_warp_barriers = 0;
};

inline State &S() { static State s; return s; }

inline void yield_as(int st)
{
    State &s = S();
    Fiber &f = s.fibers[s.tid];
    f.state = st;
    emu_switch(&f.sp, s.sched_sp);
}

inline void trampoline()
{
    State &s = S();
    s.body();
    s.fibers[s.tid].state = DONE;
    emu_switch(&s.fibers[s.tid].sp, s.sched_sp);
    abort();
}

constexpr size_t STACK_BYTES = 512 * 1024;

// runs `body` once per CUDA thread for every block of the grid; cluster > 1: grid_x counts CLUSTERS of `cluster` CTAs
// whose threads run together (block_x threads and smem_bytes of shared memory per CTA, block_x a multiple of 32)
inline void launch(unsigned grid_x, unsigned block_x, size_t smem_bytes, std::function<void()> body, int cluster = 1)
{
    State &s = S();
    s.cluster = cluster; s.cta_threads = (int)block_x;
    s.smem_stride = (smem_bytes + 127) & ~(size_t)127;
    s.gridDim.x = grid_x * cluster; s.blockDim.x = block_x;
    const unsigned nfib = block_x * cluster;
    s.nthreads = (int)nfib;
    s.body = body;
    s.fibers.resize(nfib);
    s.slots.assign((size_t)((nfib + 31) / 32) * 32, 0);
    for (auto &f : s.fibers) if (!f.stack) f.stack = (char *)malloc(STACK_BYTES);
    s.smem = (unsigned char *)malloc(s.smem_stride * cluster + 64);
    for (unsigned b = 0; b < grid_x; b++) {
        s.blockIdx.x = b * cluster;
        s.cluster_id = b;
        memset(s.smem, 0xCD, s.smem_stride * cluster + 64);
        for (unsigned t = 0; t < nfib; t++) {
            Fiber &f = s.fibers[t];
            void **top = (void **)(((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15);
            *--top = nullptr;                       // fake return address: keeps the ABI stack alignment at entry
            *--top = (void *)trampoline;            // `ret` of the first switch lands here
            for (int k = 0; k < 6; k++) *--top = nullptr;
            f.sp = top;
            f.state = RUNNABLE;
        }
        const int nwarps = (int)(nfib + 31) / 32;
        for (;;) {
            bool ran = false;
            for (int t = 0; t < s.nthreads; t++)
                if (s.fibers[t].state == RUNNABLE) {
                    s.tid = t;
                    emu_switch(&s.sched_sp, s.fibers[t].sp);
                    ran = true;
                }
            bool released = false;
            // warp barriers: every live lane of the warp waits at one
            for (int w = 0; w < nwarps; w++) {
                int live = 0, waiting = 0;
                for (int l = 0; l < 32 && w * 32 + l < s.nthreads; l++) {
                    int st = s.fibers[w * 32 + l].state;
                    if (st != DONE) live++;
                    if (st == AT_WARP_BARRIER) waiting++;
                }
                if (live > 0 && waiting == live) {
                    for (int l = 0; l < 32 && w * 32 + l < s.nthreads; l++)
                        if (s.fibers[w * 32 + l].state == AT_WARP_BARRIER) s.fibers[w * 32 + l].state = RUNNABLE;
                    s.n_warp_barriers++;
                    released = true;
                }
            }
            // named barriers: `count` threads wait at the same id
            for (int id = 1; id < 24; id++) {
                int waiting_n = 0, want = 0;
                for (int t = 0; t < s.nthreads; t++)
                    if (s.fibers[t].state == AT_NAMED_BARRIER && s.fibers[t].bar_id == id) { waiting_n++; want = s.fibers[t].bar_count; }
                if (waiting_n > 0 && waiting_n + s.bar_arrived[id] >= want) {
                    s.bar_arrived[id] = 0;
                    for (int t = 0; t < s.nthreads; t++)
                        if (s.fibers[t].state == AT_NAMED_BARRIER && s.fibers[t].bar_id == id) s.fibers[t].state = RUNNABLE;
                    released = true;
                }
            }
            // block barrier
            int live = 0, waiting = 0;
            for (int t = 0; t < s.nthreads; t++) {
                int st = s.fibers[t].state;
                if (st != DONE) live++;
                if (st == AT_BLOCK_BARRIER) waiting++;
            }
            if (live == 0) break;
            if (waiting == live) {
                for (int t = 0; t < s.nthreads; t++)
                    if (s.fibers[t].state == AT_BLOCK_BARRIER) s.fibers[t].state = RUNNABLE;
                s.n_block_barriers++;
                released = true;
            }
            if (!ran && !released) {
                fprintf(stderr, "cuda_emu: DEADLOCK in block %u: threads wait at barriers that cannot complete\n", b);
                for (int t = 0; t < s.nthreads; t++)
                    if (s.fibers[t].state != DONE && (t % 32 == 0 || s.fibers[t].state != s.fibers[t - 1].state))
                        fprintf(stderr, "   thread %d.. state %d\n", t, s.fibers[t].state);
                abort();
            }
        }
    }
    free(s.smem);
    s.smem = nullptr;
    s.cluster = 1;
}

// ---- thread-block clusters: rank of the calling thread's CTA, its and its partner's shared memory ----
inline int cta_rank() { return S().cluster > 1 ? S().tid / S().cta_threads : 0; }
inline unsigned char *dyn_smem() { return S().smem + (size_t)cta_rank() * S().smem_stride; }
inline unsigned char *cluster_smem(int rank) { return S().smem + (size_t)rank * S().smem_stride; }
inline Dim3 thread_idx() { Dim3 d; d.x = (unsigned)(S().cluster > 1 ? S().tid % S().cta_threads : S().tid); return d; }
inline Dim3 block_idx() { Dim3 d = S().blockIdx; d.x += (unsigned)cta_rank(); return d; }

inline int lane_id() { return S().tid & 31; }
inline uint64_t *warp_slots() { return &S().slots[(size_t)(S().tid >> 5) * 32]; }

template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <class T> inline T exchange(T v, int src_lane)
{
    uint64_t *sl = warp_slots();
    sl[lane_id()] = to_bits(v);
    yield_as(AT_WARP_BARRIER);
    T r = from_bits<T>(sl[src_lane & 31]);
    yield_as(AT_WARP_BARRIER);
    return r;
}

}  // namespace emu

// ---- CUDA surface ------------------------------------------------------------------------------------
#define threadIdx (emu::thread_idx())
#define blockIdx  (emu::block_idx())
#define blockDim  (emu::S().blockDim)
#define gridDim   (emu::S().gridDim)

// bar.sync id, count  (count = number of THREADS that meet at barrier `id`, 1 <= id <= 15; the emulator keeps
// ids 16.. for the per-CTA __syncthreads of a cluster launch)
inline void emu_named_barrier(int id, int count)
{
    emu::Fiber &f = emu::S().fibers[emu::S().tid];
    f.bar_id = id; f.bar_count = count;
    emu::yield_as(emu::AT_NAMED_BARRIER);
}
// bar.arrive id, count: the calling thread counts towards barrier `id` and goes on (producer side of a producer / consumer pair)
inline void emu_named_arrive(int id, int count)
{
    (void)count;
    emu::S().bar_arrived[id]++;
}
inline void __syncthreads()
{
    if (emu::S().cluster > 1) emu_named_barrier(16 + emu::cta_rank(), emu::S().cta_threads);
    else emu::yield_as(emu::AT_BLOCK_BARRIER);
}
inline void emu_cluster_sync() { emu::yield_as(emu::AT_BLOCK_BARRIER); }     // barrier.cluster arrive + wait
inline void __syncwarp(unsigned = 0xffffffffu) { emu::yield_as(emu::AT_WARP_BARRIER); }
inline void __threadfence_block() {}
inline void __threadfence() {}

template <class T> inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    int lane = emu::lane_id();
    return emu::exchange(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
// redux.sync.or over the lanes of `mask` (the calling lane is one of them; every lane of the warp executes it with its own mask)
inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
    uint64_t *sl = emu::warp_slots();
    sl[emu::lane_id()] = emu::to_bits(v);
    emu::yield_as(emu::AT_WARP_BARRIER);
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r |= emu::from_bits<unsigned>(sl[l]);
    emu::yield_as(emu::AT_WARP_BARRIER);
    return r;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32)
{
    int lane = emu::lane_id(), src = lane ^ m;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    return emu::exchange(v, src);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32)
{
    int lane = emu::lane_id();
    int src = ((lane & (width - 1)) >= (int)d) ? lane - (int)d : lane;
    return emu::exchange(v, src);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32)
{
    int lane = emu::lane_id();
    int src = ((lane & (width - 1)) + (int)d < width) ? lane + (int)d : lane;
    return emu::exchange(v, src);
}
inline unsigned __ballot_sync(unsigned, int pred)
{
    uint64_t *sl = emu::warp_slots();
    sl[emu::lane_id()] = pred ? 1 : 0;
    emu::yield_as(emu::AT_WARP_BARRIER);
    unsigned r = 0;
    int base = emu::S().tid & ~31;
    for (int l = 0; l < 32 && base + l < emu::S().nthreads; l++)
        if (emu::S().fibers[base + l].state != emu::DONE && sl[l]) r |= 1u << l;
    emu::yield_as(emu::AT_WARP_BARRIER);
    return r;
}
inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
template <class T> inline unsigned __match_any_sync(unsigned, T v)
{
    uint64_t *sl = emu::warp_slots();
    sl[emu::lane_id()] = emu::to_bits(v);
    emu::yield_as(emu::AT_WARP_BARRIER);
    unsigned r = 0;
    int base = emu::S().tid & ~31;
    for (int l = 0; l < 32 && base + l < emu::S().nthreads; l++)
        if (emu::S().fibers[base + l].state != emu::DONE && sl[l] == emu::to_bits(v)) r |= 1u << l;
    emu::yield_as(emu::AT_WARP_BARRIER);
    return r;
}

inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh)
{
    sh &= 31;
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
template <class T> inline T __ldg(const T *p) { return *p; }
using std::min;
using std::max;
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
inline int4 make_int4(int x, int y, int z, int w) { int4 v = {x, y, z, w}; return v; }
inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p |= v; return o; }
inline int atomicExch(int *p, int v) { int o = *p; *p = v; return o; }
inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p += v; return o; }
