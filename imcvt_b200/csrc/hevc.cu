// hevc.cu -- host side of the H.265 intra encoder: workspace carving and the launch of hevc_frame_kernel
// (device code: hevc_kernel.cuh).  Replaces the codec part of writeHEVCImageFile / HEVCImageEncoder
// (reference: src/imageio_hevc.c:9-53, src/HEVCe/HEVCe.c:1569-1646).
#include "common.cuh"
#include "../../include/imcvt_b200.h"
#include "hevc_kernel.cuh"
#include <algorithm>
#include <functional>
#include <queue>
#include <vector>

namespace imcvt {
namespace hevc {

#ifdef IMCVT_HEVC_PROF
static unsigned long long *g_prof = nullptr;
#endif
static inline int pad32(uint32_t v) { uint32_t c = v > 8192 ? 8192 : v; return (int)((c + 31) / 32 * 32); }   // HEVCe.c:1580-1581

// Relative cost of one job by quantiser and the gain of giving a job two SMs (the cluster-pair kernel), measured on B200 with
// 16 "natural" 1920x1080 frames on CTA pairs (tools/quick_hevc.py, end of round 2, profiles/quick_timings_r2.txt): 9.47 / 7.35 /
// 5.29 / 1.60 s at qpd6 1 / 2 / 3 / 4 (the lower the quantiser, the more bins the CABAC trials step through); one CTA per frame
// with every SM busy 2.97 s against 1.74 s on pairs at qpd6 4.  qpd6 0: extrapolated.
static const double kJobCost[5] = {7.2, 5.9, 4.6, 3.3, 1.0};
static const double kPairGain = 1.7;

// Finishing time of jobs handed out longest first to `machines` workers (what the kernel's dynamic queue does).
static double makespan(const std::vector<double> &jobs_desc, int machines)
{
    if (machines <= 0) return 1e300;
    std::priority_queue<double, std::vector<double>, std::greater<double>> load;
    for (int m = 0; m < machines; m++) load.push(0.0);
    double end = 0.0;
    for (double j : jobs_desc) { const double t = load.top() + j; load.pop(); load.push(t); end = std::max(end, t); }
    return end;
}

// d_frames: device, n gray8 images.  d_out / d_out_len / d_rcon (optional): device.  Queues everything on `st`.
int encode_device(const uint8_t *d_frames, size_t frame_stride, int n, uint32_t H, uint32_t W, int qpd6,
                  uint8_t *d_out, size_t out_stride, uint64_t *d_out_len, uint8_t *d_rcon, size_t rcon_stride,
                  int *d_overflow, cudaStream_t st, const int *qlist, int nq)
{
    // qlist / nq: a sweep (nq > 0) -- every frame is coded at each of the nq settings; job j = frame j / nq at qlist[j % nq]
    // writes slot j of d_out / d_out_len / d_rcon.  The jobs of a sweep fill the GPU where one setting's batch would not.
    Engine &E = engine();
    const int n_frames = n;
    if (nq > 0) n = n * nq;                             // from here on n counts jobs (= output slots)
    const int hp = pad32(H), wp = pad32(W);
    const size_t rbytes = (size_t)hp * wp;
    if (d_rcon && rcon_stride < rbytes) { set_error("rcon_stride smaller than one padded picture"); return 1; }
    if (out_stride > 0xFFFFFFFFull) out_stride = 0xFFFFFFFFull;             // stream positions are 32-bit inside the kernel

    if (!E.attr_hevc) {
        IMCVT_CUDA(cudaFuncSetAttribute(hevc_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        IMCVT_CUDA(cudaFuncSetAttribute(hevc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        E.attr_hevc = true;
    }
    // Few frames: two SMs per frame (a cluster of two CTAs: the spine searches the 8x8 units and decides, the upper CTA
    // evaluates the 16x16 / 32x32 candidates meanwhile).  Many frames: one CTA per frame keeps every SM on its own frame.
    // Launch shape: one CTA per job, or a cluster of two CTAs per job (each job 1.6x faster on twice the SMs).  Few jobs: pairs.
    // Otherwise whichever finishes the queue sooner -- the jobs of a sweep differ several-fold in cost, and with one CTA per job
    // the launch lasts as long as its costliest job while the SMs of the cheap ones idle.
    bool pair = 2 * n <= E.sm_count;
    int qorder[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    if (nq > 0) std::stable_sort(qorder, qorder + nq, [&](int a, int b) { return qlist[a] < qlist[b]; });
    if (!pair && n <= 8192) {
        std::vector<double> jobs;
        jobs.reserve(n);
        if (nq > 0) { for (int r = 0; r < nq; r++) for (int f = 0; f < n_frames; f++) jobs.push_back(kJobCost[qlist[qorder[r]]]); }
        else jobs.assign(n, 1.0);
        const double t_single = makespan(jobs, E.sm_count), t_pair = makespan(jobs, E.sm_count / 2) / kPairGain;
        pair = t_pair < 0.97 * t_single;
    }
    const int shape = E.hevc_path & 3, form = E.hevc_path >> 2;          // imcvtDebugHEVCPath: launch shape | search form << 2
    if (shape == 1) pair = false;
    if (shape == 2) pair = true;
    const int grid = pair ? 2 * (n < E.sm_count / 2 ? n : E.sm_count / 2)
                          : (n < E.sm_count ? n : E.sm_count);              // one CTA per SM; frames come from a queue
    const size_t rstride = (rbytes + 255) & ~(size_t)255;
    size_t bytes = 4096 + (size_t)grid * 2 * NALL * SCR_BYTES + 256 + (size_t)grid * 2 * NALL * CAND_LIST_U16 * 2 + 256;
    if (!d_rcon) bytes += (size_t)n * rstride + 256;
    if (E.ws.reserve(bytes)) return 1;
    if (E.ws.wait_previous(st)) return 1;
    uint8_t *scratch = E.ws.take<uint8_t>((size_t)grid * 2 * NALL * SCR_BYTES);
    unsigned short *lists = E.ws.take<unsigned short>((size_t)grid * 2 * NALL * CAND_LIST_U16);
    int *next = E.ws.take<int>(1);
    uint8_t *rc = d_rcon;
    size_t rcs = rcon_stride;
    if (!rc) { rc = E.ws.take<uint8_t>((size_t)n * rstride); rcs = rstride; }
    if (!scratch || !lists || !next || !rc) { set_error("workspace carve failed"); return 1; }
    IMCVT_CUDA(cudaMemsetAsync(next, 0, sizeof(int), st));

    Params P;
    P.frames = d_frames; P.frame_stride = frame_stride;
    P.n = n_frames; P.H = (int)H; P.W = (int)W; P.hp = hp; P.wp = wp; P.qpd6 = qpd6;
    P.out = d_out; P.out_stride = out_stride;
    P.out_len = (unsigned long long *)d_out_len;
    P.rcon = rc; P.rcon_stride = rcs;
    P.scratch = scratch; P.lists = lists; P.overflow = d_overflow; P.next_frame = next;
    P.prof = nullptr;
    P.nq = nq > 0 ? nq : 0;
    for (int k = 0; k < 8; k++) P.qlist[k] = (nq > 0 && k < nq) ? qlist[k] : 0;
    for (int k = 0; k < 8; k++) P.qorder[k] = qorder[k];
    // 8x8 units: the round form (warps 0-9 walk the four 4x4 rounds together, one-TU pipelines beside round 0) on both launch
    // shapes; the three-warp-group form stays selectable as a cross-check (measured round 2, 96 frames on single CTAs: 3.38 s
    // against 3.96 s; CTA pairs, 64 frames: 2.18 s against 2.62 s)
    P.cu8_lockstep = form == 2 ? 0 : 1;
    if (nq > 0) P.qpd6 = qlist[0];
#ifdef IMCVT_HEVC_PROF
    {
        static unsigned long long *d_prof = nullptr;
        if (!d_prof) { IMCVT_CUDA(cudaMalloc(&d_prof, 128 * 8)); IMCVT_CUDA(cudaMemset(d_prof, 0, 128 * 8)); }
        P.prof = d_prof;
        g_prof = d_prof;
    }
#endif
    {
        LaunchScope ls(pair ? "hevc_pair" : "hevc_frame", st);
        if (pair) hevc_pair_kernel<<<grid, NT, sizeof(Smem), st>>>(P);
        else hevc_frame_kernel<<<grid, NT, sizeof(Smem), st>>>(P);
    }
    IMCVT_CUDA(cudaGetLastError());
    return E.ws.mark_busy(st);
}

}  // namespace hevc
}  // namespace imcvt

#ifdef IMCVT_HEVC_PROF
// development build only (make PROF=1): reads and clears the 64 phase counters of hevc_frame_kernel
extern "C" int imcvtDebugHevcProf(unsigned long long *out64)
{
    if (!imcvt::hevc::g_prof) return 1;
    cudaDeviceSynchronize();
    cudaMemcpy(out64, imcvt::hevc::g_prof, 128 * 8, cudaMemcpyDeviceToHost);
    cudaMemset(imcvt::hevc::g_prof, 0, 128 * 8);
    return 0;
}
#endif
