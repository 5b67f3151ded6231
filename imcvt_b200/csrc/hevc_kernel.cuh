// hevc_kernel.cuh -- H.265/HEVC intra-frame encoder kernel for sm_100a (device code only).
//
// Replaces HEVCImageEncoder and everything below it (reference: src/HEVCe/HEVCe.c:177-1646): 8-bit luma,
// CTU 32, CU 32/16/8, TU 32..4, 35 intra modes, exhaustive RD search driven by REAL CABAC trial encodes,
// simplified RDOQ.  Streams and reconstructions are bit-identical to the reference.
//
// Mapping: one CTA per frame walks the CTUs in raster order -- the picture has ONE
// arithmetic coder and ONE adaptive context set (:1577-1578) whose exact bit counts feed every decision (:1437),
// so CTUs of a frame form a chain.  Inside a CU the 35 modes are the parallel axis:
//   phase A  warp-cooperative TU pipelines (32/TS transform units per warp, a lane owns a column, then a row):
//            reference samples + smoothing (:196-257), prediction (:262-381), residual, forward transform
//            (:497-516) as register partial butterflies, RDOQ (:540-594), de-quantisation (:599-614), inverse
//            transform, reconstruction, SSE, last significant position / coefficient-group mask;
//   phase B  one thread per candidate (spread across warps): CABAC trial encode from the CU's entry state
//            (:1422-1423), bytes into a private scratch stream so emulation-prevention bytes count (:834);
//   decide   reference candidate order, `>=` tie rule (:1439,1475,1520,1545); the winner's coder state, bytes,
//            reconstruction and size/mode maps are committed.
//
// The file is also compiled for the host by tests/emu (a fiber SIMT emulator, test infrastructure only) so the
// kernel logic is checked against the oracle in the CPU test tier.
#pragma once
#include <stdint.h>
#include <stddef.h>

// The pair kernel encodes one frame on a cluster of two CTAs that exchange state through distributed shared memory;
// the host emulator runs the two CTAs' threads as one set of fibers with two shared-memory images.
#ifdef IMCVT_HOST_EMU
#define IMCVT_CONSTANT static const
#define IMCVT_DYN_SMEM() (emu::dyn_smem())
#define HEVC_PAIR_RANK() (emu::cta_rank())
#define HEVC_PAIR_SMEM(r) (emu::cluster_smem(r))
#define HEVC_PAIR_SYNC() emu_cluster_sync()
#define HEVC_CLUSTER_DIMS
#define HEVC_SMEM(K) (*(K).sm)
#else
#include <cooperative_groups.h>
#define IMCVT_CONSTANT __constant__ const
extern __shared__ __align__(16) unsigned char imcvt_dyn_smem[];
#define IMCVT_DYN_SMEM() (imcvt_dyn_smem)
#define HEVC_PAIR_RANK() ((int)cooperative_groups::this_cluster().block_rank())
#define HEVC_PAIR_SMEM(r) (cooperative_groups::this_cluster().map_shared_rank(imcvt_dyn_smem, (r)))
#define HEVC_PAIR_SYNC() cooperative_groups::this_cluster().sync()
#define HEVC_CLUSTER_DIMS __cluster_dims__(2, 1, 1)
// A CTA's own shared memory, named through the __shared__ symbol in every function that uses it: read through the pointer kept
// in Ctx it is a generic address, and every access becomes a 64-bit generic load / store instead of LDS / STS.
#define HEVC_SMEM(K) (*reinterpret_cast<Smem *>(imcvt_dyn_smem))
#endif

namespace imcvt {
namespace hevc {

// development aid: cycles of thread 0 between phase marks, accumulated per (CU level, phase) in Params::prof
#if defined(IMCVT_HEVC_PROF) && !defined(IMCVT_HOST_EMU)
#define HPROF(K, id) do { if (threadIdx.x == 0 && (K).prof_on) { const long long t__ = clock64(); atomicAdd(&(K).P->prof[(id)], (unsigned long long)(t__ - (K).prof_t)); (K).prof_t = t__; } } while (0)
// stage counters inside the TU pipeline distort the phase counters: a separate switch (make PROF=1 TPROF=1)
#ifdef IMCVT_HEVC_TPROF
#define TPROF(sm, id) do { if (threadIdx.x == 0) { const long long t__ = clock64(); atomicAdd(&(sm).prof[(id)], (unsigned long long)(t__ - (sm).prof_t)); (sm).prof_t = t__; } } while (0)
#define TPROF_START(sm) do { if (threadIdx.x == 0) (sm).prof_t = clock64(); } while (0)
#else
#define TPROF(sm, id) do { } while (0)
#define TPROF_START(sm) do { } while (0)
#endif
#else
#define HPROF(K, id) do { } while (0)
#define TPROF(sm, id) do { } while (0)
#define TPROF_START(sm) do { } while (0)
#endif

// bar.sync id, count: `count` threads (whole warps) meet at barrier `id`.  bar.arrive: a producer warp counts towards the barrier
// and goes on; only the consumers wait.  The fence in front of the arrive orders the producer's global-memory stores (bin lists)
// as the barrier itself orders shared memory.
#ifdef IMCVT_HOST_EMU
#define HEVC_NAMED_BARRIER(id, count) emu_named_barrier(id, count)
#define HEVC_NAMED_ARRIVE(id, count) emu_named_arrive(id, count)
#else
#define HEVC_NAMED_BARRIER(id, count) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory")
#define HEVC_NAMED_ARRIVE(id, count) do { __threadfence_block(); asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); } while (0)
#endif
// the transform-unit pipeline, the binariser and the block coder are inlined into the phases that use them
#define HEVC_BIGFN __device__ __forceinline__
// the phases of a coding unit are real functions: one copy of each in the instruction cache however often they are used
#ifdef IMCVT_HOST_EMU
#define HEVC_PHASE __device__
#else
#define HEVC_PHASE __device__ __noinline__
#endif

#include "hevc_tables.h"

constexpr int NW = 12;                  // warps per CTA
constexpr int NT = NW * 32;
constexpr int NMODE = 35;
constexpr int CTU = 32;
constexpr int RCP = 72;                 // row pitch of the CTU reconstruction window (1 + 64 columns used)
constexpr int LEV_STRIDE = 1152;        // int16 per candidate: one 32x(32+2) block or four 16x(16+2) blocks
constexpr int REC_STRIDE = 1024;        // bytes per candidate reconstruction tile
constexpr int WS_STRIDE = 32 * 34;      // int16 per warp: transpose scratch
constexpr int REFS_BYTES = 448;         // per warp: reference samples of its transform units
constexpr int CXN = 142, CXP = 144;     // context bytes (padded to words)
constexpr int SCR_BYTES = 8192;         // private trial stream per candidate (reference bound: 3200, :794)
constexpr int I32MAX = 0x7fffffff;
constexpr int SMEM_LIMIT = 232448;     // 227 KB: the most dynamic shared memory a CTA can ask for on sm_100
constexpr int NSCAN = 1520;
constexpr int NCAND = 2 * NMODE;        // candidates coded together: 35 one-TU trials + 35 four-TU trials
constexpr int NXN = NCAND + NMODE;      // 8x8 CUs: slots NCAND..NXN-1 = stand-alone 4x4 trials of the round in flight, slot NXN = the NxN trial
constexpr int NALL = NXN + 1;
constexpr int PRE_MAX = 40;             // entries of a trial's prefix list: head (<= 9; NxN 16) + <= 4 cbf flags + last position (<= 24; 4x4: 6)
constexpr int SLOT = 112;               // bin-list entries per coefficient group (worst case 107)
constexpr int CAND_LIST_U16 = 64 * SLOT + 64;   // 64 group slots + 64 group headers per candidate
// Where the bin lists of a candidate live (entries of group g at CAND * (lid - BASE) + g * SLOT, its header at ... + HDR + g):
//   ListsGlobal  the per-CTA lists in global memory, any transform size
//   ListsCu8     the 4x4 blocks of an 8x8 unit's four-TU and stand-alone candidates (lid NMODE .. NXN-1, four groups each) in
//                shared memory: they are written, read by the trials of the same round and dropped -- in global memory every one
//                of those steps was an L2 round trip on the unit's critical chain.  The space is rec0, which only the 32x32
//                evaluation uses and which is dead while 8x8 units are searched (both kernels).
struct ListsGlobal { static constexpr int CAND = CAND_LIST_U16, HDR = 64 * SLOT, BASE = 0; };
struct ListsCu8 { static constexpr int CAND = 4 * SLOT + 8, HDR = 4 * SLOT, BASE = 35; };

// context offsets: member order of the reference's ContextSet (:744-758)
enum { CX_SPLIT_CU = 0, CX_PART = 3, CX_YMODE = 4, CX_UVMODE = 5, CX_SPLIT_TU = 6, CX_YCBF = 9, CX_UVCBF = 11,
       CX_LASTX = 16, CX_LASTY = 41, CX_SIGCG = 66, CX_SIG = 68, CX_ONE = 112, CX_ABS = 136 };

IMCVT_CONSTANT int c_dst4[4][4] = { {29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29} };   // :391-396
IMCVT_CONSTANT signed char c_angle[35] = {0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26, -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32};
IMCVT_CONSTANT short c_invang[35] = {0, 0, 256, 315, 390, 482, 630, 910, 1638, 4096, 0, 4096, 1638, 910, 630, 482, 390, 315, 256, 315, 390, 482, 630, 910, 1638, 4096, 0, 4096, 1638, 910, 630, 482, 390, 315, 256};
IMCVT_CONSTANT int c_wd[5] = {11, 11, 11, 5, 1};             // calcRDcost weights :178-179
IMCVT_CONSTANT int c_wb[5] = {1, 4, 16, 29, 23};

// CABAC tables (:700-714); copied to shared memory at kernel start because lanes index them divergently
IMCVT_CONSTANT unsigned char c_next_lps[128] = {1,0,0,1,2,3,4,5,4,5,8,9,8,9,10,11,12,13,14,15,16,17,18,19,18,19,22,23,22,23,24,25,26,27,26,27,30,31,30,31,32,33,32,33,36,37,36,37,38,39,38,39,42,43,42,43,44,45,44,45,46,47,48,49,48,49,50,51,52,53,52,53,54,55,54,55,56,57,58,59,58,59,60,61,60,61,60,61,62,63,64,65,64,65,66,67,66,67,66,67,68,69,68,69,70,71,70,71,70,71,72,73,72,73,72,73,74,75,74,75,74,75,76,77,76,77,126,127};
IMCVT_CONSTANT unsigned char c_lps[64][4] = {
  {128,176,208,240},{128,167,197,227},{128,158,187,216},{123,150,178,205},{116,142,169,195},{111,135,160,185},{105,128,152,175},{100,122,144,166},
  { 95,116,137,158},{ 90,110,130,150},{ 85,104,123,142},{ 81, 99,117,135},{ 77, 94,111,128},{ 73, 89,105,122},{ 69, 85,100,116},{ 66, 80, 95,110},
  { 62, 76, 90,104},{ 59, 72, 86, 99},{ 56, 69, 81, 94},{ 53, 65, 77, 89},{ 51, 62, 73, 85},{ 48, 59, 69, 80},{ 46, 56, 66, 76},{ 43, 53, 63, 72},
  { 41, 50, 59, 69},{ 39, 48, 56, 65},{ 37, 45, 54, 62},{ 35, 43, 51, 59},{ 33, 41, 48, 56},{ 32, 39, 46, 53},{ 30, 37, 43, 50},{ 29, 35, 41, 48},
  { 27, 33, 39, 45},{ 26, 31, 37, 43},{ 24, 30, 35, 41},{ 23, 28, 33, 39},{ 22, 27, 32, 37},{ 21, 26, 30, 35},{ 20, 24, 29, 33},{ 19, 23, 27, 31},
  { 18, 22, 26, 30},{ 17, 21, 25, 28},{ 16, 20, 23, 27},{ 15, 19, 22, 25},{ 14, 18, 21, 24},{ 14, 17, 20, 23},{ 13, 16, 19, 22},{ 12, 15, 18, 21},
  { 12, 14, 17, 20},{ 11, 14, 16, 19},{ 11, 13, 15, 18},{ 10, 12, 15, 17},{ 10, 12, 14, 16},{  9, 11, 13, 15},{  9, 11, 12, 14},{  8, 10, 12, 14},
  {  8,  9, 11, 13},{  7,  9, 11, 12},{  7,  9, 10, 12},{  7,  8, 10, 11},{  6,  8,  9, 11},{  6,  7,  9, 10},{  6,  7,  8,  9},{  2,  2,  2,  2}};
IMCVT_CONSTANT unsigned char c_ctx_init[CXN] = {           // newContextSet :762-776
    139, 141, 157,  184,  184,  63,  153, 138, 138,  111, 141,  94, 138, 182, 154, 154,
    110, 110, 124, 0, 0,  125, 140, 153, 0, 0,  125, 127, 140, 109, 0,  111, 143, 127, 111, 79,  108, 123, 63, 154, 0,
    110, 110, 124, 0, 0,  125, 140, 153, 0, 0,  125, 127, 140, 109, 0,  111, 143, 127, 111, 79,  108, 123, 63, 154, 0,
    91, 171,
    111, 111, 125, 110, 110, 94, 124, 108, 124, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125, 141, 140, 139, 182, 182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111, 111,
    140, 92, 137, 138, 140, 152, 138, 139, 153, 74, 149, 92, 139, 107, 122, 152, 140, 179, 166, 182, 140, 227, 122, 197,
    138, 153, 136, 167, 152, 152 };
// fixed parameter sets (putHeaderToBuffer :664-690)
IMCVT_CONSTANT unsigned char c_vps[27] = {0x00,0x00,0x01,0x40,0x01,0x0C,0x01,0xFF,0xFF,0x03,0x10,0x00,0x00,0x03,0x00,0x00,0x03,0x00,0x00,0x03,0x00,0x00,0x03,0x00,0xB4,0xF0,0x24};
IMCVT_CONSTANT unsigned char c_sps[22] = {0x00,0x00,0x01,0x42,0x01,0x01,0x03,0x10,0x00,0x00,0x03,0x00,0x00,0x03,0x00,0x00,0x03,0x00,0x00,0x03,0x00,0xB4};
IMCVT_CONSTANT unsigned char c_pps[11] = {0x00,0x00,0x01,0x44,0x01,0xC0,0x90,0x91,0x81,0xD9,0x20};
IMCVT_CONSTANT unsigned char c_slice_head[6] = {0x00,0x00,0x01,0x26,0x01,0xAC};
IMCVT_CONSTANT unsigned char c_slice_tail[5][2] = {{0x16,0xDE},{0x10,0xDE},{0x2B,0x78},{0x4D,0xE0},{0x97,0x80}};

// ------------------------------------------------------------------------------------------------
// kernel parameters and shared-memory layout
// ------------------------------------------------------------------------------------------------
struct Params {
    const uint8_t *frames; size_t frame_stride;     // n gray8 images H x W
    int n, H, W, hp, wp, qpd6;                       // hp, wp: padded to x32 (and clipped to 8192, :1580-1581)
    int nq, qlist[8];                                // nq > 0: a sweep -- job j codes frame j / nq at qpd6 = qlist[j % nq] into slot j
    int cu8_lockstep;                                // 8x8 units: 1 = all warps walk the 4x4 rounds in lock step, 0 = three warp groups side by side
    int qorder[8];                                   // ... the queue hands the jobs out setting by setting, qorder[0] first (the costliest: lowest qpd6)
    uint8_t *out; size_t out_stride;                 // output slots
    unsigned long long *out_len;                     // [n]
    uint8_t *rcon; size_t rcon_stride;               // [n][hp*wp] reconstruction (also the line buffer between CTU rows)
    uint8_t *scratch;                                // [gridDim.x][2][NALL][SCR_BYTES] private trial streams (per candidate set)
    unsigned short *lists;                           // [gridDim.x][2][NALL][CAND_LIST_U16] bin lists of the trials
    int *overflow;                                   // set to 1 when a stream did not fit its slot
    int *next_frame;                                 // dynamic frame queue
    unsigned long long *prof;                        // phase cycle counters (only with IMCVT_HEVC_PROF; else unused)
};

struct CoderRegs { int range, low, nbits, pending, held, zeros; unsigned nout; int pad_; };
struct CuPar { int y0, x0, avail, split_ctx, mode_l, mode_a; };

struct BinDesc { int active, lid, slot0, mode, lev_off, last; unsigned sg0, sg1; int set, pad_; };   // one transform unit to binarise

struct Smem {
    short lev[NW][LEV_STRIDE];                      // quantised levels of the units a warp has in flight (binarised at once)
    short ws[NW][WS_STRIDE];                        // per-warp transpose scratch
    // Candidate state exists twice: set 0 belongs to the 32x32 level, set 1 to the 16x16 and 8x8 levels.  One CTA walks
    // the levels one after the other; in the two-CTA kernel the upper CTA keeps a 32x32 and a 16x16 evaluation in flight.
    uint8_t rec0[NCAND][REC_STRIDE];                // candidate prediction -> reconstruction tiles, set 0
    uint8_t rec1[NCAND][REC_STRIDE / 4];            // ... set 1 (16x16 and smaller)
    uint8_t rec4[NMODE][16];                        // 8x8 CUs: stand-alone 4x4 trials of the round in flight
    uint8_t refs[NW][REFS_BYTES];
    BinDesc bdesc[NW][8];
    uint8_t cctx[2][NALL][CXP];                     // candidate context sets
    CoderRegs ccoder[2][NALL];
    int cbits[2][NALL], csse[2][NALL];
    int clast[2][NALL][4];
    unsigned ccg[2][NALL][4][2];                    // significant groups by position: bit yc*8+xc
    unsigned csg[2][NALL][4][2];                    // significant groups by scan index
    int sse0[NMODE];                                // 8x8 CUs: squared error of the first 4x4 block per mode
    unsigned short tpre[NALL][PRE_MAX];             // per trial: syntax waiting in front of a block's coefficient groups (PreList)
    // CTU state
    uint8_t orig[CTU][CTU];
    uint8_t rc[CTU + 1][RCP];                       // reconstruction window: row 0 / column 0 are the neighbours (:1588-1589)
    uint8_t msz[9][12], mmd[9][12];                 // CU size / luma mode per 4x4 cell; row 0 / column 0 = above / left CTU
    uint8_t left_sz[8], left_md[8];                 // right-most column of the previous CTU
    uint8_t nxn_rec[8 * 8];
    alignas(16) unsigned short nxn_list[4][SLOT];   // bin lists of the four kept 4x4 blocks
    unsigned short nxn_hdr[4];
    int nxn_mode[4], nxn_last[4];
    unsigned nxn_cg[4];
    CoderRegs cur, entry[3];
    uint8_t cur_ctx[CXP], entry_ctx[3][CXP], init_ctx[CXP];
    int best_cost[3];
    int win, red[NW];
    CuPar par[2];                                   // 32x32 / 16x16 unit whose candidates are being evaluated
    int eval_cost[2], eval_win[2];                  // ... the best of its 70 candidates (reference order, later wins ties)
    // two-CTA kernel: the winner of the partner's evaluation, written by the partner through distributed shared memory
    CoderRegs res_coder[2];
    uint8_t res_ctx[2][CXP];
    uint8_t res_rec0[REC_STRIDE], res_rec1[REC_STRIDE / 4];
    unsigned lpsw[64];                              // LPS range table, one word per state (4 range classes)
    uint8_t next_lps[128];
    uint8_t scan[NSCAN][2];                         // coefficient scans (y, x): see scan_base()
    uint8_t dord[4][64][2];                         // up-right diagonal orders of 1, 2x2, 4x4, 8x8 grids
    uint8_t above_sz[2048];                         // CU sizes along the bottom of the CTU row above (4x4 cells)
#if defined(IMCVT_HEVC_PROF) && !defined(IMCVT_HOST_EMU)
    unsigned long long *prof; long long prof_t;     // development build: stage counters of thread 0
#endif
};

static_assert(sizeof(Smem) <= SMEM_LIMIT, "hevc Smem does not fit the 227 KB a CTA can have");

// the bin lists of an 8x8 unit's 4x4 blocks (ListsCu8): (NXN - NMODE) candidates x ListsCu8::CAND entries in the space of rec0
static_assert((NXN - NMODE) * ListsCu8::CAND * 2 <= NCAND * REC_STRIDE, "ListsCu8 does not fit rec0");
static_assert((ListsCu8::CAND * 2) % 16 == 0 && (SLOT * 2) % 16 == 0, "group slots must stay 16-byte aligned");
__device__ __forceinline__ unsigned short *cu8_lists(Smem &sm) { return reinterpret_cast<unsigned short *>(&sm.rec0[0][0]); }

// a pointer into this CTA's shared memory that came through memory (a struct member, a function argument) and is therefore a
// generic address: the same address, derived from the __shared__ symbol so that the compiler knows the address space
template <class T> __device__ __forceinline__ T *own_smem(T *p)
{
    unsigned char *base = IMCVT_DYN_SMEM();
    return reinterpret_cast<T *>(base + (unsigned)(reinterpret_cast<unsigned char *>(p) - base));
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int clip16(int v) { return clampi(v, -32768, 32767); }
__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }
template <int TS> struct SizeClass { static constexpr int v = TS == 4 ? 0 : TS == 8 ? 1 : TS == 16 ? 2 : 3; };
__device__ __forceinline__ int size_class(int sz) { return sz == 4 ? 0 : sz == 8 ? 1 : sz == 16 ? 2 : 3; }

// calcRDcost :177-185 with the two divisions (I32MAX / weight) taken out of the hot loops: they only depend on qpd6
struct RdK {
    int wd, wb, limd, limb;
    __device__ __forceinline__ void init(int qpd6) { wd = c_wd[qpd6]; wb = c_wb[qpd6]; limd = I32MAX / wd; limb = I32MAX / wb; }
    __device__ __forceinline__ int cost(int dist, int bits) const
    {
        const int c1 = (limd <= dist) ? I32MAX : wd * dist;
        const int c2 = (limb <= bits) ? I32MAX : wb * bits;
        return (I32MAX - c1 <= c2) ? I32MAX : c1 + c2;
    }
};

IMCVT_CONSTANT int c_level_rate[6] = {0, 70000, 90000, 92000, 157536, 190304};
__device__ __forceinline__ int level_rate(int level)                                   // estimateCoeffRate :526-535
{
    if (level < 6) return c_level_rate[level];
    const int i = 31 - __clz(level - 5);                // number of exp-Golomb prefix steps for level - 6
    return 92000 + ((3 + i * 2 + 1) << 15);
}

__device__ __forceinline__ int scan_type_for(int sz, int mode)                         // getScanOrder :1133-1150
{
    if (sz <= 8) {
        if (iabs(mode - 26) <= 4) return 1;
        if (iabs(mode - 10) <= 4) return 2;
    }
    return 0;
}

// start of a scan table inside Smem::scan: diagonal 4/8/16/32, horizontal 4/8, vertical 4/8
__device__ __forceinline__ int scan_base(int st, int sc)
{
    if (st == 0) return sc == 0 ? 0 : sc == 1 ? 16 : sc == 2 ? 80 : 336;
    return (st == 1 ? 1360 : 1440) + (sc == 0 ? 0 : 16);
}

__device__ __forceinline__ bool uses_smoothed(int sz, int mode)                        // :274-280
{
    if (sz == 4 || mode == 1) return false;
    if (sz == 8) return mode == 0 || mode == 2 || mode == 18 || mode == 34;
    if (sz == 16) return !(mode >= 9 && mode <= 11) && !(mode >= 25 && mode <= 27);
    return mode != 10 && mode != 26;
}

// ------------------------------------------------------------------------------------------------
// CABAC engine (one thread).  State in registers, contexts and tables in shared memory.
// ------------------------------------------------------------------------------------------------
struct Coder {
    int range, low, nbits, pending, held, zeros;
    unsigned nout, cap;
    uint8_t *ctx, *out;
    const unsigned *lps;
    const uint8_t *nlps;

    __device__ __forceinline__ void load(const CoderRegs &r) { range = r.range; low = r.low; nbits = r.nbits; pending = r.pending; held = r.held; zeros = r.zeros; nout = r.nout; }
    __device__ __forceinline__ void store(CoderRegs &r) const { r.range = range; r.low = low; r.nbits = nbits; r.pending = pending; r.held = held; r.zeros = zeros; r.nout = nout; }
    __device__ __forceinline__ void fresh() { range = 510; low = 0; nbits = 23; pending = 0; held = 0xFF; zeros = 0; nout = 0; }     // newCABACcoder :808

    __device__ __forceinline__ void byte(int b)                                        // CABACput :820-831
    {
        b &= 0xFF;
        if (zeros >= 2 && b <= 3) { if (nout < cap) out[nout] = 3; nout++; zeros = 0; }
        if (nout < cap) out[nout] = (uint8_t)b;
        nout++;
        zeros = b == 0 ? zeros + 1 : 0;
    }
    __device__ __forceinline__ int bits() const { return 8 * ((int)nout + pending) + 23 - nbits; }      // CABAClen :834
    __device__ __forceinline__ void shift_out()                                        // CABACupdate :858-878
    {
        if (nbits >= 12) return;
        const int lead = low >> (24 - nbits);
        nbits += 8;
        low &= (int)(0xFFFFFFFFu >> nbits);
        if (lead == 0xFF) {
            pending++;
        } else if (pending > 0) {
            const int carry = lead >> 8;
            byte(held + carry);
            held = lead & 0xFF;
            for (; pending > 1; pending--) byte(0xFF + carry);
        } else {
            pending = 1;
            held = lead;
        }
    }
    // Shared-memory accesses of the bin step.  On the device they are ld / st.shared on 32-bit addresses kept in registers
    // (ctx_s, lps_s, nlps_s, set by bind_shared): left to the compiler, every access re-derived the CTA's shared window base
    // from a special register inside the bin loop -- on the one dependent chain a trial lane has.
#ifdef IMCVT_HOST_EMU
    __device__ __forceinline__ void bind_shared() {}
    __device__ __forceinline__ int ld_ctx(int ci) const { return ctx[ci]; }
    __device__ __forceinline__ void st_ctx(int ci, int v) { ctx[ci] = (uint8_t)v; }
    __device__ __forceinline__ unsigned ld_lps(int row) const { return lps[row]; }
    __device__ __forceinline__ int ld_nlps(int st) const { return nlps[st]; }
#else
    unsigned ctx_s, lps_s, nlps_s;
    __device__ __forceinline__ void bind_shared()
    {
        ctx_s = (unsigned)__cvta_generic_to_shared(ctx); lps_s = (unsigned)__cvta_generic_to_shared(lps); nlps_s = (unsigned)__cvta_generic_to_shared(nlps);
        // opaque to the optimiser: otherwise it re-materialises the table addresses (special-register read + shift + add) per bin
        asm volatile("mov.u32 %0, %0;" : "+r"(lps_s));
        asm volatile("mov.u32 %0, %0;" : "+r"(nlps_s));
    }
    __device__ __forceinline__ int ld_ctx(int ci) const { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(ctx_s + (unsigned)ci) : "memory"); return (int)v; }
    __device__ __forceinline__ void st_ctx(int ci, int v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(ctx_s + (unsigned)ci), "r"(v) : "memory"); }
    __device__ __forceinline__ unsigned ld_lps(int row) const { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(lps_s + 4u * (unsigned)row)); return v; }
    __device__ __forceinline__ int ld_nlps(int st) const { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(nlps_s + (unsigned)st)); return (int)v; }
#endif
    __device__ __forceinline__ void bin(int b, int ci)                                 // CABACputBin :913-932
    {
        const int st = ld_ctx(ci);
        const int l = (int)((ld_lps(st >> 1) >> (((range >> 6) & 3) * 8)) & 0xFFu);
        range -= l;
        if ((b != 0) != ((st & 1) != 0)) {
            const int n = __clz(l >> 3) - 26;           // renormalisation shift (table :712-714 in closed form)
            st_ctx(ci, ld_nlps(st));
            low = (low + range) << n;
            range = l << n;
            nbits -= n;
        } else {
            st_ctx(ci, st < 124 ? st + 2 : st);
            if (range < 256) { low <<= 1; range <<= 1; nbits--; }
        }
        shift_out();
    }
    __device__ __forceinline__ void bypass(int bins, int len)                          // CABACputBins :898-910
    {
        bins &= (1 << len) - 1;
        while (len > 0) {
            const int n = len < 8 ? len : 8;
            len -= n;
            low = (low << n) + range * ((bins >> len) & ((1 << n) - 1));
            nbits -= n;
            shift_out();
        }
    }
    __device__ __forceinline__ void bypass_chunk(int val, int n)                       // one step of CABACputBins: n <= 8 bins
    {
        low = (low << n) + range * val;
        nbits -= n;
        shift_out();
    }
    __device__ __forceinline__ void terminate(int b)                                   // CABACputTerminate :881-895
    {
        range -= 2;
        if (b) { low = (low + range) << 7; range = 256; nbits -= 7; }
        else if (range < 256) { low <<= 1; range <<= 1; nbits--; }
        shift_out();
    }
    __device__ __forceinline__ void finish()                                           // CABACfinish :839-855
    {
        int fill = 0;
        if ((low >> (32 - nbits)) > 0) {
            byte(held + 1);
            low -= 1 << (32 - nbits);
        } else {
            if (pending > 0) byte(held);
            fill = 0xFF;
        }
        for (; pending > 1; pending--) byte(fill);
        const int tail = (low >> 8) << nbits;
        byte(tail >> 16); byte(tail >> 8); byte(tail);
    }
};

// ------------------------------------------------------------------------------------------------
// syntax elements (one thread)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mpm_list(int l, int a, int m[3])                       // getProbablePmodes :957-976
{
    if (l != a) {
        m[0] = l; m[1] = a;
        m[2] = (l != 0 && a != 0) ? 0 : (l + a < 2 ? 26 : 1);
    } else if (l > 1) {
        m[0] = l; m[1] = ((l + 29) % 32) + 2; m[2] = ((l - 1) % 32) + 2;
    } else { m[0] = 0; m[1] = 1; m[2] = 26; }
}

// The lanes of a warp code different candidates: if each syntax element called the arithmetic coder where it is
// produced, lanes would sit at different call sites and run one after the other.  So the syntax in front of a block's
// coefficient groups (coding-unit head, cbf flags, last position) is first written as bin-list entries into a small
// per-trial list, and the coder is only ever stepped from the loop of code_block / emit_list, where the lanes converge.
__device__ __forceinline__ unsigned pack4(const int *v) { return (unsigned)v[0] | ((unsigned)v[1] << 8) | ((unsigned)v[2] << 16) | ((unsigned)v[3] << 24); }
struct PreList {
    unsigned short *p; int n;
    __device__ __forceinline__ void ctx(int ci, int bin) { p[n++] = (unsigned short)((bin << 8) | ci); }                       // = ent_ctx(0, ci, bin)
    __device__ __forceinline__ void byp(int val, int len) { p[n++] = (unsigned short)((3 << 14) | (len << 8) | (val & ((1 << len) - 1))); }   // = ent_byp, len <= 8
};

__device__ __forceinline__ int pos_group(int v)        // last-position group index (:1046)
{
    return v < 4 ? v : v < 6 ? 4 : v < 8 ? 5 : v < 12 ? 6 : v < 16 ? 7 : v < 24 ? 8 : 9;
}
__device__ __forceinline__ int group_min(int g)        // :1047
{
    return g < 4 ? g : g == 4 ? 4 : g == 5 ? 6 : g == 6 ? 8 : g == 7 ? 12 : g == 8 ? 16 : 24;
}

__device__ inline void build_last_pos(PreList &L, int sz, int scan_type, int y, int x)     // putLastSignificantXY :1045-1086 (luma)
{
    const int sc = size_class(sz), base = sc * 5, sft = sc == 0 ? 0 : 1;
    int ty = scan_type == 2 ? x : y, tx = scan_type == 2 ? y : x;
    const int gy = pos_group(ty), gx = pos_group(tx), gmax = pos_group(sz - 1);
    int i;
    for (i = 0; i < gx; i++) L.ctx(CX_LASTX + base + (i >> sft), 1);
    if (gx < gmax) L.ctx(CX_LASTX + base + (gx >> sft), 0);
    for (i = 0; i < gy; i++) L.ctx(CX_LASTY + base + (i >> sft), 1);
    if (gy < gmax) L.ctx(CX_LASTY + base + (gy >> sft), 0);
    if (gx > 3) { tx -= group_min(gx); for (i = ((gx - 2) >> 1) - 1; i >= 0; i--) L.byp((tx >> i) & 1, 1); }
    if (gy > 3) { ty -= group_min(gy); for (i = ((gy - 2) >> 1) - 1; i >= 0; i--) L.byp((ty >> i) & 1, 1); }
}

__device__ __forceinline__ int sig_ctx_index(int sz, int scan_type, int y, int x, int nb)      // getSigCtxIdx :1091-1121 (luma)
{
    if (y == 0 && x == 0) return 0;
    if (sz == 4) {
        // {{0,1,4,5},{2,3,4,5},{6,6,8,8},{7,7,8,8}} packed 4 bits per entry, row-major
        const unsigned long long map4 = 0x8877886654325410ull;
        return (int)((map4 >> (4 * (y * 4 + x))) & 15);
    }
    int ci = 9;
    if (sz >= 16) ci += 12;
    if (sz == 8 && scan_type != 0) ci += 6;
    if ((y >> 2) || (x >> 2)) ci += 3;
    const int py = y & 3, px = x & 3;
    int s;
    if (nb == 0) s = py + px; else if (nb == 1) s = py << 1; else if (nb == 2) s = px << 1; else return ci + 2;
    return ci + (s == 0 ? 2 : s < 3 ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------
// Bin lists.  The residual syntax of a transform unit (putCoef :1172-1268) is binarised in parallel, one
// coefficient group (16 scan positions) per half-warp, into a fixed slot of 16-bit entries; the arithmetic
// coder lane of the trial then only walks entries.  Entry: bits 15..14 type
//   0 context bin: bits 7..0 context index, bit 8 value
//   1 context bin of the greater-than-1 class, 2 of the greater-than-2 class: as 0, but the context set of these
//     depends on the PREVIOUS coded group (c1 == 0 there, :1229-1232); the coder lane adds 4 / 1 when that holds
//   3 bypass chunk: bits 7..0 value, bits 11..8 length (1..8) -- chunked exactly like CABACputBins (:898-910)
// Header per group: bits 6..0 entry count, bit 7 group has non-zero levels, bit 8 one of its first 8 levels is > 1.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned short ent_ctx(int cls, int ci, int bin) { return (unsigned short)((cls << 14) | (bin << 8) | ci); }
__device__ __forceinline__ unsigned short ent_byp(int val, int len) { return (unsigned short)((3 << 14) | (len << 8) | val); }

// number of chunk entries of one CABACputBins(value, len) call
__device__ __forceinline__ int chunks_of(int len) { return (len + 7) >> 3; }
__device__ __forceinline__ void emit_bypass(unsigned short *dst, int &o, int val, int len)
{
    val &= (1 << len) - 1;
    while (len > 0) {
        const int n = len < 8 ? len : 8;
        len -= n;
        dst[o++] = ent_byp((val >> len) & ((1 << n) - 1), n);
    }
}

// One coefficient group by the 16 lanes of a half-warp; lane hl codes scan position cg*16 + 15 - hl, so lane order is
// coding order.  Executed by all 32 lanes (ballots / shuffles inside); `has` masks a half without a task.
template <class LL>
__device__ __forceinline__ void binarize_cg(bool has, const BinDesc &D, int sz, int cg, int hl, int half,
                                            const uint8_t (*scan_all)[2], const short *levbase,
                                            unsigned cgp0, unsigned cgp1, unsigned short *lists)
{
    const int st = scan_type_for(sz, D.mode), ncg = sz >> 2, lp = sz + 2;
    const uint8_t (*scan)[2] = scan_all + scan_base(st, size_class(sz));
    const int i = cg * 16 + 15 - hl;
    const int y = scan[i][0], x = scan[i][1], yc = y >> 2, xc = x >> 2;
    const bool valid = has && i <= D.last;
    const int v = valid ? (int)levbase[D.lev_off + y * lp + x] : 0;
    const bool nzf = v != 0;
    const int mag = iabs(v);
    const bool first_cg = cg == 0;
    const unsigned lt = (1u << hl) - 1u;
    const int sh = half * 16;

    const unsigned nzball = (__ballot_sync(0xffffffffu, nzf) >> sh) & 0xFFFFu;
    const int nzidx = __popc(nzball & lt), nz = __popc(nzball);
    // significant_coeff_flag
    const int bpos = yc * 8 + xc, br = bpos + 1, bb = bpos + 8;
    const int cgs = (int)((bpos < 32 ? cgp0 >> bpos : cgp1 >> (bpos - 32)) & 1u);
    const int right = xc < ncg - 1 && ((br < 32 ? cgp0 >> br : cgp1 >> (br - 32)) & 1u);
    const int below = yc < ncg - 1 && ((bb < 32 ? cgp0 >> bb : cgp1 >> (bb - 32)) & 1u);
    const int nb = (below << 1) | right;
    const bool sig_coded = valid && i != D.last && (first_cg || (cgs && (hl != 15 || nzidx > 0)));
    const unsigned sball = (__ballot_sync(0xffffffffu, sig_coded) >> sh) & 0xFFFFu;
    const int n1 = __popc(sball);
    // greater-than-1 flags of the first 8 levels, greater-than-2 flag of the first level > 1
    const bool big = nzf && mag > 1 && nzidx < 8;
    const unsigned bball = (__ballot_sync(0xffffffffu, big) >> sh) & 0xFFFFu;
    const bool earlier_big = (bball & lt) != 0;
    const bool firstbig = big && !earlier_big;
    const unsigned g2ball = (__ballot_sync(0xffffffffu, firstbig && mag > 2) >> sh) & 0xFFFFu;
    const int anybig = bball != 0;
    const int n2 = n1 + (nz < 8 ? nz : 8), n3 = n2 + anybig;
    const bool escape = nz > 8 || __popc(bball) > 1 || g2ball != 0;
    // sign bits, first coded level in the most significant position
    int signs = (nzf && v < 0) ? 1 << (nz - 1 - nzidx) : 0;
    signs = (int)__reduce_or_sync(0xFFFFu << sh, (unsigned)signs);       // one redux over the half-warp instead of four shuffle steps
    const int n4 = n3 + (nz > 8 ? 2 : nz > 0 ? 1 : 0);

    // remaining absolute levels: Rice parameter and base level evolve along the group (:1252-1263)
    int my_rice = 0, my_base = 3;
    if (__any_sync(0xffffffffu, escape && has)) {
        int rice = 0, base = 3, seen = 0;
#pragma unroll 4
        for (int q = 0; q < 16; q++) {
            const int mq = __shfl_sync(0xffffffffu, mag, q, 16);
            if (q == hl) { my_rice = rice; my_base = base; }
            if (mq > 0) {
                const int ev = mq - (seen < 8 ? base : 1);
                if (ev >= 0 && mq > (3 << rice)) rice = rice < 4 ? rice + 1 : 4;
                if (mq >= 2) base = 2;
                seen++;
            }
        }
    }
    int lenA = 0, valA = 0, lenB = 0, valB = 0;
    if (escape && nzf) {
        const int ev = mag - (nzidx < 8 ? my_base : 1);
        if (ev >= 0) {                                                          // putRemainExGolomb :1153-1168
            if (ev < (3 << my_rice)) {
                const int len = ev >> my_rice;
                lenA = len + 1; valA = (1 << (len + 1)) - 2;
                lenB = my_rice; valB = ev & ((1 << my_rice) - 1);
            } else {
                int len = my_rice, w = ev - (3 << my_rice);
                for (; w >= (1 << len); len++) w -= 1 << len;
                lenA = 4 + len - my_rice; valA = (1 << lenA) - 2;
                lenB = len; valB = w;
            }
        }
    }
    const int mine = chunks_of(lenA) + chunks_of(lenB);
    int incl = mine, total = n4;
    if (__any_sync(0xffffffffu, mine != 0)) {                             // no escape codes in either half (the common case): nothing to scan
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d, 16); if (hl >= d) incl += t; }
        total = n4 + __shfl_sync(0xffffffffu, incl, 15, 16);
    }

    if (!has) return;
    unsigned short *slot = lists + (size_t)(D.lid - LL::BASE) * LL::CAND + (size_t)(D.slot0 + cg) * SLOT;
    if (sig_coded) slot[__popc(sball & lt)] = ent_ctx(0, CX_SIG + sig_ctx_index(sz, st, y, x, nb), nzf);
    if (nzf && nzidx < 8) {
        const int c1 = earlier_big ? 0 : (nzidx < 2 ? 1 + nzidx : 3);
        slot[n1 + nzidx] = ent_ctx(1, CX_ONE + (first_cg ? 0 : 8) + c1, mag > 1);
    }
    if (firstbig) slot[n2] = ent_ctx(2, CX_ABS + (first_cg ? 0 : 2), mag > 2);
    if (nzf && nzidx == 0) {
        if (nz > 8) { slot[n3] = ent_byp((signs >> (nz - 8)) & 0xFF, 8); slot[n3 + 1] = ent_byp(signs & ((1 << (nz - 8)) - 1), nz - 8); }
        else slot[n3] = ent_byp(signs, nz);
    }
    if (mine) {
        int o = n4 + incl - mine;
        emit_bypass(slot, o, valA, lenA);
        emit_bypass(slot, o, valB, lenB);
    }
    if (hl == 0) lists[(size_t)(D.lid - LL::BASE) * LL::CAND + LL::HDR + D.slot0 + cg] = (unsigned short)(total | (nz > 0 ? 0x80 : 0) | (anybig ? 0x100 : 0));
}

// Binarises the transform units a warp has just quantised (descriptors in sm.bdesc[warp][0..G-1], levels in sm.lev).
// Tasks are the significant coefficient groups plus group 0 of every unit; each half-warp takes its share.
template <int TS, class LL = ListsGlobal>
HEVC_BIGFN void binarize_warp(Smem &sm, int warp, int lane, unsigned short *lists)
{
    constexpr int G = 32 / TS;
    TPROF_START(sm);
    const int half = lane >> 4, hl = lane & 15;
    int gi = G == 1 ? 0 : half * (G / 2);
    const int gend = G == 1 ? 1 : gi + G / 2;
    unsigned m0 = 0, m1 = 0;
    BinDesc D = sm.bdesc[warp][gi];
    auto load_masks = [&]() {
        m0 = 0; m1 = 0;
        if (D.active && (D.sg0 | D.sg1) != 0) {             // an all-zero unit has no bins: publish_tu wrote its empty group 0
            m0 = D.sg0 | 1u; m1 = D.sg1;
            if (G == 1) { const unsigned alt = half ? 0xAAAAAAAAu : 0x55555555u; m0 &= alt; m1 &= alt; }
        }
    };
    load_masks();
    for (;;) {
        while ((m0 | m1) == 0 && gi + 1 < gend) { gi++; D = sm.bdesc[warp][gi]; load_masks(); }
        const bool has = (m0 | m1) != 0;
        if (!__any_sync(0xffffffffu, has)) break;
        int cg = 0;
        if (has) {
            if (m0) { cg = __ffs((int)m0) - 1; m0 &= m0 - 1; }
            else    { cg = 32 + __ffs((int)m1) - 1; m1 &= m1 - 1; }
        }
        // position-order group mask of this unit (for the neighbour-group context)
        const int blk = D.slot0 / ((TS / 4) * (TS / 4));
        const int cand = D.lid;
        const unsigned cgp0 = has ? sm.ccg[D.set][cand][blk][0] : 0u, cgp1 = has ? sm.ccg[D.set][cand][blk][1] : 0u;
        binarize_cg<LL>(has, D, TS, cg, hl, half, sm.scan, &sm.lev[0][0], cgp0, cgp1, lists);
    }
    TPROF(sm, 64 + SizeClass<TS>::v * 8 + 7);
}

// Eight bin-list entries at a time: the lists live in (L2-backed) global memory, and a trial lane that fetched them one
// by one would pay a memory round trip per bin.  Group slots are 16-byte aligned (SLOT * 2 = 224 bytes).
struct alignas(16) Entry8 { unsigned w[4]; };
__device__ __forceinline__ unsigned entry_of(const Entry8 &v, int k)
{
    const int i = k & 7;
    const unsigned lo = (i & 2) ? v.w[1] : v.w[0], hi = (i & 2) ? v.w[3] : v.w[2];
    const unsigned w = (i & 4) ? hi : lo;
    return (i & 1) ? w >> 16 : w & 0xFFFFu;
}

// Arithmetic coding of one transform unit from its bin lists (one lane).  slots / hdrs point at the unit's group 0.
// sz == 0: no unit, only what waits in `pre` is coded (blocks without coefficients, the tail of a trial).
// ONE copy of this function serves every trial of the kernel (a real call, HEVC_PHASE): the trials are single lanes walking
// through code that nothing else keeps warm, and the kernel is bound by instruction fetch -- inlined
// at its seven call sites the coder loop alone was 80 KB of a coding unit's instruction stream.
HEVC_PHASE void code_stream(Coder &c_io, PreList &pre_io, const uint8_t (*scan_all)[2], int sz, int mode, int last,
                            unsigned cgp0, unsigned cgp1, unsigned sg0, unsigned sg1, const unsigned short *slots, const unsigned short *hdrs)
{
    Coder c = c_io;                                                     // registers inside; the caller's copy is in local memory across the call
    PreList pre = pre_io;
    // pointers that arrive through memory are generic addresses; contexts, tables and the prefix list live in this CTA's shared
    // memory: named through the __shared__ symbol again, their accesses are LDS / STS
    Smem &sm = *reinterpret_cast<Smem *>(IMCVT_DYN_SMEM());
    c.ctx = own_smem(c.ctx); c.lps = sm.lpsw; c.nlps = sm.next_lps;
    c.bind_shared();
    pre.p = own_smem(pre.p);
    scan_all = sm.scan;
    const int st = sz ? scan_type_for(sz, mode) : 0, ncg = sz >> 2;
    const uint8_t (*scan)[2] = scan_all + (sz ? scan_base(st, size_class(sz)) : 0);
    int cg = sz ? last >> 4 : 0, prevbig = 0, add = 0, k = 0, pk = 0;
    // the first group's header and entries are requested before the last position is binarised: an L2 round trip (another warp
    // wrote them) that the prefix work covers
    unsigned h = 0;
    const Entry8 *sp = reinterpret_cast<const Entry8 *>(slots + cg * SLOT);
    Entry8 ev; ev.w[0] = ev.w[1] = ev.w[2] = ev.w[3] = 0;
    if (sz) {
        // a unit without a significant group has an empty group 0 (publish_tu): nothing to fetch
        if (sg0 | sg1) { h = hdrs[cg]; ev = sp[0]; }
        build_last_pos(pre, sz, st, scan[last][0], scan[last][1]);
    }
    int cnt = (int)(h & 0x7Fu);
    for (;;) {
        int isb = 0, a = 0, b = 0;
        bool emit = false;
        if (pk < pre.n) {
            const unsigned e = pre.p[pk++];
            if ((e >> 14) == 3) { isb = 1; a = (int)(e & 0xFFu); b = (int)((e >> 8) & 15u); }
            else { a = (int)(e & 0xFFu); b = (int)((e >> 8) & 1u); }
            emit = true;
        } else if (k < cnt) {
            if (k > 0 && (k & 7) == 0) ev = sp[k >> 3];
            const unsigned e = entry_of(ev, k);
            k++;
            const int type = (int)(e >> 14);
            if (type == 3) { isb = 1; a = (int)(e & 0xFFu); b = (int)((e >> 8) & 15u); }
            else { a = (int)(e & 0xFFu) + (type == 1 ? 4 * add : type == 2 ? add : 0); b = (int)((e >> 8) & 1u); }
            emit = true;
        } else {
            if (h & 0x80u) prevbig = (int)((h >> 8) & 1u);
            if (--cg < 0) break;
            const int yc = scan[cg * 16][0] >> 2, xc = scan[cg * 16][1] >> 2;
            const int bpos = yc * 8 + xc, br = bpos + 1, bb = bpos + 8;
            const int cgs = (int)((cg < 32 ? sg0 >> cg : sg1 >> (cg - 32)) & 1u);
            const int right = xc < ncg - 1 && ((br < 32 ? cgp0 >> br : cgp1 >> (br - 32)) & 1u);
            const int below = yc < ncg - 1 && ((bb < 32 ? cgp0 >> bb : cgp1 >> (bb - 32)) & 1u);
            if (cg > 0) { a = CX_SIGCG + ((right | below) != 0); b = cgs; emit = true; }      // coded_sub_block_flag
            if (cgs || cg == 0) {
                h = hdrs[cg];
                sp = reinterpret_cast<const Entry8 *>(slots + cg * SLOT);
                ev = sp[0];
                cnt = (int)(h & 0x7Fu); add = prevbig;
            } else { h = 0; cnt = 0; }
            k = 0;
        }
        if (emit) { if (isb) c.bypass_chunk(a, b); else c.bin(b, a); }
    }
    c_io.range = c.range; c_io.low = c.low; c_io.nbits = c.nbits; c_io.pending = c.pending; c_io.held = c.held; c_io.zeros = c.zeros; c_io.nout = c.nout;
    pre_io.n = 0;
}
__device__ __forceinline__ void code_block(Coder &c, int sz, int mode, int last, unsigned cgp0, unsigned cgp1, unsigned sg0, unsigned sg1,
                                           const unsigned short *slots, const unsigned short *hdrs, const uint8_t (*scan_all)[2], PreList &pre)
{
    code_stream(c, pre, scan_all, sz, mode, last, cgp0, cgp1, sg0, sg1, slots, hdrs);
}
// steps the coder through what is left in a prefix list (blocks without coefficients do not call code_block)
__device__ __forceinline__ void emit_list(Coder &c, PreList &pre, const uint8_t (*scan_all)[2], const unsigned short *any_list)
{
    if (pre.n) code_stream(c, pre, scan_all, 0, 0, 0, 0u, 0u, 0u, 0u, any_list, any_list);
}

// coding-unit syntax before the residual (putCU_* :1271-1339 up to the cbf flags).  modes / left / above: one byte per
// prediction unit (kind 2 = NxN: four, else one).  All prev_intra_luma_pred flags first, then mpm_idx /
// rem_intra_luma_pred_mode (putYpmode :984-1017).  One compact copy (a real call, rolled loops): it runs once per trial on one
// lane, its size is its cost.
HEVC_PHASE void build_cu_head(PreList &L_io, int sz, int kind, unsigned modes, unsigned left, unsigned above)
{
    PreList L = L_io;
    L.p = own_smem(L.p);
    if (sz == 8) L.ctx(CX_PART, kind != 2);                                            // putPartSize :951-954
    const int count = kind == 2 ? 4 : 1;
    unsigned hits = 0;                                                                 // two bits per unit: 0 = not a probable mode, else index + 1
#pragma unroll 1
    for (int i = 0; i < count; i++) {
        int m[3];
        mpm_list((int)((left >> (8 * i)) & 255u), (int)((above >> (8 * i)) & 255u), m);
        const int md = (int)((modes >> (8 * i)) & 255u);
        const unsigned hit = m[2] == md ? 3u : m[1] == md ? 2u : m[0] == md ? 1u : 0u;  // the last match of :990-993
        hits |= hit << (2 * i);
        L.ctx(CX_YMODE, hit != 0);
    }
#pragma unroll 1
    for (int i = 0; i < count; i++) {
        const int hit = (int)((hits >> (2 * i)) & 3u);
        if (hit) {
            L.byp(hit > 1, 1);
            if (hit > 1) L.byp(hit - 2, 1);
        } else {
            int m[3], t, rem = (int)((modes >> (8 * i)) & 255u);
            mpm_list((int)((left >> (8 * i)) & 255u), (int)((above >> (8 * i)) & 255u), m);
            if (m[0] < m[1]) { t = m[0]; m[0] = m[1]; m[1] = t; }
            if (m[1] < m[2]) { t = m[1]; m[1] = m[2]; m[2] = t; }
            if (m[0] < m[1]) { t = m[0]; m[0] = m[1]; m[1] = t; }
            if (rem > m[0]) rem--;
            if (rem > m[1]) rem--;
            if (rem > m[2]) rem--;
            L.byp(rem, 5);
        }
    }
    L.ctx(CX_UVMODE, 0);                                                               // chroma mode "derived" :1020-1022
    if (kind != 2) L.ctx(CX_SPLIT_TU + (sz == 32 ? 0 : sz == 16 ? 1 : 2), kind == 1);  // :1025-1032
    L.ctx(CX_UVCBF + 0, 0);                                                            // cbf_cb, cbf_cr at depth 0
    L.ctx(CX_UVCBF + 0, 0);
    L_io.n = L.n;
}

// ------------------------------------------------------------------------------------------------
// 1-D transforms on register arrays.  out[k] = sum_n M[k][n] in[n] (forward) / out[n] = sum_k M[k][n] in[k]
// (inverse), M = N-point HEVC matrix.  Even/odd decomposition regroups the reference's dense sums (:469-492)
// before any rounding, so results are identical.
// ------------------------------------------------------------------------------------------------
template <int N> struct Dct {
    template <int OS> static __device__ __forceinline__ void fwd(const int *in, int *out)
    {
        int E[N / 2], O[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; k++) { E[k] = in[k] + in[N - 1 - k]; O[k] = in[k] - in[N - 1 - k]; }
#pragma unroll
        for (int m = 0; m < N / 2; m++) {
            int s = 0;
#pragma unroll
            for (int k = 0; k < N / 2; k++) s += c_dct[(2 * m + 1) * (32 / N)][k] * O[k];
            out[(2 * m + 1) * OS] = s;
        }
        Dct<N / 2>::template fwd<2 * OS>(E, out);
    }
    template <int IS> static __device__ __forceinline__ void inv(const int *in, int *out)
    {
        int Ee[N / 2];
        Dct<N / 2>::template inv<2 * IS>(in, Ee);
#pragma unroll
        for (int n = 0; n < N / 2; n++) {
            int s = 0;
#pragma unroll
            for (int m = 0; m < N / 2; m++) s += c_dct[(2 * m + 1) * (32 / N)][n] * in[(2 * m + 1) * IS];
            out[n] = Ee[n] + s;
            out[N - 1 - n] = Ee[n] - s;
        }
    }
};
template <> struct Dct<2> {
    template <int OS> static __device__ __forceinline__ void fwd(const int *in, int *out)
    {
        out[0] = 64 * (in[0] + in[1]);
        out[OS] = 64 * (in[0] - in[1]);
    }
    template <int IS> static __device__ __forceinline__ void inv(const int *in, int *out)
    {
        out[0] = 64 * (in[0] + in[IS]);
        out[1] = 64 * (in[0] - in[IS]);
    }
};

template <int TS> __device__ __forceinline__ void fwd1d(const int *in, int *out)
{
    if constexpr (TS == 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) out[k] = c_dst4[k][0] * in[0] + c_dst4[k][1] * in[1] + c_dst4[k][2] * in[2] + c_dst4[k][3] * in[3];
    } else {
        Dct<TS>::template fwd<1>(in, out);
    }
}
template <int TS> __device__ __forceinline__ void inv1d(const int *in, int *out)
{
    if constexpr (TS == 4) {
#pragma unroll
        for (int n = 0; n < 4; n++) out[n] = c_dst4[0][n] * in[0] + c_dst4[1][n] * in[1] + c_dst4[2][n] * in[2] + c_dst4[3][n] * in[3];
    } else {
        Dct<TS>::template inv<1>(in, out);
    }
}

// ------------------------------------------------------------------------------------------------
// phase A: one transform unit per group of TS lanes
// ------------------------------------------------------------------------------------------------
// where reference samples come from: the CTU window, or (inside the CU under trial) a private tile
struct Src {
    const uint8_t *rc;          // Smem::rc
    const uint8_t *priv;        // tile of the CU being tried (or nullptr)
    int py0, px0, psz, ppitch;
    __device__ __forceinline__ int at(int y, int x) const
    {
        if (priv && (unsigned)(y - py0) < (unsigned)psz && (unsigned)(x - px0) < (unsigned)psz)
            return priv[(y - py0) * ppitch + (x - px0)];
        return rc[(y + 1) * RCP + x + 1];
    }
};

template <int TS> __device__ __forceinline__ int group_sum(int v)
{
#pragma unroll
    for (int d = 1; d < TS; d <<= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
template <int TS> __device__ __forceinline__ int group_max(int v)
{
#pragma unroll
    for (int d = 1; d < TS; d <<= 1) { const int o = __shfl_xor_sync(0xffffffffu, v, d); v = o > v ? o : v; }
    return v;
}
template <int TS> __device__ __forceinline__ unsigned group_or(unsigned v)
{
#pragma unroll
    for (int d = 1; d < TS; d <<= 1) v |= __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

struct TuResult { int sse, last; unsigned cg0, cg1, sg0, sg1; };    // cg: groups by position (bit yc*8+xc), sg: by scan index

// Must be called by all 32 lanes of a warp (it contains warp-level exchanges); `active` masks the stores of idle
// groups.  (y0, x0): position of the TU inside the CTU.  avail: bit0 left, bit1 below-left, bit2 above, bit3 above-right.
// lev: TS x TS levels with pitch TS+2.  rec: TS x TS tile (pitch rec_pitch) that receives the reconstruction.
template <int TS>
HEVC_BIGFN TuResult tu_pipeline(Smem &sm, int warp, int lane, bool active, const Src &src, int y0, int x0,
                                                int avail, int mode, int qpd6, const RdK &rd, short *lev, uint8_t *rec, int rec_pitch)
{
    constexpr int LOG = TS == 4 ? 2 : TS == 8 ? 3 : TS == 16 ? 4 : 5;
    constexpr int SC = LOG - 2;
    constexpr int GS = (11 * TS + 6 + 3) & ~3;
    constexpr int LP = TS + 2;
    const int g = lane / TS, l = lane % TS;
    uint8_t *R = sm.refs[warp] + g * GS;
    uint8_t *left = R + 4, *top = left + 2 * TS, *fleft = top + 2 * TS, *ftop = fleft + 2 * TS, *line = ftop + 2 * TS;
    short *ws = sm.ws[warp] + g * TS * LP;
    const bool hl = avail & 1, hbl = avail & 2, ha = avail & 4, har = avail & 8;
    TPROF_START(sm);

    // ---- reference samples with substitution (getBorder :212-243) ----
    const int corner = (hl && ha) ? src.at(y0 - 1, x0 - 1) : hl ? src.at(y0, x0 - 1) : ha ? src.at(y0 - 1, x0) : 128;
#pragma unroll
    for (int i = l; i < 2 * TS; i += TS) {
        int vl, vt;
        if (i < TS) {
            vl = hl ? src.at(y0 + i, x0 - 1) : corner;
            vt = ha ? src.at(y0 - 1, x0 + i) : corner;
        } else {
            vl = hbl ? src.at(y0 + i, x0 - 1) : (hl ? src.at(y0 + TS - 1, x0 - 1) : corner);
            vt = har ? src.at(y0 - 1, x0 + i) : (ha ? src.at(y0 - 1, x0 + TS - 1) : corner);
        }
        left[i] = (uint8_t)vl; top[i] = (uint8_t)vt;
    }
    if (l == 0) R[0] = (uint8_t)corner;
    __syncwarp();
    // [1 2 1]/4 smoothing (:245-256); 4x4 units never use the smoothed samples (:274)
    if constexpr (TS > 4) {
#pragma unroll
        for (int i = l; i < 2 * TS; i += TS) {
            if (i == 2 * TS - 1) { fleft[i] = left[i]; ftop[i] = top[i]; }
            else {
                const int pl = i > 0 ? left[i - 1] : corner, pt = i > 0 ? top[i - 1] : corner;
                fleft[i] = (uint8_t)((2 + 2 * left[i] + pl + left[i + 1]) >> 2);
                ftop[i]  = (uint8_t)((2 + 2 * top[i] + pt + top[i + 1]) >> 2);
            }
        }
        if (l == 0) R[1] = (uint8_t)((2 + left[0] + top[0] + 2 * corner) >> 2);
    }
    __syncwarp();

    TPROF(sm, 64 + SC * 8 + 0);
    // ---- prediction (predict :262-381), this lane owns column l ----
    const bool smo = uses_smoothed(TS, mode);
    const uint8_t *Lr = smo ? fleft : left, *Tr = smo ? ftop : top;
    const int C = smo ? R[1] : R[0];
    const int dcsum = group_sum<TS>((int)left[l] + (int)top[l]);        // DC never uses the smoothed samples
    const bool angular = mode >= 2 && mode != 10 && mode != 26;
    uint8_t *lineb = line + TS;                                         // lineb[0] = corner, lineb[1..] = main reference, lineb[<0] = projected side
    if (angular) {
        const bool horiz = mode < 18;
        const int ang = c_angle[mode], inv = c_invang[mode];
        const uint8_t *mainr = horiz ? Lr : Tr, *side = horiz ? Tr : Lr;
        for (int i = l; i < 2 * TS + 2; i += TS) lineb[i] = (uint8_t)(i == 0 ? C : (i <= 2 * TS ? mainr[i - 1] : 0));
        const int lim = (TS * ang) >> 5;
        for (int i = -1 - l; i > lim; i -= TS) lineb[i] = side[((128 - inv * i) >> 8) - 1];
    }
    __syncwarp();

    int v[TS];          // residual column
    {
        int p[TS];
        if (mode == 0) {
            const int tl = Tr[l], tr = Tr[TS], bl = Lr[TS];
#pragma unroll
            for (int y = 0; y < TS; y++) p[y] = (TS + (TS - 1 - l) * Lr[y] + (l + 1) * tr + (TS - 1 - y) * tl + (y + 1) * bl) >> (LOG + 1);
        } else if (mode == 1) {
            const int dc = (dcsum + TS) >> (LOG + 1);
#pragma unroll
            for (int y = 0; y < TS; y++) p[y] = dc;
            if (TS <= 16) {
                if (l == 0) {
                    p[0] = (2 + 2 * dc + left[0] + top[0]) >> 2;
#pragma unroll
                    for (int y = 1; y < TS; y++) p[y] = (2 + 3 * dc + left[y]) >> 2;
                } else {
                    p[0] = (2 + 3 * dc + top[l]) >> 2;
                }
            }
        } else if (mode == 10) {
#pragma unroll
            for (int y = 0; y < TS; y++) p[y] = Lr[y];
            if (TS <= 16) p[0] = clampi(p[0] + ((Tr[l] - C) >> 1), 0, 255);
        } else if (mode == 26) {
#pragma unroll
            for (int y = 0; y < TS; y++) p[y] = Tr[l];
            if (TS <= 16 && l == 0) {
#pragma unroll
                for (int y = 0; y < TS; y++) p[y] = clampi(p[y] + ((Lr[y] - C) >> 1), 0, 255);
            }
        } else {
            const int ang = c_angle[mode];
            if (mode < 18) {
                const int off = ang * (l + 1), oi = off >> 5, of = off & 31;
#pragma unroll
                for (int y = 0; y < TS; y++) p[y] = ((32 - of) * lineb[oi + y + 1] + of * lineb[oi + y + 2] + 16) >> 5;
            } else {
#pragma unroll
                for (int y = 0; y < TS; y++) {
                    const int off = ang * (y + 1), oi = off >> 5, of = off & 31;
                    p[y] = ((32 - of) * lineb[oi + l + 1] + of * lineb[oi + l + 2] + 16) >> 5;
                }
            }
        }
#pragma unroll
        for (int y = 0; y < TS; y++) {
            if (active) rec[y * rec_pitch + l] = (uint8_t)p[y];
            v[y] = (int)sm.orig[y0 + y][x0 + l] - p[y];
        }
    }

    TPROF(sm, 64 + SC * 8 + 1);
    // ---- forward transform (:497-516, inverse = 0): columns, then rows ----
    int t[TS];
    fwd1d<TS>(v, t);
#pragma unroll
    for (int i = 0; i < TS; i++) ws[i * LP + l] = (short)((t[i] + (1 << (SC + 1) >> 1)) >> (SC + 1));
    __syncwarp();
#pragma unroll
    for (int k = 0; k < TS; k++) v[k] = ws[l * LP + k];
    fwd1d<TS>(v, t);
    __syncwarp();

    TPROF(sm, 64 + SC * 8 + 2);
    // ---- RDOQ-lite (quantize :540-594): this lane owns row l of the coefficient block ----
    bool warp_nz;                                       // some unit of this warp has a non-zero level
    {
        const int dist_sft = 8 - SC, sft = 19 - SC + qpd6;
        const int add = 1 << sft >> 1, max_d = I32MAX - add, thr = 9 << sft >> 2;
        int cgsum = 0;
#pragma unroll
        for (int n = 0; n < TS; n++) {
            const int coef = (t[n] + (1 << (SC + 8) >> 1)) >> (SC + 8);
            const int av = iabs(coef);
            int d = (av & 0x1ffff) << 14;
            d = (av > 0x1ffff || d > max_d) ? max_d : d;
            int level = clip16((d + add) >> sft);
            const int lo = level - 2 > 0 ? level - 2 : 0;
            int best = I32MAX, pick = 0;
            if (level == 0) level = -1;                     // a lone level-0 candidate can only confirm pick = 0
#pragma unroll 1
            for (; level >= lo; level--) {                  // a real loop: unrolled, this body was 40% of the kernel's code
                const int d1 = iabs(d - (level << sft)) >> dist_sft;
                const int dist = (d1 < 46340 ? d1 * d1 : I32MAX) >> 7;
                const int cost = rd.cost(dist, level_rate(level));
                if (cost < best) { best = cost; pick = level; }
            }
            t[n] = coef < 0 ? -pick : pick;
            cgsum += d < thr ? d : thr;
            if ((n & 3) == 3) {                         // coefficient group (rows l&~3.., columns n-3..n) complete
                cgsum += __shfl_xor_sync(0xffffffffu, cgsum, 1);
                cgsum += __shfl_xor_sync(0xffffffffu, cgsum, 2);
                if (cgsum < thr) { t[n] = 0; t[n - 1] = 0; t[n - 2] = 0; t[n - 3] = 0; }
                cgsum = 0;
            }
        }
        int nzrow = 0;
#pragma unroll
        for (int n = 0; n < TS; n++) nzrow |= t[n];
        warp_nz = __any_sync(0xffffffffu, nzrow != 0);
        if (warp_nz) {
            if (active) {
#pragma unroll
                for (int n = 0; n < TS; n++) lev[l * LP + n] = (short)t[n];
            } else {
#pragma unroll
                for (int n = 0; n < TS; n++) ws[l * LP + n] = (short)t[n];     // idle group: keep its data private
            }
        }
    }
    __syncwarp();
    const short *lv = active ? lev : ws;

    TPROF(sm, 64 + SC * 8 + 3);
    // ---- last significant position and coefficient-group mask (first pass of putCoef :1196-1203) ----
    TuResult res;
    res.last = 0; res.cg0 = 0; res.cg1 = 0; res.sg0 = 0; res.sg1 = 0;
    if (warp_nz) {
        const uint8_t (*scan)[2] = sm.scan + scan_base(scan_type_for(TS, mode), SC);
        int last = 0;
        unsigned cg0 = 0, cg1 = 0, sg0 = 0, sg1 = 0;
#pragma unroll 4
        for (int i = l; i < TS * TS; i += TS) {
            const int y = scan[i][0], x = scan[i][1];
            if (lv[y * LP + x] != 0) {
                last = i;
                const int b = (y >> 2) * 8 + (x >> 2), g4 = i >> 4;
                if (b < 32) cg0 |= 1u << b; else cg1 |= 1u << (b - 32);
                if (g4 < 32) sg0 |= 1u << g4; else sg1 |= 1u << (g4 - 32);
            }
        }
        res.last = group_max<TS>(last);
        res.cg0 = group_or<TS>(cg0);
        res.cg1 = TS == 32 ? group_or<TS>(cg1) : 0u;
        res.sg0 = group_or<TS>(sg0);
        res.sg1 = TS == 32 ? group_or<TS>(sg1) : 0u;
    }

    TPROF(sm, 64 + SC * 8 + 4);
    // ---- de-quantise (:599-614) + inverse transform: columns, then rows (all-zero levels reconstruct the prediction) ----
    if (warp_nz) {
        {
            const int s = 5 - SC + qpd6;
#pragma unroll
            for (int k = 0; k < TS; k++) v[k] = clip16((int)lv[k * LP + l] * (1 << s));
        }
        __syncwarp();                                   // idle groups keep their levels in ws: all reads done before it is reused
        inv1d<TS>(v, t);
#pragma unroll
        for (int i = 0; i < TS; i++) ws[i * LP + l] = (short)clip16((t[i] + 64) >> 7);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < TS; k++) v[k] = ws[l * LP + k];
        inv1d<TS>(v, t);
        __syncwarp();
    } else {
#pragma unroll
        for (int j = 0; j < TS; j++) t[j] = 0;
    }

    TPROF(sm, 64 + SC * 8 + 5);
    // ---- reconstruction of row l and its squared error ----
    int sse = 0;
    if (active) {
#pragma unroll
        for (int j = 0; j < TS; j++) {
            const int r = clampi(clip16((t[j] + 2048) >> 12) + (int)rec[l * rec_pitch + j], 0, 255);
            const int d = (int)sm.orig[y0 + l][x0 + j] - r;
            sse += d * d;
            rec[l * rec_pitch + j] = (uint8_t)r;
        }
    }
    res.sse = group_sum<TS>(sse);
    __syncwarp();
    TPROF(sm, 64 + SC * 8 + 6);
    return res;
}

// ------------------------------------------------------------------------------------------------
// block-level helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_sum(Smem &sm, int v)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = group_sum<32>(v);
    __syncthreads();
    if (lane == 0) sm.red[warp] = v;
    __syncthreads();
    int s = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) s += sm.red[w];
    return s;
}

__device__ __forceinline__ void copy_words(void *dst, const void *src, int nwords, int tid, int nthreads)
{
    unsigned *d = (unsigned *)dst;
    const unsigned *s = (const unsigned *)src;
    for (int i = tid; i < nwords; i += nthreads) d[i] = s[i];
}

// Trial encodes are one lane each, and the lanes of a warp that code different candidates sit in different places of
// the syntax, i.e. run one after the other.  So the trials are spread thin: a few lanes on every warp.
// 70 candidates of a unit over all 12 warps: 6 lanes each
__device__ __forceinline__ int candidate_of_thread(int tid, int ncand)
{
    const int lane = tid & 31, warp = tid >> 5;
    if (lane >= 6) return -1;
    const int c = warp * 6 + lane;
    return c < ncand ? c : -1;
}
// 35 trials over warps w0 .. w0+8: 4 lanes each
__device__ __forceinline__ int candidate35_of_thread(int tid, int w0)
{
    const int lane = tid & 31, warp = (tid >> 5) - w0;
    if (warp < 0 || warp >= 9 || lane >= 4) return -1;
    const int c = warp * 4 + lane;
    return c < NMODE ? c : -1;
}

struct Ctx {
    Smem *sm;
    const Params *P;
    uint8_t *out;               // this frame's output slot
    uint8_t *scratch[2];        // this CTA's [NALL][SCR_BYTES] per candidate set
    unsigned short *lists[2];   // this CTA's [NALL][CAND_LIST_U16] per candidate set
    const uint8_t *peer_scratch[2];     // two-CTA kernel: the partner's trial streams
    Smem *peer;                 // two-CTA kernel: the partner's shared memory
    int qpd6;
    RdK rd;
    long long prof_t;
    bool prof_on;               // development build: this CTA's thread 0 feeds the phase counters (the spine of a pair)
};

template <int SZ> struct Level { static constexpr int v = SZ == 32 ? 0 : SZ == 16 ? 1 : 2; static constexpr int set = SZ == 32 ? 0 : 1; };
template <int SET> __device__ __forceinline__ uint8_t *rec_tile(Smem &sm, int c) { return SET == 0 ? sm.rec0[c] : sm.rec1[c]; }

__device__ __forceinline__ void attach(Coder &c, Smem &sm, uint8_t *ctx, uint8_t *out, unsigned cap)
{
    c.ctx = ctx; c.out = out; c.cap = cap; c.lps = sm.lpsw; c.nlps = sm.next_lps;
    c.bind_shared();
}

template <int SZ>
__device__ __forceinline__ void fill_maps(Smem &sm, int y0, int x0, int mode)
{
    constexpr int NC = SZ / 4;
    const int tid = threadIdx.x;
    if (tid < NC * NC) {
        const int r = 1 + y0 / 4 + tid / NC, c = 1 + x0 / 4 + tid % NC;
        sm.msz[r][c] = (uint8_t)SZ;
        sm.mmd[r][c] = (uint8_t)mode;
    }
}

// group leaders publish what binarize_warp needs about the unit they just quantised
template <int SET, class LL = ListsGlobal>
__device__ __forceinline__ void publish_tu(Smem &sm, unsigned short *lists, int warp, int g, bool active, int cand, int blk, int ncg, int mode, int lev_off, const TuResult &r)
{
    BinDesc &D = sm.bdesc[warp][g];
    D.set = SET;
    D.active = active; D.lid = cand; D.slot0 = blk * ncg; D.mode = mode; D.lev_off = lev_off; D.last = r.last; D.sg0 = r.sg0; D.sg1 = r.sg1;
    if (active) {
        if ((r.sg0 | r.sg1) == 0) lists[(size_t)(cand - LL::BASE) * LL::CAND + LL::HDR + blk * ncg] = 0;      // empty group 0 (stand-alone 4x4 trials read it)
        sm.clast[SET][cand][blk] = r.last;
        sm.ccg[SET][cand][blk][0] = r.cg0; sm.ccg[SET][cand][blk][1] = r.cg1;
        sm.csg[SET][cand][blk][0] = r.sg0; sm.csg[SET][cand][blk][1] = r.sg1;
    }
}

// Best of candidates first..first+n-1 in the reference order (`best >= cost` while walking up: a later candidate wins
// ties, :1439).  One full warp; every lane returns the result (win = -1 when n == 0).
template <int SET>
__device__ __forceinline__ void warp_best(Smem &sm, const RdK &rd, int first, int n, int lane, int &bc, int &win)
{
    bc = I32MAX; win = -1;
    for (int c = first + lane; c < first + n; c += 32) {
        const int v = rd.cost(sm.csse[SET][c], sm.cbits[SET][c]);
        if (bc >= v) { bc = v; win = c; }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, bc, d), ow = __shfl_xor_sync(0xffffffffu, win, d);
        if (ow >= 0 && (win < 0 || ob < bc || (ob == bc && ow > win))) { bc = ob; win = ow; }
    }
}

// CABAC trial encode of candidate c of a coding unit from the unit's entry state (:1422-1437, :1455-1473): c < 35 one
// transform unit, c >= 35 four.  One lane.
template <int SZ>
HEVC_PHASE void trial_cu(Ctx &K, int c, const CoderRegs &entry, int entry_bits_rel, int split_ctx, int mode_l, int mode_a)
{
    Smem &sm = HEVC_SMEM(K);
    constexpr int H = SZ / 2, SET = Level<SZ>::set;
    const int kind = c >= NMODE, mode = c - kind * NMODE;
    Coder cd; cd.load(entry); cd.nout = 0;
    attach(cd, sm, sm.cctx[SET][c], K.scratch[SET] + (size_t)c * SCR_BYTES, SCR_BYTES);
    PreList pre; pre.p = sm.tpre[c]; pre.n = 0;
    if (SZ >= 16) pre.ctx(CX_SPLIT_CU + split_ctx, 0);               // putSplitCUflag :942-946
    build_cu_head(pre, SZ, kind, (unsigned)mode, (unsigned)mode_l, (unsigned)mode_a);
    const unsigned short *lst = K.lists[SET] + (size_t)c * CAND_LIST_U16;
    const int tsz = kind ? H : SZ, ncg = (tsz / 4) * (tsz / 4);
    for (int i = 0; i < (kind ? 4 : 1); i++) {
        const unsigned g0 = sm.ccg[SET][c][i][0], g1 = sm.ccg[SET][c][i][1];
        const int cbf = (g0 | g1) != 0;
        pre.ctx(CX_YCBF + (kind ? 0 : 1), cbf);                      // :1039
        if (cbf) code_block(cd, tsz, mode, sm.clast[SET][c][i], g0, g1, sm.csg[SET][c][i][0], sm.csg[SET][c][i][1],
                            lst + (size_t)i * ncg * SLOT, lst + 64 * SLOT + i * ncg, sm.scan, pre);
    }
    emit_list(cd, pre, sm.scan, lst);
    cd.store(sm.ccoder[SET][c]);
    sm.cbits[SET][c] = cd.bits() - entry_bits_rel;
}


// 8x8 units: CABAC trial of four-TU candidate NMODE + c (:1455-1473) block by block, as the blocks arrive -- block i of every
// candidate is published by the warps that run its pipeline with bar.arrive on barrier bar0 + i; this trial waits there, so only
// the last block's bins are left to code once the last pipelines are through.  Called by WHOLE warps (`nthreads` = the waiting
// threads plus the arriving ones); lanes without a candidate (c < 0) only keep the barriers company.
HEVC_PHASE void trial_cu8_four(Ctx &K, int c, const CoderRegs &entry, int entry_bits_rel, int mode_l, int mode_a, int bar0, int nthreads)
{
    Smem &sm = HEVC_SMEM(K);
    constexpr int SET = 1;
    const bool on = c >= 0;
    const int mode = on ? c : 0, cand = NMODE + mode;
    Coder cd; cd.load(entry); cd.nout = 0;
    attach(cd, sm, sm.cctx[SET][cand], K.scratch[SET] + (size_t)cand * SCR_BYTES, SCR_BYTES);
    PreList pre; pre.p = sm.tpre[cand]; pre.n = 0;
    if (on) build_cu_head(pre, 8, 1, (unsigned)mode, (unsigned)mode_l, (unsigned)mode_a);
    const unsigned short *lst = cu8_lists(sm) + (size_t)(cand - ListsCu8::BASE) * ListsCu8::CAND;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        __syncwarp();
        HEVC_NAMED_BARRIER(bar0 + i, nthreads);
        if (on) {
            const unsigned g0 = sm.ccg[SET][cand][i][0];
            const int cbf = g0 != 0;
            pre.ctx(CX_YCBF + 0, cbf);                                   // :1039
            if (cbf) code_block(cd, 4, mode, sm.clast[SET][cand][i], g0, 0u, sm.csg[SET][cand][i][0], 0u,
                                lst + (size_t)i * SLOT, lst + ListsCu8::HDR + i, sm.scan, pre);
        }
    }
    if (on) {
        emit_list(cd, pre, sm.scan, lst);
        cd.store(sm.ccoder[SET][cand]);
        sm.cbits[SET][cand] = cd.bits() - entry_bits_rel;
    }
}


// ------------------------------------------------------------------------------------------------
// 32x32 and 16x16 coding units (processCURecurs :1349-1559).  The work of a unit comes in pieces that the whole CTA
// executes in uniform control flow:
//   cu_enter      entry state (:1363-1372): coder, contexts, neighbour modes -> sm.entry / entry_ctx / par [level]
//   eval_a1/a4    phase A of the 35 one-TU candidates (:1419-1447) / of the 35 four-TU candidates (:1454-1483)
//   eval_b        the 70 CABAC trials from the entry state and their best (reference order, a later one wins ties)
//   cu_split_cost cost of the split alternative once the four smaller units are coded (:1402-1412)
// The evaluation reads only samples OUTSIDE the unit and the entry state, never the smaller units' results, which is
// what lets the two-CTA kernel run it on the partner CTA while the smaller units are searched.
// ------------------------------------------------------------------------------------------------
template <int SZ> struct SubUnits {
    int y[4], x[4], avail[4];
    __device__ __forceinline__ SubUnits(int y0, int x0, int av)
    {
        constexpr int H = SZ / 2;
        const bool hl = av & 1, hbl = av & 2, ha = av & 4, har = av & 8;
        y[0] = y0; y[1] = y0; y[2] = y0 + H; y[3] = y0 + H;
        x[0] = x0; x[1] = x0 + H; x[2] = x0; x[3] = x0 + H;
        avail[0] = (hl ? 1 : 0) | (hl ? 2 : 0) | (ha ? 4 : 0) | (ha ? 8 : 0);          // :1375-1378
        avail[1] = 1 | 0 | (ha ? 4 : 0) | (har ? 8 : 0);
        avail[2] = (hl ? 1 : 0) | (hbl ? 2 : 0) | 4 | 8;
        avail[3] = 1 | 0 | 4 | 0;
    }
};

template <int SZ>
__device__ __forceinline__ void cu_enter(Ctx &K, int y0, int x0, int avail)
{
    Smem &sm = HEVC_SMEM(K);
    constexpr int LV = Level<SZ>::v;
    const int tid = threadIdx.x, cy = y0 / 4, cx = x0 / 4;
    __syncthreads();
    copy_words(sm.entry_ctx[LV], sm.cur_ctx, CXP / 4, tid, NT);
    if (tid == 0) {
        sm.entry[LV] = sm.cur; sm.best_cost[LV] = I32MAX;
        CuPar q; q.y0 = y0; q.x0 = x0; q.avail = avail;
        q.split_ctx = (SZ > sm.msz[1 + cy][cx]) + (SZ > sm.msz[cy][1 + cx]);
        q.mode_l = sm.mmd[1 + cy][cx]; q.mode_a = sm.mmd[cy][1 + cx];
        sm.par[LV] = q;
    }
    __syncthreads();
}

// the split flag of a unit that is being split (:1403); one thread, after cu_enter
template <int SZ>
__device__ __forceinline__ void put_split_flag(Ctx &K)
{
    Smem &sm = HEVC_SMEM(K);
    if (threadIdx.x == 0) {
        Coder c; c.load(sm.cur); attach(c, sm, sm.cur_ctx, K.out, (unsigned)K.P->out_stride);
        c.bin(1, CX_SPLIT_CU + sm.par[Level<SZ>::v].split_ctx);
        c.store(sm.cur);
    }
}

template <int SZ>
__device__ __forceinline__ void cu_split_cost(Ctx &K)
{
    Smem &sm = HEVC_SMEM(K);
    constexpr int LV = Level<SZ>::v;
    const int tid = threadIdx.x, y0 = sm.par[LV].y0, x0 = sm.par[LV].x0;
    __syncthreads();
    int sse = 0;
    for (int i = tid; i < SZ * SZ; i += NT) {
        const int y = i / SZ, x = i % SZ;
        const int d = (int)sm.orig[y0 + y][x0 + x] - (int)sm.rc[1 + y0 + y][1 + x0 + x];
        sse += d * d;
    }
    sse = block_sum(sm, sse);
    if (tid == 0) {
        const CoderRegs &e = sm.entry[LV];
        const int bits = 8 * ((int)sm.cur.nout + sm.cur.pending) + 23 - sm.cur.nbits;
        sm.best_cost[LV] = K.rd.cost(sse, bits - (8 * ((int)e.nout + e.pending) + 23 - e.nbits));
    }
    __syncthreads();
}

template <int SZ>
HEVC_PHASE void eval_a1(Ctx &K)
{
    Smem &sm = HEVC_SMEM(K);
    constexpr int LV = Level<SZ>::v, SET = Level<SZ>::set;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CuPar q = sm.par[LV];
    Src src; src.rc = &sm.rc[0][0]; src.priv = nullptr; src.py0 = q.y0; src.px0 = q.x0; src.psz = SZ; src.ppitch = SZ;
    constexpr int G = 32 / SZ, NG = NW * G;
    for (int base = 0; base < NMODE; base += NG) {
        if (base + warp * G >= NMODE) break;                                // warp-uniform
        const int item = base + warp * G + lane / SZ;
        const bool active = item < NMODE;
        const int m = active ? item : 0;
        const int loff = (lane / SZ) * SZ * (SZ + 2);
        TuResult r = tu_pipeline<SZ>(sm, warp, lane, active, src, q.y0, q.x0, q.avail, m, K.qpd6, K.rd, sm.lev[warp] + loff, rec_tile<SET>(sm, m), SZ);
        if (lane % SZ == 0) {
            publish_tu<SET>(sm, K.lists[SET], warp, lane / SZ, active, m, 0, (SZ / 4) * (SZ / 4), m, warp * LEV_STRIDE + loff, r);
            if (active) sm.csse[SET][m] = r.sse;
        }
        __syncwarp();
        binarize_warp<SZ>(sm, warp, lane, K.lists[SET]);
        __syncwarp();
    }
    __syncthreads();
}

template <int SZ>
HEVC_PHASE void eval_a4(Ctx &K)
{
    Smem &sm = HEVC_SMEM(K);
    constexpr int LV = Level<SZ>::v, SET = Level<SZ>::set, H = SZ / 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CuPar q = sm.par[LV];
    const SubUnits<SZ> sub(q.y0, q.x0, q.avail);
    Src src; src.rc = &sm.rc[0][0]; src.priv = nullptr; src.py0 = q.y0; src.px0 = q.x0; src.psz = SZ; src.ppitch = SZ;
    constexpr int G = 32 / H, NG = NW * G;
    for (int base = 0; base < NMODE; base += NG) {
        if (base + warp * G >= NMODE) break;
        const int item = base + warp * G + lane / H;
        const bool active = item < NMODE;
        const int m = active ? item : 0, cand = NMODE + m;
        // earlier sub-TUs of the same trial.  An idle group (item >= NMODE) runs the pipeline with its stores masked; it must not
        // read candidate 0's tile while that candidate's own group writes it (values unused, but a shared-memory race)
        Src ps = src; ps.priv = active ? rec_tile<SET>(sm, cand) : nullptr;
        const int loff = (lane / H) * H * (H + 2);
        int sse = 0;
#pragma unroll 1
        for (int i = 0; i < 4; i++) {                                        // one copy of the pipeline in the instruction cache, not four
            TuResult r = tu_pipeline<H>(sm, warp, lane, active, ps, sub.y[i], sub.x[i], sub.avail[i], m, K.qpd6, K.rd,
                                        sm.lev[warp] + loff, rec_tile<SET>(sm, cand) + (sub.y[i] - q.y0) * SZ + (sub.x[i] - q.x0), SZ);
            sse += r.sse;
            if (lane % H == 0) publish_tu<SET>(sm, K.lists[SET], warp, lane / H, active, cand, i, (H / 4) * (H / 4), m, warp * LEV_STRIDE + loff, r);
            __syncwarp();
            binarize_warp<H>(sm, warp, lane, K.lists[SET]);
            __syncwarp();
        }
        if (active && lane % H == 0) sm.csse[SET][cand] = sse;
    }
    __syncthreads();
}

template <int SZ>
HEVC_PHASE void eval_b(Ctx &K)
{
    Smem &sm = HEVC_SMEM(K);
    constexpr int LV = Level<SZ>::v, SET = Level<SZ>::set;
    const int tid = threadIdx.x;
    for (int i = tid; i < NCAND * (CXP / 4); i += NT)
        ((unsigned *)sm.cctx[SET])[i] = ((const unsigned *)sm.entry_ctx[LV])[i % (CXP / 4)];
    __syncthreads();
    {
        const int c = candidate_of_thread(tid, NCAND);
        if (c >= 0) {
            const CoderRegs entry = sm.entry[LV];
            const CuPar q = sm.par[LV];
            trial_cu<SZ>(K, c, entry, 8 * entry.pending + 23 - entry.nbits, q.split_ctx, q.mode_l, q.mode_a);
        }
    }
    __syncthreads();
    if (tid < 32) {
        int bc, win;
        warp_best<SET>(sm, K.rd, 0, NCAND, tid, bc, win);
        if (tid == 0) { sm.eval_cost[LV] = bc; sm.eval_win[LV] = win; }
    }
    __syncthreads();
}

// decision between the split alternative (cost in best_cost) and the evaluated candidates (:1437-1446, :1473-1482): the
// candidates come later in the reference order, so they win ties.  When a candidate wins, it becomes the current state.
//   REMOTE = false: the candidates were evaluated by this CTA;  true: by the partner CTA (two-CTA kernel), which left the
//   winner's coder, contexts and reconstruction in res_* and its bytes in its own trial stream.
template <int SZ, bool REMOTE>
__device__ __forceinline__ void cu_decide_commit(Ctx &K)
{
    Smem &sm = HEVC_SMEM(K);
    constexpr int LV = Level<SZ>::v, SET = Level<SZ>::set;
    const int tid = threadIdx.x, y0 = sm.par[LV].y0, x0 = sm.par[LV].x0;
    const int win = sm.eval_win[LV];
    if (sm.best_cost[LV] >= sm.eval_cost[LV] && win >= 0) {                  // uniform: values were written before a barrier
        const CoderRegs &wc = REMOTE ? sm.res_coder[LV] : sm.ccoder[SET][win];
        const unsigned base = sm.entry[LV].nout, n = wc.nout;
        const uint8_t *bytes = (REMOTE ? K.peer_scratch[SET] : K.scratch[SET]) + (size_t)win * SCR_BYTES;
        const unsigned cap = (unsigned)K.P->out_stride;
        for (unsigned i = tid; i < n && i < SCR_BYTES; i += NT) if (base + i < cap) K.out[base + i] = bytes[i];
        copy_words(sm.cur_ctx, REMOTE ? sm.res_ctx[LV] : sm.cctx[SET][win], CXP / 4, tid, NT);
        const uint8_t *rec = REMOTE ? (SET == 0 ? sm.res_rec0 : sm.res_rec1) : rec_tile<SET>(sm, win);
        for (int i = tid; i < SZ * SZ; i += NT) sm.rc[1 + y0 + i / SZ][1 + x0 + i % SZ] = rec[i];
        fill_maps<SZ>(sm, y0, x0, win >= NMODE ? win - NMODE : win);
        __syncthreads();                                                    // every thread has read wc.nout
        if (tid == 0) { CoderRegs c = wc; c.nout = base + n; sm.cur = c; sm.best_cost[LV] = sm.eval_cost[LV]; }
    }
    __syncthreads();
}

HEVC_PHASE void search_cu8_groups(Ctx &K, int y0, int x0, int avail);
HEVC_PHASE void search_cu8_lockstep(Ctx &K, int y0, int x0, int avail);
__device__ __forceinline__ void search_cu8(Ctx &K, int y0, int x0, int avail)
{
    if (K.P->cu8_lockstep) search_cu8_lockstep(K, y0, x0, avail); else search_cu8_groups(K, y0, x0, avail);
}

// one CTA: a unit is searched depth first, the evaluation of its own candidates after its four smaller units
template <int SZ>
__device__ void search_cu(Ctx &K, int y0, int x0, int avail)
{
    static_assert(SZ == 32 || SZ == 16, "8x8 coding units: search_cu8");
    constexpr int LV = Level<SZ>::v, H = SZ / 2;
    cu_enter<SZ>(K, y0, x0, avail);
    HPROF(K, LV * 16 + 0);
    put_split_flag<SZ>(K);
    const SubUnits<SZ> sub(y0, x0, avail);
    for (int i = 0; i < 4; i++) {
        if constexpr (H == 8) search_cu8(K, sub.y[i], sub.x[i], sub.avail[i]);
        else search_cu<H>(K, sub.y[i], sub.x[i], sub.avail[i]);
    }
    cu_split_cost<SZ>(K);
    HPROF(K, LV * 16 + 1);
    eval_a1<SZ>(K);
    HPROF(K, LV * 16 + 2);
    eval_a4<SZ>(K);
    HPROF(K, LV * 16 + 3);
    eval_b<SZ>(K);
    HPROF(K, LV * 16 + 5);
    cu_decide_commit<SZ, false>(K);
    HPROF(K, LV * 16 + 7);
}

// RD search of an 8x8 coding unit (processCURecurs :1349-1559 with sz == 8: no split; one 8x8 TU, four 4x4 TUs, or
// four 4x4 PUs).  The reference's steps 2, 3 and 4 all start from the unit's entry state and only step 4 is a chain (a PU
// predicts from the kept reconstruction of the PUs before it), so three groups of warps run them side by side (see the
// body); the decision follows the reference order: one-TU modes, four-TU modes, NxN, a later candidate wins ties.
HEVC_PHASE void search_cu8_groups(Ctx &K, int y0, int x0, int avail)
{
    Smem &sm = HEVC_SMEM(K);
    constexpr int SZ = 8, LV = 2, H = 4, SET = 1, RT = 320;                   // RT: threads of the round warps
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qpd6 = K.qpd6;
    const RdK rd = K.rd;
    const int cy = y0 / 4, cx = x0 / 4;
    const bool hl = avail & 1, hbl = avail & 2, ha = avail & 4, har = avail & 8;
    const unsigned out_cap = (unsigned)K.P->out_stride;

    __syncthreads();
    // ---- entry state (:1363-1372) ----
    copy_words(sm.entry_ctx[LV], sm.cur_ctx, CXP / 4, tid, NT);
    if (tid == 0) sm.entry[LV] = sm.cur;
    const int mode_l = sm.mmd[1 + cy][cx], mode_a = sm.mmd[cy][1 + cx];
    // candidate contexts: the trials of the whole unit start from the entry state, the stand-alone 4x4 trials are fresh
    for (int i = tid; i < NALL * (CXP / 4); i += NT) {
        const int c = i / (CXP / 4), w = i % (CXP / 4);
        ((unsigned *)sm.cctx[SET][c])[w] = (c >= NCAND && c < NXN) ? ((const unsigned *)sm.init_ctx)[w] : ((const unsigned *)sm.cur_ctx)[w];
    }
    __syncthreads();
    const CoderRegs entry = sm.entry[LV];
    const int entry_bits_rel = 8 * entry.pending + 23 - entry.nbits;
    HPROF(K, LV * 16 + 0);

    const int sub_y[4] = {y0, y0, y0 + H, y0 + H}, sub_x[4] = {x0, x0 + H, x0, x0 + H};
    const int sub_avail[4] = {                                                        // :1375-1378
        (hl ? 1 : 0) | (hl ? 2 : 0) | (ha ? 4 : 0) | (ha ? 8 : 0),
        1 | 0 | (ha ? 4 : 0) | (har ? 8 : 0),
        (hl ? 1 : 0) | (hbl ? 2 : 0) | 4 | 8,
        1 | 0 | 4 | 0 };
    Src src; src.rc = &sm.rc[0][0]; src.priv = nullptr; src.py0 = y0; src.px0 = x0; src.psz = SZ; src.ppitch = SZ;

    // Three groups of warps work side by side from the entry state and only meet where one needs another's results:
    //   A  warps 0-4    the four blocks of the 35 four-TU candidates back to back (a candidate's block i + 1 predicts from its own
    //                   block i, nothing else), then their CABAC trials
    //   B  warps 5-9    one-TU pipelines (first pass), then the four rounds of stand-alone PUs: pipelines, trials on fresh coders,
    //                   choice, kept reconstruction -- the unit's critical chain -- then the NxN trial
    //   C  warps 10-11  one-TU pipelines (first pass and the rest), then the one-TU trials
    // Barriers: 7 = B + C after the first one-TU pass (the trials read every candidate's lists), 5 = A + B once block 0 of the
    // four-TU candidates is in (round 0 of the stand-alone PUs codes those same transform units), 6 = A alone, 4 = B alone.
    constexpr int GA = 5, GB = 5;                                           // warps of group A / B
    constexpr int BT = GB * 32;                                             // threads of group B
    // Group B is the unit's critical chain and the issue arbiter of an SM sub-partition favours the highest warp ids: B sits on
    // hardware warps 7-11, C on 5-6, A on 0-4.  `lw` is the warp's number in the order A, B, C; the per-warp buffers stay
    // with the hardware warp.
    const int lw = warp < GA ? warp : (warp < NW - GB ? warp + GB : warp - (NW - GA - GB));
    if (lw < GA) {
        // ---- group A: four-TU candidates 35..69 (:1454-1483) ----
        const int item = warp * 8 + lane / H;
        const bool active = item < NMODE;
        const int m = active ? item : 0, cand = NMODE + m;
        const int loff = (lane / H) * H * (H + 2);
        int sse3 = 0;
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            Src ps = src;
            ps.priv = active ? sm.rec1[cand] : nullptr;                     // earlier blocks of the same trial (idle groups: none)
            ps.ppitch = 8;
            TuResult r = tu_pipeline<H>(sm, warp, lane, active, ps, sub_y[i], sub_x[i], sub_avail[i], m, qpd6, rd,
                                        sm.lev[warp] + loff, sm.rec1[cand] + (sub_y[i] - y0) * SZ + (sub_x[i] - x0), SZ);
            sse3 += r.sse;
            if (lane % H == 0) {
                publish_tu<SET>(sm, K.lists[SET], warp, lane / H, active, cand, i, 1, m, warp * LEV_STRIDE + loff, r);
                if (active) {
                    if (i == 0) sm.sse0[m] = r.sse;
                    else if (i == 3) sm.csse[SET][cand] = sse3;
                }
            }
            __syncwarp();
            binarize_warp<H>(sm, warp, lane, K.lists[SET]);
            __syncwarp();
            if (i == 0) HEVC_NAMED_BARRIER(5, (GA + GB) * 32);              // block 0 is in: group B's round 0 may code it
        }
        HEVC_NAMED_BARRIER(6, GA * 32);                                     // every candidate's four blocks are in
        HPROF(K, LV * 16 + 3);
        if (lane < 7) trial_cu<SZ>(K, NMODE + warp * 7 + lane, entry, entry_bits_rel, 0, mode_l, mode_a);
        HPROF(K, LV * 16 + 5);
    } else {
        // ---- groups B and C: one-TU candidates 0..34 (:1419-1447), four transform units per warp and pass ----
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1 && lw < GA + GB) break;                         // the second pass is group C's
            const int item = (pass == 0 ? (lw - GA) * 4 : 28 + (lw - GA - GB) * 4) + lane / SZ;
            const bool active = item < NMODE;
            const int m = active ? item : 0;
            const int loff = (lane / SZ) * SZ * (SZ + 2);
            TuResult r = tu_pipeline<SZ>(sm, warp, lane, active, src, y0, x0, avail, m, qpd6, rd, sm.lev[warp] + loff, sm.rec1[m], SZ);
            if (lane % SZ == 0) {
                publish_tu<SET>(sm, K.lists[SET], warp, lane / SZ, active, m, 0, 4, m, warp * LEV_STRIDE + loff, r);
                if (active) sm.csse[SET][m] = r.sse;
            }
            __syncwarp();
            binarize_warp<SZ>(sm, warp, lane, K.lists[SET]);
            __syncwarp();
            if (pass == 0) HEVC_NAMED_BARRIER(7, NT - GA * 32);             // candidates 0..27 are in
        }
        if (lw >= GA + GB) {
            // ---- group C: CABAC trials of the one-TU candidates ----
            __syncwarp();
            HEVC_NAMED_BARRIER(8, NT - (GA + GB) * 32);                     // candidates 28..34 are in (the two warps of the group)
            const int c = (lw - GA - GB) * 18 + lane;
            if (lane < 18 && c < NMODE) trial_cu<SZ>(K, c, entry, entry_bits_rel, 0, mode_l, mode_a);
        } else {
            // ---- group B: four rounds over the 4x4 blocks, stand-alone PUs (:1498-1527) ----
            const int bw = lw - GA, btid = bw * 32 + lane;
            const int item = bw * 8 + lane / H;
            const bool active = item < NMODE;
            const int m = active ? item : 0;
            const int loff = (lane / H) * H * (H + 2);
            HEVC_NAMED_BARRIER(5, (GA + GB) * 32);                          // block 0 of the four-TU candidates: the same transform units as round 0's
            for (int i = 0; i < 4; i++) {
                if (i > 0) {
                    const int cand = NCAND + m;
                    Src ps = src;
                    ps.priv = active ? sm.nxn_rec : nullptr;                // the kept blocks of the earlier PUs (idle groups: none)
                    ps.ppitch = 8;
                    TuResult r = tu_pipeline<H>(sm, warp, lane, active, ps, sub_y[i], sub_x[i], sub_avail[i], m, qpd6, rd,
                                                sm.lev[warp] + loff, sm.rec4[m], H);
                    if (lane % H == 0) {
                        publish_tu<SET>(sm, K.lists[SET], warp, lane / H, active, cand, 0, 1, m, warp * LEV_STRIDE + loff, r);
                        if (active) sm.csse[SET][cand] = r.sse;
                    }
                    __syncwarp();
                    binarize_warp<H>(sm, warp, lane, K.lists[SET]);
                    __syncwarp();
                    HEVC_NAMED_BARRIER(4, BT);
                }
                // stand-alone trial of every mode on a fresh coder: cost = absolute length of the block's residual (:1504-1518)
                if (lane < 7) {
                    const int c = bw * 7 + lane;
                    const int from = i == 0 ? NMODE + c : NCAND + c;
                    Coder cd; cd.fresh();
                    attach(cd, sm, sm.cctx[SET][NCAND + c], K.scratch[SET] + (size_t)(NCAND + c) * SCR_BYTES, SCR_BYTES);
                    const unsigned short *lst = K.lists[SET] + (size_t)from * CAND_LIST_U16;
                    PreList pre; pre.p = sm.tpre[NCAND + c]; pre.n = 0;
                    code_block(cd, 4, c, sm.clast[SET][from][0], sm.ccg[SET][from][0][0], 0u, sm.csg[SET][from][0][0], 0u, lst, lst + 64 * SLOT, sm.scan, pre);
                    sm.cbits[SET][NCAND + c] = cd.bits();
                    if (i == 0) sm.csse[SET][NCAND + c] = sm.sse0[c];
                }
                HEVC_NAMED_BARRIER(4, BT);
                if (bw == 0) {
                    int bc, win;
                    warp_best<SET>(sm, rd, NCAND, NMODE, lane, bc, win);
                    win -= NCAND;
                    if (lane == 0) {
                        const int from = i == 0 ? NMODE + win : NCAND + win;
                        sm.win = win;
                        sm.nxn_mode[i] = win; sm.nxn_last[i] = sm.clast[SET][from][0]; sm.nxn_cg[i] = sm.ccg[SET][from][0][0];
                    }
                }
                // fresh contexts for the next round's trials
                for (int k = btid; k < NMODE * (CXP / 4); k += BT)
                    ((unsigned *)sm.cctx[SET][NCAND])[k] = ((const unsigned *)sm.init_ctx)[k % (CXP / 4)];
                HEVC_NAMED_BARRIER(4, BT);
                {
                    const int w = sm.win;
                    const int from = i == 0 ? NMODE + w : NCAND + w;
                    const unsigned short *lst = K.lists[SET] + (size_t)from * CAND_LIST_U16;
                    if (btid < SLOT) sm.nxn_list[i][btid] = lst[btid];         // keep the winner's bins for the NxN trial
                    if (btid == SLOT) sm.nxn_hdr[i] = lst[64 * SLOT];
                    if (btid >= 128 && btid < 144) {
                        const int y = (btid - 128) >> 2, x = (btid - 128) & 3;
                        sm.nxn_rec[(sub_y[i] - y0 + y) * 8 + (sub_x[i] - x0 + x)] = i == 0 ? sm.rec1[NMODE + w][y * 8 + x] : sm.rec4[w][y * 4 + x];
                    }
                }
                HEVC_NAMED_BARRIER(4, BT);
            }
            // ---- the NxN coding unit (:1528-1545), first warp of the group ----
            if (bw == 0) {
                int sse = 0;
                for (int q = lane; q < 64; q += 32) { const int d = (int)sm.orig[y0 + (q >> 3)][x0 + (q & 7)] - (int)sm.nxn_rec[q]; sse += d * d; }
                sse = group_sum<32>(sse);
                if (lane == 0) {
                    Coder cd; cd.load(entry); cd.nout = 0;
                    attach(cd, sm, sm.cctx[SET][NXN], K.scratch[SET] + (size_t)NXN * SCR_BYTES, SCR_BYTES);
                    int modes[4], left[4], above[4];
                    for (int i = 0; i < 4; i++) modes[i] = sm.nxn_mode[i];
                    left[0] = mode_l;                  above[0] = mode_a;                               // :1530-1537
                    left[1] = modes[0];                above[1] = sm.mmd[cy][1 + cx + 1];
                    left[2] = sm.mmd[1 + cy + 1][cx];  above[2] = modes[0];
                    left[3] = modes[2];                above[3] = modes[1];
                    PreList pre; pre.p = sm.tpre[NXN]; pre.n = 0;
                    build_cu_head(pre, 8, 2, pack4(modes), pack4(left), pack4(above));
                    for (int i = 0; i < 4; i++) {
                        const int cbf = sm.nxn_cg[i] != 0;
                        pre.ctx(CX_YCBF + 0, cbf);
                        if (cbf) code_block(cd, 4, modes[i], sm.nxn_last[i], sm.nxn_cg[i], 0u, 1u, 0u, sm.nxn_list[i], &sm.nxn_hdr[i], sm.scan, pre);
                    }
                    emit_list(cd, pre, sm.scan, sm.nxn_list[0]);
                    cd.store(sm.ccoder[SET][NXN]);
                    sm.cbits[SET][NXN] = cd.bits() - entry_bits_rel;
                    sm.csse[SET][NXN] = sse;
                }
            }
        }
    }
    HPROF(K, LV * 16 + 12);
    __syncthreads();
    HPROF(K, LV * 16 + 5);
    // decide (:1437-1446, :1473-1482, :1543-1545)
    if (warp == 0) {
        int bc, win;
        warp_best<SET>(sm, rd, 0, NCAND, lane, bc, win);
        if (bc >= rd.cost(sm.csse[SET][NXN], sm.cbits[SET][NXN])) win = NXN;
        if (lane == 0) sm.win = win;
    }
    __syncthreads();
    HPROF(K, LV * 16 + 6);
    {
        // commit: coder, bytes, maps, reconstruction (:1439-1446, :1546-1554)
        const int win = sm.win;
        const unsigned base = entry.nout, n = sm.ccoder[SET][win].nout;
        const uint8_t *bytes = K.scratch[SET] + (size_t)win * SCR_BYTES;
        for (unsigned i = tid; i < n && i < SCR_BYTES; i += NT) if (base + i < out_cap) K.out[base + i] = bytes[i];
        copy_words(sm.cur_ctx, sm.cctx[SET][win], CXP / 4, tid, NT);
        if (tid == 0) { sm.cur = sm.ccoder[SET][win]; sm.cur.nout = base + n; }
        if (win == NXN) {
            if (tid < 4) {
                const int r = 1 + cy + (tid >> 1), c = 1 + cx + (tid & 1);
                sm.msz[r][c] = 8;
                sm.mmd[r][c] = (uint8_t)sm.nxn_mode[tid];
            }
            if (tid >= 64 && tid < 128) sm.rc[1 + y0 + ((tid - 64) >> 3)][1 + x0 + (tid & 7)] = sm.nxn_rec[tid - 64];
        } else {
            fill_maps<SZ>(sm, y0, x0, win >= NMODE ? win - NMODE : win);
            if (tid >= 64 && tid < 128) sm.rc[1 + y0 + ((tid - 64) >> 3)][1 + x0 + (tid & 7)] = sm.rec1[win][tid - 64];
        }
    }
    __syncthreads();
    HPROF(K, LV * 16 + 7);
}

// The same search with the 4x4 rounds of all warps in lock step (the first form of this kernel; Params::cu8_lockstep selects it).
// RD search of an 8x8 coding unit (processCURecurs :1349-1559 with sz == 8: no split; one 8x8 TU, four 4x4 TUs, or
// four 4x4 PUs).  The reference's steps 2, 3 and 4 all start from the unit's entry state, so they run side by side:
//   all warps     one-TU candidates (35 x TU8): phase A
//   warps 10,11   CABAC trials of the one-TU candidates                                 } concurrently, the ten
//   warps 0..9    four rounds, one per 4x4 block i: warps 0-4 block i of the 35 four-TU  } warps synchronise on a
//                 candidates (:1454-1467), warps 5-9 the 35 stand-alone modes of PU i     } named barrier
//                 (:1498-1527; for i = 0 these are the same transform units as the four-TU candidates' first block),
//                 then the stand-alone trials on fresh coders, the choice and the kept 4x4 reconstruction
//   warps 2,3     CABAC trials of the four-TU candidates;  warp 0: the NxN trial (:1528-1545)
// and the decision follows the reference order: one-TU modes, four-TU modes, NxN, a later candidate wins ties.
HEVC_PHASE void search_cu8_lockstep(Ctx &K, int y0, int x0, int avail)
{
    Smem &sm = HEVC_SMEM(K);
    constexpr int SZ = 8, LV = 2, H = 4, SET = 1, RT = 320;                   // RT: threads of the round warps
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qpd6 = K.qpd6;
    const RdK rd = K.rd;
    const int cy = y0 / 4, cx = x0 / 4;
    const bool hl = avail & 1, hbl = avail & 2, ha = avail & 4, har = avail & 8;
    const unsigned out_cap = (unsigned)K.P->out_stride;

    __syncthreads();
    // ---- entry state (:1363-1372) ----
    copy_words(sm.entry_ctx[LV], sm.cur_ctx, CXP / 4, tid, NT);
    if (tid == 0) sm.entry[LV] = sm.cur;
    const int mode_l = sm.mmd[1 + cy][cx], mode_a = sm.mmd[cy][1 + cx];
    // candidate contexts: the trials of the whole unit start from the entry state, the stand-alone 4x4 trials are fresh
    for (int i = tid; i < NALL * (CXP / 4); i += NT) {
        const int c = i / (CXP / 4), w = i % (CXP / 4);
        ((unsigned *)sm.cctx[SET][c])[w] = (c >= NCAND && c < NXN) ? ((const unsigned *)sm.init_ctx)[w] : ((const unsigned *)sm.cur_ctx)[w];
    }
    __syncthreads();
    const CoderRegs entry = sm.entry[LV];
    const int entry_bits_rel = 8 * entry.pending + 23 - entry.nbits;
    HPROF(K, LV * 16 + 0);

    const int sub_y[4] = {y0, y0, y0 + H, y0 + H}, sub_x[4] = {x0, x0 + H, x0, x0 + H};
    const int sub_avail[4] = {                                                        // :1375-1378
        (hl ? 1 : 0) | (hl ? 2 : 0) | (ha ? 4 : 0) | (ha ? 8 : 0),
        1 | 0 | (ha ? 4 : 0) | (har ? 8 : 0),
        (hl ? 1 : 0) | (hbl ? 2 : 0) | 4 | 8,
        1 | 0 | 4 | 0 };
    Src src; src.rc = &sm.rc[0][0]; src.priv = nullptr; src.py0 = y0; src.px0 = x0; src.psz = SZ; src.ppitch = SZ;

    // Warps 0-4 start the chain of the four stand-alone PUs at once (round 0 = block 0 of the four-TU candidates, the same
    // transform units); warps 5-11 run the one-TU pipelines meanwhile and warps 5-9 join the rounds when round 1 begins.
    //   barrier 4   the round warps 0-9 (RT threads); in round 0 warps 5-9 only meet its last one
    //   barrier 9   warps 0-4 inside round 0 (AT threads)
    //   barrier 10  one-TU candidates 0..19 are in: warps 5-9 arrive, warps 10-11 wait (their own 20..34 included)
    //   barriers 5-8  block i of the four-TU candidates is in: warps 0-4 arrive, warps 10-11 wait (trial_cu8_four)
    constexpr int AT = 160, CT = 64;
    if (warp >= 5) {
        // ---- one-TU candidates 0..34 (:1419-1447), phase A: four transform units per warp and pass; the second pass is warps 10-11's ----
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1 && warp < 10) break;
            const int item = (pass == 0 ? (warp - 5) * 4 : 28 + (warp - 10) * 4) + lane / SZ;
            const bool active = item < NMODE;
            const int m = active ? item : 0;
            const int loff = (lane / SZ) * SZ * (SZ + 2);
            TuResult r = tu_pipeline<SZ>(sm, warp, lane, active, src, y0, x0, avail, m, qpd6, rd, sm.lev[warp] + loff, sm.rec1[m], SZ);
            if (lane % SZ == 0) {
                publish_tu<SET>(sm, K.lists[SET], warp, lane / SZ, active, m, 0, 4, m, warp * LEV_STRIDE + loff, r);
                if (active) sm.csse[SET][m] = r.sse;
            }
            __syncwarp();
            binarize_warp<SZ>(sm, warp, lane, K.lists[SET]);
            __syncwarp();
        }
    }
    HPROF(K, LV * 16 + 2);

    if (warp >= 10) {
        // ---- CABAC trials of the one-TU candidates, then of the four-TU ones block by block as the rounds deliver them ----
        HEVC_NAMED_BARRIER(10, AT + CT);
        const int c = (warp - 10) * 18 + lane;
        const bool on = lane < 18 && c < NMODE;
        if (on) trial_cu<SZ>(K, c, entry, entry_bits_rel, 0, mode_l, mode_a);
        trial_cu8_four(K, on ? c : -1, entry, entry_bits_rel, mode_l, mode_a, 5, AT + CT);
    } else {
        // ---- four rounds over the 4x4 blocks ----
        const bool grp3 = warp < 5;                                         // four-TU candidates; else stand-alone PUs
        const int item = (grp3 ? warp : warp - 5) * 8 + lane / H;
        const bool active = item < NMODE;
        const int m = active ? item : 0;
        const int loff = (lane / H) * H * (H + 2);
        int sse3 = 0;
        for (int i = 0; i < 4; i++) {
            if (grp3 || i > 0) {
                // one call site for both groups (the pipeline is big; two inlined copies cost instruction-cache misses)
                const int cand = grp3 ? NMODE + m : NCAND + m;
                Src ps = src;
                ps.priv = !active ? nullptr : grp3 ? sm.rec1[cand] : sm.nxn_rec;   // earlier blocks of the same trial | the kept blocks of the earlier PUs (idle groups: none)
                ps.ppitch = 8;
                uint8_t *rec = grp3 ? sm.rec1[cand] + (sub_y[i] - y0) * SZ + (sub_x[i] - x0) : sm.rec4[m];
                TuResult r = tu_pipeline<H>(sm, warp, lane, active, ps, sub_y[i], sub_x[i], sub_avail[i], m, qpd6, rd,
                                            sm.lev[warp] + loff, rec, grp3 ? SZ : H);
                sse3 += r.sse;
                if (lane % H == 0) {
                    publish_tu<SET, ListsCu8>(sm, cu8_lists(sm), warp, lane / H, active, cand, grp3 ? i : 0, 1, m, warp * LEV_STRIDE + loff, r);
                    if (active) {
                        if (!grp3) sm.csse[SET][cand] = r.sse;
                        else if (i == 0) sm.sse0[m] = r.sse;
                        else if (i == 3) sm.csse[SET][cand] = sse3;
                    }
                }
                __syncwarp();
                binarize_warp<H, ListsCu8>(sm, warp, lane, cu8_lists(sm));
                __syncwarp();
                if (grp3) HEVC_NAMED_ARRIVE(5 + i, AT + CT);                // block i of the four-TU candidates is in: warps 10, 11 may code it
            }
            // round 0 is warps 0-4's alone; warps 5-9 come from the one-TU pipelines and wait for the kept block 0
            const bool solo = i == 0;
            if (solo && !grp3) {
                __syncwarp();
                HEVC_NAMED_ARRIVE(10, AT + CT);                             // one-TU candidates 0..19 are in
                HEVC_NAMED_BARRIER(4, RT);
                continue;
            }
            const int rbar = solo ? 9 : 4, rthr = solo ? AT : RT;
            HEVC_NAMED_BARRIER(rbar, rthr);
            HPROF(K, LV * 16 + 8);
            // stand-alone trial of every mode on a fresh coder: cost = absolute length of the block's residual (:1504-1518);
            // four lanes of warps 0-8, in round 0 seven lanes of warps 0-4
            {
                const int c = solo ? (lane < 7 ? warp * 7 + lane : -1) : candidate35_of_thread(tid, 0);
                if (c >= 0) {
                    const int from = i == 0 ? NMODE + c : NCAND + c;
                    Coder cd; cd.fresh();
                    attach(cd, sm, sm.cctx[SET][NCAND + c], K.scratch[SET] + (size_t)(NCAND + c) * SCR_BYTES, SCR_BYTES);
                    const unsigned short *lst = cu8_lists(sm) + (size_t)(from - ListsCu8::BASE) * ListsCu8::CAND;
                    PreList pre; pre.p = sm.tpre[NCAND + c]; pre.n = 0;
                    code_block(cd, 4, c, sm.clast[SET][from][0], sm.ccg[SET][from][0][0], 0u, sm.csg[SET][from][0][0], 0u, lst, lst + ListsCu8::HDR, sm.scan, pre);
                    sm.cbits[SET][NCAND + c] = cd.bits();
                    if (i == 0) sm.csse[SET][NCAND + c] = sm.sse0[c];
                }
            }
            HPROF(K, LV * 16 + 9);
            HEVC_NAMED_BARRIER(rbar, rthr);
            HPROF(K, LV * 16 + 10);
            if (warp == 0) {
                // the choice, and what is kept of the winner for the rounds to come and for the NxN trial: its mode, last position and
                // group mask, its bins, its reconstruction -- all by this warp, while the others refresh the trial contexts
                int bc, win;
                warp_best<SET>(sm, rd, NCAND, NMODE, lane, bc, win);
                win -= NCAND;
                const int from = i == 0 ? NMODE + win : NCAND + win;
                if (lane == 0) { sm.nxn_mode[i] = win; sm.nxn_last[i] = sm.clast[SET][from][0]; sm.nxn_cg[i] = sm.ccg[SET][from][0][0]; }
                const unsigned short *lst = cu8_lists(sm) + (size_t)(from - ListsCu8::BASE) * ListsCu8::CAND;
                for (int k = lane; k < SLOT; k += 32) sm.nxn_list[i][k] = lst[k];
                if (lane == 16) sm.nxn_hdr[i] = lst[ListsCu8::HDR];
                if (lane < 16) {
                    const int y = lane >> 2, x = lane & 3;
                    sm.nxn_rec[(sub_y[i] - y0 + y) * 8 + (sub_x[i] - x0 + x)] = i == 0 ? sm.rec1[NMODE + win][y * 8 + x] : sm.rec4[win][y * 4 + x];
                }
            }
            // fresh contexts for the next round's trials
            for (int k = tid; k < NMODE * (CXP / 4); k += rthr)
                ((unsigned *)sm.cctx[SET][NCAND])[k] = ((const unsigned *)sm.init_ctx)[k % (CXP / 4)];
            HEVC_NAMED_BARRIER(4, RT);                                      // (round 0: warps 5-9 join here)
            HPROF(K, LV * 16 + 11);
        }
    }

    // ---- the NxN coding unit (warp 0), next to the four-TU trials on warps 10, 11 ----
    if (warp == 0) {
        int sse = 0;
        for (int q = lane; q < 64; q += 32) { const int d = (int)sm.orig[y0 + (q >> 3)][x0 + (q & 7)] - (int)sm.nxn_rec[q]; sse += d * d; }
        sse = group_sum<32>(sse);
        if (lane == 0) {
            Coder cd; cd.load(entry); cd.nout = 0;
            attach(cd, sm, sm.cctx[SET][NXN], K.scratch[SET] + (size_t)NXN * SCR_BYTES, SCR_BYTES);
            int modes[4], left[4], above[4];
            for (int i = 0; i < 4; i++) modes[i] = sm.nxn_mode[i];
            left[0] = mode_l;                  above[0] = mode_a;                               // :1530-1537
            left[1] = modes[0];                above[1] = sm.mmd[cy][1 + cx + 1];
            left[2] = sm.mmd[1 + cy + 1][cx];  above[2] = modes[0];
            left[3] = modes[2];                above[3] = modes[1];
            PreList pre; pre.p = sm.tpre[NXN]; pre.n = 0;
            build_cu_head(pre, 8, 2, pack4(modes), pack4(left), pack4(above));
            for (int i = 0; i < 4; i++) {
                const int cbf = sm.nxn_cg[i] != 0;
                pre.ctx(CX_YCBF + 0, cbf);
                if (cbf) code_block(cd, 4, modes[i], sm.nxn_last[i], sm.nxn_cg[i], 0u, 1u, 0u, sm.nxn_list[i], &sm.nxn_hdr[i], sm.scan, pre);
            }
            emit_list(cd, pre, sm.scan, sm.nxn_list[0]);
            cd.store(sm.ccoder[SET][NXN]);
            sm.cbits[SET][NXN] = cd.bits() - entry_bits_rel;
            sm.csse[SET][NXN] = sse;
        }
    }
    HPROF(K, LV * 16 + 12);
    __syncthreads();
    HPROF(K, LV * 16 + 5);
    // decide (:1437-1446, :1473-1482, :1543-1545)
    if (warp == 0) {
        int bc, win;
        warp_best<SET>(sm, rd, 0, NCAND, lane, bc, win);
        if (bc >= rd.cost(sm.csse[SET][NXN], sm.cbits[SET][NXN])) win = NXN;
        if (lane == 0) sm.win = win;
    }
    __syncthreads();
    HPROF(K, LV * 16 + 6);
    {
        // commit: coder, bytes, maps, reconstruction (:1439-1446, :1546-1554)
        const int win = sm.win;
        const unsigned base = entry.nout, n = sm.ccoder[SET][win].nout;
        const uint8_t *bytes = K.scratch[SET] + (size_t)win * SCR_BYTES;
        for (unsigned i = tid; i < n && i < SCR_BYTES; i += NT) if (base + i < out_cap) K.out[base + i] = bytes[i];
        copy_words(sm.cur_ctx, sm.cctx[SET][win], CXP / 4, tid, NT);
        if (tid == 0) { sm.cur = sm.ccoder[SET][win]; sm.cur.nout = base + n; }
        if (win == NXN) {
            if (tid < 4) {
                const int r = 1 + cy + (tid >> 1), c = 1 + cx + (tid & 1);
                sm.msz[r][c] = 8;
                sm.mmd[r][c] = (uint8_t)sm.nxn_mode[tid];
            }
            if (tid >= 64 && tid < 128) sm.rc[1 + y0 + ((tid - 64) >> 3)][1 + x0 + (tid & 7)] = sm.nxn_rec[tid - 64];
        } else {
            fill_maps<SZ>(sm, y0, x0, win >= NMODE ? win - NMODE : win);
            if (tid >= 64 && tid < 128) sm.rc[1 + y0 + ((tid - 64) >> 3)][1 + x0 + (tid & 7)] = sm.rec1[win][tid - 64];
        }
    }
    __syncthreads();
    HPROF(K, LV * 16 + 7);
}

// ------------------------------------------------------------------------------------------------
// parameter sets + slice header (putHeaderToBuffer :664-690); one thread
// ------------------------------------------------------------------------------------------------
struct HeaderBits {
    uint8_t *p; int bit; unsigned acc;
    __device__ void put(unsigned v, int n)
    {
        for (n--; n >= 0; n--) {
            acc = (acc << 1) | ((v >> n) & 1u);
            if (++bit == 8) { *p++ = (uint8_t)acc; acc = 0; bit = 0; }
        }
    }
    __device__ void ue(unsigned v)                      // putUVLCtoBuffer :641-647
    {
        int len = 1;
        v++;
        for (unsigned t = v + 1; t != 1; t >>= 1) len += 2;
        put(v & ((1u << ((len + 1) >> 1)) - 1u), (len >> 1) + ((len + 1) >> 1));
    }
    __device__ void align() { if (bit) { acc <<= (8 - bit); *p++ = (uint8_t)acc; acc = 0; bit = 0; } }
};

__device__ inline unsigned put_headers(uint8_t *out, int qpd6, int hp, int wp)
{
    uint8_t *p = out;
    for (int i = 0; i < 27; i++) *p++ = c_vps[i];
    for (int i = 0; i < 22; i++) *p++ = c_sps[i];
    HeaderBits w; w.p = p; w.bit = 0; w.acc = 0;
    w.put(0x0A, 4); w.ue((unsigned)wp); w.ue((unsigned)hp);
    w.put(0x197EE4, 22); w.put(0x681ED1, 24);
    w.align();
    p = w.p;
    for (int i = 0; i < 11; i++) *p++ = c_pps[i];
    for (int i = 0; i < 6; i++) *p++ = c_slice_head[i];
    *p++ = c_slice_tail[qpd6][0]; *p++ = c_slice_tail[qpd6][1];
    return (unsigned)(p - out);
}

// ------------------------------------------------------------------------------------------------
// the frame kernel (HEVCImageEncoder :1569-1646): one CTA per frame, frames taken from a queue
// ------------------------------------------------------------------------------------------------
__device__ inline void build_tables(Smem &sm, int qpd6)
{
    const int tid = threadIdx.x;
    for (int i = tid; i < 64; i += NT) sm.lpsw[i] = (unsigned)c_lps[i][0] | ((unsigned)c_lps[i][1] << 8) | ((unsigned)c_lps[i][2] << 16) | ((unsigned)c_lps[i][3] << 24);
    for (int i = tid; i < 128; i += NT) sm.next_lps[i] = c_next_lps[i];
    if (tid < CXP) {                                                        // initContextValue :726-734
        int s = 0;
        if (tid < CXN) {
            const int v = c_ctx_init[tid], qp = qpd6 * 6 + 4;
            const int st = clampi(((((v >> 4) * 5 - 45) * qp) >> 4) + ((v & 15) << 3) - 16, 1, 126);
            s = st >= 64 ? (((st - 64) << 1) | 1) : ((63 - st) << 1);
        }
        sm.init_ctx[tid] = (uint8_t)s;
    }
    if (tid < 4) {                                                          // up-right diagonal order of an n x n grid
        const int n = 1 << tid;
        int k = 0;
        for (int d = 0; d <= 2 * n - 2; d++)
            for (int y = d < n - 1 ? d : n - 1; y >= (d - n + 1 > 0 ? d - n + 1 : 0); y--) { sm.dord[tid][k][0] = (uint8_t)y; sm.dord[tid][k][1] = (uint8_t)(d - y); k++; }
    }
    __syncthreads();
    for (int i = tid; i < NSCAN; i += NT) {                                 // scan tables (:1126-1150)
        int y, x;
        if (i < 1360) {
            const int sc = i < 16 ? 0 : i < 80 ? 1 : i < 336 ? 2 : 3;
            const int k = i - scan_base(0, sc), gidx = k >> 4, p = k & 15;
            y = sm.dord[sc][gidx][0] * 4 + sm.dord[2][p][0];
            x = sm.dord[sc][gidx][1] * 4 + sm.dord[2][p][1];
        } else {
            const int st = i < 1440 ? 1 : 2, r = i - (st == 1 ? 1360 : 1440);
            const int sc = r < 16 ? 0 : 1, k = r - (sc ? 16 : 0), ncg = 1 << sc, gidx = k >> 4, p = k & 15;
            const int a = (gidx / ncg) * 4 + p / 4, b = (gidx % ncg) * 4 + p % 4;
            y = st == 1 ? a : b; x = st == 1 ? b : a;
        }
        sm.scan[i][0] = (uint8_t)y; sm.scan[i][1] = (uint8_t)x;
    }
    __syncthreads();
}

// original block with edge replication, neighbours of the reconstruction (:1613-1621) and, for the CTA that keeps the
// size / mode maps, their borders
__device__ __forceinline__ void load_ctu(Ctx &K, const uint8_t *img, const uint8_t *rcon, int y, int x, bool maps)
{
    Smem &sm = HEVC_SMEM(K);
    const Params &P = *K.P;
    const int tid = threadIdx.x, H = P.H, W = P.W, wp = P.wp;
    const bool hl = x > 0, ha = y > 0;
    for (int i = tid; i < CTU * CTU; i += NT) {
        const int yy = min(y + i / CTU, H - 1), xx = min(x + i % CTU, W - 1);
        sm.orig[i / CTU][i % CTU] = img[(size_t)yy * W + xx];
    }
    if (tid < CTU) sm.rc[1 + tid][0] = hl ? rcon[(size_t)(y + tid) * wp + x - 1] : 0;
    if (tid >= 64 && tid < 64 + 1 + 2 * CTU) {
        const int j = tid - 64 - 1;                                     // -1 .. 63
        const int xx = clampi(x + j, 0, wp - 1);
        sm.rc[0][1 + j] = ha ? rcon[(size_t)(y - 1) * wp + xx] : 0;
    }
    if (maps && tid >= 160 && tid < 160 + 81) {
        const int r = (tid - 160) / 9, c = (tid - 160) % 9;
        uint8_t s = 32, m = 1;                                          // outside the picture: size 32, DC
        if (r == 0) { if (c > 0) s = sm.above_sz[x / 4 + c - 1]; }
        else if (c == 0) { if (hl) { s = sm.left_sz[r - 1]; m = sm.left_md[r - 1]; } }
        if (r == 0 || c == 0) { sm.msz[r][c] = s; sm.mmd[r][c] = m; }
    }
}

// reconstruction out (:1625-1627), map borders for the next CTUs, end_of_slice_segment_flag (:1629)
__device__ __forceinline__ void finish_ctu(Ctx &K, uint8_t *rcon, int y, int x)
{
    Smem &sm = HEVC_SMEM(K);
    const Params &P = *K.P;
    const int tid = threadIdx.x, hp = P.hp, wp = P.wp;
    for (int i = tid; i < CTU * CTU / 4; i += NT) {
        const int r = i / 8, c4 = (i % 8) * 4;
        const uint8_t *s = &sm.rc[1 + r][1 + c4];
        *(unsigned *)(rcon + (size_t)(y + r) * wp + x + c4) = (unsigned)s[0] | ((unsigned)s[1] << 8) | ((unsigned)s[2] << 16) | ((unsigned)s[3] << 24);
    }
    if (tid < 8) { sm.left_sz[tid] = sm.msz[1 + tid][8]; sm.left_md[tid] = sm.mmd[1 + tid][8]; sm.above_sz[x / 4 + tid] = sm.msz[8][1 + tid]; }
    if (tid == 0) {
        Coder c; c.load(sm.cur); attach(c, sm, sm.cur_ctx, K.out, (unsigned)P.out_stride);
        c.terminate(y + CTU >= hp && x + CTU >= wp);
        c.store(sm.cur);
    }
}

// ---- two-CTA kernel: what the CTAs hand each other through distributed shared memory ----
// spine -> upper: the entry state of the unit whose candidates the upper CTA is to evaluate
template <int SZ>
__device__ __forceinline__ void push_entry(Ctx &K)
{
    Smem &sm = HEVC_SMEM(K), &pe = *K.peer;
    constexpr int LV = Level<SZ>::v;
    const int tid = threadIdx.x;
    copy_words(pe.entry_ctx[LV], sm.entry_ctx[LV], CXP / 4, tid, NT);
    if (tid >= 64 && tid < 64 + (int)(sizeof(CoderRegs) / 4)) ((unsigned *)&pe.entry[LV])[tid - 64] = ((const unsigned *)&sm.entry[LV])[tid - 64];
    if (tid >= 96 && tid < 96 + (int)(sizeof(CuPar) / 4)) ((unsigned *)&pe.par[LV])[tid - 96] = ((const unsigned *)&sm.par[LV])[tid - 96];
}
// spine -> upper: the final reconstruction of a 16x16 unit (reference samples of the units that follow)
__device__ __forceinline__ void push_rc_block(Ctx &K, int y0, int x0)
{
    Smem &sm = HEVC_SMEM(K), &pe = *K.peer;
    const int tid = threadIdx.x;
    if (tid < 256) pe.rc[1 + y0 + (tid >> 4)][1 + x0 + (tid & 15)] = sm.rc[1 + y0 + (tid >> 4)][1 + x0 + (tid & 15)];
}
// upper -> spine: the best candidate of an evaluation: cost, index, coder, contexts, reconstruction (its bytes stay in
// the upper CTA's trial stream in global memory)
template <int SZ>
__device__ __forceinline__ void push_result(Ctx &K)
{
    Smem &sm = HEVC_SMEM(K), &pe = *K.peer;
    constexpr int LV = Level<SZ>::v, SET = Level<SZ>::set;
    const int tid = threadIdx.x, win = sm.eval_win[LV];
    copy_words(pe.res_ctx[LV], sm.cctx[SET][win], CXP / 4, tid, NT);
    copy_words(SET == 0 ? pe.res_rec0 : pe.res_rec1, rec_tile<SET>(sm, win), SZ * SZ / 4, tid, NT);
    if (tid >= 64 && tid < 64 + (int)(sizeof(CoderRegs) / 4)) ((unsigned *)&pe.res_coder[LV])[tid - 64] = ((const unsigned *)&sm.ccoder[SET][win])[tid - 64];
    if (tid == 96) { pe.eval_cost[LV] = sm.eval_cost[LV]; pe.eval_win[LV] = win; }
}

// One frame (HEVCImageEncoder :1569-1646).
//   PAIR = false: one CTA does everything, depth first.
//   PAIR = true : a cluster of two CTAs.  Rank 0 ("spine") owns the arithmetic coder, the contexts, the reconstruction
//     and the maps, searches the 8x8 units and takes every decision; rank 1 ("upper") evaluates the 70 candidates of the
//     32x32 unit and of each 16x16 unit meanwhile.  Per 16x16 unit the two meet twice at a cluster barrier: once when the
//     unit's entry state is in the upper CTA's shared memory, once when the upper CTA's result is in the spine's.
template <bool PAIR>
__device__ inline void encode_frame(Ctx &K, int job, int rank)
{
    Smem &sm = HEVC_SMEM(K);
    const Params &P = *K.P;
    const int tid = threadIdx.x;
    const int hp = P.hp, wp = P.wp, qpd6 = K.qpd6;
    const int f = job;                                                          // slot of the stream / reconstruction / length
    const uint8_t *img = P.frames + (size_t)(P.nq > 0 ? job / P.nq : job) * P.frame_stride;
    uint8_t *rcon = P.rcon + (size_t)f * P.rcon_stride;
    K.out = P.out + (size_t)f * P.out_stride;

    __syncthreads();
    copy_words(sm.cur_ctx, sm.init_ctx, CXP / 4, tid, NT);
    for (int i = tid; i < 2048; i += NT) sm.above_sz[i] = 32;                // :1594-1599
    if (tid == 0 && rank == 0) {
        Coder c; c.fresh(); c.store(sm.cur);
        sm.cur.nout = P.out_stride >= 128 ? put_headers(K.out, qpd6, hp, wp) : 0;
        if (P.out_stride < 128) atomicExch(P.overflow, 1);
    }
    __syncthreads();

    for (int y = 0; y < hp; y += CTU) {
        for (int x = 0; x < wp; x += CTU) {
            const bool hl = x > 0, ha = y > 0, har = ha && (x + CTU < wp);      // :1605-1608
            const int avail = (hl ? 1 : 0) | (ha ? 4 : 0) | (har ? 8 : 0);
            load_ctu(K, img, rcon, y, x, rank == 0);
            HPROF(K, 48);
            if (!PAIR) {
                search_cu<CTU>(K, 0, 0, avail);
                finish_ctu(K, rcon, y, x);
                __syncthreads();
            } else if (rank == 0) {
                cu_enter<32>(K, 0, 0, avail);
                push_entry<32>(K);
                put_split_flag<32>(K);
                const SubUnits<32> sub(0, 0, avail);
                for (int j = 0; j < 4; j++) {
                    cu_enter<16>(K, sub.y[j], sub.x[j], sub.avail[j]);
                    push_entry<16>(K);
                    HPROF(K, 50);
                    HEVC_PAIR_SYNC();                                       // the upper CTA may evaluate this unit
                    HPROF(K, 51);
                    put_split_flag<16>(K);
                    const SubUnits<16> s8(sub.y[j], sub.x[j], sub.avail[j]);
                    for (int k = 0; k < 4; k++) search_cu8(K, s8.y[k], s8.x[k], s8.avail[k]);
                    cu_split_cost<16>(K);
                    HPROF(K, 52);
                    HEVC_PAIR_SYNC();                                       // its result is here
                    HPROF(K, 53);
                    cu_decide_commit<16, true>(K);
                    push_rc_block(K, sub.y[j], sub.x[j]);
                }
                cu_split_cost<32>(K);
                cu_decide_commit<32, true>(K);
                finish_ctu(K, rcon, y, x);
                HPROF(K, 54);
                HEVC_PAIR_SYNC();                                           // the reconstruction is in global memory
                HPROF(K, 55);
            } else {
                __syncthreads();
                for (int j = 0; j < 4; j++) {
                    HEVC_PAIR_SYNC();
                    eval_a1<16>(K); eval_a4<16>(K); eval_b<16>(K);
                    push_result<16>(K);
                    // the 32x32 unit's evaluation, a third per 16x16 unit
                    if (j == 0) eval_a1<32>(K);
                    if (j == 1) eval_a4<32>(K);
                    if (j == 2) { eval_b<32>(K); push_result<32>(K); }
                    HEVC_PAIR_SYNC();
                }
                HEVC_PAIR_SYNC();
            }
            HPROF(K, 49);
        }
    }
    if (tid == 0 && rank == 0) {
        Coder c; c.load(sm.cur); attach(c, sm, sm.cur_ctx, K.out, (unsigned)P.out_stride);
        c.finish();                                                             // :1639-1640
        const bool fits = c.nout <= (unsigned)P.out_stride;
        P.out_len[f] = fits ? c.nout : 0;
        if (!fits) atomicExch(P.overflow, 1);
    }
    __syncthreads();
}

template <bool PAIR>
__device__ __forceinline__ void frame_kernel_body(const Params &P)
{
    const int rank = PAIR ? HEVC_PAIR_RANK() : 0;
    Smem &sm = *reinterpret_cast<Smem *>(IMCVT_DYN_SMEM());
    Ctx K;
    K.sm = &sm; K.P = &P; K.qpd6 = P.qpd6; K.rd.init(P.qpd6);
    K.peer = PAIR ? reinterpret_cast<Smem *>(HEVC_PAIR_SMEM(rank ^ 1)) : nullptr;
    for (int q = 0; q < 2; q++) {
        K.scratch[q] = P.scratch + ((size_t)blockIdx.x * 2 + q) * NALL * SCR_BYTES;
        K.lists[q] = P.lists + ((size_t)blockIdx.x * 2 + q) * NALL * CAND_LIST_U16;
        K.peer_scratch[q] = P.scratch + ((size_t)(blockIdx.x ^ 1) * 2 + q) * NALL * SCR_BYTES;
    }
    K.out = nullptr;
    K.prof_on = rank == 0; K.prof_t = 0;
    build_tables(sm, P.qpd6);
#if defined(IMCVT_HEVC_PROF) && !defined(IMCVT_HOST_EMU)
    K.prof_t = clock64();
    if (threadIdx.x == 0) sm.prof = P.prof;
    __syncthreads();
#endif
    for (;;) {                                                                  // frames vary 2-3x in cost: dynamic queue
        if (PAIR) HEVC_PAIR_SYNC(); else __syncthreads();
        if (threadIdx.x == 0 && rank == 0) { const int f = atomicAdd(P.next_frame, 1); sm.win = f; if (PAIR) K.peer->win = f; }
        if (PAIR) HEVC_PAIR_SYNC(); else __syncthreads();
        int f = sm.win;
        if (f >= (P.nq > 0 ? P.n * P.nq : P.n)) break;
        if (P.nq > 0) { const int r = f / P.n; f = (f - r * P.n) * P.nq + P.qorder[r]; }      // longest jobs first: the tail of the launch is short jobs
        if (P.nq > 0 && P.qlist[f % P.nq] != K.qpd6) {                          // a sweep: the context initialisation follows the QP
            K.qpd6 = P.qlist[f % P.nq];
            K.rd.init(K.qpd6);
            __syncthreads();
            build_tables(sm, K.qpd6);
        }
        encode_frame<PAIR>(K, f, rank);
    }
}

__global__ void __launch_bounds__(NT, 1) hevc_frame_kernel(Params P) { frame_kernel_body<false>(P); }
__global__ void HEVC_CLUSTER_DIMS __launch_bounds__(NT, 1) hevc_pair_kernel(Params P) { frame_kernel_body<true>(P); }

}  // namespace hevc
}  // namespace imcvt
