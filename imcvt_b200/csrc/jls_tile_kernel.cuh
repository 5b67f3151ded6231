// jls_tile_kernel.cuh -- lossless (NEAR = 0) JPEG-LS scan encoder for sm_100a: one CTA per plane, the plane streams
// through shared memory in row tiles and never touches HBM in between (device code only).
//
// Replaces JLSencodeScan + the bit writer (reference: src/imageio_jls.c:240-399, :124-197) for NEAR = 0.
// With NEAR = 0 the reconstruction equals the input, so the causal template, the context, the MED prediction and
// the run / regular classification of every sample depend on the input only; what is serial is
// one chain per context (A,B,C,N, :376-393), the run_index chain (:291-344) and the bit position.
//
// A CTA is a three-stage software pipeline over tiles of <= 6144 samples (whole rows); each `tick` (one
// __syncthreads) the three warp groups work on three consecutive tiles:
//   producers (6 warps)  tile k   : classify every sample (carry-lookahead run state over warp ballots), counting-sort
//                                   the samples by context into shared memory, stage the rows of tile k+1
//   chains   (12 warps)  tile k-1 : thread b owns context b for the whole plane (state in registers) and walks its
//                                   slice of the sorted tile; thread 364 walks the run events; each sample's Golomb
//                                   code + length go to the tile's raster-order slots
//   packer    (1 warp)   tile k-2 : prefix sum of the code lengths, MSB-first packing, marker-avoidance bit stuffing
//                                   (:156-169) straight into the plane's output stream
// HBM traffic = the pixels once + the stream once (the algorithmic minimum); the tile never leaves the SM.
//
// Also compiled for the host by tests/emu (fiber SIMT emulator, test infrastructure only).
#pragma once
#include <stdint.h>
#include <stddef.h>

// The pair kernel runs one plane on a cluster of two CTAs (rank 0 produces sorted tiles straight into the shared
// memory of rank 1).  Under the host emulator the cluster is one block of 2 x NTHREADS fibers with two Smem images.
#ifdef IMCVT_HOST_EMU
#define JLS_DYN_SMEM() (emu::S().smem)
#define JLS_NAMED_BARRIER(id, count) emu_named_barrier(id, count)
#define JLS_CONSTANT static const
#define JLS_PAIR_RANK()  (emu::S().tid / imcvt::jls::tile::NTHREADS)
#define JLS_PAIR_TID()   (emu::S().tid % imcvt::jls::tile::NTHREADS)
#define JLS_PAIR_SMEM(r) (emu::S().smem + (size_t)(r) * imcvt::jls::tile::SMEM_STRIDE)
#define JLS_OWN_SMEM(r)  JLS_PAIR_SMEM(r)
#define JLS_PAIR_SYNC()  __syncthreads()
#define JLS_CLUSTER_DIMS
#else
#include <cooperative_groups.h>
extern __shared__ __align__(16) unsigned char jls_dyn_smem[];
#define JLS_DYN_SMEM() (jls_dyn_smem)
#define JLS_NAMED_BARRIER(id, count) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory")
#define JLS_CONSTANT __constant__ const
#define JLS_PAIR_RANK()  ((int)cooperative_groups::this_cluster().block_rank())
#define JLS_PAIR_TID()   ((int)threadIdx.x)
#define JLS_PAIR_SMEM(r) (cooperative_groups::this_cluster().map_shared_rank(jls_dyn_smem, (r)))
#define JLS_OWN_SMEM(r)  (jls_dyn_smem)      /* the CTA's own image through the shared window: LDS/STS, not generic LD/ST */
#define JLS_PAIR_SYNC()  cooperative_groups::this_cluster().sync()
#define JLS_CLUSTER_DIMS __cluster_dims__(2, 1, 1)
#endif

namespace imcvt {
namespace jls {

constexpr int NCTX    = 364;       // regular contexts (:241-244)
constexpr int NBUCKET = 365;       // + one bucket holding the run events
constexpr int HSLOTS  = 366;       // + a dummy slot for samples that produce no record

JLS_CONSTANT uint8_t c_J[32] = {0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,5,5,6,6,7,7,8,9,10,11,12,13,14,15};   // :15

// ---- small device helpers (shared with jls.cu) ----------------------------------------------------
__device__ __forceinline__ int grad_q_lossless(int g)          // gradientQuantize :68-76 with T1,T2,T3 = 3,7,21, NEAR = 0
{
    const int m = g < 0 ? -g : g;
    const int q = (m > 0) + (m >= 3) + (m >= 7) + (m >= 21);
    return g < 0 ? -q : q;
}
__device__ __forceinline__ int med3(int a, int b, int c)        // predict :87-94
{
    const int mx = max(a, b), mn = min(a, b);
    return c >= mx ? mn : (c <= mn ? mx : a + b - c);
}
__device__ __forceinline__ int golomb_k(int A, int N)           // getK :114-121, closed form
{
    int k = max(0, __clz(N) - __clz(A));                        // bitlen(A) - bitlen(N)
    if ((N << k) < A) k++;
    return k;
}
// Golomb code of GolombCoding :187-197 as (value, length): `hi` zeros, a one, k low bits -- or the escape.
__device__ __forceinline__ void golomb_code(int me, int k, int limit, int qbpp, uint32_t &code, int &len)
{
    const int hi = me >> k;
    if (hi < limit) {
        len  = hi + 1 + k;
        code = (1u << k) | ((uint32_t)me & ((1u << k) - 1u));
    } else {
        len  = limit + 1 + qbpp;
        code = (1u << qbpp) | ((uint32_t)(me - 1) & ((1u << qbpp) - 1u));
    }
}

// where the scan of plane p goes: component 0 of a frame straight into the frame's output slot, components 1,2
// (RGB) into scratch, moved by assemble_kernel once the sizes are known
struct FrameLayout {
    int ncomp;              // 1 gray, 3 RGB
    int hdr_len;            // SOI + SOF55 : 15 gray / 21 RGB   (:211-232)
    uint8_t *out;           // device output slots
    size_t out_stride;
    uint8_t *scratch;       // [P][scratch_stride]
    size_t scratch_stride;
};
__device__ __forceinline__ uint8_t *plane_dst(const FrameLayout &L, int p, size_t &cap)
{
    const int f = p / L.ncomp, c = p - f * L.ncomp;
    if (c == 0) {
        const size_t o = (size_t)L.hdr_len + 10;
        cap = L.out_stride > o ? L.out_stride - o : 0;
        return L.out + (size_t)f * L.out_stride + o;
    }
    cap = L.scratch_stride;
    return L.scratch + (size_t)p * L.scratch_stride;
}

// ------------------------------------------------------------------------------------------------
// the tile kernel
// ------------------------------------------------------------------------------------------------
namespace tile {

constexpr int TSMAX = 6144;         // samples per tile (whole rows)
constexpr int WMAX  = 6144;         // widest plane this kernel takes (wider ones use the multi-kernel path)
constexpr int CHAIN_WARPS = 12, PROD_WARPS = 10, PACK_WARPS = 4;
constexpr int NWARPS = CHAIN_WARPS + PROD_WARPS + PACK_WARPS;
constexpr int NTHREADS = NWARPS * 32;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int PIXBUF = 16 + TSMAX + WMAX + 16;
constexpr int CHUNK = 1024;         // samples per packer chunk (32 per lane)
constexpr int UBITS_WORDS = TSMAX + 8;   // un-stuffed bits of one tile: <= 32 per sample
// codes / lens of a tile are indexed with one pad word per 32 samples: the packer's lanes read 32 consecutive samples
// each, and the pad spreads them over the shared-memory banks
constexpr int CODE_SLOTS = TSMAX + TSMAX / 32, LEN_SLOTS = TSMAX + TSMAX / 8;
__device__ __forceinline__ int code_idx(int lp) { return lp + (lp >> 5); }
__device__ __forceinline__ int len_idx(int lp) { return lp + ((lp >> 5) << 2); }

struct Params {
    const uint8_t *pix; size_t plane_stride;   // P planes, H x W u8
    int P, W, H;
    FrameLayout L;
    unsigned long long *plane_bytes;           // [P] stuffed scan length
    int *next_plane;                           // plane queue
    unsigned long long *prof;                  // role cycle counters (development build IMCVT_JLS_PROF only)
    int line_out;                              // 1: the output slots are page-locked HOST memory: write whole 128-byte lines
    int tma;                                   // 1: row tiles are staged by bulk async copies (cp.async.bulk + mbarrier) when the plane allows it
};

#if defined(IMCVT_JLS_PROF) && !defined(IMCVT_HOST_EMU)
#define JPROF_T0() const long long jp_t0 = clock64()
#define JPROF(id, cond) do { if (cond) atomicAdd(&P.prof[(id)], (unsigned long long)(clock64() - jp_t0)); } while (0)
#else
#define JPROF_T0() do { } while (0)
#define JPROF(id, cond) do { } while (0)
#endif

struct Smem {
    uint32_t sorted[2][TSMAX];      // records of a tile, grouped by context, raster order inside a context
    uint32_t codes[2][CODE_SLOTS];  // Golomb code per sample, raster order (code_idx)
    uint32_t temp_rec[TSMAX];       // producers: bucket | record payload << 9 per sample, raster order (pass 1 -> pass 2)
    uint16_t bstart[3][HSLOTS + 2]; // start of every context's slice inside the sorted tile (third entry: pair kernel)
    alignas(4) uint16_t cursor[NWARPS][HSLOTS + 2];       // producers: per-warp histogram, then scatter cursor (pair kernel: every warp of rank 0)
    uint32_t ubits[UBITS_WORDS];    // packer: un-stuffed bits of one tile
    uint32_t chunk_bits[8];         // packer: code bits per chunk
    uint8_t  lens[2][LEN_SLOTS];    // code length per sample (0 = nothing), raster order (len_idx)
    alignas(16) uint8_t pix[3][PIXBUF];     // rows of a tile (+ the row above, + the sample two rows above column 0); 16-byte aligned: bulk copies land here
    alignas(8) unsigned long long stage_bar;   // mbarrier of the bulk copy in flight (one at a time: issued at the start of a tick, awaited at its end)
    int ctx[NCTX][4];               // (A, B, C, N) of every regular context of the plane in flight
    int ctx_a[NCTX];                // pair kernel: A lives with the second chain stage
    int run_state[8];               // run_index, A/B/N of the two run-interruption contexts, pending run length
    uint16_t order[3][CHAIN_WARPS * 32];   // chain thread -> bucket for a tile: buckets sorted by slice size (big slices share warps)
    int cls_cnt[16], cls_pos[16];
    alignas(16) uint8_t out_ring[256];      // packer, output in host memory: stuffed bytes until a 128-byte line is complete
    uint32_t blk_tot[32];           // producers: histogram totals per block of 32 buckets
    int8_t qlut[512];               // quantised gradient of g at [g + 255]
    int plane;
};
// sorted tile number b: the pair kernel keeps three tiles in flight; the consumer CTA has no use for temp_rec
__device__ __forceinline__ uint32_t *sorted_buf(Smem &sm, int b) { return b < 2 ? sm.sorted[b] : sm.temp_rec; }
constexpr size_t SMEM_STRIDE = (sizeof(Smem) + 127) & ~(size_t)127;     // host emulator: distance of the two images of a pair

// records: regular sample  pos | (sign*(x-MED) + 256) << 13 | clamp window << 22     (rewritten by its chain as pos | (e & 255) << 13 | k << 21 | map << 25)
//          run event       pos | run length << 13 | row-end flag << 28 | partial flag << 29 (the run goes on in the next segment) | first column << 30
__device__ __forceinline__ uint8_t *pix_row(uint8_t *buf, int local_row, int W) { return buf + 16 + (local_row + 1) * W; }

// stages the rows of tile t (and the row above it) into a pix buffer; executed by `nthr` threads
__device__ __forceinline__ void stage_tile(uint8_t *buf, const uint8_t *img, int W, int H, int rows_per_tile, int t, int tid, int nthr)
{
    const int r0 = t * rows_per_tile;
    if (r0 >= H) return;
    const int r1 = min(H, r0 + rows_per_tile);
    if (r0 == 0) for (int i = tid; i < W; i += nthr) buf[16 + i] = 0;     // the row above the picture reads as zeros (:47-49)
    const int first = r0 > 0 ? r0 - 1 : 0;
    const uint8_t *src = img + (size_t)first * W;
    uint8_t *dst = buf + 16 + (first - (r0 - 1)) * W;
    const int n = (r1 - first) * W;
    // 4-byte words where source and destination agree on alignment, bytes otherwise
    const int mis = (int)((size_t)src & 3);
    if (((size_t)dst & 3) == (size_t)mis) {
        const int head = min(n, (4 - mis) & 3);
        if (tid < head) dst[tid] = src[tid];
        const int words = (n - head) >> 2;
        const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src + head);
        uint32_t *d4 = reinterpret_cast<uint32_t *>(dst + head);
        for (int i = tid; i < words; i += nthr) d4[i] = s4[i];
        const int done = head + words * 4;
        if (tid < n - done) dst[done + tid] = src[done + tid];
    } else {
        for (int i = tid; i < n; i += nthr) dst[i] = src[i];
    }
    if (tid == 0) buf[0] = r0 >= 2 ? img[(size_t)(r0 - 2) * W] : 0;
}

// stage_tile in two halves for a CTA whose every thread is a producer (the pair kernel's rank 0): the global loads of tile
// t + 1 are ISSUED at the start of the tick into registers and only stored to shared memory at its end, so their latency
// -- a PCIe round trip when the picture is read in place from page-locked host memory -- is covered by the tick's work.
template <int NREG>
struct StageRegs { uint32_t w[NREG]; uint32_t edge; bool fast; };
template <int NREG>
__device__ __forceinline__ void stage_issue(StageRegs<NREG> &g, const uint8_t *buf, const uint8_t *img, int W, int H, int rows_per_tile, int t, int tid, int nthr)
{
    g.fast = false;
    const int r0 = t * rows_per_tile;
    if (r0 >= H) return;
    const int r1 = min(H, r0 + rows_per_tile);
    const int first = r0 > 0 ? r0 - 1 : 0;
    const uint8_t *src = img + (size_t)first * W;
    const uint8_t *dst = buf + 16 + (first - (r0 - 1)) * W;
    const int n = (r1 - first) * W;
    const int mis = (int)((size_t)src & 3);
    const int head = min(n, (4 - mis) & 3), words = (n - head) >> 2, done = head + words * 4;
    if (((size_t)dst & 3) != (size_t)mis || words > NREG * nthr) return;       // stage_commit falls back to stage_tile
    g.fast = true;
    const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src + head);
#pragma unroll
    for (int k = 0; k < NREG; k++) { const int i = tid + k * nthr; g.w[k] = i < words ? s4[i] : 0u; }
    // the few bytes around the words, and the sample two rows above column 0: one byte each in `edge`
    g.edge = 0;
    if (tid < head) g.edge = src[tid];
    else if (tid >= 32 && tid - 32 < n - done) g.edge = src[done + tid - 32];
    else if (tid == 64 && r0 >= 2) g.edge = img[(size_t)(r0 - 2) * W];
}
template <int NREG>
__device__ __forceinline__ void stage_commit(const StageRegs<NREG> &g, uint8_t *buf, const uint8_t *img, int W, int H, int rows_per_tile, int t, int tid, int nthr)
{
    if (!g.fast) { stage_tile(buf, img, W, H, rows_per_tile, t, tid, nthr); return; }
    const int r0 = t * rows_per_tile;
    const int r1 = min(H, r0 + rows_per_tile);
    if (r0 == 0) for (int i = tid; i < W; i += nthr) buf[16 + i] = 0;
    const int first = r0 > 0 ? r0 - 1 : 0;
    const uint8_t *src = img + (size_t)first * W;
    uint8_t *dst = buf + 16 + (first - (r0 - 1)) * W;
    const int n = (r1 - first) * W;
    const int mis = (int)((size_t)src & 3);
    const int head = min(n, (4 - mis) & 3), words = (n - head) >> 2, done = head + words * 4;
    uint32_t *d4 = reinterpret_cast<uint32_t *>(dst + head);
#pragma unroll
    for (int k = 0; k < NREG; k++) { const int i = tid + k * nthr; if (i < words) d4[i] = g.w[k]; }
    if (tid < head) dst[tid] = (uint8_t)g.edge;
    else if (tid >= 32 && tid - 32 < n - done) dst[done + tid - 32] = (uint8_t)g.edge;
    else if (tid == 64) buf[0] = r0 >= 2 ? (uint8_t)g.edge : 0;
}

// Pair kernel, consumer CTA: the rows of tile t as the producer CTA staged them a tick earlier, fetched from ITS shared memory
// (distributed shared memory loads) instead of a second read of the picture: global memory -- or, with page-locked input
// read in place, the PCIe link -- sees every sample once.
__device__ __forceinline__ void pull_tile(uint8_t *dst, const uint8_t *remote, int W, int H, int rows_per_tile, int t, int tid, int nthr)
{
    const int r0 = t * rows_per_tile;
    if (r0 >= H) return;
    const int R = min(H, r0 + rows_per_tile) - r0;
    const int words = (16 + (R + 1) * W + 3) >> 2;          // slot of the sample two rows up | row above | R rows
    const uint32_t *s4 = reinterpret_cast<const uint32_t *>(remote);
    uint32_t *d4 = reinterpret_cast<uint32_t *>(dst);
    for (int i = tid; i < words; i += nthr) d4[i] = s4[i];
}

// ---- row tiles by bulk asynchronous copy (TMA engine: cp.async.bulk global -> shared, completion on an mbarrier) -------------
// One thread arms the barrier with the byte count and issues ONE copy for the rows of a tile (+ the row above it); nobody's
// registers or issue slots move pixels.  The copy wants 16-byte aligned source, destination and size: planes whose width is
// a multiple of 16 (and whose first byte is 16-byte aligned) qualify, the others keep the register staging.
#ifdef IMCVT_HOST_EMU
__device__ __forceinline__ void bulk_bar_init(unsigned long long *bar) { *bar = 0; }
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) { memcpy(dst, src, bytes); (void)bar; }
__device__ __forceinline__ void bulk_wait(unsigned long long *bar, uint32_t parity) { (void)bar; (void)parity; }
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_bar_init(unsigned long long *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_wait(unsigned long long *bar, uint32_t parity)
{
    const uint32_t b = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(b), "r"(parity) : "memory");
    }
}
#endif
// whether the rows of this plane can be staged by bulk copies
__device__ __forceinline__ bool bulk_ok(const Params &P, const uint8_t *img) { return P.tma && (P.W & 15) == 0 && ((size_t)img & 15) == 0; }
// issue (one thread): the rows of tile t and the row above it, as stage_tile lays them out
__device__ __forceinline__ void stage_bulk_issue(uint8_t *buf, const uint8_t *img, int W, int H, int rows_per_tile, int t, unsigned long long *bar)
{
    const int r0 = t * rows_per_tile, r1 = min(H, r0 + rows_per_tile);
    const int first = r0 > 0 ? r0 - 1 : 0;
    bulk_load(buf + 16 + (first - (r0 - 1)) * W, img + (size_t)first * W, (uint32_t)((r1 - first) * W), bar);
}
// the two things a bulk copy does not bring (all threads of the staging group; plain stores to other bytes than the copy's)
__device__ __forceinline__ void stage_bulk_edges(uint8_t *buf, const uint8_t *img, int W, int rows_per_tile, int t, int tid, int nthr)
{
    const int r0 = t * rows_per_tile;
    if (r0 == 0) for (int i = tid; i < W; i += nthr) buf[16 + i] = 0;     // the row above the picture reads as zeros (:47-49)
    if (tid == 0) buf[0] = r0 >= 2 ? img[(size_t)(r0 - 2) * W] : 0;
}

// ---- producers ------------------------------------------------------------------------------------
// Gradient quantisation (gradientQuantize :68-76 with T1,T2,T3 = 3,7,21, NEAR = 0) from a table in shared memory:
// qlut[g + 255] = quantised gradient of g in [-255, 255]; the three terms of getQ :79-84 are 81*q1 + 9*q2 + q3.
__device__ __forceinline__ void build_qlut(int8_t *qlut, int tid, int nthr)
{
    for (int i = tid; i < 511; i += nthr) qlut[i] = (int8_t)grad_q_lossless(i - 255);
}

// Per-sample classification inputs of 32 consecutive samples of a row (samplePixels :46-65, getQ :79-84).  The row above
// the picture is staged as zeros, so only column 0 and the last column need care.
struct Cls { int x, a, b, c, Q; bool active; };
__device__ __forceinline__ Cls classify_sample(const int8_t *qlut, const uint8_t *pb, const uint8_t *cur, int W, int i_local, int j)
{
    Cls s;
    s.active = j >= 0 && j < W;
    const int jc = s.active ? j : 0;
    const uint8_t *up = cur - W;
    const int jm = jc > 0 ? jc - 1 : 0, jp = jc + 1 < W ? jc + 1 : jc;
    s.x = cur[jc];
    s.b = up[jc];
    const int d = up[jp];
    int a = cur[jm], c = up[jm];
    if (jc == 0) { a = s.b; c = i_local >= 1 ? (int)up[-W] : (int)pb[0]; }
    s.a = a; s.c = c;
    s.Q = 81 * (int)qlut[d - s.b + 255] + 9 * (int)qlut[s.b - c + 255] + (int)qlut[c - a + 255];
    return s;
}

// Classifies the samples [c0, c1) of one row (c0 a multiple of 32) for one warp.  PASS 1: record + bucket per sample
// into temp_*, histogram into hist[] (16-bit counters, incremented two to a word by shared-memory atomics so that
// successive iterations do not depend on each other).  Run state (:270, :284-289): running_j = (q0_j | running_{j-1})
// & eq_j, a carry chain resolved 32 samples at a time by one 64-bit addition of the two ballot masks.  A segment that
// does not start the row finds its carry-in by looking back: the closest earlier sample with !eq (carry 0) or
// q0 & eq (carry 1) decides.
__device__ __forceinline__ void classify_segment(Smem &sm, uint8_t *pb, int W, int i_local, int c0, int c1, int lane, uint16_t *hist)
{
    const uint8_t *cur = pix_row(pb, i_local, W);
    const int8_t *qlut = sm.qlut;
    const uint32_t lane_le = 0xFFFFFFFFu >> (31 - lane);
    uint32_t cin = 0;
    for (int jb = c0 - 32; jb >= 0; jb -= 32) {
        const Cls s = classify_sample(qlut, pb, cur, W, i_local, jb + lane);
        const uint32_t eqm = __ballot_sync(0xffffffffu, s.x == s.a);
        const uint32_t q0m = __ballot_sync(0xffffffffu, s.Q == 0);
        const uint32_t dec = ~eqm | q0m;
        if (dec) { const int top = 31 - __clz((int)dec); cin = (eqm >> top) & (q0m >> top) & 1u; break; }
    }
    int run_start = c0;
    uint32_t *hist32 = reinterpret_cast<uint32_t *>(hist);
    for (int j0 = c0; j0 < c1; j0 += 32) {
        const int j = j0 + lane;
        const Cls s = classify_sample(qlut, pb, cur, W, i_local, j < c1 ? j : -1);
        const bool active = s.active;
        const int Q = s.Q;
        const bool q0 = active && Q == 0;
        const bool eq = active && s.x == s.a;
        const uint32_t act = __ballot_sync(0xffffffffu, active);
        const uint32_t Pm  = __ballot_sync(0xffffffffu, eq);
        const uint32_t Q0m = __ballot_sync(0xffffffffu, q0);
        const uint32_t Gm  = Pm & Q0m;
        const uint64_t cv  = ((uint64_t)Pm + Gm + cin) ^ Pm ^ Gm;   // bit j = carry into j = running_{j-1}
        const uint32_t prevrun = (uint32_t)cv;
        const uint32_t Rm  = (uint32_t)(cv >> 1) & act;             // running_j
        const uint32_t INM = (Q0m | prevrun) & act;                 // sample is coded in run mode (hit or interruption)
        const uint32_t S   = Q0m & ~prevrun;                        // a run-mode stretch starts here
        const bool inm = (INM >> lane) & 1u, run = (Rm >> lane) & 1u;
        const bool is_int = inm && !run;
        const bool is_eol = run && (j == W - 1);
        const bool is_part = run && !is_eol && (j == c1 - 1);       // the run goes on in the next segment of the row

        if (active) {
            int bucket = NBUCKET;                                   // dummy: plain run sample
            uint32_t rec = 0;                                       // payload: the record without its position
            const uint32_t lp = (uint32_t)(i_local * W + j);
            if (!inm) {
                // regular sample: what its chain needs -- sign * (x - MED) and the clamp window of the bias (regular_step)
                const bool neg = Q < 0;
                const int med = med3(s.a, s.b, s.c);
                bucket = (neg ? -Q : Q) - 1;
                const int d0 = neg ? med - s.x : s.x - med;
                const int nlo = neg ? 255 - med : med;
                rec = (uint32_t)(d0 + 256) | ((uint32_t)nlo << 9);
            } else if (is_int || is_eol || is_part) {
                bucket = NCTX;
                const uint32_t sm_ = S & lane_le;
                const int start_col = sm_ ? (j0 + 31 - __clz((int)sm_)) : run_start;
                rec = (uint32_t)(j - start_col + (is_int ? 0 : 1)) | ((uint32_t)is_eol << 15) | ((uint32_t)is_part << 16) | ((uint32_t)(j == 0) << 17);
            }
            sm.temp_rec[lp] = (uint32_t)bucket | (rec << 9);
            if (bucket < NBUCKET) atomicAdd(&hist32[bucket >> 1], 1u << ((bucket & 1) * 16));
        }
        cin = (uint32_t)(cv >> 32) & 1u;
        if (S) run_start = j0 + 31 - __clz((int)S);
    }
}

// PASS 2: scatter the records of one row segment into sorted[] (cursor[] = next free slot per bucket for this warp)
__device__ __forceinline__ void scatter_segment(Smem &sm, uint32_t *sorted, int W, int i_local, int c0, int c1, int lane, uint16_t *cursor)
{
    const uint32_t lane_lt = (0xFFFFFFFFu >> (31 - lane)) >> 1;
    for (int j0 = c0; j0 < c1; j0 += 32) {
        const int j = j0 + lane;
        int bucket = NBUCKET;
        uint32_t rec = 0;
        if (j < c1) { const int lp = i_local * W + j; const uint32_t t = sm.temp_rec[lp]; bucket = (int)(t & 0x1FFu); rec = (uint32_t)lp | ((t >> 9) << 13); }
        const uint32_t peers = __match_any_sync(0xffffffffu, bucket);
        const int rank = __popc(peers & lane_lt);
        const int base = cursor[bucket];
        __syncwarp();
        if (rank == 0) cursor[bucket] = (uint16_t)(base + __popc(peers));
        if (bucket < NBUCKET) sorted[base + rank] = rec;
        __syncwarp();
    }
}

constexpr int NBLK = (NBUCKET + 1 + 31) / 32;      // blocks of 32 histogram slots (the extra slot closes the last slice)

// sm: the producers' own shared memory; cons: where the sorted tile goes (the same image, or the partner CTA's)
template <int NPW>
__device__ __forceinline__ void producer_tick(Smem &sm, Smem &cons, const Params &P, const uint8_t *img, int rows_per_tile, int t, int ntiles, int nbuf, int pw, int lane,
                                              uint32_t &bulk_phase)
{
    constexpr int PROD_WARPS = NPW, PROD_THREADS = NPW * 32;
    const int W = P.W, H = P.H;
    const int ptid = pw * 32 + lane;
    JPROF_T0();
    // every thread of the CTA is a producer (pair kernel): the next tile's rows are requested now and stored at the end
    constexpr bool SPLIT_STAGE = (NPW == NWARPS);
    constexpr int NREG = SPLIT_STAGE ? (TSMAX + WMAX + 4 * NPW * 32 - 1) / (4 * NPW * 32) : 1;
    StageRegs<NREG> sregs;
    sregs.fast = false;
    // the rows of the next tile: one bulk copy issued now and awaited at the end of the tick (its buffer was last read two
    // ticks ago), or -- planes the copy engine cannot take -- loads into registers now, stores at the end
    const bool bulk = bulk_ok(P, img);
    if (bulk) { if (ptid == 0 && t + 1 < ntiles) stage_bulk_issue(sm.pix[(t + 1) % 3], img, W, H, rows_per_tile, t + 1, &sm.stage_bar); }
    else if (SPLIT_STAGE && t + 1 < ntiles) stage_issue<NREG>(sregs, sm.pix[(t + 1) % 3], img, W, H, rows_per_tile, t + 1, ptid, PROD_THREADS);
    if (t < ntiles) {
        const int buf = t % nbuf;                  // tiles in flight: 2 (one CTA) or 3 (pair)
        uint8_t *pb = sm.pix[t % 3];
        const int r0 = t * rows_per_tile, R = min(H, r0 + rows_per_tile) - r0;
        // work units = 32-sample chunks of the tile's rows in raster order; every warp takes an equal contiguous range of
        // them (keeps raster order per bucket), i.e. one row segment per row it touches
        const int chunks = (W + 31) >> 5;
        const int nunits = R * chunks;
        const int per = (nunits + PROD_WARPS - 1) / PROD_WARPS;
        const int ua = min(nunits, pw * per), ub = min(nunits, ua + per);
        const int ra = ua / chunks, rb = ub > ua ? (ub - 1) / chunks : ra - 1;       // rows [ra, rb]
        uint16_t *h = sm.cursor[pw];
        for (int k = lane; k < (HSLOTS + 2) / 2; k += 32) reinterpret_cast<uint32_t *>(h)[k] = 0;
        if (ptid < 16) { sm.cls_cnt[ptid] = 0; sm.cls_pos[ptid] = 0; }
        __syncwarp();
        for (int r = ra; r <= rb; r++) {
            const int c0 = max(ua - r * chunks, 0) * 32, c1 = min(W, min(ub - r * chunks, chunks) * 32);
            classify_segment(sm, pb, W, r, c0, c1, lane, h);
        }
        JPROF(8, ptid == 0);
        JLS_NAMED_BARRIER(1, PROD_THREADS);
        JPROF(9, ptid == 0);
        // exclusive scan over the buckets of (sum over warps), blocks of 32 buckets spread over the producer warps;
        // the per-warp histograms become scatter cursors
        for (int kb = pw; kb < NBLK; kb += PROD_WARPS) {
            const int k = kb * 32 + lane;
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < PROD_WARPS; w++) tot += k < NBUCKET ? sm.cursor[w][k] : 0;
            uint32_t inc = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
            if (lane == 31) sm.blk_tot[kb] = inc;
        }
        JLS_NAMED_BARRIER(1, PROD_THREADS);
        for (int kb = pw; kb < NBLK; kb += PROD_WARPS) {
            const int k = kb * 32 + lane;
            uint32_t bt = lane < kb ? sm.blk_tot[lane] : 0;                 // NBLK <= 32
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) bt += __shfl_xor_sync(0xffffffffu, bt, d);
            uint32_t cnt[PROD_WARPS], tot = 0;
#pragma unroll
            for (int w = 0; w < PROD_WARPS; w++) { cnt[w] = k < NBUCKET ? sm.cursor[w][k] : 0; tot += cnt[w]; }
            uint32_t inc = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
            uint32_t start = bt + inc - tot;
            if (k <= NBUCKET) { sm.bstart[buf][k] = (uint16_t)start; if (&cons != &sm) cons.bstart[buf][k] = (uint16_t)start; }
            if (k < NBUCKET) {
#pragma unroll
                for (int w = 0; w < PROD_WARPS; w++) { sm.cursor[w][k] = (uint16_t)start; start += cnt[w]; }
                if (k < NCTX) atomicAdd(&sm.cls_cnt[32 - __clz((int)tot)], 1);   // size class = bit length of the slice size
            }
        }
        JPROF(10, ptid == 0);
        JLS_NAMED_BARRIER(1, PROD_THREADS);
        for (int r = ra; r <= rb; r++) {
            const int c0 = max(ua - r * chunks, 0) * 32, c1 = min(W, min(ub - r * chunks, chunks) * 32);
            scatter_segment(sm, sorted_buf(cons, buf), W, r, c0, c1, lane, h);
        }
        JPROF(11, ptid == 0);
        // chain thread i of the tile takes the i-th biggest slice (by size class): similar slices share a warp; the run
        // events have a thread of their own in the last chain warp (next to the smallest slices)
        {
            int below = lane < 16 ? sm.cls_cnt[lane] : 0;                   // lane c: number of slices of class c
            int pos = below;                                                // -> number of slices of a bigger class
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) { const int v = __shfl_down_sync(0xffffffffu, pos, d); if (lane + d < 16) pos += v; }
            pos -= below;
            for (int kb = pw; kb < NBLK; kb += PROD_WARPS) {
                const int k = kb * 32 + lane;
                int cls = 0;
                if (k < NCTX) cls = 32 - __clz((int)sm.bstart[buf][k + 1] - (int)sm.bstart[buf][k]);
                const int first = __shfl_sync(0xffffffffu, pos, cls);
                if (k < NCTX) cons.order[buf][first + atomicAdd(&sm.cls_pos[cls], 1)] = (uint16_t)k;
            }
        }
    }
    // rows of the next tile
    if (t + 1 < ntiles) {
        if (bulk) {
            stage_bulk_edges(sm.pix[(t + 1) % 3], img, W, rows_per_tile, t + 1, ptid, PROD_THREADS);
            bulk_wait(&sm.stage_bar, bulk_phase & 1u);                  // the barrier's phase flips with every copy of this CTA
            bulk_phase++;
        }
        else if (SPLIT_STAGE) stage_commit<NREG>(sregs, sm.pix[(t + 1) % 3], img, W, H, rows_per_tile, t + 1, ptid, PROD_THREADS);
        else stage_tile(sm.pix[(t + 1) % 3], img, W, H, rows_per_tile, t + 1, ptid, PROD_THREADS);
    }
}

// ---- chains ---------------------------------------------------------------------------------------
struct RegularState { int A, B, C, N; };
struct RunState { int run_index, A0, B0, N0, A1, B1, N1, pending; };

// The serial part of a regular sample (:346-393): prediction correction, error, context update.  What the code needs
// (error, Golomb parameter, mapping flag) replaces the record; the code itself is generated later by all threads.
//
// The loop-carried dependency is C -> prediction -> error -> B -> bias test -> C; everything else is kept off it:
// with s = +-1 the sample's sign, sign * (Px - MED) = clamp(C, lo, hi) where lo/hi depend on MED only, so the error is
// sext8(sign * (x - MED) - clamp(C, lo, hi)); N and the halving flag form their own short chain; k and the mapping
// flag are outputs.  No branches: a warp's chain threads run different contexts in lock step.
__device__ __forceinline__ uint32_t regular_step(uint32_t rec, RegularState &s)
{
    const int d0 = (int)((rec >> 13) & 0x1FFu);         // sign * (x - MED) + 256: the bias vanishes in the modulo reduction
    const int nlo = (int)((rec >> 22) & 0xFFu);         // sign * (Px - MED) = clamp(C, lo, hi), lo = -nlo, hi = 255 - nlo  (:349 + the [0,255] clip)
    const int dlo = d0 + nlo, dhi = dlo - 255;          // d0 - lo, d0 - hi
    const int k = golomb_k(s.A, s.N);
    const int map = (k == 0) && (2 * s.B <= -s.N);      // :367
    const int rs = s.N >= 64;                           // :378-383: halve, then N++
    const int N1 = (s.N >> rs) + 1;
    const int e = (int)(int8_t)max(min(d0 - s.C, dlo), dhi);    // d0 - clamp(C, lo, hi), then modRange :105-111 with RANGE = 256
    const uint32_t out = (rec & 0x1FFFu) | ((uint32_t)(e & 0xFF) << 13) | ((uint32_t)k << 21) | ((uint32_t)map << 25);
    const int B1 = (s.B + e) >> rs;                     // :376-383
    s.A = (s.A + (e < 0 ? -e : e)) >> rs;
    s.N = N1;
    // bias update :384-393 without branches: dn = (B1 <= -N1), up = (B1 > 0) as sign masks
    const int sdn = (B1 + N1 - 1) >> 31, sup = (-B1) >> 31;
    const int dC = sdn - sup;                           // -1 / 0 / +1
    s.B = min(max(B1 - N1 * dC, 1 - N1), 0);
    s.C = min(max(s.C + dC, -128), 127);
    return out;
}

// Golomb code of a regular sample from what its chain left in the record (:364-373)
__device__ __forceinline__ void regular_code(uint32_t w, uint32_t *codes, uint8_t *lens)
{
    const int lp = w & 0x1FFF, e = (int)(int8_t)(w >> 13), k = (w >> 21) & 15, map = (w >> 25) & 1;
    const int ae = e < 0 ? -e : e;
    const int me = 2 * ae + (e < 0 ? -(map + 1) : map);
    uint32_t code; int len;
    golomb_code(me, k, 23, 8, code, len);
    codes[code_idx(lp)] = code;
    lens[len_idx(lp)]  = (uint8_t)len;
}

// one run event (:291-344): completed 2^J blocks -> ones in the slot of the run's last sample; interruption sample:
// run remainder in J+1 bits + Golomb code of the interruption error
// J[] (:15) from two packed constants, four bits per entry
__device__ __forceinline__ int run_order(int i)
{
    const unsigned long long t = i < 16 ? 0x3333222211110000ull : 0xFEDCBA9877665544ull;
    return (int)((t >> ((i & 15) * 4)) & 15ull);
}
__device__ __forceinline__ void run_step(uint32_t rec, RunState &s, const uint8_t *pb, int W, uint32_t *codes, uint8_t *lens)
{
    const int lp = rec & 0x1FFF;
    int L = (rec >> 13) & 0x7FFF;
    const bool eol = (rec >> 28) & 1u;
    if ((rec >> 29) & 1u) { s.pending += L; return; }   // the run goes on in the next segment of the row
    L += s.pending;
    s.pending = 0;
    int cnt = L, ones = 0;
    int jj = run_order(s.run_index);
    while (cnt >= (1 << jj)) {                          // :295-300, whole blocks at once
        cnt -= 1 << jj;
        ones++;
        if (s.run_index < 31) { s.run_index++; jj = run_order(s.run_index); }
    }
    if (eol) {                                          // :302-303
        if (cnt > 0) ones++;
        codes[code_idx(lp)] = (uint32_t)((1ull << ones) - 1ull);
        lens[len_idx(lp)]  = (uint8_t)ones;
        return;
    }
    if (ones > 0) {
        codes[code_idx(lp - 1)] = (uint32_t)((1ull << ones) - 1ull);
        lens[len_idx(lp - 1)]  = (uint8_t)ones;
    }
    const int glimit = 23 - 1 - jj;                     // :306
    if (s.run_index > 0) s.run_index--;                 // :310-311
    const uint8_t *px_ = pb + 16 + W + lp;              // the sample; the row above the picture is staged as zeros
    const int x = px_[0];
    const int b = px_[-W];
    const int a = ((rec >> 30) & 1u) ? b : (int)px_[-1];    // first column: a = b (:55)
    const int t = (a == b);                             // RItype :313
    const bool neg = a > b;                             // :314
    const int px = t ? a : b;
    int e = x - px;
    if (neg) e = -e;
    e = (int)(int8_t)e;
    const int At = t ? s.A1 : s.A0, Bt = t ? s.B1 : s.B0, Nt = t ? s.N1 : s.N0;
    const int k = golomb_k(At + (t ? (Nt >> 1) : 0), Nt);               // :329
    const int map = (e != 0) && ((e > 0) == (k == 0 && 2 * Bt < Nt));   // :331
    const int me = 2 * (e < 0 ? -e : e) - t - map;
    uint32_t gcode; int glen;
    golomb_code(me, k, glimit, 8, gcode, glen);
    codes[code_idx(lp)] = ((uint32_t)cnt << glen) | gcode;        // run remainder in J+1 bits, then the Golomb code (<= 32 bits)
    lens[len_idx(lp)]  = (uint8_t)(jj + 1 + glen);
    int An = At + ((me + 1 - t) >> 1), Bn = Bt + (e < 0), Nn = Nt;      // :336-344
    if (Nn >= 64) { An >>= 1; Bn >>= 1; Nn >>= 1; }
    Nn++;
    if (t) { s.A1 = An; s.B1 = Bn; s.N1 = Nn; } else { s.A0 = An; s.B0 = Bn; s.N0 = Nn; }
}

// ---- packer ---------------------------------------------------------------------------------------
// PACK_WARPS warps share a tile: (1) code-length totals per chunk of 1024 samples, (2) every chunk's codes are
// concatenated MSB-first at its bit offset in the tile-wide buffer ubits[], (3) one warp turns the bits into stuffed
// bytes.  Bits that do not fill a byte at the end of a tile wait in a register (`carry`) and read as word -1.
struct PackState {
    uint8_t *dst; size_t cap; size_t out_pos;
    uint32_t carry_bits;        // un-stuffed bits of earlier tiles still to be written (< 8)
    uint32_t carry;             // ... right-aligned
    uint32_t pz;                // 1: the next byte starts with a stuffed 0 bit (previous byte was 0xFF)
    // Output straight into page-locked host memory (imcvtJLSHostMode 2): byte stores would cross PCIe one by one, so the
    // stuffed bytes collect in a 256-byte shared-memory ring and leave as whole, 128-byte aligned lines (one 4-byte store
    // per lane = one write transaction per line).  ring == nullptr: bytes go straight to dst.
    uint8_t *ring;
    uint32_t mis;               // dst & 127: ring / line positions are counted from the 128-byte line dst lies in
    uint32_t flushed;           // first position (from that line's start) not yet written to dst; a multiple of 128
};
__device__ __forceinline__ void pack_state_init(PackState &ps, uint8_t *dst, size_t cap, uint8_t *ring)
{
    ps.out_pos = 0; ps.carry_bits = 0; ps.carry = 0; ps.pz = 0; ps.dst = dst; ps.cap = cap;
    ps.ring = ring; ps.mis = (uint32_t)((size_t)dst & 127u); ps.flushed = 0;
}
// lines that are complete below `vend` leave the ring (the first one may start inside a line: bytes)
__device__ __forceinline__ void ring_flush_lines(PackState &ps, uint32_t vend, int lane)
{
    while (ps.flushed + 128u <= vend) {
        uint8_t *g = ps.dst - ps.mis + ps.flushed;                      // 128-byte aligned
        const uint8_t *r = ps.ring + (ps.flushed & 255u);
        if (ps.flushed >= ps.mis) {
            reinterpret_cast<uint32_t *>(g)[lane] = reinterpret_cast<const uint32_t *>(r)[lane];
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) { const uint32_t v = 4u * lane + k; if (v >= ps.mis) g[v] = r[v]; }
        }
        ps.flushed += 128u;
    }
}
// the bytes of the last, incomplete line
__device__ __forceinline__ void ring_flush_tail(PackState &ps, uint32_t vend, int lane)
{
    for (uint32_t v = ps.flushed + lane; v < vend; v += 32) if (v >= ps.mis) ps.dst[v - ps.mis] = ps.ring[v & 255u];
    ps.flushed = vend;
}

__device__ __forceinline__ uint32_t ubits_word(const uint32_t *buf, int32_t w, uint32_t carry)
{
    return w >= 0 ? (w < UBITS_WORDS ? buf[w] : 0u) : (w == -1 ? carry : 0u);
}

// Turns the carried bits followed by the first `T` bits of ubits[] into stuffed bytes (writeBit :156-169).  Only whole
// bytes are produced unless `flush`; the left-over bits become the new carry.  One warp.
__device__ __forceinline__ void stuff_bits(const uint32_t *buf, PackState &ps, uint32_t T, bool flush, int lane)
{
    int32_t in_pos = -(int32_t)ps.carry_bits;
    const uint32_t carry = ps.carry;
    for (;;) {
        const int32_t s = in_pos - (int32_t)ps.pz;      // stream bit that lands on the MSB of this window's first byte
        const int32_t avail = (int32_t)T - s;           // bits from s to the end
        uint32_t whole = avail > 0 ? (uint32_t)avail >> 3 : 0u;
        if (whole == 0) break;
        const int32_t w0 = s >> 5;                      // >= -1
        const uint32_t sh = (uint32_t)s & 31u;
        const int32_t wi = w0 + lane;
        const uint32_t hi = ubits_word(buf, wi, carry), lo = ubits_word(buf, wi + 1, carry);
        uint32_t v = __funnelshift_l(lo, hi, sh);
        if (ps.pz && lane == 0) v &= 0x7FFFFFFFu;
        int fb = 4;
        if      ((v >> 24) == 0xFFu)           fb = 0;
        else if (((v >> 16) & 0xFFu) == 0xFFu) fb = 1;
        else if (((v >> 8) & 0xFFu) == 0xFFu)  fb = 2;
        else if ((v & 0xFFu) == 0xFFu)         fb = 3;
        // bytes past `whole` are not complete yet and must not end the window
        const uint32_t ball = __ballot_sync(0xffffffffu, fb < 4 && (uint32_t)(4 * lane + fb) < whole);
        uint32_t nbytes = whole < 128u ? whole : 128u;
        bool ends_ff = false;
        if (ball) {
            const int f = __ffs((int)ball) - 1;
            const uint32_t cut = 4u * f + (uint32_t)__shfl_sync(0xffffffffu, fb, f) + 1u;
            if (cut <= nbytes) { nbytes = cut; ends_ff = true; }
        }
        if (ps.ring) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t bi = 4u * lane + k;
                if (bi < nbytes && ps.out_pos + bi < ps.cap) ps.ring[(ps.mis + (uint32_t)ps.out_pos + bi) & 255u] = (uint8_t)(v >> (24 - 8 * k));
            }
            __syncwarp();
            const size_t end = ps.out_pos + nbytes < ps.cap ? ps.out_pos + nbytes : ps.cap;
            ring_flush_lines(ps, ps.mis + (uint32_t)end, lane);
            __syncwarp();
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t bi = 4u * lane + k;
                if (bi < nbytes && ps.out_pos + bi < ps.cap) ps.dst[ps.out_pos + bi] = (uint8_t)(v >> (24 - 8 * k));
            }
        }
        ps.out_pos += nbytes;
        in_pos = s + (int32_t)(8u * nbytes);
        ps.pz = ends_ff ? 1u : 0u;
    }
    const uint32_t left = (uint32_t)((int32_t)T - in_pos);      // < 8 (no whole byte can be formed)
    uint32_t rest = 0;                                          // the left-over bits, left-aligned
    if (left > 0) {
        const int32_t w = in_pos >> 5;
        rest = __funnelshift_l(ubits_word(buf, w + 1, carry), ubits_word(buf, w, carry), (uint32_t)in_pos & 31u) & ~(0xFFFFFFFFu >> left);
    }
    if (flush) {
        if (ps.ring) {                                  // what is left of the last line; the one or two closing bytes go direct
            const size_t end = ps.out_pos < ps.cap ? ps.out_pos : ps.cap;
            __syncwarp();
            ring_flush_tail(ps, ps.mis + (uint32_t)end, lane);
        }
        if (left > 0) {                                 // flushBits :177-184: pad the last byte with zeros
            if (lane == 0 && ps.out_pos < ps.cap) ps.dst[ps.out_pos] = (uint8_t)(ps.pz ? (rest >> 25) : (rest >> 24));
            ps.out_pos++; ps.pz = 0;
        }
        if (ps.pz) { if (lane == 0 && ps.out_pos < ps.cap) ps.dst[ps.out_pos] = 0; ps.out_pos++; ps.pz = 0; }   // a 0x00 after a trailing 0xFF (:178)
        ps.carry_bits = 0; ps.carry = 0;
        return;
    }
    ps.carry_bits = left;
    ps.carry = left ? rest >> (32 - left) : 0u;
}

// the packer warps' share of one tick: tile samples [0, n) (codes/lens in raster order) -> stuffed bytes
__device__ __forceinline__ void packer_tile(Smem &sm, const Params &P, PackState &ps, uint32_t *codes, uint8_t *lens, int n, bool last_tile, int pw, int lane)
{
    constexpr int MAXC = (TSMAX / CHUNK + PACK_WARPS - 1) / PACK_WARPS;     // chunks per warp
    uint32_t *buf = sm.ubits;
    const int nchunk = (n + CHUNK - 1) / CHUNK;
    JPROF_T0();
    // (1) lane owns 32 consecutive samples of each of its warp's chunks
    uint32_t l4[MAXC][8], mylen[MAXC];
#pragma unroll
    for (int m = 0; m < MAXC; m++) {
        const int c = pw + m * PACK_WARPS;
        const int s0 = c * CHUNK + lane * 32;
        mylen[m] = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint32_t v = 0;
            if (c < nchunk && s0 + 4 * q < n) v = *reinterpret_cast<const uint32_t *>(lens + len_idx(s0) + 4 * q);   // slots past n hold zeros
            l4[m][q] = v;
            mylen[m] += (v & 0xFF) + ((v >> 8) & 0xFF) + ((v >> 16) & 0xFF) + (v >> 24);
        }
        uint32_t tot = mylen[m];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
        if (lane == 0 && c < nchunk) sm.chunk_bits[c] = tot;
    }
    JLS_NAMED_BARRIER(3, PACK_WARPS * 32);
    // chunk offsets; zero the words the tile's bits will be OR-ed into
    uint32_t cb = lane < nchunk ? sm.chunk_bits[lane] : 0u, cinc = cb;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, cinc, d); if (lane >= d) cinc += v; }
    const uint32_t T = __shfl_sync(0xffffffffu, cinc, 7);                   // TSMAX / CHUNK <= 8 chunks
    const uint32_t cex = cinc - cb;
    for (uint32_t w = pw * 32 + lane; w < (T >> 5) + 2; w += PACK_WARPS * 32) buf[w] = 0;
    JLS_NAMED_BARRIER(3, PACK_WARPS * 32);
    JPROF(12, pw == 0 && lane == 0);
    // (2) the codes of a lane's samples are concatenated in a register window and leave as whole words
#pragma unroll
    for (int m = 0; m < MAXC; m++) {
        const int c = pw + m * PACK_WARPS;
        const uint32_t coff = __shfl_sync(0xffffffffu, cex, c & 7);
        uint32_t inc = mylen[m];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
        if (mylen[m]) {
            const int s0 = c * CHUNK + lane * 32;
            const uint32_t bit = coff + inc - mylen[m];
            uint32_t w = bit >> 5;
            uint32_t nacc = bit & 31u;                  // bits in the window (the first `bit & 31` belong to a neighbour: zeros here)
            uint64_t acc = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                uint32_t cd[4];
#pragma unroll
                for (int r = 0; r < 4; r++) cd[r] = codes[code_idx(s0) + 4 * q + r];     // stale codes of empty slots are masked by len = 0
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint32_t len = (l4[m][q] >> (8 * r)) & 0xFF;
                    acc = (acc << len) | (len ? (uint64_t)cd[r] : 0ull);
                    nacc += len;
                    if (nacc >= 32) {
                        nacc -= 32;
                        atomicOr(&buf[w], (uint32_t)(acc >> nacc));
                        w++;
                    }
                }
                // the slots are consumed: the chains of a later tile expect zero lengths
                if (l4[m][q]) *reinterpret_cast<uint32_t *>(lens + len_idx(s0) + 4 * q) = 0;
            }
            if (nacc) atomicOr(&buf[w], (uint32_t)(acc << (32 - nacc)));
        }
    }
    JLS_NAMED_BARRIER(3, PACK_WARPS * 32);
    JPROF(13, pw == 0 && lane == 0);
    // (3)
    if (pw == 0) stuff_bits(buf, ps, T, last_tile, lane);
    JPROF(14, pw == 0 && lane == 0);
}

// ---- the kernels ----------------------------------------------------------------------------------
// chain warps' share of one tick: tile t (sorted by the producers one tick earlier)
__device__ __forceinline__ void chain_tick(Smem &sm, const Params &P, int t, int W, int tid, long long jp_t0)
{
    const int buf = t & 1;
    uint32_t *rec = sm.sorted[buf];
    const int bucket = tid < NCTX ? (int)sm.order[buf][tid] : (tid == CHAIN_WARPS * 32 - 1 ? NCTX : -1);
    if (bucket >= 0) {
        const int b0 = sm.bstart[buf][bucket], b1 = sm.bstart[buf][bucket + 1];
        if (bucket < NCTX) {
            if (b1 > b0) {
                RegularState rs; rs.A = sm.ctx[bucket][0]; rs.B = sm.ctx[bucket][1]; rs.C = sm.ctx[bucket][2]; rs.N = sm.ctx[bucket][3];
                uint32_t r = rec[b0];
                for (int i = b0; i < b1; i++) {
                    const uint32_t cur = r;
                    if (i + 1 < b1) r = rec[i + 1];
                    rec[i] = regular_step(cur, rs);
                }
                sm.ctx[bucket][0] = rs.A; sm.ctx[bucket][1] = rs.B; sm.ctx[bucket][2] = rs.C; sm.ctx[bucket][3] = rs.N;
            }
        } else if (b1 > b0) {
            RunState us;
            us.run_index = sm.run_state[0]; us.A0 = sm.run_state[1]; us.B0 = sm.run_state[2]; us.N0 = sm.run_state[3];
            us.A1 = sm.run_state[4]; us.B1 = sm.run_state[5]; us.N1 = sm.run_state[6]; us.pending = sm.run_state[7];
            for (int i = b0; i < b1; i++) run_step(rec[i], us, sm.pix[t % 3], W, sm.codes[buf], sm.lens[buf]);
            sm.run_state[0] = us.run_index; sm.run_state[1] = us.A0; sm.run_state[2] = us.B0; sm.run_state[3] = us.N0;
            sm.run_state[4] = us.A1; sm.run_state[5] = us.B1; sm.run_state[6] = us.N1; sm.run_state[7] = us.pending;
        }
    }
    JPROF(0, tid == 0);
    JPROF(6, tid == CHAIN_WARPS * 32 - 1);
    JLS_NAMED_BARRIER(2, CHAIN_WARPS * 32);
    JPROF(7, tid == 0);
    // all chain threads: Golomb codes of the tile's regular samples into their raster slots
    const int nreg = sm.bstart[buf][NCTX];
    for (int i = tid; i < nreg; i += CHAIN_WARPS * 32) regular_code(rec[i], sm.codes[buf], sm.lens[buf]);
    JPROF(4, tid == 0);
}

// ---- pair kernel: the chain in two stages ----------------------------------------------------------
// A warp issues one integer instruction every other cycle, so the tick of a tile is (samples of its biggest context) x
// (instructions of a chain step) x 2 cycles.  The pair kernel therefore keeps only the true recurrence on the critical
// thread -- stage 1: C -> error -> B -> bias, plus N -- and leaves A, the Golomb parameter, the mapping and the code to
// stage 2, which walks the same slice one tick later on other warps.
struct C1State { int B, C, N; };
__device__ __forceinline__ uint32_t chain1_step(uint32_t rec, C1State &s)
{
    const int d0 = (int)((rec >> 13) & 0x1FFu);         // sign * (x - MED) + 256
    const int dlo = d0 + (int)(rec >> 22), dhi = dlo - 255;     // d0 - lo, d0 - hi of the clamp window (regular_step)
    const int rs = s.N >> 6;                            // N <= 64: halve when it is 64 (:378-383)
    const int N1 = (s.N >> rs) + 1;
    const uint32_t cond = (uint32_t)(2 * s.B + s.N - 1) >> 31;  // 2B <= -N: the mapping of :367 applies when k turns out 0
    const int e = (int)(int8_t)max(min(d0 - s.C, dlo), dhi);
    const uint32_t out = (rec & 0x1FFFu) | ((uint32_t)(e & 0xFF) << 13) | (cond << 21) | ((uint32_t)s.N << 22);
    const int B1 = (s.B + e) >> rs;
    s.N = N1;
    const int sdn = (B1 + N1 - 1) >> 31, sup = (-B1) >> 31;
    const int dC = sdn - sup;
    s.B = min(max(B1 - N1 * dC, 1 - N1), 0);
    s.C = min(max(s.C + dC, -128), 127);
    return out;
}
__device__ __forceinline__ void chain1_tick(Smem &sm, int t, int tid)
{
    const int buf = t % 3;
    uint32_t *rec = sorted_buf(sm, buf);
    if (tid < NCTX) {
        const int bucket = sm.order[buf][tid];
        const int b0 = sm.bstart[buf][bucket], b1 = sm.bstart[buf][bucket + 1];
        if (b1 > b0) {
            C1State cs; cs.B = sm.ctx[bucket][1]; cs.C = sm.ctx[bucket][2]; cs.N = sm.ctx[bucket][3];
            uint32_t r = rec[b0];
            for (int i = b0; i < b1; i++) {
                const uint32_t cur = r;
                if (i + 1 < b1) r = rec[i + 1];
                rec[i] = chain1_step(cur, cs);
            }
            sm.ctx[bucket][1] = cs.B; sm.ctx[bucket][2] = cs.C; sm.ctx[bucket][3] = cs.N;
        }
    }
}
// stage 2, CHAIN2_WARPS warps: thread j walks the j-th biggest and the j-th smallest slice of the tile (A, k, map, Golomb
// code into the raster slots); the last thread walks the run events
constexpr int CHAIN2_WARPS = 6;
__device__ __forceinline__ void chain2_slice(Smem &sm, uint32_t *rec, int bucket, int buf)
{
    const int b0 = sm.bstart[buf][bucket], b1 = sm.bstart[buf][bucket + 1];
    if (b1 <= b0) return;
    int A = sm.ctx_a[bucket];
    uint32_t w = rec[b0];
    for (int i = b0; i < b1; i++) {
        const uint32_t cur = w;
        if (i + 1 < b1) w = rec[i + 1];
        const int e = (int)(int8_t)(cur >> 13), N = (int)((cur >> 22) & 0x7Fu);
        const int k = golomb_k(A, N);
        const uint32_t map = (k == 0) ? (cur >> 21) & 1u : 0u;
        rec[i] = (cur & 0x1FFFFFu) | ((uint32_t)k << 21) | (map << 25);     // what regular_code reads: pos | error | k | map
        A = (A + (e < 0 ? -e : e)) >> (N >> 6);
    }
    sm.ctx_a[bucket] = A;
}
__device__ __forceinline__ void chain2_tick(Smem &sm, int t, int W, int j)
{
    const int buf = t % 3, cb = t & 1;
    uint32_t *rec = sorted_buf(sm, buf);
    // which stage-2 warp gets the biggest slices: not the one that shares an SM sub-partition (hardware warp id mod 4) with
    // stage 1's warp of biggest slices -- both are bound by the integer issue rate of their sub-partition
    int jr = (j + CHAIN2_WARPS * 32 - 32) % (CHAIN2_WARPS * 32);
    // ... and the run events (index 191) go to a lane of yet another sub-partition (thread 127 swaps with thread 31)
    if (j == 127) jr = CHAIN2_WARPS * 32 - 1; else if (j == 31) jr = 95;
    if (jr < NCTX / 2) {
        chain2_slice(sm, rec, sm.order[buf][jr], buf);
        chain2_slice(sm, rec, sm.order[buf][NCTX - 1 - jr], buf);
    } else if (jr == CHAIN2_WARPS * 32 - 1) {
        const int b0 = sm.bstart[buf][NCTX], b1 = sm.bstart[buf][NCTX + 1];
        if (b1 > b0) {
            RunState us;
            us.run_index = sm.run_state[0]; us.A0 = sm.run_state[1]; us.B0 = sm.run_state[2]; us.N0 = sm.run_state[3];
            us.A1 = sm.run_state[4]; us.B1 = sm.run_state[5]; us.N1 = sm.run_state[6]; us.pending = sm.run_state[7];
            for (int i = b0; i < b1; i++) run_step(rec[i], us, sm.pix[t % 3], W, sm.codes[cb], sm.lens[cb]);
            sm.run_state[0] = us.run_index; sm.run_state[1] = us.A0; sm.run_state[2] = us.B0; sm.run_state[3] = us.N0;
            sm.run_state[4] = us.A1; sm.run_state[5] = us.B1; sm.run_state[6] = us.N1; sm.run_state[7] = us.pending;
        }
    }
    JLS_NAMED_BARRIER(5, CHAIN2_WARPS * 32);
    // all stage-2 threads: Golomb codes of the tile's regular samples into their raster slots
    const int nreg = sm.bstart[buf][NCTX];
    for (int i = j; i < nreg; i += CHAIN2_WARPS * 32) regular_code(rec[i], sm.codes[cb], sm.lens[cb]);
}

// PAIR = false: one CTA per plane, its warps split into chains | producers | packers.
// PAIR = true : a cluster of two CTAs per plane.  Rank 0: every warp is a producer and scatters the sorted tile into
//               the shared memory of rank 1.  Rank 1: chains | packers | the remaining warps stage the rows the run
//               events read.  One cluster barrier per tick replaces __syncthreads.
template <bool PAIR>
__device__ __forceinline__ void tile_body(const Params &P)
{
    const int rank = PAIR ? JLS_PAIR_RANK() : 0;
    // the warp arbiter favours high warp ids: the chain warps (role index 0.., the biggest slices first) sit on the
    // highest hardware warps so that the longest context chain of a tile is never kept waiting
    const int hw_tid = PAIR ? JLS_PAIR_TID() : (int)threadIdx.x;
    const int lane = hw_tid & 31, warp = NWARPS - 1 - (hw_tid >> 5);
    const int tid = warp * 32 + lane;
    Smem &sm   = *reinterpret_cast<Smem *>(PAIR ? (unsigned char *)JLS_OWN_SMEM(rank) : (unsigned char *)JLS_DYN_SMEM());
    // where the producers put the sorted tile: the same image, or rank 1's through distributed shared memory
    Smem &cons = *reinterpret_cast<Smem *>(PAIR ? (unsigned char *)JLS_PAIR_SMEM(1) : (unsigned char *)JLS_DYN_SMEM());
    // the producer CTA's image as the consumer CTA sees it (rows of a tile are pulled from there)
    Smem &prod = *reinterpret_cast<Smem *>(PAIR ? (unsigned char *)JLS_PAIR_SMEM(0) : (unsigned char *)JLS_DYN_SMEM());
    static_assert(offsetof(Smem, pix) % 4 == 0 && PIXBUF % 4 == 0, "pull_tile copies words");
    const int W = P.W, H = P.H;
    const int rows_per_tile = max(1, TSMAX / W);
    const int ntiles = (H + rows_per_tile - 1) / rows_per_tile;
    const bool producer_cta = !PAIR || rank == 0, consumer_cta = !PAIR || rank == 1;
    constexpr int CHAIN2_0 = CHAIN_WARPS;                                         // pair, rank 1: first warp of chain stage 2
    constexpr int PACK0 = PAIR ? CHAIN_WARPS + CHAIN2_WARPS : CHAIN_WARPS + PROD_WARPS;     // first packer warp
    constexpr int STAGE0 = PACK0 + PACK_WARPS;                                    // pair, rank 1: first staging warp
    static_assert(!PAIR || STAGE0 < NWARPS, "the consumer CTA needs at least one staging warp");
    constexpr int NTICK_EXTRA = PAIR ? 3 : 2;                                     // pipeline depth behind the producers

    if (consumer_cta) for (int i = tid; i < 2 * LEN_SLOTS / 4; i += NTHREADS) reinterpret_cast<uint32_t *>(&sm.lens[0][0])[i] = 0;
    if (producer_cta) build_qlut(sm.qlut, tid, NTHREADS);
    uint32_t bulk_phase = 0;                        // bulk copies this CTA has awaited (every staging thread counts alike)
    if (producer_cta && tid == 0) bulk_bar_init(&sm.stage_bar);     // in front of the first use: the barrier at the top of the plane loop

    for (;;) {
        if (PAIR) JLS_PAIR_SYNC(); else __syncthreads();
        if (rank == 0 && tid == 0) { const int q = atomicAdd(P.next_plane, 1); sm.plane = q; if (PAIR) cons.plane = q; }
        if (PAIR) JLS_PAIR_SYNC(); else __syncthreads();
        const int p = sm.plane;
        if (p >= P.P) break;
        const uint8_t *img = P.pix + (size_t)p * P.plane_stride;

        PackState ps;
        pack_state_init(ps, nullptr, 0, nullptr);
        if (consumer_cta) {
            // :256-261 (a_init = 4 for RANGE 256), :250, :263-267
            for (int i = tid; i < NCTX; i += NTHREADS) { sm.ctx[i][0] = 4; sm.ctx[i][1] = 0; sm.ctx[i][2] = 0; sm.ctx[i][3] = 1; sm.ctx_a[i] = 4; }
            if (tid < 8) sm.run_state[tid] = (tid == 1 || tid == 4) ? 4 : (tid == 3 || tid == 6) ? 1 : 0;
            size_t cap;
            uint8_t *dst = plane_dst(P.L, p, cap);
            pack_state_init(ps, dst, cap, P.line_out ? sm.out_ring : nullptr);
        }
        if (producer_cta) stage_tile(sm.pix[0], img, W, H, rows_per_tile, 0, tid, NTHREADS);
        if (PAIR) JLS_PAIR_SYNC(); else __syncthreads();

        for (int tick = 0; tick < ntiles + NTICK_EXTRA; tick++) {
            JPROF_T0();
#if !(defined(IMCVT_JLS_PROF) && !defined(IMCVT_HOST_EMU))
            const long long jp_t0 = 0;
#endif
            if (PAIR) {
                if (rank == 0) {
                    producer_tick<NWARPS>(sm, cons, P, img, rows_per_tile, tick, ntiles, 3, warp, lane, bulk_phase);
                    JPROF(1, tid == 0);
                } else if (warp < CHAIN_WARPS) {
                    if (tick >= 1 && tick - 1 < ntiles) chain1_tick(sm, tick - 1, tid);
                    JPROF(0, tid == 0);
                } else if (warp < PACK0) {
                    if (tick >= 2 && tick - 2 < ntiles) chain2_tick(sm, tick - 2, W, tid - CHAIN2_0 * 32);
                    JPROF(6, tid == CHAIN2_0 * 32);
                    JPROF(7, tid == PACK0 * 32 - 1);
                } else if (warp < STAGE0) {
                    const int t = tick - 3;
                    if (t >= 0 && t < ntiles) {
                        const int r0 = t * rows_per_tile, n = (min(H, r0 + rows_per_tile) - r0) * W;
                        packer_tile(sm, P, ps, sm.codes[t & 1], sm.lens[t & 1], n, t == ntiles - 1, warp - PACK0, lane);
                    }
                } else {
                    // the rows of tile `tick`, read by its run events two ticks later: the producer CTA staged them during the
                    // tick before and works on them now (it stages tile tick + 1 into another buffer meanwhile)
                    if (tick < ntiles) pull_tile(sm.pix[tick % 3], prod.pix[tick % 3], W, H, rows_per_tile, tick, tid - STAGE0 * 32, NTHREADS - STAGE0 * 32);
                }
                JPROF(2, rank == 1 && warp == PACK0 && lane == 0);
                JLS_PAIR_SYNC();
                JPROF(3, rank == 1 && tid == 0);
            } else {
                if (warp >= CHAIN_WARPS && warp < CHAIN_WARPS + PROD_WARPS) {
                    producer_tick<PROD_WARPS>(sm, sm, P, img, rows_per_tile, tick, ntiles, 2, warp - CHAIN_WARPS, lane, bulk_phase);
                    JPROF(1, warp == CHAIN_WARPS && lane == 0);
                    JPROF(5, warp == CHAIN_WARPS + PROD_WARPS - 1 && lane == 0);
                } else if (warp < CHAIN_WARPS) {
                    if (tick >= 1 && tick - 1 < ntiles) chain_tick(sm, P, tick - 1, W, tid, jp_t0);
                } else {
                    const int t = tick - 2;
                    if (t >= 0 && t < ntiles) {
                        const int r0 = t * rows_per_tile, n = (min(H, r0 + rows_per_tile) - r0) * W;
                        packer_tile(sm, P, ps, sm.codes[t & 1], sm.lens[t & 1], n, t == ntiles - 1, warp - PACK0, lane);
                    }
                    JPROF(2, warp == PACK0 && lane == 0);
                }
                __syncthreads();
                JPROF(3, tid == 0);
            }
        }
        if (consumer_cta && warp == PACK0 && lane == 0) P.plane_bytes[p] = ps.out_pos;
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) jls_tile_kernel(Params P) { tile_body<false>(P); }
__global__ void JLS_CLUSTER_DIMS __launch_bounds__(NTHREADS, 1) jls_pair_kernel(Params P) { tile_body<true>(P); }

}  // namespace tile
}  // namespace jls
}  // namespace imcvt
