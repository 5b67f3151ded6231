// jls.cu -- JPEG-LS (ITU-T T.87 baseline, 8-bit, NEAR 0..4) encoder for sm_100a.
//
// Replaces the codec part of the reference's writeJLSImageFile (src/imageio_jls.c:428-465):
// JLSencodeImageGray/RGB (:402-423) -> JLSencodeScan (:240-399) and the bit writer (:124-197).
// Streams are bit-identical to the reference.
//
// Two device paths:
//
//  * NEAR == 0 (lossless): the reconstruction equals the input, so the causal template, the context,
//    the MED prediction and the run/regular classification of every sample depend on the input only.
//    The scan is decomposed into
//       classify (pixel parallel, run state by a carry-lookahead over warp ballots)
//       -> stable multisplit of the samples into 364 context buckets + 1 bucket of run events
//       -> chains: one thread per (plane, context) walks its samples in raster order carrying A,B,C,N
//          (the only true recurrence, :376-393), one thread per plane walks the run events carrying
//          run_index and the two run-interruption contexts (:291-344)
//       -> raster-order prefix sum of code lengths -> MSB-first bit packing in shared memory
//       -> marker-avoidance bit stuffing (:156-169) + headers.
//
//  * NEAR > 0 (near-lossless): the reconstructed value feeds the next sample's context (:321-323), the
//    whole plane is one dependency chain.  One warp per plane, lane 0 walks the scan.
//    The same kernel can be forced for NEAR == 0 as an on-device cross-check.
#include "common.cuh"
#include "../../include/imcvt_b200.h"

#include "jls_tile_kernel.cuh"
#include "jls_rowscan_kernel.cuh"
#include "jls_nearscan_kernel.cuh"

namespace imcvt {
namespace jls {

constexpr int ROWS_PER_TILE = 8;     // classify / multisplit tile = 8 image rows handled by one warp
constexpr int CLS_WARPS = 4;
constexpr int PK_THREADS = 256;      // bit packer: 256 threads x 8 samples
constexpr int PK_PER     = 8;
constexpr int PK_TILE    = PK_THREADS * PK_PER;
constexpr int ST_CHUNK_WORDS = 2048; // stuffing kernel: shared-memory window over the unstuffed stream

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// K0: RGB24 interleaved -> 3 planes (replaces the de-interleave loop :447-453)
// ------------------------------------------------------------------------------------------------
__global__ void deinterleave_rgb_kernel(const uint8_t *__restrict__ frames, size_t frame_stride, int n_frames,
                                        size_t npix, uint8_t *__restrict__ planes, size_t plane_stride)
{
    size_t total = (size_t)n_frames * npix;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t f = t / npix, i = t - f * npix;
        const uint8_t *src = frames + f * frame_stride + 3 * i;
        uint8_t *dst = planes + (3 * f) * plane_stride + i;
        dst[0] = src[0];
        dst[plane_stride] = src[1];
        dst[2 * plane_stride] = src[2];
    }
}

// ------------------------------------------------------------------------------------------------
// K1 / K3: classify every sample of a tile (8 rows, one warp) and either count it into the tile's
// bucket histogram (SCATTER = false) or write its record to the bucket's slice (SCATTER = true).
//
// Per sample (NEAR = 0): Q = 81 q(d-b) + 9 q(b-c) + q(c-a)  (getQ :79-84, neighbours per samplePixels :46-65),
//   q0 = (Q == 0), eq = (x == a).  Run state along a row (:270, :284-289):
//   running_j = (q0_j | running_{j-1}) & eq_j  -- a carry chain with generate q0&eq and propagate eq,
//   resolved for 32 samples at once by one 64-bit addition of the two ballot masks.
// Records:  regular sample  -> bucket |Q|-1 : pos | sign<<31 , aux = x | MED<<8
//           run interruption-> bucket 364   : pos             , aux = run length before it
//           run hits row end-> bucket 364   : pos | 1<<31     , aux = run length including it
// ------------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(CLS_WARPS * 32)
classify_kernel(const uint8_t *__restrict__ pix, size_t plane_stride, int W, int H, int tiles_per_plane, int total_tiles,
                uint32_t *__restrict__ hist, const uint32_t *__restrict__ bucket_start,
                uint32_t *__restrict__ rec_pos, uint16_t *__restrict__ rec_aux, size_t rec_stride,
                uint8_t *__restrict__ lens, size_t slot_stride)
{
    __shared__ uint32_t sh_hist[CLS_WARPS][HSLOTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gt = blockIdx.x * CLS_WARPS + warp;
    if (gt >= total_tiles) return;                      // whole warp leaves together; only __syncwarp below
    const int p = gt / tiles_per_plane, tile = gt - p * tiles_per_plane;
    uint32_t *h = sh_hist[warp];
    const size_t hrow = (size_t)p * NBUCKET;

    for (int k = lane; k < HSLOTS; k += 32) {
        uint32_t v = 0;
        if (SCATTER && k < NBUCKET) v = bucket_start[hrow + k] + hist[(hrow + k) * tiles_per_plane + tile];
        h[k] = v;
    }
    __syncwarp();

    const uint8_t *img = pix + (size_t)p * plane_stride;
    const int row0 = tile * ROWS_PER_TILE, row1 = min(H, row0 + ROWS_PER_TILE);
    const uint32_t lane_le = 0xFFFFFFFFu >> (31 - lane);        // bits 0..lane
    const uint32_t lane_lt = lane_le >> 1;                      // bits 0..lane-1

    for (int i = row0; i < row1; i++) {
        const uint8_t *cur = img + (size_t)i * W;
        const uint8_t *up  = cur - W;
        uint32_t cin = 0;
        int run_start = 0;
        for (int j0 = 0; j0 < W; j0 += 32) {
            const int j = j0 + lane;
            const bool active = j < W;
            int x = 0, a = 0, b = 0, c = 0, d = 0;
            if (active) {
                x = cur[j];
                if (i > 0) { b = up[j]; d = (j + 1 < W) ? up[j + 1] : b; }
                if (j == 0) { a = b; c = (i > 1) ? up[-W] : 0; }
                else        { a = cur[j - 1]; c = (i > 0) ? up[j - 1] : 0; }
            }
            const int Q = 81 * grad_q_lossless(d - b) + 9 * grad_q_lossless(b - c) + grad_q_lossless(c - a);
            const bool q0 = active && Q == 0;
            const bool eq = active && x == a;
            const uint32_t act = __ballot_sync(0xffffffffu, active);
            const uint32_t Pm  = __ballot_sync(0xffffffffu, eq);
            const uint32_t Gm  = __ballot_sync(0xffffffffu, q0 && eq);
            const uint32_t Q0m = __ballot_sync(0xffffffffu, q0);
            const uint64_t cv  = ((uint64_t)Pm + Gm + cin) ^ Pm ^ Gm;   // bit j = carry into j = running_{j-1}
            const uint32_t prevrun = (uint32_t)cv;
            const uint32_t R   = (uint32_t)(cv >> 1) & act;             // running_j
            const uint32_t INM = (Q0m | prevrun) & act;                 // sample is coded in run mode (hit or interruption)
            const uint32_t S   = Q0m & ~prevrun;                        // a run-mode stretch starts here
            const bool inm = (INM >> lane) & 1u, run = (R >> lane) & 1u;
            const bool is_int = inm && !run;
            const bool is_eol = run && (j == W - 1);

            int bucket = NBUCKET;                                       // dummy: plain run sample / inactive lane
            if (active) {
                if (!inm) bucket = abs(Q) - 1;
                else if (is_int || is_eol) bucket = NCTX;
            }
            const uint32_t peers = __match_any_sync(0xffffffffu, bucket);
            const int rank = __popc(peers & lane_lt);
            const uint32_t base = h[bucket];
            __syncwarp();
            if (rank == 0) h[bucket] = base + __popc(peers);
            __syncwarp();

            if (SCATTER) {
                const uint32_t pos = (uint32_t)i * (uint32_t)W + (uint32_t)j;
                if (bucket < NBUCKET) {
                    uint32_t flag; uint32_t aux;
                    if (bucket < NCTX) {
                        flag = Q < 0;
                        aux  = (uint32_t)x | ((uint32_t)med3(a, b, c) << 8);
                    } else {
                        const uint32_t sm = S & lane_le;
                        const int start_col = sm ? (j0 + 31 - __clz(sm)) : run_start;
                        flag = is_eol;
                        aux  = (uint32_t)(j - start_col + (is_eol ? 1 : 0));
                    }
                    const size_t o = (size_t)p * rec_stride + base + rank;
                    rec_pos[o] = pos | (flag << 31);
                    rec_aux[o] = (uint16_t)aux;
                }
                if (run) lens[(size_t)p * slot_stride + pos] = 0;       // run hits emit nothing by themselves
            }
            cin = (uint32_t)(cv >> 32) & 1u;
            if (S) run_start = j0 + 31 - __clz(S);
        }
    }

    if (!SCATTER) {
        __syncwarp();
        for (int k = lane; k < NBUCKET; k += 32) hist[(hrow + k) * tiles_per_plane + tile] = h[k];
    }
}

// K2a: exclusive scan over the tiles of every (plane, bucket) row, one warp per row.
__global__ void hist_scan_kernel(uint32_t *__restrict__ hist, uint32_t *__restrict__ bucket_total, int T, int nrows)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= nrows) return;
    uint32_t *h = hist + (size_t)row * T;
    uint32_t carry = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const uint32_t v = t < T ? h[t] : 0;
        const uint32_t inc = warp_incl_scan(v, lane);
        if (t < T) h[t] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) bucket_total[row] = carry;
}

// K2b: exclusive scan over the 365 bucket totals of a plane -> start of every bucket's record slice.
__global__ void bucket_offsets_kernel(const uint32_t *__restrict__ bucket_total, uint32_t *__restrict__ bucket_start, int P)
{
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= P) return;
    uint32_t carry = 0;
    for (int k0 = 0; k0 < NBUCKET; k0 += 32) {
        const int k = k0 + lane;
        const uint32_t v = k < NBUCKET ? bucket_total[(size_t)p * NBUCKET + k] : 0;
        const uint32_t inc = warp_incl_scan(v, lane);
        if (k < NBUCKET) bucket_start[(size_t)p * NBUCKET + k] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// ------------------------------------------------------------------------------------------------
// K4: the chains.  Blocks [0, reg_blocks) : one thread per (plane, regular context).
//                  Blocks [reg_blocks, ..): one warp (lane 0) per plane for the run events.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void regular_step(uint32_t rec, uint32_t aux, int &A, int &B, int &C, int &N,
                                             uint32_t *__restrict__ codes, uint8_t *__restrict__ lens)
{
    const int x = aux & 0xff, med = aux >> 8;
    const bool neg = rec >> 31;
    const uint32_t pos = rec & 0x7fffffffu;
    int px = med + (neg ? -C : C);                      // :349
    px = min(max(px, 0), 255);
    int e = x - px;
    if (neg) e = -e;
    e = (int)(int8_t)e;                                 // modRange :105-111 with RANGE = 256
    const int k = golomb_k(A, N);
    const int ae = abs(e);
    const int map = (k == 0) && (2 * B <= -N);          // :367
    const int me = 2 * ae + (e < 0 ? -(map + 1) : map);
    uint32_t code; int len;
    golomb_code(me, k, 23, 8, code, len);
    codes[pos] = code;
    lens[pos]  = (uint8_t)len;
    B += e;                                             // :376-393
    A += ae;
    if (N >= 64) { A >>= 1; B >>= 1; N >>= 1; }
    N++;
    if (B <= -N) {
        B += N; if (B <= -N) B = -N + 1;
        C = max(C - 1, -128);
    } else if (B > 0) {
        B -= N; if (B > 0) B = 0;
        C = min(C + 1, 127);
    }
}

__global__ void __launch_bounds__(128)
chain_kernel(const uint8_t *__restrict__ pix, size_t plane_stride, int W, int P, int reg_blocks,
             const uint32_t *__restrict__ bucket_start, const uint32_t *__restrict__ bucket_total,
             const uint32_t *__restrict__ rec_pos, const uint16_t *__restrict__ rec_aux, size_t rec_stride,
             uint32_t *__restrict__ codes_all, uint8_t *__restrict__ lens_all, size_t slot_stride)
{
    if ((int)blockIdx.x < reg_blocks) {
        // ---- regular contexts: A,B,C,N recurrence in registers, records streamed with a 4-deep prefetch ----
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= P * NCTX) return;
        const int p = t / NCTX, q = t - p * NCTX;
        const uint32_t n = bucket_total[(size_t)p * NBUCKET + q];
        const size_t o = (size_t)p * rec_stride + bucket_start[(size_t)p * NBUCKET + q];
        const uint32_t *rp = rec_pos + o;
        const uint16_t *ra = rec_aux + o;
        uint32_t *codes = codes_all + (size_t)p * slot_stride;
        uint8_t  *lens  = lens_all + (size_t)p * slot_stride;
        int A = 4, B = 0, C = 0, N = 1;                 // :256-261 (a_init = 4 for RANGE 256)
        uint32_t i = 0;
        uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (n >= 4) { r0 = rp[0]; r1 = rp[1]; r2 = rp[2]; r3 = rp[3]; a0 = ra[0]; a1 = ra[1]; a2 = ra[2]; a3 = ra[3]; }
        for (; i + 4 <= n; i += 4) {
            const uint32_t c0 = r0, c1 = r1, c2 = r2, c3 = r3, b0 = a0, b1 = a1, b2 = a2, b3 = a3;
            if (i + 8 <= n) {
                r0 = rp[i + 4]; r1 = rp[i + 5]; r2 = rp[i + 6]; r3 = rp[i + 7];
                a0 = ra[i + 4]; a1 = ra[i + 5]; a2 = ra[i + 6]; a3 = ra[i + 7];
            }
            regular_step(c0, b0, A, B, C, N, codes, lens);
            regular_step(c1, b1, A, B, C, N, codes, lens);
            regular_step(c2, b2, A, B, C, N, codes, lens);
            regular_step(c3, b3, A, B, C, N, codes, lens);
        }
        for (; i < n; i++) regular_step(rp[i], ra[i], A, B, C, N, codes, lens);
        return;
    }

    // ---- run events of one plane (:291-344): run_index and the two run-interruption contexts ----
    const int p = ((int)blockIdx.x - reg_blocks) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P || (threadIdx.x & 31) != 0) return;
    const uint32_t n = bucket_total[(size_t)p * NBUCKET + NCTX];
    const size_t o = (size_t)p * rec_stride + bucket_start[(size_t)p * NBUCKET + NCTX];
    const uint32_t *rp = rec_pos + o;
    const uint16_t *ra = rec_aux + o;
    const uint8_t *img = pix + (size_t)p * plane_stride;
    uint32_t *codes = codes_all + (size_t)p * slot_stride;
    uint8_t  *lens  = lens_all + (size_t)p * slot_stride;
    int run_index = 0;
    int A0 = 4, B0 = 0, N0 = 1, A1 = 4, B1 = 0, N1 = 1;                 // :263-267
    if (n == 0) return;
    uint32_t nrec = rp[0], naux = ra[0];
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t rec = nrec, L = naux;
        if (i + 1 < n) { nrec = rp[i + 1]; naux = ra[i + 1]; }
        const bool eol = rec >> 31;
        const uint32_t pos = rec & 0x7fffffffu;
        int cnt = (int)L, ones = 0;
        while (cnt >= (1 << c_J[run_index])) {                          // :295-300, whole blocks at once
            cnt -= 1 << c_J[run_index];
            ones++;
            if (run_index < 31) run_index++;
        }
        if (eol) {                                                      // :302-303
            if (cnt > 0) ones++;
            const uint32_t first = pos - L + 1;
            codes[first] = (1u << ones) - 1u;
            lens[first]  = (uint8_t)ones;
            continue;
        }
        if (ones > 0) {
            const uint32_t first = pos - L;
            codes[first] = (1u << ones) - 1u;
            lens[first]  = (uint8_t)ones;
        }
        const int jj = c_J[run_index];
        const int glimit = 23 - 1 - jj;                                 // :306
        if (run_index > 0) run_index--;                                 // :310-311
        const int row = pos / (uint32_t)W, col = pos - (uint32_t)row * (uint32_t)W;
        const int x = img[pos];
        const int b = row > 0 ? img[pos - W] : 0;
        const int a = col > 0 ? img[pos - 1] : b;
        const int t = (a == b);                                         // RItype :313
        const bool neg = a > b;                                         // :314
        const int px = t ? a : b;
        int e = x - px;
        if (neg) e = -e;
        e = (int)(int8_t)e;
        const int At = t ? A1 : A0, Bt = t ? B1 : B0, Nt = t ? N1 : N0;
        const int k = golomb_k(At + (t ? (Nt >> 1) : 0), Nt);           // :329
        const int map = (e != 0) && ((e > 0) == (k == 0 && 2 * Bt < Nt));   // :331
        const int me = 2 * abs(e) - t - map;
        uint32_t gcode; int glen;
        golomb_code(me, k, glimit, 8, gcode, glen);
        codes[pos] = ((uint32_t)cnt << glen) | gcode;                   // run remainder in J+1 bits, then the Golomb code
        lens[pos]  = (uint8_t)(jj + 1 + glen);
        int An = At + ((me + 1 - t) >> 1), Bn = Bt + (e < 0), Nn = Nt;  // :336-344
        if (Nn >= 64) { An >>= 1; Bn >>= 1; Nn >>= 1; }
        Nn++;
        if (t) { A1 = An; B1 = Bn; N1 = Nn; } else { A0 = An; B0 = Bn; N0 = Nn; }
    }
}

// ------------------------------------------------------------------------------------------------
// K5: raster-order prefix sum of the code lengths and MSB-first packing into the unstuffed stream.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t *sh_warp /*[8]*/, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) sh_warp[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < PK_THREADS / 32; w++) {
        const uint32_t s = sh_warp[w];
        if (w < warp) wbase += s;
        tot += s;
    }
    total = tot;
    return wbase + inc - v;
}

__device__ __forceinline__ void load_lens8(const uint8_t *lens, size_t idx0, size_t npix, uint32_t l[PK_PER])
{
    if (idx0 + PK_PER <= npix) {
        const uint2 v = *reinterpret_cast<const uint2 *>(lens + idx0);
#pragma unroll
        for (int k = 0; k < 4; k++) { l[k] = (v.x >> (8 * k)) & 0xff; l[4 + k] = (v.y >> (8 * k)) & 0xff; }
    } else {
#pragma unroll
        for (int k = 0; k < PK_PER; k++) l[k] = (idx0 + k < npix) ? lens[idx0 + k] : 0;
    }
}

__global__ void __launch_bounds__(PK_THREADS)
lens_sum_kernel(const uint8_t *__restrict__ lens_all, size_t slot_stride, size_t npix, int tiles_per_plane,
                uint32_t *__restrict__ tile_bits)
{
    __shared__ uint32_t sh_warp[PK_THREADS / 32];
    const int p = blockIdx.x / tiles_per_plane, tl = blockIdx.x - p * tiles_per_plane;
    const size_t idx0 = (size_t)tl * PK_TILE + (size_t)threadIdx.x * PK_PER;
    uint32_t l[PK_PER];
    load_lens8(lens_all + (size_t)p * slot_stride, idx0, npix, l);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < PK_PER; k++) s += l[k];
    uint32_t total;
    (void)block_excl_scan_256(s, sh_warp, total);
    if (threadIdx.x == 0) tile_bits[blockIdx.x] = total;
}

// one block per plane: exclusive scan of the tile bit counts (64-bit), zero the words tiles will OR into
__global__ void __launch_bounds__(PK_THREADS)
tile_scan_kernel(const uint32_t *__restrict__ tile_bits, uint64_t *__restrict__ tile_off, uint64_t *__restrict__ plane_bits,
                 int tiles_per_plane, uint32_t *__restrict__ ubits, size_t uw_stride)
{
    __shared__ uint32_t sh_warp[PK_THREADS / 32];
    const int p = blockIdx.x;
    const uint32_t *tb = tile_bits + (size_t)p * tiles_per_plane;
    uint64_t *to = tile_off + (size_t)p * tiles_per_plane;
    uint32_t *ub = ubits + (size_t)p * uw_stride;
    uint64_t carry = 0;
    for (int t0 = 0; t0 < tiles_per_plane; t0 += PK_THREADS) {
        const int t = t0 + threadIdx.x;
        const uint32_t v = t < tiles_per_plane ? tb[t] : 0;
        uint32_t total;
        const uint32_t ex = block_excl_scan_256(v, sh_warp, total);
        if (t < tiles_per_plane) {
            const uint64_t off = carry + ex;
            to[t] = off;
            ub[off >> 5] = 0;                       // first word of the tile may be shared with its predecessor
        }
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x < 4) ub[(carry >> 5) + threadIdx.x] = 0;    // tail words the stuffer may touch
    if (threadIdx.x == 0) plane_bits[p] = carry;
}

__global__ void __launch_bounds__(PK_THREADS)
pack_kernel(const uint32_t *__restrict__ codes_all, const uint8_t *__restrict__ lens_all, size_t slot_stride, size_t npix,
            int tiles_per_plane, const uint64_t *__restrict__ tile_off, uint32_t *__restrict__ ubits, size_t uw_stride)
{
    __shared__ uint32_t sh_words[PK_TILE + 2];
    __shared__ uint32_t sh_warp[PK_THREADS / 32];
    const int p = blockIdx.x / tiles_per_plane, tl = blockIdx.x - p * tiles_per_plane;
    const size_t idx0 = (size_t)tl * PK_TILE + (size_t)threadIdx.x * PK_PER;
    const uint8_t  *lens  = lens_all + (size_t)p * slot_stride;
    const uint32_t *codes = codes_all + (size_t)p * slot_stride;

    uint32_t l[PK_PER], c[PK_PER];
    load_lens8(lens, idx0, npix, l);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < PK_PER; k++) { c[k] = l[k] ? codes[idx0 + k] : 0; s += l[k]; }
    for (int w = threadIdx.x; w < PK_TILE + 2; w += PK_THREADS) sh_words[w] = 0;
    uint32_t total;
    uint32_t my = block_excl_scan_256(s, sh_warp, total);      // contains a __syncthreads after the zeroing
    if (total == 0) return;

    const uint64_t off = tile_off[blockIdx.x];
    const uint32_t s0 = (uint32_t)off & 31u;
    uint32_t bit = s0 + my;
#pragma unroll
    for (int k = 0; k < PK_PER; k++) {
        if (l[k]) {
            const uint32_t w = bit >> 5, sh = bit & 31u;
            const uint64_t v = (uint64_t)c[k] << (64 - l[k] - sh);      // left-align the code at stream bit `bit`
            atomicOr(&sh_words[w], (uint32_t)(v >> 32));
            if ((uint32_t)v) atomicOr(&sh_words[w + 1], (uint32_t)v);
            bit += l[k];
        }
    }
    __syncthreads();
    const uint32_t endbit = s0 + total;
    const uint32_t nwords = (endbit + 31) >> 5;
    uint32_t *ub = ubits + (size_t)p * uw_stride + (off >> 5);
    for (uint32_t w = threadIdx.x; w < nwords; w += PK_THREADS) {
        const uint32_t v = sh_words[w];
        const bool shared_first = (w == 0) && s0 != 0;
        const bool shared_last  = (w == nwords - 1) && (endbit & 31u) != 0;
        if (shared_first || shared_last) atomicOr(&ub[w], v);
        else ub[w] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// K6: marker-avoidance bit stuffing (writeBit :156-169, flushBits :177-184).  After an 0xFF byte the next
// byte carries a leading 0 and 7 payload bits.  One warp per plane: a window of 128 output bytes is
// produced in parallel under the current bit phase; the first 0xFF in the window ends it.
// dst of plane p: component 0 of a frame goes straight to its place in the frame's output slot,
// components 1,2 (RGB) go to scratch and are moved by assemble_kernel once the sizes are known.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
stuff_kernel(const uint32_t *__restrict__ ubits, size_t uw_stride, const uint64_t *__restrict__ plane_bits,
             FrameLayout L, uint64_t *__restrict__ plane_bytes)
{
    __shared__ uint32_t buf[ST_CHUNK_WORDS];
    const int p = blockIdx.x, lane = threadIdx.x;
    const uint64_t T = plane_bits[p];
    const uint32_t *in = ubits + (size_t)p * uw_stride;
    size_t cap;
    uint8_t *out = plane_dst(L, p, cap);

    uint64_t in_pos = 0, out_pos = 0;
    uint32_t pz = 0;                                    // 1: the next byte starts with the stuffed 0 bit
    uint64_t cw0 = ~(uint64_t)0;                        // first word held in buf (none yet)
    while (in_pos < T) {
        const uint64_t s = in_pos - pz;                 // stream bit that lands on the MSB of this window's first byte
        const uint64_t w0 = s >> 5;
        const uint32_t sh = (uint32_t)s & 31u;
        if (cw0 == ~(uint64_t)0 || w0 < cw0 || w0 + 33 > cw0 + ST_CHUNK_WORDS) {
            __syncwarp();
            cw0 = w0;
            for (int k = lane; k < ST_CHUNK_WORDS; k += 32)
                buf[k] = (cw0 + k < uw_stride) ? in[cw0 + k] : 0u;
            __syncwarp();
        }
        const uint32_t hi = buf[w0 - cw0 + lane], lo = buf[w0 - cw0 + lane + 1];
        uint32_t v = __funnelshift_l(lo, hi, sh);
        const int64_t valid = (int64_t)T - (int64_t)(s + 32u * (uint32_t)lane);     // payload bits left at this lane
        if (valid <= 0) v = 0;
        else if (valid < 32) v &= ~(0xFFFFFFFFu >> valid);
        if (pz && lane == 0) v &= 0x7FFFFFFFu;
        int fb = 4;
        if      ((v >> 24) == 0xFFu)          fb = 0;
        else if (((v >> 16) & 0xFFu) == 0xFFu) fb = 1;
        else if (((v >> 8) & 0xFFu) == 0xFFu)  fb = 2;
        else if ((v & 0xFFu) == 0xFFu)         fb = 3;
        const uint32_t ball = __ballot_sync(0xffffffffu, fb < 4);
        uint32_t nbytes = 128;
        bool ends_ff = false;
        if (ball) {
            const int f = __ffs(ball) - 1;
            nbytes = 4u * f + (uint32_t)__shfl_sync(0xffffffffu, fb, f) + 1u;
            ends_ff = true;
        }
        const uint64_t need = (T - s + 7) >> 3;         // bytes until the stream is exhausted
        if (need < nbytes) { nbytes = (uint32_t)need; ends_ff = false; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t bi = 4u * lane + k;
            if (bi < nbytes && out_pos + bi < cap) out[out_pos + bi] = (uint8_t)(v >> (24 - 8 * k));
        }
        out_pos += nbytes;
        in_pos = s + 8ull * nbytes;
        pz = ends_ff ? 1u : 0u;
    }
    if (lane == 0) {
        if (pz) { if (out_pos < cap) out[out_pos] = 0; out_pos++; }      // flushBits after a trailing 0xFF (:178)
        plane_bytes[p] = out_pos;
    }
}

// K7: headers, trailers, moving components 1,2 into place, per-frame length.  grid = (frames, slices).
__global__ void __launch_bounds__(256)
assemble_kernel(FrameLayout L, int H, int W, int near_, const uint64_t *__restrict__ plane_bytes,
                uint64_t *__restrict__ out_len, int *__restrict__ overflow)
{
    const int f = blockIdx.x;
    uint8_t *out = L.out + (size_t)f * L.out_stride;
    // offsets of the scans
    size_t scan_off[3], scan_len[3];
    size_t o = L.hdr_len;
    for (int c = 0; c < L.ncomp; c++) {
        o += 10;
        scan_off[c] = o;
        scan_len[c] = plane_bytes[(size_t)f * L.ncomp + c];
        o += scan_len[c];
    }
    const size_t total = o + 2;
    const bool fits = total <= L.out_stride;
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        out_len[f] = fits ? total : 0;
        if (!fits) { atomicExch(overflow, 1); }
        else {
            uint8_t *h = out;                                               // :211-232
            h[0] = 0xFF; h[1] = 0xD8; h[2] = 0xFF; h[3] = 0xF7; h[4] = 0x00; h[5] = (uint8_t)(8 + 3 * L.ncomp);
            h[6] = 8; h[7] = (uint8_t)(H >> 8); h[8] = (uint8_t)H; h[9] = (uint8_t)(W >> 8); h[10] = (uint8_t)W;
            h[11] = (uint8_t)L.ncomp;
            for (int c = 0; c < L.ncomp; c++) { h[12 + 3 * c] = (uint8_t)(c + 1); h[13 + 3 * c] = 0x11; h[14 + 3 * c] = 0x00; }
            for (int c = 0; c < L.ncomp; c++) {                             // :200-208
                uint8_t *s = out + scan_off[c] - 10;
                s[0] = 0xFF; s[1] = 0xDA; s[2] = 0x00; s[3] = 0x08; s[4] = 0x01; s[5] = (uint8_t)(c + 1);
                s[6] = 0x00; s[7] = (uint8_t)near_; s[8] = 0x00; s[9] = 0x00;
            }
            out[total - 2] = 0xFF; out[total - 1] = 0xD9;                   // :235-237
        }
    }
    if (!fits) return;
    for (int c = 1; c < L.ncomp; c++) {
        const uint8_t *src = L.scratch + ((size_t)f * L.ncomp + c) * L.scratch_stride;
        uint8_t *dst = out + scan_off[c];
        for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < scan_len[c]; i += (size_t)gridDim.y * blockDim.x)
            dst[i] = src[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Serial scan: any NEAR.  One warp per plane, lane 0 walks JLSencodeScan (:240-399) sample by sample with
// the context state in shared memory and a 64-bit bit accumulator that applies the stuffing rule on the fly.
// ------------------------------------------------------------------------------------------------
// SerialParams, grad_q_any and the row-pipelined scan kernel (the path for NEAR > 0) live in jls_rowscan_kernel.cuh

struct BitSink {
    uint8_t *p; size_t n, cap; uint64_t acc; int nacc; bool after_ff;
    __device__ __forceinline__ void put(uint32_t v, int nb)
    {
        acc = (acc << nb) | v;
        nacc += nb;
        for (;;) {
            const int need = after_ff ? 7 : 8;
            if (nacc < need) break;
            const uint32_t byte = (uint32_t)(acc >> (nacc - need)) & ((1u << need) - 1u);
            if (n < cap) p[n] = (uint8_t)byte;
            n++;
            nacc -= need;
            after_ff = (byte == 0xFFu);
        }
        acc &= (1ull << nacc) - 1ull;
    }
    __device__ __forceinline__ void zeros_one(int z)
    {
        while (z > 24) { put(0, 24); z -= 24; }
        put(1, z + 1);
    }
    __device__ __forceinline__ void flush()
    {
        if (nacc > 0 || after_ff) {
            const int need = after_ff ? 7 : 8;
            const uint32_t byte = (uint32_t)(acc << (need - nacc)) & ((1u << need) - 1u);
            if (n < cap) p[n] = (uint8_t)byte;
            n++;
            nacc = 0; acc = 0; after_ff = false;
        }
    }
};


__device__ __forceinline__ void golomb_put(BitSink &s, int me, int k, int limit, int qbpp)
{
    const int hi = me >> k;
    if (hi < limit) {
        s.zeros_one(hi);
        if (k) s.put((uint32_t)me & ((1u << k) - 1u), k);
    } else {
        s.zeros_one(limit);
        s.put((uint32_t)(me - 1) & ((1u << qbpp) - 1u), qbpp);
    }
}

__global__ void __launch_bounds__(32)
serial_scan_kernel(const uint8_t *__restrict__ pix, size_t plane_stride, uint8_t *__restrict__ recon_all,
                   int W, int H, SerialParams P, FrameLayout L, uint64_t *__restrict__ plane_bytes)
{
    __shared__ int4 ctx[NCTX];                          // (A, B, C, N) per context
    const int p = blockIdx.x;
    for (int k = threadIdx.x; k < NCTX; k += 32) ctx[k] = make_int4(P.a_init, 0, 0, 1);
    __syncwarp();
    if (threadIdx.x != 0) return;

    const uint8_t *src = pix + (size_t)p * plane_stride;
    uint8_t *rec = recon_all + (size_t)p * plane_stride;        // reconstruction (== input when NEAR = 0)
    size_t cap;
    BitSink s;
    s.p = plane_dst(L, p, cap); s.cap = cap; s.n = 0; s.acc = 0; s.nacc = 0; s.after_ff = false;

    int rA[2] = {P.a_init, P.a_init}, rB[2] = {0, 0}, rN[2] = {1, 1};
    int run_index = 0;
    const int near_ = P.near_;
    for (int y = 0; y < H; y++) {
        const uint8_t *in_row = src + (size_t)y * W;
        uint8_t *cur = rec + (size_t)y * W;
        const uint8_t *up = cur - W;
        bool in_run = false;
        int run_len = 0;
        int a = 0, c = 0, b = 0;                         // sliding template
        int d = 0;
        if (y > 0) { b = up[0]; }
        a = b;
        c = (y > 1) ? up[-W] : 0;
        for (int x = 0; x < W; x++) {
            d = (y > 0) ? ((x + 1 < W) ? up[x + 1] : b) : 0;
            const int v = in_row[x];
            const int qv = 81 * grad_q_any(P, d - b) + 9 * grad_q_any(P, b - c) + grad_q_any(P, c - a);
            bool neg = qv < 0;
            const int cq = abs(qv);
            if (cq == 0) in_run = true;
            bool hit = false, brk = false;
            if (in_run) { hit = abs(v - a) <= near_; brk = !hit; in_run = hit; }
            int rx;
            if (hit) {
                rx = a;
                run_len++;
                if (run_len >= (1 << c_J[run_index])) {
                    s.put(1, 1);
                    run_len -= 1 << c_J[run_index];
                    if (run_index < 31) run_index++;
                }
                if (x == W - 1 && run_len > 0) s.put(1, 1);
            } else if (brk) {
                const int jj = c_J[run_index];
                const int glimit = P.limit - 1 - jj;
                s.put((uint32_t)run_len, jj + 1);
                run_len = 0;
                if (run_index > 0) run_index--;
                const int t = abs(a - b) <= near_;
                neg = a > b + near_;
                const int px = t ? a : b;
                int e = neg ? px - v : v - px;
                e = e < 0 ? -((near_ - e) / P.quant) : (near_ + e) / P.quant;
                rx = near_ ? min(max(px + (neg ? -P.quant * e : P.quant * e), 0), 255) : v;
                if (e < 0) e += P.range;
                if (e >= (P.range + 1) / 2) e -= P.range;
                const int k = golomb_k(rA[t] + (t ? (rN[t] >> 1) : 0), rN[t]);
                const int map = (e != 0) && ((e > 0) == (k == 0 && 2 * rB[t] < rN[t]));
                const int me = 2 * abs(e) - t - map;
                golomb_put(s, me, k, glimit, P.qbpp);
                if (e < 0) rB[t]++;
                rA[t] += (me + 1 - t) >> 1;
                if (rN[t] >= 64) { rA[t] >>= 1; rB[t] >>= 1; rN[t] >>= 1; }
                rN[t]++;
            } else {
                run_len = 0;
                int4 st = ctx[cq - 1];
                int px = med3(a, b, c) + (neg ? -st.z : st.z);
                px = min(max(px, 0), 255);
                int e = neg ? px - v : v - px;
                e = e < 0 ? -((near_ - e) / P.quant) : (near_ + e) / P.quant;
                rx = near_ ? min(max(px + (neg ? -P.quant * e : P.quant * e), 0), 255) : v;
                if (e < 0) e += P.range;
                if (e >= (P.range + 1) / 2) e -= P.range;
                const int k = golomb_k(st.x, st.w);
                const int map = (k == 0) && (2 * st.y <= -st.w) && (near_ == 0);
                const int me = 2 * abs(e) + (e < 0 ? -(map + 1) : map);
                golomb_put(s, me, k, P.limit, P.qbpp);
                st.y += e * P.quant;
                st.x += abs(e);
                if (st.w >= 64) { st.x >>= 1; st.y >>= 1; st.w >>= 1; }
                st.w++;
                if (st.y <= -st.w) {
                    st.y += st.w; if (st.y <= -st.w) st.y = -st.w + 1;
                    st.z = max(st.z - 1, -128);
                } else if (st.y > 0) {
                    st.y -= st.w; if (st.y > 0) st.y = 0;
                    st.z = min(st.z + 1, 127);
                }
                ctx[cq - 1] = st;
            }
            cur[x] = (uint8_t)rx;
            c = b; b = d; a = rx;                         // slide the template (a = reconstructed left neighbour)
        }
    }
    s.flush();
    plane_bytes[p] = s.n;
}

// ------------------------------------------------------------------------------------------------
// host side: one batch through the device pipeline
// ------------------------------------------------------------------------------------------------
static inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }


#ifdef IMCVT_JLS_PROF
static unsigned long long *g_jls_prof = nullptr;
#endif
// d_frames: device, interleaved.  d_out / d_out_len: device.  Queues everything on `st`.
// d_overflow: device int set to 1 when a slot was too small.
// Whether a gray batch of n planes would run on the cluster-pair tile kernel: the one launch shape that is built for frames
// and streams in page-locked HOST memory (rows requested a tick ahead of their use, output in whole 128-byte lines).
bool host_in_place_ok(int n, int is_rgb, uint32_t W, int near_)
{
    Engine &E = engine();
    const bool lossless = (near_ == 0) && !E.jls_force_serial;
    const bool use_tile = lossless && (int)W <= tile::WMAX && E.jls_path != 1;
    const bool pair = E.jls_path == 3 || ((E.jls_path == 0 || E.jls_path == 6) && 2 * n <= E.sm_count);
    return !is_rgb && use_tile && pair;
}

int encode_device(const uint8_t *d_frames, size_t frame_stride, int n, int is_rgb, uint32_t H, uint32_t W, int near_,
                  uint8_t *d_out, size_t out_stride, uint64_t *d_out_len, int *d_overflow, cudaStream_t st, Workspace *wsp, bool out_is_host)
{
    Engine &E = engine();
    Workspace &WS = wsp ? *wsp : E.ws;                  // concurrent calls (chunks of a host batch) bring their own workspace
    const int ncomp = is_rgb ? 3 : 1;
    const int P = n * ncomp;
    const size_t npix = (size_t)H * W;
    const size_t slot_stride = round_up(npix, 16);
    const bool lossless = (near_ == 0) && !E.jls_force_serial;
    // planes up to tile::WMAX wide stream through shared memory (jls_tile_kernel.cuh); wider ones take the multi-kernel path
    const bool use_tile = lossless && (int)W <= tile::WMAX && E.jls_path != 1;

    // ---- workspace layout ----
    const int T  = (int)((H + ROWS_PER_TILE - 1) / ROWS_PER_TILE);             // classify tiles per plane
    const int PT = (int)((npix + PK_TILE - 1) / PK_TILE);                       // packer tiles per plane
    const size_t uw_stride = round_up(npix + 64, 64);                           // words: <= 32 bits per sample
    const size_t scratch_stride = round_up(npix * 4 + npix * 4 / 7 + 64, 256);  // stuffed bytes of one scan, worst case
    const bool need_planes = is_rgb || frame_stride != npix;                    // gray, tightly packed: use the input in place
    size_t bytes = 4096;
    auto add = [&](size_t b) { bytes += round_up(b, 256); };
    if (need_planes) add((size_t)P * slot_stride);
    if (use_tile) {
        add(256);                                  // plane queue counter
    } else if (lossless) {
        add((size_t)P * NBUCKET * T * 4);          // hist
        add((size_t)P * NBUCKET * 4 * 2);          // bucket_total, bucket_start
        add((size_t)P * slot_stride * 4);          // rec_pos
        add((size_t)P * slot_stride * 2);          // rec_aux
        add((size_t)P * slot_stride * 4);          // codes
        add((size_t)P * slot_stride);              // lens
        add((size_t)P * PT * 4);                   // tile_bits
        add((size_t)P * PT * 8);                   // tile_off
        add((size_t)P * 8);                        // plane_bits
        add((size_t)P * uw_stride * 4);            // unstuffed bits
    } else {
        add((size_t)P * slot_stride);              // reconstruction
    }
    add((size_t)P * 8);                            // plane_bytes
    if (ncomp > 1) add((size_t)P * scratch_stride);
    if (WS.reserve(bytes)) return 1;
    if (WS.wait_previous(st)) return 1;

    const uint8_t *planes = d_frames;
    size_t plane_stride = frame_stride;
    if (need_planes) {
        uint8_t *pl = WS.take<uint8_t>((size_t)P * slot_stride);
        plane_stride = slot_stride;
        if (is_rgb) {
            LaunchScope ls("jls_deinterleave", st);
            deinterleave_rgb_kernel<<<E.sm_count * 8, 256, 0, st>>>(d_frames, frame_stride, n, npix, pl, plane_stride);
        } else {
            IMCVT_CUDA(cudaMemcpy2DAsync(pl, plane_stride, d_frames, frame_stride, npix, n, cudaMemcpyDeviceToDevice, st));
        }
        planes = pl;
    }

    FrameLayout Lay;
    Lay.ncomp = ncomp; Lay.hdr_len = is_rgb ? 21 : 15;
    Lay.out = d_out; Lay.out_stride = out_stride;
    Lay.scratch = nullptr; Lay.scratch_stride = scratch_stride;

    uint64_t *plane_bytes = nullptr;

    if (use_tile) {
        int *next = WS.take<int>(1);
        plane_bytes = WS.take<uint64_t>((size_t)P);
        if (ncomp > 1) Lay.scratch = WS.take<uint8_t>((size_t)P * scratch_stride);
        if (!next || !plane_bytes || (ncomp > 1 && !Lay.scratch)) { set_error("workspace carve failed"); return 1; }
        if (!E.attr_jls_tile) {
            IMCVT_CUDA(cudaFuncSetAttribute(tile::jls_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tile::Smem)));
            IMCVT_CUDA(cudaFuncSetAttribute(tile::jls_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tile::Smem)));
            E.attr_jls_tile = true;
        }
        // Few planes: two SMs per plane (a cluster of two CTAs: rank 0 = producers, rank 1 = two chain stages + packers).
        // Many planes: one CTA per plane keeps every SM on a plane of its own.
        const bool pair = E.jls_path == 3 || ((E.jls_path == 0 || E.jls_path == 6) && 2 * P <= E.sm_count);
        IMCVT_CUDA(cudaMemsetAsync(next, 0, sizeof(int), st));
        tile::Params TP;
        TP.pix = planes; TP.plane_stride = plane_stride; TP.P = P; TP.W = (int)W; TP.H = (int)H;
        TP.L = Lay; TP.plane_bytes = (unsigned long long *)plane_bytes; TP.next_plane = next;
        TP.prof = nullptr;
        TP.line_out = (out_is_host && ncomp == 1) ? 1 : 0;     // RGB scans 1, 2 go through device scratch and assemble_kernel
        TP.tma = E.jls_path == 6 ? 0 : 1;                      // imcvtDebugJLSPath(6): row tiles through registers instead of bulk copies
#ifdef IMCVT_JLS_PROF
        {
            if (!g_jls_prof) { IMCVT_CUDA(cudaMalloc(&g_jls_prof, 16 * 8)); IMCVT_CUDA(cudaMemset(g_jls_prof, 0, 16 * 8)); }
            TP.prof = g_jls_prof;
        }
#endif
        if (pair) {
            LaunchScope ls("jls_pair", st);
            const int pairs = P < E.sm_count / 2 ? P : E.sm_count / 2;
            tile::jls_pair_kernel<<<2 * pairs, tile::NTHREADS, sizeof(tile::Smem), st>>>(TP);
        } else {
            LaunchScope ls("jls_tile", st);
            tile::jls_tile_kernel<<<P < E.sm_count ? P : E.sm_count, tile::NTHREADS, sizeof(tile::Smem), st>>>(TP);
        }
    } else if (lossless) {
        uint32_t *hist         = WS.take<uint32_t>((size_t)P * NBUCKET * T);
        uint32_t *bucket_total = WS.take<uint32_t>((size_t)P * NBUCKET);
        uint32_t *bucket_start = WS.take<uint32_t>((size_t)P * NBUCKET);
        uint32_t *rec_pos      = WS.take<uint32_t>((size_t)P * slot_stride);
        uint16_t *rec_aux      = WS.take<uint16_t>((size_t)P * slot_stride);
        uint32_t *codes        = WS.take<uint32_t>((size_t)P * slot_stride);
        uint8_t  *lens         = WS.take<uint8_t>((size_t)P * slot_stride);
        uint32_t *tile_bits    = WS.take<uint32_t>((size_t)P * PT);
        uint64_t *tile_off     = WS.take<uint64_t>((size_t)P * PT);
        uint64_t *plane_bits   = WS.take<uint64_t>((size_t)P);
        uint32_t *ubits        = WS.take<uint32_t>((size_t)P * uw_stride);
        plane_bytes            = WS.take<uint64_t>((size_t)P);
        if (ncomp > 1) Lay.scratch = WS.take<uint8_t>((size_t)P * scratch_stride);
        if (!ubits || !plane_bytes || (ncomp > 1 && !Lay.scratch)) { set_error("workspace carve failed"); return 1; }

        const int total_tiles = P * T;
        const int cls_blocks = (total_tiles + CLS_WARPS - 1) / CLS_WARPS;
        {
            LaunchScope ls("jls_classify_count", st);
            classify_kernel<false><<<cls_blocks, CLS_WARPS * 32, 0, st>>>(planes, plane_stride, (int)W, (int)H, T, total_tiles,
                hist, nullptr, nullptr, nullptr, slot_stride, nullptr, slot_stride);
        }
        {
            LaunchScope ls("jls_hist_scan", st);
            const int rows = P * NBUCKET;
            hist_scan_kernel<<<(rows * 32 + 255) / 256, 256, 0, st>>>(hist, bucket_total, T, rows);
        }
        {
            LaunchScope ls("jls_bucket_offsets", st);
            bucket_offsets_kernel<<<(P * 32 + 127) / 128, 128, 0, st>>>(bucket_total, bucket_start, P);
        }
        {
            LaunchScope ls("jls_classify_scatter", st);
            classify_kernel<true><<<cls_blocks, CLS_WARPS * 32, 0, st>>>(planes, plane_stride, (int)W, (int)H, T, total_tiles,
                hist, bucket_start, rec_pos, rec_aux, slot_stride, lens, slot_stride);
        }
        {
            LaunchScope ls("jls_chain", st);
            const int reg_blocks = (P * NCTX + 127) / 128;
            const int run_blocks = (P + 3) / 4;
            chain_kernel<<<reg_blocks + run_blocks, 128, 0, st>>>(planes, plane_stride, (int)W, P, reg_blocks,
                bucket_start, bucket_total, rec_pos, rec_aux, slot_stride, codes, lens, slot_stride);
        }
        {
            LaunchScope ls("jls_lens_sum", st);
            lens_sum_kernel<<<P * PT, PK_THREADS, 0, st>>>(lens, slot_stride, npix, PT, tile_bits);
        }
        {
            LaunchScope ls("jls_tile_scan", st);
            tile_scan_kernel<<<P, PK_THREADS, 0, st>>>(tile_bits, tile_off, plane_bits, PT, ubits, uw_stride);
        }
        {
            LaunchScope ls("jls_pack", st);
            pack_kernel<<<P * PT, PK_THREADS, 0, st>>>(codes, lens, slot_stride, npix, PT, tile_off, ubits, uw_stride);
        }
        {
            LaunchScope ls("jls_stuff", st);
            stuff_kernel<<<P, 32, 0, st>>>(ubits, uw_stride, plane_bits, Lay, plane_bytes);
        }
    } else {
        uint8_t *recon = WS.take<uint8_t>((size_t)P * slot_stride);
        plane_bytes    = WS.take<uint64_t>((size_t)P);
        if (ncomp > 1) Lay.scratch = WS.take<uint8_t>((size_t)P * scratch_stride);
        if (!recon || !plane_bytes || (ncomp > 1 && !Lay.scratch)) { set_error("workspace carve failed"); return 1; }
        // the serial kernel addresses the reconstruction with the input's plane stride (<= slot_stride)
        SerialParams SP = make_params(near_);
        if (E.jls_path == 4) {                          // cross-check: the plain one-thread scan
            LaunchScope ls("jls_serial_scan", st);
            serial_scan_kernel<<<P, 32, 0, st>>>(planes, plane_stride, recon, (int)W, (int)H, SP, Lay, plane_bytes);
        } else {
            if (!E.attr_jls_rowscan) {
                IMCVT_CUDA(cudaFuncSetAttribute(rowscan::rowscan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(rowscan::Smem)));
                IMCVT_CUDA(cudaFuncSetAttribute(rowscan::rowscan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(rowscan::Smem)));
                E.attr_jls_rowscan = true;
            }
            if (near_ > 0 && E.jls_path != 5) {
                // NEAR 1..4: candidate-table kernel (jls_nearscan_kernel.cuh); imcvtDebugJLSPath(5) keeps the three-warp row scan as a cross-check
                if (!E.attr_jls_nearscan) {
                    IMCVT_CUDA(cudaFuncSetAttribute(nearscan::nearscan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(nearscan::Smem)));
                    IMCVT_CUDA(cudaFuncSetAttribute(nearscan::nearscan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(nearscan::Smem)));
                    IMCVT_CUDA(cudaFuncSetAttribute(nearscan::nearscan_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(nearscan::Smem)));
                    IMCVT_CUDA(cudaFuncSetAttribute(nearscan::nearscan_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(nearscan::Smem)));
                    E.attr_jls_nearscan = true;
                }
                LaunchScope ls("jls_nearscan", st);
                const size_t smem = sizeof(nearscan::Smem);
                switch (near_) {
                case 1: nearscan::nearscan_kernel<1><<<P, nearscan::NT, smem, st>>>(planes, plane_stride, (int)W, (int)H, SP, Lay, plane_bytes); break;
                case 2: nearscan::nearscan_kernel<2><<<P, nearscan::NT, smem, st>>>(planes, plane_stride, (int)W, (int)H, SP, Lay, plane_bytes); break;
                case 3: nearscan::nearscan_kernel<3><<<P, nearscan::NT, smem, st>>>(planes, plane_stride, (int)W, (int)H, SP, Lay, plane_bytes); break;
                default: nearscan::nearscan_kernel<4><<<P, nearscan::NT, smem, st>>>(planes, plane_stride, (int)W, (int)H, SP, Lay, plane_bytes); break;
                }
            } else {
                LaunchScope ls("jls_rowscan", st);
                if (near_ > 0) rowscan::rowscan_kernel<true><<<P, rowscan::NT, sizeof(rowscan::Smem), st>>>(planes, plane_stride, (int)W, (int)H, SP, Lay, plane_bytes);
                else rowscan::rowscan_kernel<false><<<P, rowscan::NT, sizeof(rowscan::Smem), st>>>(planes, plane_stride, (int)W, (int)H, SP, Lay, plane_bytes);
            }
        }
    }
    {
        LaunchScope ls("jls_assemble", st);
        dim3 grid(n, ncomp > 1 ? 32 : 1);
        assemble_kernel<<<grid, 256, 0, st>>>(Lay, (int)H, (int)W, near_, plane_bytes, d_out_len, d_overflow);
    }
    IMCVT_CUDA(cudaGetLastError());
    return WS.mark_busy(st);
}

}  // namespace jls
}  // namespace imcvt

#ifdef IMCVT_JLS_PROF
// development build only (make PROF=1): reads and clears the role cycle counters of jls_tile_kernel
extern "C" int imcvtDebugJlsProf(unsigned long long *out16)
{
    if (!imcvt::jls::g_jls_prof) return 1;
    cudaDeviceSynchronize();
    cudaMemcpy(out16, imcvt::jls::g_jls_prof, 16 * 8, cudaMemcpyDeviceToHost);
    cudaMemset(imcvt::jls::g_jls_prof, 0, 16 * 8);
    return 0;
}
#endif
