// jls_nearscan_kernel.cuh -- JPEG-LS scan encoder for NEAR = 1..4 on sm_100a (reference: JLSencodeScan, src/imageio_jls.c:240-399
// with the reconstruction feedback of :320-323).  One CTA per plane; included by jls.cu, and compiled for the host by
// tests/emu/jls_nearscan_emu.cpp under the SIMT emulator.
//
// With NEAR > 0 the reconstructed value Rx of a sample is the left neighbour `a` of the next one, so a plane is ONE dependency
// chain.  What keeps that chain short here: |Rx - x| <= NEAR always (:97-102, :320-323), so `a` can only be one of the K = 2 NEAR + 1
// values x[i-1] - NEAR .. x[i-1] + NEAR -- and everything else the sample's decision needs (b, c, d from the finished row
// above, the sample itself) is known beforehand.  So:
//   stagers (4 warps)   for every sample of a segment and every candidate k of `a`: the whole context decision, as one 32-bit
//                       entry per (sample, k) in shared memory -- regular mode: context number (as the byte offset of its state),
//                       sign and MED prediction (:79-94); run mode (:284-344): the sample's outcome is a function of a, b and x
//                       alone, so the entry already holds the reconstruction and the next candidate index.  Two tables: the
//                       sample reached outside a run / inside one (`running`, :284-289).
//   walker (lane 0 of warp 0)   the chain proper: entry = T[sample][state]  ->  context state (B, C, N)  ->  px = clamp(MED +- C)
//                       ->  one table look-up indexed by x - px that gives the wrapped error AND the next candidate index
//                       ->  next entry.  Three dependent shared-memory loads and a handful of integer instructions per sample;
//                       the B / C / N update (:376-393) and the sample's record for the coder ride beside it.
//   coder (warp 2)      a segment later: regular samples 32 at a time -- lanes whose samples share a context find each other
//                       with match.any and step that context's (A, N) in registers in raster order, the others are independent
//                       -- Golomb parameter, mapped error, code + length (:365-375); run lengths and run interruptions
//                       (:291-344: run_index, A / B / N of the two interruption contexts) on lane 0, they are rare.
//   packer (warp 1)     a segment after that: bit-packs and 0xFF-stuffs (the tile kernel's packer, :156-184).
// The stagers work a segment ahead of the walker except for the first segment of a row, which needs the finished row above.
#pragma once
#include "jls_rowscan_kernel.cuh"

namespace imcvt {
namespace jls {
namespace nearscan {

constexpr int SEGCAP = 1024;                        // samples per segment: 1024 (NEAR <= 2) or 512
constexpr int TWORDS = 13328;                       // one row buffer: (segment + 1) rows of 2 K + 3 words
constexpr int STAGE_WARPS = 4;
constexpr int NT = 32 * (3 + STAGE_WARPS);
constexpr int WMAXR = 32768;                        // :436
constexpr uint32_t REC_BREAK = 0x8000u;

template <int NEAR> struct Cfg {
    static constexpr int K = 2 * NEAR + 1;
    static constexpr int SEG = NEAR <= 2 ? 1024 : 512;
    // One row of words per sample: K entries for a sample reached outside a run | K entries inside a run | the sample's offset
    // into the walker's table | the walker's record for the coder | the candidate the walker left.  An odd number of words:
    // the stagers' lanes (a sample each) hit different banks.
    static constexpr int RS = 2 * K + 3;
    static constexpr int XW = 2 * K, RECW = 2 * K + 1, SELW = 2 * K + 2;
    static constexpr int RUN_OFS = K * 4;           // byte offset of the in-run entries inside a row
    static_assert((SEG + 1) * RS <= TWORDS, "row buffer too small");
};

// Entries (32 bit), type in [25:24], stop mark in [26].
//   0 regular           [15:0] 0x2000 | 16 (q - 1) (= the record's high half, and the context state's byte offset + 0x2000)
//                       [23:16] MED | [31] sign
//   2 regular, sample within NEAR of 0 or 255: same fields; the walker also clamps the next candidate to what 0..255 allows
//   1 run sample        [15:0] next state (byte offset of the candidate among the row's in-run entries)
//   3 run interruption  [5:0] next candidate (x 4) | [14:6] wrapped error | [15] RItype
//   the row after the segment's last sample holds stop marks: the walker's loop over regular samples needs no counter
// Records for the coder (one per sample): low half wrapped error | high half: 0 = run sample, REC_BREAK | RItype = run
// interruption, else 0x2000 | 16 (q - 1).
// The walker leaves the state it hands to the next sample (a byte; candidate = 4 (Rx - x + NEAR), + RUN_OFS inside a run); all
// threads turn a walked segment into reconstructed samples of the row buffer at the start of the next iteration.
constexpr uint32_t ENT_STOP = 0x07000000u;
// low half of a | low half of b << 16
__device__ __forceinline__ uint32_t pack_halves(uint32_t a, uint32_t b)
{
#ifdef IMCVT_HOST_EMU
    return (a & 0xFFFFu) | (b << 16);
#else
    return __byte_perm(a, b, 0x5410);
#endif
}
struct Smem {
    uint8_t row[2][WMAXR + 16];                     // reconstructed rows y-1 / y (index 8 + x)
    uint32_t T[2][TWORDS];
    uint32_t codes[2][SEGCAP + SEGCAP / 32];
    uint8_t lens[2][SEGCAP + SEGCAP / 8];
    uint32_t ubits[tile::UBITS_WORDS];
    int4 ctx[NCTX];                                 // walker: (B, C, N, unused)
    int2 ctx2[NCTX];                                // coder: (A, N)
    uint32_t dq2[1024];                             // [(x - px + 255) | negative-sign << 9]: wrapped error (low half) | next candidate x 4 (high half)
    signed char gq[512];                            // quantised gradient at [g + 255]
    short eq[512];                                  // quantised, wrapped error at [d + 255] (:97-111)
    short qe[512];                                  // the reconstruction step it stands for
};

template <int NEAR>
__device__ __forceinline__ void stage_segment(Smem &sm, const uint8_t *__restrict__ src, int W, int y, int x0, int n, int buf, int st_tid, int st_n)
{
    constexpr int K = Cfg<NEAR>::K, RS = Cfg<NEAR>::RS;
    const uint8_t *up = sm.row[(y + 1) & 1] + 8;
    const uint8_t *in_row = src + (size_t)y * W + x0;
    uint32_t *T = sm.T[buf];
    for (int i = st_tid; i < n; i += st_n) {
        const int x = x0 + i;
        const int xv = in_row[i];
        const int b = up[x], d = up[x + 1];                       // up[W] repeats up[W - 1] (:57-58)
        const bool first = x == 0;
        // column 0 (:47-56): a = b, c = the reconstruction two rows up, still in the buffer row y is about to overwrite
        const int c = first ? (int)sm.row[y & 1][8] : (int)up[x - 1];
        const int abase = first ? b : (int)in_row[i - 1] - NEAR;
        const int q12 = 81 * sm.gq[d - b + 255] + 9 * sm.gq[b - c + 255];
        const bool edge = xv < NEAR || xv > 255 - NEAR;
        uint32_t *row = T + i * RS;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int a = first ? b : min(max(abase + k, 0), 255);          // a candidate outside 0..255 is never selected
            uint32_t er;
            if (abs(xv - a) <= NEAR) {                                      // run sample (:291-303): Rx = a
                er = (1u << 24) | (uint32_t)(Cfg<NEAR>::RUN_OFS + (a - xv + NEAR) * 4);
            } else {                                                        // run interruption (:305-344)
                const int t = abs(a - b) <= NEAR;
                const bool neg = a > b + NEAR;
                const int px = t ? a : b;
                const int di = (neg ? px - xv : xv - px) + 255;
                const int e = sm.eq[di], step = sm.qe[di];
                const int rx = min(max(px + (neg ? -step : step), 0), 255);
                er = (3u << 24) | (uint32_t)((rx - xv + NEAR) * 4) | (((uint32_t)e & 0x1FFu) << 6) | ((uint32_t)t << 15);
            }
            row[K + k] = er;
            const int qv = q12 + sm.gq[c - a + 255];
            uint32_t eg = er;                                               // Q == 0 enters run mode (:284-289)
            if (qv != 0) {
                const int cq = abs(qv);
                eg = ((uint32_t)(cq - 1) << 4) | 0x2000u | ((uint32_t)med3(a, b, c) << 16) | (qv < 0 ? 0x80000000u : 0u) | (edge ? (2u << 24) : 0u);
            }
            row[k] = eg;
        }
        row[Cfg<NEAR>::XW] = (uint32_t)((xv + 255) * 4);
    }
    for (int k = st_tid; k < 2 * K; k += st_n) T[n * RS + k] = ENT_STOP;    // the row after the last sample stops the walker
}

template <int NEAR>
__global__ void __launch_bounds__(NT, 1)
nearscan_kernel(const uint8_t *__restrict__ pix, size_t plane_stride, int W, int H, SerialParams P, FrameLayout L,
                uint64_t *__restrict__ plane_bytes)
{
    constexpr int K = Cfg<NEAR>::K, SEG = Cfg<NEAR>::SEG, RS = Cfg<NEAR>::RS;
    constexpr int quant = 2 * NEAR + 1, range = (256 + 4 * NEAR) / quant, half = (range + 1) / 2;
    Smem &sm = *reinterpret_cast<Smem *>(JLS_ROWSCAN_SMEM());
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = blockIdx.x;
    const uint8_t *src = pix + (size_t)p * plane_stride;

    for (int k = tid; k < NCTX; k += NT) { sm.ctx[k] = make_int4(0, 0, 1, 0); sm.ctx2[k] = make_int2(P.a_init, 1); }
    for (int i = tid; i < 511; i += NT) {
        const int g = i - 255;
        sm.gq[i] = (signed char)grad_q_any(P, g);
        const int q = g < 0 ? -((NEAR - g) / quant) : (NEAR + g) / quant;   // quantize :97-102
        sm.qe[i] = (short)(quant * q);
        int e = q;                                                          // modRange :105-111
        if (e < 0) e += range;
        if (e >= half) e -= range;
        sm.eq[i] = (short)e;
        // the walker's table: d = x - px at [d + 255], for either sign of the context.  Rx = px + quant q(d) whatever the sign
        // (q is odd), so the next candidate is quant q(d) - d + NEAR; the coded error is sign * q(d), wrapped.
        const uint32_t nxt = (uint32_t)((quant * q - g + NEAR) * 4) << 16;
        sm.dq2[i] = ((uint32_t)e & 0xFFFFu) | nxt;
        int en = -q;
        if (en < 0) en += range;
        if (en >= half) en -= range;
        sm.dq2[512 + i] = ((uint32_t)en & 0xFFFFu) | nxt;
    }
    for (int i = tid; i < WMAXR + 16; i += NT) { sm.row[0][i] = 0; sm.row[1][i] = 0; }      // the row above the picture reads as zeros

    tile::PackState ps;
    {
        size_t cap;
        uint8_t *dst = plane_dst(L, p, cap);
        tile::pack_state_init(ps, dst, cap, nullptr);
    }

    const int nseg = (W + SEG - 1) / SEG, total = H * nseg;
    uint32_t sel = NEAR * 4;                        // walker: byte offset of the current candidate (+ RUN_OFS inside a run)
    // coder, lane 0: run state and the two run-interruption contexts
    int rA0 = P.a_init, rA1 = P.a_init, rB0 = 0, rB1 = 0, rN0 = 1, rN1 = 1;
    int run_index = 0, run_len = 0;
    __syncthreads();

    for (int it = 0; it < total + 2; it++) {
        // the segment walked in the previous iteration becomes reconstructed samples of the row buffer: Rx = x + candidate - NEAR
        if (it >= 1 && it <= total) {
            const int pit = it - 1, py = pit / nseg, psg = pit - py * nseg, px0 = psg * SEG, pn = min(SEG, W - px0);
            const uint8_t *in_prev = src + (size_t)py * W + px0;
            uint8_t *dst = sm.row[py & 1] + 8 + px0;
            const uint32_t *T = sm.T[pit & 1];
            for (int i = tid; i < pn; i += NT) {
                int k4 = (int)(T[i * RS + Cfg<NEAR>::SELW] & 0xFFu);
                if (k4 >= Cfg<NEAR>::RUN_OFS) k4 -= Cfg<NEAR>::RUN_OFS;
                const int rx = (int)in_prev[i] + (k4 >> 2) - NEAR;
                dst[i] = (uint8_t)rx;
                if (px0 + i == W - 1) dst[i + 1] = (uint8_t)rx;         // right neighbour of the last sample of the row below (:57-58)
            }
            __syncthreads();
        }
        // the first segment of a row needs the finished row above: its entries are made here, in front of the walk, by everybody
        if (it < total && it % nseg == 0) {
            stage_segment<NEAR>(sm, src, W, it / nseg, 0, min(SEG, W), it & 1, tid, NT);
            __syncthreads();
        }
        if (warp == 1 && it >= 2) {
            // ---- pack the segment before the previous one ----
            const int pit = it - 2, py = pit / nseg, psg = pit - py * nseg, pn = min(SEG, W - psg * SEG), pb = pit & 1;
            const uint32_t *codes = sm.codes[pb];
            const uint8_t *lens = sm.lens[pb];
            const int s0 = lane * 32;
            uint32_t l4[8], mylen = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                uint32_t v = 0;
                if (s0 + 4 * q < pn) v = *reinterpret_cast<const uint32_t *>(lens + tile::len_idx(s0) + 4 * q);
                // lengths past the segment end inside the last word are stale: mask them
                if (s0 + 4 * q + 4 > pn && s0 + 4 * q < pn) v &= 0xFFFFFFFFu >> (8 * (s0 + 4 * q + 4 - pn));
                l4[q] = v;
                mylen += (v & 0xFF) + ((v >> 8) & 0xFF) + ((v >> 16) & 0xFF) + (v >> 24);
            }
            uint32_t inc = mylen;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
            const uint32_t T = __shfl_sync(0xffffffffu, inc, 31);        // bits of the segment
            for (uint32_t w = lane; w < (T >> 5) + 2; w += 32) sm.ubits[w] = 0;
            __syncwarp();
            if (mylen) {
                const uint32_t bit = inc - mylen;
                uint32_t w = bit >> 5, nacc = bit & 31u;
                uint64_t acc = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) {
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const uint32_t len = (l4[q] >> (8 * r)) & 0xFF;
                        const uint32_t cd = codes[tile::code_idx(s0) + 4 * q + r];
                        acc = (acc << len) | (len ? (uint64_t)cd : 0ull);
                        nacc += len;
                        if (nacc >= 32) { nacc -= 32; atomicOr(&sm.ubits[w], (uint32_t)(acc >> nacc)); w++; }
                    }
                }
                if (nacc) atomicOr(&sm.ubits[w], (uint32_t)(acc << (32 - nacc)));
            }
            __syncwarp();
            tile::stuff_bits(sm.ubits, ps, T, pit == total - 1, lane);
        }
        if (warp == 2 && it >= 1 && it <= total) {
            // ---- coder: records of the previous segment -> code + length per sample ----
            const int pit = it - 1, py = pit / nseg, psg = pit - py * nseg, px0 = psg * SEG, pn = min(SEG, W - px0), pb = pit & 1;
            const uint32_t *rec = sm.T[pb] + Cfg<NEAR>::RECW;            // record of sample i at rec[i * RS]
            uint32_t *codes = sm.codes[pb];
            uint8_t *lens = sm.lens[pb];
            if (px0 == 0) run_len = 0;                                   // :270
            const uint32_t lt_mask = (1u << lane) - 1u;
            for (int base = 0; base < pn; base += 32) {
                const int i = base + lane;
                const bool valid = i < pn;
                const uint32_t r = valid ? rec[i * RS] : 0u;
                const uint32_t hi = r >> 16;
                const int e = (int)(short)(r & 0xFFFFu);
                const bool isreg = valid && (hi & 0x2000u) != 0 && (hi & REC_BREAK) == 0;
                const int q = (int)((hi >> 4) & 0x1FFu);
                // regular samples: the lanes of one context step its (A, N) together, in raster order (:365-383 without B, C)
                const uint32_t m = __match_any_sync(0xffffffffu, isreg ? q : 512 + lane);
                const int rank = __popc(m & lt_mask);
                int2 st = isreg ? sm.ctx2[q] : make_int2(0, 1);
                const int ae = abs(e);
                uint32_t mm = m;
                uint32_t code = 0; int len = 0;
                for (int t = 0; __any_sync(0xffffffffu, isreg && t <= rank); t++) {
                    const int srcl = mm ? __ffs((int)mm) - 1 : lane;
                    mm &= mm - 1u;
                    const int aet = __shfl_sync(0xffffffffu, ae, srcl);
                    if (isreg && t <= rank) {
                        if (t == rank) {
                            const int k = golomb_k(st.x, st.y);
                            const int me = 2 * ae - (e < 0 ? 1 : 0);     // NEAR > 0: no mapping by the bias (:367)
                            code = (uint32_t)me | ((uint32_t)k << 16);   // finished by the expansion below
                            len = 0xFF;
                        }
                        st.x += aet;
                        if (st.y >= 64) { st.x >>= 1; st.y >>= 1; }
                        st.y++;
                    }
                }
                if (isreg && rank == __popc(m) - 1) sm.ctx2[q] = st;
                // run samples carry no code unless lane 0 finds a completed block below
                if (valid) { codes[tile::code_idx(i)] = code; lens[tile::len_idx(i)] = (uint8_t)len; }
                // run samples and run interruptions: serial, on lane 0 (run_index, run length, the two interruption contexts);
                // a stretch of run samples is counted block by block (:295-300), not sample by sample
                uint32_t rm = __ballot_sync(0xffffffffu, valid && !isreg);
                const uint32_t hm = __ballot_sync(0xffffffffu, valid && hi == 0u);
                __syncwarp();
                if (lane == 0) {
                    while (rm) {
                        const int j = __ffs((int)rm) - 1;
                        if ((hm >> j) & 1u) {
                            // run samples j .. j + cnt - 1
                            const uint32_t rest = ~(hm >> j);
                            int cnt = rest ? __ffs((int)rest) - 1 : 32 - j;
                            rm &= cnt + j >= 32 ? 0u : ~0u << (j + cnt);
                            int pos = base + j;
                            const int end = pos + cnt;
                            while (pos < end) {
                                const int jj = tile::run_order(run_index);
                                const int need = (1 << jj) - run_len;
                                if (end - pos >= need) {                  // the block completes on sample pos + need - 1: a one (:296-299)
                                    pos += need;
                                    codes[tile::code_idx(pos - 1)] = 1u;
                                    lens[tile::len_idx(pos - 1)] = 1;
                                    run_len = 0;
                                    if (run_index < 31) run_index++;
                                } else {
                                    run_len += end - pos;
                                    pos = end;
                                }
                            }
                            if (px0 + end == W && run_len > 0) {          // the run reaches the end of the row: one more one (:302-303)
                                const uint32_t c0 = codes[tile::code_idx(end - 1)];
                                const int l0 = lens[tile::len_idx(end - 1)];
                                codes[tile::code_idx(end - 1)] = (c0 << 1) | 1u;
                                lens[tile::len_idx(end - 1)] = (uint8_t)(l0 + 1);
                            }
                            continue;
                        }
                        rm &= rm - 1u;
                        const int ii = base + j;
                        const uint32_t rr = rec[ii * RS];
                        const uint32_t h2 = rr >> 16;
                        const int e2 = (int)(short)(rr & 0xFFFFu);
                        const int t = (int)(h2 & 1u);
                        const int jj = tile::run_order(run_index);
                        const int glimit = P.limit - 1 - jj;
                        const int rem = run_len;
                        run_len = 0;
                        if (run_index > 0) run_index--;
                        int At = t ? rA1 : rA0, Bt = t ? rB1 : rB0, Nt = t ? rN1 : rN0;
                        const int k = golomb_k(At + (t ? (Nt >> 1) : 0), Nt);
                        const int map = (e2 != 0) && ((e2 > 0) == (k == 0 && 2 * Bt < Nt));
                        const int me = 2 * abs(e2) - t - map;
                        uint32_t gcode; int glen;
                        golomb_code(me, k, glimit, P.qbpp, gcode, glen);
                        codes[tile::code_idx(ii)] = ((uint32_t)rem << glen) | gcode;     // run remainder in J+1 bits, then the Golomb code (<= 32 bits)
                        lens[tile::len_idx(ii)] = (uint8_t)(jj + 1 + glen);
                        if (e2 < 0) Bt++;
                        At += (me + 1 - t) >> 1;
                        if (Nt >= 64) { At >>= 1; Bt >>= 1; Nt >>= 1; }
                        Nt++;
                        if (t) { rA1 = At; rB1 = Bt; rN1 = Nt; } else { rA0 = At; rB0 = Bt; rN0 = Nt; }
                    }
                }
                __syncwarp();
            }
            for (int i = lane; i < pn; i += 32) {
                if (lens[tile::len_idx(i)] != 0xFF) continue;            // run samples carry their final code already (<= 32 bits)
                const uint32_t mk = codes[tile::code_idx(i)];
                uint32_t code; int len;
                golomb_code((int)(mk & 0xFFFFu), (int)(mk >> 16), P.limit, P.qbpp, code, len);
                codes[tile::code_idx(i)] = code;
                lens[tile::len_idx(i)] = (uint8_t)len;
            }
        }
        if (warp >= 3 && it + 1 < total && (it + 1) % nseg != 0) {
            // ---- the entries of the next segment of the same row ----
            const int nit = it + 1, ny = nit / nseg, nsg = nit - ny * nseg, nx0 = nsg * SEG;
            stage_segment<NEAR>(sm, src, W, ny, nx0, min(SEG, W - nx0), nit & 1, tid - 96, STAGE_WARPS * 32);
        }
        if (tid == 0 && it < total) {
            // ---- the walk ----
            const int x0 = (it % nseg) * SEG;
            char *tp = reinterpret_cast<char *>(sm.T[it & 1]);
            // state offsets carry the 0x2000 flag of the record's high half: taken off the base once
            char *ctxb = reinterpret_cast<char *>(sm.ctx) - 0x2000;
            const char *dqb = reinterpret_cast<const char *>(sm.dq2);
            if (x0 == 0) sel = NEAR * 4;                                 // row start (:270): not in a run; every candidate entry of column 0 is the same
            uint32_t w = *reinterpret_cast<const uint32_t *>(tp + sel);
            for (;;) {
                // regular samples (:346-394 without A, k and the code), as long as the next entry is one: the loop's only branch
#pragma unroll 2
                while ((w & 0x07000000u) == 0u) {
                    int4 *sp = reinterpret_cast<int4 *>(ctxb + (w & 0xFFFFu));
                    int4 st = *sp;
                    const int sgn = (int)w >> 31;
                    const int medn = (int)((w >> 16) & 0xFFu) + (int)(w >> 31);            // MED - sign
                    const uint32_t xo = *reinterpret_cast<const uint32_t *>(tp + Cfg<NEAR>::XW * 4) | ((uint32_t)sgn & 2048u);   // dq2[x + 255 | negative << 9]
                    const int hv = st.z >> 6;                                               // N >= 64: halve (:378-383; N never exceeds 64)
                    const int nn = (st.z >> hv) + 1;
                    int px = medn + (st.y ^ sgn);
                    px = min(max(px, 0), 255);
                    const uint32_t v = *reinterpret_cast<const uint32_t *>(dqb + xo - 4 * px);
                    sel = v >> 16;
                    const uint32_t wn = *reinterpret_cast<const uint32_t *>(tp + RS * 4 + sel);          // the next entry, before the state update
                    const int e = (int)(short)(v & 0xFFFFu);
                    const int bh = (st.x + e * quant) >> hv;                                // :376-383
                    // bias update :384-393 without branches: down = (B <= -N), up = (B > 0) as sign masks
                    const int sdn = (bh + nn - 1) >> 31, sup = (-bh) >> 31;
                    const int dc = sdn - sup;                                               // -1 / 0 / +1
                    st.x = min(max(bh - nn * dc, 1 - nn), 0);
                    st.y = min(max(st.y + dc, -128), 127);
                    st.z = nn;
                    *sp = st;
                    *reinterpret_cast<uint32_t *>(tp + Cfg<NEAR>::RECW * 4) = pack_halves(v, w);
                    *reinterpret_cast<uint8_t *>(tp + Cfg<NEAR>::SELW * 4) = (uint8_t)sel;
                    tp += RS * 4;
                    w = wn;
                }
                if (w & 0x04000000u) break;                              // the row after the last sample
                uint32_t r;
                if ((w & 0x01000000u) == 0u) {
                    // a regular sample within NEAR of 0 or 255: as above, and the reconstruction is clamped to 0..255 (:321-323)
                    int4 *sp = reinterpret_cast<int4 *>(ctxb + (w & 0xFFFFu));
                    int4 st = *sp;
                    const int sgn = (int)w >> 31;
                    const uint32_t xw = *reinterpret_cast<const uint32_t *>(tp + Cfg<NEAR>::XW * 4);
                    const int xv = (int)(xw >> 2) - 255;
                    int px = (int)((w >> 16) & 0xFFu) + ((st.y ^ sgn) - sgn);
                    px = min(max(px, 0), 255);
                    const uint32_t v = *reinterpret_cast<const uint32_t *>(dqb + (xw | ((uint32_t)sgn & 2048u)) - 4 * px);
                    const int e = (int)(short)(v & 0xFFFFu);
                    const int lo = (max(-xv, -NEAR) + NEAR) * 4, hi = (min(255 - xv, NEAR) + NEAR) * 4;
                    sel = (uint32_t)min(max((int)(v >> 16), lo), hi);
                    st.x += e * quant;
                    if (st.z >= 64) { st.x >>= 1; st.z >>= 1; }
                    st.z++;
                    if (st.x <= -st.z) {
                        st.x += st.z; if (st.x <= -st.z) st.x = -st.z + 1;
                        st.y = max(st.y - 1, -128);
                    } else if (st.x > 0) {
                        st.x -= st.z; if (st.x > 0) st.x = 0;
                        st.y = min(st.y + 1, 127);
                    }
                    *sp = st;
                    r = (v & 0xFFFFu) | (w << 16);
                } else if ((w & 0x02000000u) == 0u) {                    // run sample: counted by the coder
                    sel = w & 0xFFFFu;
                    r = 0;
                } else {                                                 // run interruption
                    sel = w & 0x3Cu;
                    r = ((uint32_t)((int)(w << 17) >> 23) & 0xFFFFu) | ((REC_BREAK | ((w >> 15) & 1u)) << 16);
                }
                *reinterpret_cast<uint32_t *>(tp + Cfg<NEAR>::RECW * 4) = r;
                *reinterpret_cast<uint8_t *>(tp + Cfg<NEAR>::SELW * 4) = (uint8_t)sel;
                tp += RS * 4;
                w = *reinterpret_cast<const uint32_t *>(tp + sel);
            }
        }
        __syncthreads();
    }
    if (tid == 32) plane_bytes[p] = ps.out_pos;
}

}  // namespace nearscan
}  // namespace jls
}  // namespace imcvt
