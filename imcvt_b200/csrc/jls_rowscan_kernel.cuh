// jls_rowscan_kernel.cuh -- JPEG-LS scan encoder for any NEAR on sm_100a, the path for NEAR > 0 (reference: JLSencodeScan,
// src/imageio_jls.c:240-399): one CTA of three warps per plane, a row segment apart (stage / walk | code | pack).
// Included by jls.cu; tests/emu/jls_rowscan_emu.cpp compiles the same source for the host under the SIMT emulator.
#pragma once
#include "jls_tile_kernel.cuh"

#ifdef IMCVT_HOST_EMU
#define JLS_ROWSCAN_SMEM() (emu::dyn_smem())
#else
#define JLS_ROWSCAN_SMEM() (jls_dyn_smem)           /* the dynamic shared memory window declared by jls_tile_kernel.cuh */
#endif

namespace imcvt {
namespace jls {

// setParameters :26-38 for 8-bit samples, as the kernels take it
struct SerialParams { int near_, t1, t2, t3, quant, range, qbpp, limit, a_init; };

inline SerialParams make_params(int near_)               // setParameters :26-38 for bpp = 8
{
    SerialParams P;
    P.near_ = near_;
    P.t1 = 3 + 3 * near_; P.t2 = 7 + 5 * near_; P.t3 = 21 + 7 * near_;
    P.quant = 2 * near_ + 1;
    P.range = (256 + 4 * near_) / P.quant;
    P.qbpp = 1; while ((1 << P.qbpp) < P.range) P.qbpp++;
    P.limit = 32 - P.qbpp - 1;
    P.a_init = (P.range + 32) / 64; if (P.a_init < 2) P.a_init = 2;
    return P;
}

__device__ __forceinline__ int grad_q_any(const SerialParams &P, int g)
{
    const int m = abs(g);
    const int q = (m > P.near_) + (m >= P.t1) + (m >= P.t2) + (m >= P.t3);
    return g < 0 ? -q : q;
}

// ------------------------------------------------------------------------------------------------
// Row-pipelined scan: any NEAR, the path for NEAR > 0.  With NEAR > 0 the reconstructed value of a sample feeds the
// context of the next one (:321-323), so a plane is ONE dependency chain; everything that is not on that chain is taken
// off the thread that walks it.  Three warps per plane, each a segment (<= 2048 samples of a row) behind the previous one:
//   warp 0       all lanes stage the input and the two gradient terms that only need the row above, 81*q(d-b) + 9*q(b-c),
//                from the reconstructed row kept in shared memory.  Then lane 0 walks the segment with ONLY the
//                recurrence: third gradient by table, run state, prediction with the bias C of the context, error
//                quantisation by table (no divisions), reconstruction, and the B / C / N update the next prediction of
//                that context needs (:346-394 without A, k and the code).  It leaves one record per sample.
//   warp 2       lane 0 turns the previous segment's records into codes: A and N per context (N evolves the same way in
//                both stages), Golomb parameter, mapping, code + length; run-length counting and the run-interruption
//                samples with their own A / B / N (:291-344) live here too -- none of them feeds the reconstruction.
//   warp 1       packs the codes of the segment before that (register windows, tile-wide bit buffer) and turns them into
//                0xFF-stuffed bytes with the packer of the tile kernel.
// ------------------------------------------------------------------------------------------------
namespace rowscan {
constexpr int SEG = 2048;
constexpr int WMAXR = 32768;                        // :436
constexpr int NT = 96;
// One 32-bit slot per sample of a segment, used twice.  Staged (all lanes of warp 0):
//     sample | (81*q(d-b) + 9*q(b-c) + 512) << 8 | d << 24            (everything the walk needs that is known beforehand)
// and, once lane 0 has walked the sample, its record for stage 2:
//     low half: wrapped error (signed) | high half: 0 = run sample, 0x8000 | RItype = run interruption,
//     else context number 1..364 (| 1 << 9: the mapping hint 2B <= -N of the lossless instance, :367)
constexpr uint32_t REC_BREAK = 0x8000u;
struct Smem {
    uint8_t row[2][WMAXR + 16];                     // reconstructed rows y-1 / y (index 8 + x)
    uint32_t rec[2][SEG];
    uint32_t codes[2][SEG + SEG / 32];
    uint8_t lens[2][SEG + SEG / 8];
    uint32_t ubits[tile::UBITS_WORDS];
    int4 ctx[NCTX];                                 // stage 1: (unused, B, C, N)
    int2 ctx2[NCTX];                                // stage 2: (A, N)
    signed char gq[512];                            // quantised gradient at [g + 255]
    short eq[512];                                  // quantised prediction error at [d + 255] (:97-102), wrapped into the code range (:105-111)
    short qe[512];                                  // the reconstruction step it stands for: quant * (quantised error, not wrapped)
};

// LOSSY = NEAR > 0 (the reconstruction differs from the input); the lossless instance only serves the forced-serial cross-check
template <bool LOSSY>
__global__ void __launch_bounds__(NT, 1)
rowscan_kernel(const uint8_t *__restrict__ pix, size_t plane_stride, int W, int H, SerialParams P, FrameLayout L,
               uint64_t *__restrict__ plane_bytes)
{
    Smem &sm = *reinterpret_cast<Smem *>(JLS_ROWSCAN_SMEM());
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = blockIdx.x;
    const uint8_t *src = pix + (size_t)p * plane_stride;
    const int near_ = P.near_, quant = P.quant, range = P.range;
    const int half = (range + 1) / 2;

    for (int k = tid; k < NCTX; k += NT) { sm.ctx[k] = make_int4(0, 0, 0, 1); sm.ctx2[k] = make_int2(P.a_init, 1); }
    for (int i = tid; i < 511; i += NT) {
        const int g = i - 255;
        sm.gq[i] = (signed char)grad_q_any(P, g);
        int e = g < 0 ? -((near_ - g) / quant) : (near_ + g) / quant;
        sm.qe[i] = (short)(quant * e);
        if (e < 0) e += range;
        if (e >= half) e -= range;
        sm.eq[i] = (short)e;
    }
    for (int i = tid; i < WMAXR + 16; i += NT) { sm.row[0][i] = 0; sm.row[1][i] = 0; }      // the row above the picture reads as zeros

    tile::PackState ps;
    {
        size_t cap;
        uint8_t *dst = plane_dst(L, p, cap);
        tile::pack_state_init(ps, dst, cap, nullptr);
    }

    const int nseg = (W + SEG - 1) / SEG, total = H * nseg;
    // stage 1 (lane 0 of warp 0): row state of the walk
    int in_run = 0;                                 // an int: a bool would be kept as a byte of a shared register
    int a = 0, b = 0, c = 0, up2 = 0;
    // stage 2 (lane 0 of warp 2): run state and the two run-interruption contexts
    int rA0 = P.a_init, rA1 = P.a_init, rB0 = 0, rB1 = 0, rN0 = 1, rN1 = 1;      // scalars: an indexed array would live in local memory
    int run_index = 0, run_len = 0;
    __syncthreads();

    for (int it = 0; it < total + 2; it++) {
        if (warp == 1 && it >= 2) {
            // ---- pack the segment before the previous one ----
            const int pit = it - 2, py = pit / nseg, psg = pit - py * nseg, pn = min(SEG, W - psg * SEG), pb = pit & 1;
            const uint32_t *codes = sm.codes[pb];
            const uint8_t *lens = sm.lens[pb];
            uint32_t T = 0;                                              // bits of the segment
            for (int c0 = 0; c0 < pn; c0 += 1024) {
                const int s0 = c0 + lane * 32;
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
                const uint32_t ctot = __shfl_sync(0xffffffffu, inc, 31);
                for (uint32_t w = (T >> 5) + 1 + lane; w < ((T + ctot) >> 5) + 2; w += 32) sm.ubits[w] = 0;
                if (T == 0 && lane == 0) sm.ubits[0] = 0;
                __syncwarp();
                if (mylen) {
                    const uint32_t bit = T + inc - mylen;
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
                T += ctot;
            }
            tile::stuff_bits(sm.ubits, ps, T, pit == total - 1, lane);
        }
        if (warp == 2 && it >= 1 && it <= total) {
            // ---- stage 2: records of the previous segment -> code + length per sample.  Lane 0 walks what is serial (A / N
            //      per context -> Golomb parameter and mapped error; run lengths; run interruptions), then the whole warp
            //      turns the (mapped error, k) pairs of the regular samples into codes ----
            const int pit = it - 1, py = pit / nseg, psg = pit - py * nseg, px0 = psg * SEG, pn = min(SEG, W - px0), pb = pit & 1;
            const uint32_t *rec = sm.rec[pb];
            uint32_t *codes = sm.codes[pb];
            uint8_t *lens = sm.lens[pb];
            if (lane == 0) {
            if (px0 == 0) run_len = 0;                                   // :270
            for (int i = 0; i < pn; i++) {
                const uint32_t r = rec[i];
                const uint32_t hi = r >> 16;
                const int e = (int)(short)(r & 0xFFFFu);
                uint32_t code = 0; int len = 0;
                if (hi != 0 && !(hi & REC_BREAK)) {
                    const int q = (int)(hi & 511u) - 1;
                    int2 st = sm.ctx2[q];
                    const int k = golomb_k(st.x, st.y);
                    const int map = (k == 0) && ((hi >> 9) & 1u);        // the hint is (2B <= -N) && NEAR == 0 (:367)
                    const int me = 2 * abs(e) + (e < 0 ? -(map + 1) : map);
                    code = (uint32_t)me | ((uint32_t)k << 16);           // finished by the warp below
                    len = 0xFF;
                    st.x += abs(e);
                    if (st.y >= 64) { st.x >>= 1; st.y >>= 1; }
                    st.y++;
                    sm.ctx2[q] = st;
                } else if (hi == 0) {
                    run_len++;
                    const int jj = tile::run_order(run_index);
                    if (run_len >= (1 << jj)) {
                        code = 1; len = 1;
                        run_len -= 1 << jj;
                        if (run_index < 31) run_index++;
                    }
                    if (px0 + i == W - 1 && run_len > 0) { code = (code << 1) | 1u; len++; }
                } else {
                    const int t = (int)(hi & 1u);
                    const int jj = tile::run_order(run_index);
                    const int glimit = P.limit - 1 - jj;
                    const int rem = run_len;
                    run_len = 0;
                    if (run_index > 0) run_index--;
                    int At = t ? rA1 : rA0, Bt = t ? rB1 : rB0, Nt = t ? rN1 : rN0;
                    const int k = golomb_k(At + (t ? (Nt >> 1) : 0), Nt);
                    const int map = (e != 0) && ((e > 0) == (k == 0 && 2 * Bt < Nt));
                    const int me = 2 * abs(e) - t - map;
                    uint32_t gcode; int glen;
                    golomb_code(me, k, glimit, P.qbpp, gcode, glen);
                    code = ((uint32_t)rem << glen) | gcode;              // run remainder in J+1 bits, then the Golomb code (<= 32 bits)
                    len = jj + 1 + glen;
                    if (e < 0) Bt++;
                    At += (me + 1 - t) >> 1;
                    if (Nt >= 64) { At >>= 1; Bt >>= 1; Nt >>= 1; }
                    Nt++;
                    if (t) { rA1 = At; rB1 = Bt; rN1 = Nt; } else { rA0 = At; rB0 = Bt; rN0 = Nt; }
                }
                codes[tile::code_idx(i)] = code;
                lens[tile::len_idx(i)] = (uint8_t)len;
            }
            }
            __syncwarp();
            for (int i = lane; i < pn; i += 32) {
                if (lens[tile::len_idx(i)] != 0xFF) continue;            // run samples carry their final code already (<= 32 bits)
                const uint32_t mk = codes[tile::code_idx(i)];
                uint32_t code; int len;
                golomb_code((int)(mk & 0xFFFFu), (int)(mk >> 16), P.limit, P.qbpp, code, len);
                codes[tile::code_idx(i)] = code;
                lens[tile::len_idx(i)] = (uint8_t)len;
            }
        }
        if (warp == 0 && it < total) {
            // ---- stage the segment (input + everything that only needs the row above), then stage 1 on lane 0 ----
            const int y = it / nseg, sg = it - y * nseg, x0 = sg * SEG, n = min(SEG, W - x0);
            const uint8_t *up = sm.row[(y + 1) & 1] + 8, *in_row = src + (size_t)y * W + x0;
            uint32_t *rec = sm.rec[it & 1];
            for (int i = lane; i < n; i += 32) {
                const int x = x0 + i;
                const int bb = up[x], cc = x > 0 ? up[x - 1] : 0, dd = up[x + 1];      // up[W] repeats up[W - 1] (:57-58)
                const int q12 = 81 * sm.gq[dd - bb + 255] + 9 * sm.gq[bb - cc + 255];
                rec[i] = (uint32_t)in_row[i] | ((uint32_t)(q12 + 512) << 8) | ((uint32_t)dd << 24);
            }
            __syncwarp();
            if (lane == 0) {
                uint8_t *curp = sm.row[y & 1] + 8 + x0;
                if (x0 == 0) {                                           // row start (:270, :47-56)
                    in_run = 0;
                    b = up[0]; a = b; c = up2;
                    up2 = b;                                             // recon(y-1, 0) is the c of row y+1's first sample
                    // the first sample of a row has its own c (two rows up): its staged gradient terms are replaced
                    const int q12 = 81 * sm.gq[up[1] - b + 255] + 9 * sm.gq[b - c + 255];
                    rec[0] = (rec[0] & ~(0x3FFu << 8)) | ((uint32_t)(q12 + 512) << 8);
                }
                uint32_t *slot = rec;
                for (int i = 0; i < n; i++, slot++) {
                    const uint32_t w = *slot;
                    const int v = (int)(w & 0xFFu), d = (int)(w >> 24);
                    const int qv = (int)((w >> 8) & 0x3FFu) - 512 + sm.gq[c - a + 255];
                    int rx;
                    uint32_t r;
                    if ((in_run | (qv == 0)) != 0) {
                        if (abs(v - a) <= near_) {                       // run sample (:291-303): counted by stage 2
                            in_run = 1;
                            rx = a;
                            r = 0;
                        } else {                                         // run interruption (:305-344)
                            in_run = 0;
                            const int t = abs(a - b) <= near_;
                            const bool neg = a > b + near_;
                            const int px = t ? a : b;
                            const int di = (neg ? px - v : v - px) + 255;
                            const int e = sm.eq[di], step = sm.qe[di];
                            rx = LOSSY ? min(max(px + (neg ? -step : step), 0), 255) : v;
                            r = ((uint32_t)e & 0xFFFFu) | ((REC_BREAK | (uint32_t)t) << 16);
                        }
                    } else {
                        const bool neg = qv < 0;
                        const int cq = neg ? -qv : qv;
                        int4 st = sm.ctx[cq - 1];
                        int px = med3(a, b, c) + (neg ? -st.z : st.z);
                        px = min(max(px, 0), 255);
                        const int di = (neg ? px - v : v - px) + 255;
                        const int e = sm.eq[di], step = sm.qe[di];
                        rx = LOSSY ? min(max(px + (neg ? -step : step), 0), 255) : v;
                        const uint32_t hint = !LOSSY && (2 * st.y <= -st.w);
                        r = ((uint32_t)e & 0xFFFFu) | (((uint32_t)cq | (hint << 9)) << 16);
                        st.y += e * quant;
                        if (st.w >= 64) { st.y >>= 1; st.w >>= 1; }
                        st.w++;
                        if (st.y <= -st.w) {
                            st.y += st.w; if (st.y <= -st.w) st.y = -st.w + 1;
                            st.z = max(st.z - 1, -128);
                        } else if (st.y > 0) {
                            st.y -= st.w; if (st.y > 0) st.y = 0;
                            st.z = min(st.z + 1, 127);
                        }
                        sm.ctx[cq - 1] = st;
                    }
                    *slot = r;
                    curp[i] = (uint8_t)rx;
                    c = b; b = d; a = rx;                                 // slide the template (a = reconstructed left neighbour)
                }
                if (x0 + n == W) curp[n] = (uint8_t)a;                   // right neighbour of the last sample of the row below
            }
        }
        __syncthreads();
    }
    if (tid == 32) plane_bytes[p] = ps.out_pos;
}
}  // namespace rowscan

}  // namespace jls
}  // namespace imcvt
