/* oracle/jls_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference JPEG-LS (ITU-T T.87 baseline, 8-bit) scan encoder,
 * written from the reference's behaviour (file:line citations are to /root/reference/src/imageio_jls.c).
 * It is the *checker* for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it.  The product library (imcvt_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_jls.py checks this file byte-for-byte against
 *   (a) the known-answer table of SURVEY.md 8(c) (reference outputs for image/P2..P6.pnm, NEAR 0..4), and
 *   (b) oracle/_ref/libimcvt_ref.so = the unmodified reference sources compiled by oracle/Makefile,
 *       on random / edge-case images.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* ---- coding parameters for 8-bit samples (setParameters :26-38, getQbpp :18-23) ---- */
typedef struct {
    int near, t1, t2, t3;   /* gradient thresholds */
    int quant;              /* 2*NEAR+1 */
    int range;              /* "qbeta" in the reference = quantised RANGE */
    int qbpp;               /* bits needed for range */
    int limit;              /* LIMIT - qbpp - 1 : length of the unary prefix before escape */
    int a_init;
} jls_params;

static jls_params jls_make_params(int near)
{
    jls_params p;
    p.near  = near;
    p.t1    = 3 + 3 * near;              /* :30 with bpp=8 -> tmp=1 */
    p.t2    = 7 + 5 * near;              /* :31 */
    p.t3    = 21 + 7 * near;             /* :32 */
    p.quant = 2 * near + 1;              /* :33 */
    p.range = (256 + 4 * near) / p.quant;/* :34 */
    p.qbpp  = 1;
    while ((1 << p.qbpp) < p.range) p.qbpp++;   /* :18-23 */
    p.limit = 32 - p.qbpp - 1;           /* :36 */
    p.a_init = (p.range + 32) / 64;      /* :37 */
    if (p.a_init < 2) p.a_init = 2;
    return p;
}

/* ---- MSB-first bit sink with the T.87 marker-avoidance rule (writeBit :156-169, flushBits :177-184) ---- */
typedef struct {
    uint8_t *dst;
    size_t   n;        /* bytes emitted */
    uint32_t cur;      /* bits collected for the current byte (right aligned) */
    int      room;     /* payload bits still free in the current byte: 8, or 7 after an 0xFF byte */
    int      cap;      /* capacity the current byte started with */
} bitsink;

static void sink_init(bitsink *s, uint8_t *dst) { s->dst = dst; s->n = 0; s->cur = 0; s->room = 8; s->cap = 8; }

static void sink_put(bitsink *s, uint32_t value, int nbits)
{
    while (nbits > 0) {
        int take = nbits < s->room ? nbits : s->room;
        uint32_t chunk = (value >> (nbits - take)) & ((1u << take) - 1u);
        s->cur = (s->cur << take) | chunk;
        s->room -= take;
        nbits -= take;
        if (s->room == 0) {
            uint8_t byte = (uint8_t)s->cur;               /* cap==7: top bit is the stuffed 0 */
            s->dst[s->n++] = byte;
            s->cap = s->room = (byte == 0xFF) ? 7 : 8;    /* :164-165 */
            s->cur = 0;
        }
    }
}

static void sink_zeros_then_one(bitsink *s, int zeros)
{
    while (zeros > 24) { sink_put(s, 0, 24); zeros -= 24; }
    sink_put(s, 1, zeros + 1);
}

static void sink_flush(bitsink *s)                        /* :177-184 */
{
    if (s->room < 8) {                                    /* partial byte, or the empty 7-bit byte after 0xFF */
        s->dst[s->n++] = (uint8_t)(s->cur << s->room);
        s->cur = 0; s->room = s->cap = 8;
    }
}

static void sink_bytes(bitsink *s, uint32_t v, int nbytes)  /* writeValue :147-153 (byte aligned only) */
{
    for (int i = nbytes - 1; i >= 0; i--) s->dst[s->n++] = (uint8_t)(v >> (8 * i));
}

/* ---- small helpers ---- */
static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline int grad_q(const jls_params *p, int g)       /* gradientQuantize :68-76 */
{
    int m = iabs(g), q;
    if      (m >= p->t3)  q = 4;
    else if (m >= p->t2)  q = 3;
    else if (m >= p->t1)  q = 2;
    else if (m >  p->near) q = 1;
    else                  q = 0;
    return g < 0 ? -q : q;
}

static inline int med_predict(int a, int b, int c)         /* predict :87-94 */
{
    int mx = a > b ? a : b, mn = a < b ? a : b;
    if (c >= mx) return mn;
    if (c <= mn) return mx;
    return a + b - c;
}

static inline int err_quant(const jls_params *p, int e)    /* quantize :97-102 */
{
    return e < 0 ? -((p->near - e) / p->quant) : (p->near + e) / p->quant;
}

static inline int err_wrap(const jls_params *p, int e)     /* modRange :105-111 */
{
    if (e < 0) e += p->range;
    if (e >= (p->range + 1) / 2) e -= p->range;
    return e;
}

static inline int golomb_k(int A, int N)                   /* getK :114-121 (ritype folded by caller) */
{
    int k = 0;
    while ((N << k) < A) k++;
    return k;
}

static void golomb_put(bitsink *s, const jls_params *p, int limit, int v, int k)   /* GolombCoding :187-197 */
{
    int hi = v >> k;
    if (hi < limit) {
        sink_zeros_then_one(s, hi);
        if (k) sink_put(s, (uint32_t)v & ((1u << k) - 1u), k);
    } else {
        sink_zeros_then_one(s, limit);
        sink_put(s, (uint32_t)(v - 1) & ((1u << p->qbpp) - 1u), p->qbpp);
    }
}

static const uint8_t RUN_J[32] = {0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,5,5,6,6,7,7,8,9,10,11,12,13,14,15};  /* :15 */

/* ---- one scan over one plane (JLSencodeScan :240-399). `plane` holds the samples on entry and the
 *      reconstruction on exit (the reference also encodes in place, :406). ---- */
static void jls_scan(bitsink *s, int near, int H, int W, int *plane)
{
    const jls_params P = jls_make_params(near);
    int A[364], B[364], C[364], N[364];
    int rA[2], rB[2], rN[2];
    int run_index = 0;                                      /* persists across rows (:250) */

    for (int q = 0; q < 364; q++) { A[q] = P.a_init; B[q] = 0; C[q] = 0; N[q] = 1; }      /* :256-261 */
    for (int t = 0; t < 2; t++)   { rA[t] = P.a_init; rB[t] = 0; rN[t] = 1; }              /* :263-267 */

    for (int y = 0; y < H; y++) {
        int *cur = plane + (size_t)y * W;
        const int *up  = y > 0 ? cur - W : NULL;
        const int *up2 = y > 1 ? cur - 2 * (size_t)W : NULL;
        int in_run = 0, run_len = 0;                        /* reset per row (:270) */

        for (int x = 0; x < W; x++) {
            /* causal neighbourhood from the reconstruction (samplePixels :46-65) */
            int b = up ? up[x] : 0;
            int d = up ? (x + 1 < W ? up[x + 1] : b) : 0;
            int a, c;
            if (x == 0) { a = b; c = up2 ? up2[0] : 0; }
            else        { a = cur[x - 1]; c = up ? up[x - 1] : 0; }
            const int v = cur[x];

            int qv = 81 * grad_q(&P, d - b) + 9 * grad_q(&P, b - c) + grad_q(&P, c - a);  /* getQ :79-84 */
            int sgn = qv < 0 ? -1 : 1;
            int ctx = iabs(qv);

            if (ctx == 0) in_run = 1;                       /* :284 */
            int hit = 0, brk = 0;
            if (in_run) {                                   /* :286-289 */
                hit = iabs(v - a) <= near;
                brk = !hit;
                in_run = hit;
            }

            if (hit) {                                      /* run continues (:291-303) */
                cur[x] = a;
                run_len++;
                if (run_len >= (1 << RUN_J[run_index])) {
                    sink_put(s, 1, 1);
                    run_len -= 1 << RUN_J[run_index];
                    if (run_index < 31) run_index++;
                }
                if (x == W - 1 && run_len > 0) sink_put(s, 1, 1);
            } else if (brk) {                               /* run interruption (:305-344) */
                int j = RUN_J[run_index];
                int glimit = P.limit - 1 - j;
                sink_put(s, (uint32_t)run_len, j + 1);
                run_len = 0;
                if (run_index > 0) run_index--;

                int t = iabs(a - b) <= near;                /* RItype */
                sgn = (a > b + near) ? -1 : 1;
                int px = t ? a : b;
                int e = err_quant(&P, sgn * (v - px));
                cur[x] = near ? clampi(px + sgn * P.quant * e, 0, 255) : v;
                e = err_wrap(&P, e);
                int k = golomb_k(rA[t] + (t ? (rN[t] >> 1) : 0), rN[t]);
                int map = (e != 0) && ((e > 0) == (k == 0 && 2 * rB[t] < rN[t]));          /* :331 */
                int me = 2 * iabs(e) - t - map;
                golomb_put(s, &P, glimit, me, k);
                if (e < 0) rB[t]++;
                rA[t] += (me + 1 - t) >> 1;
                if (rN[t] >= 64) { rA[t] >>= 1; rB[t] >>= 1; rN[t] >>= 1; }
                rN[t]++;
            } else {                                        /* regular mode (:346-394) */
                run_len = 0;
                int q = ctx - 1;
                int px = clampi(med_predict(a, b, c) + sgn * C[q], 0, 255);
                int e = err_quant(&P, sgn * (v - px));
                cur[x] = near ? clampi(px + sgn * P.quant * e, 0, 255) : v;
                e = err_wrap(&P, e);
                int k = golomb_k(A[q], N[q]);
                int map = (k == 0) && (2 * B[q] <= -N[q]) && (near == 0);                  /* :367 */
                int me = 2 * iabs(e);
                if (e < 0) me -= map + 1; else me += map;
                golomb_put(s, &P, P.limit, me, k);

                B[q] += e * P.quant;                        /* :376 */
                A[q] += iabs(e);
                if (N[q] >= 64) { A[q] >>= 1; B[q] >>= 1; N[q] >>= 1; }
                N[q]++;
                if (B[q] <= -N[q]) {
                    B[q] += N[q]; if (B[q] <= -N[q]) B[q] = -N[q] + 1;
                    C[q]--;
                } else if (B[q] > 0) {
                    B[q] -= N[q]; if (B[q] > 0) B[q] = 0;
                    C[q]++;
                }
                C[q] = clampi(C[q], -128, 127);
            }
        }
    }
    sink_flush(s);
}

/* ---- whole-image encode (headers :200-237, JLSencodeImageGray :402-409, JLSencodeImageRGB :412-423,
 *      widening / de-interleave of writeJLSImageFile :447-459) ----
 * pixels: interleaved u8, H*W (gray) or H*W*3 (RGB).  Returns the stream length, 0 on bad arguments. */
size_t oracle_jls_encode(const uint8_t *pixels, int is_rgb, uint32_t H, uint32_t W, int near,
                         uint8_t *out, size_t out_cap)
{
    if (W < 1 || W > 32767 || H < 1 || H > 32767) return 0;      /* :436 */
    const size_t n = (size_t)H * W;
    const int ncomp = is_rgb ? 3 : 1;
    uint8_t *tmp = (uint8_t *)malloc(8 * n + 65536);             /* reference's own bound (:439) */
    int *plane = (int *)malloc(n * sizeof(int));
    if (!tmp || !plane) { free(tmp); free(plane); return 0; }

    bitsink s; sink_init(&s, tmp);
    sink_bytes(&s, 0xFFD8, 2);                                   /* SOI */
    sink_bytes(&s, 0xFFF7, 2);                                   /* SOF55 */
    sink_bytes(&s, 8 + 3 * ncomp, 2);
    sink_bytes(&s, 8, 1);
    sink_bytes(&s, H, 2);
    sink_bytes(&s, W, 2);
    sink_bytes(&s, ncomp, 1);
    for (int c = 1; c <= ncomp; c++) { sink_bytes(&s, c, 1); sink_bytes(&s, 0x11, 1); sink_bytes(&s, 0, 1); }

    for (int c = 0; c < ncomp; c++) {
        for (size_t i = 0; i < n; i++) plane[i] = pixels[i * ncomp + c];
        sink_bytes(&s, 0xFFDA, 2);                               /* SOS (:200-208) */
        sink_bytes(&s, 8, 2);
        sink_bytes(&s, 1, 1);
        sink_bytes(&s, c + 1, 1);
        sink_bytes(&s, 0, 1);
        sink_bytes(&s, near, 1);
        sink_bytes(&s, 0, 2);
        jls_scan(&s, near, (int)H, (int)W, plane);
    }
    sink_bytes(&s, 0xFFD9, 2);                                   /* EOI */

    size_t len = s.n;
    if (len > out_cap) len = 0; else memcpy(out, tmp, len);
    free(tmp); free(plane);
    return len;
}
