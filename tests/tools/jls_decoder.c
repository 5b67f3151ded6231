/* jls_decoder.c -- TEST INFRASTRUCTURE ONLY: a small stand-alone JPEG-LS (ITU-T T.87) baseline decoder for the streams
 * this repository produces: 8-bit, NEAR 0..255, default thresholds (no LSE), one component per scan (gray, or RGB as three
 * non-interleaved scans).  Written from the standard's decoding procedures (Annex A.7 / F), not from the reference
 * (which has no decoder): it is the independent end of the encode -> decode round-trip tests, also at BASELINE sizes.
 *
 *   long jls_decode(const uint8_t *s, size_t n, uint8_t *out, int *w, int *h, int *ncomp, int *near)
 *       out: ncomp planes of h*w bytes, plane after plane; returns 0 on success, a negative code on a malformed stream.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int J[32] = {0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,5,5,6,6,7,7,8,9,10,11,12,13,14,15};

typedef struct { const uint8_t *p, *end; uint32_t acc; int nacc; int after_ff; int err; } Bits;

static int bit(Bits *b)
{
    if (b->nacc == 0) {
        if (b->p >= b->end) { b->err = 1; return 0; }
        uint8_t v = *b->p++;
        if (b->after_ff) { b->acc = v & 0x7F; b->nacc = 7; }        /* the bit after a 0xFF byte is stuffed */
        else { b->acc = v; b->nacc = 8; }
        b->after_ff = (v == 0xFF);
    }
    b->nacc--;
    return (b->acc >> b->nacc) & 1;
}
static int bits(Bits *b, int n) { int v = 0; while (n-- > 0) v = (v << 1) | bit(b); return v; }

static int golomb(Bits *b, int k, int limit, int qbpp)
{
    int q = 0;
    while (!bit(b)) { if (++q > 64 || b->err) { b->err = 1; return 0; } }
    if (q < limit) return (q << k) | bits(b, k);
    return bits(b, qbpp) + 1;
}
static int kparam(int A, int N) { int k = 0; while ((N << k) < A) k++; return k; }
static int gquant(int g, int near, int t1, int t2, int t3)
{
    int m = g < 0 ? -g : g, q = (m > near) + (m >= t1) + (m >= t2) + (m >= t3);
    return g < 0 ? -q : q;
}
static int clamp255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

static int decode_scan(Bits *b, uint8_t *img, int W, int H, int near)
{
    const int quant = 2 * near + 1, range = (255 + 2 * near) / quant + 1;
    int qbpp = 1; while ((1 << qbpp) < range) qbpp++;
    const int limit = 2 * (8 + 8) - qbpp - 1;                       /* LIMIT = 2*(bpp + max(8,bpp)) = 32 */
    const int t1 = 3 + 3 * near, t2 = 7 + 5 * near, t3 = 21 + 7 * near;
    int a_init = (range + 32) / 64; if (a_init < 2) a_init = 2;
    int A[365], B[365], C[365], N[365];
    for (int i = 0; i < 365; i++) { A[i] = a_init; B[i] = 0; C[i] = 0; N[i] = 1; }
    int rA[2] = {a_init, a_init}, rN[2] = {1, 1}, rNn[2] = {0, 0};
    int run_index = 0;
    for (int y = 0; y < H; y++) {
        uint8_t *cur = img + (size_t)y * W;
        const uint8_t *up = y > 0 ? cur - W : NULL;
        int x = 0;
        while (x < W) {
            const int Rb = up ? up[x] : 0;
            const int Ra = x > 0 ? cur[x - 1] : Rb;
            const int Rc = x > 0 ? (up ? up[x - 1] : 0) : (y > 1 ? cur[-2 * W] : 0);
            const int Rd = up ? (x + 1 < W ? up[x + 1] : Rb) : 0;
            const int q1 = gquant(Rd - Rb, near, t1, t2, t3), q2 = gquant(Rb - Rc, near, t1, t2, t3), q3 = gquant(Rc - Ra, near, t1, t2, t3);
            if (q1 == 0 && q2 == 0 && q3 == 0) {
                /* run mode (A.7.1) */
                int done = 0;
                while (!done) {
                    if (bit(b)) {
                        const int blk = 1 << J[run_index], rem = W - x;
                        const int cnt = blk <= rem ? blk : rem;
                        memset(cur + x, Ra, (size_t)cnt);
                        x += cnt;
                        if (blk <= rem && run_index < 31) run_index++;
                        if (x >= W) done = 1;
                    } else {
                        const int cnt = bits(b, J[run_index]);
                        if (x + cnt >= W) { b->err = 1; return -10; }
                        memset(cur + x, Ra, (size_t)cnt);
                        x += cnt;
                        /* run interruption sample (A.7.2) */
                        const int Rb2 = up ? up[x] : 0;
                        const int Ra2 = x > 0 ? cur[x - 1] : Rb2;          /* == Ra unless the run had length 0 at column 0 */
                        const int t = abs(Ra2 - Rb2) <= near;
                        const int sign = (!t && Ra2 > Rb2) ? -1 : 1;
                        const int px = t ? Ra2 : Rb2;
                        const int temp = t ? rA[1] + (rN[1] >> 1) : rA[0];
                        const int k = kparam(temp, rN[t]);
                        const int me = golomb(b, k, limit - J[run_index] - 1, qbpp);
                        int e;
                        const int tmp = me + t;
                        const int cond = (k == 0 && 2 * rNn[t] < rN[t]);
                        if (tmp & 1) { e = (tmp + 1) >> 1; if (!cond) e = -e; }       /* map = 1 */
                        else { e = tmp >> 1; if (cond) e = -e; }                       /* map = 0 */
                        int rx = px + sign * e * quant;
                        if (rx < -near) rx += range * quant; else if (rx > 255 + near) rx -= range * quant;
                        cur[x] = (uint8_t)clamp255(rx);
                        x++;
                        if (e < 0) rNn[t]++;
                        rA[t] += (me + 1 - t) >> 1;
                        if (rN[t] >= 64) { rA[t] >>= 1; rN[t] >>= 1; rNn[t] >>= 1; }
                        rN[t]++;
                        if (run_index > 0) run_index--;
                        done = 1;
                    }
                    if (b->err) return -11;
                }
            } else {
                /* regular mode (A.4 - A.6) */
                int q = 81 * q1 + 9 * q2 + q3, sign = 1;
                if (q < 0) { q = -q; sign = -1; }
                int px = Rc >= (Ra > Rb ? Ra : Rb) ? (Ra < Rb ? Ra : Rb) : Rc <= (Ra < Rb ? Ra : Rb) ? (Ra > Rb ? Ra : Rb) : Ra + Rb - Rc;
                px = clamp255(px + sign * C[q]);
                const int k = kparam(A[q], N[q]);
                const int me = golomb(b, k, limit, qbpp);
                int e = (me & 1) ? -((me + 1) >> 1) : me >> 1;
                if (near == 0 && k == 0 && 2 * B[q] <= -N[q]) e = -(e + 1);
                int rx = px + sign * e * quant;
                if (rx < -near) rx += range * quant; else if (rx > 255 + near) rx -= range * quant;
                cur[x] = (uint8_t)clamp255(rx);
                x++;
                B[q] += e * quant;
                A[q] += e < 0 ? -e : e;
                if (N[q] >= 64) { A[q] >>= 1; B[q] = B[q] >= 0 ? B[q] >> 1 : -((1 - B[q]) >> 1); N[q] >>= 1; }
                N[q]++;
                if (B[q] <= -N[q]) { B[q] += N[q]; if (C[q] > -128) C[q]--; if (B[q] <= -N[q]) B[q] = -N[q] + 1; }
                else if (B[q] > 0) { B[q] -= N[q]; if (C[q] < 127) C[q]++; if (B[q] > 0) B[q] = 0; }
                if (b->err) return -12;
            }
        }
    }
    return 0;
}

long jls_decode(const uint8_t *s, size_t n, uint8_t *out, int *w, int *h, int *ncomp, int *near_out)
{
    size_t i = 0;
    int W = 0, H = 0, NC = 0, comp = 0;
    if (n < 4 || s[0] != 0xFF || s[1] != 0xD8) return -1;
    i = 2;
    while (i + 4 <= n) {
        if (s[i] != 0xFF) return -2;
        const int m = s[i + 1];
        if (m == 0xD9) break;
        const int len = (s[i + 2] << 8) | s[i + 3];
        if (m == 0xF7) {                                            /* SOF55 */
            if (s[i + 4] != 8) return -3;
            H = (s[i + 5] << 8) | s[i + 6]; W = (s[i + 7] << 8) | s[i + 8]; NC = s[i + 9];
            i += 2 + len;
        } else if (m == 0xDA) {                                     /* SOS: one component, NEAR, ILV = 0 */
            if (s[i + 4] != 1 || !W || !H || comp >= NC) return -4;
            const int near = s[i + 7];
            if (s[i + 8] != 0) return -5;
            i += 2 + len;
            Bits b; b.p = s + i; b.end = s + n; b.acc = 0; b.nacc = 0; b.after_ff = 0; b.err = 0;
            const int rc = decode_scan(&b, out + (size_t)comp * W * H, W, H, near);
            if (rc) return rc;
            comp++;
            if (near_out) *near_out = near;
            i = (size_t)(b.p - s);                                  /* the scan ends on a byte boundary; after 0xFF a 0x00..0x7F byte was consumed */
        } else {
            i += 2 + len;
        }
    }
    if (comp != NC || NC == 0) return -6;
    *w = W; *h = H; *ncomp = NC;
    return 0;
}
