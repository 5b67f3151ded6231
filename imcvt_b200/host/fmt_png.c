/* fmt_png.c -- PNG reader / writer of the batch command line.
 *
 * Same contract as the reference's writePNGImageFile / loadPNGImageFile (src/imageio_png.c:39-101, :110-188).  The writer
 * is the reference's "store" PNG: 8-bit gray or RGB, filter 0 on every line, one IDAT of stored (uncompressed) deflate
 * blocks -- byte-identical files.  The reader replaces the reference's third-party decoder (src/uPNG/, LodePNG-derived) with
 * a small one written from RFC 1950 / 1951 and the PNG specification: zlib framing, stored / fixed / dynamic Huffman blocks,
 * the five scan-line filters.  Huffman codes are kept as a tree built the way the reference's decoder builds its own, so that damaged
 * streams it still turns into a picture decode to the same picture (tree_build()); differential runs against the reference's CLI on
 * 12 000 mutated files left only differences where the reference shows uninitialised memory (lines the data does not reach).
 *
 * Behaviour kept from the reference, because stdout and the converted picture would show it:
 *   - the checks and their order (uPNG.c:898-964, :1005-1041): 29 bytes and the signature, IHDR first, colour type x bit depth,
 *     compression and filter method 0, no interlacing, no unknown critical chunk; chunk CRCs and the Adler-32 are not verified;
 *   - "this PNG format is not-yet supported, error code = 5 | 6 | 7" for an unknown critical chunk / interlacing / a colour format
 *     outside {gray, RGB, gray+alpha, RGBA} x the depths uPNG knows (palette pictures land here) (imageio_png.c:136-137);
 *   - a picture that decodes but is not LUMA8 / RGB8 / RGBA8 is refused with its format name (:145-149);
 *   - RGBA8 loses its alpha with a warning line (:163-164).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ writer */
static uint32_t crc32_update(uint32_t crc, const uint8_t *p, size_t n)
{
    static uint32_t table[256];
    if (table[1] == 0)
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++)
                c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
            table[i] = c;
        }
    for (size_t i = 0; i < n; i++)
        crc = table[(crc ^ p[i]) & 255u] ^ (crc >> 8);
    return crc;
}

static void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

/* length | type | data | crc(type + data) */
static int write_chunk(FILE *fp, const char *type, const uint8_t *data, uint32_t len)
{
    uint8_t head[8], tail[4];
    put_be32(head, len);
    memcpy(head + 4, type, 4);
    uint32_t crc = crc32_update(0xFFFFFFFFu, head + 4, 4);
    crc = crc32_update(crc, data, len);
    put_be32(tail, ~crc);
    return fwrite(head, 1, 8, fp) != 8 || (len && fwrite(data, 1, len, fp) != len) || fwrite(tail, 1, 4, fp) != 4;
}

/* return: 0 ok, 1 failed */
int writePNGImageFile(const char *p_filename, const uint8_t *p_buf, int is_rgb, uint32_t height, uint32_t width)
{
    if (width < 1 || height < 1)
        return 1;
    const size_t line = (size_t)(is_rgb ? 3 : 1) * width + 1, raw = line * height;      /* filter byte + samples, per line */
    const size_t nblocks = (raw + 0xFFFE) / 0xFFFF;
    uint8_t *z = (uint8_t *)malloc(2 + raw + 5 * nblocks + 4);
    if (z == NULL)
        return 1;
    FILE *fp = fopen(p_filename, "wb");
    if (fp == NULL) {
        free(z);
        return 1;
    }
    uint8_t ihdr[13];
    put_be32(ihdr, width);
    put_be32(ihdr + 4, height);
    ihdr[8] = 8; ihdr[9] = is_rgb ? 2 : 0; ihdr[10] = ihdr[11] = ihdr[12] = 0;

    /* zlib stream: header 78 01, stored blocks of at most 65535 bytes, Adler-32 */
    uint8_t *p = z;
    *p++ = 0x78; *p++ = 0x01;
    uint32_t a = 1, b = 0;
    size_t x = 0;                                   /* position inside the current line */
    for (size_t done = 0; done < raw; ) {
        const size_t n = raw - done < 0xFFFF ? raw - done : 0xFFFF;
        *p++ = done + n == raw;                     /* BFINAL on the last block, BTYPE 00 */
        *p++ = (uint8_t)n; *p++ = (uint8_t)(n >> 8);
        *p++ = (uint8_t)~n; *p++ = (uint8_t)(~n >> 8);
        for (size_t i = 0; i < n; i++) {
            const uint8_t v = x == 0 ? 0 : *p_buf++;
            if (++x == line)
                x = 0;
            *p++ = v;
            a = (a + v) % 65521u;
            b = (b + a) % 65521u;
        }
        done += n;
    }
    put_be32(p, (b << 16) | a);
    p += 4;
    int failed = fwrite("\x89PNG\r\n\x1a\n", 1, 8, fp) != 8;
    failed |= write_chunk(fp, "IHDR", ihdr, 13);
    failed |= write_chunk(fp, "IDAT", z, (uint32_t)(p - z));
    failed |= write_chunk(fp, "IEND", z, 0);
    free(z);
    fclose(fp);
    (void)failed;                                   /* the reference reports success once the file is open (:97-100) */
    return 0;
}

/* ------------------------------------------------------------------------------------------------ inflate (RFC 1951) */
typedef struct { const uint8_t *p; size_t n, bit; int bad; } Bits;

static unsigned take(Bits *s, int n)                /* n <= 16 bits, least significant first */
{
    unsigned v = 0;
    for (int i = 0; i < n; i++, s->bit++) {
        const size_t byte = s->bit >> 3;
        if (byte >= s->n) { s->bad = 1; return 0; }
        v |= (unsigned)((s->p[byte] >> (s->bit & 7)) & 1u) << i;
    }
    return v;
}

/* Huffman codes as a binary tree in an array, node[2 * position + bit]: a value below n is a symbol, anything else is n + the
 * position to continue at.  A table-driven canonical decoder would do for every stream zlib accepts; the tree is here because the
 * reference's decoder (uPNG.c:205-299) is one and its treatment of code sets that are NOT proper prefix codes shows in what it
 * decodes: a bit pattern that an incomplete set leaves unassigned decodes as symbol 0 where the walk leaves the tree, and an
 * over-subscribed set is refused only when an insertion runs out of positions -- short of that the later code lands where the walk
 * takes it.  tree_build() inserts the canonical codes in symbol order by the same procedure, so damaged streams that the reference
 * still turns into a picture give the same picture here. */
typedef struct { uint16_t node[2 * 288]; unsigned n; } Tree;
enum { TREE_EMPTY = 0x7FFF };

static int tree_build(Tree *t, const uint8_t *lengths, unsigned n, int maxlen)
{
    unsigned count[17] = { 0 }, start[17] = { 0 }, made = 0, pos = 0;
    t->n = n;
    for (unsigned i = 0; i < n; i++)
        count[lengths[i]]++;
    count[0] = 0;
    for (int len = 1; len <= maxlen; len++)
        start[len] = (start[len - 1] + count[len - 1]) << 1;         /* may outgrow len bits when over-subscribed: the low len bits count */
    for (unsigned i = 0; i < 2 * n; i++)
        t->node[i] = TREE_EMPTY;
    for (unsigned sym = 0; sym < n; sym++) {
        const int len = lengths[sym];
        if (len == 0)
            continue;
        const unsigned code = start[len]++;
        for (int k = len - 1; k >= 0; k--) {
            if (pos > n - 2)
                return -1;                                          /* out of positions: over-subscribed */
            uint16_t *slot = &t->node[2 * pos + ((code >> k) & 1u)];
            if (*slot != TREE_EMPTY) {
                pos = (unsigned)*slot - n;                          /* a symbol here (or garbage) sends pos out of range: caught above, next bit */
            } else if (k == 0) {
                *slot = (uint16_t)sym;
                pos = 0;
            } else {
                *slot = (uint16_t)(++made + n);
                pos = made;
            }
        }
    }
    for (unsigned i = 0; i < 2 * n; i++)
        if (t->node[i] == TREE_EMPTY)
            t->node[i] = 0;
    return 0;
}

static int tree_decode(Bits *s, const Tree *t)
{
    unsigned pos = 0;
    for (;;) {
        const unsigned bit = take(s, 1);
        if (s->bad)
            return -1;
        const unsigned v = t->node[2 * pos + bit];
        if (v < t->n)
            return (int)v;
        pos = v - t->n;
        if (pos >= t->n)
            return -1;
    }
}

static const uint16_t k_len_base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
static const uint8_t k_len_extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
static const uint16_t k_dist_base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
static const uint8_t k_dist_extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

static int inflate_codes(Bits *s, uint8_t *out, size_t cap, size_t *pos, const Tree *lit, const Tree *dist)
{
    for (;;) {
        const int sym = tree_decode(s, lit);
        if (sym < 0)
            return 1;
        if (sym < 256) {
            if (*pos >= cap)
                return 1;
            out[(*pos)++] = (uint8_t)sym;
        } else if (sym == 256) {
            return 0;
        } else if (sym <= 285) {
            const size_t len = k_len_base[sym - 257] + take(s, k_len_extra[sym - 257]);
            const int ds = tree_decode(s, dist);
            if (ds < 0 || ds > 29)
                return 1;
            const size_t d = k_dist_base[ds] + take(s, k_dist_extra[ds]);
            if (s->bad || d > *pos || *pos + len >= cap)          /* '>=': a match may not END at the last byte of the room (uPNG.c:542) */
                return 1;
            for (size_t i = 0; i < len; i++, (*pos)++)
                out[*pos] = out[*pos - d];
        }                                                           /* symbols 286 / 287 produce nothing (uPNG.c:495-552 has no branch for them) */
    }
}

/* zlib stream -> out[cap]; returns 0 when the stream is well formed and fits */
static int zlib_inflate(const uint8_t *z, size_t n, uint8_t *out, size_t cap)
{
    if (n < 2 || ((unsigned)z[0] * 256 + z[1]) % 31 != 0 || (z[0] & 15) != 8 || (z[0] >> 4) > 7 || (z[1] & 32))
        return 1;
    Bits s = { z + 2, n - 2, 0, 0 };
    size_t pos = 0;
    for (int last = 0; !last; ) {
        last = (int)take(&s, 1);
        const unsigned type = take(&s, 2);
        if (s.bad)
            return 1;
        if (type == 0) {
            s.bit = (s.bit + 7) & ~(size_t)7;
            const size_t at = s.bit >> 3;
            if (at + 4 > s.n)
                return 1;
            const unsigned len = s.p[at] | (s.p[at + 1] << 8), nlen = s.p[at + 2] | (s.p[at + 3] << 8);
            if ((len ^ nlen) != 0xFFFFu || at + 4 + len > s.n || pos + len >= cap)   /* '>=' as uPNG.c:588: a 1 x 1 picture, whose room is exactly its two bytes, is refused */
                return 1;
            memcpy(out + pos, s.p + at + 4, len);
            pos += len;
            s.bit = (at + 4 + len) << 3;
        } else if (type == 1 || type == 2) {
            Tree lit, dist;
            uint8_t lengths[320];
            if (type == 1) {
                int i = 0;
                for (; i < 144; i++) lengths[i] = 8;
                for (; i < 256; i++) lengths[i] = 9;
                for (; i < 280; i++) lengths[i] = 7;
                for (; i < 288; i++) lengths[i] = 8;
                tree_build(&lit, lengths, 288, 15);
                for (i = 0; i < 32; i++) lengths[i] = 5;
                tree_build(&dist, lengths, 32, 15);
            } else {
                static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
                const int nlit = (int)take(&s, 5) + 257, ndist = (int)take(&s, 5) + 1, ncode = (int)take(&s, 4) + 4;
                if (s.bad)
                    return 1;
                memset(lengths, 0, sizeof lengths);                 /* counts above 286 / 30 are taken as they come (uPNG.c:318-320) */
                for (int i = 0; i < ncode; i++)
                    lengths[order[i]] = (uint8_t)take(&s, 3);
                Tree clen;
                if (s.bad || tree_build(&clen, lengths, 19, 7))
                    return 1;
                uint8_t ll[288 + 32];
                memset(ll, 0, sizeof ll);
                int i = 0;
                while (i < nlit + ndist) {
                    const int sym = tree_decode(&s, &clen);
                    if (sym < 0)
                        return 1;
                    if (sym < 16) {
                        ll[i++] = (uint8_t)sym;
                    } else {
                        uint8_t v = 0;
                        int rep;
                        if (sym == 16) {
                            if (i == 0)
                                return 1;
                            v = ll[i - 1];
                            rep = 3 + (int)take(&s, 2);
                        } else if (sym == 17) {
                            rep = 3 + (int)take(&s, 3);
                        } else {
                            rep = 11 + (int)take(&s, 7);
                        }
                        if (s.bad || i + rep > nlit + ndist)
                            return 1;
                        while (rep--)
                            ll[i++] = v;
                    }
                }
                if (ll[256] == 0)
                    return 1;
                /* both alphabets at their full size, lengths past the transmitted counts zero (uPNG.c:313-316, :473-474) */
                memset(lengths, 0, sizeof lengths);
                memcpy(lengths, ll, (size_t)nlit);
                if (tree_build(&lit, lengths, 288, 15))
                    return 1;
                memset(lengths, 0, sizeof lengths);
                memcpy(lengths, ll + nlit, (size_t)ndist);
                if (tree_build(&dist, lengths, 32, 15))
                    return 1;
            }
            if (inflate_codes(&s, out, cap, &pos, &lit, &dist))
                return 1;
        } else {
            return 1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ reader */
static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

/* undoes the scan-line filters in place: lines of 1 + stride bytes -> out (stride bytes each); bpp = bytes per complete pixel */
static int unfilter(uint8_t *out, const uint8_t *in, size_t stride, size_t height, size_t bpp)
{
    for (size_t y = 0; y < height; y++) {
        const uint8_t *src = in + y * (stride + 1) + 1, *up = y ? out + (y - 1) * stride : NULL;
        uint8_t *dst = out + y * stride;
        const int type = src[-1];
        if (type > 4)
            return 1;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? dst[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            const int pred = type == 0 ? 0 : type == 1 ? a : type == 2 ? b : type == 3 ? (a + b) / 2 : paeth(a, b, c);
            dst[i] = (uint8_t)(src[i] + pred);
        }
    }
    return 0;
}

/* the reference decoder's format numbers (uPNG.h:43-57), only as far as its messages need them */
enum { F_BAD = 0, F_RGB8, F_RGB16, F_RGBA8, F_RGBA16, F_L1, F_L2, F_L4, F_L8, F_LA1, F_LA2, F_LA4, F_LA8 };

static int format_of(int color, int depth)
{
    switch (color) {
    case 0: return depth == 1 ? F_L1 : depth == 2 ? F_L2 : depth == 4 ? F_L4 : depth == 8 ? F_L8 : F_BAD;
    case 2: return depth == 8 ? F_RGB8 : depth == 16 ? F_RGB16 : F_BAD;
    case 4: return depth == 1 ? F_LA1 : depth == 2 ? F_LA2 : depth == 4 ? F_LA4 : depth == 8 ? F_LA8 : F_BAD;
    case 6: return depth == 8 ? F_RGBA8 : depth == 16 ? F_RGBA16 : F_BAD;
    default: return F_BAD;
    }
}

/* return: NULL failed, else malloc'd pixels */
uint8_t *loadPNGImageFile(const char *p_filename, int *p_is_rgb, uint32_t *p_height, uint32_t *p_width)
{
    static const char *names[] = { "BADFORMAT", "RGB8", "RGB16", "RGBA8", "RGBA16", "LUMA1", "LUMA2", "LUMA4", "LUMA8",
                                   "LUMA_ALPHA1", "LUMA_ALPHA2", "LUMA_ALPHA4", "LUMA_ALPHA8" };
    FILE *fp = fopen(p_filename, "rb");
    if (fp == NULL)
        return NULL;
    uint8_t *file = NULL, *z = NULL, *raw = NULL, *pix = NULL, *img = NULL;
    size_t n = 0;
    if (fseek(fp, 0, SEEK_END) == 0) {
        const long sz = ftell(fp);
        if (sz > 0 && fseek(fp, 0, SEEK_SET) == 0 && (file = (uint8_t *)malloc((size_t)sz)) != NULL)
            n = fread(file, 1, (size_t)sz, fp);
    }
    fclose(fp);
    int unsupported = 0;                            /* the reference decoder's error code when it is one the caller reports */
    if (file == NULL || n < 29 || memcmp(file, "\x89PNG\r\n\x1a\n", 8) != 0 || memcmp(file + 12, "IHDR", 4) != 0)
        goto fail;
    const uint32_t width = be32(file + 16), height = be32(file + 20);
    const int depth = file[24], color = file[25], format = format_of(color, depth);
    if (format == F_BAD) { unsupported = 7; goto fail; }
    if (file[26] != 0 || file[27] != 0)
        goto fail;
    if (file[28] != 0) { unsupported = 6; goto fail; }

    /* the IDAT chunks, concatenated */
    size_t zn = 0;
    for (int pass = 0; pass < 2; pass++) {
        size_t at = 33, fill = 0;
        while (at < n) {
            if (at + 12 > n)
                goto fail;
            const uint32_t len = be32(file + at);
            if (len > 0x7FFFFFFFu || at + len + 12 > n)
                goto fail;
            if (memcmp(file + at + 4, "IDAT", 4) == 0) {
                if (pass)
                    memcpy(z + fill, file + at + 8, len);
                fill += len;
            } else if (memcmp(file + at + 4, "IEND", 4) == 0) {
                break;
            } else if ((file[at + 4] & 32) == 0) {
                unsupported = 5;
                goto fail;
            }
            at += (size_t)len + 12;
        }
        if (pass == 0) {
            zn = fill;
            if ((z = (uint8_t *)malloc(zn + 1)) == NULL)
                goto fail;
        }
    }
    {
        static const int channels[7] = { 1, 0, 3, 0, 2, 0, 4 };
        const size_t bits_pp = (size_t)channels[color] * (size_t)depth;
        const size_t stride = ((size_t)width * bits_pp + 7) / 8, raw_size = (stride + 1) * height;
        /* room for the inflated lines as the reference decoder sizes it (uPNG.c:1092): a stream that inflates to more is malformed */
        const size_t cap = (size_t)(uint32_t)(width * (height * (uint32_t)bits_pp + 7u)) / 8 + height;
        const size_t room = cap > raw_size ? cap : raw_size;
        /* deflate expands at most 1032 : 1: a picture of more than a gigabyte that the data cannot cover even then is a damaged
         * header, not a picture (the reference would un-filter all of it; smaller ones are decoded as the reference decodes them,
         * what the data does not reach reads as zeros) */
        if (raw_size > ((size_t)1 << 30) && raw_size / 1032 > zn + 64)
            goto fail;
        if (width == 0 || height == 0 || (raw = (uint8_t *)calloc(room + 1, 1)) == NULL || (pix = (uint8_t *)malloc(stride * height + 1)) == NULL)
            goto fail;
        if (zlib_inflate(z, zn, raw, cap) || unfilter(pix, raw, stride, height, bits_pp >= 8 ? bits_pp / 8 : 1))
            goto fail;
        if (format != F_RGBA8 && format != F_RGB8 && format != F_L8) {
            printf("   ***ERROR: only support LUMA8, RGB8, and RGBA8. But this PNG is %s\n", names[format]);
            goto fail;
        }
        *p_is_rgb = format != F_L8;
        *p_height = height;
        *p_width = width;
        const size_t npix = (size_t)width * height;
        if ((img = (uint8_t *)malloc(npix * (format == F_L8 ? 1 : 3))) != NULL) {
            if (format == F_RGBA8) {
                printf("   *warning: disard alpha channel of this PNG\n");
                for (size_t i = 0; i < npix; i++)
                    memcpy(img + 3 * i, pix + 4 * i, 3);
            } else {
                memcpy(img, pix, npix * (format == F_L8 ? 1 : 3));
            }
        }
    }
    free(file); free(z); free(raw); free(pix);
    return img;
fail:
    if (unsupported)
        printf("   ***ERROR: this PNG format is not-yet supported, error code = %d\n", unsupported);
    free(file); free(z); free(raw); free(pix);
    return NULL;
}
