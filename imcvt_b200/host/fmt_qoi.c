/* fmt_qoi.c -- QOI reader / writer of the batch command line.
 *
 * Same contract as the reference's loadQOIImageFile / writeQOIImageFile (src/imageio_qoi.c:6-119, :122-255); written from the
 * QOI format (14-byte header, ops INDEX / DIFF / LUMA / RUN / RGB / RGBA over a 64-entry colour table hashed by
 * 3 r + 5 g + 7 b + 11 a).
 *
 * Behaviour kept from the reference, because the files would show it:
 *   writer  always three channels, colour space 0; a gray picture is written as r = g = b (:33-38); the 8-byte end marker
 *           of the QOI specification is NOT written (:98-116); the colour table is refreshed on every pixel, runs included;
 *   reader  three or four channels, alpha dropped, the result is always RGB (:149-186); decoding stops when the picture is
 *           full or the data ends, whichever comes first (:238-251) -- pixels the data does not reach are left black here
 *           (the reference leaves them uninitialised).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint8_t r, g, b, a; } Px;

static unsigned hash_of(Px p) { return (3u * p.r + 5u * p.g + 7u * p.b + 11u * p.a) & 63u; }
static int same(Px x, Px y) { return x.r == y.r && x.g == y.g && x.b == y.b && x.a == y.a; }

/* return: 0 ok, 1 failed */
int writeQOIImageFile(const char *p_filename, const uint8_t *p_buf, int is_rgb, uint32_t height, uint32_t width)
{
    if (width < 1 || height < 1)
        return 1;
    const size_t npix = (size_t)width * height;
    uint8_t *out = (uint8_t *)malloc(14 + 5 * npix + 64);
    if (out == NULL)
        return 1;
    FILE *fp = fopen(p_filename, "wb");
    if (fp == NULL) {
        free(out);
        return 1;
    }
    uint8_t *q = out;
    memcpy(q, "qoif", 4);
    for (int i = 0; i < 4; i++) {
        q[4 + i] = (uint8_t)(width >> (24 - 8 * i));
        q[8 + i] = (uint8_t)(height >> (24 - 8 * i));
    }
    q[12] = 3;
    q[13] = 0;
    q += 14;

    Px table[64];
    memset(table, 0, sizeof table);
    Px prev = { 0, 0, 0, 255 };
    unsigned run = 0;
    for (size_t i = 0; i < npix; i++) {
        Px cur;
        cur.r = cur.g = cur.b = *p_buf++;
        if (is_rgb) { cur.g = *p_buf++; cur.b = *p_buf++; }
        cur.a = 255;
        const unsigned h = hash_of(cur);
        if (same(cur, prev)) {
            if (++run == 62) { *q++ = (uint8_t)(0xC0 | (run - 1)); run = 0; }
        } else {
            if (run) { *q++ = (uint8_t)(0xC0 | (run - 1)); run = 0; }
            if (same(cur, table[h])) {
                *q++ = (uint8_t)h;
            } else {
                /* differences wrap at 8 bits, as the format defines them */
                const uint8_t dr = (uint8_t)(cur.r - prev.r + 2), dg = (uint8_t)(cur.g - prev.g + 2), db = (uint8_t)(cur.b - prev.b + 2);
                const uint8_t lr = (uint8_t)(dr - dg + 8), lb = (uint8_t)(db - dg + 8), lg = (uint8_t)(dg + 30);
                if (dr < 4 && dg < 4 && db < 4) {
                    *q++ = (uint8_t)(0x40 | (dr << 4) | (dg << 2) | db);
                } else if (lr < 16 && lg < 64 && lb < 16) {
                    *q++ = (uint8_t)(0x80 | lg);
                    *q++ = (uint8_t)((lr << 4) | lb);
                } else {
                    *q++ = 0xFE;
                    *q++ = cur.r; *q++ = cur.g; *q++ = cur.b;
                }
            }
        }
        table[h] = prev = cur;
    }
    if (run)
        *q++ = (uint8_t)(0xC0 | (run - 1));
    const size_t total = (size_t)(q - out);
    const int failed = fwrite(out, 1, total, fp) != total;
    free(out);
    fclose(fp);
    return failed;
}

/* return: NULL failed, else malloc'd pixels (always RGB) */
uint8_t *loadQOIImageFile(const char *p_filename, int *p_is_rgb, uint32_t *p_height, uint32_t *p_width)
{
    FILE *fp = fopen(p_filename, "rb");
    if (fp == NULL)
        return NULL;
    uint8_t *file = NULL;
    size_t n = 0;
    if (fseek(fp, 0, SEEK_END) == 0) {
        const long sz = ftell(fp);
        if (sz > 0 && fseek(fp, 0, SEEK_SET) == 0 && (file = (uint8_t *)malloc((size_t)sz + 8)) != NULL) {
            n = fread(file, 1, (size_t)sz, fp);
            memset(file + n, 0, 8);            /* an op cut short by the end of the file reads zeros */
        }
    }
    fclose(fp);
    if (file == NULL || n < 4 || memcmp(file, "qoif", 4) != 0) {
        free(file);
        return NULL;
    }
    /* header bytes past the end of the file count as EOF (-1) does in the reference's arithmetic */
    uint32_t dims[2];
    for (int k = 0; k < 2; k++) {
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) {
            const size_t at = 4 + 4 * (size_t)k + (size_t)i;
            v = (i ? v << 8 : 0) + (at < n ? file[at] : 0xFFFFFFFFu);
        }
        dims[k] = v;
    }
    const uint32_t width = dims[0], height = dims[1];
    const uint8_t channels = 12 < n ? file[12] : 0xFF;
    *p_width = width;
    *p_height = height;
    if (width < 1 || height < 1 || channels < 3 || channels > 4 || n <= 14) {
        free(file);
        return NULL;
    }
    const size_t npix = (size_t)width * height;
    uint8_t *img = (uint8_t *)calloc(npix, 3);
    if (img == NULL) {
        free(file);
        return NULL;
    }
    *p_is_rgb = 1;

    Px table[64];
    memset(table, 0, sizeof table);
    Px px = { 0, 0, 0, 255 };
    const uint8_t *q = file + 14, *end = file + n;
    size_t done = 0;
    while (done < npix) {
        const uint8_t op = *q++;
        unsigned run = 1;
        switch (op >> 6) {
        case 0:
            px = table[op & 63];
            break;
        case 1:
            px.r = (uint8_t)(px.r + ((op >> 4) & 3) - 2);
            px.g = (uint8_t)(px.g + ((op >> 2) & 3) - 2);
            px.b = (uint8_t)(px.b + (op & 3) - 2);
            break;
        case 2: {
            const uint8_t second = *q++;
            const int dg = (op & 63) - 32;
            px.g = (uint8_t)(px.g + dg);
            px.r = (uint8_t)(px.r + dg - 8 + (second >> 4));
            px.b = (uint8_t)(px.b + dg - 8 + (second & 15));
            break;
        }
        default:
            if ((op & 63) < 62) {
                run = 1 + (op & 63);
            } else {
                px.r = *q++; px.g = *q++; px.b = *q++;
                if ((op & 63) == 63)
                    px.a = *q++;
            }
            break;
        }
        table[hash_of(px)] = px;
        for (; run > 0 && done < npix; run--, done++) {
            img[3 * done] = px.r; img[3 * done + 1] = px.g; img[3 * done + 2] = px.b;
        }
        if (q >= end)
            break;
    }
    free(file);
    return img;
}
