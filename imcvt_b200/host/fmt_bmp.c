/* fmt_bmp.c -- BMP reader / writer of the batch command line.
 *
 * Same contract as the reference's loadBMPImageFile / writeBMPImageFile (src/imageio_bmp.c:21-82, :85-177); written from
 * the BMP layout (14-byte file header, 40-byte BITMAPINFOHEADER, optional palette, bottom-up rows padded to 4 bytes) on
 * whole-file buffers instead of a stdio call per byte.
 *
 * Behaviour kept from the reference, because a file on disk or a later conversion would show it:
 *   writer  gray -> 8 bit with a 256-entry grey palette (B, G, R, 0xFF), RGB -> 24 bit BGR; image-size field 0, resolution
 *           0x0EC4 both ways, palette count 256 / 0 (:39-58); width or height 0 fails before the file is created (:31-35);
 *   reader  takes 8 / 24 / 32 bit uncompressed with a pixel offset >= 54, a header >= 40 bytes and at most 256 palette
 *           entries (:112); an 8-bit picture whose palette is not grey comes back as RGB (:123-130); a palette count of 0 is
 *           NOT read as 256, so such a picture decodes through an all-zero palette (:122); bytes past the end of the file
 *           read as 0xFF, as fgetc's EOF does when it is stored in a byte (:152-171).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void put_le(uint8_t *p, uint32_t v, int n)
{
    for (int i = 0; i < n; i++)
        p[i] = (uint8_t)(v >> (8 * i));
}

/* return: 0 ok, 1 failed */
int writeBMPImageFile(const char *p_filename, const uint8_t *p_buf, int is_rgb, uint32_t height, uint32_t width)
{
    if (width < 1 || height < 1)
        return 1;
    const size_t bpp = is_rgb ? 3 : 1, row = bpp * width, row_padded = (row + 3) / 4 * 4;
    const size_t ncolors = is_rgb ? 0 : 256, head = 14 + 40 + 4 * ncolors, total = head + (size_t)height * row_padded;
    FILE *fp = fopen(p_filename, "wb");
    if (fp == NULL)
        return 1;
    uint8_t *file = (uint8_t *)calloc(total, 1);
    if (file == NULL) {
        fclose(fp);
        return 1;
    }
    /* file header */
    file[0] = 'B'; file[1] = 'M';
    put_le(file + 2, (uint32_t)total, 4);
    put_le(file + 10, (uint32_t)head, 4);
    /* BITMAPINFOHEADER */
    put_le(file + 14, 40, 4);
    put_le(file + 18, width, 4);
    put_le(file + 22, height, 4);
    put_le(file + 26, 1, 2);
    put_le(file + 28, is_rgb ? 24 : 8, 2);
    put_le(file + 38, 0x0EC4, 4);
    put_le(file + 42, 0x0EC4, 4);
    put_le(file + 46, (uint32_t)ncolors, 4);
    for (size_t i = 0; i < ncolors; i++) {
        uint8_t *e = file + 54 + 4 * i;
        e[0] = e[1] = e[2] = (uint8_t)i;
        e[3] = 0xFF;
    }
    /* rows, bottom first */
    for (uint32_t y = 0; y < height; y++) {
        const uint8_t *src = p_buf + (size_t)(height - 1 - y) * row;
        uint8_t *dst = file + head + (size_t)y * row_padded;
        if (is_rgb) {
            for (uint32_t x = 0; x < width; x++, src += 3, dst += 3) {
                dst[0] = src[2]; dst[1] = src[1]; dst[2] = src[0];
            }
        } else {
            memcpy(dst, src, row);
        }
    }
    const int failed = fwrite(file, 1, total, fp) != total;
    free(file);
    return (fclose(fp) != 0) | failed;
}

/* the file as one buffer; a byte past its end reads as EOF does in the reference */
typedef struct { const uint8_t *p; size_t n, pos; } Bytes;

static uint32_t get_le(Bytes *b, uint32_t len)
{
    uint32_t v = 0;
    for (uint32_t i = 0; i < len; i++, b->pos++) {
        /* fgetc() << 8 i with fgetc() == -1 at the end of the file: every bit from 8 i upwards (only the low four bytes matter) */
        const uint32_t byte = b->pos < b->n ? b->p[b->pos] : 0xFFFFFFFFu;
        if (i < 4)
            v |= byte << (8 * i);
    }
    return v;
}
static uint8_t get_byte(Bytes *b)
{
    const uint8_t v = b->pos < b->n ? b->p[b->pos] : 0xFF;
    b->pos++;
    return v;
}

/* return: NULL failed, else malloc'd pixels */
uint8_t *loadBMPImageFile(const char *p_filename, int *p_is_rgb, uint32_t *p_height, uint32_t *p_width)
{
    FILE *fp = fopen(p_filename, "rb");
    if (fp == NULL)
        return NULL;
    uint8_t *file = NULL;
    size_t n = 0;
    if (fseek(fp, 0, SEEK_END) == 0) {
        const long sz = ftell(fp);
        if (sz > 0 && fseek(fp, 0, SEEK_SET) == 0 && (file = (uint8_t *)malloc((size_t)sz)) != NULL)
            n = fread(file, 1, (size_t)sz, fp);
    }
    fclose(fp);
    Bytes b = { file, n, 0 };
    const uint32_t magic = get_le(&b, 2);
    get_le(&b, 8);
    const uint32_t offset = get_le(&b, 4), dib = get_le(&b, 4);
    const uint32_t width = get_le(&b, 4), height = get_le(&b, 4);
    get_le(&b, 2);
    const uint32_t bits = get_le(&b, 2), compression = get_le(&b, 4);
    get_le(&b, 12);
    const uint32_t ncolors = get_le(&b, 4);
    get_le(&b, 4);
    *p_width = width;
    *p_height = height;
    if (magic != 0x4D42 || offset < 54 || dib < 40 || width < 1 || height < 1 || (bits != 8 && bits != 24 && bits != 32) ||
        compression != 0 || ncolors > 256) {
        free(file);
        return NULL;
    }
    const uint32_t bytes_pp = bits / 8;
    uint8_t pal[256][3];                       /* R, G, B */
    memset(pal, 0, sizeof pal);
    int is_rgb = bytes_pp > 1;
    if (bytes_pp == 1) {
        b.pos += dib - 40;
        for (uint32_t i = 0; i < ncolors; i++) {
            pal[i][2] = get_byte(&b); pal[i][1] = get_byte(&b); pal[i][0] = get_byte(&b);
            get_byte(&b);
            if (pal[i][0] != pal[i][1] || pal[i][1] != pal[i][2])
                is_rgb = 1;
        }
    }
    *p_is_rgb = is_rgb;
    const size_t out_pp = is_rgb ? 3 : 1;
    /* a BMP's sizes are 32-bit fields: pixel data of 4 GiB or more cannot be a BMP file (a damaged header would otherwise make
     * this loader write tens of gigabytes of 0xFF; the reference does exactly that, fgetc by fgetc) */
    if (((uint64_t)bytes_pp * width + 3) / 4 * 4 * (uint64_t)height > 0xFFFFFFFFull) {
        free(file);
        return NULL;
    }
    uint8_t *img = (uint8_t *)malloc(out_pp * width * height);
    if (img == NULL) {
        free(file);
        return NULL;
    }
    const size_t row = (size_t)bytes_pp * width, row_padded = (row + 3) / 4 * 4;
    for (uint32_t y = 0; y < height; y++) {
        uint8_t *dst = img + out_pp * (size_t)(height - 1 - y) * width;
        b.pos = (size_t)offset + (size_t)y * row_padded;
        if (b.pos >= b.n) {                     /* the whole row lies past the end of the file: every byte reads 0xFF */
            if (bytes_pp > 1 || !is_rgb) {
                memset(dst, bytes_pp > 1 ? 0xFF : pal[0xFF][0], out_pp * width);
            } else {
                for (uint32_t x = 0; x < width; x++, dst += 3)
                    memcpy(dst, pal[0xFF], 3);
            }
            continue;
        }
        for (uint32_t x = 0; x < width; x++) {
            if (bytes_pp > 1) {
                dst[2] = get_byte(&b); dst[1] = get_byte(&b); dst[0] = get_byte(&b);
                dst += 3;
                if (bytes_pp == 4)
                    get_byte(&b);
            } else {
                const uint8_t *e = pal[get_byte(&b)];
                *dst++ = e[0];
                if (is_rgb) { *dst++ = e[1]; *dst++ = e[2]; }
            }
        }
    }
    free(file);
    return img;
}
