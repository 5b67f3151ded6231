/* pnm_io.c -- PNM reader / writer of the batch command line.
 *
 * Same contract as the reference's loadPNMImageFile / writePNMImageFile (src/imageio_pnm.c:10-29, :73-148):
 * plain and raw PBM/PGM/PPM (P1..P6) in, raw P5/P6 out, maxval 1..255 taken as is (no rescale), '#' comments.
 * Written as a header step + a pixel step so that the batch front-end can read the pixels of a raw P5/P6 file
 * with ONE fread straight into its (pinned) staging slot instead of a malloc'd buffer per image.
 *
 * Behaviour kept from the reference where a stream of odd input would show it:
 *   - the header scanner skips every byte that is not a digit or '#' (src/imageio_pnm.c:34-59);
 *   - after the last header number the rest of that LINE is skipped, not just one whitespace (:97-99);
 *   - plain PBM treats each whitespace-separated NUMBER as one sample, non-zero = black (:134-138);
 *   - plain samples above 255 wrap to their low byte (:137).
 */
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include "imcvt_pnm.h"

/* next decimal number of the stream (-1 at end of file); returns the byte that ended it */
static int next_number(FILE *fp, int *num)
{
    int ch;
    *num = -1;
    while ((ch = fgetc(fp)) != EOF) {
        if (ch == '#') {
            do ch = fgetc(fp); while (ch != EOF && ch != '\n' && ch != '\r');
            if (ch == EOF)
                break;
        } else if (ch >= '0' && ch <= '9') {
            unsigned v = 0;                    /* wraps like the reference's int does in practice, without the undefined behaviour */
            do { v = v * 10u + (unsigned)(ch - '0'); ch = fgetc(fp); } while (ch >= '0' && ch <= '9');
            *num = (int)v;
            return ch;
        }
    }
    return EOF;
}

int imcvtPNMReadHeader(FILE *fp, ImcvtPNMHeader *hd)
{
    int magic = fgetc(fp), type = fgetc(fp) - '0', w, h, maxval = 1, ch;
    ch = next_number(fp, &w);
    ch = next_number(fp, &h);
    if (type == 2 || type == 3 || type == 5 || type == 6)
        ch = next_number(fp, &maxval);
    if (magic != 'P' || type < 1 || type > 6 || w < 1 || h < 1 || maxval < 1 || maxval > 255)
        return 1;
    while (ch != '\n' && ch != EOF)
        ch = fgetc(fp);
    hd->type = type;
    hd->width = (uint32_t)w;
    hd->height = (uint32_t)h;
    hd->is_rgb = (type == 3 || type == 6);
    hd->bytes = (size_t)(hd->is_rgb ? 3 : 1) * (size_t)w * (size_t)h;
    return 0;
}

int imcvtPNMReadPixels(FILE *fp, const ImcvtPNMHeader *hd, uint8_t *dst)
{
    const size_t W = hd->width, H = hd->height;
    size_t i, j;
    int failed = 0;
    if (hd->type == 5 || hd->type == 6)
        return hd->bytes != fread(dst, 1, hd->bytes, fp);
    if (hd->type == 4) {                                /* raw PBM: rows padded to whole bytes, 1 = black */
        for (i = 0; i < H; i++)
            for (j = 0; j < W; j += 8) {
                int ch = fgetc(fp), b;
                failed |= (ch == EOF);
                for (b = 0; b < 8 && j + b < W; b++)
                    dst[i * W + j + b] = ((ch >> (7 - b)) & 1) ? 0 : 255;
            }
        return failed;
    }
    for (i = 0; i < hd->bytes; i++) {                   /* plain PBM / PGM / PPM */
        int v;
        next_number(fp, &v);
        failed |= (v < 0);
        dst[i] = (uint8_t)(hd->type != 1 ? v : (v ? 0 : 255));
    }
    return failed;
}

#ifndef IMCVT_PNM_NO_REFERENCE_NAMES
/* the reference's two names (src/imageio.h:8, :16), for callers that want one malloc'd buffer per image */
uint8_t *loadPNMImageFile(const char *p_filename, int *p_is_rgb, uint32_t *p_height, uint32_t *p_width)
{
    ImcvtPNMHeader hd;
    uint8_t *buf = NULL;
    FILE *fp = fopen(p_filename, "rb");
    if (fp == NULL)
        return NULL;
    if (!imcvtPNMReadHeader(fp, &hd) && (buf = (uint8_t *)malloc(hd.bytes + 8)) != NULL) {
        if (imcvtPNMReadPixels(fp, &hd, buf)) {
            free(buf);
            buf = NULL;
        } else {
            *p_is_rgb = hd.is_rgb;
            *p_height = hd.height;
            *p_width = hd.width;
        }
    }
    fclose(fp);
    return buf;
}

int writePNMImageFile(const char *p_filename, const uint8_t *p_buf, int is_rgb, uint32_t height, uint32_t width)
{
    const size_t len = (size_t)(is_rgb ? 3 : 1) * width * height;
    int failed;
    FILE *fp;
    if (width < 1 || height < 1 || (fp = fopen(p_filename, "wb")) == NULL)
        return 1;
    fprintf(fp, "P%c\n%u %u\n255\n", is_rgb ? '6' : '5', (unsigned)width, (unsigned)height);
    failed = (len != fwrite(p_buf, 1, len, fp));
    fclose(fp);
    return failed;
}
#endif
