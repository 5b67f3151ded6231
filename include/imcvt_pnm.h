/* imcvt_pnm.h -- two-step PNM reader of the batch command line (imcvt_b200/host/pnm_io.c).
 * Replaces loadPNMImageFile's single malloc'd buffer (reference: src/imageio_pnm.c:73-148) with
 * header-then-pixels, so that the caller chooses where the pixels land (a pinned staging slot). */
#ifndef IMCVT_PNM_H
#define IMCVT_PNM_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int      type;          /* the digit of the magic number: 1..6 */
    int      is_rgb;        /* 1 for P3 / P6 */
    uint32_t width, height;
    size_t   bytes;         /* width * height * (is_rgb ? 3 : 1): what imcvtPNMReadPixels() writes */
} ImcvtPNMHeader;

/* 0 = ok and the stream stands at the first sample; 1 = not a PNM file this reader takes */
int imcvtPNMReadHeader(FILE *fp, ImcvtPNMHeader *hd);
/* 0 = ok; 1 = short or malformed sample data.  dst holds hd->bytes bytes, interleaved u8 */
int imcvtPNMReadPixels(FILE *fp, const ImcvtPNMHeader *hd, uint8_t *dst);

#ifdef __cplusplus
}
#endif
#endif /* IMCVT_PNM_H */
