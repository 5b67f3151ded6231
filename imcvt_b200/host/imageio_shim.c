/* imageio_shim.c -- host C side of the drop-in boundary: the reference's file-writer signatures
 * (src/imageio.h:20-21, src/HEVCe/HEVCe.h:5-12) implemented on top of the batch C-ABI (imcvt_b200.h).
 * Behaviour mirrored from the reference wrappers: argument checks and return codes of
 * writeJLSImageFile (src/imageio_jls.c:428-477) and writeHEVCImageFile (src/imageio_hevc.c:9-53),
 * including the RGB -> green-channel warning text (imageio_hevc.c:22).
 */
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include "imcvt_b200.h"
#include "imcvt_imageio.h"

static int dump(const char *fname, const uint8_t *buf, size_t len)
{
    int failed = 1;
    FILE *fp = fopen(fname, "wb");
    if (fp) {
        failed = (len != fwrite(buf, 1, len, fp));
        fclose(fp);
    }
    return failed;
}

int writeJLSImageFile(const char *p_filename, const uint8_t *p_buf, int is_rgb, uint32_t height, uint32_t width, int near)
{
    size_t cap, len = 0;
    uint8_t *stream;
    int failed;
    if (width < 1 || width > 32767 || height < 1 || height > 32767)
        return 1;
    cap = imcvtJLSBound(is_rgb, height, width);
    stream = (uint8_t *)malloc(cap);
    if (stream == NULL)
        return 1;
    failed = imcvtEncodeBatchJLS(p_buf, (size_t)height * width * (is_rgb ? 3 : 1), 1, is_rgb, height, width, near,
                                 stream, cap, &len);
    if (!failed)
        failed = dump(p_filename, stream, len);
    free(stream);
    return failed;
}

int HEVCImageEncoder(unsigned char *pbuffer, const unsigned char *img, unsigned char *img_rcon,
                     int *ysz, int *xsz, const int qpd6)
{
    size_t len = 0;
    int h = *ysz, w = *xsz, hp, wp;
    if (h < 1 || w < 1)
        return 0;
    /* the reference silently clips to 8192 and pads to a multiple of 32 (HEVCe.c:1580-1581) */
    hp = ((h > 8192 ? 8192 : h) + 31) / 32 * 32;
    wp = ((w > 8192 ? 8192 : w) + 31) / 32 * 32;
    if (imcvtEncodeBatchHEVC(img, (size_t)h * w, 1, (uint32_t)h, (uint32_t)w, qpd6,
                             pbuffer, imcvtHEVCBound((uint32_t)h, (uint32_t)w), &len,
                             img_rcon, (size_t)hp * wp))
        return 0;
    *ysz = hp;
    *xsz = wp;
    return (int)len;
}

int writeHEVCImageFile(const char *p_filename, const uint8_t *p_buf, int is_rgb, uint32_t height, uint32_t width, int qpd6)
{
    size_t npix = (size_t)height * width, len = 0, cap;
    uint8_t *stream;
    int failed;
    if (height < 1 || width < 1)
        return 1;
    cap = imcvtHEVCBound(height, width);
    stream = (uint8_t *)malloc(cap);
    if (stream == NULL)
        return 1;
    if (is_rgb) {       /* the green channel is picked on the device (imcvtEncodeBatchHEVCGreen), the warning stays */
        printf("   warning: this HEVCencoder currently only support gray 8-bit image instead of RGB image. Only compress the green channel of this image.\n");
        failed = imcvtEncodeBatchHEVCGreen(p_buf, npix * 3, 1, height, width, qpd6, stream, cap, &len);
    } else {
        failed = imcvtEncodeBatchHEVC(p_buf, npix, 1, height, width, qpd6, stream, cap, &len, NULL, 0);
    }
    if (!failed && len == 0)
        failed = 1;
    if (!failed)
        failed = dump(p_filename, stream, len);
    free(stream);
    return failed;
}
