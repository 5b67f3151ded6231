/* imcvt_imageio.h -- the reference's image-file API for the two accelerated writers, exported by
 * libimcvt_b200.so with the reference's exact signatures and return codes so that the reference's own
 * main.c and loaders link against this library unchanged.
 *
 *   writeJLSImageFile   replaces src/imageio_jls.c:428   (declared src/imageio.h:20)
 *   writeHEVCImageFile  replaces src/imageio_hevc.c:9    (declared src/imageio.h:21)
 *   HEVCImageEncoder    replaces src/HEVCe/HEVCe.c:1569  (declared src/HEVCe/HEVCe.h:5-12)
 *
 * return: 0 = success, 1 = failed (src/imageio.h:15).  Like the reference header this file includes
 * nothing; include <stdint.h> first.
 */
#ifndef IMCVT_IMAGEIO_H
#define IMCVT_IMAGEIO_H
#ifdef __cplusplus
extern "C" {
#endif

int writeJLSImageFile (const char *p_filename, const uint8_t *p_buf, int is_rgb, uint32_t height, uint32_t width, int near);
int writeHEVCImageFile(const char *p_filename, const uint8_t *p_buf, int is_rgb, uint32_t height, uint32_t width, int qpd6);

/* returns the HEVC stream length in bytes (<= 0 on failure); *ysz, *xsz are overwritten with the padded size;
 * img_rcon must hold the padded picture */
int HEVCImageEncoder(unsigned char *pbuffer, const unsigned char *img, unsigned char *img_rcon,
                     int *ysz, int *xsz, const int qpd6);

#ifdef __cplusplus
}
#endif
#endif
