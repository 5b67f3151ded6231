/* oracle/ref_wrap.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Wrapper translation unit that pulls the UNMODIFIED reference JPEG-LS encoder in by path
 * (REF_SRC is passed by oracle/Makefile and points at /root/reference/src) so that its
 * file-static in-memory encoders (imageio_jls.c:402 JLSencodeImageGray, :412 JLSencodeImageRGB)
 * can be reached through ctypes.  Nothing from the reference is copied into this repository:
 * the sources are compiled where they lie and the outputs land in oracle/_ref/ (git-ignored).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define STR2(x) #x
#define STR(x) STR2(x)
#include STR(REF_SRC/imageio_jls.c)

/* In-memory twin of writeJLSImageFile (imageio_jls.c:428-477): same widening / de-interleave
 * (:447-459), same encoder calls (:461-465), no file I/O.  Returns the stream length, 0 on failure. */
size_t ref_jls_encode(const uint8_t *p_buf, int is_rgb, uint32_t height, uint32_t width, int near_,
                      uint8_t *out, size_t out_cap)
{
    size_t n = (size_t)height * width, len;
    int planes = is_rgb ? 3 : 1;
    uint8_t *stream;
    int *r, *g, *b;
    if (width < 1 || width > 32767 || height < 1 || height > 32767) return 0;
    stream = (uint8_t *)malloc(8 * n + 65536 + (size_t)planes * n * sizeof(int));
    if (!stream) return 0;
    r = (int *)(stream + 8 * n + 65536); g = r + n; b = g + n;
    if (is_rgb) { for (size_t i = 0; i < n; i++) { r[i] = p_buf[3*i]; g[i] = p_buf[3*i+1]; b[i] = p_buf[3*i+2]; } }
    else        { for (size_t i = 0; i < n; i++) r[i] = p_buf[i]; }
    len = is_rgb ? (size_t)JLSencodeImageRGB(8, near_, (int)height, (int)width, r, g, b, stream)
                 : (size_t)JLSencodeImageGray(8, near_, (int)height, (int)width, r, stream);
    if (len > out_cap) len = 0; else memcpy(out, stream, len);
    free(stream);
    return len;
}
