/* fake_engine.c -- TEST DOUBLE of libimcvt_b200.so's C-ABI for the host-side tests of the batch command line.  NOT a codec and
 * never part of the product: a "stream" is a 16-byte tag (codec, setting, is_rgb / green flag, height, width) followed by the
 * frame's own bytes, so that tests/test_cli_staging.py can see on a CPU box WHICH pixels the command line staged for WHICH file, in
 * which batches (every call is logged to $FAKE_ENGINE_LOG).  Built by the test into a temporary directory as libimcvt_b200.so and
 * picked up through LD_LIBRARY_PATH; the real library has no CPU path at all. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void logcall(const char *what, int n, int a, uint32_t h, uint32_t w, int setting, int ndev)
{
    const char *path = getenv("FAKE_ENGINE_LOG");
    FILE *fp;
    if (path == NULL || (fp = fopen(path, "a")) == NULL)
        return;
    fprintf(fp, "%s n=%d rgb=%d h=%u w=%u setting=%d ndev=%d\n", what, n, a, h, w, setting, ndev);
    fclose(fp);
}
static void emit(uint8_t *slot, size_t *len, char codec, int setting, int flag, uint32_t h, uint32_t w, const uint8_t *px, size_t bytes)
{
    memset(slot, 0, 16);
    slot[0] = (uint8_t)codec; slot[1] = (uint8_t)setting; slot[2] = (uint8_t)flag;
    memcpy(slot + 4, &h, 4); memcpy(slot + 8, &w, 4);
    memcpy(slot + 16, px, bytes);
    *len = 16 + bytes;
}
int imcvtInit(int device) { (void)device; return 0; }
void imcvtShutdown(void) {}
const char *imcvtLastError(void) { return "fake engine"; }
size_t imcvtJLSBound(int is_rgb, uint32_t h, uint32_t w) { return 64 + (size_t)h * w * (is_rgb ? 3 : 1); }
size_t imcvtHEVCBound(uint32_t h, uint32_t w) { return 64 + (size_t)h * w; }
void *imcvtHostAlloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void imcvtHostFree(void *p) { free(p); }
int imcvtDeviceCount(void) { const char *e = getenv("FAKE_ENGINE_DEVICES"); return e ? atoi(e) : 1; }

int imcvtEncodeBatchJLSMulti(const uint8_t *f, size_t fs, int n, int is_rgb, uint32_t h, uint32_t w, int near_, uint8_t *out, size_t os, size_t *len,
                             const int *devices, int ndev)
{
    (void)devices;
    logcall("jls", n, is_rgb, h, w, near_, ndev);
    for (int i = 0; i < n; i++) emit(out + (size_t)i * os, &len[i], 'J', near_, is_rgb, h, w, f + (size_t)i * fs, (size_t)h * w * (is_rgb ? 3 : 1));
    return 0;
}
int imcvtEncodeBatchJLS(const uint8_t *f, size_t fs, int n, int is_rgb, uint32_t h, uint32_t w, int near_, uint8_t *out, size_t os, size_t *len)
{
    return imcvtEncodeBatchJLSMulti(f, fs, n, is_rgb, h, w, near_, out, os, len, NULL, 0);
}
int imcvtEncodeBatchHEVCMulti(const uint8_t *f, size_t fs, int n, uint32_t h, uint32_t w, int q, uint8_t *out, size_t os, size_t *len,
                              uint8_t *rcon, size_t rs, const int *devices, int ndev)
{
    (void)rcon; (void)rs; (void)devices;
    logcall("hevc", n, 0, h, w, q, ndev);
    for (int i = 0; i < n; i++) emit(out + (size_t)i * os, &len[i], 'H', q, 0, h, w, f + (size_t)i * fs, (size_t)h * w);
    return 0;
}
int imcvtEncodeBatchHEVC(const uint8_t *f, size_t fs, int n, uint32_t h, uint32_t w, int q, uint8_t *out, size_t os, size_t *len, uint8_t *rcon, size_t rs)
{
    return imcvtEncodeBatchHEVCMulti(f, fs, n, h, w, q, out, os, len, rcon, rs, NULL, 0);
}
int imcvtEncodeBatchHEVCGreenMulti(const uint8_t *f, size_t fs, int n, uint32_t h, uint32_t w, int q, uint8_t *out, size_t os, size_t *len,
                                   const int *devices, int ndev)
{
    (void)devices;
    logcall("hevc-green", n, 1, h, w, q, ndev);
    for (int i = 0; i < n; i++) {
        uint8_t *g = (uint8_t *)malloc((size_t)h * w + 1);
        for (size_t k = 0; k < (size_t)h * w; k++) g[k] = f[(size_t)i * fs + 3 * k + 1];
        emit(out + (size_t)i * os, &len[i], 'H', q, 1, h, w, g, (size_t)h * w);
        free(g);
    }
    return 0;
}
int imcvtEncodeBatchHEVCGreen(const uint8_t *f, size_t fs, int n, uint32_t h, uint32_t w, int q, uint8_t *out, size_t os, size_t *len)
{
    return imcvtEncodeBatchHEVCGreenMulti(f, fs, n, h, w, q, out, os, len, NULL, 0);
}
