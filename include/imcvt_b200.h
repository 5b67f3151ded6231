/* imcvt_b200.h -- C-ABI of the B200-native encode engine (libimcvt_b200.so).
 *
 * Drop-in boundary for the two compute-bound paths of WangXuan95/ImCvt:
 *   - JPEG-LS encoder   (reference: src/imageio_jls.c)
 *   - H.265 intra encoder (reference: src/HEVCe/HEVCe.c, wrapper src/imageio_hevc.c)
 *
 * Plain C linkage, plain pointers and sizes, no torch / C++ types.  All functions are synchronous at
 * return unless they take a `stream`.  Every function returns 0 on success and non-zero on failure
 * (the reference's convention, src/imageio.h:15); imcvtLastError() gives the reason.  There is NO CPU
 * fallback: without a CUDA device every encode entry point fails.
 *
 * The reference-compatible file API (writeJLSImageFile / writeHEVCImageFile / HEVCImageEncoder) is
 * declared in imcvt_imageio.h and exported by the same library.
 */
#ifndef IMCVT_B200_H
#define IMCVT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- life cycle ------------------------------------------------------------------------------ */
int         imcvtInit(int device);          /* optional: binds the engine to a CUDA device (default: current) */
void        imcvtShutdown(void);            /* frees the device workspaces (of every device the multi-device calls used) */
const char *imcvtLastError(void);           /* message of the last failure on this thread */
int         imcvtVersion(void);

/* ---- sizes ----------------------------------------------------------------------------------- */
/* Upper bound of one encoded JPEG-LS image: the reference's own allocation bound 8*W*H+65536
 * (src/imageio_jls.c:439) is far above what 8-bit data can produce; the engine guarantees
 * len <= imcvtJLSBound().  is_rgb: 0 gray8, 1 RGB24. */
size_t imcvtJLSBound(int is_rgb, uint32_t height, uint32_t width);
/* Upper bound of one encoded HEVC picture = the reference's stream allocation (src/imageio_hevc.c:14). */
size_t imcvtHEVCBound(uint32_t height, uint32_t width);

/* ---- JPEG-LS, batch, host buffers (replaces N calls of writeJLSImageFile's codec part,
 *      src/imageio_jls.c:428-465: same pixels in, same .jls bytes out) -------------------------------
 * frames      : n images, image i at frames + i*frame_stride, interleaved u8 H*W (gray) or H*W*3 (RGB)
 * near        : 0..4 (NEAR); 0 = lossless
 * out         : n output slots, slot i at out + i*out_stride, each >= imcvtJLSBound() bytes is always enough
 * out_len[i]  : bytes produced for image i
 * Bad dimensions (outside 1..32767, src/imageio_jls.c:436) or a too small slot fail the call. */
int imcvtEncodeBatchJLS(const uint8_t *frames, size_t frame_stride, int n, int is_rgb,
                        uint32_t height, uint32_t width, int near_,
                        uint8_t *out, size_t out_stride, size_t *out_len);

/* Same, but `d_frames`, `d_out`, `d_out_len` (uint64 per image) are DEVICE pointers and the work is queued
 * on `stream` (a cudaStream_t passed as void*; NULL = default stream) without synchronising. */
int imcvtEncodeBatchJLSDevice(const void *d_frames, size_t frame_stride, int n, int is_rgb,
                              uint32_t height, uint32_t width, int near_,
                              void *d_out, size_t out_stride, void *d_out_len, void *stream);

/* ---- H.265 intra, batch (replaces N calls of HEVCImageEncoder, src/HEVCe/HEVCe.h:5-12) ---------------
 * frames      : n gray8 images H*W, image i at frames + i*frame_stride
 * qpd6        : 0..4  (QP = 6*qpd6 + 4, src/HEVCe/HEVCe.c:727)
 * out/out_len : as above, slot >= imcvtHEVCBound()
 * rcon        : optional (may be NULL): n reconstructed pictures of the PADDED size
 *               (ceil32(H) x ceil32(W)), picture i at rcon + i*rcon_stride */
int imcvtEncodeBatchHEVC(const uint8_t *frames, size_t frame_stride, int n,
                         uint32_t height, uint32_t width, int qpd6,
                         uint8_t *out, size_t out_stride, size_t *out_len,
                         uint8_t *rcon, size_t rcon_stride);

/* Same as imcvtEncodeBatchHEVC without rcon, for interleaved RGB24 frames: the engine encodes the GREEN channel, which is
 * what the reference's wrapper does with RGB input (src/imageio_hevc.c:21-26; the caller prints its warning line).  The
 * channel is picked on the device by a small kernel instead of a host loop over every pixel. */
int imcvtEncodeBatchHEVCGreen(const uint8_t *rgb_frames, size_t frame_stride, int n,
                              uint32_t height, uint32_t width, int qpd6,
                              uint8_t *out, size_t out_stride, size_t *out_len);

/* A quantiser sweep in ONE launch (BASELINE configs[4]: every frame at several QPs): frame i is coded at each of the nq
 * settings qpd6[0..nq-1] (1 <= nq <= 8, each 0..4); the stream of (frame i, setting k) lands in slot i*nq + k of out /
 * out_len.  Same bytes as nq calls of imcvtEncodeBatchHEVC, but the n*nq jobs share the GPU: where a batch of n frames leaves
 * SMs idle (n < SMs/2), the sweep's jobs fill them. */
int imcvtEncodeBatchHEVCSweep(const uint8_t *frames, size_t frame_stride, int n,
                              uint32_t height, uint32_t width, const int *qpd6, int nq,
                              uint8_t *out, size_t out_stride, size_t *out_len);

int imcvtEncodeBatchHEVCDevice(const void *d_frames, size_t frame_stride, int n,
                               uint32_t height, uint32_t width, int qpd6,
                               void *d_out, size_t out_stride, void *d_out_len,
                               void *d_rcon, size_t rcon_stride, void *stream);

/* The device-buffer entry points above (imcvtEncodeBatchJLSDevice / imcvtEncodeBatchHEVCDevice) do not synchronise, so a slot
 * that is too small cannot fail the call: it shows as d_out_len[i] == 0 for that image (nothing is written past the slot).
 * Pointers must be device (or page-locked host) memory of the engine's device; frame_stride >= one frame. */

/* ---- every GPU of the box (replaces the reference's file-by-file loop, src/main.c:162-211, for a batch) ---------------
 * Same arguments and results as the single-device entry points; the batch is sharded over `devices[0..ndev-1]` (CUDA
 * ordinals; NULL / ndev <= 0 = all visible devices) in contiguous blocks of ceil(n / ndev) frames, one host thread and one
 * engine per device, no collective: frames are independent and the sizes land in out_len directly.  The bytes are those of
 * the single-device call whatever the number of devices.  Page-locked buffers (imcvtHostAlloc) are read and written by all
 * devices at full rate. */
int imcvtDeviceCount(void);
int imcvtEncodeBatchJLSMulti(const uint8_t *frames, size_t frame_stride, int n, int is_rgb,
                             uint32_t height, uint32_t width, int near_,
                             uint8_t *out, size_t out_stride, size_t *out_len, const int *devices, int ndev);
int imcvtEncodeBatchHEVCMulti(const uint8_t *frames, size_t frame_stride, int n,
                              uint32_t height, uint32_t width, int qpd6,
                              uint8_t *out, size_t out_stride, size_t *out_len,
                              uint8_t *rcon, size_t rcon_stride, const int *devices, int ndev);
int imcvtEncodeBatchHEVCGreenMulti(const uint8_t *rgb_frames, size_t frame_stride, int n,
                                   uint32_t height, uint32_t width, int qpd6,
                                   uint8_t *out, size_t out_stride, size_t *out_len, const int *devices, int ndev);
int imcvtEncodeBatchHEVCSweepMulti(const uint8_t *frames, size_t frame_stride, int n,
                                   uint32_t height, uint32_t width, const int *qpd6, int nq,
                                   uint8_t *out, size_t out_stride, size_t *out_len, const int *devices, int ndev);

/* How imcvtEncodeBatchJLS moves PAGE-LOCKED host buffers (cudaHostAlloc / imcvtHostAlloc / cudaHostRegister memory; pageable
 * buffers are always staged through device memory):
 *  -1 (default) = in place when the batch is one the in-place kernel is built for (gray, NEAR 0, planes up to 6144 wide, at
 *       most SMs/2 frames: the cluster-pair kernel requests rows over PCIe a tick ahead of their use and writes the stream in
 *       whole 128-byte lines), otherwise staged;
 *   0 = always upload, encode, download (chunked and overlapped);
 *   1 = frames read in place, streams downloaded;   2 = frames read and streams written in place, whatever the kernel. */
void imcvtJLSHostMode(int mode);

/* ---- staging memory ---------------------------------------------------------------------------- */
/* Page-locked host memory for the frames / out arguments of the host-buffer entry points: copies from and to it run at
 * full PCIe rate and overlap the kernels.  A loader that reads image files straight into such a buffer replaces the
 * reference's malloc per image (src/imageio_pnm.c:112) plus a staging copy.  NULL on failure (imcvtLastError()). */
void *imcvtHostAlloc(size_t bytes);
void  imcvtHostFree(void *p);

/* ---- instrumentation ------------------------------------------------------------------------- */
/* When enabled, every kernel launch of the engine is bracketed by CUDA events on its own stream;
 * imcvtProfileGet() returns launch count and summed device milliseconds per kernel name
 * ("jls_chain", "hevc_frame", ...).  Used by bench.py for the roofline line. */
void imcvtProfileEnable(int on);
void imcvtProfileReset(void);
int  imcvtProfileGet(const char *kernel, int *launches, double *total_ms);
int  imcvtProfileList(char *buf, size_t cap);        /* "name:launches:ms\n..." */
long imcvtLaunchCount(void);                         /* total kernels launched since ProfileReset */

/* Debug / cross-check: force the fully serial JPEG-LS scan kernel also for NEAR=0 (default 0). */
void imcvtDebugJLSForceSerial(int on);
/* Debug / cross-check of the NEAR=0 paths: 0 = default (planes up to 6144 wide: shared-memory tile kernel, on a
 * cluster of two CTAs per plane when the batch has at most SMs/2 planes, else one CTA per plane), 1 = force the
 * multi-kernel pipeline that wider planes use, 2 = force one CTA per plane, 3 = force the two-CTA cluster (rank 0
 * produces sorted tiles into the shared memory of rank 1), 4 = NEAR > 0 / forced-serial scans on the plain
 * one-thread-per-plane kernel instead of the row-pipelined one, 5 = NEAR > 0 on the three-warp row scan instead of the
 * candidate-table kernel, 6 = default launch shapes, but the tile kernels stage their rows through registers instead of
 * bulk asynchronous copies (cp.async.bulk). */
void imcvtDebugJLSPath(int path);
/* Debug / cross-check of the H.265 paths: 0 = default (a cluster of two CTAs per frame when the batch has at most SMs/2
 * frames, else one CTA per frame), 1 = force one CTA per frame, 2 = force the two-CTA cluster; + 4 = search the 8x8 units with
 * all warps in lock step, + 8 = with three warp groups side by side (default: lock step on pairs, groups on single CTAs). */
void imcvtDebugHEVCPath(int path);

#ifdef __cplusplus
}
#endif
#endif /* IMCVT_B200_H */
