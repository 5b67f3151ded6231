// api.cu -- extern "C" entry points of libimcvt_b200.so (declared in include/imcvt_b200.h)
#include "common.cuh"
#include "../../include/imcvt_b200.h"
#include <string.h>
#include <vector>

namespace imcvt {
namespace jls {
int encode_device(const uint8_t *d_frames, size_t frame_stride, int n, int is_rgb, uint32_t H, uint32_t W, int near_,
                  uint8_t *d_out, size_t out_stride, uint64_t *d_out_len, int *d_overflow, cudaStream_t st, Workspace *wsp = nullptr);
}
namespace hevc {
int encode_device(const uint8_t *d_frames, size_t frame_stride, int n, uint32_t H, uint32_t W, int qpd6,
                  uint8_t *d_out, size_t out_stride, uint64_t *d_out_len, uint8_t *d_rcon, size_t rcon_stride,
                  int *d_overflow, cudaStream_t st);
}
}  // namespace imcvt

using namespace imcvt;

static bool jls_dims_ok(uint32_t h, uint32_t w) { return w >= 1 && w <= 32767 && h >= 1 && h <= 32767; }   // imageio_jls.c:436

static inline uint32_t pad32(uint32_t v) { uint32_t c = v > 8192 ? 8192 : v; return (c + 31) / 32 * 32; }   // HEVCe.c:1580-1581

// shared host-buffer driver: H2D, encode, D2H of lengths then of exactly the produced bytes
template <class EncodeFn>
static int run_host_batch(const uint8_t *frames, size_t frame_stride, size_t frame_bytes, int n,
                          uint8_t *out, size_t out_stride, size_t bound, size_t *out_len,
                          uint8_t *rcon, size_t rcon_stride, size_t rcon_bytes, EncodeFn encode)
{
    Engine &E = engine();
    cudaStream_t st = E.stream;
    const size_t dev_in_stride  = (frame_bytes + 255) & ~(size_t)255;
    const size_t dev_out_stride = ((out_stride < bound ? out_stride : bound) + 255) & ~(size_t)255;
    const size_t dev_rc_stride  = rcon ? ((rcon_bytes + 255) & ~(size_t)255) : 0;
    if (E.io.reserve((size_t)n * (dev_in_stride + dev_out_stride + dev_rc_stride) + (size_t)n * 8 + 4096)) return 1;
    uint8_t  *d_in   = E.io.take<uint8_t>((size_t)n * dev_in_stride);
    uint8_t  *d_out  = E.io.take<uint8_t>((size_t)n * dev_out_stride);
    uint8_t  *d_rc   = rcon ? E.io.take<uint8_t>((size_t)n * dev_rc_stride) : nullptr;
    uint64_t *d_len  = E.io.take<uint64_t>((size_t)n);
    int      *d_flag = E.io.take<int>(1);
    if (!d_in || !d_out || !d_len || !d_flag) { set_error("io arena carve failed"); return 1; }

    IMCVT_CUDA(cudaMemcpy2DAsync(d_in, dev_in_stride, frames, frame_stride, frame_bytes, n, cudaMemcpyHostToDevice, st));
    IMCVT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    const size_t usable = out_stride < dev_out_stride ? out_stride : dev_out_stride;
    if (encode(d_in, dev_in_stride, d_out, dev_out_stride, usable, d_len, d_rc, dev_rc_stride, d_flag, st)) return 1;

    std::vector<uint64_t> lens(n);
    int flag = 0;
    IMCVT_CUDA(cudaMemcpyAsync(lens.data(), d_len, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    IMCVT_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    IMCVT_CUDA(cudaStreamSynchronize(st));
    if (flag) { set_error("output slot too small for at least one image (out_stride = %zu)", out_stride); return 1; }
    for (int i = 0; i < n; i++) {
        if (lens[i] == 0 || lens[i] > out_stride) { set_error("image %d: encoder produced an invalid length %llu", i, (unsigned long long)lens[i]); return 1; }
        out_len[i] = (size_t)lens[i];
        IMCVT_CUDA(cudaMemcpyAsync(out + (size_t)i * out_stride, d_out + (size_t)i * dev_out_stride, (size_t)lens[i], cudaMemcpyDeviceToHost, st));
    }
    if (rcon)
        IMCVT_CUDA(cudaMemcpy2DAsync(rcon, rcon_stride, d_rc, dev_rc_stride, rcon_bytes, n, cudaMemcpyDeviceToHost, st));
    IMCVT_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// JPEG-LS host batches of 8 frames or more: the batch is cut into NCHUNK chunks, each on a stream of its own with a
// workspace of its own, so that the upload of chunk c+1 and the download of chunk c-1 overlap the kernel of chunk c
// (a plane occupies one or two SMs for the whole kernel, so chunk kernels run side by side on different SMs).
static int run_host_jls_chunked(const uint8_t *frames, size_t frame_stride, size_t frame_bytes, int n, int is_rgb,
                                uint32_t height, uint32_t width, int near_, uint8_t *out, size_t out_stride, size_t bound, size_t *out_len)
{
    Engine &E = engine();
    constexpr int NCH = Engine::NCHUNK;
    for (int c = 0; c < NCH; c++)
        if (!E.chunk_stream[c]) IMCVT_CUDA(cudaStreamCreateWithFlags(&E.chunk_stream[c], cudaStreamNonBlocking));
    const size_t dev_in_stride  = (frame_bytes + 255) & ~(size_t)255;
    const size_t dev_out_stride = ((out_stride < bound ? out_stride : bound) + 255) & ~(size_t)255;
    if (E.io.reserve((size_t)n * (dev_in_stride + dev_out_stride) + (size_t)n * 8 + 4096)) return 1;
    uint8_t  *d_in   = E.io.take<uint8_t>((size_t)n * dev_in_stride);
    uint8_t  *d_out  = E.io.take<uint8_t>((size_t)n * dev_out_stride);
    uint64_t *d_len  = E.io.take<uint64_t>((size_t)n);
    int      *d_flag = E.io.take<int>(1);
    if (!d_in || !d_out || !d_len || !d_flag) { set_error("io arena carve failed"); return 1; }
    IMCVT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), E.stream));
    IMCVT_CUDA(cudaStreamSynchronize(E.stream));

    const int per = (n + NCH - 1) / NCH;
    // the lengths come back into PINNED memory: a copy into pageable memory would block the host until the chunk's kernel
    // has finished and serialise the chunks
    if (E.pinned_lens_cap < n) {
        if (E.pinned_lens) cudaFreeHost(E.pinned_lens);
        E.pinned_lens = nullptr; E.pinned_lens_cap = 0;
        IMCVT_CUDA(cudaHostAlloc((void **)&E.pinned_lens, (size_t)n * 8, cudaHostAllocDefault));
        E.pinned_lens_cap = n;
    }
    uint64_t *lens = E.pinned_lens;
    int nch = 0;
    for (int c = 0; c < NCH; c++) {
        const int first = c * per, cnt = n - first < per ? n - first : per;
        if (cnt <= 0) break;
        nch = c + 1;
        cudaStream_t st = E.chunk_stream[c];
        IMCVT_CUDA(cudaMemcpy2DAsync(d_in + (size_t)first * dev_in_stride, dev_in_stride, frames + (size_t)first * frame_stride, frame_stride,
                                     frame_bytes, cnt, cudaMemcpyHostToDevice, st));
        if (jls::encode_device(d_in + (size_t)first * dev_in_stride, dev_in_stride, cnt, is_rgb, height, width, near_,
                               d_out + (size_t)first * dev_out_stride, dev_out_stride, d_len + first, d_flag, st, &E.ws_chunk[c])) return 1;
        IMCVT_CUDA(cudaMemcpyAsync(lens + first, d_len + first, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
    }
    for (int c = 0; c < nch; c++) {
        const int first = c * per, cnt = n - first < per ? n - first : per;
        cudaStream_t st = E.chunk_stream[c];
        IMCVT_CUDA(cudaStreamSynchronize(st));
        for (int i = first; i < first + cnt; i++) {
            if (lens[i] == 0 || lens[i] > out_stride) {
                for (int k = 0; k < nch; k++) cudaStreamSynchronize(E.chunk_stream[k]);
                set_error("image %d: encoder produced an invalid length %llu (output slot too small?)", i, (unsigned long long)lens[i]);
                return 1;
            }
            out_len[i] = (size_t)lens[i];
            IMCVT_CUDA(cudaMemcpyAsync(out + (size_t)i * out_stride, d_out + (size_t)i * dev_out_stride, (size_t)lens[i], cudaMemcpyDeviceToHost, st));
        }
    }
    for (int c = 0; c < nch; c++) IMCVT_CUDA(cudaStreamSynchronize(E.chunk_stream[c]));
    int flag = 0;
    IMCVT_CUDA(cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) { set_error("output slot too small for at least one image (out_stride = %zu)", out_stride); return 1; }
    return 0;
}

extern "C" {

size_t imcvtJLSBound(int is_rgb, uint32_t height, uint32_t width)
{
    const size_t npix = (size_t)height * width;
    const size_t ncomp = is_rgb ? 3 : 1;
    // <= 32 code bits per sample (LIMIT), <= 1 stuffed bit per 7 payload bits, 10-byte SOS per scan
    return 32 + ncomp * (10 + 4 * npix + 4 * npix / 7 + 16) + 2;
}

size_t imcvtHEVCBound(uint32_t height, uint32_t width)
{
    return 2 * ((size_t)width + 32) * ((size_t)height + 32) + 65536;           // imageio_hevc.c:14
}

int imcvtEncodeBatchJLSDevice(const void *d_frames, size_t frame_stride, int n, int is_rgb,
                              uint32_t height, uint32_t width, int near_,
                              void *d_out, size_t out_stride, void *d_out_len, void *stream)
{
    if (n <= 0) { set_error("batch size must be positive"); return 1; }
    if (!jls_dims_ok(height, width)) { set_error("JPEG-LS dimensions must be 1..32767 (got %ux%u)", width, height); return 1; }
    if (near_ < 0 || near_ > 4) { set_error("NEAR must be 0..4 (got %d)", near_); return 1; }
    if (ensure_engine()) return 1;
    Engine &E = engine();
    if (!E.d_flag) { IMCVT_CUDA(cudaMalloc(&E.d_flag, sizeof(int))); }
    int *d_flag = E.d_flag;
    cudaStream_t st = (cudaStream_t)stream;
    IMCVT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    return jls::encode_device((const uint8_t *)d_frames, frame_stride, n, is_rgb, height, width, near_,
                              (uint8_t *)d_out, out_stride, (uint64_t *)d_out_len, d_flag, st);
}

int imcvtEncodeBatchJLS(const uint8_t *frames, size_t frame_stride, int n, int is_rgb,
                        uint32_t height, uint32_t width, int near_,
                        uint8_t *out, size_t out_stride, size_t *out_len)
{
    if (n <= 0 || !frames || !out || !out_len) { set_error("bad arguments"); return 1; }
    if (!jls_dims_ok(height, width)) { set_error("JPEG-LS dimensions must be 1..32767 (got %ux%u)", width, height); return 1; }
    if (near_ < 0 || near_ > 4) { set_error("NEAR must be 0..4 (got %d)", near_); return 1; }
    const size_t frame_bytes = (size_t)height * width * (is_rgb ? 3 : 1);
    if (frame_stride < frame_bytes) { set_error("frame_stride smaller than one frame"); return 1; }
    if (ensure_engine()) return 1;
    if (n >= 8)
        return run_host_jls_chunked(frames, frame_stride, frame_bytes, n, is_rgb, height, width, near_, out, out_stride,
                                    imcvtJLSBound(is_rgb, height, width), out_len);
    auto enc = [&](uint8_t *d_in, size_t in_stride, uint8_t *d_out, size_t dev_out_stride, size_t usable, uint64_t *d_len,
                   uint8_t *, size_t, int *d_flag, cudaStream_t st) {
        (void)usable;
        return jls::encode_device(d_in, in_stride, n, is_rgb, height, width, near_, d_out, dev_out_stride, d_len, d_flag, st);
    };
    return run_host_batch(frames, frame_stride, frame_bytes, n, out, out_stride, imcvtJLSBound(is_rgb, height, width), out_len,
                          nullptr, 0, 0, enc);
}

int imcvtEncodeBatchHEVCDevice(const void *d_frames, size_t frame_stride, int n,
                               uint32_t height, uint32_t width, int qpd6,
                               void *d_out, size_t out_stride, void *d_out_len,
                               void *d_rcon, size_t rcon_stride, void *stream)
{
    if (n <= 0) { set_error("batch size must be positive"); return 1; }
    if (height < 1 || width < 1) { set_error("empty image"); return 1; }
    if (qpd6 < 0 || qpd6 > 4) { set_error("qpd6 must be 0..4 (got %d)", qpd6); return 1; }
    if (ensure_engine()) return 1;
    Engine &E = engine();
    if (!E.d_flag) { IMCVT_CUDA(cudaMalloc(&E.d_flag, sizeof(int))); }
    int *d_flag = E.d_flag;
    cudaStream_t st = (cudaStream_t)stream;
    IMCVT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    return hevc::encode_device((const uint8_t *)d_frames, frame_stride, n, height, width, qpd6,
                               (uint8_t *)d_out, out_stride, (uint64_t *)d_out_len, (uint8_t *)d_rcon, rcon_stride, d_flag, st);
}

int imcvtEncodeBatchHEVC(const uint8_t *frames, size_t frame_stride, int n,
                         uint32_t height, uint32_t width, int qpd6,
                         uint8_t *out, size_t out_stride, size_t *out_len,
                         uint8_t *rcon, size_t rcon_stride)
{
    if (n <= 0 || !frames || !out || !out_len) { set_error("bad arguments"); return 1; }
    if (height < 1 || width < 1) { set_error("empty image"); return 1; }
    if (qpd6 < 0 || qpd6 > 4) { set_error("qpd6 must be 0..4 (got %d)", qpd6); return 1; }
    const size_t frame_bytes = (size_t)height * width;
    if (frame_stride < frame_bytes) { set_error("frame_stride smaller than one frame"); return 1; }
    const size_t rcon_bytes = (size_t)pad32(height) * pad32(width);
    if (rcon && rcon_stride < rcon_bytes) { set_error("rcon_stride smaller than one padded picture"); return 1; }
    if (ensure_engine()) return 1;
    auto enc = [&](uint8_t *d_in, size_t in_stride, uint8_t *d_out, size_t dev_out_stride, size_t usable, uint64_t *d_len,
                   uint8_t *d_rc, size_t rc_stride, int *d_flag, cudaStream_t st) {
        (void)usable;
        return hevc::encode_device(d_in, in_stride, n, height, width, qpd6, d_out, dev_out_stride, d_len, d_rc, rc_stride, d_flag, st);
    };
    return run_host_batch(frames, frame_stride, frame_bytes, n, out, out_stride, imcvtHEVCBound(height, width), out_len,
                          rcon, rcon_stride, rcon_bytes, enc);
}

}  // extern "C"
