// api.cu -- extern "C" entry points of libimcvt_b200.so (declared in include/imcvt_b200.h)
#include "common.cuh"
#include "../../include/imcvt_b200.h"
#include <string.h>
#include <vector>

namespace imcvt {
namespace jls {
int encode_device(const uint8_t *d_frames, size_t frame_stride, int n, int is_rgb, uint32_t H, uint32_t W, int near_,
                  uint8_t *d_out, size_t out_stride, uint64_t *d_out_len, int *d_overflow, cudaStream_t st, Workspace *wsp = nullptr,
                  bool out_is_host = false);     // out_is_host: d_out aliases page-locked host memory (whole-line stores)
}
namespace jls { bool host_in_place_ok(int n, int is_rgb, uint32_t W, int near_); }
namespace hevc {
int encode_device(const uint8_t *d_frames, size_t frame_stride, int n, uint32_t H, uint32_t W, int qpd6,
                  uint8_t *d_out, size_t out_stride, uint64_t *d_out_len, uint8_t *d_rcon, size_t rcon_stride,
                  int *d_overflow, cudaStream_t st, const int *qlist = nullptr, int nq = 0);
}
}  // namespace imcvt

using namespace imcvt;

static bool jls_dims_ok(uint32_t h, uint32_t w) { return w >= 1 && w <= 32767 && h >= 1 && h <= 32767; }   // imageio_jls.c:436

static inline uint32_t pad32(uint32_t v) { uint32_t c = v > 8192 ? 8192 : v; return (c + 31) / 32 * 32; }   // HEVCe.c:1580-1581

// Green plane of interleaved RGB24 frames (replaces the host loop of the reference's wrapper, src/imageio_hevc.c:24-26):
// one thread per four pixels, three aligned 32-bit loads -> one 32-bit store.  Frame bases are 256-byte aligned.
__global__ void green_plane_kernel(const uint8_t *rgb, size_t rgb_stride, uint8_t *gray, size_t gray_stride, size_t npix, int n)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;         // group of four pixels
    for (int f = blockIdx.y; f < n; f += gridDim.y) {                       // grid.y is capped at 65535: frames beyond it loop
        const uint8_t *src = rgb + (size_t)f * rgb_stride;
        uint8_t *dst = gray + (size_t)f * gray_stride;
        if (4 * g + 4 <= npix) {
            const uint32_t *s = reinterpret_cast<const uint32_t *>(src) + 3 * g;
            const uint32_t a = s[0], b = s[1], c = s[2];                   // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
            reinterpret_cast<uint32_t *>(dst)[g] = ((a >> 8) & 0xFF) | ((b & 0xFF) << 8) | ((b >> 24) << 16) | (((c >> 16) & 0xFF) << 24);
        } else {
            for (size_t p = 4 * g; p < npix; p++) dst[p] = src[3 * p + 1];
        }
    }
}

// limits the kernels index safely: frames per call, H.265 picture dimensions (int arithmetic on H * W inside the kernels)
static bool batch_ok(int n) { if (n <= 0 || n > (1 << 24)) { set_error("batch size must be 1..16777216 (got %d)", n); return false; } return true; }
static bool hevc_dims_ok(uint32_t h, uint32_t w)
{
    if (h < 1 || w < 1) { set_error("empty image"); return false; }
    if (h > (1u << 20) || w > (1u << 20) || (uint64_t)h * w > (1ull << 31) - 1) { set_error("H.265 picture %ux%u is beyond what the engine indexes (H * W < 2^31)", w, h); return false; }
    return true;
}

// The device-buffer entry points take one of the engine's overflow flags in turn (calls on different streams do not share
// one) and check what they can about the caller's pointers: non-null, on the engine's device.
static int *next_device_flag(Engine &E)
{
    if (!E.d_flag && cudaMalloc(&E.d_flag, Engine::NFLAG * sizeof(int)) != cudaSuccess) { (void)cudaGetLastError(); E.d_flag = nullptr; return nullptr; }
    return E.d_flag + (E.flag_next++ % Engine::NFLAG);
}
static bool on_engine_device(Engine &E, const void *p, const char *what)
{
    cudaPointerAttributes a;
    if (!p) { set_error("%s is NULL", what); return false; }
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); set_error("%s is not a CUDA pointer", what); return false; }
    if (a.type == cudaMemoryTypeDevice && a.device != E.device) { set_error("%s lives on device %d, the engine on device %d", what, a.device, E.device); return false; }
    if (a.type == cudaMemoryTypeUnregistered) { set_error("%s is pageable host memory, not a device pointer", what); return false; }
    return true;
}

// shared host-buffer driver: H2D, encode, D2H of lengths then of exactly the produced bytes
template <class EncodeFn>
static int run_host_batch(const uint8_t *frames, size_t frame_stride, size_t frame_bytes, int n,
                          uint8_t *out, size_t out_stride, size_t bound, size_t *out_len,
                          uint8_t *rcon, size_t rcon_stride, size_t rcon_bytes, size_t scratch_stride, EncodeFn encode)
{
    Engine &E = engine();
    cudaStream_t st = E.stream;
    const size_t dev_in_stride  = (frame_bytes + 255) & ~(size_t)255;
    const size_t dev_out_stride = ((out_stride < bound ? out_stride : bound) + 255) & ~(size_t)255;
    const size_t dev_rc_stride  = rcon ? ((rcon_bytes + 255) & ~(size_t)255) : 0;
    if (E.io.reserve((size_t)n * (dev_in_stride + dev_out_stride + dev_rc_stride + scratch_stride) + (size_t)n * 8 + 4096)) return 1;
    uint8_t  *d_in   = E.io.take<uint8_t>((size_t)n * dev_in_stride);
    uint8_t  *d_scr  = scratch_stride ? E.io.take<uint8_t>((size_t)n * scratch_stride) : nullptr;   // per-frame scratch of the caller's encode step
    uint8_t  *d_out  = E.io.take<uint8_t>((size_t)n * dev_out_stride);
    uint8_t  *d_rc   = rcon ? E.io.take<uint8_t>((size_t)n * dev_rc_stride) : nullptr;
    uint64_t *d_len  = E.io.take<uint64_t>((size_t)n);
    int      *d_flag = E.io.take<int>(1);
    if (!d_in || !d_out || !d_len || !d_flag) { set_error("io arena carve failed"); return 1; }

    IMCVT_CUDA(cudaMemcpy2DAsync(d_in, dev_in_stride, frames, frame_stride, frame_bytes, n, cudaMemcpyHostToDevice, st));
    IMCVT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    const size_t usable = out_stride < dev_out_stride ? out_stride : dev_out_stride;
    if (encode(d_in, dev_in_stride, d_out, dev_out_stride, usable, d_len, d_rc, dev_rc_stride, d_flag, d_scr, st)) return 1;

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
static int run_host_jls_chunked_impl(const uint8_t *frames, size_t frame_stride, size_t frame_bytes, int n, int is_rgb,
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

// Device-side alias of a page-locked host range (cudaHostAlloc / cudaHostRegister memory is mapped under unified addressing),
// or nullptr when the range is ordinary pageable memory.
static void *mapped_alias(const void *p, size_t bytes)
{
    cudaPointerAttributes a, b;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess || cudaPointerGetAttributes(&b, (const char *)p + (bytes ? bytes - 1 : 0)) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    if (a.type != cudaMemoryTypeHost || b.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return a.devicePointer;
}

// Page-locked input: the kernels read the frames IN PLACE over PCIe (each sample is read once by the producers, a second
// time out of L2 by the run-event stage), so there is no upload in front of the kernel -- the pixels cross the bus while
// the planes are being coded.  zero_copy_out: the streams are written straight into the caller's page-locked slots as well.
static int run_host_jls_mapped(const uint8_t *d_frames, size_t frame_stride, int n, int is_rgb, uint32_t height, uint32_t width, int near_,
                               uint8_t *out, uint8_t *d_out_mapped, size_t out_stride, size_t bound, size_t *out_len)
{
    Engine &E = engine();
    cudaStream_t st = E.stream;
    const size_t dev_out_stride = d_out_mapped ? out_stride : ((out_stride < bound ? out_stride : bound) + 255) & ~(size_t)255;
    if (E.io.reserve((d_out_mapped ? 0 : (size_t)n * dev_out_stride) + (size_t)n * 8 + 4096)) return 1;
    uint8_t  *d_out  = d_out_mapped ? d_out_mapped : E.io.take<uint8_t>((size_t)n * dev_out_stride);
    uint64_t *d_len  = E.io.take<uint64_t>((size_t)n);
    int      *d_flag = E.io.take<int>(1);
    if (!d_out || !d_len || !d_flag) { set_error("io arena carve failed"); return 1; }
    if (E.pinned_lens_cap < n + 1) {
        if (E.pinned_lens) cudaFreeHost(E.pinned_lens);
        E.pinned_lens = nullptr; E.pinned_lens_cap = 0;
        IMCVT_CUDA(cudaHostAlloc((void **)&E.pinned_lens, (size_t)(n + 1) * 8, cudaHostAllocDefault));
        E.pinned_lens_cap = n + 1;
    }
    uint64_t *lens = E.pinned_lens;
    IMCVT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    if (jls::encode_device(d_frames, frame_stride, n, is_rgb, height, width, near_, d_out, dev_out_stride, d_len, d_flag, st, nullptr,
                           d_out_mapped != nullptr)) return 1;
    IMCVT_CUDA(cudaMemcpyAsync(lens, d_len, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    IMCVT_CUDA(cudaMemcpyAsync(lens + n, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    IMCVT_CUDA(cudaStreamSynchronize(st));
    if ((int)lens[n]) { set_error("output slot too small for at least one image (out_stride = %zu)", out_stride); return 1; }
    for (int i = 0; i < n; i++) {
        if (lens[i] == 0 || lens[i] > out_stride) { set_error("image %d: encoder produced an invalid length %llu", i, (unsigned long long)lens[i]); return 1; }
        out_len[i] = (size_t)lens[i];
        if (!d_out_mapped)
            IMCVT_CUDA(cudaMemcpyAsync(out + (size_t)i * out_stride, d_out + (size_t)i * dev_out_stride, (size_t)lens[i], cudaMemcpyDeviceToHost, st));
    }
    if (!d_out_mapped) IMCVT_CUDA(cudaStreamSynchronize(st));
    return 0;
}

static int run_host_jls_chunked(const uint8_t *frames, size_t frame_stride, size_t frame_bytes, int n, int is_rgb,
                                uint32_t height, uint32_t width, int near_, uint8_t *out, size_t out_stride, size_t bound, size_t *out_len)
{
    const int rc = run_host_jls_chunked_impl(frames, frame_stride, frame_bytes, n, is_rgb, height, width, near_, out, out_stride, bound, out_len);
    if (rc) {
        // a failed call must not leave chunk streams running on the arena and the workspaces the next call will carve again
        Engine &E = engine();
        for (int c = 0; c < Engine::NCHUNK; c++)
            if (E.chunk_stream[c]) cudaStreamSynchronize(E.chunk_stream[c]);
        (void)cudaGetLastError();
    }
    return rc;
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
    IMCVT_API_LOCK();
    if (!batch_ok(n)) return 1;
    if (!jls_dims_ok(height, width)) { set_error("JPEG-LS dimensions must be 1..32767 (got %ux%u)", width, height); return 1; }
    if (near_ < 0 || near_ > 4) { set_error("NEAR must be 0..4 (got %d)", near_); return 1; }
    if (frame_stride < (size_t)height * width * (is_rgb ? 3 : 1)) { set_error("frame_stride smaller than one frame"); return 1; }
    if (ensure_engine()) return 1;
    Engine &E = engine();
    if (!on_engine_device(E, d_frames, "d_frames") || !on_engine_device(E, d_out, "d_out") || !on_engine_device(E, d_out_len, "d_out_len")) return 1;
    int *d_flag = next_device_flag(E);
    if (!d_flag) { set_error("device flag allocation failed"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    IMCVT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    return jls::encode_device((const uint8_t *)d_frames, frame_stride, n, is_rgb, height, width, near_,
                              (uint8_t *)d_out, out_stride, (uint64_t *)d_out_len, d_flag, st);
}

int imcvtEncodeBatchJLS(const uint8_t *frames, size_t frame_stride, int n, int is_rgb,
                        uint32_t height, uint32_t width, int near_,
                        uint8_t *out, size_t out_stride, size_t *out_len)
{
    IMCVT_API_LOCK();
    if (!frames || !out || !out_len) { set_error("bad arguments"); return 1; }
    if (!batch_ok(n)) return 1;
    if (!jls_dims_ok(height, width)) { set_error("JPEG-LS dimensions must be 1..32767 (got %ux%u)", width, height); return 1; }
    if (near_ < 0 || near_ > 4) { set_error("NEAR must be 0..4 (got %d)", near_); return 1; }
    const size_t frame_bytes = (size_t)height * width * (is_rgb ? 3 : 1);
    if (frame_stride < frame_bytes) { set_error("frame_stride smaller than one frame"); return 1; }
    if (ensure_engine()) return 1;
    // page-locked frames and slots: read and written in place by the kernels when the batch runs on the launch shape built
    // for it (mode -1, the default), or whenever the caller says so (imcvtJLSHostMode 1 / 2)
    int mode = engine().jls_host_mode;
    if (mode < 0) mode = jls::host_in_place_ok(n, is_rgb, width, near_) ? 2 : 0;
    if (mode > 0) {
        void *d_in = mapped_alias(frames, (size_t)(n - 1) * frame_stride + frame_bytes);
        void *d_o = mode > 1 ? mapped_alias(out, (size_t)n * out_stride) : nullptr;
        if (d_in)
            return run_host_jls_mapped((const uint8_t *)d_in, frame_stride, n, is_rgb, height, width, near_, out, (uint8_t *)d_o,
                                       out_stride, imcvtJLSBound(is_rgb, height, width), out_len);
    }
    if (n >= 8)
        return run_host_jls_chunked(frames, frame_stride, frame_bytes, n, is_rgb, height, width, near_, out, out_stride,
                                    imcvtJLSBound(is_rgb, height, width), out_len);
    auto enc = [&](uint8_t *d_in, size_t in_stride, uint8_t *d_out, size_t dev_out_stride, size_t usable, uint64_t *d_len,
                   uint8_t *, size_t, int *d_flag, uint8_t *, cudaStream_t st) {
        (void)usable;
        return jls::encode_device(d_in, in_stride, n, is_rgb, height, width, near_, d_out, dev_out_stride, d_len, d_flag, st);
    };
    return run_host_batch(frames, frame_stride, frame_bytes, n, out, out_stride, imcvtJLSBound(is_rgb, height, width), out_len,
                          nullptr, 0, 0, 0, enc);
}

int imcvtEncodeBatchHEVCDevice(const void *d_frames, size_t frame_stride, int n,
                               uint32_t height, uint32_t width, int qpd6,
                               void *d_out, size_t out_stride, void *d_out_len,
                               void *d_rcon, size_t rcon_stride, void *stream)
{
    IMCVT_API_LOCK();
    if (!batch_ok(n) || !hevc_dims_ok(height, width)) return 1;
    if (qpd6 < 0 || qpd6 > 4) { set_error("qpd6 must be 0..4 (got %d)", qpd6); return 1; }
    if (frame_stride < (size_t)height * width) { set_error("frame_stride smaller than one frame"); return 1; }
    if (ensure_engine()) return 1;
    Engine &E = engine();
    if (!on_engine_device(E, d_frames, "d_frames") || !on_engine_device(E, d_out, "d_out") || !on_engine_device(E, d_out_len, "d_out_len")) return 1;
    if (d_rcon && !on_engine_device(E, d_rcon, "d_rcon")) return 1;
    int *d_flag = next_device_flag(E);
    if (!d_flag) { set_error("device flag allocation failed"); return 1; }
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
    IMCVT_API_LOCK();
    if (!frames || !out || !out_len) { set_error("bad arguments"); return 1; }
    if (!batch_ok(n) || !hevc_dims_ok(height, width)) return 1;
    if (qpd6 < 0 || qpd6 > 4) { set_error("qpd6 must be 0..4 (got %d)", qpd6); return 1; }
    const size_t frame_bytes = (size_t)height * width;
    if (frame_stride < frame_bytes) { set_error("frame_stride smaller than one frame"); return 1; }
    const size_t rcon_bytes = (size_t)pad32(height) * pad32(width);
    if (rcon && rcon_stride < rcon_bytes) { set_error("rcon_stride smaller than one padded picture"); return 1; }
    if (ensure_engine()) return 1;
    auto enc = [&](uint8_t *d_in, size_t in_stride, uint8_t *d_out, size_t dev_out_stride, size_t usable, uint64_t *d_len,
                   uint8_t *d_rc, size_t rc_stride, int *d_flag, uint8_t *, cudaStream_t st) {
        (void)usable;
        return hevc::encode_device(d_in, in_stride, n, height, width, qpd6, d_out, dev_out_stride, d_len, d_rc, rc_stride, d_flag, st);
    };
    return run_host_batch(frames, frame_stride, frame_bytes, n, out, out_stride, imcvtHEVCBound(height, width), out_len,
                          rcon, rcon_stride, rcon_bytes, 0, enc);
}

int imcvtEncodeBatchHEVCGreen(const uint8_t *rgb_frames, size_t frame_stride, int n,
                              uint32_t height, uint32_t width, int qpd6,
                              uint8_t *out, size_t out_stride, size_t *out_len)
{
    IMCVT_API_LOCK();
    if (!rgb_frames || !out || !out_len) { set_error("bad arguments"); return 1; }
    if (!batch_ok(n) || !hevc_dims_ok(height, width)) return 1;
    if (qpd6 < 0 || qpd6 > 4) { set_error("qpd6 must be 0..4 (got %d)", qpd6); return 1; }
    const size_t npix = (size_t)height * width;
    if (frame_stride < 3 * npix) { set_error("frame_stride smaller than one RGB frame"); return 1; }
    if (ensure_engine()) return 1;
    const size_t gray_stride = (npix + 255) & ~(size_t)255;
    auto enc = [&](uint8_t *d_in, size_t in_stride, uint8_t *d_out, size_t dev_out_stride, size_t usable, uint64_t *d_len,
                   uint8_t *, size_t, int *d_flag, uint8_t *d_gray, cudaStream_t st) {
        (void)usable;
        {
            LaunchScope ls("green_plane", st);
            const unsigned groups = (unsigned)((npix + 3) / 4);             // npix < 2^31 (hevc_dims_ok)
            dim3 grid((groups + 255) / 256, (unsigned)(n < 65535 ? n : 65535));
            green_plane_kernel<<<grid, 256, 0, st>>>(d_in, in_stride, d_gray, gray_stride, npix, n);
        }
        IMCVT_CUDA(cudaGetLastError());
        return hevc::encode_device(d_gray, gray_stride, n, height, width, qpd6, d_out, dev_out_stride, d_len, nullptr, 0, d_flag, st);
    };
    return run_host_batch(rgb_frames, frame_stride, 3 * npix, n, out, out_stride, imcvtHEVCBound(height, width), out_len,
                          nullptr, 0, 0, gray_stride, enc);
}

int imcvtEncodeBatchHEVCSweep(const uint8_t *frames, size_t frame_stride, int n,
                              uint32_t height, uint32_t width, const int *qpd6, int nq,
                              uint8_t *out, size_t out_stride, size_t *out_len)
{
    IMCVT_API_LOCK();
    if (!frames || !out || !out_len || !qpd6) { set_error("bad arguments"); return 1; }
    if (nq < 1 || nq > 8) { set_error("a sweep takes 1..8 settings (got %d)", nq); return 1; }
    if (!batch_ok(n) || !hevc_dims_ok(height, width)) return 1;
    for (int k = 0; k < nq; k++)
        if (qpd6[k] < 0 || qpd6[k] > 4) { set_error("qpd6 must be 0..4 (got %d)", qpd6[k]); return 1; }
    if ((long long)n * nq > 0x7FFFFFFF) { set_error("too many jobs"); return 1; }
    const size_t frame_bytes = (size_t)height * width;
    if (frame_stride < frame_bytes) { set_error("frame_stride smaller than one frame"); return 1; }
    if (ensure_engine()) return 1;
    Engine &E = engine();
    cudaStream_t st = E.stream;
    const int jobs = n * nq;
    const size_t bound = imcvtHEVCBound(height, width);
    const size_t dev_in_stride  = (frame_bytes + 255) & ~(size_t)255;
    const size_t dev_out_stride = ((out_stride < bound ? out_stride : bound) + 255) & ~(size_t)255;
    if (E.io.reserve((size_t)n * dev_in_stride + (size_t)jobs * dev_out_stride + (size_t)jobs * 8 + 4096)) return 1;
    uint8_t  *d_in   = E.io.take<uint8_t>((size_t)n * dev_in_stride);
    uint8_t  *d_out  = E.io.take<uint8_t>((size_t)jobs * dev_out_stride);
    uint64_t *d_len  = E.io.take<uint64_t>((size_t)jobs);
    int      *d_flag = E.io.take<int>(1);
    if (!d_in || !d_out || !d_len || !d_flag) { set_error("io arena carve failed"); return 1; }
    IMCVT_CUDA(cudaMemcpy2DAsync(d_in, dev_in_stride, frames, frame_stride, frame_bytes, n, cudaMemcpyHostToDevice, st));
    IMCVT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    if (hevc::encode_device(d_in, dev_in_stride, n, height, width, qpd6[0], d_out, dev_out_stride, d_len, nullptr, 0, d_flag, st, qpd6, nq)) return 1;
    std::vector<uint64_t> lens(jobs);
    int flag = 0;
    IMCVT_CUDA(cudaMemcpyAsync(lens.data(), d_len, (size_t)jobs * 8, cudaMemcpyDeviceToHost, st));
    IMCVT_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    IMCVT_CUDA(cudaStreamSynchronize(st));
    if (flag) { set_error("output slot too small for at least one image (out_stride = %zu)", out_stride); return 1; }
    for (int j = 0; j < jobs; j++) {
        if (lens[j] == 0 || lens[j] > out_stride) { set_error("job %d: encoder produced an invalid length %llu", j, (unsigned long long)lens[j]); return 1; }
        out_len[j] = (size_t)lens[j];
        IMCVT_CUDA(cudaMemcpyAsync(out + (size_t)j * out_stride, d_out + (size_t)j * dev_out_stride, (size_t)lens[j], cudaMemcpyDeviceToHost, st));
    }
    IMCVT_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ---- multi-device entry points -------------------------------------------------------------------------------------
// Frames are independent units (JLSencodeScan / HEVCImageEncoder re-initialise all state per call, imageio_jls.c:256-267,
// HEVCe.c:1577-1599), so a batch shards over the devices of the box in contiguous blocks of ceil(n / ndev) frames -- the
// multi-file loop of the reference's main (main.c:162-211) spread over GPUs.  One host thread per device drives that
// device's engine through the single-device entry point on its block; results land in the caller's slots directly, the
// sizes in the caller's out_len -- one process, no collective.
}  // extern "C"

#include <thread>

static std::mutex g_multi_mu;

static int resolve_devices(const int *devices, int ndev, std::vector<int> &list)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available: libimcvt_b200 has no CPU fallback");
        return 1;
    }
    if (!devices || ndev <= 0) {
        for (int d = 0; d < count; d++) list.push_back(d);
        return 0;
    }
    for (int k = 0; k < ndev; k++) {
        if (devices[k] < 0 || devices[k] >= count) { set_error("device %d does not exist (%d visible)", devices[k], count); return 1; }
        for (int j = 0; j < k; j++)
            if (devices[j] == devices[k]) { set_error("device %d is listed twice", devices[k]); return 1; }
        list.push_back(devices[k]);
    }
    return 0;
}

// call(first, count) runs the single-device entry point on frames [first, first + count) with the worker's engine selected
template <class Call>
static int run_multi(int n, const int *devices, int ndev, Call call)
{
    std::lock_guard<std::mutex> serial(g_multi_mu);
    std::vector<int> list;
    if (resolve_devices(devices, ndev, list)) return 1;
    const int G = (int)list.size();
    const int per = (n + G - 1) / G;                                  // frame f -> block f / per (imcvt_b200/shard.py: frame_block)
    std::vector<std::thread> workers;
    std::vector<int> rc(G, 0);
    std::vector<std::string> err(G);
    for (int g = 0; g < G; g++) {
        const int first = g * per, count = n - first < per ? n - first : per;
        if (count <= 0) break;
        workers.emplace_back([&, g, first, count]() {
            Engine *E = device_engine(list[g]);
            if (!E) { rc[g] = 1; err[g] = "no engine for the device"; return; }
            select_engine(E);
            rc[g] = call(first, count);
            if (rc[g]) err[g] = imcvtLastError();
            select_engine(nullptr);
        });
    }
    for (auto &w : workers) w.join();
    for (int g = 0; g < G; g++)
        if (rc[g]) { set_error("device %d: %s", list[g], err[g].c_str()); return 1; }
    return 0;
}

extern "C" {

int imcvtDeviceCount(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return count;
}

int imcvtEncodeBatchJLSMulti(const uint8_t *frames, size_t frame_stride, int n, int is_rgb,
                             uint32_t height, uint32_t width, int near_,
                             uint8_t *out, size_t out_stride, size_t *out_len, const int *devices, int ndev)
{
    if (!frames || !out || !out_len) { set_error("bad arguments"); return 1; }
    if (!batch_ok(n)) return 1;
    return run_multi(n, devices, ndev, [&](int first, int count) {
        return imcvtEncodeBatchJLS(frames + (size_t)first * frame_stride, frame_stride, count, is_rgb, height, width, near_,
                                   out + (size_t)first * out_stride, out_stride, out_len + first);
    });
}

int imcvtEncodeBatchHEVCMulti(const uint8_t *frames, size_t frame_stride, int n,
                              uint32_t height, uint32_t width, int qpd6,
                              uint8_t *out, size_t out_stride, size_t *out_len,
                              uint8_t *rcon, size_t rcon_stride, const int *devices, int ndev)
{
    if (!frames || !out || !out_len) { set_error("bad arguments"); return 1; }
    if (!batch_ok(n)) return 1;
    return run_multi(n, devices, ndev, [&](int first, int count) {
        return imcvtEncodeBatchHEVC(frames + (size_t)first * frame_stride, frame_stride, count, height, width, qpd6,
                                    out + (size_t)first * out_stride, out_stride, out_len + first,
                                    rcon ? rcon + (size_t)first * rcon_stride : nullptr, rcon_stride);
    });
}

int imcvtEncodeBatchHEVCGreenMulti(const uint8_t *rgb_frames, size_t frame_stride, int n,
                                   uint32_t height, uint32_t width, int qpd6,
                                   uint8_t *out, size_t out_stride, size_t *out_len, const int *devices, int ndev)
{
    if (!rgb_frames || !out || !out_len) { set_error("bad arguments"); return 1; }
    if (!batch_ok(n)) return 1;
    return run_multi(n, devices, ndev, [&](int first, int count) {
        return imcvtEncodeBatchHEVCGreen(rgb_frames + (size_t)first * frame_stride, frame_stride, count, height, width, qpd6,
                                         out + (size_t)first * out_stride, out_stride, out_len + first);
    });
}

int imcvtEncodeBatchHEVCSweepMulti(const uint8_t *frames, size_t frame_stride, int n,
                                   uint32_t height, uint32_t width, const int *qpd6, int nq,
                                   uint8_t *out, size_t out_stride, size_t *out_len, const int *devices, int ndev)
{
    if (!frames || !out || !out_len || !qpd6) { set_error("bad arguments"); return 1; }
    if (nq < 1 || nq > 8) { set_error("a sweep takes 1..8 settings (got %d)", nq); return 1; }
    if (!batch_ok(n)) return 1;
    return run_multi(n, devices, ndev, [&](int first, int count) {
        return imcvtEncodeBatchHEVCSweep(frames + (size_t)first * frame_stride, frame_stride, count, height, width, qpd6, nq,
                                         out + (size_t)first * nq * out_stride, out_stride, out_len + (size_t)first * nq);
    });
}

void *imcvtHostAlloc(size_t bytes)
{
    void *p = nullptr;
    IMCVT_API_LOCK();
    if (ensure_engine()) return nullptr;
    // portable: page-locked for every device's context, so the multi-device entry points copy from / to it at full rate
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("pinned host allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}

void imcvtHostFree(void *p)
{
    if (p) cudaFreeHost(p);
}

}  // extern "C"
