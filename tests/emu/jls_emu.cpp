// tests/emu/jls_emu.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the SOURCE of the CUDA JPEG-LS tile kernel (imcvt_b200/csrc/jls_tile_kernel.cuh) under the fiber SIMT
// emulator of cuda_emu.h; the markers around the scan are written here the way assemble_kernel does on the device.
#include "cuda_emu.h"
#include "../../imcvt_b200/csrc/jls_tile_kernel.cuh"

using namespace imcvt::jls;

// gray8 image H x W -> complete .jls stream (NEAR = 0); returns its length, 0 when it did not fit
extern "C" size_t emu_jls_encode_mode(const uint8_t *img, int H, int W, uint8_t *out, size_t cap, int pair);
extern "C" size_t emu_jls_encode(const uint8_t *img, int H, int W, uint8_t *out, size_t cap) { return emu_jls_encode_mode(img, H, W, out, cap, 0); }
// pair = 1: the two-CTA cluster kernel, emulated as one block of 2 x NTHREADS fibers with two shared-memory images
extern "C" size_t emu_jls_encode_mode(const uint8_t *img, int H, int W, uint8_t *out, size_t cap, int pair)
{
    if (W > tile::WMAX || cap < 64) return 0;
    tile::Params P;
    unsigned long long bytes = 0;
    int next = 0;
    P.pix = img; P.plane_stride = (size_t)H * W; P.P = 1; P.W = W; P.H = H;
    P.L.ncomp = 1; P.L.hdr_len = 15; P.L.out = out; P.L.out_stride = cap - 2; P.L.scratch = nullptr; P.L.scratch_stride = 0;
    P.plane_bytes = &bytes; P.next_plane = &next;
    P.prof = nullptr;
    P.tma = (W % 32 == 0) ? 1 : 0;      // both staging paths get exercised by the test shapes
    P.line_out = (pair >> 1) & 1;       // bit 1: stuffed bytes leave through the 128-byte line ring (output in host memory)
    pair &= 1;
    if (pair) emu::launch(1, 2 * tile::NTHREADS, 2 * tile::SMEM_STRIDE, [&] { tile::jls_pair_kernel(P); });
    else emu::launch(1, tile::NTHREADS, sizeof(tile::Smem), [&] { tile::jls_tile_kernel(P); });
    const size_t total = 25 + (size_t)bytes + 2;
    if (total > cap) return 0;
    const uint8_t hdr[25] = {0xFF, 0xD8, 0xFF, 0xF7, 0x00, 0x0B, 8, (uint8_t)(H >> 8), (uint8_t)H, (uint8_t)(W >> 8), (uint8_t)W, 1, 1, 0x11, 0,
                             0xFF, 0xDA, 0x00, 0x08, 0x01, 0x01, 0x00, 0x00, 0x00, 0x00};
    memcpy(out, hdr, 25);
    out[total - 2] = 0xFF; out[total - 1] = 0xD9;
    return total;
}
extern "C" long emu_jls_smem_bytes(void) { return (long)sizeof(tile::Smem); }
