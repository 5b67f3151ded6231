// tests/emu/jls_rowscan_emu.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the SOURCE of the row-pipelined JPEG-LS scan kernel (imcvt_b200/csrc/jls_rowscan_kernel.cuh: the path for NEAR > 0,
// three warps a row segment apart) under the fiber SIMT emulator of cuda_emu.h; the markers around the scan are written here
// the way assemble_kernel does on the device.
#include "cuda_emu.h"
#include "../../imcvt_b200/csrc/jls_rowscan_kernel.cuh"

using namespace imcvt::jls;

// gray8 image H x W -> complete .jls stream at NEAR = near_; returns its length, 0 when it did not fit
extern "C" size_t emu_jls_rowscan_encode(const uint8_t *img, int H, int W, int near_, uint8_t *out, size_t cap)
{
    if (W > rowscan::WMAXR || cap < 64) return 0;
    FrameLayout L;
    L.ncomp = 1; L.hdr_len = 15; L.out = out; L.out_stride = cap - 2; L.scratch = nullptr; L.scratch_stride = 0;
    const SerialParams SP = make_params(near_);
    uint64_t bytes = 0;
    if (near_ > 0)
        emu::launch(1, rowscan::NT, sizeof(rowscan::Smem), [&] { rowscan::rowscan_kernel<true>(img, (size_t)H * W, W, H, SP, L, &bytes); });
    else
        emu::launch(1, rowscan::NT, sizeof(rowscan::Smem), [&] { rowscan::rowscan_kernel<false>(img, (size_t)H * W, W, H, SP, L, &bytes); });
    const size_t total = 25 + (size_t)bytes + 2;
    if (total > cap) return 0;
    const uint8_t hdr[25] = {0xFF, 0xD8, 0xFF, 0xF7, 0x00, 0x0B, 8, (uint8_t)(H >> 8), (uint8_t)H, (uint8_t)(W >> 8), (uint8_t)W, 1, 1, 0x11, 0,
                             0xFF, 0xDA, 0x00, 0x08, 0x01, 0x01, 0x00, (uint8_t)near_, 0x00, 0x00};
    memcpy(out, hdr, 25);
    out[total - 2] = 0xFF; out[total - 1] = 0xD9;
    return total;
}
