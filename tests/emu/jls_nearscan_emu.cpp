// tests/emu/jls_nearscan_emu.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the SOURCE of the candidate-table JPEG-LS scan kernel (imcvt_b200/csrc/jls_nearscan_kernel.cuh: the path for NEAR 1..4)
// under the fiber SIMT emulator of cuda_emu.h; the markers around the scan are written here the way assemble_kernel does.
#include "cuda_emu.h"
#include "../../imcvt_b200/csrc/jls_nearscan_kernel.cuh"

using namespace imcvt::jls;

// gray8 image H x W -> complete .jls stream at NEAR = near_ (1..4); returns its length, 0 when it did not fit
extern "C" size_t emu_jls_nearscan_encode(const uint8_t *img, int H, int W, int near_, uint8_t *out, size_t cap)
{
    if (W > nearscan::WMAXR || cap < 64 || near_ < 1 || near_ > 4) return 0;
    FrameLayout L;
    L.ncomp = 1; L.hdr_len = 15; L.out = out; L.out_stride = cap - 2; L.scratch = nullptr; L.scratch_stride = 0;
    const SerialParams SP = make_params(near_);
    uint64_t bytes = 0;
    const size_t plane = (size_t)H * W;
    switch (near_) {
    case 1: emu::launch(1, nearscan::NT, sizeof(nearscan::Smem), [&] { nearscan::nearscan_kernel<1>(img, plane, W, H, SP, L, &bytes); }); break;
    case 2: emu::launch(1, nearscan::NT, sizeof(nearscan::Smem), [&] { nearscan::nearscan_kernel<2>(img, plane, W, H, SP, L, &bytes); }); break;
    case 3: emu::launch(1, nearscan::NT, sizeof(nearscan::Smem), [&] { nearscan::nearscan_kernel<3>(img, plane, W, H, SP, L, &bytes); }); break;
    default: emu::launch(1, nearscan::NT, sizeof(nearscan::Smem), [&] { nearscan::nearscan_kernel<4>(img, plane, W, H, SP, L, &bytes); }); break;
    }
    const size_t total = 25 + (size_t)bytes + 2;
    if (total > cap) return 0;
    const uint8_t hdr[25] = {0xFF, 0xD8, 0xFF, 0xF7, 0x00, 0x0B, 8, (uint8_t)(H >> 8), (uint8_t)H, (uint8_t)(W >> 8), (uint8_t)W, 1, 1, 0x11, 0,
                             0xFF, 0xDA, 0x00, 0x08, 0x01, 0x01, 0x00, (uint8_t)near_, 0x00, 0x00};
    memcpy(out, hdr, 25);
    out[total - 2] = 0xFF; out[total - 1] = 0xD9;
    return total;
}
