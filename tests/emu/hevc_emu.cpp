// tests/emu/hevc_emu.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the SOURCE of the CUDA HEVC frame kernel (imcvt_b200/csrc/hevc_kernel.cuh) under the fiber SIMT emulator
// of cuda_emu.h, so that tests/test_emu_hevc.py can compare it bit-for-bit with the oracle without a GPU.
// It is not linked into libimcvt_b200.so and nothing in the product loads it.
#include "cuda_emu.h"
#include "../../imcvt_b200/csrc/hevc_kernel.cuh"

using namespace imcvt::hevc;

extern "C" long emu_hevc_encode_mode(uint8_t *stream, size_t cap, const uint8_t *img, uint8_t *img_rcon, int *ysz, int *xsz, int qpd6, int pair);
extern "C" long emu_hevc_encode(uint8_t *stream, size_t cap, const uint8_t *img, uint8_t *img_rcon, int *ysz, int *xsz, int qpd6)
{
    return emu_hevc_encode_mode(stream, cap, img, img_rcon, ysz, xsz, qpd6, 0);
}
// pair bit 0: the two-CTA cluster kernel (spine + upper), both CTAs' threads as fibers of one emulated cluster;
// bit 1: the 8x8 units searched in lock step (Params::cu8_lockstep) instead of by three warp groups
extern "C" long emu_hevc_encode_mode(uint8_t *stream, size_t cap, const uint8_t *img, uint8_t *img_rcon, int *ysz, int *xsz, int qpd6, int pair)
{
    const int lockstep = (pair >> 1) & 1;
    pair &= 1;
    const int H = *ysz, W = *xsz;
    const int hp = (std::min(H, 8192) + 31) / 32 * 32, wp = (std::min(W, 8192) + 31) / 32 * 32;
    Params P;
    unsigned long long len = 0;
    int overflow = 0, next = 0;
    std::vector<uint8_t> scratch((size_t)2 * 2 * NALL * SCR_BYTES);
    std::vector<unsigned short> lists((size_t)2 * 2 * NALL * CAND_LIST_U16, 0xDEAD);
    P.frames = img; P.frame_stride = (size_t)H * W; P.n = 1; P.H = H; P.W = W; P.hp = hp; P.wp = wp; P.qpd6 = qpd6;
    P.out = stream; P.out_stride = cap; P.out_len = &len;
    P.rcon = img_rcon; P.rcon_stride = (size_t)hp * wp;
    P.scratch = scratch.data(); P.lists = lists.data(); P.overflow = &overflow; P.next_frame = &next;
    P.prof = nullptr;
    P.nq = 0;
    P.cu8_lockstep = lockstep;
    for (int k = 0; k < 8; k++) P.qorder[k] = k;
    if (pair) emu::launch(1, NT, sizeof(Smem), [&] { hevc_pair_kernel(P); }, 2);
    else emu::launch(1, NT, sizeof(Smem), [&] { hevc_frame_kernel(P); });
    *ysz = hp; *xsz = wp;
    return overflow ? 0 : (long)len;
}

// a quantiser sweep on ONE emulated CTA (or pair): job k codes the frame at qlist[k] into slot k (stream slot `cap` bytes apart,
// reconstructions hp*wp apart); lens[k] receives the stream lengths.  Returns 0 on success.
extern "C" int emu_hevc_sweep(uint8_t *streams, size_t cap, const uint8_t *img, uint8_t *rcons, int H, int W, const int *qlist, int nq,
                              unsigned long long *lens, int pair)
{
    const int hp = (std::min(H, 8192) + 31) / 32 * 32, wp = (std::min(W, 8192) + 31) / 32 * 32;
    Params P;
    int overflow = 0, next = 0;
    std::vector<uint8_t> scratch((size_t)2 * 2 * NALL * SCR_BYTES);
    std::vector<unsigned short> lists((size_t)2 * 2 * NALL * CAND_LIST_U16, 0xDEAD);
    P.frames = img; P.frame_stride = (size_t)H * W; P.n = 1; P.H = H; P.W = W; P.hp = hp; P.wp = wp; P.qpd6 = qlist[0];
    P.out = streams; P.out_stride = cap; P.out_len = lens;
    P.rcon = rcons; P.rcon_stride = (size_t)hp * wp;
    P.scratch = scratch.data(); P.lists = lists.data(); P.overflow = &overflow; P.next_frame = &next;
    P.prof = nullptr;
    P.nq = nq;
    P.cu8_lockstep = 1;                             // the product's default on both launch shapes
    for (int k = 0; k < 8; k++) P.qlist[k] = k < nq ? qlist[k] : 0;
    // the queue order of the product's host side (hevc.cu): settings by ascending qpd6, ties in list order
    for (int k = 0; k < 8; k++) P.qorder[k] = k;
    for (int a = 1; a < nq; a++)
        for (int b = a; b > 0 && qlist[P.qorder[b]] < qlist[P.qorder[b - 1]]; b--) { const int t = P.qorder[b]; P.qorder[b] = P.qorder[b - 1]; P.qorder[b - 1] = t; }
    if (pair) emu::launch(1, NT, sizeof(Smem), [&] { hevc_pair_kernel(P); }, 2);
    else emu::launch(1, NT, sizeof(Smem), [&] { hevc_frame_kernel(P); });
    return overflow;
}

extern "C" long emu_smem_bytes(void) { return (long)sizeof(Smem); }
extern "C" long emu_barrier_count(int which) { return which ? emu::S().n_warp_barriers : emu::S().n_block_barriers; }
