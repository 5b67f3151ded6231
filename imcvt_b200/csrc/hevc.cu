// hevc.cu -- H.265 intra encoder for sm_100a (placeholder until the frame kernel lands)
#include "common.cuh"
namespace imcvt { namespace hevc {
int encode_device(const uint8_t *, size_t, int, uint32_t, uint32_t, int, uint8_t *, size_t, uint64_t *, uint8_t *, size_t, int *, cudaStream_t)
{
    set_error("HEVC device path not built yet");
    return 1;
}
}}
