// placeholder until the tcgen05 kernel lands: reports "unsupported" so every conv takes the fp32 path
#include "tc.cuh"

namespace ds {
size_t tc_packed_weight_bytes(int, int, int) { return 0; }
int tc_pack_conv_weight(const float*, uint8_t*, int, int, int, cudaStream_t) { return DS_OK; }
bool tc_conv_supported(const ConvSrc&, int, int, int, const ConvEpi&) { return false; }
int tc_launch_conv(const ConvSrc&, const uint8_t*, int, int, int, int, int, int, const ConvEpi&, float*, cudaStream_t) {
    set_error("tensor-core conv not built");
    return DS_ERR_INVALID;
}
}  // namespace ds
