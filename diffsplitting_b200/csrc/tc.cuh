// bf16 tensor-core (tcgen05 / TMEM) convolution path, precision mode DS_PREC_BF16.
#pragma once
#include "common.cuh"

namespace ds {

// bytes of the UMMA-ready bf16 pack of one OIHW conv weight
size_t tc_packed_weight_bytes(int cout, int cin, int ks);
int tc_pack_conv_weight(const float* w_oihw, uint8_t* packed, int cout, int cin, int ks, cudaStream_t st);
bool tc_conv_supported(const ConvSrc& src, int cout, int ks, int stride, const ConvEpi& epi);
int tc_launch_conv(const ConvSrc& src, const uint8_t* w_packed, int cout, int ks, int stride, int B, int Ho, int Wo,
                   const ConvEpi& epi, float* out, cudaStream_t st);

}  // namespace ds
