// bf16 tensor-core (tcgen05 / TMEM / TMA) convolution path, precision mode DS_PREC_BF16.  See tc.cu.
#pragma once
#include "common.cuh"

namespace ds {

struct TcConvPlan {
    alignas(64) uint8_t params[2048];     // a TcParams (tensor maps + geometry), filled by tc_build_conv
    int smem_bytes, grid_x, grid_y, grid_z, kc;
    int patch;                            // 1: params hold a TcpParams (tall-patch kernel)
};

// tf32 = 1: operands are fp32 tensors / fp32 weight images read by the tensor core as TF32 (kind::tf32), else bf16
int tc_pick_kc(int ca, int cb, int tf32 = 0);                      // channels per K chunk (rows of 128 / 64 / 32 bytes), 0 if unsupported
bool tc_conv_shape_supported(int ca, int cb, int ks, int stride, int up, int Hs, int Ws, int tf32 = 0);
size_t tc_packed_weight_bytes(int cout, int cin, int ks, int tf32 = 0);   // upper bound over packing variants
// cin_src (optional): channels present in w_oihw; the packed image has cin >= cin_src channels, the extra ones zero
int tc_pack_conv_weight(const float* w_oihw, uint8_t* packed, int cout, int cin, int ks, int up, int kc, int tf32, cudaStream_t st,
                        int cin_src = 0);
// the network input (one or two fp32 NCHW tensors, <= 8 channels together) -> [B,H,W,cpad] fp32 NHWC (cpad 8 or 16), zero padded,
// rounded to TF32
int tc_pack_input(const float* xa, int ca, const float* xb, int cb, float* out, int cpad, int B, int H, int W, cudaStream_t st);
// geometry + tensor maps for: out = conv(cat[src_a, src_b]) ; sources bf16 (tf32: fp32) NHWC [B,Hs,Ws,c]
// xsrc_a / xsrc_b (optional): a second input whose 1x1 conv is accumulated into the same output (the res_conv of a ResNet block);
// only with the persistent kernel: ask tc_conv_persistent first
int tc_build_conv(TcConvPlan* plan, const void* src_a, int ca, const void* src_b, int cb, int Hs, int Ws, int B, int cout, int ks,
                  int stride, int up, int tf32 = 0, const void* xsrc_a = nullptr, int xca = 0, const void* xsrc_b = nullptr, int xcb = 0);
bool tc_conv_persistent(int ca, int cb, int Hs, int Ws, int B, int cout, int ks, int stride, int up, int tf32);
// epi.bias / temb / residual (fp32 NHWC) as in the fp32 kernel; any subset of the three outputs may be given
// sums_out (optional): [B][cout][2] fp64 accumulators (zeroed by the caller) receiving per-channel sum / sum of squares
// xw_packed / bias2: the folded 1x1 conv's weights (packed with ks = 1 and the same K chunk) and bias
int tc_launch_conv(const TcConvPlan* plan, const uint8_t* w_packed, const ConvEpi& epi, float* out_f32, void* out_b16,
                   float* out_nchw, double* sums_out, cudaStream_t st, const uint8_t* xw_packed = nullptr, const float* bias2 = nullptr);

// ---- fused GroupNorm-apply + Swish -> conv, operands staged through registers (tc_halo.cu); sources fp32 NHWC
bool halo_conv_supported(int ca, int cb, int cout, int ks, int B, int H, int W, int tf32 = 0);
bool halo_conv_preferred(int ca, int cb, int cout, int ks, int B, int H, int W, int tf32 = 0);   // supported AND faster than GN-apply + TMA conv
size_t halo_packed_weight_bytes(int cout, int cin, int ks, int tf32 = 0);
int halo_pack_conv_weight(const float* w_oihw, uint8_t* packed, int cout, int cin, int ks, int tf32, cudaStream_t st);
// GroupNorm statistics of the concat input: either `stats` = (mean, rstd) [B][G], or per-channel fp64 (sum, sumsq)
// accumulators of the two sources `sums_a` [B][ca][2] / `sums_b` [B][cb][2] (as emitted by the producers' epilogues);
// all null: no normalisation.
struct HaloNorm {
    const float2* stats;
    const double* sums_a; const double* sums_b;
    const float* gamma; const float* beta;
    int G, swish;
};
int halo_launch_conv(const float* src_a, int ca, const float* src_b, int cb, const HaloNorm& norm, const uint8_t* w_packed,
                     int cout, int ks, int B, int H, int W, const ConvEpi& epi, float* out_f32, void* out_b16, float* out_nchw,
                     double* sums_out, int tf32, cudaStream_t st);

// ---- persistent, software-pipelined variant of the fused kernel for many-tile 3x3 layers whose weights fit in shared memory
// (tc_stream.cu); same weight pack and arguments as halo_launch_conv, which dispatches to it
bool stream_conv_preferred(int ca, int cb, int cout, int ks, int B, int H, int W, int tf32);
int stream_launch_conv(const float* src_a, int ca, const float* src_b, int cb, const HaloNorm& norm, const uint8_t* w_packed,
                       int cout, int ks, int B, int H, int W, const ConvEpi& epi, float* out_f32, void* out_b16, float* out_nchw,
                       double* sums_out, int tf32, cudaStream_t st);

// ---- bf16 tensor-core self-attention (attention_tc.cu): qkv bf16 [B,N,3C] -> out bf16 [B,N,C]
struct AttnTcPlan {
    alignas(64) uint8_t params[512];      // an AttnTcParams (tensor map + geometry), filled by attn_tc_build
    int smem_bytes, grid_x, grid_y, grid_z;
};
bool attn_tc_supported(int N, int C);                              // C a multiple of 64
int attn_tc_build(AttnTcPlan* plan, const void* qkv, void* out, int B, int N, int C);
int attn_tc_launch(const AttnTcPlan* plan, cudaStream_t st);

// ---- per-sample persistent chain of [GroupNorm(+Swish) ->] 3x3 / 1x1 convs for the low-resolution levels (tc_chain.cu)
constexpr int CHAIN_MAX_OPS = 16;
constexpr int CHAIN_MAX_MTILES = 3;           // 128-row tiles per sample: (H+2)(W+2) <= 384, i.e. up to 17x17
struct ChainOpDesc {
    const void* src_a; const void* src_b;     // fp32 NHWC sources (concat), or bf16 NHWC when src_b16 (then no normalisation)
    int ca, cb, src_b16;
    int norm, swish, G;                       // fused GroupNorm of the concat input, statistics from sums_a / sums_b
    const double* sums_a; const double* sums_b;
    const float* gamma; const float* beta;
    const uint8_t* w;                         // chain_pack_conv_weight() image
    int cout, ks;
    // optional second operand, raw (no normalisation), applied as a 1x1 conv and accumulated into the same output:
    // the res_conv of a ResnetBlock folded into its conv2 (unet.py:118-122:  block2(h) + res_conv(x))
    const float* xsrc_a; const float* xsrc_b; int xca, xcb;
    const uint8_t* wx;                        // chain_pack_conv_weight() image of the 1x1 weights
    const float* bias2;
    ConvEpi epi;                              // bias / temb / residual (fp32 NHWC)
    float* out_f32; void* out_b16;            // fp32 and/or bf16 NHWC outputs
    double* sums_out;                         // statistics of the output (replicated slots) or null
};
struct ChainPlan {
    alignas(64) uint8_t params[4096];         // a ChainParams
    int smem_bytes, B, CL;                    // CL = CTAs per cluster = per sample
};
bool chain_level_supported(int H, int W);
int chain_slice_rows(int cout);               // output channels per CTA (16 or 32)
int chain_cluster_size(int cout);             // CTAs per sample; all ops of one chain must agree on both
bool chain_conv_supported(int ca, int cb, int cout, int ks, int H, int W);
size_t chain_packed_weight_bytes(int cout, int cin, int ks);
int chain_pack_conv_weight(const float* w_oihw, uint8_t* packed, int cout, int cin, int ks, cudaStream_t st);
int chain_build(ChainPlan* plan, const ChainOpDesc* ops, int nops, int B, int H, int W);
int chain_launch(const ChainPlan* plan, cudaStream_t st);

}  // namespace ds
