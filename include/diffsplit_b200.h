/*
 * diffsplit_b200.h - C ABI of the B200-native DiffSplitting sampling hot path.
 *
 * The reference (rayanirban/DiffSplitting) has no FFI of its own: its boundary is the
 * duck-typed Python object graph  model.create_model(opt) -> DDPM -> netG -> denoise_fn
 * (model/__init__.py:5-9, model/model.py:12-43, model/networks.py:91-180).  The Python
 * mirror in diffsplitting_b200/model/ keeps that surface and reaches the device ONLY
 * through the entry points below (ctypes; see INTEGRATION.md for the binding stub).
 *
 * Conventions
 *  - every function returns DS_OK (0) or a negative ds_status; ds_last_error() gives the text;
 *    nothing throws across the boundary
 *  - all pointers named d_* are DEVICE pointers owned by the caller (torch tensors);
 *    all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*), with no
 *    internal synchronisation and no allocation after ds_unet_load_weights -> every call is
 *    CUDA-graph capturable
 *  - external image tensors are fp32 NCHW contiguous, exactly what the reference passes
 *  - handles are not thread-safe (the reference is single-threaded per process)
 */
#ifndef DIFFSPLIT_B200_H
#define DIFFSPLIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ds_status {
    DS_OK = 0,
    DS_ERR_INVALID = -1,        /* bad argument / unsupported shape */
    DS_ERR_CUDA = -2,           /* a CUDA runtime call failed */
    DS_ERR_WORKSPACE = -3,      /* caller workspace too small */
    DS_ERR_WEIGHT = -4,         /* missing / mis-shaped weight */
    DS_ERR_NO_DEVICE = -5       /* no sm_100 device visible */
} ds_status;

enum { DS_UNET_SR3 = 0, DS_UNET_DDPM = 1 };       /* sr3_modules/unet.py vs ddpm_modules/unet.py */
enum { DS_PREC_FP32 = 0, DS_PREC_BF16 = 1, DS_PREC_TF32 = 2 };
/* fp32: CUDA-core path (<= 1e-5 gate).  bf16: tcgen05 path, bf16 MMA operands.  tf32: tcgen05 path with fp32 tensors read as
 * TF32 operands (kind::tf32, 10-bit mantissa: what the reference's own cuDNN convolutions use on a GPU by default) - for the
 * narrow latency-bound splitting nets, where the halved MMA rate is free; needs ds_unet_desc.tf32_weights. */
enum { DS_TILE_TRIM = 0, DS_TILE_PAD = 1, DS_TILE_SHIFT = 2 };   /* data/tiling_manager.py:6-12 */

const char* ds_last_error(void);
int ds_version(void);
/* SM count / max threads per SM of the current device (needed to replay torch's Philox launch geometry). */
int ds_device_info(int* sm_count, int* max_threads_per_sm, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ UNet (replaces UNet.forward,
 * model/sr3_modules/unet.py:161-259 and model/ddpm_modules/unet.py:147-243) */
#define DS_MAX_LEVELS 8
typedef struct ds_unet_desc {
    int32_t variant;                        /* DS_UNET_SR3 | DS_UNET_DDPM */
    int32_t in_channel, out_channel, inner_channel, norm_groups;
    int32_t n_mults;  int32_t channel_mults[DS_MAX_LEVELS];
    int32_t n_attn_res; int32_t attn_res[DS_MAX_LEVELS];
    int32_t res_blocks;
    int32_t image_size;                     /* seeds the attn_res matching, as in the reference ctor */
    int32_t with_time_emb;                  /* 0: time=None (TimePredictor style) */
    int32_t tf32_weights;                   /* 1: also keep TF32 weight images (required for DS_PREC_TF32 forwards) */
} ds_unet_desc;

typedef struct ds_tensor_view {
    const char* name;                       /* reference state_dict key relative to the UNet, e.g. "downs.1.res_block.block1.block.3.weight" */
    const void* d_data;                     /* device pointer, fp32 contiguous */
    int32_t ndim;
    int64_t shape[4];
} ds_tensor_view;

typedef struct ds_unet ds_unet;

int  ds_unet_create(const ds_unet_desc* desc, ds_unet** out);
void ds_unet_destroy(ds_unet* net);
/* the state_dict keys (and shapes) this architecture expects, in reference order */
int  ds_unet_num_weights(const ds_unet* net);
const char* ds_unet_weight_name(const ds_unet* net, int i);
int  ds_unet_weight_shape(const ds_unet* net, int i, int32_t* ndim, int64_t shape[4]);
/* copy + repack reference-layout fp32 weights into library-owned device buffers */
int  ds_unet_load_weights(ds_unet* net, const ds_tensor_view* weights, int n, void* stream);
size_t ds_unet_workspace_bytes(ds_unet* net, int B, int H, int W, int precision);
/* out[B,Cout,H,W] = UNet(cat([x_a, x_b], dim=1), time).  x_b may be NULL (cb = 0); the channel
 * concat of p_mean_variance (sr3 diffusion.py:157-158) is never materialised.
 * d_time: sr3 noise level (B,1) / ddpm time (B,) or (1,) -> time_len in {1, B}; NULL iff !with_time_emb. */
int  ds_unet_forward(ds_unet* net, const float* d_xa, int ca, const float* d_xb, int cb,
                     const float* d_time, int time_len, float* d_out,
                     int B, int H, int W, int precision,
                     void* d_workspace, size_t workspace_bytes, void* stream);
/* algorithmic FLOPs (2*MAC, BASELINE.md section 3 convention) of one forward for one sample */
double ds_unet_flops(const ds_unet* net, int H, int W);
/* number of kernels one forward enqueues (for bench.py's gpu_launches) */
int  ds_unet_launches(ds_unet* net, int B, int H, int W, int precision);
/* measurement: the same forward with a CUDA-event pair around every operator (synchronises; not capturable).
 * kind: 0 conditioning MLP, 1 conv fp32 (CUDA cores), 2 group-norm(+swish) (2 launches), 3 attention, 4 conv bf16
 * (tcgen05, TMA-fed), 5 group-norm statistics only, 6 fused group-norm-apply + swish + conv bf16 (tcgen05, staged),
 * 7 the same inside a per-sample persistent chain launch (time reported on the chain's first operator, launches = 0 on the rest).
 * flops / bytes are the ALGORITHMIC figures of that operator for the whole batch. */
typedef struct ds_op_profile {
    int32_t kind, cin, cout, ksize, h, w, launches;
    float   ms;
    double  flops, bytes;
} ds_op_profile;
int  ds_unet_forward_profiled(ds_unet* net, const float* d_xa, int ca, const float* d_xb, int cb,
                              const float* d_time, int time_len, float* d_out,
                              int B, int H, int W, int precision,
                              void* d_workspace, size_t workspace_bytes, void* stream,
                              ds_op_profile* ops, int max_ops, int* n_ops);
/* debugging/parity: copy an internal activation (reference module path, e.g. "downs.1") of the LAST
 * forward out of the workspace as fp32 NCHW; returns DS_ERR_INVALID if the name is unknown */
int  ds_unet_read_tap(ds_unet* net, const char* name, float* d_out, size_t out_elems,
                      void* d_workspace, void* stream);

/* ------------------------------------------------------------------ standalone operators
 * (each is one of the kernels ds_unet_forward composes; exported so that parity tests can pin them
 * one by one).  Activations here are fp32 NHWC [B,H,W,C]. */
/* GroupNorm(+Swish) over the channel-concat of a (ca ch) and b (cb ch, may be NULL):
 * reference Block = GroupNorm -> Swish (unet.py:80-91); eps 1e-5. */
int ds_groupnorm_swish_f32(const float* d_a, int ca, const float* d_b, int cb, const float* d_gamma,
                           const float* d_beta, float* d_out, int B, int H, int W, int groups,
                           int apply_swish, void* d_scratch, size_t scratch_bytes, void* stream);
size_t ds_groupnorm_scratch_bytes(int B, int groups);
/* direct convolution, weights in reference OIHW layout (repacked on the fly into d_scratch) */
int ds_conv2d_f32(const float* d_x, const float* d_w_oihw, const float* d_bias, float* d_out,
                  int B, int H, int W, int cin, int cout, int ksize, int stride, int upsample2x,
                  void* d_scratch, size_t scratch_bytes, void* stream);
size_t ds_conv2d_scratch_bytes(int cin, int cout, int ksize);
/* bf16 tensor-core convolution (tcgen05 + TMEM + TMA), the operator DS_PREC_BF16 forwards are built from.
 * d_xa / d_xb: bf16 NHWC sources (channel concat; d_xb may be NULL), channel counts multiples of 16;
 * d_residual: fp32 NHWC [B,Ho,Wo,cout] or NULL (the residual stream stays fp32); d_out: bf16 NHWC, or fp32 NCHW if
 * out_f32_nchw; d_out_f32 (optional): an additional fp32 NHWC copy of the result.
 * stride 2 = Downsample (unet.py:68-74); upsample2x = nearest x2 + 3x3 (unet.py:58-65) computed as four 2x2
 * convolutions on the low-resolution input with pre-summed weights. */
int ds_conv2d_bf16(const void* d_xa, int ca, const void* d_xb, int cb, const float* d_w_oihw, const float* d_bias,
                   const float* d_residual, void* d_out, float* d_out_f32, int out_f32_nchw, int B, int H, int W, int cout,
                   int ksize, int stride, int upsample2x, void* d_scratch, size_t scratch_bytes, void* stream);
size_t ds_conv2d_bf16_scratch_bytes(int cin, int cout, int ksize);
/* fused Block (unet.py:80-91): GroupNorm(groups) -> Swish -> Conv(ksize 3 or 1, stride 1) over the channel concat of two
 * fp32 NHWC sources, as ONE tensor-core kernel after a statistics pass (groups == 0: plain conv of the raw input).
 * Total input channels: multiple of 16, <= 128.  Outputs: bf16 NHWC and / or fp32 NHWC. */
int ds_gnconv_bf16(const float* d_xa, int ca, const float* d_xb, int cb, const float* d_gamma, const float* d_beta,
                   int groups, int apply_swish, const float* d_w_oihw, const float* d_bias, const float* d_residual,
                   void* d_out_b16, float* d_out_f32, int B, int H, int W, int cout, int ksize, void* d_scratch,
                   size_t scratch_bytes, void* stream);
size_t ds_gnconv_bf16_scratch_bytes(int B, int groups, int cin, int cout, int ksize);
/* The same two operators with fp32 NHWC sources read by the tensor core as TF32 operands (precision mode DS_PREC_TF32:
 * kind::tf32, weights rounded to TF32 at pack time, fp32 accumulation); channel counts multiples of 8 (conv) / 16 (fused). */
int ds_conv2d_tf32(const float* d_xa, int ca, const float* d_xb, int cb, const float* d_w_oihw, const float* d_bias,
                   const float* d_residual, float* d_out_f32, int out_nchw, int B, int H, int W, int cout, int ksize,
                   int stride, int upsample2x, void* d_scratch, size_t scratch_bytes, void* stream);
size_t ds_conv2d_tf32_scratch_bytes(int cin, int cout, int ksize);
int ds_gnconv_tf32(const float* d_xa, int ca, const float* d_xb, int cb, const float* d_gamma, const float* d_beta,
                   int groups, int apply_swish, const float* d_w_oihw, const float* d_bias, const float* d_residual,
                   float* d_out_f32, int B, int H, int W, int cout, int ksize, void* d_scratch, size_t scratch_bytes,
                   void* stream);
size_t ds_gnconv_tf32_scratch_bytes(int B, int groups, int cin, int cout, int ksize);
/* debugging aid (DIFFSPLIT_B200_HALO_DBG=1): mean clock64 cycles of the 7 phases of the last fused-conv launch:
 * setup | wait for predecessor | loads + scale table | transform + stage | MMA | epilogue | teardown */
int ds_debug_stream_phases(long long* out16);   /* per-role wait / work cycles of the pipelined fused conv (tc_stream.cu) */
int ds_debug_halo_phases(double* h_out7, int* n_ctas);
/* debugging aid (DIFFSPLIT_B200_TRACE=1): GPU-timer (ns) start / end of every tensor-core conv launch, also inside
 * CUDA-graph replays.  reset(1) forgets the launch ids, reset(0) re-arms the recorded ones; read returns the count. */
int ds_debug_chain_phases(long long* h_out, int max_ops, int* n_ops);   /* DIFFSPLIT_B200_CHAIN_DBG=1, see tc_chain.cu */
int ds_debug_trace_reset(int forget_ids);
int ds_debug_trace_read(unsigned long long* h_start_end, int* h_kind, int max_n);
/* single-head attention over N=H*W tokens: qkv [B,N,3C] (q|k|v along C) -> out [B,N,C]
 * (softmax(q k^T / sqrt(C)) v, unet.py:132-139) */
int ds_attention_f32(const float* d_qkv, float* d_out, int B, int N, int C, void* stream);
/* the same on the tensor cores (tcgen05 / TMEM / TMA, precision mode DS_PREC_BF16): qkv and out are bf16, scores and
 * the output accumulate in fp32, the probabilities are rounded to bf16.  C must be a multiple of 64. */
int ds_attention_bf16(const void* d_qkv, void* d_out, int B, int N, int C, void* stream);
/* two chained fused ops in ONE per-sample persistent kernel (the low-resolution path of the bf16 UNet), shaped like a
 * ResnetBlock (unet.py:94-123):  h = conv1(swish(GN1(cat[xa, xb]))) + b1 ;  y = conv2(swish(GN2(h))) + b2 [+ residual].
 * fp32 NHWC in / out; (H+2)(W+2) <= 384; channel counts multiples of 8 (16 in total), <= 256. */
size_t ds_chain2_bf16_scratch_bytes(int B, int H, int W, int ca, int cb, int cmid, int cout, int ksize);
int ds_chain2_bf16(const float* d_xa, int ca, const float* d_xb, int cb, const float* d_gamma1, const float* d_beta1,
                   const float* d_w1_oihw, const float* d_b1, const float* d_gamma2, const float* d_beta2,
                   const float* d_w2_oihw, const float* d_b2, const float* d_residual, int groups, float* d_out_f32,
                   void* d_out_b16, int B, int H, int W, int cmid, int cout, int ksize, void* d_scratch, size_t scratch_bytes,
                   void* stream);

/* ------------------------------------------------------------------ sampler updates
 * One fused elementwise kernel per reverse step (replaces p_sample / inference_one_step:
 * sr3 diffusion.py:141-175, ddpm diffusion.py:179-203, indi.py:62-69).
 *
 *   mode 0 (posterior): x0 = c[0]*x - c[1]*net ; if clip: clamp(x0,-1,1)
 *   mode 1 (indi)     : x0 = net
 *   out = (c[2]*x0 + c[3]*x) + z * c[4]
 *
 * Every product/sum is individually rounded (no FMA contraction) in the reference's evaluation
 * order, so fp32 results are bit-identical to the eager reference for the same z.
 * Coefficients: row `k` of d_coef [n_steps,5]; k = d_state->step (device counter, CUDA-graph replay) when
 * d_state != NULL else `step`.  Noise z: d_noise (injected, same shape as x) if non-NULL; else generated
 * in-kernel with Philox4x32-10 replaying torch.randn's CUDA stream: seed, offset (d_state->offset or
 * `offset`), torch launch geometry `rng_threads` = 256*grid.  If c[4]==0 and skip_rng_if_zero the draw is
 * skipped and the offset is not advanced (sr3 t==0).  With d_state, the kernel's last block advances it
 * (step+1, offset += offset_inc) and, if d_time_out != NULL, writes the next step's UNet time input
 * d_time_table[k+1] to d_time_out[0..time_len).
 */
typedef struct ds_sampler_state {   /* lives in DEVICE memory, 24 bytes, initialise with a memcpy */
    uint64_t seed;                  /* Philox key = seed of the torch generator (read from here, not from the
                                       kernel arguments, so a captured graph follows torch.manual_seed) */
    uint64_t offset;                /* Philox offset of the torch generator */
    int32_t  step;                  /* index of the next step to run (0 = first executed step) */
    uint32_t done;                  /* scratch for last-block detection, must be 0 between launches */
} ds_sampler_state;

typedef struct ds_step_args {
    const float* d_x;          /* x_t          [numel] */
    const float* d_net;        /* UNet output  [numel] */
    float*       d_out;        /* x_{t-1}      [numel], may alias d_x */
    int64_t      numel;
    int32_t      mode;         /* 0 posterior, 1 indi */
    int32_t      clip;
    const float* d_coef;       /* [n_steps,5] */
    int32_t      n_steps;
    int32_t      step;         /* used when d_state == NULL */
    ds_sampler_state* d_state; /* device loop state or NULL */
    const float* d_noise;      /* injected noise or NULL */
    uint64_t     seed;         /* used when d_state == NULL */
    uint64_t     offset;       /* used when d_state == NULL */
    uint64_t     offset_inc;   /* torch's per-call offset increment for this numel */
    int32_t      rng_threads;  /* 256 * torch grid for this numel */
    int32_t      skip_rng_if_zero;
    const float* d_time_table; /* [n_steps+1] UNet time input per step (entry n_steps unused) or NULL */
    float*       d_time_out;   /* [time_len] or NULL */
    int32_t      time_len;
    int64_t      per_sample_numel; /* 0, or elements per sample: element i uses coefficient row i / per_sample_numel of d_coef
                                    * [B,5] (ddpm p_sample with a different t per sample, ddpm_modules/diffusion.py:64-67,
                                    * 195-203); needs d_state == NULL, step 0 */
} ds_step_args;
/* The kernel is launched with programmatic stream serialisation and reads *d_state and evaluates its first Philox block
 * BEFORE waiting for its predecessor in the stream: the predecessor must not be the kernel that last wrote *d_state (in the
 * sampling loop the whole UNet forward lies between two updates; after ds_randn_axpy(.., d_state, ..) run the forward or
 * synchronise the stream first). */
int ds_sampler_step(const ds_step_args* args, void* stream);
/* out = (base ? base : 0) + z*scale, z ~ torch.randn stream (the initial draw of p_sample_loop :194 /
 * InDI.inference :82).  Seed/offset from d_state (offset advanced by the kernel) or the arguments. */
int ds_randn_axpy(const float* d_base, float scale, float* d_out, int64_t numel, uint64_t seed,
                  uint64_t offset, ds_sampler_state* d_state, uint64_t offset_inc, int rng_threads, void* stream);

/* ------------------------------------------------------------------ tiling
 * TileIndexManager / stitch_predictions (data/tiling_manager.py:14-191, data/tile_stitcher.py:10-81).
 * Shapes are (F,H,W); integer arithmetic only, bit-exact. Host functions take host pointers. */
int ds_tile_counts(const int32_t data_shape[3], const int32_t grid_shape[3], const int32_t patch_shape[3],
                   int mode, int32_t counts[3], int64_t* total);
/* patch origin (f,h,w) of tiles [first, first+n) -> h_locations[n*3] */
int ds_tile_patch_locations(const int32_t data_shape[3], const int32_t grid_shape[3],
                            const int32_t patch_shape[3], int mode, int64_t first, int64_t n,
                            int32_t* h_locations);
/* gather tiles [first, first+n): frames (C,F,H,W) fp32|uint16 -> tiles (n,C,P,P) fp32
 * (SplitDataset.__getitem__ crop, data/split_dataset.py:239-249). elem_size 4 = fp32, 2 = uint16 */
int ds_crop_tiles(const void* d_frames, int elem_size, int C, const int32_t data_shape[3],
                  const int32_t grid_shape[3], const int32_t patch_shape[3], int mode,
                  int64_t first, int64_t n, float* d_tiles, void* stream);
/* tiles (N,C,P,P) for ALL N tiles -> frames (F,H,W,C); where tiles overlap the highest tile index wins,
 * exactly as the reference's sequential loop; uncovered pixels are 0. */
int ds_stitch_tiles(const float* d_tiles, int C, const int32_t data_shape[3], const int32_t grid_shape[3],
                    const int32_t patch_shape[3], int mode, float* d_out, void* stream);

/* Multi-GPU exchange of predicted tiles: only the destination box of a tile - the part stitch_predictions reads
 * (data/tile_stitcher.py:28-57: the inner grid cell, widened to the frame edge where the patch touches it) - travels.
 * ds_tile_regions:      host table, per tile [first, first+n): (frame, y_lo, y_hi, x_lo, x_hi) of its destination box.
 * ds_pack_tile_regions: tiles (n,C,P,P) whose GLOBAL indices are d_tile_ids[n] -> d_packed + d_offsets[i] as [C][hy][hx].
 * ds_stitch_packed:     packed boxes of ALL tiles (d_tile_offsets[total]: element offset of every global tile index) ->
 *                       frames (F,H,W,C); same last-writer-wins rule and bits as ds_stitch_tiles. */
int ds_tile_regions(const int32_t data_shape[3], const int32_t grid_shape[3], const int32_t patch_shape[3], int mode,
                    int64_t first, int64_t n, int32_t* h_regions);
int ds_pack_tile_regions(const float* d_tiles, int C, const int32_t data_shape[3], const int32_t grid_shape[3],
                         const int32_t patch_shape[3], int mode, const int64_t* d_tile_ids, const int64_t* d_offsets,
                         int64_t n, float* d_packed, void* stream);
int ds_stitch_packed(const float* d_packed, const int64_t* d_tile_offsets, int C, const int32_t data_shape[3],
                     const int32_t grid_shape[3], const int32_t patch_shape[3], int mode, float* d_out, void* stream);

/* Fused tile loader: crop + normalise + mix of tiles [first, first+n) from 2-channel frames (2,F,H,W) fp32|uint16
 * (SplitDataset.__getitem__ without transforms, data/split_dataset.py:237-278; normalize_target / normalize_inp :195-201).
 * d_input (n,1,P,P), d_target (n,2,P,P) or NULL; bit-exact with the reference's numpy arithmetic (float64 constants,
 * python-scalar channel weights). */
typedef struct {
    double  mean_target[2], std_target[2];
    double  mean_input, std_input;
    float   w0, w1;                         /* channel_weights */
    int32_t input_from_normalized_target;
} ds_tile_norm;
int ds_tile_batch(const void* d_frames, int elem_size, const int32_t data_shape[3], const int32_t grid_shape[3],
                  const int32_t patch_shape[3], int mode, int64_t first, int64_t n, const ds_tile_norm* norm,
                  float* d_input, float* d_target, void* stream);

/* ------------------------------------------------------------------ metrics
 * PSNR and RangeInvariantPsnr (core/psnr.py:46-82) of n_frames x C images of npix pixels, one pass over gt and pred.
 * Image (f, c) starts at base + f*frame_stride + c*channel_stride, pixels pixel_stride apart (elements), so both
 * (F,C,H,W) and the stitched (F,H,W,C) layout are read in place.  unnormalize: v*scale[c] + offset[c] in float64;
 * quantize_u16: the caller's cast to uint16 (split.py:198-203; the prediction is clamped to [0,65535] first).
 * d_out [n_frames*C][4] = PSNR, RangeInvariantPsnr, mse, range(gt).  fp64 accumulation. */
typedef struct {
    const float* d_gt;
    const float* d_pred;
    int64_t gt_frame_stride, gt_channel_stride, gt_pixel_stride;
    int64_t pred_frame_stride, pred_channel_stride, pred_pixel_stride;
    int32_t n_frames, C;
    int64_t npix;
    int32_t unnormalize, quantize_u16;
    const double* scale;                    /* host [C] or NULL */
    const double* offset;                   /* host [C] or NULL */
    float*  d_out;
    void*   d_workspace;
    size_t  workspace_bytes;
} ds_psnr_args;
size_t ds_psnr_workspace_bytes(int n_frames, int C, int64_t npix);
int ds_psnr(const ds_psnr_args* args, void* stream);

/* Tail of the time predictor (model/ddpm_modules/time_predictor.py:5-12, 38-45): d_out[b] = sum(relu(unet_out) * m) / sum(m),
 * m = sigmoid(conv7x7(x) + bias), sums over (channel, pixel).  d_x (B,cin,H,W), d_unet_out (B,cout,H,W) fp32 NCHW,
 * d_w_oihw (cout,cin,7,7); cin <= 8, cout <= 4.  The UNet part is ds_unet_forward with with_time_emb = 0. */
size_t ds_time_head_workspace_bytes(int B, int H, int W);
int ds_time_head_f32(const float* d_x, const float* d_unet_out, const float* d_w_oihw, const float* d_bias,
                     int B, int cin, int cout, int H, int W, float* d_out, void* d_workspace, size_t workspace_bytes,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFSPLIT_B200_H */
