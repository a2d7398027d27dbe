// Entry convolution of the UNets (model/sr3_modules/unet.py:199-200, model/ddpm_modules/unet.py:185-186): 3x3, padding 1,
// 1..16 input channels read straight from the caller's fp32 NCHW tensors (cond | x_t = two sources, the torch.cat of
// p_mean_variance is never materialised), 16..64 output channels.  Too thin for the tensor cores (K = 9..144): one thread
// per output pixel, all output channels in registers, the packed fp32 weights [tap * Cin + c][Npad] broadcast from shared
// memory.  Writes fp32 NHWC (+ bf16 NHWC) and, like the tensor-core epilogues, the per-channel fp64 (sum, sumsq) of its
// output for the first GroupNorm - replacing the generic CUDA-core conv + a separate statistics pass over its output.
#include "common.cuh"

namespace ds {

constexpr int EN_THREADS = 128;
constexpr int EN_MAX_K = 9 * 16;

struct EntryParams {
    const float* xa; const float* xb;     // fp32 NCHW [B, ca|cb, H, W]
    const float* w;                       // fp32 [9 * (ca + cb)][Npad]
    const float* bias;                    // [Cout] or null
    float* out_f32;                       // fp32 NHWC or null
    __nv_bfloat16* out_b16;               // bf16 NHWC or null
    double* sums_out;                     // [TC_SUM_COPIES][B][Cout][2] or null
    int ca, cb, Cout, Npad, B, H, W;
};

// column sums over the 32 lanes of a warp for 16 values per lane (recursive halving, 16 shuffles); lane l ends up with the
// sum of column 8 b4 + 4 b3 + 2 b2 + b1 (b_k = bit k of l)
__device__ __forceinline__ float en_colsum16(const float* v, int lane) {
    float a[8], b4[4], c2[2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool hi = lane & 16;
        a[j] = (hi ? v[j + 8] : v[j]) + __shfl_xor_sync(0xffffffffu, hi ? v[j] : v[j + 8], 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool hi = lane & 8;
        b4[j] = (hi ? a[j + 4] : a[j]) + __shfl_xor_sync(0xffffffffu, hi ? a[j] : a[j + 4], 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const bool hi = lane & 4;
        c2[j] = (hi ? b4[j + 2] : b4[j]) + __shfl_xor_sync(0xffffffffu, hi ? b4[j] : b4[j + 2], 4);
    }
    const bool hi = lane & 2;
    float d = (hi ? c2[1] : c2[0]) + __shfl_xor_sync(0xffffffffu, hi ? c2[0] : c2[1], 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

template <int NCH>      // output channels per thread = Npad (16, 32, 48 or 64)
__global__ void __launch_bounds__(EN_THREADS) conv_entry_kernel(const EntryParams p) {
    extern __shared__ float en_w[];                       // [K][NCH] weights, then [2][NCH] double accumulators
    const int Cin = p.ca + p.cb, K = 9 * Cin;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < K * NCH; i += EN_THREADS) en_w[i] = __ldg(p.w + i);
    double* acc_s = reinterpret_cast<double*>(en_w + ((K * NCH + 1) & ~1));
    for (int i = tid; i < 2 * NCH; i += EN_THREADS) acc_s[i] = 0.0;
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    const int HW = p.H * p.W;
    const int b = blockIdx.y;
    const int pix = blockIdx.x * EN_THREADS + tid;
    const bool valid = pix < HW;
    const int y = valid ? pix / p.W : 0, x = valid ? pix - y * p.W : 0;
    float acc[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) acc[j] = 0.f;
    if (valid) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
            if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) continue;
            for (int c = 0; c < Cin; ++c) {
                const float* src = c < p.ca ? p.xa + ((size_t)b * p.ca + c) * HW : p.xb + ((size_t)b * p.cb + (c - p.ca)) * HW;
                const float v = __ldg(src + (size_t)yy * p.W + xx);
                const float4* wr = reinterpret_cast<const float4*>(en_w + (size_t)(tap * Cin + c) * NCH);
#pragma unroll
                for (int j = 0; j < NCH / 4; ++j) {
                    const float4 w4 = wr[j];
                    acc[4 * j] = fmaf(v, w4.x, acc[4 * j]);
                    acc[4 * j + 1] = fmaf(v, w4.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(v, w4.z, acc[4 * j + 2]);
                    acc[4 * j + 3] = fmaf(v, w4.w, acc[4 * j + 3]);
                }
            }
        }
        if (p.bias) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) if (j < p.Cout) acc[j] += __ldg(p.bias + j);
        }
        const size_t off = ((size_t)b * HW + pix) * p.Cout;
        if (p.Cout == NCH) {
            if (p.out_f32) {
                float4* o = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
                for (int j = 0; j < NCH / 4; ++j) o[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            }
            if (p.out_b16) {
                uint4* o = reinterpret_cast<uint4*>(p.out_b16 + off);
#pragma unroll
                for (int j = 0; j < NCH / 8; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(acc[8 * j + 2 * k], acc[8 * j + 2 * k + 1]);
                        w[k] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    o[j] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                if (j < p.Cout) {
                    if (p.out_f32) p.out_f32[off + j] = acc[j];
                    if (p.out_b16) p.out_b16[off + j] = __float2bfloat16_rn(acc[j]);
                }
            }
        }
    }
    if (p.sums_out) {
        // per-channel (sum, sumsq) over the CTA's pixels: warp shuffles -> shared fp64 -> one atomic pair per channel
#pragma unroll
        for (int c0 = 0; c0 < NCH; c0 += 16) {
            float f[16], sq[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { f[j] = valid ? acc[c0 + j] : 0.f; sq[j] = f[j] * f[j]; }
            const float s1 = en_colsum16(f, lane), s2 = en_colsum16(sq, lane);
            if (!(lane & 1)) {
                const int col = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                atomicAdd(acc_s + 2 * col, (double)s1);
                atomicAdd(acc_s + 2 * col + 1, (double)s2);
            }
        }
        __syncthreads();
        const int copy = blockIdx.x % TC_SUM_COPIES;
        for (int i = tid; i < 2 * p.Cout; i += EN_THREADS)
            atomicAdd(p.sums_out + (((size_t)copy * p.B + b) * p.Cout) * 2 + i, acc_s[i]);
    }
}

bool entry_conv_supported(int ca, int cb, int cout, int ks) {
    const int npad = (cout + 15) / 16 * 16;
    return ks == 3 && ca > 0 && cb >= 0 && 9 * (ca + cb) <= EN_MAX_K && npad <= 64;
}

// w_packed: the fp32 pack of launch_pack_conv_weight_f32 with Npad = round_up(cout, 16)
int launch_conv_entry(const float* xa, int ca, const float* xb, int cb, const float* w_packed, int npad, const float* bias, int cout,
                      int B, int H, int W, float* out_f32, void* out_b16, double* sums_out, cudaStream_t st) {
    DS_REQUIRE(entry_conv_supported(ca, cb, cout, 3) && npad % 16 == 0 && npad >= cout && npad <= 64,
               "entry conv: unsupported shape %d+%d -> %d (row pitch %d)", ca, cb, cout, npad);
    EntryParams p;
    p.xa = xa; p.xb = xb; p.w = w_packed; p.bias = bias; p.out_f32 = out_f32; p.out_b16 = reinterpret_cast<__nv_bfloat16*>(out_b16);
    p.sums_out = sums_out; p.ca = ca; p.cb = cb; p.Cout = cout; p.Npad = npad; p.B = B; p.H = H; p.W = W;
    const dim3 grid((unsigned)cdiv((int64_t)H * W, EN_THREADS), (unsigned)B);
    const size_t smem = (size_t)((9 * (ca + cb) * npad + 1) & ~1) * 4 + 2 * npad * 8;
    cudaError_t e;
    switch (npad) {
        case 16: e = launch_pdl(conv_entry_kernel<16>, grid, dim3(EN_THREADS), smem, st, p); break;
        case 32: e = launch_pdl(conv_entry_kernel<32>, grid, dim3(EN_THREADS), smem, st, p); break;
        case 48: e = launch_pdl(conv_entry_kernel<48>, grid, dim3(EN_THREADS), smem, st, p); break;
        default: e = launch_pdl(conv_entry_kernel<64>, grid, dim3(EN_THREADS), smem, st, p); break;
    }
    DS_CHECK_CUDA(e);
    return DS_OK;
}

}  // namespace ds
