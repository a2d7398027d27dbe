// Head of the time predictor: relu(unet_out) * sigmoid(conv7x7(x)) summed over (channel, pixel) and divided by the sum of
// the mask.  Reference: ForegroundMask + TimePredictor.forward (model/ddpm_modules/time_predictor.py:5-12, 38-45).
// The UNet part runs through ds_unet_forward (with_time_emb = 0); this is the remaining elementwise / reduction tail
// (7x7 conv with 1-6 input channels, sigmoid, relu, product, two sums, one division: ~8 torch kernels in the reference).
// One pass: a CTA per (sample, pixel range); 7x7 weights in shared memory; fp32 conv, fp64 sums; partials then one
// finishing warp per sample.
#include "common.cuh"

namespace ds {

constexpr int TH_MAX_W = 4 * 8 * 49;       // out_channel <= 4, in_channel <= 8

__global__ void __launch_bounds__(256) time_head_partial_kernel(const float* __restrict__ x, const float* __restrict__ u,
                                                                const float* __restrict__ w, const float* __restrict__ bias,
                                                                int cin, int cout, int H, int W, double* __restrict__ partial) {
    __shared__ float ws[TH_MAX_W];
    __shared__ float bs[4];
    __shared__ double red[8][2];
    for (int i = threadIdx.x; i < cout * cin * 49; i += 256) ws[i] = w[i];
    if (threadIdx.x < cout) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int b = blockIdx.y;
    const int64_t plane = (int64_t)H * W;
    const float* xb = x + (int64_t)b * cin * plane;
    const float* ub = u + (int64_t)b * cout * plane;
    double num = 0.0, den = 0.0;
    for (int64_t k = blockIdx.x * 256LL + threadIdx.x; k < plane; k += (int64_t)gridDim.x * 256) {
        const int y = (int)(k / W), xx = (int)(k - (int64_t)y * W);
        float acc[4];
#pragma unroll
        for (int co = 0; co < 4; ++co) acc[co] = co < cout ? bs[co] : 0.f;
        for (int ci = 0; ci < cin; ++ci) {
            for (int ky = 0; ky < 7; ++ky) {
                const int yy = y + ky - 3;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
                    const int xc = xx + kx - 3;
                    if (xc < 0 || xc >= W) continue;
                    const float v = __ldg(xb + ci * plane + (int64_t)yy * W + xc);
#pragma unroll
                    for (int co = 0; co < 4; ++co)
                        if (co < cout) acc[co] = fmaf(v, ws[(co * cin + ci) * 49 + ky * 7 + kx], acc[co]);
                }
            }
        }
#pragma unroll
        for (int co = 0; co < 4; ++co) {
            if (co < cout) {
                const float m = 1.f / (1.f + expf(-acc[co]));
                const float r = fmaxf(__ldg(ub + co * plane + k), 0.f);
                num += (double)(r * m);
                den += (double)m;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = num; red[threadIdx.x >> 5][1] = den; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        partial[((int64_t)b * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(32) time_head_final_kernel(const double* __restrict__ partial, int nblk, float* __restrict__ out) {
    const int b = blockIdx.x;
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 32) {
        num += partial[((int64_t)b * nblk + i) * 2];
        den += partial[((int64_t)b * nblk + i) * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if (threadIdx.x == 0) out[b] = (float)(num / den);
}

static int time_head_blocks(int B, int64_t plane) {
    int64_t want = (plane + 255) / 256;
    int64_t cap = (148 * 8 + B - 1) / B;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

}  // namespace ds

extern "C" size_t ds_time_head_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)B * ds::time_head_blocks(B, (int64_t)H * W) * 2 * sizeof(double);
}

extern "C" int ds_time_head_f32(const float* d_x, const float* d_unet_out, const float* d_w_oihw, const float* d_bias,
                                int B, int cin, int cout, int H, int W, float* d_out, void* d_workspace,
                                size_t workspace_bytes, void* stream) {
    using namespace ds;
    DS_REQUIRE(d_x && d_unet_out && d_w_oihw && d_bias && d_out && d_workspace, "time_head: null argument");
    DS_REQUIRE(B > 0 && H > 0 && W > 0, "time_head: empty input");
    DS_REQUIRE(cin >= 1 && cin <= 8 && cout >= 1 && cout <= 4, "time_head: %d -> %d channels (supported: <= 8 -> <= 4)", cin, cout);
    DS_REQUIRE(workspace_bytes >= ds_time_head_workspace_bytes(B, H, W), "time_head: workspace too small");
    const int nblk = time_head_blocks(B, (int64_t)H * W);
    time_head_partial_kernel<<<dim3(nblk, B), 256, 0, (cudaStream_t)stream>>>(d_x, d_unet_out, d_w_oihw, d_bias, cin, cout, H, W,
                                                                            (double*)d_workspace);
    DS_CHECK_LAUNCH("time_head_partial");
    time_head_final_kernel<<<B, 32, 0, (cudaStream_t)stream>>>((const double*)d_workspace, nblk, d_out);
    DS_CHECK_LAUNCH("time_head_final");
    return DS_OK;
}
