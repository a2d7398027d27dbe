// fp32 GroupNorm(+Swish) over NHWC activations, optionally over the channel concat of two sources.
//
// Reference: Block = nn.GroupNorm(groups, dim) -> Swish (model/sr3_modules/unet.py:53-55, 80-91) and
// SelfAttention.norm (:119); eps = 1e-5, biased variance, affine.
//
// Two launches: (1) partial (sum, sum-of-squares) per (b, pixel slab, group), per-thread fp32 partials over
// short runs combined in fp64 in a fixed order (deterministic, no atomics); (2) every CTA of the apply
// kernel re-reduces the <=64 slab partials of its sample in fp64, then normalises its own pixel slab:
// y = (x - mean) * rstd * gamma + beta ; y * sigmoid(y).  Algorithmic traffic: 2 reads + 1 write.
#include "common.cuh"

namespace ds {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_SPLIT = 64;
constexpr int GN_MAX_GROUPS = 64;

int gn_nsplit(int B, int HW, int C) {
    // enough CTAs to cover the machine, but at least ~2048 elements per CTA
    int64_t per = (int64_t)HW * C;
    int want = (int)((592 + B - 1) / B);
    int cap = (int)((per + 2047) / 2048);
    int n = want < cap ? want : cap;
    if (n > GN_MAX_SPLIT) n = GN_MAX_SPLIT;
    if (n > HW) n = HW;
    if (n < 1) n = 1;
    return n;
}

size_t gn_scratch_bytes(int B, int G) { return (size_t)B * GN_MAX_SPLIT * G * 2 * sizeof(double); }

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f32(float* p, float v) { *p = v; }
__device__ __forceinline__ void from_f32(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__device__ __forceinline__ float load_cat(const T* __restrict__ a, int ca, const T* __restrict__ b, int cb, size_t pix, int c) {
    return to_f32((c < ca) ? a[pix * ca + c] : b[pix * cb + (c - ca)]);
}

template <typename T>
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const T* __restrict__ a, int ca, const T* __restrict__ b, int cb,
                                                               int HW, int G, int nsplit, double* __restrict__ partial) {
    const int C = ca + cb;
    const int cpg = C / G;
    const int bi = blockIdx.y, sp = blockIdx.x;
    const int p0 = (int)((int64_t)HW * sp / nsplit), p1 = (int)((int64_t)HW * (sp + 1) / nsplit);
    const int t = threadIdx.x;
    __shared__ double s_sum[GN_THREADS], s_sq[GN_THREADS];
    __shared__ double g_sum[GN_MAX_GROUPS], g_sq[GN_MAX_GROUPS];
    if (t < G) { g_sum[t] = 0.0; g_sq[t] = 0.0; }
    __syncthreads();
    const size_t base = (size_t)bi * HW;
    for (int c0 = 0; c0 < C; c0 += GN_THREADS) {
        const int cw = min(GN_THREADS, C - c0);          // channels handled in this sweep
        const int rows = GN_THREADS / cw;                 // pixels processed in parallel
        const int r = t / cw, c = c0 + (t - r * cw);
        double ds_ = 0.0, dq = 0.0;
        if (r < rows) {
            float s = 0.f, q = 0.f;
            int run = 0;
            for (int p = p0 + r; p < p1; p += rows) {
                const float v = load_cat(a, ca, b, cb, base + p, c);
                s += v;
                q = fmaf(v, v, q);
                if (++run == 32) { ds_ += (double)s; dq += (double)q; s = 0.f; q = 0.f; run = 0; }
            }
            ds_ += (double)s;
            dq += (double)q;
        }
        s_sum[t] = ds_;
        s_sq[t] = dq;
        __syncthreads();
        // thread g (< G) gathers the entries of its group inside this sweep, fixed order
        if (t < G) {
            const int glo = t * cpg, ghi = glo + cpg;      // channel range of group t
            const int lo = max(glo, c0), hi = min(ghi, c0 + cw);
            double s = 0.0, q = 0.0;
            for (int rr = 0; rr < rows; ++rr)
                for (int cc = lo; cc < hi; ++cc) {
                    s += s_sum[rr * cw + (cc - c0)];
                    q += s_sq[rr * cw + (cc - c0)];
                }
            g_sum[t] += s;
            g_sq[t] += q;
        }
        __syncthreads();
    }
    if (t < G) {
        double* dst = partial + (((size_t)bi * GN_MAX_SPLIT + sp) * G + t) * 2;
        dst[0] = g_sum[t];
        dst[1] = g_sq[t];
    }
}

template <typename T>
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const T* __restrict__ a, int ca, const T* __restrict__ b, int cb,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, T* __restrict__ out,
                                                               int HW, int G, int nsplit, int nchunk, int swish,
                                                               const double* __restrict__ partial) {
    const int C = ca + cb;
    const int cpg = C / G;
    const int bi = blockIdx.y;
    const int t = threadIdx.x;
    __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
    if (t < G) {
        double s = 0.0, q = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) {
            const double* src = partial + (((size_t)bi * GN_MAX_SPLIT + sp) * G + t) * 2;
            s += src[0];
            q += src[1];
        }
        const double n = (double)HW * cpg;
        const double mean = s / n;
        double var = q / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[t] = (float)mean;
        s_rstd[t] = (float)(1.0 / sqrt(var + 1e-5));
    }
    __syncthreads();
    const int64_t total = (int64_t)HW * C;
    const int64_t e0 = total * blockIdx.x / nchunk, e1 = total * (blockIdx.x + 1) / nchunk;
    const size_t base = (size_t)bi * HW;
    for (int64_t e = e0 + t; e < e1; e += GN_THREADS) {
        const int p = (int)(e / C);
        const int c = (int)(e - (int64_t)p * C);
        const int g = c / cpg;
        const float x = load_cat(a, ca, b, cb, base + p, c);
        float y = (x - s_mean[g]) * s_rstd[g];
        y = fmaf(y, gamma[c], beta[c]);
        if (swish) y = y / (1.0f + expf(-y));
        from_f32(out + (base + p) * C + c, y);
    }
}

int launch_groupnorm(const void* a, int ca, const void* b, int cb, const float* gamma, const float* beta, void* out, int B,
                     int HW, int G, int swish, void* scratch, int bf16, cudaStream_t st) {
    const int C = ca + cb;
    DS_REQUIRE(G >= 1 && G <= GN_MAX_GROUPS && C % G == 0, "groupnorm: %d channels not divisible into %d groups (max %d)",
               C, G, GN_MAX_GROUPS);
    const int nsplit = gn_nsplit(B, HW, C);
    double* partial = reinterpret_cast<double*>(scratch);
    typedef __nv_bfloat16 bf;
    if (bf16) gn_stats_kernel<bf><<<dim3(nsplit, B), GN_THREADS, 0, st>>>((const bf*)a, ca, (const bf*)b, cb, HW, G, nsplit, partial);
    else gn_stats_kernel<float><<<dim3(nsplit, B), GN_THREADS, 0, st>>>((const float*)a, ca, (const float*)b, cb, HW, G, nsplit, partial);
    DS_CHECK_LAUNCH("gn_stats");
    int64_t total = (int64_t)HW * C;
    int nchunk = (int)((total + 16383) / 16384);
    if (nchunk > 65535) nchunk = 65535;
    if (nchunk < 1) nchunk = 1;
    if (bf16)
        gn_apply_kernel<bf><<<dim3(nchunk, B), GN_THREADS, 0, st>>>((const bf*)a, ca, (const bf*)b, cb, gamma, beta, (bf*)out, HW, G,
                                                                   nsplit, nchunk, swish, partial);
    else
        gn_apply_kernel<float><<<dim3(nchunk, B), GN_THREADS, 0, st>>>((const float*)a, ca, (const float*)b, cb, gamma, beta,
                                                                      (float*)out, HW, G, nsplit, nchunk, swish, partial);
    DS_CHECK_LAUNCH("gn_apply");
    return DS_OK;
}

}  // namespace ds
