// GroupNorm(+Swish) over NHWC activations, optionally over the channel concat of two sources.
//
// Reference: Block = nn.GroupNorm(groups, dim) -> Swish (model/sr3_modules/unet.py:53-55, 80-91) and
// SelfAttention.norm (:119); eps = 1e-5, biased variance, affine.
//
// Two launches, both HBM/L2-bandwidth kernels (2 reads + 1 write algorithmic):
//  (1) gn_stats: per (sample, pixel slab) partial (sum, sum of squares) per group - 16-byte vector loads, per-thread
//      fp32 runs of <= 32 values flushed into fp64, combined in a FIXED order (no atomics on data => deterministic).
//      The last CTA of every sample (detected with one self-resetting counter) folds the <= 64 slab partials in
//      fp64 and publishes (mean, rstd) per group.
//  (2) gn_apply: y = (x - mean) * rstd * gamma + beta ; y * sigmoid(y); fp32 in, fp32 or bf16 out (bf16 = the
//      operand format of the tensor-core convolutions), 16-byte loads, 8/16-byte stores.
#include "common.cuh"

namespace ds {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_SPLIT = 64;
constexpr int GN_MAX_GROUPS = 64;
constexpr int GN_SUM_COPIES = TC_SUM_COPIES;      // replicated statistics accumulators (common.cuh)

int gn_nsplit(int B, int HW, int C) {
    int64_t per = (int64_t)HW * C;
    int want = (int)((592 + B - 1) / B);
    int cap = (int)((per + 4095) / 4096);
    int n = want < cap ? want : cap;
    if (n > GN_MAX_SPLIT) n = GN_MAX_SPLIT;
    if (n > HW) n = HW;
    if (n < 1) n = 1;
    return n;
}

// scratch layout: [B][GN_MAX_SPLIT][G] double2 partials | [B][G] float2 (mean, rstd)
static size_t gn_partial_bytes(int B, int G) { return (size_t)B * GN_MAX_SPLIT * G * 2 * sizeof(double); }
size_t gn_scratch_bytes(int B, int G) { return gn_partial_bytes(B, G) + align_up((size_t)B * G * 2 * sizeof(float), 256); }

__device__ __forceinline__ float4 load4(const float* __restrict__ a, int ca, const float* __restrict__ b, int cb, size_t pix, int c) {
    const float* p = (c < ca) ? a + pix * ca + c : b + pix * cb + (c - ca);
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float load1(const float* __restrict__ a, int ca, const float* __restrict__ b, int cb, size_t pix, int c) {
    return (c < ca) ? a[pix * ca + c] : b[pix * cb + (c - ca)];
}

// VEC: C % 4 == 0, C/4 <= 256, both sources 16-byte aligned with channel counts % 4 == 0
template <bool VEC>
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b,
                                                               int cb, int HW, int G, int nsplit, double* __restrict__ partial,
                                                               float2* __restrict__ stats, unsigned* __restrict__ counters) {
    const int C = ca + cb;
    const int cpg = C / G;
    const int bi = blockIdx.y, sp = blockIdx.x;
    const int p0 = (int)((int64_t)HW * sp / nsplit), p1 = (int)((int64_t)HW * (sp + 1) / nsplit);
    const int t = threadIdx.x;
    __shared__ double s_sum[GN_THREADS * 4], s_sq[GN_THREADS * 4];
    __shared__ double g_sum[GN_MAX_GROUPS], g_sq[GN_MAX_GROUPS];
    __shared__ bool s_last;
    pdl_wait();
    pdl_trigger();
    if (t < G) { g_sum[t] = 0.0; g_sq[t] = 0.0; }
    __syncthreads();
    const size_t base = (size_t)bi * HW;
    if (VEC) {
        const int q = C >> 2;                     // channel quads
        const int rows = GN_THREADS / q;
        const int r = t / q, cq = t - r * q;
        double ds_[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
        if (r < rows) {
            float s[4] = {0, 0, 0, 0}, qq[4] = {0, 0, 0, 0};
            int run = 0;
            auto acc = [&](const float4& v) {
                s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
                qq[0] = fmaf(v.x, v.x, qq[0]); qq[1] = fmaf(v.y, v.y, qq[1]);
                qq[2] = fmaf(v.z, v.z, qq[2]); qq[3] = fmaf(v.w, v.w, qq[3]);
            };
            int p = p0 + r;
            for (; p + 3 * rows < p1; p += 4 * rows) {      // 4 independent 16-byte loads in flight
                const float4 v0 = load4(a, ca, b, cb, base + p, cq * 4);
                const float4 v1 = load4(a, ca, b, cb, base + p + rows, cq * 4);
                const float4 v2 = load4(a, ca, b, cb, base + p + 2 * rows, cq * 4);
                const float4 v3 = load4(a, ca, b, cb, base + p + 3 * rows, cq * 4);
                acc(v0); acc(v1); acc(v2); acc(v3);
                run += 4;
                if (run >= 32) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { ds_[j] += (double)s[j]; dq[j] += (double)qq[j]; s[j] = 0.f; qq[j] = 0.f; }
                    run = 0;
                }
            }
            for (; p < p1; p += rows) acc(load4(a, ca, b, cb, base + p, cq * 4));
#pragma unroll
            for (int j = 0; j < 4; ++j) { ds_[j] += (double)s[j]; dq[j] += (double)qq[j]; }
        }
        // entry index = r * C + channel
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (r < rows) { s_sum[r * C + cq * 4 + j] = ds_[j]; s_sq[r * C + cq * 4 + j] = dq[j]; }
        }
        __syncthreads();
        if (t < G) {
            double s = 0.0, qv = 0.0;
            for (int rr = 0; rr < rows; ++rr)
                for (int cc = t * cpg; cc < (t + 1) * cpg; ++cc) { s += s_sum[rr * C + cc]; qv += s_sq[rr * C + cc]; }
            g_sum[t] = s;
            g_sq[t] = qv;
        }
    } else {
        for (int c0 = 0; c0 < C; c0 += GN_THREADS) {
            const int cw = min(GN_THREADS, C - c0);
            const int rows = GN_THREADS / cw;
            const int r = t / cw, c = c0 + (t - r * cw);
            double ds_ = 0.0, dq = 0.0;
            if (r < rows) {
                float s = 0.f, q = 0.f;
                int run = 0;
                for (int p = p0 + r; p < p1; p += rows) {
                    const float v = load1(a, ca, b, cb, base + p, c);
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
            if (t < G) {
                const int glo = t * cpg, ghi = glo + cpg;
                const int lo = max(glo, c0), hi = min(ghi, c0 + cw);
                double s = 0.0, q = 0.0;
                for (int rr = 0; rr < rows; ++rr)
                    for (int cc = lo; cc < hi; ++cc) { s += s_sum[rr * cw + (cc - c0)]; q += s_sq[rr * cw + (cc - c0)]; }
                g_sum[t] += s;
                g_sq[t] += q;
            }
            __syncthreads();
        }
    }
    if (t < G) {
        double* dst = partial + (((size_t)bi * GN_MAX_SPLIT + sp) * G + t) * 2;
        dst[0] = g_sum[t];
        dst[1] = g_sq[t];
    }
    // ---- last CTA of this sample publishes (mean, rstd)
    __threadfence();
    __syncthreads();
    if (t == 0) s_last = (atomicAdd(&counters[bi], 1u) == (unsigned)nsplit - 1u);
    __syncthreads();
    if (s_last) {
        __threadfence();
        // all threads fetch: thread (slice, g) sums every `slices`-th slab partial, then a fixed-order fold
        const int slices = GN_THREADS / G;
        const int g = t % G, sl = t / G;
        double ps = 0.0, pq = 0.0;
        if (sl < slices) {
            const double* src = partial + ((size_t)bi * GN_MAX_SPLIT * G + g) * 2;
            for (int k = sl; k < nsplit; k += slices) {
                ps += __ldcg(src + (size_t)k * G * 2);
                pq += __ldcg(src + (size_t)k * G * 2 + 1);
            }
        }
        __syncthreads();
        s_sum[t] = ps;
        s_sq[t] = pq;
        __syncthreads();
        if (t < G) {
            double s = 0.0, q = 0.0;
            for (int k = 0; k < slices; ++k) { s += s_sum[k * G + t]; q += s_sq[k * G + t]; }
            const double n = (double)HW * cpg;
            const double mean = s / n;
            double var = q / n - mean * mean;
            if (var < 0.0) var = 0.0;
            stats[(size_t)bi * G + t] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
        }
        if (t == 0) counters[bi] = 0u;
    }
}

__device__ __forceinline__ float swish_f(float y, bool fast) {
    return fast ? __fdividef(y, 1.0f + __expf(-y)) : y / (1.0f + expf(-y));
}

template <typename Tout, bool VEC>
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b,
                                                               int cb, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, Tout* __restrict__ out, int HW,
                                                               int G, int nchunk, int swish, const float2* __restrict__ stats,
                                                               const double* __restrict__ sums_a,
                                                               const double* __restrict__ sums_b) {
    const int C = ca + cb;
    const int cpg = C / G;
    const int bi = blockIdx.y;
    const int t = threadIdx.x;
    constexpr bool kFast = sizeof(Tout) == 2;          // bf16 output: fast exp is below the output rounding
    __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
    pdl_wait();
    pdl_trigger();
    if (sums_a) {        // per-channel fp64 (sum, sumsq) emitted by the producers' epilogues, replicated GN_SUM_COPIES times
        __shared__ double g_s[GN_MAX_GROUPS], g_q[GN_MAX_GROUPS];
        if (t < G) { g_s[t] = 0.0; g_q[t] = 0.0; }
        __syncthreads();
        // one thread per channel adds the copies (independent loads), then one shared-memory atomic pair per channel
        for (int c = t; c < C; c += GN_THREADS) {
            const bool first = c < ca;
            const double2* src = reinterpret_cast<const double2*>(first ? sums_a + ((size_t)bi * ca + c) * 2
                                                                        : sums_b + ((size_t)bi * cb + (c - ca)) * 2);
            const size_t cstride = (size_t)gridDim.y * (first ? ca : cb);      // gridDim.y = B
            double sm = 0.0, sq = 0.0;
#pragma unroll
            for (int k = 0; k < GN_SUM_COPIES; ++k) {
                const double2 v = src[k * cstride];
                sm += v.x;
                sq += v.y;
            }
            atomicAdd(&g_s[c / cpg], sm);
            atomicAdd(&g_q[c / cpg], sq);
        }
        __syncthreads();
        if (t < G) {
            const double n = (double)HW * cpg;
            const double mean = g_s[t] / n;
            double var = g_q[t] / n - mean * mean;
            if (var < 0.0) var = 0.0;
            s_mean[t] = (float)mean;
            s_rstd[t] = (float)(1.0 / sqrt(var + 1e-5));
        }
    } else if (t < G) {
        const float2 st = stats[(size_t)bi * G + t];
        s_mean[t] = st.x;
        s_rstd[t] = st.y;
    }
    __syncthreads();
    const size_t base = (size_t)bi * HW;
    if (VEC) {
        const int q = C >> 2;
        const int total = HW * q;                        // vectors in this sample (< 2^31 by construction)
        const int v0 = (int)((int64_t)total * blockIdx.x / nchunk), v1 = (int)((int64_t)total * (blockIdx.x + 1) / nchunk);
#pragma unroll 4
        for (int v = v0 + t; v < v1; v += GN_THREADS) {
            const int p = v / q, c = (v - p * q) * 4;
            const float4 x = load4(a, ca, b, cb, base + p, c);
            const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
            const float xs[4] = {x.x, x.y, x.z, x.w}, gs[4] = {ga.x, ga.y, ga.z, ga.w}, bs[4] = {be.x, be.y, be.z, be.w};
            float y[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int g = (c + j) / cpg;
                float yy = (xs[j] - s_mean[g]) * s_rstd[g];
                yy = fmaf(yy, gs[j], bs[j]);
                y[j] = swish ? swish_f(yy, kFast) : yy;
            }
            Tout* o = out + (base + p) * C + c;
            if constexpr (sizeof(Tout) == 2) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(y[0], y[1]), hi = __floats2bfloat162_rn(y[2], y[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(o) = pk;
            } else {
                *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
            }
        }
    } else {
        const int64_t total = (int64_t)HW * C;
        const int64_t e0 = total * blockIdx.x / nchunk, e1 = total * (blockIdx.x + 1) / nchunk;
        for (int64_t e = e0 + t; e < e1; e += GN_THREADS) {
            const int p = (int)(e / C);
            const int c = (int)(e - (int64_t)p * C);
            const int g = c / cpg;
            float y = (load1(a, ca, b, cb, base + p, c) - s_mean[g]) * s_rstd[g];
            y = fmaf(y, gamma[c], beta[c]);
            if (swish) y = swish_f(y, kFast);
            if constexpr (sizeof(Tout) == 2) out[(base + p) * C + c] = __float2bfloat16_rn(y);
            else out[(base + p) * C + c] = y;
        }
    }
}

float2* gn_stats_ptr(void* scratch, int B, int G) {
    return reinterpret_cast<float2*>((uint8_t*)scratch + gn_partial_bytes(B, G));
}

static bool gn_vec_ok(const float* a, int ca, const float* b, int cb, const float* gamma, const float* beta) {
    const int C = ca + cb;
    return (ca % 4 == 0) && (cb % 4 == 0) && (C / 4 <= GN_THREADS) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
           (b == nullptr || (reinterpret_cast<uintptr_t>(b) & 15) == 0) &&
           (gamma == nullptr || (reinterpret_cast<uintptr_t>(gamma) & 15) == 0) &&
           (beta == nullptr || (reinterpret_cast<uintptr_t>(beta) & 15) == 0);
}

// statistics only: (mean, rstd) per (sample, group) -> gn_stats_ptr(scratch); the normalisation itself is then fused
// into the consumer convolution's operand staging (tc_halo.cu)
int launch_gn_stats(const float* a, int ca, const float* b, int cb, int B, int HW, int G, void* scratch, unsigned* counters,
                    cudaStream_t st) {
    const int C = ca + cb;
    DS_REQUIRE(G >= 1 && G <= GN_MAX_GROUPS && C % G == 0, "groupnorm: %d channels not divisible into %d groups (max %d)",
               C, G, GN_MAX_GROUPS);
    DS_REQUIRE((int64_t)HW * C < (1ll << 31), "groupnorm: sample too large");
    DS_REQUIRE(B <= GN_MAX_BATCH, "groupnorm: batch %d > %d", B, GN_MAX_BATCH);
    const int nsplit = gn_nsplit(B, HW, C);
    double* partial = reinterpret_cast<double*>(scratch);
    float2* stats = gn_stats_ptr(scratch, B, G);
    if (gn_vec_ok(a, ca, b, cb, nullptr, nullptr))
        launch_pdl(gn_stats_kernel<true>, dim3(nsplit, B), dim3(GN_THREADS), 0, st, a, ca, b, cb, HW, G, nsplit, partial, stats, counters);
    else
        launch_pdl(gn_stats_kernel<false>, dim3(nsplit, B), dim3(GN_THREADS), 0, st, a, ca, b, cb, HW, G, nsplit, partial, stats, counters);
    DS_CHECK_LAUNCH("gn_stats");
    return DS_OK;
}

int launch_groupnorm(const float* a, int ca, const float* b, int cb, const float* gamma, const float* beta, void* out, int B,
                     int HW, int G, int swish, void* scratch, unsigned* counters, int out_bf16, cudaStream_t st) {
    const int C = ca + cb;
    DS_REQUIRE(G >= 1 && G <= GN_MAX_GROUPS && C % G == 0, "groupnorm: %d channels not divisible into %d groups (max %d)",
               C, G, GN_MAX_GROUPS);
    DS_REQUIRE((int64_t)HW * C < (1ll << 31), "groupnorm: sample too large");
    DS_REQUIRE(B <= GN_MAX_BATCH, "groupnorm: batch %d > %d", B, GN_MAX_BATCH);
    const int nsplit = gn_nsplit(B, HW, C);
    double* partial = reinterpret_cast<double*>(scratch);
    float2* stats = reinterpret_cast<float2*>((uint8_t*)scratch + gn_partial_bytes(B, G));
    const bool vec = (ca % 4 == 0) && (cb % 4 == 0) && (C / 4 <= GN_THREADS) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                     (b == nullptr || (reinterpret_cast<uintptr_t>(b) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(gamma) & 15) == 0) && ((reinterpret_cast<uintptr_t>(beta) & 15) == 0);
    if (vec) launch_pdl(gn_stats_kernel<true>, dim3(nsplit, B), dim3(GN_THREADS), 0, st, a, ca, b, cb, HW, G, nsplit, partial, stats, counters);
    else launch_pdl(gn_stats_kernel<false>, dim3(nsplit, B), dim3(GN_THREADS), 0, st, a, ca, b, cb, HW, G, nsplit, partial, stats, counters);
    DS_CHECK_LAUNCH("gn_stats");
    const int64_t total = (int64_t)HW * C;
    int nchunk = (int)((total + 4095) / 4096);
    if (nchunk > 65535) nchunk = 65535;
    if (nchunk < 1) nchunk = 1;
    const dim3 grid(nchunk, B);
    typedef __nv_bfloat16 bf;
    if (out_bf16) {
        if (vec) launch_pdl(gn_apply_kernel<bf, true>, grid, dim3(GN_THREADS), 0, st, a, ca, b, cb, gamma, beta, (bf*)out, HW, G, nchunk, swish, stats, nullptr, nullptr);
        else launch_pdl(gn_apply_kernel<bf, false>, grid, dim3(GN_THREADS), 0, st, a, ca, b, cb, gamma, beta, (bf*)out, HW, G, nchunk, swish, stats, nullptr, nullptr);
    } else {
        if (vec) launch_pdl(gn_apply_kernel<float, true>, grid, dim3(GN_THREADS), 0, st, a, ca, b, cb, gamma, beta, (float*)out, HW, G, nchunk, swish, stats, nullptr, nullptr);
        else launch_pdl(gn_apply_kernel<float, false>, grid, dim3(GN_THREADS), 0, st, a, ca, b, cb, gamma, beta, (float*)out, HW, G, nchunk, swish, stats, nullptr, nullptr);
    }
    DS_CHECK_LAUNCH("gn_apply");
    return DS_OK;
}

// ---- bf16 mode: statistics arrive as per-channel fp64 sums from the producers' epilogues; only the apply pass runs.
// (1) gn_fold_kernel: ONE CTA per sample folds the replicated per-channel sums into (mean, rstd) per group (before this
//     every apply CTA repeated the fold: C x 8 copies x 16 B per 16 KB of data, 0.29 TB/s on a 1536-channel tensor);
// (2) gn_apply_tab_kernel: a CTA builds the per-channel (scale, shift) table of its sample in shared memory once and walks
//     >= 32 K elements with 16-byte loads / 8- or 16-byte stores (no per-element integer division, no per-element
//     gamma / beta / mean / rstd fetches).
__global__ void __launch_bounds__(GN_THREADS) gn_fold_kernel(const double* __restrict__ sums_a, int ca,
                                                              const double* __restrict__ sums_b, int cb, int HW, int G,
                                                              float2* __restrict__ stats) {
    const int C = ca + cb, cpg = C / G, bi = blockIdx.x, t = threadIdx.x, B = gridDim.x;
    __shared__ double g_s[GN_MAX_GROUPS], g_q[GN_MAX_GROUPS];
    if (t < G) { g_s[t] = 0.0; g_q[t] = 0.0; }
    pdl_wait();
    pdl_trigger();
    __syncthreads();
    for (int c = t; c < C; c += GN_THREADS) {
        const bool first = c < ca;
        const double2* src = reinterpret_cast<const double2*>(first ? sums_a + ((size_t)bi * ca + c) * 2
                                                                    : sums_b + ((size_t)bi * cb + (c - ca)) * 2);
        const size_t cstride = (size_t)B * (first ? ca : cb);
        double2 v[GN_SUM_COPIES];
#pragma unroll
        for (int k = 0; k < GN_SUM_COPIES; ++k) v[k] = src[k * cstride];
        double sm = 0.0, sq = 0.0;
#pragma unroll
        for (int k = 0; k < GN_SUM_COPIES; ++k) { sm += v[k].x; sq += v[k].y; }
        atomicAdd(&g_s[c / cpg], sm);
        atomicAdd(&g_q[c / cpg], sq);
    }
    __syncthreads();
    if (t < G) {
        const double n = (double)HW * cpg;
        const double mean = g_s[t] / n;
        double var = g_q[t] / n - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[(size_t)bi * G + t] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
    }
}

constexpr int GN_TAB_MAX_C = 4096;

template <typename Tout>
__global__ void __launch_bounds__(GN_THREADS) gn_apply_tab_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b,
                                                                   int cb, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, Tout* __restrict__ out, int HW,
                                                                   int G, int nchunk, int swish, const float2* __restrict__ stats, int reverse) {
    extern __shared__ float2 gn_tab[];                 // [C] (scale, shift):  y = x * scale + shift
    // reverse: walk the tensor back to front.  The producer conv wrote it front to back, so its tail is what is still in L2 (126 MB
    // of a 0.13 - 0.8 GB tensor); and the consumer conv, which reads front to back, finds the head of THIS kernel's output there
    const int C = ca + cb, cpg = C / G, bi = reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y, t = threadIdx.x;
    const int chunk = reverse ? nchunk - 1 - (int)blockIdx.x : (int)blockIdx.x;
    constexpr bool kFast = sizeof(Tout) == 2;
    pdl_wait();
    pdl_trigger();
    for (int c = t; c < C; c += GN_THREADS) {
        // the reference's order: ((x - mean) * rstd) * gamma + beta; folded into one fma per element (<= 2 ulp of the
        // normalised value, far below the bf16 / tf32 operand rounding that follows)
        const float2 st = stats[(size_t)bi * G + c / cpg];
        const float sc = st.y * gamma[c];
        gn_tab[c] = make_float2(sc, fmaf(-st.x, sc, beta[c]));
    }
    __syncthreads();
    const size_t base = (size_t)bi * HW;
    const int q = C >> 2;
    const int64_t total = (int64_t)HW * q;
    const int64_t v0 = total * chunk / nchunk, v1 = total * (chunk + 1) / nchunk;
    auto one = [&](int64_t v) {
        const int p = (int)(v / q), c = (int)(v - (int64_t)p * q) * 4;
        const float4 x = load4(a, ca, b, cb, base + p, c);
        const float4 t0 = *reinterpret_cast<const float4*>(gn_tab + c), t1 = *reinterpret_cast<const float4*>(gn_tab + c + 2);
        float y[4] = {fmaf(x.x, t0.x, t0.y), fmaf(x.y, t0.z, t0.w), fmaf(x.z, t1.x, t1.y), fmaf(x.w, t1.z, t1.w)};
        if (swish) {
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] = swish_f(y[j], kFast);
        }
        Tout* o = out + (base + p) * C + c;
        if constexpr (sizeof(Tout) == 2) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(y[0], y[1]), hi = __floats2bfloat162_rn(y[2], y[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(o) = pk;
        } else {
            *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
        }
    };
    int64_t v = v0 + t;
    for (; v + 3 * GN_THREADS < v1; v += 4 * GN_THREADS) {      // four independent 16-byte loads in flight
        one(v); one(v + GN_THREADS); one(v + 2 * GN_THREADS); one(v + 3 * GN_THREADS);
    }
    for (; v < v1; v += GN_THREADS) one(v);
}

int launch_gn_apply_sums(const float* a, int ca, const float* b, int cb, const double* sums_a, const double* sums_b,
                         const float* gamma, const float* beta, void* out, int B, int HW, int G, int swish, int out_bf16,
                         void* scratch, cudaStream_t st) {
    const int C = ca + cb;
    DS_REQUIRE(G >= 1 && G <= GN_MAX_GROUPS && C % G == 0, "groupnorm: %d channels not divisible into %d groups", C, G);
    DS_REQUIRE((int64_t)HW * C < (1ll << 31), "groupnorm: sample too large");
    DS_REQUIRE(scratch, "groupnorm: scratch missing");
    float2* stats = gn_stats_ptr(scratch, B, G);
    launch_pdl(gn_fold_kernel, dim3(B), dim3(GN_THREADS), 0, st, sums_a, ca, sums_b, cb, HW, G, stats);
    DS_CHECK_LAUNCH("gn_fold");
    const int64_t total = (int64_t)HW * C;
    typedef __nv_bfloat16 bf;
    const bool tab = (ca % 4 == 0) && (cb % 4 == 0) && C <= GN_TAB_MAX_C && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                     (b == nullptr || (reinterpret_cast<uintptr_t>(b) & 15) == 0);
    if (tab) {
        // >= 32 K elements per CTA (the table costs C entries), at least ~4 CTAs per SM when the tensor allows it
        int nchunk = (int)((total + 32767) / 32768);
        if (nchunk < 1) nchunk = 1;
        if (nchunk > 65535) nchunk = 65535;
        const dim3 grid(nchunk, B);
        const size_t smem = (size_t)C * sizeof(float2);
        static int rev = -1;
        if (rev < 0) { const char* e = getenv("DIFFSPLIT_B200_GN_REVERSE"); rev = e ? atoi(e) : 1; }
        if (out_bf16) launch_pdl(gn_apply_tab_kernel<bf>, grid, dim3(GN_THREADS), smem, st, a, ca, b, cb, gamma, beta, (bf*)out, HW, G, nchunk, swish, (const float2*)stats, rev);
        else launch_pdl(gn_apply_tab_kernel<float>, grid, dim3(GN_THREADS), smem, st, a, ca, b, cb, gamma, beta, (float*)out, HW, G, nchunk, swish, (const float2*)stats, rev);
        DS_CHECK_LAUNCH("gn_apply_tab");
        return DS_OK;
    }
    int nchunk = (int)((total + 4095) / 4096);
    if (nchunk > 65535) nchunk = 65535;
    if (nchunk < 1) nchunk = 1;
    const dim3 grid(nchunk, B);
    if (out_bf16) launch_pdl(gn_apply_kernel<bf, false>, grid, dim3(GN_THREADS), 0, st, a, ca, b, cb, gamma, beta, (bf*)out, HW, G, nchunk, swish, (const float2*)stats, nullptr, nullptr);
    else launch_pdl(gn_apply_kernel<float, false>, grid, dim3(GN_THREADS), 0, st, a, ca, b, cb, gamma, beta, (float*)out, HW, G, nchunk, swish, (const float2*)stats, nullptr, nullptr);
    DS_CHECK_LAUNCH("gn_apply");
    return DS_OK;
}

// per-channel (sum, sumsq) of a tensor whose producer cannot emit them (the fp32 entry conv): out[b][c][2] += ...
__global__ void __launch_bounds__(GN_THREADS) ch_sums_kernel(const float* __restrict__ x, int C, int HW, int nsplit,
                                                              double* __restrict__ out) {
    const int bi = blockIdx.y, sp = blockIdx.x, t = threadIdx.x;
    const int p0 = (int)((int64_t)HW * sp / nsplit), p1 = (int)((int64_t)HW * (sp + 1) / nsplit);
    __shared__ double s_sum[GN_THREADS], s_sq[GN_THREADS];
    pdl_wait();
    pdl_trigger();
    const size_t base = (size_t)bi * HW;
    for (int c0 = 0; c0 < C; c0 += GN_THREADS) {
        const int cw = min(GN_THREADS, C - c0);
        const int rows = GN_THREADS / cw;
        const int r = t / cw, c = c0 + (t - r * cw);
        double ds_ = 0.0, dq = 0.0;
        if (r < rows) {
            float s = 0.f, q = 0.f;
            int run = 0;
            for (int p = p0 + r; p < p1; p += rows) {
                const float v = x[(base + p) * C + c];
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
        if (t < cw) {
            double s = 0.0, q = 0.0;
            for (int rr = 0; rr < rows; ++rr) { s += s_sum[rr * cw + t]; q += s_sq[rr * cw + t]; }
            double* dst = out + ((size_t)bi * C + c0 + t) * 2;
            atomicAdd(dst, s);
            atomicAdd(dst + 1, q);
        }
        __syncthreads();
    }
}

int launch_ch_sums(const float* x, int C, int B, int HW, double* out, cudaStream_t st) {
    const int nsplit = gn_nsplit(B, HW, C);
    launch_pdl(ch_sums_kernel, dim3(nsplit, B), dim3(GN_THREADS), 0, st, x, C, HW, nsplit, out);
    DS_CHECK_LAUNCH("ch_sums");
    return DS_OK;
}

}  // namespace ds
