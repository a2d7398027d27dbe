// fp32 single-head self-attention (precision mode DS_PREC_FP32), flash-style: the N x N score matrix of the
// reference (model/sr3_modules/unet.py:132-139, 64 MB per sample at N = 4096) is never materialised.
//
//   qkv : [B, N, 3C]  (q | k | v along the channel axis, i.e. the 1x1 qkv conv output in NHWC,
//                      matching  .view(b, 1, 3C, h, w).chunk(3, dim=2))
//   out : [B, N, C]   out[q] = sum_k softmax_k(q.k / sqrt(C)) v[k]
//
// One CTA = 32 queries x one CV-wide slice of the value/output channels; it walks the keys in blocks of 32
// with an online softmax (fp32 throughout, expf not __expf).  CTAs of different channel slices recompute the
// scores; this is the exactness path, the tensor-core path lives in attention_tc.cu.
#include "common.cuh"

namespace ds {

constexpr int AQ = 32;     // queries per CTA
constexpr int AK = 32;     // keys per inner block
constexpr int AC = 32;     // channel chunk of the score dot product

__device__ __forceinline__ float ld_f32(const float* p) { return *p; }
__device__ __forceinline__ float ld_f32(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_f32(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_f32(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <int CV, typename T>
__global__ void __launch_bounds__(256) attention_f32_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N, int C) {
    constexpr int CPT = CV / 32;                 // output channels per thread
    __shared__ float Qs[AQ][AC + 1];
    __shared__ float Ks[AK][AC + 1];
    __shared__ float Ss[AQ][AK + 1];
    __shared__ __align__(16) float Vs[AK][CV];
    __shared__ float s_alpha[AQ], s_l[AQ];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int q0 = blockIdx.x * AQ;
    const int cs = blockIdx.y * CV;
    const int b = blockIdx.z;
    const size_t row3 = (size_t)3 * C;
    const T* base = qkv + (size_t)b * N * row3;
    const float scale_div = sqrtf((float)C);

    // score role: 2x2 micro-tile
    const int sy = t >> 4, sx = t & 15;
    // softmax role: warp owns rows 4*warp .. 4*warp+3
    float m_run[4], l_run[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m_run[i] = -INFINITY; l_run[i] = 0.f; }
    // PV role
    const int og = t >> 5;            // query group: rows 4*og .. 4*og+3
    const int oc = (t & 31) * CPT;    // first channel of this thread inside the slice
    float acc[4][CPT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < N; k0 += AK) {
        float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
        for (int c0 = 0; c0 < C; c0 += AC) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = t + 256 * i;
                const int r = idx >> 5, c = idx & 31;
                const bool cin = (c0 + c) < C;
                Qs[r][c] = (q0 + r < N && cin) ? ld_f32(base + (size_t)(q0 + r) * row3 + c0 + c) : 0.f;
                Ks[r][c] = (k0 + r < N && cin) ? ld_f32(base + (size_t)(k0 + r) * row3 + C + c0 + c) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < AC; ++c) {
                const float qa = Qs[2 * sy][c], qb = Qs[2 * sy + 1][c];
                const float ka = Ks[2 * sx][c], kb = Ks[2 * sx + 1][c];
                s00 = fmaf(qa, ka, s00);
                s01 = fmaf(qa, kb, s01);
                s10 = fmaf(qb, ka, s10);
                s11 = fmaf(qb, kb, s11);
            }
            __syncthreads();
        }
        // V tile for this key block
        for (int idx = t; idx < AK * CV; idx += 256) {
            const int r = idx / CV, c = idx - r * CV;
            Vs[r][c] = (k0 + r < N) ? ld_f32(base + (size_t)(k0 + r) * row3 + 2 * C + cs + c) : 0.f;
        }
        const bool ka_ok = (k0 + 2 * sx) < N, kb_ok = (k0 + 2 * sx + 1) < N;
        Ss[2 * sy][2 * sx] = ka_ok ? s00 / scale_div : -INFINITY;
        Ss[2 * sy][2 * sx + 1] = kb_ok ? s01 / scale_div : -INFINITY;
        Ss[2 * sy + 1][2 * sx] = ka_ok ? s10 / scale_div : -INFINITY;
        Ss[2 * sy + 1][2 * sx + 1] = kb_ok ? s11 / scale_div : -INFINITY;
        __syncthreads();
        // online softmax, one warp per 4 rows, lane = key
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = 4 * warp + i;
            const float s = Ss[r][lane];
            float mx = s;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float m_new = fmaxf(m_run[i], mx);        // finite: key block has >= 1 valid key
            const float p = expf(s - m_new);                // exp(-inf) = 0 for masked keys
            float sum = p;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float alpha = expf(m_run[i] - m_new);     // first block: exp(-inf) = 0
            l_run[i] = l_run[i] * alpha + sum;
            m_run[i] = m_new;
            Ss[r][lane] = p;
            if (lane == 0) s_alpha[r] = alpha;
        }
        __syncthreads();
        // O = O * alpha + P V
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float al = s_alpha[4 * og + i];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[i][j] *= al;
        }
#pragma unroll 8
        for (int k = 0; k < AK; ++k) {
            float v[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) v[j] = Vs[k][oc + j];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float p = Ss[4 * og + i][k];
#pragma unroll
                for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(p, v[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s_l[4 * warp + i] = l_run[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + 4 * og + i;
        if (q >= N) continue;
        const float inv = s_l[4 * og + i];
#pragma unroll
        for (int j = 0; j < CPT; ++j) st_f32(out + ((size_t)b * N + q) * C + cs + oc + j, acc[i][j] / inv);
    }
}

template <typename T>
static void launch_attn_t(const T* qkv, T* out, int B, int N, int C, cudaStream_t st) {
    if (C % 128 == 0) {
        attention_f32_kernel<128, T><<<dim3(cdiv(N, AQ), C / 128, B), 256, 0, st>>>(qkv, out, N, C);
    } else if (C % 64 == 0) {
        attention_f32_kernel<64, T><<<dim3(cdiv(N, AQ), C / 64, B), 256, 0, st>>>(qkv, out, N, C);
    } else {
        attention_f32_kernel<32, T><<<dim3(cdiv(N, AQ), C / 32, B), 256, 0, st>>>(qkv, out, N, C);
    }
}

int launch_attention(const void* qkv, void* out, int B, int N, int C, int bf16, cudaStream_t st) {
    DS_REQUIRE(B > 0 && N > 0 && C > 0 && C % 32 == 0, "attention: unsupported shape B=%d N=%d C=%d (C %% 32)", B, N, C);
    if (bf16) launch_attn_t((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, B, N, C, st);
    else launch_attn_t((const float*)qkv, (float*)out, B, N, C, st);
    DS_CHECK_LAUNCH("attention_f32");
    return DS_OK;
}

}  // namespace ds
