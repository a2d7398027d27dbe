// fp32 implicit-GEMM convolution on CUDA cores (precision mode DS_PREC_FP32).
//
// Reference ops covered (model/sr3_modules/unet.py): Conv3x3 pad1 (:87,193), Downsample 3x3 s2 (:71),
// Upsample nearest-x2 + 3x3 (:61-62, the upsample is folded into the gather: src = (y>>1, x>>1)),
// 1x1 res_conv / qkv / out (:102,120-121), the skip / condition torch.cat (unet.py:255,
// diffusion.py:158: two sources read as two K ranges, never materialised), and the epilogue adds:
// bias, FiLM / time-embedding vector (:49), residual (:110,142).
//
// GEMM view: out[m, n] = sum_k A[m, k] * Wp[k, n],  m = (b, oy, ox), k = (r, s, c), n = cout.
// Accumulation is plain fp32 FFMA in k order: this is the "<= 1e-5" path, not the fast one.
#include "common.cuh"

namespace ds {

struct ConvArgs {
    ConvSrc src;
    const float* w;      // [K][Npad]
    int Npad, Cout, ks, stride, pad;
    int B, Ho, Wo;
    int M, K, Ctot, Hin, Win;
    ConvEpi epi;
    float* out;
};

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(256) conv_f32_kernel(const ConvArgs p) {
    static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
    constexpr int TPR = 256 / BM;        // loader threads per A row
    constexpr int KPT = BK / TPR;        // consecutive k per loader thread
    constexpr int LDA = BM + 4;
    __shared__ __align__(16) float As[BK][LDA];
    __shared__ __align__(16) float Bs[BK][BN];

    const int t = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- loader role: one A row, KPT consecutive k
    const int lrow = t / TPR;
    const int lk = (t % TPR) * KPT;
    const int lm = m0 + lrow;
    const bool lvalid = lm < p.M;
    int lb = 0, iy0 = 0, ix0 = 0;
    if (lvalid) {
        lb = lm / (p.Ho * p.Wo);
        int rem = lm - lb * (p.Ho * p.Wo);
        int oy = rem / p.Wo;
        int ox = rem - oy * p.Wo;
        iy0 = oy * p.stride - p.pad;
        ix0 = ox * p.stride - p.pad;
    }

    // ---- compute role
    const int tx = t % (BN / TN);
    const int ty = t / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nchunks = (p.K + BK - 1) / BK;
    for (int kc = 0; kc < nchunks; ++kc) {
        const int kbase = kc * BK;
        // A tile
        if (VEC) {
#pragma unroll
            for (int v = 0; v < KPT / 4; ++v) {
                const int k = kbase + lk + 4 * v;
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lvalid && k < p.K) {
                    const int tap = k / p.Ctot;
                    const int c = k - tap * p.Ctot;
                    const int r = tap / p.ks;
                    const int s = tap - r * p.ks;
                    int iy = iy0 + r, ix = ix0 + s;
                    if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win) {
                        if (p.src.up) { iy >>= 1; ix >>= 1; }
                        const size_t pix = ((size_t)lb * p.src.Hs + iy) * p.src.Ws + ix;
                        const float* ptr = (c < p.src.ca) ? p.src.a + pix * p.src.ca + c
                                                          : p.src.b + pix * p.src.cb + (c - p.src.ca);
                        val = __ldg(reinterpret_cast<const float4*>(ptr));
                    }
                }
                As[lk + 4 * v + 0][lrow] = val.x;
                As[lk + 4 * v + 1][lrow] = val.y;
                As[lk + 4 * v + 2][lrow] = val.z;
                As[lk + 4 * v + 3][lrow] = val.w;
            }
        } else {
#pragma unroll
            for (int v = 0; v < KPT; ++v) {
                const int k = kbase + lk + v;
                float val = 0.f;
                if (lvalid && k < p.K) {
                    const int tap = k / p.Ctot;
                    const int c = k - tap * p.Ctot;
                    const int r = tap / p.ks;
                    const int s = tap - r * p.ks;
                    int iy = iy0 + r, ix = ix0 + s;
                    if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win) {
                        if (p.src.up) { iy >>= 1; ix >>= 1; }
                        const bool first = c < p.src.ca;
                        const float* base = first ? p.src.a : p.src.b;
                        const int cc = first ? c : c - p.src.ca;
                        const int cn = first ? p.src.ca : p.src.cb;
                        size_t off;
                        if (p.src.nchw) off = (((size_t)lb * cn + cc) * p.src.Hs + iy) * p.src.Ws + ix;
                        else off = (((size_t)lb * p.src.Hs + iy) * p.src.Ws + ix) * cn + cc;
                        val = __ldg(base + off);
                    }
                }
                As[lk + v][lrow] = val;
            }
        }
        // B tile
        for (int i = t; i < BK * BN; i += 256) {
            const int r = i / BN, cidx = i - r * BN;
            const int k = kbase + r;
            Bs[r][cidx] = (k < p.K) ? __ldg(p.w + (size_t)k * p.Npad + n0 + cidx) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: + bias + conditioning vector + residual
    const int HWo = p.Ho * p.Wo;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        const int b = m / HWo;
        const int pix = m - b * HWo;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= p.Cout) continue;
            float v = acc[i][j];
            if (p.epi.bias) v += p.epi.bias[n];
            if (p.epi.temb) v += p.epi.temb[(size_t)(p.epi.temb_bcast ? 0 : b) * p.epi.temb_stride + p.epi.temb_off + n];
            if (p.epi.residual) v += p.epi.residual[(size_t)m * p.Cout + n];
            if (p.epi.out_nchw) p.out[((size_t)b * p.Cout + n) * HWo + pix] = v;
            else p.out[(size_t)m * p.Cout + n] = v;
            if (p.epi.out2_bf16) reinterpret_cast<__nv_bfloat16*>(p.epi.out2_bf16)[(size_t)m * p.Cout + n] = __float2bfloat16_rn(v);
        }
    }
}

template <int BM, int BN, int TM, int TN>
static void launch_cfg(const ConvArgs& a, bool vec, cudaStream_t st) {
    dim3 grid(cdiv(a.M, BM), a.Npad / BN);
    if (vec) conv_f32_kernel<BM, BN, TM, TN, true><<<grid, 256, 0, st>>>(a);
    else conv_f32_kernel<BM, BN, TM, TN, false><<<grid, 256, 0, st>>>(a);
}

int launch_conv_f32(const ConvSrc& src, const float* w_packed, int Npad, int Cout, int ksize, int stride,
                    int B, int Ho, int Wo, const ConvEpi& epi, float* out, cudaStream_t st) {
    ConvArgs a;
    a.src = src;
    a.w = w_packed;
    a.Npad = Npad;
    a.Cout = Cout;
    a.ks = ksize;
    a.stride = stride;
    a.pad = ksize / 2;
    a.B = B;
    a.Ho = Ho;
    a.Wo = Wo;
    a.M = B * Ho * Wo;
    a.Ctot = src.ca + src.cb;
    a.K = ksize * ksize * a.Ctot;
    a.Hin = src.up ? 2 * src.Hs : src.Hs;
    a.Win = src.up ? 2 * src.Ws : src.Ws;
    a.epi = epi;
    a.out = out;
    DS_REQUIRE(ksize == 1 || ksize == 3, "conv: ksize %d unsupported", ksize);
    DS_REQUIRE(stride == 1 || stride == 2, "conv: stride %d unsupported", stride);
    DS_REQUIRE(Npad % 16 == 0 && Npad >= Cout, "conv: bad Npad %d for Cout %d", Npad, Cout);
    const bool vec = !src.nchw && (src.ca % 4 == 0) && (src.cb % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(src.a) & 15) == 0) &&
                     (src.b == nullptr || (reinterpret_cast<uintptr_t>(src.b) & 15) == 0);
    if (Npad % 64 == 0) launch_cfg<64, 64, 4, 4>(a, vec, st);
    else if (Npad % 32 == 0) launch_cfg<64, 32, 4, 2>(a, vec, st);
    else launch_cfg<128, 16, 4, 2>(a, vec, st);
    DS_CHECK_LAUNCH("conv_f32");
    return DS_OK;
}

// OIHW [Cout][Cin][ks][ks] -> packed [ (r*ks+s)*Cin + c ][ Npad ] (zero padded columns)
__global__ void pack_conv_weight_f32_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin,
                                            int ks, int Npad) {
    const int K = ks * ks * Cin;
    const size_t total = (size_t)K * Npad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % Npad);
        const int k = (int)(i / Npad);
        float v = 0.f;
        if (n < Cout) {
            const int tap = k / Cin, c = k - tap * Cin;
            v = w[((size_t)n * Cin + c) * ks * ks + tap];
        }
        out[i] = v;
    }
}

int launch_pack_conv_weight_f32(const float* w_oihw, float* w_packed, int Cout, int Cin, int ks, int Npad,
                                cudaStream_t st) {
    const size_t total = (size_t)ks * ks * Cin * Npad;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    pack_conv_weight_f32_kernel<<<blocks, 256, 0, st>>>(w_oihw, w_packed, Cout, Cin, ks, Npad);
    DS_CHECK_LAUNCH("pack_conv_weight_f32");
    return DS_OK;
}

}  // namespace ds
