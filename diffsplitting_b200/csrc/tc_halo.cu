// Fused  GroupNorm-apply + Swish -> 3x3 / 1x1 convolution  on the tensor cores, operands staged through registers.
//
// Why a second tensor-core conv kernel: for the 16..128-channel layers of the splitting UNets the TMA im2col path
// (tc.cu) fetches every input pixel 9 times in 32..128-byte rows and needs a separate normalisation pass before it.
// Here each CTA
//   1. loads the raw fp32 input pixels it needs ONCE per filter row (3 row segments of 130 pixels, 16-byte loads),
//      applies the per-(sample, channel) GroupNorm scale/shift + Swish of the reference Block
//      (model/sr3_modules/unet.py:80-91) in registers, converts to bf16 and writes the UMMA no-swizzle K-major layout
//      [8-channel plane][pixel][16 B];
//   2. issues all 9 taps x C/16 tcgen05.mma from that one copy: in the ZERO-PADDED, row-major flattened image
//      (pitch W+2) a filter tap is a constant shift of the flat index, i.e. just a different descriptor start address;
//   3. runs the common epilogue (bias, conditioning vector, fp32 residual, fp32 / bf16 stores).
// One tile = 128 consecutive flat padded positions (border positions are computed and discarded: 6 % waste at 64^2).
// The channel concat of the up path (unet.py:255) is two source pointers; weights arrive with cp.async.bulk.
#include "tc.cuh"
#include "tc_ptx.cuh"

namespace ds {

constexpr int HALO_THREADS = 192;
constexpr int HALO_SEG_PX = 136;           // 128 outputs + 2 halo pixels, padded to a multiple of 8
constexpr int HALO_MAX_SAMPLES = 12;
constexpr int HALO_MAX_GROUPS = 64;
constexpr size_t HALO_SMEM_LIMIT = 200 * 1024;

// n / d for 0 <= n < 2^31 with a precomputed multiplier (CUTLASS FastDivmod scheme): no integer division on device
struct FastDiv {
    uint32_t d, mul, shr;
};
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    if (d == 1) { f.mul = 0; f.shr = 0; return f; }
    uint32_t lg = 0;
    while ((1u << lg) < d) ++lg;
    const uint32_t p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
    return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) { return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr); }

struct HaloParams {
    const float* src_a; const float* src_b;     // fp32 NHWC [B,H,W,ca|cb]
    int ca, cb;
    const float2* stats;                        // (mean, rstd) [B][G] of the concat input, or
    const double* sums_a; const double* sums_b; // per-channel fp64 (sum, sumsq) [B][ca|cb][2]; all null = identity
    const float* gamma; const float* beta;
    int G, swish;
    const uint8_t* w;                           // bf16 [tap][kstep][plane(2)][Npad][8]
    TcEpi epi;
    int B, H, W, Wp, HpWp, total_q;
    int C, ksteps, ntaps, BN, n_tiles, Npad;
    // operand buffer geometry: `contig` = the three filter-row segments overlap inside ONE contiguous run of the flat
    // padded index space (W + 2 <= 139): every pixel is staged exactly once; else three separate 136-pixel segments
    int contig, plane_px, seg_stride_px;
    FastDiv div_hpwp, div_wp, div_planepx;
};

// no-swizzle K-major descriptor: rows 16 B apart inside an 8-row core matrix, SBO between 8-row groups,
// LBO between the two 8-element K halves of one K=16 step
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(HALO_THREADS) conv_halo_kernel(const __grid_constant__ HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 127u) & ~127u;
    uint8_t* gbase = smem_raw + (base - raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P = p.C >> 3;
    const uint32_t plane_bytes = (uint32_t)p.plane_px * 16u;
    const uint32_t a_off = 0, b_off = P * plane_bytes;
    const uint32_t b_tile = (uint32_t)p.BN * 16u;                       // one (tap, kstep, plane) block
    const uint32_t b_bytes = (uint32_t)p.ntaps * p.ksteps * 2u * b_tile;
    const uint32_t tab_off = b_off + b_bytes;
    const int q0 = blockIdx.x * 128;
    const int nt = blockIdx.y;
    // samples touched by the loaded span [q0 - Wp - 1, q0 + 128 + Wp]
    int b_first = (q0 - p.Wp - 1) / p.HpWp;
    if (q0 - p.Wp - 1 < 0) b_first = 0;
    int b_last = (q0 + 128 + p.Wp) / p.HpWp;
    if (b_last > p.B - 1) b_last = p.B - 1;
    const int nsamp = b_last - b_first + 1;
    const uint32_t bar_off = (tab_off + (uint32_t)nsamp * p.C * 8u + 15u) & ~15u;
    const uint32_t bfull = base + bar_off, mma_done = bfull + 8u, tmem_slot = bfull + 16u;
    uint8_t* red = gbase + bar_off + 32u;                                          // epilogue reduction buffer
    const uint32_t tmem_cols = p.BN <= 32 ? 32u : (p.BN <= 64 ? 64u : 128u);

    if (tid == 0) {
        mbar_init(bfull, 1);
        mbar_init(mma_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // ---- weights: ntaps * ksteps * 2 blocks of BN x 16 B, fetched asynchronously while the operands are staged
    if (warp == 0) {
        if (lane == 0) mbar_expect_tx(bfull, b_bytes);
        __syncwarp();
        const int ncopy = p.ntaps * p.ksteps * 2;
        for (int i = lane; i < ncopy; i += 32)
            bulk_load(base + b_off + i * b_tile, p.w + ((size_t)i * p.Npad + (size_t)nt * p.BN) * 16, b_tile, bfull);
    }

    pdl_wait();          // everything above (barriers, TMEM, weight fetch) overlapped the previous kernel's tail
    pdl_trigger();

    // ---- operand staging: raw fp32 -> normalise -> Swish -> bf16 -> [plane][pixel][16 B].
    // The raw pixel loads of the first batch are issued BEFORE the scale/shift table is built, so the two global-memory
    // round trips (statistics / affine parameters and pixels) overlap.
    float2* tab = reinterpret_cast<float2*>(gbase + tab_off);
    const int total = P * p.plane_px;
    const int q_first = p.ntaps == 9 ? q0 - p.Wp - 1 : q0 - 1;
    constexpr int NB = 4;                                  // chunks in flight per thread
    struct Chunk { float4 v0, v1; uint32_t dst; int tabi; };
    auto issue = [&](int idx, Chunk& ch) {
        ch.tabi = -1;
        ch.dst = 0xFFFFFFFFu;
        if (idx >= total) return;
        const int kp = fdiv(idx, p.div_planepx);
        const int px = idx - kp * p.plane_px;
        int q;
        if (p.contig) {
            q = q_first + px;
        } else {
            const int seg = px / HALO_SEG_PX;
            q = q0 + (seg - 1) * p.Wp - 1 + (px - seg * HALO_SEG_PX);
        }
        ch.dst = base + a_off + kp * plane_bytes + px * 16;
        if (q >= 0 && q < p.total_q) {
            const int b = fdiv(q, p.div_hpwp);
            const int rq = q - b * p.HpWp;
            const int yy = fdiv(rq, p.div_wp), xx = rq - yy * p.Wp;
            if (yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W) {
                const size_t pix = ((size_t)b * p.H + (yy - 1)) * p.W + (xx - 1);
                const int c0 = kp * 8;
                const float* src = c0 < p.ca ? p.src_a + pix * p.ca + c0 : p.src_b + pix * p.cb + (c0 - p.ca);
                ch.v0 = __ldg(reinterpret_cast<const float4*>(src));
                ch.v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
                ch.tabi = (b - b_first) * p.C + c0;
            }
        }
    };
    auto finish = [&](const Chunk& ch) {
        if (ch.dst == 0xFFFFFFFFu) return;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (ch.tabi >= 0) {
            const float2* tb = tab + ch.tabi;
            float x[8] = {ch.v0.x, ch.v0.y, ch.v0.z, ch.v0.w, ch.v1.x, ch.v1.y, ch.v1.z, ch.v1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 sc = tb[j];
                float y = fmaf(x[j], sc.x, sc.y);
                if (p.swish) {                      // y * sigmoid(y) = h * tanh(h) + h, h = y / 2: one MUFU op
                    const float h = 0.5f * y;
                    float th;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
                    y = fmaf(h, th, h);
                }
                x[j] = y;
            }
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
                w[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            val = make_uint4(w[0], w[1], w[2], w[3]);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ch.dst), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
    };

    Chunk cur[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) issue(tid + u * HALO_THREADS, cur[u]);

    // ---- per (sample in tile, channel) scale / shift of the fused GroupNorm: a = rstd * gamma, sh = beta - mean * a
    {
        const bool norm = p.stats || p.sums_a;
        const int cpg = norm ? p.C / p.G : 1;
        const double cnt = (double)p.H * p.W * cpg;
        for (int i = tid; i < nsamp * p.C; i += HALO_THREADS) {
            const int s = i / p.C, c = i - s * p.C, b = b_first + s;
            float a = 1.f, sh = 0.f;
            if (norm) {
                float mean, rstd;
                if (p.sums_a) {
                    // fold the producers' per-channel fp64 sums of this channel's group (it may straddle the two sources)
                    const int g = c / cpg;
                    double sm = 0.0, sq = 0.0;
                    for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
                        const double* src = cc < p.ca ? p.sums_a + ((size_t)b * p.ca + cc) * 2
                                                      : p.sums_b + ((size_t)b * p.cb + (cc - p.ca)) * 2;
                        sm += src[0];
                        sq += src[1];
                    }
                    const double mu = sm / cnt;
                    double var = sq / cnt - mu * mu;
                    if (var < 0.0) var = 0.0;
                    mean = (float)mu;
                    rstd = (float)(1.0 / sqrt(var + 1e-5));
                } else {
                    const float2 st = p.stats[(size_t)b * p.G + c / cpg];
                    mean = st.x;
                    rstd = st.y;
                }
                a = rstd * p.gamma[c];
                sh = p.beta[c] - mean * a;
            }
            tab[i] = make_float2(a, sh);
        }
    }
    __syncthreads();

#pragma unroll
    for (int u = 0; u < NB; ++u) finish(cur[u]);
    for (int i0 = tid + NB * HALO_THREADS; i0 < total; i0 += NB * HALO_THREADS) {
#pragma unroll
        for (int u = 0; u < NB; ++u) issue(i0 + u * HALO_THREADS, cur[u]);
#pragma unroll
        for (int u = 0; u < NB; ++u) finish(cur[u]);
    }
    fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();

    if (warp == 1) {
        if (lane == 0) {
            mbar_wait(bfull, 0);
            tc_fence_after();
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            for (int tap = 0; tap < p.ntaps; ++tap) {
                const int r = p.ntaps == 9 ? tap / 3 : 0;
                const int s = p.ntaps == 9 ? tap % 3 : 1;
                for (int kk = 0; kk < p.ksteps; ++kk) {
                    const uint32_t a_addr = base + a_off + (2 * kk) * plane_bytes + (uint32_t)(r * p.seg_stride_px + s) * 16u;
                    const uint32_t b_addr = base + b_off + (uint32_t)((tap * p.ksteps + kk) * 2) * b_tile;
                    umma_bf16(tmem_base, make_desc_nosw(a_addr, plane_bytes, 128u), make_desc_nosw(b_addr, b_tile, 128u),
                              idesc, (tap | kk) ? 1u : 0u);
                }
            }
            umma_commit(mma_done);
        }
        __syncwarp();
    } else if (warp >= 2) {
        const int qd = warp & 3;
        const int m = qd * 32 + lane;
        const int q = q0 + m;
        bool valid = q < p.total_q;
        int b = 0, oy = 0, ox = 0;
        if (valid) {
            b = fdiv(q, p.div_hpwp);
            const int rq = q - b * p.HpWp;
            const int yy = fdiv(rq, p.div_wp), xx = rq - yy * p.Wp;
            valid = yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W;
            oy = yy - 1;
            ox = xx - 1;
        }
        float add[16];
        if (valid) tc_epilogue_addend(p.epi, b, oy, ox, nt * p.BN, add);      // in flight while the MMAs run
        mbar_wait(mma_done, 0);
        tc_fence_after();
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, v);
            float f[16];
            if (valid) {
                if (c0) tc_epilogue_addend(p.epi, b, oy, ox, nt * p.BN + c0, add);
                tc_epilogue_write(p.epi, v, add, b, oy, ox, nt * p.BN + c0, f);
            }
            if (p.epi.sums_out) tc_epilogue_stats(p.epi, f, valid, b, nt * p.BN + c0, m, tid - 64, red);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------ host side
static int halo_plane_px(int ntaps, int W) {
    if (ntaps == 1) return HALO_SEG_PX;
    const int contig = 130 + 2 * (W + 2);
    return contig <= 3 * HALO_SEG_PX ? (contig + 7) / 8 * 8 : 3 * HALO_SEG_PX;
}

static size_t halo_smem_bytes(int C, int ntaps, int W, int BN, int nsamp) {
    return (size_t)(C / 8) * halo_plane_px(ntaps, W) * 16 + (size_t)ntaps * (C / 16) * 2 * BN * 16 + (size_t)nsamp * C * 8 +
           64 + TC_RED_BYTES + 128;
}

static int halo_samples_per_tile(int H, int W) {
    const int Wp = W + 2, HpWp = (H + 2) * Wp;
    return (128 + 2 * Wp + 2) / HpWp + 2;
}

static int halo_pick_bn(int cout, int C, int ntaps, int W, int nsamp, int64_t m_tiles) {
    const int npad = (cout + 15) / 16 * 16;
    int bn = 16;
    for (int c = 128; c >= 16; c >>= 1)
        if (npad % c == 0) { bn = c; break; }
    // every N tile re-stages (and re-normalises) the same operand pixels: only split N for shared memory or when the
    // grid would leave most SMs idle
    while (bn > 16 && (halo_smem_bytes(C, ntaps, W, bn, nsamp) > HALO_SMEM_LIMIT || m_tiles * (npad / bn) < 48)) bn >>= 1;
    return bn;
}

bool halo_conv_supported(int ca, int cb, int cout, int ks, int B, int H, int W) {
    const int C = ca + cb;
    if (ca <= 0 || ca % 8 || cb % 8 || C % 16 || C > 128) return false;
    if (!(ks == 1 || ks == 3)) return false;
    if ((int64_t)B * (H + 2) * (W + 2) >= (1ll << 31) - 4096) return false;
    const int nsamp = halo_samples_per_tile(H, W);
    if (nsamp > HALO_MAX_SAMPLES) return false;
    return halo_smem_bytes(C, ks * ks, W, 16, nsamp) <= HALO_SMEM_LIMIT;
}

size_t halo_packed_weight_bytes(int cout, int cin, int ks) {
    return (size_t)ks * ks * cin * ((cout + 15) / 16 * 16) * 2;
}

__global__ void pack_halo_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin, int ks,
                                        int npad) {
    const int ntaps = ks * ks;
    const size_t total = (size_t)ntaps * cin * npad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int j = (int)(r % 8); r /= 8;
        const int n = (int)(r % npad); r /= npad;
        const int plane = (int)(r % 2); r /= 2;
        const int kk = (int)(r % (cin / 16)); r /= (cin / 16);
        const int tap = (int)r;
        const int c = kk * 16 + plane * 8 + j;
        const float v = n < cout ? w[((size_t)n * cin + c) * ntaps + tap] : 0.f;
        out[i] = __float2bfloat16_rn(v);
    }
}

int halo_pack_conv_weight(const float* w_oihw, uint8_t* packed, int cout, int cin, int ks, cudaStream_t st) {
    DS_REQUIRE(cin % 16 == 0, "halo_pack: cin %d not a multiple of 16", cin);
    const int npad = (cout + 15) / 16 * 16;
    const size_t total = (size_t)ks * ks * cin * npad;
    int blocks = (int)((total + 255) / 256 > 2048 ? 2048 : (total + 255) / 256);
    pack_halo_weight_kernel<<<blocks, 256, 0, st>>>(w_oihw, reinterpret_cast<__nv_bfloat16*>(packed), cout, cin, ks, npad);
    DS_CHECK_LAUNCH("pack_halo_weight");
    return DS_OK;
}

int halo_launch_conv(const float* src_a, int ca, const float* src_b, int cb, const HaloNorm& norm, const uint8_t* w_packed,
                     int cout, int ks, int B, int H, int W, const ConvEpi& epi, float* out_f32, void* out_b16, float* out_nchw,
                     double* sums_out, cudaStream_t st) {
    DS_REQUIRE(halo_conv_supported(ca, cb, cout, ks, B, H, W), "halo conv: unsupported shape");
    HaloParams p;
    memset(&p, 0, sizeof(p));
    p.src_a = src_a; p.src_b = src_b; p.ca = ca; p.cb = cb;
    p.stats = norm.stats; p.sums_a = norm.sums_a; p.sums_b = norm.sums_b;
    p.gamma = norm.gamma; p.beta = norm.beta; p.G = norm.G > 0 ? norm.G : 1; p.swish = norm.swish;
    DS_REQUIRE(p.G <= HALO_MAX_GROUPS, "halo conv: %d groups > %d", p.G, HALO_MAX_GROUPS);
    p.epi.sums_out = sums_out;
    p.w = w_packed;
    p.epi.bias = epi.bias; p.epi.temb = epi.temb; p.epi.temb_off = epi.temb_off; p.epi.temb_stride = epi.temb_stride;
    p.epi.temb_bcast = epi.temb_bcast; p.epi.residual = epi.residual;
    p.epi.out_f32 = out_f32; p.epi.out_b16 = reinterpret_cast<__nv_bfloat16*>(out_b16); p.epi.out_nchw = out_nchw;
    p.epi.Cout = cout; p.epi.Ho = H; p.epi.Wo = W;
    p.B = B; p.H = H; p.W = W; p.Wp = W + 2; p.HpWp = (H + 2) * (W + 2);
    p.total_q = B * p.HpWp;
    p.C = ca + cb; p.ksteps = p.C / 16; p.ntaps = ks * ks;
    p.Npad = (cout + 15) / 16 * 16;
    const int nsamp = halo_samples_per_tile(H, W);
    const int64_t m_tiles = ((int64_t)p.total_q + 127) / 128;
    p.BN = halo_pick_bn(cout, p.C, p.ntaps, W, nsamp, m_tiles);
    p.n_tiles = p.Npad / p.BN;
    p.plane_px = halo_plane_px(p.ntaps, W);
    p.contig = (p.ntaps == 1 || p.plane_px != 3 * HALO_SEG_PX) ? 1 : 0;
    p.seg_stride_px = p.ntaps == 1 ? 0 : (p.contig ? p.Wp : HALO_SEG_PX);
    p.div_hpwp = make_fastdiv((uint32_t)p.HpWp);
    p.div_wp = make_fastdiv((uint32_t)p.Wp);
    p.div_planepx = make_fastdiv((uint32_t)p.plane_px);
    const size_t smem = halo_smem_bytes(p.C, p.ntaps, W, p.BN, nsamp);
    static bool attr_set = false;
    if (!attr_set) {
        DS_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        attr_set = true;
    }
    DS_CHECK_CUDA(launch_pdl(conv_halo_kernel, dim3((unsigned)m_tiles, p.n_tiles, 1), dim3(HALO_THREADS), smem, st, p));
    return DS_OK;
}

}  // namespace ds
