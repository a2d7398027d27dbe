// Per-sample persistent chain of  [GroupNorm(+Swish) ->] 3x3 / 1x1 convolutions  for the LOW-RESOLUTION levels of the UNet.
//
// At 8x8 (and 16x16) a layer of the splitting networks is 1..3 tiles of 128 output rows per sample: launched one layer per
// kernel it occupies a third of the SMs for ~10 us of which most is launch / drain / dependency latency (profiles/README.md).
// But GroupNorm, the convolutions and the residual adds never mix samples (model/sr3_modules/unet.py:80-123), so the whole
// run of consecutive low-resolution layers of ONE sample needs no grid-wide dependency: one CTA per sample walks the ops
// of the run, with CTA barriers where a kernel boundary used to be.
//
//   warp 0        weight producer: cp.async.bulk of (tap, 64-channel chunk) slabs [plane][Cout][16 B] through an mbarrier
//                 ring; runs ahead across op boundaries (weights are constants), also before griddepcontrol.wait
//   warp 1        MMA issuer: tcgen05.mma M=128, N=Cout (<= 256), accumulators in TMEM
//   warps 2..15   workers: build the scale/shift table from the producers' fp64 (sum, sumsq), stage the A operand
//                 (fp32 -> normalise -> Swish -> bf16, flat zero-padded index space exactly as tc_halo.cu, so a filter
//                 tap is a shift of the descriptor start address); warps 2..5 also run the epilogue (bias, time vector,
//                 fp32 residual, fp32 / bf16 stores, statistics of the output for the next GroupNorm)
//
// Activations stay in the planner's global buffers (L2 resident); tensors written earlier in the same launch are read with
// ld.global.cg (never the non-coherent path).  Statistics go to the same replicated fp64 slots the other kernels use.
#include <cuda.h>

#include "tc.cuh"
#include "tc_ptx.cuh"

namespace ds {

constexpr int CH_THREADS = 512;
constexpr int CH_WORKERS = CH_THREADS - 64;       // warps 2..15
constexpr int CH_MAX_GROUPS = 64;
constexpr int CH_MAX_C = 256;                     // input channels (concat) and output channels
constexpr int CH_INFLIGHT = 4;                    // staged work items in flight per worker thread
constexpr int CH_KC = 64;                         // channels per weight slab
constexpr size_t CH_SMEM_LIMIT = 208 * 1024;

struct ChainDev {                                 // device view of one op
    const void* src_a; const void* src_b;         // fp32 NHWC (src_b16 = 0) or bf16 NHWC (src_b16 = 1, no normalisation)
    const double* sums_a; const double* sums_b;   // replicated per-channel fp64 (sum, sumsq) of the sources (norm = 1)
    const float* gamma; const float* beta;
    const uint8_t* w;                             // bf16 [tap][chunk][plane][Npad][8]
    TcEpi epi;
    int ca, cb, G, swish, norm, ntaps, Npad, src_b16;
};

struct ChainParams {
    int nops, B, H, W, Wp, HpWp, mtiles, plane_px, stages;
    uint32_t stage_bytes, a_bytes, tmem_cols;
    TraceSlot trace;
    ChainDev ops[CHAIN_MAX_OPS];
};
static_assert(sizeof(ChainParams) <= 4000, "ChainParams must stay a by-value kernel parameter");

__device__ __forceinline__ void ch_bar_workers() { asm volatile("bar.sync 2, %0;" ::"n"(CH_WORKERS) : "memory"); }
__device__ __forceinline__ void ch_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(CH_THREADS) conv_chain_kernel(const __grid_constant__ ChainParams p) {
    extern __shared__ uint8_t ch_smem[];
    const uint32_t raw = smem_u32(ch_smem);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = ch_smem + (base - raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;

    const uint32_t a_off = 0;
    const uint32_t ring_off = p.a_bytes;
    const uint32_t tab_off = ring_off + (uint32_t)p.stages * p.stage_bytes;        // float2 [CH_MAX_C]
    const uint32_t gst_off = tab_off + CH_MAX_C * 8u;                              // float2 [CH_MAX_GROUPS]
    const uint32_t chs_off = gst_off + CH_MAX_GROUPS * 8u;                         // double2 [CH_MAX_C]
    const uint32_t bar_off = chs_off + CH_MAX_C * 16u;
    auto full_bar = [&](int s) { return base + bar_off + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return base + bar_off + 64u + 8u * (uint32_t)s; };
    const uint32_t a_full = base + bar_off + 128u, mma_done = a_full + 8u, tmem_slot = a_full + 16u;
    uint8_t* red = gbase + bar_off + 160u;

    trace_begin(p.trace);
    if (warp == 0 && elect_one()) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(a_full, 1);
        mbar_init(mma_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    pdl_trigger();
    if (warp == 0) {
        // ------------------------------------------------------------------ weight producer (constants: no dependency wait)
        if (elect_one()) {
            int u = 0;
            for (int oi = 0; oi < p.nops; ++oi) {
                const ChainDev& o = p.ops[oi];
                const int C = o.ca + o.cb;
                for (int t = 0; t < p.mtiles; ++t) {
                    const uint8_t* src = o.w;
                    for (int tap = 0; tap < o.ntaps; ++tap) {
                        for (int c0 = 0; c0 < C; c0 += CH_KC, ++u) {
                            const int kc = min(CH_KC, C - c0);
                            const uint32_t bytes = (uint32_t)kc * o.Npad * 2u;
                            const int s = u % p.stages;
                            mbar_wait(empty_bar(s), (((uint32_t)(u / p.stages)) & 1u) ^ 1u);
                            mbar_expect_tx(full_bar(s), bytes);
                            bulk_load(base + ring_off + (uint32_t)s * p.stage_bytes, src, bytes, full_bar(s));
                            src += bytes;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            int u = 0, it = 0;
            const uint32_t plane_bytes = (uint32_t)p.plane_px * 16u;
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);                      // SBO = 128 B, descriptor version 1
            const uint32_t a_lo0 = (((base + a_off) & 0x3FFFFu) >> 4) | ((plane_bytes >> 4) << 16);
            for (int oi = 0; oi < p.nops; ++oi) {
                const ChainDev& o = p.ops[oi];
                const int C = o.ca + o.cb;
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(o.Npad >> 3) << 17) | ((128u >> 4) << 24);
                const uint32_t b_lbo = ((uint32_t)o.Npad * 16u) >> 4;                // between the two 8-channel planes of a K step
                for (int t = 0; t < p.mtiles; ++t, ++it) {
                    mbar_wait(a_full, (uint32_t)it & 1u);
                    tc_fence_after();
                    uint32_t first = 1;
                    for (int tap = 0; tap < o.ntaps; ++tap) {
                        const int r = o.ntaps == 9 ? tap / 3 : 1;
                        const int sx = o.ntaps == 9 ? tap - 3 * r : 1;
                        const uint32_t a_tap = a_lo0 + (uint32_t)(r * p.Wp + sx);
                        for (int c0 = 0; c0 < C; c0 += CH_KC, ++u) {
                            const int kc = min(CH_KC, C - c0);
                            const int s = u % p.stages;
                            mbar_wait(full_bar(s), ((uint32_t)(u / p.stages)) & 1u);
                            tc_fence_after();
                            uint32_t a_lo = a_tap + (uint32_t)(c0 >> 3) * (plane_bytes >> 4);
                            uint32_t b_lo = (((base + ring_off + (uint32_t)s * p.stage_bytes) & 0x3FFFFu) >> 4) | (b_lbo << 16);
                            for (int k = 0; k < kc; k += 16) {
                                umma_bf16(tmem_base, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, first ? 0u : 1u);
                                first = 0;
                                a_lo += (2u * plane_bytes) >> 4;
                                b_lo += 2u * b_lbo;
                            }
                            umma_commit(empty_bar(s));
                        }
                    }
                    umma_commit(mma_done);
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ workers
        const int wt = tid - 64;
        pdl_wait();
        float2* tab = reinterpret_cast<float2*>(gbase + tab_off);
        float2* gst = reinterpret_cast<float2*>(gbase + gst_off);
        double2* chs = reinterpret_cast<double2*>(gbase + chs_off);
        const uint32_t plane_bytes = (uint32_t)p.plane_px * 16u;
        int it = 0;
        for (int oi = 0; oi < p.nops; ++oi) {
            const ChainDev& o = p.ops[oi];
            const int C = o.ca + o.cb;
            // ---- scale / shift table of the fused GroupNorm (a = rstd * gamma, sh = beta - mean * a)
            if (o.norm) {
                const int cpg = C / o.G;
                for (int cc = wt; cc < C; cc += CH_WORKERS) {
                    const bool firsts = cc < o.ca;
                    const double2* src = reinterpret_cast<const double2*>(
                        firsts ? o.sums_a + ((size_t)b * o.ca + cc) * 2 : o.sums_b + ((size_t)b * o.cb + (cc - o.ca)) * 2);
                    const size_t cstride = (size_t)p.B * (firsts ? o.ca : o.cb);
                    double2 v[TC_SUM_COPIES];
#pragma unroll
                    for (int k = 0; k < TC_SUM_COPIES; ++k) v[k] = __ldcg(src + k * cstride);
                    double sm = 0.0, sq = 0.0;
#pragma unroll
                    for (int k = 0; k < TC_SUM_COPIES; ++k) { sm += v[k].x; sq += v[k].y; }
                    chs[cc] = make_double2(sm, sq);
                }
                ch_bar_workers();
                const double inv_cnt = 1.0 / ((double)p.H * p.W * cpg);
                for (int g = wt; g < o.G; g += CH_WORKERS) {
                    double sm = 0.0, sq = 0.0;
                    for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) { sm += chs[cc].x; sq += chs[cc].y; }
                    const double mu = sm * inv_cnt;
                    double var = sq * inv_cnt - mu * mu;
                    if (var < 0.0) var = 0.0;
                    gst[g] = make_float2((float)mu, rsqrtf((float)var + 1e-5f));
                }
                ch_bar_workers();
                for (int cc = wt; cc < C; cc += CH_WORKERS) {
                    const float2 st = gst[cc / cpg];
                    const float a = st.y * __ldg(o.gamma + cc);
                    tab[cc] = make_float2(a, __ldg(o.beta + cc) - st.x * a);
                }
                ch_bar_workers();
            }
            for (int t = 0; t < p.mtiles; ++t, ++it) {
                // ---- stage the A operand of this tile: flat padded positions [128 t - Wp - 1, 128 t + 128 + Wp + 1)
                const int q_first = t * 128 - p.Wp - 1;
                const int npl = C >> 3;
                const int items = p.plane_px * npl;
                // work item = (8-channel plane, pixel); CH_INFLIGHT items in flight per thread (the loads are L2 round trips)
                for (int i0 = wt; i0 < items; i0 += CH_INFLIGHT * CH_WORKERS) {
                    uint4 raw0[CH_INFLIGHT], raw1[CH_INFLIGHT];
                    int kpv[CH_INFLIGHT], pxv[CH_INFLIGHT];
                    bool okv[CH_INFLIGHT];
#pragma unroll
                    for (int e = 0; e < CH_INFLIGHT; ++e) {
                        const int i = i0 + e * CH_WORKERS;
                        okv[e] = false;
                        kpv[e] = -1;
                        if (i >= items) continue;
                        const int kp = i / p.plane_px, px = i - kp * p.plane_px;
                        kpv[e] = kp; pxv[e] = px;
                        const int q = q_first + px;
                        if (q < 0 || q >= p.HpWp) continue;
                        const int yy = q / p.Wp, xx = q - yy * p.Wp;
                        if (yy < 1 || yy > p.H || xx < 1 || xx > p.W) continue;
                        okv[e] = true;
                        const size_t pix = ((size_t)b * p.H + (yy - 1)) * p.W + (xx - 1);
                        const int c0 = kp * 8;
                        if (o.src_b16) {
                            const __nv_bfloat16* src = c0 < o.ca ? reinterpret_cast<const __nv_bfloat16*>(o.src_a) + pix * o.ca + c0
                                                                 : reinterpret_cast<const __nv_bfloat16*>(o.src_b) + pix * o.cb + (c0 - o.ca);
                            raw0[e] = __ldcg(reinterpret_cast<const uint4*>(src));
                        } else {
                            const float* src = c0 < o.ca ? reinterpret_cast<const float*>(o.src_a) + pix * o.ca + c0
                                                         : reinterpret_cast<const float*>(o.src_b) + pix * o.cb + (c0 - o.ca);
                            raw0[e] = __ldcg(reinterpret_cast<const uint4*>(src));
                            raw1[e] = __ldcg(reinterpret_cast<const uint4*>(src) + 1);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < CH_INFLIGHT; ++e) {
                        if (kpv[e] < 0) continue;
                        uint4 val = make_uint4(0u, 0u, 0u, 0u);
                        if (okv[e]) {
                            if (o.src_b16) {
                                val = raw0[e];
                            } else {
                                float x[8] = {__uint_as_float(raw0[e].x), __uint_as_float(raw0[e].y), __uint_as_float(raw0[e].z),
                                              __uint_as_float(raw0[e].w), __uint_as_float(raw1[e].x), __uint_as_float(raw1[e].y),
                                              __uint_as_float(raw1[e].z), __uint_as_float(raw1[e].w)};
                                if (o.norm) {
                                    const float4* tb = reinterpret_cast<const float4*>(tab + kpv[e] * 8);
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const float4 sc = tb[j];
                                        x[2 * j] = fmaf(x[2 * j], sc.x, sc.y);
                                        x[2 * j + 1] = fmaf(x[2 * j + 1], sc.z, sc.w);
                                    }
                                }
                                if (o.swish) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {           // y * sigmoid(y) = h * tanh(h) + h, h = y / 2
                                        const float h = 0.5f * x[j];
                                        float th;
                                        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
                                        x[j] = fmaf(h, th, h);
                                    }
                                }
                                uint32_t w[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
                                    w[j] = *reinterpret_cast<const uint32_t*>(&h);
                                }
                                val = make_uint4(w[0], w[1], w[2], w[3]);
                            }
                        }
                        const uint32_t dst = base + a_off + (uint32_t)kpv[e] * plane_bytes + (uint32_t)pxv[e] * 16u;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
                    }
                }
                fence_proxy_async();
                ch_bar_workers();
                if (wt == 0) ch_mbar_arrive(a_full);
                // ---- epilogue (warps 2..5: TMEM lane quadrant = warp & 3)
                if (warp < 6) {
                    const int qd = warp & 3;
                    const int m = qd * 32 + lane;
                    const int q = t * 128 + m;
                    bool valid = q < p.HpWp;
                    int oy = 0, ox = 0;
                    if (valid) {
                        const int yy = q / p.Wp, xx = q - yy * p.Wp;
                        valid = yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W;
                        oy = yy - 1;
                        ox = xx - 1;
                    }
                    float add[16];
                    if (valid) tc_epilogue_addend<true>(o.epi, b, oy, ox, 0, add);
                    mbar_wait(mma_done, (uint32_t)it & 1u);
                    tc_fence_after();
                    for (int c0 = 0; c0 < o.Npad; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, v);
                        float f[16];
                        if (valid) {
                            if (c0) tc_epilogue_addend<true>(o.epi, b, oy, ox, c0, add);
                            tc_epilogue_write(o.epi, v, add, b, oy, ox, c0, f);
                        }
                        if (o.epi.sums_out) tc_epilogue_stats(o.epi, f, valid, b, c0, m, wt, b, b % TC_SUM_COPIES, red);
                    }
                    tc_fence_before();
                    __threadfence();
                }
                ch_bar_workers();          // outputs + statistics of this tile are visible to the CTA; A buffer and TMEM are free
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    trace_end(p.trace);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------ host side
static int chain_plane_px(int W) { return (130 + 2 * (W + 2) + 7) / 8 * 8; }

bool chain_level_supported(int H, int W) { return (H + 2) * (W + 2) <= 128 * CHAIN_MAX_MTILES; }

bool chain_conv_supported(int ca, int cb, int cout, int ks, int H, int W) {
    const int C = ca + cb;
    if (!chain_level_supported(H, W)) return false;
    if (ca <= 0 || ca % 8 || cb % 8 || C % 16 || C > CH_MAX_C) return false;
    if (cout <= 0 || (cout + 15) / 16 * 16 > CH_MAX_C) return false;
    return ks == 1 || ks == 3;
}

size_t chain_packed_weight_bytes(int cout, int cin, int ks) { return (size_t)ks * ks * cin * ((cout + 15) / 16 * 16) * 2; }

// w_oihw fp32 [cout][cin][ks][ks] -> bf16 [tap][chunk of <= 64 channels][8-channel plane][Npad rows][8]
__global__ void pack_chain_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin, int ks,
                                         int npad) {
    const int ntaps = ks * ks;
    const size_t total = (size_t)ntaps * cin * npad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        // planes are contiguous over the whole K range of a tap, so [chunk][plane] == [global plane]
        size_t r = i;
        const int j = (int)(r % 8); r /= 8;
        const int row = (int)(r % npad); r /= npad;
        const int plane = (int)(r % (cin / 8)); r /= (cin / 8);
        const int tap = (int)r;
        const int c = plane * 8 + j;
        const float v = row < cout ? w[((size_t)row * cin + c) * ntaps + tap] : 0.f;
        out[i] = __float2bfloat16_rn(v);
    }
}

int chain_pack_conv_weight(const float* w_oihw, uint8_t* packed, int cout, int cin, int ks, cudaStream_t st) {
    DS_REQUIRE(cin % 16 == 0, "chain_pack: cin %d not a multiple of 16", cin);
    const int npad = (cout + 15) / 16 * 16;
    const size_t total = (size_t)ks * ks * cin * npad;
    int blocks = (int)((total + 255) / 256 > 2048 ? 2048 : (total + 255) / 256);
    pack_chain_weight_kernel<<<blocks, 256, 0, st>>>(w_oihw, reinterpret_cast<__nv_bfloat16*>(packed), cout, cin, ks, npad);
    DS_CHECK_LAUNCH("pack_chain_weight");
    return DS_OK;
}

int chain_build(ChainPlan* plan, const ChainOpDesc* ops, int nops, int B, int H, int W) {
    DS_REQUIRE(nops >= 1 && nops <= CHAIN_MAX_OPS, "chain: %d ops (max %d)", nops, CHAIN_MAX_OPS);
    DS_REQUIRE(chain_level_supported(H, W), "chain: %dx%d does not fit %d tiles per sample", H, W, CHAIN_MAX_MTILES);
    static_assert(sizeof(ChainParams) <= sizeof(ChainPlan::params), "ChainPlan::params too small");
    ChainParams* p = reinterpret_cast<ChainParams*>(plan->params);
    memset(p, 0, sizeof(ChainParams));
    p->nops = nops; p->B = B; p->H = H; p->W = W; p->Wp = W + 2; p->HpWp = (H + 2) * (W + 2);
    p->mtiles = (p->HpWp + 127) / 128;
    p->plane_px = chain_plane_px(W);
    int cmax = 0, nmax = 0;
    for (int i = 0; i < nops; ++i) {
        const ChainOpDesc& d = ops[i];
        DS_REQUIRE(chain_conv_supported(d.ca, d.cb, d.cout, d.ks, H, W), "chain: op %d has an unsupported shape (%d+%d -> %d, k%d)", i,
                   d.ca, d.cb, d.cout, d.ks);
        DS_REQUIRE(!d.norm || (d.G > 0 && d.G <= CH_MAX_GROUPS && (d.ca + d.cb) % d.G == 0 && d.sums_a && (d.cb == 0 || d.sums_b)),
                   "chain: op %d has an invalid GroupNorm (G=%d)", i, d.G);
        DS_REQUIRE(!(d.src_b16 && (d.norm || d.swish)), "chain: op %d normalises a bf16 source", i);
        ChainDev& o = p->ops[i];
        o.src_a = d.src_a; o.src_b = d.src_b; o.ca = d.ca; o.cb = d.cb;
        o.sums_a = d.sums_a; o.sums_b = d.sums_b; o.gamma = d.gamma; o.beta = d.beta;
        o.G = d.G > 0 ? d.G : 1; o.swish = d.swish; o.norm = d.norm; o.src_b16 = d.src_b16;
        o.w = d.w; o.ntaps = d.ks * d.ks; o.Npad = (d.cout + 15) / 16 * 16;
        o.epi.bias = d.epi.bias; o.epi.temb = d.epi.temb; o.epi.temb_off = d.epi.temb_off; o.epi.temb_stride = d.epi.temb_stride;
        o.epi.temb_bcast = d.epi.temb_bcast; o.epi.residual = d.epi.residual;
        o.epi.out_f32 = d.out_f32; o.epi.out_b16 = reinterpret_cast<__nv_bfloat16*>(d.out_b16); o.epi.out_nchw = nullptr;
        o.epi.sums_out = d.sums_out; o.epi.sums_B = B;
        o.epi.Cout = d.cout; o.epi.Ho = H; o.epi.Wo = W;
        cmax = cmax > d.ca + d.cb ? cmax : d.ca + d.cb;
        nmax = nmax > o.Npad ? nmax : o.Npad;
    }
    p->a_bytes = (uint32_t)align_up((size_t)(cmax / 8) * p->plane_px * 16, 1024);
    p->stage_bytes = (uint32_t)align_up((size_t)nmax * CH_KC * 2, 1024);
    const size_t fixed = p->a_bytes + CH_MAX_C * 8 + CH_MAX_GROUPS * 8 + CH_MAX_C * 16 + 160 + TC_RED_BYTES + 1024 + 64;
    DS_REQUIRE(fixed + 2 * (size_t)p->stage_bytes <= CH_SMEM_LIMIT, "chain: shared memory (A %u B + 2 x %u B)", p->a_bytes, p->stage_bytes);
    int stages = (int)((CH_SMEM_LIMIT - fixed) / p->stage_bytes);
    p->stages = stages > 8 ? 8 : stages;
    p->tmem_cols = nmax <= 32 ? 32u : (nmax <= 64 ? 64u : (nmax <= 128 ? 128u : 256u));
    plan->smem_bytes = (int)(fixed + (size_t)p->stages * p->stage_bytes);
    plan->B = B;
    return DS_OK;
}

int chain_launch(const ChainPlan* plan, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        DS_CHECK_CUDA(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024));
        attr_set = true;
    }
    ChainParams p = *reinterpret_cast<const ChainParams*>(plan->params);
    p.trace = trace_next(7);
    DS_CHECK_CUDA(launch_pdl(conv_chain_kernel, dim3(plan->B), dim3(CH_THREADS), (size_t)plan->smem_bytes, st, p));
    return DS_OK;
}

}  // namespace ds
