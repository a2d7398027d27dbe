// Per-sample persistent chain of  [GroupNorm(+Swish) ->] 3x3 / 1x1 convolutions  for the LOW-RESOLUTION levels of the UNet.
//
// At 8x8 (and 16x16) a layer of the splitting networks is 1..3 tiles of 128 output rows per sample: launched one layer per
// kernel it occupies a third of the SMs for ~10 us of which most is launch / drain / dependency latency (profiles/README.md).
// But GroupNorm, the convolutions and the residual adds never mix samples (model/sr3_modules/unet.py:80-123), so the whole
// run of consecutive low-resolution layers of ONE sample needs no grid-wide dependency.
//
// One thread-block CLUSTER per sample walks the ops of the run; the CTAs of the cluster split the OUTPUT channels
// (slices of `ns` = 16 or 32): an SM can only ingest ~25-30 B/clk from L2 (measured: one CTA streaming the 295 KB of a
// 128x128x3x3 layer took 12k cycles), so the weights of a layer must be spread over several SMs.  Every CTA stages the
// full A operand (all input channels of its sample), multiplies it with its weight slice, writes its output channels and
// their statistics to global memory, and the cluster meets at an mbarrier-based cluster barrier where a kernel boundary
// used to be.
//
//   warp 0        weight producer: cp.async.bulk of (tap, 128-channel chunk) slabs [plane][ns][16 B] through an mbarrier
//                 ring; runs ahead across op boundaries (weights are constants), also before griddepcontrol.wait
//   warp 1        MMA issuer: tcgen05.mma M=128, N=ns, accumulators in TMEM
//   warps 2..15   workers: build the scale/shift table from the producers' fp64 (sum, sumsq), stage the A operand
//                 (fp32 -> normalise -> Swish -> bf16, flat zero-padded index space exactly as tc_halo.cu, so a filter
//                 tap is a shift of the descriptor start address); warps 2..5 also run the epilogue (bias, time vector,
//                 fp32 residual, fp32 / bf16 stores, statistics of the output for the next GroupNorm)
//
// Activations stay in the planner's global buffers (L2 resident); tensors written earlier in the same launch are read with
// ld.global.cg (never the non-coherent path).  Statistics go to copy 0 of the replicated fp64 slots the other kernels use
// (plain stores: every channel has exactly one owner CTA).
#include <cuda.h>

#include "tc.cuh"
#include "tc_ptx.cuh"

namespace ds {

constexpr int CH_THREADS = 512;
constexpr int CH_WORKERS = CH_THREADS - 64;       // warps 2..15
constexpr int CH_MAX_GROUPS = 64;
constexpr int CH_MAX_C = 256;                     // input channels (concat) and output channels
constexpr int CH_INFLIGHT = 3;                    // staged work items in flight per worker thread
constexpr int CH_KC = 128;                        // channels per weight unit = (tap, channel chunk)
constexpr int CH_UNITS_PER_STAGE = 3;             // full-size units per ring slot (fewer, larger copies and barrier waits)
constexpr int CH_MAX_PX = 176;                    // staged pixels per tile: 130 + 2 (W + 2), W <= 20
constexpr size_t CH_SMEM_LIMIT = 208 * 1024;

struct ChainDev {                                 // device view of one op
    const void* src_a; const void* src_b;         // fp32 NHWC (src_b16 = 0) or bf16 NHWC (src_b16 = 1, no normalisation)
    const double* sums_a; const double* sums_b;   // replicated per-channel fp64 (sum, sumsq) of the sources (norm = 1)
    const float* gamma; const float* beta;
    const uint8_t* w;                             // bf16 [slice][tap][plane][ns][8]
    TcEpi epi;
    int ca, cb, G, swish, norm, ntaps, src_b16;
    const float* xsrc_a; const float* xsrc_b;     // optional raw second operand (folded 1x1 res_conv), fp32 NHWC
    const uint8_t* wx;                            // its weights, bf16 [slice][plane][ns][8]
    int xca, xcb;
};

struct ChainParams {
    int nops, B, H, W, Wp, HpWp, mtiles, plane_px, stages, CL, ns;
    uint32_t stage_bytes, a_bytes;
    FastDiv div_px, div_wp;
    TraceSlot trace;
    long long* dbg;                               // optional: clock64 stamps of CTA 0, 6 per op (first tile)
    ChainDev ops[CHAIN_MAX_OPS];
};
static_assert(sizeof(ChainParams) <= 4000, "ChainParams must stay a by-value kernel parameter");
#define CH_STAMP(k) do { if (p.dbg && blockIdx.x == 0 && wt == 0 && t == 0) p.dbg[oi * 6 + (k)] = clock64(); } while (0)

__device__ __forceinline__ void ch_bar_workers() { asm volatile("bar.sync 2, %0;" ::"n"(CH_WORKERS) : "memory"); }
__device__ __forceinline__ void ch_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release at cluster scope) on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void cluster_mbar_arrive(uint32_t local_bar, uint32_t rank) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}" ::"r"(local_bar), "r"(rank)
        : "memory");
}
__device__ __forceinline__ void cluster_mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

// bias + conditioning vector + fp32 residual of one (pixel, 16-channel chunk); the chunk lies inside Cout or is a tail
__device__ __forceinline__ void ch_addend(const TcEpi& e, int b, int oy, int ox, int n0, float (&add)[16]) {
    if (n0 + 16 > e.Cout) {
        tc_epilogue_addend<true>(e, b, oy, ox, n0, add);
        if (e.bias2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n0 + j < e.Cout) add[j] += __ldg(e.bias2 + n0 + j);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) add[j] = 0.f;
    if (e.bias) {
        const float4* s = reinterpret_cast<const float4*>(e.bias + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float4 v = __ldg(s + j); add[4 * j] = v.x; add[4 * j + 1] = v.y; add[4 * j + 2] = v.z; add[4 * j + 3] = v.w; }
    }
    if (e.bias2) {
        const float4* s = reinterpret_cast<const float4*>(e.bias2 + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float4 v = __ldg(s + j); add[4 * j] += v.x; add[4 * j + 1] += v.y; add[4 * j + 2] += v.z; add[4 * j + 3] += v.w; }
    }
    if (e.temb) {
        const float4* s = reinterpret_cast<const float4*>(e.temb + (size_t)(e.temb_bcast ? 0 : b) * e.temb_stride + e.temb_off + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float4 v = __ldg(s + j); add[4 * j] += v.x; add[4 * j + 1] += v.y; add[4 * j + 2] += v.z; add[4 * j + 3] += v.w; }
    }
    if (e.residual) {
        const float4* s = reinterpret_cast<const float4*>(e.residual + (((size_t)b * e.Ho + oy) * e.Wo + ox) * e.Cout + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float4 v = __ldcg(s + j); add[4 * j] += v.x; add[4 * j + 1] += v.y; add[4 * j + 2] += v.z; add[4 * j + 3] += v.w; }
    }
}

__global__ void __launch_bounds__(CH_THREADS) conv_chain_kernel(const __grid_constant__ ChainParams p) {
    extern __shared__ uint8_t ch_smem[];
    const uint32_t raw = smem_u32(ch_smem);
    const uint32_t base = (raw + 1023u) & ~1023u;      // identical offset in every CTA of the cluster (same kernel, same layout)
    uint8_t* gbase = ch_smem + (base - raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / p.CL, slice = blockIdx.x - b * p.CL;       // cluster = CL consecutive CTAs = one sample
    const int ncol0 = slice * p.ns;                                       // first output channel of this CTA

    const uint32_t a_off = 0;
    const uint32_t ring_off = p.a_bytes;
    const uint32_t tab_off = ring_off + (uint32_t)p.stages * p.stage_bytes;        // float2 [CH_MAX_C]
    const uint32_t gst_off = tab_off + CH_MAX_C * 8u;                              // float2 [CH_MAX_GROUPS]
    const uint32_t chs_off = gst_off + CH_MAX_GROUPS * 8u;                         // double2 [CH_MAX_C]
    const uint32_t pix_off = chs_off + CH_MAX_C * 16u;                             // int [CH_MAX_PX]
    const uint32_t stat_off = pix_off + CH_MAX_PX * 4u;                            // double2 [4 quadrants][32 columns]
    const uint32_t bar_off = stat_off + 4u * 32u * 16u;
    auto full_bar = [&](int s) { return base + bar_off + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return base + bar_off + 64u + 8u * (uint32_t)s; };
    const uint32_t a_full = base + bar_off + 128u, mma_done = a_full + 8u, xbar = a_full + 16u, tmem_slot = a_full + 24u;

    trace_begin(p.trace);
    if (warp == 0 && elect_one()) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(a_full, 1);
        mbar_init(mma_done, 1);
        mbar_init(xbar, (uint32_t)p.CL);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.CL > 1) cluster_sync_all();             // every CTA's cluster barrier is initialised before anyone arrives on it
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
                const uint8_t* wslice = o.w + (size_t)slice * o.ntaps * C * p.ns * 2;
                const int Cx = o.xca + o.xcb;
                const uint8_t* xslice = o.wx ? o.wx + (size_t)slice * Cx * p.ns * 2 : nullptr;
                // stream the (tap, <=128-channel chunk) units of the weight image(s), several units per ring slot
                for (int t = 0; t < p.mtiles; ++t) {
                    for (int img = 0; img < (xslice ? 2 : 1); ++img) {
                        const uint8_t* src = img ? xslice : wslice;
                        const int C_ = img ? Cx : C, ntaps_ = img ? 1 : o.ntaps;
                        const int upt = (C_ + CH_KC - 1) / CH_KC;
                        const int U = ntaps_ * upt;
                        const int upst = max(1, (int)(p.stage_bytes / ((uint32_t)min(C_, CH_KC) * p.ns * 2u)));
                        for (int u0 = 0; u0 < U; u0 += upst, ++u) {
                            uint32_t bytes = 0;
                            for (int uu = u0; uu < min(u0 + upst, U); ++uu) {
                                const int c0 = (uu % upt) * CH_KC;
                                bytes += (uint32_t)min(CH_KC, C_ - c0) * p.ns * 2u;
                            }
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
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.ns >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t b_lbo = (uint32_t)p.ns;                                  // (ns * 16 B) >> 4: between the two planes of a K step
            const uint32_t a_kstep = (2u * plane_bytes) >> 4, b_kstep = 2u * b_lbo;
            for (int oi = 0; oi < p.nops; ++oi) {
                const ChainDev& o = p.ops[oi];
                const int C = o.ca + o.cb;
                for (int t = 0; t < p.mtiles; ++t, ++it) {
                    mbar_wait(a_full, (uint32_t)it & 1u);
                    tc_fence_after();
                    long long wait_full = 0;
                    uint32_t first = 1;
                    // the MMAs of the weight image(s): operand planes [plane0, plane0 + C_/8), all taps or (folded 1x1) the centre
                    for (int img = 0; img < (o.wx ? 2 : 1); ++img) {
                        const int C_ = img ? o.xca + o.xcb : C, ntaps_ = img ? 1 : o.ntaps, plane0 = img ? (C >> 3) : 0;
                        const int upt = (C_ + CH_KC - 1) / CH_KC;
                        const int U = ntaps_ * upt;
                        const int upst = max(1, (int)(p.stage_bytes / ((uint32_t)min(C_, CH_KC) * p.ns * 2u)));
                        for (int u0 = 0; u0 < U; u0 += upst, ++u) {
                            const int s = u % p.stages;
                            const long long w0 = p.dbg ? clock64() : 0;
                            mbar_wait(full_bar(s), ((uint32_t)(u / p.stages)) & 1u);
                            if (p.dbg) wait_full += clock64() - w0;
                            tc_fence_after();
                            uint32_t b_lo = (((base + ring_off + (uint32_t)s * p.stage_bytes) & 0x3FFFFu) >> 4) | (b_lbo << 16);
                            for (int uu = u0; uu < min(u0 + upst, U); ++uu) {
                                const int tap = uu / upt, c0 = (uu - tap * upt) * CH_KC;
                                const int r = ntaps_ == 9 ? tap / 3 : 1;
                                const int sx = ntaps_ == 9 ? tap - 3 * r : 1;
                                const uint32_t a_lo = a_lo0 + (uint32_t)(r * p.Wp + sx) + (uint32_t)(plane0 + (c0 >> 3)) * (plane_bytes >> 4);
                                const int ksteps = min(CH_KC, C_ - c0) >> 4;
                                // all descriptors of the unit are independent adds: the up-to-8 MMAs issue back to back
#pragma unroll
                                for (int k = 0; k < CH_KC / 16; ++k) {
                                    if (k < ksteps)
                                        umma_bf16(tmem_base, ((uint64_t)desc_hi << 32) | (a_lo + (uint32_t)k * a_kstep),
                                                  ((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)k * b_kstep), idesc, (first && k == 0) ? 0u : 1u);
                                }
                                first = 0;
                                b_lo += (uint32_t)ksteps * b_kstep;
                            }
                            umma_commit(empty_bar(s));
                        }
                    }
                    umma_commit(mma_done);
                    if (p.dbg && blockIdx.x == 0 && t == 0) p.dbg[oi * 6 + 5] = wait_full;
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ workers
        const int wt = tid - 64;
        pdl_wait();
        float2* tab = reinterpret_cast<float2*>(gbase + tab_off);
            double2* chs = reinterpret_cast<double2*>(gbase + chs_off);
        int* pixinfo = reinterpret_cast<int*>(gbase + pix_off);
        double2* stat = reinterpret_cast<double2*>(gbase + stat_off);
        const uint32_t plane_bytes = (uint32_t)p.plane_px * 16u;
        int it = 0;
        int pix_tile = -1;                        // tile the pixel table currently describes
        for (int oi = 0; oi < p.nops; ++oi) {
            const ChainDev& o = p.ops[oi];
            const int C = o.ca + o.cb;
            { const int t = 0; CH_STAMP(0); }
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
                for (int cc = wt; cc < C; cc += CH_WORKERS) {       // every channel folds its own group (cpg smem reads)
                    const int g0 = cc / cpg * cpg;
                    double sm = 0.0, sq = 0.0;
                    for (int k = g0; k < g0 + cpg; ++k) { sm += chs[k].x; sq += chs[k].y; }
                    const double mu = sm * inv_cnt;
                    double var = sq * inv_cnt - mu * mu;
                    if (var < 0.0) var = 0.0;
                    const float a = rsqrtf((float)var + 1e-5f) * __ldg(o.gamma + cc);
                    tab[cc] = make_float2(a, __ldg(o.beta + cc) - (float)mu * a);
                }
            }
            if (o.epi.sums_out && wt < 128) stat[wt] = make_double2(0.0, 0.0);
            for (int t = 0; t < p.mtiles; ++t, ++it) {
                // ---- pixel table of this tile: staged position -> pixel index inside the sample, or -1 = zero padding
                const int q_first = t * 128 - p.Wp - 1;
                if (pix_tile != t) {
                    for (int px = wt; px < p.plane_px; px += CH_WORKERS) {
                        const int q = q_first + px;
                        int v = -1;
                        if (q >= 0 && q < p.HpWp) {
                            const int yy = fdiv(q, p.div_wp), xx = q - yy * p.Wp;
                            if (yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W) v = (yy - 1) * p.W + (xx - 1);
                        }
                        pixinfo[px] = v;
                    }
                    pix_tile = t;
                }
                ch_bar_workers();                  // table + pixel table visible
                CH_STAMP(1);
                // ---- stage the A operand of this tile: flat padded positions [128 t - Wp - 1, 128 t + 128 + Wp + 1).
                // Segment 0 = the (normalised) conv input, segment 1 = the raw block input of a folded 1x1 res_conv.
                // Work item = (8-channel plane, pixel); CH_INFLIGHT items (two 16-byte L2 loads each) in flight per thread.
                const size_t pix0 = (size_t)b * p.H * p.W;
                // Coalesced mapping: the channel vector of a pixel is contiguous in NHWC, so the lanes of a warp take
                // consecutive 16-byte chunks (4 fp32 / 8 bf16 channels) of ONE pixel (two pixels per warp when a pixel has <= 16
                // chunks) - one request = whole 128-byte lines instead of 32 separate sectors -; warps stride over the pixels,
                // CH_INFLIGHT pixels (loads) in flight per lane.
                const int wwarp = warp - 2;                                    // 0..13
                for (int seg = 0; seg < (o.wx ? 2 : 1); ++seg) {
                    const void* sa = seg ? (const void*)o.xsrc_a : o.src_a;
                    const void* sb = seg ? (const void*)o.xsrc_b : o.src_b;
                    const int sca = seg ? o.xca : o.ca, scb = seg ? o.xcb : o.cb;
                    const bool snorm = !seg && o.norm, sswish = !seg && o.swish, sb16 = !seg && o.src_b16;
                    const int plane0 = seg ? (C >> 3) : 0;
                    const int esz = sb16 ? 2 : 4, cpi = 16 / esz;               // bytes per element, channels per 16-byte chunk
                    const int nchunk = (sca + scb) / cpi;                       // chunks per pixel
                    const int ppw = nchunk <= 16 ? 2 : 1;                       // pixels per warp iteration
                    const int lsub = ppw == 2 ? (lane & 15) : lane, psub = ppw == 2 ? (lane >> 4) : 0;
                    for (int c0 = lsub; c0 < nchunk; c0 += (ppw == 2 ? 16 : 32)) {
                        const int ch = c0 * cpi;                                // first channel of this lane's chunk
                        const bool in_a = ch < sca;
                        const uint8_t* sbase = reinterpret_cast<const uint8_t*>(in_a ? sa : sb) + (size_t)(in_a ? ch : ch - sca) * esz;
                        const size_t pstride = (size_t)(in_a ? sca : scb) * esz;
                        float4 t0 = make_float4(1.f, 0.f, 1.f, 0.f), t1 = t0;
                        if (snorm) {
                            const float4* tb = reinterpret_cast<const float4*>(tab + ch);
                            t0 = tb[0];
                            t1 = tb[1];
                        }
                        for (int px0 = wwarp * ppw + psub; px0 < p.plane_px; px0 += CH_INFLIGHT * 14 * ppw) {
                            uint4 raw[CH_INFLIGHT];
                            int pi[CH_INFLIGHT];
#pragma unroll
                            for (int e = 0; e < CH_INFLIGHT; ++e) {
                                const int px = px0 + e * 14 * ppw;
                                pi[e] = px < p.plane_px ? pixinfo[px] : -2;
                                if (pi[e] >= 0) raw[e] = __ldcg(reinterpret_cast<const uint4*>(sbase + (pix0 + (size_t)pi[e]) * pstride));
                            }
#pragma unroll
                            for (int e = 0; e < CH_INFLIGHT; ++e) {
                                if (pi[e] == -2) continue;
                                const int px = px0 + e * 14 * ppw;
                                if (sb16) {
                                    const uint4 val = pi[e] >= 0 ? raw[e] : make_uint4(0u, 0u, 0u, 0u);
                                    const uint32_t dst = base + a_off + (uint32_t)(plane0 + c0) * plane_bytes + (uint32_t)px * 16u;
                                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
                                } else {
                                    uint32_t w0 = 0u, w1 = 0u;
                                    if (pi[e] >= 0) {
                                        float x[4] = {__uint_as_float(raw[e].x), __uint_as_float(raw[e].y), __uint_as_float(raw[e].z), __uint_as_float(raw[e].w)};
                                        if (snorm) {
                                            x[0] = fmaf(x[0], t0.x, t0.y); x[1] = fmaf(x[1], t0.z, t0.w);
                                            x[2] = fmaf(x[2], t1.x, t1.y); x[3] = fmaf(x[3], t1.z, t1.w);
                                        }
                                        if (sswish) {
#pragma unroll
                                            for (int j = 0; j < 4; ++j) {       // y * sigmoid(y) = h * tanh(h) + h, h = y / 2
                                                const float h = 0.5f * x[j];
                                                float th;
                                                asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
                                                x[j] = fmaf(h, th, h);
                                            }
                                        }
                                        const __nv_bfloat162 h0 = __floats2bfloat162_rn(x[0], x[1]), h1 = __floats2bfloat162_rn(x[2], x[3]);
                                        w0 = *reinterpret_cast<const uint32_t*>(&h0);
                                        w1 = *reinterpret_cast<const uint32_t*>(&h1);
                                    }
                                    // chunk c0 = channels 4 c0 .. 4 c0 + 3 = half (c0 & 1) of plane c0 / 2
                                    const uint32_t dst = base + a_off + (uint32_t)(plane0 + (c0 >> 1)) * plane_bytes + (uint32_t)px * 16u + (uint32_t)(c0 & 1) * 8u;
                                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(w0), "r"(w1) : "memory");
                                }
                            }
                        }
                    }
                }
                fence_proxy_async();
                ch_bar_workers();
                if (wt == 0) ch_mbar_arrive(a_full);
                CH_STAMP(2);
                // ---- epilogue (warps 2..5: TMEM lane quadrant = warp & 3): this CTA's ns output channels
                if (warp < 6) {
                    const int qd = warp & 3;
                    const int m = qd * 32 + lane;
                    const int q = t * 128 + m;
                    bool valid = q < p.HpWp;
                    int oy = 0, ox = 0;
                    if (valid) {
                        const int yy = fdiv(q, p.div_wp), xx = q - yy * p.Wp;
                        valid = yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W;
                        oy = yy - 1;
                        ox = xx - 1;
                    }
                    float add0[16], add1[16];                  // both chunks' addends are in flight while the MMAs run
                    const bool two = p.ns > 16;
                    if (valid) {
                        ch_addend(o.epi, b, oy, ox, ncol0, add0);
                        if (two) ch_addend(o.epi, b, oy, ox, ncol0 + 16, add1);
                    }
                    mbar_wait(mma_done, (uint32_t)it & 1u);
                    tc_fence_after();
                    CH_STAMP(3);
#pragma unroll
                    for (int ck = 0; ck < 2; ++ck) {
                        if (ck == 1 && !two) break;
                        uint32_t v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(ck * 16), v);
                        float f[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = 0.f;
                        if (valid) tc_epilogue_write(o.epi, v, ck ? add1 : add0, b, oy, ox, ncol0 + ck * 16, f);
                        if (o.epi.sums_out) {
                            float sq[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) sq[j] = f[j] * f[j];
                            const float s1 = tc_colsum16(f, lane), s2 = tc_colsum16(sq, lane);
                            if (!(lane & 1)) {
                                const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                                double2& d = stat[qd * 32 + ck * 16 + col];       // owned by this lane for the whole op
                                d.x += (double)s1;
                                d.y += (double)s2;
                            }
                        }
                    }
                    tc_fence_before();
                    // ---- after the last tile: statistics of this CTA's output channels -> copy 0 of the replicated slots
                    // (plain stores: every channel has one owner CTA)
                    if (o.epi.sums_out && t == p.mtiles - 1) {
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        if (wt < p.ns && ncol0 + wt < o.epi.Cout) {
                            double sm = 0.0, sq = 0.0;
#pragma unroll
                            for (int k = 0; k < 4; ++k) { sm += stat[k * 32 + wt].x; sq += stat[k * 32 + wt].y; }
                            double* dst = o.epi.sums_out + ((size_t)b * o.epi.Cout + ncol0 + wt) * 2;
                            __stcg(reinterpret_cast<double2*>(dst), make_double2(sm, sq));
                        }
                    }
                    __threadfence();
                    CH_STAMP(4);
                }
                ch_bar_workers();          // outputs (and statistics) are fenced; A buffer and TMEM are free
            }
            // ---- cluster barrier: every CTA of the sample has written (and fenced) its channels of this op
            if (p.CL > 1) {
                if (wt == 0) {
                    for (int r = 0; r < p.CL; ++r) cluster_mbar_arrive(xbar, (uint32_t)r);
                }
                cluster_mbar_wait(xbar, (uint32_t)oi & 1u);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    trace_end(p.trace);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 32);
    }
    if (p.CL > 1) cluster_sync_all();             // nobody exits while a peer may still arrive on its barrier
}

// ------------------------------------------------------------------------------------------ host side
// staged positions per tile; the count is made == 1 (mod 8) so that the plane pitch is 16 B off a multiple of 128 B: the
// lanes of a warp write the SAME pixel of consecutive planes and would otherwise all hit the same banks
static int chain_plane_px(int W) { return (130 + 2 * (W + 2) + 7) / 8 * 8 + 1; }

bool chain_level_supported(int H, int W) { return (H + 2) * (W + 2) <= 128 * CHAIN_MAX_MTILES && chain_plane_px(W) <= CH_MAX_PX; }

int chain_slice_rows(int cout) {
    const int npad = (cout + 15) / 16 * 16;
    return npad % 32 == 0 ? 32 : 16;
}

int chain_cluster_size(int cout) { return (cout + 15) / 16 * 16 / chain_slice_rows(cout); }

bool chain_conv_supported(int ca, int cb, int cout, int ks, int H, int W) {
    const int C = ca + cb;
    if (!chain_level_supported(H, W)) return false;
    if (ca <= 0 || ca % 8 || cb % 8 || C % 16 || C > CH_MAX_C) return false;
    const int npad = (cout + 15) / 16 * 16;
    if (cout <= 0 || npad > CH_MAX_C || chain_cluster_size(cout) > 8) return false;
    return ks == 1 || ks == 3;
}

size_t chain_packed_weight_bytes(int cout, int cin, int ks) { return (size_t)ks * ks * cin * ((cout + 15) / 16 * 16) * 2; }

// w_oihw fp32 [cout][cin][ks][ks] -> bf16 [slice of ns output channels][tap][8-channel plane][ns rows][8]
__global__ void pack_chain_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin, int ks,
                                         int npad, int ns) {
    const int ntaps = ks * ks;
    const size_t total = (size_t)ntaps * cin * npad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int j = (int)(r % 8); r /= 8;
        const int row = (int)(r % ns); r /= ns;
        const int plane = (int)(r % (cin / 8)); r /= (cin / 8);
        const int tap = (int)(r % ntaps); r /= ntaps;
        const int n = (int)r * ns + row;
        const int c = plane * 8 + j;
        const float v = n < cout ? w[((size_t)n * cin + c) * ntaps + tap] : 0.f;
        out[i] = __float2bfloat16_rn(v);
    }
}

int chain_pack_conv_weight(const float* w_oihw, uint8_t* packed, int cout, int cin, int ks, cudaStream_t st) {
    DS_REQUIRE(cin % 16 == 0, "chain_pack: cin %d not a multiple of 16", cin);
    const int npad = (cout + 15) / 16 * 16;
    const size_t total = (size_t)ks * ks * cin * npad;
    int blocks = (int)((total + 255) / 256 > 2048 ? 2048 : (total + 255) / 256);
    pack_chain_weight_kernel<<<blocks, 256, 0, st>>>(w_oihw, reinterpret_cast<__nv_bfloat16*>(packed), cout, cin, ks, npad,
                                                     chain_slice_rows(cout));
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
    p->div_px = make_fastdiv((uint32_t)p->plane_px);
    p->div_wp = make_fastdiv((uint32_t)p->Wp);
    p->ns = chain_slice_rows(ops[0].cout);
    p->CL = chain_cluster_size(ops[0].cout);
    int cmax = 0;
    for (int i = 0; i < nops; ++i) {
        const ChainOpDesc& d = ops[i];
        DS_REQUIRE(chain_conv_supported(d.ca, d.cb, d.cout, d.ks, H, W), "chain: op %d has an unsupported shape (%d+%d -> %d, k%d)", i,
                   d.ca, d.cb, d.cout, d.ks);
        DS_REQUIRE(chain_slice_rows(d.cout) == p->ns && chain_cluster_size(d.cout) == p->CL,
                   "chain: op %d (%d output channels) does not split like op 0 (%d slices of %d)", i, d.cout, p->CL, p->ns);
        DS_REQUIRE(!d.norm || (d.G > 0 && d.G <= CH_MAX_GROUPS && (d.ca + d.cb) % d.G == 0 && d.sums_a && (d.cb == 0 || d.sums_b)),
                   "chain: op %d has an invalid GroupNorm (G=%d)", i, d.G);
        DS_REQUIRE(!(d.src_b16 && (d.norm || d.swish)), "chain: op %d normalises a bf16 source", i);
        ChainDev& o = p->ops[i];
        o.src_a = d.src_a; o.src_b = d.src_b; o.ca = d.ca; o.cb = d.cb;
        o.sums_a = d.sums_a; o.sums_b = d.sums_b; o.gamma = d.gamma; o.beta = d.beta;
        o.G = d.G > 0 ? d.G : 1; o.swish = d.swish; o.norm = d.norm; o.src_b16 = d.src_b16;
        o.w = d.w; o.ntaps = d.ks * d.ks;
        if (d.wx) {
            const int Cx = d.xca + d.xcb;
            DS_REQUIRE(d.xsrc_a && d.xca > 0 && d.xca % 8 == 0 && d.xcb % 8 == 0 && Cx % 16 == 0 && (d.xcb == 0 || d.xsrc_b) &&
                           d.ca + d.cb + Cx <= CH_MAX_C + 256,
                       "chain: op %d has an invalid folded 1x1 operand (%d+%d channels)", i, d.xca, d.xcb);
            o.xsrc_a = d.xsrc_a; o.xsrc_b = d.xsrc_b; o.xca = d.xca; o.xcb = d.xcb; o.wx = d.wx;
            o.epi.bias2 = d.bias2;
            cmax = cmax > d.ca + d.cb + Cx ? cmax : d.ca + d.cb + Cx;
        }
        o.epi.bias = d.epi.bias; o.epi.temb = d.epi.temb; o.epi.temb_off = d.epi.temb_off; o.epi.temb_stride = d.epi.temb_stride;
        o.epi.temb_bcast = d.epi.temb_bcast; o.epi.residual = d.epi.residual;
        o.epi.out_f32 = d.out_f32; o.epi.out_b16 = reinterpret_cast<__nv_bfloat16*>(d.out_b16); o.epi.out_nchw = nullptr;
        o.epi.sums_out = d.sums_out; o.epi.sums_B = B;
        o.epi.Cout = d.cout; o.epi.Ho = H; o.epi.Wo = W;
        cmax = cmax > d.ca + d.cb ? cmax : d.ca + d.cb;
    }
    p->a_bytes = (uint32_t)align_up((size_t)(cmax / 8) * p->plane_px * 16, 1024);
    p->stage_bytes = (uint32_t)align_up((size_t)p->ns * CH_KC * 2 * CH_UNITS_PER_STAGE, 1024);
    const size_t fixed = p->a_bytes + CH_MAX_C * 8 + CH_MAX_GROUPS * 8 + CH_MAX_C * 16 + CH_MAX_PX * 4 + 4 * 32 * 16 + 192 + 1024 + 64;
    DS_REQUIRE(fixed + 2 * (size_t)p->stage_bytes <= CH_SMEM_LIMIT, "chain: shared memory (A %u B + 2 x %u B)", p->a_bytes, p->stage_bytes);
    int stages = (int)((CH_SMEM_LIMIT - fixed) / p->stage_bytes);
    p->stages = stages > 8 ? 8 : stages;
    plan->smem_bytes = (int)(fixed + (size_t)p->stages * p->stage_bytes);
    plan->B = B;
    plan->CL = p->CL;
    return DS_OK;
}

static long long* g_chain_dbg = nullptr;
static int g_chain_dbg_nops = 0;

int chain_launch(const ChainPlan* plan, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        DS_CHECK_CUDA(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024));
        attr_set = true;
    }
    ChainParams p = *reinterpret_cast<const ChainParams*>(plan->params);
    p.trace = trace_next(7);
    static const bool chain_dbg = getenv("DIFFSPLIT_B200_CHAIN_DBG") != nullptr;
    if (chain_dbg) {
        if (!g_chain_dbg) DS_CHECK_CUDA(cudaMalloc(&g_chain_dbg, CHAIN_MAX_OPS * 6 * sizeof(long long)));
        DS_CHECK_CUDA(cudaMemsetAsync(g_chain_dbg, 0, CHAIN_MAX_OPS * 6 * sizeof(long long), st));
        p.dbg = g_chain_dbg;
        g_chain_dbg_nops = p.nops;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(plan->B * plan->CL));
    cfg.blockDim = dim3(CH_THREADS);
    cfg.dynamicSmemBytes = (size_t)plan->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)plan->CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
    if (pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    DS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel, p));
    return DS_OK;
}

}  // namespace ds

// debugging aid (DIFFSPLIT_B200_CHAIN_DBG=1): clock64 stamps of CTA 0 for every op of the LAST chain launch:
// [op][0 start, 1 table built, 2 operand staged, 3 MMAs done, 4 epilogue done, 5 = cycles the MMA issuer waited for weight slabs]
extern "C" int ds_debug_chain_phases(long long* out, int max_ops, int* n_ops) {
    using namespace ds;
    DS_REQUIRE(g_chain_dbg && out && n_ops, "chain debug buffer not active");
    DS_CHECK_CUDA(cudaDeviceSynchronize());
    const int n = g_chain_dbg_nops < max_ops ? g_chain_dbg_nops : max_ops;
    DS_CHECK_CUDA(cudaMemcpy(out, g_chain_dbg, (size_t)n * 6 * sizeof(long long), cudaMemcpyDeviceToHost));
    *n_ops = n;
    return DS_OK;
}
