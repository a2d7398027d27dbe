// Fused  GroupNorm-apply + Swish -> 3x3 / 1x1 convolution  on the tensor cores, operands staged through registers.
//
// Why a second tensor-core conv kernel: for the 16..224-channel layers of the splitting UNets the TMA im2col path
// (tc.cu) fetches every input pixel 9 times in 32..128-byte rows and needs a separate normalisation pass before it.
// Here each CTA
//   1. loads the raw fp32 input pixels of its tile + halo ONCE (16-byte loads), applies the per-(sample, channel)
//      GroupNorm scale/shift + Swish of the reference Block (model/sr3_modules/unet.py:80-91) in registers, converts to
//      bf16 and writes the UMMA no-swizzle K-major layout [8-channel plane][pixel][16 B];
//   2. issues all 9 taps x C/16 tcgen05.mma from that one copy: in the ZERO-PADDED, row-major flattened image
//      (pitch W+2) - or, for wide images, in a 9 x 18 patch around a tile of 7 x 16 outputs (pitch 18) - a filter tap is
//      a constant shift of the flat index, i.e. just a different descriptor start address;
//   3. runs the common epilogue (bias, conditioning vector, fp32 residual, fp32 / bf16 stores, GroupNorm statistics of
//      its own output for the consumer).
// Flat tiles = 128 consecutive padded positions (border positions are computed and discarded: 6 % waste at 64^2).
// The channel concat of the up path (unet.py:255) is two source pointers; the weights arrive with one 3-D TMA load.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <vector>

#include "tc.cuh"
#include "tc_ptx.cuh"

namespace ds {

constexpr int HALO_THREADS = 192;        // base block: TMA/setup warp, MMA warp, 4 epilogue warps (all of them stage)
constexpr int HALO_MAX_THREADS = 512;    // wide block for layers that fit one CTA per SM anyway: extra warps only stage
constexpr int HALO_SEG_PX = 136;           // 128 outputs + 2 halo pixels, padded to a multiple of 8
constexpr int HALO_TW = 16, HALO_TH = 7, HALO_PW = HALO_TW + 2;      // 2-D tiles: 7 x 16 outputs, patch 9 x 18
constexpr int HALO_2D_PX = ((HALO_TH + 2) * HALO_PW + 2 + 7) / 8 * 8;   // 168 staged positions
constexpr int HALO_MAX_SAMPLES = 12;
constexpr int HALO_MAX_GROUPS = 64;
constexpr size_t HALO_SMEM_LIMIT = 200 * 1024;

struct HaloParams {
    CUtensorMap wmap;                           // packed weights as a 3-D tensor (128 bf16 = one 16-row x 8-channel plane
                                                // of a block, 16-channel block, (tap, kstep, plane)): one TMA load lands the
                                                // CTA's BN output channels as [tap][kstep][plane][BN rows][16 B]
    const float* src_a; const float* src_b;     // fp32 NHWC [B,H,W,ca|cb]
    int ca, cb;
    const float2* stats;                        // (mean, rstd) [B][G] of the concat input, or
    const double* sums_a; const double* sums_b; // per-channel fp64 (sum, sumsq) [B][ca|cb][2]; all null = identity
    const float* gamma; const float* beta;
    int G, swish;
    const uint8_t* w;                           // bf16 [16-channel block][tap][kstep][plane(2)][16 rows][8]
    TcEpi epi;
    int B, H, W, Wp, HpWp, total_q;
    int C, ksteps, ntaps, BN, n_tiles, Npad;
    int tf32;                                   // operands staged / packed as fp32 (TF32 MMA): a 16-byte plane row holds 4 channels
                                                // and a K step is 8 channels (two planes), instead of 8 and 16 for bf16
    int wloads;                                 // the weights arrive in `wloads` TMA boxes of (taps / wloads) taps each
    // operand buffer geometry, flat tiles: the 128 outputs and their three filter rows are ONE contiguous run of the flat
    // padded index space (130 + 2 (W + 2) positions, W + 2 <= 139; wider images use the 2-D tiles below)
    int plane_px, seg_stride_px;
    // 2-D tile geometry (tile2d = 1): a tile is HALO_TH rows x HALO_TW columns of ONE sample; the staged patch has
    // (HALO_TH + 2) rows of pitch HALO_PW = HALO_TW + 2, so a tap is still a constant shift (r * HALO_PW + s) and the 128 MMA
    // rows are patch positions 0..127 (columns HALO_TW, HALO_TW + 1 of every row are computed and discarded).  Stages
    // 1.5 pixels per output instead of 1.2 + 2 (W + 2) / 128 (flat tiles) or 3.2 (wide images).
    int tile2d, tiles_x, tiles_y;
    FastDiv div_tiles_x, div_tiles_xy;
    // A operand layout: no-swizzle [8-channel plane][pixel][16 B].  (The swizzled layouts [64-channel chunk][pixel][128 B]
    // also work for the shifted windows - the tensor core's swizzle is a pure function of the shared-memory address, the
    // descriptor base offset stays 0 - but were slower here: profiles/README.md, findings 3 and 7; conv_tcp_kernel in tc.cu
    // uses them because its patches come from TMA.)
    FastDiv div_hpwp, div_wp, div_planepx;
    long long* dbg;                             // optional per-CTA phase timestamps (clock64), 8 per CTA
    TraceSlot trace;
};
#define HALO_STAMP(i) do { if (p.dbg && tid == (i >= 5 ? 64 : 0)) p.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (i)] = clock64(); } while (0)

// no-swizzle K-major descriptor: rows 16 B apart inside an 8-row core matrix, SBO between 8-row groups,
// LBO between the two 8-element K halves of one K=16 step
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

template <int MAXT, bool TILE2D, bool TF32>
__global__ void __launch_bounds__(MAXT, (MAXT <= 192 ? 4 : 1)) conv_halo_kernel(const __grid_constant__ HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
    constexpr int CPP = TF32 ? 4 : 8;               // channels per 16-byte plane row
    const int P = p.C / CPP;
    const uint32_t plane_bytes = (uint32_t)p.plane_px * 16u;
    const uint32_t a_bytes_total = P * plane_bytes;
    const uint32_t a_off = 0, b_off = a_bytes_total;
    // weights of one 16-output-channel block: [tap][kstep][plane(2)][16 rows][16 B]; a CTA owns BN/16 consecutive blocks
    const uint32_t blk_bytes = (uint32_t)p.ntaps * p.ksteps * 512u;
    const uint32_t b_bytes = (uint32_t)(p.BN >> 4) * blk_bytes;
    const uint32_t tab_off = b_off + b_bytes;
    const int q0 = blockIdx.x * 128;
    const int nt = blockIdx.y;
    // samples touched by the loaded span [q0 - Wp - 1, q0 + 128 + Wp]
    int b_first = (q0 - p.Wp - 1) / p.HpWp;
    if (q0 - p.Wp - 1 < 0) b_first = 0;
    int b_last = (q0 + 128 + p.Wp) / p.HpWp;
    if (b_last > p.B - 1) b_last = p.B - 1;
    int tile_x0 = 0, tile_y0 = 0;                   // 2-D tiles: first output column / row of this tile
    if (TILE2D) {
        const int tb = fdiv((int)blockIdx.x, p.div_tiles_xy);
        const int rem = (int)blockIdx.x - tb * p.tiles_x * p.tiles_y;
        const int ty = fdiv(rem, p.div_tiles_x);
        tile_y0 = ty * HALO_TH;
        tile_x0 = (rem - ty * p.tiles_x) * HALO_TW;
        b_first = b_last = tb;
    }
    const int nsamp = b_last - b_first + 1;
    const uint32_t gst_off = tab_off + (uint32_t)nsamp * p.C * 8u;                 // (mean, rstd) per (sample, group)
    const uint32_t chs_off = (gst_off + (uint32_t)nsamp * HALO_MAX_GROUPS * 8u + 15u) & ~15u;   // fp64 (sum, sq) per (sample, ch)
    const uint32_t bar_off = (chs_off + (uint32_t)nsamp * p.C * 16u + 15u) & ~15u;
    const uint32_t bfull = base + bar_off, mma_done = bfull + 8u, tmem_slot = bfull + 16u;
    uint8_t* red = gbase + bar_off + 32u;                                          // epilogue reduction buffer
    const uint32_t tmem_cols = (uint32_t)p.BN <= 32u ? 32u : ((uint32_t)p.BN <= 64u ? 64u : 128u);

    HALO_STAMP(0);
    trace_begin(p.trace);
    if (warp == 0 && elect_one()) {
        mbar_init(bfull, 1);
        mbar_init(mma_done, 1);
        fence_barrier_init();
        // ---- weights: ONE bulk copy (constant data: may run before the predecessor kernel has finished)
        mbar_expect_tx(bfull, b_bytes);
        const int per = p.ntaps * p.ksteps * 2 / p.wloads;           // (tap, kstep, plane) images per box
        for (int l = 0; l < p.wloads; ++l)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                    base + b_off + (uint32_t)l * (uint32_t)per * (uint32_t)p.BN * 16u),
                "l"(reinterpret_cast<uint64_t>(&p.wmap)), "r"(bfull), "r"(0), "r"(nt * (p.BN >> 4)), "r"(l * per)
                : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    HALO_STAMP(1);
    pdl_wait();          // everything above (barriers, TMEM, weight fetch) overlapped the previous kernel's tail
    pdl_trigger();
    HALO_STAMP(2);

    // ---- operand staging: raw fp32 -> normalise -> Swish -> bf16 -> [8-channel plane][pixel][16 B].
    // One thread = one pixel of the staged run (index decode once per pixel), looping over the channel planes with two
    // planes (4 x 16-byte loads) in flight.  The first loads are issued BEFORE the scale/shift table is built, so the two
    // global-memory round trips (statistics / affine parameters and pixels) overlap.
    float2* tab = reinterpret_cast<float2*>(gbase + tab_off);
    const int q_first = p.ntaps == 9 ? q0 - p.Wp - 1 : q0 - 1;
    struct Pix { const float* a; const float* b; int tabi; uint32_t dst; int pxi; };   // a == nullptr: zero padding
    auto decode = [&](int px, Pix& px_) {
        px_.a = nullptr;
        px_.b = nullptr;
        px_.tabi = 0;
        px_.dst = base + a_off + px * 16;
        px_.pxi = px;
        if (TILE2D) {
            const int pr = px / HALO_PW, pc = px - pr * HALO_PW;
            const int yy = tile_y0 + pr - 1, xx = tile_x0 + pc - 1;
            if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
                const size_t pix = ((size_t)b_first * p.H + yy) * p.W + xx;
                px_.a = p.src_a + pix * p.ca;
                px_.b = p.src_b ? p.src_b + pix * p.cb : nullptr;
            }
            return;
        }
        const int q = q_first + px;
        if (q >= 0 && q < p.total_q) {
            const int b = fdiv(q, p.div_hpwp);
            const int rq = q - b * p.HpWp;
            const int yy = fdiv(rq, p.div_wp), xx = rq - yy * p.Wp;
            if (yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W) {
                const size_t pix = ((size_t)b * p.H + (yy - 1)) * p.W + (xx - 1);
                px_.a = p.src_a + pix * p.ca;
                px_.b = p.src_b ? p.src_b + pix * p.cb : nullptr;
                px_.tabi = (b - b_first) * p.C;
            }
        }
    };
    // bf16: a plane = 8 channels = two 16-byte loads (v0, v1);  tf32: a plane = 4 channels = one load (v0)
    auto load_plane = [&](const Pix& px_, int kp, float4& v0, float4& v1) {
        if (!px_.a) return;
        const int c0 = kp * CPP;
        const float* src = c0 < p.ca ? px_.a + c0 : px_.b + (c0 - p.ca);
        v0 = __ldg(reinterpret_cast<const float4*>(src));
        if (!TF32) v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
    };
    auto store_plane = [&](const Pix& px_, int kp, const float4& v0, const float4& v1) {
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (px_.a) {
            const float4* tb = reinterpret_cast<const float4*>(tab + px_.tabi + kp * CPP);   // (a, sh) pairs of 2 channels
            float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int j = 0; j < CPP / 2; ++j) {
                const float4 sc = tb[j];
                x[2 * j] = fmaf(x[2 * j], sc.x, sc.y);
                x[2 * j + 1] = fmaf(x[2 * j + 1], sc.z, sc.w);
            }
            if (p.swish) {
#pragma unroll
                for (int j = 0; j < CPP; ++j) {
                    if (TF32) {
                        // y * sigmoid(y) with the exponential and the division at full fp32 accuracy of the fast intrinsics
                        // (relative error ~1e-6 << the 2^-11 of the TF32 rounding that follows; tanh.approx is only ~2^-11)
                        x[j] = __fdividef(x[j], 1.0f + __expf(-x[j]));
                    } else {                            // y * sigmoid(y) = h * tanh(h) + h, h = y / 2: one MUFU op
                        const float h = 0.5f * x[j];
                        float th;
                        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
                        x[j] = fmaf(h, th, h);
                    }
                }
            }
            if (TF32) {
                val = make_uint4(__float_as_uint(to_tf32(x[0])), __float_as_uint(to_tf32(x[1])), __float_as_uint(to_tf32(x[2])),
                                 __float_as_uint(to_tf32(x[3])));
            } else {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
                    w[j] = *reinterpret_cast<const uint32_t*>(&h);
                }
                val = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        const uint32_t dst = px_.dst + kp * plane_bytes;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
    };

    // work item = (pixel, pair of channel planes): decode once, four 16-byte loads; two items in flight per thread
    const int npairs = P >> 1;
    const int items = npairs * p.plane_px;
    struct Item { Pix px; int kp; float4 v[4]; bool on; };
    auto issue = [&](int idx, Item& it) {
        it.on = idx < items;
        if (!it.on) return;
        const int pp = fdiv(idx, p.div_planepx);
        decode(idx - pp * p.plane_px, it.px);
        it.kp = 2 * pp;
        load_plane(it.px, it.kp, it.v[0], it.v[1]);
        load_plane(it.px, it.kp + 1, it.v[2], it.v[3]);
    };
    auto finish = [&](const Item& it) {
        if (!it.on) return;
        store_plane(it.px, it.kp, it.v[0], it.v[1]);
        store_plane(it.px, it.kp + 1, it.v[2], it.v[3]);
    };
    Item it0, it1;
    issue(tid, it0);
    issue(tid + nthr, it1);

    // ---- per (sample in tile, channel) scale / shift of the fused GroupNorm: a = rstd * gamma, sh = beta - mean * a
    {
        const bool norm = p.stats || p.sums_a;
        const int cpg = norm ? p.C / p.G : 1;
        if (p.sums_a) {
            // fold the producers' replicated per-channel fp64 sums; a group may straddle the two sources of a concat.
            // Stage 1: one thread per (sample, channel) adds the replicated copies (all loads independent: ONE round trip).
            double2* chs = reinterpret_cast<double2*>(gbase + chs_off);
            for (int i = tid; i < nsamp * p.C; i += nthr) {
                const int s = i / p.C, cc = i - s * p.C, b = b_first + s;
                const bool first = cc < p.ca;
                const double2* src = reinterpret_cast<const double2*>(
                    first ? p.sums_a + ((size_t)b * p.ca + cc) * 2 : p.sums_b + ((size_t)b * p.cb + (cc - p.ca)) * 2);
                const size_t cstride = (size_t)p.B * (first ? p.ca : p.cb);
                double2 v[TC_SUM_COPIES];
#pragma unroll
                for (int k = 0; k < TC_SUM_COPIES; ++k) v[k] = src[k * cstride];
                double sm = 0.0, sq = 0.0;
#pragma unroll
                for (int k = 0; k < TC_SUM_COPIES; ++k) { sm += v[k].x; sq += v[k].y; }
                chs[i] = make_double2(sm, sq);
            }
            __syncthreads();
        }
        const double inv_cnt = norm ? 1.0 / ((double)p.H * p.W * cpg) : 0.0;
        for (int i = tid; i < nsamp * p.C; i += nthr) {
            const int s = i / p.C, c = i - s * p.C;
            float a = 1.f, sh = 0.f;
            if (p.sums_a) {                       // every channel folds its own group (cpg reads of shared memory)
                const double2* cs = reinterpret_cast<const double2*>(gbase + chs_off) + s * p.C + c / cpg * cpg;
                double sm = 0.0, sq = 0.0;
                for (int k = 0; k < cpg; ++k) { sm += cs[k].x; sq += cs[k].y; }
                const double mu = sm * inv_cnt;
                double var = sq * inv_cnt - mu * mu;
                if (var < 0.0) var = 0.0;
                a = rsqrtf((float)var + 1e-5f) * p.gamma[c];
                sh = p.beta[c] - (float)mu * a;
            } else if (norm) {
                const float2 st = p.stats[(size_t)(b_first + s) * p.G + c / cpg];
                a = st.y * p.gamma[c];
                sh = p.beta[c] - st.x * a;
            }
            tab[i] = make_float2(a, sh);
        }
    }
    __syncthreads();
    HALO_STAMP(3);

    finish(it0);
    finish(it1);
    for (int i0 = tid + 2 * nthr; i0 < items; i0 += 2 * nthr) {
        issue(i0, it0);
        issue(i0 + nthr, it1);
        finish(it0);
        finish(it1);
    }
    fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    HALO_STAMP(4);

    if (warp == 1) {
        if (elect_one()) {
            mbar_wait(bfull, 0);
            tc_fence_after();
            // Descriptors differ only in the 14-bit start-address field: 32-bit adds.  B image of (tap, kstep): two planes of
            // BN rows x 16 B (LBO = BN * 16), 8-row groups 128 B apart.
            const uint32_t idesc = umma_idesc(p.BN, TF32);
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);                       // SBO = 128 B, descriptor version 1
            const uint32_t a_lo0 = (((base + a_off) & 0x3FFFFu) >> 4) | ((plane_bytes >> 4) << 16);
            const uint32_t b_lo0 = (((base + b_off) & 0x3FFFFu) >> 4) | ((uint32_t)p.BN << 16);      // LBO = BN * 16 B
            const uint32_t a_kstep = (2u * plane_bytes) >> 4;
            const uint32_t b_kstep = (uint32_t)p.BN * 2u;                                              // 2 planes of BN * 16 B
            uint32_t b_lo = b_lo0;
            for (int tap = 0; tap < p.ntaps; ++tap) {
                const int r = p.ntaps == 9 ? tap / 3 : (TILE2D ? 1 : 0);
                const int sx = p.ntaps == 9 ? tap - 3 * r : 1;
                uint32_t a_lo = a_lo0 + (uint32_t)(r * p.seg_stride_px + sx);
                for (int kk = 0; kk < p.ksteps; ++kk) {
                    if (TF32) umma_tf32(tmem_base, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, (tap | kk) ? 1u : 0u);
                    else umma_bf16(tmem_base, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, (tap | kk) ? 1u : 0u);
                    a_lo += a_kstep;
                    b_lo += b_kstep;
                }
            }
            umma_commit(mma_done);
        }
        __syncwarp();
    } else if (warp >= 2 && warp < 6) {
        const int qd = warp & 3;
        const int m = qd * 32 + lane;
        const int q = q0 + m;
        bool valid = q < p.total_q;
        int b = 0, oy = 0, ox = 0;
        if (TILE2D) {
            const int pr = m / HALO_PW, pc = m - pr * HALO_PW;
            b = b_first;
            oy = tile_y0 + pr;
            ox = tile_x0 + pc;
            valid = pr < HALO_TH && pc < HALO_TW && oy < p.H && ox < p.W;
        } else if (valid) {
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
        HALO_STAMP(5);
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, v);
            float f[16];
            if (valid) {
                if (c0) tc_epilogue_addend(p.epi, b, oy, ox, nt * p.BN + c0, add);
                tc_epilogue_write(p.epi, v, add, b, oy, ox, nt * p.BN + c0, f);
            }
            if (p.epi.sums_out) {
                const int bt0 = TILE2D ? b_first : fdiv(q0, p.div_hpwp);
                const int nsr = TILE2D ? 1 : fdiv(min(q0 + 127, p.total_q - 1), p.div_hpwp) - bt0 + 1;
                if (nsr == 1) tc_epilogue_stats_smem(p.epi, f, valid, bt0, nt * p.BN + c0, tid - 64, (int)(blockIdx.x % TC_SUM_COPIES), red);
                else if (nsr == 2) tc_epilogue_stats_shfl(p.epi, f, valid, b, nt * p.BN + c0, tid - 64, bt0, nsr, (int)(blockIdx.x % TC_SUM_COPIES), red);
                else tc_epilogue_stats(p.epi, f, valid, b, nt * p.BN + c0, m, tid - 64, bt0, (int)(blockIdx.x % TC_SUM_COPIES), red);
            }
        }
        HALO_STAMP(6);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
    HALO_STAMP(7);
    trace_end(p.trace);
}


static long long* g_halo_dbg = nullptr;
static size_t g_halo_dbg_ctas = 0;

// ------------------------------------------------------------------------------------------ host side
// 2-D tiles for wide images (the flat tiles stage 130 + 2 (W + 2) positions for 128 outputs, or three separate segments of
// 136; the 2-D ones 168 for 112) - always from DIFFSPLIT_B200_HALO_2D_MINW (138) columns on, and from 48 columns on when
// the layer is many waves of tiles (throughput regime; in the latency regime of the 64 x 64 benchmark layers the 14 % extra
// tiles cost more than the smaller patches save: 2017 vs 2097 steps/s), from 96 columns on already at 5 waves
static bool halo_use_2d(int B, int H, int W) {
    static int minw = -1;
    if (minw < 0) { const char* e = getenv("DIFFSPLIT_B200_HALO_2D_MINW"); minw = e ? atoi(e) : 138; }
    if (W >= minw || 130 + 2 * (W + 2) > 3 * HALO_SEG_PX) return true;       // the flat run is capped at 408 staged positions
    const int64_t flat_tiles = ((int64_t)B * (H + 2) * (W + 2) + 127) / 128;
    // W >= 96: the flat tiles stage >= 2.5x the outputs, and 5 waves are enough for the 2-D tiles to win
    // (8 x 64 x 128 x 128, 64 -> 64: 125 -> 79 us; hagen_joint_512: 5.29 -> 5.09 ms per step)
    return minw == 138 && ((W >= 48 && flat_tiles >= 8 * 148) || (W >= 96 && flat_tiles >= 5 * 148));
}
static int64_t halo_m_tiles(int B, int H, int W) {
    if (halo_use_2d(B, H, W)) return (int64_t)B * ((H + HALO_TH - 1) / HALO_TH) * ((W + HALO_TW - 1) / HALO_TW);
    return ((int64_t)B * (H + 2) * (W + 2) + 127) / 128;
}
static int halo_plane_px(int ntaps, int W, bool two_d) {
    if (two_d) return HALO_2D_PX;
    if (ntaps == 1) return HALO_SEG_PX;
    return (130 + 2 * (W + 2) + 7) / 8 * 8;
}

// es = bytes per staged operand element: 2 (bf16) or 4 (tf32)
static size_t halo_a_bytes(int C, int ntaps, int W, bool two_d, int es) { return (size_t)(C * es / 16) * halo_plane_px(ntaps, W, two_d) * 16; }

static size_t halo_smem_bytes(int C, int ntaps, int W, int BN, int nsamp, bool two_d, int es) {
    return halo_a_bytes(C, ntaps, W, two_d, es) + 1024 + (size_t)ntaps * C * BN * es + (size_t)nsamp * C * 8 +
           (size_t)nsamp * HALO_MAX_GROUPS * 8 + (size_t)nsamp * C * 16 + 96 + TC_RED_BYTES + 128;
}

static int halo_samples_per_tile(int H, int W, bool two_d) {
    if (two_d) return 1;
    const int Wp = W + 2, HpWp = (H + 2) * Wp;
    return (128 + 2 * Wp + 2) / HpWp + 2;
}

static int halo_pick_bn(int cout, int C, int ntaps, int W, int nsamp, int64_t m_tiles, bool two_d, int es) {
    const int npad = (cout + 15) / 16 * 16;
    int bn = 16;
    for (int c = 128; c >= 16; c >>= 1)
        if (npad % c == 0) { bn = c; break; }
    // every N tile re-stages (and re-normalises) the same operand pixels: only split N for shared memory or when the
    // grid would leave most SMs idle
    static int min_ctas = -1;
    if (min_ctas < 0) { const char* e = getenv("DIFFSPLIT_B200_HALO_MIN_CTAS"); min_ctas = e ? atoi(e) : 48; }
    while (bn > 16 && (halo_smem_bytes(C, ntaps, W, bn, nsamp, two_d, es) > HALO_SMEM_LIMIT || m_tiles * (npad / bn) < min_ctas)) bn >>= 1;
    return bn;
}

// images of one (tap, kstep, plane) the weight TMA box may hold (box dimension limit 256): the taps are spread over 1, 3 or 9 boxes
static int halo_weight_loads(int C, int ntaps, int es) {
    const int per_tap = C * es / 16;                      // (kstep, plane) images per tap
    for (int l = 1; l <= ntaps; ++l)
        if (ntaps % l == 0 && (ntaps / l) * per_tap <= 256) return l;
    return 0;
}

bool halo_conv_supported(int ca, int cb, int cout, int ks, int B, int H, int W, int tf32) {
    const int C = ca + cb, es = tf32 ? 4 : 2;
    if (ca <= 0 || ca % 8 || cb % 8 || C % 16 || C > 224 || halo_weight_loads(C, ks * ks, es) == 0) return false;
    if (!(ks == 1 || ks == 3)) return false;
    if ((int64_t)B * (H + 2) * (W + 2) >= (1ll << 31) - 4096) return false;
    const bool two_d = halo_use_2d(B, H, W);
    const int nsamp = halo_samples_per_tile(H, W, two_d);
    if (nsamp > HALO_MAX_SAMPLES) return false;
    return halo_smem_bytes(C, ks * ks, W, 16, nsamp, two_d, es) <= HALO_SMEM_LIMIT;
}

// Is the fused kernel also the FASTER choice (vs GroupNorm-apply into a bf16 tensor + the TMA-fed conv)?  Each CTA stages
// plane_px pixels for 128 outputs (1.2x at 8 x 8 ... 3.2x for wide images, whose three filter rows are separate segments)
// and every N tile repeats that work.  For grids of at most a few waves the saved launch and HBM round trip win regardless;
// for many waves, measured on B200 (8 x C x S x S, us, fused vs unfused): C=16 S=512 190 vs 249, C=32 S=512 458 vs 550,
// C=64 S=256 257 vs 344, C=64 S=512 ~1850 vs 1360 (the re-reads spill from L2 to HBM), C=128 S=128 570 vs 166,
// C=128 S=256 2260 vs 645.
bool halo_conv_preferred(int ca, int cb, int cout, int ks, int B, int H, int W, int tf32) {
    // throughput regime: GroupNorm-apply + the persistent TMA-fed conv (conv_tcs_kernel) instead of a fused kernel for inputs of at
    // least DIFFSPLIT_B200_UNFUSE_MINC channels (bf16 operands only: the TF32 nets are the narrow latency-bound ones)
    static int unfuse_minc = -1;
    if (unfuse_minc < 0) { const char* e = getenv("DIFFSPLIT_B200_UNFUSE_MINC"); unfuse_minc = e ? atoi(e) : 64; }
    static int unfuse_minc32 = -1;
    if (unfuse_minc32 < 0) { const char* e = getenv("DIFFSPLIT_B200_UNFUSE_MINC_TF32"); unfuse_minc32 = e ? atoi(e) : 64; }
    if (ks == 3 && ca + cb >= (tf32 ? unfuse_minc32 : unfuse_minc) && cout >= 32 && (int64_t)B * H * W >= (int64_t)(tf32 ? 2 : 8) * 148 * 128)
        return false;
    if (ca + cb <= 224 && stream_conv_preferred(ca, cb, cout, ks, B, H, W, tf32)) return true;      // pipelined variant (tc_stream.cu)
    if (!halo_conv_supported(ca, cb, cout, ks, B, H, W, tf32)) return false;
    const int C = ca + cb, ntaps = ks * ks, es = tf32 ? 4 : 2;
    const int64_t m_tiles = halo_m_tiles(B, H, W);
    const bool two_d = halo_use_2d(B, H, W);
    const int nsamp = halo_samples_per_tile(H, W, two_d);
    const int bn = halo_pick_bn(cout, C, ntaps, W, nsamp, m_tiles, two_d, es);
    const int n_tiles = (cout + 15) / 16 * 16 / bn;
    if (m_tiles * n_tiles <= 4 * 148) return true;                           // latency regime
    // throughput regime, wide inputs the pipelined variant cannot take: GroupNorm-apply + the persistent TMA-fed conv
    // (conv_tcs_kernel) beats this kernel's serial load -> normalise -> MMA -> epilogue (128 -> 64 at 8 x 512^2: 2.0 ms fused)
    if (C >= 128 && m_tiles * n_tiles >= 8 * 148) return false;
    const double redo = (double)halo_plane_px(ntaps, W, two_d) / (two_d ? (double)(HALO_TH * HALO_TW) : 128.0);   // staged per output pixel
    if (redo * n_tiles * C > 320.0) return false;
    const double in_bytes = (double)B * H * W * C * 4.0;
    if (redo > 1.5 && in_bytes > 400e6) return false;                        // 3x re-reads of a tensor that does not fit L2
    return true;
}

size_t halo_packed_weight_bytes(int cout, int cin, int ks, int tf32) {
    return (size_t)ks * ks * cin * ((cout + 15) / 16 * 16) * (tf32 ? 4 : 2);
}

// [16-channel block][tap][kstep][plane(2)][16 rows][CPP channels]; CPP = 8 (bf16) or 4 (tf32, values rounded to TF32)
template <bool TF32>
__global__ void pack_halo_weight_kernel(const float* __restrict__ w, void* __restrict__ out, int cout, int cin, int ks, int npad) {
    constexpr int CPP = TF32 ? 4 : 8;
    const int ntaps = ks * ks;
    const size_t total = (size_t)ntaps * cin * npad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int j = (int)(r % CPP); r /= CPP;
        const int row = (int)(r % 16); r /= 16;
        const int plane = (int)(r % 2); r /= 2;
        const int kk = (int)(r % (cin / (2 * CPP))); r /= (cin / (2 * CPP));
        const int tap = (int)(r % ntaps); r /= ntaps;
        const int n = (int)r * 16 + row;
        const int c = kk * 2 * CPP + plane * CPP + j;
        const float v = n < cout ? w[((size_t)n * cin + c) * ntaps + tap] : 0.f;
        if (TF32) reinterpret_cast<float*>(out)[i] = to_tf32(v);
        else reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
    }
}

int halo_pack_conv_weight(const float* w_oihw, uint8_t* packed, int cout, int cin, int ks, int tf32, cudaStream_t st) {
    DS_REQUIRE(cin % 16 == 0, "halo_pack: cin %d not a multiple of 16", cin);
    const int npad = (cout + 15) / 16 * 16;
    const size_t total = (size_t)ks * ks * cin * npad;
    int blocks = (int)((total + 255) / 256 > 2048 ? 2048 : (total + 255) / 256);
    if (tf32) pack_halo_weight_kernel<true><<<blocks, 256, 0, st>>>(w_oihw, packed, cout, cin, ks, npad);
    else pack_halo_weight_kernel<false><<<blocks, 256, 0, st>>>(w_oihw, packed, cout, cin, ks, npad);
    DS_CHECK_LAUNCH("pack_halo_weight");
    return DS_OK;
}

int halo_launch_conv(const float* src_a, int ca, const float* src_b, int cb, const HaloNorm& norm, const uint8_t* w_packed,
                     int cout, int ks, int B, int H, int W, const ConvEpi& epi, float* out_f32, void* out_b16, float* out_nchw,
                     double* sums_out, int tf32, cudaStream_t st) {
    if (ca + cb <= 224 && stream_conv_preferred(ca, cb, cout, ks, B, H, W, tf32))
        return stream_launch_conv(src_a, ca, src_b, cb, norm, w_packed, cout, ks, B, H, W, epi, out_f32, out_b16, out_nchw, sums_out, tf32, st);
    DS_REQUIRE(halo_conv_supported(ca, cb, cout, ks, B, H, W, tf32), "halo conv: unsupported shape");
    HaloParams p;
    memset(&p, 0, sizeof(p));
    const int es = tf32 ? 4 : 2;
    p.tf32 = tf32;
    p.src_a = src_a; p.src_b = src_b; p.ca = ca; p.cb = cb;
    p.stats = norm.stats; p.sums_a = norm.sums_a; p.sums_b = norm.sums_b;
    p.gamma = norm.gamma; p.beta = norm.beta; p.G = norm.G > 0 ? norm.G : 1; p.swish = norm.swish;
    DS_REQUIRE(p.G <= HALO_MAX_GROUPS, "halo conv: %d groups > %d", p.G, HALO_MAX_GROUPS);
    p.epi.sums_out = sums_out;
    p.epi.sums_B = B;
    p.w = w_packed;
    p.epi.bias = epi.bias; p.epi.temb = epi.temb; p.epi.temb_off = epi.temb_off; p.epi.temb_stride = epi.temb_stride;
    p.epi.temb_bcast = epi.temb_bcast; p.epi.residual = epi.residual;
    p.epi.out_f32 = out_f32; p.epi.out_b16 = reinterpret_cast<__nv_bfloat16*>(out_b16); p.epi.out_nchw = out_nchw;
    p.epi.Cout = cout; p.epi.Ho = H; p.epi.Wo = W;
    p.B = B; p.H = H; p.W = W; p.Wp = W + 2; p.HpWp = (H + 2) * (W + 2);
    p.total_q = B * p.HpWp;
    p.C = ca + cb; p.ksteps = p.C * es / 32; p.ntaps = ks * ks;
    p.wloads = halo_weight_loads(p.C, p.ntaps, es);
    p.Npad = (cout + 15) / 16 * 16;
    const bool two_d = halo_use_2d(B, H, W);
    const int nsamp = halo_samples_per_tile(H, W, two_d);
    const int64_t m_tiles = halo_m_tiles(B, H, W);
    p.tile2d = two_d ? 1 : 0;
    p.tiles_x = (W + HALO_TW - 1) / HALO_TW;
    p.tiles_y = (H + HALO_TH - 1) / HALO_TH;
    p.div_tiles_x = make_fastdiv((uint32_t)p.tiles_x);
    p.div_tiles_xy = make_fastdiv((uint32_t)(p.tiles_x * p.tiles_y));
    p.BN = halo_pick_bn(cout, p.C, p.ntaps, W, nsamp, m_tiles, two_d, es);
    p.n_tiles = p.Npad / p.BN;
    p.plane_px = halo_plane_px(p.ntaps, W, two_d);
    p.seg_stride_px = p.tile2d ? HALO_PW : (p.ntaps == 1 ? 0 : p.Wp);
    p.div_hpwp = make_fastdiv((uint32_t)p.HpWp);
    p.div_wp = make_fastdiv((uint32_t)p.Wp);
    p.div_planepx = make_fastdiv((uint32_t)p.plane_px);
    {
        // 3-D view of the packed weights [block][tap*kstep*plane][16 rows x 8 ch = 128 bf16]
        static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
        if (!enc) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            DS_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
            DS_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
            enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
        }
        // (128 bf16 = 256 bytes = one 16-row x 16-byte plane image; the element type only sizes the box, nothing is converted)
        const cuuint64_t tkp = (cuuint64_t)p.ntaps * p.ksteps * 2;
        cuuint64_t dims[3] = {128, (cuuint64_t)(p.Npad / 16), tkp};
        cuuint64_t strides[2] = {tkp * 256, 256};                     // bytes: next 16-channel block, next plane image
        cuuint32_t box[3] = {128, (cuuint32_t)(p.BN / 16), (cuuint32_t)(tkp / p.wloads)};
        cuuint32_t estr[3] = {1, 1, 1};
        DS_REQUIRE(p.wloads > 0 && tkp / p.wloads <= 256, "halo conv: %d taps x %d k-steps exceed the TMA boxes", p.ntaps, p.ksteps);
        CUresult r = enc(&p.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<uint8_t*>(w_packed), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("halo conv: cuTensorMapEncodeTiled for the weights failed (%d)", (int)r);
            return DS_ERR_CUDA;
        }
    }
    const size_t smem = halo_smem_bytes(p.C, p.ntaps, W, p.BN, nsamp, two_d, es);
    p.trace = trace_next(6);
    static const bool halo_dbg = getenv("DIFFSPLIT_B200_HALO_DBG") != nullptr;
    if (halo_dbg) {
        const size_t ctas = (size_t)m_tiles * p.n_tiles;
        if (!g_halo_dbg || g_halo_dbg_ctas < ctas) {
            if (g_halo_dbg) cudaFree(g_halo_dbg);
            DS_CHECK_CUDA(cudaMalloc(&g_halo_dbg, ctas * 8 * sizeof(long long)));
            g_halo_dbg_ctas = ctas;
        }
        p.dbg = g_halo_dbg;
    }
    // Staging is latency bound (global loads of the raw activations): single-wave grids whose shared-memory footprint leaves
    // one or two CTAs per SM anyway get extra warps that only stage.
    static int thr_env = -1;
    if (thr_env < 0) { const char* e = getenv("DIFFSPLIT_B200_HALO_THREADS"); thr_env = e ? atoi(e) : 0; }
    const size_t n_ctas = (size_t)m_tiles * p.n_tiles;
    int threads = (n_ctas <= 2 * 148 && smem >= 48 * 1024) ? HALO_MAX_THREADS : HALO_THREADS;
    if (thr_env >= HALO_THREADS && thr_env <= HALO_MAX_THREADS && thr_env % 32 == 0) threads = thr_env;
    const dim3 grid((unsigned)m_tiles, p.n_tiles, 1);
    typedef void (*KernelT)(const HaloParams);
    static const KernelT kernels[8] = {
        conv_halo_kernel<HALO_THREADS, false, false>, conv_halo_kernel<HALO_THREADS, true, false>,
        conv_halo_kernel<HALO_MAX_THREADS, false, false>, conv_halo_kernel<HALO_MAX_THREADS, true, false>,
        conv_halo_kernel<HALO_THREADS, false, true>, conv_halo_kernel<HALO_THREADS, true, true>,
        conv_halo_kernel<HALO_MAX_THREADS, false, true>, conv_halo_kernel<HALO_MAX_THREADS, true, true>};
    static bool attr_set = false;
    if (!attr_set) {
        for (int i = 0; i < 8; ++i)
            DS_CHECK_CUDA(cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        attr_set = true;
    }
    const KernelT kern = kernels[(tf32 ? 4 : 0) + (threads == HALO_THREADS ? 0 : 2) + (two_d ? 1 : 0)];
    const cudaError_t err = launch_pdl(kern, grid, dim3(threads), smem, st, p);
    DS_CHECK_CUDA(err);
    return DS_OK;
}

}  // namespace ds

// debugging aid: average per-phase cycle counts of the last conv_halo launch (DIFFSPLIT_B200_HALO_DBG=1)
extern "C" int ds_debug_halo_phases(double* out7, int* n_ctas) {
    using namespace ds;
    DS_REQUIRE(g_halo_dbg && out7 && n_ctas, "halo debug buffer not active");
    std::vector<long long> h(g_halo_dbg_ctas * 8);
    DS_CHECK_CUDA(cudaDeviceSynchronize());
    DS_CHECK_CUDA(cudaMemcpy(h.data(), g_halo_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 7; ++i) out7[i] = 0;
    for (size_t c = 0; c < g_halo_dbg_ctas; ++c)
        for (int i = 0; i < 7; ++i) out7[i] += (double)(h[c * 8 + i + 1] - h[c * 8 + i]);
    for (int i = 0; i < 7; ++i) out7[i] /= (double)g_halo_dbg_ctas;
    *n_ctas = (int)g_halo_dbg_ctas;
    return DS_OK;
}
