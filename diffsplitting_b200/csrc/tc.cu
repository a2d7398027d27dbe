// bf16 implicit-GEMM convolution on the 5th-generation tensor cores (precision mode DS_PREC_BF16).
//
//   out[m, n] = sum_{tap, c} A_tap[m, c] * W[n, tap, c]      m = output pixel, n = output channel
//
//  * A operand: bf16 NHWC activations, one TMA 4-D tile load per (filter tap, channel chunk): box
//    {KC channels, tw, th, tb} at coordinates {c0, x0+dx, y0+dy, b0}; out-of-bounds coordinates are zero-filled
//    by the TMA unit, which IS the conv's zero padding.  tw*th*tb = 128 output pixels = the UMMA M.
//  * B operand: weights pre-packed (once, at load time) into the exact swizzled shared-memory image of every
//    (tap, chunk) unit, fetched with a 1-D bulk copy (cp.async.bulk) - no descriptor needed.
//  * tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN (16..128), K=16 per instruction, fp32 accumulators in TMEM,
//    issued by one thread; smem slots are released by tcgen05.commit -> mbarrier.
//  * Epilogue: 4 warps, tcgen05.ld (32 lanes x 32 bit x 16 columns), + bias + conditioning vector + residual,
//    bf16 NHWC (or fp32 NCHW for the network output) stores.
//
// Folded into the load path (reference model/sr3_modules/unet.py):
//   - skip / condition concat (:255): two source tensors = two tensor maps, consecutive K ranges
//   - Downsample 3x3 stride 2 (:68-74): four parity views of the input (base pointer offset, doubled strides)
//   - Upsample nearest x2 + 3x3 (:58-65): four output-parity classes, each a 2x2 conv on the LOW-resolution
//     input with pre-summed weights (2.25x fewer MACs, no 4x intermediate)
#include <cuda.h>
#include <cudaTypedefs.h>

#include "tc.cuh"
#include "tc_ptx.cuh"

namespace ds {

// ------------------------------------------------------------------------------------------ kernel
constexpr int TC_THREADS = 192;       // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int TC_MAX_TAPS = 9;

struct TcParams {
    CUtensorMap maps[8];              // [source (a=0, b=1) * 4 + view]; view = stride-2 parity plane or 0
    const uint8_t* w;                 // packed weights: [class][unit][16-row block][16 x KC swizzled image]
    TcEpi epi;                        // bias / conditioning vector / residual / outputs
    int nb16;                         // 16-row weight blocks per (class, unit) = ceil(Cout / 16)
    int B, H, W;                      // tile space (= output size; for upsample: the LOW-res size)
    int BN, n_tiles;
    int KC, chunks_a, chunks_b;       // channel chunking of the two sources
    int ntaps, nclasses;              // taps per class, output parity classes (4 for upsample, else 1)
    int tw, th, tb, tiles_x, tiles_y; // 128-pixel tile = tw x th x tb
    int stages;
    int8_t tap_dx[4][TC_MAX_TAPS], tap_dy[4][TC_MAX_TAPS], tap_view[4][TC_MAX_TAPS];   // per class
    int up;                           // output pixel = (2y + class/2, 2x + class%2)
    int tf32;                         // operands are fp32 read as TF32 (KC <= 32 channels = 128-byte rows), else bf16
    TraceSlot trace;
};

// MMA issue loop of conv_tc_kernel: U (tap, chunk) units through a ring of `stages` slots [A 128 x KC | B BN x KC]; barriers
// full[s] at bar_base + 8 s, empty[s] behind them.
template <bool TF32, int KS>
__device__ __forceinline__ void tc_issue_units(int BN, int U, int stages, uint32_t base, uint32_t st16, uint32_t a16, uint32_t row_bytes,
                                               uint32_t bar_base, uint32_t tmem_base) {
    const uint32_t idesc = umma_idesc(BN, TF32);
    const uint64_t hi = make_smem_desc(0, row_bytes) + (uint64_t)(base >> 4);
    const uint32_t empty0 = bar_base + 8u * (uint32_t)stages;
    uint32_t s = 0, ph = 0, first = 0;
    for (int u = 0; u < U; ++u) {
        mbar_wait_fast(bar_base + 8u * s, ph);
        tc_fence_after();
        const uint64_t adesc = hi + (uint64_t)(s * st16), bdesc = adesc + (uint64_t)a16;
#pragma unroll
        for (int k = 0; k < KS; ++k) umma<TF32>(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k == 0 ? first : 1u);
        first = 1u;
        umma_commit(empty0 + 8u * s);             // frees the smem slot once these MMAs have read it
        if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
    }
}

__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row_bytes = (uint32_t)p.KC * (p.tf32 ? 4u : 2u);
    const uint32_t a_bytes = 128u * row_bytes, b_bytes = (uint32_t)p.BN * row_bytes;
    const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
    const uint32_t bar_base = base + p.stages * stage_bytes;           // full[stages], empty[stages], tmem_full, tmem slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
    const uint32_t tfull_bar = bar_base + 16u * p.stages;
    const uint32_t tmem_slot = tfull_bar + 8u;
    uint8_t* red = smem_raw + (((tmem_slot + 16u + 15u) & ~15u) - smem_u32(smem_raw));      // epilogue reduction buffer (TC_RED_BYTES, 16-byte aligned)

    // tile coordinates
    int bid = blockIdx.x;
    const int tx_i = bid % p.tiles_x; bid /= p.tiles_x;
    const int ty_i = bid % p.tiles_y; bid /= p.tiles_y;
    const int tb_i = bid;
    const int x0 = tx_i * p.tw, y0 = ty_i * p.th, b0 = tb_i * p.tb;
    const int nt = blockIdx.y, cls = blockIdx.z;
    const int units_per_tap = p.chunks_a + p.chunks_b;
    const int U = p.ntaps * units_per_tap;
    const uint32_t tmem_cols = (uint32_t)p.BN <= 32u ? 32u : ((uint32_t)p.BN <= 64u ? 64u : 128u);

    trace_begin(p.trace);
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (elect_one()) {
            const size_t blk16 = (size_t)16 * row_bytes;
            const uint8_t* wsrc = p.w + ((size_t)cls * U * p.nb16 + (size_t)nt * (p.BN / 16)) * blk16;
            const size_t ustride = (size_t)p.nb16 * blk16;
            int s = 0;
            uint32_t eph = 1u;                                        // parity to wait for on empty[s]
            for (int tap = 0; tap < p.ntaps; ++tap) {
                const int xs = x0 + p.tap_dx[cls][tap], ys = y0 + p.tap_dy[cls][tap], view = p.tap_view[cls][tap];
                for (int cc = 0; cc < units_per_tap; ++cc) {
                    mbar_wait_fast(empty_bar(s), eph);
                    const int src = cc < p.chunks_a ? 0 : 1;
                    const int c0 = (src == 0 ? cc : cc - p.chunks_a) * p.KC;
                    const uint32_t dstA = base + s * stage_bytes;
                    mbar_expect_tx(full_bar(s), a_bytes + b_bytes);
                    tma_load_4d(dstA, &p.maps[src * 4 + view], full_bar(s), c0, xs, ys, b0);
                    bulk_load(dstA + a_bytes, wsrc, b_bytes, full_bar(s));
                    wsrc += ustride;
                    if (++s == p.stages) { s = 0; eph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            // 32 bytes of every row per MMA (K = 16 bf16 / 8 tf32); one thread issues everything, so the loop is kept lean
            // (tc_issue_units: running stage index and descriptor words)
            const int ksteps = (int)(row_bytes >> 5);
            const uint32_t st16 = stage_bytes >> 4, a16 = a_bytes >> 4;
            if (p.tf32) {
                if (ksteps == 4) tc_issue_units<true, 4>(p.BN, U, p.stages, base, st16, a16, row_bytes, bar_base, tmem_base);
                else if (ksteps == 2) tc_issue_units<true, 2>(p.BN, U, p.stages, base, st16, a16, row_bytes, bar_base, tmem_base);
                else tc_issue_units<true, 1>(p.BN, U, p.stages, base, st16, a16, row_bytes, bar_base, tmem_base);
            } else {
                if (ksteps == 4) tc_issue_units<false, 4>(p.BN, U, p.stages, base, st16, a16, row_bytes, bar_base, tmem_base);
                else if (ksteps == 2) tc_issue_units<false, 2>(p.BN, U, p.stages, base, st16, a16, row_bytes, bar_base, tmem_base);
                else tc_issue_units<false, 1>(p.BN, U, p.stages, base, st16, a16, row_bytes, bar_base, tmem_base);
            }
            umma_commit(tfull_bar);               // accumulator complete
        }
        __syncwarp();
    } else {
        // ---- epilogue: warp q owns TMEM lanes [32q, 32q+32) == output rows
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int xl = m % p.tw, yl = (m / p.tw) % p.th, bl = m / (p.tw * p.th);
        const int x = x0 + xl, y = y0 + yl, b = b0 + bl;
        const bool valid = x < p.W && y < p.H && b < p.B;
        int oy = y, ox = x;
        if (p.up) { oy = 2 * y + (cls >> 1); ox = 2 * x + (cls & 1); }
        const int n_base = nt * p.BN;
        float add[16];
        if (valid) tc_epilogue_addend(p.epi, b, oy, ox, n_base, add);       // in flight while the MMAs run
        mbar_wait(tfull_bar, 0);
        tc_fence_after();
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            float f[16];
            if (valid) {
                if (c0) tc_epilogue_addend(p.epi, b, oy, ox, n_base + c0, add);
                tc_epilogue_write(p.epi, v, add, b, oy, ox, n_base + c0, f);
            }
            if (p.epi.sums_out) {
                const int nsr = min(p.tb, p.B - b0);
                if (nsr == 1) tc_epilogue_stats_smem(p.epi, f, valid, b0, n_base + c0, (int)threadIdx.x - 64, (int)(blockIdx.x % TC_SUM_COPIES), red);
                else if (nsr == 2) tc_epilogue_stats_shfl(p.epi, f, valid, b, n_base + c0, (int)threadIdx.x - 64, b0, nsr, (int)(blockIdx.x % TC_SUM_COPIES), red);
                else tc_epilogue_stats(p.epi, f, valid, b, n_base + c0, m, (int)threadIdx.x - 64, b0, (int)(blockIdx.x % TC_SUM_COPIES), red);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
    trace_end(p.trace);
}

// ------------------------------------------------------------------------------------------ tall-patch variant
// For 3x3 stride-1 convolutions on large images the kernel above is bound by what an SM can pull from L2 (~25-30 B/clk): it
// fetches every input pixel once per filter tap and the whole weight tile per 128-pixel tile.  Here a CTA owns MT
// vertically adjacent tiles of 7 x 16 outputs and, per 64-channel chunk, loads ONE patch of (7 MT + 2) x 18 pixels with a
// single TMA box (out-of-bounds rows / columns zero-filled = the padding).  In that patch (row pitch 18) a filter tap is a
// constant row shift, so the A operand of (tile t, tap r,s) is just the descriptor start  patch + (126 t + 18 r + s) rows:
// the tensor core's swizzle is a pure function of the shared-memory address, a window may start on any row of a swizzle
// atom (profiles/README.md, finding 3).  Every weight stage is applied to all MT tiles (MT accumulators in TMEM), so the
// weights are fetched once per 7 MT x 16 pixels.  Ingest per 64-channel chunk at 128 -> 128, MT = 3: 53 KB of pixels +
// 147 KB of weights for 6.9k cycles of MMAs instead of 288 KB for 2.3k.
constexpr int TP_TW = 16, TP_TH = 7, TP_PW = TP_TW + 2;
constexpr int TP_MAX_MT = 3;
constexpr int TP_GROUPS = 4;               // epilogue groups of four warps (one warp per TMEM lane quadrant each)
constexpr int TP_THREADS = 64 + 128 * TP_GROUPS;   // warp 0 TMA, warp 1 MMA, then the epilogue groups

struct TcpParams {
    CUtensorMap pmap[2];              // sources a / b as (C, W, H, B), box {KC, 18, 7 MT + 2, 1}
    const uint8_t* w;                 // the same packed weights as conv_tc_kernel (class 0)
    TcEpi epi;
    int nb16, B, H, W, BN, n_tiles, KC, chunks_a, chunks_b, MT, tiles_x, tiles_y, wstages;
    uint32_t patch_bytes;
    int tf32;
    // persistent variant (conv_tcs_kernel)
    CUtensorMap resmap;               // fp32 residual as (Cout, W, H, B), box {20, 8, 16, 1} = one chunk in the staging layout
    int res_tma, n_items;             // res_tma: the fp32 residual arrives through resmap (box {20, 8, 16, 1}) in the staging buffers
           // work items = (M supertile, N tile[, output parity class]), N (class) fastest
    int ngroups;                      // epilogue groups in use: 4, or 3 (BN = 128 with 128-byte rows: shared memory goes to the weight ring)
    int up;                           // nearest x2 upsample + 3x3 as four 2x2 convs on the low-res patch (geo 1 only): H, W = low-res
                                      //    size, item class (a, b) writes the output pixels (2y + a, 2x + b); w = [class][tap][chunk]
    uint32_t class_bytes;             // packed weight bytes per class
    // folded 1x1 conv of a second input (the res_conv of a ResNet block, out = conv3x3(h) + conv1x1(cat[xa, xb])): extra K chunks
    // after the 3x3 ones, one tap each, A = the 16 x 16 centre box of the supertile (8-row groups 16 rows apart)
    CUtensorMap xmap[2];              // sources xa / xb as (C, W, H, B), box {KC, 16, 16, 1}
    const uint8_t* xw;                // their packed 1x1 weights [chunk][16-row block] (same KC)
    int xchunks_a, xchunks_b;
    int geo;                          // 0: supertile = 2 stacked tiles of 7 x 16 outputs (128 MMA rows = patch positions of pitch 18, 112
                                      //    valid);  1: supertile = 16 x 16 outputs as two 8-wide tiles side by side: MMA row 8 g + i =
                                      //    output (row g, column i) -> an 8-row core-matrix group is 8 consecutive patch pixels and the
                                      //    groups are one patch row (18 pixels) apart: SBO = 18 rows instead of 8, all 128 rows valid
    FastDiv div_ntiles, div_tiles_x, div_tiles_xy;
    TraceSlot trace;
};

// MMA issue loop of conv_tcp_kernel (one thread; see tcs_issue for why it is written this way)
struct TcpIssue {
    uint32_t base, wring, wstage_bytes, bar_base, pfull0, tmem_base, row_bytes;
    int nchunks;
};
template <bool TF32, int KS>
__device__ __forceinline__ void tcp_issue(const TcpParams& p, const TcpIssue& ii) {
    const uint32_t idesc = umma_idesc(p.BN, TF32);
    const uint32_t rb16 = ii.row_bytes >> 4;
    const uint64_t a_hi = make_smem_desc(0, ii.row_bytes);
    const uint64_t b_hi = a_hi + (uint64_t)(ii.wring >> 4);
    const uint32_t t_step = (uint32_t)(TP_TH * TP_PW) * rb16, wst16 = ii.wstage_bytes >> 4;
    const uint32_t nw = (uint32_t)p.wstages, wempty0 = ii.bar_base + 8u * nw;
    const uint32_t BN = (uint32_t)p.BN;
    const int MT = p.MT;
    uint32_t s = 0, wph = 0, first = 0;
    for (int cc = 0; cc < ii.nchunks; ++cc) {
        const uint32_t pb = (uint32_t)cc & 1u;
        mbar_wait_fast(ii.pfull0 + 8u * pb, ((uint32_t)cc >> 1) & 1u);
        tc_fence_after();
        const uint64_t abase = a_hi + (uint64_t)((ii.base + pb * p.patch_bytes) >> 4);
        for (int ty = 0; ty < 3; ++ty) {
            for (int tx = 0; tx < 3; ++tx) {
                mbar_wait_fast(ii.bar_base + 8u * s, wph);
                tc_fence_after();
                const uint64_t bdesc = b_hi + (uint64_t)(s * wst16);
                const uint64_t adesc = abase + (uint64_t)((uint32_t)(ty * TP_PW + tx) * rb16);
                for (int t = 0; t < MT; ++t) {
#pragma unroll
                    for (int k = 0; k < KS; ++k)
                        umma<TF32>(ii.tmem_base + (uint32_t)t * BN, adesc + (uint64_t)((uint32_t)t * t_step + 2u * k), bdesc + (uint64_t)(2 * k), idesc,
                                   k == 0 ? first : 1u);
                }
                first = 1u;
                umma_commit(wempty0 + 8u * s);
                if (++s == nw) { s = 0; wph ^= 1u; }
            }
        }
        umma_commit(ii.pfull0 + 16u + 8u * pb);                                 // pempty
    }
}

__global__ void __launch_bounds__(TP_THREADS) conv_tcp_kernel(const __grid_constant__ TcpParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row_bytes = (uint32_t)p.KC * (p.tf32 ? 4u : 2u);
    const uint32_t w_bytes = (uint32_t)p.BN * row_bytes;                    // one (tap, chunk) weight tile
    const uint32_t wstage_bytes = (w_bytes + 1023u) & ~1023u;
    const uint32_t wring = base + 2u * p.patch_bytes;
    const uint32_t bar_base = wring + (uint32_t)p.wstages * wstage_bytes;
    auto wfull = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto wempty = [&](int s) { return bar_base + 8u * (uint32_t)(p.wstages + s); };
    const uint32_t pfull0 = bar_base + 16u * (uint32_t)p.wstages;          // pfull[2], pempty[2], tfull, tmem slot
    auto pfull = [&](int i) { return pfull0 + 8u * (uint32_t)i; };
    auto pempty = [&](int i) { return pfull0 + 16u + 8u * (uint32_t)i; };
    const uint32_t tfull = pfull0 + 32u, tmem_slot = pfull0 + 40u;
    uint8_t* red = smem_raw + (((tmem_slot + 16u + 15u) & ~15u) - smem_u32(smem_raw));       // 16-byte aligned

    int bid = blockIdx.x;
    const int tx_i = bid % p.tiles_x; bid /= p.tiles_x;
    const int ty_i = bid % p.tiles_y; bid /= p.tiles_y;
    const int b = bid;
    const int x0 = tx_i * TP_TW, y0 = ty_i * TP_TH * p.MT;
    const int nt = blockIdx.y;
    const int nchunks = p.chunks_a + p.chunks_b;
    const uint32_t ncols = (uint32_t)(p.MT * p.BN);
    const uint32_t tmem_cols = ncols <= 32u ? 32u : (ncols <= 64u ? 64u : (ncols <= 128u ? 128u : (ncols <= 256u ? 256u : 512u)));

    trace_begin(p.trace);
    if (warp == 0 && elect_one()) {
        for (int s = 0; s < p.wstages; ++s) { mbar_init(wfull(s), 1); mbar_init(wempty(s), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(pfull(i), 1); mbar_init(pempty(i), 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (elect_one()) {
            const size_t blk16 = (size_t)16 * row_bytes;
            const uint8_t* wsrc = p.w + (size_t)nt * (p.BN / 16) * blk16;
            const uint32_t box_bytes = (uint32_t)(TP_PW * (TP_TH * p.MT + 2)) * row_bytes;
            int s = 0;
            uint32_t eph = 1u;
            for (int cc = 0; cc < nchunks; ++cc) {
                const int pb = cc & 1;
                mbar_wait_fast(pempty(pb), (((uint32_t)(cc >> 1)) & 1u) ^ 1u);
                const int src = cc < p.chunks_a ? 0 : 1;
                const int c0 = (src == 0 ? cc : cc - p.chunks_a) * p.KC;
                mbar_expect_tx(pfull(pb), box_bytes);
                tma_load_4d(base + (uint32_t)pb * p.patch_bytes, &p.pmap[src], pfull(pb), c0, x0 - 1, y0 - 1, b);
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait_fast(wempty(s), eph);
                    mbar_expect_tx(wfull(s), w_bytes);
                    bulk_load(wring + (uint32_t)s * wstage_bytes, wsrc + (size_t)(tap * nchunks + cc) * p.nb16 * blk16, w_bytes, wfull(s));
                    if (++s == p.wstages) { s = 0; eph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            TcpIssue ii;
            ii.base = base; ii.wring = wring; ii.wstage_bytes = wstage_bytes; ii.bar_base = bar_base; ii.pfull0 = pfull0;
            ii.tmem_base = tmem_base; ii.row_bytes = row_bytes; ii.nchunks = nchunks;
            const int ksteps = (int)(row_bytes >> 5);
            if (p.tf32) {
                if (ksteps == 4) tcp_issue<true, 4>(p, ii); else if (ksteps == 2) tcp_issue<true, 2>(p, ii); else tcp_issue<true, 1>(p, ii);
            } else {
                if (ksteps == 4) tcp_issue<false, 4>(p, ii); else if (ksteps == 2) tcp_issue<false, 2>(p, ii); else tcp_issue<false, 1>(p, ii);
            }
            umma_commit(tfull);
        }
        __syncwarp();
    } else {
        // TP_GROUPS groups of four warps: group g takes the 16-column chunks g, g + TP_GROUPS, ... of the CTA's MT tiles.  The
        // epilogue is bound by the latency of the residual loads (HBM), so the addends of the next TWO chunks are in flight
        // while one is processed, and the groups multiply the loads in flight per SM.
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int m = q * 32 + lane;
        const int te = (int)threadIdx.x - 64 - grp * 128;
        uint8_t* redg = red + (size_t)grp * TC_RED_BYTES;
        const int pr = m / TP_PW, pc = m - pr * TP_PW;
        const int x = x0 + pc;
        const int cpt = p.BN >> 4;                                   // chunks per tile
        const int G = p.MT * cpt;
        const int nwork = (G - grp + TP_GROUPS - 1) / TP_GROUPS;
        auto where = [&](int i, int& t, int& c0, int& y, bool& valid) {
            const int g = grp + TP_GROUPS * i;
            t = g / cpt;
            c0 = (g - t * cpt) * 16;
            y = y0 + t * TP_TH + pr;
            valid = pr < TP_TH && pc < TP_TW && y < p.H && x < p.W;
        };
        float addA[16], addB[16];
        {
            int t, c0, y; bool valid;
            if (nwork > 0) { where(0, t, c0, y, valid); if (valid) tc_epilogue_addend(p.epi, b, y, x, nt * p.BN + c0, addA); }
            if (nwork > 1) { where(1, t, c0, y, valid); if (valid) tc_epilogue_addend(p.epi, b, y, x, nt * p.BN + c0, addB); }
        }
        mbar_wait(tfull, 0);
        tc_fence_after();
        auto work = [&](int i, float (&add)[16]) {
            int t, c0, y; bool valid;
            where(i, t, c0, y, valid);
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.BN + c0), v);
            float f[16];
            if (valid) tc_epilogue_write(p.epi, v, add, b, y, x, nt * p.BN + c0, f);
            if (i + 2 < nwork) {                                       // refill this slot with the addend of chunk i + 2
                int t2, c2, y2; bool v2;
                where(i + 2, t2, c2, y2, v2);
                if (v2) tc_epilogue_addend(p.epi, b, y2, x, nt * p.BN + c2, add);
            }
            if (p.epi.sums_out) {
                const int copy = (int)(blockIdx.x % TC_SUM_COPIES);
                switch (grp) {
                    case 0: tc_epilogue_stats_smem<1>(p.epi, f, valid, b, nt * p.BN + c0, te, copy, redg); break;
                    case 1: tc_epilogue_stats_smem<2>(p.epi, f, valid, b, nt * p.BN + c0, te, copy, redg); break;
                    case 2: tc_epilogue_stats_smem<3>(p.epi, f, valid, b, nt * p.BN + c0, te, copy, redg); break;
                    default: tc_epilogue_stats_smem<4>(p.epi, f, valid, b, nt * p.BN + c0, te, copy, redg); break;
                }
            }
        };
        for (int i = 0; i < nwork; i += 2) {
            work(i, addA);
            if (i + 1 < nwork) work(i + 1, addB);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
    trace_end(p.trace);
}

// ------------------------------------------------------------------------------------------ persistent tall-patch variant
// conv_tcp_kernel spends most of a CTA's life outside the MMAs (ncu, 128 -> 128 at 8 x 256^2: tensor pipe 20 % active): TMEM
// allocation, barrier set-up and pipeline fill per CTA, then an epilogue (residual loads, fp32 + bf16 stores, statistics) that
// only starts when the last MMA has retired - 12 waves of that on 148 SMs.  Here ONE CTA per SM walks the work items
// (M supertile of MT = 2 tiles x N tile) w = blockIdx.x, + gridDim.x, ... with TWO accumulator sets in TMEM (2 x MT x BN <= 512
// columns), so the epilogue of item i runs under the MMAs of item i + 1:
//   warp 0       TMA producer: per 64-channel chunk one (7 MT + 2) x 18 patch (ring of 2) + 9 weight tiles (ring of wstages);
//                L2 prefetch of the item's residual tile
//   warp 1       MMA issuer: 9 taps x MT tiles x 4 k-steps per chunk into ACC[i & 1]
//   warps 2-17   epilogue, four (or three) groups of four warps (one per TMEM lane quadrant) sharing the item's 16-column chunks, each chunk
//                in two views (as conv_stream_kernel): ROW view  tcgen05.ld + (bias + conditioning vector) -> staging buffer;
//                QUAD view  + residual (requested one chunk ahead), fp32 / bf16 stores with 64 contiguous bytes per four lanes,
//                statistics in registers, folded once per item
constexpr int TS_MT = 2;
constexpr int TS_THREADS = 576;            // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue: up to four groups of four warps
constexpr int TS_EGROUPS = 4;              // most epilogue groups (TcpParams::ngroups = 3 when four staging areas do not fit)
constexpr int TS_STG_LD = 20;
constexpr uint32_t TS_STG_BUF = 128 * TS_STG_LD * 4;                 // one staging buffer [128][20] fp32
constexpr uint32_t TS_STG_BYTES = 2 * TS_STG_BUF + 128 * 4;          // per group: two staging buffers + (bias + temb) of the item's BN channels

__device__ __noinline__ void tcs_store_tail(const TcEpi& e, float4 x, int pix, int n0) {
    const float xs[4] = {x.x, x.y, x.z, x.w};
    const int HW = e.Ho * e.Wo;
    if (e.out_nchw) {
        const int pb = pix / HW, pp = pix - pb * HW;
        for (int k = 0; k < 4; ++k)
            if (n0 + k < e.Cout) e.out_nchw[((size_t)pb * e.Cout + n0 + k) * HW + pp] = xs[k];
        return;
    }
    const size_t off = (size_t)pix * e.Cout + n0;
    for (int k = 0; k < 4; ++k) {
        if (n0 + k < e.Cout) {
            if (e.out_f32) e.out_f32[off + k] = xs[k];
            if (e.out_b16) e.out_b16[off + k] = __float2bfloat16_rn(xs[k]);
        }
    }
}
__device__ __noinline__ float4 tcs_load_tail(const float* src, int n) {
    return make_float4(__ldg(src), n > 1 ? __ldg(src + 1) : 0.f, n > 2 ? __ldg(src + 2) : 0.f, n > 3 ? __ldg(src + 3) : 0.f);
}

// The MMA issuer is ONE thread: every instruction between two tcgen05.mma is serial latency in front of the tensor pipe (a
// 128 x 128 x 16 MMA is 64 cycles of tensor work, a 128 x 64 x 16 one 32), so the loop carries running ring indices and
// descriptor words instead of recomputing them: no division, no descriptor assembly, no clock reads on the fast path.
struct TcsIssue {
    uint32_t base, wring, wstage_bytes, bar_base, b2, tmem_base, acc_cols, row_bytes;
    int my_items, nchunks;
};
template <bool TF32, int KS>
__device__ __forceinline__ void tcs_issue(const TcpParams& p, const TcsIssue& ii) {
    const uint32_t idesc = umma_idesc(p.BN, TF32);
    const uint32_t rb16 = ii.row_bytes >> 4;                      // descriptor address units (16 B) per row
    const uint64_t a_hi = p.geo ? make_smem_desc_sbo(0, ii.row_bytes, TP_PW * ii.row_bytes) : make_smem_desc(0, ii.row_bytes);
    const uint64_t b_hi = make_smem_desc(0, ii.row_bytes) + (uint64_t)(ii.wring >> 4);
    const uint32_t t_step = (p.geo ? 8u : (uint32_t)(TP_TH * TP_PW)) * rb16;
    const uint32_t wst16 = ii.wstage_bytes >> 4;
    const uint32_t nw = (uint32_t)p.wstages;
    const uint32_t wfull0 = ii.bar_base, wempty0 = ii.bar_base + 8u * nw;
    const int nt1 = p.up ? 2 : 3;                                 // taps per axis
    const uint32_t BN = (uint32_t)p.BN;
    const uint64_t ax_hi = make_smem_desc_sbo(0, ii.row_bytes, 16u * ii.row_bytes);
    const int xchunks = p.xchunks_a + p.xchunks_b;
    uint32_t s = 0, wph = 0, cg = 0;
    for (int i = 0; i < ii.my_items; ++i) {
        const uint32_t a = (uint32_t)i & 1u;
        // upsample: class (oa, ob) = low two bits of the item index; tap (ty, tx) reads the low-res pixel (ty + oa - 1, tx + ob - 1)
        uint32_t o0 = 0;
        if (p.up) {
            const uint32_t w = blockIdx.x + (uint32_t)i * gridDim.x;
            o0 = ((w >> 1) & 1u) * TP_PW + (w & 1u);
        }
        mbar_wait(ii.b2 + 48u + 8u * a, (((uint32_t)i >> 1) & 1u) ^ 1u);       // aempty
        tc_fence_after();
        const uint32_t acc = ii.tmem_base + a * ii.acc_cols;
        uint32_t first = 0;                                       // accumulate flag of the k = 0 MMAs: 0 on the item's first tap
        for (int cc = 0; cc < ii.nchunks; ++cc, ++cg) {
            const uint32_t pb = cg & 1u;
            mbar_wait_fast(ii.b2 + 8u * pb, (cg >> 1) & 1u);                    // pfull
            tc_fence_after();
            const uint64_t abase = a_hi + (uint64_t)(((ii.base + pb * p.patch_bytes) >> 4) + o0 * rb16);
            for (int ty = 0; ty < nt1; ++ty) {
                for (int tx = 0; tx < nt1; ++tx) {
                    mbar_wait_fast(wfull0 + 8u * s, wph);
                    tc_fence_after();
                    const uint64_t bdesc = b_hi + (uint64_t)(s * wst16);
                    const uint64_t adesc = abase + (uint64_t)((uint32_t)(ty * TP_PW + tx) * rb16);
#pragma unroll
                    for (int t = 0; t < TS_MT; ++t) {
#pragma unroll
                        for (int k = 0; k < KS; ++k)
                            umma<TF32>(acc + (uint32_t)t * BN, adesc + (uint64_t)((uint32_t)t * t_step + 2u * k), bdesc + (uint64_t)(2 * k), idesc,
                                       k == 0 ? first : 1u);
                    }
                    first = 1u;
                    umma_commit(wempty0 + 8u * s);
                    if (++s == nw) { s = 0; wph ^= 1u; }
                }
            }
            umma_commit(ii.b2 + 16u + 8u * pb);                                 // pempty
        }
        for (int xc = 0; xc < xchunks; ++xc, ++cg) {                            // folded 1x1 conv: one tap per chunk
            const uint32_t pb = cg & 1u;
            mbar_wait_fast(ii.b2 + 8u * pb, (cg >> 1) & 1u);
            mbar_wait_fast(wfull0 + 8u * s, wph);
            tc_fence_after();
            const uint64_t bdesc = b_hi + (uint64_t)(s * wst16);
            const uint64_t adesc = ax_hi + (uint64_t)((ii.base + pb * p.patch_bytes) >> 4);
#pragma unroll
            for (int t = 0; t < TS_MT; ++t) {
#pragma unroll
                for (int k = 0; k < KS; ++k)
                    umma<TF32>(acc + (uint32_t)t * BN, adesc + (uint64_t)((uint32_t)t * 8u * rb16 + 2u * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
            }
            umma_commit(wempty0 + 8u * s);
            if (++s == nw) { s = 0; wph ^= 1u; }
            umma_commit(ii.b2 + 16u + 8u * pb);
        }
        umma_commit(ii.b2 + 32u + 8u * a);                                      // afull
    }
}

__global__ void __launch_bounds__(TS_THREADS, 1) conv_tcs_kernel(const __grid_constant__ TcpParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t row_bytes = (uint32_t)p.KC * (p.tf32 ? 4u : 2u);
    const uint32_t w_bytes = (uint32_t)p.BN * row_bytes;
    const uint32_t wstage_bytes = (w_bytes + 1023u) & ~1023u;
    const uint32_t wring = base + 2u * p.patch_bytes;
    const int NG = p.ngroups;
    const uint32_t stg_off = 2u * p.patch_bytes + (uint32_t)p.wstages * wstage_bytes;      // NG x TS_STG_BYTES
    const uint32_t bar_base = base + stg_off + (uint32_t)NG * TS_STG_BYTES;
    auto wfull = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto wempty = [&](int s) { return bar_base + 8u * (uint32_t)(p.wstages + s); };
    const uint32_t b2 = bar_base + 16u * (uint32_t)p.wstages;
    auto pfull = [&](int i) { return b2 + 8u * (uint32_t)i; };
    auto pempty = [&](int i) { return b2 + 16u + 8u * (uint32_t)i; };
    auto afull = [&](int i) { return b2 + 32u + 8u * (uint32_t)i; };
    auto aempty = [&](int i) { return b2 + 48u + 8u * (uint32_t)i; };
    const uint32_t tmem_slot = b2 + 64u;
    auto rfull = [&](int g, uint32_t buf) { return b2 + 80u + 16u * (uint32_t)g + 8u * buf; };
    const int nchunks = p.chunks_a + p.chunks_b;
    const int cpt = p.BN >> 4;                                    // 16-column chunks per tile
    const uint32_t acc_cols = (uint32_t)(TS_MT * p.BN);
    const uint32_t tmem_cols = 2u * acc_cols <= 32u ? 32u : (2u * acc_cols <= 64u ? 64u : (2u * acc_cols <= 128u ? 128u : (2u * acc_cols <= 256u ? 256u : 512u)));
    const int my_items = ((int)blockIdx.x < p.n_items) ? (p.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int egroups = cpt * TS_MT >= NG ? NG : cpt * TS_MT;                         // epilogue groups with work

    trace_begin(p.trace);
    if (warp == 0 && elect_one()) {
        for (int s = 0; s < p.wstages; ++s) { mbar_init(wfull(s), 1); mbar_init(wempty(s), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(pfull(i), 1); mbar_init(pempty(i), 1);
            mbar_init(afull(i), 1); mbar_init(aempty(i), 4 * egroups);
            for (int g = 0; g < TS_EGROUPS; ++g) mbar_init(rfull(g, (uint32_t)i), 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();
    pdl_trigger();

    auto item_coords = [&](int i, int& nt, int& b, int& y0, int& x0) {
        const int w = (int)blockIdx.x + i * (int)gridDim.x;
        const int mi = fdiv(w, p.div_ntiles);
        nt = w - mi * (p.up ? 4 * p.n_tiles : p.n_tiles);                      // up: N tile * 4 + class
        b = fdiv(mi, p.div_tiles_xy);
        const int rem = mi - b * p.tiles_x * p.tiles_y;
        const int ty = fdiv(rem, p.div_tiles_x);
        y0 = ty * (p.geo ? 16 : TP_TH * TS_MT);
        x0 = (rem - ty * p.tiles_x) * TP_TW;
    };

    if (warp == 0) {
        if (elect_one()) {
            const size_t blk16 = (size_t)16 * row_bytes;
            const uint32_t box_bytes = (uint32_t)(TP_PW * (p.geo ? 18 : TP_TH * TS_MT + 2)) * row_bytes;
            uint32_t cg = 0, s = 0, wph = 1;                     // chunk counter, weight ring stage and (empty-barrier) phase
            for (int i = 0; i < my_items; ++i) {
                int nt, b, y0, x0;
                item_coords(i, nt, b, y0, x0);
                const int ntaps = p.up ? 4 : 9;
                const uint8_t* wsrc = p.w + (size_t)(p.up ? nt >> 2 : nt) * (p.BN / 16) * blk16 + (p.up ? (size_t)(nt & 3) * p.class_bytes : 0);
                for (int cc = 0; cc < nchunks; ++cc, ++cg) {
                    const int pb = (int)(cg & 1u);
                    mbar_wait_relaxed(pempty(pb), ((cg >> 1) & 1u) ^ 1u);
                    const int src = cc < p.chunks_a ? 0 : 1;
                    const int c0 = (src == 0 ? cc : cc - p.chunks_a) * p.KC;
                    mbar_expect_tx(pfull(pb), box_bytes);
                    tma_load_4d(base + (uint32_t)pb * p.patch_bytes, &p.pmap[src], pfull(pb), c0, x0 - 1, y0 - 1, b);
                    for (int tap = 0; tap < ntaps; ++tap) {
                        mbar_wait_relaxed(wempty((int)s), wph);
                        mbar_expect_tx(wfull((int)s), w_bytes);
                        bulk_load(wring + s * wstage_bytes, wsrc + (size_t)(tap * nchunks + cc) * p.nb16 * blk16, w_bytes, wfull((int)s));
                        if (++s == (uint32_t)p.wstages) { s = 0; wph ^= 1u; }
                    }
                }
                const int xchunks = p.xchunks_a + p.xchunks_b;
                for (int xc = 0; xc < xchunks; ++xc, ++cg) {
                    const int pb = (int)(cg & 1u);
                    mbar_wait_relaxed(pempty(pb), ((cg >> 1) & 1u) ^ 1u);
                    const int src = xc < p.xchunks_a ? 0 : 1;
                    const int c0 = (src == 0 ? xc : xc - p.xchunks_a) * p.KC;
                    mbar_expect_tx(pfull(pb), 256u * row_bytes);
                    tma_load_4d(base + (uint32_t)pb * p.patch_bytes, &p.xmap[src], pfull(pb), c0, x0, y0, b);
                    mbar_wait_relaxed(wempty((int)s), wph);
                    mbar_expect_tx(wfull((int)s), w_bytes);
                    bulk_load(wring + s * wstage_bytes, p.xw + ((size_t)xc * p.nb16 + (size_t)nt * (p.BN / 16)) * blk16, w_bytes, wfull((int)s));
                    if (++s == (uint32_t)p.wstages) { s = 0; wph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            TcsIssue ii;
            ii.base = base; ii.wring = wring; ii.wstage_bytes = wstage_bytes; ii.bar_base = bar_base; ii.b2 = b2;
            ii.tmem_base = tmem_base; ii.acc_cols = acc_cols; ii.my_items = my_items; ii.row_bytes = row_bytes; ii.nchunks = nchunks;
            if (p.tf32) {
                if (row_bytes == 128) tcs_issue<true, 4>(p, ii); else tcs_issue<true, 2>(p, ii);
            } else {
                if (row_bytes == 128) tcs_issue<false, 4>(p, ii); else tcs_issue<false, 2>(p, ii);
            }
        }
        __syncwarp();
    } else if (warp >= 2 && ((warp - 2) >> 2) < NG) {
        const int gi = (warp - 2) >> 2;                          // epilogue group 0 .. NG - 1 (any four consecutive warps cover the
        const int q = warp & 3;                                  // four TMEM lane quadrants: a warp reads lanes [32 (warp % 4), + 32))
        const int m = q * 32 + lane;
        const int te = tid - 64 - gi * 128;
        float* stg = reinterpret_cast<float*>(gbase + stg_off + (uint32_t)gi * TS_STG_BYTES);
        float* btv = stg + 2 * 128 * TS_STG_LD;                  // [BN <= 128] bias + conditioning vector of the item
        const int quad = te & 3, r0 = te >> 2;
        const int Cout = p.epi.Cout;
        const bool vec_ok = (Cout & 3) == 0 && !p.epi.out_nchw;
        const int lcpt = 31 - __clz(cpt);                        // BN is a power of two
        // the item's chunks (tile t, 16-column block c) are dealt to the groups so that a group touches at most 3 columns (its
        // statistics stay in 3 register slots): four groups - by chunk index g = t cpt + c = gi, gi + 4, ... (columns gi mod cpt and
        // + 4); three groups (only with 8 columns per tile) - by column c = gi, gi + 3, gi + 6, both tiles of a column in turn
        const bool bycol = NG == 3;
        const int n_mine = bycol ? TS_MT * ((cpt - gi + 2) / 3) : (TS_MT * cpt > gi ? (TS_MT * cpt - gi + 3) / 4 : 0);
        const float* const resp = p.epi.residual;
        float* const o32 = p.epi.out_f32;
        __nv_bfloat16* const o16 = p.epi.out_b16;
        double* const sums = p.epi.sums_out;
        const int copy = (int)(blockIdx.x % TC_SUM_COPIES);
        const uint32_t bar_id = (uint32_t)gi + 1u;
        const float* const srow = stg + r0 * TS_STG_LD + quad * 4;
        const bool rtma = p.res_tma != 0;
        const int t_first = bycol ? 0 : gi >> lcpt, c_first = bycol ? gi : gi & (cpt - 1);
        if (n_mine > 0) {
            // residual through TMA: thread 0 of the group runs a cursor two chunks ahead of the group over its (item, chunk)
            // sequence; chunk n lands in staging buffer n & 1 (in the staging layout: 80-byte rows = 16 + 4 channels), the ROW view
            // adds the accumulator to it
            int qi = 0, qt = t_first, qc = c_first;
            uint32_t qn = 0, cn = 0;
            auto issue_next = [&]() {
                if (qi >= my_items) return;
                int nt, b, y0, x0;
                item_coords(qi, nt, b, y0, x0);
                const uint32_t buf = qn & 1u;
                mbar_expect_tx(rfull(gi, buf), TS_STG_BUF);
                tma_load_4d(smem_u32(stg) + buf * TS_STG_BUF, &p.resmap, rfull(gi, buf), nt * p.BN + qc * 16, x0 + 8 * qt, y0, b);
                ++qn;
                if (bycol) {
                    qc += 3;
                    if (qc >= 8) { qc = gi; ++qt; }
                } else {
                    const int g = (qt << lcpt) + qc + 4;
                    qt = g >> lcpt; qc = g & (cpt - 1);
                }
                if (qt >= TS_MT) { ++qi; qt = t_first; qc = c_first; }
            };
            if (rtma && te == 0) { issue_next(); issue_next(); }
#pragma unroll 1
            for (int i = 0; i < my_items; ++i) {
                const int a = i & 1;
                int nt, b, y0, x0;
                item_coords(i, nt, b, y0, x0);
                int oa = 0, ob = 0;                               // upsample: output parity of the item
                if (p.up) { oa = (nt >> 1) & 1; ob = nt & 1; nt >>= 2; }
                // (bias + conditioning vector) of the item's channels -> btv (the previous item's last chunk barrier orders this
                // write after its last read; the barrier below makes it visible)
                if (te < p.BN) {
                    const int n = nt * p.BN + te;
                    float v = 0.f;
                    if (n < Cout) {
                        if (p.epi.bias) v = __ldg(p.epi.bias + n);
                        if (p.epi.bias2) v += __ldg(p.epi.bias2 + n);
                        if (p.epi.temb) v += __ldg(p.epi.temb + (size_t)(p.epi.temb_bcast ? 0 : b) * p.epi.temb_stride + p.epi.temb_off + n);
                    }
                    btv[te] = v;
                }
                // QUAD-view pixel offsets of rows r0 + 32 j, per tile t: pixel index * Cout (or -1)
                int off[TS_MT][4];
#pragma unroll
                for (int t = 0; t < TS_MT; ++t) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int r = r0 + 32 * jj;
                        if (p.geo) {
                            const int y = y0 + (r >> 3), x = x0 + 8 * t + (r & 7);
                            if (p.up) off[t][jj] = (y < p.H && x < p.W) ? ((b * p.epi.Ho + 2 * y + oa) * p.epi.Wo + 2 * x + ob) * Cout : -1;
                            else off[t][jj] = (y < p.H && x < p.W) ? ((b * p.H + y) * p.W + x) * Cout : -1;
                        } else {
                            const int pr = r / TP_PW, pc = r - pr * TP_PW;
                            const int y = y0 + t * TP_TH + pr, x = x0 + pc;
                            off[t][jj] = (pr < TP_TH && pc < TP_TW && y < p.H && x < p.W) ? ((b * p.H + y) * p.W + x) * Cout : -1;
                        }
                    }
                }
                const int nbase = nt * p.BN + quad * 4;
                float acc[3][8];                                  // statistics per column slot of this group
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
                int t = t_first, c = c_first, slot = bycol ? 0 : c >> 2;
                // residual that cannot come by TMA (channel count not a multiple of 4, unaligned, upsample): plain loads where it is
                // added - slow, and not a case the networks produce
                const bool rldg = resp != nullptr && !rtma;
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");           // btv visible
                mbar_wait_relaxed(afull(a), (uint32_t)(i >> 1) & 1u);
                tc_fence_after();
#pragma unroll 1
                for (int kk = 0; kk < n_mine; ++kk, ++cn) {
                    const int c0 = c * 16;
                    const uint32_t buf = cn & 1u;
                    float* const sb = stg + buf * (128 * TS_STG_LD);
                    // ---- ROW view
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * acc_cols + (uint32_t)(t * p.BN + c0), v);
                    if (kk == n_mine - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(aempty(a)) : "memory");
                    }
                    {
                        const float4* bv = reinterpret_cast<const float4*>(btv + c0);
                        float4* row = reinterpret_cast<float4*>(sb + m * TS_STG_LD);
                        if (rtma) {
                            mbar_wait_relaxed(rfull(gi, buf), (cn >> 1) & 1u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float4 ad = bv[e], rs = row[e];
                                row[e] = make_float4(__uint_as_float(v[4 * e]) + ad.x + rs.x, __uint_as_float(v[4 * e + 1]) + ad.y + rs.y,
                                                     __uint_as_float(v[4 * e + 2]) + ad.z + rs.z, __uint_as_float(v[4 * e + 3]) + ad.w + rs.w);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float4 ad = bv[e];
                                row[e] = make_float4(__uint_as_float(v[4 * e]) + ad.x, __uint_as_float(v[4 * e + 1]) + ad.y,
                                                     __uint_as_float(v[4 * e + 2]) + ad.z, __uint_as_float(v[4 * e + 3]) + ad.w);
                            }
                        }
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    // the group's next chunk
                    int tn = t, cn_ = c, slotn = slot;
                    if (bycol) {
                        cn_ += 3; ++slotn;
                        if (cn_ >= 8) { cn_ = gi; slotn = 0; ++tn; }
                    } else {
                        const int g = (t << lcpt) + c + 4;
                        tn = g >> lcpt; cn_ = g & (cpt - 1); slotn = cn_ >> 2;
                    }
                    // ---- QUAD view
                    const int n0 = nbase + c0;
                    float sm[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        float4 x = *reinterpret_cast<const float4*>(srow + buf * (128 * TS_STG_LD) + 32 * jj * TS_STG_LD);
                        const int o = t == 0 ? off[0][jj] : off[TS_MT - 1][jj];
                        if (o >= 0 && n0 < Cout) {
                            if (rldg) {
                                const float* src = resp + o + n0;
                                const float4 r4 = vec_ok ? __ldg(reinterpret_cast<const float4*>(src)) : tcs_load_tail(src, Cout - n0);
                                x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
                            }
                            if (vec_ok) {
                                if (o32) *reinterpret_cast<float4*>(o32 + o + n0) = x;
                                if (o16) {
                                    const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
                                    uint2 pk;
                                    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                                    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                                    *reinterpret_cast<uint2*>(o16 + o + n0) = pk;
                                }
                            } else {
                                tcs_store_tail(p.epi, x, o / Cout, n0);
                                if (n0 + 1 >= Cout) x.y = 0.f;
                                if (n0 + 2 >= Cout) x.z = 0.f;
                                if (n0 + 3 >= Cout) x.w = 0.f;
                            }
                            sm[0] += x.x; sm[1] += x.y; sm[2] += x.z; sm[3] += x.w;
                            sq[0] = fmaf(x.x, x.x, sq[0]); sq[1] = fmaf(x.y, x.y, sq[1]); sq[2] = fmaf(x.z, x.z, sq[2]); sq[3] = fmaf(x.w, x.w, sq[3]);
                        }
                    }
                    if (sums) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            if (k == slot) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) { acc[k][e] += sm[e]; acc[k][4 + e] += sq[e]; }
                            }
                        }
                    }
                    t = tn; c = cn_; slot = slotn;
                    // two staging buffers: the next chunk writes the other one, and its mid-chunk barrier orders this chunk's reads
                    // before the chunk after next; only the TMA refill of THIS buffer needs everyone to be done with it now
                    if (rtma) {
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                        if (te == 0) { fence_proxy_async(); issue_next(); }
                    }
                }
                // ---- fold the item's statistics: per column slot, lanes of equal quad, then one fp64 atomic pair per channel
                if (sums) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int cidx = bycol ? gi + 3 * k : (gi & (cpt - 1)) + 4 * k;       // column held by slot k
                        if (cidx >= cpt || (!bycol && k >= 2)) continue;
#pragma unroll
                        for (int o = 4; o <= 16; o <<= 1)
#pragma unroll
                            for (int e = 0; e < 8; ++e) acc[k][e] += __shfl_xor_sync(0xffffffffu, acc[k][e], o);
                        if (lane < 4) {
                            const int n0 = nt * p.BN + cidx * 16 + lane * 4;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (n0 + e < Cout && (acc[k][e] != 0.f || acc[k][4 + e] != 0.f)) {
                                    double* dst = sums + (((size_t)copy * p.epi.sums_B + b) * Cout + n0 + e) * 2;
                                    atomicAdd(dst, (double)acc[k][e]);
                                    atomicAdd(dst + 1, (double)acc[k][4 + e]);
                                }
                            }
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
    trace_end(p.trace);
}

// ------------------------------------------------------------------------------------------ weight packing
// Swizzle<B,4,3> on byte offsets inside a tile whose rows are `row_bytes` wide (B = log2(row_bytes/16)).
__host__ __device__ inline uint32_t swizzle_offset(uint32_t row, uint32_t byte_in_row, uint32_t row_bytes) {
    const uint32_t off = row * row_bytes + byte_in_row;
    const uint32_t bits = row_bytes == 128 ? 7u : (row_bytes == 64 ? 3u : 1u);
    return off ^ (((off >> 7) & bits) << 4);
}

struct PackGeom {
    int cout, cin, ks, up, nb16, KC, chunks, ntaps, nclasses, tf32;
    int cin_src;                      // input channels of the OIHW tensor (<= cin: the rest are zero weights)
};

// largest N tile (16..128) that divides the 16-padded channel count
static int tc_bn(int cout) {
    const int npad = (cout + 15) / 16 * 16;
    for (int bn = 128; bn > 16; bn >>= 1)
        if (npad % bn == 0) return bn;
    return 16;
}

static PackGeom pack_geom(int cout, int cin, int ks, int up, int kc, int tf32) {
    PackGeom g;
    g.tf32 = tf32;
    g.cin_src = cin;
    g.cout = cout; g.cin = cin; g.ks = ks; g.up = up;
    g.nb16 = (cout + 15) / 16;
    g.KC = kc;
    g.chunks = cin / kc;
    g.ntaps = up ? 4 : ks * ks;
    g.nclasses = up ? 4 : 1;
    return g;
}

// channels per K chunk: bf16: 64 / 32 / 16 channels = rows of 128 / 64 / 32 bytes; tf32: 16 / 8 channels = rows of 64 / 32 bytes
// (32-channel = 128-byte fp32 rows give wrong results in conv_tc_kernel - not in conv_tcp_kernel - on the B200, cause not found;
// the TF32 layers are the narrow latency-bound ones, where the chunk size does not matter)
int tc_pick_kc(int ca, int cb, int tf32) {
    for (int kc = tf32 ? 16 : 64; kc >= (tf32 ? 8 : 16); kc >>= 1)
        if (ca % kc == 0 && cb % kc == 0) return kc;
    return 0;
}

size_t tc_packed_weight_bytes(int cout, int cin, int ks, int tf32) {
    // worst case over the variants a layer can be packed for (plain taps vs. 4 upsample classes x 4 taps)
    const size_t npad = (size_t)(cout + 15) / 16 * 16;
    const size_t es = tf32 ? 4 : 2;
    const size_t plain = npad * ks * ks * cin * es;
    const size_t upv = ks == 3 ? (size_t)4 * npad * 4 * cin * es : 0;
    return plain > upv ? plain : upv;
}

__global__ void pack_tc_weight_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, PackGeom g) {
    const int U = g.ntaps * g.chunks;
    const size_t total = (size_t)g.nclasses * U * g.nb16 * 16 * g.KC;
    const uint32_t es = g.tf32 ? 4u : 2u;
    const uint32_t row_bytes = g.KC * es;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int k = (int)(r % g.KC); r /= g.KC;
        const int row = (int)(r % 16); r /= 16;
        const int nb = (int)(r % g.nb16); r /= g.nb16;
        const int u = (int)(r % U); r /= U;
        const int cls = (int)r;
        const int tap = u / g.chunks, c = (u - tap * g.chunks) * g.KC + k;
        const int n = nb * 16 + row;
        float v = 0.f;
        if (n < g.cout && c < g.cin_src) {
            const float* wn = w + ((size_t)n * g.cin_src + c) * g.ks * g.ks;
            if (!g.up) {
                v = wn[tap];
            } else {
                // class (a, b) = output parity; tap (ty, tx) of the 2x2 conv on the low-res input.
                // rows: a=0: ty0 <- r{0}, ty1 <- r{1,2};  a=1: ty0 <- r{0,1}, ty1 <- r{2}   (same for columns)
                const int a = cls >> 1, b = cls & 1, ty = tap >> 1, tx = tap & 1;
                const int r_lo = a == 0 ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2);
                const int r_hi = a == 0 ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
                const int s_lo = b == 0 ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2);
                const int s_hi = b == 0 ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
                for (int rr = r_lo; rr <= r_hi; ++rr)
                    for (int ss = s_lo; ss <= s_hi; ++ss) v += wn[rr * 3 + ss];
            }
        }
        // a BN-row operand tile is BN/16 consecutive blocks: the swizzle has an 8-row period, so blocks are independent
        const size_t blk = (((size_t)cls * U + u) * g.nb16 + nb) * ((size_t)16 * row_bytes);
        if (g.tf32) *reinterpret_cast<float*>(out + blk + swizzle_offset(row, k * 4, row_bytes)) = to_tf32(v);
        else *reinterpret_cast<__nv_bfloat16*>(out + blk + swizzle_offset(row, k * 2, row_bytes)) = __float2bfloat16_rn(v);
    }
}

int tc_pack_conv_weight(const float* w_oihw, uint8_t* packed, int cout, int cin, int ks, int up, int kc, int tf32, cudaStream_t st,
                        int cin_src) {
    DS_REQUIRE(tf32 ? (kc == 8 || kc == 16 || kc == 32) : (kc == 16 || kc == 32 || kc == 64), "tc_pack: KC %d", kc);
    DS_REQUIRE(cin % kc == 0, "tc_pack: cin %d not a multiple of KC %d", cin, kc);
    DS_REQUIRE(cin_src >= 0 && cin_src <= cin, "tc_pack: %d source channels > %d", cin_src, cin);
    PackGeom g = pack_geom(cout, cin, ks, up, kc, tf32);
    if (cin_src) g.cin_src = cin_src;
    const size_t total = (size_t)g.nclasses * g.ntaps * g.chunks * g.nb16 * 16 * g.KC;
    int blocks = (int)((total + 255) / 256 > 2048 ? 2048 : (total + 255) / 256);
    pack_tc_weight_kernel<<<blocks, 256, 0, st>>>(w_oihw, packed, g);
    DS_CHECK_LAUNCH("pack_tc_weight");
    return DS_OK;
}

// ------------------------------------------------------------------------------------------ host launch
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int get_encoder() {
    if (g_encode) return DS_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DS_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    DS_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return DS_OK;
}

static int encode_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int B, size_t sx, size_t sy, size_t sb, int kc,
                      int tw, int th, int tb, int tf32) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)sx, (cuuint64_t)sy, (cuuint64_t)sb};
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tb};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const int rowb = kc * (tf32 ? 4 : 2);
    const CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = g_encode(m, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for C=%d W=%d H=%d B=%d box=(%d,%d,%d,%d)", (int)r, C, W, H, B, kc, tw, th, tb);
        return DS_ERR_CUDA;
    }
    return DS_OK;
}

bool tc_conv_shape_supported(int ca, int cb, int ks, int stride, int up, int Hs, int Ws, int tf32) {
    if (tc_pick_kc(ca, cb, tf32) == 0 || ca <= 0) return false;
    if (!(ks == 1 || ks == 3)) return false;
    if (stride == 2 && (ks != 3 || up || (Hs & 1) || (Ws & 1))) return false;
    if (up && (ks != 3 || stride != 1)) return false;
    return true;
}

static int pow2_floor(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}

// does tc_build_conv pick the persistent kernel (conv_tcs_kernel) for this layer?  geo / items: its tile geometry
static bool tcs_selected(int ca, int cb, int Hs, int Ws, int B, int cout, int ks, int stride, int up, int tf32, int* geo_out, int64_t* items_out) {
    static int patch_env = -1, persist_env = -1, persist_min = -1, geo_env = -1;
    if (patch_env < 0) { const char* e = getenv("DIFFSPLIT_B200_TC_PATCH"); patch_env = e ? atoi(e) : 1; }
    if (persist_env < 0) { const char* e2 = getenv("DIFFSPLIT_B200_TC_PERSIST"); persist_env = e2 ? atoi(e2) : 1; }
    // fewest work items for the persistent kernel (one per CTA below 148): fewer still and the one-tile-per-CTA kernels,
    // which cut smaller tiles, spread the work over more SMs
    if (persist_min < 0) { const char* e4 = getenv("DIFFSPLIT_B200_TC_PERSIST_MIN"); persist_min = e4 ? atoi(e4) : 64; }
    if (geo_env < 0) { const char* e3 = getenv("DIFFSPLIT_B200_TC_GEO"); geo_env = e3 ? atoi(e3) : 1; }
    const int kc = tc_pick_kc(ca, cb, tf32);
    const size_t e = tf32 ? 4 : 2;
    const int geo = geo_env ? 1 : 0;
    const int tiles_x = (Ws + TP_TW - 1) / TP_TW;
    const int bn = tc_bn(cout);
    const int ty2 = geo ? (Hs + 15) / 16 : (Hs + TP_TH * TS_MT - 1) / (TP_TH * TS_MT);
    const int64_t items = (int64_t)B * tiles_x * ty2 * (((cout + 15) / 16 * 16) / bn) * (up ? 4 : 1);
    const bool small_idx = (int64_t)B * Hs * Ws * ((cout + 3) / 4 * 4) * (up ? 4 : 1) < (1ll << 31);
    const bool can = ks == 3 && stride == 1 && Ws >= 8 && (!up || geo);
    if (geo_out) *geo_out = geo;
    if (items_out) *items_out = items;
    return kc > 0 && can && patch_env != 0 && persist_env != 0 && (size_t)kc * e >= 64 && small_idx && 2 * TS_MT * bn <= 512 &&
           (persist_env == 2 || items >= persist_min);
}

bool tc_conv_persistent(int ca, int cb, int Hs, int Ws, int B, int cout, int ks, int stride, int up, int tf32) {
    if (!tc_conv_shape_supported(ca, cb, ks, stride, up, Hs, Ws, tf32)) return false;
    const int kc = tc_pick_kc(ca, cb, tf32);
    const size_t e = tf32 ? 4 : 2;
    int geo = 0;
    if (!tcs_selected(ca, cb, Hs, Ws, B, cout, ks, stride, up, tf32, &geo, nullptr)) return false;
    // at least three weight stages must fit beside the patches and the staging buffers
    const int prows = geo ? TP_PW * 18 : TP_PW * (TP_TH * TS_MT + 2) + 8;
    const size_t patch_bytes = align_up((size_t)prows * kc * e, 1024), wstage = align_up((size_t)tc_bn(cout) * kc * e, 1024);
    return (225 * 1024 - 2 * patch_bytes - 3 * TS_STG_BYTES - 2048) / wstage >= 3;
}

int tc_build_conv(TcConvPlan* plan, const void* src_a, int ca, const void* src_b, int cb, int Hs, int Ws, int B, int cout, int ks,
                  int stride, int up, int tf32, const void* xsrc_a, int xca, const void* xsrc_b, int xcb) {
    int rc = get_encoder();
    if (rc != DS_OK) return rc;
    DS_REQUIRE(tc_conv_shape_supported(ca, cb, ks, stride, up, Hs, Ws, tf32), "tc conv: unsupported shape");
    plan->patch = 0;
    const int kc = tc_pick_kc(ca, cb, tf32);
    const size_t e = tf32 ? 4 : 2;                                  // bytes per source element
    {
        // tall-patch variant: 3x3 stride-1 layers in the throughput regime (DIFFSPLIT_B200_TC_PATCH: 0 never, 2 whenever possible)
        static int patch_env = -1;
        if (patch_env < 0) { const char* e = getenv("DIFFSPLIT_B200_TC_PATCH"); patch_env = e ? atoi(e) : 1; }
        const int tiles_x = (Ws + TP_TW - 1) / TP_TW;
        int mt = (Hs + TP_TH - 1) / TP_TH;
        if (mt > TP_MAX_MT) mt = TP_MAX_MT;
        const int bn = tc_bn(cout);
        if (mt * bn > 512) mt = 512 / bn;
        const int tiles_y = (Hs + TP_TH * mt - 1) / (TP_TH * mt);
        const int64_t ctas = (int64_t)B * tiles_x * tiles_y * (((cout + 15) / 16 * 16) / bn);
        const bool can = ks == 3 && stride == 1 && !up && Ws >= 8 && mt >= 1;
        {
            int geo = 0;
            int64_t items = 0;
            if (tcs_selected(ca, cb, Hs, Ws, B, cout, ks, stride, up, tf32, &geo, &items)) {
                const int ty2 = geo ? (Hs + 15) / 16 : (Hs + TP_TH * TS_MT - 1) / (TP_TH * TS_MT);
                TcpParams& q = *reinterpret_cast<TcpParams*>(plan->params);
                memset(&q, 0, sizeof(q));
                q.B = B; q.H = Hs; q.W = Ws; q.MT = TS_MT; q.tiles_x = tiles_x; q.tiles_y = ty2;
                q.epi.Ho = up ? 2 * Hs : Hs; q.epi.Wo = up ? 2 * Ws : Ws; q.epi.Cout = cout;
                q.nb16 = (cout + 15) / 16;
                q.BN = bn; q.n_tiles = (q.nb16 * 16) / bn;
                q.KC = kc; q.chunks_a = ca / kc; q.chunks_b = cb / kc;
                q.tf32 = tf32;
                q.n_items = (int)items;
                q.geo = geo;
                q.up = up ? 1 : 0;
                q.class_bytes = (uint32_t)((size_t)4 * (ca + cb) / kc * q.nb16 * 16 * kc * e);
                q.div_ntiles = make_fastdiv((uint32_t)(q.n_tiles * (up ? 4 : 1)));
                q.div_tiles_x = make_fastdiv((uint32_t)tiles_x);
                q.div_tiles_xy = make_fastdiv((uint32_t)(tiles_x * ty2));
                const int prows = geo ? TP_PW * 18 : TP_PW * (TP_TH * TS_MT + 2) + 8;
                q.patch_bytes = (uint32_t)align_up((size_t)prows * kc * e, 1024);
                const uint32_t wstage = (uint32_t)align_up((size_t)bn * kc * e, 1024);
                // four epilogue groups when their staging areas leave at least five weight stages, else three (with three groups the
                // chunks are dealt by column, which needs eight columns per tile)
                int ng = 4;
                int wst = (int)((225 * 1024 - 2 * (size_t)q.patch_bytes - 4 * TS_STG_BYTES - 2048) / wstage);
                static int ng_env = -1;
                if (ng_env < 0) { const char* e6 = getenv("DIFFSPLIT_B200_TC_NG"); ng_env = e6 ? atoi(e6) : 0; }
                if ((ng_env == 3 || (ng_env != 4 && wst < 5) || wst < 3) && bn == 128) {
                    ng = 3;
                    wst = (int)((225 * 1024 - 2 * (size_t)q.patch_bytes - 3 * TS_STG_BYTES - 2048) / wstage);
                }
                q.ngroups = ng;
                if (wst > 12) wst = 12;
                if (wst >= 3) {
                    q.wstages = wst;
                    for (int s_ = 0; s_ < 2; ++s_) {
                        const void* ptr = s_ == 0 ? src_a : src_b;
                        const int C = s_ == 0 ? ca : cb;
                        if (!ptr || C == 0) continue;
                        rc = encode_map(&q.pmap[s_], ptr, C, Ws, Hs, B, (size_t)C * e, (size_t)Ws * C * e, (size_t)Hs * Ws * C * e, kc, TP_PW,
                                        geo ? 18 : TP_TH * TS_MT + 2, 1, tf32);
                        if (rc != DS_OK) return rc;
                    }
                    if (xsrc_a) {
                        DS_REQUIRE(!up && geo && xca > 0 && tc_pick_kc(xca, xcb, tf32) == kc && (xcb == 0 || xsrc_b),
                                   "tc conv: folded 1x1 input %d+%d does not match the layer's K chunk %d", xca, xcb, kc);
                        for (int s_ = 0; s_ < 2; ++s_) {
                            const void* ptr = s_ == 0 ? xsrc_a : xsrc_b;
                            const int C = s_ == 0 ? xca : xcb;
                            if (!ptr || C == 0) continue;
                            rc = encode_map(&q.xmap[s_], ptr, C, Ws, Hs, B, (size_t)C * e, (size_t)Ws * C * e, (size_t)Hs * Ws * C * e, kc, 16, 16, 1,
                                            tf32);
                            if (rc != DS_OK) return rc;
                        }
                        q.xchunks_a = xca / kc; q.xchunks_b = xcb / kc;
                    }
                    plan->patch = 2;
                    plan->smem_bytes = (int)(2 * (size_t)q.patch_bytes + (size_t)wst * wstage + ng * TS_STG_BYTES + 16 * wst + 192 + 1024);
                    static int sms = 0;
                    if (!sms) {
                        int dev = 0;
                        DS_CHECK_CUDA(cudaGetDevice(&dev));
                        DS_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
                    }
                    plan->grid_x = items < sms ? (int)items : sms;
                    plan->grid_y = 1;
                    plan->grid_z = 1;
                    plan->kc = kc;
                    return DS_OK;
                }
            }
        }
        if (can && (patch_env == 2 || (patch_env == 1 && ctas >= 148 && (size_t)(ca + cb) * e >= 128 && (size_t)kc * e >= 64))) {
            static_assert(sizeof(TcpParams) <= sizeof(plan->params), "TcConvPlan::params too small");
            TcpParams& q = *reinterpret_cast<TcpParams*>(plan->params);
            memset(&q, 0, sizeof(q));
            q.B = B; q.H = Hs; q.W = Ws; q.MT = mt; q.tiles_x = tiles_x; q.tiles_y = tiles_y;
            q.epi.Ho = Hs; q.epi.Wo = Ws; q.epi.Cout = cout;
            q.nb16 = (cout + 15) / 16;
            q.BN = bn; q.n_tiles = (q.nb16 * 16) / bn;
            q.KC = kc; q.chunks_a = ca / kc; q.chunks_b = cb / kc;
            q.tf32 = tf32;
            const int prows = TP_PW * (TP_TH * mt + 2) + 8;             // + the rows the last tile's window runs past the patch
            q.patch_bytes = (uint32_t)align_up((size_t)prows * kc * e, 1024);
            const uint32_t wstage = (uint32_t)align_up((size_t)bn * kc * e, 1024);
            int wst = (int)((196 * 1024 - 2 * (size_t)q.patch_bytes - 2048 - TP_GROUPS * TC_RED_BYTES) / wstage);
            if (wst > 9) wst = 9;
            if (wst >= 2) {
                q.wstages = wst;
                for (int s_ = 0; s_ < 2; ++s_) {
                    const void* ptr = s_ == 0 ? src_a : src_b;
                    const int C = s_ == 0 ? ca : cb;
                    if (!ptr || C == 0) continue;
                    rc = encode_map(&q.pmap[s_], ptr, C, Ws, Hs, B, (size_t)C * e, (size_t)Ws * C * e, (size_t)Hs * Ws * C * e, kc, TP_PW,
                                    TP_TH * mt + 2, 1, tf32);
                    if (rc != DS_OK) return rc;
                }
                plan->patch = 1;
                plan->smem_bytes = (int)(2 * (size_t)q.patch_bytes + (size_t)wst * wstage + 16 * wst + 80 + 1024 + TP_GROUPS * TC_RED_BYTES);
                plan->grid_x = tiles_x * tiles_y * B;
                plan->grid_y = q.n_tiles;
                plan->grid_z = 1;
                plan->kc = kc;
                return DS_OK;
            }
        }
    }
    DS_REQUIRE(!xsrc_a, "tc conv: a folded 1x1 input needs the persistent kernel (check tc_conv_persistent first)");
    TcParams& p = *reinterpret_cast<TcParams*>(plan->params);
    static_assert(sizeof(TcParams) <= sizeof(plan->params), "TcConvPlan::params too small");
    memset(&p, 0, sizeof(p));
    // tile space
    const int H = stride == 2 ? Hs / 2 : Hs, W = stride == 2 ? Ws / 2 : Ws;
    p.B = B; p.H = H; p.W = W;
    p.epi.Ho = up ? 2 * H : H; p.epi.Wo = up ? 2 * W : W;
    p.up = up;
    p.tf32 = tf32;
    // 128-pixel tile: pick the power-of-two width with the least padded columns, prefer wide
    int best_tw = 1, best_waste = 1 << 30;
    for (int tw = 128; tw >= 1; tw >>= 1) {
        if (tw > 2 * pow2_floor(W) && tw > 1) continue;
        const int waste = ((W + tw - 1) / tw) * tw - W;
        if (waste < best_waste) { best_waste = waste; best_tw = tw; }
    }
    p.tw = best_tw;
    int th = 128 / p.tw;
    const int hp = pow2_floor(H) < H ? 2 * pow2_floor(H) : H;       // smallest power of two >= H
    if (th > hp) th = hp;
    p.th = th;
    p.tb = 128 / (p.tw * p.th);
    p.tiles_x = (W + p.tw - 1) / p.tw;
    p.tiles_y = (H + p.th - 1) / p.th;
    const int tiles_b = (B + p.tb - 1) / p.tb;
    p.epi.Cout = cout;
    p.nb16 = (cout + 15) / 16;
    // N tile: as wide as divides the padded channel count, narrowed until the grid covers the machine
    int bn = tc_bn(cout);
    const int m_tiles = p.tiles_x * p.tiles_y * tiles_b * (up ? 4 : 1);
    while (bn > 16 && m_tiles * ((p.nb16 * 16) / bn) < 120) bn >>= 1;
    p.BN = bn;
    p.n_tiles = (p.nb16 * 16) / bn;
    p.KC = kc;
    p.chunks_a = ca / kc;
    p.chunks_b = cb / kc;
    p.ntaps = up ? 4 : ks * ks;
    p.nclasses = up ? 4 : 1;
    for (int cls = 0; cls < p.nclasses; ++cls) {
        for (int t = 0; t < p.ntaps; ++t) {
            int dy, dx, view = 0;
            if (up) {
                const int a = cls >> 1, b = cls & 1, ty = t >> 1, tx = t & 1;
                dy = a == 0 ? ty - 1 : ty;
                dx = b == 0 ? tx - 1 : tx;
            } else if (stride == 2) {
                const int r = t / 3, s = t % 3;
                // input row 2*oy + r - 1:  r=0 -> parity 1 of row oy-1;  r=1 -> parity 0 of oy;  r=2 -> parity 1 of oy
                const int py = r == 1 ? 0 : 1, px = s == 1 ? 0 : 1;
                dy = r == 0 ? -1 : 0;
                dx = s == 0 ? -1 : 0;
                view = py * 2 + px;
            } else {
                dy = ks == 3 ? t / 3 - 1 : 0;
                dx = ks == 3 ? t % 3 - 1 : 0;
            }
            p.tap_dy[cls][t] = (int8_t)dy; p.tap_dx[cls][t] = (int8_t)dx; p.tap_view[cls][t] = (int8_t)view;
        }
    }
    // tensor maps
    for (int s = 0; s < 2; ++s) {
        const void* ptr = s == 0 ? src_a : src_b;
        const int C = s == 0 ? ca : cb;
        if (!ptr || C == 0) continue;
        if (stride == 2) {
            for (int v = 0; v < 4; ++v) {
                const int py = v >> 1, px = v & 1;
                const uint8_t* bp = (const uint8_t*)ptr + ((size_t)py * Ws + px) * C * e;
                rc = encode_map(&p.maps[s * 4 + v], bp, C, Ws / 2, Hs / 2, B, 2 * (size_t)C * e, 2 * (size_t)Ws * C * e,
                                (size_t)Hs * Ws * C * e, kc, p.tw, p.th, p.tb, tf32);
                if (rc != DS_OK) return rc;
            }
        } else {
            rc = encode_map(&p.maps[s * 4], ptr, C, Ws, Hs, B, (size_t)C * e, (size_t)Ws * C * e, (size_t)Hs * Ws * C * e, kc, p.tw,
                            p.th, p.tb, tf32);
            if (rc != DS_OK) return rc;
        }
    }
    // pipeline depth: every (tap, chunk) unit is one TMA round trip (~1 us), so small units need many slots in flight:
    // up to 16 stages within 72 KB (3 CTAs / SM) for small tiles, up to 160 KB (1 CTA / SM) for the large ones
    const uint32_t stage_bytes = (uint32_t)align_up((128u + (uint32_t)p.BN) * kc * (uint32_t)e, 1024);
    const int U = p.ntaps * (p.chunks_a + p.chunks_b);
    const uint32_t budget = stage_bytes <= 8192 ? 73728u : 163840u;
    int stages = (int)(budget / stage_bytes);
    if (stages > 16) stages = 16;
    if (stages > U) stages = U;
    if (stages < 1) stages = 1;
    p.stages = stages;
    plan->smem_bytes = stages * stage_bytes + 16 * stages + 48 + 1024 + TC_RED_BYTES;
    plan->grid_x = p.tiles_x * p.tiles_y * tiles_b;
    plan->grid_y = p.n_tiles;
    plan->grid_z = p.nclasses;
    plan->kc = kc;
    return DS_OK;
}

int tc_launch_conv(const TcConvPlan* plan, const uint8_t* w_packed, const ConvEpi& epi, float* out_f32, void* out_b16,
                   float* out_nchw, double* sums_out, cudaStream_t st, const uint8_t* xw_packed, const float* bias2) {
    DS_REQUIRE(plan->patch == 2 || !xw_packed, "tc conv: folded 1x1 weights without the persistent kernel");
    if (plan->patch == 2) {
        TcpParams q = *reinterpret_cast<const TcpParams*>(plan->params);
        static_assert(sizeof(TcpParams) <= sizeof(plan->params), "TcConvPlan::params too small");
        DS_REQUIRE((q.xchunks_a + q.xchunks_b > 0) == (xw_packed != nullptr), "tc conv: plan / launch disagree about the folded 1x1 input");
        q.w = w_packed;
        q.xw = xw_packed;
        q.epi.bias2 = bias2;
        q.epi.bias = epi.bias; q.epi.temb = epi.temb; q.epi.temb_off = epi.temb_off; q.epi.temb_stride = epi.temb_stride;
        q.epi.temb_bcast = epi.temb_bcast; q.epi.residual = epi.residual;
        q.epi.out_f32 = out_f32; q.epi.out_b16 = reinterpret_cast<__nv_bfloat16*>(out_b16); q.epi.out_nchw = out_nchw;
        q.epi.sums_out = sums_out; q.epi.sums_B = q.B;
        q.trace = trace_next(4);
        q.res_tma = 0;
        const int cout = q.epi.Cout;
        static int respf_env = -1;
        if (respf_env < 0) { const char* e5 = getenv("DIFFSPLIT_B200_TC_RESPF"); respf_env = e5 ? atoi(e5) : 1; }
        if (respf_env && epi.residual && !q.up && q.geo && (cout * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(epi.residual) & 15) == 0) {
            cuuint64_t dims[4] = {(cuuint64_t)cout, (cuuint64_t)q.W, (cuuint64_t)q.H, (cuuint64_t)q.B};
            cuuint64_t strides[3] = {(cuuint64_t)cout * 4, (cuuint64_t)q.W * cout * 4, (cuuint64_t)q.H * q.W * cout * 4};
            cuuint32_t box[4] = {TS_STG_LD, 8, 16, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            if (g_encode(&q.resmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(epi.residual), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                q.res_tma = 1;
        }
        static bool sattr = false;
        if (!sattr) {
            DS_CHECK_CUDA(cudaFuncSetAttribute(conv_tcs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            sattr = true;
        }
        DS_CHECK_CUDA(launch_pdl(conv_tcs_kernel, dim3(plan->grid_x, 1, 1), dim3(TS_THREADS), (size_t)plan->smem_bytes, st, q));
        return DS_OK;
    }
    if (plan->patch) {
        TcpParams q = *reinterpret_cast<const TcpParams*>(plan->params);
        q.w = w_packed;
        q.epi.bias = epi.bias; q.epi.temb = epi.temb; q.epi.temb_off = epi.temb_off; q.epi.temb_stride = epi.temb_stride;
        q.epi.temb_bcast = epi.temb_bcast; q.epi.residual = epi.residual;
        q.epi.out_f32 = out_f32; q.epi.out_b16 = reinterpret_cast<__nv_bfloat16*>(out_b16); q.epi.out_nchw = out_nchw;
        q.epi.sums_out = sums_out; q.epi.sums_B = q.B;
        q.trace = trace_next(4);
        static bool pattr = false;
        if (!pattr) {
            DS_CHECK_CUDA(cudaFuncSetAttribute(conv_tcp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            pattr = true;
        }
        DS_CHECK_CUDA(launch_pdl(conv_tcp_kernel, dim3(plan->grid_x, plan->grid_y, 1), dim3(TP_THREADS), (size_t)plan->smem_bytes, st, q));
        return DS_OK;
    }
    TcParams p = *reinterpret_cast<const TcParams*>(plan->params);
    p.w = w_packed;
    p.epi.bias = epi.bias;
    p.epi.temb = epi.temb;
    p.epi.temb_off = epi.temb_off;
    p.epi.temb_stride = epi.temb_stride;
    p.epi.temb_bcast = epi.temb_bcast;
    p.epi.residual = epi.residual;
    p.epi.out_f32 = out_f32;
    p.epi.out_b16 = reinterpret_cast<__nv_bfloat16*>(out_b16);
    p.epi.out_nchw = out_nchw;
    p.epi.sums_out = sums_out;
    p.epi.sums_B = p.B;
    p.trace = trace_next(4);
    static bool attr_set = false;
    if (!attr_set) {
        DS_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    DS_CHECK_CUDA(launch_pdl(conv_tc_kernel, dim3(plan->grid_x, plan->grid_y, plan->grid_z), dim3(TC_THREADS),
                             (size_t)plan->smem_bytes, st, p));
    return DS_OK;
}

// ------------------------------------------------------------------------------------------ network input -> TF32 NHWC, 8 channels
// The first conv of the net reads the caller's fp32 NCHW tensors (condition and x_t: the torch.cat of p_mean_variance).  With
// <= 8 channels in total they are repacked once into an NHWC tensor of 8 or 16 channels (zero padded, values rounded to TF32), so the
// conv runs on the tensor cores as a TF32 implicit GEMM (16 channels = 64-byte rows: the persistent kernel) instead of on the CUDA cores.
template <int CP>
__global__ void pack_input_kernel(const float* __restrict__ xa, int ca, const float* __restrict__ xb, int cb, float* __restrict__ out,
                                  int B, int HW) {
    pdl_wait();
    pdl_trigger();
    const size_t total = (size_t)B * HW;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / HW, px = i - b * HW;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float x = 0.f;
            if (c < ca) x = __ldg(xa + (b * ca + c) * HW + px);
            else if (c < ca + cb) x = __ldg(xb + (b * cb + (c - ca)) * HW + px);
            v[c] = to_tf32(x);
        }
        float4* dst = reinterpret_cast<float4*>(out + i * CP);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
        for (int c = 2; c < CP / 4; ++c) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

int tc_pack_input(const float* xa, int ca, const float* xb, int cb, float* out, int cpad, int B, int H, int W, cudaStream_t st) {
    DS_REQUIRE(xa && ca > 0 && cb >= 0 && ca + cb <= 8 && (cb == 0 || xb) && out && (cpad == 8 || cpad == 16), "pack_input: %d+%d -> %d channels",
               ca, cb, cpad);
    const size_t total = (size_t)B * H * W;
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    if (cpad == 8) DS_CHECK_CUDA(launch_pdl(pack_input_kernel<8>, dim3(blocks), dim3(256), 0, st, xa, ca, xb, cb, out, B, H * W));
    else DS_CHECK_CUDA(launch_pdl(pack_input_kernel<16>, dim3(blocks), dim3(256), 0, st, xa, ca, xb, cb, out, B, H * W));
    return DS_OK;
}

}  // namespace ds
