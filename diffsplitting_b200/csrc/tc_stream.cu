// Persistent, warp-specialised  GroupNorm-apply + Swish -> 3x3 convolution  for the THROUGHPUT regime (many waves of tiles):
// the software-pipelined successor of conv_halo_kernel (tc_halo.cu) for layers whose whole weight image fits in shared
// memory next to the pipeline buffers (9 C BN es <= ~100 KB: the 16..64-channel layers of the 256^2 / 512^2 levels).
//
// conv_halo_kernel runs  load -> normalise -> MMA -> epilogue  strictly in sequence inside a CTA and re-fetches the weights
// for every 112-output tile (74 KB of weights for 43 KB of activations at 64 -> 64); the 64-channel 512^2 layers of
// sr_sr3_64_512 ran at 135-156 TFLOP/s = 1.1-1.4 TB/s, a third of what HBM allows.  Here ONE CTA per SM keeps the weights
// resident and walks tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...  through four concurrently running stages:
//
//   warp 0            TMA producer   raw fp32 patch (7+2) x (16+2) x C of tile i+2 -> RAW ring (out-of-image = zero fill)
//   warps 2-3, 12-15  transform      RAW[i+1]: scale/shift (GroupNorm) + Swish -> bf16 | tf32 -> OP[(i+1)&1], the no-swizzle
//                                    K-major operand image [channel plane][patch pixel][16 B] (padding pixels = zeros)
//   warp 1            MMA issuer     OP[i&1]: 9 taps x C/16 tcgen05.mma, a tap = a descriptor start shifted by (18 r + s)
//                                    pixels; accumulators ACC[i&1] in TMEM (two buffers)
//   warps 4-7, 8-11   epilogue       ACC[(i-1)&1]: two groups of four warps (one warp per TMEM lane quadrant each) take the even
//                                    / odd 16-channel chunks: tcgen05.ld, + bias + conditioning vector + fp32 residual
//                                    (requested four chunks ahead), fp32 / bf16 NHWC or fp32 NCHW stores, GroupNorm statistics
//                                    of the output for the consumer.  The epilogue is instruction-issue bound (~150
//                                    instructions per warp and chunk): eight warps, not four (first version: 1.10 ms for the
//                                    64 -> 64 layer at 8 x 512^2, no faster than conv_halo_kernel, 7 % tensor-pipe activity)
//
// linked by mbarriers (TMA complete_tx, tcgen05.commit, warp arrivals).  Geometry, operand layout, weight pack and epilogue
// are those of conv_halo_kernel's 2-D tiles, so the two kernels are interchangeable per layer (halo_launch_conv picks).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "tc.cuh"
#include "tc_ptx.cuh"

namespace ds {

constexpr int SK_THREADS = 512;
constexpr int SK_XF_WARPS = 6, SK_XF_THREADS = SK_XF_WARPS * 32;      // warps 2, 3, 12 .. 15
constexpr int SK_EPI_GROUPS = 2;                                       // warps 4-7 and 8-11
constexpr int SK_TW = 16, SK_TH = 7, SK_PW = SK_TW + 2, SK_PH = SK_TH + 2;
constexpr int SK_PATCH_PX = SK_PW * SK_PH;               // 162 staged pixels
constexpr int SK_PLANE_PX = 169;                         // + the rows the last window runs past the patch; 169 * 16 B = 16 mod 128:
                                                         // the 8 planes a quarter-warp stores to hit 8 different bank groups
constexpr int SK_NRAW = 2;
constexpr int SK_STG_LD = 20;                            // row pitch (floats) of the epilogue staging buffer: conflict-free STS.128
constexpr uint32_t SK_RED_BYTES = 128 * SK_STG_LD * 4;                     // one staging buffer; two per epilogue group
constexpr uint32_t SK_RES_BOX_BYTES = SK_STG_LD * SK_PW * SK_TH * 4;       // residual box {20 ch, 18, 7}: rows of the staging layout
constexpr size_t SK_SMEM_LIMIT = 225 * 1024;

struct StreamParams {
    CUtensorMap wmap;                 // packed weights (halo_pack_conv_weight image), as in conv_halo_kernel
    CUtensorMap rmap[2];              // raw fp32 sources as (C, W, H, B) tensors, box {c, 18, 9, 1}
    CUtensorMap resmap;               // the fp32 residual as a (Cout, W, H, B) tensor, box {20, 18, 7, 1} = one chunk in the staging layout
    int ca, cb, C;
    const double* sums_a; const double* sums_b;       // per-channel fp64 (sum, sumsq) [copies][B][c][2], or
    const float2* stats;                               // (mean, rstd) [B][G]; all null: no normalisation
    const float* gamma; const float* beta;
    int G, swish;
    TcEpi epi;
    int B, H, W;
    int ksteps, BN, wloads;
    int tiles_x, tiles_y, n_tiles_m;
    FastDiv div_tiles_x, div_tiles_xy;
    uint32_t raw_bytes, raw_a_bytes, op_bytes, b_bytes;
    int res_prefetch;                 // resmap is valid: the residual arrives by TMA in the staging buffers
    long long* dbg;                   // DIFFSPLIT_B200_STREAM_DBG: per-role wait / work cycle sums of CTA 0 (16 slots)
    TraceSlot trace;
};
#define SK_T0() const long long _t0 = p.dbg ? clock64() : 0
#define SK_ACC(slot) do { if (p.dbg) acc_dbg[slot] += clock64() - _t0; } while (0)

// rare epilogue paths kept out of line: fp32 NCHW output (the network's last conv) and channel counts that are not multiples of 4
__device__ __noinline__ void stream_store_tail(const TcEpi& e, float4 x, int pix, int n0, int HW) {
    const float xs[4] = {x.x, x.y, x.z, x.w};
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

__device__ __noinline__ float4 stream_load_tail(const float* src, int n) {
    return make_float4(__ldg(src), n > 1 ? __ldg(src + 1) : 0.f, n > 2 ? __ldg(src + 2) : 0.f, n > 3 ? __ldg(src + 3) : 0.f);
}

template <bool TF32>
__global__ void __launch_bounds__(SK_THREADS, 1) conv_stream_kernel(const __grid_constant__ StreamParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int CPP = TF32 ? 4 : 8;                    // channels per 16-byte plane row
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t base = (raw0 + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - raw0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P = p.C / CPP;
    constexpr uint32_t plane_bytes = SK_PLANE_PX * 16u;
    // shared-memory map
    const uint32_t w_off = 0;
    const uint32_t op_off = (w_off + p.b_bytes + 127u) & ~127u;
    const uint32_t raw_off = (op_off + 2u * p.op_bytes + 127u) & ~127u;
    const uint32_t tab_off = (raw_off + SK_NRAW * p.raw_bytes + 15u) & ~15u;
    const uint32_t bt_off = (tab_off + (uint32_t)p.B * p.C * 8u + 15u) & ~15u;       // [B][BN] bias + conditioning vector
    const uint32_t bar_off = (bt_off + (uint32_t)p.B * p.BN * 4u + 15u) & ~15u;
    const uint32_t bars = base + bar_off;
    auto full_raw = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty_raw = [&](int s) { return bars + 8u * (uint32_t)(SK_NRAW + s); };
    auto full_op = [&](int s) { return bars + 8u * (uint32_t)(2 * SK_NRAW + s); };
    auto empty_op = [&](int s) { return bars + 8u * (uint32_t)(2 * SK_NRAW + 2 + s); };
    auto full_acc = [&](int s) { return bars + 8u * (uint32_t)(2 * SK_NRAW + 4 + s); };
    auto empty_acc = [&](int s) { return bars + 8u * (uint32_t)(2 * SK_NRAW + 6 + s); };
    const uint32_t wfull = bars + 8u * (uint32_t)(2 * SK_NRAW + 8), tmem_slot = wfull + 8u;
    auto rfull = [&](int g, uint32_t buf) { return bars + 8u * (uint32_t)(2 * SK_NRAW + 10) + 16u * (uint32_t)g + 8u * buf; };
    uint8_t* red = gbase + ((bar_off + 8u * (uint32_t)(2 * SK_NRAW + 14) + 127u) & ~127u);    // SK_EPI_GROUPS x 2 x SK_RED_BYTES, 128-byte aligned (TMA destination)
    const int nt = blockIdx.y;
    const uint32_t tmem_cols = 2u * (uint32_t)p.BN <= 32u ? 32u : (2u * (uint32_t)p.BN <= 64u ? 64u : (2u * (uint32_t)p.BN <= 128u ? 128u : (2u * (uint32_t)p.BN <= 256u ? 256u : 512u)));
    const int my_tiles = ((int)blockIdx.x < p.n_tiles_m) ? (p.n_tiles_m - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    long long acc_dbg[4] = {0, 0, 0, 0};
    const long long t_start = p.dbg ? clock64() : 0;
    trace_begin(p.trace);
    if (warp == 0 && elect_one()) {
        for (int s = 0; s < SK_NRAW; ++s) { mbar_init(full_raw(s), 1); mbar_init(empty_raw(s), SK_XF_WARPS); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(full_op(s), SK_XF_WARPS); mbar_init(empty_op(s), 1);
            mbar_init(full_acc(s), 1); mbar_init(empty_acc(s), 4 * ((p.BN >> 4) >= SK_EPI_GROUPS ? SK_EPI_GROUPS : 1));   // BN = 16: one group per buffer
        }
        mbar_init(wfull, 1);
        for (int g = 0; g < SK_EPI_GROUPS; ++g) { mbar_init(rfull(g, 0), 1); mbar_init(rfull(g, 1), 1); }
        fence_barrier_init();
        // weights: constant data, fetched before the predecessor kernel has finished
        mbar_expect_tx(wfull, p.b_bytes);
        const int per = 9 * p.ksteps * 2 / p.wloads;
        for (int l = 0; l < p.wloads; ++l)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                    base + w_off + (uint32_t)l * (uint32_t)per * (uint32_t)p.BN * 16u),
                "l"(reinterpret_cast<uint64_t>(&p.wmap)), "r"(wfull), "r"(0), "r"(nt * (p.BN >> 4)), "r"(l * per)
                : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    // the operand rows behind the patch (read by the last windows of rows 126, 127: discarded outputs) must hold finite values
    for (uint32_t i = tid; i < 2u * (uint32_t)P * (SK_PLANE_PX - SK_PATCH_PX); i += SK_THREADS) {
        const uint32_t s = i / ((uint32_t)P * (SK_PLANE_PX - SK_PATCH_PX)), r = i - s * (uint32_t)P * (SK_PLANE_PX - SK_PATCH_PX);
        const uint32_t pl = r / (SK_PLANE_PX - SK_PATCH_PX), px = SK_PATCH_PX + (r - pl * (SK_PLANE_PX - SK_PATCH_PX));
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + op_off + s * p.op_bytes + pl * plane_bytes + px * 16u), "r"(0u) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    pdl_wait();
    pdl_trigger();

    auto tile_coords = [&](int i, int& b, int& y0, int& x0) {
        const int t = (int)blockIdx.x + i * (int)gridDim.x;
        b = fdiv(t, p.div_tiles_xy);
        const int rem = t - b * p.tiles_x * p.tiles_y;
        const int ty = fdiv(rem, p.div_tiles_x);
        y0 = ty * SK_TH;
        x0 = (rem - ty * p.tiles_x) * SK_TW;
    };
    auto issue_raw = [&](int i) {              // one elected thread of warp 0
        const int s = i % SK_NRAW;
        int b, y0, x0;
        tile_coords(i, b, y0, x0);
        mbar_expect_tx(full_raw(s), p.raw_bytes);
        const uint32_t dst = base + raw_off + (uint32_t)s * p.raw_bytes;
        tma_load_4d(dst, &p.rmap[0], full_raw(s), 0, x0 - 1, y0 - 1, b);
        if (p.cb) tma_load_4d(dst + p.raw_a_bytes, &p.rmap[1], full_raw(s), 0, x0 - 1, y0 - 1, b);
    };
    if (warp == 0 && elect_one()) {            // the first patches travel while the scale / shift table is built
        for (int i = 0; i < SK_NRAW && i < my_tiles; ++i) issue_raw(i);
    }

    // ---- per (sample, channel) scale / shift of the fused GroupNorm: a = rstd * gamma, sh = beta - mean * a  (all threads)
    float2* tab = reinterpret_cast<float2*>(gbase + tab_off);
    {
        const bool norm = p.stats || p.sums_a;
        const int cpg = norm ? p.C / p.G : 1;
        double2* chs = reinterpret_cast<double2*>(gbase + op_off + p.op_bytes);      // OP[1] is idle until the second tile
        if (p.sums_a) {
            for (int i = tid; i < p.B * p.C; i += SK_THREADS) {
                const int b = i / p.C, cc = i - b * p.C;
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
        for (int i = tid; i < p.B * p.C; i += SK_THREADS) {
            const int b = i / p.C, c = i - b * p.C;
            float a = 1.f, sh = 0.f;
            if (p.sums_a) {
                const double2* cs = chs + b * p.C + c / cpg * cpg;
                double sm = 0.0, sq = 0.0;
                for (int k = 0; k < cpg; ++k) { sm += cs[k].x; sq += cs[k].y; }
                const double mu = sm * inv_cnt;
                double var = sq * inv_cnt - mu * mu;
                if (var < 0.0) var = 0.0;
                a = rsqrtf((float)var + 1e-5f) * p.gamma[c];
                sh = p.beta[c] - (float)mu * a;
            } else if (norm) {
                const float2 st = p.stats[(size_t)b * p.G + c / cpg];
                a = st.y * p.gamma[c];
                sh = p.beta[c] - st.x * a;
            }
            tab[i] = make_float2(a, sh);
        }
        // what the epilogue adds per (sample, output channel): bias + conditioning vector (zero beyond Cout)
        float* btw = reinterpret_cast<float*>(gbase + bt_off);
        for (int i = tid; i < p.B * p.BN; i += SK_THREADS) {
            const int b = i / p.BN, n = nt * p.BN + (i - b * p.BN);
            float v = 0.f;
            if (n < p.epi.Cout) {
                if (p.epi.bias) v = __ldg(p.epi.bias + n);
                if (p.epi.temb) v += __ldg(p.epi.temb + (size_t)(p.epi.temb_bcast ? 0 : b) * p.epi.temb_stride + p.epi.temb_off + n);
            }
            btw[i] = v;
        }
    }
    __syncthreads();

    if (warp == 0) {
        // ===== TMA producer: patch of tile i once the transform warps have released its ring slot
        if (elect_one()) {
            for (int i = SK_NRAW; i < my_tiles; ++i) {
                { SK_T0(); mbar_wait_relaxed(empty_raw(i % SK_NRAW), (uint32_t)((i / SK_NRAW) & 1) ^ 1u); SK_ACC(0); }
                issue_raw(i);
            }
            if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0) { p.dbg[0] = acc_dbg[0]; p.dbg[15] = clock64() - t_start; p.dbg[14] = my_tiles; }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer
        if (elect_one()) {
            mbar_wait(wfull, 0);
            const uint32_t idesc = umma_idesc(p.BN, TF32);
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);                       // SBO = 128 B, descriptor version 1
            const uint32_t b_lo0 = (((base + w_off) & 0x3FFFFu) >> 4) | ((uint32_t)p.BN << 16);      // LBO = BN * 16 B
            const uint32_t a_kstep = (2u * plane_bytes) >> 4, b_kstep = (uint32_t)p.BN * 2u;
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i & 1;
                const uint32_t ph = (uint32_t)(i >> 1) & 1u;
                { SK_T0(); mbar_wait(full_op(s), ph); SK_ACC(0); }
                { SK_T0(); mbar_wait(empty_acc(s), ph ^ 1u); SK_ACC(1); }
                tc_fence_after();
                SK_T0();
                const uint32_t a_lo0 = (((base + op_off + (uint32_t)s * p.op_bytes) & 0x3FFFFu) >> 4) | ((plane_bytes >> 4) << 16);
                const uint32_t acc = tmem_base + (uint32_t)(s * p.BN);
                // one thread issues everything: taps unrolled (constant shifts), descriptors = one 64-bit add per operand and MMA
                const uint64_t a_d0 = ((uint64_t)desc_hi << 32) | a_lo0, b_d0 = ((uint64_t)desc_hi << 32) | b_lo0;
                const int ks = p.ksteps;
                uint64_t b_d = b_d0;
                uint32_t first = 0;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    uint64_t a_d = a_d0 + (uint64_t)((tap / 3) * SK_PW + tap % 3);
#pragma unroll 2
                    for (int kk = 0; kk < ks; ++kk) {
                        umma<TF32>(acc, a_d, b_d, idesc, first);
                        first = 1u;
                        a_d += a_kstep;
                        b_d += b_kstep;
                    }
                }
                umma_commit(empty_op(s));         // the operand image may be overwritten once these MMAs have read it
                umma_commit(full_acc(s));         // accumulator complete
                SK_ACC(2);
            }
            if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0) { p.dbg[1] = acc_dbg[0]; p.dbg[2] = acc_dbg[1]; p.dbg[3] = acc_dbg[2]; }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 12) {
        // ===== epilogue: group gi = chunks gi, gi + 2 (BN <= 64: at most two chunks per group and tile), two phases per chunk:
        //  1. ROW view - warp q owns TMEM lanes [32q, 32q + 32) = patch positions: tcgen05.ld, + (bias + conditioning vector) from
        //     the shared-memory table, four 16-byte stores into the group's staging buffer [128 rows][20];
        //  2. QUAD view - thread (row te / 4 + 32 j, quad te % 4), j = 0..3: reads its 4 channels back, adds the fp32 residual and
        //     stores fp32 / bf16 NHWC (or fp32 NCHW): four neighbouring lanes cover the 64 contiguous bytes of one pixel, so
        //     one instruction touches 8 lines instead of 32.
        // The kernel is instruction-issue bound (first pipelined version: 17.5 k warp instructions per tile, 8.2 k cycles), so:
        // the residual of the NEXT tile is loaded into the registers the current chunk has just consumed (pure loads, one tile
        // of latency hidden, no rotation); pixel offsets are computed once per tile; the GroupNorm statistics of the output
        // stay in registers (4 channels x 4 rows per thread and chunk) and are folded - shuffles + fp64 atomics - only when
        // the CTA moves on to another sample.
        const int gi = (warp - 4) >> 2;
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int te = tid - 128 - gi * 128;
        float* stg = reinterpret_cast<float*>(red + (size_t)gi * 2 * SK_RED_BYTES);      // two [128][20] staging buffers
        const int nchunks = p.BN >> 4;
        // BN >= 32: group gi takes chunks gi, gi + 2 of EVERY tile;  BN = 16 (one chunk): the groups take alternate TILES (group gi =
        // accumulator buffer gi), so both halves of the epilogue warps work and a tile may take two tile periods to drain
        const bool alt = nchunks == 1;
        const int ncg = alt ? 1 : (nchunks - gi + SK_EPI_GROUPS - 1) / SK_EPI_GROUPS;       // this group's chunks per tile: 1 or 2
        const int i_first = alt ? gi : 0, i_step = alt ? 2 : 1;
        const int copy = (int)(blockIdx.x % TC_SUM_COPIES);
        const int quad = te & 3, r0 = te >> 2;
        const int Cout = p.epi.Cout;
        const float* bt = reinterpret_cast<const float*>(gbase + bt_off);
        const bool vec_ok = (Cout & 3) == 0 && !p.epi.out_nchw;
        const int nq = nt * p.BN + (alt ? 0 : gi * 16) + quad * 4;     // first channel of this thread's quad in chunk kk = 0 (+ 32 kk)
        float4 res[2][4];
        float acc[2][8];                                           // (sum x4, sum of squares x4) per chunk of the current sample
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
        // element offsets (pixel * Cout + nq, or -1) of rows r0 + 32 j of a tile
        auto offsets = [&](int i, int& b, int (&off)[4]) {
            int y0, x0;
            tile_coords(i, b, y0, x0);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int r = r0 + 32 * jj;
                const int pr = r / SK_PW, pc = r - pr * SK_PW;
                const int y = y0 + pr, x = x0 + pc;
                off[jj] = (pr < SK_TH && pc < SK_TW && y < p.H && x < p.W) ? ((b * p.H + y) * p.W + x) * Cout + nq : -1;
            }
        };
        auto load_res = [&](const int (&off)[4], int kk, float4 (&rr)[4]) {
            if (!p.epi.residual) return;
            const int n0 = nq + 32 * kk;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                if (off[jj] < 0 || n0 >= Cout) continue;
                const float* src = p.epi.residual + off[jj] + 32 * kk;
                if (vec_ok) rr[jj] = __ldg(reinterpret_cast<const float4*>(src));
                else rr[jj] = stream_load_tail(src, Cout - n0);
            }
        };
        auto flush = [&](int b) {                                  // fold the register statistics of sample b
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k < ncg) {
#pragma unroll
                    for (int o = 4; o <= 16; o <<= 1)
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[k][e] += __shfl_xor_sync(0xffffffffu, acc[k][e], o);
                    if (lane < 4) {
                        const int n0 = nq + 32 * k;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (n0 + e < Cout) {
                                double* dst = p.epi.sums_out + (((size_t)copy * p.epi.sums_B + b) * Cout + n0 + e) * 2;
                                atomicAdd(dst, (double)acc[k][e]);
                                atomicAdd(dst + 1, (double)acc[k][4 + e]);
                            }
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
                }
            }
        };
        // The fp32 residual: read with ordinary loads from HBM it is bound by the misses an SM's L1 can track (~6 B/clk per SM: the
        // epilogue took 5 k cycles per chunk); instead thread 0 of the group has TMA bring the chunk (box {20 ch, 18, 7} = the
        // staging layout, rows m = 18 pr + pc) into the staging buffer two chunks ahead, and the ROW view adds the accumulator to it.
        const bool rtma = p.res_prefetch != 0;
        const bool rldg = p.epi.residual != nullptr && !rtma;
        int qi = i_first, qk = 0;
        uint32_t qn = 0, cn = 0;
        auto issue_next = [&]() {
            if (qi >= my_tiles) return;
            int qb, qy0, qx0;
            tile_coords(qi, qb, qy0, qx0);
            const int c0 = alt ? 0 : (gi + SK_EPI_GROUPS * qk) * 16;
            const uint32_t buf = qn & 1u;
            mbar_expect_tx(rfull(gi, buf), SK_RES_BOX_BYTES);
            tma_load_4d(smem_u32(stg) + buf * SK_RED_BYTES, &p.resmap, rfull(gi, buf), nt * p.BN + c0, qx0, qy0, qb);
            ++qn;
            if (++qk == ncg) { qk = 0; qi += i_step; }
        };
        if (rtma && te == 0 && ncg > 0) { issue_next(); issue_next(); }
        if (i_first < my_tiles) {
            int b = 0, off[4], b_acc = -1;
            offsets(i_first, b, off);
            b_acc = b;
            if (rldg) {
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (k < ncg) load_res(off, k, res[k]);
            }
#pragma unroll 1
            for (int i = i_first;; i += i_step) {
                if (p.epi.sums_out && (i >= my_tiles || b != b_acc)) { flush(b_acc); b_acc = b; }      // the one call site of flush
                if (i >= my_tiles) break;
                const int s = i & 1;
                int b_next = b, off_next[4] = {-1, -1, -1, -1};
                if (i + i_step < my_tiles) offsets(i + i_step, b_next, off_next);
                {
                    SK_T0();
                    mbar_wait_relaxed(full_acc(s), (uint32_t)(i >> 1) & 1u);
                    SK_ACC(0);
                }
                tc_fence_after();
                SK_T0();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    if (kk < ncg) {
                        const int c0 = alt ? 0 : (gi + SK_EPI_GROUPS * kk) * 16;
                        const uint32_t buf = cn & 1u;
                        float* const sb = stg + buf * (128 * SK_STG_LD);
                        // ---- phase 1 (ROW view)
                        uint32_t v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * p.BN + c0), v);
                        if (kk == ncg - 1) {          // this warp's last read of the accumulator buffer: hand it back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_acc(s)) : "memory");
                        }
                        {
                            const float4* btv = reinterpret_cast<const float4*>(bt + b * p.BN + c0);
                            float4* row = reinterpret_cast<float4*>(sb + m * SK_STG_LD);
                            if (rtma) {
                                mbar_wait_relaxed(rfull(gi, buf), (cn >> 1) & 1u);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float4 a = btv[e], rs = row[e];
                                    row[e] = make_float4(__uint_as_float(v[4 * e]) + a.x + rs.x, __uint_as_float(v[4 * e + 1]) + a.y + rs.y,
                                                         __uint_as_float(v[4 * e + 2]) + a.z + rs.z, __uint_as_float(v[4 * e + 3]) + a.w + rs.w);
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float4 a = btv[e];
                                    row[e] = make_float4(__uint_as_float(v[4 * e]) + a.x, __uint_as_float(v[4 * e + 1]) + a.y,
                                                         __uint_as_float(v[4 * e + 2]) + a.z, __uint_as_float(v[4 * e + 3]) + a.w);
                                }
                            }
                        }
                        if (gi == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
                        // ---- phase 2 (QUAD view)
                        const int n0 = nq + 32 * kk;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            float4 x = *reinterpret_cast<const float4*>(sb + (r0 + 32 * jj) * SK_STG_LD + quad * 4);
                            if (off[jj] < 0 || n0 >= Cout) continue;
                            if (rldg) { x.x += res[kk][jj].x; x.y += res[kk][jj].y; x.z += res[kk][jj].z; x.w += res[kk][jj].w; }
                            if (vec_ok) {
                                const int o = off[jj] + 32 * kk;
                                if (p.epi.out_f32) *reinterpret_cast<float4*>(p.epi.out_f32 + o) = x;
                                if (p.epi.out_b16) {
                                    const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
                                    uint2 pk;
                                    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                                    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                                    *reinterpret_cast<uint2*>(p.epi.out_b16 + o) = pk;
                                }
                            } else {
                                stream_store_tail(p.epi, x, (off[jj] - nq) / Cout, n0, p.H * p.W);
                                if (n0 + 1 >= Cout) x.y = 0.f;
                                if (n0 + 2 >= Cout) x.z = 0.f;
                                if (n0 + 3 >= Cout) x.w = 0.f;
                            }
                            acc[kk][0] += x.x; acc[kk][1] += x.y; acc[kk][2] += x.z; acc[kk][3] += x.w;
                            acc[kk][4] = fmaf(x.x, x.x, acc[kk][4]); acc[kk][5] = fmaf(x.y, x.y, acc[kk][5]);
                            acc[kk][6] = fmaf(x.z, x.z, acc[kk][6]); acc[kk][7] = fmaf(x.w, x.w, acc[kk][7]);
                        }
                        if (rldg) load_res(off_next, kk, res[kk]);      // next tile's residual into the registers just consumed
                        // two staging buffers: the next chunk writes the other one and its mid-chunk barrier orders this chunk's reads
                        // before the chunk after next; only the TMA refill of THIS buffer needs everyone to be done with it now
                        if (rtma) {
                            if (gi == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
                            if (te == 0) { fence_proxy_async(); issue_next(); }
                        }
                        ++cn;
                    }
                }
                b = b_next;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) off[jj] = off_next[jj];
                SK_ACC(1);
            }
        }
        if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && q == 0) { p.dbg[4 + 2 * gi] = acc_dbg[0]; p.dbg[5 + 2 * gi] = acc_dbg[1]; }
        tc_fence_before();
    } else {
        // ===== transform (warps 2, 3, 12 .. 15): raw fp32 patch -> normalise -> Swish -> operand image.  Thread xt owns channel plane
        // kp = xt % P (its scale / shift pairs stay in registers for the tile) and the patch pixels xt / P, + 192 / P, ...
        const int xt = warp < 4 ? tid - 64 : tid - 384 + 64;
        const int kp = xt % P, px_first = xt / P, px_step = SK_XF_THREADS / P;
        const int c0 = kp * CPP;
        const bool from_a = c0 < p.ca;
        const uint32_t src_pitch = (uint32_t)(from_a ? p.ca : p.cb) * 4u;
        const uint32_t src_first = (from_a ? (uint32_t)c0 * 4u : p.raw_a_bytes + (uint32_t)(c0 - p.ca) * 4u) + (uint32_t)px_first * src_pitch;
        const uint32_t dst_first = (uint32_t)kp * plane_bytes + (uint32_t)px_first * 16u;
        for (int i = 0; i < my_tiles; ++i) {
            const int r = i % SK_NRAW, s = i & 1;
            int b, y0, x0;
            tile_coords(i, b, y0, x0);
            float4 sc[CPP / 2];
#pragma unroll
            for (int j = 0; j < CPP / 2; ++j) sc[j] = reinterpret_cast<const float4*>(tab + b * p.C + c0)[j];
            { SK_T0(); mbar_wait_relaxed(full_raw(r), (uint32_t)((i / SK_NRAW) & 1)); SK_ACC(0); }
            { SK_T0(); mbar_wait_relaxed(empty_op(s), ((uint32_t)(i >> 1) & 1u) ^ 1u); SK_ACC(1); }
            SK_T0();
            uint32_t src = base + raw_off + (uint32_t)r * p.raw_bytes + src_first;
            uint32_t dst = base + op_off + (uint32_t)s * p.op_bytes + dst_first;
            // two patch pixels per iteration: the two load -> fma -> ex2 -> rcp -> cvt -> store chains are independent, so their
            // latencies overlap (one pixel per iteration: ~400 cycles each, the transform bound the 16-channel layers)
            auto xform = [&](float (&x)[CPP]) -> uint4 {
#pragma unroll
                for (int j = 0; j < CPP / 2; ++j) {
                    x[2 * j] = fmaf(x[2 * j], sc[j].x, sc[j].y);
                    x[2 * j + 1] = fmaf(x[2 * j + 1], sc[j].z, sc[j].w);
                }
                if (p.swish) {
#pragma unroll
                    for (int j = 0; j < CPP; ++j) {
                        if (TF32) {
                            x[j] = __fdividef(x[j], 1.0f + __expf(-x[j]));
                        } else {                    // y * sigmoid(y) = h * tanh(h) + h, h = y / 2: one MUFU op
                            const float h = 0.5f * x[j];
                            float th;
                            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
                            x[j] = fmaf(h, th, h);
                        }
                    }
                }
                if (TF32) {
                    return make_uint4(__float_as_uint(to_tf32(x[0])), __float_as_uint(to_tf32(x[1])), __float_as_uint(to_tf32(x[2])),
                                      __float_as_uint(to_tf32(x[3])));
                } else {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(x[(2 * j) % CPP], x[(2 * j + 1) % CPP]);
                        w[j] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    return make_uint4(w[0], w[1], w[2], w[3]);
                }
            };
            auto load_px = [&](uint32_t a, float (&x)[CPP]) {
                float4 v0;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(a));
                x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
                if (!TF32) {
                    float4 v1;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(a + 16u));
                    x[CPP - 4] = v1.x; x[CPP - 3] = v1.y; x[CPP - 2] = v1.z; x[CPP - 1] = v1.w;
                }
            };
            auto inside = [&](int px) {
                const int pr = px / SK_PW, pc = px - pr * SK_PW;
                return (unsigned)(y0 + pr - 1) < (unsigned)p.H && (unsigned)(x0 + pc - 1) < (unsigned)p.W;
            };
            const uint32_t sstep = (uint32_t)px_step * src_pitch, dstep = (uint32_t)px_step * 16u;
#pragma unroll 1
            for (int px = px_first; px < SK_PATCH_PX; px += 2 * px_step, src += 2u * sstep, dst += 2u * dstep) {
                const int px2 = px + px_step;
                const bool have2 = px2 < SK_PATCH_PX;
                float xa[CPP], xb[CPP];
                load_px(src, xa);
                load_px(have2 ? src + sstep : src, xb);            // (a repeated in-range address when there is no second pixel)
                const bool in1 = inside(px), in2 = have2 && inside(px2);
                uint4 va = xform(xa), vb = xform(xb);
                if (!in1) va = make_uint4(0u, 0u, 0u, 0u);         // the conv's zero padding is applied AFTER the normalisation
                if (!in2) vb = make_uint4(0u, 0u, 0u, 0u);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(va.x), "r"(va.y), "r"(va.z), "r"(va.w) : "memory");
                if (have2)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + dstep), "r"(vb.x), "r"(vb.y), "r"(vb.z), "r"(vb.w) : "memory");
            }
            fence_proxy_async();                  // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_op(s)) : "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_raw(r)) : "memory");
            }
            SK_ACC(2);
        }
        if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && xt == 0) { p.dbg[8] = acc_dbg[0]; p.dbg[9] = acc_dbg[1]; p.dbg[10] = acc_dbg[2]; }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
    trace_end(p.trace);
}

// ------------------------------------------------------------------------------------------ host side
static size_t stream_smem_bytes(int C, int BN, int B, int es) {
    const size_t b_bytes = (size_t)9 * C * BN * es;
    const size_t op_bytes = (size_t)(C * es / 16) * SK_PLANE_PX * 16;
    const size_t raw_bytes = (size_t)SK_PATCH_PX * C * 4;
    return 1024 + align_up(b_bytes, 128) + align_up(2 * op_bytes, 128) + SK_NRAW * raw_bytes + 128 + (size_t)B * C * 8 + 16 +
           (size_t)B * BN * 4 + 16 + 8 * (2 * SK_NRAW + 14) + 128 + SK_EPI_GROUPS * 2 * SK_RED_BYTES + 256;
}

static int stream_weight_loads(int C, int es) {
    const int per_tap = C * es / 16;
    for (int l = 1; l <= 9; ++l)
        if (9 % l == 0 && (9 / l) * per_tap <= 256) return l;
    return 0;
}

static int stream_pick_bn(int cout, int C, int B, int es) {
    const int npad = (cout + 15) / 16 * 16;
    for (int bn = npad; bn >= 16; bn -= 16) {
        if (npad % bn || bn > 64) continue;          // two epilogue groups x at most two 16-channel chunks per tile
        if (stream_smem_bytes(C, bn, B, es) <= SK_SMEM_LIMIT) return bn;
    }
    return 0;
}

// The pipelined kernel pays off when every CTA walks several tiles (weights / table / TMEM once per CTA, stages overlapped)
// and the input need not be staged more than twice (N split only when the weights force it).
bool stream_conv_preferred(int ca, int cb, int cout, int ks, int B, int H, int W, int tf32) {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("DIFFSPLIT_B200_STREAM"); mode = e ? atoi(e) : 1; }
    if (mode == 0) return false;
    const int C = ca + cb, es = tf32 ? 4 : 2;
    if (ks != 3 || ca <= 0 || ca % 8 || cb % 8 || C % 16 || C > 256 || W < 8) return false;
    if (ca * 4 > 1024 || cb * 4 > 1024) return false;                 // TMA box: <= 256 elements per dimension
    if (stream_weight_loads(C, es) == 0) return false;
    if (SK_XF_THREADS % (C * es / 16)) return false;                  // a transform thread owns one channel plane
    if ((int64_t)B * H * W * ((cout + 3) / 4 * 4) >= (1ll << 31)) return false;     // 32-bit element offsets in the epilogue
    const int bn = stream_pick_bn(cout, C, B, es);
    if (bn == 0) return false;
    const int n_tiles = (cout + 15) / 16 * 16 / bn;
    if (n_tiles > 2) return false;
    if ((size_t)B * C * 16 > (size_t)(C * es / 16) * SK_PLANE_PX * 16) return false;     // statistics fold scratch = OP[1]
    const int64_t m_tiles = (int64_t)B * ((H + SK_TH - 1) / SK_TH) * ((W + SK_TW - 1) / SK_TW);
    return mode == 2 || m_tiles * n_tiles >= 4 * 148;
}

static PFN_cuTensorMapEncodeTiled_v12000 g_enc = nullptr;
static long long* g_stream_dbg = nullptr;
static int stream_encoder() {
    if (g_enc) return DS_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DS_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    DS_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    g_enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return DS_OK;
}

int stream_launch_conv(const float* src_a, int ca, const float* src_b, int cb, const HaloNorm& norm, const uint8_t* w_packed,
                       int cout, int ks, int B, int H, int W, const ConvEpi& epi, float* out_f32, void* out_b16, float* out_nchw,
                       double* sums_out, int tf32, cudaStream_t st) {
    DS_REQUIRE(ks == 3 && stream_conv_preferred(ca, cb, cout, ks, B, H, W, tf32) , "stream conv: unsupported shape");
    int rc = stream_encoder();
    if (rc != DS_OK) return rc;
    StreamParams p;
    memset(&p, 0, sizeof(p));
    const int es = tf32 ? 4 : 2;
    p.ca = ca; p.cb = cb; p.C = ca + cb;
    p.stats = norm.stats; p.sums_a = norm.sums_a; p.sums_b = norm.sums_b;
    p.gamma = norm.gamma; p.beta = norm.beta; p.G = norm.G > 0 ? norm.G : 1; p.swish = norm.swish;
    p.epi.sums_out = sums_out; p.epi.sums_B = B;
    p.epi.bias = epi.bias; p.epi.temb = epi.temb; p.epi.temb_off = epi.temb_off; p.epi.temb_stride = epi.temb_stride;
    p.epi.temb_bcast = epi.temb_bcast; p.epi.residual = epi.residual;
    p.epi.out_f32 = out_f32; p.epi.out_b16 = reinterpret_cast<__nv_bfloat16*>(out_b16); p.epi.out_nchw = out_nchw;
    p.epi.Cout = cout; p.epi.Ho = H; p.epi.Wo = W;
    p.B = B; p.H = H; p.W = W;
    p.ksteps = p.C * es / 32;
    p.wloads = stream_weight_loads(p.C, es);
    p.BN = stream_pick_bn(cout, p.C, B, es);
    const int npad = (cout + 15) / 16 * 16;
    const int n_tiles = npad / p.BN;
    p.tiles_x = (W + SK_TW - 1) / SK_TW;
    p.tiles_y = (H + SK_TH - 1) / SK_TH;
    p.n_tiles_m = B * p.tiles_x * p.tiles_y;
    p.div_tiles_x = make_fastdiv((uint32_t)p.tiles_x);
    p.div_tiles_xy = make_fastdiv((uint32_t)(p.tiles_x * p.tiles_y));
    p.b_bytes = (uint32_t)(9 * p.C * p.BN * es);
    p.op_bytes = (uint32_t)((p.C * es / 16) * SK_PLANE_PX * 16);
    p.raw_a_bytes = (uint32_t)(SK_PATCH_PX * ca * 4);
    p.raw_bytes = (uint32_t)(SK_PATCH_PX * p.C * 4);
    {
        const cuuint64_t tkp = (cuuint64_t)9 * p.ksteps * 2;
        cuuint64_t dims[3] = {128, (cuuint64_t)(npad / 16), tkp};
        cuuint64_t strides[2] = {tkp * 256, 256};
        cuuint32_t box[3] = {128, (cuuint32_t)(p.BN / 16), (cuuint32_t)(tkp / p.wloads)};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = g_enc(&p.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<uint8_t*>(w_packed), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("stream conv: cuTensorMapEncodeTiled for the weights failed (%d)", (int)r); return DS_ERR_CUDA; }
    }
    for (int s = 0; s < 2; ++s) {
        const float* ptr = s == 0 ? src_a : src_b;
        const int c = s == 0 ? ca : cb;
        if (!ptr || c == 0) continue;
        cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)c * 4, (cuuint64_t)W * c * 4, (cuuint64_t)H * W * c * 4};
        cuuint32_t box[4] = {(cuuint32_t)c, SK_PW, SK_PH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = g_enc(&p.rmap[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("stream conv: cuTensorMapEncodeTiled for source %d failed (%d)", s, (int)r); return DS_ERR_CUDA; }
    }
    if (epi.residual) {
        cuuint64_t dims[4] = {(cuuint64_t)cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)cout * 4, (cuuint64_t)W * cout * 4, (cuuint64_t)H * W * cout * 4};
        cuuint32_t box[4] = {SK_STG_LD, SK_PW, SK_TH, 1};         // one 16-channel chunk (+ 4) of a tile in the staging layout
        cuuint32_t estr[4] = {1, 1, 1, 1};
        static int respf_env = -1;
        if (respf_env < 0) { const char* e5 = getenv("DIFFSPLIT_B200_STREAM_RESPF"); respf_env = e5 ? atoi(e5) : 1; }
        if (respf_env && (cout * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(epi.residual) & 15) == 0) {
            CUresult r = g_enc(&p.resmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(epi.residual), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("stream conv: cuTensorMapEncodeTiled for the residual failed (%d)", (int)r); return DS_ERR_CUDA; }
            p.res_prefetch = 1;
        }
    }
    const size_t smem = stream_smem_bytes(p.C, p.BN, B, es);
    p.trace = trace_next(6);
    static const bool dbg_on = getenv("DIFFSPLIT_B200_STREAM_DBG") != nullptr;
    if (dbg_on) {
        if (!g_stream_dbg) DS_CHECK_CUDA(cudaMalloc(&g_stream_dbg, 16 * sizeof(long long)));
        DS_CHECK_CUDA(cudaMemsetAsync(g_stream_dbg, 0, 16 * sizeof(long long), st));
        p.dbg = g_stream_dbg;
    }
    static bool attr_set = false;
    if (!attr_set) {
        DS_CHECK_CUDA(cudaFuncSetAttribute(conv_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM_LIMIT + 2048));
        DS_CHECK_CUDA(cudaFuncSetAttribute(conv_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM_LIMIT + 2048));
        attr_set = true;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        DS_CHECK_CUDA(cudaGetDevice(&dev));
        DS_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int gx = sms / n_tiles;
    if (gx > p.n_tiles_m) gx = p.n_tiles_m;
    if (gx < 1) gx = 1;
    const dim3 grid((unsigned)gx, (unsigned)n_tiles, 1);
    const cudaError_t err = tf32 ? launch_pdl(conv_stream_kernel<true>, grid, dim3(SK_THREADS), smem, st, p)
                                 : launch_pdl(conv_stream_kernel<false>, grid, dim3(SK_THREADS), smem, st, p);
    DS_CHECK_CUDA(err);
    return DS_OK;
}

}  // namespace ds

// debugging aid (DIFFSPLIT_B200_STREAM_DBG=1): cycle sums of CTA 0 of the last conv_stream launch:
// [0] producer wait empty_raw | [1..3] MMA wait full_op, wait empty_acc, issue+commit | [4,5] / [6,7] epilogue group 0 / 1 wait
// full_acc, work | [8..10] transform wait full_raw, wait empty_op, work | [14] tiles of the CTA | [15] producer lifetime
extern "C" int ds_debug_stream_phases(long long* out16) {
    using namespace ds;
    DS_REQUIRE(g_stream_dbg && out16, "stream debug buffer not active (DIFFSPLIT_B200_STREAM_DBG=1)");
    DS_CHECK_CUDA(cudaDeviceSynchronize());
    DS_CHECK_CUDA(cudaMemcpy(out16, g_stream_dbg, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    return DS_OK;
}
