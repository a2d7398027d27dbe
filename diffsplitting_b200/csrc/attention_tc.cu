// bf16 tensor-core single-head self-attention (precision mode DS_PREC_BF16): tcgen05.mma + TMEM + TMA, flash-style.
// Replaces the reference's  einsum("bnchw,bncyx->bnhwyx") / sqrt(C) -> softmax -> einsum("bnhwyx,bncyx->bnchw")
// (model/sr3_modules/unet.py:126-140, model/ddpm_modules/unet.py:113-127) without the N x N score matrix in HBM.
//
//   qkv : bf16 [B, N, 3C]  (q | k | v along the channel axis = the NHWC output of the 1x1 qkv conv)
//   out : bf16 [B, N, C]
//
// One CTA = 128 queries (TMEM lanes) x one CV-wide slice of the value/output channels (CTAs of different slices recompute
// the scores).  Keys are walked in blocks of 128, TWICE: pass A computes the exact row maximum and the softmax
// denominator (scores only), pass B recomputes the scores, writes P = exp(s - max) as a bf16 A operand into shared memory
// and accumulates O += P V in TMEM - so the accumulator is never rescaled.  With N <= 128 (the 64 x 64 splitting networks:
// N = 64) the scores of pass A are still in TMEM and pass B does not recompute them.
//
//   warp 0      : TMA producer  (Q and K 64-channel panels through a ring, V block)
//   warp 1      : MMA issuer    (S = Q K^T: both operands K-major SW128;  O += P V: V is consumed MN-major - it arrives
//                                as [key][channel] rows, which IS the canonical MN-major SW128 layout)
//   warps 2..5  : softmax / epilogue, one thread per query row (tcgen05.ld 32x32b)
//
// For C <= 256 the query panels are loaded ONCE and stay resident (the ring then carries keys only); with more than one key
// block the scores are double-buffered in TMEM, so the tensor core computes block j+1 while the softmax warps work on
// block j (an SM ingests only ~25-30 B/clk from L2: reloading Q per key block and serialising MMA and softmax made the
// N = 4096 case 2x slower).
// TMEM: one key block: columns [0,128) scores, [128,128+CV) output; otherwise [0,128) and [128,256) scores, [256,256+CV) output.  Everything accumulates in fp32; P is rounded to bf16
// (values in [0,1]); the normalisation 1/l is applied to the fp32 accumulator in the epilogue.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "tc.cuh"
#include "tc_ptx.cuh"

namespace ds {

constexpr int AT_THREADS = 192;
constexpr int AT_BK = 128;                       // keys per block (= N of the score MMA)
constexpr uint32_t AT_PANEL = 128 * 128;         // one 64-channel (or 64-key) panel of 128 rows x 128 B

struct AttnTcParams {
    CUtensorMap map;                             // qkv as [3C, N, B] bf16, box {64, 128, 1}, SWIZZLE_128B
    __nv_bfloat16* out;
    int B, N, C, CV, stages, q_res;
    float scale_log2;                            // log2(e) / sqrt(C)
    TraceSlot trace;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// 2^x for x <= 0 (scores minus their row maximum): the bare SFU instruction; exp2f() wraps it in range scaling that the
// softmax does not need (a flushed denormal is 0 either way after the bf16 rounding / in a sum of values <= 1)
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// SW128 descriptor, K-major (rows of 128 B, 8-row groups 1024 B apart) or MN-major (lbo = stride between 64-element
// groups along M/N, 8-row K groups 1024 B apart)
__device__ __forceinline__ uint64_t at_desc(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

__global__ void __launch_bounds__(AT_THREADS) attention_tc_kernel(const __grid_constant__ AttnTcParams p) {
    extern __shared__ __align__(1024) uint8_t at_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * 128, cs = blockIdx.y * p.CV, b = blockIdx.z;
    const int nkb = (p.N + AT_BK - 1) / AT_BK;
    const bool single = nkb == 1;
    const int nchunk = p.C >> 6;                                  // 64-channel panels of the score contraction
    const int vpanels = p.CV >> 6;

    const uint32_t base = (smem_u32(at_smem) + 1023u) & ~1023u;
    const uint32_t qbuf = base;                                   // resident Q panels (q_res)
    const uint32_t ring = base + (p.q_res ? (uint32_t)nchunk * AT_PANEL : 0u);   // stages x (K panel) or (Q panel | K panel)
    const uint32_t slot = p.q_res ? AT_PANEL : 2u * AT_PANEL;
    const uint32_t vbuf = ring + (uint32_t)p.stages * slot;
    const uint32_t pbuf = vbuf + (uint32_t)vpanels * AT_PANEL;    // P: 2 key panels
    const uint32_t bars = pbuf + 2u * AT_PANEL;
    auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(4 + s); };
    auto s_full = [&](int i) { return bars + 64u + 8u * (uint32_t)i; };        // score buffer i written
    auto s_free = [&](int i) { return bars + 80u + 8u * (uint32_t)i; };        // score buffer i consumed
    const uint32_t p_full = bars + 96, pv_done = bars + 104, v_full = bars + 112, q_full = bars + 120;
    const uint32_t tmem_slot = bars + 128;
    const uint32_t tmem_cols = single && p.CV <= 128 ? 256u : 512u;

    trace_begin(p.trace);
    if (warp == 0 && elect_one()) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(s_full(i), 1); mbar_init(s_free(i), 128); }
        mbar_init(q_full, 1);
        mbar_init(p_full, 128);
        mbar_init(pv_done, 1);
        mbar_init(v_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + (single ? 128u : 256u);
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (elect_one()) {
            if (p.q_res) {
                mbar_expect_tx(q_full, (uint32_t)nchunk * AT_PANEL);
                for (int cc = 0; cc < nchunk; ++cc) tma_load_3d(qbuf + (uint32_t)cc * AT_PANEL, &p.map, q_full, cc * 64, q0, b);
            }
            int u = 0;
            auto load_k = [&](int j) {
                for (int cc = 0; cc < nchunk; ++cc, ++u) {
                    const int s = u % p.stages;
                    mbar_wait(empty_bar(s), (((uint32_t)(u / p.stages)) & 1u) ^ 1u);
                    const uint32_t dst = ring + (uint32_t)s * slot;
                    mbar_expect_tx(full_bar(s), slot);
                    if (!p.q_res) tma_load_3d(dst, &p.map, full_bar(s), cc * 64, q0, b);
                    tma_load_3d(dst + (p.q_res ? 0u : AT_PANEL), &p.map, full_bar(s), p.C + cc * 64, j * AT_BK, b);
                }
            };
            for (int j = 0; j < nkb; ++j) load_k(j);                       // pass A
            if (!single) load_k(0);                                        // pass B: keys run one block ahead of V
            for (int j = 0; j < nkb; ++j) {
                if (!single && j + 1 < nkb) load_k(j + 1);
                if (j > 0) mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);     // V buffer free
                mbar_expect_tx(v_full, (uint32_t)vpanels * AT_PANEL);
                for (int pc = 0; pc < vpanels; ++pc)
                    tma_load_3d(vbuf + (uint32_t)pc * AT_PANEL, &p.map, v_full, 2 * p.C + cs + pc * 64, j * AT_BK, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            // D = f32, A = B = bf16, M = 128; scores: N = 128, both K-major; output: N = CV, B MN-major (bit 16)
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BK >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(p.CV >> 3) << 17) | ((128u >> 4) << 24);
            int u = 0, it = 0;
            if (p.q_res) { mbar_wait(q_full, 0u); tc_fence_after(); }
            auto scores = [&]() {
                const int sb = single ? 0 : (it & 1);                       // score buffer of this block
                if (it >= 2) { mbar_wait(s_free(sb), (uint32_t)((it >> 1) - 1) & 1u); tc_fence_after(); }
                for (int cc = 0; cc < nchunk; ++cc, ++u) {
                    const int s = u % p.stages;
                    mbar_wait(full_bar(s), ((uint32_t)(u / p.stages)) & 1u);
                    tc_fence_after();
                    const uint32_t sa = ring + (uint32_t)s * slot;
                    const uint64_t qd = at_desc(p.q_res ? qbuf + (uint32_t)cc * AT_PANEL : sa, 16);
                    const uint64_t kd = at_desc(p.q_res ? sa : sa + AT_PANEL, 16);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_s + (uint32_t)sb * 128u, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, (cc | k) ? 1u : 0u);
                    umma_commit(empty_bar(s));
                }
                umma_commit(s_full(sb));
                ++it;
            };
            for (int j = 0; j < nkb; ++j) scores();                        // pass A
            if (!single) scores();                                         // pass B: scores run one block ahead of P V
            for (int j = 0; j < nkb; ++j) {
                if (!single && j + 1 < nkb) scores();
                mbar_wait(p_full, (uint32_t)j & 1u);
                mbar_wait(v_full, (uint32_t)j & 1u);
                tc_fence_after();
                const int ksteps = (min(AT_BK, p.N - j * AT_BK) + 15) >> 4;      // P is zero beyond the last valid key
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t pd = at_desc(pbuf + (uint32_t)(k >> 2) * AT_PANEL + (uint32_t)(k & 3) * 32u, 16);
                    const uint64_t vd = at_desc(vbuf + (uint32_t)k * 2048u, AT_PANEL);
                    umma_bf16(tmem_o, pd, vd, idesc_o, (j | k) ? 1u : 0u);
                }
                umma_commit(pv_done);
            }
        }
        __syncwarp();
    } else {
        const int qd = warp & 3;
        const int m = qd * 32 + lane;                                       // query row == TMEM lane
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        int it = 0;
        uint32_t v[16];
        uint32_t sv[AT_BK];
        for (int j = 0; j < nkb; ++j, ++it) {                               // pass A: exact max and denominator
            const int sb = single ? 0 : (it & 1);
            const uint32_t tmem_sj = tmem_s + lane_addr + (uint32_t)sb * 128u;
            mbar_wait(s_full(sb), (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const int kvalid = min(AT_BK, p.N - j * AT_BK);
            // the whole 128-column row in registers: 8 pipelined TMEM loads, one wait, one read of the scores
#pragma unroll
            for (int c = 0; c < AT_BK / 16; ++c) tmem_ld16_nowait(tmem_sj + (uint32_t)(c * 16), reinterpret_cast<uint32_t(&)[16]>(sv[c * 16]));
            tmem_ld_wait();
            float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
            if (kvalid == AT_BK) {
#pragma unroll
                for (int i = 0; i < AT_BK; i += 4) {
                    b0 = fmaxf(b0, __uint_as_float(sv[i]));
                    b1 = fmaxf(b1, __uint_as_float(sv[i + 1]));
                    b2 = fmaxf(b2, __uint_as_float(sv[i + 2]));
                    b3 = fmaxf(b3, __uint_as_float(sv[i + 3]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < AT_BK; ++i)
                    if (i < kvalid) b0 = fmaxf(b0, __uint_as_float(sv[i]));
            }
            const float m_new = fmaxf(m_run, fmaxf(fmaxf(b0, b1), fmaxf(b2, b3)));
            const float mneg = -m_new * p.scale_log2;
            // four independent partial sums: one dependent FADD chain over 128 SFU results made this loop latency bound
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            if (kvalid == AT_BK) {
#pragma unroll
                for (int i = 0; i < AT_BK; i += 4) {
                    s0 += ex2_fast(fmaf(__uint_as_float(sv[i]), p.scale_log2, mneg));
                    s1 += ex2_fast(fmaf(__uint_as_float(sv[i + 1]), p.scale_log2, mneg));
                    s2 += ex2_fast(fmaf(__uint_as_float(sv[i + 2]), p.scale_log2, mneg));
                    s3 += ex2_fast(fmaf(__uint_as_float(sv[i + 3]), p.scale_log2, mneg));
                }
            } else {
#pragma unroll
                for (int i = 0; i < AT_BK; ++i)
                    if (i < kvalid) s0 += ex2_fast(fmaf(__uint_as_float(sv[i]), p.scale_log2, mneg));
            }
            const float sum = (s0 + s1) + (s2 + s3);
            l_run = l_run * ex2_fast((m_run - m_new) * p.scale_log2) + sum;    // first block: 2^(-inf) = 0
            m_run = m_new;
            if (!single) { tc_fence_before(); mbar_arrive(s_free(sb)); }
        }
        const uint32_t prow = (uint32_t)m * 128u;
        const float mneg_run = -m_run * p.scale_log2;
        for (int j = 0; j < nkb; ++j) {                                     // pass B: P = exp(s - max) -> smem
            const int sb = single ? 0 : (it & 1);
            const uint32_t tmem_sj = tmem_s + lane_addr + (uint32_t)sb * 128u;
            if (!single) {
                mbar_wait(s_full(sb), (uint32_t)(it >> 1) & 1u);
                tc_fence_after();
                ++it;
            }
            if (j > 0) mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);         // P buffer free
            const int kvalid = min(AT_BK, p.N - j * AT_BK);
#pragma unroll
            for (int c = 0; c < AT_BK / 16; ++c) tmem_ld16_nowait(tmem_sj + (uint32_t)(c * 16), reinterpret_cast<uint32_t(&)[16]>(sv[c * 16]));
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < AT_BK / 16; ++c) {
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float e0 = ex2_fast(fmaf(__uint_as_float(sv[c * 16 + 2 * i]), p.scale_log2, mneg_run));
                    float e1 = ex2_fast(fmaf(__uint_as_float(sv[c * 16 + 2 * i + 1]), p.scale_log2, mneg_run));
                    if (c * 16 + 2 * i >= kvalid) e0 = 0.f;
                    if (c * 16 + 2 * i + 1 >= kvalid) e1 = 0.f;
                    __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
                    pk[i] = *reinterpret_cast<uint32_t*>(&h);
                }
                // keys [16c, 16c+16) of row m: panel c/4, 16-byte chunks 2(c%4), 2(c%4)+1, XOR-swizzled with the row
                const uint32_t pan = pbuf + (uint32_t)(c >> 2) * AT_PANEL + prow;
                const uint32_t ch0 = (uint32_t)(2 * (c & 3)), sw = (uint32_t)(m & 7);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pan + ((ch0 ^ sw) << 4)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pan + (((ch0 + 1u) ^ sw) << 4)), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(p_full);
            if (!single) mbar_arrive(s_free(sb));
        }
        mbar_wait(pv_done, (uint32_t)(nkb - 1) & 1u);
        tc_fence_after();
        const float inv = 1.f / l_run;
        const int q = q0 + m;
        __nv_bfloat16* orow = p.out + ((size_t)b * p.N + (size_t)min(q, p.N - 1)) * p.C + cs;
        for (int c = 0; c < p.CV / 16; ++c) {
            tmem_ld16(tmem_o + lane_addr + (uint32_t)(c * 16), v);         // warp-collective: rows >= N load too
            if (q < p.N) {
                uint32_t ob[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * i]) * inv, __uint_as_float(v[2 * i + 1]) * inv);
                    ob[i] = *reinterpret_cast<uint32_t*>(&h);
                }
                uint4* dst = reinterpret_cast<uint4*>(orow + c * 16);
                dst[0] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
                dst[1] = make_uint4(ob[4], ob[5], ob[6], ob[7]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    trace_end(p.trace);
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------ host
static_assert(sizeof(AttnTcParams) <= sizeof(AttnTcPlan::params), "AttnTcPlan::params too small");

bool attn_tc_supported(int N, int C) { return N > 0 && C >= 64 && C % 64 == 0; }

static int attn_pick_cv(int C) { return C % 256 == 0 ? 256 : (C % 128 == 0 ? 128 : 64); }

int attn_tc_build(AttnTcPlan* plan, const void* qkv, void* out, int B, int N, int C) {
    DS_REQUIRE(attn_tc_supported(N, C), "attention (bf16): unsupported shape N=%d C=%d (C %% 64)", N, C);
    DS_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "attention (bf16): buffers must be 16-byte aligned");
    static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        DS_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        DS_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
        enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    AttnTcParams* p = reinterpret_cast<AttnTcParams*>(plan->params);
    memset(p, 0, sizeof(AttnTcParams));
    cuuint64_t dims[3] = {(cuuint64_t)(3 * C), (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)3 * C * 2, (cuuint64_t)N * 3 * C * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&p->map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("attention (bf16): cuTensorMapEncodeTiled failed (%d) for B=%d N=%d C=%d", (int)r, B, N, C);
        return DS_ERR_CUDA;
    }
    p->out = reinterpret_cast<__nv_bfloat16*>(out);
    p->B = B; p->N = N; p->C = C;
    p->CV = attn_pick_cv(C);
    p->q_res = C <= 256;
    const size_t fixed = (size_t)(p->q_res ? (C / 64) * AT_PANEL : 0) + (size_t)(p->CV / 64) * AT_PANEL + 2 * AT_PANEL;
    const size_t slot_bytes = p->q_res ? AT_PANEL : 2 * AT_PANEL;
    int stages = (int)((200 * 1024 - fixed) / slot_bytes);
    const int want = p->q_res ? 4 : 3;
    if (stages > want) stages = want;
    if (N <= AT_BK && stages > C / 64) stages = C / 64;                          // one key block: nothing to prefetch
    DS_REQUIRE(stages >= 1, "attention (bf16): no room for the operand ring at C=%d", C);
    p->stages = stages;
    p->scale_log2 = (float)(1.4426950408889634 / sqrt((double)C));
    plan->smem_bytes = (int)(fixed + (size_t)p->stages * slot_bytes + 256 + 1024);
    plan->grid_x = cdiv(N, 128); plan->grid_y = C / p->CV; plan->grid_z = B;
    return DS_OK;
}

int attn_tc_launch(const AttnTcPlan* plan, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        DS_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        attr_set = true;
    }
    AttnTcParams p = *reinterpret_cast<const AttnTcParams*>(plan->params);
    p.trace = trace_next(3);
    DS_CHECK_CUDA(launch_pdl(attention_tc_kernel, dim3(plan->grid_x, plan->grid_y, plan->grid_z), dim3(AT_THREADS), (size_t)plan->smem_bytes, st, p));
    return DS_OK;
}

}  // namespace ds
