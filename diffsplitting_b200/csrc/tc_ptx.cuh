// PTX wrappers shared by the tensor-core kernels (mbarrier, TMA / bulk copies, TMEM, tcgen05.mma / ld / commit) and the
// common epilogue (bias + conditioning vector + fp32 residual -> fp32 NHWC / bf16 NHWC / fp32 NCHW stores).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ds {

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a lost TMA / MMA completion traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// hot-loop wait: one poll without the clock reads; the bounded loop only when the phase is not there yet
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    mbar_wait(bar, parity);
}
// the same for waits that are expected to be long (a whole pipeline stage): the hardware may park the thread for up to
// ~20 us per try (it still wakes when the phase completes), so a waiting warp does not spend issue slots on the poll loop
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// One lane of a fully converged warp.  Selecting the issuing thread with elect.sync (instead of `lane == 0`) lets ptxas prove
// that a single thread is active: without it every UTCHMMA / UTMALDG is wrapped in an ELECT / R2UR / BRA.U.ANY loop over
// the "possibly several" active lanes, which alone costs more than a small-N MMA (tools/mma_probe.cu: 52 vs 39 cycles).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with fp32 operands read as TF32 (10-bit mantissa; K = 8 per instruction = 32 bytes per row, as K = 16 bf16)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (1 << 4); A / B format at [7,10) / [10,13): 1 = BF16, 2 = TF32;
// both K-major; N >> 3 at [17,23); M >> 4 at [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int N, bool tf32) {
    const uint32_t fmt = tf32 ? 2u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
// round-to-nearest fp32 -> tf32 (the tensor core would otherwise truncate the low 13 mantissa bits)
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
template <bool TF32>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (TF32) umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
    else umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: issue several, then tmem_ld_wait() once (the loads pipeline)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version | [61,64) layout (2=SW128, 4=SW64, 6=SW32)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    const uint64_t sbo = (8u * row_bytes) >> 4;          // 8-row group stride
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}


// the same with an explicit stride between the 8-row groups (any multiple of the row size: the swizzle is a function of the
// shared-memory address, so the groups need not be contiguous)
__device__ __forceinline__ uint64_t make_smem_desc_sbo(uint32_t saddr, uint32_t row_bytes, uint32_t sbo_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (layout << 61);
}

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

// ------------------------------------------------------------------------------------------ shared epilogue
struct TcEpi {
    const float* bias;                // [Cout] or null
    const float* bias2;               // second bias (the folded 1x1 res_conv of tc_chain.cu) or null; chain kernel only
    const float* temb;                // conditioning vectors or null
    int temb_off, temb_stride, temb_bcast;
    const float* residual;            // fp32 NHWC [B,Ho,Wo,Cout] or null (the residual stream stays fp32)
    float* out_f32;                   // fp32 NHWC or null   (consumers: GroupNorm statistics, residual adds)
    __nv_bfloat16* out_b16;           // bf16 NHWC or null   (consumers: TMA-fed convolutions, attention)
    float* out_nchw;                  // fp32 NCHW or null   (the network output)
    double* sums_out;                 // [TC_SUM_COPIES][B][Cout][2] (sum, sum of squares) accumulators of the OUTPUT tensor
                                      // or null: the GroupNorm statistics of the consumer, produced here instead of by a
                                      // pass over HBM
    int sums_B;                       // batch size (stride of the replicated accumulators)
    int Cout, Ho, Wo;
};

// column sums of a 32-row x 16-column block held as v[16] per lane: recursive halving, 16 shuffles; afterwards lane l holds
// the sum of column  8 b4 + 4 b3 + 2 b2 + b1  (b_k = bit k of l), duplicated in the lane pair (l, l ^ 1)
__device__ __forceinline__ float tc_colsum16(const float (&v)[16], int lane) {
    float a[8];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float mine = hi ? v[j + 8] : v[j], other = hi ? v[j] : v[j + 8];
            a[j] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
        }
    }
    float b4[4];
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float mine = hi ? a[j + 4] : a[j], other = hi ? a[j] : a[j + 4];
            b4[j] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
        }
    }
    float c2[2];
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float mine = hi ? b4[j + 2] : b4[j], other = hi ? b4[j] : b4[j + 2];
            c2[j] = mine + __shfl_xor_sync(0xffffffffu, other, 4);
        }
    }
    const bool hi = lane & 2;
    const float mine = hi ? c2[1] : c2[0], other = hi ? c2[0] : c2[1];
    float d = mine + __shfl_xor_sync(0xffffffffu, other, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

constexpr int TC_RED_LD = 17;                                   // padded row of the epilogue reduction buffer
constexpr int TC_RED_LDT = 132;                                 // column pitch of the transposed variant: [16 columns][128 rows + 4]
// [16][132] fp32 (or [128][17] fp32 + [128] sample ids) followed by [8 row groups][2 kinds][16 columns] partial sums
constexpr uint32_t TC_RED_BYTES = 16 * TC_RED_LDT * 4 + 8 * 2 * 16 * 4;

// one thread = one output pixel (b, oy, ox), 16 consecutive output channels starting at n0, accumulators in v[16]
// Everything the epilogue ADDS to the accumulators of one (pixel, 16-channel chunk): bias + conditioning vector + fp32
// residual.  Split from the store so that these global loads are issued BEFORE the thread waits for the MMAs.
// COHERENT: the residual may have been written earlier by the SAME kernel (tc_chain.cu): ld.global.cg instead of the
// non-coherent path.
template <bool COHERENT = false>
__device__ __forceinline__ void tc_epilogue_addend(const TcEpi& p, int b, int oy, int ox, int n0, float (&add)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) add[j] = 0.f;
    const bool full = n0 + 16 <= p.Cout;
    if (p.bias) {
        // 16-byte loads where the 16 channels exist and the vector is aligned (the library's own arenas always are; an epilogue
        // warp issues ~40 instructions per chunk less than with scalar loads, and it is instruction-issue bound)
        if (full && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
            const float4* v = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 t = __ldg(v + j);
                add[4 * j] = t.x; add[4 * j + 1] = t.y; add[4 * j + 2] = t.z; add[4 * j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n0 + j < p.Cout) add[j] = __ldg(p.bias + n0 + j);
        }
    }
    if (p.temb) {
        const float* te = p.temb + (size_t)(p.temb_bcast ? 0 : b) * p.temb_stride + p.temb_off + n0;
        if (full && (reinterpret_cast<uintptr_t>(te) & 15) == 0) {
            const float4* v = reinterpret_cast<const float4*>(te);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 t = COHERENT ? __ldcg(v + j) : __ldg(v + j);
                add[4 * j] += t.x; add[4 * j + 1] += t.y; add[4 * j + 2] += t.z; add[4 * j + 3] += t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n0 + j < p.Cout) add[j] += COHERENT ? __ldcg(te + j) : __ldg(te + j);
        }
    }
    if (p.residual) {
        const size_t off = (((size_t)b * p.Ho + oy) * p.Wo + ox) * p.Cout + n0;
        if (full) {
            const float4* r = reinterpret_cast<const float4*>(p.residual + off);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 rv = COHERENT ? __ldcg(r + j) : __ldg(r + j);
                add[4 * j] += rv.x; add[4 * j + 1] += rv.y; add[4 * j + 2] += rv.z; add[4 * j + 3] += rv.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n0 + j < p.Cout) add[j] += COHERENT ? __ldcg(p.residual + off + j) : __ldg(p.residual + off + j);
        }
    }
}

// f[] = accumulators + addend, stored as fp32 NHWC / bf16 NHWC / fp32 NCHW
__device__ __forceinline__ void tc_epilogue_write(const TcEpi& p, const uint32_t (&v)[16], const float (&add)[16], int b, int oy,
                                                  int ox, int n0, float (&f)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + add[j];
    if (p.out_nchw) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (n0 + j < p.Cout) p.out_nchw[(((size_t)b * p.Cout + n0 + j) * p.Ho + oy) * p.Wo + ox] = f[j];
        return;
    }
    const size_t off = (((size_t)b * p.Ho + oy) * p.Wo + ox) * p.Cout + n0;
    if (n0 + 16 <= p.Cout) {
        if (p.out_f32) {
            float4* o = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
        if (p.out_b16) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                w[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            uint4* o = reinterpret_cast<uint4*>(p.out_b16 + off);
            o[0] = make_uint4(w[0], w[1], w[2], w[3]);
            o[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (n0 + j < p.Cout) {
                if (p.out_f32) p.out_f32[off + j] = f[j];
                if (p.out_b16) p.out_b16[off + j] = __float2bfloat16_rn(f[j]);
            }
        }
    }
}

// Per-channel (sum, sum of squares) of one 128-row x 16-channel epilogue chunk, per sample, accumulated into sums_out.
// Called by the 128 epilogue threads (te = 0..127, row m of the tile); `red` = TC_RED_BYTES of shared memory; `b_tile0` =
// sample of the tile's first row (rows are ordered by sample).
//  1. transposed pass: thread (channel c = te & 15, row group te >> 4) walks 16 consecutive rows in fp32;
//  2. the 8 row groups of the two first samples of the tile are folded through shared memory -> ONE fp64 atomic pair per
//     (sample, channel) and CTA (further samples of tiny images: direct atomics);
//  3. the accumulators are replicated TC_SUM_COPIES times (copy = CTA index mod copies): same-address L2 atomics
//     serialise, 34 CTAs x 8 row groups hammering one address cost ~20 us per layer before this.
__device__ __forceinline__ void tc_stats_atomic(const TcEpi& p, int copy, int b, int ch, float s, float q) {
    double* dst = p.sums_out + (((size_t)copy * p.sums_B + b) * p.Cout + ch) * 2;
    atomicAdd(dst, (double)s);
    atomicAdd(dst + 1, (double)q);
}
__device__ __forceinline__ void tc_epilogue_stats(const TcEpi& p, const float (&f)[16], bool valid, int b, int n0, int m, int te,
                                                  int b_tile0, int copy, uint8_t* red_raw) {
    float* red = reinterpret_cast<float*>(red_raw);
    int* sb = reinterpret_cast<int*>(red_raw + 128 * TC_RED_LD * 4);
#pragma unroll
    for (int j = 0; j < 16; ++j) red[m * TC_RED_LD + j] = valid ? f[j] : 0.f;
    sb[m] = valid ? b : -1;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int c = te & 15, rg = te >> 4, r0 = rg * 16;
    const bool cok = n0 + c < p.Cout;
    float ps[2] = {0.f, 0.f}, pq[2] = {0.f, 0.f};          // partials of the tile's first two samples
    {
        int cur = -1;
        float s = 0.f, q = 0.f;
        auto flush = [&]() {
            if (cur < 0) return;
            const int rel = cur - b_tile0;
            if (rel == 0 || rel == 1) { ps[rel] = s; pq[rel] = q; }
            else if (cok) tc_stats_atomic(p, copy, cur, n0 + c, s, q);
        };
        for (int r = r0; r < r0 + 16; ++r) {
            const int bb = sb[r];
            if (bb < 0) continue;
            if (bb != cur) { flush(); cur = bb; s = 0.f; q = 0.f; }
            const float v = red[r * TC_RED_LD + c];
            s += v;
            q = fmaf(v, v, q);
        }
        flush();
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");          // everyone is done reading `red`: reuse it for the fold
    // fold[rel][kind (sum|sq)][c][rg]
    red[((0 * 2 + 0) * 16 + c) * 8 + rg] = ps[0];
    red[((0 * 2 + 1) * 16 + c) * 8 + rg] = pq[0];
    red[((1 * 2 + 0) * 16 + c) * 8 + rg] = ps[1];
    red[((1 * 2 + 1) * 16 + c) * 8 + rg] = pq[1];
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (te < 32) {
        const int rel = te >> 4, cc = te & 15;
        const int bb = b_tile0 + rel;
        if (n0 + cc < p.Cout && bb < p.sums_B) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                s += red[((rel * 2 + 0) * 16 + cc) * 8 + k];
                q += red[((rel * 2 + 1) * 16 + cc) * 8 + k];
            }
            if (s != 0.f || q != 0.f) tc_stats_atomic(p, copy, bb, n0 + cc, s, q);
        }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

// The same statistics with warp shuffles instead of the transposed shared-memory pass, for tiles whose 128 rows belong to at
// most two samples (`nsamp_rows`, uniform over the CTA): per sample a masked column sum (16 shuffles each for sum and sum of
// squares), one 128-thread barrier to fold the four warps, one fp64 atomic pair per (sample, channel).
template <int BAR = 1>      // named barrier of the calling group of 128 epilogue threads
__device__ __forceinline__ void tc_epilogue_stats_shfl(const TcEpi& p, const float (&f)[16], bool valid, int b, int n0, int te,
                                                       int b_tile0, int nsamp_rows, int copy, uint8_t* red_raw) {
    float* red = reinterpret_cast<float*>(red_raw);          // [warp 4][sample 2][kind 2][16]
    const int lane = te & 31, w = te >> 5;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (s < nsamp_rows) {
            const bool mine = valid && b == b_tile0 + s;
            float v[16], q[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] = mine ? f[j] : 0.f; q[j] = v[j] * v[j]; }
            const float s1 = tc_colsum16(v, lane), s2 = tc_colsum16(q, lane);
            if (!(lane & 1)) {
                const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                red[((w * 2 + s) * 2 + 0) * 16 + col] = s1;
                red[((w * 2 + s) * 2 + 1) * 16 + col] = s2;
            }
        }
    }
    asm volatile("bar.sync %0, 128;" ::"n"(BAR) : "memory");
    if (te < 32 * nsamp_rows) {
        const int s = te >> 5, kind = (te >> 4) & 1, col = te & 15;
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) tot += red[((k * 2 + s) * 2 + kind) * 16 + col];
        const int bb = b_tile0 + s;
        if (n0 + col < p.Cout && bb < p.sums_B && tot != 0.f)
            atomicAdd(p.sums_out + (((size_t)copy * p.sums_B + bb) * p.Cout + n0 + col) * 2 + kind, (double)tot);
    }
    asm volatile("bar.sync %0, 128;" ::"n"(BAR) : "memory");            // `red` may be reused
}


// The same statistics for tiles whose 128 rows belong to ONE sample, through shared memory instead of shuffles: every thread
// stores its 16 values transposed ([column][row], pitch 132: conflict free), 128 threads then add 16 rows x 1 column each with
// four 16-byte loads, 32 threads fold the 8 row groups -> ONE fp64 atomic per (kind, channel).  ~75 instructions and two
// 128-thread barriers per chunk instead of ~190 and two: the epilogue warps are instruction-issue bound.
// `red_raw` must be 16-byte aligned.
template <int BAR = 1>
__device__ __forceinline__ void tc_epilogue_stats_smem(const TcEpi& p, const float (&f)[16], bool valid, int b, int n0, int te, int copy,
                                                       uint8_t* red_raw) {
    float* red = reinterpret_cast<float*>(red_raw);
    float* part = red + 16 * TC_RED_LDT;                        // [8 groups][2 kinds][16 columns]
#pragma unroll
    for (int j = 0; j < 16; ++j) red[j * TC_RED_LDT + te] = valid ? f[j] : 0.f;
    asm volatile("bar.sync %0, 128;" ::"n"(BAR) : "memory");
    {
        const int c = te & 15, g = te >> 4;
        const float4* src = reinterpret_cast<const float4*>(red + c * TC_RED_LDT + g * 16);
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 v = src[k];
            s += (v.x + v.y) + (v.z + v.w);
            q = fmaf(v.x, v.x, q); q = fmaf(v.y, v.y, q); q = fmaf(v.z, v.z, q); q = fmaf(v.w, v.w, q);
        }
        part[(g * 2 + 0) * 16 + c] = s;
        part[(g * 2 + 1) * 16 + c] = q;
    }
    asm volatile("bar.sync %0, 128;" ::"n"(BAR) : "memory");
    if (te < 32) {
        const int kind = te >> 4, c = te & 15;
        float tot = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) tot += part[(g * 2 + kind) * 16 + c];
        if (n0 + c < p.Cout && tot != 0.f)
            atomicAdd(p.sums_out + (((size_t)copy * p.sums_B + b) * p.Cout + n0 + c) * 2 + kind, (double)tot);
    }
    // no third barrier: the next call's first barrier orders these reads of `part` before its writes to it, and its writes to
    // `red` only follow reads that completed before the second barrier above
}

}  // namespace ds
