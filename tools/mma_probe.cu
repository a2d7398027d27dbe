// Micro-probe: what one tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, both operands in shared memory) costs the ISSUING
// thread, as a function of N and of how the issuing thread is selected:
//   variant 0: `if (threadIdx.x == 0)`              - ptxas cannot prove a single active thread and wraps every UTCHMMA
//                                                     in an ELECT / R2UR / BRA.U.ANY "waterfall" loop
//   variant 1: `if (warp == 0 && elect.sync)`       - no waterfall, descriptors still go through R2UR
//   variant 2: whole warp runs the loop, the MMA is predicated on the elected lane (uniform control flow)
// One CTA; cycles from the first issue to the mbarrier arrival of the commit, for chains of 8 and 72 MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I diffsplitting_b200/csrc -I include -o tools/mma_probe tools/mma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include "tc_ptx.cuh"
using namespace ds;

struct Res { long long issue, total; };

__device__ __forceinline__ void umma_pred(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}

template <int VARIANT>
__global__ void __launch_bounds__(128) probe(int n_mma, int N, Res* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(smem_u32(&tslot), 128);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const int PX = 152;
    // no-swizzle operands: A [plane][pixel][16 B] (LBO = PX*16, SBO = 128), B [plane][N][16 B]
    const uint32_t a_hi = (uint32_t)(128 >> 4) | (1u << 14), b_hi = a_hi;
    const uint32_t a_lo0 = ((base & 0x3FFFFu) >> 4) | ((uint32_t)((PX * 16) >> 4) << 16);
    const uint32_t b_lo0 = (((base + 96 * 1024) & 0x3FFFFu) >> 4) | ((uint32_t)((N * 16) >> 4) << 16);
    const uint32_t a_step = (2 * PX * 16) >> 4, b_step = (uint32_t)(2 * N * 16) >> 4;
    uint32_t elected = 0;
    if (tid < 32) asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
    bool run;
    if (VARIANT == 0) run = tid == 0;
    else if (VARIANT == 1) run = tid < 32 && elected;
    else run = tid < 32;
    if (run) {
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            uint32_t a_lo = a_lo0, b_lo = b_lo0;
#pragma unroll 8
            for (int i = 0; i < n_mma; ++i) {
                if ((i & 7) == 0) { a_lo = a_lo0; b_lo = b_lo0; }
                if (VARIANT == 2) umma_pred(tmem, a_lo, a_hi, b_lo, b_hi, idesc, i > 0, elected);
                else umma_bf16(tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, i > 0);
                a_lo += a_step;
                b_lo += b_step;
            }
            const long long t1 = clock64();
            if (VARIANT != 2 || elected) umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), rep & 1);
            const long long t2 = clock64();
            if (VARIANT != 2 || elected) { out->issue = t1 - t0; out->total = t2 - t0; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 128);
}

template <int V>
static int run_variant(Res* d) {
    cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024);
    for (int N : {16, 32, 64, 128}) {
        Res r8, r72;
        probe<V><<<1, 128, 161 * 1024>>>(8, N, d);
        cudaMemcpy(&r8, d, sizeof(Res), cudaMemcpyDeviceToHost);
        probe<V><<<1, 128, 161 * 1024>>>(72, N, d);
        if (cudaMemcpy(&r72, d, sizeof(Res), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        printf("variant %d N %3d | per MMA: issue %.1f total %.1f cycles  (n=8: issue %lld total %lld; n=72: issue %lld total %lld)\n", V, N,
               (r72.issue - r8.issue) / 64.0, (r72.total - r8.total) / 64.0, r8.issue, r8.total, r72.issue, r72.total);
    }
    return 0;
}

int main() {
    Res* d; cudaMalloc(&d, sizeof(Res));
    if (run_variant<0>(d) || run_variant<1>(d) || run_variant<2>(d)) return 1;
    return 0;
}
