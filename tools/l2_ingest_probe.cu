// What can an SM pull out of L2 when every SM streams the SAME buffer (the weight-streaming pattern of the TMA-fed conv)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_ingest_probe tools/l2_ingest_probe.cu && tools/l2_ingest_probe
// Every CTA (one per SM, 200 KB of shared memory) walks a buffer of S bytes in 16 KB bulk copies through a ring of 8 slots;
// nothing is computed.  mode 0: every CTA reads the same addresses in the same order;  mode 1: every CTA starts at a different
// offset (same buffer);  mode 2: disjoint per-CTA regions;  cluster size CL > 1: the CTAs of a cluster split every stage and
// multicast their part to the whole cluster (each SM issues 1 / CL of the reads and receives all of them).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

constexpr int STAGE = 16384, SLOTS = 8, GROUP = 4;

// one warp per CTA; all lanes walk the loop (the cluster barrier is warp-aligned), lane 0 issues the copies
template <int CL>
__global__ void __launch_bounds__(32, 1) probe(const uint8_t* buf, size_t S, int reps, int mode, unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
    const uint32_t bars = base + SLOTS * STAGE;
    const uint32_t rank = CL > 1 ? cluster_rank() : 0;
    const int lane = threadIdx.x;
    if (lane == 0) {
        for (int s = 0; s < SLOTS; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    const size_t nst = S / STAGE;
    const size_t total = nst * (size_t)reps;                  // a multiple of GROUP (S is a multiple of 64 KB)
    const size_t cl_id = blockIdx.x / CL;
    size_t start = 0;
    if (mode == 1) start = (cl_id * 977) % nst;
    const uint8_t* my = buf;
    if (mode == 2) my = buf + cl_id * S;
    const unsigned long long t0 = clock64();
    const size_t ngroups = total / GROUP;
    for (size_t gidx = 0; gidx < ngroups + SLOTS / GROUP; ++gidx) {
        if (gidx >= SLOTS / GROUP) {                          // the group issued SLOTS / GROUP groups ago has landed ...
            const size_t g0 = (gidx - SLOTS / GROUP) * GROUP;
            for (int k = 0; k < GROUP; ++k) {
                const size_t i = g0 + k;
                const long long w0 = clock64();
                uint32_t ok = 0;
                while (!ok) {
                    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(bars + 8 * (uint32_t)(i % SLOTS)), "r"((uint32_t)((i / SLOTS) & 1)) : "memory");
                    if (clock64() - w0 > 2000000000ll) __trap();
                }
            }
            // ... in every CTA of the cluster before anyone's multicast may overwrite those slots
            if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
        }
        if (gidx < ngroups && lane == 0) {
            for (int k = 0; k < GROUP; ++k) {
                const size_t i = gidx * GROUP + k;
                const int s = (int)(i % SLOTS);
                const uint8_t* src = my + ((start + i) % nst) * STAGE;
                mbar_expect(bars + 8 * s, STAGE);
                if (CL == 1) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + s * STAGE), "l"(src), "r"(STAGE), "r"(bars + 8 * s) : "memory");
                } else {
                    const uint32_t part = STAGE / CL;
                    const uint16_t mask = (uint16_t)((1u << CL) - 1u);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(base + s * STAGE + rank * part), "l"(src + rank * part), "r"(part), "r"(bars + 8 * s), "h"(mask) : "memory");
                }
            }
        }
        __syncwarp();
    }
    const unsigned long long t1 = clock64();
    if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CL>
static void run(const uint8_t* buf, size_t S, int reps, int mode, int sms, unsigned long long* dcy) {
    const int smem = SLOTS * STAGE + 1024 + 256;
    CK(cudaFuncSetAttribute(probe<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    const int grid = sms / CL * CL;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int it = 0; it < 2; ++it) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, probe<CL>, buf, S, reps, mode, dcy));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)S * reps * grid;
    printf("CL=%d mode=%d S=%6.1f MB reps=%3d grid=%d: %8.3f ms  %7.2f TB/s delivered to SMs  %6.1f GB/s per SM (%5.1f B/clk @1.965 GHz)  L2 reads %7.2f TB/s\n",
           CL, mode, S / 1048576.0, reps, grid, ms, bytes / ms / 1e9, bytes / ms / 1e6 / grid, bytes / ms / 1e6 / grid / 1.965, bytes / CL / ms / 1e9);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint8_t* buf;
    const size_t cap = (size_t)148 * 64 * 1048576ull;
    CK(cudaMalloc(&buf, cap));
    CK(cudaMemset(buf, 1, cap));
    unsigned long long* dcy;
    CK(cudaMalloc(&dcy, 1024 * 8));
    const size_t sizes[] = {2u << 20, 16u << 20, 48u << 20};
    for (size_t S : sizes) {
        const int reps = (int)((512ull << 20) / S);
        run<1>(buf, S, reps, 0, sms, dcy);
        run<1>(buf, S, reps, 1, sms, dcy);
        run<2>(buf, S, reps, 0, sms, dcy);
        run<4>(buf, S, reps, 0, sms, dcy);
        run<2>(buf, S, reps, 1, sms, dcy);
    }
    run<1>(buf, 4u << 20, 32, 2, sms, dcy);          // disjoint regions, each L2 resident in total: 148 * 4 MB > L2 -> HBM
    run<1>(buf, 512u << 10, 256, 2, sms, dcy);       // disjoint, 74 MB total: L2 resident
    return 0;
}
