// Shared helpers for the diffsplit_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>

#include "../../include/diffsplit_b200.h"

namespace ds {

void set_error(const char* fmt, ...);

#define DS_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ds::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DS_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define DS_CHECK_LAUNCH(what)                                                            \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            ds::set_error("launch of %s failed: %s", what, cudaGetErrorString(_e));      \
            return DS_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define DS_REQUIRE(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            ds::set_error(__VA_ARGS__);                                                  \
            return DS_ERR_INVALID;                                                       \
        }                                                                                \
    } while (0)

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the
// stream is still draining; everything it does before pdl_wait() (barrier / TMEM setup, constant weight fetches) overlaps
// the predecessor's tail; pdl_wait() returns once the predecessor grid has completed and its writes are visible.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
// ---- optional kernel timeline (DIFFSPLIT_B200_TRACE=1): every traced launch gets an id; its CTAs atomically record the
// earliest start / latest end in nanoseconds of the GPU global timer.  Works inside CUDA-graph replays.
struct TraceSlot { unsigned long long* buf; int id; };
TraceSlot trace_next(int kind);                 // host: {nullptr, 0} when tracing is off
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_begin(const TraceSlot& t) {
    if (t.buf && threadIdx.x == 0) atomicMin(t.buf + 2 * t.id, gtime_ns());
}
__device__ __forceinline__ void trace_end(const TraceSlot& t) {
    if (t.buf && threadIdx.x == 0) atomicMax(t.buf + 2 * t.id + 1, gtime_ns());
}
#endif
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// replication factor of the per-channel fp64 (sum, sumsq) GroupNorm accumulators [copies][B][C][2] (contention spreading)
constexpr int TC_SUM_COPIES = 8;

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ----------------------------------------------------------------- fp32 kernels (conv_f32.cu, norm.cu, attention.cu, temb.cu)
struct ConvSrc {
    const float* a;     // first source  (ca channels)
    const float* b;     // second source (cb channels) or nullptr: channel concat [a, b]
    int ca, cb;
    int nchw;           // 1: sources are NCHW (external boundary tensors), 0: NHWC
    int Hs, Ws;         // stored spatial size of the sources
    int up;             // 1: nearest x2 upsample folded into the gather (logical input = 2Hs x 2Ws)
};

struct ConvEpi {
    const float* bias;      // [Npad] or nullptr
    const float* temb;      // [tlen, temb_stride] conditioning vectors or nullptr
    int temb_off, temb_stride, temb_bcast;   // value = temb[(temb_bcast ? 0 : b)*temb_stride + temb_off + n]
    const float* residual;  // NHWC [B,Ho,Wo,Cout] or nullptr
    int out_nchw;           // 1: write NCHW (external), 0: NHWC
    void* out2_bf16;        // fp32 kernel only: additionally write a bf16 NHWC copy (entry conv of the bf16 path)
};

int launch_conv_f32(const ConvSrc& src, const float* w_packed /*[K][Npad]*/, int Npad, int Cout,
                    int ksize, int stride, int B, int Ho, int Wo, const ConvEpi& epi, float* out,
                    cudaStream_t st);
int launch_pack_conv_weight_f32(const float* w_oihw, float* w_packed, int Cout, int Cin, int ks, int Npad,
                                cudaStream_t st);

// entry conv (conv_entry.cu): 3x3, <= 16 input channels read from fp32 NCHW, <= 64 output channels, statistics epilogue
bool entry_conv_supported(int ca, int cb, int cout, int ks);
int launch_conv_entry(const float* xa, int ca, const float* xb, int cb, const float* w_packed, int npad, const float* bias, int cout,
                      int B, int H, int W, float* out_f32, void* out_b16, double* sums_out, cudaStream_t st);

int gn_nsplit(int B, int HW, int C);
size_t gn_scratch_bytes(int B, int G);
constexpr int GN_MAX_BATCH = 4096;
// fp32 in (the residual stream is fp32 in both precision modes); out fp32, or bf16 = tensor-core operand format.
// counters: GN_MAX_BATCH zero-initialised, self-resetting uint32 (library-owned, see gn_counters()).
int launch_groupnorm(const float* a, int ca, const float* b, int cb, const float* gamma, const float* beta,
                     void* out, int B, int HW, int G, int swish, void* scratch, unsigned* counters, int out_bf16,
                     cudaStream_t st);
int launch_gn_stats(const float* a, int ca, const float* b, int cb, int B, int HW, int G, void* scratch, unsigned* counters,
                    cudaStream_t st);
float2* gn_stats_ptr(void* scratch, int B, int G);
int launch_gn_apply_sums(const float* a, int ca, const float* b, int cb, const double* sums_a, const double* sums_b,
                         const float* gamma, const float* beta, void* out, int B, int HW, int G, int swish, int out_bf16,
                         void* scratch, cudaStream_t st);
int launch_ch_sums(const float* x, int C, int B, int HW, double* out /*[B][C][2], pre-zeroed*/, cudaStream_t st);
int gn_counters(unsigned** out);     // per-device buffer, allocated on first use (never inside graph capture)

int launch_attention(const void* qkv, void* out, int B, int N, int C, int bf16, cudaStream_t st);

struct TembParams {
    int variant;            // DS_UNET_SR3 / DS_UNET_DDPM
    int dim;                // inner_channel
    const float* inv_freq;  // ddpm: [dim/2]
    const float* w1; const float* b1;   // [4dim, dim]
    const float* w2; const float* b2;   // [dim, 4dim]
    const float* wf; const float* bf;   // all per-block projections stacked: [total, dim], [total]
    int total;
};
int launch_temb_f32(const TembParams& p, const float* time, int tlen, float* out /*[tlen,total]*/, cudaStream_t st);

}  // namespace ds
