// Fused reverse-step update with in-kernel counter-based noise.
//
// Replaces, per step, the ~10 elementwise launches + RNG launch + H2D copy of
//   sr3  p_sample / p_mean_variance / predict_start_from_noise / q_posterior  (sr3 diffusion.py:141-175)
//   ddpm p_sample                                                            (ddpm diffusion.py:179-203)
//   InDI inference_one_step                                                  (indi.py:62-69)
// Traffic: read x_t, read the UNet output, write x_{t-1} = 12 B / element; noise never touches memory.
//
// Noise replays torch.randn's CUDA stream bit-for-bit: Philox4x32-10 keyed by the generator seed, subsequence
// = torch's thread index, counter = offset/4 + round, curand_normal4's Box-Muller (logf + __sincosf), element
// (thread j, lane ii, round r) -> flat index j + T*ii + 4*T*r with T = 256*torch_grid
// (ATen/native/cuda/DistributionTemplates.h:34-90; curand_philox4x32_x.h:160-192; curand_normal.h:70-87).
#include "common.cuh"

namespace ds {

struct uint4_ { unsigned x, y, z, w; };

__device__ __forceinline__ uint4_ philox_round(uint4_ c, unsigned k0, unsigned k1) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    uint4_ r;
    r.x = hi1 ^ c.y ^ k0;
    r.y = lo1;
    r.z = hi0 ^ c.w ^ k1;
    r.w = lo0;
    return r;
}

__device__ __forceinline__ uint4_ philox4x32_10(uint4_ c, unsigned k0, unsigned k1) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        c = philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return philox_round(c, k0, k1);
}

// curand's _curand_box_muller, device branch, same expressions so that nvcc contracts them identically
__device__ __forceinline__ float2 box_muller(unsigned x, unsigned y) {
    float2 r;
    const float u = x * 2.3283064e-10f + (2.3283064e-10f / 2);
    const float v = y * (2.3283064e-10f * 6.2831855f) + ((2.3283064e-10f * 6.2831855f) / 2);
    const float s = sqrtf(-2.0f * logf(u));
    __sincosf(v, &r.x, &r.y);
    r.x *= s;
    r.y *= s;
    return r;
}

__device__ __forceinline__ float4 normal4(uint64_t seed, uint64_t subseq, uint64_t counter) {
    uint4_ c;
    c.x = (unsigned)counter;
    c.y = (unsigned)(counter >> 32);
    c.z = (unsigned)subseq;
    c.w = (unsigned)(subseq >> 32);
    const uint4_ o = philox4x32_10(c, (unsigned)seed, (unsigned)(seed >> 32));
    const float2 a = box_muller(o.x, o.y), b = box_muller(o.z, o.w);
    return make_float4(a.x, a.y, b.x, b.y);
}

constexpr int ST_THREADS = 256;

__global__ void __launch_bounds__(ST_THREADS) sampler_step_kernel(const ds_step_args a) {
    const int k = a.d_state ? a.d_state->step : a.step;
    const float* cf = a.d_coef + (size_t)k * 5;
    const float c0 = cf[0], c1 = cf[1], c2 = cf[2], c3 = cf[3], c4 = cf[4];
    const uint64_t offset = a.d_state ? a.d_state->offset : a.offset;
    const uint64_t seed = a.d_state ? a.d_state->seed : a.seed;
    const bool gen = (a.d_noise == nullptr) && !(a.skip_rng_if_zero && c4 == 0.0f);
    const int64_t T = a.rng_threads;
    const int64_t rounds = gen ? (a.numel - 1) / (4 * T) + 1 : 0;

    auto update = [&](int64_t i, float z) {
        const float x = a.d_x[i];
        const float n = a.d_net[i];
        float x0;
        if (a.mode == 0) {
            x0 = __fsub_rn(__fmul_rn(c0, x), __fmul_rn(c1, n));
            if (a.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
        } else {
            x0 = n;
        }
        const float mean = __fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, x));
        a.d_out[i] = __fadd_rn(mean, __fmul_rn(z, c4));
    };

    // Launched with programmatic stream serialisation: everything above and the first Philox / Box-Muller evaluation (the
    // bulk of this kernel's arithmetic; it needs only the loop state the PREVIOUS step's update left behind) overlap the
    // tail of the UNet's last conv; x_t and the network output are touched only after the wait.
    const int64_t items = rounds * T;
    const int64_t w_first = blockIdx.x * (int64_t)ST_THREADS + threadIdx.x;
    float4 z_first = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gen && w_first < items) {
        const int64_t r = w_first / T, j = w_first - r * T;
        z_first = normal4(seed, (uint64_t)j, offset / 4 + (uint64_t)r);
    }
    pdl_trigger();
    pdl_wait();
    if (gen) {
        // work item = (round r, torch thread j): one Philox call -> 4 elements
        for (int64_t w = w_first; w < items; w += (int64_t)gridDim.x * ST_THREADS) {
            const int64_t r = w / T, j = w - r * T;
            const float4 z = w == w_first ? z_first : normal4(seed, (uint64_t)j, offset / 4 + (uint64_t)r);
            const int64_t i0 = j + 4 * T * r;
            if (i0 < a.numel) update(i0, z.x);
            if (i0 + T < a.numel) update(i0 + T, z.y);
            if (i0 + 2 * T < a.numel) update(i0 + 2 * T, z.z);
            if (i0 + 3 * T < a.numel) update(i0 + 3 * T, z.w);
        }
    } else {
        for (int64_t i = blockIdx.x * (int64_t)ST_THREADS + threadIdx.x; i < a.numel; i += (int64_t)gridDim.x * ST_THREADS)
            update(i, a.d_noise ? a.d_noise[i] : 0.0f);
    }

    // last block to finish advances the device-side loop state (graph replay needs no host scalars)
    if (a.d_state) {
        __shared__ bool last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last = (atomicAdd(&a.d_state->done, 1u) == gridDim.x - 1);
        __syncthreads();
        if (last) {
            if (threadIdx.x == 0) {
                a.d_state->done = 0u;
                a.d_state->step = k + 1;
                if (gen) a.d_state->offset = offset + a.offset_inc;
            }
            if (a.d_time_out && a.d_time_table)
                for (int i = threadIdx.x; i < a.time_len; i += ST_THREADS) a.d_time_out[i] = a.d_time_table[k + 1];
        }
    }
}

__global__ void __launch_bounds__(ST_THREADS) randn_axpy_kernel(const float* __restrict__ base, float scale,
                                                                float* __restrict__ out, int64_t numel, uint64_t seed_h,
                                                                uint64_t offset_h, ds_sampler_state* d_state,
                                                                uint64_t offset_inc, int64_t T) {
    const uint64_t offset = d_state ? d_state->offset : offset_h;
    const uint64_t seed = d_state ? d_state->seed : seed_h;
    const int64_t rounds = (numel - 1) / (4 * T) + 1;
    const int64_t items = rounds * T;
    for (int64_t w = blockIdx.x * (int64_t)ST_THREADS + threadIdx.x; w < items; w += (int64_t)gridDim.x * ST_THREADS) {
        const int64_t r = w / T, j = w - r * T;
        const float4 z = normal4(seed, (uint64_t)j, offset / 4 + (uint64_t)r);
        const float zz[4] = {z.x, z.y, z.z, z.w};
        const int64_t i0 = j + 4 * T * r;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int64_t i = i0 + ii * T;
            if (i < numel) {
                const float nz = __fmul_rn(zz[ii], scale);
                out[i] = base ? __fadd_rn(base[i], nz) : nz;
            }
        }
    }
    if (d_state) {
        __shared__ bool last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last = (atomicAdd(&d_state->done, 1u) == gridDim.x - 1);
        __syncthreads();
        if (last && threadIdx.x == 0) {
            d_state->done = 0u;
            d_state->offset = offset + offset_inc;
        }
    }
}

static int step_grid(int64_t work) {
    int64_t g = (work + ST_THREADS - 1) / ST_THREADS;
    if (g > 148 * 8) g = 148 * 8;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace ds

extern "C" int ds_sampler_step(const ds_step_args* a, void* stream) {
    using namespace ds;
    DS_REQUIRE(a && a->d_x && a->d_net && a->d_out && a->numel > 0 && a->d_coef, "sampler_step: null argument");
    DS_REQUIRE(a->mode == 0 || a->mode == 1, "sampler_step: mode %d", a->mode);
    DS_REQUIRE(a->d_state || (a->step >= 0 && a->step < a->n_steps), "sampler_step: step %d outside [0,%d)", a->step, a->n_steps);
    DS_REQUIRE(a->d_noise || a->rng_threads > 0, "sampler_step: rng_threads must be set when noise is generated");
    DS_REQUIRE(!a->d_time_out || (a->d_time_table && a->d_state), "sampler_step: d_time_out needs d_time_table and d_state");
    int64_t work = a->numel;
    if (!a->d_noise) work = ((a->numel - 1) / (4 * (int64_t)a->rng_threads) + 1) * (int64_t)a->rng_threads;
    DS_CHECK_CUDA(launch_pdl(sampler_step_kernel, dim3((unsigned)step_grid(work)), dim3(ST_THREADS), 0, (cudaStream_t)stream, *a));
    return DS_OK;
}

extern "C" int ds_randn_axpy(const float* d_base, float scale, float* d_out, int64_t numel, uint64_t seed, uint64_t offset,
                             ds_sampler_state* d_state, uint64_t offset_inc, int rng_threads, void* stream) {
    using namespace ds;
    DS_REQUIRE(d_out && numel > 0 && rng_threads > 0, "randn_axpy: bad argument");
    const int64_t work = ((numel - 1) / (4 * (int64_t)rng_threads) + 1) * (int64_t)rng_threads;
    randn_axpy_kernel<<<step_grid(work), ST_THREADS, 0, (cudaStream_t)stream>>>(
        d_base, scale, d_out, numel, seed, offset, d_state, offset_inc, (int64_t)rng_threads);
    DS_CHECK_LAUNCH("randn_axpy");
    return DS_OK;
}
