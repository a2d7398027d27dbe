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

// one element of the update, every product / sum individually rounded in the reference's order
__device__ __forceinline__ float step_update(float x, float n, float z, float c0, float c1, float c2, float c3, float c4, int mode,
                                             int clip) {
    float x0;
    if (mode == 0) {
        x0 = __fsub_rn(__fmul_rn(c0, x), __fmul_rn(c1, n));
        if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    } else {
        x0 = n;
    }
    const float mean = __fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, x));
    return __fadd_rn(mean, __fmul_rn(z, c4));
}

// VEC: one CUDA thread = FOUR consecutive torch threads j .. j+3 (four Philox calls): for each of torch's four unrolled lanes
// ii the four results are the consecutive elements j + T ii + 4 T r .. +3, i.e. one 16-byte load of x_t, one of the network
// output and one 16-byte store, all eight loads of a work item in flight together (the scalar mapping issued 4-byte accesses
// at stride T).  The element <-> (thread, lane, round) mapping and the arithmetic are unchanged: bit-identical results.
template <bool VEC>
__global__ void __launch_bounds__(ST_THREADS) sampler_step_kernel(const ds_step_args a) {
    int k = a.d_state ? a.d_state->step : a.step;
    if (k >= a.n_steps) k = a.n_steps - 1;                  // a loop run past its table repeats the last row instead of reading beyond it
    const float* cf = a.d_coef + (size_t)k * 5;
    const bool per_sample = a.per_sample_numel > 0;         // rows indexed by sample (ddpm p_sample with unequal t)
    const float c0 = cf[0], c1 = cf[1], c2 = cf[2], c3 = cf[3], c4 = cf[4];
    const uint64_t offset = a.d_state ? a.d_state->offset : a.offset;
    const uint64_t seed = a.d_state ? a.d_state->seed : a.seed;
    const bool gen = (a.d_noise == nullptr) && !(a.skip_rng_if_zero && c4 == 0.0f);
    const int64_t T = a.rng_threads;
    const int64_t rounds = gen ? (a.numel - 1) / (4 * T) + 1 : 0;

    auto update = [&](int64_t i, float z) {
        if (per_sample) {
            const float* r = a.d_coef + (size_t)(i / a.per_sample_numel) * 5;
            a.d_out[i] = step_update(a.d_x[i], a.d_net[i], z, __ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3), __ldg(r + 4), a.mode, a.clip);
        } else {
            a.d_out[i] = step_update(a.d_x[i], a.d_net[i], z, c0, c1, c2, c3, c4, a.mode, a.clip);
        }
    };

    // Launched with programmatic stream serialisation: everything above and the first Philox / Box-Muller evaluation (the
    // bulk of this kernel's arithmetic; it needs only the loop state the PREVIOUS step's update left behind) overlap the
    // tail of the UNet's last conv; x_t and the network output are touched only after the wait.
    const int64_t w_first = blockIdx.x * (int64_t)ST_THREADS + threadIdx.x;
    if (VEC) {
        const int64_t Tq = T >> 2;                          // T = 256 * grid: a multiple of 4
        const int64_t items = rounds * Tq;
        float4 zf[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) zf[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gen && w_first < items) {
            const int64_t r = w_first / Tq, q = w_first - r * Tq;
#pragma unroll
            for (int u = 0; u < 4; ++u) zf[u] = normal4(seed, (uint64_t)(4 * q + u), offset / 4 + (uint64_t)r);
        }
        pdl_trigger();
        pdl_wait();
        if (gen) {
            for (int64_t w = w_first; w < items; w += (int64_t)gridDim.x * ST_THREADS) {
                const int64_t r = w / Tq, q = w - r * Tq;
                float4 z[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) z[u] = w == w_first ? zf[u] : normal4(seed, (uint64_t)(4 * q + u), offset / 4 + (uint64_t)r);
                const int64_t i0 = 4 * q + 4 * T * r;
                float4 xv[4], nv[4];
                bool full[4];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int64_t i = i0 + ii * T;
                    full[ii] = i + 3 < a.numel;
                    if (full[ii]) {
                        xv[ii] = *reinterpret_cast<const float4*>(a.d_x + i);
                        nv[ii] = *reinterpret_cast<const float4*>(a.d_net + i);
                    }
                }
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int64_t i = i0 + ii * T;
                    const float zz[4] = {ii == 0 ? z[0].x : ii == 1 ? z[0].y : ii == 2 ? z[0].z : z[0].w,
                                         ii == 0 ? z[1].x : ii == 1 ? z[1].y : ii == 2 ? z[1].z : z[1].w,
                                         ii == 0 ? z[2].x : ii == 1 ? z[2].y : ii == 2 ? z[2].z : z[2].w,
                                         ii == 0 ? z[3].x : ii == 1 ? z[3].y : ii == 2 ? z[3].z : z[3].w};
                    if (full[ii]) {
                        float4 o;
                        o.x = step_update(xv[ii].x, nv[ii].x, zz[0], c0, c1, c2, c3, c4, a.mode, a.clip);
                        o.y = step_update(xv[ii].y, nv[ii].y, zz[1], c0, c1, c2, c3, c4, a.mode, a.clip);
                        o.z = step_update(xv[ii].z, nv[ii].z, zz[2], c0, c1, c2, c3, c4, a.mode, a.clip);
                        o.w = step_update(xv[ii].w, nv[ii].w, zz[3], c0, c1, c2, c3, c4, a.mode, a.clip);
                        *reinterpret_cast<float4*>(a.d_out + i) = o;
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (i + u < a.numel) update(i + u, zz[u]);
                    }
                }
            }
        } else {
            const int64_t n4 = a.numel >> 2;
            for (int64_t v = w_first; v < n4; v += (int64_t)gridDim.x * ST_THREADS) {
                const float4 xv = *reinterpret_cast<const float4*>(a.d_x + 4 * v);
                const float4 nv = *reinterpret_cast<const float4*>(a.d_net + 4 * v);
                float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.d_noise) zv = *reinterpret_cast<const float4*>(a.d_noise + 4 * v);
                float4 o;
                o.x = step_update(xv.x, nv.x, zv.x, c0, c1, c2, c3, c4, a.mode, a.clip);
                o.y = step_update(xv.y, nv.y, zv.y, c0, c1, c2, c3, c4, a.mode, a.clip);
                o.z = step_update(xv.z, nv.z, zv.z, c0, c1, c2, c3, c4, a.mode, a.clip);
                o.w = step_update(xv.w, nv.w, zv.w, c0, c1, c2, c3, c4, a.mode, a.clip);
                *reinterpret_cast<float4*>(a.d_out + 4 * v) = o;
            }
            if (w_first == 0)
                for (int64_t i = n4 << 2; i < a.numel; ++i) update(i, a.d_noise ? a.d_noise[i] : 0.0f);
        }
    } else {
        const int64_t items = rounds * T;
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
            for (int64_t i = w_first; i < a.numel; i += (int64_t)gridDim.x * ST_THREADS)
                update(i, a.d_noise ? a.d_noise[i] : 0.0f);
        }
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
    DS_REQUIRE(a->n_steps > 0, "sampler_step: n_steps %d", a->n_steps);
    DS_REQUIRE(a->d_state || (a->step >= 0 && a->step < a->n_steps), "sampler_step: step %d outside [0,%d)", a->step, a->n_steps);
    DS_REQUIRE(a->d_noise || a->rng_threads > 0, "sampler_step: rng_threads must be set when noise is generated");
    DS_REQUIRE(!a->d_time_out || (a->d_time_table && a->d_state), "sampler_step: d_time_out needs d_time_table and d_state");
    DS_REQUIRE(a->per_sample_numel == 0 || (!a->d_state && a->step == 0 && !a->skip_rng_if_zero && a->per_sample_numel > 0 &&
                                            a->numel % a->per_sample_numel == 0 && a->numel / a->per_sample_numel <= a->n_steps),
               "sampler_step: per-sample rows need d_state == NULL, step 0, skip_rng_if_zero 0 and numel = n_steps * per_sample_numel");
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec = a->per_sample_numel == 0 && al16(a->d_x) && al16(a->d_net) && al16(a->d_out) && (!a->d_noise || al16(a->d_noise)) &&
                     (a->d_noise || a->rng_threads % 4 == 0);
    int64_t work;
    if (a->d_noise) work = vec ? (a->numel + 3) / 4 : a->numel;
    else work = ((a->numel - 1) / (4 * (int64_t)a->rng_threads) + 1) * (int64_t)(vec ? a->rng_threads / 4 : a->rng_threads);
    if (vec)
        DS_CHECK_CUDA(launch_pdl(sampler_step_kernel<true>, dim3((unsigned)step_grid(work)), dim3(ST_THREADS), 0, (cudaStream_t)stream, *a));
    else
        DS_CHECK_CUDA(launch_pdl(sampler_step_kernel<false>, dim3((unsigned)step_grid(work)), dim3(ST_THREADS), 0, (cudaStream_t)stream, *a));
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
