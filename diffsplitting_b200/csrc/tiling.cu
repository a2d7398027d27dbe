// Tile index arithmetic, tile gather and stitching.
//
// Reference: TileIndexManager (data/tiling_manager.py:14-191), SplitDatasetTiledPred.patch_location
// (data/split_dataset_tiledpred.py:27-32) + the crop of SplitDataset.__getitem__ (data/split_dataset.py:239-249),
// stitch_predictions (data/tile_stitcher.py:10-81).  Everything is integer index work + copies: bit-exact.
//
// The reference's np.ceil / np.floor of float quotients are replaced by exact integer division.  Stitching
// is PULL based: one thread per output pixel finds the tile that the reference's sequential loop would have
// written last (the highest tile index whose destination box covers the pixel), so overlapping boxes
// (ShiftBoundary's shifted last row/column) resolve exactly as in the reference, without write races.
#include "common.cuh"

namespace ds {

struct TileGeom {
    int data[3], grid[3], patch[3];
    int mode;
    int counts[3];
    int64_t strides[3];
    int64_t total;
    int off[3];
};

__host__ __device__ inline int floordiv(int a, int b) {
    int q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
}
__host__ __device__ inline int ceildiv(int a, int b) { return -floordiv(-a, b); }

static int make_geom(const int32_t data[3], const int32_t grid[3], const int32_t patch[3], int mode, TileGeom* g) {
    DS_REQUIRE(mode == DS_TILE_TRIM || mode == DS_TILE_PAD || mode == DS_TILE_SHIFT, "tiling: unknown mode %d", mode);
    for (int d = 0; d < 3; ++d) {
        DS_REQUIRE(data[d] > 0 && grid[d] > 0 && patch[d] > 0, "tiling: non-positive shape in dim %d", d);
        DS_REQUIRE(patch[d] >= grid[d], "tiling: patch %d < grid %d in dim %d", patch[d], grid[d], d);
        DS_REQUIRE((patch[d] - grid[d]) % 2 == 0, "tiling: odd padding in dim %d", d);
        g->data[d] = data[d];
        g->grid[d] = grid[d];
        g->patch[d] = patch[d];
        g->off[d] = (patch[d] - grid[d]) / 2;                          // patch_offset, tiling_manager.py:31-32
    }
    g->mode = mode;
    for (int d = 0; d < 3; ++d) {                                      // get_individual_dim_grid_count :34-50
        const int D = data[d], G = grid[d], P = patch[d];
        int c;
        if (G == 1 && P == 1) c = D;
        else if (mode == DS_TILE_PAD) c = ceildiv(D, G);
        else if (mode == DS_TILE_SHIFT) c = ceildiv(D - (P - G), G);
        else c = floordiv(D - (P - G), G);
        DS_REQUIRE(c > 0, "tiling: data dim %d (%d) too small for patch %d", d, D, P);
        g->counts[d] = c;
    }
    g->strides[2] = 1;                                                 // grid_count :58-67
    g->strides[1] = g->counts[2];
    g->strides[0] = (int64_t)g->counts[1] * g->counts[2];
    g->total = g->strides[0] * g->counts[0];                           // total_grid_count :52-56
    return DS_OK;
}

// get_gridstart_location_from_dim_index, tiling_manager.py:120-143
__host__ __device__ inline int grid_start(const TileGeom& g, int d, int k) {
    const int G = g.grid[d], P = g.patch[d];
    if (G == 1 && P == 1) return k;
    if (g.mode == DS_TILE_PAD) return k * G;
    const int ex = (P - G) / 2;
    if (g.mode == DS_TILE_TRIM || k < g.counts[d] - 1) return k * G + ex;
    return g.data[d] - G - ex;
}

// destination interval [lo, hi) that tile k writes along dim d (tile_stitcher.py:28-52)
__host__ __device__ inline void dst_interval(const TileGeom& g, int d, int k, int* lo, int* hi) {
    const int gs = grid_start(g, d, k);
    int vs = gs, ve = gs + g.grid[d];
    if (g.mode == DS_TILE_SHIFT) {
        const int ps = gs - g.off[d];
        if (ps == 0) vs = 0;
        if (ps + g.patch[d] == g.data[d]) ve = g.data[d];
    }
    *lo = vs;
    *hi = ve;
}

// highest k whose destination interval covers coordinate y, or -1
__host__ __device__ inline int owner(const TileGeom& g, int d, int y) {
    const int n = g.counts[d];
    int lo, hi;
    dst_interval(g, d, n - 1, &lo, &hi);
    if (y >= lo && y < hi) return n - 1;
    const int G = g.grid[d], P = g.patch[d];
    int k;
    if (G == 1 && P == 1) k = y;
    else if (g.mode == DS_TILE_PAD) k = y / G;
    else k = floordiv(y - (P - G) / 2, G);
    if (k < 0) k = 0;
    if (k > n - 1) k = n - 1;
    // regular cells are disjoint, so only k (or, for an extended first box, 0) can cover y
    for (int kk = k; kk >= 0 && kk >= k - 1; --kk) {
        dst_interval(g, d, kk, &lo, &hi);
        if (y >= lo && y < hi) return kk;
    }
    return -1;
}

__global__ void __launch_bounds__(256) stitch_kernel(const float* __restrict__ tiles, int C, const TileGeom g,
                                                     float* __restrict__ out) {
    const int W = g.data[2], H = g.data[1];
    const int64_t npix = (int64_t)g.data[0] * H * W;
    const int P1 = g.patch[1], P2 = g.patch[2];
    for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < npix; i += (int64_t)gridDim.x * 256) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int f = (int)(i / ((int64_t)W * H));
        const int kf = owner(g, 0, f), ky = owner(g, 1, y), kx = owner(g, 2, x);
        float* dst = out + i * C;
        if (kf < 0 || ky < 0 || kx < 0) {
            for (int c = 0; c < C; ++c) dst[c] = 0.f;
            continue;
        }
        const int64_t n = kf * g.strides[0] + ky * g.strides[1] + kx;
        const int py = y - (grid_start(g, 1, ky) - g.off[1]);
        const int px = x - (grid_start(g, 2, kx) - g.off[2]);
        const float* src = tiles + ((n * C) * P1 + py) * (int64_t)P2 + px;
        for (int c = 0; c < C; ++c) dst[c] = src[(int64_t)c * P1 * P2];
    }
}

// ---- multi-GPU exchange format: only the part of a predicted tile that the stitcher reads (its destination box: the inner
// grid cell, widened to the frame edge where the patch touches it, tile_stitcher.py:28-57) is packed as [C][hy][hx] and
// travels; 490 x 2 x 512^2 tiles = 1.03 GB shrink to the 335 MB of the stitched frames.  Tiles are addressed by their GLOBAL
// index, so the packed buffers of any partition of the tiles stitch to the same bits.
__global__ void __launch_bounds__(256) pack_regions_kernel(const float* __restrict__ tiles, int C, const TileGeom g,
                                                           const int64_t* __restrict__ tile_ids, const int64_t* __restrict__ offsets,
                                                           int64_t n, float* __restrict__ packed) {
    const int P1 = g.patch[1], P2 = g.patch[2];
    // one CTA walks one (tile, channel) region at a time: its geometry is CTA-uniform
    for (int64_t job = blockIdx.x; job < n * C; job += gridDim.x) {
        const int64_t ti = job / C;
        const int c = (int)(job - ti * C);
        int64_t idx = tile_ids[ti];
        idx -= (idx / g.strides[0]) * g.strides[0];
        const int ky = (int)(idx / g.strides[1]), kx = (int)(idx - (int64_t)ky * g.strides[1]);
        int ylo, yhi, xlo, xhi;
        dst_interval(g, 1, ky, &ylo, &yhi);
        dst_interval(g, 2, kx, &xlo, &xhi);
        const int hy = yhi - ylo, hx = xhi - xlo;
        const int py0 = ylo - (grid_start(g, 1, ky) - g.off[1]), px0 = xlo - (grid_start(g, 2, kx) - g.off[2]);
        const float* src = tiles + ((ti * C + c) * P1 + py0) * (int64_t)P2 + px0;
        float* dst = packed + offsets[ti] + (int64_t)c * hy * hx;
        for (int i = threadIdx.x; i < hy * hx; i += 256) {
            const int y = i / hx, x = i - y * hx;
            dst[i] = src[(int64_t)y * P2 + x];
        }
    }
}

__global__ void __launch_bounds__(256) stitch_packed_kernel(const float* __restrict__ packed, const int64_t* __restrict__ tile_off,
                                                            int C, const TileGeom g, float* __restrict__ out) {
    const int W = g.data[2], H = g.data[1];
    const int64_t npix = (int64_t)g.data[0] * H * W;
    for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < npix; i += (int64_t)gridDim.x * 256) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int f = (int)(i / ((int64_t)W * H));
        const int kf = owner(g, 0, f), ky = owner(g, 1, y), kx = owner(g, 2, x);
        float* dst = out + i * C;
        if (kf < 0 || ky < 0 || kx < 0) {
            for (int c = 0; c < C; ++c) dst[c] = 0.f;
            continue;
        }
        const int64_t n = kf * g.strides[0] + ky * g.strides[1] + kx;
        int ylo, yhi, xlo, xhi;
        dst_interval(g, 1, ky, &ylo, &yhi);
        dst_interval(g, 2, kx, &xlo, &xhi);
        const int hy = yhi - ylo, hx = xhi - xlo;
        const float* src = packed + tile_off[n] + (int64_t)(y - ylo) * hx + (x - xlo);
        for (int c = 0; c < C; ++c) dst[c] = src[(int64_t)c * hy * hx];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) crop_kernel(const T* __restrict__ frames, int C, const TileGeom g, int64_t first,
                                                   int64_t n, float* __restrict__ tiles) {
    const int P1 = g.patch[1], P2 = g.patch[2];
    const int H = g.data[1], W = g.data[2], F = g.data[0];
    const int64_t per = (int64_t)C * P1 * P2;
    const int64_t total = n * per;
    for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t ti = i / per;
        int64_t r = i - ti * per;
        const int c = (int)(r / ((int64_t)P1 * P2));
        r -= (int64_t)c * P1 * P2;
        const int py = (int)(r / P2), px = (int)(r - (int64_t)py * P2);
        int64_t idx = first + ti;
        const int kf = (int)(idx / g.strides[0]);
        idx -= kf * g.strides[0];
        const int ky = (int)(idx / g.strides[1]);
        const int kx = (int)(idx - ky * g.strides[1]);
        const int f = grid_start(g, 0, kf) - g.off[0];
        const int y = grid_start(g, 1, ky) - g.off[1] + py;
        const int x = grid_start(g, 2, kx) - g.off[2] + px;
        float v = 0.f;
        if (f >= 0 && f < F && y >= 0 && y < H && x >= 0 && x < W)
            v = (float)frames[(((int64_t)c * F + f) * H + y) * W + x];
        tiles[i] = v;
    }
}


// ---- fused tile loader: crop + normalise + mix (SplitDataset.__getitem__, data/split_dataset.py:237-278, without transforms)
// patch_c = frames[c, f, y0:y0+P, x0:x0+P].astype(float32); target_c = float32((float64(patch_c) - mean_t[c]) / std_t[c])
// (normalize_target :199-201: a float32 array against float64 constants promotes to float64, one rounding at the astype);
// input = w0 * target_0 + w1 * target_1 in float32 (:268-269, python-scalar weights stay weak) or
// float32((float64(w0 * patch_0 + w1 * patch_1) - mean_inp) / std_inp) (:270-272).  Every operation is a single IEEE
// operation in the reference, so the intrinsics below (no FMA contraction) reproduce it bit for bit.
struct TileNorm {
    double mean_t[2], std_t[2], mean_i, std_i;
    float w0, w1;
    int from_norm_target;
};

template <typename T>
__global__ void __launch_bounds__(256) tile_batch_kernel(const T* __restrict__ frames, const TileGeom g, int64_t first,
                                                         int64_t n, const TileNorm nm, float* __restrict__ inp,
                                                         float* __restrict__ target) {
    const int P1 = g.patch[1], P2 = g.patch[2];
    const int H = g.data[1], W = g.data[2], F = g.data[0];
    const int64_t plane = (int64_t)P1 * P2;
    const int64_t chan = (int64_t)F * H * W;
    const int xq = P2 >> 2;                                            // 4 pixels per thread (P2 % 4 == 0)
    const int64_t total = n * P1 * (int64_t)xq;
    for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t ti = i / ((int64_t)P1 * xq);
        int64_t r = i - ti * (int64_t)P1 * xq;
        const int py = (int)(r / xq), px = (int)(r - (int64_t)py * xq) * 4;
        int64_t idx = first + ti;
        const int kf = (int)(idx / g.strides[0]);
        idx -= kf * g.strides[0];
        const int ky = (int)(idx / g.strides[1]);
        const int kx = (int)(idx - ky * g.strides[1]);
        const int f = grid_start(g, 0, kf) - g.off[0];
        const int y = grid_start(g, 1, ky) - g.off[1] + py;
        const int x = grid_start(g, 2, kx) - g.off[2] + px;
        float vi[4], v0[4], v1[4];
        const bool row_in = f >= 0 && f < F && y >= 0 && y < H;
        const T* src = frames + ((int64_t)f * H + y) * W + x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float p0 = 0.f, p1 = 0.f;
            if (row_in && x + j >= 0 && x + j < W) {
                p0 = (float)__ldg(src + j);
                p1 = (float)__ldg(src + chan + j);
            }
            const float t0 = (float)__ddiv_rn(__dsub_rn((double)p0, nm.mean_t[0]), nm.std_t[0]);
            const float t1 = (float)__ddiv_rn(__dsub_rn((double)p1, nm.mean_t[1]), nm.std_t[1]);
            v0[j] = t0;
            v1[j] = t1;
            if (nm.from_norm_target) {
                vi[j] = __fadd_rn(__fmul_rn(nm.w0, t0), __fmul_rn(nm.w1, t1));
            } else {
                const float mix = __fadd_rn(__fmul_rn(nm.w0, p0), __fmul_rn(nm.w1, p1));
                vi[j] = (float)__ddiv_rn(__dsub_rn((double)mix, nm.mean_i), nm.std_i);
            }
        }
        const int64_t o = (int64_t)py * P2 + px;
        *reinterpret_cast<float4*>(inp + ti * plane + o) = make_float4(vi[0], vi[1], vi[2], vi[3]);
        if (target) {
            *reinterpret_cast<float4*>(target + (ti * 2) * plane + o) = make_float4(v0[0], v0[1], v0[2], v0[3]);
            *reinterpret_cast<float4*>(target + (ti * 2 + 1) * plane + o) = make_float4(v1[0], v1[1], v1[2], v1[3]);
        }
    }
}

// ---- PSNR / RangeInvariantPsnr of whole images (core/psnr.py:46-82) with the caller's un-normalisation
// (split.py:198-203: v * std + mean in float64, prediction clamped to [0, 65535], both truncated to uint16) folded in.
// One pass: per image the 5 shifted moments of (gt, pred), the sum of squared differences, min and max of gt, in
// fp64; the two metrics follow in closed form (the zero-mean / rescale steps of RangeInvariantPsnr reduce to
// 10 log10(range^2 n / (Sgg - Sgp^2 / Spp)); the standard deviation cancels).
constexpr int PSNR_VALS = 8;        // sum dg, sum dg^2, sum dp, sum dp^2, sum dg dp, sum (g - p)^2, min g, max g
struct PsnrView {
    const float* gt;
    const float* pred;
    int64_t gt_fs, gt_cs, gt_es, pr_fs, pr_cs, pr_es;                 // frame / channel / pixel strides in elements
    int C;
    int64_t npix;
    double scale[4], offset[4];
    int unnorm, quantize;
};

__device__ __forceinline__ double psnr_value(const PsnrView& v, float raw, int c, bool clamp) {
    double x = (double)raw;
    if (v.unnorm) x = __dadd_rn(__dmul_rn(x, v.scale[c]), v.offset[c]);
    if (v.quantize) {
        if (clamp) x = fmin(fmax(x, 0.0), 65535.0);
        x = trunc(x);
    }
    return x;
}

__global__ void __launch_bounds__(256) psnr_partial_kernel(const PsnrView v, double* __restrict__ partial) {
    const int img = blockIdx.y, f = img / v.C, c = img - f * v.C;
    const float* g = v.gt + f * v.gt_fs + c * v.gt_cs;
    const float* p = v.pred + f * v.pr_fs + c * v.pr_cs;
    const double g0 = psnr_value(v, __ldg(g), c, false), p0 = psnr_value(v, __ldg(p), c, true);
    double a[PSNR_VALS] = {0, 0, 0, 0, 0, 0, g0, g0};
    // 4 strided pixels per thread and iteration: 8 independent loads in flight before the fp64 arithmetic
    const int64_t step = (int64_t)gridDim.x * 256;
    for (int64_t k = blockIdx.x * 256LL + threadIdx.x; k < v.npix; k += 4 * step) {
        float gr[4], pr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t kk = k + j * step;
            const bool in = kk < v.npix;
            gr[j] = in ? __ldg(g + kk * v.gt_es) : 0.f;
            pr[j] = in ? __ldg(p + kk * v.pr_es) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (k + j * step >= v.npix) break;
            const double gv = psnr_value(v, gr[j], c, false);
            const double pv = psnr_value(v, pr[j], c, true);
            const double dg = gv - g0, dp = pv - p0, d = gv - pv;
            a[0] += dg;
            a[1] = fma(dg, dg, a[1]);
            a[2] += dp;
            a[3] = fma(dp, dp, a[3]);
            a[4] = fma(dg, dp, a[4]);
            a[5] = fma(d, d, a[5]);
            a[6] = fmin(a[6], gv);
            a[7] = fmax(a[7], gv);
        }
    }
    __shared__ double red[8][PSNR_VALS];
#pragma unroll
    for (int i = 0; i < PSNR_VALS; ++i) {
        double x = a[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_xor_sync(0xffffffffu, x, o);
            x = i < 6 ? x + y : (i == 6 ? fmin(x, y) : fmax(x, y));
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < PSNR_VALS) {
        const int i = threadIdx.x;
        double x = red[0][i];
        for (int w = 1; w < 8; ++w) x = i < 6 ? x + red[w][i] : (i == 6 ? fmin(x, red[w][i]) : fmax(x, red[w][i]));
        partial[((int64_t)img * gridDim.x + blockIdx.x) * PSNR_VALS + i] = x;
    }
}

// out[img] = (PSNR, RangeInvariantPsnr, mse, range)
__global__ void __launch_bounds__(32) psnr_final_kernel(const double* __restrict__ partial, int nblk, int64_t npix,
                                                        float* __restrict__ out) {
    const int img = blockIdx.x, lane = threadIdx.x;
    double a[PSNR_VALS];
    const double* pp = partial + (int64_t)img * nblk * PSNR_VALS;
#pragma unroll
    for (int i = 0; i < PSNR_VALS; ++i) a[i] = i < 6 ? 0.0 : pp[i];
    for (int b = lane; b < nblk; b += 32) {
#pragma unroll
        for (int i = 0; i < PSNR_VALS; ++i) {
            const double y = pp[(int64_t)b * PSNR_VALS + i];
            a[i] = i < 6 ? a[i] + y : (i == 6 ? fmin(a[i], y) : fmax(a[i], y));
        }
    }
#pragma unroll
    for (int i = 0; i < PSNR_VALS; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_xor_sync(0xffffffffu, a[i], o);
            a[i] = i < 6 ? a[i] + y : (i == 6 ? fmin(a[i], y) : fmax(a[i], y));
        }
    }
    if (lane == 0) {
        const double n = (double)npix;
        const double sgg = a[1] - a[0] * a[0] / n, spp = a[3] - a[2] * a[2] / n, sgp = a[4] - a[0] * a[2] / n;
        const double range = a[7] - a[6], mse = a[5] / n;
        const double resid = sgg - sgp * sgp / spp;
        out[img * 4 + 0] = (float)(20.0 * log10(range / sqrt(mse)));
        out[img * 4 + 1] = (float)(10.0 * log10(range * range * n / resid));
        out[img * 4 + 2] = (float)mse;
        out[img * 4 + 3] = (float)range;
    }
}

static int psnr_blocks(int n_images, int64_t npix) {
    int64_t want = (npix + 256 * 8 - 1) / (256 * 8);
    int64_t cap = (148 * 8 + n_images - 1) / n_images;
    if (cap < 1) cap = 1;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

}  // namespace ds

extern "C" int ds_tile_counts(const int32_t data_shape[3], const int32_t grid_shape[3], const int32_t patch_shape[3],
                              int mode, int32_t counts[3], int64_t* total) {
    ds::TileGeom g;
    int rc = ds::make_geom(data_shape, grid_shape, patch_shape, mode, &g);
    if (rc != DS_OK) return rc;
    for (int d = 0; d < 3; ++d) counts[d] = g.counts[d];
    if (total) *total = g.total;
    return DS_OK;
}

extern "C" int ds_tile_patch_locations(const int32_t data_shape[3], const int32_t grid_shape[3],
                                       const int32_t patch_shape[3], int mode, int64_t first, int64_t n,
                                       int32_t* h_locations) {
    using namespace ds;
    TileGeom g;
    int rc = make_geom(data_shape, grid_shape, patch_shape, mode, &g);
    if (rc != DS_OK) return rc;
    DS_REQUIRE(first >= 0 && n >= 0 && first + n <= g.total, "tiling: tile range [%lld, %lld) outside [0, %lld)",
               (long long)first, (long long)(first + n), (long long)g.total);
    for (int64_t i = 0; i < n; ++i) {                                  // get_location_from_dataset_idx :145-154
        int64_t idx = first + i;
        for (int d = 0; d < 3; ++d) {
            const int k = (int)(idx / g.strides[d]);
            idx %= g.strides[d];
            h_locations[i * 3 + d] = grid_start(g, d, k) - g.off[d];   // get_patch_location_from_dataset_idx :106-112
        }
    }
    return DS_OK;
}

extern "C" int ds_crop_tiles(const void* d_frames, int elem_size, int C, const int32_t data_shape[3],
                             const int32_t grid_shape[3], const int32_t patch_shape[3], int mode, int64_t first,
                             int64_t n, float* d_tiles, void* stream) {
    using namespace ds;
    TileGeom g;
    int rc = make_geom(data_shape, grid_shape, patch_shape, mode, &g);
    if (rc != DS_OK) return rc;
    DS_REQUIRE(d_frames && d_tiles && C > 0, "crop_tiles: null argument");
    DS_REQUIRE(elem_size == 4 || elem_size == 2, "crop_tiles: elem_size %d (4 = fp32, 2 = uint16)", elem_size);
    DS_REQUIRE(first >= 0 && n >= 0 && first + n <= g.total, "crop_tiles: tile range outside [0, %lld)", (long long)g.total);
    if (n == 0) return DS_OK;
    const int64_t total = n * C * (int64_t)g.patch[1] * g.patch[2];
    int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    if (elem_size == 4)
        crop_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)d_frames, C, g, first, n, d_tiles);
    else
        crop_kernel<unsigned short><<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned short*)d_frames, C, g, first, n, d_tiles);
    DS_CHECK_LAUNCH("crop_tiles");
    return DS_OK;
}

extern "C" int ds_stitch_tiles(const float* d_tiles, int C, const int32_t data_shape[3], const int32_t grid_shape[3],
                               const int32_t patch_shape[3], int mode, float* d_out, void* stream) {
    using namespace ds;
    TileGeom g;
    int rc = make_geom(data_shape, grid_shape, patch_shape, mode, &g);
    if (rc != DS_OK) return rc;
    DS_REQUIRE(d_tiles && d_out && C > 0, "stitch_tiles: null argument");
    DS_REQUIRE(g.patch[0] == 1 && g.grid[0] == 1, "stitch_tiles: only (1,P,P) patches over (F,H,W) data are supported");
    if (mode != DS_TILE_SHIFT) {
        // the reference asserts every grid cell lies inside the data (tile_stitcher.py:41-42)
        for (int d = 0; d < 3; ++d) {
            const int last = grid_start(g, d, g.counts[d] - 1);
            DS_REQUIRE(last + g.grid[d] <= g.data[d] && grid_start(g, d, 0) >= 0,
                       "stitch_tiles: grid cell outside the data in dim %d (reference asserts)", d);
        }
    }
    const int64_t npix = (int64_t)g.data[0] * g.data[1] * g.data[2];
    int blocks = (int)((npix + 255) / 256 > 148 * 16 ? 148 * 16 : (npix + 255) / 256);
    stitch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_tiles, C, g, d_out);
    DS_CHECK_LAUNCH("stitch_tiles");
    return DS_OK;
}

static int stitch_geom(const int32_t data_shape[3], const int32_t grid_shape[3], const int32_t patch_shape[3], int mode, ds::TileGeom* g,
                       const char* who) {
    using namespace ds;
    int rc = make_geom(data_shape, grid_shape, patch_shape, mode, g);
    if (rc != DS_OK) return rc;
    DS_REQUIRE(g->patch[0] == 1 && g->grid[0] == 1, "%s: only (1,P,P) patches over (F,H,W) data are supported", who);
    if (mode != DS_TILE_SHIFT) {
        for (int d = 0; d < 3; ++d) {
            const int last = grid_start(*g, d, g->counts[d] - 1);
            DS_REQUIRE(last + g->grid[d] <= g->data[d] && grid_start(*g, d, 0) >= 0,
                       "%s: grid cell outside the data in dim %d (reference asserts)", who, d);
        }
    }
    return DS_OK;
}

extern "C" int ds_tile_regions(const int32_t data_shape[3], const int32_t grid_shape[3], const int32_t patch_shape[3], int mode,
                               int64_t first, int64_t n, int32_t* h_regions) {
    using namespace ds;
    TileGeom g;
    int rc = stitch_geom(data_shape, grid_shape, patch_shape, mode, &g, "tile_regions");
    if (rc != DS_OK) return rc;
    DS_REQUIRE(h_regions && first >= 0 && n >= 0 && first + n <= g.total, "tile_regions: tile range outside [0, %lld)", (long long)g.total);
    for (int64_t i = 0; i < n; ++i) {
        int64_t idx = first + i;
        const int kf = (int)(idx / g.strides[0]);
        idx -= kf * g.strides[0];
        const int ky = (int)(idx / g.strides[1]), kx = (int)(idx - (int64_t)ky * g.strides[1]);
        int lo, hi;
        dst_interval(g, 0, kf, &lo, &hi);
        h_regions[i * 5 + 0] = lo;
        dst_interval(g, 1, ky, &lo, &hi);
        h_regions[i * 5 + 1] = lo; h_regions[i * 5 + 2] = hi;
        dst_interval(g, 2, kx, &lo, &hi);
        h_regions[i * 5 + 3] = lo; h_regions[i * 5 + 4] = hi;
    }
    return DS_OK;
}

extern "C" int ds_pack_tile_regions(const float* d_tiles, int C, const int32_t data_shape[3], const int32_t grid_shape[3],
                                    const int32_t patch_shape[3], int mode, const int64_t* d_tile_ids, const int64_t* d_offsets,
                                    int64_t n, float* d_packed, void* stream) {
    using namespace ds;
    TileGeom g;
    int rc = stitch_geom(data_shape, grid_shape, patch_shape, mode, &g, "pack_tile_regions");
    if (rc != DS_OK) return rc;
    DS_REQUIRE(C > 0 && n >= 0, "pack_tile_regions: bad argument");
    if (n == 0) return DS_OK;
    DS_REQUIRE(d_tiles && d_tile_ids && d_offsets && d_packed, "pack_tile_regions: null argument");
    const int64_t jobs = n * C;
    const int blocks = (int)(jobs > 148 * 8 ? 148 * 8 : jobs);
    pack_regions_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_tiles, C, g, d_tile_ids, d_offsets, n, d_packed);
    DS_CHECK_LAUNCH("pack_tile_regions");
    return DS_OK;
}

extern "C" int ds_stitch_packed(const float* d_packed, const int64_t* d_tile_offsets, int C, const int32_t data_shape[3],
                                const int32_t grid_shape[3], const int32_t patch_shape[3], int mode, float* d_out, void* stream) {
    using namespace ds;
    TileGeom g;
    int rc = stitch_geom(data_shape, grid_shape, patch_shape, mode, &g, "stitch_packed");
    if (rc != DS_OK) return rc;
    DS_REQUIRE(d_packed && d_tile_offsets && d_out && C > 0, "stitch_packed: null argument");
    const int64_t npix = (int64_t)g.data[0] * g.data[1] * g.data[2];
    int blocks = (int)((npix + 255) / 256 > 148 * 16 ? 148 * 16 : (npix + 255) / 256);
    stitch_packed_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_packed, d_tile_offsets, C, g, d_out);
    DS_CHECK_LAUNCH("stitch_packed");
    return DS_OK;
}

extern "C" int ds_tile_batch(const void* d_frames, int elem_size, const int32_t data_shape[3], const int32_t grid_shape[3],
                             const int32_t patch_shape[3], int mode, int64_t first, int64_t n, const ds_tile_norm* norm,
                             float* d_input, float* d_target, void* stream) {
    using namespace ds;
    TileGeom g;
    int rc = make_geom(data_shape, grid_shape, patch_shape, mode, &g);
    if (rc != DS_OK) return rc;
    DS_REQUIRE(d_frames && d_input && norm, "tile_batch: null argument");
    DS_REQUIRE(elem_size == 4 || elem_size == 2, "tile_batch: elem_size %d (4 = fp32, 2 = uint16)", elem_size);
    DS_REQUIRE(g.patch[0] == 1 && g.grid[0] == 1, "tile_batch: only (1,P,P) patches over (F,H,W) data are supported");
    DS_REQUIRE(g.patch[2] % 4 == 0, "tile_batch: patch width %d is not a multiple of 4", g.patch[2]);
    DS_REQUIRE(first >= 0 && n >= 0 && first + n <= g.total, "tile_batch: tile range outside [0, %lld)", (long long)g.total);
    DS_REQUIRE(norm->std_target[0] != 0.0 && norm->std_target[1] != 0.0 && norm->std_input != 0.0, "tile_batch: zero std");
    if (n == 0) return DS_OK;
    TileNorm nm;
    for (int c = 0; c < 2; ++c) { nm.mean_t[c] = norm->mean_target[c]; nm.std_t[c] = norm->std_target[c]; }
    nm.mean_i = norm->mean_input;
    nm.std_i = norm->std_input;
    nm.w0 = norm->w0;
    nm.w1 = norm->w1;
    nm.from_norm_target = norm->input_from_normalized_target;
    const int64_t total = n * g.patch[1] * (int64_t)(g.patch[2] / 4);
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    if (elem_size == 4)
        tile_batch_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)d_frames, g, first, n, nm, d_input, d_target);
    else
        tile_batch_kernel<unsigned short><<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned short*)d_frames, g, first, n, nm, d_input, d_target);
    DS_CHECK_LAUNCH("tile_batch");
    return DS_OK;
}

extern "C" size_t ds_psnr_workspace_bytes(int n_frames, int C, int64_t npix) {
    if (n_frames <= 0 || C <= 0 || npix <= 0) return 0;
    return (size_t)n_frames * C * ds::psnr_blocks(n_frames * C, npix) * ds::PSNR_VALS * sizeof(double);
}

extern "C" int ds_psnr(const ds_psnr_args* a, void* stream) {
    using namespace ds;
    DS_REQUIRE(a && a->d_gt && a->d_pred && a->d_out && a->d_workspace, "psnr: null argument");
    DS_REQUIRE(a->n_frames > 0 && a->C > 0 && a->C <= 4 && a->npix > 0, "psnr: %d frames x %d channels (1..4) x %lld pixels",
               a->n_frames, a->C, (long long)a->npix);
    DS_REQUIRE(a->workspace_bytes >= ds_psnr_workspace_bytes(a->n_frames, a->C, a->npix), "psnr: workspace too small");
    DS_REQUIRE(!a->unnormalize || (a->scale && a->offset), "psnr: unnormalize without scale / offset");
    PsnrView v;
    v.gt = a->d_gt; v.pred = a->d_pred;
    v.gt_fs = a->gt_frame_stride; v.gt_cs = a->gt_channel_stride; v.gt_es = a->gt_pixel_stride;
    v.pr_fs = a->pred_frame_stride; v.pr_cs = a->pred_channel_stride; v.pr_es = a->pred_pixel_stride;
    v.C = a->C; v.npix = a->npix; v.unnorm = a->unnormalize; v.quantize = a->quantize_u16;
    for (int c = 0; c < 4; ++c) {
        v.scale[c] = a->unnormalize && c < a->C ? a->scale[c] : 1.0;
        v.offset[c] = a->unnormalize && c < a->C ? a->offset[c] : 0.0;
    }
    const int n_images = a->n_frames * a->C;
    const int nblk = psnr_blocks(n_images, a->npix);
    psnr_partial_kernel<<<dim3(nblk, n_images), 256, 0, (cudaStream_t)stream>>>(v, (double*)a->d_workspace);
    DS_CHECK_LAUNCH("psnr_partial");
    psnr_final_kernel<<<n_images, 32, 0, (cudaStream_t)stream>>>((const double*)a->d_workspace, nblk, a->npix, a->d_out);
    DS_CHECK_LAUNCH("psnr_final");
    return DS_OK;
}
