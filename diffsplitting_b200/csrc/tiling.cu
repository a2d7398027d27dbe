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
