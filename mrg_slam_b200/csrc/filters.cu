// Prefiltering kernels: distance filter, VoxelGrid, radius / statistical outlier removal.
//
// Replaces: A19 distance_filter (apps/prefiltering_component.cpp:206-229, in-tree),
// A16 pcl::VoxelGrid<PointXYZI>::filter (:167-171; apps/scan_matching_odometry_component.cpp:175-179),
// A17 pcl::RadiusOutlierRemoval (:195-199), A18 pcl::StatisticalOutlierRemoval (:190-194).
#include <cfloat>
#include <climits>
#include <cmath>
#include <algorithm>

#include "internal.hpp"
#include "knn.cuh"

namespace b2r {

void compute_bbox(Ctx& ctx, const float4* pts, int n, float mn[3], float mx[3]);  // cloud.cu
void compute_bbox_range(Ctx& ctx, const float4* pts, int n, double near_t, double far_t, float mn[3], float mx[3], int* count);  // cloud.cu

// ------------------------------------------------------------------------------------------------ distance filter
// keep iff near < |p| < far, norm in float as x^2 + (y^2 + z^2) (Eigen 3-vector reduction order), widened to double
__global__ void distance_flag_kernel(const float4* __restrict__ in, int n, double near_t, double far_t, uint8_t* __restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  const float s = __fadd_rn(__fmul_rn(p.x, p.x), __fadd_rn(__fmul_rn(p.y, p.y), __fmul_rn(p.z, p.z)));
  const double d = (double)__fsqrt_rn(s);
  keep[i] = (d > near_t && d < far_t) ? 1 : 0;
}

void filter_distance(Ctx& ctx, const float4* in, int n, double near_t, double far_t, DevCloud& out) {
  out.n = 0;
  if (n == 0) return;
  DBuf<uint8_t> keep; keep.alloc(n, ctx.stream);
  B2R_LAUNCH(ctx, distance_flag_kernel, (n + 255) / 256, 256, 0, in, n, near_t, far_t, keep.p);
  compact_points(ctx, in, keep.p, n, out);
}

// ------------------------------------------------------------------------------------------------ VoxelGrid
struct VgParams {
  float inv_leaf;
  int min_b[3];
  int mul[3];
  int div[3];     // voxels per axis
  int shift, cx;  // coarse cells of the sort: runs of 2^shift x-voxels, cx of them per (y, z) row
};

// key = dense voxel index (pcl::VoxelGrid: float math, floor(p*inv_leaf) - min_b), value = point index
// RANGE: the distance filter (near < |p| < far, distance_flag_kernel's arithmetic) is applied here instead of compacting first
template <bool RANGE>
__global__ void vg_key_kernel(const float4* __restrict__ in, int n, VgParams prm, double near_t, double far_t, unsigned* __restrict__ keys,
                              int* __restrict__ coarse, int* __restrict__ cell_cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  unsigned key = 0xffffffffu;  // non-finite (and out-of-range) points sort last and are dropped
  bool ok = isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
  if (RANGE) {
    const float s = __fadd_rn(__fmul_rn(p.x, p.x), __fadd_rn(__fmul_rn(p.y, p.y), __fmul_rn(p.z, p.z)));
    const double d = (double)__fsqrt_rn(s);
    ok = d > near_t && d < far_t;
  }
  int cc = 0;
  if (ok) {
    const int i0 = (int)(floorf(__fmul_rn(p.x, prm.inv_leaf)) - (float)prm.min_b[0]);
    const int i1 = (int)(floorf(__fmul_rn(p.y, prm.inv_leaf)) - (float)prm.min_b[1]);
    const int i2 = (int)(floorf(__fmul_rn(p.z, prm.inv_leaf)) - (float)prm.min_b[2]);
    key = (unsigned)(i0 * prm.mul[0] + i1 * prm.mul[1] + i2 * prm.mul[2]);
    cc = (i2 * prm.div[1] + i1) * prm.cx + (i0 >> prm.shift);
    atomicAdd(&cell_cnt[cc], 1);  // the sort's counting pass (VoxelSort)
  }
  keys[i] = key;
  coarse[i] = cc;
}
// sorted keys [0, *n_valid): 1 where a new voxel starts; 0 beyond the valid entries
template <typename KeyT>
__global__ void vg_head_kernel(const KeyT* __restrict__ keys, int n, const int* __restrict__ n_valid, uint8_t* __restrict__ head) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i < n_valid[0] && (i == 0 || keys[i - 1] != keys[i])) ? 1 : 0;
}
// counts per block of 256 -> (after scan) offsets; then each head writes its position
__global__ void vg_seg_kernel(const uint8_t* __restrict__ head, int n, const int* __restrict__ block_off, int* __restrict__ seg_start) {
  __shared__ int wsum[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = (i < n && head[i]) ? 1 : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, v);
  if (lane == 0) wsum[warp] = __popc(bal);
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += wsum[w];
  if (v) seg_start[block_off[blockIdx.x] + woff + __popc(bal & ((1u << lane) - 1u))] = i;
}
// one WARP per voxel: pcl::CentroidPoint<PointXYZI> — float sums in (stable-sorted = ascending point index) order, divided by
// float(count).  The order of the additions is fixed, so they stay sequential, but the points of a voxel are gathered 32 at a
// time by the lanes (the loads are the expensive part: a thread per voxel spent its time waiting for them one by one, and a
// voxel near the sensor holds hundreds of points) and handed to the running sums through shuffles.
__global__ void __launch_bounds__(256) vg_centroid_kernel(const float4* __restrict__ in, const int* __restrict__ vals, const int* __restrict__ seg_start,
                                                          int nseg, int n_valid, int min_pts, float4* __restrict__ out, uint8_t* __restrict__ keep) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= nseg) return;
  const int b = seg_start[s], e = (s + 1 < nseg) ? seg_start[s + 1] : n_valid;
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  for (int j0 = b; j0 < e; j0 += 32) {
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j0 + lane < e) p = __ldg(&in[__ldg(&vals[j0 + lane])]);
    const int m = min(32, e - j0);
    for (int l = 0; l < m; ++l) {
      sx = __fadd_rn(sx, __shfl_sync(0xffffffffu, p.x, l)); sy = __fadd_rn(sy, __shfl_sync(0xffffffffu, p.y, l));
      sz = __fadd_rn(sz, __shfl_sync(0xffffffffu, p.z, l)); si = __fadd_rn(si, __shfl_sync(0xffffffffu, p.w, l));
    }
  }
  if (lane != 0) return;
  const float cnt = (float)(e - b);
  out[s] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), __fdiv_rn(si, cnt));
  if (keep) keep[s] = (e - b) >= min_pts ? 1 : 0;
}

void flags_block_offsets(Ctx& ctx, const uint8_t* flags, int n, int* block_off);  // cloud.cu

// ------------------------------------------------------------------------------------------------ voxel sort
// Stable sort of (voxel key, point index) pairs by key — the std::sort of pcl::VoxelGrid::applyFilter and the
// accumulation order of ApproximateMeanVoxelGrid's hash map — WITHOUT a general radix sort.  The key is a dense voxel index
// x + DX * (y + DY * z), so ascending key = ascending (z, y, x): the points are counting-sorted in ONE pass into coarse cells
// that are runs of 2^shift consecutive x-voxels of one (y, z) row (coarse cell order = key order), and every entry then finds
// its rank inside its coarse cell by (key, point index).  Whatever the number of key bits this is one pass over the data:
//   vs_count     cell counts (atomics)                     vs_scan      exclusive scan (tile scans; the last block to finish
//   vs_scatter   entries into their cell (atomic cursor)                scans the tile totals: one launch)
//   vs_rank      in-cell rank -> final position, keys and point indices written sorted
// (replaces cub::DeviceRadixSort::SortPairs, which needs 4 digit passes for a 32-bit key and 8 for a 64-bit one).
// shift is chosen so that the table has at most max(4 M, 32 n) cells: a coarse cell holds a handful of entries, and the quadratic
// in-cell ranking stays cheap.  A cell with more than kVsHeavy entries (a leaf size of many metres: thousands of points per
// voxel) is not ranked; the call then falls back to a plain bitonic sort of the (key, index) pairs (vs_bitonic_*): slow,
// O(n log^2 n), but only reached by inputs no mrg_slam configuration produces.
constexpr int kVsTile = 2048;   // cells per scan tile (256 threads x 8)
constexpr int kVsHeavy = 8192;  // entries per coarse cell beyond which the in-cell ranking is refused

// start[c] = exclusive scan of cnt inside its tile; tile_off[t] = exclusive scan of the tile totals (written by the last block);
// cnt is reset to 0 (it becomes the scatter cursor); total[0] = number of valid entries
__global__ void __launch_bounds__(256) vs_scan_kernel(int* __restrict__ cnt, int ncell, int* __restrict__ start, int* __restrict__ tile_tot,
                                                      int* __restrict__ tile_off, int ntiles, unsigned* __restrict__ blocks_done, int* __restrict__ total) {
  __shared__ int wsum[8];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int base = blockIdx.x * kVsTile + threadIdx.x * 8;  // the table is padded to whole tiles (zero counts beyond ncell)
  int v[8], run = 0;
  {
    const int4 a = *reinterpret_cast<const int4*>(cnt + base), b = *reinterpret_cast<const int4*>(cnt + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
    for (int k = 0; k < 8; ++k) run += v[k];
  }
  int incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { woff += w < warp ? wsum[w] : 0; tot += wsum[w]; }
  int ex = woff + incl - run;
  int o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { o[k] = ex; ex += v[k]; }
  *reinterpret_cast<int4*>(start + base) = make_int4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<int4*>(start + base + 4) = make_int4(o[4], o[5], o[6], o[7]);
  *reinterpret_cast<int4*>(cnt + base) = make_int4(0, 0, 0, 0);
  *reinterpret_cast<int4*>(cnt + base + 4) = make_int4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tile_tot[blockIdx.x] = tot;
    __threadfence();
    last = atomicAdd(blocks_done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block: exclusive scan of the tile totals (sequential chunks of 256)
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b = 0; b < ntiles; b += 256) {
    const int i = b + threadIdx.x;
    const int x = i < ntiles ? ((volatile int*)tile_tot)[i] : 0;
    int in2 = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, in2, o);
      if (lane >= o) in2 += t;
    }
    if (lane == 31) wsum[warp] = in2;
    __syncthreads();
    int wo = 0, tt = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { wo += w < warp ? wsum[w] : 0; tt += wsum[w]; }
    if (i < ntiles) tile_off[i] = carry + wo + in2 - x;
    __syncthreads();
    if (threadIdx.x == 0) carry += tt;
    __syncthreads();
  }
  if (threadIdx.x == 0) { total[0] = carry; *blocks_done = 0u; }
}
template <typename KeyT>
__global__ void vs_scatter_kernel(const KeyT* __restrict__ keys, const int* __restrict__ coarse, int n, const int* __restrict__ start,
                                  const int* __restrict__ tile_off, int* __restrict__ cur, int* __restrict__ tmp_idx, KeyT* __restrict__ tmp_key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const KeyT k = keys[i];
  if (k == (KeyT)~(KeyT)0) return;
  const int c = coarse[i];
  const int pos = start[c] + tile_off[c / kVsTile] + atomicAdd(&cur[c], 1);
  tmp_idx[pos] = i;
  tmp_key[pos] = k;
}
// one thread per entry (scatter order inside a cell is arbitrary): rank by (key, point index) among the cell's entries
template <typename KeyT>
__global__ void vs_rank_kernel(const KeyT* __restrict__ tmp_key, const int* __restrict__ coarse, const int* __restrict__ tmp_idx,
                               const int* __restrict__ total, const int* __restrict__ start, const int* __restrict__ tile_off,
                               const int* __restrict__ cur /* = cell counts after the scatter */, KeyT* __restrict__ keys_out,
                               int* __restrict__ vals_out, int* __restrict__ heavy_flag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total[0]) return;
  const int i = tmp_idx[j];
  const KeyT k = tmp_key[j];
  const int c = coarse[i];
  const int s = start[c] + tile_off[c / kVsTile], e = s + cur[c];
  if (e - s > kVsHeavy) { *heavy_flag = 1; return; }
  int rank = 0;
#pragma unroll 8
  for (int t = s; t < e; ++t) {  // the cell's entries: read by all of its threads (L1 broadcast)
    const int it = __ldg(&tmp_idx[t]);
    const KeyT kt = __ldg(&tmp_key[t]);
    rank += (kt < k || (kt == k && it < i)) ? 1 : 0;
  }
  keys_out[s + rank] = k;
  vals_out[s + rank] = i;
}

// keys[i]: dense voxel index of point i or all-ones (dropped); coarse[i]: its coarse cell (any value for dropped points).
// On return keys_out / vals_out[0 .. *d_total) hold the valid pairs in ascending (key, point index) order.
// ---- fallback: bitonic sort of (key, point index) pairs, padded to a power of two with all-ones keys (which sort last)
template <typename KeyT>
__global__ void vs_bitonic_init_kernel(const KeyT* __restrict__ keys, int n, int npad, KeyT* __restrict__ k, int* __restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  k[i] = i < n ? keys[i] : (KeyT)~(KeyT)0;
  v[i] = i;
}
template <typename KeyT>
__global__ void vs_bitonic_step_kernel(KeyT* __restrict__ k, int* __restrict__ v, int npad, int jj, int kk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = i ^ jj;
  if (i >= npad || p <= i) return;
  const KeyT ka = k[i], kb = k[p];
  const int va = v[i], vb = v[p];
  const bool gt = ka > kb || (ka == kb && va > vb);
  if (((i & kk) == 0) == gt) { k[i] = kb; k[p] = ka; v[i] = vb; v[p] = va; }
}
template <typename KeyT>
__global__ void vs_count_valid_kernel(const KeyT* __restrict__ k, int n, int* __restrict__ total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool valid = k[i] != (KeyT)~(KeyT)0;
  const bool next_valid = (i + 1 < n) ? k[i + 1] != (KeyT)~(KeyT)0 : false;
  if (valid && !next_valid) total[0] = i + 1;
  if (i == 0 && !valid) total[0] = 0;
}
template <typename KeyT>
static void voxel_sort_fallback(Ctx& ctx, const KeyT* keys, int n, KeyT* keys_out, int* vals_out, int* d_total) {
  int npad = 1;
  while (npad < n) npad <<= 1;
  DBuf<KeyT> k; k.alloc(npad, ctx.stream);
  DBuf<int> v; v.alloc(npad, ctx.stream);
  const int nb = (npad + 255) / 256;
  B2R_LAUNCH(ctx, vs_bitonic_init_kernel<KeyT>, nb, 256, 0, keys, n, npad, k.p, v.p);
  for (int kk = 2; kk <= npad; kk <<= 1)
    for (int jj = kk >> 1; jj > 0; jj >>= 1) B2R_LAUNCH(ctx, vs_bitonic_step_kernel<KeyT>, nb, 256, 0, k.p, v.p, npad, jj, kk);
  B2R_CUDA(cudaMemcpyAsync(keys_out, k.p, sizeof(KeyT) * n, cudaMemcpyDeviceToDevice, ctx.stream));
  B2R_CUDA(cudaMemcpyAsync(vals_out, v.p, sizeof(int) * n, cudaMemcpyDeviceToDevice, ctx.stream));
  B2R_LAUNCH(ctx, vs_count_valid_kernel<KeyT>, (n + 255) / 256, 256, 0, keys_out, n, d_total);
}

// The sort in two calls around the caller's key kernel, which counts the cells itself (cnt[coarse] += 1 per valid point):
//   VoxelSort vs; vs.begin(ctx, n, ncoarse);  <key kernel writing keys / coarse and counting into vs.cnt.p>;  vs.finish(...)
// d_heavy (device int, zeroed by begin): set when a coarse cell was too crowded to rank — the caller reads it at its next
// synchronisation and, if set, calls voxel_sort_fallback and repeats what it derived from the sorted arrays.
struct VoxelSort {
  DBuf<int> cnt, start, tile, tmp;
  int n = 0, ncoarse = 0, ntiles = 0;
  void begin(Ctx& ctx, int n_, int ncoarse_, int* d_heavy) {
    n = n_; ncoarse = ncoarse_;
    ntiles = (ncoarse + kVsTile - 1) / kVsTile;
    cnt.alloc((size_t)ntiles * kVsTile, ctx.stream);
    start.alloc((size_t)ntiles * kVsTile, ctx.stream);
    tile.alloc((size_t)2 * ntiles + 1, ctx.stream);
    tmp.alloc((size_t)n, ctx.stream);
    cnt.zero(ctx.stream);
    B2R_CUDA(cudaMemsetAsync(tile.p + 2 * ntiles, 0, sizeof(unsigned), ctx.stream));
    B2R_CUDA(cudaMemsetAsync(d_heavy, 0, sizeof(int), ctx.stream));
  }
  template <typename KeyT>
  void finish(Ctx& ctx, const KeyT* keys, const int* coarse, KeyT* keys_out, int* vals_out, int* d_total, int* d_heavy) {
    DBuf<KeyT> tmpk;
    tmpk.alloc((size_t)n, ctx.stream);
    unsigned* done = reinterpret_cast<unsigned*>(tile.p + 2 * ntiles);
    const int nb = (n + 255) / 256;
    B2R_LAUNCH(ctx, vs_scan_kernel, ntiles, 256, 0, cnt.p, ncoarse, start.p, tile.p, tile.p + ntiles, ntiles, done, d_total);
    B2R_LAUNCH(ctx, vs_scatter_kernel<KeyT>, nb, 256, 0, keys, coarse, n, start.p, tile.p + ntiles, cnt.p, tmp.p, tmpk.p);
    B2R_LAUNCH(ctx, vs_rank_kernel<KeyT>, nb, 256, 0, tmpk.p, coarse, tmp.p, d_total, start.p, tile.p + ntiles, cnt.p, keys_out, vals_out, d_heavy);
  }
};
// x-voxels per coarse cell (a power of two) and the table size for a DX x DY x DZ voxel grid holding n points
static void voxel_sort_dims(long long DX, long long DY, long long DZ, int n, int& shift, long long& cx, long long& ncoarse) {
  const long long want = std::max<long long>(1 << 22, 32ll * n);
  shift = 0;
  for (;;) {
    cx = ((DX - 1) >> shift) + 1;
    ncoarse = cx * DY * DZ;
    if (ncoarse <= want || cx == 1) break;
    ++shift;
  }
  if (ncoarse > (1ll << 27)) throw Error(B2R_ERR_CAPACITY, "voxel grid has too many (y, z) rows for the sort table");
}

// range != nullptr: {near, far} of the distance filter that precedes VoxelGrid in the prefilter chain
// (apps/prefiltering_component.cpp:149-151), folded into the bounding-box and key passes; the result is the one of
// filter_distance followed by filter_voxelgrid (same points, same relative order => same float sums).
void filter_voxelgrid(Ctx& ctx, const float4* in, int n, float leaf, int min_pts, DevCloud& out, bool& overflow, const double* range) {
  overflow = false;
  out.n = 0;
  if (n == 0) return;
  float mn[3], mx[3];
  if (range) {
    int kept = 0;
    compute_bbox_range(ctx, in, n, range[0], range[1], mn, mx, &kept);
    if (kept == 0) return;
  } else {
    compute_bbox(ctx, in, n, mn, mx);
  }
  const float inv_leaf = 1.0f / leaf;
  const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv_leaf) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv_leaf) + 1,
                dz = (int64_t)((mx[2] - mn[2]) * inv_leaf) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) {  // PCL: "Leaf size is too small ... Integer indices would overflow": output = input
    overflow = true;
    if (range) {  // ... the input of VoxelGrid being the distance-filtered cloud
      filter_distance(ctx, in, n, range[0], range[1], out);
      return;
    }
    out.pts.alloc(n, ctx.stream);
    B2R_CUDA(cudaMemcpyAsync(out.pts.p, in, (size_t)n * 16, cudaMemcpyDeviceToDevice, ctx.stream));
    out.n = n;
    return;
  }
  // the centroids lie inside the box of the points they average (up to an ulp of float rounding: padded)
  out.has_box = true;
  for (int d = 0; d < 3; ++d) {
    out.box_min[d] = mn[d] - 1e-5f * std::max(1.f, std::fabs(mn[d]));
    out.box_max[d] = mx[d] + 1e-5f * std::max(1.f, std::fabs(mx[d]));
  }
  VgParams prm;
  prm.inv_leaf = inv_leaf;
  int div_b[3];
  for (int d = 0; d < 3; ++d) {
    prm.min_b[d] = (int)std::floor(mn[d] * inv_leaf);
    div_b[d] = (int)std::floor(mx[d] * inv_leaf) - prm.min_b[d] + 1;
  }
  prm.mul[0] = 1; prm.mul[1] = div_b[0]; prm.mul[2] = div_b[0] * div_b[1];
  for (int d = 0; d < 3; ++d) prm.div[d] = div_b[d];
  long long cx = 1, ncoarse = 1;
  voxel_sort_dims(div_b[0], div_b[1], div_b[2], n, prm.shift, cx, ncoarse);
  prm.cx = (int)cx;

  DBuf<unsigned> k0, k1;
  DBuf<int> c0, v1;
  k0.alloc(n, ctx.stream); k1.alloc(n, ctx.stream); c0.alloc(n, ctx.stream); v1.alloc(n, ctx.stream);
  const int nb = (n + 255) / 256;
  DBuf<int> cnt; cnt.alloc((size_t)nb + 3, ctx.stream);  // block offsets | number of voxels | valid points | crowded-cell flag
  // ascending voxel index, ties in ascending point index (what a stable sort gives): one-pass counting sort + in-cell ranks
  VoxelSort vs;
  vs.begin(ctx, n, (int)ncoarse, cnt.p + nb + 2);
  if (range) B2R_LAUNCH(ctx, vg_key_kernel<true>, nb, 256, 0, in, n, prm, range[0], range[1], k0.p, c0.p, vs.cnt.p);
  else B2R_LAUNCH(ctx, vg_key_kernel<false>, nb, 256, 0, in, n, prm, 0.0, 0.0, k0.p, c0.p, vs.cnt.p);
  vs.finish<unsigned>(ctx, k0.p, c0.p, k1.p, v1.p, cnt.p + nb + 1, cnt.p + nb + 2);
  DBuf<uint8_t> head; head.alloc(n, ctx.stream);
  int h2[3] = {0, 0, 0};
  for (int attempt = 0; attempt < 2; ++attempt) {
    B2R_LAUNCH(ctx, vg_head_kernel<unsigned>, nb, 256, 0, k1.p, n, cnt.p + nb + 1, head.p);
    flags_block_offsets(ctx, head.p, n, cnt.p);
    B2R_CUDA(cudaMemcpyAsync(h2, cnt.p + nb, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(ctx.stream));
    if (!h2[2]) break;
    voxel_sort_fallback<unsigned>(ctx, k0.p, n, k1.p, v1.p, cnt.p + nb + 1);  // a crowded cell was left unranked
    B2R_CUDA(cudaMemsetAsync(cnt.p + nb + 2, 0, sizeof(int), ctx.stream));
  }
  const int nseg = h2[0], n_valid = h2[1];
  if (nseg == 0) return;
  DBuf<int> seg; seg.alloc(nseg, ctx.stream);
  B2R_LAUNCH(ctx, vg_seg_kernel, nb, 256, 0, head.p, n, cnt.p, seg.p);
  if (min_pts <= 1) {
    out.pts.alloc(nseg, ctx.stream);
    B2R_LAUNCH(ctx, vg_centroid_kernel, (nseg + 7) / 8, 256, 0, in, v1.p, seg.p, nseg, n_valid, min_pts, out.pts.p, (uint8_t*)nullptr);
    out.n = nseg;
  } else {
    DBuf<float4> cen; cen.alloc(nseg, ctx.stream);
    DBuf<uint8_t> keep; keep.alloc(nseg, ctx.stream);
    B2R_LAUNCH(ctx, vg_centroid_kernel, (nseg + 7) / 8, 256, 0, in, v1.p, seg.p, nseg, n_valid, min_pts, cen.p, keep.p);
    compact_points(ctx, cen.p, keep.p, nseg, out);
  }
}

// ------------------------------------------------------------------------------------------------ radius outlier removal
// one thread per (cell-sorted) point: count neighbours with d2 < r2 (incl. itself), keep iff count > min_neighbors
__global__ void radius_keep_kernel(const CloudView* __restrict__ views, float r2, int min_nb, uint8_t* __restrict__ keep) {
  const CloudView& c = views[0];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c.n) return;
  const float4 p = __ldg(&c.spts[j]);
  int k;
  if (min_nb == 1) {
    // PCL special case: nearestKSearch(2) and d2[1] <= r^2.  Equivalent count with a non-strict bound.
    k = radius_count(c, p.x, p.y, p.z, nextafterf(r2, INFINITY), 1);
  } else {
    k = radius_count(c, p.x, p.y, p.z, r2, min_nb);
  }
  keep[__float_as_int(p.w)] = k > min_nb ? 1 : 0;
}

// a cloud object over the caller's device points (borrowed: the filters only need it for the duration of the call)
static void with_temp_cloud(Ctx& ctx, const float4* in, int n, Cloud& tmp, const DevCloud* box = nullptr) {
  tmp.device = ctx.device;
  tmp.n = n;
  tmp.pts.p = const_cast<float4*>(in);
  if (box && box->has_box) {  // any box containing the points serves the exact search grid
    for (int d = 0; d < 3; ++d) { tmp.bmin[d] = box->box_min[d]; tmp.bmax[d] = box->box_max[d]; }
    tmp.has_bbox = true;
  }
}

void filter_radius(Ctx& ctx, const b2r_config& cfg, const float4* in, int n, double radius, int min_nb, DevCloud& out, const DevCloud* box) {
  out.n = 0;
  if (n == 0) return;
  Cloud tmp;
  with_temp_cloud(ctx, in, n, tmp, box);
  b2r_config c2 = cfg;
  c2.nn_cell_size = radius;  // 3x3x3 cells of size >= r cover the search ball
  std::vector<Cloud*> cl{&tmp};
  std::vector<Needs> nd(1);
  nd[0].grid = true;
  DBuf<CloudView> dv;
  clouds_prepare(ctx, c2, cl, nd, dv);
  DBuf<uint8_t> keep; keep.alloc(n, ctx.stream);
  const float r2 = (float)(radius * radius);
  B2R_LAUNCH(ctx, radius_keep_kernel, (n + 127) / 128, 128, 0, dv.p, r2, min_nb, keep.p);
  compact_points(ctx, in, keep.p, n, out);
}

// ------------------------------------------------------------------------------------------------ statistical outlier removal
// one thread per (cell-sorted) point: (mean_k+1)-NN; distances[i] = (float)(sum_{j=1..k} sqrt(d2_j) / k), double sum in
// ascending-distance order with float sqrt, as PCL does
template <int K>
__global__ void __launch_bounds__(128) sor_distance_kernel(const CloudView* __restrict__ views, int mean_k, float* __restrict__ distances) {
  const CloudView& c = views[0];
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= c.n) return;
  const float4 p = __ldg(&c.spts[qi]);
  TopkVisitor<K> v(p.x, p.y, p.z);
  knn_topk<K>(c, query_cell(c, p.x, p.y, p.z), mean_k + 1, v);
  double sum = 0.0;
#pragma unroll
  for (int j = 1; j < K; ++j)
    if (j <= mean_k) sum += (double)__fsqrt_rn(v.d[j]);
  distances[__float_as_int(p.w)] = (float)(sum / (double)mean_k);
}
// two-stage deterministic reduction of sum and sum of (float) squares
__global__ void __launch_bounds__(256) sor_stats_kernel(const float* __restrict__ distances, int n, double* __restrict__ partials) {
  __shared__ double red[2 * 8];
  double acc[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float d = distances[i];
    acc[0] += (double)d;
    acc[1] += (double)__fmul_rn(d, d);
  }
  block_reduce_to<2>(acc, red, partials + blockIdx.x * 2);
}
__global__ void sor_threshold_kernel(const double* __restrict__ partials, int nblocks, int n, double stddev_mul, double* __restrict__ thr_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sum = 0.0, sq = 0.0;
  for (int b = 0; b < nblocks; ++b) { sum += partials[b * 2]; sq += partials[b * 2 + 1]; }
  const double nn = (double)n;
  const double mean = sum / nn;
  const double variance = (sq - sum * sum / nn) / (nn - 1.0);
  *thr_out = mean + stddev_mul * sqrt(variance);
}
__global__ void sor_keep_kernel(const float* __restrict__ distances, int n, const double* __restrict__ thr, uint8_t* __restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keep[i] = ((double)distances[i] > *thr) ? 0 : 1;
}

void filter_statistical(Ctx& ctx, const b2r_config& cfg, const float4* in, int n, int mean_k, double stddev_mul, DevCloud& out,
                        const DevCloud* box) {
  out.n = 0;
  if (n == 0) return;
  if (mean_k < 1 || mean_k > 31) throw Error(B2R_ERR_INVALID_ARG, "statistical_mean_k must be in [1,31]");
  if (n < mean_k + 1) throw Error(B2R_ERR_INVALID_ARG, "cloud has fewer than mean_k+1 points");
  Cloud tmp;
  with_temp_cloud(ctx, in, n, tmp, box);
  std::vector<Cloud*> cl{&tmp};
  std::vector<Needs> nd(1);
  nd[0].grid = true;
  DBuf<CloudView> dv;
  clouds_prepare(ctx, cfg, cl, nd, dv);
  DBuf<float> dist; dist.alloc(n, ctx.stream);
  {
    const int kk = mean_k + 1, nb128 = (n + 127) / 128;
    if (kk <= 8) B2R_LAUNCH(ctx, sor_distance_kernel<8>, nb128, 128, 0, dv.p, mean_k, dist.p);
    else if (kk <= 16) B2R_LAUNCH(ctx, sor_distance_kernel<16>, nb128, 128, 0, dv.p, mean_k, dist.p);
    else if (kk <= 24) B2R_LAUNCH(ctx, sor_distance_kernel<24>, nb128, 128, 0, dv.p, mean_k, dist.p);
    else B2R_LAUNCH(ctx, sor_distance_kernel<32>, nb128, 128, 0, dv.p, mean_k, dist.p);
  }
  const int nb = std::max(1, std::min(2 * ctx.num_sms, (n + 1023) / 1024));
  DBuf<double> part; part.alloc((size_t)nb * 2 + 1, ctx.stream);
  B2R_LAUNCH(ctx, sor_stats_kernel, nb, 256, 0, dist.p, n, part.p);
  B2R_LAUNCH(ctx, sor_threshold_kernel, 1, 32, 0, part.p, nb, n, stddev_mul, part.p + (size_t)nb * 2);
  DBuf<uint8_t> keep; keep.alloc(n, ctx.stream);
  B2R_LAUNCH(ctx, sor_keep_kernel, (n + 255) / 256, 256, 0, dist.p, n, part.p + (size_t)nb * 2, keep.p);
  compact_points(ctx, in, keep.p, n, out);
}

}  // namespace b2r

// ------------------------------------------------------------------------------------------------ map cloud
// MapCloudGenerator::generate (src/mrg_slam/map_cloud_generator.cpp:14-86) + pcl::ApproximateMeanVoxelGrid<PointXYZI>
// (include/pcl/filters/ApproximateMeanVoxelGrid.hpp:63-126), both in the reference tree.
namespace b2r {

struct MapSegment {  // one keyframe inside the concatenated input
  int begin;         // first point of the keyframe in the concatenation
  int n;
  float pose[16];    // keyframe->pose.matrix().cast<float>(), column-major
};

// One thread per input point: far-distance test in the sensor frame (float Vector3f::squaredNorm, x^2 + (y^2 + z^2) as
// Eigen's unrolled 3-element reduction associates it — the same order as the distance filter above; :40-42),
// then dst = pose * (x, y, z, 1) as Eigen's fixed-size product evaluates it: ((c0 x + c1 y) + c2 z) + c3 (:44).
__global__ void map_transform_kernel(const float4* __restrict__ in, int n, const MapSegment* __restrict__ segs, int nseg, int use_far,
                                     float far_sq, int drop_nonfinite, float4* __restrict__ out, uint8_t* __restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = nseg - 1;  // segment of point i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].begin <= i) lo = mid; else hi = mid - 1;
  }
  const float* T = segs[lo].pose;
  const float4 p = in[i];
  bool k = true;
  if (use_far) {
    const float sq = __fadd_rn(__fmul_rn(p.x, p.x), __fadd_rn(__fmul_rn(p.y, p.y), __fmul_rn(p.z, p.z)));
    if (sq > far_sq) k = false;
  }
  float4 o;
  o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], p.x), __fmul_rn(T[4], p.y)), __fmul_rn(T[8], p.z)), T[12]);
  o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[1], p.x), __fmul_rn(T[5], p.y)), __fmul_rn(T[9], p.z)), T[13]);
  o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[2], p.x), __fmul_rn(T[6], p.y)), __fmul_rn(T[10], p.z)), T[14]);
  o.w = p.w;
  // the voxel filter casts floor(p * inv_leaf) to int, undefined for non-finite points upstream: they are dropped here
  if (drop_nonfinite && !(isfinite(o.x) && isfinite(o.y) && isfinite(o.z))) k = false;
  out[i] = o;
  keep[i] = k ? 1 : 0;
}

struct AmvgParams {
  float inv_leaf;
  int min_b[3];
  long long mul[3];
  long long cx;  // coarse cells of the sort per (y, z) row
  long long ext1;
  int shift;
};
// key = (ix - min) + nx * ((iy - min) + ny * (iz - min)), ixyz = (int)floor(p * inv_leaf) (hpp:88-90): one key per
// ApproximateMeanVoxelGrid hash-map entry; sorting by it only fixes the (implementation-defined) output order
__global__ void amvg_key_kernel(const float4* __restrict__ in, int n, AmvgParams prm, unsigned long long* __restrict__ keys, int* __restrict__ coarse,
                                int* __restrict__ cell_cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  const long long i0 = (long long)(int)floorf(__fmul_rn(p.x, prm.inv_leaf)) - prm.min_b[0];
  const long long i1 = (long long)(int)floorf(__fmul_rn(p.y, prm.inv_leaf)) - prm.min_b[1];
  const long long i2 = (long long)(int)floorf(__fmul_rn(p.z, prm.inv_leaf)) - prm.min_b[2];
  keys[i] = (unsigned long long)(i0 * prm.mul[0] + i1 * prm.mul[1] + i2 * prm.mul[2]);
  const int cc = (int)((i2 * prm.ext1 + i1) * prm.cx + (i0 >> prm.shift));
  coarse[i] = cc;
  atomicAdd(&cell_cnt[cc], 1);  // the sort's counting pass (VoxelSort)
}

void map_cloud(Ctx& ctx, const void* const* clouds, const size_t* n, const double* poses_colmajor, const uint8_t* first_keyframe, size_t count,
               size_t stride_bytes, int memspace, float resolution, int min_points_per_voxel, float distance_far_thresh, int skip_first_cloud,
               DevCloud& out, bool& null_result) {
  null_result = false;
  out.n = 0;
  if (count == 0) { null_result = true; return; }  // :20-23
  // ---- concatenate the keyframes that take part (:31-35) and describe them for the transform kernel
  std::vector<MapSegment> segs;
  size_t total = 0;
  for (size_t k = 0; k < count; ++k) {
    if (first_keyframe && first_keyframe[k] && skip_first_cloud) continue;
    if (n[k] == 0) continue;
    MapSegment s;
    s.begin = (int)total;
    s.n = (int)n[k];
    for (int i = 0; i < 16; ++i) s.pose[i] = (float)poses_colmajor[k * 16 + i];
    segs.push_back(s);
    total += n[k];
    if (total > (size_t)INT32_MAX) throw Error(B2R_ERR_CAPACITY, "map cloud: more than 2^31 input points");
  }
  if (total == 0) { null_result = count > 1; return; }  // :58-61
  DBuf<float4> cat; cat.alloc(total, ctx.stream);
  {
    size_t si = 0;
    for (size_t k = 0; k < count; ++k) {
      if (first_keyframe && first_keyframe[k] && skip_first_cloud) continue;
      if (n[k] == 0) continue;
      float4* dst = cat.p + segs[si].begin;
      if (stride_bytes == 16) {
        B2R_CUDA(cudaMemcpyAsync(dst, clouds[k], n[k] * 16, memspace == B2R_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx.stream));
      } else {
        DBuf<float4> tmp;
        load_points(ctx, clouds[k], n[k], stride_bytes, memspace, tmp);
        B2R_CUDA(cudaMemcpyAsync(dst, tmp.p, n[k] * 16, cudaMemcpyDeviceToDevice, ctx.stream));
      }
      ++si;
    }
  }
  DBuf<MapSegment> dsegs; dsegs.alloc(segs.size(), ctx.stream);
  B2R_CUDA(cudaMemcpyAsync(dsegs.p, segs.data(), sizeof(MapSegment) * segs.size(), cudaMemcpyHostToDevice, ctx.stream));
  const int N = (int)total;
  DBuf<float4> tr; tr.alloc(total, ctx.stream);
  DBuf<uint8_t> keep; keep.alloc(total, ctx.stream);
  const bool filter = resolution > 0.0f;
  B2R_LAUNCH(ctx, map_transform_kernel, (N + 255) / 256, 256, 0, cat.p, N, dsegs.p, (int)segs.size(), distance_far_thresh > 0 ? 1 : 0,
             distance_far_thresh * distance_far_thresh, filter ? 1 : 0, tr.p, keep.p);
  DevCloud world;
  compact_points(ctx, tr.p, keep.p, N, world);  // order-preserving: keyframe order, then point order, like push_back
  if (world.n == 0) { null_result = count > 1; return; }
  if (!filter) { out = std::move(world); return; }  // :67-71 full-resolution cloud

  // ---- ApproximateMeanVoxelGrid: per voxel float sums of x, y, z, intensity in input order, / float(count)
  const int M = world.n;
  float mn[3], mx[3];
  compute_bbox(ctx, world.pts.p, M, mn, mx);
  AmvgParams prm;
  prm.inv_leaf = 1.0f / resolution;  // Array3f::Ones() / leaf_size_.array()
  long long ext[3];
  for (int d = 0; d < 3; ++d) {
    prm.min_b[d] = (int)std::floor(mn[d] * prm.inv_leaf);
    ext[d] = (long long)(int)std::floor(mx[d] * prm.inv_leaf) - prm.min_b[d] + 1;
  }
  if ((double)ext[0] * (double)ext[1] * (double)ext[2] > 9.0e18) throw Error(B2R_ERR_CAPACITY, "map cloud: voxel index range exceeds 63 bits");
  prm.mul[0] = 1; prm.mul[1] = ext[0]; prm.mul[2] = ext[0] * ext[1];
  long long ncoarse = 1;
  prm.ext1 = ext[1];
  voxel_sort_dims(ext[0], ext[1], ext[2], M, prm.shift, prm.cx, ncoarse);
  DBuf<unsigned long long> k0, k1;
  DBuf<int> c0, v1;
  k0.alloc(M, ctx.stream); k1.alloc(M, ctx.stream); c0.alloc(M, ctx.stream); v1.alloc(M, ctx.stream);
  const int nb = (M + 255) / 256;
  DBuf<int> cnt; cnt.alloc((size_t)nb + 3, ctx.stream);
  // equal keys keep ascending input order = the hash map's accumulation order
  VoxelSort vs;
  vs.begin(ctx, M, (int)ncoarse, cnt.p + nb + 2);
  B2R_LAUNCH(ctx, amvg_key_kernel, nb, 256, 0, world.pts.p, M, prm, k0.p, c0.p, vs.cnt.p);
  vs.finish<unsigned long long>(ctx, k0.p, c0.p, k1.p, v1.p, cnt.p + nb + 1, cnt.p + nb + 2);
  DBuf<uint8_t> head; head.alloc(M, ctx.stream);
  int h3[3] = {0, 0, 0};
  for (int attempt = 0; attempt < 2; ++attempt) {
    B2R_LAUNCH(ctx, vg_head_kernel<unsigned long long>, nb, 256, 0, k1.p, M, cnt.p + nb + 1, head.p);
    flags_block_offsets(ctx, head.p, M, cnt.p);
    B2R_CUDA(cudaMemcpyAsync(h3, cnt.p + nb, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(ctx.stream));
    if (!h3[2]) break;
    voxel_sort_fallback<unsigned long long>(ctx, k0.p, M, k1.p, v1.p, cnt.p + nb + 1);
    B2R_CUDA(cudaMemsetAsync(cnt.p + nb + 2, 0, sizeof(int), ctx.stream));
  }
  const int nseg = h3[0];
  DBuf<int> seg; seg.alloc(nseg, ctx.stream);
  B2R_LAUNCH(ctx, vg_seg_kernel, nb, 256, 0, head.p, M, cnt.p, seg.p);
  DBuf<float4> cen; cen.alloc(nseg, ctx.stream);
  DBuf<uint8_t> vkeep; vkeep.alloc(nseg, ctx.stream);
  // count_threshold_: keep iff count >= min_points_per_voxel (hpp:114)
  B2R_LAUNCH(ctx, vg_centroid_kernel, (nseg + 7) / 8, 256, 0, world.pts.p, v1.p, seg.p, nseg, M, min_points_per_voxel, cen.p, vkeep.p);
  compact_points(ctx, cen.p, vkeep.p, nseg, out);
}

}  // namespace b2r

// ------------------------------------------------------------------------------------------------ other-robot points
// mrg_slam_component.cpp:395-427: a point is removed when it lies within robot_remove_points_radius of any
// other robot's position (sensor frame, float): distSqr = (point - other).squaredNorm() < radius^2, first hit wins.
namespace b2r {

__global__ void robot_flag_kernel(const float4* __restrict__ in, int n, const float* __restrict__ others, int n_others, float r2,
                                  uint8_t* __restrict__ keep, uint8_t* __restrict__ removed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  bool hit = false;
  for (int k = 0; k < n_others && !hit; ++k) {
    const float dx = __fsub_rn(p.x, others[3 * k]), dy = __fsub_rn(p.y, others[3 * k + 1]), dz = __fsub_rn(p.z, others[3 * k + 2]);
    const float d = __fadd_rn(__fmul_rn(dx, dx), __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dz, dz)));  // Vector3f::squaredNorm
    hit = d < r2;
  }
  keep[i] = hit ? 0 : 1;
  removed[i] = hit ? 1 : 0;
}

void filter_robot_points(Ctx& ctx, const float4* in, int n, const float* others_xyz_host, int n_others, float radius_sqr, DevCloud& kept,
                         DevCloud* removed) {
  kept.n = 0;
  if (removed) removed->n = 0;
  if (n == 0) return;
  DBuf<float> dothers; dothers.alloc((size_t)std::max(1, 3 * n_others), ctx.stream);
  if (n_others) B2R_CUDA(cudaMemcpyAsync(dothers.p, others_xyz_host, sizeof(float) * 3 * n_others, cudaMemcpyHostToDevice, ctx.stream));
  DBuf<uint8_t> keep, rem;
  keep.alloc(n, ctx.stream); rem.alloc(n, ctx.stream);
  B2R_LAUNCH(ctx, robot_flag_kernel, (n + 255) / 256, 256, 0, in, n, dothers.p, n_others, radius_sqr, keep.p, rem.p);
  compact_points(ctx, in, keep.p, n, kept);
  if (removed) compact_points(ctx, in, rem.p, n, *removed);
}

}  // namespace b2r
