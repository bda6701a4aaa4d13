// Cloud-side structures: upload, bounding boxes, dense bucket grids (counting sort), exact kNN with
// per-point covariances, VGICP Gaussian voxel maps and NDT voxel grids.
//
// Replaces (reference call sites -> upstream classes):
//   A3/A5  fast_gicp setInputSource/Target + calculate_covariances   (registrations.cpp:55-63,76-84)
//   A6     fast_gicp GaussianVoxelMap::create_voxelmap
//   A10    pclomp::VoxelGridCovariance::filter                        (registrations.cpp:139)
//   E13    pcl::search::KdTree (FLANN) exact kNN                      -> uniform grid, ring expansion
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <algorithm>

#include "internal.hpp"
#include "knn.cuh"

namespace b2r {

CloudView Cloud::view() const {
  CloudView v;
  memset(&v, 0, sizeof(v));
  v.pts = pts.p; v.n = n;
  for (int d = 0; d < 3; ++d) { v.bmin[d] = bmin[d]; v.bmax[d] = bmax[d]; }
  v.h = h; v.inv_h = h > 0 ? 1.0f / h : 0.f;
  v.hx = hx; v.inv_hx = hx > 0 ? 1.0f / hx : 0.f;
  for (int d = 0; d < 3; ++d) v.gd[d] = gd[d];
  v.ncell = ncell; v.cell_start = cell_start.p; v.cell_cnt = cell_cnt.p; v.cell_tmp = cell_tmp.p; v.spts = spts.p; v.spair = spair.p;
  v.cov = cov.p; v.nrm = nrm.p;
  v.vres = vres;
  for (int d = 0; d < 3; ++d) { v.vmin[d] = vmin[d]; v.vd[d] = vd[d]; }
  v.vcell = vcell; v.v_start = v_start.p; v.v_cnt = v_cnt.p; v.v_order = v_order.p; v.v_table = v_table.p; v.vrec = vrec.p;
  v.v_nrec = v_nrec.p; v.v_reccell = v_reccell.p;
  v.leaf = leaf; v.inv_leaf = leaf > 0 ? 1.0f / leaf : 0.f;
  for (int d = 0; d < 3; ++d) { v.min_b[d] = min_b[d]; v.max_b[d] = max_b[d]; v.div_b[d] = div_b[d]; }
  v.ncell_ndt = ncell_ndt; v.ndt_centroids = ndt_centroids ? 1 : 0; v.n_start = n_start.p; v.n_cnt = n_cnt.p; v.n_order = n_order.p; v.n_table = n_table.p; v.nrec = nrec.p;
  v.n_nrec = n_nrec.p; v.n_reccell = n_reccell.p;
  return v;
}

// ------------------------------------------------------------------------------------------------ I/O
__global__ void repack32_kernel(const uint8_t* __restrict__ raw, int n, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = *reinterpret_cast<const float4*>(raw + (size_t)i * 32);
  const float inten = *reinterpret_cast<const float*>(raw + (size_t)i * 32 + 16);
  out[i] = make_float4(a.x, a.y, a.z, inten);
}
__global__ void unpack32_kernel(const float4* __restrict__ in, int n, uint8_t* __restrict__ raw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = in[i];
  *reinterpret_cast<float4*>(raw + (size_t)i * 32) = make_float4(p.x, p.y, p.z, 1.0f);
  *reinterpret_cast<float4*>(raw + (size_t)i * 32 + 16) = make_float4(p.w, 0.f, 0.f, 0.f);
}

void load_points(Ctx& ctx, const void* points, size_t n, size_t stride_bytes, int memspace, DBuf<float4>& dst) {
  if (stride_bytes != 16 && stride_bytes != 32) throw Error(B2R_ERR_INVALID_ARG, "stride_bytes must be 16 or 32");
  dst.alloc(n, ctx.stream);
  if (n == 0) return;
  cudaMemcpyKind kind = memspace == B2R_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (stride_bytes == 16) {
    B2R_CUDA(cudaMemcpyAsync(dst.p, points, n * 16, kind, ctx.stream));
  } else {
    DBuf<uint8_t> raw;
    const uint8_t* src = (const uint8_t*)points;
    if (memspace != B2R_DEVICE) {
      raw.alloc(n * 32, ctx.stream);
      B2R_CUDA(cudaMemcpyAsync(raw.p, points, n * 32, kind, ctx.stream));
      src = raw.p;
    }
    B2R_LAUNCH(ctx, repack32_kernel, (unsigned)((n + 255) / 256), 256, 0, src, (int)n, dst.p);
  }
}

void store_points(Ctx& ctx, const float4* src, size_t n, void* out, size_t stride_bytes, int memspace) {
  if (stride_bytes != 16 && stride_bytes != 32) throw Error(B2R_ERR_INVALID_ARG, "stride_bytes must be 16 or 32");
  if (n == 0) return;
  cudaMemcpyKind kind = memspace == B2R_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (stride_bytes == 16) {
    B2R_CUDA(cudaMemcpyAsync(out, src, n * 16, kind, ctx.stream));
  } else {
    DBuf<uint8_t> raw;
    uint8_t* dst = (uint8_t*)out;
    if (memspace != B2R_DEVICE) { raw.alloc(n * 32, ctx.stream); dst = raw.p; }
    B2R_LAUNCH(ctx, unpack32_kernel, (unsigned)((n + 255) / 256), 256, 0, src, (int)n, dst);
    if (memspace != B2R_DEVICE) B2R_CUDA(cudaMemcpyAsync(out, raw.p, n * 32, kind, ctx.stream));
  }
  if (memspace != B2R_DEVICE) B2R_CUDA(cudaStreamSynchronize(ctx.stream));
}

// ------------------------------------------------------------------------------------------------ bbox
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
static inline float ordered_to_float(int i) {
  int b = i >= 0 ? i : i ^ 0x7fffffff;
  float f;
  memcpy(&f, &b, 4);
  return f;
}

// bbox_out[cloud][6] ordered-int encoded (min x,y,z, max x,y,z); must be pre-initialised.
// COPY: the points are still in the caller's device buffers srcs[cloud]; the same pass copies them into the cloud's own
// storage (one kernel for a whole batch of clouds instead of one memcpy per cloud plus a second read for the boxes).
template <bool COPY>
__global__ void bbox_kernel(const CloudView* __restrict__ views, int* __restrict__ bbox_out, const float4* const* __restrict__ srcs) {
  const CloudView& c = views[blockIdx.y];
  const float4* __restrict__ src = COPY ? srcs[blockIdx.y] : c.pts;
  float4* __restrict__ dst = const_cast<float4*>(c.pts);
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(&src[i]);
    if (COPY) dst[i] = p;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  if ((threadIdx.x & 31) == 0) {
    int* o = bbox_out + blockIdx.y * 6;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      atomicMin(&o[d], float_to_ordered(mn[d]));
      atomicMax(&o[3 + d], float_to_ordered(mx[d]));
    }
  }
}
__global__ void bbox_init_kernel(int* bbox_out, int nclouds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nclouds * 6) bbox_out[i] = (i % 6) < 3 ? INT_MAX : INT_MIN;
}

// ------------------------------------------------------------------------------------------------ bucket grids
enum { GRID_NN = 0, GRID_VGICP = 1, GRID_NDT = 2 };

template <int MODE>
__device__ __forceinline__ int cell_key(const CloudView& c, const float4& p) {
  if (MODE == GRID_NN) {
    int cx, cy, cz;
    nn_cell_of(c, p.x, p.y, p.z, cx, cy, cz);
    return (cz * c.gd[1] + cy) * c.gd[0] + cx;
  } else if (MODE == GRID_VGICP) {
    int x = vgicp_coord_d((double)p.x, c.vres) - c.vmin[0];
    int y = vgicp_coord_d((double)p.y, c.vres) - c.vmin[1];
    int z = vgicp_coord_d((double)p.z, c.vres) - c.vmin[2];
    return (z * c.vd[1] + y) * c.vd[0] + x;
  } else {
    // pclomp::VoxelGridCovariance build key (float math, multiply by inverse leaf): SURVEY A.4
    int i0 = (int)(floorf(__fmul_rn(p.x, c.inv_leaf)) - (float)c.min_b[0]);
    int i1 = (int)(floorf(__fmul_rn(p.y, c.inv_leaf)) - (float)c.min_b[1]);
    int i2 = (int)(floorf(__fmul_rn(p.z, c.inv_leaf)) - (float)c.min_b[2]);
    return i0 + i1 * c.div_b[0] + i2 * c.div_b[0] * c.div_b[1];
  }
}
template <int MODE>
__device__ __forceinline__ int* grid_cnt(const CloudView& c) { return MODE == GRID_NN ? c.cell_cnt : (MODE == GRID_VGICP ? c.v_cnt : c.n_cnt); }
template <int MODE>
__device__ __forceinline__ int* grid_start(const CloudView& c) { return MODE == GRID_NN ? c.cell_start : (MODE == GRID_VGICP ? c.v_start : c.n_start); }
template <int MODE>
__device__ __forceinline__ int grid_ncell(const CloudView& c) { return MODE == GRID_NN ? c.ncell : (MODE == GRID_VGICP ? c.vcell : c.ncell_ndt); }

template <int MODE>
__global__ void grid_count_kernel(const CloudView* __restrict__ views) {
  const CloudView& c = views[blockIdx.y];
  int* cnt = grid_cnt<MODE>(c);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(&c.pts[i]);
    atomicAdd(&cnt[cell_key<MODE>(c, p)], 1);
  }
}

// Exclusive scan of the per-cell counts into start[0..ncell] in three phases (tile-local scan, scan of the tile
// totals, apply); counts are reset to 0 (they become the scatter cursors).  For voxel grids the same pass assigns
// compact record ids to the occupied cells (64-bit packed scan: low word points, high word occupied cells).
constexpr int kScanItems = 8;
constexpr int kScanTile = 1024 * kScanItems;

template <int MODE>
__global__ void __launch_bounds__(1024) grid_scan_tile_kernel(const CloudView* __restrict__ views, unsigned long long* __restrict__ tile_tot,
                                                              int max_tiles) {
  const CloudView& c = views[blockIdx.y];
  const int ncell = grid_ncell<MODE>(c);
  const int base = blockIdx.x * kScanTile;
  if (base >= ncell) return;
  int* cnt = grid_cnt<MODE>(c);
  int* start = grid_start<MODE>(c);
  int* table = MODE == GRID_VGICP ? c.v_table : (MODE == GRID_NDT ? c.n_table : nullptr);
  __shared__ unsigned long long warp_tot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int v[kScanItems];
  unsigned long long tsum = 0ull;
  const int i0 = base + threadIdx.x * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (i0 + k < ncell) ? cnt[i0 + k] : 0;
    tsum += (unsigned long long)v[k] + (v[k] > 0 ? (1ull << 32) : 0ull);
  }
  unsigned long long incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = warp_tot[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_tot[lane] = wi - w;
    if (lane == 31) tile_tot[(size_t)blockIdx.y * max_tiles + blockIdx.x] = wi;
  }
  __syncthreads();
  unsigned long long run = warp_tot[warp] + (incl - tsum);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (i0 + k < ncell) {
      start[i0 + k] = (int)(run & 0xffffffffull);
      if (MODE != GRID_NN) table[i0 + k] = v[k] > 0 ? (int)(run >> 32) : -1;
      cnt[i0 + k] = 0;
    }
    run += (unsigned long long)v[k] + (v[k] > 0 ? (1ull << 32) : 0ull);
  }
}

// one block per cloud: exclusive scan of the tile totals (in place); writes start[ncell] and the record count
template <int MODE>
__global__ void __launch_bounds__(1024) grid_scan_tops_kernel(const CloudView* __restrict__ views, unsigned long long* __restrict__ tile_tot,
                                                              int max_tiles) {
  const CloudView& c = views[blockIdx.x];
  const int ncell = grid_ncell<MODE>(c);
  const int ntiles = (ncell + kScanTile - 1) / kScanTile;
  unsigned long long* tt = tile_tot + (size_t)blockIdx.x * max_tiles;
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < ntiles; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned long long v = i < ntiles ? tt[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    const unsigned long long excl = carry_s + warp_tot[warp] + incl - v;
    if (i < ntiles) tt[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    grid_start<MODE>(c)[ncell] = (int)(carry_s & 0xffffffffull);
    if (MODE == GRID_VGICP) c.v_nrec[0] = (int)(carry_s >> 32);
    if (MODE == GRID_NDT) c.n_nrec[0] = (int)(carry_s >> 32);
  }
}

template <int MODE>
__global__ void __launch_bounds__(1024) grid_scan_apply_kernel(const CloudView* __restrict__ views, const unsigned long long* __restrict__ tile_tot,
                                                               int max_tiles) {
  const CloudView& c = views[blockIdx.y];
  const int ncell = grid_ncell<MODE>(c);
  const int base = blockIdx.x * kScanTile;
  if (base >= ncell) return;
  const unsigned long long off = tile_tot[(size_t)blockIdx.y * max_tiles + blockIdx.x];
  const int off_lo = (int)(off & 0xffffffffull), off_hi = (int)(off >> 32);
  int* start = grid_start<MODE>(c);
  int* table = MODE == GRID_VGICP ? c.v_table : (MODE == GRID_NDT ? c.n_table : nullptr);
  int* reccell = MODE == GRID_VGICP ? c.v_reccell : (MODE == GRID_NDT ? c.n_reccell : nullptr);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int i = base + k * 1024 + threadIdx.x;
    if (i < ncell) {
      start[i] += off_lo;
      if (MODE != GRID_NN) {
        const int t = table[i];
        if (t >= 0) { table[i] = t + off_hi; reccell[t + off_hi] = i; }
      }
    }
  }
}

// scatter of the point indices into their cells (atomic cursor: arbitrary order inside a cell) ...
template <int MODE>
__global__ void grid_scatter_kernel(const CloudView* __restrict__ views) {
  const CloudView& c = views[blockIdx.y];
  int* cnt = grid_cnt<MODE>(c);
  const int* start = grid_start<MODE>(c);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(&c.pts[i]);
    const int key = cell_key<MODE>(c, p);
    const int pos = start[key] + atomicAdd(&cnt[key], 1);
    if (MODE == GRID_NN) c.cell_tmp[pos] = make_int2(float_order_key(p.x), i);
    else if (MODE == GRID_VGICP) c.v_order[c.n + pos] = i;
    else c.n_order[c.n + pos] = i;
  }
}
// ... then every entry finds its rank inside its cell and moves there, so that the result is identical from run to run
// (tie-breaks by position, summation orders).  Voxel lists: ascending point index.  NN grid: ascending (x, point index),
// which makes every row of cells along x one run sorted by x (knn.cuh sweeps it outwards from the query).  One thread
// per entry; the cell's list is read by all of its entries (L1-resident).
template <int MODE>
__global__ void grid_rank_kernel(const CloudView* __restrict__ views) {
  const CloudView& c = views[blockIdx.y];
  const int* start = grid_start<MODE>(c);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < c.n; j += gridDim.x * blockDim.x) {
    if (MODE == GRID_NN) {
      const int2 me = c.cell_tmp[j];
      const float4 p = __ldg(&c.pts[me.y]);
      const int key = cell_key<MODE>(c, p);
      const int s = start[key], e = start[key + 1];
      int rank = 0;
      for (int t = s; t < e; ++t) {
        const int2 o = c.cell_tmp[t];
        rank += (o.x < me.x || (o.x == me.x && o.y < me.y)) ? 1 : 0;
      }
      const int pos = s + rank;
      c.spts[pos] = make_float4(p.x, p.y, p.z, __int_as_float(me.y));
      float* pr = reinterpret_cast<float*>(c.spair) + (size_t)(pos >> 1) * 8 + (pos & 1);
      pr[0] = p.x; pr[2] = p.y; pr[4] = p.z; pr[6] = __int_as_float(me.y);
      if (j == 0 && (c.n & 1)) {  // the odd cloud's last record: its second point lies at +inf
        float* pe = reinterpret_cast<float*>(c.spair) + (size_t)(c.n >> 1) * 8 + 1;
        pe[0] = INFINITY; pe[2] = INFINITY; pe[4] = INFINITY; pe[6] = __int_as_float(-1);
      }
    } else {
      int* order = MODE == GRID_VGICP ? c.v_order : c.n_order;
      const int* tmp = order + c.n;
      const int i = tmp[j];
      const int key = cell_key<MODE>(c, __ldg(&c.pts[i]));
      const int s = start[key], e = start[key + 1];
      int rank = 0;
      for (int t = s; t < e; ++t) rank += (tmp[t] < i) ? 1 : 0;
      order[s + rank] = i;
    }
  }
}

// VGICP: one warp per occupied voxel: sum of the points' positions and covariances (fast_gicp create_voxelmap,
// SURVEY A.2) over the index-sorted list, lane-strided partials + fixed-order warp reduction, then divide by the count.
__global__ void __launch_bounds__(256) vgicp_reduce_kernel(const CloudView* __restrict__ views) {
  const CloudView& c = views[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const int nrec = c.v_nrec[0];
  for (int rec = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rec < nrec; rec += nwarps) {
    const int cell = c.v_reccell[rec];
    const int s = c.v_start[cell], n = c.v_start[cell + 1] - s;
    const int* idx = c.v_order + s;  // ascending point index: fixed summation order
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = lane; j < n; j += 32) {
      const int i = idx[j];
      const float4 p = __ldg(&c.pts[i]);
      a[0] += (double)p.x; a[1] += (double)p.y; a[2] += (double)p.z;
      const double* pc = c.cov + (size_t)i * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k) a[3 + k] += pc[k];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = warp_sum(a[k]);
    if (lane == 0) {
      const double nn = (double)n;
      VoxRec r;
      r.mean[0] = a[0] / nn; r.mean[1] = a[1] / nn; r.mean[2] = a[2] / nn;
#pragma unroll
      for (int k = 0; k < 6; ++k) r.cov[k] = a[3 + k] / nn;
      r.w = sqrt(nn);
      r.n = n;
      r.cell = cell;
      r.pad = 0.0;
      c.vrec[rec] = r;
    }
  }
}

// NDT: per-voxel mean, single-pass covariance, eigenvalue clamp and inverse (pclomp::VoxelGridCovariance)
__global__ void __launch_bounds__(256) ndt_reduce_kernel(const CloudView* __restrict__ views) {
  const CloudView& c = views[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const int nrec = c.n_nrec[0];
  for (int rec = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rec < nrec; rec += nwarps) {
    const int cell = c.n_reccell[rec];
    const int s = c.n_start[cell], e = c.n_start[cell + 1];
    const int* idx = c.n_order + s;  // ascending point index: fixed summation order
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // sum x,y,z ; sum xx,xy,xz,yy,yz,zz
    for (int j = lane; j < e - s; j += 32) {
      const float4 p = __ldg(&c.pts[idx[j]]);
      const double qx = (double)p.x, qy = (double)p.y, qz = (double)p.z;
      acc[0] += qx; acc[1] += qy; acc[2] += qz;
      acc[3] += qx * qx; acc[4] += qx * qy; acc[5] += qx * qz; acc[6] += qy * qy; acc[7] += qy * qz; acc[8] += qz * qz;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = warp_sum(acc[k]);
    if (lane != 0) continue;
    float cen[3] = {0.f, 0.f, 0.f};
    if (c.ndt_centroids) {  // the kd-tree's points: sequential FLOAT sums in point order, divided by the float count
      for (int j = 0; j < e - s; ++j) {
        const float4 p = __ldg(&c.pts[idx[j]]);
        cen[0] = __fadd_rn(cen[0], p.x); cen[1] = __fadd_rn(cen[1], p.y); cen[2] = __fadd_rn(cen[2], p.z);
      }
      const float fn = (float)(e - s);
      cen[0] = __fdiv_rn(cen[0], fn); cen[1] = __fdiv_rn(cen[1], fn); cen[2] = __fdiv_rn(cen[2], fn);
    }
    double sum[3] = {acc[0], acc[1], acc[2]};
    double cov[9] = {acc[3], acc[4], acc[5], acc[4], acc[6], acc[7], acc[5], acc[7], acc[8]};
    NdtRec r;
    const int npts = e - s;
    const double nn = (double)npts;
    double mean[3] = {sum[0] / nn, sum[1] / nn, sum[2] / nn};
    r.mean[0] = mean[0]; r.mean[1] = mean[1]; r.mean[2] = mean[2];
    r.n = npts;
    r.cell = cell;
    r.centroid[0] = cen[0]; r.centroid[1] = cen[1]; r.centroid[2] = cen[2]; r.pad = 0.f;
#pragma unroll
    for (int a = 0; a < 9; ++a) { r.icov[a] = 0.f; r.icov_d[a] = 0.0; }
    if (npts >= 6) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) cov[a * 3 + b] = (cov[a * 3 + b] - 2 * (sum[a] * mean[b])) / nn + mean[a] * mean[b];
#pragma unroll
      for (int a = 0; a < 9; ++a) cov[a] *= (nn - 1.0) / nn;
      double S[9];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) S[a * 3 + b] = (a >= b) ? cov[a * 3 + b] : cov[b * 3 + a];
      double ev[3], V[9];
      sym3_eigen_dev(S, ev, V);
      if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) {
        r.n = -1;
      } else {
        double min_ev = 0.01 * ev[2];
        if (ev[0] < min_ev) {
          ev[0] = min_ev;
          if (ev[1] < min_ev) ev[1] = min_ev;
          // cov = V diag(ev) V^-1 (V orthonormal: inverse by cofactors as upstream does through Eigen)
          double VL[9], Vi[9];
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < 3; ++j) VL[a * 3 + j] = V[a * 3 + j] * ev[j];
          {
            double c00 = V[4] * V[8] - V[5] * V[7], c01 = V[5] * V[6] - V[3] * V[8], c02 = V[3] * V[7] - V[4] * V[6];
            double id = 1.0 / (V[0] * c00 + V[1] * c01 + V[2] * c02);
            Vi[0] = c00 * id; Vi[1] = (V[2] * V[7] - V[1] * V[8]) * id; Vi[2] = (V[1] * V[5] - V[2] * V[4]) * id;
            Vi[3] = c01 * id; Vi[4] = (V[0] * V[8] - V[2] * V[6]) * id; Vi[5] = (V[2] * V[3] - V[0] * V[5]) * id;
            Vi[6] = c02 * id; Vi[7] = (V[1] * V[6] - V[0] * V[7]) * id; Vi[8] = (V[0] * V[4] - V[1] * V[3]) * id;
          }
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) cov[a * 3 + b] = VL[a * 3 + 0] * Vi[0 * 3 + b] + VL[a * 3 + 1] * Vi[1 * 3 + b] + VL[a * 3 + 2] * Vi[2 * 3 + b];
        }
        double c00 = cov[4] * cov[8] - cov[5] * cov[7], c01 = cov[5] * cov[6] - cov[3] * cov[8], c02 = cov[3] * cov[7] - cov[4] * cov[6];
        double id = 1.0 / (cov[0] * c00 + cov[1] * c01 + cov[2] * c02);
        double ic[9];
        ic[0] = c00 * id; ic[1] = (cov[2] * cov[7] - cov[1] * cov[8]) * id; ic[2] = (cov[1] * cov[5] - cov[2] * cov[4]) * id;
        ic[3] = c01 * id; ic[4] = (cov[0] * cov[8] - cov[2] * cov[6]) * id; ic[5] = (cov[2] * cov[3] - cov[0] * cov[5]) * id;
        ic[6] = c02 * id; ic[7] = (cov[1] * cov[6] - cov[0] * cov[7]) * id; ic[8] = (cov[0] * cov[4] - cov[1] * cov[3]) * id;
        bool bad = false;
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          if (isinf(ic[a])) bad = true;
          r.icov[a] = (float)ic[a];
          r.icov_d[a] = ic[a];
        }
        if (bad) r.n = -1;
      }
    }
    c.nrec[rec] = r;
  }
}

// ------------------------------------------------------------------------------------------------ kNN covariances
// fast_gicp calculate_covariances in one kernel, one thread per (cell-sorted) point:
//   pass 1  exact k-th nearest squared distance dk (knn.cuh, K floats in registers);
//   pass 2  the same cells again, accumulating sum(d) and sum(d d^T) of d = p - query in double over the points with
//           d2 <= dk (the k nearest; ties beyond k are resolved towards the lower position on a rare slow path);
//   then    covariance = E[d d^T] - E[d] E[d]^T (shift-invariant, so identical to the centred sum of the reference up
//           to rounding ~1e-15), symmetric 3x3 Jacobi eigen-decomposition, PLANE regularisation (1, 1, 1e-3).
struct CovAccum {
  double s[9];
  int cnt;
};
struct CovVisitor {
  float qx, qy, qz, dk;
  bool ties;  // accept d2 == dk as well as d2 < dk
  int k;
  int32_t* knn_row;
  CovAccum a;
  __device__ __forceinline__ float thr() const { return dk * (1.f + 1e-6f); }
  __device__ __forceinline__ bool stop(float s) const { return s > dk * kGapSlack; }
  __device__ __forceinline__ void add(const float4& p) {
    const double dx = (double)p.x - (double)qx, dy = (double)p.y - (double)qy, dz = (double)p.z - (double)qz;
    a.s[0] += dx; a.s[1] += dy; a.s[2] += dz;
    a.s[3] += dx * dx; a.s[4] += dx * dy; a.s[5] += dx * dz; a.s[6] += dy * dy; a.s[7] += dy * dz; a.s[8] += dz * dz;
    if (knn_row && a.cnt < k) knn_row[a.cnt] = __float_as_int(p.w);
    ++a.cnt;
  }
  __device__ __forceinline__ float qy_f() const { return qy; }
  __device__ __forceinline__ float qz_f() const { return qz; }
  __device__ __forceinline__ void test(const float4& p, int, float d2) {
    if (d2 < dk || (ties && d2 == dk)) add(p);
  }
};
// lowest position > after whose distance equals dk exactly
struct TieVisitor {
  float qx, qy, qz, dk;
  int after, found;
  __device__ __forceinline__ float thr() const { return dk * (1.f + 1e-6f); }
  __device__ __forceinline__ bool stop(float s) const { return s > dk * kGapSlack; }
  __device__ __forceinline__ float qy_f() const { return qy; }
  __device__ __forceinline__ float qz_f() const { return qz; }
  __device__ __forceinline__ void test(const float4&, int j, float d2) {
    if (j > after && j < found && d2 == dk) found = j;
  }
};
// more than k points within dk (exact ties at the k-th distance): strictly closer ones, then ties by ascending position.
// Rare; everything goes in and out by value so that the caller's accumulators stay in registers.
__device__ __noinline__ CovAccum cov_accumulate_ties(const CloudView* cp, float qx, float qy, float qz, int r, float dk, int k,
                                                     int32_t* knn_row) {
  const CloudView& c = *cp;
  const QueryCell q = query_cell(c, qx, qy, qz);
  CovVisitor v{qx, qy, qz, dk, false, k, knn_row, {{0, 0, 0, 0, 0, 0, 0, 0, 0}, 0}};
  visit_rows(c, q, qx, r, false, v);
  int last = -1;
  while (v.a.cnt < k) {
    TieVisitor tv{qx, qy, qz, dk, last, INT_MAX};
    visit_rows(c, q, qx, r, false, tv);
    if (tv.found == INT_MAX) break;
    v.add(__ldg(&c.spts[tv.found]));
    last = tv.found;
  }
  return v.a;
}

// covariance = E[d d^T] - E[d] E[d]^T of the k neighbours, symmetric 3x3 Jacobi eigen-decomposition, PLANE regularisation
// (singular values replaced by (1, 1, 1e-3), smallest = plane normal), stored in original point order
__device__ __forceinline__ void cov_store(const CloudView& c, int orig, const CovAccum& a, int k) {
  const double kk = (double)k;
  const double mx = a.s[0] / kk, my = a.s[1] / kk, mz = a.s[2] / kk;
  const double m0 = a.s[3] / kk - mx * mx, m1 = a.s[4] / kk - mx * my, m2 = a.s[5] / kk - mx * mz;
  const double m3 = a.s[6] / kk - my * my, m4 = a.s[7] / kk - my * mz, m5 = a.s[8] / kk - mz * mz;
  double S[9] = {m0, m1, m2, m1, m3, m4, m2, m4, m5};
  double ev[3], V[9];
  sym3_eigen_dev(S, ev, V);
  const double vals[3] = {1e-3, 1.0, 1.0};
  double o[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double aa = V[0 * 3 + j], b = V[1 * 3 + j], cc = V[2 * 3 + j];
    o[0] += vals[j] * aa * aa; o[1] += vals[j] * aa * b; o[2] += vals[j] * aa * cc;
    o[3] += vals[j] * b * b; o[4] += vals[j] * b * cc; o[5] += vals[j] * cc * cc;
  }
  double* dst = c.cov + (size_t)orig * 6;
#pragma unroll
  for (int t = 0; t < 6; ++t) dst[t] = o[t];
  if (c.nrm) {  // the eigenvector that got 1e-3: o = V V^T - (1 - 1e-3) n n^T
    double4* dn = reinterpret_cast<double4*>(c.nrm) + orig;
    *dn = make_double4(V[0], V[3], V[6], 0.0);
  }
}

__device__ unsigned long long g_knn_list_overflows = 0;  // queries whose candidate log overflowed (second traversal taken)
#ifdef B2R_KNN_STATS
__device__ unsigned long long g_knn_hist[66];
#endif
constexpr int kKnnThreads = 128;
#ifndef B2R_KNN_CAP
#define B2R_KNN_CAP 24
#endif
#ifndef B2R_KNN_BLOCKS
#define B2R_KNN_BLOCKS 8
#endif
// logged candidates per query kept in shared memory (12 KB per block).  Measured 8..48: the more of the unified L1 / shared
// memory array is left to L1 the better (the candidates are re-read from neighbouring queries): 48 -> 1.03 ms, 32 -> 0.97,
// 24 -> 0.96, 8 -> 0.97 on the headline step (the overflow goes to L1-backed local memory and costs the same)
constexpr int kKnnListCap = B2R_KNN_CAP;
constexpr int kKnnListSpill = 120 - B2R_KNN_CAP;   // further ones in per-thread local memory

template <int K>
__global__ void __launch_bounds__(kKnnThreads, B2R_KNN_BLOCKS) knn_cov_kernel(const CloudView* __restrict__ views, int k, int32_t* __restrict__ knn_out) {
  __shared__ int s_list[kKnnListCap * kKnnThreads];
  const CloudView& c = views[blockIdx.y];
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= c.n) return;
  const float4 sp = __ldg(&c.spts[qi]);
  const int orig = __float_as_int(sp.w);
  const QueryCell q = query_cell(c, sp.x, sp.y, sp.z);
  int spill[kKnnListSpill];
  TopkListVisitor<K, kKnnListCap, kKnnListSpill> tv(sp.x, sp.y, sp.z, s_list + threadIdx.x, kKnnThreads, spill);
  const int r = knn_topk<K>(c, q, k, tv);
  const float dk = topk_kth<K>(tv.d, k);
  const int logged = tv.cnt;
#ifdef B2R_KNN_STATS
  atomicAdd(&g_knn_hist[min(logged / 4, 63)], 1ull);
  atomicAdd(&g_knn_hist[64], (unsigned long long)tv.tested);
  atomicAdd(&g_knn_hist[65], (unsigned long long)r);
#endif
  CovVisitor v{sp.x, sp.y, sp.z, dk, true, k, knn_out ? knn_out + (size_t)orig * k : nullptr, {{0, 0, 0, 0, 0, 0, 0, 0, 0}, 0}};
  if (logged <= kKnnListCap + kKnnListSpill) {
    // pass 2 over the logged candidates only, in traversal order.  Branch-free: a rejected candidate is replaced by the
    // query itself, whose differences are exactly +0.0 and leave every sum unchanged.
    const double qxd = (double)sp.x, qyd = (double)sp.y, qzd = (double)sp.z;
    for (int t = 0; t < logged; ++t) {
      const float4 p = __ldg(&c.spts[tv.logged(t)]);
      const float d2 = dist2_flann(sp.x, sp.y, sp.z, p.x, p.y, p.z);
      const bool ok = d2 <= dk;
      const double dx = (double)(ok ? p.x : sp.x) - qxd, dy = (double)(ok ? p.y : sp.y) - qyd, dz = (double)(ok ? p.z : sp.z) - qzd;
      v.a.s[0] += dx; v.a.s[1] += dy; v.a.s[2] += dz;
      v.a.s[3] += dx * dx; v.a.s[4] += dx * dy; v.a.s[5] += dx * dz; v.a.s[6] += dy * dy; v.a.s[7] += dy * dz; v.a.s[8] += dz * dz;
      if (v.knn_row && ok && v.a.cnt < k) v.knn_row[v.a.cnt] = __float_as_int(p.w);
      v.a.cnt += ok ? 1 : 0;
    }
  } else {
    atomicAdd(&g_knn_list_overflows, 1ull);
    visit_rows(c, q, sp.x, r, false, v);  // log overflowed: same rows again, radius = k-th distance
  }
  if (v.a.cnt > k) v.a = cov_accumulate_ties(&c, sp.x, sp.y, sp.z, r, dk, k, v.knn_row);
  cov_store(c, orig, v.a, k);
}

// ---- the same computation with the candidate rows staged in shared memory by TMA (north star: "shared-memory-staged tiles, TMA
// where it fits").  A block's 128 cell-sorted queries sit in one or two (y, z) rows and a few metres of x, so the points all of
// them can touch form, per neighbouring row, ONE contiguous range of spts: the strip [x_min - h, x_max + h] of that row.  One
// thread computes the strips of the rectangle of rows around the block and issues one cp.async.bulk (global -> shared, completion
// on an mbarrier) per row; the search and the covariance pass then read their candidates from shared memory (positions outside
// the staged strip, e.g. a second ring, fall back to global memory: the tile is a cache, never a correctness condition).
// Compared with knn_cov_kernel: no dependent L1 / L2 round trip per candidate (ncu: 3.0 long-scoreboard stalls per issue
// there), and no candidate log at all — the covariance pass simply sweeps the staged rows again with the proven k-th distance
// as its radius, which removes the log's shared-memory columns, its local-memory spill (72 ints per thread, 1.7x DRAM traffic)
// and the log bookkeeping in the inner loop.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned phase) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  return ok != 0;
}

constexpr int kTilePts = 2560;  // staged points per block (40 KB)
constexpr int kTileRows = 32;   // rows of the staged rectangle

template <int K>
__global__ void __launch_bounds__(kKnnThreads, 5) knn_cov_tile_kernel(const CloudView* __restrict__ views, int k, int32_t* __restrict__ knn_out) {
  __shared__ __align__(128) float4 s_tile[kTilePts];
  __shared__ TileRow s_rows[kTileRows];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_ext[6];  // min cy, max cy, min cz, max cz, min / max ordered x bits
  __shared__ int s_total;
  const CloudView& c = views[blockIdx.y];
  const int q0 = blockIdx.x * blockDim.x;
  if (q0 >= c.n) return;
  const int qi = min(q0 + (int)threadIdx.x, c.n - 1);  // the last block's spare threads repeat its last query (no store)
  const bool live = q0 + (int)threadIdx.x < c.n;
  const float4 sp = __ldg(&c.spts[qi]);
  const int orig = __float_as_int(sp.w);
  const QueryCell q = query_cell(c, sp.x, sp.y, sp.z);
  // ---- the block's rectangle of rows and its x extent
  if (threadIdx.x == 0) { s_ext[0] = INT_MAX; s_ext[1] = INT_MIN; s_ext[2] = INT_MAX; s_ext[3] = INT_MIN; s_ext[4] = INT_MAX; s_ext[5] = INT_MIN; }
  __syncthreads();
  {
    int cy0 = q.cy, cy1 = q.cy, cz0 = q.cz, cz1 = q.cz, x0 = float_order_key(sp.x), x1 = x0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cy0 = min(cy0, __shfl_xor_sync(0xffffffffu, cy0, o)); cy1 = max(cy1, __shfl_xor_sync(0xffffffffu, cy1, o));
      cz0 = min(cz0, __shfl_xor_sync(0xffffffffu, cz0, o)); cz1 = max(cz1, __shfl_xor_sync(0xffffffffu, cz1, o));
      x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&s_ext[0], cy0); atomicMax(&s_ext[1], cy1); atomicMin(&s_ext[2], cz0); atomicMax(&s_ext[3], cz1);
      atomicMin(&s_ext[4], x0); atomicMax(&s_ext[5], x1);
    }
  }
  __syncthreads();
  const int y_base = max(s_ext[0] - 1, 0), y_end = min(s_ext[1] + 1, c.gd[1] - 1);
  const int z_base = max(s_ext[2] - 1, 0), z_end = min(s_ext[3] + 1, c.gd[2] - 1);
  const int ny = y_end - y_base + 1, nz = z_end - z_base + 1;
  const int nrows = ny * nz <= kTileRows ? ny * nz : 0;  // a block that straddles many rows (sparse regions) is not staged
  if ((int)threadIdx.x < nrows) {
    const int xb0 = s_ext[4], xb1 = s_ext[5];
    const float xmin = __int_as_float(xb0 >= 0 ? xb0 : xb0 ^ 0x7fffffff), xmax = __int_as_float(xb1 >= 0 ? xb1 : xb1 ^ 0x7fffffff);
    const int cx_lo = clampi((int)floorf((xmin - c.h - c.bmin[0]) * c.inv_hx), 0, c.gd[0] - 1);
    const int cx_hi = clampi((int)floorf((xmax + c.h - c.bmin[0]) * c.inv_hx), 0, c.gd[0] - 1);
    const int rowbase = ((z_base + (int)threadIdx.x / ny) * c.gd[1] + y_base + (int)threadIdx.x % ny) * c.gd[0];
    s_rows[threadIdx.x] = TileRow{__ldg(&c.cell_start[rowbase + cx_lo]), __ldg(&c.cell_start[rowbase + cx_hi + 1]), 0};
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int off = 0;
    for (int t = 0; t < nrows; ++t) {  // strips in row order until the tile is full; the rest stays in global memory
      TileRow r = s_rows[t];
      if (off + (r.hi - r.lo) > kTilePts) r.hi = r.lo;
      r.off = off;
      off += r.hi - r.lo;
      s_rows[t] = r;
    }
    s_total = off;
    if (off > 0) {
      mbar_init(&s_bar, 1);
      mbar_expect_tx(&s_bar, (unsigned)off * 16u);
      for (int t = 0; t < nrows; ++t) {
        const TileRow r = s_rows[t];
        if (r.hi > r.lo) tma_bulk_g2s(&s_tile[r.off], &c.spts[r.lo], (unsigned)(r.hi - r.lo) * 16u, &s_bar);
      }
    }
  }
  __syncthreads();
  if (s_total > 0) {
    while (!mbar_try_wait(&s_bar, 0)) {}
  }
  const TilePts src{s_tile, s_rows, y_base, z_base, nrows ? ny : 0, nrows ? nz : 0};
  // ---- pass 1: exact k-th distance; pass 2: moments of the points within it (both from the tile)
  TopkVisitor<K> tv(sp.x, sp.y, sp.z);
  const int r = knn_topk<K>(c, q, k, tv, src);
  const float dk = topk_kth<K>(tv.d, k);
  CovVisitor v{sp.x, sp.y, sp.z, dk, true, k, (knn_out && live) ? knn_out + (size_t)orig * k : nullptr, {{0, 0, 0, 0, 0, 0, 0, 0, 0}, 0}};
  visit_rows(c, q, sp.x, r, false, v, src);
  if (v.a.cnt > k) v.a = cov_accumulate_ties(&c, sp.x, sp.y, sp.z, r, dk, k, v.knn_row);
  if (live) cov_store(c, orig, v.a, k);
}

// pcl::GeneralizedIterativeClosestPoint::computeCovariances (registrations.cpp:93-116 "GICP" / "GICP_OMP") from the
// neighbour lists knn_cov_kernel writes: raw moments with FLOAT products widened to double, E[x x^T] - mean mean^T, the direction
// of the smallest singular value (|eigenvalue|: the float products can make the matrix slightly indefinite) gets gicp_epsilon,
// the other two 1.  One thread per point (original order); overwrites the fast_gicp covariance knn_cov_kernel left there.
__global__ void pcl_cov_kernel(const CloudView* __restrict__ views, const int32_t* __restrict__ knn, int k, double eps) {
  const CloudView& c = views[0];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  double mean[3] = {0, 0, 0}, m[6] = {0, 0, 0, 0, 0, 0};  // xx, yx, yy, zx, zy, zz
  for (int j = 0; j < k; ++j) {
    const float4 p = __ldg(&c.pts[knn[(size_t)i * k + j]]);
    mean[0] += (double)p.x; mean[1] += (double)p.y; mean[2] += (double)p.z;
    m[0] += (double)__fmul_rn(p.x, p.x);
    m[1] += (double)__fmul_rn(p.y, p.x); m[2] += (double)__fmul_rn(p.y, p.y);
    m[3] += (double)__fmul_rn(p.z, p.x); m[4] += (double)__fmul_rn(p.z, p.y); m[5] += (double)__fmul_rn(p.z, p.z);
  }
  const double kk = (double)k;
  for (int a = 0; a < 3; ++a) mean[a] /= kk;
  const double cxx = m[0] / kk - mean[0] * mean[0], cyx = m[1] / kk - mean[1] * mean[0], cyy = m[2] / kk - mean[1] * mean[1];
  const double czx = m[3] / kk - mean[2] * mean[0], czy = m[4] / kk - mean[2] * mean[1], czz = m[5] / kk - mean[2] * mean[2];
  double S[9] = {cxx, cyx, czx, cyx, cyy, czy, czx, czy, czz};
  double ev[3], V[9];
  sym3_eigen_dev(S, ev, V);
  int small = 0;
  if (fabs(ev[1]) < fabs(ev[small])) small = 1;
  if (fabs(ev[2]) < fabs(ev[small])) small = 2;
  double o[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double v = j == small ? eps : 1.0;
    const double a = V[0 * 3 + j], b = V[1 * 3 + j], cc = V[2 * 3 + j];
    o[0] += v * a * a; o[1] += v * a * b; o[2] += v * a * cc;
    o[3] += v * b * b; o[4] += v * b * cc; o[5] += v * cc * cc;
  }
  double* dst = c.cov + (size_t)i * 6;
#pragma unroll
  for (int t = 0; t < 6; ++t) dst[t] = o[t];
}

// arbitrary queries (debug / tests): one thread per query, neighbours as (distance, position) keys
__global__ void __launch_bounds__(64) knn_query_kernel(const CloudView* __restrict__ views, const float4* __restrict__ queries, int nq, int k,
                                                        int32_t* __restrict__ idx_out, float* __restrict__ d2_out) {
  const CloudView& c = views[0];
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  const float4 p = __ldg(&queries[qi]);
  const QueryCell q = query_cell(c, p.x, p.y, p.z);
  TopkKeyVisitor<32> v(p.x, p.y, p.z);
  int r = rows_outside(c, q) + 1;
  visit_rows(c, q, p.x, r, false, v);
  for (;;) {
    unsigned long long kth = v.d[31];
#pragma unroll
    for (int i = 30; i >= 0; --i) kth = (i >= k - 1) ? v.d[i] : kth;
    if (kth != ~0ull && __uint_as_float((unsigned)(kth >> 32)) <= ring_safe_d2(r, c.h) + q.xout2) break;
    if (ring_covers_grid(c, q, r)) break;
    ++r;
    visit_rows(c, q, p.x, r, true, v);
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i < k) {
      const unsigned long long key = v.d[i];
      if (key != ~0ull) {
        idx_out[(size_t)qi * k + i] = __float_as_int(c.spts[(unsigned)(key & 0xffffffffull)].w);
        d2_out[(size_t)qi * k + i] = __uint_as_float((unsigned)(key >> 32));
      } else {
        idx_out[(size_t)qi * k + i] = -1;
        d2_out[(size_t)qi * k + i] = INFINITY;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host orchestration
// Cell size along y and z.  Rows of cells are swept along x from the query outwards (knn.cuh), so the y/z size trades
// the number of rows a query touches against the width of the strip it sweeps; ~2 mean point spacings works best.
static float auto_cell_size(const Cloud& c, const b2r_config& cfg) {
  if (cfg.nn_cell_size > 0) return (float)cfg.nn_cell_size;
  double dx = std::max(1e-3f, c.bmax[0] - c.bmin[0]), dy = std::max(1e-3f, c.bmax[1] - c.bmin[1]);
  static const double factor = [] { const char* e = getenv("B2R_NN_CELL_FACTOR"); return e ? atof(e) : 2.5; }();
  double h = factor * std::sqrt(dx * dy / std::max(1, c.n));
  return (float)std::min(4.0, std::max(0.05, h));
}

static void grid_dims(Cloud& c, float h) {
  // cells along x only index into the sorted rows: smaller ones shorten the part of a row scanned unconditionally
  static const float xfactor = [] { const char* e = getenv("B2R_NN_XCELL_FACTOR"); return e ? (float)atof(e) : 0.34f; }();
  const long cap = 1l << 24;
  for (;;) {
    const float hx = h * xfactor;
    long tot = 1;
    for (int d = 0; d < 3; ++d) {
      c.gd[d] = (int)std::floor((c.bmax[d] - c.bmin[d]) / (d == 0 ? hx : h)) + 1;
      tot *= c.gd[d];
    }
    if (tot <= cap) { c.ncell = (int)tot; c.h = h; c.hx = hx; return; }
    h *= 1.26f;
  }
}

static inline unsigned blocks_for(int n, int per_block, int cap) { return (unsigned)std::max(1, std::min(cap, (n + per_block - 1) / per_block)); }

// kNN covariances of a batch of clouds: one launch, one thread per point, clouds on grid.y
static void launch_knn_cov(Ctx& ctx, const CloudView* dviews, const std::vector<Cloud*>& clouds, int k, int maxn, int32_t* knn_out) {
  const int nc = (int)clouds.size();
  double bytes = 0.0;  // 16 B point in + 24 B covariance out per point
  for (Cloud* c : clouds) bytes += 40.0 * c->n;
  const dim3 grid((unsigned)((maxn + 127) / 128), (unsigned)nc);
  ProfScope ps(ctx, PROF_KNN_COV, bytes);
  // B2R_KNN_TILE=1: the variant with the rows staged in shared memory by TMA bulk copies.  Measured on the 4096-pair batch
  // (272 clouds of 22.6k points) 5.69 ms against 3.32 ms for knn_cov_kernel, on the 33-cloud chain 1.61 against 0.96 ms: the
  // search is bound by instruction issue, not by load latency, and the tile adds instructions (range test per load, a second
  // sweep instead of the candidate log, a serial prologue) while leaving room for 5 instead of 8 blocks per SM.  Kept as a
  // measured alternative (profiles/r2/ncu_knn_cov_tile_*.md, SASS with UBLKCP in profiles/sass/); the default stays the log kernel.
  static const bool tile = [] { const char* e = getenv("B2R_KNN_TILE"); return e && atoi(e) != 0; }();
  if (tile) {
    if (k <= 8) B2R_LAUNCH(ctx, knn_cov_tile_kernel<8>, grid, 128, 0, dviews, k, knn_out);
    else if (k <= 16) B2R_LAUNCH(ctx, knn_cov_tile_kernel<16>, grid, 128, 0, dviews, k, knn_out);
    else if (k <= 20) B2R_LAUNCH(ctx, knn_cov_tile_kernel<20>, grid, 128, 0, dviews, k, knn_out);
    else if (k <= 24) B2R_LAUNCH(ctx, knn_cov_tile_kernel<24>, grid, 128, 0, dviews, k, knn_out);
    else B2R_LAUNCH(ctx, knn_cov_tile_kernel<32>, grid, 128, 0, dviews, k, knn_out);
    return;
  }
  if (k <= 8) B2R_LAUNCH(ctx, knn_cov_kernel<8>, grid, 128, 0, dviews, k, knn_out);
  else if (k <= 16) B2R_LAUNCH(ctx, knn_cov_kernel<16>, grid, 128, 0, dviews, k, knn_out);
  else if (k <= 20) B2R_LAUNCH(ctx, knn_cov_kernel<20>, grid, 128, 0, dviews, k, knn_out);
  else if (k <= 24) B2R_LAUNCH(ctx, knn_cov_kernel<24>, grid, 128, 0, dviews, k, knn_out);
  else B2R_LAUNCH(ctx, knn_cov_kernel<32>, grid, 128, 0, dviews, k, knn_out);
}

// bounding boxes of the clouds that lack one: one kernel over all of them, one D2H, one synchronisation.
// device_srcs (optional, one per cloud, same order): packed device buffers the points are copied from in the same pass.
// a cloud whose box is still on its way (split upload): wait for its group and take the box
static void cloud_resolve_box(Cloud& c) {
  if (!c.pending) return;
  c.pending->resolve();
  const int* v = c.pending->vals.data() + (size_t)6 * c.pending_idx;
  for (int d = 0; d < 3; ++d) {
    c.bmin[d] = c.n ? ordered_to_float(v[d]) : 0.f;
    c.bmax[d] = c.n ? ordered_to_float(v[3 + d]) : 0.f;
  }
  c.has_bbox = true;
  c.pending.reset();
}

void clouds_compute_bbox(Ctx& ctx, const std::vector<Cloud*>& clouds, const void* const* device_srcs = nullptr) {
  for (Cloud* c : clouds) cloud_resolve_box(*c);
  std::vector<Cloud*> todo;
  std::vector<const float4*> srcs;
  for (size_t i = 0; i < clouds.size(); ++i) {
    Cloud* c = clouds[i];
    if (!c->has_bbox && std::find(todo.begin(), todo.end(), c) == todo.end()) {
      todo.push_back(c);
      if (device_srcs) srcs.push_back((const float4*)device_srcs[i]);
    }
  }
  if (todo.empty()) return;
  const int nc = (int)todo.size();
  DBuf<const float4*> dsrc;
  if (device_srcs) {
    dsrc.alloc(nc, ctx.stream);
    B2R_CUDA(cudaMemcpyAsync(dsrc.p, srcs.data(), sizeof(const float4*) * nc, cudaMemcpyHostToDevice, ctx.stream));
  }
  std::vector<CloudView> hv(nc);
  int maxn = 1;
  for (int i = 0; i < nc; ++i) { hv[i] = todo[i]->view(); maxn = std::max(maxn, todo[i]->n); }
  DBuf<CloudView> dv; dv.alloc(nc, ctx.stream);
  DBuf<int> db; db.alloc((size_t)nc * 6, ctx.stream);
  B2R_CUDA(cudaMemcpyAsync(dv.p, hv.data(), sizeof(CloudView) * nc, cudaMemcpyHostToDevice, ctx.stream));
  B2R_LAUNCH(ctx, bbox_init_kernel, (nc * 6 + 255) / 256, 256, 0, db.p, nc);
  dim3 g(blocks_for(maxn, 256 * 8, std::max(1, 4 * ctx.num_sms / nc)), nc);
  if (device_srcs) B2R_LAUNCH(ctx, bbox_kernel<true>, g, 256, 0, dv.p, db.p, dsrc.p);
  else B2R_LAUNCH(ctx, bbox_kernel<false>, g, 256, 0, dv.p, db.p, (const float4* const*)nullptr);
  std::vector<int> hb((size_t)nc * 6);
  B2R_CUDA(cudaMemcpyAsync(hb.data(), db.p, sizeof(int) * nc * 6, cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  for (int i = 0; i < nc; ++i) {
    Cloud* c = todo[i];
    for (int d = 0; d < 3; ++d) {
      c->bmin[d] = c->n ? ordered_to_float(hb[i * 6 + d]) : 0.f;
      c->bmax[d] = c->n ? ordered_to_float(hb[i * 6 + 3 + d]) : 0.f;
    }
    c->has_bbox = true;
  }
}

void clouds_upload(Ctx& ctx, Cloud* const* clouds, const void* const* points, const size_t* n, size_t count, size_t stride_bytes, int memspace) {
  if (stride_bytes != 16 && stride_bytes != 32) throw Error(B2R_ERR_INVALID_ARG, "stride_bytes must be 16 or 32");
  ArenaPlan plan;
  std::vector<uint8_t*> raw(count, nullptr);
  const bool stage = stride_bytes == 32 && memspace != B2R_DEVICE;
  for (size_t i = 0; i < count; ++i) {
    if (n[i] > (size_t)INT_MAX / 8) throw Error(B2R_ERR_INVALID_ARG, "cloud too large");
    Cloud& c = *clouds[i];
    c = Cloud();
    c.device = ctx.device;
    c.n = (int)n[i];
    plan.want(c.pts.p, n[i]);
  }
  std::shared_ptr<Arena> arena = plan.commit(ctx);
  DBuf<uint8_t> staging;
  if (stage) {
    size_t tot = 0;
    for (size_t i = 0; i < count; ++i) tot += ArenaPlan::up(n[i] * 32);
    staging.alloc(tot, ctx.stream);
    size_t off = 0;
    for (size_t i = 0; i < count; ++i) { raw[i] = staging.p + off; off += ArenaPlan::up(n[i] * 32); }
  }
  const cudaMemcpyKind kind = memspace == B2R_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  // Copy inside the bounding-box kernel: device buffers, or PINNED host buffers (device-readable under unified addressing —
  // the kernel then pulls them over PCIe itself: one launch instead of one DMA call per cloud).  Pageable host memory
  // keeps the cudaMemcpyAsync path.
  bool fused_copy = stride_bytes == 16 && count > 1;
  if (fused_copy && memspace != B2R_DEVICE) {
    static const bool zero_copy = [] { const char* e = getenv("B2R_ZERO_COPY_UPLOAD"); return !e || atoi(e) != 0; }();
    fused_copy = zero_copy;
    for (size_t i = 0; i < count && fused_copy; ++i) {
      if (n[i] == 0) continue;
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, points[i]) != cudaSuccess) { cudaGetLastError(); fused_copy = false; break; }
      if (at.type != cudaMemoryTypeHost || at.devicePointer != points[i]) fused_copy = false;
    }
  }
  std::vector<Cloud*> all;
  for (size_t i = 0; i < count; ++i) {
    Cloud& c = *clouds[i];
    c.mem_pts = arena;
    all.push_back(&c);
    if (n[i] == 0 || fused_copy) continue;
    if (stride_bytes == 16) {
      B2R_CUDA(cudaMemcpyAsync(c.pts.p, points[i], n[i] * 16, kind, ctx.stream));
    } else {
      const uint8_t* src = (const uint8_t*)points[i];
      if (stage) {
        B2R_CUDA(cudaMemcpyAsync(raw[i], points[i], n[i] * 32, kind, ctx.stream));
        src = raw[i];
      }
      B2R_LAUNCH(ctx, repack32_kernel, (unsigned)((n[i] + 255) / 256), 256, 0, src, (int)n[i], c.pts.p);
    }
  }
  // Pinned host clouds, large batch: the batch is copied in two halves.  The first half goes as before (kernel on the handle's
  // stream, boxes read back, one synchronisation); the second half's copy + box kernel runs on a second stream and is NOT waited
  // for: its clouds carry a PendingBoxes, and the batch path prepares the first half's structures (grid, covariances, voxel maps)
  // while the second half is still coming over PCIe (api.cu: run_align).  B2R_SPLIT_UPLOAD=0 restores the single kernel.
  static const bool split_on = [] { const char* e = getenv("B2R_SPLIT_UPLOAD"); return !e || atoi(e) != 0; }();
  if (!(fused_copy && memspace != B2R_DEVICE && split_on && count >= 16)) {
    clouds_compute_bbox(ctx, all, fused_copy ? points : nullptr);  // synchronises
    return;
  }
  if (auto old = ctx.pending_boxes.lock()) old->resolve();  // the pinned area is about to be reused
  const int nc = (int)count;
  size_t tot_pts = 0, acc_pts = 0;
  for (size_t i = 0; i < count; ++i) tot_pts += n[i];
  int k = 0;  // first cloud of the second half
  while (k < nc - 1 && acc_pts + n[k] <= tot_pts / 2) acc_pts += n[k++];
  k = std::max(k, 1);
  if (!ctx.copy_stream) B2R_CUDA(cudaStreamCreateWithFlags(&ctx.copy_stream, cudaStreamNonBlocking));
  if (ctx.pinned_boxes_cap < (size_t)nc * 6) {
    if (ctx.pinned_boxes) cudaFreeHost(ctx.pinned_boxes);
    ctx.pinned_boxes = nullptr;
    ctx.pinned_boxes_cap = 0;
    B2R_CUDA(cudaHostAlloc((void**)&ctx.pinned_boxes, sizeof(int) * 6 * (size_t)nc * 2, cudaHostAllocDefault));
    ctx.pinned_boxes_cap = (size_t)nc * 6 * 2;
  }
  std::vector<CloudView> hv(nc);
  std::vector<const float4*> srcs(nc);
  int maxn_a = 1, maxn_b = 1;
  for (int i = 0; i < nc; ++i) {
    hv[i] = all[i]->view();
    srcs[i] = (const float4*)points[i];
    (i < k ? maxn_a : maxn_b) = std::max(i < k ? maxn_a : maxn_b, all[i]->n);
  }
  DBuf<CloudView> dv; dv.alloc(nc, ctx.stream);
  DBuf<const float4*> dsrc; dsrc.alloc(nc, ctx.stream);
  DBuf<int> db; db.alloc((size_t)nc * 6, ctx.stream);
  B2R_CUDA(cudaMemcpyAsync(dv.p, hv.data(), sizeof(CloudView) * nc, cudaMemcpyHostToDevice, ctx.stream));
  B2R_CUDA(cudaMemcpyAsync(dsrc.p, srcs.data(), sizeof(const float4*) * nc, cudaMemcpyHostToDevice, ctx.stream));
  B2R_LAUNCH(ctx, bbox_init_kernel, (nc * 6 + 255) / 256, 256, 0, db.p, nc);
  auto pb = std::make_shared<PendingBoxes>();
  pb->device = ctx.device;
  pb->count = nc - k;
  pb->pinned = ctx.pinned_boxes + (size_t)6 * k;
  // ---- first half on the handle's stream
  {
    dim3 g(blocks_for(maxn_a, 256 * 8, std::max(1, 4 * ctx.num_sms / k)), k);
    B2R_LAUNCH(ctx, bbox_kernel<true>, g, 256, 0, dv.p, db.p, dsrc.p);
  }
  // ---- second half on the copy stream, behind the first half's kernel (the two would only share the PCIe link otherwise, and
  // the first half would arrive later); not waited for here
  cudaEvent_t first_done = ctx.get_event();
  B2R_CUDA(cudaEventRecord(first_done, ctx.stream));
  B2R_CUDA(cudaStreamWaitEvent(ctx.copy_stream, first_done, 0));
  ctx.ev_pool.push_back(first_done);
  {
    dim3 g(blocks_for(maxn_b, 256 * 8, std::max(1, 4 * ctx.num_sms / (nc - k))), nc - k);
    bbox_kernel<true><<<g, 256, 0, ctx.copy_stream>>>(dv.p + k, db.p + (size_t)6 * k, dsrc.p + k);
    ++ctx.launches;
    B2R_CUDA(cudaGetLastError());
    B2R_CUDA(cudaMemcpyAsync(ctx.pinned_boxes + (size_t)6 * k, db.p + (size_t)6 * k, sizeof(int) * 6 * (nc - k), cudaMemcpyDeviceToHost, ctx.copy_stream));
    B2R_CUDA(cudaEventCreateWithFlags(&pb->ev, cudaEventDisableTiming));
    B2R_CUDA(cudaEventRecord(pb->ev, ctx.copy_stream));
  }
  // ---- the first half's boxes: one synchronisation of the handle's stream
  {
    B2R_CUDA(cudaMemcpyAsync(ctx.pinned_boxes, db.p, sizeof(int) * 6 * k, cudaMemcpyDeviceToHost, ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(ctx.stream));
    for (int i = 0; i < k; ++i) {
      Cloud* c = all[i];
      for (int d = 0; d < 3; ++d) {
        c->bmin[d] = c->n ? ordered_to_float(ctx.pinned_boxes[i * 6 + d]) : 0.f;
        c->bmax[d] = c->n ? ordered_to_float(ctx.pinned_boxes[i * 6 + 3 + d]) : 0.f;
      }
      c->has_bbox = true;
    }
  }
  // the kernel on the copy stream still reads dv / dsrc and writes db: they are released when its event has passed
  pb->scratch.push_back({dv.p, ctx.stream}); dv.p = nullptr;
  pb->scratch.push_back({dsrc.p, ctx.stream}); dsrc.p = nullptr;
  pb->scratch.push_back({db.p, ctx.stream}); db.p = nullptr;
  for (int i = k; i < nc; ++i) { all[i]->pending = pb; all[i]->pending_idx = i - k; }
  ctx.pending_boxes = pb;
}

template <int MODE>
static void run_grid_build(Ctx& ctx, const CloudView* dviews, int nc, int maxn, int maxcell, unsigned long long* tile_tot, int max_tiles) {
  dim3 g(blocks_for(maxn, 256 * 4, std::max(1, 8 * ctx.num_sms / nc)), nc);
  ProfScope ps(ctx, PROF_GRID_BUILD, 0.0);
  B2R_LAUNCH(ctx, grid_count_kernel<MODE>, g, 256, 0, dviews);
  B2R_LAUNCH(ctx, grid_scan_tile_kernel<MODE>, dim3((maxcell + kScanTile - 1) / kScanTile, nc), 1024, 0, dviews, tile_tot, max_tiles);
  B2R_LAUNCH(ctx, grid_scan_tops_kernel<MODE>, nc, 1024, 0, dviews, tile_tot, max_tiles);
  B2R_LAUNCH(ctx, grid_scan_apply_kernel<MODE>, dim3((maxcell + kScanTile - 1) / kScanTile, nc), 1024, 0, dviews, tile_tot, max_tiles);
  B2R_LAUNCH(ctx, grid_scatter_kernel<MODE>, g, 256, 0, dviews);
  B2R_LAUNCH(ctx, grid_rank_kernel<MODE>, dim3(blocks_for(maxn, 128, std::max(1, 16 * ctx.num_sms / nc)), nc), 128, 0, dviews);
}

void clouds_prepare(Ctx& ctx, const b2r_config& cfg, const std::vector<Cloud*>& clouds_in, const std::vector<Needs>& needs_in,
                    DBuf<CloudView>& dviews) {
  // ---- merge duplicate clouds
  std::vector<Cloud*> clouds;
  std::vector<Needs> needs;
  std::vector<int> slot_of(clouds_in.size());
  for (size_t i = 0; i < clouds_in.size(); ++i) {
    Cloud* c = clouds_in[i];
    if (c->device != ctx.device) throw Error(B2R_ERR_INVALID_ARG, "cloud lives on another device");
    size_t j = 0;
    for (; j < clouds.size(); ++j)
      if (clouds[j] == c) break;
    if (j == clouds.size()) { clouds.push_back(c); needs.push_back(Needs()); }
    slot_of[i] = (int)j;
    Needs& nd = needs[j];
    const Needs& in = needs_in[i];
    nd.grid = nd.grid || in.grid || in.cov_k > 0;
    nd.cov_k = std::max(nd.cov_k, in.cov_k);
    nd.cov_mode = std::max(nd.cov_mode, in.cov_mode);
    if (in.vres > 0) nd.vres = in.vres;
    if (in.leaf > 0) nd.leaf = in.leaf;
    nd.leaf_centroids = nd.leaf_centroids || in.leaf_centroids;
    if (nd.vres > 0 && nd.cov_k == 0) throw Error(B2R_ERR_STATE, "voxel map needs covariances");
  }
  clouds_compute_bbox(ctx, clouds);  // the only synchronisation (skipped when the clouds came through clouds_upload)

  // ---- plan: dimensions of everything that is missing, and ONE allocation for all of it
  std::vector<int> todo_grid, todo_cov, todo_vox, todo_ndt;
  ArenaPlan plan;
  int cov_k = 0, cov_mode = 0;
  for (size_t i = 0; i < clouds.size(); ++i) {
    Cloud* c = clouds[i];
    const Needs& nd = needs[i];
    if (c->n == 0) continue;
    if (nd.grid && !c->has_grid) {
      grid_dims(*c, auto_cell_size(*c, cfg));
      plan.want(c->cell_start.p, (size_t)c->ncell + 1);
      plan.want_zeroed(c->cell_cnt.p, (size_t)c->ncell);
      plan.want(c->cell_tmp.p, (size_t)c->n);
      plan.want(c->spts.p, (size_t)c->n);
      plan.want(c->spair.p, ((size_t)c->n + 1) / 2 * 2);
      todo_grid.push_back((int)i);
    }
    const bool new_cov = nd.cov_k > 0 && (c->cov_k != nd.cov_k || c->cov_mode != nd.cov_mode);
    if (new_cov) {
      if (nd.cov_k > 32) throw Error(B2R_ERR_INVALID_ARG, "correspondence_randomness must be <= 32");
      if (c->n < nd.cov_k) throw Error(B2R_ERR_INVALID_ARG, "cloud has fewer points than correspondence_randomness");
      if (cov_k && cov_k != nd.cov_k) throw Error(B2R_ERR_INVALID_ARG, "one correspondence_randomness per call");
      if (cov_k && cov_mode != nd.cov_mode) throw Error(B2R_ERR_INVALID_ARG, "one covariance mode per call");
      cov_k = nd.cov_k;
      cov_mode = nd.cov_mode;
      plan.want(c->cov.p, (size_t)c->n * 6);
      if (nd.cov_mode == 0) plan.want(c->nrm.p, (size_t)c->n * 4); else c->nrm.p = nullptr;
      c->vres = 0.0;  // a voxel map built from older covariances is stale
      todo_cov.push_back((int)i);
    }
    if (nd.vres > 0 && c->vres != nd.vres) {
      const double res = nd.vres;
      long tot = 1;
      for (int d = 0; d < 3; ++d) {
        c->vmin[d] = (int)std::floor((double)c->bmin[d] / res - 0.5);
        int vmax = (int)std::floor((double)c->bmax[d] / res - 0.5);
        c->vd[d] = vmax - c->vmin[d] + 1;
        tot *= c->vd[d];
      }
      if (tot > (1l << 26)) throw Error(B2R_ERR_CAPACITY, "VGICP voxel table too large for this extent/resolution");
      c->vcell = (int)tot;
      const size_t maxrec = (size_t)std::min<long>(tot, c->n);
      plan.want(c->v_start.p, (size_t)tot + 1);
      plan.want_zeroed(c->v_cnt.p, (size_t)tot);
      plan.want(c->v_table.p, (size_t)tot);
      plan.want(c->v_order.p, (size_t)c->n * 2);
      plan.want(c->v_reccell.p, maxrec);
      plan.want(c->vrec.p, maxrec);
      plan.want(c->v_nrec.p, 1);
      todo_vox.push_back((int)i);
    }
    if (nd.leaf > 0 && (c->leaf != nd.leaf || (nd.leaf_centroids && !c->ndt_centroids))) {
      c->ndt_centroids = nd.leaf_centroids;
      const float leaf = nd.leaf, inv_leaf = 1.0f / leaf;
      int64_t dx = (int64_t)((c->bmax[0] - c->bmin[0]) * inv_leaf) + 1, dy = (int64_t)((c->bmax[1] - c->bmin[1]) * inv_leaf) + 1,
              dz = (int64_t)((c->bmax[2] - c->bmin[2]) * inv_leaf) + 1;
      const bool overflow = dx * dy * dz > (int64_t)INT32_MAX;  // PCL: warn, grid stays empty
      long tot = 1;
      for (int d = 0; d < 3; ++d) {
        c->min_b[d] = (int)std::floor(c->bmin[d] * inv_leaf);
        c->max_b[d] = (int)std::floor(c->bmax[d] * inv_leaf);
        c->div_b[d] = c->max_b[d] - c->min_b[d] + 1;
        tot *= c->div_b[d];
      }
      c->ndt_overflow = overflow;
      if (overflow) { c->ncell_ndt = 0; c->leaf = leaf; continue; }
      if (tot > (1l << 26)) throw Error(B2R_ERR_CAPACITY, "NDT voxel table too large for this extent/resolution");
      c->ncell_ndt = (int)tot;
      const size_t maxrec = (size_t)std::min<long>(tot, c->n);
      plan.want(c->n_start.p, (size_t)tot + 1);
      plan.want_zeroed(c->n_cnt.p, (size_t)tot);
      plan.want(c->n_table.p, (size_t)tot);
      plan.want(c->n_order.p, (size_t)c->n * 2);
      plan.want(c->n_reccell.p, maxrec);
      plan.want(c->nrec.p, maxrec);
      plan.want(c->n_nrec.p, 1);
      todo_ndt.push_back((int)i);
    }
  }
  if (!plan.empty()) {
    std::shared_ptr<Arena> arena = plan.commit(ctx);
    for (int i : todo_grid) clouds[i]->mem_grid = arena;
    for (int i : todo_cov) clouds[i]->mem_cov = arena;
    for (int i : todo_vox) clouds[i]->mem_vox = arena;
    for (int i : todo_ndt) clouds[i]->mem_ndt = arena;
  }
  // the flags go up now: the views below must describe the finished structures
  for (int i : todo_grid) clouds[i]->has_grid = true;
  for (int i : todo_cov) { clouds[i]->cov_k = cov_k; clouds[i]->cov_mode = cov_mode; }
  for (int i : todo_vox) clouds[i]->vres = needs[i].vres;
  for (int i : todo_ndt) clouds[i]->leaf = needs[i].leaf;

  // ---- one upload: [views in caller order | grid todo | cov todo | voxel todo | ndt todo]
  std::vector<CloudView> hv;
  hv.reserve(clouds_in.size() + todo_grid.size() + todo_cov.size() + todo_vox.size() + todo_ndt.size());
  for (size_t i = 0; i < clouds_in.size(); ++i) hv.push_back(clouds[slot_of[i]]->view());
  size_t off_grid = hv.size();
  for (int i : todo_grid) hv.push_back(clouds[i]->view());
  size_t off_cov = hv.size();
  for (int i : todo_cov) hv.push_back(clouds[i]->view());
  size_t off_vox = hv.size();
  for (int i : todo_vox) hv.push_back(clouds[i]->view());
  size_t off_ndt = hv.size();
  for (int i : todo_ndt) hv.push_back(clouds[i]->view());
  dviews.alloc(std::max<size_t>(1, hv.size()), ctx.stream);
  if (!hv.empty()) B2R_CUDA(cudaMemcpyAsync(dviews.p, hv.data(), sizeof(CloudView) * hv.size(), cudaMemcpyHostToDevice, ctx.stream));

  // scratch for the scans: the largest (clouds x tiles) of the three grid kinds
  auto stage_dims = [&](const std::vector<int>& todo, int kind, int& maxn, int& maxcell) {
    maxn = 1; maxcell = 1;
    for (int i : todo) {
      maxn = std::max(maxn, clouds[i]->n);
      maxcell = std::max(maxcell, kind == 0 ? clouds[i]->ncell : (kind == 1 ? clouds[i]->vcell : clouds[i]->ncell_ndt));
    }
  };
  size_t tile_need = 0;
  int mn[3], mc[3];
  const std::vector<int>* todos[3] = {&todo_grid, &todo_vox, &todo_ndt};
  for (int kd = 0; kd < 3; ++kd) {
    stage_dims(*todos[kd], kd, mn[kd], mc[kd]);
    tile_need = std::max(tile_need, todos[kd]->size() * (size_t)((mc[kd] + kScanTile - 1) / kScanTile));
  }
  DBuf<unsigned long long> tile_tot;
  if (tile_need) tile_tot.alloc(tile_need, ctx.stream);

  if (!todo_grid.empty())
    run_grid_build<GRID_NN>(ctx, dviews.p + off_grid, (int)todo_grid.size(), mn[0], mc[0], tile_tot.p, (mc[0] + kScanTile - 1) / kScanTile);
  if (!todo_cov.empty()) {
    std::vector<Cloud*> cl;
    int maxn = 1;
    for (int i : todo_cov) { cl.push_back(clouds[i]); maxn = std::max(maxn, clouds[i]->n); }
    if (cov_mode == 0) {
      launch_knn_cov(ctx, dviews.p + off_cov, cl, cov_k, maxn, nullptr);
    } else {
      // pcl::GeneralizedIterativeClosestPoint covariances: the same exact kNN, cloud by cloud with the neighbour lists written
      // out, then PCL's own moment arithmetic over those lists
      DBuf<int32_t> knn; knn.alloc((size_t)maxn * cov_k, ctx.stream);
      for (size_t j = 0; j < cl.size(); ++j) {
        std::vector<Cloud*> one{cl[j]};
        launch_knn_cov(ctx, dviews.p + off_cov + j, one, cov_k, cl[j]->n, knn.p);
        B2R_LAUNCH(ctx, pcl_cov_kernel, (cl[j]->n + 127) / 128, 128, 0, dviews.p + off_cov + j, knn.p, cov_k, cfg.gicp_epsilon);
      }
    }
  }
  if (!todo_vox.empty()) {
    const int nc = (int)todo_vox.size();
    run_grid_build<GRID_VGICP>(ctx, dviews.p + off_vox, nc, mn[1], mc[1], tile_tot.p, (mc[1] + kScanTile - 1) / kScanTile);
    dim3 g(blocks_for(std::min(mc[1], mn[1]), 8, std::max(1, 8 * ctx.num_sms / nc)), (unsigned)nc);
    ProfScope ps(ctx, PROF_VOXEL_REDUCE, 0.0);
    B2R_LAUNCH(ctx, vgicp_reduce_kernel, g, 256, 0, dviews.p + off_vox);
  }
  if (!todo_ndt.empty()) {
    const int nc = (int)todo_ndt.size();
    run_grid_build<GRID_NDT>(ctx, dviews.p + off_ndt, nc, mn[2], mc[2], tile_tot.p, (mc[2] + kScanTile - 1) / kScanTile);
    dim3 g(blocks_for(std::min(mc[2], mn[2]), 8, std::max(1, 8 * ctx.num_sms / nc)), (unsigned)nc);
    ProfScope ps(ctx, PROF_VOXEL_REDUCE, 0.0);
    B2R_LAUNCH(ctx, ndt_reduce_kernel, g, 256, 0, dviews.p + off_ndt);
  }
}

void debug_knn(Ctx& ctx, const b2r_config& cfg, Cloud& c, const float* queries, size_t nq, int k, int32_t* idx_out, float* d2_out) {
  if (k < 1 || k > 32) throw Error(B2R_ERR_INVALID_ARG, "k must be in [1,32]");
  std::vector<Cloud*> cl{&c};
  std::vector<Needs> nd(1);
  nd[0].grid = true;
  DBuf<CloudView> dv;
  clouds_prepare(ctx, cfg, cl, nd, dv);
  DBuf<float4> dq; dq.alloc(nq, ctx.stream);
  DBuf<int32_t> di; di.alloc(nq * k, ctx.stream);
  DBuf<float> dd; dd.alloc(nq * k, ctx.stream);
  B2R_CUDA(cudaMemcpyAsync(dq.p, queries, nq * 16, cudaMemcpyHostToDevice, ctx.stream));
  B2R_LAUNCH(ctx, knn_query_kernel, (unsigned)((nq + 63) / 64), 64, 0, dv.p, dq.p, (int)nq, k, di.p, dd.p);
  B2R_CUDA(cudaMemcpyAsync(idx_out, di.p, nq * k * 4, cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaMemcpyAsync(d2_out, dd.p, nq * k * 4, cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
}

unsigned long long debug_knn_list_overflows(Ctx& ctx) {
  unsigned long long v = 0;
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  B2R_CUDA(cudaMemcpyFromSymbol(&v, g_knn_list_overflows, sizeof(v)));
#ifdef B2R_KNN_STATS
  unsigned long long hst[66];
  B2R_CUDA(cudaMemcpyFromSymbol(hst, g_knn_hist, sizeof(hst)));
  unsigned long long tot = 0;
  for (int i = 0; i < 64; ++i) tot += hst[i];
  fprintf(stderr, "knn stats: queries %llu, tested/query %.1f, mean ring %.2f\nlogged histogram (bucket of 4):", tot, (double)hst[64] / tot, (double)hst[65] / tot);
  for (int i = 0; i < 64; ++i) fprintf(stderr, " %d:%.4f", i * 4, (double)hst[i] / tot);
  fprintf(stderr, "\n");
#endif
  return v;
}

void debug_cov_knn(Ctx& ctx, const b2r_config& cfg, Cloud& c, int k, int32_t* knn_out) {
  std::vector<Cloud*> cl{&c};
  std::vector<Needs> nd(1);
  nd[0].grid = true;
  DBuf<CloudView> dv0;
  clouds_prepare(ctx, cfg, cl, nd, dv0);
  if (!c.cov.p || c.cov_k == 0) {
    ArenaPlan plan;
    plan.want(c.cov.p, (size_t)c.n * 6);
    plan.want(c.nrm.p, (size_t)c.n * 4);
    c.mem_cov = plan.commit(ctx);
  }
  CloudView hv = c.view();
  DBuf<CloudView> dv; dv.alloc(1, ctx.stream);
  B2R_CUDA(cudaMemcpyAsync(dv.p, &hv, sizeof(hv), cudaMemcpyHostToDevice, ctx.stream));
  DBuf<int32_t> dk; dk.alloc((size_t)c.n * k, ctx.stream);
  std::vector<Cloud*> one{&c};
  launch_knn_cov(ctx, dv.p, one, k, c.n, dk.p);
  B2R_CUDA(cudaMemcpyAsync(knn_out, dk.p, (size_t)c.n * k * 4, cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  c.cov_k = k;
  c.vres = 0.0;
}

// ------------------------------------------------------------------------------------------------ compaction
// order-preserving stream compaction: block counts -> single-block scan -> scatter
__global__ void compact_count_kernel(const uint8_t* __restrict__ keep, int n, int* __restrict__ block_cnt) {
  __shared__ int wsum[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int v = (i < n && keep[i]) ? 1 : 0;
  int s = warp_sum(v);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0;
    t = warp_sum(t);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) compact_scan_kernel(int* __restrict__ block_cnt, int nblocks, int* __restrict__ total) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? block_cnt[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    const int excl = carry + warp_tot[warp] + incl - v;
    if (i < nblocks) block_cnt[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void compact_scatter_kernel(const float4* __restrict__ in, const uint8_t* __restrict__ keep, int n,
                                       const int* __restrict__ block_off, float4* __restrict__ out) {
  __shared__ int wsum[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = (i < n && keep[i]) ? 1 : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, v);
  if (lane == 0) wsum[warp] = __popc(bal);
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += wsum[w];
  if (v) out[block_off[blockIdx.x] + woff + __popc(bal & ((1u << lane) - 1u))] = in[i];
}

void compact_points(Ctx& ctx, const float4* in, const uint8_t* keep, int n, DevCloud& out) {
  out.n = 0;
  if (n == 0) return;
  const int nb = (n + 255) / 256;
  DBuf<int> cnt; cnt.alloc((size_t)nb + 1, ctx.stream);
  B2R_LAUNCH(ctx, compact_count_kernel, nb, 256, 0, keep, n, cnt.p);
  B2R_LAUNCH(ctx, compact_scan_kernel, 1, 1024, 0, cnt.p, nb, cnt.p + nb);
  out.pts.alloc((size_t)n, ctx.stream);
  B2R_LAUNCH(ctx, compact_scatter_kernel, nb, 256, 0, in, keep, n, cnt.p, out.pts.p);
  int total = 0;
  B2R_CUDA(cudaMemcpyAsync(&total, cnt.p + nb, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  out.n = total;
}

}  // namespace b2r

namespace b2r {

// block_off: nb + 1 ints (nb = ceil(n/256)): exclusive per-block offsets of the set flags, then the total
void flags_block_offsets(Ctx& ctx, const uint8_t* flags, int n, int* block_off) {
  const int nb = (n + 255) / 256;
  B2R_LAUNCH(ctx, compact_count_kernel, nb, 256, 0, flags, n, block_off);
  B2R_LAUNCH(ctx, compact_scan_kernel, 1, 1024, 0, block_off, nb, block_off + nb);
}

void compute_bbox(Ctx& ctx, const float4* pts, int n, float mn[3], float mx[3]) {
  CloudView hv;
  memset(&hv, 0, sizeof(hv));
  hv.pts = pts;
  hv.n = n;
  DBuf<CloudView> dv; dv.alloc(1, ctx.stream);
  DBuf<int> db; db.alloc(6, ctx.stream);
  B2R_CUDA(cudaMemcpyAsync(dv.p, &hv, sizeof(hv), cudaMemcpyHostToDevice, ctx.stream));
  B2R_LAUNCH(ctx, bbox_init_kernel, 1, 32, 0, db.p, 1);
  B2R_LAUNCH(ctx, bbox_kernel<false>, dim3(blocks_for(n, 256 * 8, 2 * ctx.num_sms), 1), 256, 0, dv.p, db.p, (const float4* const*)nullptr);
  int hb[6];
  B2R_CUDA(cudaMemcpyAsync(hb, db.p, sizeof(hb), cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  for (int d = 0; d < 3; ++d) { mn[d] = ordered_to_float(hb[d]); mx[d] = ordered_to_float(hb[3 + d]); }
}

// getMinMax3D over the points the distance filter keeps (near < |p| < far, same float norm as distance_flag_kernel): lets
// b2r_prefilter fold the distance filter into VoxelGrid without compacting the cloud in between.  count = points kept.
__global__ void bbox_range_kernel(const float4* __restrict__ pts, int n, double near_t, double far_t, int* __restrict__ bbox_out,
                                  int* __restrict__ count_out) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(&pts[i]);
    const float s = __fadd_rn(__fmul_rn(p.x, p.x), __fadd_rn(__fmul_rn(p.y, p.y), __fmul_rn(p.z, p.z)));
    const double d = (double)__fsqrt_rn(s);
    if (d > near_t && d < far_t) {  // false for non-finite points
      ++cnt;
      mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      atomicMin(&bbox_out[d], float_to_ordered(mn[d]));
      atomicMax(&bbox_out[3 + d], float_to_ordered(mx[d]));
    }
    if (cnt) atomicAdd(count_out, cnt);
  }
}

void compute_bbox_range(Ctx& ctx, const float4* pts, int n, double near_t, double far_t, float mn[3], float mx[3], int* count) {
  DBuf<int> db; db.alloc(7, ctx.stream);
  B2R_LAUNCH(ctx, bbox_init_kernel, 1, 32, 0, db.p, 1);
  B2R_CUDA(cudaMemsetAsync(db.p + 6, 0, sizeof(int), ctx.stream));
  B2R_LAUNCH(ctx, bbox_range_kernel, blocks_for(n, 256 * 8, 2 * ctx.num_sms), 256, 0, pts, n, near_t, far_t, db.p, db.p + 6);
  int hb[7];
  B2R_CUDA(cudaMemcpyAsync(hb, db.p, sizeof(hb), cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  for (int d = 0; d < 3; ++d) { mn[d] = ordered_to_float(hb[d]); mx[d] = ordered_to_float(hb[3 + d]); }
  *count = hb[6];
}

}  // namespace b2r
