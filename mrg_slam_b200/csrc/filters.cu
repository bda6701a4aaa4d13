// Prefiltering kernels: distance filter, VoxelGrid, radius / statistical outlier removal.
//
// Replaces: A19 distance_filter (apps/prefiltering_component.cpp:206-229, in-tree),
// A16 pcl::VoxelGrid<PointXYZI>::filter (:167-171; apps/scan_matching_odometry_component.cpp:175-179),
// A17 pcl::RadiusOutlierRemoval (:195-199), A18 pcl::StatisticalOutlierRemoval (:190-194).
#include <cfloat>
#include <climits>
#include <cmath>
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

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
};

// key = dense voxel index (pcl::VoxelGrid: float math, floor(p*inv_leaf) - min_b), value = point index
// RANGE: the distance filter (near < |p| < far, distance_flag_kernel's arithmetic) is applied here instead of compacting first
template <bool RANGE>
__global__ void vg_key_kernel(const float4* __restrict__ in, int n, VgParams prm, double near_t, double far_t, unsigned* __restrict__ keys,
                              int* __restrict__ vals) {
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
  if (ok) {
    const int i0 = (int)(floorf(__fmul_rn(p.x, prm.inv_leaf)) - (float)prm.min_b[0]);
    const int i1 = (int)(floorf(__fmul_rn(p.y, prm.inv_leaf)) - (float)prm.min_b[1]);
    const int i2 = (int)(floorf(__fmul_rn(p.z, prm.inv_leaf)) - (float)prm.min_b[2]);
    key = (unsigned)(i0 * prm.mul[0] + i1 * prm.mul[1] + i2 * prm.mul[2]);
  }
  keys[i] = key;
  vals[i] = i;
}
__global__ void vg_head_kernel(const unsigned* __restrict__ keys, int n, uint8_t* __restrict__ head) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned k = keys[i];
  head[i] = (k != 0xffffffffu && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
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
// one thread per voxel: pcl::CentroidPoint<PointXYZI> — float sums in (stable-sorted = ascending point index) order,
// divided by float(count)
__global__ void vg_centroid_kernel(const float4* __restrict__ in, const int* __restrict__ vals, const int* __restrict__ seg_start, int nseg,
                                   int n_valid, int min_pts, float4* __restrict__ out, uint8_t* __restrict__ keep) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int b = seg_start[s], e = (s + 1 < nseg) ? seg_start[s + 1] : n_valid;
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  for (int j = b; j < e; ++j) {
    const float4 p = __ldg(&in[vals[j]]);
    sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); si = __fadd_rn(si, p.w);
  }
  const float cnt = (float)(e - b);
  out[s] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), __fdiv_rn(si, cnt));
  if (keep) keep[s] = (e - b) >= min_pts ? 1 : 0;
}
__global__ void count_valid_kernel(const unsigned* __restrict__ keys, int n, int* __restrict__ n_valid) {
  // keys sorted ascending: the first 0xffffffff marks the end of the finite points
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool valid = keys[i] != 0xffffffffu;
  const bool next_valid = (i + 1 < n) ? keys[i + 1] != 0xffffffffu : false;
  if (valid && !next_valid) *n_valid = i + 1;
}

void flags_block_offsets(Ctx& ctx, const uint8_t* flags, int n, int* block_off);  // cloud.cu

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
  const int64_t max_idx = (int64_t)div_b[0] * div_b[1] * div_b[2];
  int end_bit = 1;
  while (end_bit < 32 && ((int64_t)1 << end_bit) <= max_idx) ++end_bit;
  end_bit = 32;  // the 0xffffffff sentinel of non-finite points needs all bits

  DBuf<unsigned> k0, k1;
  DBuf<int> v0, v1;
  k0.alloc(n, ctx.stream); k1.alloc(n, ctx.stream); v0.alloc(n, ctx.stream); v1.alloc(n, ctx.stream);
  const int nb = (n + 255) / 256;
  if (range) B2R_LAUNCH(ctx, vg_key_kernel<true>, nb, 256, 0, in, n, prm, range[0], range[1], k0.p, v0.p);
  else B2R_LAUNCH(ctx, vg_key_kernel<false>, nb, 256, 0, in, n, prm, 0.0, 0.0, k0.p, v0.p);
  // stable LSD radix sort (CUB, library primitive): ascending voxel index, ties keep ascending point index
  size_t tmp_bytes = 0;
  B2R_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0.p, k1.p, v0.p, v1.p, n, 0, end_bit, ctx.stream));
  DBuf<uint8_t> tmp; tmp.alloc(tmp_bytes, ctx.stream);
  B2R_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k0.p, k1.p, v0.p, v1.p, n, 0, end_bit, ctx.stream));
  ctx.launches += 5;  // CUB onesweep: histogram + 4 digit passes

  DBuf<uint8_t> head; head.alloc(n, ctx.stream);
  DBuf<int> cnt; cnt.alloc((size_t)nb + 2, ctx.stream);
  B2R_CUDA(cudaMemsetAsync(cnt.p + nb + 1, 0, sizeof(int), ctx.stream));
  B2R_LAUNCH(ctx, vg_head_kernel, nb, 256, 0, k1.p, n, head.p);
  B2R_LAUNCH(ctx, count_valid_kernel, nb, 256, 0, k1.p, n, cnt.p + nb + 1);
  flags_block_offsets(ctx, head.p, n, cnt.p);
  int h2[2] = {0, 0};
  B2R_CUDA(cudaMemcpyAsync(h2, cnt.p + nb, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  const int nseg = h2[0], n_valid = h2[1];
  if (nseg == 0) return;
  DBuf<int> seg; seg.alloc(nseg, ctx.stream);
  B2R_LAUNCH(ctx, vg_seg_kernel, nb, 256, 0, head.p, n, cnt.p, seg.p);
  if (min_pts <= 1) {
    out.pts.alloc(nseg, ctx.stream);
    B2R_LAUNCH(ctx, vg_centroid_kernel, (nseg + 127) / 128, 128, 0, in, v1.p, seg.p, nseg, n_valid, min_pts, out.pts.p, (uint8_t*)nullptr);
    out.n = nseg;
  } else {
    DBuf<float4> cen; cen.alloc(nseg, ctx.stream);
    DBuf<uint8_t> keep; keep.alloc(nseg, ctx.stream);
    B2R_LAUNCH(ctx, vg_centroid_kernel, (nseg + 127) / 128, 128, 0, in, v1.p, seg.p, nseg, n_valid, min_pts, cen.p, keep.p);
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
};
// key = (ix - min) + nx * ((iy - min) + ny * (iz - min)), ixyz = (int)floor(p * inv_leaf) (hpp:88-90): one key per
// ApproximateMeanVoxelGrid hash-map entry; sorting by it only fixes the (implementation-defined) output order
__global__ void amvg_key_kernel(const float4* __restrict__ in, int n, AmvgParams prm, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  const long long i0 = (long long)(int)floorf(__fmul_rn(p.x, prm.inv_leaf)) - prm.min_b[0];
  const long long i1 = (long long)(int)floorf(__fmul_rn(p.y, prm.inv_leaf)) - prm.min_b[1];
  const long long i2 = (long long)(int)floorf(__fmul_rn(p.z, prm.inv_leaf)) - prm.min_b[2];
  keys[i] = (unsigned long long)(i0 * prm.mul[0] + i1 * prm.mul[1] + i2 * prm.mul[2]);
  vals[i] = i;
}
__global__ void amvg_head_kernel(const unsigned long long* __restrict__ keys, int n, uint8_t* __restrict__ head) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || keys[i - 1] != keys[i]) ? 1 : 0;
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
  const unsigned long long max_key = (unsigned long long)(ext[0] * ext[1] * ext[2]);
  int end_bit = 1;
  while (end_bit < 64 && (1ull << end_bit) <= max_key) ++end_bit;
  DBuf<unsigned long long> k0, k1;
  DBuf<int> v0, v1;
  k0.alloc(M, ctx.stream); k1.alloc(M, ctx.stream); v0.alloc(M, ctx.stream); v1.alloc(M, ctx.stream);
  const int nb = (M + 255) / 256;
  B2R_LAUNCH(ctx, amvg_key_kernel, nb, 256, 0, world.pts.p, M, prm, k0.p, v0.p);
  size_t tmp_bytes = 0;  // stable LSD radix sort (CUB): equal keys keep ascending input order = the hash map's accumulation order
  B2R_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0.p, k1.p, v0.p, v1.p, M, 0, end_bit, ctx.stream));
  DBuf<uint8_t> tmp; tmp.alloc(tmp_bytes, ctx.stream);
  B2R_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k0.p, k1.p, v0.p, v1.p, M, 0, end_bit, ctx.stream));
  ctx.launches += 1 + (end_bit + 7) / 8;
  DBuf<uint8_t> head; head.alloc(M, ctx.stream);
  DBuf<int> cnt; cnt.alloc((size_t)nb + 1, ctx.stream);
  B2R_LAUNCH(ctx, amvg_head_kernel, nb, 256, 0, k1.p, M, head.p);
  flags_block_offsets(ctx, head.p, M, cnt.p);
  int nseg = 0;
  B2R_CUDA(cudaMemcpyAsync(&nseg, cnt.p + nb, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  DBuf<int> seg; seg.alloc(nseg, ctx.stream);
  B2R_LAUNCH(ctx, vg_seg_kernel, nb, 256, 0, head.p, M, cnt.p, seg.p);
  DBuf<float4> cen; cen.alloc(nseg, ctx.stream);
  DBuf<uint8_t> vkeep; vkeep.alloc(nseg, ctx.stream);
  // count_threshold_: keep iff count >= min_points_per_voxel (hpp:114)
  B2R_LAUNCH(ctx, vg_centroid_kernel, (nseg + 127) / 128, 128, 0, world.pts.p, v1.p, seg.p, nseg, M, min_points_per_voxel, cen.p, vkeep.p);
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
