// Exact nearest-neighbour search on the uniform grid of a CloudView.
//
// Replaces pcl::search::KdTree / FLANN exact search: kNN for fast_gicp covariances (A5) and
// StatisticalOutlierRemoval (A18); 1-NN for FastGICP correspondences (A9) and getFitnessScore (A13);
// radius counts for RadiusOutlierRemoval (A17).  Distances use FLANN's L2_Simple float association so that
// neighbour sets and squared distances are bit-identical to the kd-tree oracle (ties aside).
//
// All searches are ONE THREAD PER QUERY.  Queries are issued in cell-sorted order, so the 32 lanes of a warp sit in
// the same few cells: their candidate runs coincide (broadcast loads, L1-resident) and their control flow is nearly
// uniform.  A search scans the 3x3x3 cube of cells around the query (centre row first), then grows shell by shell
// until the k-th distance is proven (<= the distance to the border of the scanned cube).  A row of cells along x is a
// contiguous run of the cell-sorted array, so a ring costs two table loads per (y,z) row; rows whose distance
// lower bound already exceeds the current k-th distance are skipped.
//
// topk<K>: the K smallest squared distances, ascending, in registers; insertion is a 2-instruction-per-slot
// min/max chain and is skipped when the candidate does not beat the current K-th value.  Callers that need the
// neighbours themselves (covariances) make a second pass over the same cells with the proven k-th distance as the
// acceptance radius: no index list is kept, which keeps the register footprint at K floats.
#pragma once
#include <climits>

#include "common.cuh"

namespace b2r {

#define B2R_FULL 0xffffffffu
#ifndef B2R_NN1_XPRUNE
#define B2R_NN1_XPRUNE true
#endif

// Everything closer than this (squared) to a query is inside the cube of Chebyshev radius r around the query's cell.
// The 2e-3-cell margin covers the float rounding of (p - bmin) * inv_h for grids up to a few thousand cells per axis.
__device__ __forceinline__ float ring_safe_d2(int r, float h) {
  const float b = ((float)r - 2e-3f) * h;
  return b * b;
}

__device__ __forceinline__ int cells_outside(const CloudView& c, int cx, int cy, int cz) {
  int o = 0;
  o = max(o, max(-cx, cx - (c.gd[0] - 1)));
  o = max(o, max(-cy, cy - (c.gd[1] - 1)));
  o = max(o, max(-cz, cz - (c.gd[2] - 1)));
  return o;
}
__device__ __forceinline__ bool ring_covers_grid(const CloudView& c, int cx, int cy, int cz, int r) {
  return cx - r <= 0 && cx + r >= c.gd[0] - 1 && cy - r <= 0 && cy + r >= c.gd[1] - 1 && cz - r <= 0 && cz + r >= c.gd[2] - 1;
}

// Query position in cell units: integer cell (unclamped, may lie outside the grid) + fraction inside the cell.
struct QueryCell {
  int cx, cy, cz;
  float fx, fy, fz;  // fractional position inside the cell, in [0,1]
};
__device__ __forceinline__ QueryCell query_cell(const CloudView& c, float x, float y, float z) {
  const float big = 1.0e9f;
  QueryCell q;
  const float vx = (x - c.bmin[0]) * c.inv_h, vy = (y - c.bmin[1]) * c.inv_h, vz = (z - c.bmin[2]) * c.inv_h;
  const float ux = floorf(vx), uy = floorf(vy), uz = floorf(vz);
  q.cx = (int)fminf(fmaxf(ux, -big), big);
  q.cy = (int)fminf(fmaxf(uy, -big), big);
  q.cz = (int)fminf(fmaxf(uz, -big), big);
  q.fx = fminf(fmaxf(vx - ux, 0.f), 1.f);
  q.fy = fminf(fmaxf(vy - uy, 0.f), 1.f);
  q.fz = fminf(fmaxf(vz - uz, 0.f), 1.f);
  return q;
}
// lower bound (in cell units, with a 2e-3-cell safety margin for the float rounding of the cell assignment) of the
// distance along one axis from a query at fraction f of its cell to the cell d cells away
__device__ __forceinline__ float axis_gap(float f, int d) {
  const float g = d > 0 ? (float)d - f : (d < 0 ? f - (float)(d + 1) : 0.f);
  return fmaxf(g - 2e-3f, 0.f);
}

// Visits the runs of the cube (shell == false) or of the shell (shell == true) of Chebyshev radius r around the query's
// cell.  V::thr() is the squared distance beyond which points cannot matter any more (it may shrink while scanning);
// V::run(s, e) scans positions [s, e) of spts.  Rows whose distance lower bound exceeds thr() are skipped.  For the
// cube the query's own row is visited first: a good k-th distance early prunes most of the rest.
// XPRUNE additionally trims each row to the cells that can still matter.  That pays for scattered queries (1-NN of
// transformed points); for a cloud's own points in cell order it does not: lanes of one cell then stop sharing
// identical runs (broadcast loads, uniform trip counts), which costs more than the skipped candidates save.
template <bool XPRUNE, typename V>
__device__ __forceinline__ void visit_ring(const CloudView& c, const QueryCell& q, int r, bool shell, V& v) {
  const int z0 = max(q.cz - r, 0), z1 = min(q.cz + r, c.gd[2] - 1);
  const int y0 = max(q.cy - r, 0), y1 = min(q.cy + r, c.gd[1] - 1);
  const int fx0 = max(q.cx - r, 0), fx1 = min(q.cx + r, c.gd[0] - 1);
  const float inv_h2 = c.inv_h * c.inv_h;
  const bool centre_first = !shell && q.cy >= y0 && q.cy <= y1 && q.cz >= z0 && q.cz <= z1;
  if (centre_first && fx0 <= fx1) {
    const int rowbase = (q.cz * c.gd[1] + q.cy) * c.gd[0];
    v.run(__ldg(&c.cell_start[rowbase + fx0]), __ldg(&c.cell_start[rowbase + fx1 + 1]));
  }
  for (int z = z0; z <= z1; ++z) {
    const int dz = z - q.cz;
    const bool zface = dz == r || dz == -r;
    const float gz = axis_gap(q.fz, dz);
    for (int y = y0; y <= y1; ++y) {
      const int dy = y - q.cy;
      if (centre_first && dy == 0 && dz == 0) continue;
      const float gy = axis_gap(q.fy, dy);
      // budget left for the x axis, in cells^2 (thr() == INFINITY: everything)
      const float bx2 = v.thr() * inv_h2 - (gy * gy + gz * gz);
      if (!(bx2 > 0.f)) continue;
      const int rowbase = (z * c.gd[1] + y) * c.gd[0];
      if (!shell || zface || dy == r || dy == -r) {
        int x0 = fx0, x1 = fx1;
        if (XPRUNE) {
          const float bx = fminf(sqrtf(bx2), (float)r + 1.f) + 2e-3f;
          x0 = max(q.cx - min(r, (int)floorf(bx + 1.f - q.fx)), 0);
          x1 = min(q.cx + min(r, (int)floorf(bx + q.fx)), c.gd[0] - 1);
        }
        if (x0 <= x1) v.run(__ldg(&c.cell_start[rowbase + x0]), __ldg(&c.cell_start[rowbase + x1 + 1]));
      } else {
        const int xa = q.cx - r, xb = q.cx + r;
        const float ga = axis_gap(q.fx, -r), gb = axis_gap(q.fx, r);
        if (xa >= 0 && xa < c.gd[0] && ga * ga < bx2) v.run(__ldg(&c.cell_start[rowbase + xa]), __ldg(&c.cell_start[rowbase + xa + 1]));
        if (xb >= 0 && xb < c.gd[0] && gb * gb < bx2) v.run(__ldg(&c.cell_start[rowbase + xb]), __ldg(&c.cell_start[rowbase + xb + 1]));
      }
    }
  }
}

// ---- K smallest squared distances, ascending, in registers
template <int K>
__device__ __forceinline__ void topk_insert(float (&d)[K], float v) {
#pragma unroll
  for (int i = K - 1; i > 0; --i) d[i] = fmaxf(d[i - 1], fminf(d[i], v));
  d[0] = fminf(d[0], v);
}
// d[k-1] for a runtime k in [1, K]
template <int K>
__device__ __forceinline__ float topk_kth(const float (&d)[K], int k) {
  float r = d[K - 1];
#pragma unroll
  for (int i = K - 2; i >= 0; --i)
    r = (i >= k - 1) ? d[i] : r;  // (an equality test here lets the compiler fold it into a dynamic index = local memory)
  return r;
}

template <int K>
struct TopkVisitor {
  const CloudView& c;
  float qx, qy, qz;
  float d[K];
  __device__ __forceinline__ TopkVisitor(const CloudView& cv, float x, float y, float z) : c(cv), qx(x), qy(y), qz(z) {
#pragma unroll
    for (int i = 0; i < K; ++i) d[i] = INFINITY;
  }
  __device__ __forceinline__ float thr() const { return d[K - 1]; }
  __device__ __forceinline__ void run(int s, int e) {
    for (int j = s; j < e; ++j) {
      const float4 p = __ldg(&c.spts[j]);
      const float d2 = dist2_flann(qx, qy, qz, p.x, p.y, p.z);
      if (d2 < d[K - 1]) topk_insert<K>(d, d2);
    }
  }
};

// Exact k smallest squared distances (k <= K, ascending in v.d[0..k)) of the query in cloud c.
// Returns the Chebyshev radius of the scanned cube; v.d[k-1] == INFINITY if the cloud has fewer than k points.
template <int K>
__device__ __forceinline__ int knn_topk(const CloudView& c, const QueryCell& q, int k, TopkVisitor<K>& v) {
  int r = cells_outside(c, q.cx, q.cy, q.cz) + 1;
  visit_ring<false>(c, q, r, false, v);
  for (;;) {
    if (topk_kth<K>(v.d, k) <= ring_safe_d2(r, c.h)) break;
    if (ring_covers_grid(c, q.cx, q.cy, q.cz, r)) break;
    ++r;
    visit_ring<false>(c, q, r, true, v);
  }
  return r;
}

// ---- same search keeping (distance, position) pairs as 64-bit keys: float bits of d^2 << 32 | position in spts.
// Used where the neighbour identities are wanted in ascending order (debug / test entry point).
__device__ __forceinline__ unsigned long long u64min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long u64max(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

template <int K>
struct TopkKeyVisitor {
  const CloudView& c;
  float qx, qy, qz;
  unsigned long long d[K];
  __device__ __forceinline__ TopkKeyVisitor(const CloudView& cv, float x, float y, float z) : c(cv), qx(x), qy(y), qz(z) {
#pragma unroll
    for (int i = 0; i < K; ++i) d[i] = ~0ull;
  }
  __device__ __forceinline__ float worst() const { return __uint_as_float((unsigned)(d[K - 1] >> 32)); }
  __device__ __forceinline__ float thr() const { return d[K - 1] == ~0ull ? INFINITY : worst() * (1.f + 1e-6f); }
  __device__ __forceinline__ void run(int s, int e) {
    for (int j = s; j < e; ++j) {
      const float4 p = __ldg(&c.spts[j]);
      const float d2 = dist2_flann(qx, qy, qz, p.x, p.y, p.z);
      const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)j;
      if (key < d[K - 1]) {
#pragma unroll
        for (int i = K - 1; i > 0; --i) d[i] = u64max(d[i - 1], u64min(d[i], key));
        d[0] = u64min(d[0], key);
      }
    }
  }
};

// ---- one thread per query: exact 1-NN within max_d2 (pass INFINITY for unbounded).  Returns position in spts or -1.
// Ties go to the lower position.
struct Nn1Visitor {
  const CloudView& c;
  float qx, qy, qz, cut;  // rows farther than `cut` (squared) cannot matter
  float best;
  int best_pos;
  __device__ __forceinline__ float thr() const { return fminf(best, cut) * (1.f + 1e-6f); }
  __device__ __forceinline__ void run(int s, int e) {
    for (int j = s; j < e; ++j) {
      const float4 p = __ldg(&c.spts[j]);
      const float d2 = dist2_flann(qx, qy, qz, p.x, p.y, p.z);
      if (d2 < best || (d2 == best && j < best_pos)) { best = d2; best_pos = j; }
    }
  }
};
__device__ __forceinline__ int nn1_search(const CloudView& c, float qx, float qy, float qz, float max_d2, float& best_out) {
  Nn1Visitor v{c, qx, qy, qz, max_d2, INFINITY, -1};
  if (c.n == 0) { best_out = v.best; return -1; }
  const QueryCell q = query_cell(c, qx, qy, qz);
  int r = cells_outside(c, q.cx, q.cy, q.cz) + 1;
  // nothing closer than (r-2)*h can exist when the query is outside the grid
  {
    const float lb = (float)(r - 2) * c.h;
    if (r >= 3 && lb * lb > max_d2) { best_out = v.best; return -1; }
  }
  visit_ring<B2R_NN1_XPRUNE>(c, q, r, false, v);
  for (;;) {
    const float b2 = ring_safe_d2(r, c.h);
    if (v.best <= b2) break;     // proven nearest
    if (b2 > max_d2) break;      // anything farther is out of range anyway
    if (ring_covers_grid(c, q.cx, q.cy, q.cz, r)) break;
    ++r;
    visit_ring<B2R_NN1_XPRUNE>(c, q, r, true, v);
  }
  best_out = v.best;
  return v.best_pos;
}

// ---- one thread per query: number of points with d2 < r2 (strict), early exit once count > stop_above.
// Requires c.h >= radius so that the 3x3x3 block around the query cell suffices.
__device__ __forceinline__ int radius_count(const CloudView& c, float qx, float qy, float qz, float r2, int stop_above) {
  int cx, cy, cz;
  nn_cell_of(c, qx, qy, qz, cx, cy, cz);
  int cnt = 0;
  const int z0 = max(cz - 1, 0), z1 = min(cz + 1, c.gd[2] - 1);
  const int y0 = max(cy - 1, 0), y1 = min(cy + 1, c.gd[1] - 1);
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, c.gd[0] - 1);
  for (int z = z0; z <= z1; ++z)
    for (int y = y0; y <= y1; ++y) {
      const int rowbase = (z * c.gd[1] + y) * c.gd[0];
      const int s = __ldg(&c.cell_start[rowbase + x0]), e = __ldg(&c.cell_start[rowbase + x1 + 1]);
      for (int j = s; j < e; ++j) {
        const float4 p = __ldg(&c.spts[j]);
        if (dist2_flann(qx, qy, qz, p.x, p.y, p.z) < r2) {
          if (++cnt > stop_above) return cnt;
        }
      }
    }
  return cnt;
}

}  // namespace b2r
