// Exact nearest-neighbour search on the uniform grid of a CloudView.
//
// Replaces pcl::search::KdTree / FLANN exact search: kNN for fast_gicp covariances (A5) and
// StatisticalOutlierRemoval (A18); 1-NN for FastGICP correspondences (A9) and getFitnessScore (A13);
// radius counts for RadiusOutlierRemoval (A17).  Distances use FLANN's L2_Simple float association so that
// neighbour sets and squared distances are bit-identical to the kd-tree oracle (ties aside).
//
// warp_knn: one warp per query.  Lane i ends up holding the i-th nearest neighbour as a 64-bit key
// (float bits of d^2 << 32 | position in the cell-sorted array); ~0 = none.  The search scans the cube of
// cells around the query, then grows shell by shell until the k-th distance is proven (<= r*h).
// A ring is enumerated as runs (contiguous spans of the cell-sorted array): every lane fetches the bounds of
// one run (one memory latency for up to 32 rows), then the warp streams the concatenated candidates in full
// 32-wide chunks.  Candidates that beat the current k-th distance are appended to a per-warp shared-memory
// buffer; a full buffer is bitonic-sorted and merged into the sorted top-32 list held across the lanes.
#pragma once
#include "common.cuh"

namespace b2r {

#define B2R_FULL 0xffffffffu

__device__ __forceinline__ unsigned long long u64min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long u64max(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

// ascending bitonic sort of one key per lane
__device__ __forceinline__ unsigned long long warp_sort32(unsigned long long v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long o = __shfl_xor_sync(B2R_FULL, v, j);
      const bool keep_min = ((lane & k) == 0) == ((lane & j) == 0);
      v = ((v > o) == keep_min) ? o : v;  // one 64-bit compare per exchange
    }
  }
  return v;
}
// both ascending; returns the 32 smallest of the union, ascending
__device__ __forceinline__ unsigned long long warp_merge32(unsigned long long list, unsigned long long cand, int lane) {
  const unsigned long long rev = __shfl_sync(B2R_FULL, cand, 31 - lane);
  unsigned long long m = u64min(list, rev);
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const unsigned long long o = __shfl_xor_sync(B2R_FULL, m, j);
    m = ((m > o) == ((lane & j) == 0)) ? o : m;
  }
  return m;
}

struct KnnScratch {  // per-warp shared memory
  unsigned long long pend[32];
};

struct KnnState {
  unsigned long long list;  // lane i: i-th best so far
  unsigned long long kth;   // key of the current k-th best (warp-uniform)
  int npend;                // candidates waiting in the scratch buffer
};

__device__ __forceinline__ void knn_flush(KnnState& s, KnnScratch& sm, int k, int lane) {
  if (s.npend == 0) return;
  __syncwarp();
  const unsigned long long v = lane < s.npend ? sm.pend[lane] : ~0ull;
  __syncwarp();
  s.list = warp_merge32(s.list, warp_sort32(v, lane), lane);
  s.kth = __shfl_sync(B2R_FULL, s.list, k - 1);
  s.npend = 0;
}

__device__ __forceinline__ void knn_push(KnnState& s, KnnScratch& sm, unsigned long long key, int k, int lane) {
  const bool want = key < s.kth;
  const unsigned m = __ballot_sync(B2R_FULL, want);
  if (!m) return;
  const int cnum = __popc(m);
  if (s.npend + cnum > 32) knn_flush(s, sm, k, lane);
  if (want) sm.pend[s.npend + __popc(m & ((1u << lane) - 1u))] = key;
  s.npend += cnum;
}

// Every lane owns one run [s,e) of the cell-sorted array; the warp scans the concatenation of the 32 runs.
__device__ __forceinline__ void knn_scan_runs(const CloudView& c, float qx, float qy, float qz, int s, int e, KnnState& st, KnnScratch& sm,
                                              int k, int lane) {
  const int len = e - s;
  int incl = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(B2R_FULL, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(B2R_FULL, incl, 31);
  const int excl = incl - len;
  for (int base = 0; base < total; base += 32) {
    const int j = base + lane;
    int lo = 0;  // number of runs that end at or before j  (= index of the run containing j)
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int v = __shfl_sync(B2R_FULL, incl, lo + step - 1);
      if (v <= j) lo += step;
    }
    const int rs = __shfl_sync(B2R_FULL, s, lo), rx = __shfl_sync(B2R_FULL, excl, lo);
    unsigned long long key = ~0ull;
    if (j < total) {
      const int pos = rs + (j - rx);
      const float4 p = __ldg(&c.spts[pos]);
      const float d2 = dist2_flann(qx, qy, qz, p.x, p.y, p.z);
      key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)pos;
    }
    knn_push(st, sm, key, k, lane);
  }
}

// Bounds of run `id` of the cube (shell == false) or shell (shell == true) of Chebyshev radius r around (cx,cy,cz).
// Rows are enumerated over the part of the (y,z) window that lies inside the grid; a shell contributes one full
// x-span for face rows and two single cells (x = cx-r, cx+r) for inner rows.
struct RingEnum {
  int ylo, zlo, wy, nrows, nruns;
};
__device__ __forceinline__ RingEnum ring_enum(const CloudView& c, int cy, int cz, int r, bool shell) {
  RingEnum e;
  e.ylo = max(cy - r, 0);
  e.zlo = max(cz - r, 0);
  const int yhi = min(cy + r, c.gd[1] - 1), zhi = min(cz + r, c.gd[2] - 1);
  e.wy = max(yhi - e.ylo + 1, 0);
  const int wz = max(zhi - e.zlo + 1, 0);
  e.nrows = e.wy * wz;
  e.nruns = shell ? 2 * e.nrows : e.nrows;
  return e;
}
__device__ __forceinline__ void ring_run_bounds(const CloudView& c, const RingEnum& en, int cx, int cy, int cz, int r, bool shell, int id,
                                                int& s, int& e) {
  s = e = 0;
  if (id >= en.nruns) return;
  int row = shell ? (id >> 1) : id;
  const int seg = shell ? (id & 1) : 0;
  if (!shell && r == 1 && en.nrows == 9) {
    // centre row first: a good k-th distance early prunes the rest
    row = (int)((0x862075314ull >> (4 * row)) & 0xfull);  // order 4,1,3,5,7,0,2,6,8
  }
  const int y = en.ylo + row % en.wy, z = en.zlo + row / en.wy;
  const int rowbase = (z * c.gd[1] + y) * c.gd[0];
  const bool face = !shell || (y - cy == r) || (cy - y == r) || (z - cz == r) || (cz - z == r);
  if (face) {
    if (seg) return;
    const int x0 = max(cx - r, 0), x1 = min(cx + r, c.gd[0] - 1);
    if (x0 > x1) return;
    s = __ldg(&c.cell_start[rowbase + x0]);
    e = __ldg(&c.cell_start[rowbase + x1 + 1]);
  } else {
    const int x = seg ? cx + r : cx - r;
    if (x < 0 || x >= c.gd[0]) return;
    s = __ldg(&c.cell_start[rowbase + x]);
    e = __ldg(&c.cell_start[rowbase + x + 1]);
  }
}

__device__ __forceinline__ void knn_scan_ring(const CloudView& c, float qx, float qy, float qz, int cx, int cy, int cz, int r, bool shell,
                                              KnnState& st, KnnScratch& sm, int k, int lane) {
  const RingEnum en = ring_enum(c, cy, cz, r, shell);
  for (int base = 0; base < en.nruns; base += 32) {
    int s, e;
    ring_run_bounds(c, en, cx, cy, cz, r, shell, base + lane, s, e);
    if (__ballot_sync(B2R_FULL, e > s)) knn_scan_runs(c, qx, qy, qz, s, e, st, sm, k, lane);
  }
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

__device__ __forceinline__ unsigned long long warp_knn(const CloudView& c, float qx, float qy, float qz, int k, int lane, KnnScratch& sm) {
  KnnState st;
  st.list = ~0ull; st.kth = ~0ull; st.npend = 0;
  int cx, cy, cz;
  nn_cell_of_unclamped(c, qx, qy, qz, cx, cy, cz);
  int r = cells_outside(c, cx, cy, cz) + 1;
  knn_scan_ring(c, qx, qy, qz, cx, cy, cz, r, false, st, sm, k, lane);
  knn_flush(st, sm, k, lane);
  for (;;) {
    if (st.kth != ~0ull) {
      const float kd2 = __uint_as_float((unsigned)(st.kth >> 32));
      const float bound = (float)r * c.h;
      if (kd2 <= bound * bound * (1.0f - 1e-5f)) break;
    }
    if (ring_covers_grid(c, cx, cy, cz, r)) break;
    ++r;
    knn_scan_ring(c, qx, qy, qz, cx, cy, cz, r, true, st, sm, k, lane);
    knn_flush(st, sm, k, lane);
  }
  return st.list;
}

// ---- one thread per query: exact 1-NN within max_d2 (pass INFINITY for unbounded).  Returns position in spts or -1.
__device__ __forceinline__ void nn1_scan_run(const CloudView& c, float qx, float qy, float qz, int s, int e, float& best, int& best_pos) {
  for (int j = s; j < e; ++j) {
    const float4 p = __ldg(&c.spts[j]);
    const float d2 = dist2_flann(qx, qy, qz, p.x, p.y, p.z);
    if (d2 < best || (d2 == best && j < best_pos)) { best = d2; best_pos = j; }
  }
}
__device__ __forceinline__ void nn1_scan_ring(const CloudView& c, float qx, float qy, float qz, int cx, int cy, int cz, int r, bool shell,
                                              float& best, int& best_pos) {
  const int z0 = max(cz - r, 0), z1 = min(cz + r, c.gd[2] - 1);
  const int y0 = max(cy - r, 0), y1 = min(cy + r, c.gd[1] - 1);
  for (int z = z0; z <= z1; ++z) {
    const bool zface = (z - cz == r) || (cz - z == r);
    for (int y = y0; y <= y1; ++y) {
      const int rowbase = (z * c.gd[1] + y) * c.gd[0];
      const bool full = !shell || zface || (y - cy == r) || (cy - y == r);
      if (full) {
        const int x0 = max(cx - r, 0), x1 = min(cx + r, c.gd[0] - 1);
        if (x0 <= x1) nn1_scan_run(c, qx, qy, qz, __ldg(&c.cell_start[rowbase + x0]), __ldg(&c.cell_start[rowbase + x1 + 1]), best, best_pos);
      } else {
        const int xa = cx - r, xb = cx + r;
        if (xa >= 0 && xa < c.gd[0]) nn1_scan_run(c, qx, qy, qz, __ldg(&c.cell_start[rowbase + xa]), __ldg(&c.cell_start[rowbase + xa + 1]), best, best_pos);
        if (xb >= 0 && xb < c.gd[0]) nn1_scan_run(c, qx, qy, qz, __ldg(&c.cell_start[rowbase + xb]), __ldg(&c.cell_start[rowbase + xb + 1]), best, best_pos);
      }
    }
  }
}
__device__ __forceinline__ int nn1_search(const CloudView& c, float qx, float qy, float qz, float max_d2, float& best_out) {
  float best = INFINITY;
  int best_pos = -1;
  if (c.n == 0) { best_out = best; return -1; }
  int cx, cy, cz;
  nn_cell_of_unclamped(c, qx, qy, qz, cx, cy, cz);
  int r = cells_outside(c, cx, cy, cz) + 1;
  // nothing closer than (r-2)*h can exist when the query is outside the grid
  {
    const float lb = (float)(r - 2) * c.h;
    if (r >= 3 && lb * lb > max_d2) { best_out = best; return -1; }
  }
  nn1_scan_ring(c, qx, qy, qz, cx, cy, cz, r, false, best, best_pos);
  for (;;) {
    const float bound = (float)r * c.h;
    const float b2 = bound * bound * (1.0f - 1e-5f);
    if (best <= b2) break;       // proven nearest
    if (b2 > max_d2) break;      // anything farther is out of range anyway
    if (ring_covers_grid(c, cx, cy, cz, r)) break;
    ++r;
    nn1_scan_ring(c, qx, qy, qz, cx, cy, cz, r, true, best, best_pos);
  }
  best_out = best;
  return best_pos;
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
