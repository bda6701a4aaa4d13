// Exact nearest-neighbour search on the cell grid of a CloudView.
//
// Replaces pcl::search::KdTree / FLANN exact search: kNN for fast_gicp covariances (A5) and
// StatisticalOutlierRemoval (A18); 1-NN for FastGICP correspondences (A9) and getFitnessScore (A13);
// radius counts for RadiusOutlierRemoval (A17).  Distances use FLANN's L2_Simple float association so that
// neighbour sets and squared distances are bit-identical to the kd-tree oracle (ties aside).
//
// Layout (cloud.cu): points are copied in cell order (z, y, x) and, inside a cell, in ascending x.  A ROW of cells
// along x is therefore one contiguous run of the array, globally sorted by x; the cells along x (small: hx ~ h/3) are
// only an index into it.
//
// Search: ONE THREAD PER QUERY.  For every row within Chebyshev distance r (in y, z cells of size h) of the query's
// row, the thread scans the x-cell the query falls into and then SWEEPS the sorted row outwards in both directions,
// stopping a direction as soon as dx^2 + gap^2 exceeds the current bound (k-th distance, search radius ...), gap being the
// row's (y, z) distance lower bound: |dx| only grows along a sweep, so nothing beyond can qualify, and the strip of a row
// one cell away is narrower than the bound itself.  So the candidates of a row are just that strip, at point granularity,
// whatever the cell size.  Rows whose gap already exceeds the bound are skipped.  r grows ring by ring (whole rows, so a ring
// is just the rows with max(|dy|,|dz|) == r) until the k-th distance is proven: <= the distance to the nearest unvisited row.
// Queries are issued in cell order (also the SOURCE points of a fitness / FAST_GICP pass: in their own cloud's cell order),
// so the lanes of a warp walk nearly the same rows and runs (L1-resident, partly broadcast loads).
//
// Candidate arithmetic: FLANN's float squared distance, formed by Blackwell's packed-pair FP32 instructions — per candidate
// (x, y) together (cand_d2), or, in the 1-NN searches, TWO neighbouring points of a row per instruction from the
// pair-interleaved copy `spair` (VISIT_PAIRS) — every element rounded exactly like the scalar chain, so distances stay
// bit-identical.
//
// topk<K>: the K smallest squared distances, ascending, in registers; insertion is a 2-instruction-per-slot
// min/max chain and is skipped when the candidate does not beat the current K-th value.  The covariance kernel logs every
// candidate that entered the top-k (TopkListVisitor) and re-reads only those in its second pass.
#pragma once
#include <climits>

#include "common.cuh"

namespace b2r {

#define B2R_FULL 0xffffffffu

// Everything closer than this (squared) to a query has its (y, z) cell within Chebyshev distance r of the query's.
// The 2e-3-cell margin covers the float rounding of (p - bmin) * inv_h for grids up to a few thousand cells per axis.
__device__ __forceinline__ float ring_safe_d2(int r, float h) {
  const float b = ((float)r - 2e-3f) * h;
  return b * b;
}

// Query position in cell units: integer cell (y, z unclamped: may lie outside the grid; x clamped: it only indexes
// into the row) + fraction inside the cell along y and z.
struct QueryCell {
  int cx, cy, cz;
  float fy, fz;  // fractional position inside the cell, in [0,1]
  float xout2;   // squared distance from the query to the grid's x extent (0 inside): lower bound for every point
};
__device__ __forceinline__ QueryCell query_cell(const CloudView& c, float x, float y, float z) {
  const float big = 1.0e9f;
  QueryCell q;
  const float vy = (y - c.bmin[1]) * c.inv_h, vz = (z - c.bmin[2]) * c.inv_h;
  const float uy = floorf(vy), uz = floorf(vz);
  q.cx = clampi((int)fminf(fmaxf(floorf((x - c.bmin[0]) * c.inv_hx), -big), big), 0, c.gd[0] - 1);
  q.cy = (int)fminf(fmaxf(uy, -big), big);
  q.cz = (int)fminf(fmaxf(uz, -big), big);
  q.fy = fminf(fmaxf(vy - uy, 0.f), 1.f);
  q.fz = fminf(fmaxf(vz - uz, 0.f), 1.f);
  const float ox = fmaxf(fmaxf(c.bmin[0] - x, x - c.bmax[0]), 0.f) * (1.f - 1e-6f);
  q.xout2 = ox * ox;
  return q;
}
__device__ __forceinline__ int rows_outside(const CloudView& c, const QueryCell& q) {
  int o = 0;
  o = max(o, max(-q.cy, q.cy - (c.gd[1] - 1)));
  o = max(o, max(-q.cz, q.cz - (c.gd[2] - 1)));
  return o;
}
__device__ __forceinline__ bool ring_covers_grid(const CloudView& c, const QueryCell& q, int r) {
  return q.cy - r <= 0 && q.cy + r >= c.gd[1] - 1 && q.cz - r <= 0 && q.cz + r >= c.gd[2] - 1;
}
// lower bound (in cell units, with a 2e-3-cell safety margin for the float rounding of the cell assignment) of the
// distance along one axis from a query at fraction f of its cell to the cell d cells away
__device__ __forceinline__ float axis_gap(float f, int d) {
  const float g = d > 0 ? (float)d - f : (d < 0 ? f - (float)(d + 1) : 0.f);
  return fmaxf(g - 2e-3f, 0.f);
}

// Where a visitor's candidate points come from.  GlobalPts: the cloud's cell-sorted copy in global memory (L1 / L2).
// TilePts (knn_cov_tile_kernel): the strips of the rows around a block's queries, staged in shared memory by TMA bulk copies
// (one cp.async.bulk per row: a row's strip is a contiguous range of spts); positions outside the staged strip of a row
// fall back to global memory, so the staging is a cache, never a correctness condition.
struct GlobalPts {
  struct Row {};
  __device__ __forceinline__ Row row(int, int) const { return Row{}; }
  __device__ __forceinline__ float4 load(const CloudView& c, const Row&, int j) const { return __ldg(&c.spts[j]); }
};
struct TileRow { int lo, hi, off; };  // positions [lo, hi) of spts live at tile[off ...]
struct TilePts {
  const float4* tile;    // shared memory
  const TileRow* rows;   // shared memory: the (ny x nz) rectangle of rows starting at (y_base, z_base)
  int y_base, z_base, ny, nz;
  struct Row { int lo, hi, delta; };
  __device__ __forceinline__ Row row(int y, int z) const {
    const int iy = y - y_base, iz = z - z_base;
    if (iy < 0 || iy >= ny || iz < 0 || iz >= nz) return Row{0, 0, 0};
    const TileRow t = rows[iz * ny + iy];
    return Row{t.lo, t.hi, t.off - t.lo};
  }
  __device__ __forceinline__ float4 load(const CloudView& c, const Row& r, int j) const {
    return (j >= r.lo && j < r.hi) ? tile[j + r.delta] : __ldg(&c.spts[j]);
  }
};

// Visits the rows at Chebyshev (y, z) distance <= r (ring == false) or == r (ring == true) of the query's row.
// Visitor interface:  thr()      squared distance beyond which a ROW cannot matter (may shrink while scanning);
//                     stop(s)    true when a point whose squared distance is at least s cannot matter, s = dx^2 + rg2: the
//                                squared x-distance plus the row's squared (y, z) gap, a lower bound of the distance to
//                                anything in that row (the strip of a row one cell away is narrower than the bound itself);
//                     test(p,j)  candidate p = spts[j].
// s is a bound on the REAL distance while candidates are compared by their float-evaluated distance, which can fall short of
// it by a few ulps; every visitor therefore compares s with its bound times kGapSlack (the 2e-3-cell margin of the gaps is
// metres, not ulps, but it vanishes for a gap of zero, where s = fl(dx^2) <= the float distance exactly).
// The query's own row comes first: a good k-th distance early prunes most of the rest.
constexpr float kGapSlack = 1.f + 2e-6f;
// One candidate: FLANN's float squared distance to the query (bit-identical to dist2_flann) and, on the way, fl(dx^2) for the
// sweep's stop test — the x / y differences and squares come out of one FFMA2 + one FMUL2 (packed-pair FP32, sm_100).
struct CandD2 { float dx, dx2, d2; };
__device__ __forceinline__ CandD2 cand_d2(float qx, float qy, float qz, const float4& p) {
#if defined(B2R_NO_F32X2)
  const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
  const float dx2 = __fmul_rn(dx, dx);
  return CandD2{dx, dx2, __fadd_rn(__fadd_rn(dx2, __fmul_rn(dy, dy)), __fmul_rn(dz, dz))};
#else
  const float2 d = __ffma2_rn(make_float2(p.x, p.y), make_float2(-1.f, -1.f), make_float2(qx, qy));
  const float2 sq = __fmul2_rn(d, d);
  const float dz = __fsub_rn(qz, p.z);
  return CandD2{d.x, sq.x, __fadd_rn(__fadd_rn(sq.x, sq.y), __fmul_rn(dz, dz))};
#endif
}
// How a row is swept (A/B-measured on the B200):
//   VISIT_CELL3   the query's x-cell unconditionally (independent loads, unrolled by the compiler), then a sweep towards smaller
//                 and one towards larger x, each stopped by stop(dx^2 + rg2)
//   VISIT_MERGED  one loop from the start of the query's x-cell towards larger x (points before the query are skipped, not
//                 tested, when they cannot matter) and one towards smaller x: fewer candidates, but every load is serialised
//                 behind the previous test
//   VISIT_AHEAD   VISIT_CELL3 with the sweeps loading one point ahead of the test
//   VISIT_PAIRS   (1-NN visitors only) candidates two at a time from the pair-interleaved copy `spair`: the three differences, their
//                 squares and the two sums are packed-pair FP32 instructions (8 issue slots for two candidates instead of 12), and
//                 the loop overhead halves.  A pair may reach one position beyond the row's ends: that is another real point of the
//                 cloud (or the +inf pad), harmless for a minimum — which is why k-NN / radius visitors, where a point must not be
//                 counted twice, keep the single-candidate sweeps.
enum { VISIT_CELL3 = 0, VISIT_MERGED = 1, VISIT_AHEAD = 2, VISIT_SWEEP_MASK = 3, VISIT_LANE_RING = 4, VISIT_PAIRS = 8 };
#ifndef B2R_VISIT_DEFAULT
#define B2R_VISIT_DEFAULT VISIT_MERGED
#endif
template <int MODE, typename V, typename S>
__device__ __forceinline__ void visit_row(const CloudView& c, const QueryCell& q, float qx, int y, int z, float rg2, V& v, const S& src) {
  const int rowbase = (z * c.gd[1] + y) * c.gd[0];
  const int row_s = __ldg(&c.cell_start[rowbase]), row_e = __ldg(&c.cell_start[rowbase + c.gd[0]]);
  if (row_s == row_e) return;
  const typename S::Row sr = src.row(y, z);
  const float qy = v.qy_f(), qz = v.qz_f();
  if constexpr ((MODE & VISIT_PAIRS) != 0) {
    const int cs = __ldg(&c.cell_start[rowbase + q.cx]), ce = __ldg(&c.cell_start[rowbase + q.cx + 1]);
    const unsigned long long qx2 = f2_pack(qx, qx), qy2 = f2_pack(qy, qy), qz2 = f2_pack(qz, qz);
    const float4* __restrict__ sp = c.spair;
    // FLANN's ((dx^2 + dy^2) + dz^2) for the two points of record m, every operation rounded like its scalar form
#define B2R_PAIR_D2(m)                                                                          \
    const float4 pa = __ldg(&sp[2 * (size_t)(m)]), pb = __ldg(&sp[2 * (size_t)(m) + 1]);          \
    const unsigned long long dxp = f2_sub(qx2, f2_pack(pa.x, pa.y));                             \
    const unsigned long long dyp = f2_sub(qy2, f2_pack(pa.z, pa.w));                             \
    const unsigned long long dzp = f2_sub(qz2, f2_pack(pb.x, pb.y));                             \
    const float2 sx = f2_unpack(f2_mul(dxp, dxp)), sy = f2_unpack(f2_mul(dyp, dyp)), sz = f2_unpack(f2_mul(dzp, dzp)); \
    /* the sums stay scalar: ptxas contracts a packed multiply feeding a packed add into FFMA2 even when both carry .rn */      \
    const float2 d2 = make_float2(__fadd_rn(__fadd_rn(sx.x, sy.x), sz.x), __fadd_rn(__fadd_rn(sx.y, sy.y), sz.y));
    const int mc0 = cs >> 1;
    int mu = mc0;  // first record of the sweep towards larger x
    if (ce > cs) {
      const int mc1 = (ce - 1) >> 1;
      for (int m = mc0; m <= mc1; ++m) {  // the query's x-cell: unconditional, independent loads
        B2R_PAIR_D2(m)
        v.test2(d2.x, d2.y, 2 * m);
      }
      mu = mc1 + 1;
    }
    for (int m = mu; 2 * m < row_e; ++m) {  // towards larger x
      B2R_PAIR_D2(m)
      // the record's first point decides: it is the nearer one in x, provided it lies beyond the query's x-cell (the first
      // record after an EMPTY cell may begin one position earlier — in a lower cell, or in the previous row)
      if (2 * m >= cs && v.stop(__fadd_rn(sx.x, rg2))) break;
      v.test2(d2.x, d2.y, 2 * m);
    }
    for (int m = mc0 - 1; m >= 0 && 2 * m + 1 >= row_s; --m) {  // towards smaller x: the record's second point is the nearer one
      B2R_PAIR_D2(m)
      if (v.stop(__fadd_rn(sx.y, rg2))) break;
      v.test2(d2.x, d2.y, 2 * m);
    }
#undef B2R_PAIR_D2
  } else if constexpr ((MODE & VISIT_SWEEP_MASK) == VISIT_MERGED) {
    const int cs = __ldg(&c.cell_start[rowbase + q.cx]);
    for (int j = cs; j < row_e; ++j) {  // the query's x-cell, then towards larger x
      const float4 p = src.load(c, sr, j);
      const CandD2 d = cand_d2(qx, qy, qz, p);
      if (v.stop(__fadd_rn(d.dx2, rg2))) {
        if (!(d.dx > 0.f)) break;  // at or beyond the query: everything after it is farther still
        continue;                  // inside the cell, before the query: the following points are closer
      }
      v.test(p, j, d.d2);
    }
    for (int j = cs - 1; j >= row_s; --j) {  // towards smaller x
      const float4 p = src.load(c, sr, j);
      const CandD2 d = cand_d2(qx, qy, qz, p);
      if (v.stop(__fadd_rn(d.dx2, rg2))) break;
      v.test(p, j, d.d2);
    }
  } else {
    const int cs = __ldg(&c.cell_start[rowbase + q.cx]), ce = __ldg(&c.cell_start[rowbase + q.cx + 1]);
    for (int j = cs; j < ce; ++j) {
      const float4 p = src.load(c, sr, j);
      v.test(p, j, cand_d2(qx, qy, qz, p).d2);
    }
    if constexpr ((MODE & VISIT_SWEEP_MASK) == VISIT_CELL3) {
      for (int j = cs - 1; j >= row_s; --j) {  // towards smaller x
        const float4 p = src.load(c, sr, j);
        const CandD2 d = cand_d2(qx, qy, qz, p);
        if (v.stop(__fadd_rn(d.dx2, rg2))) break;
        v.test(p, j, d.d2);
      }
      for (int j = ce; j < row_e; ++j) {  // towards larger x
        const float4 p = src.load(c, sr, j);
        const CandD2 d = cand_d2(qx, qy, qz, p);
        if (v.stop(__fadd_rn(d.dx2, rg2))) break;
        v.test(p, j, d.d2);
      }
    } else {
      if (cs > row_s) {
        int j = cs - 1;
        float4 p = src.load(c, sr, j);
        for (;;) {
          const bool more = j > row_s;
          float4 pn = p;
          if (more) pn = src.load(c, sr, j - 1);  // in flight while p is tested
          const CandD2 d = cand_d2(qx, qy, qz, p);
          if (v.stop(__fadd_rn(d.dx2, rg2))) break;
          v.test(p, j, d.d2);
          if (!more) break;
          p = pn; --j;
        }
      }
      if (ce < row_e) {
        int j = ce;
        float4 p = src.load(c, sr, j);
        for (;;) {
          const bool more = j + 1 < row_e;
          float4 pn = p;
          if (more) pn = src.load(c, sr, j + 1);
          const CandD2 d = cand_d2(qx, qy, qz, p);
          if (v.stop(__fadd_rn(d.dx2, rg2))) break;
          v.test(p, j, d.d2);
          if (!more) break;
          p = pn; ++j;
        }
      }
    }
  }
}
// The rows at Chebyshev distance exactly r.  Every lane walks ITS OWN list of them: a lane keeps stepping along the perimeter
// until it finds a row that can still matter (inside the grid, gap below the current bound) and only then joins the others in
// visit_row — which row that is differs from lane to lane.  (A lock-step double loop over (z, y) runs every slot of the ring for
// the whole warp with the few lanes that need it: ncu showed 7-9 of 32 lanes in the candidate tests.)
template <int MODE, typename V, typename S>
__device__ __forceinline__ void visit_ring(const CloudView& c, const QueryCell& q, float qx, int r, V& v, const S& src) {
  const float h2 = c.h * c.h;
  const int n = 2 * r;
  int side = 0, k = 0;
  for (;;) {
    int y = 0, z = 0;
    float rg2 = 0.f;
    bool found = false;
    while (side < 4) {
      int dy, dz;
      if (side == 0) { dy = k - r; dz = -r; }
      else if (side == 1) { dy = r; dz = k - r; }
      else if (side == 2) { dy = r - k; dz = r; }
      else { dy = -r; dz = r - k; }
      if (k == 0) {  // a side that lies outside the grid as a whole is skipped at once
        const bool out = side == 0 ? q.cz - r < 0 : (side == 1 ? q.cy + r >= c.gd[1] : (side == 2 ? q.cz + r >= c.gd[2] : q.cy - r < 0));
        if (out) { ++side; continue; }
      }
      if (++k == n) { k = 0; ++side; }
      y = q.cy + dy; z = q.cz + dz;
      if (y < 0 || y >= c.gd[1] || z < 0 || z >= c.gd[2]) continue;
      const float gy = axis_gap(q.fy, dy), gz = axis_gap(q.fz, dz);
      rg2 = (gy * gy + gz * gz) * h2;
      if (rg2 + q.xout2 < v.thr()) { found = true; break; }
    }
    if (!found) break;
    visit_row<MODE>(c, q, qx, y, z, rg2, v, src);
  }
}
template <int MODE, typename V, typename S>
__device__ __forceinline__ void visit_rows(const CloudView& c, const QueryCell& q, float qx, int r, bool ring, V& v, const S& src) {
  const int z0 = max(q.cz - r, 0), z1 = min(q.cz + r, c.gd[2] - 1);
  const int y0 = max(q.cy - r, 0), y1 = min(q.cy + r, c.gd[1] - 1);
  const float h2 = c.h * c.h;
  if (ring) {
    if constexpr ((MODE & VISIT_LANE_RING) != 0) {
      visit_ring<MODE>(c, q, qx, r, v, src);
    } else {  // the two z edges, then the two y edges without the corners, the warp in lock step
#pragma unroll 1
      for (int e = 0; e < 2; ++e) {
        const int dz = e ? r : -r, z = q.cz + dz;
        if (z < 0 || z >= c.gd[2]) continue;
        const float gz = axis_gap(q.fz, dz);
        for (int y = y0; y <= y1; ++y) {
          const float gy = axis_gap(q.fy, y - q.cy);
          const float rg2 = (gy * gy + gz * gz) * h2;
          if (!(rg2 + q.xout2 < v.thr())) continue;
          visit_row<MODE>(c, q, qx, y, z, rg2, v, src);
        }
      }
      const int zi0 = max(q.cz - r + 1, 0), zi1 = min(q.cz + r - 1, c.gd[2] - 1);
#pragma unroll 1
      for (int e = 0; e < 2; ++e) {
        const int dy = e ? r : -r, y = q.cy + dy;
        if (y < 0 || y >= c.gd[1]) continue;
        const float gy = axis_gap(q.fy, dy);
        for (int z = zi0; z <= zi1; ++z) {
          const float gz = axis_gap(q.fz, z - q.cz);
          const float rg2 = (gy * gy + gz * gz) * h2;
          if (!(rg2 + q.xout2 < v.thr())) continue;
          visit_row<MODE>(c, q, qx, y, z, rg2, v, src);
        }
      }
    }
    return;
  }
  if ((MODE & VISIT_LANE_RING) != 0 && r == 1) {  // the common case (query inside the grid): its own row, then the ring around it, lane by lane
    if (q.cy >= 0 && q.cy < c.gd[1] && q.cz >= 0 && q.cz < c.gd[2]) visit_row<MODE>(c, q, qx, q.cy, q.cz, 0.f, v, src);
    visit_ring<MODE>(c, q, qx, 1, v, src);
    return;
  }
  const bool own_first = q.cy >= y0 && q.cy <= y1 && q.cz >= z0 && q.cz <= z1;
  if (own_first) visit_row<MODE>(c, q, qx, q.cy, q.cz, 0.f, v, src);
  for (int z = z0; z <= z1; ++z) {
    const int dz = z - q.cz;
    const float gz = axis_gap(q.fz, dz);
    for (int y = y0; y <= y1; ++y) {
      const int dy = y - q.cy;
      if (own_first && dy == 0 && dz == 0) continue;
      const float gy = axis_gap(q.fy, dy);
      const float rg2 = (gy * gy + gz * gz) * h2;
      if (!(rg2 + q.xout2 < v.thr())) continue;
      visit_row<MODE>(c, q, qx, y, z, rg2, v, src);
    }
  }
}
template <typename V, typename S>
__device__ __forceinline__ void visit_rows(const CloudView& c, const QueryCell& q, float qx, int r, bool ring, V& v, const S& src) {
  visit_rows<B2R_VISIT_DEFAULT>(c, q, qx, r, ring, v, src);
}
template <typename V>
__device__ __forceinline__ void visit_rows(const CloudView& c, const QueryCell& q, float qx, int r, bool ring, V& v) {
  visit_rows<B2R_VISIT_DEFAULT>(c, q, qx, r, ring, v, GlobalPts{});
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
  float qx, qy, qz;
  float d[K];
  float lim;  // d[K - 1] * kGapSlack
  __device__ __forceinline__ TopkVisitor(float x, float y, float z) : qx(x), qy(y), qz(z), lim(INFINITY) {
#pragma unroll
    for (int i = 0; i < K; ++i) d[i] = INFINITY;
  }
  __device__ __forceinline__ float thr() const { return d[K - 1]; }
  __device__ __forceinline__ bool stop(float s) const { return s > lim; }
  __device__ __forceinline__ float qy_f() const { return qy; }
  __device__ __forceinline__ float qz_f() const { return qz; }
  __device__ __forceinline__ void test(const float4&, int, float d2) {
    if (d2 < d[K - 1]) { topk_insert<K>(d, d2); lim = d[K - 1] * kGapSlack; }
  }
};

// Same top-K, and every candidate that entered it (or tied with its K-th value) is also logged by position: the first
// CAP in a caller-provided shared-memory column (element t at list[t * stride]), the next SPILL in a per-thread
// local-memory array (rare; L1-resident).  Every point whose distance is <= the final k-th distance was necessarily
// logged when it was seen, in traversal order, so a caller that needs the neighbours themselves re-reads just the
// logged ones instead of traversing the rows a second time.  cnt counts all logged candidates; cnt > CAP + SPILL
// means entries were dropped and the caller must fall back to a second traversal.
template <int K, int CAP, int SPILL>
struct TopkListVisitor {
  float qx, qy, qz;
  float d[K];
  int* list;
  int stride;
  int cnt;
  int* spill;  // SPILL ints of per-thread local memory, kept outside this struct so that d[] stays in registers
  float lim;   // d[K - 1] * kGapSlack
#ifdef B2R_KNN_STATS
  int tested = 0;
#endif
  __device__ __forceinline__ TopkListVisitor(float x, float y, float z, int* l, int s, int* sp)
      : qx(x), qy(y), qz(z), list(l), stride(s), cnt(0), spill(sp), lim(INFINITY) {
#pragma unroll
    for (int i = 0; i < K; ++i) d[i] = INFINITY;
  }
  // tie-inclusive pruning: a candidate at exactly the K-th distance must still be seen (and logged)
  __device__ __forceinline__ float thr() const { return lim; }
  __device__ __forceinline__ bool stop(float s) const { return s > lim; }
  __device__ __forceinline__ float qy_f() const { return qy; }
  __device__ __forceinline__ float qz_f() const { return qz; }
  __device__ __forceinline__ void test(const float4&, int j, float d2) {
#ifdef B2R_KNN_STATS
    ++tested;
#endif
    if (d2 <= d[K - 1]) {
      topk_insert<K>(d, d2);
      lim = d[K - 1] * kGapSlack;
      if (cnt < CAP) list[cnt * stride] = j;
      else if (cnt < CAP + SPILL) spill[cnt - CAP] = j;
      ++cnt;
    }
  }
  __device__ __forceinline__ int logged(int t) const { return t < CAP ? list[t * stride] : spill[t - CAP]; }
};

// Exact k smallest squared distances (k <= K, ascending in v.d[0..k)) of the query in cloud c.
// Returns the (y, z) radius of the scanned rows; v.d[k-1] == INFINITY if the cloud has fewer than k points.
template <int K, typename V, typename S>
__device__ __forceinline__ int knn_topk(const CloudView& c, const QueryCell& q, int k, V& v, const S& src) {
  int r = rows_outside(c, q) + 1;
  visit_rows(c, q, v.qx, r, false, v, src);
  for (;;) {
    if (topk_kth<K>(v.d, k) <= ring_safe_d2(r, c.h) + q.xout2) break;
    if (ring_covers_grid(c, q, r)) break;
    ++r;
    visit_rows(c, q, v.qx, r, true, v, src);
  }
  return r;
}
template <int K, typename V>
__device__ __forceinline__ int knn_topk(const CloudView& c, const QueryCell& q, int k, V& v) {
  return knn_topk<K>(c, q, k, v, GlobalPts{});
}

// ---- same search keeping (distance, position) pairs as 64-bit keys: float bits of d^2 << 32 | position in spts.
// Used where the neighbour identities are wanted in ascending order (debug / test entry point).
__device__ __forceinline__ unsigned long long u64min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long u64max(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

template <int K>
struct TopkKeyVisitor {
  float qx, qy, qz;
  unsigned long long d[K];
  __device__ __forceinline__ TopkKeyVisitor(float x, float y, float z) : qx(x), qy(y), qz(z) {
#pragma unroll
    for (int i = 0; i < K; ++i) d[i] = ~0ull;
  }
  __device__ __forceinline__ float worst() const { return d[K - 1] == ~0ull ? INFINITY : __uint_as_float((unsigned)(d[K - 1] >> 32)); }
  __device__ __forceinline__ float thr() const { return worst() * (1.f + 1e-6f); }  // equal distances still compete on position
  __device__ __forceinline__ bool stop(float s) const { return s > worst() * kGapSlack; }
  __device__ __forceinline__ float qy_f() const { return qy; }
  __device__ __forceinline__ float qz_f() const { return qz; }
  __device__ __forceinline__ void test(const float4&, int j, float d2) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)j;
    if (key < d[K - 1]) {
#pragma unroll
      for (int i = K - 1; i > 0; --i) d[i] = u64max(d[i - 1], u64min(d[i], key));
      d[0] = u64min(d[0], key);
    }
  }
};

// ---- one thread per query: exact 1-NN within max_d2 (pass INFINITY for unbounded).  Returns position in spts or -1.
// Ties go to the lower position.
struct Nn1Visitor {
  float qx, qy, qz, cut;  // points farther than `cut` (squared) cannot matter
  float best;
  int best_pos;
  float lim;  // min(best, cut) * kGapSlack
  __device__ __forceinline__ float thr() const { return lim; }
  __device__ __forceinline__ bool stop(float s) const { return s > lim; }
  __device__ __forceinline__ float qy_f() const { return qy; }
  __device__ __forceinline__ float qz_f() const { return qz; }
  __device__ __forceinline__ void test(const float4&, int j, float d2) {
    if (d2 < best || (d2 == best && j < best_pos)) { best = d2; best_pos = j; lim = fminf(d2, cut) * kGapSlack; }
  }
  __device__ __forceinline__ void test2(float da, float db, int j) {
    if (da < best || (da == best && j < best_pos)) { best = da; best_pos = j; }
    if (db < best || (db == best && j + 1 < best_pos)) { best = db; best_pos = j + 1; }
    lim = fminf(best, cut) * kGapSlack;
  }
};
template <int MODE = B2R_VISIT_DEFAULT>
__device__ __forceinline__ int nn1_search(const CloudView& c, float qx, float qy, float qz, float max_d2, float& best_out) {
  Nn1Visitor v{qx, qy, qz, max_d2, INFINITY, -1, max_d2 * kGapSlack};
  if (c.n == 0) { best_out = v.best; return -1; }
  const QueryCell q = query_cell(c, qx, qy, qz);
  int r = rows_outside(c, q) + 1;
  // nothing closer than (r-2)*h can exist when the query is outside the grid
  {
    const float lb = (float)(r - 2) * c.h;
    if ((r >= 3 && lb * lb > max_d2) || q.xout2 > max_d2) { best_out = v.best; return -1; }
  }
  visit_rows<MODE>(c, q, qx, r, false, v, GlobalPts{});
  for (;;) {
    const float b2 = ring_safe_d2(r, c.h) + q.xout2;
    if (v.best <= b2) break;     // proven nearest
    if (b2 > max_d2) break;      // anything farther is out of range anyway
    if (ring_covers_grid(c, q, r)) break;
    ++r;
    visit_rows<MODE>(c, q, qx, r, true, v, GlobalPts{});
  }
  best_out = v.best;
  return v.best_pos;
}

// ---- the same search when only the DISTANCE is wanted (getFitnessScore, inlier fraction): no position, no tie rule, a
// branch-free candidate test.  Returns the squared distance of the nearest point, INFINITY if the cloud is empty or nothing lies
// within max_d2 of the rows that had to be looked at (a value beyond max_d2 may come back: the caller compares).
struct Nn1DistVisitor {
  float qx, qy, qz, cut;
  float best;
  float lim;  // min(best, cut) * kGapSlack
  __device__ __forceinline__ float thr() const { return lim; }
  __device__ __forceinline__ bool stop(float s) const { return s > lim; }
  __device__ __forceinline__ float qy_f() const { return qy; }
  __device__ __forceinline__ float qz_f() const { return qz; }
  __device__ __forceinline__ void test(const float4&, int, float d2) {
    best = fminf(best, d2);
    lim = fminf(best, cut) * kGapSlack;
  }
  __device__ __forceinline__ void test2(float da, float db, int) {
    best = fminf(best, fminf(da, db));
    lim = fminf(best, cut) * kGapSlack;
  }
};
template <int MODE = B2R_VISIT_DEFAULT>
__device__ __forceinline__ float nn1_dist(const CloudView& c, float qx, float qy, float qz, float max_d2) {
  Nn1DistVisitor v{qx, qy, qz, max_d2, INFINITY, max_d2 * kGapSlack};
  if (c.n == 0) return INFINITY;
  const QueryCell q = query_cell(c, qx, qy, qz);
  int r = rows_outside(c, q) + 1;
  {
    const float lb = (float)(r - 2) * c.h;
    if ((r >= 3 && lb * lb > max_d2) || q.xout2 > max_d2) return INFINITY;
  }
  visit_rows<MODE>(c, q, qx, r, false, v, GlobalPts{});
  for (;;) {
    const float b2 = ring_safe_d2(r, c.h) + q.xout2;
    if (v.best <= b2) break;
    if (b2 > max_d2) break;
    if (ring_covers_grid(c, q, r)) break;
    ++r;
    visit_rows<MODE>(c, q, qx, r, true, v, GlobalPts{});
  }
  return v.best;
}

// ---- one thread per query: exact 1-NN of a DOUBLE query under double squared distance, Eigen Vector4d/Packet2d order
// (dx^2 + dz^2) + dy^2 — small_gicp's kd-tree metric.  The rows are pruned with float arithmetic on the rounded query and a
// margin that covers the rounding of the query (<= 4e-6 m at 64 m) and of the float differences; candidates are compared
// in double.  Ties go to the lower position.  Returns the position in spts or -1 (nothing within max_d2).
struct Nn1VisitorD {
  double qx, qy, qz, best;
  float fx, fy, fz, cutf;  // cutf: float bound on the squared distance of anything that can still win
  int best_pos;
  __device__ __forceinline__ static float bound_of(double d2) {
    const double r = sqrt(d2) + 1e-4;
    return (float)(r * r * (1.0 + 1e-6));
  }
  __device__ __forceinline__ float thr() const { return cutf; }
  __device__ __forceinline__ bool stop(float s) const { return s > cutf; }  // cutf carries a 0.1 mm margin of its own
  __device__ __forceinline__ float qy_f() const { return fy; }
  __device__ __forceinline__ float qz_f() const { return fz; }
  __device__ __forceinline__ void test(const float4& p, int j, float) {
    const double d0 = (double)p.x - qx, d1 = (double)p.y - qy, d2 = (double)p.z - qz;
    const double d = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d2, d2)), __dmul_rn(d1, d1));
    if (d < best || (d == best && j < best_pos)) {
      best = d;
      best_pos = j;
      cutf = fminf(cutf, bound_of(d));
    }
  }
};
__device__ __forceinline__ int nn1_search_d(const CloudView& c, double qx, double qy, double qz, double max_d2, double& best_out) {
  Nn1VisitorD v;
  v.qx = qx; v.qy = qy; v.qz = qz; v.best = INFINITY; v.best_pos = -1;
  v.fx = (float)qx; v.fy = (float)qy; v.fz = (float)qz;
  v.cutf = max_d2 < 1e30 ? Nn1VisitorD::bound_of(max_d2) : INFINITY;
  best_out = v.best;
  if (c.n == 0) return -1;
  const QueryCell q = query_cell(c, v.fx, v.fy, v.fz);
  int r = rows_outside(c, q) + 1;
  {
    const float lb = (float)(r - 2) * c.h;
    if ((r >= 3 && lb * lb > v.cutf) || q.xout2 > v.cutf) return -1;
  }
  visit_rows(c, q, v.fx, r, false, v);
  for (;;) {
    const float b2 = ring_safe_d2(r, c.h) + q.xout2;
    if (v.cutf <= b2) break;  // everything that could still win lies inside the visited rows
    if (ring_covers_grid(c, q, r)) break;
    ++r;
    visit_rows(c, q, v.fx, r, true, v);
  }
  best_out = v.best;
  return (v.best_pos >= 0 && v.best <= max_d2) ? v.best_pos : -1;
}

// ---- one thread per query: number of points with d2 < r2 (strict), early exit once count > stop_above.
// Requires c.h >= radius so that the 3x3 rows around the query's row suffice.
struct RadiusVisitor {
  float qx, qy, qz, r2;
  int cnt, stop_above;
  __device__ __forceinline__ float thr() const { return cnt > stop_above ? 0.f : r2; }
  __device__ __forceinline__ bool stop(float s) const { return s > r2 * kGapSlack || cnt > stop_above; }
  __device__ __forceinline__ float qy_f() const { return qy; }
  __device__ __forceinline__ float qz_f() const { return qz; }
  __device__ __forceinline__ void test(const float4&, int, float d2) {
    if (d2 < r2) ++cnt;
  }
};
__device__ __forceinline__ int radius_count(const CloudView& c, float qx, float qy, float qz, float r2, int stop_above) {
  RadiusVisitor v{qx, qy, qz, r2, 0, stop_above};
  QueryCell q = query_cell(c, qx, qy, qz);
  q.cy = clampi(q.cy, 0, c.gd[1] - 1);
  q.cz = clampi(q.cz, 0, c.gd[2] - 1);
  visit_rows(c, q, qx, 1, false, v);
  return v.cnt;
}

}  // namespace b2r
