// NDT_OMP: derivative kernels over the target voxel grid and the on-device Newton + More-Thuente state machine.
//
// Replaces: A11 pclomp::NormalDistributionsTransform::computeDerivatives / updateDerivatives /
// computePointDerivatives / computeAngleDerivatives / computeHessian and A12 computeTransformation /
// computeStepLengthMT / updateIntervalMT / trialValueSelectionMT (ndt_omp, selected at
// src/mrg_slam/registrations.cpp:130-147).  Inner math is float as in current ndt_omp, accumulation double.
#include <cfloat>
#include <cmath>
#include <algorithm>

#include "internal.hpp"

namespace b2r {

enum { NP_INIT = 0, NP_MT_FIRST = 1, NP_MT_LOOP = 2, NP_MT_HESS = 3, NP_DONE = 4 };
enum { EV_GRAD_HESS = 0, EV_GRAD = 1, EV_HESS = 2 };
constexpr int kNdtAcc = 43;   // score, g(6), H(36)
constexpr int kNdtPart = 44;  // per-block partials: the kNdtAcc sums + the number of voxel hits (measurement only)

struct NdtState {
  double p[6], x_t[6], dir[6], p_eval[6];
  double g[6], H[36], score;
  double phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t, psi_t, d_psi_t;
  float M[16];  // column-major transform used by the pending evaluation (= final_transformation_)
  int interval_converged, open_interval, step_iterations;
  int phase, eval_mode, nr_iterations, converged, evals;
  double work_pts, work_hits;  // source points / voxel hits processed over all evaluations (algorithmic-bytes accounting)
};

struct NdtParams {
  int neighbor_search;
  double gauss_d1, gauss_d2;
  double step_size, trans_eps;
  int max_iterations;
};

// ---- angle derivative tables (computeAngleDerivatives), double -> float ----
struct AngTables { float j[8][3]; float h[15][3]; };

__device__ void ndt_angle_tables(const double* p, AngTables& t) {
  double cx, cy, cz, sx, sy, sz;
  if (fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = cos(p[3]); sx = sin(p[3]); }
  if (fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = cos(p[4]); sy = sin(p[4]); }
  if (fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = cos(p[5]); sz = sin(p[5]); }
  const double J[8][3] = {{-sx * sz + cx * sy * cz, -sx * cz - cx * sy * sz, -cx * cy},
                          {cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy},
                          {-sy * cz, sy * sz, cy},
                          {sx * cy * cz, -sx * cy * sz, sx * sy},
                          {-cx * cy * cz, cx * cy * sz, -cx * sy},
                          {-cy * sz, -cy * cz, 0},
                          {cx * cz - sx * sy * sz, -cx * sz - sx * sy * cz, 0},
                          {sx * cz + cx * sy * sz, cx * sy * cz - sx * sz, 0}};
  const double Hh[15][3] = {{-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, sx * cy},
                            {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, -cx * cy},
                            {cx * cy * cz, -cx * cy * sz, cx * sy},
                            {sx * cy * cz, -sx * cy * sz, sx * sy},
                            {-sx * cz - cx * sy * sz, sx * sz - cx * sy * cz, 0},
                            {cx * cz - sx * sy * sz, -sx * sy * cz - cx * sz, 0},
                            {-cy * cz, cy * sz, sy},
                            {-sx * sy * cz, sx * sy * sz, sx * cy},
                            {cx * sy * cz, -cx * sy * sz, -cx * cy},
                            {sy * sz, sy * cz, 0},
                            {-sx * cy * sz, -sx * cy * cz, 0},
                            {cx * cy * sz, cx * cy * cz, 0},
                            {-cy * cz, cy * sz, 0},
                            {-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, 0},
                            {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, 0}};
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 3; ++c) t.j[r][c] = (float)J[r][c];
  for (int r = 0; r < 15; ++r)
    for (int c = 0; c < 3; ++c) t.h[r][c] = (float)Hh[r][c];
}

__device__ __forceinline__ float dot3f(const float* row, float x, float y, float z) {
  float s = __fmul_rn(row[0], x);
  s = __fadd_rn(s, __fmul_rn(row[1], y));
  s = __fadd_rn(s, __fmul_rn(row[2], z));
  return s;
}
__device__ __forceinline__ float mad3(float a0, float b0, float a1, float b1, float a2, float b2) {
  float s = __fmul_rn(a0, b0);
  s = __fadd_rn(s, __fmul_rn(a1, b1));
  s = __fadd_rn(s, __fmul_rn(a2, b2));
  return s;
}

__device__ __forceinline__ float mad2(float a1, float b1, float a2, float b2) {
  return __fadd_rn(__fmul_rn(a1, b1), __fmul_rn(a2, b2));
}

// updateDerivatives for one (point, leaf) hit.  xt: transformed point; pg / ph: the point's computePointDerivatives rows.
template <bool SPLIT>
__device__ __forceinline__ void ndt_hit(double* acc, double* sacc, const NdtRec& L, float xt0, float xt1, float xt2, const float* pg, const float* ph,
                                        float gd2, double gd1, bool do_grad, bool do_hess) {
  // updateDerivatives (float inner math)
  const float x0 = (float)((double)xt0 - __ldg(&L.mean[0]));
  const float x1 = (float)((double)xt1 - __ldg(&L.mean[1]));
  const float x2 = (float)((double)xt2 - __ldg(&L.mean[2]));
  float C[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) C[t] = __ldg(&L.icov[t]);
  const float xC0 = mad3(x0, C[0], x1, C[3], x2, C[6]);
  const float xC1 = mad3(x0, C[1], x1, C[4], x2, C[7]);
  const float xC2 = mad3(x0, C[2], x1, C[5], x2, C[8]);
  const float xCx = mad3(x0, xC0, x1, xC1, x2, xC2);
  float e = (float)exp((double)__fmul_rn(__fmul_rn(-gd2, xCx), 0.5f));
  const float score_inc = (float)(-gd1 * (double)e);
  e = __fmul_rn(gd2, e);
  if (e > 1.0f || e < 0.0f || e != e) return;
  e = (float)((double)e * gd1);
  acc[0] += (double)score_inc;
  // point_gradient (3x6): columns 0..2 identity; col 3 = (0, pg0, pg1); col 4 = (pg2, pg3, pg4); col 5 = (pg5, pg6, pg7).
  // The reference multiplies through the full matrices in float; the products with the structural 1 and 0 entries
  // are exact (x*1 = x, x*0 = +-0, y + +-0 = y), so they are skipped here and every remaining operation is the
  // reference's own, in its order: cg(:,c) = c_inv * point_gradient.col(c) is column c of c_inv for c < 3, etc.
  float cg[3][6];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    cg[a][0] = C[a * 3 + 0]; cg[a][1] = C[a * 3 + 1]; cg[a][2] = C[a * 3 + 2];
    cg[a][3] = mad2(C[a * 3 + 1], pg[0], C[a * 3 + 2], pg[1]);
    cg[a][4] = mad3(C[a * 3 + 0], pg[2], C[a * 3 + 1], pg[3], C[a * 3 + 2], pg[4]);
    cg[a][5] = mad3(C[a * 3 + 0], pg[5], C[a * 3 + 1], pg[6], C[a * 3 + 2], pg[7]);
  }
  float xcg[6] = {xC0, xC1, xC2, 0.f, 0.f, 0.f};  // x . cg(:,c); for c < 3 this is x . c_inv(:,c), computed above
#pragma unroll
  for (int c = 3; c < 6; ++c) xcg[c] = mad3(x0, cg[0][c], x1, cg[1][c], x2, cg[2][c]);
  if (do_grad) {
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[1 + c] += (double)__fmul_rn(e, xcg[c]);
  }
  if (do_hess) {
    // point_hessian blocks (first component of a, b, c is structurally 0):
    // (3,3)=a (4,3)=b (5,3)=c (3,4)=b (4,4)=d (5,4)=e (3,5)=c (4,5)=e (5,5)=f
    const float xa = mad2(xC1, ph[0], xC2, ph[1]);
    const float xb = mad2(xC1, ph[2], xC2, ph[3]);
    const float xc = mad2(xC1, ph[4], xC2, ph[5]);
    const float xd = mad3(xC0, ph[6], xC1, ph[7], xC2, ph[8]);
    const float xe = mad3(xC0, ph[9], xC1, ph[10], xC2, ph[11]);
    const float xf = mad3(xC0, ph[12], xC1, ph[13], xC2, ph[14]);
    const float XH[3][3] = {{xa, xb, xc}, {xb, xd, xe}, {xc, xe, xf}};
#pragma unroll
    for (int i2_ = 0; i2_ < 6; ++i2_) {
      const float ti = __fmul_rn(-gd2, xcg[i2_]);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        // G = point_gradient.col(j) . cg(:,i)
        float G;
        if (j < 3) G = cg[j][i2_];
        else if (j == 3) G = mad2(pg[0], cg[1][i2_], pg[1], cg[2][i2_]);
        else if (j == 4) G = mad3(pg[2], cg[0][i2_], pg[3], cg[1][i2_], pg[4], cg[2][i2_]);
        else G = mad3(pg[5], cg[0][i2_], pg[6], cg[1][i2_], pg[7], cg[2][i2_]);
        float v = __fmul_rn(ti, xcg[j]);
        if (i2_ >= 3 && j >= 3) v = __fadd_rn(v, XH[i2_ - 3][j - 3]);  // elsewhere the block of point_hessian is 0
        v = __fadd_rn(v, G);
        if (SPLIT && i2_ >= 3) sacc[((i2_ - 3) * 6 + j) * 128] += (double)__fmul_rn(e, v);  // stride = kNdtThreads
        else acc[7 + i2_ * 6 + j] += (double)__fmul_rn(e, v);
      }
    }
  }
}

// grid = (chunks, pairs).  Evaluates score / gradient / Hessian of the pending transform of each active pair.
constexpr int kNdtThreads = 128;
__global__ void __launch_bounds__(kNdtThreads) ndt_eval_kernel(const CloudView* __restrict__ views, const PairDesc* __restrict__ pairs,
                                                        const NdtState* __restrict__ states, NdtParams prm, double* __restrict__ partials,
                                                        int32_t* __restrict__ hits_out) {
  const int pair = blockIdx.y;
  const NdtState& st = states[pair];
  if (st.phase == NP_DONE) return;
  const int mode = st.eval_mode;
  const bool do_grad = mode != EV_HESS, do_hess = mode != EV_GRAD;
  const CloudView& src = views[pairs[pair].src];
  const CloudView& tgt = views[pairs[pair].tgt];
  __shared__ AngTables tab;
  __shared__ float T[16];
  __shared__ double red[kNdtAcc * 4];
  __shared__ int s_rec[27 * kNdtThreads];  // usable leaves of each thread's current point (<= 27 with DIRECT27)
  if (threadIdx.x == 0) ndt_angle_tables(st.p_eval, tab);
  if (threadIdx.x < 16) T[threadIdx.x] = st.M[threadIdx.x];
  __syncthreads();
  const float gd2 = (float)prm.gauss_d2;
  const double gd1 = prm.gauss_d1;
  double acc[kNdtAcc];
#pragma unroll
  for (int t = 0; t < kNdtAcc; ++t) acc[t] = 0.0;
  const int noff = prm.neighbor_search == B2R_DIRECT1 ? 1 : (prm.neighbor_search == B2R_DIRECT7 ? 7 : 27);
  const bool have_grid = tgt.ncell_ndt > 0;
  // KDTREE (pclomp::KDTREE, registrations.cpp:140-141; also what pcl::NormalDistributionsTransform does): radius search over the
  // leaves' float centroids, FLANN L2_Simple distance < (float)(resolution^2).  A centroid lies inside its own cell, so the 27
  // cells around the query's hold every centroid within one leaf size: probe them, keep those that pass the distance test.
  const bool kdtree = prm.neighbor_search == B2R_KDTREE;
  const float kd_r2 = (float)((double)tgt.leaf * (double)tgt.leaf);
  int nhits = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < src.n; i += gridDim.x * blockDim.x) {
    const float4 xo = __ldg(&src.pts[i]);
    // transformPointCloud, PCL association (x*c0 + y*c1) + (z*c2 + c3)
    const float xt0 = __fadd_rn(__fadd_rn(__fmul_rn(xo.x, T[0]), __fmul_rn(xo.y, T[4])), __fadd_rn(__fmul_rn(xo.z, T[8]), T[12]));
    const float xt1 = __fadd_rn(__fadd_rn(__fmul_rn(xo.x, T[1]), __fmul_rn(xo.y, T[5])), __fadd_rn(__fmul_rn(xo.z, T[9]), T[13]));
    const float xt2 = __fadd_rn(__fadd_rn(__fmul_rn(xo.x, T[2]), __fmul_rn(xo.y, T[6])), __fadd_rn(__fmul_rn(xo.z, T[10]), T[14]));
    // lookup key: float DIVISION by the leaf size (getNeighborhoodAtPoint)
    const int i0 = (int)floorf(__fdiv_rn(xt0, tgt.leaf)), i1 = (int)floorf(__fdiv_rn(xt1, tgt.leaf)), i2 = (int)floorf(__fdiv_rn(xt2, tgt.leaf));
    // getNeighborhoodAtPoint: collect the usable leaves first (same order as the reference visits them), then process
    // them hit by hit — the lanes of a warp then work on their k-th hit together instead of idling through the
    // neighbour offsets where only some of them have a leaf
    int hits = 0;
    for (int o = 0; o < noff && have_grid; ++o) {
      int ox, oy, oz;
      neighbor_offset(prm.neighbor_search, o, ox, oy, oz);
      const int c0 = i0 + ox, c1 = i1 + oy, c2 = i2 + oz;
      if (c0 < tgt.min_b[0] || c0 > tgt.max_b[0] || c1 < tgt.min_b[1] || c1 > tgt.max_b[1] || c2 < tgt.min_b[2] || c2 > tgt.max_b[2]) continue;
      const int idx = (c0 - tgt.min_b[0]) + (c1 - tgt.min_b[1]) * tgt.div_b[0] + (c2 - tgt.min_b[2]) * tgt.div_b[0] * tgt.div_b[1];
      const int rec = __ldg(&tgt.n_table[idx]);
      if (rec < 0) continue;
      if (__ldg(&tgt.nrec[rec].n) < 6) continue;
      if (kdtree) {
        const float* cen = tgt.nrec[rec].centroid;
        if (!(dist2_flann(xt0, xt1, xt2, __ldg(&cen[0]), __ldg(&cen[1]), __ldg(&cen[2])) < kd_r2)) continue;
      }
      s_rec[hits * kNdtThreads + threadIdx.x] = rec;
      ++hits;
    }
    float pg[8];   // point_gradient entries (1,3) (2,3) (0,4) (1,4) (2,4) (0,5) (1,5) (2,5)
    float ph[15];  // a(1,2) b(1,2) c(1,2) d(0..2) e(0..2) f(0..2)
    if (hits > 0) {  // computePointDerivatives (float), once per point
#pragma unroll
      for (int r = 0; r < 8; ++r) pg[r] = dot3f(tab.j[r], xo.x, xo.y, xo.z);
      if (do_hess) {
#pragma unroll
        for (int r = 0; r < 15; ++r) ph[r] = dot3f(tab.h[r], xo.x, xo.y, xo.z);
      }
    }
    for (int hh = 0; hh < hits; ++hh) {
      ndt_hit<false>(acc, nullptr, tgt.nrec[s_rec[hh * kNdtThreads + threadIdx.x]], xt0, xt1, xt2, pg, ph, gd2, gd1, do_grad, do_hess);
    }
    if (hits_out) hits_out[i] = hits;
    nhits += hits;
  }
  double* out = partials + ((size_t)pair * gridDim.x + blockIdx.x) * kNdtPart;
  block_reduce_to<kNdtAcc>(acc, red, out);
  const int bh = block_sum_int(nhits, (int*)red);
  if (threadIdx.x == 0) out[43] = (double)bh;
}

// Same evaluation with the (point, leaf) hits compacted across the block.  A point has 0..7 usable leaves (DIRECT7), so in
// ndt_eval_kernel the lanes of a warp idle while the one with the most hits finishes.  Here every round of kNdtThreads
// points only LISTS its hits: a block-wide exclusive scan of the hit counts gives each thread the slots of a ring buffer in
// shared memory (order = point order, then the reference's neighbour order: deterministic), and the hits are then processed
// kNdtThreads at a time, one per thread, whichever point they belong to; what does not fill a round waits for the next
// points.  The point's own derivative rows are recomputed per hit (23 dot products against ~800 instructions per hit).
constexpr int kNdtQueueOff = 7;  // DIRECT1 / DIRECT7 (registrations.cpp:140-146); DIRECT27 keeps ndt_eval_kernel
constexpr int kNdtQueue = 1024;   // >= (kNdtThreads - 1) left over + kNdtQueueOff * kNdtThreads new hits
__global__ void __launch_bounds__(kNdtThreads, 4) ndt_eval_queue_kernel(const CloudView* __restrict__ views, const PairDesc* __restrict__ pairs,
                                                              const NdtState* __restrict__ states, NdtParams prm,
                                                              double* __restrict__ partials, int32_t* __restrict__ hits_out) {
  const int pair = blockIdx.y;
  const NdtState& st = states[pair];
  if (st.phase == NP_DONE) return;
  const int mode = st.eval_mode;
  const bool do_grad = mode != EV_HESS, do_hess = mode != EV_GRAD;
  const CloudView& src = views[pairs[pair].src];
  const CloudView& tgt = views[pairs[pair].tgt];
  __shared__ AngTables tab;
  __shared__ float T[16];
  __shared__ double red[kNdtAcc * 4];
  __shared__ int2 s_q[kNdtQueue];  // (source point, leaf record)
  __shared__ int s_wsum[kNdtThreads / 32];
  // rows 3..5 of the Hessian are accumulated in shared memory (one column per thread): 36 registers less, which is what
  // lets a fourth block (16 warps) fit on the SM
  __shared__ double s_hess[18 * kNdtThreads];
  if (threadIdx.x == 0) ndt_angle_tables(st.p_eval, tab);
  if (threadIdx.x < 16) T[threadIdx.x] = st.M[threadIdx.x];
  __syncthreads();
  const float gd2 = (float)prm.gauss_d2;
  const double gd1 = prm.gauss_d1;
  double acc[25];  // score, g(6), H rows 0..2
#pragma unroll
  for (int t = 0; t < 25; ++t) acc[t] = 0.0;
#pragma unroll
  for (int t = 0; t < 18; ++t) s_hess[t * kNdtThreads + threadIdx.x] = 0.0;
  const int noff = prm.neighbor_search == B2R_DIRECT1 ? 1 : 7;
  const bool have_grid = tgt.ncell_ndt > 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int nhits = 0;
  int head = 0, tail = 0;  // ring positions (block-uniform)
  const int n = src.n;
  for (int base = blockIdx.x * blockDim.x; base < n || tail - head > 0; base += gridDim.x * blockDim.x) {
    const bool more = base < n;
    if (more) {
      // ---- list the usable leaves of this round's points
      const int i = base + threadIdx.x;
      int recs[kNdtQueueOff];
#pragma unroll
      for (int o = 0; o < kNdtQueueOff; ++o) recs[o] = -1;  // threads past the end of the cloud list nothing
      int hits = 0;
      if (i < n && have_grid) {
        const float4 xo = __ldg(&src.pts[i]);
        const float xt0 = __fadd_rn(__fadd_rn(__fmul_rn(xo.x, T[0]), __fmul_rn(xo.y, T[4])), __fadd_rn(__fmul_rn(xo.z, T[8]), T[12]));
        const float xt1 = __fadd_rn(__fadd_rn(__fmul_rn(xo.x, T[1]), __fmul_rn(xo.y, T[5])), __fadd_rn(__fmul_rn(xo.z, T[9]), T[13]));
        const float xt2 = __fadd_rn(__fadd_rn(__fmul_rn(xo.x, T[2]), __fmul_rn(xo.y, T[6])), __fadd_rn(__fmul_rn(xo.z, T[10]), T[14]));
        const int i0 = (int)floorf(__fdiv_rn(xt0, tgt.leaf)), i1 = (int)floorf(__fdiv_rn(xt1, tgt.leaf)), i2 = (int)floorf(__fdiv_rn(xt2, tgt.leaf));
        // probe all neighbour cells first (independent loads), then the leaves' point counts
#pragma unroll
        for (int o = 0; o < kNdtQueueOff; ++o) {
          recs[o] = -1;
          if (o < noff) {
            int ox, oy, oz;
            neighbor_offset(prm.neighbor_search, o, ox, oy, oz);
            const int c0 = i0 + ox, c1 = i1 + oy, c2 = i2 + oz;
            if (!(c0 < tgt.min_b[0] || c0 > tgt.max_b[0] || c1 < tgt.min_b[1] || c1 > tgt.max_b[1] || c2 < tgt.min_b[2] || c2 > tgt.max_b[2]))
              recs[o] = __ldg(&tgt.n_table[(c0 - tgt.min_b[0]) + (c1 - tgt.min_b[1]) * tgt.div_b[0] + (c2 - tgt.min_b[2]) * tgt.div_b[0] * tgt.div_b[1]]);
          }
        }
#pragma unroll
        for (int o = 0; o < kNdtQueueOff; ++o)
          if (o < noff && recs[o] >= 0 && __ldg(&tgt.nrec[recs[o]].n) < 6) recs[o] = -1;
#pragma unroll
        for (int o = 0; o < kNdtQueueOff; ++o) hits += (o < noff && recs[o] >= 0) ? 1 : 0;
        if (hits_out) hits_out[i] = hits;
      }
      nhits += hits;
      // ---- block-wide exclusive scan of the hit counts
      int incl = hits;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      int off = incl - hits, total = 0;
#pragma unroll
      for (int w = 0; w < kNdtThreads / 32; ++w) {
        const int ws = s_wsum[w];
        if (w < warp) off += ws;
        total += ws;
      }
      int slot = tail + off;
#pragma unroll
      for (int o = 0; o < kNdtQueueOff; ++o)
        if (o < noff && recs[o] >= 0) { s_q[slot & (kNdtQueue - 1)] = make_int2(i, recs[o]); ++slot; }
      tail += total;
      __syncthreads();
    }
    // ---- process full rounds (and the remainder once the points are exhausted)
    const bool last = base + (int)(gridDim.x * blockDim.x) >= n;
    while (tail - head >= kNdtThreads || (last && tail - head > 0)) {
      const int pos = head + threadIdx.x;
      if (pos < tail) {
        const int2 it = s_q[pos & (kNdtQueue - 1)];
        const float4 xo = __ldg(&src.pts[it.x]);
        const float xt0 = __fadd_rn(__fadd_rn(__fmul_rn(xo.x, T[0]), __fmul_rn(xo.y, T[4])), __fadd_rn(__fmul_rn(xo.z, T[8]), T[12]));
        const float xt1 = __fadd_rn(__fadd_rn(__fmul_rn(xo.x, T[1]), __fmul_rn(xo.y, T[5])), __fadd_rn(__fmul_rn(xo.z, T[9]), T[13]));
        const float xt2 = __fadd_rn(__fadd_rn(__fmul_rn(xo.x, T[2]), __fmul_rn(xo.y, T[6])), __fadd_rn(__fmul_rn(xo.z, T[10]), T[14]));
        float pg[8], ph[15];
#pragma unroll
        for (int r = 0; r < 8; ++r) pg[r] = dot3f(tab.j[r], xo.x, xo.y, xo.z);
        if (do_hess) {
#pragma unroll
          for (int r = 0; r < 15; ++r) ph[r] = dot3f(tab.h[r], xo.x, xo.y, xo.z);
        }
        ndt_hit<true>(acc, s_hess + threadIdx.x, tgt.nrec[it.y], xt0, xt1, xt2, pg, ph, gd2, gd1, do_grad, do_hess);
      }
      head = min(head + kNdtThreads, tail);
    }
    if (last) { __syncthreads(); }
    if (!more) break;
  }
  double* out = partials + ((size_t)pair * gridDim.x + blockIdx.x) * kNdtPart;
  double all[kNdtAcc];
#pragma unroll
  for (int t = 0; t < 25; ++t) all[t] = acc[t];
#pragma unroll
  for (int t = 0; t < 18; ++t) all[25 + t] = s_hess[t * kNdtThreads + threadIdx.x];
  block_reduce_to<kNdtAcc>(all, red, out);
  const int bh = block_sum_int(nhits, (int*)red);
  if (threadIdx.x == 0) out[43] = (double)bh;
}

static bool ndt_use_queue() {
  static const bool q = [] { const char* e = getenv("B2R_NDT_QUEUE"); return !e || atoi(e) != 0; }();
  return q;
}

// ---- step-kernel helpers (one thread per pair) ----
__device__ void svd6_solve_dev(const double* Ain, const double* rhs, double* x) {
  double U[36], V[36];
  for (int i = 0; i < 36; ++i) { U[i] = Ain[i]; V[i] = 0.0; }
  for (int i = 0; i < 6; ++i) V[i * 6 + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 5; ++p)
      for (int q = p + 1; q < 6; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 6; ++i) {
          alpha += U[i * 6 + p] * U[i * 6 + p];
          beta += U[i * 6 + q] * U[i * 6 + q];
          gamma += U[i * 6 + p] * U[i * 6 + q];
        }
        if (gamma == 0.0 || fabs(gamma) <= 1e-15 * sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 6; ++i) {
          const double up = U[i * 6 + p], uq = U[i * 6 + q];
          U[i * 6 + p] = c * up - s * uq;
          U[i * 6 + q] = s * up + c * uq;
          const double vp = V[i * 6 + p], vq = V[i * 6 + q];
          V[i * 6 + p] = c * vp - s * vq;
          V[i * 6 + q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sig[6], smax = 0.0;
  for (int j = 0; j < 6; ++j) {
    double s = 0;
    for (int i = 0; i < 6; ++i) s += U[i * 6 + j] * U[i * 6 + j];
    sig[j] = sqrt(s);
    smax = fmax(smax, sig[j]);
  }
  const double thr = 2.220446049250313e-16 * 6 * smax;
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    if (!(sig[j] > thr)) continue;
    double proj = 0;
    for (int i = 0; i < 6; ++i) proj += U[i * 6 + j] * rhs[i];
    proj /= (sig[j] * sig[j]);
    for (int i = 0; i < 6; ++i) x[i] += V[i * 6 + j] * proj;
  }
}

// Translation3f(p0..2) * AngleAxisf(p3,X) * AngleAxisf(p4,Y) * AngleAxisf(p5,Z) in float, column-major out.
__host__ __device__ inline void ndt_matrix_from_p(const double* p, float* M) {
  const float rx = (float)p[3], ry = (float)p[4], rz = (float)p[5];
  const float cx = (float)cos((double)rx), sx = (float)sin((double)rx);
  const float cy = (float)cos((double)ry), sy = (float)sin((double)ry);
  const float cz = (float)cos((double)rz), sz = (float)sin((double)rz);
  const float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  const float Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  const float Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  float A[9], B[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
#ifdef __CUDA_ARCH__
      float s = __fmul_rn(Rx[i * 3 + 0], Ry[0 * 3 + j]);
      s = __fadd_rn(s, __fmul_rn(Rx[i * 3 + 1], Ry[1 * 3 + j]));
      s = __fadd_rn(s, __fmul_rn(Rx[i * 3 + 2], Ry[2 * 3 + j]));
#else
      float s = Rx[i * 3 + 0] * Ry[0 * 3 + j];
      s = s + Rx[i * 3 + 1] * Ry[1 * 3 + j];
      s = s + Rx[i * 3 + 2] * Ry[2 * 3 + j];
#endif
      A[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
#ifdef __CUDA_ARCH__
      float s = __fmul_rn(A[i * 3 + 0], Rz[0 * 3 + j]);
      s = __fadd_rn(s, __fmul_rn(A[i * 3 + 1], Rz[1 * 3 + j]));
      s = __fadd_rn(s, __fmul_rn(A[i * 3 + 2], Rz[2 * 3 + j]));
#else
      float s = A[i * 3 + 0] * Rz[0 * 3 + j];
      s = s + A[i * 3 + 1] * Rz[1 * 3 + j];
      s = s + A[i * 3 + 2] * Rz[2 * 3 + j];
#endif
      B[i * 3 + j] = s;
    }
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) M[c * 4 + r] = B[r * 3 + c];
    M[12 + r] = (float)p[r];
    M[r * 4 + 3] = 0.f;
  }
  M[15] = 1.f;
}

__device__ double mt_psi(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
__device__ double mt_dpsi(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

__device__ bool mt_update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t,
                                   double g_t) {
  if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
  if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  return true;
}
__device__ double mt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    const double w = sqrt(z * z - g_t * g_l);
    const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    return (fabs(a_c - a_l) < fabs(a_q - a_l)) ? a_c : 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {
    const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    const double w = sqrt(z * z - g_t * g_l);
    const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    return (fabs(a_c - a_t) >= fabs(a_s - a_t)) ? a_c : a_s;
  } else if (fabs(g_t) <= fabs(g_l)) {
    const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    const double w = sqrt(z * z - g_t * g_l);
    const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    const double a_t_next = (fabs(a_c - a_t) < fabs(a_s - a_t)) ? a_c : a_s;
    return (a_t > a_l) ? fmin(a_t + 0.66 * (a_u - a_t), a_t_next) : fmax(a_t + 0.66 * (a_u - a_t), a_t_next);
  }
  const double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
  const double w = sqrt(z * z - g_t * g_u);
  return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
}

constexpr double kMtMu = 1.e-4, kMtNu = 0.9;
constexpr int kMtMaxIter = 10;

__device__ void ndt_request_eval(NdtState& s, int mode, int phase) {
  for (int i = 0; i < 6; ++i) s.p_eval[i] = s.x_t[i];
  s.eval_mode = mode;
  s.phase = phase;
}
__device__ void ndt_set_trial(NdtState& s, const NdtParams& prm) {
  s.a_t = fmin(s.a_t, prm.step_size);
  s.a_t = fmax(s.a_t, prm.trans_eps / 2);
  for (int i = 0; i < 6; ++i) s.x_t[i] = s.p[i] + s.dir[i] * s.a_t;
  ndt_matrix_from_p(s.x_t, s.M);
}

// returns true when the alignment finished
__device__ bool ndt_newton_start(NdtState& s, const NdtParams& prm);

__device__ bool ndt_mt_end(NdtState& s, const NdtParams& prm) {
  for (int i = 0; i < 6; ++i) s.p[i] += s.dir[i] * s.a_t;
  if (s.nr_iterations > prm.max_iterations || (s.nr_iterations && (fabs(s.a_t) < prm.trans_eps))) s.converged = 1;
  s.nr_iterations++;
  if (s.converged) return true;
  return ndt_newton_start(s, prm);
}

__device__ bool ndt_newton_start(NdtState& s, const NdtParams& prm) {
  for (;;) {
    double ng[6], delta[6];
    for (int i = 0; i < 6; ++i) ng[i] = -s.g[i];
    svd6_solve_dev(s.H, ng, delta);
    double nrm = 0;
    for (int i = 0; i < 6; ++i) nrm += delta[i] * delta[i];
    nrm = sqrt(nrm);
    if (nrm == 0 || nrm != nrm) {
      s.converged = (nrm == nrm) ? 1 : 0;
      return true;
    }
    for (int i = 0; i < 6; ++i) s.dir[i] = delta[i] / nrm;
    // computeStepLengthMT prologue
    s.phi_0 = -s.score;
    double d = 0;
    for (int i = 0; i < 6; ++i) d += s.g[i] * s.dir[i];
    s.d_phi_0 = -d;
    if (s.d_phi_0 >= 0) {
      if (s.d_phi_0 == 0) {
        // step length 0: no evaluation; p unchanged.  Same bookkeeping as ndt_mt_end without recursion.
        s.a_t = 0.0;
        if (s.nr_iterations > prm.max_iterations || (s.nr_iterations && (0.0 < prm.trans_eps))) s.converged = 1;
        s.nr_iterations++;
        if (s.converged) return true;
        continue;
      }
      s.d_phi_0 = -s.d_phi_0;
      for (int i = 0; i < 6; ++i) s.dir[i] = -s.dir[i];
    }
    s.step_iterations = 0;
    s.a_l = 0; s.a_u = 0;
    s.f_l = mt_psi(s.a_l, s.phi_0, s.phi_0, s.d_phi_0, kMtMu);
    s.g_l = mt_dpsi(s.d_phi_0, s.d_phi_0, kMtMu);
    s.f_u = mt_psi(s.a_u, s.phi_0, s.phi_0, s.d_phi_0, kMtMu);
    s.g_u = mt_dpsi(s.d_phi_0, s.d_phi_0, kMtMu);
    s.interval_converged = (prm.step_size - prm.trans_eps / 2) < 0;
    s.open_interval = 1;
    s.a_t = nrm;
    ndt_set_trial(s, prm);
    ndt_request_eval(s, EV_GRAD_HESS, NP_MT_FIRST);
    return false;
  }
}

__device__ bool ndt_mt_continue(NdtState& s, const NdtParams& prm) {
  if (!s.interval_converged && s.step_iterations < kMtMaxIter && !(s.psi_t <= 0 && s.d_phi_t <= -kMtNu * s.d_phi_0)) {
    if (s.open_interval) s.a_t = mt_trial_value(s.a_l, s.f_l, s.g_l, s.a_u, s.f_u, s.g_u, s.a_t, s.psi_t, s.d_psi_t);
    else s.a_t = mt_trial_value(s.a_l, s.f_l, s.g_l, s.a_u, s.f_u, s.g_u, s.a_t, s.phi_t, s.d_phi_t);
    ndt_set_trial(s, prm);
    ndt_request_eval(s, EV_GRAD, NP_MT_LOOP);
    return false;
  }
  if (s.step_iterations) {
    ndt_request_eval(s, EV_HESS, NP_MT_HESS);  // computeHessian at x_t
    return false;
  }
  return ndt_mt_end(s, prm);
}

__device__ void ndt_step_body(NdtState* __restrict__ states, int npairs, const NdtParams& prm, const double* __restrict__ partials, int chunks,
                              const int* __restrict__ src_n, int* __restrict__ done_count) {
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pair >= npairs) return;
  NdtState& s = states[pair];
  const int phase = s.phase;
  if (phase == NP_DONE) return;
  // fixed-order sum of the chunk partials: lane l sums elements l and l+32
  double v0 = 0.0, v1 = 0.0;
  {
    const double* p = partials + (size_t)pair * chunks * kNdtPart;
    for (int c = 0; c < chunks; ++c) {
      v0 += p[(size_t)c * kNdtPart + lane];
      if (lane + 32 < kNdtPart) v1 += p[(size_t)c * kNdtPart + lane + 32];
    }
  }
  const double nhits = __shfl_sync(0xffffffffu, v1, 43 - 32);
  double a[kNdtAcc];
#pragma unroll
  for (int t = 0; t < kNdtAcc; ++t) a[t] = t < 32 ? __shfl_sync(0xffffffffu, v0, t) : __shfl_sync(0xffffffffu, v1, t - 32);
  if (lane != 0) return;
  s.evals++;
  s.work_pts += (double)src_n[pair];
  s.work_hits += nhits;
  const int mode = s.eval_mode;
  if (mode != EV_HESS) {
    s.score = a[0];
    for (int i = 0; i < 6; ++i) s.g[i] = a[1 + i];
  }
  // computeDerivatives zeroes the Hessian even when it does not compute it
  for (int i = 0; i < 36; ++i) s.H[i] = (mode == EV_GRAD) ? 0.0 : a[7 + i];
  bool finished = false;
  if (phase == NP_INIT) {
    finished = ndt_newton_start(s, prm);
  } else if (phase == NP_MT_FIRST || phase == NP_MT_LOOP) {
    s.phi_t = -s.score;
    double d = 0;
    for (int i = 0; i < 6; ++i) d += s.g[i] * s.dir[i];
    s.d_phi_t = -d;
    s.psi_t = mt_psi(s.a_t, s.phi_t, s.phi_0, s.d_phi_0, kMtMu);
    s.d_psi_t = mt_dpsi(s.d_phi_t, s.d_phi_0, kMtMu);
    if (phase == NP_MT_LOOP) {
      if (s.open_interval && (s.psi_t <= 0 && s.d_psi_t >= 0)) {
        s.open_interval = 0;
        s.f_l = s.f_l + s.phi_0 - kMtMu * s.d_phi_0 * s.a_l; s.g_l = s.g_l + kMtMu * s.d_phi_0;
        s.f_u = s.f_u + s.phi_0 - kMtMu * s.d_phi_0 * s.a_u; s.g_u = s.g_u + kMtMu * s.d_phi_0;
      }
      if (s.open_interval) s.interval_converged = mt_update_interval(s.a_l, s.f_l, s.g_l, s.a_u, s.f_u, s.g_u, s.a_t, s.psi_t, s.d_psi_t) ? 1 : 0;
      else s.interval_converged = mt_update_interval(s.a_l, s.f_l, s.g_l, s.a_u, s.f_u, s.g_u, s.a_t, s.phi_t, s.d_phi_t) ? 1 : 0;
      s.step_iterations++;
    }
    finished = ndt_mt_continue(s, prm);
  } else {  // NP_MT_HESS
    finished = ndt_mt_end(s, prm);
  }
  if (finished) {
    s.phase = NP_DONE;
    atomicAdd(done_count, 1);
  }
}

__global__ void ndt_step_kernel(NdtState* __restrict__ states, NdtParams prm, const double* __restrict__ partials, int chunks,
                                const int* __restrict__ src_n, LoopArgs la) {
  ndt_step_body(states, la.npairs, prm, partials, chunks, src_n, &la.ctl->done);
  loop_tail(la);
}
__global__ void ndt_rows_kernel(const NdtState* __restrict__ states, int npairs, b2r_result* __restrict__ rows) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= npairs) return;
  const NdtState& s = states[pair];
  b2r_result& r = rows[pair];
  for (int t = 0; t < 16; ++t) r.T[t] = s.M[t];
  r.converged = s.converged;
  r.iterations = s.nr_iterations;
  r.error = s.score;
  r.evals = s.evals;
  clear_row_padding(r);
  r.fitness = 0.0;
}

// Eigen::Matrix3f::eulerAngles(0,1,2) (Eigen >= 3.3), float
static void euler_angles_012(const float* R /*row-major*/, float* res) {
  auto c = [&](int r, int cc) { return R[r * 3 + cc]; };
  res[0] = std::atan2(c(1, 2), c(2, 2));
  const float c2 = std::sqrt(c(0, 0) * c(0, 0) + c(0, 1) * c(0, 1));
  if (res[0] > 0.f) {
    res[0] -= (float)M_PI;
    res[1] = std::atan2(-c(0, 2), -c2);
  } else {
    res[1] = std::atan2(-c(0, 2), c2);
  }
  const float s1 = std::sin(res[0]), c1 = std::cos(res[0]);
  res[2] = std::atan2(s1 * c(2, 0) - c1 * c(1, 0), c1 * c(1, 1) - s1 * c(2, 1));
  res[0] = -res[0]; res[1] = -res[1]; res[2] = -res[2];
}

static NdtParams make_ndt_params(const b2r_config& cfg) {
  NdtParams p;
  p.neighbor_search = cfg.neighbor_search;
  const double res = (double)(float)cfg.resolution;
  const double c1 = 10 * (1 - cfg.ndt_outlier_ratio);
  const double c2 = cfg.ndt_outlier_ratio / std::pow(res, 3);
  const double d3 = -std::log(c2);
  p.gauss_d1 = -std::log(c1 + c2) - d3;
  p.gauss_d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / p.gauss_d1);
  p.step_size = cfg.ndt_step_size;
  p.trans_eps = cfg.transformation_epsilon;
  p.max_iterations = cfg.maximum_iterations;
  return p;
}

static int ndt_chunks(const Ctx& ctx, int npairs, int maxn) {
  int by_size = std::max(1, (maxn + 511) / 512);
  static const int fill = [] { const char* e = getenv("B2R_FILL_PER_SM"); return e ? atoi(e) : 48; }();  // blocks per SM a launch should offer: short tails when few pairs are active
  int by_fill = std::max(1, (fill * ctx.num_sms + npairs - 1) / npairs);
  static const int tail_pts = [] { const char* e = getenv("B2R_CHUNK_POINTS"); return e ? atoi(e) : 8192; }();
  by_fill = std::max(by_fill, (maxn + tail_pts - 1) / tail_pts);  // see pick_chunks (lsq.cu): bounds the tail of large batches
  return std::max(1, std::min(by_size, by_fill));
}

static void ndt_init_state(NdtState& s, const float* g) {
  memset(&s, 0, sizeof(s));
  bool ident = true;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      if (g[c * 4 + r] != (r == c ? 1.f : 0.f)) ident = false;
  // align(): final_transformation_ = I; computeTransformation: if (guess != I) final_transformation_ = guess
  for (int i = 0; i < 16; ++i) s.M[i] = ident ? ((i % 5 == 0) ? 1.f : 0.f) : g[i];
  float Rf[9], eul[3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Rf[r * 3 + c] = s.M[c * 4 + r];
  euler_angles_012(Rf, eul);
  s.p[0] = s.M[12]; s.p[1] = s.M[13]; s.p[2] = s.M[14];
  s.p[3] = eul[0]; s.p[4] = eul[1]; s.p[5] = eul[2];
  for (int i = 0; i < 6; ++i) { s.p_eval[i] = s.p[i]; s.x_t[i] = s.p[i]; }
  s.phase = NP_INIT;
  s.eval_mode = EV_GRAD_HESS;
}

void ndt_align_batch(Ctx& ctx, const b2r_config& cfg, const BatchArgs& b) {
  const int np = b.np;
  if (np == 0) return;
  const int chunks = ndt_chunks(ctx, np, b.maxn);
  NdtParams prm = make_ndt_params(cfg);
  std::vector<NdtState> hs(np);
  for (int i = 0; i < np; ++i) ndt_init_state(hs[i], b.guesses + (size_t)i * 16);
  DBuf<NdtState> ds; ds.alloc(np, ctx.stream);
  DBuf<double> part; part.alloc((size_t)np * chunks * kNdtPart, ctx.stream);
  DBuf<LoopCtl> ctl; ctl.alloc(1, ctx.stream);
  ctl.zero(ctx.stream);
  B2R_CUDA(cudaMemcpyAsync(ds.p, hs.data(), sizeof(NdtState) * np, cudaMemcpyHostToDevice, ctx.stream));
  // each outer iteration evaluates at most 1 + 10 + 1 times
  const long max_rounds = 2 + (long)(std::max(0, cfg.maximum_iterations) + 3) * (kMtMaxIter + 2);
  const CloudView* a_views = b.d_views;
  const PairDesc* a_pairs = b.d_pairs;
  const NdtState* a_states_c = ds.p;
  NdtState* a_states = ds.p;
  double* a_part = part.p;
  const double* a_part_c = part.p;
  int32_t* a_null = nullptr;
  int a_chunks = chunks;
  const int* a_src_n = b.d_src_n;
  LoopArgs la;
  memset(&la, 0, sizeof(la));
  void* eval_args[] = {&a_views, &a_pairs, &a_states_c, &prm, &a_part, &a_null};
  void* step_args[] = {&a_states, &prm, &a_part_c, &a_chunks, &a_src_n, &la};
  const bool queue = ndt_use_queue() && (prm.neighbor_search == B2R_DIRECT1 || prm.neighbor_search == B2R_DIRECT7);
  run_device_loop(ctx, queue ? (const void*)ndt_eval_queue_kernel : (const void*)ndt_eval_kernel, dim3(chunks, np), dim3(kNdtThreads), eval_args,
                  (const void*)ndt_step_kernel, dim3((np + 3) / 4), dim3(128), step_args, la, ctl.p, np, max_rounds, PROF_NDT_EVAL);
  B2R_LAUNCH(ctx, ndt_rows_kernel, (np + 127) / 128, 128, 0, ds.p, np, b.d_rows);
  if (ctx.profile) {
    // per derivative pass 16 B per source point + 64 B per voxel hit, over the passes each pair ran
    B2R_CUDA(cudaMemcpyAsync(hs.data(), ds.p, sizeof(NdtState) * np, cudaMemcpyDeviceToHost, ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(ctx.stream));
    double wp = 0.0, wh = 0.0;
    for (int i = 0; i < np; ++i) { wp += hs[i].work_pts; wh += hs[i].work_hits; }
    ctx.prof_bytes[PROF_NDT_EVAL] += 16.0 * wp + 64.0 * wh;
  }
}

void ndt_debug_derivatives(Ctx& ctx, const b2r_config& cfg, const CloudView* d_views, int n_src, const double* p6, double* score,
                           double* grad6, double* hess36, int32_t* hits_out) {
  const NdtParams prm = make_ndt_params(cfg);
  NdtState s;
  memset(&s, 0, sizeof(s));
  for (int i = 0; i < 6; ++i) { s.p[i] = p6[i]; s.p_eval[i] = p6[i]; }
  ndt_matrix_from_p(p6, s.M);
  s.phase = NP_INIT;
  s.eval_mode = EV_GRAD_HESS;
  const int chunks = ndt_chunks(ctx, 1, n_src);
  PairDesc pd{0, 1};
  DBuf<PairDesc> dp; dp.alloc(1, ctx.stream);
  DBuf<NdtState> ds; ds.alloc(1, ctx.stream);
  DBuf<double> part; part.alloc((size_t)chunks * kNdtPart, ctx.stream);
  DBuf<int32_t> dh;
  if (hits_out) dh.alloc(n_src, ctx.stream);
  B2R_CUDA(cudaMemcpyAsync(dp.p, &pd, sizeof(pd), cudaMemcpyHostToDevice, ctx.stream));
  B2R_CUDA(cudaMemcpyAsync(ds.p, &s, sizeof(s), cudaMemcpyHostToDevice, ctx.stream));
  if (ndt_use_queue() && (prm.neighbor_search == B2R_DIRECT1 || prm.neighbor_search == B2R_DIRECT7)) B2R_LAUNCH(ctx, ndt_eval_queue_kernel, dim3(chunks, 1), 128, 0, d_views, dp.p, ds.p, prm, part.p, hits_out ? dh.p : nullptr);
  else B2R_LAUNCH(ctx, ndt_eval_kernel, dim3(chunks, 1), 128, 0, d_views, dp.p, ds.p, prm, part.p, hits_out ? dh.p : nullptr);
  std::vector<double> hp((size_t)chunks * kNdtPart);
  B2R_CUDA(cudaMemcpyAsync(hp.data(), part.p, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost, ctx.stream));
  if (hits_out) B2R_CUDA(cudaMemcpyAsync(hits_out, dh.p, sizeof(int32_t) * n_src, cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  double a[kNdtAcc] = {0};
  for (int c = 0; c < chunks; ++c)
    for (int t = 0; t < kNdtAcc; ++t) a[t] += hp[(size_t)c * kNdtPart + t];
  *score = a[0];
  for (int i = 0; i < 6; ++i) grad6[i] = a[1 + i];
  for (int i = 0; i < 36; ++i) hess36[i] = a[7 + i];
}

}  // namespace b2r
