// pcl::GeneralizedIterativeClosestPoint ("GICP" / "GICP_OMP": /root/reference/src/mrg_slam/registrations.cpp:93-116, SURVEY 8a
// row G): the two kernels that answer the requests of the state machine in gicp_pcl_sm.hpp, and the batch driver.
//
//   gicp_pcl_eval_kernel   grid (chunks, pairs).  REQ_CORRESPOND: q = transformation_ * (guess * p) in float (Eigen's product
//                          order), exact 1-NN in the target (knn.cuh), kept if d2 < max_correspondence_distance^2; the position is
//                          stored per source point, the block counts its correspondences.  REQ_EVAL: the functor's sums at x —
//                          d = T(x) (guess p) - q_target (float differences widened), M = (R C_src R^T + C_tgt)^-1 with
//                          R = transformation_ * guess (double), f += d^T M d, g_t += M d, D += (guess p) (M d)^T — 13 doubles
//                          per block, fixed-order reduction.
//   gicp_pcl_step_kernel   one thread per pair: sums the block partials in a fixed order, turns them into f and g
//                          (computeRDerivative) and advances the pair's state machine to its next request.
// The control flow (outer loop, BFGS, line search) is gp::advance, verified on the host against the oracle bit for bit
// (tests/test_gicp_pcl_sm.py); the covariances are PCL's own (cloud.cu: pcl_cov_kernel).
#include <cfloat>
#include <cmath>
#include <algorithm>
#include <vector>

#include "internal.hpp"
#include "knn.cuh"
#include "gicp_pcl_sm.hpp"

namespace b2r {

constexpr int kGpAcc = 13;   // fs, gt(3), D(9)
constexpr int kGpPart = 14;  // + correspondences counted by a REQ_CORRESPOND pass

// Matrix4f (column-major) * (x, y, z, 1): ((c0 x + c1 y) + c2 z) + c3, no contraction
__device__ __forceinline__ void m4f_point_dev(const float* T, float x, float y, float z, float& ox, float& oy, float& oz) {
  float o[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float s = __fmul_rn(T[0 * 4 + r], x);
    s = __fadd_rn(s, __fmul_rn(T[1 * 4 + r], y));
    s = __fadd_rn(s, __fmul_rn(T[2 * 4 + r], z));
    o[r] = __fadd_rn(s, T[3 * 4 + r]);
  }
  ox = o[0]; oy = o[1]; oz = o[2];
}

__global__ void __launch_bounds__(256, 2) gicp_pcl_eval_kernel(const CloudView* __restrict__ views, const PairDesc* __restrict__ pairs,
                                                             const gp::State* __restrict__ states, double corr_thr2, float corr_max_d2,
                                                             double* __restrict__ partials, int32_t* __restrict__ corr,
                                                             const long long* __restrict__ corr_off) {
  const int pair = blockIdx.y;
  const gp::State& st = states[pair];
  const int req = st.request;
  if (req != gp::REQ_CORRESPOND && req != gp::REQ_EVAL) return;
  const CloudView& src = views[pairs[pair].src];
  const CloudView& tgt = views[pairs[pair].tgt];
  __shared__ float s_guess[16], s_trans[16], s_T[16];
  __shared__ double s_R[9];
  __shared__ double red[kGpAcc * 8];
  if (threadIdx.x < 16) { s_guess[threadIdx.x] = st.guess[threadIdx.x]; s_trans[threadIdx.x] = st.transformation[threadIdx.x]; }
  if (threadIdx.x == 0 && req == gp::REQ_EVAL) gp::apply_state_identity(st.xreq, s_T);
  __syncthreads();
  if (threadIdx.x < 9) {  // transform_R = transformation_ * guess, rotation block, double
    const int i = threadIdx.x / 3, j = threadIdx.x % 3;
    double s = 0;
    for (int k = 0; k < 4; ++k) s += (double)s_trans[k * 4 + i] * (double)s_guess[j * 4 + k];
    s_R[threadIdx.x] = s;
  }
  __syncthreads();
  int32_t* cc = corr + corr_off[pair];
  double acc[kGpAcc];
#pragma unroll
  for (int t = 0; t < kGpAcc; ++t) acc[t] = 0.0;
  int ncorr = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < src.n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(&src.pts[i]);
    float ox, oy, oz;  // `output` = guess * input (pcl::transformPointCloud)
    m4f_point_dev(s_guess, p.x, p.y, p.z, ox, oy, oz);
    if (req == gp::REQ_CORRESPOND) {
      float qx, qy, qz;
      m4f_point_dev(s_trans, ox, oy, oz, qx, qy, qz);
      float d2;
      const int pos = nn1_search(tgt, qx, qy, qz, corr_max_d2, d2);
      const bool ok = pos >= 0 && (double)d2 < corr_thr2;
      cc[i] = ok ? pos : -1;
      ncorr += ok ? 1 : 0;
      continue;
    }
    const int pos = cc[i];
    if (pos < 0) continue;
    const float4 tq = __ldg(&tgt.spts[pos]);
    // Mahalanobis matrix of the correspondence: (R C1 R^T + C2)^-1
    double C1[6], RCR[6], S[6], M[6];
    {
      const double* pc = src.cov + (size_t)i * 6;
      const double* pt = tgt.cov + (size_t)__float_as_int(tq.w) * 6;
#pragma unroll
      for (int t = 0; t < 6; ++t) C1[t] = __ldg(&pc[t]);
      rsrt(s_R, C1, RCR);
#pragma unroll
      for (int t = 0; t < 6; ++t) S[t] = RCR[t] + __ldg(&pt[t]);
      sym3_inverse(S, M);
    }
    float ex, ey, ez;
    m4f_point_dev(s_T, ox, oy, oz, ex, ey, ez);
    const double d0 = (double)__fsub_rn(ex, tq.x), d1 = (double)__fsub_rn(ey, tq.y), d2 = (double)__fsub_rn(ez, tq.z);
    const double Md0 = M[0] * d0 + M[1] * d1 + M[2] * d2;
    const double Md1 = M[1] * d0 + M[3] * d1 + M[4] * d2;
    const double Md2 = M[2] * d0 + M[4] * d1 + M[5] * d2;
    acc[0] += d0 * Md0 + d1 * Md1 + d2 * Md2;
    acc[1] += Md0; acc[2] += Md1; acc[3] += Md2;
    const double px = (double)ox, py = (double)oy, pz = (double)oz;  // p_base_src (base_transformation_ = identity)
    acc[4] += px * Md0; acc[5] += px * Md1; acc[6] += px * Md2;
    acc[7] += py * Md0; acc[8] += py * Md1; acc[9] += py * Md2;
    acc[10] += pz * Md0; acc[11] += pz * Md1; acc[12] += pz * Md2;
  }
  double* out = partials + ((size_t)pair * gridDim.x + blockIdx.x) * kGpPart;
  if (req == gp::REQ_CORRESPOND) {
    const int bc = block_sum_int(ncorr, (int*)red);
    if (threadIdx.x == 0) out[13] = (double)bc;
  } else {
    block_reduce_to<kGpAcc>(acc, red, out);
  }
}

__device__ void gicp_pcl_step_body(gp::State* __restrict__ states, int npairs, const gp::Params& prm, const double* __restrict__ partials,
                                   int chunks, int* __restrict__ m_arr, int* __restrict__ done_count) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= npairs) return;
  gp::State s = states[pair];
  if (s.request == gp::REQ_DONE) return;
  const double* p = partials + (size_t)pair * chunks * kGpPart;
  if (s.request == gp::REQ_NONE) {
    gp::advance(s, prm, 0, 0.0, nullptr);  // the first request
  } else if (s.request == gp::REQ_CORRESPOND) {
    double m = 0;
    for (int c = 0; c < chunks; ++c) m += p[(size_t)c * kGpPart + 13];
    m_arr[pair] = (int)m;
    gp::advance(s, prm, (int)m, 0.0, nullptr);
  } else {
    double a[kGpAcc];
    for (int t = 0; t < kGpAcc; ++t) {
      double v = 0;
      for (int c = 0; c < chunks; ++c) v += p[(size_t)c * kGpPart + t];
      a[t] = v;
    }
    double f, g[6];
    gp::finish_fdf(s.xreq, a[0], a + 1, a + 4, m_arr[pair], &f, g);
    gp::advance(s, prm, 0, f, g);
  }
  if (s.request == gp::REQ_DONE) atomicAdd(done_count, 1);
  states[pair] = s;
}

__global__ void gicp_pcl_step_kernel(gp::State* __restrict__ states, gp::Params prm, const double* __restrict__ partials, int chunks,
                                     int* __restrict__ m_arr, LoopArgs la) {
  gicp_pcl_step_body(states, la.npairs, prm, partials, chunks, m_arr, &la.ctl->done);
  loop_tail(la);
}
// the first request of every pair (REQ_NONE -> REQ_CORRESPOND), before the loop
__global__ void gicp_pcl_first_kernel(gp::State* __restrict__ states, int npairs, gp::Params prm, const double* __restrict__ partials, int chunks,
                                      int* __restrict__ m_arr, int* __restrict__ done_count) {
  gicp_pcl_step_body(states, npairs, prm, partials, chunks, m_arr, done_count);
}
__global__ void gicp_pcl_rows_kernel(const gp::State* __restrict__ states, int npairs, b2r_result* __restrict__ rows) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= npairs) return;
  const gp::State& s = states[pair];
  b2r_result& r = rows[pair];
  gp::final_transformation(s, r.T);
  r.converged = s.request == gp::REQ_DONE ? s.converged : 0;
  r.iterations = s.nr_iterations;
  r.error = s.f;
  r.evals = s.evals;
  clear_row_padding(r);
  r.fitness = 0.0;
}

void gicp_pcl_align_batch(Ctx& ctx, const b2r_config& cfg, const BatchArgs& b) {
  const int np = b.np;
  if (np == 0) return;
  const int chunks = std::max(1, std::min((b.maxn + 1023) / 1024, std::max(1, (32 * ctx.num_sms + np - 1) / np)));
  gp::Params prm;
  gp::default_params(prm);
  prm.transformation_epsilon = cfg.transformation_epsilon;
  prm.rotation_epsilon = cfg.rotation_epsilon;
  prm.maximum_iterations = cfg.maximum_iterations;
  prm.max_optimizer_iterations = cfg.max_optimizer_iterations;
  double thr2 = cfg.max_correspondence_distance * cfg.max_correspondence_distance;
  float max_d2 = thr2 >= (double)FLT_MAX ? INFINITY : (float)(thr2 * 1.0001);
  std::vector<gp::State> hs(np);
  for (int i = 0; i < np; ++i) gp::init(hs[i], b.guesses + (size_t)i * 16);
  std::vector<long long> hoff(np);
  long long tot = 0;
  for (int i = 0; i < np; ++i) { hoff[i] = tot; tot += b.src_sizes[i]; }
  DBuf<gp::State> ds; ds.alloc(np, ctx.stream);
  DBuf<double> part; part.alloc((size_t)np * chunks * kGpPart, ctx.stream);
  DBuf<int32_t> corr; corr.alloc((size_t)std::max(1ll, tot), ctx.stream);
  DBuf<long long> coff; coff.alloc(np, ctx.stream);
  DBuf<int> marr; marr.alloc(np, ctx.stream);
  DBuf<LoopCtl> ctl; ctl.alloc(1, ctx.stream);
  ctl.zero(ctx.stream);
  B2R_CUDA(cudaMemsetAsync(marr.p, 0, sizeof(int) * np, ctx.stream));
  B2R_CUDA(cudaMemsetAsync(part.p, 0, sizeof(double) * (size_t)np * chunks * kGpPart, ctx.stream));
  B2R_CUDA(cudaMemcpyAsync(ds.p, hs.data(), sizeof(gp::State) * np, cudaMemcpyHostToDevice, ctx.stream));
  B2R_CUDA(cudaMemcpyAsync(coff.p, hoff.data(), sizeof(long long) * np, cudaMemcpyHostToDevice, ctx.stream));
  const int step_blocks = (np + 63) / 64;
  // the first step only issues the first request (REQ_NONE -> REQ_CORRESPOND)
  B2R_LAUNCH(ctx, gicp_pcl_first_kernel, step_blocks, 64, 0, ds.p, np, prm, part.p, chunks, marr.p, &ctl.p->done);
  // per outer iteration: 1 correspondence pass + per BFGS step (<= max_optimizer_iterations) a line search of a few evaluations
  const long max_rounds = 4 + (long)std::max(1, cfg.maximum_iterations) * (2 + (long)std::max(1, cfg.max_optimizer_iterations) * 210);
  const CloudView* a_views = b.d_views;
  const PairDesc* a_pairs = b.d_pairs;
  const gp::State* a_states_c = ds.p;
  gp::State* a_states = ds.p;
  double* a_part = part.p;
  const double* a_part_c = part.p;
  int32_t* a_corr = corr.p;
  const long long* a_coff = coff.p;
  int* a_marr = marr.p;
  int a_chunks = chunks;
  LoopArgs la;
  memset(&la, 0, sizeof(la));
  void* eval_args[] = {&a_views, &a_pairs, &a_states_c, &thr2, &max_d2, &a_part, &a_corr, &a_coff};
  void* step_args[] = {&a_states, &prm, &a_part_c, &a_chunks, &a_marr, &la};
  run_device_loop(ctx, (const void*)gicp_pcl_eval_kernel, dim3(chunks, np), dim3(256), eval_args, (const void*)gicp_pcl_step_kernel,
                  dim3(step_blocks), dim3(64), step_args, la, ctl.p, np, max_rounds, -1);
  B2R_LAUNCH(ctx, gicp_pcl_rows_kernel, (np + 127) / 128, 128, 0, ds.p, np, b.d_rows);
}

}  // namespace b2r
