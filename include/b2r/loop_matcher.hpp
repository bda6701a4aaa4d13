// Header-only C++ mirror of LoopDetector::matching + perform_loop_closure_consistency_check
// (/root/reference/src/mrg_slam/loop_detector.cpp:97-303) on top of the libb2r C ABI.
//
// The candidate loop of :126-145 (K independent aligns + getFitnessScore against one target) becomes ONE
// b2r_align_batch call; the consistency check's extra aligns (:235-241, :278-284) go through the same call.  Everything
// around them is restated statement by statement: normalize_estimate (:182-188), the initial guess (:129-133), the
// best-candidate rule (`!hasConverged() || score > best_score -> skip`, :138), the Matrix4f identity checks (:247-250,
// :290-296) and the acceptance rules (:156-166).  No Eigen / PCL dependency: 4x4 matrices are column-major arrays, the
// layout of Eigen::Matrix4f / Isometry3d::matrix(), so `.data()` of the reference's objects can be passed straight in.
//
// mrg_slam_b200/loop_closure.py holds the same logic for many new keyframes at once (sharded over GPUs);
// tests/test_loop_consistency.py checks that one against a sequential restatement, tests/cpp/host_mirror_test.cpp this one.
#pragma once
#include <array>
#include <cfloat>
#include <cmath>
#include <vector>

#include "../b2r.h"

namespace b2r {

using Mat4d = std::array<double, 16>;  // column-major
using Mat4f = std::array<float, 16>;   // column-major

namespace detail {
template <typename S>
inline S at(const std::array<S, 16>& M, int r, int c) { return M[c * 4 + r]; }

// Eigen::Quaternion<S>(rotation block of M): (w, x, y, z), not normalised
template <typename S>
inline void quat_from_matrix(const std::array<S, 16>& M, S q[4]) {
  const S t = at(M, 0, 0) + at(M, 1, 1) + at(M, 2, 2);
  if (t > S(0)) {
    S r = std::sqrt(t + S(1));
    q[0] = S(0.5) * r;
    r = S(0.5) / r;
    q[1] = (at(M, 2, 1) - at(M, 1, 2)) * r;
    q[2] = (at(M, 0, 2) - at(M, 2, 0)) * r;
    q[3] = (at(M, 1, 0) - at(M, 0, 1)) * r;
    return;
  }
  int i = 0;
  if (at(M, 1, 1) > at(M, 0, 0)) i = 1;
  if (at(M, 2, 2) > at(M, i, i)) i = 2;
  const int j = (i + 1) % 3, k = (i + 2) % 3;
  S r = std::sqrt(at(M, i, i) - at(M, j, j) - at(M, k, k) + S(1));
  q[1 + i] = S(0.5) * r;
  r = S(0.5) / r;
  q[0] = (at(M, k, j) - at(M, j, k)) * r;
  q[1 + j] = (at(M, j, i) + at(M, i, j)) * r;
  q[1 + k] = (at(M, k, i) + at(M, i, k)) * r;
}

inline Mat4f mul(const Mat4f& A, const Mat4f& B) {
  Mat4f C{};
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += A[k * 4 + r] * B[c * 4 + k];
      C[c * 4 + r] = s;
    }
  return C;
}

// general 4x4 inverse (Matrix4f::inverse()), cofactor expansion in float
inline Mat4f inverse(const Mat4f& m) {
  Mat4f inv;
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  const float id = 1.f / det;
  for (float& v : inv) v *= id;
  return inv;
}

inline Mat4f to_float(const Mat4d& M) {
  Mat4f F;
  for (int i = 0; i < 16; ++i) F[i] = (float)M[i];
  return F;
}
}  // namespace detail

// LoopDetector::normalize_estimate (:182-188): rotation -> Quaterniond -> normalized() -> toRotationMatrix()
inline Mat4d normalize_estimate(const Mat4d& T) {
  double q[4];
  detail::quat_from_matrix<double>(T, q);
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
  Mat4d o = T;
  o[0] = 1 - 2 * (y * y + z * z); o[4] = 2 * (x * y - w * z);     o[8] = 2 * (x * z + w * y);
  o[1] = 2 * (x * y + w * z);     o[5] = 1 - 2 * (x * x + z * z); o[9] = 2 * (y * z - w * x);
  o[2] = 2 * (x * z - w * y);     o[6] = 2 * (y * z + w * x);     o[10] = 1 - 2 * (x * x + y * y);
  return o;
}

// ( new_keyframe_estimate.inverse() * other_estimate ).matrix().cast<float>() (:129-133, :235-240, :278-283)
inline Mat4f registration_guess(const Mat4d& new_estimate, const Mat4d& other_estimate, bool planar = false) {
  const Mat4d a = normalize_estimate(new_estimate), b = normalize_estimate(other_estimate);
  Mat4d inv{};  // Isometry inverse: [R^T | -R^T t]
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) inv[c * 4 + r] = a[r * 4 + c];
  for (int r = 0; r < 3; ++r) inv[12 + r] = -(inv[0 * 4 + r] * a[12] + inv[1 * 4 + r] * a[13] + inv[2 * 4 + r] * a[14]);
  inv[15] = 1.0;
  Mat4d g{};
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += inv[k * 4 + r] * b[c * 4 + k];
      g[c * 4 + r] = s;
    }
  Mat4f f = detail::to_float(g);
  if (planar) f[14] = 0.f;  // guess(2, 3) = 0
  return f;
}

// what the matching reads from a KeyFrame (keyframe.hpp): cloud, graph estimate, flags and the odometry edges
struct KeyframeRef {
  b2r_cloud* cloud = nullptr;            // device-resident copy of KeyFrame::cloud (b2r_cloud_create)
  Mat4d estimate{};                      // node->estimate().matrix()
  bool first_keyframe = false, static_keyframe = false;
  const KeyframeRef* prev = nullptr;     // prev_edge->to_keyframe, or nullptr
  Mat4d rel_pose_to_prev{};              // prev_edge->relative_pose().matrix()
  const KeyframeRef* next = nullptr;     // next_edge->from_keyframe, or nullptr
  Mat4d rel_pose_from_next{};            // next_edge->relative_pose().matrix()
};

struct LoopMatchParams {  // config/mrg_slam.yaml:172-179
  double fitness_score_max_range = DBL_MAX;  // .inf in the YAML: PCL compares it with the squared distance
  double fitness_score_thresh = 1.25;
  bool use_planar_registration_guess = false;
  bool enable_loop_closure_consistency_check = true;
  double loop_closure_consistency_max_delta_trans = 0.3;
  double loop_closure_consistency_max_delta_angle = 0.0523599;
};

struct LoopMatch {
  b2r_status status = B2R_OK;
  int best = -1;                  // index into the candidate list, -1: no converged candidate
  double best_score = DBL_MAX;
  Mat4f rel_pose_new_to_best{};   // getFinalTransformation() of the best candidate
  bool consistency_passed = false;
  bool loop_found = false;        // what matching() returns a Loop for
  int aligns = 0;                 // registrations run (candidates + consistency)
  float delta_trans[2] = {-1.f, -1.f}, delta_angle[2] = {-1.f, -1.f};  // [0] prev check, [1] next check; -1: not run
};

namespace detail {
// translation norm and Quaternionf(rotation).angularDistance(Identity) = 2 atan2(|vec|, |w|)
inline void identity_delta(const Mat4f& M, float& dtrans, float& dangle) {
  dtrans = std::sqrt(M[12] * M[12] + M[13] * M[13] + M[14] * M[14]);
  float q[4];
  quat_from_matrix<float>(M, q);
  dangle = 2.f * std::atan2(std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), std::fabs(q[0]));
}
}  // namespace detail

// LoopDetector::matching (:97-180) for one new keyframe.  `h` is the loop detector's registration handle.
inline LoopMatch match_keyframe(b2r_handle* h, const KeyframeRef& new_keyframe, const std::vector<const KeyframeRef*>& candidates,
                                const LoopMatchParams& prm = LoopMatchParams()) {
  LoopMatch out;
  if (candidates.empty()) return out;  // :99-101
  const size_t K = candidates.size();
  std::vector<b2r_cloud*> src(K), tgt(K, new_keyframe.cloud);  // setInputTarget( new_keyframe->cloud ) (:104)
  std::vector<float> guesses(K * 16);
  for (size_t i = 0; i < K; ++i) {
    src[i] = candidates[i]->cloud;  // setInputSource( candidate->cloud ) (:127)
    const Mat4f g = registration_guess(new_keyframe.estimate, candidates[i]->estimate, prm.use_planar_registration_guess);
    for (int t = 0; t < 16; ++t) guesses[i * 16 + t] = g[t];
  }
  std::vector<b2r_result> res(K);
  out.status = b2r_align_batch(h, src.data(), tgt.data(), guesses.data(), K, /*with_fitness=*/1, prm.fitness_score_max_range, res.data());
  out.aligns = (int)K;
  if (out.status != B2R_OK) return out;  // PCL style: nothing converged, no loop
  for (size_t i = 0; i < K; ++i) {       // :137-144
    if (!res[i].converged || res[i].fitness > out.best_score) continue;
    out.best_score = res[i].fitness;
    out.best = (int)i;
    for (int t = 0; t < 16; ++t) out.rel_pose_new_to_best[t] = res[i].T[t];
  }
  // ---- perform_loop_closure_consistency_check (:190-218)
  const KeyframeRef* best = out.best >= 0 ? candidates[out.best] : nullptr;
  auto one_align = [&](const KeyframeRef* other, Mat4f& T) {  // :233-244, :276-287 (converged flag not consulted)
    const Mat4f g = registration_guess(new_keyframe.estimate, other->estimate, prm.use_planar_registration_guess);
    b2r_cloud* s = other->cloud;
    b2r_cloud* t = new_keyframe.cloud;
    b2r_result r;
    const b2r_status st = b2r_align_batch(h, &s, &t, g.data(), 1, 0, 0.0, &r);
    ++out.aligns;
    for (int i = 0; i < 16; ++i) T[i] = st == B2R_OK ? r.T[i] : g[i];
    return st;
  };
  auto check = [&]() -> bool {
    if (best && (best->first_keyframe || best->static_keyframe)) return true;  // :197-199
    if (!best || !prm.enable_loop_closure_consistency_check || out.best_score > prm.fitness_score_thresh) return false;  // :201-204
    if (best->prev) {  // check_consistency_with_prev_keyframe (:220-262)
      Mat4f T_new_prev;
      one_align(best->prev, T_new_prev);
      const Mat4f M = detail::mul(detail::mul(detail::inverse(T_new_prev), out.rel_pose_new_to_best), detail::to_float(best->rel_pose_to_prev));
      detail::identity_delta(M, out.delta_trans[0], out.delta_angle[0]);
      if (!(out.delta_trans[0] > prm.loop_closure_consistency_max_delta_trans || out.delta_angle[0] > prm.loop_closure_consistency_max_delta_angle))
        return true;
    }
    if (!best->next) return false;  // check_consistency_with_next_keyframe (:264-303)
    Mat4f T_new_next;
    one_align(best->next, T_new_next);
    const Mat4f M = detail::mul(detail::mul(detail::inverse(out.rel_pose_new_to_best), T_new_next), detail::to_float(best->rel_pose_from_next));
    detail::identity_delta(M, out.delta_trans[1], out.delta_angle[1]);
    return !(out.delta_trans[1] > prm.loop_closure_consistency_max_delta_trans || out.delta_angle[1] > prm.loop_closure_consistency_max_delta_angle);
  };
  out.consistency_passed = check();
  if (out.best_score > prm.fitness_score_thresh) return out;  // :156-160 (also covers best == nullptr: DBL_MAX)
  if (prm.enable_loop_closure_consistency_check && best && !best->first_keyframe && !out.consistency_passed) return out;  // :162-166
  out.loop_found = best != nullptr;
  return out;
}


// ------------------------------------------------------------------------------------------------------------------------------
// LoopDetector::matching (:97-180) for MANY new keyframes at once, sharded over the ranks of a b2r_comm (one process or thread per
// GPU).  Every rank calls this with the same keyframe lists; a KeyframeRef's `cloud` only has to be valid on the rank
// that b2r_partition_by_target assigns its target to (`cloud_of` is asked for exactly those).  Stages, as the reference orders them
// per keyframe: (1) all candidate aligns + getFitnessScore as ONE sharded batch, best-candidate rule on the gathered table
// (b2r_select_best_candidates); (2) perform_loop_closure_consistency_check: the `prev` aligns of all surviving loops as a second
// (much smaller) sharded batch, then the `next` aligns of those that failed or have no prev edge as a third; (3) acceptance
// (:156-166).  Every rank returns the same decisions.
struct KeyframeJob {
  const KeyframeRef* new_keyframe = nullptr;
  std::vector<const KeyframeRef*> candidates;
};

inline std::vector<LoopMatch> match_keyframes_sharded(b2r_handle* h, b2r_comm* comm, const std::vector<KeyframeJob>& jobs,
                                                      const LoopMatchParams& prm = LoopMatchParams()) {
  std::vector<LoopMatch> out(jobs.size());
  const int rank = b2r_comm_rank(comm), nranks = b2r_comm_size(comm);
  // ---- stage 1: the candidate pairs of every job, in job order then candidate order (the tie rule depends on it)
  struct PairRef { size_t job; int cand; };
  std::vector<PairRef> ref;
  std::vector<int64_t> ids;
  std::vector<double> weights;
  std::vector<float> guesses;
  std::vector<const KeyframeRef*> src_kf, tgt_kf;
  auto sharded = [&](std::vector<b2r_result>& res, int with_fitness) -> b2r_status {
    const size_t n = ids.size();
    res.assign(n, b2r_result());
    if (n == 0) return B2R_OK;
    std::vector<int32_t> rank_of(n);
    b2r_status st = b2r_partition_by_target(ids.data(), weights.data(), n, nranks, rank_of.data());
    if (st != B2R_OK) return st;
    std::vector<b2r_cloud*> s(n, nullptr), t(n, nullptr);
    for (size_t i = 0; i < n; ++i)
      if (rank_of[i] == rank) { s[i] = src_kf[i]->cloud; t[i] = tgt_kf[i]->cloud; }
    return b2r_align_batch_sharded(h, comm, s.data(), t.data(), ids.data(), weights.data(), guesses.data(), n, with_fitness,
                                   prm.fitness_score_max_range, res.data());
  };
  auto push_pair = [&](size_t job, int cand, const KeyframeRef* target, const KeyframeRef* source) {
    ref.push_back(PairRef{job, cand});
    ids.push_back((int64_t)job);
    weights.push_back(source->cloud ? (double)b2r_cloud_size(source->cloud) : 1.0);
    const Mat4f g = registration_guess(target->estimate, source->estimate, prm.use_planar_registration_guess);
    guesses.insert(guesses.end(), g.begin(), g.end());
    src_kf.push_back(source);
    tgt_kf.push_back(target);
  };
  for (size_t j = 0; j < jobs.size(); ++j)
    for (size_t c = 0; c < jobs[j].candidates.size(); ++c) push_pair(j, (int)c, jobs[j].new_keyframe, jobs[j].candidates[c]);
  std::vector<b2r_result> res;
  b2r_status st = sharded(res, 1);
  for (size_t i = 0; i < ref.size(); ++i) out[ref[i].job].aligns++;
  if (st != B2R_OK) {
    for (LoopMatch& m : out) m.status = st;
    return out;
  }
  for (size_t i = 0; i < ref.size(); ++i) {  // :137-144 per job, candidates in order; DBL_MAX threshold: acceptance comes later
    LoopMatch& m = out[ref[i].job];
    if (!res[i].converged || res[i].fitness > m.best_score) continue;
    m.best_score = res[i].fitness;
    m.best = ref[i].cand;
    for (int t = 0; t < 16; ++t) m.rel_pose_new_to_best[t] = res[i].T[t];
  }
  // ---- stage 2: consistency checks (:190-303), prev first, next for those that did not pass
  std::vector<size_t> todo;
  for (size_t j = 0; j < jobs.size(); ++j) {
    LoopMatch& m = out[j];
    const KeyframeRef* best = m.best >= 0 ? jobs[j].candidates[m.best] : nullptr;
    if (best && (best->first_keyframe || best->static_keyframe)) { m.consistency_passed = true; continue; }  // :197-199
    if (!best || !prm.enable_loop_closure_consistency_check || m.best_score > prm.fitness_score_thresh) continue;  // :201-204
    todo.push_back(j);
  }
  for (int stage = 0; stage < 2; ++stage) {
    ref.clear(); ids.clear(); weights.clear(); guesses.clear(); src_kf.clear(); tgt_kf.clear();
    for (size_t j : todo) {
      const KeyframeRef* best = jobs[j].candidates[out[j].best];
      if (out[j].consistency_passed) continue;
      const KeyframeRef* other = stage == 0 ? best->prev : best->next;
      if (other) push_pair(j, stage, jobs[j].new_keyframe, other);
    }
    st = sharded(res, 0);
    for (size_t i = 0; i < ref.size(); ++i) {
      const size_t j = ref[i].job;
      LoopMatch& m = out[j];
      ++m.aligns;
      const KeyframeRef* best = jobs[j].candidates[m.best];
      Mat4f T;
      for (int t = 0; t < 16; ++t) T[t] = st == B2R_OK ? res[i].T[t] : guesses[i * 16 + t];
      const Mat4f M = stage == 0 ? detail::mul(detail::mul(detail::inverse(T), m.rel_pose_new_to_best), detail::to_float(best->rel_pose_to_prev))
                                 : detail::mul(detail::mul(detail::inverse(m.rel_pose_new_to_best), T), detail::to_float(best->rel_pose_from_next));
      detail::identity_delta(M, m.delta_trans[stage], m.delta_angle[stage]);
      m.consistency_passed = !(m.delta_trans[stage] > prm.loop_closure_consistency_max_delta_trans ||
                               m.delta_angle[stage] > prm.loop_closure_consistency_max_delta_angle);
    }
  }
  // ---- stage 3: acceptance (:156-166)
  for (size_t j = 0; j < jobs.size(); ++j) {
    LoopMatch& m = out[j];
    const KeyframeRef* best = m.best >= 0 ? jobs[j].candidates[m.best] : nullptr;
    if (m.best_score > prm.fitness_score_thresh) continue;
    if (prm.enable_loop_closure_consistency_check && best && !best->first_keyframe && !m.consistency_passed) continue;
    m.loop_found = best != nullptr;
  }
  return out;
}

}  // namespace b2r
