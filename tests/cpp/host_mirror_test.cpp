// Exercises the C++ host mirror (include/b2r/registration.hpp) and the PCL adapter (against tests/cpp/pcl_stub)
// the way the reference's callers drive a registration object:
//   apps/scan_matching_odometry_component.cpp:203-275 (setInputTarget / setInputSource / align / hasConverged /
//   getFinalTransformation, keyframe switch) and src/mrg_slam/loop_detector.cpp:126-145.
// and the loop matcher (include/b2r/loop_matcher.hpp) against src/mrg_slam/loop_detector.cpp:97-303.
// Usage: host_mirror_test [--expect-gpu]
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>

#include <b2r/information_matrix.hpp>
#include <b2r/loop_matcher.hpp>
#include <b2r/pcl_adapter.hpp>
#include <b2r/registration.hpp>

#include <cfloat>

extern "C" int b2r_synth_num_rays(int sensor);
extern "C" int b2r_synth_scan(int sensor, uint64_t seed, int scan_idx, float* out_xyzi);
extern "C" void b2r_synth_pose(uint64_t seed, int scan_idx, double* T16);

static int fails = 0;
#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) { std::printf("CHECK failed: %s (line %d)\n", #cond, __LINE__); ++fails; } \
  } while (0)

static b2r::PointCloud::Ptr make_cloud(int scan_idx) {
  const uint64_t seed = 0x5EED0000ull;
  std::vector<float> raw((size_t)b2r_synth_num_rays(0) * 4);
  int n = b2r_synth_scan(0, seed, scan_idx, raw.data());
  auto c = std::make_shared<b2r::PointCloud>();
  for (int i = 0; i < n; ++i) {
    const float* p = &raw[(size_t)i * 4];
    float d = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    if (d < 0.5f || d > 35.f || (i % 2)) continue;  // crude thinning, enough for a smoke-level check
    b2r::PointXYZI q; q.x = p[0]; q.y = p[1]; q.z = p[2]; q.intensity = p[3];
    c->points.push_back(q);
  }
  return c;
}


static b2r::Mat4d pose_colmajor(int scan_idx) {
  double T[16];
  b2r_synth_pose(0x5EED0000ull, scan_idx, T);  // row-major
  b2r::Mat4d M;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) M[c * 4 + r] = T[r * 4 + c];
  return M;
}
static b2r::Mat4d mul_d(const b2r::Mat4d& A, const b2r::Mat4d& B) {
  b2r::Mat4d C{};
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      for (int k = 0; k < 4; ++k) C[c * 4 + r] += A[k * 4 + r] * B[c * 4 + k];
  return C;
}
static b2r::Mat4d inv_rigid(const b2r::Mat4d& A) {
  b2r::Mat4d I{};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) I[c * 4 + r] = A[r * 4 + c];
  for (int r = 0; r < 3; ++r) I[12 + r] = -(I[r] * A[12] + I[4 + r] * A[13] + I[8 + r] * A[14]);
  I[15] = 1.0;
  return I;
}

// pure host math of the loop matcher: runs everywhere
static void loop_matcher_math() {
  b2r::Mat4d T = pose_colmajor(7);
  for (int i = 0; i < 12; ++i)
    if (i % 4 != 3) T[i] *= 1.0 + 1e-6;  // slightly denormalised rotation, what normalize_estimate is for
  const b2r::Mat4d N = b2r::normalize_estimate(T);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += N[a * 4 + k] * N[b * 4 + k];
      CHECK(std::fabs(s - (a == b ? 1.0 : 0.0)) < 1e-14);
    }
  CHECK(N[12] == T[12] && N[13] == T[13] && N[14] == T[14]);
  const b2r::Mat4f g = b2r::registration_guess(pose_colmajor(3), pose_colmajor(4));
  const b2r::Mat4d gd = mul_d(inv_rigid(pose_colmajor(3)), pose_colmajor(4));
  for (int i = 0; i < 16; ++i) CHECK(std::fabs(g[i] - (float)gd[i]) < 1e-6f);
  const b2r::Mat4f gi = b2r::detail::mul(b2r::detail::inverse(g), g);
  for (int i = 0; i < 16; ++i) CHECK(std::fabs(gi[i] - (i % 5 == 0 ? 1.f : 0.f)) < 1e-5f);
  b2r::Mat4f R = b2r::identity4();
  const float ang = 0.05f;
  R[0] = std::cos(ang); R[4] = -std::sin(ang); R[1] = std::sin(ang); R[5] = std::cos(ang); R[12] = 0.3f; R[13] = -0.4f;
  float dt, da;
  b2r::detail::identity_delta(R, dt, da);
  CHECK(std::fabs(dt - 0.5f) < 1e-6f && std::fabs(da - ang) < 1e-5f);
  b2r::KeyframeRef kf;
  b2r::LoopMatch none = b2r::match_keyframe(nullptr, kf, {});
  CHECK(!none.loop_found && none.best == -1 && none.aligns == 0);  // no candidates: nullptr (:99-101)
}

// ---- an in-process "world" of N ranks (one thread each) for the host transport of b2r_comm: every rank deposits its block,
// the last one in releases all of them.  What MPI_Allgather / a gloo all_gather would do between processes.
struct ThreadWorld {
  int n;
  std::mutex mu;
  std::condition_variable cv;
  std::vector<std::vector<char>> blocks;
  int arrived = 0, generation = 0;
  explicit ThreadWorld(int n_) : n(n_), blocks(n_) {}
};
struct ThreadRank { ThreadWorld* w; int rank; };
static int thread_allgather(void* user, const void* send, void* recv, size_t bytes) {
  ThreadRank* r = static_cast<ThreadRank*>(user);
  ThreadWorld& w = *r->w;
  std::unique_lock<std::mutex> lk(w.mu);
  w.blocks[r->rank].assign((const char*)send, (const char*)send + bytes);
  const int gen = w.generation;
  if (++w.arrived == w.n) {
    w.arrived = 0;
    ++w.generation;
    w.cv.notify_all();
  } else {
    w.cv.wait(lk, [&] { return w.generation != gen; });
  }
  for (int i = 0; i < w.n; ++i) std::memcpy((char*)recv + (size_t)i * bytes, w.blocks[i].data(), bytes);
  // nobody may overwrite a block before everyone has copied it: second rendezvous
  const int gen2 = w.generation;
  if (++w.arrived == w.n) {
    w.arrived = 0;
    ++w.generation;
    w.cv.notify_all();
  } else {
    w.cv.wait(lk, [&] { return w.generation != gen2; });
  }
  return 0;
}

// host-side logic of the sharded batch (no GPU needed): partition, gather through the C entry point at world size 2, best-candidate rule
static void sharding_host_logic() {
  const size_t n_targets = 9, k = 5, n = n_targets * k;
  std::vector<int64_t> ids(n);
  std::vector<double> w(n);
  for (size_t i = 0; i < n; ++i) { ids[i] = 100 + (int64_t)(i / k); w[i] = 1.0 + (double)(i % 3); }
  std::vector<int32_t> rank_of(n, -1);
  CHECK(b2r_partition_by_target(ids.data(), w.data(), n, 2, rank_of.data()) == B2R_OK);
  size_t c0 = 0;
  for (size_t i = 0; i < n; ++i) {
    CHECK(rank_of[i] == 0 || rank_of[i] == 1);
    CHECK(rank_of[i] == rank_of[(i / k) * k]);           // a target never straddles two ranks
    if (i) CHECK(rank_of[i] >= rank_of[i - 1]);          // contiguous blocks of targets
    c0 += rank_of[i] == 0;
  }
  CHECK(c0 >= 2 * k && c0 <= n - 2 * k);                 // both ranks get a real share
  auto fake = [&](size_t i) {
    b2r_result r;
    std::memset(&r, 0, sizeof(r));
    for (int t = 0; t < 16; ++t) r.T[t] = (float)(i * 16 + t);
    r.converged = (i % 7) != 3;
    r.iterations = (int)(i % 5);
    r.error = 0.5 * (double)i;
    r.evals = 2 + (int)(i % 3);
    r.fitness = 0.25 * (double)((i * 2654435761u >> 7) % 4);  // coarse values: ties occur
    return r;
  };
  ThreadWorld world(2);
  std::vector<b2r_result> table[2];
  std::thread th[2];
  for (int rk = 0; rk < 2; ++rk) {
    th[rk] = std::thread([&, rk] {
      ThreadRank me{&world, rk};
      b2r_comm* comm = nullptr;
      CHECK(b2r_comm_init_host(thread_allgather, &me, rk, 2, &comm) == B2R_OK);
      CHECK(b2r_comm_rank(comm) == rk && b2r_comm_size(comm) == 2);
      std::vector<b2r_result> local;
      for (size_t i = 0; i < n; ++i)
        if (rank_of[i] == rk) local.push_back(fake(i));
      table[rk].resize(n);
      CHECK(b2r_gather_results(nullptr, comm, rank_of.data(), n, local.data(), table[rk].data()) == B2R_OK);
      CHECK(b2r_comm_collectives(comm) == 1);
      b2r_comm_destroy(comm);
    });
  }
  for (auto& t : th) t.join();
  for (size_t i = 0; i < n; ++i) {
    const b2r_result want = fake(i);
    CHECK(std::memcmp(&table[0][i], &want, sizeof(want)) == 0 && std::memcmp(&table[1][i], &want, sizeof(want)) == 0);
  }
  std::vector<int64_t> best(n);
  std::vector<double> score(n);
  size_t nt = 0;
  CHECK(b2r_select_best_candidates(table[0].data(), ids.data(), n, 0.5, best.data(), score.data(), &nt) == B2R_OK && nt == n_targets);
  for (size_t t = 0; t < n_targets; ++t) {  // the reference's loop (:106-145) followed by the threshold (:156)
    double bs = DBL_MAX;
    int64_t bi = -1;
    for (size_t c = 0; c < k; ++c) {
      const b2r_result& r = table[0][t * k + c];
      if (!r.converged || r.fitness > bs) continue;
      bs = r.fitness;
      bi = (int64_t)(t * k + c);
    }
    CHECK(score[t] == bs && best[t] == (bs > 0.5 ? -1 : bi));
  }
  CHECK(b2r_comm_init_host(nullptr, nullptr, 0, 2, nullptr) == B2R_ERR_INVALID_ARG);
  std::printf("sharding host logic ok\n");
}

int main(int argc, char** argv) {
  const bool expect_gpu = argc > 1 && std::strcmp(argv[1], "--expect-gpu") == 0;
  loop_matcher_math();
  sharding_host_logic();
  // ---- factory dispatch (registrations.cpp:46-148)
  b2r::RegistrationParams prm;
  prm.registration_method = "FAST_VGICP";
  prm.reg_resolution = 1.0;
  auto vg = b2r::select_registration_method(prm);
  CHECK(vg && vg->config().method == B2R_FAST_VGICP && vg->config().correspondence_randomness == 20);
  prm.registration_method = "NDT_OMP"; prm.reg_nn_search_method = "DIRECT1"; prm.reg_resolution = 0.5;
  auto ndt = b2r::select_registration_method(prm);
  CHECK(ndt && ndt->config().method == B2R_NDT_OMP && ndt->config().neighbor_search == B2R_DIRECT1 && ndt->config().resolution == 0.5);
  prm.registration_method = "NDT_OMP"; prm.reg_nn_search_method = "KDTREE";
  auto ndtk = b2r::select_registration_method(prm);
  CHECK(ndtk && ndtk->config().neighbor_search == B2R_KDTREE);  // registrations.cpp:140-141
  prm.registration_method = "BOGUS"; prm.reg_nn_search_method = "DIRECT1";
  auto fb = b2r::select_registration_method(prm);
  // unknown -> warning + NDT; no "OMP" in the string -> pcl::NormalDistributionsTransform (:122-128) = kd-tree radius search
  CHECK(fb && fb->config().method == B2R_NDT_OMP && fb->config().neighbor_search == B2R_KDTREE);
  prm.registration_method = "FAST_VGICP_CUDA";
  auto vc = b2r::select_registration_method(prm);
  CHECK(vc && vc->config().method == B2R_FAST_VGICP);
  prm.registration_method = "SMALL_GICP";  // the YAML default (registrations.cpp:46-54)
  prm.reg_max_correspondence_distance = 1.5;
  auto sg = b2r::select_registration_method(prm);
  CHECK(sg && sg->config().method == B2R_SMALL_GICP && sg->config().max_correspondence_distance == 1.5);
  prm.registration_method = "GICP_OMP";  // pclomp's BFGS GICP (registrations.cpp:104-116)
  prm.reg_max_optimizer_iterations = 7;
  auto pg = b2r::select_registration_method(prm);
  CHECK(pg && pg->config().method == B2R_GICP_PCL && pg->config().max_optimizer_iterations == 7 && pg->config().gicp_epsilon == 1e-3);
  prm.registration_method = "ICP";  // outside the engine
  CHECK(b2r::select_registration_method(prm) == nullptr);

  auto a = make_cloud(3), b = make_cloud(4), c = make_cloud(5);
  double Ta[16], Tb[16];
  b2r_synth_pose(0x5EED0000ull, 3, Ta);
  b2r_synth_pose(0x5EED0000ull, 4, Tb);
  const double gt_dx = Tb[3] - Ta[3];  // dominant forward motion, world frame ~ sensor frame for this trajectory

  // ---- odometry-style use of the mirror
  b2r::PointCloud aligned;
  vg->setInputTarget(a);
  vg->setInputSource(b);
  vg->align(aligned, b2r::identity4());
  if (!expect_gpu && !vg->hasConverged()) {
    // CPU-only box: the engine must refuse loudly, never fall back
    CHECK(vg->lastError().find("no CUDA device") != std::string::npos || !vg->lastError().empty());
    b2r::Matrix4f T = vg->getFinalTransformation();
    CHECK(T == b2r::identity4());  // failure leaves the guess
    std::printf("host mirror: no-device path ok (%s)\n", vg->lastError().c_str());
  } else {
    CHECK(vg->hasConverged());
    b2r::Matrix4f T = vg->getFinalTransformation();
    CHECK(std::fabs(T[12] - gt_dx) < 0.15);
    CHECK(aligned.size() == b->size());
    CHECK(std::fabs(aligned.points[0].x - (T[0] * b->points[0].x + T[4] * b->points[0].y + T[8] * b->points[0].z + T[12])) < 1e-4);
    const double f = vg->getFitnessScore();
    CHECK(f > 0 && f < 1.0);
    // keyframe switch: the current source becomes the target (scan_matching_odometry_component.cpp:332-333)
    vg->setInputTarget(b);
    vg->setInputSource(c);
    vg->align(aligned, b2r::identity4());
    CHECK(vg->hasConverged());
    // ---- the PCL adapter through pcl::Registration's own (stubbed) align()
    auto pa = std::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
    auto pb = std::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
    pa->points.resize(a->size()); pb->points.resize(b->size());
    std::memcpy(static_cast<void*>(pa->points.data()), a->points.data(), a->size() * 32);
    std::memcpy(static_cast<void*>(pb->points.data()), b->points.data(), b->size() * 32);
    b2r::PclRegistration::Ptr reg(new b2r::PclRegistration(B2R_FAST_VGICP));
    reg->setResolution(1.0);
    reg->setTransformationEpsilon(0.1);
    reg->setMaximumIterations(64);
    reg->setCorrespondenceRandomness(20);
    pcl::Registration<pcl::PointXYZI, pcl::PointXYZI>::Ptr base = reg;  // what select_registration_method returns
    base->setInputTarget(pa);
    base->setInputSource(pb);
    pcl::PointCloud<pcl::PointXYZI> out;
    base->align(out);
    CHECK(base->hasConverged());
    auto Tp = base->getFinalTransformation();
    for (int i = 0; i < 16; ++i) CHECK(Tp.data()[i] == T[i]);  // same engine, same bits
    CHECK(std::fabs(reg->fitness() - f) < 1e-12);
    // pcl::Registration's own (non-virtual) getFitnessScore walks tree_: every query must be served from the GPU table, the
    // result must be the GPU fitness to the last bit of the host's summation order, and no host kd-tree may have been built
    {
      const double f_pcl = base->getFitnessScore();
      CHECK(std::fabs(f_pcl - f) <= 1e-12 * f);
      CHECK(reg->gpuTree().served == (long)pb->size() && reg->gpuTree().host_builds == 0 && reg->gpuTree().host_queries == 0);
      const double f_rng = base->getFitnessScore(0.01);  // max_range (squared) below most distances: a second pass over the table
      CHECK(f_rng < f_pcl && reg->gpuTree().served == 2 * (long)pb->size());
      // the inlier loop of scan_matching_odometry_component.cpp:409-415 through getSearchMethodTarget()
      pcl::Indices ki; std::vector<float> kd;
      int inl = 0;
      for (size_t i = 0; i < out.size(); ++i) {
        base->getSearchMethodTarget()->nearestKSearch(out.points[i], 1, ki, kd);
        if (kd[0] < 0.5 * 0.5) ++inl;
      }
      double frac = 0, fit2 = 0;
      CHECK(b2r_inlier_fraction(reg->engine().handle(), 0.5, &frac, &fit2) == B2R_OK);
      CHECK(frac == (double)((float)inl / (float)out.size()) && reg->gpuTree().host_queries == 0);
      // a query the table cannot serve falls back to the host tree (built now, once)
      pcl::PointXYZI stray; stray.x = 1.f; stray.y = 2.f; stray.z = 3.f;
      base->getSearchMethodTarget()->nearestKSearch(stray, 1, ki, kd);
      CHECK(reg->gpuTree().host_builds == 1 && reg->gpuTree().host_queries == 1 && ki[0] >= 0);
    }
    // ---- loop matcher (loop_detector.cpp:97-303): new keyframe = scan 4 seen again, candidates = scans 3 and 5
    {
      b2r_handle* h = vg->handle();
      auto up = [&](const b2r::PointCloud::Ptr& pc) {
        b2r_cloud* cl = nullptr;
        CHECK(b2r_cloud_create(h, pc->points.data(), pc->size(), 32, B2R_HOST, &cl) == B2R_OK);
        return cl;
      };
      auto d2 = make_cloud(2), d6 = make_cloud(6);
      b2r::KeyframeRef k2, k3, k4, k5, k6;
      b2r::KeyframeRef* all[5] = {&k2, &k3, &k4, &k5, &k6};
      b2r::PointCloud::Ptr pcs[5] = {d2, a, b, c, d6};
      for (int i = 0; i < 5; ++i) { all[i]->cloud = up(pcs[i]); all[i]->estimate = pose_colmajor(2 + i); }
      auto rel = [&](int from, int to) { return mul_d(inv_rigid(pose_colmajor(from)), pose_colmajor(to)); };  // from <- to
      k3.prev = &k2; k3.rel_pose_to_prev = rel(3, 2); k3.next = &k5; k3.rel_pose_from_next = rel(5, 3);
      k5.prev = &k3; k5.rel_pose_to_prev = rel(5, 3); k5.next = &k6; k5.rel_pose_from_next = rel(6, 5);
      const std::vector<const b2r::KeyframeRef*> cands = {&k3, &k5};
      b2r::LoopMatch m = b2r::match_keyframe(h, k4, cands);
      CHECK(m.status == B2R_OK && m.best >= 0 && m.best_score < 1.25);
      CHECK(m.loop_found && m.consistency_passed && m.aligns == 3);  // two candidates + the prev check
      CHECK(m.delta_trans[0] >= 0.f && m.delta_trans[0] < 0.3f && m.delta_trans[1] < 0.f);
      // ---- InformationMatrixCalculator::calc_information_matrix (information_matrix_calculator.cpp:14-44) of the loop edge
      {
        double inf[36], fit = -1.0;
        b2r::Mat4d relpose;
        for (int i = 0; i < 16; ++i) relpose[i] = (double)m.rel_pose_new_to_best[i];
        b2r::InformationMatrixParams ip;
        CHECK(b2r::calc_information_matrix(h, k4.cloud, cands[m.best]->cloud, relpose.data(), ip, inf, &fit) == B2R_OK);
        double want_fit = 0.0;
        CHECK(b2r_fitness_pair(h, k4.cloud, cands[m.best]->cloud, m.rel_pose_new_to_best.data(), DBL_MAX, &want_fit) == B2R_OK);
        CHECK(fit == want_fit && fit > 0.0 && fit < 1.25);
        const double y = (1.0 - std::exp(-2.0 * fit)) / (1.0 - std::exp(-2.0 * 1.25));
        const double wx = 0.1 * 0.1 + (0.75 * 0.75 - 0.1 * 0.1) * y, wq = 0.05 * 0.05 + (0.2 * 0.2 - 0.05 * 0.05) * y;
        for (int r = 0; r < 6; ++r)
          for (int c2 = 0; c2 < 6; ++c2)
            CHECK(std::fabs(inf[r * 6 + c2] - (r != c2 ? 0.0 : (r < 3 ? 1.0 / wx : 1.0 / wq))) <= 1e-12 / (r < 3 ? wx : wq));
        ip.use_const_inf_matrix = true;
        CHECK(b2r::calc_information_matrix(h, k4.cloud, cands[m.best]->cloud, relpose.data(), ip, inf) == B2R_OK);
        CHECK(inf[0] == 1.0 / 0.5 && inf[21] == 1.0 / 0.1 && inf[1] == 0.0);
      }
      // the best candidate's pose is the single-pair result, bit for bit
      b2r_result r1;
      const b2r::Mat4f g = b2r::registration_guess(k4.estimate, cands[m.best]->estimate);
      CHECK(b2r_set_target_cloud(h, k4.cloud) == B2R_OK && b2r_set_source_cloud(h, cands[m.best]->cloud) == B2R_OK);
      CHECK(b2r_align(h, g.data(), &r1) == B2R_OK);
      for (int i = 0; i < 16; ++i) CHECK(r1.T[i] == m.rel_pose_new_to_best[i]);
      // a prev edge that cannot close: the next keyframe is tried and accepts
      b2r::KeyframeRef* best = all[m.best == 0 ? 1 : 3];
      const b2r::Mat4d good_prev = best->rel_pose_to_prev, good_next = best->rel_pose_from_next;
      best->rel_pose_to_prev[12] += 1.0;
      m = b2r::match_keyframe(h, k4, cands);
      CHECK(m.loop_found && m.aligns == 4 && m.delta_trans[0] > 0.3f && m.delta_trans[1] >= 0.f && m.delta_trans[1] < 0.3f);
      // both edges wrong: rejected although the score is below the threshold (:162-166)
      best->rel_pose_from_next[12] += 1.0;
      m = b2r::match_keyframe(h, k4, cands);
      CHECK(!m.loop_found && !m.consistency_passed && m.best >= 0 && m.best_score < 1.25);
      // ... unless the candidate is a robot's first keyframe (:197-199) or the check is disabled
      best->first_keyframe = true;
      m = b2r::match_keyframe(h, k4, cands);
      CHECK(m.loop_found && m.aligns == 2);
      best->first_keyframe = false;
      b2r::LoopMatchParams off;
      off.enable_loop_closure_consistency_check = false;
      m = b2r::match_keyframe(h, k4, cands, off);
      CHECK(m.loop_found && m.aligns == 2);
      // a threshold nobody meets: loop not found (:156-160)
      b2r::LoopMatchParams strict;
      strict.fitness_score_thresh = 1e-9;
      m = b2r::match_keyframe(h, k4, cands, strict);
      CHECK(!m.loop_found && m.best >= 0);
      best->rel_pose_to_prev = good_prev; best->rel_pose_from_next = good_next;
      // ---- the sharded multi-keyframe matcher (b2r_align_batch_sharded) must take the same decisions
      {
        const b2r::LoopMatch single4 = b2r::match_keyframe(h, k4, cands);
        const std::vector<const b2r::KeyframeRef*> cands3 = {&k2, &k4};
        const b2r::LoopMatch single3 = b2r::match_keyframe(h, k3, cands3);
        std::vector<b2r::KeyframeJob> jobs(2);
        jobs[0].new_keyframe = &k4; jobs[0].candidates = cands;
        jobs[1].new_keyframe = &k3; jobs[1].candidates = cands3;
        auto same = [&](const b2r::LoopMatch& a, const b2r::LoopMatch& b) {
          bool ok = a.best == b.best && a.best_score == b.best_score && a.loop_found == b.loop_found &&
                    a.consistency_passed == b.consistency_passed && a.aligns == b.aligns;
          for (int i = 0; i < 16; ++i) ok = ok && a.rel_pose_new_to_best[i] == b.rel_pose_new_to_best[i];
          for (int i = 0; i < 2; ++i) ok = ok && a.delta_trans[i] == b.delta_trans[i] && a.delta_angle[i] == b.delta_angle[i];
          return ok;
        };
        // (a) NCCL transport, one rank: ncclCommInitRank + ncclAllGather on this GPU
        unsigned char id[B2R_UNIQUE_ID_BYTES];
        b2r_comm* nc = nullptr;
        CHECK(b2r_comm_unique_id(id) == B2R_OK);
        CHECK(b2r_comm_init(h, id, 0, 1, &nc) == B2R_OK);
        if (nc) {
          const std::vector<b2r::LoopMatch> got = b2r::match_keyframes_sharded(h, nc, jobs);
          CHECK(got.size() == 2 && got[0].status == B2R_OK && same(got[0], single4) && same(got[1], single3));
          CHECK(b2r_comm_collectives(nc) >= 1);
          b2r_comm_destroy(nc);
        }
        // (b) host transport, two ranks = two threads with one handle (stream) each on this GPU; every rank must end up with
        // the decisions of the single-rank run although each aligned only its own keyframe's candidates
        ThreadWorld world(2);
        std::vector<b2r::LoopMatch> res2[2];
        std::thread th[2];
        for (int rk = 0; rk < 2; ++rk) {
          th[rk] = std::thread([&, rk] {
            b2r_config cfg;
            b2r_default_config(B2R_FAST_VGICP, &cfg);
            b2r_handle* hr = nullptr;
            CHECK(b2r_create(&cfg, &hr) == B2R_OK);
            ThreadRank me{&world, rk};
            b2r_comm* comm = nullptr;
            CHECK(b2r_comm_init_host(thread_allgather, &me, rk, 2, &comm) == B2R_OK);
            res2[rk] = b2r::match_keyframes_sharded(hr, comm, jobs);
            b2r_comm_destroy(comm);
            b2r_destroy(hr);
          });
        }
        for (auto& t : th) t.join();
        for (int rk = 0; rk < 2; ++rk) CHECK(res2[rk].size() == 2 && same(res2[rk][0], single4) && same(res2[rk][1], single3));
        std::printf("sharded matcher: nccl(1 rank) and host transport (2 ranks) ok\n");
      }
      for (auto* k : all) b2r_cloud_destroy(k->cloud);
      std::printf("loop matcher: gpu path ok\n");
    }
    std::printf("host mirror: gpu path ok, tx=%.4f (gt %.4f), fitness=%.5f\n", T[12], gt_dx, f);
  }
  std::printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
  return fails ? 1 : 0;
}
