// libb2r host-side internals: cloud objects, handle, and the entry points of each kernel module.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"

namespace b2r {

// One stream-ordered device allocation shared by the structures carved out of it (cudaFreeAsync when the last user dies).
struct Arena {
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t s = nullptr;
  std::shared_ptr<StreamOwner> keep;  // keeps s alive until the free below has been enqueued
  Arena() = default;
  Arena(const Arena&) = delete;
  Arena& operator=(const Arena&) = delete;
  ~Arena() { if (p) cudaFreeAsync(p, s); }
};
// Collects the buffers of a build step, then carves all of them out of ONE allocation: a batch of clouds costs one
// cudaMallocAsync and one memset (the zero-initialised buffers come first) instead of a dozen per cloud.
struct ArenaPlan {
  struct Slot { void** dst; size_t off; bool zeroed; };
  std::vector<Slot> slots;
  size_t zbytes = 0, bytes = 0;
  static size_t up(size_t v) { return (v + 255) & ~(size_t)255; }
  template <typename T> void want(T*& dst, size_t count) { slots.push_back({(void**)&dst, bytes, false}); bytes = up(bytes + count * sizeof(T)); }
  template <typename T> void want_zeroed(T*& dst, size_t count) { slots.push_back({(void**)&dst, zbytes, true}); zbytes = up(zbytes + count * sizeof(T)); }
  bool empty() const { return slots.empty(); }
  std::shared_ptr<Arena> commit(const Ctx& ctx) {
    cudaStream_t stream = ctx.stream;
    auto a = std::make_shared<Arena>();
    a->s = stream;
    a->keep = ctx.stream_owner;
    a->bytes = zbytes + bytes;
    if (a->bytes) {
      B2R_CUDA(pool_malloc(&a->p, a->bytes, stream));
      if (zbytes) B2R_CUDA(cudaMemsetAsync(a->p, 0, zbytes, stream));
    }
    for (const Slot& sl : slots) *sl.dst = (char*)a->p + (sl.zeroed ? sl.off : zbytes + sl.off);
    return a;
  }
};

template <typename T>
struct Ref {  // non-owning device pointer into one of the cloud's arenas
  T* p = nullptr;
};

// Device-resident cloud with lazily built, cached search structures.
struct Cloud {
  int device = 0;
  int n = 0;
  // scratch of run_align: this cloud's slot in the current batch's list of distinct clouds (valid while batch_epoch matches)
  int batch_slot = 0;
  uint64_t batch_epoch = 0;
  // The allocation each group of buffers below was carved out of (shared with the other structures / clouds of the same build
  // step).  Rebuilding a structure with other parameters (k, covariance mode, resolution, leaf) replaces its entry, so the
  // superseded allocation is released as soon as no other structure uses it: device memory does not grow with alternating use.
  std::shared_ptr<Arena> mem_pts, mem_grid, mem_cov, mem_vox, mem_ndt;
  // declared after the arenas: released first, and its destructor waits for an upload that may still be writing into mem_pts
  std::shared_ptr<PendingBoxes> pending;
  int pending_idx = 0;
  Ref<float4> pts;
  bool has_bbox = false;
  float bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
  // NN grid
  bool has_grid = false;
  float h = 0.f, hx = 0.f;
  int gd[3] = {0, 0, 0};
  int ncell = 0;
  Ref<int> cell_start, cell_cnt;
  Ref<int2> cell_tmp;
  Ref<float4> spts, spair;
  // covariances
  int cov_k = 0;  // k the covariances were built with (0 = none)
  int cov_mode = 0;  // Needs::cov_mode they were built with
  Ref<double> cov, nrm;
  // VGICP voxel map
  double vres = 0.0;  // resolution the map was built with (0 = none)
  int vmin[3] = {0, 0, 0}, vd[3] = {0, 0, 0};
  int vcell = 0;
  Ref<int> v_start, v_cnt, v_order, v_table, v_nrec, v_reccell;
  Ref<VoxRec> vrec;
  // NDT grid
  float leaf = 0.f;  // leaf the grid was built with (0 = none)
  int min_b[3] = {0, 0, 0}, max_b[3] = {0, 0, 0}, div_b[3] = {0, 0, 0};
  int ncell_ndt = 0;
  bool ndt_overflow = false;
  bool ndt_centroids = false;  // the grid was built with the leaves' float centroids
  Ref<int> n_start, n_cnt, n_order, n_table, n_nrec, n_reccell;
  Ref<NdtRec> nrec;

  CloudView view() const;
};

// What a role in an alignment needs from a cloud.
struct Needs {
  bool grid = false;
  int cov_k = 0;
  int cov_mode = 0;  // 0: fast_gicp / small_gicp covariances, 1: pcl::GeneralizedIterativeClosestPoint::computeCovariances
  double vres = 0.0;
  float leaf = 0.f;
  bool leaf_centroids = false;  // NDT KDTREE search: the leaves' float centroids as well
};

struct Handle {
  b2r_config cfg;
  Ctx ctx;
  std::string last_error;
  std::unique_ptr<Cloud> owned_source, owned_target;
  Cloud* source = nullptr;
  Cloud* target = nullptr;
  // pcl::Registration state
  float final_T[16];
  bool converged = false;
  bool has_result = false;
  float timings[4] = {0, 0, 0, 0};
  bool timings_pending = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t user_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

}  // namespace b2r
struct b2r_handle { b2r::Handle h; };
struct b2r_cloud { b2r::Cloud c; };
namespace b2r {
inline Handle* b2r_handle_impl(b2r_handle* h) { return &h->h; }
inline Cloud* b2r_cloud_impl(b2r_cloud* c) { return &c->c; }

// ---- api.cu: the batch path behind b2r_align / b2r_align_batch / b2r_align_batch_sharded ----
//   out_all   host array of n rows, or nullptr (then the call returns without synchronising)
//   rows_dev  device array of n rows the results are (also) left in, or nullptr
void b2r_run_align(Handle& h, const std::vector<Cloud*>& sources, const std::vector<Cloud*>& targets, const float* guesses_colmajor,
                   int with_fitness, double fitness_max_range, b2r_result* out_all, b2r_result* rows_dev);

// ---- cloud.cu ----
// uploads `count` clouds into one shared allocation and computes their bounding boxes; one synchronisation at the end
void clouds_upload(Ctx& ctx, Cloud* const* clouds, const void* const* points, const size_t* n, size_t count, size_t stride_bytes, int memspace);
// Builds whatever of `needs[i]` is missing in clouds[i] (batched launches over all clouds, no synchronisation after the
// bounding boxes are known) and returns the device array of the clouds' views, in the order given: dviews.p[i].
void clouds_prepare(Ctx& ctx, const b2r_config& cfg, const std::vector<Cloud*>& clouds, const std::vector<Needs>& needs,
                    DBuf<CloudView>& dviews);
void debug_knn(Ctx& ctx, const b2r_config& cfg, Cloud& c, const float* queries, size_t nq, int k, int32_t* idx_out, float* d2_out);
unsigned long long debug_knn_list_overflows(Ctx& ctx);  // per device, since library load
void debug_cov_knn(Ctx& ctx, const b2r_config& cfg, Cloud& c, int k, int32_t* knn_out);

// ---- the batch every optimiser works on: all of it in device memory, results included (api.cu: run_align) ----
struct PairDesc {
  int src;  // index into the views array
  int tgt;
};
struct BatchArgs {
  const CloudView* d_views;
  const PairDesc* d_pairs;
  const int* d_src_n;       // source points per pair
  const float* d_guesses;   // 16 floats per pair, column-major
  const float* guesses;     // host copy of d_guesses (NDT / GICP_PCL initialise their state machines on the host)
  b2r_result* d_rows;       // one result row per pair; the optimisers fill everything but `fitness`
  int np;
  int maxn;                 // largest source cloud of the batch
  const int* src_sizes;     // host copy of d_src_n
};

// ---- loop.cu: { eval ; step } until every pair is done, as one CUDA-graph launch (WHILE node) or, when per-kernel profiling is
// on or B2R_GRAPH_LOOP=0, as a host-polled loop.  `la` must be the LoopArgs object the step kernel's argument array points at;
// prof_id < 0: the evaluation launches are not bracketed.  Stream-ordered: returns without synchronising in graph mode.
void run_device_loop(Ctx& ctx, const void* eval_fn, dim3 eval_grid, dim3 eval_block, void** eval_args, const void* step_fn, dim3 step_grid,
                     dim3 step_block, void** step_args, LoopArgs& la, LoopCtl* d_ctl, int npairs, long max_rounds, int prof_id,
                     const void* eval2_fn = nullptr);  // eval2_fn: a second evaluation kernel per round (same grid, block, arguments)

// ---- lsq.cu (FAST_GICP / FAST_VGICP / SMALL_GICP) ----
void lsq_align_batch(Ctx& ctx, const b2r_config& cfg, const BatchArgs& b);
void lsq_debug_linearize(Ctx& ctx, const b2r_config& cfg, const CloudView* d_views, int n_src, const double* T_lin, const double* T_trial,
                         bool trial, double* H, double* b, double* err, int32_t* corr_out, uint8_t* corr_valid);

// ---- gicp_pcl.cu / ndt.cu ----
void gicp_pcl_align_batch(Ctx& ctx, const b2r_config& cfg, const BatchArgs& b);
void ndt_align_batch(Ctx& ctx, const b2r_config& cfg, const BatchArgs& b);
void ndt_debug_derivatives(Ctx& ctx, const b2r_config& cfg, const CloudView* d_views, int n_src, const double* p6, double* score,
                           double* grad6, double* hess36, int32_t* hits_out);

// ---- fitness / inlier fraction (in lsq.cu): reads T from the rows, writes rows[i].fitness.  inlier_d2 > 0 additionally counts the
// source points whose nearest target point is closer than sqrt(inlier_d2) into inlier_out[i] (device, np ints) ----
void fitness_batch(Ctx& ctx, const BatchArgs& b, double max_range, float inlier_d2 = 0.f, int* d_inlier_out = nullptr);
void transform_cloud(Ctx& ctx, const float4* in, int n, const float* T_colmajor, float4* out);
// per source point: nearest target point (original index), squared distance, transformed point (device arrays of n / n / 3n)
void nearest_neighbors(Ctx& ctx, const CloudView* d_views, int n_src, const float* T_colmajor, int32_t* d_idx, float* d_d2, float* d_xyz);

// ---- filters.cu ----
struct DevCloud {  // packed device points + count
  DBuf<float4> pts;
  int n = 0;
  // a box known to contain every point (VoxelGrid leaves the box of its input: centroids lie inside it), so that the
  // outlier filter that follows need not run (and wait for) another bounding-box pass for its search grid
  bool has_box = false;
  float box_min[3] = {0, 0, 0}, box_max[3] = {0, 0, 0};
};
void filter_distance(Ctx& ctx, const float4* in, int n, double near_t, double far_t, DevCloud& out);
void filter_voxelgrid(Ctx& ctx, const float4* in, int n, float leaf, int min_pts, DevCloud& out, bool& overflow,
                      const double* range = nullptr);  // range: {near, far} of a preceding distance filter, folded in
void filter_radius(Ctx& ctx, const b2r_config& cfg, const float4* in, int n, double radius, int min_nb, DevCloud& out,
                   const DevCloud* box = nullptr);
void filter_statistical(Ctx& ctx, const b2r_config& cfg, const float4* in, int n, int mean_k, double stddev_mul, DevCloud& out,
                        const DevCloud* box = nullptr);

// MapCloudGenerator::generate + ApproximateMeanVoxelGrid; null_result mirrors the reference's nullptr returns
void map_cloud(Ctx& ctx, const void* const* clouds, const size_t* n, const double* poses_colmajor, const uint8_t* first_keyframe, size_t count,
               size_t stride_bytes, int memspace, float resolution, int min_points_per_voxel, float distance_far_thresh, int skip_first_cloud,
               DevCloud& out, bool& null_result);

// other robots' points (mrg_slam_component.cpp:395-427); others_xyz_host: n_others * 3 floats, sensor frame
void filter_robot_points(Ctx& ctx, const float4* in, int n, const float* others_xyz_host, int n_others, float radius_sqr, DevCloud& kept,
                         DevCloud* removed);

// shared utilities (cloud.cu)
void compact_points(Ctx& ctx, const float4* in, const uint8_t* keep, int n, DevCloud& out);
void load_points(Ctx& ctx, const void* points, size_t n, size_t stride_bytes, int memspace, DBuf<float4>& dst);
void store_points(Ctx& ctx, const float4* src, size_t n, void* out, size_t stride_bytes, int memspace);

}  // namespace b2r
