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
      B2R_CUDA(cudaMallocAsync(&a->p, a->bytes, stream));
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
  std::vector<std::shared_ptr<Arena>> mem;  // keeps every buffer below alive
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
  Ref<float4> spts;
  // covariances
  int cov_k = 0;  // k the covariances were built with (0 = none)
  int cov_mode = 0;  // Needs::cov_mode they were built with
  Ref<double> cov;
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
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t user_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

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

// ---- lsq.cu (FAST_GICP / FAST_VGICP) ----
struct PairDesc {
  int src;  // index into the views array
  int tgt;
};
void lsq_align_batch(Ctx& ctx, const b2r_config& cfg, const CloudView* d_views, const std::vector<PairDesc>& pairs, const int* src_sizes,
                     const float* guesses_colmajor, b2r_result* out);
void lsq_debug_linearize(Ctx& ctx, const b2r_config& cfg, const CloudView* d_views, int n_src, const double* T_lin, const double* T_trial,
                         bool trial, double* H, double* b, double* err, int32_t* corr_out, uint8_t* corr_valid);

// ---- ndt.cu ----
void gicp_pcl_align_batch(Ctx& ctx, const b2r_config& cfg, const CloudView* d_views, const std::vector<PairDesc>& pairs, const int* src_sizes,
                          const float* guesses_colmajor, b2r_result* out);
void ndt_align_batch(Ctx& ctx, const b2r_config& cfg, const CloudView* d_views, const std::vector<PairDesc>& pairs, const int* src_sizes,
                     const float* guesses_colmajor, b2r_result* out);
void ndt_debug_derivatives(Ctx& ctx, const b2r_config& cfg, const CloudView* d_views, int n_src, const double* p6, double* score,
                           double* grad6, double* hess36, int32_t* hits_out);

// ---- fitness (in lsq.cu) ----
void fitness_batch(Ctx& ctx, const CloudView* d_views, const std::vector<PairDesc>& pairs, const int* src_sizes, const float* T_colmajor,
                   double max_range, double* out);
void transform_cloud(Ctx& ctx, const float4* in, int n, const float* T_colmajor, float4* out);

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
