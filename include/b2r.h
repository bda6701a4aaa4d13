/* libb2r — B200-native scan registration + prefiltering for mrg_slam's hot path.
 *
 * C ABI (plain pointers and sizes; no C++/torch types).  Every entry point names
 * the reference interface it replaces (paths relative to /root/reference).
 * All computation runs in hand-written sm_100a CUDA kernels; there is no CPU
 * fallback: if no CUDA device is present b2r_create fails with B2R_ERR_NO_DEVICE.
 *
 * Conventions
 *  - points: float32 x,y,z,intensity.  stride_bytes = 16 (packed, the KITTI layout of
 *    python_scripts/kitti_singlerobot_processor.py:174-183; intensity at byte 12) or
 *    32 (pcl::PointXYZI: x,y,z,1 | intensity,pad; intensity at byte 16).
 *  - 4x4 transforms: 16 floats, COLUMN-major (Eigen::Matrix4f storage), as
 *    pcl::Registration::align / getFinalTransformation use.
 *  - memspace: where the caller's buffer lives (host or this handle's CUDA device).
 *  - a handle is thread-compatible (one thread at a time); distinct handles are
 *    independent (own CUDA stream, own workspace).
 */
#ifndef B2R_H_
#define B2R_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum b2r_status {
  B2R_OK = 0,
  B2R_ERR_INVALID_ARG = 1,
  B2R_ERR_CUDA = 2,
  B2R_ERR_NO_DEVICE = 3,
  B2R_ERR_CAPACITY = 4, /* a dense voxel table would exceed its cap (cloud extent / resolution) */
  B2R_ERR_STATE = 5,    /* e.g. align before set_target */
  B2R_ERR_COMM = 6      /* NCCL missing / a collective failed */
} b2r_status;

/* registration_method strings handled by select_registration_method (src/mrg_slam/registrations.cpp:46-148) */
typedef enum b2r_method {
  B2R_NDT_OMP = 0,   /* "NDT_OMP"    registrations.cpp:130-147 -> pclomp::NormalDistributionsTransform */
  B2R_FAST_GICP = 1, /* "FAST_GICP"  registrations.cpp:55-63   -> fast_gicp::FastGICP ("GICP" row) */
  B2R_FAST_VGICP = 2, /* "FAST_VGICP" registrations.cpp:76-84   -> fast_gicp::FastVGICP */
  B2R_SMALL_GICP = 3, /* "SMALL_GICP" registrations.cpp:46-54   -> small_gicp::RegistrationPCL (GICP), the YAML default */
  B2R_GICP_PCL = 4    /* "GICP" / "GICP_OMP" registrations.cpp:93-116 -> pcl / pclomp GeneralizedIterativeClosestPoint (BFGS) */
} b2r_method;

/* reg_nn_search_method (registrations.cpp:140-146).  KDTREE (NDT only): pclomp's radius search over the leaf centroids with
 * radius = resolution, which is also what pcl::NormalDistributionsTransform (method string "NDT") always does. */
typedef enum b2r_neighbor_search { B2R_DIRECT1 = 0, B2R_DIRECT7 = 1, B2R_DIRECT27 = 2, B2R_KDTREE = 3 } b2r_neighbor_search;
typedef enum b2r_memspace { B2R_HOST = 0, B2R_DEVICE = 1 } b2r_memspace;

/* The 10 reg_* ROS parameters read at registrations.cpp:34-43 plus the upstream
 * defaults mrg_slam never overrides (parameter table). */
typedef struct b2r_config {
  int method;                         /* b2r_method */
  int device;                         /* CUDA device ordinal */
  double transformation_epsilon;      /* reg_transformation_epsilon (0.1) */
  int maximum_iterations;             /* reg_maximum_iterations (64) */
  double max_correspondence_distance; /* reg_max_correspondence_distance (2.0), FAST_GICP only */
  int correspondence_randomness;      /* reg_correspondence_randomness (20): k of the covariance kNN, <= 32 */
  double resolution;                  /* reg_resolution (1.0): VGICP voxel / NDT leaf */
  int neighbor_search;                /* reg_nn_search_method for NDT_OMP (DIRECT7); FAST_VGICP keeps DIRECT1 */
  double rotation_epsilon;            /* fast_gicp default 2e-3 */
  int lm_max_iterations;              /* fast_gicp default 10 */
  double lm_init_lambda_factor;       /* fast_gicp default 1e-9 */
  double ndt_step_size;               /* ndt_omp default 0.1 */
  double ndt_outlier_ratio;           /* ndt_omp default 0.55 */
  double nn_cell_size;                /* uniform-grid cell for exact kNN / 1-NN; 0 = auto from density */
  int max_optimizer_iterations;       /* reg_max_optimizer_iterations (20): BFGS steps per outer iteration, GICP_PCL only */
  double gicp_epsilon;                /* PCL default 1e-3: the regularised smallest singular value, GICP_PCL only */
} b2r_config;

typedef struct b2r_result {
  float T[16];    /* getFinalTransformation(), column-major */
  int converged;  /* hasConverged() */
  int iterations; /* nr_iterations_ as the upstream class leaves it */
  double error;   /* last LM error (GICP/VGICP) or NDT score */
  int evals;      /* cost-function passes over the source cloud */
  int reserved;   /* 0 (explicit padding: result rows are all-gathered and compared as bytes) */
  double fitness; /* getFitnessScore(max_range) when requested by a batch call, else 0 */
} b2r_result;

typedef struct b2r_handle b2r_handle; /* one pcl::Registration object */
typedef struct b2r_cloud b2r_cloud;   /* a device-resident point cloud + its cached search structures */

/* ---- lifecycle: replaces the constructor + setter block of select_registration_method ---- */
b2r_status b2r_default_config(int method, b2r_config* cfg);
b2r_status b2r_create(const b2r_config* cfg, b2r_handle** out);
void b2r_destroy(b2r_handle* h);
const char* b2r_last_error(const b2r_handle* h);
const char* b2r_version(void);

/* ---- clouds.  A cloud may be shared by several handles' calls on the same device.  Search structures
 * (kNN grid, covariances, voxel maps) are built lazily and cached, so promoting a source to target at a
 * keyframe switch (scan_matching_odometry_component.cpp:332-333) reuses them. ---- */
b2r_status b2r_cloud_create(b2r_handle* h, const void* points, size_t n, size_t stride_bytes, int memspace, b2r_cloud** out);
/* same for `count` clouds with a single device synchronisation at the end (loop-closure batches, graph loading) */
b2r_status b2r_cloud_create_batch(b2r_handle* h, const void* const* points, const size_t* n, size_t count, size_t stride_bytes, int memspace,
                                  b2r_cloud** out);
void b2r_cloud_destroy(b2r_cloud* c);
void b2r_cloud_destroy_batch(b2r_cloud* const* clouds, size_t count); /* NULL entries are skipped */
size_t b2r_cloud_size(const b2r_cloud* c);

/* ---- pcl::Registration surface (apps/scan_matching_odometry_component.cpp:203,208,266,270,275;
 *      src/mrg_slam/loop_detector.cpp:104,127,134,137,138,144) ---- */
b2r_status b2r_set_target(b2r_handle* h, const void* points, size_t n, size_t stride_bytes, int memspace); /* setInputTarget */
b2r_status b2r_set_source(b2r_handle* h, const void* points, size_t n, size_t stride_bytes, int memspace); /* setInputSource */
b2r_status b2r_set_target_cloud(b2r_handle* h, b2r_cloud* c); /* setInputTarget with a retained cloud (not owned) */
b2r_status b2r_set_source_cloud(b2r_handle* h, b2r_cloud* c);
b2r_status b2r_align(b2r_handle* h, const float guess_colmajor[16], b2r_result* out); /* align(output, guess) */
/* getFitnessScore(max_range): max_range is compared against the SQUARED distance, as PCL does
 * (same code in-tree: src/mrg_slam/information_matrix_calculator.cpp:46-81). */
b2r_status b2r_fitness(b2r_handle* h, double max_range, double* out);
/* the `output` cloud of align(): source transformed by the final transformation (float, PCL association) */
b2r_status b2r_transform_source(b2r_handle* h, void* out_points, size_t stride_bytes, int memspace);
/* InformationMatrixCalculator::calc_fitness_score(cloud1, cloud2, relpose, max_range)
 * (src/mrg_slam/information_matrix_calculator.cpp:46-81): fitness of T*source against target. */
b2r_status b2r_fitness_pair(b2r_handle* h, b2r_cloud* target, b2r_cloud* source, const float T_colmajor[16], double max_range, double* out);

/* ---- loop-closure batch: the candidate loop of LoopDetector::matching (loop_detector.cpp:126-145)
 * for many (target, source, guess) triples at once.  If with_fitness != 0, out[i].fitness =
 * getFitnessScore(fitness_max_range) of pair i (loop_detector.cpp:137). ---- */
b2r_status b2r_align_batch(b2r_handle* h, b2r_cloud* const* sources, b2r_cloud* const* targets, const float* guesses_colmajor,
                           size_t n_pairs, int with_fitness, double fitness_max_range, b2r_result* out);

/* ScanMatchingOdometryComponent::publish_scan_matching_status (apps/scan_matching_odometry_component.cpp:403-415): the fraction of
 * the aligned source points (T = the last final transformation) whose nearest target point is closer than
 * max_correspondence_dist (0.5 there; k_sq_dists[0] < dist * dist), replacing the N host calls of
 * getSearchMethodTarget()->nearestKSearch(pt, 1, ...).  fitness_out (optional) receives getFitnessScore() of the same pass (:403). */
b2r_status b2r_inlier_fraction(b2r_handle* h, double max_correspondence_dist, double* fraction_out, double* fitness_out);

/* Per-point nearest neighbours of the aligned source (T = the last final transformation, float transform with PCL's association)
 * in the target: idx_out[i] = index of the nearest target point (-1: empty target), d2_out[i] = its squared distance (FLANN float
 * association), xyz_out (optional, 3 floats per point) = the transformed point itself.  This is the table behind
 * pcl::Registration::getFitnessScore's loop of tree_->nearestKSearch(point, 1, ...) calls (and the inlier loop of
 * scan_matching_odometry_component.cpp:409-415): include/b2r/pcl_adapter.hpp serves those calls from it through a
 * pcl::search::KdTree subclass, so unmodified callers of getFitnessScore() get GPU answers (option (i)). */
b2r_status b2r_nearest_neighbors(b2r_handle* h, int32_t* idx_out, float* d2_out, float* xyz_out);

/* ---- multi-GPU loop-closure batches: the candidate loops of LoopDetector::matching (loop_detector.cpp:97-180) of many
 * new keyframes, sharded over the GPUs of one box, one process (or thread) per GPU.  The pairs are partitioned by target id (all
 * candidates of a new keyframe on one rank: its target structures are built once), every rank aligns its slice, the fixed-size
 * result rows are written on the device into the send buffer of ONE ncclAllGather (NVLink) and every rank receives the whole
 * table in pair order.  There is no collective inside the optimiser.
 *   - b2r_comm_unique_id: rank 0 creates the NCCL id; the launcher distributes the 128 bytes (MPI_Bcast, torch.distributed, a file).
 *   - b2r_comm_init: ncclCommInitRank on the handle's device; collectives run on the handle's stream.  NCCL is dlopen'ed
 *     (libnccl.so.2; a copy the host process already loaded is reused), B2R_ERR_COMM if absent.
 *   - b2r_comm_init_host: the same entry points over a caller-provided HOST all-gather (MPI, gloo, tests): fn(user, send, recv,
 *     bytes_per_rank) must fill recv[r * bytes_per_rank ...] with rank r's send block on every rank and return 0. ---- */
#define B2R_UNIQUE_ID_BYTES 128
typedef struct b2r_comm b2r_comm;
typedef int (*b2r_allgather_fn)(void* user, const void* send, void* recv, size_t bytes_per_rank);
b2r_status b2r_comm_unique_id(void* id_out /* B2R_UNIQUE_ID_BYTES */);
b2r_status b2r_comm_init(b2r_handle* h, const void* unique_id, int rank, int nranks, b2r_comm** out);
b2r_status b2r_comm_init_host(b2r_allgather_fn fn, void* user, int rank, int nranks, b2r_comm** out);
void b2r_comm_destroy(b2r_comm* c);
int b2r_comm_rank(const b2r_comm* c);
int b2r_comm_size(const b2r_comm* c);
const char* b2r_comm_last_error(const b2r_comm* c);
uint64_t b2r_comm_collectives(const b2r_comm* c); /* all-gathers issued so far */
/* Static block partition by target id, optionally balanced by per-pair weights (e.g. source points): targets keep their order of
 * first appearance, each rank gets a contiguous block of targets.  Pure host code, identical on every rank. */
b2r_status b2r_partition_by_target(const int64_t* target_ids, const double* weights /* or NULL */, size_t n_pairs, int nranks,
                                   int32_t* rank_of_pair);
/* b2r_align_batch over the ranks of `comm`.  All ranks pass the same pair list (target_ids, weights, guesses); sources[i] /
 * targets[i] must be valid on the rank that b2r_partition_by_target assigns pair i to and may be NULL elsewhere.  out receives all
 * n_pairs rows on every rank.  A failure on one rank does not keep it out of the collective: its rows come back as
 * converged = 0, T = guess, fitness = DBL_MAX and that rank returns the error status. */
b2r_status b2r_align_batch_sharded(b2r_handle* h, b2r_comm* comm, b2r_cloud* const* sources, b2r_cloud* const* targets,
                                   const int64_t* target_ids, const double* weights /* or NULL */, const float* guesses_colmajor,
                                   size_t n_pairs, int with_fitness, double fitness_max_range, b2r_result* out);
/* the gather step alone: local = this rank's rows in pair order; out = all n_pairs rows.  h may be NULL with a host transport. */
b2r_status b2r_gather_results(b2r_handle* h, b2r_comm* comm, const int32_t* rank_of_pair, size_t n_pairs, const b2r_result* local,
                              b2r_result* out);
/* The candidate reduction of loop_detector.cpp:106-160 over a gathered table: per target (order of first appearance) the pair index
 * of the best converged candidate — `score > best_score -> skip`, so equal scores go to the LATER candidate — or -1 when there is
 * none or its score exceeds fitness_score_thresh.  best_pair_out / best_score_out: room for n_pairs entries. */
b2r_status b2r_select_best_candidates(const b2r_result* results, const int64_t* target_ids, size_t n_pairs, double fitness_score_thresh,
                                      int64_t* best_pair_out, double* best_score_out /* or NULL */, size_t* n_targets_out);

/* ---- prefiltering (apps/prefiltering_component.cpp:149-151).  `out` has room for n points in `memspace`,
 * packed 16 B; *m receives the number written. ---- */
/* distance_filter(), prefiltering_component.cpp:206-229 */
b2r_status b2r_distance_filter(b2r_handle* h, const void* in, size_t n, size_t stride_bytes, int memspace, double near_thresh,
                               double far_thresh, void* out, size_t* m);
/* pcl::VoxelGrid::filter, prefiltering_component.cpp:167-171 / scan_matching_odometry_component.cpp:175-179.
 * *overflow = 1 reproduces PCL's "leaf size too small" branch: output = input unchanged. */
b2r_status b2r_voxelgrid(b2r_handle* h, const void* in, size_t n, size_t stride_bytes, int memspace, float leaf, int min_points_per_voxel,
                         void* out, size_t* m, int* overflow);
/* pcl::RadiusOutlierRemoval::filter, prefiltering_component.cpp:195-199 */
b2r_status b2r_radius_outlier(b2r_handle* h, const void* in, size_t n, size_t stride_bytes, int memspace, double radius, int min_neighbors,
                              void* out, size_t* m);
/* pcl::StatisticalOutlierRemoval::filter, prefiltering_component.cpp:190-194 */
b2r_status b2r_statistical_outlier(b2r_handle* h, const void* in, size_t n, size_t stride_bytes, int memspace, int mean_k,
                                   double stddev_mul, void* out, size_t* m);

typedef struct b2r_prefilter_config {
  int enable_distance_filter; double distance_near_thresh, distance_far_thresh; /* :208-216 */
  int downsample_method;  /* 0 NONE, 1 VOXELGRID */                               /* :160-171 */
  float downsample_resolution; int downsample_min_points_per_voxel;
  int outlier_removal_method; /* 0 NONE, 1 STATISTICAL, 2 RADIUS */             /* :184-199 */
  int statistical_mean_k; double statistical_stddev;
  double radius_radius; int radius_min_neighbors;
} b2r_prefilter_config;
b2r_status b2r_default_prefilter_config(b2r_prefilter_config* cfg); /* values of config/mrg_slam.yaml:48-64 */
/* cloud_callback's filter chain distance_filter -> downsample -> outlier_removal (prefiltering_component.cpp:149-151)
 * with intermediates kept on the device. */
b2r_status b2r_prefilter(b2r_handle* h, const b2r_prefilter_config* cfg, const void* in, size_t n, size_t stride_bytes, int memspace,
                         void* out, size_t* m);

/* ---- other robots' points (apps/mrg_slam_component.cpp:395-427): before a keyframe is created, every point within
 * robot_remove_points_radius of another robot is dropped.  others_xyz: n_others * 3 floats on the HOST, the other robots' positions
 * already in the sensor frame ((odom^-1 * map2odom * p).cast<float>(), :397-403); radius_sqr = radius^2 as a float (:405-406).
 * kept_out / removed_out (optional) have room for n packed points in `memspace`; both keep the input order. ---- */
b2r_status b2r_remove_robot_points(b2r_handle* h, const void* in, size_t n, size_t stride_bytes, int memspace, const float* others_xyz,
                                   size_t n_others, float radius_sqr, void* kept_out, size_t* n_kept, void* removed_out, size_t* n_removed);

/* ---- map cloud.  MapCloudGenerator::generate (src/mrg_slam/map_cloud_generator.cpp:14-86): every keyframe cloud is
 * transformed by keyframe->pose (Isometry3d matrix, 16 doubles column-major each, cast to float as :36 does), points farther than
 * distance_far_thresh from their sensor are skipped when distance_far_thresh > 0 (:38-42), keyframes with first_keyframe[k] != 0 are
 * skipped when skip_first_cloud != 0 (:33-35), the rest is concatenated in keyframe order and, when resolution > 0, reduced by
 * pcl::ApproximateMeanVoxelGrid (include/pcl/filters/ApproximateMeanVoxelGrid.hpp:63-126: voxel = floor(p / leaf), float sums in
 * input order, mean = sum / count, kept iff count >= min_points_per_voxel).  The hash-map iteration order of the output is
 * implementation-defined upstream; here voxels come out in ascending (z, y, x) voxel order.  `out` has room for sum(n) packed
 * 16 B points in `memspace`; *is_null = 1 where generate() returns nullptr (no keyframes, or nothing left of several). ---- */
b2r_status b2r_map_cloud(b2r_handle* h, const void* const* clouds, const size_t* n, const double* poses_colmajor, const uint8_t* first_keyframe,
                         size_t count, size_t stride_bytes, int memspace, float resolution, int min_points_per_voxel,
                         float distance_far_thresh, int skip_first_cloud, void* out, size_t* m, int* is_null);

/* ---- introspection for parity tests and benchmarks ---- */
uint64_t b2r_kernel_launches(const b2r_handle* h); /* kernel + graph launches issued by this handle so far */
uint64_t b2r_graph_launches(const b2r_handle* h);  /* of which optimiser loops launched as one CUDA graph (WHILE node; the {eval, step}
                                                      rounds inside run without the host) */
b2r_status b2r_synchronize(b2r_handle* h);
/* kNN-covariance queries (on this device, since library load) whose in-kernel candidate log overflowed and took the
 * second-traversal path: a tuning counter for the exact-kNN grid, results are identical either way */
b2r_status b2r_debug_knn_list_overflows(b2r_handle* h, uint64_t* out);
/* which: 0 = source, 1 = target.  cov6 = xx,xy,xz,yy,yz,zz per point; knn (n*k int32, ascending distance) optional */
b2r_status b2r_debug_covariances(b2r_handle* h, int which, double* cov6_out, int32_t* knn_out);
/* VGICP voxel map of the target, sorted by (x,y,z) voxel coordinate; arrays sized for *V <= n_target voxels */
b2r_status b2r_debug_voxelmap(b2r_handle* h, int32_t* coords_out, int32_t* npts_out, double* mean_out, double* cov6_out, size_t* V);
/* linearize at pose T (4x4 row-major double): H 36 row-major, b 6, *err.  corr_out: FAST_GICP n int32 target index
 * or -1; FAST_VGICP n*3 int32 voxel coordinates with corr_valid[n] */
b2r_status b2r_debug_linearize(b2r_handle* h, const double* T_rowmajor, double* H, double* b, double* err, int32_t* corr_out,
                               uint8_t* corr_valid);
b2r_status b2r_debug_compute_error(b2r_handle* h, const double* T_lin_rowmajor, const double* T_trial_rowmajor, double* err);
/* NDT target cells sorted by dense index; icov 9 doubles (float-rounded values) */
b2r_status b2r_debug_ndt_grid(b2r_handle* h, int32_t* idx_out, int32_t* npts_out, double* mean_out, double* icov_out, int32_t* min_b,
                              int32_t* div_b, size_t* V);
b2r_status b2r_debug_ndt_derivatives(b2r_handle* h, const double* p6, double* score, double* grad6, double* hess36, int32_t* hits_out);
/* exact kNN of `queries` in cloud c (ascending squared distance) */
b2r_status b2r_debug_knn(b2r_handle* h, b2r_cloud* c, const float* queries_xyzi, size_t nq, int k, int32_t* idx_out, float* d2_out);
/* per-stage device time of the last align/align_batch in milliseconds: [0] cloud prep (grid+covariances+maps),
 * [1] optimiser loop, [2] fitness, [3] total */
b2r_status b2r_last_timings(const b2r_handle* h, float ms_out[4]);
/* CUDA events on the handle's own stream (the stream every kernel of this handle is launched on), for callers that
 * must time a region on the device: record into slot 0..7, then read the elapsed time between two slots. */
b2r_status b2r_event_record(b2r_handle* h, int slot);
b2r_status b2r_event_elapsed_ms(b2r_handle* h, int slot_start, int slot_stop, float* ms);
/* Per-kernel device timing (CUDA events around each launch of the named kernel family) for roofline reporting.
 * kernel_id: 0 knn_cov (kNN + covariances), 1 lsq_eval (GICP/VGICP correspondence+reduction pass), 2 ndt_eval,
 * 3 grid_build (count/scan/scatter), 4 voxel_reduce (VGICP/NDT per-voxel statistics), 5 fitness.
 * b2r_profile_read returns the summed duration, number of launches and summed algorithmic bytes since enable. */
b2r_status b2r_profile_enable(b2r_handle* h, int on);
b2r_status b2r_profile_read(b2r_handle* h, int kernel_id, double* ms_sum, uint64_t* launches, double* algorithmic_bytes);

#ifdef __cplusplus
}
#endif
#endif /* B2R_H_ */
