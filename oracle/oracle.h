/* ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
 *
 * C interface of the CPU restatement of mrg_slam's scan-registration hot path.
 * The reference's own arithmetic lives in PCL / ndt_omp / fast_gicp, none of
 * which are vendored or pinned in /root/reference (CMakeLists.txt:26,84-85;
 * docker/noetic/Dockerfile:14-15), so these functions restate the published
 * upstream algorithms and anchor on the reference's call
 * sites: src/mrg_slam/registrations.cpp:46-148,
 * apps/scan_matching_odometry_component.cpp:195-350,
 * src/mrg_slam/loop_detector.cpp:97-180, apps/prefiltering_component.cpp:149-229.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library.  The product (libb2r.so) never does.
 *
 * Points are packed float32 x,y,z,intensity (16 B), the KITTI layout shown at
 * python_scripts/kitti_singlerobot_processor.py:174-183.  4x4 transforms are
 * column-major float (Eigen::Matrix4f).
 */
#ifndef ORACLE_H_
#define ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_NDT_OMP = 0, ORC_FAST_GICP = 1, ORC_FAST_VGICP = 2, ORC_SMALL_GICP = 3 };
enum { ORC_DIRECT1 = 0, ORC_DIRECT7 = 1, ORC_DIRECT27 = 2, ORC_KDTREE = 3 };

typedef struct orc_params {
  int method;                     /* ORC_* */
  int num_threads;                /* reg_num_threads; 0 => omp max */
  double transformation_epsilon;  /* reg_transformation_epsilon */
  int maximum_iterations;         /* reg_maximum_iterations */
  double max_correspondence_distance; /* FAST_GICP only (registrations.cpp:61) */
  int correspondence_randomness;  /* k for covariance kNN */
  double resolution;              /* VGICP voxel / NDT leaf */
  int neighbor_search;            /* ORC_DIRECT* (NDT_OMP: reg_nn_search_method; VGICP upstream default DIRECT1) */
  double rotation_epsilon;        /* fast_gicp default 2e-3 */
  int lm_max_iterations;          /* fast_gicp default 10 */
  double lm_init_lambda_factor;   /* fast_gicp default 1e-9 */
  double ndt_step_size;           /* ndt_omp default 0.1 */
  double ndt_outlier_ratio;       /* ndt_omp default 0.55 */
} orc_params;

typedef struct orc_result {
  float T[16];       /* final transformation, column-major */
  int converged;
  int iterations;    /* nr_iterations_ as the upstream class leaves it */
  double error;      /* last LM error y0 (GICP/VGICP) or NDT score */
  int lm_evals;      /* linearize + compute_error passes executed (NDT: derivative passes) */
} orc_result;

typedef struct orc_reg orc_reg;

void orc_default_params(int method, orc_params* p);
orc_reg* orc_reg_create(const orc_params* p);
void orc_reg_destroy(orc_reg* r);
void orc_reg_set_target(orc_reg* r, const float* xyzi, int n);
void orc_reg_set_source(orc_reg* r, const float* xyzi, int n);
int orc_reg_align(orc_reg* r, const float* guess_colmajor, orc_result* out);
/* pcl::Registration::getFitnessScore(max_range) on the last alignment */
double orc_reg_fitness(orc_reg* r, double max_range);

/* ---- intermediates for parity tests ---- */
/* per-point kNN covariances (PLANE regularised), 6 unique doubles xx,xy,xz,yy,yz,zz; knn_idx (n*k) optional */
void orc_knn_covariances(const float* xyzi, int n, int k, double* cov6_out, int* knn_idx_out);
/* fast_gicp GaussianVoxelMap: returns V; arrays sorted by (x,y,z) voxel coordinate. capacity = n voxels */
int orc_vgicp_voxelmap(const float* xyzi, int n, const double* cov6, double resolution, int* coords_out, int* npts_out,
                       double* mean_out, double* cov6_out);
/* linearize at pose T (row-major 4x4 double): H(36 row-major), b(6), returns error.  corr_out: FAST_GICP => n ints
 * (target index or -1); FAST_VGICP => n*3 voxel coords and corr_valid (n) */
double orc_reg_linearize(orc_reg* r, const double* T_rowmajor, double* H, double* b, int* corr_out, uint8_t* corr_valid);
double orc_reg_compute_error(orc_reg* r, const double* T_rowmajor);
/* NDT voxel grid (VoxelGridCovariance): returns number of leaves; sorted by dense index */
int orc_ndt_grid(const float* xyzi, int n, double resolution, int* idx_out, int* npts_out, double* mean_out, double* icov_out,
                 int* min_b_out, int* div_b_out);
/* NDT derivatives at transform vector p(6) with source transformed by the float matrix built from p */
double orc_reg_ndt_derivatives(orc_reg* r, const double* p6, double* grad6, double* hess36, int* hits_out);

/* ---- filters ---- */
int orc_distance_filter(const float* xyzi, int n, double near_thresh, double far_thresh, float* out);
int orc_voxelgrid(const float* xyzi, int n, float leaf, int min_points_per_voxel, float* out, int* voxel_index_out);
int orc_radius_outlier(const float* xyzi, int n, double radius, int min_neighbors, uint8_t* keep);
int orc_statistical_outlier(const float* xyzi, int n, int mean_k, double stddev_mul, uint8_t* keep, float* distances_out,
                            double* thr_out);
void orc_transform_cloud(const float* xyzi, int n, const float* T_colmajor, float* out);
double orc_fitness_score(const float* target, int nt, const float* source, int ns, const float* T_colmajor, double max_range,
                         int* nr_out);
void orc_knn(const float* xyzi, int n, const float* queries, int nq, int k, int* idx_out, float* d2_out);

/* MapCloudGenerator::generate + pcl::ApproximateMeanVoxelGrid (both in-tree: src/mrg_slam/map_cloud_generator.cpp:14-86,
 * include/pcl/filters/ApproximateMeanVoxelGrid.hpp:63-126); see mapcloud.cpp */
int orc_map_cloud(const float* const* clouds, const int* n, const double* poses_colmajor, const uint8_t* first_keyframe, int count,
                  float resolution, int min_points_per_voxel, float distance_far_thresh, int skip_first_cloud, float* out,
                  int* voxel_keys_out);

/* ---- pcl::GeneralizedIterativeClosestPoint ("GICP" / "GICP_OMP", registrations.cpp:93-116) — gicp_pcl.cpp.
 * Oracle only: the product has no engine for this method yet. */
typedef struct orc_gicp_pcl_params {
  double transformation_epsilon;       /* reg_transformation_epsilon */
  int maximum_iterations;              /* reg_maximum_iterations */
  int use_reciprocal_correspondences;  /* reg_use_reciprocal_correspondences (false in the YAML; not restated) */
  double max_correspondence_distance;  /* reg_max_correspondence_distance */
  int correspondence_randomness;       /* reg_correspondence_randomness */
  int max_optimizer_iterations;        /* reg_max_optimizer_iterations (BFGS steps per outer iteration) */
  double rotation_epsilon;             /* PCL default 2e-3 */
  double gicp_epsilon;                 /* PCL default 1e-3 */
} orc_gicp_pcl_params;
void orc_gicp_pcl_default_params(orc_gicp_pcl_params* p);
int orc_gicp_pcl_align(const float* target, int nt, const float* source, int ns, const orc_gicp_pcl_params* p, const float* guess_colmajor,
                       orc_result* out);
void orc_gicp_pcl_apply_state(const double* x6, float* T_colmajor);
double orc_gicp_pcl_fdf(const float* src, const float* tgt, const int* idx_src, const int* idx_tgt, int m, const double* mahalanobis, int ns,
                        const double* x6, double* grad6);
int orc_gicp_pcl_bfgs(const float* src, const float* tgt, const int* idx_src, const int* idx_tgt, int m, const double* mahalanobis, int ns,
                      double* x6, int max_inner, double gradient_tol, int* status, double* f_out, int* evals);
void orc_gicp_pcl_covariances(const float* xyzi, int n, int k, double gicp_epsilon, double* cov9_out);
/* a persistent problem: correspondences for a given transformation_, then the functor on them (host test of the device
 * state machine in mrg_slam_b200/csrc/gicp_pcl_sm.hpp).  target / source must outlive the object. */
typedef struct orc_gicp_pcl orc_gicp_pcl;
orc_gicp_pcl* orc_gicp_pcl_create(const float* target, int nt, const float* source, int ns, const orc_gicp_pcl_params* p,
                                  const float* guess_colmajor);
void orc_gicp_pcl_destroy(orc_gicp_pcl* o);
int orc_gicp_pcl_correspond(orc_gicp_pcl* o, const float* transformation_colmajor);
double orc_gicp_pcl_eval(orc_gicp_pcl* o, const double* x6, double* grad6);

/* ---- upstream-version variants: the alternative reading of every detail SURVEY.md Appendix A marks as version-dependent.
 * Test infrastructure for tests/test_oracle_variants.py (how far does a result move if upstream differs?); 0 = documented choice. */
enum {
  ORC_VAR_VGICP_COORD_NO_HALF = 0, /* voxel_coord = floor(x / res) instead of floor(x / res - 0.5)                      (A.2) */
  ORC_VAR_NDT_ANGLE_EPS_1E5 = 1,   /* small-angle cut-off 1e-5 instead of the literal 10e-5                              (A.3) */
  ORC_VAR_NDT_INNER_DOUBLE = 2,    /* updateDerivatives / point derivatives in double (PCL, older ndt_omp) instead of float (A.3) */
  ORC_VAR_MT_CLAMP_MAX_FIRST = 3,  /* a_t = min(max(a_t, step_min), step_max) instead of max(min(a_t, step_max), step_min)  (A.3) */
  ORC_VAR_NDT_COV_NEWER_PCL = 4,   /* leaf covariance (cov - pt_sum mean^T) / (n - 1) instead of the older single-pass form  (A.4) */
  ORC_VAR_NDT_LOOKUP_MUL = 5,      /* neighbourhood lookup key floor(p * inv_leaf) instead of floor(p / leaf)               (A.4) */
  ORC_VAR_VOXELGRID_DESCENDING = 6,/* points of a voxel summed in descending instead of ascending index order (unstable sort) (A.6) */
  ORC_VAR_RADIUS_NONSTRICT = 7,    /* radius search d2 <= r2 instead of d2 < r2                                             (A.7) */
  ORC_VAR_TRANSFORM_LEFT_TO_RIGHT = 8, /* transformPointCloud ((x c0 + y c1) + z c2) + c3 instead of the SSE association      (A.10) */
  ORC_VAR_NORM_LEFT_TO_RIGHT = 9,  /* distance filter norm (x^2 + y^2) + z^2 instead of x^2 + (y^2 + z^2)                    (A.11) */
  ORC_VAR_EULER_NO_FIXUP = 10,     /* eulerAngles(0,1,2) without the first-angle fix-up of Eigen >= 3.3                     (A.3) */
  ORC_VAR_COUNT = 11
};
void orc_set_variant(int id, int value);
int orc_get_variant(int id);

void orc_set_num_threads(int n);
int orc_get_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
