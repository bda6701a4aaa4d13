// Header-only C++ mirror of InformationMatrixCalculator::calc_information_matrix
// (/root/reference/src/mrg_slam/information_matrix_calculator.cpp:14-44, weight(): :83-88) on top of the libb2r C ABI.
//
// The reference builds the 6x6 information matrix of every odometry and loop edge from a fitness score — a FLANN kd-tree build
// over cloud1 plus one nearest-neighbour query per point of cloud2 on one host thread (calc_fitness_score, :46-81; callers
// graph_database.cpp:140,580).  Here the score is ONE b2r_fitness_pair call on device-resident clouds; the weighting around it is
// restated statement by statement.  No Eigen / PCL dependency: `relpose` is the column-major 4x4 of Eigen::Isometry3d::matrix(),
// the result the row-major (= column-major: it is diagonal) 6x6 of the returned Eigen::MatrixXd.
#pragma once
#include <cmath>

#include "../b2r.h"

namespace b2r {

// the node parameters calc_information_matrix reads, with the values of config/mrg_slam.yaml:216-223 and :173
struct InformationMatrixParams {
  bool use_const_inf_matrix = false;
  double const_stddev_x = 0.5, const_stddev_q = 0.1;
  double var_gain_a = 2.0;
  double min_stddev_x = 0.1, max_stddev_x = 0.75;
  double min_stddev_q = 0.05, max_stddev_q = 0.2;
  double fitness_score_thresh = 1.25;
};

// InformationMatrixCalculator::weight (:83-88)
inline double information_weight(double a, double max_x, double min_y, double max_y, double x) {
  const double y = (1.0 - std::exp(-a * x)) / (1.0 - std::exp(-a * max_x));
  return min_y + (max_y - min_y) * y;
}

// inf36: 36 doubles.  cloud1 = the edge's first keyframe (the kd-tree side), cloud2 = the second (transformed by relpose).
// fitness_out (optional) receives calc_fitness_score(cloud1, cloud2, relpose) (max_range = numeric_limits<double>::max()).
inline b2r_status calc_information_matrix(b2r_handle* h, b2r_cloud* cloud1, b2r_cloud* cloud2, const double relpose_colmajor[16],
                                          const InformationMatrixParams& p, double inf36[36], double* fitness_out = nullptr) {
  for (int i = 0; i < 36; ++i) inf36[i] = (i % 7 == 0) ? 1.0 : 0.0;
  if (p.use_const_inf_matrix) {  // :18-23
    for (int i = 0; i < 3; ++i) inf36[i * 7] /= p.const_stddev_x;
    for (int i = 3; i < 6; ++i) inf36[i * 7] /= p.const_stddev_q;
    return B2R_OK;
  }
  float T[16];
  for (int i = 0; i < 16; ++i) T[i] = (float)relpose_colmajor[i];  // relpose.cast<float>() (:58)
  double fitness = 0.0;
  const b2r_status st = b2r_fitness_pair(h, cloud1, cloud2, T, 1.79769313486231570815e+308, &fitness);
  if (st != B2R_OK) return st;
  if (fitness_out) *fitness_out = fitness;
  const double min_var_x = std::pow(p.min_stddev_x, 2), max_var_x = std::pow(p.max_stddev_x, 2);
  const double min_var_q = std::pow(p.min_stddev_q, 2), max_var_q = std::pow(p.max_stddev_q, 2);
  const double w_x = information_weight(p.var_gain_a, p.fitness_score_thresh, min_var_x, max_var_x, fitness);
  const double w_q = information_weight(p.var_gain_a, p.fitness_score_thresh, min_var_q, max_var_q, fitness);
  for (int i = 0; i < 3; ++i) inf36[i * 7] /= w_x;
  for (int i = 3; i < 6; ++i) inf36[i * 7] /= w_q;
  return B2R_OK;
}

}  // namespace b2r
