// Host harness for mrg_slam_b200/csrc/gicp_pcl_sm.hpp (the device state machine of PCL's BFGS GICP): the machine's requests are
// answered by the ORACLE's own correspondence search and functor (oracle/gicp_pcl.cpp), so that it must reproduce
// orc_gicp_pcl_align bit for bit — the control flow is checked without a GPU.  Built and driven by tests/test_gicp_pcl_sm.py.
#include <cstring>

#include "../../mrg_slam_b200/csrc/gicp_pcl_sm.hpp"
#include "../../oracle/oracle.h"

extern "C" int gp_sm_align(const float* target, int nt, const float* source, int ns, const orc_gicp_pcl_params* p, const float* guess,
                           orc_result* out, int* rounds_out) {
  orc_gicp_pcl* o = orc_gicp_pcl_create(target, nt, source, ns, p, guess);
  if (!o) return -1;
  gp::Params prm;
  gp::default_params(prm);
  prm.transformation_epsilon = p->transformation_epsilon;
  prm.rotation_epsilon = p->rotation_epsilon;
  prm.maximum_iterations = p->maximum_iterations;
  prm.max_optimizer_iterations = p->max_optimizer_iterations;
  gp::State s;
  gp::init(s, guess);
  gp::advance(s, prm, 0, 0.0, nullptr);  // PH_START: the first request
  int rounds = 0;
  while (s.request != gp::REQ_DONE && rounds < 1000000) {
    ++rounds;
    if (s.request == gp::REQ_CORRESPOND) {
      const int m = orc_gicp_pcl_correspond(o, s.transformation);
      if (m < 0) { orc_gicp_pcl_destroy(o); return -1; }
      gp::advance(s, prm, m, 0.0, nullptr);
    } else if (s.request == gp::REQ_EVAL) {
      double g[6];
      const double f = orc_gicp_pcl_eval(o, s.xreq, g);
      gp::advance(s, prm, 0, f, g);
    } else {
      break;
    }
  }
  gp::final_transformation(s, out->T);
  out->converged = s.converged;
  out->iterations = s.nr_iterations;
  out->error = 0;
  out->lm_evals = s.evals;
  if (rounds_out) *rounds_out = rounds;
  orc_gicp_pcl_destroy(o);
  return s.request == gp::REQ_DONE ? 0 : -2;
}
