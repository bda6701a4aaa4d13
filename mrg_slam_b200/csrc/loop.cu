// The optimiser loop of every registration method as ONE CUDA-graph launch (common.cuh: LoopCtl / loop_tail).
//
// Replaces the host loops of fast_gicp LsqRegistration::computeTransformation, pclomp NDT::computeTransformation and
// pcl GICP::computeTransformation (selected at /root/reference/src/mrg_slam/registrations.cpp:46-148): there the host runs
// "linearise, solve, test convergence" iteration by iteration; here the graph's WHILE node repeats {eval kernel, step kernel}
// and the step kernel's last block clears the condition when every pair of the batch has finished.
#include <cstdlib>

#include "internal.hpp"

namespace b2r {

static bool graph_loop_enabled() {
  static const bool on = [] { const char* e = getenv("B2R_GRAPH_LOOP"); return !e || atoi(e) != 0; }();
  return on;
}

#define B2R_GRAPH(expr, g, ge)                                                                              \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) {                                                                                \
      if (ge) cudaGraphExecDestroy(ge);                                                                     \
      if (g) cudaGraphDestroy(g);                                                                           \
      char _b[512];                                                                                         \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      throw ::b2r::Error(B2R_ERR_CUDA, _b);                                                                 \
    }                                                                                                       \
  } while (0)

void run_device_loop(Ctx& ctx, const void* eval_fn, dim3 eval_grid, dim3 eval_block, void** eval_args, const void* step_fn, dim3 step_grid,
                     dim3 step_block, void** step_args, LoopArgs& la, LoopCtl* d_ctl, int npairs, long max_rounds, int prof_id,
                     const void* eval2_fn) {
  la.ctl = d_ctl;
  la.npairs = npairs;
  la.max_rounds = (int)std::min<long>(max_rounds, 1l << 30);
  la.handle = 0;
  la.rounds_total = ctx.d_graph_rounds;
  la.kernels_per_round = eval2_fn ? 3 : 2;
  if (graph_loop_enabled() && !ctx.profile) {
    la.use_graph = 1;
    ctx.reap_graphs(false);
    cudaKernelNodeParams ke = {}, ke2 = {}, ks = {};
    ke.func = const_cast<void*>(eval_fn); ke.gridDim = eval_grid; ke.blockDim = eval_block; ke.kernelParams = eval_args;
    ke2 = ke; ke2.func = const_cast<void*>(eval2_fn);
    ks.func = const_cast<void*>(step_fn); ks.gridDim = step_grid; ks.blockDim = step_block; ks.kernelParams = step_args;
    // ---- a graph of this kernel pair instantiated earlier: rewrite the two nodes' parameters and launch
    if (ctx.loop_graph_cache_ok) {
      for (Ctx::LoopGraph& lg : ctx.loop_graphs) {
        if (lg.eval_fn != eval_fn || lg.eval2_fn != eval2_fn || lg.step_fn != step_fn) continue;
        la.handle = lg.handle;  // baked into the step kernel's arguments below
        if (cudaGraphExecKernelNodeSetParams(lg.exec, lg.eval_node, &ke) == cudaSuccess &&
            (!eval2_fn || cudaGraphExecKernelNodeSetParams(lg.exec, lg.eval2_node, &ke2) == cudaSuccess) &&
            cudaGraphExecKernelNodeSetParams(lg.exec, lg.step_node, &ks) == cudaSuccess) {
          B2R_CUDA(cudaGraphLaunch(lg.exec, ctx.stream));
          ctx.launches += 1;
          ++ctx.graph_launches;
          return;
        }
        cudaGetLastError();
        ctx.loop_graph_cache_ok = false;  // build a fresh graph per call from now on (the stale entries go with the handle)
        break;
      }
    }
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    B2R_GRAPH(cudaGraphCreate(&g, 0), g, ge);
    B2R_GRAPH(cudaGraphConditionalHandleCreate(&la.handle, g, 1, cudaGraphCondAssignDefault), g, ge);
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = la.handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    cudaGraphNode_t wnode = nullptr;
    B2R_GRAPH(cudaGraphAddNode(&wnode, g, nullptr, 0, &cp), g, ge);
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    cudaGraphNode_t ne = nullptr, ne2 = nullptr, ns = nullptr;
    B2R_GRAPH(cudaGraphAddKernelNode(&ne, body, nullptr, 0, &ke), g, ge);
    if (eval2_fn) {
      // the two evaluation kernels touch disjoint pairs (each serves the pairs of one phase), so they are independent nodes: in
      // the lock-step early rounds one of them has nothing to do and its launch hides behind the other instead of preceding it
      static const bool parallel = [] { const char* e = getenv("B2R_LOOP_PARALLEL_EVAL"); return !e || atoi(e) != 0; }();
      B2R_GRAPH(cudaGraphAddKernelNode(&ne2, body, parallel ? nullptr : &ne, parallel ? 0 : 1, &ke2), g, ge);
      cudaGraphNode_t deps[2] = {ne, ne2};
      B2R_GRAPH(cudaGraphAddKernelNode(&ns, body, deps, 2, &ks), g, ge);
    } else {
      B2R_GRAPH(cudaGraphAddKernelNode(&ns, body, &ne, 1, &ks), g, ge);
    }
    B2R_GRAPH(cudaGraphInstantiate(&ge, g, 0), g, ge);
    B2R_GRAPH(cudaGraphLaunch(ge, ctx.stream), g, ge);
    ctx.launches += 1;  // one graph launch; the rounds it ran are counted on the device (LoopArgs::rounds_total)
    ++ctx.graph_launches;
    if (ctx.loop_graph_cache_ok) ctx.loop_graphs.push_back(Ctx::LoopGraph{eval_fn, eval2_fn, step_fn, g, ge, ne, ne2, ns, la.handle});
    else ctx.graph_graveyard.emplace_back(ge, g);  // released once the stream has drained (destroying it now would wait for the launch)
    return;
  }
  // ---- host-polled loop: groups of rounds, then one read of the done counter
  la.use_graph = 0;
  int rounds_per_check = 6;
  long rounds = 0;
  int hdone = 0;
  while (hdone < npairs && rounds < max_rounds) {
    for (int r = 0; r < rounds_per_check; ++r) {
      {
        ProfScope ps(ctx, prof_id >= 0 ? prof_id : 0, 0.0, prof_id >= 0);  // bytes are added by the caller from the work the device did
        B2R_CUDA(cudaLaunchKernel(eval_fn, eval_grid, eval_block, eval_args, 0, ctx.stream));
        ++ctx.launches;
        if (eval2_fn) {
          B2R_CUDA(cudaLaunchKernel(eval2_fn, eval_grid, eval_block, eval_args, 0, ctx.stream));
          ++ctx.launches;
        }
      }
      B2R_CUDA(cudaLaunchKernel(step_fn, step_grid, step_block, step_args, 0, ctx.stream));
      ++ctx.launches;
    }
    rounds += rounds_per_check;
    B2R_CUDA(cudaMemcpyAsync(&hdone, &d_ctl->done, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(ctx.stream));
    rounds_per_check = 4;
  }
}

}  // namespace b2r
