// C ABI of libb2r (include/b2r.h): handle / cloud lifecycle, the pcl::Registration surface, the loop-closure batch
// call, the prefilter chain and introspection entry points.  All work is delegated to the CUDA modules; there is
// no CPU code path for any computation.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <atomic>
#include <map>
#include <mutex>
#include <cstring>

#include "internal.hpp"

using namespace b2r;

namespace b2r {
// One private pool per device, shared by every handle of the process, never trimmed (release threshold = max): a batch
// re-uses the blocks of the previous one without going back to the driver.
cudaMemPool_t device_pool() {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) throw Error(B2R_ERR_CUDA, "cudaGetDevice failed");
  std::lock_guard<std::mutex> lk(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    B2R_CUDA(cudaMemPoolCreate(&pools[dev], &props));
    uint64_t thr = UINT64_MAX;
    B2R_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &thr));
  }
  return pools[dev];
}
}  // namespace b2r


namespace {

Needs needs_for(const b2r_config& cfg, bool is_target, bool want_fitness) {
  Needs nd;
  switch (cfg.method) {
    case B2R_FAST_VGICP:
      nd.cov_k = cfg.correspondence_randomness;
      if (is_target) nd.vres = cfg.resolution;
      break;
    case B2R_SMALL_GICP:
    case B2R_FAST_GICP:
      nd.cov_k = cfg.correspondence_randomness;
      if (is_target) nd.grid = true;
      break;
    case B2R_GICP_PCL:  // PCL's own covariance arithmetic (float products, E[xx^T] - mean mean^T)
      nd.cov_k = cfg.correspondence_randomness;
      nd.cov_mode = 1;
      if (is_target) nd.grid = true;
      break;
    default:  // NDT_OMP
      if (is_target) { nd.leaf = (float)cfg.resolution; nd.leaf_centroids = cfg.neighbor_search == B2R_KDTREE; }
      break;
  }
  // getFitnessScore: the target's NN grid is searched; the SOURCE's grid only provides the order its points are queried in (cell
  // order: a warp's queries are neighbours in space) — 0.3 ms of grid builds against 4 ms of fitness on the 4096-pair NDT batch
  if (want_fitness) nd.grid = true;
  return nd;
}

template <typename F>
b2r_status guarded(b2r_handle* hh, F&& f) {
  if (!hh) return B2R_ERR_INVALID_ARG;
  Handle& h = hh->h;
  try {
    if (cudaSetDevice(h.ctx.device) != cudaSuccess) throw Error(B2R_ERR_CUDA, "cudaSetDevice failed");
    f(h);
    return B2R_OK;
  } catch (const Error& e) {
    h.last_error = e.what();
    cudaGetLastError();
    return e.status;
  } catch (const std::exception& e) {
    h.last_error = e.what();
    return B2R_ERR_INVALID_ARG;
  }
}

void check_cfg(const b2r_config& cfg) {
  if (cfg.neighbor_search < B2R_DIRECT1 || cfg.neighbor_search > B2R_KDTREE) throw Error(B2R_ERR_INVALID_ARG, "unknown neighbor_search");
  if (cfg.neighbor_search == B2R_KDTREE && cfg.method != B2R_NDT_OMP) throw Error(B2R_ERR_INVALID_ARG, "KDTREE search is an NDT mode");
  if (cfg.method < B2R_NDT_OMP || cfg.method > B2R_GICP_PCL) throw Error(B2R_ERR_INVALID_ARG, "unknown method");
  if (cfg.resolution <= 0) throw Error(B2R_ERR_INVALID_ARG, "resolution must be > 0");
  if (cfg.method != B2R_NDT_OMP && (cfg.correspondence_randomness < 4 || cfg.correspondence_randomness > 32))
    throw Error(B2R_ERR_INVALID_ARG, "correspondence_randomness must be in [4,32]");
}

// A pair the configured method cannot run on: an empty cloud, or (GICP family) fewer points than the covariance neighbourhood.
// The reference aligns candidates one by one (loop_detector.cpp:126-145), so a degenerate candidate only loses itself: such a
// pair comes back as converged = 0, T = guess, fitness = DBL_MAX and the rest of the batch runs.
bool degenerate_pair(const b2r_config& cfg, const Cloud* s, const Cloud* t) {
  const int need = cfg.method == B2R_NDT_OMP ? 1 : cfg.correspondence_randomness;
  return s->n < need || t->n < need;
}
void fail_result(const float* guess, b2r_result& r) {
  memset(&r, 0, sizeof(r));
  memcpy(r.T, guess, 64);
  r.converged = 0; r.iterations = 0; r.error = 0.0; r.evals = 0; r.fitness = DBL_MAX;
}

// Runs the optimiser on a list of (source, target, guess) triples; clouds are prepared (batched) first.  Everything between
// the upload of the pair table and the final copy of the result rows is stream-ordered device work: structure builds, the
// optimiser loop (one graph launch), result rows, fitness.
//   out_all   host array of n rows, or nullptr
//   rows_dev  device array of n rows the results are (also) left in, or nullptr — the all-gather send buffer of a sharded batch
// With out_all == nullptr the call returns without synchronising.
void run_align(Handle& h, const std::vector<Cloud*>& sources_in, const std::vector<Cloud*>& targets_in, const float* guesses_in, int with_fitness,
               double fitness_max_range, b2r_result* out_all, b2r_result* rows_dev) {
  Ctx& ctx = h.ctx;
  HostTrace tr;
  B2R_CUDA(cudaEventRecord(h.ev[0], ctx.stream));
  h.timings_pending = true;
  const size_t n_all = sources_in.size();
  // ---- per-pair validation: only the usable pairs go to the device
  std::vector<int> live;
  std::vector<Cloud*> sources, targets;
  std::vector<b2r_result> failed;  // rows of the pairs that did not run (only materialised when there are any)
  for (size_t i = 0; i < n_all; ++i) {
    if (!sources_in[i] || !targets_in[i]) throw Error(B2R_ERR_INVALID_ARG, "null cloud");
    if (degenerate_pair(h.cfg, sources_in[i], targets_in[i])) continue;
    live.push_back((int)i);
    sources.push_back(sources_in[i]);
    targets.push_back(targets_in[i]);
  }
  const int np = (int)live.size();
  const bool all_live = (size_t)np == n_all;
  std::vector<Cloud*> uniq;
  std::vector<Needs> needs;
  static std::atomic<uint64_t> epoch_counter{0};
  const uint64_t epoch = ++epoch_counter;  // marks the clouds already listed in this call (O(1) per pair, no map)
  const Needs nd_src = needs_for(h.cfg, false, with_fitness != 0), nd_tgt = needs_for(h.cfg, true, with_fitness != 0);
  auto add = [&](Cloud* c, bool is_target) {
    int id;
    if (c->batch_epoch != epoch) {
      c->batch_epoch = epoch;
      c->batch_slot = id = (int)uniq.size();
      uniq.push_back(c);
      needs.push_back(Needs());
    } else {
      id = c->batch_slot;
    }
    const Needs& nd = is_target ? nd_tgt : nd_src;
    Needs& cur = needs[id];
    cur.grid = cur.grid || nd.grid;
    cur.cov_k = std::max(cur.cov_k, nd.cov_k);
    cur.cov_mode = std::max(cur.cov_mode, nd.cov_mode);
    if (nd.vres > 0) cur.vres = nd.vres;
    if (nd.leaf > 0) cur.leaf = nd.leaf;
    cur.leaf_centroids = cur.leaf_centroids || nd.leaf_centroids;
    return id;
  };
  // ---- one host block [pairs | source sizes | guesses] -> one H2D copy
  const size_t off_n = sizeof(PairDesc) * (size_t)np, off_g = off_n + sizeof(int) * (size_t)np, tot = off_g + 64 * (size_t)np;
  std::vector<uint8_t> blk(std::max<size_t>(tot, 1));
  PairDesc* pairs = reinterpret_cast<PairDesc*>(blk.data());
  int* src_sizes = reinterpret_cast<int*>(blk.data() + off_n);
  float* guesses = reinterpret_cast<float*>(blk.data() + off_g);
  int maxn = 1;
  for (int i = 0; i < np; ++i) {
    pairs[i].src = add(sources[i], false);
    pairs[i].tgt = add(targets[i], true);
    src_sizes[i] = sources[i]->n;
    maxn = std::max(maxn, sources[i]->n);
    memcpy(guesses + (size_t)i * 16, guesses_in + (size_t)live[i] * 16, 64);
  }
  DBuf<uint8_t> dblk;
  DBuf<b2r_result> rows_tmp;
  b2r_result* d_rows = nullptr;
  tr.mark("pairs");
  if (np > 0) {
    dblk.alloc(tot, ctx.stream);
    B2R_CUDA(cudaMemcpyAsync(dblk.p, blk.data(), tot, cudaMemcpyHostToDevice, ctx.stream));
    if (rows_dev && all_live) d_rows = rows_dev;
    else { rows_tmp.alloc(np, ctx.stream); d_rows = rows_tmp.p; }
    DBuf<CloudView> dv;
    {
      // clouds of a split upload (cloud.cu): the search structures of those that have arrived are built (enqueued) first, while
      // the copy of the others is still running on the copy stream; the second call then waits for them and builds the rest
      std::vector<Cloud*> ready;
      std::vector<Needs> ready_needs;
      bool waiting = false;
      for (size_t i = 0; i < uniq.size(); ++i) {
        if (uniq[i]->pending && !uniq[i]->pending->resolved) waiting = true;
        else { ready.push_back(uniq[i]); ready_needs.push_back(needs[i]); }
      }
      if (waiting && !ready.empty()) {
        DBuf<CloudView> first;
        clouds_prepare(ctx, h.cfg, ready, ready_needs, first);
        tr.mark("prepare(arrived)");
      }
    }
    clouds_prepare(ctx, h.cfg, uniq, needs, dv);
    tr.mark("prepare");
    B2R_CUDA(cudaEventRecord(h.ev[1], ctx.stream));
    BatchArgs b;
    b.d_views = dv.p;
    b.d_pairs = reinterpret_cast<const PairDesc*>(dblk.p);
    b.d_src_n = reinterpret_cast<const int*>(dblk.p + off_n);
    b.d_guesses = reinterpret_cast<const float*>(dblk.p + off_g);
    b.guesses = guesses;
    b.d_rows = d_rows;
    b.np = np;
    b.maxn = maxn;
    b.src_sizes = src_sizes;
    if (h.cfg.method == B2R_NDT_OMP) ndt_align_batch(ctx, h.cfg, b);
    else if (h.cfg.method == B2R_GICP_PCL) gicp_pcl_align_batch(ctx, h.cfg, b);
    else lsq_align_batch(ctx, h.cfg, b);
    B2R_CUDA(cudaEventRecord(h.ev[2], ctx.stream));
    tr.mark("optimise");
    if (with_fitness) fitness_batch(ctx, b, fitness_max_range);
    B2R_CUDA(cudaEventRecord(h.ev[3], ctx.stream));
    tr.mark("fitness");
    // dv, dblk and the optimisers' buffers are released stream-ordered: after everything enqueued above
    if (all_live) {
      if (out_all) {
        const size_t tot = sizeof(b2r_result) * np;
        void* stage = ctx.pinned_buf(tot);
        B2R_CUDA(cudaMemcpyAsync(stage ? stage : (void*)out_all, d_rows, tot, cudaMemcpyDeviceToHost, ctx.stream));
        B2R_CUDA(cudaStreamSynchronize(ctx.stream));
        if (stage) memcpy(out_all, stage, tot);
        tr.mark("rows_d2h+sync");
      }
      tr.print("run_align (host ms; device work is asynchronous)");
      return;
    }
  } else {
    for (int e = 1; e < 4; ++e) B2R_CUDA(cudaEventRecord(h.ev[e], ctx.stream));
  }
  // ---- some pairs were degenerate: merge on the host (rare path)
  std::vector<b2r_result> merged(n_all), got(np);
  if (np > 0) {
    B2R_CUDA(cudaMemcpyAsync(got.data(), d_rows, sizeof(b2r_result) * np, cudaMemcpyDeviceToHost, ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  for (size_t i = 0; i < n_all; ++i) fail_result(guesses_in + i * 16, merged[i]);
  for (int j = 0; j < np; ++j) merged[live[j]] = got[j];
  if (out_all) memcpy(out_all, merged.data(), sizeof(b2r_result) * n_all);
  if (rows_dev) {
    B2R_CUDA(cudaMemcpyAsync(rows_dev, merged.data(), sizeof(b2r_result) * n_all, cudaMemcpyHostToDevice, ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(ctx.stream));  // `merged` goes out of scope
  }
}

// two-entry view array {source, target} of the handle's current clouds, prepared for the configured method
void prepare_current(Handle& h, DBuf<CloudView>& dv, bool want_fitness) {
  if (!h.source || !h.target) throw Error(B2R_ERR_STATE, "source and target must be set first");
  std::vector<Cloud*> cl{h.source, h.target};
  std::vector<Needs> nd{needs_for(h.cfg, false, false), needs_for(h.cfg, true, want_fitness)};
  clouds_prepare(h.ctx, h.cfg, cl, nd, dv);  // source == target (same object) still yields two identical views
}

}  // namespace

namespace b2r {
void b2r_run_align(Handle& h, const std::vector<Cloud*>& sources, const std::vector<Cloud*>& targets, const float* guesses_colmajor,
                   int with_fitness, double fitness_max_range, b2r_result* out_all, b2r_result* rows_dev) {
  run_align(h, sources, targets, guesses_colmajor, with_fitness, fitness_max_range, out_all, rows_dev);
}
}  // namespace b2r

extern "C" {

const char* b2r_version(void) { return "b2r 0.2 (sm_100a)"; }

b2r_status b2r_default_config(int method, b2r_config* cfg) {
  if (!cfg) return B2R_ERR_INVALID_ARG;
  memset(cfg, 0, sizeof(*cfg));
  cfg->method = method;
  cfg->device = 0;
  cfg->transformation_epsilon = 0.1;       // config/mrg_slam.yaml:102
  cfg->maximum_iterations = 64;            // :103
  cfg->max_correspondence_distance = 2.0;  // :104
  cfg->correspondence_randomness = 20;     // :107
  cfg->resolution = 1.0;                   // :108
  cfg->neighbor_search = method == B2R_NDT_OMP ? B2R_DIRECT7 : B2R_DIRECT1;  // :109 / fast_gicp default
  cfg->rotation_epsilon = 2e-3;
  cfg->lm_max_iterations = 10;
  cfg->lm_init_lambda_factor = 1e-9;
  cfg->ndt_step_size = 0.1;
  cfg->ndt_outlier_ratio = 0.55;
  cfg->nn_cell_size = 0.0;
  cfg->max_optimizer_iterations = 20;      // config/mrg_slam.yaml:105
  cfg->gicp_epsilon = 1e-3;
  return B2R_OK;
}

b2r_status b2r_create(const b2r_config* cfg, b2r_handle** out) {
  if (!cfg || !out) return B2R_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return B2R_ERR_NO_DEVICE;  // no CPU fallback
  }
  if (cfg->device < 0 || cfg->device >= ndev) return B2R_ERR_INVALID_ARG;
  b2r_handle* hh = new b2r_handle();
  Handle& h = hh->h;
  try {
    check_cfg(*cfg);
    h.cfg = *cfg;
    h.ctx.device = cfg->device;
    B2R_CUDA(cudaSetDevice(cfg->device));
    B2R_CUDA(cudaStreamCreateWithFlags(&h.ctx.stream, cudaStreamNonBlocking));
    h.ctx.stream_owner = std::make_shared<StreamOwner>();
    h.ctx.stream_owner->s = h.ctx.stream;
    h.ctx.stream_owner->device = cfg->device;
    int sms = 0;
    B2R_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
    h.ctx.num_sms = sms > 0 ? sms : 148;
    B2R_CUDA(cudaMalloc((void**)&h.ctx.d_graph_rounds, sizeof(unsigned long long)));
    B2R_CUDA(cudaMemset(h.ctx.d_graph_rounds, 0, sizeof(unsigned long long)));
    (void)device_pool();  // the library's private memory pool of this device (the process's default pool is not touched)
    for (int i = 0; i < 4; ++i) B2R_CUDA(cudaEventCreate(&h.ev[i]));
    for (int i = 0; i < 16; ++i) h.final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
  } catch (const Error& e) {
    b2r_status s = e.status;
    delete hh;
    return s;
  }
  *out = hh;
  return B2R_OK;
}

void b2r_destroy(b2r_handle* hh) {
  if (!hh) return;
  Handle& h = hh->h;
  cudaSetDevice(h.ctx.device);
  if (h.ctx.stream) cudaStreamSynchronize(h.ctx.stream);
  h.owned_source.reset();
  h.owned_target.reset();
  for (int i = 0; i < 4; ++i)
    if (h.ev[i]) cudaEventDestroy(h.ev[i]);
  for (int i = 0; i < 8; ++i)
    if (h.user_ev[i]) cudaEventDestroy(h.user_ev[i]);
  h.ctx.prof_resolve();
  h.ctx.reap_graphs(true);
  if (auto pb = h.ctx.pending_boxes.lock()) pb->resolve();  // an upload still in flight uses the copy stream and the pinned area
  if (h.ctx.copy_stream) { cudaStreamSynchronize(h.ctx.copy_stream); cudaStreamDestroy(h.ctx.copy_stream); h.ctx.copy_stream = nullptr; }
  if (h.ctx.pinned_boxes) { cudaFreeHost(h.ctx.pinned_boxes); h.ctx.pinned_boxes = nullptr; h.ctx.pinned_boxes_cap = 0; }
  if (h.ctx.pinned) { cudaFreeHost(h.ctx.pinned); h.ctx.pinned = nullptr; h.ctx.pinned_bytes = 0; }
  if (h.ctx.d_graph_rounds) cudaFree(h.ctx.d_graph_rounds);
  for (cudaEvent_t e : h.ctx.ev_pool) cudaEventDestroy(e);
  h.ctx.ev_pool.clear();
  if (h.ctx.stream) cudaStreamSynchronize(h.ctx.stream);
  h.ctx.stream_owner.reset();  // the stream itself goes away with the last allocation made on it (clouds may outlive the handle)
  delete hh;
}

const char* b2r_last_error(const b2r_handle* hh) { return hh ? hh->h.last_error.c_str() : "null handle"; }

b2r_status b2r_cloud_create(b2r_handle* hh, const void* points, size_t n, size_t stride_bytes, int memspace, b2r_cloud** out) {
  if (!out) return B2R_ERR_INVALID_ARG;
  *out = nullptr;
  return guarded(hh, [&](Handle& h) {
    if (!points && n) throw Error(B2R_ERR_INVALID_ARG, "null points");
    b2r_cloud* c = new b2r_cloud();
    try {
      Cloud* cp = &c->c;
      clouds_upload(h.ctx, &cp, &points, &n, 1, stride_bytes, memspace);
    } catch (...) {
      delete c;
      throw;
    }
    *out = c;
  });
}

b2r_status b2r_cloud_create_batch(b2r_handle* hh, const void* const* points, const size_t* n, size_t count, size_t stride_bytes, int memspace,
                                  b2r_cloud** out) {
  if (!out || (count && (!points || !n))) return B2R_ERR_INVALID_ARG;
  for (size_t i = 0; i < count; ++i) out[i] = nullptr;
  return guarded(hh, [&](Handle& h) {
    try {
      std::vector<Cloud*> cl(count);
      for (size_t i = 0; i < count; ++i) {
        if (!points[i] && n[i]) throw Error(B2R_ERR_INVALID_ARG, "null points");
        out[i] = new b2r_cloud();
        cl[i] = &out[i]->c;
      }
      clouds_upload(h.ctx, cl.data(), points, n, count, stride_bytes, memspace);
    } catch (...) {
      for (size_t i = 0; i < count; ++i) { delete out[i]; out[i] = nullptr; }
      throw;
    }
  });
}

void b2r_cloud_destroy(b2r_cloud* c) {
  if (!c) return;
  cudaSetDevice(c->c.device);
  delete c;
}
void b2r_cloud_destroy_batch(b2r_cloud* const* clouds, size_t count) {
  if (!clouds) return;
  int dev = -1;
  for (size_t i = 0; i < count; ++i) {
    if (!clouds[i]) continue;
    if (clouds[i]->c.device != dev) { dev = clouds[i]->c.device; cudaSetDevice(dev); }
    delete clouds[i];
  }
}
size_t b2r_cloud_size(const b2r_cloud* c) { return c ? (size_t)c->c.n : 0; }

b2r_status b2r_set_target(b2r_handle* hh, const void* points, size_t n, size_t stride_bytes, int memspace) {
  return guarded(hh, [&](Handle& h) {
    if (!points || n == 0) throw Error(B2R_ERR_INVALID_ARG, "empty target");
    std::unique_ptr<Cloud> c(new Cloud());
    Cloud* cp = c.get();
    clouds_upload(h.ctx, &cp, &points, &n, 1, stride_bytes, memspace);
    h.owned_target = std::move(c);
    h.target = h.owned_target.get();
  });
}
b2r_status b2r_set_source(b2r_handle* hh, const void* points, size_t n, size_t stride_bytes, int memspace) {
  return guarded(hh, [&](Handle& h) {
    if (!points || n == 0) throw Error(B2R_ERR_INVALID_ARG, "empty source");
    std::unique_ptr<Cloud> c(new Cloud());
    Cloud* cp = c.get();
    clouds_upload(h.ctx, &cp, &points, &n, 1, stride_bytes, memspace);
    h.owned_source = std::move(c);
    h.source = h.owned_source.get();
  });
}
b2r_status b2r_set_target_cloud(b2r_handle* hh, b2r_cloud* c) {
  return guarded(hh, [&](Handle& h) {
    if (!c) throw Error(B2R_ERR_INVALID_ARG, "null cloud");
    h.target = &c->c;
  });
}
b2r_status b2r_set_source_cloud(b2r_handle* hh, b2r_cloud* c) {
  return guarded(hh, [&](Handle& h) {
    if (!c) throw Error(B2R_ERR_INVALID_ARG, "null cloud");
    h.source = &c->c;
  });
}

b2r_status b2r_align(b2r_handle* hh, const float guess[16], b2r_result* out) {
  b2r_status st = guarded(hh, [&](Handle& h) {
    if (!guess || !out) throw Error(B2R_ERR_INVALID_ARG, "null argument");
    if (!h.source || !h.target) throw Error(B2R_ERR_STATE, "source and target must be set before align");
    // pcl::Registration::align: converged_ = false, final_transformation_ = I before computeTransformation
    h.converged = false;
    for (int i = 0; i < 16; ++i) h.final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
    if (degenerate_pair(h.cfg, h.source, h.target))  // a single align reports it as an error (converged_ stays false)
      throw Error(B2R_ERR_INVALID_ARG, h.source->n == 0 || h.target->n == 0 ? "empty cloud" : "cloud has fewer points than correspondence_randomness");
    std::vector<Cloud*> s{h.source}, t{h.target};
    run_align(h, s, t, guess, 0, 0.0, out, nullptr);
    memcpy(h.final_T, out->T, sizeof(h.final_T));
    h.converged = out->converged != 0;
    h.has_result = true;
  });
  if (st != B2R_OK && out && guess) {  // PCL style: failure leaves converged_ = false; report the guess
    memcpy(out->T, guess, 64);
    out->converged = 0; out->iterations = 0; out->error = 0; out->evals = 0; out->fitness = 0;
  }
  return st;
}

b2r_status b2r_align_batch(b2r_handle* hh, b2r_cloud* const* sources, b2r_cloud* const* targets, const float* guesses, size_t n_pairs,
                           int with_fitness, double fitness_max_range, b2r_result* out) {
  return guarded(hh, [&](Handle& h) {
    if (n_pairs == 0) return;
    if (!sources || !targets || !guesses || !out) throw Error(B2R_ERR_INVALID_ARG, "null argument");
    std::vector<Cloud*> s(n_pairs), t(n_pairs);
    for (size_t i = 0; i < n_pairs; ++i) {
      if (!sources[i] || !targets[i]) throw Error(B2R_ERR_INVALID_ARG, "null cloud in batch");
      s[i] = &sources[i]->c;
      t[i] = &targets[i]->c;
    }
    run_align(h, s, t, guesses, with_fitness, fitness_max_range, out, nullptr);
  });
}

// getFitnessScore / calc_fitness_score / inlier fraction of one (source, target, T) triple
static void fitness_single(Handle& h, Cloud* source, Cloud* target, const float* T, double max_range, double* fitness_out, float inlier_d2,
                           double* inlier_fraction_out) {
  if (source->n == 0 || target->n == 0) {  // PCL: no correspondences -> max(); 0 / 0 inliers is reported as 0
    if (fitness_out) *fitness_out = DBL_MAX;
    if (inlier_fraction_out) *inlier_fraction_out = 0.0;
    return;
  }
  Ctx& ctx = h.ctx;
  std::vector<Cloud*> cl{source, target};
  std::vector<Needs> nd(2);
  nd[1].grid = true;
  DBuf<CloudView> dv;
  clouds_prepare(ctx, h.cfg, cl, nd, dv);
  struct Block { PairDesc pair; int n; int pad; b2r_result row; } hb;
  memset(&hb, 0, sizeof(hb));
  hb.pair = PairDesc{0, 1};
  hb.n = source->n;
  memcpy(hb.row.T, T, 64);
  DBuf<Block> db; db.alloc(1, ctx.stream);
  DBuf<int> din; din.alloc(1, ctx.stream);
  B2R_CUDA(cudaMemcpyAsync(db.p, &hb, sizeof(hb), cudaMemcpyHostToDevice, ctx.stream));
  BatchArgs b;
  memset(&b, 0, sizeof(b));
  b.d_views = dv.p;
  b.d_pairs = &db.p->pair;
  b.d_src_n = &db.p->n;
  b.d_rows = &db.p->row;
  b.np = 1;
  b.maxn = source->n;
  b.src_sizes = &hb.n;
  fitness_batch(ctx, b, max_range, inlier_d2, inlier_fraction_out ? din.p : nullptr);
  int inl = 0;
  B2R_CUDA(cudaMemcpyAsync(&hb.row, &db.p->row, sizeof(b2r_result), cudaMemcpyDeviceToHost, ctx.stream));
  if (inlier_fraction_out) B2R_CUDA(cudaMemcpyAsync(&inl, din.p, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  if (fitness_out) *fitness_out = hb.row.fitness;
  // static_cast<float>(num_inliers) / aligned->size()  (scan_matching_odometry_component.cpp:415)
  if (inlier_fraction_out) *inlier_fraction_out = (double)((float)inl / (float)source->n);
}

b2r_status b2r_fitness(b2r_handle* hh, double max_range, double* out) {
  return guarded(hh, [&](Handle& h) {
    if (!out) throw Error(B2R_ERR_INVALID_ARG, "null argument");
    if (!h.source || !h.target) throw Error(B2R_ERR_STATE, "source and target must be set first");
    fitness_single(h, h.source, h.target, h.final_T, max_range, out, 0.f, nullptr);
  });
}

b2r_status b2r_fitness_pair(b2r_handle* hh, b2r_cloud* target, b2r_cloud* source, const float T[16], double max_range, double* out) {
  return guarded(hh, [&](Handle& h) {
    if (!out || !target || !source || !T) throw Error(B2R_ERR_INVALID_ARG, "null argument");
    fitness_single(h, &source->c, &target->c, T, max_range, out, 0.f, nullptr);
  });
}

b2r_status b2r_inlier_fraction(b2r_handle* hh, double max_correspondence_dist, double* fraction_out, double* fitness_out) {
  return guarded(hh, [&](Handle& h) {
    if (!fraction_out) throw Error(B2R_ERR_INVALID_ARG, "null argument");
    if (!(max_correspondence_dist > 0)) throw Error(B2R_ERR_INVALID_ARG, "max_correspondence_dist must be > 0");
    if (!h.source || !h.target) throw Error(B2R_ERR_STATE, "source and target must be set first");
    // k_sq_dists[0] < max_correspondence_dist * max_correspondence_dist: a float against a double product (:413)
    const double thr = max_correspondence_dist * max_correspondence_dist;
    float thr_f = (float)thr;                    // smallest float t with (d2 < t) == ((double)d2 < thr) for every float d2:
    if ((double)thr_f < thr) thr_f = std::nextafterf(thr_f, INFINITY);  // round the double threshold UP to a float
    fitness_single(h, h.source, h.target, h.final_T, DBL_MAX, fitness_out, thr_f, fraction_out);
  });
}

b2r_status b2r_nearest_neighbors(b2r_handle* hh, int32_t* idx_out, float* d2_out, float* xyz_out) {
  return guarded(hh, [&](Handle& h) {
    if (!idx_out || !d2_out) throw Error(B2R_ERR_INVALID_ARG, "null argument");
    if (!h.source || !h.target) throw Error(B2R_ERR_STATE, "source and target must be set first");
    const int n = h.source->n;
    if (n == 0) return;
    if (h.target->n == 0) {
      for (int i = 0; i < n; ++i) { idx_out[i] = -1; d2_out[i] = INFINITY; }
      return;
    }
    Ctx& ctx = h.ctx;
    std::vector<Cloud*> cl{h.source, h.target};
    std::vector<Needs> nd(2);
    nd[1].grid = true;
    DBuf<CloudView> dv;
    clouds_prepare(ctx, h.cfg, cl, nd, dv);
    DBuf<int32_t> di; di.alloc(n, ctx.stream);
    DBuf<float> dd; dd.alloc(n, ctx.stream);
    DBuf<float> dx;
    if (xyz_out) dx.alloc((size_t)n * 3, ctx.stream);
    nearest_neighbors(ctx, dv.p, n, h.final_T, di.p, dd.p, xyz_out ? dx.p : nullptr);
    B2R_CUDA(cudaMemcpyAsync(idx_out, di.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx.stream));
    B2R_CUDA(cudaMemcpyAsync(d2_out, dd.p, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx.stream));
    if (xyz_out) B2R_CUDA(cudaMemcpyAsync(xyz_out, dx.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  });
}

b2r_status b2r_transform_source(b2r_handle* hh, void* out_points, size_t stride_bytes, int memspace) {
  return guarded(hh, [&](Handle& h) {
    if (!h.source) throw Error(B2R_ERR_STATE, "source must be set first");
    if (!out_points) throw Error(B2R_ERR_INVALID_ARG, "null output");
    DBuf<float4> tmp;
    tmp.alloc(h.source->n, h.ctx.stream);
    transform_cloud(h.ctx, h.source->pts.p, h.source->n, h.final_T, tmp.p);
    store_points(h.ctx, tmp.p, h.source->n, out_points, stride_bytes, memspace);
    B2R_CUDA(cudaStreamSynchronize(h.ctx.stream));
  });
}

// ------------------------------------------------------------------------------------------------ filters
static void finish_filter(Handle& h, DevCloud& res, void* out, size_t* m, int memspace) {
  if (out && res.n) store_points(h.ctx, res.pts.p, res.n, out, 16, memspace);
  B2R_CUDA(cudaStreamSynchronize(h.ctx.stream));
  if (m) *m = (size_t)res.n;
}

b2r_status b2r_distance_filter(b2r_handle* hh, const void* in, size_t n, size_t stride_bytes, int memspace, double near_thresh,
                               double far_thresh, void* out, size_t* m) {
  return guarded(hh, [&](Handle& h) {
    DBuf<float4> src;
    load_points(h.ctx, in, n, stride_bytes, memspace, src);
    DevCloud res;
    filter_distance(h.ctx, src.p, (int)n, near_thresh, far_thresh, res);
    finish_filter(h, res, out, m, memspace);
  });
}

b2r_status b2r_voxelgrid(b2r_handle* hh, const void* in, size_t n, size_t stride_bytes, int memspace, float leaf, int min_points_per_voxel,
                         void* out, size_t* m, int* overflow) {
  return guarded(hh, [&](Handle& h) {
    if (!(leaf > 0)) throw Error(B2R_ERR_INVALID_ARG, "leaf must be > 0");
    DBuf<float4> src;
    load_points(h.ctx, in, n, stride_bytes, memspace, src);
    DevCloud res;
    bool ovf = false;
    filter_voxelgrid(h.ctx, src.p, (int)n, leaf, min_points_per_voxel, res, ovf);
    if (overflow) *overflow = ovf ? 1 : 0;
    finish_filter(h, res, out, m, memspace);
  });
}

b2r_status b2r_radius_outlier(b2r_handle* hh, const void* in, size_t n, size_t stride_bytes, int memspace, double radius, int min_neighbors,
                              void* out, size_t* m) {
  return guarded(hh, [&](Handle& h) {
    if (!(radius > 0)) throw Error(B2R_ERR_INVALID_ARG, "radius must be > 0");
    DBuf<float4> src;
    load_points(h.ctx, in, n, stride_bytes, memspace, src);
    DevCloud res;
    filter_radius(h.ctx, h.cfg, src.p, (int)n, radius, min_neighbors, res);
    finish_filter(h, res, out, m, memspace);
  });
}

b2r_status b2r_statistical_outlier(b2r_handle* hh, const void* in, size_t n, size_t stride_bytes, int memspace, int mean_k, double stddev_mul,
                                   void* out, size_t* m) {
  return guarded(hh, [&](Handle& h) {
    DBuf<float4> src;
    load_points(h.ctx, in, n, stride_bytes, memspace, src);
    DevCloud res;
    filter_statistical(h.ctx, h.cfg, src.p, (int)n, mean_k, stddev_mul, res);
    finish_filter(h, res, out, m, memspace);
  });
}

b2r_status b2r_default_prefilter_config(b2r_prefilter_config* c) {
  if (!c) return B2R_ERR_INVALID_ARG;
  memset(c, 0, sizeof(*c));
  c->enable_distance_filter = 1;  // config/mrg_slam.yaml:48-64
  c->distance_near_thresh = 0.1;
  c->distance_far_thresh = 35.0;
  c->downsample_method = 1;
  c->downsample_resolution = 0.1f;
  c->downsample_min_points_per_voxel = 1;
  c->outlier_removal_method = 2;
  c->statistical_mean_k = 30;
  c->statistical_stddev = 1.2;
  c->radius_radius = 0.5;
  c->radius_min_neighbors = 2;
  return B2R_OK;
}

b2r_status b2r_prefilter(b2r_handle* hh, const b2r_prefilter_config* cfg, const void* in, size_t n, size_t stride_bytes, int memspace,
                         void* out, size_t* m) {
  return guarded(hh, [&](Handle& h) {
    if (!cfg) throw Error(B2R_ERR_INVALID_ARG, "null config");
    DevCloud cur;
    HostTrace tr;  // B2R_TRACE=1: stage times with a synchronisation after each stage (attribution only)
    auto mark = [&](const char* name) { if (tr.on) { cudaStreamSynchronize(h.ctx.stream); tr.mark(name); } };
    load_points(h.ctx, in, n, stride_bytes, memspace, cur.pts);
    cur.n = (int)n;
    mark("load");
    // cloud_callback order (apps/prefiltering_component.cpp:149-151): distance_filter, downsample, outlier_removal.
    // When VoxelGrid follows the distance filter, the filter is folded into VoxelGrid's own passes (no compaction,
    // no count read-back in between); the result is identical.
    const double range[2] = {cfg->distance_near_thresh, cfg->distance_far_thresh};
    const bool fold = cfg->enable_distance_filter && cfg->downsample_method == 1;
    if (cfg->enable_distance_filter && !fold) {
      DevCloud nxt;
      filter_distance(h.ctx, cur.pts.p, cur.n, cfg->distance_near_thresh, cfg->distance_far_thresh, nxt);
      cur = std::move(nxt);
    }
    if (cfg->downsample_method == 1) {
      DevCloud nxt;
      bool ovf = false;
      filter_voxelgrid(h.ctx, cur.pts.p, cur.n, cfg->downsample_resolution, cfg->downsample_min_points_per_voxel, nxt, ovf,
                       fold ? range : nullptr);
      cur = std::move(nxt);
      mark("voxelgrid");
    }
    if (cfg->outlier_removal_method == 1) {
      DevCloud nxt;
      filter_statistical(h.ctx, h.cfg, cur.pts.p, cur.n, cfg->statistical_mean_k, cfg->statistical_stddev, nxt, &cur);
      cur = std::move(nxt);
    } else if (cfg->outlier_removal_method == 2) {
      DevCloud nxt;
      filter_radius(h.ctx, h.cfg, cur.pts.p, cur.n, cfg->radius_radius, cfg->radius_min_neighbors, nxt, &cur);
      cur = std::move(nxt);
    }
    mark("outlier");
    finish_filter(h, cur, out, m, memspace);
    mark("store");
    tr.print("prefilter");
  });
}

b2r_status b2r_remove_robot_points(b2r_handle* hh, const void* in, size_t n, size_t stride_bytes, int memspace, const float* others_xyz,
                                   size_t n_others, float radius_sqr, void* kept_out, size_t* n_kept, void* removed_out, size_t* n_removed) {
  return guarded(hh, [&](Handle& h) {
    if (n_others && !others_xyz) throw Error(B2R_ERR_INVALID_ARG, "null robot positions");
    DBuf<float4> src;
    load_points(h.ctx, in, n, stride_bytes, memspace, src);
    DevCloud kept, removed;
    filter_robot_points(h.ctx, src.p, (int)n, others_xyz, (int)n_others, radius_sqr, kept, removed_out ? &removed : nullptr);
    if (removed_out && removed.n) store_points(h.ctx, removed.pts.p, removed.n, removed_out, 16, memspace);
    if (n_removed) *n_removed = removed_out ? (size_t)removed.n : (size_t)((int)n - kept.n);
    finish_filter(h, kept, kept_out, n_kept, memspace);
  });
}

b2r_status b2r_map_cloud(b2r_handle* hh, const void* const* clouds, const size_t* n, const double* poses_colmajor, const uint8_t* first_keyframe,
                         size_t count, size_t stride_bytes, int memspace, float resolution, int min_points_per_voxel, float distance_far_thresh,
                         int skip_first_cloud, void* out, size_t* m, int* is_null) {
  return guarded(hh, [&](Handle& h) {
    if (count && (!clouds || !n || !poses_colmajor)) throw Error(B2R_ERR_INVALID_ARG, "null keyframe arrays");
    DevCloud res;
    bool null_result = false;
    map_cloud(h.ctx, clouds, n, poses_colmajor, first_keyframe, count, stride_bytes, memspace, resolution, min_points_per_voxel,
              distance_far_thresh, skip_first_cloud, res, null_result);
    if (is_null) *is_null = null_result ? 1 : 0;
    finish_filter(h, res, out, m, memspace);
  });
}

// ------------------------------------------------------------------------------------------------ introspection
// kernels launched by the host + the kernels executed inside the optimiser graphs (two per round, counted on the device)
uint64_t b2r_kernel_launches(const b2r_handle* hh) {
  if (!hh) return 0;
  const Ctx& c = hh->h.ctx;
  unsigned long long in_graphs = 0;  // kernels executed by the optimiser-loop graphs, counted on the device
  if (c.d_graph_rounds) {
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    if (cudaMemcpy(&in_graphs, c.d_graph_rounds, sizeof(in_graphs), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); in_graphs = 0; }
  }
  return c.launches - c.graph_launches + in_graphs;
}
uint64_t b2r_graph_launches(const b2r_handle* hh) { return hh ? hh->h.ctx.graph_launches : 0; }

b2r_status b2r_debug_knn_list_overflows(b2r_handle* hh, uint64_t* out) {
  if (!out) return B2R_ERR_INVALID_ARG;
  return guarded(hh, [&](Handle& h) { *out = debug_knn_list_overflows(h.ctx); });
}

b2r_status b2r_synchronize(b2r_handle* hh) {
  return guarded(hh, [&](Handle& h) { B2R_CUDA(cudaStreamSynchronize(h.ctx.stream)); });
}

b2r_status b2r_last_timings(const b2r_handle* hh, float ms_out[4]) {
  if (!hh || !ms_out) return B2R_ERR_INVALID_ARG;
  Handle& h = const_cast<Handle&>(hh->h);
  if (h.timings_pending) {  // the stage events of the last align call are read on demand (the call itself does not wait for them)
    cudaSetDevice(h.ctx.device);
    if (cudaEventSynchronize(h.ev[3]) == cudaSuccess) {
      cudaEventElapsedTime(&h.timings[0], h.ev[0], h.ev[1]);
      cudaEventElapsedTime(&h.timings[1], h.ev[1], h.ev[2]);
      cudaEventElapsedTime(&h.timings[2], h.ev[2], h.ev[3]);
      cudaEventElapsedTime(&h.timings[3], h.ev[0], h.ev[3]);
    }
    cudaGetLastError();
    h.timings_pending = false;
  }
  for (int i = 0; i < 4; ++i) ms_out[i] = h.timings[i];
  return B2R_OK;
}

b2r_status b2r_event_record(b2r_handle* hh, int slot) {
  return guarded(hh, [&](Handle& h) {
    if (slot < 0 || slot >= 8) throw Error(B2R_ERR_INVALID_ARG, "event slot must be in [0,8)");
    if (!h.user_ev[slot]) B2R_CUDA(cudaEventCreate(&h.user_ev[slot]));
    B2R_CUDA(cudaEventRecord(h.user_ev[slot], h.ctx.stream));
  });
}
b2r_status b2r_event_elapsed_ms(b2r_handle* hh, int a, int b, float* ms) {
  return guarded(hh, [&](Handle& h) {
    if (a < 0 || a >= 8 || b < 0 || b >= 8 || !ms || !h.user_ev[a] || !h.user_ev[b]) throw Error(B2R_ERR_INVALID_ARG, "bad event slots");
    B2R_CUDA(cudaEventSynchronize(h.user_ev[b]));
    B2R_CUDA(cudaEventElapsedTime(ms, h.user_ev[a], h.user_ev[b]));
  });
}
b2r_status b2r_profile_enable(b2r_handle* hh, int on) {
  return guarded(hh, [&](Handle& h) {
    h.ctx.prof_resolve();
    h.ctx.profile = on != 0;
    for (int i = 0; i < PROF_COUNT; ++i) { h.ctx.prof_ms[i] = 0; h.ctx.prof_n[i] = 0; h.ctx.prof_bytes[i] = 0; }
  });
}
b2r_status b2r_profile_read(b2r_handle* hh, int id, double* ms_sum, uint64_t* launches, double* bytes) {
  return guarded(hh, [&](Handle& h) {
    if (id < 0 || id >= PROF_COUNT) throw Error(B2R_ERR_INVALID_ARG, "unknown kernel id");
    h.ctx.prof_resolve();
    if (ms_sum) *ms_sum = h.ctx.prof_ms[id];
    if (launches) *launches = h.ctx.prof_n[id];
    if (bytes) *bytes = h.ctx.prof_bytes[id];
  });
}

b2r_status b2r_debug_covariances(b2r_handle* hh, int which, double* cov6_out, int32_t* knn_out) {
  return guarded(hh, [&](Handle& h) {
    Cloud* c = which == 0 ? h.source : h.target;
    if (!c) throw Error(B2R_ERR_STATE, "cloud not set");
    const int k = h.cfg.correspondence_randomness;
    if (knn_out) {
      debug_cov_knn(h.ctx, h.cfg, *c, k, knn_out);
    } else {
      std::vector<Cloud*> cl{c};
      std::vector<Needs> nd(1);
      nd[0].cov_k = k;
      nd[0].cov_mode = h.cfg.method == B2R_GICP_PCL ? 1 : 0;
      DBuf<CloudView> dvtmp;
      clouds_prepare(h.ctx, h.cfg, cl, nd, dvtmp);
    }
    if (cov6_out) {
      B2R_CUDA(cudaMemcpyAsync(cov6_out, c->cov.p, sizeof(double) * 6 * c->n, cudaMemcpyDeviceToHost, h.ctx.stream));
      B2R_CUDA(cudaStreamSynchronize(h.ctx.stream));
    }
  });
}

b2r_status b2r_debug_voxelmap(b2r_handle* hh, int32_t* coords_out, int32_t* npts_out, double* mean_out, double* cov6_out, size_t* V) {
  return guarded(hh, [&](Handle& h) {
    if (!h.target) throw Error(B2R_ERR_STATE, "target not set");
    Cloud* c = h.target;
    std::vector<Cloud*> cl{c};
    std::vector<Needs> nd(1);
    nd[0].cov_k = h.cfg.correspondence_randomness;
    nd[0].vres = h.cfg.resolution;
    DBuf<CloudView> dvtmp;
    clouds_prepare(h.ctx, h.cfg, cl, nd, dvtmp);
    int nrec = 0;
    B2R_CUDA(cudaMemcpyAsync(&nrec, c->v_nrec.p, sizeof(int), cudaMemcpyDeviceToHost, h.ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(h.ctx.stream));
    std::vector<VoxRec> recs(nrec);
    B2R_CUDA(cudaMemcpyAsync(recs.data(), c->vrec.p, sizeof(VoxRec) * nrec, cudaMemcpyDeviceToHost, h.ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(h.ctx.stream));
    struct Item { int x, y, z; const VoxRec* r; };
    std::vector<Item> items(nrec);
    for (int i = 0; i < nrec; ++i) {
      const int cell = recs[i].cell;
      const int x = cell % c->vd[0], y = (cell / c->vd[0]) % c->vd[1], z = cell / (c->vd[0] * c->vd[1]);
      items[i] = Item{x + c->vmin[0], y + c->vmin[1], z + c->vmin[2], &recs[i]};
    }
    std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.x != b.x ? a.x < b.x : (a.y != b.y ? a.y < b.y : a.z < b.z); });
    for (int i = 0; i < nrec; ++i) {
      coords_out[i * 3] = items[i].x; coords_out[i * 3 + 1] = items[i].y; coords_out[i * 3 + 2] = items[i].z;
      npts_out[i] = items[i].r->n;
      for (int d = 0; d < 3; ++d) mean_out[i * 3 + d] = items[i].r->mean[d];
      for (int d = 0; d < 6; ++d) cov6_out[i * 6 + d] = items[i].r->cov[d];
    }
    *V = (size_t)nrec;
  });
}

b2r_status b2r_debug_linearize(b2r_handle* hh, const double* T, double* H, double* b, double* err, int32_t* corr_out, uint8_t* corr_valid) {
  return guarded(hh, [&](Handle& h) {
    if (h.cfg.method == B2R_NDT_OMP) throw Error(B2R_ERR_STATE, "linearize is a GICP/VGICP entry point");
    DBuf<CloudView> dv;
    prepare_current(h, dv, false);
    lsq_debug_linearize(h.ctx, h.cfg, dv.p, h.source->n, T, T, false, H, b, err, corr_out, corr_valid);
  });
}
b2r_status b2r_debug_compute_error(b2r_handle* hh, const double* T_lin, const double* T_trial, double* err) {
  return guarded(hh, [&](Handle& h) {
    if (h.cfg.method == B2R_NDT_OMP) throw Error(B2R_ERR_STATE, "compute_error is a GICP/VGICP entry point");
    DBuf<CloudView> dv;
    prepare_current(h, dv, false);
    lsq_debug_linearize(h.ctx, h.cfg, dv.p, h.source->n, T_lin, T_trial, true, nullptr, nullptr, err, nullptr, nullptr);
  });
}

b2r_status b2r_debug_ndt_grid(b2r_handle* hh, int32_t* idx_out, int32_t* npts_out, double* mean_out, double* icov_out, int32_t* min_b,
                              int32_t* div_b, size_t* V) {
  return guarded(hh, [&](Handle& h) {
    if (!h.target) throw Error(B2R_ERR_STATE, "target not set");
    Cloud* c = h.target;
    std::vector<Cloud*> cl{c};
    std::vector<Needs> nd(1);
    nd[0].leaf = (float)h.cfg.resolution;
    DBuf<CloudView> dvtmp;
    clouds_prepare(h.ctx, h.cfg, cl, nd, dvtmp);
    for (int d = 0; d < 3; ++d) { min_b[d] = c->min_b[d]; div_b[d] = c->div_b[d]; }
    *V = 0;
    if (c->ncell_ndt == 0) return;
    int nrec = 0;
    B2R_CUDA(cudaMemcpyAsync(&nrec, c->n_nrec.p, sizeof(int), cudaMemcpyDeviceToHost, h.ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(h.ctx.stream));
    std::vector<NdtRec> recs(nrec);
    B2R_CUDA(cudaMemcpyAsync(recs.data(), c->nrec.p, sizeof(NdtRec) * nrec, cudaMemcpyDeviceToHost, h.ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(h.ctx.stream));
    std::sort(recs.begin(), recs.end(), [](const NdtRec& a, const NdtRec& b) { return a.cell < b.cell; });
    for (int i = 0; i < nrec; ++i) {
      idx_out[i] = recs[i].cell;
      npts_out[i] = recs[i].n;
      for (int d = 0; d < 3; ++d) mean_out[i * 3 + d] = recs[i].mean[d];
      for (int d = 0; d < 9; ++d) icov_out[i * 9 + d] = recs[i].icov_d[d];
    }
    *V = (size_t)nrec;
  });
}

b2r_status b2r_debug_ndt_derivatives(b2r_handle* hh, const double* p6, double* score, double* grad6, double* hess36, int32_t* hits_out) {
  return guarded(hh, [&](Handle& h) {
    if (h.cfg.method != B2R_NDT_OMP) throw Error(B2R_ERR_STATE, "ndt derivatives need an NDT_OMP handle");
    DBuf<CloudView> dv;
    prepare_current(h, dv, false);
    ndt_debug_derivatives(h.ctx, h.cfg, dv.p, h.source->n, p6, score, grad6, hess36, hits_out);
  });
}

b2r_status b2r_debug_knn(b2r_handle* hh, b2r_cloud* c, const float* queries, size_t nq, int k, int32_t* idx_out, float* d2_out) {
  return guarded(hh, [&](Handle& h) {
    if (!c || !queries || !idx_out || !d2_out) throw Error(B2R_ERR_INVALID_ARG, "null argument");
    debug_knn(h.ctx, h.cfg, c->c, queries, nq, k, idx_out, d2_out);
  });
}

}  // extern "C"
