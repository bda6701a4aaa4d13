// Multi-GPU loop-closure batches behind the C ABI: communicator, partition of the pair list by target, the sharded batch
// call and the best-candidate reduction.
//
// Replaces the candidate loop of LoopDetector::matching (/root/reference/src/mrg_slam/loop_detector.cpp:97-180) for MANY new
// keyframes at once, spread over the GPUs of one box: the (target, candidate) pairs are independent
// (:126-145 touches only per-iteration state and a running best), so they are partitioned by target — all candidates of one
// new keyframe on one rank, whose voxel map / covariances / NN grid are then built once — every rank runs its slice through
// the same batch path as b2r_align_batch, the result rows are written by the rank's last kernels straight into the send
// buffer of ONE ncclAllGather over NVLink, and every rank applies the reference's best-candidate rule to the same table.
// There is no collective inside the optimiser.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2; a copy already loaded by the host process — e.g. PyTorch's — is reused),
// so libb2r.so itself has no link-time dependency on it and single-GPU users never touch it.  A second transport takes a
// host all-gather callback (MPI, gloo, tests): same entry points, same table.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cfloat>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>

#include "internal.hpp"

using namespace b2r;

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string error;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {  // a copy the process already holds (PyTorch bundles one) wins: one NCCL per process
      api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if (api.lib) break;
    }
    for (int i = 0; i < 2 && !api.lib; ++i) api.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) { api.error = std::string("libnccl.so.2 not found: ") + dlerror(); return; }
    auto sym = [&](const char* s) { void* p = dlsym(api.lib, s); if (!p) api.error = std::string("missing NCCL symbol ") + s; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  });
  return api;
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw Error(B2R_ERR_COMM, std::string(what) + " failed: " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "?"));
}

}  // namespace

struct b2r_comm {
  int rank = 0, nranks = 1;
  // NCCL transport
  ncclComm_t nccl = nullptr;
  int device = 0;
  // host transport
  b2r_allgather_fn host_fn = nullptr;
  void* host_user = nullptr;
  std::string last_error;
  uint64_t collectives = 0;
};

namespace {

template <typename F>
b2r_status comm_guarded(b2r_comm* c, F&& f) {
  try {
    f();
    return B2R_OK;
  } catch (const Error& e) {
    if (c) c->last_error = e.what();
    cudaGetLastError();
    return e.status;
  } catch (const std::exception& e) {
    if (c) c->last_error = e.what();
    return B2R_ERR_INVALID_ARG;
  }
}

// Static block partition of the pair list by target id (SURVEY 8e "Partitioning"): targets keep their order of first appearance,
// every rank gets a contiguous block of targets whose summed weight is as close to total / nranks as whole targets allow (a
// target goes where most of it falls), never starving the later ranks.  Every rank computes the same answer.
void partition_pairs(const int64_t* target_ids, const double* weights, size_t n, int nranks, int32_t* rank_of_pair) {
  // targets in order of first appearance, each with its summed weight (one pass, no per-target lists)
  std::map<int64_t, int> slot;
  std::vector<double> gw;
  std::vector<int> slot_of(n);
  double total = 0.0;
  int last_slot = -1;
  int64_t last_id = 0;
  for (size_t i = 0; i < n; ++i) {
    const double w = weights ? weights[i] : 1.0;
    int k;
    if (last_slot >= 0 && target_ids[i] == last_id) {
      k = last_slot;  // the candidates of a target are usually consecutive
    } else {
      auto it = slot.find(target_ids[i]);
      if (it == slot.end()) { k = (int)gw.size(); slot[target_ids[i]] = k; gw.push_back(0.0); }
      else k = it->second;
      last_slot = k; last_id = target_ids[i];
    }
    gw[k] += w;
    slot_of[i] = k;
    total += w;
  }
  std::vector<int32_t> rank_of_slot(gw.size());
  int rank = 0;
  double acc = 0.0;
  for (size_t k = 0; k < gw.size(); ++k) {
    while (rank < nranks - 1 && acc + 0.5 * gw[k] >= (rank + 1) * total / nranks) ++rank;
    rank_of_slot[k] = rank;
    acc += gw[k];
  }
  for (size_t i = 0; i < n; ++i) rank_of_pair[i] = rank_of_slot[slot_of[i]];
}

// All-gathers `cnt_max` rows per rank.  send: this rank's rows (device for NCCL, host for the callback transport).  Returns the
// gathered table [rank][cnt_max]: the handle's pinned staging buffer when the table arrived there by DMA (valid until the
// handle's next call), `all` otherwise.
const b2r_result* gather_rows(Handle* h, b2r_comm* comm, const b2r_result* d_send, const b2r_result* h_send, size_t cnt_max,
                              std::vector<b2r_result>& all) {
  const size_t bytes = cnt_max * sizeof(b2r_result);
  const size_t rows = cnt_max * (size_t)comm->nranks;
  const b2r_result* table = nullptr;
  if (comm->nccl) {
    Ctx& ctx = h->ctx;
    DBuf<b2r_result> recv; recv.alloc(rows, ctx.stream);
    HostTrace tr;  // B2R_TRACE=1: device time of the collective itself (includes waiting for the slowest rank)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (tr.on) { e0 = ctx.get_event(); e1 = ctx.get_event(); cudaEventRecord(e0, ctx.stream); }
    nccl_check(nccl().AllGather(d_send, recv.p, bytes, ncclChar, comm->nccl, ctx.stream), "ncclAllGather");
    if (tr.on) cudaEventRecord(e1, ctx.stream);
    const size_t tot = rows * sizeof(b2r_result);
    void* stage = ctx.pinned_buf(tot);
    if (!stage) { all.resize(rows); stage = all.data(); }
    B2R_CUDA(cudaMemcpyAsync(stage, recv.p, tot, cudaMemcpyDeviceToHost, ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(ctx.stream));
    if (tr.on) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      fprintf(stderr, "[b2r trace] rank %d ncclAllGather (device, incl. waiting for the other ranks): %.3f ms, %zu B per rank\n", comm->rank, ms, bytes);
      ctx.ev_pool.push_back(e0); ctx.ev_pool.push_back(e1);
    }
    table = static_cast<const b2r_result*>(stage);
  } else {
    all.resize(rows);
    if (comm->host_fn(comm->host_user, h_send, all.data(), bytes) != 0) throw Error(B2R_ERR_COMM, "host all-gather callback failed");
    table = all.data();
  }
  ++comm->collectives;
  return table;
}

}  // namespace

extern "C" {

b2r_status b2r_comm_unique_id(void* id_out) {
  if (!id_out) return B2R_ERR_INVALID_ARG;
  NcclApi& api = nccl();
  if (!api.error.empty() || !api.GetUniqueId) return B2R_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == B2R_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  if (api.GetUniqueId(&id) != ncclSuccess) return B2R_ERR_COMM;
  memcpy(id_out, &id, sizeof(id));
  return B2R_OK;
}

b2r_status b2r_comm_init(b2r_handle* hh, const void* unique_id, int rank, int nranks, b2r_comm** out) {
  if (!out) return B2R_ERR_INVALID_ARG;
  *out = nullptr;
  if (!hh || !unique_id || nranks < 1 || rank < 0 || rank >= nranks) return B2R_ERR_INVALID_ARG;
  Handle& h = *b2r_handle_impl(hh);
  b2r_comm* c = new b2r_comm();
  b2r_status st = comm_guarded(c, [&] {
    NcclApi& api = nccl();
    if (!api.error.empty()) throw Error(B2R_ERR_COMM, api.error);
    B2R_CUDA(cudaSetDevice(h.ctx.device));
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    c->rank = rank;
    c->nranks = nranks;
    c->device = h.ctx.device;
    nccl_check(api.CommInitRank(&c->nccl, nranks, id, rank), "ncclCommInitRank");
  });
  if (st != B2R_OK) { h.last_error = c->last_error; delete c; return st; }
  *out = c;
  return B2R_OK;
}

b2r_status b2r_comm_init_host(b2r_allgather_fn fn, void* user, int rank, int nranks, b2r_comm** out) {
  if (!out) return B2R_ERR_INVALID_ARG;
  *out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks || (!fn && nranks > 1)) return B2R_ERR_INVALID_ARG;
  b2r_comm* c = new b2r_comm();
  c->rank = rank;
  c->nranks = nranks;
  c->host_fn = fn;
  c->host_user = user;
  *out = c;
  return B2R_OK;
}

void b2r_comm_destroy(b2r_comm* c) {
  if (!c) return;
  if (c->nccl) {
    cudaSetDevice(c->device);
    nccl().CommDestroy(c->nccl);
  }
  delete c;
}

int b2r_comm_rank(const b2r_comm* c) { return c ? c->rank : -1; }
int b2r_comm_size(const b2r_comm* c) { return c ? c->nranks : 0; }
const char* b2r_comm_last_error(const b2r_comm* c) { return c ? c->last_error.c_str() : "null communicator"; }
uint64_t b2r_comm_collectives(const b2r_comm* c) { return c ? c->collectives : 0; }

b2r_status b2r_partition_by_target(const int64_t* target_ids, const double* weights, size_t n_pairs, int nranks, int32_t* rank_of_pair) {
  if (nranks < 1 || (n_pairs && (!target_ids || !rank_of_pair))) return B2R_ERR_INVALID_ARG;
  try {
    partition_pairs(target_ids, weights, n_pairs, nranks, rank_of_pair);
  } catch (const std::exception&) {
    return B2R_ERR_INVALID_ARG;
  }
  return B2R_OK;
}

b2r_status b2r_gather_results(b2r_handle* hh, b2r_comm* comm, const int32_t* rank_of_pair, size_t n_pairs, const b2r_result* local, b2r_result* out) {
  if (!comm || (n_pairs && (!rank_of_pair || !out))) return B2R_ERR_INVALID_ARG;
  if (comm->nccl && !hh) return B2R_ERR_INVALID_ARG;
  return comm_guarded(comm, [&] {
    std::vector<size_t> cnt(comm->nranks, 0);
    for (size_t i = 0; i < n_pairs; ++i) {
      if (rank_of_pair[i] < 0 || rank_of_pair[i] >= comm->nranks) throw Error(B2R_ERR_INVALID_ARG, "rank_of_pair out of range");
      ++cnt[rank_of_pair[i]];
    }
    const size_t cnt_max = std::max<size_t>(1, *std::max_element(cnt.begin(), cnt.end()));
    if (cnt[comm->rank] && !local) throw Error(B2R_ERR_INVALID_ARG, "null local rows");
    std::vector<b2r_result> send(cnt_max), all;
    memset(send.data(), 0, sizeof(b2r_result) * cnt_max);
    if (cnt[comm->rank]) memcpy(send.data(), local, sizeof(b2r_result) * cnt[comm->rank]);
    const b2r_result* table = send.data();
    if (comm->nranks == 1 && !comm->nccl) {
    } else if (comm->nccl) {
      Handle& h = *b2r_handle_impl(hh);
      B2R_CUDA(cudaSetDevice(h.ctx.device));
      DBuf<b2r_result> dsend; dsend.alloc(cnt_max, h.ctx.stream);
      B2R_CUDA(cudaMemcpyAsync(dsend.p, send.data(), sizeof(b2r_result) * cnt_max, cudaMemcpyHostToDevice, h.ctx.stream));
      table = gather_rows(&h, comm, dsend.p, nullptr, cnt_max, all);
    } else {
      table = gather_rows(nullptr, comm, nullptr, send.data(), cnt_max, all);
    }
    std::vector<size_t> cur(comm->nranks, 0);
    for (size_t i = 0; i < n_pairs; ++i) {
      const int r = rank_of_pair[i];
      out[i] = table[(size_t)r * cnt_max + cur[r]++];
    }
  });
}

b2r_status b2r_align_batch_sharded(b2r_handle* hh, b2r_comm* comm, b2r_cloud* const* sources, b2r_cloud* const* targets, const int64_t* target_ids,
                                   const double* weights, const float* guesses, size_t n_pairs, int with_fitness, double fitness_max_range,
                                   b2r_result* out) {
  if (!hh || !comm) return B2R_ERR_INVALID_ARG;
  if (n_pairs == 0) return B2R_OK;
  if (!sources || !targets || !target_ids || !guesses || !out) return B2R_ERR_INVALID_ARG;
  Handle& h = *b2r_handle_impl(hh);
  return comm_guarded(comm, [&] {
    B2R_CUDA(cudaSetDevice(h.ctx.device));
    Ctx& ctx = h.ctx;
    HostTrace tr;
    // ---- the same partition on every rank
    std::vector<int32_t> rank_of(n_pairs);
    partition_pairs(target_ids, weights, n_pairs, comm->nranks, rank_of.data());
    std::vector<size_t> cnt(comm->nranks, 0);
    for (size_t i = 0; i < n_pairs; ++i) ++cnt[rank_of[i]];
    const size_t cnt_max = std::max<size_t>(1, *std::max_element(cnt.begin(), cnt.end()));
    const size_t mine = cnt[comm->rank];
    // ---- this rank's slice, in pair order
    std::vector<Cloud*> s, t;
    std::vector<float> g(mine * 16);
    s.reserve(mine); t.reserve(mine);
    std::string slice_error;
    for (size_t i = 0, j = 0; i < n_pairs; ++i) {
      if (rank_of[i] != comm->rank) continue;
      if (!sources[i] || !targets[i]) slice_error = "a pair of this rank's slice has a null cloud";
      s.push_back(sources[i] ? b2r_cloud_impl(sources[i]) : nullptr);
      t.push_back(targets[i] ? b2r_cloud_impl(targets[i]) : nullptr);
      memcpy(&g[j * 16], guesses + i * 16, 64);
      ++j;
    }
    tr.mark("partition+slice");
    // ---- align: the result rows are written on the device straight into the all-gather send buffer.  A failure on this rank must
    // not keep it out of the collective (the others would wait forever): its rows become failure rows and the error is reported
    // after the gather.
    DBuf<b2r_result> send; send.alloc(cnt_max, ctx.stream);
    std::vector<b2r_result> host_rows;  // host transport only
    const bool host_transport = comm->nccl == nullptr;
    b2r_status local_status = B2R_OK;
    std::string local_error;
    try {
      if (!slice_error.empty()) throw Error(B2R_ERR_INVALID_ARG, slice_error);
      if (mine) {
        if (host_transport) {
          host_rows.resize(cnt_max);
          memset(host_rows.data(), 0, sizeof(b2r_result) * cnt_max);
          b2r_run_align(h, s, t, g.data(), with_fitness, fitness_max_range, host_rows.data(), nullptr);
        } else {
          b2r_run_align(h, s, t, g.data(), with_fitness, fitness_max_range, nullptr, send.p);
        }
      } else if (host_transport) {
        host_rows.assign(cnt_max, b2r_result());
      }
    } catch (const Error& e) {
      local_status = e.status;
      local_error = e.what();
      cudaGetLastError();
      std::vector<b2r_result> fail(cnt_max);
      memset(fail.data(), 0, sizeof(b2r_result) * cnt_max);
      for (size_t j = 0; j < mine; ++j) {
        memcpy(fail[j].T, &g[j * 16], 64);
        fail[j].fitness = DBL_MAX;
      }
      if (host_transport) host_rows = fail;
      else {
        B2R_CUDA(cudaMemcpyAsync(send.p, fail.data(), sizeof(b2r_result) * cnt_max, cudaMemcpyHostToDevice, ctx.stream));
        B2R_CUDA(cudaStreamSynchronize(ctx.stream));
      }
    }
    tr.mark("run_align");
    // ---- one all-gather of fixed-size rows; every rank ends up with the whole table in pair order
    std::vector<b2r_result> all;
    const b2r_result* table = host_rows.data();
    if (!(comm->nranks == 1 && host_transport)) table = gather_rows(&h, comm, send.p, host_rows.data(), cnt_max, all);
    tr.mark("allgather+d2h+sync");
    std::vector<size_t> cur(comm->nranks, 0);
    for (size_t i = 0; i < n_pairs; ++i) {
      const int r = rank_of[i];
      out[i] = table[(size_t)r * cnt_max + cur[r]++];
    }
    tr.mark("scatter");
    tr.print("align_batch_sharded");
    if (local_status != B2R_OK) throw Error(local_status, local_error);
  });
}

b2r_status b2r_select_best_candidates(const b2r_result* results, const int64_t* target_ids, size_t n_pairs, double fitness_score_thresh,
                                      int64_t* best_pair_out, double* best_score_out, size_t* n_targets_out) {
  if (n_pairs && (!results || !target_ids)) return B2R_ERR_INVALID_ARG;
  if (!best_pair_out || !n_targets_out) return B2R_ERR_INVALID_ARG;
  // loop_detector.cpp:106-145 per target, candidates in list order: `!hasConverged() || score > best_score -> continue`, so among
  // equal scores the LATER candidate wins; :156-160: best_score > fitness_score_thresh -> no loop (-1)
  std::vector<int64_t> order;
  std::map<int64_t, size_t> slot;
  std::vector<double> best;
  std::vector<int64_t> arg;
  for (size_t i = 0; i < n_pairs; ++i) {
    auto it = slot.find(target_ids[i]);
    size_t k;
    if (it == slot.end()) {
      k = order.size();
      slot[target_ids[i]] = k;
      order.push_back(target_ids[i]);
      best.push_back(DBL_MAX);
      arg.push_back(-1);
    } else {
      k = it->second;
    }
    const b2r_result& r = results[i];
    if (!r.converged || r.fitness > best[k]) continue;
    best[k] = r.fitness;
    arg[k] = (int64_t)i;
  }
  for (size_t k = 0; k < order.size(); ++k) {
    best_pair_out[k] = best[k] > fitness_score_thresh ? -1 : arg[k];
    if (best_score_out) best_score_out[k] = best[k];
  }
  *n_targets_out = order.size();
  return B2R_OK;
}

}  // extern "C"
