// libb2r internal definitions shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b2r.h"

namespace b2r {

struct Error : std::runtime_error {
  b2r_status status;
  Error(b2r_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};

#define B2R_CUDA(expr)                                                                                       \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess) {                                                                                 \
      char _b[512];                                                                                          \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      throw ::b2r::Error(B2R_ERR_CUDA, _b);                                                                  \
    }                                                                                                        \
  } while (0)

// Private stream-ordered memory pool of the current device (created on first use, release threshold = keep everything): the
// process's default pool and its attributes are left alone.
cudaMemPool_t device_pool();
inline cudaError_t pool_malloc(void** p, size_t bytes, cudaStream_t s) { return cudaMallocFromPoolAsync(p, bytes, device_pool(), s); }

enum { PROF_KNN_COV = 0, PROF_LSQ_EVAL, PROF_NDT_EVAL, PROF_GRID_BUILD, PROF_VOXEL_REDUCE, PROF_FITNESS, PROF_COUNT };

struct ProfRec {
  int id;
  cudaEvent_t a, b;
  double bytes;
};

// The handle's stream, shared with every allocation made on it: a cloud may outlive the handle that created it
// (b2r_cloud_destroy after b2r_destroy), and its stream-ordered frees still need a live stream.
struct StreamOwner {
  cudaStream_t s = nullptr;
  int device = 0;
  ~StreamOwner() {
    if (s) { cudaSetDevice(device); cudaStreamDestroy(s); }
  }
};

// Bounding boxes of a group of clouds whose upload is still running on the handle's copy stream (clouds_upload, split upload of
// pinned host clouds: the second half of a batch is pulled over PCIe while the first half's search structures are being built).
// Shared by the clouds of the group; whoever needs a box first waits for the event and reads them all.
struct PendingBoxes {
  int device = 0;
  cudaEvent_t ev = nullptr;     // after the group's copy + box kernel and the read-back of the boxes
  const int* pinned = nullptr;  // where the boxes arrive (the Ctx's pinned area; valid until resolved)
  int count = 0;                // clouds in the group
  std::vector<int> vals;        // 6 ordered ints per cloud, once resolved
  bool resolved = false;
  std::vector<std::pair<void*, cudaStream_t>> scratch;  // device scratch of the group's kernel, released after the event
  void resolve() {
    if (resolved) return;
    cudaSetDevice(device);
    if (ev) cudaEventSynchronize(ev);
    if (pinned) vals.assign(pinned, pinned + (size_t)6 * count);
    for (auto& sc : scratch) cudaFreeAsync(sc.first, sc.second);
    scratch.clear();
    resolved = true;
  }
  ~PendingBoxes() {
    resolve();  // the copy must have finished before the clouds' storage can be released
    if (ev) cudaEventDestroy(ev);
  }
};

// Host-side stage timer (B2R_TRACE=1): wall-clock milliseconds between marks, printed to stderr by the caller.
struct HostTrace {
  bool on;
  std::vector<std::pair<const char*, double>> marks;
  double t0;
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
  HostTrace() {
    static const bool enabled = [] { const char* e = getenv("B2R_TRACE"); return e && atoi(e) != 0; }();
    on = enabled;
    t0 = on ? now() : 0.0;
  }
  void mark(const char* name) {
    if (!on) return;
    const double t = now();
    marks.emplace_back(name, t - t0);
    t0 = t;
  }
  void print(const char* what) const {
    if (!on) return;
    fprintf(stderr, "[b2r trace] %s:", what);
    for (const auto& m : marks) fprintf(stderr, " %s %.3f", m.first, m.second);
    fprintf(stderr, "\n");
  }
};

// Execution context of one handle: device, stream, launch counter, optional per-kernel event timing.
struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::shared_ptr<StreamOwner> stream_owner;
  uint64_t launches = 0;        // kernel launches + graph launches issued by this handle
  uint64_t graph_launches = 0;  // of which: optimiser loops launched as one CUDA graph
  unsigned long long* d_graph_rounds = nullptr;  // device counter: kernels executed inside those graphs
  // One instantiated loop graph per (eval kernel, step kernel) pair of this handle: later alignments only rewrite the two kernel
  // nodes' parameters (cudaGraphExecKernelNodeSetParams) instead of building and instantiating a new graph (~0.26 ms each).
  struct LoopGraph {
    const void* eval_fn;
    const void* eval2_fn;
    const void* step_fn;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    cudaGraphNode_t eval_node, eval2_node, step_node;
    cudaGraphConditionalHandle handle;
  };
  std::vector<LoopGraph> loop_graphs;
  bool loop_graph_cache_ok = true;  // cleared if the driver refuses to update nodes inside a conditional body
  // executable graphs of loops that may still be running: destroying one in flight makes the host wait for it, so they are
  // released later, once the stream has drained (reap_graphs)
  std::vector<std::pair<cudaGraphExec_t, cudaGraph_t>> graph_graveyard;
  void reap_graphs(bool force) {
    if (graph_graveyard.empty()) return;
    if (!force && graph_graveyard.size() < 64 && cudaStreamQuery(stream) != cudaSuccess) { cudaGetLastError(); return; }
    if (force || graph_graveyard.size() >= 64) cudaStreamSynchronize(stream);
    for (auto& g : graph_graveyard) { cudaGraphExecDestroy(g.first); cudaGraphDestroy(g.second); }
    graph_graveyard.clear();
    if (force) {
      for (auto& g : loop_graphs) { cudaGraphExecDestroy(g.exec); cudaGraphDestroy(g.graph); }
      loop_graphs.clear();
    }
  }
  int num_sms = 148;
  // split upload (cloud.cu): a second stream for the copy of a batch's second half, a pinned area its boxes arrive in, and the
  // group that may still be in flight (resolved before the area is reused and when the handle goes away)
  cudaStream_t copy_stream = nullptr;
  int* pinned_boxes = nullptr;
  size_t pinned_boxes_cap = 0;
  std::weak_ptr<PendingBoxes> pending_boxes;
  // pinned host staging for small device->host results (result tables, counters): a copy into pageable memory goes through the
  // driver's own bounce buffer and costs several times the transfer
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  void* pinned_buf(size_t bytes) {
    if (bytes > pinned_bytes) {
      if (pinned) cudaFreeHost(pinned);
      pinned = nullptr;
      pinned_bytes = 0;
      const size_t want = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
      if (cudaHostAlloc(&pinned, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); pinned = nullptr; return nullptr; }
      pinned_bytes = want;
    }
    return pinned;
  }
  bool profile = false;
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[PROF_COUNT] = {0};
  uint64_t prof_n[PROF_COUNT] = {0};
  double prof_bytes[PROF_COUNT] = {0};

  cudaEvent_t get_event() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  void prof_resolve() {
    if (prof_pending.empty()) return;
    cudaStreamSynchronize(stream);
    for (ProfRec& r : prof_pending) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { prof_ms[r.id] += ms; prof_n[r.id]++; prof_bytes[r.id] += r.bytes; }
      ev_pool.push_back(r.a);
      ev_pool.push_back(r.b);
    }
    prof_pending.clear();
  }
};

// Brackets the launches issued during its lifetime with CUDA events on the handle's stream (only when profiling is on).
struct ProfScope {
  Ctx& c;
  int id;
  double bytes;
  cudaEvent_t a = nullptr;
  ProfScope(Ctx& ctx, int kernel_id, double algorithmic_bytes, bool enabled = true) : c(ctx), id(kernel_id), bytes(algorithmic_bytes) {
    if (c.profile && enabled) { a = c.get_event(); cudaEventRecord(a, c.stream); }
  }
  ~ProfScope() {
    if (a) {
      cudaEvent_t b = c.get_event();
      cudaEventRecord(b, c.stream);
      c.prof_pending.push_back(ProfRec{id, a, b, bytes});
      if (c.prof_pending.size() > 4096) c.prof_resolve();
    }
  }
};

#define B2R_LAUNCH(ctx, kernel, grid, block, smem, ...)            \
  do {                                                             \
    kernel<<<(grid), (block), (smem), (ctx).stream>>>(__VA_ARGS__); \
    ++(ctx).launches;                                              \
    B2R_CUDA(cudaGetLastError());                                  \
  } while (0)

// Stream-ordered device buffer.
template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t count, cudaStream_t stream) {
    if (count <= n && p) { s = stream; return; }
    release();
    s = stream;
    n = count;
    if (count) B2R_CUDA(pool_malloc((void**)&p, count * sizeof(T), stream));
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
  void zero(cudaStream_t stream) { if (p) B2R_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), stream)); }
};

// ------------------------------------------------------------------------------------------------
// Device-side view of a cloud and of its cached structures.  Kernels that work on many clouds take an
// array of these and pick views[blockIdx.y].
struct VoxRec {      // VGICP Gaussian voxel (fast_gicp GaussianVoxelMap); 96 B, the first 80 are what an evaluation reads,
  double cov[6];     // as five 16-byte loads in the order it uses them: covariance (xx,xy,xz,yy,yz,zz) for the Mahalanobis matrix,
  double mean[3];    // then mean and weight for the residual.
  double w;          // w = sqrt(n), the weight of the voxel's residual
  int n;             // (Measured and rejected on the 4096-pair batch, where the evaluation kernels run at 82-95 % of the L1 data
  int cell;          // pipe's wavefront rate: one 128-byte line per record, 21.8 -> 23.8 ms per step — the larger footprint costs more
  double pad;        // L1 / L2 hits than the straddling records cost wavefronts.)  cell = dense table index (export)
};
static_assert(sizeof(VoxRec) == 96, "VoxRec layout");
struct NdtRec {      // NDT leaf (pclomp::VoxelGridCovariance)
  double mean[3];
  float icov[9];     // inverse covariance rounded to float (updateDerivatives casts it on use)
  int n;             // nr_points, -1 if unusable
  int cell;
  double icov_d[9];  // full-precision copy for export
  float centroid[3]; // leaf.centroid of VoxelGridCovariance: float sum in point order / float count (KDTREE search only)
  float pad;
};

struct CloudView {
  const float4* pts;  // original order: x,y,z,intensity
  int n;
  float bmin[3], bmax[3];
  // exact-NN grid (dense cell table; cells of size hx along x and h along y, z; points copied in cell order and, inside a
  // cell, in ascending x: every row of cells along x is one run of the array sorted by x; w = original index bits)
  float h, inv_h;
  float hx, inv_hx;
  int gd[3];
  int ncell;
  int* cell_start;  // ncell + 1
  int* cell_cnt;    // ncell (count, then scatter cursor)
  int2* cell_tmp;   // n: (ordered x bits, point index) grouped by cell in scatter (arbitrary) order, before the in-cell ranking
  float4* spts;
  // the same points, two per record, component-wise: record m = {x[2m] x[2m+1] y[2m] y[2m+1]} {z[2m] z[2m+1] w[2m] w[2m+1]}, so that
  // ONE packed-pair FP32 instruction (FFMA2 / FMUL2 / FADD2, sm_100) serves two neighbouring candidates of a row; an odd cloud
  // ends with a point at +inf (never nearest)
  float4* spair;
  // per-point covariances (original order) and, for the PLANE-regularised ones, the direction that got the small eigenvalue
  // (4 doubles per point: unit normal + padding): cov = I - (1 - 1e-3) n n^T up to rounding
  double* cov;
  double* nrm;
  // VGICP voxel map
  double vres;
  int vmin[3], vd[3];
  int vcell;
  int* v_start;   // vcell + 1
  int* v_cnt;     // vcell
  int* v_order;   // 2n: [0,n) point indices grouped by voxel, ascending inside a voxel; [n,2n) the same in scatter order
  int* v_table;   // vcell: record id or -1
  VoxRec* vrec;
  int* v_nrec;    // [1] number of records
  int* v_reccell; // record id -> table cell
  // NDT grid
  float leaf, inv_leaf;
  int min_b[3], max_b[3], div_b[3];
  int ncell_ndt;
  int ndt_centroids;  // the leaves carry their float centroids (needed by the KDTREE neighbourhood search)
  int* n_start;
  int* n_cnt;
  int* n_order;
  int* n_table;
  NdtRec* nrec;
  int* n_nrec;
  int* n_reccell;
};

// ------------------------------------------------------------------------------------------------
// Optimiser loops that never leave the device.  Every method's alignment is { evaluate the cost over the source cloud ;
// advance each pair's state machine } repeated until every pair of the batch has finished.  The pair of kernels is the body
// of a CUDA-graph WHILE node: the step kernel's last block decides whether another round is needed and sets the node's
// condition itself (cudaGraphSetConditional), so a whole alignment — whatever its iteration count — is ONE graph launch
// with no host polling in between (north star: "the Gauss-Newton/Newton update and convergence test stay on the device").
struct LoopCtl {        // device memory, zero-initialised before the loop
  int done;             // pairs that reached their final state
  int blocks_finished;  // step-kernel blocks of the current round that are through (last-block detection)
  int rounds;           // rounds executed
  int pad;
};
struct LoopArgs {       // passed by value to the step kernels
  LoopCtl* ctl;
  int npairs;
  int max_rounds;
  int use_graph;        // 0: the host polls ctl->done between groups of rounds (profiling / fallback path)
  int kernels_per_round;  // kernel nodes in the loop body (launch accounting)
  cudaGraphConditionalHandle handle;
  unsigned long long* rounds_total;  // per-handle counter of kernels run inside graphs (kernel-launch accounting), or nullptr
};
// Called by EVERY thread of a step kernel after its work (no early returns before it).
__device__ __forceinline__ void loop_tail(const LoopArgs& la) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned nblocks = gridDim.x * gridDim.y;
    const int t = atomicAdd(&la.ctl->blocks_finished, 1);
    if (t == (int)nblocks - 1) {  // last block of this round: every pair's state (and ctl->done) is final
      __threadfence();
      la.ctl->blocks_finished = 0;
      const int rounds = atomicAdd(&la.ctl->rounds, 1) + 1;
      const int done = atomicAdd(&la.ctl->done, 0);
      if (la.use_graph && la.rounds_total) atomicAdd(la.rounds_total, (unsigned long long)la.kernels_per_round);
      if (la.use_graph) cudaGraphSetConditional(la.handle, (done < la.npairs && rounds < la.max_rounds) ? 1u : 0u);
    }
  }
}

static_assert(sizeof(b2r_result) == 96, "b2r_result layout");
__host__ __device__ inline void clear_row_padding(b2r_result& r) { r.reserved = 0; }

// ------------------------------------------------------------------------------------------------
// device helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of K doubles held per thread in v[0..K).  Result valid in thread 0..K-1 of warp 0
// as return value for element threadIdx.x; fixed order => deterministic.  smem: K * (blockDim/32) doubles.
template <int K>
__device__ __forceinline__ void block_reduce_to(double* v, double* smem, double* out /*global, K values*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) smem[warp * K + k] = s;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarp; ++w) s += smem[w * K + k];
    out[k] = s;
  }
  __syncthreads();
}

// The same block-wide sum for K <= 32 values per thread with a transposing butterfly inside each warp: at every step a lane
// keeps one half of its values and hands the other half to its partner, so 32 x 32 partial sums collapse in 31 exchanges per lane
// (16 + 8 + 4 + 2 + 1) instead of K x 5; lane l ends up with the warp's total of value l.  Fixed tree order => deterministic.
// (ncu on the 4096-pair batch: the K x 5 shuffle reduction was 15 % of vgicp_eval_kernel's instructions.)
template <int K>
__device__ __forceinline__ void block_reduce_butterfly(const double* v_in, double* smem /* K * nwarp */, double* out /*global, K values*/) {
  static_assert(K <= 32, "at most one value per lane");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  double v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = i < K ? v_in[i] : 0.0;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;  // this lane keeps the upper half of its remaining values
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = up ? v[i] : v[i + half];
      const double keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  if (lane < K) smem[warp * K + lane] = v[0];
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarp; ++w) s += smem[w * K + k];
    out[k] = s;
  }
  __syncthreads();
}

// Block-wide integer sum (every thread gets the result).  smem: blockDim/32 ints; ends with a barrier.
__device__ __forceinline__ int block_sum_int(int v, int* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  int s = 0;
  for (int w = 0; w < nwarp; ++w) s += smem[w];
  __syncthreads();
  return s;
}

// Packed-pair FP32 (sm_100: FFMA2 / FMUL2 and the add as FFMA2 with a unit factor): two IEEE round-to-nearest results per issue
// slot.  Written as PTX with explicit .rn: the CUDA intrinsics (__fmul2_rn / __fadd2_rn) are contracted by the compiler into fused
// multiply-adds like ordinary float arithmetic (seen in SASS, and as a last-bit difference from FLANN's distance), the .rn PTX
// forms are not.
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(unsigned long long r) {
  float2 v;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
  return v;
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {  // a - b
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// FLANN L2_Simple squared distance: float accumulation in dimension order, no FMA contraction.
// The x and y differences and their squares are formed by Blackwell's packed-pair FP32 instructions (FFMA2 / FMUL2: two IEEE
// round-to-nearest results per issue slot, sm_100 only): q - p = fma(p, -1, q) is rounded once, exactly like the subtraction,
// so the value is bit-identical to the scalar chain — 6 instead of 8 issue slots per candidate in kernels that are bound by
// instruction issue (ncu: knn_cov, fitness).  B2R_NO_F32X2 restores the scalar form.
__device__ __forceinline__ float dist2_flann(float ax, float ay, float az, float bx, float by, float bz) {
#if defined(B2R_NO_F32X2)
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  r = __fadd_rn(r, __fmul_rn(dz, dz));
  return r;
#else
  const float2 d = __ffma2_rn(make_float2(bx, by), make_float2(-1.f, -1.f), make_float2(ax, ay));
  const float2 sq = __fmul2_rn(d, d);
  const float dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(sq.x, sq.y), __fmul_rn(dz, dz));
#endif
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ void nn_cell_of(const CloudView& c, float x, float y, float z, int& cx, int& cy, int& cz) {
  cx = clampi((int)floorf((x - c.bmin[0]) * c.inv_hx), 0, c.gd[0] - 1);
  cy = clampi((int)floorf((y - c.bmin[1]) * c.inv_h), 0, c.gd[1] - 1);
  cz = clampi((int)floorf((z - c.bmin[2]) * c.inv_h), 0, c.gd[2] - 1);
}
// total order on floats as ints (negative values reversed; NaNs sort to the ends): keeps "ascending x" well defined
__device__ __forceinline__ int float_order_key(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}

// fast_gicp voxel_coord: floor(x / resolution - 0.5) in double
__device__ __forceinline__ int vgicp_coord_d(double x, double res) { return (int)floor(x / res - 0.5); }

// neighbour offsets: DIRECT1 {0}; DIRECT7 {0,+x,-x,+y,-y,+z,-z}; DIRECT27 full 3x3x3
__device__ __forceinline__ void neighbor_offset(int mode, int o, int& ox, int& oy, int& oz) {
  ox = oy = oz = 0;
  if (mode == B2R_DIRECT7) {
    if (o == 1) ox = 1; else if (o == 2) ox = -1; else if (o == 3) oy = 1; else if (o == 4) oy = -1; else if (o == 5) oz = 1; else if (o == 6) oz = -1;
  } else if (mode == B2R_DIRECT27 || mode == B2R_KDTREE) {
    ox = o / 9 - 1; oy = (o / 3) % 3 - 1; oz = o % 3 - 1;
  }
}

// symmetric 3x3 helpers on 6-vectors (xx,xy,xz,yy,yz,zz)
__device__ __forceinline__ void sym3_inverse(const double* s, double* inv) {
  double c00 = s[3] * s[5] - s[4] * s[4];
  double c01 = s[2] * s[4] - s[1] * s[5];
  double c02 = s[1] * s[4] - s[2] * s[3];
  double det = s[0] * c00 + s[1] * c01 + s[2] * c02;
  double id = 1.0 / det;
  inv[0] = c00 * id;
  inv[1] = c01 * id;
  inv[2] = c02 * id;
  inv[3] = (s[0] * s[5] - s[2] * s[2]) * id;
  inv[4] = (s[1] * s[2] - s[0] * s[4]) * id;
  inv[5] = (s[0] * s[3] - s[1] * s[1]) * id;
}
// R S R^T for symmetric S (6) and row-major R (9) -> symmetric (6)
__device__ __forceinline__ void rsrt(const double* R, const double* s, double* o) {
  double A[9];  // A = R * S
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    A[r * 3 + 0] = R[r * 3 + 0] * s[0] + R[r * 3 + 1] * s[1] + R[r * 3 + 2] * s[2];
    A[r * 3 + 1] = R[r * 3 + 0] * s[1] + R[r * 3 + 1] * s[3] + R[r * 3 + 2] * s[4];
    A[r * 3 + 2] = R[r * 3 + 0] * s[2] + R[r * 3 + 1] * s[4] + R[r * 3 + 2] * s[5];
  }
  o[0] = A[0] * R[0] + A[1] * R[1] + A[2] * R[2];
  o[1] = A[0] * R[3] + A[1] * R[4] + A[2] * R[5];
  o[2] = A[0] * R[6] + A[1] * R[7] + A[2] * R[8];
  o[3] = A[3] * R[3] + A[4] * R[4] + A[5] * R[5];
  o[4] = A[3] * R[6] + A[4] * R[7] + A[5] * R[8];
  o[5] = A[6] * R[6] + A[7] * R[7] + A[8] * R[8];
}

// cyclic Jacobi eigen-decomposition of a symmetric 3x3 (full 9, row-major).  evals ascending, V columns.
__device__ inline void sym3_eigen_dev(const double* Ain, double* evals, double* V) {
  double A[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = Ain[i];
  double Q[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    double diag = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
    if (off <= 1e-34 * diag || off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
      double apq = A[p * 3 + q];
      if (apq == 0.0) continue;
      double app = A[p * 3 + p], aqq = A[q * 3 + q];
      double theta = (aqq - app) / (2.0 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double akp = A[k * 3 + p], akq = A[k * 3 + q];
        A[k * 3 + p] = c * akp - s * akq;
        A[k * 3 + q] = s * akp + c * akq;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double apk = A[p * 3 + k], aqk = A[q * 3 + k];
        A[p * 3 + k] = c * apk - s * aqk;
        A[q * 3 + k] = s * apk + c * aqk;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double qkp = Q[k * 3 + p], qkq = Q[k * 3 + q];
        Q[k * 3 + p] = c * qkp - s * qkq;
        Q[k * 3 + q] = s * qkp + c * qkq;
      }
    }
  }
  // ascending eigenvalues; columns swapped with compile-time indices (a permutation array would be indexed
  // dynamically and push Q into local memory)
  double d0 = A[0], d1 = A[4], d2 = A[8];
#define B2R_EIG_CSWAP(da, db, ca, cb)                                   \
  if (db < da) {                                                        \
    double t = da; da = db; db = t;                                     \
    _Pragma("unroll") for (int i = 0; i < 3; ++i) { t = Q[i * 3 + ca]; Q[i * 3 + ca] = Q[i * 3 + cb]; Q[i * 3 + cb] = t; } \
  }
  B2R_EIG_CSWAP(d0, d1, 0, 1)
  B2R_EIG_CSWAP(d1, d2, 1, 2)
  B2R_EIG_CSWAP(d0, d1, 0, 1)
#undef B2R_EIG_CSWAP
  evals[0] = d0; evals[1] = d1; evals[2] = d2;
#pragma unroll
  for (int i = 0; i < 9; ++i) V[i] = Q[i];
}

}  // namespace b2r
