// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// Exact kd-tree (kNN sorted by distance, radius search) standing in for
// pcl::search::KdTree / pcl::KdTreeFLANN (FLANN KDTreeSingleIndex, exact search,
// sorted results) which the reference uses implicitly through pcl::Registration
// (apps/scan_matching_odometry_component.cpp:412,
// src/mrg_slam/information_matrix_calculator.cpp:51-67) and inside fast_gicp /
// PCL filters.  Distance = FLANN L2_Simple: float accumulation in dimension
// order ((dx*dx) + dy*dy) + dz*dz.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace orc {

struct KdTree {
  struct Node {
    int left = -1, right = -1;  // children, or -1
    int begin = 0, end = 0;     // leaf range in idx_
    int dim = -1;
    float split = 0.f;
    float lo = 0.f, hi = 0.f;  // max of left along dim, min of right along dim
  };
  const float* pts_ = nullptr;  // stride 4 floats (x,y,z,intensity)
  int n_ = 0;
  std::vector<int> idx_;
  std::vector<Node> nodes_;
  static constexpr int kLeaf = 12;

  void build(const float* xyzi, int n) {
    pts_ = xyzi;
    n_ = n;
    idx_.resize(n);
    std::iota(idx_.begin(), idx_.end(), 0);
    nodes_.clear();
    nodes_.reserve(n / 4 + 16);
    if (n > 0) build_rec(0, n);
  }

  static inline float dist2(const float* a, const float* b) {
    float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    float r = d0 * d0;
    r = r + d1 * d1;
    r = r + d2 * d2;
    return r;
  }

  // k nearest, ascending (distance, index).  Returns number found (min(k, n)).
  int knn(const float* q, int k, int* out_idx, float* out_d2) const {
    if (n_ == 0 || k <= 0) return 0;
    Heap h{out_idx, out_d2, 0, k};
    search_knn(0, q, h);
    // heap -> ascending order
    int m = h.size;
    std::vector<std::pair<float, int>> tmp(m);
    for (int i = 0; i < m; ++i) tmp[i] = {out_d2[i], out_idx[i]};
    std::sort(tmp.begin(), tmp.end());
    for (int i = 0; i < m; ++i) { out_d2[i] = tmp[i].first; out_idx[i] = tmp[i].second; }
    return m;
  }

  // Exact nearest neighbour of a DOUBLE query under double squared distance (small_gicp's own kd-tree searches in
  // double: `(traits::point(points, index) - query).squaredNorm()` on Vector4d, Packet2d order (dx^2 + dz^2) + dy^2).
  // Ties go to the lower index.  Returns the index or -1 if the tree is empty.
  int nn_double(const double* q, double* d2_out) const {
    int best = -1;
    double bd = INFINITY;
    if (n_) search_nn_d(0, q, best, bd);
    *d2_out = bd;
    return best;
  }

  // Count of points with d2 < r2 (strict, FLANN RadiusResultSet); stops early
  // once the count exceeds `stop_above` (pass INT32_MAX for a full count).
  int radius_count(const float* q, float r2, int stop_above) const {
    int cnt = 0;
    if (n_) search_radius(0, q, r2, stop_above, cnt);
    return cnt;
  }

 private:
  struct Heap {  // bounded max-heap on (d2, idx)
    int* idx; float* d2; int size; int cap;
    inline bool less(int a, int b) const { return d2[a] < d2[b] || (d2[a] == d2[b] && idx[a] < idx[b]); }
    inline float worst() const { return size < cap ? INFINITY : d2[0]; }
    void push(float d, int i) {
      if (size < cap) {
        int c = size++;
        d2[c] = d; idx[c] = i;
        while (c > 0) {
          int p = (c - 1) / 2;
          if (less(p, c)) { std::swap(d2[p], d2[c]); std::swap(idx[p], idx[c]); c = p; } else break;
        }
      } else if (d < d2[0] || (d == d2[0] && i < idx[0])) {
        d2[0] = d; idx[0] = i;
        int c = 0;
        for (;;) {
          int l = 2 * c + 1, r = l + 1, m = c;
          if (l < size && less(m, l)) m = l;
          if (r < size && less(m, r)) m = r;
          if (m == c) break;
          std::swap(d2[m], d2[c]); std::swap(idx[m], idx[c]); c = m;
        }
      }
    }
  };

  int build_rec(int b, int e) {
    int id = (int)nodes_.size();
    nodes_.push_back(Node());
    if (e - b <= kLeaf) {
      nodes_[id].begin = b; nodes_[id].end = e;
      return id;
    }
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = b; i < e; ++i) {
      const float* p = pts_ + 4 * (size_t)idx_[i];
      for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], p[d]); mx[d] = std::max(mx[d], p[d]); }
    }
    int dim = 0;
    if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
    if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
    int mid = (b + e) / 2;
    std::nth_element(idx_.begin() + b, idx_.begin() + mid, idx_.begin() + e, [&](int a, int c) {
      float va = pts_[4 * (size_t)a + dim], vc = pts_[4 * (size_t)c + dim];
      return va < vc || (va == vc && a < c);
    });
    float lo = -INFINITY, hi = INFINITY;
    for (int i = b; i < mid; ++i) lo = std::max(lo, pts_[4 * (size_t)idx_[i] + dim]);
    for (int i = mid; i < e; ++i) hi = std::min(hi, pts_[4 * (size_t)idx_[i] + dim]);
    nodes_[id].dim = dim;
    nodes_[id].split = pts_[4 * (size_t)idx_[mid] + dim];
    nodes_[id].lo = lo; nodes_[id].hi = hi;
    nodes_[id].begin = b; nodes_[id].end = e;
    int l = build_rec(b, mid);
    int r = build_rec(mid, e);
    nodes_[id].left = l; nodes_[id].right = r;
    return id;
  }

  void search_knn(int id, const float* q, Heap& h) const {
    const Node& nd = nodes_[id];
    if (nd.dim < 0) {
      for (int i = nd.begin; i < nd.end; ++i) {
        int pi = idx_[i];
        h.push(dist2(q, pts_ + 4 * (size_t)pi), pi);
      }
      return;
    }
    // exact distance from q to each child's slab along dim (conservative in double)
    double dl = (double)q[nd.dim] - (double)nd.lo;  // >0 when q is right of the left child
    double dr = (double)nd.hi - (double)q[nd.dim];  // >0 when q is left of the right child
    int first = nd.left, second = nd.right;
    double dsec = dr;
    if (q[nd.dim] >= nd.split) { first = nd.right; second = nd.left; dsec = dl; }
    search_knn(first, q, h);
    double bound = dsec > 0 ? dsec * dsec : 0.0;
    if (bound * (1.0 - 1e-6) <= (double)h.worst()) search_knn(second, q, h);
  }

  void search_nn_d(int id, const double* q, int& best, double& bd) const {
    const Node& nd = nodes_[id];
    if (nd.dim < 0) {
      for (int i = nd.begin; i < nd.end; ++i) {
        const int pi = idx_[i];
        const float* p = pts_ + 4 * (size_t)pi;
        const double d0 = (double)p[0] - q[0], d1 = (double)p[1] - q[1], d2 = (double)p[2] - q[2];
        const double d = (d0 * d0 + d2 * d2) + d1 * d1;
        if (d < bd || (d == bd && pi < best)) { bd = d; best = pi; }
      }
      return;
    }
    const double dl = q[nd.dim] - (double)nd.lo, dr = (double)nd.hi - q[nd.dim];
    int first = nd.left, second = nd.right;
    double dsec = dr;
    if (q[nd.dim] >= (double)nd.split) { first = nd.right; second = nd.left; dsec = dl; }
    search_nn_d(first, q, best, bd);
    const double bound = dsec > 0 ? dsec * dsec : 0.0;
    if (bound * (1.0 - 1e-12) <= bd) search_nn_d(second, q, best, bd);
  }

  void search_radius(int id, const float* q, float r2, int stop_above, int& cnt) const {
    if (cnt > stop_above) return;
    const Node& nd = nodes_[id];
    if (nd.dim < 0) {
      for (int i = nd.begin; i < nd.end; ++i)
        if (dist2(q, pts_ + 4 * (size_t)idx_[i]) < r2) ++cnt;
      return;
    }
    double dl = (double)q[nd.dim] - (double)nd.lo;
    double dr = (double)nd.hi - (double)q[nd.dim];
    if (!(dl > 0 && dl * dl * (1.0 - 1e-6) > (double)r2)) search_radius(nd.left, q, r2, stop_above, cnt);
    if (!(dr > 0 && dr * dr * (1.0 - 1e-6) > (double)r2)) search_radius(nd.right, q, r2, stop_above, cnt);
  }
};

}  // namespace orc
