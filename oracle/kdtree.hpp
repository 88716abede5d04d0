// ORACLE — TEST INFRASTRUCTURE ONLY.
// Exact nearest-neighbour search standing in for pcl::KdTreeFLANN (FLANN single kd-tree, eps = 0,
// L2_Simple float distance; not vendored). Consumed by the reference at
//   voxel_grid_covariance_omp.h:295,485 (centroid radius search)
//   gicp_omp_impl.hpp:79 (k-NN covariances), gicp_omp.h:343 (1-NN correspondences)
//   pcl::Registration::getFitnessScore (graph_based_slam_component.cpp:231)
// Squared distances are accumulated in float exactly as FLANN's L2_Simple does
// (((0 + dx*dx) + dy*dy) + dz*dz); ties are broken towards the lower point index
// (FLANN's tie order is unspecified).
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <queue>
#include <utility>
#include <vector>

namespace oracle {

struct P3 {
  float x, y, z;
};

inline float dist2f(const P3& a, const P3& b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  float r = dx * dx;
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

class KdTree {
 public:
  void build(const std::vector<P3>& pts) {
    pts_ = &pts;
    idx_.resize(pts.size());
    std::iota(idx_.begin(), idx_.end(), 0);
    nodes_.clear();
    nodes_.reserve(pts.size() / 4 + 16);
    if (!pts.empty()) build_rec(0, (int)pts.size());
  }
  bool empty() const { return idx_.empty(); }

  // k nearest, ascending by (d2, index). Returns number found.
  int knn(const P3& q, int k, std::vector<int>& out_idx, std::vector<float>& out_d2) const {
    out_idx.clear();
    out_d2.clear();
    if (nodes_.empty() || k <= 0) return 0;
    Heap heap;
    search_knn(0, q, k, heap);
    int n = (int)heap.size();
    out_idx.resize(n);
    out_d2.resize(n);
    for (int i = n - 1; i >= 0; i--) {
      out_d2[i] = heap.top().first;
      out_idx[i] = heap.top().second;
      heap.pop();
    }
    return n;
  }

  // all points with d2 <= r2, ascending by (d2, index)
  int radius(const P3& q, float r2, std::vector<int>& out_idx, std::vector<float>& out_d2) const {
    std::vector<std::pair<float, int>> res;
    if (!nodes_.empty()) search_radius(0, q, r2, res);
    std::sort(res.begin(), res.end());
    out_idx.resize(res.size());
    out_d2.resize(res.size());
    for (size_t i = 0; i < res.size(); i++) {
      out_d2[i] = res[i].first;
      out_idx[i] = res[i].second;
    }
    return (int)res.size();
  }

 private:
  struct Node {
    int lo, hi;        // leaf: idx_ range
    int left, right;   // children (-1 for leaf)
    int dim;
    float split;
    float bmin[3], bmax[3];
  };
  using Heap = std::priority_queue<std::pair<float, int>>;  // max-heap on (d2, index)

  int build_rec(int lo, int hi) {
    Node n;
    n.lo = lo;
    n.hi = hi;
    n.left = n.right = -1;
    n.dim = 0;
    n.split = 0;
    for (int d = 0; d < 3; d++) {
      n.bmin[d] = 3.4e38f;
      n.bmax[d] = -3.4e38f;
    }
    for (int i = lo; i < hi; i++) {
      const P3& p = (*pts_)[idx_[i]];
      const float c[3] = {p.x, p.y, p.z};
      for (int d = 0; d < 3; d++) {
        n.bmin[d] = std::min(n.bmin[d], c[d]);
        n.bmax[d] = std::max(n.bmax[d], c[d]);
      }
    }
    int me = (int)nodes_.size();
    nodes_.push_back(n);
    if (hi - lo > 12) {
      int dim = 0;
      float ext = -1;
      for (int d = 0; d < 3; d++)
        if (n.bmax[d] - n.bmin[d] > ext) {
          ext = n.bmax[d] - n.bmin[d];
          dim = d;
        }
      if (ext > 0) {
        int mid = (lo + hi) / 2;
        auto key = [&](int i) {
          const P3& p = (*pts_)[i];
          return dim == 0 ? p.x : (dim == 1 ? p.y : p.z);
        };
        std::nth_element(idx_.begin() + lo, idx_.begin() + mid, idx_.begin() + hi,
                         [&](int a, int b) { return key(a) < key(b); });
        nodes_[me].dim = dim;
        nodes_[me].split = key(idx_[mid]);
        int l = build_rec(lo, mid);
        int r = build_rec(mid, hi);
        nodes_[me].left = l;
        nodes_[me].right = r;
      }
    }
    return me;
  }

  static float box_d2(const Node& n, const P3& q) {
    const float c[3] = {q.x, q.y, q.z};
    double s = 0;  // conservative lower bound (double so rounding never prunes a valid candidate)
    for (int d = 0; d < 3; d++) {
      double e = 0;
      if (c[d] < n.bmin[d]) e = (double)n.bmin[d] - c[d];
      else if (c[d] > n.bmax[d]) e = (double)c[d] - n.bmax[d];
      s += e * e;
    }
    return (float)(s * (1.0 - 1e-6));
  }

  void search_knn(int ni, const P3& q, int k, Heap& heap) const {
    const Node& n = nodes_[ni];
    if ((int)heap.size() == k && box_d2(n, q) > heap.top().first) return;
    if (n.left < 0) {
      for (int i = n.lo; i < n.hi; i++) {
        int id = idx_[i];
        std::pair<float, int> cand(dist2f(q, (*pts_)[id]), id);
        if ((int)heap.size() < k) heap.push(cand);
        else if (cand < heap.top()) {
          heap.pop();
          heap.push(cand);
        }
      }
      return;
    }
    float qv = n.dim == 0 ? q.x : (n.dim == 1 ? q.y : q.z);
    int first = qv < n.split ? n.left : n.right;
    int second = qv < n.split ? n.right : n.left;
    search_knn(first, q, k, heap);
    search_knn(second, q, k, heap);
  }

  void search_radius(int ni, const P3& q, float r2, std::vector<std::pair<float, int>>& res) const {
    const Node& n = nodes_[ni];
    if (box_d2(n, q) > r2) return;
    if (n.left < 0) {
      for (int i = n.lo; i < n.hi; i++) {
        int id = idx_[i];
        float d2 = dist2f(q, (*pts_)[id]);
        if (d2 <= r2) res.emplace_back(d2, id);
      }
      return;
    }
    search_radius(n.left, q, r2, res);
    search_radius(n.right, q, r2, res);
  }

  const std::vector<P3>* pts_ = nullptr;
  std::vector<int> idx_;
  std::vector<Node> nodes_;
};

}  // namespace oracle
