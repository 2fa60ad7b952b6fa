// TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/README.md).
//
// Stand-in for the subset of the FLANN C++ API that the reference calls:
//   src/obvision/registration/icp/assign/FlannPairAssignment.cpp:48-53,72-82
//   src/obvision/registration/ransacMatching/RandomNormalMatching.cpp:60-63,287-305
// FLANN is a third-party dependency that is absent from /root/reference and from
// this image (CMakeLists.txt:94, docker/Dockerfile:24 `libflann-dev`, most likely
// 1.9.1, version UNPINNED by the reference).
//
// Published behaviour that is restated:
//   * Index<L2<double>> with KDTreeSingleIndexParams, knnSearch(query, indices,
//     dists, 1, SearchParams(-1 /*unlimited checks*/, eps = 0)) returns the EXACT
//     nearest neighbour and its SQUARED Euclidean distance.
//   * L2<T>::operator() accumulates (a[i]-b[i])^2 left to right for dims < 4:
//     d2 = ((0 + dx*dx) + dy*dy)  (flann/algorithms/dist.h, the scalar tail loop).
//   * KDTreeSingleIndex copies the data set when `reorder` is true (default), so
//     the caller may free its buffer after buildIndex()
//     (RandomNormalMatching.cpp:64 relies on that).
// Tie rule: FLANN's result on exactly equal distances depends on its tree layout
// and is unspecified; this shim defines "lowest data index wins" -- the CUDA path
// uses the same rule ("parity unpinned" at this boundary).
//
// The search structure is a small exact kd-tree (median split on the widest
// dimension, leaves of <= 10 points), so the CPU baseline is not penalised by a
// brute-force scan.
#ifndef ORACLE_FLANN_SHIM_HPP
#define ORACLE_FLANN_SHIM_HPP

#include <algorithm>
#include <cassert>
#include <cstddef>
#include <limits>
#include <vector>

namespace flann
{

template <typename T>
class Matrix
{
public:
  size_t rows;
  size_t cols;
  size_t stride;

  Matrix() : rows(0), cols(0), stride(0), data_(NULL) {}
  Matrix(T* data, size_t rows_, size_t cols_, size_t stride_ = 0)
      : rows(rows_), cols(cols_), stride(stride_ ? stride_ : cols_), data_(data) {}

  T* operator[](size_t row) const { return data_ + row * stride; }
  T* ptr() const { return data_; }

private:
  T* data_;
};

template <typename T>
struct L2
{
  typedef T ElementType;
  typedef T ResultType;

  ResultType operator()(const T* a, const T* b, size_t size) const
  {
    ResultType result = ResultType();
    for(size_t i = 0; i < size; i++)
    {
      const ResultType diff = a[i] - b[i];
      result += diff * diff;
    }
    return result;
  }
};

struct IndexParams
{
};

struct KDTreeSingleIndexParams : public IndexParams
{
  KDTreeSingleIndexParams(int leaf_max_size_ = 10, bool reorder_ = true, int dim_ = -1)
      : leaf_max_size(leaf_max_size_), reorder(reorder_), dim(dim_) {}
  int leaf_max_size;
  bool reorder;
  int dim;
};

struct SearchParams
{
  SearchParams(int checks_ = 32, float eps_ = 0.0f, bool sorted_ = true) : checks(checks_), eps(eps_), sorted(sorted_) {}
  int checks;
  float eps;
  bool sorted;
};

template <typename Distance>
class Index
{
public:
  typedef typename Distance::ElementType ElementType;
  typedef typename Distance::ResultType DistanceType;

  Index(const Matrix<ElementType>& features, const KDTreeSingleIndexParams& params, Distance d = Distance())
      : src_(features), leafMax_(params.leaf_max_size > 0 ? params.leaf_max_size : 10), distance_(d), dim_(features.cols)
  {
  }

  void buildIndex()
  {
    const size_t n = src_.rows;
    data_.resize(n * dim_);
    for(size_t i = 0; i < n; i++)
      for(size_t j = 0; j < dim_; j++) data_[i * dim_ + j] = src_[i][j];
    order_.resize(n);
    for(size_t i = 0; i < n; i++) order_[i] = (int)i;
    nodes_.clear();
    if(n > 0) build(0, n);
  }

  int knnSearch(const Matrix<ElementType>& queries, Matrix<int>& indices, Matrix<DistanceType>& dists, size_t knn,
                const SearchParams& /*params*/) const
  {
    assert(knn == 1);
    (void)knn;
    int count = 0;
    for(size_t q = 0; q < queries.rows; q++)
    {
      int best = -1;
      DistanceType bestDist = std::numeric_limits<DistanceType>::infinity();
      if(!nodes_.empty()) search(0, queries[q], best, bestDist);
      if(best >= 0)
      {
        indices[q][0] = best;
        dists[q][0] = bestDist;
        count++;
      }
    }
    return count;
  }

private:
  struct Node
  {
    int left, right;   // children (node indices) or -1 for a leaf
    size_t begin, end; // leaf: range in order_
    int splitDim;
    ElementType splitVal;
  };

  int build(size_t begin, size_t end)
  {
    Node node;
    node.left = node.right = -1;
    node.begin = begin;
    node.end = end;
    node.splitDim = 0;
    node.splitVal = ElementType();
    const int id = (int)nodes_.size();
    nodes_.push_back(node);
    if(end - begin <= (size_t)leafMax_) return id;

    // widest dimension
    int bestDim = 0;
    ElementType bestSpan = -1;
    for(size_t d = 0; d < dim_; d++)
    {
      ElementType lo = data_[order_[begin] * dim_ + d], hi = lo;
      for(size_t i = begin + 1; i < end; i++)
      {
        const ElementType v = data_[order_[i] * dim_ + d];
        if(v < lo) lo = v;
        if(v > hi) hi = v;
      }
      if(hi - lo > bestSpan) { bestSpan = hi - lo; bestDim = (int)d; }
    }
    if(!(bestSpan > 0)) return id; // all points coincide: keep as a (large) leaf

    const size_t mid = begin + (end - begin) / 2;
    const size_t dimN = dim_;
    const std::vector<ElementType>& dat = data_;
    std::nth_element(order_.begin() + begin, order_.begin() + mid, order_.begin() + end,
                     [&dat, dimN, bestDim](int a, int b) { return dat[a * dimN + bestDim] < dat[b * dimN + bestDim]; });
    const ElementType splitVal = data_[order_[mid] * dim_ + bestDim];
    const int left = build(begin, mid);
    const int right = build(mid, end);
    nodes_[id].left = left;
    nodes_[id].right = right;
    nodes_[id].splitDim = bestDim;
    nodes_[id].splitVal = splitVal;
    return id;
  }

  void search(int id, const ElementType* q, int& best, DistanceType& bestDist) const
  {
    const Node& node = nodes_[id];
    if(node.left < 0)
    {
      for(size_t i = node.begin; i < node.end; i++)
      {
        const int idx = order_[i];
        const DistanceType d = distance_(q, &data_[idx * dim_], dim_);
        if(d < bestDist || (d == bestDist && idx < best))
        {
          bestDist = d;
          best = idx;
        }
      }
      return;
    }
    const ElementType diff = q[node.splitDim] - node.splitVal;
    const int nearChild = (diff < 0) ? node.left : node.right;
    const int farChild = (diff < 0) ? node.right : node.left;
    search(nearChild, q, best, bestDist);
    // <= keeps exact ties reachable so that the lowest index can win
    if(diff * diff <= bestDist) search(farChild, q, best, bestDist);
  }

  Matrix<ElementType> src_;
  int leafMax_;
  Distance distance_;
  size_t dim_;
  std::vector<ElementType> data_;
  std::vector<int> order_;
  std::vector<Node> nodes_;
};

} // namespace flann

#endif
