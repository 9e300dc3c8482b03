// oracle/env_shim/nanoflann.hpp — stand-in for the slice of nanoflann that base::Tree uses (/root/reference/include/state_spaces/Tree.h:16-17,
// src/state_spaces/Tree.cpp:35-51,86-100): KDTreeSingleIndexDynamicAdaptor<L2_Simple_Adaptor<float, Tree>, Tree> with addPoints /
// removePoint / findNeighbors(KNNResultSet).  nanoflann answers EXACT nearest-neighbour queries in the squared L2 metric, so an
// exhaustive scan over the dataset adaptor (kdtree_get_point_count / kdtree_get_pt) returns the same state whenever the minimum is
// unique; the first index wins a tie.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
#include <cstddef>
#include <fstream>      // the planners' outputPlannerData uses std::ofstream and relies on a transitive include
#include <limits>
#include <vector>
namespace nanoflann {
template <class T, class DS> struct L2_Simple_Adaptor {};
struct KDTreeSingleIndexAdaptorParams { explicit KDTreeSingleIndexAdaptorParams(size_t = 10) {} };
struct SearchParams { explicit SearchParams(int = 32, float = 0, bool = true) {} };
template <class T> struct KNNResultSet {
    size_t capacity, *indices = nullptr;
    T *dists = nullptr;
    explicit KNNResultSet(size_t capacity_) : capacity(capacity_) {}
    void init(size_t *indices_, T *dists_) { indices = indices_; dists = dists_; }
};
template <class Dist, class DS, int DIM = -1> class KDTreeSingleIndexDynamicAdaptor {
    int dim_;
    DS &data_;
    std::vector<char> removed_;

public:
    KDTreeSingleIndexDynamicAdaptor(int dim, DS &data, const KDTreeSingleIndexAdaptorParams & = KDTreeSingleIndexAdaptorParams()) : dim_(dim), data_(data) {}
    void addPoints(size_t first, size_t last) { if (removed_.size() <= last) removed_.resize(last + 1, 0); for (size_t i = first; i <= last; i++) removed_[i] = 0; }
    void removePoint(size_t idx) { if (idx < removed_.size()) removed_[idx] = 1; }
    // only the 1-nearest query of Tree::getNearestState is served
    template <class RS> bool findNeighbors(RS &result, const float *query, const SearchParams &) const {
        float best = std::numeric_limits<float>::infinity();
        size_t best_idx = 0;
        bool any = false;
        const size_t count = data_.kdtree_get_point_count();
        for (size_t i = 0; i < count; i++) {
            if (i < removed_.size() && removed_[i]) continue;
            float s = 0.f;                                           // L2_Simple_Adaptor::evalMetric: sequential sum of squared differences
            for (int d = 0; d < dim_; d++) { const float diff = query[d] - data_.kdtree_get_pt(i, size_t(d)); s += diff * diff; }
            if (s < best) { best = s; best_idx = i; any = true; }
        }
        if (any && result.indices) { result.indices[0] = best_idx; result.dists[0] = best; }
        return any;
    }
};
}  // namespace nanoflann
