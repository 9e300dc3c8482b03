// oracle/env_shim/nanoflann.hpp — stand-in for nanoflann (base::Tree's optional kd-tree): declarations only, enough for
// /root/reference/src/state_spaces/Tree.cpp to compile; the harness never builds a kd-tree (base::Tree falls back to its own linear
// search when kd_tree == nullptr).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <fstream>      // the planners' outputPlannerData uses std::ofstream and rely on a transitive include
namespace nanoflann {
template <class T, class DS> struct L2_Simple_Adaptor {};
struct KDTreeSingleIndexAdaptorParams { explicit KDTreeSingleIndexAdaptorParams(size_t = 10) {} };
struct SearchParams { explicit SearchParams(int = 32, float = 0, bool = true) {} };
template <class T> struct KNNResultSet {
    explicit KNNResultSet(size_t) {}
    void init(size_t *, T *) {}
};
template <class Dist, class DS, int DIM = -1> class KDTreeSingleIndexDynamicAdaptor {
public:
    KDTreeSingleIndexDynamicAdaptor(int, DS &, const KDTreeSingleIndexAdaptorParams & = KDTreeSingleIndexAdaptorParams()) {}
    void addPoints(size_t, size_t) {}
    void removePoint(size_t) {}
    template <class RS> bool findNeighbors(RS &, const float *, const SearchParams &) const { return false; }
};
}  // namespace nanoflann
