// Rrtx.h — the edge-batch consumers of RRTx written for a BATCHED state space (SpaceBackend.h).
//
// What is restated from RPMPLv2 (control flow only; every collision answer comes from the backend):
//   RRTx::solve, the tree-growth block of an iteration      src/planners/rrtx/RRTx.cpp:157-196 (phase 2) and :113-127 (phase 1)
//   RRTx::findNeighbors                                     src/planners/rrtx/RRTx.cpp:424-452
//   RRTx::chooseParent                                      src/planners/rrtx/RRTx.cpp:454-491
//   RRTx::rewireNeighbors / propagateCostChanges            src/planners/rrtx/RRTx.cpp:493-548
//   RRTx::updatePath / computePath                          src/planners/rrtx/RRTx.cpp:550-593
//   RRTx::shrinkingBallRadius and the constants behind it   src/planners/rrtx/RRTx.cpp:71-75, :636-639
//   RRTConnect::extend                                      src/planners/rrt/RRTConnect.cpp:92-106
//   base::Tree::getNearestState / upgradeTree               src/state_spaces/Tree.cpp:35-50, :90-103
// Not restated: the trajectory side of an iteration (UpdatingState, splines, MotionValidity — the latter is rvs_motion_validity), and
// handleDynamicObstacles, which is dead code in the reference (nothing ever seeds orphan_set: updateObstacles is declared, never defined).
//
// What is new: the reference asks the state space for ONE edge at a time, and only while a candidate's cost still beats the running
// minimum.  The verdict of an edge is a pure function of the edge, so here every candidate edge of an iteration goes out as ONE
// batched query (the extension edge, neighbour -> q_new for chooseParent, q_new -> neighbour for rewireNeighbors: up to
// 1 + 2 * MAX_NEIGHBORS edges of NUM_INTERPOLATION_VALIDITY_CHECKS points each) and the reference's sequential cost logic is then
// replayed over the verdicts: the resulting tree (parents, costs, children lists) is the reference's, state by state.
// Reference behaviour that is kept on purpose because it shapes the costs (tests/test_rrtx.py pins it against the compiled RRTx.cpp):
//   * chooseParent adds q_new to its parent's children, upgradeTree adds it AGAIN: a re-parented node stays in its old parent's list
//     once (removeChild erases one entry), so a later propagateCostChanges from the old parent can still lower its cost;
//   * the root keeps State's default cost of -1 (State.cpp:12), so every cost-to-come is offset by -1;
//   * a q_new none of whose neighbours gives a valid edge enters the tree WITHOUT a parent (chooseParent starts from a null parent).
#ifndef RVS_PLANNING_RRTX_H
#define RVS_PLANNING_RRTX_H

#include <cstdint>
#include <vector>

#include "SpaceBackend.h"

namespace rvsplan
{
	struct RrtxConfig                    // RRTxConfig (data/configurations/configuration_rrtx.yaml; src/configurations/RRTxConfig.cpp)
	{
		float eps_step = 0.1f;           // EPS_STEP
		float r_rewire = 0.f;            // R_REWIRE: 0 = shrinking ball
		int max_neighbors = 1000;        // MAX_NEIGHBORS
		bool one_call_per_iteration = true;   // every candidate edge of an iteration in one backend call (false: three calls, fewer edges)
	};

	// Nearest-neighbour / radius queries over the states of the tree. The host index scans linearly; the CUDA index keeps the states
	// in HBM (rvs_tree_*, csrc/rvs_nn.cuh) — same answers, the arithmetic order of both is the reference's.
	class NeighborIndex
	{
	public:
		virtual ~NeighborIndex() {}
		virtual void append(const float *q) = 0;
		virtual int nearest(const float *q) = 0;                                           // squared L2, sequential sum; ties -> lowest index
		// states with ss->getNorm(q, state) <= radius in tree order, except `exclude`
		virtual void radius(const float *q, float radius, int exclude, std::vector<int> &idx, std::vector<float> &dist) = 0;
	};

	class HostNeighborIndex : public NeighborIndex
	{
	public:
		explicit HostNeighborIndex(int n_) : n(n_) {}
		void append(const float *q) override { pts.insert(pts.end(), q, q + n); }
		int nearest(const float *q) override;
		void radius(const float *q, float radius, int exclude, std::vector<int> &idx, std::vector<float> &dist) override;
	private:
		int n;
		std::vector<float> pts;
	};

	struct RrtxStats
	{
		int64_t iterations = 0, trapped = 0, backend_calls = 0, edges_checked = 0, edges_reference = 0, reparented = 0, rewired = 0, reseated = 0;
	};

	class RrtxTree
	{
	public:
		RrtxTree(SpaceBackend &backend_, NeighborIndex &index_, const RrtxConfig &cfg_);
		int addRoot(const float *q);                               // RRTx ctor: tree->upgradeTree(q_goal, nullptr); cost = State's default -1
		// one tree-growth block (RRTx.cpp:166-190) towards q_rand: index of q_new, or -1 when the extension is Trapped
		int grow(const float *q_rand);
		float rewireRadius() const;                                // R_REWIRE, or shrinkingBallRadius(number of states)
		// updatePath's walk (:558-570): is any edge from `start` to the root invalid?  One batched query for the whole chain.
		bool pathInvalid(int start);
		std::vector<int> pathToRoot(int start) const;              // computePath (:575-592)

		int size() const { return int(parent.size()); }
		const float *coord(int i) const { return &q[size_t(i) * size_t(n)]; }
		std::vector<int> parent;
		std::vector<float> cost;
		std::vector<std::vector<int>> children;                    // State::children, duplicates and stale entries included (see above)
		RrtxStats stats;
	private:
		float dist(int a, int b) const;
		void findNeighbors(const float *q_new, int self, float radius, std::vector<int> &nb, std::vector<float> &nb_dist);
		void removeChild(int p, int child);
		void propagateCostChanges(int node);
		SpaceBackend &backend;
		NeighborIndex &index;
		RrtxConfig cfg;
		int n;
		float eta = 0.f, gamma_rrt = 0.f;
		std::vector<float> q;
	};
}

#endif
