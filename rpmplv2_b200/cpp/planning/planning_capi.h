// planning_capi.h — shared by the two C entry points that run the planners of Planners.h: rvsplan_solve_batch (CUDA
// backend, this directory) and the CPU-oracle twin under oracle/ (tests and CPU baseline only).
#ifndef RVS_PLANNING_CAPI_H
#define RVS_PLANNING_CAPI_H

#include <algorithm>
#include <vector>

#include "Planners.h"
#include "Rrtx.h"
#include "rvs_planning.h"

namespace rvsplan
{
	inline PlannerConfig config_from_c(const rvsplan_config &c)
	{
		PlannerConfig cfg;
		cfg.kind = PlannerKind(c.kind);
		cfg.max_num_iter = c.max_num_iter; cfg.max_num_states = c.max_num_states; cfg.max_planning_time = c.max_planning_time;
		cfg.eps_step = c.eps_step; cfg.max_extension_steps = c.max_extension_steps; cfg.delta = c.delta; cfg.num_spines = c.num_spines;
		cfg.spine.num_layers = c.spine.num_layers; cfg.spine.num_iter_spine = c.spine.num_iter_spine;
		cfg.spine.use_expanded_bubble = c.spine.use_expanded_bubble != 0; cfg.spine.d_crit = c.spine.d_crit;
		return cfg;
	}

	inline void results_to_c(const std::vector<PlannerResult> &res, int n, rvsplan_result *out, float *paths, int32_t max_path)
	{
		for (size_t p = 0; p < res.size(); p++)
		{
			const PlannerResult &r = res[p];
			rvsplan_result &o = out[p];
			o.success = r.success ? 1 : 0;
			o.invalid_ends = (!r.success && r.iterations == 0 && r.states == 0) ? 1 : 0;
			o.iterations = r.iterations; o.states = r.states; o.planning_time = r.planning_time;
			o.path_len = int32_t(r.path.size() / size_t(n));
			o.rounds = r.rounds; o.num_valid = r.num_valid; o.num_edges = r.num_edges; o.num_distances = r.num_distances; o.num_spines = r.num_spines;
			if (paths && max_path > 0)
			{
				const size_t keep = std::min<size_t>(size_t(o.path_len), size_t(max_path));
				std::copy_n(r.path.begin(), keep * size_t(n), paths + p * size_t(max_path) * size_t(n));
			}
		}
	}

	inline void trees_to_c(const std::vector<PlannerResult> &res, int n, float *tree_coord, int32_t *tree_parent, int32_t *tree_sizes, int32_t max_tree_states)
	{
		for (size_t p = 0; p < res.size(); p++)
			for (int t = 0; t < 2; t++)
			{
				const size_t slot = 2 * p + size_t(t);
				const size_t count = res[p].tree_parent[t].size(), keep = std::min<size_t>(count, size_t(max_tree_states));
				tree_sizes[slot] = int32_t(count);
				std::copy_n(res[p].tree_coord[t].begin(), keep * size_t(n), tree_coord + slot * size_t(max_tree_states) * size_t(n));
				std::copy_n(res[p].tree_parent[t].begin(), keep, tree_parent + slot * size_t(max_tree_states));
			}
	}

	// the body of rvsplan_rrtx_grow / its CPU-oracle twin: 0, or 1 when the tree outgrows max_states
	inline int rrtx_grow(SpaceBackend &backend, NeighborIndex &index, const rvsplan_rrtx_config &c, const float *q_root, const float *samples, int32_t S,
						 int32_t max_states, float *coord, int32_t *parent, float *cost, int32_t *new_index, int32_t path_start, int32_t *path,
						 int32_t *path_len, int32_t *path_invalid, rvsplan_rrtx_stats *stats)
	{
		RrtxConfig cfg;
		cfg.eps_step = c.eps_step; cfg.r_rewire = c.r_rewire; cfg.max_neighbors = c.max_neighbors; cfg.one_call_per_iteration = c.one_call_per_iteration != 0;
		RrtxTree tree(backend, index, cfg);
		const int n = backend.numDofs();
		tree.addRoot(q_root);
		for (int32_t s = 0; s < S; s++)
		{
			if (tree.size() >= max_states) return 1;
			const int id = tree.grow(samples + size_t(s) * size_t(n));
			if (new_index) new_index[s] = id;
		}
		for (int k = 0; k < tree.size(); k++)
		{
			std::copy_n(tree.coord(k), n, coord + size_t(k) * size_t(n));
			parent[k] = tree.parent[size_t(k)];
			cost[k] = tree.cost[size_t(k)];
		}
		if (path_start >= 0 && path_start < tree.size() && path && path_len && path_invalid)
		{
			*path_invalid = tree.pathInvalid(path_start) ? 1 : 0;
			const std::vector<int> p = tree.pathToRoot(path_start);
			*path_len = int32_t(p.size());
			std::copy(p.begin(), p.end(), path);
		}
		if (stats)
		{
			const RrtxStats &t = tree.stats;
			*stats = {int64_t(tree.size()), t.iterations, t.trapped, t.backend_calls, t.edges_checked, t.edges_reference, t.reparented, t.rewired, t.reseated};
		}
		return 0;
	}
}

#endif
