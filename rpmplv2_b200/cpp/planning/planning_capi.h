// planning_capi.h — shared by the two C entry points that run the planners of Planners.h: rvsplan_solve_batch (CUDA
// backend, this directory) and the CPU-oracle twin under oracle/ (tests and CPU baseline only).
#ifndef RVS_PLANNING_CAPI_H
#define RVS_PLANNING_CAPI_H

#include <algorithm>
#include <vector>

#include "Planners.h"
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
}

#endif
