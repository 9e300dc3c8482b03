// planning_capi.cpp — librvs_planning.so: include/rvs_planning.h over CudaBackend. There is no CPU path in this library.
#include "planning_capi.h"

#include <exception>
#include <string>

#include "CudaBackend.h"

namespace
{
	thread_local std::string g_error;
}

extern "C" {

void rvsplan_default_config(int32_t kind, rvsplan_config *out)
{
	if (!out) return;
	out->kind = kind;
	out->max_num_iter = 1000000000; out->max_num_states = 1000000000; out->max_planning_time = 10.f;   // configuration_*.yaml:1-3
	out->eps_step = 0.1f; out->max_extension_steps = 50;                                               // configuration_rrtconnect.yaml:4-5
	out->delta = 3.14159f; out->num_spines = 7;                                                        // configuration_rbtconnect.yaml:5-6
	out->spine.num_layers = kind == RVSPLAN_RBT_CONNECT ? 1 : 5;                                       // configuration_rgbtconnect.yaml:4
	out->spine.num_iter_spine = 1; out->spine.use_expanded_bubble = 1; out->spine.d_crit = 0.05f;      // configuration_rbtconnect.yaml:4,7,8
}

const char *rvsplan_last_error(void) { return g_error.c_str(); }

static int solve_batch_impl(rvs_handle h, const rvsplan_config *cfg, const float *q_start, const float *q_goal, const uint64_t *seeds, int32_t P,
							rvsplan_result *results, float *paths, int32_t max_path, int64_t *pool_stats, float *tree_coord, int32_t *tree_parent,
							int32_t *tree_sizes, int32_t max_tree_states)
{
	if (!h || !cfg || !q_start || !q_goal || !results || P < 0) { g_error = "rvsplan_solve_batch: NULL argument or P < 0"; return RVS_ERR_INVALID_ARGUMENT; }
	if (cfg->kind < RVSPLAN_RRT_CONNECT || cfg->kind > RVSPLAN_RGBT_CONNECT || cfg->num_spines < 1 || cfg->spine.num_layers < 1 ||
		cfg->max_extension_steps < 1 || !(cfg->eps_step > 0)) { g_error = "rvsplan_solve_batch: bad planner configuration"; return RVS_ERR_INVALID_ARGUMENT; }
	const bool want_trees = tree_coord && tree_parent && tree_sizes && max_tree_states > 0;
	try
	{
		rvsplan::CudaBackend backend(h);
		rvsplan::PlannerConfig pc = rvsplan::config_from_c(*cfg);
		pc.keep_trees = want_trees;
		rvsplan::PlannerPool pool(backend, pc);
		const int64_t launches0 = rvs_launch_count(h);
		const std::vector<rvsplan::PlannerResult> res = pool.solve(q_start, q_goal, seeds, P, true);
		rvsplan::results_to_c(res, backend.numDofs(), results, paths, max_path);
		if (want_trees) rvsplan::trees_to_c(res, backend.numDofs(), tree_coord, tree_parent, tree_sizes, max_tree_states);
		if (pool_stats) { pool_stats[0] = pool.rounds(); pool_stats[1] = pool.backendCalls(); pool_stats[2] = rvs_launch_count(h) - launches0; }
	}
	catch (const std::exception &e)
	{
		g_error = e.what();
		return RVS_ERR_CUDA;
	}
	return RVS_OK;
}

int rvsplan_solve_batch(rvs_handle h, const rvsplan_config *cfg, const float *q_start, const float *q_goal, const uint64_t *seeds,
						int32_t P, rvsplan_result *results, float *paths, int32_t max_path, int64_t *pool_stats)
{
	return solve_batch_impl(h, cfg, q_start, q_goal, seeds, P, results, paths, max_path, pool_stats, nullptr, nullptr, nullptr, 0);
}

int rvsplan_solve_batch_trees(rvs_handle h, const rvsplan_config *cfg, const float *q_start, const float *q_goal, const uint64_t *seeds,
							  int32_t P, rvsplan_result *results, float *paths, int32_t max_path, int64_t *pool_stats,
							  float *tree_coord, int32_t *tree_parent, int32_t *tree_sizes, int32_t max_tree_states)
{
	if (!tree_coord || !tree_parent || !tree_sizes || max_tree_states < 1) { g_error = "rvsplan_solve_batch_trees: NULL tree buffers"; return RVS_ERR_INVALID_ARGUMENT; }
	return solve_batch_impl(h, cfg, q_start, q_goal, seeds, P, results, paths, max_path, pool_stats, tree_coord, tree_parent, tree_sizes, max_tree_states);
}

void rvsplan_rrtx_default_config(rvsplan_rrtx_config *out)
{
	if (!out) return;
	out->eps_step = 1.0f; out->r_rewire = 3.14f; out->max_neighbors = 1000;                           // configuration_rrtx.yaml:4-6
	out->one_call_per_iteration = 1; out->device_index = 0;
}

int rvsplan_rrtx_grow(rvs_handle h, const rvsplan_rrtx_config *cfg, const float *q_root, const float *samples, int32_t S, int32_t max_states,
					  float *coord, int32_t *parent, float *cost, int32_t *new_index, int32_t path_start, int32_t *path, int32_t *path_len,
					  int32_t *path_invalid, rvsplan_rrtx_stats *stats)
{
	if (!h || !cfg || !q_root || (S > 0 && !samples) || S < 0 || max_states < 1 || !coord || !parent || !cost)
	{ g_error = "rvsplan_rrtx_grow: NULL argument, S < 0 or max_states < 1"; return RVS_ERR_INVALID_ARGUMENT; }
	if (!(cfg->eps_step > 0) || cfg->r_rewire < 0 || cfg->max_neighbors < 1) { g_error = "rvsplan_rrtx_grow: bad configuration"; return RVS_ERR_INVALID_ARGUMENT; }
	try
	{
		rvsplan::CudaBackend backend(h);
		int rc;
		if (cfg->device_index)
		{
			rvsplan::CudaNeighborIndex index(h);
			rc = rvsplan::rrtx_grow(backend, index, *cfg, q_root, samples, S, max_states, coord, parent, cost, new_index, path_start, path, path_len, path_invalid, stats);
		}
		else
		{
			rvsplan::HostNeighborIndex index(backend.numDofs());
			rc = rvsplan::rrtx_grow(backend, index, *cfg, q_root, samples, S, max_states, coord, parent, cost, new_index, path_start, path, path_len, path_invalid, stats);
		}
		if (rc) { g_error = "rvsplan_rrtx_grow: the tree outgrew max_states"; return RVS_ERR_INVALID_ARGUMENT; }
	}
	catch (const std::exception &e)
	{
		g_error = e.what();
		return RVS_ERR_CUDA;
	}
	return RVS_OK;
}

}
