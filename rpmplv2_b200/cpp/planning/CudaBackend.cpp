#include "CudaBackend.h"

#include <stdexcept>
#include <string>

namespace rvsplan
{
	CudaBackend::CudaBackend(rvs_handle h) : engine(h)
	{
		if (!h) throw std::runtime_error("CudaBackend: NULL engine handle (there is no CPU fallback)");
		n = rvs_num_dofs(h);
		lo.resize(size_t(n));
		hi.resize(size_t(n));
		check(rvs_get_limits(h, lo.data(), hi.data()), "rvs_get_limits");
		rvs_params prm;
		check(rvs_get_params(h, &prm), "rvs_get_params");
		eq_thr = prm.equality_threshold;
		int32_t enabled = 0;
		float eps = 0.f;
		check(rvs_get_self_collision(h, &enabled, &eps), "rvs_get_self_collision");
		self_collision = enabled != 0;
	}

	void CudaBackend::check(int status, const char *what) const
	{
		if (status != RVS_OK) throw std::runtime_error(std::string(what) + ": " + rvs_last_error(engine));
	}

	void CudaBackend::reserveSlots(int capacity) { check(rvs_store_reserve(engine, capacity), "rvs_store_reserve"); }

	// the planners always pair the obstacle test with the robot's self-collision test (RRTConnect.cpp:13-15, :100): one engine call each
	void CudaBackend::isValid(const float *q, int N, uint8_t *valid) { check(rvs_check_states(engine, q, N, valid), "rvs_check_states"); }

	void CudaBackend::isValidEdges(const float *q1, const float *q2, int E, uint8_t *valid)
	{
		check(rvs_check_edges(engine, q1, q2, E, valid), "rvs_check_edges");
	}

	void CudaBackend::isValidEdgesDetailed(const float *q1, const float *q2, int E, uint8_t *obstacle_ok, uint8_t *self_hit, float *q2_out)
	{
		check(rvs_check_edges_detailed(engine, q1, q2, E, obstacle_ok, self_hit, q2_out), "rvs_check_edges_detailed");
	}

	void CudaBackend::distanceToSlots(const float *q, int N, const int32_t *slots, float *d_c)
	{
		check(rvs_store_distance(engine, q, N, slots, d_c, nullptr), "rvs_store_distance");
	}

	void CudaBackend::extendGenSpines(const int32_t *root_slots, const float *q_targets, int S, const SpineParams &sp, bool gen2,
									  float *q_layers, int32_t *n_layers, float *d_c_last, uint8_t *status)
	{
		rvs_spine_params p;
		p.num_layers = sp.num_layers; p.num_iter_spine = sp.num_iter_spine; p.use_expanded_bubble = sp.use_expanded_bubble ? 1 : 0; p.d_crit = sp.d_crit;
		check(rvs_store_extend_spines(engine, root_slots, q_targets, S, &p, gen2 ? 1 : 0, q_layers, n_layers, d_c_last, status), "rvs_store_extend_spines");
	}

	// the whole loop in ONE engine call: its 2 x steps kernels gate themselves on device state (rvs_store_connect)
	void CudaBackend::connectGenSpines(const int32_t *root_slots, const float *, const float *q_targets, const int32_t *new_slots, int R, int steps,
									   const SpineParams &sp, float *q_layers, int32_t *n_layers, float *d_c_new, uint8_t *status)
	{
		rvs_spine_params p;
		p.num_layers = sp.num_layers; p.num_iter_spine = sp.num_iter_spine; p.use_expanded_bubble = sp.use_expanded_bubble ? 1 : 0; p.d_crit = sp.d_crit;
		check(rvs_store_connect(engine, root_slots, q_targets, new_slots, R, steps, &p, q_layers, n_layers, d_c_new, status), "rvs_store_connect");
	}

	CudaNeighborIndex::CudaNeighborIndex(rvs_handle h) : engine(h)
	{
		if (!h) throw std::runtime_error("CudaNeighborIndex: NULL engine handle");
		check(rvs_tree_create(h, &tree), "rvs_tree_create");
	}
	CudaNeighborIndex::~CudaNeighborIndex() { if (tree >= 0) rvs_tree_release(engine, tree); }
	void CudaNeighborIndex::check(int status, const char *what) const
	{
		if (status != RVS_OK) throw std::runtime_error(std::string(what) + ": " + rvs_last_error(engine));
	}
	void CudaNeighborIndex::append(const float *q) { check(rvs_tree_append(engine, tree, q, 1), "rvs_tree_append"); }
	int CudaNeighborIndex::nearest(const float *q)
	{
		int32_t idx = -1;
		check(rvs_tree_nearest(engine, &tree, q, 1, &idx, nullptr), "rvs_tree_nearest");
		return idx;
	}
	void CudaNeighborIndex::radius(const float *q, float radius_, int exclude, std::vector<int> &idx, std::vector<float> &dist)
	{
		for (;;)
		{
			buf_idx.resize(size_t(capacity)); buf_dist.resize(size_t(capacity));
			int32_t count = 0;
			check(rvs_tree_radius(engine, tree, q, radius_, exclude, capacity, buf_idx.data(), buf_dist.data(), &count), "rvs_tree_radius");
			if (count <= capacity)
			{
				idx.assign(buf_idx.begin(), buf_idx.begin() + count);
				dist.assign(buf_dist.begin(), buf_dist.begin() + count);
				return;
			}
			capacity = count + count / 2;
		}
	}
}
