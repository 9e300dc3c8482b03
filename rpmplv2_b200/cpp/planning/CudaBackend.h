// CudaBackend.h — SpaceBackend over the C ABI of librvs_cuda.so (include/rvs_cuda.h). No CPU fallback: every call that
// fails throws std::runtime_error with rvs_last_error().
#ifndef RVS_PLANNING_CUDABACKEND_H
#define RVS_PLANNING_CUDABACKEND_H

#include <vector>

#include "Rrtx.h"
#include "SpaceBackend.h"
#include "rvs_cuda.h"

namespace rvsplan
{
	class CudaBackend : public SpaceBackend
	{
	public:
		explicit CudaBackend(rvs_handle h);            // the handle stays owned by the caller
		const char *name() const override { return "RealVectorSpaceCUDA"; }
		int numDofs() const override { return n; }
		const float *limitsLo() const override { return lo.data(); }
		const float *limitsHi() const override { return hi.data(); }
		float equalityThreshold() const override { return eq_thr; }
		void reserveSlots(int capacity) override;
		void isValid(const float *q, int N, uint8_t *valid) override;
		void isValidEdges(const float *q1, const float *q2, int E, uint8_t *valid) override;
		bool hasSelfCollision() const override { return self_collision; }
		void isValidEdgesDetailed(const float *q1, const float *q2, int E, uint8_t *obstacle_ok, uint8_t *self_hit, float *q2_out) override;
		void distanceToSlots(const float *q, int N, const int32_t *slots, float *d_c) override;
		void extendGenSpines(const int32_t *root_slots, const float *q_targets, int S, const SpineParams &sp, bool gen2,
							 float *q_layers, int32_t *n_layers, float *d_c_last, uint8_t *status) override;
		void connectGenSpines(const int32_t *root_slots, const float *root_d_c, const float *q_targets, const int32_t *new_slots, int R, int steps,
							  const SpineParams &sp, float *q_layers, int32_t *n_layers, float *d_c_new, uint8_t *status) override;
	private:
		void check(int status, const char *what) const;
		rvs_handle engine;
		int n = 0;
		float eq_thr = 0.f;
		bool self_collision = false;
		std::vector<float> lo, hi;
	};

	// NeighborIndex over a device-resident tree (rvs_tree_*, csrc/rvs_nn.cuh): the states never leave HBM, a query moves one
	// configuration in and the hits out.
	class CudaNeighborIndex : public NeighborIndex
	{
	public:
		explicit CudaNeighborIndex(rvs_handle h);
		~CudaNeighborIndex() override;
		void append(const float *q) override;
		int nearest(const float *q) override;
		void radius(const float *q, float radius, int exclude, std::vector<int> &idx, std::vector<float> &dist) override;
	private:
		void check(int status, const char *what) const;
		rvs_handle engine;
		int32_t tree = -1;
		int32_t capacity = 1024;
		std::vector<int32_t> buf_idx;
		std::vector<float> buf_dist;
	};
}

#endif
