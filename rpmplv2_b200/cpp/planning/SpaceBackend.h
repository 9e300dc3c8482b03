// SpaceBackend.h — the batched state-space queries the bur / RRT planners of this directory are written against.
//
// It is the batched face of RPMPLv2's base::StateSpace hot methods (include/state_spaces/StateSpace.h:50-55):
//   isValid(q)                       -> isValid            RealVectorSpace.cpp:277-309
//   isValid(q1, q2)                  -> isValidEdges       RealVectorSpace.cpp:261-275
//   computeDistance(q)               -> distanceToSlots    RealVectorSpace.cpp:318-392 (result cached in a State "slot")
//   extendGenSpine / extendGenSpine2 -> extendGenSpines    RGBTConnect.cpp:73-118 (+ RBTConnect::extendSpine, RBTConnect.cpp:85-150)
// The product implementation is CudaBackend (librvs_cuda.so through include/rvs_cuda.h). The CPU oracle implements the same
// interface under oracle/ for the tests and for the CPU baseline of bench.py; nothing in this directory links it.
#ifndef RVS_PLANNING_SPACEBACKEND_H
#define RVS_PLANNING_SPACEBACKEND_H

#include <algorithm>
#include <cstdint>
#include <vector>

namespace rvsplan
{
	struct SpineParams            // RBTConnectConfig / RGBTConnectConfig (data/configurations/configuration_rbtconnect.yaml:4-8, _rgbtconnect.yaml:4)
	{
		int num_layers = 5;
		int num_iter_spine = 1;
		bool use_expanded_bubble = true;
		float d_crit = 0.05f;
	};

	class SpaceBackend
	{
	public:
		virtual ~SpaceBackend() {}
		virtual const char *name() const = 0;
		virtual int numDofs() const = 0;
		virtual const float *limitsLo() const = 0;
		virtual const float *limitsHi() const = 0;
		virtual float equalityThreshold() const = 0;

		// State cache: slots 0 .. capacity-1 hold (q, d_c, d_c_profile, nearest points) of states that served as bur roots.
		virtual void reserveSlots(int capacity) = 0;

		// valid[i] = isValid(q_i) && !robot->checkSelfCollision(q_i)                      (RRTConnect.cpp:13-15)
		virtual void isValid(const float *q, int N, uint8_t *valid) = 0;
		// valid[e] = isValid(q1_e, q2_e) && !robot->checkSelfCollision(q1_e, q2_e)        (RRTConnect.cpp:100)
		virtual void isValidEdges(const float *q1, const float *q2, int E, uint8_t *valid) = 0;
		// Does robot->checkSelfCollision ever answer true (xArm6 with self_collision_checking)?  Only then isValidEdgesDetailed differs
		// from isValidEdges.
		virtual bool hasSelfCollision() const = 0;
		// The two halves of the edge test apart, for callers whose control flow depends on WHICH one failed (RRTx::chooseParent,
		// RRTx.cpp:466-468): obstacle_ok[e] = isValid(q1_e, q2_e); self_hit[e] = robot->checkSelfCollision(q1_e, q2_e) and, on a hit,
		// q2_out[e] = the last collision-free point of the edge the reference re-seats q2 on (xArm6.cpp:284-352; q2_e otherwise).
		virtual void isValidEdgesDetailed(const float *q1, const float *q2, int E, uint8_t *obstacle_ok, uint8_t *self_hit, float *q2_out) = 0;
		// d_c[i] = computeDistance(q_i); everything the spines need later stays in slot slots[i]
		virtual void distanceToSlots(const float *q, int N, const int32_t *slots, float *d_c) = 0;
		// spine s: from the cached state root_slots[s] towards q_targets[s]; q_layers[S][num_layers][n]
		virtual void extendGenSpines(const int32_t *root_slots, const float *q_targets, int S, const SpineParams &sp, bool gen2,
									 float *q_layers, int32_t *n_layers, float *d_c_last, uint8_t *status) = 0;

		// connectGenSpine's loop (RGBTConnect.cpp:120-150) for R requests: request r starts from the cached state root_slots[r] (whose
		// real distance is root_d_c[r]) and repeats, at most `steps` times, { extendGenSpine2 towards its target; computeDistance of the
		// state it reached, cached in new_slots[r][k] (-1: fewer steps for this request) } while the spine Advanced and the current d_c is
		// above D_CRIT.  Outputs R x steps: q_layers (x num_layers x n), n_layers (-1: step not executed), d_c_new, status.
		// This default walks the loop through the two calls above, one step at a time; a backend may do the whole loop in one call.
		virtual void connectGenSpines(const int32_t *root_slots, const float *root_d_c, const float *q_targets, const int32_t *new_slots, int R, int steps,
									  const SpineParams &sp, float *q_layers, int32_t *n_layers, float *d_c_new, uint8_t *status)
		{
			const int n = numDofs();
			const size_t stride = size_t(sp.num_layers) * size_t(n);
			std::vector<int32_t> cur(root_slots, root_slots + R);
			std::vector<float> d_cur(root_d_c, root_d_c + R);
			std::vector<uint8_t> alive(size_t(R), 1);
			std::fill(n_layers, n_layers + size_t(R) * steps, -1);
			std::fill(status, status + size_t(R) * steps, uint8_t(0));
			std::fill(d_c_new, d_c_new + size_t(R) * steps, -1.f);
			for (int k = 0; k < steps; k++)
			{
				std::vector<int> who;
				for (int r = 0; r < R; r++)
				{
					if (alive[size_t(r)] && new_slots[size_t(r) * steps + k] >= 0 && d_cur[size_t(r)] > sp.d_crit) who.push_back(r);
					else alive[size_t(r)] = 0;
				}
				if (who.empty()) break;
				const int S = int(who.size());
				std::vector<int32_t> roots(static_cast<size_t>(S)), fresh(static_cast<size_t>(S)), nl(static_cast<size_t>(S));
				std::vector<float> targets(size_t(S) * n), layers(size_t(S) * stride), d_last(static_cast<size_t>(S)), q_new(size_t(S) * n), d_new(static_cast<size_t>(S));
				std::vector<uint8_t> st(static_cast<size_t>(S));
				for (int i = 0; i < S; i++)
				{
					const int r = who[size_t(i)];
					roots[size_t(i)] = cur[size_t(r)]; fresh[size_t(i)] = new_slots[size_t(r) * steps + k];
					std::copy_n(q_targets + size_t(r) * n, n, targets.begin() + size_t(i) * n);
				}
				extendGenSpines(roots.data(), targets.data(), S, sp, true, layers.data(), nl.data(), d_last.data(), st.data());
				for (int i = 0; i < S; i++)
					std::copy_n(layers.begin() + size_t(i) * stride + size_t(std::max(1, int(nl[size_t(i)])) - 1) * n, n, q_new.begin() + size_t(i) * n);
				distanceToSlots(q_new.data(), S, fresh.data(), d_new.data());
				for (int i = 0; i < S; i++)
				{
					const size_t r = size_t(who[size_t(i)]), o = r * size_t(steps) + size_t(k);
					std::copy_n(layers.begin() + size_t(i) * stride, stride, q_layers + o * stride);
					n_layers[o] = nl[size_t(i)]; d_c_new[o] = d_new[size_t(i)]; status[o] = st[size_t(i)];
					cur[r] = fresh[size_t(i)]; d_cur[r] = d_new[size_t(i)];
					if (st[size_t(i)] != 1) alive[r] = 0;
				}
			}
		}
	};
}

#endif
