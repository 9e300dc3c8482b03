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

#include <cstdint>

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
	};
}

#endif
