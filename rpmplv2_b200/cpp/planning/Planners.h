// Planners.h — RRT-Connect, RBT-Connect and RGBT-Connect written for a BATCHED state space (SpaceBackend.h).
//
// What is restated from RPMPLv2 (control flow only; every geometric answer comes from the backend):
//   RRTConnect::solve / extend / connect / computePath            src/planners/rrt/RRTConnect.cpp:38-150
//   RBTConnect::solve / getRandomState / connectSpine             src/planners/rbt/RBTConnect.cpp:14-185
//   RGBTConnect::solve / extendGenSpine2 / connectGenSpine        src/planners/rbt/RGBTConnect.cpp:14-150
//   RealVectorSpace::getRandomState / interpolateEdge[2] / pruneEdge / isEqual   src/state_spaces/real_vector_space/RealVectorSpace.cpp:27-163
//
// Execution model (this is the part that is new): a planner run is a C++20 coroutine that SUSPENDS at every state-space
// query. PlannerPool drives P independent runs (same scene; their own start/goal and random seed — the reference's
// `testing: max_num` repetitions, apps/test_rgbtconnect.cpp) in lockstep: each round it collects the pending queries of
// all runs, issues ONE backend call per query kind (validity / edges / distances / spines) and resumes the runs with
// their answers. On the GPU a round costs a few kernel launches whatever P is, so P runs finish in about the time of the
// slowest one. Inside one run the planner also batches what the reference does one call at a time: the NUM_SPINES spines
// of a bur are one spine query, and the straight-line marching of connect() is one edge query (the marching points do not
// depend on the verdicts, only the stopping index does). Between two backend calls the runs are resumed by a small pool of host
// threads (tree growth and the linear nearest-neighbour scans are the host-side cost of a round).
//
// Differences from the reference, all outside the collision/distance path: nearest neighbours are found by a linear scan
// instead of nanoflann's kd-tree (both exact); random numbers come from a counter-based generator per run (independent of P and of
// the thread schedule) unless RVS_PLAN_RNG=libc, which makes a single run draw from std::rand() exactly like Eigen::VectorXf::Random —
// that run then reproduces the reference planner's own run state by state (tests/test_planners.py against the compiled solve()).
// Robot self-collision (xArm6's capsule pre-filter and edge marching, xArm6.cpp:284-424) is part of the backend's isValid / isValidEdges.
#ifndef RVS_PLANNING_PLANNERS_H
#define RVS_PLANNING_PLANNERS_H

#include <atomic>
#include <coroutine>
#include <cstdint>
#include <exception>
#include <memory>
#include <utility>
#include <vector>

#include "SpaceBackend.h"

namespace rvsplan
{
	enum class PlannerKind : int32_t { RRTConnect = 0, RBTConnect = 1, RGBTConnect = 2 };
	enum class Status : uint8_t { None = 0, Advanced = 1, Trapped = 2, Reached = 3 };   // base::State::Status

	struct PlannerConfig          // RRTConnectConfig + RBTConnectConfig + RGBTConnectConfig (data/configurations/*.yaml)
	{
		PlannerKind kind = PlannerKind::RGBTConnect;
		int64_t max_num_iter = 1000000000;
		int64_t max_num_states = 1000000000;
		float max_planning_time = 10.f;      // [s], wall clock of the pool
		float eps_step = 0.1f;               // RRTConnectConfig::EPS_STEP
		int max_extension_steps = 50;        // RRTConnectConfig::MAX_EXTENSION_STEPS
		float delta = 3.14159f;              // RBTConnectConfig::DELTA
		int num_spines = 7;                  // RBTConnectConfig::NUM_SPINES
		SpineParams spine;                   // D_CRIT, NUM_ITER_SPINE, USE_EXPANDED_BUBBLE, NUM_LAYERS (1 for RBT-Connect)
		int connect_chunk = 0;               // > 0: that many steps of connectSpine / connectGenSpine's loop per backend query (Request::CONNECT, one
		                                     // rvs_store_connect call on the CUDA backend); 0: one spine query + one distance query per step.
		                                     // Off by default: on the shipped scenes the loop seldom makes a second step, so a chain saves no round
		                                     // and its idle launches cost (measured, profiles/r2: 1024 lockstep xArm6 runs 0.65 -> 1.14 s with 2)
		bool keep_trees = false;             // return both trees with the result (RBTConnect::outputPlannerData's `*trees[0]`, `*trees[1]`)
	};

	struct PlannerResult
	{
		bool success = false;
		int64_t iterations = 0;
		int64_t states = 0;
		float planning_time = 0.f;           // [s] from the start of the pool to the end of this run
		int64_t rounds = 0;                  // backend rounds this run took part in
		int64_t num_valid = 0, num_edges = 0, num_distances = 0, num_spines = 0;   // queries it issued
		std::vector<float> path;             // path_len x n
		std::vector<float> tree_coord[2];    // the two trees as outputPlannerData dumps them (only when PlannerConfig::keep_trees): states in tree order
		std::vector<int32_t> tree_parent[2]; // parent index inside the same tree, -1 for the root
	};

	// ---- coroutine plumbing --------------------------------------------------------------------------------------------
	template <class T>
	class Task
	{
	public:
		struct promise_type
		{
			T value{};
			std::coroutine_handle<> continuation{};
			std::exception_ptr error{};
			Task get_return_object() { return Task(std::coroutine_handle<promise_type>::from_promise(*this)); }
			std::suspend_always initial_suspend() noexcept { return {}; }
			struct Final
			{
				bool await_ready() const noexcept { return false; }
				std::coroutine_handle<> await_suspend(std::coroutine_handle<promise_type> h) const noexcept
				{
					auto c = h.promise().continuation;
					return c ? c : std::noop_coroutine();
				}
				void await_resume() const noexcept {}
			};
			Final final_suspend() noexcept { return {}; }
			void return_value(T v) { value = std::move(v); }
			void unhandled_exception() { error = std::current_exception(); }
		};
		explicit Task(std::coroutine_handle<promise_type> h_) : h(h_) {}
		Task(Task &&o) noexcept : h(std::exchange(o.h, {})) {}
		Task(const Task &) = delete;
		~Task() { if (h) h.destroy(); }
		bool await_ready() const noexcept { return false; }
		std::coroutine_handle<> await_suspend(std::coroutine_handle<> awaiting) noexcept
		{
			h.promise().continuation = awaiting;
			return h;
		}
		T await_resume()
		{
			if (h.promise().error) std::rethrow_exception(h.promise().error);
			return std::move(h.promise().value);
		}
		std::coroutine_handle<promise_type> handle() const { return h; }
	private:
		std::coroutine_handle<promise_type> h;
	};

	// ---- one pending state-space query of one run -----------------------------------------------------------------------
	struct Request
	{
		enum Kind { NONE, VALID, EDGES, DISTANCE, SPINES, CONNECT } kind = NONE;
		std::vector<float> q1, q2;           // VALID/DISTANCE: q1 = N x n;  EDGES: q1, q2 = E x n;  SPINES: q1 = targets S x n
		std::vector<int32_t> slots;          // DISTANCE: destination slots;  SPINES: root slots;  CONNECT: root slot, then one slot per step
		int steps = 0;                       // CONNECT: steps of connectGenSpine's loop asked for in this query (q1 = the target, d[0] = root d_c)
		bool gen2 = true;
		int count = 0;
		// answers
		std::vector<uint8_t> flags;          // VALID / EDGES verdicts; SPINES status
		std::vector<float> d;                // DISTANCE d_c; SPINES d_c_last; CONNECT: real d_c of the state every step reached
		std::vector<float> q_layers;         // SPINES: S x num_layers x n
		std::vector<int32_t> n_layers;       // SPINES
	};

	struct Node
	{
		int parent = -1;
		int slot = -1;                       // state-cache slot once a real distance was computed
		float d_c = -1.f;                    // State::d_c (-1 = unset)
		bool is_real = false;                // State::is_real_d_c
	};

	class Tree                               // base::Tree (src/state_spaces/Tree.cpp) without the kd-tree
	{
	public:
		explicit Tree(int n_) : n(n_) {}
		int add(const float *q, int parent);                 // upgradeTree
		int nearest(const float *q) const;                   // getNearestState (exact, linear scan)
		const float *coord(int idx) const { return &q[size_t(idx) * n]; }
		int size() const { return int(nodes.size()); }
		const std::vector<float> &coords() const { return q; }
		std::vector<Node> nodes;
	private:
		int n;
		std::vector<float> q;
	};

	class SlotAllocator                      // shared by the runs of a pool, which are resumed from several host threads
	{
	public:
		int next() { return count.fetch_add(1, std::memory_order_relaxed); }
		int used() const { return count.load(std::memory_order_relaxed); }
	private:
		std::atomic<int> count{0};
	};

	class PlannerRun;

	// Drives P runs in lockstep over one backend.
	class PlannerPool
	{
	public:
		// host_threads: threads that resume the runs between two backend calls (0 = all hardware threads, 1 = serial). The runs
		// are independent (own tree, own random stream), so this only changes the wall time, never a result.
		PlannerPool(SpaceBackend &backend_, const PlannerConfig &cfg_, int host_threads_ = 0) : backend(backend_), cfg(cfg_), host_threads(host_threads_) {}
		// q_start / q_goal: P x n. Throws std::domain_error when a start or goal is invalid (RRTConnect.cpp:13-16) unless
		// skip_invalid is set, in which case that run reports success = false with 0 iterations.
		std::vector<PlannerResult> solve(const float *q_start, const float *q_goal, const uint64_t *seeds, int P, bool skip_invalid = false);
		int64_t rounds() const { return num_rounds; }
		int64_t backendCalls() const { return num_calls; }
	private:
		SpaceBackend &backend;
		PlannerConfig cfg;
		int host_threads;
		int64_t num_rounds = 0, num_calls = 0;
	};
}

#endif
