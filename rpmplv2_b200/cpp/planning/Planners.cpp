// Planners.cpp — see Planners.h. Reference citations are relative to the RPMPLv2 tree.
#include "Planners.h"
#include "SpaceMath.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <stdexcept>
#include <string>
#include <thread>

namespace rvsplan
{
	namespace
	{
		using Clock = std::chrono::steady_clock;
		using Coord = std::vector<float>;

		inline uint64_t splitmix64(uint64_t x)
		{
			x += 0x9E3779B97F4A7C15ull;
			x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
			x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
			return x ^ (x >> 31);
		}

		// RVS_PLAN_RNG=libc: draw from std::rand() exactly as Eigen::VectorXf::Random does (one call per coefficient, in index order),
		// so that a SINGLE run (P = 1, after srand(seed)) consumes the reference's own random stream and can be compared with
		// RRTConnect/RBTConnect/RGBTConnect::solve() state by state (tests/test_planners.py).  Not meaningful for P > 1.
		inline bool libcStream()
		{
			const char *e = std::getenv("RVS_PLAN_RNG");          // read per run: a test switches it on for single runs only
			return e && std::string(e) == "libc";
		}

		struct Rng                       // counter based: value k of run `seed` is a pure function of (seed, k)
		{
			uint64_t seed = 0, counter = 0;
			bool libc = false;
			float uniform()              // [-1, 1) like Eigen::VectorXf::Random
			{
				if (libc) return 2.0f * float(std::rand()) / float(RAND_MAX) - 1.0f;
				const uint64_t h = splitmix64(seed ^ (0xD1B54A32D192ED03ull * ++counter));
				return float(h >> 40) * (2.0f / 16777216.0f) - 1.0f;
			}
		};

		using math::norm;                // RealVectorSpace::getNorm in Eigen's summation order (SpaceMath.h)
	}

	// ---- Tree ------------------------------------------------------------------------------------------------------------
	int Tree::add(const float *coord_, int parent)
	{
		q.insert(q.end(), coord_, coord_ + n);
		Node nd;
		nd.parent = parent;
		nodes.push_back(nd);
		return int(nodes.size()) - 1;
	}

	int Tree::nearest(const float *coord_) const
	{
		int best = -1;
		float best_d = INFINITY;
		const int count = int(nodes.size());
		for (int k = 0; k < count; k++)
		{
			const float *p = &q[size_t(k) * n];
			float s = 0.f;
			for (int i = 0; i < n; i++) { const float d = p[i] - coord_[i]; s += d * d; }
			if (s < best_d) { best_d = s; best = k; }
		}
		return best;
	}

	// ---- one planner run ---------------------------------------------------------------------------------------------------
	class PlannerRun
	{
	public:
		PlannerRun(const PlannerConfig &cfg_, const SpaceBackend &space, SlotAllocator &slots_, const float *q_start_, const float *q_goal_,
				   uint64_t seed, Clock::time_point t0_, bool skip_invalid_)
			: cfg(cfg_), n(space.numDofs()), lo(space.limitsLo()), hi(space.limitsHi()), eq_thr(space.equalityThreshold()), slots(slots_),
			  trees{Tree(space.numDofs()), Tree(space.numDofs())}, q_start(q_start_, q_start_ + space.numDofs()),
			  q_goal(q_goal_, q_goal_ + space.numDofs()), t0(t0_), skip_invalid(skip_invalid_)
		{
			rng.seed = splitmix64(seed);
			if (libcStream()) { rng.libc = true; std::srand(unsigned(seed)); }
			if (cfg.kind == PlannerKind::RBTConnect) cfg.spine.num_layers = 1;     // a bur spine is a one-layer generalized spine
			if (const char *e = std::getenv("RVS_PLAN_CONNECT_CHUNK")) cfg.connect_chunk = std::max(0, std::min(64, std::atoi(e)));   // A/B knob (0: a query per step)
		}

		Task<int> solve()
		{
			// RRTConnect ctor (RRTConnect.cpp:13-25): both ends must be valid
			req.kind = Request::VALID; req.count = 2;
			req.q1 = q_start; req.q1.insert(req.q1.end(), q_goal.begin(), q_goal.end());
			co_await Query{*this};
			res.num_valid += 2;
			if (!req.flags[0] || !req.flags[1])
			{
				invalid_ends = req.flags[0] ? "Goal position is invalid!" : "Start position is invalid!";
				finish(false);
				co_return 0;
			}
			trees[0].add(q_start.data(), -1);
			trees[1].add(q_goal.data(), -1);
			res.states = 2;
			int t = 0;                   // 0: tree from q_start, 1: tree from q_goal
			Coord q_e(n), q_new(n);
			while (true)
			{
				Status status = Status::None;
				randomState(nullptr, q_e.data());
				int near = trees[t].nearest(q_e.data());
				bool bur = false;
				if (cfg.kind != PlannerKind::RRTConnect)
				{
					const float d_c = co_await computeDistance(t, near);
					bur = d_c > cfg.spine.d_crit;
				}
				if (bur)
				{
					// RBTConnect.cpp:32-40 / RGBTConnect.cpp:33-45: NUM_SPINES (generalized) spines from q_near, one spine query
					req.kind = Request::SPINES; req.count = cfg.num_spines; req.gen2 = true;
					req.q1.resize(size_t(cfg.num_spines) * n);
					req.slots.assign(size_t(cfg.num_spines), trees[t].nodes[near].slot);
					for (int s = 0; s < cfg.num_spines; s++) randomStateAround(trees[t].coord(near), &req.q1[size_t(s) * n]);
					co_await Query{*this};
					res.num_spines += cfg.num_spines;
					for (int s = 0; s < cfg.num_spines; s++)
					{
						int parent = near, last = near;
						last = addLayers(t, s, parent);
						status = Status(req.flags[s]);
						if (s == cfg.num_spines - 1) std::copy_n(trees[t].coord(last), n, q_new.begin());
					}
				}
				else
				{
					const Coord from(trees[t].coord(near), trees[t].coord(near) + n);
					status = co_await extend(from, q_e, q_new);
					if (status != Status::Trapped) trees[t].add(q_new.data(), near);
				}
				t = 1 - t;
				if (status != Status::Trapped)
				{
					near = trees[t].nearest(q_new.data());
					if (cfg.kind == PlannerKind::RRTConnect) status = co_await connect(t, near, q_new, cfg.max_extension_steps);
					else status = co_await connectGenSpine(t, near, q_new);
				}
				res.iterations++;
				res.states = trees[0].size() + trees[1].size();
				if (status == Status::Reached)
				{
					computePath();
					finish(true);
					co_return 0;
				}
				if (elapsed() >= cfg.max_planning_time || res.states >= cfg.max_num_states || res.iterations >= cfg.max_num_iter)
				{
					finish(false);
					co_return 0;
				}
			}
		}

		struct Query
		{
			PlannerRun &r;
			bool await_ready() const noexcept { return false; }
			void await_suspend(std::coroutine_handle<> h) const noexcept { r.resume_at = h; }
			void await_resume() const noexcept {}
		};

		Request req;
		std::coroutine_handle<> resume_at{};
		PlannerResult res;
		std::string invalid_ends;
		bool finished = false;

	private:
		float elapsed() const { return std::chrono::duration<float>(Clock::now() - t0).count(); }
		void finish(bool ok)
		{
			res.success = ok; res.planning_time = elapsed(); finished = true; req.kind = Request::NONE;
			if (cfg.keep_trees)
				for (int t = 0; t < 2; t++)
				{
					res.tree_coord[t] = trees[t].coords();
					res.tree_parent[t].clear();
					for (const Node &nd : trees[t].nodes) res.tree_parent[t].push_back(nd.parent);
				}
		}

		// RealVectorSpace::getRandomState (RealVectorSpace.cpp:27-41)
		void randomState(const float *center, float *out)
		{
			for (int i = 0; i < n; i++)
			{
				out[i] = ((hi[i] - lo[i]) * rng.uniform() + lo[i] + hi[i]) / 2;
				if (center) out[i] += center[i];
			}
		}

		// RBTConnect::getRandomState(q_center) (RBTConnect.cpp:69-81): saturate to DELTA, prune to the joint limits
		void randomStateAround(const float *center, float *out)
		{
			Coord tmp(n);
			do
			{
				randomState(center, tmp.data());
				const float dist = norm(center, tmp.data(), n);                                    // interpolateEdge (:78-91)
				if (dist > 0) for (int i = 0; i < n; i++) tmp[i] = center[i] + (tmp[i] - center[i]) * (cfg.delta / dist);
				pruneEdge(center, tmp.data(), out);
			} while (norm(center, out, n) < eq_thr);                                                // isEqual (:53-59)
		}

		// RealVectorSpace::pruneEdge (:118-163) with the robot's joint limits
		void pruneEdge(const float *q1, const float *q2, float *out) const
		{
			std::vector<float> bounds(n);
			std::vector<int> indices;
			for (int k = 0; k < n; k++)
			{
				if (q2[k] > hi[k]) { bounds[k] = hi[k]; indices.push_back(k); }
				else if (q2[k] < lo[k]) { bounds[k] = lo[k]; indices.push_back(k); }
			}
			for (int idx : indices)
			{
				const float t = (bounds[idx] - q1[idx]) / (q2[idx] - q1[idx]);
				bool found = true;
				for (int k = 0; k < n; k++)
				{
					out[k] = q1[k] + t * (q2[k] - q1[k]);
					if (out[k] < lo[k] - eq_thr || out[k] > hi[k] + eq_thr) { found = false; break; }
				}
				if (found)
				{
					for (int k = 0; k < n; k++) out[k] = q1[k] + t * (q2[k] - q1[k]);
					return;
				}
			}
			std::copy_n(q2, n, out);
		}

		// RealVectorSpace::interpolateEdge2 (:95-114)
		Status interpolateEdge2(const float *q1, const float *q2, float step, float *out) const
		{
			const float dist = norm(q1, q2, n);
			if (step < dist)
			{
				if (dist > 0) for (int i = 0; i < n; i++) out[i] = q1[i] + (q2[i] - q1[i]) * (step / dist);
				else std::copy_n(q2, n, out);
				return Status::Advanced;
			}
			std::copy_n(q2, n, out);
			return Status::Reached;
		}

		// computeDistance with the State cache rule (RealVectorSpace.cpp:320-321): reuse d_c when it is real and > 0
		Task<float> computeDistance(int t, int idx)
		{
			Node &nd = trees[t].nodes[idx];
			if (nd.d_c > 0 && nd.is_real) co_return nd.d_c;
			int slot = nd.slot;
			if (slot < 0) slot = slots.next();
			req.kind = Request::DISTANCE; req.count = 1;
			req.q1.assign(trees[t].coord(idx), trees[t].coord(idx) + n);
			req.slots.assign(1, slot);
			co_await Query{*this};
			res.num_distances++;
			Node &nd2 = trees[t].nodes[idx];                    // the vector may have grown meanwhile
			nd2.slot = slot; nd2.d_c = req.d[0]; nd2.is_real = true;
			co_return nd2.d_c;
		}

		// RRTConnect::extend (RRTConnect.cpp:92-106): one EPS_STEP towards q_e, kept when the edge is valid
		Task<Status> extend(Coord from, Coord q_e, Coord &q_new)
		{
			const Status status = interpolateEdge2(from.data(), q_e.data(), cfg.eps_step, q_new.data());
			req.kind = Request::EDGES; req.count = 1;
			req.q1 = from; req.q2 = q_new;
			co_await Query{*this};
			res.num_edges++;
			co_return req.flags[0] ? status : Status::Trapped;
		}

		// RRTConnect::connect (RRTConnect.cpp:108-125). The marching points q_{k+1} = interpolateEdge2(q_k, q_e, EPS_STEP) do not
		// depend on the verdicts, so all (up to `steps`) edges go out as ONE query and the longest valid prefix is kept.
		Task<Status> connect(int t, int idx, Coord q_e, int steps)
		{
			if (steps <= 0) co_return Status::Advanced;
			Coord cur(trees[t].coord(idx), trees[t].coord(idx) + n), next(n);
			req.kind = Request::EDGES; req.q1.clear(); req.q2.clear();
			std::vector<Status> st;
			for (int k = 0; k < steps; k++)
			{
				st.push_back(interpolateEdge2(cur.data(), q_e.data(), cfg.eps_step, next.data()));
				req.q1.insert(req.q1.end(), cur.begin(), cur.end());
				req.q2.insert(req.q2.end(), next.begin(), next.end());
				if (st.back() == Status::Reached) break;
				cur = next;
			}
			const int E = int(st.size());
			req.count = E;
			co_await Query{*this};
			res.num_edges += E;
			int parent = idx;
			for (int k = 0; k < E; k++)
			{
				if (!req.flags[k]) co_return Status::Trapped;
				parent = trees[t].add(&req.q2[size_t(k) * n], parent);
			}
			co_return st.back();               // Reached, or Advanced when MAX_EXTENSION_STEPS ran out
		}

		// adds the states of spine s (answer of the last SPINES query) below `parent`; returns the last node
		int addLayers(int t, int s, int parent)
		{
			const int layers = std::max(1, int(req.n_layers[s]));
			const size_t stride = size_t(cfg.spine.num_layers) * n;
			for (int l = 0; l < layers; l++) parent = trees[t].add(&req.q_layers[size_t(s) * stride + size_t(l) * n], parent);
			return parent;
		}

		// RBTConnect::connectSpine (RBTConnect.cpp:152-180) / RGBTConnect::connectGenSpine (RGBTConnect.cpp:120-150)
		// The loop { spine from q_new; computeDistance(q_new) } depends on nothing but the state space, so up to cfg.connect_chunk of
		// its steps go out as ONE query (Request::CONNECT; on the CUDA backend one engine call whose kernels gate themselves on device
		// state) and the tree is grown from the answer.
		Task<Status> connectGenSpine(int t, int idx, Coord q_e)
		{
			float d_c = co_await computeDistance(t, idx);
			int q_new = idx;
			Status status = Status::Advanced;
			int num_ext = 0;                                   // extension steps made so far (the reference's num_ext after its increment)
			while (status == Status::Advanced && num_ext < cfg.max_extension_steps)
			{
				if (!(d_c > cfg.spine.d_crit))
				{
					// below D_CRIT the loop stays in RRT mode until it ends (d_c is not refreshed, :141-146): march the rest at once
					status = co_await connect(t, q_new, q_e, cfg.max_extension_steps - num_ext);
					co_return status;
				}
				if (cfg.connect_chunk <= 0)
				{
					const int q_temp = q_new;
					req.kind = Request::SPINES; req.count = 1; req.gen2 = true;
					req.q1 = q_e;
					req.slots.assign(1, trees[t].nodes[q_temp].slot);
					co_await Query{*this};
					res.num_spines++;
					num_ext++;
					status = Status(req.flags[0]);
					q_new = addLayers(t, 0, q_temp);
					d_c = co_await computeDistance(t, q_new);
					continue;
				}
				const int steps = std::min(cfg.connect_chunk, cfg.max_extension_steps - num_ext);
				req.kind = Request::CONNECT; req.count = 1; req.steps = steps;
				req.q1 = q_e;
				req.d.assign(1, d_c);
				req.slots.assign(1, trees[t].nodes[q_new].slot);
				for (int k = 0; k < steps; k++)
				{
					if (spare_slots.empty()) req.slots.push_back(slots.next());
					else { req.slots.push_back(spare_slots.back()); spare_slots.pop_back(); }
				}
				co_await Query{*this};
				const size_t stride = size_t(cfg.spine.num_layers) * n;
				int used = 0;
				for (int k = 0; k < steps; k++)
				{
					if (req.n_layers[k] < 0) break;              // the loop had ended before this step
					used++; num_ext++;
					res.num_spines++; res.num_distances++;
					status = Status(req.flags[k]);
					const int layers = std::max(1, int(req.n_layers[k]));
					for (int l = 0; l < layers; l++) q_new = trees[t].add(&req.q_layers[size_t(k) * stride + size_t(l) * n], q_new);
					Node &nd = trees[t].nodes[q_new];
					nd.slot = req.slots[size_t(1 + k)]; nd.d_c = req.d[size_t(k)]; nd.is_real = true;
					d_c = nd.d_c;
					if (status != Status::Advanced) break;
				}
				for (int k = steps; k > used; k--) spare_slots.push_back(req.slots[size_t(k)]);     // slots the loop did not get to
				if (used == 0) break;                            // (cannot happen: d_c > D_CRIT was checked above)
			}
			co_return status;
		}

		// RRTConnect::computePath (RRTConnect.cpp:127-145)
		void computePath()
		{
			std::vector<int> chain;
			for (int k = trees[0].nodes[trees[0].size() - 1].parent; k >= 0; k = trees[0].nodes[k].parent) chain.push_back(k);
			std::reverse(chain.begin(), chain.end());
			for (int k : chain) res.path.insert(res.path.end(), trees[0].coord(k), trees[0].coord(k) + n);
			for (int k = trees[1].size() - 1; k >= 0; k = trees[1].nodes[k].parent) res.path.insert(res.path.end(), trees[1].coord(k), trees[1].coord(k) + n);
		}

		PlannerConfig cfg;
		int n;
		const float *lo, *hi;
		float eq_thr;
		SlotAllocator &slots;
		Rng rng;
		std::vector<int32_t> spare_slots;      // State slots a connect query was given but did not use
		Tree trees[2];
		Coord q_start, q_goal;
		Clock::time_point t0;
		bool skip_invalid;
	};

	// ---- the pool -----------------------------------------------------------------------------------------------------------
	namespace
	{
		// Persistent helper threads for the resume step: run(count, fn) calls fn(i) for i in [0, count) on the caller and on whichever
		// helpers are awake (dynamic chunks of 4) and returns when every item is done. The caller never waits for a helper to wake
		// up: work is claimed through one atomic word (round << 32 | next index), so a helper that arrives late finds nothing to
		// claim. Helpers spin for ~600 us after a round before sleeping on the word: rounds are separated by one backend call (150-400 us), and a
		// helper that has to be woken through the kernel arrives after the caller has done the round alone.
		class ResumeCrew
		{
		public:
			explicit ResumeCrew(int threads)
			{
				for (int t = 1; t < threads; t++) helpers.emplace_back([this] { loop(); });
			}
			~ResumeCrew()
			{
				stop.store(true);
				state.fetch_add(uint64_t(1) << 32);
				state.notify_all();
				for (auto &th : helpers) th.join();
			}
			template <class F>
			void run(int count, F &&fn)
			{
				if (helpers.empty() || count < 32) { for (int i = 0; i < count; i++) fn(i); return; }
				job = [&fn](int i) { fn(i); };
				total.store(count, std::memory_order_relaxed);
				done.store(0, std::memory_order_relaxed);
				const uint64_t round = (state.load(std::memory_order_relaxed) >> 32) + 1;
				state.store(round << 32, std::memory_order_release);
				if (sleepers.load(std::memory_order_acquire) > 0) state.notify_all();
				work(round);
				for (int spin = 0; done.load(std::memory_order_acquire) < count; spin++)
					if (spin > 4000) std::this_thread::yield();
			}
		private:
			void work(uint64_t round)
			{
				for (;;)
				{
					uint64_t cur = state.load(std::memory_order_acquire);
					const int n = total.load(std::memory_order_relaxed);
					if ((cur >> 32) != round || int(cur & 0xffffffffu) >= n) return;
					if (!state.compare_exchange_weak(cur, cur + 4, std::memory_order_acq_rel)) continue;
					const int b = int(cur & 0xffffffffu), e = std::min(n, b + 4);
					for (int i = b; i < e; i++) job(i);
					done.fetch_add(e - b, std::memory_order_release);
				}
			}
			void loop()
			{
				uint64_t seen = 0;
				for (;;)
				{
					uint64_t cur;
					const auto t_spin = std::chrono::steady_clock::now();
					for (int spin = 0;; spin++)
					{
						cur = state.load(std::memory_order_acquire);
						if ((cur >> 32) != seen) break;
						if ((spin & 255) == 255 && std::chrono::steady_clock::now() - t_spin > std::chrono::microseconds(600))
						{
							sleepers.fetch_add(1, std::memory_order_acq_rel);
							state.wait(cur);
							sleepers.fetch_sub(1, std::memory_order_acq_rel);
						}
					}
					seen = cur >> 32;
					if (stop.load()) return;
					work(seen);
				}
			}
			std::vector<std::thread> helpers;
			std::function<void(int)> job;
			std::atomic<int> total{0}, done{0}, sleepers{0};
			std::atomic<uint64_t> state{0};
			std::atomic<bool> stop{false};
		};
	}

	std::vector<PlannerResult> PlannerPool::solve(const float *q_start, const float *q_goal, const uint64_t *seeds, int P, bool skip_invalid)
	{
		const int n = backend.numDofs();
		SlotAllocator slots;
		const auto t0 = Clock::now();
		std::vector<std::unique_ptr<PlannerRun>> runs;
		std::vector<Task<int>> tasks;
		runs.reserve(size_t(P));
		tasks.reserve(size_t(P));
		for (int p = 0; p < P; p++)
		{
			runs.emplace_back(new PlannerRun(cfg, backend, slots, q_start + size_t(p) * n, q_goal + size_t(p) * n, seeds ? seeds[p] : uint64_t(p), t0, skip_invalid));
			tasks.emplace_back(runs.back()->solve());
			tasks.back().handle().resume();
		}
		int nthreads = host_threads > 0 ? host_threads : int(std::thread::hardware_concurrency());
		if (const char *env = std::getenv("RVS_PLAN_THREADS")) { nthreads = std::max(1, std::atoi(env)); host_threads = nthreads; }    // tuning / debugging knob
		ResumeCrew crew(P < 256 ? 1 : std::max(1, std::min({host_threads > 0 ? nthreads : nthreads - 2, 32, P / 16})));   // default: leave two hardware threads to the caller and the CUDA runtime
		std::vector<float> q1, q2, d, q_layers;
		std::vector<int32_t> sl, n_layers;
		std::vector<uint8_t> flags;
		std::vector<int> who;
		int slot_cap = 0;
		const size_t layer_stride = size_t(cfg.kind == PlannerKind::RBTConnect ? 1 : cfg.spine.num_layers) * n;
		SpineParams sp = cfg.spine;
		if (cfg.kind == PlannerKind::RBTConnect) sp.num_layers = 1;
		double t_gather = 0, t_backend = 0, t_resume = 0, t_connect = 0, t_kind[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		long long n_kind[8] = {0, 0, 0, 0, 0, 0, 0, 0}, items_kind[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		const bool profile = std::getenv("RVS_PLAN_PROFILE") != nullptr;
		auto now = [] { return Clock::now(); };
		auto secs = [](Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
		while (true)
		{
			bool any = false;
			for (int kind = Request::VALID; kind <= Request::CONNECT; kind++)
			{
				const auto tg0 = now();
				who.clear(); q1.clear(); q2.clear(); sl.clear();
				int total = 0;
				for (int p = 0; p < P; p++)
				{
					Request &r = runs[p]->req;
					if (runs[p]->finished || r.kind != kind) continue;
					who.push_back(p);
					total += r.count;
					q1.insert(q1.end(), r.q1.begin(), r.q1.begin() + size_t(r.count) * n);
					if (kind == Request::EDGES) q2.insert(q2.end(), r.q2.begin(), r.q2.begin() + size_t(r.count) * n);
					if (kind == Request::DISTANCE || kind == Request::SPINES) sl.insert(sl.end(), r.slots.begin(), r.slots.begin() + r.count);
				}
				if (who.empty()) continue;
				t_gather += secs(tg0, now());
				if (kind == Request::CONNECT)
				{
					// every run's chain in one backend call: R requests x the longest chain asked for (shorter ones are padded with -1)
					any = true;
					num_calls++;
					if (slots.used() > slot_cap)
					{
						slot_cap = std::max(256, 2 * slots.used());
						backend.reserveSlots(slot_cap);
					}
					const int R = int(who.size());
					int steps = 0;
					for (int p : who) steps = std::max(steps, runs[p]->req.steps);
					std::vector<int32_t> roots(static_cast<size_t>(R)), fresh(size_t(R) * steps, -1), nl(size_t(R) * steps);
					std::vector<float> root_d(static_cast<size_t>(R)), d_new(size_t(R) * steps), layers(size_t(R) * steps * layer_stride);
					std::vector<uint8_t> st(size_t(R) * steps);
					for (int i = 0; i < R; i++)
					{
						const Request &r = runs[size_t(who[size_t(i)])]->req;
						roots[size_t(i)] = r.slots[0]; root_d[size_t(i)] = r.d[0];
						for (int k = 0; k < r.steps; k++) fresh[size_t(i) * steps + k] = r.slots[size_t(1 + k)];
					}
					const auto tc0 = now();
					backend.connectGenSpines(roots.data(), root_d.data(), q1.data(), fresh.data(), R, steps, sp, layers.data(), nl.data(), d_new.data(), st.data());
					const auto tc1 = now();
					t_backend += secs(tc0, tc1); t_connect += secs(tc0, tc1);
					for (int i = 0; i < R; i++)
					{
						Request &r = runs[size_t(who[size_t(i)])]->req;
						const size_t o = size_t(i) * steps;
						r.n_layers.assign(nl.begin() + o, nl.begin() + o + r.steps);
						r.d.assign(d_new.begin() + o, d_new.begin() + o + r.steps);
						r.flags.assign(st.begin() + o, st.begin() + o + r.steps);
						r.q_layers.assign(layers.begin() + o * layer_stride, layers.begin() + (o + r.steps) * layer_stride);
						r.kind = Request::NONE;
						runs[size_t(who[size_t(i)])]->res.rounds++;
					}
					crew.run(R, [&](int i) { runs[size_t(who[size_t(i)])]->resume_at.resume(); });
					t_resume += secs(tc1, now());
					continue;
				}
				any = true;
				num_calls++;
				if (slots.used() > slot_cap)
				{
					slot_cap = std::max(256, 2 * slots.used());
					backend.reserveSlots(slot_cap);
				}
				const auto tb0 = now();
				switch (kind)
				{
					case Request::VALID:
						flags.resize(size_t(total));
						backend.isValid(q1.data(), total, flags.data());
						break;
					case Request::EDGES:
						flags.resize(size_t(total));
						backend.isValidEdges(q1.data(), q2.data(), total, flags.data());
						break;
					case Request::DISTANCE:
						d.resize(size_t(total));
						backend.distanceToSlots(q1.data(), total, sl.data(), d.data());
						break;
					default:
						flags.resize(size_t(total)); d.resize(size_t(total)); n_layers.resize(size_t(total));
						q_layers.resize(size_t(total) * layer_stride);
						backend.extendGenSpines(sl.data(), q1.data(), total, sp, true, q_layers.data(), n_layers.data(), d.data(), flags.data());
						break;
				}
				const auto tr0 = now();
				t_backend += secs(tb0, tr0); t_kind[kind] += secs(tb0, tr0); n_kind[kind]++; items_kind[kind] += total;
				size_t off = 0;
				for (int p : who)
				{
					Request &r = runs[p]->req;
					const size_t c = size_t(r.count);
					if (kind == Request::VALID || kind == Request::EDGES || kind == Request::SPINES) r.flags.assign(flags.begin() + off, flags.begin() + off + c);
					if (kind == Request::DISTANCE || kind == Request::SPINES) r.d.assign(d.begin() + off, d.begin() + off + c);
					if (kind == Request::SPINES)
					{
						r.n_layers.assign(n_layers.begin() + off, n_layers.begin() + off + c);
						r.q_layers.assign(q_layers.begin() + off * layer_stride, q_layers.begin() + (off + c) * layer_stride);
					}
					off += c;
					r.kind = Request::NONE;
					runs[p]->res.rounds++;
				}
				// resume after the scatter: a resumed run overwrites its request with the next one (possibly of a kind served later in this round)
				crew.run(int(who.size()), [&](int i) { runs[size_t(who[size_t(i)])]->resume_at.resume(); });
				t_resume += secs(tr0, now());
			}
			if (!any) break;
			num_rounds++;
		}
		if (profile) std::fprintf(stderr, "[rvsplan] P=%d rounds=%lld calls=%lld gather %.1f ms backend %.1f ms (connect chains %.1f ms) scatter+resume %.1f ms total %.1f ms\n", P, (long long)num_rounds, (long long)num_calls, 1e3 * t_gather, 1e3 * t_backend, 1e3 * t_connect, 1e3 * t_resume, 1e3 * secs(t0, now()));
		if (profile)
			for (int k = Request::VALID; k <= Request::SPINES; k++)
				std::fprintf(stderr, "[rvsplan]   kind %d (1 valid, 2 edges, 3 distance, 4 spines): %lld calls, %lld items, %.1f ms (%.1f us per call)\n", k, n_kind[k], items_kind[k], 1e3 * t_kind[k], n_kind[k] ? 1e6 * t_kind[k] / double(n_kind[k]) : 0.0);
		std::vector<PlannerResult> out;
		out.reserve(size_t(P));
		for (int p = 0; p < P; p++)
		{
			auto &pr = tasks[p].handle().promise();
			if (pr.error) std::rethrow_exception(pr.error);
			if (!runs[p]->invalid_ends.empty() && !skip_invalid) throw std::domain_error(runs[p]->invalid_ends);
			out.push_back(std::move(runs[p]->res));
		}
		return out;
	}
}
