// Rrtx.cpp — see Rrtx.h. Reference citations are relative to the RPMPLv2 tree.
#include "Rrtx.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <queue>

#include "SpaceMath.h"

namespace rvsplan
{
	// ---- host index: exhaustive scans in the reference's arithmetic ------------------------------------------------------------------
	int HostNeighborIndex::nearest(const float *q_)
	{
		// nanoflann's L2_Simple_Adaptor: sequential sum of squared differences; an exhaustive scan keeps the first minimum
		int best = -1;
		float best_d = INFINITY;
		const int count = int(pts.size() / size_t(n));
		for (int k = 0; k < count; k++)
		{
			const float *p = &pts[size_t(k) * size_t(n)];
			float s = 0.f;
			for (int i = 0; i < n; i++) { const float d = q_[i] - p[i]; s += d * d; }
			if (s < best_d) { best_d = s; best = k; }
		}
		return best;
	}

	void HostNeighborIndex::radius(const float *q_, float radius_, int exclude, std::vector<int> &idx, std::vector<float> &dist)
	{
		idx.clear(); dist.clear();
		const int count = int(pts.size() / size_t(n));
		for (int k = 0; k < count; k++)
		{
			if (k == exclude) continue;
			const float d = math::norm(q_, &pts[size_t(k) * size_t(n)], n);
			if (d <= radius_) { idx.push_back(k); dist.push_back(d); }
		}
	}

	// ---- the tree ------------------------------------------------------------------------------------------------------------------
	RrtxTree::RrtxTree(SpaceBackend &backend_, NeighborIndex &index_, const RrtxConfig &cfg_) : backend(backend_), index(index_), cfg(cfg_), n(backend_.numDofs())
	{
		// RRTx.cpp:71-75 with the reference's own mix of float and double intermediates
		const size_t d = size_t(n);
		eta = float(std::sqrt(double(d)));
		const float mi = float(std::pow(2 * M_PI, double(d)));
		const float zeta = float(std::pow(M_PI, double(float(d) / 2.0f)) / double(std::tgamma(float(d) / 2.0f + 1)));
		gamma_rrt = float(2 * std::pow((1 + 1.0 / double(d)) * double(mi / zeta), double(1.0f / float(d))));
	}

	int RrtxTree::addRoot(const float *q_)
	{
		q.insert(q.end(), q_, q_ + n);
		parent.push_back(-1);
		cost.push_back(-1.f);                    // the reference never sets the root's cost: it keeps State's default (State.cpp:12), and every cost-to-come is measured from -1
		children.emplace_back();
		index.append(q_);
		return size() - 1;
	}

	float RrtxTree::rewireRadius() const
	{
		if (cfg.r_rewire > 0) return cfg.r_rewire;
		const size_t num_states = size_t(size());                                  // shrinkingBallRadius (:636-639)
		return std::min(gamma_rrt * float(std::pow(std::log(double(num_states)) / double(num_states), double(1.0f / float(size_t(n))))), eta);
	}

	float RrtxTree::dist(int a, int b) const { return math::norm(coord(a), coord(b), n); }

	// findNeighbors (:424-452): within `radius` in tree order; more than MAX_NEIGHBORS -> the nearest MAX_NEIGHBORS, ordered by distance
	void RrtxTree::findNeighbors(const float *q_new, int self, float radius, std::vector<int> &nb, std::vector<float> &nb_dist)
	{
		index.radius(q_new, radius, self, nb, nb_dist);
		if (int(nb.size()) > cfg.max_neighbors)
		{
			std::vector<int> order(nb.size());
			std::iota(order.begin(), order.end(), 0);
			std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nb_dist[size_t(a)] < nb_dist[size_t(b)]; });
			order.resize(size_t(cfg.max_neighbors));
			std::vector<int> nb2; std::vector<float> d2;
			for (int o : order) { nb2.push_back(nb[size_t(o)]); d2.push_back(nb_dist[size_t(o)]); }
			nb.swap(nb2); nb_dist.swap(d2);
		}
	}

	void RrtxTree::removeChild(int p, int child)                                   // State::removeChild: the FIRST entry only
	{
		auto &ch = children[size_t(p)];
		auto it = std::find(ch.begin(), ch.end(), child);
		if (it != ch.end()) ch.erase(it);
	}

	void RrtxTree::propagateCostChanges(int node)                                  // :531-548
	{
		std::queue<int> cost_queue;
		cost_queue.push(node);
		while (!cost_queue.empty())
		{
			const int current = cost_queue.front();
			cost_queue.pop();
			for (int child : children[size_t(current)])
			{
				const float new_cost = cost[size_t(current)] + dist(current, child);
				if (new_cost < cost[size_t(child)]) { cost[size_t(child)] = new_cost; cost_queue.push(child); }
			}
		}
	}

	int RrtxTree::grow(const float *q_rand)
	{
		stats.iterations++;
		const float inf = std::numeric_limits<float>::infinity();
		const int near = index.nearest(q_rand);
		std::vector<float> q_new(static_cast<size_t>(n), 0.f);
		math::interpolateEdge2(coord(near), q_rand, cfg.eps_step, n, q_new.data());
		const float cost_init = cost[size_t(near)] + math::norm(coord(near), q_new.data(), n);       // :172
		const float radius = rewireRadius();
		std::vector<int> nb; std::vector<float> nb_dist;
		findNeighbors(q_new.data(), -1, radius, nb, nb_dist);
		const int K = int(nb.size());

		// ---- the candidate edges of this iteration, as batched queries --------------------------------------------------------------
		// verdicts of the edges neighbour -> q_new (up: obstacle test and self-collision test apart, the latter only for robots that have
		// one) and q_new -> neighbour (down: 1 = isValid && !checkSelfCollision, 0 = blocked, 255 = not asked)
		const bool self = backend.hasSelfCollision();
		std::vector<uint8_t> up_ok(size_t(K), 0), up_self(size_t(K), 0), down(size_t(K), 255);
		std::vector<float> up_q2(self ? size_t(K) * size_t(n) : 0);
		std::vector<float> e1, e2;
		std::vector<uint8_t> flags;
		auto push = [&](const float *a, const float *b) { e1.insert(e1.end(), a, a + n); e2.insert(e2.end(), b, b + n); };
		auto ask = [&]() {                                                           // combined verdicts of (e1, e2) -> flags
			const int E = int(e1.size() / size_t(n));
			flags.assign(size_t(E), 0);
			if (E > 0) { backend.isValidEdges(e1.data(), e2.data(), E, flags.data()); stats.backend_calls++; stats.edges_checked += E; }
			e1.clear(); e2.clear();
		};
		auto askUp = [&](int first, const float *endpoint) {                         // edges neighbour_k -> endpoint for k >= first
			if (first >= K) return;
			for (int k = first; k < K; k++) push(coord(nb[size_t(k)]), endpoint);
			if (self)
			{
				backend.isValidEdgesDetailed(e1.data(), e2.data(), K - first, &up_ok[size_t(first)], &up_self[size_t(first)], &up_q2[size_t(first) * size_t(n)]);
				stats.backend_calls++; stats.edges_checked += K - first;
				e1.clear(); e2.clear();
			}
			else
			{
				ask();
				std::copy(flags.begin(), flags.end(), up_ok.begin() + first);
			}
		};
		uint8_t ext_ok = 0;
		if (cfg.one_call_per_iteration && !self)
		{
			float lb = cost_init;                                                    // lower bound of the cost q_new can end up with
			for (int k = 0; k < K; k++) lb = std::min(lb, cost[size_t(nb[size_t(k)])] + nb_dist[size_t(k)]);
			std::vector<int> want_down;
			push(coord(near), q_new.data());
			for (int k = 0; k < K; k++) push(coord(nb[size_t(k)]), q_new.data());
			for (int k = 0; k < K; k++)
				if (lb + nb_dist[size_t(k)] < cost[size_t(nb[size_t(k)])]) { want_down.push_back(k); push(q_new.data(), coord(nb[size_t(k)])); }
			ask();
			ext_ok = flags[0];
			for (int k = 0; k < K; k++) up_ok[size_t(k)] = flags[size_t(1 + k)];
			for (size_t w = 0; w < want_down.size(); w++) down[size_t(want_down[w])] = flags[size_t(1 + K) + w];
		}
		else
		{
			push(coord(near), q_new.data());
			ask();
			ext_ok = flags[0];
		}
		stats.edges_reference++;
		if (!ext_ok) { stats.trapped++; return -1; }                                 // RRTConnect::extend: Trapped
		if (!cfg.one_call_per_iteration || self) askUp(0, q_new.data());

		// ---- chooseParent (:454-491), replayed over the verdicts --------------------------------------------------------------------
		// q_new has no parent yet (extend returns a fresh state), so the running minimum starts at infinity.
		int min_parent = -1;
		float min_cost = inf;
		bool replaced = false;                 // xArm6 only: checkSelfCollision(neighbor, q_new) re-seats the LOCAL q_new on a hit (xArm6.cpp:284-352)
		std::vector<float> cur = q_new;
		for (int k = 0; k < K; k++)
		{
			const int nbk = nb[size_t(k)];
			const float pc = cost[size_t(nbk)] + (replaced ? math::norm(coord(nbk), cur.data(), n) : nb_dist[size_t(k)]);
			if (!(pc < min_cost)) continue;
			stats.edges_reference++;
			if (!up_ok[size_t(k)]) continue;                                         // isValid(neighbor, q_new) fails: checkSelfCollision is not called
			if (self && up_self[size_t(k)])
			{
				// the rest of the loop sees the last collision-free point of this edge instead of q_new: new edges, new verdicts
				cur.assign(up_q2.begin() + size_t(k) * size_t(n), up_q2.begin() + size_t(k + 1) * size_t(n));
				replaced = true;
				stats.reseated++;
				askUp(k + 1, cur.data());
				continue;
			}
			min_cost = pc; min_parent = nbk;
		}

		// ---- tree->upgradeTree(q_new, q_new->getParent()) (:181) --------------------------------------------------------------------
		const int id = size();
		q.insert(q.end(), q_new.begin(), q_new.end());
		children.emplace_back();
		index.append(q_new.data());
		if (min_parent >= 0 && !replaced)
		{
			parent.push_back(min_parent);
			cost.push_back(min_cost);
			children[size_t(min_parent)].push_back(id);                              // chooseParent: min_parent->addChild(q_new)
			children[size_t(min_parent)].push_back(id);                              // upgradeTree: q_parent->addChild(q_new) once more
			stats.reparented += min_parent != near;
		}
		else
		{
			// no valid parent edge, or the parent went to the re-seated copy of q_new: the state enters the tree as the reference leaves it
			parent.push_back(-1);
			cost.push_back(cost_init);
		}

		// ---- rewireNeighbors(q_new, neighbors) (:493-529) ---------------------------------------------------------------------------
		if (!cfg.one_call_per_iteration || self)
		{
			std::vector<int> want;
			for (int k = 0; k < K; k++)
				if (nb[size_t(k)] != parent[size_t(id)] && cost[size_t(id)] + nb_dist[size_t(k)] < cost[size_t(nb[size_t(k)])])
				{ want.push_back(k); push(q_new.data(), coord(nb[size_t(k)])); }
			ask();
			for (size_t w = 0; w < want.size(); w++) down[size_t(want[w])] = flags[w];
		}
		for (int k = 0; k < K; k++)
		{
			const int nbk = nb[size_t(k)];
			if (nbk == parent[size_t(id)]) continue;
			const float pc = cost[size_t(id)] + nb_dist[size_t(k)];
			if (!(pc < cost[size_t(nbk)])) continue;
			stats.edges_reference++;
			if (down[size_t(k)] == 255)                                              // (cannot happen: the batches above ask a superset)
			{
				push(q_new.data(), coord(nbk));
				ask();
				down[size_t(k)] = flags[0];
			}
			if (!down[size_t(k)]) continue;
			if (parent[size_t(nbk)] >= 0) removeChild(parent[size_t(nbk)], nbk);
			parent[size_t(nbk)] = id;
			cost[size_t(nbk)] = pc;
			children[size_t(id)].push_back(nbk);
			stats.rewired++;
			propagateCostChanges(nbk);
		}
		return id;
	}

	bool RrtxTree::pathInvalid(int start)
	{
		std::vector<float> e1, e2;
		for (int cur = start; parent[size_t(cur)] >= 0; cur = parent[size_t(cur)])
		{
			e1.insert(e1.end(), coord(cur), coord(cur) + n);
			e2.insert(e2.end(), coord(parent[size_t(cur)]), coord(parent[size_t(cur)]) + n);
		}
		const int E = int(e1.size() / size_t(n));
		if (E == 0) return false;
		std::vector<uint8_t> flags(size_t(E), 0);
		backend.isValidEdges(e1.data(), e2.data(), E, flags.data());
		stats.backend_calls++; stats.edges_checked += E;
		return std::find(flags.begin(), flags.end(), uint8_t(0)) != flags.end();
	}

	std::vector<int> RrtxTree::pathToRoot(int start) const
	{
		std::vector<int> path;
		if (parent[size_t(start)] < 0) return path;                                  // "No path to start"
		for (int cur = start; cur >= 0; cur = parent[size_t(cur)]) path.push_back(cur);
		return path;
	}
}
