// RealVectorSpaceCUDA.h — the `space: "RealVectorSpaceCUDA"` state space: RPMPLv2's base::RealVectorSpace with its
// collision / distance hot path served by the B200 engine (librvs_cuda.so, include/rvs_cuda.h).
//
// Drop-in contract (reference paths are relative to the RPMPLv2 tree):
//   * derives from base::RealVectorSpace (include/state_spaces/real_vector_space/RealVectorSpace.h:17-52) and overrides
//     exactly the four hot methods of base::StateSpace (include/state_spaces/StateSpace.h:50-55):
//         bool  isValid(q)                        RealVectorSpace.cpp:277-309
//         bool  isValid(q1, q2)                   RealVectorSpace.cpp:261-275
//         float computeDistance(q, compute_again) RealVectorSpace.cpp:318-392
//         float computeDistanceUnderestimation(q, nearest_points)   RealVectorSpace.cpp:402-446
//     with the same return values, the same side effects on base::State (setDistance / setDistanceProfile /
//     setIsRealDistance / setNearestPoints) and the same caching rules (Q20 of SURVEY.md §8.1). Everything else
//     (sampling, interpolation, pruning, path preprocessing) is inherited unchanged.
//   * adds the batched overloads the bur planners call for whole spine sets / edge batches:
//         isValidBatch, isValidEdges, computeDistanceBatch, computeDistanceUnderestimationBatch, extendSpinesBatch,
//         checkStatesBatch / checkEdgesBatch (validity AND robot->checkSelfCollision in one engine call), checkSelfCollisionEdges.
//   * robot chain: taken from the robot's URDF with rvs::chain_from_urdf_file (rvs_urdf.h) — the same chain the
//     reference's robot classes hand to KDL; radii, gripper length, ground_included and limits are read from the
//     robots::AbstractRobot the scenario built. Obstacles are read from env::Environment (AABB of every
//     fcl::CollisionObject, label == "ground", min_dist_tol).  Every query first compares the environment's current AABBs with the
//     table it uploaded last and re-uploads when they differ (Environment::addObject / removeObject / updateEnvironment between two
//     queries, as MotionValidity::check, DRGBT and RRTx do): a stale table would be a silent false `valid`.  refreshObstacles()
//     forces the upload.  robot->getSelfCollisionChecking() and RRTConnectConfig::EPS_STEP are mirrored the same way.
//   * dynamic scenes: attachDynamicEnvironment + checkMotion replace the K x { env->updateEnvironment; isValid } loop of
//     planning::trajectory::MotionValidity::check by one engine call (rvs_motion_validity).
//   * there is NO CPU fallback: the constructor throws std::runtime_error when the engine cannot be created.
#ifndef RPMPL_REALVECTORSPACECUDA_H
#define RPMPL_REALVECTORSPACECUDA_H

#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "RealVectorSpace.h"
#include "rvs_cuda.h"

namespace base
{
	class RealVectorSpaceCUDA : public base::RealVectorSpace
	{
	public:
		RealVectorSpaceCUDA(size_t num_dimensions_, const std::shared_ptr<robots::AbstractRobot> robot_,
							const std::shared_ptr<env::Environment> env_, const std::string &urdf_path,
							int device = 0, rvs_precision precision = RVS_PREC_REF32);
		RealVectorSpaceCUDA(size_t num_dimensions_, const std::shared_ptr<robots::AbstractRobot> robot_,
							const std::shared_ptr<env::Environment> env_, const rvs_robot_desc &chain,
							int device = 0, rvs_precision precision = RVS_PREC_REF32);
		~RealVectorSpaceCUDA() override;
		RealVectorSpaceCUDA(const RealVectorSpaceCUDA &) = delete;
		RealVectorSpaceCUDA &operator=(const RealVectorSpaceCUDA &) = delete;

		bool isValid(const std::shared_ptr<base::State> q1, const std::shared_ptr<base::State> q2) override;
		bool isValid(const std::shared_ptr<base::State> q) override;
		float computeDistance(const std::shared_ptr<base::State> q, bool compute_again = false) override;
		float computeDistanceUnderestimation(const std::shared_ptr<base::State> q,
			const std::shared_ptr<std::vector<Eigen::MatrixXf>> nearest_points) override;

		// ---- batched overloads (new call sites for RGBT / RGBMT* / DRGBT / RRTx) ----
		// valid[i] = isValid(states[i])
		std::vector<uint8_t> isValidBatch(const std::vector<std::shared_ptr<base::State>> &states);
		// valid[e] = isValid(edges[e].first, edges[e].second)
		std::vector<uint8_t> isValidEdges(const std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> &edges);
		// computeDistance for every state (same caching rule and the same State side effects as the single call).
		// with_nearest_points = false skips the per-(obstacle, link) nearest points (RVS_DIST_PROFILE): the states
		// then get d_c, the profile and is_real_d_c, and their nearest_points are left untouched.
		std::vector<float> computeDistanceBatch(const std::vector<std::shared_ptr<base::State>> &states,
												bool compute_again = false, bool with_nearest_points = true);
		// computeDistanceUnderestimation of every state w.r.t. ONE root's planes (the spine set of a bur)
		std::vector<float> computeDistanceUnderestimationBatch(const std::vector<std::shared_ptr<base::State>> &states,
			const std::shared_ptr<std::vector<Eigen::MatrixXf>> nearest_points);

		// One RBTConnect::extendSpine / RGBTConnect::extendGenSpine per target from ONE root whose real distance, profile and nearest
		// points are on the State (computeDistance(root) is called when they are not): RBTConnect.cpp:85-150, RGBTConnect.cpp:73-92,
		// xArm6 self-collision truncation included (RBTConnect.cpp:129-131, capsule stage).  Returns {status, reached state} per target;
		// a reached state carries the distance-to-planes (setDistance / setIsRealDistance(false)) like the reference's q_new.
		std::vector<std::pair<base::State::Status, std::shared_ptr<base::State>>> extendSpinesBatch(const std::shared_ptr<base::State> root,
			const std::vector<std::shared_ptr<base::State>> &targets, size_t num_layers, size_t num_iter_spine, bool use_expanded_bubble, float d_crit);
		// valid[i] = isValid(states[i]) && !robot->checkSelfCollision(states[i])                 (RRTConnect.cpp:13-15, MotionValidity.cpp:42)
		std::vector<uint8_t> checkStatesBatch(const std::vector<std::shared_ptr<base::State>> &states);
		// valid[e] = isValid(first, second) && !robot->checkSelfCollision(first, second)         (RRTConnect.cpp:100, RRTx.cpp:467-468, 507-508)
		std::vector<uint8_t> checkEdgesBatch(const std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> &edges);
		// robot->checkSelfCollision(first, second&) for every edge (xArm6.cpp:284-368, capsule stage): hit[e]; on a hit edges[e].second is
		// replaced by a new state at the last collision-free configuration, as the reference does with its q2 reference
		std::vector<uint8_t> checkSelfCollisionEdges(std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> &edges);

		// ---- dynamic environment on the device ----
		// Hands env's objects (position, velocity, max_vel, label == "dynamic_obstacle"; box side = extent of the current AABB,
		// use fcl::Box::side with the real FCL) and the workspace parameters to the engine, which from then on moves the objects
		// itself. motion_type is the scenario YAML's testing.motion_type (env::Environment keeps it private, Environment.h:62-70).
		void attachDynamicEnvironment(rvs_motion_type motion_type);
		// planning::trajectory::MotionValidity::check(q_previous, q_current) (MotionValidity.cpp:30-51): num_checks times
		// { updateEnvironment(delta_time); interpolateEdge; isValid } in ONE engine call. Afterwards env's objects sit where the
		// reference leaves them (positions / velocities mirrored back with setPosition / setVelocity); `path`, if given, receives
		// the validity points that passed, as the reference appends them. robot->checkSelfCollision stays with the caller.
		bool checkMotion(const std::shared_ptr<base::State> q_previous, const std::shared_ptr<base::State> q_current, size_t num_checks,
						 float delta_time, std::vector<std::shared_ptr<base::State>> *path = nullptr);
		// The other overload, MotionValidity::check(traj_points_current_iter) (MotionValidity.cpp:60-82; traj_interpolation "Spline" /
		// "Ruckig"): the trajectory's own points, one update of delta_time before each.
		bool checkMotion(const std::vector<Eigen::VectorXf> &traj_points, float delta_time, std::vector<std::shared_ptr<base::State>> *path = nullptr);
		// The same check for many candidate motions against ONE obstacle timeline; the environment is not moved.
		std::vector<uint8_t> checkMotionBatch(const std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> &motions,
											  size_t num_checks, float delta_time);

		// re-read env->getObjects() into the device obstacle table (24 B per object)
		void refreshObstacles();
		// push the current RealVectorSpaceConfig globals to the engine (they are mutable statics in the reference)
		void refreshConfig();
		rvs_handle handle() const { return engine; }

	private:
		void create(const rvs_robot_desc &chain, int device, rvs_precision precision);
		void check(int status, const char *what) const;
		void gather(const std::vector<std::shared_ptr<base::State>> &states, std::vector<float> &q) const;
		void packPlanes(const std::vector<Eigen::MatrixXf> &nearest_points, std::vector<float> &planes) const;
		void mirrorEnvironment();
		void syncIfStale();          // staleness guard: obstacle table + self-collision flags
		void storeDistance(const std::shared_ptr<base::State> &q, float d_c, const float *profile, const float *nearest) const;

		rvs_handle engine = nullptr;
		size_t num_links = 0;
		size_t num_objects = 0;
		std::vector<float> uploaded_aabb;          // [M][6] what the device table holds (staleness guard)
		std::vector<float> uploaded_tol;
		std::vector<uint8_t> uploaded_ground;
		bool device_moves_obstacles = false;       // after attachDynamicEnvironment the ENGINE owns the obstacle motion
		int uploaded_self_collision = -1;
		float uploaded_eps_step = 0.f;
	};
}

#endif //RPMPL_REALVECTORSPACECUDA_H
