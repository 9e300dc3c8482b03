// RealVectorSpaceCUDA.cpp — see RealVectorSpaceCUDA.h. Host-side glue only: gathers base::State coordinates into flat
// arrays, calls the C ABI, scatters results back into the State caches exactly as RealVectorSpace.cpp does.
#include "RealVectorSpaceCUDA.h"

#include <cstring>
#include <stdexcept>

#include "RealVectorSpaceConfig.h"
#include "RRTConnectConfig.h"
#include "rvs_urdf.h"

// base::StateSpaceType::RealVectorSpaceCUDA is the enum value INTEGRATION.md adds to include/state_spaces/StateSpaceType.h.  A build
// against an UNPATCHED reference tree (oracle/Makefile: _ref/libdropin_ref.so) passes -DRVS_STATE_SPACE_TYPE=base::StateSpaceType::RealVectorSpace.
#ifndef RVS_STATE_SPACE_TYPE
#define RVS_STATE_SPACE_TYPE base::StateSpaceType::RealVectorSpaceCUDA
#endif

namespace
{
	rvs_params currentParams(rvs_precision precision)
	{
		rvs_params p;
		p.equality_threshold = RealVectorSpaceConfig::EQUALITY_THRESHOLD;
		p.max_distance = RealVectorSpaceConfig::MAX_DISTANCE;
		p.num_interpolation_validity_checks = int32_t(RealVectorSpaceConfig::NUM_INTERPOLATION_VALIDITY_CHECKS);
		p.precision = int32_t(precision);
		return p;
	}
}

base::RealVectorSpaceCUDA::RealVectorSpaceCUDA(size_t num_dimensions_, const std::shared_ptr<robots::AbstractRobot> robot_,
	const std::shared_ptr<env::Environment> env_, const std::string &urdf_path, int device, rvs_precision precision)
	: RealVectorSpace(num_dimensions_, robot_, env_)
{
	rvs_robot_desc chain;
	std::string error;
	if (!rvs::chain_from_urdf_file(urdf_path, robot->getType(), &chain, &error))
		throw std::runtime_error("RealVectorSpaceCUDA: " + error);
	create(chain, device, precision);
}

base::RealVectorSpaceCUDA::RealVectorSpaceCUDA(size_t num_dimensions_, const std::shared_ptr<robots::AbstractRobot> robot_,
	const std::shared_ptr<env::Environment> env_, const rvs_robot_desc &chain, int device, rvs_precision precision)
	: RealVectorSpace(num_dimensions_, robot_, env_)
{
	create(chain, device, precision);
}

void base::RealVectorSpaceCUDA::create(const rvs_robot_desc &chain, int device, rvs_precision precision)
{
	setStateSpaceType(RVS_STATE_SPACE_TYPE);
	if (precision == RVS_PREC_F64)
		throw std::runtime_error("RealVectorSpaceCUDA: base::State stores float distances; use RVS_PREC_REF32 or RVS_PREC_FAST32");
	if (size_t(chain.num_dofs) != robot->getNumDOFs() || size_t(chain.num_dofs) != num_dimensions)
		throw std::runtime_error("RealVectorSpaceCUDA: URDF chain and robot disagree on the number of DOFs");
	num_links = robot->getNumLinks();
	if (num_links != size_t(chain.num_dofs))
		throw std::runtime_error("RealVectorSpaceCUDA: the engine expects one capsule per joint (getNumLinks() == getNumDOFs())");

	// scenario-level fields come from the robot object the scenario configured (Scenario.cpp:36-75)
	rvs_robot_desc desc = chain;
	desc.ground_included = int32_t(robot->getGroundIncluded());
	desc.gripper_length = robot->getGripperLength();
	for (size_t i = 0; i < num_links; i++)
	{
		desc.capsule_radius[i] = robot->getCapsuleRadius(i);
		desc.limit_lo[i] = robot->getLimits()[i].first;
		desc.limit_hi[i] = robot->getLimits()[i].second;
	}

	const rvs_params params = currentParams(precision);
	const int status = rvs_create(&desc, &params, device, &engine);
	if (status != RVS_OK)
		throw std::runtime_error(std::string("RealVectorSpaceCUDA: cannot create the engine: ") + rvs_last_error(nullptr));
	refreshObstacles();
}

base::RealVectorSpaceCUDA::~RealVectorSpaceCUDA()
{
	if (engine != nullptr)
		rvs_destroy(engine);
}

void base::RealVectorSpaceCUDA::check(int status, const char *what) const
{
	if (status != RVS_OK)
		throw std::runtime_error(std::string("RealVectorSpaceCUDA::") + what + ": " + rvs_last_error(engine));
}

void base::RealVectorSpaceCUDA::refreshConfig()
{
	rvs_params params = currentParams(RVS_PREC_REF32);
	params.precision = -1;		// keep the precision the handle was created with
	check(rvs_set_params(engine, &params), "refreshConfig");
}

/// Obstacle table = what RealVectorSpace.cpp:293-295,344-346 reads per pair: the AABB of every collision object,
/// whether its label is "ground" (:285,336) and its min_dist_tol (:367).
void base::RealVectorSpaceCUDA::refreshObstacles()
{
	num_objects = env->getNumObjects();
	std::vector<float> aabb_min(3 * num_objects), aabb_max(3 * num_objects), tol(num_objects);
	std::vector<uint8_t> is_ground(num_objects);
	for (size_t j = 0; j < num_objects; j++)
	{
		if (env->getCollObject(j)->getNodeType() != fcl::GEOM_BOX)
			throw std::runtime_error("RealVectorSpaceCUDA: only box obstacles are supported (the reference's sphere branch is a TODO stub)");
		const auto &AABB { env->getCollObject(j)->getAABB() };        // fcl::AABB<float> with the real FCL
		for (int k = 0; k < 3; k++)
		{
			aabb_min[3 * j + k] = AABB.min_[k];
			aabb_max[3 * j + k] = AABB.max_[k];
		}
		is_ground[j] = env->getObject(j)->getLabel() == "ground" ? 1 : 0;
		tol[j] = env->getObject(j)->getMinDistTol();
	}
	check(rvs_set_obstacles(engine, aabb_min.data(), aabb_max.data(), is_ground.data(), tol.data(), int32_t(num_objects)), "refreshObstacles");
	uploaded_aabb.resize(6 * num_objects);
	for (size_t j = 0; j < num_objects; j++)
		for (int k = 0; k < 3; k++)
		{
			uploaded_aabb[6 * j + k] = aabb_min[3 * j + k];
			uploaded_aabb[6 * j + 3 + k] = aabb_max[3 * j + k];
		}
	uploaded_tol = tol;
	uploaded_ground = is_ground;
	device_moves_obstacles = false;
}

/// Staleness guard.  The reference's own callers move the environment between two queries (MotionValidity.cpp:37-46,
/// DRGBT.cpp:79-175, RRTx.cpp:97-175: env->updateEnvironment(dt) then ss->isValid(q)); the engine answers from a device copy, so
/// every query first compares the environment's AABBs / labels / tolerances (24 M bytes of host reads) with what was uploaded and
/// re-uploads on any difference.  The self-collision switch of the robot and RRTConnectConfig::EPS_STEP are mirrored the same way.
void base::RealVectorSpaceCUDA::syncIfStale()
{
	const int sc { robot->getSelfCollisionChecking() ? 1 : 0 };
	if (sc != uploaded_self_collision || RRTConnectConfig::EPS_STEP != uploaded_eps_step)
	{
		check(rvs_set_self_collision(engine, sc, RRTConnectConfig::EPS_STEP), "syncIfStale");
		uploaded_self_collision = sc;
		uploaded_eps_step = RRTConnectConfig::EPS_STEP;
	}
	if (device_moves_obstacles)
		return;
	bool stale { env->getNumObjects() != num_objects };
	for (size_t j = 0; j < num_objects && !stale; j++)
	{
		const auto &AABB { env->getCollObject(j)->getAABB() };
		for (int k = 0; k < 3 && !stale; k++)
			stale = AABB.min_[k] != uploaded_aabb[6 * j + k] || AABB.max_[k] != uploaded_aabb[6 * j + 3 + k];
		stale = stale || env->getObject(j)->getMinDistTol() != uploaded_tol[j] ||
				(env->getObject(j)->getLabel() == "ground" ? 1 : 0) != uploaded_ground[j];
	}
	if (stale)
		refreshObstacles();
}

/// env::Environment's dynamic state (Object.h:27-31, Environment.h:54-60) -> rvs_env_set.
void base::RealVectorSpaceCUDA::attachDynamicEnvironment(rvs_motion_type motion_type)
{
	num_objects = env->getNumObjects();
	std::vector<float> dim(3 * num_objects), pos(3 * num_objects), vel(3 * num_objects), max_vel(num_objects), tol(num_objects);
	std::vector<uint8_t> is_ground(num_objects), is_dynamic(num_objects);
	for (size_t j = 0; j < num_objects; j++)
	{
		if (env->getCollObject(j)->getNodeType() != fcl::GEOM_BOX)
			throw std::runtime_error("RealVectorSpaceCUDA: only box obstacles are supported (the reference's sphere branch is a TODO stub)");
		const auto &AABB { env->getCollObject(j)->getAABB() };        // fcl::AABB<float> with the real FCL
		const std::shared_ptr<env::Object> object { env->getObject(j) };
		for (int k = 0; k < 3; k++)
		{
			dim[3 * j + k] = AABB.max_[k] - AABB.min_[k];
			pos[3 * j + k] = object->getPosition()(k);
			vel[3 * j + k] = object->getVelocity()(k);
		}
		max_vel[j] = object->getMaxVel();
		is_ground[j] = object->getLabel() == "ground" ? 1 : 0;
		is_dynamic[j] = object->getLabel() == "dynamic_obstacle" ? 1 : 0;
		tol[j] = object->getMinDistTol();
	}
	rvs_env_params ep;
	ep.motion_type = int32_t(motion_type);
	for (int k = 0; k < 3; k++)
		ep.ws_center[k] = env->getWSCenter()(k);
	ep.ws_radius = env->getWSRadius();
	ep.base_radius = env->getBaseRadius();
	ep.robot_max_vel = env->getRobotMaxVel();
	check(rvs_env_set(engine, &ep, dim.data(), pos.data(), nullptr, vel.data(), max_vel.data(), is_dynamic.data(), is_ground.data(), tol.data(),
					  int32_t(num_objects)), "attachDynamicEnvironment");
	device_moves_obstacles = true;		// from here on the engine advances the obstacles and checkMotion mirrors them back into env
}

bool base::RealVectorSpaceCUDA::checkMotion(const std::shared_ptr<base::State> q_previous, const std::shared_ptr<base::State> q_current,
	size_t num_checks, float delta_time, std::vector<std::shared_ptr<base::State>> *path)
{
	uint8_t valid { 0 };
	int32_t first_invalid { 0 };
	check(rvs_motion_validity(engine, q_previous->getCoord().data(), q_current->getCoord().data(), 1, int32_t(num_checks), delta_time, 0,
							  &valid, &first_invalid), "checkMotion");

	mirrorEnvironment();

	if (path != nullptr)
	{
		const float dist { getNorm(q_previous, q_current) };
		const size_t passed { valid ? num_checks : size_t(first_invalid - 1) };
		for (size_t num_check = 1; num_check <= passed; num_check++)
			path->emplace_back(interpolateEdge(q_previous, q_current, dist * num_check / num_checks, dist));
	}
	return valid != 0;
}

/// planning::trajectory::MotionValidity::check(traj_points_current_iter) (MotionValidity.cpp:60-82): the trajectory's own points,
/// one environment update of delta_time before each, in ONE engine call.
bool base::RealVectorSpaceCUDA::checkMotion(const std::vector<Eigen::VectorXf> &traj_points, float delta_time,
	std::vector<std::shared_ptr<base::State>> *path)
{
	if (traj_points.empty())
		return true;
	std::vector<float> points(traj_points.size() * num_dimensions);
	for (size_t k = 0; k < traj_points.size(); k++)
		std::memcpy(&points[k * num_dimensions], traj_points[k].data(), num_dimensions * sizeof(float));
	uint8_t valid { 0 };
	int32_t first_invalid { 0 };
	check(rvs_motion_validity_points(engine, points.data(), 1, int32_t(traj_points.size()), delta_time, 0, &valid, &first_invalid), "checkMotion");
	mirrorEnvironment();

	if (path != nullptr)
	{
		const size_t passed { valid ? traj_points.size() : size_t(first_invalid - 1) };
		for (size_t k = 0; k < passed; k++)
			path->emplace_back(getNewState(traj_points[k]));
	}
	return valid != 0;
}

/// the objects are where the reference's loop would have left them: mirror them back for logging / the next planner step
void base::RealVectorSpaceCUDA::mirrorEnvironment()
{
	std::vector<float> pos(3 * num_objects), vel(3 * num_objects);
	check(rvs_env_get(engine, pos.data(), vel.data()), "checkMotion");
	for (size_t j = 0; j < num_objects; j++)
	{
		if (env->getObject(j)->getLabel() != "dynamic_obstacle")
			continue;
		env->getObject(j)->setPosition(fcl::Vector3f(pos[3 * j], pos[3 * j + 1], pos[3 * j + 2]));
		env->getObject(j)->setVelocity(fcl::Vector3f(vel[3 * j], vel[3 * j + 1], vel[3 * j + 2]));
	}
}

std::vector<uint8_t> base::RealVectorSpaceCUDA::checkMotionBatch
	(const std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> &motions, size_t num_checks, float delta_time)
{
	std::vector<float> q1(motions.size() * num_dimensions), q2(motions.size() * num_dimensions);
	for (size_t e = 0; e < motions.size(); e++)
	{
		std::memcpy(&q1[e * num_dimensions], motions[e].first->getCoord().data(), num_dimensions * sizeof(float));
		std::memcpy(&q2[e * num_dimensions], motions[e].second->getCoord().data(), num_dimensions * sizeof(float));
	}
	std::vector<uint8_t> valid(motions.size());
	check(rvs_motion_validity(engine, q1.data(), q2.data(), int64_t(motions.size()), int32_t(num_checks), delta_time, -1, valid.data(), nullptr),
		  "checkMotionBatch");
	return valid;
}

void base::RealVectorSpaceCUDA::gather(const std::vector<std::shared_ptr<base::State>> &states, std::vector<float> &q) const
{
	q.resize(states.size() * num_dimensions);
	for (size_t i = 0; i < states.size(); i++)
		std::memcpy(&q[i * num_dimensions], states[i]->getCoord().data(), num_dimensions * sizeof(float));
}

/// nearest_points: one 6 x L matrix per obstacle, column i = (R, O) of link i (RealVectorSpace.cpp:381) -> [M][L][6]
void base::RealVectorSpaceCUDA::packPlanes(const std::vector<Eigen::MatrixXf> &nearest_points, std::vector<float> &planes) const
{
	planes.resize(nearest_points.size() * num_links * 6);
	for (size_t j = 0; j < nearest_points.size(); j++)
		for (size_t i = 0; i < num_links; i++)
			for (size_t k = 0; k < 6; k++)
				planes[(j * num_links + i) * 6 + k] = nearest_points[j](Eigen::Index(k), Eigen::Index(i));
}

/// The State side effects of RealVectorSpace::computeDistance (RealVectorSpace.cpp:371-376, 386-389).
void base::RealVectorSpaceCUDA::storeDistance(const std::shared_ptr<base::State> &q, float d_c, const float *profile, const float *nearest) const
{
	q->setDistance(d_c);
	q->setDistanceProfile(std::vector<float>(profile, profile + num_links));
	q->setIsRealDistance(true);
	if (d_c <= 0)
	{
		q->setNearestPoints(nullptr);
		return;
	}
	if (nearest == nullptr)
	{
		q->setNearestPoints(nullptr);		// planes were not asked for: never leave those of an older environment / configuration behind
		return;
	}
	std::shared_ptr<std::vector<Eigen::MatrixXf>> nearest_points { std::make_shared<std::vector<Eigen::MatrixXf>>
		(num_objects, Eigen::MatrixXf(6, num_links)) };
	for (size_t j = 0; j < num_objects; j++)
		for (size_t i = 0; i < num_links; i++)
			for (size_t k = 0; k < 6; k++)
				nearest_points->at(j)(Eigen::Index(k), Eigen::Index(i)) = nearest[(j * num_links + i) * 6 + k];
	q->setNearestPoints(nearest_points);
}

bool base::RealVectorSpaceCUDA::isValid(const std::shared_ptr<base::State> q)
{
	syncIfStale();
	uint8_t valid = 0;
	check(rvs_is_valid(engine, q->getCoord().data(), 1, &valid), "isValid");
	return valid != 0;
}

bool base::RealVectorSpaceCUDA::isValid(const std::shared_ptr<base::State> q1, const std::shared_ptr<base::State> q2)
{
	syncIfStale();
	uint8_t valid = 0;
	check(rvs_is_valid_edges(engine, q1->getCoord().data(), q2->getCoord().data(), 1, &valid), "isValid(q1, q2)");
	return valid != 0;
}

float base::RealVectorSpaceCUDA::computeDistance(const std::shared_ptr<base::State> q, bool compute_again)
{
	// (a state filled by computeDistanceBatch(with_nearest_points = false) has a real distance but no planes: the next single call
	// computes them — a state the reference itself filled always has both, so the reference's cache rule is unchanged)
	if (!compute_again && q->getDistance() > 0 && q->getIsRealDistance() && q->getNearestPoints() != nullptr)
		return q->getDistance();

	syncIfStale();
	float d_c = 0;
	std::vector<float> profile(num_links), nearest(num_objects * num_links * 6);
	check(rvs_distance(engine, q->getCoord().data(), 1, RVS_DIST_FULL, &d_c, profile.data(), nearest.data()), "computeDistance");
	storeDistance(q, d_c, profile.data(), nearest.data());
	return d_c;
}

float base::RealVectorSpaceCUDA::computeDistanceUnderestimation(const std::shared_ptr<base::State> q,
	const std::shared_ptr<std::vector<Eigen::MatrixXf>> nearest_points)
{
	if (q->getDistance() > 0 && q->getIsRealDistance()) 	// Real distance was already computed
		return q->getDistance();

	if (nearest_points == nullptr)
		throw std::runtime_error("RealVectorSpaceCUDA::computeDistanceUnderestimation: nearest_points is null (root state without planes)");
	syncIfStale();
	std::vector<float> planes;
	packPlanes(*nearest_points, planes);
	float d_c = 0;
	uint8_t rejected = 0;
	std::vector<float> profile(num_links);
	check(rvs_underestimation(engine, q->getCoord().data(), 1, planes.data(), nullptr, 1, &d_c, profile.data(), &rejected), "computeDistanceUnderestimation");
	if (rejected)
		return 0;			// a plane distance < 0: the reference returns 0 without touching the state (:428-429)

	q->setDistance(d_c);
	q->setDistanceProfile(profile);
	q->setIsRealDistance(false);
	return d_c;
}

std::vector<uint8_t> base::RealVectorSpaceCUDA::isValidBatch(const std::vector<std::shared_ptr<base::State>> &states)
{
	syncIfStale();
	std::vector<float> q;
	gather(states, q);
	std::vector<uint8_t> valid(states.size());
	check(rvs_is_valid(engine, q.data(), int64_t(states.size()), valid.data()), "isValidBatch");
	return valid;
}

std::vector<uint8_t> base::RealVectorSpaceCUDA::isValidEdges
	(const std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> &edges)
{
	std::vector<float> q1(edges.size() * num_dimensions), q2(edges.size() * num_dimensions);
	for (size_t e = 0; e < edges.size(); e++)
	{
		std::memcpy(&q1[e * num_dimensions], edges[e].first->getCoord().data(), num_dimensions * sizeof(float));
		std::memcpy(&q2[e * num_dimensions], edges[e].second->getCoord().data(), num_dimensions * sizeof(float));
	}
	std::vector<uint8_t> valid(edges.size());
	syncIfStale();
	check(rvs_is_valid_edges(engine, q1.data(), q2.data(), int64_t(edges.size()), valid.data()), "isValidEdges");
	return valid;
}

std::vector<uint8_t> base::RealVectorSpaceCUDA::checkStatesBatch(const std::vector<std::shared_ptr<base::State>> &states)
{
	syncIfStale();
	std::vector<float> q;
	gather(states, q);
	std::vector<uint8_t> valid(states.size());
	check(rvs_check_states(engine, q.data(), int64_t(states.size()), valid.data()), "checkStatesBatch");
	return valid;
}

std::vector<uint8_t> base::RealVectorSpaceCUDA::checkEdgesBatch
	(const std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> &edges)
{
	syncIfStale();
	std::vector<float> q1(edges.size() * num_dimensions), q2(edges.size() * num_dimensions);
	for (size_t e = 0; e < edges.size(); e++)
	{
		std::memcpy(&q1[e * num_dimensions], edges[e].first->getCoord().data(), num_dimensions * sizeof(float));
		std::memcpy(&q2[e * num_dimensions], edges[e].second->getCoord().data(), num_dimensions * sizeof(float));
	}
	std::vector<uint8_t> valid(edges.size());
	check(rvs_check_edges(engine, q1.data(), q2.data(), int64_t(edges.size()), valid.data()), "checkEdgesBatch");
	return valid;
}

std::vector<uint8_t> base::RealVectorSpaceCUDA::checkSelfCollisionEdges
	(std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> &edges)
{
	syncIfStale();
	std::vector<float> q1(edges.size() * num_dimensions), q2(edges.size() * num_dimensions), q2_out(edges.size() * num_dimensions);
	for (size_t e = 0; e < edges.size(); e++)
	{
		std::memcpy(&q1[e * num_dimensions], edges[e].first->getCoord().data(), num_dimensions * sizeof(float));
		std::memcpy(&q2[e * num_dimensions], edges[e].second->getCoord().data(), num_dimensions * sizeof(float));
	}
	std::vector<uint8_t> hit(edges.size());
	check(rvs_self_collision_edges(engine, q1.data(), q2.data(), int64_t(edges.size()), hit.data(), q2_out.data()), "checkSelfCollisionEdges");
	for (size_t e = 0; e < edges.size(); e++)
	{
		if (!hit[e])
			continue;
		Eigen::VectorXf coord(num_dimensions);
		for (size_t i = 0; i < num_dimensions; i++)
			coord(i) = q2_out[e * num_dimensions + i];
		edges[e].second = getNewState(coord);			// xArm6.cpp:355: q2 = make_shared<RealVectorSpaceState>(q1_temp_coord)
	}
	return hit;
}

std::vector<std::pair<base::State::Status, std::shared_ptr<base::State>>> base::RealVectorSpaceCUDA::extendSpinesBatch
	(const std::shared_ptr<base::State> root, const std::vector<std::shared_ptr<base::State>> &targets, size_t num_layers, size_t num_iter_spine,
	 bool use_expanded_bubble, float d_crit)
{
	std::vector<std::pair<base::State::Status, std::shared_ptr<base::State>>> result;
	if (targets.empty())
		return result;
	if (root->getDistance() < 0 || root->getNearestPoints() == nullptr)		// RBTConnect.cpp:88-89
		computeDistance(root, root->getDistance() > 0);
	if (root->getDistance() <= 0)											// RBTConnect.cpp:91-92
	{
		for (size_t s = 0; s < targets.size(); s++)
			result.emplace_back(base::State::Status::Trapped, root);
		return result;
	}
	syncIfStale();
	std::vector<float> q_e, planes;
	gather(targets, q_e);
	packPlanes(*root->getNearestPoints(), planes);
	rvs_spine_params sp;
	sp.num_layers = int32_t(num_layers);
	sp.num_iter_spine = int32_t(num_iter_spine);
	sp.use_expanded_bubble = use_expanded_bubble ? 1 : 0;
	sp.d_crit = d_crit;
	const size_t S { targets.size() };
	std::vector<float> q_reached(S * num_dimensions), d_c(S);
	std::vector<uint8_t> status(S);
	check(rvs_extend_spines(engine, root->getCoord().data(), root->getDistanceProfile().data(), root->getDistance(), planes.data(), q_e.data(),
							int32_t(S), &sp, q_reached.data(), d_c.data(), status.data()), "extendSpinesBatch");
	for (size_t s = 0; s < S; s++)
	{
		Eigen::VectorXf coord(num_dimensions);
		for (size_t i = 0; i < num_dimensions; i++)
			coord(i) = q_reached[s * num_dimensions + i];
		std::shared_ptr<base::State> q_new { getNewState(coord) };
		if (d_c[s] > 0)					// a rejected underestimation returns 0 and leaves the state untouched (RealVectorSpace.cpp:428-429)
		{
			q_new->setDistance(d_c[s]);
			q_new->setIsRealDistance(false);
		}
		result.emplace_back(static_cast<base::State::Status>(status[s]), q_new);
	}
	return result;
}

std::vector<float> base::RealVectorSpaceCUDA::computeDistanceBatch(const std::vector<std::shared_ptr<base::State>> &states,
	bool compute_again, bool with_nearest_points)
{
	std::vector<float> result(states.size());
	std::vector<std::shared_ptr<base::State>> todo;
	std::vector<size_t> where;
	for (size_t i = 0; i < states.size(); i++)
	{
		if (!compute_again && states[i]->getDistance() > 0 && states[i]->getIsRealDistance())
			result[i] = states[i]->getDistance();		// cached (RealVectorSpace.cpp:320-321)
		else
		{
			todo.emplace_back(states[i]);
			where.emplace_back(i);
		}
	}
	if (todo.empty())
		return result;

	syncIfStale();
	std::vector<float> q;
	gather(todo, q);
	const size_t per_state = num_objects * num_links * 6;
	std::vector<float> d_c(todo.size()), profile(todo.size() * num_links), nearest(with_nearest_points ? todo.size() * per_state : 0);
	check(rvs_distance(engine, q.data(), int64_t(todo.size()), with_nearest_points ? RVS_DIST_FULL : RVS_DIST_PROFILE,
					   d_c.data(), profile.data(), with_nearest_points ? nearest.data() : nullptr), "computeDistanceBatch");
	for (size_t k = 0; k < todo.size(); k++)
	{
		storeDistance(todo[k], d_c[k], &profile[k * num_links], with_nearest_points ? &nearest[k * per_state] : nullptr);
		result[where[k]] = d_c[k];
	}
	return result;
}

std::vector<float> base::RealVectorSpaceCUDA::computeDistanceUnderestimationBatch(const std::vector<std::shared_ptr<base::State>> &states,
	const std::shared_ptr<std::vector<Eigen::MatrixXf>> nearest_points)
{
	std::vector<float> result(states.size());
	std::vector<std::shared_ptr<base::State>> todo;
	std::vector<size_t> where;
	for (size_t i = 0; i < states.size(); i++)
	{
		if (states[i]->getDistance() > 0 && states[i]->getIsRealDistance())
			result[i] = states[i]->getDistance();
		else
		{
			todo.emplace_back(states[i]);
			where.emplace_back(i);
		}
	}
	if (todo.empty())
		return result;

	if (nearest_points == nullptr)
		throw std::runtime_error("RealVectorSpaceCUDA::computeDistanceUnderestimationBatch: nearest_points is null (root state without planes)");
	syncIfStale();
	std::vector<float> q, planes;
	gather(todo, q);
	packPlanes(*nearest_points, planes);
	std::vector<float> d_c(todo.size()), profile(todo.size() * num_links);
	std::vector<uint8_t> rejected(todo.size());
	check(rvs_underestimation(engine, q.data(), int64_t(todo.size()), planes.data(), nullptr, 1, d_c.data(), profile.data(), rejected.data()),
		  "computeDistanceUnderestimationBatch");
	for (size_t k = 0; k < todo.size(); k++)
	{
		if (rejected[k])
		{
			result[where[k]] = 0;
			continue;
		}
		todo[k]->setDistance(d_c[k]);
		todo[k]->setDistanceProfile(std::vector<float>(&profile[k * num_links], &profile[(k + 1) * num_links]));
		todo[k]->setIsRealDistance(false);
		result[where[k]] = d_c[k];
	}
	return result;
}
