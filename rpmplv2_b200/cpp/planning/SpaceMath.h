// SpaceMath.h — host-side arithmetic of base::RealVectorSpace that the planners of this directory share (no geometry: that is the backend's).
#ifndef RVS_PLANNING_SPACEMATH_H
#define RVS_PLANNING_SPACEMATH_H

#include <algorithm>
#include <cmath>

namespace rvsplan::math
{
	// RealVectorSpace::getNorm (:47-50): (q1 - q2).norm() of a dynamic VectorXf, summed in Eigen's linear-vectorised order (packets of
	// 4, two accumulators, (l0 + l2) + (l1 + l3), scalar tail) — the same order as coord_norm in csrc/rvs_kernels.cuh, so that the
	// host's interpolation fractions and the device's edge discretisation see the same distance
	inline float norm(const float *a, const float *b, int n)
	{
		float sq[64] = {};                                               // RVS_MAX_DOFS is 32
		for (int i = 0; i < n; i++) { const float d = a[i] - b[i]; sq[i] = d * d; }
		const int aligned = (n / 4) * 4, aligned2 = (n / 8) * 8;
		float res;
		if (aligned)
		{
			float p0[4], p1[4];
			for (int l = 0; l < 4; l++) p0[l] = sq[l];
			if (aligned > 4)
			{
				for (int l = 0; l < 4; l++) p1[l] = sq[4 + l];
				for (int k = 8; k < aligned2; k += 8)
					for (int l = 0; l < 4; l++) { p0[l] = p0[l] + sq[k + l]; p1[l] = p1[l] + sq[k + 4 + l]; }
				for (int l = 0; l < 4; l++) p0[l] = p0[l] + p1[l];
				if (aligned > aligned2)
					for (int l = 0; l < 4; l++) p0[l] = p0[l] + sq[aligned2 + l];
			}
			res = (p0[0] + p0[2]) + (p0[1] + p0[3]);
			for (int i = aligned; i < n; i++) res = res + sq[i];
		}
		else
		{
			res = sq[0];
			for (int i = 1; i < n; i++) res = res + sq[i];
		}
		return std::sqrt(res);
	}

	// RealVectorSpace::interpolateEdge2 (RealVectorSpace.cpp:95-114): true = Reached (q2 itself), false = Advanced by `step`
	inline bool interpolateEdge2(const float *q1, const float *q2, float step, int n, float *out)
	{
		const float dist = norm(q1, q2, n);
		if (step < dist)
		{
			if (dist > 0) for (int i = 0; i < n; i++) out[i] = q1[i] + (q2[i] - q1[i]) * (step / dist);
			else std::copy_n(q2, n, out);
			return false;
		}
		std::copy_n(q2, n, out);
		return true;
	}
}

#endif
