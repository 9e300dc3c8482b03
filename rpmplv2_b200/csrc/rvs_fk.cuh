// rvs_fk.cuh — device forward kinematics of a serial revolute chain, streamed into the capsule skeleton.
// Kept free of launch-specific constructs so that tests/host_sim can also compile it for the host (logic check only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "rvs_types.h"

namespace rvs {

// ---- forward kinematics -----------------------------------------------------------------------------------------
__device__ __forceinline__ void sincos_of(double a, double *s, double *c) { sincos(a, s, c); }
__device__ __forceinline__ void sincos_of(float a, float *s, float *c) { sincosf(a, s, c); }

// Streams the chain and writes the skeleton (n+1 points) as R.  F is the FK scalar.  Component c of point k goes to
// sk[(3k + c) * stride] (stride 1: a private array; stride = block size: a shared-memory column per thread).
// Reference semantics: KDL tree FK (T_k = T_{k-1} Trans(xyz_k) RPY(rpy_k) Rot(axis_k, q_k)) and
// computeSkeleton (xArm6.cpp:209-235, Planar2DOF.cpp:117-130).
// `joint_sincos(k, &s, &c)` supplies sin / cos of joint k's angle: computed on the spot (fk_skeleton) or fetched from the lane that
// computed it (fk_skeleton_warp: the 2n transcendental evaluations are the long part of a 22-joint chain and are independent).
template <class F, class R, class SinCos>
__device__ __forceinline__ void fk_skeleton_with(const RobotDev *__restrict__ rb, SinCos joint_sincos, R *sk_base, const int stride = 1) {
    struct Strided {
        R *p; int s;
        __device__ __forceinline__ R &operator[](int k) const { return p[k * s]; }
    } sk = {sk_base, stride};
    const int n = rb->n;
    const bool xarm = rb->kind == RVS_ROBOT_XARM6;
    F r00 = 1, r01 = 0, r02 = 0, r10 = 0, r11 = 1, r12 = 0, r20 = 0, r21 = 0, r22 = 1;
    F px = 0, py = 0, pz = 0;
    if (xarm) { sk[0] = 0; sk[1] = 0; sk[2] = 0; }
#pragma unroll 1
    for (int k = 0; k <= n; k++) {
        const F x = F(rb->xyz[k][0]), y = F(rb->xyz[k][1]), z = F(rb->xyz[k][2]);
        px = px + ((r00 * x + r01 * y) + r02 * z);
        py = py + ((r10 * x + r11 * y) + r12 * z);
        pz = pz + ((r20 * x + r21 * y) + r22 * z);
        if (!rb->rfix_identity[k]) {
            const double *m = rb->rfix[k];
            const F m00 = F(m[0]), m01 = F(m[1]), m02 = F(m[2]), m10 = F(m[3]), m11 = F(m[4]), m12 = F(m[5]), m20 = F(m[6]), m21 = F(m[7]), m22 = F(m[8]);
            const F a00 = (r00 * m00 + r01 * m10) + r02 * m20, a01 = (r00 * m01 + r01 * m11) + r02 * m21, a02 = (r00 * m02 + r01 * m12) + r02 * m22;
            const F a10 = (r10 * m00 + r11 * m10) + r12 * m20, a11 = (r10 * m01 + r11 * m11) + r12 * m21, a12 = (r10 * m02 + r11 * m12) + r12 * m22;
            const F a20 = (r20 * m00 + r21 * m10) + r22 * m20, a21 = (r20 * m01 + r21 * m11) + r22 * m21, a22 = (r20 * m02 + r21 * m12) + r22 * m22;
            r00 = a00; r01 = a01; r02 = a02; r10 = a10; r11 = a11; r12 = a12; r20 = a20; r21 = a21; r22 = a22;
        }
        if (k < n) {
            F s, c;
            joint_sincos(k, &s, &c);
            const int code = rb->axis_code[k];
            if (code == AXIS_Z) {            // columns 0,1 mix; the products dropped are exact zeros
                const F a00 = r00 * c + r01 * s, a01 = r01 * c - r00 * s;
                const F a10 = r10 * c + r11 * s, a11 = r11 * c - r10 * s;
                const F a20 = r20 * c + r21 * s, a21 = r21 * c - r20 * s;
                r00 = a00; r01 = a01; r10 = a10; r11 = a11; r20 = a20; r21 = a21;
            } else if (code == AXIS_X) {
                const F a01 = r01 * c + r02 * s, a02 = r02 * c - r01 * s;
                const F a11 = r11 * c + r12 * s, a12 = r12 * c - r11 * s;
                const F a21 = r21 * c + r22 * s, a22 = r22 * c - r21 * s;
                r01 = a01; r02 = a02; r11 = a11; r12 = a12; r21 = a21; r22 = a22;
            } else if (code == AXIS_Y) {
                const F a00 = r00 * c - r02 * s, a02 = r00 * s + r02 * c;
                const F a10 = r10 * c - r12 * s, a12 = r10 * s + r12 * c;
                const F a20 = r20 * c - r22 * s, a22 = r20 * s + r22 * c;
                r00 = a00; r02 = a02; r10 = a10; r12 = a12; r20 = a20; r22 = a22;
            } else {                          // KDL::Rotation::Rot2
                const F vx = F(rb->axis[k][0]), vy = F(rb->axis[k][1]), vz = F(rb->axis[k][2]), vt = F(1) - c;
                const F m00 = c + vt * vx * vx, m01 = -vz * s + vt * vx * vy, m02 = vy * s + vt * vx * vz;
                const F m10 = vz * s + vt * vy * vx, m11 = c + vt * vy * vy, m12 = -vx * s + vt * vy * vz;
                const F m20 = -vy * s + vt * vz * vx, m21 = vx * s + vt * vz * vy, m22 = c + vt * vz * vz;
                const F a00 = (r00 * m00 + r01 * m10) + r02 * m20, a01 = (r00 * m01 + r01 * m11) + r02 * m21, a02 = (r00 * m02 + r01 * m12) + r02 * m22;
                const F a10 = (r10 * m00 + r11 * m10) + r12 * m20, a11 = (r10 * m01 + r11 * m11) + r12 * m21, a12 = (r10 * m02 + r11 * m12) + r12 * m22;
                const F a20 = (r20 * m00 + r21 * m10) + r22 * m20, a21 = (r20 * m01 + r21 * m11) + r22 * m21, a22 = (r20 * m02 + r21 * m12) + r22 * m22;
                r00 = a00; r01 = a01; r02 = a02; r10 = a10; r11 = a11; r12 = a12; r20 = a20; r21 = a21; r22 = a22;
            }
        }
        if (!xarm) {
            sk[3 * k] = R(px); sk[3 * k + 1] = R(py); sk[3 * k + 2] = R(pz);
        } else if (k == 1 || k == 2) {
            sk[3 * k] = R(px); sk[3 * k + 1] = R(py); sk[3 * k + 2] = R(pz);
        } else if (k == 3) {                                     // xArm6.cpp:220
            sk[9] = R(px - r02 * F(0.25)); sk[10] = R(py - r12 * F(0.25)); sk[11] = R(pz - r22 * F(0.25));
        } else if (k == 4) {                                     // xArm6.cpp:222-225
            sk[12] = R(px); sk[13] = R(py); sk[14] = R(pz);
            sk[15] = R(px + r00 * F(0.076)); sk[16] = R(py + r10 * F(0.076)); sk[17] = R(pz + r20 * F(0.076));
        } else if (k == 5) {                                     // xArm6.cpp:130,226-231
            const F g = F(rb->gripper_length);
            const R kk = R(0.3 * double(rb->gripper_length));
            sk[18] = R(px + g * r02) - kk * R(r02); sk[19] = R(py + g * r12) - kk * R(r12); sk[20] = R(pz + g * r22) - kk * R(r22);
        }
    }
}

template <class F, class R>
__device__ __forceinline__ void fk_skeleton(const RobotDev *__restrict__ rb, const float *__restrict__ q, R *sk_base, const int stride = 1) {
    fk_skeleton_with<F, R>(rb, [q](int k, F *s, F *c) { sincos_of(F(q[k]), s, c); }, sk_base, stride);
}

#ifdef __CUDACC__
// Warp-cooperative variant for kernels in which a whole warp works on ONE configuration (every lane holds the same q): lane k
// evaluates sincos(q_k), the chain products then run (redundantly, in registers) on every lane with the values shuffled in.
// Same operations in the same order as fk_skeleton -> bit-identical skeleton.  All 32 lanes must call it.
__device__ __forceinline__ double shfl_f(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ float shfl_f(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <class F, class R>
__device__ __forceinline__ void fk_skeleton_warp(const RobotDev *__restrict__ rb, const float *__restrict__ q, R *sk_base, const int stride = 1) {
    const int lane = threadIdx.x & 31, n = rb->n;
    F s_mine = 0, c_mine = 1;
    if (lane < n) sincos_of(F(q[lane]), &s_mine, &c_mine);
    fk_skeleton_with<F, R>(rb, [=](int k, F *s, F *c) { *s = shfl_f(s_mine, k); *c = shfl_f(c_mine, k); }, sk_base, stride);
}
#endif

}  // namespace rvs
