// rvs_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT A PRODUCT PATH).
//
// A from-scratch restatement of RPMPLv2's `RealVectorSpace` collision/distance hot path in plain C++
// (no Eigen / FCL / KDL), templated on the geometry scalar T in {float, double}:
//   T = float  : reference-faithful (the reference's geometry is FP32, CollisionAndDistance.h:24-36)
//   T = double : the FP64 sibling used to check the RVS_PREC_F64 kernels.
// Forward kinematics always runs in double, as orocos-KDL does, and is narrowed to T for the skeleton
// (src/robots/xArm6.cpp:216-227).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
//
// PARITY STATUS: the reference ships no golden vectors for this path and cannot be built with its own build system here (Eigen, FCL,
// KDL, yaml-cpp absent).  Instead its own sources are compiled, unmodified and where they lie, against work-alikes of the API slices
// they use (oracle/eigen_shim, oracle/env_shim; oracle/Makefile target `ref` -> oracle/_ref/):
//   * CollisionAndDistance.cpp                      -> libcad_ref.so: every primitive, 2.4 M random + degenerate pairs, bit for bit;
//   * RealVectorSpace.cpp + State / StateSpace / Environment / Object / Box / AbstractRobot sources -> librvs_ref.so: the LOOPS
//     (isValid(q), isValid(q1,q2), computeDistance with profile / nearest points / nullptr rule, computeDistanceUnderestimation),
//     nine scenes, bit for bit (tests/test_oracle_golden.py, tests/golden/rvs_golden.npz);
//   * RBTConnect.cpp / RGBTConnect.cpp (+ RRTConnect, AbstractPlanner, PlannerInfo, PlanningTypes, Tree, the *ConnectConfig sources)
//     -> the same librvs_ref.so: extendSpine / extendGenSpine / extendGenSpine2, every layer, status and cached distance, bit for
//     bit (tests/test_spines.py, tests/golden/spine_golden.npz);
//   * Environment.cpp's motion laws                 -> libenv_ref.so (rvs_oracle_env.hpp).
// Third-party arithmetic (Eigen reduction order, KDL FK and the robot classes' computeEnclosingRadii that sit on it, FCL AABB) is
// restated from published behaviour / the robot sources and stays "parity unpinned" (SURVEY.md §8c, Q22-Q24).
//
// All `file:line` citations are relative to /root/reference/.
// Build: g++ -O3 -std=c++17 -ffp-contract=off  (the reference is an x86-64 Release build without FMA).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------------
// 3-vector with the two reduction orders Eigen 3.4 produces:
//   fixed-size Vector3f expressions reduce as  x0 + (x1 + x2)   (redux_novec_unroller halves the range)
//   dynamic-size blocks (MatrixXf::col)  reduce left to right    (x0 + x1) + x2
// ------------------------------------------------------------------------------------------------
template <class T>
struct V3 {
    T x, y, z;
    T operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    T &at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> inline V3<T> operator+(const V3<T> &a, const V3<T> &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline V3<T> operator-(const V3<T> &a, const V3<T> &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline V3<T> operator*(T s, const V3<T> &a) { return {s * a.x, s * a.y, s * a.z}; }
template <class T> inline T dotF(const V3<T> &a, const V3<T> &b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
template <class T> inline T dotD(const V3<T> &a, const V3<T> &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
template <class T> inline T sqnF(const V3<T> &a) { return dotF(a, a); }
template <class T> inline T normF(const V3<T> &a) { return std::sqrt(dotF(a, a)); }
template <class T> inline T normD(const V3<T> &a) { return std::sqrt(dotD(a, a)); }

template <class T>
struct Config {           // RealVectorSpaceConfig (src/configurations/RealVectorSpaceConfig.cpp:3-5)
    T eq_thr;             // EQUALITY_THRESHOLD
    T max_dist;           // MAX_DISTANCE
    int num_checks;       // NUM_INTERPOLATION_VALIDITY_CHECKS
};

template <class T>
struct Box6 {             // obs = (x_min, y_min, z_min, x_max, y_max, z_max) (CollisionAndDistance.cpp:3)
    T o[6];
    T lo(int k) const { return o[k]; }
    T hi(int k) const { return o[3 + k]; }
};

template <class T>
struct Near {             // {distance, nearest points or null}
    T d;
    bool has;             // false <=> the reference returns nullptr
    V3<T> P1, P2;         // P1 on the robot segment axis, P2 on the obstacle
};

// ------------------------------------------------------------------------------------------------
// distanceLineSegToLineSeg  (CollisionAndDistance.cpp:217-320), SURVEY Q8-Q10
// ------------------------------------------------------------------------------------------------
template <class T>
Near<T> seg_seg(const V3<T> &A, const V3<T> &B, const V3<T> &C, const V3<T> &D, const Config<T> &cfg) {
    Near<T> out{cfg.max_dist, true, {0, 0, 0}, {0, 0, 0}};
    const V3<T> AB = B - A, CD = D - C, DC = C - D, CA = A - C, BA = A - B;
    const T alpha1 = sqnF(AB);
    const T alpha2 = dotF(AB, CD);
    const T beta1 = dotF(DC, AB);
    const T beta2 = dotF(DC, CD);
    const T gamma1 = dotF(CA, BA);
    const T gamma2 = dotF(CA, DC);
    // no parallel guard: 0/0 -> NaN and x/0 -> +-inf both fail the open-interval test below (:229-232)
    const T s = (alpha1 * gamma2 - alpha2 * gamma1) / (alpha1 * beta2 - alpha2 * beta1);
    const T t = (gamma1 - beta1 * s) / alpha1;

    if (t > 0 && t < 1 && s > 0 && s < 1) {
        out.P1 = A + t * AB;
        out.P2 = C + s * CD;
        out.d = normD(out.P2 - out.P1);            // columns of a MatrixXf(3,2): dynamic reduction
        if (out.d < cfg.eq_thr) return {T(0), false, {0, 0, 0}, {0, 0, 0}};   // collision snap (:236-238)
        return out;
    }
    const T alpha3 = sqnF(DC);
    const T opt[4] = {dotF(CA, BA) / alpha1,           // s = 0
                      dotF(A - D, BA) / alpha1,        // s = 1
                      dotF(CA, CD) / alpha3,           // t = 0
                      dotF(B - C, CD) / alpha3};       // t = 1
    for (int i = 0; i < 4; i++) {
        V3<T> p1, p2;
        if (opt[i] < 0) {                              // (:251-268)
            if (i == 0 || i == 2) { p1 = A; p2 = C; }
            else if (i == 1)      { p1 = A; p2 = D; }
            else                  { p1 = B; p2 = C; }
        } else if (opt[i] > 1) {                       // (:269-286)
            if (i == 1 || i == 3) { p1 = B; p2 = D; }
            else if (i == 0)      { p1 = B; p2 = C; }
            else                  { p1 = A; p2 = D; }
        } else {                                       // (:287-309); NaN parameters land here
            if (i == 0)      { p1 = A + opt[i] * AB; p2 = C; }
            else if (i == 1) { p1 = A + opt[i] * AB; p2 = D; }
            else if (i == 2) { p1 = A; p2 = C + opt[i] * CD; }
            else             { p1 = B; p2 = C + opt[i] * CD; }
        }
        const T d = normD(p2 - p1);
        if (d < out.d) { out.d = d; out.P1 = p1; out.P2 = p2; }   // strict: first minimum wins (:312-316)
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// distanceLineSegToPoint  (CollisionAndDistance.cpp:323-342)
// ------------------------------------------------------------------------------------------------
template <class T>
Near<T> seg_point(const V3<T> &A, const V3<T> &B, const V3<T> &C, const Config<T> &cfg) {
    const V3<T> AB = B - A;
    const T t = dotF(C - A, AB) / sqnF(AB);
    V3<T> p1;
    if (t < 0) p1 = A;
    else if (t > 1) p1 = B;
    else p1 = A + t * AB;
    const T d = normD(C - p1);
    if (d < cfg.eq_thr) return {T(0), false, {0, 0, 0}, {0, 0, 0}};
    return {d, true, p1, C};
}

// in-plane axes of the face whose normal is axis c: (a, b) in ascending order (CollisionAndDistance.cpp:54-56)
inline void plane_axes(int c, int &a, int &b) { a = (c == 0) ? 1 : 0; b = (c == 2) ? 1 : 2; }

template <class T>
inline V3<T> lift(T pa, T pb, T value, int c) {     // get3DPoint (:150-155)
    V3<T> p;
    int a, b;
    plane_axes(c, a, b);
    p.at(a) = pa; p.at(b) = pb; p.at(c) = value;
    return p;
}

// ------------------------------------------------------------------------------------------------
// checkCases (CollisionAndDistance.cpp:113-148): distances from AB to the rectangle edges lying on
// the side where the 2-D `point` sticks out of the rectangle rec = (min_a, min_b, max_a, max_b).
// ------------------------------------------------------------------------------------------------
template <class T>
T rect_edge_cases(const V3<T> &A, const V3<T> &B, const T rec[4], T pa, T pb, T value, int c, const Config<T> &cfg) {
    T d1 = cfg.max_dist, d2 = cfg.max_dist;
    if (pa < rec[0])
        d1 = seg_seg(A, B, lift(rec[0], rec[1], value, c), lift(rec[0], rec[3], value, c), cfg).d;
    else if (pa > rec[2])
        d1 = seg_seg(A, B, lift(rec[2], rec[1], value, c), lift(rec[2], rec[3], value, c), cfg).d;

    if (d1 > 0 && pb < rec[1])
        d2 = seg_seg(A, B, lift(rec[0], rec[1], value, c), lift(rec[2], rec[1], value, c), cfg).d;
    else if (d1 > 0 && pb > rec[3])
        d2 = seg_seg(A, B, lift(rec[0], rec[3], value, c), lift(rec[2], rec[3], value, c), cfg).d;
    return std::min(d1, d2);
}

// ------------------------------------------------------------------------------------------------
// collisionCapsuleToRectangle (CollisionAndDistance.cpp:46-111), SURVEY Q5-Q7.
// `face` in 0..5 = {x_min, y_min, z_min, x_max, y_max, z_max}.
// ------------------------------------------------------------------------------------------------
template <class T>
bool capsule_hits_rectangle(const V3<T> &A, const V3<T> &B, T radius, const Box6<T> &obs, int face, const Config<T> &cfg) {
    const T value = obs.o[face];
    const int c = face > 2 ? face - 3 : face;
    int a, b;
    plane_axes(c, a, b);
    const T rec[4] = {obs.lo(a), obs.lo(b), obs.hi(a), obs.hi(b)};
    const T Aa = A[a], Ab = A[b], Ba = B[a], Bb = B[b];
    const T t = (B[c] - A[c]) != 0 ? (value - A[c]) / (B[c] - A[c]) : cfg.max_dist;
    const T M0 = Aa + t * (Ba - Aa), M1 = Ab + t * (Bb - Ab);
    const V3<T> Aproj = lift(Aa, Ab, value, c), Bproj = lift(Ba, Bb, value, c);
    auto norm2 = [](T u, T v) { return std::sqrt(u * u + v * v); };

    if (t > 0 && t < 1) {                            // AB crosses the face plane (:62-75)
        if (M0 > rec[0] - radius && M0 < rec[2] + radius && M1 > rec[1] - radius && M1 < rec[3] + radius) {
            // NB the four corner-disc clauses index `rec` as (min_a, max_a, min_b, max_b) although it is
            // laid out (min_a, min_b, max_a, max_b): reproduced verbatim (Q5, :69-72).
            if ((M0 > rec[0] - radius && M0 < rec[2] + radius && M1 > rec[1] && M1 < rec[3]) ||
                (M1 > rec[1] - radius && M1 < rec[3] + radius && M0 > rec[0] && M0 < rec[2]) ||
                (M0 < rec[0] && M1 < rec[2] && norm2(M0 - rec[0], M1 - rec[2]) < radius) ||
                (M0 < rec[0] && M1 > rec[3] && norm2(M0 - rec[0], M1 - rec[3]) < radius) ||
                (M0 > rec[1] && M1 < rec[2] && norm2(M0 - rec[1], M1 - rec[2]) < radius) ||
                (M0 > rec[1] && M1 > rec[3] && norm2(M0 - rec[1], M1 - rec[3]) < radius))
                return true;
        }
    } else if (std::min(normF(A - Aproj), normF(B - Bproj)) > radius)   // (:76-77)
        return false;

    if (radius > 0) {                                // (:80-108)
        const bool Ain = Aa > rec[0] && Aa < rec[2] && Ab > rec[1] && Ab < rec[3];
        const bool Bin = Ba > rec[0] && Ba < rec[2] && Bb > rec[1] && Bb < rec[3];
        if (Ain) {
            if (Bin) {
                if (std::min(normF(A - Aproj), normF(B - Bproj)) < radius) return true;
            } else {
                if (std::min(rect_edge_cases(A, B, rec, Ba, Bb, value, c, cfg), normF(A - Aproj)) < radius) return true;
            }
        } else if (Bin) {
            if (std::min(rect_edge_cases(A, B, rec, Aa, Ab, value, c, cfg), normF(B - Bproj)) < radius) return true;
        } else {
            if (rect_edge_cases(A, B, rec, Aa, Ab, value, c, cfg) < radius) return true;
            if (rect_edge_cases(A, B, rec, Ba, Bb, value, c, cfg) < radius) return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// collisionCapsuleToBox (CollisionAndDistance.cpp:4-42), SURVEY Q2-Q4
// ------------------------------------------------------------------------------------------------
template <class T>
bool capsule_hits_box(const V3<T> &A, const V3<T> &B, T radius, const Box6<T> &obs, const Config<T> &cfg) {
    // `radius * sqrt(3) / 3` is evaluated in double and narrowed (:7)
    const T rn = T(double(radius) * std::sqrt(3.0) / 3);
    auto inflated = [&](const V3<T> &P) {
        return P.x > obs.o[0] - rn && P.y > obs.o[1] - rn && P.z > obs.o[2] - rn &&
               P.x < obs.o[3] + rn && P.y < obs.o[4] + rn && P.z < obs.o[5] + rn;
    };
    if (inflated(A) || inflated(B)) return true;
    for (int k = 0; k < 3; k++)
        if ((A[k] < obs.lo(k) - radius && B[k] < obs.lo(k) - radius) ||
            (A[k] > obs.hi(k) + radius && B[k] > obs.hi(k) + radius))
            return false;
    // faces chosen by A's side only, x then y then z, stop at the first hit (:19-40)
    for (int k = 0; k < 3; k++) {
        bool hit = false;
        if (A[k] < obs.lo(k)) hit = capsule_hits_rectangle(A, B, radius, obs, k, cfg);
        else if (A[k] > obs.hi(k)) hit = capsule_hits_rectangle(A, B, radius, obs, k + 3, cfg);
        if (hit) return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// CapsuleToBox::compute and helpers (CollisionAndDistance.cpp:395-755), SURVEY Q11-Q16
// ------------------------------------------------------------------------------------------------
template <class T>
struct CapsuleBox {
    const V3<T> A, B;
    const T radius;
    const Box6<T> obs;
    const Config<T> &cfg;
    T d_c;
    Near<T> near;                 // near.has == false <=> nearest_pts == nullptr
    int proj[6][2] = {};          // projections(face, endpoint)
    T dist_AB[2];

    CapsuleBox(const V3<T> &A_, const V3<T> &B_, T r, const Box6<T> &o, const Config<T> &c)
        : A(A_), B(B_), radius(r), obs(o), cfg(c), d_c(c.max_dist), near{c.max_dist, true, {0, 0, 0}, {0, 0, 0}} {
        dist_AB[0] = dist_AB[1] = c.max_dist;
    }
    const V3<T> &P(int i) const { return i == 0 ? A : B; }

    // projectionLineSegOnSide (:451-475). c = constant axis; returns true when an endpoint is inside the box.
    void project_on_axis(int c) {
        int a, b;
        plane_axes(c, a, b);
        for (int i = 0; i < 2; i++) {
            const V3<T> &Q = P(i);
            if (Q[a] >= obs.lo(a) && Q[a] <= obs.hi(a) && Q[b] >= obs.lo(b) && Q[b] <= obs.hi(b)) {
                if (Q[c] > obs.lo(c) && Q[c] < obs.hi(c)) { d_c = 0; return; }
                else if (Q[c] <= obs.lo(c)) { proj[c][i] = 1; dist_AB[i] = obs.lo(c) - Q[c]; }
                else if (Q[c] >= obs.hi(c)) { proj[c + 3][i] = 1; dist_AB[i] = Q[c] - obs.hi(c); }
            }
        }
    }

    // distanceToMoreLineSegments (:576-595)
    void more_segments(const V3<T> *segs, int count) {
        for (int k = 0; k < count; k++) {
            Near<T> r = seg_seg(A, B, segs[2 * k], segs[2 * k + 1], cfg);
            if (r.d <= 0) { d_c = 0; return; }
            else if (r.d < d_c) { d_c = r.d; near.P1 = r.P1; near.P2 = r.P2; }
        }
    }

    // checkEdges + getLineSegments (:477-574). `other` is the endpoint WITHOUT a projection.
    void check_edges(const V3<T> &other, int idx) {
        static const int order[6] = {0, 3, 1, 4, 2, 5};      // x_min, x_max, y_min, y_max, z_min, z_max
        for (int f : order) {
            if (!proj[f][idx]) continue;
            if (capsule_hits_rectangle(A, B, T(0), obs, f, cfg)) { d_c = 0; return; }
            const int c = f > 2 ? f - 3 : f;
            int a, b;
            plane_axes(c, a, b);
            const T value = obs.o[f], min1 = obs.lo(a), min2 = obs.lo(b), max1 = obs.hi(a), max2 = obs.hi(b);
            V3<T> segs[4];
            int count = 0;
            if (other[a] < min1) { segs[0] = lift(min1, min2, value, c); segs[1] = lift(min1, max2, value, c); count = 1; }
            else if (other[a] > max1) { segs[0] = lift(max1, min2, value, c); segs[1] = lift(max1, max2, value, c); count = 1; }
            if (other[b] < min2) { segs[2 * count] = lift(min1, min2, value, c); segs[2 * count + 1] = lift(max1, min2, value, c); count++; }
            else if (other[b] > max2) { segs[2 * count] = lift(min1, max2, value, c); segs[2 * count + 1] = lift(max1, max2, value, c); count++; }
            // count == 0: the reference feeds an UNINITIALISED 3x2 block to seg-seg here (Q13); reachable only
            // on measure-zero boundaries. The restatement evaluates no edge candidates.
            more_segments(segs, count);
            return;
        }
    }

    V3<T> corner(int cx, int cy, int cz) const { return {obs.o[cx ? 3 : 0], obs.o[cy ? 4 : 1], obs.o[cz ? 5 : 2]}; }

    // the 4 edges of the face `c = value side` in the reference's ring order (:631-634, 711-713, 731-733)
    void face_ring(int c, int side, V3<T> *segs) const {
        int a, b;
        plane_axes(c, a, b);
        const T v = side ? obs.hi(c) : obs.lo(c);
        const V3<T> p00 = lift(obs.lo(a), obs.lo(b), v, c), p10 = lift(obs.hi(a), obs.lo(b), v, c),
                    p11 = lift(obs.hi(a), obs.hi(b), v, c), p01 = lift(obs.lo(a), obs.hi(b), v, c);
        segs[0] = p00; segs[1] = p10; segs[2] = p10; segs[3] = p11;
        segs[4] = p11; segs[5] = p01; segs[6] = p01; segs[7] = p00;
    }

    // checkOtherCases (:597-755)
    void other_cases() {
        int code[3];                                      // 0: both < min, 1: both > max, 2: mixed
        for (int k = 0; k < 3; k++)
            code[k] = (A[k] < obs.lo(k) && B[k] < obs.lo(k)) ? 0 : ((A[k] > obs.hi(k) && B[k] > obs.hi(k)) ? 1 : 2);
        const int mixed = (code[0] == 2) + (code[1] == 2) + (code[2] == 2);
        if (mixed == 0) {                                 // corner region -> segment-to-point
            Near<T> r = seg_point(A, B, corner(code[0], code[1], code[2]), cfg);
            d_c = r.d; near = r;
        } else if (mixed == 1) {                          // edge region -> one segment-to-segment
            int lo[3], hi[3];
            for (int k = 0; k < 3; k++) { lo[k] = code[k] == 2 ? 0 : code[k]; hi[k] = code[k] == 2 ? 1 : code[k]; }
            Near<T> r = seg_seg(A, B, corner(lo[0], lo[1], lo[2]), corner(hi[0], hi[1], hi[2]), cfg);
            d_c = r.d; near = r;
        } else if (mixed == 2) {                          // face region -> ring of 4 edges
            const int c = code[0] != 2 ? 0 : (code[1] != 2 ? 1 : 2);
            V3<T> segs[8];
            face_ring(c, code[c], segs);
            more_segments(segs, 4);
        } else {                                          // fully mixed (:739-751)
            for (int f = 0; f < 6; f++)
                if (capsule_hits_rectangle(A, B, T(0), obs, f, cfg)) { d_c = 0; return; }
            V3<T> segs[24];
            face_ring(2, 0, segs);
            face_ring(2, 1, segs + 8);
            const int cx[4] = {0, 1, 1, 0}, cy[4] = {0, 0, 1, 1};
            for (int k = 0; k < 4; k++) { segs[16 + 2 * k] = corner(cx[k], cy[k], 0); segs[17 + 2 * k] = corner(cx[k], cy[k], 1); }
            more_segments(segs, 12);
        }
    }

    // compute (:409-448)
    void compute() {
        project_on_axis(0);
        project_on_axis(1);
        project_on_axis(2);
        if (d_c == 0) { near.has = false; return; }       // endpoint inside the box: exactly 0 (Q11)

        int num_proj = 0;
        for (int f = 0; f < 6; f++) num_proj = std::max(num_proj, proj[f][0] + proj[f][1]);
        if (num_proj > 0) {
            const int ip = (dist_AB[0] < dist_AB[1]) ? 0 : 1;
            d_c = std::min(dist_AB[0], dist_AB[1]);
            int face = 0;                                 // maxCoeff(&idx): first face holding the maximum (:427)
            for (int f = 0; f < 6; f++) if (proj[f][ip]) { face = f; break; }
            const int c = face > 2 ? face - 3 : face;
            near.P1 = P(ip);
            near.P2 = P(ip);
            near.P2.at(c) = obs.o[face];
            if (num_proj == 1) check_edges(ip == 0 ? B : A, ip);
        } else
            other_cases();
        d_c -= radius;                                    // (:447)
    }
};

template <class T>
Near<T> capsule_box_distance(const V3<T> &A, const V3<T> &B, T radius, const Box6<T> &obs, const Config<T> &cfg) {
    CapsuleBox<T> cb(A, B, radius, obs, cfg);
    cb.compute();
    Near<T> out = cb.near;
    out.d = cb.d_c;
    return out;
}

// ------------------------------------------------------------------------------------------------
// Robots: FK (double), skeleton, enclosing radii, xArm6 capsule self-collision pre-filter
// ------------------------------------------------------------------------------------------------
enum RobotKind { XARM6 = 0, PLANAR = 1, SPATIAL = 2 };

struct M3 { double m[3][3]; };
inline M3 matmul(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = (a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j];
    return r;
}
// KDL::Rotation::RPY(roll, pitch, yaw) = Rz(yaw) Ry(pitch) Rx(roll)
inline M3 rpy_matrix(double roll, double pitch, double yaw) {
    const double ca = std::cos(yaw), sa = std::sin(yaw), cb = std::cos(pitch), sb = std::sin(pitch),
                 cc = std::cos(roll), sc = std::sin(roll);
    return {{{ca * cb, ca * sb * sc - sa * cc, ca * sb * cc + sa * sc},
             {sa * cb, sa * sb * sc + ca * cc, sa * sb * cc - ca * sc},
             {-sb, cb * sc, cb * cc}}};
}
// KDL::Rotation::Rot2(unit axis v, angle)
inline M3 axis_rotation(const double v[3], double angle) {
    const double ct = std::cos(angle), st = std::sin(angle), vt = 1 - ct;
    return {{{ct + vt * v[0] * v[0], -v[2] * st + vt * v[0] * v[1], v[1] * st + vt * v[0] * v[2]},
             {v[2] * st + vt * v[1] * v[0], ct + vt * v[1] * v[1], -v[0] * st + vt * v[1] * v[2]},
             {-v[1] * st + vt * v[2] * v[0], v[0] * st + vt * v[2] * v[1], ct + vt * v[2] * v[2]}}};
}

struct Robot {
    int kind = PLANAR;
    int n = 0;                           // DOFs == links
    int ground_included = 0;
    float gripper_length = 0;
    std::vector<M3> fixed_rot;           // RPY of joint k (n+1 entries, last = fixed tip)
    std::vector<std::array<double, 3>> xyz;   // n+1
    std::vector<std::array<double, 3>> axis;  // n (unit)
    std::vector<float> radius;           // n
    bool self_collision_checking = false;     // AbstractRobot::getSelfCollisionChecking (xArm6 ctor: true; Scenario.cpp:101-103 overrides)
    float eps_step = 0.1f;                    // RRTConnectConfig::EPS_STEP, the marching step of xArm6::checkSelfCollision(q1, q2) (xArm6.cpp:340-341)
};

struct Frame { M3 R; double p[3]; };

// KDL semantics: T_k = T_{k-1} Trans(xyz_k) RPY(rpy_k) Rot(axis_k, q_k); frames[k] = pose of
// chain segment k (k = 0..n-1 links, k = n the fixed tip).  q is float, widened to double
// (src/robots/xArm6.cpp:119-120).
inline void forward_kinematics(const Robot &rb, const float *q, Frame *frames /* n+1 */) {
    M3 R = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
    double p[3] = {0, 0, 0};
    for (int k = 0; k <= rb.n; k++) {
        for (int i = 0; i < 3; i++)
            p[i] = p[i] + ((R.m[i][0] * rb.xyz[k][0] + R.m[i][1] * rb.xyz[k][1]) + R.m[i][2] * rb.xyz[k][2]);
        R = matmul(R, rb.fixed_rot[k]);
        if (k < rb.n) R = matmul(R, axis_rotation(rb.axis[k].data(), double(q[k])));
        frames[k].R = R;
        for (int i = 0; i < 3; i++) frames[k].p[i] = p[i];
    }
}

// computeSkeleton: xArm6.cpp:209-235, Planar2DOF.cpp:117-130. skeleton has n+1 points.
template <class T>
void skeleton(const Robot &rb, const float *q, V3<T> *sk) {
    std::vector<Frame> fr(rb.n + 1);
    forward_kinematics(rb, q, fr.data());
    auto pt = [](const double *p) { return V3<T>{T(p[0]), T(p[1]), T(p[2])}; };
    if (rb.kind == XARM6) {
        const double g = double(rb.gripper_length);
        double p5[3];
        for (int i = 0; i < 3; i++) p5[i] = fr[5].p[i] + g * fr[5].R.m[i][2];            // xArm6.cpp:130
        sk[0] = {0, 0, 0};
        sk[1] = pt(fr[1].p);
        sk[2] = pt(fr[2].p);
        double p3[3], p4x[3];
        for (int i = 0; i < 3; i++) p3[i] = fr[3].p[i] - fr[3].R.m[i][2] * 0.25;         // :220
        sk[3] = pt(p3);
        sk[4] = pt(fr[4].p);
        for (int i = 0; i < 3; i++) p4x[i] = fr[4].p[i] + fr[4].R.m[i][0] * 0.076;       // :224
        sk[5] = pt(p4x);
        sk[6] = pt(p5);
        // `0.3 * gripper_length * Vector3f(a)`: the double scalar is narrowed to the vector's scalar type
        // before the product (Eigen promote_scalar_arg), the subtraction runs in T (:230-231).
        const T k = T(0.3 * double(rb.gripper_length));
        const V3<T> a = {T(fr[5].R.m[0][2]), T(fr[5].R.m[1][2]), T(fr[5].R.m[2][2])};
        sk[6] = sk[6] - k * a;
    } else {
        for (int k = 0; k <= rb.n; k++) sk[k] = pt(fr[k].p);
    }
}

// computeEnclosingRadii: xArm6.cpp:237-275 (with capsules_radius_new, :98-105), Planar2DOF.cpp:132-149,
// Spatial10DOF.cpp:11-30.  R is n x (n+1) row-major, zero-initialised.
template <class T>
void enclosing_radii(const Robot &rb, const V3<T> *sk, T *R) {
    const int n = rb.n, W = n + 1;
    std::fill(R, R + n * W, T(0));
    auto norm_xy = [](const V3<T> &p) { return std::sqrt(p.x * p.x + p.y * p.y); };
    if (rb.kind == XARM6) {
        std::vector<T> rn(n);
        for (int i = 0; i < n - 1; i++) rn[i] = std::max(T(rb.radius[i]), T(rb.radius[i + 1]));
        rn[n - 1] = T(rb.radius[n - 1]);
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j <= n; j++) {
                if (i == 0) {
                    if (j >= 2) R[i * W + j] = std::max(R[i * W + j - 1], norm_xy(sk[j]) + rn[j - 1]);
                } else if (i == 3) {
                    if (j >= 5) {
                        const V3<T> AB = sk[4] - sk[3];
                        const T t = dotD(sk[j] - sk[3], AB) / sqnF(AB);
                        const V3<T> proj = sk[3] + t * AB;
                        R[i * W + j] = std::max(R[i * W + j - 1], normD(sk[j] - proj) + rn[j - 1]);
                    }
                } else
                    R[i * W + j] = std::max(R[i * W + j - 1], normD(sk[j] - sk[i]) + rn[j - 1]);
            }
    } else {
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j <= n; j++) {
                const T len = (rb.kind == SPATIAL && i == 0) ? norm_xy(sk[j]) : normD(sk[j] - sk[i]);
                R[i * W + j] = std::max(R[i * W + j - 1], len + T(rb.radius[j - 1]));
            }
    }
}

// xArm6::checkSelfCollision(q) capsule stage (xArm6.cpp:374-424): bit k set when computeCapsulesDistance < 0
// for pairs {0-5, 1-5, 0-4, 1-4}. (The reference then confirms with FCL meshes: out of contract.)
template <class T>
uint8_t self_collision_mask(const Robot &rb, const V3<T> *sk, const Config<T> &cfg) {
    if (rb.kind != XARM6) return 0;
    static const int pairs[4][2] = {{0, 5}, {1, 5}, {0, 4}, {1, 4}};
    uint8_t mask = 0;
    for (int k = 0; k < 4; k++) {
        const int a = pairs[k][0], b = pairs[k][1];
        const T d = seg_seg(sk[a], sk[a + 1], sk[b], sk[b + 1], cfg).d - T(rb.radius[a]) - T(rb.radius[b]);
        if (d < 0) mask |= uint8_t(1u << k);
    }
    return mask;
}

// ------------------------------------------------------------------------------------------------
// Environment + state-space level
// ------------------------------------------------------------------------------------------------
template <class T>
struct Scene {
    std::vector<Box6<T>> boxes;
    std::vector<uint8_t> is_ground;      // label == "ground"
    std::vector<T> min_dist_tol;         // env::Object::getMinDistTol
};

// RealVectorSpace::isValid(q) on a skeleton (RealVectorSpace.cpp:277-309)
template <class T>
bool is_valid(const Robot &rb, const Scene<T> &sc, const V3<T> *sk, const Config<T> &cfg) {
    for (int i = 0; i < rb.n; i++)
        for (size_t j = 0; j < sc.boxes.size(); j++) {
            if (sc.is_ground[j] && i < rb.ground_included) continue;
            if (capsule_hits_box(sk[i], sk[i + 1], T(rb.radius[i]), sc.boxes[j], cfg)) return false;
        }
    return true;
}

// interpolation used by isValid(q1,q2): RealVectorSpace.cpp:75-88 and :261-275, all in float.
inline void edge_point(const float *q1, const float *q2, int n, float k, int num_checks, float dist, float *out) {
    const float step = k / num_checks * dist;
    if (dist > 0) {
        const float f = step / dist;
        for (int i = 0; i < n; i++) out[i] = q1[i] + (q2[i] - q1[i]) * f;
    } else
        for (int i = 0; i < n; i++) out[i] = q2[i];
}
// (q1 - q2).norm() on dynamic VectorXf (RealVectorSpace.cpp:49-52).  Eigen's SSE2 linear-vectorised sum
// (Redux.h, packet size 4, two packet accumulators, predux = (l0+l2)+(l1+l3), scalar tail) is emulated so
// that the step fraction matches to the last bit (assumption, see Q24).
inline float coord_norm(const float *a, const float *b, int n) {
    float sq[64];
    for (int i = 0; i < n; i++) { const float d = a[i] - b[i]; sq[i] = d * d; }
    const int aligned = (n / 4) * 4, aligned2 = (n / 8) * 8;
    float res;
    if (aligned) {
        float p0[4], p1[4];
        for (int l = 0; l < 4; l++) p0[l] = sq[l];
        if (aligned > 4) {
            for (int l = 0; l < 4; l++) p1[l] = sq[4 + l];
            for (int idx = 8; idx < aligned2; idx += 8)
                for (int l = 0; l < 4; l++) { p0[l] = p0[l] + sq[idx + l]; p1[l] = p1[l] + sq[idx + 4 + l]; }
            for (int l = 0; l < 4; l++) p0[l] = p0[l] + p1[l];
            if (aligned > aligned2)
                for (int l = 0; l < 4; l++) p0[l] = p0[l] + sq[aligned2 + l];
        }
        res = (p0[0] + p0[2]) + (p0[1] + p0[3]);
        for (int i = aligned; i < n; i++) res = res + sq[i];
    } else {
        res = sq[0];
        for (int i = 1; i < n; i++) res = res + sq[i];
    }
    return std::sqrt(res);
}

template <class T>
bool is_valid_edge(const Robot &rb, const Scene<T> &sc, const float *q1, const float *q2, const Config<T> &cfg) {
    const float dist = coord_norm(q1, q2, rb.n);
    std::vector<float> q(rb.n);
    std::vector<V3<T>> sk(rb.n + 1);
    for (float k = float(cfg.num_checks); k >= 1; k--) {
        edge_point(q1, q2, rb.n, k, cfg.num_checks, dist, q.data());
        skeleton(rb, q.data(), sk.data());
        if (!is_valid(rb, sc, sk.data(), cfg)) return false;
    }
    return true;
}

// RealVectorSpace::computeDistance (RealVectorSpace.cpp:318-392).
// profile[L]; nearest (optional) [M][L][6] = (R, O). Returns d_c. `collided` reports the early exit, after
// which `nearest` is unspecified and profile follows the reference fill (later links stay 0).
template <class T>
T distance(const Robot &rb, const Scene<T> &sc, const V3<T> *sk, const Config<T> &cfg, T *profile, T *nearest, bool *collided,
           uint64_t *visited = nullptr) {
    const int L = rb.n;
    const size_t M = sc.boxes.size();
    T d_c = cfg.max_dist;
    for (int i = 0; i < L; i++) profile[i] = 0;
    if (collided) *collided = false;
    for (int i = 0; i < L; i++) {
        profile[i] = cfg.max_dist;
        for (size_t j = 0; j < M; j++) {
            T d;
            V3<T> Rp{0, 0, 0}, Op{0, 0, 0};
            if (sc.is_ground[j] && i < rb.ground_included) {          // (:336-341)
                d = cfg.max_dist;
                Op = {0, 0, -cfg.max_dist};
            } else {
                Near<T> r = capsule_box_distance(sk[i], sk[i + 1], T(rb.radius[i]), sc.boxes[j], cfg);
                d = r.d; Rp = r.P1; Op = r.P2;
                if (visited) ++*visited;                              // one distanceCapsuleToBox call
            }
            if (d > sc.min_dist_tol[j]) d = cfg.max_dist;            // (:367-368)
            profile[i] = std::min(profile[i], d);
            if (profile[i] <= 0) {                                   // (:371-378)
                if (collided) *collided = true;
                return 0;
            }
            if (nearest) {
                T *dst = nearest + (j * L + i) * 6;
                dst[0] = Rp.x; dst[1] = Rp.y; dst[2] = Rp.z; dst[3] = Op.x; dst[4] = Op.y; dst[5] = Op.z;
            }
        }
        d_c = std::min(d_c, profile[i]);
    }
    return d_c;
}

// RealVectorSpace::computeDistanceUnderestimation (RealVectorSpace.cpp:402-446) without the State cache.
// planes [M][L][6]. Returns d_c; *rejected = true when a plane distance < 0 (the reference returns 0 and
// leaves the state untouched, :428-429) — profile is then unspecified.
template <class T>
T underestimation(const Robot &rb, size_t M, const V3<T> *sk, const T *planes, const Config<T> &cfg, T *profile, bool *rejected) {
    const int L = rb.n;
    T d_c = cfg.max_dist;
    *rejected = false;
    for (int i = 0; i < L; i++) profile[i] = 0;
    for (int i = 0; i < L; i++) {
        profile[i] = cfg.max_dist;
        for (size_t j = 0; j < M; j++) {
            const T *src = planes + (j * L + i) * 6;
            const V3<T> Rp{src[0], src[1], src[2]}, Op{src[3], src[4], src[5]};
            if (normF(Op) == cfg.max_dist) continue;                 // (:420-422)
            V3<T> nrm = Rp - Op;                                     // .normalized(): (:425-426)
            const T z = sqnF(nrm);
            if (z > 0) { const T len = std::sqrt(z); nrm = {nrm.x / len, nrm.y / len, nrm.z / len}; }
            const T d = std::min(dotD(sk[i] - Op, nrm), dotD(sk[i + 1] - Op, nrm)) - T(rb.radius[i]);
            if (d < 0) { *rejected = true; return 0; }
            profile[i] = std::min(profile[i], d);
        }
        d_c = std::min(d_c, profile[i]);
    }
    return d_c;
}

// ------------------------------------------------------------------------------------------------
// Generalized-bur spines (row f-1): RBTConnect::extendSpine (src/planners/rbt/RBTConnect.cpp:85-150) and
// RGBTConnect::extendGenSpine with use_real_dist = false (src/planners/rbt/RGBTConnect.cpp:73-92).
// Self-collision truncation (RBTConnect.cpp:129-131) is not restated: it is `false` for every robot but
// xArm6, whose check ends in an FCL mesh test.
// ------------------------------------------------------------------------------------------------
struct SpineConfig {      // RBTConnectConfig / RGBTConnectConfig
    int num_layers = 5, num_iter_spine = 1;
    bool use_expanded_bubble = true;
    float d_crit = 0.05f;
};

// R.col(c).dot(delta) on dynamic Eigen objects: same SSE2 linear-vectorised order as coord_norm (assumption Q24).
inline float dot_dyn(const float *a, int stride, const float *b, int n) {
    float pr[64];
    for (int i = 0; i < n; i++) pr[i] = a[i * stride] * b[i];
    const int aligned = (n / 4) * 4, aligned2 = (n / 8) * 8;
    float res;
    if (aligned) {
        float p0[4], p1[4];
        for (int l = 0; l < 4; l++) p0[l] = pr[l];
        if (aligned > 4) {
            for (int l = 0; l < 4; l++) p1[l] = pr[4 + l];
            for (int idx = 8; idx < aligned2; idx += 8)
                for (int l = 0; l < 4; l++) { p0[l] = p0[l] + pr[idx + l]; p1[l] = p1[l] + pr[idx + 4 + l]; }
            for (int l = 0; l < 4; l++) p0[l] = p0[l] + p1[l];
            if (aligned > aligned2)
                for (int l = 0; l < 4; l++) p0[l] = p0[l] + pr[aligned2 + l];
        }
        res = (p0[0] + p0[2]) + (p0[1] + p0[3]);
        for (int i = aligned; i < n; i++) res = res + pr[i];
    } else {
        res = pr[0];
        for (int i = 1; i < n; i++) res = res + pr[i];
    }
    return res;
}

// xArm6::computeCapsulesDistance (xArm6.cpp:417-424) on a skeleton
inline float capsules_distance(const Robot &rb, const V3<float> *sk, int a, int b, const Config<float> &cfg) {
    return seg_seg(sk[a], sk[a + 1], sk[b], sk[b + 1], cfg).d - rb.radius[a] - rb.radius[b];
}

// xArm6::checkSelfCollision(q1, q2&) (xArm6.cpp:284-368), capsule stage (the FCL mesh confirmation of checkRealSelfCollision is
// third-party arithmetic and out of contract: it is taken to answer "collide", which makes the check conservative).
//   1. up to 5 bubble steps from q1 towards q2: step = min over the pairs {0-4, 0-5, 1-4, 1-5} of capsule distance / covered distance
//      (phi_k = R.col(k) . |q2 - q1_temp| from the enclosing radii); step > 1 => no self-collision possible;
//   2. from the configuration the bubble steps reached, march towards q2 in EPS_STEP pieces and test the pairs in the order
//      {0-5, 1-5, 0-4, 1-4}, skipping a pair while its own bubble step still covers the marched fraction; on a hit q2 is replaced by
//      the last collision-free configuration and the function returns true.
// std::min / std::max are the reference's ((b < a) ? b : a and (a < b) ? b : a): NaN / inf from zero-length motions propagate as there.
// Every other robot class answers the constant false (Planar2DOF.cpp:151-161, Spatial10DOF.cpp:32-43).
inline bool self_collision_edge(const Robot &rb, const Config<float> &cfg, const float *q1, float *q2) {
    if (rb.kind != XARM6 || !rb.self_collision_checking) return false;
    const int n = rb.n, W = n + 1;
    std::vector<V3<float>> sk(n + 1);
    std::vector<float> R(size_t(n) * W), delta(n), q1_temp(q1, q1 + n), q_new(n);
    float step04 = 0, step05 = 0, step14 = 0, step15 = 0;
    int num_iter = 0;
    while (true) {
        skeleton(rb, q1_temp.data(), sk.data());
        enclosing_radii(rb, sk.data(), R.data());
        for (int i = 0; i < n; i++) delta[i] = std::fabs(q2[i] - q1_temp[i]);
        const float phi2 = dot_dyn(R.data() + 2, W, delta.data(), n), phi4 = dot_dyn(R.data() + 4, W, delta.data(), n),
                    phi5 = dot_dyn(R.data() + 5, W, delta.data(), n), phi6 = dot_dyn(R.data() + 6, W, delta.data(), n);
        step04 = capsules_distance(rb, sk.data(), 0, 4, cfg) / std::max(phi4, phi5);
        step05 = capsules_distance(rb, sk.data(), 0, 5, cfg) / std::max(phi5, phi6);
        step14 = capsules_distance(rb, sk.data(), 1, 4, cfg) / (phi2 + std::max(phi4, phi5));
        step15 = capsules_distance(rb, sk.data(), 1, 5, cfg) / (phi2 + std::max(phi5, phi6));
        const float step = std::min(std::min(step04, step05), std::min(step14, step15));
        if (step > 1) return false;
        if (++num_iter == 5 || step <= 0) break;
        if (step > 0)
            for (int i = 0; i < n; i++) q1_temp[i] = q1_temp[i] + step * (q2[i] - q1_temp[i]);
    }
    const float D = coord_norm(q2, q1_temp.data(), n);
    float dist = D;
    bool advanced = true;
    while (advanced) {
        if (dist > rb.eps_step)
            for (int i = 0; i < n; i++) q_new[i] = q1_temp[i] + (q2[i] - q1_temp[i]) / dist * rb.eps_step;
        else {
            for (int i = 0; i < n; i++) q_new[i] = q2[i];
            advanced = false;
        }
        dist = coord_norm(q2, q_new.data(), n);
        const float step_crit = 1 - dist / D;
        const bool skip[4] = {step05 > step_crit, step15 > step_crit, step04 > step_crit, step14 > step_crit};
        static const int pairs[4][2] = {{0, 5}, {1, 5}, {0, 4}, {1, 4}};
        skeleton(rb, q_new.data(), sk.data());
        for (int k = 0; k < 4; k++)
            if (!skip[k] && capsules_distance(rb, sk.data(), pairs[k][0], pairs[k][1], cfg) < 0) {
                for (int i = 0; i < n; i++) q2[i] = q1_temp[i];
                return true;
            }
        q1_temp = q_new;
    }
    return false;
}

// One extendSpine call from state (q, profile, d_c) towards q_e. Returns base::State::Status (1 Advanced, 3 Reached).
inline int extend_spine(const Robot &rb, const Config<float> &cfg, const float *q, const float *profile, float d_c, const float *q_e, const SpineConfig &sp, float *q_new) {
    const int n = rb.n, W = n + 1;
    std::vector<V3<float>> sk(n + 1), sk_new(n + 1);
    std::vector<float> R(size_t(n) * W), delta(n), rho_profile(n, 0.f), q_temp(q, q + n);
    float rho = 0.f;
    int status = 1;
    skeleton(rb, q, sk.data());
    for (int counter = 1;; counter++) {
        skeleton(rb, q_temp.data(), sk_new.data());
        enclosing_radii(rb, sk_new.data(), R.data());
        for (int i = 0; i < n; i++) delta[i] = std::fabs(q_e[i] - q_temp[i]);
        float step;
        if (sp.use_expanded_bubble) {
            step = INFINITY;
            for (int i = 0; i < n; i++) step = std::min(step, (profile[i] - rho_profile[i]) / dot_dyn(R.data() + (i + 1), W, delta.data(), n));
        } else
            step = (d_c - rho) / dot_dyn(R.data() + n, W, delta.data(), n);
        if (step > 1) {
            for (int i = 0; i < n; i++) q_new[i] = q_e[i];
            status = 3;
        } else
            for (int i = 0; i < n; i++) q_new[i] = q_temp[i] + step * (q_e[i] - q_temp[i]);
        // RBTConnect.cpp:129-131: a self-collision on [q_temp, q_new] truncates q_new to the last collision-free configuration
        const bool self_collision = self_collision_edge(rb, cfg, q_temp.data(), q_new);
        if (self_collision) status = !(coord_norm(q_temp.data(), q_new, n) < cfg.eq_thr) ? 1 : 2;
        if (counter == sp.num_iter_spine || status != 1 || self_collision) return status;
        skeleton(rb, q_new, sk_new.data());
        std::fill(rho_profile.begin(), rho_profile.end(), 0.f);
        float rho_k_prev = 0.f;
        for (int k = 1; k <= n; k++) {
            const float rho_k = normD(sk[k] - sk_new[k]);
            rho_profile[k - 1] = std::max(rho_k_prev, rho_k);
            rho_k_prev = rho_k;
            rho = std::max(rho, rho_k);
        }
        for (int i = 0; i < n; i++) q_temp[i] = q_new[i];
    }
}

// extendGenSpine from a root (its profile, d_c and nearest-point planes) towards q_e. q_out[n]; *d_out = the last
// distance-to-planes (0 when it was rejected). Returns the status (2 Trapped when the root has no clearance).
inline int extend_gen_spine(const Robot &rb, size_t M, const Config<float> &cfg, const SpineConfig &sp, const float *q_root, const float *root_profile,
                            float root_d_c, const float *root_planes, const float *q_e, float *q_out, float *d_out) {
    const int n = rb.n;
    for (int i = 0; i < n; i++) q_out[i] = q_root[i];
    *d_out = root_d_c;
    if (!(root_d_c > 0)) return 2;
    std::vector<float> q_cur(q_root, q_root + n), prof_cur(root_profile, root_profile + n), prof_new(n);
    std::vector<V3<float>> sk(n + 1);
    float dc_cur = root_d_c;
    int status = 0;
    for (int layer = 0; layer < sp.num_layers; layer++) {
        status = extend_spine(rb, cfg, q_cur.data(), prof_cur.data(), dc_cur, q_e, sp, q_out);
        skeleton(rb, q_out, sk.data());
        bool rejected;
        *d_out = underestimation(rb, M, sk.data(), root_planes, cfg, prof_new.data(), &rejected);
        if (*d_out < sp.d_crit || status != 1) break;
        q_cur.assign(q_out, q_out + n);
        prof_cur = prof_new;
        dc_cur = *d_out;
    }
    return status;
}

// extendGenSpine (gen2 = false, RGBTConnect.cpp:73-92) / extendGenSpine2 (gen2 = true, :96-118) keeping the state after
// every layer: q_layers[num_layers][n], *n_layers = how many were made (0 and status Trapped when the root has no
// clearance; the reference's extendGenSpine2 is never called for such a root, RGBTConnect.cpp:33,141).
inline int extend_gen_spine_layers(const Robot &rb, size_t M, const Config<float> &cfg, const SpineConfig &sp, bool gen2, const float *q_root,
                                   const float *root_profile, float root_d_c, const float *root_planes, const float *q_e, float *q_layers,
                                   int *n_layers, float *d_out) {
    const int n = rb.n;
    *n_layers = 0;
    *d_out = root_d_c;
    if (!(root_d_c > 0)) { for (int i = 0; i < n; i++) q_layers[i] = q_root[i]; return 2; }
    std::vector<float> q_cur(q_root, q_root + n), prof_cur(root_profile, root_profile + n), prof_new(n);
    std::vector<V3<float>> sk(n + 1);
    float dc_cur = root_d_c;
    int status = 0;
    for (int layer = 0; layer < sp.num_layers; layer++) {
        float *q_out = q_layers + size_t(layer) * n;
        status = extend_spine(rb, cfg, q_cur.data(), prof_cur.data(), dc_cur, q_e, sp, q_out);
        *n_layers = layer + 1;
        skeleton(rb, q_out, sk.data());
        bool rejected;
        *d_out = underestimation(rb, M, sk.data(), root_planes, cfg, prof_new.data(), &rejected);
        if (*d_out < sp.d_crit || (gen2 ? status == 3 : status != 1)) break;
        q_cur.assign(q_out, q_out + n);
        prof_cur = prof_new;
        dc_cur = *d_out;
    }
    return status;
}

// splitmix64-based stateless synthetic configuration generator
inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline float unit24(uint64_t seed, uint64_t counter) { return float(splitmix64(seed ^ counter) >> 40) * (1.0f / 16777216.0f); }
inline void generate_config(uint64_t seed, int64_t i, int n, const float *lo, const float *hi, float *q) {
    for (int j = 0; j < n; j++) q[j] = lo[j] + (hi[j] - lo[j]) * unit24(seed, uint64_t(i) * 32u + uint64_t(j));
}

}  // namespace orc
