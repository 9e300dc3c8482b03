// rvs_geom.cuh — device-side capsule / axis-aligned-box geometry of the RealVectorSpaceCUDA engine (sm_100a).
//
// What it computes is what base::CollisionAndDistance computes in the reference
// (/root/reference/src/state_spaces/real_vector_space/CollisionAndDistance.cpp); how it computes it is
// re-designed for SIMT execution:
//   * every obstacle is an AABB, so every "other segment" the reference feeds to its generic segment-segment
//     routine is an axis-aligned box edge.  seg_edge() below evaluates that routine with the structural zeros
//     removed (the products it drops are exact +-0 terms, so the surviving arithmetic is bit-identical in the
//     no-FMA REF32/F64 builds);
//   * the reference's 27-region case tree is turned into "classify -> build a list of box edges (6-bit corner
//     pairs) -> one uniform loop over seg_edge()", so that lanes of a warp that sit in different regions still
//     execute the same instruction stream and differ only in trip count;
//   * reductions keep the summation orders the oracle documents (x0+(x1+x2) for fixed 3-vectors, left-to-right
//     for matrix columns) so that REF32 results can be compared bit for bit.
//
// Template parameter R is the geometry scalar (float: REF32/FAST32, double: F64).  PTS selects whether nearest
// points are produced (RVS_DIST_FULL) — without it all point bookkeeping is compiled out.
#pragma once
#include <cuda_runtime.h>

namespace rvs {

template <class R> struct V3 { R x, y, z; };

template <class R> __device__ __forceinline__ R sel3(const V3<R> &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
template <class R> __device__ __forceinline__ void put3(V3<R> &v, int k, R s) { if (k == 0) v.x = s; else if (k == 1) v.y = s; else v.z = s; }
template <class R> __device__ __forceinline__ V3<R> sub(const V3<R> &a, const V3<R> &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class R> __device__ __forceinline__ V3<R> lerp_from(const V3<R> &a, R t, const V3<R> &d) { return {a.x + t * d.x, a.y + t * d.y, a.z + t * d.z}; }
// fixed-size reduction order  x0 + (x1 + x2)
template <class R> __device__ __forceinline__ R dotF(const V3<R> &a, const V3<R> &b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
// dynamic (matrix column) reduction order  (x0 + x1) + x2
template <class R> __device__ __forceinline__ R sqD(R dx, R dy, R dz) { return (dx * dx + dy * dy) + dz * dz; }
__device__ __forceinline__ float rsqrt_of(float x) { return sqrtf(x); }
__device__ __forceinline__ double rsqrt_of(double x) { return sqrt(x); }

template <class R>
struct Box {
    V3<R> lo, hi;
    __device__ __forceinline__ V3<R> corner(int code) const {
        return {(code & 1) ? hi.x : lo.x, (code & 2) ? hi.y : lo.y, (code & 4) ? hi.z : lo.z};
    }
};

// Per-link quantities hoisted out of the obstacle loop.
template <class R>
struct Capsule {
    V3<R> A, B, AB, BA;   // AB = B - A, BA = A - B
    R alpha1;             // |AB|^2 in fixed order
    R radius;
    R radius_in;          // radius * sqrt(3) / 3 (collision predicate only)
    __device__ __forceinline__ void set(const V3<R> &a, const V3<R> &b, R r, R rin) {
        A = a; B = b; AB = sub(b, a); BA = sub(a, b); alpha1 = dotF(AB, AB); radius = r; radius_in = rin;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Segment AB against the box edge C -> D, where D equals C except along axis e (D_e = De).
// Restates distanceLineSegToLineSeg (CollisionAndDistance.cpp:217-320) for an axis-aligned second segment.
// Returns the distance; 0 means "collision snap" (interior solution closer than EQUALITY_THRESHOLD, or an exact
// touch).  When PTS, P1/P2 receive the nearest points (unspecified when 0 is returned).
// ---------------------------------------------------------------------------------------------------------------
template <class R, bool PTS>
__device__ __forceinline__ R seg_edge(const Capsule<R> &cp, const V3<R> &C, int e, R De, R eq_thr, R max_dist, V3<R> &P1, V3<R> &P2) {
    const V3<R> &A = cp.A, &B = cp.B, &AB = cp.AB, &BA = cp.BA;
    const R alpha1 = cp.alpha1;
    const V3<R> CA = sub(A, C);
    const R Ce = sel3(C, e);
    const R len = De - Ce;           // (D - C)_e ; the other two components are exactly 0
    const R nlen = Ce - De;          // (C - D)_e
    const R ABe = sel3(AB, e), BAe = sel3(BA, e), CAe = sel3(CA, e), Ae = sel3(A, e), Be = sel3(B, e);
    const R px = CA.x * BA.x, py = CA.y * BA.y, pz = CA.z * BA.z;
    const R alpha2 = ABe * len;
    const R beta1 = nlen * ABe;
    const R beta2 = nlen * len;
    const R gamma1 = px + (py + pz);
    const R gamma2 = CAe * nlen;
    // no parallel guard on purpose: 0/0 -> NaN and x/0 -> +-inf both fail the open-interval test (Q8)
    const R s = (alpha1 * gamma2 - alpha2 * gamma1) / (alpha1 * beta2 - alpha2 * beta1);
    const R t = (gamma1 - beta1 * s) / alpha1;

    if (t > R(0) && t < R(1) && s > R(0) && s < R(1)) {
        const V3<R> p1 = lerp_from(A, t, AB);
        V3<R> p2 = C;
        put3(p2, e, Ce + s * len);
        const R d = rsqrt_of(sqD(p2.x - p1.x, p2.y - p1.y, p2.z - p1.z));
        if (d < eq_thr) return R(0);
        if (PTS) { P1 = p1; P2 = p2; }
        return d;
    }
    const R alpha3 = nlen * nlen;
    const R opt0 = gamma1 / alpha1;                                       // s = 0 (corner C)
    const R pe1 = (Ae - De) * BAe;                                        // e-th product of (A - D).(A - B)
    const R qx = e == 0 ? pe1 : px, qy = e == 1 ? pe1 : py, qz = e == 2 ? pe1 : pz;
    const R opt1 = (qx + (qy + qz)) / alpha1;                             // s = 1 (corner D)
    const R opt2 = (CAe * len) / alpha3;                                  // t = 0 (endpoint A)
    const R opt3 = ((Be - Ce) * len) / alpha3;                            // t = 1 (endpoint B)

    // candidate i = 0, 1: a point of AB against corner C / D
    const V3<R> r0 = opt0 < R(0) ? A : (opt0 > R(1) ? B : lerp_from(A, opt0, AB));
    const V3<R> r1 = opt1 < R(0) ? A : (opt1 > R(1) ? B : lerp_from(A, opt1, AB));
    // candidate i = 2, 3: endpoint A / B against a point of the edge
    const R e2 = opt2 < R(0) ? Ce : (opt2 > R(1) ? De : Ce + opt2 * len);
    const R e3 = opt3 < R(0) ? Ce : (opt3 > R(1) ? De : Ce + opt3 * len);
    V3<R> D = C;  put3(D, e, De);
    V3<R> c2 = C; put3(c2, e, e2);
    V3<R> c3 = C; put3(c3, e, e3);
    const R s0 = sqD(C.x - r0.x, C.y - r0.y, C.z - r0.z);
    const R s1 = sqD(D.x - r1.x, D.y - r1.y, D.z - r1.z);
    const R s2 = sqD(c2.x - A.x, c2.y - A.y, c2.z - A.z);
    const R s3 = sqD(c3.x - B.x, c3.y - B.y, c3.z - B.z);
    if (!PTS) {
        // sqrt is monotone, so min_i sqrt(s_i) == sqrt(min_i s_i) bit for bit; NaN candidates never win.
        R m = s0;
        if (!(m == m)) m = s1; else if (s1 < m) m = s1;
        if (!(m == m)) m = s2; else if (s2 < m) m = s2;
        if (!(m == m)) m = s3; else if (s3 < m) m = s3;
        const R d = rsqrt_of(m);
        return d < max_dist ? d : max_dist;
    } else {
        R d = max_dist;          // strict '<' keeps the first minimum (Q10)
        const R d0 = rsqrt_of(s0), d1 = rsqrt_of(s1), d2 = rsqrt_of(s2), d3 = rsqrt_of(s3);
        if (d0 < d) { d = d0; P1 = r0; P2 = C; }
        if (d1 < d) { d = d1; P1 = r1; P2 = D; }
        if (d2 < d) { d = d2; P1 = A; P2 = c2; }
        if (d3 < d) { d = d3; P1 = B; P2 = c3; }
        return d;
    }
}

// distanceLineSegToPoint (CollisionAndDistance.cpp:323-342)
template <class R, bool PTS>
__device__ __forceinline__ R seg_point(const Capsule<R> &cp, const V3<R> &C, R eq_thr, V3<R> &P1, V3<R> &P2) {
    const R t = dotF(sub(C, cp.A), cp.AB) / cp.alpha1;
    const V3<R> p1 = t < R(0) ? cp.A : (t > R(1) ? cp.B : lerp_from(cp.A, t, cp.AB));
    const R d = rsqrt_of(sqD(C.x - p1.x, C.y - p1.y, C.z - p1.z));
    if (d < eq_thr) return R(0);
    if (PTS) { P1 = p1; P2 = C; }
    return d;
}

// in-plane axes (a < b) of a face with normal axis c
__device__ __forceinline__ int axis_a(int c) { return c == 0 ? 1 : 0; }
__device__ __forceinline__ int axis_b(int c) { return c == 2 ? 1 : 2; }

// collisionCapsuleToRectangle with radius == 0 (CollisionAndDistance.cpp:46-111, Q6-Q7): true iff the segment
// crosses the face plane strictly inside the face rectangle.  face in 0..5 = x_min,y_min,z_min,x_max,y_max,z_max.
template <class R>
__device__ __forceinline__ bool segment_crosses_face(const Capsule<R> &cp, const Box<R> &bx, int face, R max_dist) {
    const int c = face > 2 ? face - 3 : face, a = axis_a(c), b = axis_b(c);
    const R value = face > 2 ? sel3(bx.hi, c) : sel3(bx.lo, c);
    const R ABc = sel3(cp.AB, c);
    const R t = ABc != R(0) ? (value - sel3(cp.A, c)) / ABc : max_dist;
    if (!(t > R(0) && t < R(1))) return false;
    const R M0 = sel3(cp.A, a) + t * sel3(cp.AB, a), M1 = sel3(cp.A, b) + t * sel3(cp.AB, b);
    return M0 > sel3(bx.lo, a) && M0 < sel3(bx.hi, a) && M1 > sel3(bx.lo, b) && M1 < sel3(bx.hi, b);
}

// 6-bit edge code: low 3 bits = corner C, high 3 bits = corner D (bit k set = max side on axis k)
__device__ __forceinline__ unsigned edge_code(unsigned c, unsigned d) { return c | (d << 3); }
// ring of the 4 edges of face (c, side) in the reference's order p00->p10->p11->p01->p00, packed 6 bits each
__device__ __forceinline__ unsigned ring_codes(int c, int side) {
    const unsigned s = unsigned(side) << c, ua = 1u << axis_a(c), ub = 1u << axis_b(c);
    return edge_code(s, s | ua) | (edge_code(s | ua, s | ua | ub) << 6) | (edge_code(s | ua | ub, s | ub) << 12) |
           (edge_code(s | ub, s) << 18);
}

// ---------------------------------------------------------------------------------------------------------------
// CapsuleToBox::compute (CollisionAndDistance.cpp:409-755), split in two so that the work-queue kernel can run the
// two halves in separately compacted phases:
//   classify_pair()  stage 1 (endpoint projections) + the case selection of stage 2: everything except the
//                    segment-edge evaluations.  It either finishes the pair (`done`, `d` = final distance) or
//                    leaves `d` = the best candidate so far (face distance or max_dist, radius NOT yet subtracted)
//                    plus the list of box edges the reference would feed to distanceLineSegToLineSeg;
//   seg_edge()       one such evaluation.
// capsule_box_distance() below is the sequential composition (used by FULL mode, the primitives and the replay of
// a colliding pair).
// ---------------------------------------------------------------------------------------------------------------
template <class R>
struct PairPlan {
    R d;              // done: final pair distance (radius subtracted; exactly 0 when an endpoint is inside, Q11)
    unsigned list;    // up to 4 edge codes, 6 bits each (ignored when twelve)
    int count;        // edges to evaluate
    bool twelve;      // all 12 edges in the reference's order (:746-751)
    bool done;
    unsigned cand;    // PTS only: where the candidate points behind `d` come from: bits 0-2 = face (x_min, y_min, z_min, x_max, y_max, z_max) the
                      // endpoint projects onto, 7 = no candidate yet (P1 = P2 = 0); bit 3 = the endpoint is B
};

// k-th edge code of a plan
template <class R>
__device__ __forceinline__ unsigned plan_edge(const PairPlan<R> &pl, int k) {
    if (!pl.twelve) return (pl.list >> (6 * k)) & 63u;
    // z_min ring, z_max ring, then the four vertical edges (:746-751)
    unsigned code = k < 4 ? (ring_codes(2, 0) >> (6 * k)) : (k < 8 ? (ring_codes(2, 1) >> (6 * (k - 4)))
                    : edge_code(((0x6u >> (k - 8)) & 1u) | (((0xCu >> (k - 8)) & 1u) << 1), 0));
    if (k >= 8) code |= ((code & 7u) | 4u) << 3;
    return code & 63u;
}

template <class R, bool PTS>
__device__ __forceinline__ PairPlan<R> classify_pair(const Capsule<R> &cp, const Box<R> &bx, R eq_thr, R max_dist, V3<R> &P1, V3<R> &P2) {
    const V3<R> &A = cp.A, &B = cp.B;
    PairPlan<R> pl;
    pl.list = 0; pl.count = 0; pl.twelve = false; pl.done = true; pl.d = R(0); pl.cand = 7u;
    // --- stage 1: endpoint classification (projectionLineSegOnSide x3, :451-475)
    const bool ainx = A.x >= bx.lo.x && A.x <= bx.hi.x, ainy = A.y >= bx.lo.y && A.y <= bx.hi.y, ainz = A.z >= bx.lo.z && A.z <= bx.hi.z;
    const bool binx = B.x >= bx.lo.x && B.x <= bx.hi.x, biny = B.y >= bx.lo.y && B.y <= bx.hi.y, binz = B.z >= bx.lo.z && B.z <= bx.hi.z;
    bool inside = false;
    unsigned mA = 0, mB = 0;
    R dA = max_dist, dB = max_dist;
#define RVS_PROJECT(c, COND_A, COND_B, LO, HI, PA, PB)                                            \
    if (COND_A) {                                                                                 \
        if (PA > LO && PA < HI) inside = true;                                                    \
        else if (PA <= LO) { mA |= 1u << (c); dA = LO - PA; }                                     \
        else if (PA >= HI) { mA |= 8u << (c); dA = PA - HI; }                                     \
    }                                                                                             \
    if (COND_B) {                                                                                 \
        if (PB > LO && PB < HI) inside = true;                                                    \
        else if (PB <= LO) { mB |= 1u << (c); dB = LO - PB; }                                     \
        else if (PB >= HI) { mB |= 8u << (c); dB = PB - HI; }                                     \
    }
    RVS_PROJECT(0, ainy && ainz, biny && binz, bx.lo.x, bx.hi.x, A.x, B.x)
    RVS_PROJECT(1, ainx && ainz, binx && binz, bx.lo.y, bx.hi.y, A.y, B.y)
    RVS_PROJECT(2, ainx && ainy, binx && biny, bx.lo.z, bx.hi.z, A.z, B.z)
#undef RVS_PROJECT
    if (inside) return pl;                                              // exactly 0 (Q11)

    if ((mA | mB) != 0u) {
        // --- stage 2a: some endpoint projects onto a face (:420-443)
        const bool useB = !(dA < dB);
        pl.d = dB < dA ? dB : dA;
        const unsigned m = useB ? mB : mA;
        const V3<R> &Pn = useB ? B : A;
        if (PTS) {
            const int face = __ffs(int(m)) - 1;
            const int c = face > 2 ? face - 3 : face;
            P1 = Pn; P2 = Pn;
            put3(P2, c, face > 2 ? sel3(bx.hi, c) : sel3(bx.lo, c));
            pl.cand = unsigned(face) | (useB ? 8u : 0u);
        }
        if ((mA & mB) == 0u) {
            // checkEdges (:477-574): first face in the order x_min, x_max, y_min, y_max, z_min, z_max
            const int f = (m & 1u) ? 0 : (m & 8u) ? 3 : (m & 2u) ? 1 : (m & 16u) ? 4 : (m & 4u) ? 2 : 5;
            if (segment_crosses_face(cp, bx, f, max_dist)) { pl.d = R(0) - cp.radius; return pl; }
            const int c = f > 2 ? f - 3 : f, a = axis_a(c), b = axis_b(c);
            const V3<R> &Po = useB ? A : B;                         // the endpoint without projection
            const R oa = sel3(Po, a), ob = sel3(Po, b);
            const unsigned s = unsigned(f > 2) << c, ua = 1u << a, ub = 1u << b;
            if (oa < sel3(bx.lo, a)) { pl.list = edge_code(s, s | ub); pl.count = 1; }
            else if (oa > sel3(bx.hi, a)) { pl.list = edge_code(s | ua, s | ua | ub); pl.count = 1; }
            if (ob < sel3(bx.lo, b)) { pl.list |= edge_code(s, s | ua) << (6 * pl.count); pl.count++; }
            else if (ob > sel3(bx.hi, b)) { pl.list |= edge_code(s | ub, s | ub | ua) << (6 * pl.count); pl.count++; }
        }
        if (pl.count == 0) { pl.d = pl.d - cp.radius; return pl; }
        pl.done = false;
        return pl;
    }
    // --- stage 2b: no projection (checkOtherCases, :597-755)
    const int cx = (A.x < bx.lo.x && B.x < bx.lo.x) ? 0 : ((A.x > bx.hi.x && B.x > bx.hi.x) ? 1 : 2);
    const int cy = (A.y < bx.lo.y && B.y < bx.lo.y) ? 0 : ((A.y > bx.hi.y && B.y > bx.hi.y) ? 1 : 2);
    const int cz = (A.z < bx.lo.z && B.z < bx.lo.z) ? 0 : ((A.z > bx.hi.z && B.z > bx.hi.z) ? 1 : 2);
    const int mixed = (cx == 2) + (cy == 2) + (cz == 2);
    const unsigned base = unsigned(cx == 1) | (unsigned(cy == 1) << 1) | (unsigned(cz == 1) << 2);
    pl.d = max_dist;
    if (mixed == 0) {
        pl.d = seg_point<R, PTS>(cp, bx.corner(int(base)), eq_thr, P1, P2) - cp.radius;
        return pl;
    } else if (mixed == 1) {
        const unsigned um = cx == 2 ? 1u : (cy == 2 ? 2u : 4u);
        pl.list = edge_code(base, base | um); pl.count = 1;
    } else if (mixed == 2) {
        const int c = cx != 2 ? 0 : (cy != 2 ? 1 : 2);
        pl.list = ring_codes(c, int((base >> c) & 1u)); pl.count = 4;
    } else {
        bool crossing = false;
#pragma unroll 1
        for (int f = 0; f < 6; f++) crossing = crossing || segment_crosses_face(cp, bx, f, max_dist);
        if (crossing) { pl.d = R(0) - cp.radius; return pl; }
        pl.twelve = true; pl.count = 12;
    }
    pl.done = false;
    return pl;
}

// one planned edge: decode the 6-bit code and evaluate distanceLineSegToLineSeg against it
template <class R, bool PTS>
__device__ __forceinline__ R plan_edge_distance(const Capsule<R> &cp, const Box<R> &bx, unsigned code, R eq_thr, R max_dist, V3<R> &q1, V3<R> &q2) {
    const unsigned cc = code & 7u, dc = (code >> 3) & 7u;
    const int e = __ffs(int(cc ^ dc)) - 1;
    const V3<R> C = bx.corner(int(cc));
    const R De = ((dc >> e) & 1u) ? sel3(bx.hi, e) : sel3(bx.lo, e);
    return seg_edge<R, PTS>(cp, C, e, De, eq_thr, max_dist, q1, q2);
}

// The candidate points classify_pair<R, true> leaves behind an UNDECIDED pair, rebuilt from PairPlan::cand — for callers that evaluate the
// planned edges later (rvs_distance_full.cuh defers them): the endpoint that projects onto a face and its projection, or no points yet.
template <class R>
__device__ __forceinline__ void plan_candidate_points(const Capsule<R> &cp, const Box<R> &bx, unsigned cand, V3<R> &P1, V3<R> &P2) {
    P1 = {0, 0, 0}; P2 = {0, 0, 0};
    if ((cand & 7u) != 7u) {
        const int face = int(cand & 7u), c = face > 2 ? face - 3 : face;
        P1 = (cand & 8u) ? cp.B : cp.A; P2 = P1;
        put3(P2, c, face > 2 ? sel3(bx.hi, c) : sel3(bx.lo, c));
    }
}

// distance from the capsule to the box, radius already subtracted; exactly 0 when an endpoint lies inside the box
// (Q11); -radius on the snap/crossing paths (Q14).
template <class R, bool PTS>
__device__ __forceinline__ R capsule_box_distance(const Capsule<R> &cp, const Box<R> &bx, R eq_thr, R max_dist, V3<R> &P1, V3<R> &P2) {
    const PairPlan<R> pl = classify_pair<R, PTS>(cp, bx, eq_thr, max_dist, P1, P2);
    if (pl.done) return pl.d;
    R d = pl.d;
    // --- uniform edge loop (distanceToMoreLineSegments, :576-595)
#pragma unroll 1
    for (int k = 0; k < pl.count; k++) {
        V3<R> q1, q2;
        const R de = plan_edge_distance<R, PTS>(cp, bx, plan_edge(pl, k), eq_thr, max_dist, q1, q2);
        if (de <= R(0)) return R(0) - cp.radius;
        if (de < d) { d = de; if (PTS) { P1 = q1; P2 = q2; } }
    }
    return d - cp.radius;
}

// ---------------------------------------------------------------------------------------------------------------
// Collision predicate: collisionCapsuleToBox / collisionCapsuleToRectangle / checkCases
// (CollisionAndDistance.cpp:4-148).  A different decision tree from the distance (Q3-Q7) — kept as such.
// ---------------------------------------------------------------------------------------------------------------
// The predicate is evaluated in two halves, like the distance (classify_pair / seg_edge):
//   hit_box_plan()        everything that needs no segment-edge distance: the endpoint accept, the slab reject and, per face chosen by
//                         A's side, the crossing test and the projection tests.  It returns 1 when the collision is already decided;
//                         otherwise it has emitted the box edges (6-bit codes, at most 4 per face) whose distance the reference would
//                         compare with the radius, and the pair collides iff one of them is closer than the radius.
//   plan_edge_distance()  one such evaluation.
// Why this is the same verdict in any order: checkCases (:113-148) returns min(d_c1, d_c2) and computes d_c2 only when d_c1 > 0 — but
// d_c1 <= 0 is already < radius (the radius > 0 block), so "min < radius" == "d_c1 < radius or d_c2 < radius"; the callers combine
// faces, endpoints and checkCases results with min / || only (:86-107, :19-40).
template <class R, class Emit>
__device__ __forceinline__ void hit_edge_cases(const Box<R> &bx, int c, unsigned s, R pa, R pb, Emit &emit) {
    const int a = axis_a(c), b = axis_b(c);
    const unsigned ua = 1u << a, ub = 1u << b;
    if (pa < sel3(bx.lo, a)) emit(edge_code(s, s | ub));                      // d_c1: the edge at a_min ...
    else if (pa > sel3(bx.hi, a)) emit(edge_code(s | ua, s | ua | ub));       // ... or at a_max
    if (pb < sel3(bx.lo, b)) emit(edge_code(s, s | ua));                      // d_c2: the edge at b_min ...
    else if (pb > sel3(bx.hi, b)) emit(edge_code(s | ub, s | ub | ua));       // ... or at b_max
}

// collisionCapsuleToRectangle (:46-111) without its segment-edge distances: 1 = true, 0 = not decided here (edges emitted, or false)
template <class R, class Emit>
__device__ __forceinline__ int hit_rectangle_plan(const Capsule<R> &cp, const Box<R> &bx, int face, R max_dist, Emit &emit) {
    const R radius = cp.radius;
    const int c = face > 2 ? face - 3 : face, a = axis_a(c), b = axis_b(c);
    const unsigned s = unsigned(face > 2) << c;
    const R value = face > 2 ? sel3(bx.hi, c) : sel3(bx.lo, c);
    const R r0 = sel3(bx.lo, a), r1 = sel3(bx.lo, b), r2 = sel3(bx.hi, a), r3 = sel3(bx.hi, b);   // rec
    const R Aa = sel3(cp.A, a), Ab = sel3(cp.A, b), Ba = sel3(cp.B, a), Bb = sel3(cp.B, b);
    const R Ac = sel3(cp.A, c), Bc = sel3(cp.B, c), ABc = sel3(cp.AB, c);
    const R t = ABc != R(0) ? (value - Ac) / ABc : max_dist;
    const R M0 = Aa + t * sel3(cp.AB, a), M1 = Ab + t * sel3(cp.AB, b);
    const R nA = rsqrt_of((Ac - value) * (Ac - value)), nB = rsqrt_of((Bc - value) * (Bc - value));   // |A - A_proj|, |B - B_proj|
    const R nmin = nB < nA ? nB : nA;

    if (t > R(0) && t < R(1)) {
        if (M0 > r0 - radius && M0 < r2 + radius && M1 > r1 - radius && M1 < r3 + radius) {
            // corner-disc clauses index `rec` as the reference does (Q5): verbatim, including the slip
            if ((M0 > r0 - radius && M0 < r2 + radius && M1 > r1 && M1 < r3) ||
                (M1 > r1 - radius && M1 < r3 + radius && M0 > r0 && M0 < r2) ||
                (M0 < r0 && M1 < r2 && rsqrt_of((M0 - r0) * (M0 - r0) + (M1 - r2) * (M1 - r2)) < radius) ||
                (M0 < r0 && M1 > r3 && rsqrt_of((M0 - r0) * (M0 - r0) + (M1 - r3) * (M1 - r3)) < radius) ||
                (M0 > r1 && M1 < r2 && rsqrt_of((M0 - r1) * (M0 - r1) + (M1 - r2) * (M1 - r2)) < radius) ||
                (M0 > r1 && M1 > r3 && rsqrt_of((M0 - r1) * (M0 - r1) + (M1 - r3) * (M1 - r3)) < radius))
                return 1;
        }
    } else if (nmin > radius)
        return 0;

    if (radius > R(0)) {
        const bool Ain = Aa > r0 && Aa < r2 && Ab > r1 && Ab < r3;
        const bool Bin = Ba > r0 && Ba < r2 && Bb > r1 && Bb < r3;
        if (Ain) {
            if (Bin) return nmin < radius ? 1 : 0;
            if (nA < radius) return 1;
            hit_edge_cases(bx, c, s, Ba, Bb, emit);
        } else if (Bin) {
            if (nB < radius) return 1;
            hit_edge_cases(bx, c, s, Aa, Ab, emit);
        } else {
            hit_edge_cases(bx, c, s, Aa, Ab, emit);
            hit_edge_cases(bx, c, s, Ba, Bb, emit);
        }
    }
    return 0;
}

// collisionCapsuleToBox (:4-42) without the segment-edge distances: 1 = collision, 0 = the emitted edges (possibly none) decide
template <class R, class Emit>
__device__ __forceinline__ int hit_box_plan(const Capsule<R> &cp, const Box<R> &bx, R max_dist, Emit &emit) {
    const V3<R> &A = cp.A, &B = cp.B;
    const R rn = cp.radius_in, radius = cp.radius;
    if ((A.x > bx.lo.x - rn && A.y > bx.lo.y - rn && A.z > bx.lo.z - rn && A.x < bx.hi.x + rn && A.y < bx.hi.y + rn && A.z < bx.hi.z + rn) ||
        (B.x > bx.lo.x - rn && B.y > bx.lo.y - rn && B.z > bx.lo.z - rn && B.x < bx.hi.x + rn && B.y < bx.hi.y + rn && B.z < bx.hi.z + rn))
        return 1;
    if ((A.x < bx.lo.x - radius && B.x < bx.lo.x - radius) || (A.x > bx.hi.x + radius && B.x > bx.hi.x + radius) ||
        (A.y < bx.lo.y - radius && B.y < bx.lo.y - radius) || (A.y > bx.hi.y + radius && B.y > bx.hi.y + radius) ||
        (A.z < bx.lo.z - radius && B.z < bx.lo.z - radius) || (A.z > bx.hi.z + radius && B.z > bx.hi.z + radius))
        return 0;
    // faces chosen by A's side only, x then y then z (Q4)
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        const R ak = sel3(A, k);
        const int face = ak < sel3(bx.lo, k) ? k : (ak > sel3(bx.hi, k) ? k + 3 : -1);
        if (face >= 0 && hit_rectangle_plan(cp, bx, face, max_dist, emit)) return 1;
    }
    return 0;
}

// up to 12 edge codes of one pair (4 per face), 6 bits each, in two 64-bit words
struct HitEdges {
    unsigned long long lo = 0ull, hi = 0ull;
    int count = 0;
    __device__ __forceinline__ void operator()(unsigned code) {
        if (count < 10) lo |= (unsigned long long)code << (6 * count);
        else hi |= (unsigned long long)code << (6 * (count - 10));
        count++;
    }
    __device__ __forceinline__ unsigned at(int k) const { return unsigned((k < 10 ? lo >> (6 * k) : hi >> (6 * (k - 10))) & 63ull); }
};

// the sequential composition (one-thread-per-item kernels, primitives, motion check)
template <class R>
__device__ __forceinline__ bool capsule_hits_box(const Capsule<R> &cp, const Box<R> &bx, R eq_thr, R max_dist) {
    HitEdges edges;
    if (hit_box_plan(cp, bx, max_dist, edges)) return true;
#pragma unroll 1
    for (int k = 0; k < edges.count; k++) {
        V3<R> u, v;
        if (plan_edge_distance<R, false>(cp, bx, edges.at(k), eq_thr, max_dist, u, v) < cp.radius) return true;
    }
    return false;
}

}  // namespace rvs
