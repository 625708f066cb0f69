// rt_intersect.h — FP32 ray/primitive tests with the reference's acceptance rules.
//
// Each test reproduces the DECISIONS of the reference routine it names (strict interior test,
// |A| < 1e-7 rejection, SMALLEST_DIST < t < current t on float t, root order), while the
// arithmetic is arranged for FP32 accuracy (the oracle is FP64; the closer the FP32 value is to
// the exact one, the fewer epsilon-tie mismatches).  Algebraic identities used are stated inline.
#pragma once

#include "rt_scene.h"

#define RT_SMALLEST_DIST 1e-4f   // ray.h:10
#define RT_TRI_EPS 1e-7f         // triangle.h:12, plane.cpp:9

// Ray::setParameter (ray.cpp:3-13) on a running nearest t.
RT_HD bool accept_t(float par, float& tcur) {
    if (par < tcur && par > RT_SMALLEST_DIST) { tcur = par; return true; }
    return false;
}

// Triangle::intersect (triangle.cpp:10-24) on a pre-differenced record: a, e1 = a-b, e2 = a-c.
//   A       = det(e1, e2, d)          = e1 . c,  c = cofactors of (e2, d)   (utilities.cpp:17-22)
//   beta*A  = det(a-o, e2, d)         = s  . c,  s = a - o
//   gamma*A = det(e1, s, d)           = d  . q,  q = e1 x s
//   t*A     = det(e1, e2, s)          = -(e2 . q)
// Returns 0 = barycentric test failed, 1 = inside but t rejected, 2 = accepted (tcur/beta/gamma updated),
// 3 = inside and EXACTLY as far as the current hit (tie_bg, if given, receives this triangle's beta/gamma): the
// caller breaks the tie by the lower object index — the order in which the reference's linear loop meets them
// (world.cpp:7-14 keeps the first of two equal t), and independent of the order a traversal visits them in.
RT_HD int tri_test(f3 a, f3 e1, f3 e2, f3 o, f3 d, float& tcur, float& beta, float& gamma, float* tie_bg = nullptr) {
    float cx = e2.y * d.z - d.y * e2.z;
    float cy = e2.x * d.z - d.x * e2.z;
    float cz = e2.x * d.y - d.x * e2.y;
    float A = e1.x * cx - e1.y * cy + e1.z * cz;
    if (fabsf(A) < RT_TRI_EPS) return 0;
    f3 s = a - o;
    float bA = s.x * cx - s.y * cy + s.z * cz;
    f3 q = cross(e1, s);
    float gA = dot(d, q);
    // beta > 0 && gamma > 0 && beta + gamma < 1   (signs decided without the division)
    float sg = A > 0.0f ? 1.0f : -1.0f;
    float b = bA * sg, g = gA * sg, aa = A * sg;
    if (!(b > 0.0f && g > 0.0f && b + g < aa)) return 0;
    float inv = rt_rcp(A);
    float t = -dot(e2, q) * inv;
    if (!accept_t(t, tcur)) {
        if (t == tcur && t > RT_SMALLEST_DIST && tie_bg) { tie_bg[0] = bA * inv; tie_bg[1] = gA * inv; return 3; }
        return 1;
    }
    beta = bA * inv;
    gamma = gA * inv;
    return 2;
}

// Sphere::intersect (sphere.cpp:5-39).  With b = 2 d.oc and c = |oc|^2 - r^2 the reference's
// discriminant is b^2 - 4c = 4 (r^2 - |oc - (d.oc) d|^2) for unit d; the right-hand form has no
// cancellation between |oc|^2 and r^2 in FP32.  Roots are tried far-then-near like :31-32.
RT_HD bool sphere_test(f3 centre, float radius, f3 o, f3 d, float& tcur) {
    f3 oc = o - centre;
    float bh = dot(d, oc);
    f3 l = oc - d * bh;
    float dq = radius * radius - dot(l, l);
    if (!(dq >= 0.0f)) return false;
    if (dq == 0.0f) return accept_t(-bh, tcur);
    float D = sqrtf(dq);
    bool b1 = accept_t(-bh + D, tcur);
    bool b2 = accept_t(-bh - D, tcur);
    return b1 || b2;
}

// Plane::intersect (plane.cpp:12-27): triangles (p1,p2,p3) then (p1,p3,p4); the || short-circuits
// as soon as the first one contains the ray, accepted or not.
RT_HD bool plane_test(f3 p1, f3 p2, f3 p3, f3 p4, f3 o, f3 d, float& tcur) {
    float b, g;
    int r = tri_test(p1, p1 - p2, p1 - p3, o, d, tcur, b, g);
    if (r) return r == 2;
    return tri_test(p1, p1 - p3, p1 - p4, o, d, tcur, b, g) == 2;
}

// Cylinder::intersect + solveQuadratic (cylinder.cpp:4-32): infinite cylinder about `up` (not
// normalised).  With t1 = d - (d.up)up, t2 = oc - (oc.up)up: A = t1.t1, B = 2 t1.t2,
// C = t2.t2 - r^2 and B^2 - 4AC = 4 (A r^2 - |t1 x t2|^2)  (Lagrange identity), again free of the
// t2.t2 - r^2 cancellation.  Picks the smaller root if it is > 0, else the larger (:27-28); the
// larger root is never tried when 0 < smaller <= SMALLEST_DIST.
RT_HD bool cylinder_test(f3 pos, float radius, f3 up, f3 o, f3 d, float& tcur) {
    f3 t1v = d - up * dot(d, up);
    f3 oc = o - pos;
    f3 t2v = oc - up * dot(oc, up);
    float A = dot(t1v, t1v);
    float Bh = dot(t1v, t2v);
    f3 w = cross(t1v, t2v);
    float Dq = A * radius * radius - dot(w, w);
    if (Dq < 0.0f) return false;
    float sq = sqrtf(Dq);
    float r1 = (-Bh + sq) / A;
    float r2 = (-Bh - sq) / A;
    if (r1 > r2) { float tmp = r1; r1 = r2; r2 = tmp; }
    if (r1 > 0.0f) return accept_t(r1, tcur);
    return accept_t(r2, tcur);
}

// One analytic primitive against the running nearest hit.
RT_HD bool analytic_test(const AnalyticPrim& p, f3 o, f3 d, float& tcur, float& beta, float& gamma) {
    switch (p.kind) {
        case RT_KIND_SPHERE: return sphere_test(mk3(p.a), p.a.w, o, d, tcur);
        case RT_KIND_PLANE: return plane_test(mk3(p.a), mk3(p.b), mk3(p.c), mk3(p.d), o, d, tcur);
        case RT_KIND_CYLINDER: return cylinder_test(mk3(p.a), p.a.w, mk3(p.b), o, d, tcur);
        default: return tri_test(mk3(p.a), mk3(p.b), mk3(p.c), o, d, tcur, beta, gamma) == 2;
    }
}

// getNormalAtPosition of each kind (sphere.cpp:41-44, plane.h:24, cylinder.cpp:34-38,
// triangle.cpp:26-29) — never normalised, never flipped towards the viewer.
RT_HD f3 analytic_normal(const AnalyticPrim& p, f3 P) {
    switch (p.kind) {
        case RT_KIND_SPHERE: return P - mk3(p.a);
        case RT_KIND_PLANE: return cross(mk3(p.c) - mk3(p.a), mk3(p.b) - mk3(p.a));
        case RT_KIND_CYLINDER: {
            f3 up = mk3(p.b), rel = P - mk3(p.a);
            float t = dot(rel, up) / dot(up, up);
            return rel - up * t;
        }
        default: return cross(mk3(p.b), mk3(p.c));
    }
}
