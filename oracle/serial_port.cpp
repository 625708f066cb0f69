// serial_port.cpp — CPU restatement of RealTrace's Serial render core.
//
// TEST INFRASTRUCTURE ONLY (see oracle_abi.h).  Written from scratch as flat
// FP64 C-style code; every function names the reference lines
// (/root/reference/Serial/<file>:<line>) whose arithmetic it follows, operation
// order included, because tests/test_oracle_pinning.py requires this port to be
// BIT-IDENTICAL (RGB8 frame, first-hit id, float t, ray count) to the
// reference's own sources compiled by oracle/build_ref.py, and to the frame
// hashes pinned in tests/golden/.  The port exists so that a checker travels
// to machines where /root/reference does not (the GPU box) and so the
// algorithm is stated in one readable place.
//
// Build: g++ -std=c++14 -O3 -fPIC -shared -pthread (no -march=native, no
// -ffast-math: FMA contraction would break bit-equality with the reference).

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <limits>
#include <thread>
#include <utility>
#include <vector>

#include "oracle_abi.h"

namespace {

// ----------------------------------------------------------------------------- math
struct V3 { double x, y, z; };                       // vector3D.h:7-56 / color.h:5-35 (same 3 doubles)

inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }          // vector3D.cpp:32-35
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }          // :37-40
inline V3 mul(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }            // :47-55
inline V3 mulc(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }         // :57-60, color.cpp:36-37
inline V3 divs(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }           // :42-45
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }                               // :16-17
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      // :112-113
inline double length(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }   // :88-89
inline V3 unit(V3 a) { return divs(a, length(a)); }                              // :94-95
inline V3 cross(V3 a, V3 b) {                                                    // :103-110
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double det2(double a, double b, double c, double d) { return a * d - b * c; }   // utilities.cpp:4-9
inline double det3(V3 c1, V3 c2, V3 c3) {                                        // utilities.cpp:17-22
    return c1.x * det2(c2.y, c3.y, c2.z, c3.z) - c1.y * det2(c2.x, c3.x, c2.z, c3.z) +
           c1.z * det2(c2.x, c3.x, c2.y, c3.y);
}

const float kSmallestDist = 1e-4;   // ray.h:10 (a float constant)
const double kEps = 1e-7;           // triangle.h:12, plane.cpp:9

// ----------------------------------------------------------------------------- ray
struct RayRec {                     // ray.h:11-47
    V3 o, d;
    float t;
    bool hit;
    int obj;                        // index into Scene::objects (the `object` pointer)
    int level;
};
inline RayRec make_ray(V3 o, V3 d, int level = 0) {   // ray.h:25-29: direction normalised, t = FLT_MAX
    RayRec r;
    r.o = o; r.d = unit(d); r.t = FLT_MAX; r.hit = false; r.obj = -1; r.level = level;
    return r;
}
inline V3 position(const RayRec& r) { return add(r.o, mul(r.d, (double)r.t)); }   // ray.h:32
inline bool set_parameter(RayRec& r, float par, int obj) {                          // ray.cpp:3-13
    if (par < r.t && par > kSmallestDist) { r.hit = true; r.t = par; r.obj = obj; return true; }
    return false;
}

// ----------------------------------------------------------------------------- scene
enum Kind { TRI, SPH, PLN, CYL };
struct Mat { V3 color; double ka, kd, ks, kr, kt, eta; bool bary; };
struct Obj {
    Kind kind;
    int mat;
    V3 a, b, c, d;        // TRI: a,b,c   PLN: p1..p4   SPH: a = centre   CYL: a = position, b = up
    double radius;
    V3 nrm;               // PLN: constant normal (plane.h:24)
    V3 col[3];            // per-face vertex colours when the material is barycentric
    uint32_t object_id;
};
struct Box { double lo[3], hi[3]; };

struct Grid {             // uniform-grid.h:22-49
    Box bounds;
    double delta[3], width[3], inv_width[3];
    int n[3];
    int nv = 0;
    std::vector<std::vector<int>> cell;   // object indices, ascending (Voxel::idx / ::primitives)
    bool present = false;
};

struct SceneRec {
    std::vector<Obj> objects;             // World::objectList order
    std::vector<int> analytic;            // non-triangle object indices
    std::vector<Mat> mats;
    std::vector<std::pair<V3, V3>> lights;   // position, intensity
    V3 ambient, background;
    Grid grid;
    int mode = 0;
    int max_depth = 10;
};

struct Counters { uint64_t primary = 0, shadow = 0, secondary = 0; };
enum RayKind { K_PRIMARY, K_SHADOW, K_SECONDARY, K_UNCOUNTED };

// ----------------------------------------------------------------------------- primitive tests
// triangle.cpp:10-24 (also plane.cpp:12-22).  `accept_flag` is what the caller sees as the
// return value: Triangle returns setParameter's result, Plane's helper returns true once the
// barycentric test passes.
inline bool tri_test(RayRec& r, V3 A, V3 B, V3 C, int obj, bool plane_semantics) {
    double a = det3(sub(A, B), sub(A, C), r.d);
    if (std::abs(a) < kEps) return false;
    double beta = det3(sub(A, r.o), sub(A, C), r.d) / a;
    double gamma = det3(sub(A, B), sub(A, r.o), r.d) / a;
    double t = det3(sub(A, B), sub(A, C), sub(A, r.o)) / a;
    if (beta > 0.0 && gamma > 0.0 && beta + gamma < 1.0) {
        bool acc = set_parameter(r, (float)t, obj);
        return plane_semantics ? true : acc;
    }
    return false;
}

inline bool sphere_test(RayRec& r, const Obj& s, int obj) {          // sphere.cpp:5-39
    V3 cv = sub(r.o, s.a);
    double a = 1.0;
    double b = 2 * dot(r.d, cv);
    double c = dot(cv, cv) - s.radius * s.radius;
    double disc = b * b - 4.0 * a * c;
    if (disc >= 0.0) {
        if (disc == 0) {
            double t = -b / (2.0 * a);
            set_parameter(r, (float)t, obj);
            return true;
        }
        double D = std::sqrt(disc);
        double t1 = (-b + D) / (2.0 * a);
        double t2 = (-b - D) / (2.0 * a);
        bool b1 = set_parameter(r, (float)t1, obj);
        bool b2 = set_parameter(r, (float)t2, obj);
        return b1 || b2;
    }
    return false;
}

inline bool solve_quadratic(double A, double B, double C, double& r1, double& r2) {   // cylinder.cpp:4-12
    double D = B * B - 4 * A * C;
    if (D < 0) return false;
    r1 = (-B + std::sqrt(D)) / (2 * A);
    r2 = (-B - std::sqrt(D)) / (2 * A);
    if (r1 > r2) std::swap(r1, r2);
    return true;
}

inline bool cylinder_test(RayRec& r, const Obj& c, int obj) {        // cylinder.cpp:14-32
    V3 up = c.b;
    V3 t1v = sub(r.d, mul(up, dot(r.d, up)));
    V3 oc = sub(r.o, c.a);
    V3 t2v = sub(oc, mul(up, dot(oc, up)));
    double A = dot(t1v, t1v);
    double B = 2 * dot(t1v, t2v);
    double C = dot(t2v, t2v) - c.radius * c.radius;
    double t1, t2;
    if (solve_quadratic(A, B, C, t1, t2)) {
        if (t1 > 0) set_parameter(r, (float)t1, obj);
        else set_parameter(r, (float)t2, obj);
        return true;
    }
    return false;
}

inline bool object_test(const SceneRec& s, RayRec& r, int i) {       // Object::intersect dispatch
    const Obj& o = s.objects[i];
    switch (o.kind) {
        case TRI: return tri_test(r, o.a, o.b, o.c, i, false);
        case SPH: return sphere_test(r, o, i);
        case PLN:                                                    // plane.cpp:24-27 (|| short-circuits)
            return tri_test(r, o.a, o.b, o.c, i, true) || tri_test(r, o.a, o.c, o.d, i, true);
        case CYL: return cylinder_test(r, o, i);
    }
    return false;
}

inline V3 normal_at(const Obj& o, V3 p) {
    switch (o.kind) {
        case TRI: return cross(sub(o.a, o.b), sub(o.a, o.c));        // triangle.cpp:26-29
        case SPH: return sub(p, o.a);                                // sphere.cpp:41-44
        case PLN: return o.nrm;                                      // plane.cpp:29-32
        case CYL: {                                                  // cylinder.cpp:34-38
            double t = dot(sub(p, o.a), o.b) / dot(o.b, o.b);
            return sub(sub(p, o.a), mul(o.b, t));
        }
    }
    return {0, 0, 0};
}

// ----------------------------------------------------------------------------- uniform grid
inline Box tri_bound(const Obj& o, bool fixed) {                     // triangle.cpp:31-40 + utilities.h:13-17
    Box b;
    for (int k = 0; k < 3; k++) {
        b.lo[k] = std::numeric_limits<double>::max();
        b.hi[k] = fixed ? std::numeric_limits<double>::lowest() : std::numeric_limits<double>::min();
    }
    const V3 v[3] = {o.a, o.b, o.c};
    for (int axis = 0; axis < 3; axis++)
        for (int n = 0; n < 3; n++) {
            double e = axis == 0 ? v[n].x : axis == 1 ? v[n].y : v[n].z;
            b.lo[axis] = std::min(b.lo[axis], e);
            b.hi[axis] = std::max(b.hi[axis], e);
        }
    return b;
}

inline int clampi(int a, int lo, int hi) { if (a > hi) return hi; if (a < lo) return lo; return a; }   // utilities.h:20-26

inline int pos_to_voxel(const Grid& g, double p, int axis) {         // uniform-grid.cpp:41-44
    int v = (int)((p - g.bounds.lo[axis]) * g.inv_width[axis]);
    return clampi(v, 0, g.n[axis] - 1);
}
inline float voxel_to_pos(const Grid& g, int p, int axis) {          // :46-48 (returns float)
    return (float)(g.bounds.lo[axis] + p * g.width[axis]);
}
inline int cell_offset(const Grid& g, int x, int y, int z) {         // :50-52
    return z * g.n[0] * g.n[1] + y * g.n[0] + x;
}

void build_grid(SceneRec& s, const std::vector<int>& tris, bool fixed) {   // uniform-grid.cpp:54-147
    Grid& g = s.grid;
    for (int k = 0; k < 3; k++) {
        g.bounds.lo[k] = std::numeric_limits<double>::max();
        g.bounds.hi[k] = fixed ? std::numeric_limits<double>::lowest() : std::numeric_limits<double>::min();
    }
    for (int i : tris) {
        Box b = tri_bound(s.objects[i], fixed);
        for (int k = 0; k < 3; k++) {
            g.bounds.lo[k] = std::min(g.bounds.lo[k], b.lo[k]);
            g.bounds.hi[k] = std::max(g.bounds.hi[k], b.hi[k]);
        }
    }
    for (int k = 0; k < 3; k++) g.delta[k] = g.bounds.hi[k] - g.bounds.lo[k];
    // findVoxelsPerUnitDist :33-39
    double max_axis = std::max(g.delta[0], std::max(g.delta[1], g.delta[2]));
    double inv_max = 1.0 / max_axis;
    double cube_root = 3.0 * std::pow((double)(int)tris.size(), 1.0 / 3.0);
    double vpud = cube_root * inv_max;
    for (int k = 0; k < 3; k++) {
        g.n[k] = (int)std::ceil(g.delta[k] * vpud);                 // :87
        g.n[k] = clampi(g.n[k], 1, 64);                             // :89
    }
    g.nv = 1;
    for (int k = 0; k < 3; k++) {
        g.width[k] = g.delta[k] / g.n[k];                           // :95
        g.inv_width[k] = (g.width[k] == 0.0L) ? 0.0L : 1.0 / g.width[k];
        g.nv *= g.n[k];
    }
    g.cell.assign(g.nv, {});
    for (int i : tris) {                                            // :106-135
        Box b = tri_bound(s.objects[i], fixed);
        int vmin[3], vmax[3];
        for (int k = 0; k < 3; k++) {
            vmin[k] = pos_to_voxel(g, b.lo[k], k);
            vmax[k] = pos_to_voxel(g, b.hi[k], k);
        }
        for (int z = vmin[2]; z <= vmax[2]; z++)
            for (int y = vmin[1]; y <= vmax[1]; y++)
                for (int x = vmin[0]; x <= vmax[0]; x++) g.cell[cell_offset(g, x, y, z)].push_back(i);
    }
    g.present = true;
}

bool grid_intersect(const SceneRec& s, RayRec& r) {                  // uniform-grid.cpp:149-256
    const Grid& g = s.grid;
    const double o[3] = {r.o.x, r.o.y, r.o.z};
    const double d[3] = {r.d.x, r.d.y, r.d.z};
    double rayT;
    bool flag = true;
    {
        double tmin = (g.bounds.lo[0] - o[0]) / d[0];               // :156-158 (double)
        double tmax = (g.bounds.hi[0] - o[0]) / d[0];
        if (tmin > tmax) std::swap(tmin, tmax);
        float tymin = (float)((g.bounds.lo[1] - o[1]) / d[1]);      // :160-161 (float)
        float tymax = (float)((g.bounds.hi[1] - o[1]) / d[1]);
        if (tymin > tymax) std::swap(tymin, tymax);
        if ((tmin > tymax) || (tymin > tmax)) flag = false;
        if (tymin > tmin) tmin = tymin;
        if (tymax < tmax) tmax = tymax;
        float tzmin = (float)((g.bounds.lo[2] - o[2]) / d[2]);      // :174-175 (float)
        float tzmax = (float)((g.bounds.hi[2] - o[2]) / d[2]);
        if (tzmin > tzmax) std::swap(tzmin, tzmax);
        if ((tmin > tzmax) || (tzmin > tmax)) flag = false;
        if (tzmin > tmin) tmin = tzmin;
        if (tzmax < tmax) tmax = tzmax;
        rayT = tmin;
        if (flag) r.t = (float)rayT;                                // :190 strictSetParameter(float)
    }
    if (!flag) return false;
    V3 gi = position(r);                                            // :199
    const double gip[3] = {gi.x, gi.y, gi.z};
    int pos[3], step[3], out[3];
    double next_t[3], delta_t[3];
    for (int axis = 0; axis < 3; axis++) {                          // :204-221
        pos[axis] = pos_to_voxel(g, gip[axis], axis);
        if (d[axis] >= 0) {
            next_t[axis] = rayT + (voxel_to_pos(g, pos[axis] + 1, axis) - gip[axis]) / d[axis];
            delta_t[axis] = g.width[axis] / d[axis];
            step[axis] = 1;
            out[axis] = g.n[axis];
        } else {
            next_t[axis] = rayT + (voxel_to_pos(g, pos[axis], axis) - gip[axis]) / d[axis];
            delta_t[axis] = -g.width[axis] / d[axis];
            step[axis] = -1;
            out[axis] = -1;
        }
    }
    bool hit_something = false;
    r.t = FLT_MAX;                                                  // :224
    static const int cmp_to_axis[8] = {2, 1, 2, 1, 2, 2, 0, 0};    // :243
    for (;;) {                                                      // :226-253
        const std::vector<int>& cell = g.cell[cell_offset(g, pos[0], pos[1], pos[2])];
        if (!cell.empty()) {
            bool any = false;                                       // Voxel::intersect :9-31
            for (int i : cell) {
                const Obj& t = s.objects[i];
                if (tri_test(r, t.a, t.b, t.c, i, false)) any = true;
            }
            hit_something |= any;
        }
        int bits = ((next_t[0] < next_t[1]) << 2) + ((next_t[0] < next_t[2]) << 1) + ((next_t[1] < next_t[2]));
        int axis = cmp_to_axis[bits];
        pos[axis] += step[axis];
        if (pos[axis] == out[axis]) break;
        if (hit_something) break;                                   // :251 early exit
        next_t[axis] += delta_t[axis];
    }
    return hit_something;
}

// World::firstIntersection, world.cpp:5-17, as patched for the reference build (build_ref.py H3):
// grid first (it overwrites Ray::t), then the analytic objects; TRUE_NEAREST = the linear loop :7-14.
inline void first_intersection(const SceneRec& s, RayRec& r, Counters& n, RayKind kind) {
    if (kind == K_PRIMARY) n.primary++;
    else if (kind == K_SHADOW) n.shadow++;
    else if (kind == K_SECONDARY) n.secondary++;
    if (s.mode == ORACLE_MODE_TRUE_NEAREST) {
        for (int i = 0; i < (int)s.objects.size(); i++) object_test(s, r, i);
    } else {
        if (s.grid.present) grid_intersect(s, r);
        for (int i : s.analytic) object_test(s, r, i);
    }
}

// ----------------------------------------------------------------------------- materials
inline V3 material_shade(const SceneRec& s, const Obj& o, const RayRec& incident) {
    const Mat& m = s.mats[o.mat];
    if (!m.bary) return m.color;                                    // material.cpp:5-8
    // BarycentricMaterial::shade, material.cpp:10-22
    double a = det3(sub(o.a, o.b), sub(o.a, o.c), incident.d);
    if (std::abs(a) < kEps) return {0, 0, 0};                       // `return false` -> Color(0)
    double beta = det3(sub(o.a, incident.o), sub(o.a, o.c), incident.d) / a;
    double gamma = det3(sub(o.a, o.b), sub(o.a, incident.o), incident.d) / a;
    if (!(beta > 0.0 && gamma > 0.0 && beta + gamma < 1.0)) return {0, 0, 0};
    double alpha = 1.0 - (beta + gamma);
    return add(add(mul(o.col[0], alpha), mul(o.col[1], beta)), mul(o.col[2], gamma));
}

inline V3 reflect(V3 I, V3 N) { return sub(I, mul(N, 2.0 * dot(N, I))); }        // world.cpp:27-30
inline bool refract(V3 I, V3 N, double eta, V3& T) {                               // world.cpp:19-25
    double k = 1.0 - eta * eta * (1.0 - dot(N, I) * dot(N, I));
    if (k < 0) return false;
    T = sub(mul(I, eta), mul(N, eta * dot(N, I) + std::sqrt(k)));
    return true;
}

// World::get_light_shade, world.cpp:126-137
inline V3 light_shade(const SceneRec& s, V3 P, V3 normal, V3 Lpos, V3 intensity, const Obj& o, V3 view) {
    const Mat& m = s.mats[o.mat];
    V3 n = unit(normal);
    V3 r = unit(reflect(neg(unit(sub(Lpos, P))), n));
    float diffuse = (float)std::max(dot(n, unit(Lpos)), 0.0);                     // :133 (light POSITION vector)
    float specular = (float)std::max(std::pow(dot(unit(view), r), 128), 0.0);     // :134
    RayRec probe = make_ray(P, view);                                             // :136 Ray(position, viewVector)
    V3 albedo = material_shade(s, o, probe);
    V3 diff = mulc(mul(intensity, m.kd * diffuse), albedo);
    V3 spec = mul(intensity, m.ks * specular);
    return add(diff, spec);
}

const double kEuler = 2.718282;     // world.cpp:2

// World::shade_ray, world.cpp:32-111
V3 shade_ray(const SceneRec& s, RayRec ray, Counters& n, RayKind kind) {
    if (ray.level > s.max_depth) return s.background;               // :34
    first_intersection(s, ray, n, kind);                            // :35
    if (!ray.hit) return s.background;                              // :110
    const Obj& o = s.objects[ray.obj];
    const Mat& m = s.mats[o.mat];
    V3 P = position(ray);
    V3 shadow_color = {0, 0, 0};
    bool is_shadow = false;
    for (const auto& L : s.lights) {                                // :44-51
        V3 toL = sub(L.first, P);
        RayRec sr = make_ray(add(P, mul(toL, 0.01)), toL);
        first_intersection(s, sr, n, K_SHADOW);
        if (sr.hit) {
            is_shadow = true;
            shadow_color = mul(mulc(s.ambient, material_shade(s, o, ray)), m.ka);
        }
    }
    V3 light_color = {0, 0, 0};
    V3 Nraw = normal_at(o, P);
    for (const auto& L : s.lights)                                  // :54-58
        light_color = add(light_color, light_shade(s, P, Nraw, L.first, L.second, o, ray.d));
    light_color = add(light_color, mul(mulc(s.ambient, material_shade(s, o, ray)), m.ka));   // :59
    V3 final_color = light_color;
    if (is_shadow) final_color = add(mul(final_color, 1e-4), mul(shadow_color, 1 - 1e-4));   // :63

    V3 N = unit(Nraw);                                              // :66-69
    V3 I = unit(ray.d);
    double eta = m.eta;
    V3 T = {0.0, 0.0, 0.0};
    double t = ray.t;
    double c = 0;
    V3 k = {1.0, 1.0, 1.0};
    int level = ray.level;
    if (m.kr > 0 && m.kt > 0) {                                     // :77-100 dielectric
        V3 R = reflect(I, N);
        if (dot(ray.d, N) < 0) {
            refract(I, N, eta, T);                                  // :83 result ignored
            c = -dot(I, N);
        } else {
            k = {std::pow(kEuler, -1.0 * 0.27 * t), std::pow(kEuler, -1.0 * 0.45 * t),
                 std::pow(kEuler, -1.0 * 0.55 * t)};                // :88
            if (refract(I, mul(N, -1.0), 1 / eta, T)) c = dot(T, N);
            else {
                RayRec tmp = make_ray(add(P, mul(R, 1e-4)), R, level + 1);
                return mulc(k, shade_ray(s, tmp, n, K_SECONDARY));  // :91-92
            }
        }
        double R0 = ((eta - 1) * (eta - 1)) / ((eta + 1) * (eta + 1));
        double Rs = R0 + (1 - R0) * std::pow(1 - c, 5);             // :96
        RayRec t1 = make_ray(add(P, mul(R, 1e-4)), R, level + 1);
        RayRec t2 = make_ray(add(P, mul(T, 1e-4)), T, level * 2);   // :98
        V3 A = shade_ray(s, t1, n, K_SECONDARY);
        V3 B = shade_ray(s, t2, n, K_SECONDARY);
        return mulc(k, add(mul(A, Rs), mul(B, 1 - Rs)));            // :99
    } else if (m.kr > 0) {                                          // :101-107 mirror
        V3 R = reflect(I, N);
        RayRec rr = make_ray(add(P, mul(R, 1e-4)), R, level + 1);
        final_color = add(final_color, mul(shade_ray(s, rr, n, K_SECONDARY), m.kr));
    }
    return final_color;
}

// ----------------------------------------------------------------------------- camera
struct Cam {                         // camera.h:7-34
    V3 pos, u, v, w;
    float focal, aspect;
    int width, height;
};
Cam make_camera(const oracle_camera* c) {                            // camera.cpp:4-25
    Cam cam;
    cam.pos = {c->pos[0], c->pos[1], c->pos[2]};
    V3 target = {c->target[0], c->target[1], c->target[2]};
    V3 up = unit(V3{c->up[0], c->up[1], c->up[2]});
    V3 los = sub(target, cam.pos);
    cam.w = unit(neg(los));
    cam.u = unit(cross(up, cam.w));
    cam.v = unit(cross(cam.w, cam.u));
    cam.width = c->width;
    cam.height = c->height;
    float focal_height = 1.0;
    cam.aspect = float(c->width) / float(c->height);
    float fovy = c->fovy;
    cam.focal = focal_height / (2.0 * std::tan(fovy * M_PI / (180.0 * 2.0)));
    return cam;
}
inline V3 ray_direction(const Cam& c, int i, int j) {                // camera.cpp:33-44
    V3 dir = {0.0, 0.0, 0.0};
    dir = add(dir, mul(neg(c.w), (double)c.focal));
    float xw = c.aspect * (i - c.width / 2.0 + 0.5) / c.width;
    float yw = (j - c.height / 2.0 + 0.5) / c.height;
    dir = add(dir, mul(c.u, (double)xw));
    dir = add(dir, mul(c.v, (double)yw));
    return unit(dir);
}
inline void clamp_color(V3& c) {                                     // color.cpp:19-28
    if (c.x > 1.0f) c.x = 1.0f;
    if (c.y > 1.0f) c.y = 1.0f;
    if (c.z > 1.0f) c.z = 1.0f;
    if (c.x < 0.0f) c.x = 0.0f;
    if (c.y < 0.0f) c.y = 0.0f;
    if (c.z < 0.0f) c.z = 0.0f;
}

// ----------------------------------------------------------------------------- scene construction
inline V3 v3f(const float* p) { return {(double)p[0], (double)p[1], (double)p[2]}; }

bool build_scene(const oracle_scene* in, int mode, int max_depth, SceneRec& s) {
    s.mode = mode;
    s.max_depth = max_depth;
    s.ambient = v3f(in->ambient);
    s.background = v3f(in->background);
    for (uint32_t i = 0; i < in->n_materials; i++) {
        const oracle_material& m = in->materials[i];
        s.mats.push_back({v3f(m.color), m.ka, m.kd, m.ks, m.kr, m.kt, m.eta, m.barycentric != 0});
    }
    std::vector<Obj> objs;
    for (uint32_t i = 0; i < in->n_tri; i++) {
        const float* v = in->tri_v + 9 * (size_t)i;
        Obj o{};
        o.kind = TRI; o.mat = (int)in->tri_material[i];
        if (o.mat >= (int)in->n_materials) return false;
        o.a = v3f(v); o.b = v3f(v + 3); o.c = v3f(v + 6);
        if (s.mats[o.mat].bary) {
            if (!in->tri_rgb) return false;
            const float* c = in->tri_rgb + 9 * (size_t)i;
            o.col[0] = v3f(c); o.col[1] = v3f(c + 3); o.col[2] = v3f(c + 6);
        }
        o.object_id = in->tri_object_id ? in->tri_object_id[i] : i;
        objs.push_back(o);
    }
    uint32_t next_id = in->n_tri;
    for (uint32_t i = 0; i < in->n_sph; i++, next_id++) {
        const float* p = in->sph + 4 * (size_t)i;
        Obj o{};
        o.kind = SPH; o.mat = (int)in->sph_material[i];
        if (o.mat >= (int)in->n_materials || s.mats[o.mat].bary) return false;
        o.a = v3f(p); o.radius = p[3];
        o.object_id = in->sph_object_id ? in->sph_object_id[i] : next_id;
        objs.push_back(o);
    }
    for (uint32_t i = 0; i < in->n_pln; i++, next_id++) {
        const float* p = in->pln + 12 * (size_t)i;
        Obj o{};
        o.kind = PLN; o.mat = (int)in->pln_material[i];
        if (o.mat >= (int)in->n_materials || s.mats[o.mat].bary) return false;
        o.a = v3f(p); o.b = v3f(p + 3); o.c = v3f(p + 6); o.d = v3f(p + 9);
        o.nrm = cross(sub(o.c, o.a), sub(o.b, o.a));                // plane.h:24
        o.object_id = in->pln_object_id ? in->pln_object_id[i] : next_id;
        objs.push_back(o);
    }
    for (uint32_t i = 0; i < in->n_cyl; i++, next_id++) {
        const float* p = in->cyl + 7 * (size_t)i;
        Obj o{};
        o.kind = CYL; o.mat = (int)in->cyl_material[i];
        if (o.mat >= (int)in->n_materials || s.mats[o.mat].bary) return false;
        o.a = v3f(p); o.radius = p[3]; o.b = v3f(p + 4);
        o.object_id = in->cyl_object_id ? in->cyl_object_id[i] : next_id;
        objs.push_back(o);
    }
    std::stable_sort(objs.begin(), objs.end(), [](const Obj& a, const Obj& b) { return a.object_id < b.object_id; });
    s.objects.swap(objs);
    std::vector<int> tris;
    for (int i = 0; i < (int)s.objects.size(); i++) {
        if (s.objects[i].kind == TRI) tris.push_back(i);
        else s.analytic.push_back(i);
    }
    for (uint32_t i = 0; i < in->n_lights; i++) {
        const float* p = in->lights + 6 * (size_t)i;
        s.lights.push_back({v3f(p), v3f(p + 3)});
    }
    if (!tris.empty() && mode != ORACLE_MODE_TRUE_NEAREST) build_grid(s, tris, mode == ORACLE_MODE_BBOX_FIXED);
    return true;
}

}  // namespace

extern "C" const char* oracle_name(void) { return "port:serial_port.cpp"; }

extern "C" int oracle_render(const oracle_scene* scene, const oracle_camera* c, int max_depth, int mode,
                             int col_begin, int col_step, int nthreads, uint8_t* rgb, int32_t* prim_id,
                             float* t_hit, oracle_result* result) {
    if (!scene || !c || !rgb || col_step < 1 || col_begin < 0 || c->width < 1 || c->height < 1) return -1;
    auto t0 = std::chrono::steady_clock::now();
    SceneRec s;
    if (!build_scene(scene, mode, max_depth, s)) return -2;
    double build_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    Cam cam = make_camera(c);
    const int W = c->width, H = c->height;
    if (nthreads < 1) nthreads = 1;
    std::vector<int> cols;
    for (int i = col_begin; i < W; i += col_step) cols.push_back(i);
    std::vector<Counters> per_thread(nthreads);

    auto t1 = std::chrono::steady_clock::now();
    auto worker = [&](int tid) {
        Counters n;
        for (size_t k = tid; k < cols.size(); k += nthreads) {
            int i = cols[k];
            for (int j = 0; j < H; j++) {                           // renderengine.cpp:13-18
                RayRec ray = make_ray(cam.pos, ray_direction(cam, i, j));   // :5-6
                V3 color = shade_ray(s, ray, n, K_PRIMARY);
                clamp_color(color);
                size_t o = ((size_t)i + (size_t)j * W) * 3;         // camera.cpp:46-52
                rgb[o + 0] = (unsigned char)(255 * color.x);
                rgb[o + 1] = (unsigned char)(255 * color.y);
                rgb[o + 2] = (unsigned char)(255 * color.z);
            }
        }
        per_thread[tid] = n;
    };
    if (nthreads == 1) worker(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++) pool.emplace_back(worker, t);
        for (auto& t : pool) t.join();
    }
    double render_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();

    if (prim_id || t_hit) {
        Counters dummy;
        for (int i : cols)
            for (int j = 0; j < H; j++) {
                RayRec ray = make_ray(cam.pos, ray_direction(cam, i, j));
                first_intersection(s, ray, dummy, K_UNCOUNTED);
                size_t o = (size_t)i + (size_t)j * W;
                if (prim_id) prim_id[o] = ray.hit ? ray.obj : -1;
                if (t_hit) t_hit[o] = ray.hit ? ray.t : FLT_MAX;
            }
    }
    if (result) {
        Counters n;
        for (auto& p : per_thread) { n.primary += p.primary; n.shadow += p.shadow; n.secondary += p.secondary; }
        result->rays_primary = n.primary;
        result->rays_shadow = n.shadow;
        result->rays_secondary = n.secondary;
        result->rays_total = n.primary + n.shadow + n.secondary;
        result->build_seconds = build_s;
        result->render_seconds = render_s;
        result->columns_rendered = (uint32_t)cols.size();
        result->threads_used = (uint32_t)nthreads;
    }
    return 0;
}

extern "C" int oracle_trace_rays(const oracle_scene* scene, int mode, const float* rays, uint32_t n_rays,
                                 int32_t* prim_id, float* t_hit) {
    if (!scene || !rays) return -1;
    SceneRec s;
    if (!build_scene(scene, mode, 0, s)) return -2;
    Counters dummy;
    for (uint32_t r = 0; r < n_rays; r++) {
        const float* p = rays + 6 * (size_t)r;
        RayRec ray = make_ray(v3f(p), v3f(p + 3));
        first_intersection(s, ray, dummy, K_UNCOUNTED);
        if (prim_id) prim_id[r] = ray.hit ? ray.obj : -1;
        if (t_hit) t_hit[r] = ray.hit ? ray.t : FLT_MAX;
    }
    return 0;
}

extern "C" int oracle_shade_rays(const oracle_scene* scene, int max_depth, int mode, const float* rays,
                                 uint32_t n_rays, double* rgb_out) {
    if (!scene || !rays || !rgb_out) return -1;
    SceneRec s;
    if (!build_scene(scene, mode, max_depth, s)) return -2;
    Counters dummy;
    for (uint32_t r = 0; r < n_rays; r++) {
        const float* p = rays + 6 * (size_t)r;
        RayRec ray = make_ray(v3f(p), v3f(p + 3));
        V3 c = shade_ray(s, ray, dummy, K_PRIMARY);
        rgb_out[3 * (size_t)r] = c.x; rgb_out[3 * (size_t)r + 1] = c.y; rgb_out[3 * (size_t)r + 2] = c.z;
    }
    return 0;
}
