// rt_scene.h — device-resident scene records (the data layout in HBM).
//
// Everything the traversal and shading kernels read is a 16-byte-aligned float4 record:
//   BVH node      64 B = 4 x float4   both children's boxes + child codes (two-child-AABB node)
//   triangle      48 B = 3 x float4   vertex a, e1 = a-b, e2 = a-c, stored in Morton order;
//                                     the w lanes carry object id / material id / original index
//   vertex rgb    48 B = 3 x float4   BarycentricMaterial colours, indexed by ORIGINAL triangle
//   material      48 B = 3 x float4
//   light         32 B = 2 x float4
//   analytic      80 B                spheres / quads / cylinders / oversized triangles, tested
//                                     linearly (few of them)
#pragma once

#include "rt_hd.h"

// ---- BVH node ---------------------------------------------------------------------------------
// n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
// n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
// n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
// n3 = (as_float(code0), as_float(code1), as_float(first), as_float(last))   [first,last] = sorted range
// child code >= 0: internal node index; < 0: leaf, ~code = (first << 3) | (count - 1)
#define RT_NODE_FLOAT4S 4
// 4-wide node (128 B = 8 x float4), indexed by the binary node it was collapsed from (grandchildren become
// children; only nodes at even depth are ever reached):
//   q0 = lo.x[0..3]  q1 = hi.x[0..3]  q2 = lo.y  q3 = hi.y  q4 = lo.z  q5 = hi.z  q6 = child codes  q7 = unused
// Empty slots carry the point box (+FLT_MAX)^3 and RT_EMPTY_CODE.
#define RT_NODE4_FLOAT4S 8
#define RT_LEAF_MAX 8
#define RT_LEAF_SHIFT 3

RT_HD int rt_leaf_code(uint32_t first, uint32_t count) { return ~(int)((first << RT_LEAF_SHIFT) | (count - 1u)); }
RT_HD uint32_t rt_leaf_first(int code) { return ((uint32_t)~code) >> RT_LEAF_SHIFT; }
RT_HD uint32_t rt_leaf_count(int code) { return (((uint32_t)~code) & (RT_LEAF_MAX - 1u)) + 1u; }
// An empty child (used when the BVH holds a single leaf): its box is the point (+FLT_MAX)^3, whose
// slab distances are all +-inf, so it is never entered.
#define RT_EMPTY_CODE 0x7fffffff

// ---- analytic primitives ----------------------------------------------------------------------
enum { RT_KIND_SPHERE = 0, RT_KIND_PLANE = 1, RT_KIND_CYLINDER = 2, RT_KIND_TRIANGLE = 3 };

struct AnalyticPrim {
    uint32_t kind, object_id, material, tri_index;   // tri_index: original triangle (vertex colours)
    float4 a, b, c, d;
    // sphere:   a = (centre, radius)
    // plane:    a..d = the four corners (plane.h:21)
    // cylinder: a = (position, radius), b = up (NOT normalised, cylinder.h:17-21)
    // triangle: a = vertex a, b = e1 = a-b, c = e2 = a-c
};

// ---- hit record -------------------------------------------------------------------------------
// prim >= 0 : position of the triangle in the sorted triangle array
// prim == -1: miss
// prim <= -2: analytic primitive (-2 - index)
struct HitRec {
    float t;
    int prim;
    float beta, gamma;   // barycentrics of triangle hits (BarycentricMaterial::shade, material.cpp:15-21)
};
#define RT_MISS (-1)
RT_HD int rt_analytic_code(int k) { return -2 - k; }
RT_HD int rt_analytic_index(int code) { return -2 - code; }

// ---- the scene as the kernels see it -------------------------------------------------------------
struct SceneDev {
    const float4* nodes;          // RT_NODE_FLOAT4S per node; node 0 = root
    const float4* nodes4;         // optional 4-wide view of the same tree (RT_NODE4_FLOAT4S per node), or nullptr
    const float4* tris;           // 3 per triangle, Morton order
    const float4* tri_rgb;        // 3 per ORIGINAL triangle, or nullptr
    const AnalyticPrim* analytic;
    const float4* materials;      // 3 per material: (r,g,b,ka) (kd,ks,kr,kt) (eta,flags,-,-)
    const float4* lights;         // 2 per light: (pos,-) (intensity,-)
    int n_bvh_tris;
    int n_nodes;
    int n_analytic;
    int n_lights;
    float ambient[3];
    float background[3];
};

struct MaterialRec {
    f3 color;
    float ka, kd, ks, kr, kt, eta;
    uint32_t flags;
};
RT_HD MaterialRec load_material(const SceneDev& s, uint32_t id) {
    float4 m0 = ldg(s.materials + 3 * id), m1 = ldg(s.materials + 3 * id + 1), m2 = ldg(s.materials + 3 * id + 2);
    MaterialRec m;
    m.color = mk3(m0);
    m.ka = m0.w; m.kd = m1.x; m.ks = m1.y; m.kr = m1.z; m.kt = m1.w; m.eta = m2.x;
    m.flags = as_uint(m2.y);
    return m;
}

// ---- work counters (RT_FLAG_COUNT_WORK) ----------------------------------------------------------
struct WorkCount {
    uint32_t nodes, tris;
};
