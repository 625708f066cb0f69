// rt_traverse.h — nearest-hit / any-hit queries: stack-based walk of the two-child-AABB LBVH,
// then the (few) analytic primitives linearly.
//
// Replaces World::firstIntersection -> UniformGrid::intersect -> Voxel::intersect
// (/root/reference/Serial/world.cpp:5-17, uniform-grid.cpp:149-256, :9-31).  Unlike the grid walk,
// which stops at the first voxel that reported any accepted hit (:251), this returns the true
// nearest accepted hit; SURVEY Appendix A Q13 measures that difference (<= 0.05 % of hit pixels).
// Any-hit mode reproduces the shadow test of world.cpp:44-51: no maximum distance, any accepted
// hit (t > SMALLEST_DIST) along the ray counts, even beyond the light.
#pragma once

#include "rt_intersect.h"

#define RT_STACK_SIZE 64

struct RayPrep {
    f3 o, d;
    f3 idir;   // 1 / d, with |d| clamped away from 0 so that products stay finite
    f3 ood;    // o * idir
};

RT_HD float safe_rcp(float v) {
    const float tiny = 8.271806e-25f;   // 2^-80 (Aila & Laine's guard); keeps lo*idir - o*idir finite
    float a = fabsf(v) > tiny ? v : copysignf(tiny, v);
    return 1.0f / a;
}

RT_HD RayPrep prep_ray(f3 o, f3 d) {
    RayPrep r;
    r.o = o; r.d = d;
    r.idir = mk3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    r.ood = o * r.idir;
    return r;
}

// Slab test of one child box against [0, tmax]; returns entry distance in tnear.
RT_HD bool slab(float lox, float hix, float loy, float hiy, float loz, float hiz, const RayPrep& r, float tmax,
                float& tnear) {
    float x0 = fmaf(lox, r.idir.x, -r.ood.x), x1 = fmaf(hix, r.idir.x, -r.ood.x);
    float y0 = fmaf(loy, r.idir.y, -r.ood.y), y1 = fmaf(hiy, r.idir.y, -r.ood.y);
    float z0 = fmaf(loz, r.idir.z, -r.ood.z), z1 = fmaf(hiz, r.idir.z, -r.ood.z);
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    tnear = tn;
    return tn <= tf;
}

RT_HD bool leaf_test(const SceneDev& s, int code, const RayPrep& r, HitRec& hit, bool any_hit, WorkCount* wc) {
    uint32_t first = rt_leaf_first(code), count = rt_leaf_count(code);
    bool found = false;
    for (uint32_t k = 0; k < count; k++) {
        const float4* rec = s.tris + 3 * (size_t)(first + k);
        float4 r0 = ldg(rec), r1 = ldg(rec + 1), r2 = ldg(rec + 2);
        if (wc) wc->tris++;
        float tie[2];
        int rc = tri_test(mk3(r0), mk3(r1), mk3(r2), r.o, r.d, hit.t, hit.beta, hit.gamma, tie);
        if (rc == 2) {
            hit.prim = (int)(first + k);
            found = true;
            if (any_hit) return true;
        } else if (rc == 3 && !any_hit && hit.prim >= 0 && hit.prim != (int)(first + k)) {
            // two triangles at exactly the same distance: the one that comes first in the object list wins, whatever
            // order the walk (or several lanes sharing one ray's walk) met them in
            if (as_uint(r0.w) < as_uint(ldg(s.tris + 3 * (size_t)hit.prim).w)) {
                hit.prim = (int)(first + k);
                hit.beta = tie[0];
                hit.gamma = tie[1];
            }
        }
    }
    return found;
}

// Traversal stacks.  The step functions below take any type with clear() / empty() / push(v) -> false when full /
// pop() (needs !empty()); the persistent kernels bring their own (render.cu: LaneStack), everything else uses this.
struct StackView {             // over an array and a depth the caller owns (kept apart: a struct holding the array would
                               // be placed in local memory as a whole on the device, the depth included)
    int* e;
    int& sp;
    RT_HD StackView(int* e_, int& sp_) : e(e_), sp(sp_) {}
    RT_HD void clear() { sp = 0; }
    RT_HD bool empty() const { return sp == 0; }
    RT_HD bool push(int v) {
        if (sp >= RT_STACK_SIZE) return false;
        e[sp++] = v;
        return true;
    }
    RT_HD int pop() { return e[--sp]; }
};

// Traversal state codes: 0 <= node < RT_DONE is an internal node, node < 0 a leaf, RT_DONE = finished.
#define RT_DONE 0x7fffffff
RT_HD bool rt_is_internal(int node) { return (unsigned)node < (unsigned)RT_DONE; }

// One internal-node step: tests both child boxes against [0, tmax] and returns where to go next —
// the nearer hit child (the farther one is pushed), the only hit child, or the popped stack top.
template <class STK>
RT_HD int bvh_node_step(const SceneDev& s, const RayPrep& r, float tmax, int node, STK& stk, bool* overflow) {
    const float4* n = s.nodes + RT_NODE_FLOAT4S * (size_t)node;
    float4 n0 = ldg(n), n1 = ldg(n + 1), n2 = ldg(n + 2), n3 = ldg(n + 3);
    float t0, t1;
    bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, r, tmax, t0);
    bool h1 = slab(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, r, tmax, t1);
    int c0 = (int)as_uint(n3.x), c1 = (int)as_uint(n3.y);
    if (h0 && h1) {
        if (t1 < t0) { int tmp = c0; c0 = c1; c1 = tmp; }
        if (!stk.push(c1) && overflow) *overflow = true;
        return c0;
    }
    if (h0) return c0;
    if (h1) return c1;
    return stk.empty() ? RT_DONE : stk.pop();
}

// One step of the "if-if" walk with ONE fetch for both kinds of step.  In a warp whose lanes disagree (some at an internal
// node, some at a leaf) the separate functions above run one after the other and each waits for its own loads; here the
// three float4 loads every lane needs — the two child boxes of a node, or the 48-byte record of a leaf's first triangle —
// are issued together from an address chosen per lane, so a diverged iteration waits for memory once instead of twice.
// Same arithmetic, same order of tests, same results as bvh_node_step / leaf_test.
template <class STK>
RT_HD int bvh_step_unified(const SceneDev& s, const RayPrep& r, HitRec& hit, int node, STK& stk, bool any_hit, bool& found,
                           bool* overflow, WorkCount* wc) {
    const bool internal = rt_is_internal(node);
    const uint32_t first = rt_leaf_first(node);
    const float4* p = internal ? s.nodes + RT_NODE_FLOAT4S * (size_t)node : s.tris + 3 * (size_t)first;
    const float4 q0 = ldg(p), q1 = ldg(p + 1), q2 = ldg(p + 2);
    if (internal) {
        if (wc) wc->nodes++;
        const float4 q3 = ldg(p + 3);
        float t0, t1;
        bool h0 = slab(q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, r, hit.t, t0);
        bool h1 = slab(q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, r, hit.t, t1);
        int c0 = (int)as_uint(q3.x), c1 = (int)as_uint(q3.y);
        if (h0 && h1) {
            if (t1 < t0) { int tmp = c0; c0 = c1; c1 = tmp; }
            if (!stk.push(c1) && overflow) *overflow = true;
            return c0;
        }
        if (h0) return c0;
        if (h1) return c1;
        return stk.empty() ? RT_DONE : stk.pop();
    }
    // leaf: the first triangle from the registers, any further ones (RT_LEAF_SIZE > 1) fetched one by one
    const uint32_t count = rt_leaf_count(node);
    float4 r0 = q0, r1 = q1, r2 = q2;
    for (uint32_t k = 0;;) {
        if (wc) wc->tris++;
        float tie[2];
        int rc = tri_test(mk3(r0), mk3(r1), mk3(r2), r.o, r.d, hit.t, hit.beta, hit.gamma, tie);
        if (rc == 2) {
            hit.prim = (int)(first + k);
            found = true;
            if (any_hit) return RT_DONE;
        } else if (rc == 3 && !any_hit && hit.prim >= 0 && hit.prim != (int)(first + k)) {
            if (as_uint(r0.w) < as_uint(ldg(s.tris + 3 * (size_t)hit.prim).w)) {
                hit.prim = (int)(first + k);
                hit.beta = tie[0];
                hit.gamma = tie[1];
            }
        }
        if (++k >= count) break;
        const float4* rec = s.tris + 3 * (size_t)(first + k);
        r0 = ldg(rec); r1 = ldg(rec + 1); r2 = ldg(rec + 2);
    }
    return stk.empty() ? RT_DONE : stk.pop();
}

// The same step on the 4-wide view: four slab tests per fetch, hit children visited nearest first (the
// others are pushed farthest first).  Halves the number of dependent fetches per ray, which is what
// latency-bound work (bounce paths, small frames, per-rank shares) is made of.
RT_HD void sort2(float& ta, int& ca, float& tb, int& cb) {
    if (tb < ta) { float t = ta; ta = tb; tb = t; int c = ca; ca = cb; cb = c; }
}
template <class STK>
RT_HD int bvh4_node_step(const SceneDev& s, const RayPrep& r, float tmax, int node, STK& stk, bool* overflow) {
    const float4* n = s.nodes4 + RT_NODE4_FLOAT4S * (size_t)node;
    float4 lx = ldg(n), hx = ldg(n + 1), ly = ldg(n + 2), hy = ldg(n + 3), lz = ldg(n + 4), hz = ldg(n + 5), cd = ldg(n + 6);
    const float inf = RT_FLT_MAX;
    float t0, t1, t2, t3;
    if (!slab(lx.x, hx.x, ly.x, hy.x, lz.x, hz.x, r, tmax, t0)) t0 = inf;
    if (!slab(lx.y, hx.y, ly.y, hy.y, lz.y, hz.y, r, tmax, t1)) t1 = inf;
    if (!slab(lx.z, hx.z, ly.z, hy.z, lz.z, hz.z, r, tmax, t2)) t2 = inf;
    if (!slab(lx.w, hx.w, ly.w, hy.w, lz.w, hz.w, r, tmax, t3)) t3 = inf;
    int c0 = (int)as_uint(cd.x), c1 = (int)as_uint(cd.y), c2 = (int)as_uint(cd.z), c3 = (int)as_uint(cd.w);
    // 5-comparator network: ascending entry distance, misses (inf) last
    sort2(t0, c0, t1, c1); sort2(t2, c2, t3, c3); sort2(t0, c0, t2, c2); sort2(t1, c1, t3, c3); sort2(t1, c1, t2, c2);
    if (t0 == inf) return stk.empty() ? RT_DONE : stk.pop();
    if (t3 != inf) { if (!stk.push(c3) && overflow) *overflow = true; }
    if (t2 != inf) { if (!stk.push(c2) && overflow) *overflow = true; }
    if (t1 != inf) { if (!stk.push(c1) && overflow) *overflow = true; }
    return c0;
}

// The node steps over a caller-owned array (the CPU emulation's loop models keep their stacks that way).
RT_HD int bvh_node_step(const SceneDev& s, const RayPrep& r, float tmax, int node, int* stack, int& sp, bool* overflow) {
    StackView v(stack, sp);
    return bvh_node_step(s, r, tmax, node, v, overflow);
}
RT_HD int bvh4_node_step(const SceneDev& s, const RayPrep& r, float tmax, int node, int* stack, int& sp, bool* overflow) {
    StackView v(stack, sp);
    return bvh4_node_step(s, r, tmax, node, v, overflow);
}

// The unified step on the 4-wide view: the first three float4 of the wide node (lo.x, hi.x, lo.y of the four children) or
// the leaf's triangle record come from one fetch; a node lane then adds its other four float4.
template <class STK>
RT_HD int bvh4_step_unified(const SceneDev& s, const RayPrep& r, HitRec& hit, int node, STK& stk, bool any_hit, bool& found,
                            bool* overflow, WorkCount* wc) {
    const bool internal = rt_is_internal(node);
    const uint32_t first = rt_leaf_first(node);
    const float4* p = internal ? s.nodes4 + RT_NODE4_FLOAT4S * (size_t)node : s.tris + 3 * (size_t)first;
    const float4 q0 = ldg(p), q1 = ldg(p + 1), q2 = ldg(p + 2);
    if (internal) {
        if (wc) wc->nodes++;
        const float4 lx = q0, hx = q1, ly = q2, hy = ldg(p + 3), lz = ldg(p + 4), hz = ldg(p + 5), cd = ldg(p + 6);
        const float inf = RT_FLT_MAX;
        float t0, t1, t2, t3;
        if (!slab(lx.x, hx.x, ly.x, hy.x, lz.x, hz.x, r, hit.t, t0)) t0 = inf;
        if (!slab(lx.y, hx.y, ly.y, hy.y, lz.y, hz.y, r, hit.t, t1)) t1 = inf;
        if (!slab(lx.z, hx.z, ly.z, hy.z, lz.z, hz.z, r, hit.t, t2)) t2 = inf;
        if (!slab(lx.w, hx.w, ly.w, hy.w, lz.w, hz.w, r, hit.t, t3)) t3 = inf;
        int c0 = (int)as_uint(cd.x), c1 = (int)as_uint(cd.y), c2 = (int)as_uint(cd.z), c3 = (int)as_uint(cd.w);
        sort2(t0, c0, t1, c1); sort2(t2, c2, t3, c3); sort2(t0, c0, t2, c2); sort2(t1, c1, t3, c3); sort2(t1, c1, t2, c2);
        if (t0 == inf) return stk.empty() ? RT_DONE : stk.pop();
        if (t3 != inf) { if (!stk.push(c3) && overflow) *overflow = true; }
        if (t2 != inf) { if (!stk.push(c2) && overflow) *overflow = true; }
        if (t1 != inf) { if (!stk.push(c1) && overflow) *overflow = true; }
        return c0;
    }
    const uint32_t count = rt_leaf_count(node);
    float4 r0 = q0, r1 = q1, r2 = q2;
    for (uint32_t k = 0;;) {
        if (wc) wc->tris++;
        float tie[2];
        int rc = tri_test(mk3(r0), mk3(r1), mk3(r2), r.o, r.d, hit.t, hit.beta, hit.gamma, tie);
        if (rc == 2) {
            hit.prim = (int)(first + k);
            found = true;
            if (any_hit) return RT_DONE;
        } else if (rc == 3 && !any_hit && hit.prim >= 0 && hit.prim != (int)(first + k)) {
            if (as_uint(r0.w) < as_uint(ldg(s.tris + 3 * (size_t)hit.prim).w)) {
                hit.prim = (int)(first + k);
                hit.beta = tie[0];
                hit.gamma = tie[1];
            }
        }
        if (++k >= count) break;
        const float4* rec = s.tris + 3 * (size_t)(first + k);
        r0 = ldg(rec); r1 = ldg(rec + 1); r2 = ldg(rec + 2);
    }
    return stk.empty() ? RT_DONE : stk.pop();
}

// hit.t must hold the current upper bound (RT_FLT_MAX for a fresh ray), hit.prim = RT_MISS.
template <bool ANY_HIT>
RT_HD bool bvh_walk(const SceneDev& s, const RayPrep& r, HitRec& hit, WorkCount* wc, bool* overflow) {
    if (s.n_bvh_tris <= 0) return false;
    int entries[RT_STACK_SIZE];
    int depth = 0;
    StackView stk(entries, depth);
    int node = 0;
    bool found = false;
    // the steps the persistent kernels run (one fetch per step for node and leaf alike)
    if (!s.nodes4) {
        while (node != RT_DONE) node = bvh_step_unified(s, r, hit, node, stk, ANY_HIT, found, overflow, wc);
    } else {
        while (node != RT_DONE) node = bvh4_step_unified(s, r, hit, node, stk, ANY_HIT, found, overflow, wc);
    }
    return found;
}

// Linear reference walk used by RT_FLAG_BRUTE_FORCE (and by the structure tests): every triangle.
template <bool ANY_HIT>
RT_HD bool brute_walk(const SceneDev& s, const RayPrep& r, HitRec& hit, WorkCount* wc) {
    bool found = false;
    for (int i = 0; i < s.n_bvh_tris; i++) {
        const float4* rec = s.tris + 3 * (size_t)i;
        float4 r0 = ldg(rec), r1 = ldg(rec + 1), r2 = ldg(rec + 2);
        if (wc) wc->tris++;
        float tie[2];
        int rc = tri_test(mk3(r0), mk3(r1), mk3(r2), r.o, r.d, hit.t, hit.beta, hit.gamma, tie);
        if (rc == 2) {
            hit.prim = i;
            found = true;
            if (ANY_HIT) return true;
        } else if (rc == 3 && !ANY_HIT && hit.prim >= 0 && hit.prim != i) {
            if (as_uint(r0.w) < as_uint(ldg(s.tris + 3 * (size_t)hit.prim).w)) {
                hit.prim = i;
                hit.beta = tie[0];
                hit.gamma = tie[1];
            }
        }
    }
    return found;
}

// The full query.  Direction NaN (a failed refraction, world.cpp:83/:98) misses everything, like
// the reference where every comparison against NaN is false.
template <bool ANY_HIT>
RT_HD bool trace_ray(const SceneDev& s, f3 o, f3 d, bool brute, HitRec& hit, WorkCount* wc, bool* overflow) {
    hit.t = RT_FLT_MAX;
    hit.prim = RT_MISS;
    hit.beta = 0.0f;
    hit.gamma = 0.0f;
    if (!(d.x == d.x && d.y == d.y && d.z == d.z)) return false;
    RayPrep r = prep_ray(o, d);
    bool found = brute ? brute_walk<ANY_HIT>(s, r, hit, wc) : bvh_walk<ANY_HIT>(s, r, hit, wc, overflow);
    if (ANY_HIT && found) return true;
    for (int k = 0; k < s.n_analytic; k++) {
        const AnalyticPrim p = s.analytic[k];
        if (wc) wc->tris++;
        if (analytic_test(p, o, d, hit.t, hit.beta, hit.gamma)) {
            hit.prim = rt_analytic_code(k);
            found = true;
            if (ANY_HIT) return true;
        }
    }
    return found;
}
