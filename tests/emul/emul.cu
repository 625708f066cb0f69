// emul.cu — TEST-ONLY host emulation of the GPU pipeline.
//
// Compiled by nvcc as HOST code: it runs the very same RT_HD functions the kernels call
// (rt_intersect.h, rt_traverse.h, rt_shade.h, rt_bvh.h) serially on the CPU, with std::stable_sort
// standing in for the device radix sort and plain loops for the launch grids.  It exists so that
// the logic (LBVH construction, traversal, shading, bounce bookkeeping, FP32 parity rates) can be
// debugged on a machine without a GPU.  It is never linked into, loaded by or shipped with the
// product library; only tests/test_emulation.py loads it.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/realtrace_b200.h"
#include "../../oracle/oracle_abi.h"
#include "../../realtrace_b200/csrc/rt_bvh.h"
#include "../../realtrace_b200/csrc/rt_shade.h"

namespace {

struct EmulScene {
    std::vector<float4> nodes, nodes4, tris, tri_rgb, materials, lights;
    std::vector<AnalyticPrim> analytic;
    std::vector<uint64_t> keys;
    std::vector<uint32_t> order;
    SceneDev dev{};
    uint32_t n_bvh = 0, n_large = 0;
};

f3 vtx(const float* v, uint32_t i, int k) { const float* p = v + 9 * (size_t)i + 3 * k; return mk3(p[0], p[1], p[2]); }

// ---- experiment only: top-down binned-SAH build over the Morton-sorted triangles (same node format), to
// measure how far the LBVH is from a high-quality tree.  Leaves hold one triangle.
struct SahTmp { Aabb box; f3 ctr; uint32_t k; };
static float half_area(const Aabb& b) { f3 d = b.hi - b.lo; return d.x * d.y + d.y * d.z + d.z * d.x; }
static int sah_build(std::vector<SahTmp>& prims, int lo, int hi, std::vector<float4>& nodes, Aabb& out_box, float pad_abs) {
    // returns child code for the range [lo, hi)
    Aabb box; box.lo = mk3(RT_FLT_MAX, RT_FLT_MAX, RT_FLT_MAX); box.hi = mk3(-RT_FLT_MAX, -RT_FLT_MAX, -RT_FLT_MAX);
    Aabb cb = box;
    for (int i = lo; i < hi; i++) { box = aabb_union(box, prims[i].box); Aabb c; c.lo = c.hi = prims[i].ctr; cb = aabb_union(cb, c); }
    out_box = box;
    if (hi - lo == 1) return rt_leaf_code(prims[lo].k, 1u);
    int best_axis = -1, best_split = -1; float best_cost = RT_FLT_MAX;
    const int NB = 16;
    float ext[3] = {cb.hi.x - cb.lo.x, cb.hi.y - cb.lo.y, cb.hi.z - cb.lo.z};
    float clo[3] = {cb.lo.x, cb.lo.y, cb.lo.z};
    for (int ax = 0; ax < 3; ax++) {
        if (!(ext[ax] > 0)) continue;
        Aabb bb[NB]; int cnt[NB];
        for (int b = 0; b < NB; b++) { bb[b].lo = mk3(RT_FLT_MAX, RT_FLT_MAX, RT_FLT_MAX); bb[b].hi = mk3(-RT_FLT_MAX, -RT_FLT_MAX, -RT_FLT_MAX); cnt[b] = 0; }
        for (int i = lo; i < hi; i++) {
            float c = ax == 0 ? prims[i].ctr.x : ax == 1 ? prims[i].ctr.y : prims[i].ctr.z;
            int b = (int)((c - clo[ax]) / ext[ax] * NB); if (b >= NB) b = NB - 1; if (b < 0) b = 0;
            bb[b] = aabb_union(bb[b], prims[i].box); cnt[b]++;
        }
        float la[NB], ra[NB]; int lc[NB], rc[NB];
        Aabb acc; acc.lo = mk3(RT_FLT_MAX, RT_FLT_MAX, RT_FLT_MAX); acc.hi = mk3(-RT_FLT_MAX, -RT_FLT_MAX, -RT_FLT_MAX); int c = 0;
        for (int b = 0; b < NB; b++) { if (cnt[b]) acc = aabb_union(acc, bb[b]); c += cnt[b]; la[b] = c ? half_area(acc) : 0; lc[b] = c; }
        acc.lo = mk3(RT_FLT_MAX, RT_FLT_MAX, RT_FLT_MAX); acc.hi = mk3(-RT_FLT_MAX, -RT_FLT_MAX, -RT_FLT_MAX); c = 0;
        for (int b = NB - 1; b >= 0; b--) { if (cnt[b]) acc = aabb_union(acc, bb[b]); c += cnt[b]; ra[b] = c ? half_area(acc) : 0; rc[b] = c; }
        for (int b = 0; b < NB - 1; b++) {
            if (lc[b] == 0 || rc[b + 1] == 0) continue;
            float cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
            if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = b; }
        }
    }
    int mid;
    if (best_axis < 0) mid = (lo + hi) / 2;
    else {
        auto it = std::partition(prims.begin() + lo, prims.begin() + hi, [&](const SahTmp& p) {
            float c = best_axis == 0 ? p.ctr.x : best_axis == 1 ? p.ctr.y : p.ctr.z;
            int b = (int)((c - clo[best_axis]) / ext[best_axis] * NB); if (b >= NB) b = NB - 1; if (b < 0) b = 0;
            return b <= best_split; });
        mid = (int)(it - prims.begin());
        if (mid == lo || mid == hi) mid = (lo + hi) / 2;
    }
    int me = (int)(nodes.size() / 4);
    nodes.resize(nodes.size() + 4);
    Aabb L, R;
    int cl = sah_build(prims, lo, mid, nodes, L, pad_abs), cr = sah_build(prims, mid, hi, nodes, R, pad_abs);
    auto padded = [&](Aabb b) { Aabb r = aabb_pad(b); r.lo = r.lo - mk3(pad_abs, pad_abs, pad_abs); r.hi = r.hi + mk3(pad_abs, pad_abs, pad_abs); return r; };
    Aabb Lp = padded(L), Rp = padded(R);
    nodes[4 * me] = make_float4(Lp.lo.x, Lp.hi.x, Lp.lo.y, Lp.hi.y);
    nodes[4 * me + 1] = make_float4(Rp.lo.x, Rp.hi.x, Rp.lo.y, Rp.hi.y);
    nodes[4 * me + 2] = make_float4(Lp.lo.z, Lp.hi.z, Rp.lo.z, Rp.hi.z);
    nodes[4 * me + 3] = make_float4(as_float((uint32_t)cl), as_float((uint32_t)cr), 0, 0);
    return me;
}

// ---- experiment only: PLOC (Meister & Bittner 2018) over the Morton-sorted leaves — bottom-up merging of
// mutual nearest neighbours (surface area of the union) inside a window of +-radius positions.  Same node
// format; node 0 is the root (indices are handed out from the top down as merges happen last for the root).
struct PlocCluster { Aabb box; int code; };
static void ploc_build(std::vector<PlocCluster> c, int radius, std::vector<float4>& nodes, float pad_abs) {
    int n = (int)c.size();
    nodes.assign(4 * (size_t)(n - 1), make_float4(0, 0, 0, 0));
    int next = n - 2;
    auto padded = [&](Aabb b) { Aabb r = aabb_pad(b); r.lo = r.lo - mk3(pad_abs, pad_abs, pad_abs); r.hi = r.hi + mk3(pad_abs, pad_abs, pad_abs); return r; };
    std::vector<int> nn(n);
    std::vector<PlocCluster> out;
    while ((int)c.size() > 1) {
        int m = (int)c.size();
        for (int i = 0; i < m; i++) {
            float best = RT_FLT_MAX; int bj = -1;
            for (int j = std::max(0, i - radius); j <= std::min(m - 1, i + radius); j++) {
                if (j == i) continue;
                float a = half_area(aabb_union(c[i].box, c[j].box));
                if (a < best) { best = a; bj = j; }
            }
            nn[i] = bj;
        }
        out.clear();
        for (int i = 0; i < m; i++) {
            int j = nn[i];
            if (nn[j] == i) {
                if (i < j) {
                    int me = next--;
                    Aabb Lp = padded(c[i].box), Rp = padded(c[j].box);
                    nodes[4 * (size_t)me] = make_float4(Lp.lo.x, Lp.hi.x, Lp.lo.y, Lp.hi.y);
                    nodes[4 * (size_t)me + 1] = make_float4(Rp.lo.x, Rp.hi.x, Rp.lo.y, Rp.hi.y);
                    nodes[4 * (size_t)me + 2] = make_float4(Lp.lo.z, Lp.hi.z, Rp.lo.z, Rp.hi.z);
                    nodes[4 * (size_t)me + 3] = make_float4(as_float((uint32_t)c[i].code), as_float((uint32_t)c[j].code), 0, 0);
                    PlocCluster p; p.box = aabb_union(c[i].box, c[j].box); p.code = me;
                    out.push_back(p);
                }
            } else out.push_back(c[i]);
        }
        c.swap(out);
    }
}

void build(const oracle_scene* in, int leaf_size, EmulScene& S) {
    uint32_t n = in->n_tri;
    // analytic list: spheres, planes, cylinders (object ids default to after the triangles)
    uint32_t next = n;
    for (uint32_t i = 0; i < in->n_sph; i++, next++) {
        AnalyticPrim p{}; const float* q = in->sph + 4 * (size_t)i;
        p.kind = RT_KIND_SPHERE; p.material = in->sph_material[i]; p.object_id = in->sph_object_id ? in->sph_object_id[i] : next;
        p.a = make_float4(q[0], q[1], q[2], q[3]); S.analytic.push_back(p);
    }
    for (uint32_t i = 0; i < in->n_pln; i++, next++) {
        AnalyticPrim p{}; const float* q = in->pln + 12 * (size_t)i;
        p.kind = RT_KIND_PLANE; p.material = in->pln_material[i]; p.object_id = in->pln_object_id ? in->pln_object_id[i] : next;
        p.a = make_float4(q[0], q[1], q[2], 0); p.b = make_float4(q[3], q[4], q[5], 0);
        p.c = make_float4(q[6], q[7], q[8], 0); p.d = make_float4(q[9], q[10], q[11], 0); S.analytic.push_back(p);
    }
    for (uint32_t i = 0; i < in->n_cyl; i++, next++) {
        AnalyticPrim p{}; const float* q = in->cyl + 7 * (size_t)i;
        p.kind = RT_KIND_CYLINDER; p.material = in->cyl_material[i]; p.object_id = in->cyl_object_id ? in->cyl_object_id[i] : next;
        p.a = make_float4(q[0], q[1], q[2], q[3]); p.b = make_float4(q[4], q[5], q[6], 0); S.analytic.push_back(p);
    }
    for (uint32_t i = 0; i < in->n_materials; i++) {
        const oracle_material& m = in->materials[i];
        S.materials.push_back(make_float4(m.color[0], m.color[1], m.color[2], m.ka));
        S.materials.push_back(make_float4(m.kd, m.ks, m.kr, m.kt));
        S.materials.push_back(make_float4(m.eta, as_float(m.barycentric ? 1u : 0u), 0, 0));
    }
    for (uint32_t i = 0; i < in->n_lights; i++) {
        const float* p = in->lights + 6 * (size_t)i;
        S.lights.push_back(make_float4(p[0], p[1], p[2], 0)); S.lights.push_back(make_float4(p[3], p[4], p[5], 0));
    }
    if (in->tri_rgb)
        for (uint32_t i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) { const float* p = in->tri_rgb + 9 * (size_t)i + 3 * k; S.tri_rgb.push_back(make_float4(p[0], p[1], p[2], 0)); }

    // k_scene_bounds
    Aabb sb, cb;
    sb.lo = cb.lo = mk3(RT_FLT_MAX, RT_FLT_MAX, RT_FLT_MAX);
    sb.hi = cb.hi = mk3(-RT_FLT_MAX, -RT_FLT_MAX, -RT_FLT_MAX);
    for (uint32_t i = 0; i < n; i++) {
        Aabb b = tri_aabb(vtx(in->tri_v, i, 0), vtx(in->tri_v, i, 1), vtx(in->tri_v, i, 2));
        sb = aabb_union(sb, b);
        f3 c = (b.lo + b.hi) * 0.5f;
        Aabb cc; cc.lo = c; cc.hi = c;
        cb = aabb_union(cb, cc);
    }
    float amax = 0;
    for (float v : {sb.lo.x, sb.lo.y, sb.lo.z, sb.hi.x, sb.hi.y, sb.hi.z}) amax = fmaxf(amax, fabsf(v));
    float pad_abs = amax * 4.0e-6f;
    // k_morton (+ oversized-triangle classification)
    float large_frac = n >= 64 ? 0.25f : RT_FLT_MAX;
    for (int attempt = 0; attempt < 2; attempt++) {
        S.keys.assign(n, 0); S.order.resize(n); S.n_large = 0;
        float scene_ext = fmaxf(sb.hi.x - sb.lo.x, fmaxf(sb.hi.y - sb.lo.y, sb.hi.z - sb.lo.z));
        f3 ext = cb.hi - cb.lo;
        float emax = fmaxf(ext.x, fmaxf(ext.y, ext.z));
        float iu = emax > 0 ? 1.0f / emax : 0.0f;
        f3 inv = mk3(iu, iu, iu);   // cubic cells: one scale for all axes
        for (uint32_t i = 0; i < n; i++) {
            Aabb b = tri_aabb(vtx(in->tri_v, i, 0), vtx(in->tri_v, i, 1), vtx(in->tri_v, i, 2));
            float te = fmaxf(b.hi.x - b.lo.x, fmaxf(b.hi.y - b.lo.y, b.hi.z - b.lo.z));
            if (te > large_frac * scene_ext) { S.keys[i] = ~0ull; S.n_large++; }
            else S.keys[i] = morton63((b.lo + b.hi) * 0.5f, cb.lo, inv);
            S.order[i] = i;
        }
        if (S.n_large <= 16) break;
        large_frac = RT_FLT_MAX;
    }
    // sort (stable, by key)
    std::stable_sort(S.order.begin(), S.order.end(), [&](uint32_t a, uint32_t b) { return S.keys[a] < S.keys[b]; });
    std::vector<uint64_t> sk(n);
    for (uint32_t k = 0; k < n; k++) sk[k] = S.keys[S.order[k]];
    S.keys.swap(sk);
    uint32_t nb = S.n_bvh = n - S.n_large;
    // k_tri_records
    S.tris.resize(3 * (size_t)n);
    for (uint32_t k = 0; k < n; k++) {
        uint32_t i = S.order[k];
        f3 a = vtx(in->tri_v, i, 0), b = vtx(in->tri_v, i, 1), c = vtx(in->tri_v, i, 2);
        f3 e1 = a - b, e2 = a - c;
        uint32_t obj = in->tri_object_id ? in->tri_object_id[i] : i;
        S.tris[3 * k] = make_float4(a.x, a.y, a.z, as_float(obj));
        S.tris[3 * k + 1] = make_float4(e1.x, e1.y, e1.z, as_float(in->tri_material[i]));
        S.tris[3 * k + 2] = make_float4(e2.x, e2.y, e2.z, as_float(i));
    }
    for (uint32_t k = 0; k < S.n_large; k++) {   // k_emit_large
        AnalyticPrim p{};
        const float4* r = &S.tris[3 * (size_t)(nb + k)];
        p.kind = RT_KIND_TRIANGLE; p.object_id = as_uint(r[0].w); p.material = as_uint(r[1].w); p.tri_index = as_uint(r[2].w);
        p.a = r[0]; p.b = r[1]; p.c = r[2];
        S.analytic.push_back(p);
    }
    auto padded = [&](Aabb b) { Aabb r = aabb_pad(b); r.lo = r.lo - mk3(pad_abs, pad_abs, pad_abs); r.hi = r.hi + mk3(pad_abs, pad_abs, pad_abs); return r; };
    if (nb >= 2) {
        std::vector<KarrasNode> kn(nb - 1);
        std::vector<int> leaf_parent(nb), node_parent(nb - 1, -1);
        for (int i = 0; i < (int)nb - 1; i++) {
            kn[i] = karras_node(S.keys.data(), (int)nb, i);
            if (kn[i].left < 0) leaf_parent[~kn[i].left] = i; else node_parent[kn[i].left] = i;
            if (kn[i].right < 0) leaf_parent[~kn[i].right] = i; else node_parent[kn[i].right] = i;
        }
        node_parent[0] = -1;
        std::vector<Aabb> box(2 * (size_t)nb - 1);
        std::vector<int> visit(nb - 1, 0);
        S.nodes.assign(4 * (size_t)(nb - 1), make_float4(0, 0, 0, 0));
        auto code = [&](int child) {
            if (child < 0) return rt_leaf_code((uint32_t)~child, 1u);
            int cnt = kn[child].last - kn[child].first + 1;
            return cnt <= leaf_size ? rt_leaf_code((uint32_t)kn[child].first, (uint32_t)cnt) : child;
        };
        for (uint32_t k = 0; k < nb; k++) {
            uint32_t i = S.order[k];
            box[(nb - 1) + k] = tri_aabb(vtx(in->tri_v, i, 0), vtx(in->tri_v, i, 1), vtx(in->tri_v, i, 2));
            int cur = leaf_parent[k];
            while (cur >= 0) {
                if (visit[cur]++ == 0) break;
                const KarrasNode& nd = kn[cur];
                Aabb L = box[nd.left < 0 ? (nb - 1) + ~nd.left : nd.left], R = box[nd.right < 0 ? (nb - 1) + ~nd.right : nd.right];
                box[cur] = aabb_union(L, R);
                Aabb Lp = padded(L), Rp = padded(R);
                float4* o = &S.nodes[4 * (size_t)cur];
                o[0] = make_float4(Lp.lo.x, Lp.hi.x, Lp.lo.y, Lp.hi.y);
                o[1] = make_float4(Rp.lo.x, Rp.hi.x, Rp.lo.y, Rp.hi.y);
                o[2] = make_float4(Lp.lo.z, Lp.hi.z, Rp.lo.z, Rp.hi.z);
                o[3] = make_float4(as_float((uint32_t)code(nd.left)), as_float((uint32_t)code(nd.right)), as_float((uint32_t)nd.first), as_float((uint32_t)nd.last));
                cur = node_parent[cur];
            }
        }
    } else if (nb == 1) {
        uint32_t i = S.order[0];
        Aabb p = padded(tri_aabb(vtx(in->tri_v, i, 0), vtx(in->tri_v, i, 1), vtx(in->tri_v, i, 2)));
        S.nodes = {make_float4(p.lo.x, p.hi.x, p.lo.y, p.hi.y), make_float4(RT_FLT_MAX, RT_FLT_MAX, RT_FLT_MAX, RT_FLT_MAX),
                   make_float4(p.lo.z, p.hi.z, RT_FLT_MAX, RT_FLT_MAX),
                   make_float4(as_float((uint32_t)rt_leaf_code(0, 1)), as_float((uint32_t)RT_EMPTY_CODE), 0, 0)};
    }
    if (getenv("EMUL_SAH") && nb >= 2) {
        std::vector<SahTmp> prims(nb);
        for (uint32_t k = 0; k < nb; k++) {
            uint32_t i = S.order[k];
            prims[k].box = tri_aabb(vtx(in->tri_v, i, 0), vtx(in->tri_v, i, 1), vtx(in->tri_v, i, 2));
            prims[k].ctr = (prims[k].box.lo + prims[k].box.hi) * 0.5f;
            prims[k].k = k;
        }
        std::vector<float4> nodes;
        Aabb rb;
        sah_build(prims, 0, (int)nb, nodes, rb, pad_abs);
        S.nodes.swap(nodes);
    }
    if (getenv("EMUL_PLOC") && nb >= 2) {
        std::vector<PlocCluster> cl(nb);
        for (uint32_t k = 0; k < nb; k++) {
            uint32_t i = S.order[k];
            cl[k].box = tri_aabb(vtx(in->tri_v, i, 0), vtx(in->tri_v, i, 1), vtx(in->tri_v, i, 2));
            cl[k].code = rt_leaf_code(k, 1u);
        }
        std::vector<float4> nodes;
        ploc_build(cl, atoi(getenv("EMUL_PLOC")) > 0 ? atoi(getenv("EMUL_PLOC")) : 8, nodes, pad_abs);
        S.nodes.swap(nodes);
    }
    if (getenv("EMUL_WIDE") && !S.nodes.empty()) {      // 4-wide view (k_collapse4 of the GPU build)
        size_t nn = S.nodes.size() / 4;
        S.nodes4.resize(8 * nn);
        for (size_t i = 0; i < nn; i++) build_wide_node(S.nodes.data(), (int)i, &S.nodes4[8 * i]);
    }
    SceneDev& d = S.dev;
    d.nodes4 = S.nodes4.empty() ? nullptr : S.nodes4.data();
    d.nodes = S.nodes.data(); d.tris = S.tris.data(); d.tri_rgb = S.tri_rgb.empty() ? nullptr : S.tri_rgb.data();
    d.analytic = S.analytic.data(); d.materials = S.materials.data(); d.lights = S.lights.data();
    d.n_bvh_tris = (int)nb; d.n_nodes = nb >= 2 ? (int)nb - 1 : (int)nb; d.n_analytic = (int)S.analytic.size();
    d.n_lights = (int)in->n_lights;
    for (int k = 0; k < 3; k++) { d.ambient[k] = in->ambient[k]; d.background[k] = in->background[k]; }
}

struct Pending { f3 o, d, w; int level; };

long long to_fixed(float x) {
    if (!(x == x)) x = 0.0f;
    x = fminf(fmaxf(x, -1048576.0f), 1048576.0f);
    return llrintf(x * 4294967296.0f);
}

}  // namespace

extern "C" int emul_render(const oracle_scene* in, const rt_camera* cam, int max_depth, int leaf_size, int brute,
                           uint8_t* rgb, int32_t* prim_id, float* t_hit, uint64_t* counts /*primary, shadow, secondary, nodes, tris*/) {
    EmulScene S;
    build(in, leaf_size, S);
    const SceneDev& s = S.dev;
    const int W = cam->width, H = cam->height;
    uint64_t n_shadow = 0, n_secondary = 0;
    WorkCount wc{0, 0};
    uint64_t nodes = 0, tris = 0;
    bool overflow = false;
    f3 bg = mk3(s.background[0], s.background[1], s.background[2]);
    std::vector<Pending> stack;
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            // primary_ray of render.cu
            float xw = (float)((double)cam->aspect * (i - W / 2.0 + 0.5) / W), yw = (float)((j - H / 2.0 + 0.5) / H);
            double dd[3];
            for (int k = 0; k < 3; k++) dd[k] = -(double)cam->w[k] * (double)cam->focal_distance + (double)cam->u[k] * (double)xw + (double)cam->v[k] * (double)yw;
            double l = std::sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
            for (int k = 0; k < 3; k++) dd[k] /= l;
            l = std::sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
            Pending r0;
            r0.d = mk3((float)(dd[0] / l), (float)(dd[1] / l), (float)(dd[2] / l));
            r0.o = mk3(cam->pos[0], cam->pos[1], cam->pos[2]);
            r0.w = mk3(1, 1, 1); r0.level = 0;
            long long acc[3] = {0, 0, 0};
            stack.clear();
            stack.push_back(r0);
            bool first = true;
            while (!stack.empty()) {
                Pending r = stack.back();
                stack.pop_back();
                if (!first) n_secondary++;
                HitRec h;
                wc.nodes = wc.tris = 0;
                bool found = trace_ray<false>(s, r.o, r.d, brute != 0, h, &wc, &overflow);
                nodes += wc.nodes; tris += wc.tris;
                if (first) {
                    size_t at = (size_t)i + (size_t)j * W;
                    if (prim_id) prim_id[at] = !found ? -1 : (h.prim >= 0 ? (int)as_uint(s.tris[3 * (size_t)h.prim].w) : (int)s.analytic[rt_analytic_index(h.prim)].object_id);
                    if (t_hit) t_hit[at] = found ? h.t : RT_FLT_MAX;
                    first = false;
                }
                f3 add;
                if (!found) add = r.w * bg;
                else {
                    ShadeOut out;
                    auto any_hit = [&](f3 so, f3 sd) { HitRec sh; wc.nodes = wc.tris = 0; bool f = trace_ray<true>(s, so, sd, brute != 0, sh, &wc, &overflow); nodes += wc.nodes; tris += wc.tris; return f; };
                    shade_hit(s, r.o, r.d, r.level, h, max_depth, any_hit, out);
                    n_shadow += out.shadow_rays;
                    add = r.w * (out.local + out.bg_weight * bg);
                    for (int c = 0; c < out.n_children; c++) stack.push_back({out.child[c].o, out.child[c].d, r.w * out.child[c].w, out.child[c].level});
                }
                acc[0] += to_fixed(add.x); acc[1] += to_fixed(add.y); acc[2] += to_fixed(add.z);
            }
            for (int k = 0; k < 3; k++) {
                double c = (double)acc[k] * (1.0 / 4294967296.0);
                if (c > 1.0) c = 1.0;
                if (c < 0.0) c = 0.0;
                rgb[((size_t)i + (size_t)j * W) * 3 + k] = (uint8_t)(255.0 * c);
            }
        }
    if (counts) { counts[0] = (uint64_t)W * H; counts[1] = n_shadow; counts[2] = n_secondary; counts[3] = nodes; counts[4] = tris; }
    return overflow ? -5 : 0;
}

extern "C" int emul_trace_rays(const oracle_scene* in, int leaf_size, int brute, const float* rays, uint32_t n, int32_t* prim_id, float* t_hit) {
    EmulScene S;
    build(in, leaf_size, S);
    const SceneDev& s = S.dev;
    bool overflow = false;
    for (uint32_t r = 0; r < n; r++) {
        const float* p = rays + 6 * (size_t)r;
        double x = p[3], y = p[4], z = p[5], l = std::sqrt(x * x + y * y + z * z);
        HitRec h;
        bool found = trace_ray<false>(s, mk3(p[0], p[1], p[2]), mk3((float)(x / l), (float)(y / l), (float)(z / l)), brute != 0, h, nullptr, &overflow);
        prim_id[r] = !found ? -1 : (h.prim >= 0 ? (int)as_uint(s.tris[3 * (size_t)h.prim].w) : (int)s.analytic[rt_analytic_index(h.prim)].object_id);
        t_hit[r] = found ? h.t : RT_FLT_MAX;
    }
    return overflow ? -5 : 0;
}

// Experiment (round-2 preparation): a nearest-hit walk that takes `switch_after` steps on the binary tree and then
// continues on the 4-wide view of the same tree with the SAME stack — both views index the same nodes, so a pushed
// or current node code means the same subtree in either.  Returns per ray the hit and the number of dependent steps.
extern "C" int emul_hybrid_walk(const oracle_scene* in, int switch_after, const float* rays, uint32_t n, int32_t* prim,
                                float* t_hit, uint32_t* steps) {
    EmulScene S;
    setenv("EMUL_WIDE", "1", 1);
    build(in, 1, S);
    unsetenv("EMUL_WIDE");
    const SceneDev& s = S.dev;
    bool overflow = false;
    for (uint32_t k = 0; k < n; k++) {
        const float* p = rays + 6 * (size_t)k;
        double x = p[3], y = p[4], z = p[5], l = std::sqrt(x * x + y * y + z * z);
        RayPrep r = prep_ray(mk3(p[0], p[1], p[2]), mk3((float)(x / l), (float)(y / l), (float)(z / l)));
        HitRec hit;
        hit.t = RT_FLT_MAX; hit.prim = RT_MISS; hit.beta = hit.gamma = 0.0f;
        int stack[RT_STACK_SIZE], sp = 0, node = s.n_bvh_tris > 0 ? 0 : RT_DONE;
        uint32_t st = 0;
        while (node != RT_DONE) {
            if (rt_is_internal(node)) {
                node = ((int)st >= switch_after) ? bvh4_node_step(s, r, hit.t, node, stack, sp, &overflow)
                                                 : bvh_node_step(s, r, hit.t, node, stack, sp, &overflow);
            } else {
                leaf_test(s, node, r, hit, false, nullptr);
                node = sp ? stack[--sp] : RT_DONE;
            }
            st++;
        }
        prim[k] = hit.prim;
        t_hit[k] = hit.t;
        steps[k] = st;
    }
    return overflow ? -5 : 0;
}

// Analysis (round-2 preparation): SIMD efficiency of one warp = one 8x4 pixel block under different loop
// structures, from the exact step sequences (N = node step, L = leaf test) of its 32 primary rays and their shadow
// rays (walked one after the other by the same lane, like the fused kernel).  out[policy][0..3] = executions of the
// node branch, of the leaf branch, lane-steps spent in node steps, lane-steps spent in leaf tests.
//   policy 0: if-if — per iteration every live lane takes its next step; the warp issues the node branch if any lane
//             is at a node and the leaf branch if any lane is at a leaf (the shipped loop)
//   policy 1: while-while — all lanes run node steps until each has reached a leaf (or finished), then all run leaf
//             tests until each is back at a node
//   policy 2: if-if with the leaf branch postponed until at least `quorum` lanes wait at a leaf (or nobody can take a
//             node step)
static void step_sequence(const SceneDev& s, f3 o, f3 d, bool any_hit, std::vector<uint8_t>& seq, HitRec& hit, bool& found) {
    RayPrep r = prep_ray(o, d);
    hit.t = RT_FLT_MAX; hit.prim = RT_MISS; hit.beta = hit.gamma = 0.0f;
    found = false;
    bool overflow = false;
    int stack[RT_STACK_SIZE], sp = 0, node = s.n_bvh_tris > 0 ? 0 : RT_DONE;
    while (node != RT_DONE) {
        if (rt_is_internal(node)) { seq.push_back(0); node = bvh_node_step(s, r, hit.t, node, stack, sp, &overflow); }
        else {
            seq.push_back(1);
            if (leaf_test(s, node, r, hit, any_hit, nullptr)) { found = true; if (any_hit) break; }
            node = sp ? stack[--sp] : RT_DONE;
        }
    }
    for (int k = 0; k < s.n_analytic && !(any_hit && found); k++) {
        const AnalyticPrim p = s.analytic[k];
        if (analytic_test(p, o, d, hit.t, hit.beta, hit.gamma)) { hit.prim = rt_analytic_code(k); found = true; }
    }
}
// Variant walk for the model: leaf children are tested inside the node step that finds their box hit ("eager
// leaves"; only internal nodes are ever pushed).  seq gets one entry per iteration: the number of triangle tests
// made in it (0, 1 or 2).  Nearest hits must equal the shipped walk's.
static void eager_sequence(const SceneDev& s, f3 o, f3 d, bool any_hit, std::vector<uint8_t>& seq, HitRec& hit, bool& found) {
    RayPrep r = prep_ray(o, d);
    hit.t = RT_FLT_MAX; hit.prim = RT_MISS; hit.beta = hit.gamma = 0.0f;
    found = false;
    int stack[RT_STACK_SIZE], sp = 0, node = s.n_bvh_tris > 1 ? 0 : RT_DONE;
    if (s.n_bvh_tris == 1) {   // single-leaf tree: the root record holds the leaf
        seq.push_back(1);
        if (leaf_test(s, (int)as_uint(s.nodes[3].x), r, hit, any_hit, nullptr)) found = true;
    }
    while (node != RT_DONE) {
        const float4* n = s.nodes + RT_NODE_FLOAT4S * (size_t)node;
        float4 n0 = n[0], n1 = n[1], n2 = n[2], n3 = n[3];
        float t0, t1;
        bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, r, hit.t, t0);
        bool h1 = slab(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, r, hit.t, t1);
        int c0 = (int)as_uint(n3.x), c1 = (int)as_uint(n3.y);
        if (h0 && h1 && t1 < t0) { int tc = c0; c0 = c1; c1 = tc; float tt = t0; t0 = t1; t1 = tt; }
        else if (!h0 && h1) { c0 = c1; t0 = t1; h0 = true; h1 = false; }
        uint8_t tests = 0;
        int next_a = RT_DONE, next_b = RT_DONE;     // internal children still to visit, near first
        bool stop = false;
        if (h0) {
            if (c0 < 0) { tests++; if (leaf_test(s, c0, r, hit, any_hit, nullptr)) { found = true; stop = any_hit; } }
            else next_a = c0;
        }
        if (h1 && !stop && t1 <= hit.t) {
            if (c1 < 0) { tests++; if (leaf_test(s, c1, r, hit, any_hit, nullptr)) { found = true; stop = any_hit; } }
            else { if (next_a == RT_DONE) next_a = c1; else next_b = c1; }
        }
        seq.push_back(tests);
        if (stop) break;
        if (next_a != RT_DONE) { if (next_b != RT_DONE && sp < RT_STACK_SIZE) stack[sp++] = next_b; node = next_a; }
        else node = sp ? stack[--sp] : RT_DONE;
    }
    for (int k = 0; k < s.n_analytic && !(any_hit && found); k++) {
        const AnalyticPrim p = s.analytic[k];
        if (analytic_test(p, o, d, hit.t, hit.beta, hit.gamma)) { hit.prim = rt_analytic_code(k); found = true; }
    }
}
// out: [0] iterations (warp), [1] leaf sub-branch executions, [2] lane-iterations, [3] lane-tests, [4] rays whose
// nearest hit differs from the shipped walk's, [5] rays, [6] longest lane chain (iterations), [7] triangle tests of the
// shipped walk for comparison
extern "C" int emul_simd_model_eager(const oracle_scene* in, const rt_camera* cam, double* out /*8*/) {
    EmulScene S;
    build(in, 1, S);
    const SceneDev& s = S.dev;
    const int W = cam->width, H = cam->height;
    for (int k = 0; k < 8; k++) out[k] = 0.0;
    std::vector<uint8_t> seq[32], ref;
    f3 o = mk3(cam->pos[0], cam->pos[1], cam->pos[2]);
    for (int by = 0; by + 4 <= H; by += 4)
        for (int bx = 0; bx + 8 <= W; bx += 8) {
            for (int l = 0; l < 32; l++) {
                int i = bx + (l & 7), j = by + (l >> 3);
                seq[l].clear();
                float xw = (float)((double)cam->aspect * (i - W / 2.0 + 0.5) / W), yw = (float)((j - H / 2.0 + 0.5) / H);
                double dd[3];
                for (int k = 0; k < 3; k++) dd[k] = -(double)cam->w[k] * (double)cam->focal_distance + (double)cam->u[k] * (double)xw + (double)cam->v[k] * (double)yw;
                double len = std::sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
                f3 d = mk3((float)(dd[0] / len), (float)(dd[1] / len), (float)(dd[2] / len));
                HitRec h, h2; bool found, f2;
                eager_sequence(s, o, d, false, seq[l], h, found);
                ref.clear();
                step_sequence(s, o, d, false, ref, h2, f2);
                for (uint8_t v : ref) out[7] += v;
                out[5] += 1;
                if (found != f2 || (found && (h.prim != h2.prim || h.t != h2.t))) out[4] += 1;
                if (found)
                    for (int li = 0; li < s.n_lights; li++) {
                        f3 P = fma3(d, h.t, o), toL = mk3(s.lights[2 * li]) - P;
                        HitRec sh; bool sf;
                        eager_sequence(s, fma3(toL, 0.01f, P), normalize(toL), true, seq[l], sh, sf);
                        ref.clear();
                        step_sequence(s, fma3(toL, 0.01f, P), normalize(toL), true, ref, h2, f2);
                        for (uint8_t v : ref) out[7] += v;
                        if (sf != f2) out[4] += 1;
                    }
            }
            size_t maxlen = 0;
            for (int l = 0; l < 32; l++) { maxlen = std::max(maxlen, seq[l].size()); out[2] += (double)seq[l].size(); for (uint8_t v : seq[l]) out[3] += v; }
            out[6] = std::max(out[6], (double)maxlen);
            for (size_t it = 0; it < maxlen; it++) {
                int mt = 0;
                for (int l = 0; l < 32; l++) if (it < seq[l].size()) mt = std::max(mt, (int)seq[l][it]);
                out[0] += 1;
                out[1] += mt;
            }
        }
    return 0;
}

// Refill policy model: ONE warp works through the whole frame (8x4 blocks in raster order, fused primary + shadow
// sequences per ray, if-if loop); idle lanes take the next rays of the stream as soon as at least `threshold` lanes
// are idle (32 = whole batches, the shipped setting for primary rays).  out: [0] node-branch executions, [1]
// leaf-branch executions, [2] refill events, [3] lane-steps.
extern "C" int emul_simd_refill_model(const oracle_scene* in, const rt_camera* cam, int threshold, double* out /*4*/) {
    EmulScene S;
    build(in, 1, S);
    const SceneDev& s = S.dev;
    const int W = cam->width, H = cam->height;
    for (int k = 0; k < 4; k++) out[k] = 0.0;
    f3 o = mk3(cam->pos[0], cam->pos[1], cam->pos[2]);
    std::vector<uint8_t> seq[32];
    size_t pos[32];
    for (int l = 0; l < 32; l++) pos[l] = 0;
    // the stream: pixel k of the frame in 8x4-block raster order
    const int bxn = W / 8, byn = H / 4;
    const long total = (long)bxn * byn * 32;
    long next = 0;
    auto fetch = [&](int l) {
        long k = next++;
        long b = k / 32; int q = (int)(k % 32);
        int i = (int)(b % bxn) * 8 + (q & 7), j = (int)(b / bxn) * 4 + (q >> 3);
        seq[l].clear(); pos[l] = 0;
        float xw = (float)((double)cam->aspect * (i - W / 2.0 + 0.5) / W), yw = (float)((j - H / 2.0 + 0.5) / H);
        double dd[3];
        for (int c = 0; c < 3; c++) dd[c] = -(double)cam->w[c] * (double)cam->focal_distance + (double)cam->u[c] * (double)xw + (double)cam->v[c] * (double)yw;
        double len = std::sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
        f3 d = mk3((float)(dd[0] / len), (float)(dd[1] / len), (float)(dd[2] / len));
        HitRec h; bool found;
        step_sequence(s, o, d, false, seq[l], h, found);
        if (found)
            for (int li = 0; li < s.n_lights; li++) {
                f3 P = fma3(d, h.t, o), toL = mk3(s.lights[2 * li]) - P;
                HitRec sh; bool sf;
                step_sequence(s, fma3(toL, 0.01f, P), normalize(toL), true, seq[l], sh, sf);
            }
    };
    for (;;) {
        int idle = 0;
        for (int l = 0; l < 32; l++) if (pos[l] >= seq[l].size()) idle++;
        if (next < total && (idle >= threshold || idle == 32)) {
            out[2] += 1;
            for (int l = 0; l < 32 && next < total; l++) if (pos[l] >= seq[l].size()) fetch(l);
        }
        int nn = 0, nl = 0;
        for (int l = 0; l < 32; l++) if (pos[l] < seq[l].size()) { if (seq[l][pos[l]]) nl++; else nn++; }
        if (!nn && !nl) { if (next >= total) break; continue; }
        if (nn) out[0] += 1;
        if (nl) out[1] += 1;
        out[3] += nn + nl;
        for (int l = 0; l < 32; l++) if (pos[l] < seq[l].size()) pos[l]++;
    }
    return 0;
}

extern "C" int emul_simd_model(const oracle_scene* in, const rt_camera* cam, int quorum, double* out /*3 x 4*/) {
    EmulScene S;
    build(in, 1, S);
    const SceneDev& s = S.dev;
    const int W = cam->width, H = cam->height;
    for (int k = 0; k < 12; k++) out[k] = 0.0;
    std::vector<uint8_t> seq[32];
    f3 o = mk3(cam->pos[0], cam->pos[1], cam->pos[2]);
    const int bw = getenv("EMUL_BLOCK_W") ? atoi(getenv("EMUL_BLOCK_W")) : 8, bh = 32 / bw;   // pixel block of a warp
    for (int by = 0; by + bh <= H; by += bh)
        for (int bx = 0; bx + bw <= W; bx += bw) {
            for (int l = 0; l < 32; l++) {
                int i = bx + (l % bw), j = by + (l / bw);
                seq[l].clear();
                float xw = (float)((double)cam->aspect * (i - W / 2.0 + 0.5) / W), yw = (float)((j - H / 2.0 + 0.5) / H);
                double dd[3];
                for (int k = 0; k < 3; k++) dd[k] = -(double)cam->w[k] * (double)cam->focal_distance + (double)cam->u[k] * (double)xw + (double)cam->v[k] * (double)yw;
                double len = std::sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
                f3 d = mk3((float)(dd[0] / len), (float)(dd[1] / len), (float)(dd[2] / len));
                HitRec h; bool found;
                step_sequence(s, o, d, false, seq[l], h, found);
                if (found)
                    for (int li = 0; li < s.n_lights; li++) {
                        f3 P = fma3(d, h.t, o), toL = mk3(s.lights[2 * li]) - P;
                        HitRec sh; bool sf;
                        step_sequence(s, fma3(toL, 0.01f, P), normalize(toL), true, seq[l], sh, sf);
                    }
            }
            size_t pos[32];
            // policy 0: if-if
            for (int l = 0; l < 32; l++) pos[l] = 0;
            for (;;) {
                int nn = 0, nl = 0;
                for (int l = 0; l < 32; l++) if (pos[l] < seq[l].size()) { if (seq[l][pos[l]]) nl++; else nn++; }
                if (!nn && !nl) break;
                if (nn) { out[0] += 1; out[2] += nn; }
                if (nl) { out[1] += 1; out[3] += nl; }
                for (int l = 0; l < 32; l++) if (pos[l] < seq[l].size()) pos[l]++;
            }
            // policy 1: while-while
            for (int l = 0; l < 32; l++) pos[l] = 0;
            for (;;) {
                bool any = false;
                for (;;) {   // node phase
                    int nn = 0;
                    for (int l = 0; l < 32; l++) if (pos[l] < seq[l].size() && seq[l][pos[l]] == 0) { nn++; pos[l]++; }
                    if (!nn) break;
                    out[4] += 1; out[6] += nn; any = true;
                }
                for (;;) {   // leaf phase
                    int nl = 0;
                    for (int l = 0; l < 32; l++) if (pos[l] < seq[l].size() && seq[l][pos[l]] == 1) { nl++; pos[l]++; }
                    if (!nl) break;
                    out[5] += 1; out[7] += nl; any = true;
                }
                if (!any) break;
            }
            // policy 2: if-if, leaf branch waits for a quorum
            for (int l = 0; l < 32; l++) pos[l] = 0;
            for (;;) {
                int nn = 0, nl = 0;
                for (int l = 0; l < 32; l++) if (pos[l] < seq[l].size()) { if (seq[l][pos[l]]) nl++; else nn++; }
                if (!nn && !nl) break;
                if (nn) {
                    out[8] += 1; out[10] += nn;
                    for (int l = 0; l < 32; l++) if (pos[l] < seq[l].size() && seq[l][pos[l]] == 0) pos[l]++;
                }
                if (nl && (nl >= quorum || !nn)) {
                    out[9] += 1; out[11] += nl;
                    for (int l = 0; l < 32; l++) if (pos[l] < seq[l].size() && seq[l][pos[l]] == 1) pos[l]++;
                }
            }
        }
    return 0;
}

// BVH introspection for structure checks: returns node count; copies nodes (16 floats each) and order.
extern "C" int emul_bvh(const oracle_scene* in, int leaf_size, float* nodes, uint32_t* order, uint64_t* keys, uint32_t* n_bvh) {
    EmulScene S;
    build(in, leaf_size, S);
    if (nodes) memcpy(nodes, S.nodes.data(), S.nodes.size() * sizeof(float4));
    if (order) memcpy(order, S.order.data(), S.n_bvh * sizeof(uint32_t));
    if (keys) memcpy(keys, S.keys.data(), S.n_bvh * sizeof(uint64_t));
    if (n_bvh) *n_bvh = S.n_bvh;
    return (int)(S.nodes.size() / 4);
}

// Per-pixel traversal cost of the primary rays (node visits, triangle tests) and, where shadow_steps is given,
// of the shadow rays of the primary hit (node visits + triangle tests, summed over the lights) — used to study
// load balance and the length of the per-lane dependency chains (profiles/r1_tuning.md).
extern "C" int emul_primary_cost(const oracle_scene* in, const rt_camera* cam, int leaf_size, uint32_t* nodes, uint32_t* tris,
                                 uint32_t* shadow_steps) {
    EmulScene S;
    build(in, leaf_size, S);
    const SceneDev& s = S.dev;
    const int W = cam->width, H = cam->height;
    bool overflow = false;
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            float xw = (float)((double)cam->aspect * (i - W / 2.0 + 0.5) / W), yw = (float)((j - H / 2.0 + 0.5) / H);
            double dd[3];
            for (int k = 0; k < 3; k++) dd[k] = -(double)cam->w[k] * (double)cam->focal_distance + (double)cam->u[k] * (double)xw + (double)cam->v[k] * (double)yw;
            double l = std::sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
            f3 d = mk3((float)(dd[0] / l), (float)(dd[1] / l), (float)(dd[2] / l));
            f3 o = mk3(cam->pos[0], cam->pos[1], cam->pos[2]);
            HitRec h;
            WorkCount wc{0, 0};
            bool found = trace_ray<false>(s, o, d, false, h, &wc, &overflow);
            size_t at = (size_t)i + (size_t)j * W;
            nodes[at] = wc.nodes;
            tris[at] = wc.tris;
            if (shadow_steps) {
                uint32_t steps = 0;
                if (found) {
                    f3 P = fma3(d, h.t, o);
                    for (int li = 0; li < s.n_lights; li++) {
                        f3 toL = mk3(s.lights[2 * li]) - P;                  // world.cpp:45
                        HitRec sh;
                        WorkCount ws{0, 0};
                        trace_ray<true>(s, fma3(toL, 0.01f, P), normalize(toL), false, sh, &ws, &overflow);
                        steps += ws.nodes + ws.tris;
                    }
                }
                shadow_steps[at] = steps;
            }
        }
    return 0;
}
