// rt_bvh.h — per-element pieces of the LBVH builder (Morton keys, Karras' hierarchy, node
// encoding), RT_HD so tests/emul can run them on the CPU.  The LBVH replaces the reference's
// uniform grid (/root/reference/Serial/uniform-grid.cpp:54-147) and its k-d tree stub
// (kdtree.cpp, never built).
#pragma once

#include "rt_scene.h"

// ---- Morton keys ---------------------------------------------------------------------------------
RT_HD uint64_t spread21(uint32_t v) {   // 21 bits -> every third bit of 63
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// centroid c inside [lo, lo + 1/inv_ext] per axis -> 63-bit key (x most significant of each triple)
RT_HD uint64_t morton63(f3 c, f3 lo, f3 inv_ext) {
    const float scale = 2097152.0f;   // 2^21
    float fx = (c.x - lo.x) * inv_ext.x * scale;
    float fy = (c.y - lo.y) * inv_ext.y * scale;
    float fz = (c.z - lo.z) * inv_ext.z * scale;
    uint32_t ix = (uint32_t)fminf(fmaxf(fx, 0.0f), scale - 1.0f);
    uint32_t iy = (uint32_t)fminf(fmaxf(fy, 0.0f), scale - 1.0f);
    uint32_t iz = (uint32_t)fminf(fmaxf(fz, 0.0f), scale - 1.0f);
    return (spread21(ix) << 2) | (spread21(iy) << 1) | spread21(iz);
}

// ---- Karras 2012: one internal node per thread ---------------------------------------------------
// delta(i, j): length of the common prefix of keys i and j; ties are broken by the index so that
// duplicate keys still produce a balanced subtree.  -1 outside [0, n).
RT_HD int karras_delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = ldg(keys + i), b = ldg(keys + j);
    if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz64(a ^ b);
}

struct KarrasNode {
    int left, right;       // >= 0: internal node; < 0: ~leaf index (sorted position)
    int first, last;       // sorted range covered
};

RT_HD KarrasNode karras_node(const uint64_t* keys, int n, int i) {
    int dl = karras_delta(keys, n, i, i - 1), dr = karras_delta(keys, n, i, i + 1);
    int d = dr > dl ? 1 : -1;
    int dmin = d > 0 ? dl : dr;
    int lmax = 2;
    while (karras_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (karras_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = karras_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (karras_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + (d < 0 ? -1 : 0);
    KarrasNode k;
    k.first = i < j ? i : j;
    k.last = i < j ? j : i;
    k.left = (k.first == gamma) ? ~gamma : gamma;
    k.right = (k.last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    return k;
}

// ---- boxes ---------------------------------------------------------------------------------------
struct Aabb {
    f3 lo, hi;
};
RT_HD Aabb aabb_union(Aabb a, Aabb b) {
    Aabb r;
    r.lo = mk3(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z));
    r.hi = mk3(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z));
    return r;
}
RT_HD Aabb tri_aabb(f3 a, f3 b, f3 c) {
    Aabb r;
    r.lo = mk3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
    r.hi = mk3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
    return r;
}
// Boxes stored in nodes are widened by a few ulps so that the FP32 slab test can never reject a
// box whose triangle the (more precise) triangle test would accept.
RT_HD Aabb aabb_pad(Aabb b) {
    const float rel = 4.0e-7f, absv = 1.0e-30f;
    Aabb r;
    r.lo = mk3(b.lo.x - (fabsf(b.lo.x) * rel + absv), b.lo.y - (fabsf(b.lo.y) * rel + absv),
               b.lo.z - (fabsf(b.lo.z) * rel + absv));
    r.hi = mk3(b.hi.x + (fabsf(b.hi.x) * rel + absv), b.hi.y + (fabsf(b.hi.y) * rel + absv),
               b.hi.z + (fabsf(b.hi.z) * rel + absv));
    return r;
}

// ---- 4-wide collapse ---------------------------------------------------------------------------------
// Wide node of binary node i: each child of i that is itself an internal node is replaced by its two
// children.  Everything needed sits in the finished binary records: nodes[i] holds the (padded) boxes and
// codes of i's children, nodes[c] those of c's children.
RT_HD void wide_slot(float4* q, int k, float lox, float hix, float loy, float hiy, float loz, float hiz, int code) {
    float* p;
    p = (float*)&q[0]; p[k] = lox;  p = (float*)&q[1]; p[k] = hix;
    p = (float*)&q[2]; p[k] = loy;  p = (float*)&q[3]; p[k] = hiy;
    p = (float*)&q[4]; p[k] = loz;  p = (float*)&q[5]; p[k] = hiz;
    p = (float*)&q[6]; p[k] = as_float((uint32_t)code);
}
RT_HD void build_wide_node(const float4* nodes, int i, float4* q /*8*/) {
    const float big = RT_FLT_MAX;
    for (int k = 0; k < 4; k++) wide_slot(q, k, big, big, big, big, big, big, RT_EMPTY_CODE);
    q[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const float4* n = nodes + RT_NODE_FLOAT4S * (size_t)i;
    float4 n0 = n[0], n1 = n[1], n2 = n[2], n3 = n[3];
    int code[2] = {(int)as_uint(n3.x), (int)as_uint(n3.y)};
    int k = 0;
    for (int c = 0; c < 2; c++) {
        if (code[c] == RT_EMPTY_CODE) continue;
        if (code[c] < 0) {   // leaf child of i: keep it with the box i stores for it
            if (c == 0) wide_slot(q, k++, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, code[c]);
            else wide_slot(q, k++, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, code[c]);
        } else {             // internal child: its two children move up
            const float4* m = nodes + RT_NODE_FLOAT4S * (size_t)code[c];
            float4 m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3];
            int g0 = (int)as_uint(m3.x), g1 = (int)as_uint(m3.y);
            if (g0 != RT_EMPTY_CODE) wide_slot(q, k++, m0.x, m0.y, m0.z, m0.w, m2.x, m2.y, g0);
            if (g1 != RT_EMPTY_CODE) wide_slot(q, k++, m1.x, m1.y, m1.z, m1.w, m2.z, m2.w, g1);
        }
    }
}
