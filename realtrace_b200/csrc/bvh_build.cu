// bvh_build.cu — on-GPU LBVH construction and refit.
//
// Replaces the reference's acceleration-structure builders: UniformGrid::UniformGrid
// (/root/reference/Serial/uniform-grid.cpp:54-147) and the GPU grid build of the CUDA tracer
// (/root/reference/Parellel/kernel.cu:457-520: get_bounds, 7 x thrust::reduce, count_sizes,
// thrust::exclusive_scan, build_grid).  Pipeline, all on the context's stream:
//   k_scene_bounds   per-triangle AABB -> scene box + centroid box (block reduce, ordered-int atomics)
//   k_morton         63-bit Morton key of the AABB centre; oversized triangles get key ~0 and are
//                    later tested linearly (they would otherwise inflate every ancestor box)
//   radix sort       radix_sort.cu
//   k_tri_records    48-byte triangle records in Morton order
//   k_karras         Karras 2012 hierarchy, one thread per internal node
//   k_refit          bottom-up boxes with one atomic visit counter per node; the second arriver
//                    writes the 64-byte two-child-AABB node; subtrees of <= leaf_size triangles
//                    are referenced as one leaf (their triangles are contiguous in Morton order)
// A REFIT commit re-runs only k_tri_records + k_refit on the existing order ("SAH-free refit").
#include <cfloat>
#include <cstring>

#include "rt_context.h"

namespace {

constexpr int TPB = 256;
constexpr uint32_t RT_MAX_LARGE = 16;

__device__ __forceinline__ uint32_t ord_encode(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_decode(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ void load_tri(const float* __restrict__ v, uint32_t i, f3& a, f3& b, f3& c) {
    const float* p = v + 9 * (size_t)i;
    a = mk3(p[0], p[1], p[2]);
    b = mk3(p[3], p[4], p[5]);
    c = mk3(p[6], p[7], p[8]);
}

// bounds[0..2] = min of boxes, [3..5] = max of boxes, [6..8] = min of centres, [9..11] = max of centres
__global__ void __launch_bounds__(TPB) k_scene_bounds(const float* __restrict__ v, uint32_t n,
                                                     uint32_t* __restrict__ bounds) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
        f3 a, b, c;
        load_tri(v, i, a, b, c);
        Aabb bx = tri_aabb(a, b, c);
        float l[3] = {bx.lo.x, bx.lo.y, bx.lo.z}, h[3] = {bx.hi.x, bx.hi.y, bx.hi.z};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float ctr = 0.5f * (l[k] + h[k]);
            lo[k] = fminf(lo[k], l[k]); hi[k] = fmaxf(hi[k], h[k]);
            clo[k] = fminf(clo[k], ctr); chi[k] = fmaxf(chi[k], ctr);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
            clo[k] = fminf(clo[k], __shfl_xor_sync(0xffffffffu, clo[k], o));
            chi[k] = fmaxf(chi[k], __shfl_xor_sync(0xffffffffu, chi[k], o));
        }
    }
    // one set of 12 atomics per CTA (one per warp serialised 113 k same-address atomics at 1 M triangles: 78 us)
    __shared__ float red[TPB / 32][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            red[warp][k] = lo[k]; red[warp][3 + k] = hi[k]; red[warp][6 + k] = clo[k]; red[warp][9 + k] = chi[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const int k = threadIdx.x;
        const bool is_min = (k % 6) < 3;
        float r = red[0][k];
#pragma unroll
        for (int w = 1; w < TPB / 32; w++) r = is_min ? fminf(r, red[w][k]) : fmaxf(r, red[w][k]);
        if (is_min) atomicMin(&bounds[k], ord_encode(r)); else atomicMax(&bounds[k], ord_encode(r));
    }
}

__global__ void k_bounds_init(uint32_t* __restrict__ bounds) {
    int k = threadIdx.x;
    if (k < 12) bounds[k] = (k % 6) < 3 ? 0xffffffffu : 0u;
}

// Absolute pad of the node boxes: 4e-6 of the largest scene coordinate, derived on the device from the scene
// box of the CURRENT vertices (a refit of a mesh that moved or grew must not keep the pad of the old extent).
__device__ __forceinline__ float pad_from_bounds(const uint32_t* __restrict__ bounds) {
    float amax = 0.0f;
#pragma unroll
    for (int k = 0; k < 6; k++) amax = fmaxf(amax, fabsf(ord_decode(bounds[k])));
    return amax * 4.0e-6f;
}

__global__ void __launch_bounds__(TPB) k_morton(const float* __restrict__ v, uint32_t n,
                                               const uint32_t* __restrict__ bounds, float large_frac,
                                               uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                               uint32_t* __restrict__ n_large) {
    uint32_t i = blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    f3 slo = mk3(ord_decode(bounds[0]), ord_decode(bounds[1]), ord_decode(bounds[2]));
    f3 shi = mk3(ord_decode(bounds[3]), ord_decode(bounds[4]), ord_decode(bounds[5]));
    f3 clo = mk3(ord_decode(bounds[6]), ord_decode(bounds[7]), ord_decode(bounds[8]));
    f3 chi = mk3(ord_decode(bounds[9]), ord_decode(bounds[10]), ord_decode(bounds[11]));
    f3 a, b, c;
    load_tri(v, i, a, b, c);
    Aabb bx = tri_aabb(a, b, c);
    float scene_ext = fmaxf(shi.x - slo.x, fmaxf(shi.y - slo.y, shi.z - slo.z));
    float tri_ext = fmaxf(bx.hi.x - bx.lo.x, fmaxf(bx.hi.y - bx.lo.y, bx.hi.z - bx.lo.z));
    uint64_t key;
    if (tri_ext > large_frac * scene_ext) {
        key = ~0ull;
        atomicAdd(n_large, 1u);
    } else {
        f3 ext = chi - clo;
        // one scale for all axes (cubic Morton cells): a flat scene must not spend its top splits on
        // slicing the thin axis into pancakes
        float emax = fmaxf(ext.x, fmaxf(ext.y, ext.z));
        float iu = emax > 0.0f ? 1.0f / emax : 0.0f;
        f3 inv = mk3(iu, iu, iu);
        f3 ctr = (bx.lo + bx.hi) * 0.5f;
        key = morton63(ctr, clo, inv);
    }
    keys[i] = key;
    vals[i] = i;
}

// Triangle records in Morton order: (a, object id) (a-b, material id) (a-c, original index).
__global__ void __launch_bounds__(TPB) k_tri_records(const float* __restrict__ v, const uint32_t* __restrict__ mat,
                                                    const uint32_t* __restrict__ obj,
                                                    const uint32_t* __restrict__ order, uint32_t n,
                                                    float4* __restrict__ tris) {
    uint32_t k = blockIdx.x * TPB + threadIdx.x;
    if (k >= n) return;
    uint32_t i = order[k];
    f3 a, b, c;
    load_tri(v, i, a, b, c);
    f3 e1 = a - b, e2 = a - c;
    tris[3 * (size_t)k + 0] = make_float4(a.x, a.y, a.z, __uint_as_float(obj[i]));
    tris[3 * (size_t)k + 1] = make_float4(e1.x, e1.y, e1.z, __uint_as_float(mat[i]));
    tris[3 * (size_t)k + 2] = make_float4(e2.x, e2.y, e2.z, __uint_as_float(i));
}

// Oversized triangles (sorted positions n_bvh .. n_bvh + n_large - 1, ascending original index)
// become analytic primitives.
__global__ void k_emit_large(const float4* __restrict__ tris, uint32_t n_bvh, uint32_t n_large,
                             AnalyticPrim* __restrict__ out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_large) return;
    const float4* r = tris + 3 * (size_t)(n_bvh + k);
    AnalyticPrim p;
    p.kind = RT_KIND_TRIANGLE;
    p.object_id = __float_as_uint(r[0].w);
    p.material = __float_as_uint(r[1].w);
    p.tri_index = __float_as_uint(r[2].w);
    p.a = r[0]; p.b = r[1]; p.c = r[2];
    p.d = make_float4(0, 0, 0, 0);
    out[k] = p;
}

__global__ void __launch_bounds__(TPB) k_karras(const uint64_t* __restrict__ keys, int n, KarrasNode* __restrict__ kn,
                                               int* __restrict__ leaf_parent, int* __restrict__ node_parent) {
    int i = blockIdx.x * TPB + threadIdx.x;
    if (i >= n - 1) return;
    KarrasNode k = karras_node(keys, n, i);
    kn[i] = k;
    if (k.left < 0) leaf_parent[~k.left] = i; else node_parent[k.left] = i;
    if (k.right < 0) leaf_parent[~k.right] = i; else node_parent[k.right] = i;
    if (i == 0) node_parent[0] = -1;
}

__device__ __forceinline__ Aabb pad_box(Aabb b, float pad_abs) {
    Aabb r = aabb_pad(b);
    r.lo = r.lo - mk3(pad_abs, pad_abs, pad_abs);
    r.hi = r.hi + mk3(pad_abs, pad_abs, pad_abs);
    return r;
}

// Child reference + box as the parent stores them.
__device__ __forceinline__ int child_code(int child, const KarrasNode* __restrict__ kn, int leaf_size) {
    if (child < 0) return rt_leaf_code((uint32_t)~child, 1u);
    int cnt = kn[child].last - kn[child].first + 1;
    if (cnt <= leaf_size) return rt_leaf_code((uint32_t)kn[child].first, (uint32_t)cnt);
    return child;
}

// One thread per leaf; boxes of node i live at box[i], of leaf k at box[(n-1) + k].
__global__ void __launch_bounds__(TPB) k_refit(const float* __restrict__ v, const uint32_t* __restrict__ order, int n,
                                              const KarrasNode* __restrict__ kn, const int* __restrict__ leaf_parent,
                                              const int* __restrict__ node_parent, uint32_t* __restrict__ visit,
                                              float4* __restrict__ box_lo, float4* __restrict__ box_hi,
                                              float4* __restrict__ nodes, int leaf_size, const uint32_t* __restrict__ bounds) {
    int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= n) return;
    const float pad_abs = pad_from_bounds(bounds);
    f3 a, b, c;
    load_tri(v, order[k], a, b, c);
    Aabb bx = tri_aabb(a, b, c);
    box_lo[(n - 1) + k] = make_float4(bx.lo.x, bx.lo.y, bx.lo.z, 0.0f);
    box_hi[(n - 1) + k] = make_float4(bx.hi.x, bx.hi.y, bx.hi.z, 0.0f);
    // A node whose leaf range lies inside this CTA's 256 leaves is finished by two threads of this CTA: a
    // CTA-scope fence orders their box stores.  Only nodes that span CTAs (well under 1 % of them) pay for the
    // GPU-scope fence (MEMBAR.SC.GPU + L1 invalidate), which used to be issued on every level by every thread.
    const int cta_lo = blockIdx.x * TPB, cta_hi = min(cta_lo + TPB, n) - 1;
    int cur = leaf_parent[k];
    while (cur >= 0) {
        KarrasNode nd = kn[cur];
        if (nd.first >= cta_lo && nd.last <= cta_hi) __threadfence_block(); else __threadfence();
        if (atomicAdd(&visit[cur], 1u) == 0u) return;   // first arriver leaves; the sibling finishes the node
        int li = nd.left < 0 ? (n - 1) + ~nd.left : nd.left;
        int ri = nd.right < 0 ? (n - 1) + ~nd.right : nd.right;
        float4 l0 = __ldcg(box_lo + li), l1 = __ldcg(box_hi + li);
        float4 r0 = __ldcg(box_lo + ri), r1 = __ldcg(box_hi + ri);
        Aabb L, R;
        L.lo = mk3(l0); L.hi = mk3(l1); R.lo = mk3(r0); R.hi = mk3(r1);
        Aabb U = aabb_union(L, R);
        box_lo[cur] = make_float4(U.lo.x, U.lo.y, U.lo.z, 0.0f);
        box_hi[cur] = make_float4(U.hi.x, U.hi.y, U.hi.z, 0.0f);
        Aabb Lp = pad_box(L, pad_abs), Rp = pad_box(R, pad_abs);
        float4* out = nodes + RT_NODE_FLOAT4S * (size_t)cur;
        out[0] = make_float4(Lp.lo.x, Lp.hi.x, Lp.lo.y, Lp.hi.y);
        out[1] = make_float4(Rp.lo.x, Rp.hi.x, Rp.lo.y, Rp.hi.y);
        out[2] = make_float4(Lp.lo.z, Lp.hi.z, Rp.lo.z, Rp.hi.z);
        out[3] = make_float4(__int_as_float(child_code(nd.left, kn, leaf_size)),
                             __int_as_float(child_code(nd.right, kn, leaf_size)), __int_as_float(nd.first),
                             __int_as_float(nd.last));
        cur = node_parent[cur];
    }
}

// 4-wide view of the finished binary tree (rt_bvh.h: build_wide_node), one thread per binary node.
__global__ void __launch_bounds__(TPB) k_collapse4(const float4* __restrict__ nodes, int n_nodes, float4* __restrict__ nodes4) {
    int i = blockIdx.x * TPB + threadIdx.x;
    if (i >= n_nodes) return;
    float4 q[RT_NODE4_FLOAT4S];
    build_wide_node(nodes, i, q);
#pragma unroll
    for (int k = 0; k < RT_NODE4_FLOAT4S; k++) nodes4[RT_NODE4_FLOAT4S * (size_t)i + k] = q[k];
}

// A BVH of one triangle: root with one leaf child and one empty child.
__global__ void k_single_leaf(const float* __restrict__ v, const uint32_t* __restrict__ order,
                              float4* __restrict__ nodes, const uint32_t* __restrict__ bounds) {
    const float pad_abs = pad_from_bounds(bounds);
    f3 a, b, c;
    load_tri(v, order[0], a, b, c);
    Aabb p = pad_box(tri_aabb(a, b, c), pad_abs);
    nodes[0] = make_float4(p.lo.x, p.hi.x, p.lo.y, p.hi.y);
    // empty child: a point box at +FLT_MAX — every slab distance is +-inf, so the interval is empty
    nodes[1] = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
    nodes[2] = make_float4(p.lo.z, p.hi.z, FLT_MAX, FLT_MAX);
    nodes[3] = make_float4(__int_as_float(rt_leaf_code(0u, 1u)), __int_as_float(RT_EMPTY_CODE), __int_as_float(0),
                           __int_as_float(0));
}

inline int blocks_for(uint32_t n) { return (int)((n + TPB - 1) / TPB); }

}  // namespace

static void upload_static_scene(rt_ctx* c) {
    cudaStream_t st = c->stream;
    // materials: 3 float4 each
    std::vector<float4> m(3 * c->h_materials.size());
    c->has_reflective = c->has_dielectric = false;
    for (size_t i = 0; i < c->h_materials.size(); i++) {
        const rt_material& s = c->h_materials[i];
        m[3 * i + 0] = make_float4(s.color[0], s.color[1], s.color[2], s.ka);
        m[3 * i + 1] = make_float4(s.kd, s.ks, s.kr, s.kt);
        float fl;
        uint32_t f = s.flags;
        memcpy(&fl, &f, 4);
        m[3 * i + 2] = make_float4(s.eta, fl, 0.0f, 0.0f);
        if (s.kr > 0.0f) c->has_reflective = true;
        if (s.kr > 0.0f && s.kt > 0.0f) c->has_dielectric = true;
    }
    c->d_materials.reserve(m.size() ? m.size() : 1);
    if (!m.empty()) RT_CUDA(cudaMemcpyAsync(c->d_materials.p, m.data(), m.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    // lights: 2 float4 each
    size_t nl = c->h_lights.size() / 6;
    std::vector<float4> l(2 * nl);
    for (size_t i = 0; i < nl; i++) {
        const float* p = c->h_lights.data() + 6 * i;
        l[2 * i] = make_float4(p[0], p[1], p[2], 0.0f);
        l[2 * i + 1] = make_float4(p[3], p[4], p[5], 0.0f);
    }
    c->d_lights.reserve(l.size() ? l.size() : 1);
    if (!l.empty()) RT_CUDA(cudaMemcpyAsync(c->d_lights.p, l.data(), l.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    // analytic primitives: spheres, planes, cylinders (+ room for oversized triangles)
    std::vector<AnalyticPrim> a;
    a.insert(a.end(), c->h_spheres.begin(), c->h_spheres.end());
    a.insert(a.end(), c->h_planes.begin(), c->h_planes.end());
    a.insert(a.end(), c->h_cylinders.begin(), c->h_cylinders.end());
    c->n_fixed_analytic = (uint32_t)a.size();
    c->d_analytic.reserve(a.size() + RT_MAX_LARGE);
    if (!a.empty()) RT_CUDA(cudaMemcpyAsync(c->d_analytic.p, a.data(), a.size() * sizeof(AnalyticPrim), cudaMemcpyHostToDevice, st));
    // triangles
    c->n_tri = (uint32_t)(c->h_tri_v.size() / 9);
    uint32_t n = c->n_tri;
    c->d_tri_v.reserve(n ? 9 * (size_t)n : 1);
    c->d_tri_mat.reserve(n ? n : 1);
    c->d_tri_obj.reserve(n ? n : 1);
    if (n) {
        RT_CUDA(cudaMemcpyAsync(c->d_tri_v.p, c->h_tri_v.data(), 9 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
        RT_CUDA(cudaMemcpyAsync(c->d_tri_mat.p, c->h_tri_mat.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        RT_CUDA(cudaMemcpyAsync(c->d_tri_obj.p, c->h_tri_obj.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
    if (!c->h_tri_rgb.empty()) {
        std::vector<float4> rgb(3 * (size_t)n);
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) {
                const float* p = c->h_tri_rgb.data() + 9 * i + 3 * k;
                rgb[3 * i + k] = make_float4(p[0], p[1], p[2], 0.0f);
            }
        c->d_tri_rgb.reserve(rgb.size());
        RT_CUDA(cudaMemcpyAsync(c->d_tri_rgb.p, rgb.data(), rgb.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
        RT_CUDA(cudaStreamSynchronize(st));   // rgb is a local
    }
    RT_CUDA(cudaStreamSynchronize(st));       // m, l, a are locals
}

// The 4-wide view serves k_paths (latency-bound bounce paths; mode 1, only when something reflects), every fused
// walk (mode 2), or the heaviest tiles of a multi-GPU share (want_nodes4, set by the first frame that asks).
static bool wants_wide_view(const rt_ctx* c) {
    return (c->wide_bvh == 1 && c->has_reflective) || c->wide_bvh >= 2 || c->want_nodes4;
}

static void publish_scene(rt_ctx* c) {
    SceneDev& s = c->scene;
    s.nodes = c->d_nodes.p;
    s.nodes4 = (wants_wide_view(c) && c->n_bvh >= 1) ? c->d_nodes4.p : nullptr;
    s.tris = c->d_tris.p;
    s.tri_rgb = c->h_tri_rgb.empty() ? nullptr : c->d_tri_rgb.p;
    s.analytic = c->d_analytic.p;
    s.materials = c->d_materials.p;
    s.lights = c->d_lights.p;
    s.n_bvh_tris = (int)c->n_bvh;
    s.n_nodes = c->n_bvh >= 2 ? (int)c->n_bvh - 1 : (c->n_bvh == 1 ? 1 : 0);
    s.n_analytic = (int)(c->n_fixed_analytic + c->n_large);
    s.n_lights = (int)(c->h_lights.size() / 6);
    for (int k = 0; k < 3; k++) { s.ambient[k] = c->ambient[k]; s.background[k] = c->background[k]; }
}

static void launch_scene_bounds(rt_ctx* c) {
    cudaStream_t st = c->stream;
    uint32_t n = c->n_tri;
    k_bounds_init<<<1, 32, 0, st>>>(c->d_bounds.p);
    int bb = blocks_for(n);
    if (bb > c->sm_count * 8) bb = c->sm_count * 8;
    k_scene_bounds<<<bb, TPB, 0, st>>>(c->d_tri_v.p, n, c->d_bounds.p);
    RT_CUDA(cudaGetLastError());
}

static void launch_collapse4(rt_ctx* c) {
    uint32_t nb = c->n_bvh;
    if (nb < 1) return;
    int nn = nb >= 2 ? (int)nb - 1 : 1;
    c->d_nodes4.reserve(RT_NODE4_FLOAT4S * (size_t)nn);
    k_collapse4<<<blocks_for((uint32_t)nn), TPB, 0, c->stream>>>(c->d_nodes.p, nn, c->d_nodes4.p);
    RT_CUDA(cudaGetLastError());
    c->launch_total++;
}

static void run_refit(rt_ctx* c, bool new_vertices) {
    cudaStream_t st = c->stream;
    uint32_t n = c->n_tri, nb = c->n_bvh;
    const uint32_t* order = c->d_vals[c->sorted_buf].p;
    if (new_vertices && n) {       // REFIT commit: the pad of the node boxes follows the current extent
        launch_scene_bounds(c);
        c->launch_total += 2;
    }
    if (n) {
        k_tri_records<<<blocks_for(n), TPB, 0, st>>>(c->d_tri_v.p, c->d_tri_mat.p, c->d_tri_obj.p, order, n, c->d_tris.p);
        RT_CUDA(cudaGetLastError());
        c->launch_total++;
    }
    if (c->n_large) {
        k_emit_large<<<1, 32, 0, st>>>(c->d_tris.p, nb, c->n_large, c->d_analytic.p + c->n_fixed_analytic);
        RT_CUDA(cudaGetLastError());
        c->launch_total++;
    }
    if (nb >= 2) {
        RT_CUDA(cudaMemsetAsync(c->d_visit.p, 0, (nb - 1) * sizeof(uint32_t), st));
        k_refit<<<blocks_for(nb), TPB, 0, st>>>(c->d_tri_v.p, order, (int)nb, c->d_karras.p, c->d_leaf_parent.p,
                                                c->d_node_parent.p, c->d_visit.p, c->d_box_lo.p, c->d_box_hi.p,
                                                c->d_nodes.p, c->leaf_size, c->d_bounds.p);
        RT_CUDA(cudaGetLastError());
        c->launch_total++;
    } else if (nb == 1) {
        k_single_leaf<<<1, 1, 0, st>>>(c->d_tri_v.p, order, c->d_nodes.p, c->d_bounds.p);
        RT_CUDA(cudaGetLastError());
        c->launch_total++;
    }
    if (wants_wide_view(c)) launch_collapse4(c);
}

void rt_ensure_nodes4(rt_ctx* c) {
    c->want_nodes4 = true;
    if (c->scene.nodes4 || c->n_bvh < 1) return;
    launch_collapse4(c);
    publish_scene(c);
}

void rt_build_bvh(rt_ctx* c, bool refit_only) {
    cudaStream_t st = c->stream;
    if (refit_only) {
        // asynchronous: only enqueues; rt_scene_build_stats reads the event pair later
        RT_CUDA(cudaEventRecord(c->ev[8], st));
        run_refit(c, true);
        RT_CUDA(cudaEventRecord(c->ev[9], st));
        c->refit_pending = true;
        publish_scene(c);
        return;
    }
    upload_static_scene(c);
    uint32_t n = c->n_tri;
    c->n_bvh = 0;
    c->n_large = 0;
    c->build_stats = rt_build_stats{};
    RT_CUDA(cudaEventRecord(c->ev[4], st));
    if (n) {
        c->d_keys[0].reserve(n);
        c->d_vals[0].reserve(n);
        c->d_bounds.reserve(12);
        c->d_misc.reserve(4);
        c->d_tris.reserve(3 * (size_t)n);
        c->d_nodes.reserve(RT_NODE_FLOAT4S * (size_t)(n > 1 ? n - 1 : 1));
        c->d_box_lo.reserve(2 * (size_t)n);
        c->d_box_hi.reserve(2 * (size_t)n);
        c->d_karras.reserve(n);
        c->d_leaf_parent.reserve(n);
        c->d_node_parent.reserve(n);
        c->d_visit.reserve(n);

        launch_scene_bounds(c);

        float large_frac = n >= 64 ? 0.25f : FLT_MAX;
        uint32_t h_large = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            RT_CUDA(cudaMemsetAsync(c->d_misc.p, 0, 4 * sizeof(uint32_t), st));
            k_morton<<<blocks_for(n), TPB, 0, st>>>(c->d_tri_v.p, n, c->d_bounds.p, large_frac, c->d_keys[0].p,
                                                    c->d_vals[0].p, c->d_misc.p);
            RT_CUDA(cudaGetLastError());
            RT_CUDA(cudaMemcpyAsync(&h_large, c->d_misc.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            RT_CUDA(cudaStreamSynchronize(st));
            if (h_large <= RT_MAX_LARGE) break;
            large_frac = FLT_MAX;   // too many outliers to test linearly: keep them all in the hierarchy
        }
        c->n_large = h_large;
        c->n_bvh = n - h_large;

        rt_sort_pairs_device(c, n);
        if (c->n_bvh >= 2) {
            k_karras<<<blocks_for(c->n_bvh - 1), TPB, 0, st>>>(c->d_keys[c->sorted_buf].p, (int)c->n_bvh, c->d_karras.p,
                                                              c->d_leaf_parent.p, c->d_node_parent.p);
            RT_CUDA(cudaGetLastError());
        }
    }
    run_refit(c, false);
    RT_CUDA(cudaEventRecord(c->ev[5], st));
    RT_CUDA(cudaEventSynchronize(c->ev[5]));
    RT_CUDA(cudaEventElapsedTime(&c->build_stats.ms_build, c->ev[4], c->ev[5]));
    c->build_stats.sort_passes = n >= 2 ? (uint32_t)rt_sort_passes_done(c) : 0u;
    c->build_stats.n_triangles = c->n_bvh;
    c->build_stats.leaf_size = (uint32_t)c->leaf_size;
    c->build_stats.n_large_triangles = c->n_large;
    c->build_stats.n_nodes = c->n_bvh >= 2 ? c->n_bvh - 1 : (c->n_bvh == 1 ? 1u : 0u);
    publish_scene(c);
}
