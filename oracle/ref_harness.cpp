// ref_harness.cpp — headless driver around the reference's OWN Serial classes.
//
// TEST INFRASTRUCTURE ONLY (see oracle_abi.h).  This file is compiled by
// oracle/build_ref.py together with the reference sources taken from
// /root/reference/Serial (staged + line-patched in a temp dir, never copied
// into this repository) into oracle/_ref/libserial_ref.so.
//
// It restates what Serial/lumina.cpp does around the render core:
//   scene construction        lumina.cpp:245-272, 289, 302-310, 360-362
//   the per-pixel driver      renderengine.cpp:3-8, 10-18 (trace is private and
//                             renderLoop keeps a function-static column cursor,
//                             so the column-subset / threaded paths call
//                             World::shade_ray + Color::clamp + drawPixel
//                             directly; the full single-thread path calls
//                             RenderEngine::renderLoop() itself)
// Everything the harness computes comes out of the reference's classes.

#include "world.h"
#include "camera.h"
#include "renderengine.h"
#include "material.h"
#include "sphere.h"
#include "plane.h"
#include "cylinder.h"
#include "triangle.h"
#include "pointlightsource.h"
#include "uniform-grid.h"

#include <algorithm>
#include <chrono>
#include <thread>
#include <unordered_map>
#include <vector>
#include <cfloat>

#include "oracle_abi.h"

// Globals the staged patches refer to (see build_ref.py for each hunk).
int g_oracle_depth = 10;                 // replaces `#define RECURSION_DEPTH 10` (world.h:11)
int g_oracle_mode = ORACLE_MODE_AS_SHIPPED;
int g_oracle_bbox_fixed = 0;             // utilities.h:16 min() -> lowest() when 1
bool g_oracle_has_grid = false;          // default-constructed UniformGrid must not be walked
std::vector<int> g_oracle_analytic;      // objectList indices of non-triangle objects
thread_local unsigned long long g_oracle_rays = 0;

namespace {

struct BuiltScene {
    World* world = nullptr;
    std::vector<Material*> plain_materials;
    std::vector<BarycentricMaterial*> bary_materials;
    std::vector<Triangle*> triangles;
    std::vector<Sphere*> spheres;
    std::vector<Plane*> planes;
    std::vector<Cylinder*> cylinders;
    std::vector<PointLightSource*> lights;
    std::unordered_map<const Object*, int> object_index;
    double build_seconds = 0;

    ~BuiltScene() {
        for (auto p : triangles) delete p;
        for (auto p : spheres) delete p;
        for (auto p : planes) delete p;
        for (auto p : cylinders) delete p;
        for (auto p : lights) delete p;
        for (auto p : plain_materials) delete p;
        for (auto p : bary_materials) delete p;
        delete world;
    }
};

Vector3D vec(const float* p) { return Vector3D((double)p[0], (double)p[1], (double)p[2]); }
Color col(const float* p) { return Color((double)p[0], (double)p[1], (double)p[2]); }

void fill_material(Material* m, const oracle_material& s) {
    m->color = col(s.color);
    m->ka = s.ka; m->kd = s.kd; m->ks = s.ks; m->kr = s.kr; m->kt = s.kt; m->eta = s.eta;
}

struct Pending { uint32_t object_id; Object* obj; bool is_triangle; };

bool build_scene(const oracle_scene* s, int mode, BuiltScene& out) {
    auto t0 = std::chrono::steady_clock::now();
    g_oracle_mode = mode;
    g_oracle_bbox_fixed = (mode == ORACLE_MODE_BBOX_FIXED) ? 1 : 0;
    out.world = new World;
    World* w = out.world;
    w->setAmbient(col(s->ambient));
    w->setBackground(col(s->background));

    // Shared flat materials, one reference Material per record.
    std::vector<Material*> shared(s->n_materials, nullptr);
    for (uint32_t i = 0; i < s->n_materials; i++) {
        if (s->materials[i].barycentric) continue;  // instantiated per triangle below
        Material* m = new Material(w);
        fill_material(m, s->materials[i]);
        shared[i] = m;
        out.plain_materials.push_back(m);
    }

    std::vector<Pending> pending;
    pending.reserve(s->n_tri + s->n_sph + s->n_pln + s->n_cyl);
    for (uint32_t i = 0; i < s->n_tri; i++) {
        const float* v = s->tri_v + 9 * (size_t)i;
        uint32_t mi = s->tri_material[i];
        if (mi >= s->n_materials) return false;
        Material* m = shared[mi];
        if (s->materials[mi].barycentric) {
            if (!s->tri_rgb) return false;
            const float* c = s->tri_rgb + 9 * (size_t)i;
            // lumina.cpp:249 — one BarycentricMaterial per textured face
            BarycentricMaterial* bm = new BarycentricMaterial(w, vec(v), vec(v + 3), vec(v + 6),
                                                              col(c), col(c + 3), col(c + 6));
            fill_material(bm, s->materials[mi]);
            out.bary_materials.push_back(bm);
            m = bm;
        }
        Triangle* t = new Triangle(vec(v), vec(v + 3), vec(v + 6), m);  // lumina.cpp:263
        out.triangles.push_back(t);
        pending.push_back({s->tri_object_id ? s->tri_object_id[i] : i, t, true});
    }
    uint32_t next_id = s->n_tri;
    for (uint32_t i = 0; i < s->n_sph; i++) {
        const float* p = s->sph + 4 * (size_t)i;
        if (s->sph_material[i] >= s->n_materials || !shared[s->sph_material[i]]) return false;
        Sphere* o = new Sphere(vec(p), (double)p[3], shared[s->sph_material[i]]);
        out.spheres.push_back(o);
        pending.push_back({s->sph_object_id ? s->sph_object_id[i] : next_id, o, false});
        next_id++;
    }
    for (uint32_t i = 0; i < s->n_pln; i++) {
        const float* p = s->pln + 12 * (size_t)i;
        if (s->pln_material[i] >= s->n_materials || !shared[s->pln_material[i]]) return false;
        Plane* o = new Plane(vec(p), vec(p + 3), vec(p + 6), vec(p + 9), shared[s->pln_material[i]]);
        out.planes.push_back(o);
        pending.push_back({s->pln_object_id ? s->pln_object_id[i] : next_id, o, false});
        next_id++;
    }
    for (uint32_t i = 0; i < s->n_cyl; i++) {
        const float* p = s->cyl + 7 * (size_t)i;
        if (s->cyl_material[i] >= s->n_materials || !shared[s->cyl_material[i]]) return false;
        Cylinder* o = new Cylinder(vec(p), (double)p[3], vec(p + 4), shared[s->cyl_material[i]]);
        out.cylinders.push_back(o);
        pending.push_back({s->cyl_object_id ? s->cyl_object_id[i] : next_id, o, false});
        next_id++;
    }
    std::stable_sort(pending.begin(), pending.end(),
                     [](const Pending& a, const Pending& b) { return a.object_id < b.object_id; });

    g_oracle_analytic.clear();
    std::vector<Triangle*> grid_triangles;
    for (size_t k = 0; k < pending.size(); k++) {
        w->addObject(pending[k].obj);                       // lumina.cpp:269
        out.object_index[pending[k].obj] = (int)k;
        if (pending[k].is_triangle) grid_triangles.push_back(static_cast<Triangle*>(pending[k].obj));
        else g_oracle_analytic.push_back((int)k);
    }
    for (uint32_t i = 0; i < s->n_lights; i++) {
        const float* p = s->lights + 6 * (size_t)i;
        PointLightSource* l = new PointLightSource(w, vec(p), col(p + 3));  // lumina.cpp:360
        out.lights.push_back(l);
        w->addLight(l);
    }
    g_oracle_has_grid = !grid_triangles.empty() && mode != ORACLE_MODE_TRUE_NEAREST;
    if (g_oracle_has_grid) w->uniform_grid = UniformGrid(grid_triangles);   // lumina.cpp:289
    out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

int index_of(const BuiltScene& b, const Object* o) {
    if (!o) return -1;
    auto it = b.object_index.find(o);
    return it == b.object_index.end() ? -1 : it->second;
}

}  // namespace

extern "C" const char* oracle_name(void) { return "reference:Serial"; }

extern "C" int oracle_render(const oracle_scene* scene, const oracle_camera* c, int max_depth, int mode,
                             int col_begin, int col_step, int nthreads, uint8_t* rgb, int32_t* prim_id,
                             float* t_hit, oracle_result* result) {
    if (!scene || !c || !rgb || col_step < 1 || col_begin < 0 || c->width < 1 || c->height < 1) return -1;
    BuiltScene b;
    if (!build_scene(scene, mode, b)) return -2;
    g_oracle_depth = max_depth;

    Camera cam(Vector3D(c->pos[0], c->pos[1], c->pos[2]), Vector3D(c->target[0], c->target[1], c->target[2]),
               Vector3D(c->up[0], c->up[1], c->up[2]), c->fovy, c->width, c->height);
    const int W = c->width, H = c->height;
    unsigned long long rays = 0;
    unsigned columns = 0;
    if (nthreads < 1) nthreads = 1;

    auto t0 = std::chrono::steady_clock::now();
    if (col_begin == 0 && col_step == 1 && nthreads == 1) {
        // The reference's own frame loop (lumina.cpp:464): one column per call.
        g_oracle_rays = 0;
        RenderEngine engine(b.world, &cam);
        while (!engine.renderLoop()) {}
        rays = g_oracle_rays;
        columns = (unsigned)W;
    } else {
        std::vector<int> cols;
        for (int i = col_begin; i < W; i += col_step) cols.push_back(i);
        columns = (unsigned)cols.size();
        std::vector<unsigned long long> per_thread(nthreads, 0);
        auto worker = [&](int tid) {
            g_oracle_rays = 0;
            for (size_t k = tid; k < cols.size(); k += nthreads) {
                int i = cols[k];
                for (int j = 0; j < H; j++) {
                    // renderengine.cpp:3-8 then :15-17
                    Vector3D dir = cam.get_ray_direction(i, j);
                    Ray ray(cam.get_position(), dir);
                    Color color = b.world->shade_ray(ray);
                    color.clamp();
                    cam.drawPixel(i, j, color);
                }
            }
            per_thread[tid] = g_oracle_rays;
        };
        if (nthreads == 1) worker(0);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads; t++) pool.emplace_back(worker, t);
            for (auto& t : pool) t.join();
        }
        for (auto v : per_thread) rays += v;
    }
    double render_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    const unsigned char* bm = cam.getBitmap();
    for (int i = col_begin; i < W; i += col_step)
        for (int j = 0; j < H; j++) {
            size_t o = ((size_t)i + (size_t)j * W) * 3;
            rgb[o] = bm[o]; rgb[o + 1] = bm[o + 1]; rgb[o + 2] = bm[o + 2];
        }

    if (prim_id || t_hit) {
        for (int i = col_begin; i < W; i += col_step)
            for (int j = 0; j < H; j++) {
                Ray ray(cam.get_position(), cam.get_ray_direction(i, j));
                b.world->firstIntersection(ray);
                size_t o = (size_t)i + (size_t)j * W;
                if (prim_id) prim_id[o] = ray.didHit() ? index_of(b, ray.intersected()) : -1;
                if (t_hit) t_hit[o] = ray.didHit() ? ray.getParameter() : FLT_MAX;
            }
    }
    if (result) {
        result->rays_total = rays;
        result->rays_primary = result->rays_shadow = result->rays_secondary = 0;
        result->build_seconds = b.build_seconds;
        result->render_seconds = render_s;
        result->columns_rendered = columns;
        result->threads_used = (uint32_t)nthreads;
    }
    return 0;
}

extern "C" int oracle_trace_rays(const oracle_scene* scene, int mode, const float* rays, uint32_t n_rays,
                                 int32_t* prim_id, float* t_hit) {
    if (!scene || !rays) return -1;
    BuiltScene b;
    if (!build_scene(scene, mode, b)) return -2;
    for (uint32_t r = 0; r < n_rays; r++) {
        const float* p = rays + 6 * (size_t)r;
        Ray ray(vec(p), vec(p + 3));
        b.world->firstIntersection(ray);
        if (prim_id) prim_id[r] = ray.didHit() ? index_of(b, ray.intersected()) : -1;
        if (t_hit) t_hit[r] = ray.didHit() ? ray.getParameter() : FLT_MAX;
    }
    return 0;
}

extern "C" int oracle_shade_rays(const oracle_scene* scene, int max_depth, int mode, const float* rays,
                                 uint32_t n_rays, double* rgb_out) {
    if (!scene || !rays || !rgb_out) return -1;
    BuiltScene b;
    if (!build_scene(scene, mode, b)) return -2;
    g_oracle_depth = max_depth;
    for (uint32_t r = 0; r < n_rays; r++) {
        const float* p = rays + 6 * (size_t)r;
        Ray ray(vec(p), vec(p + 3));
        Color c = b.world->shade_ray(ray);
        rgb_out[3 * (size_t)r] = c.r; rgb_out[3 * (size_t)r + 1] = c.g; rgb_out[3 * (size_t)r + 2] = c.b;
    }
    return 0;
}
