// rt_shade.h — one surface interaction of the Whitted integrator, restated for a wavefront.
//
// The reference recursion (World::shade_ray, /root/reference/Serial/world.cpp:32-111) is linear in
// the colours its children return:  colour(ray) = local + sum_i w_i * colour(child_i).  A ray
// therefore carries an RGB throughput; shading a hit adds throughput*local to the pixel and emits
// children with throughput*w_i.  Every quirk of SURVEY Appendix A is kept (Q5-Q13); where this
// file deviates it says so:
//   * dielectric hits discard their local colour (world.cpp:77-100), so the shadow rays the
//     reference still traces for them (world.cpp:44-51) are neither traced nor counted here;
//   * BarycentricMaterial::shade is evaluated once per hit from the hit's own barycentrics; the
//     reference re-derives them from Ray(P, view) inside get_light_shade (world.cpp:136), which is
//     the same point up to rounding (differs only at edge epsilon-ties);
//   * children that the reference would start and immediately terminate (level > max depth,
//     world.cpp:34; NaN direction after an ignored refract() failure, world.cpp:83) are resolved
//     in place as throughput*background and are not counted as rays.
#pragma once

#include "rt_traverse.h"

struct ShadeChild {
    f3 o, d, w;    // origin, unit direction, throughput factor relative to the parent ray
    int level;
};

struct ShadeOut {
    f3 local;          // colour added for this hit (to be scaled by the ray's throughput)
    f3 bg_weight;      // sum of child weights that resolve to the background immediately
    int n_children;
    ShadeChild child[2];
    int shadow_rays;   // any-hit queries issued
};

RT_HD float pow128(float x) {   // pow(x, 128), world.cpp:134: integer exponent, sign lost
    x *= x; x *= x; x *= x; x *= x; x *= x; x *= x; x *= x;
    return x;
}
RT_HD float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }   // world.cpp:96

RT_HD f3 reflect3(f3 I, f3 N) { return I - N * (2.0f * dot(N, I)); }             // world.cpp:27-30
RT_HD bool refract3(f3 I, f3 N, float eta, f3& T) {                               // world.cpp:19-25
    float ni = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - ni * ni);
    if (k < 0.0f) return false;
    T = I * eta - N * (eta * ni + sqrtf(k));
    return true;
}

RT_HD void push_child(ShadeOut& out, const SceneDev& s, f3 P, f3 dir, f3 w, int level, int max_depth) {
    f3 dn = normalize(dir);                                        // Ray ctor, ray.h:25-29
    bool dead = level > max_depth || !(dn.x == dn.x && dn.y == dn.y && dn.z == dn.z);
    if (dead) { out.bg_weight = out.bg_weight + w; return; }
    ShadeChild& c = out.child[out.n_children++];
    c.o = fma3(dir, 1e-4f, P);                                     // P + 1e-4 * R  (world.cpp:91,97,98,104)
    c.d = dn;
    c.w = w;
    c.level = level;
}

// Surface data of a hit: raw normal, material, albedo (Material::shade / BarycentricMaterial::shade).
struct Surface {
    f3 N;
    MaterialRec m;
    f3 albedo;
};

RT_HD Surface load_surface(const SceneDev& s, const HitRec& h, f3 P) {
    Surface sf;
    uint32_t mat, orig;
    if (h.prim >= 0) {
        const float4* rec = s.tris + 3 * (size_t)h.prim;
        float4 r1 = ldg(rec + 1), r2 = ldg(rec + 2);
        sf.N = cross(mk3(r1), mk3(r2));                            // (a-b) x (a-c), triangle.cpp:28
        mat = as_uint(r1.w);
        orig = as_uint(r2.w);
    } else {
        const AnalyticPrim p = s.analytic[rt_analytic_index(h.prim)];
        sf.N = analytic_normal(p, P);
        mat = p.material;
        orig = p.tri_index;
    }
    sf.m = load_material(s, mat);
    if ((sf.m.flags & 1u) && s.tri_rgb) {                          // material.cpp:19-21
        const float4* c = s.tri_rgb + 3 * (size_t)orig;
        float alpha = 1.0f - (h.beta + h.gamma);
        sf.albedo = mk3(ldg(c)) * alpha + mk3(ldg(c + 1)) * h.beta + mk3(ldg(c + 2)) * h.gamma;
    } else {
        sf.albedo = sf.m.color;                                    // material.cpp:5-8
    }
    return sf;
}

// AnyHit: callable bool(f3 origin, f3 unit_direction) — the shadow query.
template <class AnyHit>
RT_HD void shade_hit(const SceneDev& s, f3 o, f3 d, int level, const HitRec& h, int max_depth, AnyHit any_hit,
                     ShadeOut& out) {
    out.local = mk3(0.0f, 0.0f, 0.0f);
    out.bg_weight = mk3(0.0f, 0.0f, 0.0f);
    out.n_children = 0;
    out.shadow_rays = 0;

    f3 P = fma3(d, h.t, o);                                        // Ray::getPosition, ray.h:32
    Surface sf = load_surface(s, h, P);
    f3 N = normalize(sf.N);                                        // world.cpp:66-68, :129
    f3 I = normalize(d);                                           // world.cpp:67,69
    const MaterialRec& m = sf.m;
    f3 ambient = mk3(s.ambient[0], s.ambient[1], s.ambient[2]);

    if (m.kr > 0.0f && m.kt > 0.0f) {                              // dielectric, world.cpp:77-100
        f3 R = reflect3(I, N);
        f3 T = mk3(0.0f, 0.0f, 0.0f);
        f3 k = mk3(1.0f, 1.0f, 1.0f);
        float c;
        if (dot(d, N) < 0.0f) {                                    // entering, :81-85
            refract3(I, N, m.eta, T);                              // failure ignored: T stays 0 -> NaN ray
            c = -dot(I, N);
        } else {                                                   // leaving, :86-94
            const float e = 2.718282f;                             // world.cpp:2
            k = mk3(powf(e, -0.27f * h.t), powf(e, -0.45f * h.t), powf(e, -0.55f * h.t));
            if (refract3(I, -N, 1.0f / m.eta, T)) c = dot(T, N);
            else {
                push_child(out, s, P, R, k, level + 1, max_depth);
                return;
            }
        }
        float R0 = ((m.eta - 1.0f) * (m.eta - 1.0f)) / ((m.eta + 1.0f) * (m.eta + 1.0f));
        float Rs = R0 + (1.0f - R0) * pow5(1.0f - c);
        push_child(out, s, P, R, k * Rs, level + 1, max_depth);
        push_child(out, s, P, T, k * (1.0f - Rs), level * 2, max_depth);   // :98 level*2
        return;
    }

    // ---- local illumination, world.cpp:40-63 and get_light_shade :126-137
    f3 amb = ambient * sf.albedo * m.ka;
    f3 light_color = mk3(0.0f, 0.0f, 0.0f);
    bool is_shadow = false;
    for (int li = 0; li < s.n_lights; li++) {
        f3 Lp = mk3(ldg(s.lights + 2 * li));
        f3 Li = mk3(ldg(s.lights + 2 * li + 1));
        f3 toL = Lp - P;
        f3 ldir = normalize(toL);
        out.shadow_rays++;
        if (any_hit(fma3(toL, 0.01f, P), ldir)) is_shadow = true;  // :45-47 (no tmax, Q6)
        f3 r = normalize(reflect3(-ldir, N));                      // :130
        float diffuse = fmaxf(dot(N, normalize(Lp)), 0.0f);        // :133 light POSITION vector (Q8)
        float specular = pow128(dot(I, r));                        // :134 (Q9)
        light_color = light_color + Li * (m.kd * diffuse) * sf.albedo + Li * (m.ks * specular);   // :136
    }
    light_color = light_color + amb;                               // :59
    f3 final_color = light_color;
    if (is_shadow) final_color = final_color * 1e-4f + amb * (1.0f - 1e-4f);   // :63 (Q7)
    out.local = final_color;

    if (m.kr > 0.0f) {                                             // mirror, world.cpp:101-107
        f3 R = reflect3(I, N);
        push_child(out, s, P, R, mk3(m.kr, m.kr, m.kr), level + 1, max_depth);
    }
}
