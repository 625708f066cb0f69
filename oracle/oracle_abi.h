/*
 * oracle_abi.h — flat C interface shared by the two CPU checkers.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load these libraries, and only as the checker.
 *
 * Two shared objects export exactly this interface:
 *   oracle/_ref/libserial_ref.so   the reference's own Serial/ sources
 *                                  (compiled by oracle/build_ref.py from
 *                                  /root/reference, git-ignored)
 *   oracle/libserial_port.so       oracle/serial_port.cpp, a from-scratch FP64
 *                                  restatement of the same algorithm that is
 *                                  pinned bit-for-bit against the former.
 *
 * Scene inputs are float32 (the same arrays the GPU library receives through
 * include/realtrace_b200.h); both checkers widen them to double, so the CPU
 * and GPU sides see identical geometry.
 */
#ifndef ORACLE_ABI_H
#define ORACLE_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Material record; mirrors the public fields of Serial/material.h:18-25.   */
typedef struct oracle_material {
    float color[3];
    float ka, kd, ks, kr, kt, eta;
    uint32_t barycentric; /* 1: BarycentricMaterial (material.h:35-50); the
                             three vertex colours come from tri_rgb          */
} oracle_material;

typedef struct oracle_scene {
    /* triangles: 9 floats each (ax ay az bx by bz cx cy cz)                 */
    const float*    tri_v;
    const uint32_t* tri_material;   /* index into materials                  */
    const float*    tri_rgb;        /* 9 floats per triangle or NULL         */
    const uint32_t* tri_object_id;  /* index in World::objectList, or NULL
                                       (then 0..n_tri-1)                     */
    uint32_t        n_tri;
    /* spheres: cx cy cz r                                                   */
    const float*    sph;  const uint32_t* sph_material; const uint32_t* sph_object_id; uint32_t n_sph;
    /* planes (quads): 4 corners, 12 floats                                  */
    const float*    pln;  const uint32_t* pln_material; const uint32_t* pln_object_id; uint32_t n_pln;
    /* cylinders: px py pz r ux uy uz                                        */
    const float*    cyl;  const uint32_t* cyl_material; const uint32_t* cyl_object_id; uint32_t n_cyl;
    const oracle_material* materials; uint32_t n_materials;
    const float*    lights;         /* 6 floats each: position, intensity    */
    uint32_t        n_lights;
    float           ambient[3];
    float           background[3];
} oracle_scene;

/* Camera as the reference constructs it (Serial/camera.h:25).               */
typedef struct oracle_camera {
    double pos[3], target[3], up[3];
    float  fovy;
    int32_t width, height;
} oracle_camera;

enum {
    ORACLE_MODE_AS_SHIPPED   = 0, /* uniform grid incl. the BBox init quirk  */
    ORACLE_MODE_BBOX_FIXED   = 1, /* grid with axis_max = lowest()           */
    ORACLE_MODE_TRUE_NEAREST = 2  /* linear loop over every object           */
};

typedef struct oracle_result {
    uint64_t rays_total;      /* World::firstIntersection calls while shading */
    uint64_t rays_primary;    /* port only (0 from the reference build)       */
    uint64_t rays_shadow;     /* port only                                    */
    uint64_t rays_secondary;  /* port only                                    */
    double   build_seconds;   /* scene + grid construction                    */
    double   render_seconds;  /* shading loop only                            */
    uint32_t columns_rendered;
    uint32_t threads_used;
} oracle_result;

/*
 * Render columns col_begin, col_begin+col_step, ... of the frame.
 *   rgb      W*H*3 bytes, byte (i + j*W)*3 (Serial/camera.cpp:46-52); columns
 *            that are not rendered are left untouched.  May not be NULL.
 *   prim_id  optional W*H int32: objectList index of the primary hit, -1 miss
 *   t_hit    optional W*H float: primary-ray parameter, FLT_MAX on a miss
 *   nthreads 1 = the reference's single-threaded loop; >1 splits the selected
 *            columns round-robin over std::threads
 * Returns 0 on success, negative on bad arguments.
 */
int oracle_render(const oracle_scene* scene, const oracle_camera* cam,
                  int max_depth, int mode,
                  int col_begin, int col_step, int nthreads,
                  uint8_t* rgb, int32_t* prim_id, float* t_hit,
                  oracle_result* result);

/*
 * Nearest-hit queries for explicit rays (KATs).  rays: 6 floats each
 * (origin, direction; the direction is normalised like Ray's constructor).
 * Outputs per ray: objectList index or -1, and t (FLT_MAX on miss).
 */
int oracle_trace_rays(const oracle_scene* scene, int mode,
                      const float* rays, uint32_t n_rays,
                      int32_t* prim_id, float* t_hit);

/* Shade explicit rays with World::shade_ray at level 0; rgb_out 3 doubles
 * per ray, unclamped.                                                       */
int oracle_shade_rays(const oracle_scene* scene, int max_depth, int mode,
                      const float* rays, uint32_t n_rays, double* rgb_out);

const char* oracle_name(void);

#ifdef __cplusplus
}
#endif
#endif
