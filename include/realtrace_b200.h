/*
 * realtrace_b200.h — C ABI of the B200-native ray-tracing core for RealTrace.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.
 * The reference (rjalfa/RealTrace, Serial/) has no FFI of its own — its render
 * core is called in-process by Serial/lumina.cpp — so each entry point below
 * names the reference interface it replaces (file:line under
 * /root/reference/Serial).  The C++ classes of realtrace_b200/host/ (World,
 * Triangle, Sphere, Plane, Cylinder, Material, PointLightSource, Camera,
 * RenderEngine) keep the reference's signatures and call only these functions.
 *
 * Conventions
 *   - every function returns 0 on success or a negative rt_status; the text of
 *     the last error is available from rt_last_error()
 *   - one rt_ctx per process and per GPU; a context is not thread-safe
 *   - primitive ids are indices into World::objectList in insertion order
 *     (world.h:34-37); -1 means "no hit"
 *   - frame layout: byte (i + j*W)*3 + c, row j = 0 is the bottom row
 *     (camera.cpp:46-52)
 *   - there is no CPU fallback: without a CUDA device rt_create fails
 */
#ifndef REALTRACE_B200_H
#define REALTRACE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_ctx rt_ctx;

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID_ARGUMENT = -1,
    RT_ERR_NO_DEVICE = -2,
    RT_ERR_CUDA = -3,
    RT_ERR_NOT_COMMITTED = -4,
    RT_ERR_QUEUE_OVERFLOW = -5,
    RT_ERR_OUT_OF_MEMORY = -6
} rt_status;

/* Material — the public data fields of class Material (material.h:18-25).
 * `n` (Phong exponent) is omitted: the reference never reads it
 * (world.cpp:134 hard-codes 128).                                            */
typedef struct rt_material {
    float color[3];
    float ka, kd, ks, kr, kt, eta;
    uint32_t flags;               /* RT_MATERIAL_BARYCENTRIC: BarycentricMaterial
                                     (material.h:35-50); colours come from the
                                     triangle's vertex_rgb                     */
} rt_material;
#define RT_MATERIAL_BARYCENTRIC 1u

/* Camera — the state Camera's constructor derives (camera.cpp:4-25): basis
 * u, v, w (w = -line of sight), focalDistance and aspect (both float members,
 * camera.h:20-22).                                                           */
typedef struct rt_camera {
    float pos[3];
    float u[3], v[3], w[3];
    float focal_distance;
    float aspect;
    int32_t width, height;
} rt_camera;

typedef struct rt_render_params {
    int32_t max_depth;            /* replaces #define RECURSION_DEPTH (world.h:11):
                                     rays with level > max_depth return background  */
    int32_t tile_w, tile_h;       /* screen tile; 0 = default 64 x 32; multiples of 8 / 4 */
    int32_t rank, world_size;     /* interleaved tile ownership: this context renders the
                                     tiles with (tile_id % world_size) == rank; {0,1} = all */
    uint32_t flags;
    /* Dynamic tile stealing (multi-GPU, needs a frame every rank can write, see
     * rt_shared_buffer_*): with steal_pool_div = k > 0 every k-th group of world_size tiles is not
     * owned by anybody; ranks that run out of their own tiles claim 8x4-pixel blocks of those
     * tiles from the shared cursor (64 x uint32, zero-initialised, slot frame_index % 64) with
     * system-scope atomics.  All ranks must pass the same k, cursor and frame_index.           */
    int32_t steal_pool_div;
    uint32_t frame_index;
    void* steal_cursor;
} rt_render_params;
#define RT_FLAG_BRUTE_FORCE   1u  /* test every triangle linearly instead of walking the BVH */
#define RT_FLAG_COUNT_WORK    2u  /* count BVH node visits / triangle tests (slower)         */
#define RT_FLAG_WARP_TIMES    8u  /* debug: record start/end time of every warp of the primary
                                     traversal kernel (read back with rt_debug_warp_times)      */
#define RT_FLAG_PACKED_TILES  4u  /* output buffer holds this rank's tiles back to back
                                     (tile-local row-major RGB8) instead of a full frame    */

typedef struct rt_aux_out {       /* optional per-pixel primary-hit outputs, W*H each, index i + j*W */
    int32_t* prim_id;
    float*   t;
} rt_aux_out;

typedef struct rt_frame_stats {
    uint64_t rays_primary;        /* one ray = one World::firstIntersection call (world.cpp:35,:46) */
    uint64_t rays_shadow;
    uint64_t rays_secondary;
    uint64_t node_visits;         /* RT_FLAG_COUNT_WORK only: nearest-hit rays (primary + secondary) */
    uint64_t tri_tests;           /* RT_FLAG_COUNT_WORK only: nearest-hit rays */
    uint64_t shadow_node_visits;  /* RT_FLAG_COUNT_WORK only: any-hit (shadow) rays */
    uint64_t shadow_tri_tests;    /* RT_FLAG_COUNT_WORK only: any-hit (shadow) rays */
    uint32_t waves;               /* wavefront iterations (bounce generations) */
    uint32_t tiles;               /* tiles rendered by this context */
    uint32_t kernel_launches;     /* CUDA kernels launched for this frame */
    uint32_t max_queue;           /* largest ray population of any wave */
    uint32_t stolen_blocks;       /* 8x4-pixel blocks this context claimed from the shared pool */
    uint32_t reserved0;
    float ms_device;              /* device time of the frame, CUDA events on the render stream */
    float ms_trace;               /* device time of the primary nearest-hit kernel */
    float ms_shadow;              /* device time of the primary wave's shadow any-hit kernel */
    float ms_shade;               /* device time of the primary wave's shading kernel */
    float ms_secondary;           /* device time of all bounce waves (trace + shadow + shade) */
    float ms_resolve;
} rt_frame_stats;

typedef struct rt_build_stats {
    uint32_t n_triangles;         /* triangles in the LBVH */
    uint32_t n_large_triangles;   /* outliers tested linearly */
    uint32_t n_nodes;
    uint32_t sort_passes;
    uint32_t leaf_size;           /* triangles per leaf at most (RT_LEAF_SIZE, default 1) */
    float ms_build;               /* device time: bounds + Morton + sort + hierarchy + refit */
    float ms_refit;               /* device time of the last refit */
} rt_build_stats;

/* ---- life cycle -------------------------------------------------------------------------- */
/* device: CUDA ordinal.  Replaces nothing in the reference (single-threaded CPU code).        */
int rt_create(rt_ctx** out, int device);
/* One context that drives n_devices GPUs from the calling thread (one rt_ctx per process drives N
 * GPUs): the scene is replicated on every device, each frame is cut into interleaved screen tiles (tile t belongs
 * to device t % n), every device stores its finished pixels into the frame that lives on device_ids[0] over
 * NVLink peer access (cudaDeviceEnablePeerAccess), and a flag handshake in peer memory completes the frame.
 * device_ids == NULL means 0 .. n_devices-1; n_devices == 0 means every visible device.  Every rt_* call works
 * on such a context: scene calls are mirrored to all devices, rt_scene_commit builds the LBVH on all of them
 * concurrently, rt_render / rt_render_device / rt_render_enqueue split the frame; per-ray queries and
 * introspection run on device_ids[0].  With one device it is exactly rt_create.                              */
int rt_create_multi(rt_ctx** out, const int* device_ids, int n_devices);
int rt_device_count(rt_ctx* ctx, int32_t* n_devices);
int rt_destroy(rt_ctx* ctx);
/* Page-locked host memory for frames (Camera::bitmap, camera.cpp:19): the device-to-host copy of a frame into
 * such a buffer is an asynchronous DMA at full PCIe rate; a pageable buffer (the reference's new[]) is staged
 * by the driver.  rt_render page-locks a pageable rgb_out on first use (cudaHostRegister) where it can.     */
int rt_host_alloc(void** out, uint64_t bytes);
int rt_host_free(void* p);
const char* rt_last_error(rt_ctx* ctx);      /* ctx may be NULL: last error of rt_create        */
/* Run all further work of this context on an existing CUDA stream (a cudaStream_t passed as a
 * pointer, e.g. torch.cuda.current_stream().cuda_stream); NULL restores the context's own stream.
 * The legacy default stream is named by CUDA's handle cudaStreamLegacy, (void*)0x1.           */
int rt_set_stream(rt_ctx* ctx, void* cuda_stream);

/* ---- scene: replaces World::addObject / addLight / setAmbient / setBackground (world.h:26-37)
 * and the object constructors.  Arrays are copied; object_id may be NULL (triangles: 0..n-1;
 * other kinds: numbered after the triangles in the order spheres, planes, cylinders).          */
/* Triangle(a,b,c,mat) triangle.h:22 — v: 9 floats per triangle; vertex_rgb: 9 floats per
 * triangle (BarycentricMaterial colours, material.h:44) or NULL.                               */
int rt_scene_set_triangles(rt_ctx* ctx, const float* v, const uint32_t* material_id, const float* vertex_rgb,
                           const uint32_t* object_id, uint32_t n);
/* Sphere(pos, r, mat) sphere.h:17 — 4 floats: cx cy cz r                                        */
int rt_scene_set_spheres(rt_ctx* ctx, const float* packed, const uint32_t* material_id, const uint32_t* object_id,
                         uint32_t n);
/* Plane(p1,p2,p3,p4, mat) plane.h:21 — 12 floats                                                */
int rt_scene_set_planes(rt_ctx* ctx, const float* packed, const uint32_t* material_id, const uint32_t* object_id,
                        uint32_t n);
/* Cylinder(pos, r, up, mat) cylinder.h:17 — 7 floats: px py pz r ux uy uz                       */
int rt_scene_set_cylinders(rt_ctx* ctx, const float* packed, const uint32_t* material_id, const uint32_t* object_id,
                           uint32_t n);
int rt_scene_set_materials(rt_ctx* ctx, const rt_material* m, uint32_t n);
/* PointLightSource(world, pos, intensity) pointlightsource.h:11 — 6 floats per light            */
int rt_scene_set_lights(rt_ctx* ctx, const float* pos_rgb, uint32_t n);
int rt_scene_set_environment(rt_ctx* ctx, const float ambient[3], const float background[3]);

/* Replaces `world->uniform_grid = UniformGrid(all_triangles)` (lumina.cpp:289,
 * uniform-grid.cpp:54-147): uploads the scene and builds the LBVH on the GPU.
 * RT_COMMIT_REFIT keeps the hierarchy and recomputes boxes from the current vertices
 * (set with rt_scene_update_vertices).                                                         */
#define RT_COMMIT_BUILD 0
#define RT_COMMIT_REFIT 1
int rt_scene_commit(rt_ctx* ctx, int mode);
int rt_scene_update_vertices(rt_ctx* ctx, const float* v, uint32_t n);
/* The same from DEVICE memory (vertices deformed by the caller's own kernel): an asynchronous device-to-device
 * copy on the context's stream, no host round trip; follow with rt_scene_commit(RT_COMMIT_REFIT).  A later
 * RT_COMMIT_BUILD without new rt_scene_set_triangles rebuilds from these vertices.                  */
int rt_scene_update_vertices_device(rt_ctx* ctx, const float* v_dev, uint32_t n);
int rt_scene_build_stats(rt_ctx* ctx, rt_build_stats* out);

/* ---- rendering: replaces `while(!engine->renderLoop()){}` (renderengine.cpp:10-26,
 * lumina.cpp:464) plus Camera::drawPixel into Camera::getBitmap().
 * rt_render: rgb_out / aux are HOST buffers (W*H*3 bytes; aux arrays W*H).
 * rt_render_device: rgb_out_dev / aux are DEVICE buffers; nothing is copied to the host and
 * the call returns once the work is enqueued and its statistics (if requested) are read.       */
int rt_render(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, uint8_t* rgb_out,
              const rt_aux_out* aux, rt_frame_stats* stats);
int rt_render_device(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, void* rgb_out_dev,
                     const rt_aux_out* aux_dev, rt_frame_stats* stats);

/* rt_render_device with stats == NULL only enqueues the frame on the context's stream and returns;
 * rt_synchronize waits for it and reports errors raised by the kernels meanwhile.              */
int rt_synchronize(rt_ctx* ctx);
/* Pipelined host-buffer frames: rt_render_enqueue enqueues the frame AND its copy into rgb_out (page-locked:
 * rt_host_alloc, else it is page-locked here) and returns; rt_render_wait(slot) blocks until that frame has
 * arrived.  Two slots (0, 1) with their own device frames: enqueue frame k+1 into the other slot before waiting
 * for frame k and the copy of frame k overlaps the rendering of frame k+1 (the display path of
 * Parellel/main.cu:62-70,122-133 without GL: the host buffer stands in for the PBO).                        */
int rt_render_enqueue(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, uint8_t* rgb_out, int32_t slot);
int rt_render_wait(rt_ctx* ctx, int32_t slot);

/* ---- multi-GPU (one process per GPU)                                                         */
/* A device buffer other processes on the node can map (CUDA IPC over NVLink peer access): the
 * owner creates it and publishes the 64-byte handle; peers open it and pass the mapped pointer as
 * rgb_out_dev of rt_render_device, so that their resolve kernel stores its tiles straight into the
 * owner's frame.  rt_download copies device memory of this context to the host.                */
int rt_shared_buffer_create(rt_ctx* ctx, uint64_t bytes, void** dev_ptr, unsigned char handle[64]);
int rt_shared_buffer_open(rt_ctx* ctx, const unsigned char handle[64], void** dev_ptr);
int rt_download(rt_ctx* ctx, const void* dev_ptr, void* host, uint64_t bytes);
/* Frame-completion handshake through a shared buffer (>= 1 KiB from rt_shared_buffer_create on rank 0,
 * opened by the others) instead of a collective: enqueue phase 0 before and phase 1 after the kernels
 * that write frame `frame_index` into rank 0's memory.
 *   phase 0: rank 0 publishes "frames < frame_index are consumed" — in stream order, i.e. after everything it
 *            enqueued for the previous frame (a download, a display copy); ranks > 0 wait for that.
 *   phase 1: ranks > 0 signal their arrival; rank 0 waits for world_size - 1 arrivals.
 * After phase 1 has run on rank 0's stream the frame is complete there; ranks > 0 never run more than one
 * frame ahead of rank 0.  frame_index must count 0, 1, 2, ... identically on all ranks.  The waiting kernels
 * are single threads that give up after ~2 s (reported by the next synchronising call).          */
int rt_peer_sync(rt_ctx* ctx, void* sync_buf, int32_t rank, int32_t world_size, uint32_t frame_index, int32_t phase);
/* A symmetric barrier over the same buffer: every rank enqueues it with the same epoch = 0, 1, 2, ...     */
int rt_peer_barrier(rt_ctx* ctx, void* sync_buf, int32_t world_size, uint32_t epoch);
/* One call for a whole multi-GPU frame step, all of it only enqueued (p->flags must carry RT_FLAG_PACKED_TILES).
 * Scene with mirror/dielectric bounces: render this rank's tiles into the local packed buffer, rt_peer_sync
 * phase 0, push the tiles into the shared frame, rt_peer_sync phase 1.  Bounce-free scene (1..8 lights): phase 0,
 * ONE kernel that traces, shades and stores every finished 8x4 pixel block straight into frame_dev (packed_dev
 * is not written), phase 1.                                                                               */
int rt_render_push(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, void* packed_dev, void* frame_dev,
                   void* sync_buf, uint32_t frame_index);
/* tile helpers for the gather-based assembly (the gather itself is NCCL, outside)               */
/* number of tiles in the frame / owned by `rank`, and bytes of one packed tile                  */
int rt_tile_layout(int32_t width, int32_t height, int32_t tile_w, int32_t tile_h, int32_t rank, int32_t world_size,
                   uint32_t* tiles_total, uint32_t* tiles_owned, uint32_t* tile_bytes);
/* Scatter packed tiles (device) of rank `src_rank` into a full frame (device, W*H*3).          */
int rt_assemble_tiles(rt_ctx* ctx, const void* packed_dev, int32_t src_rank, int32_t world_size, int32_t width,
                      int32_t height, int32_t tile_w, int32_t tile_h, void* frame_dev);

/* ---- per-ray queries: World::firstIntersection (world.cpp:5-17) and World::shade_ray
 * (world.cpp:32-111) for batches of explicit rays.  rays: 6 floats each (origin, direction;
 * the direction is normalised like Ray's constructor, ray.h:25-29).  Host buffers.            */
int rt_trace_rays(rt_ctx* ctx, const float* rays, uint32_t n, uint32_t flags, int32_t* prim_id, float* t);
int rt_shade_rays(rt_ctx* ctx, const float* rays, uint32_t n, int32_t max_depth, uint32_t flags, float* rgb_out);

/* ---- introspection for the structure tests.  Any pointer may be NULL.
 * nodes: n_nodes * 16 floats (4 x float4 per node); tri_order: sorted position ->
 * index into the triangle array given to rt_scene_set_triangles; keys: sorted Morton keys.     */
int rt_bvh_download(rt_ctx* ctx, float* nodes, uint32_t* tri_order, uint64_t* keys, uint32_t* n_nodes,
                    uint32_t* n_bvh_triangles);

/* Roofline denominators measured on this context's device, a few hundred ms:
 *   l2_read_gbs             float4 __ldg sweeps over a 64 MiB buffer, all SMs (L2-resident after the first sweep)
 *   l2_random_node_gbs      random 64-byte records (the BVH node format) out of the same buffer, one dependent
 *                           chain per thread at full occupancy; l2_dependent_fetch_ns = time of one chain step
 *   hbm_read_gbs            one sweep over 2 GiB
 *   fma_lane_instr_per_s    8 independent FFMA chains per thread at full occupancy = 128 lanes x SMs x real clock;
 *                           issue_warp_instr_per_s = that / 32 (one warp instruction per scheduler and clock)   */
typedef struct rt_microbench_result {
    double l2_read_gbs, l2_random_node_gbs, l2_dependent_fetch_ns, hbm_read_gbs;
    double fma_lane_instr_per_s, issue_warp_instr_per_s, implied_sm_mhz;
    int32_t sm_count, l2_buffer_mib;
} rt_microbench_result;
int rt_microbench(rt_ctx* ctx, rt_microbench_result* out);

/* Kernels enqueued so far by rt_render / rt_render_device / rt_render_push / rt_peer_sync / rt_assemble_tiles
 * and the refit kernels of rt_scene_commit (a host-side running count, no synchronisation; rt_peer_barrier, the
 * sort/hierarchy kernels of a full build and per-ray queries are not counted):
 * the difference across a timed loop of asynchronous frames is its exact launch count.          */
int rt_debug_frame_launches(rt_ctx* ctx, uint64_t* n_kernels);

/* Debug timeline (RT_FLAG_WARP_TIMES): out receives {start_ns, end_ns} per warp of the last primary
 * traversal kernel; *n_warps in: capacity, out: warps written.                                   */
int rt_debug_warp_times(rt_ctx* ctx, uint64_t* out, uint32_t* n_warps);

/* Device radix sort used by the builder, exposed for its own test: sorts (key, value) pairs
 * given in HOST memory, stable, ascending.                                                     */
int rt_debug_sort_pairs(rt_ctx* ctx, uint64_t* keys, uint32_t* values, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
