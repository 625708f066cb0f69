// rt_context.h — host-side state behind an rt_ctx (internal; not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/realtrace_b200.h"
#include "rt_bvh.h"
#include "rt_scene.h"

// ---- error plumbing ---------------------------------------------------------------------------
struct RtError {
    int code;
    std::string msg;
};

#define RT_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char buf__[512];                                                                       \
            snprintf(buf__, sizeof buf__, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,     \
                     cudaGetErrorString(e__));                                                     \
            throw RtError{RT_ERR_CUDA, buf__};                                                     \
        }                                                                                          \
    } while (0)

// ---- a growable device buffer -----------------------------------------------------------------
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    void reserve(size_t n) {
        if (n <= cap) return;
        release();
        cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            char buf[256];
            snprintf(buf, sizeof buf, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
            throw RtError{RT_ERR_OUT_OF_MEMORY, buf};
        }
        cap = n;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

// ---- per-wave device counters -------------------------------------------------------------------
struct WaveCounters {
    uint32_t n_rays;        // population of this wave (written by the previous wave's shade kernel)
    uint32_t fetch_trace;   // persistent-thread work cursor of the trace kernel
    uint32_t n_hits;        // hit-queue length
    uint32_t fetch_shade;   // work cursor of the shadow (any-hit) kernel
    uint32_t n_next;        // rays appended for the next wave
    uint32_t flags;         // unused (errors go to the context's sticky word)
    uint32_t pad[2];
};
#define RT_WAVE_SLOTS 64

struct FrameCounters {
    unsigned long long rays_primary, rays_shadow, rays_secondary;
    unsigned int stolen_blocks, pad;
    unsigned long long node_visits[2], tri_tests[2];   // [0] nearest-hit rays, [1] any-hit (shadow) rays
};

// k_frame's own counters (two sets: the launch of frame k clears the set of frame k+1)
struct FrameKernelCounters {
    WaveCounters wave;
    FrameCounters fc;
};

// Ray queue in SoA float4 records (48 B per ray).
struct RayQueue {
    float4* o_pix;   // origin.xyz, as_float(local pixel index)
    float4* d_lvl;   // direction.xyz, as_float(level)
    float4* w;       // throughput rgb, unused
};

struct TileLayout {
    int width = 0, height = 0, tile_w = 0, tile_h = 0, rank = 0, world = 1, pool_div = 0;
    int tiles_x = 0, tiles_y = 0;
    uint32_t n_tiles_owned = 0;    // statically owned by this rank
    uint32_t n_pool_tiles = 0;     // shared pool (dynamic stealing)
    bool operator==(const TileLayout& o) const {
        return width == o.width && height == o.height && tile_w == o.tile_w && tile_h == o.tile_h && rank == o.rank &&
               world == o.world && pool_div == o.pool_div;
    }
};

#define RT_MAX_DEVICES 16

struct rt_ctx {
    int device = 0;
    // ---- one context driving several GPUs (rt_create_multi): the context the caller holds is rank 0, `kids` are
    // ranks 1..n-1 on the other devices (scene replicated, frame split into interleaved tiles); kids have parent set
    std::vector<rt_ctx*> kids;
    rt_ctx* parent = nullptr;
    void* mg_sync = nullptr;               // handshake words in rank 0's memory (rt_peer_sync layout), mapped by every rank
    struct RtRankPool* mg_pool = nullptr;  // one enqueue thread per extra device (multi_device.cu), created by the first frame
    int mg_threads = 0;                    // RT_MULTI_THREADS: 0 = the caller's thread issues every rank's calls (default: measured equal), 1 = threads from 4 devices on, 2 = always
    uint32_t mg_frame = 0;                 // frames rendered through the multi-device path (handshake slot counter)
    cudaEvent_t mg_ev[2] = {};
    std::vector<void*> host_registered;    // caller buffers page-locked on first use by rt_render_enqueue (unregistered at destroy)
    std::vector<void*> host_checked;       // destinations whose memory kind has been looked up already
    cudaStream_t copy_stream = nullptr;    // frame -> host copies of this rank (its own PCIe link), overlapping the next frame
    DevBuf<float> d_cam_tab;               // window coordinate per column, then per row (k_camera_tables), for the cached size/aspect
    int cam_tab_W = 0, cam_tab_H = 0;
    double cam_tab_aspect = 0.0;
    DevBuf<uint8_t> d_packed;              // this rank's tiles back to back (scenes with bounces: resolve target before the push)
    // A frame every rank can store into: for n > 1 a virtual address range whose granules are physically spread
    // round robin over the ranks' GPUs (multi_device.cu), else a plain allocation on this device.
    struct SharedFrame {
        void* va = nullptr;
        size_t bytes = 0, gran = 0, granules = 0;
        bool vmm = false;
        std::vector<unsigned long long> handles;   // CUmemGenericAllocationHandle per granule
        std::vector<int> owner;                    // rank whose GPU backs the granule
    };
    struct FrameSlot {                     // rt_render_enqueue / rt_render_wait: two frames in flight
        SharedFrame frame;
        cudaEvent_t ready = nullptr;       // frame complete in device memory (rank 0's render stream)
        cudaEvent_t done[RT_MAX_DEVICES] = {};     // rank r's share of the host copy has landed
        uint32_t* h_sticky = nullptr;      // pinned, RT_MAX_DEVICES words: each rank's error word after the frame
        bool in_flight = false;
    };
    FrameSlot slots[2];
    bool pipelined = false;                // rt_render_enqueue has been used: keep both slots allocated
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;
    int sm_count = 148;

    // ---- host staging of the scene (copied by the rt_scene_set_* calls)
    std::vector<float> h_tri_v;            // 9 per triangle
    bool host_vertices_stale = false;      // the device copy is newer (rt_scene_update_vertices_device)
    std::vector<uint32_t> h_tri_mat, h_tri_obj;
    std::vector<float> h_tri_rgb;          // 9 per triangle or empty
    std::vector<AnalyticPrim> h_spheres, h_planes, h_cylinders;
    std::vector<rt_material> h_materials;
    std::vector<float> h_lights;           // 6 per light
    float ambient[3] = {0, 0, 0}, background[3] = {0, 0, 0};
    bool has_reflective = false, has_dielectric = false;
    bool committed = false;

    // ---- device scene
    uint32_t n_tri = 0, n_bvh = 0, n_large = 0, n_fixed_analytic = 0;
    DevBuf<float> d_tri_v;
    DevBuf<uint32_t> d_tri_mat, d_tri_obj;
    DevBuf<float4> d_tri_rgb, d_tris, d_nodes, d_nodes4, d_box_lo, d_box_hi, d_materials, d_lights;
    int wide_bvh = 1;                      // 4-wide collapse of the tree: 0 off, 1 for k_paths (bounce paths), 2 everywhere (RT_WIDE_BVH)
    int wide_heavy = 1;                    // RT_WIDE_HEAVY: the heaviest tiles of a share walk the 4-wide view (k_frame / k_frame_push):
                                           // 0 never, 1 shares of a multi-rank frame, 2 every frame with a cost-sorted tile order
    int wide_heavy_div = 64;               // RT_WIDE_HEAVY_DIV: ... the first n_tiles_owned / div tiles of the heavy-tiles-first order
    int wide_after_bursts = 0;             // RT_WIDE_AFTER_BURSTS: ... and any batch still running after this many bursts of RT_LOOP_PRIMARY
                                           // steps continues on the wide view (0 = off)
    bool want_nodes4 = false;              // the 4-wide view is kept current for the heavy tiles (set by the first frame that wants it)
    DevBuf<AnalyticPrim> d_analytic;
    DevBuf<uint64_t> d_keys[2];
    DevBuf<uint32_t> d_vals[2];
    int sorted_buf = 0;                    // which of d_keys/d_vals holds the sorted result
    DevBuf<uint32_t> d_hist, d_digit_hist, d_bounds, d_misc;
    DevBuf<KarrasNode> d_karras;
    DevBuf<int> d_leaf_parent, d_node_parent;
    DevBuf<uint32_t> d_visit;
    SceneDev scene{};
    rt_build_stats build_stats{};
    int leaf_size = 1;                     // measured best on B200 with the if-if walk (profiles/r1_tuning.md)

    // ---- render state
    TileLayout layout;
    DevBuf<uint32_t> d_tile_ids, d_pool_ids, d_stolen_map, d_tile_cost;
    DevBuf<uint32_t> d_tile_ids0, d_tile_ord;   // the canonical (ascending) tile list; ordinal of the tile at each position of d_tile_ids
    int path_share = 1;                    // k_paths: parked rays go to idle lanes of the warp (RT_PATH_SHARE)
    int tile_sort_every = 8;               // RT_TILE_SORT_EVERY: the order is renewed after the first two frames of a layout, then every k-th
    bool tile_cost_valid = false;
    uint32_t frames_in_layout = 0;
    int tile_feedback = 1;                 // heavy-tiles-first reordering from last frame's cost (RT_TILE_FEEDBACK)
    DevBuf<uint32_t> d_local_cursor;       // stand-in steal cursor when the caller passes none
    bool local_cursor_valid = false;
    uint32_t local_cursor_frame = 0;
    DevBuf<long long> d_accum;             // 3 per local pixel, 32.32 fixed point
    DevBuf<float4> d_q[2][3];              // two ray queues x (o_pix, d_lvl, w)
    DevBuf<float4> d_hits;                 // HitRec per ray slot
    DevBuf<uint32_t> d_hitq;
    DevBuf<uint8_t> d_occl;                // shadow results, [light][hit]
    size_t queue_cap = 0;
    DevBuf<WaveCounters> d_waves;
    DevBuf<FrameCounters> d_frame;
    DevBuf<uint8_t> d_rgb;
    DevBuf<int32_t> d_aux_prim;
    DevBuf<float> d_aux_t;
    DevBuf<float> d_rays_in;               // rt_trace_rays / rt_shade_rays staging
    DevBuf<float> d_rgbf_out;
    DevBuf<unsigned long long> d_warp_times;   // debug timeline of the primary traversal kernel
    WaveCounters* h_waves = nullptr;       // pinned
    FrameCounters* h_frame = nullptr;      // pinned
    DevBuf<uint32_t> d_sticky;             // bit0 ray-queue overflow, bit1 traversal stack overflow
    uint32_t* h_sticky = nullptr;          // pinned
    std::vector<void*> ipc_opened, ipc_created;
    cudaEvent_t ev[10] = {};               // 0-3,6,7 frame; 4,5 build; 8,9 refit
    bool refit_pending = false;
    int trace_blocks = 0, fused_blocks = 0, shadow_blocks = 0, shade_blocks = 0, path_blocks = 0;
    int wide_blocks = 0, path_wide_blocks = 0, fused_shade_blocks = 0, wide_shade_blocks = 0;
    int tile_bucket_bits = 5;              // heavy-tiles-first: mantissa bits of the cost kept in the sort key (RT_TILE_BUCKET_BITS)
    int fuse_shade = 1;                    // primary hits shaded inside the fused traversal kernel (RT_FUSE_SHADE): 0 never,
                                           // 1 when that kernel also pushes finished tiles to a remote frame, 2 always
    bool remote_output = false;            // rt_render_push: the frame being rendered into sits in another GPU's memory
    int frame_kernel = 2;                  // whole bounce-free frames in one k_frame launch (RT_FRAME_KERNEL): 0 never, 1 when pushed to a shared frame, 2 always
    int frame_blocks = 0, frame_push_blocks = 0, frame_blocks_dense = 0, frame_push_blocks_dense = 0;
    long long dense_min_pixels = 1500000;  // shares of at least this many pixels run the 9-CTAs-per-SM build of k_frame / k_frame_push (RT_DENSE_MIN_PIXELS)
    int push_inline = 1;                   // multi-GPU bounce-free frames: finished 8x4 blocks go straight into the shared frame (k_frame_push) instead of a push phase at the end (RT_PUSH_INLINE)
    DevBuf<uint32_t> d_fsync;              // k_frame's phase counters
    uint32_t fsync_target[3] = {0, 0, 0};  // where the cumulative barrier counters stand after the launch being built (wrap)
    DevBuf<FrameKernelCounters> d_fk;      // k_frame's double-buffered counters
    uint32_t fk_epoch = 0;
    FrameCounters* last_frame_ctr = nullptr;   // device counters of the frame enqueued last (d_frame or a d_fk set)
    struct PushTarget {                    // set by rt_push_frame around rt_render_frame: where this rank's tiles go
        void* frame = nullptr;             // the shared frame (nullptr: not pushing)
        void* sync = nullptr;
        uint32_t frame_index = 0;
        int rank = 0, world = 1;
        bool done = false;                 // rt_render_frame pushed (and signalled) inside its one kernel
    } push;
    uint64_t launch_total = 0;             // kernels enqueued by frame-level calls since rt_create (rt_debug_frame_launches)
    int path_kernel = 1;                   // bounce generations in one k_paths launch (RT_PATH_KERNEL=0: wave loop)
    int fuse_shadow = 1;                   // shadow rays ride in the lane that found the hit (RT_FUSE_SHADOW)
    int refill_primary_fused = 32;
    int blocks_per_sm = 0;                 // 0 = as many persistent CTAs as fit
    // idle lanes a warp waits for before fetching new rays (k_traverse); measured on B200 (profiles/r1_tuning.md):
    // coherent primary rays are best refilled as whole warps, shadow and bounce rays lane by lane in groups
    int refill_primary = 32, refill_queue = 16, refill_shadow = 16;
    int loop_primary = 64, loop_queue = 64, loop_shadow = 64;   // k_traverse loop style per ray class (0 = while-while)
};

// ---- entry points implemented across the .cu files ------------------------------------------------
void rt_build_bvh(rt_ctx* c, bool refit_only);                      // bvh_build.cu
void rt_sort_pairs_device(rt_ctx* c, uint32_t n);   // radix_sort.cu: d_keys[0]/d_vals[0] -> sorted_buf; only enqueues
void rt_multi_pool_stop(rt_ctx* c);                  // multi_device.cu: joins the per-rank enqueue threads
void rt_ensure_nodes4(rt_ctx* c);                    // bvh_build.cu: builds the 4-wide view now if the scene has none (enqueues)
int rt_sort_passes_done(rt_ctx* c);                  // passes of the last sort that moved keys (synchronises)
void rt_render_frame(rt_ctx* c, const rt_camera* cam, const rt_render_params* p, void* rgb_dev,
                     const rt_aux_out* aux_dev, rt_frame_stats* stats);   // render.cu
void rt_query_rays(rt_ctx* c, const float* rays_host, uint32_t n, int max_depth, uint32_t flags, bool shade,
                   int32_t* prim_out, float* t_out, float* rgb_out);      // render.cu
void rt_render_init(rt_ctx* c);                                     // render.cu
void rt_sync_and_check(rt_ctx* c);                                  // render.cu
void rt_peer_barrier_enqueue(rt_ctx* c, void* sync_buf, int world, uint32_t epoch);   // render.cu
void rt_peer_sync_enqueue(rt_ctx* c, void* sync_buf, int rank, int world, uint32_t frame_index, int phase);   // render.cu
bool rt_scene_bounces(const rt_ctx* c, int max_depth);                     // render.cu
bool rt_frame_kernel_ok(const rt_ctx* c, const rt_render_params* p, bool pushing);   // render.cu
bool rt_frame_pushes_inline(const rt_ctx* c, const rt_render_params* p);   // render.cu
void rt_assemble(rt_ctx* c, const void* packed, int src_rank, int world, int width, int height, int tile_w,
                 int tile_h, void* frame);                          // render.cu
RtError rt_sticky_error(uint32_t flags);                             // render.cu: device error word -> error (code RT_OK if 0)
void rt_push_frame(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, void* packed_dev, void* frame_dev,
                   void* sync_buf, uint32_t frame_index, const rt_aux_out* aux_dev);   // api.cu: body of rt_render_push
void rt_run_microbench(rt_ctx* c, rt_microbench_result* out);         // microbench.cu
// multi_device.cu
void rt_frame_reserve(rt_ctx* c, rt_ctx::SharedFrame& f, size_t bytes);
void rt_frame_release(rt_ctx* c, rt_ctx::SharedFrame& f);
void rt_multi_attach(rt_ctx* c, const std::vector<int>& devices, rt_ctx* (*make_ctx)(int device));
void rt_multi_build(rt_ctx* c, bool refit_only);
void rt_multi_enqueue_frame(rt_ctx* c, const rt_camera* cam, const rt_render_params* p, void* frame_dev,
                            const rt_aux_out* aux_dev);
void rt_multi_collect(rt_ctx* c, rt_frame_stats* stats);
void rt_frame_download_async(rt_ctx* c, const rt_ctx::SharedFrame& f, uint8_t* host, size_t bytes, cudaEvent_t ready,
                             cudaEvent_t* done, uint32_t* h_sticky);
