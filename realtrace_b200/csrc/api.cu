// api.cu — the extern "C" boundary (include/realtrace_b200.h).  No exception crosses it: every
// entry point converts failures into a negative rt_status and records the text for rt_last_error.
#include <cmath>
#include <cstring>
#include <ctime>

#include "rt_context.h"

static thread_local std::string g_create_error;

template <class F>
static int guarded(rt_ctx* ctx, F&& body) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    try {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) throw RtError{RT_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)};
        body();
        return RT_OK;
    } catch (const RtError& e) {
        ctx->err = e.msg;
        cudaGetLastError();   // clear a sticky-less error so the next call starts clean
        return e.code;
    } catch (const std::bad_alloc&) {
        ctx->err = "host allocation failed";
        return RT_ERR_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        ctx->err = e.what();
        return RT_ERR_INVALID_ARGUMENT;
    }
}

static void need(bool ok, const char* what) {
    if (!ok) throw RtError{RT_ERR_INVALID_ARGUMENT, what};
}

// rank 0 (the context the caller holds) and the contexts of the other devices (rt_create_multi)
static std::vector<rt_ctx*> all_ranks(rt_ctx* c) {
    std::vector<rt_ctx*> v{c};
    v.insert(v.end(), c->kids.begin(), c->kids.end());
    return v;
}

// A context for one device (throws RtError).
static rt_ctx* make_ctx(int device) {
    rt_ctx* c = new rt_ctx;
    try {
        c->device = device;
        RT_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        RT_CUDA(cudaGetDeviceProperties(&prop, device));
        c->sm_count = prop.multiProcessorCount;
        RT_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        RT_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        for (auto& ev : c->ev) RT_CUDA(cudaEventCreate(&ev));
        for (auto& ev : c->mg_ev) RT_CUDA(cudaEventCreate(&ev));
        const char* ls = getenv("RT_LEAF_SIZE");
        if (ls) {
            int v = atoi(ls);
            if (v >= 1 && v <= RT_LEAF_MAX) c->leaf_size = v;
        }
        auto env_int = [](const char* name, int lo, int hi, int& dst) {
            const char* e = getenv(name);
            if (e) { int v = atoi(e); if (v >= lo && v <= hi) dst = v; }
        };
        env_int("RT_REFILL_PRIMARY", 1, 32, c->refill_primary);
        env_int("RT_REFILL_QUEUE", 1, 32, c->refill_queue);
        env_int("RT_REFILL_SHADOW", 1, 32, c->refill_shadow);
        env_int("RT_BLOCKS_PER_SM", 1, 32, c->blocks_per_sm);
        env_int("RT_WIDE_BVH", 0, 2, c->wide_bvh);
        env_int("RT_MULTI_THREADS", 0, 2, c->mg_threads);
        env_int("RT_WIDE_HEAVY", 0, 2, c->wide_heavy);
        env_int("RT_WIDE_HEAVY_DIV", 1, 65536, c->wide_heavy_div);
        env_int("RT_WIDE_AFTER_BURSTS", 0, 1024, c->wide_after_bursts);
        env_int("RT_FUSE_SHADOW", 0, 1, c->fuse_shadow);
        env_int("RT_FUSE_SHADE", 0, 2, c->fuse_shade);
        env_int("RT_FRAME_KERNEL", 0, 2, c->frame_kernel);
        env_int("RT_PUSH_INLINE", 0, 1, c->push_inline);
        env_int("RT_TILE_BUCKET_BITS", 0, 8, c->tile_bucket_bits);
        env_int("RT_TILE_SORT_EVERY", 1, 1024, c->tile_sort_every);
        env_int("RT_PATH_KERNEL", 0, 1, c->path_kernel);
        env_int("RT_PATH_SHARE", 0, 1, c->path_share);
        env_int("RT_TILE_FEEDBACK", 0, 1, c->tile_feedback);
        env_int("RT_REFILL_PRIMARY_FUSED", 1, 32, c->refill_primary_fused);
        env_int("RT_LOOP_PRIMARY", 0, 1024, c->loop_primary);
        env_int("RT_LOOP_QUEUE", 0, 1024, c->loop_queue);
        env_int("RT_LOOP_SHADOW", 0, 1024, c->loop_shadow);
        if (const char* e = getenv("RT_DENSE_MIN_PIXELS")) c->dense_min_pixels = atoll(e);
        rt_render_init(c);
    } catch (...) {
        delete c;
        throw;
    }
    return c;
}

static void destroy_one(rt_ctx* ctx) {
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& s : ctx->slots) {
        rt_frame_release(ctx, s.frame);
        if (s.ready) cudaEventDestroy(s.ready);
        for (auto& e : s.done)
            if (e) cudaEventDestroy(e);
        if (s.h_sticky) cudaFreeHost(s.h_sticky);
    }
    for (void* p : ctx->host_registered) cudaHostUnregister(p);
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->mg_ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->h_waves) cudaFreeHost(ctx->h_waves);
    if (ctx->h_frame) cudaFreeHost(ctx->h_frame);
    if (ctx->h_sticky) cudaFreeHost(ctx->h_sticky);
    for (void* p : ctx->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void* p : ctx->ipc_created) cudaFree(p);
    if (ctx->mg_sync) cudaFree(ctx->mg_sync);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaGetLastError();
    delete ctx;
}

// One rank's part of a multi-GPU frame, only enqueued (see rt_render_push in the header).
void rt_push_frame(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, void* packed_dev, void* frame_dev,
                   void* sync_buf, uint32_t frame_index, const rt_aux_out* aux_dev) {
    int world = p->world_size > 1 ? p->world_size : 1, rank = world > 1 ? p->rank : 0;
    if (rt_frame_kernel_ok(ctx, p, true)) {
        // bounce-free scene: ONE launch traces, shades, copies this rank's tiles into the shared frame with 16-byte
        // stores (over NVLink where they live on another GPU) and does its part of the handshake (k_frame)
        ctx->push.frame = frame_dev; ctx->push.sync = sync_buf; ctx->push.frame_index = frame_index;
        ctx->push.rank = rank; ctx->push.world = world; ctx->push.done = false;
        try { rt_render_frame(ctx, cam, p, packed_dev, aux_dev, nullptr); }
        catch (...) { ctx->push = rt_ctx::PushTarget{}; throw; }
        bool done = ctx->push.done;
        ctx->push = rt_ctx::PushTarget{};
        if (done) return;
        // (the frame took the multi-kernel path after all: finish like a scene with bounces)
        rt_peer_sync_enqueue(ctx, sync_buf, rank, world, frame_index, 0);
        rt_assemble(ctx, packed_dev, rank, world, cam->width, cam->height, p->tile_w > 0 ? p->tile_w : 64,
                    p->tile_h > 0 ? p->tile_h : 32, frame_dev);
        rt_peer_sync_enqueue(ctx, sync_buf, rank, world, frame_index, 1);
        return;
    }
    if (rt_frame_pushes_inline(ctx, p)) {
        // bounce-free scene: ONE kernel traces, shades and stores every finished 8x4 block straight into the
        // shared frame (over NVLink where that block lives on another GPU); the packed buffer is not used
        rt_render_params q = *p;
        q.flags &= ~(uint32_t)RT_FLAG_PACKED_TILES;
        rt_peer_sync_enqueue(ctx, sync_buf, rank, world, frame_index, 0);
        ctx->remote_output = true;
        try { rt_render_frame(ctx, cam, &q, frame_dev, aux_dev, nullptr); }
        catch (...) { ctx->remote_output = false; throw; }
        ctx->remote_output = false;
        rt_peer_sync_enqueue(ctx, sync_buf, rank, world, frame_index, 1);
        return;
    }
    rt_render_frame(ctx, cam, p, packed_dev, aux_dev, nullptr);
    rt_peer_sync_enqueue(ctx, sync_buf, rank, world, frame_index, 0);
    rt_assemble(ctx, packed_dev, rank, world, cam->width, cam->height, p->tile_w > 0 ? p->tile_w : 64,
                p->tile_h > 0 ? p->tile_h : 32, frame_dev);
    rt_peer_sync_enqueue(ctx, sync_buf, rank, world, frame_index, 1);
}

extern "C" {

int rt_create_multi(rt_ctx** out, const int* device_ids, int n_devices) {
    if (!out) return RT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                         "); realtrace_b200 has no CPU fallback";
        cudaGetLastError();
        return RT_ERR_NO_DEVICE;
    }
    if (n_devices == 0) n_devices = count < RT_MAX_DEVICES ? count : RT_MAX_DEVICES;
    if (n_devices < 0 || n_devices > RT_MAX_DEVICES) {
        g_create_error = "rt_create_multi: between 1 and 16 devices";
        return RT_ERR_INVALID_ARGUMENT;
    }
    std::vector<int> devs((size_t)n_devices);
    for (int i = 0; i < n_devices; i++) {
        devs[i] = device_ids ? device_ids[i] : i;
        bool dup = false;
        for (int j = 0; j < i; j++) dup |= devs[j] == devs[i];
        if (devs[i] < 0 || devs[i] >= count || dup) {
            g_create_error = "device ordinal out of range (or listed twice)";
            return RT_ERR_INVALID_ARGUMENT;
        }
    }
    rt_ctx* c = nullptr;
    try {
        c = make_ctx(devs[0]);
        if (n_devices > 1) rt_multi_attach(c, devs, make_ctx);
        RT_CUDA(cudaSetDevice(devs[0]));
    } catch (const RtError& err) {
        g_create_error = err.msg;
        if (c) {
            for (rt_ctx* k : c->kids) destroy_one(k);
            destroy_one(c);
        }
        cudaGetLastError();
        return err.code;
    }
    *out = c;
    return RT_OK;
}

int rt_create(rt_ctx** out, int device) { return rt_create_multi(out, &device, 1); }

int rt_device_count(rt_ctx* ctx, int32_t* n) {
    if (!ctx || !n) return RT_ERR_INVALID_ARGUMENT;
    *n = 1 + (int32_t)ctx->kids.size();
    return RT_OK;
}

int rt_destroy(rt_ctx* ctx) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    rt_multi_pool_stop(ctx);
    for (rt_ctx* k : ctx->kids) {
        cudaSetDevice(k->device);
        cudaDeviceSynchronize();
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    // shared frames first: they are mapped on every device
    for (auto& s : ctx->slots) rt_frame_release(ctx, s.frame);
    for (rt_ctx* k : ctx->kids) destroy_one(k);
    destroy_one(ctx);
    return RT_OK;
}

int rt_host_alloc(void** out, uint64_t bytes) {
    if (!out || bytes == 0) return RT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        g_create_error = std::string("rt_host_alloc: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? RT_ERR_OUT_OF_MEMORY : RT_ERR_NO_DEVICE;
    }
    *out = p;
    return RT_OK;
}

int rt_host_free(void* p) {
    if (!p) return RT_OK;
    cudaError_t e = cudaFreeHost(p);
    cudaGetLastError();
    return e == cudaSuccess ? RT_OK : RT_ERR_CUDA;
}

const char* rt_last_error(rt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rt_set_stream(rt_ctx* ctx, void* cuda_stream) {
    return guarded(ctx, [&] { ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream; });
}

int rt_scene_set_triangles(rt_ctx* ctx, const float* v, const uint32_t* material_id, const float* vertex_rgb,
                           const uint32_t* object_id, uint32_t n) {
    return guarded(ctx, [&] {
        need(n == 0 || (v && material_id), "rt_scene_set_triangles: v and material_id are required");
        need(n < (1u << 28), "rt_scene_set_triangles: at most 2^28 - 1 triangles");
        for (rt_ctx* c : all_ranks(ctx)) {
            c->h_tri_v.assign(v, v + 9 * (size_t)n);
            c->host_vertices_stale = false;
            c->h_tri_mat.assign(material_id, material_id + n);
            c->h_tri_obj.resize(n);
            for (uint32_t i = 0; i < n; i++) c->h_tri_obj[i] = object_id ? object_id[i] : i;
            if (vertex_rgb) c->h_tri_rgb.assign(vertex_rgb, vertex_rgb + 9 * (size_t)n);
            else c->h_tri_rgb.clear();
            c->committed = false;
        }
    });
}

static void fill_ids(std::vector<AnalyticPrim>& dst, uint32_t kind, const uint32_t* material_id, const uint32_t* object_id,
                     uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        dst[i].kind = kind;
        dst[i].material = material_id[i];
        dst[i].object_id = object_id ? object_id[i] : 0xffffffffu;   // resolved at commit
        dst[i].tri_index = 0;
    }
}

int rt_scene_set_spheres(rt_ctx* ctx, const float* p, const uint32_t* material_id, const uint32_t* object_id, uint32_t n) {
    return guarded(ctx, [&] {
        need(n == 0 || (p && material_id), "rt_scene_set_spheres: packed and material_id are required");
        ctx->h_spheres.assign(n, AnalyticPrim{});
        fill_ids(ctx->h_spheres, RT_KIND_SPHERE, material_id, object_id, n);
        for (uint32_t i = 0; i < n; i++) {
            const float* q = p + 4 * (size_t)i;
            ctx->h_spheres[i].a = make_float4(q[0], q[1], q[2], q[3]);
        }
        ctx->committed = false;
        for (rt_ctx* k : ctx->kids) { k->h_spheres = ctx->h_spheres; k->committed = false; }
    });
}

int rt_scene_set_planes(rt_ctx* ctx, const float* p, const uint32_t* material_id, const uint32_t* object_id, uint32_t n) {
    return guarded(ctx, [&] {
        need(n == 0 || (p && material_id), "rt_scene_set_planes: packed and material_id are required");
        ctx->h_planes.assign(n, AnalyticPrim{});
        fill_ids(ctx->h_planes, RT_KIND_PLANE, material_id, object_id, n);
        for (uint32_t i = 0; i < n; i++) {
            const float* q = p + 12 * (size_t)i;
            ctx->h_planes[i].a = make_float4(q[0], q[1], q[2], 0);
            ctx->h_planes[i].b = make_float4(q[3], q[4], q[5], 0);
            ctx->h_planes[i].c = make_float4(q[6], q[7], q[8], 0);
            ctx->h_planes[i].d = make_float4(q[9], q[10], q[11], 0);
        }
        ctx->committed = false;
        for (rt_ctx* k : ctx->kids) { k->h_planes = ctx->h_planes; k->committed = false; }
    });
}

int rt_scene_set_cylinders(rt_ctx* ctx, const float* p, const uint32_t* material_id, const uint32_t* object_id, uint32_t n) {
    return guarded(ctx, [&] {
        need(n == 0 || (p && material_id), "rt_scene_set_cylinders: packed and material_id are required");
        ctx->h_cylinders.assign(n, AnalyticPrim{});
        fill_ids(ctx->h_cylinders, RT_KIND_CYLINDER, material_id, object_id, n);
        for (uint32_t i = 0; i < n; i++) {
            const float* q = p + 7 * (size_t)i;
            ctx->h_cylinders[i].a = make_float4(q[0], q[1], q[2], q[3]);
            ctx->h_cylinders[i].b = make_float4(q[4], q[5], q[6], 0);
        }
        ctx->committed = false;
        for (rt_ctx* k : ctx->kids) { k->h_cylinders = ctx->h_cylinders; k->committed = false; }
    });
}

int rt_scene_set_materials(rt_ctx* ctx, const rt_material* m, uint32_t n) {
    return guarded(ctx, [&] {
        need(n > 0 && m, "rt_scene_set_materials: at least one material is required");
        for (rt_ctx* c : all_ranks(ctx)) {
            c->h_materials.assign(m, m + n);
            c->committed = false;
        }
    });
}

int rt_scene_set_lights(rt_ctx* ctx, const float* pos_rgb, uint32_t n) {
    return guarded(ctx, [&] {
        need(n == 0 || pos_rgb, "rt_scene_set_lights: pos_rgb is required");
        for (rt_ctx* c : all_ranks(ctx)) {
            c->h_lights.assign(pos_rgb, pos_rgb + 6 * (size_t)n);
            c->committed = false;
        }
    });
}

int rt_scene_set_environment(rt_ctx* ctx, const float ambient[3], const float background[3]) {
    return guarded(ctx, [&] {
        need(ambient && background, "rt_scene_set_environment: both colours are required");
        for (rt_ctx* c : all_ranks(ctx)) {
            for (int k = 0; k < 3; k++) { c->ambient[k] = ambient[k]; c->background[k] = background[k]; }
            if (c->committed) for (int k = 0; k < 3; k++) { c->scene.ambient[k] = ambient[k]; c->scene.background[k] = background[k]; }
        }
    });
}

int rt_scene_commit(rt_ctx* ctx, int mode) {
    return guarded(ctx, [&] {
        need(mode == RT_COMMIT_BUILD || mode == RT_COMMIT_REFIT, "rt_scene_commit: unknown mode");
        if (mode == RT_COMMIT_REFIT) {
            if (!ctx->committed) throw RtError{RT_ERR_NOT_COMMITTED, "rt_scene_commit(REFIT) before a BUILD commit"};
            for (rt_ctx* c : all_ranks(ctx)) {      // only enqueues kernels: no helper threads needed
                RT_CUDA(cudaSetDevice(c->device));
                rt_build_bvh(c, true);
            }
            RT_CUDA(cudaSetDevice(ctx->device));
            return;
        }
        need(!ctx->h_materials.empty(), "rt_scene_commit: no materials set");
        if (ctx->host_vertices_stale && ctx->n_tri && ctx->h_tri_v.size() == 9 * (size_t)ctx->n_tri &&
            ctx->d_tri_v.cap >= 9 * (size_t)ctx->n_tri) {
            // vertices were last written on the device (rt_scene_update_vertices_device): a rebuild starts from them
            RT_CUDA(cudaMemcpyAsync(ctx->h_tri_v.data(), ctx->d_tri_v.p, ctx->h_tri_v.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            RT_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        ctx->host_vertices_stale = false;
        uint32_t nm = (uint32_t)ctx->h_materials.size();
        for (uint32_t m : ctx->h_tri_mat) need(m < nm, "rt_scene_commit: triangle material id out of range");
        bool any_bary = false;
        for (const auto& m : ctx->h_materials) any_bary |= (m.flags & RT_MATERIAL_BARYCENTRIC) != 0;
        if (any_bary)
            for (uint32_t m : ctx->h_tri_mat)
                if (ctx->h_materials[m].flags & RT_MATERIAL_BARYCENTRIC)
                    need(!ctx->h_tri_rgb.empty(), "rt_scene_commit: barycentric material without vertex_rgb");
        // default object ids of the analytic kinds: after the triangles, in the order spheres, planes, cylinders
        uint32_t next = (uint32_t)(ctx->h_tri_v.size() / 9);
        for (auto* list : {&ctx->h_spheres, &ctx->h_planes, &ctx->h_cylinders})
            for (auto& p : *list) {
                need(p.material < nm, "rt_scene_commit: analytic material id out of range");
                need(!(ctx->h_materials[p.material].flags & RT_MATERIAL_BARYCENTRIC),
                     "rt_scene_commit: barycentric materials are for triangles only");
                if (p.object_id == 0xffffffffu) p.object_id = next;
                next++;
            }
        for (rt_ctx* k : ctx->kids) {               // same staging on every rank (ids resolved, vertices current)
            k->h_tri_v = ctx->h_tri_v;
            k->host_vertices_stale = false;
            k->h_spheres = ctx->h_spheres; k->h_planes = ctx->h_planes; k->h_cylinders = ctx->h_cylinders;
        }
        if (ctx->kids.empty()) rt_build_bvh(ctx, false);
        else rt_multi_build(ctx, false);
        ctx->committed = true;
    });
}

int rt_scene_update_vertices(rt_ctx* ctx, const float* v, uint32_t n) {
    return guarded(ctx, [&] {
        if (!ctx->committed) throw RtError{RT_ERR_NOT_COMMITTED, "rt_scene_update_vertices before commit"};
        need(v && n == ctx->n_tri, "rt_scene_update_vertices: vertex count must equal the committed triangle count");
        for (rt_ctx* c : all_ranks(ctx)) {
            RT_CUDA(cudaSetDevice(c->device));
            c->h_tri_v.assign(v, v + 9 * (size_t)n);
            c->host_vertices_stale = false;
            RT_CUDA(cudaMemcpyAsync(c->d_tri_v.p, c->h_tri_v.data(), 9 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        }
        for (rt_ctx* c : all_ranks(ctx)) RT_CUDA(cudaStreamSynchronize(c->stream));
        RT_CUDA(cudaSetDevice(ctx->device));
    });
}

int rt_scene_update_vertices_device(rt_ctx* ctx, const float* v_dev, uint32_t n) {
    return guarded(ctx, [&] {
        if (!ctx->committed) throw RtError{RT_ERR_NOT_COMMITTED, "rt_scene_update_vertices_device before commit"};
        need(v_dev && n == ctx->n_tri, "rt_scene_update_vertices_device: vertex count must equal the committed triangle count");
        // device to device on the context's stream, nothing waits: the caller's deformation kernel must have been
        // enqueued on the same stream (rt_set_stream) or be complete
        RT_CUDA(cudaMemcpyAsync(ctx->d_tri_v.p, v_dev, 9 * (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->host_vertices_stale = true;
        if (!ctx->kids.empty()) {
            // the other devices take their copy from rank 0's over NVLink, after rank 0's copy has landed
            RT_CUDA(cudaEventRecord(ctx->mg_ev[0], ctx->stream));
            for (rt_ctx* k : ctx->kids) {
                RT_CUDA(cudaSetDevice(k->device));
                RT_CUDA(cudaStreamWaitEvent(k->stream, ctx->mg_ev[0], 0));
                RT_CUDA(cudaMemcpyPeerAsync(k->d_tri_v.p, k->device, ctx->d_tri_v.p, ctx->device, 9 * (size_t)n * sizeof(float), k->stream));
                k->host_vertices_stale = true;
            }
            RT_CUDA(cudaSetDevice(ctx->device));
        }
    });
}

int rt_scene_build_stats(rt_ctx* ctx, rt_build_stats* out) {
    return guarded(ctx, [&] {
        need(out != nullptr, "rt_scene_build_stats: out is NULL");
        if (ctx->refit_pending) {
            RT_CUDA(cudaEventSynchronize(ctx->ev[9]));
            RT_CUDA(cudaEventElapsedTime(&ctx->build_stats.ms_refit, ctx->ev[8], ctx->ev[9]));
            ctx->refit_pending = false;
        }
        *out = ctx->build_stats;
    });
}

static void check_render_args(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p) {
    if (!ctx->committed) throw RtError{RT_ERR_NOT_COMMITTED, "render before rt_scene_commit"};
    need(cam && p, "render: camera and params are required");
    need(cam->width > 0 && cam->height > 0 && (int64_t)cam->width * cam->height < (1ll << 31), "render: bad frame size");
}

int rt_render_device(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, void* rgb_out_dev,
                     const rt_aux_out* aux_dev, rt_frame_stats* stats) {
    return guarded(ctx, [&] {
        check_render_args(ctx, cam, p);
        if (ctx->kids.empty()) {
            rt_render_frame(ctx, cam, p, rgb_out_dev, aux_dev, stats);
            return;
        }
        // several devices: rgb_out_dev / aux_dev live on rank 0's device; every rank stores its tiles there
        need(rgb_out_dev != nullptr, "rt_render_device on a multi-device context needs an output buffer");
        rt_multi_enqueue_frame(ctx, cam, p, rgb_out_dev, aux_dev);
        if (stats) rt_multi_collect(ctx, stats);
    });
}

// Device-side aux buffers of rt_render (rank 0), cleared to "miss".
static rt_aux_out prepare_aux(rt_ctx* ctx, const rt_aux_out* aux, size_t npix) {
    rt_aux_out aux_dev{nullptr, nullptr};
    if (aux && (aux->prim_id || aux->t)) {
        ctx->d_aux_prim.reserve(npix);
        ctx->d_aux_t.reserve(npix);
        aux_dev.prim_id = ctx->d_aux_prim.p;
        aux_dev.t = ctx->d_aux_t.p;
        // pixels of tiles this rank does not own stay "miss"
        RT_CUDA(cudaMemsetAsync(ctx->d_aux_prim.p, 0xff, npix * sizeof(int32_t), ctx->stream));
        RT_CUDA(cudaMemsetAsync(ctx->d_aux_t.p, 0, npix * sizeof(float), ctx->stream));
    }
    return aux_dev;
}

static void ensure_slot(rt_ctx* ctx, int slot, size_t bytes) {
    rt_ctx::FrameSlot& s = ctx->slots[slot];
    rt_frame_reserve(ctx, s.frame, bytes);
    if (!s.ready) {
        RT_CUDA(cudaSetDevice(ctx->device));
        RT_CUDA(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
        std::vector<rt_ctx*> ranks = all_ranks(ctx);
        for (size_t r = 0; r < ranks.size(); r++) {
            RT_CUDA(cudaSetDevice(ranks[r]->device));
            RT_CUDA(cudaEventCreateWithFlags(&s.done[r], cudaEventDisableTiming));
        }
        RT_CUDA(cudaSetDevice(ctx->device));
        RT_CUDA(cudaMallocHost((void**)&s.h_sticky, RT_MAX_DEVICES * sizeof(uint32_t)));
        memset(s.h_sticky, 0, RT_MAX_DEVICES * sizeof(uint32_t));
    }
}

// Waits for the host copy of `slot` and raises what the kernels of that frame reported.
static void wait_slot(rt_ctx* ctx, int slot) {
    rt_ctx::FrameSlot& s = ctx->slots[slot];
    if (!s.in_flight) return;
    s.in_flight = false;
    std::vector<rt_ctx*> ranks = all_ranks(ctx);
    RtError first{RT_OK, ""};
    for (size_t r = 0; r < ranks.size(); r++) {
        RT_CUDA(cudaEventSynchronize(s.done[r]));
        uint32_t fl = s.h_sticky[r];
        if (fl) {
            s.h_sticky[r] = 0;
            RT_CUDA(cudaSetDevice(ranks[r]->device));
            RT_CUDA(cudaMemsetAsync(ranks[r]->d_sticky.p, 0, sizeof(uint32_t), ranks[r]->stream));
            if (first.code == RT_OK) first = rt_sticky_error(fl);
        }
    }
    RT_CUDA(cudaSetDevice(ctx->device));
    if (first.code != RT_OK) throw first;
}

// Frame `slot`: render into the slot's device frame, then copy it to the host on the copy streams.

static double host_now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void enqueue_slot(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, uint8_t* rgb_out, int slot,
                         const rt_aux_out* aux_dev) {
    static const bool trace = getenv("RT_TRACE_HOST") != nullptr;
    double t0 = trace ? host_now_ms() : 0.0;
    size_t bytes = (size_t)cam->width * cam->height * 3;
    ensure_slot(ctx, slot, bytes);
    if (slot == 0 && ctx->pipelined) ensure_slot(ctx, 1, bytes);   // (allocating a shared frame takes milliseconds)
    rt_ctx::FrameSlot& s = ctx->slots[slot];
    if (s.in_flight) wait_slot(ctx, slot);           // the slot's previous frame must have left its device frame
    double t1 = trace ? host_now_ms() : 0.0;
    if (ctx->kids.empty()) {
        rt_render_params q = *p;
        q.flags &= ~(uint32_t)RT_FLAG_PACKED_TILES;
        q.world_size = 1; q.rank = 0;
        rt_render_frame(ctx, cam, &q, s.frame.va, aux_dev, nullptr);
    } else {
        rt_multi_enqueue_frame(ctx, cam, p, s.frame.va, aux_dev);
    }
    double t2 = trace ? host_now_ms() : 0.0;
    RT_CUDA(cudaEventRecord(s.ready, ctx->stream));
    rt_frame_download_async(ctx, s.frame, rgb_out, bytes, s.ready, s.done, s.h_sticky);
    s.in_flight = true;
    if (trace) fprintf(stderr, "[rt host] enqueue_slot %d: slot wait %.3f ms, frame enqueue %.3f ms, download enqueue %.3f ms\n", slot,
                       t1 - t0, t2 - t1, host_now_ms() - t2);
}

int rt_render(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, uint8_t* rgb_out, const rt_aux_out* aux,
              rt_frame_stats* stats) {
    return guarded(ctx, [&] {
        check_render_args(ctx, cam, p);
        need(rgb_out != nullptr, "rt_render: rgb_out is NULL");
        size_t npix = (size_t)cam->width * cam->height;
        if (!ctx->kids.empty()) {
            // several devices: whole frames only (the tile split is the library's business here)
            need(!(p->flags & RT_FLAG_PACKED_TILES) && p->world_size <= 1, "rt_render on a multi-device context renders whole frames");
            rt_aux_out aux_dev = prepare_aux(ctx, aux, npix);
            if (aux_dev.prim_id) {
                // the other ranks write their pixels' aux values into rank 0's buffers: after the clear above
                RT_CUDA(cudaEventRecord(ctx->mg_ev[0], ctx->stream));
                for (rt_ctx* k : ctx->kids) {
                    RT_CUDA(cudaSetDevice(k->device));
                    RT_CUDA(cudaStreamWaitEvent(k->stream, ctx->mg_ev[0], 0));
                }
                RT_CUDA(cudaSetDevice(ctx->device));
            }
            enqueue_slot(ctx, cam, p, rgb_out, 0, aux_dev.prim_id ? &aux_dev : nullptr);
            if (aux && aux->prim_id)
                RT_CUDA(cudaMemcpyAsync(aux->prim_id, ctx->d_aux_prim.p, npix * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            if (aux && aux->t)
                RT_CUDA(cudaMemcpyAsync(aux->t, ctx->d_aux_t.p, npix * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            RtError deferred{RT_OK, ""};
            try { rt_multi_collect(ctx, stats); } catch (const RtError& e) { deferred = e; }
            try { wait_slot(ctx, 0); } catch (const RtError& e) { if (deferred.code == RT_OK) deferred = e; }
            if (deferred.code != RT_OK) throw deferred;
            return;
        }
        bool packed = (p->flags & RT_FLAG_PACKED_TILES) != 0;
        size_t bytes = npix * 3;
        if (packed) {
            uint32_t total, owned, tb;
            rt_tile_layout(cam->width, cam->height, p->tile_w, p->tile_h, p->world_size > 1 ? p->rank : 0,
                           p->world_size > 1 ? p->world_size : 1, &total, &owned, &tb);
            bytes = (size_t)owned * tb;
        }
        ctx->d_rgb.reserve(bytes ? bytes : 1);
        rt_aux_out aux_dev = prepare_aux(ctx, aux, npix);
        if (!packed && p->world_size > 1) RT_CUDA(cudaMemsetAsync(ctx->d_rgb.p, 0, bytes, ctx->stream));
        RtError deferred{RT_OK, ""};
        try {
            rt_render_frame(ctx, cam, p, ctx->d_rgb.p, (aux_dev.prim_id ? &aux_dev : nullptr), stats);
        } catch (const RtError& e) {
            if (e.code != RT_ERR_QUEUE_OVERFLOW) throw;
            deferred = e;
        }
        RT_CUDA(cudaMemcpyAsync(rgb_out, ctx->d_rgb.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (aux && aux->prim_id)
            RT_CUDA(cudaMemcpyAsync(aux->prim_id, ctx->d_aux_prim.p, npix * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (aux && aux->t)
            RT_CUDA(cudaMemcpyAsync(aux->t, ctx->d_aux_t.p, npix * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        // the copies are in flight: wait for them and surface errors the kernels raised meanwhile (with stats == NULL
        // rt_render_frame did not synchronise, so the sticky error word has not been looked at yet)
        try {
            rt_sync_and_check(ctx);
        } catch (const RtError& e) {
            if (deferred.code == RT_OK) deferred = e;
        }
        if (deferred.code != RT_OK) throw deferred;
    });
}

int rt_render_enqueue(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, uint8_t* rgb_out, int32_t slot) {
    return guarded(ctx, [&] {
        check_render_args(ctx, cam, p);
        need(rgb_out != nullptr && (slot == 0 || slot == 1), "rt_render_enqueue: rgb_out is NULL or slot is not 0/1");
        need(!(p->flags & RT_FLAG_PACKED_TILES) && p->world_size <= 1, "rt_render_enqueue renders whole frames");
        ctx->pipelined = true;                         // both frame slots are set up at the first call
        // an asynchronous copy needs a page-locked destination: lock the caller's buffer the first time it is seen.
        // (The pointer query takes a driver-wide lock and was measured at 0.1 - 15 ms while frames are in flight, so
        // every destination is looked at once.)
        bool known = false;
        for (void* q : ctx->host_checked) known |= q == (void*)rgb_out;
        if (!known) {
            cudaPointerAttributes at;
            cudaError_t e = cudaPointerGetAttributes(&at, rgb_out);
            if (e != cudaSuccess || at.type == cudaMemoryTypeUnregistered) {
                cudaGetLastError();
                size_t bytes = (size_t)cam->width * cam->height * 3;
                if (cudaHostRegister(rgb_out, bytes, cudaHostRegisterPortable) == cudaSuccess) ctx->host_registered.push_back(rgb_out);
                else cudaGetLastError();              // stays pageable: the copy is then staged by the driver
            }
            if (ctx->host_checked.size() >= 64) ctx->host_checked.clear();
            ctx->host_checked.push_back(rgb_out);
        }
        enqueue_slot(ctx, cam, p, rgb_out, slot, nullptr);
    });
}

int rt_render_wait(rt_ctx* ctx, int32_t slot) {
    return guarded(ctx, [&] {
        need(slot == 0 || slot == 1, "rt_render_wait: slot is not 0/1");
        wait_slot(ctx, slot);
    });
}

int rt_synchronize(rt_ctx* ctx) {
    return guarded(ctx, [&] {
        RtError first{RT_OK, ""};
        for (rt_ctx* c : all_ranks(ctx)) {
            try {
                RT_CUDA(cudaSetDevice(c->device));
                rt_sync_and_check(c);
            } catch (const RtError& e) {
                if (first.code == RT_OK) first = e;
            }
        }
        for (int s = 0; s < 2; s++) {
            try { wait_slot(ctx, s); } catch (const RtError& e) { if (first.code == RT_OK) first = e; }
        }
        RT_CUDA(cudaSetDevice(ctx->device));
        if (first.code != RT_OK) throw first;
    });
}

int rt_shared_buffer_create(rt_ctx* ctx, uint64_t bytes, void** dev_ptr, unsigned char handle[64]) {
    return guarded(ctx, [&] {
        need(dev_ptr && handle && bytes > 0, "rt_shared_buffer_create: bad arguments");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        void* p = nullptr;
        RT_CUDA(cudaMalloc(&p, bytes));
        RT_CUDA(cudaMemset(p, 0, bytes));
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) { cudaFree(p); throw RtError{RT_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)}; }
        memcpy(handle, &h, 64);
        ctx->ipc_created.push_back(p);
        *dev_ptr = p;
    });
}

int rt_shared_buffer_open(rt_ctx* ctx, const unsigned char handle[64], void** dev_ptr) {
    return guarded(ctx, [&] {
        need(dev_ptr && handle, "rt_shared_buffer_open: bad arguments");
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, 64);
        void* p = nullptr;
        RT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_opened.push_back(p);
        *dev_ptr = p;
    });
}

int rt_peer_sync(rt_ctx* ctx, void* sync_buf, int32_t rank, int32_t world_size, uint32_t frame_index, int32_t phase) {
    return guarded(ctx, [&] {
        need(sync_buf && world_size >= 1 && rank >= 0 && rank < world_size && (phase == 0 || phase == 1), "rt_peer_sync: bad arguments");
        rt_peer_sync_enqueue(ctx, sync_buf, rank, world_size, frame_index, phase);
    });
}

int rt_peer_barrier(rt_ctx* ctx, void* sync_buf, int32_t world_size, uint32_t epoch) {
    return guarded(ctx, [&] {
        need(sync_buf && world_size >= 1, "rt_peer_barrier: bad arguments");
        rt_peer_barrier_enqueue(ctx, sync_buf, world_size, epoch);
    });
}

int rt_render_push(rt_ctx* ctx, const rt_camera* cam, const rt_render_params* p, void* packed_dev, void* frame_dev,
                   void* sync_buf, uint32_t frame_index) {
    return guarded(ctx, [&] {
        check_render_args(ctx, cam, p);
        need(packed_dev && frame_dev && sync_buf, "rt_render_push: NULL buffer");
        need((p->flags & RT_FLAG_PACKED_TILES) != 0 && p->world_size >= 1, "rt_render_push: needs RT_FLAG_PACKED_TILES and a world size");
        rt_push_frame(ctx, cam, p, packed_dev, frame_dev, sync_buf, frame_index, nullptr);
    });
}

int rt_download(rt_ctx* ctx, const void* dev_ptr, void* host, uint64_t bytes) {
    return guarded(ctx, [&] {
        need(dev_ptr && host, "rt_download: NULL buffer");
        RT_CUDA(cudaMemcpyAsync(host, dev_ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        RT_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int rt_tile_layout(int32_t width, int32_t height, int32_t tile_w, int32_t tile_h, int32_t rank, int32_t world_size,
                   uint32_t* tiles_total, uint32_t* tiles_owned, uint32_t* tile_bytes) {
    if (tile_w <= 0) tile_w = 64;
    if (tile_h <= 0) tile_h = 32;
    if (width <= 0 || height <= 0 || world_size < 1 || rank < 0 || rank >= world_size) return RT_ERR_INVALID_ARGUMENT;
    uint32_t tx = (uint32_t)((width + tile_w - 1) / tile_w), ty = (uint32_t)((height + tile_h - 1) / tile_h);
    uint32_t total = tx * ty;
    if (tiles_total) *tiles_total = total;
    if (tiles_owned) *tiles_owned = total > (uint32_t)rank ? (total - (uint32_t)rank + (uint32_t)world_size - 1) / (uint32_t)world_size : 0u;
    if (tile_bytes) *tile_bytes = (uint32_t)(tile_w * tile_h * 3);
    return RT_OK;
}

int rt_assemble_tiles(rt_ctx* ctx, const void* packed_dev, int32_t src_rank, int32_t world_size, int32_t width,
                      int32_t height, int32_t tile_w, int32_t tile_h, void* frame_dev) {
    return guarded(ctx, [&] {
        need(packed_dev && frame_dev, "rt_assemble_tiles: NULL buffer");
        if (tile_w <= 0) tile_w = 64;
        if (tile_h <= 0) tile_h = 32;
        need(world_size >= 1 && src_rank >= 0 && src_rank < world_size && width > 0 && height > 0, "rt_assemble_tiles: bad layout");
        rt_assemble(ctx, packed_dev, src_rank, world_size, width, height, tile_w, tile_h, frame_dev);
    });
}

int rt_trace_rays(rt_ctx* ctx, const float* rays, uint32_t n, uint32_t flags, int32_t* prim_id, float* t) {
    return guarded(ctx, [&] {
        if (!ctx->committed) throw RtError{RT_ERR_NOT_COMMITTED, "rt_trace_rays before rt_scene_commit"};
        need(n == 0 || rays, "rt_trace_rays: rays is NULL");
        rt_query_rays(ctx, rays, n, 0, flags, false, prim_id, t, nullptr);
    });
}

int rt_shade_rays(rt_ctx* ctx, const float* rays, uint32_t n, int32_t max_depth, uint32_t flags, float* rgb_out) {
    return guarded(ctx, [&] {
        if (!ctx->committed) throw RtError{RT_ERR_NOT_COMMITTED, "rt_shade_rays before rt_scene_commit"};
        need(n == 0 || (rays && rgb_out), "rt_shade_rays: NULL buffer");
        rt_query_rays(ctx, rays, n, max_depth, flags, true, nullptr, nullptr, rgb_out);
    });
}

int rt_bvh_download(rt_ctx* ctx, float* nodes, uint32_t* tri_order, uint64_t* keys, uint32_t* n_nodes,
                    uint32_t* n_bvh_triangles) {
    return guarded(ctx, [&] {
        if (!ctx->committed) throw RtError{RT_ERR_NOT_COMMITTED, "rt_bvh_download before rt_scene_commit"};
        uint32_t nn = (uint32_t)ctx->scene.n_nodes, nb = ctx->n_bvh;
        if (n_nodes) *n_nodes = nn;
        if (n_bvh_triangles) *n_bvh_triangles = nb;
        cudaStream_t st = ctx->stream;
        if (nodes && nn) RT_CUDA(cudaMemcpyAsync(nodes, ctx->d_nodes.p, (size_t)nn * 16 * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (tri_order && nb) RT_CUDA(cudaMemcpyAsync(tri_order, ctx->d_vals[ctx->sorted_buf].p, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (keys && nb) RT_CUDA(cudaMemcpyAsync(keys, ctx->d_keys[ctx->sorted_buf].p, nb * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        RT_CUDA(cudaStreamSynchronize(st));
    });
}

int rt_microbench(rt_ctx* ctx, rt_microbench_result* out) {
    return guarded(ctx, [&] {
        need(out != nullptr, "rt_microbench: out is NULL");
        rt_run_microbench(ctx, out);
    });
}

int rt_debug_frame_launches(rt_ctx* ctx, uint64_t* n_kernels) {
    return guarded(ctx, [&] {
        need(n_kernels, "rt_debug_frame_launches: NULL argument");
        *n_kernels = ctx->launch_total;
    });
}

int rt_debug_warp_times(rt_ctx* ctx, uint64_t* out, uint32_t* n_warps) {
    return guarded(ctx, [&] {
        need(out && n_warps, "rt_debug_warp_times: NULL argument");
        uint32_t have = (uint32_t)(ctx->d_warp_times.cap / 2);
        uint32_t n = have < *n_warps ? have : *n_warps;
        if (n) {
            RT_CUDA(cudaMemcpyAsync(out, ctx->d_warp_times.p, (size_t)n * 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
            RT_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        *n_warps = n;
    });
}

int rt_debug_sort_pairs(rt_ctx* ctx, uint64_t* keys, uint32_t* values, uint32_t n) {
    return guarded(ctx, [&] {
        need(n == 0 || (keys && values), "rt_debug_sort_pairs: NULL buffer");
        if (n == 0) return;
        // the scene's sort buffers are reused: a later commit rebuilds them, but an already
        // committed hierarchy keeps pointing at d_vals[sorted_buf], so refuse in that state
        need(!ctx->committed, "rt_debug_sort_pairs: use a context without a committed scene");
        cudaStream_t st = ctx->stream;
        ctx->d_keys[0].reserve(n);
        ctx->d_vals[0].reserve(n);
        RT_CUDA(cudaMemcpyAsync(ctx->d_keys[0].p, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        RT_CUDA(cudaMemcpyAsync(ctx->d_vals[0].p, values, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        rt_sort_pairs_device(ctx, n);
        RT_CUDA(cudaMemcpyAsync(keys, ctx->d_keys[ctx->sorted_buf].p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        RT_CUDA(cudaMemcpyAsync(values, ctx->d_vals[ctx->sorted_buf].p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        RT_CUDA(cudaStreamSynchronize(st));
    });
}

}  // extern "C"
