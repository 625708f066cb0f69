// multi_device.cu — one rt_ctx driving several GPUs from one process (rt_create_multi), and the
// pipelined host-frame path (rt_render_enqueue / rt_render_wait).
//
// Replaces nothing in the reference (a single-threaded CPU renderer, Serial/renderengine.cpp:10-26): it is
// how `RenderEngine::render()` of the drop-in uses every GPU of the node.  SURVEY 8(e): the scene is
// replicated, tile t of the frame belongs to rank t % n, every pixel is computed by exactly one rank with the
// same code and data, so the assembled frame is bit-identical for any n.
//
// Where the frame lives.  For n > 1 the frame is ONE virtual address range (CUDA virtual memory management)
// whose 2 MiB granules are physically placed round robin on the n GPUs and mapped read/write on all of them
// over NVLink.  The render kernels are unchanged — they store finished 8x4 pixel blocks to frame + offset and
// the store lands on whichever GPU backs that granule — but the copy to the host now runs on n PCIe links at
// once: each GPU copies the granules it owns straight into the caller's page-locked buffer.  If the VMM calls
// are not available the frame falls back to a plain allocation on rank 0 (one PCIe link).
//
// Completion: the flag handshake of render.cu in rank 0's memory (no collective, no host round trip between
// the ranks); rank 0's stream then records one event that every rank's copy stream waits for.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <condition_variable>
#include <cstring>
#include <ctime>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>

#include "rt_context.h"

namespace {

// driver entry points, resolved at run time through the runtime (no link-time dependency on libcuda)
struct Vmm {
    PFN_cuMemGetAllocationGranularity granularity = nullptr;
    PFN_cuMemAddressReserve reserve = nullptr;
    PFN_cuMemCreate create = nullptr;
    PFN_cuMemMap map = nullptr;
    PFN_cuMemSetAccess set_access = nullptr;
    PFN_cuMemUnmap unmap = nullptr;
    PFN_cuMemRelease release = nullptr;
    PFN_cuMemAddressFree address_free = nullptr;
    bool ok = false;
    Vmm() {
        struct { const char* name; void** fn; } want[] = {
            {"cuMemGetAllocationGranularity", (void**)&granularity}, {"cuMemAddressReserve", (void**)&reserve},
            {"cuMemCreate", (void**)&create}, {"cuMemMap", (void**)&map}, {"cuMemSetAccess", (void**)&set_access},
            {"cuMemUnmap", (void**)&unmap}, {"cuMemRelease", (void**)&release}, {"cuMemAddressFree", (void**)&address_free}};
        ok = getenv("RT_NO_VMM") == nullptr;
        for (auto& w : want) {
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(w.name, w.fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*w.fn)
                ok = false;
        }
        cudaGetLastError();
    }
};
Vmm& vmm() {
    static Vmm v;
    return v;
}

std::vector<rt_ctx*> ranks_of(rt_ctx* c) {
    std::vector<rt_ctx*> v{c};
    v.insert(v.end(), c->kids.begin(), c->kids.end());
    return v;
}

}  // namespace

// ---- one enqueue thread per extra device --------------------------------------------------------------------
// A frame of an n-device context is ~8 runtime calls per rank for the kernels and ~5 + its granules for the copy to the
// host: 0.117 ms of the caller's time per frame at 8 GPUs when one thread issues them all.  Rank r's calls are independent
// of the other ranks' (own device, own streams), so with RT_MULTI_THREADS rank 0's are issued by the caller and every other
// rank's by a thread of its own; the caller returns when all of them have been enqueued (0.062 ms at 8 GPUs).  Measured on
// 8 x B200 (profiles/r2_tuning.md section 18) the pipelined frame rate does not change — it is bound by the frame's way into
// host memory, not by the enqueue — so the default is 0 (one thread); 1 = threads from four devices on, 2 = always.
struct RtRankPool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable go, done;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
    const std::function<void(int)>* job = nullptr;
    std::vector<RtError> errs;
};

static void rank_pool_worker(RtRankPool* p, int r, int device) {
    cudaSetDevice(device);
    uint64_t seen = 0;
    for (;;) {
        const std::function<void(int)>* job;
        {
            std::unique_lock<std::mutex> lk(p->m);
            p->go.wait(lk, [&] { return p->stop || p->gen != seen; });
            if (p->stop) return;
            seen = p->gen;
            job = p->job;
        }
        RtError err{RT_OK, ""};
        try {
            (*job)(r);
        } catch (const RtError& e) {
            err = e;
        } catch (const std::exception& e) {
            err = RtError{RT_ERR_CUDA, e.what()};
        } catch (...) {
            err = RtError{RT_ERR_CUDA, "unknown exception on a rank's enqueue thread"};
        }
        {
            std::lock_guard<std::mutex> lk(p->m);
            p->errs[r] = err;
            if (--p->pending == 0) p->done.notify_one();
        }
    }
}

// job(r) for every rank: rank 0 on the calling thread, the others on their threads (or all on the caller without a
// pool).  Returns when every job has returned; rethrows the lowest rank's error.
static void run_on_ranks(rt_ctx* c, const std::function<void(int)>& job) {
    const int n = 1 + (int)c->kids.size();
    // (the two wake-ups of a frame cost what two ranks' calls cost: worth it from four devices on; 2 = always)
    if (n == 1 || !c->mg_threads || (c->mg_threads == 1 && n < 4)) {
        for (int r = 0; r < n; r++) job(r);
        return;
    }
    if (!c->mg_pool) {
        RtRankPool* p = new RtRankPool;
        p->errs.assign(n, RtError{RT_OK, ""});
        for (int r = 1; r < n; r++) p->th.emplace_back(rank_pool_worker, p, r, c->kids[r - 1]->device);
        c->mg_pool = p;
    }
    RtRankPool* p = c->mg_pool;
    {
        std::lock_guard<std::mutex> lk(p->m);
        for (auto& e : p->errs) e = RtError{RT_OK, ""};
        p->job = &job;
        p->pending = n - 1;
        p->gen++;
    }
    p->go.notify_all();
    std::exception_ptr mine;                   // whatever rank 0's job throws: the others still refer to `job`, wait first
    try {
        job(0);
    } catch (...) {
        mine = std::current_exception();
    }
    {
        std::unique_lock<std::mutex> lk(p->m);
        p->done.wait(lk, [&] { return p->pending == 0; });
        p->job = nullptr;
    }
    if (mine) std::rethrow_exception(mine);
    for (auto& e : p->errs)
        if (e.code != RT_OK) throw e;
}

void rt_multi_pool_stop(rt_ctx* c) {
    RtRankPool* p = c->mg_pool;
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->stop = true;
    }
    p->go.notify_all();
    for (auto& t : p->th) t.join();
    delete p;
    c->mg_pool = nullptr;
}

// ---- the striped frame ------------------------------------------------------------------------------------
void rt_frame_release(rt_ctx* c, rt_ctx::SharedFrame& f) {
    if (f.vmm) {
        Vmm& v = vmm();
        if (f.va) {
            v.unmap((CUdeviceptr)f.va, f.granules * f.gran);
            v.address_free((CUdeviceptr)f.va, f.granules * f.gran);
        }
        for (unsigned long long h : f.handles) v.release((CUmemGenericAllocationHandle)h);
    } else if (f.va) {
        cudaSetDevice(c->device);
        cudaFree((void*)f.va);
    }
    f = rt_ctx::SharedFrame{};
}

// A frame of `bytes` every rank of c can store into.  Striped over the ranks' GPUs where VMM works.
void rt_frame_reserve(rt_ctx* c, rt_ctx::SharedFrame& f, size_t bytes) {
    if (f.va && f.bytes >= bytes) return;
    rt_frame_release(c, f);
    std::vector<rt_ctx*> ranks = ranks_of(c);
    const int n = (int)ranks.size();
    Vmm& v = vmm();
    if (n > 1 && v.ok) {
        CUmemAllocationProp prop;
        memset(&prop, 0, sizeof prop);
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = c->device;
        size_t gran = 0;
        bool good = v.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) == CUDA_SUCCESS && gran > 0;
        size_t granules = good ? (bytes + gran - 1) / gran : 0;
        CUdeviceptr va = 0;
        good = good && v.reserve(&va, granules * gran, 0, 0, 0) == CUDA_SUCCESS;
        std::vector<unsigned long long> handles;
        std::vector<int> owner;
        size_t mapped = 0;
        for (size_t g = 0; good && g < granules; g++) {
            int r = (int)(g % (size_t)n);
            prop.location.id = ranks[r]->device;
            CUmemGenericAllocationHandle h;
            if (v.create(&h, gran, &prop, 0) != CUDA_SUCCESS) { good = false; break; }
            handles.push_back((unsigned long long)h);
            owner.push_back(r);
            if (v.map(va + g * gran, gran, 0, h, 0) != CUDA_SUCCESS) { good = false; break; }
            mapped = g + 1;
        }
        if (good) {
            std::vector<CUmemAccessDesc> acc(n);
            for (int r = 0; r < n; r++) {
                acc[r].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
                acc[r].location.id = ranks[r]->device;
                acc[r].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            }
            good = v.set_access(va, granules * gran, acc.data(), (size_t)n) == CUDA_SUCCESS;
        }
        if (good) {
            f.va = (void*)va; f.bytes = bytes; f.gran = gran; f.granules = granules; f.vmm = true;
            f.handles = handles; f.owner = owner;
            RT_CUDA(cudaMemsetAsync(f.va, 0, granules * gran, c->stream));
            RT_CUDA(cudaStreamSynchronize(c->stream));
            return;
        }
        if (va) {
            if (mapped) v.unmap(va, mapped * gran);
            v.address_free(va, granules * gran);
        }
        for (unsigned long long h : handles) v.release((CUmemGenericAllocationHandle)h);
        cudaGetLastError();
    }
    // one device (or no VMM): a plain allocation on rank 0; peers reach it through cudaDeviceEnablePeerAccess
    RT_CUDA(cudaSetDevice(c->device));
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) throw RtError{RT_ERR_OUT_OF_MEMORY, std::string("frame allocation failed: ") + cudaGetErrorString(e)};
    RT_CUDA(cudaMemset(p, 0, bytes));
    f.va = p; f.bytes = bytes; f.gran = bytes; f.granules = 1; f.vmm = false;
    f.handles.clear();
    f.owner.assign(1, 0);
}

// ---- life cycle ---------------------------------------------------------------------------------------------
// Called by rt_create_multi after the rank-0 context exists: contexts for the other devices + peer access.
void rt_multi_attach(rt_ctx* c, const std::vector<int>& devices, rt_ctx* (*make_ctx)(int device)) {
    for (size_t i = 1; i < devices.size(); i++) {
        rt_ctx* k = make_ctx(devices[i]);
        k->parent = c;
        c->kids.push_back(k);
    }
    std::vector<rt_ctx*> ranks = ranks_of(c);
    // every pair, both directions: frames are striped over all GPUs, flags live on rank 0
    for (rt_ctx* a : ranks)
        for (rt_ctx* b : ranks) {
            if (a == b) continue;
            int can = 0;
            RT_CUDA(cudaDeviceCanAccessPeer(&can, a->device, b->device));
            if (!can) throw RtError{RT_ERR_CUDA, "rt_create_multi: the devices have no peer access to each other (NVLink/PCIe P2P)"};
            RT_CUDA(cudaSetDevice(a->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) RT_CUDA(e);
            cudaGetLastError();
        }
    RT_CUDA(cudaSetDevice(c->device));
    void* s = nullptr;
    RT_CUDA(cudaMalloc(&s, 1024));
    RT_CUDA(cudaMemset(s, 0, 1024));
    c->mg_sync = s;
}

// LBVH build on every device at once (the build synchronises with the host several times; one helper thread per
// extra device keeps the builds concurrent).
void rt_multi_build(rt_ctx* c, bool refit_only) {
    // every rank renders a share: the heaviest tiles will walk the 4-wide view (RT_WIDE_HEAVY), so build it with the
    // tree instead of inside the second frame
    if (!c->kids.empty() && c->wide_heavy >= 1) {
        c->want_nodes4 = true;
        for (rt_ctx* k : c->kids) k->want_nodes4 = true;
    }
    std::vector<RtError> errs(c->kids.size(), RtError{RT_OK, ""});
    std::vector<std::thread> th;
    for (size_t i = 0; i < c->kids.size(); i++) {
        rt_ctx* k = c->kids[i];
        th.emplace_back([k, refit_only, &errs, i] {
            try {
                RT_CUDA(cudaSetDevice(k->device));
                rt_build_bvh(k, refit_only);
                k->committed = true;
            } catch (const RtError& e) {
                errs[i] = e;
            }
        });
    }
    RtError mine{RT_OK, ""};
    try {
        RT_CUDA(cudaSetDevice(c->device));
        rt_build_bvh(c, refit_only);
    } catch (const RtError& e) {
        mine = e;
    }
    for (auto& t : th) t.join();
    cudaSetDevice(c->device);
    if (mine.code != RT_OK) throw mine;
    for (auto& e : errs)
        if (e.code != RT_OK) throw e;
}

// ---- one frame over all ranks ---------------------------------------------------------------------------------
// Enqueues the frame on every rank's stream: rank r renders the tiles t % n == r and stores them into frame_dev
// (handshake phases as in rt_render_push).  Nothing here waits for the GPU.
void rt_multi_enqueue_frame(rt_ctx* c, const rt_camera* cam, const rt_render_params* p, void* frame_dev,
                            const rt_aux_out* aux_dev) {
    std::vector<rt_ctx*> ranks = ranks_of(c);
    const int n = (int)ranks.size();
    rt_render_params q = *p;
    q.world_size = n;
    q.flags |= RT_FLAG_PACKED_TILES;
    q.steal_pool_div = 0;
    q.steal_cursor = nullptr;
    if (q.tile_w <= 0 || q.tile_h <= 0) {      // measured (profiles/r1_tuning.md 13): finer tiles from 4 GPUs on
        q.tile_w = n >= 4 ? 32 : 64;
        q.tile_h = n >= 4 ? 16 : 32;
    }
    const uint32_t k = c->mg_frame++;
    static const bool trace = getenv("RT_TRACE_HOST") != nullptr;
    run_on_ranks(c, [&](int r) {
        rt_ctx* x = ranks[r];
        timespec ts0;
        if (trace) clock_gettime(CLOCK_MONOTONIC, &ts0);
        RT_CUDA(cudaSetDevice(x->device));
        rt_render_params qr = q;
        qr.rank = r;
        uint32_t total, owned, tb;
        rt_tile_layout(cam->width, cam->height, qr.tile_w, qr.tile_h, r, n, &total, &owned, &tb);
        x->d_packed.reserve((size_t)(owned ? owned : 1) * tb);
        if (r == 0) RT_CUDA(cudaEventRecord(c->mg_ev[0], x->stream));
        rt_push_frame(x, cam, &qr, x->d_packed.p, frame_dev, c->mg_sync, k, aux_dev);
        if (r == 0) RT_CUDA(cudaEventRecord(c->mg_ev[1], x->stream));
        if (trace) {
            timespec ts1;
            clock_gettime(CLOCK_MONOTONIC, &ts1);
            fprintf(stderr, "[rt host] frame %u rank %d enqueued in %.3f ms\n", k, r,
                    (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);
        }
    });
    RT_CUDA(cudaSetDevice(c->device));
}

// Statistics of the frame just enqueued: waits for every rank, raises their errors, sums their counters.
void rt_multi_collect(rt_ctx* c, rt_frame_stats* stats) {
    std::vector<rt_ctx*> ranks = ranks_of(c);
    if (stats) memset(stats, 0, sizeof *stats);
    RtError first{RT_OK, ""};
    for (rt_ctx* x : ranks) {
        try {
            RT_CUDA(cudaSetDevice(x->device));
            RT_CUDA(cudaMemcpyAsync(x->h_frame, x->last_frame_ctr ? x->last_frame_ctr : x->d_frame.p, sizeof(FrameCounters), cudaMemcpyDeviceToHost, x->stream));
            rt_sync_and_check(x);
            if (stats) {
                stats->rays_primary += x->h_frame->rays_primary;
                stats->rays_shadow += x->h_frame->rays_shadow;
                stats->rays_secondary += x->h_frame->rays_secondary;
                stats->node_visits += x->h_frame->node_visits[0];
                stats->tri_tests += x->h_frame->tri_tests[0];
                stats->shadow_node_visits += x->h_frame->node_visits[1];
                stats->shadow_tri_tests += x->h_frame->tri_tests[1];
                stats->tiles += x->layout.n_tiles_owned;
            }
        } catch (const RtError& e) {
            if (first.code == RT_OK) first = e;
        }
    }
    cudaSetDevice(c->device);
    if (first.code != RT_OK) throw first;
    if (stats) {
        stats->waves = 1;
        RT_CUDA(cudaEventElapsedTime(&stats->ms_device, c->mg_ev[0], c->mg_ev[1]));
        stats->ms_trace = stats->ms_device;
    }
}

// ---- frame -> host ----------------------------------------------------------------------------------------------
// Copies `bytes` of the shared frame into host memory: every rank copies the granules its GPU backs on its own copy
// stream (its own PCIe link), after `ready` (recorded on rank 0's render stream once the frame is complete).
// done[r] is recorded on rank r's copy stream afterwards.
void rt_frame_download_async(rt_ctx* c, const rt_ctx::SharedFrame& f, uint8_t* host, size_t bytes, cudaEvent_t ready,
                             cudaEvent_t* done, uint32_t* h_sticky) {
    std::vector<rt_ctx*> ranks = ranks_of(c);
    const int n = (int)ranks.size();
    // (`ready` has been recorded by the caller: every rank's wait refers to that record)
    run_on_ranks(c, [&](int r) {
        rt_ctx* x = ranks[r];
        RT_CUDA(cudaSetDevice(x->device));
        RT_CUDA(cudaStreamWaitEvent(x->copy_stream, ready, 0));
        if (!f.vmm) {
            if (r == 0) RT_CUDA(cudaMemcpyAsync(host, f.va, bytes, cudaMemcpyDeviceToHost, x->copy_stream));
        } else {
            for (size_t g = (size_t)r; g < f.granules; g += (size_t)n) {
                size_t off = g * f.gran;
                if (off >= bytes) break;
                size_t len = bytes - off < f.gran ? bytes - off : f.gran;
                RT_CUDA(cudaMemcpyAsync(host + off, (const uint8_t*)f.va + off, len, cudaMemcpyDeviceToHost, x->copy_stream));
            }
        }
        // `ready` follows the arrival of every rank, so each rank's kernels of this frame have finished: its error
        // word is final
        RT_CUDA(cudaMemcpyAsync(h_sticky + r, x->d_sticky.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, x->copy_stream));
        RT_CUDA(cudaEventRecord(done[r], x->copy_stream));
    });
    RT_CUDA(cudaSetDevice(c->device));
}
