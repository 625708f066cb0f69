// render.cu — the wavefront pipeline: primary raygen + persistent-thread traversal (nearest hit and
// shadow any-hit), shading + bounce spawning, RGB8 resolve.
//
// Replaces the reference's frame loop RenderEngine::renderLoop -> trace -> World::shade_ray
// (/root/reference/Serial/renderengine.cpp:3-26, world.cpp:32-111) and its display conversion
// Camera::drawPixel (camera.cpp:46-52).  One "wave" is one generation of rays:
//   k_traverse<PRIMARY|QUEUE>  nearest hit.  Persistent warps; every LANE owns one ray and, when it
//             finishes, takes the next one from a global cursor (ballot + one atomic per refill), so
//             short and long rays do not hold each other up.  Wave 0 generates its rays in registers
//             from the pixel index (tile-ordered, 8x4 pixels per 32 consecutive indices); later waves
//             read 48-byte SoA ray records.  Hits are compacted into a hit queue with warp-aggregated
//             atomics; misses add throughput*background at once (world.cpp:110).
//             FUSE: the lane then walks its hit's shadow rays itself (one kernel, one tail per wave).
//             SHADE (primary wave, used where the frame sits in another GPU's memory): the warp also
//             shades its hits at the end of every 32-pixel batch and stores the finished 8x4 block
//             with 8-byte stores — trace, shade and delivery over NVLink in ONE kernel per frame.
//   k_traverse<SHADOW>  one any-hit query per (hit, light): world.cpp:44-51 (no tmax).  Same kernel
//             body, early exit on the first accepted hit, one occlusion byte out (> 8 lights, or
//             RT_FUSE_SHADOW=0).
//   k_shade   per hit: the local Phong-like term (world.cpp:126-137) from the occlusion bytes, then
//             the mirror / dielectric children (:77-107) appended to the next wave's ray queue with
//             one atomic per warp (shfl prefix sum).
//   k_paths   every bounce generation in one launch: a lane follows the whole subtree of one bounce ray.
//   k_resolve 32.32 fixed-point accumulators -> clamp -> (uint8)(255 c) (color.cpp:19-28,
//             camera.cpp:49-51) into the frame or into this rank's packed tile buffer.
// Pixel sums use integer atomics, so the frame is bit-identical for any tile split, queue order
// or GPU count.
#include <cstring>

#include "rt_context.h"
#include "rt_shade.h"

namespace {

constexpr int TRAV_TPB = 128;
// The traversal stack of a lane in the persistent kernels (RT_STACK_KIND, compile time):
//   0  per-thread array (local memory: STL/LDL through L1, up to 32 different lines per warp access)
//   1  the same with the top entry cached in a register (a pop's load leaves the dependent chain)
//   2  the first RT_SMEM_STACK entries in shared memory, laid out [entry][thread] so that a warp's access is free of
//      bank conflicts whatever each lane's depth, deeper entries in local memory
//   3  2 + the top entry in a register
#ifndef RT_STACK_KIND
#define RT_STACK_KIND 0
#endif
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 16
#endif
// min CTAs/SM in __launch_bounds__ measured: 10 (48 regs) = no bound (55 regs, 9 CTAs/SM); 12 and 16 spill and lose 15-50 %
constexpr int SHADE_TPB = 128;

// Scalars and pointers only (the arrays are declared next to it by RT_LANE_STACK_DECL): a struct that also held the
// array would be placed in local memory as a whole, `sp` included.
struct LaneStack {
    int* e;                                    // local-memory entries (kinds 2/3: the entries beyond RT_SMEM_STACK)
    int* sm;                                   // kinds 2/3: this thread's column of the CTA's shared block
    int sp;
    int top;                                   // kinds 1/3: entry sp-1; memory holds entries 0 .. sp-2
#if RT_STACK_KIND >= 2
    __device__ __forceinline__ int ld(int i) const { return i < RT_SMEM_STACK ? sm[i * TRAV_TPB] : e[i - RT_SMEM_STACK]; }
    __device__ __forceinline__ void st(int i, int v) {
        if (i < RT_SMEM_STACK) sm[i * TRAV_TPB] = v;
        else e[i - RT_SMEM_STACK] = v;
    }
#else
    __device__ __forceinline__ int ld(int i) const { return e[i]; }
    __device__ __forceinline__ void st(int i, int v) { e[i] = v; }
#endif
#if RT_STACK_KIND & 1
    __device__ __forceinline__ bool push(int v) {
        if (sp >= RT_STACK_SIZE) return false;
        if (sp > 0) st(sp - 1, top);
        top = v;
        sp++;
        return true;
    }
    __device__ __forceinline__ int pop() {
        int v = top;
        if (--sp > 0) top = ld(sp - 1);
        return v;
    }
#else
    __device__ __forceinline__ bool push(int v) {
        if (sp >= RT_STACK_SIZE) return false;
        st(sp++, v);
        return true;
    }
    __device__ __forceinline__ int pop() { return ld(--sp); }
#endif
    __device__ __forceinline__ void clear() { sp = 0; }
    __device__ __forceinline__ bool empty() const { return sp == 0; }
};
#if RT_STACK_KIND >= 2
#define RT_LANE_STACK_DECL(stk)                                   \
    __shared__ int s_lane_stack[RT_SMEM_STACK * TRAV_TPB];        \
    int lane_stack_mem[RT_STACK_SIZE - RT_SMEM_STACK];            \
    LaneStack stk;                                                \
    stk.e = lane_stack_mem;                                       \
    stk.sm = s_lane_stack + threadIdx.x;                          \
    stk.sp = 0;                                                   \
    stk.top = 0
#else
#define RT_LANE_STACK_DECL(stk)                                   \
    int lane_stack_mem[RT_STACK_SIZE];                            \
    LaneStack stk;                                                \
    stk.e = lane_stack_mem;                                       \
    stk.sm = nullptr;                                             \
    stk.sp = 0;                                                   \
    stk.top = 0
#endif

enum { MODE_PRIMARY = 0, MODE_QUEUE = 1, MODE_SHADOW = 2 };
#define RT_STEAL_RUN 1        // 8x4-pixel blocks claimed per system-scope atomic.  Measured at 2 GPUs: runs of 4 lengthen
                              // the kernel tail (1.55 ms vs 1.29 ms per frame): stolen work is the LAST work, keep it fine-grained

struct CamDev {
    double pos[3], u[3], v[3], w[3];
    double focal, aspect;
    int W, H;
    const float* xw;      // per column / per row window coordinates (k_camera_tables): they depend on the frame size and
    const float* yw;      // the aspect only, so the two FP64 divisions of camera.cpp:36-37 leave the per-pixel path
};

struct FrameDev {
    int W, H, tile_w, tile_h, tiles_x, tile_pix, world;
    uint32_t n_tiles_owned, n_local_pix;   // the statically owned tiles and their pixel count
    const uint32_t* tile_ids;
    const uint32_t* tile_ord;              // canonical ordinal of the tile at each position of tile_ids (index into tile_cost)
    // dynamic tile stealing: blocks of 8x4 pixels of the pool tiles, claimed through steal_cursor
    const uint32_t* pool_ids;              // pool tile ids (the same list on every rank)
    uint32_t n_pool_blocks;                // pool tiles * blocks per tile
    uint32_t* stolen_map;                  // slot -> pool block claimed by this rank
    uint32_t* steal_cursor;                // system-scope cursor of this frame, nullptr = no stealing
    uint32_t* tile_cost;                   // per owned tile: clocks its warps spent this frame (feedback for the
                                           // heavy-tiles-first order of the next frame), nullptr = off
    uint32_t n_wide_pix;                   // batches below this local pixel index (the heaviest tiles of the order) walk
                                           // the 4-wide view (traverse_body<WIDE = 2>); 0 = none
    uint32_t wide_after_bursts;            // ... and any batch moves to it once it has run this many bursts of
                                           // loop_style steps (a long batch the order did not predict); 0 = never
};

// (tile, 8x4 block, lane) -> frame pixel; lane = x%8 + 8*(y%4).
__device__ __forceinline__ bool block_to_pixel(const FrameDev& f, uint32_t tile, uint32_t b, uint32_t l, int& i, int& j) {
    int tx = (int)(tile % (uint32_t)f.tiles_x), ty = (int)(tile / (uint32_t)f.tiles_x);
    int bpr = f.tile_w >> 3;
    i = tx * f.tile_w + (int)(b % (uint32_t)bpr) * 8 + (int)(l & 7);
    j = ty * f.tile_h + (int)(b / (uint32_t)bpr) * 4 + (int)(l >> 3);
    return i < f.W && j < f.H;
}

// local pixel index -> frame pixel.  Indices below n_local_pix walk the owned tiles in 8x4 blocks;
// indices above address blocks stolen from the pool, 32 per slot of stolen_map.
__device__ __forceinline__ bool local_to_pixel(const FrameDev& f, uint32_t lp, int& i, int& j) {
    if (lp < f.n_local_pix) {
        uint32_t tl = lp / (uint32_t)f.tile_pix, p = lp % (uint32_t)f.tile_pix;
        return block_to_pixel(f, __ldg(f.tile_ids + tl), p >> 5, p & 31, i, j);
    }
    uint32_t slot = (lp - f.n_local_pix) >> 5;
    uint32_t pb = f.stolen_map[slot], bpt = (uint32_t)f.tile_pix >> 5;
    return block_to_pixel(f, __ldg(f.pool_ids + pb / bpt), pb % bpt, lp & 31, i, j);
}
__device__ __forceinline__ uint32_t pixel_in_tile_to_local(const FrameDev& f, uint32_t tl, int x, int y) {
    int bpr = f.tile_w >> 3;
    int b = (y >> 2) * bpr + (x >> 3);
    int l = (y & 3) * 8 + (x & 7);
    return tl * (uint32_t)f.tile_pix + (uint32_t)(b * 32 + l);
}

// Output byte offset of an owned pixel: frame layout (camera.cpp:48) or this rank's packed tiles.  The
// packed slot of tile t is t / world (ascending owned order, what k_assemble expects) whatever order the
// tiles are processed in.
__device__ __forceinline__ size_t pixel_byte_offset(const FrameDev& f, int i, int j, int packed) {
    if (!packed) return ((size_t)i + (size_t)j * f.W) * 3;
    int tx = i / f.tile_w, ty = j / f.tile_h;
    uint32_t slot = (uint32_t)(ty * f.tiles_x + tx) / (uint32_t)f.world;
    return ((size_t)slot * f.tile_pix + (size_t)(j - ty * f.tile_h) * f.tile_w + (size_t)(i - tx * f.tile_w)) * 3;
}

// Camera::get_ray_direction (camera.cpp:33-44) + the Ray constructor's normalisation (ray.h:25-29),
// evaluated in FP64 like the reference and rounded once to FP32.
__device__ __forceinline__ void primary_ray(const CamDev& c, int i, int j, f3& o, f3& d) {
    float xw = __ldg(c.xw + i);
    float yw = __ldg(c.yw + j);
    double dx = -c.w[0] * c.focal + c.u[0] * (double)xw + c.v[0] * (double)yw;
    double dy = -c.w[1] * c.focal + c.u[1] * (double)xw + c.v[1] * (double)yw;
    double dz = -c.w[2] * c.focal + c.u[2] * (double)xw + c.v[2] * (double)yw;
    // get_ray_direction normalises and Ray's constructor normalises again; the second pass moves the
    // FP64 value by <= 1 ulp, far below the final rounding to FP32, so one reciprocal length is used
    double inv = rsqrt(dx * dx + dy * dy + dz * dz);
    d = mk3((float)(dx * inv), (float)(dy * inv), (float)(dz * inv));
    o = mk3((float)c.pos[0], (float)c.pos[1], (float)c.pos[2]);
}

// xw[i] = (float)(aspect (i - W/2 + 0.5) / W), yw[j] = (float)((j - H/2 + 0.5) / H): camera.cpp:36-37 in FP64 with the
// reference's conversion to float; rebuilt only when the frame size or the aspect changes.
__global__ void k_camera_tables(float* __restrict__ xw, float* __restrict__ yw, int W, int H, double aspect) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < W) xw[k] = (float)(aspect * (k - W / 2.0 + 0.5) / W);
    else if (k < W + H) yw[k - W] = (float)((k - W - H / 2.0 + 0.5) / H);
}

// Color::clamp + Camera::drawPixel (color.cpp:19-28, camera.cpp:46-52): truncating 8-bit conversion of
// a 32.32 fixed-point channel.
__device__ __forceinline__ uint32_t to_u8(long long q) {
    double c = (double)q * (1.0 / 4294967296.0);
    if (c > 1.0) c = 1.0;
    if (c < 0.0) c = 0.0;
    return (uint32_t)(255.0 * c);
}

__device__ __forceinline__ long long to_fixed(float x) {
    if (!(x == x)) x = 0.0f;                           // NaN -> 0
    x = fminf(fmaxf(x, -1048576.0f), 1048576.0f);
    return __float2ll_rn(x * 4294967296.0f);
}

// Scenes without bounces get exactly one contribution per pixel (primary miss or primary hit): it is
// converted and stored as RGB8 on the spot — same rounding chain as accumulate + k_resolve — and the
// accumulator buffer and the resolve kernel are skipped altogether.
__device__ __forceinline__ void write_pixel_direct(uint8_t* out, size_t byte0, f3 c) {
    out[byte0 + 0] = (uint8_t)to_u8(to_fixed(c.x));
    out[byte0 + 1] = (uint8_t)to_u8(to_fixed(c.y));
    out[byte0 + 2] = (uint8_t)to_u8(to_fixed(c.z));
}

__device__ __forceinline__ uint32_t pack_rgb8(f3 c) {      // the same three bytes, as R | G << 8 | B << 16
    return to_u8(to_fixed(c.x)) | to_u8(to_fixed(c.y)) << 8 | to_u8(to_fixed(c.z)) << 16;
}

// FIRST: the wave-0 contribution of a pixel (exactly one per in-frame pixel: the primary miss in
// k_traverse or the primary hit in k_shade) is a plain store, which also initialises the
// accumulator — no memset of the frame; every later contribution is an integer atomic add.
template <bool FIRST>
__device__ __forceinline__ void accumulate(long long* accum, uint32_t pix, f3 c) {
    float v[3] = {c.x, c.y, c.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        long long q = to_fixed(v[k]);
        if (FIRST) accum[3 * (size_t)pix + k] = q;
        else if (q) atomicAdd((unsigned long long*)(accum + 3 * (size_t)pix + k), (unsigned long long)q);
    }
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct TravArgs {
    SceneDev s;
    CamDev cam;
    FrameDev f;
    RayQueue q;               // QUEUE: the rays; SHADOW: the rays of the wave whose hits are tested
    const float4* hits_in;    // SHADOW: hit records of the wave
    const uint32_t* hitq_in;  // SHADOW: ray index of each hit
    float4* hits;             // nearest modes: hit records out (compacted)
    uint32_t* hitq;
    uint8_t* occl;            // SHADOW: [light][hit] 1 = occluded
    WaveCounters* wave;
    long long* accum;
    FrameCounters* fc;
    uint32_t* sticky;         // error bits that survive until the next synchronising call
    int32_t* aux_prim;
    float* aux_t;
    uint32_t brute;
    uint32_t cap;             // ray-queue capacity (bounds n_rays after an overflow)
    uint32_t primary_wave;    // SHADOW: the wave's rays are primary rays (regenerated from the pixel)
    int refill_min;           // idle lanes a warp waits for before it fetches new rays
    int loop_style;           // 0: while-while; k > 0: if-if in bursts of k steps
    unsigned long long* warp_times;   // debug (RT_FLAG_WARP_TIMES): per warp {start, end} in ns, 2 per warp
    uint8_t* direct_rgb;      // PRIMARY, scene without bounces: RGB8 output written in place of the accumulator
    int direct_packed;
    RayQueue qout;            // SHADE: bounce rays of the shaded hits
    WaveCounters* next;       // SHADE: their counter
    int max_depth;            // SHADE
    int remote_out;           // SHADE: direct_rgb is another GPU's frame (informational: same code path)
};

// Appends the bounce rays of this warp's shaded hits to the next wave's queue: exclusive prefix over the
// warp, one atomic per warp.  w is the throughput of the shaded ray.
__device__ __forceinline__ void append_children(const RayQueue& qout, WaveCounters* next, uint32_t cap, const ShadeOut& out,
                                                bool valid, f3 w, uint32_t pix, int lane, bool& qfull) {
    uint32_t k = valid ? (uint32_t)out.n_children : 0u;
    uint32_t incl = k;
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o2);
        if (lane >= o2) incl += t;
    }
    uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total) {
        uint32_t qbase = 0;
        if (lane == 31) qbase = atomicAdd(&next->n_rays, total);
        qbase = __shfl_sync(0xffffffffu, qbase, 31);
        uint32_t at = qbase + incl - k;
        for (uint32_t c = 0; c < k; c++) {
            if (at + c >= cap) { qfull = true; break; }
            const ShadeChild& ch = out.child[c];
            f3 cw = w * ch.w;
            qout.o_pix[at + c] = make_float4(ch.o.x, ch.o.y, ch.o.z, __uint_as_float(pix));
            qout.d_lvl[at + c] = make_float4(ch.d.x, ch.d.y, ch.d.z, __int_as_float(ch.level));
            qout.w[at + c] = make_float4(cw.x, cw.y, cw.z, 0.0f);
        }
    }
}

// SHADE mode, RGB8 output: the warp owns an 8x4 pixel block (lane = x%8 + 8*(y%4)) whose colours are all final
// at the end of the batch.  Instead of 96 single-byte stores scattered over the batch's lifetime, the four
// 24-byte rows leave as twelve 8-byte stores of one instruction — what a frame that sits in another GPU's
// memory (NVLink) needs, and cheaper for the local L2 too.  rgb: this lane's pixel as R | G<<8 | B<<16; have: the
// lane's pixel is inside the frame; (i0, j0): the block's top-left pixel; row_stride in bytes.
__device__ __forceinline__ void store_block_rgb8(uint8_t* out, size_t block_byte0, size_t row_stride, uint32_t rgb, bool have,
                                                 bool aligned8, int lane) {
    const unsigned FULL = 0xffffffffu;
    if (__all_sync(FULL, have) && aligned8) {
        // lanes 0..11: row = lane / 3, bytes [8 * part, 8 * part + 8) of the row's 24, part = lane % 3
        int row = (lane < 12 ? lane : 0) / 3, part = (lane < 12 ? lane : 0) % 3;
        int p0 = (8 * part) / 3, off = (8 * part) % 3;                 // first pixel of the span, byte offset inside it
        uint32_t c0 = __shfl_sync(FULL, rgb, row * 8 + p0);
        uint32_t c1 = __shfl_sync(FULL, rgb, row * 8 + p0 + 1);
        uint32_t c2 = __shfl_sync(FULL, rgb, row * 8 + p0 + 2);
        uint32_t c3 = __shfl_sync(FULL, rgb, row * 8 + ((p0 + 3) & 7));
        // 12 bytes of pixel data starting at pixel p0, as 96 bits; take 8 bytes from byte `off`
        unsigned long long lo = (unsigned long long)c0 | ((unsigned long long)c1 << 24) | ((unsigned long long)c2 << 48);
        unsigned long long hi = ((unsigned long long)c2 >> 16) | ((unsigned long long)c3 << 8);
        unsigned long long v = off == 0 ? lo : (lo >> (8 * off)) | (hi << (64 - 8 * off));
        if (lane < 12) *reinterpret_cast<unsigned long long*>(out + block_byte0 + (size_t)row * row_stride + 8 * part) = v;
    } else if (have) {
        uint8_t* o = out + block_byte0 + (size_t)(lane >> 3) * row_stride + (size_t)(lane & 7) * 3;
        o[0] = (uint8_t)rgb; o[1] = (uint8_t)(rgb >> 8); o[2] = (uint8_t)(rgb >> 16);
    }
}

// SHADE mode of k_traverse: the warp's batch is done, shade its pending primary hits together.
// Returns true if the bounce-ray queue overflowed.
__device__ __forceinline__ bool shade_batch(const TravArgs& a, bool pending, uint32_t pm, HitRec nh, uint32_t occl_mask,
                                         f3 pd, uint32_t pix, uint32_t& px_rgb, bool& px_have) {
    const int lane = threadIdx.x & 31;
    bool qfull = false;
    ShadeOut out;
    out.n_children = 0;
    if (pending) {
        const f3 bg = mk3(a.s.background[0], a.s.background[1], a.s.background[2]);
        f3 po = mk3((float)a.cam.pos[0], (float)a.cam.pos[1], (float)a.cam.pos[2]);   // primary_ray's origin
        uint32_t li = 0;
        auto any_hit = [&](f3, f3) -> bool { bool o2 = ((occl_mask >> li) & 1u) != 0; li++; return o2; };
        shade_hit(a.s, po, pd, 0, nh, a.max_depth, any_hit, out);
        f3 contrib = out.local + out.bg_weight * bg;
        if (a.direct_rgb) { px_rgb = pack_rgb8(contrib); px_have = true; }     // stored with the rest of the block
        else accumulate<true>(a.accum, pix, contrib);
    }
    append_children(a.qout, a.next, a.cap, out, pending, mk3(1, 1, 1), pix, lane, qfull);
    if (lane == 0) atomicAdd(&a.wave->n_hits, (uint32_t)__popc(pm));
    return qfull;
}

// FUSE (nearest-hit modes only): a lane whose ray hit something does not go idle — it turns into
// the shadow ray(s) of that hit (one per light, same walk with early exit) and only then emits the
// hit together with its occlusion bits.  One kernel and one tail per wave instead of two, no second
// derivation of the hit point, and lanes whose rays missed keep pulling new rays meanwhile.
// WIDE = 1: walk the 4-wide view of the tree (SceneDev::nodes4); a template parameter so that the binary walk
// keeps its register budget.  (A hybrid that moved a ray to the wide view after 32 steps was measured in round 2:
// it removes the kernel's tail but costs the bulk 20 % at 80 registers — profiles/r2_tuning.md — and was deleted.)
// WIDE = 2 (whole-batch refill only): both walks, chosen per 32-pixel batch.  The batches of the tiles at the head of the
// heavy-tiles-first order walk the wide view, everything else the binary tree.  The end of a rank's share of a multi-GPU
// frame is ONE batch whose longest ray is a chain of ~240 dependent steps; on the wide view the same ray takes ~40 %
// fewer, and the few hundred batches that walk it do not cost the bulk its cheaper binary step.  The choice is
// warp-uniform and fixed for the batch, so each walk keeps its own loop (no per-step branch, unlike the hybrid).
// SHADE (fused primary rays, whole-batch refill): the hit is not queued for k_shade either.  A lane keeps its
// hit and occlusion bits until the warp's 32-pixel batch is done, then the warp shades all its hits together
// (World::shade_ray, world.cpp:32-111, same shade_hit as k_shade), writes the pixels and queues the bounce
// rays.  One kernel per primary wave: no hit queue traffic, no second launch, no second tail.
#ifndef RT_DENSE_MIN_BLOCKS
#define RT_DENSE_MIN_BLOCKS 9           // k_frame / k_frame_push for large shares: 56 registers, 9 CTAs per SM
#endif
#ifndef RT_FRAME_WIDE
#define RT_FRAME_WIDE 2                 // k_frame / k_frame_push: 2 = heavy batches walk the wide view, 0 = binary walk only
#endif
// Both walks are compiled into the 72-register build only (small shares: that is where one straggler batch ends the
// launch).  In the 56-register build the second loop's registers push per-burst state into local memory: measured 5 % on
// the whole 4K frame (1.081 -> 1.137 ms) with the wide walk never taken, and a 1/8 share does not gain from it there.
__host__ __device__ constexpr int frame_wide_mode(int minb) { return minb >= RT_DENSE_MIN_BLOCKS ? 0 : RT_FRAME_WIDE; }
#ifndef RT_SHADE_FUSED_MIN_BLOCKS
#define RT_SHADE_FUSED_MIN_BLOCKS 7     // SHADE: hold the kernel to the traversal loop's 72 registers (the once-per-batch shading spills)
#endif
// The body of the traversal kernels (k_traverse, and phase 1 of k_frame).  Returns the number of work items this
// warp claimed from the cursor (warp-uniform): k_frame's "everything is traced" barrier counts them.
template <int MODE, bool COUNT, bool FUSE, int WIDE, bool SHADE>
__device__ __forceinline__ uint32_t traverse_body(const TravArgs& a, float* s_pdir) {
    constexpr bool ANY = MODE == MODE_SHADOW;
    static_assert(!(FUSE && ANY), "FUSE applies to the nearest-hit modes");
    static_assert(!WIDE || FUSE, "the wide walk is instantiated for the fused kernels only");
    static_assert(WIDE >= 0 && WIDE <= 2, "WIDE: 0 binary walk, 1 4-wide view, 2 chosen per batch");
    static_assert(WIDE != 2 || MODE == MODE_PRIMARY, "the per-batch choice needs whole batches of primary rays");
    static_assert(!SHADE || (FUSE && MODE == MODE_PRIMARY), "in-kernel shading is for the fused primary wave");
    const unsigned FULL = 0xffffffffu;
    uint32_t claimed = 0;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t n, n_hits_in = 0;
    if (MODE == MODE_PRIMARY) n = a.f.n_local_pix;
    else if (MODE == MODE_QUEUE) n = min(a.wave->n_rays, a.cap);
    else { n_hits_in = a.wave->n_hits; n = n_hits_in * (uint32_t)a.s.n_lights; }
    uint32_t* cursor = ANY ? &a.wave->fetch_shade : &a.wave->fetch_trace;
    unsigned long long t_start = 0;
    if (a.warp_times) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    const f3 bg = mk3(a.s.background[0], a.s.background[1], a.s.background[2]);
    const bool use_bvh = !a.brute && a.s.n_bvh_tris > 0;

    // per-lane ray state
    RayPrep r;
    HitRec hit;
    RT_LANE_STACK_DECL(stk);
    int node = RT_DONE;
    bool active = false, found = false, exhausted = false, overflow = false;
    bool steal_done = !(MODE == MODE_PRIMARY && a.f.steal_cursor != nullptr && a.f.n_pool_blocks > 0);
    uint32_t steal_left = 0, steal_slot = 0;
    uint32_t batch_tile = 0xffffffffu;
    long long batch_t0 = 0;
    bool wide_now = false;                 // WIDE == 2: the current batch walks the 4-wide view (warp-uniform)
    uint32_t bursts = 0;                   // WIDE == 2: traversal bursts the current batch has run (warp-uniform)
    uint32_t item = 0, pix = 0;
    int pi = 0, pj = 0;
    f3 w = mk3(1, 1, 1);
    WorkCount wc, wcs;
    wc.nodes = wc.tris = wcs.nodes = wcs.tris = 0;
    uint32_t traced = 0, traced_shadow = 0;
    // FUSE: phase < 0 = nearest-hit query, phase = li >= 0 = shadow ray towards light li of the saved hit
    int phase = -1;
    HitRec nh;
    nh.t = RT_FLT_MAX; nh.prim = RT_MISS; nh.beta = nh.gamma = 0.0f;
    f3 P = mk3(0, 0, 0);
    uint32_t occl_mask = 0;
    bool pending = false, qfull = false;   // SHADE: this lane's hit waits for the end of the batch
    uint32_t px_rgb = 0;                   // SHADE, RGB8 output: the lane's finished pixel, stored with its 8x4 block
    bool px_have = false;
    auto start_shadow = [&](int li) {
        f3 toL = mk3(__ldg(a.s.lights + 2 * li)) - P;
        f3 sd = normalize(toL);
        r = prep_ray(fma3(toL, 0.01f, P), sd);                 // world.cpp:45
        hit.t = RT_FLT_MAX; hit.prim = RT_MISS;
        found = false;
        stk.clear();
        node = (use_bvh && sd.x == sd.x && sd.y == sd.y && sd.z == sd.z) ? 0 : RT_DONE;
        traced_shadow++;
    };

    // the warp has just finished a whole 32-pixel batch (all lanes idle, hits shaded)
    auto finish_batch = [&]() {
        if (SHADE && a.direct_rgb && __any_sync(FULL, px_have)) {
            // lane 0 is the block's top-left pixel (always inside the frame when any pixel of the block is)
            int i0 = __shfl_sync(FULL, pi, 0), j0 = __shfl_sync(FULL, pj, 0);
            size_t b0 = pixel_byte_offset(a.f, i0, j0, a.direct_packed);
            size_t stride = (size_t)(a.direct_packed ? a.f.tile_w : a.f.W) * 3;
            bool al = (((uintptr_t)a.direct_rgb + b0) & 7) == 0 && (stride & 7) == 0;
            store_block_rgb8(a.direct_rgb, b0, stride, px_rgb, px_have, al, lane);
            px_have = false;
        }
        if (batch_tile == 0xffffffffu) return;
        // charge its duration to the tile
        if (a.f.tile_cost && lane == 0) {
            uint32_t cost = (uint32_t)((clock64() - batch_t0) >> 6);
            // a batch that walked the wide view is charged what the binary walk would have taken (about 3/2), so that a
            // heavy tile does not drop out of the head of the order because it was treated as heavy
            if (WIDE == 2 && wide_now) cost += cost >> 1;
            atomicMax(a.f.tile_cost + __ldg(a.f.tile_ord + batch_tile), cost);
        }
        batch_tile = 0xffffffffu;
    };

    for (;;) {
        // ---- refill idle lanes from the global cursor
        uint32_t need = __ballot_sync(FULL, !active);
        uint32_t my = 0;
        bool have = false;
        if (SHADE && need == FULL) {
            uint32_t pm = __ballot_sync(FULL, pending);
            if (pm) {
                f3 pd = mk3(s_pdir[threadIdx.x], s_pdir[TRAV_TPB + threadIdx.x], s_pdir[2 * TRAV_TPB + threadIdx.x]);
                if (shade_batch(a, pending, pm, nh, occl_mask, pd, pix, px_rgb, px_have)) qfull = true;
                pending = false;
            }
        }
        if (MODE == MODE_PRIMARY && need == FULL) finish_batch();
        if (!exhausted && ((!SHADE && __popc(need) >= a.refill_min) || need == FULL)) {
            int cnt = __popc(need), leader = __ffs(need) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(cursor, (uint32_t)cnt);
            base = __shfl_sync(FULL, base, leader);
            if (MODE == MODE_PRIMARY && a.f.tile_cost && need == FULL && base < n) {
                batch_tile = base / (uint32_t)a.f.tile_pix;
                batch_t0 = clock64();
            }
            if (WIDE == 2) { wide_now = need == FULL && base < a.f.n_wide_pix; bursts = 0; }
            if (base + (uint32_t)cnt >= n) exhausted = true;
            if (base < n) claimed += min((uint32_t)cnt, n - base);
            my = base + __popc(need & lt);
            have = !active && my < n;
        } else if (MODE == MODE_PRIMARY && exhausted && !steal_done && need == FULL) {
            // own tiles are done: work through the shared pool in runs of RT_STEAL_RUN 8x4 blocks, one
            // system-scope atomic per run (the cursor lives in rank 0's memory, reached over NVLink)
            if (steal_left == 0) {
                uint32_t pb = 0, slot = 0, cnt = 0;
                if (lane == 0) {
                    pb = atomicAdd_system(a.f.steal_cursor, (uint32_t)RT_STEAL_RUN);
                    if (pb < a.f.n_pool_blocks) {
                        cnt = min((uint32_t)RT_STEAL_RUN, a.f.n_pool_blocks - pb);
                        slot = atomicAdd(&a.fc->stolen_blocks, cnt);
                        for (uint32_t k = 0; k < cnt; k++) a.f.stolen_map[slot + k] = pb + k;
                    }
                }
                __syncwarp();
                steal_left = __shfl_sync(FULL, cnt, 0);
                steal_slot = __shfl_sync(FULL, slot, 0);
                if (steal_left == 0) steal_done = true;
            }
            if (steal_left) {
                if (WIDE == 2) { wide_now = false; bursts = 0; }
                my = a.f.n_local_pix + steal_slot * 32u + (uint32_t)lane;
                have = true;
                steal_slot++;
                steal_left--;
            }
        }
        {
            if (have) {
                f3 o, d;
                bool ok = true;
                item = my;
                if (MODE == MODE_PRIMARY) {
                    ok = local_to_pixel(a.f, my, pi, pj);
                    if (ok) { primary_ray(a.cam, pi, pj, o, d); traced++; }
                    pix = my;
                    w = mk3(1, 1, 1);
                } else if (MODE == MODE_QUEUE) {
                    float4 ro = a.q.o_pix[my], rd = a.q.d_lvl[my], rw = a.q.w[my];
                    o = mk3(ro); d = mk3(rd); w = mk3(rw);
                    pix = __float_as_uint(ro.w);
                } else {
                    uint32_t li = my / n_hits_in, pos = my - li * n_hits_in;
                    float4 hr = a.hits_in[pos];
                    uint32_t idx = a.hitq_in[pos];
                    f3 po, pd;
                    if (a.primary_wave) {
                        int qi, qj;
                        local_to_pixel(a.f, idx, qi, qj);
                        primary_ray(a.cam, qi, qj, po, pd);
                    } else {
                        po = mk3(a.q.o_pix[idx]);
                        pd = mk3(a.q.d_lvl[idx]);
                    }
                    // dielectric hits discard their local colour (world.cpp:77-100): no shadow query
                    int prim = __float_as_int(hr.y);
                    uint32_t mat = prim >= 0 ? __float_as_uint(__ldg(a.s.tris + 3 * (size_t)prim + 1).w)
                                             : a.s.analytic[rt_analytic_index(prim)].material;
                    float4 m1 = __ldg(a.s.materials + 3 * mat + 1);
                    if (m1.z > 0.0f && m1.w > 0.0f) {
                        a.occl[my] = 0;
                        ok = false;
                    } else {
                        f3 P = fma3(pd, hr.x, po);
                        f3 toL = mk3(__ldg(a.s.lights + 2 * li)) - P;
                        o = fma3(toL, 0.01f, P);              // world.cpp:45
                        d = normalize(toL);
                        traced++;
                    }
                }
                if (ok) {
                    hit.t = RT_FLT_MAX; hit.prim = RT_MISS; hit.beta = hit.gamma = 0.0f;
                    found = false;
                    active = true;
                    phase = -1;
                    stk.clear();
                                // a NaN direction (ignored refract() failure, world.cpp:83) misses everything
                    bool finite = d.x == d.x && d.y == d.y && d.z == d.z;
                    r = prep_ray(o, d);
                    node = (use_bvh && finite) ? 0 : RT_DONE;
                    if (!finite) r.d = mk3(d.x, d.y, d.z);
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted && steal_done) {
                if (MODE == MODE_PRIMARY) finish_batch();      // a last batch that lies outside the frame
                break;
            }
            continue;
        }
        // ---- traversal of the lanes that own a ray
        const bool any = FUSE ? phase >= 0 : ANY;
        WorkCount* wcp = COUNT ? (any ? &wcs : &wc) : nullptr;
        // a batch that is still running after wide_after_bursts bursts is a long one the order did not announce (a moving
        // camera, the first frames of a layout): it continues on the wide view — both views index the same nodes and
        // leaves, so the lanes' stacks and current nodes carry over
        if (WIDE == 2 && a.f.wide_after_bursts && bursts++ >= a.f.wide_after_bursts) wide_now = true;
        if (active) {
            if (WIDE == 2 && wide_now) {
                // a heavy batch: the same walk on the 4-wide view (its own loop: the binary loop below keeps its code)
                const int iters = a.loop_style == 0 ? 64 : a.loop_style;
                for (int it = 0; it < iters && node != RT_DONE; it++)
                    node = bvh4_step_unified(a.s, r, hit, node, stk, any, found, &overflow, wcp);
            } else if (a.loop_style == 0) {
                while (rt_is_internal(node)) {
                    if (COUNT) wcp->nodes++;
                    node = WIDE == 1
                               ? bvh4_node_step(a.s, r, hit.t, node, stk, &overflow)
                               : bvh_node_step(a.s, r, hit.t, node, stk, &overflow);
                }
                while (node < 0) {
                    if (leaf_test(a.s, node, r, hit, any, wcp)) {
                        found = true;
                        if (any) { node = RT_DONE; break; }
                    }
                    node = stk.empty() ? RT_DONE : stk.pop();
                }
            } else {
                // "if-if": every lane advances one step of whatever kind per iteration, for a bounded
                // number of iterations before the warp looks at its refill state again
                for (int it = 0; it < a.loop_style && node != RT_DONE; it++)
                    node = WIDE == 1 ? bvh4_step_unified(a.s, r, hit, node, stk, any, found, &overflow, wcp)
                                     : bvh_step_unified(a.s, r, hit, node, stk, any, found, &overflow, wcp);
            }
        }
        // ---- rays that ran out of nodes: linear primitives, then the result
        bool fin = active && node == RT_DONE;
        bool emit = false;
        if (fin) {
            bool finite = r.d.x == r.d.x && r.d.y == r.d.y && r.d.z == r.d.z;
            if (finite && !(any && found)) {
                if (a.brute) found |= any ? brute_walk<true>(a.s, r, hit, wcp) : brute_walk<false>(a.s, r, hit, wcp);
                for (int k = 0; k < a.s.n_analytic && !(any && found); k++) {
                    const AnalyticPrim p = a.s.analytic[k];
                    if (COUNT) wcp->tris++;
                    if (analytic_test(p, r.o, r.d, hit.t, hit.beta, hit.gamma)) {
                        hit.prim = rt_analytic_code(k);
                        found = true;
                    }
                }
            }
            if (any) {
                if (FUSE) {
                    occl_mask |= (found ? 1u : 0u) << phase;
                    phase++;
                    if (phase < a.s.n_lights) { start_shadow(phase); fin = false; }
                    else emit = true;
                } else {
                    a.occl[item] = found ? 1 : 0;
                }
            } else {
                if (a.aux_prim) {
                    size_t at = MODE == MODE_PRIMARY ? (size_t)pi + (size_t)pj * a.f.W : (size_t)pix;
                    int id = -1;
                    if (found) {
                        if (hit.prim >= 0) id = (int)__float_as_uint(__ldg(a.s.tris + 3 * (size_t)hit.prim).w);
                        else id = (int)a.s.analytic[rt_analytic_index(hit.prim)].object_id;
                    }
                    a.aux_prim[at] = id;
                    if (a.aux_t) a.aux_t[at] = found ? hit.t : RT_FLT_MAX;
                }
                if (!found) {
                    if (SHADE && a.direct_rgb) { px_rgb = pack_rgb8(w * bg); px_have = true; }
                    else if (MODE == MODE_PRIMARY && a.direct_rgb)
                        write_pixel_direct(a.direct_rgb, pixel_byte_offset(a.f, pi, pj, a.direct_packed), w * bg);
                    else
                        accumulate<MODE == MODE_PRIMARY>(a.accum, pix, w * bg);      // world.cpp:110
                } else {
                    nh = hit;
                    occl_mask = 0;
                    emit = true;
                    if (SHADE) {
                        s_pdir[threadIdx.x] = r.d.x; s_pdir[TRAV_TPB + threadIdx.x] = r.d.y; s_pdir[2 * TRAV_TPB + threadIdx.x] = r.d.z;
                    }
                    if (FUSE && a.s.n_lights > 0) {
                        // dielectric hits discard their local colour (world.cpp:77-100): no shadow query
                        uint32_t mat = hit.prim >= 0 ? __float_as_uint(__ldg(a.s.tris + 3 * (size_t)hit.prim + 1).w)
                                                     : a.s.analytic[rt_analytic_index(hit.prim)].material;
                        float4 m1 = __ldg(a.s.materials + 3 * mat + 1);
                        if (!(m1.z > 0.0f && m1.w > 0.0f)) {
                            P = fma3(r.d, hit.t, r.o);
                            phase = 0;
                            start_shadow(0);
                            emit = false;
                            fin = false;
                        }
                    }
                }
            }
            if (fin) active = false;
        }
        if (SHADE) {
            pending |= emit;
        } else if (!ANY) {
            // hits: one atomic per warp reserves a run of the hit queue
            uint32_t mask = __ballot_sync(FULL, emit);
            if (mask) {
                uint32_t qbase = 0;
                int leader = __ffs(mask) - 1;
                if (lane == leader) qbase = atomicAdd(&a.wave->n_hits, (uint32_t)__popc(mask));
                qbase = __shfl_sync(FULL, qbase, leader);
                if (emit) {
                    uint32_t pos = qbase + __popc(mask & lt);
                    a.hitq[pos] = item;
                    a.hits[pos] = make_float4(nh.t, __int_as_float(nh.prim), nh.beta, nh.gamma);
                    if (FUSE) a.occl[pos] = (uint8_t)occl_mask;
                }
            }
        }
    }
    if (ANY || MODE == MODE_PRIMARY) {
        uint32_t ns = warp_sum(traced);
        if (lane == 0 && ns) atomicAdd(ANY ? &a.fc->rays_shadow : &a.fc->rays_primary, (unsigned long long)ns);
    }
    if (FUSE) {
        uint32_t ns = warp_sum(traced_shadow);
        if (lane == 0 && ns) atomicAdd(&a.fc->rays_shadow, (unsigned long long)ns);
    }
    if (COUNT) {
        uint32_t nn = warp_sum(wc.nodes), nt = warp_sum(wc.tris);
        uint32_t sn = warp_sum(wcs.nodes), st = warp_sum(wcs.tris);
        if (lane == 0) {
            atomicAdd(&a.fc->node_visits[ANY ? 1 : 0], (unsigned long long)nn);
            atomicAdd(&a.fc->tri_tests[ANY ? 1 : 0], (unsigned long long)nt);
            if (FUSE) {
                atomicAdd(&a.fc->node_visits[1], (unsigned long long)sn);
                atomicAdd(&a.fc->tri_tests[1], (unsigned long long)st);
            }
        }
    }
    if (__any_sync(FULL, overflow) && lane == 0) atomicOr(a.sticky, 2u);
    if (SHADE && __any_sync(FULL, qfull) && lane == 0) atomicOr(a.sticky, 1u);
    if (a.warp_times && lane == 0) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        a.warp_times[2 * wid] = t_start;
        a.warp_times[2 * wid + 1] = t_end;
    }
    return claimed;
}

template <int MODE, bool COUNT, bool FUSE, int WIDE = 0, bool SHADE = false>
__global__ void __launch_bounds__(TRAV_TPB, (SHADE || (FUSE && MODE == MODE_PRIMARY)) ? RT_SHADE_FUSED_MIN_BLOCKS : 0) k_traverse(const __grid_constant__ TravArgs a) {
    __shared__ float s_pdir[SHADE ? 3 * TRAV_TPB : 1];   // SHADE: primary direction of the lane's pending hit
    traverse_body<MODE, COUNT, FUSE, WIDE, SHADE>(a, s_pdir);
}

struct ShadeArgs {
    SceneDev s;
    CamDev cam;
    FrameDev f;
    RayQueue qin, qout;
    const float4* hits;
    const uint32_t* hitq;
    const uint8_t* occl;
    WaveCounters* wave;
    WaveCounters* next;
    long long* accum;
    uint32_t* sticky;
    uint32_t cap;
    int max_depth;
    int occl_bits;            // 1: occl[hit] is a bit mask over lights (fused traversal); 0: occl[light][hit] bytes
    uint8_t* direct_rgb;      // see TravArgs
    int direct_packed;
};

// Shading proper: no traversal in here, the shadow answers come from k_traverse<SHADOW>.
#ifndef RT_SHADE_MIN_BLOCKS
#define RT_SHADE_MIN_BLOCKS 1
#endif
template <bool PRIMARY>
__global__ void __launch_bounds__(SHADE_TPB, RT_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ ShadeArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t n = a.wave->n_hits;
    const f3 bg = mk3(a.s.background[0], a.s.background[1], a.s.background[2]);
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    bool qfull = false;
    for (uint32_t base = warp * 32u; base < n; base += warps * 32u) {
        uint32_t pos = base + lane;
        bool valid = pos < n;
        ShadeOut out;
        out.n_children = 0;
        f3 w = mk3(1, 1, 1);
        uint32_t pix = 0;
        if (valid) {
            uint32_t idx = a.hitq[pos];
            float4 hr = a.hits[pos];
            HitRec h;
            h.t = hr.x; h.prim = __float_as_int(hr.y); h.beta = hr.z; h.gamma = hr.w;
            f3 o, d;
            int level = 0;
            int pi = 0, pj = 0;
            if (PRIMARY) {
                local_to_pixel(a.f, idx, pi, pj);
                primary_ray(a.cam, pi, pj, o, d);
                pix = idx;
            } else {
                float4 ro = a.qin.o_pix[idx], rd = a.qin.d_lvl[idx], rw = a.qin.w[idx];
                o = mk3(ro); d = mk3(rd); w = mk3(rw);
                pix = __float_as_uint(ro.w);
                level = __float_as_int(rd.w);
            }
            uint32_t li = 0;
            auto any_hit = [&](f3, f3) -> bool {
                bool o2 = a.occl_bits ? ((a.occl[pos] >> li) & 1u) != 0 : a.occl[(size_t)li * n + pos] != 0;
                li++;
                return o2;
            };
            shade_hit(a.s, o, d, level, h, a.max_depth, any_hit, out);
            f3 contrib = w * (out.local + out.bg_weight * bg);
            if (PRIMARY && a.direct_rgb) write_pixel_direct(a.direct_rgb, pixel_byte_offset(a.f, pi, pj, a.direct_packed), contrib);
            else accumulate<PRIMARY>(a.accum, pix, contrib);
        }
        append_children(a.qout, a.next, a.cap, out, valid, w, pix, lane, qfull);
    }
    if (__any_sync(0xffffffffu, qfull) && lane == 0) atomicOr(a.sticky, 1u);
}


// ---- k_frame: a whole bounce-free frame in ONE launch --------------------------------------------------------------
// The north-star workload (config 4: primary + shadow rays, no bounces) used to take a traversal kernel, a shading
// kernel and — when the frame lives on another GPU — a push kernel between two single-thread handshake kernels: five
// launches whose gaps and tails are a third of an 8-GPU frame.  Here the persistent warps run all of it:
//   phase 1  traverse_body<PRIMARY, FUSE>: nearest hits + their shadow rays; misses write their pixel, hits go to
//            the compacted hit queue (same code as k_traverse)
//   barrier  over the CTAs of the grid (one atomic + one poller per CTA on a cumulative counter); the launch is
//            cooperative, so all CTAs are resident and the barrier cannot deadlock
//   phase 2  the hit queue is shaded with full warps (k_shade's code), rounds of 32 hits dealt out statically, pixels
//            written as RGB8
//   barrier  (frame elsewhere) the same
//   phase 3  (frame elsewhere) this rank's packed tiles -> the shared frame with 16-byte stores over NVLink, after
//            rank 0's "previous frame consumed" flag
//   exit     the last CTA to finish (done[2]) signals rank 0's arrival slot with a system-scope atomic; on rank 0 it
//            waits for the other ranks' arrivals instead, so the completion of rank 0's kernel IS the completion of
//            the frame.  No handshake kernels, no collective.
struct FrameSyncDev {
    uint32_t* done;                 // CTAs past [0] tracing, [1] shading, [2] the push — cumulative over all frames, never
                                    // reset: this launch's barriers release at target[k] (the counters wrap)
    uint32_t target[3];
    uint32_t* zero_words;           // the counter set of the NEXT frame (double-buffered): cleared by this launch
    uint32_t n_zero_words;
    uint32_t* sync;                 // rt_peer_sync layout in rank 0's memory (peer mapped), or nullptr: no other ranks
    uint32_t frame, rank, world;
};
struct PushDev {                    // phase 3; packed == nullptr: pixels were written in place, nothing to push
    const uint8_t* packed;
    uint8_t* frame;
    uint32_t tiles_total;
    int wide16;                     // rows of tile and frame are 16-byte multiples, both buffers 16-byte aligned
};
struct FrameArgs {
    TravArgs t;
    FrameSyncDev y;
    PushDev push;
    int max_depth;
    unsigned long long* phase_times;   // debug (RT_FLAG_WARP_TIMES): 8 time stamps per warp
    int phase1_only;                   // debug (RT_FK_SPLIT=1): stop after the tracing phase; the host launches k_shade
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *(const volatile uint32_t*)p; }

// Spin until *p >= want; gives up after ~2 s (sticky bit 4) like the handshake kernels.
__device__ __forceinline__ void spin_until_ge(const uint32_t* p, uint32_t want, uint32_t* sticky, unsigned ns) {
    long long t0 = clock64();
    while (ld_volatile_u32(p) < want) {
        if (clock64() - t0 > 4000000000ll) { atomicOr(sticky, 4u); break; }
        __nanosleep(ns);
    }
}

// The same for a cumulative counter that wraps: until *p has reached `target`.
__device__ __forceinline__ void spin_until_reached(const uint32_t* p, uint32_t target, uint32_t* sticky, unsigned ns) {
    long long t0 = clock64();
    while ((int32_t)(ld_volatile_u32(p) - target) < 0) {
        if (clock64() - t0 > 4000000000ll) { atomicOr(sticky, 4u); break; }
        __nanosleep(ns);
    }
}

// MINB: resident CTAs per SM the register budget is held to.  RT_SHADE_FUSED_MIN_BLOCKS (7 -> 72 registers) gives the
// fastest single warp and with it the shortest tail: right for a rank's share of a multi-GPU frame.  RT_DENSE_MIN_BLOCKS
// (9 -> 56 registers, cold per-lane state spilled, the hot loop unchanged) gives 29 % more warps per SM: the bulk of a
// large frame runs 6.5 % faster, the stragglers 15 % slower.  The host picks by the pixels this launch owns.
template <bool COUNT, int MINB = RT_SHADE_FUSED_MIN_BLOCKS>
__global__ void __launch_bounds__(TRAV_TPB, MINB) k_frame(const __grid_constant__ FrameArgs a) {
    // (nothing is kept live across phase 1 that phase 1 does not need: its loop sits exactly at the register budget)
    auto stamp = [&](int k) {
        if (a.phase_times && (threadIdx.x & 31) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.phase_times[8 * (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) + k] = t;
        }
    };
    stamp(0);
    if (blockIdx.x == 0 && threadIdx.x < a.y.n_zero_words) a.y.zero_words[threadIdx.x] = 0u;   // next frame's counters
    // rank 0 opens the frame for the other ranks: everything enqueued on its stream for the previous frame (a copy to
    // the host, say) has finished before this kernel started
    if (a.y.sync && a.y.rank == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        volatile uint32_t* consumed = a.y.sync + 64;
        if (*consumed < a.y.frame) *consumed = a.y.frame;
        __threadfence_system();
    }
    // ---- phase 1
    traverse_body<MODE_PRIMARY, COUNT, true, frame_wide_mode(MINB), false>(a.t, nullptr);
    stamp(1);
    if (a.phase1_only) return;
    // every hit of the frame is in the queue once all CTAs are here
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.y.done + 0, 1u);
        spin_until_reached(a.y.done + 0, a.y.target[0], a.t.sticky, 200);
        __threadfence();
    }
    __syncthreads();
    stamp(2);
    // ---- phase 2: shade the hit queue (World::shade_ray's local term, world.cpp:40-63, :126-137), 32 hits per warp
    // and round, rounds dealt out statically
    const int lane = threadIdx.x & 31;
    const bool pushing = a.push.packed != nullptr;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_hits = __ldcg(&a.t.wave->n_hits);
    const f3 bg = mk3(a.t.s.background[0], a.t.s.background[1], a.t.s.background[2]);
    for (uint32_t base = gwarp * 32u; base < n_hits; base += n_warps * 32u) {
        uint32_t pos = base + lane;
        if (pos < n_hits) {
            uint32_t idx = __ldcg(a.t.hitq + pos);
            float4 hr = __ldcg(a.t.hits + pos);
            uint32_t om = (uint32_t)__ldcg(a.t.occl + pos);
            HitRec h;
            h.t = hr.x; h.prim = __float_as_int(hr.y); h.beta = hr.z; h.gamma = hr.w;
            int pi = 0, pj = 0;
            f3 o, d;
            local_to_pixel(a.t.f, idx, pi, pj);
            primary_ray(a.t.cam, pi, pj, o, d);
            ShadeOut out;
            uint32_t li = 0;
            auto any_hit = [&](f3, f3) -> bool { bool o2 = ((om >> li) & 1u) != 0; li++; return o2; };
            shade_hit(a.t.s, o, d, 0, h, a.max_depth, any_hit, out);
            f3 contrib = out.local + out.bg_weight * bg;
            write_pixel_direct(a.t.direct_rgb, pixel_byte_offset(a.t.f, pi, pj, a.t.direct_packed), contrib);
        }
    }
    stamp(3);
    if (!pushing) return;                                  // pixels were written in place: the frame is done
    // barrier: every CTA has shaded its rounds
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        // frame open?  one CTA asks rank 0 over NVLink, BEFORE it joins the barrier: the others learn it from the barrier
        if (blockIdx.x == 0 && a.y.sync && a.y.rank != 0) spin_until_ge(a.y.sync + 64, a.y.frame, a.t.sticky, 100);
        atomicAdd(a.y.done + 1, 1u);
        spin_until_reached(a.y.done + 1, a.y.target[1], a.t.sticky, 200);
        __threadfence();
    }
    __syncthreads();
    stamp(4);
    // ---- phase 3: packed tiles -> the shared frame
    {
        const FrameDev& f = a.t.f;
        const uint32_t world = (uint32_t)f.world, src = a.y.rank;
        const uint32_t n_owned = a.push.tiles_total > src ? (a.push.tiles_total - src + world - 1) / world : 0u;
        const uint32_t stride = gridDim.x * blockDim.x;
        if (a.push.wide16) {
            const uint32_t chunks_per_row = (uint32_t)(f.tile_w * 3) >> 4, chunks_per_tile = chunks_per_row * (uint32_t)f.tile_h;
            const uint32_t total = n_owned * chunks_per_tile;
            for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
                uint32_t tl = q / chunks_per_tile, r = q % chunks_per_tile;
                uint32_t tile = src + tl * world;
                int y = (int)(r / chunks_per_row), cb = (int)(r % chunks_per_row) * 16;
                int tx = (int)(tile % (uint32_t)f.tiles_x), ty = (int)(tile / (uint32_t)f.tiles_x);
                int j = ty * f.tile_h + y;
                if (j >= f.H) continue;
                int row_bytes = min(f.tile_w, f.W - tx * f.tile_w) * 3;
                if (cb >= row_bytes) continue;
                size_t src0 = ((size_t)tl * f.tile_pix + (size_t)y * f.tile_w) * 3 + cb;
                size_t dst0 = ((size_t)tx * f.tile_w + (size_t)j * f.W) * 3 + cb;
                if (cb + 16 <= row_bytes) {
                    *reinterpret_cast<uint4*>(a.push.frame + dst0) = __ldcg(reinterpret_cast<const uint4*>(a.push.packed + src0));
                } else {
                    for (int k = 0; k < row_bytes - cb; k++) a.push.frame[dst0 + k] = __ldcg(a.push.packed + src0 + k);
                }
            }
        } else {
            const uint32_t quads_per_row = (uint32_t)f.tile_w >> 2, quads_per_tile = quads_per_row * (uint32_t)f.tile_h;
            const uint32_t total = n_owned * quads_per_tile;
            for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
                uint32_t tl = q / quads_per_tile, r = q % quads_per_tile;
                uint32_t tile = src + tl * world;
                int y = (int)(r / quads_per_row), x0 = (int)(r % quads_per_row) * 4;
                int tx = (int)(tile % (uint32_t)f.tiles_x), ty = (int)(tile / (uint32_t)f.tiles_x);
                int j = ty * f.tile_h + y, i0 = tx * f.tile_w + x0;
                if (j >= f.H || i0 >= f.W) continue;
                int npx = min(4, f.W - i0);
                size_t src0 = ((size_t)tl * f.tile_pix + (size_t)y * f.tile_w + x0) * 3;
                size_t dst0 = ((size_t)i0 + (size_t)j * f.W) * 3;
                for (int k = 0; k < 3 * npx; k++) a.push.frame[dst0 + k] = __ldcg(a.push.packed + src0 + k);
            }
        }
    }
    stamp(5);
    // ---- exit: the last CTA completes the frame
    __syncthreads();
    if (threadIdx.x == 0) {
        // this CTA's stores are ordered before its count at GPU scope; the last CTA, which has observed every count,
        // publishes all of them at system scope with ONE fence before it signals (causality order is cumulative) —
        // a system-scope fence per CTA cost 7 us at the end of every frame
        __threadfence();
        uint32_t prev = atomicAdd(a.y.done + 2, 1u);
        if (prev + 1u == a.y.target[2]) {
            __threadfence_system();
            if (a.y.sync && a.y.world > 1) {
                uint32_t* slot = a.y.sync + (a.y.frame % 64u);
                if (a.y.rank != 0) {
                    atomicAdd_system(slot, 1u);
                } else {
                    spin_until_ge(slot, a.y.world - 1, a.t.sticky, 100);
                    __threadfence_system();
                    *(volatile uint32_t*)(a.y.sync + ((a.y.frame + 32u) % 64u)) = 0u;   // the slot that comes into use 32 frames on
                }
            }
        }
    }
    stamp(6);
}

// ---- k_frame_push: the multi-GPU frame of a bounce-free scene in ONE launch ------------------------------------------
// k_frame's push phase sends a rank's whole share at the END of the frame: at 8 GPUs seven ranks then store 21.8 MB into
// rank 0 at the same moment, and that burst is bound by rank 0's NVLink ingress (measured: +50 us on a 190 us frame).
// Here the stores are spread over the frame instead: the traversal body runs in its SHADE form — a warp shades the
// hits of its 32-pixel batch when the batch is done and stores the finished 8x4 block straight into the shared frame
// with 8-byte stores — and the handshake is folded into the same launch:
//   start   rank 0 publishes "frame open" (everything it enqueued for the previous frame is done); the other ranks'
//           CTAs wait for that flag before their first store (one poll per CTA)
//   exit    every CTA counts itself after a GPU-scope fence; the last one publishes all stores with one system-scope
//           fence and signals rank 0's arrival slot — on rank 0 it waits for the other ranks' arrivals instead, so the
//           end of rank 0's launch IS the completion of the frame.  No handshake kernels, no collective, no barrier
//           inside the grid (an ordinary launch).
template <bool COUNT, int MINB = RT_SHADE_FUSED_MIN_BLOCKS>
__global__ void __launch_bounds__(TRAV_TPB, MINB) k_frame_push(const __grid_constant__ FrameArgs a) {
    __shared__ float s_pdir[3 * TRAV_TPB];
    auto stamp = [&](int k) {
        if (a.phase_times && (threadIdx.x & 31) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.phase_times[8 * (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) + k] = t;
        }
    };
    stamp(0);
    if (blockIdx.x == 0 && threadIdx.x < a.y.n_zero_words) a.y.zero_words[threadIdx.x] = 0u;   // next frame's counters
    if (a.y.sync && a.y.world > 1) {
        if (a.y.rank == 0) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                volatile uint32_t* consumed = a.y.sync + 64;
                if (*consumed < a.y.frame) *consumed = a.y.frame;
                __threadfence_system();
            }
        } else {
            // ONE thread of the rank polls rank 0's flag over NVLink and republishes it in local memory, where the other
            // CTAs wait for it (thousands of remote pollers cost 20 us at the start of every frame)
            if (threadIdx.x == 0) {
                if (blockIdx.x == 0) {
                    spin_until_ge(a.y.sync + 64, a.y.frame, a.t.sticky, 100);
                    __threadfence_system();
                    *(volatile uint32_t*)(a.y.done + 3) = a.y.frame + 1u;
                    __threadfence();
                } else {
                    spin_until_reached(a.y.done + 3, a.y.frame + 1u, a.t.sticky, 100);
                    __threadfence();
                }
            }
            __syncthreads();
        }
    }
    stamp(2);
    traverse_body<MODE_PRIMARY, COUNT, true, frame_wide_mode(MINB), true>(a.t, s_pdir);
    stamp(1);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        uint32_t prev = atomicAdd(a.y.done + 2, 1u);
        if (prev + 1u == a.y.target[2]) {
            __threadfence_system();
            if (a.y.sync && a.y.world > 1) {
                uint32_t* slot = a.y.sync + (a.y.frame % 64u);
                if (a.y.rank != 0) {
                    atomicAdd_system(slot, 1u);
                } else {
                    spin_until_ge(slot, a.y.world - 1, a.t.sticky, 100);
                    __threadfence_system();
                    *(volatile uint32_t*)(a.y.sync + ((a.y.frame + 32u) % 64u)) = 0u;
                }
            }
        }
    }
    stamp(6);
}

// ---- k_paths: every bounce generation in ONE launch --------------------------------------------------
// The wave loop pays two kernel launches (and, with dielectrics, a host round trip) per generation,
// each bounded by the latency of its slowest 32-ray batch; mirror/dielectric scenes have few bounce
// rays (10^4-10^6), so that latency is all there is.  Here a lane takes one ray of the first bounce
// generation and follows its whole subtree by itself: nearest hit -> shadow rays (fused as in
// k_traverse) -> shade_hit inline -> continue with the first child, park a second child (dielectrics,
// world.cpp:97-99) on a small per-lane stack of pending rays.  No queue between lanes, no spinning,
// no host involvement; idle lanes pull new rays from the cursor.  The arithmetic per ray is the wave
// path's, and pixel sums are integer atomics, so the frame is bit-identical to the wave loop's.
#define RT_PATH_STACK 24      // pending rays per lane (binary dielectric tree, depth-first)
#pragma nv_diag_suppress 549  // `pend` is only read below psp, i.e. after it was written

struct PathArgs {
    SceneDev s;
    RayQueue q;               // the first bounce generation (written by k_shade of wave 0)
    const WaveCounters* wave; // wave->n_rays = its population
    uint32_t* cursor;
    long long* accum;
    FrameCounters* fc;
    uint32_t* sticky;
    uint32_t cap;
    int max_depth;
    int refill_min;
    int loop_style;
    uint32_t brute;
    int share;                // parked rays are handed to idle lanes of the warp (RT_PATH_SHARE)
};

template <bool COUNT, bool WIDE = false>
__global__ void __launch_bounds__(TRAV_TPB) k_paths(const __grid_constant__ PathArgs a) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t n = min(a.wave->n_rays, a.cap);
    const f3 bg = mk3(a.s.background[0], a.s.background[1], a.s.background[2]);
    const bool use_bvh = !a.brute && a.s.n_bvh_tris > 0;
    const int burst = a.loop_style > 0 ? a.loop_style : 64;

    RayPrep r;
    HitRec hit, nh;
    RT_LANE_STACK_DECL(stk);
    float4 pend[3 * RT_PATH_STACK];           // parked rays: (o, pix) (d, level) (w, -)
    int node = RT_DONE, psp = 0;
    bool active = false, found = false, exhausted = false, overflow = false, pend_overflow = false;
    int phase = -1;                           // < 0 nearest-hit query, li >= 0 shadow ray of light li
    f3 ro = mk3(0, 0, 0), rd = mk3(0, 0, 1), w = mk3(1, 1, 1), P = mk3(0, 0, 0);
    int level = 0;
    uint32_t pix = 0, occl_mask = 0;
    uint32_t n_secondary = 0, n_shadow = 0;
    WorkCount wc, wcs;
    wc.nodes = wc.tris = wcs.nodes = wcs.tris = 0;
    hit.t = nh.t = RT_FLT_MAX; hit.prim = nh.prim = RT_MISS; hit.beta = hit.gamma = nh.beta = nh.gamma = 0.0f;

    bool spawned = false;                     // the current hit's children are parked already (a.share)
    auto start_nearest = [&]() {              // (ro, rd) is the lane's current ray
        spawned = false;
        bool finite = rd.x == rd.x && rd.y == rd.y && rd.z == rd.z;
        r = prep_ray(ro, rd);
        hit.t = RT_FLT_MAX; hit.prim = RT_MISS; hit.beta = hit.gamma = 0.0f;
        found = false;
        phase = -1;
        stk.clear();
        node = (use_bvh && finite) ? 0 : RT_DONE;
        active = true;
        n_secondary++;
    };
    auto start_shadow = [&](int li) {
        f3 toL = mk3(__ldg(a.s.lights + 2 * li)) - P;
        f3 sd = normalize(toL);
        r = prep_ray(fma3(toL, 0.01f, P), sd);
        hit.t = RT_FLT_MAX; hit.prim = RT_MISS;
        found = false;
        stk.clear();
        node = (use_bvh && sd.x == sd.x && sd.y == sd.y && sd.z == sd.z) ? 0 : RT_DONE;
        n_shadow++;
    };
    auto park = [&](const ShadeChild& ch) {   // a child of the lane's current ray (pix, w) waits on the lane's stack
        if (psp < RT_PATH_STACK) {
            f3 cw = w * ch.w;
            pend[3 * psp] = make_float4(ch.o.x, ch.o.y, ch.o.z, __uint_as_float(pix));
            pend[3 * psp + 1] = make_float4(ch.d.x, ch.d.y, ch.d.z, __int_as_float(ch.level));
            pend[3 * psp + 2] = make_float4(cw.x, cw.y, cw.z, 0.0f);
            psp++;
        } else {
            pend_overflow = true;
        }
    };
    auto next_ray = [&]() {                   // current ray is finished: continue with a parked one or go idle
        if (psp > 0) {
            psp--;
            float4 p0 = pend[3 * psp], p1 = pend[3 * psp + 1], p2 = pend[3 * psp + 2];
            ro = mk3(p0); pix = __float_as_uint(p0.w);
            rd = mk3(p1); level = __float_as_int(p1.w);
            w = mk3(p2);
            start_nearest();
        } else {
            active = false;
        }
    };

    for (;;) {
        uint32_t need = __ballot_sync(FULL, !active);
        if (!exhausted && (__popc(need) >= a.refill_min || need == FULL)) {
            int cnt = __popc(need), leader = __ffs(need) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(a.cursor, (uint32_t)cnt);
            base = __shfl_sync(FULL, base, leader);
            if (base + (uint32_t)cnt >= n) exhausted = true;
            uint32_t my = base + __popc(need & lt);
            if (!active && my < n) {
                float4 q0 = a.q.o_pix[my], q1 = a.q.d_lvl[my], q2 = a.q.w[my];
                ro = mk3(q0); pix = __float_as_uint(q0.w);
                rd = mk3(q1); level = __float_as_int(q1.w);
                w = mk3(q2);
                psp = 0;
                start_nearest();
            }
        }
        // ---- parked rays are shared inside the warp: the k-th idle lane takes the OLDEST parked ray (the root of the
        // largest pending subtree) of the k-th lane that has one, so that no lane walks a whole dielectric subtree
        // alone while its neighbours have nothing to do.  Shuffles only: no queue, no atomics, no polling; pixel sums
        // are order-independent (32.32 fixed point), so frames do not change.
        if (a.share) {
            const uint32_t idle = __ballot_sync(FULL, !active);
            const uint32_t donors = __ballot_sync(FULL, active && psp > 0);
            if (idle && donors) {
                const int pairs = min(__popc(idle), __popc(donors));
                const int my_idle = __popc(idle & lt);
                const bool give = active && psp > 0 && __popc(donors & lt) < pairs;
                const bool take = !active && my_idle < pairs;
                float4 g0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), g1 = g0, g2 = g0;
                if (give) {
                    g0 = pend[0]; g1 = pend[1]; g2 = pend[2];
                    psp--;
                    if (psp > 0) { pend[0] = pend[3 * psp]; pend[1] = pend[3 * psp + 1]; pend[2] = pend[3 * psp + 2]; }
                }
                const int src = take ? (int)__fns(donors, 0, my_idle + 1) : lane;
                g0.x = __shfl_sync(FULL, g0.x, src); g0.y = __shfl_sync(FULL, g0.y, src); g0.z = __shfl_sync(FULL, g0.z, src);
                g0.w = __shfl_sync(FULL, g0.w, src);
                g1.x = __shfl_sync(FULL, g1.x, src); g1.y = __shfl_sync(FULL, g1.y, src); g1.z = __shfl_sync(FULL, g1.z, src);
                g1.w = __shfl_sync(FULL, g1.w, src);
                g2.x = __shfl_sync(FULL, g2.x, src); g2.y = __shfl_sync(FULL, g2.y, src); g2.z = __shfl_sync(FULL, g2.z, src);
                if (take) {
                    ro = mk3(g0); pix = __float_as_uint(g0.w);
                    rd = mk3(g1); level = __float_as_int(g1.w);
                    w = mk3(g2);
                    psp = 0;
                    start_nearest();
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }
        const bool any = phase >= 0;
        WorkCount* wcp = COUNT ? (any ? &wcs : &wc) : nullptr;
        if (active) {
            for (int it = 0; it < burst && node != RT_DONE; it++)
                node = WIDE ? bvh4_step_unified(a.s, r, hit, node, stk, any, found, &overflow, wcp)
                            : bvh_step_unified(a.s, r, hit, node, stk, any, found, &overflow, wcp);
        }
        if (active && node == RT_DONE) {
            bool finite = r.d.x == r.d.x && r.d.y == r.d.y && r.d.z == r.d.z;
            if (finite && !(any && found)) {
                if (a.brute) found |= any ? brute_walk<true>(a.s, r, hit, wcp) : brute_walk<false>(a.s, r, hit, wcp);
                for (int k = 0; k < a.s.n_analytic && !(any && found); k++) {
                    const AnalyticPrim p = a.s.analytic[k];
                    if (COUNT) wcp->tris++;
                    if (analytic_test(p, r.o, r.d, hit.t, hit.beta, hit.gamma)) {
                        hit.prim = rt_analytic_code(k);
                        found = true;
                    }
                }
            }
            bool shade_now = false;
            if (any) {
                occl_mask |= (found ? 1u : 0u) << phase;
                phase++;
                if (phase < a.s.n_lights) start_shadow(phase);
                else shade_now = true;
            } else if (!found) {
                accumulate<false>(a.accum, pix, w * bg);           // world.cpp:110
                next_ray();
            } else {
                nh = hit;
                occl_mask = 0;
                shade_now = true;
                if (a.s.n_lights > 0) {
                    uint32_t mat = hit.prim >= 0 ? __float_as_uint(__ldg(a.s.tris + 3 * (size_t)hit.prim + 1).w)
                                                 : a.s.analytic[rt_analytic_index(hit.prim)].material;
                    float4 m1 = __ldg(a.s.materials + 3 * mat + 1);
                    if (!(m1.z > 0.0f && m1.w > 0.0f)) {           // dielectrics discard the local colour
                        P = fma3(rd, hit.t, ro);
                        if (a.share) {
                            // The mirror child needs the hit, not its shading: it is parked NOW, so that an idle lane
                            // can walk it while this lane walks the hit's shadow rays (a path's chain of dependent
                            // traversals shrinks from depth x (1 + lights) to depth + lights).
                            ShadeOut eo;
                            auto lit = [](f3, f3) -> bool { return false; };
                            shade_hit(a.s, ro, rd, level, nh, a.max_depth, lit, eo);
                            for (int k = 0; k < eo.n_children; k++) park(eo.child[k]);
                            spawned = true;
                        }
                        phase = 0;
                        start_shadow(0);
                        shade_now = false;
                    }
                }
            }
            if (shade_now) {
                ShadeOut out;
                uint32_t li = 0;
                auto occluded = [&](f3, f3) -> bool { bool o2 = ((occl_mask >> li) & 1u) != 0; li++; return o2; };
                shade_hit(a.s, ro, rd, level, nh, a.max_depth, occluded, out);
                accumulate<false>(a.accum, pix, w * (out.local + out.bg_weight * bg));
                if (spawned) out.n_children = 0;                   // its children were parked when the hit was found
                if (out.n_children == 2) park(out.child[1]);       // park the second child
                if (out.n_children >= 1) {
                    const ShadeChild& c0 = out.child[0];
                    ro = c0.o; rd = c0.d; w = w * c0.w; level = c0.level;
                    start_nearest();
                } else {
                    next_ray();
                }
            }
        }
    }
    uint32_t s1 = warp_sum(n_secondary), s2 = warp_sum(n_shadow);
    if (lane == 0) {
        if (s1) atomicAdd(&a.fc->rays_secondary, (unsigned long long)s1);
        if (s2) atomicAdd(&a.fc->rays_shadow, (unsigned long long)s2);
    }
    if (COUNT) {
        uint32_t nn = warp_sum(wc.nodes), nt = warp_sum(wc.tris), sn = warp_sum(wcs.nodes), st = warp_sum(wcs.tris);
        if (lane == 0) {
            atomicAdd(&a.fc->node_visits[0], (unsigned long long)nn);
            atomicAdd(&a.fc->tri_tests[0], (unsigned long long)nt);
            atomicAdd(&a.fc->node_visits[1], (unsigned long long)sn);
            atomicAdd(&a.fc->tri_tests[1], (unsigned long long)st);
        }
    }
    uint32_t fl = (__any_sync(FULL, overflow) ? 2u : 0u) | (__any_sync(FULL, pend_overflow) ? 8u : 0u);
    if (fl && lane == 0) atomicOr(a.sticky, fl);
}

// Heavy tiles first: reorders this rank's tile list by the cost measured in the frame that just ended
// (slowest 32-pixel batch of the tile, descending), so that the LAST batches the persistent warps pick
// up are cheap ones and the kernel does not end on a long tail of expensive batches.  One CTA, bitonic sort of <= 4096 keys
// (cost << 32 | tile id) in shared memory.  Only the order of work changes, never a pixel.
#define RT_SORT_TILES_MAX 16384
#define RT_SORT_CLASSES 256
// Heavy tiles first: a STABLE counting sort of the owned tiles by cost class, descending; tiles of one class keep their
// canonical (ascending tile id) order, which is the cache-friendly one.  One CTA, two passes of n / 1024 steps, ~5 us
// for 2 k tiles (the bitonic sort this replaces took ~50 us and was limited to 4096 tiles).  Launch + run still cost a
// 1/8 share 9 us, and a fresher order does not help a moving camera (profiles/r2_tuning.md section 14), so the order
// is renewed every RT_TILE_SORT_EVERY-th frame (8).
//   cost[o]      slowest batch of the tile with canonical ordinal o in the frame just rendered (cleared here)
//   class        255 - (distance in cost buckets from the most expensive tile), clamped at 0; a bucket is the exponent
//                and the top sub_bits mantissa bits of the cost
//   ids0[o]      the canonical list;  tile_ids[p] / tile_ord[p]: tile and ordinal at position p of the new order
__global__ void __launch_bounds__(1024) k_sort_tiles(uint32_t* __restrict__ tile_ids, uint32_t* __restrict__ tile_ord,
                                                     const uint32_t* __restrict__ ids0, uint32_t* __restrict__ cost, uint32_t n,
                                                     int sub_bits) {
    __shared__ uint8_t cls[RT_SORT_TILES_MAX];
    __shared__ uint16_t hist[32][RT_SORT_CLASSES];      // per warp: tiles of each class in the warp's range, then offsets
    __shared__ uint32_t class_base[RT_SORT_CLASSES];
    __shared__ uint32_t s_max;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto bucket = [&](uint32_t c) -> uint32_t {
        if (!c) return 0u;
        int e = 31 - __clz((int)c);
        return ((uint32_t)e << sub_bits) + (e >= sub_bits ? (c >> (e - sub_bits)) & ((1u << sub_bits) - 1u) : 0u) + 1u;
    };
    if (threadIdx.x == 0) s_max = 0;
    for (uint32_t i = threadIdx.x; i < 32 * RT_SORT_CLASSES; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    uint32_t bm = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) bm = max(bm, bucket(cost[i]));
    for (int o = 16; o; o >>= 1) bm = max(bm, __shfl_xor_sync(0xffffffffu, bm, o));
    if (lane == 0) atomicMax(&s_max, bm);
    __syncthreads();
    bm = s_max;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        uint32_t b = bucket(cost[i]);
        uint32_t d = bm - b;
        cls[i] = (uint8_t)(b == 0 || d > 255u ? 0u : 255u - d);
        cost[i] = 0;
    }
    __syncthreads();
    // every warp owns a contiguous range of ordinals (a multiple of 32 long)
    const uint32_t per_warp = ((n + 1023u) / 1024u) * 32u;
    const uint32_t lo = warp * per_warp, hi = min(n, lo + per_warp);
    for (uint32_t g = lo; g < hi; g += 32) {
        uint32_t o = g + lane;
        bool in = o < hi;
        uint32_t c = in ? cls[o] : 0xffffffffu;
        uint32_t same = __match_any_sync(0xffffffffu, c);
        if (in && lane == __ffs(same) - 1) hist[warp][c] += (uint16_t)__popc(same);
        __syncwarp();
    }
    __syncthreads();
    // offsets: classes descending, inside a class warps ascending
    if (threadIdx.x < RT_SORT_CLASSES) {
        uint32_t c = threadIdx.x, run = 0;
        for (int w = 0; w < 32; w++) { uint32_t h = hist[w][c]; hist[w][c] = (uint16_t)run; run += h; }
        class_base[c] = run;                            // total of the class for now
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int c = RT_SORT_CLASSES - 1; c >= 0; c--) { uint32_t t = class_base[c]; class_base[c] = run; run += t; }
    }
    __syncthreads();
    for (uint32_t g = lo; g < hi; g += 32) {
        uint32_t o = g + lane;
        bool in = o < hi;
        uint32_t c = in ? cls[o] : 0xffffffffu;
        uint32_t same = __match_any_sync(0xffffffffu, c);
        if (in) {
            uint32_t pos = class_base[c] + hist[warp][c] + (uint32_t)__popc(same & ((1u << lane) - 1u));
            tile_ids[pos] = ids0[o];
            tile_ord[pos] = o;
        }
        __syncwarp();
        if (in && lane == __ffs(same) - 1) hist[warp][c] += (uint16_t)__popc(same);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_resolve(FrameDev f, const long long* __restrict__ accum,
                                                uint8_t* __restrict__ out, int packed,
                                                const FrameCounters* __restrict__ fc) {
    uint32_t quads_per_row = (uint32_t)f.tile_w >> 2;
    uint32_t quads_per_tile = quads_per_row * (uint32_t)f.tile_h;
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t static_quads = f.n_tiles_owned * quads_per_tile;
    uint32_t tl = 0, tile, lp0;
    int y, x0;
    if (q < static_quads) {
        tl = q / quads_per_tile;
        uint32_t r = q % quads_per_tile;
        y = (int)(r / quads_per_row); x0 = (int)(r % quads_per_row) * 4;
        tile = __ldg(f.tile_ids + tl);
        lp0 = pixel_in_tile_to_local(f, tl, x0, y);
    } else {
        // blocks claimed from the shared pool: 8 quads (4 rows x 2) per 8x4 block
        uint32_t sq = q - static_quads, slot = sq >> 3, qq = sq & 7;
        if (f.steal_cursor == nullptr || slot >= fc->stolen_blocks) return;
        uint32_t pb = f.stolen_map[slot], bpt = (uint32_t)f.tile_pix >> 5;
        tile = __ldg(f.pool_ids + pb / bpt);
        uint32_t blk = pb % bpt;
        int bpr = f.tile_w >> 3, row = (int)(qq >> 1), xq = (int)(qq & 1) * 4;
        x0 = (int)(blk % (uint32_t)bpr) * 8 + xq;
        y = (int)(blk / (uint32_t)bpr) * 4 + row;
        lp0 = f.n_local_pix + slot * 32u + (uint32_t)(row * 8 + xq);
    }
    int tx = (int)(tile % (uint32_t)f.tiles_x), ty = (int)(tile / (uint32_t)f.tiles_x);
    int j = ty * f.tile_h + y, i0 = tx * f.tile_w + x0;
    if (j >= f.H || i0 >= f.W) return;
    uint32_t px[12];
    int npx = min(4, f.W - i0);
    for (int k = 0; k < 4; k++) {
        const long long* a = accum + 3 * (size_t)(lp0 + k);
        bool in = k < npx;
        px[3 * k + 0] = in ? to_u8(a[0]) : 0u;
        px[3 * k + 1] = in ? to_u8(a[1]) : 0u;
        px[3 * k + 2] = in ? to_u8(a[2]) : 0u;
    }
    size_t byte0 = pixel_byte_offset(f, i0, j, packed);
    if (npx == 4 && (byte0 & 3) == 0) {
        uint32_t* o32 = reinterpret_cast<uint32_t*>(out + byte0);
        o32[0] = px[0] | px[1] << 8 | px[2] << 16 | px[3] << 24;
        o32[1] = px[4] | px[5] << 8 | px[6] << 16 | px[7] << 24;
        o32[2] = px[8] | px[9] << 8 | px[10] << 16 | px[11] << 24;
    } else {
        for (int k = 0; k < 3 * npx; k++) out[byte0 + k] = (uint8_t)px[k];
    }
}

__global__ void k_resolve_float(const long long* __restrict__ accum, uint32_t n3, float* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) out[i] = (float)((double)accum[i] * (1.0 / 4294967296.0));
}

// Tiles of rank `src` (packed) -> full frame.
__global__ void __launch_bounds__(256) k_assemble(const uint8_t* __restrict__ packed, uint8_t* __restrict__ frame,
                                                 int W, int H, int tile_w, int tile_h, int tiles_x,
                                                 uint32_t tiles_total, int src, int world) {
    uint32_t quads_per_row = (uint32_t)tile_w >> 2, quads_per_tile = quads_per_row * (uint32_t)tile_h;
    uint32_t n_owned = tiles_total > (uint32_t)src ? (tiles_total - (uint32_t)src + (uint32_t)world - 1) / (uint32_t)world : 0u;
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_owned * quads_per_tile) return;
    uint32_t tl = q / quads_per_tile, r = q % quads_per_tile;
    uint32_t tile = (uint32_t)src + tl * (uint32_t)world;
    int y = (int)(r / quads_per_row), x0 = (int)(r % quads_per_row) * 4;
    int tx = (int)(tile % (uint32_t)tiles_x), ty = (int)(tile / (uint32_t)tiles_x);
    int j = ty * tile_h + y, i0 = tx * tile_w + x0;
    if (j >= H || i0 >= W) return;
    int npx = min(4, W - i0);
    size_t src0 = ((size_t)tl * tile_w * tile_h + (size_t)y * tile_w + x0) * 3;
    size_t dst0 = ((size_t)i0 + (size_t)j * W) * 3;
    if (npx == 4 && (dst0 & 3) == 0 && (src0 & 3) == 0) {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(packed + src0);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(frame + dst0);
        d32[0] = s32[0]; d32[1] = s32[1]; d32[2] = s32[2];
    } else {
        for (int k = 0; k < 3 * npx; k++) frame[dst0 + k] = packed[src0 + k];
    }
}

// ---- frame-completion handshake between ranks through peer memory (NVLink), no collective library:
// sync[0..63] arrival counters (slot = frame % 64), sync[64] = number of frames rank 0 has consumed.
//   phase 0 (before a rank writes frame k into rank 0's buffer): rank 0 publishes consumed = k — everything
//   it enqueued on its stream for frame k-1 (a download, say) has finished by then; ranks > 0 wait for it
//   phase 1 (after the write): ranks > 0 add 1 to slot k; rank 0 waits for world-1 arrivals and clears the
//   slot that comes into use 32 frames later
// The spinning kernels are single threads on DIFFERENT GPUs and give up after ~2 s (sticky bit 4).
__global__ void k_peer_wait_consumed(volatile uint32_t* sync, uint32_t frame, uint32_t* sticky) {
    long long t0 = clock64();
    while (sync[64] < frame) {
        if (clock64() - t0 > 4000000000ll) { atomicOr(sticky, 4u); break; }
        __nanosleep(200);
    }
    __threadfence_system();
}
__global__ void k_peer_arrive(uint32_t* sync, uint32_t frame) {
    __threadfence_system();
    atomicAdd_system(sync + (frame % 64u), 1u);
}
__global__ void k_peer_wait_all(volatile uint32_t* sync, uint32_t frame, uint32_t expect, uint32_t* sticky) {
    long long t0 = clock64();
    while (sync[frame % 64u] < expect) {
        if (clock64() - t0 > 4000000000ll) { atomicOr(sticky, 4u); break; }
        __nanosleep(100);
    }
    __threadfence_system();
    sync[(frame + 32u) % 64u] = 0u;
    __threadfence_system();
}
__global__ void k_peer_publish_consumed(volatile uint32_t* sync, uint32_t frame) {
    if (sync[64] < frame) sync[64] = frame;
    __threadfence_system();
}

// Symmetric barrier over the same buffer (sync[80] counts arrivals of all epochs): used to line ranks up.
__global__ void k_peer_barrier(uint32_t* sync, uint32_t target, uint32_t* sticky) {
    __threadfence_system();
    atomicAdd_system(sync + 80, 1u);
    volatile uint32_t* v = sync + 80;
    long long t0 = clock64();
    while (*v < target) {
        if (clock64() - t0 > 4000000000ll) { atomicOr(sticky, 4u); break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

// Same scatter with 16-byte transactions (what NVLink peer stores want): one thread per 16 B of a tile row.
// Requires rows of the frame and of the tile to be 16-byte multiples; partial edge rows fall back to bytes.
__global__ void __launch_bounds__(256) k_assemble16(const uint8_t* __restrict__ packed, uint8_t* __restrict__ frame,
                                                   int W, int H, int tile_w, int tile_h, int tiles_x,
                                                   uint32_t tiles_total, int src, int world) {
    const uint32_t chunks_per_row = (uint32_t)(tile_w * 3) >> 4, chunks_per_tile = chunks_per_row * (uint32_t)tile_h;
    uint32_t n_owned = tiles_total > (uint32_t)src ? (tiles_total - (uint32_t)src + (uint32_t)world - 1) / (uint32_t)world : 0u;
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_owned * chunks_per_tile) return;
    uint32_t tl = q / chunks_per_tile, r = q % chunks_per_tile;
    uint32_t tile = (uint32_t)src + tl * (uint32_t)world;
    int y = (int)(r / chunks_per_row), cb = (int)(r % chunks_per_row) * 16;     // byte offset inside the tile row
    int tx = (int)(tile % (uint32_t)tiles_x), ty = (int)(tile / (uint32_t)tiles_x);
    int j = ty * tile_h + y;
    if (j >= H) return;
    int row_bytes = min(tile_w, W - tx * tile_w) * 3;                           // valid bytes of this tile row
    if (cb >= row_bytes) return;
    size_t src0 = ((size_t)tl * tile_w * tile_h + (size_t)y * tile_w) * 3 + cb;
    size_t dst0 = ((size_t)tx * tile_w + (size_t)j * W) * 3 + cb;
    if (cb + 16 <= row_bytes) {
        *reinterpret_cast<uint4*>(frame + dst0) = *reinterpret_cast<const uint4*>(packed + src0);
    } else {
        for (int k = 0; k < row_bytes - cb; k++) frame[dst0 + k] = packed[src0 + k];
    }
}

CamDev make_cam(rt_ctx* ctx, const rt_camera* c) {
    CamDev d;
    for (int k = 0; k < 3; k++) { d.pos[k] = c->pos[k]; d.u[k] = c->u[k]; d.v[k] = c->v[k]; d.w[k] = c->w[k]; }
    d.focal = c->focal_distance;
    d.aspect = c->aspect;
    d.W = c->width;
    d.H = c->height;
    const int W = c->width > 0 ? c->width : 0, H = c->height > 0 ? c->height : 0;
    if (ctx->cam_tab_W != W || ctx->cam_tab_H != H || ctx->cam_tab_aspect != c->aspect || !ctx->d_cam_tab.p) {
        ctx->d_cam_tab.reserve((size_t)W + H + 1);
        if (W + H > 0) {
            k_camera_tables<<<(W + H + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_cam_tab.p, ctx->d_cam_tab.p + W, W, H, c->aspect);
            RT_CUDA(cudaGetLastError());
        }
        ctx->cam_tab_W = W; ctx->cam_tab_H = H; ctx->cam_tab_aspect = c->aspect;
    }
    d.xw = ctx->d_cam_tab.p;
    d.yw = ctx->d_cam_tab.p + W;
    return d;
}

template <typename K>
int persistent_blocks(K kernel, int tpb, int sm_count) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, tpb, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    return per_sm * sm_count;
}

void setup_layout(rt_ctx* c, const rt_camera* cam, const rt_render_params* p) {
    TileLayout L;
    L.width = cam->width; L.height = cam->height;
    L.tile_w = p->tile_w > 0 ? p->tile_w : 64;
    L.tile_h = p->tile_h > 0 ? p->tile_h : 32;
    L.rank = p->world_size > 1 ? p->rank : 0;
    L.world = p->world_size > 1 ? p->world_size : 1;
    L.pool_div = (L.world > 1 && p->steal_pool_div > 0) ? p->steal_pool_div : 0;
    if (L.tile_w % 8 || L.tile_h % 4) throw RtError{RT_ERR_INVALID_ARGUMENT, "tile_w must be a multiple of 8 and tile_h of 4"};
    if (L.rank < 0 || L.rank >= L.world) throw RtError{RT_ERR_INVALID_ARGUMENT, "rank outside [0, world_size)"};
    if (L.pool_div && (p->flags & RT_FLAG_PACKED_TILES))
        throw RtError{RT_ERR_INVALID_ARGUMENT, "tile stealing needs a frame-layout output, not RT_FLAG_PACKED_TILES"};
    L.tiles_x = (L.width + L.tile_w - 1) / L.tile_w;
    L.tiles_y = (L.height + L.tile_h - 1) / L.tile_h;
    if (L == c->layout && c->d_tile_ids.p) return;
    // tile t belongs to group t / world; every pool_div-th group is the shared pool, the other groups are
    // dealt out round robin (tile t -> rank t % world)
    std::vector<uint32_t> ids, pool;
    uint32_t total = (uint32_t)L.tiles_x * (uint32_t)L.tiles_y;
    for (uint32_t t = 0; t < total; t++) {
        uint32_t g = t / (uint32_t)L.world;
        if (L.pool_div && g % (uint32_t)L.pool_div == (uint32_t)L.pool_div - 1) pool.push_back(t);
        else if (t % (uint32_t)L.world == (uint32_t)L.rank) ids.push_back(t);
    }
    L.n_tiles_owned = (uint32_t)ids.size();
    L.n_pool_tiles = (uint32_t)pool.size();
    c->d_tile_ids.reserve(ids.size() ? ids.size() : 1);
    c->d_tile_ids0.reserve(ids.size() ? ids.size() : 1);
    c->d_tile_ord.reserve(ids.size() ? ids.size() : 1);
    c->d_pool_ids.reserve(pool.size() ? pool.size() : 1);
    uint32_t bpt = (uint32_t)(L.tile_w * L.tile_h) / 32u;
    c->d_stolen_map.reserve(pool.size() ? pool.size() * bpt : 1);
    if (!ids.empty()) {
        std::vector<uint32_t> ord(ids.size());
        for (size_t k = 0; k < ord.size(); k++) ord[k] = (uint32_t)k;
        RT_CUDA(cudaMemcpyAsync(c->d_tile_ids.p, ids.data(), ids.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        RT_CUDA(cudaMemcpyAsync(c->d_tile_ids0.p, ids.data(), ids.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        RT_CUDA(cudaMemcpyAsync(c->d_tile_ord.p, ord.data(), ord.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        RT_CUDA(cudaStreamSynchronize(c->stream));      // `ord` and `ids` are pageable host memory
    }
    if (!pool.empty())
        RT_CUDA(cudaMemcpyAsync(c->d_pool_ids.p, pool.data(), pool.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    c->layout = L;
    c->tile_cost_valid = false;
    c->frames_in_layout = 0;
}

FrameDev frame_dev(rt_ctx* c, const rt_render_params* p) {
    const TileLayout& L = c->layout;
    FrameDev f;
    f.W = L.width; f.H = L.height; f.tile_w = L.tile_w; f.tile_h = L.tile_h; f.tiles_x = L.tiles_x;
    f.tile_pix = L.tile_w * L.tile_h;
    f.world = L.world;
    f.n_tiles_owned = L.n_tiles_owned;
    f.n_local_pix = L.n_tiles_owned * (uint32_t)f.tile_pix;
    f.tile_ids = c->d_tile_ids.p;
    f.tile_ord = c->d_tile_ord.p;
    f.pool_ids = c->d_pool_ids.p;
    f.n_pool_blocks = L.n_pool_tiles * ((uint32_t)f.tile_pix / 32u);
    f.stolen_map = c->d_stolen_map.p;
    f.tile_cost = nullptr;
    f.n_wide_pix = 0;
    f.wide_after_bursts = 0;
    if (c->tile_feedback && L.n_tiles_owned > 1 && L.n_tiles_owned <= RT_SORT_TILES_MAX && c->refill_primary == 32 &&
        (!c->fuse_shadow || c->refill_primary_fused == 32)) {
        c->d_tile_cost.reserve(L.n_tiles_owned);
        if (!c->tile_cost_valid) {
            RT_CUDA(cudaMemsetAsync(c->d_tile_cost.p, 0, L.n_tiles_owned * sizeof(uint32_t), c->stream));
            c->tile_cost_valid = true;
        }
        f.tile_cost = c->d_tile_cost.p;
    }
    f.steal_cursor = nullptr;
    if (L.pool_div && L.n_pool_tiles) {
        uint32_t slot = p->frame_index % 64u;
        if (p->steal_cursor) {
            f.steal_cursor = (uint32_t*)p->steal_cursor + slot;
            // rank 0 owns the cursor array: clear the slot that comes into use 32 frames from now (every
            // rank has passed this frame's completion barrier long before it gets there)
            if (L.rank == 0)
                RT_CUDA(cudaMemsetAsync((uint32_t*)p->steal_cursor + (p->frame_index + 32u) % 64u, 0, sizeof(uint32_t), c->stream));
        } else {
            // no shared cursor: logical ranks rendered one after the other by this context
            c->d_local_cursor.reserve(64);
            if (!c->local_cursor_valid || c->local_cursor_frame != p->frame_index) {
                RT_CUDA(cudaMemsetAsync(c->d_local_cursor.p, 0, 64 * sizeof(uint32_t), c->stream));
                c->local_cursor_valid = true;
                c->local_cursor_frame = p->frame_index;
            }
            f.steal_cursor = c->d_local_cursor.p + slot;
        }
    }
    return f;
}

void ensure_queues(rt_ctx* c, size_t n_rays0, bool need_bounce) {
    size_t hits_cap = n_rays0 ? n_rays0 : 1;
    size_t cap = 0;
    if (need_bounce) cap = (c->has_dielectric ? 2 : 1) * hits_cap;
    if (cap > hits_cap) hits_cap = cap;
    c->d_hits.reserve(hits_cap);
    c->d_hitq.reserve(hits_cap);
    size_t nl = c->scene.n_lights > 0 ? (size_t)c->scene.n_lights : 1;
    c->d_occl.reserve(hits_cap * nl);
    if (cap) {
        for (int b = 0; b < 2; b++)
            for (int k = 0; k < 3; k++) c->d_q[b][k].reserve(cap);
    }
    c->queue_cap = cap;
}

RayQueue queue_of(rt_ctx* c, int b) {
    RayQueue q;
    q.o_pix = c->d_q[b][0].p; q.d_lvl = c->d_q[b][1].p; q.w = c->d_q[b][2].p;
    return q;
}

// Does the primary wave shade its hits inside the traversal kernel?  (RT_FUSE_SHADE: 0 never, 1 when the
// kernel also pushes finished tiles into a remote frame, 2 always.)
bool wave0_shades_inline(const rt_ctx* c, bool pushing) {
    return c->fuse_shadow && c->scene.n_lights > 0 && c->scene.n_lights <= 8 && c->refill_primary_fused == 32 &&
           (c->fuse_shade == 2 || (c->fuse_shade == 1 && pushing));
}
template <int MODE, bool FUSE, bool SHADE = false>
void launch_traverse(rt_ctx* c, const TravArgs& a, bool count) {
    int blocks = MODE == MODE_SHADOW ? c->shadow_blocks : (FUSE ? (SHADE ? c->fused_shade_blocks : c->fused_blocks) : c->trace_blocks);
    if (FUSE && a.s.nodes4 != nullptr && c->wide_bvh == 2) {
        constexpr int W = FUSE ? 1 : 0;   // only the fused kernels have a wide instantiation
        blocks = SHADE ? c->wide_shade_blocks : c->wide_blocks;
        if (count) k_traverse<MODE, true, FUSE, W, SHADE><<<blocks, TRAV_TPB, 0, c->stream>>>(a);
        else k_traverse<MODE, false, FUSE, W, SHADE><<<blocks, TRAV_TPB, 0, c->stream>>>(a);
    } else {
        if (count) k_traverse<MODE, true, FUSE, 0, SHADE><<<blocks, TRAV_TPB, 0, c->stream>>>(a);
        else k_traverse<MODE, false, FUSE, 0, SHADE><<<blocks, TRAV_TPB, 0, c->stream>>>(a);
    }
    RT_CUDA(cudaGetLastError());
}
template <bool PRIMARY>
void launch_shade(rt_ctx* c, const ShadeArgs& a) {
    k_shade<PRIMARY><<<c->shade_blocks, SHADE_TPB, 0, c->stream>>>(a);
    RT_CUDA(cudaGetLastError());
}

// The three kernels of one wave whose rays sit in queue `cur` (or are the primary rays).
template <bool PRIMARY>
uint32_t launch_wave(rt_ctx* c, TravArgs ta, ShadeArgs sa, int slot_in, int slot_out, int cur, bool count, bool shade,
                     cudaEvent_t after_trace, cudaEvent_t after_shadow) {
    uint32_t launches = 0;
    ta.q = queue_of(c, cur);
    ta.wave = c->d_waves.p + slot_in;
    ta.hits = c->d_hits.p; ta.hitq = c->d_hitq.p;
    ta.refill_min = PRIMARY ? c->refill_primary : c->refill_queue;
    ta.loop_style = PRIMARY ? c->loop_primary : c->loop_queue;
    const bool fuse = shade && c->fuse_shadow && c->scene.n_lights > 0 && c->scene.n_lights <= 8;
    // primary wave, whole-batch refill: shading happens inside the traversal kernel too
    const bool fuse_shade = PRIMARY && fuse && wave0_shades_inline(c, ta.remote_out != 0);
    if (fuse_shade) {
        ta.refill_min = 32;
        ta.qout = queue_of(c, cur ^ 1);
        ta.next = c->d_waves.p + slot_out;
        ta.max_depth = sa.max_depth;
        launch_traverse<MODE_PRIMARY, true, PRIMARY>(c, ta, count);     // (PRIMARY is true here)
        launches++;
        if (after_trace) RT_CUDA(cudaEventRecord(after_trace, c->stream));
        if (after_shadow) RT_CUDA(cudaEventRecord(after_shadow, c->stream));
        return launches;
    }
    if (fuse) {
        ta.occl = c->d_occl.p;
        ta.refill_min = PRIMARY ? c->refill_primary_fused : c->refill_queue;
        launch_traverse<PRIMARY ? MODE_PRIMARY : MODE_QUEUE, true>(c, ta, count);
    } else {
        launch_traverse<PRIMARY ? MODE_PRIMARY : MODE_QUEUE, false>(c, ta, count);
    }
    launches++;
    if (after_trace) RT_CUDA(cudaEventRecord(after_trace, c->stream));
    if (!shade) return launches;
    sa.occl_bits = fuse ? 1 : 0;
    if (!fuse && c->scene.n_lights > 0) {
        TravArgs sh = ta;
        sh.s.nodes4 = nullptr;
        sh.hits_in = c->d_hits.p; sh.hitq_in = c->d_hitq.p; sh.occl = c->d_occl.p;
        sh.primary_wave = PRIMARY ? 1u : 0u;
        sh.refill_min = c->refill_shadow;
        sh.loop_style = c->loop_shadow;
        sh.aux_prim = nullptr; sh.aux_t = nullptr;
        sh.warp_times = nullptr;
        launch_traverse<MODE_SHADOW, false>(c, sh, count);
        launches++;
    }
    if (after_shadow) RT_CUDA(cudaEventRecord(after_shadow, c->stream));
    sa.qin = queue_of(c, cur);
    sa.qout = queue_of(c, cur ^ 1);
    sa.wave = c->d_waves.p + slot_in;
    sa.next = c->d_waves.p + slot_out;
    sa.occl = c->d_occl.p;
    launch_shade<PRIMARY>(c, sa);
    launches++;
    return launches;
}

struct WaveResult {
    uint32_t waves = 0, launches = 0, max_queue = 0;
    uint64_t secondary = 0;
};

// All bounce generations in one launch: the rays sit in queue `cur`, their count in wave slot `slot`.
void launch_paths(rt_ctx* c, int slot, int cur, int max_depth, uint32_t brute, bool count) {
    PathArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.s = c->scene;
    pa.q = queue_of(c, cur);
    pa.wave = c->d_waves.p + slot;
    pa.cursor = &c->d_waves.p[slot].fetch_trace;
    pa.accum = c->d_accum.p;
    pa.fc = c->d_frame.p;
    pa.sticky = c->d_sticky.p;
    pa.cap = (uint32_t)c->queue_cap;
    pa.max_depth = max_depth;
    pa.refill_min = c->refill_queue;
    pa.loop_style = c->loop_queue;
    pa.brute = brute;
    pa.share = c->path_share;
    if (pa.s.nodes4) {
        if (count) k_paths<true, true><<<c->path_wide_blocks, TRAV_TPB, 0, c->stream>>>(pa);
        else k_paths<false, true><<<c->path_wide_blocks, TRAV_TPB, 0, c->stream>>>(pa);
    } else {
        if (count) k_paths<true><<<c->path_blocks, TRAV_TPB, 0, c->stream>>>(pa);
        else k_paths<false><<<c->path_blocks, TRAV_TPB, 0, c->stream>>>(pa);
    }
    RT_CUDA(cudaGetLastError());
}

// Runs waves first_wave.. over rays already sitting in queue `cur`.  Mirror-only scenes need no
// host round trip: a ray of wave w has level w, so exactly max_depth waves can be populated and
// they are launched blind (empty ones exit at once).  With dielectrics (level*2, world.cpp:98) the
// population of each wave is read back before it is launched.
WaveResult run_bounce_waves(rt_ctx* c, const TravArgs& ta, const ShadeArgs& sa, int first_wave, int cur, int max_depth,
                            bool count, bool levels_equal_waves) {
    WaveResult r;
    cudaStream_t st = c->stream;
    bool poll = c->has_dielectric || max_depth > 16 || !levels_equal_waves;
    for (int w = first_wave;; w++) {
        int slot_in = w % RT_WAVE_SLOTS, slot_out = (w + 1) % RT_WAVE_SLOTS;
        if (poll) {
            RT_CUDA(cudaMemcpyAsync(&c->h_waves[slot_in], c->d_waves.p + slot_in, sizeof(WaveCounters), cudaMemcpyDeviceToHost, st));
            RT_CUDA(cudaStreamSynchronize(st));
            uint32_t n = c->h_waves[slot_in].n_rays;
            if (n == 0) break;
            if (n > c->queue_cap) n = (uint32_t)c->queue_cap;
            r.secondary += n;
            if (n > r.max_queue) r.max_queue = n;
        } else if (w > max_depth) {
            break;
        }
        if (w + 1 >= RT_WAVE_SLOTS) RT_CUDA(cudaMemsetAsync(c->d_waves.p + slot_out, 0, sizeof(WaveCounters), st));
        r.launches += launch_wave<false>(c, ta, sa, slot_in, slot_out, cur, count, true, nullptr, nullptr);
        r.waves++;
        cur ^= 1;
    }
    return r;
}

}  // namespace

// Can any hit spawn a live child ray?  A mirror child has level + 1 (world.cpp:105), alive from max_depth 1 on; the
// refracted child of a dielectric has level * 2 (world.cpp:98), which stays 0 below a primary ray and is therefore
// alive even at max_depth 0 (the reference's guard is `level > RECURSION_DEPTH`, world.cpp:33).
bool rt_scene_bounces(const rt_ctx* c, int max_depth) {
    return c->has_reflective && (max_depth >= 1 || (c->has_dielectric && max_depth >= 0));
}

// rt_render_push: can the frame's only kernel push its tiles itself?  Bounce-free scenes only: their pixels are
// final the moment the batch is shaded.
bool rt_frame_pushes_inline(const rt_ctx* c, const rt_render_params* p) {
    const bool bounce = rt_scene_bounces(c, p->max_depth);
    return !bounce && (p->flags & RT_FLAG_PACKED_TILES) && !(p->world_size > 1 && p->steal_pool_div > 0) &&
           wave0_shades_inline(c, true);
}


// Can the whole frame run as one k_frame launch?  Bounce-free scenes whose shadow rays ride in the primary lanes.
// RT_FRAME_KERNEL: 0 never, 1 where the frame is pushed to a shared (multi-GPU) frame, 2 always.
bool rt_frame_kernel_ok(const rt_ctx* c, const rt_render_params* p, bool pushing) {
    const bool stealing = p->world_size > 1 && p->steal_pool_div > 0;
    return (c->frame_kernel == 2 || (c->frame_kernel == 1 && pushing)) && !rt_scene_bounces(c, p->max_depth) && c->fuse_shadow &&
           c->scene.n_lights > 0 && c->scene.n_lights <= 8 && !stealing;
}

// Waits for the context's stream and turns the sticky device error word into an exception.
void rt_sync_and_check(rt_ctx* c) {
    cudaStream_t st = c->stream;
    RT_CUDA(cudaMemcpyAsync(c->h_sticky, c->d_sticky.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    uint32_t fl = *c->h_sticky;
    if (fl) {
        RT_CUDA(cudaMemsetAsync(c->d_sticky.p, 0, sizeof(uint32_t), st));
        throw rt_sticky_error(fl);
    }
}

RtError rt_sticky_error(uint32_t fl) {
    if (!fl) return RtError{RT_OK, ""};
    if (fl & 4u) return RtError{RT_ERR_CUDA, "peer frame handshake timed out (a rank did not arrive within ~2 s)"};
    if (fl & 1u) return RtError{RT_ERR_QUEUE_OVERFLOW, "ray queue overflow: a wave spawned more rays than the queue holds"};
    if (fl & 8u) return RtError{RT_ERR_QUEUE_OVERFLOW, "pending-ray stack overflow in k_paths (dielectric chain deeper than RT_PATH_STACK); set RT_PATH_KERNEL=0"};
    return RtError{RT_ERR_QUEUE_OVERFLOW, "traversal stack overflow (BVH deeper than RT_STACK_SIZE)"};
}

void rt_render_init(rt_ctx* c) {
    auto lo = [](int a, int b) { return a < b ? a : b; };
    c->trace_blocks = lo(persistent_blocks(k_traverse<MODE_PRIMARY, false, false>, TRAV_TPB, c->sm_count),
                         persistent_blocks(k_traverse<MODE_QUEUE, false, false>, TRAV_TPB, c->sm_count));
    c->fused_blocks = lo(persistent_blocks(k_traverse<MODE_PRIMARY, false, true>, TRAV_TPB, c->sm_count),
                         persistent_blocks(k_traverse<MODE_QUEUE, false, true>, TRAV_TPB, c->sm_count));
    c->shadow_blocks = persistent_blocks(k_traverse<MODE_SHADOW, false, false>, TRAV_TPB, c->sm_count);
    c->path_blocks = persistent_blocks(k_paths<false>, TRAV_TPB, c->sm_count);
    c->path_wide_blocks = persistent_blocks(k_paths<false, true>, TRAV_TPB, c->sm_count);
    c->wide_blocks = lo(persistent_blocks(k_traverse<MODE_PRIMARY, false, true, 1>, TRAV_TPB, c->sm_count),
                        persistent_blocks(k_traverse<MODE_QUEUE, false, true, 1>, TRAV_TPB, c->sm_count));
    c->fused_shade_blocks = persistent_blocks(k_traverse<MODE_PRIMARY, false, true, 0, true>, TRAV_TPB, c->sm_count);
    c->wide_shade_blocks = persistent_blocks(k_traverse<MODE_PRIMARY, false, true, 1, true>, TRAV_TPB, c->sm_count);
    c->shade_blocks = lo(persistent_blocks(k_shade<true>, SHADE_TPB, c->sm_count),
                         persistent_blocks(k_shade<false>, SHADE_TPB, c->sm_count));
    c->frame_blocks = persistent_blocks(k_frame<false>, TRAV_TPB, c->sm_count);
    c->frame_push_blocks = persistent_blocks(k_frame_push<false>, TRAV_TPB, c->sm_count);
    c->frame_blocks_dense = persistent_blocks(k_frame<false, RT_DENSE_MIN_BLOCKS>, TRAV_TPB, c->sm_count);
    c->frame_push_blocks_dense = persistent_blocks(k_frame_push<false, RT_DENSE_MIN_BLOCKS>, TRAV_TPB, c->sm_count);
    c->d_fsync.reserve(4);
    RT_CUDA(cudaMemset(c->d_fsync.p, 0, 4 * sizeof(uint32_t)));
    c->d_fk.reserve(2);
    RT_CUDA(cudaMemset(c->d_fk.p, 0, 2 * sizeof(FrameKernelCounters)));
    if (c->blocks_per_sm > 0) {   // RT_BLOCKS_PER_SM: cap the persistent grids (tuning)
        int cap = c->blocks_per_sm * c->sm_count;
        c->trace_blocks = lo(c->trace_blocks, cap);
        c->fused_blocks = lo(c->fused_blocks, cap);
        c->fused_shade_blocks = lo(c->fused_shade_blocks, cap);
        c->frame_blocks = lo(c->frame_blocks, cap);
        c->frame_push_blocks = lo(c->frame_push_blocks, cap);
        c->frame_blocks_dense = lo(c->frame_blocks_dense, cap);
        c->frame_push_blocks_dense = lo(c->frame_push_blocks_dense, cap);
        c->shadow_blocks = lo(c->shadow_blocks, cap);
        c->path_blocks = lo(c->path_blocks, cap);
    }
    c->d_waves.reserve(RT_WAVE_SLOTS);
    c->d_frame.reserve(1);
    RT_CUDA(cudaMallocHost((void**)&c->h_waves, RT_WAVE_SLOTS * sizeof(WaveCounters)));
    RT_CUDA(cudaMallocHost((void**)&c->h_frame, sizeof(FrameCounters)));
    RT_CUDA(cudaMallocHost((void**)&c->h_sticky, sizeof(uint32_t)));
    c->d_sticky.reserve(1);
    RT_CUDA(cudaMemset(c->d_sticky.p, 0, sizeof(uint32_t)));
}

void rt_render_frame(rt_ctx* c, const rt_camera* cam, const rt_render_params* p, void* rgb_dev,
                     const rt_aux_out* aux_dev, rt_frame_stats* stats) {
    cudaStream_t st = c->stream;
    setup_layout(c, cam, p);
    FrameDev f = frame_dev(c, p);
    const bool count = (p->flags & RT_FLAG_COUNT_WORK) != 0;
    const bool bounce = rt_scene_bounces(c, p->max_depth);
    const size_t pix_cap = (size_t)f.n_local_pix + (size_t)f.n_pool_blocks * 32u;   // own tiles + everything stealable
    ensure_queues(c, pix_cap, bounce);
    c->d_accum.reserve(3 * (pix_cap ? pix_cap : 1));

    // no bounces => one contribution per pixel => RGB8 straight from the wave-0 kernels (stolen blocks keep
    // the accumulator path: their packed offset is not defined)
    const bool direct = !bounce && rgb_dev != nullptr && f.steal_cursor == nullptr;
    const bool has_work = f.n_local_pix > 0 || f.steal_cursor != nullptr;
    const bool pushing = c->push.frame != nullptr;
    const bool use_fk = direct && rt_frame_kernel_ok(c, p, pushing) && (has_work || pushing);
    // event records only where somebody reads them: each one is a serialisation point on the stream
    auto mark = [&](int k) { if (stats) RT_CUDA(cudaEventRecord(c->ev[k], st)); };
    mark(0);
    WaveCounters* waves_dev = c->d_waves.p;
    FrameCounters* frame_dev_ctr = c->d_frame.p;
    uint32_t* zero_next = nullptr;
    if (use_fk) {
        // k_frame's counters are double-buffered and cleared by the previous launch: no memset on the stream
        FrameKernelCounters* set = c->d_fk.p + (c->fk_epoch & 1u);
        zero_next = (uint32_t*)(c->d_fk.p + ((c->fk_epoch + 1u) & 1u));
        c->fk_epoch++;
        waves_dev = &set->wave;
        frame_dev_ctr = &set->fc;
    } else {
        RT_CUDA(cudaMemsetAsync(c->d_waves.p, 0, RT_WAVE_SLOTS * sizeof(WaveCounters), st));
        RT_CUDA(cudaMemsetAsync(c->d_frame.p, 0, sizeof(FrameCounters), st));
    }
    c->last_frame_ctr = frame_dev_ctr;

    TravArgs ta;
    memset(&ta, 0, sizeof ta);
    ta.s = c->scene; ta.cam = make_cam(c, cam); ta.f = f;
    ta.accum = c->d_accum.p; ta.fc = frame_dev_ctr; ta.sticky = c->d_sticky.p;
    ta.aux_prim = aux_dev ? aux_dev->prim_id : nullptr;
    ta.aux_t = aux_dev ? aux_dev->t : nullptr;
    ta.brute = (p->flags & RT_FLAG_BRUTE_FORCE) ? 1u : 0u;
    ta.cap = (uint32_t)c->queue_cap;
    ShadeArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.s = c->scene; sa.cam = ta.cam; sa.f = f;
    sa.hits = c->d_hits.p; sa.hitq = c->d_hitq.p; sa.accum = c->d_accum.p; sa.cap = (uint32_t)c->queue_cap;
    sa.sticky = c->d_sticky.p;
    sa.max_depth = p->max_depth;

    if (p->flags & RT_FLAG_WARP_TIMES) {
        int mb = c->trace_blocks;
        for (int b : {c->fused_blocks, c->fused_shade_blocks, c->wide_blocks, c->wide_shade_blocks})
            mb = b > mb ? b : mb;
        c->d_warp_times.reserve(2 * (size_t)mb * (TRAV_TPB / 32));
        ta.warp_times = c->d_warp_times.p;
    }
    if (direct) {
        ta.direct_rgb = sa.direct_rgb = (uint8_t*)rgb_dev;
        ta.direct_packed = sa.direct_packed = (p->flags & RT_FLAG_PACKED_TILES) ? 1 : 0;
    }
    ta.remote_out = (c->remote_output && direct) ? 1 : 0;
    uint32_t launches = 0;
    if (use_fk) {
        // the whole frame in one launch: trace + shadow rays, shade, (push + handshake)
        FrameArgs fa;
        memset(&fa, 0, sizeof fa);
        ta.q = queue_of(c, 0);
        ta.wave = waves_dev;
        ta.hits = c->d_hits.p; ta.hitq = c->d_hitq.p; ta.occl = c->d_occl.p;
        ta.refill_min = c->refill_primary_fused;
        ta.loop_style = c->loop_primary;
        // the tiles at the head of the heavy-tiles-first order walk the 4-wide view (traverse_body<WIDE = 2>): needs a
        // cost-sorted order (from the second frame of a layout on) and whole-batch refill; the view is built on first use
        const bool dense = (long long)f.n_local_pix >= c->dense_min_pixels;
        ta.f.n_wide_pix = 0;
        if (RT_FRAME_WIDE == 2 && !dense && f.tile_cost && c->frames_in_layout >= 1 && ta.refill_min == 32 && !ta.brute &&
            (c->wide_heavy == 2 || (c->wide_heavy == 1 && f.world > 1)) && c->n_bvh >= 1) {
            rt_ensure_nodes4(c);
            ta.s = c->scene;
            ta.f.n_wide_pix = ((f.n_tiles_owned + (uint32_t)c->wide_heavy_div - 1) / (uint32_t)c->wide_heavy_div) * (uint32_t)f.tile_pix;
        }
        // long batches the order did not announce move to the wide view after a few bursts (no order needed)
        ta.f.wide_after_bursts = 0;
        if (RT_FRAME_WIDE == 2 && !dense && c->wide_after_bursts > 0 && ta.refill_min == 32 && !ta.brute && ta.loop_style > 0 &&
            (c->wide_heavy == 2 || (c->wide_heavy == 1 && f.world > 1)) && c->n_bvh >= 1) {
            rt_ensure_nodes4(c);
            ta.s = c->scene;
            ta.f.wide_after_bursts = (uint32_t)c->wide_after_bursts;
        }
        fa.t = ta;
        fa.max_depth = p->max_depth;
        static const bool fk_split = getenv("RT_FK_SPLIT") != nullptr;
        fa.phase1_only = (fk_split && !pushing) ? 1 : 0;
        fa.y.done = c->d_fsync.p;
        fa.y.zero_words = zero_next;
        fa.y.n_zero_words = (uint32_t)(sizeof(FrameKernelCounters) / sizeof(uint32_t));
        // large shares run the 9-CTAs-per-SM build of the kernel, small ones the 7-CTA build (see k_frame)
        const int fk_blocks = dense ? c->frame_blocks_dense : c->frame_blocks;
        const int fkp_blocks = dense ? c->frame_push_blocks_dense : c->frame_push_blocks;
        if (p->flags & RT_FLAG_WARP_TIMES) {
            const size_t nw = 8 * (size_t)(fk_blocks > fkp_blocks ? fk_blocks : fkp_blocks) * (TRAV_TPB / 32);
            c->d_warp_times.reserve(nw);
            RT_CUDA(cudaMemsetAsync(c->d_warp_times.p, 0, nw * sizeof(unsigned long long), st));
            fa.phase_times = c->d_warp_times.p;
            fa.t.warp_times = nullptr;
        }
        if (pushing) {
            fa.y.sync = c->push.world > 1 ? (uint32_t*)c->push.sync : nullptr;
            fa.y.frame = c->push.frame_index; fa.y.rank = (uint32_t)c->push.rank; fa.y.world = (uint32_t)c->push.world;
        }
        if (pushing && c->push_inline) {
            // stores spread over the frame: every finished 8x4 block goes straight into the shared frame
            fa.t.direct_rgb = (uint8_t*)c->push.frame;
            fa.t.direct_packed = 0;
            fa.t.remote_out = 1;
            fa.t.refill_min = 32;
            fa.t.qout = queue_of(c, 1);
            fa.t.next = waves_dev;                 // (bounce-free: nothing is ever appended)
            fa.t.max_depth = p->max_depth;
            c->fsync_target[2] += (uint32_t)fkp_blocks;
            fa.y.target[2] = c->fsync_target[2];
            if (count) {
                if (dense) k_frame_push<true, RT_DENSE_MIN_BLOCKS><<<fkp_blocks, TRAV_TPB, 0, st>>>(fa);
                else k_frame_push<true><<<fkp_blocks, TRAV_TPB, 0, st>>>(fa);
            } else {
                if (dense) k_frame_push<false, RT_DENSE_MIN_BLOCKS><<<fkp_blocks, TRAV_TPB, 0, st>>>(fa);
                else k_frame_push<false><<<fkp_blocks, TRAV_TPB, 0, st>>>(fa);
            }
        } else {
            if (pushing) {
                fa.push.packed = (const uint8_t*)rgb_dev;
                fa.push.frame = (uint8_t*)c->push.frame;
                fa.push.tiles_total = (uint32_t)c->layout.tiles_x * (uint32_t)c->layout.tiles_y;
                fa.push.wide16 = (f.W * 3) % 16 == 0 && (f.tile_w * 3) % 16 == 0 && ((uintptr_t)rgb_dev & 15) == 0 &&
                                 ((uintptr_t)c->push.frame & 15) == 0;
            }
            if (!fa.phase1_only) c->fsync_target[0] += (uint32_t)fk_blocks;
            if (pushing) { c->fsync_target[1] += (uint32_t)fk_blocks; c->fsync_target[2] += (uint32_t)fk_blocks; }
            for (int k = 0; k < 3; k++) fa.y.target[k] = c->fsync_target[k];
            // the phases are separated by barriers over CTAs: a cooperative launch guarantees that all of them are resident
            void* kargs[] = {(void*)&fa};
            const void* fn = dense ? (count ? (const void*)k_frame<true, RT_DENSE_MIN_BLOCKS> : (const void*)k_frame<false, RT_DENSE_MIN_BLOCKS>)
                                   : (count ? (const void*)k_frame<true> : (const void*)k_frame<false>);
            RT_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)fk_blocks), dim3(TRAV_TPB), kargs, 0, st));
        }
        RT_CUDA(cudaGetLastError());
        launches++;
        c->push.done = true;
        mark(1);
        mark(2);
        if (fa.phase1_only) {
            sa.qin = queue_of(c, 0); sa.qout = queue_of(c, 1);
            sa.wave = waves_dev; sa.next = waves_dev; sa.occl = c->d_occl.p; sa.occl_bits = 1;
            launch_shade<true>(c, sa);
            launches++;
        }
    } else if (has_work) {
        launches += launch_wave<true>(c, ta, sa, 0, 1, 0, count, true, stats ? c->ev[1] : nullptr, stats ? c->ev[2] : nullptr);
    } else {
        mark(1);
        mark(2);
    }
    mark(3);

    WaveResult wr;
    const bool use_paths = c->path_kernel && c->scene.n_lights <= 8;
    if (bounce && has_work) {
        ta.aux_prim = nullptr; ta.aux_t = nullptr;
        ta.warp_times = nullptr;
        // wave 0 wrote its children into queue 1
        if (use_paths) {
            launch_paths(c, 1, 1, p->max_depth, ta.brute, count);
            launches++;
            wr.waves = 1;
        } else {
            wr = run_bounce_waves(c, ta, sa, 1, 1, p->max_depth, count, true);
            launches += wr.launches;
        }
    }
    mark(6);

    if (has_work && rgb_dev && !direct) {
        uint32_t quads = f.n_tiles_owned * (uint32_t)(f.tile_pix / 4) + (f.steal_cursor ? f.n_pool_blocks * 8u : 0u);
        k_resolve<<<(quads + 255) / 256, 256, 0, st>>>(f, c->d_accum.p, (uint8_t*)rgb_dev,
                                                       (p->flags & RT_FLAG_PACKED_TILES) ? 1 : 0, c->d_frame.p);
        RT_CUDA(cudaGetLastError());
        launches++;
    }
    // heavy tiles first in the frames that follow: re-sorted after the first two frames of a layout and then every
    // RT_TILE_SORT_EVERY-th frame (costs accumulate as maxima in between); it runs after the last kernel that indexes
    // pixels through tile_ids
    if (f.tile_cost && (c->frames_in_layout < 2 || c->frames_in_layout % (uint32_t)c->tile_sort_every == 0)) {
        k_sort_tiles<<<1, 1024, 0, st>>>(c->d_tile_ids.p, c->d_tile_ord.p, c->d_tile_ids0.p, c->d_tile_cost.p, f.n_tiles_owned,
                                         c->tile_bucket_bits);
        RT_CUDA(cudaGetLastError());
        launches++;
    }
    c->frames_in_layout++;
    c->launch_total += launches;
    mark(7);

    // Without a stats request the frame is left in flight: nothing below synchronises, errors stay in
    // the sticky word until rt_synchronize / the next stats-bearing call.
    if (!stats) return;
    if (use_fk) {
        memset(c->h_waves, 0, RT_WAVE_SLOTS * sizeof(WaveCounters));
        RT_CUDA(cudaMemcpyAsync(c->h_waves, waves_dev, sizeof(WaveCounters), cudaMemcpyDeviceToHost, st));
    } else {
        RT_CUDA(cudaMemcpyAsync(c->h_waves, c->d_waves.p, RT_WAVE_SLOTS * sizeof(WaveCounters), cudaMemcpyDeviceToHost, st));
    }
    RT_CUDA(cudaMemcpyAsync(c->h_frame, frame_dev_ctr, sizeof(FrameCounters), cudaMemcpyDeviceToHost, st));
    rt_sync_and_check(c);
    uint64_t secondary = wr.secondary;
    uint32_t max_queue = wr.max_queue;
    if (bounce && use_paths) {
        secondary = c->h_frame->rays_secondary;
        max_queue = c->h_waves[1].n_rays;
    } else if (bounce && !c->has_dielectric && p->max_depth <= 16) {   // blind mode: read populations now
        for (int w = 1; w <= p->max_depth && w < RT_WAVE_SLOTS; w++) {
            secondary += c->h_waves[w].n_rays;
            if (c->h_waves[w].n_rays > max_queue) max_queue = c->h_waves[w].n_rays;
        }
    }
    memset(stats, 0, sizeof *stats);
    stats->rays_primary = c->h_frame->rays_primary;
    stats->stolen_blocks = c->h_frame->stolen_blocks;
    stats->rays_shadow = c->h_frame->rays_shadow;
    stats->rays_secondary = secondary;
    stats->node_visits = c->h_frame->node_visits[0];
    stats->tri_tests = c->h_frame->tri_tests[0];
    stats->shadow_node_visits = c->h_frame->node_visits[1];
    stats->shadow_tri_tests = c->h_frame->tri_tests[1];
    stats->waves = 1 + wr.waves;
    stats->tiles = c->layout.n_tiles_owned;
    stats->kernel_launches = launches;
    stats->max_queue = max_queue > c->h_waves[0].n_hits ? max_queue : c->h_waves[0].n_hits;
    RT_CUDA(cudaEventElapsedTime(&stats->ms_device, c->ev[0], c->ev[7]));
    RT_CUDA(cudaEventElapsedTime(&stats->ms_trace, c->ev[0], c->ev[1]));
    RT_CUDA(cudaEventElapsedTime(&stats->ms_shadow, c->ev[1], c->ev[2]));
    RT_CUDA(cudaEventElapsedTime(&stats->ms_shade, c->ev[2], c->ev[3]));
    RT_CUDA(cudaEventElapsedTime(&stats->ms_secondary, c->ev[3], c->ev[6]));
    RT_CUDA(cudaEventElapsedTime(&stats->ms_resolve, c->ev[6], c->ev[7]));
}

void rt_query_rays(rt_ctx* c, const float* rays_host, uint32_t n, int max_depth, uint32_t flags, bool shade,
                   int32_t* prim_out, float* t_out, float* rgb_out) {
    cudaStream_t st = c->stream;
    if (n == 0) return;
    const bool count = false;
    const bool bounce = shade && rt_scene_bounces(c, max_depth);
    ensure_queues(c, n, true);
    size_t cap = c->queue_cap;
    c->d_accum.reserve(3 * (size_t)n);
    c->d_aux_prim.reserve(n);
    c->d_aux_t.reserve(n);
    // Ray's constructor normalises the direction in FP64 (ray.h:25-29)
    std::vector<float4> o(n), d(n), w(n);
    for (uint32_t i = 0; i < n; i++) {
        const float* r = rays_host + 6 * (size_t)i;
        double x = r[3], y = r[4], z = r[5];
        double l = std::sqrt(x * x + y * y + z * z);
        uint32_t pix = i;
        float pf;
        memcpy(&pf, &pix, 4);
        o[i] = make_float4(r[0], r[1], r[2], pf);
        d[i] = make_float4((float)(x / l), (float)(y / l), (float)(z / l), 0.0f);   // level 0 (bit pattern 0)
        w[i] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    }
    RT_CUDA(cudaMemcpyAsync(c->d_q[0][0].p, o.data(), n * sizeof(float4), cudaMemcpyHostToDevice, st));
    RT_CUDA(cudaMemcpyAsync(c->d_q[0][1].p, d.data(), n * sizeof(float4), cudaMemcpyHostToDevice, st));
    RT_CUDA(cudaMemcpyAsync(c->d_q[0][2].p, w.data(), n * sizeof(float4), cudaMemcpyHostToDevice, st));
    RT_CUDA(cudaMemsetAsync(c->d_waves.p, 0, RT_WAVE_SLOTS * sizeof(WaveCounters), st));
    RT_CUDA(cudaMemsetAsync(c->d_frame.p, 0, sizeof(FrameCounters), st));
    RT_CUDA(cudaMemsetAsync(c->d_accum.p, 0, 3 * (size_t)n * sizeof(long long), st));
    WaveCounters w0;
    memset(&w0, 0, sizeof w0);
    w0.n_rays = n;
    RT_CUDA(cudaMemcpyAsync(c->d_waves.p, &w0, sizeof w0, cudaMemcpyHostToDevice, st));

    TravArgs ta;
    memset(&ta, 0, sizeof ta);
    ta.s = c->scene;
    ta.accum = c->d_accum.p; ta.fc = c->d_frame.p; ta.sticky = c->d_sticky.p;
    ta.aux_prim = c->d_aux_prim.p; ta.aux_t = c->d_aux_t.p;
    ta.brute = (flags & RT_FLAG_BRUTE_FORCE) ? 1u : 0u;
    ta.cap = (uint32_t)cap;
    ShadeArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.s = c->scene; sa.hits = c->d_hits.p; sa.hitq = c->d_hitq.p; sa.accum = c->d_accum.p;
    sa.sticky = c->d_sticky.p;
    sa.cap = (uint32_t)cap; sa.max_depth = max_depth;
    // wave 0 runs through the generic (queue-fed) kernels
    launch_wave<false>(c, ta, sa, 0, 1, 0, count, shade, nullptr, nullptr);
    if (bounce) {
        ta.aux_prim = nullptr; ta.aux_t = nullptr;
        if (c->path_kernel && c->scene.n_lights <= 8) launch_paths(c, 1, 1, max_depth, ta.brute, count);
        else run_bounce_waves(c, ta, sa, 1, 1, max_depth, count, false);
    }
    if (prim_out) RT_CUDA(cudaMemcpyAsync(prim_out, c->d_aux_prim.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (t_out) RT_CUDA(cudaMemcpyAsync(t_out, c->d_aux_t.p, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (shade && rgb_out) {
        c->d_rgbf_out.reserve(3 * (size_t)n);
        k_resolve_float<<<(3 * n + 255) / 256, 256, 0, st>>>(c->d_accum.p, 3 * n, c->d_rgbf_out.p);
        RT_CUDA(cudaGetLastError());
        RT_CUDA(cudaMemcpyAsync(rgb_out, c->d_rgbf_out.p, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    rt_sync_and_check(c);
}

void rt_peer_sync_enqueue(rt_ctx* c, void* sync_buf, int rank, int world, uint32_t frame_index, int phase) {
    uint32_t* sync = (uint32_t*)sync_buf;
    cudaStream_t st = c->stream;
    if (world <= 1) return;
    if (phase == 0) {
        if (rank != 0) k_peer_wait_consumed<<<1, 1, 0, st>>>(sync, frame_index, c->d_sticky.p);
        else k_peer_publish_consumed<<<1, 1, 0, st>>>(sync, frame_index);
        c->launch_total++;
    } else {
        c->launch_total++;
        if (rank != 0) k_peer_arrive<<<1, 1, 0, st>>>(sync, frame_index);
        else k_peer_wait_all<<<1, 1, 0, st>>>(sync, frame_index, (uint32_t)(world - 1), c->d_sticky.p);
    }
    RT_CUDA(cudaGetLastError());
}

void rt_peer_barrier_enqueue(rt_ctx* c, void* sync_buf, int world, uint32_t epoch) {
    if (world <= 1) return;
    k_peer_barrier<<<1, 1, 0, c->stream>>>((uint32_t*)sync_buf, (uint32_t)world * (epoch + 1u), c->d_sticky.p);
    RT_CUDA(cudaGetLastError());
}

void rt_assemble(rt_ctx* c, const void* packed, int src_rank, int world, int width, int height, int tile_w,
                 int tile_h, void* frame) {
    int tiles_x = (width + tile_w - 1) / tile_w, tiles_y = (height + tile_h - 1) / tile_h;
    uint32_t total = (uint32_t)tiles_x * (uint32_t)tiles_y;
    uint32_t n_owned = total > (uint32_t)src_rank ? (total - (uint32_t)src_rank + (uint32_t)world - 1) / (uint32_t)world : 0u;
    if (!n_owned) return;
    bool wide = (width * 3) % 16 == 0 && (tile_w * 3) % 16 == 0 && ((uintptr_t)packed & 15) == 0 && ((uintptr_t)frame & 15) == 0;
    if (wide) {
        uint32_t chunks = n_owned * (uint32_t)(tile_w * 3 / 16) * (uint32_t)tile_h;
        k_assemble16<<<(chunks + 255) / 256, 256, 0, c->stream>>>((const uint8_t*)packed, (uint8_t*)frame, width, height,
                                                                  tile_w, tile_h, tiles_x, total, src_rank, world);
    } else {
        uint32_t quads = n_owned * (uint32_t)(tile_w * tile_h / 4);
        k_assemble<<<(quads + 255) / 256, 256, 0, c->stream>>>((const uint8_t*)packed, (uint8_t*)frame, width, height,
                                                               tile_w, tile_h, tiles_x, total, src_rank, world);
    }
    c->launch_total++;
    RT_CUDA(cudaGetLastError());
}
