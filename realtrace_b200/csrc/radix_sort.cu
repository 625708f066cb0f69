// radix_sort.cu — stable LSD radix sort of (64-bit Morton key, 32-bit triangle index) pairs.
//
// Hand-written for the LBVH builder (north star: "radix sort"); replaces the thrust reduce/scan +
// atomics the reference's CUDA grid builder leans on (/root/reference/Parellel/kernel.cu:486-514).
// 8-bit digits.  One up-front pass histograms all eight digit positions (k_rs_digit_hist); everything a
// pass needs to know about the key distribution is read from that table ON THE DEVICE — the sort never
// comes back to the host.  A pass in which every key has the same digit (the high bytes of 32-bit keys,
// of a tiny scene) is recognised by each kernel from the table: histogram and offsets return at once and
// the scatter degenerates to a copy of its tile, so the result always lands in buffer 0 after eight passes.
// Per pass:
//   k_rs_hist     per-block digit histogram (shared-memory atomics)         hist[digit][block]
//   k_rs_offsets  one CTA per digit: exclusive scan of that digit's row of block counts, started at the
//                 digit's global base (sum of the smaller digits' totals from the up-front table)
//   k_rs_scatter  each warp owns a contiguous 512-key chunk of the CTA's 4096-key tile and ranks
//                 it 32 keys at a time with __match_any_sync; chunk bases come from a per-digit
//                 prefix over the CTA's 8 warps.  Order inside a digit = original order => stable.
// Keys are streamed twice per pass (24 B/key read + 12 B/key written); at N = 1 M that is 36 MB per
// pass — L2 resident on B200 — so the sort is latency/launch bound, not HBM bound.
#include "rt_context.h"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;
constexpr int RS_WARP_CHUNK = 32 * RS_ROUNDS;       // 512
constexpr int RS_TILE = RS_WARP_CHUNK * RS_WARPS;   // 4096

__global__ void __launch_bounds__(RS_THREADS) k_rs_digit_hist(const uint64_t* __restrict__ keys, uint32_t n,
                                                             uint32_t* __restrict__ digit_hist /*[8][256]*/) {
    __shared__ uint32_t h[8 * 256];
    for (int i = threadIdx.x; i < 8 * 256; i += RS_THREADS) h[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += gridDim.x * RS_THREADS) {
        uint64_t k = keys[i];
#pragma unroll
        for (int d = 0; d < 8; d++) atomicAdd(&h[d * 256 + (uint32_t)((k >> (8 * d)) & 0xff)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 256; i += RS_THREADS)
        if (h[i]) atomicAdd(&digit_hist[i], h[i]);
}

// True (for the whole CTA of 256 threads) when one digit value holds all n keys of this pass.
__device__ __forceinline__ bool pass_is_trivial(const uint32_t* __restrict__ digit_row, uint32_t n) {
    return __syncthreads_or(digit_row[threadIdx.x] == n) != 0;
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift,
                                                       uint32_t* __restrict__ hist, uint32_t nblocks,
                                                       const uint32_t* __restrict__ digit_row) {
    __shared__ uint32_t h[256];
    if (pass_is_trivial(digit_row, n)) return;
    h[threadIdx.x] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int it = 0; it < RS_TILE / RS_THREADS; it++) {
        uint32_t i = base + it * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)((keys[i] >> shift) & 0xff)], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// CTA d turns row d of hist (the count of digit d in every block, block-major) into the global offsets of
// those blocks' digit-d keys: digit base + exclusive prefix over the blocks.  256 CTAs of 256 threads.
__global__ void __launch_bounds__(RS_THREADS) k_rs_offsets(uint32_t* __restrict__ hist, uint32_t nblocks,
                                                          const uint32_t* __restrict__ digit_row, uint32_t n,
                                                          uint32_t* __restrict__ passes_done) {
    __shared__ uint32_t ws[2][RS_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t d = blockIdx.x;
    const uint32_t cnt = digit_row[threadIdx.x];
    if (__syncthreads_or(cnt == n)) return;
    if (d == 0 && threadIdx.x == 0) atomicAdd(passes_done, 1u);
    // digit base = keys with a smaller digit
    uint32_t v = threadIdx.x < d ? cnt : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) ws[0][warp] = v;
    __syncthreads();
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) run += ws[0][w];
    uint32_t* __restrict__ row = hist + (size_t)d * nblocks;
    int buf = 1;
    for (uint32_t base = 0; base < nblocks; base += RS_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < nblocks ? row[i] : 0u;
        uint32_t incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) ws[buf][warp] = incl;
        __syncthreads();                       // ws[buf] complete; ws[buf ^ 1] of the previous round fully read
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t t = ws[buf][w];
            if (w < warp) before += t;
            total += t;
        }
        if (i < nblocks) row[i] = run + before + incl - x;
        run += total;
        buf ^= 1;
    }
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keys_in,
                                                          const uint32_t* __restrict__ vals_in,
                                                          uint64_t* __restrict__ keys_out,
                                                          uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                          const uint32_t* __restrict__ offsets, uint32_t nblocks,
                                                          const uint32_t* __restrict__ digit_row) {
    __shared__ uint32_t wh[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (pass_is_trivial(digit_row, n)) {       // one digit value only: the stable permutation is the identity
        const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll 4
        for (int it = 0; it < RS_TILE / RS_THREADS; it++) {
            uint32_t i = base + it * RS_THREADS + threadIdx.x;
            if (i < n) { keys_out[i] = keys_in[i]; vals_out[i] = vals_in[i]; }
        }
        return;
    }
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();
    const uint32_t chunk = blockIdx.x * RS_TILE + warp * RS_WARP_CHUNK;
    const uint32_t lt = (1u << lane) - 1u;

    // phase A: per-warp digit counts of its chunk
    for (int r = 0; r < RS_ROUNDS; r++) {
        uint32_t i = chunk + r * 32 + lane;
        bool valid = i < n;
        uint32_t digit = valid ? (uint32_t)((keys_in[i] >> shift) & 0xff) : 256u + lane;
        uint32_t peers = __match_any_sync(0xffffffffu, digit);
        if (valid && (peers & lt) == 0) wh[warp][digit] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // phase B: digit `threadIdx.x`: global base of this CTA + exclusive prefix over its warps
    {
        uint32_t run = offsets[threadIdx.x * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = wh[w][threadIdx.x];
            wh[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase C: rank and scatter
    for (int r = 0; r < RS_ROUNDS; r++) {
        uint32_t i = chunk + r * 32 + lane;
        bool valid = i < n;
        uint64_t k = valid ? keys_in[i] : 0ull;
        uint32_t digit = valid ? (uint32_t)((k >> shift) & 0xff) : 256u + lane;
        uint32_t peers = __match_any_sync(0xffffffffu, digit);
        uint32_t base = valid ? wh[warp][digit] : 0u;
        __syncwarp();
        if (valid) {
            uint32_t dst = base + __popc(peers & lt);
            keys_out[dst] = k;
            vals_out[dst] = vals_in[i];
            if ((peers & lt) == 0) wh[warp][digit] = base + __popc(peers);
        }
        __syncwarp();
    }
}

}  // namespace

// Sorts d_keys[0]/d_vals[0] (n pairs); the result is back in d_keys[0]/d_vals[0] (eight passes, sorted_buf = 0).
// Only enqueues: nothing is read back.  d_digit_hist[RS_PASS_COUNTER] counts the passes that moved keys; the
// caller fetches it with rt_sort_passes_done() behind a synchronisation of its own.
constexpr uint32_t RS_PASS_COUNTER = 8 * 256;

void rt_sort_pairs_device(rt_ctx* c, uint32_t n) {
    c->sorted_buf = 0;
    if (n < 2) return;
    cudaStream_t st = c->stream;
    c->d_keys[1].reserve(n);
    c->d_vals[1].reserve(n);
    c->d_digit_hist.reserve(8 * 256 + 1);
    uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
    c->d_hist.reserve((size_t)256 * nblocks);

    RT_CUDA(cudaMemsetAsync(c->d_digit_hist.p, 0, (8 * 256 + 1) * sizeof(uint32_t), st));
    int hist_blocks = (int)((n + RS_THREADS * 8 - 1) / (RS_THREADS * 8));
    if (hist_blocks > c->sm_count * 4) hist_blocks = c->sm_count * 4;
    k_rs_digit_hist<<<hist_blocks, RS_THREADS, 0, st>>>(c->d_keys[0].p, n, c->d_digit_hist.p);
    RT_CUDA(cudaGetLastError());

    int cur = 0;
    for (int d = 0; d < 8; d++) {
        const int shift = 8 * d;
        const uint32_t* row = c->d_digit_hist.p + 256 * d;
        k_rs_hist<<<nblocks, RS_THREADS, 0, st>>>(c->d_keys[cur].p, n, shift, c->d_hist.p, nblocks, row);
        k_rs_offsets<<<256, RS_THREADS, 0, st>>>(c->d_hist.p, nblocks, row, n, c->d_digit_hist.p + RS_PASS_COUNTER);
        k_rs_scatter<<<nblocks, RS_THREADS, 0, st>>>(c->d_keys[cur].p, c->d_vals[cur].p, c->d_keys[cur ^ 1].p,
                                                     c->d_vals[cur ^ 1].p, n, shift, c->d_hist.p, nblocks, row);
        RT_CUDA(cudaGetLastError());
        cur ^= 1;
    }
    c->sorted_buf = cur;   // 0
}

// Number of passes of the last rt_sort_pairs_device that had more than one digit value (synchronises the stream).
int rt_sort_passes_done(rt_ctx* c) {
    if (!c->d_digit_hist.p || c->d_digit_hist.cap < 8 * 256 + 1) return 0;
    uint32_t h = 0;
    RT_CUDA(cudaMemcpyAsync(&h, c->d_digit_hist.p + RS_PASS_COUNTER, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    return (int)h;
}
