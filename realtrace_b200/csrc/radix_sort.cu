// radix_sort.cu — stable LSD radix sort of (64-bit Morton key, 32-bit triangle index) pairs.
//
// Hand-written for the LBVH builder (north star: "radix sort"); replaces the thrust reduce/scan +
// atomics the reference's CUDA grid builder leans on (/root/reference/Parellel/kernel.cu:486-514).
// 8-bit digits.  One up-front pass histograms all eight digit positions so that passes in which
// every key has the same digit (most of the high bytes of a small scene) are skipped.  Per
// executed pass:
//   k_rs_hist     per-block digit histogram (shared-memory atomics)         hist[digit][block]
//   k_rs_scan     exclusive scan of hist in digit-major order (one CTA)
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

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift,
                                                       uint32_t* __restrict__ hist, uint32_t nblocks) {
    __shared__ uint32_t h[256];
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

// In-place exclusive scan of `total` counters by one CTA of 1024 threads.
__global__ void __launch_bounds__(1024) k_rs_scan(uint32_t* __restrict__ hist, uint32_t total) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < total; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < total ? hist[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sum[lane];
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sum[lane] = wi - w;   // exclusive prefix of warp totals
        }
        __syncthreads();
        uint32_t r = running;
        if (i < total) hist[i] = r + warp_sum[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) running = r + warp_sum[31] + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keys_in,
                                                          const uint32_t* __restrict__ vals_in,
                                                          uint64_t* __restrict__ keys_out,
                                                          uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                          const uint32_t* __restrict__ offsets, uint32_t nblocks) {
    __shared__ uint32_t wh[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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

// Sorts d_keys[0]/d_vals[0] (n pairs); leaves the result in d_keys[sorted_buf]/d_vals[sorted_buf].
void rt_sort_pairs_device(rt_ctx* c, uint32_t n, int* sort_passes) {
    c->sorted_buf = 0;
    if (sort_passes) *sort_passes = 0;
    if (n < 2) return;
    cudaStream_t st = c->stream;
    c->d_keys[1].reserve(n);
    c->d_vals[1].reserve(n);
    c->d_digit_hist.reserve(8 * 256);
    uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
    c->d_hist.reserve((size_t)256 * nblocks);

    RT_CUDA(cudaMemsetAsync(c->d_digit_hist.p, 0, 8 * 256 * sizeof(uint32_t), st));
    int hist_blocks = (int)((n + RS_THREADS * 8 - 1) / (RS_THREADS * 8));
    if (hist_blocks > c->sm_count * 4) hist_blocks = c->sm_count * 4;
    k_rs_digit_hist<<<hist_blocks, RS_THREADS, 0, st>>>(c->d_keys[0].p, n, c->d_digit_hist.p);
    RT_CUDA(cudaGetLastError());
    std::vector<uint32_t> dh(8 * 256);
    RT_CUDA(cudaMemcpyAsync(dh.data(), c->d_digit_hist.p, dh.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));

    int cur = 0, passes = 0;
    for (int d = 0; d < 8; d++) {
        bool trivial = false;
        for (int b = 0; b < 256; b++)
            if (dh[d * 256 + b] == n) { trivial = true; break; }
        if (trivial) continue;
        int shift = 8 * d;
        k_rs_hist<<<nblocks, RS_THREADS, 0, st>>>(c->d_keys[cur].p, n, shift, c->d_hist.p, nblocks);
        k_rs_scan<<<1, 1024, 0, st>>>(c->d_hist.p, 256 * nblocks);
        k_rs_scatter<<<nblocks, RS_THREADS, 0, st>>>(c->d_keys[cur].p, c->d_vals[cur].p, c->d_keys[cur ^ 1].p,
                                                     c->d_vals[cur ^ 1].p, n, shift, c->d_hist.p, nblocks);
        RT_CUDA(cudaGetLastError());
        cur ^= 1;
        passes++;
    }
    c->sorted_buf = cur;
    if (sort_passes) *sort_passes = passes;
}
