// microbench.cu — the roofline denominators of the traversal kernels, measured on the device the context runs on
// ("memory = measured L2 read bandwidth (float4 __ldg over a 64 MB buffer, all SMs)"; FMA-pipe issue
// at the clock the GPU really runs).  Nothing here is on the render path; bench.py calls rt_microbench once per
// run and reports the figures next to MEASURED_PEAKS.json's HBM copy bandwidth.
#include "rt_context.h"

namespace {

// Every thread streams float4 records with the read-only path (`__ldg`, what the node/triangle fetches use);
// `passes` sweeps over the same buffer: a 64 MiB buffer stays in the 126 MB L2 after the first sweep, a 2 GiB
// buffer never does.
__global__ void __launch_bounds__(256) k_mb_read(const float4* __restrict__ buf, size_t n_vec, int passes, float* __restrict__ sink) {
    float acc = 0.0f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; p++) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        // four independent loads in flight per thread
        for (; i + 3 * stride < n_vec; i += 4 * stride) {
            float4 a = __ldg(buf + i), b = __ldg(buf + i + stride), c = __ldg(buf + i + 2 * stride), d = __ldg(buf + i + 3 * stride);
            acc += a.x + b.y + c.z + d.w;
        }
        for (; i < n_vec; i += stride) acc += __ldg(buf + i).x;
    }
    if (acc == 123.456f) *sink = acc;      // never true: keeps the loads alive
}

// Random 64-byte records (4 x float4, the BVH node format) out of an L2-resident buffer: the access pattern of
// a traversal step, with `chains` independent dependent-chains... each thread follows ONE chain (the next index
// depends on the loaded data), so this measures latency-bound record throughput at full occupancy.
__global__ void __launch_bounds__(128) k_mb_chase(const float4* __restrict__ nodes, uint32_t n_nodes, int steps, uint32_t* __restrict__ sink) {
    uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u % n_nodes;
    float acc = 0.0f;
    for (int s = 0; s < steps; s++) {
        const float4* p = nodes + 4 * (size_t)idx;
        float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
        acc += a.x + b.x + c.x;
        idx = (__float_as_uint(d.x) + (uint32_t)s) % n_nodes;      // d.x holds a pseudo-random next index
    }
    if (acc == 123.456f) *sink = idx;
    if (idx == 0xffffffffu) *sink = 1;
}

__global__ void k_mb_fill(float4* __restrict__ nodes, uint32_t n_nodes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    uint32_t h = i * 747796405u + 2891336453u;
    h = ((h >> ((h >> 28u) + 4u)) ^ h) * 277803737u;
    h = (h >> 22u) ^ h;
    nodes[4 * (size_t)i + 0] = make_float4(1.0f, 2.0f, 3.0f, 4.0f);
    nodes[4 * (size_t)i + 1] = make_float4(1.0f, 2.0f, 3.0f, 4.0f);
    nodes[4 * (size_t)i + 2] = make_float4(1.0f, 2.0f, 3.0f, 4.0f);
    nodes[4 * (size_t)i + 3] = make_float4(__uint_as_float(h % n_nodes), 0.0f, 0.0f, 0.0f);
}

// 8 independent FFMA chains per thread: the FP32 FMA pipe at full issue rate (128 lanes per SM and clock).
__global__ void __launch_bounds__(256) k_mb_fma(int iters, float seed, float* __restrict__ sink) {
    float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const float m = 1.0000001f, c = 1e-7f;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456f) *sink = s;
}

float time_ms(rt_ctx* c, cudaEvent_t a, cudaEvent_t b) {
    float ms = 0;
    RT_CUDA(cudaEventSynchronize(b));
    RT_CUDA(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

}  // namespace

void rt_run_microbench(rt_ctx* c, rt_microbench_result* out) {
    cudaStream_t st = c->stream;
    memset(out, 0, sizeof *out);
    DevBuf<float4> buf;
    DevBuf<float> sink;
    sink.reserve(4);
    const int blocks = c->sm_count * 8;
    cudaEvent_t e0 = c->ev[0], e1 = c->ev[1];

    // ---- L2 read bandwidth: 64 MiB, 1 warm sweep + 20 timed sweeps
    {
        size_t bytes = 64ull << 20, n_vec = bytes / sizeof(float4);
        buf.reserve(n_vec);
        RT_CUDA(cudaMemsetAsync(buf.p, 0, bytes, st));
        k_mb_read<<<blocks, 256, 0, st>>>(buf.p, n_vec, 2, sink.p);
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            RT_CUDA(cudaEventRecord(e0, st));
            k_mb_read<<<blocks, 256, 0, st>>>(buf.p, n_vec, 20, sink.p);
            RT_CUDA(cudaEventRecord(e1, st));
            RT_CUDA(cudaGetLastError());
            float ms = time_ms(c, e0, e1);
            if (ms < best) best = ms;
        }
        out->l2_read_gbs = (double)bytes * 20 / (best * 1e-3) / 1e9;
        out->l2_buffer_mib = 64;
    }
    // ---- random 64-byte records out of the same 64 MiB (1 Mi nodes), one dependent chain per thread
    {
        uint32_t n_nodes = (uint32_t)((64ull << 20) / 64);
        k_mb_fill<<<(n_nodes + 255) / 256, 256, 0, st>>>(buf.p, n_nodes);
        int per_sm = 8, steps = 256;
        int grid = c->sm_count * per_sm;
        k_mb_chase<<<grid, 128, 0, st>>>(buf.p, n_nodes, 32, (uint32_t*)sink.p);
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            RT_CUDA(cudaEventRecord(e0, st));
            k_mb_chase<<<grid, 128, 0, st>>>(buf.p, n_nodes, steps, (uint32_t*)sink.p);
            RT_CUDA(cudaEventRecord(e1, st));
            RT_CUDA(cudaGetLastError());
            float ms = time_ms(c, e0, e1);
            if (ms < best) best = ms;
        }
        double records = (double)grid * 128 * steps;
        out->l2_random_node_gbs = records * 64 / (best * 1e-3) / 1e9;
        out->l2_dependent_fetch_ns = best * 1e6 / steps;      // one chain step: the L2 round trip of a 64-byte record under load
    }
    // ---- HBM read bandwidth: 2 GiB, one sweep (nothing of it is in L2 when its turn comes)
    {
        size_t bytes = 2ull << 30, n_vec = bytes / sizeof(float4);
        bool ok = true;
        try { buf.reserve(n_vec); } catch (const RtError&) { ok = false; cudaGetLastError(); }
        if (ok) {
            RT_CUDA(cudaMemsetAsync(buf.p, 0, bytes, st));
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                RT_CUDA(cudaEventRecord(e0, st));
                k_mb_read<<<blocks, 256, 0, st>>>(buf.p, n_vec, 1, sink.p);
                RT_CUDA(cudaEventRecord(e1, st));
                RT_CUDA(cudaGetLastError());
                float ms = time_ms(c, e0, e1);
                if (ms < best) best = ms;
            }
            out->hbm_read_gbs = (double)bytes / (best * 1e-3) / 1e9;
        }
    }
    // ---- FMA issue: 8 CTAs of 256 threads per SM (2048 threads, full occupancy), 64 FFMA per iteration
    {
        int iters = 4096;
        k_mb_fma<<<blocks, 256, 0, st>>>(64, 1.0f, sink.p);
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            RT_CUDA(cudaEventRecord(e0, st));
            k_mb_fma<<<blocks, 256, 0, st>>>(iters, 1.0f, sink.p);
            RT_CUDA(cudaEventRecord(e1, st));
            RT_CUDA(cudaGetLastError());
            float ms = time_ms(c, e0, e1);
            if (ms < best) best = ms;
        }
        double lane_instr = (double)blocks * 256 * iters * 64;
        out->fma_lane_instr_per_s = lane_instr / (best * 1e-3);
        out->issue_warp_instr_per_s = out->fma_lane_instr_per_s / 32.0;
        out->implied_sm_mhz = out->fma_lane_instr_per_s / (128.0 * c->sm_count) / 1e6;
    }
    out->sm_count = c->sm_count;
    RT_CUDA(cudaStreamSynchronize(st));
}
