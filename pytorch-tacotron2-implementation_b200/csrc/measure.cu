// Measurement aids used by bench.py for the roofline denominators that MEASURED_PEAKS.json
// does not hold: aggregate L2 -> SM read bandwidth on a buffer that fits in
// L2, and the latency of the decoder's grid-wide barrier.
#include "handle.h"

namespace b200taco {

namespace {

__global__ void __launch_bounds__(512) l2_read_kernel(const uint4* __restrict__ buf, size_t n_vec, int iters,
                                                      unsigned long long* sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) {
        size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n_vec; i += 4 * stride) {
            const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride),
                        d = __ldcg(buf + i + 3 * stride);
            acc.x ^= a.x ^ b.x ^ c.x ^ d.x; acc.y ^= a.y ^ b.y ^ c.y ^ d.y;
            acc.z ^= a.z ^ b.z ^ c.z ^ d.z; acc.w ^= a.w ^ b.w ^ c.w ^ d.w;
        }
        for (; i < n_vec; i += stride) {
            const uint4 a = __ldcg(buf + i);
            acc.x ^= a.x; acc.y ^= a.y; acc.z ^= a.z; acc.w ^= a.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) *sink = acc.x;   // keeps the loads alive
}

// Per-SM streaming probe: every CTA repeatedly reads its own `per_cta` bytes (L2 resident after the
// first pass).  mode 0: LDG.128, 8 independent loads per thread; mode >= 1: TMA bulk copies of
// `chunk` bytes with `depth` copies in flight (one issuing thread), the pattern of the decoder pipeline.
__global__ void __launch_bounds__(256, 1) stream_probe_kernel(const uint8_t* __restrict__ buf, size_t per_cta, int iters,
                                                              int mode, int chunk, int depth,
                                                              unsigned long long* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint8_t* src = buf + (size_t)blockIdx.x * per_cta;
    if (mode == 0) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        const size_t n = per_cta / 16;
        for (int it = 0; it < iters; ++it) {
            for (size_t i = threadIdx.x; i + 7 * 256 < n; i += 8 * 256) {
                uint4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldcg(s4 + i + j * 256);
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc.x ^= v[j].x; acc.y ^= v[j].y; acc.z ^= v[j].z; acc.w ^= v[j].w; }
            }
        }
        if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) *sink = acc.x;
        return;
    }
    // mode = number of issuing warps (lane 0 of warps 0..mode-1), each with its own `depth` slots and
    // its own interleaved share of the chunks
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);          // mode * depth barriers
    uint8_t* stage = smem + 1024;
    const int nw = mode;
    if (threadIdx.x == 0) {
        for (int d = 0; d < nw * (depth < 0 ? -depth : depth); ++d) mbar_init(&bars[d], 1);
        fence_mbar_init();
    }
    __syncthreads();
    // depth < 0 selects lanes 0..mode-1 of warp 0 as issuers instead of lane 0 of warps 0..mode-1
    const bool same_warp = depth < 0;
    depth = depth < 0 ? -depth : depth;
    const int w = same_warp ? (int)threadIdx.x : (int)(threadIdx.x >> 5);
    if ((same_warp ? threadIdx.x < (unsigned)nw : ((threadIdx.x & 31) == 0 && w < nw))) {
        const int nchunks = (int)(per_cta / chunk);
        const long long total = (long long)(nchunks / nw) * iters;
        uint64_t* mybars = bars + w * depth;
        uint8_t* mystage = stage + (size_t)w * depth * chunk;
        for (long long c = 0; c < total + depth; ++c) {
            const int d = (int)(c % depth);
            if (c >= depth) {
                const uint32_t parity = (uint32_t)(((c - depth) / depth) & 1);
                while (!mbar_try_wait(&mybars[d], parity)) {}
            }
            if (c < total) {
                mbar_arrive_expect_tx(&mybars[d], (uint32_t)chunk);
                const size_t idx = (size_t)((c % (nchunks / nw)) * nw + w);
                bulk_g2s(mystage + (size_t)d * chunk, src + idx * chunk, (uint32_t)chunk, &mybars[d]);
            }
        }
    }
}

// cycles per operation of the synchronisation primitives the decoder uses (single thread, after a store)
__global__ void sync_cost_kernel(uint32_t* scratch, long long* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    constexpr int N = 256;
    long long t0, t1;
    t0 = clock64();
    for (int i = 0; i < N; ++i) { scratch[i & 63] = i; asm volatile("fence.proxy.async;" ::: "memory"); }
    t1 = clock64(); out[0] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) { scratch[i & 63] = i; asm volatile("fence.proxy.async.global;" ::: "memory"); }
    t1 = clock64(); out[1] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) { scratch[i & 63] = i; asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    t1 = clock64(); out[2] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) { scratch[i & 63] = i; __threadfence(); }
    t1 = clock64(); out[3] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) { scratch[i & 63] = i; red_release_add_u32(scratch + 64, 1u); }
    t1 = clock64(); out[4] = (t1 - t0) / N;
    t0 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < N; ++i) { acc += ld_acquire_u32(scratch + 64 + (acc & 1)); }
    t1 = clock64(); out[5] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) { acc += atomicAdd(scratch + 66, 1u); }
    t1 = clock64(); out[6] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) { scratch[i & 63] = i; }
    t1 = clock64(); out[7] = (t1 - t0) / N;
    scratch[70] = acc;
}

}  // namespace

int measure_sync_costs(long long* out8_host, cudaStream_t stream) {
    uint32_t* scratch = nullptr;
    long long* out = nullptr;
    TACO_CUDA(cudaMalloc(&scratch, 1024));
    TACO_CUDA(cudaMalloc(&out, 64));
    TACO_CUDA(cudaMemsetAsync(scratch, 0, 1024, stream));
    sync_cost_kernel<<<1, 32, 0, stream>>>(scratch, out);
    TACO_LAUNCHED();
    TACO_CUDA(cudaMemcpyAsync(out8_host, out, 64, cudaMemcpyDeviceToHost, stream));
    TACO_CUDA(cudaStreamSynchronize(stream));
    cudaFree(scratch);
    cudaFree(out);
    return B200TACO_OK;
}

int measure_stream_probe(void* d_buf, size_t per_cta, int ctas, int iters, int mode, int chunk, int depth, float* gbps_per_sm,
                         cudaStream_t stream) {
    TACO_REQUIRE(d_buf && gbps_per_sm && per_cta >= 65536 && ctas >= 1 && iters >= 1, B200TACO_ERR_INVALID, "stream_probe: bad arguments");
    TACO_REQUIRE(mode == 0 || (chunk >= 512 && chunk % 16 == 0 && depth != 0 && depth >= -32 && depth <= 32 && (size_t)chunk * (depth < 0 ? -depth : depth) * mode <= 200 * 1024 && mode <= 8 &&
                               per_cta % chunk == 0),
                 B200TACO_ERR_INVALID, "stream_probe: bad chunk/depth");
    unsigned long long* sink = nullptr;
    TACO_CUDA(cudaMalloc(&sink, 8));
    const int smem = mode == 0 ? 0 : 1024 + chunk * (depth < 0 ? -depth : depth) * mode;
    TACO_CUDA(cudaFuncSetAttribute(stream_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaEvent_t e0, e1;
    TACO_CUDA(cudaEventCreate(&e0));
    TACO_CUDA(cudaEventCreate(&e1));
    stream_probe_kernel<<<ctas, 256, smem, stream>>>(static_cast<const uint8_t*>(d_buf), per_cta, 1, mode, chunk, depth, sink);
    TACO_LAUNCHED();
    TACO_CUDA(cudaEventRecord(e0, stream));
    stream_probe_kernel<<<ctas, 256, smem, stream>>>(static_cast<const uint8_t*>(d_buf), per_cta, iters, mode, chunk, depth, sink);
    TACO_LAUNCHED();
    TACO_CUDA(cudaEventRecord(e1, stream));
    TACO_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    TACO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *gbps_per_sm = (float)((double)per_cta * iters / (ms * 1e-3) / 1e9);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return B200TACO_OK;
}

int measure_l2_read(void* d_buf, size_t bytes, int iters, float* gbps, cudaStream_t stream) {
    TACO_REQUIRE(d_buf && gbps && bytes >= 4096 && iters >= 1, B200TACO_ERR_INVALID, "measure_l2_read: bad arguments");
    int dev = 0, sms = 0;
    TACO_CUDA(cudaGetDevice(&dev));
    TACO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned long long* sink = nullptr;
    TACO_CUDA(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1;
    TACO_CUDA(cudaEventCreate(&e0));
    TACO_CUDA(cudaEventCreate(&e1));
    const size_t n_vec = bytes / 16;
    l2_read_kernel<<<sms * 2, 512, 0, stream>>>(static_cast<const uint4*>(d_buf), n_vec, 2, sink);   // warm L2
    TACO_LAUNCHED();
    TACO_CUDA(cudaEventRecord(e0, stream));
    l2_read_kernel<<<sms * 2, 512, 0, stream>>>(static_cast<const uint4*>(d_buf), n_vec, iters, sink);
    TACO_LAUNCHED();
    TACO_CUDA(cudaEventRecord(e1, stream));
    TACO_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    TACO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *gbps = (float)((double)n_vec * 16.0 * iters / (ms * 1e-3) / 1e9);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return B200TACO_OK;
}

}  // namespace b200taco
