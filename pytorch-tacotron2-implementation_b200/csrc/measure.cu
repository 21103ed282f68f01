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

}  // namespace

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
