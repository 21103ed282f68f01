// Do ld.relaxed.gpu loads of one thread overlap?  N independent loads (different lines) per round, timed.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
template <int V> __device__ __forceinline__ uint64_t ld64(const uint64_t* p) {
    uint64_t w;
    if (V == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (V == 1) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (V == 2) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(w) : "l"(p));
    if (V == 3) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p));          // no memory clobber
    return w;
}
template <int V, int N>
__global__ void loads(const uint64_t* buf, int iters, int threads_active, unsigned long long* out) {
    if (threadIdx.x >= threads_active) return;
    uint64_t acc = 0;
    const uint64_t* p = buf + threadIdx.x;
    unsigned long long t0 = gtime();
    for (int i = 0; i < iters; ++i) {
        uint64_t w[N];
#pragma unroll
        for (int q = 0; q < N; ++q) w[q] = ld64<V>(p + (size_t)q * 4096 + (acc & 1));
#pragma unroll
        for (int q = 0; q < N; ++q) acc += w[q];
    }
    if (threadIdx.x == 0) { out[0] = gtime() - t0; out[1] = acc; }
}
template <int V, int N> void run(const char* name, uint64_t* buf, unsigned long long* out, int threads) {
    const int iters = 2000; unsigned long long h;
    loads<V, N><<<1, 512>>>(buf, iters, threads, out); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
    printf("%-34s N=%2d threads=%3d: %.0f ns per round\n", name, N, threads, (double)h / iters);
}
int main() {
    uint64_t* buf; unsigned long long* out;
    CK(cudaMalloc(&buf, 1 << 22)); CK(cudaMalloc(&out, 64)); CK(cudaMemset(buf, 0, 1 << 22));
    for (int th : {32, 224, 512}) {
        run<0, 1>("ld.relaxed.gpu (memory clobber)", buf, out, th); run<0, 4>("ld.relaxed.gpu (memory clobber)", buf, out, th);
        run<0, 12>("ld.relaxed.gpu (memory clobber)", buf, out, th);
        run<3, 12>("ld.relaxed.gpu (no clobber)", buf, out, th);
        run<1, 12>("ld.volatile", buf, out, th); run<2, 12>("ld.cg", buf, out, th);
    }
    return 0;
}
