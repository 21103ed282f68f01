// LL exchange microbenchmarks: what does one cross-CTA hop cost on B200?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void ll_st(uint64_t* p, float v, uint32_t tag) {
    const uint64_t w = ((uint64_t)tag << 32) | (uint64_t)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint64_t ll_ld1(const uint64_t* p) {
    uint64_t w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ void ll_ld2(const uint64_t* p, uint64_t& a, uint64_t& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// (a) ping-pong between CTA 0 and CTA peer, one thread each
__global__ void pingpong(uint64_t* buf, int peer, int iters, unsigned long long* out) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me != 0 && me != peer) return;
    uint64_t* mine = buf + (me == 0 ? 0 : 16);      // separate 128-byte lines
    uint64_t* theirs = buf + (me == 0 ? 16 : 0);
    unsigned long long t0 = gtime();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            ll_st(theirs, 1.f, i);
            while ((uint32_t)(ll_ld1(mine) >> 32) != (uint32_t)i) {}
        } else {
            while ((uint32_t)(ll_ld1(mine) >> 32) != (uint32_t)i) {}
            ll_st(theirs, 1.f, i);
        }
    }
    if (me == 0) out[0] = gtime() - t0;
}

// (b) all-gather stage: every CTA publishes U words (x R replicas), every CTA gathers N = U * nCTA words, repeat.
template <int RT>
__global__ void __launch_bounds__(RT, 1) allgather(uint64_t* buf, int nprod, int U, int R, int iters, int work, unsigned long long* out) {
    __shared__ float x[4096];
    const int tid = threadIdx.x, cta = blockIdx.x, n = nprod * U;
    const int rep = cta % R;
    unsigned long long t0 = gtime();
    float accum = 0.f;
    for (int i = 1; i <= iters; ++i) {
        uint64_t* X = buf + (size_t)(i & 3) * (R * n);
        // publish
        if (cta < nprod && tid < U * R) ll_st(X + (size_t)(tid / U) * n + cta * U + (tid % U), accum + 1.f, i);
        // gather
        const uint64_t* src = X + (size_t)rep * n;
        for (int k = tid; k < n / 2; k += RT) {
            uint64_t a, b;
            ll_ld2(src + 2 * k, a, b);
            while ((uint32_t)(a >> 32) != (uint32_t)i || (uint32_t)(b >> 32) != (uint32_t)i) ll_ld2(src + 2 * k, a, b);
            x[2 * k] = __uint_as_float((uint32_t)a);
            x[2 * k + 1] = __uint_as_float((uint32_t)b);
        }
        __syncthreads();
        if (work) {     // a dependent warp-level reduction + second sync, like a GEMV stage
            float s = 0.f;
            for (int k = tid & 31; k < n; k += 32) s += x[k];
            for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
            accum = s * 1e-9f;
            __syncthreads();
        }
    }
    if (cta == 0 && tid == 0) out[0] = gtime() - t0;
}

// (c) cluster-of-2 DSMEM LL ping-pong
__global__ void __cluster_dims__(2, 1, 1) dsmem_pingpong(int iters, unsigned long long* out) {
    __shared__ uint64_t slot[16];
    cg::cluster_group cl = cg::this_cluster();
    const int me = cl.block_rank();
    if (threadIdx.x == 0) slot[0] = 0;
    cl.sync();
    if (threadIdx.x == 0) {
        uint64_t* remote = cl.map_shared_rank(slot, me ^ 1);
        volatile uint64_t* mine = slot;
        unsigned long long t0 = gtime();
        for (int i = 1; i <= iters; ++i) {
            const uint64_t w = ((uint64_t)i << 32) | 5u;
            if (me == 0) {
                *(volatile uint64_t*)remote = w;
                while ((uint32_t)(*mine >> 32) != (uint32_t)i) {}
            } else {
                while ((uint32_t)(*mine >> 32) != (uint32_t)i) {}
                *(volatile uint64_t*)remote = w;
            }
        }
        if (me == 0 && blockIdx.x == 0) out[0] = gtime() - t0;
    }
    cl.sync();
}

// (d) timer cost
__global__ void timer_cost(unsigned long long* out) {
    unsigned long long t0 = gtime(), s = 0;
    long long c0 = clock64();
    for (int i = 0; i < 1000; ++i) s += gtime();
    long long c1 = clock64();
    out[0] = gtime() - t0; out[1] = c1 - c0; out[2] = s;
}

int main() {
    uint64_t* buf; unsigned long long* out; 
    CK(cudaMalloc(&buf, 1 << 24)); CK(cudaMalloc(&out, 64));
    unsigned long long h[4];
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d\n", nsm);
    timer_cost<<<1, 1>>>(out); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost));
    printf("globaltimer read: %.1f ns, %.1f cycles each\n", h[0] / 1000.0, h[1] / 1000.0);
    const int iters = 4000;
    for (int peer : {1, 2, 37, 74, 100, 147}) {
        CK(cudaMemset(buf, 0, 1 << 16));
        void* args[] = {&buf, (void*)&peer, (void*)&iters, &out};
        CK(cudaLaunchCooperativeKernel((void*)pingpong, dim3(nsm), dim3(32), args, 0, 0));
        CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
        printf("pingpong cta0<->cta%d: round trip %.1f ns -> hop %.1f ns\n", peer, (double)h[0] / iters, (double)h[0] / iters / 2);
    }
    for (int work : {0, 1})
    for (int R : {1, 4})
    for (int U : {8, 2}) {
        int nprod = 128; 
        CK(cudaMemset(buf, 0, 1 << 22));
        void* args[] = {&buf, &nprod, &U, (void*)&R, (void*)&iters, (void*)&work, &out};
        CK(cudaLaunchCooperativeKernel((void*)allgather<512>, dim3(nsm), dim3(512), args, 0, 0));
        CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
        printf("allgather 148 CTAs x512 thr, %d producers x %d words, %d replicas, work=%d: %.1f ns / stage\n", nprod, U, R, work, (double)h[0] / iters);
    }
    {
        dsmem_pingpong<<<2, 32>>>(iters, out); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
        printf("DSMEM LL pingpong (cluster of 2): hop %.1f ns\n", (double)h[0] / iters / 2);
    }
    return 0;
}
