// All-gather stage variants: what makes a 128-producer LL all-gather cost 900 ns when one hop costs 350?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void ll_st(uint64_t* p, float v, uint32_t tag) {
    const uint64_t w = ((uint64_t)tag << 32) | (uint64_t)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void ll_ld2(const uint64_t* p, uint64_t& a, uint64_t& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// nprod producers publish 8 words each into R replicas (stride words between producers); the first ncons CTAs (all CTAs
// if ncons == 0) gather all nprod * 8 words: thread t polls pair t (K-split pattern: 512 threads x 2 words = 1024 words).
// Non-consumers only produce; everybody waits for the round through ONE extra flag word published by consumer... to keep the
// stage serial without another all-gather, the producers wait for a single token that consumer CTA 0 (thread 0) publishes
// after its own gather completed.
__global__ void __launch_bounds__(512, 1) stage(uint64_t* buf, uint64_t* token, int nprod, int stride, int R, int ncons, int iters,
                                               int sleep_ns, unsigned long long* out) {
    const int tid = threadIdx.x, cta = blockIdx.x;
    const int slot_words = R * nprod * stride;
    const bool consumer = ncons == 0 || cta >= (int)gridDim.x - ncons;       // the LAST ncons CTAs consume (like the decoder's roles)
    const int rep = cta % R;
    unsigned long long t0 = gtime(), tmax = 0, tsum = 0;
    for (int i = 1; i <= iters; ++i) {
        uint64_t* X = buf + (size_t)(i & 3) * slot_words;
        // wait for the token of round i - 1 (everybody), then publish
        if (i > 1) {
            if (tid == 0) {
                uint64_t w;
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(token + (cta % 4) * 16) : "memory"); } while ((uint32_t)w != (uint32_t)(i - 1));
            }
            __syncthreads();
        }
        const unsigned long long tp = gtime();
        if (cta < nprod && tid < 8 * R) ll_st(X + (size_t)(tid / 8) * nprod * stride + cta * stride + (tid % 8), 1.f, i);
        if (consumer) {
            if (sleep_ns) __nanosleep(sleep_ns);
            const int word = 2 * tid;                        // pair index -> producer word / 8, offset word % 8
            const uint64_t* src = X + (size_t)rep * nprod * stride + (word / 8) * stride + (word % 8);
            uint64_t a, b;
            if (word < nprod * 8) {
                do { ll_ld2(src, a, b); } while ((uint32_t)(a >> 32) != (uint32_t)i || (uint32_t)(b >> 32) != (uint32_t)i);
            }
            __syncthreads();
            const unsigned long long dt = gtime() - tp;
            tsum += dt; if (dt > tmax) tmax = dt;
            if (cta == (int)gridDim.x - 1 && tid == 0) {
                for (int r = 0; r < 4; ++r) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(token + r * 16), "l"((uint64_t)i) : "memory");
            }
        }
    }
    if (cta == (int)gridDim.x - 1 && tid == 0) { out[0] = gtime() - t0; out[1] = tsum; out[2] = tmax; }
}

int main() {
    uint64_t *buf, *token; unsigned long long* out;
    CK(cudaMalloc(&buf, 1 << 24)); CK(cudaMalloc(&token, 4096)); CK(cudaMalloc(&out, 64));
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    const int iters = 3000;
    struct Cfg { int nprod, stride, R, ncons, sleep; const char* name; };
    Cfg cfgs[] = {
        {128, 8, 4, 0, 0, "128 prod x 64 B, 4 replicas, all 148 CTAs consume"},
        {128, 8, 4, 20, 0, "128 prod x 64 B, 4 replicas, 20 consumers"},
        {128, 8, 1, 20, 0, "128 prod x 64 B, 1 replica, 20 consumers"},
        {128, 16, 4, 20, 0, "128 prod, one 128 B line each, 4 replicas, 20 consumers"},
        {128, 16, 4, 0, 0, "128 prod, one 128 B line each, 4 replicas, all consume"},
        {128, 8, 4, 20, 300, "128 prod x 64 B, 4 replicas, 20 consumers, first poll after 300 ns"},
        {128, 8, 4, 1, 0, "128 prod x 64 B, 4 replicas, 1 consumer"},
        {8, 8, 4, 0, 0, "8 prod x 64 B (64 words), all consume"},
        {8, 8, 4, 20, 0, "8 prod x 64 B, 20 consumers"},
    };
    for (auto& c : cfgs) {
        CK(cudaMemset(buf, 0, 1 << 24)); CK(cudaMemset(token, 0, 4096));
        void* args[] = {&buf, &token, &c.nprod, &c.stride, &c.R, &c.ncons, (void*)&iters, &c.sleep, &out};
        CK(cudaLaunchCooperativeKernel((void*)stage, dim3(nsm), dim3(512), args, 0, 0));
        CK(cudaDeviceSynchronize());
        unsigned long long h[3]; CK(cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost));
        printf("%-70s round %.0f ns | last CTA: publish->gathered mean %.0f ns, max %.0f ns\n", c.name, (double)h[0] / iters, (double)h[1] / iters, (double)h[2]);
    }
    return 0;
}
