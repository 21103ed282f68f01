// LL hop variants: which load/store flavour and polling pattern gives the shortest cross-CTA hop through L2?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

template <int V> __device__ __forceinline__ void st64(uint64_t* p, uint64_t w) {
    if (V == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    if (V == 1) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    if (V == 2) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    if (V == 3) asm volatile("st.global.wt.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    if (V == 4) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    if (V == 5) { unsigned long long old; asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(w) : "memory"); }
    if (V == 6) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
template <int V> __device__ __forceinline__ uint64_t ld64(const uint64_t* p) {
    uint64_t w;
    if (V == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (V == 1) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (V == 2) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (V == 3) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (V == 4) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (V == 5) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}

template <int SV, int LV>
__global__ void pingpong(uint64_t* buf, int peer, int iters, unsigned long long* out) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me != 0 && me != peer) return;
    uint64_t* mine = buf + (me == 0 ? 0 : 64);
    uint64_t* theirs = buf + (me == 0 ? 64 : 0);
    unsigned long long t0 = gtime();
    for (int i = 1; i <= iters; ++i) {
        const uint64_t w = ((uint64_t)i << 32) | 7u;
        if (me == 0) {
            st64<SV>(theirs, w);
            while ((uint32_t)(ld64<LV>(mine) >> 32) != (uint32_t)i) {}
        } else {
            while ((uint32_t)(ld64<LV>(mine) >> 32) != (uint32_t)i) {}
            st64<SV>(theirs, w);
        }
    }
    if (me == 0) out[0] = gtime() - t0;
}

// polling by NW warps (lane 0 of each), staggered naturally; the first to see the tag raises a shared flag
template <int NW>
__global__ void pingpong_mw(uint64_t* buf, int peer, int iters, unsigned long long* out) {
    __shared__ volatile int seen;
    const int me = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (me != 0 && me != peer) return;
    if (lane != 0) return;
    uint64_t* mine = buf + (me == 0 ? 0 : 64);
    uint64_t* theirs = buf + (me == 0 ? 64 : 0);
    if (threadIdx.x == 0) seen = 0;
    __threadfence_block();
    // all polling warps of a CTA spin; warp 0 is the one that stores
    unsigned long long t0 = gtime();
    for (int i = 1; i <= iters; ++i) {
        const uint64_t w = ((uint64_t)i << 32) | 7u;
        if (me == 0 && warp == 0) st64<0>(theirs, w);
        while (seen < i) {
            if ((uint32_t)(ld64<0>(mine) >> 32) == (uint32_t)i) seen = i;
        }
        if (me != 0 && warp == 0) st64<0>(theirs, w);
    }
    if (me == 0 && warp == 0) out[0] = gtime() - t0;
}

// store-to-visibility: CTA 0 stores and then immediately polls the same word itself (own store round trip)
__global__ void self_rt(uint64_t* buf, int iters, unsigned long long* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long t0 = gtime();
    for (int i = 1; i <= iters; ++i) {
        st64<0>(buf, ((uint64_t)i << 32));
        while ((uint32_t)(ld64<0>(buf) >> 32) != (uint32_t)i) {}
    }
    out[0] = gtime() - t0;
}
// plain dependent load latency through L2 (pointer chase with ld.relaxed.gpu)
__global__ void chase(uint64_t* buf, int iters, unsigned long long* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t idx = 0;
    unsigned long long t0 = gtime();
    for (int i = 0; i < iters; ++i) idx = ld64<0>(buf + idx);
    out[0] = gtime() - t0; out[1] = idx;
}

template <int SV, int LV> void run_pp(const char* name, uint64_t* buf, unsigned long long* out, int nsm) {
    const int iters = 4000; unsigned long long h;
    double best = 1e9, sum = 0; int n = 0;
    for (int peer : {1, 2, 37, 74, 100, 147}) {
        CK(cudaMemset(buf, 0, 4096));
        void* args[] = {&buf, (void*)&peer, (void*)&iters, &out};
        CK(cudaLaunchCooperativeKernel((void*)pingpong<SV, LV>, dim3(nsm), dim3(32), args, 0, 0));
        CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
        double hop = (double)h / iters / 2; sum += hop; ++n; if (hop < best) best = hop;
    }
    printf("%-44s hop mean %.1f ns, best %.1f ns\n", name, sum / n, best);
}
template <int NW> void run_mw(uint64_t* buf, unsigned long long* out, int nsm) {
    const int iters = 4000; unsigned long long h; double sum = 0; int n = 0;
    for (int peer : {1, 37, 74, 147}) {
        CK(cudaMemset(buf, 0, 4096));
        void* args[] = {&buf, (void*)&peer, (void*)&iters, &out};
        CK(cudaLaunchCooperativeKernel((void*)pingpong_mw<NW>, dim3(nsm), dim3(32 * NW), args, 0, 0));
        CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
        sum += (double)h / iters / 2; ++n;
    }
    printf("polling by %d warps: hop mean %.1f ns\n", NW, sum / n);
}

int main() {
    uint64_t* buf; unsigned long long* out;
    CK(cudaMalloc(&buf, 1 << 24)); CK(cudaMalloc(&out, 64));
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long h[2];
    CK(cudaMemset(buf, 0, 1 << 24));
    chase<<<1, 32>>>(buf, 4000, out); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
    printf("dependent ld.relaxed.gpu (same word, L2): %.1f ns\n", (double)h[0] / 4000);
    self_rt<<<1, 32>>>(buf, 4000, out); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
    printf("own store -> own poll sees it: %.1f ns\n", (double)h[0] / 4000);
    run_pp<0, 0>("st.relaxed.gpu / ld.relaxed.gpu", buf, out, nsm);
    run_pp<1, 1>("st.volatile / ld.volatile", buf, out, nsm);
    run_pp<2, 2>("st.cg / ld.cg", buf, out, nsm);
    run_pp<3, 3>("st.wt / ld.cv", buf, out, nsm);
    run_pp<4, 4>("st.relaxed.sys / ld.relaxed.sys", buf, out, nsm);
    run_pp<5, 0>("atom.exch / ld.relaxed.gpu", buf, out, nsm);
    run_pp<6, 5>("st.release.gpu / ld.acquire.gpu", buf, out, nsm);
    run_pp<0, 1>("st.relaxed.gpu / ld.volatile", buf, out, nsm);
    run_pp<1, 0>("st.volatile / ld.relaxed.gpu", buf, out, nsm);
    run_mw<1>(buf, out, nsm); run_mw<2>(buf, out, nsm); run_mw<4>(buf, out, nsm); run_mw<8>(buf, out, nsm);
    return 0;
}
