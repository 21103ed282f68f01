// Weight gradients of the dense / convolution layers on the 5th-generation tensor cores (tcgen05, accumulators in TMEM) at
// fp32 accuracy ("3xTF32", as dense_tc.cu):
//
//   dW[tap][ci][co] = sum_n X[n + tap - pad][ci] * dY[n][co]        (n over all rows; X rows outside the utterance are zero)
//
// the reduction runs over the ROWS (51,200 of them in a 64 x 800-frame training step), so both operands are "transposed"
// activations: A = X^T (M = 128 input channels, K = rows), B = dY^T (N = 128 output channels, K = rows).  One CTA owns a
// 128 x 128 tile of (ci, co) for one tap and a chunk of the rows (split-K over gridDim.z, partial tiles added with atomics):
//   warps 0-3  load A: lane = input channel, 4 consecutive rows per 16-byte k group (global reads coalesced along the channel),
//              split into big / small TF32 planes, stored in the UMMA K-major no-swizzle core-matrix layout
//   warps 4-7  the same for B (output channels)
//   warp 8     one elected lane issues 3 x (KC / 8) tcgen05.mma.kind::tf32 per stage (small x big, big x small, big x big) into
//              4 rotating TMEM accumulators (the tensor core truncates when it accumulates; shorter chains, summed with
//              round-to-nearest adds in the epilogue) and releases the stage with tcgen05.commit
// Replaces the fp32 CUDA-core reductions of train_dense.cu (22 TFLOP/s) where the shape allows (channels multiples of 128).
#include <cstdlib>

#include "handle.h"
#include "umma.cuh"

namespace b200taco {

namespace {

constexpr int WT = 128;                       // tile edge (UMMA M and N)
constexpr int WKC = 16;                       // rows per pipeline stage
constexpr int WKG = WKC / 4;                  // 16-byte k groups per stage
constexpr int WPLANE = WT * WKC * 4;          // bytes of one TF32 plane
constexpr int WSTAGE = 4 * WPLANE;            // A big, A small, B big, B small
constexpr int WSTAGES = 6;
constexpr int WOFF = 1024;
constexpr int WTOTAL = WOFF + WSTAGES * WSTAGE;
constexpr int WNACC = 4;                      // 4 x 128 columns = all of TMEM
constexpr int WLOADERS = 16;                  // two groups (even / odd stages) x (4 warps for A + 4 for B)
constexpr int WTHREADS = (WLOADERS + 1) * 32;
static_assert(WTOTAL <= 227 * 1024, "shared memory budget");

__host__ __device__ constexpr uint32_t w_instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void w_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void w_tmem_ld_32x8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void w_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void w_split(float x, float& big, float& small) {
    big = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    small = x - big;
}
__device__ __forceinline__ int w_plane_off(int row, int kg) { return (((row >> 3) * WKG + kg) << 7) + ((row & 7) << 4); }

template <bool CONV>
__global__ void __launch_bounds__(WTHREADS, 1)
    wgrad_tc_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ dY, int ldy, float* __restrict__ dW, int n_rows,
                    int seg_len, int c_in, int c_out, int taps, int rows_per, int atomic) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + WSTAGES;
    uint64_t* acc_bar = empty + WSTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    uint8_t* stages = smem + WOFF;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ci0 = blockIdx.x * WT, co0 = blockIdx.y * WT;
    const int tap = blockIdx.z % taps, chunk = blockIdx.z / taps;
    const int sh = tap - (taps - 1) / 2;
    const int r0 = chunk * rows_per, r1 = min(n_rows, r0 + rows_per);
    const int iters = (r1 - r0 + WKC - 1) / WKC;

    if (tid == 0) {
        for (int s = 0; s < WSTAGES; ++s) {
            mbar_init(&full[s], 8);                          // the eight loader warps
            mbar_init(&empty[s], 1);                         // tcgen05.commit
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp < WLOADERS) {
        // ------------------------------------------------------------ loaders.  Group 0 (warps 0-7) fills the even stages, group 1
        // the odd ones; inside a group warps 0-3 load operand A (X), warps 4-7 operand B (dY).  A lane owns one channel and reads
        // the 16 rows of a stage (each load coalesced along the channels); the loads of the group's next stage are in flight while
        // the current one is split and stored, so ~2-3 stages of global loads are outstanding per CTA.
        const int grp = warp >> 3;
        const bool isA = (warp & 4) == 0;
        const int m = (warp & 3) * 32 + lane;                   // channel inside the tile
        const int ld = isA ? ldx : ldy;
        const int shift = (CONV && isA) ? sh : 0;
        const float* src = (isA ? X + ci0 + m : dY + co0 + m) + (ptrdiff_t)shift * ld;
        float v[WKC];
        auto fetch = [&](int it) {
            const int rb = r0 + it * WKC;
            const float* p_ = src + (size_t)rb * ld;
            int t = CONV ? rb % seg_len + shift : 0;             // position of the row the tap reads, inside its utterance
#pragma unroll
            for (int k = 0; k < WKC; ++k) {
                bool ok = rb + k < r1;
                if (CONV) {
                    ok = ok && t >= 0 && t < seg_len;
                    if (++t - shift == seg_len) t = shift;       // next utterance
                }
                v[k] = ok ? __ldg(p_) : 0.f;
                p_ += ld;
            }
        };
        if (grp < iters) fetch(grp);
        for (int it = grp; it < iters; it += 2) {
            const int stage = it % WSTAGES;
            const uint32_t parity = ((it / WSTAGES) & 1u) ^ 1u;
            float4 big[WKG], small[WKG];
#pragma unroll
            for (int g = 0; g < WKG; ++g) {
                w_split(v[4 * g + 0], big[g].x, small[g].x);
                w_split(v[4 * g + 1], big[g].y, small[g].y);
                w_split(v[4 * g + 2], big[g].z, small[g].z);
                w_split(v[4 * g + 3], big[g].w, small[g].w);
            }
            if (it + 2 < iters) fetch(it + 2);               // this group's next stage: loads in flight over the wait
            w_wait(&empty[stage], parity);
            uint8_t* st = stages + stage * WSTAGE + (isA ? 0 : 2 * WPLANE);
#pragma unroll
            for (int g = 0; g < WKG; ++g) {
                const int off = w_plane_off(m, g);
                *reinterpret_cast<float4*>(st + off) = big[g];
                *reinterpret_cast<float4*>(st + WPLANE + off) = small[g];
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------ MMA issuer (warp 8)
        constexpr uint32_t idesc = w_instr_desc_tf32(WT, WT);
        for (int it = 0; it < iters; ++it) {
            const int stage = it % WSTAGES;
            w_wait(&full[stage], (it / WSTAGES) & 1u);
            umma::fence_after_sync();
            const uint32_t a_big = smem_u32(stages + stage * WSTAGE), a_small = a_big + WPLANE;
            const uint32_t b_big = a_big + 2 * WPLANE, b_small = b_big + WPLANE;
#pragma unroll
            for (int ks = 0; ks < WKC / 8; ++ks) {
                const uint64_t da = umma::smem_desc(a_big + ks * 256, 128, WKG * 128);
                const uint64_t ds = umma::smem_desc(a_small + ks * 256, 128, WKG * 128);
                const uint64_t db = umma::smem_desc(b_big + ks * 256, 128, WKG * 128);
                const uint64_t dv = umma::smem_desc(b_small + ks * 256, 128, WKG * 128);
                const int step = it * (WKC / 8) + ks;
                const uint32_t acc = tmem + (uint32_t)(step % WNACC) * WT;
                w_mma_tf32(acc, ds, db, idesc, step >= WNACC);          // small terms first
                w_mma_tf32(acc, da, dv, idesc, true);
                w_mma_tf32(acc, da, db, idesc, true);
            }
            umma::commit(&empty[stage]);
        }
        umma::commit(acc_bar);
    }
    __syncwarp();

    // ---------------------------------------------------------------- epilogue (warps 0-7): TMEM -> dW
    if (warp < 8 && iters > 0) {
        w_wait(acc_bar, 0);
        umma::fence_after_sync();
        const int q = warp & 3, half = warp >> 2;                 // TMEM lane quarter, column half
        const int ci = ci0 + q * 32 + lane;
        const int nacc = min(WNACC, iters * (WKC / 8));
        float* drow = dW + ((size_t)tap * c_in + ci) * c_out + co0 + half * (WT / 2);
#pragma unroll 1
        for (int c = 0; c < WT / 2; c += 8) {
            float v[8];
            w_tmem_ld_32x8(tmem + ((uint32_t)(q * 32) << 16) + half * (WT / 2) + c, v);
            for (int a = 1; a < nacc; ++a) {
                float u[8];
                w_tmem_ld_32x8(tmem + ((uint32_t)(q * 32) << 16) + a * WT + half * (WT / 2) + c, u);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += u[i];
            }
            if (atomic) {
#pragma unroll
                for (int i = 0; i < 8; ++i) atomicAdd(drow + c + i, v[i]);
            } else {
                *reinterpret_cast<float4*>(drow + c) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(drow + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        umma::fence_before_sync();
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

bool dense_wgrad_tc_supported(int n_rows, int c_in, int c_out) {
    static const bool allow = [] { const char* e = getenv("TACO_WGRAD_TC"); return !(e && e[0] == '0'); }();
    return allow && c_in % WT == 0 && c_out % WT == 0 && n_rows >= 512;
}

// dW[tap][ci][co] = sum_n X[n + tap - pad][ci] * dY[n][co]  (overwrites dW)
int dense_wgrad_tc(const float* X, int ldx, const float* dY, int ldy, float* dW, int n_rows, int seg_len, int c_in, int c_out, int taps,
                   int num_sms, cudaStream_t st) {
    TACO_REQUIRE(dense_wgrad_tc_supported(n_rows, c_in, c_out), B200TACO_ERR_INVALID, "dense_wgrad_tc: unsupported shape %d x %d x %d", n_rows,
                 c_in, c_out);
    const int tiles = (c_in / WT) * (c_out / WT) * taps;
    // row chunks: enough CTAs for two waves, and a chunk of X and dY rows that the tiles of one chunk re-read stays inside L2
    int chunks = 1;
    while ((tiles * chunks < 2 * num_sms || (size_t)(n_rows / chunks) * (c_in + c_out) * 4 > (48u << 20)) && n_rows / (chunks * 2) >= 256)
        chunks *= 2;
    const int rows_per = ((n_rows + chunks - 1) / chunks + WKC - 1) / WKC * WKC;
    chunks = (n_rows + rows_per - 1) / rows_per;
    TACO_REQUIRE(taps * chunks <= 65535, B200TACO_ERR_UNSUPPORTED, "dense_wgrad_tc: too many row chunks");
    if (chunks > 1) TACO_CUDA(cudaMemsetAsync(dW, 0, (size_t)taps * c_in * c_out * sizeof(float), st));
    const dim3 grid(c_in / WT, c_out / WT, taps * chunks);
    if (taps > 1) {
        TACO_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WTOTAL));
        wgrad_tc_kernel<true><<<grid, WTHREADS, WTOTAL, st>>>(X, ldx, dY, ldy, dW, n_rows, seg_len, c_in, c_out, taps, rows_per,
                                                              chunks > 1 ? 1 : 0);
    } else {
        TACO_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WTOTAL));
        wgrad_tc_kernel<false><<<grid, WTHREADS, WTOTAL, st>>>(X, ldx, dY, ldy, dW, n_rows, seg_len, c_in, c_out, taps, rows_per,
                                                               chunks > 1 ? 1 : 0);
    }
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // namespace b200taco
