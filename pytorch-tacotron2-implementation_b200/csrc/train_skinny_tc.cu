// Per-step data gradients of the two decoder LSTM cells on the tensor cores (tcgen05, accumulators in TMEM):
//
//   out[b][n] += sum_k dZ[b][k] * W[k][n]        b < B <= 64 rows, K = 4096 gate rows, N = 2560 / 1792 inputs of the cell
//
// called twice per decode step of the reverse walk (train_api.cu), i.e. 1600 times per training step; the fp32 CUDA-core kernel
// (skinny_gemm_kernel, train_decoder.cu) re-reads 71 MB of fp32 weights per decode step at 17 TFLOP/s.  Here the weights are
// packed ONCE per backward pass as bf16 UMMA tiles (the values the mixed-precision forward multiplied, so the gradient is that
// of the function that was evaluated) and the activations enter as hi/lo bf16 pairs (rows b and 64 + b of the M = 128 operand,
// dZ = hi + lo to 2^-17 relative), products exact, accumulation fp32 in TMEM:
//   warps 0-3  A loaders: dZ rows (k contiguous) -> hi / lo bf16 -> K-major no-swizzle core matrices
//   warp 4     one elected lane streams the pre-packed weight tile (16 KB per stage) with cp.async.bulk
//   warp 5     one elected lane issues 4 tcgen05.mma.kind::f16 (K = 16 each) per stage
//   epilogue   all 16 loader warps read TMEM (lane quarter x 32-column group) into a padded shared-memory tile, hi and lo rows are
//              added there, and the 64 x 128 result goes onto the destination with COALESCED atomics (a warp = 128 consecutive
//              bytes of one row).  The first version added hi and lo separately with one row per lane: 2.6 M scattered atomics
//              per call were most of its 45 us.
//   grid       (N / 128 column tiles) x (K splits); the 64 k-chunks are dealt as evenly as possible to floor(SMs / tiles) splits so
//              that the whole call is ONE wave (it was 160 CTAs = two waves for the 2560-column cell)
#include <cstdlib>

#include "handle.h"
#include "umma.cuh"

namespace b200taco {

namespace {

constexpr int ST_N = 128;                     // weight columns per tile (UMMA N)
constexpr int ST_K = 64;                      // k per stage
constexpr int ST_KG = ST_K / 8;               // 16-byte k groups (8 bf16) per stage
constexpr int ST_TILE = 128 * ST_K * 2;       // bytes of one operand tile (A: 128 hi/lo rows; B: 128 columns)
constexpr int ST_STAGE = 2 * ST_TILE;
constexpr int ST_STAGES = 6;
constexpr int ST_OFF = 1024;
constexpr int ST_TOTAL = ST_OFF + ST_STAGES * ST_STAGE;
constexpr int ST_GROUPS = 4;                   // loader groups: group g fills the stages it = g (mod 4): four stages of loads in flight
constexpr int ST_LOADERS = 4 * ST_GROUPS;
constexpr int ST_THREADS = (ST_LOADERS + 2) * 32;
constexpr int ST_NACC = 4;
// A loader group goes from iteration it to it + ST_GROUPS after the stage of it - ST_STAGES was released; the parity wait on the next
// stage is only unambiguous if the release of it + ST_GROUPS - 2 ST_STAGES has happened by then, i.e. ST_GROUPS <= ST_STAGES (with 4
// groups on 3 stages a group ran two phases ahead, passed the wait on the "preceding phase" and over-arrived on the full barrier).
static_assert(ST_GROUPS <= ST_STAGES, "loader groups may not outnumber the pipeline stages");
constexpr int ST_RED_LD = ST_N + 1;            // row stride of the epilogue tile: lanes = rows on the way in, = columns on the way out
static_assert(128 * ST_RED_LD * 4 <= ST_STAGES * ST_STAGE, "the epilogue tile reuses the pipeline stages");

struct SkinnySegs {
    float* out[3];
    int ldo[3];
    int n_end[3];
};

__device__ __forceinline__ void st_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void st_tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// W fp32 [K][N] (ldw) -> bf16 tiles [N / 128][K / 64][128 n x 64 k] in the UMMA K-major no-swizzle layout
__global__ void pack_skinny_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ W, int ldw, int K, int N) {
    const size_t total = (size_t)K * N;
    const int kchunks = K / ST_K;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % N), k = (int)(i / N);
        const size_t tile = (size_t)(n / ST_N) * kchunks + k / ST_K;
        dst[tile * (128 * ST_K) + umma::tile_offset(n % ST_N, k % ST_K, ST_KG)] = __float2bfloat16_rn(W[(size_t)k * ldw + n]);
    }
}

__global__ void __launch_bounds__(ST_THREADS, 1)
    skinny_tc_kernel(const float* __restrict__ X, int ldx, const __nv_bfloat16* __restrict__ Wp, const SkinnySegs so, int B, int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + ST_STAGES;
    uint64_t* acc_bar = empty + ST_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    uint8_t* stages = smem + ST_OFF;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = blockIdx.x, n0 = nt * ST_N;
    const int kchunks = K / ST_K;
    const int c0 = (int)(((long long)blockIdx.y * kchunks) / gridDim.y), c1 = (int)(((long long)(blockIdx.y + 1) * kchunks) / gridDim.y);
    const int k0 = c0 * ST_K;
    const int iters = c1 - c0;

    pdl_trigger();
    if (tid == 0) {
        for (int s = 0; s < ST_STAGES; ++s) {
            mbar_init(&full[s], 4 + 1);                      // four loader warps + the weight producer's expect_tx arrive
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp < ST_LOADERS) {
        const int grp = warp >> 2, w4 = warp & 3;
        // ------------------------------------------------------------ A loaders: item = (row b, k group g): 8 consecutive k
        // lane = row b & 7 | ... so that the 8 lanes of a quarter warp write one 128-byte core matrix
        float4 va[4][2];
        auto fetch = [&](int it) {
            const int kb = k0 + it * ST_K;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int item = j * 128 + w4 * 32 + lane;            // 512 items = 64 rows x 8 k groups
                const int b = (item & 7) | ((item >> 6) << 3), g = (item >> 3) & 7;
                va[j][0] = va[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b < B) {
                    const float4* p = reinterpret_cast<const float4*>(X + (size_t)b * ldx + kb + 8 * g);
                    va[j][0] = p[0];
                    va[j][1] = p[1];
                }
            }
        };
        pdl_wait();               // dZ comes from the preceding kernel; barriers, TMEM and the first weight tiles are set up by now
        if (grp < iters) fetch(grp);
        for (int it = grp; it < iters; it += ST_GROUPS) {
            const int stage = it % ST_STAGES;
            const uint32_t parity = ((it / ST_STAGES) & 1u) ^ 1u;
            uint4 hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x[8] = {va[j][0].x, va[j][0].y, va[j][0].z, va[j][0].w, va[j][1].x, va[j][1].y, va[j][1].z, va[j][1].w};
                uint32_t h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    umma::split_bf16(x[2 * e], h0, l0);
                    umma::split_bf16(x[2 * e + 1], h1, l1);
                    h[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    l[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                }
                hi[j] = make_uint4(h[0], h[1], h[2], h[3]);
                lo[j] = make_uint4(l[0], l[1], l[2], l[3]);
            }
            if (it + ST_GROUPS < iters) fetch(it + ST_GROUPS);
            st_wait(&empty[stage], parity);
            uint8_t* st = stages + stage * ST_STAGE;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int item = j * 128 + w4 * 32 + lane;
                const int b = (item & 7) | ((item >> 6) << 3), g = (item >> 3) & 7;
                *reinterpret_cast<uint4*>(st + umma::tile_offset(b, 8 * g, ST_KG) * 2) = hi[j];
                *reinterpret_cast<uint4*>(st + umma::tile_offset(64 + b, 8 * g, ST_KG) * 2) = lo[j];
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        }
    } else if (warp == ST_LOADERS) {
        if (lane == 0) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(Wp) + ((size_t)nt * kchunks + k0 / ST_K) * ST_TILE;
            for (int it = 0; it < iters; ++it) {
                const int stage = it % ST_STAGES;
                const uint32_t parity = ((it / ST_STAGES) & 1u) ^ 1u;
                st_wait(&empty[stage], parity);
                mbar_arrive_expect_tx(&full[stage], ST_TILE);
                bulk_g2s(stages + stage * ST_STAGE + ST_TILE, src + (size_t)it * ST_TILE, ST_TILE, &full[stage]);
            }
        }
    } else if (lane == 0) {
        constexpr uint32_t idesc = umma::instr_desc_bf16_f32(128, ST_N);
        for (int it = 0; it < iters; ++it) {
            const int stage = it % ST_STAGES;
            st_wait(&full[stage], (it / ST_STAGES) & 1u);
            umma::fence_after_sync();
            const uint32_t a = smem_u32(stages + stage * ST_STAGE), w = a + ST_TILE;
#pragma unroll
            for (int ks = 0; ks < ST_K / 16; ++ks) {
                const int step = it * (ST_K / 16) + ks;
                const uint32_t acc = tmem + (uint32_t)(step % ST_NACC) * ST_N;
                umma::mma_bf16(acc, umma::smem_desc(a + ks * 256, 128, ST_KG * 128), umma::smem_desc(w + ks * 256, 128, ST_KG * 128), idesc,
                               step >= ST_NACC);
            }
            umma::commit(&empty[stage]);
        }
        umma::commit(acc_bar);
    }
    __syncwarp();

    // ---------------------------------------------------------------- epilogue (iters is uniform over the CTA)
    if (iters > 0) {
        float* red = reinterpret_cast<float*>(stages);                  // [128 hi|lo rows][ST_RED_LD]; every MMA has read its stage
        if (warp < ST_LOADERS) {
            st_wait(acc_bar, 0);
            umma::fence_after_sync();
            const int q = warp & 3, cg = warp >> 2;                     // TMEM lane quarter, 32-column group
            const int nacc = min(ST_NACC, iters * (ST_K / 16));
            float* rrow = red + (q * 32 + lane) * ST_RED_LD + cg * 32;
#pragma unroll 1
            for (int hc = 0; hc < 32; hc += 16) {
                float v[16];
                st_tmem_ld_32x16(tmem + ((uint32_t)(q * 32) << 16) + cg * 32 + hc, v);
                for (int a = 1; a < nacc; ++a) {
                    float u[16];
                    st_tmem_ld_32x16(tmem + ((uint32_t)(q * 32) << 16) + a * ST_N + cg * 32 + hc, u);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += u[i];
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) rrow[hc + i] = v[i];
            }
            umma::fence_before_sync();
        }
        __syncthreads();
        if (warp < ST_LOADERS) {
            const int seg = n0 < so.n_end[0] ? 0 : (n0 < so.n_end[1] ? 1 : 2);
            const int nb = seg == 0 ? 0 : so.n_end[seg - 1];
            float* dst = so.out[seg] + (n0 - nb);
            for (int b = warp; b < B; b += ST_LOADERS) {
#pragma unroll
                for (int c = 0; c < ST_N; c += 32)
                    atomicAdd(dst + (size_t)b * so.ldo[seg] + c + lane, red[b * ST_RED_LD + c + lane] + red[(64 + b) * ST_RED_LD + c + lane]);
            }
        }
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

size_t skinny_tc_packed_bytes(int K, int N) { return (size_t)K * N * sizeof(__nv_bfloat16); }

bool skinny_tc_supported(int B, int K, int N, int nseg, const int* n_ends) {
    static const bool allow = [] { const char* e = getenv("TACO_SKINNY_TC"); return !(e && e[0] == '0'); }();
    if (!allow || B > 64 || K % ST_K != 0 || N % ST_N != 0 || nseg < 1 || nseg > 3) return false;
    for (int i = 0; i < nseg; ++i)
        if (n_ends[i] % ST_N != 0) return false;
    return true;
}

int pack_skinny_tc(void* dst, const float* W, int ldw, int K, int N, cudaStream_t st) {
    pack_skinny_kernel<<<2048, 256, 0, st>>>(static_cast<__nv_bfloat16*>(dst), W, ldw, K, N);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

// out_i[b][n - n_begin_i] (+)= sum_k X[b][k] * W[k][n] for the column segments of one packed [K][N] matrix (see skinny_gemm_multi)
int skinny_gemm_multi_tc(const float* X, int ldx, const void* Wp, int B, int K, int nseg, float* const* outs, const int* ldos,
                         const int* n_ends, const bool* acc, cudaStream_t st) {
    const int N = n_ends[nseg - 1];
    TACO_REQUIRE(skinny_tc_supported(B, K, N, nseg, n_ends) && ldx % 4 == 0, B200TACO_ERR_INVALID, "skinny_gemm_tc: unsupported shape B %d K %d N %d",
                 B, K, N);
    SkinnySegs so{};
    for (int i = 0; i < 3; ++i) {
        const int j = i < nseg ? i : nseg - 1;
        so.out[i] = outs[j]; so.ldo[i] = ldos[j]; so.n_end[i] = i < nseg ? n_ends[i] : N;
        if (i < nseg && !acc[i]) {
            const int nb = i == 0 ? 0 : n_ends[i - 1];
            TACO_CUDA(cudaMemset2DAsync(outs[i], (size_t)ldos[i] * sizeof(float), 0, (size_t)(n_ends[i] - nb) * sizeof(float), B, st));
        }
    }
    const int ntiles = N / ST_N;
    static const int sms = [] { int d = 0, n = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n > 0 ? n : 148; }();
    // one wave: as many K splits as fit next to the column tiles, at least 4 k-chunks each (TACO_SKINNY_KS overrides: measurements)
    static const int forced_ks = [] { const char* e = getenv("TACO_SKINNY_KS"); return e ? atoi(e) : 0; }();
    int ks = forced_ks > 0 ? forced_ks : sms / ntiles;
    ks = ks < 1 ? 1 : (ks > K / (4 * ST_K) ? K / (4 * ST_K) : ks);
    if (ks < 1) ks = 1;
    static bool attr = false;
    if (!attr) {
        TACO_CUDA(cudaFuncSetAttribute(skinny_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_TOTAL));
        attr = true;
    }
    TACO_CUDA(launch_pdl(skinny_tc_kernel, dim3(ntiles, ks), dim3(ST_THREADS), ST_TOTAL, st, X, ldx, static_cast<const __nv_bfloat16*>(Wp), so, B, K));
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // namespace b200taco
