// Per-step data gradients of the two decoder LSTM cells on the tensor cores (tcgen05, accumulators in TMEM):
//
//   out[b][n] += sum_k dZ[b][k] * W[k][n]        b < B <= 64 rows, K = 4096 gate rows, N = 2560 / 1792 inputs of the cell
//
// called twice per decode step of the reverse walk (train_api.cu), i.e. 1600 times per training step; the fp32 CUDA-core kernel
// (skinny_gemm_kernel, train_decoder.cu) re-reads 71 MB of fp32 weights per decode step at 17 TFLOP/s.  Here the weights are
// packed ONCE per backward pass as bf16 UMMA tiles (the values the mixed-precision forward multiplied, so the gradient is that
// of the function that was evaluated) and the activations enter as hi/lo bf16 pairs (rows b and 64 + b of the M = 128 operand,
// dZ = hi + lo to 2^-17 relative), products exact, accumulation fp32 in TMEM.
//
// BOTH operands arrive as ready-made UMMA tiles: the LSTM-cell backward kernel that produces dZ also writes it as hi/lo bf16
// tiles ([K / 64][128 rows x 64 k], K-major, no swizzle - lstm_cell_bwd_kernel, train_decoder.cu), so a stage is two bulk
// copies by the TMA engine and nothing passes through registers on the way in (the first version converted dZ in 16 loader warps:
// 96 registers x 576 threads, 197 KB of shared memory and three dependent L2 round trips per CTA):
//   warps 0-3  epilogue: TMEM (lane quarter) -> padded shared-memory tile, hi + lo rows added, COALESCED atomics onto the
//              destination (a warp = 128 consecutive bytes of one row; split-K)
//   warp 4     one lane streams the weight tile and the dZ tile of a stage with cp.async.bulk (16 KB each)
//   warp 5     one lane issues 4 tcgen05.mma.kind::f16 (K = 16 each) per stage
//   grid       (N / 128 column tiles) x (K splits); the 64 k-chunks are dealt as evenly as possible to floor(SMs / tiles) splits so
//              that the whole call is ONE wave
// 192 threads and 97 KB of shared memory per CTA: a CTA fits on an SM NEXT TO a CTA of the attention backward (512 threads, 121 KB
// at T = 150), so under programmatic dependent launch the neighbours of this kernel in the reverse walk set themselves up (weights,
// windows, location features; here: barriers, TMEM, the first weight tiles) while it runs.
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
constexpr int ST_STAGES = 3;
constexpr int ST_OFF = 1024;
constexpr int ST_TOTAL = ST_OFF + ST_STAGES * ST_STAGE;
constexpr int ST_EPI = 4;                      // epilogue warps (one per TMEM lane quarter)
constexpr int ST_THREADS = (ST_EPI + 2) * 32;
constexpr int ST_NACC = 4;
constexpr int ST_RED_LD = ST_N + 1;            // row stride of the epilogue tile: lanes = rows on the way in, = columns on the way out
static_assert(128 * ST_RED_LD * 4 <= ST_STAGES * ST_STAGE, "the epilogue tile reuses the pipeline stages");

struct SkinnySegs {
    float* out[3];
    int ldo[3];
    int n_end[3];
};
// one GEMM of a launch: activations (tile image), packed weights, destinations, number of 128-column tiles.  A launch carries one or
// two independent problems of the same K (the column tiles of the second follow those of the first in blockIdx.x): the two cells'
// data gradients of the reverse walk are adjacent and independent once the cell backward kernels are fused into the attention backward.
struct SkinnyProb {
    const __nv_bfloat16* Ap;
    const __nv_bfloat16* Wp;
    SkinnySegs so;
    int ntiles;
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
    skinny_tc_kernel(const SkinnyProb p0, const SkinnyProb p1, int B, int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + ST_STAGES;
    uint64_t* acc_bar = empty + ST_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    uint8_t* stages = smem + ST_OFF;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool second = (int)blockIdx.x >= p0.ntiles;
    const SkinnyProb& pr = second ? p1 : p0;
    const __nv_bfloat16* __restrict__ Ap = pr.Ap;          // dZ as hi/lo tiles [K / 64][128 x 64]
    const __nv_bfloat16* __restrict__ Wp = pr.Wp;
    const SkinnySegs& so = pr.so;
    const int nt = second ? blockIdx.x - p0.ntiles : blockIdx.x, n0 = nt * ST_N;
    const int kchunks = K / ST_K;
    const int c0 = (int)(((long long)blockIdx.y * kchunks) / gridDim.y), c1 = (int)(((long long)(blockIdx.y + 1) * kchunks) / gridDim.y);
    const int iters = c1 - c0;
    if (tid == 0) {
        for (int s = 0; s < ST_STAGES; ++s) {
            mbar_init(&full[s], 1);                          // the producer's expect_tx arrive (both tiles complete on it)
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(Wp) + ((size_t)nt * kchunks + c0) * ST_TILE;
    const uint8_t* asrc = reinterpret_cast<const uint8_t*>(Ap) + (size_t)c0 * ST_TILE;
    const int pre = min(iters, ST_STAGES);
    // the weight tiles of the first stages do not depend on the preceding kernel: they are in flight before the wait
    if (warp == ST_EPI && lane == 0) {
        for (int it = 0; it < pre; ++it) {
            mbar_arrive_expect_tx(&full[it], 2 * ST_TILE);
            bulk_g2s(stages + it * ST_STAGE + ST_TILE, wsrc + (size_t)it * ST_TILE, ST_TILE, &full[it]);
        }
    }
    __syncwarp();
    pdl_wait();           // dZ (and the consumed destinations) come from the preceding kernels: ONE convergent wait for the whole CTA
    // TMEM is allocated only now: a CTA that held its 512 columns while it waited for the predecessor cost 6 us per call (measured)
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    pdl_trigger();        // the next kernel of the chain may set itself up while this one works (run-ahead of one kernel)

    if (warp == ST_EPI) {
        if (lane == 0) {
            for (int it = 0; it < pre; ++it) bulk_g2s(stages + it * ST_STAGE, asrc + (size_t)it * ST_TILE, ST_TILE, &full[it]);
            for (int it = pre; it < iters; ++it) {
                const int stage = it % ST_STAGES;
                st_wait(&empty[stage], ((it / ST_STAGES) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&full[stage], 2 * ST_TILE);
                bulk_g2s(stages + stage * ST_STAGE + ST_TILE, wsrc + (size_t)it * ST_TILE, ST_TILE, &full[stage]);
                bulk_g2s(stages + stage * ST_STAGE, asrc + (size_t)it * ST_TILE, ST_TILE, &full[stage]);
            }
        }
    } else if (warp == ST_EPI + 1) {
        if (lane == 0) {
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
    }
    __syncwarp();

    // ---------------------------------------------------------------- epilogue (iters is uniform over the CTA)
    if (iters > 0) {
        float* red = reinterpret_cast<float*>(stages);                  // [128 hi|lo rows][ST_RED_LD]; every MMA has read its stage
        if (warp < ST_EPI) {
            st_wait(acc_bar, 0);
            umma::fence_after_sync();
            const int nacc = min(ST_NACC, iters * (ST_K / 16));
            float* rrow = red + (warp * 32 + lane) * ST_RED_LD;
#pragma unroll 1
            for (int c = 0; c < ST_N; c += 16) {
                float v[16];
                st_tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
                for (int a = 1; a < nacc; ++a) {
                    float u[16];
                    st_tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16) + a * ST_N + c, u);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += u[i];
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) rrow[c + i] = v[i];
            }
            umma::fence_before_sync();
        }
        __syncthreads();
        if (warp < ST_EPI) {
            const int seg = n0 < so.n_end[0] ? 0 : (n0 < so.n_end[1] ? 1 : 2);
            const int nb = seg == 0 ? 0 : so.n_end[seg - 1];
            float* dst = so.out[seg] + (n0 - nb);
            for (int b = warp; b < B; b += ST_EPI) {
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
// bytes of the hi/lo bf16 tile image of a [64][K] activation matrix (written by lstm_cell_bwd, read here)
size_t skinny_tc_act_bytes(int K) { return (size_t)(K / ST_K) * ST_TILE; }

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
// Xp: the activations as hi/lo bf16 tiles (skinny_tc_act_bytes; rows >= B zero).  skinny_gemm_pair_tc: two such problems in one launch.
static int make_prob(SkinnyProb& pr, const void* Xp, const void* Wp, int B, int K, int nseg, float* const* outs, const int* ldos, const int* n_ends,
                     const bool* acc, cudaStream_t st) {
    const int N = n_ends[nseg - 1];
    TACO_REQUIRE(skinny_tc_supported(B, K, N, nseg, n_ends), B200TACO_ERR_INVALID, "skinny_gemm_tc: unsupported shape B %d K %d N %d",
                 B, K, N);
    pr.Ap = static_cast<const __nv_bfloat16*>(Xp);
    pr.Wp = static_cast<const __nv_bfloat16*>(Wp);
    pr.ntiles = N / ST_N;
    for (int i = 0; i < 3; ++i) {
        const int j = i < nseg ? i : nseg - 1;
        pr.so.out[i] = outs[j]; pr.so.ldo[i] = ldos[j]; pr.so.n_end[i] = i < nseg ? n_ends[i] : N;
        if (i < nseg && !acc[i]) {
            const int nb = i == 0 ? 0 : n_ends[i - 1];
            TACO_CUDA(cudaMemset2DAsync(outs[i], (size_t)ldos[i] * sizeof(float), 0, (size_t)(n_ends[i] - nb) * sizeof(float), B, st));
        }
    }
    return B200TACO_OK;
}
static int launch_probs(const SkinnyProb& p0, const SkinnyProb& p1, int B, int K, cudaStream_t st) {
    const int ntiles = p0.ntiles + p1.ntiles;
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
    TACO_CUDA(launch_pdl(PDL_SKINNY_TC, skinny_tc_kernel, dim3(ntiles, ks), dim3(ST_THREADS), ST_TOTAL, st, p0, p1, B, K));
    TACO_LAUNCHED();
    return B200TACO_OK;
}
int skinny_gemm_multi_tc(const void* Xp, const void* Wp, int B, int K, int nseg, float* const* outs, const int* ldos,
                         const int* n_ends, const bool* acc, cudaStream_t st) {
    SkinnyProb p0{}, p1{};
    if (int e = make_prob(p0, Xp, Wp, B, K, nseg, outs, ldos, n_ends, acc, st)) return e;
    return launch_probs(p0, p1, B, K, st);
}
int skinny_gemm_pair_tc(const void* Xp0, const void* Wp0, int nseg0, float* const* outs0, const int* ldos0, const int* n_ends0, const bool* acc0,
                        const void* Xp1, const void* Wp1, int nseg1, float* const* outs1, const int* ldos1, const int* n_ends1, const bool* acc1,
                        int B, int K, cudaStream_t st) {
    SkinnyProb p0{}, p1{};
    if (int e = make_prob(p0, Xp0, Wp0, B, K, nseg0, outs0, ldos0, n_ends0, acc0, st)) return e;
    if (int e = make_prob(p1, Xp1, Wp1, B, K, nseg1, outs1, ldos1, n_ends1, acc1, st)) return e;
    return launch_probs(p0, p1, B, K, st);
}

}  // namespace b200taco
