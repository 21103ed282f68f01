// Dense layers on the 5th-generation tensor cores (tcgen05, accumulators in TMEM) at fp32 accuracy.
//
//   Y[n][co] = act( (sum_{tap,ci} X[n + tap - pad][ci] * W[tap][ci][co]) * scale[co] + shift[co] ) (+ R[n][co])
//
// the same contract as the SIMT kernel in dense.cu (reference src/model.py:25-39 encoder convolutions,
// :42-48 BiLSTM input projection, :112 memory layer, :166-171 prenet, :347-407 postnet).
//
// Precision: "3xTF32".  Every fp32 operand is split x = big + small with big = round-to-nearest TF32
// (10 explicit mantissa bits) and small = x - big (exact in fp32, itself read as TF32 by the tensor core):
//   x*w ~= big_x*big_w + small_x*big_w + big_x*small_w          (dropped term <= 2^-22 |x w|)
// accumulated in fp32 in TMEM.  Measured against an fp64 evaluation this is as accurate as the fp32 SIMT
// kernel (8e-8 vs 1.5e-7 on the encoder stack), so the parity tolerances do not move.
//
// One CTA computes a 128-row x NT-column output tile as an implicit GEMM over K = taps x c_in:
//   warps 0-7  A loaders: read X rows (shifted by the tap, zero outside the utterance), split, store the
//              big/small TF32 planes in the UMMA K-major core-matrix layout; warps 0-3 fill even pipeline
//              iterations, warps 4-7 odd ones.  Afterwards the same warps run the epilogue
//              (tcgen05.ld -> scale/shift/activation/residual -> global).
//   warp 8     W producer: one elected lane streams the pre-split, pre-tiled weight planes with
//              cp.async.bulk (TMA engine) into the same stage.
//   warp 9     MMA issuer: one elected lane issues 3 x (KC/8) tcgen05.mma.kind::tf32 per stage and
//              releases the stage with tcgen05.commit.
#include <cstdlib>

#include "handle.h"
#include "umma.cuh"

namespace b200taco {

namespace {

constexpr int TM = 128;            // rows per tile (UMMA M)
constexpr int KC = 16;             // input channels per pipeline stage
constexpr int KG = KC / 4;         // 16-byte k groups per stage (4 TF32 each)
constexpr int A_PLANE = TM * KC * 4;
constexpr int N_LOAD_WARPS = 8, W_WARP = 8, MMA_WARP = 9, NTHREADS = 320;

template <int NT>
struct Cfg {
    static constexpr int W_PLANE = NT * KC * 4;
    static constexpr int STAGE = 2 * A_PLANE + 2 * W_PLANE;
    static constexpr int STAGES = NT >= 256 ? 4 : 6;
    static constexpr int OFF_STAGE = 1024;
    static constexpr int TOTAL = OFF_STAGE + STAGES * STAGE;
    // The tensor core adds each MMA's products into the fp32 accumulator with truncation (measured: the error of
    // one long accumulation chain grows linearly with its length, ~1e-5 relative at K = 2560).  The K steps are
    // therefore dealt round-robin to NACC independent accumulators (all 512 TMEM columns) that the epilogue sums
    // with round-to-nearest adds: NACC times shorter chains, NACC times smaller bias.
    static constexpr int NACC = 512 / NT;
    static constexpr int TMEM_COLS = 512;
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
    static_assert(NT % 16 == 0 && NT <= 256, "UMMA N");
};

// 32-bit instruction descriptor for kind::tf32: D fp32, A/B TF32, both K-major, dense.
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// x = big + small, big = x rounded to nearest TF32 (ties away), small exact in fp32
__device__ __forceinline__ void split_tf32(float x, float& big, float& small) {
    big = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    small = x - big;
}

// byte offset of (row, k group) inside a K-major no-swizzle plane with KG k-groups: core matrix = 8 rows x 16 B
__device__ __forceinline__ int plane_off(int row, int kg) { return (((row >> 3) * KG + kg) << 7) + ((row & 7) << 4); }

template <int NT, int ACT>
__global__ void __launch_bounds__(NTHREADS, 1)
    conv_tc_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ Wtc, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ R, int ldres, float* __restrict__ Y, int ldy,
                   int n_rows, int seg_len, int c_in, int c_out, int taps, const ChannelMajorOut cm) {
    using C = Cfg<NT>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_bar = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    uint8_t* stages = smem + C::OFF_STAGE;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * TM, co0 = blockIdx.y * NT;
    const int pad = (taps - 1) / 2;
    const int kchunks = c_in / KC;
    const int iters = taps * kchunks;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], N_LOAD_WARPS / 2 + 1);      // 4 loader warps + the W producer's expect_tx arrive
            mbar_init(&empty[s], 1);                        // tcgen05.commit
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, C::TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp < N_LOAD_WARPS) {
        // ------------------------------------------------------------ A loaders
        const int grp = warp >> 2, w4 = warp & 3;           // grp 0: even iterations, grp 1: odd
        const int kg = lane >> 3;
        int rows[4], ts[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rows[j] = w4 * 32 + j * 8 + (lane & 7);
            const int gr = row0 + rows[j];
            ts[j] = gr < n_rows ? gr % seg_len : -(1 << 28);   // rows past the end never load
        }
        float4 v[4];
        auto fetch = [&](int it) {
            const int cic = it / taps, tap = it - cic * taps;
            const int sh = tap - pad;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int tt = ts[j] + sh;
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tt >= 0 && tt < seg_len)
                    v[j] = __ldg(reinterpret_cast<const float4*>(X + (size_t)(row0 + rows[j] + sh) * ldx + cic * KC + kg * 4));
            }
        };
        if (grp < iters) fetch(grp);
        for (int it = grp; it < iters; it += 2) {
            const int stage = it % STAGES;
            const uint32_t parity = ((it / STAGES) & 1u) ^ 1u;
            float4 big[4], small[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                split_tf32(v[j].x, big[j].x, small[j].x);
                split_tf32(v[j].y, big[j].y, small[j].y);
                split_tf32(v[j].z, big[j].z, small[j].z);
                split_tf32(v[j].w, big[j].w, small[j].w);
            }
            if (it + 2 < iters) fetch(it + 2);              // next tile of this group: loads in flight over the wait
            mbar_wait_spin(&empty[stage], parity);
            uint8_t* st = stages + stage * C::STAGE;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int off = plane_off(rows[j], kg);
                *reinterpret_cast<float4*>(st + off) = big[j];
                *reinterpret_cast<float4*>(st + A_PLANE + off) = small[j];
            }
            fence_proxy_async_smem();                       // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        }
    } else if (warp == W_WARP) {
        // ------------------------------------------------------------ W producer (TMA bulk copies)
        if (lane == 0) {
            const size_t plane_all = (size_t)c_out * KC;    // floats per (tap, chunk, plane)
            for (int it = 0; it < iters; ++it) {
                const int stage = it % STAGES;
                const uint32_t parity = ((it / STAGES) & 1u) ^ 1u;
                const int cic = it / taps, tap = it - cic * taps;
                mbar_wait_spin(&empty[stage], parity);
                mbar_arrive_expect_tx(&full[stage], 2 * C::W_PLANE);
                const float* src = Wtc + ((size_t)(tap * kchunks + cic) * 2) * plane_all + (size_t)co0 * KC;
                uint8_t* dst = stages + stage * C::STAGE + 2 * A_PLANE;
                bulk_g2s(dst, src, C::W_PLANE, &full[stage]);
                bulk_g2s(dst + C::W_PLANE, src + plane_all, C::W_PLANE, &full[stage]);
            }
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc_tf32(TM, NT);
            for (int it = 0; it < iters; ++it) {
                const int stage = it % STAGES;
                mbar_wait_spin(&full[stage], (it / STAGES) & 1u);
                umma::fence_after_sync();
                const uint32_t a_big = smem_u32(stages + stage * C::STAGE), a_small = a_big + A_PLANE;
                const uint32_t w_big = a_big + 2 * A_PLANE, w_small = w_big + C::W_PLANE;
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t da = umma::smem_desc(a_big + ks * 256, 128, KG * 128);
                    const uint64_t ds = umma::smem_desc(a_small + ks * 256, 128, KG * 128);
                    const uint64_t dw = umma::smem_desc(w_big + ks * 256, 128, KG * 128);
                    const uint64_t dv = umma::smem_desc(w_small + ks * 256, 128, KG * 128);
                    const int step = it * (KC / 8) + ks;
                    const uint32_t acc = tmem + (uint32_t)(step % C::NACC) * NT;
                    mma_tf32(acc, ds, dw, idesc, step >= C::NACC);       // small terms first
                    mma_tf32(acc, da, dv, idesc, true);
                    mma_tf32(acc, da, dw, idesc, true);
                }
                umma::commit(&empty[stage]);
            }
            umma::commit(acc_bar);
        }
        __syncwarp();
    }

    // ---------------------------------------------------------------- epilogue (warps 0-7)
    if (warp < N_LOAD_WARPS) {
        mbar_wait_spin(acc_bar, 0);
        umma::fence_after_sync();
        const int q = warp & 3, half = warp >> 2;                 // TMEM lane quarter, column half
        const int row = row0 + q * 32 + lane;
        constexpr int CH = NT / 2;
#pragma unroll 1
        for (int c = 0; c < CH; c += 8) {
            float v[8];
            tmem_ld_32x8(tmem + ((uint32_t)(q * 32) << 16) + half * CH + c, v);
            const int nacc = min(C::NACC, iters * (KC / 8));          // accumulators that received at least one step
            for (int a = 1; a < nacc; ++a) {
                float u[8];
                tmem_ld_32x8(tmem + ((uint32_t)(q * 32) << 16) + a * NT + half * CH + c, u);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += u[i];
            }
            const int co = co0 + half * CH + c;
            if (row < n_rows && co < c_out) {
                float sc[8], sh[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    sc[i] = scale ? __ldg(scale + co + i) : 1.f;
                    sh[i] = shift ? __ldg(shift + co + i) : 0.f;
                }
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float x = fmaf(v[i], sc[i], sh[i]);
                    if (ACT == ACT_RELU) x = fmaxf(x, 0.f);
                    if (ACT == ACT_TANH) x = tanhf(x);
                    o[i] = x;
                }
                if (R) {
                    const float4 r0 = __ldg(reinterpret_cast<const float4*>(R + (size_t)row * ldres + co));
                    const float4 r1 = __ldg(reinterpret_cast<const float4*>(R + (size_t)row * ldres + co + 4));
                    o[0] += r0.x; o[1] += r0.y; o[2] += r0.z; o[3] += r0.w;
                    o[4] += r1.x; o[5] += r1.y; o[6] += r1.z; o[7] += r1.w;
                }
                const int ub = row / seg_len, us = row - ub * seg_len;
                const bool keep = cm.lengths == nullptr || us < cm.lengths[ub];
                if (cm.zero_tail && !keep) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = 0.f;
                }
                float4* dst = reinterpret_cast<float4*>(Y + (size_t)row * ldy + co);
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                if (cm.y != nullptr && us < cm.ld) {
                    // second copy of the result in the vocoder's layout (B, c_out, cm.ld) (inference.py:71), zero from the
                    // utterance's own length on (train.py:71-79); consecutive lanes = consecutive frames: coalesced per channel
#pragma unroll
                    for (int i = 0; i < 8; ++i) cm.y[((size_t)ub * c_out + co + i) * cm.ld + us] = keep ? o[i] : 0.f;
                }
            }
        }
        umma::fence_before_sync();
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, C::TMEM_COLS);
}

// W[tap][ci][co] fp32 -> [tap][chunk][plane big|small][co/8][KG][8 co][4 ci]  (UMMA K-major core matrices)
__global__ void pack_dense_tc_kernel(float* __restrict__ dst, const float* __restrict__ w, int taps, int c_in, int c_out) {
    const size_t total = (size_t)taps * c_in * c_out;
    const int kchunks = c_in / KC;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int co = i % c_out, ci = (i / c_out) % c_in, tap = i / ((size_t)c_out * c_in);
        const int cic = ci / KC, k = ci % KC;
        float big, small;
        split_tf32(w[i], big, small);
        const size_t plane_all = (size_t)c_out * KC;
        const size_t base = ((size_t)(tap * kchunks + cic) * 2) * plane_all;
        const size_t off = (size_t)((co >> 3) * KG + (k >> 2)) * 32 + (co & 7) * 4 + (k & 3);
        dst[base + off] = big;
        dst[base + plane_all + off] = small;
    }
}

template <int NT>
int launch_nt(const float* X, int ldx, const float* Wtc, const float* scale, const float* shift, const float* R, int ldres,
              float* Y, int ldy, int n_rows, int seg_len, int c_in, int c_out, int taps, int act, cudaStream_t stream,
              const ChannelMajorOut& cm) {
    using C = Cfg<NT>;
    dim3 grid((n_rows + TM - 1) / TM, c_out / NT);
#define TACO_TC_CASE(A)                                                                                                 \
    {                                                                                                                   \
        auto kern = conv_tc_kernel<NT, A>;                                                                              \
        TACO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));                   \
        kern<<<grid, NTHREADS, C::TOTAL, stream>>>(X, ldx, Wtc, scale, shift, R, ldres, Y, ldy, n_rows, seg_len, c_in,  \
                                                   c_out, taps, cm);                                                    \
    }
    switch (act) {
        case ACT_NONE: TACO_TC_CASE(ACT_NONE) break;
        case ACT_RELU: TACO_TC_CASE(ACT_RELU) break;
        default: TACO_TC_CASE(ACT_TANH) break;
    }
#undef TACO_TC_CASE
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // namespace

int conv1d_channels_last_tc(const float* X, int ldx, const float* Wtc, const float* scale, const float* shift,
                            const float* residual, int ldres, float* Y, int ldy, int n_rows, int seg_len, int c_in,
                            int c_out, int taps, int act, int num_sms, cudaStream_t stream, const ChannelMajorOut& cm);

bool dense_tc_supported(int c_in, int c_out, int ldx, int ldy) {
    return c_in % KC == 0 && (c_out % 64 == 0 || c_out == 80) && ldx % 4 == 0 && ldy % 4 == 0;
}

size_t dense_tc_weight_floats(int taps, int c_in, int c_out) { return (size_t)2 * taps * c_in * c_out; }

int pack_dense_tc(float* dst, const float* w_tap_ci_co, int taps, int c_in, int c_out, cudaStream_t stream) {
    const size_t total = (size_t)taps * c_in * c_out;
    const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    pack_dense_tc_kernel<<<blocks, 256, 0, stream>>>(dst, w_tap_ci_co, taps, c_in, c_out);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

int dense_layer(const b200taco_handle* h, const float* X, int ldx, const float* W, const float* Wtc, const float* scale,
                const float* shift, const float* residual, int ldres, float* Y, int ldy, int n_rows, int seg_len, int c_in,
                int c_out, int taps, int act, bool precise, cudaStream_t stream, const ChannelMajorOut& cm) {
    // AUTO: layers whose output feeds the 1000-step decoder recurrence ("precise": encoder, memory layer) stay on the
    // fp32 CUDA cores (round-to-nearest accumulation); everything else runs on the tensor cores.  TC: all of them.
    const bool tc = h->dense_path == B200TACO_PATH_TC || (h->dense_path == B200TACO_PATH_AUTO && !precise);
    if (tc && Wtc != nullptr && dense_tc_supported(c_in, c_out, ldx, ldy))
        return conv1d_channels_last_tc(X, ldx, Wtc, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out,
                                       taps, act, precise ? 0 : h->num_sms, stream, cm);
    if (int e = conv1d_channels_last(X, ldx, W, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out, taps, act,
                                     stream))
        return e;
    // CUDA-core path: tail zeroing and the vocoder layout are separate passes over the result
    if (cm.zero_tail && cm.lengths != nullptr)
        if (int e = zero_tail_rows(Y, ldy, cm.lengths, n_rows / seg_len, seg_len, c_out, stream)) return e;
    if (cm.y != nullptr) return vocoder_layout(Y, ldy, cm.lengths, n_rows / seg_len, seg_len, c_out, cm.ld, cm.y, stream);
    return B200TACO_OK;
}

int conv1d_channels_last_tc(const float* X, int ldx, const float* Wtc, const float* scale, const float* shift,
                            const float* residual, int ldres, float* Y, int ldy, int n_rows, int seg_len, int c_in,
                            int c_out, int taps, int act, int num_sms, cudaStream_t stream, const ChannelMajorOut& cm) {
    TACO_REQUIRE(dense_tc_supported(c_in, c_out, ldx, ldy), B200TACO_ERR_INVALID,
                 "conv1d (tensor-core path): unsupported shape c_in %d c_out %d ldx %d ldy %d", c_in, c_out, ldx, ldy);
    if (n_rows <= 0) return B200TACO_OK;
    const int mtiles = (n_rows + TM - 1) / TM;
    // num_sms == 0 ("precise"): always the narrowest tile = 16 accumulators; used for the encoder, whose output
    // feeds a 1000-step recurrence (measured |error| vs fp32 SIMT: 8e-7 at NT=64, 3e-6 at NT=256)
    if (c_out == 80)
        return launch_nt<80>(X, ldx, Wtc, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out, taps, act, stream, cm);
    // measurement knob: TACO_DENSE_NT=64|128|256 forces the tile width (narrower tile = more accumulators)
    static const int forced = [] { const char* e = getenv("TACO_DENSE_NT"); return e ? atoi(e) : 0; }();
    if (forced == 256 && c_out % 256 == 0)
        return launch_nt<256>(X, ldx, Wtc, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out, taps, act, stream, cm);
    if (forced == 128 && c_out % 128 == 0)
        return launch_nt<128>(X, ldx, Wtc, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out, taps, act, stream, cm);
    if (forced == 32 && c_out % 32 == 0)
        return launch_nt<32>(X, ldx, Wtc, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out, taps, act, stream, cm);
    if (forced == 64)
        return launch_nt<64>(X, ldx, Wtc, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out, taps, act, stream, cm);
    // wide tiles when there is enough work to fill the machine, narrow ones to spread small problems over more SMs
    if (num_sms > 0 && c_out % 256 == 0 && mtiles * (c_out / 256) >= num_sms)
        return launch_nt<256>(X, ldx, Wtc, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out, taps, act, stream, cm);
    if (num_sms > 0 && c_out % 128 == 0 && mtiles * (c_out / 128) >= num_sms)
        return launch_nt<128>(X, ldx, Wtc, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out, taps, act, stream, cm);
    if (num_sms > 0 || c_out % 32 != 0)
        return launch_nt<64>(X, ldx, Wtc, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out, taps, act, stream, cm);
    return launch_nt<32>(X, ldx, Wtc, scale, shift, residual, ldres, Y, ldy, n_rows, seg_len, c_in, c_out, taps, act, stream, cm);
}

}  // namespace b200taco
