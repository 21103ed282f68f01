// Persistent Tacotron 2 decoder: the whole frame loop of Decoder.inference / Decoder.forward
// (reference src/model.py:277-333) in ONE cooperative kernel launch.
//
// Per decode step (Decoder._decode_step, src/model.py:246-275) the grid runs four phases
// separated by a hand-rolled gpu-scope barrier:
//   P1  attention LSTMCell   z = W[4096 x 1792] . [prenet; ctx; ah]      (128 "GEMM" CTAs, 32 gate rows each)
//       + each CTA's slice of the query projection Wq . ah (partials, summed in P2)
//   P2  location-sensitive attention (src/model.py:117-148), one work item per
//       (utterance, slice of encoder positions); partial softmax/context per slice, the last
//       slice to arrive (atomic ticket) merges them - no extra grid barrier
//   P3  decoder LSTMCell     z = W[4096 x 2560] . [ctx; ah; dh]  + each CTA's slice of the
//       mel/gate projection (partials); spare CTAs compute the context part of the projection
//   P4  per utterance: sum projection partials -> mel frame, gate logit, stop rule, prenet
//       of the next input frame
// LSTM weights stream from L2 (they fit: 35.7 MB bf16) into shared memory through a
// 4-stage mbarrier pipeline filled by TMA bulk copies (cp.async.bulk, SASS UBLKCP) issued by
// a dedicated producer warp; weight chunks of the next GEMM phase are prefetched across the
// grid barrier.  Cell state c never leaves the owning CTA; h, ctx and the prenet output are
// exchanged through the double-buffered "act" matrix [2816][BP] in global memory (L2).
#include <cooperative_groups.h>

#include <cstdlib>

#include "handle.h"
#include "umma.cuh"

namespace b200taco {

namespace {

constexpr int NW = 8;                       // consumer warps
constexpr int NTC = NW * 32;                // consumer threads
constexpr int NT = NTC + 32;                // + producer warp
// Producer lanes: a thread's bulk copies complete one at a time (~0.4 us each whatever the size, measured:
// profiles/r1_stream_probe.txt), so every pipeline stage has its own issuing lane (lane == stage) and the
// copies of different stages proceed in parallel; one lane per stage also keeps each stage's uses in order.
constexpr int TC_CK = 64;                   // k per pipeline chunk on the tensor-core path
constexpr int TC_KG = TC_CK / 8;
constexpr int GS = 4;                       // tensor-core path: CTAs per split-K group
constexpr int GN = GS * ROWS_PER_CTA;       // gate rows per group = N of one tcgen05.mma (128)
constexpr int TILE_POS = 40;                // encoder positions per attention tile (8 groups x 5)
constexpr int MAX_NS = 16;                  // max slices per utterance
constexpr int TC_MIN_BATCH = 4;             // smallest batch that takes the tensor-core kernel
constexpr int APART = 4 + EMB;              // per-slice partial: m, s, pad, pad, ctx[512]
constexpr uint64_t SPIN_TIMEOUT_NS = 4000000000ull;

// control words in scratch
constexpr int CTRL_BARRIER = 0, CTRL_STOP = 1, CTRL_FINISHED = 2, CTRL_TICKETS = 32, CTRL_DONE = 32 + MAX_BATCH,
              CTRL_LEN = 32 + 2 * MAX_BATCH, CTRL_PROF = 32 + 3 * MAX_BATCH /* 5 x u64 */,
              CTRL_GCNT = CTRL_PROF + 48 /* 32 group counters */,
              // utterance queue (continuous batching, KParams::n_utts > 0): queue head, then per decoder row the utterance
              // it carries (negative: -1 - last utterance, row retired) and the decode step at which that utterance began
              CTRL_QHEAD = CTRL_GCNT + GEMM_CTAS / GS, CTRL_UTT = CTRL_QHEAD + 1, CTRL_START = CTRL_UTT + MAX_BATCH,
              CTRL_WORDS = CTRL_START + MAX_BATCH;
constexpr int ROW_START = MAX_BATCH;       // offsets inside a CTA's shared copy of the row table: start steps,
constexpr int ROW_ABASE = 2 * MAX_BATCH;   // first alignment row of the utterance ((u / B) * max_steps * B + u % B)

// The persistent kernel's code (~165 KB) is far beyond the instruction caches and every warp walks it once per decode
// step (ncu: sm__icc_request_hit_rate 72-84 %, stall_no_instruction up to 1.1 per issued instruction): what runs more than
// once per step - MMA issue, the accumulator epilogue and the LSTM pointwise part of both cells - is kept as ONE out-of-line
// copy with rolled loops: 10.3 k -> 7.2 k instructions, decode step 31.5 -> 30.7 us at B=32 (the attention and projection
// phases, whose code did not change, got 0.8 us faster).
#ifndef TACO_COMPACT
#define TACO_COMPACT 1
#endif
#if TACO_COMPACT
#define TACO_NI __noinline__
#define TACO_ROLLED _Pragma("unroll 1")
#else
#define TACO_NI __forceinline__
#define TACO_ROLLED
#endif


struct KParams {
    int B, T, Tp, NS, SL, NQ, max_steps, mode, stop_mode;
    // continuous batching (0 = off): n_utts utterances of equal text length stream through the B decoder rows; a row whose
    // utterance stopped takes the next one from the queue at the following decode step.  memory / pm / mel / gate /
    // lengths are then indexed by utterance, align is (n_utts / B, max_steps, B, T).
    int n_utts;
    float thr;
    const void* w_attn;
    const void* w_dec;
    const float *bias_attn, *bias_dec, *wq_slices, *wproj_slices, *wproj_ctxT, *proj_bias, *prenet1T, *prenet2T,
        *loc_conv, *loc_denseT, *attn_v;
    const float *memory, *pm;
    const void* p_all;                 // teacher forcing: prenet outputs in the GEMM operand layout of the chosen path
    const void *w_attn_tc, *w_dec_tc;  // tcgen05 path: bf16 weights as UMMA K-major tiles
    __nv_bfloat16* act16;              // tcgen05 path: hi/lo bf16 activation tiles [2][44 chunks][2*BP rows x 64 k]
    const int64_t* text_lengths;
    float *mel, *gate, *align;
    int32_t *lengths, *steps_run;
    // scratch
    float *act, *pqpart, *melpart, *ctx_bj, *attn_prev, *attn_cum, *wtmp, *apart;
    float* zpart;                      // tcgen05 path: split-K partial gate sums [dst cta][src member][32 rows][BP]
    // training (teacher-forced) only, null otherwise: per-step state kept for the backward pass and the dropout masks
    // of the carried hidden states (src/model.py:262, :270).  Layouts [step][b][...]: gates = gate*1024 + unit (i,f,g,o
    // post-activation), c, h (after dropout), ctx; masks are 0/1 floats.
    float *sv_ga, *sv_ca, *sv_ah, *sv_gd, *sv_cd, *sv_dh, *sv_ctx;
    const float *mask_a, *mask_d;
    float drop_scale;
    uint32_t* ctrl;
};

template <int BP, int WBYTES = 2>
struct Tile {
    static constexpr int BPL = BP >= 64 ? 8 : (BP >= 4 ? 4 : BP);   // batch columns per lane
    static constexpr int LB = BP / BPL;                              // lanes along batch
    static constexpr int LK = 32 / (4 * LB);                         // lanes along k (4 lanes along rows)
    static constexpr int CK = (BP <= 4 ? 256 : (BP <= 16 ? 128 : 64)) * 2 / WBYTES;  // k rows per pipeline chunk
    static constexpr int KPW = CK / NW;                              // k rows per warp per chunk
    static constexpr int KPL = KPW / LK;                             // k rows per lane per chunk
    static_assert(LB * BPL == BP && LK * LB * 4 == 32 && KPL * LK == KPW, "tile");
};

// TC = tensor-core path (tcgen05, accumulators in TMEM): operands are bf16 K-major UMMA tiles,
// activations split hi/lo (rows b and BP+b of the M=128 operand), N = this CTA's 32 gate rows.
template <int BP, typename WT, bool TC>
struct Smem {
    using TL = Tile<BP, sizeof(WT)>;
    static constexpr int CK = TC ? TC_CK : TL::CK;
    static constexpr int NSTAGES = TC ? (BP <= 32 ? 4 : 3) : (BP >= 64 ? 4 : 6);
    static constexpr int WROWS = TC ? GN : ROWS_PER_CTA;                       // weight rows per chunk
    static constexpr int W_BYTES = CK * WROWS * (int)sizeof(WT);
    static constexpr int A_BYTES = TC ? 2 * BP * CK * 2 : CK * BP * 4;
    static constexpr int STAGE_BYTES = W_BYTES + A_BYTES;
    static_assert(!TC || (sizeof(WT) == 2 && BP >= 32), "tensor-core path: bf16 weights, batch tile >= 32");
    // persistent small arrays (floats)
    static constexpr int OFF_BAR = 0;                                          // full[], empty[], acc bar, tmem slot
    static constexpr int OFF_ACCBAR = 2 * NSTAGES * 8;
    static constexpr int OFF_TMEM = OFF_ACCBAR + 8;
    static_assert(OFF_TMEM + 4 <= 256, "barrier area");
    static constexpr int OFF_STAGE = 256;
    static constexpr int OFF_RED = OFF_STAGE + NSTAGES * STAGE_BYTES;          // NW*32*BP floats, aliased by P2/P4 scratch
    static constexpr int RED_BYTES_MIN = 48 * 1024;
    static constexpr int RED_SIMT = TC ? 0 : NW * 32 * BP * 4;
    static constexpr int RED_BYTES = RED_SIMT > RED_BYTES_MIN ? RED_SIMT : RED_BYTES_MIN;
    static constexpr int OFF_Z = OFF_RED + RED_BYTES;                          // 32*BP
    static constexpr int OFF_CA = OFF_Z + 32 * BP * 4;                         // 8*BP
    static constexpr int OFF_CD = OFF_CA + 8 * BP * 4;
    static constexpr int OFF_H = OFF_CD + 8 * BP * 4;
    static constexpr int OFF_BIAS = OFF_H + 12 * BP * 4;                      // h[8][BP] + this CTA's 4 context rows cx[4][BP]                        // 64
    static constexpr int OFF_WQ = OFF_BIAS + 64 * 4;                           // 128*8
    static constexpr int OFF_WPROJ = OFF_WQ + 128 * 8 * 4;                     // 81*8 (pad 656)
    static constexpr int OFF_WLOC = OFF_WPROJ + 984 * 4;                      // + ctx slice 81*4 (pad 328)                       // 32*63
    static constexpr int OFF_WLD = OFF_WLOC + 32 * 63 * 4;                     // 32*128
    static constexpr int OFF_V = OFF_WLD + 32 * 128 * 4;                       // 132
    static constexpr int OFF_MISC = OFF_V + 132 * 4;                           // 64 floats of flags/scalars
    static constexpr int OFF_XCH = OFF_MISC + 64 * 4;                          // tensor-core path: split-K partials of the
    static constexpr int XCH_BYTES = TC ? GS * ROWS_PER_CTA * BP * 4 : 0;      // four group members, written through DSMEM
    static constexpr int OFF_ROWS = OFF_XCH + XCH_BYTES;                       // tensor-core path: row table (utterance, start step)
    static constexpr int TOTAL = OFF_ROWS + (TC ? 3 * MAX_BATCH * 4 : 0);
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
    static_assert(STAGE_BYTES % 128 == 0 && W_BYTES % 128 == 0, "bulk copy alignment");
};

// element offset of (row m, activation index k) inside one act16 buffer: [chunk][2*BP rows x 64 k tile]
template <int BP>
__device__ __forceinline__ size_t act16_off(int m, int k) {
    return (size_t)(k / TC_CK) * (2 * BP * TC_CK) + umma::tile_offset(m, k % TC_CK, TC_KG);
}

// Scratch carved from the "red" region during P2 (floats)
struct P2Scratch {
    static constexpr int PQ = 0;                                  // 256 (two halves)
    static constexpr int WIN = 256;                               // 2 * WIN_LEN
    static constexpr int WIN_LEN = MAX_TEXT_LEN + 2 * TILE_POS + 32;
    static constexpr int E = WIN + 2 * WIN_LEN;                   // MAX_TEXT_LEN + TILE_POS
    static constexpr int LOC_LD = 36;                             // row stride of the location-feature tile (16 B aligned)
    static constexpr int LOC = E + MAX_TEXT_LEN + TILE_POS;       // TILE_POS * LOC_LD
    static constexpr int PM = LOC + TILE_POS * LOC_LD;            // TILE_POS * 128
    static constexpr int EPART = PM + TILE_POS * 128;             // TILE_POS * 4
    static constexpr int SC = EPART + TILE_POS * 4;               // MAX_NS scale factors + m,s
    static constexpr int TOTAL = SC + 64;
};
static_assert(P2Scratch::TOTAL * 4 <= 48 * 1024 && 128 * 33 * 4 <= 48 * 1024, "P2 scratch must fit the red region");

__device__ __forceinline__ void spin_fail() { __trap(); }

#ifdef TACO_PROFILE_FINE
#define TACO_FINE_ARGS t_fine, t_flast
#define FT_RESET() do { if (timing) t_flast = global_timer_ns(); } while (0)
#define FT(n) do { if (timing) { const unsigned long long _n = global_timer_ns(); t_fine[n] += _n - t_flast; t_flast = _n; } } while (0)
#else
#define TACO_FINE_ARGS nullptr, t_last
#define FT_RESET()
#define FT(n)
#endif

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = global_timer_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) spin_fail();
    }
}

// All NT threads of every CTA call this the same number of times.
// Ordering: every thread's writes happen-before thread 0's red.release (bar.sync + cumulativity);
// thread 0's ld.acquire happens-before every thread's later reads (second bar.sync).
__device__ __forceinline__ void grid_barrier(uint32_t* counter, uint32_t& epoch, uint32_t ncta) {
    __syncthreads();
    epoch += 1;
    if (threadIdx.x == 0) {
        fence_proxy_async();                       // generic writes -> visible to other CTAs' TMA reads
        red_release_add_u32(counter, 1u);
        const uint32_t target = epoch * ncta;
        if (ld_acquire_u32(counter) < target) {
            uint64_t t0 = global_timer_ns();
            while (ld_acquire_u32(counter) < target) {
                if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) spin_fail();
            }
        }
    }
    __syncthreads();
}

// sum of N values spaced `stride` floats apart, 16 loads in flight (L2 latency bound otherwise)
template <int N, int BATCH = 32>
__device__ __forceinline__ float strided_sum_cg(const float* src, int stride) {
    static_assert(N % BATCH == 0 && BATCH % 4 == 0, "N");
    float total = 0.f;
#pragma unroll 1
    for (int i0 = 0; i0 < N; i0 += BATCH) {
        float v[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) v[i] = __ldcg(src + (size_t)(i0 + i) * stride);
#pragma unroll
        for (int i = 0; i < BATCH; i += 4) total += (v[i] + v[i + 1]) + (v[i + 2] + v[i + 3]);
    }
    return total;
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t atom_add_acq_rel_gpu(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ void consumer_sync() { named_bar_sync(1, NTC); }

// where in the attention phase the background weight streaming of the tensor-core kernel may start:
// 0 = after the first loads, 1 = after the energies, 2 = after the context partial
#ifndef TACO_SIG_POINT
#define TACO_SIG_POINT 2
#endif

// ---- thread-block cluster helpers (tensor-core kernel: a split-K group is one cluster of GS CTAs) ----
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CTA-local signal in shared memory (written by consumer thread 0, polled by the producer warp)
__device__ __forceinline__ void wait_signal(const int* flag, int value) {
    const volatile int* f = flag;
    const uint64_t t0 = global_timer_ns();
    while (*f < value) {
        if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) spin_fail();
    }
}

template <typename WT>
__device__ __forceinline__ void load_w8(const WT* p, float (&w)[8]);
template <>
__device__ __forceinline__ void load_w8<__nv_bfloat16>(const __nv_bfloat16* p, float (&w)[8]) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    w[0] = bf16lo(v.x); w[1] = bf16hi(v.x); w[2] = bf16lo(v.y); w[3] = bf16hi(v.y);
    w[4] = bf16lo(v.z); w[5] = bf16hi(v.z); w[6] = bf16lo(v.w); w[7] = bf16hi(v.w);
}
template <>
__device__ __forceinline__ void load_w8<float>(const float* p, float (&w)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

template <int N>
__device__ __forceinline__ void load_a(const float* p, float (&a)[N]) {
    if constexpr (N == 1) {
        a[0] = p[0];
    } else if constexpr (N == 2) {
        float2 v = *reinterpret_cast<const float2*>(p);
        a[0] = v.x; a[1] = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < N; i += 4) {
            float4 v = *reinterpret_cast<const float4*>(p + i);
            a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Pipeline bookkeeping shared by the producer warp and the consumer warps: both walk the
// same deterministic sequence of chunks, so a running chunk counter gives stage and parity.
struct Pipe {
    uint64_t* full;
    uint64_t* empty;
    uint8_t* stages;
};

// Consumer side of one LSTM gate contraction: acc[r][j] += W[k][lr*8+r] * act[k][lb*BPL+j].
template <int BP, typename WT, bool ZERO = true>
__device__ __forceinline__ void gemm_consume(const Pipe& pipe, uint32_t& g, int nchunks, int warp, int lane,
                                             float (&acc)[8][Tile<BP>::BPL]) {
    using TL = Tile<BP, sizeof(WT)>;
    using SM = Smem<BP, WT, false>;
    constexpr int STAGES = SM::NSTAGES;
    const int lk = lane % TL::LK;
    const int lb = (lane / TL::LK) % TL::LB;
    const int lr = lane / (TL::LK * TL::LB);
    if constexpr (ZERO) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int j = 0; j < TL::BPL; ++j) acc[r][j] = 0.f;
    }
    for (int ch = 0; ch < nchunks; ++ch, ++g) {
        const uint32_t stage = g % STAGES, parity = (g / STAGES) & 1u;
        mbar_wait(&pipe.full[stage], parity);
#ifdef TACO_DBG_SLEEP
        __nanosleep(TACO_DBG_SLEEP);
#endif
        const WT* Ws = reinterpret_cast<const WT*>(pipe.stages + stage * SM::STAGE_BYTES);
        const float* As = reinterpret_cast<const float*>(pipe.stages + stage * SM::STAGE_BYTES + SM::W_BYTES);
#pragma unroll
        for (int j = 0; j < TL::KPL; ++j) {
            const int k = warp * TL::KPW + j * TL::LK + lk;
            float w[8], a[TL::BPL];
            load_w8<WT>(Ws + k * ROWS_PER_CTA + lr * 8, w);
            load_a<TL::BPL>(As + k * BP + lb * TL::BPL, a);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < TL::BPL; ++c) acc[r][c] = fmaf(w[r], a[c], acc[r][c]);
        }
        // the loads above are generic-proxy reads; the refill of this stage is an async-proxy
        // (TMA) write: without this fence the bulk copy can overtake loads still in flight
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pipe.empty[stage]);
    }
}

// Reduce the per-warp partial tiles: over the k-lanes by shuffle, over warps through `red`;
// result z[row][b] in shared memory.  Consumer warps only.
template <int BP>
__device__ __forceinline__ void gemm_reduce(float (&acc)[8][Tile<BP>::BPL], float* red, float* z, int warp, int lane,
                                            int tid) {
    using TL = Tile<BP>;
    const int lk = lane % TL::LK;
    const int lb = (lane / TL::LK) % TL::LB;
    const int lr = lane / (TL::LK * TL::LB);
#pragma unroll
    for (int o = 1; o < TL::LK; o <<= 1)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < TL::BPL; ++c) acc[r][c] += __shfl_xor_sync(0xffffffffu, acc[r][c], o);
    if (lk == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < TL::BPL; ++c)
                red[(warp * ROWS_PER_CTA + lr * 8 + r) * BP + lb * TL::BPL + c] = acc[r][c];
    }
    consumer_sync();
    for (int i = tid; i < ROWS_PER_CTA * BP; i += NTC) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[w * ROWS_PER_CTA * BP + i];
        z[i] = s;
    }
    consumer_sync();
}

// LSTMCell pointwise part (gate order i,f,g,o; src/model.py:195-203 via nn.LSTMCell).
// z rows are unit*4+gate; c state stays in shared memory of the owning CTA.
// training: where this step's gate activations / states go, and the dropout mask of the carried h
struct PwSave {
    float *gates, *c, *h;          // [b][4096], [b][1024], [b][1024] of this step (null: inference)
    const float* mask;             // [b][1024] of this step or null
    float scale;
    int unit0;                     // first hidden unit of this CTA
};
__device__ __forceinline__ PwSave pw_save(float* g, float* c, float* h, const float* mask, float scale, int step, int B,
                                          int unit0) {
    PwSave sv;
    sv.gates = g ? g + (size_t)step * B * 4 * RNN : nullptr;
    sv.c = c ? c + (size_t)step * B * RNN : nullptr;
    sv.h = h ? h + (size_t)step * B * RNN : nullptr;
    sv.mask = mask ? mask + (size_t)step * B * RNN : nullptr;
    sv.scale = scale;
    sv.unit0 = unit0;
    return sv;
}

template <int BP, bool TC>
__device__ __forceinline__ void lstm_pointwise(const float* z, const float* bias, float* c_state, float* h_out,
                                               float* act_dst /* &act[buf][row0][0] */, __nv_bfloat16* act16_buf,
                                               int k0, int B, int tid, const PwSave& sv, const int* rows = nullptr,
                                               int step = 0, bool zero_z = false) {
    // rows != null (continuous batching): a row that starts a new utterance at this step has zero carried state
    // (Decoder._initialize_decoder_states, src/model.py:219-234): cell state 0, and - zero_z - all inputs of this
    // contraction zero, whatever the accumulator gathered from the previous utterance in the background
    for (int i = tid; i < UNITS_PER_CTA * BP; i += NTC) {
        const int u = i / BP, b = i % BP;
        if (b < B) {
            const bool fresh = rows != nullptr && rows[ROW_START + b] == step;
            const float zs = fresh && zero_z ? 0.f : 1.f;
            const float zi = zs * z[(u * 4 + 0) * BP + b] + bias[u * 4 + 0];
            const float zf = zs * z[(u * 4 + 1) * BP + b] + bias[u * 4 + 1];
            const float zg = zs * z[(u * 4 + 2) * BP + b] + bias[u * 4 + 2];
            const float zo = zs * z[(u * 4 + 3) * BP + b] + bias[u * 4 + 3];
            const float gi = sigmoid_acc(zi), gf = sigmoid_acc(zf), gg = tanhf(zg), go = sigmoid_acc(zo);
            // explicit contraction: the plain and the utterance-queue instantiations must round identically
            const float c = __fmaf_rn(gf, fresh ? 0.f : c_state[i], __fmul_rn(gi, gg));
            float h = go * tanhf(c);
            if (sv.gates != nullptr) {          // training: keep what the backward needs, drop the carried h
                const int unit = sv.unit0 + u;
                float* gs = sv.gates + (size_t)b * 4 * RNN + unit;
                gs[0] = gi; gs[RNN] = gf; gs[2 * RNN] = gg; gs[3 * RNN] = go;
                sv.c[(size_t)b * RNN + unit] = c;
                if (sv.mask != nullptr) h *= sv.mask[(size_t)b * RNN + unit] * sv.scale;
                sv.h[(size_t)b * RNN + unit] = h;
            }
            c_state[i] = c;
            h_out[i] = h;
            if constexpr (TC) {
                __nv_bfloat16 hi, lo;
                umma::split_bf16(h, hi, lo);
                act16_buf[act16_off<BP>(b, k0 + u)] = hi;
                act16_buf[act16_off<BP>(BP + b, k0 + u)] = lo;
            } else {
                act_dst[u * BP + b] = h;
            }
        } else {
            h_out[i] = 0.f;
        }
    }
    consumer_sync();
}

// one out-of-line copy for both cells of the tensor-core kernel (see TACO_COMPACT)
template <int BP>
__device__ TACO_NI void lstm_pointwise_tc(const float* z, const float* bias, float* c_state, float* h_out,
                                          __nv_bfloat16* act16_buf, int k0, int B, int tid, const PwSave sv, const int* rows,
                                          int step, bool zero_z) {
    lstm_pointwise<BP, true>(z, bias, c_state, h_out, nullptr, act16_buf, k0, B, tid, sv, rows, step, zero_z);
}

// ---- producer warp -------------------------------------------------------------------------------
// Every pipeline stage has its own issuing lane (lane == stage): a thread's bulk copies complete one at a
// time (~0.4 us each whatever the size, profiles/r1_stream_probe.txt), copies of different threads overlap,
// and one lane per stage keeps each stage's uses in order.  The lanes run ONE convergent loop and poll their
// empty barriers with test_wait: a blocking try_wait would put the whole warp, i.e. all lanes, to sleep.
// ch = position in this CTA's chunk order, pc(ch) = physical chunk (k range): every CTA walks the K range
// from a different starting chunk so that the 128 CTAs do not request the same activation chunk at once.
template <int BP, typename WT, bool TC>
struct Producer {
    using SM = Smem<BP, WT, TC>;
    static constexpr int STAGES = SM::NSTAGES;

    __device__ static __forceinline__ int first_chunk(uint32_t g0, int lane) {
        return (lane + STAGES - (int)(g0 % STAGES)) % STAGES;
    }
    __device__ static __forceinline__ void issue_w(const Pipe& pipe, uint32_t stage, int pc, const WT* wsrc) {
        mbar_arrive_expect_tx(&pipe.full[stage], SM::STAGE_BYTES);     // bytes of BOTH copies of the stage
        bulk_g2s(pipe.stages + stage * SM::STAGE_BYTES, wsrc + (size_t)pc * SM::CK * SM::WROWS, SM::W_BYTES,
                 &pipe.full[stage]);
    }
    __device__ static __forceinline__ void issue_a(const Pipe& pipe, uint32_t stage, int pc, const uint8_t* src) {
        bulk_g2s(pipe.stages + stage * SM::STAGE_BYTES + SM::W_BYTES, src + (size_t)pc * SM::A_BYTES, SM::A_BYTES,
                 &pipe.full[stage]);
    }
    // All 32 lanes of the producer warp call this.  Chunks [0, pre) already have their weights in flight.
    // This CTA's nch chunks are the physical chunks abase .. abase+nch-1 of the activation vector (abase != 0
    // on the split-K tensor-core path); those with physical index < split come from a_lo, the others from a_hi.
    __device__ static __forceinline__ void phase(const Pipe& pipe, uint32_t g0, int nch, int pre, int rot, int lane,
                                                 const WT* wsrc, const uint8_t* a_lo, const uint8_t* a_hi, int split,
                                                 int abase) {
        int ch = lane < STAGES ? first_chunk(g0, lane) : nch;
        const uint64_t t0 = global_timer_ns();
        while (__any_sync(0xffffffffu, ch < nch)) {
            if (ch < nch) {
                const uint32_t g = g0 + ch, stage = g % STAGES, parity = ((g / STAGES) & 1u) ^ 1u;
                if (ch < pre || mbar_test_wait(&pipe.empty[stage], parity)) {
                    int pc = ch + rot;
                    pc = pc >= nch ? pc - nch : pc;
                    if (ch >= pre) issue_w(pipe, stage, pc, wsrc);
                    issue_a(pipe, stage, abase + pc, abase + pc < split ? a_lo : a_hi);
                    ch += STAGES;
                } else if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) {
                    spin_fail();
                }
            }
        }
    }
    // weights of the first `pre` chunks of the NEXT phase (its activations are not ready yet)
    __device__ static __forceinline__ void prefetch(const Pipe& pipe, uint32_t g0, int nch, int pre, int rot, int lane,
                                                    const WT* wsrc) {
        int ch = lane < STAGES ? first_chunk(g0, lane) : pre;
        const uint64_t t0 = global_timer_ns();
        while (__any_sync(0xffffffffu, ch < pre)) {
            if (ch < pre) {
                const uint32_t g = g0 + ch, stage = g % STAGES, parity = ((g / STAGES) & 1u) ^ 1u;
                if (mbar_test_wait(&pipe.empty[stage], parity)) {
                    int pc = ch + rot;
                    pc = pc >= nch ? pc - nch : pc;
                    issue_w(pipe, stage, pc, wsrc);
                    ch += STAGES;
                } else if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) {
                    spin_fail();
                }
            }
        }
    }
    // a phase that will never run was prefetched: complete its stages with dummy activation copies
    __device__ static __forceinline__ void drain(const Pipe& pipe, uint32_t g0, int pre, int lane, const uint8_t* dummy) {
        if (lane < STAGES) {
            const int ch = first_chunk(g0, lane);
            if (ch < pre) {
                const uint32_t g = g0 + ch, stage = g % STAGES;
                issue_a(pipe, stage, 0, dummy);
                mbar_wait(&pipe.full[stage], (g / STAGES) & 1u);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// P2: location-sensitive attention (src/model.py:117-148), consumer warps of every CTA.  One work item per
// (utterance, slice of encoder positions); partial softmax/context per slice, the last slice to arrive
// (atomic ticket) merges them - no extra grid barrier.
template <int BP, bool TC, bool CL = false>
__device__ __forceinline__ void attention_phase(const KParams& p, float* red, const float* wld_s, const float* wloc_s,
                                                const float* v_s, int* misc_s, int cur, int out_idx, int cta, int ncta,
                                                int tid, int warp, int lane, float* act_cur, __nv_bfloat16* act16_cur,
                                                bool timing, unsigned long long* t_fine, unsigned long long& t_flast,
                                                volatile int* sig = nullptr, int sig_val = 0, float* apx = nullptr,
                                                uint64_t* xbar2 = nullptr, bool rev = false, const int* rows = nullptr,
                                                int step = 0) {
    // rev: work items are dealt from the LAST CTA downwards, so that the CTAs beyond the 128 GEMM CTAs take them first
    // (small batches of the SIMT kernel: the GEMM CTAs then spend this phase on their background contractions)
    static_assert(!CL || TC, "cluster merge belongs to the tensor-core kernel");
    // sig: set to sig_val once this CTA's latency-critical loads (pq partials, windows, first pm tile) have landed;
    // the tensor-core kernel's producer warp starts its background weight streaming only then
    const int B = p.B, T = p.T, prv = cur ^ 1;
    (void)timing; (void)t_fine; (void)t_flast; (void)act_cur; (void)act16_cur;
    float* sc = red;
    const int NS = p.NS, SL = p.SL;
    const float* prev_in = p.attn_prev + (size_t)prv * B * p.Tp;
    const float* cum_in = p.attn_cum + (size_t)prv * B * p.Tp;
    float* prev_out = p.attn_prev + (size_t)cur * B * p.Tp;
    float* cum_out = p.attn_cum + (size_t)cur * B * p.Tp;
    FT_RESET();
    for (int item = rev ? ncta - 1 - cta : cta; item < B * NS; item += ncta) {
        const int b = item / NS, sl = item % NS;
        // continuous batching: row b carries utterance mb since decode step rows[ROW_START + b]; its first step sees
        // zero previous / cumulative attention weights (src/model.py:104-115); a retired row repeats its last
        // utterance without writing anything the host reads
        int mb = b, oi = out_idx;
        bool fresh = false;
        if (rows != nullptr) {
            const int u = rows[b], st = rows[ROW_START + b];
            mb = u >= 0 ? u : -1 - u;
            fresh = step == st;
            oi = u >= 0 ? step - st - 1 : -1;
        }
        // pad mask (src/model.py:131-132, :443-450); ragged continuous batching: the utterances of a launch are padded to a common
        // T and every utterance attends to its own length only, which is exactly its exact-length solo decode (weights of the
        // padded positions are 0, like the zero padding of the location convolution at the end of an exact-length utterance)
        const int len = p.text_lengths ? (int)p.text_lengths[mb] : T;
        // positions of this slice that are worked on.  Ragged launch: an utterance shorter than the launch's padded T only pays
        // for its own positions plus the reach of the location convolution (15; those positions are masked, get weight 0 and
        // keep the windows behind the last real position zero); whatever lies further back can only influence masked
        // positions.  The output buffers of a continuous-batching launch are zero initialised by the caller.
        // The NS slices of a ragged utterance split ITS OWN span evenly (not the launch's padded T): the phase ends with the slowest
        // CTA, and with the queue sorted by length the live utterances are of similar length at any time.
        int SLu = SL;
        if (rows != nullptr && p.text_lengths != nullptr) SLu = (min(T, len + 16) + NS - 1) / NS;
        const int t0 = sl * SLu, t1 = min(T, t0 + SLu);
        int n = max(0, t1 - t0);
        if (rows != nullptr && p.text_lengths != nullptr) n = max(0, min(n, len + 16 - t0));
        const size_t align_row = (size_t)(rows != nullptr ? rows[ROW_ABASE + b] : b) + (size_t)(oi < 0 ? 0 : oi) * B;
        consumer_sync();   // previous item's scratch is free
        const int ntiles = (n + TILE_POS - 1) / TILE_POS;
        // processed-memory tile 0: asynchronous copy, overlapped with the loads below
        auto issue_pm_tile = [&](int tb) {
            const int rows = min(TILE_POS, n - tb);
            const float* src = p.pm + ((size_t)mb * T + t0 + tb) * ATT_DIM;
            for (int i = tid; i < rows * (ATT_DIM / 4); i += NTC) cp_async16(sc + P2Scratch::PM + i * 4, src + i * 4);
        };
        issue_pm_tile(0);
        // previous / cumulative attention windows with a 15-position halo, zero padded
        const int wlen = ntiles * TILE_POS + LOC_K - 1 + 4;
        for (int i = tid; i < wlen; i += NTC) {
            const int t = t0 - (LOC_K - 1) / 2 + i;
            const bool ok = t >= 0 && t < T && !fresh;
            sc[P2Scratch::WIN + i] = ok ? __ldcg(prev_in + (size_t)b * p.Tp + t) : 0.f;
            sc[P2Scratch::WIN + P2Scratch::WIN_LEN + i] = ok ? __ldcg(cum_in + (size_t)b * p.Tp + t) : 0.f;
        }
        // processed query = sum of the 128 per-CTA partials
        {
            const int a = tid & 127, hf = tid >> 7;
            const float* src = p.pqpart + ((size_t)b * GEMM_CTAS + hf * 64) * ATT_DIM + a;
            sc[P2Scratch::PQ + tid] = strided_sum_cg<64, 64>(src, ATT_DIM);
        }
        consumer_sync();
        if (TACO_SIG_POINT == 0 && sig != nullptr && tid == 0) *sig = sig_val;
        FT(1);
        // energies: every thread owns two attention dims (a0, a0 + 64) and, per pass, four positions
        const int a0 = tid & 63;
        float wld[2][LOC_F];
#pragma unroll
        for (int f = 0; f < LOC_F; ++f) {
            wld[0][f] = wld_s[f * ATT_DIM + a0];
            wld[1][f] = wld_s[f * ATT_DIM + a0 + 64];
        }
        const float pq0 = sc[P2Scratch::PQ + a0] + sc[P2Scratch::PQ + 128 + a0];
        const float pq1 = sc[P2Scratch::PQ + a0 + 64] + sc[P2Scratch::PQ + 128 + a0 + 64];
        const float vw0 = v_s[a0], vw1 = v_s[a0 + 64];
        for (int tile = 0; tile < ntiles; ++tile) {
            const int tb = tile * TILE_POS;          // offset inside the slice
            // ---- location conv (src/model.py:125-126): 32 filters x (2 x 31 taps)
            {
                const int f = tid & 31, tg = tid >> 5;
                float o5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float x[LOC_K + 4];
                    const float* wsrc = sc + P2Scratch::WIN + c * P2Scratch::WIN_LEN + tb + tg * 5;
#pragma unroll
                    for (int i = 0; i < LOC_K + 4; ++i) x[i] = wsrc[i];
#pragma unroll
                    for (int k = 0; k < LOC_K; ++k) {
                        const float w = wloc_s[f * 63 + c * LOC_K + k];
#pragma unroll
                        for (int j = 0; j < 5; ++j) o5[j] = fmaf(w, x[k + j], o5[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 5; ++j) sc[P2Scratch::LOC + (tg * 5 + j) * P2Scratch::LOC_LD + f] = o5[j];
            }
            if (tile > 0) issue_pm_tile(tb);
            cp_async_wait_all();
            consumer_sync();
        FT(2);
            // ---- energies (src/model.py:127-130): v . tanh(pq + pm + Wld loc).  The 128-bit broadcast reads of the
            // location features bound this loop (LSU write-back), so each read feeds 8 FMAs: 2 dims x 4 positions
            // per thread.  Sum over the 128 dims: two dims in the thread, shuffle within the warp, 2 warp partials.
            const int npl = min(TILE_POS, n - tb);
            {
                const int hf = tid >> 6, w2 = (tid >> 5) & 1;
                for (int pl0 = hf * 4; pl0 < npl; pl0 += 16) {
                    float x0[4], x1[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        x0[j] = pq0 + sc[P2Scratch::PM + (pl0 + j) * ATT_DIM + a0];
                        x1[j] = pq1 + sc[P2Scratch::PM + (pl0 + j) * ATT_DIM + a0 + 64];
                    }
#pragma unroll
                    for (int f4 = 0; f4 < LOC_F; f4 += 4) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 l = *reinterpret_cast<const float4*>(sc + P2Scratch::LOC + (pl0 + j) * P2Scratch::LOC_LD + f4);
                            x0[j] = fmaf(wld[0][f4], l.x, x0[j]);
                            x1[j] = fmaf(wld[1][f4], l.x, x1[j]);
                            x0[j] = fmaf(wld[0][f4 + 1], l.y, x0[j]);
                            x1[j] = fmaf(wld[1][f4 + 1], l.y, x1[j]);
                            x0[j] = fmaf(wld[0][f4 + 2], l.z, x0[j]);
                            x1[j] = fmaf(wld[1][f4 + 2], l.z, x1[j]);
                            x0[j] = fmaf(wld[0][f4 + 3], l.w, x0[j]);
                            x1[j] = fmaf(wld[1][f4 + 3], l.w, x1[j]);
                        }
                    }
                    float x[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = fmaf(vw0, tanh_fast(x0[j]), vw1 * tanh_fast(x1[j]));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) x[j] += __shfl_xor_sync(0xffffffffu, x[j], o);
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) sc[P2Scratch::EPART + (pl0 + j) * 4 + w2] = x[j];
                    }
                }
            }
            consumer_sync();
        FT(3);
            if (tid < npl) {
                const float* ep = sc + P2Scratch::EPART + tid * 4;
                float e = ((ep[0] + ep[1]) + v_s[128]) * v_s[129];
                if (t0 + tb + tid >= len) e = -1e9f;                  // masked_fill, src/model.py:132
                sc[P2Scratch::E + tb + tid] = e;
            }
            consumer_sync();
        FT(4);
        }
        // ---- first 20 memory rows of the context sum: issued now, in flight during the softmax statistics ----
        const float2* mrow = reinterpret_cast<const float2*>(p.memory + ((size_t)mb * T + t0) * EMB) + tid;
        float2 mv0[20];
#pragma unroll
        for (int j = 0; j < 20; ++j) mv0[j] = j < n ? __ldg(mrow + (size_t)j * (EMB / 2)) : make_float2(0.f, 0.f);
        // ---- slice-local softmax statistics (warp 0) ----
        if (warp == 0) {
            float m = -INFINITY;
            for (int i = lane; i < n; i += 32) m = fmaxf(m, sc[P2Scratch::E + i]);
            m = warp_max(m);
            float ssum = 0.f;
            for (int i = lane; i < n; i += 32) {
                const float ew = expf(sc[P2Scratch::E + i] - m);
                sc[P2Scratch::E + i] = ew;
                ssum += ew;
            }
            ssum = warp_sum(ssum);
            if (lane == 0) {
                sc[P2Scratch::SC + 32] = n > 0 ? m : -INFINITY;
                sc[P2Scratch::SC + 33] = ssum;
            }
        }
        consumer_sync();
        if (TACO_SIG_POINT == 1 && sig != nullptr && tid == 0) *sig = sig_val;
        FT(5);
        // ---- slice-local un-normalised context: sum_t ew[t] * memory[b][t][:] ----
        float2 cacc = make_float2(0.f, 0.f);
        {
#pragma unroll
            for (int j = 0; j < 20; ++j) {
                const float ew = j < n ? sc[P2Scratch::E + j] : 0.f;
                cacc.x = fmaf(ew, mv0[j].x, cacc.x);
                cacc.y = fmaf(ew, mv0[j].y, cacc.y);
            }
            // remaining rows, 20 (40 KB per CTA) in flight at a time; the last batch is predicated, not serialised
            for (int i = 20; i < n; i += 20) {
                float2 mv[20];
#pragma unroll
                for (int j = 0; j < 20; ++j)
                    mv[j] = i + j < n ? __ldg(mrow + (size_t)(i + j) * (EMB / 2)) : make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 20; ++j) {
                    const float ew = i + j < n ? sc[P2Scratch::E + i + j] : 0.f;
                    cacc.x = fmaf(ew, mv[j].x, cacc.x);
                    cacc.y = fmaf(ew, mv[j].y, cacc.y);
                }
            }
        }
        if (TACO_SIG_POINT == 2 && sig != nullptr && tid == 0) *sig = sig_val;
        if constexpr (CL) {
            // The NS (2 or 4) slices of an utterance sit in one thread-block cluster (CTA = item).  Every slice
            // sends its partial (max, sum, un-normalised context) to all its siblings through DSMEM and arrives on
            // their mbarriers; then every CTA finishes ITS OWN share: the attention weights of its positions and
            // 512/NS context rows.  No ticket atomic, no partials through L2.
            if (NS > 1) {
                const uint32_t rank = (uint32_t)cta % GS, base = rank / (uint32_t)NS * (uint32_t)NS;
                const uint32_t slot = smem_u32(apx) + (uint32_t)sl * APART * 4u;
                for (int q = 0; q < NS; ++q) {
                    const uint32_t dst = mapa_u32(slot, base + q);
                    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(dst + (4u + 2u * tid) * 4u), "f"(cacc.x),
                                 "f"(cacc.y)
                                 : "memory");
                    if (tid == 0) {
                        st_cluster_f32(dst, sc[P2Scratch::SC + 32]);
                        st_cluster_f32(dst + 4u, sc[P2Scratch::SC + 33]);
                    }
                }
                consumer_sync();
                if (tid == 0) {
                    const uint32_t bar_local = smem_u32(xbar2);
                    for (int q = 0; q < NS; ++q)
                        if (q != sl) mbar_arrive_remote_release(mapa_u32(bar_local, base + q));
                }
                if (!mbar_try_wait_acquire_cluster(xbar2, (uint32_t)cur)) {
                    const uint64_t tw0 = global_timer_ns();
                    while (!mbar_try_wait_acquire_cluster(xbar2, (uint32_t)cur)) {
                        if (global_timer_ns() - tw0 > SPIN_TIMEOUT_NS) spin_fail();
                    }
                }
        FT(6);
                float M = -INFINITY;
                for (int q = 0; q < NS; ++q) M = fmaxf(M, apx[q * APART]);
                float scale[GS], S = 0.f;
#pragma unroll
                for (int q = 0; q < GS; ++q) {
                    scale[q] = q < NS ? expf(apx[q * APART] - M) : 0.f;
                    S += q < NS ? scale[q] * apx[q * APART + 1] : 0.f;
                }
                const float mine = expf(sc[P2Scratch::SC + 32] - M);
                // attention weights of this slice: alignment output, prev, cum (src/model.py:142-145)
                for (int i = tid; i < n; i += NTC) {
                    const int t = t0 + i;
                    const float w = sc[P2Scratch::E + i] * mine / S;
                    prev_out[(size_t)b * p.Tp + t] = w;
                    // cumulative weights of this slice are still in the window (zero for a fresh row): no reload from L2
                    cum_out[(size_t)b * p.Tp + t] = sc[P2Scratch::WIN + P2Scratch::WIN_LEN + (LOC_K - 1) / 2 + i] + w;
                    if (oi >= 0) p.align[align_row * T + t] = w;
                }
                // context rows [sl * 512/NS, (sl+1) * 512/NS)
                const int per = EMB / NS;
                for (int d2 = tid; d2 < per / 2; d2 += NTC) {
                    const int d = sl * per + 2 * d2;
                    float2 c = make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < GS; ++q) {
                        if (q < NS) {
                            const float2 v = *reinterpret_cast<const float2*>(apx + q * APART + 4 + d);
                            c.x = fmaf(scale[q], v.x, c.x);
                            c.y = fmaf(scale[q], v.y, c.y);
                        }
                    }
                    c.x = c.x / S;
                    c.y = c.y / S;
                    __nv_bfloat16 hx, lx, hy, ly;
                    umma::split_bf16(c.x, hx, lx);
                    umma::split_bf16(c.y, hy, ly);
                    *reinterpret_cast<__nv_bfloat162*>(act16_cur + act16_off<BP>(b, ACT_CTX + d)) = __nv_bfloat162(hx, hy);
                    *reinterpret_cast<__nv_bfloat162*>(act16_cur + act16_off<BP>(BP + b, ACT_CTX + d)) = __nv_bfloat162(lx, ly);
                    *reinterpret_cast<float2*>(p.ctx_bj + (size_t)b * EMB + d) = c;
                    if (p.sv_ctx != nullptr && out_idx >= 0)
                        *reinterpret_cast<float2*>(p.sv_ctx + ((size_t)out_idx * B + b) * EMB + d) = c;
                }
        FT(8);
                continue;
            }
        }
        float* part = p.apart + ((size_t)b * MAX_NS + sl) * APART;
        if (NS > 1) {
            // Slices of an utterance on different CTAs without a cluster: every slice publishes its partial (max, sum,
            // un-normalised context) in L2, counts itself in, waits until all NS slices of the utterance are in, and then
            // finishes ITS OWN share - the attention weights of its positions and 512/NS context dims - so nothing is
            // serialised on a last arriver and the weights never leave shared memory.  One work item per CTA
            // (B * NS <= grid size by construction), so the wait cannot deadlock.
            reinterpret_cast<float2*>(part + 4)[tid] = cacc;
            if (tid == 0) { part[0] = sc[P2Scratch::SC + 32]; part[1] = sc[P2Scratch::SC + 33]; }
            consumer_sync();
            if (tid == 0) {
                // release: this CTA's partial (ordered by the bar.sync above); acquire: the others'
                const uint32_t target = (uint32_t)NS * (uint32_t)((p.mode == 1 ? out_idx : out_idx + 1) + 1);
                uint32_t seen = atom_add_acq_rel_gpu(p.ctrl + CTRL_TICKETS + b, 1u) + 1u;
                if (seen < target) {
                    const uint64_t tw0 = global_timer_ns();
                    while (ld_acquire_u32(p.ctrl + CTRL_TICKETS + b) < target) {
                        if (global_timer_ns() - tw0 > SPIN_TIMEOUT_NS) spin_fail();
                    }
                }
            }
            consumer_sync();
        FT(6);
            // all partials: context partial of this thread's dims first (independent of the scales), then (max, sum)
            const int per = (EMB + NS - 1) / NS;                  // context dims finished by this slice
            const int d = sl * per + tid;
            const bool mine_d = tid < per && d < EMB;
            float cv[MAX_NS];
#pragma unroll
            for (int q = 0; q < MAX_NS; ++q)
                cv[q] = q < NS && mine_d ? __ldcg(p.apart + ((size_t)b * MAX_NS + q) * APART + 4 + d) : 0.f;
            if (tid < 32) {
                float m = -INFINITY;
                if (tid < NS) m = __ldcg(p.apart + ((size_t)b * MAX_NS + tid) * APART);
                const float M = warp_max(m);
                float scale = 0.f, ssum = 0.f;
                if (tid < NS) {
                    scale = expf(m - M);
                    ssum = scale * __ldcg(p.apart + ((size_t)b * MAX_NS + tid) * APART + 1);
                }
                ssum = warp_sum(ssum);
                if (tid < NS) sc[P2Scratch::SC + tid] = scale;
                if (tid == 0) sc[P2Scratch::SC + 33] = ssum;
            }
            consumer_sync();
        FT(7);
            const float S = sc[P2Scratch::SC + 33];
            if (mine_d) {
                float c = 0.f;
#pragma unroll
                for (int q = 0; q < MAX_NS; ++q) c = fmaf(q < NS ? sc[P2Scratch::SC + q] : 0.f, cv[q], c);
                c = c / S;
                if constexpr (TC) {
                    __nv_bfloat16 hi, lo;
                    umma::split_bf16(c, hi, lo);
                    act16_cur[act16_off<BP>(b, ACT_CTX + d)] = hi;
                    act16_cur[act16_off<BP>(BP + b, ACT_CTX + d)] = lo;
                } else {
                    act_cur[(size_t)(ACT_CTX + d) * BP + b] = c;
                }
                p.ctx_bj[(size_t)b * EMB + d] = c;
                if (p.sv_ctx != nullptr && out_idx >= 0) p.sv_ctx[((size_t)out_idx * B + b) * EMB + d] = c;
            }
            // attention weights of this slice: alignment output, prev, cum (src/model.py:142-145)
            const float mine = sc[P2Scratch::SC + sl];
            for (int i = tid; i < n; i += NTC) {
                const int t = t0 + i;
                const float w = sc[P2Scratch::E + i] * mine / S;
                prev_out[(size_t)b * p.Tp + t] = w;
                cum_out[(size_t)b * p.Tp + t] = sc[P2Scratch::WIN + P2Scratch::WIN_LEN + (LOC_K - 1) / 2 + i] + w;
                if (oi >= 0) p.align[align_row * T + t] = w;
            }
        FT(8);
            continue;
        }
        // a single slice: the whole utterance is here
        if (tid == 0) sc[P2Scratch::SC + 0] = 1.f;
        consumer_sync();
        const float S = sc[P2Scratch::SC + 33];
        cacc.x = cacc.x / S;
        cacc.y = cacc.y / S;
        // context -> act (column b) and the utterance-major copy used by the projections
        if constexpr (TC) {
            __nv_bfloat16 hx, lx, hy, ly;
            umma::split_bf16(cacc.x, hx, lx);
            umma::split_bf16(cacc.y, hy, ly);
            *reinterpret_cast<__nv_bfloat162*>(act16_cur + act16_off<BP>(b, ACT_CTX + 2 * tid)) = __nv_bfloat162(hx, hy);
            *reinterpret_cast<__nv_bfloat162*>(act16_cur + act16_off<BP>(BP + b, ACT_CTX + 2 * tid)) = __nv_bfloat162(lx, ly);
        } else {
            act_cur[(size_t)(ACT_CTX + 2 * tid) * BP + b] = cacc.x;
            act_cur[(size_t)(ACT_CTX + 2 * tid + 1) * BP + b] = cacc.y;
        }
        reinterpret_cast<float2*>(p.ctx_bj + (size_t)b * EMB)[tid] = cacc;
        if (p.sv_ctx != nullptr && out_idx >= 0) reinterpret_cast<float2*>(p.sv_ctx + ((size_t)out_idx * B + b) * EMB)[tid] = cacc;
        // attention weights: alignment output, prev, cum (src/model.py:142-145)
        for (int t = tid; t < T; t += NTC) {
            const float w = sc[P2Scratch::E + t] / S;
            prev_out[(size_t)b * p.Tp + t] = w;
            cum_out[(size_t)b * p.Tp + t] = (fresh ? 0.f : __ldcg(cum_in + (size_t)b * p.Tp + t)) + w;
            if (oi >= 0) p.align[align_row * T + t] = w;
        }
        FT(8);
    }
    if (sig != nullptr && tid == 0) *sig = sig_val;       // CTAs without a work item
}

// P4: per utterance: sum projection partials -> mel frame, gate logit, stop rule (src/model.py:327 /
// train.py:71-79), prenet of the next input frame (src/model.py:166-171).  Consumer warps of every CTA.
template <int BP, bool TC>
__device__ __forceinline__ void projection_phase(const KParams& p, float* red, bool tf, int out_idx, int cta, int ncta,
                                                 int tid, float* act_cur, __nv_bfloat16* act16_cur,
                                                 volatile int* sig = nullptr, int sig_val = 0, bool rev = false,
                                                 const int* rows = nullptr, int step = 0,
                                                 __nv_bfloat16* act16_nxt = nullptr) {
    const int B = p.B;
    (void)act_cur; (void)act16_cur;
    float* sc = red;
    const int NQ = p.NQ, RQ = PRENET / NQ;
    for (int item = rev ? ncta - 1 - cta : cta; item < B * NQ; item += ncta) {
        const int b = item / NQ, q = item % NQ;
        // continuous batching: outputs go to the utterance the row carries, at the utterance's own frame index
        int ub = b, oi = out_idx;
        bool live = true;
        if (rows != nullptr) {
            const int u = rows[b], st = rows[ROW_START + b];
            live = u >= 0;
            ub = live ? u : -1 - u;
            oi = live ? step - st - 1 : -1;
        }
        consumer_sync();
        // The prenet weights stream from L2 every step; their loads do not depend on this step's values, so each
        // batch is issued one stage ahead of its use (first half of layer 1 before the partial sums, second half
        // of layer 1 and the first batch of layer 2 before layer 1 is finished).
        const float* w1src = p.prenet1T + tid;
        const int nl = tid % RQ, kg = tid / RQ, kn = PRENET / NQ;
        const float* w2src = p.prenet2T + (size_t)(kg * kn) * PRENET + q * RQ + nl;
        float w1a[40];
        if (!tf) {
#pragma unroll
            for (int i = 0; i < 40; ++i) w1a[i] = __ldg(w1src + (size_t)i * PRENET);
        }
        {
            const int o = tid & 127, ig = tid >> 7;
            if (o < PROJ_OUT) {
                const float* src = p.melpart + ((size_t)b * GEMM_CTAS + ig * 64) * PROJ_PAD + o;
                sc[ig * 128 + o] = strided_sum_cg<64, 64>(src, PROJ_PAD);
            }
        }
        consumer_sync();
        if (sig != nullptr && tid == 0) *sig = sig_val;
        if (tid < PROJ_OUT) {
            const float v = sc[tid] + sc[128 + tid] + __ldg(p.proj_bias + tid);
            sc[256 + tid] = v;                                  // mel frame (0..79), gate logit (80)
            if (q == 0 && oi >= 0) {
                if (tid < N_MELS) p.mel[((size_t)ub * p.max_steps + oi) * N_MELS + tid] = v;
                else p.gate[(size_t)ub * p.max_steps + oi] = v;
            }
            if (rows != nullptr) {
                // utterance queue: a row whose utterance stopped (its own gate, train.py:71-79 with the steps > 1 guard of
                // src/model.py:327, or the frame cap) takes the next queued utterance from the following decode step on
                if (q == 0 && tid == N_MELS) {
                    int refill = 0;
                    if (live && ((oi >= 1 && sigmoid_acc(v) > p.thr) || oi + 1 >= p.max_steps)) {
                        refill = 2;
                        const uint32_t nxt = atomicAdd(p.ctrl + CTRL_QHEAD, 1u) + (uint32_t)B;
                        if (nxt < (uint32_t)p.n_utts) {
                            p.ctrl[CTRL_UTT + b] = nxt;
                            p.ctrl[CTRL_START + b] = (uint32_t)(step + 1);
                            refill = 3;
                        } else {
                            p.ctrl[CTRL_UTT + b] = (uint32_t)(-1 - ub);
                        }
                        if (atomicAdd(p.ctrl + CTRL_FINISHED, 1u) + 1u == (uint32_t)p.n_utts) atomicExch(p.ctrl + CTRL_STOP, 1u);
                    }
                    reinterpret_cast<int*>(sc)[340] = refill;
                }
            } else if (q == 0 && !tf && tid == N_MELS && out_idx >= 1) {
                // stop rule: src/model.py:327 (batch-wide) or train.py:71-79 (per utterance)
                const bool fired = sigmoid_acc(v) > p.thr;
                if (fired) {
                    if (p.stop_mode == B200TACO_STOP_BATCH_WIDE) {
                        atomicExch(p.ctrl + CTRL_STOP, 1u);
                    } else if (p.ctrl[CTRL_DONE + b] == 0u) {
                        p.ctrl[CTRL_DONE + b] = 1u;
                        p.ctrl[CTRL_LEN + b] = (uint32_t)(out_idx + 1);
                        if (atomicAdd(p.ctrl + CTRL_FINISHED, 1u) + 1u == (uint32_t)B)
                            atomicExch(p.ctrl + CTRL_STOP, 1u);
                    }
                }
            }
        }
        if (tf) continue;
        // second half of layer 1 and first batch of layer 2: in flight over the barrier and the first-half FMAs
        float w1b[40], w2a[32];
#pragma unroll
        for (int i = 0; i < 40; ++i) w1b[i] = __ldg(w1src + (size_t)(40 + i) * PRENET);
#pragma unroll
        for (int i = 0; i < 32; ++i) w2a[i] = __ldg(w2src + (size_t)i * PRENET);
        consumer_sync();
        if constexpr (TC) {
            // the utterance's length is published only now: the last frame's mel / gate stores of the other threads are
            // ordered before it by the barrier above, its alignment rows by the grid barrier after phase B.  d_lengths may
            // be mapped host memory that the host polls to post-process finished utterances while the launch goes on.
            if (rows != nullptr && q == 0 && tid == N_MELS && (reinterpret_cast<const int*>(sc)[340] & 2) != 0) {
                __threadfence_system();
                *reinterpret_cast<volatile int32_t*>(p.lengths + ub) = oi + 1;
            }
            // the new utterance starts from a zero decoder hidden state: its rows of the operand that the next step's
            // background contraction W[dh] . dh reads (the attention LSTM ignores its accumulator for a fresh row)
            if (rows != nullptr && q == 0 && (reinterpret_cast<const int*>(sc)[340] & 1) != 0) {
                const __nv_bfloat16 zero = __float2bfloat16(0.f);
                for (int k = tid; k < RNN; k += NTC) {
                    act16_nxt[act16_off<BP>(b, ACT_DH + k)] = zero;
                    act16_nxt[act16_off<BP>(BP + b, ACT_DH + k)] = zero;
                }
            }
        }
        // prenet layer 1 (all 256 outputs in every quarter CTA), eval mode: no dropout (F4)
        {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 40; ++i) acc = fmaf(w1a[i], sc[256 + i], acc);
#pragma unroll
            for (int i = 0; i < 40; ++i) acc = fmaf(w1b[i], sc[256 + 40 + i], acc);
            sc[384 + tid] = fmaxf(acc, 0.f);
        }
        consumer_sync();
        // prenet layer 2: this CTA's RQ outputs, k split over NQ thread groups
        {
            const float* x = sc + 384 + kg * kn;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc = fmaf(w2a[i], x[i], acc);
#pragma unroll 1
            for (int m0 = 32; m0 < kn; m0 += 32) {
                float w[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) w[i] = __ldg(w2src + (size_t)(m0 + i) * PRENET);
#pragma unroll
                for (int i = 0; i < 32; ++i) acc = fmaf(w[i], x[m0 + i], acc);
            }
            sc[640 + kg * RQ + nl] = acc;
        }
        consumer_sync();
        if (tid < RQ) {
            float v = 0.f;
            for (int kg = 0; kg < NQ; ++kg) v += sc[640 + kg * RQ + tid];
            v = fmaxf(v, 0.f);
            if constexpr (TC) {
                __nv_bfloat16 hi, lo;
                umma::split_bf16(v, hi, lo);
                act16_cur[act16_off<BP>(b, ACT_P + q * RQ + tid)] = hi;
                act16_cur[act16_off<BP>(BP + b, ACT_P + q * RQ + tid)] = lo;
            } else {
                act_cur[(size_t)(ACT_P + q * RQ + tid) * BP + b] = v;
            }
        }
    }
    if (sig != nullptr && tid == 0) *sig = sig_val;       // CTAs without a work item
}

// ---------------------------------------------------------------------------------------
template <int BP, typename WT, bool TC>
__global__ void __launch_bounds__(NT, 1) decoder_kernel(const KParams p) {
    using SM = Smem<BP, WT, TC>;
    constexpr int STAGES = SM::NSTAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const bool is_gemm = cta < GEMM_CTAS;
    const bool is_producer = warp == NW;
    const int B = p.B, T = p.T;

    Pipe pipe;
    pipe.full = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    pipe.empty = pipe.full + STAGES;
    pipe.stages = smem + SM::OFF_STAGE;
    float* red = reinterpret_cast<float*>(smem + SM::OFF_RED);
    float* z_s = reinterpret_cast<float*>(smem + SM::OFF_Z);
    float* ca_s = reinterpret_cast<float*>(smem + SM::OFF_CA);
    float* cd_s = reinterpret_cast<float*>(smem + SM::OFF_CD);
    float* h_s = reinterpret_cast<float*>(smem + SM::OFF_H);
    float* bias_s = reinterpret_cast<float*>(smem + SM::OFF_BIAS);
    float* wq_s = reinterpret_cast<float*>(smem + SM::OFF_WQ);
    float* wproj_s = reinterpret_cast<float*>(smem + SM::OFF_WPROJ);
    float* wpc_s = wproj_s + 656;
    float* cx_s = h_s + 8 * BP;
    float* wloc_s = reinterpret_cast<float*>(smem + SM::OFF_WLOC);
    float* wld_s = reinterpret_cast<float*>(smem + SM::OFF_WLD);
    float* v_s = reinterpret_cast<float*>(smem + SM::OFF_V);
    int* misc_s = reinterpret_cast<int*>(smem + SM::OFF_MISC);
    uint64_t* acc_bar = reinterpret_cast<uint64_t*>(smem + SM::OFF_ACCBAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::OFF_TMEM);

    // ---- one-time setup ----
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&pipe.full[s], 1);
            mbar_init(&pipe.empty[s], TC ? 1 : NW);
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    static_assert(!TC, "the tensor-core path is decoder_tc_kernel");
    for (int i = tid; i < UNITS_PER_CTA * BP; i += NT) { ca_s[i] = 0.f; cd_s[i] = 0.f; h_s[i] = 0.f; }
    if (is_gemm) {
        for (int i = tid; i < 32; i += NT) {
            bias_s[i] = p.bias_attn[cta * 32 + i];
            bias_s[32 + i] = p.bias_dec[cta * 32 + i];
        }
        for (int i = tid; i < 128 * 8; i += NT) wq_s[i] = p.wq_slices[(size_t)cta * 1024 + i];
        for (int i = tid; i < PROJ_OUT * 8; i += NT) wproj_s[i] = p.wproj_slices[(size_t)cta * PROJ_OUT * 8 + i];
        // context part of the projections for this CTA's four context rows: wpc[o][j] = W[o][1024 + 4*cta + j]
        for (int i = tid; i < PROJ_OUT * 4; i += NT) wpc_s[i] = p.wproj_ctxT[(size_t)(4 * cta + (i & 3)) * PROJ_PAD + (i >> 2)];
    }
    for (int i = tid; i < LOC_F * 2 * LOC_K; i += NT) wloc_s[(i / 62) * 63 + (i % 62)] = p.loc_conv[i];
    for (int i = tid; i < LOC_F * ATT_DIM; i += NT) wld_s[i] = p.loc_denseT[i];
    for (int i = tid; i < 130; i += NT) v_s[i] = p.attn_v[i];
    __syncthreads();

    // per-CTA weight bytes are the same on both paths: 32 rows x K, or 128 rows x K/4 (split-K group)
    const WT* w_attn = reinterpret_cast<const WT*>(TC ? p.w_attn_tc : p.w_attn) + (size_t)cta * K_ATTN * ROWS_PER_CTA;
    const WT* w_dec = reinterpret_cast<const WT*>(TC ? p.w_dec_tc : p.w_dec) + (size_t)cta * K_DEC * ROWS_PER_CTA;
    constexpr int KSPLIT = TC ? GS : 1;
    constexpr int NCH1 = K_ATTN / SM::CK / KSPLIT, NCH3 = K_DEC / SM::CK / KSPLIT, PCH = PRENET / SM::CK;
    static_assert((K_ATTN / SM::CK) % KSPLIT == 0 && (K_DEC / SM::CK) % KSPLIT == 0, "K split");
    const int ab1 = TC ? (cta % GS) * NCH1 : 0, ab3 = TC ? (cta % GS) * NCH3 : 0;    // first physical chunk of this CTA
    // bytes of one activation buffer in the GEMM operand layout and of one step of p_all
    constexpr size_t ABUF_BYTES = TC ? (size_t)(ACT_ROWS / TC_CK) * SM::A_BYTES : (size_t)ACT_ROWS * BP * 4;
    uint8_t* const abase = TC ? reinterpret_cast<uint8_t*>(p.act16) : reinterpret_cast<uint8_t*>(p.act);
#ifdef TACO_NO_PREFETCH
    constexpr int PRE1 = 0, PRE3 = 0;
#else
    constexpr int PRE1 = NCH1 < STAGES ? NCH1 : STAGES, PRE3 = NCH3 < STAGES ? NCH3 : STAGES;
#endif
    static_assert(PRENET % SM::CK == 0 && ACT_CTX % SM::CK == 0, "prenet / context rows must be whole chunks");

    uint32_t epoch = 0;
    uint32_t g = 0;             // chunks consumed (consumers) / chunks fully issued (producer)
    const bool tf = p.mode == 1;
    const int total_steps = tf ? p.max_steps : p.max_steps + 1;
    int steps_done = 0;         // decode steps executed
    bool stopped = false;

    // phase timers (CTA 0, thread 0): time from the end of one grid barrier to the end of the next
    const bool timing = cta == 0 && tid == 0;
    unsigned long long t_phase[4] = {0ull, 0ull, 0ull, 0ull}, t_last = timing ? global_timer_ns() : 0ull;
#ifdef TACO_PROFILE_FINE
    unsigned long long t_fine[16] = {}, t_flast = 0ull;
#endif
    auto stamp = [&](int ph) {
        if (timing) {
            const unsigned long long now = global_timer_ns();
            t_phase[ph] += now - t_last;
            t_last = now;
        }
    };

    // producer: weight chunks of the very first phase
    const int rot1 = ((cta / KSPLIT) * NCH1 * KSPLIT) / GEMM_CTAS, rot3 = ((cta / KSPLIT) * NCH3 * KSPLIT) / GEMM_CTAS;
    using PR = Producer<BP, WT, TC>;
    if (is_producer && is_gemm) PR::prefetch(pipe, g, NCH1, PRE1, rot1, lane, w_attn);

    for (int s = 0; s < total_steps; ++s) {
        const int cur = s & 1, prv = cur ^ 1;
        float* act_cur = p.act + (size_t)cur * ACT_ROWS * BP;
        float* act_prv = p.act + (size_t)prv * ACT_ROWS * BP;
        const uint8_t* a_cur = abase + (size_t)cur * ABUF_BYTES;        // GEMM operand view of the same buffers
        const uint8_t* a_prv = abase + (size_t)prv * ABUF_BYTES;
        __nv_bfloat16* act16_cur = p.act16 + (size_t)cur * (ABUF_BYTES / 2);
        __nv_bfloat16* act16_prv = p.act16 + (size_t)prv * (ABUF_BYTES / 2);
        const int out_idx = tf ? s : s - 1;          // F6: decode step 0 of inference is discarded

        // =============================== P1: attention LSTM ===============================
        if (is_gemm) {
            if (is_producer) {
                fence_proxy_async();
                // prenet rows: previous buffer, or the precomputed buffer when teacher forcing
                const uint8_t* a0 = tf ? static_cast<const uint8_t*>(p.p_all) + (size_t)s * PCH * SM::A_BYTES : a_prv;
                PR::phase(pipe, g, NCH1, PRE1, rot1, lane, w_attn, a0, a_prv, PCH, ab1);
                g += NCH1;
                PR::prefetch(pipe, g, NCH3, PRE3, rot3, lane, w_dec);          // across P2
            } else {
                FT_RESET();
                {
                    float acc[8][Tile<BP>::BPL];
                    gemm_consume<BP, WT>(pipe, g, NCH1, warp, lane, acc);
                    gemm_reduce<BP>(acc, red, z_s, warp, lane, tid);
                }
                FT(9);
                lstm_pointwise<BP, TC>(z_s, bias_s, ca_s, h_s, act_cur + (size_t)(ACT_AH + cta * UNITS_PER_CTA) * BP,
                                       act16_cur, ACT_AH + cta * UNITS_PER_CTA, B, tid,
                                       pw_save(p.sv_ga, p.sv_ca, p.sv_ah, p.mask_a, p.drop_scale, s, B, cta * UNITS_PER_CTA));
                FT(10);
                // this CTA's slice of the query projection (src/model.py:123): pqpart[b][cta][a]
                for (int i = tid; i < ATT_DIM * B; i += NTC) {
                    const int a = i % ATT_DIM, b = i / ATT_DIM;
                    float sum = 0.f;
#pragma unroll
                    for (int u = 0; u < UNITS_PER_CTA; ++u) sum = fmaf(wq_s[a * 8 + u], h_s[u * BP + b], sum);
                    p.pqpart[((size_t)b * GEMM_CTAS + cta) * ATT_DIM + a] = sum;
                }
                FT(11);
            }
        }
        grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
        stamp(0);

        // =============================== P2: attention ===============================
        if (!is_producer)
            attention_phase<BP, TC>(p, red, wld_s, wloc_s, v_s, misc_s, cur, out_idx, cta, ncta, tid, warp, lane, act_cur,
                                    act16_cur, timing, TACO_FINE_ARGS);
        grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
        stamp(1);

        // =============================== P3: decoder LSTM ===============================
        if (is_gemm) {
            if (is_producer) {
                fence_proxy_async();
                const uint8_t* a = a_cur + (size_t)(ACT_CTX / SM::CK) * SM::A_BYTES;     // starts at the context rows
                PR::phase(pipe, g, NCH3, PRE3, rot3, lane, w_dec, a, a, 0, ab3);
                g += NCH3;
                PR::prefetch(pipe, g, NCH1, PRE1, rot1, lane, w_attn);         // next step's P1, across P4
            } else {
                for (int i = tid; i < 4 * B; i += NTC) cx_s[(i & 3) * BP + (i >> 2)] = __ldcg(p.ctx_bj + (size_t)(i >> 2) * EMB + 4 * cta + (i & 3));
                FT_RESET();
                {
                    float acc[8][Tile<BP>::BPL];
                    gemm_consume<BP, WT>(pipe, g, NCH3, warp, lane, acc);
                    gemm_reduce<BP>(acc, red, z_s, warp, lane, tid);
                }
                FT(12);
                lstm_pointwise<BP, TC>(z_s, bias_s + 32, cd_s, h_s, act_prv + (size_t)(ACT_DH + cta * UNITS_PER_CTA) * BP,
                                       act16_prv, ACT_DH + cta * UNITS_PER_CTA, B, tid,
                                       pw_save(p.sv_gd, p.sv_cd, p.sv_dh, p.mask_d, p.drop_scale, s, B, cta * UNITS_PER_CTA));
                FT(13);
                // this CTA's slice of linear_projection / gate_layer (src/model.py:272-274): 8 decoder-h units + 4 context rows
                for (int i = tid; i < PROJ_PAD * B; i += NTC) {
                    const int o = i % PROJ_PAD, b = i / PROJ_PAD;
                    if (o < PROJ_OUT) {
                        float sum = 0.f;
#pragma unroll
                        for (int u = 0; u < UNITS_PER_CTA; ++u) sum = fmaf(wproj_s[o * 8 + u], h_s[u * BP + b], sum);
#pragma unroll
                        for (int j = 0; j < 4; ++j) sum = fmaf(wpc_s[o * 4 + j], cx_s[j * BP + b], sum);
                        p.melpart[((size_t)b * GEMM_CTAS + cta) * PROJ_PAD + o] = sum;
                    }
                }
            }
        }
        grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
        stamp(2);

        // =============================== P4: projection sums, stop rule, prenet ===============================
        if (!is_producer) projection_phase<BP, TC>(p, red, tf, out_idx, cta, ncta, tid, act_cur, act16_cur);
        steps_done = s + 1;
        if (!tf) {
            grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
            stamp(3);
            if (__ldcg(p.ctrl + CTRL_STOP) != 0u) { stopped = true; break; }
        }
    }

    // ---- drain: the producer prefetched weight chunks of a phase that will not run ----
    if (is_producer && is_gemm) PR::drain(pipe, g, PRE1, lane, abase);
    if (cta == 0 && tid == 0) {
        unsigned long long* prof = reinterpret_cast<unsigned long long*>(p.ctrl + CTRL_PROF);
        for (int i = 0; i < 4; ++i) prof[i] = t_phase[i];
        prof[4] = (unsigned long long)steps_done;
#ifdef TACO_PROFILE_FINE
        for (int i = 0; i < 16; ++i) prof[5 + i] = t_fine[i];
        prof[5 + 14] = g_dbg[0]; prof[5 + 15] = g_dbg[1];
        g_dbg[0] = 0; g_dbg[1] = 0;
#endif
        const int S = tf ? p.max_steps : steps_done - 1;     // output frames (decode step 0 discarded)
        *p.steps_run = S;
        for (int b = 0; b < B; ++b) {
            int len = S;
            if (!tf && p.stop_mode == B200TACO_STOP_PER_UTTERANCE && __ldcg(p.ctrl + CTRL_DONE + b))
                len = (int)__ldcg(p.ctrl + CTRL_LEN + b);
            p.lengths[b] = len;
        }
    }
    (void)stopped;
}

template <int BP, typename WT, bool TC>
int launch(const KParams& kp, int ncta, cudaStream_t stream) {
    using SM = Smem<BP, WT, TC>;
    auto kern = decoder_kernel<BP, WT, TC>;
    TACO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    int per_sm = 0;
    TACO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, SM::TOTAL));
    TACO_REQUIRE(per_sm >= 1, B200TACO_ERR_UNSUPPORTED, "decoder kernel does not fit on an SM (smem %d B)", SM::TOTAL);
    void* args[] = {const_cast<KParams*>(&kp)};
    TACO_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(ncta), dim3(NT), args, SM::TOTAL, stream));
    launch_counter()++;
    return B200TACO_OK;
}


// =======================================================================================================
// Tensor-core decoder kernel (batch tile >= 32, bf16 weights): same four phases, but the gate contractions
// are HOISTED.  Both LSTM pre-activations are sums over K ranges that become known at different barriers:
//     attention LSTM (step t+1):  ah_t (after barrier 1 of step t) | ctx_t (after barrier 2) | p_{t+1} (after barrier 4)
//     decoder LSTM   (step t)  :  dh_{t-1} (after barrier 3 of step t-1) | ah_t (after barrier 1) | ctx_t (after barrier 2)
// The tcgen05 contraction needs no consumer warp (TMA producer warp + one MMA-issuing thread + TMEM), so every
// K range is multiplied as soon as it exists, in the BACKGROUND of the attention / projection phases, into two
// TMEM accumulators.  Only the last range of each contraction is on the critical path:
//     phase A: 1 chunk (prenet rows)  -> attention LSTM epilogue        (was 7 chunks)
//     phase C: 2 chunks (context rows) -> decoder LSTM epilogue          (was 10 chunks)
// Split-K over groups of GS=4 CTAs as before, but every availability class is dealt evenly to the four members
// (dims.cuh tc_*_act_chunk), so the critical chunks load in parallel.
constexpr int NT_TC = NTC + 64;                 // + producer warp (8) + MMA warp (9)
constexpr int TC_PRODUCER_WARP = NW, TC_MMA_WARP = NW + 1;

template <int BP>
struct TcPipe {
    using SM = Smem<BP, __nv_bfloat16, true>;
    static constexpr int STAGES = SM::NSTAGES;

    // All 32 lanes of the producer warp call this (one issuing lane per stage, see Producer).  Issues the n chunks
    // of one job at ring positions g0.. ; mode 0: weights + activations, 1: weights only (activations not ready
    // yet), 2: activations only (weights went out earlier with mode 1).
    // (inlined: every call site has constant n / mode, and the producer warp is on the latency-critical path of the
    // critical chunks - measured 0.56 us per step faster than one out-of-line copy)
    __device__ static __forceinline__ void produce(const Pipe& pipe, uint32_t g0, int n, int mode, int lane,
                                           const __nv_bfloat16* wsrc, const uint8_t* abuf, int achunk0) {
        int c = lane < STAGES ? (lane + STAGES - (int)(g0 % STAGES)) % STAGES : n;
        const uint64_t t0 = global_timer_ns();
        while (__any_sync(0xffffffffu, c < n)) {
            if (c < n) {
                const uint32_t g = g0 + c, stage = g % STAGES, parity = ((g / STAGES) & 1u) ^ 1u;
                if (mode == 2 || mbar_test_wait(&pipe.empty[stage], parity)) {
                    uint8_t* dst = pipe.stages + stage * SM::STAGE_BYTES;
                    if (mode != 2) {
                        mbar_arrive_expect_tx(&pipe.full[stage], SM::STAGE_BYTES);
                        bulk_g2s(dst, wsrc + (size_t)c * TC_CK * GN, SM::W_BYTES, &pipe.full[stage]);
                    }
                    if (mode != 1)
                        bulk_g2s(dst + SM::W_BYTES, abuf + (size_t)(achunk0 + c) * SM::A_BYTES, SM::A_BYTES, &pipe.full[stage]);
                    c += STAGES;
                } else if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) {
                    spin_fail();
                }
            }
        }
    }

    // MMA-issuing thread: D[acc] (+)= A[chunk] . W[chunk]^T for the n chunks at ring positions g..; fresh = the
    // first MMA overwrites the accumulator.
    __device__ static TACO_NI uint32_t consume(const Pipe pipe, uint32_t g, int n, uint32_t acc, bool fresh) {
        constexpr uint32_t idesc = umma::instr_desc_bf16_f32(128, GN);
        TACO_ROLLED
        for (int c = 0; c < n; ++c, ++g) {
            const uint32_t stage = g % STAGES, parity = (g / STAGES) & 1u;
            mbar_wait(&pipe.full[stage], parity);
            umma::fence_after_sync();
            const uint32_t w_addr = smem_u32(pipe.stages + stage * SM::STAGE_BYTES);
            const uint32_t a_addr = w_addr + SM::W_BYTES;
#pragma unroll
            for (int ks = 0; ks < TC_CK / 16; ++ks)
                umma::mma_bf16(acc, umma::smem_desc(a_addr + ks * 256, 128, TC_KG * 128),
                               umma::smem_desc(w_addr + ks * 256, 128, TC_KG * 128), idesc, !(fresh && c == 0 && ks == 0));
            umma::commit(&pipe.empty[stage]);       // stage is free once these MMAs have read it
        }
        return g;
    }
};

// Consumer warps: accumulator (2*BP hi|lo rows x 128 gate rows of the group, this CTA's K share) -> split-K
// exchange with the three other members -> z[row][b] of this CTA's own 8 hidden units.
// CL (group = thread-block cluster): the partial of member `dst` is stored straight into dst's shared memory
// (DSMEM), followed by one remote mbarrier arrive per peer; no L2 round trip, no polling on a global counter.
// !CL: partials through L2 (zpart) and a release/acquire group counter.
// acc_phase / gepoch: the caller's counters BEFORE this exchange (it flips / increments them afterwards)
template <int BP, bool CL>
__device__ TACO_NI void tc_epilogue(uint32_t acc, uint64_t* acc_bar, uint32_t acc_phase, float* red, float* z,
                                    float* zpart, uint32_t* gcnt, uint32_t gepoch, float* xch, uint64_t* xbar,
                                    int cta, int warp, int lane, int tid) {
    const int member = cta % GS, group = cta / GS;
    constexpr int EW = 2 * BP / 32;                     // warps whose TMEM lane quarter holds rows
    constexpr int PB = BP <= 32 ? 4 : 2;                // 32-column blocks staged in `red` per pass
    constexpr int RS = PB * 32 + 1;                     // red row stride
    constexpr int SLICE = ROWS_PER_CTA * BP;            // floats of one member's partial
    static_assert(2 * BP * RS * 4 <= Smem<BP, __nv_bfloat16, true>::RED_BYTES, "epilogue staging must fit the red region");
    if (warp < EW) {
        mbar_wait(acc_bar, acc_phase);
        umma::fence_after_sync();
    }
    const uint32_t xch_local = smem_u32(xch);
#pragma unroll 1
    for (int pass = 0; pass < GS / PB; ++pass) {
        if (warp < EW) {
            const int m = warp * 32 + lane;
            TACO_ROLLED
            for (int cbl = 0; cbl < PB; ++cbl) {
                float v[32];
                umma::tmem_ld_32x32(acc + ((uint32_t)(warp * 32) << 16) + (pass * PB + cbl) * 32, v);
#pragma unroll
                for (int c = 0; c < 32; ++c) red[m * RS + cbl * 32 + c] = v[c];
            }
            umma::fence_before_sync();
        }
        consumer_sync();
        // hi + lo rows; columns [32*dst, 32*dst+32) are the gate rows of member dst
        for (int i = tid; i < PB * SLICE; i += NTC) {
            const int b = i % BP, c = (i / BP) % ROWS_PER_CTA, cbl = i / SLICE;
            const int dst = pass * PB + cbl;
            const float v = red[b * RS + cbl * 32 + c] + red[(BP + b) * RS + cbl * 32 + c];
            if constexpr (CL) {
                // slot `member` of dst's exchange buffer (the own slot is a plain local store through the same window)
                st_cluster_f32(mapa_u32(xch_local + (uint32_t)(member * SLICE + c * BP + b) * 4u, (uint32_t)dst), v);
            } else {
                if (dst == member) z[c * BP + b] = v;       // own slice stays in shared memory
                else zpart[(((size_t)(group * GS + dst)) * GS + member) * SLICE + c * BP + b] = v;
            }
        }
        consumer_sync();
    }
    gepoch += 1;
    if constexpr (CL) {
        // the stores of all consumer threads are ordered before thread 0's release by the bar.sync above
        if (tid == 0) {
            const uint32_t bar_local = smem_u32(xbar);
#pragma unroll
            for (int r = 1; r < GS; ++r) mbar_arrive_remote_release(mapa_u32(bar_local, (uint32_t)((member + r) % GS)));
        }
        const uint32_t parity = (gepoch - 1u) & 1u;
        if (!mbar_try_wait_acquire_cluster(xbar, parity)) {
            const uint64_t t0 = global_timer_ns();
            while (!mbar_try_wait_acquire_cluster(xbar, parity)) {
                if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) spin_fail();
            }
        }
        for (int i = tid; i < SLICE; i += NTC) {
            float accv = 0.f;
#pragma unroll
            for (int src = 0; src < GS; ++src) accv += xch[src * SLICE + i];      // fixed order: bitwise reproducible
            z[i] = accv;
        }
        consumer_sync();
    } else {
        // group rendezvous: all four partial slices are in L2
        if (tid == 0) {
            red_release_add_u32(gcnt + group, 1u);
            const uint32_t target = gepoch * GS;
            if (ld_acquire_u32(gcnt + group) < target) {
                const uint64_t t0 = global_timer_ns();
                while (ld_acquire_u32(gcnt + group) < target) {
                    if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) spin_fail();
                }
            }
        }
        consumer_sync();
        for (int i = tid; i < SLICE; i += NTC) {
            float accv = 0.f;
#pragma unroll
            for (int src = 0; src < GS; ++src) {            // fixed order: bitwise reproducible
                const float v = src == member ? z[i] : __ldcg(zpart + ((size_t)cta * GS + src) * SLICE + i);
                accv += v;
            }
            red[i] = accv;
        }
        consumer_sync();
        for (int i = tid; i < SLICE; i += NTC) z[i] = red[i];
        consumer_sync();
    }
}

// ST: utterance queue (continuous batching) compiled in; the plain instantiation carries none of its code
template <int BP, bool CL, bool ST>
__global__ void __launch_bounds__(NT_TC, 1) decoder_tc_kernel(const KParams p) {
    using WT = __nv_bfloat16;
    using SM = Smem<BP, WT, true>;
    using TP = TcPipe<BP>;
    constexpr int STAGES = SM::NSTAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const bool is_gemm = cta < GEMM_CTAS;
    const bool is_producer = warp == TC_PRODUCER_WARP, is_mma = warp == TC_MMA_WARP;
    const bool is_consumer = warp < NW;
    const int B = p.B;

    Pipe pipe;
    pipe.full = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    pipe.empty = pipe.full + STAGES;
    pipe.stages = smem + SM::OFF_STAGE;
    float* red = reinterpret_cast<float*>(smem + SM::OFF_RED);
    float* z_s = reinterpret_cast<float*>(smem + SM::OFF_Z);
    float* ca_s = reinterpret_cast<float*>(smem + SM::OFF_CA);
    float* cd_s = reinterpret_cast<float*>(smem + SM::OFF_CD);
    float* h_s = reinterpret_cast<float*>(smem + SM::OFF_H);
    float* bias_s = reinterpret_cast<float*>(smem + SM::OFF_BIAS);
    float* wq_s = reinterpret_cast<float*>(smem + SM::OFF_WQ);
    float* wproj_s = reinterpret_cast<float*>(smem + SM::OFF_WPROJ);
    float* wpc_s = wproj_s + 656;
    float* cx_s = h_s + 8 * BP;
    float* wloc_s = reinterpret_cast<float*>(smem + SM::OFF_WLOC);
    float* wld_s = reinterpret_cast<float*>(smem + SM::OFF_WLD);
    float* v_s = reinterpret_cast<float*>(smem + SM::OFF_V);
    int* misc_s = reinterpret_cast<int*>(smem + SM::OFF_MISC);
    uint64_t* bar_attn = reinterpret_cast<uint64_t*>(smem + SM::OFF_ACCBAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::OFF_TMEM);
    uint64_t* bar_dec = reinterpret_cast<uint64_t*>(smem + SM::OFF_ACCBAR + 16);
    uint64_t* bar_exit = reinterpret_cast<uint64_t*>(smem + SM::OFF_ACCBAR + 24);
    uint64_t* xbar = reinterpret_cast<uint64_t*>(smem + SM::OFF_ACCBAR + 32);
    uint64_t* xbar2 = reinterpret_cast<uint64_t*>(smem + SM::OFF_ACCBAR + 40);
    float* xch = reinterpret_cast<float*>(smem + SM::OFF_XCH);
    int* rows_s = reinterpret_cast<int*>(smem + SM::OFF_ROWS);
    constexpr bool streaming = ST;
    const int* const rows = ST ? rows_s : nullptr;
    static_assert(SM::OFF_ACCBAR + 48 <= 256, "barrier area");
    static_assert(GS * APART * 4 <= SM::XCH_BYTES, "attention partials of the slice siblings alias the exchange buffer");

    // ---- one-time setup ----
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&pipe.full[s], 1);
            mbar_init(&pipe.empty[s], 1);
        }
        mbar_init(bar_attn, 1);
        mbar_init(bar_dec, 1);
        mbar_init(bar_exit, 1);
        mbar_init(xbar, GS - 1);                 // one remote arrive per peer and exchange
        mbar_init(xbar2, p.NS > 1 ? p.NS - 1 : 1);   // attention merge: one arrive per sibling slice and step
        fence_mbar_init();
        misc_s[1] = 0;
    }
    if (is_gemm && warp == 0) umma::tmem_alloc(tmem_slot, 2 * GN);          // two accumulators: 128 lanes x (128 + 128) columns
    umma::fence_before_sync();
    for (int i = tid; i < UNITS_PER_CTA * BP; i += NT_TC) { ca_s[i] = 0.f; cd_s[i] = 0.f; h_s[i] = 0.f; }
    if (is_gemm) {
        for (int i = tid; i < 32; i += NT_TC) {
            bias_s[i] = p.bias_attn[cta * 32 + i];
            bias_s[32 + i] = p.bias_dec[cta * 32 + i];
        }
        for (int i = tid; i < 128 * 8; i += NT_TC) wq_s[i] = p.wq_slices[(size_t)cta * 1024 + i];
        for (int i = tid; i < PROJ_OUT * 8; i += NT_TC) wproj_s[i] = p.wproj_slices[(size_t)cta * PROJ_OUT * 8 + i];
        for (int i = tid; i < PROJ_OUT * 4; i += NT_TC) wpc_s[i] = p.wproj_ctxT[(size_t)(4 * cta + (i & 3)) * PROJ_PAD + (i >> 2)];
    }
    for (int i = tid; i < LOC_F * 2 * LOC_K; i += NT_TC) wloc_s[(i / 62) * 63 + (i % 62)] = p.loc_conv[i];
    for (int i = tid; i < LOC_F * ATT_DIM; i += NT_TC) wld_s[i] = p.loc_denseT[i];
    for (int i = tid; i < 130; i += NT_TC) v_s[i] = p.attn_v[i];
    // utterance queue: row b starts with utterance b at step 0 (start steps and the queue head are zero-initialised)
    if (streaming && cta == 0 && tid < B) p.ctrl[CTRL_UTT + tid] = (uint32_t)tid;
    __syncthreads();
    if constexpr (CL) cluster_sync_all();        // peers' mbarriers are initialised before anyone arrives on them
    umma::fence_after_sync();
    const uint32_t tmem = is_gemm ? *tmem_slot : 0u;
    const uint32_t acc_attn = tmem, acc_dec = tmem + GN;
    uint32_t ph_attn = 0, ph_dec = 0, gepoch = 0;

    // this CTA's weights: [local chunk j][128 rows x 64 k tile]; jobs are contiguous ranges of local chunks
    const int member = cta % GS;
    constexpr int NCH_A = K_ATTN / TC_CK / GS, NCH_D = K_DEC / TC_CK / GS;         // 7, 10
    constexpr size_t WCH = (size_t)TC_CK * GN;                                        // elements per weight chunk
    const WT* w_attn = reinterpret_cast<const WT*>(p.w_attn_tc) + (size_t)cta * NCH_A * WCH;
    const WT* w_dec = reinterpret_cast<const WT*>(p.w_dec_tc) + (size_t)cta * NCH_D * WCH;
    // activation chunks (64 rows of the act vector) behind each job, dims.cuh
    const int ch_p = tc_attn_act_chunk(member, 0), ch_ctx = tc_attn_act_chunk(member, 1), ch_ah = tc_attn_act_chunk(member, 3),
              ch_dh = tc_dec_act_chunk(member, 6);
    constexpr size_t ABUF_BYTES = (size_t)(ACT_ROWS / TC_CK) * SM::A_BYTES;
    uint8_t* const abase = reinterpret_cast<uint8_t*>(p.act16);

    uint32_t epoch = 0;
    uint32_t g = 0;             // ring position: chunks issued (producer) / consumed (MMA thread)
    const bool tf = p.mode == 1;
    // utterance queue: every utterance takes at most max_steps + 1 decode steps of one row
    const int total_steps = tf ? p.max_steps
                               : (streaming ? (p.n_utts + B - 1) / B * (p.max_steps + 1) : p.max_steps + 1);
    int steps_done = 0;

    const bool timing = cta == 0 && tid == 0;
    unsigned long long t_phase[4] = {0ull, 0ull, 0ull, 0ull}, t_last = timing ? global_timer_ns() : 0ull;
#ifdef TACO_PROFILE_FINE
    unsigned long long t_fine[16] = {}, t_flast = 0ull;
#endif
    auto stamp = [&](int ph) {
        if (timing) {
            const unsigned long long now = global_timer_ns();
            t_phase[ph] += now - t_last;
            t_last = now;
        }
    };

    // ---- prologue = the background jobs of a virtual step -1 (all activations are zero): establishes the
    // invariant "at the start of phase A the attention accumulator holds the ah/ctx ranges".
    if (is_gemm) {
        const uint8_t* a0 = abase + ABUF_BYTES;          // buffer 1 = prv of step 0 (zero-initialised)
        if (is_producer) {
            TP::produce(pipe, g, 4, 0, lane, w_attn + 3 * WCH, a0, ch_ah); g += 4;
            TP::produce(pipe, g, 2, 0, lane, w_attn + 1 * WCH, a0, ch_ctx); g += 2;
            TP::produce(pipe, g, 1, 1, lane, w_attn, nullptr, 0);            // weights of the first critical chunk
        } else if (is_mma && lane == 0) {
            g = TP::consume(pipe, g, 4, acc_attn, true);
            g = TP::consume(pipe, g, 2, acc_attn, false);
        }
    }

    for (int s = 0; s < total_steps; ++s) {
        const int cur = s & 1, prv = cur ^ 1;
        float* act_cur = p.act + (size_t)cur * ACT_ROWS * BP;
        const uint8_t* a_cur = abase + (size_t)cur * ABUF_BYTES;
        const uint8_t* a_prv = abase + (size_t)prv * ABUF_BYTES;
        __nv_bfloat16* act16_cur = p.act16 + (size_t)cur * (ABUF_BYTES / 2);
        __nv_bfloat16* act16_prv = p.act16 + (size_t)prv * (ABUF_BYTES / 2);
        const int out_idx = tf ? s : s - 1;          // F6: decode step 0 of inference is discarded
        // utterance queue: this step's row table (rows that stopped in the previous step's phase D were re-assigned
        // there, before the barrier).  Loaded by the last two consumer warps - not TMEM readers of the epilogue at
        // BP = 32 - while the critical chunk is in flight; the first use follows a consumer / CTA barrier.
        if (streaming && tid >= NTC - MAX_BATCH && tid < NTC) {
            const int i = tid - (NTC - MAX_BATCH);
            if (i < B) {
                const int u = s == 0 ? i : (int)__ldcg(p.ctrl + CTRL_UTT + i);
                rows_s[i] = u;
                rows_s[ROW_START + i] = s == 0 ? 0 : (int)__ldcg(p.ctrl + CTRL_START + i);
                const int ua = u >= 0 ? u : -1 - u;
                rows_s[ROW_ABASE + i] = ua / B * p.max_steps * B + ua % B;
            }
        }

        // =============================== A: attention LSTM (critical: prenet rows) ===============================
        if (is_gemm) {
            if (is_producer) {
                fence_proxy_async();
                // prenet rows: previous buffer, or the precomputed buffer when teacher forcing (4 chunks per step)
                const uint8_t* a0 = tf ? static_cast<const uint8_t*>(p.p_all) + (size_t)s * (PRENET / TC_CK) * SM::A_BYTES : a_prv;
                TP::produce(pipe, g, 1, 2, lane, nullptr, a0, ch_p); g += 1;
                TP::produce(pipe, g, 4, 0, lane, w_dec + 6 * WCH, a_cur, ch_dh); g += 4;   // background: decoder LSTM = W[dh] . dh_{t-1}
            } else if (is_mma) {
                if (lane == 0) {
                    g = TP::consume(pipe, g, 1, acc_attn, false);
                    umma::commit(bar_attn);
                    g = TP::consume(pipe, g, 4, acc_dec, true);
                }
                __syncwarp();
            } else {
                FT_RESET();
                tc_epilogue<BP, CL>(acc_attn, bar_attn, ph_attn, red, z_s, p.zpart, p.ctrl + CTRL_GCNT, gepoch, xch, xbar, cta, warp,
                                    lane, tid);
                ph_attn ^= 1u; gepoch += 1;
                FT(9);
                lstm_pointwise_tc<BP>(z_s, bias_s, ca_s, h_s, act16_cur, ACT_AH + cta * UNITS_PER_CTA, B, tid,
                                       pw_save(p.sv_ga, p.sv_ca, p.sv_ah, p.mask_a, p.drop_scale, s, B, cta * UNITS_PER_CTA),
                                       rows, s, true);
                FT(10);
                // this CTA's slice of the query projection (src/model.py:123): pqpart[b][cta][a]
                for (int i = tid; i < ATT_DIM * B; i += NTC) {
                    const int a = i % ATT_DIM, b = i / ATT_DIM;
                    float sum = 0.f;
#pragma unroll
                    for (int u = 0; u < UNITS_PER_CTA; ++u) sum = fmaf(wq_s[a * 8 + u], h_s[u * BP + b], sum);
                    p.pqpart[((size_t)b * GEMM_CTAS + cta) * ATT_DIM + a] = sum;
                }
                FT(11);
            }
        }
        grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
        stamp(0);

        // =============================== B: attention; background: ah_t ranges ===============================
        if (is_gemm && is_producer) {
            fence_proxy_async();
            wait_signal(misc_s + 1, 2 * s + 1);      // let the attention's own loads go first
            TP::produce(pipe, g, 4, 0, lane, w_dec + 2 * WCH, a_cur, ch_ah); g += 4;     // decoder LSTM   += W[ah] . ah_t
            TP::produce(pipe, g, 2, 1, lane, w_dec, nullptr, 0);                         // weights of phase C's critical chunks
        } else if (is_gemm && is_mma) {
            if (lane == 0) g = TP::consume(pipe, g, 4, acc_dec, false);
            __syncwarp();
        } else if (is_consumer) {
            attention_phase<BP, true, CL>(p, red, wld_s, wloc_s, v_s, misc_s, cur, out_idx, cta, ncta, tid, warp, lane, act_cur,
                                          act16_cur, timing, TACO_FINE_ARGS, misc_s + 1, 2 * s + 1, xch, xbar2, false, rows, s);
        }
        grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
        stamp(1);

        // =============================== C: decoder LSTM (critical: context rows) ===============================
        if (is_gemm) {
            if (is_producer) {
                fence_proxy_async();
                TP::produce(pipe, g, 2, 2, lane, nullptr, a_cur, ch_ctx); g += 2;
                // background: attention LSTM(t+1) = W[ah] . ah_t + W[ctx] . ctx_t ; then the weights of the next phase A
                TP::produce(pipe, g, 4, 0, lane, w_attn + 3 * WCH, a_cur, ch_ah); g += 4;
                TP::produce(pipe, g, 2, 0, lane, w_attn + 1 * WCH, a_cur, ch_ctx); g += 2;
                TP::produce(pipe, g, 1, 1, lane, w_attn, nullptr, 0);
            } else if (is_mma) {
                if (lane == 0) {
                    g = TP::consume(pipe, g, 2, acc_dec, false);
                    umma::commit(bar_dec);
                    g = TP::consume(pipe, g, 4, acc_attn, true);
                    g = TP::consume(pipe, g, 2, acc_attn, false);
                }
                __syncwarp();
            } else {
                for (int i = tid; i < 4 * B; i += NTC) cx_s[(i & 3) * BP + (i >> 2)] = __ldcg(p.ctx_bj + (size_t)(i >> 2) * EMB + 4 * cta + (i & 3));
                FT_RESET();
                tc_epilogue<BP, CL>(acc_dec, bar_dec, ph_dec, red, z_s, p.zpart, p.ctrl + CTRL_GCNT, gepoch, xch, xbar, cta, warp,
                                    lane, tid);
                ph_dec ^= 1u; gepoch += 1;
                FT(12);
                lstm_pointwise_tc<BP>(z_s, bias_s + 32, cd_s, h_s, act16_prv, ACT_DH + cta * UNITS_PER_CTA, B, tid,
                                       pw_save(p.sv_gd, p.sv_cd, p.sv_dh, p.mask_d, p.drop_scale, s, B, cta * UNITS_PER_CTA),
                                       rows, s, false);
                FT(13);
                // this CTA's slice of linear_projection / gate_layer (src/model.py:272-274): 8 decoder-h units + 4 context rows
                for (int i = tid; i < PROJ_PAD * B; i += NTC) {
                    const int o = i % PROJ_PAD, b = i / PROJ_PAD;
                    if (o < PROJ_OUT) {
                        float sum = 0.f;
#pragma unroll
                        for (int u = 0; u < UNITS_PER_CTA; ++u) sum = fmaf(wproj_s[o * 8 + u], h_s[u * BP + b], sum);
#pragma unroll
                        for (int j = 0; j < 4; ++j) sum = fmaf(wpc_s[o * 4 + j], cx_s[j * BP + b], sum);
                        p.melpart[((size_t)b * GEMM_CTAS + cta) * PROJ_PAD + o] = sum;
                    }
                }
            }
        }
        grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
        stamp(2);

        // =============================== D: projection sums, stop rule, prenet; background: dh_t range ============
        if (is_consumer)
            projection_phase<BP, true>(p, red, tf, out_idx, cta, ncta, tid, act_cur, act16_cur, nullptr, 0, false, rows, s,
                                       act16_prv);
        steps_done = s + 1;
        if (!tf) {
            grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
            stamp(3);
            if (__ldcg(p.ctrl + CTRL_STOP) != 0u) break;
        }
    }

    // ---- drain: background MMAs may still be running; the weights of a phase-A chunk that will never run are
    // in flight: wait for the former, complete the stage of the latter ----
    if (is_gemm && is_mma && lane == 0) {
        umma::commit(bar_exit);
        mbar_wait(bar_exit, 0);
    }
    if (is_gemm && is_producer) {
        const uint32_t stage = g % STAGES;
        if (lane == (int)stage) {
            bulk_g2s(pipe.stages + stage * SM::STAGE_BYTES + SM::W_BYTES, abase, SM::A_BYTES, &pipe.full[stage]);
            mbar_wait(&pipe.full[stage], (g / STAGES) & 1u);
        }
    }
    if (cta == 0 && tid == 0) {
        unsigned long long* prof = reinterpret_cast<unsigned long long*>(p.ctrl + CTRL_PROF);
        for (int i = 0; i < 4; ++i) prof[i] = t_phase[i];
        prof[4] = (unsigned long long)steps_done;
#ifdef TACO_PROFILE_FINE
        for (int i = 0; i < 16; ++i) prof[5 + i] = t_fine[i];
#endif
        const int S = tf ? p.max_steps : steps_done - 1;     // output frames (decode step 0 discarded)
        *p.steps_run = streaming ? steps_done : S;           // utterance queue: decode steps of the whole launch
        for (int b = 0; b < B && !streaming; ++b) {
            int len = S;
            if (!tf && p.stop_mode == B200TACO_STOP_PER_UTTERANCE && __ldcg(p.ctrl + CTRL_DONE + b))
                len = (int)__ldcg(p.ctrl + CTRL_LEN + b);
            p.lengths[b] = len;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if constexpr (CL) cluster_sync_all();        // no CTA leaves while a peer may still write into its shared memory
    if (is_gemm && warp == 0) umma::tmem_dealloc(tmem, 2 * GN);
}

// Cooperative launch with the split-K groups as thread-block clusters of GS CTAs.  The grid is the largest
// multiple of GS that is co-resident (B200: 33 clusters = 132 CTAs at this shared-memory size).
template <int BP>
int tc_cluster_ctas(int num_sms) {
    using SM = Smem<BP, __nv_bfloat16, true>;
    auto kern = decoder_tc_kernel<BP, true, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms / GS * GS);
    cfg.blockDim = dim3(NT_TC);
    cfg.dynamicSmemBytes = SM::TOTAL;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = GS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    const int ctas = ncl * GS;
    return ctas > num_sms / GS * GS ? num_sms / GS * GS : ctas;
}

template <int BP>
int launch_tc(const KParams& kp, int ncta, bool clusters, cudaStream_t stream) {
    using SM = Smem<BP, __nv_bfloat16, true>;
    if (clusters) {
        auto kern = kp.n_utts > 0 ? decoder_tc_kernel<BP, true, true> : decoder_tc_kernel<BP, true, false>;
        TACO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ncta);
        cfg.blockDim = dim3(NT_TC);
        cfg.dynamicSmemBytes = SM::TOTAL;
        cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = GS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative;
        at[1].val.cooperative = 1;
        // Nsight Compute cannot launch a grid that is BOTH cooperative and clustered (the driver answers LaunchFailed and the
        // tool kills the process; measured with kernel and application replay), but it profiles the same grid as a plain cluster
        // launch.  Under the tool's injection (it exports NV_NSIGHT_INJECTION_TRANSPORT_TYPE / NV_TPS_LAUNCH_TOKEN to its target),
        // or with TACO_TC_COOP=0, the cooperative attribute is dropped: the grid was sized by cudaOccupancyMaxActiveClusters and
        // a profiler serialises kernels, so every CTA is still co-resident.
        static const bool coop = [] {
            const char* e = getenv("TACO_TC_COOP");
            if (e) return e[0] != '0';
            return getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") == nullptr && getenv("NV_TPS_LAUNCH_TOKEN") == nullptr;
        }();
        cfg.attrs = at; cfg.numAttrs = coop ? 2 : 1;
        TACO_CUDA(cudaLaunchKernelEx(&cfg, kern, kp));
        launch_counter()++;
        return B200TACO_OK;
    }
    auto kern = kp.n_utts > 0 ? decoder_tc_kernel<BP, false, true> : decoder_tc_kernel<BP, false, false>;
    TACO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    int per_sm = 0;
    TACO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT_TC, SM::TOTAL));
    TACO_REQUIRE(per_sm >= 1, B200TACO_ERR_UNSUPPORTED, "decoder kernel does not fit on an SM (smem %d B)", SM::TOTAL);
    void* args[] = {const_cast<KParams*>(&kp)};
    TACO_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(ncta), dim3(NT_TC), args, SM::TOTAL, stream));
    launch_counter()++;
    return B200TACO_OK;
}


// =======================================================================================================
// SIMT decoder kernel with HOISTED gate contractions (batch tile <= 16: B <= 3, or fp32 weights / forced SIMT).
// Same idea as decoder_tc_kernel, with fp32 register accumulators instead of TMEM: each CTA keeps the partial
// pre-activations of its 32 gate rows for BOTH LSTMs in registers across the phases and adds every K range as
// soon as its activations exist:
//     phase A: += W_a[prenet] . p_t            -> attention LSTM cell      (1792 -> 256 rows on the critical path)
//     phase B: += W_d[ah] . ah_t,  W_a[ah] . ah_t   (background of the attention phase)
//     phase C: += W_d[ctx] . ctx_t             -> decoder LSTM cell        (2560 -> 512 rows on the critical path)
//     phase D: += W_a[ctx] . ctx_t, W_d[dh] . dh_t  (background of the projection phase)
// Attention / projection work items are dealt from the last CTA downwards (the >= 20 CTAs beyond the GEMM CTAs
// first), so at B=1 the GEMM CTAs have nothing else to do in phases B and D.
template <typename SM>
struct SimtPipe {
    static constexpr int STAGES = SM::NSTAGES;
    // producer warp, all 32 lanes; mode 0: weights + activations, 1: weights only, 2: activations only
    template <typename WT>
    __device__ static __forceinline__ void produce(const Pipe& pipe, uint32_t g0, int n, int mode, int lane, const WT* wsrc,
                                                   const uint8_t* abuf, int achunk0) {
        int c = lane < STAGES ? (lane + STAGES - (int)(g0 % STAGES)) % STAGES : n;
        const uint64_t t0 = global_timer_ns();
        while (__any_sync(0xffffffffu, c < n)) {
            if (c < n) {
                const uint32_t g = g0 + c, stage = g % STAGES, parity = ((g / STAGES) & 1u) ^ 1u;
                if (mode == 2 || mbar_test_wait(&pipe.empty[stage], parity)) {
                    uint8_t* dst = pipe.stages + stage * SM::STAGE_BYTES;
                    if (mode != 2) {
                        mbar_arrive_expect_tx(&pipe.full[stage], SM::STAGE_BYTES);
                        bulk_g2s(dst, reinterpret_cast<const uint8_t*>(wsrc) + (size_t)c * SM::W_BYTES, SM::W_BYTES, &pipe.full[stage]);
                    }
                    if (mode != 1)
                        bulk_g2s(dst + SM::W_BYTES, abuf + (size_t)(achunk0 + c) * SM::A_BYTES, SM::A_BYTES, &pipe.full[stage]);
                    c += STAGES;
                } else if (global_timer_ns() - t0 > SPIN_TIMEOUT_NS) {
                    spin_fail();
                }
            }
        }
    }
};

template <int BP, typename WT>
__global__ void __launch_bounds__(NT, 1) decoder_simt_kernel(const KParams p) {
    using SM = Smem<BP, WT, false>;
    using SP = SimtPipe<SM>;
    using TL = Tile<BP, sizeof(WT)>;
    constexpr int STAGES = SM::NSTAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const bool is_gemm = cta < GEMM_CTAS;
    const bool is_producer = warp == NW;
    const int B = p.B;

    Pipe pipe;
    pipe.full = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    pipe.empty = pipe.full + STAGES;
    pipe.stages = smem + SM::OFF_STAGE;
    float* red = reinterpret_cast<float*>(smem + SM::OFF_RED);
    float* z_s = reinterpret_cast<float*>(smem + SM::OFF_Z);
    float* ca_s = reinterpret_cast<float*>(smem + SM::OFF_CA);
    float* cd_s = reinterpret_cast<float*>(smem + SM::OFF_CD);
    float* h_s = reinterpret_cast<float*>(smem + SM::OFF_H);
    float* bias_s = reinterpret_cast<float*>(smem + SM::OFF_BIAS);
    float* wq_s = reinterpret_cast<float*>(smem + SM::OFF_WQ);
    float* wproj_s = reinterpret_cast<float*>(smem + SM::OFF_WPROJ);
    float* wpc_s = wproj_s + 656;
    float* cx_s = h_s + 8 * BP;
    float* wloc_s = reinterpret_cast<float*>(smem + SM::OFF_WLOC);
    float* wld_s = reinterpret_cast<float*>(smem + SM::OFF_WLD);
    float* v_s = reinterpret_cast<float*>(smem + SM::OFF_V);
    int* misc_s = reinterpret_cast<int*>(smem + SM::OFF_MISC);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&pipe.full[s], 1);
            mbar_init(&pipe.empty[s], NW);
        }
        fence_mbar_init();
    }
    for (int i = tid; i < UNITS_PER_CTA * BP; i += NT) { ca_s[i] = 0.f; cd_s[i] = 0.f; h_s[i] = 0.f; }
    if (is_gemm) {
        for (int i = tid; i < 32; i += NT) {
            bias_s[i] = p.bias_attn[cta * 32 + i];
            bias_s[32 + i] = p.bias_dec[cta * 32 + i];
        }
        for (int i = tid; i < 128 * 8; i += NT) wq_s[i] = p.wq_slices[(size_t)cta * 1024 + i];
        for (int i = tid; i < PROJ_OUT * 8; i += NT) wproj_s[i] = p.wproj_slices[(size_t)cta * PROJ_OUT * 8 + i];
        for (int i = tid; i < PROJ_OUT * 4; i += NT) wpc_s[i] = p.wproj_ctxT[(size_t)(4 * cta + (i & 3)) * PROJ_PAD + (i >> 2)];
    }
    for (int i = tid; i < LOC_F * 2 * LOC_K; i += NT) wloc_s[(i / 62) * 63 + (i % 62)] = p.loc_conv[i];
    for (int i = tid; i < LOC_F * ATT_DIM; i += NT) wld_s[i] = p.loc_denseT[i];
    for (int i = tid; i < 130; i += NT) v_s[i] = p.attn_v[i];
    __syncthreads();

    // chunk geometry: K ranges of the act vector [prenet | ctx | ah | dh] in chunks of CK rows
    constexpr int CK = SM::CK;
    constexpr int NP = PRENET / CK, NC = EMB / CK, NH = RNN / CK;
    static_assert(PRENET % CK == 0 && EMB % CK == 0 && RNN % CK == 0, "K ranges must be whole chunks");
    constexpr int CH_P = 0, CH_CTX = NP, CH_AH = NP + NC, CH_DH = NP + NC + NH;        // act chunk of each range
    constexpr size_t WCH = (size_t)CK * ROWS_PER_CTA;                                    // weight elements per chunk
    const WT* w_attn = reinterpret_cast<const WT*>(p.w_attn) + (size_t)cta * K_ATTN * ROWS_PER_CTA;   // [p | ctx | ah]
    const WT* w_dec = reinterpret_cast<const WT*>(p.w_dec) + (size_t)cta * K_DEC * ROWS_PER_CTA;      // [ctx | ah | dh]
    const WT* wa_p = w_attn, *wa_ctx = w_attn + NP * WCH, *wa_ah = w_attn + (NP + NC) * WCH;
    const WT* wd_ctx = w_dec, *wd_ah = w_dec + NC * WCH, *wd_dh = w_dec + (NC + NH) * WCH;
    constexpr size_t ABUF_BYTES = (size_t)ACT_ROWS * BP * 4;
    uint8_t* const abase = reinterpret_cast<uint8_t*>(p.act);

    uint32_t epoch = 0;
    uint32_t g = 0;             // ring position
    const bool tf = p.mode == 1;
    const int total_steps = tf ? p.max_steps : p.max_steps + 1;
    int steps_done = 0;

    const bool timing = cta == 0 && tid == 0;
    unsigned long long t_phase[4] = {0ull, 0ull, 0ull, 0ull}, t_last = timing ? global_timer_ns() : 0ull;
#ifdef TACO_PROFILE_FINE
    unsigned long long t_fine[16] = {}, t_flast = 0ull;
#endif
    auto stamp = [&](int ph) {
        if (timing) {
            const unsigned long long now = global_timer_ns();
            t_phase[ph] += now - t_last;
            t_last = now;
        }
    };

    // partial pre-activations of this CTA's 32 gate rows (this lane's share), both LSTMs; all state starts at zero,
    // so the K ranges that a "step -1" would have contributed are zero as well
    float acc_a[8][TL::BPL], acc_d[8][TL::BPL];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < TL::BPL; ++j) { acc_a[r][j] = 0.f; acc_d[r][j] = 0.f; }
    auto zero = [](float (&a)[8][TL::BPL]) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int j = 0; j < TL::BPL; ++j) a[r][j] = 0.f;
    };
    // ---- producer bookkeeping.  The weight stream of a step is a fixed cyclic sequence of L chunks:
    //   [attn p (NP)] | [dec ah (NH)] [attn ah (NH)] | [dec ctx (NC)] | [attn ctx (NC)] [dec dh (NH)]
    //        phase A            phase B                    phase C              phase D
    // Weights never depend on the step, so they run ahead of the activation tiles by up to the ring depth
    // (gw = chunks whose weights are issued, ga = chunks whose activation tile is issued, gw <= ga + STAGES: the
    // stage of chunk q is free once chunk q - STAGES was consumed, which needs no further help from the producer
    // iff its activations are already issued).
    constexpr int L = NP + 3 * NH + 2 * NC;
    constexpr int Q_B = NP, Q_C = NP + 2 * NH, Q_D = Q_C + NC;
    uint32_t gw = 0, ga = 0;
    // issue chunks [q0, q1) (global chunk numbers), mode as in SimtPipe::produce; a_cur/a_prv/p_all of step `st`
    auto issue = [&](uint32_t q0, uint32_t q1, int mode, int st) {
        const uint8_t* ac = abase + (size_t)(st & 1) * ABUF_BYTES;
        const uint8_t* ap = abase + (size_t)((st & 1) ^ 1) * ABUF_BYTES;
        while (q0 < q1) {
            const int q = (int)(q0 % L);
            const WT* w; const uint8_t* ab; int ach, jend;
            if (q < Q_B) { w = wa_p + (size_t)q * WCH; ab = tf ? static_cast<const uint8_t*>(p.p_all) + (size_t)st * NP * SM::A_BYTES : ap; ach = CH_P + q; jend = Q_B; }
            else if (q < Q_B + NH) { w = wd_ah + (size_t)(q - Q_B) * WCH; ab = ac; ach = CH_AH + q - Q_B; jend = Q_B + NH; }
            else if (q < Q_C) { w = wa_ah + (size_t)(q - Q_B - NH) * WCH; ab = ac; ach = CH_AH + q - Q_B - NH; jend = Q_C; }
            else if (q < Q_D) { w = wd_ctx + (size_t)(q - Q_C) * WCH; ab = ac; ach = CH_CTX + q - Q_C; jend = Q_D; }
            else if (q < Q_D + NC) { w = wa_ctx + (size_t)(q - Q_D) * WCH; ab = ac; ach = CH_CTX + q - Q_D; jend = Q_D + NC; }
            else { w = wd_dh + (size_t)(q - Q_D - NC) * WCH; ab = ap; ach = CH_DH + q - Q_D - NC; jend = L; }
            const uint32_t n = min((uint32_t)(jend - q), q1 - q0);
            SP::produce(pipe, q0, (int)n, mode, lane, w, ab, ach);
            q0 += n;
        }
    };
    // the activation tiles of the chunks of step `st` up to position qend (exclusive) have become available
    auto advance = [&](int st, int qend) {
        const uint32_t a_limit = (uint32_t)st * L + qend;
        if (ga < gw) { const uint32_t e = min(gw, a_limit); issue(ga, e, 2, st); ga = e; }       // weights already in flight
        if (ga < a_limit) { issue(ga, a_limit, 0, st); ga = a_limit; gw = a_limit; }
        const uint32_t w_limit = min(ga + STAGES, (uint32_t)total_steps * L);
        if (gw < w_limit) { issue(gw, w_limit, 1, st); gw = w_limit; }                          // run ahead with weights
    };
    if (is_gemm && is_producer) advance(0, 0);

    for (int s = 0; s < total_steps; ++s) {
        const int cur = s & 1, prv = cur ^ 1;
        float* act_cur = p.act + (size_t)cur * ACT_ROWS * BP;
        float* act_prv = p.act + (size_t)prv * ACT_ROWS * BP;
        const uint8_t* a_cur = abase + (size_t)cur * ABUF_BYTES;
        const uint8_t* a_prv = abase + (size_t)prv * ABUF_BYTES;
        const int out_idx = tf ? s : s - 1;          // F6: decode step 0 of inference is discarded

        // =============================== A: attention LSTM (critical: prenet rows) ===============================
        if (is_gemm) {
            if (is_producer) {
                fence_proxy_async();
                advance(s, Q_B);
            } else {
                FT_RESET();
                gemm_consume<BP, WT, false>(pipe, g, NP, warp, lane, acc_a);
                gemm_reduce<BP>(acc_a, red, z_s, warp, lane, tid);
                zero(acc_a);
                FT(9);
                lstm_pointwise<BP, false>(z_s, bias_s, ca_s, h_s, act_cur + (size_t)(ACT_AH + cta * UNITS_PER_CTA) * BP, nullptr,
                                          ACT_AH + cta * UNITS_PER_CTA, B, tid,
                                       pw_save(p.sv_ga, p.sv_ca, p.sv_ah, p.mask_a, p.drop_scale, s, B, cta * UNITS_PER_CTA));
                FT(10);
                for (int i = tid; i < ATT_DIM * B; i += NTC) {       // query projection slice (src/model.py:123)
                    const int a = i % ATT_DIM, b = i / ATT_DIM;
                    float sum = 0.f;
#pragma unroll
                    for (int u = 0; u < UNITS_PER_CTA; ++u) sum = fmaf(wq_s[a * 8 + u], h_s[u * BP + b], sum);
                    p.pqpart[((size_t)b * GEMM_CTAS + cta) * ATT_DIM + a] = sum;
                }
                FT(11);
            }
        }
        grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
        stamp(0);

        // =============================== B: attention; background: ah_t ranges ===============================
        if (is_producer) {
            if (is_gemm) {
                fence_proxy_async();
                advance(s, Q_C);
            }
        } else {
            attention_phase<BP, false>(p, red, wld_s, wloc_s, v_s, misc_s, cur, out_idx, cta, ncta, tid, warp, lane, act_cur,
                                       nullptr, timing, TACO_FINE_ARGS, nullptr, 0, nullptr, nullptr, true);
            if (is_gemm) {
                gemm_consume<BP, WT, false>(pipe, g, NH, warp, lane, acc_d);
                gemm_consume<BP, WT, false>(pipe, g, NH, warp, lane, acc_a);
            }
        }
        grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
        stamp(1);

        // =============================== C: decoder LSTM (critical: context rows) ===============================
        if (is_gemm) {
            if (is_producer) {
                fence_proxy_async();
                advance(s, Q_D);
            } else {
                for (int i = tid; i < 4 * B; i += NTC) cx_s[(i & 3) * BP + (i >> 2)] = __ldcg(p.ctx_bj + (size_t)(i >> 2) * EMB + 4 * cta + (i & 3));
                FT_RESET();
                gemm_consume<BP, WT, false>(pipe, g, NC, warp, lane, acc_d);
                gemm_reduce<BP>(acc_d, red, z_s, warp, lane, tid);
                zero(acc_d);
                FT(12);
                lstm_pointwise<BP, false>(z_s, bias_s + 32, cd_s, h_s, act_prv + (size_t)(ACT_DH + cta * UNITS_PER_CTA) * BP, nullptr,
                                          ACT_DH + cta * UNITS_PER_CTA, B, tid,
                                       pw_save(p.sv_gd, p.sv_cd, p.sv_dh, p.mask_d, p.drop_scale, s, B, cta * UNITS_PER_CTA));
                FT(13);
                // linear_projection / gate_layer slice (src/model.py:272-274): 8 decoder-h units + 4 context rows
                for (int i = tid; i < PROJ_PAD * B; i += NTC) {
                    const int o = i % PROJ_PAD, b = i / PROJ_PAD;
                    if (o < PROJ_OUT) {
                        float sum = 0.f;
#pragma unroll
                        for (int u = 0; u < UNITS_PER_CTA; ++u) sum = fmaf(wproj_s[o * 8 + u], h_s[u * BP + b], sum);
#pragma unroll
                        for (int j = 0; j < 4; ++j) sum = fmaf(wpc_s[o * 4 + j], cx_s[j * BP + b], sum);
                        p.melpart[((size_t)b * GEMM_CTAS + cta) * PROJ_PAD + o] = sum;
                    }
                }
            }
        }
        grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
        stamp(2);

        // =============================== D: projection sums, stop rule, prenet; background: ctx_t, dh_t ranges =====
        if (is_producer) {
            if (is_gemm) {
                fence_proxy_async();
                advance(s, L);                                                       // dh_t lives in the other buffer
            }
        } else {
            projection_phase<BP, false>(p, red, tf, out_idx, cta, ncta, tid, act_cur, nullptr, nullptr, 0, true);
            if (is_gemm) {
                gemm_consume<BP, WT, false>(pipe, g, NC, warp, lane, acc_a);
                gemm_consume<BP, WT, false>(pipe, g, NH, warp, lane, acc_d);
            }
        }
        steps_done = s + 1;
        if (!tf) {
            grid_barrier(p.ctrl + CTRL_BARRIER, epoch, ncta);
            stamp(3);
            if (__ldcg(p.ctrl + CTRL_STOP) != 0u) break;
        }
    }

    // ---- drain: weights of chunks that will never run are in flight: complete their stages ----
    if (is_gemm && is_producer) {
        for (uint32_t q = ga; q < gw; ++q) {
            const uint32_t stage = q % STAGES;
            if (lane == (int)stage) {
                bulk_g2s(pipe.stages + stage * SM::STAGE_BYTES + SM::W_BYTES, abase, SM::A_BYTES, &pipe.full[stage]);
                mbar_wait(&pipe.full[stage], (q / STAGES) & 1u);
            }
        }
    }
    if (cta == 0 && tid == 0) {
        unsigned long long* prof = reinterpret_cast<unsigned long long*>(p.ctrl + CTRL_PROF);
        for (int i = 0; i < 4; ++i) prof[i] = t_phase[i];
        prof[4] = (unsigned long long)steps_done;
#ifdef TACO_PROFILE_FINE
        for (int i = 0; i < 16; ++i) prof[5 + i] = t_fine[i];
#endif
        const int S = tf ? p.max_steps : steps_done - 1;     // output frames (decode step 0 discarded)
        *p.steps_run = S;
        for (int b = 0; b < B; ++b) {
            int len = S;
            if (!tf && p.stop_mode == B200TACO_STOP_PER_UTTERANCE && __ldcg(p.ctrl + CTRL_DONE + b))
                len = (int)__ldcg(p.ctrl + CTRL_LEN + b);
            p.lengths[b] = len;
        }
    }
}

template <int BP, typename WT>
int launch_simt(const KParams& kp, int ncta, cudaStream_t stream) {
    using SM = Smem<BP, WT, false>;
    auto kern = decoder_simt_kernel<BP, WT>;
    TACO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    int per_sm = 0;
    TACO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, SM::TOTAL));
    TACO_REQUIRE(per_sm >= 1, B200TACO_ERR_UNSUPPORTED, "decoder kernel does not fit on an SM (smem %d B)", SM::TOTAL);
    void* args[] = {const_cast<KParams*>(&kp)};
    TACO_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(ncta), dim3(NT), args, SM::TOTAL, stream));
    launch_counter()++;
    return B200TACO_OK;
}

// prenet outputs Y[s*B + b][256] (fp32) -> the operand layout the chosen path streams per step
__global__ void p_all_pack_simt_kernel(const float* __restrict__ Y, float* __restrict__ P, int B, int BP, int S) {
    const size_t total = (size_t)S * PRENET * BP;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = i % BP, k = (i / BP) % PRENET, s = i / ((size_t)BP * PRENET);
        P[i] = b < B ? Y[((size_t)s * B + b) * PRENET + k] : 0.f;
    }
}
__global__ void p_all_pack_tc_kernel(const float* __restrict__ Y, __nv_bfloat16* __restrict__ P, int B, int BP, int S) {
    const size_t per_step = (size_t)PRENET * 2 * BP;          // 4 chunks x (2*BP rows x 64 k)
    const size_t total = (size_t)S * PRENET * BP;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = i % PRENET, b = (i / PRENET) % BP, s = i / ((size_t)PRENET * BP);
        const float v = b < B ? Y[((size_t)s * B + b) * PRENET + k] : 0.f;
        __nv_bfloat16 hi, lo;
        umma::split_bf16(v, hi, lo);
        const size_t chunk = (size_t)(k / TC_CK) * (2 * BP * TC_CK);
        P[s * per_step + chunk + umma::tile_offset(b, k % TC_CK, TC_KG)] = hi;
        P[s * per_step + chunk + umma::tile_offset(BP + b, k % TC_CK, TC_KG)] = lo;
    }
}

template <typename WT>
int dispatch_bp(int BP, const KParams& kp, int ncta, cudaStream_t stream) {
    switch (BP) {
        // batch tiles 1-2: hoisted contractions (measured B=1: 25.7 -> 22.0 us/step); from 4 rows up the attention /
        // projection work items no longer fit on the CTAs beyond the GEMM CTAs and the one-contraction-per-phase
        // kernel is faster (B=8 forced SIMT: 33 vs 37 us) - with bf16 weights those batches take the tensor-core kernel
        case 1: return launch_simt<1, WT>(kp, ncta, stream);
        case 2: return launch_simt<2, WT>(kp, ncta, stream);
        case 4: return launch<4, WT, false>(kp, ncta, stream);
        case 8: return launch<8, WT, false>(kp, ncta, stream);
        case 16: return launch<16, WT, false>(kp, ncta, stream);
        case 32: return launch<32, WT, false>(kp, ncta, stream);
        case 64: return launch<64, WT, false>(kp, ncta, stream);
    }
    set_error("unsupported batch padding %d", BP);
    return B200TACO_ERR_UNSUPPORTED;
}

struct ScratchLayout {
    size_t act16, zpart;
    size_t act, pqpart, melpart, ctx_bj, attn_prev, attn_cum, wtmp, apart, ctrl, res, total;
    size_t zero_begin, zero_end;
};

ScratchLayout scratch_layout(int B, int T) {
    const int BP = decoder_batch_pad(B);
    const int Tp = (T + 3) / 4 * 4;
    ScratchLayout L{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    // zero-initialised region first (state), then write-before-read buffers
    L.zero_begin = off;
    L.act = take((size_t)2 * ACT_ROWS * BP * 4);
    const int BPT = BP < 32 ? 32 : BP;                                    // batch tile of the tensor-core path
    L.act16 = take((size_t)2 * ACT_ROWS * 2 * BPT * 2);                   // hi/lo bf16 operand tiles (tensor-core path)
    L.attn_prev = take((size_t)2 * B * Tp * 4);
    L.attn_cum = take((size_t)2 * B * Tp * 4);
    L.ctrl = take((size_t)CTRL_WORDS * 4);
    L.zero_end = off;
    L.pqpart = take((size_t)B * GEMM_CTAS * ATT_DIM * 4);
    L.melpart = take((size_t)B * GEMM_CTAS * PROJ_PAD * 4);
    L.ctx_bj = take((size_t)B * EMB * 4);
    L.wtmp = take((size_t)B * Tp * 4);
    L.apart = take((size_t)B * MAX_NS * APART * 4);
    L.zpart = take((size_t)GEMM_CTAS * GS * ROWS_PER_CTA * BPT * 4);
    L.res = take(decoder_res_scratch_bytes(B, T));                        // exchange ring of the resident kernel (decoder_res.cu)
    L.total = off;
    return L;
}

}  // namespace

bool decoder_uses_tensor_cores(const b200taco_handle* h, int B) {
    // tcgen05 only where batch x 4H is a real GEMM (north_star c); vectorised SIMT contraction otherwise
    // (measured, profiles/r1_v8_decoder_phase_times.jsonl: the tensor-core kernel at its minimum batch tile of 32 is
    // faster than the SIMT kernel from B = TC_MIN_BATCH up; TACO_TC_MIN_B overrides for measurements)
    if (h->decoder_path == 1) return false;
    static const int min_b = [] { const char* e = getenv("TACO_TC_MIN_B"); return e ? atoi(e) : TC_MIN_BATCH; }();
    return h->weight_dtype == 0 && B >= min_b;
}

// batch tile (columns of the activation operand) of the kernel that will run this batch
static int batch_tile(const b200taco_handle* h, int B) {
    const int bp = decoder_batch_pad(B);
    return decoder_uses_tensor_cores(h, B) && bp < 32 ? 32 : bp;
}

int decoder_pack_p_all(const b200taco_handle* h, const float* Y, void* P, int B, int S, cudaStream_t stream) {
    const int BP = batch_tile(h, B);
    if (decoder_uses_tensor_cores(h, B))
        p_all_pack_tc_kernel<<<2048, 256, 0, stream>>>(Y, static_cast<__nv_bfloat16*>(P), B, BP, S);
    else
        p_all_pack_simt_kernel<<<2048, 256, 0, stream>>>(Y, static_cast<float*>(P), B, BP, S);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

int decoder_batch_pad(int B) {
    int bp = 1;
    while (bp < B) bp <<= 1;
    return bp;
}

size_t decoder_scratch_bytes(int B, int T) { return scratch_layout(B, T).total; }

int decoder_run(const b200taco_handle* h, const DecoderArgs& a, cudaStream_t stream) {
    TACO_REQUIRE(a.B >= 1 && a.B <= MAX_BATCH, B200TACO_ERR_UNSUPPORTED, "batch %d outside [1, %d]", a.B, MAX_BATCH);
    TACO_REQUIRE(a.T >= 1 && a.T <= MAX_TEXT_LEN, B200TACO_ERR_UNSUPPORTED, "text length %d outside [1, %d]", a.T,
                 MAX_TEXT_LEN);
    TACO_REQUIRE(a.max_steps >= 1, B200TACO_ERR_INVALID, "max_steps must be >= 1");
    if (a.n_utts > 0) {
        TACO_REQUIRE(a.mode == 0 && decoder_uses_tensor_cores(h, a.B), B200TACO_ERR_UNSUPPORTED,
                     "the utterance queue runs on the tensor-core decoder kernel only (free-running, bf16 LSTM weights, "
                     ">= %d rows; got %d rows)", TC_MIN_BATCH, a.B);
        TACO_REQUIRE(a.n_utts % a.B == 0, B200TACO_ERR_INVALID, "utterance count %d is not a multiple of the %d decoder rows",
                     a.n_utts, a.B);
    }
    TACO_REQUIRE(h->num_sms >= MIN_CTAS, B200TACO_ERR_UNSUPPORTED,
                 "the persistent decoder needs >= %d SMs (device has %d)", MIN_CTAS, h->num_sms);
    const ScratchLayout L = scratch_layout(a.B, a.T);
    TACO_REQUIRE(a.scratch_bytes >= L.total, B200TACO_ERR_STATE, "decoder scratch too small: %zu < %zu",
                 a.scratch_bytes, L.total);
    if (decoder_res_supported(h, a)) return decoder_res_run(h, a, static_cast<uint8_t*>(a.scratch) + L.res, stream);
    int ncta = h->num_sms;
    const int BP = batch_tile(h, a.B);
    // tensor-core path: split-K groups as thread-block clusters (DSMEM exchange) when enough clusters are
    // co-resident to hold the 128 GEMM CTAs plus at least one cluster of spare CTAs; TACO_TC_CLUSTERS=0 disables
    bool clusters = false;
    if (decoder_uses_tensor_cores(h, a.B)) {
        static const bool allow = [] { const char* e = getenv("TACO_TC_CLUSTERS"); return !(e && e[0] == '0'); }();
        const int cl = allow ? (BP == 32 ? tc_cluster_ctas<32>(h->num_sms) : tc_cluster_ctas<64>(h->num_sms)) : 0;
        if (cl >= GEMM_CTAS) {
            clusters = true;
            ncta = cl;
        }
    }
    uint8_t* base = static_cast<uint8_t*>(a.scratch);
    TACO_CUDA(cudaMemsetAsync(base + L.zero_begin, 0, L.zero_end - L.zero_begin, stream));

    KParams kp{};
    kp.B = a.B; kp.T = a.T; kp.Tp = (a.T + 3) / 4 * 4;
    kp.n_utts = a.n_utts;
    int ns = ncta / a.B;
    ns = ns < 1 ? 1 : (ns > MAX_NS ? MAX_NS : ns);
    static const int ns_max = [] { const char* e = getenv("TACO_NS_MAX"); return e ? atoi(e) : MAX_NS; }();   // measurement knob
    ns = ns > ns_max ? ns_max : ns;
    // hoisted SIMT kernel (batch tiles 1-2): the work items are dealt to the CTAs beyond the 128 GEMM CTAs first, so one round of
    // items on those CTAs beats more, shorter slices (measured B=2, T=200: 16 slices 13.5 us, 8 slices 11.5 us; the phase costs
    // ~7.9 us + 0.16 us per position of a slice)
    if (!decoder_uses_tensor_cores(h, a.B) && BP <= 2 && ncta > GEMM_CTAS) {
        const int spare = (ncta - GEMM_CTAS) / a.B;
        if (spare >= 4 && spare < ns) ns = spare;
    }
    // keep slices at least 8 positions long
    while (ns > 1 && (a.T + ns - 1) / ns < 8) --ns;
    if (clusters) {
        // the slices of an utterance must sit inside one cluster (CTA = work item): 1, 2 or 4 slices
        ns = ns >= 4 ? 4 : (ns >= 2 ? 2 : 1);
    }
    kp.NS = ns;
    kp.SL = (a.T + ns - 1) / ns;
    int nq = ncta / a.B;
    kp.NQ = nq >= 4 ? 4 : (nq >= 2 ? 2 : 1);
    kp.max_steps = a.max_steps; kp.mode = a.mode; kp.stop_mode = a.stop_mode; kp.thr = a.gate_threshold;
    kp.w_attn = h->w_attn; kp.w_dec = h->w_dec;
    kp.bias_attn = h->bias_attn; kp.bias_dec = h->bias_dec; kp.wq_slices = h->wq_slices;
    kp.wproj_slices = h->wproj_slices; kp.wproj_ctxT = h->wproj_ctxT; kp.proj_bias = h->proj_bias;
    kp.prenet1T = h->prenet1T; kp.prenet2T = h->prenet2T; kp.loc_conv = h->loc_conv; kp.loc_denseT = h->loc_denseT;
    kp.attn_v = h->attn_v;
    kp.memory = a.memory; kp.pm = a.pm; kp.p_all = a.p_all; kp.text_lengths = a.text_lengths;
    kp.mel = a.mel; kp.gate = a.gate; kp.align = a.align; kp.lengths = a.lengths; kp.steps_run = a.steps_run;
    kp.sv_ga = a.sv_ga; kp.sv_ca = a.sv_ca; kp.sv_ah = a.sv_ah; kp.sv_gd = a.sv_gd; kp.sv_cd = a.sv_cd; kp.sv_dh = a.sv_dh;
    kp.sv_ctx = a.sv_ctx; kp.mask_a = a.mask_a; kp.mask_d = a.mask_d; kp.drop_scale = a.drop_scale;
    kp.act = reinterpret_cast<float*>(base + L.act);
    kp.act16 = reinterpret_cast<__nv_bfloat16*>(base + L.act16);
    kp.w_attn_tc = h->w_attn_tc; kp.w_dec_tc = h->w_dec_tc;
    kp.zpart = reinterpret_cast<float*>(base + L.zpart);
    kp.pqpart = reinterpret_cast<float*>(base + L.pqpart);
    kp.melpart = reinterpret_cast<float*>(base + L.melpart);
    kp.ctx_bj = reinterpret_cast<float*>(base + L.ctx_bj);
    kp.attn_prev = reinterpret_cast<float*>(base + L.attn_prev);
    kp.attn_cum = reinterpret_cast<float*>(base + L.attn_cum);
    kp.wtmp = reinterpret_cast<float*>(base + L.wtmp);
    kp.apart = reinterpret_cast<float*>(base + L.apart);
    kp.ctrl = reinterpret_cast<uint32_t*>(base + L.ctrl);
    const_cast<b200taco_handle*>(h)->last_profile = reinterpret_cast<const unsigned long long*>(kp.ctrl + CTRL_PROF);
    if (decoder_uses_tensor_cores(h, a.B)) {
        TACO_REQUIRE(h->w_attn_tc && h->w_dec_tc, B200TACO_ERR_STATE, "tensor-core weights not packed");
        if (BP == 32) return launch_tc<32>(kp, ncta, clusters, stream);
        return launch_tc<64>(kp, ncta, clusters, stream);
    }
    if (h->weight_dtype == 1) return dispatch_bp<float>(BP, kp, ncta, stream);
    return dispatch_bp<__nv_bfloat16>(BP, kp, ncta, stream);
}

}  // namespace b200taco
