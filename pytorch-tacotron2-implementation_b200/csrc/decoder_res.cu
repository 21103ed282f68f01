// Weight-resident, barrier-free persistent decoder for one or two utterances (free-running inference,
// Decoder.inference / Decoder._decode_step, reference src/model.py:246-333).
//
// The three persistent kernels in decoder.cu stream the 36.4 MB of bf16 LSTM weights from L2 every decode step and
// separate the four phases of a step with all-to-all grid barriers (1.25 us each plus skew and the dependent L2 round
// trip behind it): at B=1 the step is latency, 20.7 us against a 2.2 us L2-bandwidth floor.  This kernel removes both:
//
//  * 128 "gate" CTAs own 8 hidden units (32 gate rows) of BOTH LSTM cells and keep their 272 KB weight slice ON CHIP for
//    the whole launch: the rows multiplied on the critical path (attention LSTM x prenet, decoder LSTM x context) and the
//    attention LSTM x context rows in registers (40 packed bf16x2 registers per thread), the three 1024-wide recurrent /
//    attention-hidden ranges in shared memory (192 KB).  No weight byte moves during the frame loop.
//  * the remaining CTAs (20 on a B200) split into "attention" CTAs and 8 "frame" CTAs.  The attention CTAs of an utterance
//    split the ATTENTION DIMENSIONS (and the context dimensions), not the encoder positions: each owns rows of the query
//    projection, the matching columns of processed memory / location features for ALL positions, and a column slice of the
//    encoder memory.  It needs nobody else's processed query (no exchange), publishes its partial energies e_j[t], sums
//    everybody's partials, runs the (redundant, deterministic) softmax over all positions and finishes its own context
//    dimensions: two exchanges per attention instead of four, and the attention weights never leave the CTA.  Location
//    convolution and location dense layer (both bias free, src/model.py:96-101) are applied as ONE composed filter
//    G[a][c][k] = sum_f W_dense[a][f] W_conv[f][c][k], so a CTA only computes its own attention dimensions.
//    The frame CTAs own rows of the mel / gate projection and of the prenet (weights in registers / shared memory).
//  * every value that crosses CTAs travels as a 64-bit word {fp32 value, step tag} written with one relaxed store and
//    polled by its consumers (the "LL" protocol of collective libraries): data and flag arrive together, so an exchange
//    costs one L2 write + one L2 read instead of fence + atomic + spin + dependent load, there is no grid barrier, and a
//    consumer only waits for the producers it actually reads.  Tags are step + 1 in a 4-deep ring of buffers.
//  * everything that does not depend on the newest value is computed while waiting for it: the recurrent contractions
//    W[ah].ah, W[dh].dh, W_a[ctx].ctx accumulate in registers in the background, the location convolution and the
//    location/processed-memory part of the energies are ready before the attention hidden state arrives.
//
// Arithmetic per step is the same fp32 math as decoder_simt_kernel (bf16 weights exact, fp32 state and accumulation);
// only summation orders differ.  Stop rule, F6 discarded first step and output layouts are those of decoder.cu.
#include <cstdlib>

#include "handle.h"

namespace b200taco {

namespace {

constexpr int RT = 512;                       // threads per CTA
constexpr int RWARPS = RT / 32;
constexpr int DEPTH = 4;                      // exchange ring depth (steps)
constexpr int NREP = 4;                       // replicas of every broadcast vector: consumer CTA c polls replica c % NREP
constexpr int NFR = 8;                        // frame CTAs
constexpr int MPR = (PROJ_OUT + NFR - 1) / NFR;   // mel / gate projection rows per frame CTA (one warp per row, weights in registers)
constexpr int MELW = 12;                      // exchange words per frame CTA (MPR padded to whole 32-byte sectors)
constexpr int P2R = PRENET / NFR;             // prenet layer-2 rows per frame CTA (one half-warp per row)
constexpr int MAX_NSL = 16;                   // attention CTAs per utterance
constexpr int HALO = (LOC_K - 1) / 2;
constexpr int W0 = 16;                        // zeros in front of the attention-weight windows
constexpr int TB = 4;                         // positions per work item of the location filter
constexpr int NTAP = 2 * LOC_K;
constexpr uint32_t STOP_BIT = 0x80000000u;
constexpr uint32_t SPIN_LIMIT = 1u << 25;     // polls of one word before the kernel traps (~10 s)
constexpr int PROF_WORDS = 24;
constexpr int TRACE_EV = 64, TRACE_WORDS = 160 * TRACE_EV + 2 * 512;   // per-CTA event times of one decode step (measurement aid)

static_assert(NREP == 4, "the four gate lanes of a hidden unit publish the four replicas");
static_assert(RT == 2 * PRENET && MPR <= RWARPS && MPR <= MELW && P2R == 2 * RWARPS && MELW % 4 == 0, "thread maps");

struct ASmem {                                // shared-memory layout of an attention CTA (floats), computed by the host
    int G, red, pq, v, winp, winc, pm, pre, ew, mx, cpart, spart, mem, SP, NTG, PT, total;
};

struct RParams {
    int B, T, Tp, Te, NATT, NSL, DA, DC, max_steps, stop_mode, mem_resident, bg_sleep, bg_backoff;
    float thr;
    const uint32_t *w_attn, *w_dec;           // bf16 pairs: [128 cta][K][16 words], word w = rows (2w, 2w+1) of the CTA
    const float *bias_attn, *bias_dec, *wq_slices, *wproj_slices, *wproj_ctxT, *proj_bias, *prenet1T, *prenet2T, *loc_conv,
        *loc_denseT, *attn_v;
    const float *memory, *pm;
    const int64_t* text_lengths;
    float *mel, *gate, *align;
    int32_t *lengths, *steps_run;
    uint64_t* xch;                            // exchange ring
    unsigned long long* prof;
    unsigned long long* trace;               // [cta][TRACE_EV] globaltimer of step trace_step, or nullptr
    int trace_step;
    ASmem as;
};

// word offsets of the exchanged vectors inside one ring slot
struct XL {
    int p, ah, ctx, dh, ep, mel, words;
};
__host__ __device__ inline XL xlayout(int BP, int Te, int NSL) {
    XL x;
    int o = 0;
    x.p = o;    o += NREP * BP * PRENET;      // prenet output            frame -> gate
    x.ah = o;   o += NREP * BP * RNN;         // attention hidden         gate -> gate, attention
    x.ctx = o;  o += NREP * BP * EMB;         // attention context        attention -> gate, frame
    x.dh = o;   o += NREP * BP * RNN;         // decoder hidden           gate -> gate, frame
    x.ep = o;   o += BP * NSL * Te;           // partial energies         attention -> attention (same utterance); rows of Te = 16 n
                                              // words: every 128-byte line is written by ONE store instruction
    x.mel = o;  o += BP * NFR * MELW;         // mel frame + gate logit   frame -> frame, attention (gate)
    x.words = (o + 15) & ~15;
    return x;
}

// ---- tagged 64-bit words --------------------------------------------------------------------------------------------
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
__device__ __noinline__ void res_spin_fail() { __trap(); }
__device__ __forceinline__ bool tag_ok(uint64_t w, uint32_t tag) { return (((uint32_t)(w >> 32) ^ tag) & ~STOP_BIT) == 0u; }

// one word; every lane of the warp calls it (with its own address) and the warp leaves together
__device__ __forceinline__ float ll_wait1(const uint64_t* p, uint32_t tag) {
    uint64_t w = ll_ld1(p);
    uint32_t n = 0;
    while (!__all_sync(0xffffffffu, tag_ok(w, tag))) {
        w = ll_ld1(p);
        if (++n > SPIN_LIMIT) res_spin_fail();
    }
    return __uint_as_float((uint32_t)w);
}

// n words (even, 16-byte aligned) -> n floats in shared memory; returns the OR of the words' stop bits.  backoff_ns > 0:
// a gather nobody is waiting for (the recurrent contractions of the NEXT cell evaluation) sleeps between polls, so that its
// polling traffic does not sit in front of the stores and polls of the CTAs on the critical path
__device__ __forceinline__ uint32_t ll_gather(const uint64_t* src, float* dst, int n, uint32_t tag, int tid, unsigned backoff_ns = 0) {
    uint32_t stop = 0u;
    for (int i = tid; i < n / 2; i += RT) {              // n / 2 is a multiple of 32: whole warps
        uint64_t a, b;
        ll_ld2(src + 2 * i, a, b);
        uint32_t it = 0;
        // the warp leaves the loop as ONE: a lane-wise exit splits the warp, and the shuffles / barriers behind a split warp
        // were measured to cost ~2 us (divergent-shuffle slow path)
        while (!__all_sync(0xffffffffu, tag_ok(a, tag) && tag_ok(b, tag))) {
            if (backoff_ns) __nanosleep(backoff_ns);
            ll_ld2(src + 2 * i, a, b);
            if (++it > SPIN_LIMIT) res_spin_fail();
        }
        *reinterpret_cast<float2*>(dst + 2 * i) = make_float2(__uint_as_float((uint32_t)a), __uint_as_float((uint32_t)b));
        stop |= (uint32_t)((a | b) >> 32) & STOP_BIT;
    }
    return stop;
}

// Every broadcast vector exists in NREP copies at different addresses, and the L2 slice (and die) that owns an address is a
// hash of it: a consumer on the far die sees a new value ~0.35 us later than one next to the owning slice (measured: the gate
// CTAs of one die receive the prenet output that much later than the other die's).  Each thread therefore times a few loads
// of ITS words in every copy once, per ring slot, and polls the copy that answers fastest: 2 bits per slot.
__device__ __forceinline__ uint32_t pick_replicas(const uint64_t* xch, size_t slot_words, size_t vec_off, size_t rep_stride) {
    uint32_t choice = 0u;
    for (int slot = 0; slot < DEPTH; ++slot) {
        long long best_t = 1ll << 40;
        uint32_t best = 0u;
        for (int r = 0; r < NREP; ++r) {
            const uint64_t* q = xch + (size_t)slot * slot_words + vec_off + (size_t)r * rep_stride;
            long long m = 1ll << 40;
            for (int k = 0; k < 4; ++k) {
                const long long t0 = clock64();
                uint64_t z;
                do {                                                  // the branch on the loaded value orders the clock reads
                    const uint64_t w = ll_ld1(q);
                    asm volatile("and.b64 %0, %1, 0;" : "=l"(z) : "l"(w));
                } while (z != 0);
                const long long dt = clock64() - t0;
                m = dt < m ? dt : m;
            }
            if (m < best_t) { best_t = m; best = (uint32_t)r; }
        }
        choice |= best << (2 * slot);
    }
    return choice;
}
__device__ __forceinline__ int replica_of(uint32_t choice, int s) { return (int)((choice >> (2 * (s & (DEPTH - 1)))) & 3u); }

// ---- gate-row contractions.  Warp w owns gate rows 2w, 2w+1 of the CTA (packed in one 32-bit word per k); lane l owns
// k = 128 j + 4 l + i, so that both the weight vector and the activation vector of a lane are conflict-free 16-byte
// shared-memory accesses.
template <int BP>
__device__ __forceinline__ void fma_word(uint32_t w, const float (&x)[BP], float (&acc)[2][BP]) {
#pragma unroll
    for (int b = 0; b < BP; ++b) {
        acc[0][b] = fmaf(bf16lo(w), x[b], acc[0][b]);
        acc[1][b] = fmaf(bf16hi(w), x[b], acc[1][b]);
    }
}
template <int BP, int NV>
__device__ __forceinline__ void gemv_reg(const uint32_t (&w)[NV * 4], const float* x, int ldx, float (&acc)[2][BP], int lane) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        float4 x4[BP];
#pragma unroll
        for (int b = 0; b < BP; ++b) x4[b] = reinterpret_cast<const float4*>(x + b * ldx)[j * 32 + lane];
        float xv[BP];
#pragma unroll
        for (int b = 0; b < BP; ++b) xv[b] = x4[b].x;
        fma_word<BP>(w[j * 4 + 0], xv, acc);
#pragma unroll
        for (int b = 0; b < BP; ++b) xv[b] = x4[b].y;
        fma_word<BP>(w[j * 4 + 1], xv, acc);
#pragma unroll
        for (int b = 0; b < BP; ++b) xv[b] = x4[b].z;
        fma_word<BP>(w[j * 4 + 2], xv, acc);
#pragma unroll
        for (int b = 0; b < BP; ++b) xv[b] = x4[b].w;
        fma_word<BP>(w[j * 4 + 3], xv, acc);
    }
}
template <int BP>
__device__ __forceinline__ void gemv_smem(const uint32_t* wrow, const float* x, float (&acc)[2][BP], int lane) {
    const uint4* wv = reinterpret_cast<const uint4*>(wrow);
#pragma unroll
    for (int j = 0; j < RNN / 128; ++j) {
        const uint4 w4 = wv[j * 32 + lane];
        float4 x4[BP];
#pragma unroll
        for (int b = 0; b < BP; ++b) x4[b] = reinterpret_cast<const float4*>(x + b * RNN)[j * 32 + lane];
        float xv[BP];
#pragma unroll
        for (int b = 0; b < BP; ++b) xv[b] = x4[b].x;
        fma_word<BP>(w4.x, xv, acc);
#pragma unroll
        for (int b = 0; b < BP; ++b) xv[b] = x4[b].y;
        fma_word<BP>(w4.y, xv, acc);
#pragma unroll
        for (int b = 0; b < BP; ++b) xv[b] = x4[b].z;
        fma_word<BP>(w4.z, xv, acc);
#pragma unroll
        for (int b = 0; b < BP; ++b) xv[b] = x4[b].w;
        fma_word<BP>(w4.w, xv, acc);
    }
}
// warp-reduce the lane partials of this warp's two rows, leave them in z[b][32] and restart the accumulators
template <int BP>
__device__ __forceinline__ void reduce_rows(float (&acc)[2][BP], float* z, int warp, int lane) {
#pragma unroll
    for (int b = 0; b < BP; ++b) {
        const float r0 = warp_sum(acc[0][b]), r1 = warp_sum(acc[1][b]);
        if (lane == 0) *reinterpret_cast<float2*>(z + b * ROWS_PER_CTA + 2 * warp) = make_float2(r0, r1);
        acc[0][b] = 0.f;
        acc[1][b] = 0.f;
    }
}
// LSTM cell of utterance b = warp (one lane per gate row, nn.LSTMCell semantics, src/model.py:257-262 / :266-270);
// the four gate lanes of a unit publish the new hidden value to the four replicas
__device__ __forceinline__ float lstm_cell(const float* z, const float* bias, float& c, uint64_t* xdst, int ld_rep, int unit0,
                                           uint32_t tag, int lane) {
    const int gate = lane & 3, u = lane >> 2;
    const float pre = z[lane] + bias[lane];
    const float act = gate == 2 ? tanhf(pre) : sigmoid_acc(pre);
    const int l0 = lane & ~3;
    const float gi = __shfl_sync(0xffffffffu, act, l0), gf = __shfl_sync(0xffffffffu, act, l0 + 1),
                gg = __shfl_sync(0xffffffffu, act, l0 + 2), go = __shfl_sync(0xffffffffu, act, l0 + 3);
    c = __fmaf_rn(gf, c, __fmul_rn(gi, gg));
    const float h = go * tanhf(c);
    ll_st(xdst + (size_t)gate * ld_rep + unit0 + u, h, tag);
    return h;
}

#define RES_EVW(i)                                                                                 \
    do {                                                                                           \
        if (p.trace != nullptr && (threadIdx.x & 31) == 0 && s == p.trace_step)                    \
            p.trace[(size_t)blockIdx.x * TRACE_EV + (i) + (threadIdx.x >> 5)] = global_timer_ns(); \
    } while (0)
#define RES_EVT(slot)                                                                              \
    do {                                                                                           \
        if (p.trace != nullptr && s == p.trace_step)                                               \
            p.trace[160 * TRACE_EV + (slot) * 512 + threadIdx.x] = global_timer_ns();              \
    } while (0)
#define RES_EV(i)                                                                                  \
    do {                                                                                           \
        if (p.trace != nullptr && threadIdx.x == 0 && s == p.trace_step)                           \
            p.trace[(size_t)blockIdx.x * TRACE_EV + (i)] = global_timer_ns();                      \
    } while (0)

// =====================================================================================================================
// stop rule, evaluated identically by every frame / attention CTA from the published gate logits: batch wide
// (src/model.py:327) or per utterance (train.py:71-79 with the steps > 1 guard)
template <int BP>
__device__ __forceinline__ bool stop_rule(const RParams& p, const float (&gate)[BP], int out_idx, bool (&done)[BP], int (&ulen)[BP]) {
    if (out_idx < 1) return false;
    bool any = false, all_done = true;
#pragma unroll
    for (int q = 0; q < BP; ++q) {
        const bool fired = sigmoid_acc(gate[q]) > p.thr;
        any |= fired;
        if (fired && !done[q]) { done[q] = true; ulen[q] = out_idx + 1; }
        all_done &= done[q];
    }
    return p.stop_mode == B200TACO_STOP_BATCH_WIDE ? any : all_done;
}

// =====================================================================================================================
// gate CTAs
template <int BP>
__device__ __forceinline__ void gate_role(const RParams& p, uint8_t* smem, const XL& xl) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, cta = blockIdx.x;
    uint32_t* sW = reinterpret_cast<uint32_t*>(smem);                 // [3 ranges][16 warps][1024 k]
    float* xah = reinterpret_cast<float*>(smem + 3 * RWARPS * RNN * 4);
    float* xdh = xah + BP * RNN;
    float* xctx = xdh + BP * RNN;
    float* xp = xctx + BP * EMB;
    float* z_s = xp + BP * PRENET;
    float* bias_s = z_s + BP * ROWS_PER_CTA;

    const uint32_t* wa = p.w_attn + (size_t)cta * K_ATTN * (ROWS_PER_CTA / 2);      // K order [prenet | ctx | ah]
    const uint32_t* wd = p.w_dec + (size_t)cta * K_DEC * (ROWS_PER_CTA / 2);        // K order [ctx | ah | dh]
    for (int i = tid; i < RWARPS * RNN; i += RT) {
        const int w = i & 15, k = i >> 4;
        sW[(0 * RWARPS + w) * RNN + k] = __ldg(wa + (size_t)(PRENET + EMB + k) * 16 + w);   // attention LSTM x ah
        sW[(1 * RWARPS + w) * RNN + k] = __ldg(wd + (size_t)(EMB + k) * 16 + w);            // decoder LSTM x ah
        sW[(2 * RWARPS + w) * RNN + k] = __ldg(wd + (size_t)(EMB + RNN + k) * 16 + w);      // decoder LSTM x dh
    }
    uint32_t wa_p[8], wa_c[16], wd_c[16];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) wa_p[j * 4 + i] = __ldg(wa + (size_t)(j * 128 + lane * 4 + i) * 16 + warp);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            wa_c[j * 4 + i] = __ldg(wa + (size_t)(PRENET + j * 128 + lane * 4 + i) * 16 + warp);
            wd_c[j * 4 + i] = __ldg(wd + (size_t)(j * 128 + lane * 4 + i) * 16 + warp);
        }
    for (int i = tid; i < ROWS_PER_CTA; i += RT) {
        bias_s[i] = p.bias_attn[cta * ROWS_PER_CTA + i];
        bias_s[ROWS_PER_CTA + i] = p.bias_dec[cta * ROWS_PER_CTA + i];
    }
    __syncthreads();

    // the two vectors this CTA waits for on the critical path: thread i gathers words 2 i, 2 i + 1
    const uint32_t rep_p = tid < BP * PRENET / 2 ? pick_replicas(p.xch, xl.words, xl.p + 2 * tid, BP * PRENET) : 0u;
    const uint32_t rep_c = tid < BP * EMB / 2 ? pick_replicas(p.xch, xl.words, xl.ctx + 2 * tid, BP * EMB) : 0u;
    float acc_a[2][BP], acc_d[2][BP];
#pragma unroll
    for (int b = 0; b < BP; ++b) { acc_a[0][b] = acc_a[1][b] = acc_d[0][b] = acc_d[1][b] = 0.f; }
    float c_a = 0.f, c_d = 0.f;
    const int rep = cta % NREP, unit0 = cta * UNITS_PER_CTA;
    const int total_steps = p.max_steps + 1;
    // phase timers of CTA 0 (b200taco_decoder_profile): accumulated in shared memory by thread 0, so that they cost no registers
    __shared__ unsigned long long t_acc[PROF_WORDS];
    const bool timing = cta == 0 && tid == 0;
    if (timing) {
        for (int i = 0; i < PROF_WORDS; ++i) t_acc[i] = 0ull;
        t_acc[PROF_WORDS - 1] = global_timer_ns();
    }
    auto stamp = [&](int ph) {
        if (timing) {
            const unsigned long long now = global_timer_ns();
            t_acc[ph] += now - t_acc[PROF_WORDS - 1];
            t_acc[PROF_WORDS - 1] = now;
        }
    };
    int steps_done = 0;
    for (int s = 0; s < total_steps; ++s) {
        const uint32_t tag = (uint32_t)s + 1u;
        uint64_t* X = p.xch + (size_t)(s & (DEPTH - 1)) * xl.words;
        // ---- attention LSTM: only the prenet rows wait for this step's input
        const uint32_t stop = ll_gather(X + xl.p + replica_of(rep_p, s) * BP * PRENET, xp, BP * PRENET, tag, tid);
        if (__syncthreads_or((int)stop)) break;
        if (s > 0) stamp(3);
        RES_EV(0);
        if (p.trace != nullptr && tid == 0 && s == p.trace_step + 1) p.trace[(size_t)blockIdx.x * TRACE_EV + 6] = global_timer_ns();
        gemv_reg<BP, 2>(wa_p, xp, PRENET, acc_a, lane);
        reduce_rows<BP>(acc_a, z_s, warp, lane);
        __syncthreads();
        if (warp < BP) lstm_cell(z_s + warp * ROWS_PER_CTA, bias_s, c_a, X + xl.ah + warp * RNN, BP * RNN, unit0, tag, lane);
        stamp(0);
        RES_EV(1);
        // ---- background: both cells' rows of the new attention hidden state
        if (p.bg_sleep > 0) __nanosleep(p.bg_sleep);
        ll_gather(X + xl.ah + rep * BP * RNN, xah, BP * RNN, tag, tid, p.bg_backoff);
        __syncthreads();
        gemv_smem<BP>(sW + (1 * RWARPS + warp) * RNN, xah, acc_d, lane);
        gemv_smem<BP>(sW + (0 * RWARPS + warp) * RNN, xah, acc_a, lane);
        RES_EV(5);
        // ---- decoder LSTM: only the context rows wait for the attention
        ll_gather(X + xl.ctx + replica_of(rep_c, s) * BP * EMB, xctx, BP * EMB, tag, tid);
        __syncthreads();
        stamp(1);
        RES_EV(2);
        gemv_reg<BP, 4>(wd_c, xctx, EMB, acc_d, lane);
        reduce_rows<BP>(acc_d, z_s, warp, lane);
        __syncthreads();
        if (warp < BP)
            lstm_cell(z_s + warp * ROWS_PER_CTA, bias_s + ROWS_PER_CTA, c_d, X + xl.dh + warp * RNN, BP * RNN, unit0, tag, lane);
        stamp(2);
        RES_EV(3);
        // ---- background: attention LSTM x context, decoder LSTM x new decoder hidden (both for the next step)
        gemv_reg<BP, 4>(wa_c, xctx, EMB, acc_a, lane);
        if (p.bg_sleep > 0) __nanosleep(p.bg_sleep);
        ll_gather(X + xl.dh + rep * BP * RNN, xdh, BP * RNN, tag, tid, p.bg_backoff);
        __syncthreads();
        gemv_smem<BP>(sW + (2 * RWARPS + warp) * RNN, xdh, acc_d, lane);
        RES_EV(4);
        steps_done = s + 1;
    }
    if (timing) {
        for (int i = 0; i < 4; ++i) p.prof[i] = t_acc[i];
        p.prof[4] = (unsigned long long)steps_done;
    }
}

// ---- reduction of N values per lane over the 32 lanes of a warp with N - 1 (+ log2(32 / N)) shuffles instead of 5 N:
// every halving step a lane keeps one half of its values and receives the partner's partials of that half.  On return
// the lane holds the warp total of value index  lane >> log2(32 / N)  (fixed summation order).
template <int N>
__device__ __forceinline__ float multi_reduce(float (&v)[N], int lane) {
    static_assert(N == 8 || N == 16 || N == 32, "multi_reduce width");
    constexpr int STEPS = N == 32 ? 5 : (N == 16 ? 4 : 3);
#pragma unroll
    for (int st = 0; st < STEPS; ++st) {
        const int cnt = N >> st, off = 16 >> st;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < cnt / 2; ++i) {
            const float send = up ? v[i] : v[i + cnt / 2];
            const float keep = up ? v[i + cnt / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float r = v[0];
#pragma unroll
    for (int off = 16 >> STEPS; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

// two tagged words (16-byte aligned pair) straight into registers
__device__ __forceinline__ uint32_t ll_wait2(const uint64_t* p, uint32_t tag, float& x0, float& x1) {
    uint64_t a, b;
    ll_ld2(p, a, b);
    uint32_t it = 0;
    while (!__all_sync(0xffffffffu, tag_ok(a, tag) && tag_ok(b, tag))) {
        ll_ld2(p, a, b);
        if (++it > SPIN_LIMIT) res_spin_fail();
    }
    x0 = __uint_as_float((uint32_t)a);
    x1 = __uint_as_float((uint32_t)b);
    return (uint32_t)((a | b) >> 32) & STOP_BIT;
}

// =====================================================================================================================
// attention CTAs: location-sensitive attention (src/model.py:117-148) of ONE utterance, split over its NSL CTAs by
// attention dimension (DA <= RW dims each: query projection rows, processed memory, composed location filter, energies)
// and by context dimension (DC dims each).  Every contraction is split over K so that a thread multiplies exactly the
// words it polled (no staging of the incoming vector, no shared-memory traffic for weights: they sit in registers).
// T <= 256 (one lane pair per encoder position).  The loop body must stay well inside the 32 KB instruction cache: a body
// of 46 KB was measured to run 3x slower than its instruction count (every line refetched from L2 every decode step), so
// layouts come precomputed from the host and the memory columns have ONE home per instantiation (MM).
constexpr int RES_MAX_T = 256;
template <int RW>
__host__ inline ASmem a_smem(int T, int DA, int DC, bool mem_in_smem) {
    ASmem a;
    const int Tp = (T + 3) & ~3;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
    a.SP = ((Tp + TB + 15) / 32) * 32 + 16;           // row stride of pm / pre: 16 mod 32 (two rows never share a bank)
    a.NTG = RT / DC;                                  // position groups of the context contraction
    a.PT = (((T + a.NTG - 1) / a.NTG) + 3) & ~3;      // positions per group
    a.G = take(NTAP * RW);
    a.red = take(RWARPS * RW);
    a.pq = take(RW);
    a.v = take(RW);
    a.winp = take(Tp + 64);
    a.winc = take(Tp + 64);
    a.pm = take(DA * a.SP);
    a.pre = take(DA * a.SP);
    a.ew = take((a.NTG * a.PT > RES_MAX_T ? a.NTG * a.PT : RES_MAX_T) + 16);   // read as whole float4s; zero beyond T
    a.mx = take(RWARPS);
    a.cpart = take(a.NTG * DC);
    a.spart = take(RWARPS + 4);
    a.mem = take(mem_in_smem ? a.NTG * a.PT * DC : 0);
    a.total = o;
    return a;
}

template <int BP, int RW, int MM>
__device__ __forceinline__ void attn_role(const RParams& p, uint8_t* smem, const XL& xl) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j = (int)blockIdx.x - GEMM_CTAS, cta = blockIdx.x;
    const int NSL = p.NSL, T = p.T, DA = p.DA, DC = p.DC;
    const int b = j / NSL, sl = j % NSL;                       // NATT == BP * NSL
    const int a0 = sl * DA, na = max(0, min(ATT_DIM, a0 + DA) - a0);
    const int d0 = sl * DC, nd = max(0, min(EMB, d0 + DC) - d0);
    const int len = p.text_lengths ? (int)p.text_lengths[b] : T;
    const ASmem& L = p.as;
    const int SP = L.SP, NTG = L.NTG, PT = L.PT;
    float* sm = reinterpret_cast<float*>(smem);
    float *G_s = sm + L.G, *red_s = sm + L.red, *pq_s = sm + L.pq, *v_s = sm + L.v, *winp = sm + L.winp, *winc = sm + L.winc,
          *pm_s = sm + L.pm, *pre_s = sm + L.pre, *ew_s = sm + L.ew, *mx_s = sm + L.mx, *cpart_s = sm + L.cpart, *spart_s = sm + L.spart,
          *mem_s = sm + L.mem;

    // ---- one-off.  Query projection: thread owns k = 2 tid, 2 tid + 1 of all own rows (registers)
    float wq0[RW], wq1[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        const int k = 2 * tid;
        wq0[r] = r < na ? __ldg(p.wq_slices + ((size_t)(k >> 3) * ATT_DIM + a0 + r) * 8 + (k & 7)) : 0.f;
        wq1[r] = r < na ? __ldg(p.wq_slices + ((size_t)(k >> 3) * ATT_DIM + a0 + r) * 8 + (k & 7) + 1) : 0.f;
    }
    for (int i = tid; i < RW; i += RT) v_s[i] = i < na ? __ldg(p.attn_v + a0 + i) : 0.f;        // energies: v of the own dims
    // composed location filter G[(c,k)][a] = sum_f W_dense[a][f] W_conv[f][c][k]
    for (int i = tid; i < NTAP * RW; i += RT) {
        const int ck = i / RW, a = i % RW;
        float g = 0.f;
        if (a < na)
            for (int f = 0; f < LOC_F; ++f) g = fmaf(__ldg(p.loc_denseT + f * ATT_DIM + a0 + a), __ldg(p.loc_conv + f * NTAP + ck), g);
        G_s[i] = g;
    }
    for (int i = tid; i < p.Tp + 64; i += RT) { winp[i] = 0.f; winc[i] = 0.f; }
    for (int i = tid; i < DA * SP; i += RT) { pm_s[i] = 0.f; pre_s[i] = 0.f; }
    for (int i = tid; i < (NTG * PT > RES_MAX_T ? NTG * PT : RES_MAX_T) + 16; i += RT) ew_s[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < T * na; i += RT) {
        const int t = i / na, a = i % na;
        const float v = __ldg(p.pm + ((size_t)b * T + t) * ATT_DIM + a0 + a);
        pm_s[a * SP + t] = v;
        pre_s[a * SP + t] = v;                                  // decode step 0: zero attention history
    }
    // context: thread (ctg, cd) owns positions [ctg PT, ctg PT + PT) of memory column d0 + cd, held in registers (MM == 2) or
    // shared memory (MM == 1); the same pair is the (replica, dim) of the thread that publishes a context dim.  (A variant that
    // finished every dim with one warp reduction - no shared-memory partials, no block barrier - measured 1 us SLOWER per step.)
    const int ctg = tid / DC, cd = tid % DC;
    const float* memG = p.memory + (size_t)b * T * EMB + d0;
    constexpr int MEMR = 20;                                   // MM == 2: positions per thread held in registers
    float memr[MM == 2 ? MEMR : 1];
    if (MM == 2) {
#pragma unroll
        for (int i = 0; i < MEMR; ++i) {
            const int t = ctg * PT + i;
            memr[i] = (ctg < NTG && cd < nd && i < PT && t < T) ? __ldg(memG + (size_t)t * EMB + cd) : 0.f;
        }
    } else {
        for (int i = tid; i < NTG * PT * DC; i += RT) {
            const int t = i / DC, d = i % DC;
            mem_s[i] = (t < T && d < nd) ? __ldg(memG + (size_t)t * EMB + d) : 0.f;
        }
    }
    const float v_bias = p.attn_v[ATT_DIM], e_scale = p.attn_v[ATT_DIM + 1];
    __syncthreads();

    const int total_steps = p.max_steps + 1;
    bool done[BP];
    int ulen[BP];
#pragma unroll
    for (int q = 0; q < BP; ++q) { done[q] = false; ulen[q] = 0; }
    const uint32_t rep_ah = pick_replicas(p.xch, xl.words, xl.ah + (size_t)b * RNN + 2 * tid, BP * RNN);
    const int nitems = ((T + TB - 1) / TB) * na;
    const int al0 = sl * ((T + NSL - 1) / NSL), al1 = min(T, al0 + (T + NSL - 1) / NSL);     // alignment columns written by this CTA
    const int et = tid >> 1, eh = lane & 1;                    // energies: lane pair (position, dim / slice parity)
    const uint64_t* Xm_gate = p.xch + xl.mel + (N_MELS / MPR) * MELW + N_MELS % MPR;        // gate logit of utterance 0 in slot 0

#pragma unroll 1
    for (int s = 0; s < total_steps; ++s) {
        const uint32_t tag = (uint32_t)s + 1u;
        uint64_t* X = p.xch + (size_t)(s & (DEPTH - 1)) * xl.words;
        const int out_idx = s - 1;                                   // F6: decode step 0 is discarded

        if (s > 0) {
            // ---- (0) before the attention hidden state arrives: pre[a][t] = pm[a][t] + sum_{c,k} G[a][c][k] win_c[t + k - 15]
            // (location conv + location dense, src/model.py:125-127); a work item = (TB positions, one attention dim)
#pragma unroll 1
            for (int it = tid; it < nitems; it += RT) {
                const int tb = it / na, a = it - tb * na, t0 = tb * TB;
                float acc[TB];
#pragma unroll
                for (int q = 0; q < TB; ++q) acc[q] = 0.f;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {                      // rolled: the loop body is walked once per decode step
                    const float* w = (c ? winc : winp) + W0 - HALO + t0;
                    const float* gp = G_s + c * LOC_K * RW + a;
                    float x[TB + LOC_K - 1];
#pragma unroll
                    for (int q = 0; q < TB - 1; ++q) x[q] = w[q];
#pragma unroll
                    for (int k = 0; k < LOC_K; ++k) {
                        x[k + TB - 1] = w[k + TB - 1];
                        const float g = gp[k * RW];
#pragma unroll
                        for (int q = 0; q < TB; ++q) acc[q] = fmaf(g, x[q + k], acc[q]);
                    }
                }
                const float* pm_r = pm_s + a * SP + t0;
                float* pre_r = pre_s + a * SP + t0;
#pragma unroll
                for (int q = 0; q < TB; ++q) pre_r[q] = pm_r[q] + acc[q];
            }
            __syncthreads();
            RES_EV(0);
            // the previous frame's gate logits: every CTA evaluates the stop rule itself
            float g[BP];
#pragma unroll
            for (int q = 0; q < BP; ++q) g[q] = ll_wait1(Xm_gate + (size_t)((s - 1) & (DEPTH - 1)) * xl.words + q * NFR * MELW, tag - 1u);
            if (stop_rule<BP>(p, g, out_idx - 1, done, ulen)) break;
        }
        // ---- (1) attention hidden state -> own rows of the processed query (src/model.py:123): K split over the threads
        {
            float x0, x1;
            ll_wait2(X + xl.ah + (size_t)(replica_of(rep_ah, s) * BP + b) * RNN + 2 * tid, tag, x0, x1);
            RES_EV(1);
            float v[RW];
#pragma unroll
            for (int r = 0; r < RW; ++r) v[r] = fmaf(wq1[r], x1, wq0[r] * x0);
            const float tot = multi_reduce<RW>(v, lane);
            if ((lane & (32 / RW - 1)) == 0) red_s[warp * RW + lane / (32 / RW)] = tot;
        }
        __syncthreads();
        if (tid < RW) {
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < RWARPS; ++w) acc += red_s[w * RW + tid];
            pq_s[tid] = acc;
        }
        __syncthreads();
        RES_EV(2);
        // ---- (2) partial energies over own dims, all positions (src/model.py:128): two lanes per position
        if (warp * 16 < T) {
            const int tt = et < T ? et : T - 1;
            float x0 = 0.f;
#pragma unroll
            for (int u = 0; u < RW / 2; ++u) {
                const int a = eh + 2 * u;
                x0 = fmaf(v_s[a], tanh_fast(pq_s[a] + pre_s[(a < DA ? a : 0) * SP + tt]), x0);
            }
            x0 += __shfl_xor_sync(0xffffffffu, x0, 1);
            if (eh == 0 && et < T) ll_st(X + xl.ep + (size_t)(b * NSL + sl) * p.Te + et, x0, tag);
        }
        RES_EV(3);
        // ---- (3) everybody's partials -> energies (src/model.py:129-132), summed in slice order by every CTA alike (the
        // lane pair splits the slices even / odd); softmax numerator (src/model.py:139)
        float e_own = -INFINITY;
        if (warp * 16 < T) {
            const uint64_t* Xe = X + xl.ep + (size_t)b * NSL * p.Te + (et < T ? et : T - 1);
            float e;
            uint32_t it = 0;
            bool ok;
            do {                                                 // all loads of a round in flight together
                e = 0.f;
                ok = true;
#pragma unroll
                for (int q = 0; q < MAX_NSL / 2; ++q)
                    if (2 * q + eh < NSL) {
                        const uint64_t w = ll_ld1(Xe + (size_t)(2 * q + eh) * p.Te);
                        ok = ok && tag_ok(w, tag);
                        e += __uint_as_float((uint32_t)w);
                    }
                if (++it > SPIN_LIMIT) res_spin_fail();
            } while (!__all_sync(0xffffffffu, ok));
            e += __shfl_xor_sync(0xffffffffu, e, 1);
            if (et < T) e_own = et >= len ? -1e9f : (e + v_bias) * e_scale;
        }
        {
            const float m = warp_max(e_own);
            if (lane == 0) mx_s[warp] = m;
        }
        __syncthreads();
        RES_EV(4);
        {
            const float4 m0 = *reinterpret_cast<const float4*>(mx_s), m1 = *reinterpret_cast<const float4*>(mx_s + 4),
                         m2 = *reinterpret_cast<const float4*>(mx_s + 8), m3 = *reinterpret_cast<const float4*>(mx_s + 12);
            const float m = fmaxf(fmaxf(fmaxf(fmaxf(m0.x, m0.y), fmaxf(m0.z, m0.w)), fmaxf(fmaxf(m1.x, m1.y), fmaxf(m1.z, m1.w))),
                                  fmaxf(fmaxf(fmaxf(m2.x, m2.y), fmaxf(m2.z, m2.w)), fmaxf(fmaxf(m3.x, m3.y), fmaxf(m3.z, m3.w))));
            if (et < T && eh == 0) ew_s[et] = expf(e_own - m);
        }
        __syncthreads();
        RES_EV(5);
        // ---- (4) own context dims (src/model.py:146) and the softmax denominator
        float Ssum;
        {
            if (ctg < NTG) {
                float c0 = 0.f, c1 = 0.f, s0 = 0.f, s1 = 0.f;
                const float* ew_g = ew_s + ctg * PT;
                if (MM == 2) {
#pragma unroll
                    for (int i = 0; i < MEMR; i += 4) {
                        if (i < PT) {
                            const float4 e4 = *reinterpret_cast<const float4*>(ew_g + i);      // zero beyond T
                            c0 = fmaf(e4.x, memr[i], c0);
                            c1 = fmaf(e4.y, memr[i + 1], c1);
                            c0 = fmaf(e4.z, memr[i + 2], c0);
                            c1 = fmaf(e4.w, memr[i + 3], c1);
                            s0 += e4.x + e4.z;
                            s1 += e4.y + e4.w;
                        }
                    }
                } else {
                    const float* mp = mem_s + (size_t)ctg * PT * DC + cd;
#pragma unroll 4
                    for (int i = 0; i < PT; i += 2) {
                        const float2 e2 = *reinterpret_cast<const float2*>(ew_g + i);
                        c0 = fmaf(e2.x, mp[(size_t)i * DC], c0);
                        c1 = fmaf(e2.y, mp[(size_t)(i + 1) * DC], c1);
                        s0 += e2.x;
                        s1 += e2.y;
                    }
                }
                cpart_s[ctg * DC + cd] = c0 + c1;
                if (cd == 0) spart_s[ctg] = s0 + s1;
            }
            __syncthreads();
            RES_EV(6);
            Ssum = 0.f;
#pragma unroll
            for (int g = 0; g < RWARPS; ++g)                         // NTG <= 16 (DC >= 32)
                if (g < NTG) Ssum += spart_s[g];
            if (ctg < NREP && cd < nd) {
                float c = 0.f;
#pragma unroll
                for (int g = 0; g < RWARPS; ++g)
                    if (g < NTG) c += cpart_s[g * DC + cd];
                ll_st(X + xl.ctx + (size_t)(ctg * BP + b) * EMB + d0 + cd, c / Ssum, tag);
            }
        }
        RES_EV(7);
        // ---- (5) attention weights: alignment output (this CTA's columns), windows of the next step (src/model.py:142-145)
        if (tid < T) {
            const float w = ew_s[tid] / Ssum;
            winp[W0 + tid] = w;
            winc[W0 + tid] += w;
            if (out_idx >= 0 && tid >= al0 && tid < al1) p.align[((size_t)out_idx * p.B + b) * T + tid] = w;
        }
        __syncthreads();
        RES_EV(8);
    }
}

// =====================================================================================================================
// frame CTAs: rows of linear_projection / gate_layer (src/model.py:272-274), the stop rule, and the prenet of the new frame
// (src/model.py:166-171, eval mode: no dropout) -> input of the next decode step.  All weights in registers, contractions
// split over K like the attention CTAs'.
constexpr int PRW = 16;                       // reduction width of the projection rows (MPR <= 16)
struct FSmem {
    int red, mel, y1, red2, total;
};
__host__ __device__ inline FSmem f_smem(int BP) {
    FSmem f;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
    f.red = take(BP * RWARPS * PRW);
    f.mel = take(BP * PROJ_PAD);
    f.y1 = take(BP * PRENET);
    f.red2 = take(BP * RWARPS * 16);
    f.total = o;
    return f;
}

template <int BP>
__device__ __forceinline__ void frame_role(const RParams& p, uint8_t* smem, const XL& xl) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, cta = blockIdx.x;
    const int jf = (int)blockIdx.x - GEMM_CTAS - p.NATT;
    const FSmem L = f_smem(BP);
    float* sm = reinterpret_cast<float*>(smem);
    float *red_s = sm + L.red, *mel_s = sm + L.mel, *y1_s = sm + L.y1, *red2_s = sm + L.red2;
    static_assert(MPR <= PRW && P2R == 32, "frame CTA thread maps");

    // ---- one-off: projection rows MPR * jf + r: thread owns decoder-hidden columns 2 tid, 2 tid + 1 and context column tid;
    // prenet layer 1: thread = (output tid / 2, mel bins 40 (tid & 1) ..); prenet layer 2: thread = (input tid & 255, rows
    // P2R jf + 16 (tid >> 8) ..)
    float wp0[MPR], wp1[MPR], wpc[MPR], w1[40], w2[16];
#pragma unroll
    for (int r = 0; r < MPR; ++r) {
        const int prow = MPR * jf + r, k = 2 * tid;
        const bool ok = prow < PROJ_OUT;
        wp0[r] = ok ? __ldg(p.wproj_slices + ((size_t)(k >> 3) * PROJ_OUT + prow) * 8 + (k & 7)) : 0.f;
        wp1[r] = ok ? __ldg(p.wproj_slices + ((size_t)(k >> 3) * PROJ_OUT + prow) * 8 + (k & 7) + 1) : 0.f;
        wpc[r] = ok ? __ldg(p.wproj_ctxT + (size_t)tid * PROJ_PAD + prow) : 0.f;
    }
    {
        const int o = tid >> 1, half = tid & 1;
#pragma unroll
        for (int i = 0; i < 40; ++i) w1[i] = __ldg(p.prenet1T + (size_t)(40 * half + i) * PRENET + o);
        const int k = tid & (PRENET - 1), rh = tid >> 8;
#pragma unroll
        for (int i = 0; i < 16; ++i) w2[i] = __ldg(p.prenet2T + (size_t)k * PRENET + P2R * jf + 16 * rh + i);
    }
    // prenet of the all-zero go frame (src/model.py:226) is zero (no biases): input of decode step 0
    for (int i = tid; i < P2R * BP * NREP; i += RT) {
        const int r = i % P2R, rb = i / P2R;                                // rb = rep * BP + b
        ll_st(p.xch + xl.p + (size_t)rb * PRENET + P2R * jf + r, 0.f, 1u);
    }
    const float pbias = lane < MPR && MPR * jf + lane < PROJ_OUT ? __ldg(p.proj_bias + MPR * jf + lane) : 0.f;
    __syncthreads();

    const int total_steps = p.max_steps + 1;
    int steps_done = 0;
    bool done[BP];
    int ulen[BP];
#pragma unroll
    for (int q = 0; q < BP; ++q) { done[q] = false; ulen[q] = 0; }
    uint32_t rep_fc[BP], rep_fd[BP];
#pragma unroll
    for (int q = 0; q < BP; ++q) {
        rep_fc[q] = pick_replicas(p.xch, xl.words, xl.ctx + (size_t)q * EMB + tid, BP * EMB);
        rep_fd[q] = pick_replicas(p.xch, xl.words, xl.dh + (size_t)q * RNN + 2 * tid, BP * RNN);
    }

    for (int s = 0; s < total_steps; ++s) {
        const uint32_t tag = (uint32_t)s + 1u;
        uint64_t* X = p.xch + (size_t)(s & (DEPTH - 1)) * xl.words;
        const int out_idx = s - 1;                                   // F6: decode step 0 is discarded
        // ---- (1) context (arrives first) and decoder hidden state -> this CTA's projection rows
#pragma unroll
        for (int q = 0; q < BP; ++q) {
            const float xc = ll_wait1(X + xl.ctx + (size_t)(replica_of(rep_fc[q], s) * BP + q) * EMB + tid, tag);
        RES_EV(0);
            float x0, x1;
            ll_wait2(X + xl.dh + (size_t)(replica_of(rep_fd[q], s) * BP + q) * RNN + 2 * tid, tag, x0, x1);
            RES_EVW(16);
            if (jf == 0) RES_EVT(1);
        RES_EV(1);
            float v[PRW];
#pragma unroll
            for (int r = 0; r < PRW; ++r) v[r] = r < MPR ? fmaf(wpc[r < MPR ? r : 0], xc, fmaf(wp1[r < MPR ? r : 0], x1, wp0[r < MPR ? r : 0] * x0)) : 0.f;
            const float tot = multi_reduce<PRW>(v, lane);
            if ((lane & 1) == 0) red_s[(q * RWARPS + warp) * PRW + (lane >> 1)] = tot;
            RES_EVW(32);
        }
        __syncthreads();
        RES_EV(8);
        if (tid < BP * 32) {
            const int q = tid >> 5, r = tid & 31, prow = MPR * jf + r;
            if (r < MPR && prow < PROJ_OUT) {
                float acc = 0.f;
#pragma unroll
                for (int w = 0; w < RWARPS; ++w) acc += red_s[(q * RWARPS + w) * PRW + r];
                const float v = acc + pbias;
                if (out_idx >= 0) {
                    if (prow < N_MELS) p.mel[((size_t)q * p.max_steps + out_idx) * N_MELS + prow] = v;
                    else p.gate[(size_t)q * p.max_steps + out_idx] = v;
                }
                ll_st(X + xl.mel + (size_t)(q * NFR + jf) * MELW + r, v, tag);
            }
        }
        RES_EV(2);
        // ---- (2) the whole frame: stop rule, prenet
        for (int i0 = warp * 32; i0 < BP * PROJ_OUT; i0 += RT) {          // whole warps; lanes beyond the end re-read the last word
            const int i = min(i0 + lane, BP * PROJ_OUT - 1);
            const int q = i / PROJ_OUT, row = i % PROJ_OUT;
            mel_s[q * PROJ_PAD + row] = ll_wait1(X + xl.mel + (size_t)(q * NFR + row / MPR) * MELW + row % MPR, tag);
        }
        __syncthreads();
        RES_EV(3);
        float g[BP];
#pragma unroll
        for (int q = 0; q < BP; ++q) g[q] = mel_s[q * PROJ_PAD + N_MELS];
        const bool stop = stop_rule<BP>(p, g, out_idx, done, ulen);
        steps_done = s + 1;
        if (s + 1 < total_steps) {
            const int o = tid >> 1, half = tid & 1;
#pragma unroll
            for (int q = 0; q < BP; ++q) {
                float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
                for (int i = 0; i < 40; i += 4) {
                    const float4 m4 = *reinterpret_cast<const float4*>(mel_s + q * PROJ_PAD + 40 * half + i);
                    c0 = fmaf(w1[i], m4.x, c0);
                    c1 = fmaf(w1[i + 1], m4.y, c1);
                    c2 = fmaf(w1[i + 2], m4.z, c2);
                    c3 = fmaf(w1[i + 3], m4.w, c3);
                }
                float y = (c0 + c1) + (c2 + c3);
                y += __shfl_xor_sync(0xffffffffu, y, 1);
                if (half == 0) y1_s[q * PRENET + o] = fmaxf(y, 0.f);
            }
            __syncthreads();
        RES_EV(4);
            uint64_t* Xn = p.xch + (size_t)((s + 1) & (DEPTH - 1)) * xl.words + xl.p;
            const uint32_t ntag = (tag + 1u) | (stop ? STOP_BIT : 0u);
#pragma unroll
            for (int q = 0; q < BP; ++q) {
                const float y = y1_s[q * PRENET + (tid & (PRENET - 1))];
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = w2[i] * y;
                const float tot = multi_reduce<16>(v, lane);
                if ((lane & 1) == 0) red2_s[(q * RWARPS + warp) * 16 + (lane >> 1)] = tot;
            }
            __syncthreads();
            if (tid < BP * NREP * P2R) {       // thread = (utterance, replica, row): 8 warps hold the K halves of a row's 16-row group
                const int r = tid % P2R, rq = tid / P2R, rr = rq % NREP, q = rq / NREP;
                const int w0 = (r >> 4) * (RWARPS / 2);
                float acc = 0.f;
#pragma unroll
                for (int w = 0; w < RWARPS / 2; ++w) acc += red2_s[(q * RWARPS + w0 + w) * 16 + (r & 15)];
                ll_st(Xn + (size_t)(rr * BP + q) * PRENET + P2R * jf + r, fmaxf(acc, 0.f), ntag);
            }
        }
        RES_EV(5);
        if (stop) break;
    }
    if (jf == 0 && tid == 0) {
        const int S = steps_done - 1;                  // output frames (decode step 0 discarded)
        *p.steps_run = S;
        for (int q = 0; q < BP; ++q)
            p.lengths[q] = (p.stop_mode == B200TACO_STOP_PER_UTTERANCE && done[q]) ? ulen[q] : S;
    }
}

template <int BP, int MM>
__global__ void __launch_bounds__(RT, 1) decoder_res_kernel(const RParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const XL xl = xlayout(BP, p.Te, p.NSL);
    const int cta = blockIdx.x;
    if (cta < GEMM_CTAS) gate_role<BP>(p, smem, xl);
    else if (cta < GEMM_CTAS + p.NATT) attn_role<BP, 16 * BP, MM>(p, smem, xl);
    else frame_role<BP>(p, smem, xl);
}

constexpr size_t gate_smem_bytes(int BP) {
    return (size_t)3 * RWARPS * RNN * 4 + (size_t)BP * (2 * RNN + EMB + PRENET + ROWS_PER_CTA) * 4 + 2 * ROWS_PER_CTA * 4;
}
constexpr size_t SMEM_LIMIT = 227 * 1024;

struct Plan {
    int NATT, NSL, DA, DC, resident;
    ASmem as;
    size_t smem;
};
bool make_plan(const b200taco_handle* h, int B, int T, Plan& pl) {
    const int spare = h->num_sms - GEMM_CTAS - NFR;
    if (spare < 2 * B || T > RES_MAX_T) return false;
    pl.NSL = spare / B;
    if (pl.NSL > MAX_NSL) pl.NSL = MAX_NSL;
    pl.NATT = pl.NSL * B;
    pl.DA = (ATT_DIM + pl.NSL - 1) / pl.NSL;
    pl.DC = ((EMB + pl.NSL - 1) / pl.NSL + 3) & ~3;                 // whole 32-byte sectors per producer
    if (pl.DA > 16 * B || NREP * pl.DC > RT || pl.DC < 32) return false;
    const int NTG = RT / pl.DC, PT = (((T + NTG - 1) / NTG) + 3) & ~3;
    const size_t g = gate_smem_bytes(B), f = (size_t)f_smem(B).total * 4;
    static const int mm_knob = [] { const char* e = getenv("TACO_RES_MM"); return e ? atoi(e) : 2; }();       // measurement knob
    pl.resident = (B == 1 && mm_knob >= 2 && PT <= 20) ? 2 : 1;    // memory columns in registers / in shared memory
    pl.as = B == 1 ? a_smem<16>(T, pl.DA, pl.DC, pl.resident == 1) : a_smem<32>(T, pl.DA, pl.DC, pl.resident == 1);
    const size_t a = (size_t)pl.as.total * 4;
    pl.smem = g > a ? g : a;
    if (f > pl.smem) pl.smem = f;
    return pl.smem <= SMEM_LIMIT;
}

template <int BP, int MM>
int launch_res(const RParams& rp, int ncta, size_t smem, cudaStream_t stream) {
    auto kern = decoder_res_kernel<BP, MM>;
    TACO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    TACO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RT, smem));
    TACO_REQUIRE(per_sm >= 1, B200TACO_ERR_UNSUPPORTED, "resident decoder kernel does not fit on an SM (smem %zu B)", smem);
    void* args[] = {const_cast<RParams*>(&rp)};
    TACO_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(ncta), dim3(RT), args, smem, stream));
    launch_counter()++;
    return B200TACO_OK;
}

}  // namespace

bool decoder_res_supported(const b200taco_handle* h, const DecoderArgs& a) {
    static const bool allow = [] { const char* e = getenv("TACO_RES"); return !(e && e[0] == '0'); }();
    if (!allow || h->decoder_path == 1 || h->weight_dtype != 0) return false;
    if (a.mode != 0 || a.n_utts > 0 || a.B < 1 || a.B > 2 || a.sv_ga != nullptr || a.mask_a != nullptr || a.mask_d != nullptr)
        return false;
    Plan pl;
    return make_plan(h, a.B, a.T, pl);
}

size_t decoder_res_scratch_bytes(int B, int T) {
    if (B > 2) return 0;
    const int Te = (T + 15) / 16 * 16;
    return (size_t)DEPTH * xlayout(B, Te, MAX_NSL).words * 8 + (PROF_WORDS + TRACE_WORDS) * 8;
}

int decoder_res_run(const b200taco_handle* h, const DecoderArgs& a, void* scratch, cudaStream_t stream) {
    Plan pl;
    TACO_REQUIRE(make_plan(h, a.B, a.T, pl), B200TACO_ERR_UNSUPPORTED, "resident decoder: unsupported shape");
    const size_t bytes = decoder_res_scratch_bytes(a.B, a.T);
    TACO_CUDA(cudaMemsetAsync(scratch, 0, bytes, stream));
    RParams rp{};
    rp.B = a.B; rp.T = a.T; rp.Tp = (a.T + 3) / 4 * 4; rp.Te = (a.T + 15) / 16 * 16; rp.NATT = pl.NATT; rp.NSL = pl.NSL; rp.DA = pl.DA; rp.DC = pl.DC;
    rp.max_steps = a.max_steps; rp.stop_mode = a.stop_mode; rp.mem_resident = pl.resident; rp.thr = a.gate_threshold;
    static const int bg_sleep = [] { const char* e = getenv("TACO_RES_BG_SLEEP"); return e ? atoi(e) : 0; }();
    static const int bg_backoff = [] { const char* e = getenv("TACO_RES_BG_BACKOFF"); return e ? atoi(e) : 0; }();
    rp.bg_sleep = bg_sleep; rp.bg_backoff = bg_backoff;
    rp.w_attn = static_cast<const uint32_t*>(h->w_attn); rp.w_dec = static_cast<const uint32_t*>(h->w_dec);
    rp.bias_attn = h->bias_attn; rp.bias_dec = h->bias_dec; rp.wq_slices = h->wq_slices; rp.wproj_slices = h->wproj_slices;
    rp.wproj_ctxT = h->wproj_ctxT; rp.proj_bias = h->proj_bias; rp.prenet1T = h->prenet1T; rp.prenet2T = h->prenet2T;
    rp.loc_conv = h->loc_conv; rp.loc_denseT = h->loc_denseT; rp.attn_v = h->attn_v;
    rp.memory = a.memory; rp.pm = a.pm; rp.text_lengths = a.text_lengths;
    rp.mel = a.mel; rp.gate = a.gate; rp.align = a.align; rp.lengths = a.lengths; rp.steps_run = a.steps_run;
    rp.xch = static_cast<uint64_t*>(scratch);
    rp.prof = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(scratch) + bytes - (PROF_WORDS + TRACE_WORDS) * 8);
    static const int trace_step = [] { const char* e = getenv("TACO_RES_TRACE"); return e ? atoi(e) : -1; }();
    rp.trace = trace_step >= 0 ? rp.prof + PROF_WORDS : nullptr;
    rp.trace_step = trace_step;
    const_cast<b200taco_handle*>(h)->last_trace = rp.trace;
    const_cast<b200taco_handle*>(h)->last_profile = rp.prof;
    const int ncta = GEMM_CTAS + pl.NATT + NFR;
    rp.as = pl.as;
    if (a.B == 2) return launch_res<2, 1>(rp, ncta, pl.smem, stream);
    return pl.resident == 2 ? launch_res<1, 2>(rp, ncta, pl.smem, stream) : launch_res<1, 1>(rp, ncta, pl.smem, stream);
}

}  // namespace b200taco
