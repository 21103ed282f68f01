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
//  * the remaining CTAs (20 on a B200) split into "attention" CTAs - each owns a slice of the encoder positions of one
//    utterance (memory and processed memory resident in shared memory) and a few rows of the query projection - and
//    "frame" CTAs, which own rows of the mel / gate projection and of the prenet (weights in registers / shared memory).
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
constexpr int PQR_MAX = RWARPS;               // query projection rows per attention CTA (one warp per row)
constexpr int MPR_MAX = RWARPS;               // mel / gate projection rows per frame CTA (one warp per row, weights in registers)
constexpr int P2R_MAX = 2 * RWARPS;           // prenet layer-2 rows per frame CTA
constexpr int MIN_A = 18, MAX_A = 32;
constexpr int HALO = (LOC_K - 1) / 2;
constexpr int PARTW = 8 + EMB;                // words per attention partial: max, sum, pad, context[512] at +8
constexpr int LOC_LD = 36;
constexpr uint32_t STOP_BIT = 0x80000000u;
constexpr uint32_t SPIN_LIMIT = 1u << 25;     // polls of one word before the kernel traps (~10 s)
constexpr int PROF_WORDS = 24;

static_assert(NREP == 4, "the four gate lanes of a hidden unit publish the four replicas");
static_assert(RT == 4 * ATT_DIM && RT == EMB && RT == 2 * PRENET && 2 * RT == RNN, "thread maps");

struct RParams {
    int B, T, Tp, NATT, NFR, NSL, SL, DPC, PQR, MPR, P2R, max_steps, stop_mode, mem_resident, fine;
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
};

// word offsets of the exchanged vectors inside one ring slot
struct XL {
    int p, ah, ctx, dh, pq, part, mel, w, words;
};
__host__ __device__ inline XL xlayout(int BP, int Tp) {
    XL x;
    int o = 0;
    x.p = o;    o += NREP * BP * PRENET;      // prenet output            frame -> gate
    x.ah = o;   o += NREP * BP * RNN;         // attention hidden         gate -> gate, attention
    x.ctx = o;  o += NREP * BP * EMB;         // attention context        attention -> gate, frame
    x.dh = o;   o += NREP * BP * RNN;         // decoder hidden           gate -> gate, frame
    x.pq = o;   o += BP * ATT_DIM;            // processed query          attention -> attention
    x.part = o; o += MAX_A * PARTW;           // per-slice softmax/context partials  attention -> attention
    x.mel = o;  o += BP * PROJ_PAD;           // mel frame + gate logit   frame -> frame, attention (gate)
    x.w = o;    o += BP * Tp;                 // attention weights        attention -> attention (window halos)
    x.words = (o + 1) & ~1;
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

__device__ __forceinline__ float ll_wait1(const uint64_t* p, uint32_t tag) {
    uint64_t w = ll_ld1(p);
    if (!tag_ok(w, tag)) {
        uint32_t n = 0;
        do {
            w = ll_ld1(p);
            if (++n > SPIN_LIMIT) res_spin_fail();
        } while (!tag_ok(w, tag));
    }
    return __uint_as_float((uint32_t)w);
}

// n words (even, 16-byte aligned) -> n floats in shared memory; returns the OR of the words' stop bits
__device__ __forceinline__ uint32_t ll_gather(const uint64_t* src, float* dst, int n, uint32_t tag, int tid) {
    uint32_t stop = 0u;
    for (int i = tid; i < n / 2; i += RT) {
        uint64_t a, b;
        ll_ld2(src + 2 * i, a, b);
        if (!(tag_ok(a, tag) && tag_ok(b, tag))) {
            uint32_t it = 0;
            do {
                ll_ld2(src + 2 * i, a, b);
                if (++it > SPIN_LIMIT) res_spin_fail();
            } while (!(tag_ok(a, tag) && tag_ok(b, tag)));
        }
        *reinterpret_cast<float2*>(dst + 2 * i) = make_float2(__uint_as_float((uint32_t)a), __uint_as_float((uint32_t)b));
        stop |= (uint32_t)((a | b) >> 32) & STOP_BIT;
    }
    return stop;
}

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

    float acc_a[2][BP], acc_d[2][BP];
#pragma unroll
    for (int b = 0; b < BP; ++b) { acc_a[0][b] = acc_a[1][b] = acc_d[0][b] = acc_d[1][b] = 0.f; }
    float c_a = 0.f, c_d = 0.f;
    const int rep = cta % NREP, unit0 = cta * UNITS_PER_CTA;
    const int total_steps = p.max_steps + 1;
    const bool timing = cta == 0 && tid == 0;
    unsigned long long t_phase[4] = {0ull, 0ull, 0ull, 0ull}, t_last = timing ? global_timer_ns() : 0ull;
    auto stamp = [&](int ph) {
        if (timing) {
            const unsigned long long now = global_timer_ns();
            t_phase[ph] += now - t_last;
            t_last = now;
        }
    };
    int steps_done = 0;
    for (int s = 0; s < total_steps; ++s) {
        const uint32_t tag = (uint32_t)s + 1u;
        uint64_t* X = p.xch + (size_t)(s & (DEPTH - 1)) * xl.words;
        // ---- attention LSTM: only the prenet rows wait for this step's input
        const uint32_t stop = ll_gather(X + xl.p + rep * BP * PRENET, xp, BP * PRENET, tag, tid);
        if (__syncthreads_or((int)stop)) break;
        if (s > 0) stamp(3);
        gemv_reg<BP, 2>(wa_p, xp, PRENET, acc_a, lane);
        reduce_rows<BP>(acc_a, z_s, warp, lane);
        __syncthreads();
        if (warp < BP) lstm_cell(z_s + warp * ROWS_PER_CTA, bias_s, c_a, X + xl.ah + warp * RNN, BP * RNN, unit0, tag, lane);
        stamp(0);
        // ---- background: both cells' rows of the new attention hidden state
        ll_gather(X + xl.ah + rep * BP * RNN, xah, BP * RNN, tag, tid);
        __syncthreads();
        gemv_smem<BP>(sW + (1 * RWARPS + warp) * RNN, xah, acc_d, lane);
        gemv_smem<BP>(sW + (0 * RWARPS + warp) * RNN, xah, acc_a, lane);
        // ---- decoder LSTM: only the context rows wait for the attention
        ll_gather(X + xl.ctx + rep * BP * EMB, xctx, BP * EMB, tag, tid);
        __syncthreads();
        stamp(1);
        gemv_reg<BP, 4>(wd_c, xctx, EMB, acc_d, lane);
        reduce_rows<BP>(acc_d, z_s, warp, lane);
        __syncthreads();
        if (warp < BP)
            lstm_cell(z_s + warp * ROWS_PER_CTA, bias_s + ROWS_PER_CTA, c_d, X + xl.dh + warp * RNN, BP * RNN, unit0, tag, lane);
        stamp(2);
        // ---- background: attention LSTM x context, decoder LSTM x new decoder hidden (both for the next step)
        gemv_reg<BP, 4>(wa_c, xctx, EMB, acc_a, lane);
        ll_gather(X + xl.dh + rep * BP * RNN, xdh, BP * RNN, tag, tid);
        __syncthreads();
        gemv_smem<BP>(sW + (2 * RWARPS + warp) * RNN, xdh, acc_d, lane);
        steps_done = s + 1;
    }
    if (timing) {
        for (int i = 0; i < 4; ++i) p.prof[i] = t_phase[i];
        p.prof[4] = (unsigned long long)steps_done;
    }
}

// =====================================================================================================================
// attention CTAs: location-sensitive attention (src/model.py:117-148) on a slice of one utterance's encoder positions
// plus PQR rows of the query projection
struct ASmem {
    int wloc, v, xah, pq, winp, winc, loc, pre, e, ew, stage, ms, mem, pm, total;
};
__host__ __device__ inline ASmem a_smem(int BP, int SL, int NSL, int DPC, bool resident) {
    ASmem a;
    const int SLp = (SL + 3) & ~3;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
    a.wloc = take(LOC_F * 63);
    a.v = take(ATT_DIM + 4);
    a.xah = take(BP * RNN);
    a.pq = take(ATT_DIM);
    a.winp = take(SLp + 96);
    a.winc = take(SLp + 96);
    a.loc = take((SLp + 16) * LOC_LD);
    a.pre = take(SLp * ATT_DIM);
    a.e = take(SLp + 4);
    a.ew = take(SLp + 4);
    a.stage = take(NSL * DPC);
    a.ms = take(2 * MAX_A);
    a.mem = take(resident ? SL * EMB : 0);
    a.pm = take(resident ? SL * ATT_DIM : 0);
    a.total = o;
    return a;
}

template <int BP>
__device__ __forceinline__ void attn_role(const RParams& p, uint8_t* smem, const XL& xl) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j = (int)blockIdx.x - GEMM_CTAS, cta = blockIdx.x;
    const int NSL = p.NSL, SL = p.SL, T = p.T, DPC = p.DPC, PQR = p.PQR;
    const int b = j / NSL, sl = j % NSL;                       // NATT == BP * NSL
    const int t0 = sl * SL, n = max(0, min(T, t0 + SL) - t0);
    const int len = p.text_lengths ? (int)p.text_lengths[b] : T;
    const ASmem L = a_smem(BP, SL, NSL, DPC, p.mem_resident != 0);
    float* sm = reinterpret_cast<float*>(smem);
    float *wloc_s = sm + L.wloc, *v_s = sm + L.v, *xah = sm + L.xah, *pq_s = sm + L.pq, *winp = sm + L.winp, *winc = sm + L.winc,
          *loc_s = sm + L.loc, *pre_s = sm + L.pre, *e_s = sm + L.e, *ew_s = sm + L.ew, *stage_s = sm + L.stage, *ms_s = sm + L.ms;
    const int SLp = (SL + 3) & ~3;

    // ---- one-off: this warp's query-projection row and this thread's location-dense column into registers
    const int pq_row = PQR * j + warp;
    const bool has_row = warp < PQR && pq_row < ATT_DIM;
    float wq[32], wld[LOC_F];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = jj * 128 + lane * 4 + i;
            wq[jj * 4 + i] = has_row ? __ldg(p.wq_slices + ((size_t)(k >> 3) * ATT_DIM + pq_row) * 8 + (k & 7)) : 0.f;
        }
#pragma unroll
    for (int f = 0; f < LOC_F; ++f) wld[f] = __ldg(p.loc_denseT + f * ATT_DIM + (tid & (ATT_DIM - 1)));
    for (int i = tid; i < LOC_F * 2 * LOC_K; i += RT) wloc_s[(i / 62) * 63 + (i % 62)] = p.loc_conv[i];
    for (int i = tid; i < ATT_DIM + 2; i += RT) v_s[i] = p.attn_v[i];
    for (int i = tid; i < SLp + 96; i += RT) { winp[i] = 0.f; winc[i] = 0.f; }
    const float* memP = p.memory + ((size_t)b * T + t0) * EMB;
    const float* pmP = p.pm + ((size_t)b * T + t0) * ATT_DIM;
    if (p.mem_resident) {
        float *mem_s = sm + L.mem, *pm_s = sm + L.pm;
        for (int i = tid; i < n * EMB / 4; i += RT) reinterpret_cast<float4*>(mem_s)[i] = __ldg(reinterpret_cast<const float4*>(memP) + i);
        for (int i = tid; i < n * ATT_DIM / 4; i += RT) reinterpret_cast<float4*>(pm_s)[i] = __ldg(reinterpret_cast<const float4*>(pmP) + i);
        memP = mem_s;
        pmP = pm_s;
    }
    __syncthreads();
    const float4 va = *reinterpret_cast<const float4*>(v_s + (lane & 15) * 8), vb = *reinterpret_cast<const float4*>(v_s + (lane & 15) * 8 + 4);
    const float v_bias = v_s[ATT_DIM], e_scale = v_s[ATT_DIM + 1];

    const int total_steps = p.max_steps + 1;
    bool done[BP];
    int ulen[BP];
#pragma unroll
    for (int q = 0; q < BP; ++q) { done[q] = false; ulen[q] = 0; }
    const int rep = cta % NREP;
    const bool ftiming = p.fine && j == 0 && tid == 0;
    auto fnow = [&]() -> unsigned long long { return p.fine == 2 ? (unsigned long long)clock64() : global_timer_ns(); };
    unsigned long long t_fine[9] = {}, t_fl = ftiming ? fnow() : 0ull;
    auto fstamp = [&](int i) {
        if (ftiming) {
            const unsigned long long now = fnow();
            t_fine[i] += now - t_fl;
            t_fl = now;
        }
    };

    for (int s = 0; s < total_steps; ++s) {
        const uint32_t tag = (uint32_t)s + 1u;
        uint64_t* X = p.xch + (size_t)(s & (DEPTH - 1)) * xl.words;
        const int out_idx = s - 1;                                   // F6: decode step 0 is discarded

        // ---- (0) before the attention hidden state arrives: windows of the previous / cumulative weights with their
        // halos, location convolution, and pre[t][a] = pm[t][a] + Wld . loc[:, t]
        if (s > 0) {
            const uint64_t* Xw = p.xch + (size_t)((s - 1) & (DEPTH - 1)) * xl.words + xl.w + (size_t)b * p.Tp;
            for (int i = tid; i < n + 2 * HALO; i += RT) {
                const int t = t0 - HALO + i;
                if (t >= 0 && t < T) {
                    const float w = ll_wait1(Xw + t, tag - 1u);
                    winp[i] = w;
                    winc[i] += w;
                }
            }
            __syncthreads();
        }
        {   // location conv (src/model.py:125-126): 32 filters x (2 x 31 taps); a thread owns filter f at positions t, t+16
            const int f = lane;
            for (int t = warp; t < n; t += 2 * RWARPS) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float* x = (c ? winc : winp) + t;
#pragma unroll
                    for (int k = 0; k < LOC_K; ++k) {
                        const float w = wloc_s[f * 63 + c * LOC_K + k];
                        a0 = fmaf(w, x[k], a0);
                        a1 = fmaf(w, x[k + RWARPS], a1);
                    }
                }
                loc_s[t * LOC_LD + f] = a0;
                loc_s[(t + RWARPS) * LOC_LD + f] = a1;       // rows n .. n+15 exist (never read when beyond n)
            }
        }
        __syncthreads();
        {   // location dense (src/model.py:127) + processed memory
            const int a = tid & (ATT_DIM - 1), tq = tid >> 7;
            for (int t = tq; t < n; t += 4) {
                float x = pmP[(size_t)t * ATT_DIM + a];
#pragma unroll
                for (int f4 = 0; f4 < LOC_F; f4 += 4) {
                    const float4 l = *reinterpret_cast<const float4*>(loc_s + t * LOC_LD + f4);
                    x = fmaf(wld[f4], l.x, x);
                    x = fmaf(wld[f4 + 1], l.y, x);
                    x = fmaf(wld[f4 + 2], l.z, x);
                    x = fmaf(wld[f4 + 3], l.w, x);
                }
                pre_s[t * ATT_DIM + a] = x;
            }
        }
        fstamp(0);
        // the previous frame's gate logits: every CTA evaluates the stop rule itself
        if (s > 0) {
            const uint64_t* Xm = p.xch + (size_t)((s - 1) & (DEPTH - 1)) * xl.words + xl.mel;
            float g[BP];
#pragma unroll
            for (int q = 0; q < BP; ++q) g[q] = ll_wait1(Xm + q * PROJ_PAD + N_MELS, tag - 1u);
            if (stop_rule<BP>(p, g, out_idx - 1, done, ulen)) break;
        }
        // ---- (1) attention hidden state -> this CTA's rows of the processed query (src/model.py:123), one warp per row
        ll_gather(X + xl.ah + rep * BP * RNN, xah, BP * RNN, tag, tid);
        __syncthreads();
        fstamp(1);
        if (warp < PQR) {
#pragma unroll
            for (int q = 0; q < BP; ++q) {
                float acc = 0.f;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float4 x4 = reinterpret_cast<const float4*>(xah + q * RNN)[jj * 32 + lane];
                    acc = fmaf(wq[jj * 4 + 0], x4.x, acc);
                    acc = fmaf(wq[jj * 4 + 1], x4.y, acc);
                    acc = fmaf(wq[jj * 4 + 2], x4.z, acc);
                    acc = fmaf(wq[jj * 4 + 3], x4.w, acc);
                }
                acc = warp_sum(acc);
                if (lane == 0 && has_row) ll_st(X + xl.pq + q * ATT_DIM + pq_row, acc, tag);
            }
        }
        fstamp(2);
        // ---- (2) energies v . tanh(pq + pre) (src/model.py:128-132): one half-warp per position, 8 attention dims per lane
        ll_gather(X + xl.pq + b * ATT_DIM, pq_s, ATT_DIM, tag, tid);
        __syncthreads();
        fstamp(3);
        {
            const int hl = lane & 15;
            const float4 qa = *reinterpret_cast<const float4*>(pq_s + hl * 8), qb = *reinterpret_cast<const float4*>(pq_s + hl * 8 + 4);
            for (int t = 2 * warp + (lane >> 4); t < ((n + 1) & ~1); t += 2 * RWARPS) {
                const int tt = t < n ? t : n - 1;
                const float4 ra = *reinterpret_cast<const float4*>(pre_s + tt * ATT_DIM + hl * 8);
                const float4 rb = *reinterpret_cast<const float4*>(pre_s + tt * ATT_DIM + hl * 8 + 4);
                float x0 = va.x * tanh_fast(qa.x + ra.x), x1 = va.y * tanh_fast(qa.y + ra.y);
                x0 = fmaf(va.z, tanh_fast(qa.z + ra.z), x0);
                x1 = fmaf(va.w, tanh_fast(qa.w + ra.w), x1);
                x0 = fmaf(vb.x, tanh_fast(qb.x + rb.x), x0);
                x1 = fmaf(vb.y, tanh_fast(qb.y + rb.y), x1);
                x0 = fmaf(vb.z, tanh_fast(qb.z + rb.z), x0);
                x1 = fmaf(vb.w, tanh_fast(qb.w + rb.w), x1);
                float x = x0 + x1;
#pragma unroll
                for (int sh = 8; sh > 0; sh >>= 1) x += __shfl_xor_sync(0xffffffffu, x, sh);
                if (hl == 0 && t < n) e_s[t] = t0 + t >= len ? -1e9f : (x + v_bias) * e_scale;     // masked_fill, src/model.py:132
            }
        }
        __syncthreads();
        // ---- (3) slice-local softmax statistics (every warp, redundantly) and un-normalised context partial
        float m_loc = -INFINITY;
        for (int i = lane; i < n; i += 32) m_loc = fmaxf(m_loc, e_s[i]);
        m_loc = warp_max(m_loc);
        if (tid < n) ew_s[tid] = expf(e_s[tid] - m_loc);
        __syncthreads();
        fstamp(4);
        {
            float c0 = 0.f, c1 = 0.f;
            int i = 0;
            for (; i + 1 < n; i += 2) {
                c0 = fmaf(ew_s[i], memP[(size_t)i * EMB + tid], c0);
                c1 = fmaf(ew_s[i + 1], memP[(size_t)(i + 1) * EMB + tid], c1);
            }
            if (i < n) c0 = fmaf(ew_s[i], memP[(size_t)i * EMB + tid], c0);
            ll_st(X + xl.part + (size_t)j * PARTW + 8 + tid, c0 + c1, tag);
            if (warp == RWARPS - 1) {
                float ssum = 0.f;
                for (int q = lane; q < n; q += 32) ssum += ew_s[q];
                ssum = warp_sum(ssum);
                if (lane == 0) {
                    ll_st(X + xl.part + (size_t)j * PARTW, n > 0 ? m_loc : -INFINITY, tag);
                    ll_st(X + xl.part + (size_t)j * PARTW + 1, ssum, tag);
                }
            }
        }
        fstamp(5);
        // ---- (4) merge: this CTA finishes context dims [sl * DPC, +DPC) of its utterance
        const int d0 = sl * DPC, nd = max(0, min(EMB, d0 + DPC) - d0);
        const uint64_t* Xp = X + xl.part + (size_t)b * NSL * PARTW;
        for (int i = tid; i < NSL * nd; i += RT) {
            const int q = i / nd, dd = i % nd;
            stage_s[q * DPC + dd] = ll_wait1(Xp + (size_t)q * PARTW + 8 + d0 + dd, tag);
        }
        for (int i = RT - 1 - tid; i < 2 * NSL; i += RT) ms_s[i] = ll_wait1(Xp + (size_t)(i >> 1) * PARTW + (i & 1), tag);
        __syncthreads();
        fstamp(6);
        const float mq = lane < NSL ? ms_s[2 * lane] : -INFINITY;
        const float Msc = warp_max(mq);
        const float scale = lane < NSL ? expf(mq - Msc) : 0.f;
        const float Ssum = warp_sum(lane < NSL ? scale * ms_s[2 * lane + 1] : 0.f);
        const float mine = __shfl_sync(0xffffffffu, scale, sl);
        if (warp * 32 < nd) {
            const int dd = warp * 32 + lane;
            float c = 0.f;
            for (int q = 0; q < NSL; ++q) {
                const float sq = __shfl_sync(0xffffffffu, scale, q);
                c = fmaf(sq, dd < nd ? stage_s[q * DPC + dd] : 0.f, c);
            }
            c = c / Ssum;
            if (dd < nd) {
#pragma unroll
                for (int r = 0; r < NREP; ++r) ll_st(X + xl.ctx + (size_t)(r * BP + b) * EMB + d0 + dd, c, tag);
            }
        }
        fstamp(8);
        // ---- (5) attention weights of the slice: alignment output, next step's windows (src/model.py:142-145)
        for (int i = tid; i < n; i += RT) {
            const float w = ew_s[i] * mine / Ssum;
            if (out_idx >= 0) p.align[((size_t)out_idx * p.B + b) * T + t0 + i] = w;
            ll_st(X + xl.w + (size_t)b * p.Tp + t0 + i, w, tag);
        }
        fstamp(7);
    }
    if (ftiming) {
        for (int i = 0; i < 8; ++i) p.prof[5 + i] = t_fine[i];
        p.prof[19] = t_fine[8];
    }
}

// =====================================================================================================================
// frame CTAs: rows of linear_projection / gate_layer (src/model.py:272-274), the stop rule, and the prenet of the new frame
// (src/model.py:166-171, eval mode: no dropout) -> input of the next decode step
struct FSmem {
    int w2, xcat, mel, y1p, y1, total;
};
__host__ __device__ inline FSmem f_smem(int BP, int P2R) {
    FSmem f;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
    f.w2 = take(P2R_MAX * PRENET);          // rows beyond P2R are zero
    f.xcat = take(BP * (RNN + EMB));
    f.mel = take(BP * PROJ_PAD);
    f.y1p = take(2 * BP * PRENET);
    f.y1 = take(BP * PRENET);
    f.total = o;
    return f;
}

template <int BP>
__device__ __forceinline__ void frame_role(const RParams& p, uint8_t* smem, const XL& xl) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, cta = blockIdx.x;
    const int jf = (int)blockIdx.x - GEMM_CTAS - p.NATT;
    const int MPR = p.MPR, P2R = p.P2R;
    const FSmem L = f_smem(BP, P2R);
    float* sm = reinterpret_cast<float*>(smem);
    float *w2_s = sm + L.w2, *xcat = sm + L.xcat, *mel_s = sm + L.mel, *y1p_s = sm + L.y1p, *y1_s = sm + L.y1;
    constexpr int KC = RNN + EMB;                       // projection input [decoder hidden | context]

    // ---- one-off: this warp's projection row (warp w < MPR owns row MPR * jf + w; lane l owns k = 128 jj + 4 l + i of
    // [decoder hidden | context]) and prenet layer 1 into registers
    float wp[48], w1[40];
    const int prow = MPR * jf + warp;
    const bool has_prow = warp < MPR && prow < PROJ_OUT;
#pragma unroll
    for (int jj = 0; jj < 12; ++jj)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = jj * 128 + lane * 4 + i;
            float w = 0.f;
            if (has_prow)
                w = k < RNN ? __ldg(p.wproj_slices + ((size_t)(k >> 3) * PROJ_OUT + prow) * 8 + (k & 7))
                            : __ldg(p.wproj_ctxT + (size_t)(k - RNN) * PROJ_PAD + prow);
            wp[jj * 4 + i] = w;
        }
    const float pbias = has_prow ? __ldg(p.proj_bias + prow) : 0.f;
    {
        const int o = tid & (PRENET - 1), half = tid >> 8;
#pragma unroll
        for (int i = 0; i < 40; ++i) w1[i] = __ldg(p.prenet1T + (size_t)(40 * half + i) * PRENET + o);
    }
    for (int i = tid; i < P2R_MAX * PRENET; i += RT) {
        const int k = i & (PRENET - 1), r = i >> 8, row = P2R * jf + r;
        w2_s[r * PRENET + k] = r < P2R && row < PRENET ? __ldg(p.prenet2T + (size_t)k * PRENET + row) : 0.f;
    }
    // prenet of the all-zero go frame (src/model.py:226) is zero (no biases): input of decode step 0
    for (int i = tid; i < P2R * BP * NREP; i += RT) {
        const int r = i % P2R, rb = i / P2R, row = P2R * jf + r;          // rb = rep * BP + b
        if (row < PRENET) ll_st(p.xch + xl.p + (size_t)rb * PRENET + row, 0.f, 1u);
    }
    __syncthreads();

    const int total_steps = p.max_steps + 1;
    int steps_done = 0;
    bool done[BP];
    int ulen[BP];
#pragma unroll
    for (int q = 0; q < BP; ++q) { done[q] = false; ulen[q] = 0; }
    const int rep = cta % NREP;
    const bool ftiming = p.fine && jf == 0 && tid == 0;
    auto fnow = [&]() -> unsigned long long { return p.fine == 2 ? (unsigned long long)clock64() : global_timer_ns(); };
    unsigned long long t_fine[6] = {}, t_fl = ftiming ? fnow() : 0ull;
    auto fstamp = [&](int i) {
        if (ftiming) {
            const unsigned long long now = fnow();
            t_fine[i] += now - t_fl;
            t_fl = now;
        }
    };

    for (int s = 0; s < total_steps; ++s) {
        const uint32_t tag = (uint32_t)s + 1u;
        uint64_t* X = p.xch + (size_t)(s & (DEPTH - 1)) * xl.words;
        const int out_idx = s - 1;                                   // F6: decode step 0 is discarded
        // ---- (1) context (arrives first) and decoder hidden state -> this CTA's projection rows
#pragma unroll
        for (int q = 0; q < BP; ++q) ll_gather(X + xl.ctx + (size_t)(rep * BP + q) * EMB, xcat + q * KC + RNN, EMB, tag, tid);
        fstamp(0);
#pragma unroll
        for (int q = 0; q < BP; ++q) ll_gather(X + xl.dh + (size_t)(rep * BP + q) * RNN, xcat + q * KC, RNN, tag, tid);
        __syncthreads();
        fstamp(1);
        if (has_prow) {
#pragma unroll
            for (int q = 0; q < BP; ++q) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int jj = 0; jj < 12; ++jj) {
                    const float4 x4 = reinterpret_cast<const float4*>(xcat + q * KC)[jj * 32 + lane];
                    a0 = fmaf(wp[jj * 4 + 0], x4.x, a0);
                    a1 = fmaf(wp[jj * 4 + 1], x4.y, a1);
                    a2 = fmaf(wp[jj * 4 + 2], x4.z, a2);
                    a3 = fmaf(wp[jj * 4 + 3], x4.w, a3);
                }
                const float v = warp_sum((a0 + a1) + (a2 + a3)) + pbias;
                if (lane == 0) {
                    if (out_idx >= 0) {
                        if (prow < N_MELS) p.mel[((size_t)q * p.max_steps + out_idx) * N_MELS + prow] = v;
                        else p.gate[(size_t)q * p.max_steps + out_idx] = v;
                    }
                    ll_st(X + xl.mel + q * PROJ_PAD + prow, v, tag);
                    if (prow == N_MELS) ll_st(X + xl.mel + q * PROJ_PAD + prow + 1, 0.f, tag);      // pad to an even word count
                }
            }
        }
        fstamp(2);
        // ---- (2) the whole frame: stop rule, prenet
#pragma unroll
        for (int q = 0; q < BP; ++q) ll_gather(X + xl.mel + q * PROJ_PAD, mel_s + q * PROJ_PAD, N_MELS + 2, tag, tid);
        __syncthreads();
        fstamp(3);
        float g[BP];
#pragma unroll
        for (int q = 0; q < BP; ++q) g[q] = mel_s[q * PROJ_PAD + N_MELS];
        const bool stop = stop_rule<BP>(p, g, out_idx, done, ulen);
        steps_done = s + 1;
        if (s + 1 < total_steps) {
            const int o = tid & (PRENET - 1), half = tid >> 8;
#pragma unroll
            for (int q = 0; q < BP; ++q) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int i = 0; i < 40; i += 4) {
                    const float4 m4 = *reinterpret_cast<const float4*>(mel_s + q * PROJ_PAD + 40 * half + i);
                    a0 = fmaf(w1[i], m4.x, a0);
                    a1 = fmaf(w1[i + 1], m4.y, a1);
                    a0 = fmaf(w1[i + 2], m4.z, a0);
                    a1 = fmaf(w1[i + 3], m4.w, a1);
                }
                y1p_s[(half * BP + q) * PRENET + o] = a0 + a1;
            }
            __syncthreads();
            if (tid < PRENET) {
#pragma unroll
                for (int q = 0; q < BP; ++q) y1_s[q * PRENET + tid] = fmaxf(y1p_s[q * PRENET + tid] + y1p_s[(BP + q) * PRENET + tid], 0.f);
            }
            __syncthreads();
            fstamp(4);
            uint64_t* Xn = p.xch + (size_t)((s + 1) & (DEPTH - 1)) * xl.words + xl.p;
            const uint32_t ntag = (tag + 1u) | (stop ? STOP_BIT : 0u);
            {   // one half-warp per row: lane hl owns k = 64 jj + 4 hl + i
                const int r = 2 * warp + (lane >> 4), hl = lane & 15, row = P2R * jf + r;
#pragma unroll
                for (int q = 0; q < BP; ++q) {
                    float a0 = 0.f, a1 = 0.f;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float4 w4 = reinterpret_cast<const float4*>(w2_s + r * PRENET)[jj * 16 + hl];
                        const float4 x4 = reinterpret_cast<const float4*>(y1_s + q * PRENET)[jj * 16 + hl];
                        a0 = fmaf(w4.x, x4.x, a0);
                        a1 = fmaf(w4.y, x4.y, a1);
                        a0 = fmaf(w4.z, x4.z, a0);
                        a1 = fmaf(w4.w, x4.w, a1);
                    }
                    float acc = a0 + a1;
#pragma unroll
                    for (int sh = 8; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
                    acc = fmaxf(acc, 0.f);
                    if (hl < NREP && r < P2R && row < PRENET) ll_st(Xn + (size_t)(hl * BP + q) * PRENET + row, acc, ntag);
                }
            }
        }
        fstamp(5);
        if (stop) break;
    }
    if (ftiming) for (int i = 0; i < 6; ++i) p.prof[13 + i] = t_fine[i];
    if (jf == 0 && tid == 0) {
        const int S = steps_done - 1;                  // output frames (decode step 0 discarded)
        *p.steps_run = S;
        for (int q = 0; q < BP; ++q)
            p.lengths[q] = (p.stop_mode == B200TACO_STOP_PER_UTTERANCE && done[q]) ? ulen[q] : S;
    }
}

template <int BP>
__global__ void __launch_bounds__(RT, 1) decoder_res_kernel(const RParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const XL xl = xlayout(BP, p.Tp);
    const int cta = blockIdx.x;
    if (cta < GEMM_CTAS) gate_role<BP>(p, smem, xl);
    else if (cta < GEMM_CTAS + p.NATT) attn_role<BP>(p, smem, xl);
    else frame_role<BP>(p, smem, xl);
}

constexpr size_t gate_smem_bytes(int BP) {
    return (size_t)3 * RWARPS * RNN * 4 + (size_t)BP * (2 * RNN + EMB + PRENET + ROWS_PER_CTA) * 4 + 2 * ROWS_PER_CTA * 4;
}
constexpr size_t SMEM_LIMIT = 227 * 1024;

struct Plan {
    int NATT, NFR, NSL, SL, DPC, PQR, MPR, P2R, resident;
    size_t smem;
};
bool make_plan(const b200taco_handle* h, int B, int T, Plan& pl) {
    int na = h->num_sms - GEMM_CTAS;
    if (na > MAX_A) na = MAX_A;
    if (na < MIN_A) return false;
    pl.NFR = na / 2;                                               // frame CTAs; the rest (a multiple of B) are attention CTAs
    pl.NATT = (na - pl.NFR) / B * B;
    pl.NSL = pl.NATT / B;
    if (pl.NSL < 1 || pl.NSL > 32) return false;
    pl.SL = (T + pl.NSL - 1) / pl.NSL;
    pl.DPC = ((EMB + pl.NSL - 1) / pl.NSL + 1) & ~1;
    pl.PQR = (ATT_DIM + pl.NATT - 1) / pl.NATT;
    pl.MPR = (PROJ_OUT + pl.NFR - 1) / pl.NFR;
    pl.P2R = (PRENET + pl.NFR - 1) / pl.NFR;
    if (pl.PQR > PQR_MAX || pl.MPR > MPR_MAX || pl.P2R > P2R_MAX) return false;
    const size_t g = gate_smem_bytes(B), f = (size_t)f_smem(B, pl.P2R).total * 4;
    size_t a = (size_t)a_smem(B, pl.SL, pl.NSL, pl.DPC, true).total * 4;
    pl.resident = 1;
    if (a > SMEM_LIMIT) {
        a = (size_t)a_smem(B, pl.SL, pl.NSL, pl.DPC, false).total * 4;
        pl.resident = 0;
    }
    pl.smem = g > a ? g : a;
    if (f > pl.smem) pl.smem = f;
    return pl.smem <= SMEM_LIMIT;
}

template <int BP>
int launch_res(const RParams& rp, int ncta, size_t smem, cudaStream_t stream) {
    auto kern = decoder_res_kernel<BP>;
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
    const int Tp = (T + 3) / 4 * 4;
    return (size_t)DEPTH * xlayout(B, Tp).words * 8 + PROF_WORDS * 8;
}

int decoder_res_run(const b200taco_handle* h, const DecoderArgs& a, void* scratch, cudaStream_t stream) {
    Plan pl;
    TACO_REQUIRE(make_plan(h, a.B, a.T, pl), B200TACO_ERR_UNSUPPORTED, "resident decoder: unsupported shape");
    const size_t bytes = decoder_res_scratch_bytes(a.B, a.T);
    TACO_CUDA(cudaMemsetAsync(scratch, 0, bytes, stream));
    RParams rp{};
    rp.B = a.B; rp.T = a.T; rp.Tp = (a.T + 3) / 4 * 4; rp.NATT = pl.NATT; rp.NFR = pl.NFR; rp.NSL = pl.NSL; rp.SL = pl.SL;
    rp.DPC = pl.DPC; rp.PQR = pl.PQR; rp.MPR = pl.MPR; rp.P2R = pl.P2R;
    rp.max_steps = a.max_steps; rp.stop_mode = a.stop_mode; rp.mem_resident = pl.resident; rp.thr = a.gate_threshold;
    static const int fine = [] { const char* e = getenv("TACO_RES_FINE"); return e ? atoi(e) : 0; }();
    rp.fine = fine;
    rp.w_attn = static_cast<const uint32_t*>(h->w_attn); rp.w_dec = static_cast<const uint32_t*>(h->w_dec);
    rp.bias_attn = h->bias_attn; rp.bias_dec = h->bias_dec; rp.wq_slices = h->wq_slices; rp.wproj_slices = h->wproj_slices;
    rp.wproj_ctxT = h->wproj_ctxT; rp.proj_bias = h->proj_bias; rp.prenet1T = h->prenet1T; rp.prenet2T = h->prenet2T;
    rp.loc_conv = h->loc_conv; rp.loc_denseT = h->loc_denseT; rp.attn_v = h->attn_v;
    rp.memory = a.memory; rp.pm = a.pm; rp.text_lengths = a.text_lengths;
    rp.mel = a.mel; rp.gate = a.gate; rp.align = a.align; rp.lengths = a.lengths; rp.steps_run = a.steps_run;
    rp.xch = static_cast<uint64_t*>(scratch);
    rp.prof = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(scratch) + bytes - PROF_WORDS * 8);
    const_cast<b200taco_handle*>(h)->last_profile = rp.prof;
    const int ncta = GEMM_CTAS + pl.NATT + pl.NFR;
    return a.B == 1 ? launch_res<1>(rp, ncta, pl.smem, stream) : launch_res<2>(rp, ncta, pl.smem, stream);
}

}  // namespace b200taco
