// Backward-pass kernels of the teacher-forced decoder (reference Decoder.forward / _decode_step in train mode,
// src/model.py:246-297; LocationSensitiveAttention :117-148).  First correct fp32 versions: the host walks the decode
// steps in reverse and launches these kernels plus dense-layer calls per step (train_api.cu).
#include <cooperative_groups.h>

#include <cstdlib>

#include "handle.h"
#include "umma.cuh"

namespace cg = cooperative_groups;

namespace b200taco {

constexpr int ATTN_PACC = ATT_DIM * LOC_F + LOC_F * 2 * LOC_K + ATT_DIM + 64;     // 4096 + 1984 + 128 + pad

// TACO_PDL = bit mask of the kernels launched with programmatic stream serialization (PdlKernel; measurements), default all
bool pdl_enabled(int which) {
    static const int mask = [] { const char* e = getenv("TACO_PDL"); return e ? atoi(e) : 15; }();
    return (mask >> which) & 1;
}

namespace {

inline int blocks_for(size_t n) { return (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256); }

// One (utterance, unit) of the LSTMCell backward: dh = gradient w.r.t. the hidden state BEFORE dropout.  Shared by
// lstm_cell_bwd_kernel and the tail of attention_bwd_v3_kernel.
__device__ __forceinline__ void lstm_cell_bwd_unit(const float* __restrict__ gates, const float* __restrict__ c,
                                                   const float* __restrict__ c_prev, float dh, float* __restrict__ dc,
                                                   float* __restrict__ dz, __nv_bfloat16* __restrict__ dzp, int b, int u) {
    const int i = b * RNN + u;
    const float* g4 = gates + (size_t)b * 4 * RNN + u;
    const float gi = g4[0], gf = g4[RNN], gg = g4[2 * RNN], go = g4[3 * RNN];
    const float tc = tanhf(c[i]);
    const float dct = dc[i] + dh * go * (1.f - tc * tc);
    const float cp = c_prev ? c_prev[i] : 0.f;
    float* d4 = dz + (size_t)b * 4 * RNN + u;
    const float dzv[4] = {dct * gg * gi * (1.f - gi), dct * cp * gf * (1.f - gf), dct * gi * (1.f - gg * gg), dh * tc * go * (1.f - go)};
#pragma unroll
    for (int g = 0; g < 4; ++g) d4[g * RNN] = dzv[g];
    dc[i] = dct * gf;
    // second copy for the tensor-core data-gradient GEMM that follows (skinny_tc_kernel): hi/lo bf16, rows b and 64 + b of the
    // M = 128 operand, as UMMA K-major tiles [k / 64][128 x 64] that the TMA engine copies straight into its pipeline stages
    if (dzp != nullptr) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int k = g * RNN + u;
            __nv_bfloat16 hi, lo;
            umma::split_bf16(dzv[g], hi, lo);
            __nv_bfloat16* tile = dzp + (size_t)(k >> 6) * (128 * 64);
            tile[umma::tile_offset(b, k & 63, 8)] = hi;
            tile[umma::tile_offset(64 + b, k & 63, 8)] = lo;
        }
    }
}

// ---- LSTMCell backward (pointwise part).  gates [B][4096] = i,f,g,o post-activation (gate*1024 + unit), c [B][1024],
// c_prev [B][1024] or null (zero state), dh_post [B][1024] = gradient w.r.t. the (dropped) hidden state, mask or null,
// dc [B][1024] carried cell gradient (in/out), dz [B][4096] gradient w.r.t. the gate pre-activations (out).
// dh_extra [B][ld_extra] (nullable): a second addend of dh_post that is not carried (the projection's gradient of this step).
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c, const float* __restrict__ c_prev,
                                     float* __restrict__ dh_post, const float* __restrict__ dh_extra, int ld_extra,
                                     const float* __restrict__ mask, float scale, float* __restrict__ dc, float* __restrict__ dz,
                                     __nv_bfloat16* __restrict__ dzp, int B) {
    pdl_wait();
    pdl_trigger();          // only now: the NEXT kernel of the chain may start its set-up while this one works (run-ahead of one kernel)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * RNN) return;
    const int b = i / RNN, u = i % RNN;
    const float dh = (dh_post[i] + (dh_extra ? dh_extra[(size_t)b * ld_extra + u] : 0.f)) * (mask ? mask[i] * scale : 1.f);
    dh_post[i] = 0.f;          // consumed: the data-gradient GEMM that follows accumulates the previous step's gradient onto it
    lstm_cell_bwd_unit(gates, c, c_prev, dh, dc, dz, dzp, b, u);
}

// cum[t] = sum_{tau < t} align[tau]  (exclusive prefix sum over the step axis), one thread per (b, position)
__global__ void attn_cumsum_kernel(const float* __restrict__ align, float* __restrict__ cum, int S, int BT) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BT) return;
    float acc = 0.f;
    for (int t = 0; t < S; ++t) {
        cum[(size_t)t * BT + i] = acc;
        acc += align[(size_t)t * BT + i];
    }
}

// dst[i] (+)= a[i] elementwise helpers
__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ a, int n, int ld_dst, int ld_a, int cols, int accumulate) {
    // dst[r][c] (+)= a[r][c] for r < n, c < cols with row strides ld_dst / ld_a
    const size_t total = (size_t)n * cols;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = i / cols, cc = i % cols;
        const float v = a[(size_t)r * ld_a + cc];
        float* d = dst + (size_t)r * ld_dst + cc;
        *d = accumulate ? *d + v : v;
    }
}

// relu + dropout backward: dx = dy * mask * scale * (y > 0)   (y is the layer output AFTER dropout; y > 0 <=> relu' = 1 and kept)
__global__ void relu_drop_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ mask,
                                     float scale, float* __restrict__ dx, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float keep = mask ? mask[i] * scale : 1.f;
        dx[i] = y[i] > 0.f ? dy[i] * keep : 0.f;
    }
}
__global__ void relu_drop_fwd_kernel(float* __restrict__ y, const float* __restrict__ mask, float scale, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = mask ? y[i] * mask[i] * scale : y[i];
}

// ---- attention backward of ONE decode step, one CTA per utterance.
// Recomputes the location features and tanh arguments from the saved attention weights.
//   in : w [B][T] (this step's weights), prev [B][T] or null (step 0), cum [B][T], pq [B][128] (Wq . ah_t), pm [B][T][128],
//        memory [B][T][512], dctx [B][512], dw_ext [B][T] (external d(align) or null), dprev_next [B][T] (gradient that the
//        NEXT step's location conv sent to its `prev` input = this step's weights), gsum [B][T] (sum over later steps of
//        the gradients w.r.t. their `cum` inputs), lengths or null
//   out: dmemory += , dpm += , dpq [B][128], dprev_out [B][T] (gradient w.r.t. THIS step's prev input), gsum += gradient
//        w.r.t. this step's cum input, and the parameter gradients (atomically accumulated): dWld [128][32],
//        dWloc [32][2][31], dv [128], dvb, dscale
//   scratch per utterance: loc [T][32], dloc [T][32], dw [T], g [T][62]
__global__ void __launch_bounds__(256) attention_bwd_kernel(
    const float* __restrict__ w, const float* __restrict__ prev, const float* __restrict__ cum, const float* __restrict__ pq,
    const float* __restrict__ pm, const float* __restrict__ memory, const float* __restrict__ dctx,
    const float* __restrict__ dw_ext, const float* __restrict__ dprev_next, float* __restrict__ gsum,
    const int64_t* __restrict__ lengths, const float* __restrict__ Wloc /* [32][2][31] */,
    const float* __restrict__ Wld /* [128][32] */, const float* __restrict__ v /* [128] */, const float* __restrict__ vb_p,
    const float* __restrict__ escale_p,
    float* __restrict__ dmemory, float* __restrict__ dpm, float* __restrict__ dpq, float* __restrict__ dprev_out,
    float* __restrict__ pacc /* [B][ATTN_PACC]: per-utterance parameter-gradient accumulators (dWld | dWloc | dv | dvb, dscale) */,
    float* __restrict__ scratch, int B, int T) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int len = lengths ? (int)lengths[b] : T;
    const float vb = vb_p[0], escale = escale_p[0];
    __shared__ float dctx_s[EMB];
    __shared__ float wloc_s[LOC_F * 2 * LOC_K];
    __shared__ float wld_s[ATT_DIM * (LOC_F + 1)];  // row stride 33: conflict-free both along a (lf) and along f (dloc)
    __shared__ float loc_s[8][LOC_F];
    __shared__ float v_s[ATT_DIM], pq_s[ATT_DIM], dpq_s[ATT_DIM], dv_s[ATT_DIM];
    __shared__ float red_s[256];
    __shared__ float tile_du[8][ATT_DIM];          // du of 8 positions at a time
    __shared__ float tile_de[8];
    float* dWld = pacc + (size_t)b * ATTN_PACC;            // only this CTA ever touches slice b: plain read-modify-write
    float* dWloc = dWld + ATT_DIM * LOC_F;
    float* dv = dWloc + LOC_F * 2 * LOC_K;
    float* dvb_scale = dv + ATT_DIM;
    float dwld_r[ATT_DIM * LOC_F / 256];                   // this thread's 16 entries of dWld, summed over all positions
#pragma unroll
    for (int i = 0; i < ATT_DIM * LOC_F / 256; ++i) dwld_r[i] = 0.f;
    float* loc = scratch + (size_t)b * T * (2 * LOC_F + 1 + 2 * LOC_K);
    float* dloc = loc + (size_t)T * LOC_F;
    float* dwv = dloc + (size_t)T * LOC_F;
    const float* wb = w + (size_t)b * T;
    const float* pb = prev ? prev + (size_t)b * T : nullptr;
    const float* cb = cum + (size_t)b * T;

    for (int i = tid; i < EMB; i += 256) dctx_s[i] = dctx[(size_t)b * EMB + i];
    for (int i = tid; i < LOC_F * 2 * LOC_K; i += 256) wloc_s[i] = Wloc[i];
    for (int i = tid; i < ATT_DIM * LOC_F; i += 256) wld_s[(i / LOC_F) * (LOC_F + 1) + (i % LOC_F)] = Wld[i];
    if (tid < ATT_DIM) { v_s[tid] = v[tid]; pq_s[tid] = pq[(size_t)b * ATT_DIM + tid]; dpq_s[tid] = 0.f; dv_s[tid] = 0.f; }
    __syncthreads();

    // (a) dw[tau] = dctx . memory[tau] + external terms ; dmemory[tau] += w[tau] * dctx ; s = sum_tau w dw
    float s_part = 0.f;
    for (int tau = tid >> 5; tau < T; tau += 8) {          // one warp per position
        const int lane = tid & 31;
        const float* mrow = memory + ((size_t)b * T + tau) * EMB;
        float* dmrow = dmemory + ((size_t)b * T + tau) * EMB;
        const float wt = wb[tau];
        float acc = 0.f;
        float mv[EMB / 32], dv_[EMB / 32];
#pragma unroll
        for (int q = 0; q < EMB / 32; ++q) {              // all 32 loads of the row pair in flight
            mv[q] = mrow[lane + 32 * q];
            dv_[q] = dmrow[lane + 32 * q];
        }
#pragma unroll
        for (int q = 0; q < EMB / 32; ++q) {
            acc = fmaf(dctx_s[lane + 32 * q], mv[q], acc);
            dmrow[lane + 32 * q] = dv_[q] + wt * dctx_s[lane + 32 * q];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            float d = acc + (dw_ext ? dw_ext[(size_t)b * T + tau] : 0.f) + (dprev_next ? dprev_next[(size_t)b * T + tau] : 0.f) +
                      gsum[(size_t)b * T + tau];
            dwv[tau] = d;
            s_part += wt * d;
        }
    }
    red_s[tid] = s_part;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red_s[tid] += red_s[tid + o];
        __syncthreads();
    }
    const float s_all = red_s[0];
    __syncthreads();

    // (b) location conv recomputed: loc[tau][f] = sum_c sum_k Wloc[f][c][k] * in_c[tau + k - 15]
    for (int i = tid; i < T * LOC_F; i += 256) {
        const int tau = i / LOC_F, f = i % LOC_F;
        float acc = 0.f;
        for (int k = 0; k < LOC_K; ++k) {
            const int tt = tau + k - (LOC_K - 1) / 2;
            if (tt >= 0 && tt < T) {
                acc = fmaf(wloc_s[(f * 2 + 0) * LOC_K + k], pb ? pb[tt] : 0.f, acc);
                acc = fmaf(wloc_s[(f * 2 + 1) * LOC_K + k], cb[tt], acc);
            }
        }
        loc[i] = acc;
    }
    __syncthreads();

    // (c) per position: de -> du[a], dpq, dv, dvb, dscale, dpm ; 8 positions per pass, thread = (position p8, 32 dims)
    float dvb_part = 0.f, dscale_part = 0.f;
    for (int t0 = 0; t0 < T; t0 += 8) {
        const int p8 = tid >> 5, lane = tid & 31, tau = t0 + p8;
        loc_s[p8][lane] = tau < T ? loc[tau * LOC_F + lane] : 0.f;
        __syncwarp();
        float e_raw_part = 0.f;
        float th[4], lfv[4];
        if (tau < T) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int a = lane + 32 * q;
                float lf = 0.f;
                for (int f = 0; f < LOC_F; ++f) lf = fmaf(wld_s[a * (LOC_F + 1) + f], loc_s[p8][f], lf);
                lfv[q] = lf;
                th[q] = tanhf(pq_s[a] + pm[((size_t)b * T + tau) * ATT_DIM + a] + lf);
                e_raw_part = fmaf(v_s[a], th[q], e_raw_part);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e_raw_part += __shfl_xor_sync(0xffffffffu, e_raw_part, o);
        float de_raw = 0.f;
        if (tau < T) {
            const float e_raw = e_raw_part + vb;
            // softmax backward; masked positions (energy filled with -1e9, src/model.py:132) pass no gradient
            const float de = tau < len ? wb[tau] * (dwv[tau] - s_all) : 0.f;
            de_raw = de * escale;
            if (lane == 0) { dscale_part += de * e_raw; dvb_part += de_raw; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int a = lane + 32 * q;
                const float du = de_raw * v_s[a] * (1.f - th[q] * th[q]);
                tile_du[p8][a] = du;
                dpm[((size_t)b * T + tau) * ATT_DIM + a] += du;
                atomicAdd(&dv_s[a], de_raw * th[q]);
                atomicAdd(&dpq_s[a], du);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) tile_du[p8][lane + 32 * q] = 0.f;
        }
        if (lane == 0) tile_de[p8] = de_raw;
        __syncthreads();
        // dloc[tau][f] = sum_a du[a] * Wld[a][f] ; dWld[a][f] += du[tau][a] * loc[tau][f]
        {
            const int f = tid & 31, pp = tid >> 5, tt = t0 + pp;
            if (tt < T) {
                float acc = 0.f;
                for (int a = 0; a < ATT_DIM; ++a) acc = fmaf(tile_du[pp][a], wld_s[a * (LOC_F + 1) + f], acc);
                dloc[tt * LOC_F + f] = acc;
            }
        }
#pragma unroll
        for (int q = 0; q < ATT_DIM * LOC_F / 256; ++q) {
            const int i = tid + q * 256, a = i / LOC_F, f = i % LOC_F;
            float acc = 0.f;
            for (int pp = 0; pp < 8; ++pp)
                acc = fmaf(tile_du[pp][a], loc_s[pp][f], acc);        // rows past T hold zeros in both tiles
            dwld_r[q] += acc;
        }
        __syncthreads();
        (void)lfv;
    }
    // block sums of dvb / dscale
    red_s[tid] = dvb_part;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (tid < o) red_s[tid] += red_s[tid + o]; __syncthreads(); }
    if (tid == 0) dvb_scale[0] += red_s[0];
    __syncthreads();
    red_s[tid] = dscale_part;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (tid < o) red_s[tid] += red_s[tid + o]; __syncthreads(); }
    if (tid == 0) dvb_scale[1] += red_s[0];
    if (tid < ATT_DIM) {
        dpq[(size_t)b * ATT_DIM + tid] = dpq_s[tid];
        dv[tid] += dv_s[tid];
    }
#pragma unroll
    for (int q = 0; q < ATT_DIM * LOC_F / 256; ++q) dWld[tid + q * 256] += dwld_r[q];
    __syncthreads();

    // (d) location conv backward: dWloc[f][c][k] += sum_tau dloc[tau][f] * in_c[tau + k - 15]
    for (int i = tid; i < LOC_F * 2 * LOC_K; i += 256) {
        const int k = i % LOC_K, c = (i / LOC_K) % 2, f = i / (2 * LOC_K);
        float acc = 0.f;
        for (int tau = 0; tau < T; ++tau) {
            const int tt = tau + k - (LOC_K - 1) / 2;
            if (tt >= 0 && tt < T) acc = fmaf(dloc[tau * LOC_F + f], c == 0 ? (pb ? pb[tt] : 0.f) : cb[tt], acc);
        }
        dWloc[i] += acc;
    }
    //     d in_c[tt] = sum_k g[tt - k + 15][c][k],  g[tau][c][k] = sum_f dloc[tau][f] * Wloc[f][c][k]  (two parallel passes)
    float* gck = dwv + T;                                      // [T][62]
    for (int i = tid; i < T * 2 * LOC_K; i += 256) {
        const int ck = i % (2 * LOC_K), tau = i / (2 * LOC_K);
        float acc = 0.f;
#pragma unroll 8
        for (int f = 0; f < LOC_F; ++f) acc = fmaf(dloc[tau * LOC_F + f], wloc_s[f * 2 * LOC_K + ck], acc);
        gck[i] = acc;
    }
    __syncthreads();
    for (int i = tid; i < 2 * T; i += 256) {
        const int c = i / T, tt = i % T;
        float acc = 0.f;
        for (int k = 0; k < LOC_K; ++k) {
            const int tau = tt - k + (LOC_K - 1) / 2;
            if (tau >= 0 && tau < T) acc += gck[tau * 2 * LOC_K + c * LOC_K + k];
        }
        if (c == 0) dprev_out[(size_t)b * T + tt] = acc;
        else gsum[(size_t)b * T + tt] += acc;
    }
}

// ---- attention backward of ONE decode step, second version: one CTA of 512 threads per utterance, every intermediate
// (attention windows, location features and their gradients) in SHARED memory, per-thread register accumulators for the
// parameter gradients, no atomics.  The memory gradient  dmemory[b][tau][:] = sum_t w_t[tau] * dctx_t[:]  is NOT accumulated
// here (it was a 2 x 19.6 MB read-modify-write per step at B=64): the kernel stores dctx_t into dctx_all[t] and the host
// contracts  align^T . dctx_all  once per utterance after the walk.  T <= ATTN_V2_MAX_T (shared-memory budget).
constexpr int ATTN_V2_NT = 512, ATTN_V2_NW = ATTN_V2_NT / 32, ATTN_V2_MAX_T = 320;
constexpr int LOC_LD = LOC_F + 4;                       // row stride of loc / dloc / Wld tiles: 16-byte rows, 9 x 16 B apart
constexpr int WLOC_LD = 2 * LOC_K + 1;                  // row stride of the location-conv weights (odd)
__host__ __device__ constexpr int attn_v2_smem_floats(int T) {
    return EMB + LOC_F * WLOC_LD + ATT_DIM * LOC_LD + 4 * ATT_DIM + 2 * (T + LOC_K - 1) + 2 * T + 2 * T * LOC_LD +
           2 * ATTN_V2_NW * ATT_DIM + 2 * ATTN_V2_NW * ATT_DIM + ATTN_V2_NT;
}
__global__ void __launch_bounds__(ATTN_V2_NT) attention_bwd_v2_kernel(
    const float* __restrict__ w, const float* __restrict__ prev, const float* __restrict__ cum, const float* __restrict__ pq,
    const float* __restrict__ pm, const float* __restrict__ memory, float* __restrict__ dctx,
    const float* __restrict__ dw_ext, const float* __restrict__ dprev_next, float* __restrict__ gsum,
    const int64_t* __restrict__ lengths, const float* __restrict__ Wloc, const float* __restrict__ Wld,
    const float* __restrict__ v, const float* __restrict__ vb_p, const float* __restrict__ escale_p,
    float* __restrict__ dctx_out /* [B][512]: this step's slice of dctx_all */, float* __restrict__ dpm,
    float* __restrict__ dpq, float* __restrict__ dprev_out, float* __restrict__ pacc, int B, int T) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int len = lengths ? (int)lengths[b] : T;
    const float vb = vb_p[0], escale = escale_p[0];
    constexpr int HALO = (LOC_K - 1) / 2;
    const int WL = T + LOC_K - 1;
    float* dctx_s = sm;
    float* wloc_s = dctx_s + EMB;
    float* wld_s = wloc_s + LOC_F * WLOC_LD;
    float* v_s = wld_s + ATT_DIM * LOC_LD;
    float* pq_s = v_s + ATT_DIM;
    float* win_p = pq_s + ATT_DIM + 2 * ATT_DIM;       // (two spare rows kept for alignment of the layout)
    float* win_c = win_p + WL;
    float* w_s = win_c + WL;
    float* dw_s = w_s + T;
    float* loc_s = dw_s + T;
    float* dloc_s = loc_s + T * LOC_LD;
    float* tile_du = dloc_s + T * LOC_LD;              // [2 NW][128]: du of the 32 positions of a pass
    float* accx = tile_du + 2 * ATTN_V2_NW * ATT_DIM;  // [2][NW][128]: dv / dpq partials of the warps
    float* red_s = accx + 2 * ATTN_V2_NW * ATT_DIM;    // [NT]
    float* dWld = pacc + (size_t)b * ATTN_PACC;        // only this CTA ever touches slice b: plain read-modify-write
    float* dWloc = dWld + ATT_DIM * LOC_F;
    float* dv = dWloc + LOC_F * 2 * LOC_K;
    float* dvb_scale = dv + ATT_DIM;
    const float* wb = w + (size_t)b * T;
    const float* pb = prev ? prev + (size_t)b * T : nullptr;
    const float* cb = cum + (size_t)b * T;

    for (int i = tid; i < EMB; i += ATTN_V2_NT) {
        const float d = dctx[(size_t)b * EMB + i];
        dctx_s[i] = d;
        dctx_out[(size_t)b * EMB + i] = d;
        dctx[(size_t)b * EMB + i] = 0.f;      // consumed: the attention cell's data gradient accumulates d ctx_{t-1} onto it
    }
    for (int i = tid; i < LOC_F * 2 * LOC_K; i += ATTN_V2_NT) wloc_s[(i / (2 * LOC_K)) * WLOC_LD + i % (2 * LOC_K)] = Wloc[i];
    for (int i = tid; i < ATT_DIM * LOC_F; i += ATTN_V2_NT) wld_s[(i / LOC_F) * LOC_LD + (i % LOC_F)] = Wld[i];
    if (tid < ATT_DIM) { v_s[tid] = v[tid]; pq_s[tid] = pq[(size_t)b * ATT_DIM + tid]; }
    for (int i = tid; i < WL; i += ATTN_V2_NT) {
        const int t = i - HALO;
        const bool ok = t >= 0 && t < T;
        win_p[i] = ok && pb ? pb[t] : 0.f;
        win_c[i] = ok ? cb[t] : 0.f;
    }
    for (int i = tid; i < T; i += ATTN_V2_NT) w_s[i] = wb[i];
    __syncthreads();

    // (a) dw[tau] = dctx . memory[tau] + external terms ; s = sum_tau w dw.  One warp per position, two positions in flight.
    float s_part = 0.f;
    for (int tau0 = 2 * warp; tau0 < T; tau0 += 2 * ATTN_V2_NW) {
        float4 m[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int tau = tau0 + r < T ? tau0 + r : tau0;
            const float4* mrow = reinterpret_cast<const float4*>(memory + ((size_t)b * T + tau) * EMB);
#pragma unroll
            for (int q = 0; q < 4; ++q) m[r][q] = __ldg(mrow + lane + 32 * q);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 d = *reinterpret_cast<const float4*>(dctx_s + 4 * (lane + 32 * q));
                acc = fmaf(d.x, m[r][q].x, acc); acc = fmaf(d.y, m[r][q].y, acc);
                acc = fmaf(d.z, m[r][q].z, acc); acc = fmaf(d.w, m[r][q].w, acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            const int tau = tau0 + r;
            if (lane == 0 && tau < T) {
                const float d = acc + (dw_ext ? dw_ext[(size_t)b * T + tau] : 0.f) +
                                (dprev_next ? dprev_next[(size_t)b * T + tau] : 0.f) + gsum[(size_t)b * T + tau];
                dw_s[tau] = d;
                s_part += w_s[tau] * d;
            }
        }
    }
    red_s[tid] = s_part;
    // (b) location conv recomputed from the windows: loc[tau][f] = sum_c sum_k Wloc[f][c][k] * in_c[tau + k - 15]
    for (int i = tid; i < T * LOC_F; i += ATTN_V2_NT) {
        const int tau = i >> 5, f = i & 31;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < LOC_K; ++k) {
            acc = fmaf(wloc_s[f * WLOC_LD + k], win_p[tau + k], acc);
            acc = fmaf(wloc_s[f * WLOC_LD + LOC_K + k], win_c[tau + k], acc);
        }
        loc_s[tau * LOC_LD + f] = acc;
    }
    __syncthreads();
    for (int o = ATTN_V2_NT / 2; o > 0; o >>= 1) {
        if (tid < o) red_s[tid] += red_s[tid + o];
        __syncthreads();
    }
    const float s_all = red_s[0];
    __syncthreads();

    // (c) per position: de -> du, dpm, partial dv / dpq / dvb / dscale; dloc[tau][f] = sum_a du[a] Wld[a][f];
    //     dWld[a][f] += du[tau][a] loc[tau][f].  A pass covers 32 positions, TWO per warp (lane = 4 attention dims), so every
    //     128-bit read of Wld feeds both; all shared-memory reads are 128 bits wide (the loop is shared-memory-bandwidth bound).
    float dvb_part = 0.f, dscale_part = 0.f;
    float dv_r[4] = {0.f, 0.f, 0.f, 0.f}, dpq_r[4] = {0.f, 0.f, 0.f, 0.f};
    float dwld_r[ATT_DIM * LOC_F / ATTN_V2_NT];            // entries a = 4 (warp + 16 g) + j, f = lane
#pragma unroll
    for (int i = 0; i < ATT_DIM * LOC_F / ATTN_V2_NT; ++i) dwld_r[i] = 0.f;
    for (int t0 = 0; t0 < T; t0 += 2 * ATTN_V2_NW) {
        int tau[2];
        bool ok[2];
        float* du_row[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            tau[r] = t0 + r * ATTN_V2_NW + warp;
            ok[r] = tau[r] < T;
            if (!ok[r]) tau[r] = T - 1;                    // loads stay in range; results are discarded
            du_row[r] = tile_du + (r * ATTN_V2_NW + warp) * ATT_DIM;
        }
        float pmv[2][4], lf[2][4], th[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                pmv[r][q] = __ldg(pm + ((size_t)b * T + tau[r]) * ATT_DIM + lane + 32 * q);
                lf[r][q] = 0.f;
            }
        const float4* l0 = reinterpret_cast<const float4*>(loc_s + tau[0] * LOC_LD);
        const float4* l1 = reinterpret_cast<const float4*>(loc_s + tau[1] * LOC_LD);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4* wr = reinterpret_cast<const float4*>(wld_s + (lane + 32 * q) * LOC_LD);
#pragma unroll
            for (int f4 = 0; f4 < LOC_F / 4; ++f4) {
                const float4 wv = wr[f4], x = l0[f4], y = l1[f4];
                lf[0][q] = fmaf(wv.x, x.x, lf[0][q]); lf[0][q] = fmaf(wv.y, x.y, lf[0][q]);
                lf[0][q] = fmaf(wv.z, x.z, lf[0][q]); lf[0][q] = fmaf(wv.w, x.w, lf[0][q]);
                lf[1][q] = fmaf(wv.x, y.x, lf[1][q]); lf[1][q] = fmaf(wv.y, y.y, lf[1][q]);
                lf[1][q] = fmaf(wv.z, y.z, lf[1][q]); lf[1][q] = fmaf(wv.w, y.w, lf[1][q]);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float e_raw_part = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int a = lane + 32 * q;
                th[r][q] = tanhf(pq_s[a] + pmv[r][q] + lf[r][q]);
                e_raw_part = fmaf(v_s[a], th[r][q], e_raw_part);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e_raw_part += __shfl_xor_sync(0xffffffffu, e_raw_part, o);
            const float e_raw = e_raw_part + vb;
            // softmax backward; masked positions (energy filled with -1e9, src/model.py:132) pass no gradient
            const float de = ok[r] && tau[r] < len ? w_s[tau[r]] * (dw_s[tau[r]] - s_all) : 0.f;
            const float de_raw = de * escale;
            if (lane == 0) { dscale_part += de * e_raw; dvb_part += de_raw; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int a = lane + 32 * q;
                const float du = de_raw * v_s[a] * (1.f - th[r][q] * th[r][q]);     // 0 for positions past T (de = 0)
                du_row[r][a] = du;
                if (ok[r]) {
                    float* dpm_p = dpm + ((size_t)b * T + tau[r]) * ATT_DIM + a;
                    *dpm_p += du;
                }
                dv_r[q] = fmaf(de_raw, th[r][q], dv_r[q]);
                dpq_r[q] += du;
            }
        }
        __syncwarp();
        {
            float acc0 = 0.f, acc1 = 0.f;
            const float4* d0 = reinterpret_cast<const float4*>(du_row[0]);
            const float4* d1 = reinterpret_cast<const float4*>(du_row[1]);
#pragma unroll 4
            for (int a4 = 0; a4 < ATT_DIM / 4; ++a4) {
                const float4 x = d0[a4], y = d1[a4];
                const float w0 = wld_s[(4 * a4 + 0) * LOC_LD + lane], w1 = wld_s[(4 * a4 + 1) * LOC_LD + lane],
                            w2 = wld_s[(4 * a4 + 2) * LOC_LD + lane], w3 = wld_s[(4 * a4 + 3) * LOC_LD + lane];
                acc0 = fmaf(x.x, w0, acc0); acc0 = fmaf(x.y, w1, acc0); acc0 = fmaf(x.z, w2, acc0); acc0 = fmaf(x.w, w3, acc0);
                acc1 = fmaf(y.x, w0, acc1); acc1 = fmaf(y.y, w1, acc1); acc1 = fmaf(y.z, w2, acc1); acc1 = fmaf(y.w, w3, acc1);
            }
            if (ok[0]) dloc_s[tau[0] * LOC_LD + lane] = acc0;
            if (ok[1]) dloc_s[tau[1] * LOC_LD + lane] = acc1;
        }
        __syncthreads();
#pragma unroll 4
        for (int pp = 0; pp < 2 * ATTN_V2_NW; ++pp) {
            const int tt = t0 + pp < T ? t0 + pp : T - 1;         // rows past T hold zeros in tile_du
            const float l = loc_s[tt * LOC_LD + lane];
            const float4* dr = reinterpret_cast<const float4*>(tile_du + pp * ATT_DIM);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float4 d = dr[warp + ATTN_V2_NW * g];
                dwld_r[4 * g + 0] = fmaf(d.x, l, dwld_r[4 * g + 0]);
                dwld_r[4 * g + 1] = fmaf(d.y, l, dwld_r[4 * g + 1]);
                dwld_r[4 * g + 2] = fmaf(d.z, l, dwld_r[4 * g + 2]);
                dwld_r[4 * g + 3] = fmaf(d.w, l, dwld_r[4 * g + 3]);
            }
        }
        __syncthreads();
    }
    // warp partials of dv / dpq -> block sums
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        accx[warp * ATT_DIM + lane + 32 * q] = dv_r[q];
        accx[(ATTN_V2_NW + warp) * ATT_DIM + lane + 32 * q] = dpq_r[q];
    }
    red_s[tid] = dvb_part;
    __syncthreads();
    if (tid < 2 * ATT_DIM) {
        const int which = tid / ATT_DIM, a = tid % ATT_DIM;
        float acc = 0.f;
#pragma unroll
        for (int ww = 0; ww < ATTN_V2_NW; ++ww) acc += accx[(which * ATTN_V2_NW + ww) * ATT_DIM + a];
        if (which == 0) dv[a] += acc;
        else dpq[(size_t)b * ATT_DIM + a] = acc;
    }
    // dvb / dscale are produced by lane 0 of every warp only
    if (tid == 0) {
        float acc = 0.f;
        for (int ww = 0; ww < ATTN_V2_NW; ++ww) acc += red_s[ww * 32];
        dvb_scale[0] += acc;
    }
    __syncthreads();
    red_s[tid] = dscale_part;
    __syncthreads();
    if (tid == 0) {
        float acc = 0.f;
        for (int ww = 0; ww < ATTN_V2_NW; ++ww) acc += red_s[ww * 32];
        dvb_scale[1] += acc;
    }
#pragma unroll
    for (int q = 0; q < ATT_DIM * LOC_F / ATTN_V2_NT; ++q)
        dWld[(4 * (warp + ATTN_V2_NW * (q >> 2)) + (q & 3)) * LOC_F + lane] += dwld_r[q];

    // (d) location conv backward: dWloc[f][c][k] += sum_tau dloc[tau][f] * in_c[tau + k - 15]
    for (int i = tid; i < LOC_F * 2 * LOC_K; i += ATTN_V2_NT) {
        const int k = i % LOC_K, c = (i / LOC_K) % 2, f = i / (2 * LOC_K);
        const float* win = c == 0 ? win_p : win_c;
        float acc = 0.f;
#pragma unroll 4
        for (int tau = 0; tau < T; ++tau) acc = fmaf(dloc_s[tau * LOC_LD + f], win[tau + k], acc);
        dWloc[i] += acc;
    }
    //     d in_c[tt] = sum_k sum_f dloc[tt - k + 15][f] * Wloc[f][c][k]
    for (int i = tid; i < 2 * T; i += ATTN_V2_NT) {
        const int c = i / T, tt = i % T;
        float acc = 0.f;
        for (int k = 0; k < LOC_K; ++k) {
            const int tau = tt - k + HALO;
            if (tau >= 0 && tau < T) {
                const float4* dl = reinterpret_cast<const float4*>(dloc_s + tau * LOC_LD);
                const float* wl = wloc_s + c * LOC_K + k;
#pragma unroll
                for (int f4 = 0; f4 < LOC_F / 4; ++f4) {
                    const float4 d = dl[f4];
                    acc = fmaf(d.x, wl[(4 * f4 + 0) * WLOC_LD], acc); acc = fmaf(d.y, wl[(4 * f4 + 1) * WLOC_LD], acc);
                    acc = fmaf(d.z, wl[(4 * f4 + 2) * WLOC_LD], acc); acc = fmaf(d.w, wl[(4 * f4 + 3) * WLOC_LD], acc);
                }
            }
        }
        if (c == 0) dprev_out[(size_t)b * T + tt] = acc;
        else gsum[(size_t)b * T + tt] += acc;
    }
}

// ---- third version: the same computation with the positions of an utterance split over a CLUSTER OF TWO CTAs (2 B = 128 CTAs
// instead of 64 on the 148 SMs; the kernel is bound by shared-memory bandwidth and dependent latency inside one SM, so two SMs per
// utterance nearly halve it).  CTA r owns positions [r Th, r Th + n).  What crosses the pair goes through DSMEM:
//   - the softmax-backward scalar  s = sum_tau w dw   (each CTA adds its partial into the peer, cluster.sync #1),
//   - the query gradient dpq (CTA 1 sends its 128 partial sums to CTA 0, which stores the total),
//   - 15 halo rows of dloc on each side of the cut for the transposed location convolution (read from the peer after
//     cluster.sync #2; cluster.sync #3 keeps the peer resident until they are read).
// Parameter-gradient partials go to accumulator slice 2 b + r (2 B slices, summed once after the walk).  T >= ATTN_V3_MIN_T.
constexpr int ATTN_V3_MIN_T = 48;
static_assert(2 * ATTN_V2_NT == RNN, "fused cell tail: one thread of the CTA pair per attention-LSTM unit");
constexpr int ATTN_V3_G_LD = 2 * LOC_K + 3;      // odd row stride of the g[tau][c][k] tile (location-input gradient)
static_assert(ATTN_V2_NT == LOC_F * 2 * 8 && 8 * 4 >= LOC_K, "dWloc mapping: thread = (feature, channel, group of 4 taps)");
static_assert(ATTN_V2_NT == 8 * 64 && 2 * LOC_K <= 64 && LOC_F == 32, "g mapping: thread = (row group, channel-tap)");
__host__ __device__ constexpr int attn_v3_half(int T) { return (((T + 1) >> 1) + 3) & ~3; }
__host__ __device__ constexpr int attn_v3_smem_floats(int T) {
    return EMB + LOC_F * WLOC_LD + ATT_DIM * LOC_LD + 5 * ATT_DIM + 2 * (attn_v3_half(T) + LOC_K - 1) + 2 * attn_v3_half(T) +
           attn_v3_half(T) * LOC_LD + (attn_v3_half(T) + LOC_K - 1) * LOC_LD + 2 * ATTN_V2_NW * ATT_DIM + 2 * ATTN_V2_NW * ATT_DIM +
           ATTN_V2_NT + (attn_v3_half(T) + LOC_K - 1) * ATTN_V3_G_LD;
}
// phase times of CTA 0 (TIMED instantiation only, TACO_ATTN_TIMES=1; read with b200taco_attn_bwd_phase_ns, debug header)
__device__ unsigned long long g_attn_phase_ns[8];
#define ATTN_T(i)                                                            \
    if (TIMED && blockIdx.x == 0 && tid == 0) {                              \
        const unsigned long long now_ = global_timer_ns();                   \
        atomicAdd(&g_attn_phase_ns[i], now_ - t_prev_);                      \
        t_prev_ = now_;                                                      \
    }
template <bool TIMED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ATTN_V2_NT) attention_bwd_v3_kernel(
    const float* __restrict__ w, const float* __restrict__ prev, const float* __restrict__ cum, const float* __restrict__ pq,
    const float* __restrict__ pm, const float* __restrict__ memory, float* __restrict__ dctx,
    const float* __restrict__ dw_ext, const float* __restrict__ dprev_next, float* __restrict__ gsum,
    const int64_t* __restrict__ lengths, const float* __restrict__ Wloc, const float* __restrict__ Wld,
    const float* __restrict__ v, const float* __restrict__ vb_p, const float* __restrict__ escale_p,
    float* __restrict__ dctx_out, float* __restrict__ dpm, float* __restrict__ dpq, float* __restrict__ dprev_out,
    float* __restrict__ pacc /* [2 B][ATTN_PACC] */, const float* __restrict__ dctx_extra /* [B][ld_extra], nullable: added to dctx */,
    int ld_extra, const AttnCellTail tail, int B, int T) {
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    // "both CTAs are running" (required before the first DSMEM access): arrive now, wait just before that access
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    const int b = blockIdx.x >> 1, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int len = lengths ? (int)lengths[b] : T;
    const float vb = vb_p[0], escale = escale_p[0];
    constexpr int HALO = (LOC_K - 1) / 2;
    const int Th = attn_v3_half(T);
    const int p0 = rank * Th, n = rank == 0 ? Th : T - Th;          // own positions [p0, p0 + n); both >= HALO (T >= 48)
    const int WLh = Th + LOC_K - 1;                                  // layout strides use Th on both CTAs (peer addressing)
    float* dctx_s = sm;
    float* wloc_s = dctx_s + EMB;
    float* wld_s = wloc_s + LOC_F * WLOC_LD;
    float* v_s = wld_s + ATT_DIM * LOC_LD;
    float* pq_s = v_s + ATT_DIM;
    float* xch_dpq = pq_s + ATT_DIM;                   // written by the peer (CTA 1 -> CTA 0)
    float* own_dpq = xch_dpq + ATT_DIM;
    float* xch_s = own_dpq + ATT_DIM;                  // [0] written by the peer; rest of the row spare
    float* win_p = xch_s + ATT_DIM;                    // [n + 30]: conv inputs at positions p0 - 15 ...
    float* win_c = win_p + WLh;
    float* w_s = win_c + WLh;
    float* dw_s = w_s + Th;
    float* loc_s = dw_s + Th;
    float* dloc_s = loc_s + Th * LOC_LD;               // [n + 30] rows: row i = position p0 - 15 + i (halo rows from the peer / zero)
    float* tile_du = dloc_s + WLh * LOC_LD;
    float* accx = tile_du + 2 * ATTN_V2_NW * ATT_DIM;
    float* red_s = accx + 2 * ATTN_V2_NW * ATT_DIM;
    float* g_s = red_s + ATTN_V2_NT;                   // [n + 30][ATTN_V3_G_LD]
    float* dWld = pacc + (size_t)blockIdx.x * ATTN_PACC;
    float* dWloc = dWld + ATT_DIM * LOC_F;
    float* dv = dWloc + LOC_F * 2 * LOC_K;
    float* dvb_scale = dv + ATT_DIM;
    const float* wb = w + (size_t)b * T;
    const float* pb = prev ? prev + (size_t)b * T : nullptr;
    const float* cb = cum + (size_t)b * T;
    unsigned long long t_prev_ = TIMED ? global_timer_ns() : 0ull;

    // ---- everything that depends only on weights and saved forward activations comes first: under programmatic dependent
    //      launch this part (and the location features (b)) overlaps the preceding kernel of the walk
    for (int i = tid; i < LOC_F * 2 * LOC_K; i += ATTN_V2_NT) wloc_s[(i / (2 * LOC_K)) * WLOC_LD + i % (2 * LOC_K)] = Wloc[i];
    for (int i = tid; i < ATT_DIM * LOC_F; i += ATTN_V2_NT) wld_s[(i / LOC_F) * LOC_LD + (i % LOC_F)] = Wld[i];
    if (tid < ATT_DIM) { v_s[tid] = v[tid]; pq_s[tid] = pq[(size_t)b * ATT_DIM + tid]; }
    for (int i = tid; i < n + LOC_K - 1; i += ATTN_V2_NT) {
        const int t = p0 - HALO + i;
        const bool ok = t >= 0 && t < T;
        win_p[i] = ok && pb ? pb[t] : 0.f;
        win_c[i] = ok ? cb[t] : 0.f;
    }
    for (int i = tid; i < n; i += ATTN_V2_NT) w_s[i] = wb[p0 + i];
    // halo rows of dloc that no CTA owns (positions < 0 and >= T) are zero
    for (int i = tid; i < HALO * LOC_LD; i += ATTN_V2_NT) dloc_s[(rank == 0 ? 0 : (HALO + n) * LOC_LD) + i] = 0.f;
    __syncthreads();
    ATTN_T(0)                                                        // loads
    // (b) location features of the own positions: lane = feature, a warp takes 8 consecutive positions at a time and slides an
    //     8-entry window over the conv input, so one weight and ONE new input feed 8 FMAs (the first version read two shared-memory
    //     words per FMA and was bound by that: 4.1 us)
    for (int j0 = warp * 8; j0 < n; j0 += ATTN_V2_NW * 8) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            const float* win = (c == 0 ? win_p : win_c) + j0;     // reads past the own window (tail block) land in the next array; discarded
            const float* wl = wloc_s + lane * WLOC_LD + c * LOC_K;
            float x[8];
#pragma unroll
            for (int i = 0; i < 7; ++i) x[i] = win[i];
#pragma unroll
            for (int kk = 0; kk < LOC_K; ++kk) {
                x[(kk + 7) & 7] = win[kk + 7];
                const float wv = wl[kk];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(wv, x[(kk + i) & 7], acc[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (j0 + i < n) loc_s[(j0 + i) * LOC_LD + lane] = acc[i];
    }
    ATTN_T(1)                                                        // (b) location features
    pdl_wait();                                                      // ---- from here on: gradients produced by the preceding kernels
    pdl_trigger();
    for (int i = tid; i < EMB; i += ATTN_V2_NT) {
        const float d = dctx[(size_t)b * EMB + i] + (dctx_extra ? dctx_extra[(size_t)b * ld_extra + i] : 0.f);
        dctx_s[i] = d;
        if (rank == 0) dctx_out[(size_t)b * EMB + i] = d;
    }
    __syncthreads();
    // (a) dw of the own positions, partial s
    float s_part = 0.f;
    for (int j0 = 2 * warp; j0 < n; j0 += 2 * ATTN_V2_NW) {
        float4 m[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = j0 + r < n ? j0 + r : j0;
            const float4* mrow = reinterpret_cast<const float4*>(memory + ((size_t)b * T + p0 + j) * EMB);
#pragma unroll
            for (int q = 0; q < 4; ++q) m[r][q] = __ldg(mrow + lane + 32 * q);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 d = *reinterpret_cast<const float4*>(dctx_s + 4 * (lane + 32 * q));
                acc = fmaf(d.x, m[r][q].x, acc); acc = fmaf(d.y, m[r][q].y, acc);
                acc = fmaf(d.z, m[r][q].z, acc); acc = fmaf(d.w, m[r][q].w, acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            const int j = j0 + r;
            if (lane == 0 && j < n) {
                const size_t g = (size_t)b * T + p0 + j;
                const float d = acc + (dw_ext ? dw_ext[g] : 0.f) + (dprev_next ? dprev_next[g] : 0.f) + gsum[g];
                dw_s[j] = d;
                s_part += w_s[j] * d;
            }
        }
    }
    red_s[tid] = s_part;
    __syncthreads();
    ATTN_T(2)                                                        // waiting for the predecessor + dctx + (a)
    for (int o = ATTN_V2_NT / 2; o > 0; o >>= 1) {
        if (tid < o) red_s[tid] += red_s[tid + o];
        __syncthreads();
    }
    const float s_own = red_s[0];
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    if (tid == 0) *cluster.map_shared_rank(xch_s, rank ^ 1) = s_own;
    cluster.sync();                                                  // #1 (also orders the reads of red_s[0] before its reuse)
    const float s_all = rank == 0 ? s_own + xch_s[0] : xch_s[0] + s_own;
    if (rank == 0)                                                   // consumed (the peer read it before it arrived at #1)
        for (int i = tid; i < EMB; i += ATTN_V2_NT) dctx[(size_t)b * EMB + i] = 0.f;

    ATTN_T(3)                                                        // block reduction + cluster.sync #1
    // (c) per position, exactly as in the second version, over the own positions
    float dvb_part = 0.f, dscale_part = 0.f;
    float dv_r[4] = {0.f, 0.f, 0.f, 0.f}, dpq_r[4] = {0.f, 0.f, 0.f, 0.f};
    float dwld_r[ATT_DIM * LOC_F / ATTN_V2_NT];
#pragma unroll
    for (int i = 0; i < ATT_DIM * LOC_F / ATTN_V2_NT; ++i) dwld_r[i] = 0.f;
    // the processed-memory rows of a pass are fetched one pass ahead (an L2 round trip per pass was exposed otherwise)
    float pmn[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q)
            pmn[r][q] = __ldg(pm + ((size_t)b * T + p0 + min(r * ATTN_V2_NW + warp, n - 1)) * ATT_DIM + lane + 32 * q);
    for (int t0 = 0; t0 < n; t0 += 2 * ATTN_V2_NW) {
        int jj[2];
        bool ok[2];
        float* du_row[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            jj[r] = t0 + r * ATTN_V2_NW + warp;
            ok[r] = jj[r] < n;
            if (!ok[r]) jj[r] = n - 1;
            du_row[r] = tile_du + (r * ATTN_V2_NW + warp) * ATT_DIM;
        }
        float pmv[2][4], lf[2][4], th[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                pmv[r][q] = pmn[r][q];
                lf[r][q] = 0.f;
            }
        if (t0 + 2 * ATTN_V2_NW < n) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    pmn[r][q] = __ldg(pm + ((size_t)b * T + p0 + min(t0 + 2 * ATTN_V2_NW + r * ATTN_V2_NW + warp, n - 1)) * ATT_DIM + lane + 32 * q);
        }
        const float4* l0 = reinterpret_cast<const float4*>(loc_s + jj[0] * LOC_LD);
        const float4* l1 = reinterpret_cast<const float4*>(loc_s + jj[1] * LOC_LD);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4* wr = reinterpret_cast<const float4*>(wld_s + (lane + 32 * q) * LOC_LD);
#pragma unroll
            for (int f4 = 0; f4 < LOC_F / 4; ++f4) {
                const float4 wv = wr[f4], x = l0[f4], y = l1[f4];
                lf[0][q] = fmaf(wv.x, x.x, lf[0][q]); lf[0][q] = fmaf(wv.y, x.y, lf[0][q]);
                lf[0][q] = fmaf(wv.z, x.z, lf[0][q]); lf[0][q] = fmaf(wv.w, x.w, lf[0][q]);
                lf[1][q] = fmaf(wv.x, y.x, lf[1][q]); lf[1][q] = fmaf(wv.y, y.y, lf[1][q]);
                lf[1][q] = fmaf(wv.z, y.z, lf[1][q]); lf[1][q] = fmaf(wv.w, y.w, lf[1][q]);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float e_raw_part = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int a = lane + 32 * q;
                th[r][q] = tanhf(pq_s[a] + pmv[r][q] + lf[r][q]);
                e_raw_part = fmaf(v_s[a], th[r][q], e_raw_part);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e_raw_part += __shfl_xor_sync(0xffffffffu, e_raw_part, o);
            const float e_raw = e_raw_part + vb;
            const float de = ok[r] && p0 + jj[r] < len ? w_s[jj[r]] * (dw_s[jj[r]] - s_all) : 0.f;
            const float de_raw = de * escale;
            if (lane == 0) { dscale_part += de * e_raw; dvb_part += de_raw; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int a = lane + 32 * q;
                const float du = de_raw * v_s[a] * (1.f - th[r][q] * th[r][q]);
                du_row[r][a] = du;
                // exactly one add per element and launch (deterministic); a reduction does not make the warp wait for the old value
                if (ok[r]) atomicAdd(dpm + ((size_t)b * T + p0 + jj[r]) * ATT_DIM + a, du);
                dv_r[q] = fmaf(de_raw, th[r][q], dv_r[q]);
                dpq_r[q] += du;
            }
        }
        __syncwarp();
        {
            float acc0 = 0.f, acc1 = 0.f;
            const float4* d0 = reinterpret_cast<const float4*>(du_row[0]);
            const float4* d1 = reinterpret_cast<const float4*>(du_row[1]);
#pragma unroll 4
            for (int a4 = 0; a4 < ATT_DIM / 4; ++a4) {
                const float4 x = d0[a4], y = d1[a4];
                const float w0 = wld_s[(4 * a4 + 0) * LOC_LD + lane], w1 = wld_s[(4 * a4 + 1) * LOC_LD + lane],
                            w2 = wld_s[(4 * a4 + 2) * LOC_LD + lane], w3 = wld_s[(4 * a4 + 3) * LOC_LD + lane];
                acc0 = fmaf(x.x, w0, acc0); acc0 = fmaf(x.y, w1, acc0); acc0 = fmaf(x.z, w2, acc0); acc0 = fmaf(x.w, w3, acc0);
                acc1 = fmaf(y.x, w0, acc1); acc1 = fmaf(y.y, w1, acc1); acc1 = fmaf(y.z, w2, acc1); acc1 = fmaf(y.w, w3, acc1);
            }
            if (ok[0]) dloc_s[(HALO + jj[0]) * LOC_LD + lane] = acc0;
            if (ok[1]) dloc_s[(HALO + jj[1]) * LOC_LD + lane] = acc1;
        }
        __syncthreads();
#pragma unroll 4
        for (int pp = 0; pp < 2 * ATTN_V2_NW; ++pp) {
            const int j = t0 + pp < n ? t0 + pp : n - 1;          // rows past n hold zeros in tile_du
            const float l = loc_s[j * LOC_LD + lane];
            const float4* dr = reinterpret_cast<const float4*>(tile_du + pp * ATT_DIM);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float4 d = dr[warp + ATTN_V2_NW * g];
                dwld_r[4 * g + 0] = fmaf(d.x, l, dwld_r[4 * g + 0]);
                dwld_r[4 * g + 1] = fmaf(d.y, l, dwld_r[4 * g + 1]);
                dwld_r[4 * g + 2] = fmaf(d.z, l, dwld_r[4 * g + 2]);
                dwld_r[4 * g + 3] = fmaf(d.w, l, dwld_r[4 * g + 3]);
            }
        }
        __syncthreads();
    }
    ATTN_T(4)                                                        // (c)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        accx[warp * ATT_DIM + lane + 32 * q] = dv_r[q];
        accx[(ATTN_V2_NW + warp) * ATT_DIM + lane + 32 * q] = dpq_r[q];
    }
    red_s[tid] = dvb_part;
    __syncthreads();
    if (tid < 2 * ATT_DIM) {
        const int which = tid / ATT_DIM, a = tid % ATT_DIM;
        float acc = 0.f;
#pragma unroll
        for (int ww = 0; ww < ATTN_V2_NW; ++ww) acc += accx[(which * ATTN_V2_NW + ww) * ATT_DIM + a];
        if (which == 0) dv[a] += acc;
        else {
            own_dpq[a] = acc;
            if (rank == 1 || tail.Wq != nullptr) cluster.map_shared_rank(xch_dpq, rank ^ 1)[a] = acc;
        }
    }
    if (tid == 0) {
        float acc = 0.f;
        for (int ww = 0; ww < ATTN_V2_NW; ++ww) acc += red_s[ww * 32];
        dvb_scale[0] += acc;
    }
    __syncthreads();
    red_s[tid] = dscale_part;
    __syncthreads();
    if (tid == 0) {
        float acc = 0.f;
        for (int ww = 0; ww < ATTN_V2_NW; ++ww) acc += red_s[ww * 32];
        dvb_scale[1] += acc;
    }
#pragma unroll
    for (int q = 0; q < ATT_DIM * LOC_F / ATTN_V2_NT; ++q)
        dWld[(4 * (warp + ATTN_V2_NW * (q >> 2)) + (q & 3)) * LOC_F + lane] += dwld_r[q];

    cluster.sync();                                                  // #2: dloc of both halves complete, dpq partial delivered
    if (rank == 0 && tid < ATT_DIM) dpq[(size_t)b * ATT_DIM + tid] = own_dpq[tid] + xch_dpq[tid];
    if (tail.Wq != nullptr) {
        // the same sum on both CTAs (CTA 0's partial first), then this thread's unit of the query^T product and of the cell
        const int u = rank * ATTN_V2_NT + tid;
        float q = 0.f;
        const float* wq = tail.Wq + u;
#pragma unroll 8
        for (int a = 0; a < ATT_DIM; ++a) {
            const float d = rank == 0 ? own_dpq[a] + xch_dpq[a] : xch_dpq[a] + own_dpq[a];
            q = fmaf(d, __ldg(wq + (size_t)a * RNN), q);
        }
        const int i = b * RNN + u;
        const float dh = (tail.d_ah[i] + q) * (tail.mask ? tail.mask[i] * tail.scale : 1.f);
        tail.d_ah[i] = 0.f;
        lstm_cell_bwd_unit(tail.gates, tail.c, tail.c_prev, dh, tail.dc, tail.dz, tail.dzp, b, u);
        if (tail.gates2 != nullptr) {
            const float dh2 = (tail.dh2[i] + (tail.dh2_extra ? tail.dh2_extra[(size_t)b * tail.ld2 + u] : 0.f)) *
                              (tail.mask2 ? tail.mask2[i] * tail.scale : 1.f);
            tail.dh2[i] = 0.f;
            lstm_cell_bwd_unit(tail.gates2, tail.c2, tail.c_prev2, dh2, tail.dc2, tail.dz2, tail.dzp2, b, u);
        }
    }
    {   // the 15 rows next to the cut come from the peer: CTA 0 needs the peer's first rows, CTA 1 the last rows of CTA 0
        const float* peer = cluster.map_shared_rank(dloc_s, rank ^ 1);
        const int src_row = rank == 0 ? HALO : Th, dst_row = rank == 0 ? HALO + n : 0;
        for (int i = tid; i < HALO * LOC_F; i += ATTN_V2_NT) {
            const int rr = i >> 5, f = i & 31;
            dloc_s[(dst_row + rr) * LOC_LD + f] = peer[(src_row + rr) * LOC_LD + f];
        }
    }
    cluster.sync();                                                  // #3: halo rows in place; the peer may exit
    ATTN_T(5)                                                        // reductions, cluster.sync #2, halo copy, cluster.sync #3

    // (d) location conv backward over the own positions: dWloc[f][c][k] += sum_j dloc[j][f] * in_c[j + k - 15].
    //     thread = (f, c, 4 consecutive taps): per position one dloc value and one NEW window value feed 4 FMAs
    {
        const int f = tid >> 4, c = (tid >> 3) & 1, k0 = (tid & 7) * 4;
        const float* win = (c == 0 ? win_p : win_c) + k0;            // tap 31 does not exist: its column reads one word past the window
        const float* dl = dloc_s + HALO * LOC_LD + f;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        float w0 = win[0], w1 = win[1], w2 = win[2];
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float w3 = win[j + 3];
            const float d = dl[j * LOC_LD];
            a0 = fmaf(d, w0, a0); a1 = fmaf(d, w1, a1); a2 = fmaf(d, w2, a2); a3 = fmaf(d, w3, a3);
            w0 = w1; w1 = w2; w2 = w3;
        }
        float* o = dWloc + f * 2 * LOC_K + c * LOC_K + k0;
        o[0] += a0; o[1] += a1; o[2] += a2;
        if (k0 + 3 < LOC_K) o[3] += a3;
    }
    ATTN_T(6)                                                        // dWloc
    //     d in_c[tt] = sum_k g[tt - k + 15][c][k],  g[tau][c][k] = sum_f dloc[tau][f] * Wloc[f][c][k]: the filter column of a
    //     thread stays in registers while it walks over the rows (halo rows included; rows outside [0, T) are zero)
    {
        const int ck = tid & 63, grp = tid >> 6;
        if (ck < 2 * LOC_K) {
            float wreg[LOC_F];
#pragma unroll
            for (int f = 0; f < LOC_F; ++f) wreg[f] = wloc_s[f * WLOC_LD + ck];
            for (int r = grp; r < n + 2 * HALO; r += 8) {
                const float4* dl = reinterpret_cast<const float4*>(dloc_s + r * LOC_LD);
                float acc = 0.f;
#pragma unroll
                for (int f4 = 0; f4 < LOC_F / 4; ++f4) {
                    const float4 d = dl[f4];
                    acc = fmaf(d.x, wreg[4 * f4 + 0], acc); acc = fmaf(d.y, wreg[4 * f4 + 1], acc);
                    acc = fmaf(d.z, wreg[4 * f4 + 2], acc); acc = fmaf(d.w, wreg[4 * f4 + 3], acc);
                }
                g_s[r * ATTN_V3_G_LD + ck] = acc;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * n; i += ATTN_V2_NT) {
        const int c = i / n, j = i % n;
        const float* gp = g_s + (j + 2 * HALO) * ATTN_V3_G_LD + c * LOC_K;     // row of tap 0; tap k sits k rows up, k words right
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < LOC_K; ++k) acc += gp[k - k * ATTN_V3_G_LD];
        if (c == 0) dprev_out[(size_t)b * T + p0 + j] = acc;
        else gsum[(size_t)b * T + p0 + j] += acc;
    }
    __syncthreads();
    ATTN_T(7)                                                        // d in_c
}
#undef ATTN_T

// Deferred context gradient (src/model.py:147): dmemory[b][tau][d] = sum_s w[s][b][tau] * dctx[s][b][d], one 64 x 64 tile of
// (tau, d) per CTA, all S steps reduced in the CTA (plain stores, no atomics).  w (S,B,T), dctx (S,B,512), dmemory (B,T,512).
__global__ void __launch_bounds__(256) ctx_dmemory_kernel(const float* __restrict__ w, const float* __restrict__ dctx,
                                                          float* __restrict__ dmemory, int S, int B, int T) {
    __shared__ __align__(16) float Ws[16][64 + 4];
    __shared__ __align__(16) float Ds[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int t0 = blockIdx.x * 64, d0 = blockIdx.y * 64, b = blockIdx.z;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 4, lc = (tid & 15) * 4;
    for (int sb = 0; sb < S; sb += 16) {
        const int sidx = sb + lr;
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f), dv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sidx < S) {
            const float* wrow = w + ((size_t)sidx * B + b) * T;          // rows are not 16-byte aligned in general
            const int t = t0 + lc;
            if (t < T) wv.x = wrow[t];
            if (t + 1 < T) wv.y = wrow[t + 1];
            if (t + 2 < T) wv.z = wrow[t + 2];
            if (t + 3 < T) wv.w = wrow[t + 3];
            dv = *reinterpret_cast<const float4*>(dctx + ((size_t)sidx * B + b) * EMB + d0 + lc);
        }
        *reinterpret_cast<float4*>(&Ws[lr][lc]) = wv;
        *reinterpret_cast<float4*>(&Ds[lr][lc]) = dv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&Ws[k][4 * ty]);
            const float4 c = *reinterpret_cast<const float4*>(&Ds[k][4 * tx]);
            const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], cv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + 4 * ty + i;
        if (t < T)
            *reinterpret_cast<float4*>(dmemory + ((size_t)b * T + t) * EMB + d0 + 4 * tx) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// out[b][n] += sum_k X[b][k] * W[k][n] for b < B <= 64: one 64 x 64 output tile per (blockIdx.x, K split blockIdx.y),
// partial sums added atomically (out must hold the values to accumulate onto, e.g. zeros)
// Column ranges of the product may go to different destinations (one launch for all contractions that share X, e.g. the
// three data gradients of an LSTM cell): segment i covers columns [n_end[i-1], n_end[i]), boundaries are multiples of 64.
struct SkinnyOut {
    float* out[3];
    int ldo[3];
    int n_end[3];
};
__global__ void __launch_bounds__(256) skinny_gemm_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                          int ldw, const SkinnyOut so, int B, int K, int N, int k_per) {
    __shared__ __align__(16) float Xs[16][64 + 4];
    __shared__ __align__(16) float Ws[16][64];
    pdl_wait();
    pdl_trigger();          // only now: the NEXT kernel of the chain may start its set-up while this one works (run-ahead of one kernel)
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * 64, k0 = blockIdx.y * k_per, k1 = min(K, k0 + k_per);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int xm = tid >> 2, xq = tid & 3;              // X loader: row xm, k 4*xq..
    const int wk = tid >> 4, wq = tid & 15;             // W loader: k row wk, cols 4*wq..
    // register double buffering: the loads of tile kb+16 are in flight while tile kb is multiplied
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), wv = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](int kb) {
        xv = make_float4(0.f, 0.f, 0.f, 0.f);
        wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xm < B) xv = *reinterpret_cast<const float4*>(X + (size_t)xm * ldx + kb + 4 * xq);
        if (n0 + 4 * wq < N) wv = __ldg(reinterpret_cast<const float4*>(W + (size_t)(kb + wk) * ldw + n0 + 4 * wq));
    };
    if (k0 < k1) fetch(k0);
    for (int kb = k0; kb < k1; kb += 16) {
        Xs[4 * xq + 0][xm] = xv.x; Xs[4 * xq + 1][xm] = xv.y; Xs[4 * xq + 2][xm] = xv.z; Xs[4 * xq + 3][xm] = xv.w;
        *reinterpret_cast<float4*>(&Ws[wk][4 * wq]) = wv;
        __syncthreads();
        if (kb + 16 < k1) fetch(kb + 16);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&Xs[kk][4 * ty]);
            const float4 b = *reinterpret_cast<const float4*>(&Ws[kk][4 * tx]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int seg = n0 < so.n_end[0] ? 0 : (n0 < so.n_end[1] ? 1 : 2);
    const int nb = seg == 0 ? 0 : so.n_end[seg - 1];
    float* out = so.out[seg];
    const int ldo = so.ldo[seg];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = 4 * ty + i;
        if (b >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 4 * tx + j;
            if (n < N) atomicAdd(out + (size_t)b * ldo + (n - nb), acc[i][j]);
        }
    }
}

// dst[c][r] = src[r][c] for r < R, c < C (row strides ldd / lds)
__global__ void transpose_block_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, int R, int C) {
    const size_t total = (size_t)R * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = i % R, c = i / R;
        dst[(size_t)c * ldd + r] = src[(size_t)r * lds + c];
    }
}
// dout[s][b][96]: columns 0..79 = d_coarse[b][s][:], 80 = d_gate[b][s], rest 0
__global__ void assemble_dout_kernel(const float* __restrict__ d_coarse, const float* __restrict__ d_gate, float* __restrict__ dout,
                                     int B, int S) {
    const size_t total = (size_t)S * B * PROJ_PAD;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int o = i % PROJ_PAD, b = (i / PROJ_PAD) % B, s = i / ((size_t)PROJ_PAD * B);
        float v = 0.f;
        if (o < N_MELS) v = d_coarse[((size_t)b * S + s) * N_MELS + o];
        else if (o == N_MELS) v = d_gate[(size_t)b * S + s];
        dout[i] = v;
    }
}
// Wpg[96][1536]: rows 0..79 = linear_projection.weight, row 80 = gate_layer.weight, rest 0
__global__ void assemble_wpg_kernel(const float* __restrict__ wp, const float* __restrict__ wg, float* __restrict__ dst) {
    const int total = PROJ_PAD * (RNN + EMB);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % (RNN + EMB), o = i / (RNN + EMB);
        dst[i] = o < N_MELS ? wp[(size_t)o * (RNN + EMB) + c] : (o == N_MELS ? wg[c] : 0.f);
    }
}

}  // namespace

// out (+)= X . W ; X [B][K] (ldx), W [K][N] (ldw), out [B][N] (ldo), B <= 64, K % 16 == 0, N % 4 == 0
// out_i[b][n - n_begin_i] (+)= sum_k X[b][k] * W[k][n] for the column segments [n_begin_i, n_end_i) of one [K][N] matrix:
// up to three destinations, each accumulated onto (acc[i]) or overwritten; n_end[nseg-1] = N.
int skinny_gemm_multi(const float* X, int ldx, const float* W, int ldw, int B, int K, int nseg, float* const* outs, const int* ldos,
                      const int* n_ends, const bool* acc, cudaStream_t st) {
    const int N = n_ends[nseg - 1];
    TACO_REQUIRE(B <= 64 && K % 16 == 0 && N % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && nseg >= 1 && nseg <= 3,
                 B200TACO_ERR_INVALID, "skinny_gemm: unsupported shape B %d K %d N %d", B, K, N);
    SkinnyOut so{};
    for (int i = 0; i < 3; ++i) {
        const int j = i < nseg ? i : nseg - 1;
        so.out[i] = outs[j]; so.ldo[i] = ldos[j]; so.n_end[i] = i < nseg ? n_ends[i] : N;
        if (i < nseg) {
            const int nb = i == 0 ? 0 : n_ends[i - 1];
            TACO_REQUIRE(i == nseg - 1 || n_ends[i] % 64 == 0, B200TACO_ERR_INVALID, "skinny_gemm: segment boundary %d", n_ends[i]);
            if (!acc[i])
                TACO_CUDA(cudaMemset2DAsync(outs[i], (size_t)ldos[i] * sizeof(float), 0, (size_t)(n_ends[i] - nb) * sizeof(float), B, st));
        }
    }
    const int ntiles = (N + 63) / 64;
    static const int want = [] { const char* e = getenv("TACO_SKINNY_CTAS"); return e ? atoi(e) : 592; }();    // measured at B=64: 128 -> 592 CTAs = -20 ms per backward
    int ks = 1;
    while (ntiles * ks < want && K / (ks * 2) >= 64) ks *= 2;
    const int k_per = ((K + ks - 1) / ks + 15) / 16 * 16;
    TACO_CUDA(launch_pdl(PDL_SKINNY, skinny_gemm_kernel, dim3(ntiles, (K + k_per - 1) / k_per), dim3(256), 0, st, X, ldx, W, ldw, so, B, K, N, k_per));
    TACO_LAUNCHED();
    return B200TACO_OK;
}
int skinny_gemm(const float* X, int ldx, const float* W, int ldw, float* out, int ldo, int B, int K, int N, bool accumulate,
                cudaStream_t st) {
    float* outs[1] = {out};
    const int ldos[1] = {ldo}, n_ends[1] = {N};
    const bool acc[1] = {accumulate};
    return skinny_gemm_multi(X, ldx, W, ldw, B, K, 1, outs, ldos, n_ends, acc, st);
}

int transpose_block(float* dst, int ldd, const float* src, int lds, int R, int C, cudaStream_t st) {
    transpose_block_kernel<<<blocks_for((size_t)R * C), 256, 0, st>>>(dst, ldd, src, lds, R, C);
    TACO_LAUNCHED();
    return B200TACO_OK;
}
int assemble_dout(const float* d_coarse, const float* d_gate, float* dout, int B, int S, cudaStream_t st) {
    assemble_dout_kernel<<<blocks_for((size_t)S * B * PROJ_PAD), 256, 0, st>>>(d_coarse, d_gate, dout, B, S);
    TACO_LAUNCHED();
    return B200TACO_OK;
}
int assemble_wpg(const float* wp, const float* wg, float* dst, cudaStream_t st) {
    assemble_wpg_kernel<<<256, 256, 0, st>>>(wp, wg, dst);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

int lstm_cell_bwd(const float* gates, const float* c, const float* c_prev, float* dh_post, const float* dh_extra, int ld_extra, const float* mask, float scale,
                  float* dc, float* dz, void* dz_tiles, int B, cudaStream_t st) {
    TACO_REQUIRE(dz_tiles == nullptr || B <= 64, B200TACO_ERR_INVALID, "lstm_cell_bwd: tile image needs B <= 64");
    TACO_CUDA(launch_pdl(PDL_CELL, lstm_cell_bwd_kernel, dim3((B * RNN + 255) / 256), dim3(256), 0, st, gates, c, c_prev, dh_post, dh_extra, ld_extra, mask,
                         scale, dc, dz, static_cast<__nv_bfloat16*>(dz_tiles), B));
    TACO_LAUNCHED();
    return B200TACO_OK;
}
int attn_cumsum(const float* align, float* cum, int S, int BT, cudaStream_t st) {
    attn_cumsum_kernel<<<(BT + 255) / 256, 256, 0, st>>>(align, cum, S, BT);
    TACO_LAUNCHED();
    return B200TACO_OK;
}
int train_axpy(float* dst, int ld_dst, const float* a, int ld_a, int n, int cols, bool accumulate, cudaStream_t st) {
    axpy_kernel<<<blocks_for((size_t)n * cols), 256, 0, st>>>(dst, a, n, ld_dst, ld_a, cols, accumulate ? 1 : 0);
    TACO_LAUNCHED();
    return B200TACO_OK;
}
int relu_drop_bwd(const float* dy, const float* y, const float* mask, float scale, float* dx, size_t n, cudaStream_t st) {
    relu_drop_bwd_kernel<<<blocks_for(n), 256, 0, st>>>(dy, y, mask, scale, dx, n);
    TACO_LAUNCHED();
    return B200TACO_OK;
}
int relu_drop_fwd(float* y, const float* mask, float scale, size_t n, cudaStream_t st) {
    relu_drop_fwd_kernel<<<blocks_for(n), 256, 0, st>>>(y, mask, scale, n);
    TACO_LAUNCHED();
    return B200TACO_OK;
}
size_t attention_bwd_scratch_floats(int B, int T) { return (size_t)B * T * (2 * LOC_F + 1 + 2 * LOC_K); }
// the shared-memory version is used whenever the utterance fits (T <= 320); it defers the memory gradient (dctx_t is stored
// to dctx_out, see attention_bwd_v2_kernel) - attention_bwd_defers_dmemory tells the caller which contract applies
bool attention_bwd_defers_dmemory(int T) {
    static const bool off = [] { const char* e = getenv("TACO_ATTN_BWD_V1"); return e && e[0] == '1'; }();
    return !off && T <= ATTN_V2_MAX_T;
}
// true when attention_bwd runs the cluster kernel for this T, i.e. when the caller may hand it the query^T / cell tail
bool attention_bwd_fuses_cell(int T) {
    static const bool no_v3 = [] { const char* e = getenv("TACO_ATTN_BWD_V2"); return e && e[0] == '1'; }();
    static const bool no_fuse = [] { const char* e = getenv("TACO_ATTN_FUSE_CELL"); return e && e[0] == '0'; }();
    return attention_bwd_defers_dmemory(T) && T >= ATTN_V3_MIN_T && !no_v3 && !no_fuse;
}
int attn_bwd_phase_ns(unsigned long long* out8, int reset) {
    if (out8 != nullptr) TACO_CUDA(cudaMemcpyFromSymbol(out8, g_attn_phase_ns, sizeof(g_attn_phase_ns)));
    if (reset) {
        const unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        TACO_CUDA(cudaMemcpyToSymbol(g_attn_phase_ns, z, sizeof(z)));
    }
    return B200TACO_OK;
}
int ctx_dmemory(const float* w, const float* dctx_all, float* dmemory, int S, int B, int T, cudaStream_t st) {
    ctx_dmemory_kernel<<<dim3((T + 63) / 64, EMB / 64, B), 256, 0, st>>>(w, dctx_all, dmemory, S, B, T);
    TACO_LAUNCHED();
    return B200TACO_OK;
}
int attention_bwd(const float* w, const float* prev, const float* cum, const float* pq, const float* pm, const float* memory,
                  float* dctx, const float* dw_ext, const float* dprev_next, float* gsum, const int64_t* lengths,
                  const float* Wloc, const float* Wld, const float* v, const float* vb, const float* escale, float* dmemory, float* dpm,
                  float* dpq, float* dprev_out, float* pacc, float* scratch, float* dctx_out, const float* dctx_extra, int ld_extra,
                  const AttnCellTail* tail, int B, int T, cudaStream_t st) {
    static const bool no_v3 = [] { const char* e = getenv("TACO_ATTN_BWD_V2"); return e && e[0] == '1'; }();
    if (attention_bwd_defers_dmemory(T) && T >= ATTN_V3_MIN_T && !no_v3) {
        const int smem = attn_v3_smem_floats(T) * (int)sizeof(float);
        static const bool timed = [] { const char* e = getenv("TACO_ATTN_TIMES"); return e && e[0] == '1'; }();
        static bool configured3 = false;
        if (!configured3) {
            TACO_CUDA(cudaFuncSetAttribute(attention_bwd_v3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           attn_v3_smem_floats(ATTN_V2_MAX_T) * (int)sizeof(float)));
            TACO_CUDA(cudaFuncSetAttribute(attention_bwd_v3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           attn_v3_smem_floats(ATTN_V2_MAX_T) * (int)sizeof(float)));
            configured3 = true;
        }
        auto kern = timed ? attention_bwd_v3_kernel<true> : attention_bwd_v3_kernel<false>;
        TACO_CUDA(launch_pdl(PDL_ATTENTION, kern, dim3(2 * B), dim3(ATTN_V2_NT), smem, st, w, prev, cum, pq, pm, memory, dctx, dw_ext, dprev_next, gsum, lengths, Wloc,
                             Wld, v, vb, escale, dctx_out, dpm, dpq, dprev_out, pacc, dctx_extra, ld_extra, tail ? *tail : AttnCellTail{}, B, T));
        TACO_LAUNCHED();
        return B200TACO_OK;
    }
    TACO_REQUIRE(tail == nullptr, B200TACO_ERR_STATE, "attention_bwd: the fused cell tail needs the cluster kernel (attention_bwd_fuses_cell)");
    // the older kernels take the extra addend as a separate pass
    if (dctx_extra != nullptr)
        if (int e = train_axpy(dctx, EMB, dctx_extra, ld_extra, B, EMB, true, st)) return e;
    if (attention_bwd_defers_dmemory(T)) {
        const int smem = attn_v2_smem_floats(T) * (int)sizeof(float);
        static bool configured = false;
        if (!configured) {
            TACO_CUDA(cudaFuncSetAttribute(attention_bwd_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           attn_v2_smem_floats(ATTN_V2_MAX_T) * (int)sizeof(float)));
            configured = true;
        }
        attention_bwd_v2_kernel<<<B, ATTN_V2_NT, smem, st>>>(w, prev, cum, pq, pm, memory, dctx, dw_ext, dprev_next, gsum, lengths,
                                                            Wloc, Wld, v, vb, escale, dctx_out, dpm, dpq, dprev_out, pacc, B, T);
        TACO_LAUNCHED();
        return B200TACO_OK;
    }
    attention_bwd_kernel<<<B, 256, 0, st>>>(w, prev, cum, pq, pm, memory, dctx, dw_ext, dprev_next, gsum, lengths, Wloc, Wld, v, vb,
                                            escale, dmemory, dpm, dpq, dprev_out, pacc, scratch, B, T);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // namespace b200taco
