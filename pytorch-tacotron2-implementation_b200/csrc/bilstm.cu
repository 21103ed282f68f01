// Encoder BiLSTM recurrence (reference src/model.py:42-48, :75: nn.LSTM, 1 layer,
// bidirectional, hidden 256 per direction, zero initial state, full padded length).
//
// The input projections x.W_ih^T + b_ih + b_hh of all positions are one dense GEMM
// (dense.cu) done beforehand; what is left is strictly serial in T.  One thread-block
// cluster of 8 CTAs runs one direction for a group of 8 utterances:
//   * CTA `rank` owns hidden units [32*rank, 32*rank+32) = 128 gate rows.  Its slice of
//     W_hh (128 x 256 fp32 = 128 KB) lives in REGISTERS for the whole sequence: thread
//     (row, k-half) keeps 128 weights, so a step reads no weights from memory at all.
//   * h_{t-1} (256 x 8) sits in shared memory of every CTA, double buffered; after the
//     cell update each CTA scatters its 32 x 8 new h values into all 8 CTAs through
//     distributed shared memory, then one cluster barrier ends the step.
//
// Ragged batches (lengths != null, b200taco_encoder_fwd_ragged): every utterance is run over ITS OWN positions [0, len) - the
// reverse direction starts at len - 1 with the zero state - which is exactly its exact-length solo run (the reference's
// inference encodes one unpadded utterance at a time, inference.py:66); a cluster walks max(len) of its 8 utterances and rows
// behind an utterance's length are not written (the caller zeroes them).
#include <cooperative_groups.h>

#include "handle.h"

namespace cg = cooperative_groups;

namespace b200taco {

namespace {

constexpr int CL = 8;          // CTAs per cluster
constexpr int BG = 8;          // utterances per cluster
constexpr int UPC = ENC_H / CL;    // 32 units per CTA
constexpr int ROWS = 4 * UPC;      // 128 gate rows per CTA
constexpr int KH = ENC_H / 2;      // 128 k per thread

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256, 1)
    bilstm_kernel(const float* __restrict__ whh /* [2][8][128][256] */, const float* __restrict__ xproj /* [B*T][2048] */,
                  float* __restrict__ out /* [B][T][512] */, const int32_t* __restrict__ lengths /* [B] or null */, int B, int T) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = cluster.block_rank();
    const int cid = blockIdx.x / CL;            // cluster index: direction-major
    const int ngroups = (B + BG - 1) / BG;
    const int dir = cid / ngroups, b0 = (cid % ngroups) * BG;
    const int tid = threadIdx.x;

    __shared__ __align__(16) float h_s[2][BG][ENC_H];      // h_{t-1}, [buffer][utterance][k]
    __shared__ float z_s[ROWS][BG + 1];

    // weights -> registers: thread (row = tid/2, half = tid%2)
    const int row = tid >> 1, half = tid & 1;
    float w[KH];
    {
        const float4* src =
            reinterpret_cast<const float4*>(whh + (((size_t)dir * CL + rank) * ROWS + row) * ENC_H + half * KH);
#pragma unroll
        for (int i = 0; i < KH / 4; ++i) {
            const float4 v = __ldg(src + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
    for (int i = tid; i < 2 * BG * ENC_H; i += 256) (&h_s[0][0][0])[i] = 0.f;
    // cell role: thread (unit = tid/8, utterance = tid%8)
    const int cu = tid >> 3, cb = tid & 7;
    const bool cvalid = b0 + cb < B;
    const int len_b = cvalid ? (lengths ? min(T, max(0, (int)lengths[b0 + cb])) : T) : 0;
    int steps = T;
    if (lengths) {                              // the same value in every CTA of the cluster
        steps = 0;
        for (int b = b0; b < min(B, b0 + BG); ++b) steps = max(steps, min(T, max(0, (int)lengths[b])));
    }
    float c_state = 0.f;
    cluster.sync();

    for (int step = 0; step < steps; ++step) {
        const int t = dir ? len_b - 1 - step : step;
        const bool active = step < len_b;       // an utterance that ran out of positions idles (its h column is read by nobody else)
        const int cur = step & 1, nxt = cur ^ 1;
        // input projection of this position (independent of h: issue the loads first)
        float xin[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
            const float* xp = xproj + ((size_t)(b0 + cb) * T + t) * (8 * ENC_H) + dir * 4 * ENC_H + rank * UPC + cu;
#pragma unroll
            for (int g = 0; g < 4; ++g) xin[g] = __ldg(xp + g * ENC_H);
        }
        // recurrent mat-vec for 8 utterances: acc[b] = sum_k w[k] * h[b][half*128 + k]
        float acc[BG];
#pragma unroll
        for (int b = 0; b < BG; ++b) acc[b] = 0.f;
#pragma unroll
        for (int k = 0; k < KH; k += 4) {
#pragma unroll
            for (int b = 0; b < BG; ++b) {
                const float4 hv = *reinterpret_cast<const float4*>(&h_s[cur][b][half * KH + k]);
                acc[b] = fmaf(w[k], hv.x, acc[b]);
                acc[b] = fmaf(w[k + 1], hv.y, acc[b]);
                acc[b] = fmaf(w[k + 2], hv.z, acc[b]);
                acc[b] = fmaf(w[k + 3], hv.w, acc[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < BG; ++b) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], 1);
        if (half == 0) {
#pragma unroll
            for (int b = 0; b < BG; ++b) z_s[row][b] = acc[b];
        }
        __syncthreads();
        // LSTM cell, gate order i,f,g,o (rows of a unit are unit*4 + gate)
        float h = 0.f;
        if (active) {
            const float zi = z_s[cu * 4 + 0][cb] + xin[0];
            const float zf = z_s[cu * 4 + 1][cb] + xin[1];
            const float zg = z_s[cu * 4 + 2][cb] + xin[2];
            const float zo = z_s[cu * 4 + 3][cb] + xin[3];
            c_state = sigmoid_acc(zf) * c_state + sigmoid_acc(zi) * tanhf(zg);
            h = sigmoid_acc(zo) * tanhf(c_state);
            out[((size_t)(b0 + cb) * T + t) * EMB + dir * ENC_H + rank * UPC + cu] = h;
        }
        // scatter h into every CTA of the cluster (distributed shared memory)
        float* local = &h_s[nxt][cb][rank * UPC + cu];
#pragma unroll
        for (int r = 0; r < CL; ++r) *cluster.map_shared_rank(local, r) = h;
        cluster.sync();
    }
}

}  // namespace

int bilstm_recurrence(const b200taco_handle* h, const float* xproj, float* memory, int B, int T, cudaStream_t stream,
                      const int32_t* lengths) {
    const int ngroups = (B + BG - 1) / BG;
    bilstm_kernel<<<2 * ngroups * CL, 256, 0, stream>>>(h->enc_lstm_whh, xproj, memory, lengths, B, T);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // namespace b200taco
