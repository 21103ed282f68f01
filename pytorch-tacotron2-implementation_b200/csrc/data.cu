// Batch assembly and hand-off layouts on the device.
//
//  * collate_kernel: TextMelCollate.__call__ of the reference (src/data_utils.py:45-84): the ragged (text ids, mel) samples
//    of a batch, delivered as two flat arrays plus offsets, become text_padded (B, Lmax) int64 and mel_padded (B, 80, Smax)
//    fp32, zero padded, row i holding sample order[i] (the text-length sort of src/data_utils.py:54-56).  One launch instead
//    of a Python loop of B slice assignments on the host followed by four host->device copies of padded tensors.
//  * vocoder_layout_kernel: (B, S, C) frames -> (B, C, S_out) channel major (inference.py:71 hands mel.transpose(1, 2) to the
//    vocoder), zero from each utterance's own length on (the per-utterance trim of train.py:71-79).  The tensor-core postnet
//    writes this layout from its epilogue (dense_tc.cu, ChannelMajorOut); this kernel serves the CUDA-core path and callers
//    that already hold a (B, S, C) tensor.
#include "handle.h"

namespace b200taco {
namespace {

constexpr int CT = 256;

// grid (B, chunks): CTA (i, c) pads text row i (chunk 0 only) and a slice of the 80 x Smax mel plane of row i
__global__ void __launch_bounds__(CT) collate_kernel(const int64_t* __restrict__ text_flat, const int64_t* __restrict__ text_off,
                                                     const float* __restrict__ mel_flat, const int64_t* __restrict__ mel_off,
                                                     const int32_t* __restrict__ order, int Lmax, int Smax, int n_mels,
                                                     int64_t* __restrict__ text_padded, float* __restrict__ mel_padded) {
    const int i = blockIdx.x, src = order[i], tid = threadIdx.x;
    if (blockIdx.y == 0) {
        const int64_t t0 = text_off[src];
        const int L = (int)(text_off[src + 1] - t0);
        for (int p = tid; p < Lmax; p += CT) text_padded[(size_t)i * Lmax + p] = p < L ? text_flat[t0 + p] : 0;   // :76
    }
    const int64_t m0 = mel_off[src] * n_mels;                      // sample layout (80, S_src), contiguous
    const int S = (int)(mel_off[src + 1] - mel_off[src]);
    const long long plane = (long long)n_mels * Smax;
    for (long long e = (long long)blockIdx.y * CT + tid; e < plane; e += (long long)gridDim.y * CT) {
        const int m = (int)(e / Smax), s = (int)(e % Smax);
        mel_padded[(size_t)i * plane + e] = s < S ? mel_flat[m0 + (int64_t)m * S + s] : 0.f;                      // :80
    }
}

// 32 x 32 tiles through shared memory: reads coalesced along C, writes coalesced along S
__global__ void __launch_bounds__(256) vocoder_layout_kernel(const float* __restrict__ mel, int ldm, const int32_t* __restrict__ lengths,
                                                             int S, int C, int S_out, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int len = lengths ? lengths[b] : S;
    for (int r = ty; r < 32; r += 8) {
        const int s = s0 + r, c = c0 + tx;
        tile[r][tx] = (s < S && c < C) ? mel[((size_t)b * S + s) * ldm + c] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, s = s0 + tx;
        if (c < C && s < S_out) out[((size_t)b * C + c) * S_out + s] = s < len ? tile[tx][r] : 0.f;
    }
}

__global__ void zero_tail_kernel(float* __restrict__ Y, int ldy, const int32_t* __restrict__ lengths, int S, int C) {
    const int b = blockIdx.y, len = lengths[b];
    const long long n = (long long)(S - len) * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int s = len + (int)(e / C), c = (int)(e % C);
        Y[((size_t)b * S + s) * ldy + c] = 0.f;
    }
}

}  // namespace

int zero_tail_rows(float* Y, int ldy, const int32_t* lengths, int B, int S, int C, cudaStream_t stream) {
    TACO_REQUIRE(B >= 1 && B <= 65535, B200TACO_ERR_INVALID, "zero_tail_rows: B %d", B);
    zero_tail_kernel<<<dim3(32, B), 256, 0, stream>>>(Y, ldy, lengths, S, C);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

int vocoder_layout(const float* mel, int ldm, const int32_t* lengths, int B, int S, int C, int S_out, float* out, cudaStream_t stream) {
    TACO_REQUIRE(B >= 1 && B <= 65535 && S >= 1 && C >= 1 && S_out >= 1, B200TACO_ERR_INVALID, "vocoder_layout: B %d S %d C %d S_out %d", B, S,
                 C, S_out);
    dim3 grid((S_out + 31) / 32, (C + 31) / 32, B);
    vocoder_layout_kernel<<<grid, 256, 0, stream>>>(mel, ldm, lengths, S, C, S_out, out);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // namespace b200taco

using namespace b200taco;

extern "C" {

B200TACO_API int32_t b200taco_collate(const int64_t* d_text_flat, const int64_t* d_text_offsets, const float* d_mel_flat,
                                      const int64_t* d_mel_offsets, const int32_t* d_order, int32_t B, int32_t Lmax, int32_t Smax,
                                      int64_t* d_text_padded, float* d_mel_padded, void* stream) {
    TACO_REQUIRE(d_text_flat && d_text_offsets && d_mel_flat && d_mel_offsets && d_order && d_text_padded && d_mel_padded,
                 B200TACO_ERR_INVALID, "collate: null pointer");
    TACO_REQUIRE(B >= 1 && B <= 65535 && Lmax >= 1 && Smax >= 1, B200TACO_ERR_INVALID, "collate: B %d, Lmax %d, Smax %d", B, Lmax, Smax);
    int chunks = (N_MELS * Smax + CT * 8 - 1) / (CT * 8);
    chunks = chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks);
    collate_kernel<<<dim3(B, chunks), CT, 0, static_cast<cudaStream_t>(stream)>>>(d_text_flat, d_text_offsets, d_mel_flat, d_mel_offsets,
                                                                                  d_order, Lmax, Smax, N_MELS, d_text_padded, d_mel_padded);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

B200TACO_API int32_t b200taco_vocoder_layout(const float* d_mel, const int32_t* d_lengths, int32_t B, int32_t S, int32_t S_out,
                                             float* d_out, void* stream) {
    TACO_REQUIRE(d_mel && d_out, B200TACO_ERR_INVALID, "vocoder_layout: null pointer");
    return vocoder_layout(d_mel, N_MELS, d_lengths, B, S, N_MELS, S_out, d_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
