// Guided-attention target of the reference's Tacotron2Loss (train.py:148-171): for every (utterance b, decoder step t) a
// Gaussian over the phoneme positions p < L_b centred on the diagonal position  e = min(int(t * L_b / S), L_b - 1), width
// sigma_b, normalised by (sum + 1e-8); zero for p >= L_b.  The reference builds it with a Python double loop over B x S
// (51,200 iterations at B=64, S=800: seconds per training step); here it is one CTA per (b, t).
#include "handle.h"

namespace b200taco {
namespace {

constexpr int GA_THREADS = 128;

__global__ void __launch_bounds__(GA_THREADS) guided_attention_target_kernel(const int64_t* __restrict__ text_lengths,
                                                                            const float* __restrict__ sigma,
                                                                            float* __restrict__ target, int S, int Tmax) {
    __shared__ float g_s[MAX_TEXT_LEN];
    __shared__ float red_s[GA_THREADS / 32];
    const int b = blockIdx.x / S, t = blockIdx.x % S, tid = threadIdx.x;
    long long L = text_lengths[b];
    L = L < 1 ? 1 : (L > Tmax ? Tmax : L);
    const long long e_ll = ((long long)t * L) / S;                       // int(t * Lb / num_steps), train.py:166
    const float e = (float)(e_ll < L - 1 ? e_ll : L - 1);
    const float sg = sigma[b];
    float part = 0.f;
    for (int p = tid; p < (int)L; p += GA_THREADS) {
        const float x = ((float)p - e) / sg;
        const float g = expf(-0.5f * (x * x));                           // train.py:167
        g_s[p] = g;
        part += g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red_s[tid >> 5] = part;
    __syncthreads();
    float total = 0.f;
#pragma unroll
    for (int i = 0; i < GA_THREADS / 32; ++i) total += red_s[i];
    const float denom = total + 1e-8f;                                   // train.py:168
    float* row = target + ((size_t)b * S + t) * Tmax;
    for (int p = tid; p < Tmax; p += GA_THREADS) row[p] = p < (int)L ? g_s[p] / denom : 0.f;
}

}  // namespace
}  // namespace b200taco

using namespace b200taco;

extern "C" {

B200TACO_API int32_t b200taco_guided_attention_target(const int64_t* d_text_lengths, const float* d_sigma, int32_t B, int32_t S,
                                                      int32_t Tmax, float* d_target, void* stream) {
    TACO_REQUIRE(d_text_lengths && d_sigma && d_target, B200TACO_ERR_INVALID, "guided_attention_target: null pointer");
    TACO_REQUIRE(B >= 1 && S >= 1 && Tmax >= 1 && Tmax <= MAX_TEXT_LEN, B200TACO_ERR_UNSUPPORTED,
                 "guided_attention_target: B %d, S %d, Tmax %d (max %d)", B, S, Tmax, MAX_TEXT_LEN);
    TACO_REQUIRE((long long)B * S < (1ll << 31), B200TACO_ERR_UNSUPPORTED, "guided_attention_target: B * S too large");
    guided_attention_target_kernel<<<B * S, GA_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(d_text_lengths, d_sigma, d_target,
                                                                                               S, Tmax);
    TACO_LAUNCHED();
    return B200TACO_OK;
}

}  // extern "C"
