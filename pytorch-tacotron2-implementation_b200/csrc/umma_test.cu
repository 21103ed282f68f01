// Bring-up / unit-test kernel for the tcgen05 helpers (umma.cuh): D[128 x 32] = A[128 x K] . B[32 x K]^T with
// bf16 operands and fp32 accumulation in TMEM, one CTA.  Exercised by tests/test_gpu_umma.py; it pins the
// descriptor conventions the decoder's tensor-core path relies on.
#include "handle.h"
#include "umma.cuh"

namespace b200taco {
namespace {

constexpr int TM = 128, TN = 32, TK = 64, KG = TK / 8;

__global__ void __launch_bounds__(128, 1)
umma_test_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D, int K,
                 int variant) {
    __shared__ __align__(1024) __nv_bfloat16 a_s[TM * TK];
    __shared__ __align__(1024) __nv_bfloat16 b_s[TN * TK];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, 32);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = umma::instr_desc_bf16_f32(TM, TN);
    uint32_t phase = 0;
    for (int k0 = 0; k0 < K; k0 += TK) {
        for (int i = tid; i < TM * TK; i += 128) {
            const int r = i / TK, k = i % TK;
            a_s[umma::tile_offset(r, k, KG)] = A[(size_t)r * K + k0 + k];
        }
        for (int i = tid; i < TN * TK; i += 128) {
            const int r = i / TK, k = i % TK;
            b_s[umma::tile_offset(r, k, KG)] = Bm[(size_t)r * K + k0 + k];
        }
        fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t lbo = (variant & 1) ? KG * 128 : 128, sbo = (variant & 1) ? 128 : KG * 128;
#pragma unroll
            for (int ks = 0; ks < TK / 16; ++ks) {
                const uint64_t ad = umma::smem_desc(smem_u32(a_s) + ks * 256, lbo, sbo);
                const uint64_t bd = umma::smem_desc(smem_u32(b_s) + ks * 256, lbo, sbo);
                umma::mma_bf16(tmem, ad, bd, idesc, k0 > 0 || ks > 0);
            }
            umma::commit(&bar);
        }
        {
            const uint64_t t0 = global_timer_ns();
            while (!mbar_try_wait(&bar, phase)) {
                if (global_timer_ns() - t0 > 2000000000ull) __trap();
            }
        }
        phase ^= 1;
        __syncthreads();
    }
    umma::fence_after_sync();
    float v[32];
    umma::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int n = 0; n < TN; ++n) D[(size_t)tid * TN + n] = v[n];
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 32);
}

}  // namespace
}  // namespace b200taco

extern "C" B200TACO_API int32_t b200taco_debug_umma_gemm(const void* d_A, const void* d_B, float* d_D, int32_t K,
                                                         int32_t variant, void* stream) {
    using namespace b200taco;
    TACO_REQUIRE(d_A && d_B && d_D && K >= 64 && K % 64 == 0, B200TACO_ERR_INVALID, "debug_umma_gemm: bad arguments");
    umma_test_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(d_A), static_cast<const __nv_bfloat16*>(d_B), d_D, K, variant);
    TACO_LAUNCHED();
    return B200TACO_OK;
}
