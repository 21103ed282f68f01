/*
 * b200taco_debug.h -- measurement and bring-up aids of libb200taco.so.
 *
 * Exported by the same shared library, used by bench.py (the live L2 read-bandwidth probe), profiles/tools/ and
 * tests/test_gpu_umma.py; NOT part of the product ABI of b200taco.h (no reference interface stands behind them, and they may
 * change without an ABI version bump).
 */
#ifndef B200TACO_DEBUG_H
#define B200TACO_DEBUG_H

#include "b200taco.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Measurement aid used by bench.py for the roofline denominator (the L2
 * peak is not in MEASURED_PEAKS.json): all SMs repeatedly read a `bytes`-sized buffer
 * that fits in L2; returns GB/s through *gbps.  d_buf must hold `bytes` bytes.
 */
B200TACO_API int32_t b200taco_measure_l2_read(void* d_buf, size_t bytes, int32_t iters, float* gbps, void* stream);

/*
 * Measurement aid: per-SM L2 streaming rate.  `ctas` CTAs each re-read their own `per_cta` bytes of d_buf
 * (ctas * per_cta bytes, L2 resident) `iters` times.  mode 0 = LDG.128 with 8 loads in flight per thread;
 * mode 1 = TMA bulk copies of `chunk` bytes, `depth` in flight (the decoder's pipeline pattern).
 * Returns GB/s per CTA (= per SM).
 */
B200TACO_API int32_t b200taco_measure_stream(void* d_buf, size_t per_cta, int32_t ctas, int32_t iters, int32_t mode,
                                             int32_t chunk, int32_t depth, float* gbps_per_sm, void* stream);

/*
 * Measurement aid: SM cycles per operation (one thread, each preceded by a global store) of
 * out8[0..7] = fence.proxy.async, fence.proxy.async.global, fence.proxy.async.shared::cta, __threadfence,
 * red.release.gpu, ld.acquire.gpu, atomicAdd (returning), plain store.
 */
B200TACO_API int32_t b200taco_measure_sync_costs(int64_t* out8, void* stream);

/*
 * Measurement aid: per-phase device time of the LAST decoder launch on this handle, measured
 * by CTA 0 between grid barriers (globaltimer, ns, summed over decode steps):
 *   out[0..3] = attention LSTM, attention, decoder LSTM, projection+stop+prenet;
 *   out[4]    = decode steps executed; out[5..20] = sub-phase timers, zero unless the library
 * was built with -DTACO_PROFILE_FINE.  5 <= n <= 21.  Synchronises `stream`.  The workspace of
 * that launch must still be alive.
 */
B200TACO_API int32_t b200taco_decoder_profile(b200taco_handle* h, int64_t* out, int32_t n, void* stream);

/*
 * Measurement aid: globaltimer values (ns) that thread 0 of every CTA of the weight-resident small-batch decoder kernel
 * recorded at its stage boundaries during ONE decode step (environment TACO_RES_TRACE=<step> at launch time), 16 per CTA.
 */
B200TACO_API int32_t b200taco_decoder_trace(b200taco_handle* h, int64_t* out, int32_t n, void* stream);

/*
 * Measurement aid: phase times (ns, summed over launches since the last reset) of CTA 0 of the training attention backward
 * (attention_bwd_v3_kernel) when the process runs with TACO_ATTN_TIMES=1: out8[0] loads of weights and saved activations, [1] location
 * features, [2] wait for the preceding kernel + d ctx + dw / s, [3] reduction + first cluster exchange, [4] per-position pass,
 * [5] reductions + halo exchange, [6] location-filter gradient, [7] location-input gradient.  Synchronises the device.
 */
B200TACO_API int32_t b200taco_attn_bwd_phase_ns(uint64_t* out8, int32_t reset);

/*
 * Unit-test aid for the tcgen05 path: D[128][32] (fp32) = A[128][K] . B[32][K]^T with bf16
 * operands (row-major device arrays), one CTA, accumulators in TMEM.  K a multiple of 64.
 * variant 0 is the production descriptor convention; other values are bring-up experiments.
 */
B200TACO_API int32_t b200taco_debug_umma_gemm(const void* d_A, const void* d_B, float* d_D, int32_t K,
                                              int32_t variant, void* stream);

#ifdef __cplusplus
}
#endif

#endif  /* B200TACO_DEBUG_H */
