// tcgen05 (5th-generation tensor core) helpers for sm_100a: TMEM allocation, shared-memory
// matrix descriptors, instruction descriptors, MMA issue/commit and TMEM loads -- inline PTX,
// no CUTLASS.  Only what this library needs: kind::f16 with bf16 operands, fp32 accumulate,
// both operands K-major in shared memory, no swizzle ("interleaved" 8x16-byte core matrices).
//
// Operand layout in shared memory (both A [M x K] and B [N x K], K-major):
//   core matrix  = 8 rows x 8 bf16 (16 bytes per row), 128 contiguous bytes
//   tile         = [row group rg][k group kg][8 rows][8 k]
//   LBO (leading-dimension byte offset)  = bytes between core matrices adjacent in K  = 128
//   SBO (stride-dimension  byte offset)  = bytes between core matrices adjacent in M/N = KG * 128
// One tcgen05.mma consumes K = 16 = two k groups; successive K steps advance the start address
// by 256 bytes.
#pragma once

#include "common.cuh"

namespace b200taco {
namespace umma {

#ifdef __CUDACC__

// 64-bit shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor").
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);          // bits  0-13 start address >> 4
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;    // bits 16-29 leading byte offset >> 4
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;    // bits 32-45 stride byte offset >> 4
    d |= (uint64_t)1 << 46;                               // bits 46-47 descriptor version (Blackwell = 1)
    // base offset 0, lbo mode 0, layout type 0 = SWIZZLE_NONE (bits 61-63)
    return d;
}

// 32-bit instruction descriptor for kind::f16: D fp32, A/B bf16, both K-major, dense.
__host__ __device__ constexpr uint32_t instr_desc_bf16_f32(int M, int N) {
    return (1u << 4)                       // D format: f32
           | (1u << 7)                     // A format: bf16
           | (1u << 10)                    // B format: bf16
           | ((uint32_t)(N >> 3) << 17)    // N / 8
           | ((uint32_t)(M >> 4) << 24);   // M / 16
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// mbarrier arrive (count 1) when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// TMEM -> registers: 32 lanes (this warp's quarter: lanes 32*(warp%4)..) x 32 consecutive columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// split an fp32 value into two bf16 parts with hi + lo == x to ~2^-17 relative
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

#endif  // __CUDACC__

// Offset (in elements) of (row, k) inside a K-major no-swizzle tile with KG k-groups per chunk.
__host__ __device__ constexpr int tile_offset(int row, int k, int KG) {
    return (((row >> 3) * KG + (k >> 3)) << 6) + ((row & 7) << 3) + (k & 7);
}

}  // namespace umma
}  // namespace b200taco
