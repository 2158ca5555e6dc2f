// tcgen05 / TMEM helpers shared by the tensor-core update kernels (update_umma.cu: 64-wide Fisher-vector product;
// update_umma32.cu: 32-wide gradient and Fisher-vector product).  Inline PTX for sm_100a: UMMA shared-memory and instruction
// descriptors (cute/arch/mma_sm100_desc.hpp), tcgen05.mma with the A operand in TMEM, commit -> mbarrier, bounded waits,
// tcgen05.ld / tcgen05.st in the 32x32b shape (thread i of a warp <-> TMEM lane base + i).
#pragma once
#include <stdint.h>

#include "update_common.cuh"

namespace b200rl {

__device__ __forceinline__ uint32_t u_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// hi word of the operand split: the float32 word with its 13 low mantissa bits cleared (exactly representable in TF32);
// lo = x - hi is exact in float32.  (Rounding hi to nearest -- cvt.rna.tf32 -- instead of truncating removes the 2^-22
// shrink of every operand that the tensor core's own truncation of the lo word causes, at one more instruction per
// element; measured on the Swimmer learning curve it changes nothing: 22.8 +- 0.4 vs 23.7 +- 2.7 over 8 seeds)
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp): start >> 4 | LBO >> 4 << 16 |
// SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t u_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D f32 (1 @4), A/B tf32 (2 @7 / @10), both K-major, N >> 3 @17, M >> 4 @24
__host__ __device__ constexpr uint32_t u_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t U_IDESC = u_idesc(128, 64);

__device__ __forceinline__ void u_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate,
                                         uint32_t idesc = U_IDESC) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void u_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(u_smem_u32(bar))
               : "memory");
}
// bounded wait: a kernel that hangs costs the whole GPU box; on time-out the caller poisons its output instead
__device__ __forceinline__ bool u_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (int it = 0; it < (1 << 26) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(u_smem_u32(bar)), "r"(parity)
        : "memory");
  }
  return done != 0;
}
__device__ __forceinline__ void u_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void u_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

#define U_R16(r, o) "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), \
                    "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]), "r"(r[o + 12]),              \
                    "r"(r[o + 13]), "r"(r[o + 14]), "r"(r[o + 15])
#define U_W16(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]),            \
                    "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]),          \
                    "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])

// 32 lanes x 16 columns of 32-bit: thread i of the warp <-> TMEM lane (base lane + i)
__device__ __forceinline__ void u_st16(uint32_t taddr, const uint32_t (&r)[32], int o) {
  if (o == 0)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                 "%15, %16};" ::"r"(taddr), U_R16(r, 0) : "memory");
  else
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                 "%15, %16};" ::"r"(taddr), U_R16(r, 16) : "memory");
}
__device__ __forceinline__ void u_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  u_st16(taddr, r, 0);
  u_st16(taddr + 16, r, 16);
}
__device__ __forceinline__ void u_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
               "[%16];" : U_W16(r, 0) : "r"(taddr) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
               "[%16];" : U_W16(r, 16) : "r"(taddr + 16) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void u_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace b200rl
