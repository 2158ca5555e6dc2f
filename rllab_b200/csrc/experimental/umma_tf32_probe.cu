// Stand-alone probe for the round-2 tensor-core update kernels.  NOT part of libb200rl.so,
// not built by default, never run on a GPU yet.  It checks the three mechanisms the design relies on, on one tile:
//
//   D[128 x 64] (TMEM accumulator, fp32) = A[128 x 64] (TF32, read from TMEM: lane = sample row, column = k)
//                                        x B[64 x 64]  (TF32, shared memory, MN-major canonical no-swizzle layout)
//
//   1. tcgen05.st of a thread-per-row activation tile as the A operand (32x32b: lane i of warp w = TMEM lane 32w + i),
//   2. the shared-memory matrix descriptor of cute/atom/mma_traits_sm100.hpp for a [K][N] row-major weight matrix
//      (MN-major, SWIZZLE_NONE: 8 K-rows x 16 B core matrices, SBO between 4-column groups, LBO between 8-row K blocks),
//   3. tcgen05.mma.cta_group::1.kind::tf32 issued by one thread, tcgen05.commit -> mbarrier, tcgen05.ld of the result.
//
// split = 1 runs the three-pass split of DESIGN.md (a_hi b_hi + a_lo b_hi + a_hi b_lo, hi = the word truncated to TF32,
// lo = x - hi): A_lo goes to a second TMEM operand block, B_lo to a second shared-memory image; the result must then
// match the full-precision product to float32 accuracy (~4e-7 of the output scale, scripts/tf32_split_study.py).
//
// Every wait is bounded: if the MMA never signals the mbarrier the kernel reports status -1 instead of hanging.
// Build:  make -C rllab_b200/csrc umma_probe      Run on a B200:  timeout 120 python scripts/umma_probe.py
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int M = 128, N = 64, K = 64, KSTEP = 8;            // one tcgen05.mma.kind::tf32 consumes K = 8 (32 bytes)
constexpr int TMEM_COLS = 256;                                // D: [0, 64), A (hi): [64, 128), A_lo: [128, 192)
constexpr uint32_t SBO = 128, LBO = (N / 4) * 128;            // bytes: next 4-column group, next 8-row K block

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                                     // descriptor version (Blackwell)
  return d;                                                   // layout_type (bits 61..63) = 0: SWIZZLE_NONE
}

// UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format TF32 (2) @7/@10, a_major K (0) @15, b_major MN (1) @16,
// n_dim = N >> 3 @17, m_dim = M >> 4 @24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ D, int* __restrict__ status, int split) {
  __shared__ __align__(1024) float sB[K * N];                 // canonical layout, 16 KB
  __shared__ __align__(1024) float sBlo[K * N];               // lo halves (split mode)
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_holder;
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- B: [K][N] row-major in global -> element (n, k) at (n % 4) * 4 + (k % 8) * 16 + (n / 4) * SBO + (k / 8) * LBO
  for (int e = tid; e < K * N; e += blockDim.x) {
    const int k = e / N, n = e % N;
    const uint32_t off = (n & 3) * 4 + (k & 7) * 16 + (n >> 2) * SBO + (k >> 3) * LBO;
    sB[off >> 2] = B[e];
    sBlo[off >> 2] = B[e] - tf32_hi(B[e]);                    // exact; truncated again to TF32 by the tensor core
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {                                            // one full warp allocates (and later frees) TMEM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_holder)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of sB -> visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_holder;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;     // TMEM address = lane << 16 | column
  const uint32_t tD = tbase + lane_base, tA = tbase + lane_base + 64;

  // ---- A: thread t holds row t (64 values) and stores it to TMEM columns [64, 128) of lane t
  {
    uint32_t r[16];
#pragma unroll
    for (int c = 0; c < K; c += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(A[tid * K + c + j]);
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
          "%15, %16};" ::"r"(tA + c),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
          : "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {                          // lo = x - hi (exact), into the second operand block
        const float x = __uint_as_float(r[j]);
        r[j] = __float_as_uint(x - tf32_hi(x));
      }
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
          "%15, %16};" ::"r"(tA + 64 + c),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- MMA: one thread issues K / 8 instructions, then commits to the mbarrier
  if (tid == 0) {
    const uint64_t desc_hi = make_desc(smem_u32(sB)), desc_lo = make_desc(smem_u32(sBlo));
    auto mma = [&](uint32_t a_addr, uint64_t descB, uint32_t accumulate) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tbase),
          "r"(a_addr), "l"(descB), "r"(IDESC), "r"(accumulate)
          : "memory");
    };
#pragma unroll
    for (int ks = 0; ks < K / KSTEP; ++ks) {
      const uint64_t koff = (uint64_t)((ks * LBO) >> 4);                 // start address advances one K block
      const uint32_t a_hi = tbase + 64 + ks * KSTEP, a_lo = tbase + 128 + ks * KSTEP;   // A columns of this K step
      if (split) {                                                       // small terms first, then the main product
        mma(a_lo, desc_hi + koff, ks > 0 ? 1u : 0u);
        mma(a_hi, desc_lo + koff, 1u);
        mma(a_hi, desc_hi + koff, 1u);
      } else {
        mma(a_hi, desc_hi + koff, ks > 0 ? 1u : 0u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                 : "memory");
  }

  // ---- bounded wait for the accumulator
  uint32_t done = 0;
  for (int it = 0; it < (1 << 22) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(&mbar)), "r"(0u)
        : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!__syncthreads_and((int)done)) {
    if (tid == 0) *status = -1;                               // the MMA never completed: report instead of hanging
  } else {
    uint32_t r[16];
#pragma unroll
    for (int c = 0; c < N; c += 16) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
          "[%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tD + c)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) D[tid * N + c + j] = __uint_as_float(r[j]);
    }
    if (tid == 0) *status = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(TMEM_COLS));
}

}  // namespace

extern "C" int umma_probe(const float* A, const float* B, float* D, int* status, int split, void* stream) {
  umma_probe_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(A, B, D, status, split);
  return (int)cudaGetLastError();
}
