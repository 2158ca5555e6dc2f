// Diagnostic probe for the tcgen05 TF32 update kernels: D[128 x 64] = A[128 x 64] . B[64 x 64] on ONE tile, with the
// operand sources / majors the update kernel design needs, so that each mechanism is verified in isolation:
//   mode 0  SS: A smem K-major,  B smem K-major      (the textbook TN case)
//   mode 1  SS: A smem K-major,  B smem MN-major     (chain GEMM: activations [sample][unit] x weights [k][n])
//   mode 2  TS: A from TMEM,     B smem MN-major     (activations written by tcgen05.st)
//   mode 3  SS: A smem MN-major, B smem K-major      (Gram product: both operands stored [k = sample-major]...)
//   mode 4  SS: A smem MN-major, B smem MN-major
// Canonical no-swizzle layouts (cute/atom/mma_traits_sm100.hpp make_umma_desc, units of 16 bytes):
//   K-major : element (mn, k) at (k%4)*4 + (mn%8)*16 + (mn/8)*SBO + (k/4)*LBO
//   MN-major: element (mn, k) at (mn%4)*4 + (k%8)*16 + (mn/4)*SBO + (k/8)*LBO
// The accumulator is pre-filled with a sentinel so that "MMA wrote nothing" and "MMA multiplied zeros" are told apart;
// the A region of TMEM is read back (mode 2) to verify tcgen05.st.  Every wait is bounded.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int M = 128, N = 64, K = 64, KSTEP = 8;
constexpr int TMEM_COLS = 256;   // D [0,64), A [64,128)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ uint32_t make_idesc(int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offsets of element (mn, k) for a [MN x 64] operand
__device__ __forceinline__ uint32_t off_kmajor(int mn, int k, uint32_t sbo, uint32_t lbo) {
  return (k & 3) * 4 + (mn & 7) * 16 + (mn >> 3) * sbo + (k >> 2) * lbo;
}
__device__ __forceinline__ uint32_t off_mnmajor(int mn, int k, uint32_t sbo, uint32_t lbo) {
  return (mn & 3) * 4 + (k & 7) * 16 + (mn >> 2) * sbo + (k >> 3) * lbo;
}

__global__ void __launch_bounds__(128, 1)
    umma_modes_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                      float* __restrict__ Aback, int* __restrict__ status, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);                 // 128 x 64 floats = 32 KB
  float* sB = reinterpret_cast<float*>(smem + 32768);         // 64 x 64 floats = 16 KB
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + 49152 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool a_tmem = (mode == 2);
  const bool a_mn = (mode == 3 || mode == 4);
  const bool b_mn = (mode == 1 || mode == 2 || mode == 4);
  // K-major A: core matrices 8 rows x 16 B contiguous; SBO = 128 (next 8 rows), LBO = (M/8)*128 (next 4 k's)
  const uint32_t a_sbo = a_mn ? 128u : 128u, a_lbo = a_mn ? (M / 4) * 128u : (M / 8) * 128u;
  const uint32_t b_sbo = 128u, b_lbo = b_mn ? (N / 4) * 128u : (N / 8) * 128u;

  for (int e = tid; e < M * K; e += blockDim.x) {
    const int m = e / K, k = e % K;
    const uint32_t off = a_mn ? off_mnmajor(m, k, a_sbo, a_lbo) : off_kmajor(m, k, a_sbo, a_lbo);
    sA[off >> 2] = A[e];
  }
  for (int e = tid; e < K * N; e += blockDim.x) {
    const int k = e / N, n = e % N;
    const uint32_t off = b_mn ? off_mnmajor(n, k, b_sbo, b_lbo) : off_kmajor(n, k, b_sbo, b_lbo);
    sB[off >> 2] = B[e];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *tmem_holder;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t tD = tbase + lane_base, tA = tbase + lane_base + 64;
  if (tid == 0) status[1] = (int)tbase;

  // sentinel into D, A rows into TMEM (mode 2)
  {
    uint32_t r[16];
#pragma unroll
    for (int c = 0; c < N; c += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(7.0f);
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
          "%15, %16};" ::"r"(tD + c),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
          : "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(A[tid * K + c + j]);
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
          "%15, %16};" ::"r"(tA + c),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (tid == 0) {
    const uint32_t idesc = make_idesc(a_mn ? 1 : 0, b_mn ? 1 : 0);
    const uint64_t dA = make_desc(smem_u32(sA), a_lbo, a_sbo), dB = make_desc(smem_u32(sB), b_lbo, b_sbo);
#pragma unroll
    for (int ks = 0; ks < K / KSTEP; ++ks) {
      // advance one K step (8 elements): K-major = 2 LBO strides (2 x 4 k's), MN-major = 1 LBO stride (8 k rows)
      const uint64_t a_koff = (uint64_t)(((a_mn ? 1 : 2) * ks * a_lbo) >> 4);
      const uint64_t b_koff = (uint64_t)(((b_mn ? 1 : 2) * ks * b_lbo) >> 4);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      if (a_tmem) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tbase),
            "r"(tbase + 64 + ks * KSTEP), "l"(dB + b_koff), "r"(idesc), "r"(acc)
            : "memory");
      } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase),
            "l"(dA + a_koff), "l"(dB + b_koff), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
                 : "memory");
  }

  uint32_t done = 0;
  for (int it = 0; it < (1 << 22) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(mbar)), "r"(0u)
        : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int all_done = __syncthreads_and((int)done);
  {
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
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
          "[%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tA + c)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) Aback[tid * K + c + j] = __uint_as_float(r[j]);
    }
  }
  if (tid == 0) status[0] = all_done ? 1 : -1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(TMEM_COLS));
}

}  // namespace

extern "C" int umma_modes_probe(const float* A, const float* B, float* D, float* Aback, int* status, int mode,
                                void* stream) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(umma_modes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
    attr = true;
  }
  umma_modes_kernel<<<1, 128, 50 * 1024, (cudaStream_t)stream>>>(A, B, D, Aback, status, mode);
  return (int)cudaGetLastError();
}
