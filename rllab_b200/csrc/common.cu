// Error plumbing, device queries and the fixed-order finalize kernels shared by every reduction.
#include <stdarg.h>

#include "common.cuh"

namespace b200rl {

static thread_local char g_err[512] = "";
unsigned long long g_kernel_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return B200RL_ECUDA;
}

int num_sms() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
  cached = sms;
  return sms;
}

// out[k] = op over blocks of partial[b][k].  32 outputs x 8 block-slices per CTA: each slice walks every 8th block
// (coalesced 256 B rows, 4 loads in flight), the slices are combined through shared memory in fixed order, so the result
// is deterministic and the dependent-add chain is nblocks/8 long instead of nblocks (the one-thread-per-output version
// cost ~56 us per call, 4 % of the cfg2 iteration -- profiles/r01c).
template <bool IS_MAX>
__global__ void __launch_bounds__(256) finalize_kernel(const double* __restrict__ partial, int nblocks, int K,
                                                       double* __restrict__ out, double scale) {
  __shared__ double sm[8][33];
  const int kx = threadIdx.x & 31, by = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + kx;
  double acc = IS_MAX ? -1.0e300 : 0.0;
  if (k < K) {
#pragma unroll 4
    for (int b = by; b < nblocks; b += 8) {
      const double v = partial[(size_t)b * K + k];
      acc = IS_MAX ? fmax(acc, v) : acc + v;
    }
  }
  sm[by][kx] = acc;
  __syncthreads();
  if (by == 0 && k < K) {
    double r = sm[0][kx];
#pragma unroll
    for (int y = 1; y < 8; ++y) r = IS_MAX ? fmax(r, sm[y][kx]) : r + sm[y][kx];
    out[k] = IS_MAX ? r : r * scale;
  }
}

int launch_finalize_sum(const double* partial, int nblocks, int K, double* out, double scale, cudaStream_t s) {
  finalize_kernel<false><<<(K + 31) / 32, 256, 0, s>>>(partial, nblocks, K, out, scale);
  B200RL_LAUNCH_CHECK("finalize_kernel<sum>");
  return 0;
}

int launch_finalize_max(const double* partial, int nblocks, int K, double* out, cudaStream_t s) {
  finalize_kernel<true><<<(K + 31) / 32, 256, 0, s>>>(partial, nblocks, K, out, 1.0);
  B200RL_LAUNCH_CHECK("finalize_kernel<max>");
  return 0;
}

}  // namespace b200rl

extern "C" {

const char* b200rl_last_error(void) { return b200rl::g_err; }

int b200rl_version(void) { return B200RL_VERSION; }

int b200rl_device_sms(int* sms_out) {
  int dev = 0, sms = 0;
  B200RL_CUDA_CHECK(cudaGetDevice(&dev));
  B200RL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sms_out) *sms_out = sms;
  return 0;
}

unsigned long long b200rl_kernel_launches(void) { return b200rl::g_kernel_launches; }

long long b200rl_ws_doubles(void) {
  return (long long)b200rl::MAX_PARTIAL_BLOCKS * b200rl::MAX_PARTIAL_K;
}
}
