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

__global__ void finalize_sum_kernel(const double* __restrict__ partial, int nblocks, int K, double* __restrict__ out,
                                    double scale) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * K + k];
  out[k] = s * scale;
}

__global__ void finalize_max_kernel(const double* __restrict__ partial, int nblocks, int K, double* __restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double s = -1.0e300;
  for (int b = 0; b < nblocks; ++b) s = fmax(s, partial[(size_t)b * K + k]);
  out[k] = s;
}

int launch_finalize_sum(const double* partial, int nblocks, int K, double* out, double scale, cudaStream_t s) {
  finalize_sum_kernel<<<(K + 127) / 128, 128, 0, s>>>(partial, nblocks, K, out, scale);
  B200RL_LAUNCH_CHECK("finalize_sum_kernel");
  return 0;
}

int launch_finalize_max(const double* partial, int nblocks, int K, double* out, cudaStream_t s) {
  finalize_max_kernel<<<(K + 127) / 128, 128, 0, s>>>(partial, nblocks, K, out);
  B200RL_LAUNCH_CHECK("finalize_max_kernel");
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
