// process_samples numeric core on the lane layout: LinearFeatureBaseline.predict + GAE/returns reverse scan +
// the reductions behind the tabular statistics; advantage centering; LinearFeatureBaseline.fit normal equations.
//
// Replaces: rllab/sampler/base.py:48-93,163-180 ; rllab/misc/special.py:51-59,107-111 ; rllab/algos/util.py:7-12 ;
//           rllab/baselines/linear_feature_baseline.py:19-43.
// All three kernels are HBM-streaming (16-40 B per sample); sums are float64, two-stage, fixed order.
#include "common.cuh"

namespace b200rl {

constexpr int PS_THREADS = 128;
constexpr int OMAX = 32;  // max obs_dim handled by the runtime-O feature code

// LinearFeatureBaseline features . w  (linear_feature_baseline.py:19-23): [clip(o,+-10), o^2, al, al^2, al^3, 1]
__device__ __forceinline__ double lfb_predict(const float* __restrict__ obs, size_t plane, size_t idx, int O,
                                              unsigned short ts, const double* __restrict__ w) {
  double acc = 0.0;
  for (int k = 0; k < O; ++k) {
    double o = (double)fminf(fmaxf(obs[k * plane + idx], -10.0f), 10.0f);
    acc += o * w[k] + (o * o) * w[O + k];
  }
  double al = (double)ts / 100.0;
  acc += al * w[2 * O] + (al * al) * w[2 * O + 1] + (al * al * al) * w[2 * O + 2] + w[2 * O + 3];
  return acc;
}

// One thread per lane, reverse scan over t.  Recurrences in float64 (the reference runs them in float64).
__global__ void __launch_bounds__(PS_THREADS)
    process_samples_kernel(int O, int N, int T, const float* __restrict__ obs, const float* __restrict__ rew,
                           const unsigned char* __restrict__ flags, const unsigned short* __restrict__ tstep,
                           const double* __restrict__ w, double discount, double gl, float* __restrict__ adv,
                           float* __restrict__ ret, float* __restrict__ base, double* __restrict__ partial_sum,
                           double* __restrict__ partial_max) {
  __shared__ double sw[2 * OMAX + 4];
  __shared__ double scratch[B200RL_PS_NSUM * 32];
  const bool have_w = (w != nullptr);
  if (have_w)
    for (int i = threadIdx.x; i < 2 * O + 4; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double s[B200RL_PS_NSUM];
  double m[B200RL_PS_NMAX];
#pragma unroll
  for (int i = 0; i < B200RL_PS_NSUM; ++i) s[i] = 0.0;
#pragma unroll
  for (int i = 0; i < B200RL_PS_NMAX; ++i) m[i] = -1.0e300;
  if (n < N) {
    const size_t plane = (size_t)T * N;
    double a_next = 0.0, r_next = 0.0, u_next = 0.0, b_next = 0.0;
    // The recurrences are sequential in t but the loads are not: fetch CH steps ahead of the arithmetic so that CH
    // (x up to O+3) independent loads are in flight per thread (the naive loop is long-scoreboard bound, ncu r01b).
    constexpr int CH = 4;
    for (int tb = T - 1; tb >= 0; tb -= CH) {
      unsigned char fl[CH];
      unsigned short tsv[CH];
      float rw[CH];
      double bs[CH];
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int t = tb - u;
        const size_t idx = (size_t)(t >= 0 ? t : 0) * N + n;
        fl[u] = flags[idx]; tsv[u] = tstep[idx]; rw[u] = rew[idx];
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int t = tb - u;
        const size_t idx = (size_t)(t >= 0 ? t : 0) * N + n;
        bs[u] = have_w ? lfb_predict(obs, plane, idx, O, tsv[u], sw) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int t = tb - u;
        if (t < 0) break;
        const size_t idx = (size_t)t * N + n;
        const unsigned char f = fl[u];
        const unsigned short ts = tsv[u];
        const double r = (double)rw[u];
        const double b = bs[u];
        if (f & B200RL_FLAG_END) { a_next = 0.0; r_next = 0.0; u_next = 0.0; b_next = 0.0; }
        const double delta = r + discount * b_next - b;   // base.py:59-61
        a_next = delta + gl * a_next;                      // discount_cumsum(deltas, discount*lambda)
        r_next = r + discount * r_next;                    // discount_cumsum(rewards, discount)
        u_next = r + u_next;
        b_next = b;
        adv[idx] = (float)a_next; ret[idx] = (float)r_next; base[idx] = (float)b;
        // statistics use the float64 values (as the reference does)
        s[0] += a_next; s[1] += a_next * a_next; s[2] += 1.0;
        s[7] += r_next; s[8] += r_next * r_next; s[9] += b; s[10] += b * b;
        const double res = r_next - b;
        s[11] += res; s[12] += res * res;
        m[2] = fmax(m[2], -a_next); m[3] = fmax(m[3], a_next);
        if (ts == 0) {  // first sample of a path
          s[3] += 1.0; s[4] += r_next; s[5] += u_next; s[6] += u_next * u_next;
          m[0] = fmax(m[0], u_next); m[1] = fmax(m[1], -u_next);
        }
      }
    }
  }
  block_reduce_store<B200RL_PS_NSUM, false>(s, scratch, partial_sum + (size_t)blockIdx.x * B200RL_PS_NSUM);
  block_reduce_store<B200RL_PS_NMAX, true>(m, scratch, partial_max + (size_t)blockIdx.x * B200RL_PS_NMAX);
}

// (adv - mean) / (std + 1e-8), then optionally (adv - min) + 1e-8   (algos/util.py:7-12)
__global__ void center_adv_kernel(float* __restrict__ adv, long long B, const double* __restrict__ sums,
                                  const double* __restrict__ maxs, int center, int positive) {
  const double cnt = sums[2];
  const double mean = sums[0] / cnt;
  double var = sums[1] / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  const double stdv = sqrt(var) + 1e-8;
  double mn = -maxs[2];
  if (center) mn = (mn - mean) / stdv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += stride) {
    double a = (double)adv[i];
    if (center) a = (a - mean) / stdv;
    if (positive) a = (a - mn) + 1e-8;
    adv[i] = (float)a;
  }
}

// Gram matrix of f = [features(d), ret] over samples: upper triangle, float64.
// Tile = 128 samples per block iteration; features staged in shared memory [d+1][TILE+4]; thread p owns pairs
// p, p+blockDim, ...; per-tile float32 dot products, float64 accumulation across tiles.
constexpr int GRAM_TILE = 128;
constexpr int GRAM_LD = GRAM_TILE + 4;
constexpr int GRAM_THREADS = 128;
constexpr int GRAM_MAXPAIRS_PER_THREAD = 9;  // (2*20+5)*(2*20+6)/2 = 1035 pairs / 128 threads

__global__ void __launch_bounds__(GRAM_THREADS)
    lfb_gram_kernel(int O, long long B, const float* __restrict__ obs, const unsigned short* __restrict__ tstep,
                    const float* __restrict__ ret, double* __restrict__ partial) {
  extern __shared__ __align__(16) float F[];  // [(d+1)][GRAM_LD]
  const int d1 = 2 * O + 5;
  const int npairs = d1 * (d1 + 1) / 2;
  double acc[GRAM_MAXPAIRS_PER_THREAD];
  int pi[GRAM_MAXPAIRS_PER_THREAD], pj[GRAM_MAXPAIRS_PER_THREAD];
#pragma unroll
  for (int q = 0; q < GRAM_MAXPAIRS_PER_THREAD; ++q) {
    acc[q] = 0.0;
    int p = threadIdx.x + q * GRAM_THREADS;
    // decode p -> (i, j), i <= j, row-major upper triangle
    int i = 0, rem = p;
    if (p < npairs) {
      while (rem >= d1 - i) { rem -= d1 - i; ++i; }
      pi[q] = i; pj[q] = i + rem;
    } else {
      pi[q] = 0; pj[q] = 0;
    }
  }
  const long long ntiles = (B + GRAM_TILE - 1) / GRAM_TILE;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long sidx = tile * GRAM_TILE + threadIdx.x;
    __syncthreads();
    if (sidx < B) {
      for (int k = 0; k < O; ++k) {
        float o = fminf(fmaxf(obs[(size_t)k * B + sidx], -10.0f), 10.0f);
        F[k * GRAM_LD + threadIdx.x] = o;
        F[(O + k) * GRAM_LD + threadIdx.x] = o * o;
      }
      float al = (float)tstep[sidx] / 100.0f;
      F[(2 * O) * GRAM_LD + threadIdx.x] = al;
      F[(2 * O + 1) * GRAM_LD + threadIdx.x] = al * al;
      F[(2 * O + 2) * GRAM_LD + threadIdx.x] = al * al * al;
      F[(2 * O + 3) * GRAM_LD + threadIdx.x] = 1.0f;
      F[(2 * O + 4) * GRAM_LD + threadIdx.x] = ret[sidx];
    } else {
      for (int k = 0; k < d1; ++k) F[k * GRAM_LD + threadIdx.x] = 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < GRAM_MAXPAIRS_PER_THREAD; ++q) {
      if (threadIdx.x + q * GRAM_THREADS < npairs) {
        const float4* ra = reinterpret_cast<const float4*>(F + pi[q] * GRAM_LD);
        const float4* rb = reinterpret_cast<const float4*>(F + pj[q] * GRAM_LD);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
        for (int k = 0; k < GRAM_TILE / 4; ++k) {
          float4 a = ra[k], b = rb[k];
          s0 = fmaf(a.x, b.x, s0); s1 = fmaf(a.y, b.y, s1); s2 = fmaf(a.z, b.z, s2); s3 = fmaf(a.w, b.w, s3);
        }
        acc[q] += (double)((s0 + s1) + (s2 + s3));
      }
    }
  }
#pragma unroll
  for (int q = 0; q < GRAM_MAXPAIRS_PER_THREAD; ++q) {
    int p = threadIdx.x + q * GRAM_THREADS;
    if (p < npairs) partial[(size_t)blockIdx.x * npairs + p] = acc[q];
  }
}

// Small observation spaces (O <= 4, i.e. d+1 <= 13 features): the whole upper triangle (<= 91 products) fits in one
// thread's registers, so each thread streams its samples (coalesced across the warp), accumulates the outer products
// in float32 registers (<= ~100 samples per thread), and the block folds them once into float64 -- no shared-memory
// staging, no bank conflicts (the staged kernel above spends its time in 13 M conflicts and LSU latency).
template <int O>
__global__ void __launch_bounds__(128) lfb_gram_reg_kernel(long long B, const float* __restrict__ obs,
                                                           const unsigned short* __restrict__ tstep,
                                                           const float* __restrict__ ret, double* __restrict__ partial) {
  constexpr int D1 = 2 * O + 5, NP = D1 * (D1 + 1) / 2;
  __shared__ double red[NP];
  float acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.f;
  for (int p = threadIdx.x; p < NP; p += blockDim.x) red[p] = 0.0;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  constexpr int UNR = 4;   // 4 samples' loads in flight per thread (memory-level parallelism; one is latency bound)
  for (long long s0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; s0 < B; s0 += UNR * stride) {
    float raw[UNR][O + 2];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long s = s0 + u * stride;
      const bool ok = s < B;
      const long long sl = ok ? s : s0;
#pragma unroll
      for (int k = 0; k < O; ++k) raw[u][k] = obs[(size_t)k * B + sl];
      raw[u][O] = (float)tstep[sl];
      raw[u][O + 1] = ret[sl];
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (s0 + u * stride < B) {
        float f[D1];
#pragma unroll
        for (int k = 0; k < O; ++k) {
          const float o = fminf(fmaxf(raw[u][k], -10.0f), 10.0f);
          f[k] = o;
          f[O + k] = o * o;
        }
        const float al = raw[u][O] / 100.0f;
        f[2 * O] = al; f[2 * O + 1] = al * al; f[2 * O + 2] = al * al * al; f[2 * O + 3] = 1.0f; f[2 * O + 4] = raw[u][O + 1];
        int p = 0;
#pragma unroll
        for (int i = 0; i < D1; ++i)
#pragma unroll
          for (int j = i; j < D1; ++j) { acc[p] = fmaf(f[i], f[j], acc[p]); ++p; }
      }
    }
  }
  // fold: warp shuffle in float64, then one shared-memory add per warp in fixed warp order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
    if (warp == w) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const double v = warp_sum((double)acc[p]);
        if (lane == 0) red[p] += v;
      }
    }
    __syncthreads();
  }
  for (int p = threadIdx.x; p < NP; p += blockDim.x) partial[(size_t)blockIdx.x * NP + p] = red[p];
}

__global__ void planes_to_rows_kernel(int dim, long long B, const float* __restrict__ src, double* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  for (int k = 0; k < dim; ++k) dst[i * dim + k] = (double)src[(size_t)k * B + i];
}

}  // namespace b200rl

using namespace b200rl;

extern "C" {

int b200rl_process_samples(int obs_dim, int N, int T, const float* obs, const float* rew, const unsigned char* flags,
                           const unsigned short* tstep, const double* w, double discount, double gae_lambda,
                           float* adv, float* ret, float* base, double* sums_out, double* maxs_out, double* ws,
                           void* stream) {
  B200RL_REQUIRE(obs && rew && flags && tstep && adv && ret && base && sums_out && maxs_out && ws,
                 "process_samples: null buffer");
  B200RL_REQUIRE(N > 0 && T > 0 && obs_dim > 0 && obs_dim <= OMAX, "process_samples: bad sizes");
  const int grid = (N + PS_THREADS - 1) / PS_THREADS;
  // lanes beyond the partial-block budget: use a wider block instead of more blocks
  B200RL_REQUIRE(grid <= 65536, "process_samples: too many lanes for one call (N <= 8M)");
  cudaStream_t st = (cudaStream_t)stream;
  double* psum = ws;
  double* pmax = ws + (size_t)grid * B200RL_PS_NSUM;
  B200RL_REQUIRE((long long)grid * (B200RL_PS_NSUM + B200RL_PS_NMAX) <= b200rl_ws_doubles(), "workspace too small");
  process_samples_kernel<<<grid, PS_THREADS, 0, st>>>(obs_dim, N, T, obs, rew, flags, tstep, w, discount,
                                                      discount * gae_lambda, adv, ret, base, psum, pmax);
  B200RL_LAUNCH_CHECK("process_samples_kernel");
  int rc = launch_finalize_sum(psum, grid, B200RL_PS_NSUM, sums_out, 1.0, st);
  if (rc) return rc;
  return launch_finalize_max(pmax, grid, B200RL_PS_NMAX, maxs_out, st);
}

int b200rl_center_advantages(float* adv, long long B, const double* sums, const double* maxs, int center,
                             int positive, void* stream) {
  B200RL_REQUIRE(adv && sums && maxs && B > 0, "center_advantages: bad arguments");
  if (!center && !positive) return 0;
  long long blocks = (B + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  center_adv_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(adv, B, sums, maxs, center, positive);
  B200RL_LAUNCH_CHECK("center_adv_kernel");
  return 0;
}

int b200rl_lfb_gram(int obs_dim, long long B, const float* obs, const unsigned short* tstep, const float* ret,
                    double* gram_out, double* ws, void* stream) {
  B200RL_REQUIRE(obs && tstep && ret && gram_out && ws && B > 0, "lfb_gram: bad arguments");
  B200RL_REQUIRE(obs_dim > 0 && obs_dim <= 20, "lfb_gram: obs_dim must be in 1..20");
  const int d1 = 2 * obs_dim + 5;
  const int npairs = d1 * (d1 + 1) / 2;
  const long long ntiles = (B + GRAM_TILE - 1) / GRAM_TILE;
  int grid = num_sms() * 4;
  if (grid > ntiles) grid = (int)ntiles;
  if (grid > MAX_PARTIAL_BLOCKS) grid = MAX_PARTIAL_BLOCKS;
  const size_t smem = (size_t)d1 * GRAM_LD * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (obs_dim <= 4) {
    long long g = (long long)num_sms() * 4;   // 2 resident CTAs (233 registers) x 2 waves
    const long long need = (B + 127) / 128;
    if (g > need) g = need;
    grid = (int)g;
    switch (obs_dim) {
      case 1: lfb_gram_reg_kernel<1><<<grid, 128, 0, st>>>(B, obs, tstep, ret, ws); break;
      case 2: lfb_gram_reg_kernel<2><<<grid, 128, 0, st>>>(B, obs, tstep, ret, ws); break;
      case 3: lfb_gram_reg_kernel<3><<<grid, 128, 0, st>>>(B, obs, tstep, ret, ws); break;
      default: lfb_gram_reg_kernel<4><<<grid, 128, 0, st>>>(B, obs, tstep, ret, ws); break;
    }
    B200RL_LAUNCH_CHECK("lfb_gram_reg_kernel");
    return launch_finalize_sum(ws, grid, npairs, gram_out, 1.0, st);
  }
  lfb_gram_kernel<<<grid, GRAM_THREADS, smem, st>>>(obs_dim, B, obs, tstep, ret, ws);
  B200RL_LAUNCH_CHECK("lfb_gram_kernel");
  return launch_finalize_sum(ws, grid, npairs, gram_out, 1.0, st);
}

int b200rl_planes_to_rows_f64(int dim, long long B, const float* src, double* dst, void* stream) {
  B200RL_REQUIRE(src && dst && dim > 0 && B > 0, "planes_to_rows: bad arguments");
  planes_to_rows_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dim, B, src, dst);
  B200RL_LAUNCH_CHECK("planes_to_rows_kernel");
  return 0;
}
}
