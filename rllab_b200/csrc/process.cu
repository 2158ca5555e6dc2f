// process_samples numeric core on the lane layout: LinearFeatureBaseline.predict + GAE/returns reverse scan +
// the reductions behind the tabular statistics; advantage centering; LinearFeatureBaseline.fit normal equations.
//
// Replaces: rllab/sampler/base.py:48-93,163-180 ; rllab/misc/special.py:51-59,107-111 ; rllab/algos/util.py:7-12 ;
//           rllab/baselines/linear_feature_baseline.py:19-43.
// All three kernels are HBM-streaming (16-40 B per sample); sums are float64, two-stage, fixed order.
#include "common.cuh"

namespace b200rl {

constexpr int OMAX = 32;  // max obs_dim handled by the runtime-O feature code

// LinearFeatureBaseline features . w  (linear_feature_baseline.py:19-23): [clip(o,+-10), o^2, al, al^2, al^3, 1]
// OT > 0: compile-time obs_dim -- the O loads of a step are issued back to back (and, with the caller's step loop
// unrolled, hoisted across steps) instead of one load -> use chain per feature (in-order issue stalls at the first use of
// every load: the runtime-O loop exposed 8 x O serial DRAM latencies per window -- 0.55 ms on cfg2, round 2 measurement)
template <int OT>
__device__ __forceinline__ double lfb_predict(const float* __restrict__ obs, size_t plane, size_t idx, int O_rt,
                                              unsigned short ts, const double* __restrict__ w) {
  double acc = 0.0;
  const int O = OT > 0 ? OT : O_rt;
  if constexpr (OT > 0) {
    float ov[OT];
#pragma unroll
    for (int k = 0; k < OT; ++k) ov[k] = obs[k * plane + idx];
#pragma unroll
    for (int k = 0; k < OT; ++k) {
      const double o = (double)fminf(fmaxf(ov[k], -10.0f), 10.0f);
      acc += o * w[k] + (o * o) * w[OT + k];
    }
  } else {
    for (int k = 0; k < O; ++k) {
      double o = (double)fminf(fmaxf(obs[k * plane + idx], -10.0f), 10.0f);
      acc += o * w[k] + (o * o) * w[O + k];
    }
  }
  double al = (double)ts / 100.0;
  acc += al * w[2 * O] + (al * al) * w[2 * O + 1] + (al * al * al) * w[2 * O + 2] + w[2 * O + 3];
  return acc;
}

// GAE / returns reverse scan, parallel over lanes AND over time (round 2; the one-thread-per-lane walk of round 1 kept
// only ~14 warps per SM busy at 65 536 lanes and sat at 26 % of the measured HBM bandwidth).
//
// A block owns 32 lanes (one warp-width, so every row access is one coalesced 128 B transaction) and walks T backwards
// in WINDOWS of PS_WARPS x PS_L steps: warp c of the block owns the c-th chunk of PS_L consecutive steps of the window
// (warp 0 = the latest steps).  The recurrences  x_t = c_t + m_t x_{t+1}  (m_t = factor, or 0 at the last sample of a
// path) are linear, so each window is resolved with a two-pass chunked scan:
//   P1  every thread loads its PS_L steps (all loads independent: PS_L x (O+3) in flight per thread), evaluates the
//       baseline b_t (float64) and publishes the b of its earliest step (the delta of the chunk before it needs it),
//   P2  local scan with zero carry-in -> values at the chunk's earliest step + "no path end inside the chunk",
//   P3  every thread folds the aggregates of the later chunks of its lane (<= PS_WARPS-1 three-term updates) into its
//       true carry-in,
//   P4  second scan from the registers with the true carry: writes adv / ret, accumulates the statistics.
// Inputs are read once and outputs written once: HBM traffic = the algorithmic 4*O+19 B per sample.
// Recurrences and statistics in float64, as the reference runs them (sampler/base.py:57-66, special.py:107-111).
//
// drop_cut != 0: a path that carries FLAG_CUT on its last sample (cut by the end of the lane buffer) is dropped, the
// way the reference's samplers only ever return whole paths (batch_polopt.py:30-34 with whole_paths=True;
// vectorized_sampler.py drops unfinished running_paths): its samples get FLAG_MASKED, adv = 0, and are excluded from
// every statistic (count, path counts, returns); downstream kernels skip masked samples.
constexpr int PS_L = 8, PS_WARPS = 8, PS_THREADS2 = 32 * PS_WARPS, PS_WIN = PS_L * PS_WARPS;

template <int OT>
__global__ void __launch_bounds__(PS_THREADS2, 2)
    process_samples_kernel(int O, int N, int T, const float* __restrict__ obs, const float* __restrict__ rew,
                           unsigned char* __restrict__ flags, const unsigned short* __restrict__ tstep,
                           const double* __restrict__ w, double discount, double gl, int drop_cut,
                           float* __restrict__ adv, float* __restrict__ ret, float* __restrict__ base,
                           double* __restrict__ partial_sum, double* __restrict__ partial_max) {
  __shared__ double sw[2 * OMAX + 4];
  __shared__ double scratch[B200RL_PS_NSUM * 32];
  __shared__ double s_bfirst[PS_WARPS][32];
  __shared__ double s_agg[PS_WARPS][3][32];
  __shared__ unsigned char s_st[PS_WARPS][32];     // bit0: no path end inside the chunk, bit1: local "dropped" state
  __shared__ double s_carry[4][32];                // window carry: adv, ret, undiscounted return, b of the next step
  __shared__ unsigned char s_mcarry[32];
  const bool have_w = (w != nullptr);
  if (have_w)
    for (int i = threadIdx.x; i < 2 * O + 4; i += blockDim.x) sw[i] = w[i];
  const int ln = threadIdx.x & 31, c = threadIdx.x >> 5;
  if (c == 0) {
    s_carry[0][ln] = 0.0; s_carry[1][ln] = 0.0; s_carry[2][ln] = 0.0; s_carry[3][ln] = 0.0;
    s_mcarry[ln] = 0;
  }
  __syncthreads();
  const int n = blockIdx.x * 32 + ln;
  const bool lane_ok = n < N;
  const size_t plane = (size_t)T * N;
  double s[B200RL_PS_NSUM];
  double m[B200RL_PS_NMAX];
#pragma unroll
  for (int i = 0; i < B200RL_PS_NSUM; ++i) s[i] = 0.0;
#pragma unroll
  for (int i = 0; i < B200RL_PS_NMAX; ++i) m[i] = -1.0e300;
  double gl8 = gl * gl, d8 = discount * discount;     // factor ^ PS_L
  gl8 *= gl8; gl8 *= gl8; d8 *= d8; d8 *= d8;
  static_assert(PS_L == 8, "the factor powers above assume 8 steps per chunk");

  for (int win_hi = T; win_hi > 0; win_hi -= PS_WIN) {
    const int t_hi = win_hi - 1 - c * PS_L;            // latest step of this thread's chunk
    // ---------------- P1: loads + baseline
    float rw[PS_L];
    double bs[PS_L];
    unsigned int endm = 0, startm = 0, cutm = 0;
#pragma unroll
    for (int u = 0; u < PS_L; ++u) {
      const int t = t_hi - u;
      const bool act = lane_ok && t >= 0;
      const size_t idx = (size_t)(act ? t : 0) * N + (lane_ok ? n : 0);
      const unsigned char f = act ? flags[idx] : (unsigned char)0;
      const unsigned short ts = act ? tstep[idx] : (unsigned short)1;
      rw[u] = act ? rew[idx] : 0.f;
      endm |= (unsigned)((f & B200RL_FLAG_END) ? 1 : 0) << u;
      cutm |= (unsigned)((f & B200RL_FLAG_CUT) ? 1 : 0) << u;
      startm |= (unsigned)(ts == 0 ? 1 : 0) << u;
      bs[u] = (have_w && act) ? lfb_predict<OT>(obs, plane, idx, O, ts, sw) : 0.0;
      if (act) base[idx] = (float)bs[u];
    }
    if (!drop_cut) cutm = 0;
    s_bfirst[c][ln] = bs[PS_L - 1];
    __syncthreads();
    // ---------------- P2: local scan (zero carry-in)
    const double b_in = (c == 0) ? s_carry[3][ln] : s_bfirst[c - 1][ln];
    {
      double a_n = 0.0, r_n = 0.0, u_n = 0.0, b_n = b_in;
      bool alive = true, dropped = false;
#pragma unroll
      for (int u = 0; u < PS_L; ++u) {
        if ((endm >> u) & 1u) { a_n = 0.0; r_n = 0.0; u_n = 0.0; b_n = 0.0; alive = false; dropped = (cutm >> u) & 1u; }
        const double r = (double)rw[u];
        a_n = (r + discount * b_n - bs[u]) + gl * a_n;
        r_n = r + discount * r_n;
        u_n = r + u_n;
        b_n = bs[u];
      }
      s_agg[c][0][ln] = a_n; s_agg[c][1][ln] = r_n; s_agg[c][2][ln] = u_n;
      s_st[c][ln] = (unsigned char)((alive ? 1 : 0) | (dropped ? 2 : 0));
    }
    __syncthreads();
    // ---------------- P3: true carry-in of this chunk
    double a_n = s_carry[0][ln], r_n = s_carry[1][ln], u_n = s_carry[2][ln];
    bool dropped = s_mcarry[ln] != 0;
    for (int cc = 0; cc < c; ++cc) {
      const unsigned char st = s_st[cc][ln];
      const bool alive = st & 1;
      a_n = s_agg[cc][0][ln] + (alive ? gl8 * a_n : 0.0);
      r_n = s_agg[cc][1][ln] + (alive ? d8 * r_n : 0.0);
      u_n = s_agg[cc][2][ln] + (alive ? u_n : 0.0);
      dropped = alive ? dropped : ((st & 2) != 0);
    }
    // ---------------- P4: second scan with the true carry, outputs + statistics
    {
      double b_n = b_in;
#pragma unroll
      for (int u = 0; u < PS_L; ++u) {
        const int t = t_hi - u;
        if (!(lane_ok && t >= 0)) continue;
        const size_t idx = (size_t)t * N + n;
        if ((endm >> u) & 1u) { a_n = 0.0; r_n = 0.0; u_n = 0.0; b_n = 0.0; dropped = (cutm >> u) & 1u; }
        const double r = (double)rw[u], b = bs[u];
        a_n = (r + discount * b_n - b) + gl * a_n;          // base.py:59-61, discount_cumsum(deltas, discount*lambda)
        r_n = r + discount * r_n;                            // discount_cumsum(rewards, discount)
        u_n = r + u_n;
        b_n = b;
        ret[idx] = (float)r_n;
        if (dropped) {
          adv[idx] = 0.f;
          flags[idx] = (unsigned char)(((endm >> u) & 1u ? B200RL_FLAG_END : 0) | ((cutm >> u) & 1u ? B200RL_FLAG_CUT : 0) |
                                       B200RL_FLAG_MASKED | (flags[idx] & B200RL_FLAG_DONE));
          continue;
        }
        adv[idx] = (float)a_n;
        // statistics use the float64 values (as the reference does)
        s[0] += a_n; s[1] += a_n * a_n; s[2] += 1.0;
        s[7] += r_n; s[8] += r_n * r_n; s[9] += b; s[10] += b * b;
        const double res = r_n - b;
        s[11] += res; s[12] += res * res;
        m[2] = fmax(m[2], -a_n); m[3] = fmax(m[3], a_n);
        if ((startm >> u) & 1u) {  // first sample of a path
          s[3] += 1.0; s[4] += r_n; s[5] += u_n; s[6] += u_n * u_n;
          m[0] = fmax(m[0], u_n); m[1] = fmax(m[1], -u_n);
        }
      }
    }
    // window carry-out: the chunk that holds the earliest steps of the window (read again only after the next
    // window's first barrier)
    if (c == PS_WARPS - 1) {
      s_carry[0][ln] = a_n; s_carry[1][ln] = r_n; s_carry[2][ln] = u_n; s_carry[3][ln] = bs[PS_L - 1];
      s_mcarry[ln] = dropped ? 1 : 0;
    }
  }
  __syncthreads();
  block_reduce_store<B200RL_PS_NSUM, false>(s, scratch, partial_sum + (size_t)blockIdx.x * B200RL_PS_NSUM);
  block_reduce_store<B200RL_PS_NMAX, true>(m, scratch, partial_max + (size_t)blockIdx.x * B200RL_PS_NMAX);
}

// (adv - mean) / (std + 1e-8), then optionally (adv - min) + 1e-8   (algos/util.py:7-12)
__global__ void center_adv_kernel(float* __restrict__ adv, long long B, const unsigned char* __restrict__ flags,
                                  const double* __restrict__ sums, const double* __restrict__ maxs, int center,
                                  int positive) {
  const double cnt = sums[2];
  const double mean = sums[0] / cnt;
  double var = sums[1] / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  const double stdv = sqrt(var) + 1e-8;
  double mn = -maxs[2];
  if (center) mn = (mn - mean) / stdv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += stride) {
    if (flags != nullptr && (flags[i] & B200RL_FLAG_MASKED)) continue;   // dropped path: adv stays 0
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
                    const float* __restrict__ ret, const unsigned char* __restrict__ flags,
                    double* __restrict__ partial) {
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
    if (sidx < B && !(flags != nullptr && (flags[sidx] & B200RL_FLAG_MASKED))) {
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
                                                           const float* __restrict__ ret,
                                                           const unsigned char* __restrict__ flags,
                                                           double* __restrict__ partial) {
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
    bool use[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long s = s0 + u * stride;
      const bool ok = s < B;
      const long long sl = ok ? s : s0;
      use[u] = ok && !(flags != nullptr && (flags[sl] & B200RL_FLAG_MASKED));
#pragma unroll
      for (int k = 0; k < O; ++k) raw[u][k] = obs[(size_t)k * B + sl];
      raw[u][O] = (float)tstep[sl];
      raw[u][O + 1] = ret[sl];
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (use[u]) {
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

int b200rl_process_samples(int obs_dim, int N, int T, const float* obs, const float* rew, unsigned char* flags,
                           const unsigned short* tstep, const double* w, double discount, double gae_lambda,
                           int drop_cut_paths, float* adv, float* ret, float* base, double* sums_out, double* maxs_out,
                           double* ws, void* stream) {
  B200RL_REQUIRE(obs && rew && flags && tstep && adv && ret && base && sums_out && maxs_out && ws,
                 "process_samples: null buffer");
  B200RL_REQUIRE(N > 0 && T > 0 && obs_dim > 0 && obs_dim <= OMAX, "process_samples: bad sizes");
  const int grid = (N + 31) / 32;          // one block per 32 lanes (PS_WARPS warps walk T in windows)
  B200RL_REQUIRE(grid <= 262144, "process_samples: too many lanes for one call (N <= 8M)");
  cudaStream_t st = (cudaStream_t)stream;
  double* psum = ws;
  double* pmax = ws + (size_t)grid * B200RL_PS_NSUM;
  B200RL_REQUIRE((long long)grid * (B200RL_PS_NSUM + B200RL_PS_NMAX) <= b200rl_ws_doubles(), "workspace too small");
#define B200RL_PS_LAUNCH(OT)                                                                                     \
  process_samples_kernel<OT><<<grid, PS_THREADS2, 0, st>>>(obs_dim, N, T, obs, rew, flags, tstep, w, discount,        \
                                                           discount * gae_lambda, drop_cut_paths, adv, ret, base, psum, \
                                                           pmax)
  switch (obs_dim) {                 // the obs dims of the compiled envs get an unrolled baseline predictor
    case 2: B200RL_PS_LAUNCH(2); break;
    case 3: B200RL_PS_LAUNCH(3); break;
    case 4: B200RL_PS_LAUNCH(4); break;
    case 13: B200RL_PS_LAUNCH(13); break;
    case 20: B200RL_PS_LAUNCH(20); break;
    default: B200RL_PS_LAUNCH(0); break;
  }
#undef B200RL_PS_LAUNCH
  B200RL_LAUNCH_CHECK("process_samples_kernel");
  int rc = launch_finalize_sum(psum, grid, B200RL_PS_NSUM, sums_out, 1.0, st);
  if (rc) return rc;
  return launch_finalize_max(pmax, grid, B200RL_PS_NMAX, maxs_out, st);
}

int b200rl_center_advantages(float* adv, long long B, const unsigned char* flags, const double* sums,
                             const double* maxs, int center, int positive, void* stream) {
  B200RL_REQUIRE(adv && sums && maxs && B > 0, "center_advantages: bad arguments");
  if (!center && !positive) return 0;
  long long blocks = (B + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  center_adv_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(adv, B, flags, sums, maxs, center,
                                                                         positive);
  B200RL_LAUNCH_CHECK("center_adv_kernel");
  return 0;
}

int b200rl_lfb_gram(int obs_dim, long long B, const float* obs, const unsigned short* tstep, const float* ret,
                    const unsigned char* flags, double* gram_out, double* ws, void* stream) {
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
      case 1: lfb_gram_reg_kernel<1><<<grid, 128, 0, st>>>(B, obs, tstep, ret, flags, ws); break;
      case 2: lfb_gram_reg_kernel<2><<<grid, 128, 0, st>>>(B, obs, tstep, ret, flags, ws); break;
      case 3: lfb_gram_reg_kernel<3><<<grid, 128, 0, st>>>(B, obs, tstep, ret, flags, ws); break;
      default: lfb_gram_reg_kernel<4><<<grid, 128, 0, st>>>(B, obs, tstep, ret, flags, ws); break;
    }
    B200RL_LAUNCH_CHECK("lfb_gram_reg_kernel");
    return launch_finalize_sum(ws, grid, npairs, gram_out, 1.0, st);
  }
  // smem <= 45 rows x 528 B = 23.8 KB: below the default 48 KB dynamic limit, no attribute needed (and a per-kernel
  // attribute set for a small obs_dim would cap a later, larger one)
  lfb_gram_kernel<<<grid, GRAM_THREADS, smem, st>>>(obs_dim, B, obs, tstep, ret, flags, ws);
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
