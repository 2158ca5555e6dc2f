// Policy-update passes over the sample batch: surrogate loss + KL, surrogate gradient, Fisher-vector product.
// One WARP per sample (lane j owns hidden unit j of each layer, +32 for 64-wide layers); 32 consecutive samples are
// loaded coalesced by the warp (one sample per lane) and broadcast with shuffles; hidden activations are exchanged
// through a per-warp shared-memory line (broadcast LDS.128); weight-gradient accumulators live in the owning lane's
// registers (float32 per 32-sample group, flushed to per-warp float64 shared-memory accumulators).
//
// Replaces the Theano functions compiled by rllab/optimizers/conjugate_gradient_optimizer.py:184-215 (f_loss, f_grad,
// f_loss_constraint), :22-55 (PerlmutterHvp f_Hx_plain), rllab/optimizers/first_order_optimizer.py:62-76 (grad part
// of f_opt), rllab/algos/vpg.py:100-103 (f_kl), over rllab/algos/npo.py:72-82 / vpg.py:88-99 and
// rllab/distributions/diagonal_gaussian.py:14-34,58-69.
#include <stdlib.h>

#include "update_common.cuh"

namespace b200rl {

constexpr int UPD_WARPS = 2;  // 64-wide nets only: 2 warps x P float64 accumulators fit 227 KB with params + tangent
constexpr int UPD_THREADS = UPD_WARPS * 32;
constexpr int FLUSH_GROUPS = 4;  // float32 register accumulators are folded into float64 every 4*32 samples


template <class N, int MODE>
struct UpdSmem {
  static constexpr int U = N::H1 / 32;
  static constexpr int W1P_LD = N::H2 + 4;
  // float region
  static constexpr int o_sp = 0;
  static constexpr int o_woT = o_sp + ((N::P + 3) & ~3);                          // [A][H2]
  static constexpr int o_w1p = o_woT + N::A * N::H2;                              // [H1][H2+4] (MODE>=1)
  static constexpr int o_sv = o_w1p + (MODE >= MODE_GRAD ? N::H1 * W1P_LD : 0);   // [P] (FVP)
  static constexpr int o_voT = o_sv + (MODE == MODE_FVP ? ((N::P + 3) & ~3) : 0); // [A][H2] (FVP)
  static constexpr int o_bufs = o_voT + (MODE == MODE_FVP ? N::A * N::H2 : 0);
  static constexpr int NBUF = (MODE == MODE_LOSS ? 2 : (MODE == MODE_GRAD ? 3 : 5));  // h1,h2,(d2),(t1,t2)
  static constexpr int buf_floats = NBUF * N::H1;
  static constexpr int n_floats = o_bufs + UPD_WARPS * buf_floats;
  static constexpr int acc_off_bytes = ((n_floats * 4 + 15) / 16) * 16;
  static constexpr int n_acc = (MODE >= MODE_GRAD ? UPD_WARPS * N::P : 0);
  static constexpr size_t bytes = (size_t)acc_off_bytes + (size_t)n_acc * 8 + 32 * 8 * 4;  // + block-reduce scratch
};

// canonical two-chain dot product of a broadcast shared-memory vector with a per-lane register column
template <int NIN>
__device__ __forceinline__ float dot_bcast_reg(const float* __restrict__ hb, const float (&col)[NIN], float bias) {
  float2 acc = make_float2(bias, 0.f);
#pragma unroll
  for (int i = 0; i < NIN; i += 4) {
    float4 h = *reinterpret_cast<const float4*>(hb + i);
    acc = ffma2(make_float2(h.x, h.y), make_float2(col[i], col[i + 1]), acc);
    acc = ffma2(make_float2(h.z, h.w), make_float2(col[i + 2], col[i + 3]), acc);
  }
  return acc.x + acc.y;
}
// same, column read from shared memory with stride `ld` (lane-consecutive -> conflict-free)
template <int NIN>
__device__ __forceinline__ float dot_bcast_smem(const float* __restrict__ hb, const float* __restrict__ colp, int ld,
                                                float bias) {
  float2 acc = make_float2(bias, 0.f);
#pragma unroll 8
  for (int i = 0; i < NIN; i += 4) {
    float4 h = *reinterpret_cast<const float4*>(hb + i);
    acc = ffma2(make_float2(h.x, h.y), make_float2(colp[i * ld], colp[(i + 1) * ld]), acc);
    acc = ffma2(make_float2(h.z, h.w), make_float2(colp[(i + 2) * ld], colp[(i + 3) * ld]), acc);
  }
  return acc.x + acc.y;
}

template <class N, int MODE>
__global__ void __launch_bounds__(UPD_THREADS) update_kernel(UpdArgs a) {
  using SM = UpdSmem<N, MODE>;
  constexpr int U = SM::U, O = N::O, A = N::A, H = N::H1, P = N::P;
  constexpr bool W1REG = (U == 1);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sf = reinterpret_cast<float*>(smem_raw);
  float* sp = sf + SM::o_sp;
  float* woT = sf + SM::o_woT;
  float* w1p = sf + SM::o_w1p;
  float* sv = sf + SM::o_sv;
  float* voT = sf + SM::o_voT;
  double* accw_all = reinterpret_cast<double*>(smem_raw + SM::acc_off_bytes);
  double* red_scratch = accw_all + SM::n_acc;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- stage parameters
  for (int i = threadIdx.x; i < P; i += blockDim.x) sp[i] = a.params[i];
  if constexpr (MODE == MODE_FVP)
    for (int i = threadIdx.x; i < P; i += blockDim.x) sv[i] = (float)a.xvec[i];
  __syncthreads();
  for (int i = threadIdx.x; i < A * H; i += blockDim.x) {
    int aa = i / H, j = i % H;
    woT[i] = sp[N::oWo + j * A + aa];
    if constexpr (MODE == MODE_FVP) voT[i] = sv[N::oWo + j * A + aa];
  }
  if constexpr (MODE >= MODE_GRAD) {
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) w1p[(i / H) * SM::W1P_LD + (i % H)] = sp[N::oW1 + i];
    for (int i = threadIdx.x; i < SM::n_acc; i += blockDim.x) accw_all[i] = 0.0;
  }
  __syncthreads();

  float* hb1 = sf + SM::o_bufs + warp * SM::buf_floats;
  float* hb2 = hb1 + H;
  float* db2 = hb2 + H;   // MODE >= GRAD
  float* tb1 = db2 + H;   // FVP
  float* tb2 = tb1 + H;   // FVP
  double* accw = accw_all + (size_t)warp * P;

  // ---- per-lane parameter registers (unit j = lane + 32u)
  float w0c[U][O], b0r[U], b1r[U], woutr[U][A];
  float w1c[W1REG ? H : 1];
  float v0c[MODE == MODE_FVP ? U : 1][O], vb0r[U], vb1r[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int j = lane + 32 * u;
#pragma unroll
    for (int o = 0; o < O; ++o) {
      w0c[u][o] = sp[N::oW0 + o * H + j];
      if constexpr (MODE == MODE_FVP) v0c[u][o] = sv[N::oW0 + o * H + j];
    }
    b0r[u] = sp[N::ob0 + j];
    b1r[u] = sp[N::ob1 + j];
    vb0r[u] = (MODE == MODE_FVP) ? sv[N::ob0 + j] : 0.f;
    vb1r[u] = (MODE == MODE_FVP) ? sv[N::ob1 + j] : 0.f;
#pragma unroll
    for (int k = 0; k < A; ++k) woutr[u][k] = sp[N::oWo + j * A + k];
  }
  if constexpr (W1REG) {
#pragma unroll
    for (int i = 0; i < H; ++i) w1c[i] = sp[N::oW1 + i * H + lane];
  }
  // distribution constants
  float ls_new[A], inv_std[A], var_new[A], var_new2[A], ls_old[A], inv_std_old[A], var_old[A], Mmu[A];
  float sum_ls_new = 0.f, sum_ls_old = 0.f;
#pragma unroll
  for (int k = 0; k < A; ++k) {
    ls_new[k] = clamp_log_std(sp[N::ols + k], a.log_min_std);
    float sd = expf(ls_new[k]);
    inv_std[k] = 1.0f / sd;
    var_new[k] = sd * sd;
    var_new2[k] = 2.0f * sd * sd + 1e-8f;
    ls_old[k] = (MODE == MODE_FVP) ? ls_new[k] : a.old_log_std[k];
    float so = expf(ls_old[k]);
    inv_std_old[k] = 1.0f / so;
    var_old[k] = so * so;
    Mmu[k] = 2.0f / var_new2[k];
    sum_ls_new += ls_new[k];
    sum_ls_old += ls_old[k];
  }
  const float half_log2pi_A = 0.5f * (float)A * 1.8378770664093453f;
  const int amin = lane < A ? lane : A - 1;

  // ---- accumulators
  float g_w0[U][O], g_b0[U], g_b1[U], g_wo[U][A];
  float g_w1[MODE >= MODE_GRAD ? H : 1];   // dW1[i][own unit]; for U==2 only the unit selected by a.unit_half
  float g_bo = 0.f, g_ls = 0.f;            // lanes < A: dbout[lane], dlog_std[lane]
#pragma unroll
  for (int u = 0; u < U; ++u) {
#pragma unroll
    for (int o = 0; o < O; ++o) g_w0[u][o] = 0.f;
    g_b0[u] = g_b1[u] = 0.f;
#pragma unroll
    for (int k = 0; k < A; ++k) g_wo[u][k] = 0.f;
  }
  if constexpr (MODE >= MODE_GRAD) {
#pragma unroll
    for (int i = 0; i < H; ++i) g_w1[i] = 0.f;
  }
  double s_loss = 0.0, s_kl = 0.0, m_kl = -1.0e300;  // lane-redundant; lane 0's copy is used

  auto flush = [&]() {
    if constexpr (MODE >= MODE_GRAD) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = lane + 32 * u;
#pragma unroll
        for (int o = 0; o < O; ++o) { accw[N::oW0 + o * H + j] += (double)g_w0[u][o]; g_w0[u][o] = 0.f; }
        accw[N::ob0 + j] += (double)g_b0[u]; g_b0[u] = 0.f;
        accw[N::ob1 + j] += (double)g_b1[u]; g_b1[u] = 0.f;
#pragma unroll
        for (int k = 0; k < A; ++k) { accw[N::oWo + j * A + k] += (double)g_wo[u][k]; g_wo[u][k] = 0.f; }
      }
      const int jw = lane + 32 * (U == 1 ? 0 : a.unit_half);
#pragma unroll
      for (int i = 0; i < H; ++i) { accw[N::oW1 + i * H + jw] += (double)g_w1[i]; g_w1[i] = 0.f; }
      if (lane < A) {
        accw[N::obo + lane] += (double)g_bo;
        accw[N::ols + lane] += (double)g_ls;
      }
      g_bo = 0.f; g_ls = 0.f;
    }
  };

  const long long ngroups = (a.B + 31) / 32;
  const long long gw = (long long)blockIdx.x * UPD_WARPS + warp;
  const long long nw = (long long)gridDim.x * UPD_WARPS;
  int since_flush = 0;
  for (long long g = gw; g < ngroups; g += nw) {
    const long long sidx = g * 32 + lane;
    const bool valid = sidx < a.B;
    const long long sl = valid ? sidx : a.B - 1;
    float xr[O], ar[A], mr[A], advr = 0.f;
#pragma unroll
    for (int o = 0; o < O; ++o) xr[o] = a.obs[(size_t)o * a.B + sl];
    if constexpr (MODE != MODE_FVP) {
#pragma unroll
      for (int k = 0; k < A; ++k) {
        ar[k] = a.act[(size_t)k * a.B + sl];
        mr[k] = a.old_mean[(size_t)k * a.B + sl];
      }
      advr = a.adv[sl];
    }
    const int cnt = (int)((a.B - g * 32) < 32 ? (a.B - g * 32) : 32);
    for (int k32 = 0; k32 < cnt; ++k32) {
      asm volatile("" ::: "memory");
      float x[O];
#pragma unroll
      for (int o = 0; o < O; ++o) x[o] = __shfl_sync(0xffffffffu, xr[o], k32);
      // ---- forward layer 1 (canonical order: even / odd input chains)
      float h1v[U], h2v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float s0 = b0r[u], s1 = 0.f;
#pragma unroll
        for (int o = 0; o < O; ++o) {
          if ((o & 1) == 0) s0 = fmaf(x[o], w0c[u][o], s0);
          else s1 = fmaf(x[o], w0c[u][o], s1);
        }
        h1v[u] = tanh_f(s0 + s1);
        hb1[lane + 32 * u] = h1v[u];
      }
      __syncwarp();
      // ---- forward layer 2
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float pre;
        if constexpr (W1REG) pre = dot_bcast_reg<H>(hb1, w1c, b1r[u]);
        else pre = dot_bcast_smem<H>(hb1, sp + N::oW1 + lane + 32 * u, H, b1r[u]);
        h2v[u] = tanh_f(pre);
        hb2[lane + 32 * u] = h2v[u];
      }
      __syncwarp();
      // ---- mean: lanes < A own one action dimension each (others duplicate lane A-1), then broadcast
      float mu[A];
      {
        float m_own = dot_bcast_smem<H>(hb2, woT + amin * H, 1, sp[N::obo + amin]);
#pragma unroll
        for (int k = 0; k < A; ++k) mu[k] = __shfl_sync(0xffffffffu, m_own, k);
      }
      float dmu[A];
      float w_s = 0.f;
      if constexpr (MODE != MODE_FVP) {
        // ---- distribution math (lane-redundant)
        float z[A], zsq = 0.f, zsq_old = 0.f, kl = 0.f;
        const float adv_s = __shfl_sync(0xffffffffu, advr, k32);
#pragma unroll
        for (int k = 0; k < A; ++k) {
          const float act = __shfl_sync(0xffffffffu, ar[k], k32);
          const float om = __shfl_sync(0xffffffffu, mr[k], k32);
          z[k] = (act - mu[k]) * inv_std[k];
          zsq += z[k] * z[k];
          const float zo = (act - om) * inv_std_old[k];
          zsq_old += zo * zo;
          const float dm = om - mu[k];
          kl += (dm * dm + var_old[k] - var_new[k]) / var_new2[k] + ls_new[k] - ls_old[k];
        }
        const float logp_new = -sum_ls_new - 0.5f * zsq - half_log2pi_A;
        float term;
        if (a.loss_kind == B200RL_LOSS_TRPO) {
          const float logp_old = -sum_ls_old - 0.5f * zsq_old - half_log2pi_A;
          const float lr = expf(logp_new - logp_old);
          w_s = lr * adv_s;
          term = -w_s;
        } else {
          w_s = adv_s;
          term = -logp_new * adv_s;
        }
        s_loss += (double)term;
        s_kl += (double)kl;
        m_kl = fmax(m_kl, (double)kl);
        if constexpr (MODE == MODE_GRAD) {
#pragma unroll
          for (int k = 0; k < A; ++k) {
            dmu[k] = -w_s * z[k] * inv_std[k];
            if (lane == k) { g_bo += dmu[k]; g_ls += -w_s * (z[k] * z[k] - 1.0f); }
          }
        }
      }
      if constexpr (MODE == MODE_FVP) {
        // ---- tangent forward: J x
        float t1v[U], t2v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float s0 = vb0r[u], s1 = 0.f;
#pragma unroll
          for (int o = 0; o < O; ++o) {
            if ((o & 1) == 0) s0 = fmaf(x[o], v0c[u][o], s0);
            else s1 = fmaf(x[o], v0c[u][o], s1);
          }
          t1v[u] = (1.0f - h1v[u] * h1v[u]) * (s0 + s1);
          tb1[lane + 32 * u] = t1v[u];
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float p2;
          if constexpr (W1REG) p2 = dot_bcast_reg<H>(tb1, w1c, vb1r[u]);
          else p2 = dot_bcast_smem<H>(tb1, sp + N::oW1 + lane + 32 * u, H, vb1r[u]);
          p2 += dot_bcast_smem<H>(hb1, sv + N::oW1 + lane + 32 * u, H, 0.f);
          t2v[u] = (1.0f - h2v[u] * h2v[u]) * p2;
          tb2[lane + 32 * u] = t2v[u];
        }
        __syncwarp();
        float md = dot_bcast_smem<H>(tb2, woT + amin * H, 1, sv[N::obo + amin]);
        md += dot_bcast_smem<H>(hb2, voT + amin * H, 1, 0.f);
#pragma unroll
        for (int k = 0; k < A; ++k) {
          dmu[k] = __shfl_sync(0xffffffffu, md, k) * Mmu[k];
          if (lane == k) g_bo += dmu[k];
        }
      }
      if constexpr (MODE >= MODE_GRAD) {
        // ---- backward
        float d2v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < A; ++k) {
            s = fmaf(dmu[k], woutr[u][k], s);
            g_wo[u][k] = fmaf(h2v[u], dmu[k], g_wo[u][k]);
          }
          d2v[u] = s * (1.0f - h2v[u] * h2v[u]);
          g_b1[u] += d2v[u];
          db2[lane + 32 * u] = d2v[u];
        }
        __syncwarp();
        // dW1[i][own unit] += h1[i] * d2[own unit]
        {
          const float d2own = (U == 1) ? d2v[0] : (a.unit_half == 0 ? d2v[0] : d2v[U - 1]);
#pragma unroll
          for (int i = 0; i < H; i += 4) {
            float4 h = *reinterpret_cast<const float4*>(hb1 + i);
            g_w1[i] = fmaf(h.x, d2own, g_w1[i]);
            g_w1[i + 1] = fmaf(h.y, d2own, g_w1[i + 1]);
            g_w1[i + 2] = fmaf(h.z, d2own, g_w1[i + 2]);
            g_w1[i + 3] = fmaf(h.w, d2own, g_w1[i + 3]);
          }
        }
        // d1[i] = (sum_j d2[j] W1[i][j]) (1 - h1[i]^2) for own units i
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float* row = w1p + (lane + 32 * u) * SM::W1P_LD;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int j = 0; j < H; j += 4) {
            float4 d = *reinterpret_cast<const float4*>(db2 + j);
            float4 w = *reinterpret_cast<const float4*>(row + j);
            s0 = fmaf(d.x, w.x, s0); s1 = fmaf(d.y, w.y, s1); s2 = fmaf(d.z, w.z, s2); s3 = fmaf(d.w, w.w, s3);
          }
          const float d1 = ((s0 + s1) + (s2 + s3)) * (1.0f - h1v[u] * h1v[u]);
          g_b0[u] += d1;
#pragma unroll
          for (int o = 0; o < O; ++o) g_w0[u][o] = fmaf(x[o], d1, g_w0[u][o]);
        }
      }
      __syncwarp();
    }
    if (MODE >= MODE_GRAD && ++since_flush == FLUSH_GROUPS) {
      flush();
      since_flush = 0;
    }
  }
  if constexpr (MODE >= MODE_GRAD) flush();
  __syncthreads();

  // ---- block-level combine (fixed order over warps) and write this block's partial vector
  if constexpr (MODE >= MODE_GRAD) {
    double* out = a.partial + (size_t)blockIdx.x * P;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < UPD_WARPS; ++w) s += accw_all[(size_t)w * P + p];
      out[p] = s;
    }
  }
  if constexpr (MODE != MODE_FVP) {
    // loss/KL scalars: lane 0 of each warp holds the warp's sums
    double v[2] = {lane == 0 ? s_loss : 0.0, lane == 0 ? s_kl : 0.0};
    double mx[1] = {m_kl};
    const size_t K = (MODE == MODE_GRAD) ? (size_t)P + 3 : 3;
    double* base = (MODE == MODE_GRAD) ? a.partial + (size_t)gridDim.x * P : a.partial;
    block_reduce_store<2, false>(v, red_scratch, base + (size_t)blockIdx.x * 3);
    block_reduce_store<1, true>(mx, red_scratch, base + (size_t)blockIdx.x * 3 + 2);
    (void)K;
  }
}

// Surrogate loss + KL, one THREAD per sample (forward only: no cross-sample reduction of per-weight quantities, so
// the thread-per-lane forward of the rollout kernel is the cheapest formulation; same canonical summation order).
#ifdef B200RL_CONST_WEIGHTS
B200RL_DEFINE_CONST_THETA
#endif
constexpr int LOSS_THREADS = 128;
template <class N>
constexpr int loss_minblocks() { return (N::H1 == 32 && N::O <= 4) ? 4 : 1; }   // 128 registers: 1.47 -> 1.28 ms (A/B)
template <class N>
__global__ void __launch_bounds__(LOSS_THREADS, loss_minblocks<N>()) loss_thread_kernel(UpdArgs a) {
  constexpr int O = N::O, A = N::A;
#ifdef B200RL_CONST_WEIGHTS
  const float* sp = c_theta;
  __shared__ double red_scratch[3 * 32];
#else
  __shared__ __align__(16) float sp[N::P];
  __shared__ double red_scratch[3 * 32];
  for (int i = threadIdx.x; i < N::P; i += blockDim.x) sp[i] = a.params[i];
  __syncthreads();
#endif
  float ls_new[A], inv_std[A], var_new[A], var_new2[A], ls_old[A], inv_std_old[A], var_old[A];
  float sum_ls_new = 0.f, sum_ls_old = 0.f;
#pragma unroll
  for (int k = 0; k < A; ++k) {
    ls_new[k] = clamp_log_std(sp[N::ols + k], a.log_min_std);
    const float sd = expf(ls_new[k]);
    inv_std[k] = 1.0f / sd;
    var_new[k] = sd * sd;
    var_new2[k] = 2.0f * sd * sd + 1e-8f;
    ls_old[k] = a.old_log_std[k];
    const float so = expf(ls_old[k]);
    inv_std_old[k] = 1.0f / so;
    var_old[k] = so * so;
    sum_ls_new += ls_new[k];
    sum_ls_old += ls_old[k];
  }
  const float half_log2pi_A = 0.5f * (float)A * 1.8378770664093453f;
  double s_loss = 0.0, s_kl = 0.0, m_kl = -1.0e300;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < a.B; s += stride) {
    asm volatile("" ::: "memory");
    float x[O], h1[N::H1], h2[N::H2], mu[A];
#pragma unroll
    for (int o = 0; o < O; ++o) x[o] = a.obs[(size_t)o * a.B + s];
    mlp_forward_thread<N>(sp, x, h1, h2, mu);
    float zsq = 0.f, zsq_old = 0.f, kl = 0.f;
#pragma unroll
    for (int k = 0; k < A; ++k) {
      const float act = a.act[(size_t)k * a.B + s];
      const float om = a.old_mean[(size_t)k * a.B + s];
      const float z = (act - mu[k]) * inv_std[k];
      zsq += z * z;
      const float zo = (act - om) * inv_std_old[k];
      zsq_old += zo * zo;
      const float dm = om - mu[k];
      kl += (dm * dm + var_old[k] - var_new[k]) / var_new2[k] + ls_new[k] - ls_old[k];
    }
    const float adv_s = a.adv[s];
    const float logp_new = -sum_ls_new - 0.5f * zsq - half_log2pi_A;
    float term;
    if (a.loss_kind == B200RL_LOSS_TRPO) {
      const float logp_old = -sum_ls_old - 0.5f * zsq_old - half_log2pi_A;
      term = -expf(logp_new - logp_old) * adv_s;
    } else {
      term = -logp_new * adv_s;
    }
    s_loss += (double)term;
    s_kl += (double)kl;
    m_kl = fmax(m_kl, (double)kl);
  }
  double v[2] = {s_loss, s_kl};
  double mx[1] = {m_kl};
  block_reduce_store<2, false>(v, red_scratch, a.partial + (size_t)blockIdx.x * 3);
  block_reduce_store<1, true>(mx, red_scratch, a.partial + (size_t)blockIdx.x * 3 + 2);
}

// Hx += diag_scale * (reg * x  (+)  M_l x_l on the un-clamped log_std entries)
__global__ void fvp_diag_kernel(int P, int ols, int A, const float* __restrict__ params, float log_min_std,
                                const double* __restrict__ x, double reg, double diag_scale, double* __restrict__ Hx) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double add = reg * x[p];
  if (p >= ols && p < ols + A) {
    const float par = params[p];
    if (par > log_min_std) {
      const double s = exp(2.0 * (double)fmaxf(par, log_min_std));
      const double eps = 1e-8;
      add += 4.0 * s * (2.0 * s - eps) / ((2.0 * s + eps) * (2.0 * s + eps)) * x[p];
    }
  }
  Hx[p] += diag_scale * add;
}

// zero the log_std gradient where the min_std clamp is active (TT.maximum routes the gradient to the constant)
__global__ void mask_logstd_grad_kernel(int ols, int A, const float* __restrict__ params, float log_min_std,
                                        double* __restrict__ g) {
  const int k = threadIdx.x;
  if (k < A && !(params[ols + k] > log_min_std)) g[ols + k] = 0.0;
}

template <class N, int MODE>
static int launch_update(const UpdArgs& a0, int grid, cudaStream_t st) {
  using SM = UpdSmem<N, MODE>;
  static bool attr_done = false;
  if (!attr_done) {
    B200RL_CUDA_CHECK(cudaFuncSetAttribute(update_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)SM::bytes));
    attr_done = true;
  }
  update_kernel<N, MODE><<<grid, UPD_THREADS, SM::bytes, st>>>(a0);
  B200RL_LAUNCH_CHECK("update_kernel");
  return 0;
}

int update_impl() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("B200RL_UPDATE_IMPL");
    // default (auto, -1): tile kernel for 32-wide nets (4 % faster there), GEMM kernel for 64-wide nets
    cached = (e == nullptr || !strcmp(e, "auto")) ? 3 : (!strcmp(e, "gemm") ? 0 : (!strcmp(e, "tile") ? 1 : 2));
  }
  return cached;
}

template <class N, int MODE>
static int update_grid(long long B) {
  using SM = UpdSmem<N, MODE>;
  int per_sm = (int)((227 * 1024) / (SM::bytes + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  long long grid = (long long)num_sms() * per_sm;
  const long long ngroups = (B + 31) / 32;
  const long long need = (ngroups + UPD_WARPS - 1) / UPD_WARPS;
  if (grid > need) grid = need;
  if (grid > MAX_PARTIAL_BLOCKS) grid = MAX_PARTIAL_BLOCKS;
  if (grid < 1) grid = 1;
  return (int)grid;
}

}  // namespace b200rl

using namespace b200rl;

extern "C" {

int b200rl_loss_kl(int loss_kind, const float* params_f32, int obs_dim, int h1, int h2, int act_dim, float min_std,
                   long long B, const float* obs, const float* act, const float* adv, const float* old_mean,
                   const float* old_log_std, double scale, double* out, double* ws, void* stream) {
  B200RL_REQUIRE(params_f32 && obs && act && adv && old_mean && old_log_std && out && ws && B > 0,
                 "loss_kl: bad arguments");
  B200RL_REQUIRE(loss_kind == B200RL_LOSS_TRPO || loss_kind == B200RL_LOSS_VPG, "loss_kl: bad loss kind");
  cudaStream_t st = (cudaStream_t)stream;
  UpdArgs a{};
  a.params = params_f32; a.log_min_std = min_std > 0.f ? logf(min_std) : -INFINITY; a.B = B;
  a.obs = obs; a.act = act; a.adv = adv; a.old_mean = old_mean; a.old_log_std = old_log_std;
  a.loss_kind = loss_kind; a.partial = ws;
  int grid = 0;
  {
    long long g = (long long)num_sms() * 4;
    const long long need = (B + LOSS_THREADS - 1) / LOSS_THREADS;
    if (g > need) g = need;
    if (g > MAX_PARTIAL_BLOCKS) g = MAX_PARTIAL_BLOCKS;
    grid = (int)g;
  }
#ifdef B200RL_CONST_WEIGHTS
  B200RL_DISPATCH_NET({
    int rc_up = upload_theta(params_f32, NetT::P, st);
    if (rc_up) return rc_up;
  });
#endif
  B200RL_DISPATCH_NET({ loss_thread_kernel<NetT><<<grid, LOSS_THREADS, 0, st>>>(a); });
  B200RL_LAUNCH_CHECK("loss_thread_kernel");
  // partial layout [grid][3] = (sum loss, sum kl, max kl): strided finalize
  int rc = launch_finalize_sum(ws, grid, 3, out, scale, st);  // out[2] is overwritten below
  if (rc) return rc;
  // max over blocks of column 2
  // (reuse finalize_max on the same [grid][3] layout, then keep column 2)
  double* tmp = ws + (size_t)grid * 3;
  rc = launch_finalize_max(ws, grid, 3, tmp, st);
  if (rc) return rc;
  B200RL_CUDA_CHECK(cudaMemcpyAsync(out + 2, tmp + 2, sizeof(double), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int b200rl_grad(int loss_kind, const float* params_f32, int obs_dim, int h1, int h2, int act_dim, float min_std,
                long long B, const float* obs, const float* act, const float* adv, const float* old_mean,
                const float* old_log_std, double scale, double* g_out, double* loss_out, float* h_cache_out, double* ws,
                void* stream) {
  B200RL_REQUIRE(params_f32 && obs && act && adv && old_mean && old_log_std && g_out && ws && B > 0,
                 "grad: bad arguments");
  B200RL_REQUIRE(loss_kind == B200RL_LOSS_TRPO || loss_kind == B200RL_LOSS_VPG, "grad: bad loss kind");
  cudaStream_t st = (cudaStream_t)stream;
  UpdArgs a{};
  a.params = params_f32; a.log_min_std = min_std > 0.f ? logf(min_std) : -INFINITY; a.B = B;
  a.obs = obs; a.act = act; a.adv = adv; a.old_mean = old_mean; a.old_log_std = old_log_std;
  a.loss_kind = loss_kind; a.partial = ws; a.h_cache = h_cache_out;
  int grid = 0, P = 0, ols = 0;
  int impl = update_impl();
  if (impl == 3) impl = (h1 == 32) ? 1 : 0;
  if (h1 == h2 && (impl == 0 || (impl == 1 && h1 == 32))) {
    int rc = (impl == 0) ? update_gemm_launch(MODE_GRAD, obs_dim, h1, act_dim, a, &grid, &P, &ols, st)
                         : update_tile_launch(MODE_GRAD, obs_dim, act_dim, a, &grid, &P, &ols, st);
    if (rc) return rc;
    rc = launch_finalize_sum(ws, grid, P, g_out, scale, st);
    if (rc) return rc;
  } else {
    B200RL_DISPATCH_NET_H(64, {
      grid = update_grid<NetT, MODE_GRAD>(B);
      P = NetT::P; ols = NetT::ols;
      constexpr int U = NetT::H1 / 32;
      for (int half = 0; half < U; ++half) {
        a.unit_half = half;
        // 64-wide nets: the second pass recomputes everything but only its dW1 half differs; both passes write the
        // full vector, the halves are merged below.
        a.partial = ws + (size_t)half * ((size_t)grid * (NetT::P + 3));
        int rc = launch_update<NetT, MODE_GRAD>(a, grid, st);
        if (rc) return rc;
        rc = launch_finalize_sum(a.partial, grid, NetT::P,
                                 half == 0 ? g_out : ws + 2 * ((size_t)grid * (NetT::P + 3)), scale, st);
        if (rc) return rc;
      }
      if (U == 2) {
        // take dW1 columns 32..63 from the second pass
        const double* g2 = ws + 2 * ((size_t)grid * (NetT::P + 3));
        B200RL_CUDA_CHECK(cudaMemcpy2DAsync(g_out + NetT::oW1 + 32, NetT::H2 * sizeof(double), g2 + NetT::oW1 + 32,
                                            NetT::H2 * sizeof(double), 32 * sizeof(double), NetT::H1,
                                            cudaMemcpyDeviceToDevice, st));
      }
    });
  }
  mask_logstd_grad_kernel<<<1, 32, 0, st>>>(ols, act_dim, params_f32, a.log_min_std, g_out);
  B200RL_LAUNCH_CHECK("mask_logstd_grad_kernel");
  if (loss_out != nullptr) {
    // per-block (sum loss, sum kl, max kl) triples follow the first pass's [grid][P] partial vectors
    const double* sc = ws + (size_t)grid * P;
    double* tmp = ws + (size_t)grid * (P + 3) * 3 + P + 8;
    int rc = launch_finalize_sum(sc, grid, 3, loss_out, scale, st);
    if (rc) return rc;
    rc = launch_finalize_max(sc, grid, 3, tmp, st);
    if (rc) return rc;
    B200RL_CUDA_CHECK(cudaMemcpyAsync(loss_out + 2, tmp + 2, sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int b200rl_fvp(const float* params_f32, int obs_dim, int h1, int h2, int act_dim, float min_std, long long B,
               const float* obs, const double* x, double scale, double reg_coeff, double diag_scale, double* Hx_out,
               const float* h_cache, double* ws, void* stream) {
  B200RL_REQUIRE(params_f32 && obs && x && Hx_out && ws && B > 0, "fvp: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  UpdArgs a{};
  a.params = params_f32; a.xvec = x; a.log_min_std = min_std > 0.f ? logf(min_std) : -INFINITY; a.B = B;
  a.obs = obs; a.partial = ws; a.h_cache = const_cast<float*>(h_cache);
  int grid = 0, P = 0, ols = 0;
  int impl = update_impl();
  if (impl == 3) impl = (h1 == 32) ? 1 : 0;
  if (h1 == h2 && (impl == 0 || (impl == 1 && h1 == 32))) {
    int rc = (impl == 0) ? update_gemm_launch(MODE_FVP, obs_dim, h1, act_dim, a, &grid, &P, &ols, st)
                         : update_tile_launch(MODE_FVP, obs_dim, act_dim, a, &grid, &P, &ols, st);
    if (rc) return rc;
    rc = launch_finalize_sum(ws, grid, P, Hx_out, scale, st);
    if (rc) return rc;
  } else {
    B200RL_DISPATCH_NET_H(64, {
      grid = update_grid<NetT, MODE_FVP>(B);
      P = NetT::P; ols = NetT::ols;
      constexpr int U = NetT::H1 / 32;
      for (int half = 0; half < U; ++half) {
        a.unit_half = half;
        a.partial = ws + (size_t)half * ((size_t)grid * NetT::P);
        int rc = launch_update<NetT, MODE_FVP>(a, grid, st);
        if (rc) return rc;
        rc = launch_finalize_sum(a.partial, grid, NetT::P, half == 0 ? Hx_out : ws + 2 * ((size_t)grid * NetT::P),
                                 scale, st);
        if (rc) return rc;
      }
      if (U == 2) {
        const double* g2 = ws + 2 * ((size_t)grid * NetT::P);
        B200RL_CUDA_CHECK(cudaMemcpy2DAsync(Hx_out + NetT::oW1 + 32, NetT::H2 * sizeof(double), g2 + NetT::oW1 + 32,
                                            NetT::H2 * sizeof(double), 32 * sizeof(double), NetT::H1,
                                            cudaMemcpyDeviceToDevice, st));
      }
    });
  }
  // the log_std slot of the sample sum is zero (mean does not depend on log_std); add reg*x and the M_l block
  fvp_diag_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, ols, act_dim, params_f32, a.log_min_std, x, reg_coeff,
                                                   diag_scale, Hx_out);
  B200RL_LAUNCH_CHECK("fvp_diag_kernel");
  return 0;
}
}
