// Fisher-vector product for 32-wide policies with the per-sample chain in FLOAT64 ("chain64"): the default product of the
// TRPO conjugate-gradient solve.
//
// Why: the reference evaluates its Hessian-vector product in float64 (Theano floatX default).  The CG recursion on
// F + 1e-5 I (condition number 1e6 .. 1e7 once the policy has sharpened) needs the operator it is given to be symmetric and
// linear far below float32 resolution: with the float32 kernels (FFMA or 3xTF32 alike) the tangent-forward chain J p and the
// backward chain J^T u round differently, the computed operator is J_b^T M J_f with |J_b - J_f| ~ 1e-7 |J|, and that defect,
// divided by the small eigenvalues, costs the 10-iteration solve about three iterations of depth.  On Swimmer that is the
// difference between 23.7 +- 2.7 and 30.5 +- 1.5 AverageReturn at iteration 40 (8 sampler seeds; float64 oracle 31.6;
// scripts/exp_seed_sweep.py) -- the learning speed of TRPO on that task follows the effective depth of
// the CG solve (6 float64 iterations: 21.1; 20 float32 iterations: 36.0).  Neither a two-word CG direction nor an unbiased
// TF32 split changes it; evaluating the chain in float64 does.
//
// What runs in float64: everything that depends on the direction p through the network -- the tangent forward
// t1 = (x V0 + vb0)(1-h1^2), t2 = (t1 W1 + h1 V1 + vb1)(1-h2^2), mu_dot = t2 Wout + h2 Vout + vbout, dmu = M mu_dot, and the
// backward d2 = (dmu Wout^T)(1-h2^2), d1 = (d2 W1^T)(1-h1^2): 3 600 DFMA per sample (B200 issues DFMA at half the FFMA
// rate), one thread per sample, theta and p as float64 in shared memory.  The activations h1, h2 are the float32 values the
// gradient pass cached: the same values in both chains, i.e. a consistent operator.  What stays float32: the sample-axis
// Gram products dW = sum_s a_s (x) d_s (tile_gram.cuh) on the float32-rounded d1 / d2 / dmu -- those roundings are
// independent from sample to sample and average out over the batch; their sums are float64 as everywhere.
//
// Replaces f_Hx_plain of rllab/optimizers/conjugate_gradient_optimizer.py:22-55 (PerlmutterHvp) for hidden (32,32).
#include "tile_gram.cuh"

namespace b200rl {

constexpr int C_THREADS = 128, C_TILE = 128, C_LD = C_TILE + 4;
// compiler barrier per weight row: without it ptxas hoists the (loop-invariant) shared-memory loads of all weights ahead of
// the fully unrolled layers and spills 8 KB per thread
#define C64_FENCE() asm volatile("" ::: "memory")

template <class N>
struct Chain64Smem {
  static constexpr int O = N::O, H = 32, A = N::A;
  static_assert(N::H1 == 32 && N::H2 == 32, "chain64 kernel is specialised for 32-wide layers");
  static constexpr int P2 = (N::P + 1) & ~1;                       // doubles per parameter vector (16 B aligned)
  static constexpr int rX = 0, rH1 = rX + O, rH2 = rH1 + H, rD1 = rH2 + H, rD2 = rD1 + H, rDM = rD2 + H, rDL = rDM + A,
                       R = rDL + A;
  static constexpr size_t o_sp = 0, o_sv = (size_t)P2 * 8, o_stage = 2 * (size_t)P2 * 8;
  static constexpr size_t bytes = o_stage + (size_t)R * C_LD * 4;
  static_assert(2 * 64 * 16 * 8 <= R * C_LD * 4, "stage region must hold the K-half combine scratch");
};

template <class N>
__global__ void __launch_bounds__(C_THREADS, 2) fvp_chain64_kernel(UpdArgs a) {
  using SM = Chain64Smem<N>;
  constexpr int O = N::O, H = 32, A = N::A, P = N::P, LD = C_LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sp = reinterpret_cast<double*>(smem_raw + SM::o_sp);      // theta (the float32 shadow, widened: the rollout's theta)
  double* sv = reinterpret_cast<double*>(smem_raw + SM::o_sv);      // direction p
  float* stage = reinterpret_cast<float*>(smem_raw + SM::o_stage);
  const int tid = threadIdx.x;
  for (int i = tid; i < P; i += C_THREADS) {
    sp[i] = (double)a.params[i];
    sv[i] = a.xvec[i];
  }
  __syncthreads();
  double Mmu[A];
#pragma unroll
  for (int k = 0; k < A; ++k) {
    const double ls = fmax(sp[N::ols + k], (double)a.log_min_std);
    const double s2 = exp(2.0 * ls);
    Mmu[k] = 2.0 / (2.0 * s2 + 1e-8);
  }
  TileGram<N, SM::rX, SM::rH1, SM::rH2, SM::rD1, SM::rD2, SM::rDM, LD, true> gram;
  gram.init();
  float* colX = stage + SM::rX * LD + tid;
  float* colH1 = stage + SM::rH1 * LD + tid;
  float* colH2 = stage + SM::rH2 * LD + tid;
  float* colD1 = stage + SM::rD1 * LD + tid;
  float* colD2 = stage + SM::rD2 * LD + tid;
  float* colDM = stage + SM::rDM * LD + tid;
  float* colDL = stage + SM::rDL * LD + tid;

  const long long ntiles = n_tiles_of(a, C_TILE);
  for (long long ti_ = blockIdx.x; ti_ < ntiles; ti_ += gridDim.x) {
    asm volatile("" ::: "memory");
    const long long s = tile_at(a, ti_) * C_TILE + tid;
    const bool valid = sample_valid(a, s);
    const long long sl = s < a.B ? s : a.B - 1;
    // ---- loads: observations and the activations cached by the gradient pass (float32)
    // (staged to this thread's own shared-memory column at once and re-read from there: 64 registers less to keep live)
    float x[O];
#pragma unroll
    for (int o = 0; o < O; ++o) {
      x[o] = a.obs[(size_t)o * a.B + sl];
      colX[o * LD] = x[o];
    }
    {
      const float* hc = a.h_cache + sl;
      float hv[2 * H];
#pragma unroll
      for (int j = 0; j < 2 * H; ++j) hv[j] = hc[(size_t)j * a.B];
#pragma unroll
      for (int j = 0; j < H; ++j) {
        colH1[j * LD] = hv[j];
        colH2[j * LD] = hv[H + j];
      }
    }
    // ---- tangent forward (float64): t1 = (x V0 + vb0)(1 - h1^2)
    double t[H];
#pragma unroll
    for (int j = 0; j < H; ++j) t[j] = sv[N::ob0 + j];
#pragma unroll
    for (int o = 0; o < O; ++o) {
      C64_FENCE();
      const double xd = (double)x[o];
#pragma unroll
      for (int j = 0; j < H; ++j) t[j] = fma(xd, sv[N::oW0 + o * H + j], t[j]);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const double h1d = (double)colH1[j * LD];
      t[j] *= 1.0 - h1d * h1d;
    }
    // t2 = (t1 W1 + h1 V1 + vb1)(1 - h2^2)
    double u[H];
#pragma unroll
    for (int j = 0; j < H; ++j) u[j] = sv[N::ob1 + j];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      C64_FENCE();
      const double hd = (double)colH1[i * LD], ti = t[i];
#pragma unroll
      for (int j = 0; j < H; ++j) u[j] = fma(hd, sv[N::oW1 + i * H + j], fma(ti, sp[N::oW1 + i * H + j], u[j]));
    }
    double dmu[A];
#pragma unroll
    for (int k = 0; k < A; ++k) dmu[k] = sv[N::obo + k];
    C64_FENCE();
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const double h2d = (double)colH2[j * LD];
      u[j] *= 1.0 - h2d * h2d;
#pragma unroll
      for (int k = 0; k < A; ++k) dmu[k] = fma(u[j], sp[N::oWo + j * A + k], fma(h2d, sv[N::oWo + j * A + k], dmu[k]));
    }
#pragma unroll
    for (int k = 0; k < A; ++k) {
      dmu[k] = valid ? dmu[k] * Mmu[k] : 0.0;
      colDM[k * LD] = (float)dmu[k];
      colDL[k * LD] = 0.f;
    }
    // ---- backward (float64): d2 = (dmu Wout^T)(1 - h2^2); d1 = (d2 W1^T)(1 - h1^2)
#pragma unroll
    for (int j = 0; j < H; ++j) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k < A; ++k) sacc = fma(dmu[k], sp[N::oWo + j * A + k], sacc);
      const double h2d = (double)colH2[j * LD];
      u[j] = sacc * (1.0 - h2d * h2d);                    // d2 (u is free)
      colD2[j * LD] = (float)u[j];
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
      C64_FENCE();
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int j = 0; j < H; j += 2) {
        s0 = fma(u[j], sp[N::oW1 + i * H + j], s0);
        s1 = fma(u[j + 1], sp[N::oW1 + i * H + j + 1], s1);
      }
      const double h1d = (double)colH1[i * LD];
      colD1[i * LD] = (float)((s0 + s1) * (1.0 - h1d * h1d));
    }
    __syncthreads();
    // ---- Gram products over the tile (float32 products of the rounded per-sample vectors, float64 across tiles)
    gram.accumulate_a(stage, tid);
    gram.accumulate_b(stage, tid);
    __syncthreads();
  }
  double* out = a.partial + (size_t)blockIdx.x * P;
  gram.write(out, reinterpret_cast<double*>(stage), tid);
}

template <class N>
static int launch_chain64(const UpdArgs& a, int* grid_out, cudaStream_t st) {
  using SM = Chain64Smem<N>;
  B200RL_SET_MAX_SMEM((fvp_chain64_kernel<N>), SM::bytes);
  int per_sm = (int)((228 * 1024) / (SM::bytes + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)num_sms() * per_sm;
  const long long ntiles = host_n_tiles(a, C_TILE);
  if (grid > ntiles) grid = ntiles;
  if (grid > MAX_PARTIAL_BLOCKS) grid = MAX_PARTIAL_BLOCKS;
  if (grid < 1) grid = 1;
  fvp_chain64_kernel<N><<<(unsigned)grid, C_THREADS, SM::bytes, st>>>(a);
  B200RL_LAUNCH_CHECK("fvp_chain64_kernel");
  *grid_out = (int)grid;
  return 0;
}

int update_fvp64_launch(int obs_dim, int act_dim, const UpdArgs& a, int* grid_out, int* P_out, int* ols_out,
                        cudaStream_t st) {
  const int h1 = 32, h2 = 32;
  B200RL_DISPATCH_NET_H(32, {
    *P_out = NetT::P;
    *ols_out = NetT::ols;
    int rc = launch_chain64<NetT>(a, grid_out, st);
    if (rc) return rc;
  });
  return 0;
}

}  // namespace b200rl
