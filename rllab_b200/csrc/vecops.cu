// P-vector kernels (float64): conjugate gradient state updates, TRPO step size, line-search parameter update,
// Adam.  P <= a few thousand, so each op is ONE block with warp-shuffle + shared-memory reductions; everything
// stays on the device so that a whole CG solve needs no host synchronisation.
//
// Replaces: rllab/misc/krylov.py:7-39 (cg), rllab/optimizers/conjugate_gradient_optimizer.py:258-275
// (initial step size, backtracking candidates), lasagne.updates.adam as used by
// rllab/optimizers/first_order_optimizer.py:21-22,62-65, rllab/core/parameterized.py:60-70 (set_param_values cast).
#include "common.cuh"

namespace b200rl {

constexpr int VEC_THREADS = 512;

__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < nwarp; ++w) s += scratch[w];  // fixed order, every thread computes the same value
  return s;
}

__global__ void __launch_bounds__(VEC_THREADS)
    cg_init_kernel(long long P, const double* __restrict__ g, double* __restrict__ x, double* __restrict__ r,
                   double* __restrict__ p, double* __restrict__ st, int p_f32) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < P; i += blockDim.x) {
    const double gi = g[i];
    x[i] = 0.0; r[i] = gi; p[i] = p_f32 ? (double)(float)gi : gi;
    acc += gi * gi;
  }
  const double rdotr = block_sum(acc, scratch);
  if (threadIdx.x == 0) { st[0] = rdotr; st[1] = 0.0; st[2] = 0.0; st[3] = 0.0; }
}

// one krylov.cg iteration given z = A p  (krylov.py:25-35); a no-op once rdotr < tol was seen (the reference breaks)
__global__ void __launch_bounds__(VEC_THREADS)
    cg_step_kernel(long long P, const double* __restrict__ z, double* __restrict__ x, double* __restrict__ r,
                   double* __restrict__ p, double* __restrict__ st, double tol, int p_f32) {
  __shared__ double scratch[32];
  if (st[1] != 0.0) return;  // uniform across the block
  const double rdotr = st[0];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < P; i += blockDim.x) acc += p[i] * z[i];
  const double pz = block_sum(acc, scratch);
  const double v = rdotr / pz;
  acc = 0.0;
  for (long long i = threadIdx.x; i < P; i += blockDim.x) {
    x[i] += v * p[i];
    const double ri = r[i] - v * z[i];
    r[i] = ri;
    acc += ri * ri;
  }
  const double newrdotr = block_sum(acc, scratch);
  const double mu = newrdotr / rdotr;
  // p_f32: the search direction is kept exactly representable in float32 -- the precision in which the Fisher-vector
  // kernel reads it -- so that z = A p is the product with the very vector the recurrences use (an A p~ paired with an
  // unrounded p is an inconsistent matvec, which CG on this ill-conditioned system amplifies)
  for (long long i = threadIdx.x; i < P; i += blockDim.x) {
    const double pn = r[i] + mu * p[i];
    p[i] = p_f32 ? (double)(float)pn : pn;
  }
  if (threadIdx.x == 0) {
    st[0] = newrdotr;
    st[2] = pz;
    st[3] += 1.0;
    if (newrdotr < tol) st[1] = 1.0;
  }
}

__global__ void __launch_bounds__(VEC_THREADS)
    trpo_step_size_kernel(long long P, const double* __restrict__ x, const double* __restrict__ Hx, double delta,
                          double* __restrict__ step, double* __restrict__ info) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < P; i += blockDim.x) acc += x[i] * Hx[i];
  const double xHx = block_sum(acc, scratch);
  double beta = sqrt(2.0 * delta * (1.0 / (xHx + 1e-8)));  // conjugate_gradient_optimizer.py:260-263
  if (isnan(beta)) beta = 1.0;                              // :264-265
  for (long long i = threadIdx.x; i < P; i += blockDim.x) step[i] = beta * x[i];
  if (threadIdx.x == 0) { info[0] = beta; info[1] = xHx; }
}

__global__ void axpy_params_kernel(long long P, const double* __restrict__ prev, const double* __restrict__ step,
                                   double ratio, double* __restrict__ out, float* __restrict__ out32) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const double v = prev[i] - ratio * step[i];
  out[i] = v;
  out32[i] = (float)v;
}

__global__ void adam_kernel(long long P, double* __restrict__ th, float* __restrict__ th32, const double* __restrict__ g,
                            double* __restrict__ m, double* __restrict__ v, double a_t, double b1, double b2, double eps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const double gi = g[i];
  const double mi = b1 * m[i] + (1.0 - b1) * gi;
  const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const double t = th[i] - a_t * mi / (sqrt(vi) + eps);
  th[i] = t;
  th32[i] = (float)t;
}

// LinearFeatureBaseline.fit solve (linear_feature_baseline.py:26-37): (A^T A + reg I) w = A^T y from the packed upper
// triangle of the (d+1)x(d+1) Gram matrix [features | returns]; float64 Cholesky by one block; like the reference, the
// regularisation is multiplied by 10 (up to 5 attempts) while the solve fails (non-positive pivot / NaN).  The
// reference calls np.linalg.lstsq on the same regularised normal equations; for this SPD system the two agree to
// rounding.  d <= 44.
constexpr int LFB_DMAX = 44;
__global__ void __launch_bounds__(64) lfb_solve_kernel(const double* __restrict__ gram, int d, double reg0,
                                                       double* __restrict__ w_out, double* __restrict__ info) {
  __shared__ double A[LFB_DMAX][LFB_DMAX + 1];
  __shared__ double bvec[LFB_DMAX], wv[LFB_DMAX];
  __shared__ int ok;
  const int d1 = d + 1, tid = threadIdx.x;
  double reg = reg0;
  int attempt = 0;
  for (; attempt < 5; ++attempt) {
    // unpack: packed index of (i, j), i <= j, row-major upper triangle of a d1 x d1 matrix
    for (int e = tid; e < d * d; e += blockDim.x) {
      const int i = e / d, j = e % d;
      const int r = i < j ? i : j, c = i < j ? j : i;
      const int p = r * d1 - r * (r - 1) / 2 + (c - r);
      A[i][j] = gram[p] + (i == j ? reg : 0.0);
    }
    for (int i = tid; i < d; i += blockDim.x) bvec[i] = gram[i * d1 - i * (i - 1) / 2 + (d - i)];
    if (tid == 0) ok = 1;
    __syncthreads();
    for (int k = 0; k < d; ++k) {
      if (tid == 0) {
        const double piv = A[k][k];
        if (!(piv > 0.0)) ok = 0;
        A[k][k] = sqrt(piv);
      }
      __syncthreads();
      if (!ok) break;
      const double lkk = A[k][k];
      for (int i = k + 1 + tid; i < d; i += blockDim.x) A[i][k] /= lkk;
      __syncthreads();
      for (int i = k + 1 + tid; i < d; i += blockDim.x) {
        const double lik = A[i][k];
        for (int j = k + 1; j <= i; ++j) A[i][j] -= lik * A[j][k];
      }
      __syncthreads();
    }
    if (ok && tid == 0) {
      for (int i = 0; i < d; ++i) {            // L y = b
        double s = bvec[i];
        for (int k = 0; k < i; ++k) s -= A[i][k] * wv[k];
        wv[i] = s / A[i][i];
      }
      for (int i = d - 1; i >= 0; --i) {       // L^T w = y
        double s = wv[i];
        for (int k = i + 1; k < d; ++k) s -= A[k][i] * wv[k];
        wv[i] = s / A[i][i];
      }
      for (int i = 0; i < d; ++i)
        if (isnan(wv[i]) || isinf(wv[i])) ok = 0;
    }
    __syncthreads();
    if (ok) break;
    reg *= 10.0;
    __syncthreads();
  }
  // all 5 attempts failed: the triangular solve never ran (wv is not defined) -- write zeros (= "no baseline"), never
  // uninitialised memory; info[2] = 0 reports it
  for (int i = tid; i < d; i += blockDim.x) w_out[i] = ok ? wv[i] : 0.0;
  if (tid == 0) { info[0] = reg; info[1] = (double)attempt; info[2] = (double)ok; }
}

// out[i] = sum over ranks (i < n_sum) or max over ranks (i >= n_sum) of gathered[r][i], in rank order: the local half of
// the all-gather based "mixed all-reduce" (rllab_b200/parallel.py) -- one collective for a vector that carries sums and
// maxima, and a reduction order that is fixed by construction (bit-identical on every rank).
__global__ void reduce_ranks_kernel(const double* __restrict__ gathered, int world, long long n, long long n_sum,
                                    double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = gathered[i];
  for (int r = 1; r < world; ++r) {
    const double v = gathered[(size_t)r * n + i];
    acc = (i < n_sum) ? acc + v : fmax(acc, v);
  }
  out[i] = acc;
}

__global__ void f64_to_f32_kernel(long long n, const double* __restrict__ s, float* __restrict__ d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = (float)s[i];
}

}  // namespace b200rl

using namespace b200rl;

extern "C" {

int b200rl_cg_init(long long P, const double* g, double* x, double* r, double* p, double* cg_state, int p_f32,
                   void* stream) {
  B200RL_REQUIRE(P > 0 && g && x && r && p && cg_state, "cg_init: bad arguments");
  cg_init_kernel<<<1, VEC_THREADS, 0, (cudaStream_t)stream>>>(P, g, x, r, p, cg_state, p_f32);
  B200RL_LAUNCH_CHECK("cg_init_kernel");
  return 0;
}

int b200rl_cg_step(long long P, const double* z, double* x, double* r, double* p, double* cg_state,
                   double residual_tol, int p_f32, void* stream) {
  B200RL_REQUIRE(P > 0 && z && x && r && p && cg_state, "cg_step: bad arguments");
  cg_step_kernel<<<1, VEC_THREADS, 0, (cudaStream_t)stream>>>(P, z, x, r, p, cg_state, residual_tol, p_f32);
  B200RL_LAUNCH_CHECK("cg_step_kernel");
  return 0;
}

int b200rl_trpo_step_size(long long P, const double* x, const double* Hx, double max_constraint_val,
                          double* step_out, double* info_out, void* stream) {
  B200RL_REQUIRE(P > 0 && x && Hx && step_out && info_out, "trpo_step_size: bad arguments");
  trpo_step_size_kernel<<<1, VEC_THREADS, 0, (cudaStream_t)stream>>>(P, x, Hx, max_constraint_val, step_out,
                                                                      info_out);
  B200RL_LAUNCH_CHECK("trpo_step_size_kernel");
  return 0;
}

int b200rl_axpy_params(long long P, const double* theta_prev, const double* step, double ratio, double* theta_out,
                       float* theta_f32_out, void* stream) {
  B200RL_REQUIRE(P > 0 && theta_prev && step && theta_out && theta_f32_out, "axpy_params: bad arguments");
  axpy_params_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, theta_prev, step, ratio,
                                                                                     theta_out, theta_f32_out);
  B200RL_LAUNCH_CHECK("axpy_params_kernel");
  return 0;
}

int b200rl_adam_step(long long P, double* theta, float* theta_f32, const double* g, double* m, double* v, long long t,
                     double lr, double b1, double b2, double eps, void* stream) {
  B200RL_REQUIRE(P > 0 && theta && theta_f32 && g && m && v && t >= 1, "adam_step: bad arguments");
  const double a_t = lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
  adam_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, theta, theta_f32, g, m, v, a_t, b1,
                                                                              b2, eps);
  B200RL_LAUNCH_CHECK("adam_kernel");
  return 0;
}

int b200rl_lfb_solve(int obs_dim, const double* gram, double reg_coeff, double* w_out, double* info_out, void* stream) {
  B200RL_REQUIRE(gram && w_out && info_out && obs_dim > 0 && 2 * obs_dim + 4 <= LFB_DMAX, "lfb_solve: bad arguments");
  lfb_solve_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(gram, 2 * obs_dim + 4, reg_coeff, w_out, info_out);
  B200RL_LAUNCH_CHECK("lfb_solve_kernel");
  return 0;
}

int b200rl_reduce_ranks(const double* gathered, int world, long long n, long long n_sum, double* out, void* stream) {
  B200RL_REQUIRE(gathered && out && world >= 1 && n > 0 && n_sum >= 0 && n_sum <= n, "reduce_ranks: bad arguments");
  reduce_ranks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gathered, world, n, n_sum, out);
  B200RL_LAUNCH_CHECK("reduce_ranks_kernel");
  return 0;
}

int b200rl_f64_to_f32(long long n, const double* src, float* dst, void* stream) {
  B200RL_REQUIRE(n > 0 && src && dst, "f64_to_f32: bad arguments");
  f64_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, src, dst);
  B200RL_LAUNCH_CHECK("f64_to_f32_kernel");
  return 0;
}
}
