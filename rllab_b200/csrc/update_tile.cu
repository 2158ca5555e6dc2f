// Surrogate gradient and Fisher-vector product for 32-wide policies: one THREAD per sample for the per-sample math
// (forward, tangent-forward, backward -- the same canonical summation order as the rollout kernel), then a
// block-cooperative accumulation of the weight gradients as small Gram products over a 128-sample tile staged in
// shared memory:
//     dW0 = X^T D1, db0 = 1^T D1, dW1 = H1^T D2, db1 = 1^T D2, dWout = H2^T DM, dbout = 1^T DM, dlog_std = 1^T DL
// dW1 (32x32 outputs, the bulk) uses 4x4 register tiles, rows interleaved by 8 so that every LDS.128 of a warp is
// conflict-free (one wavefront), split in two K-halves over the 128 threads; per-tile float32 partial products are
// folded into float64 register accumulators; per-block float64 partials are reduced in fixed order by the caller.
//
// Replaces f_grad / f_Hx_plain of rllab/optimizers/conjugate_gradient_optimizer.py:184-215,22-55 and the gradient
// half of f_opt in rllab/optimizers/first_order_optimizer.py:62-76.
#include "tile_phase_a.cuh"

namespace b200rl {

constexpr int T_THREADS = 128, T_TILE = 128, T_LD = T_TILE + 4;

template <class N, int MODE>
struct TileSmem {
  static constexpr int O = N::O, H = N::H1, A = N::A;
  static_assert(N::H1 == 32 && N::H2 == 32, "tile kernel is specialised for 32-wide layers");
  static constexpr int rX = 0, rH1 = rX + O, rH2 = rH1 + H, rD1 = rH2 + H, rD2 = rD1 + H, rDM = rD2 + H, rDL = rDM + A,
                       R = rDL + A;
  static constexpr int P4 = (N::P + 3) & ~3;
  static constexpr int o_sp = 0, o_sv = P4, o_stage = (MODE == MODE_FVP ? 2 : 1) * P4;
  static constexpr int n_floats = o_stage + R * T_LD;
  static constexpr int scratch_off = ((n_floats * 4 + 15) / 16) * 16;
  static constexpr size_t bytes = (size_t)scratch_off + 3 * 32 * 8;
  static_assert(2 * 64 * 16 * 8 <= R * T_LD * 4, "stage region must hold the K-half combine scratch");
};

// Resident CTAs per SM: 3 when three tiles fit the 228 KB of shared memory (small O, A: CartPole / Pendulum / PointEnv
// gradient -> 168 registers, 12 warps/SM: 3.88 -> 3.44 ms on cfg2, A/B measured), otherwise 2 (<= 255 registers).
template <class N, int MODE>
constexpr int tile_minblocks() { return (TileSmem<N, MODE>::bytes + 1024) * 3 <= 228 * 1024 ? 3 : 2; }

template <class N, int MODE>
__global__ void __launch_bounds__(T_THREADS, tile_minblocks<N, MODE>()) update_tile_kernel(UpdArgs a) {
  using SM = TileSmem<N, MODE>;
  constexpr int O = N::O, H = 32, A = N::A, P = N::P, LD = T_LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sf = reinterpret_cast<float*>(smem_raw);
  float* sp = sf + SM::o_sp;
  float* sv = sf + SM::o_sv;
  float* stage = sf + SM::o_stage;
  double* red_scratch = reinterpret_cast<double*>(smem_raw + SM::scratch_off);
  const int tid = threadIdx.x;

  for (int i = tid; i < P; i += T_THREADS) sp[i] = a.params[i];
  if constexpr (MODE == MODE_FVP)
    for (int i = tid; i < P; i += T_THREADS) sv[i] = (float)a.xvec[i];
  __syncthreads();

  // distribution constants
  TileDist D;
  D.sum_ls_new = 0.f; D.sum_ls_old = 0.f;
#pragma unroll
  for (int k = 0; k < A; ++k) {
    D.ls_new[k] = clamp_log_std(sp[N::ols + k], a.log_min_std);
    const float sd = expf(D.ls_new[k]);
    D.inv_std[k] = 1.0f / sd;
    D.var_new[k] = sd * sd;
    D.var_new2[k] = 2.0f * sd * sd + 1e-8f;
    D.Mmu[k] = 2.0f / D.var_new2[k];
    D.ls_old[k] = (MODE == MODE_FVP) ? D.ls_new[k] : a.old_log_std[k];
    const float so = expf(D.ls_old[k]);
    D.inv_std_old[k] = 1.0f / so;
    D.var_old[k] = so * so;
    D.sum_ls_new += D.ls_new[k];
    D.sum_ls_old += D.ls_old[k];
  }
  D.half_log2pi_A = 0.5f * (float)A * 1.8378770664093453f;

  // ---- Gram ownership
  const int w1_tile = tid & 63, kh = tid >> 6;
  const int ti = w1_tile >> 3, tj = w1_tile & 7;
  double accW1[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) accW1[r][c] = 0.0;
  constexpr int NS = (O + 1 > A + 1) ? O + 1 : A + 1;
  double accS[NS];   // small-output accumulators of this thread's task (see below)
#pragma unroll
  for (int k = 0; k < NS; ++k) accS[k] = 0.0;
  double s_loss = 0.0, s_kl = 0.0, m_kl = -1.0e300;

  const long long ntiles = n_tiles_of(a, T_TILE);
  for (long long ti_ = blockIdx.x; ti_ < ntiles; ti_ += gridDim.x) {
    asm volatile("" ::: "memory");
    const long long s = tile_at(a, ti_) * T_TILE + tid;
    const bool inrange = s < a.B;
    const bool valid = sample_valid(a, s);
    // ================= phase A: per-sample forward / (tangent) / backward, staged to shared memory
    tile_phase_a<N, MODE, SM, LD>(a, sp, sv, stage, D, inrange ? s : a.B - 1, inrange, valid, tid, s_loss, s_kl, m_kl);
    __syncthreads();
    // ================= phase B: Gram accumulation over the tile
    {
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
      const float* U = stage + (SM::rH1 + ti) * LD + kh * 64;
      const float* V = stage + (SM::rD2 + tj) * LD + kh * 64;
#pragma unroll 4
      for (int k = 0; k < 64; k += 4) {
        float4 u[4], v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) u[r] = *reinterpret_cast<const float4*>(U + r * 8 * LD + k);
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = *reinterpret_cast<const float4*>(V + c * 8 * LD + k);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[r][c] = fmaf(u[r].x, v[c].x, acc[r][c]);
            acc[r][c] = fmaf(u[r].y, v[c].y, acc[r][c]);
            acc[r][c] = fmaf(u[r].z, v[c].z, acc[r][c]);
            acc[r][c] = fmaf(u[r].w, v[c].w, acc[r][c]);
          }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) accW1[r][c] += (double)acc[r][c];
      // small outputs: warp 0 -> (dW0[:,j], db0[j]); warp 1 -> (dWout[j,:], db1[j]); warp 2 lanes < 2A -> dbout / dlog_std
      if (tid < 32) {
        float sa[O + 1];
#pragma unroll
        for (int o = 0; o <= O; ++o) sa[o] = 0.f;
        const float* D = stage + (SM::rD1 + tid) * LD;
#pragma unroll 4
        for (int k = 0; k < T_TILE; k += 4) {
          const float4 d = *reinterpret_cast<const float4*>(D + k);
#pragma unroll
          for (int o = 0; o < O; ++o) {
            const float4 xv = *reinterpret_cast<const float4*>(stage + (SM::rX + o) * LD + k);
            sa[o] = fmaf(xv.x, d.x, sa[o]); sa[o] = fmaf(xv.y, d.y, sa[o]);
            sa[o] = fmaf(xv.z, d.z, sa[o]); sa[o] = fmaf(xv.w, d.w, sa[o]);
          }
          sa[O] += (d.x + d.y) + (d.z + d.w);
        }
#pragma unroll
        for (int o = 0; o <= O; ++o) accS[o] += (double)sa[o];
      } else if (tid < 64) {
        const int j = tid - 32;
        float sa[A + 1];
#pragma unroll
        for (int k = 0; k <= A; ++k) sa[k] = 0.f;
        const float* Hh = stage + (SM::rH2 + j) * LD;
        const float* D = stage + (SM::rD2 + j) * LD;
#pragma unroll 4
        for (int k = 0; k < T_TILE; k += 4) {
          const float4 hv = *reinterpret_cast<const float4*>(Hh + k);
          const float4 d = *reinterpret_cast<const float4*>(D + k);
#pragma unroll
          for (int q = 0; q < A; ++q) {
            const float4 m = *reinterpret_cast<const float4*>(stage + (SM::rDM + q) * LD + k);
            sa[q] = fmaf(hv.x, m.x, sa[q]); sa[q] = fmaf(hv.y, m.y, sa[q]);
            sa[q] = fmaf(hv.z, m.z, sa[q]); sa[q] = fmaf(hv.w, m.w, sa[q]);
          }
          sa[A] += (d.x + d.y) + (d.z + d.w);
        }
#pragma unroll
        for (int k = 0; k <= A; ++k) accS[k] += (double)sa[k];
      } else if (tid < 64 + 2 * A) {
        const float* D = stage + (SM::rDM + (tid - 64)) * LD;   // rows DM[0..A-1], DL[0..A-1] are contiguous
        float s0 = 0.f;
#pragma unroll 4
        for (int k = 0; k < T_TILE; k += 4) {
          const float4 d = *reinterpret_cast<const float4*>(D + k);
          s0 += (d.x + d.y) + (d.z + d.w);
        }
        accS[0] += (double)s0;
      }
    }
    __syncthreads();
  }

  // ================= write this block's partial vector (float64)
  double* out = a.partial + (size_t)blockIdx.x * P;
  double* scr = reinterpret_cast<double*>(stage);   // [2][64][16]
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) scr[(kh * 64 + w1_tile) * 16 + r * 4 + c] = accW1[r][c];
  __syncthreads();
  if (tid < 64) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        out[N::oW1 + (ti + 8 * r) * H + (tj + 8 * c)] = scr[w1_tile * 16 + r * 4 + c] + scr[(64 + w1_tile) * 16 + r * 4 + c];
  }
  if (tid < 32) {
#pragma unroll
    for (int o = 0; o < O; ++o) out[N::oW0 + o * H + tid] = accS[o];
    out[N::ob0 + tid] = accS[O];
  } else if (tid < 64) {
    const int j = tid - 32;
#pragma unroll
    for (int k = 0; k < A; ++k) out[N::oWo + j * A + k] = accS[k];
    out[N::ob1 + j] = accS[A];
  } else if (tid < 64 + 2 * A) {
    out[N::obo + (tid - 64)] = accS[0];   // obo.. then ols.. are contiguous in the flat layout
  }
  if constexpr (MODE == MODE_GRAD) {
    __syncthreads();
    double v[2] = {s_loss, s_kl};
    double mx[1] = {m_kl};
    double* sc = a.partial + (size_t)gridDim.x * P + (size_t)blockIdx.x * 3;
    block_reduce_store<2, false>(v, red_scratch, sc);
    block_reduce_store<1, true>(mx, red_scratch, sc + 2);
  }
}

template <class N, int MODE>
static int launch_tile(const UpdArgs& a, int* grid_out, cudaStream_t st) {
  using SM = TileSmem<N, MODE>;
  B200RL_SET_MAX_SMEM((update_tile_kernel<N, MODE>), SM::bytes);
  int per_sm = (int)((228 * 1024) / (SM::bytes + 1024));   // 228 KB per SM, 1 KB reserved per resident CTA
  if (per_sm < 1) per_sm = 1;
  if (per_sm > tile_minblocks<N, MODE>()) per_sm = tile_minblocks<N, MODE>();
  long long grid = (long long)num_sms() * per_sm;
  const long long ntiles = host_n_tiles(a, T_TILE);
  if (grid > ntiles) grid = ntiles;
  if (grid > MAX_PARTIAL_BLOCKS) grid = MAX_PARTIAL_BLOCKS;
  if (grid < 1) grid = 1;
  update_tile_kernel<N, MODE><<<(unsigned)grid, T_THREADS, SM::bytes, st>>>(a);
  B200RL_LAUNCH_CHECK("update_tile_kernel");
  *grid_out = (int)grid;
  return 0;
}

int update_tile_launch(int mode, int obs_dim, int act_dim, const UpdArgs& a, int* grid_out, int* P_out, int* ols_out,
                       cudaStream_t st) {
  const int h1 = 32, h2 = 32;
  B200RL_DISPATCH_NET_H(32, {
    *P_out = NetT::P;
    *ols_out = NetT::ols;
    int rc = (mode == MODE_GRAD) ? launch_tile<NetT, MODE_GRAD>(a, grid_out, st)
                                 : launch_tile<NetT, MODE_FVP>(a, grid_out, st);
    if (rc) return rc;
  });
  return 0;
}

}  // namespace b200rl
