// Shared declarations of the policy-update kernels (update.cu: warp-per-sample, 64-wide nets; update_tile.cu:
// thread-per-sample + shared-memory Gram accumulation, 32-wide nets).
#pragma once
#include "mlp.cuh"

namespace b200rl {

constexpr int MODE_LOSS = 0, MODE_GRAD = 1, MODE_FVP = 2;

struct UpdArgs {
  const float* params;
  const double* xvec;  // FVP: tangent vector (float64, P)
  float* h_cache;      // [H1+H2][B] hidden activations: written by GRAD (if non-null), read by FVP (if non-null)
  float log_min_std;
  long long B;
  const float *obs, *act, *adv, *old_mean, *old_log_std;
  int loss_kind;
  int unit_half;  // 64-wide nets: which half of the layer-2 units accumulates dW1 in this pass (0/1)
  double* partial;
};

// 32-wide nets.  Launches the tile kernel on `a` (a.partial = workspace); returns the grid size used (blocks that
// wrote partial[block][P] (+ [grid][3] loss scalars after them in MODE_GRAD)), P and the log_std offset.
int update_tile_launch(int mode, int obs_dim, int act_dim, const UpdArgs& a, int* grid_out, int* P_out, int* ols_out,
                       cudaStream_t st);

// 32- and 64-wide nets: the tiled-GEMM formulation (update_gemm.cu); same contract as update_tile_launch.
int update_gemm_launch(int mode, int obs_dim, int h, int act_dim, const UpdArgs& a, int* grid_out, int* P_out,
                       int* ols_out, cudaStream_t st);

// which implementation b200rl_grad / b200rl_fvp use: env B200RL_UPDATE_IMPL = auto (default: tile for 32-wide, gemm for
// 64-wide nets) | gemm | tile | warp
int update_impl();

}  // namespace b200rl
