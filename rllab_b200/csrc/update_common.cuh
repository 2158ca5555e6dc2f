// Shared declarations of the policy-update kernels (update.cu: loss/KL pass + the C entry points; update_tile.cu:
// thread-per-sample + shared-memory Gram accumulation, 32-wide nets; update_gemm.cu: tiled-GEMM formulation, 64-wide nets).
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
  const unsigned char* flags;  // [B] or NULL: samples carrying B200RL_FLAG_MASKED contribute nothing
  const int* tile_list;        // FVP sub-sampling: indices of the 128-sample tiles to visit (device), or NULL = all tiles
  int n_list;
  double* partial;
};

// valid sample: inside the batch and not masked out by process_samples(drop_cut_paths)
__device__ __forceinline__ bool sample_valid(const UpdArgs& a, long long s) {
  return s < a.B && !(a.flags != nullptr && (a.flags[s] & B200RL_FLAG_MASKED));
}
// number of tiles a kernel iterates over and the i-th of them
__device__ __forceinline__ long long n_tiles_of(const UpdArgs& a, int tile) {
  return a.tile_list != nullptr ? (long long)a.n_list : (a.B + tile - 1) / tile;
}
__device__ __forceinline__ long long tile_at(const UpdArgs& a, long long i) {
  return a.tile_list != nullptr ? (long long)a.tile_list[i] : i;
}
inline long long host_n_tiles(const UpdArgs& a, int tile) {
  return a.tile_list != nullptr ? (long long)a.n_list : (a.B + tile - 1) / tile;
}

// 32-wide nets.  Launches the tile kernel on `a` (a.partial = workspace); returns the grid size used (blocks that
// wrote partial[block][P] (+ [grid][3] loss scalars after them in MODE_GRAD)), P and the log_std offset.
int update_tile_launch(int mode, int obs_dim, int act_dim, const UpdArgs& a, int* grid_out, int* P_out, int* ols_out,
                       cudaStream_t st);

// 32- and 64-wide nets: the tiled-GEMM formulation (update_gemm.cu); same contract as update_tile_launch.
int update_gemm_launch(int mode, int obs_dim, int h, int act_dim, const UpdArgs& a, int* grid_out, int* P_out,
                       int* ols_out, cudaStream_t st);


// 64-wide nets, gradient and (with cached activations) Fisher-vector product: dense layer chain on the tensor cores
// (update_umma.cu); same contract as update_tile_launch.
int update_umma64_launch(int mode, int obs_dim, int act_dim, const UpdArgs& a, int* grid_out, int* P_out, int* ols_out,
                         cudaStream_t st);


// 32-wide nets, gradient and (with cached activations) Fisher-vector product: dense layer chain on the tensor cores
// (update_umma32.cu); same contract as update_tile_launch.
int update_umma32_launch(int mode, int obs_dim, int act_dim, const UpdArgs& a, int* grid_out, int* P_out, int* ols_out,
                         cudaStream_t st);


}  // namespace b200rl
