// Phase B of the 32-wide update kernels (update_tile.cu: FP32 layer chain; update_umma32.cu: tcgen05 layer chain): the
// weight gradients as Gram products over one 128-sample tile staged feature-major in shared memory,
//     dW0 = X^T D1, db0 = 1^T D1, dW1 = H1^T D2, db1 = 1^T D2, dWout = H2^T DM, dbout = 1^T DM, dlog_std = 1^T DL,
// accumulated by the 128 threads of the CTA: dW1 (32 x 32 outputs, the bulk) in 4x4 register tiles, rows interleaved by 8
// so that every LDS.128 of a warp is conflict-free, split in two K-halves over the threads; the small outputs by warps
// 0 / 1 / 2.  Per-tile float32 partial products are folded into float64 register accumulators that live across the
// persistent tile loop; write() combines the K-halves through shared memory and stores the block's float64 partial vector.
#pragma once
#include "update_common.cuh"

namespace b200rl {

// distribution constants of one pass (A <= 3); FVP: old == new
struct TileDist {
  float ls_new[3], inv_std[3], ls_old[3], inv_std_old[3], Mmu[3], var_new[3], var_new2[3], var_old[3];
  float sum_ls_new, sum_ls_old, half_log2pi_A;
};

template <class N, int MODE>
__device__ __forceinline__ void tile_dist_init(TileDist& D, const float* log_std_params, const UpdArgs& a) {
  constexpr int A = N::A;
  D.sum_ls_new = 0.f;
  D.sum_ls_old = 0.f;
#pragma unroll
  for (int k = 0; k < A; ++k) {
    D.ls_new[k] = clamp_log_std(log_std_params[k], a.log_min_std);
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
}

// RX..RDL: first stage row of X, H1, H2, D1, D2, DM (DL rows follow DM); LD: row pitch in floats (tile + 4)
template <class N, int RX, int RH1, int RH2, int RD1, int RD2, int RDM, int LD>
struct TileGram {
  static constexpr int O = N::O, H = 32, A = N::A, TILE = 128;
  static constexpr int NS = (O + 1 > A + 1) ? O + 1 : A + 1;
  static_assert(N::H1 == 32 && N::H2 == 32, "32-wide layers");
  double accW1[4][4];
  double accS[NS];   // small-output accumulators of this thread's task

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) accW1[r][c] = 0.0;
#pragma unroll
    for (int k = 0; k < NS; ++k) accS[k] = 0.0;
  }

  __device__ __forceinline__ void accumulate(const float* stage, int tid) {
    const int w1_tile = tid & 63, kh = tid >> 6;
    const int ti = w1_tile >> 3, tj = w1_tile & 7;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    const float* U = stage + (RH1 + ti) * LD + kh * 64;
    const float* V = stage + (RD2 + tj) * LD + kh * 64;
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
      const float* D = stage + (RD1 + tid) * LD;
#pragma unroll 4
      for (int k = 0; k < TILE; k += 4) {
        const float4 d = *reinterpret_cast<const float4*>(D + k);
#pragma unroll
        for (int o = 0; o < O; ++o) {
          const float4 xv = *reinterpret_cast<const float4*>(stage + (RX + o) * LD + k);
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
      const float* Hh = stage + (RH2 + j) * LD;
      const float* D = stage + (RD2 + j) * LD;
#pragma unroll 4
      for (int k = 0; k < TILE; k += 4) {
        const float4 hv = *reinterpret_cast<const float4*>(Hh + k);
        const float4 d = *reinterpret_cast<const float4*>(D + k);
#pragma unroll
        for (int q = 0; q < A; ++q) {
          const float4 m = *reinterpret_cast<const float4*>(stage + (RDM + q) * LD + k);
          sa[q] = fmaf(hv.x, m.x, sa[q]); sa[q] = fmaf(hv.y, m.y, sa[q]);
          sa[q] = fmaf(hv.z, m.z, sa[q]); sa[q] = fmaf(hv.w, m.w, sa[q]);
        }
        sa[A] += (d.x + d.y) + (d.z + d.w);
      }
#pragma unroll
      for (int k = 0; k <= A; ++k) accS[k] += (double)sa[k];
    } else if (tid < 64 + 2 * A) {
      const float* D = stage + (RDM + (tid - 64)) * LD;   // rows DM[0..A-1], DL[0..A-1] are contiguous
      float s0 = 0.f;
#pragma unroll 4
      for (int k = 0; k < TILE; k += 4) {
        const float4 d = *reinterpret_cast<const float4*>(D + k);
        s0 += (d.x + d.y) + (d.z + d.w);
      }
      accS[0] += (double)s0;
    }
  }

  // out: this block's partial vector [P]; scr: >= 2 * 64 * 16 doubles of shared memory no thread still reads
  __device__ __forceinline__ void write(double* out, double* scr, int tid) {
    const int w1_tile = tid & 63, kh = tid >> 6;
    const int ti = w1_tile >> 3, tj = w1_tile & 7;
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
  }
};

}  // namespace b200rl
