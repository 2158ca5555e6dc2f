// Phase A of update_tile_kernel: the per-sample math of one thread (forward, optional tangent-forward, backward).
// Register discipline: every H-vector is staged to this thread's own column of the shared-memory tile as soon as it
// exists and re-read from there (conflict-free) where it is needed again, so that at most one input vector (32) and
// one dense-layer accumulator set (64) are live at any time (~110 registers + constants).  ptxas' allocation for the
// fully unrolled fused body is otherwise erratic (it picked 72..255 registers with up to 9 KB of spills).
#pragma once
#include "tile_gram.cuh"

namespace b200rl {

#define B200RL_SECTION_BARRIER() asm volatile("" ::: "memory")  // keep smem weights from staying live in registers

template <class N, int MODE, class SM, int LD>
__device__ __forceinline__ void tile_phase_a(const UpdArgs& a, const float* sp, const float* sv, float* stage,
                                             const TileDist& D, long long sl, bool inrange, bool valid, int tid,
                                             double& s_loss, double& s_kl, double& m_kl) {
  constexpr int O = N::O, H = 32, A = N::A;
  float* colX = stage + SM::rX * LD + tid;
  float* colH1 = stage + SM::rH1 * LD + tid;
  float* colH2 = stage + SM::rH2 * LD + tid;
  float* colD1 = stage + SM::rD1 * LD + tid;   // FVP: holds h1 V1 temporarily before d1 overwrites it
  float* colD2 = stage + SM::rD2 * LD + tid;
  float* colDM = stage + SM::rDM * LD + tid;
  float* colDL = stage + SM::rDL * LD + tid;
  float dmu[A];
  {
    float x[O], h1[H];
#pragma unroll
    for (int o = 0; o < O; ++o) {
      x[o] = a.obs[(size_t)o * a.B + sl];
      colX[o * LD] = x[o];
    }
    // activation cache (theta is fixed between the gradient and the CG solve): GRAD writes tanh outputs, FVP reads them
    const bool cached = (MODE == MODE_FVP) && (a.h_cache != nullptr);
    float* hc = a.h_cache ? a.h_cache + sl : nullptr;
    if (cached) {
#pragma unroll
      for (int j = 0; j < H; ++j) h1[j] = hc[(size_t)j * a.B];
    } else {
      dense_thread<O, H>(sp + N::oW0, sp + N::ob0, x, h1);
#pragma unroll
      for (int j = 0; j < H; ++j) h1[j] = tanh_f(h1[j]);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) colH1[j * LD] = h1[j];
    if constexpr (MODE == MODE_FVP) {
      // h1 V1 -> parked in the D1 rows
      float p2b[H];
      dense_thread<H, H, false>(sv + N::oW1, nullptr, h1, p2b);
#pragma unroll
      for (int j = 0; j < H; ++j) colD1[j * LD] = p2b[j];
      B200RL_SECTION_BARRIER();
    }
    float h2[H];
    if (cached) {
#pragma unroll
      for (int j = 0; j < H; ++j) h2[j] = hc[(size_t)(H + j) * a.B];
    } else {
      dense_thread<H, H>(sp + N::oW1, sp + N::ob1, h1, h2);
#pragma unroll
      for (int j = 0; j < H; ++j) h2[j] = tanh_f(h2[j]);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) colH2[j * LD] = h2[j];
    if (MODE == MODE_GRAD && hc != nullptr && inrange) {   // masked samples too: the FVP pass reads their rows back
#pragma unroll
      for (int j = 0; j < H; ++j) {
        hc[(size_t)j * a.B] = colH1[j * LD];
        hc[(size_t)(H + j) * a.B] = h2[j];
      }
    }
    B200RL_SECTION_BARRIER();
    if constexpr (MODE == MODE_GRAD) {
      float mu[A];
#pragma unroll
      for (int k = 0; k < A; ++k) {
        float s0 = sp[N::obo + k], s1 = 0.f;
#pragma unroll
        for (int j = 0; j < H; j += 2) {
          s0 = fmaf(h2[j], sp[N::oWo + j * A + k], s0);
          s1 = fmaf(h2[j + 1], sp[N::oWo + (j + 1) * A + k], s1);
        }
        mu[k] = s0 + s1;
      }
      float z[A], zsq = 0.f, zsq_old = 0.f, kl = 0.f;
#pragma unroll
      for (int k = 0; k < A; ++k) {
        const float act = a.act[(size_t)k * a.B + sl];
        const float om = a.old_mean[(size_t)k * a.B + sl];
        z[k] = (act - mu[k]) * D.inv_std[k];
        zsq += z[k] * z[k];
        const float zo = (act - om) * D.inv_std_old[k];
        zsq_old += zo * zo;
        const float dm = om - mu[k];
        kl += (dm * dm + D.var_old[k] - D.var_new[k]) / D.var_new2[k] + D.ls_new[k] - D.ls_old[k];
      }
      const float adv_s = a.adv[sl];
      const float logp_new = -D.sum_ls_new - 0.5f * zsq - D.half_log2pi_A;
      float w_s, term;
      if (a.loss_kind == B200RL_LOSS_TRPO) {
        const float logp_old = -D.sum_ls_old - 0.5f * zsq_old - D.half_log2pi_A;
        w_s = expf(logp_new - logp_old) * adv_s;
        term = -w_s;
      } else {
        w_s = adv_s;
        term = -logp_new * adv_s;
      }
      if (!valid) { w_s = 0.f; term = 0.f; }
      s_loss += (double)term;
      if (valid) { s_kl += (double)kl; m_kl = fmax(m_kl, (double)kl); }
#pragma unroll
      for (int k = 0; k < A; ++k) {
        dmu[k] = -w_s * z[k] * D.inv_std[k];
        colDM[k * LD] = dmu[k];
        colDL[k * LD] = -w_s * (z[k] * z[k] - 1.0f);
      }
    } else {
      // tangent forward J x (x = sv): t1 = (1-h1^2)(x V0 + vb0); t2 = (1-h2^2)(t1 W1 + h1 V1 + vb1)
      float t1[H];
      dense_thread<O, H>(sv + N::oW0, sv + N::ob0, x, t1);
#pragma unroll
      for (int j = 0; j < H; ++j) t1[j] *= (1.0f - h1[j] * h1[j]);
      B200RL_SECTION_BARRIER();
      float t2[H];
      dense_thread<H, H>(sp + N::oW1, sv + N::ob1, t1, t2);
      float md[A];
#pragma unroll
      for (int k = 0; k < A; ++k) md[k] = sv[N::obo + k];
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const float h2j = colH2[j * LD];
        const float t2j = (t2[j] + colD1[j * LD]) * (1.0f - h2j * h2j);
#pragma unroll
        for (int k = 0; k < A; ++k) md[k] = fmaf(t2j, sp[N::oWo + j * A + k], fmaf(h2j, sv[N::oWo + j * A + k], md[k]));
      }
#pragma unroll
      for (int k = 0; k < A; ++k) {
        dmu[k] = valid ? md[k] * D.Mmu[k] : 0.f;
        colDM[k * LD] = dmu[k];
        colDL[k * LD] = 0.f;
      }
    }
  }
  B200RL_SECTION_BARRIER();
  // backward: d2 = (dmu Wout^T) (1-h2^2); d1 = (d2 W1^T) (1-h1^2)
  float d2[H];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    float sacc = 0.f;
#pragma unroll
    for (int k = 0; k < A; ++k) sacc = fmaf(dmu[k], sp[N::oWo + j * A + k], sacc);
    const float h2j = colH2[j * LD];
    d2[j] = sacc * (1.0f - h2j * h2j);
    colD2[j * LD] = d2[j];
  }
#pragma unroll
  for (int i = 0; i < H; ++i) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < H; j += 4) {
      const float4 w = *reinterpret_cast<const float4*>(sp + N::oW1 + i * H + j);
      acc = ffma2(make_float2(d2[j], d2[j + 1]), make_float2(w.x, w.y), acc);
      acc = ffma2(make_float2(d2[j + 2], d2[j + 3]), make_float2(w.z, w.w), acc);
    }
    const float h1i = colH1[i * LD];
    colD1[i * LD] = (acc.x + acc.y) * (1.0f - h1i * h1i);
  }
}

}  // namespace b200rl
