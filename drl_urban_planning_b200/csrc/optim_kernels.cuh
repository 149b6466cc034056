// Small kernels around the fused SGNN kernel: cross-CTA gradient reduction, attention chain rule,
// clip + Adam, GAE.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/upb200.h"
#include "layout.h"

namespace upb {

// gsum[idx] = sum over CTAs of gpart[cta][idx]   (fixed order -> deterministic)
__global__ void __launch_bounds__(256) k_reduce_partials(const float* __restrict__ gpart, int nparts,
                                                         float* __restrict__ gsum) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G_ROW) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = 0;
  for (; c + 4 <= nparts; c += 4) {
    s0 += gpart[(size_t)(c + 0) * G_ROW + idx];
    s1 += gpart[(size_t)(c + 1) * G_ROW + idx];
    s2 += gpart[(size_t)(c + 2) * G_ROW + idx];
    s3 += gpart[(size_t)(c + 3) * G_ROW + idx];
  }
  for (; c < nparts; ++c) s0 += gpart[(size_t)c * G_ROW + idx];
  gsum[idx] = (s0 + s1) + (s2 + s3);
}

// Flat gradient buffer = real gradients (+ attention tensors chained from the composed-projection gradients)
// followed by the loss statistics.  One CTA of 256 threads.
//   q' = Win_q (Wq hc + bq) + bin_q = Qc hc + qbc   =>  g_Wq = Win_q^T g_Qc,  g_bq = Win_q^T g_qbc,
//   g_Win_q = g_Qc Wq^T + g_qbc bq^T,  g_bin_q = g_qbc;   same for V;  K has no bias gradient (softmax shift
//   invariance).
__global__ void __launch_bounds__(256) k_finish_grad(const float* __restrict__ gsum, const float* __restrict__ P,
                                                     float* __restrict__ grad) {
  const int t = threadIdx.x;
  for (int i = t; i < UPB_GRAD_STRIDE; i += 256) {
    float v = 0.f;
    if (i < NUM_PARAMS) v = gsum[i];
    else if (i >= UPB_STAT_OFFSET && i < UPB_STAT_OFFSET + 8) v = gsum[G_STATS + (i - UPB_STAT_OFFSET)];
    grad[i] = v;
  }
  __syncthreads();
  const int r = t >> 4, c = t & 15;   // 256 threads = 16 x 16
  const int gC[3] = {G_QC, G_KC, G_VC};
  const int gB[3] = {G_QBC, -1, G_VBC};
  const int pW[3] = {P_ATT_Q_W, P_ATT_K_W, P_ATT_V_W};
  const int pB[3] = {P_ATT_Q_B, P_ATT_K_B, P_ATT_V_B};
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const float* Win = P + P_MHA_IN_W + s * 256;     // rows 16s..16s+15 of in_proj_weight
    const float* gc = gsum + gC[s];
    // g_W[m=r][c] = sum_rr Win[rr][m] gC[rr][c]
    float a = 0.f, b = 0.f;
    for (int rr = 0; rr < 16; ++rr) {
      a = fmaf(Win[rr * 16 + r], gc[rr * 16 + c], a);
      // g_Win[r][m=c] = sum_cc gC[r][cc] W[m][cc]
      b = fmaf(gc[r * 16 + rr], P[pW[s] + c * 16 + rr], b);
    }
    if (gB[s] >= 0) b = fmaf(gsum[gB[s] + r], P[pB[s] + c], b);
    grad[pW[s] + r * 16 + c] = a;
    grad[P_MHA_IN_W + s * 256 + r * 16 + c] = b;
    if (t < 16) {
      float gb = 0.f, gbin = 0.f;
      if (gB[s] >= 0) {
        for (int rr = 0; rr < 16; ++rr) gb = fmaf(Win[rr * 16 + t], gsum[gB[s] + rr], gb);
        gbin = gsum[gB[s] + t];
      }
      grad[pB[s] + t] = gb;
      grad[P_MHA_IN_B + s * 16 + t] = gbin;
    }
  }
}

struct ApplyArgs {
  float* params;
  const float* grad;       // [UPB_GRAD_STRIDE]
  float* m;
  float* v;
  long long* steps;        // [4] global, encoder+value, land-use head, road head
  float lr, beta1, beta2, eps;
  int clip_mode;
};

__device__ __forceinline__ float block_sum_1024(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < 32; ++w) s += red[w];
  __syncthreads();
  return s;
}

// clip_policy_grad (agent_ppo.py:43-46: clip_grad_norm_(policy params, 1) then clip_grad_norm_(value params, 1);
// the shared encoder is in both groups) followed by torch.optim.Adam.step (urban_planning_agent.py:145-149,337).
__global__ void __launch_bounds__(1024) k_apply(const ApplyArgs a) {
  __shared__ float red[32];
  __shared__ float sh[8];
  const int t = threadIdx.x;
  const float* st = a.grad + UPB_STAT_OFFSET;
  const bool live_lu = st[5] > 0.f, live_rd = st[6] > 0.f;
  const long long gstep = a.steps[0];
  const bool do_clip = a.clip_mode == UPB_CLIP_ALWAYS || (a.clip_mode == UPB_CLIP_REFERENCE && gstep == 0);
  float c_enc = 1.f, c_pol = 1.f, c_val = 1.f;
  if (do_clip) {
    float se = 0.f, sp = 0.f, sv = 0.f;
    for (int i = t; i < NUM_PARAMS; i += 1024) {
      const float g = a.grad[i];
      if (i < ENCODER_END) se += g * g;
      else if (i < POLICY_END) sp += g * g;
      else sv += g * g;
    }
    se = block_sum_1024(se, red);
    sp = block_sum_1024(sp, red);
    sv = block_sum_1024(sv, red);
    const float n1 = sqrtf(se + sp);
    const float k1 = fminf(1.f / (n1 + 1e-6f), 1.f);               // policy group
    const float n2 = sqrtf(k1 * k1 * se + sv);
    const float k2 = fminf(1.f / (n2 + 1e-6f), 1.f);               // value group, encoder already scaled
    c_enc = k1 * k2; c_pol = k1; c_val = k2;
  }
  if (t < 3) {
    // per-segment Adam step counts: a head whose stage is absent has grad None and is skipped entirely
    const bool live = t == 0 ? true : (t == 1 ? live_lu : live_rd);
    const long long stp = a.steps[1 + t] + (live ? 1 : 0);
    const double bc1 = 1.0 - pow((double)a.beta1, (double)(stp > 0 ? stp : 1));
    const double bc2 = 1.0 - pow((double)a.beta2, (double)(stp > 0 ? stp : 1));
    sh[t * 2 + 0] = (float)((double)a.lr / bc1);
    sh[t * 2 + 1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float w1 = 1.f - a.beta1, w2 = 1.f - a.beta2;
  for (int i = t; i < NUM_PARAMS; i += 1024) {
    int seg = 0;
    bool live = true;
    float coef = i < ENCODER_END ? c_enc : (i < POLICY_END ? c_pol : c_val);
    if (i >= P_LU_W0 && i < P_RD_W0) { seg = 1; live = live_lu; }
    else if (i >= P_RD_W0 && i < POLICY_END) { seg = 2; live = live_rd; }
    if (!live) continue;
    const float g = __fmul_rn(a.grad[i], coef);
    float m = a.m[i], v = a.v[i];
    m = __fadd_rn(m, __fmul_rn(w1, __fsub_rn(g, m)));                           // lerp_(grad, 1-beta1)
    v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(w2, g), g));       // mul_(beta2).addcmul_(g, g, 1-beta2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), sh[seg * 2 + 1]), a.eps);
    a.params[i] = __fadd_rn(a.params[i], __fmul_rn(-sh[seg * 2 + 0], __fdiv_rn(m, denom)));
    a.m[i] = m;
    a.v[i] = v;
  }
  __syncthreads();
  if (t == 0) {
    a.steps[0] = gstep + 1;
    a.steps[1] += 1;
    if (live_lu) a.steps[2] += 1;
    if (live_rd) a.steps[3] += 1;
  }
}

// estimate_advantages (khrylib/rl/core/common.py:5-26).  The recurrence only chains inside an episode
// (masks[i] == 0 at its last step), so one thread walks one episode backwards with the reference's exact fp32
// operation order; episodes run in parallel.
__global__ void __launch_bounds__(256) k_gae(const float* __restrict__ rewards, const float* __restrict__ masks,
                                             const float* __restrict__ values, int T, float gamma, float gamma_tau,
                                             float* __restrict__ adv, float* __restrict__ ret) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  if (!(i == T - 1 || masks[i] == 0.f)) return;       // not the last step of a segment
  float prev_v = 0.f, prev_a = 0.f;
  if (masks[i] != 0.f) { prev_v = 0.f; prev_a = 0.f; } // i == T-1: the scan starts from zeros
  for (int j = i; j >= 0; --j) {
    const float mk = masks[j];
    if (j != i && mk == 0.f) break;                    // previous segment
    const float vj = values[j];
    float d = __fmul_rn(__fmul_rn(gamma, prev_v), mk);
    d = __fsub_rn(__fadd_rn(rewards[j], d), vj);
    const float aj = __fadd_rn(d, __fmul_rn(__fmul_rn(gamma_tau, prev_a), mk));
    adv[j] = aj;
    ret[j] = __fadd_rn(vj, aj);
    prev_v = vj;
    prev_a = aj;
  }
}

}  // namespace upb
