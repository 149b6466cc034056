// Small kernels around the fused SGNN kernel: cross-CTA gradient reduction, attention chain rule,
// clip + Adam, GAE.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/upb200.h"
#include "layout.h"

namespace upb {

// Gradient tail, one launch: every block sums its 256 columns of gpart over the CTAs (fixed order ->
// deterministic) and writes real-parameter columns straight into the flat gradient buffer; the block that
// finishes last (ticket counter) chains the composed-attention ("virtual") gradients to the six real tensors:
//   q' = Win_q (Wq hc + bq) + bin_q = Qc hc + qbc   =>  g_Wq = Win_q^T g_Qc,  g_bq = Win_q^T g_qbc,
//   g_Win_q = g_Qc Wq^T + g_qbc bq^T,  g_bin_q = g_qbc;   same for V;  K has no bias gradient (softmax shift
//   invariance).
// grad = [13,729 gradients | 3 pad | 28 statistics] (upb200.h).
constexpr int RF_THREADS = 256;
constexpr int RF_BLOCKS = (G_ROW + RF_THREADS - 1) / RF_THREADS;

__global__ void __launch_bounds__(RF_THREADS) k_reduce_finish(const float* __restrict__ gpart, int nparts,
                                                              float* __restrict__ gsum, const float* __restrict__ P,
                                                              float* __restrict__ grad, unsigned int* ticket) {
  __shared__ float sG[816];        // Qc | qbc | Kc | Vc | vbc gradients
  __shared__ float sWin[768];      // in_proj_weight
  __shared__ float sW[768];        // Wq | Wk | Wv
  __shared__ float sB[48];         // bq | bk | bv
  __shared__ bool is_last;
  const int t = threadIdx.x;
  const int idx = blockIdx.x * RF_THREADS + t;
  if (idx < G_ROW) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = 0;
    for (; c + 4 <= nparts; c += 4) {
      s0 += gpart[(size_t)(c + 0) * G_ROW + idx];
      s1 += gpart[(size_t)(c + 1) * G_ROW + idx];
      s2 += gpart[(size_t)(c + 2) * G_ROW + idx];
      s3 += gpart[(size_t)(c + 3) * G_ROW + idx];
    }
    for (; c < nparts; ++c) s0 += gpart[(size_t)c * G_ROW + idx];
    const float v = (s0 + s1) + (s2 + s3);
    gsum[idx] = v;
    const bool attn = (idx >= P_MHA_IN_W && idx < P_MHA_OUT_W) || (idx >= P_ATT_Q_W && idx < P_LU_W0);
    if (idx < NUM_PARAMS && !attn) grad[idx] = v;                      // attention tensors are written by the chain
    else if (idx >= NUM_PARAMS && idx < UPB_STAT_OFFSET) grad[idx] = 0.f;
    if (idx >= G_STATS && idx < G_STATS + 8) grad[UPB_STAT_OFFSET + (idx - G_STATS)] = v;
    if (idx >= G_STATS + 8 && idx < G_STATS + UPB_STAT_COUNT) grad[UPB_STAT_OFFSET + (idx - G_STATS)] = 0.f;
  }
  __threadfence();
  __syncthreads();
  if (t == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (t == 0) *ticket = 0u;        // ready for the next launch
  for (int i = t; i < 816; i += RF_THREADS) sG[i] = __ldcg(gsum + G_QC + i);
  for (int i = t; i < 768; i += RF_THREADS) sWin[i] = P[P_MHA_IN_W + i];
  {
    const int r = t >> 4, c = t & 15;      // 256 threads = 16 x 16
    sW[t] = P[P_ATT_Q_W + t];
    sW[256 + t] = P[P_ATT_K_W + t];
    sW[512 + t] = P[P_ATT_V_W + t];
    if (t < 16) { sB[t] = P[P_ATT_Q_B + t]; sB[16 + t] = P[P_ATT_K_B + t]; sB[32 + t] = P[P_ATT_V_B + t]; }
    __syncthreads();
    const int gC[3] = {0, 272, 528};       // offsets inside sG: Qc, Kc, Vc
    const int gB[3] = {256, -1, 784};      // qbc, -, vbc
    const int pW[3] = {P_ATT_Q_W, P_ATT_K_W, P_ATT_V_W};
    const int pB[3] = {P_ATT_Q_B, P_ATT_K_B, P_ATT_V_B};
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const float* Win = sWin + s * 256;
      const float* gc = sG + gC[s];
      const float* W = sW + s * 256;
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        a = fmaf(Win[rr * 16 + r], gc[rr * 16 + c], a);      // g_W[m=r][c]   = sum_rr Win[rr][m] gC[rr][c]
        b = fmaf(gc[r * 16 + rr], W[c * 16 + rr], b);        // g_Win[r][m=c] = sum_cc gC[r][cc] W[m][cc]
      }
      if (gB[s] >= 0) b = fmaf(sG[gB[s] + r], sB[s * 16 + c], b);
      grad[pW[s] + t] = a;
      grad[P_MHA_IN_W + s * 256 + t] = b;
      if (t < 16) {
        float gb = 0.f, gbin = 0.f;
        if (gB[s] >= 0) {
          for (int rr = 0; rr < 16; ++rr) gb = fmaf(Win[rr * 16 + t], sG[gB[s] + rr], gb);
          gbin = sG[gB[s] + t];
        }
        grad[pB[s] + t] = gb;
        grad[P_MHA_IN_B + s * 16 + t] = gbin;
      }
    }
  }
}

struct ApplyArgs {
  float* params;
  const float* grad;          // [UPB_GRAD_STRIDE]
  float* m;
  float* v;
  const long long* steps_in;  // [4] global, encoder+value, land-use head, road head
  long long* steps_out;       // [4] written by block 0 (ping-pong with steps_in across calls)
  float lr, beta1, beta2, eps;
  int clip_now;               // 1: two-group clip on this step (decided on the host: mode + first-step latch)
  // flat layout of the model being updated (SGNN: layout.h; rl-mlp: mlp_kernel.cuh)
  int num_params, encoder_end, policy_end, lu_begin, rd_begin, stat_offset;
};

constexpr int AP_THREADS = 512;
constexpr int AP_PER_THREAD = 2;
constexpr int AP_BLOCKS = (NUM_PARAMS + AP_THREADS * AP_PER_THREAD - 1) / (AP_THREADS * AP_PER_THREAD);

__device__ __forceinline__ float block_sum_ap(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < AP_THREADS / 32; ++w) s += red[w];
  __syncthreads();
  return s;
}

// clip_policy_grad (agent_ppo.py:43-46: clip_grad_norm_(policy params, 1) then clip_grad_norm_(value params, 1);
// the shared encoder is in both groups) followed by torch.optim.Adam.step (urban_planning_agent.py:145-149,337).
// Several blocks: each one recomputes the (rarely needed) clip norms itself, so there is no inter-block
// dependency; step counters are read from steps_in and written to steps_out.
__global__ void __launch_bounds__(AP_THREADS) k_apply(const ApplyArgs a) {
  __shared__ float red[AP_THREADS / 32];
  __shared__ float sh[8];
  const int t = threadIdx.x;
  const float* st = a.grad + a.stat_offset;
  const bool live_lu = st[5] > 0.f, live_rd = st[6] > 0.f;
  const long long gstep = a.steps_in[0];
  const bool do_clip = a.clip_now != 0;
  float c_enc = 1.f, c_pol = 1.f, c_val = 1.f;
  if (do_clip) {
    float se = 0.f, sp = 0.f, sv = 0.f;
    for (int i = t; i < a.num_params; i += AP_THREADS) {
      const float g = a.grad[i];
      if (i < a.encoder_end) se += g * g;
      else if (i < a.policy_end) sp += g * g;
      else sv += g * g;
    }
    se = block_sum_ap(se, red);
    sp = block_sum_ap(sp, red);
    sv = block_sum_ap(sv, red);
    const float n1 = sqrtf(se + sp);
    const float k1 = fminf(1.f / (n1 + 1e-6f), 1.f);               // policy group
    const float n2 = sqrtf(k1 * k1 * se + sv);
    const float k2 = fminf(1.f / (n2 + 1e-6f), 1.f);               // value group, encoder already scaled
    c_enc = k1 * k2; c_pol = k1; c_val = k2;
  }
  if (t < 3) {
    // per-segment Adam step counts: a head whose stage is absent has grad None and is skipped entirely
    const bool live = t == 0 ? true : (t == 1 ? live_lu : live_rd);
    const long long stp = a.steps_in[1 + t] + (live ? 1 : 0);
    const double bc1 = 1.0 - ipow((double)a.beta1, stp > 0 ? stp : 1);
    const double bc2 = 1.0 - ipow((double)a.beta2, stp > 0 ? stp : 1);
    sh[t * 2 + 0] = (float)((double)a.lr / bc1);
    sh[t * 2 + 1] = (float)sqrt(bc2);
    if (blockIdx.x == 0) a.steps_out[1 + t] = stp;
  }
  if (blockIdx.x == 0 && t == 3) a.steps_out[0] = gstep + 1;
  __syncthreads();
  const float w1 = 1.f - a.beta1, w2 = 1.f - a.beta2;
#pragma unroll
  for (int j = 0; j < AP_PER_THREAD; ++j) {
    const int i = (blockIdx.x * AP_PER_THREAD + j) * AP_THREADS + t;
    if (i >= a.num_params) break;
    int seg = 0;
    bool live = true;
    const float coef = i < a.encoder_end ? c_enc : (i < a.policy_end ? c_pol : c_val);
    if (i >= a.lu_begin && i < a.rd_begin) { seg = 1; live = live_lu; }
    else if (i >= a.rd_begin && i < a.policy_end) { seg = 2; live = live_rd; }
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
