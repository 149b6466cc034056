// Flat parameter / gradient layout (ActorCritic.parameters() order).
// Mirrors drl_urban_planning_b200/params.py; tests/test_abi.py checks both against upb_param_slot().
#pragma once

namespace upb {

constexpr int D = 16;           // gcn_node_dim
constexpr int F = 23;           // node feature dim
constexpr int FS = 24;          // padded node feature stride
constexpr int NUMD = 52;        // numerical feature dim
constexpr int NH0 = 64;         // numeric encoder hidden
constexpr int HID = 32;         // policy / value head hidden
constexpr int SVD = 67;         // value-head input: 16 + 16 + 16 + 16 + 3

// ---- parameter offsets (floats)
constexpr int P_NUM_W0 = 0;         // [64][52]
constexpr int P_NUM_B0 = 3328;      // [64]
constexpr int P_NUM_W1 = 3392;      // [16][64]
constexpr int P_NUM_B1 = 4416;      // [16]
constexpr int P_ENC_W = 4432;       // [16][23]
constexpr int P_ENC_B = 4800;       // [16]
constexpr int P_GCN0_W = 4816;      // [16][32]
constexpr int P_GCN0_B = 5328;      // [16]
constexpr int P_GCN1_W = 5344;
constexpr int P_GCN1_B = 5856;
constexpr int P_MHA_IN_W = 5872;    // [48][16]
constexpr int P_MHA_IN_B = 6640;    // [48]
constexpr int P_MHA_OUT_W = 6688;   // [16][16]
constexpr int P_MHA_OUT_B = 6944;   // [16]
constexpr int P_ATT_Q_W = 6960;
constexpr int P_ATT_Q_B = 7216;
constexpr int P_ATT_K_W = 7232;
constexpr int P_ATT_K_B = 7488;
constexpr int P_ATT_V_W = 7504;
constexpr int P_ATT_V_B = 7760;
constexpr int P_LU_W0 = 7776;       // [32][64]
constexpr int P_LU_B0 = 9824;       // [32]
constexpr int P_LU_W1 = 9856;       // [1][32]
constexpr int P_RD_W0 = 9888;       // [32][16]
constexpr int P_RD_B0 = 10400;      // [32]
constexpr int P_RD_W1 = 10432;      // [1][32]
constexpr int P_VAL_W0 = 10464;     // [32][67]
constexpr int P_VAL_B0 = 12608;     // [32]
constexpr int P_VAL_W1 = 12640;     // [32][32]
constexpr int P_VAL_B1 = 13664;     // [32]
constexpr int P_VAL_W2 = 13696;     // [1][32]
constexpr int P_VAL_B2 = 13728;     // [1]
constexpr int NUM_PARAMS = 13729;

constexpr int ENCODER_END = P_LU_W0;   // [0, ENCODER_END) shared encoder
constexpr int POLICY_END = P_VAL_W0;   // [ENCODER_END, POLICY_END) policy heads; rest value head

// ---- per-CTA partial gradient row: real parameters, then "virtual" gradients of the composed attention
// projections (chained to the real tensors once per step in k_finish_grad), then loss statistics.
constexpr int G_QC = 13744;            // [16][16]  d/d(Win_q Wq)
constexpr int G_QBC = G_QC + 256;      // [16]      d/d(Win_q bq + bin_q)
constexpr int G_KC = G_QBC + 16;       // [16][16]  d/d(Win_k Wk)
constexpr int G_VC = G_KC + 256;       // [16][16]  d/d(Win_v Wv)
constexpr int G_VBC = G_VC + 256;      // [16]      d/d(Win_v bv + bin_v)
constexpr int G_STATS = G_VBC + 16;    // 14544: [8] statistics (see upb200.h)
constexpr int G_ROW = 14592;           // row stride (multiple of 64)

// beta^n for an integer step count by repeated squaring in double (a handful of multiplies; libdevice pow(double) costs
// thousands of cycles in a one-thread critical path)
#ifdef __CUDACC__
__host__ __device__
#endif
inline double ipow(double b, long long n) {
  double r = 1.0;
  while (n > 0) {
    if (n & 1) r *= b;
    b *= b;
    n >>= 1;
  }
  return r;
}

}  // namespace upb
