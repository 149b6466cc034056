// Measured experiment: the dense per-node block of the SGNN kernel on the 5th-generation tensor
// cores (tcgen05.mma kind::tf32, accumulators in TMEM) against the packed-FMA version the kernel ships.
//
//   EPQ phase:  EPQ[n][32] = exp2(c * (H[n][16] . W^T[16][32] + b)),  n ~ 310 nodes of one graph per CTA
//               (state_encoder.py:110-130: the per-edge Linear of the reference, moved to the nodes by the P/Q split)
//
//   variant F : the kernel's epq_phase -- 8 lanes per node pair, weights in registers, FFMA2
//   variant T : H rows are split into TF32 head / tail and written to shared memory in the UMMA canonical K-major
//               layout (8 x 16-byte core matrices, no swizzle); one thread issues 3 passes (hi.hi, lo.hi, hi.lo: "3xTF32",
//               fp32-grade products) x 2 k-steps of tcgen05.mma M128 N32 K8 per 128-row tile; tcgen05.commit -> mbarrier;
//               the epilogue warps read the accumulators with tcgen05.ld 32x32b.x32, apply bias / exp2 and write EPQ rows.
// Both variants run REPS times inside one launch on every SM (persistent CTA, 512 threads, data in shared memory) and report
// cycles per phase (clock64, max over CTAs) and the max relative difference of their outputs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o epq_tc epq_tc.cu && ./epq_tc [n] [reps]
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int NT = 512;
constexpr int NMAX = 512;                 // rows kept per CTA (4 tiles of 128)
constexpr int TILE_BYTES = 128 * 16 * 4;  // one 128 x 16 fp32 operand tile

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float exp2a(float a) {
  const float t = fminf(fmaxf(a * 2.8853900817779268f, -115.41560327111707f), 115.41560327111707f);
  return ex2_approx(t);
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tf32_of(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

// ---------------------------------------------------------------------------------------------------- variant F
__device__ void epq_ffma(int n, const float* H, const float* WT, const float* b, float* EPQ) {
  const int og = threadIdx.x & 7;
  const float4 bias = og < 4 ? *reinterpret_cast<const float4*>(b + og * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  float2 wlo[16], whi[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float4 w = *reinterpret_cast<const float4*>(WT + c * 32 + og * 4);
    wlo[c] = make_float2(w.x, w.y); whi[c] = make_float2(w.z, w.w);
  }
  const int npair = (n + 1) >> 1;
  for (int task = threadIdx.x; task < npair * 8; task += NT) {
    const int i0 = (task >> 3) * 2, i1 = min(i0 + 1, n - 1);
    float4 ha[4], hb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ha[j] = *reinterpret_cast<const float4*>(H + i0 * 16 + j * 4);
      hb[j] = *reinterpret_cast<const float4*>(H + i1 * 16 + j * 4);
    }
    float2 a0 = make_float2(bias.x, bias.y), a1 = make_float2(bias.z, bias.w), b0 = a0, b1 = a1;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 va = ha[c >> 2], vb = hb[c >> 2];
      const float xa = (c & 3) == 0 ? va.x : (c & 3) == 1 ? va.y : (c & 3) == 2 ? va.z : va.w;
      const float xb = (c & 3) == 0 ? vb.x : (c & 3) == 1 ? vb.y : (c & 3) == 2 ? vb.z : vb.w;
      const float2 xa2 = make_float2(xa, xa), xb2 = make_float2(xb, xb);
      a0 = __ffma2_rn(wlo[c], xa2, a0); a1 = __ffma2_rn(whi[c], xa2, a1);
      b0 = __ffma2_rn(wlo[c], xb2, b0); b1 = __ffma2_rn(whi[c], xb2, b1);
    }
    *reinterpret_cast<float4*>(EPQ + i0 * 32 + og * 4) = make_float4(exp2a(a0.x), exp2a(a0.y), exp2a(a1.x), exp2a(a1.y));
    if (i1 != i0) *reinterpret_cast<float4*>(EPQ + i1 * 32 + og * 4) = make_float4(exp2a(b0.x), exp2a(b0.y), exp2a(b1.x), exp2a(b1.y));
  }
}

// ---------------------------------------------------------------------------------------------------- variant T
// canonical K-major, no swizzle: core matrix (mi, ki) = 8 rows x 4 tf32, 128 contiguous bytes, at ((mi * 4 + ki) * 128)
__device__ __forceinline__ uint32_t canon_off(int row, int kchunk) { return (uint32_t)(((row >> 3) * 4 + kchunk) * 128 + (row & 7) * 16); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct TcState {
  uint32_t tmem;       // base TMEM address (column 0 of the allocation)
  unsigned parity;
  long long t_stage, t_mma, t_ld, t_epi;     // accumulated cycles of thread 0: staging | issue + tensor pipe + commit
};                                           // wake-up | TMEM -> registers | exp2 + stores + closing barrier

__device__ void epq_tc(int n, const float* H, const float* b, float* EPQ, uint8_t* opA, const uint8_t* opB, uint64_t* mbar,
                       TcState& st, int lbo_is_k) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntile = (n + 127) >> 7;
  const long long c0 = clock64();
  // 1. operand staging: H rows -> TF32 head / tail tiles (hi tiles first, then lo tiles)
  for (int item = tid; item < ntile * 128 * 4; item += NT) {
    // 8 consecutive threads write the 8 rows of one core matrix (128 contiguous bytes): conflict-free 16-byte stores
    const int kc = (item >> 3) & 3, row = ((item >> 5) << 3) | (item & 7), t = row >> 7, r = row & 127;
    float4 v = row < n ? *reinterpret_cast<const float4*>(H + row * 16 + kc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 hi, lo;
    hi.x = tf32_of(v.x); hi.y = tf32_of(v.y); hi.z = tf32_of(v.z); hi.w = tf32_of(v.w);
    lo.x = tf32_of(v.x - __uint_as_float(hi.x)); lo.y = tf32_of(v.y - __uint_as_float(hi.y));
    lo.z = tf32_of(v.z - __uint_as_float(hi.z)); lo.w = tf32_of(v.w - __uint_as_float(hi.w));
    *reinterpret_cast<uint4*>(opA + (size_t)t * TILE_BYTES + canon_off(r, kc)) = hi;
    *reinterpret_cast<uint4*>(opA + (size_t)(4 + t) * TILE_BYTES + canon_off(r, kc)) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
  __syncthreads();
  const long long c1 = clock64();
  // 2. one thread issues the MMAs of all tiles, then one commit
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t kstride = 128, mstride = 512;                   // bytes between core matrices along K / along M,N
    const uint32_t lbo = lbo_is_k ? kstride : mstride, sbo = lbo_is_k ? mstride : kstride;
    const uint32_t aB = smem_u32(opB);
    for (int t = 0; t < ntile; ++t) {
      const uint32_t d = st.tmem + (uint32_t)t * 32u;
      const uint32_t aHi = smem_u32(opA + (size_t)t * TILE_BYTES), aLo = smem_u32(opA + (size_t)(4 + t) * TILE_BYTES);
      const uint32_t bHi = aB, bLo = aB + 32 * 16 * 4;
      uint32_t acc = 0;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = pass == 1 ? aLo : aHi, bb = pass == 2 ? bLo : bHi;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          mma_tf32(d, make_desc(a + ks * 256, lbo, sbo), make_desc(bb + ks * 256, lbo, sbo), idesc, acc);
          acc = 1;
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
  }
  // 3. epilogue: warps 4t .. 4t+3 own tile t (a warp reads the TMEM lanes of its quarter)
  mbar_wait(mbar, st.parity);
  st.parity ^= 1u;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const long long c2 = clock64();
  long long c3 = c2;
  const int t = warp >> 2;
  if (t < ntile) {
    const int row = t * 128 + (warp & 3) * 32 + lane;
    uint32_t v[32];
    const uint32_t taddr = st.tmem + (uint32_t)t * 32u + ((uint32_t)((warp & 3) * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    c3 = clock64();
    if (row < n) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = (jj + lane) & 7;                             // rotated chunk order: fewer bank conflicts
        float4 o;
        o.x = exp2a(__uint_as_float(v[j * 4 + 0]) + (j < 4 ? b[j * 4 + 0] : 0.f));
        o.y = exp2a(__uint_as_float(v[j * 4 + 1]) + (j < 4 ? b[j * 4 + 1] : 0.f));
        o.z = exp2a(__uint_as_float(v[j * 4 + 2]) + (j < 4 ? b[j * 4 + 2] : 0.f));
        o.w = exp2a(__uint_as_float(v[j * 4 + 3]) + (j < 4 ? b[j * 4 + 3] : 0.f));
        *reinterpret_cast<float4*>(EPQ + row * 32 + j * 4) = o;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  const long long c4 = clock64();
  st.t_stage += c1 - c0; st.t_mma += c2 - c1; st.t_ld += c3 - c2; st.t_epi += c4 - c3;
}

// ---------------------------------------------------------------------------------------------------- driver kernel
// smem: H [NMAX][16] | WT [16][32] | b [16] | EPQ [NMAX][32] | opA 8 tiles | opB (hi, lo) [32][16] canonical
constexpr size_t SM_H = 0, SM_WT = SM_H + NMAX * 16 * 4, SM_B = SM_WT + 16 * 32 * 4, SM_EPQ = SM_B + 64,
                 SM_OPA = SM_EPQ + NMAX * 32 * 4, SM_OPB = SM_OPA + 8 * TILE_BYTES, SM_END = SM_OPB + 2 * 32 * 16 * 4;

__global__ void __launch_bounds__(NT, 1) probe(int n, int reps, int variant, int lbo_is_k, const float* gH, const float* gW /*[32][16] row o, col c*/,
                                               const float* gb, float* gOut, long long* gCycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  float* H = reinterpret_cast<float*>(sm + SM_H);
  float* WT = reinterpret_cast<float*>(sm + SM_WT);
  float* b = reinterpret_cast<float*>(sm + SM_B);
  float* EPQ = reinterpret_cast<float*>(sm + SM_EPQ);
  const int tid = threadIdx.x;
  for (int i = tid; i < n * 16; i += NT) H[i] = gH[i];
  for (int i = tid; i < 512; i += NT) { const int o = i >> 4, c = i & 15; WT[c * 32 + o] = gW[i]; }
  if (tid < 16) b[tid] = gb[tid];
  // B operand: W [32 o][16 c] is already "N rows x K columns" (K-major); TF32 head and tail in the canonical layout
  for (int i = tid; i < 32 * 4; i += NT) {
    const int o = i >> 2, kc = i & 3;
    uint4 hi, lo;
    const float4 w = *reinterpret_cast<const float4*>(gW + o * 16 + kc * 4);
    hi.x = tf32_of(w.x); hi.y = tf32_of(w.y); hi.z = tf32_of(w.z); hi.w = tf32_of(w.w);
    lo.x = tf32_of(w.x - __uint_as_float(hi.x)); lo.y = tf32_of(w.y - __uint_as_float(hi.y));
    lo.z = tf32_of(w.z - __uint_as_float(hi.z)); lo.w = tf32_of(w.w - __uint_as_float(hi.w));
    *reinterpret_cast<uint4*>(sm + SM_OPB + canon_off(o, kc)) = hi;
    *reinterpret_cast<uint4*>(sm + SM_OPB + 32 * 16 * 4 + canon_off(o, kc)) = lo;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  TcState st{0u, 0u, 0, 0, 0, 0};
  if (variant == 1) {
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    st.tmem = tmem_base;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (variant == 0) { epq_ffma(n, H, WT, b, EPQ); __syncthreads(); }
    else epq_tc(n, H, b, EPQ, sm + SM_OPA, sm + SM_OPB, &mbar, st, lbo_is_k);
  }
  const long long t1 = clock64();
  if (tid == 0) gCycles[blockIdx.x] = t1 - t0;
  if (tid == 0 && blockIdx.x == 0 && variant == 1) {
    gCycles[gridDim.x + 0] = st.t_stage; gCycles[gridDim.x + 1] = st.t_mma; gCycles[gridDim.x + 2] = st.t_ld; gCycles[gridDim.x + 3] = st.t_epi;
  }
  if (blockIdx.x == 0) for (int i = tid; i < n * 32; i += NT) gOut[i] = EPQ[i];
  if (variant == 1) {
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(st.tmem) : "memory");
  }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 310, reps = argc > 2 ? atoi(argv[2]) : 200;
  if (n < 1 || n > NMAX) { printf("n must be in [1, %d]\n", NMAX); return 1; }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int grid = prop.multiProcessorCount;
  std::vector<float> H(n * 16), W(512), b(16);
  srand(7);
  for (auto& x : H) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& x : W) x = (rand() / (float)RAND_MAX - 0.5f) * 0.5f;
  for (auto& x : b) x = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
  float *dH, *dW, *db, *dOut; long long* dCyc;
  CK(cudaMalloc(&dH, H.size() * 4)); CK(cudaMalloc(&dW, 2048)); CK(cudaMalloc(&db, 64));
  CK(cudaMalloc(&dOut, (size_t)n * 32 * 4)); CK(cudaMalloc(&dCyc, (grid + 4) * 8));
  CK(cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W.data(), 2048, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), 64, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_END));
  // float64 reference
  std::vector<double> ref((size_t)n * 32);
  for (int i = 0; i < n; ++i)
    for (int o = 0; o < 32; ++o) {
      double s = o < 16 ? b[o] : 0.0;
      for (int c = 0; c < 16; ++c) s += (double)W[o * 16 + c] * H[i * 16 + c];
      ref[(size_t)i * 32 + o] = std::exp(2.0 * s);
    }
  std::vector<float> out((size_t)n * 32);
  std::vector<long long> cyc(grid + 4);
  const char* names[3] = {"F  packed FMA (shipping epq_phase)", "T  tcgen05 3xTF32, LBO = K stride", "T' tcgen05 3xTF32, LBO = M/N stride"};
  for (int v = 0; v < 3; ++v) {
    const int variant = v == 0 ? 0 : 1, lbo_is_k = v == 1 ? 1 : 0;
    CK(cudaMemset(dOut, 0, (size_t)n * 32 * 4));
    probe<<<grid, NT, SM_END>>>(n, 3, variant, lbo_is_k, dH, dW, db, dOut, dCyc);      // warm-up
    CK(cudaDeviceSynchronize());
    probe<<<grid, NT, SM_END>>>(n, reps, variant, lbo_is_k, dH, dW, db, dOut, dCyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc.data(), dCyc, (grid + 4) * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; double mean = 0;
    for (int i = 0; i < grid; ++i) { const long long c = cyc[i]; mx = c > mx ? c : mx; mean += (double)c / grid; }
    double worst = 0;
    for (size_t i = 0; i < out.size(); ++i) worst = std::fmax(worst, std::fabs(out[i] - ref[i]) / std::fabs(ref[i]));
    printf("%-40s n=%d  cycles/phase: mean %.0f  max %.0f   max rel err vs f64 %.3g\n", names[v], n, mean / reps, (double)mx / reps, worst);
    if (v == 1)
      printf("   CTA 0, thread 0, per phase: operand staging + proxy fence + barrier %.0f | MMA issue + tensor pipe + commit wake-up %.0f | "
             "tcgen05.ld %.0f | exp2 epilogue + stores + barrier %.0f cycles\n", (double)cyc[grid] / reps, (double)cyc[grid + 1] / reps,
             (double)cyc[grid + 2] / reps, (double)cyc[grid + 3] / reps);
  }
  return 0;
}
