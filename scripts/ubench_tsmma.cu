// Feasibility probe for a tensor-core denoise loop: one CTA, the shapes of one
// denoiser GEMM stage with the WEIGHTS as the TMEM-resident A operand (M = 128 output features) and the
// 10 activation rows (padded to N = 16) as the shared-memory B operand, 3xTF32.
//   (1) cost of turning a 128 x 128 fp32 weight tile that sits in shared memory (as the TMA stream of the
//       current kernel delivers it) into TF32 hi/lo planes in TMEM: LDS -> split -> tcgen05.st
//   (2) latency of one dependent stage: 48 tcgen05.mma (TS, M128 N16 K8) -> commit -> mbarrier wait ->
//       tcgen05.ld of the 16 accumulator columns -> write the next B operand (hi/lo, swizzled) ->
//       fence.proxy.async + barrier, chained 64 times
//   (3) numerics of the stage against fp64 (checks the TMEM A layout and the swizzled B layout)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_tsmma scripts/ubench_tsmma.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../amuse_b200/csrc/tc_ptx.cuh"

using namespace amuse;
using namespace amuse::tcp;

constexpr int K = 128, NF = 128, NR = 16;          // K, output features, activation rows (10 live)
constexpr int WLD = K + 4;                         // padded row stride of the fp32 weight tile in smem
constexpr uint32_t kIdesc = idesc_tf32(128, NR);
// TMEM columns: A_hi [0,128) A_lo [128,256) D [256,272)
constexpr int kColAlo = 128, kColD = 256;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of element (row r, k) in a K-major SWIZZLE_128B operand: boxes of 32 k (128 B rows),
// 8-row groups 1024 B apart, 16-B chunk index XOR (r & 7)
__device__ __forceinline__ uint32_t b_off(int r, int k) {
  const int box = k >> 5, kk = k & 31;
  const int chunk = (kk >> 2) ^ (r & 7);
  return box * (NR * 128) + (r >> 3) * 1024 + (r & 7) * 128 + chunk * 16 + (kk & 3) * 4;
}

__global__ void __launch_bounds__(256, 1) k_probe(const float* __restrict__ Wg, const float* __restrict__ Xg,
                                                  float* __restrict__ Y, long long* cyc, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Bh = smem;                         // [4 boxes][16 rows][128 B] = 8 KB
  uint8_t* Bl = smem + 8192;
  float* Ws = reinterpret_cast<float*>(smem + 16384);              // [128][WLD] fp32 weight tile
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + NF * WLD * 4);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < NF * K; i += 256) Ws[(i / K) * WLD + (i % K)] = Wg[i];
  for (int i = tid; i < NR * K; i += 256) {
    const int r = i / K, k = i % K;
    float h, l;
    split_tf32(r < 10 ? Xg[r * K + k] : 0.f, h, l);
    *reinterpret_cast<float*>(Bh + b_off(r, k)) = h;
    *reinterpret_cast<float*>(Bl + b_off(r, k)) = l;
  }
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int q = warp & 3, half = warp >> 2;          // TMEM lane quadrant, K half handled by this warp
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);
  const int f = q * 32 + lane;                       // my feature row

  // ---- (1) weight tile smem -> TF32 hi/lo planes in TMEM (A operand: lane = feature, column = k)
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int k0 = half * 64 + c * 32;
      float hi[32], lo[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + f * WLD + k0 + i * 4);
        split_tf32(w.x, hi[i * 4 + 0], lo[i * 4 + 0]);
        split_tf32(w.y, hi[i * 4 + 1], lo[i * 4 + 1]);
        split_tf32(w.z, hi[i * 4 + 2], lo[i * 4 + 2]);
        split_tf32(w.w, hi[i * 4 + 3], lo[i * 4 + 3]);
      }
      tmem_st32(lane_base + k0, hi);
      tmem_st32(lane_base + kColAlo + k0, lo);
    }
    tmem_st_wait();
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cyc[0] = (t1 - t0) / reps;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- (2) chained dependent stages
  const uint64_t d0 = umma_desc(smem_u32(smem));
  uint32_t phase = 0;
  long long tA = clock64(), t_issue = 0, t_wait = 0, t_epi = 0;
  for (int rep = 0; rep < reps; ++rep) {
    long long s0 = clock64();
    if (warp == 1) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < K / 8; ++kk) {
          const uint64_t bh = d0 + (((kk >> 2) * (NR * 128) + (kk & 3) * 32) >> 4);
          const uint64_t bl = bh + (8192 >> 4);
          umma_tf32_ts(tmem + kColD, tmem + kk * 8, bh, kIdesc, kk ? 1u : 0u);
          umma_tf32_ts(tmem + kColD, tmem + kColAlo + kk * 8, bh, kIdesc, 1u);
          umma_tf32_ts(tmem + kColD, tmem + kk * 8, bl, kIdesc, 1u);
        }
        umma_commit(&bar[0]);
      }
      __syncwarp();
    }
    long long s1 = clock64();
    mbar_wait(&bar[0], phase);
    phase ^= 1;
    tc_fence_after();
    long long s2 = clock64();
    if (warp < 4) {   // epilogue: thread = feature; 16 row values; becomes column f of the next B operand
      float v[16];
      tmem_ld16(lane_base + kColD, v);
      if (rep == reps - 1) {
#pragma unroll
        for (int r = 0; r < 10; ++r) Y[r * NF + f] = v[r];
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        float h, l;
        split_tf32(r < 10 ? v[r] * 0.05f : 0.f, h, l);      // keep magnitudes bounded over the chain
        if (rep + 1 < reps) {
          *reinterpret_cast<float*>(Bh + b_off(r, f)) = h;
          *reinterpret_cast<float*>(Bl + b_off(r, f)) = l;
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long s3 = clock64();
    t_issue += s1 - s0;
    t_wait += s2 - s1;
    t_epi += s3 - s2;
  }
  long long tB = clock64();
  if (tid == 32) {
    cyc[1] = (tB - tA) / reps;
    cyc[2] = t_issue / reps;
    cyc[3] = t_wait / reps;
    cyc[4] = t_epi / reps;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  std::vector<float> W(NF * K), X(10 * K);
  srand(3);
  for (auto& v : W) v = (rand() / (float)RAND_MAX - 0.5f) * 0.3f;
  for (auto& v : X) v = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
  float *dW, *dX, *dY;
  long long *dc, hc[8];
  cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dY, 10 * NF * 4); cudaMalloc(&dc, 64);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 16384 + NF * WLD * 4 + 64 + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // numerics: a single stage
  k_probe<<<1, 256, smem>>>(dW, dX, dY, dc, 1);
  std::vector<float> Y(10 * NF);
  cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
  double e = 0, m = 0;
  for (int r = 0; r < 10; ++r)
    for (int f = 0; f < NF; ++f) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)X[r * K + k] * (double)W[f * K + k];
      e = fmax(e, fabs(Y[r * NF + f] - s));
      m = fmax(m, fabs(s));
    }
  printf("stage numerics (Y = X . W^T, 10 x 128 x 128, 3xTF32, A = W from TMEM): max|err| vs fp64 = %.3e (|y|max %.2f)\n", e, m);
  k_probe<<<1, 256, smem>>>(dW, dX, dY, dc, 64);
  cudaMemcpy(hc, dc, 40, cudaMemcpyDeviceToHost);
  printf("(1) 128x128 fp32 weight tile smem -> TF32 hi/lo in TMEM (8 warps, LDS + split + tcgen05.st): %lld cycles\n", hc[0]);
  printf("(2) dependent stage (48 TS MMAs N=16 + commit + wait + tcgen05.ld + B rewrite + fences): %lld cycles\n", hc[1]);
  printf("    of which  MMA issue %lld | wait for completion %lld | epilogue + B rewrite + barrier %lld\n", hc[2], hc[3], hc[4]);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
