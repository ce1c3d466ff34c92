// Micro-benchmark: is the legacy warp-level tensor path (mma.sync.m16n8k8 tf32) a faster engine than
// FFMA2 for the denoise loop's 10-row GEMM stages on B200?  One CTA of 8 warps on one SM.
//   (1) raw issue rate of mma.sync.m16n8k8.tf32 with 16 independent accumulator tiles per warp
//   (2) the candidate inner loop: B fragments from shared memory (LDS.64), on-the-fly hi/lo split,
//       3 MMAs (3xTF32) per (k8, n8) tile  -- cycles per stage of K=128 x N=128 with 8-way K split
//   (3) numerics of (2) against fp64 on random data (checks fragment layout + truncation split)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_mma scripts/ubench_mma.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void k_mma_raw(float* out, long long* cyc, int iters) {
  float c[16][4];
  for (int j = 0; j < 16; ++j)
    for (int q = 0; q < 4; ++q) c[j][q] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = threadIdx.x * 11,
           b1 = threadIdx.x * 13;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) mma_tf32(c[j], a0, a1, a2, a3, b0 + j, b1);
  }
  long long t1 = clock64();
  __syncthreads();
  float r = 0;
  for (int j = 0; j < 16; ++j) r += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// Candidate stage: Y[16 x 128] (+)= A[16 x K] . W[K x 128], 8 warps = 8-way K split (16 k per warp).
// Ws: MMA-packed [K/8][16 ntiles][32 lanes][2]  (lane (g,t): W[k0+2t][8j+g], W[k0+2t+1][8j+g])
// As: [16][lda] fp32, rows >= 10 are zero.
constexpr int kLda = 136;
template <bool SPLIT3>
__global__ void k_stage(const float* __restrict__ Wg, const float* __restrict__ Ag, float* __restrict__ Y, long long* cyc,
                        int reps) {
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                  // 128*128
  float* As = sm + 128 * 128;      // 16*kLda
  float* Red = As + 16 * kLda;     // 8 * 16 * 128
  for (int i = threadIdx.x; i < 128 * 128; i += blockDim.x) Ws[i] = Wg[i];
  for (int i = threadIdx.x; i < 16 * kLda; i += blockDim.x) As[i] = Ag[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  float c[16][4];
  long long t0 = clock64();
#pragma unroll 1
  for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[j][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int kstep = warp * 2 + ks;
      const float2 alo2 = *reinterpret_cast<const float2*>(As + g * kLda + kstep * 8 + 2 * t);
      const float2 ahi2 = *reinterpret_cast<const float2*>(As + (g + 8) * kLda + kstep * 8 + 2 * t);
      // a0:(row g, slot t) a1:(row g+8, slot t) a2:(row g, slot t+4) a3:(row g+8, slot t+4)
      const float av[4] = {alo2.x, ahi2.x, alo2.y, ahi2.y};
      uint32_t ah[4], al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ah[q] = __float_as_uint(av[q]) & 0xffffe000u;
        al[q] = __float_as_uint(av[q] - __uint_as_float(ah[q]));
      }
      const float* wr = Ws + (kstep * 16) * 64 + lane * 2;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 w = *reinterpret_cast<const float2*>(wr + j * 64);
        const uint32_t wh0 = __float_as_uint(w.x) & 0xffffe000u, wh1 = __float_as_uint(w.y) & 0xffffe000u;
        mma_tf32(c[j], ah[0], ah[1], ah[2], ah[3], wh0, wh1);
        if (SPLIT3) {
          const uint32_t wl0 = __float_as_uint(w.x - __uint_as_float(wh0)), wl1 = __float_as_uint(w.y - __uint_as_float(wh1));
          mma_tf32(c[j], al[0], al[1], al[2], al[3], wh0, wh1);
          mma_tf32(c[j], ah[0], ah[1], ah[2], ah[3], wl0, wl1);
        }
      }
    }
    // park the K-split partials: rows g (c0,c1) and g+8 (c2,c3), columns 8j+2t, +1
    float* rd = Red + warp * 16 * 128;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      *reinterpret_cast<float2*>(rd + g * 128 + 8 * j + 2 * t) = make_float2(c[j][0], c[j][1]);
      if (g < 2) *reinterpret_cast<float2*>(rd + (g + 8) * 128 + 8 * j + 2 * t) = make_float2(c[j][2], c[j][3]);
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // reduce the 8 partials (rows 0..9)
  for (int i = threadIdx.x; i < 10 * 128; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += Red[w * 16 * 128 + i];
    Y[i] = s;
  }
}

int main() {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  for (int threads : {128, 256, 512}) {
    const int iters = 256;
    k_mma_raw<<<1, threads>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const int wps = threads / 128;
    printf("threads %3d  mma.m16n8k8.tf32 raw: %.2f cyc/mma/SMSP  -> %.0f MAC/clk/SM\n", threads,
           (double)h / (iters * 16.0 * wps), 1024.0 * iters * 16 * (threads / 32) / h);
  }
  // stage benchmark + numerics
  std::vector<float> W(128 * 128), A(10 * 128), Wp(128 * 128), Ap(16 * kLda, 0.f);
  srand(1);
  for (auto& v : W) v = (rand() / (float)RAND_MAX - 0.5f) * 0.3f;
  for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
  for (int ks = 0; ks < 16; ++ks)
    for (int j = 0; j < 16; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        Wp[((ks * 16 + j) * 32 + lane) * 2 + 0] = W[(ks * 8 + 2 * t) * 128 + 8 * j + g];       // W[k][n]
        Wp[((ks * 16 + j) * 32 + lane) * 2 + 1] = W[(ks * 8 + 2 * t + 1) * 128 + 8 * j + g];
      }
  for (int r = 0; r < 10; ++r)
    for (int k = 0; k < 128; ++k) Ap[r * kLda + k] = A[r * 128 + k];
  float *dW, *dA, *dY;
  cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dA, Ap.size() * 4); cudaMalloc(&dY, 10 * 128 * 4);
  cudaMemcpy(dW, Wp.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dA, Ap.data(), Ap.size() * 4, cudaMemcpyHostToDevice);
  const int smem = (128 * 128 + 16 * kLda + 8 * 16 * 128) * 4;
  cudaFuncSetAttribute(k_stage<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_stage<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<double> ref(10 * 128, 0.0);
  for (int r = 0; r < 10; ++r)
    for (int n = 0; n < 128; ++n) {
      double s = 0;
      for (int k = 0; k < 128; ++k) s += (double)A[r * 128 + k] * (double)W[k * 128 + n];
      ref[r * 128 + n] = s;
    }
  std::vector<float> Y(10 * 128);
  for (int mode = 0; mode < 2; ++mode) {
    const int reps = 200;
    if (mode == 0) k_stage<true><<<1, 256, smem>>>(dW, dA, dY, cyc, reps);
    else k_stage<false><<<1, 256, smem>>>(dW, dA, dY, cyc, reps);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
    double e = 0, m = 0;
    for (int i = 0; i < 10 * 128; ++i) { e = fmax(e, fabs(Y[i] - ref[i])); m = fmax(m, fabs(ref[i])); }
    printf("stage K=128 N=128 rows=10 %s: %.0f cyc/stage (incl. park + barrier); max|err| vs fp64 = %.3e (|y|max %.2f)\n",
           mode == 0 ? "3xTF32" : "1xTF32", (double)h / reps, e, m);
  }
  // fp32 FFMA reference error for scale
  {
    double e = 0;
    for (int r = 0; r < 10; ++r)
      for (int n = 0; n < 128; ++n) {
        float s = 0;
        for (int k = 0; k < 128; ++k) s = fmaf(A[r * 128 + k], W[k * 128 + n], s);
        e = fmax(e, fabs(s - ref[r * 128 + n]));
      }
    printf("fp32 fmaf chain max|err| vs fp64 = %.3e\n", e);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
