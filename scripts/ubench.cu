// Micro-benchmarks that ground the denoise-loop design on B200: issue rate of FFMA vs packed
// FFMA2, and shared-memory wavefront cost of warp-uniform (broadcast) vs contiguous LDS.128.
// One CTA on one SM, clock64() around the measured region.   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 512

__global__ void k_ffma(float* out, long long* cyc, float s) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
  float b = s, c = s * 0.5f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], c, b);
  }
  long long t1 = clock64();
  __syncthreads();
  float r = 0;
  for (int i = 0; i < 8; ++i) r += a[i];
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// 3 distinct register operands per FFMA (like a GEMM inner loop: acc += a * w)
__global__ void k_ffma3(float* out, long long* cyc, float s) {
  float acc[16], a[4], w[4];
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int i = 0; i < 4; ++i) { a[i] = threadIdx.x * 0.001f + i + s; w[i] = s * (i + 1); }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i * 4 + j] = fmaf(a[i], w[j], acc[i * 4 + j]);
  }
  long long t1 = clock64();
  __syncthreads();
  float r = 0;
  for (int i = 0; i < 16; ++i) r += acc[i];
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_ffma2_3(float* out, long long* cyc, float s) {
  float2 acc[16], a[4], w[4];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
  for (int i = 0; i < 4; ++i) { a[i] = make_float2(threadIdx.x * 0.001f + i + s, s); w[i] = make_float2(s * (i + 1), s + i); }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i * 4 + j] = __ffma2_rn(a[i], w[j], acc[i * 4 + j]);
  }
  long long t1 = clock64();
  __syncthreads();
  float r = 0;
  for (int i = 0; i < 16; ++i) r += acc[i].x + acc[i].y;
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>   // 0: LDS.128 warp-uniform  1: LDS.128 contiguous  2: LDS.64 contiguous  3: LDS.32 contiguous  4: LDS.32 uniform
__global__ void k_lds(float* out, long long* cyc) {
  __shared__ __align__(16) float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  float r = 0.f;
  const int lane = threadIdx.x & 31;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int base = ((it * 8 + u) * 128) & 4095;
      if (MODE == 0) { float4 v = *reinterpret_cast<const float4*>(&sm[base]); r += v.x + v.w; }
      if (MODE == 1) { float4 v = *reinterpret_cast<const float4*>(&sm[base + lane * 4]); r += v.x + v.w; }
      if (MODE == 2) { float2 v = *reinterpret_cast<const float2*>(&sm[base + lane * 2]); r += v.x + v.y; }
      if (MODE == 3) { r += sm[base + lane]; }
      if (MODE == 4) { r += sm[base]; }
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// the gemm5 inner loop in isolation: 5 uniform A rows + 4 contiguous W float4 per 4 k, FFMA or FFMA2
template <bool PACKED>
__global__ void k_gemm5(float* out, long long* cyc) {
  __shared__ __align__(16) float As[5 * 128];
  __shared__ __align__(16) float Ws[128 * 64];
  for (int i = threadIdx.x; i < 5 * 128; i += blockDim.x) As[i] = i * 1e-3f;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) Ws[i] = i * 1e-4f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[5][4];
  float2 acc2[5][4];
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; acc2[i][j] = make_float2(0.f, 0.f); }
  const int k0 = (warp & 1) * 32;
  long long t0 = clock64();
#pragma unroll 1
  for (int rep = 0; rep < 64; ++rep) {
#pragma unroll
    for (int kk = 0; kk < 32; kk += 4) {
      float4 a[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) a[i] = *reinterpret_cast<const float4*>(&As[i * 128 + k0 + kk]);
      if (!PACKED) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 t = *reinterpret_cast<const float4*>(&Ws[(k0 + kk + k4) * 128 + lane * 4]);
          const float wv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float av = (k4 == 0) ? a[i].x : (k4 == 1) ? a[i].y : (k4 == 2) ? a[i].z : a[i].w;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av, wv[j], acc[i][j]);
          }
        }
      } else {
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          const float* wr = &Ws[((k0 + kk) / 2 + kp) * 256 + lane * 8];
          const float4 t0v = *reinterpret_cast<const float4*>(wr);
          const float4 t1v = *reinterpret_cast<const float4*>(wr + 4);
          const float2 wv[4] = {make_float2(t0v.x, t0v.y), make_float2(t0v.z, t0v.w), make_float2(t1v.x, t1v.y), make_float2(t1v.z, t1v.w)};
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float2 av = (kp == 0) ? make_float2(a[i].x, a[i].y) : make_float2(a[i].z, a[i].w);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[i][j] = __ffma2_rn(av, wv[j], acc2[i][j]);
          }
        }
      }
    }
  }
  long long t1 = clock64();
  float r = 0;
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 4; ++j) r += acc[i][j] + acc2[i][j].x + acc2[i][j].y;
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  for (int threads : {128, 256, 512}) {
    int wps = threads / 128;   // warps per SMSP
    k_ffma<<<1, threads>>>(out, cyc, 1.0001f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  FFMA (a=a*b+c, 2 reg src + imm?)  %.2f cyc/warp-instr/SMSP\n", threads, (double)h / (ITERS * 16.0 * wps));
    k_ffma3<<<1, threads>>>(out, cyc, 1.0001f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  FFMA 3-reg (acc+=a*w)              %.2f cyc/warp-instr/SMSP\n", threads, (double)h / (ITERS * 16.0 * wps));
    k_ffma2_3<<<1, threads>>>(out, cyc, 1.0001f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  FFMA2 3-reg                        %.2f cyc/warp-instr/SMSP\n", threads, (double)h / (ITERS * 16.0 * wps));
  }
  for (int threads : {32, 256}) {
    int warps = threads / 32;
    k_lds<0><<<1, threads>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  LDS.128 warp-uniform  %.2f cyc/warp-instr/SM\n", threads, (double)h / (ITERS * 8.0 * warps));
    k_lds<1><<<1, threads>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  LDS.128 contiguous    %.2f cyc/warp-instr/SM\n", threads, (double)h / (ITERS * 8.0 * warps));
    k_lds<2><<<1, threads>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  LDS.64  contiguous    %.2f cyc/warp-instr/SM\n", threads, (double)h / (ITERS * 8.0 * warps));
    k_lds<3><<<1, threads>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  LDS.32  contiguous    %.2f cyc/warp-instr/SM\n", threads, (double)h / (ITERS * 8.0 * warps));
    k_lds<4><<<1, threads>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  LDS.32  warp-uniform  %.2f cyc/warp-instr/SM\n", threads, (double)h / (ITERS * 8.0 * warps));
  }
  for (int threads : {128, 256, 512}) {
    k_gemm5<false><<<1, threads>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  gemm5 FFMA : %lld cyc for 64 x (8 iter of 4k)  -> %.1f cyc per 4k-iteration per warp-slot; FMA/clk/SM = %.1f\n", threads, h,
           (double)h / (64 * 8), (double)(64 * 8 * 80 * 32) * (threads / 32) / h);
    k_gemm5<true><<<1, threads>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  gemm5 FFMA2: %lld cyc                            -> %.1f cyc per 4k-iteration; FMA/clk/SM = %.1f\n", threads, h,
           (double)h / (64 * 8), (double)(64 * 8 * 80 * 32) * (threads / 32) / h);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
