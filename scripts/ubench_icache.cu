// Instruction-cache probe: straight-line code of N FFMAs (16 B each) executed REPS times by 8 warps.
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void __launch_bounds__(256) k_code(float* out, long long* cyc, float s, int reps) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  float b = s, c = 1.0f - s;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i & 7] = fmaf(a[i & 7], b, a[(i + 3) & 7] * c);   // 2 instr (FMUL+FFMA)
  }
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 8; ++i) acc += a[i];
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int N>
void run(float* out, long long* cyc, int threads) {
  long long h;
  const int reps = 64;
  k_code<N><<<1, threads>>>(out, cyc, 0.999f, reps);
  k_code<N><<<1, threads>>>(out, cyc, 0.999f, reps);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("threads %3d  footprint %4d KB: %.2f cycles per (FMUL+FFMA) pair per warp-slot\n", threads, N * 32 / 1024,
         (double)h / (reps * (double)N));
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  for (int threads : {32, 128, 256}) {
    run<128>(out, cyc, threads);
    run<512>(out, cyc, threads);
    run<1024>(out, cyc, threads);
    run<2048>(out, cyc, threads);
    run<4096>(out, cyc, threads);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
