// L2 -> shared-memory weight streaming probe for the denoise loop.
// Today every SM of a 4-CTA cluster streams 1/4 of the 8.77 MB of weights per step (1.9 MB per SM per ~60 us,
// 3.9 TB/s over 128 SMs).  A 2-CTA cluster with one clip would exchange 6x fewer bytes over DSMEM but every SM
// would stream 1/2 of the weights (3.8 MB per step; 7.8 TB/s in aggregate at today's step time).  This probe
// answers whether the L2 fabric sustains that: CTAs stream a per-rank blob through the kernel's 2-slot ring of
// 64 KB tiles with cp.async.bulk (the same UBLKCP path), each tile "consumed" by reading it once from shared
// memory with LDS.64 into FFMA2s (the 5-row micro-kernel's load pattern), and the time per pass over the blob is
// reported for fractions 1/4 and 1/2 and for 32 / 64 / 128 / 148 CTAs.  A pass that is faster than ~50 us for the
// 1/2 blob on 128 CTAs means the 2-CTA design is not L2-bound.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_l2stream scripts/ubench_l2stream.cu
#include <cstdio>
#include <cstdlib>

#include "../amuse_b200/csrc/common.cuh"

using namespace amuse;

constexpr int kThr = 320, kTileFloats = 16768;   // the kernel's ring slot (largest tile + tail)

// ROWS = activation rows every weight word is used for (FMA work per tile: 5 = one clip, 10 = two clips)
template <int ROWS>
__global__ void __launch_bounds__(kThr, 1) k_stream(const float* __restrict__ blob, int ranks, int tiles_per_pass, int passes,
                                                    float* sink, long long* cyc) {
  extern __shared__ __align__(128) float sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * kTileFloats);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* mine = blob + static_cast<size_t>(blockIdx.x % ranks) * tiles_per_pass * kTileFloats;
  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t total = static_cast<uint32_t>(tiles_per_pass) * passes;
  auto issue = [&](uint32_t n) {
    mbar_arrive_expect_tx(bars + (n & 1), kTileFloats * 4u);
    bulk_g2s(sm + (n & 1) * kTileFloats, mine + static_cast<size_t>(n % tiles_per_pass) * kTileFloats, kTileFloats * 4u,
             bars + (n & 1));
  };
  if (tid == 0) {
    issue(0);
    if (total > 1) issue(1);
  }
  float2 acc[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) acc[i] = make_float2(0.f, 0.f);
  const long long t0 = clock64();
  for (uint32_t g = 0; g < total; ++g) {
    mbar_wait(bars + (g & 1), (g >> 1) & 1);
    const float* w = sm + (g & 1) * kTileFloats;
    if (warp < 8) {
      // 8 warps split the tile's k-pairs; lane reads 4 x LDS.64 per k-pair like gemm_rows (128-wide tile)
#pragma unroll 4
      for (int kp = warp; kp < 64; kp += 8) {
        const float* wr = w + kp * 256 + lane * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 wv = *reinterpret_cast<const float2*>(wr + 64 * j);
#pragma unroll
          for (int i = 0; i < ROWS; ++i) acc[i] = __ffma2_rn(make_float2(1.0f + i, 0.5f), wv, acc[i]);
        }
      }
    }
    __syncthreads();
    if (tid == 9 * 32 && g + 2 < total) issue(g + 2);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ROWS; ++i) s += acc[i].x + acc[i].y;
  if (s == 123.456f) sink[0] = s;
  if (blockIdx.x == 0 && tid == 0) cyc[0] = t1 - t0;
}

template <int ROWS>
static void run(const float* blob, int ranks, int tiles_per_pass, int grid, float* sink, long long* cyc) {
  const size_t smem = 2 * kTileFloats * sizeof(float) + 64;
  cudaFuncSetAttribute(k_stream<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const int passes = 200;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k_stream<ROWS><<<grid, kThr, smem>>>(blob, ranks, tiles_per_pass, 20, sink, cyc);
  cudaEventRecord(a);
  k_stream<ROWS><<<grid, kThr, smem>>>(blob, ranks, tiles_per_pass, passes, sink, cyc);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double mb = tiles_per_pass * kTileFloats * 4.0 / 1e6;
  printf("rows %2d  ranks %d  %5.2f MB per CTA per pass  grid %3d : %7.2f us / pass  (%6.0f cycles)  %6.2f TB/s aggregate  [%s]\n",
         ROWS, ranks, mb, grid, ms * 1e3 / passes, static_cast<double>(h) / passes, grid * mb / (ms * 1e3 / passes) / 1e6,
         cudaGetErrorString(e));
}

int main() {
  // 8.77 MB of weights = 131 ring tiles of 67 KB; a rank of a 4-CTA cluster consumes ~33 per step, of a 2-CTA cluster ~66
  const int tiles_total = 132;
  float* blob;
  cudaMalloc(&blob, static_cast<size_t>(tiles_total) * kTileFloats * sizeof(float));
  cudaMemset(blob, 0, static_cast<size_t>(tiles_total) * kTileFloats * sizeof(float));
  float* sink;
  long long* cyc;
  cudaMalloc(&sink, 64);
  cudaMalloc(&cyc, 64);
  for (int grid : {32, 64, 128, 148}) {
    run<10>(blob, 4, tiles_total / 4, grid, sink, cyc);   // today: 4 ranks, two clips per cluster
    run<5>(blob, 2, tiles_total / 2, grid, sink, cyc);    // candidate: 2 ranks, one clip per cluster
  }
  return 0;
}
