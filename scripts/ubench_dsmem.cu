// DSMEM exchange probe for the denoise loop: what does the all-to-all exchange of
// K-split partial sums inside a 4-CTA cluster cost as a function of the transport and of the volume?
// The kernel's exchange moves 10 rows x 512 B to each of 3 peers (15 KB out, 15 KB in per CTA) and waits
// ~500 cycles after its own stores -- close to 15 KB / (17..21 B/clk), the DSMEM rate in the microarchitecture
// notes, so it looks bandwidth-bound.  Transports compared (same protocol as the kernel: receive buffers and
// mbarriers alternate with the exchange index, no cluster barrier in the loop):
//   mode 0  st.async .v2.b32 (8 B / thread)   -- what denoise_loop.cu does today, one warp per row
//   mode 1  st.async .v4.b32 (16 B / thread)  -- one instruction per row per peer
//   mode 2  cp.async.bulk shared::cta -> shared::cluster, one 512-B copy per row per peer, issued by lane 0
//           of the row's warp after the row was written to a local staging buffer (+ fence.proxy.async)
//   mode 3  cp.async.bulk, one copy of all rows per peer, issued by one thread after a __syncthreads()
// Volume: ROWS x ROWB bytes per peer, swept over ROWB = 512, 256, 128 (latency vs bandwidth).
// 32 clusters run concurrently (the benchmark grid of the loop kernel).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_dsmem scripts/ubench_dsmem.cu
#include <cstdio>
#include <cstdlib>

#include "../amuse_b200/csrc/common.cuh"

using namespace amuse;

constexpr int kThr = 320, kRows = 10;   // kRows = buffer geometry; NROWS of them are exchanged

__device__ __forceinline__ void st_async_v2(uint32_t dst, float2 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(dst),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t dst, float4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
               "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
               : "memory");
}

// smem: stage [2][10][128] | recv [2][3][10][128] | bars[2].  The staging buffer alternates like the receive
// buffers: an outgoing bulk copy of round xe has been consumed by every peer before any of them can send round
// xe+1, which this CTA waits for before it writes the staging buffer of round xe+2.
// kCl = cluster size (4 = today, 2 = the one-clip-per-2-CTA-cluster candidate), NROWS = rows per exchange (10 / 5)
template <int MODE, int ROWB, int kCl = 4, int NROWS = kRows>
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThr, 1) k_exchange(float* sink, long long* cyc, int reps) {
  extern __shared__ __align__(128) float sm[];
  float* recv = sm + 2 * kRows * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * kRows * 128 + 2 * 3 * kRows * 128);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  constexpr int RF = ROWB / 4;   // floats per row actually exchanged
  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < kRows * 128 * 8; i += kThr) sm[i] = 0.f;
  __syncthreads();
  cluster_sync_all();

  float acc = 0.f;
  long long t_sum = 0;
  for (int xe = 0; xe < reps; ++xe) {
    float* rb = recv + (xe & 1) * (3 * kRows * 128);
    float* stage = sm + (xe & 1) * (kRows * 128);
    uint64_t* bar = bars + (xe & 1);
    __syncthreads();
    const long long t0 = clock64();
    if (tid == 0) mbar_arrive_expect_tx(bar, static_cast<uint32_t>(kCl - 1) * NROWS * ROWB);
    const int row = warp;   // 10 warps, one row each
    const bool owner = row < NROWS;
    // the value a row owner would have after its gather (depends on the previous round so nothing is hoisted)
    const float base = acc * 1e-30f + static_cast<float>(xe + row);
    if (MODE == 0) {
      if (owner && lane * 2 < RF) {
#pragma unroll
        for (uint32_t d = 1; d < kCl; ++d) {
          const uint32_t peer = (rank + d) & (kCl - 1), slot = (rank < peer) ? rank : rank - 1;
          const uint32_t dst = map_to_rank(rb + (slot * kRows + row) * 128 + 2 * lane, peer);
          const uint32_t rbar = map_to_rank(bar, peer);
          st_async_v2(dst, make_float2(base, base + 1.f), rbar);
          if (RF > 64) st_async_v2(dst + 64 * 4, make_float2(base + 2.f, base + 3.f), rbar);
        }
      }
    } else if (MODE == 1) {
      if (owner && lane * 4 < RF) {
#pragma unroll
        for (uint32_t d = 1; d < kCl; ++d) {
          const uint32_t peer = (rank + d) & (kCl - 1), slot = (rank < peer) ? rank : rank - 1;
          const uint32_t dst = map_to_rank(rb + (slot * kRows + row) * 128 + 4 * lane, peer);
          const uint32_t rbar = map_to_rank(bar, peer);
          st_async_v4(dst, make_float4(base, base + 1.f, base + 2.f, base + 3.f), rbar);
        }
      }
    } else {
      // write the row into the local staging buffer, make it visible to the async proxy
      if (owner && lane * 4 < RF) *reinterpret_cast<float4*>(stage + row * RF + 4 * lane) = make_float4(base, base + 1.f, base + 2.f, base + 3.f);
      fence_proxy_async();
      if (MODE == 2) {
        __syncwarp();
        if (owner && lane == 0) {
#pragma unroll
          for (uint32_t d = 1; d < kCl; ++d) {
            const uint32_t peer = (rank + d) & (kCl - 1), slot = (rank < peer) ? rank : rank - 1;
            bulk_s2c(map_to_rank(rb + (slot * kRows + row) * RF, peer), smem_u32(stage + row * RF), ROWB, map_to_rank(bar, peer));
          }
        }
      } else {
        __syncthreads();
        if (tid < kCl - 1) {
          const uint32_t d = tid + 1;
          const uint32_t peer = (rank + d) & (kCl - 1), slot = (rank < peer) ? rank : rank - 1;
          bulk_s2c(map_to_rank(rb + slot * kRows * RF, peer), smem_u32(stage), NROWS * ROWB, map_to_rank(bar, peer));
        }
      }
    }
    mbar_wait(bar, (xe >> 1) & 1);
    const long long t1 = clock64();
    t_sum += t1 - t0;
    // consume what arrived (the kernel's add_peers): the next round's data depends on it
    if (owner && lane * 2 < RF) {
#pragma unroll
      for (int q = 0; q < kCl - 1; ++q) acc += rb[(q * kRows + row) * (MODE >= 2 ? RF : 128) + 2 * lane];
    }
  }
  if (blockIdx.x == 0 && tid == 0) cyc[0] = t_sum;
  if (acc == 123.456f) sink[0] = acc;
  cluster_sync_all();
}

template <int MODE, int ROWB, int kCl = 4, int NROWS = kRows>
static void run(const char* name, float* sink, long long* cyc) {
  const int reps = 2000;
  const size_t smem = (kRows * 128 * 8) * sizeof(float) + 64;
  cudaFuncSetAttribute(k_exchange<MODE, ROWB, kCl, NROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  for (int grid : {kCl, 128}) {
    k_exchange<MODE, ROWB, kCl, NROWS><<<grid, kThr, smem>>>(sink, cyc, 200);
    k_exchange<MODE, ROWB, kCl, NROWS><<<grid, kThr, smem>>>(sink, cyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s row %3d B  (%5d B out/CTA)  clusters %2d : %7.1f cycles / exchange   [%s]\n", name, ROWB, (kCl - 1) * NROWS * ROWB,
           grid / kCl, static_cast<double>(h) / reps, cudaGetErrorString(e));
  }
}

int main() {
  float* sink;
  long long* cyc;
  cudaMalloc(&sink, 64);
  cudaMalloc(&cyc, 64);
  run<0, 512>("st.async.v2 (today)", sink, cyc);
  run<0, 256>("st.async.v2", sink, cyc);
  run<1, 512>("st.async.v4", sink, cyc);
  run<1, 256>("st.async.v4", sink, cyc);
  run<1, 128>("st.async.v4", sink, cyc);
  run<2, 512>("bulk per row (lane 0)", sink, cyc);
  run<2, 256>("bulk per row (lane 0)", sink, cyc);
  run<2, 128>("bulk per row (lane 0)", sink, cyc);
  run<3, 512>("bulk per peer (+syncthreads)", sink, cyc);
  run<3, 256>("bulk per peer (+syncthreads)", sink, cyc);
  run<3, 128>("bulk per peer (+syncthreads)", sink, cyc);
  // the 2-CTA-cluster candidate of DESIGN.md section 6.1a: one peer, 5 rows (one clip)
  run<0, 512, 2, 5>("2-CTA cluster, 5 rows, v2", sink, cyc);
  run<1, 512, 2, 5>("2-CTA cluster, 5 rows, v4", sink, cyc);
  run<0, 512, 4, 5>("4-CTA cluster, 5 rows, v2", sink, cyc);
  return 0;
}
