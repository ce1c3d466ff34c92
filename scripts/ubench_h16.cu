// Probe for the tensor-core denoise loop (round 2): one denoiser GEMM stage as
//   D[128 features x 16 rows] (+)= W[128 x K] (A operand, TMEM-resident) . X[16 x K]^T (B operand, shared memory)
// with an fp16 hi/lo split ("3xFP16": x = hi + lo'/2048, hi = fp16(x), lo' = fp16((x - hi) * 2048);
// x.w ~= hi.hi [acc 0] + 2^-11 (lo'.hi + hi.lo') [acc 1]), kind::f16 => K = 16 per MMA: half the MMAs and half
// the TMEM columns of 3xTF32, and the hi/lo planes together are exactly as many bytes as the fp32 weights.
//   (A) numerics against fp64 for K = 128 / 64 / 32, checks the TMEM A layout (two fp16 per 32-bit column) and
//       the SWIZZLE_128B K-major B layout
//   (B) latency of a dependent stage (24 MMAs + commit + wait + tcgen05.ld + fp16 split + B rewrite + fences),
//       one chain alone, two independent chains on one SM, and both with the weight producers running
//   (C) weight path global(L2) -> registers -> tcgen05.st: per-SM and whole-chip rate with 128 CTAs streaming the
//       four 1.9 MB rank streams of the real kernel (decides whether L2 can feed a ~30 us step)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_h16 scripts/ubench_h16.cu
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../amuse_b200/csrc/tc_ptx.cuh"

using namespace amuse;
using namespace amuse::tcp;

constexpr int NF = 128, NR = 16, ROWS = 5;
constexpr uint32_t kIdesc = (1u << 4) | (static_cast<uint32_t>(NR >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);

__device__ int g_timeout = 0;
__device__ __forceinline__ void mbar_wait_to(uint64_t* bar, uint32_t parity) {
  for (long long i = 0; i < 4000000; ++i)
    if (mbar_try_wait(bar, parity)) return;
  g_timeout = 1;
  __trap();
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void bar_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void tmem_st32u(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void split_h(float x, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn((x - __half2float(h)) * 2048.0f);
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
}
// byte offset of element (row r, k) of a K-major SWIZZLE_128B fp16 operand of 16 rows: boxes of 64 k (128-B rows),
// 8-row groups 1024 B apart, 16-B chunk index XOR (r & 7)
__device__ __host__ __forceinline__ uint32_t b_off(int r, int k) {
  const int box = k >> 6, kk = k & 63;
  const int chunk = (kk >> 3) ^ (r & 7);
  return box * 2048 + (r >> 3) * 1024 + (r & 7) * 128 + chunk * 16 + (kk & 7) * 2;
}

// weight tile (K columns) of one producer thread: batch b = 32 TMEM columns = 8 x 16 B
__device__ __forceinline__ void load_batch(const uint4* src, int q, int lane, int b, uint32_t (&r)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 v = ldg_stream(src + ((b * 4 + q) * 8 + i) * 32 + lane);
    r[i * 4 + 0] = v.x;
    r[i * 4 + 1] = v.y;
    r[i * 4 + 2] = v.z;
    r[i * 4 + 3] = v.w;
  }
}

// =====================================================================================================
// (A)+(B): 12 warps: 0-3 chain 0, 4-7 chain 1, 8-11 weight producers (TMEM lane quadrant = warp & 3)
// TMEM columns: weight slots 3 x 128 at [0,384); D of chain c at [384 + 32c, +32) (acc0 16 | acc1 16)
// =====================================================================================================
struct ChainArgs {
  const uint4* wblob;     // tile in producer layout, K columns
  const float* X;         // [ROWS][K]
  float* Y;               // [2 chains][ROWS][NF]
  long long* cyc;         // stamps
  const uint4* stream;    // producer stream for the interference mode
  int K, reps, chains, producers, stream_batches;
};

__global__ void __launch_bounds__(384, 1) k_chain(const ChainArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // per chain: Bh 4 KB | Bl 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K;
  for (int i = tid; i < 16384 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (tid == 0) {
    mbar_init(&bars[0], 1);   // D ready, chain 0
    mbar_init(&bars[1], 1);   // chain 1
    mbar_init(&bars[2], 4);   // weights ready (4 producer warps)
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int q = warp & 3;
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);

  if (warp >= 8) {
    // ---- producers: the tile for the chains into slot 0, then (interference mode) stream into slots 1, 2
    for (int b = 0; b < K / 32; ++b) {
      uint32_t r[32];
      load_batch(a.wblob, q, lane, b, r);
      tmem_st32u(lane_base + b * 32, r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[2]);
    if (a.producers) {
      long long t0 = clock64();
      uint32_t r0[32], r1[32];
      load_batch(a.stream, q, lane, 0, r0);
      for (int b = 0; b + 1 < a.stream_batches; b += 2) {
        load_batch(a.stream, q, lane, b + 1, r1);
        tmem_st32u(lane_base + 128 + ((b * 32) & 255), r0);
        if (b + 2 < a.stream_batches) load_batch(a.stream, q, lane, b + 2, r0);
        tmem_st32u(lane_base + 128 + (((b + 1) * 32) & 255), r1);
      }
      tmem_st_wait();
      long long t1 = clock64();
      if (tid == 8 * 32) a.cyc[16] = t1 - t0;
    }
  } else {
    const int c = warp >> 2;            // chain
    if (c < a.chains) {
      const int f = q * 32 + lane;      // my feature
      uint8_t* Bh = smem + c * 8192;
      uint8_t* Bl = Bh + 4096;
      for (int i = tid & 127; i < ROWS * K; i += 128) {
        const int r = i / K, k = i % K;
        uint16_t h, l;
        split_h(a.X[r * K + k], h, l);
        *reinterpret_cast<uint16_t*>(Bh + b_off(r, k)) = h;
        *reinterpret_cast<uint16_t*>(Bl + b_off(r, k)) = l;
      }
      fence_proxy_async();
      bar_named(1 + c, 128);
      const uint64_t d0 = umma_desc(smem_u32(Bh));
      const uint32_t colD = tmem + 384 + c * 32;
      uint32_t phase = 0;
      long long t_issue = 0, t_wait = 0, t_epi = 0;
      long long tA = clock64();
      for (int rep = 0; rep < a.reps; ++rep) {
        long long s0 = clock64();
        if (q == 0) {
          if (rep == 0) {
            mbar_wait_to(&bars[2], 0);
            tc_fence_after();
          }
          if (elect_one()) {
            for (int kk = 0; kk < K / 16; ++kk) {
              const uint64_t bh = d0 + (((kk >> 2) * 2048 + (kk & 3) * 32) >> 4);
              const uint64_t bl = bh + (4096 >> 4);
              const uint32_t ah = tmem + kk * 8, al = tmem + K / 2 + kk * 8;
              umma_f16_ts(colD, ah, bh, kIdesc, kk ? 1u : 0u);
              umma_f16_ts(colD + 16, al, bh, kIdesc, kk ? 1u : 0u);
              umma_f16_ts(colD + 16, ah, bl, kIdesc, 1u);
            }
            umma_commit(&bars[c]);
          }
          __syncwarp();
        }
        long long s1 = clock64();
        mbar_wait_to(&bars[c], phase);
        phase ^= 1;
        tc_fence_after();
        long long s2 = clock64();
        float v[32];
        tmem_ld32(lane_base + 384 + c * 32, v);
        float y[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) y[r] = fmaf(v[16 + r], 1.0f / 2048.0f, v[r]);
        if (rep == a.reps - 1) {
#pragma unroll
          for (int r = 0; r < ROWS; ++r) a.Y[(c * ROWS + r) * NF + f] = y[r];
        } else if (f < K) {
#pragma unroll
          for (int r = 0; r < ROWS; ++r) {
            uint16_t h, l;
            split_h(y[r] * 0.05f, h, l);
            *reinterpret_cast<uint16_t*>(Bh + b_off(r, f)) = h;
            *reinterpret_cast<uint16_t*>(Bl + b_off(r, f)) = l;
          }
        }
        fence_proxy_async();
        tc_fence_before();
        bar_named(1 + c, 128);
        tc_fence_after();
        long long s3 = clock64();
        t_issue += s1 - s0;
        t_wait += s2 - s1;
        t_epi += s3 - s2;
      }
      long long tB = clock64();
      if ((tid & 127) == 0) {
        a.cyc[c * 4 + 0] = (tB - tA) / a.reps;
        a.cyc[c * 4 + 1] = t_issue / a.reps;
        a.cyc[c * 4 + 2] = t_wait / a.reps;
        a.cyc[c * 4 + 3] = t_epi / a.reps;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// =====================================================================================================
// (C): weight streaming global(L2) -> registers -> TMEM, PW producer warps per CTA, grid CTAs, `steps` passes over
// the rank stream (rank = blockIdx.x & 3), double-buffered batches of 8 x LDG.128 per thread
// =====================================================================================================
__global__ void __launch_bounds__(256, 1) k_stream(const uint4* blob, size_t rank_vec4, int batches, int steps, long long* cyc) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int q = warp & 3, grp = warp >> 2, ngrp = blockDim.x >> 7;
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);
  const uint4* src = blob + static_cast<size_t>(blockIdx.x & 3) * rank_vec4;
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    uint32_t r0[32], r1[32];
    int b = grp;
    load_batch(src, q, lane, b, r0);
    for (; b < batches; b += 2 * ngrp) {
      if (b + ngrp < batches) load_batch(src, q, lane, b + ngrp, r1);
      tmem_st32u(lane_base + ((b * 32) & 511), r0);
      if (b + 2 * ngrp < batches) load_batch(src, q, lane, b + 2 * ngrp, r0);
      if (b + ngrp < batches) tmem_st32u(lane_base + (((b + ngrp) * 32) & 511), r1);
    }
  }
  tmem_st_wait();
  long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

static uint16_t h_bits(float x) { return __half_as_ushort(__float2half_rn(x)); }
static float h_val(float x) { return __half2float(__float2half_rn(x)); }

// producer layout of a [NF x K] weight tile: TMEM column j < K/2 holds hi(k=2j) | hi(k=2j+1) << 16, column K/2 + j the lo' pair
static std::vector<uint32_t> make_tile(const std::vector<float>& W, int K) {
  std::vector<uint32_t> t(static_cast<size_t>(K / 32) * 4 * 8 * 32 * 4);
  for (int b = 0; b < K / 32; ++b)
    for (int q = 0; q < 4; ++q)
      for (int i = 0; i < 8; ++i)
        for (int lane = 0; lane < 32; ++lane)
          for (int w = 0; w < 4; ++w) {
            const int col = b * 32 + i * 4 + w, f = q * 32 + lane;
            const bool lo = col >= K / 2;
            const int k = 2 * (lo ? col - K / 2 : col);
            uint32_t word = 0;
            for (int e = 0; e < 2; ++e) {
              const float x = W[f * K + k + e];
              const float hi = h_val(x);
              const uint16_t bits = lo ? h_bits((x - hi) * 2048.0f) : h_bits(x);
              word |= static_cast<uint32_t>(bits) << (16 * e);
            }
            t[((((static_cast<size_t>(b) * 4 + q) * 8 + i) * 32) + lane) * 4 + w] = word;
          }
  return t;
}

int main() {
  srand(7);
  long long* dc;
  cudaMalloc(&dc, 4096);
  const int smem = 16384 + 256 + 1024;
  cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // a 1.9 MB-per-rank stream x 4 ranks, L2 resident
  const size_t rank_vec4 = static_cast<size_t>(116) * 4 * 8 * 32;   // 116 batches x 16 KB = 1.9 MB
  uint4* dstream;
  cudaMalloc(&dstream, 4 * rank_vec4 * 16);
  cudaMemset(dstream, 0x11, 4 * rank_vec4 * 16);

  for (int K : {128, 64, 32}) {
    std::vector<float> W(NF * K), X(ROWS * K);
    for (auto& v : W) v = (rand() / (float)RAND_MAX - 0.5f) * 0.3f;
    for (auto& v : X) v = (rand() / (float)RAND_MAX - 0.5f) * 6.f;
    auto tile = make_tile(W, K);
    uint4* dT;
    float *dX, *dY;
    cudaMalloc(&dT, tile.size() * 4);
    cudaMalloc(&dX, X.size() * 4);
    cudaMalloc(&dY, 2 * ROWS * NF * 4);
    cudaMemcpy(dT, tile.data(), tile.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    ChainArgs a{dT, dX, dY, dc, dstream, K, 1, 2, 0, 116};
    k_chain<<<1, 384, smem>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("K=%d numerics launch failed: %s\n", K, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> Y(2 * ROWS * NF);
    cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, err32 = 0, m = 0;
    for (int c = 0; c < 2; ++c)
      for (int r = 0; r < ROWS; ++r)
        for (int f = 0; f < NF; ++f) {
          double s = 0;
          float s32 = 0.f;
          for (int k = 0; k < K; ++k) {
            s += (double)X[r * K + k] * (double)W[f * K + k];
            s32 = fmaf(X[r * K + k], W[f * K + k], s32);
          }
          err = fmax(err, fabs(Y[(c * ROWS + r) * NF + f] - s));
          err32 = fmax(err32, fabs((double)s32 - s));
          m = fmax(m, fabs(s));
        }
    printf("(A) K=%3d: 3xFP16 (A = W in TMEM) max|err| vs fp64 = %.3e   [fp32 FMA chain: %.3e]   |y|max %.2f\n", K, err, err32, m);
    if (K == 128) {
      for (int mode = 0; mode < 4; ++mode) {
        ChainArgs b{dT, dX, dY, dc, dstream, K, 200, (mode & 1) ? 2 : 1, (mode & 2) ? 1 : 0, 116 * 4};
        cudaMemset(dc, 0, 4096);
        k_chain<<<1, 384, smem>>>(b);
        e = cudaDeviceSynchronize();
        long long hc[32];
        cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
        printf("(B) chains=%d producers=%d: stage %lld cycles (issue %lld | wait %lld | ld+split+rewrite+barrier %lld)", b.chains, b.producers,
               hc[0], hc[1], hc[2], hc[3]);
        if (b.chains == 2) printf("  chain1 %lld (%lld | %lld | %lld)", hc[4], hc[5], hc[6], hc[7]);
        if (b.producers) printf("  producer: %.1f B/clk", 116.0 * 4 * 16384 / (double)hc[16]);
        printf("  [%s]\n", cudaGetErrorString(e));
      }
    }
    cudaFree(dT);
    cudaFree(dX);
    cudaFree(dY);
  }
  // (C)
  for (int threads : {128, 256})
    for (int grid : {1, 4, 32, 128, 148}) {
      const int steps = 40;
      k_stream<<<grid, threads>>>(dstream, rank_vec4, 116, 2, dc);   // warm L2
      cudaDeviceSynchronize();
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k_stream<<<grid, threads>>>(dstream, rank_vec4, 116, steps, dc);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> hc(grid);
      cudaMemcpy(hc.data(), dc, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (auto v : hc) mx = v > mx ? v : mx;
      const double bytes = 116.0 * 16384 * steps;
      printf("(C) grid %3d x %d threads: %.1f B/clk/SM (slowest CTA), %.2f TB/s aggregate, %.1f us per 1.9 MB pass  [%s]\n", grid, threads,
             bytes / (double)mx, bytes * grid / (ms * 1e-3) / 1e12, ms * 1e3 / steps, cudaGetErrorString(e));
    }
  return 0;
}
