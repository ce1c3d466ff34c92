// CTA-pair (cta_group::2) persistent tcgen05 GEMM with 3xTF32 accuracy, for the large AST GEMMs
// (M = clips*1214, N in {768, 2304, 3072}, K in {256, 768, 3072}).
//
// Why a second GEMM kernel: with hi/lo operand planes a 128x128 tile needs 64 KB of operands per
// 768 MMA cycles -- 83 B/clk/SM, i.e. ~12 KB/clk chip-wide against an L2 fabric that delivers ~6 KB/clk,
// and 128 B/clk of shared-memory reads per MMA on top of the TMA writes.  ncu on tc_gemm_kernel
// (profiles/r01_ast6_full.txt): tensor pipe 35-61 % busy, limited by operand delivery.  A CTA pair
// computing a 256x256 tile halves the bytes per MMA cycle on both paths: each CTA loads its own 128 rows
// of A and HALF of the W tile (128 of the 256 output columns); tcgen05.mma.cta_group::2 reads A from the
// CTA's own shared memory and W from both.
//
// Structure (per CTA of the pair; cluster = 2 CTAs on one TPC, persistent over a static tile schedule):
//   warp 0      TMA producer: A_hi/A_lo (own rows) + W_hi/W_lo (own column half) per 32-wide K block
//               into a 3-stage ring; bytes of BOTH CTAs are credited to the leader's "full" barrier
//   warp 1      TMEM allocation (2 x 256 accumulator columns, cta_group::2); the LEADER's warp 1 issues
//               12 tcgen05.mma.cta_group::2.kind::tf32 (M256 N256 K8) per K block and multicasts
//               tcgen05.commit to both CTAs' "empty" / "accumulator full" barriers
//   warps 2..5  epilogue on the CTA's own 128 accumulator rows, overlapped with the next tile's main
//               loop through the second accumulator buffer; one output row per thread, global loads / stores
//               transposed through per-warp staging tiles so that they move whole 128-byte lines
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace amuse {
namespace tc {

namespace {

using namespace tcp;

constexpr int BM = 128;           // rows per CTA (256 per pair)
constexpr int BN = 256;           // columns per pair tile; each CTA stages BN/2 rows of W
constexpr int BK = 32;            // fp32 per 128-B swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = 128 * BK * 4;        // 16 KB: A (128 rows) and W half (128 rows) tiles
constexpr int STAGE_BYTES = 4 * TILE_BYTES;     // A_hi, A_lo, W_hi, W_lo  (per CTA)
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;                  // 2 accumulator buffers x 256 columns
constexpr int kBarOff = STAGES * STAGE_BYTES;
constexpr int kStageOff = kBarOff + 256;         // 4 epilogue warps x [32 rows][36] floats: the transposing staging tiles
constexpr int kSmemBytes = kStageOff + 4 * 32 * 36 * 4 + 1024;
constexpr uint32_t kIdesc = idesc_tf32(2 * BM, BN);

// Epilogue of one 32-column chunk of my warp's 32 rows.  A thread owns a ROW of the accumulator, so all global traffic
// is transposed through the warp's staging tile (tc_ptx.cuh: stage_*): whole 128-byte lines per instruction instead of
// 32 lines x 16 bytes -- at K = 768 the epilogue's LSU wavefronts, not the MMAs, bounded the tile (proj GEMM 54 %
// tensor-active, profiles/r01_ast_kernels_full.txt).  `res` = this chunk's residual (hi + lo), fetched one chunk ahead
// in the staging layout (row 4 i + lane / 8, columns 4 (lane % 8) ..).
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmDesc& d, int row0, int rows_valid, int lane, int nc, float (&v)[32],
                                               float* stg, const float4 (&res)[8]) {
  const int m = row0 + lane;
  const bool row_ok = lane < rows_valid;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] += __ldg(d.bias + nc + i);
  if (EPI == EPI_GELU_PLANES) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
  if (EPI == EPI_RES_PLANES) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(stg + (4 * i + (lane >> 3)) * 36 + 4 * (lane & 7)) = res[i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 r4 = *reinterpret_cast<const float4*>(stg + lane * 36 + 4 * i);
      v[i * 4 + 0] += r4.x;
      v[i * 4 + 1] += r4.y;
      v[i * 4 + 2] += r4.z;
      v[i * 4 + 3] += r4.w;
    }
  }
  if (EPI == EPI_QKV_HEADS) {
    const int D = d.heads * 64;
    const int which = nc / D, rem = nc - which * D, hd = rem >> 6, d0 = rem & 63;   // warp-uniform
    if (which == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= d.q_scale;
    }
    if (which < 2) {   // q / k planes [clip x head][token][64]: a warp's rows may straddle two clips -> per-row pointers
      float* const ph = which == 0 ? d.q_hi : d.k_hi;
      float* const pl = which == 0 ? d.q_lo : d.k_lo;
      auto off_of = [&](int r) -> long long {
        if (r >= rows_valid) return -1;
        const int mr = row0 + r, b = mr / d.tok, t = mr - b * d.tok;
        return static_cast<long long>((static_cast<size_t>(b) * d.heads + hd) * d.tokp + t) * 64 + d0;
      };
      store_planes_coalesced_rows<32>(
          stg, lane, v, [&](int r) { const long long o = off_of(r); return o < 0 ? nullptr : ph + o; },
          [&](int r) { const long long o = off_of(r); return o < 0 ? nullptr : pl + o; });
    } else if (row_ok) {   // v transposed: consecutive lanes = consecutive tokens -> coalesced 128-B stores
      const int b = m / d.tok, t = m - b * d.tok;
      const size_t off = ((static_cast<size_t>(b) * d.heads + hd) * 64 + d0) * d.tokp + t;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float h, l;
        split_tf32(v[i], h, l);
        d.vt_hi[off + static_cast<size_t>(i) * d.tokp] = h;
        d.vt_lo[off + static_cast<size_t>(i) * d.tokp] = l;
      }
    }
  } else if (EPI == EPI_PLAIN) {
    __syncwarp();
    stage_put_row<32>(stg, lane, v);
    __syncwarp();
    stage_copy_out<32>(stg, lane, d.C + static_cast<size_t>(row0) * d.ldc + nc, d.ldc, rows_valid);
  } else {
    const size_t off = static_cast<size_t>(row0) * d.ldc + nc;
    store_planes_coalesced<32>(stg, lane, v, d.C_hi + off, d.C_lo + off, d.ldc, rows_valid);
  }
}

}  // namespace

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                    const GemmDesc d, const int tiles_m, const int tiles_n) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* full = bars;                  // [STAGES]  used in the leader: bytes of both CTAs
  uint64_t* empty = bars + STAGES;        // [STAGES]  per CTA, multicast commit from the leader
  uint64_t* acc_full = bars + 2 * STAGES; // [2]       per CTA, multicast commit from the leader
  uint64_t* acc_empty = acc_full + 2;     // [2]       used in the leader: 4 epilogue warps x 2 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();            // 0 = leader
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_tiles = tiles_m * tiles_n;
  const int nkb = d.K / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA_hi);
    prefetch_tmap(&tmA_lo);
    prefetch_tmap(&tmW_hi);
    prefetch_tmap(&tmW_lo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem2_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // both CTAs' barriers are initialised before any remote signal / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    uint32_t it = 0;      // running K-block counter across tiles
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;   // N fastest: the pairs running together share A rows
      const int m0 = tm * (2 * BM) + static_cast<int>(rank) * BM;
      const int n0 = tn * BN + static_cast<int>(rank) * (BN / 2);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        const uint32_t fb = map_to_rank(&full[s], 0);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * STAGE_BYTES);
          tma2_load_2d(st, &tmA_hi, fb, kb * BK, m0);
          tma2_load_2d(st + TILE_BYTES, &tmA_lo, fb, kb * BK, m0);
          tma2_load_2d(st + 2 * TILE_BYTES, &tmW_hi, fb, kb * BK, n0);
          tma2_load_2d(st + 3 * TILE_BYTES, &tmW_lo, fb, kb * BK, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      const uint64_t d0 = umma_desc(smem_u32(smem));
      uint32_t it = 0, lt = 0;   // K-block counter, local tile counter
      for (int tile = pair; tile < n_tiles; tile += n_pairs, ++lt) {
        const uint32_t buf = lt & 1;
        mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);   // both CTAs' epilogues have drained this buffer
        tc_fence_after();
        const uint32_t dacc = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t base = d0 + ((s * STAGE_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint64_t a_hi = base + ((k * 32) >> 4);
              const uint64_t a_lo = base + ((TILE_BYTES + k * 32) >> 4);
              const uint64_t w_hi = base + ((2 * TILE_BYTES + k * 32) >> 4);
              const uint64_t w_lo = base + ((3 * TILE_BYTES + k * 32) >> 4);
              umma2_tf32_ss(dacc, a_hi, w_hi, kIdesc, (kb | k) ? 1u : 0u);
              umma2_tf32_ss(dacc, a_lo, w_hi, kIdesc, 1u);
              umma2_tf32_ss(dacc, a_hi, w_lo, kIdesc, 1u);
            }
            umma2_commit_mc(&empty[s], 3);                       // stage free in both CTAs
            if (kb == nkb - 1) umma2_commit_mc(&acc_full[buf], 3);   // accumulator ready in both CTAs
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5, both CTAs) =====================
    const int q = warp & 3;                // TMEM lanes [32q, 32q+32) are the ones this warp may access
    const uint32_t ae0 = map_to_rank(&acc_empty[0], 0), ae1 = map_to_rank(&acc_empty[1], 0);
    float* stg = reinterpret_cast<float*>(smem + kStageOff) + (warp - 2) * (32 * 36);
    float4 res[8];                         // EPI_RES_PLANES: the residual of the chunk about to be processed
    auto load_res = [&](int row0, int rows_valid, int nc) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + (lane >> 3);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (r < rows_valid) {   // plain loads: C may alias R (in-place residual update; this warp owns these rows / columns)
          const size_t off = static_cast<size_t>(row0 + r) * d.ldr + nc + 4 * (lane & 7);
          a = *reinterpret_cast<const float4*>(d.R_hi + off);
          b = *reinterpret_cast<const float4*>(d.R_lo + off);
        }
        res[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
    };
    uint32_t lt = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, ++lt) {
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
      const int row0 = tm * (2 * BM) + static_cast<int>(rank) * BM + q * 32;   // my warp's first output row
      const int rows_valid = d.M - row0;
      const uint32_t buf = lt & 1;
      if (EPI == EPI_RES_PLANES) load_res(row0, rows_valid, tn * BN);   // under the tile's MMAs
      mbar_wait(&acc_full[buf], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(trow + c * 32, v);
        if (c == BN / 32 - 1) {            // accumulator buffer is in registers: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(buf ? ae1 : ae0);
        }
        float4 cur[8];
        if (EPI == EPI_RES_PLANES) {
#pragma unroll
          for (int i = 0; i < 8; ++i) cur[i] = res[i];
          if (c + 1 < BN / 32) load_res(row0, rows_valid, tn * BN + (c + 1) * 32);   // next chunk's, under this chunk's work
        }
        epilogue_chunk<EPI>(d, row0, rows_valid, lane, tn * BN + c * 32, v, stg, cur);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();      // the peer may still be reading my shared memory / signalling my barriers
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc<kTmemCols>(tmem_base);
  }
}

namespace {

template <int EPI>
cudaError_t launch2(const CUtensorMap* tm, const GemmDesc& d, cudaStream_t st) {
  static PerDeviceOnce once;
  static int pairs_of_dev[64] = {};
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev)) return e;
  if (cudaError_t e = once.run([dev] {
        cudaError_t e2 =
            cudaFuncSetAttribute(tc_gemm2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e2 != cudaSuccess) return e2;
        int sms = 0;
        if ((e2 = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e2;
        pairs_of_dev[dev & 63] = sms / 2;
        return cudaSuccess;
      }))
    return e;
  const int n_pairs_max = pairs_of_dev[dev & 63];
  const int tiles_m = (d.M + 2 * BM - 1) / (2 * BM), tiles_n = d.N / BN;
  const int n_pairs = tiles_m * tiles_n < n_pairs_max ? tiles_m * tiles_n : n_pairs_max;
  tc_gemm2_kernel<EPI><<<2 * n_pairs, kThreads, kSmemBytes, st>>>(tm[0], tm[1], tm[2], tm[3], d, tiles_m, tiles_n);
  return cudaGetLastError();
}

}  // namespace

cudaError_t gemm2(int epi, const GemmDesc& din, cudaStream_t st) {
  GemmDesc d = din;
  if (d.K % BK != 0 || d.M < 1 || d.N < BN || d.N % BN != 0 || d.A2_hi) return cudaErrorInvalidValue;
  if (epi == EPI_QKV_HEADS && (d.N != 3 * d.heads * 64 || d.tok < 1 || d.tokp < d.tok || !d.q_hi || !d.vt_lo))
    return cudaErrorInvalidValue;
  if (epi == EPI_PLAIN && (d.ldc & 3)) return cudaErrorInvalidValue;
  CUtensorMap tm[4];
  cudaError_t e;
  if ((e = make_map_2d(&tm[0], d.A_hi, d.M, d.K, d.lda, 128)) != cudaSuccess) return e;
  if ((e = make_map_2d(&tm[1], d.A_lo, d.M, d.K, d.lda, 128)) != cudaSuccess) return e;
  if ((e = make_map_2d(&tm[2], d.W_hi, d.N, d.K, d.ldw, 128)) != cudaSuccess) return e;
  if ((e = make_map_2d(&tm[3], d.W_lo, d.N, d.K, d.ldw, 128)) != cudaSuccess) return e;
  switch (epi) {
    case EPI_PLAIN: return launch2<EPI_PLAIN>(tm, d, st);
    case EPI_GELU_PLANES: return launch2<EPI_GELU_PLANES>(tm, d, st);
    case EPI_RES_PLANES: return launch2<EPI_RES_PLANES>(tm, d, st);
    case EPI_QKV_HEADS: return launch2<EPI_QKV_HEADS>(tm, d, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace tc
}  // namespace amuse
