// tcgen05 / TMEM / TMA GEMM with 3xTF32 accuracy (see tc_gemm.cuh).
//
// One CTA computes one 128 x 128 output tile.  192 threads, warp-specialised:
//   warp 0      TMA producer: per 32-wide K block four 2-D tensor-map loads (A_hi, A_lo, W_hi, W_lo;
//               128 rows x 128 B boxes, SWIZZLE_128B) into a 3-stage shared-memory ring, completion
//               on the stage's "full" mbarrier (complete_tx::bytes)
//   warp 1      TMEM allocation (128 columns) + MMA issue: one elected lane issues, per 8-wide k step,
//               tcgen05.mma.cta_group::1.kind::tf32  D += A_hi.W_hi, A_lo.W_hi, A_hi.W_lo
//               (operands straight from shared memory through UMMA descriptors), then
//               tcgen05.commit -> the stage's "empty" mbarrier; after the last K block
//               tcgen05.commit -> "accumulator ready"
//   warps 2..9  epilogue: two warps per 32 TMEM lanes (= output rows; a warp may address lanes 32 (warp_id % 4) ..),
//               each taking 64 of the tile's 128 columns, tcgen05.ld 32 columns at a time; a thread holds half an
//               output row, so bias / GELU / residual need no cross-thread step and LayerNorm exchanges two partial
//               sums per row with its partner warp through shared memory (named barrier per lane quadrant).  One warp
//               per scheduler with whole rows (round 1) left the epilogue latency-bound: 39k cycles per tile against
//               ~5k for loads + MMAs (profiles/r02_decode_kernels_full.txt)
// Accuracy: hi = cvt.rna.tf32(x), lo = x - hi (exact); dropping lo.lo leaves ~2^-21 relative error
// per product, accumulated in fp32 in TMEM.
#include "tc_gemm.cuh"

#include <cudaTypedefs.h>

#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace amuse {
namespace tc {

namespace {

constexpr int BM = 128, BN = 128, BK = 32;        // BK fp32 = 128 B = one swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;           // 16 KB (A and W tiles have the same shape)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;       // A_hi, A_lo, W_hi, W_lo
constexpr int kThreads = 320;                    // TMA warp + MMA warp + 8 epilogue warps
constexpr int kTmemCols = 128;
constexpr int kRedFloats = 4 * 2 * 128;          // LayerNorm partial sums: [exchange][column half][row]
constexpr int kSmemBytes = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + kRedFloats * 4;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// UMMA shared-memory descriptor: K-major operand, SWIZZLE_128B, 8-row groups 1024 B apart.
// Bit layout = cute::UMMA::SmemDescriptor (start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout_type [61,64), SWIZZLE_128B = 2).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=B=TF32 [7,10)/[10,13)=2,
// both K-major, N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

using tcp::stage_copy_out;
using tcp::stage_put_row;
using tcp::store_planes_coalesced;

__device__ __forceinline__ void bar_pair(int q) { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); }

// LayerNorm of a 128-wide row held as two halves of 64 by the two epilogue warps of a lane quadrant (two-pass mean /
// biased variance); `red` = [2 exchanges][2 halves][128 rows] floats.  g, b point at this half's 64 entries.
__device__ __forceinline__ void row_layernorm_halves(float (&v)[64], const float* __restrict__ g, const float* __restrict__ b,
                                                     float eps, float* red, int half, int row, int q) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += v[i];
  red[half * 128 + row] = s;
  bar_pair(q);
  const float mean = (red[row] + red[128 + row]) * (1.0f / 128.0f);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    v[i] -= mean;
    sq = fmaf(v[i], v[i], sq);
  }
  red[256 + half * 128 + row] = sq;
  bar_pair(q);
  const float rstd = rsqrtf((red[256 + row] + red[384 + row]) * (1.0f / 128.0f) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 gg = __ldg(g4 + i), bb = __ldg(b4 + i);
    v[4 * i + 0] = v[4 * i + 0] * rstd * gg.x + bb.x;
    v[4 * i + 1] = v[4 * i + 1] * rstd * gg.y + bb.y;
    v[4 * i + 2] = v[4 * i + 2] * rstd * gg.z + bb.z;
    v[4 * i + 3] = v[4 * i + 3] * rstd * gg.w + bb.w;
  }
}

}  // namespace

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmA2_hi, const __grid_constant__ CUtensorMap tmA2_lo,
                   const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                   const GemmDesc d) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* acc_ready = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  float* red = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x walks the N tiles of one M block: the CTAs that share an A tile are launched together, so
  // the (large, streamed) A planes are read from HBM once and hit L2 for the other N tiles, while the
  // (small) W planes stay L2-resident.  (With M fastest the A planes were re-read once per N tile:
  // 3.0 GB of DRAM reads for fc2 -- ncu, profiles/r01_tc_gemm_ast_full.txt.)
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
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
    mbar_init(acc_ready, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (whole warp in uniform control flow, one elected lane issues: under `if (lane == 0)` the compiler wraps
    //  every UTMALDG / UTCHMMA in an ELECT waterfall loop, ~50 issue cycles per instruction)
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
      uint8_t* st = smem + s * STAGE_BYTES;
      const int k0 = kb * BK;
      if (tcp::elect_one()) {
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        if (k0 < d.k_split) {
          tma_load_2d(st, &tmA_hi, &full[s], k0, m0);
          tma_load_2d(st + TILE_BYTES, &tmA_lo, &full[s], k0, m0);
        } else {
          tma_load_2d(st, &tmA2_hi, &full[s], k0 - d.k_split, m0);
          tma_load_2d(st + TILE_BYTES, &tmA2_lo, &full[s], k0 - d.k_split, m0);
        }
        tma_load_2d(st + 2 * TILE_BYTES, &tmW_hi, &full[s], k0, n0);
        tma_load_2d(st + 3 * TILE_BYTES, &tmW_lo, &full[s], k0, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint64_t d0 = umma_desc(smem_u32(smem));
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full[s], (kb / STAGES) & 1);
      tc_fence_after();
      if (tcp::elect_one()) {
        const uint64_t base = d0 + ((s * STAGE_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {   // UMMA_K = 8 for TF32: advance 32 B inside the 128-B swizzle row
          const uint64_t a_hi = base + ((k * 32) >> 4);
          const uint64_t a_lo = base + ((TILE_BYTES + k * 32) >> 4);
          const uint64_t w_hi = base + ((2 * TILE_BYTES + k * 32) >> 4);
          const uint64_t w_lo = base + ((3 * TILE_BYTES + k * 32) >> 4);
          umma_tf32(tmem_base, a_hi, w_hi, (kb | k) ? 1u : 0u);
          umma_tf32(tmem_base, a_lo, w_hi, 1u);
          umma_tf32(tmem_base, a_hi, w_lo, 1u);
        }
        umma_commit(&empty[s]);            // frees the stage once the MMAs above have read it
        if (kb == nkb - 1) umma_commit(acc_ready);   // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                // TMEM lanes [32q, 32q+32) are the ones this warp may access
    const int half = (warp - 2) >> 2;      // my 64 of the tile's 128 columns
    const int m = m0 + q * 32 + lane;      // my output row
    const bool row_ok = m < d.M;
    const int rows_valid = d.M - (m0 + q * 32);                         // of my warp's 32 rows (may exceed 32)
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 68);   // my staging tile (operand ring, after acc_ready)
    // LayerNorm epilogues: the residual rows (hi + lo) of my 32 x 64 block, fetched with whole-line loads while the
    // mainloop runs (the epilogue warps are idle until the accumulator is complete)
    float4 rr[(EPI == EPI_RES_LN_PLANES || EPI == EPI_RES_LN_CROSS_LN_PLANES) ? 16 : 1];
    if (EPI == EPI_RES_LN_PLANES || EPI == EPI_RES_LN_CROSS_LN_PLANES) {
      const int r0 = lane >> 4, c4 = lane & 15;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = 2 * i + r0;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (r < rows_valid) {
          const size_t off = static_cast<size_t>(m0 + q * 32 + r) * d.ldr + half * 64 + 4 * c4;
          a = __ldg(reinterpret_cast<const float4*>(d.R_hi + off));
          b = __ldg(reinterpret_cast<const float4*>(d.R_lo + off));
        }
        rr[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
    }
    mbar_wait(acc_ready, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    if (EPI == EPI_RES_LN_PLANES || EPI == EPI_RES_LN_CROSS_LN_PLANES) {
      float v[64];
      const int c0 = half * 64;            // N == 128: the tile is the whole row
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float t[32];
        tmem_ld32(trow + c0 + c * 32, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[c * 32 + i] = t[i];
      }
      const size_t mr = row_ok ? m : 0;
      const int row = q * 32 + lane;
      {
        const int r0 = lane >> 4, c4 = lane & 15;
#pragma unroll
        for (int i = 0; i < 16; ++i) *reinterpret_cast<float4*>(stg + (2 * i + r0) * 68 + 4 * c4) = rr[i];
      }
      __syncwarp();
      const float4* b4 = reinterpret_cast<const float4*>(d.bias + c0);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 r4 = *reinterpret_cast<const float4*>(stg + lane * 68 + 4 * i), bi = __ldg(b4 + i);
        v[i * 4 + 0] += bi.x + r4.x;
        v[i * 4 + 1] += bi.y + r4.y;
        v[i * 4 + 2] += bi.z + r4.z;
        v[i * 4 + 3] += bi.w + r4.w;
      }
      row_layernorm_halves(v, d.ln_g + c0, d.ln_b + c0, d.ln_eps, red, half, row, q);
      if (EPI == EPI_RES_LN_CROSS_LN_PLANES) {
        const float4* cv = reinterpret_cast<const float4*>(d.cvec + static_cast<size_t>(mr / d.rows_per_clip) * 128 + c0);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 c4 = __ldg(cv + i);
          v[i * 4 + 0] += c4.x;
          v[i * 4 + 1] += c4.y;
          v[i * 4 + 2] += c4.z;
          v[i * 4 + 3] += c4.w;
        }
        row_layernorm_halves(v, d.ln2_g + c0, d.ln2_b + c0, d.ln_eps, red + 512, half, row, q);
      }
      {
        const size_t off = static_cast<size_t>(m0 + q * 32) * d.ldc + c0;
        store_planes_coalesced<64>(stg, lane, v, d.C_hi + off, d.C_lo + off, d.ldc, rows_valid);
      }
    } else {
#pragma unroll 1
      for (int c = half * 2; c < half * 2 + 2; ++c) {
        const int nc = n0 + c * 32;
        if (nc >= d.N) break;                    // warp-uniform
        float v[32];
        tmem_ld32(trow + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += (nc + i < d.N) ? __ldg(d.bias + nc + i) : 0.f;
        if (EPI == EPI_QKV) {
          if (nc < d.q_cols) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= d.q_scale;
          }
        }
        if (EPI == EPI_GELU_PLANES) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        }
        if (EPI == EPI_RES_PLANES) {
          const size_t mr = row_ok ? m : 0;
          const float4* rh = reinterpret_cast<const float4*>(d.R_hi + mr * d.ldr + nc);
          const float4* rl = reinterpret_cast<const float4*>(d.R_lo + mr * d.ldr + nc);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a = rh[i], b = rl[i];   // plain loads: C may alias R (in-place residual update)
            v[i * 4 + 0] += a.x + b.x;
            v[i * 4 + 1] += a.y + b.y;
            v[i * 4 + 2] += a.z + b.z;
            v[i * 4 + 3] += a.w + b.w;
          }
        }
        if (EPI == EPI_QKV_HEADS) {
          if (row_ok) {
          const int D = d.heads * 64;
          const int which = nc / D, rem = nc - which * D, hd = rem >> 6, d0 = rem & 63;   // warp-uniform
          const int b = m / d.tok, t = m - b * d.tok;
          const size_t bh = static_cast<size_t>(b) * d.heads + hd;
          if (which == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= d.q_scale;
          }
          if (which < 2) {
            const size_t off = (bh * d.tokp + t) * 64 + d0;
            float4* oh = reinterpret_cast<float4*>((which == 0 ? d.q_hi : d.k_hi) + off);
            float4* ol = reinterpret_cast<float4*>((which == 0 ? d.q_lo : d.k_lo) + off);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 h, l;
              split_tf32(v[i * 4 + 0], h.x, l.x);
              split_tf32(v[i * 4 + 1], h.y, l.y);
              split_tf32(v[i * 4 + 2], h.z, l.z);
              split_tf32(v[i * 4 + 3], h.w, l.w);
              oh[i] = h;
              ol[i] = l;
            }
          } else {   // v transposed: consecutive lanes = consecutive tokens -> coalesced 128-B stores
            const size_t off = (bh * 64 + d0) * d.tokp + t;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float h, l;
              split_tf32(v[i], h, l);
              d.vt_hi[off + static_cast<size_t>(i) * d.tokp] = h;
              d.vt_lo[off + static_cast<size_t>(i) * d.tokp] = l;
            }
          }
          }
        } else if (EPI == EPI_PLAIN || EPI == EPI_QKV) {
          if (nc + 32 <= d.N && (d.ldc & 3) == 0) {
            __syncwarp();
            stage_put_row<32>(stg, lane, v);
            __syncwarp();
            stage_copy_out<32>(stg, lane, d.C + static_cast<size_t>(m0 + q * 32) * d.ldc + nc, d.ldc, rows_valid);
          } else {   // the 333-wide feature rows of the final projection: unaligned rows, ragged last chunk
            __syncwarp();
            stage_put_row<32>(stg, lane, v);
            __syncwarp();
            tcp::stage_copy_out_scalar32(stg, lane, d.C + static_cast<size_t>(m0 + q * 32) * d.ldc + nc, d.ldc, rows_valid,
                                         d.N - nc);
          }
        } else {
          const size_t off = static_cast<size_t>(m0 + q * 32) * d.ldc + nc;
          store_planes_coalesced<32>(stg, lane, v, d.C_hi + off, d.C_lo + off, d.ldc, rows_valid);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------- host side
namespace {

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

cudaError_t load_encode() {
  if (g_encode) return cudaSuccess;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess) return e;
  if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return cudaSuccess;
}

// row-major fp32 [rows][cols] with leading dimension ld (floats): box = 32 cols x 128 rows, 128-B swizzle
cudaError_t make_map(CUtensorMap* tm, const float* ptr, int rows, int cols, int ld, int box_rows = BM) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 3)) return cudaErrorInvalidValue;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 4};
  const cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int EPI>
cudaError_t launch(const CUtensorMap* tm, const GemmDesc& d, cudaStream_t st) {
  static PerDeviceOnce once;
  if (cudaError_t e = once.run([] {
        return cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      }))
    return e;
  dim3 grid((d.N + BN - 1) / BN, (d.M + BM - 1) / BM);
  tc_gemm_kernel<EPI><<<grid, kThreads, kSmemBytes, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], d);
  return cudaGetLastError();
}

__global__ void split_planes_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                                    size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    float h, l;
    split_tf32(src[i], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

}  // namespace

cudaError_t make_map_2d(CUtensorMap* tm, const float* ptr, int rows, int cols, int ld, int box_rows) {
  cudaError_t e = load_encode();
  if (e != cudaSuccess) return e;
  if (box_rows < 8 || box_rows > 256 || (box_rows & 7)) return cudaErrorInvalidValue;
  return make_map(tm, ptr, rows, cols, ld, box_rows);
}

cudaError_t gemm(int epi, const GemmDesc& din, cudaStream_t st) {
  GemmDesc d = din;
  if (d.K % BK != 0 || d.M < 1 || d.N < 1) return cudaErrorInvalidValue;
  if (!d.A2_hi) d.k_split = d.K;
  if (d.k_split % BK != 0) return cudaErrorInvalidValue;
  if ((epi == EPI_RES_LN_PLANES || epi == EPI_RES_LN_CROSS_LN_PLANES) && d.N != 128) return cudaErrorInvalidValue;
  if (d.ln_eps == 0.f) d.ln_eps = kLnEps;
  if (epi == EPI_QKV_HEADS && (d.N != 3 * d.heads * 64 || d.tok < 1 || d.tokp < d.tok || !d.q_hi || !d.vt_lo))
    return cudaErrorInvalidValue;
  cudaError_t e = load_encode();
  if (e != cudaSuccess) return e;
  CUtensorMap tm[6];
  if ((e = make_map(&tm[0], d.A_hi, d.M, d.k_split, d.lda)) != cudaSuccess) return e;
  if ((e = make_map(&tm[1], d.A_lo, d.M, d.k_split, d.lda)) != cudaSuccess) return e;
  if (d.A2_hi) {
    if ((e = make_map(&tm[2], d.A2_hi, d.M, d.K - d.k_split, d.lda2)) != cudaSuccess) return e;
    if ((e = make_map(&tm[3], d.A2_lo, d.M, d.K - d.k_split, d.lda2)) != cudaSuccess) return e;
  } else {
    tm[2] = tm[0];
    tm[3] = tm[1];
  }
  if ((e = make_map(&tm[4], d.W_hi, d.N, d.K, d.ldw)) != cudaSuccess) return e;
  if ((e = make_map(&tm[5], d.W_lo, d.N, d.K, d.ldw)) != cudaSuccess) return e;
  switch (epi) {
    case EPI_PLAIN: return launch<EPI_PLAIN>(tm, d, st);
    case EPI_QKV: return launch<EPI_QKV>(tm, d, st);
    case EPI_PLANES: return launch<EPI_PLANES>(tm, d, st);
    case EPI_GELU_PLANES: return launch<EPI_GELU_PLANES>(tm, d, st);
    case EPI_RES_LN_PLANES: return launch<EPI_RES_LN_PLANES>(tm, d, st);
    case EPI_RES_LN_CROSS_LN_PLANES: return launch<EPI_RES_LN_CROSS_LN_PLANES>(tm, d, st);
    case EPI_RES_PLANES: return launch<EPI_RES_PLANES>(tm, d, st);
    case EPI_QKV_HEADS: return launch<EPI_QKV_HEADS>(tm, d, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t split_planes(const float* src, float* hi, float* lo, size_t n, cudaStream_t st) {
  split_planes_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(src, hi, lo, n);
  return cudaGetLastError();
}

void split_host(const float* src, float* hi, float* lo, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    uint32_t u;
    std::memcpy(&u, &src[i], 4);
    // round to nearest (ties away, like cvt.rna) on the 13 dropped mantissa bits
    uint32_t h = (u + 0x1000u) & 0xFFFFE000u;
    if ((u & 0x7F800000u) == 0x7F800000u) h = u;   // inf / nan untouched
    float hf;
    std::memcpy(&hf, &h, 4);
    hi[i] = hf;
    lo[i] = src[i] - hf;
  }
}

}  // namespace tc
}  // namespace amuse
