// PTX wrappers shared by the tcgen05 kernels (tc_gemm.cu, ast_attn.cu): TMA tensor loads, UMMA
// shared-memory / instruction descriptors, tcgen05.mma (operands from shared memory, or A from
// TMEM), commit, TMEM load / store, TF32 hi/lo split.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace amuse {
namespace tcp {

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
// One lane of a fully converged warp.  Code that issues tcgen05.mma / TMA must sit under warp-uniform
// control flow with this predicate (not under `if (lane == 0)`): the operands of those instructions are
// uniform registers, and in a divergent region the compiler wraps every one of them in an
// ELECT / BRA.U.ANY "waterfall" loop (~50 issue cycles per MMA -- measured in the attention kernel).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T   (A: row m on TMEM lane m, one TF32 element per 32-bit column)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster on one TPC act as one MMA unit ------------
// D[256 x N]: rows [0,128) accumulate in the leader's TMEM, rows [128,256) in the peer's; A = each CTA's
// own 128 rows, B = N/2 rows from each CTA (same shared-memory offsets in both).  Issued by the leader only.
__device__ __forceinline__ void umma2_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this shared-memory offset in
// every CTA of `cta_mask`
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// TMA load into THIS CTA's shared memory; the bytes are credited to the barrier at `bar_cluster_addr`
// (a shared::cluster address -- the pair leader's "full" barrier)
__device__ __forceinline__ void tma2_load_2d(void* smem, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem2_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem2_dealloc(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}

// 32 lanes x 32 columns: thread `lane` of the warp gets columns [col, col+32) of its TMEM lane.
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
// the store twin; call tmem_st_wait() before signalling another thread
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
      "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
      "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
      "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
      "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// hi = round-to-nearest TF32 of x, lo = x - hi (exact in fp32; the MMA keeps its top 11 significand bits, so the
// pair carries ~22 bits of x)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

// ---- coalescing the epilogue's global traffic.  A thread owns an output ROW (its TMEM lane), so direct loads / stores
// touch 32 different 128-byte lines per warp instruction: 32 LSU wavefronts for 512 bytes, and the tile's epilogue was
// bound by exactly that (profiles/r02_decode_kernels_full.txt: 39k cycles per tile, 20k wavefronts).  Instead every
// warp transposes through its own staging tile in shared memory:
// [32 rows][W + 4] floats -- row-per-thread float4 accesses and row-contiguous float4 accesses are both conflict-free.
template <int W>
__device__ __forceinline__ void stage_put_row(float* stg, int lane, const float (&v)[W]) {
#pragma unroll
  for (int i = 0; i < W / 4; ++i)
    *reinterpret_cast<float4*>(stg + lane * (W + 4) + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// rows [row0, row0 + 32) x W columns from the staging tile to dst (row stride ld floats), whole 128-byte lines per instruction
template <int W>
__device__ __forceinline__ void stage_copy_out(const float* stg, int lane, float* dst, int ld, int rows_valid) {
  constexpr int LPR = W / 4, RPI = 32 / LPR;   // lanes per row, rows per instruction
  const int r0 = lane / LPR, c4 = lane % LPR;
#pragma unroll
  for (int i = 0; i < 32 / RPI; ++i) {
    const int r = i * RPI + r0;
    const float4 x = *reinterpret_cast<const float4*>(stg + r * (W + 4) + 4 * c4);
    if (r < rows_valid) *reinterpret_cast<float4*>(dst + static_cast<size_t>(r) * ld + 4 * c4) = x;
  }
}
// v -> TF32 hi / lo planes at (row0.., col0..) of C_hi / C_lo
template <int W>
__device__ __forceinline__ void store_planes_coalesced(float* stg, int lane, const float (&v)[W], float* c_hi, float* c_lo,
                                                       int ld, int rows_valid) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < W / 4; ++i) {   // hi plane (the split is recomputed for the lo plane: cheaper than 64 more live registers)
    float4 h, l;
    split_tf32(v[4 * i + 0], h.x, l.x);
    split_tf32(v[4 * i + 1], h.y, l.y);
    split_tf32(v[4 * i + 2], h.z, l.z);
    split_tf32(v[4 * i + 3], h.w, l.w);
    *reinterpret_cast<float4*>(stg + lane * (W + 4) + 4 * i) = h;
  }
  __syncwarp();
  stage_copy_out<W>(stg, lane, c_hi, ld, rows_valid);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < W / 4; ++i) {
    float4 h, l;
    split_tf32(v[4 * i + 0], h.x, l.x);
    split_tf32(v[4 * i + 1], h.y, l.y);
    split_tf32(v[4 * i + 2], h.z, l.z);
    split_tf32(v[4 * i + 3], h.w, l.w);
    *reinterpret_cast<float4*>(stg + lane * (W + 4) + 4 * i) = l;
  }
  __syncwarp();
  stage_copy_out<W>(stg, lane, c_lo, ld, rows_valid);
}

// rows x 32 columns with no alignment or edge assumption (ld not a multiple of 4, N edge): a warp writes one row's 32
// consecutive floats per instruction
__device__ __forceinline__ void stage_copy_out_scalar32(const float* stg, int lane, float* dst, int ld, int rows_valid,
                                                        int cols_valid) {
  if (lane < cols_valid) {
#pragma unroll 8
    for (int r = 0; r < 32; ++r)
      if (r < rows_valid) dst[static_cast<size_t>(r) * ld + lane] = stg[r * 36 + lane];
  }
}
// the same with a row -> pointer map (rows of a warp that land in different output blocks; nullptr = skip the row)
template <int W, class RowPtr>
__device__ __forceinline__ void stage_copy_out_rows(const float* stg, int lane, RowPtr rowptr) {
  constexpr int LPR = W / 4, RPI = 32 / LPR;
  const int r0 = lane / LPR, c4 = lane % LPR;
#pragma unroll
  for (int i = 0; i < 32 / RPI; ++i) {
    const int r = i * RPI + r0;
    const float4 x = *reinterpret_cast<const float4*>(stg + r * (W + 4) + 4 * c4);
    float* dst = rowptr(r);
    if (dst) *reinterpret_cast<float4*>(dst + 4 * c4) = x;
  }
}
template <int W, class RowPtrHi, class RowPtrLo>
__device__ __forceinline__ void store_planes_coalesced_rows(float* stg, int lane, const float (&v)[W], RowPtrHi hi_row,
                                                            RowPtrLo lo_row) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < W / 4; ++i) {
    float4 h, l;
    split_tf32(v[4 * i + 0], h.x, l.x);
    split_tf32(v[4 * i + 1], h.y, l.y);
    split_tf32(v[4 * i + 2], h.z, l.z);
    split_tf32(v[4 * i + 3], h.w, l.w);
    *reinterpret_cast<float4*>(stg + lane * (W + 4) + 4 * i) = h;
  }
  __syncwarp();
  stage_copy_out_rows<W>(stg, lane, hi_row);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < W / 4; ++i) {
    float4 h, l;
    split_tf32(v[4 * i + 0], h.x, l.x);
    split_tf32(v[4 * i + 1], h.y, l.y);
    split_tf32(v[4 * i + 2], h.z, l.z);
    split_tf32(v[4 * i + 3], h.w, l.w);
    *reinterpret_cast<float4*>(stg + lane * (W + 4) + 4 * i) = l;
  }
  __syncwarp();
  stage_copy_out_rows<W>(stg, lane, lo_row);
}

}  // namespace tcp
}  // namespace amuse
