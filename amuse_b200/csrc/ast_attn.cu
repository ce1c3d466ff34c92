// AST self-attention on tcgen05 / TMEM / TMA with 3xTF32 accuracy (see ast_attn.cuh).
//
// One CTA = one (clip, head, block of 256 queries) = two 128-row query tiles that share every
// key/value tile (halves the L2 -> SM traffic per MMA; with one tile per CTA the K/V stream alone
// would need ~42 B/clk/SM, the whole L2 fabric).  Key/value tiles are 32 keys wide.
//   warp 0      TMA producer: per key tile K_hi, K_lo (32 x 64) and V^T_hi, V^T_lo (64 x 32) into a
//               7-stage ring (32 KB / stage) -- shared memory holds nothing else
//   warp 1      TMEM allocation + MMA issue (one elected lane).  EVERY MMA takes its A operand from TMEM:
//                 S_t  = Q_t . K^T    3 x 8 tcgen05.mma M128 N32 K8, A = Q_hi / Q_lo (stored in TMEM once
//                                     by the softmax warps), B = K_hi / K_lo from shared memory
//                 O_t += P_t . V      3 x 4 tcgen05.mma M128 N64 K8, A = P_hi / P_lo, B = V^T_hi / V^T_lo
//               An MMA whose A operand comes from shared memory costs ~60 cycles here whatever N is (the
//               4 KB A slice is read at ~64 B/clk: ncu source view of the first version, whose 24 score MMAs
//               per tile were of that kind) against 16 for the TMEM form.
//               Issue order  PV_0(j) S_0(j+1) PV_1(j) S_1(j+1):  P_t(j) lives in the TMEM columns of S_t
//               (P_hi over S, P_lo beside it), and the in-order tensor pipe makes S_t(j+1) overwrite them
//               only after PV_t(j) has consumed them; the softmax of one tile overlaps the MMAs of the other.
//               One barrier per direction: s_full[t] (commit after S_t(j), which also implies PV_t(j-1) is
//               complete, so P may be rewritten and O rescaled) and p_ready[t].
//   warps 2-5   softmax of query tile 0, warps 6-9 of query tile 1: one query row per thread
//               (TMEM lane = row): tcgen05.ld S -> online softmax in the log2 domain (q is pre-scaled
//               by 64^-0.5 * log2 e) -> P split into TF32 hi/lo planes -> tcgen05.st into TMEM.
//               The running maximum is only raised (and O rescaled through tcgen05.ld/st) when it
//               grows by more than 2^8 -- exact after the final division by the row sum.
// TMEM (512 columns): O_t 2 x 64 | S_t/P_t 2 x 64 | Q_hi 2 x 64 | Q_lo 2 x 64.
// 3xTF32: x = hi + lo; x.y ~= hi.hi + lo.hi + hi.lo with fp32 accumulation in TMEM.
#include "ast_attn.cuh"

#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace amuse {
namespace attn {

namespace {

using namespace tcp;

constexpr int kThreads = 320;
constexpr int kKT = 32;                         // keys per tile
constexpr int kNT = (kTok + kKT - 1) / kKT;     // 38 key tiles
constexpr int kStages = 7;
constexpr int kKBox = kKT * 32 * 4;             // 32 keys x 32 columns: 4 KB
constexpr int kVBox = kHD * kKT * 4;            // 64 d x 32 keys: 8 KB
constexpr int kStageBytes = 4 * kKBox + 2 * kVBox;   // 32 KB
constexpr int kBarOff = kStages * kStageBytes;
constexpr int kSmemBytes = kBarOff + 256 + 1024;
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");

// TMEM columns
constexpr uint32_t kColO = 0;      // O_t  : [t*64, +64)
constexpr uint32_t kColSP = 128;   // S_t  : [128 + t*64, +32);  P_t: hi over S_t, lo at +32
constexpr uint32_t kColQh = 256;   // Q_hi : [256 + t*64, +64)
constexpr uint32_t kColQl = 384;   // Q_lo : [384 + t*64, +64)
constexpr int kTmemCols = 512;

constexpr uint32_t kIdescS = idesc_tf32(128, kKT);
constexpr uint32_t kIdescPV = idesc_tf32(128, kHD);

}  // namespace
// debug timeline (amuse_debug_attn_profile): clock64 stamps of CTA (0,0,0) for key tiles j in [8, 12)
__device__ long long* g_prof = nullptr;
namespace {
#define ATTN_PROF(slot)                                                                      \
  do {                                                                                       \
    if (prof && j >= 8 && j < 12) prof[((j - 8) * 16 + (slot))] = clock64();                 \
  } while (0)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
    ast_attention_kernel(const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
                         const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo,
                         const float* __restrict__ q_hi, const float* __restrict__ q_lo, float* __restrict__ o_hi,
                         float* __restrict__ o_lo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* q_ready = bars;                  // [1]  softmax warps -> MMA: Q_hi / Q_lo of both tiles are in TMEM
  uint64_t* s_full = bars + 1;               // [2]  MMA -> softmax t: S_t(j) complete (and PV_t(j-1) before it)
  uint64_t* p_ready = bars + 3;              // [2]  softmax t -> MMA: P_t(j) stored (and O_t rescaled)
  uint64_t* o_done = bars + 5;               // [1]  MMA -> softmax: the last PV of both tiles is complete
  uint64_t* kv_full = bars + 6;              // [kStages]
  uint64_t* kv_empty = kv_full + kStages;    // [kStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + kStages);
  static_assert((6 + 2 * kStages) * 8 + 4 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, head = blockIdx.y, clip = blockIdx.z;
  const int bh = clip * kHeads + head;
  long long* const prof = (g_prof && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0) ? g_prof : nullptr;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmK_hi);
    prefetch_tmap(&tmK_lo);
    prefetch_tmap(&tmV_hi);
    prefetch_tmap(&tmV_lo);
    mbar_init(q_ready, 8);
    mbar_init(o_done, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_ready[t], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp in uniform control flow, one elected lane issues) =====
    const int krow = bh * kTokP, vrow = bh * kHD;
    for (int j = 0; j < kNT; ++j) {
      const int s = j % kStages;
      mbar_wait(&kv_empty[s], ((j / kStages) & 1) ^ 1);
      uint8_t* st = smem + s * kStageBytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[s], kStageBytes);
        tma_load_2d(st + 0 * kKBox, &tmK_hi, &kv_full[s], 0, krow + j * kKT);
        tma_load_2d(st + 1 * kKBox, &tmK_hi, &kv_full[s], 32, krow + j * kKT);
        tma_load_2d(st + 2 * kKBox, &tmK_lo, &kv_full[s], 0, krow + j * kKT);
        tma_load_2d(st + 3 * kKBox, &tmK_lo, &kv_full[s], 32, krow + j * kKT);
        tma_load_2d(st + 4 * kKBox, &tmV_hi, &kv_full[s], j * kKT, vrow);
        tma_load_2d(st + 4 * kKBox + kVBox, &tmV_lo, &kv_full[s], j * kKT, vrow);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp in uniform control flow, one elected lane issues) =====
    // shared-memory descriptors differ only in the 14-bit start-address field (bytes >> 4); every operand
    // offset is a compile-time constant from the 1024-B aligned base, so they are formed by one 64-bit add
    const uint64_t d0 = umma_desc(smem_u32(smem));
    auto issue_S = [&](int t, int s) {
      const uint64_t k0 = d0 + ((s * kStageBytes) >> 4);
      const uint32_t dS = tmem_base + kColSP + t * 64;
      const uint32_t aQh = tmem_base + kColQh + t * kHD, aQl = tmem_base + kColQl + t * kHD;
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t b_hi = k0 + ((kb * kKBox + k * 32) >> 4);
            const uint64_t b_lo = k0 + (((2 + kb) * kKBox + k * 32) >> 4);
            umma_tf32_ts(dS, aQh + kb * 32 + k * 8, b_hi, kIdescS, (kb | k) ? 1u : 0u);
            umma_tf32_ts(dS, aQl + kb * 32 + k * 8, b_hi, kIdescS, 1u);
            umma_tf32_ts(dS, aQh + kb * 32 + k * 8, b_lo, kIdescS, 1u);
          }
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_PV = [&](int t, int s, bool first) {
      const uint64_t v0 = d0 + ((s * kStageBytes + 4 * kKBox) >> 4);
      const uint32_t dO = tmem_base + kColO + t * kHD;
      const uint32_t aP = tmem_base + kColSP + t * 64;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t b_hi = v0 + ((k * 32) >> 4);
          const uint64_t b_lo = v0 + ((kVBox + k * 32) >> 4);
          umma_tf32_ts(dO, aP + k * 8, b_hi, kIdescPV, (!first || k) ? 1u : 0u);
          umma_tf32_ts(dO, aP + kKT + k * 8, b_hi, kIdescPV, 1u);
          umma_tf32_ts(dO, aP + k * 8, b_lo, kIdescPV, 1u);
        }
      }
      __syncwarp();
    };
    mbar_wait(q_ready, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_S(0, 0);
    issue_S(1, 0);
    for (int j = 0; j < kNT; ++j) {
      const int s = j % kStages, s1 = (j + 1) % kStages;
      const bool more = j + 1 < kNT;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        ATTN_PROF(8 + t * 4 + 0);            // MMA warp: about to wait for P_t(j)
        mbar_wait(&p_ready[t], j & 1);
        ATTN_PROF(8 + t * 4 + 1);            //           P_t(j) ready
        tc_fence_after();
        issue_PV(t, s, j == 0);
        ATTN_PROF(8 + t * 4 + 2);            //           PV_t(j) issued
        if (more) {
          if (t == 0) {
            mbar_wait(&kv_full[s1], ((j + 1) / kStages) & 1);
            tc_fence_after();
          }
          issue_S(t, s1);       // overwrites P_t(j): the tensor pipe runs it after PV_t(j) above
          ATTN_PROF(8 + t * 4 + 3);          //           S_t(j+1) issued
        }
      }
      if (elect_one()) {
        umma_commit(&kv_empty[s]);          // stage s is free once S(j), PV(j) of both tiles have read it
        if (!more) umma_commit(o_done);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    const int t = (warp - 2) >> 2;          // query tile
    const int q = warp & 3;                 // TMEM lane quadrant this warp may address
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t aS = lane_base + kColSP + t * 64;
    const uint32_t aO = lane_base + kColO + t * kHD;
    {   // Q_hi / Q_lo row of this thread -> TMEM (pad rows of the planes are zero)
      const size_t qoff = (static_cast<size_t>(bh) * kTokP + qblk * 256 + t * 128 + q * 32 + lane) * kHD;
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        const float4* src = reinterpret_cast<const float4*>((pl ? q_lo : q_hi) + qoff);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = __ldg(src + c * 8 + i);
            v[i * 4 + 0] = x.x;
            v[i * 4 + 1] = x.y;
            v[i * 4 + 2] = x.z;
            v[i * 4 + 3] = x.w;
          }
          tmem_st32(lane_base + (pl ? kColQl : kColQh) + t * kHD + c * 32, v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
    }
    float m = 0.f, l = 0.f;
    for (int j = 0; j < kNT; ++j) {
      if (warp == 2) ATTN_PROF(0);          // softmax warp 2 (tile 0): waiting for S_0(j)
      mbar_wait(&s_full[t], j & 1);         // S_t(j) complete; so is PV_t(j-1): P_t and O_t are ours
      if (warp == 2) ATTN_PROF(1);          //   S_0(j) complete
      tc_fence_after();
      float sc[32];
      tmem_ld32(aS, sc);
      if (warp == 2) ATTN_PROF(2);          //   S in registers
      if (j == kNT - 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (j * kKT + i >= kTok) sc[i] = -INFINITY;
      }
      float mx = sc[0];
#pragma unroll
      for (int i = 1; i < 32; ++i) mx = fmaxf(mx, sc[i]);
      const bool raise = (j == 0) || (mx > m + 8.0f);
      const float m_use = raise ? mx : m;
      float ph[32], pl[32], ls = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float p = ex2(sc[i] - m_use);
        ls += p;
        split_tf32(p, ph[i], pl[i]);
      }
      if (j > 0 && __any_sync(0xffffffffu, raise)) {
        const float f = raise ? ex2(m - m_use) : 1.0f;
        l *= f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float o[32];
          tmem_ld32(aO + c * 32, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= f;
          tmem_st32(aO + c * 32, o);
        }
      }
      l += ls;
      m = m_use;
      if (warp == 2) ATTN_PROF(3);          //   P computed (and O rescaled)
      tmem_st32(aS, ph);
      tmem_st32(aS + kKT, pl);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[t]);
      if (warp == 2) ATTN_PROF(4);          //   P stored, p_ready signalled
    }
    mbar_wait(o_done, 0);
    tc_fence_after();
    const int tok = qblk * 256 + t * 128 + q * 32 + lane;
    const float inv = 1.0f / l;
    const size_t off = (static_cast<size_t>(clip) * kTok + (tok < kTok ? tok : 0)) * (kHeads * kHD) + head * kHD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float o[32];
      tmem_ld32(aO + c * 32, o);
      if (tok < kTok) {
        float4* dh = reinterpret_cast<float4*>(o_hi + off + c * 32);
        float4* dl = reinterpret_cast<float4*>(o_lo + off + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 h, lo4;
          split_tf32(o[i * 4 + 0] * inv, h.x, lo4.x);
          split_tf32(o[i * 4 + 1] * inv, h.y, lo4.y);
          split_tf32(o[i * 4 + 2] * inv, h.z, lo4.z);
          split_tf32(o[i * 4 + 3] * inv, h.w, lo4.w);
          dh[i] = h;
          dl[i] = lo4;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

cudaError_t attention(const AttnArgs& a, cudaStream_t st) {
  static PerDeviceOnce once;
  if (cudaError_t e = once.run([] {
        return cudaFuncSetAttribute(ast_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      }))
    return e;
  if (a.nb < 1) return cudaErrorInvalidValue;
  CUtensorMap tm[4];
  const int qk_rows = a.nb * kHeads * kTokP, v_rows = a.nb * kHeads * kHD;
  cudaError_t e;
  if ((e = tc::make_map_2d(&tm[0], a.k_hi, qk_rows, kHD, kHD, kKT)) != cudaSuccess) return e;
  if ((e = tc::make_map_2d(&tm[1], a.k_lo, qk_rows, kHD, kHD, kKT)) != cudaSuccess) return e;
  if ((e = tc::make_map_2d(&tm[2], a.vt_hi, v_rows, kTokP, kTokP, kHD)) != cudaSuccess) return e;
  if ((e = tc::make_map_2d(&tm[3], a.vt_lo, v_rows, kTokP, kTokP, kHD)) != cudaSuccess) return e;
  dim3 grid(kTokP / 256, kHeads, a.nb);
  ast_attention_kernel<<<grid, kThreads, kSmemBytes, st>>>(tm[0], tm[1], tm[2], tm[3], a.q_hi, a.q_lo, a.o_hi, a.o_lo);
  return cudaGetLastError();
}

cudaError_t debug_profile(int enable, long long* host_out, int n) {
  static long long* dbuf = nullptr;
  cudaError_t e;
  if (enable) {
    if (!dbuf) {
      if ((e = cudaMalloc(&dbuf, 64 * sizeof(long long))) != cudaSuccess) return e;
    }
    if ((e = cudaMemset(dbuf, 0, 64 * sizeof(long long))) != cudaSuccess) return e;
    return cudaMemcpyToSymbol(g_prof, &dbuf, sizeof(dbuf));
  }
  long long* null = nullptr;
  if ((e = cudaMemcpyToSymbol(g_prof, &null, sizeof(null))) != cudaSuccess) return e;
  if (dbuf && host_out && n > 0)
    return cudaMemcpy(host_out, dbuf, sizeof(long long) * (n < 64 ? n : 64), cudaMemcpyDeviceToHost);
  return cudaSuccess;
}

}  // namespace attn
}  // namespace amuse
