// K1+K2 on the tensor cores: the whole N-step denoising loop of PretrainedLPDM_v1.diffusion_backward
// (reference infer_ldm.py:142-161) as ONE persistent launch, every GEMM a tcgen05.mma.
//
// What one step computes (reference denoiser.py:135-204, cross_attention.py:41-64,259-272):
//   x = [z | time token | con | emo | sty] + learned PE        (<= 5 tokens x 128 per clip)
//   9 post-LN encoder layers with U-Net skips, final LayerNorm, eps = token 0
//   scheduler update of z (DDIM eta / DDPM ancestral), clamp(x0) optional
//
// B200 mapping.  The loop is a chain of ~41 dependent small GEMMs per step with M = 5 rows per clip: it is bound
// by the latency of that chain, and every evaluation needs all 8.8 MB of weights.
//   * One clip per cluster of 2 CTAs (64 clips = 128 SMs, one wave up to 74 clips).  The layers split 4 ways as
//     in the FFMA kernel (QKV / FFN1 by output feature = attention head / hidden slice, out_proj / FFN2 / skip-linear
//     by input feature = K-split); CTA r owns the two "ranks" 2r and 2r+1, so a K-split stage exchanges its partial
//     sums with ONE peer (2.5 KB through distributed shared memory) instead of three.  (Measured first: 4-CTA
//     clusters with two one-clip chains sharing an SM -- 60.7 us/step: the 7.7 KB exchanges ran at the 21 B/clk
//     DSMEM port rate and the two chains slowed each other by 26%.)
//   * GEMM orientation: D[128 output features x rows] = W[128 x K] . X[rows x K]^T with the WEIGHTS as the A operand
//     in TMEM (M = 128 lanes) and the 5 activation rows as the shared-memory B operand (an MMA whose A operand is
//     read from shared memory costs ~60 cycles whatever N is, a TMEM one ~19).  The accumulator comes back with
//     tcgen05.ld as "thread f holds feature f of all rows", and the whole epilogue chain (bias, residual, exchange,
//     LayerNorm, GELU, next B operand) stays in that layout: no transposition through shared memory.
//   * fp32-class accuracy from fp16 operands ("3xFP16"): x = hi + lo'/2048 with hi = fp16(x), lo' = fp16((x-hi)*2048);
//       W.x ~= W_hi.x_hi  +  2^-11 (W_hi.x_lo' + W_lo'.x_hi)
//     The B operand carries x_hi in rows 0..15 and x_lo' in rows 16..31, so ONE N = 32 MMA with A = W_hi yields
//     W_hi.x_hi (columns 0..15) and W_hi.x_lo' (columns 16..31), and one N = 16 MMA with A = W_lo' adds W_lo'.x_hi
//     onto columns 16..31: 2 MMAs per 16 k (3xTF32 needs 6).  Measured error of a K = 128 stage against fp64:
//     1.5e-6, the fp32 FMA chain 1.9e-6 (profiles/r02_ubench_h16.txt).  The hi/lo' planes are together exactly as
//     many bytes as the fp32 weights.  Operands saturate at +-65504 (cvt.satfinite).
//   * Warp roles (17 warps):
//       0-7   epilogue warps (q = TMEM lane quadrant, t = group).  N-split stages: group t drains the accumulator of
//             rank 2r+t (its head / hidden slice, all 5 rows) as soon as THAT tile's MMAs commit; K-split stages: one
//             accumulator, group 0 takes rows 0..2 and group 1 rows 3..4 (exchange, residual, LayerNorm per group).
//       8-15  weight producers: the CTA's two rank streams (2 x 1.86 MB per step, pre-split fp16 planes laid out at
//             amuse_finalize_weights in read order) from L2 with coalesced LDG.128 straight into a 3-slot TMEM ring
//             with tcgen05.st -- no shared-memory staging (measured 100 B/clk/SM, 20 TB/s chip-wide,
//             profiles/r02_ubench_h16.txt).  full[slot] / empty[slot] mbarriers, empty signalled by tcgen05.commit.
//       16    MMA issuer: pre-waits the weight tiles, waits for "B operand ready" (one arrive per epilogue warp),
//             issues both tiles of the stage and commits to the groups' accumulator barriers.
//   * K-split stages end in an exchange of partial sums with the peer CTA: st.async into its shared memory, bytes
//     credited to ITS mbarrier (no cluster barrier inside the loop); receive buffers alternate per exchange.
//   * LayerNorm in the feature-per-thread layout: per-warp shifted sums with a transposing butterfly, the 4 warps of a
//     group merged with Chan's parallel-variance formula by one lane per row.
#include "denoise_tc.cuh"

#include <cuda_fp16.h>

#include <mutex>

#include "common.cuh"
#include "philox.cuh"
#include "tc_ptx.cuh"

namespace amuse {
namespace dn2 {

namespace {

using namespace tcp;

constexpr int kRows = kTMax;
constexpr int kNR = 3;                            // rows per thread in the row-split (K-split) epilogues: group 0 rows 0..2, group 1 rows 3..4
constexpr uint32_t kSlotCols = 128, kSlots = 3;
constexpr uint32_t kColD = kSlots * kSlotCols;   // accumulator i in columns [kColD + 32 i, +32)
constexpr uint32_t kColSide = kColD + 64;        // two 32-column side slots for the out_proj tiles (K = 32): they do not
                                                 // occupy a 128-column ring slot, so the ring runs one big tile further ahead
constexpr int kTmemCols = 512;
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {   // cute::UMMA::InstrDescriptor: D = F32, A = B = F16, K-major
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
constexpr uint32_t kIdescN32 = idesc_f16(128, 32), kIdescN16 = idesc_f16(128, 16);

// ---- shared memory (bytes from the 1024-B aligned base).  B operands: K-major SWIZZLE_128B, 32 rows (x_hi rows 0..15,
// x_lo' rows 16..31), one 4 KB box per 64 input features.  Three buffers, so that a group that finishes its tile of an
// N-split stage early never writes the operand the other tile's MMAs are still reading.
constexpr int kQkvLd = 100;                       // q|k|v row stride (floats)
constexpr int oBx = 0;                            // [2 boxes] K = 128: the token rows (input of QKV / FFN1 / skip-linear)
constexpr int oBo = oBx + 8192;                   // [1 box]   K = 64: attention output of my two heads (input of out_proj)
constexpr int oBh = oBo + 4096;                   // [4 boxes] K = 256: my two hidden slices after GELU (input of FFN2)
// receive slot of the peer's partial sums: rows (0,1) as [128 features][2] | rows (3,4) as [128][2] | row 2 as [128]:
// every st.async of a warp writes one contiguous 256-B / 128-B run of the peer's shared memory (a feature-major
// [128][5] slot -- 8-B stores 24 B apart -- made every exchange 4x slower, measured)
constexpr int kPsSlot = 5 * 128 * 4;              // 2560
constexpr int oPs = oBh + 16384;                  // [2 buffers][slot]
constexpr int oQKV = oPs + 2 * kPsSlot;           // [2 heads][5][100] floats
constexpr int oSK = oQKV + 4096;                  // [4][5][128] floats: skip stack of the input blocks
constexpr int oStat = oSK + 4 * kRows * 128 * 4;  // [3 buffers][2 groups][4 warps][16] floats (see Chain::layernorm)
constexpr int oBars = oStat + 3 * 512;
constexpr int kSmemUsed = oBars + 256;
constexpr int kSmemBytes = 120 * 1024;            // > half of the SM's shared memory: one CTA (= one 512-column TMEM allocation) per SM
static_assert(oBo % 1024 == 0 && oBh % 1024 == 0, "B operands must be 1024-B aligned");
static_assert(kSmemUsed + 1024 <= kSmemBytes, "shared-memory carve-up");
constexpr uint32_t kXchgBytes = kPsSlot;          // what one exchange delivers: the peer's 128 features x 5 rows


// ---- small PTX helpers
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bar_named(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
// streaming 16-B load at base + OFF bytes (immediate offset: one address register pair serves all loads of a tile)
template <int OFF>
__device__ __forceinline__ uint4 ldg_stream(const uint4* base) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4+%5];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(base), "n"(OFF));
  return v;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4 (&r)[4]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0].x), "r"(r[0].y), "r"(r[0].z), "r"(r[0].w), "r"(r[1].x), "r"(r[1].y), "r"(r[1].z), "r"(r[1].w), "r"(r[2].x),
      "r"(r[2].y), "r"(r[2].z), "r"(r[2].w), "r"(r[3].x), "r"(r[3].y), "r"(r[3].z), "r"(r[3].w)
      : "memory");
}
// Columns [col, col+8) and [col+16, col+24) of my TMEM lane.  The registers are only defined after tcgen05.wait::ld,
// so they are threaded through the wait as read-write operands: nothing that reads them can be scheduled above it.
__device__ __forceinline__ void tmem_ld_2x8(uint32_t taddr, float (&a)[8], float (&b)[8]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr + 16)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(r[8 + i]);
  }
}
__device__ __forceinline__ void st_async_v2(uint32_t dst, float a, float b, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(dst),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t dst, float a, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst),
               "r"(__float_as_uint(a)), "r"(mbar)
               : "memory");
}
// fp16 hi / lo' planes of x (saturating: the operands of the MMAs must stay finite)
__device__ __forceinline__ void split_h(float x, uint16_t& hi, uint16_t& lo) {
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(hi) : "f"(x));
  const float r = (x - __half2float(__ushort_as_half(hi))) * 2048.0f;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(lo) : "f"(r));
}

struct Ctx {
  const Params* p;
  uint8_t* smem;
  uint64_t* bars;
  uint32_t tmem;
  int* status;
  __device__ __forceinline__ uint64_t* full(uint32_t s) const { return bars + s; }
  __device__ __forceinline__ uint64_t* empty(uint32_t s) const { return bars + 3 + s; }
  __device__ __forceinline__ uint64_t* dbar(int i) const { return bars + 6 + i; }
  __device__ __forceinline__ uint64_t* xbar(uint32_t i) const { return bars + 8 + i; }
  __device__ __forceinline__ uint64_t* sfull(int i) const { return bars + 11 + i; }    // side slots: the two out_proj tiles
  __device__ __forceinline__ uint64_t* sempty(int i) const { return bars + 13 + i; }
};
// mbarrier wait (plain try_wait: the warp is suspended by the hardware for a bounded time per probe.  A suspend-time
// hint is NOT used: ptxas turns it into try_wait + NANOSLEEP(hint) + re-check, which added microseconds to every
// producer hand-off -- measured).  Bounded: a protocol bug must end the launch (trap -> the host sees a launch
// failure), never hang the GPU.
__device__ __forceinline__ void wait_timeout(const Ctx& k) {
  if (k.status) *k.status = 1;
  __trap();
}
// 4 instructions per probe (the C++ loop around mbar_try_wait compiled to 8, and the probes of the waiting warps are a
// visible share of all issued instructions: profiles/r02_denoise_tc_lines.txt)
__device__ __forceinline__ void wait_bar(const Ctx& k, uint64_t* bar, uint32_t parity) {
  uint32_t expired;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
      "mov.u32 c, 0;\n"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "add.u32 c, c, 1;\n\t"
      "setp.lt.u32 p, c, 16777216;\n\t"
      "@p bra WAIT_%=;\n\t"
      "mov.u32 %0, 1;\n\t"
      "bra END_%=;\n"
      "DONE_%=:\n\t"
      "mov.u32 %0, 0;\n"
      "END_%=:\n\t}"
      : "=r"(expired)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  if (expired) wait_timeout(k);
}

// a wait that is not on the critical path (the producers run ahead of the MMAs): sleep between probes
__device__ __forceinline__ void wait_bar_relaxed(const Ctx& k, uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    if (mbar_try_wait(bar, parity)) return;
    __nanosleep(100);
  }
  wait_timeout(k);
}

// ---------------------------------------------------------------- weight producers (8 warps)
// The CTA's weight stream is one linear sequence of tiles (stage-major, rank 2r then 2r+1); tile g goes to TMEM slot
// g % 3.  Each warp owns one TMEM lane quadrant (32 features) and every second 16-column unit of the tile: 4 x LDG.128
// per unit and thread (512 contiguous bytes per warp instruction), all of the warp's units of a tile in flight at once,
// then one tcgen05.st.x16 per unit.  Straight-line code per tile shape: the first version (one generic predicated loop)
// executed ~200 instructions per warp and tile, a third of everything the SM issued (profiles/r02_denoise_tc_lines.txt).
template <int UNITS, int QUADS>   // UNITS = my units in the tile, QUADS = feature quadrants stored
__device__ __forceinline__ void produce_tile(const Ctx& k, const uint4* src, int q, int grp, uint32_t dst, uint64_t* empty_bar,
                                             uint32_t empty_parity, bool wait_empty) {
  if (q >= QUADS) {   // the q|k|v tile has 96 features: nothing for the warps of the fourth quadrant
    if (wait_empty) wait_bar_relaxed(k, empty_bar, empty_parity);
    return;
  }
  constexpr int S = 2 * QUADS * 128 * 16;   // bytes between two of my units
  const uint4* s = src + (grp * QUADS + q) * 128;
  // (the "no loads" / "no tcgen05.st" timing experiments of profiles/r02_sanitizer_summary.txt were runtime flags here;
  //  their zero-initialised registers and predicated loads were 9 % of everything the SM issued, so they are now a
  //  developer build: -DAMUSE_DN2_PRODUCER_DEBUG)
#ifdef AMUSE_DN2_PRODUCER_DEBUG
  const int dbg = k.p->debug_flags;
  uint4 a[4] = {}, b[4] = {}, c[4] = {}, d[4] = {};
  if (!(dbg & 2))
#else
  uint4 a[4], b[4], c[4], d[4];
#endif
  {
#define DN2_LOAD(R, J)                           \
  R[0] = ldg_stream<(J) * S>(s);                 \
  R[1] = ldg_stream<(J) * S + 512>(s);           \
  R[2] = ldg_stream<(J) * S + 1024>(s);          \
  R[3] = ldg_stream<(J) * S + 1536>(s);
  DN2_LOAD(a, 0)
  if constexpr (UNITS > 1) { DN2_LOAD(b, 1) }
  if constexpr (UNITS > 2) { DN2_LOAD(c, 2) DN2_LOAD(d, 3) }
#undef DN2_LOAD
  }
  if (wait_empty) wait_bar_relaxed(k, empty_bar, empty_parity);   // the MMAs on the previous occupant are complete
  tc_fence_after();
#ifdef AMUSE_DN2_PRODUCER_DEBUG
  if (dbg & 4) {   // timing experiment: loads only
    asm volatile("" ::"r"(a[0].x ^ a[3].w ^ b[0].x ^ b[3].w ^ c[0].x ^ c[3].w ^ d[0].x ^ d[3].w));
    return;
  }
#endif
  tmem_st16(dst + grp * 16, a);
  if constexpr (UNITS > 1) tmem_st16(dst + (grp + 2) * 16, b);
  if constexpr (UNITS > 2) {
    tmem_st16(dst + (grp + 4) * 16, c);
    tmem_st16(dst + (grp + 6) * 16, d);
  }
}
struct ProducerState {
  const uint4* src;             // my lane's position in the CTA's weight stream
  uint32_t slot = 0, use = 0;   // ring slot and use count of the next ring tile
  uint32_t suse = 0;            // use count of the side slots
};
// the two tiles (ranks 2r, 2r + 1) of one stage, specialised on the tile kind like the issuer's stages
template <int KIND>
__device__ __forceinline__ void produce_stage(const Ctx& k, ProducerState& st, int q, int grp, int lane, uint32_t lane_base,
                                              bool off) {
  constexpr int UNITS = (KIND == kWO) ? 1 : (KIND == kSK) ? 2 : 4;   // my units of the tile
  constexpr int QUADS = (KIND == kQKV) ? 3 : 4;
#pragma unroll
  for (int t = 0; t < kVirt; ++t) {
    if (KIND == kWO) {
      if (off) {
        if (st.suse > 0) wait_bar_relaxed(k, k.sempty(t), (st.suse - 1) & 1);
      } else {
        produce_tile<UNITS, QUADS>(k, st.src, q, grp, lane_base + kColSide + t * 32, k.sempty(t), (st.suse - 1) & 1,
                                   st.suse > 0);
      }
    } else {
      const bool we = st.use > 0;
      const uint32_t par = (st.use - 1) & 1;
      if (off) {
        if (we) wait_bar_relaxed(k, k.empty(st.slot), par);
      } else {
        produce_tile<UNITS, QUADS>(k, st.src, q, grp, lane_base + st.slot * kSlotCols, k.empty(st.slot), par, we);
      }
    }
    st.src += tile_vec4(KIND);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (KIND == kWO) {
      if (lane == 0) mbar_arrive(k.sfull(t));
    } else {
      if (lane == 0) mbar_arrive(k.full(st.slot));
      if (++st.slot == kSlots) {
        st.slot = 0;
        ++st.use;
      }
    }
  }
  if (KIND == kWO) ++st.suse;
}
__device__ __forceinline__ void producer_loop(const Ctx& k, int pw, int lane, uint32_t rank) {
  const int q = pw & 3, grp = pw >> 2;
  const uint32_t lane_base = k.tmem + (static_cast<uint32_t>(q * 32) << 16);
  const uint4* const src0 = k.p->blob + static_cast<size_t>(rank) * (kVirt * kRankVec4) + lane;
  const bool off = (k.p->debug_flags & 1) != 0;
  ProducerState st;
  for (int step = 0; step < k.p->n_steps; ++step) {
    st.src = src0;
#pragma unroll 1
    for (int layer = 0; layer < kLayers; ++layer) {   // the stream order of tile_info(): layers 5..8 start with the skip tile
      if (layer >= 5) produce_stage<kSK>(k, st, q, grp, lane, lane_base, off);
      produce_stage<kQKV>(k, st, q, grp, lane, lane_base, off);
      produce_stage<kWO>(k, st, q, grp, lane, lane_base, off);
      produce_stage<kW1>(k, st, q, grp, lane, lane_base, off);
      produce_stage<kW2>(k, st, q, grp, lane, lane_base, off);
    }
  }
}

// ---------------------------------------------------------------- MMA issuer (1 warp)
template <int K>
__device__ __forceinline__ void issue_tile(uint32_t a, uint32_t d, uint64_t bdesc, int jbase, bool acc0) {
#pragma unroll
  for (int kk = 0; kk < K / 16; ++kk) {
    const int j = jbase + kk;   // 16-wide k-step of the B buffer: box j / 4, 32 bytes per step inside the 128-B row
    const uint64_t bd = bdesc + static_cast<uint64_t>(((j >> 2) * 4096 + (j & 3) * 32) >> 4);
    umma_f16_ts(d, a + kk * 8, bd, kIdescN32, (acc0 || kk) ? 1u : 0u);     // W_hi . [x_hi ; x_lo']
    umma_f16_ts(d + 16, a + K / 2 + kk * 8, bd, kIdescN16, 1u);            // W_lo' . x_hi
  }
}
// One stage of the issuer, specialised on the tile kind: everything but the ring position is a compile-time constant,
// so the path from the "B operand ready" barrier to the first tcgen05.mma is a handful of uniform-datapath instructions
// (with the kind looked up per stage at run time it was ~25, on the critical path of all 41 stages).
struct IssuerState {
  uint32_t slot = 0, use = 0, suse = 0;
};
template <int KIND, bool PROF>
__device__ __forceinline__ void issue_stage(const Ctx& k, IssuerState& st, uint64_t bdesc, int step, int stage, bool prof_cta) {
  const Params& p = *k.p;
  constexpr bool nsplit = (KIND == kQKV) || (KIND == kW1);
  constexpr bool side = (KIND == kWO);
  constexpr int ksteps = (KIND == kWO) ? 2 : (KIND == kSK) ? 4 : 8;
  constexpr int K = tile_K(KIND);
  // the stage's two weight tiles (usually long in TMEM): observed here, while the epilogue warps are still busy
  uint32_t s2 = st.slot + 1, u2 = st.use;
  if (s2 == kSlots) {
    s2 = 0;
    ++u2;
  }
  const uint32_t a0 = k.tmem + (side ? kColSide : st.slot * kSlotCols), a1 = k.tmem + (side ? kColSide + 32 : s2 * kSlotCols);
  const uint32_t d0 = k.tmem + kColD, d1 = k.tmem + kColD + (nsplit ? 32 : 0);
  // profiled step: per stage, how long the weights kept the issuer (slots 300 + 2 stage) and how long it then waited
  // for the B operand (301 + 2 stage): a stage whose second number is ~0 was held up by its weights
  const bool wprof = PROF && prof_cta && step == p.prof_step;
  long long t_in = 0, t_full = 0;
  if (wprof) t_in = clock64();
  if (side) {
    wait_bar(k, k.sfull(0), st.suse & 1);
    wait_bar(k, k.sfull(1), st.suse & 1);
  } else {
    wait_bar(k, k.full(st.slot), st.use & 1);
    wait_bar(k, k.full(s2), u2 & 1);
  }
  if (wprof) t_full = clock64();
  asm volatile("bar.sync 4, 288;" ::: "memory");   // "B operand ready": all 8 epilogue warps have arrived (signal_b)
  if (wprof && elect_one()) {
    p.prof[300 + 2 * stage] = t_full - t_in;
    p.prof[301 + 2 * stage] = clock64() - t_full;
  }
  tc_fence_after();
  const bool fine = PROF && prof_cta && step == p.prof_step && (stage == 5 || stage == 6);
  if (fine && elect_one()) p.prof[stage == 5 ? 108 : 112] = clock64();
  if (elect_one()) {
    issue_tile<K>(a0, d0, bdesc, 0, false);
    if (nsplit) umma_commit(k.dbar(0));
    umma_commit(side ? k.sempty(0) : k.empty(st.slot));
    issue_tile<K>(a1, d1, bdesc, nsplit ? 0 : ksteps, !nsplit);
    umma_commit(nsplit ? k.dbar(1) : k.dbar(0));
    umma_commit(side ? k.sempty(1) : k.empty(s2));
  }
  __syncwarp();
  if (fine && elect_one()) p.prof[stage == 5 ? 109 : 113] = clock64();
  if (side) {
    ++st.suse;
  } else {
    st.slot = s2 + 1;
    st.use = u2;
    if (st.slot == kSlots) {
      st.slot = 0;
      ++st.use;
    }
  }
}
template <bool PROF>
__device__ __forceinline__ void issuer_loop(const Ctx& k) {
  const Params& p = *k.p;
  const bool prof_cta = PROF && (p.prof != nullptr) && cluster_id_x() == 0 && cluster_ctarank() == 0;
  const uint64_t dBx = umma_desc(smem_u32(k.smem + oBx)), dBo = umma_desc(smem_u32(k.smem + oBo)),
                 dBh = umma_desc(smem_u32(k.smem + oBh));
  IssuerState st;
  for (int step = 0; step < p.n_steps; ++step) {
    int stage = 0;   // the stage order of a step: layers 0..4 qkv|wo|w1|w2, layers 5..8 sk|qkv|wo|w1|w2 (tile_info)
#pragma unroll 1
    for (int layer = 0; layer < kLayers; ++layer) {
      if (layer >= 5) issue_stage<kSK, PROF>(k, st, dBx, step, stage++, prof_cta);
      issue_stage<kQKV, PROF>(k, st, dBx, step, stage++, prof_cta);
      issue_stage<kWO, PROF>(k, st, dBo, step, stage++, prof_cta);
      issue_stage<kW1, PROF>(k, st, dBx, step, stage++, prof_cta);
      issue_stage<kW2, PROF>(k, st, dBh, step, stage++, prof_cta);
    }
  }
}

// ---------------------------------------------------------------- the clip's epilogue warps (8 warps)
#define DN2_PROF(slot_)                                        \
  do {                                                         \
    if (do_prof) k.p->prof[(slot_)] = clock64();               \
  } while (0)
// stamps inside the stages of layer 1 (slots 100..127; scripts/quick_bench.py prints them)
#define DN2_FINE(slot_)                                        \
  do {                                                         \
    if (do_prof && layer == 1) k.p->prof[(slot_)] = clock64(); \
  } while (0)

struct Chain {
  const Ctx& k;
  const int q, t, lane, f;    // TMEM lane quadrant, group, lane, feature = TMEM lane
  const uint32_t rank;
  uint8_t* const smem;
  const uint32_t lane_taddr;  // TMEM address of my lane, column 0
  const uint32_t peer_ps, peer_ps2, peer_xbar;   // shared::cluster addresses in the peer CTA: my receive-slot entries (buffer 0), its xbar(0)
  uint32_t g_tile = 0;        // stage in the step
  uint32_t dph0 = 0, dph1 = 0, xe = 0;
  long long* wprof = nullptr; // debug: this warp's stamp row (lane 0 of every epilogue warp of cluster 0 / CTA 0), armed for
                              // the out_proj and FFN1 stages of layer 1 of the profiled step
  __device__ __forceinline__ void stamp(int point) const {
#ifdef AMUSE_DN2_FINE   // developer build: per-warp timeline of two stages (scripts/quick_bench.py prints it)
    if (wprof) wprof[point] = clock64();
#endif
  }

  __device__ Chain(const Ctx& k_, int q_, int t_, int lane_, uint32_t rank_)
      : k(k_), q(q_), t(t_), lane(lane_), f(q_ * 32 + lane_), rank(rank_), smem(k_.smem),
        lane_taddr(k_.tmem + (static_cast<uint32_t>(q_ * 32) << 16)),
        peer_ps(map_to_rank(k_.smem + oPs, rank_ ^ 1u) + t_ * 1024 + (q_ * 32 + lane_) * 8),
        peer_ps2(map_to_rank(k_.smem + oPs, rank_ ^ 1u) + 2048 + (q_ * 32 + lane_) * 4),
        peer_xbar(map_to_rank(k_.xbar(0), rank_ ^ 1u)) {}

  __device__ __forceinline__ void bar_group() const { bar_named(2 + t); }
  __device__ __forceinline__ void bar_all() const { asm volatile("bar.sync 1, 256;" ::: "memory"); }
  // bias | LN weight | LN bias of the stage about to run, for weight rank 2 rank + v
  __device__ __forceinline__ const float* vec(int v) const {
    return k.p->vecs + static_cast<size_t>(rank * kVirt + v) * kRankVecFloats + g_tile * 384 + f;
  }

  // B-operand element (row, input feature kcol) of buffer `buf`: x_hi at the computed offset, x_lo' 16 rows (2048 B) further
  template <int N>
  __device__ __forceinline__ void write_b(int buf, int kcol, int row_first, const float (&v)[N], int n) const {
    uint8_t* b0 = smem + buf + (kcol >> 6) * 4096 + (kcol & 7) * 2;
    const int cf = (kcol & 63) >> 3;
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i < n) {
        const int r = row_first + i;
        uint16_t h, l;
        split_h(v[i], h, l);
        uint8_t* d = b0 + r * 128 + ((cf ^ r) << 4);
        *reinterpret_cast<uint16_t*>(d) = h;
        *reinterpret_cast<uint16_t*>(d + 2048) = l;
      }
  }

  // Waiting for an mbarrier: every epilogue warp polls it itself.  (Mid-round, with the MMA issuer polling "B ready" on
  // the same schedulers, one poller per group + a named barrier for the other three was as fast and issued less; with
  // the issuer blocked on a hardware barrier instead, the extra hop costs 2.7 %: 56.5 -> 55.0 us/step, same box.)
  __device__ __forceinline__ void wait_group(uint64_t* bar, uint32_t parity) const { wait_bar(k, bar, parity); }
  // "my part of the B operand is written, and I am done with the accumulators": one arrive per warp
  __device__ __forceinline__ void signal_b() const {
    fence_proxy_async();   // my B-operand stores -> async proxy
    tc_fence_before();     // my tcgen05.ld of the previous accumulator -> before the MMAs that overwrite it
    // hardware named barrier 4 = 8 epilogue warps (arrive, non-blocking) + the issuer warp (sync): 2.3 % faster than an
    // mbarrier arrive per warp + the issuer's try_wait loop (the hop is on the critical path of all 41 stages, and a
    // blocked issuer issues nothing).  At most one phase is outstanding: no epilogue warp reaches the next signal_b before
    // the issuer has passed this one (its MMAs produce what they wait for).
    asm volatile("bar.arrive 4, 288;" ::: "memory");
  }
  // N-split stage: accumulator i (the tile of weight rank 2 rank + i) as soon as ITS MMAs have committed, my group's rows;
  // the epilogue of tile 0 runs under the MMAs of tile 1
  __device__ __forceinline__ void acc_nsplit(int i, float (&y)[kNR]) {
    wait_group(k.dbar(i), i ? dph1 : dph0);
    if (i) dph1 ^= 1;
    else dph0 ^= 1;
    tc_fence_after();
    float a[8], b[8];
    tmem_ld_2x8(lane_taddr + kColD + i * 32, a, b);
#pragma unroll
    for (int j = 0; j < kNR; ++j) y[j] = t ? fmaf(b[3 + j], 1.0f / 2048.0f, a[3 + j]) : fmaf(b[j], 1.0f / 2048.0f, a[j]);
  }
  // K-split stage: one accumulator (the sum over my two ranks), my group's rows
  __device__ __forceinline__ void gemm_ksplit(float (&y)[kNR]) {
    stamp(0);
    signal_b();
    wait_group(k.dbar(0), dph0);
    stamp(1);
    dph0 ^= 1;
    ++g_tile;
    tc_fence_after();
    float a[8], b[8];
    tmem_ld_2x8(lane_taddr + kColD, a, b);
#pragma unroll
    for (int i = 0; i < kNR; ++i) y[i] = t ? fmaf(b[3 + i], 1.0f / 2048.0f, a[3 + i]) : fmaf(b[i], 1.0f / 2048.0f, a[i]);
  }

  // ---- exchange of the K-split partial sums with the peer CTA
  __device__ __forceinline__ void xchg_arm() const {
    if (f == 0 && t == 0) mbar_arrive_expect_tx(k.xbar(xe & 1), kXchgBytes);
  }
  __device__ __forceinline__ void xchg_send(const float (&y)[kNR]) const {
    const uint32_t b = xe & 1u, off = b * kPsSlot, rb = peer_xbar + b * 8u;   // peer addresses: mapped once (constructor)
    st_async_v2(peer_ps + off, y[0], y[1], rb);
    if (t == 0) st_async_b32(peer_ps2 + off, y[2], rb);
  }
  __device__ __forceinline__ void xchg_recv(float (&v)[kNR]) {
    wait_group(k.xbar(xe & 1), (xe >> 1) & 1);
    const uint8_t* ps = smem + oPs + (xe & 1) * kPsSlot;
    const float2 a = *reinterpret_cast<const float2*>(ps + t * 1024 + f * 8);
    v[0] += a.x;
    v[1] += a.y;
    if (t == 0) v[2] += *reinterpret_cast<const float*>(ps + 2048 + f * 4);
    ++xe;
  }

  // ---- LayerNorm over the 128 features of my group's NR rows, feature f in this thread (nn.LayerNorm: biased
  // variance, eps 1e-5).  Per warp: shift by the warp's first feature, sum d and d^2 over the 32 lanes (transposing
  // butterfly: the xor-16 step hands the sums to the lower half-warp and the sums of squares to the upper one); lane r
  // then merges the 4 warps' (mean, M2) of row r with Chan's formula -- no cancellation whatever the row mean is --
  // and (mean, rstd) are broadcast to the warp.
  // The moments go through one of three shared-memory buffers: LayerNorm 1 and 2 of a layer alternate between two (a
  // warp may write the moments of the next LayerNorm while a slower warp of its group still reads the previous ones only
  // if nothing separates the two -- every pair of consecutive LayerNorms of a layer is separated by a GEMM stage, i.e.
  // by the B-ready barrier / accumulator mbarriers, but the final encoder.norm follows LayerNorm 2 of the last layer directly:
  // compute-sanitizer racecheck flagged exactly that write-after-read), the final norm has its own.
  template <int NR>
  __device__ __forceinline__ void layernorm(float (&v)[kNR], float gam, float bet, int buf) const {
    float tt[NR], sh[NR];
    const bool upper = (lane & 16) != 0;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      sh[r] = __shfl_sync(0xffffffffu, v[r], 0);
      const float d = v[r] - sh[r];
      const float s1 = d, s2 = d * d;
      const float keep = upper ? s2 : s1, send = upper ? s1 : s2;
      tt[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < NR; ++r) tt[r] += __shfl_xor_sync(0xffffffffu, tt[r], o);
    float* st = reinterpret_cast<float*>(smem + oStat) + buf * 128 + (t * 4 + q) * 16;
    if ((lane & 15) == 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) st[(upper ? 4 : 0) + r] = tt[r];
    }
    if (lane == 1) {
#pragma unroll
      for (int r = 0; r < NR; ++r) st[8 + r] = sh[r];
    }
    stamp(5);
    bar_group();
    stamp(6);
    const float* sa = reinterpret_cast<const float*>(smem + oStat) + buf * 128 + t * 64 + ((lane < NR) ? lane : 0);
    float mw[4], m2 = 0.f, mean = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float s1 = sa[w * 16], s2 = sa[w * 16 + 4];
      mw[w] = fmaf(s1, 1.0f / 32.0f, sa[w * 16 + 8]);
      m2 += fmaf(-s1 * (1.0f / 32.0f), s1, s2);
      mean += mw[w];
    }
    mean *= 0.25f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float d = mw[w] - mean;
      m2 = fmaf(32.0f * d, d, m2);
    }
    const float rstd = rsqrtf(fmaxf(m2 * (1.0f / 128.0f), 0.f) + kLnEps);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float mr = __shfl_sync(0xffffffffu, mean, r), rr = __shfl_sync(0xffffffffu, rstd, r);
      v[r] = (v[r] - mr) * rr * gam + bet;
    }
  }

  // ---- attention of head 2 rank + t for the T tokens of the clip (cross_attention.py:264-266; nn.MultiheadAttention, 4
  // heads of 32).  q (pre-scaled) | k | v rows are in shared memory; warps q = 0, 1 of the group: 3 query rows per warp,
  // 10 lanes per row = 5 keys x 2 halves of the head dimension.  The output goes straight into the B operand of out_proj.
  __device__ __forceinline__ void attention(int T) const {
    if (q >= 2) return;
    const float* QKVs = reinterpret_cast<const float*>(smem + oQKV) + t * (kRows * kQkvLd);
    const int at_slot = (lane < 30) ? lane / 10 : 0;
    const int at_l = lane - (lane / 10) * 10;                   // position inside the slot: j * 2 + half
    const bool at_live = (lane < 30) && (q * 3 + at_slot < T);
    const int at_row = at_live ? q * 3 + at_slot : 0;
    const int at_c = at_l & 1;
    const bool at_key = at_live && (at_l >> 1) < T;
    const int at_j = at_key ? (at_l >> 1) : 0;
    const int at_src = at_slot * 10;
    const bool at_pv = at_live && at_l < 8;
    const float4* qv = reinterpret_cast<const float4*>(QKVs + at_row * kQkvLd + 16 * at_c);
    const float4* kv = reinterpret_cast<const float4*>(QKVs + at_j * kQkvLd + 32 + 16 * at_c);
    // the value rows do not depend on the scores: fetched up front, under the score chain (8 lanes per row: 4 head
    // dimensions each)
    float4 v4[5];
    {
      const float* vb = QKVs + 64 + 4 * (at_l & 7);
#pragma unroll
      for (int jj = 0; jj < 5; ++jj) v4[jj] = *reinterpret_cast<const float4*>(vb + ((jj < T) ? jj : 0) * kQkvLd);
    }
    float p0, p1, p2, p3;
    {
      const float4 a = qv[0], b = kv[0];
      p0 = a.x * b.x;
      p1 = a.y * b.y;
      p2 = a.z * b.z;
      p3 = a.w * b.w;
    }
#pragma unroll
    for (int i = 1; i < 4; ++i) {
      const float4 a = qv[i], b = kv[i];
      p0 = fmaf(a.x, b.x, p0);
      p1 = fmaf(a.y, b.y, p1);
      p2 = fmaf(a.z, b.z, p2);
      p3 = fmaf(a.w, b.w, p3);
    }
    float sc = (p0 + p1) + (p2 + p3);
    sc += __shfl_xor_sync(0xffffffffu, sc, 1);   // the two halves of the head dimension
    float sj[5];
#pragma unroll
    for (int jj = 0; jj < 5; ++jj) sj[jj] = __shfl_sync(0xffffffffu, sc, at_src + 2 * jj);
    float m = sj[0];
#pragma unroll
    for (int jj = 1; jj < 5; ++jj) m = (jj < T) ? fmaxf(m, sj[jj]) : m;
    // every lane of the row exponentiates the 5 scores itself (q carries log2 e, so exp is one ex2.approx): a second
    // gather of 5 shuffles would sit on the dependency chain
    float ej[5];
#pragma unroll
    for (int jj = 0; jj < 5; ++jj) ej[jj] = (jj < T) ? ex2_approx(sj[jj] - m) : 0.f;
    const float sum = ((ej[0] + ej[1]) + (ej[2] + ej[3])) + ej[4];
    const float inv = __frcp_rn(sum);
    if (at_pv) {   // 8 lanes per row: 4 head dimensions each
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int jj = 0; jj < 5; ++jj)
        if (jj < T) {
          acc.x = fmaf(ej[jj], v4[jj].x, acc.x);
          acc.y = fmaf(ej[jj], v4[jj].y, acc.y);
          acc.z = fmaf(ej[jj], v4[jj].z, acc.z);
          acc.w = fmaf(ej[jj], v4[jj].w, acc.w);
        }
      const float o[4] = {acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv};
      uint16_t h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_h(o[i], h[i], l[i]);
      // element (row, k = 32 t + 4 at_l .. +3): 16-B chunk (4 t + (at_l >> 1)) ^ row, 8 bytes at (at_l & 1) * 8
      uint8_t* d = smem + oBo + at_row * 128 + ((((4 * t + (at_l >> 1)) ^ at_row) & 7) << 4) + (at_l & 1) * 8;
      *reinterpret_cast<uint2*>(d) = make_uint2(h[0] | (static_cast<uint32_t>(h[1]) << 16), h[2] | (static_cast<uint32_t>(h[3]) << 16));
      *reinterpret_cast<uint2*>(d + 2048) = make_uint2(l[0] | (static_cast<uint32_t>(l[1]) << 16), l[2] | (static_cast<uint32_t>(l[3]) << 16));
    }
  }

  template <bool PROF>
  __device__ void run();
};

// PROF = the kernel instantiation with the in-kernel clock stamps (amuse_profile_arm); the product launch compiles them out
template <bool PROF>
__device__ void Chain::run() {
  const Params& p = *k.p;
  const int T = p.T;
  const int clip = static_cast<int>(cluster_id_x());
  const bool do_prof_chain = PROF && (p.prof != nullptr) && clip == 0 && rank == 0 && t == 0 && f == 0;
  float* const SK = reinterpret_cast<float*>(smem + oSK);
  const int row0 = t * kNR, nr = t ? 2 : 3;   // my rows in the row-split epilogues

  // per-thread constants: feature f of the PE rows, the condition tokens and the final norm
  const float pe0 = p.pe01[f], pe1 = p.pe01[128 + f];
  float ct[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) ct[j] = (j < T - 2) ? p.cond[(static_cast<size_t>(clip) * 3 + j) * 128 + f] : 0.f;
  const float fn_g = p.final_norm[f], fn_b = p.final_norm[128 + f];
  float z = p.latents0[static_cast<size_t>(clip) * 128 + f];   // group 0 owns the latent (replicated in the two CTAs)
  const bool use_rng = (p.step_noise == nullptr);
  const unsigned long long elem = p.seed_elem_base + static_cast<unsigned long long>(clip) * 128ull + f;

  // software prefetch (one step ahead) of the per-step global reads
  float temb_next = __ldg(p.temb + f);
  float coef_next[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) coef_next[i] = __ldg(p.coef + i);
  float noise_next = use_rng ? 0.f : __ldg(p.step_noise + static_cast<size_t>(clip) * 128 + f);

  float x[kNR];   // residual stream: feature f of my group's token rows (replicated in the two CTAs)

  for (int step = 0; step < p.n_steps; ++step) {
    const bool do_prof = do_prof_chain && step == p.prof_step;
    long long* const fine_warp = (p.prof != nullptr && clip == 0 && rank == 0 && lane == 0 && step == p.prof_step)
                                     ? p.prof + 128 + (t * 4 + q) * 16
                                     : nullptr;
    DN2_PROF(0);
    float coef[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) coef[i] = coef_next[i];
    const float noise = noise_next;
    const float temb = temb_next;
    if (step + 1 < p.n_steps) {
      temb_next = __ldg(p.temb + static_cast<size_t>(step + 1) * 128 + f);
#pragma unroll
      for (int i = 0; i < 5; ++i) coef_next[i] = __ldg(p.coef + static_cast<size_t>(step + 1) * 5 + i);
      if (!use_rng) noise_next = __ldg(p.step_noise + (static_cast<size_t>(step + 1) * p.B + clip) * 128 + f);
    }
    g_tile = 0;
    // ---- token rows (denoiser.py:174-181 + position_encoding.py:156): z, t, con | emo, sty
    x[0] = t ? ct[1] : z + pe0;
    x[1] = t ? ct[2] : temb + pe1;
    x[2] = t ? 0.f : ct[0];
    write_b(oBx, f, row0, x, nr);
    DN2_PROF(1);
    float zn = noise;   // this step's noise draw (in-kernel Philox: evaluated under the first MMA phase, see the QKV stage)

    for (int layer = 0; layer < kLayers; ++layer) {
      // =============== output blocks: x = Linear(256->128)(cat(x, xs.pop())), K-split: CTA 0 holds x, CTA 1 the skip
      if (layer >= 5) {
        float y[kNR];
        const float bias = __ldg(vec(0));
        xchg_arm();
        gemm_ksplit(y);
        xchg_send(y);
#pragma unroll
        for (int i = 0; i < kNR; ++i) y[i] += bias;
        xchg_recv(y);
#pragma unroll
        for (int i = 0; i < kNR; ++i) x[i] = y[i];
        write_b(oBx, f, row0, x, nr);
      }
      DN2_PROF(2 + layer * 10 + 0);

      // =============== QKV of head 2 rank + t (nn.MultiheadAttention in_proj; q scaled by head_dim^-0.5 after the bias)
      {
        float y[kNR];
        const float b0 = __ldg(vec(0)), b1 = __ldg(vec(1));
        const float sc = (q == 0) ? 0.17677669529663687f * 1.4426950408889634f : 1.0f;   // head_dim^-0.5, and log2 e for the softmax
        float* const Q0 = reinterpret_cast<float*>(smem + oQKV) + row0 * kQkvLd + f;
        signal_b();
        // the noise draw does not depend on the chain: off the critical path, while the first MMAs of the step run
        if (layer == 0 && t == 0 && use_rng && coef[4] != 0.f) zn = philox_normal(p.seed, elem, static_cast<uint32_t>(step));
#pragma unroll
        for (int i = 0; i < kVirt; ++i) {
          acc_nsplit(i, y);
          if (f < 96) {
#pragma unroll
            for (int j = 0; j < kNR; ++j)
              if (j < nr) Q0[(i * kRows + j) * kQkvLd] = (y[j] + (i ? b1 : b0)) * sc;
          }
        }
        ++g_tile;
        bar_all();
      }
      DN2_PROF(2 + layer * 10 + 1);
      attention(T);
      DN2_PROF(2 + layer * 10 + 2);

      // =============== out_proj, K-split by head -> exchange -> + bias + residual -> LayerNorm 1
      {
        float y[kNR];
        const float bias = __ldg(vec(0)), gam = __ldg(vec(0) + 128), bet = __ldg(vec(0) + 256);
        xchg_arm();
        if (fine_warp && layer == 1) wprof = fine_warp;
        gemm_ksplit(y);
        stamp(2);
        xchg_send(y);
        stamp(3);
#pragma unroll
        for (int i = 0; i < kNR; ++i) y[i] += bias + x[i];
        xchg_recv(y);
        stamp(4);
        DN2_PROF(2 + layer * 10 + 3);
        layernorm<kNR>(y, gam, bet, 0);
        asm volatile("" ::"f"(y[0]), "f"(y[1]), "f"(y[2]));
        stamp(7);
#pragma unroll
        for (int i = 0; i < kNR; ++i) x[i] = y[i];
        write_b(oBx, f, row0, x, nr);
        stamp(8);
      }
      DN2_PROF(2 + layer * 10 + 4);

      // =============== FFN1: hidden slice 2 rank + t, erf-GELU
      {
        float y[kNR];
        const float b0 = __ldg(vec(0)), b1 = __ldg(vec(1));
        signal_b();
#pragma unroll
        for (int i = 0; i < kVirt; ++i) {
          acc_nsplit(i, y);
#pragma unroll
          for (int j = 0; j < kNR; ++j) y[j] = gelu_erf(y[j] + (i ? b1 : b0));
          write_b(oBh, i * 128 + f, row0, y, nr);
        }
        ++g_tile;
        wprof = nullptr;
      }
      DN2_PROF(2 + layer * 10 + 5);

      // =============== FFN2, K-split over my 256 hidden units -> exchange -> + bias + residual -> LayerNorm 2
      {
        float y[kNR];
        const float bias = __ldg(vec(0)), gam = __ldg(vec(0) + 128), bet = __ldg(vec(0) + 256);
        xchg_arm();
        gemm_ksplit(y);
        xchg_send(y);
#pragma unroll
        for (int i = 0; i < kNR; ++i) y[i] += bias + x[i];
        xchg_recv(y);
        DN2_PROF(2 + layer * 10 + 6);
        layernorm<kNR>(y, gam, bet, 1);
#pragma unroll
        for (int i = 0; i < kNR; ++i) x[i] = y[i];
        if (layer < 4) {
#pragma unroll
          for (int i = 0; i < kNR; ++i)
            if (i < nr) SK[(layer * kRows + row0 + i) * 128 + f] = x[i];
        }
        if (layer + 1 < kLayers) {
          if (layer + 1 >= 5 && rank == 1) {   // next stage = skip-linear: CTA 1's K slice of cat(x, skip) is the skip of layer 7 - layer
            float sv[kNR];
#pragma unroll
            for (int i = 0; i < kNR; ++i) sv[i] = (i < nr) ? SK[((7 - layer) * kRows + row0 + i) * 128 + f] : 0.f;
            write_b(oBx, f, row0, sv, nr);
          } else {
            write_b(oBx, f, row0, x, nr);
          }
        }
      }
      DN2_PROF(2 + layer * 10 + 7);
    }   // layers

    // ---- encoder.norm on token 0 -> eps (cross_attention.py:62-63, denoiser.py:188), then the scheduler step (K2),
    //      replicated in both CTAs; op order of diffusers' step(): x0 = (x - sqrt(1-a) e) / sqrt(a); clamp;
    //      x' = c2 x0 + c3 (e | x) + sigma z
    if (t == 0) {
      float e3[kNR] = {x[0], 0.f, 0.f};
      layernorm<1>(e3, fn_g, fn_b, 2);
      const float e = e3[0];
      float x0 = __fdiv_rn(__fsub_rn(z, __fmul_rn(coef[1], e)), coef[0]);
      if (p.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
      float out = __fadd_rn(__fmul_rn(coef[2], x0), __fmul_rn(coef[3], p.dir_uses_eps ? e : z));
      if (coef[4] != 0.f) out = __fadd_rn(out, __fmul_rn(coef[4], zn));
      z = out;
    }
    DN2_PROF(2 + kLayers * 10);
  }   // steps
  if (rank == 0 && t == 0) p.latents_out[static_cast<size_t>(clip) * 128 + f] = z;
}

}  // namespace

// ================================================================= the kernel
template <bool PROF>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    denoise_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();

  Ctx k;
  k.p = &p;
  k.smem = smem;
  k.bars = reinterpret_cast<uint64_t*>(smem + oBars);
  k.status = p.status;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + oBars + 128);

  for (int i = tid; i < kSmemUsed / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  if (tid == 0) {
    for (uint32_t s = 0; s < kSlots; ++s) {
      mbar_init(k.full(s), kProdWarps);
      mbar_init(k.empty(s), 1);
    }
    mbar_init(k.dbar(0), 1);
    mbar_init(k.dbar(1), 1);
    mbar_init(k.xbar(0), 1);
    mbar_init(k.xbar(1), 1);
    for (int i = 0; i < kVirt; ++i) {
      mbar_init(k.sfull(i), kProdWarps);
      mbar_init(k.sempty(i), 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  k.tmem = *tmem_slot;
  cluster_sync_all();   // both CTAs are resident, zero-filled and have their mbarriers initialised before either
                        // stores into the other's shared memory

  // warp ids: producers 0..7, epilogue warps 8..15, issuer 16 -- the warp scheduler prefers the highest id among the
  // eligible warps, and the producers are the one role that is never on the critical path
  if (warp < kProdWarps) {
    producer_loop(k, warp, lane, rank);
  } else if (warp < kProdWarps + kChainWarps) {
    Chain ch(k, warp & 3, (warp - kProdWarps) >> 2, lane, rank);
    ch.template run<PROF>();
  } else {
    issuer_loop<PROF>(k);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(k.tmem);
  }
  cluster_sync_all();   // nobody leaves while the peer could still address its shared memory
}

size_t smem_bytes() { return static_cast<size_t>(kSmemBytes); }

cudaError_t launch(const Params& p, cudaStream_t stream) {
  static std::mutex mu;
  static bool configured_dev[64] = {};   // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!configured_dev[dev & 63]) {
      cudaError_t e = cudaFuncSetAttribute(denoise_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(denoise_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      if (e != cudaSuccess) return e;
      configured_dev[dev & 63] = true;
    }
  }
  if (p.B < 1 || p.T < 2 || p.T > kTMax || p.n_steps < 1) return cudaErrorInvalidValue;
  if (p.prof) denoise_tc_kernel<true><<<dim3(p.B * kCluster), dim3(kThreads), kSmemBytes, stream>>>(p);
  else denoise_tc_kernel<false><<<dim3(p.B * kCluster), dim3(kThreads), kSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

void split_fp16(float x, uint16_t& hi, uint16_t& lo) {
  auto sat = [](float v) { return v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v); };
  const __half h = __float2half_rn(sat(x));
  const float r = (x - __half2float(h)) * 2048.0f;
  const __half l = __float2half_rn(sat(r));
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
}

}  // namespace dn2
}  // namespace amuse
