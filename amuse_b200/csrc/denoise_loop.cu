// K1+K2: the whole N-step denoising loop of PretrainedLPDM_v1.diffusion_backward
// (reference infer_ldm.py:142-161) as ONE persistent launch.
//
// What one step computes (reference denoiser.py:135-204, cross_attention.py:41-64,259-272):
//   x = [z | time token | con | emo | sty] + learned PE        (<= 5 tokens x 128 per clip)
//   9 post-LN encoder layers with U-Net skips, final LayerNorm, eps = token 0
//   scheduler update of z (DDIM eta / DDPM ancestral), clamp(x0) optional
//
// B200 mapping.  The loop is a chain of ~40 dependent small GEMMs per step with M = 5 rows per
// clip: it is dependency-latency bound, and every evaluation needs all 8.8 MB of fp32 weights.
// A thread-block cluster of 4 CTAs (4 SMs, one attention head each) owns up to 2 clips for all
// steps; 33 such clusters are co-resident on a B200, so a 64-clip batch runs as one wave.
//   * Weights are split 4 ways across the cluster, so each SM streams 1/4 of them per step from
//     L2 through the TMA bulk-copy engine into a 2-deep shared-memory ring that runs two tiles
//     ahead of the math (mbarrier complete_tx signalling).
//   * Activations (10 x 128 floats) are replicated in every CTA's shared memory.
//       QKV      N-split: CTA c computes q|k|v of head c for all rows             -> local
//       attn     T x T per clip, one warp per clip                                 -> local
//       out_proj K-split by head: every CTA produces a full-width partial sum and stores it into
//                the 3 peers' shared memory with st.async (DSMEM); the bytes are credited to the
//                receiver's mbarrier, so there is NO cluster barrier: each CTA waits on its own
//                mbarrier, adds the 4 partials + bias + residual and applies LayerNorm 1 (replicated)
//       FFN1     N-split: 128 hidden units per CTA, erf-GELU                        -> local
//       FFN2     K-split over the same 128 units -> st.async partial broadcast -> sum + LN2
//       skip     Linear(256->128)(cat(x, skip)) K-split 64 per CTA, same exchange, no LN
//     => 22 point-to-point exchanges per step, zero cluster barriers inside the loop (a
//        barrier.cluster.arrive.release costs a MEMBAR.ALL.GPU on sm_100 -- measured ~1.5k cycles),
//        and no global-memory traffic for activations.
//   * All arithmetic is fp32 FFMA: 50..1000 recurrent steps with clamp() do not survive bf16
//     (SURVEY.md App. C), and at M <= 10 rows the tensor pipe would be operand-bandwidth bound.
#include "denoise_loop.cuh"

#include <mutex>

#include "common.cuh"
#include "philox.cuh"

namespace amuse {
namespace dn {

namespace {

constexpr int kWBufFloats = 16768;   // == kTileW2 (largest tile), multiple of 32 floats
static_assert(kWBufFloats >= kTileMax && kWBufFloats % 32 == 0, "weight ring slot too small");
constexpr int kQkvLd = 100;          // q|k|v row stride: 16-B aligned rows, conflict-free T x T dot products
constexpr int kOhLd = 36;            // attention-output row stride (16-B aligned rows)
constexpr int kRedFloats = 8 * 5 * 128;      // K-split partial sums: KSPLIT x rows x 128  (= 4 x 10 x 128)
constexpr int kPsFloats = 3 * kRMax * 128;   // partial sums received from the 3 peer CTAs
constexpr int kQkvOhFloats = kRMax * kQkvLd + kRMax * kOhLd;
static_assert(kQkvOhFloats >= kRMax * 128, "Hs aliases the q|k|v + attention-output region");

// ---------------------------------------------------------------- shared-memory carve-up
// Offsets (floats) from the dynamic shared-memory base.  Every pointer is formed as
// `smem + constant`, so the compiler keeps the .shared address space (LDS/STS, not generic LD/ST).
constexpr int kRedSlot = kRMax * 128;        // one K-slice's partial tile: 10 rows x 128 (WIDE layout, see the kernel)
static_assert(kRedFloats == 4 * kRedSlot && kPsFloats == 3 * kRedSlot, "WIDE slot map: 4 in RED, 3 in the idle Ps buffer, 1 extra");
constexpr int kActFloats = kRMax * 128 + 4 * kRMax * 128 + 2 * kPsFloats + kQkvOhFloats + kRedFloats + kRedSlot +
                           kSMax * 3 * 128 + 256 + kSMax * 128 + kSMax * 128 + 128 + 256 + 96 + kTileTail;
constexpr int kSmemFloats = 2 * kWBufFloats + kActFloats + 16 /*mbarriers + pad*/;
static_assert(kSmemFloats * 4 <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");

struct Smem {
  float* base;
  __device__ __forceinline__ float* wbuf(uint32_t i) const { return base + i * kWBufFloats; }   // weight ring
  __device__ __forceinline__ float* at(int off) const { return base + 2 * kWBufFloats + off; }
  static constexpr int oXs = 0;                              // [10][128] residual stream (replicated)
  static constexpr int oSK = oXs + kRMax * 128;              // [4][10][128] skip stack of the input blocks
  static constexpr int oPs = oSK + 4 * kRMax * 128;          // [2][3][10][128] peer partials (remote-written)
  static constexpr int oQKV = oPs + 2 * kPsFloats;           // [10][100] q|k|v of my head
  static constexpr int oOh = oQKV + kRMax * kQkvLd;          // [10][36] attention output of my head
  static constexpr int oHs = oQKV;                           // [10][128] my 128 hidden units -- ALIASES q|k|v/Oh
                                                             // (dead between out_proj and the next layer's QKV)
  static constexpr int oRED = oQKV + kQkvOhFloats;           // K-split partial sums
  static constexpr int oRed7 = oRED + kRedFloats;            // [10][128] eighth K-slice of the WIDE layout
  static constexpr int oCs = oRed7 + kRedSlot;               // [2][3][128] condition tokens (+PE)
  static constexpr int oPe = oCs + kSMax * 3 * 128;          // [2][128]
  static constexpr int oZs = oPe + 256;                      // [2][128] current latents
  static constexpr int oEs = oZs + kSMax * 128;              // [2][128] eps
  static constexpr int oTemb = oEs + kSMax * 128;            // [128]
  static constexpr int oFn = oTemb + 128;                    // [256] final norm
  static constexpr int oBqkv = oFn + 256;                    // [96]  q|k|v bias of the current layer
  static constexpr int oTail = oBqkv + 96;                   // [384] bias | LN weight | LN bias of the current stage
  static constexpr int oBar = oTail + kTileTail;             // 4 mbarriers: 2 weight ring, 2 exchange
  static_assert(oBar == kActFloats, "carve-up does not match kActFloats");
  __device__ __forceinline__ uint64_t* wbar(uint32_t i) const { return reinterpret_cast<uint64_t*>(at(oBar)) + i; }
  __device__ __forceinline__ uint64_t* xbar(uint32_t i) const { return reinterpret_cast<uint64_t*>(at(oBar)) + 2 + i; }
  __device__ __forceinline__ float* Ps(uint32_t i) const { return at(oPs) + i * kPsFloats; }
};

// ---------------------------------------------------------------- weight pipeline
struct WPipe {
  const float* blob;   // this rank's stream
  uint32_t g;          // sequence number of the tile about to be consumed
  uint32_t total;      // n_steps * 40
};

__constant__ int2 c_tile_tab[kTilesPerStep];   // (offset, count) of tile i -- tile_info() evaluated once on the host

__device__ __forceinline__ void wp_issue(const Smem& s, const WPipe& w, uint32_t n) {
  const int2 tab = c_tile_tab[n % kTilesPerStep];
  const int off = tab.x, cnt = tab.y;
  uint64_t* bar = s.wbar(n & 1);
  mbar_arrive_expect_tx(bar, static_cast<uint32_t>(cnt) * 4u);
  bulk_g2s(s.wbuf(n & 1), w.blob + off, static_cast<uint32_t>(cnt) * 4u, bar);
}
__device__ __forceinline__ const float* wp_acquire(const Smem& s, const WPipe& w) {
  mbar_wait(s.wbar(w.g & 1), (w.g >> 1) & 1);
  return s.wbuf(w.g & 1);
}
// Pre-waited tiles.  An mbarrier try_wait costs ~100 cycles even when the tile landed long ago (measured: 117..278
// cycles between the end of a stage and the first GEMM instruction of the next), and every GEMM stage used to start
// with one in all 10 warps.  Warp 9 (no GEMM role; its lane 0 issues the TMA) instead observes the completed phase
// at a point where it would idle anyway, always before a __syncthreads() that precedes the stage, and the stage
// starts without touching the mbarrier: the barrier orders warp 9's observation before every other warp's reads of
// the tile.  Two placements, selected at compile time (both pass the full GPU suite; measured on one box, B = 64,
// against 62.5 us/step with the all-thread wait):
//   default                  while warps 0..7 run the GEMM of tile g, warp 9 waits for tile g+1 (issued a whole
//                            stage earlier): covers every tile                                       61.3 us/step
//   AMUSE_PREWAIT_IDLE_SITES under the exchange wait of the previous stage / during the attention stage; the FFN2
//                            tile, which follows a stage without idle time, keeps the all-thread wait 62.0 us/step
#ifdef AMUSE_PREWAIT_IDLE_SITES
constexpr bool kPrewaitInGemm = false;
#else
constexpr bool kPrewaitInGemm = true;
#endif
__device__ __forceinline__ void wp_prewait(const Smem& s, const WPipe& w, uint32_t ahead) {
  const uint32_t n = w.g + ahead;
  if (n < w.total) mbar_wait(s.wbar(n & 1), (n >> 1) & 1);
}
__device__ __forceinline__ const float* wp_acquire_prewaited(const Smem& s, const WPipe& w) { return s.wbuf(w.g & 1); }
// the FFN2 tile: pre-waited only by the in-GEMM placement
__device__ __forceinline__ const float* wp_acquire_ffn2(const Smem& s, const WPipe& w) {
  return kPrewaitInGemm ? wp_acquire_prewaited(s, w) : wp_acquire(s, w);
}
constexpr int kIssuerWarp = 9;
// Call after a __syncthreads() that follows the last read of tile g: hands the buffer back
// to the TMA engine for tile g+2.  Issued by one lane of warp 9, which has no GEMM role (measured
// earlier: with thread 0 issuing, warp 0's epilogue was 700 cycles late in every stage).  No proxy
// fence: the buffer was only READ by this CTA and every reader has passed the barrier.
constexpr int kIssuerTid = 9 * 32;
// (Looking the table entry of tile g+2 up during the GEMM, so that the issuing lane only arms the mbarrier and
// fires the copy after the barrier, was measured slower: 61.9 vs 60.6 us/step on the same box.)
__device__ __forceinline__ void wp_release(const Smem& s, WPipe& w, int tid) {
  if (tid == kIssuerTid && w.g + 2 < w.total) wp_issue(s, w, w.g + 2);
  ++w.g;
}

// ---------------------------------------------------------------- 5-row FFMA2 micro-kernel
// acc[i][j] = sum_{k in my K slice} A[rb*5+i][k] * W[k][col(lane,j)]
//
// Measured on B200 (scripts/ubench.cu): a 3-register FFMA issues every 1.44 cycles per SM
// sub-partition and the packed FFMA2 (two fp32 FMAs on 64-bit register pairs) every 2.75, i.e. both
// top out near 90 FMA/clk/SM, but FFMA2 halves the issue slots; LDS.64 sustains 125 B/clk/SM while
// LDS.128 only 64 B/clk, and a warp-uniform LDS.128 costs 2.2 cycles.  Hence:
//   * the two halves of an FFMA2 are two consecutive k: weight tiles are stored k-pair interleaved,
//     Wt2[k/2][p][2], so one 64-bit word holds (W[k][c], W[k+1][c]); an A-row float4 supplies
//     (A[k],A[k+1]) and (A[k+2],A[k+3]).  Even- and odd-k products accumulate in the two halves;
//   * weights are read with LDS.64 only: lane reads tile positions p = lane + 32*j.  For the
//     128-wide tiles the host packs the tile so that lane owns the output columns
//     {2*lane, 2*lane+1, 64+2*lane, 64+2*lane+1} (j = 0..3): every epilogue access (K-split park /
//     gather, bias, residual, LayerNorm params, DSMEM stores) is then a conflict-free 64-bit access
//     too;  for the q|k|v tile position p is the natural (q,k,v)[lane] triple;
//   * A-row loads are warp-uniform 128-bit broadcasts.
// ITERS 4-k iterations starting at a0 = A + (row block, my K slice) / w = Wt2 + my K slice.  Fully
// unrolled: with a runtime trip count (or unroll 2) the loads are not hoisted far enough ahead of
// the FFMA2s and every GEMM stage got 15-25% slower (measured), which outweighs the extra
// instruction-cache pressure of the larger body.
// NR = rows per warp: 5 (one clip's tokens) everywhere, except in the pruned last layer (1 or 2 rows).
template <int NR, int NCOL, int TC, int ITERS>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ a0, int lda, const float* __restrict__ w,
                                          float (&acc)[NR][TC]) {
  static_assert(NCOL == 32 * TC, "lane owns tile positions lane + 32*j");
  float2 acc2[NR][TC];
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc2[i][j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    float4 a[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * lda + it * 4);
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
      float2 wv[TC];
      const float* wr = w + (it * 2 + kp) * (NCOL * 2);
#pragma unroll
      for (int j = 0; j < TC; ++j) wv[j] = *reinterpret_cast<const float2*>(wr + 64 * j);
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const float2 av = (kp == 0) ? make_float2(a[i].x, a[i].y) : make_float2(a[i].z, a[i].w);
#pragma unroll
        for (int j = 0; j < TC; ++j) acc2[i][j] = __ffma2_rn(av, wv[j], acc2[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[i][j] = acc2[i][j].x + acc2[i][j].y;
}
template <int NCOL, int TC, int ITERS>
__device__ __forceinline__ void gemm5(const float* __restrict__ a0, int lda, const float* __restrict__ w,
                                      float (&acc)[5][TC]) {
  gemm_rows<5, NCOL, TC, ITERS>(a0, lda, w, acc);
}

// K-split reduction through shared memory.  Every warp parks its partial tile in
// RED[ks][row][NCOL]; after a __syncthreads() the row-owner warps (warp w owns rows w and w+8)
// add the KSPLIT partials of their rows, so the epilogues run on all 8 warps.
// A lane's 4 columns as two float2: lo = columns (2*lane, 2*lane+1), hi = (64+2*lane, 64+2*lane+1).
struct Row4 {
  float2 lo, hi;
};
__device__ __forceinline__ Row4 ld_row4(const float* row, int lane) {
  Row4 r;
  r.lo = *reinterpret_cast<const float2*>(row + 2 * lane);
  r.hi = *reinterpret_cast<const float2*>(row + 64 + 2 * lane);
  return r;
}
__device__ __forceinline__ void st_row4(float* row, int lane, const Row4& v) {
  *reinterpret_cast<float2*>(row + 2 * lane) = v.lo;
  *reinterpret_cast<float2*>(row + 64 + 2 * lane) = v.hi;
}
__device__ __forceinline__ Row4 add4(const Row4& a, const Row4& b) {
  Row4 r;
  r.lo = make_float2(a.lo.x + b.lo.x, a.lo.y + b.lo.y);
  r.hi = make_float2(a.hi.x + b.hi.x, a.hi.y + b.hi.y);
  return r;
}
template <int RT>
__device__ __forceinline__ void park4(float* RED, int rb, int ks, int lane, const float (&acc)[5][4]) {
  float* dst = RED + ((ks * RT + rb * 5) * 128);
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    Row4 v;
    v.lo = make_float2(acc[i][0], acc[i][1]);
    v.hi = make_float2(acc[i][2], acc[i][3]);
    st_row4(dst + i * 128, lane, v);
  }
}
// pruned last layer: NR rows per warp, all 8 warps split K; RED[ks][NR][128]
template <int NR>
__device__ __forceinline__ void park_rows(float* RED, int ks, int lane, const float (&acc)[NR][4]) {
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    Row4 v;
    v.lo = make_float2(acc[i][0], acc[i][1]);
    v.hi = make_float2(acc[i][2], acc[i][3]);
    st_row4(RED + (ks * NR + i) * 128, lane, v);
  }
}
template <int RT, int KS>
__device__ __forceinline__ Row4 gather4(const float* RED, int row, int lane) {
  Row4 v = ld_row4(RED + row * 128, lane);
#pragma unroll
  for (int q = 1; q < KS; ++q) v = add4(v, ld_row4(RED + (q * RT + row) * 128, lane));
  return v;
}

// ---------------------------------------------------------------- DSMEM exchange
// Exchange of K-split partial sums across the cluster, without a barrier: a row-owner warp stores
// its row of the CTA's full-width partial straight into the 3 peers' receive buffers with st.async,
// which credits the bytes to the PEER'S mbarrier (complete_tx).  Each CTA arms its own mbarrier with
// the byte count it expects and waits on it.  Receive buffers / mbarriers alternate with the
// running exchange index xe (buffer = xe & 1): a peer can only start exchange xe+2 (same buffer)
// after it completed xe+1, which needs MY xe+1 data, which I send after I have consumed xe -- so a
// buffer is never overwritten before it has been read, and phases never mix.
// Cost (scripts/ubench_dsmem.cu, measured): ~290 cycles + bytes / 21.6 B/clk per exchange -- 1000 cycles for
// the 15 KB a CTA sends and receives here, 650 for 7.5 KB, 460 for 3.75 KB; st.async.v4 is no faster than .v2
// and cp.async.bulk shared::cta -> shared::cluster is 10% slower, so the volume is what matters.  mapa compiles
// to one PRMT, hoisting it out of the loop was measured slightly slower (63.3 vs 62.5 us/step).
__device__ __forceinline__ void st_async_f2(uint32_t dst, float2 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(dst),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void send_row(const Smem& s, uint32_t xe, uint32_t rank, int row, int lane, const Row4& v) {
  float* ps = s.Ps(xe & 1);
  uint64_t* bar = s.xbar(xe & 1);
#pragma unroll
  for (uint32_t d = 1; d < kCluster; ++d) {
    const uint32_t peer = (rank + d) & (kCluster - 1);
    const uint32_t slot = (rank < peer) ? rank : rank - 1;   // my slot in the peer's [3][rows][128] buffer
    const uint32_t dst = map_to_rank(ps + (slot * kRMax + row) * 128 + 2 * lane, peer);
    const uint32_t rbar = map_to_rank(bar, peer);
    st_async_f2(dst, v.lo, rbar);
    st_async_f2(dst + 64 * 4, v.hi, rbar);
  }
}
template <int RT>
__device__ __forceinline__ void exchange_arm(const Smem& s, uint32_t xe, int tid) {
  if (tid == kIssuerTid) mbar_arrive_expect_tx(s.xbar(xe & 1), 3u * RT * 128u * 4u);
}
__device__ __forceinline__ void exchange_wait(const Smem& s, uint32_t xe) {
  mbar_wait(s.xbar(xe & 1), (xe >> 1) & 1);
}
__device__ __forceinline__ Row4 add_peers(const float* Ps, int row, int lane, Row4 v) {
#pragma unroll
  for (int q = 0; q < 3; ++q) v = add4(v, ld_row4(Ps + (q * kRMax + row) * 128, lane));
  return v;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// LayerNorm of one 128-wide row held 4 values per lane (nn.LayerNorm: biased variance, eps 1e-5).
// One butterfly instead of two: sum and sum of squares of d = x - c are reduced together, with the
// shift c = first element of the row (one extra shuffle), so that var = E[d^2] - E[d]^2 does not
// cancel (|E[d]| is of the order of the row's standard deviation, not of its mean).
__device__ __forceinline__ void layernorm1(Row4& v, const float* lnp, int lane) {
  const Row4 g = ld_row4(lnp, lane), be = ld_row4(lnp + 128, lane);
  const float c = __shfl_sync(0xffffffffu, v.lo.x, 0);
  v.lo.x -= c;
  v.lo.y -= c;
  v.hi.x -= c;
  v.hi.y -= c;
  float s1 = (v.lo.x + v.lo.y) + (v.hi.x + v.hi.y);
  float s2 = (v.lo.x * v.lo.x + v.lo.y * v.lo.y) + (v.hi.x * v.hi.x + v.hi.y * v.hi.y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float mean = s1 * (1.0f / 128.0f);
  const float var = fmaxf(fmaf(-mean, mean, s2 * (1.0f / 128.0f)), 0.f);
  const float rstd = rsqrtf(var + kLnEps);
  v.lo = make_float2((v.lo.x - mean) * rstd * g.lo.x + be.lo.x, (v.lo.y - mean) * rstd * g.lo.y + be.lo.y);
  v.hi = make_float2((v.hi.x - mean) * rstd * g.hi.x + be.hi.x, (v.hi.y - mean) * rstd * g.hi.y + be.hi.y);
}

// The bias / LayerNorm vectors of a stage are copied out of the weight tile so that the ring slot can be handed
// back to the TMA engine before the epilogue runs.  All 10 warps copy: leaving it to the two warps without a
// GEMM role was measured 3% slower per step (64.5 vs 62.5 us), their 6 dependent LDS/STS rounds end after the GEMM.
// AMUSE_COPY_AFTER_GEMM (untested candidate): issue the copy after the GEMM call instead of
// before it, so that its LDS -> STS round trip is not in front of every warp's first GEMM load.
#ifdef AMUSE_COPY_AFTER_GEMM
constexpr bool kCopyAfterGemm = true;
#else
constexpr bool kCopyAfterGemm = false;
#endif
__device__ __forceinline__ void copy_params(float* dst, const float* src, int n, int tid) {
  for (int i = tid; i < n; i += kThreads) dst[i] = src[i];
}


#define AMUSE_PROF(slot)                                     \
  do {                                                       \
    if (do_prof && tid == 0) p.prof[(slot)] = clock64();     \
  } while (0)

// Developer build (-DAMUSE_FINE_PROF): extra stamps inside the stages of layer 1 (slots 112..127).
#ifdef AMUSE_FINE_PROF
#define AMUSE_FINE(slot)                                                  \
  do {                                                                    \
    if (do_prof && layer == 1 && tid == 0) p.prof[(slot)] = clock64();    \
  } while (0)
#else
#define AMUSE_FINE(slot) do { } while (0)
#endif

}  // namespace

// ================================================================= the kernel
// RB = row blocks of 5 rows: 1 (one clip per cluster) or 2 (two clips).  8 warps = RB x KS.
// PRUNE: after the last layer only token 0 of every clip is read (encoder.norm -> eps, denoiser.py:188),
// so its out_proj / FFN1 / FFN2 are evaluated for RB rows instead of 5*RB: all 8 warps split K, every
// weight tile is read from shared memory once, and the exchanges carry RB rows.  K and V of the last
// layer still need all tokens, so the skip-linear, QKV and attention stages are unchanged.
// WIDE (RB == 2 only): instead of 2 row blocks x 4 K slices, every GEMM warp computes all 10 rows over 1/8
// of K.  Each weight word is then read from shared memory once per CTA instead of twice and a warp issues 80
// FFMA2 per 18 shared-memory loads instead of 40 per 13: with 5-row warps the LDS pipe is as busy as the FMA
// pipe (8 warps x 27 LSU cycles vs 2 warps/SMSP x 40 x 2.75 issue cycles per 4-k round).  The price is 8 partial
// tiles to park instead of 4; the 4 extra slots live in the Ps receive buffer of the exchange AFTER the next
// one (3 slots) and in one extra 5 KB buffer.  That Ps buffer is idle: with X = the next exchange this CTA
// will send, peers may already be writing X data into Ps[X & 1], but nobody can send X+1 data before it has
// received this CTA's X rows, and every row is sent by its owner warp after that warp's last read of the
// parked partials (the sent value depends on those loads).  The previous contents (exchange X-1) were consumed
// before the __syncthreads() that ended that exchange's epilogue.
template <int RB, bool PRUNE, bool WIDE>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    denoise_loop_kernel(const Params p) {
  static_assert(!WIDE || RB == 2, "WIDE is the 2-clip layout");
  constexpr int KS = WIDE ? 8 : 8 / RB;   // K-split factor of every GEMM
  constexpr int RT = RB * 5;              // activation rows computed by the GEMMs
  constexpr int NRW = WIDE ? 10 : 5;      // rows per GEMM warp
  extern __shared__ __align__(128) float smem_raw[];
  Smem s;
  s.base = smem_raw;
  float* const Xs = s.at(Smem::oXs);
  float* const SK = s.at(Smem::oSK);
  float* const QKVs = s.at(Smem::oQKV);
  float* const Oh = s.at(Smem::oOh);
  float* const Hs = s.at(Smem::oHs);
  float* const RED = s.at(Smem::oRED);
  float* const Cs = s.at(Smem::oCs);
  float* const pe01 = s.at(Smem::oPe);
  float* const zs = s.at(Smem::oZs);
  float* const Es = s.at(Smem::oEs);
  float* const tembs = s.at(Smem::oTemb);
  float* const fn = s.at(Smem::oFn);
  float* const par_bqkv = s.at(Smem::oBqkv);
  float* const par_tail = s.at(Smem::oTail);   // bias | LN weight | LN bias of the stage in flight

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();   // == attention head owned by this CTA
  const int cid = static_cast<int>(cluster_id_x());
  const bool gw = warp < kGemmWarps;         // warps 0..7 run the GEMMs, warps 8..9 only epilogues
  const int rb = WIDE ? 0 : warp / KS, ks = warp % KS;  // GEMM role: row block, K slice
  // reduce / epilogue role: warp w owns activation row w -- one row per warp, so no warp carries a
  // second row on the critical path of the ~40 epilogues per step
  const int row0 = warp;
  const bool own0 = row0 < RT;
  const int T = p.T;
  const int s_base = cid * p.S;
  const int S = min(p.S, p.B - s_base);      // clips of this cluster (>= 1 by grid construction)
  const int R = S * T;                       // live activation rows (compact: row = clip*T + token)

  // ---- one-time setup
  for (int i = tid; i < kActFloats; i += kThreads) Xs[i] = 0.f;   // every activation buffer
  if (tid == 0) {
    mbar_init(s.wbar(0), 1);
    mbar_init(s.wbar(1), 1);
    mbar_init(s.xbar(0), 1);
    mbar_init(s.xbar(1), 1);
    fence_mbar_init();
  }
  __syncthreads();
  for (int i = tid; i < S * 3 * 128; i += kThreads) Cs[i] = p.cond[static_cast<size_t>(s_base) * 384 + i];
  for (int i = tid; i < 256; i += kThreads) {
    pe01[i] = p.pe01[i];
    fn[i] = p.final_norm[i];
  }
  for (int i = tid; i < S * 128; i += kThreads) zs[i] = p.latents0[static_cast<size_t>(s_base) * 128 + i];

  WPipe wp;
  wp.blob = p.blob + static_cast<size_t>(rank) * kBlobRankFloats;
  wp.g = 0;
  wp.total = static_cast<uint32_t>(p.n_steps) * kTilesPerStep;
  if (tid == 0) {
    wp_issue(s, wp, 0);
    wp_issue(s, wp, 1);
  }
  uint32_t xe = 0;   // running index of the DSMEM exchange (selects receive buffer + mbarrier phase)
  __syncthreads();
  if (warp == kIssuerWarp) wp_prewait(s, wp, 0);   // tile 0 of step 0 (ordered by the cluster barrier below)
  cluster_sync_all();   // every CTA of the cluster is resident, zero-filled and has its mbarriers
                        // initialised before any peer stores into its shared memory

  // Stateless Philox draw per (global latent element, step) -- philox.cuh -- so the noise a clip sees does not
  // depend on how clips are packed into clusters or sharded over GPUs.
  const bool use_rng = (p.step_noise == nullptr);
  const bool owns_elem = tid < S * 128;      // thread <-> latent element (S*128 <= 256)
  const unsigned long long rng_elem = p.seed_elem_base + static_cast<unsigned long long>(s_base) * 128ull + tid;

  // software prefetch (one step ahead) of the tiny per-step global reads
  float temb_next = (tid < 128) ? __ldg(p.temb + tid) : 0.f;
  float coef_next[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) coef_next[q] = __ldg(p.coef + q);
  float noise_next = 0.f;
  if (!use_rng && owns_elem) noise_next = __ldg(p.step_noise + static_cast<size_t>(s_base) * 128 + tid);

  // attention roles (see the attention stage): lane -> (row slot, key j, half c of the head dimension)
  const int at_slot = (lane < 30) ? lane / 10 : 0;
  const int at_l = lane - (lane / 10) * 10;                    // position inside the slot: j * 2 + c
  const bool at_live = (lane < 30) && (warp * 3 + at_slot < R);
  const int at_row = at_live ? warp * 3 + at_slot : 0;         // my query row (compact: clip * T + token)
  const int at_cb = (at_row / T) * T;                          // first row of that clip
  const int at_c = at_l & 1;
  const bool at_key = at_live && (at_l >> 1) < T;              // my key exists
  const int at_j = at_key ? (at_l >> 1) : 0;
  const int at_src = at_slot * 10;                             // lane holding (j = 0, c = 0) of my slot
  const bool at_pv = at_live && at_l < 8;

  // Parked K-split partials.  WIDE: slot q of 8 (see the kernel comment); otherwise RED[ks][RT][128].
  auto wslot = [&](int q) -> float* {
    const int off = (q < 4) ? Smem::oRED + q * kRedSlot
                            : (q < 7) ? Smem::oPs + static_cast<int>((xe + 1) & 1) * kPsFloats + (q - 4) * kRedSlot
                                      : Smem::oRed7;
    return s.at(off);
  };
  auto park_tile = [&](const float (&acc)[NRW][4]) {
    if constexpr (WIDE) {
      float* dst = wslot(ks);
#pragma unroll
      for (int i = 0; i < NRW; ++i) {
        Row4 v;
        v.lo = make_float2(acc[i][0], acc[i][1]);
        v.hi = make_float2(acc[i][2], acc[i][3]);
        st_row4(dst + i * 128, lane, v);
      }
    } else {
      park4<RT>(RED, rb, ks, lane, acc);
    }
  };
  auto gather_row = [&](int row) -> Row4 {
    if constexpr (WIDE) {
      Row4 v = ld_row4(wslot(0) + row * 128, lane);
#pragma unroll
      for (int q = 1; q < 8; ++q) v = add4(v, ld_row4(wslot(q) + row * 128, lane));
      return v;
    } else {
      return gather4<RT, KS>(RED, row, lane);
    }
  };

  // K-split partial -> st.async exchange -> sum + bias [+ residual] [+ LayerNorm] -> Xs [, skip stack].
  // A lambda so that the three users (skip fusion, out_proj, FFN2) share one source form.
  auto exchange_epilogue = [&](const float* bias, const float* resid, const float* lnp, float* dst2, int prof_slot,
                               bool do_prof) {
    Row4 v;
    if (own0) {
      v = gather_row(row0);
      send_row(s, xe, rank, row0, lane, v);
      v = add4(v, ld_row4(bias, lane));
      if (resid) v = add4(v, ld_row4(resid + row0 * 128, lane));
    }
    if (!kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 0);   // next stage's tile, under the exchange wait
    if (own0) {
      if (prof_slot >= 0 && prof_slot < 12 && do_prof && tid == 0) p.prof[prof_slot + 100] = clock64();   // layer 0 only
      exchange_wait(s, xe);
      if (prof_slot >= 0 && do_prof && tid == 0) p.prof[prof_slot] = clock64();
      v = add_peers(s.Ps(xe & 1), row0, lane, v);
      if (lnp) layernorm1(v, lnp, lane);
      st_row4(Xs + row0 * 128, lane, v);
      if (dst2) st_row4(dst2 + row0 * 128, lane, v);
    }
    ++xe;
    __syncthreads();
  };

  // Pruned-row variant (last layer): warp i < RB owns token 0 of clip i, i.e. activation row i*T.  A
  // cluster with fewer clips than RB still sends RB rows (the spare row is finite filler that nobody
  // reads), so the byte count every CTA arms its mbarrier with is a compile-time constant.
  auto exchange_epilogue_pruned = [&](const float* bias, const float* lnp, int prof_slot, bool do_prof) {
    if (warp < RB) {
      const int arow = warp * T;
      Row4 v = gather4<RB, 8>(RED, warp, lane);
      send_row(s, xe, rank, arow, lane, v);
      v = add4(add4(v, ld_row4(bias, lane)), ld_row4(Xs + arow * 128, lane));
      exchange_wait(s, xe);
      if (prof_slot >= 0 && do_prof && tid == 0) p.prof[prof_slot] = clock64();
      v = add_peers(s.Ps(xe & 1), arow, lane, v);
      layernorm1(v, lnp, lane);
      st_row4(Xs + arow * 128, lane, v);
    }
    else if (!kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 0);
    ++xe;
    __syncthreads();
  };

  for (int step = 0; step < p.n_steps; ++step) {
    const bool do_prof = (p.prof != nullptr) && cid == 0 && rank == 0 && step == p.prof_step;
    AMUSE_PROF(0);
    float coef[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) coef[q] = coef_next[q];
    const float noise = noise_next;
    if (tid < 128) tembs[tid] = temb_next;
    if (step + 1 < p.n_steps) {
      if (tid < 128) temb_next = __ldg(p.temb + static_cast<size_t>(step + 1) * 128 + tid);
#pragma unroll
      for (int q = 0; q < 5; ++q) coef_next[q] = __ldg(p.coef + static_cast<size_t>(step + 1) * 5 + q);
      if (!use_rng && owns_elem)
        noise_next = __ldg(p.step_noise + (static_cast<size_t>(step + 1) * p.B + s_base) * 128 + tid);
    }
    __syncthreads();

    // ---- assemble the token rows (denoiser.py:174-181 + position_encoding.py:156)
    for (int idx = tid; idx < R * 32; idx += kThreads) {
      const int r = idx >> 5, c4 = (idx & 31) * 4;
      const int sl = r / T, tok = r - sl * T;
      float4 v;
      if (tok == 0) {
        v = add4(*reinterpret_cast<const float4*>(zs + sl * 128 + c4), *reinterpret_cast<const float4*>(pe01 + c4));
      } else if (tok == 1) {
        v = add4(*reinterpret_cast<const float4*>(tembs + c4), *reinterpret_cast<const float4*>(pe01 + 128 + c4));
      } else {
        v = *reinterpret_cast<const float4*>(Cs + (sl * 3 + tok - 2) * 128 + c4);
      }
      *reinterpret_cast<float4*>(Xs + r * 128 + c4) = v;
    }
    __syncthreads();
    AMUSE_PROF(1);

    for (int layer = 0; layer < kLayers; ++layer) {
      // Straight-line stages (a "stage machine" with one shared copy of the GEMM / epilogue code was
      // tried to keep the loop inside the 32 KB instruction cache: it was 7% slower overall).
      // =============== output blocks: x = Linear(256->128)(cat(x, xs.pop())), K-split 64 per CTA
      if (layer >= 5) {
        const float* wt = wp_acquire_prewaited(s, wp);
        exchange_arm<RT>(s, xe, tid);
        if (!kCopyAfterGemm) copy_params(par_tail, wt + 64 * 128, 128, tid);
        // my K slice of cat(x, skip): ranks 0,1 -> x[:, 64*rank ..], ranks 2,3 -> skip[:, 64*(rank-2) ..]
        const float* src = (rank < 2) ? (Xs + rank * 64) : (SK + (8 - layer) * kRMax * 128 + (rank - 2) * 64);
        if (!gw) {
          if (kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 1);   // next tile, while the GEMM runs
        } else {
          float acc[NRW][4];
          gemm_rows<NRW, 128, 4, (64 / KS) / 4>(src + (rb * 5) * 128 + ks * (64 / KS), 128,
                                                wt + (ks * (32 / KS)) * 256 + lane * 2, acc);
          park_tile(acc);
        }
        if (kCopyAfterGemm) copy_params(par_tail, wt + 64 * 128, 128, tid);
        __syncthreads();
        wp_release(s, wp, tid);
        exchange_epilogue(par_tail, nullptr, nullptr, nullptr, -1, do_prof);
      }
      AMUSE_PROF(2 + layer * 10 + 0);

      // =============== QKV of my head (cross_attention.py:264-266, nn.MultiheadAttention in_proj)
      {
        const float* wt = wp_acquire_prewaited(s, wp);
        AMUSE_FINE(120);
        if (!kCopyAfterGemm) copy_params(par_bqkv, wt + 128 * 96, 96, tid);
        if (!gw) {
          if (kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 1);   // next tile, while the GEMM runs
        } else {
          float acc[NRW][3];
          gemm_rows<NRW, 96, 3, (128 / KS) / 4>(Xs + (rb * 5) * 128 + ks * (128 / KS), 128,
                                                wt + (ks * (64 / KS)) * 192 + lane * 2, acc);
          AMUSE_FINE(121);
          float* dst = (WIDE ? wslot(ks) : RED + ((ks * RT + rb * 5) * 96)) + lane;
#pragma unroll
          for (int i = 0; i < NRW; ++i) {
            dst[i * 96] = acc[i][0];
            dst[i * 96 + 32] = acc[i][1];
            dst[i * 96 + 64] = acc[i][2];
          }
        }
        if (kCopyAfterGemm) copy_params(par_bqkv, wt + 128 * 96, 96, tid);
        __syncthreads();
        AMUSE_FINE(122);
        wp_release(s, wp, tid);
        // (The same sums as 24 lanes x float4 instead of 32 lanes x 3 scalars were measured slower: 61.1 vs 60.6 us/step.)
        if (own0) {
          float q = par_bqkv[lane], k = par_bqkv[32 + lane], v = par_bqkv[64 + lane];
#pragma unroll
          for (int c = 0; c < KS; ++c) {
            const float* src = (WIDE ? wslot(c) + row0 * 96 : RED + ((c * RT + row0) * 96)) + lane;
            q += src[0];
            k += src[32];
            v += src[64];
          }
          float* dst = QKVs + row0 * kQkvLd + lane;
          dst[0] = q * 0.17677669529663687f;   // nn.MultiheadAttention scales q (after bias) by head_dim^-0.5
          dst[32] = k;
          dst[64] = v;
        }
        AMUSE_FINE(123);
        __syncthreads();
      }
      AMUSE_PROF(2 + layer * 10 + 1);

      // =============== attention of my head.  A lone warp runs at ~6 cycles per dependent instruction
      // here (measured), so the work is laid out for the shortest per-warp instruction chain: 3 query rows
      // per warp (warps 0..3), 10 lanes per row = 5 keys x 2 halves of the head dimension.
      if (warp < 4) {
        const float4* qv = reinterpret_cast<const float4*>(QKVs + at_row * kQkvLd + 16 * at_c);
        const float4* kv = reinterpret_cast<const float4*>(QKVs + (at_cb + at_j) * kQkvLd + 32 + 16 * at_c);
        float p0, p1, p2, p3;
        {
          const float4 a = qv[0], b = kv[0];
          p0 = a.x * b.x;
          p1 = a.y * b.y;
          p2 = a.z * b.z;
          p3 = a.w * b.w;
        }
#pragma unroll
        for (int c = 1; c < 4; ++c) {
          const float4 a = qv[c], b = kv[c];
          p0 = fmaf(a.x, b.x, p0);
          p1 = fmaf(a.y, b.y, p1);
          p2 = fmaf(a.z, b.z, p2);
          p3 = fmaf(a.w, b.w, p3);
        }
        float sc = (p0 + p1) + (p2 + p3);
        sc += __shfl_xor_sync(0xffffffffu, sc, 1);          // the two halves of the head dimension
        float sj[5];
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) sj[jj] = __shfl_sync(0xffffffffu, sc, at_src + 2 * jj);
        float m = sj[0];
#pragma unroll
        for (int jj = 1; jj < 5; ++jj) m = (jj < T) ? fmaxf(m, sj[jj]) : m;
        const float e = at_key ? expf(sc - m) : 0.f;
        float ej[5];
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) ej[jj] = __shfl_sync(0xffffffffu, e, at_src + 2 * jj);
        const float sum = ((ej[0] + ej[1]) + (ej[2] + ej[3])) + ej[4];
        if (at_pv) {   // 8 lanes per row: 4 head dimensions each
          const float* vb = QKVs + at_cb * kQkvLd + 64 + 4 * at_l;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int jj = 0; jj < 5; ++jj)
            if (jj < T) {
              const float4 v4 = *reinterpret_cast<const float4*>(vb + jj * kQkvLd);
              acc.x = fmaf(ej[jj], v4.x, acc.x);
              acc.y = fmaf(ej[jj], v4.y, acc.y);
              acc.z = fmaf(ej[jj], v4.z, acc.z);
              acc.w = fmaf(ej[jj], v4.w, acc.w);
            }
          const float inv = __frcp_rn(sum);
          *reinterpret_cast<float4*>(Oh + at_row * kOhLd + 4 * at_l) =
              make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        }
      }
      else if (!kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 0);   // out_proj tile, during the attention
      __syncthreads();
      AMUSE_PROF(2 + layer * 10 + 2);

      if (PRUNE && layer == kLayers - 1) {
        // =============== last layer, token 0 of every clip only (see PRUNE above)
        {   // out_proj
          const float* wt = wp_acquire_prewaited(s, wp);
          exchange_arm<RB>(s, xe, tid);
          if (!kCopyAfterGemm) copy_params(par_tail, wt + 32 * 128, kTileTail, tid);
          if (!gw) {
          if (kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 1);   // next tile, while the GEMM runs
        } else {
            float acc[RB][4];
            gemm_rows<RB, 128, 4, 1>(Oh + warp * 4, T * kOhLd, wt + (warp * 2) * 256 + lane * 2, acc);
            park_rows<RB>(RED, warp, lane, acc);
          }
          if (kCopyAfterGemm) copy_params(par_tail, wt + 32 * 128, kTileTail, tid);
          __syncthreads();
          wp_release(s, wp, tid);
          exchange_epilogue_pruned(par_tail, par_tail + 128, 2 + layer * 10 + 3, do_prof);
        }
        AMUSE_PROF(2 + layer * 10 + 4);
        {   // FFN1 + erf-GELU
          const float* wt = wp_acquire_prewaited(s, wp);
          if (!kCopyAfterGemm) copy_params(par_tail, wt + 128 * 128, 128, tid);
          if (!gw) {
          if (kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 1);   // next tile, while the GEMM runs
        } else {
            float acc[RB][4];
            gemm_rows<RB, 128, 4, 4>(Xs + warp * 16, T * 128, wt + (warp * 8) * 256 + lane * 2, acc);
            park_rows<RB>(RED, warp, lane, acc);
          }
          if (kCopyAfterGemm) copy_params(par_tail, wt + 128 * 128, 128, tid);
          __syncthreads();
          wp_release(s, wp, tid);
          if (warp < RB) {
            Row4 v = add4(gather4<RB, 8>(RED, warp, lane), ld_row4(par_tail, lane));
            v.lo = make_float2(gelu_erf(v.lo.x), gelu_erf(v.lo.y));
            v.hi = make_float2(gelu_erf(v.hi.x), gelu_erf(v.hi.y));
            st_row4(Hs + (warp * T) * 128, lane, v);
          }
          __syncthreads();
        }
        AMUSE_PROF(2 + layer * 10 + 5);
        {   // FFN2
          const float* wt = wp_acquire_ffn2(s, wp);
          exchange_arm<RB>(s, xe, tid);
          if (!kCopyAfterGemm) copy_params(par_tail, wt + 128 * 128, kTileTail, tid);
          if (!gw) {
          if (kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 1);   // next tile, while the GEMM runs
        } else {
            float acc[RB][4];
            gemm_rows<RB, 128, 4, 4>(Hs + warp * 16, T * 128, wt + (warp * 8) * 256 + lane * 2, acc);
            park_rows<RB>(RED, warp, lane, acc);
          }
          if (kCopyAfterGemm) copy_params(par_tail, wt + 128 * 128, kTileTail, tid);
          __syncthreads();
          wp_release(s, wp, tid);
          exchange_epilogue_pruned(par_tail, par_tail + 128, 2 + layer * 10 + 6, do_prof);
        }
        AMUSE_PROF(2 + layer * 10 + 7);
        continue;
      }
      // =============== out_proj, K-split by head -> st.async partial exchange -> sum + LN1
      {
        const float* wt = wp_acquire_prewaited(s, wp);
        AMUSE_FINE(124);
        exchange_arm<RT>(s, xe, tid);
        if (!kCopyAfterGemm) copy_params(par_tail, wt + 32 * 128, kTileTail, tid);
        if (!gw) {
          if (kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 1);   // next tile, while the GEMM runs
        } else {
          float acc[NRW][4];
          gemm_rows<NRW, 128, 4, (32 / KS) / 4>(Oh + (rb * 5) * kOhLd + ks * (32 / KS), kOhLd,
                                                wt + (ks * (16 / KS)) * 256 + lane * 2, acc);
          AMUSE_FINE(125);
          park_tile(acc);
        }
        if (kCopyAfterGemm) copy_params(par_tail, wt + 32 * 128, kTileTail, tid);
        __syncthreads();
        AMUSE_FINE(126);
        wp_release(s, wp, tid);
        exchange_epilogue(par_tail, Xs, par_tail + 128, nullptr, 2 + layer * 10 + 3, do_prof);
      }
      AMUSE_PROF(2 + layer * 10 + 4);

      // =============== FFN1: my 128 hidden units, erf-GELU
      {
        const float* wt = wp_acquire_prewaited(s, wp);
        AMUSE_FINE(112);
        if (!kCopyAfterGemm) copy_params(par_tail, wt + 128 * 128, 128, tid);
        if (!gw) {
          if (kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 1);   // next tile, while the GEMM runs
        } else {
          float acc[NRW][4];
          gemm_rows<NRW, 128, 4, (128 / KS) / 4>(Xs + (rb * 5) * 128 + ks * (128 / KS), 128,
                                                 wt + (ks * (64 / KS)) * 256 + lane * 2, acc);
          AMUSE_FINE(113);
          park_tile(acc);
        }
        if (kCopyAfterGemm) copy_params(par_tail, wt + 128 * 128, 128, tid);
        __syncthreads();
        AMUSE_FINE(114);
        wp_release(s, wp, tid);
        if (own0) {
          Row4 v = add4(gather_row(row0), ld_row4(par_tail, lane));
          v.lo = make_float2(gelu_erf(v.lo.x), gelu_erf(v.lo.y));
          v.hi = make_float2(gelu_erf(v.hi.x), gelu_erf(v.hi.y));
          st_row4(Hs + row0 * 128, lane, v);
        }
        AMUSE_FINE(115);
        __syncthreads();
      }
      AMUSE_PROF(2 + layer * 10 + 5);

      // =============== FFN2, K-split over my 128 hidden units -> exchange -> sum + LN2
      {
        const float* wt = wp_acquire_ffn2(s, wp);
        AMUSE_FINE(116);
        exchange_arm<RT>(s, xe, tid);
        if (!kCopyAfterGemm) copy_params(par_tail, wt + 128 * 128, kTileTail, tid);
        if (!gw) {
          if (kPrewaitInGemm && warp == kIssuerWarp) wp_prewait(s, wp, 1);   // next tile, while the GEMM runs
        } else {
          float acc[NRW][4];
          gemm_rows<NRW, 128, 4, (128 / KS) / 4>(Hs + (rb * 5) * 128 + ks * (128 / KS), 128,
                                                 wt + (ks * (64 / KS)) * 256 + lane * 2, acc);
          AMUSE_FINE(117);
          park_tile(acc);
        }
        if (kCopyAfterGemm) copy_params(par_tail, wt + 128 * 128, kTileTail, tid);
        __syncthreads();
        AMUSE_FINE(118);
        wp_release(s, wp, tid);
        exchange_epilogue(par_tail, Xs, par_tail + 128, (layer < 4) ? (SK + layer * kRMax * 128) : nullptr,
                          2 + layer * 10 + 6, do_prof);
      }
      AMUSE_PROF(2 + layer * 10 + 7);
    }   // layers

    // ---- encoder.norm on token 0 of every clip -> eps (cross_attention.py:62-63, denoiser.py:188)
    if (warp < S) {
      const float4 v = *reinterpret_cast<const float4*>(Xs + (warp * T) * 128 + lane * 4);
      const float4 g = *reinterpret_cast<const float4*>(fn + lane * 4);
      const float4 b = *reinterpret_cast<const float4*>(fn + 128 + lane * 4);
      *reinterpret_cast<float4*>(Es + warp * 128 + lane * 4) = warp_layernorm128(v, g, b);
    }
    __syncthreads();
    // ---- scheduler step (K2), replicated in every CTA; op order of diffusers' step():
    //      x0 = (x - sqrt(1-a) e) / sqrt(a); clamp; x' = c2 x0 + c3 (e | x) + sigma z
    if (owns_elem) {
      const float x = zs[tid], e = Es[tid];
      float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(coef[1], e)), coef[0]);
      if (p.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
      float out = __fadd_rn(__fmul_rn(coef[2], x0), __fmul_rn(coef[3], p.dir_uses_eps ? e : x));
      if (coef[4] != 0.f) {
        const float zn = use_rng ? philox_normal(p.seed, rng_elem, static_cast<uint32_t>(step)) : noise;
        out = __fadd_rn(out, __fmul_rn(coef[4], zn));
      }
      zs[tid] = out;
    }
    __syncthreads();
    AMUSE_PROF(2 + kLayers * 10);
  }   // steps

  if (rank == 0)
    for (int i = tid; i < S * 128; i += kThreads) p.latents_out[static_cast<size_t>(s_base) * 128 + i] = zs[i];
  cluster_sync_all();   // nobody leaves while a peer could still address its shared memory
}

size_t smem_bytes() { return static_cast<size_t>(kSmemFloats) * sizeof(float); }

cudaError_t launch(const Params& p, cudaStream_t stream) {
  using Kernel = void (*)(const Params);
  static const Kernel kernels[3][2] = {{denoise_loop_kernel<1, false, false>, denoise_loop_kernel<1, true, false>},
                                       {denoise_loop_kernel<2, false, false>, denoise_loop_kernel<2, true, false>},
                                       {denoise_loop_kernel<2, false, true>, denoise_loop_kernel<2, true, true>}};
  static std::mutex mu;                  // two contexts on two threads may reach their first launch together
  static bool configured_dev[64] = {};   // attributes and __constant__ data are per device
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  bool& configured = configured_dev[dev & 63];
  if (!configured) {
    int2 tab[kTilesPerStep];
    for (int i = 0; i < kTilesPerStep; ++i) tile_info(i, tab[i].x, tab[i].y);
    cudaError_t e0 = cudaMemcpyToSymbol(c_tile_tab, tab, sizeof(tab));
    if (e0 != cudaSuccess) return e0;
    for (int i = 0; i < 6; ++i) {
      cudaError_t e = cudaFuncSetAttribute(kernels[i >> 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem_bytes()));
      if (e != cudaSuccess) return e;
    }
    configured = true;
  }
  if (p.S < 1 || p.S > kSMax) return cudaErrorInvalidValue;
  const int n_clusters = (p.B + p.S - 1) / p.S;
  kernels[(p.S == 2 && p.wide_rows) ? 2 : p.S - 1][p.prune_last ? 1 : 0]<<<dim3(n_clusters * kCluster), dim3(kThreads), smem_bytes(), stream>>>(p);
  return cudaGetLastError();
}

}  // namespace dn
}  // namespace amuse
