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

#include <curand_kernel.h>

#include "common.cuh"

namespace amuse {
namespace dn {

namespace {

constexpr int kWBufFloats = 16768;   // == kTileW2 (largest tile), multiple of 32 floats
static_assert(kWBufFloats >= kTileMax && kWBufFloats % 32 == 0, "weight ring slot too small");
constexpr int kQkvLd = 100;          // q|k|v row stride: 16-B aligned rows, conflict-free T x T dot products
constexpr int kOhLd = 36;            // attention-output row stride (16-B aligned rows)
constexpr int kRedFloats = 7 * 5 * 128;   // K-split partial sums: (KSPLIT-1) x rows x 128
constexpr int kPsFloats = 3 * kRMax * 128;   // partial sums received from the 3 peer CTAs

// ---------------------------------------------------------------- shared-memory carve-up
// Offsets (floats) from the dynamic shared-memory base.  Every pointer is formed as
// `smem + constant`, so the compiler keeps the .shared address space (LDS/STS, not generic LD/ST).
constexpr int kActFloats = kRMax * 128 + 4 * kRMax * 128 + 2 * kPsFloats + kRMax * kQkvLd + kRMax * kOhLd +
                           kRMax * 128 + kRedFloats + kSMax * 3 * 128 + 256 + kSMax * 128 + kSMax * 128 + 128 + 256 +
                           96 + 128 + 256 + 128 + 128 + 256 + 128;
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
  static constexpr int oHs = oOh + kRMax * kOhLd;            // [10][128] my 128 hidden units
  static constexpr int oRED = oHs + kRMax * 128;             // K-split partial sums
  static constexpr int oCs = oRED + kRedFloats;              // [2][3][128] condition tokens (+PE)
  static constexpr int oPe = oCs + kSMax * 3 * 128;          // [2][128]
  static constexpr int oZs = oPe + 256;                      // [2][128] current latents
  static constexpr int oEs = oZs + kSMax * 128;              // [2][128] eps
  static constexpr int oTemb = oEs + kSMax * 128;            // [128]
  static constexpr int oFn = oTemb + 128;                    // [256] final norm
  static constexpr int oBqkv = oFn + 256;                    // [96]
  static constexpr int oBo = oBqkv + 96;                     // [128]
  static constexpr int oLn1 = oBo + 128;                     // [256]
  static constexpr int oB1 = oLn1 + 256;                     // [128]
  static constexpr int oB2 = oB1 + 128;                      // [128]
  static constexpr int oLn2 = oB2 + 128;                     // [256]
  static constexpr int oBsk = oLn2 + 256;                    // [128]
  static constexpr int oBar = oBsk + 128;                    // 4 mbarriers: 2 weight ring, 2 exchange
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

__device__ __forceinline__ void wp_issue(const Smem& s, const WPipe& w, uint32_t n) {
  int off, cnt;
  tile_info(static_cast<int>(n % kTilesPerStep), off, cnt);
  uint64_t* bar = s.wbar(n & 1);
  mbar_arrive_expect_tx(bar, static_cast<uint32_t>(cnt) * 4u);
  bulk_g2s(s.wbuf(n & 1), w.blob + off, static_cast<uint32_t>(cnt) * 4u, bar);
}
__device__ __forceinline__ const float* wp_acquire(const Smem& s, const WPipe& w) {
  mbar_wait(s.wbar(w.g & 1), (w.g >> 1) & 1);
  return s.wbuf(w.g & 1);
}
// Call after a __syncthreads() that follows the last read of tile g: hands the buffer back
// to the TMA engine for tile g+2.
__device__ __forceinline__ void wp_release(const Smem& s, WPipe& w, int tid) {
  if (tid == 0 && w.g + 2 < w.total) {
    fence_proxy_async();
    wp_issue(s, w, w.g + 2);
  }
  ++w.g;
}

// ---------------------------------------------------------------- warp roles
// 8 warps = RB row-blocks (5 rows each) x KSPLIT = 8/RB K-slices.  The warp that folds the
// K-split partials of row-block rb (and runs its epilogue / LayerNorm) is the one with
// ks == rb, so the folding warps sit on different SM sub-partitions (warp % 4).
template <int RB>
struct Role {
  static constexpr int KS = 8 / RB;
  int rb, ks;
  bool fold;
  int slot;   // parking slot of a non-folding warp: 0 .. KS-2
  __device__ __forceinline__ explicit Role(int warp) {
    rb = warp / KS;
    ks = warp % KS;
    fold = (ks == rb);
    slot = (ks - rb - 1 + KS) % KS;
  }
};

// ---------------------------------------------------------------- 5-row FFMA2 micro-kernel
// acc[i][j] = sum_{k in my K slice} A[rb*5+i][k] * W[k][col(lane,j)]
// Blackwell issues a 3-register FFMA at half rate; the packed FFMA2 (fma.rn.f32x2, two fp32 FMAs
// on 64-bit register pairs) is what reaches 128 FMA/clk/SM.  The two lanes of an FFMA2 are two
// consecutive k: weight tiles are stored k-pair interleaved, Wt2[k/2][n][2], so one 64-bit
// word holds (W[k][n], W[k+1][n]); an A-row float4 supplies (A[k],A[k+1]) and (A[k+2],A[k+3]).
// Even-k and odd-k products accumulate in the two halves and are added at the end.
//   CONTIG (TC == 4): lane owns columns 4*lane .. 4*lane+3   (NCOL = 128, two LDS.128 per k-pair)
//   strided         : lane owns columns lane + 32*j           (NCOL = 32*TC, one LDS.64 per column)
// A-row loads are warp-uniform 128-bit broadcasts; weight loads are conflict-free.
template <int NCOL, int KDIM, int TC, int KSPLIT, bool CONTIG>
__device__ __forceinline__ void gemm5(const float* __restrict__ A, int lda, const float* __restrict__ Wt2, int rb,
                                      int ks, int lane, float (&acc)[5][TC]) {
  constexpr int KPER = KDIM / KSPLIT;
  static_assert(KPER % 4 == 0, "K slice must be a multiple of 4");
  static_assert(!CONTIG || TC == 4, "contiguous mapping owns 4 columns per lane");
  float2 acc2[5][TC];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc2[i][j] = make_float2(0.f, 0.f);
  const float* a0 = A + (rb * 5) * lda + ks * KPER;
  const float* w = Wt2 + (ks * (KPER / 2)) * (NCOL * 2) + (CONTIG ? lane * 8 : lane * 2);
#pragma unroll
  for (int kk = 0; kk < KPER; kk += 4) {
    float4 a[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * lda + kk);
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
      float2 wv[TC];
      const float* wr = w + (kk / 2 + kp) * (NCOL * 2);
      if (CONTIG) {
        const float4 t0 = *reinterpret_cast<const float4*>(wr);
        const float4 t1 = *reinterpret_cast<const float4*>(wr + 4);
        wv[0] = make_float2(t0.x, t0.y);
        wv[1] = make_float2(t0.z, t0.w);
        wv[2] = make_float2(t1.x, t1.y);
        wv[3] = make_float2(t1.z, t1.w);
      } else {
#pragma unroll
        for (int j = 0; j < TC; ++j) wv[j] = *reinterpret_cast<const float2*>(wr + 64 * j);
      }
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float2 av = (kp == 0) ? make_float2(a[i].x, a[i].y) : make_float2(a[i].z, a[i].w);
#pragma unroll
        for (int j = 0; j < TC; ++j) acc2[i][j] = __ffma2_rn(av, wv[j], acc2[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[i][j] = acc2[i][j].x + acc2[i][j].y;
}

// K-split reduction through shared memory: non-folding warps park their partials ...
template <int NCOL, int TC, bool CONTIG, int RB>
__device__ __forceinline__ void park(float* RED, const Role<RB>& r, int lane, const float (&acc)[5][TC]) {
  if (r.fold) return;
  float* dst = RED + ((r.slot * RB + r.rb) * 5) * NCOL;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    if (CONTIG) {
      *reinterpret_cast<float4*>(dst + i * NCOL + lane * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
#pragma unroll
      for (int j = 0; j < TC; ++j) dst[i * NCOL + lane + 32 * j] = acc[i][j];
    }
  }
}
// ... and after a __syncthreads() the folding warp of each row-block adds them up.
template <int NCOL, int TC, bool CONTIG, int RB>
__device__ __forceinline__ void fold(const float* RED, const Role<RB>& r, int lane, float (&acc)[5][TC]) {
#pragma unroll
  for (int q = 0; q < Role<RB>::KS - 1; ++q) {
    const float* src = RED + ((q * RB + r.rb) * 5) * NCOL;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      if (CONTIG) {
        const float4 t = *reinterpret_cast<const float4*>(src + i * NCOL + lane * 4);
        acc[i][0] += t.x;
        acc[i][1] += t.y;
        acc[i][2] += t.z;
        acc[i][3] += t.w;
      } else {
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] += src[i * NCOL + lane + 32 * j];
      }
    }
  }
}

// Exchange of K-split partial sums across the cluster, without a barrier: the folding warp stores
// its full-width partial (5 rows x 4 columns per lane) straight into the 3 peers' receive buffers
// with st.async, which credits the bytes to the PEER'S mbarrier (complete_tx).  Each CTA arms its
// own mbarrier with the byte count it expects and waits on it.  Receive buffers / mbarriers
// alternate with the running exchange index xe (buffer = xe & 1): a peer can only start exchange
// xe+2 (same buffer) after it completed xe+1, which needs MY xe+1 data, which I send after I have
// consumed xe -- so a buffer is never overwritten before it has been read, and phases never mix.
__device__ __forceinline__ void st_async_f4(uint32_t dst, float4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(dst), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
               "r"(__float_as_uint(v.w)), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void broadcast_partial(const Smem& s, uint32_t xe, uint32_t rank, int rb, int lane,
                                                  const float (&acc)[5][4]) {
  float* ps = s.Ps(xe & 1);
  uint64_t* bar = s.xbar(xe & 1);
#pragma unroll
  for (uint32_t d = 1; d < kCluster; ++d) {
    const uint32_t peer = (rank + d) & (kCluster - 1);
    const uint32_t slot = (rank < peer) ? rank : rank - 1;   // my slot in the peer's [3][rows][128] buffer
    const uint32_t base = map_to_rank(ps + (slot * kRMax + rb * 5) * 128 + lane * 4, peer);
    const uint32_t rbar = map_to_rank(bar, peer);
#pragma unroll
    for (int i = 0; i < 5; ++i)
      st_async_f4(base + i * 128 * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]), rbar);
  }
}
template <int RB>
__device__ __forceinline__ void exchange_arm(const Smem& s, uint32_t xe, int tid) {
  if (tid == 0) mbar_arrive_expect_tx(s.xbar(xe & 1), 3u * RB * 5u * 128u * 4u);
}
__device__ __forceinline__ void exchange_wait(const Smem& s, uint32_t xe) {
  mbar_wait(s.xbar(xe & 1), (xe >> 1) & 1);
}

// acc (my partial) + 3 peer partials + bias [+ residual] -> optional LayerNorm -> dst rows.
// One warp holds 5 full rows (32 lanes x 4 columns); the 5 LayerNorm chains are interleaved.
template <bool RESIDUAL, bool LN>
__device__ __forceinline__ void sum_norm_store(const float* Ps, const float* bias, const float* resid,
                                               const float* lnp, float* dst, float* dst2, int rb, int lane,
                                               float (&acc)[5][4]) {
  const float4 b = *reinterpret_cast<const float4*>(bias + lane * 4);
  float4 v[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    v[i] = make_float4(acc[i][0] + b.x, acc[i][1] + b.y, acc[i][2] + b.z, acc[i][3] + b.w);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(Ps + ((q * kRMax) + rb * 5 + i) * 128 + lane * 4);
      v[i].x += t.x;
      v[i].y += t.y;
      v[i].z += t.z;
      v[i].w += t.w;
    }
    if (RESIDUAL) {
      const float4 x = *reinterpret_cast<const float4*>(resid + (rb * 5 + i) * 128 + lane * 4);
      v[i].x += x.x;
      v[i].y += x.y;
      v[i].z += x.z;
      v[i].w += x.w;
    }
  }
  if (LN) {
    const float4 g = *reinterpret_cast<const float4*>(lnp + lane * 4);
    const float4 be = *reinterpret_cast<const float4*>(lnp + 128 + lane * 4);
    float s1[5], s2[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) s1[i] = v[i].x + v[i].y + v[i].z + v[i].w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < 5; ++i) s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float mean = s1[i] * (1.0f / 128.0f);
      v[i].x -= mean;
      v[i].y -= mean;
      v[i].z -= mean;
      v[i].w -= mean;
      s2[i] = v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < 5; ++i) s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float rstd = 1.0f / sqrtf(s2[i] * (1.0f / 128.0f) + kLnEps);
      v[i] = make_float4(v[i].x * rstd * g.x + be.x, v[i].y * rstd * g.y + be.y, v[i].z * rstd * g.z + be.z,
                         v[i].w * rstd * g.w + be.w);
    }
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    *reinterpret_cast<float4*>(dst + (rb * 5 + i) * 128 + lane * 4) = v[i];
    if (dst2) *reinterpret_cast<float4*>(dst2 + (rb * 5 + i) * 128 + lane * 4) = v[i];
  }
}

__device__ __forceinline__ void copy_params(float* dst, const float* src, int n, int tid) {
  for (int i = tid; i < n; i += kThreads) dst[i] = src[i];
}

#define AMUSE_PROF(slot)                                     \
  do {                                                       \
    if (do_prof && tid == 0) p.prof[(slot)] = clock64();     \
  } while (0)

}  // namespace

// ================================================================= the kernel
template <int RB>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    denoise_loop_kernel(const Params p) {
  constexpr int KS = 8 / RB;
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
  float* const par_bo = s.at(Smem::oBo);
  float* const par_ln1 = s.at(Smem::oLn1);
  float* const par_b1 = s.at(Smem::oB1);
  float* const par_b2 = s.at(Smem::oB2);
  float* const par_ln2 = s.at(Smem::oLn2);
  float* const par_bsk = s.at(Smem::oBsk);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();   // == attention head owned by this CTA
  const int cid = static_cast<int>(cluster_id_x());
  const Role<RB> role(warp);
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
  cluster_sync_all();   // every CTA of the cluster is resident, zero-filled and has its mbarriers
                        // initialised before any peer stores into its shared memory

  // Philox stream per latent element (subsequence = global element index), so the noise a clip
  // sees does not depend on how clips are packed into clusters or sharded over GPUs.
  curandStatePhilox4_32_10_t rng;
  const bool use_rng = (p.step_noise == nullptr);
  const bool owns_elem = tid < S * 128;      // thread <-> latent element (S*128 <= 256)
  if (use_rng && owns_elem)
    curand_init(p.seed, p.seed_elem_base + static_cast<unsigned long long>(s_base) * 128ull + tid, 0, &rng);

  // software prefetch (one step ahead) of the tiny per-step global reads
  float temb_next = (tid < 128) ? __ldg(p.temb + tid) : 0.f;
  float coef_next[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) coef_next[q] = __ldg(p.coef + q);
  float noise_next = 0.f;
  if (!use_rng && owns_elem) noise_next = __ldg(p.step_noise + static_cast<size_t>(s_base) * 128 + tid);

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
        const float4 z = *reinterpret_cast<const float4*>(zs + sl * 128 + c4);
        const float4 e = *reinterpret_cast<const float4*>(pe01 + c4);
        v = make_float4(z.x + e.x, z.y + e.y, z.z + e.z, z.w + e.w);
      } else if (tok == 1) {
        const float4 z = *reinterpret_cast<const float4*>(tembs + c4);
        const float4 e = *reinterpret_cast<const float4*>(pe01 + 128 + c4);
        v = make_float4(z.x + e.x, z.y + e.y, z.z + e.z, z.w + e.w);
      } else {
        v = *reinterpret_cast<const float4*>(Cs + (sl * 3 + tok - 2) * 128 + c4);
      }
      *reinterpret_cast<float4*>(Xs + r * 128 + c4) = v;
    }
    __syncthreads();
    AMUSE_PROF(1);

    for (int layer = 0; layer < kLayers; ++layer) {
      // =============== output blocks: x = Linear(256->128)(cat(x, xs.pop())), K-split 64 per CTA
      if (layer >= 5) {
        const float* wt = wp_acquire(s, wp);
        exchange_arm<RB>(s, xe, tid);
        copy_params(par_bsk, wt + 64 * 128, 128, tid);
        // my K slice of cat(x, skip): ranks 0,1 -> x[:, 64*rank ..], ranks 2,3 -> skip[:, 64*(rank-2) ..]
        const float* src = (rank < 2) ? (Xs + rank * 64) : (SK + (8 - layer) * kRMax * 128 + (rank - 2) * 64);
        float acc[5][4];
        gemm5<128, 64, 4, KS, true>(src, 128, wt, role.rb, role.ks, lane, acc);
        park<128, 4, true, RB>(RED, role, lane, acc);
        __syncthreads();
        wp_release(s, wp, tid);
        if (role.fold) {
          fold<128, 4, true, RB>(RED, role, lane, acc);
          broadcast_partial(s, xe, rank, role.rb, lane, acc);
          exchange_wait(s, xe);
          sum_norm_store<false, false>(s.Ps(xe & 1), par_bsk, nullptr, nullptr, Xs, nullptr, role.rb, lane, acc);
        }
        ++xe;
        __syncthreads();
      }
      AMUSE_PROF(2 + layer * 10 + 0);

      // =============== QKV of my head (cross_attention.py:264-266, nn.MultiheadAttention in_proj)
      {
        const float* wt = wp_acquire(s, wp);
        copy_params(par_bqkv, wt + 128 * 96, 96, tid);
        float acc[5][3];
        gemm5<96, 128, 3, KS, false>(Xs, 128, wt, role.rb, role.ks, lane, acc);
        park<96, 3, false, RB>(RED, role, lane, acc);
        __syncthreads();
        wp_release(s, wp, tid);
        if (role.fold) {
          fold<96, 3, false, RB>(RED, role, lane, acc);
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            float* q = QKVs + (role.rb * 5 + i) * kQkvLd;
            // nn.MultiheadAttention scales q (after bias) by head_dim^-0.5 before q.k^T
            q[lane] = (acc[i][0] + par_bqkv[lane]) * 0.17677669529663687f;
            q[32 + lane] = acc[i][1] + par_bqkv[32 + lane];
            q[64 + lane] = acc[i][2] + par_bqkv[64 + lane];
          }
        }
        __syncthreads();
      }
      AMUSE_PROF(2 + layer * 10 + 1);

      // =============== attention: T x T per clip for my head, one warp per clip
      if (warp < S) {
        const float* base = QKVs + warp * T * kQkvLd;
        const int i = (lane < 25) ? lane / 5 : 0, j = lane % 5;
        const bool valid = (lane < 25) && (i < T) && (j < T);
        const int ic = valid ? i : 0, jc = valid ? j : 0;
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        const float4* qv = reinterpret_cast<const float4*>(base + ic * kQkvLd);
        const float4* kv = reinterpret_cast<const float4*>(base + jc * kQkvLd + 32);
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          const float4 a = qv[c], b = kv[c], a2 = qv[c + 1], b2 = kv[c + 1];
          p0 = fmaf(a.x, b.x, p0);
          p1 = fmaf(a.y, b.y, p1);
          p2 = fmaf(a.z, b.z, p2);
          p3 = fmaf(a.w, b.w, p3);
          p0 = fmaf(a2.x, b2.x, p0);
          p1 = fmaf(a2.y, b2.y, p1);
          p2 = fmaf(a2.z, b2.z, p2);
          p3 = fmaf(a2.w, b2.w, p3);
        }
        const float sc = valid ? ((p0 + p1) + (p2 + p3)) : -INFINITY;
        float sj[5];
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) sj[jj] = __shfl_sync(0xffffffffu, sc, i * 5 + jj);
        const float m = fmaxf(fmaxf(fmaxf(sj[0], sj[1]), fmaxf(sj[2], sj[3])), sj[4]);
        const float e = valid ? expf(sc - m) : 0.f;
        float ej[5];
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) ej[jj] = __shfl_sync(0xffffffffu, e, i * 5 + jj);
        const float sum = (ej[0] + ej[1]) + (ej[2] + ej[3]) + ej[4];
        const float pr = valid ? e / sum : 0.f;
        float vj[5];
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) vj[jj] = (jj < T) ? base[jj * kQkvLd + 64 + lane] : 0.f;
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          float o = 0.f;
#pragma unroll
          for (int jj = 0; jj < 5; ++jj) o = fmaf(__shfl_sync(0xffffffffu, pr, ii * 5 + jj), vj[jj], o);
          if (ii < T) Oh[(warp * T + ii) * kOhLd + lane] = o;
        }
      }
      __syncthreads();
      AMUSE_PROF(2 + layer * 10 + 2);

      // =============== out_proj, K-split by head -> DSMEM broadcast of the partial -> sum + LN1
      {
        const float* wt = wp_acquire(s, wp);
        exchange_arm<RB>(s, xe, tid);
        copy_params(par_bo, wt + 32 * 128, 128, tid);
        copy_params(par_ln1, wt + 32 * 128 + 128, 256, tid);
        float acc[5][4];
        gemm5<128, 32, 4, KS, true>(Oh, kOhLd, wt, role.rb, role.ks, lane, acc);
        park<128, 4, true, RB>(RED, role, lane, acc);
        __syncthreads();
        wp_release(s, wp, tid);
        if (role.fold) {
          fold<128, 4, true, RB>(RED, role, lane, acc);
          broadcast_partial(s, xe, rank, role.rb, lane, acc);
          exchange_wait(s, xe);
          AMUSE_PROF(2 + layer * 10 + 3);
          sum_norm_store<true, true>(s.Ps(xe & 1), par_bo, Xs, par_ln1, Xs, nullptr, role.rb, lane, acc);
        }
        ++xe;
        __syncthreads();
      }
      AMUSE_PROF(2 + layer * 10 + 4);

      // =============== FFN1: my 128 hidden units, erf-GELU
      {
        const float* wt = wp_acquire(s, wp);
        copy_params(par_b1, wt + 128 * 128, 128, tid);
        float acc[5][4];
        gemm5<128, 128, 4, KS, true>(Xs, 128, wt, role.rb, role.ks, lane, acc);
        park<128, 4, true, RB>(RED, role, lane, acc);
        __syncthreads();
        wp_release(s, wp, tid);
        if (role.fold) {
          fold<128, 4, true, RB>(RED, role, lane, acc);
          const float4 b = *reinterpret_cast<const float4*>(par_b1 + lane * 4);
#pragma unroll
          for (int i = 0; i < 5; ++i)
            *reinterpret_cast<float4*>(Hs + (role.rb * 5 + i) * 128 + lane * 4) =
                make_float4(gelu_erf(acc[i][0] + b.x), gelu_erf(acc[i][1] + b.y), gelu_erf(acc[i][2] + b.z),
                            gelu_erf(acc[i][3] + b.w));
        }
        __syncthreads();
      }
      AMUSE_PROF(2 + layer * 10 + 5);

      // =============== FFN2, K-split over my 128 hidden units -> broadcast -> sum + LN2
      {
        const float* wt = wp_acquire(s, wp);
        exchange_arm<RB>(s, xe, tid);
        copy_params(par_b2, wt + 128 * 128, 128, tid);
        copy_params(par_ln2, wt + 128 * 128 + 128, 256, tid);
        float acc[5][4];
        gemm5<128, 128, 4, KS, true>(Hs, 128, wt, role.rb, role.ks, lane, acc);
        park<128, 4, true, RB>(RED, role, lane, acc);
        __syncthreads();
        wp_release(s, wp, tid);
        if (role.fold) {
          fold<128, 4, true, RB>(RED, role, lane, acc);
          broadcast_partial(s, xe, rank, role.rb, lane, acc);
          exchange_wait(s, xe);
          AMUSE_PROF(2 + layer * 10 + 6);
          sum_norm_store<true, true>(s.Ps(xe & 1), par_b2, Xs, par_ln2, Xs,
                                     (layer < 4) ? (SK + layer * kRMax * 128) : nullptr, role.rb, lane, acc);
        }
        ++xe;
        __syncthreads();
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
        const float zn = use_rng ? curand_normal(&rng) : noise;
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
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(denoise_loop_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bytes()));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(denoise_loop_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem_bytes()));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (p.S < 1 || p.S > kSMax) return cudaErrorInvalidValue;
  const int n_clusters = (p.B + p.S - 1) / p.S;
  if (p.S == 1)
    denoise_loop_kernel<1><<<dim3(n_clusters * kCluster), dim3(kThreads), smem_bytes(), stream>>>(p);
  else
    denoise_loop_kernel<2><<<dim3(n_clusters * kCluster), dim3(kThreads), smem_bytes(), stream>>>(p);
  return cudaGetLastError();
}

}  // namespace dn
}  // namespace amuse
