// K1+K2: the whole N-step denoising loop of PretrainedLPDM_v1.diffusion_backward
// (reference infer_ldm.py:142-161) as ONE persistent launch.
//
// What one step computes (reference denoiser.py:135-204, cross_attention.py:41-64,259-272):
//   x = [z | time token | con | emo | sty] + learned PE        (<= 5 tokens x 128 per clip)
//   9 post-LN encoder layers with U-Net skips, final LayerNorm, eps = token 0
//   scheduler update of z (DDIM eta / DDPM ancestral), clamp(x0) optional
//
// B200 mapping.  The loop is a chain of ~40 dependent small GEMMs per step with M = 5 rows
// per clip: it is latency-bound, and each evaluation needs all 8.8 MB of fp32 weights.  A
// thread-block cluster of 8 CTAs (8 SMs) owns up to 4 clips for all steps.  Weights are
// split 8 ways across the cluster so each SM streams only 1/8 of them per step (from L2,
// via the TMA bulk-copy engine into a 2-deep shared-memory ring that runs two tiles ahead
// of the math); activations (20 x 128 floats) are replicated in every CTA's shared memory
// and the partial results are exchanged through distributed shared memory:
//   QKV      N-split by head (CTA pair = one head, the pair splits the clips)   -> local
//   attn     5x5 per (clip, head), one warp                                      -> local
//   out_proj K-split by head  -> reduce-scatter (DSMEM) -> all-gather (DSMEM) -> LN1
//   FFN1     N-split (64 hidden units per CTA)                                   -> local
//   FFN2     K-split (same 64 units) -> reduce-scatter -> all-gather -> LN2
// i.e. 4 hardware cluster barriers per layer and no global-memory traffic for activations.
// All arithmetic is fp32 FFMA: 50..1000 recurrent steps with clamp() do not survive bf16,
// and at M = 20 rows the tensor pipe would be operand-bandwidth bound.
#include "denoise_loop.cuh"

#include <curand_kernel.h>

#include "common.cuh"

namespace amuse {
namespace dn {

namespace {

constexpr int kWBufFloats = 12416;   // >= kTileQKV, multiple of 32 floats (128 B)
constexpr int kQkvLd = 97;           // row stride of the local q|k|v buffer (bank-conflict-free 5x5 dots)
constexpr int kOhLd = 36;            // row stride of the local attention output (16-B aligned rows)
constexpr int kRedFloats = 3 * 10 * 128;

// ---------------------------------------------------------------- shared-memory carve-up
struct Smem {
  float* wbuf[2];    // weight ring
  float* Xs;         // [20][128] residual stream, replicated in every CTA
  float* SK;         // [4][20][128] saved skips of the input blocks
  float* Ys;         // [20][128] all-gathered pre-LayerNorm sums (remote-written)
  float* Zs;         // [20][128] all-gathered skip-linear output (remote-written)
  float* Ps;         // [8][20][16] reduce-scatter receive buffer (remote-written)
  float* QKVs;       // [10][97] q|k|v of my head for my half of the clips
  float* Oh;         // [10][36] attention output of my head, my clips
  float* Hs;         // [20][64] my 64 hidden units
  float* RED;        // K-split partial sums
  float* Cs;         // [4][3][128] condition tokens (+PE)
  float* pe01;       // [2][128]
  float* zs;         // [4][128] current latents
  float* Es;         // [4][128] eps
  float* tembs;      // [128]
  float* fn;         // [256] final norm
  float* par_bqkv;   // [96]
  float* par_bo;     // [128]
  float* par_ln1;    // [256]
  float* par_b1;     // [64]
  float* par_b2;     // [128]
  float* par_ln2;    // [256]
  float* par_bsk;    // [16]
  uint64_t* bar;     // [2]
};

constexpr int kSmemFloats = 2 * kWBufFloats + 2560 + 4 * 2560 + 2560 + 2560 + 2560 + 10 * kQkvLd + 6 /*pad*/ +
                            10 * kOhLd + 20 * 64 + kRedFloats + 4 * 3 * 128 + 256 + 512 + 512 + 128 + 256 + 96 +
                            128 + 256 + 64 + 128 + 256 + 16 + 16 /*bars + pad*/;

__device__ __forceinline__ void carve(float* base, Smem& s) {
  float* p = base;
  auto take = [&](int n) {
    float* r = p;
    p += n;
    return r;
  };
  s.wbuf[0] = take(kWBufFloats);
  s.wbuf[1] = take(kWBufFloats);
  s.Xs = take(2560);
  s.SK = take(4 * 2560);
  s.Ys = take(2560);
  s.Zs = take(2560);
  s.Ps = take(2560);
  s.QKVs = take(10 * kQkvLd + 6);   // 976: keeps the following buffers 16-B aligned
  s.Oh = take(10 * kOhLd);
  s.Hs = take(20 * 64);
  s.RED = take(kRedFloats);
  s.Cs = take(4 * 3 * 128);
  s.pe01 = take(256);
  s.zs = take(512);
  s.Es = take(512);
  s.tembs = take(128);
  s.fn = take(256);
  s.par_bqkv = take(96);
  s.par_bo = take(128);
  s.par_ln1 = take(256);
  s.par_b1 = take(64);
  s.par_b2 = take(128);
  s.par_ln2 = take(256);
  s.par_bsk = take(16);
  s.bar = reinterpret_cast<uint64_t*>(take(16));
}

// ---------------------------------------------------------------- weight pipeline
struct WPipe {
  const float* blob;   // this rank's stream
  uint32_t g;          // sequence number of the tile about to be consumed
  uint32_t total;      // n_steps * 40
};

__device__ __forceinline__ void wp_issue(const Smem& s, const WPipe& w, uint32_t n) {
  int off, cnt;
  tile_info(static_cast<int>(n % kTilesPerStep), off, cnt);
  uint64_t* bar = &s.bar[n & 1];
  mbar_arrive_expect_tx(bar, static_cast<uint32_t>(cnt) * 4u);
  bulk_g2s(s.wbuf[n & 1], w.blob + off, static_cast<uint32_t>(cnt) * 4u, bar);
}
__device__ __forceinline__ const float* wp_acquire(const Smem& s, const WPipe& w) {
  mbar_wait(&s.bar[w.g & 1], (w.g >> 1) & 1);
  return s.wbuf[w.g & 1];
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

// ---------------------------------------------------------------- 5-row FFMA micro-kernel
// acc[i][j] = sum_k A[rb*5+i][k0..k0+KDIM/KSPLIT) * Wt[k][col(lane,j)]
// The 8 warps are arranged RBLK x KSPLIT; the 32 lanes own the columns:
//   CONTIG (TC == 4): lane owns columns 4*lane .. 4*lane+3          (NCOL = 128)
//   strided         : lane owns columns lane + 32*j, j < TC          (NCOL = 32*TC)
// A-row loads are warp-uniform 128-bit broadcasts; weight loads are conflict-free.
template <int NCOL, int KDIM, int TC, int KSPLIT, bool CONTIG>
__device__ __forceinline__ void gemm5(const float* __restrict__ A, int lda, int nrows,
                                      const float* __restrict__ Wt, int warp, int lane, float (&acc)[5][TC]) {
  constexpr int RBLK = 8 / KSPLIT;
  constexpr int KPER = KDIM / KSPLIT;
  static_assert(KPER % 4 == 0, "K slice must be a multiple of 4");
  const int rb = warp % RBLK, ks = warp / RBLK;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;
  if (rb * 5 >= nrows) return;
  const float* arow[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) arow[i] = A + min(rb * 5 + i, nrows - 1) * lda + ks * KPER;
  const float* w = Wt + (ks * KPER) * NCOL + (CONTIG ? lane * 4 : lane);
#pragma unroll 2
  for (int kk = 0; kk < KPER; kk += 4) {
    float4 a[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) a[i] = *reinterpret_cast<const float4*>(arow[i] + kk);
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      float wv[TC];
      if (CONTIG) {
        const float4 t = *reinterpret_cast<const float4*>(w + (kk + k4) * NCOL);
        wv[0] = t.x;
        wv[1] = t.y;
        wv[2] = t.z;
        wv[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < TC; ++j) wv[j] = w[(kk + k4) * NCOL + 32 * j];
      }
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float av = (k4 == 0) ? a[i].x : (k4 == 1) ? a[i].y : (k4 == 2) ? a[i].z : a[i].w;
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(av, wv[j], acc[i][j]);
      }
    }
  }
}

// K-split reduction through shared memory.  Warps with ks > 0 park their partials; after the
// barrier the ks == 0 warps fold them in.  The barrier is also the point where the weight
// tile is no longer needed.
template <int NCOL, int TC, int KSPLIT, bool CONTIG>
__device__ __forceinline__ void park_partials(float* RED, int warp, int lane, const float (&acc)[5][TC]) {
  constexpr int RBLK = 8 / KSPLIT;
  const int rb = warp % RBLK, ks = warp / RBLK;
  if (ks == 0) return;
  float* dst = RED + ((ks - 1) * RBLK * 5 + rb * 5) * NCOL;
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
template <int NCOL, int TC, int KSPLIT, bool CONTIG>
__device__ __forceinline__ void fold_partials(const float* RED, int warp, int lane, float (&acc)[5][TC]) {
  constexpr int RBLK = 8 / KSPLIT;
  const int rb = warp % RBLK;
#pragma unroll
  for (int q = 0; q < KSPLIT - 1; ++q) {
    const float* src = RED + (q * RBLK * 5 + rb * 5) * NCOL;
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

__device__ __forceinline__ void copy_params(float* dst, const float* src, int n, int tid) {
  for (int i = tid; i < n; i += kThreads) dst[i] = src[i];
}

#define AMUSE_PROF(slot)                                                   \
  do {                                                                     \
    if (do_prof && tid == 0) p.prof[(slot)] = clock64();                   \
  } while (0)

}  // namespace

// ================================================================= the kernel
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    denoise_loop_kernel(const Params p) {
  extern __shared__ __align__(128) float smem_raw[];
  Smem s;
  carve(smem_raw, s);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int cid = static_cast<int>(cluster_id_x());
  const int head = rank >> 1, half = rank & 1;
  const int T = p.T;
  const int s_base = cid * p.S;
  const int S = min(p.S, p.B - s_base);            // clips of this cluster (>= 1 by grid construction)
  const int R = S * T;                             // activation rows
  const int S0 = (S + 1) >> 1;                     // clips handled by the even CTA of each head pair
  const int my_s0 = half ? S0 : 0;
  const int my_ns = half ? (S - S0) : S0;
  const int row0 = my_s0 * T;                      // first row of my half
  const int RH = my_ns * T;                        // rows of my half (<= 10)

  // ---- one-time setup
  for (int i = tid; i < kSmemFloats - 2 * kWBufFloats - 16; i += kThreads) s.Xs[i] = 0.f;   // all activation buffers
  if (tid == 0) {
    mbar_init(&s.bar[0], 1);
    mbar_init(&s.bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  for (int i = tid; i < S * 3 * 128; i += kThreads) s.Cs[i] = p.cond[static_cast<size_t>(s_base) * 384 + i];
  for (int i = tid; i < 256; i += kThreads) {
    s.pe01[i] = p.pe01[i];
    s.fn[i] = p.final_norm[i];
  }
  for (int i = tid; i < S * 128; i += kThreads) s.zs[i] = p.latents0[static_cast<size_t>(s_base) * 128 + i];

  WPipe wp;
  wp.blob = p.blob + static_cast<size_t>(rank) * kBlobRankFloats;
  wp.g = 0;
  wp.total = static_cast<uint32_t>(p.n_steps) * kTilesPerStep;
  if (tid == 0) {
    wp_issue(s, wp, 0);
    wp_issue(s, wp, 1);
  }
  __syncthreads();
  cluster_sync_all();   // every CTA of the cluster is resident before any DSMEM store

  // per-thread element ownership for the scheduler update: idx = tid, tid + 256 (< S*128)
  // Philox stream per latent element (subsequence = global element index), so the noise a clip
  // sees does not depend on how clips are packed into clusters or sharded over GPUs.
  curandStatePhilox4_32_10_t rng[2];
  const bool use_rng = (p.step_noise == nullptr);
  if (use_rng) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
      curand_init(p.seed, p.seed_elem_base + static_cast<unsigned long long>(s_base) * 128ull + tid + q * kThreads,
                  0, &rng[q]);
  }

  // software prefetch (one step ahead) of the tiny per-step global reads
  float temb_next = (tid < 128) ? __ldg(p.temb + tid) : 0.f;
  float coef_next[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) coef_next[q] = __ldg(p.coef + q);
  float noise_next[2] = {0.f, 0.f};
  if (!use_rng) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * kThreads;
      if (idx < S * 128) noise_next[q] = __ldg(p.step_noise + static_cast<size_t>(s_base) * 128 + idx);
    }
  }

  for (int step = 0; step < p.n_steps; ++step) {
    const bool do_prof = (p.prof != nullptr) && cid == 0 && rank == 0 && step == p.prof_step;
    AMUSE_PROF(0);
    float coef[5], noise[2];
#pragma unroll
    for (int q = 0; q < 5; ++q) coef[q] = coef_next[q];
    noise[0] = noise_next[0];
    noise[1] = noise_next[1];
    if (tid < 128) s.tembs[tid] = temb_next;
    if (step + 1 < p.n_steps) {
      if (tid < 128) temb_next = __ldg(p.temb + static_cast<size_t>(step + 1) * 128 + tid);
#pragma unroll
      for (int q = 0; q < 5; ++q) coef_next[q] = __ldg(p.coef + static_cast<size_t>(step + 1) * 5 + q);
      if (!use_rng) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int idx = tid + q * kThreads;
          if (idx < S * 128)
            noise_next[q] = __ldg(p.step_noise + (static_cast<size_t>(step + 1) * p.B + s_base) * 128 + idx);
        }
      }
    }
    __syncthreads();

    // ---- assemble the token rows (denoiser.py:174-181 + position_encoding.py:156)
    for (int idx = tid; idx < R * 32; idx += kThreads) {
      const int r = idx >> 5, c4 = (idx & 31) * 4;
      const int sl = r / T, tok = r - sl * T;
      float4 v;
      if (tok == 0) {
        const float4 z = *reinterpret_cast<const float4*>(s.zs + sl * 128 + c4);
        const float4 e = *reinterpret_cast<const float4*>(s.pe01 + c4);
        v = make_float4(z.x + e.x, z.y + e.y, z.z + e.z, z.w + e.w);
      } else if (tok == 1) {
        const float4 z = *reinterpret_cast<const float4*>(s.tembs + c4);
        const float4 e = *reinterpret_cast<const float4*>(s.pe01 + 128 + c4);
        v = make_float4(z.x + e.x, z.y + e.y, z.z + e.z, z.w + e.w);
      } else {
        v = *reinterpret_cast<const float4*>(s.Cs + (sl * 3 + tok - 2) * 128 + c4);
      }
      *reinterpret_cast<float4*>(s.Xs + r * 128 + c4) = v;
    }
    __syncthreads();
    AMUSE_PROF(1);

    for (int layer = 0; layer < kLayers; ++layer) {
      // =============== skip fusion of the output blocks: x = Linear(256->128)(cat(x, skip))
      if (layer >= 5) {
        const float* wt = wp_acquire(s, wp);
        const float* skip = s.SK + (8 - layer) * 2560;     // xs.pop(): block 5 takes skip 3, ... block 8 takes skip 0
        copy_params(s.par_bsk, wt + 256 * 16, 16, tid);
        // narrow N-split GEMM: 16 columns per CTA, K = 256 split over 4 warp groups
        const int col = lane & 15, rb = (warp & 1) * 2 + (lane >> 4), ks = warp >> 1;
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (rb * 5 < R) {
          const float* src = (ks < 2) ? s.Xs : skip;
          const int kofs = (ks & 1) * 64;
          const float* wk = wt + (ks * 64) * 16 + col;
          const float* ar[5];
#pragma unroll
          for (int i = 0; i < 5; ++i) ar[i] = src + min(rb * 5 + i, R - 1) * 128 + kofs;
#pragma unroll 2
          for (int kk = 0; kk < 64; kk += 4) {
            float4 a[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) a[i] = *reinterpret_cast<const float4*>(ar[i] + kk);
            const float w0 = wk[(kk + 0) * 16], w1 = wk[(kk + 1) * 16], w2 = wk[(kk + 2) * 16], w3 = wk[(kk + 3) * 16];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              acc[i] = fmaf(a[i].x, w0, acc[i]);
              acc[i] = fmaf(a[i].y, w1, acc[i]);
              acc[i] = fmaf(a[i].z, w2, acc[i]);
              acc[i] = fmaf(a[i].w, w3, acc[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < 5; ++i) s.RED[(ks * 20 + rb * 5 + i) * 16 + col] = acc[i];
        }
        __syncthreads();
        wp_release(s, wp, tid);
        if (tid < R * 4) {
          const int r = tid >> 2, c4 = (tid & 3) * 4;
          float4 v = *reinterpret_cast<const float4*>(s.par_bsk + c4);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(s.RED + (q * 20 + r) * 16 + c4);
            v.x += t.x;
            v.y += t.y;
            v.z += t.z;
            v.w += t.w;
          }
          float* dst = s.Zs + r * 128 + rank * 16 + c4;
#pragma unroll
          for (int j = 0; j < kCluster; ++j) st_cluster_f4(map_to_rank(dst, j), v);
        }
        cluster_sync_all();
        for (int idx = tid; idx < R * 32; idx += kThreads)
          reinterpret_cast<float4*>(s.Xs)[idx] = reinterpret_cast<const float4*>(s.Zs)[idx];
        __syncthreads();
      }
      AMUSE_PROF(2 + layer * 10 + 0);

      // =============== QKV of my head for my half of the clips (cross_attention.py:264-266)
      {
        const float* wt = wp_acquire(s, wp);
        copy_params(s.par_bqkv, wt + 128 * 96, 96, tid);
        float acc[5][3];
        gemm5<96, 128, 3, 4, false>(s.Xs + row0 * 128, 128, RH, wt, warp, lane, acc);
        park_partials<96, 3, 4, false>(s.RED, warp, lane, acc);
        __syncthreads();
        wp_release(s, wp, tid);
        const int rb = warp % 2;
        if (warp < 2 && rb * 5 < RH) {
          fold_partials<96, 3, 4, false>(s.RED, warp, lane, acc);
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const int r = rb * 5 + i;
            if (r < RH) {
              // nn.MultiheadAttention scales q (after bias) by head_dim^-0.5 before q.k^T
              s.QKVs[r * kQkvLd + lane] = (acc[i][0] + s.par_bqkv[lane]) * 0.17677669529663687f;
              s.QKVs[r * kQkvLd + 32 + lane] = acc[i][1] + s.par_bqkv[32 + lane];
              s.QKVs[r * kQkvLd + 64 + lane] = acc[i][2] + s.par_bqkv[64 + lane];
            }
          }
        }
        __syncthreads();
      }
      AMUSE_PROF(2 + layer * 10 + 1);

      // =============== attention: T x T per (clip, head), one warp per clip
      if (warp < my_ns) {
        const float* base = s.QKVs + warp * T * kQkvLd;
        const int i = (lane < 25) ? lane / 5 : 0, j = lane % 5;
        const bool valid = (lane < 25) && (i < T) && (j < T);
        float sc = -INFINITY;
        if (valid) {
          sc = 0.f;
#pragma unroll
          for (int d = 0; d < 32; ++d) sc = fmaf(base[i * kQkvLd + d], base[j * kQkvLd + 32 + d], sc);
        }
        float m = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) m = fmaxf(m, __shfl_sync(0xffffffffu, sc, i * 5 + jj));
        const float e = valid ? expf(sc - m) : 0.f;
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) sum += __shfl_sync(0xffffffffu, e, i * 5 + jj);
        const float pr = valid ? e / sum : 0.f;
        for (int ii = 0; ii < T; ++ii) {
          float o = 0.f;
          for (int jj = 0; jj < T; ++jj)
            o = fmaf(__shfl_sync(0xffffffffu, pr, ii * 5 + jj), base[jj * kQkvLd + 64 + lane], o);
          s.Oh[(warp * T + ii) * kOhLd + lane] = o;
        }
      }
      __syncthreads();
      AMUSE_PROF(2 + layer * 10 + 2);

      // =============== out_proj, K-split by head: partial[rows of my half][128] -> reduce-scatter
      {
        const float* wt = wp_acquire(s, wp);
        copy_params(s.par_bo, wt + 32 * 128, 128, tid);
        copy_params(s.par_ln1, wt + 32 * 128 + 128, 256, tid);
        float acc[5][4];
        gemm5<128, 32, 4, 4, true>(s.Oh, kOhLd, RH, wt, warp, lane, acc);
        park_partials<128, 4, 4, true>(s.RED, warp, lane, acc);
        __syncthreads();
        wp_release(s, wp, tid);
        const int rb = warp % 2;
        if (warp < 2 && rb * 5 < RH) {
          fold_partials<128, 4, 4, true>(s.RED, warp, lane, acc);
          const uint32_t dst_rank = lane >> 2;
          const int cofs = (lane & 3) * 4;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const int r = rb * 5 + i;
            if (r < RH)
              st_cluster_f4(map_to_rank(s.Ps + (head * kRMax + row0 + r) * 16 + cofs, dst_rank),
                            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
          }
        }
      }
      cluster_sync_all();
      AMUSE_PROF(2 + layer * 10 + 3);
      // reduce my 16 columns over the 4 heads, add bias + residual, all-gather
      if (tid < R * 4) {
        const int r = tid >> 2, c4 = (tid & 3) * 4, c = rank * 16 + c4;
        float4 v = *reinterpret_cast<const float4*>(s.par_bo + c);
        const float4 x = *reinterpret_cast<const float4*>(s.Xs + r * 128 + c);
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          const float4 t = *reinterpret_cast<const float4*>(s.Ps + (hh * kRMax + r) * 16 + c4);
          v.x += t.x;
          v.y += t.y;
          v.z += t.z;
          v.w += t.w;
        }
        v.x += x.x;
        v.y += x.y;
        v.z += x.z;
        v.w += x.w;
        float* dst = s.Ys + r * 128 + c;
#pragma unroll
        for (int j = 0; j < kCluster; ++j) st_cluster_f4(map_to_rank(dst, j), v);
      }
      cluster_sync_all();
      AMUSE_PROF(2 + layer * 10 + 4);
      // LayerNorm 1 (replicated): warp per row
      for (int r = warp; r < R; r += 8) {
        const float4 v = *reinterpret_cast<const float4*>(s.Ys + r * 128 + lane * 4);
        const float4 g = *reinterpret_cast<const float4*>(s.par_ln1 + lane * 4);
        const float4 b = *reinterpret_cast<const float4*>(s.par_ln1 + 128 + lane * 4);
        *reinterpret_cast<float4*>(s.Xs + r * 128 + lane * 4) = warp_layernorm128(v, g, b);
      }
      __syncthreads();
      AMUSE_PROF(2 + layer * 10 + 5);

      // =============== FFN1: my 64 hidden units, erf-GELU
      {
        const float* wt = wp_acquire(s, wp);
        copy_params(s.par_b1, wt + 128 * 64, 64, tid);
        float acc[5][2];
        gemm5<64, 128, 2, 2, false>(s.Xs, 128, R, wt, warp, lane, acc);
        park_partials<64, 2, 2, false>(s.RED, warp, lane, acc);
        __syncthreads();
        wp_release(s, wp, tid);
        const int rb = warp % 4;
        if (warp < 4 && rb * 5 < R) {
          fold_partials<64, 2, 2, false>(s.RED, warp, lane, acc);
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const int r = rb * 5 + i;
            if (r < R) {
              s.Hs[r * 64 + lane] = gelu_erf(acc[i][0] + s.par_b1[lane]);
              s.Hs[r * 64 + 32 + lane] = gelu_erf(acc[i][1] + s.par_b1[32 + lane]);
            }
          }
        }
        __syncthreads();
      }
      AMUSE_PROF(2 + layer * 10 + 6);

      // =============== FFN2, K-split over my 64 hidden units -> reduce-scatter
      {
        const float* wt = wp_acquire(s, wp);
        copy_params(s.par_b2, wt + 64 * 128, 128, tid);
        copy_params(s.par_ln2, wt + 64 * 128 + 128, 256, tid);
        float acc[5][4];
        gemm5<128, 64, 4, 2, true>(s.Hs, 64, R, wt, warp, lane, acc);
        park_partials<128, 4, 2, true>(s.RED, warp, lane, acc);
        __syncthreads();
        wp_release(s, wp, tid);
        const int rb = warp % 4;
        if (warp < 4 && rb * 5 < R) {
          fold_partials<128, 4, 2, true>(s.RED, warp, lane, acc);
          const uint32_t dst_rank = lane >> 2;
          const int cofs = (lane & 3) * 4;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const int r = rb * 5 + i;
            if (r < R)
              st_cluster_f4(map_to_rank(s.Ps + (static_cast<int>(rank) * kRMax + r) * 16 + cofs, dst_rank),
                            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
          }
        }
      }
      cluster_sync_all();
      AMUSE_PROF(2 + layer * 10 + 7);
      if (tid < R * 4) {
        const int r = tid >> 2, c4 = (tid & 3) * 4, c = rank * 16 + c4;
        float4 v = *reinterpret_cast<const float4*>(s.par_b2 + c);
        const float4 x = *reinterpret_cast<const float4*>(s.Xs + r * 128 + c);
#pragma unroll
        for (int q = 0; q < kCluster; ++q) {
          const float4 t = *reinterpret_cast<const float4*>(s.Ps + (q * kRMax + r) * 16 + c4);
          v.x += t.x;
          v.y += t.y;
          v.z += t.z;
          v.w += t.w;
        }
        v.x += x.x;
        v.y += x.y;
        v.z += x.z;
        v.w += x.w;
        float* dst = s.Ys + r * 128 + c;
#pragma unroll
        for (int j = 0; j < kCluster; ++j) st_cluster_f4(map_to_rank(dst, j), v);
      }
      cluster_sync_all();
      AMUSE_PROF(2 + layer * 10 + 8);
      // LayerNorm 2 (replicated); input blocks also push the result on the skip stack
      for (int r = warp; r < R; r += 8) {
        const float4 v = *reinterpret_cast<const float4*>(s.Ys + r * 128 + lane * 4);
        const float4 g = *reinterpret_cast<const float4*>(s.par_ln2 + lane * 4);
        const float4 b = *reinterpret_cast<const float4*>(s.par_ln2 + 128 + lane * 4);
        const float4 y = warp_layernorm128(v, g, b);
        *reinterpret_cast<float4*>(s.Xs + r * 128 + lane * 4) = y;
        if (layer < 4) *reinterpret_cast<float4*>(s.SK + layer * 2560 + r * 128 + lane * 4) = y;
      }
      __syncthreads();
      AMUSE_PROF(2 + layer * 10 + 9);
    }   // layers

    // ---- encoder.norm on token 0 of every clip -> eps (cross_attention.py:62-63, denoiser.py:188)
    if (warp < S) {
      const float4 v = *reinterpret_cast<const float4*>(s.Xs + (warp * T) * 128 + lane * 4);
      const float4 g = *reinterpret_cast<const float4*>(s.fn + lane * 4);
      const float4 b = *reinterpret_cast<const float4*>(s.fn + 128 + lane * 4);
      *reinterpret_cast<float4*>(s.Es + warp * 128 + lane * 4) = warp_layernorm128(v, g, b);
    }
    __syncthreads();
    // ---- scheduler step (K2), replicated in every CTA; op order of diffusers' step():
    //      x0 = (x - sqrt(1-a) e) / sqrt(a); clamp; x' = c2 x0 + c3 (e | x) + sigma z
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * kThreads;
      if (idx < S * 128) {
        const float x = s.zs[idx], e = s.Es[idx];
        float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(coef[1], e)), coef[0]);
        if (p.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
        float out = __fadd_rn(__fmul_rn(coef[2], x0), __fmul_rn(coef[3], p.dir_uses_eps ? e : x));
        if (coef[4] != 0.f) {
          const float zn = use_rng ? curand_normal(&rng[q]) : noise[q];
          out = __fadd_rn(out, __fmul_rn(coef[4], zn));
        }
        s.zs[idx] = out;
      }
    }
    __syncthreads();
    AMUSE_PROF(2 + kLayers * 10);
  }   // steps

  if (rank == 0)
    for (int i = tid; i < S * 128; i += kThreads) p.latents_out[static_cast<size_t>(s_base) * 128 + i] = s.zs[i];
  cluster_sync_all();   // nobody leaves while a peer could still address its shared memory
}

size_t smem_bytes() { return static_cast<size_t>(kSmemFloats) * sizeof(float); }

cudaError_t launch(const Params& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(denoise_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bytes()));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int n_clusters = (p.B + p.S - 1) / p.S;
  denoise_loop_kernel<<<dim3(n_clusters * kCluster), dim3(kThreads), smem_bytes(), stream>>>(p);
  return cudaGetLastError();
}

}  // namespace dn
}  // namespace amuse
