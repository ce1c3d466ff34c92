// Interface of the persistent denoise-loop kernel (K1 + K2 of SURVEY.md section 2.2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace amuse {
namespace dn {

constexpr int kCluster = 4;            // CTAs per thread-block cluster = attention heads (one head per CTA)
constexpr int kThreads = 320;        // 8 GEMM warps (row block x K slice) + 2: ten warps = one per activation row in the epilogues
constexpr int kGemmWarps = 8;
constexpr int kSMax = 2;               // clips per cluster
constexpr int kTMax = 5;               // tokens per clip: z, t, con, emo, sty (denoiser.py:174,180)
constexpr int kRMax = kSMax * kTMax;   // activation rows per cluster
constexpr int kTilesPerStep = 40;      // 9 layers x 4 weight tiles + 4 skip-linear tiles
constexpr int kMaxClusters = 33;       // co-resident 4-CTA clusters with ~225 KB smem on B200 (measured,
                                       // scripts/occ_probe.cu: cudaOccupancyMaxActiveClusters = 33)

// Per-CTA-rank weight stream ("blob"): the tiles one CTA consumes during one denoiser
// evaluation, in consumption order, each tile K-major ([k][n_local]) followed by the
// bias / LayerNorm vectors its epilogue needs.  Sizes in floats.
// Every 128-wide tile ends with the same 384-float tail: bias[128] | LayerNorm weight[128] | bias[128]
// (zeros where a stage has no LayerNorm), so one generic stage routine serves all of them.
constexpr int kTileTail = 384;
constexpr int kTileQKV = 128 * 96 + 96;          // in_proj rows of head `rank`: q|k|v 32 each   + bias
constexpr int kTileWO = 32 * 128 + kTileTail;    // out_proj columns of head `rank` (K-split)    + bias + norm1
constexpr int kTileW1 = 128 * 128 + kTileTail;   // linear1 rows [128*rank, +128)                + bias
constexpr int kTileW2 = 128 * 128 + kTileTail;   // linear2 columns [128*rank, +128) (K-split)   + bias + norm2
constexpr int kTileSK = 64 * 128 + kTileTail;    // linear_blocks columns [64*rank, +64) (K-split) + bias
constexpr int kLayerFloats = kTileQKV + kTileWO + kTileW1 + kTileW2;
constexpr int kBlobRankFloats = 9 * kLayerFloats + 4 * kTileSK;
constexpr int kTileMax = kTileW2;

__host__ __device__ inline void tile_info(int i, int& off, int& n) {
  // i in [0, 40): layers 0..4 have 4 tiles, layers 5..8 have 5 (skip-linear first)
  int l, j;
  if (i < 20) {
    l = i >> 2;
    j = i & 3;
    off = l * kLayerFloats;
  } else {
    l = 5 + (i - 20) / 5;
    j = (i - 20) % 5 - 1;   // -1 = skip tile
    off = 5 * kLayerFloats + (l - 5) * (kLayerFloats + kTileSK);
    if (j < 0) {
      n = kTileSK;
      return;
    }
    off += kTileSK;
  }
  const int sz[4] = {kTileQKV, kTileWO, kTileW1, kTileW2};
  for (int q = 0; q < j; ++q) off += sz[q];
  n = sz[j];
}

struct Params {
  const float* blob;         // [kCluster][kBlobRankFloats]
  const float* temb;         // [n_steps][128]   time tokens (a3), batch-invariant
  const float* cond;         // [B][3][128]      condition tokens + their PE rows (a4+a5); first T-2 valid
  const float* pe01;         // [2][128]         query_pos.pe rows 0 and 1
  const float* final_norm;   // [256]            encoder.norm weight | bias
  const float* latents0;     // [B][128]
  const float* step_noise;   // nullable [n_steps][B][128]
  const float* coef;         // [n_steps][5]     sqrt(a), sqrt(1-a), c_x0, c_dir, sigma
  float* latents_out;        // [B][128]
  long long* prof;           // nullable: clock64 stamps of cluster 0 / rank 0 (debug)
  int B, S, T, n_steps;
  int dir_uses_eps;          // 1: x' = c2 x0 + c3 eps (DDIM);  0: x' = c2 x0 + c3 x (DDPM posterior mean)
  int clip;                  // clamp x0 to [-1, 1]
  unsigned long long seed;   // Philox seed when step_noise == nullptr and sigma > 0
  unsigned long long seed_elem_base;   // global index of this launch's first latent element (multi-GPU shards)
  int prof_step;
  int prune_last;            // last layer evaluated for token 0 only (same result; see denoise_loop.cu PRUNE)
  int wide_rows;             // 2-clip clusters: every GEMM warp computes all 10 rows over 1/8 of K (see denoise_loop.cu WIDE)
};

size_t smem_bytes();
cudaError_t launch(const Params& p, cudaStream_t stream);

}  // namespace dn
}  // namespace amuse
