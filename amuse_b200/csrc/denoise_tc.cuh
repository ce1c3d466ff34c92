// Interface of the tensor-core denoise-loop kernel (K1 + K2 of SURVEY.md section 2.2): the whole N-step loop of
// PretrainedLPDM_v1.diffusion_backward (reference infer_ldm.py:142-161) in one persistent launch, every GEMM on
// tcgen05 with the weights as the TMEM-resident A operand.  See denoise_tc.cu for the design.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace amuse {
namespace dn2 {

constexpr int kCluster = 2;            // CTAs per cluster; one clip per cluster
constexpr int kVirt = 2;               // weight "ranks" per CTA: CTA r owns attention heads / hidden slices 2r and 2r + 1
constexpr int kRanks = kCluster * kVirt;   // the layers split 4 ways (one attention head per rank)
constexpr int kChainWarps = 8;         // epilogue warps: (TMEM lane quadrant q, tile / row group t)
constexpr int kProdWarps = 8;          // weight producers: global (L2) -> registers -> tcgen05.st
constexpr int kThreads = (kChainWarps + kProdWarps + 1) * 32;   // + the MMA issuer warp
constexpr int kTMax = 5;               // tokens per clip: z, t, con, emo, sty (denoiser.py:174,180)
constexpr int kTilesPerStep = 40;      // 9 layers x 4 weight tiles + 4 skip-linear tiles

// ---- per-rank weight stream.  A tile is the A operand of one GEMM stage: M = 128 output features (TMEM lanes) x K
// input features as fp16 hi / lo' planes (x = hi + lo' / 2048), two consecutive k per 32-bit TMEM column:
//   columns [0, K/2)  hi pairs (k = 2j, 2j+1),   columns [K/2, K)  lo' pairs.
// The stream stores it in the order the producers read it: units of 16 columns; inside a unit
//   [quadrant q = feature / 32][i = 0..3][lane = feature % 32] x uint4  (word w of vector i = column 16u + 4i + w)
// so every LDG.128 of a producer warp reads 512 contiguous bytes.  The q|k|v tile has 96 features: 3 quadrants.
enum TileKind { kQKV = 0, kWO = 1, kW1 = 2, kW2 = 3, kSK = 4 };
__host__ __device__ constexpr int tile_K(int kind) { return kind == kWO ? 32 : kind == kSK ? 64 : 128; }
__host__ __device__ constexpr int tile_quads(int kind) { return kind == kQKV ? 3 : 4; }
__host__ __device__ constexpr int tile_units(int kind) { return tile_K(kind) / 16; }
__host__ __device__ constexpr int tile_vec4(int kind) { return tile_units(kind) * tile_quads(kind) * 4 * 32; }
constexpr int kLayerVec4 = tile_vec4(kQKV) + tile_vec4(kWO) + tile_vec4(kW1) + tile_vec4(kW2);   // 12288 (192 KB)
constexpr int kRankVec4 = 9 * kLayerVec4 + 4 * tile_vec4(kSK);                                    // 118784 (1.86 MB)
// bias | LayerNorm weight | LayerNorm bias of tile t (zeros where a stage has none): [40][3][128] floats per rank
constexpr int kRankVecFloats = kTilesPerStep * 3 * 128;

// tile i in [0, 40) of one denoiser evaluation: layers 0..4 have 4 tiles, layers 5..8 have 5 (skip-linear first)
__host__ __device__ inline void tile_info(int i, int& kind, int& off_vec4) {
  int l, j;
  if (i < 20) {
    l = i >> 2;
    j = i & 3;
    off_vec4 = l * kLayerVec4;
  } else {
    l = 5 + (i - 20) / 5;
    j = (i - 20) % 5 - 1;   // -1 = skip tile
    off_vec4 = 5 * kLayerVec4 + (l - 5) * (kLayerVec4 + tile_vec4(kSK));
    if (j < 0) {
      kind = kSK;
      return;
    }
    off_vec4 += tile_vec4(kSK);
  }
  for (int q = 0; q < j; ++q) off_vec4 += tile_vec4(q);
  kind = j;
}

struct Params {
  const uint4* blob;         // [kCluster][40 stages][kVirt] tiles: each CTA's two rank streams interleaved in consumption order
  const float* vecs;         // [kRanks][kRankVecFloats]
  const float* temb;         // [n_steps][128]   time tokens (a3), batch-invariant
  const float* cond;         // [B][3][128]      condition tokens + their PE rows (a4+a5); first T-2 valid
  const float* pe01;         // [2][128]         query_pos.pe rows 0 and 1
  const float* final_norm;   // [256]            encoder.norm weight | bias
  const float* latents0;     // [B][128]
  const float* step_noise;   // nullable [n_steps][B][128]
  const float* coef;         // [n_steps][5]     sqrt(a), sqrt(1-a), c_x0, c_dir, sigma
  float* latents_out;        // [B][128]
  long long* prof;           // nullable: clock64 stamps of cluster 0 / CTA 0 / thread 0 (debug)
  int* status;               // nullable: set to 1 if a bounded wait expired (the kernel then traps)
  int B, T, n_steps;
  int dir_uses_eps;          // 1: x' = c2 x0 + c3 eps (DDIM);  0: x' = c2 x0 + c3 x (DDPM posterior mean)
  int clip;                  // clamp x0 to [-1, 1]
  unsigned long long seed;   // Philox key when step_noise == nullptr and sigma > 0
  unsigned long long seed_elem_base;   // global index of this launch's first latent element (multi-GPU shards)
  int prof_step;
  int prune_last;            // last layer evaluated for token 0 only (same result)
  int debug_flags;           // timing experiments only (results are garbage): 1 = producers idle (AMUSE_DN2_DEBUG=1);
                             // 2 = no global loads, 4 = no tcgen05.st only in a -DAMUSE_DN2_PRODUCER_DEBUG build
};

size_t smem_bytes();
cudaError_t launch(const Params& p, cudaStream_t stream);

// host side of the fp16 split the kernel uses: hi = fp16(x) (round to nearest, saturating), lo' = fp16((x - hi) * 2048)
void split_fp16(float x, uint16_t& hi, uint16_t& lo);

}  // namespace dn2
}  // namespace amuse
