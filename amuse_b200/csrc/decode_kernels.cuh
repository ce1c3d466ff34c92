// K3: MotionPrior.decode (reference vae.py:216-278) -- kernel-level interface.
#pragma once
#include <cuda_runtime.h>

namespace amuse {
namespace dec {

enum Epi { EPI_BIAS = 0, EPI_QKV = 1, EPI_GELU = 2, EPI_RES_LN = 3, EPI_RES_LN_CROSS_LN = 4 };

struct GemmArgs {
  const float* A;       // [M][lda]  (columns 0..127 of K when A2 != nullptr)
  int lda;
  const float* A2;      // optional second K half (skip concat: cat(x, skip), cross_attention.py:116)
  int lda2;
  const float* Wt;      // [K][ldw]  weights stored K-major (transposed at finalize)
  int ldw;
  const float* bias;    // [N]
  float* C;             // [M][ldc]
  int ldc;
  int M, N, K;
  const float* R;       // residual [M][ldr]      (EPI_RES_LN*)
  int ldr;
  const float* ln_g;    // LayerNorm after the residual add
  const float* ln_b;
  const float* cvec;    // [clips][128] collapsed cross-attention vector (EPI_RES_LN_CROSS_LN)
  const float* ln2_g;
  const float* ln2_b;
  int rows_per_clip;
};

cudaError_t launch_gemm(int epi, const GemmArgs& a, cudaStream_t st);

// softmax(q k^T) v for 4 heads x 32 dims over `frames` tokens per clip; q is pre-scaled.
cudaError_t launch_self_attention(const float* qkv /*[M][384]*/, float* out /*[M][128]*/, int clips, int frames,
                                  cudaStream_t st);

cudaError_t launch_self_attention_planes(const float* qkv, float* out_hi, float* out_lo, int clips, int frames,
                                         cudaStream_t st);

// the same on tcgen05 (attn_tc.cu; frames <= 320): one CTA per (clip, head), fp16 hi/lo operands, fp32-class accuracy.
// The default of the tensor-core decode / encode path; AMUSE_ATTN_FFMA=1 (read at amuse_create) keeps the fp32 kernel.
cudaError_t launch_self_attention_tc(const float* qkv, float* out_hi, float* out_lo, int clips, int frames,
                                     cudaStream_t st);

// cvec[l][b] = out_proj_l(W_v,l z_b + b_v,l) + b_o,l for the 9 decoder layers (1-key cross attention).
cudaError_t launch_cross_vectors(const float* z /*[B][128]*/, const float* wv_t /*[9][128][128]*/,
                                 const float* bv /*[9][128]*/, const float* wo_t /*[9][128][128]*/,
                                 const float* bo /*[9][128]*/, float* cvec /*[9][B][128]*/, int B, cudaStream_t st);

// x[b][t][:] = pe[t][:]  (queries = zeros + learned PE, vae.py:221,253)
cudaError_t launch_broadcast_rows(const float* src /*[frames][128]*/, float* dst, int clips, int frames,
                                  cudaStream_t st);

}  // namespace dec
}  // namespace amuse
