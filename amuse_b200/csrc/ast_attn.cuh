// AST self-attention (1214 tokens, 12 heads x 64; reference models/audio/audio_main_new.py:190 ->
// timm 0.4.5 Attention.forward) as a tcgen05 flash-attention kernel with fp32-class (3xTF32) accuracy.
#pragma once
#include <cuda_runtime.h>

namespace amuse {
namespace attn {

constexpr int kTok = 1214;    // cls + dist + 12 x 101 patches
constexpr int kTokP = 1280;   // per-(clip, head) row pitch of the operand planes: 5 query blocks of 256
constexpr int kHeads = 12;
constexpr int kHD = 64;

struct AttnArgs {
  // operand planes written by the qkv GEMM epilogue (tc::EPI_QKV_HEADS); pad rows / columns are zero
  const float *q_hi, *q_lo;     // [nb][12][1280][64], q pre-scaled by 64^-0.5 * log2(e)
  const float *k_hi, *k_lo;     // [nb][12][1280][64]
  const float *vt_hi, *vt_lo;   // [nb][12][64][1280]
  float *o_hi, *o_lo;           // [nb*1214][768] planes of softmax(q k^T) v, head h in columns [64h, 64h+64)
  int nb;                       // clips
};

// elements of one q / k / v^T plane for nb clips
inline size_t plane_elems(int nb) { return static_cast<size_t>(nb) * kHeads * kTokP * kHD; }

cudaError_t attention(const AttnArgs& a, cudaStream_t st);

// debug: enable != 0 arms a 64-slot clock64 timeline of CTA (0,0,0) for the next launches; enable == 0 disarms and
// copies the stamps out (layout in ast_attn.cu)
cudaError_t debug_profile(int enable, long long* host_out, int n);

}  // namespace attn
}  // namespace amuse
