#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace amuse {

struct CondArgs {
  const float* z[3];      // present condition inputs, reference order con, emo, sty  [B][256]
  const float* wt[3];     // emb_proj_*.1.weight transposed to [256][128]
  const float* bias[3];   // [128]
  const float* pe;        // query_pos.pe as [500][128]
  float* out;             // [B][3][128]
};

cudaError_t launch_time_table(const int* timesteps_dev, int n_steps, const float* freqs, const float* w1t,
                              const float* b1, const float* w2t, const float* b2, float* temb, cudaStream_t st);
cudaError_t launch_cond_tokens(const CondArgs& a, int B, int n_cond, cudaStream_t st);
cudaError_t launch_rot6d(const float* feats, int feat_ld, long long n_frames, float* poses, float* trans,
                         cudaStream_t st);

cudaError_t launch_rot6d_flat(const float* d6, long long n, float* aa, cudaStream_t st);
cudaError_t launch_add_planes(const float* hi, const float* lo, float* out, size_t n, cudaStream_t st);

}  // namespace amuse
