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

// ---- MotionPrior.encode side (vae.py:154-214; infer_ldm.py:454-462)
// poses [n_frames][55][3] axis-angle + trans [n_frames][3] -> feats [n_frames][333] (55 x 6D | trans)
cudaError_t launch_motion_to_feats(const float* poses, const float* trans, long long n_frames, float* feats,
                                   cudaStream_t st);
// feats [rows][333] -> TF32 hi/lo planes [rows][352] (columns 333..351 zero): the K-padded A operand of skel_embedding
cudaError_t launch_pack_feats(const float* feats, long long rows, float* hi, float* lo, cudaStream_t st);
// xseq = cat(global_motion_token[2], emb[300]) + pe[:302] per clip -> planes [nb*302][128]
cudaError_t launch_encoder_tokens(const float* emb, const float* gtok, const float* pe, int nb, float* hi, float* lo,
                                  cudaStream_t st);
// mu[b] = x[b*302 + 0], logvar[b] = x[b*302 + 1]  (hi + lo)
cudaError_t launch_encoder_dist(const float* hi, const float* lo, int nb, float* mu, float* logvar, cudaStream_t st);
// out [n_steps][B][128]: the N(0,1) draws the sampler kernels make for (seed, clip_offset) -- debug / test hook
cudaError_t launch_philox_export(unsigned long long seed, unsigned long long clip_offset, int B, int n_steps, float* out,
                                 cudaStream_t st);
cudaError_t launch_add_planes(const float* hi, const float* lo, float* out, size_t n, cudaStream_t st);

}  // namespace amuse
