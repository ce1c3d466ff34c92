// Hoisted (step- or batch-invariant) pieces of the denoiser and the pose conversion:
//   time_table_kernel   a3  Timesteps + TimestepEmbedding for every step of the schedule
//   cond_tokens_kernel  a4  emb_proj_{con,emo,sty} (+ the token's learned PE row, a5)
//   rot6d_kernel        a11 6D -> rotation matrix -> quaternion -> axis-angle (+ trans copy)
#include "small_kernels.cuh"

#include "common.cuh"
#include "philox.cuh"

namespace amuse {

// One block per scheduler step.  temb[step] = W2 silu(W1 [cos(t f) | sin(t f)] + b1) + b2
// (reference embeddings.py:245-322, flip_sin_to_cos=True, freq_shift=0).  The product t*f is
// formed in fp32 exactly like the reference (`timesteps[:, None].float() * emb[None, :]`).
__global__ void __launch_bounds__(128) time_table_kernel(const int* __restrict__ timesteps,
                                                         const float* __restrict__ freqs,   // [128]
                                                         const float* __restrict__ w1t,     // [256][128]
                                                         const float* __restrict__ b1,
                                                         const float* __restrict__ w2t,     // [128][128]
                                                         const float* __restrict__ b2, float* __restrict__ temb) {
  __shared__ float e[256];
  __shared__ float h[128];
  const int tid = threadIdx.x;
  const float t = static_cast<float>(timesteps[blockIdx.x]);
  const float arg = __fmul_rn(t, freqs[tid]);
  e[tid] = cosf(arg);
  e[128 + tid] = sinf(arg);
  __syncthreads();
  float acc = b1[tid];
#pragma unroll 8
  for (int k = 0; k < 256; ++k) acc = fmaf(e[k], w1t[k * 128 + tid], acc);
  h[tid] = acc / (1.0f + expf(-acc));   // SiLU
  __syncthreads();
  float out = b2[tid];
#pragma unroll 8
  for (int k = 0; k < 128; ++k) out = fmaf(h[k], w2t[k * 128 + tid], out);
  temb[static_cast<size_t>(blockIdx.x) * 128 + tid] = out;
}

// grid (B, n_cond).  cond[b][ci] = Linear(ReLU(z)) + pe[2 + ci]   (denoiser.py:74-79,153-171)
__global__ void __launch_bounds__(128) cond_tokens_kernel(CondArgs a) {
  __shared__ float r[256];
  const int tid = threadIdx.x, b = blockIdx.x, ci = blockIdx.y;
  const float* z = a.z[ci] + static_cast<size_t>(b) * 256;
  r[tid] = fmaxf(z[tid], 0.f);
  r[tid + 128] = fmaxf(z[tid + 128], 0.f);
  __syncthreads();
  const float* wt = a.wt[ci];
  float acc = a.bias[ci][tid];
#pragma unroll 8
  for (int k = 0; k < 256; ++k) acc = fmaf(r[k], wt[k * 128 + tid], acc);
  a.out[(static_cast<size_t>(b) * 3 + ci) * 128 + tid] = acc + a.pe[(2 + ci) * 128 + tid];
}

// 6D -> axis-angle for one rotation.  Follows dm/utils/transforms.py:187-208 (Gram-Schmidt,
// F.normalize eps 1e-12), :259-309 (sqrt-positive-part + copysign quaternion) and :156-184
// (axis-angle with the 0.5 - x^2/48 small-angle branch) in the reference's operation order.
__device__ __forceinline__ void rot6d_to_axis_angle(const float* __restrict__ s6, float* __restrict__ dst) {
  const float a1x = s6[0], a1y = s6[1], a1z = s6[2];
  const float a2x = s6[3], a2y = s6[4], a2z = s6[5];
  const float n1 = fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);
  const float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
  const float dp = b1x * a2x + b1y * a2y + b1z * a2z;
  float b2x = a2x - dp * b1x, b2y = a2y - dp * b1y, b2z = a2z - dp * b1z;
  const float n2 = fmaxf(sqrtf(b2x * b2x + b2y * b2y + b2z * b2z), 1e-12f);
  b2x /= n2;
  b2y /= n2;
  b2z /= n2;
  const float b3x = b1y * b2z - b1z * b2y, b3y = b1z * b2x - b1x * b2z, b3z = b1x * b2y - b1y * b2x;
  // rotation matrix rows = b1, b2, b3
  const float m00 = b1x, m01 = b1y, m02 = b1z, m10 = b2x, m11 = b2y, m12 = b2z, m20 = b3x, m21 = b3y, m22 = b3z;
  const float qw = 0.5f * sqrtf(fmaxf(0.f, 1.f + m00 + m11 + m22));
  float qx = 0.5f * sqrtf(fmaxf(0.f, 1.f + m00 - m11 - m22));
  float qy = 0.5f * sqrtf(fmaxf(0.f, 1.f - m00 + m11 - m22));
  float qz = 0.5f * sqrtf(fmaxf(0.f, 1.f - m00 - m11 + m22));
  if ((m21 - m12) < 0.f) qx = -qx;
  if ((m02 - m20) < 0.f) qy = -qy;
  if ((m10 - m01) < 0.f) qz = -qz;
  const float nrm = sqrtf(qx * qx + qy * qy + qz * qz);
  const float half = atan2f(nrm, qw);
  const float ang = 2.f * half;
  const float s = (fabsf(ang) < 1e-6f) ? (0.5f - (ang * ang) / 48.f) : (sinf(half) / ang);
  dst[0] = qx / s;
  dst[1] = qy / s;
  dst[2] = qz / s;
}

// One thread per joint of a [frames][333] feature tensor; lane 55 of each frame copies trans.
__global__ void __launch_bounds__(256) rot6d_kernel(const float* __restrict__ feats, int feat_ld, long long n_frames,
                                                    float* __restrict__ poses, float* __restrict__ trans) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = n_frames * 56;     // 55 joints + 1 "trans" lane per frame
  if (gid >= total) return;
  const long long f = gid / 56;
  const int j = static_cast<int>(gid - f * 56);
  const float* src = feats + f * feat_ld;
  if (j == 55) {
    if (trans) {
      trans[f * 3 + 0] = src[330];
      trans[f * 3 + 1] = src[331];
      trans[f * 3 + 2] = src[332];
    }
    return;
  }
  rot6d_to_axis_angle(src + j * 6, poses + (f * 55 + j) * 3);
}

// generic [n][6] -> [n][3]
__global__ void __launch_bounds__(256) rot6d_flat_kernel(const float* __restrict__ d6, long long n,
                                                         float* __restrict__ aa) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid < n) rot6d_to_axis_angle(d6 + gid * 6, aa + gid * 3);
}

cudaError_t launch_time_table(const int* timesteps_dev, int n_steps, const float* freqs, const float* w1t,
                              const float* b1, const float* w2t, const float* b2, float* temb, cudaStream_t st) {
  time_table_kernel<<<n_steps, 128, 0, st>>>(timesteps_dev, freqs, w1t, b1, w2t, b2, temb);
  return cudaGetLastError();
}
cudaError_t launch_cond_tokens(const CondArgs& a, int B, int n_cond, cudaStream_t st) {
  cond_tokens_kernel<<<dim3(B, n_cond), 128, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_rot6d(const float* feats, int feat_ld, long long n_frames, float* poses, float* trans,
                         cudaStream_t st) {
  const long long total = n_frames * 56;
  const int blocks = static_cast<int>((total + 255) / 256);
  rot6d_kernel<<<blocks, 256, 0, st>>>(feats, feat_ld, n_frames, poses, trans);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) add_planes_kernel(const float* __restrict__ hi, const float* __restrict__ lo,
                                                         float* __restrict__ out, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = hi[i] + lo[i];
}
cudaError_t launch_add_planes(const float* hi, const float* lo, float* out, size_t n, cudaStream_t st) {
  add_planes_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(hi, lo, out, n);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- MotionPrior.encode side
// axis-angle -> quaternion -> matrix -> first two rows (dm/utils/transforms.py:228-257, 96-124, 211-226),
// same operation order as the reference in fp32
__global__ void __launch_bounds__(256) motion_to_feats_kernel(const float* __restrict__ poses,
                                                              const float* __restrict__ trans, long long n_frames,
                                                              float* __restrict__ feats) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= n_frames * 56) return;
  const long long f = gid / 56;
  const int j = static_cast<int>(gid - f * 56);
  float* dst = feats + f * 333;
  if (j == 55) {
    dst[330] = trans[f * 3 + 0];
    dst[331] = trans[f * 3 + 1];
    dst[332] = trans[f * 3 + 2];
    return;
  }
  const float* a = poses + (f * 55 + j) * 3;
  const float x = a[0], y = a[1], z = a[2];
  const float ang = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  const float half = 0.5f * ang;
  const float sh = (fabsf(ang) < 1e-6f) ? (0.5f - __fmul_rn(ang, ang) / 48.0f) : __fdiv_rn(sinf(half), ang);
  const float r = cosf(half), i = __fmul_rn(x, sh), jq = __fmul_rn(y, sh), k = __fmul_rn(z, sh);
  const float nn = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, r), __fmul_rn(i, i)), __fmul_rn(jq, jq)), __fmul_rn(k, k));
  const float two_s = __fdiv_rn(2.0f, nn);
  dst[j * 6 + 0] = 1.0f - __fmul_rn(two_s, __fadd_rn(__fmul_rn(jq, jq), __fmul_rn(k, k)));
  dst[j * 6 + 1] = __fmul_rn(two_s, __fsub_rn(__fmul_rn(i, jq), __fmul_rn(k, r)));
  dst[j * 6 + 2] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(i, k), __fmul_rn(jq, r)));
  dst[j * 6 + 3] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(i, jq), __fmul_rn(k, r)));
  dst[j * 6 + 4] = 1.0f - __fmul_rn(two_s, __fadd_rn(__fmul_rn(i, i), __fmul_rn(k, k)));
  dst[j * 6 + 5] = __fmul_rn(two_s, __fsub_rn(__fmul_rn(jq, k), __fmul_rn(i, r)));
}
cudaError_t launch_motion_to_feats(const float* poses, const float* trans, long long n_frames, float* feats,
                                   cudaStream_t st) {
  const long long total = n_frames * 56;
  motion_to_feats_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(poses, trans, n_frames, feats);
  return cudaGetLastError();
}

__device__ __forceinline__ void split_rna(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

__global__ void __launch_bounds__(352) pack_feats_kernel(const float* __restrict__ feats, float* __restrict__ hi,
                                                         float* __restrict__ lo) {
  const long long row = blockIdx.x;
  const int c = threadIdx.x;
  const float v = (c < 333) ? feats[row * 333 + c] : 0.f;
  float h, l;
  split_rna(v, h, l);
  hi[row * 352 + c] = h;
  lo[row * 352 + c] = l;
}
cudaError_t launch_pack_feats(const float* feats, long long rows, float* hi, float* lo, cudaStream_t st) {
  pack_feats_kernel<<<static_cast<unsigned>(rows), 352, 0, st>>>(feats, hi, lo);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128) encoder_tokens_kernel(const float* __restrict__ emb, const float* __restrict__ gtok,
                                                             const float* __restrict__ pe, float* __restrict__ hi,
                                                             float* __restrict__ lo) {
  const int row = blockIdx.x, c = threadIdx.x;      // row = clip * 302 + token
  const int b = row / 302, tok = row - b * 302;
  const float v = (tok < 2) ? gtok[tok * 128 + c] : emb[(static_cast<size_t>(b) * 300 + tok - 2) * 128 + c];
  float h, l;
  split_rna(v + pe[tok * 128 + c], h, l);
  hi[static_cast<size_t>(row) * 128 + c] = h;
  lo[static_cast<size_t>(row) * 128 + c] = l;
}
cudaError_t launch_encoder_tokens(const float* emb, const float* gtok, const float* pe, int nb, float* hi, float* lo,
                                  cudaStream_t st) {
  encoder_tokens_kernel<<<nb * 302, 128, 0, st>>>(emb, gtok, pe, hi, lo);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) encoder_dist_kernel(const float* __restrict__ hi, const float* __restrict__ lo,
                                                           float* __restrict__ mu, float* __restrict__ logvar) {
  const int b = blockIdx.x, t = threadIdx.x >> 7, c = threadIdx.x & 127;
  const size_t src = (static_cast<size_t>(b) * 302 + t) * 128 + c;
  (t ? logvar : mu)[b * 128 + c] = hi[src] + lo[src];
}
cudaError_t launch_encoder_dist(const float* hi, const float* lo, int nb, float* mu, float* logvar, cudaStream_t st) {
  encoder_dist_kernel<<<nb, 256, 0, st>>>(hi, lo, mu, logvar);
  return cudaGetLastError();
}

// Debug export of the sampler's in-kernel noise: out[step][b][e] = philox_normal(seed, (clip_offset + b) * 128 + e, step),
// the very function both denoise-loop kernels call (philox.cuh).
__global__ void __launch_bounds__(128) philox_export_kernel(unsigned long long seed, unsigned long long clip_offset, int B,
                                                            float* __restrict__ out) {
  const int step = blockIdx.y, b = blockIdx.x, e = threadIdx.x;
  out[(static_cast<size_t>(step) * B + b) * 128 + e] =
      philox_normal(seed, (clip_offset + static_cast<unsigned long long>(b)) * 128ull + e, static_cast<uint32_t>(step));
}
cudaError_t launch_philox_export(unsigned long long seed, unsigned long long clip_offset, int B, int n_steps, float* out,
                                 cudaStream_t st) {
  philox_export_kernel<<<dim3(B, n_steps), 128, 0, st>>>(seed, clip_offset, B, out);
  return cudaGetLastError();
}

cudaError_t launch_rot6d_flat(const float* d6, long long n, float* aa, cudaStream_t st) {
  rot6d_flat_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(d6, n, aa);
  return cudaGetLastError();
}

}  // namespace amuse
