// K3 (first implementation): MotionPrior.decode as batched fp32 kernels over M = clips*300 rows.
//   gemm_nt_kernel         C = epi(A[M][K] . Wt[K][N] + bias)  64x128x16 tiles, 3-stage cp.async ring,
//                          4x8 register tiles; epilogues fuse bias, q-scaling, erf-GELU, residual +
//                          LayerNorm and the collapsed 1-key cross-attention + second LayerNorm.
//   self_attention_kernel  softmax(q k^T) v per (clip, head): K/V of the head resident in shared
//                          memory (77 KB), one query row per thread, online softmax.
//   cross_vectors_kernel   TransformerDecoderLayer.multihead_attn over a 1-token memory: the softmax
//                          over one key is 1, so the sub-block reduces to out_proj(W_v z + b_v) added
//                          to every frame of the clip (SURVEY.md App. C identity (i); exact).
// Reference: models/latent_diffusion/utils/cross_attention.py:89-125,323-345.
#include "decode_kernels.cuh"

#include "common.cuh"

namespace amuse {
namespace dec {

namespace {

constexpr int BM = 64, BN = 128, BK = 16, STAGES = 3;
constexpr int ALD = BK + 4;   // A tile row stride (floats): 16-B aligned rows, conflict-free k-vector reads

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const int sz = pred ? 16 : 0;   // src-size 0 => zero-fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void split_tf32_rna(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm of a 128-wide row spread over 16 lanes x 8 values (lane owns cols tx*4+{0..3}, 64+tx*4+{0..3})
__device__ __forceinline__ void row_layernorm(float (&v)[8], const float* __restrict__ g, const float* __restrict__ b,
                                              int tx) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  const float mean = group16_sum(s) * (1.0f / 128.0f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] -= mean;
    q += v[j] * v[j];
  }
  const float rstd = 1.0f / sqrtf(group16_sum(q) * (1.0f / 128.0f) + kLnEps);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4);
    v[j] = v[j] * rstd * g[c] + b[c];
  }
}

}  // namespace

template <int EPI>
__global__ void __launch_bounds__(256) gemm_nt_kernel(const GemmArgs a) {
  __shared__ __align__(16) float As[STAGES][BM * ALD];
  __shared__ __align__(16) float Ws[STAGES][BK * BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int nk = a.K / BK;

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
    {   // A: 64 rows x 16 k  = 256 x 16-B chunks
      const int row = tid >> 2, kc = (tid & 3) * 4;
      const int m = m0 + row;
      const bool ok = m < a.M;
      const float* src;
      if (a.A2 != nullptr && k0 >= 128)
        src = a.A2 + static_cast<size_t>(ok ? m : 0) * a.lda2 + (k0 - 128) + kc;
      else
        src = a.A + static_cast<size_t>(ok ? m : 0) * a.lda + k0 + kc;
      cp_async16(&As[stage][row * ALD + kc], src, ok);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {   // W: 16 k x 128 n = 512 chunks
      const int c = tid + q * 256;
      const int k = c >> 5, n4 = (c & 31) * 4;
      const bool ok = (n0 + n4) < a.ldw;
      const float* src = a.Wt + static_cast<size_t>(k0 + k) * a.ldw + (ok ? (n0 + n4) : 0);
      cp_async16(&Ws[stage][k * BN + n4], src, ok);
    }
  };

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    const float* as = As[kt % STAGES] + (ty * 4) * ALD;
    const float* ws = Ws[kt % STAGES] + tx * 4;
#pragma unroll
    for (int kq = 0; kq < BK; kq += 4) {
      float4 av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(as + i * ALD + kq);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const float4 b0 = *reinterpret_cast<const float4*>(ws + (kq + k4) * BN);
        const float4 b1 = *reinterpret_cast<const float4*>(ws + (kq + k4) * BN + 64);
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x = (k4 == 0) ? av[i].x : (k4 == 1) ? av[i].y : (k4 == 2) ? av[i].z : av[i].w;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(x, bv[j], acc[i][j]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // ------------------------------------------------------------------ epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    const bool row_ok = m < a.M;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4));
      v[j] = acc[i][j] + ((c < a.N) ? a.bias[c] : 0.f);
    }
    if (EPI == EPI_QKV) {
      if (n0 == 0) {   // columns 0..127 are q: nn.MultiheadAttention scales q by head_dim^-0.5
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= 0.17677669529663687f;
      }
    } else if (EPI == EPI_GELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
    } else if (EPI == EPI_RES_LN || EPI == EPI_RES_LN_CROSS_LN) {
      const float* r = a.R + static_cast<size_t>(row_ok ? m : 0) * a.ldr;
      const float4 r0 = *reinterpret_cast<const float4*>(r + tx * 4);
      const float4 r1 = *reinterpret_cast<const float4*>(r + 64 + tx * 4);
      v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
      v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
      row_layernorm(v, a.ln_g, a.ln_b, tx);
      if (EPI == EPI_RES_LN_CROSS_LN) {
        const float* cv = a.cvec + static_cast<size_t>((row_ok ? m : 0) / a.rows_per_clip) * 128;
        const float4 c0 = *reinterpret_cast<const float4*>(cv + tx * 4);
        const float4 c1 = *reinterpret_cast<const float4*>(cv + 64 + tx * 4);
        v[0] += c0.x; v[1] += c0.y; v[2] += c0.z; v[3] += c0.w;
        v[4] += c1.x; v[5] += c1.y; v[6] += c1.z; v[7] += c1.w;
        row_layernorm(v, a.ln2_g, a.ln2_b, tx);
      }
    }
    if (row_ok) {
      float* dst = a.C + static_cast<size_t>(m) * a.ldc + n0;
      if (n0 + BN <= a.N && (a.ldc & 3) == 0) {
        *reinterpret_cast<float4*>(dst + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 64 + tx * 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int cl = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4);
          if (n0 + cl < a.N) dst[cl] = v[j];
        }
      }
    }
  }
}

cudaError_t launch_gemm(int epi, const GemmArgs& a, cudaStream_t st) {
  if (a.K % BK != 0) return cudaErrorInvalidValue;
  if ((epi == EPI_RES_LN || epi == EPI_RES_LN_CROSS_LN) && a.N != 128) return cudaErrorInvalidValue;
  dim3 grid((a.M + BM - 1) / BM, (a.N + BN - 1) / BN);
  switch (epi) {
    case EPI_BIAS: gemm_nt_kernel<EPI_BIAS><<<grid, 256, 0, st>>>(a); break;
    case EPI_QKV: gemm_nt_kernel<EPI_QKV><<<grid, 256, 0, st>>>(a); break;
    case EPI_GELU: gemm_nt_kernel<EPI_GELU><<<grid, 256, 0, st>>>(a); break;
    case EPI_RES_LN: gemm_nt_kernel<EPI_RES_LN><<<grid, 256, 0, st>>>(a); break;
    case EPI_RES_LN_CROSS_LN: gemm_nt_kernel<EPI_RES_LN_CROSS_LN><<<grid, 256, 0, st>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------- self attention
template <bool PLANES>
__global__ void __launch_bounds__(128) self_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                             float* __restrict__ out_lo, int frames) {
  extern __shared__ __align__(16) float kv[];
  float* Ks = kv;
  float* Vs = kv + frames * 32;
  const int tid = threadIdx.x, h = blockIdx.y, b = blockIdx.z;
  const float* base = qkv + static_cast<size_t>(b) * frames * 384;
  for (int idx = tid; idx < frames * 8; idx += 128) {
    const int row = idx >> 3, c4 = (idx & 7) * 4;
    const float* src = base + static_cast<size_t>(row) * 384 + h * 32 + c4;
    *reinterpret_cast<float4*>(Ks + row * 32 + c4) = *reinterpret_cast<const float4*>(src + 128);
    *reinterpret_cast<float4*>(Vs + row * 32 + c4) = *reinterpret_cast<const float4*>(src + 256);
  }
  const int qi = blockIdx.x * 128 + tid;
  const bool valid = qi < frames;
  float q[32];
  {
    const float* src = base + static_cast<size_t>(valid ? qi : 0) * 384 + h * 32;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 t = *reinterpret_cast<const float4*>(src + c * 4);
      q[c * 4 + 0] = t.x; q[c * 4 + 1] = t.y; q[c * 4 + 2] = t.z; q[c * 4 + 3] = t.w;
    }
  }
  __syncthreads();
  float m = -INFINITY, l = 0.f;
  float acc[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) acc[d] = 0.f;
  for (int j0 = 0; j0 < frames; j0 += 4) {
    float sc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = min(j0 + u, frames - 1);
      const float* kr = Ks + j * 32;
      float s0 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(kr + c * 4);
        s0 = fmaf(q[c * 4 + 0], t.x, s0);
        s0 = fmaf(q[c * 4 + 1], t.y, s0);
        s0 = fmaf(q[c * 4 + 2], t.z, s0);
        s0 = fmaf(q[c * 4 + 3], t.w, s0);
      }
      sc[u] = (j0 + u < frames) ? s0 : -INFINITY;
    }
    const float cm = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
    if (cm > m) {
      const float f = expf(m - cm);
      l *= f;
#pragma unroll
      for (int d = 0; d < 32; ++d) acc[d] *= f;
      m = cm;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float pexp = expf(sc[u] - m);
      l += pexp;
      const float* vr = Vs + min(j0 + u, frames - 1) * 32;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(vr + c * 4);
        acc[c * 4 + 0] = fmaf(pexp, t.x, acc[c * 4 + 0]);
        acc[c * 4 + 1] = fmaf(pexp, t.y, acc[c * 4 + 1]);
        acc[c * 4 + 2] = fmaf(pexp, t.z, acc[c * 4 + 2]);
        acc[c * 4 + 3] = fmaf(pexp, t.w, acc[c * 4 + 3]);
      }
    }
  }
  if (valid) {
    const float inv = 1.0f / l;
    const size_t off = (static_cast<size_t>(b) * frames + qi) * 128 + h * 32;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 v = make_float4(acc[c * 4 + 0] * inv, acc[c * 4 + 1] * inv, acc[c * 4 + 2] * inv, acc[c * 4 + 3] * inv);
      if (PLANES) {   // TF32 hi/lo planes for the tcgen05 out_proj GEMM
        float4 hi, lo;
        split_tf32_rna(v.x, hi.x, lo.x);
        split_tf32_rna(v.y, hi.y, lo.y);
        split_tf32_rna(v.z, hi.z, lo.z);
        split_tf32_rna(v.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(out + off + c * 4) = hi;
        *reinterpret_cast<float4*>(out_lo + off + c * 4) = lo;
      } else {
        *reinterpret_cast<float4*>(out + off + c * 4) = v;
      }
    }
  }
}

constexpr int kAttnMaxFrames = 320;
template <bool PLANES>
static cudaError_t launch_attn(const float* qkv, float* out, float* out_lo, int clips, int frames, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(frames) * 64 * sizeof(float);
  static PerDeviceOnce once;   // sized for the longest sequence either caller uses (302 encoder tokens)
  if (smem > static_cast<size_t>(kAttnMaxFrames) * 64 * sizeof(float)) return cudaErrorInvalidValue;
  if (cudaError_t e = once.run([] {
        return cudaFuncSetAttribute(self_attention_kernel<PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kAttnMaxFrames * 64 * static_cast<int>(sizeof(float)));
      }))
    return e;
  dim3 grid((frames + 127) / 128, kHeads, clips);
  self_attention_kernel<PLANES><<<grid, 128, smem, st>>>(qkv, out, out_lo, frames);
  return cudaGetLastError();
}
cudaError_t launch_self_attention(const float* qkv, float* out, int clips, int frames, cudaStream_t st) {
  return launch_attn<false>(qkv, out, nullptr, clips, frames, st);
}
cudaError_t launch_self_attention_planes(const float* qkv, float* out_hi, float* out_lo, int clips, int frames,
                                         cudaStream_t st) {
  return launch_attn<true>(qkv, out_hi, out_lo, clips, frames, st);
}

// ------------------------------------------------------------------------- 1-key cross attention
__global__ void __launch_bounds__(128) cross_vectors_kernel(const float* __restrict__ z, const float* __restrict__ wv_t,
                                                            const float* __restrict__ bv, const float* __restrict__ wo_t,
                                                            const float* __restrict__ bo, float* __restrict__ cvec,
                                                            int B) {
  __shared__ float zs[128];
  __shared__ float t1[128];
  const int tid = threadIdx.x, b = blockIdx.x, l = blockIdx.y;
  zs[tid] = z[static_cast<size_t>(b) * 128 + tid];
  __syncthreads();
  const float* wv = wv_t + static_cast<size_t>(l) * 128 * 128;
  float acc = bv[l * 128 + tid];
#pragma unroll 8
  for (int k = 0; k < 128; ++k) acc = fmaf(zs[k], wv[k * 128 + tid], acc);
  t1[tid] = acc;
  __syncthreads();
  const float* wo = wo_t + static_cast<size_t>(l) * 128 * 128;
  float o = bo[l * 128 + tid];
#pragma unroll 8
  for (int k = 0; k < 128; ++k) o = fmaf(t1[k], wo[k * 128 + tid], o);
  cvec[(static_cast<size_t>(l) * B + b) * 128 + tid] = o;
}

cudaError_t launch_cross_vectors(const float* z, const float* wv_t, const float* bv, const float* wo_t,
                                 const float* bo, float* cvec, int B, cudaStream_t st) {
  cross_vectors_kernel<<<dim3(B, kLayers), 128, 0, st>>>(z, wv_t, bv, wo_t, bo, cvec, B);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) broadcast_rows_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                             int per_clip4, long long total4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total4) dst[i] = src[i % per_clip4];
}

cudaError_t launch_broadcast_rows(const float* src, float* dst, int clips, int frames, cudaStream_t st) {
  const int per4 = frames * 32;
  const long long total4 = static_cast<long long>(clips) * per4;
  broadcast_rows_kernel<<<static_cast<int>((total4 + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), per4, total4);
  return cudaGetLastError();
}

}  // namespace dec
}  // namespace amuse
