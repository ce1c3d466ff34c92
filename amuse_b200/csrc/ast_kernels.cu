// K5: the three AST (DeiT-base-distilled) audio encoders of AST_EVP.eval_func
// (reference models/audio/AST_EVP.py:84-90, models/audio/audio_main_new.py:174-204; the ViT blocks
// are timm 0.4.5's -- restated, see oracle/ast_ref.py).
//
// All dense layers run on the tcgen05 3xTF32 GEMM (tc_gemm.cu): patch embedding as an im2col GEMM
// (K = 256), qkv / proj / fc1 / fc2 with bias, q-scaling, erf-GELU and the residual add fused in the
// epilogues.  Activations travel as TF32 hi/lo planes (value = hi + lo).  The 1214-token attention
// (12 heads x 64) is the tcgen05 flash-attention kernel of ast_attn.cu, fed by per-head q / k / v^T
// planes written straight from the qkv GEMM epilogue; LayerNorm (eps 1e-6) is an fp32 CUDA-core kernel.
// The three branches share one im2col of the filterbank; clips are processed in chunks.
#include "ast_kernels.cuh"

#include <algorithm>
#include <cstring>

#include "../../include/amuse_b200.h"
#include "ast_attn.cuh"
#include "common.cuh"
#include "tc_gemm.cuh"

namespace amuse {
namespace ast {

namespace {

constexpr int D = 768, HEADS = 12, HD = 64, TOK = 1214, PATCH = 1212, FF = 3072, FEAT = 256;
// Clips per pass.  One pass costs whole waves: the GEMMs run ceil(tiles / 74 CTA pairs) rounds of 256x256
// tiles and the attention ceil(60 * clips / 148) rounds of CTAs, so the batch is cut into equal passes of
// at most 32 clips (32 clips: attention 12.97 waves, N = 768 GEMMs 6.16 waves; workspace ~2 GB)
constexpr int kChunkMax = 32;

struct Buf {
  float* p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t want) {
    if (want <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(&p, want * sizeof(float));
    if (e == cudaSuccess) n = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct Block {
  const float *ln1, *ln1b, *ln2, *ln2b, *qkv_b, *proj_b, *fc1_b, *fc2_b;
  float *qkv_w, *proj_w, *fc1_w, *fc2_w;   // planes (hi at p, lo at p + n)
};
struct Branch {
  const float *patch_b, *cls, *dist, *pos, *norm_w, *norm_b, *head_ln_w, *head_ln_b, *head_w, *head_b;
  float* patch_w;   // planes [768][256]
  std::vector<Block> blk;
};
struct Impl {
  int depth = 0;
  Branch br[3];   // con, emo, sty
  std::vector<float*> owned;
  Buf P, tmp, X, Hn, Qp, Kp, Vp, O, Hid, pooled;
};

__device__ __forceinline__ void split2(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

// patch rows of the strided 16x16 convolution: row = b*1212 + f*101 + t, k = df*16 + dt,
// value = x[b][0][f*10+df][t*10+dt] = fbank[b][t*10+dt][f*10+df]   (audio_main_new.py:181-182,185)
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ fbank, float* __restrict__ hi,
                                                     float* __restrict__ lo) {
  const int row = blockIdx.x, k = threadIdx.x;
  const int b = row / PATCH, p = row - b * PATCH;
  const int f = p / 101, t = p - f * 101;
  const int df = k >> 4, dt = k & 15;
  const float v = fbank[(static_cast<size_t>(b) * 1024 + t * 10 + dt) * 128 + f * 10 + df];
  float h, l;
  split2(v, h, l);
  hi[static_cast<size_t>(row) * 256 + k] = h;
  lo[static_cast<size_t>(row) * 256 + k] = l;
}

// x = cat(cls, dist, patches) + pos_embed  -> planes   (audio_main_new.py:186-189)
__global__ void __launch_bounds__(192) assemble_kernel(const float* __restrict__ patches, const float* __restrict__ cls,
                                                       const float* __restrict__ dist, const float* __restrict__ pos,
                                                       float* __restrict__ hi, float* __restrict__ lo) {
  const int row = blockIdx.x;          // b*1214 + tok
  const int b = row / TOK, tok = row - b * TOK;
  const int c = threadIdx.x * 4;
  float4 v;
  if (tok == 0)
    v = *reinterpret_cast<const float4*>(cls + c);
  else if (tok == 1)
    v = *reinterpret_cast<const float4*>(dist + c);
  else
    v = *reinterpret_cast<const float4*>(patches + (static_cast<size_t>(b) * PATCH + tok - 2) * D + c);
  const float4 pe = *reinterpret_cast<const float4*>(pos + static_cast<size_t>(tok) * D + c);
  float4 h, l;
  split2(v.x + pe.x, h.x, l.x);
  split2(v.y + pe.y, h.y, l.y);
  split2(v.z + pe.z, h.z, l.z);
  split2(v.w + pe.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi + static_cast<size_t>(row) * D + c) = h;
  *reinterpret_cast<float4*>(lo + static_cast<size_t>(row) * D + c) = l;
}

// LayerNorm over 768 columns, one warp per row, planes in -> planes out
__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ xh, const float* __restrict__ xl,
                                                      const float* __restrict__ g, const float* __restrict__ bta,
                                                      float eps, float* __restrict__ oh, float* __restrict__ ol,
                                                      int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* ph = reinterpret_cast<const float4*>(xh + static_cast<size_t>(row) * D);
  const float4* pl = reinterpret_cast<const float4*>(xl + static_cast<size_t>(row) * D);
  float4 v[6];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 a = ph[lane + 32 * i], b = pl[lane + 32 * i];
    v[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    v[i].x -= mean;
    v[i].y -= mean;
    v[i].z -= mean;
    v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + eps);
  float4* qh = reinterpret_cast<float4*>(oh + static_cast<size_t>(row) * D);
  float4* ql = reinterpret_cast<float4*>(ol + static_cast<size_t>(row) * D);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 gg = reinterpret_cast<const float4*>(g)[lane + 32 * i];
    const float4 bb = reinterpret_cast<const float4*>(bta)[lane + 32 * i];
    float4 h, l;
    split2(v[i].x * rstd * gg.x + bb.x, h.x, l.x);
    split2(v[i].y * rstd * gg.y + bb.y, h.y, l.y);
    split2(v[i].z * rstd * gg.z + bb.z, h.z, l.z);
    split2(v[i].w * rstd * gg.w + bb.w, h.w, l.w);
    qh[lane + 32 * i] = h;
    ql[lane + 32 * i] = l;
  }
}

// pooled[b][c] = mean over tokens 2..1213 of (hi + lo)   (frame_based_feats=True, audio_main_new.py:194-196)
__global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ xh, const float* __restrict__ xl,
                                                   float* __restrict__ pooled) {
  const int b = blockIdx.y, c = blockIdx.x * 256 + threadIdx.x;
  const size_t base = (static_cast<size_t>(b) * TOK + 2) * D + c;
  float s = 0.f;
  for (int r = 0; r < PATCH; ++r) s += xh[base + static_cast<size_t>(r) * D] + xl[base + static_cast<size_t>(r) * D];
  pooled[static_cast<size_t>(b) * D + c] = s * (1.0f / PATCH);
}

// feature = Linear(768->256)(LayerNorm_1e-5(pooled))   (audio_main_new.py:74,197)
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ pooled, const float* __restrict__ g,
                                                   const float* __restrict__ bta, const float* __restrict__ w,
                                                   const float* __restrict__ bias, float* __restrict__ out) {
  __shared__ float hbuf[D];
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float v[3], s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = pooled[static_cast<size_t>(b) * D + tid + 256 * i];
    s += v[i];
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float mean = tot * (1.0f / D);
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] -= mean;
    q += v[i] * v[i];
  }
  q = warp_sum(q);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float rstd = 1.0f / sqrtf(tot * (1.0f / D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 3; ++i) hbuf[tid + 256 * i] = v[i] * rstd * g[tid + 256 * i] + bta[tid + 256 * i];
  __syncthreads();
  const float4* wr = reinterpret_cast<const float4*>(w + static_cast<size_t>(tid) * D);
  float a0 = 0.f, a1 = 0.f;
  for (int c = 0; c < D / 4; ++c) {
    const float4 t = wr[c];
    a0 = fmaf(t.x, hbuf[c * 4 + 0], a0);
    a1 = fmaf(t.y, hbuf[c * 4 + 1], a1);
    a0 = fmaf(t.z, hbuf[c * 4 + 2], a0);
    a1 = fmaf(t.w, hbuf[c * 4 + 3], a1);
  }
  out[static_cast<size_t>(b) * FEAT + tid] = a0 + a1 + bias[tid];
}

const DevTensor* get(Weights& w, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = w.raw.find(key);
  if (it == w.raw.end()) {
    w.err = "missing weight ast." + key;
    return nullptr;
  }
  if (it->second.shape != std::vector<int64_t>(shape)) {
    w.err = "wrong shape for ast." + key;
    return nullptr;
  }
  return &it->second;
}

}  // namespace

int stage(Weights& w, const std::string& key, const void* data, const int64_t* shape, int ndim) {
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  DevTensor t;
  t.shape.assign(shape, shape + ndim);
  if (cudaMalloc(&t.p, static_cast<size_t>(n) * sizeof(float)) != cudaSuccess) return AMUSE_E_CUDA;
  if (cudaMemcpy(t.p, data, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDefault) != cudaSuccess) {
    cudaFree(t.p);
    return AMUSE_E_CUDA;
  }
  auto it = w.raw.find(key);
  if (it != w.raw.end()) cudaFree(it->second.p);
  w.raw[key] = t;
  w.is_ready = false;
  return AMUSE_OK;
}
bool staged(const Weights& w) { return !w.raw.empty(); }
bool ready(const Weights& w) { return w.is_ready; }

static void free_impl(Weights& w) {
  Impl* im = static_cast<Impl*>(w.impl);
  if (!im) return;
  for (float* p : im->owned) cudaFree(p);
  Buf* bufs[] = {&im->P, &im->tmp, &im->X, &im->Hn, &im->Qp, &im->Kp, &im->Vp, &im->O, &im->Hid, &im->pooled};
  for (Buf* b : bufs) b->release();
  delete im;
  w.impl = nullptr;
}

int finalize(Weights& w, cudaStream_t st) {
  free_impl(w);
  Impl* im = new Impl();
  w.impl = im;
  const char* names[3] = {"con_enc", "emo_enc", "sty_enc"};
  int depth = 0;
  while (w.raw.count(std::string(names[0]) + ".v.blocks." + std::to_string(depth) + ".norm1.weight")) ++depth;
  if (depth == 0) {
    w.err = "no ast.con_enc.v.blocks.* weights loaded";
    return AMUSE_E_MISSING;
  }
  im->depth = depth;
  auto planes = [&](const DevTensor* t, size_t n, float** out) -> int {
    float* p = nullptr;
    if (cudaMalloc(&p, 2 * n * sizeof(float)) != cudaSuccess) return AMUSE_E_CUDA;
    im->owned.push_back(p);
    if (tc::split_planes(t->p, p, p + n, n, st) != cudaSuccess) return AMUSE_E_CUDA;
    *out = p;
    return AMUSE_OK;
  };
#define GET(var, key, ...)                                  \
  const DevTensor* var = get(w, key, {__VA_ARGS__});        \
  if (!var) return AMUSE_E_MISSING;
  for (int b = 0; b < 3; ++b) {
    Branch& br = im->br[b];
    const std::string P = names[b], V = P + ".v";
    GET(pw, V + ".patch_embed.proj.weight", D, 1, 16, 16);
    GET(pb, V + ".patch_embed.proj.bias", D);
    GET(cls, V + ".cls_token", 1, 1, D);
    GET(dist, V + ".dist_token", 1, 1, D);
    GET(pos, V + ".pos_embed", 1, TOK, D);
    GET(nw, V + ".norm.weight", D);
    GET(nb, V + ".norm.bias", D);
    GET(hlw, P + ".feature_head.0.weight", D);
    GET(hlb, P + ".feature_head.0.bias", D);
    GET(hw, P + ".feature_head.1.weight", FEAT, D);
    GET(hb, P + ".feature_head.1.bias", FEAT);
    if (int rc = planes(pw, static_cast<size_t>(D) * 256, &br.patch_w)) return rc;
    br.patch_b = pb->p; br.cls = cls->p; br.dist = dist->p; br.pos = pos->p;
    br.norm_w = nw->p; br.norm_b = nb->p; br.head_ln_w = hlw->p; br.head_ln_b = hlb->p;
    br.head_w = hw->p; br.head_b = hb->p;
    br.blk.resize(depth);
    for (int i = 0; i < depth; ++i) {
      const std::string B = V + ".blocks." + std::to_string(i);
      GET(l1w, B + ".norm1.weight", D);
      GET(l1b, B + ".norm1.bias", D);
      GET(l2w, B + ".norm2.weight", D);
      GET(l2b, B + ".norm2.bias", D);
      GET(qw, B + ".attn.qkv.weight", 3 * D, D);
      GET(qb, B + ".attn.qkv.bias", 3 * D);
      GET(ow, B + ".attn.proj.weight", D, D);
      GET(ob, B + ".attn.proj.bias", D);
      GET(f1w, B + ".mlp.fc1.weight", FF, D);
      GET(f1b, B + ".mlp.fc1.bias", FF);
      GET(f2w, B + ".mlp.fc2.weight", D, FF);
      GET(f2b, B + ".mlp.fc2.bias", D);
      Block& k = br.blk[i];
      k.ln1 = l1w->p; k.ln1b = l1b->p; k.ln2 = l2w->p; k.ln2b = l2b->p;
      k.qkv_b = qb->p; k.proj_b = ob->p; k.fc1_b = f1b->p; k.fc2_b = f2b->p;
      if (int rc = planes(qw, static_cast<size_t>(3) * D * D, &k.qkv_w)) return rc;
      if (int rc = planes(ow, static_cast<size_t>(D) * D, &k.proj_w)) return rc;
      if (int rc = planes(f1w, static_cast<size_t>(FF) * D, &k.fc1_w)) return rc;
      if (int rc = planes(f2w, static_cast<size_t>(D) * FF, &k.fc2_w)) return rc;
    }
  }
#undef GET
  if (cudaStreamSynchronize(st) != cudaSuccess) return AMUSE_E_CUDA;
  w.is_ready = true;
  return AMUSE_OK;
}

int forward(Weights& w, int B, const float* fbank, float* con, float* emo, float* sty, cudaStream_t st,
            int64_t* launches) {
  Impl* im = static_cast<Impl*>(w.impl);
  if (!im || !w.is_ready) return AMUSE_E_STATE;
#define CK(call)                                             \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) {                                \
      w.err = std::string(#call) + ": " + cudaGetErrorString(e__); \
      return AMUSE_E_CUDA;                                   \
    }                                                        \
  } while (0)
  const int n_pass = (B + kChunkMax - 1) / kChunkMax;
  const int kChunk = (B + n_pass - 1) / n_pass;
  const int cb = kChunk;
  const size_t Mp = static_cast<size_t>(cb) * PATCH, M = static_cast<size_t>(cb) * TOK;
  CK(im->P.ensure(2 * Mp * 256));
  CK(im->tmp.ensure(Mp * D));
  CK(im->X.ensure(2 * M * D));
  CK(im->Hn.ensure(2 * M * D));
  {   // per-head q / k / v^T operand planes of the attention kernel; the pad rows / columns (tokens
      // 1214..1279) are never written afterwards and must stay zero
    const size_t pe = 2 * attn::plane_elems(cb);
    Buf* qkv[3] = {&im->Qp, &im->Kp, &im->Vp};
    for (Buf* b : qkv) {
      if (b->n >= pe) continue;
      CK(b->ensure(pe));
      CK(cudaMemsetAsync(b->p, 0, pe * sizeof(float), st));
    }
  }
  CK(im->O.ensure(2 * M * D));
  CK(im->Hid.ensure(2 * M * FF));
  CK(im->pooled.ensure(static_cast<size_t>(cb) * D));
  float* outs[3] = {con, emo, sty};
  int64_t n_launch = 0;
  for (int b0 = 0; b0 < B; b0 += kChunk) {
    const int nb = std::min(kChunk, B - b0);
    const int mp = nb * PATCH, m = nb * TOK;
    float *Ph = im->P.p, *Pl = im->P.p + Mp * 256;
    float *Xh = im->X.p, *Xl = im->X.p + M * D, *Hh = im->Hn.p, *Hl = im->Hn.p + M * D;
    float *Oh = im->O.p, *Ol = im->O.p + M * D, *Fh = im->Hid.p, *Fl = im->Hid.p + M * FF;
    const size_t pe1 = attn::plane_elems(cb);
    im2col_kernel<<<mp, 256, 0, st>>>(fbank + static_cast<size_t>(b0) * 1024 * 128, Ph, Pl);
    CK(cudaGetLastError());
    ++n_launch;
    for (int r = 0; r < 3; ++r) {
      const Branch& br = im->br[r];
      tc::GemmDesc g{};
      g.A_hi = Ph; g.A_lo = Pl; g.lda = 256;
      g.W_hi = br.patch_w; g.W_lo = br.patch_w + static_cast<size_t>(D) * 256; g.ldw = 256;
      g.M = mp; g.N = D; g.K = 256; g.bias = br.patch_b; g.C = im->tmp.p; g.ldc = D;
      CK(tc::gemm2(tc::EPI_PLAIN, g, st));
      assemble_kernel<<<m, 192, 0, st>>>(im->tmp.p, br.cls, br.dist, br.pos, Xh, Xl);
      CK(cudaGetLastError());
      n_launch += 2;
      for (int i = 0; i < im->depth; ++i) {
        const Block& k = br.blk[i];
        ln_rows_kernel<<<(m + 7) / 8, 256, 0, st>>>(Xh, Xl, k.ln1, k.ln1b, 1e-6f, Hh, Hl, m);
        CK(cudaGetLastError());
        g = tc::GemmDesc{};
        g.A_hi = Hh; g.A_lo = Hl; g.lda = D;
        g.W_hi = k.qkv_w; g.W_lo = k.qkv_w + static_cast<size_t>(3) * D * D; g.ldw = D;
        g.M = m; g.N = 3 * D; g.K = D; g.bias = k.qkv_b;
        g.q_scale = 0.125f * 1.4426950408889634f;   // head_dim^-0.5 (timm Attention.scale) * log2(e): ex2 softmax
        g.q_hi = im->Qp.p; g.q_lo = im->Qp.p + pe1; g.k_hi = im->Kp.p; g.k_lo = im->Kp.p + pe1;
        g.vt_hi = im->Vp.p; g.vt_lo = im->Vp.p + pe1; g.tok = TOK; g.tokp = attn::kTokP; g.heads = HEADS;
        CK(tc::gemm2(tc::EPI_QKV_HEADS, g, st));
        attn::AttnArgs aa{g.q_hi, g.q_lo, g.k_hi, g.k_lo, g.vt_hi, g.vt_lo, Oh, Ol, nb};
        CK(attn::attention(aa, st));
        g = tc::GemmDesc{};
        g.A_hi = Oh; g.A_lo = Ol; g.lda = D;
        g.W_hi = k.proj_w; g.W_lo = k.proj_w + static_cast<size_t>(D) * D; g.ldw = D;
        g.M = m; g.N = D; g.K = D; g.bias = k.proj_b;
        g.R_hi = Xh; g.R_lo = Xl; g.ldr = D; g.C_hi = Xh; g.C_lo = Xl; g.ldc = D;   // x += proj(o), in place
        CK(tc::gemm2(tc::EPI_RES_PLANES, g, st));
        ln_rows_kernel<<<(m + 7) / 8, 256, 0, st>>>(Xh, Xl, k.ln2, k.ln2b, 1e-6f, Hh, Hl, m);
        CK(cudaGetLastError());
        g = tc::GemmDesc{};
        g.A_hi = Hh; g.A_lo = Hl; g.lda = D;
        g.W_hi = k.fc1_w; g.W_lo = k.fc1_w + static_cast<size_t>(FF) * D; g.ldw = D;
        g.M = m; g.N = FF; g.K = D; g.bias = k.fc1_b; g.C_hi = Fh; g.C_lo = Fl; g.ldc = FF;
        CK(tc::gemm2(tc::EPI_GELU_PLANES, g, st));
        g = tc::GemmDesc{};
        g.A_hi = Fh; g.A_lo = Fl; g.lda = FF;
        g.W_hi = k.fc2_w; g.W_lo = k.fc2_w + static_cast<size_t>(D) * FF; g.ldw = FF;
        g.M = m; g.N = D; g.K = FF; g.bias = k.fc2_b;
        g.R_hi = Xh; g.R_lo = Xl; g.ldr = D; g.C_hi = Xh; g.C_lo = Xl; g.ldc = D;   // x += fc2(h), in place
        CK(tc::gemm2(tc::EPI_RES_PLANES, g, st));
        n_launch += 7;
      }
      ln_rows_kernel<<<(m + 7) / 8, 256, 0, st>>>(Xh, Xl, br.norm_w, br.norm_b, 1e-6f, Hh, Hl, m);
      CK(cudaGetLastError());
      pool_kernel<<<dim3(D / 256, nb), 256, 0, st>>>(Hh, Hl, im->pooled.p);
      CK(cudaGetLastError());
      head_kernel<<<nb, 256, 0, st>>>(im->pooled.p, br.head_ln_w, br.head_ln_b, br.head_w, br.head_b,
                                      outs[r] + static_cast<size_t>(b0) * FEAT);
      CK(cudaGetLastError());
      n_launch += 3;
    }
  }
#undef CK
  if (launches) *launches = n_launch;
  return AMUSE_OK;
}

void release(Weights& w) {
  free_impl(w);
  for (auto& kv : w.raw) cudaFree(kv.second.p);
  w.raw.clear();
  w.is_ready = false;
}
const char* last_error(const Weights& w) { return w.err.c_str(); }

}  // namespace ast
}  // namespace amuse
