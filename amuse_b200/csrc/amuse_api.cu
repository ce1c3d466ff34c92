// C ABI of the engine (include/amuse_b200.h): context, weight staging / repacking, scheduler
// tables, workspaces and the launch sequences of the sampler and the decoder.
#include "../../include/amuse_b200.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "ast_attn.cuh"
#include "ast_kernels.cuh"
#include "common.cuh"
#include "decode_kernels.cuh"
#include "denoise_loop.cuh"
#include "denoise_tc.cuh"
#include "fbank_kernel.cuh"
#include "small_kernels.cuh"
#include "tc_gemm.cuh"

using namespace amuse;

namespace {

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct DevBuf {
  float* p = nullptr;
  size_t n = 0;   // floats
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

struct Schedule {
  int n_steps = 0;
  std::vector<int32_t> timesteps;
  std::vector<float> coef;   // [n][5]
  int* d_timesteps = nullptr;
  float* d_coef = nullptr;
  float* d_temb = nullptr;
  bool dir_uses_eps = true;
  bool has_sigma = false;
};

const char* kBlocks[9] = {"input_blocks.0", "input_blocks.1", "input_blocks.2", "input_blocks.3", "middle_block",
                          "output_blocks.0", "output_blocks.1", "output_blocks.2", "output_blocks.3"};

// ---- decoder weights on the device (all K-major) ----------------------------------------
struct DecLayerW {
  size_t wqkv_t, bqkv, wo_t, bo, ln1, w1_t, b1, w2_t, b2, ln2, ln3;   // offsets (floats) into dec arena
  // tcgen05 path: the reference's own [N][K] layout, split into TF32 hi/lo planes (hi at off, lo at off + n)
  size_t p_qkv, p_wo, p_w1, p_w2;
};
struct DecW {
  DevBuf arena;
  DecLayerW L[9];
  size_t skip_t[4], bskip[4];
  size_t wv_t, bv, wco_t, bco;   // [9][128][128] / [9][128] cross-attention value + out projections
  size_t norm, final_t, bfinal, pe;
  size_t p_skip[4], p_final, p_pe;   // planes (hi | lo)
  bool ready = false;
};

// ---- VAE encoder weights (MotionPrior.encode, vae.py:154-214): tcgen05 path only -----------
struct EncLayerW {
  size_t p_qkv, bqkv, p_wo, bo, p_w1, b1, p_w2, b2, ln1, ln2;   // planes (hi | lo) and fp32 vectors
};
struct EncW {
  DevBuf arena;
  EncLayerW L[9];
  size_t p_skip[4], bskip[4];
  size_t norm, p_embed, bembed, gtok, pe;   // skel_embedding planes are [128][352] (K zero-padded from 333)
  bool ready = false;
};

struct DenW {
  DevBuf blob, misc;
  DevBuf blob2, vecs2;   // tensor-core loop kernel (denoise_tc.cu): fp16 hi/lo' weight stream + per-tile vectors
  // offsets into misc
  size_t freqs, w1t, b1, w2t, b2, condw[3], condb[3], pe, fnorm;
  bool ready = false;
};

}  // namespace

struct amuse_ctx {
  int device = 0;
  std::string err;
  std::unordered_map<std::string, HostTensor> raw;
  // which weight groups changed since the last finalize: a re-finalize repacks only those (the
  // training-time caller refreshes the denoiser every iteration while the VAE stays frozen, ldm.py:118-153)
  bool dirty_den = false, dirty_vae = false;
  std::vector<float> alphas_cumprod;   // 1000 entries; default computed at create
  DenW den;
  DecW dec;
  EncW enc;
  ast::Weights astw;
  std::map<std::tuple<int, int, int>, Schedule> schedules;   // (sampler, n_steps, eta bits)
  // workspaces
  DevBuf cond, lat_tmp, lat_out, one_coef;
  DevBuf dXA, dXB, dXC, dQKV, dO, dH, dSkip, dFeats, dCvec, dZero;
  DevBuf tX[3], tSkip, tO, tH;   // tcgen05 decoder path: activation planes (hi | lo halves)
  DevBuf eFeat, eEmb;            // encoder: packed feature planes [rows][352] x 2, embedded frames [rows][128]
  bool dec_use_tc = true;        // decoder GEMMs on tcgen05 (3xTF32); false = fp32 FFMA kernels
  int prune_last = 1;            // denoise loop: last layer for token 0 only (AMUSE_PRUNE_LAST=0 disables; tuning hook)
  int wide_rows = 0;             // denoise loop: 10-row GEMM warps in 2-clip clusters (AMUSE_WIDE_ROWS=0/1; tuning hook)
  // decoder pass as a CUDA graph: the ~100 launches of a 64-clip decode (each with up to 6 host-side cuTensorMapEncode
  // calls) are captured once per (batch, buffer set) and replayed with one cudaGraphLaunch (AMUSE_DECODE_GRAPH=0 disables)
  bool dec_graph = true;
  struct DecGraph {
    std::vector<uint64_t> key;
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
  };
  std::vector<DecGraph> dec_graphs;
  cudaStream_t cap_stream = nullptr;
  bool attn_ffma = false;        // MotionPrior self-attention on the fp32 kernel (decode_kernels.cu) instead of attn_tc.cu
  bool den_ffma = false;         // denoise loop on the fp32 FFMA2 kernel (denoise_loop.cu) instead of the tcgen05 one
                                 // (denoise_tc.cu): AMUSE_DENOISE_FFMA=1, A/B switch for measurements
  DevBuf h2d;   // staging for the *_host entry point
  DevBuf mel_t; // [257][128] mel filterbank weights (K-major)
  long long* d_prof = nullptr;
  int prof_step = -1;
  int64_t launches = 0;
  int dec_chunk = 32;   // clips per decoder pass (keeps the working set inside the 126 MB L2)
  // two decoder passes run concurrently on side streams: a 32-clip pass has 75 GEMM tiles, half the SMs
  static constexpr int kMaxLanes = 4;
  int dec_lanes = 2;
  cudaStream_t side[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
};

namespace {

// NVTX ranges around the phases of the hot path (tracing hooks); AMUSE_NVTX=1 switches them on
// (read once).  They show up in Nsight Systems / `ncu --nvtx` timelines: amuse.denoise, amuse.decode, amuse.ast, ...
struct NvtxRange {
  static bool enabled() {
    static const bool on = [] {
      const char* e = getenv("AMUSE_NVTX");
      return e && e[0] == '1';
    }();
    return on;
  }
  bool active;
  explicit NvtxRange(const char* name) : active(enabled()) {
    if (active) nvtxRangePushA(name);
  }
  ~NvtxRange() {
    if (active) nvtxRangePop();
  }
};

int fail(amuse_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) return fail(ctx, AMUSE_E_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

const HostTensor* find(amuse_ctx* c, const std::string& k) {
  auto it = c->raw.find(k);
  return it == c->raw.end() ? nullptr : &it->second;
}

// Default alphas_cumprod: scaled_linear betas (configs/diff_latent_v2.json:57-66) evaluated in
// double and rounded once to fp32.  torch's own fp32 linspace differs from this in the last bit
// of some entries (its vectorised kernel is build/CPU dependent), so the Python host passes the
// table torch produces on the host machine via "scheduler.alphas_cumprod" for exact parity.
void default_alphas(std::vector<float>& ac) {
  const int n = 1000;
  const double s = std::sqrt(0.00085), e = std::sqrt(0.012);
  ac.resize(n);
  double prod = 1.0;
  for (int i = 0; i < n; ++i) {
    const float b = static_cast<float>(s + (e - s) * i / (n - 1));
    const float beta = b * b;
    const float alpha = 1.0f - beta;
    prod *= static_cast<double>(alpha);   // torch CPU cumprod accumulates fp32 inputs in double
    ac[i] = static_cast<float>(prod);
  }
}

// Scheduler scalars in fp32, in the operation order of diffusers 0.17.1 `step()` (restated,
// SURVEY.md App. B.1 / B.2; oracle twin: oracle/lpdm_ref.py ddim_coeffs / ddpm_coeffs).
int build_schedule(amuse_ctx* ctx, int n_steps, int sampler, float eta, Schedule& sc) {
  const std::vector<float>& ac = ctx->alphas_cumprod;
  const int N = static_cast<int>(ac.size());
  if (n_steps < 1 || n_steps > N) return fail(ctx, AMUSE_E_INVALID, "n_steps %d out of range", n_steps);
  const int r = N / n_steps;
  sc.n_steps = n_steps;
  sc.timesteps.resize(n_steps);
  sc.coef.assign(static_cast<size_t>(n_steps) * 5, 0.f);
  sc.has_sigma = false;
  if (sampler == AMUSE_SAMPLER_DDIM) {
    sc.dir_uses_eps = true;
    for (int i = 0; i < n_steps; ++i) {
      const int t = (n_steps - 1 - i) * r + 1;   // steps_offset = 1 ("leading" spacing)
      if (t >= N)
        return fail(ctx, AMUSE_E_INVALID,
                    "DDIM with %d steps indexes alphas_cumprod[%d] (steps_offset=1): invalid in the reference too", n_steps, t);
      sc.timesteps[i] = t;
      const int p = t - r;
      const float a = ac[t], ap = (p >= 0) ? ac[p] : ac[0];   // set_alpha_to_one = False
      const float bp = 1.0f - a, bpp = 1.0f - ap;
      const float var = (bpp / bp) * (1.0f - a / ap);
      const float sd = eta * std::sqrt(var);
      float* c = &sc.coef[static_cast<size_t>(i) * 5];
      c[0] = std::sqrt(a);
      c[1] = std::sqrt(bp);
      c[2] = std::sqrt(ap);
      c[3] = std::sqrt(1.0f - ap - sd * sd);
      c[4] = sd;
      if (sd != 0.f) sc.has_sigma = true;
    }
  } else if (sampler == AMUSE_SAMPLER_DDPM) {
    sc.dir_uses_eps = false;
    for (int i = 0; i < n_steps; ++i) {
      const int t = (n_steps - 1 - i) * r;
      sc.timesteps[i] = t;
      const int p = t - r;
      const float a = ac[t], ap = (p >= 0) ? ac[p] : 1.0f;
      const float bp = 1.0f - a, bpp = 1.0f - ap;
      const float alpha_t = a / ap, beta_t = 1.0f - alpha_t;
      float* c = &sc.coef[static_cast<size_t>(i) * 5];
      c[0] = std::sqrt(a);
      c[1] = std::sqrt(bp);
      c[2] = (std::sqrt(ap) * beta_t) / bp;
      c[3] = std::sqrt(alpha_t) * bpp / bp;
      float var = (1.0f - ap) / (1.0f - a) * beta_t;   // fixed_small
      if (var < 1e-20f) var = 1e-20f;
      c[4] = (t > 0) ? std::sqrt(var) : 0.f;
      if (c[4] != 0.f) sc.has_sigma = true;
    }
  } else {
    return fail(ctx, AMUSE_E_INVALID, "unknown sampler %d", sampler);
  }
  return AMUSE_OK;
}

int get_schedule(amuse_ctx* ctx, int n_steps, int sampler, float eta, cudaStream_t st, Schedule** out) {
  uint32_t eb;
  std::memcpy(&eb, &eta, 4);
  auto key = std::make_tuple(sampler, n_steps, static_cast<int>(eb));
  auto it = ctx->schedules.find(key);
  if (it != ctx->schedules.end()) {
    *out = &it->second;
    return AMUSE_OK;
  }
  Schedule sc;
  int rc = build_schedule(ctx, n_steps, sampler, eta, sc);
  if (rc) return rc;
  CU(cudaMalloc(&sc.d_timesteps, sizeof(int) * n_steps));
  CU(cudaMalloc(&sc.d_coef, sizeof(float) * 5 * n_steps));
  CU(cudaMalloc(&sc.d_temb, sizeof(float) * 128 * n_steps));
  CU(cudaMemcpyAsync(sc.d_timesteps, sc.timesteps.data(), sizeof(int) * n_steps, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(sc.d_coef, sc.coef.data(), sizeof(float) * 5 * n_steps, cudaMemcpyHostToDevice, st));
  // a3: the time tokens of every step, once per schedule (batch-invariant)
  const float* m = ctx->den.misc.p;
  CU(launch_time_table(sc.d_timesteps, n_steps, m + ctx->den.freqs, m + ctx->den.w1t, m + ctx->den.b1,
                       m + ctx->den.w2t, m + ctx->den.b2, sc.d_temb, st));
  ctx->launches++;
  CU(cudaStreamSynchronize(st));   // host vectors above must outlive the async copies
  auto ins = ctx->schedules.emplace(key, std::move(sc));
  *out = &ins.first->second;
  return AMUSE_OK;
}

void drop_schedules(amuse_ctx* ctx) {
  for (auto& kv : ctx->schedules) {
    cudaFree(kv.second.d_timesteps);
    cudaFree(kv.second.d_coef);
    cudaFree(kv.second.d_temb);
  }
  ctx->schedules.clear();
}

// -------------------------------------------------------------------- denoiser packing
int need(amuse_ctx* ctx, const std::string& k, std::initializer_list<int64_t> shape, const HostTensor** out) {
  const HostTensor* t = find(ctx, k);
  if (!t) return fail(ctx, AMUSE_E_MISSING, "weight '%s' was not loaded", k.c_str());
  std::vector<int64_t> want(shape);
  if (t->shape != want) return fail(ctx, AMUSE_E_INVALID, "weight '%s' has the wrong shape", k.c_str());
  *out = t;
  return AMUSE_OK;
}
#define NEED(var, key, ...)                                  \
  const HostTensor* var = nullptr;                           \
  if (int rc__ = need(ctx, key, {__VA_ARGS__}, &var)) return rc__;

void transpose_into(float* dst, const float* src, int rows, int cols) {   // src [rows][cols] -> dst [cols][rows]
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) dst[static_cast<size_t>(c) * rows + r] = src[static_cast<size_t>(r) * cols + c];
}

int pack_denoiser(amuse_ctx* ctx, cudaStream_t st) {
  using namespace dn;
  const std::string P = "denoiser.";
  std::vector<float> blob(static_cast<size_t>(kCluster) * kBlobRankFloats, 0.f);
  for (int l = 0; l < 9; ++l) {
    const std::string b = P + "encoder." + kBlocks[l];
    NEED(inw, b + ".self_attn.in_proj_weight", 384, 128);
    NEED(inb, b + ".self_attn.in_proj_bias", 384);
    NEED(ow, b + ".self_attn.out_proj.weight", 128, 128);
    NEED(ob, b + ".self_attn.out_proj.bias", 128);
    NEED(w1, b + ".linear1.weight", 512, 128);
    NEED(b1, b + ".linear1.bias", 512);
    NEED(w2, b + ".linear2.weight", 128, 512);
    NEED(b2, b + ".linear2.bias", 128);
    NEED(n1w, b + ".norm1.weight", 128);
    NEED(n1b, b + ".norm1.bias", 128);
    NEED(n2w, b + ".norm2.weight", 128);
    NEED(n2b, b + ".norm2.bias", 128);
    const HostTensor *skw = nullptr, *skb = nullptr;
    if (l >= 5) {
      const std::string s = P + "encoder.linear_blocks." + std::to_string(l - 5);
      if (int rc = need(ctx, s + ".weight", {128, 256}, &skw)) return rc;
      if (int rc = need(ctx, s + ".bias", {128}, &skb)) return rc;
    }
    // Tiles are K-major and k-pair interleaved for the FFMA2 micro-kernel: element (k, p) of a
    // [K][NCOL] tile lives at ((k >> 1) * NCOL + p) * 2 + (k & 1).  For the 128-wide tiles the LDS.64
    // a lane issues at positions lane + 32*j (j = 0..3) must hand it the output columns
    // {2*lane, 2*lane+1, 64+2*lane, 64+2*lane+1}: logical column c sits at
    //   p = (c < 64) ? (c & 1) * 32 + c / 2  :  (2 + (c & 1)) * 32 + (c - 64) / 2.
    auto at = [](float* t, int ncol, int k, int c) -> float& {
      int p = c;
      if (ncol == 128) p = (c < 64) ? ((c & 1) * 32 + (c >> 1)) : ((2 + (c & 1)) * 32 + ((c - 64) >> 1));
      return t[((k >> 1) * ncol + p) * 2 + (k & 1)];
    };
    for (int rank = 0; rank < kCluster; ++rank) {   // rank == attention head owned by the CTA
      float* base = blob.data() + static_cast<size_t>(rank) * kBlobRankFloats;
      const int head = rank;
      int off, n;
      const int t0 = (l < 5) ? 4 * l : 20 + 5 * (l - 5) + 1;   // index of this layer's QKV tile
      if (l >= 5) {   // skip tile, K-split: W(kk, c) = Wsk[c][64*rank + kk]; + full bias
        tile_info(t0 - 1, off, n);
        float* t = base + off;
        for (int kk = 0; kk < 64; ++kk)
          for (int c = 0; c < 128; ++c) at(t, 128, kk, c) = skw->data[static_cast<size_t>(c) * 256 + rank * 64 + kk];
        std::memcpy(t + 64 * 128, skb->data.data(), 128 * 4);
      }
      {   // QKV tile of head `head`: local col j -> in_proj row  (j/32)*128 + head*32 + j%32
        tile_info(t0, off, n);
        float* t = base + off;
        for (int j = 0; j < 96; ++j) {
          const int row = (j / 32) * 128 + head * 32 + (j % 32);
          for (int k = 0; k < 128; ++k) at(t, 96, k, j) = inw->data[static_cast<size_t>(row) * 128 + k];
          t[128 * 96 + j] = inb->data[row];
        }
      }
      {   // out_proj columns of the head: W(kk, n) = Wo[n][head*32 + kk]; + bo + norm1
        tile_info(t0 + 1, off, n);
        float* t = base + off;
        for (int kk = 0; kk < 32; ++kk)
          for (int c = 0; c < 128; ++c) at(t, 128, kk, c) = ow->data[static_cast<size_t>(c) * 128 + head * 32 + kk];
        std::memcpy(t + 32 * 128, ob->data.data(), 128 * 4);
        std::memcpy(t + 32 * 128 + 128, n1w->data.data(), 128 * 4);
        std::memcpy(t + 32 * 128 + 256, n1b->data.data(), 128 * 4);
      }
      {   // linear1 rows [128*rank, +128): W(k, j) = W1[128*rank + j][k]
        tile_info(t0 + 2, off, n);
        float* t = base + off;
        for (int j = 0; j < 128; ++j) {
          for (int k = 0; k < 128; ++k) at(t, 128, k, j) = w1->data[static_cast<size_t>(rank * 128 + j) * 128 + k];
          t[128 * 128 + j] = b1->data[rank * 128 + j];   // tail: bias | (no LayerNorm: zeros)
        }
      }
      {   // linear2 columns [128*rank, +128): W(kk, n) = W2[n][128*rank + kk]; + b2 + norm2
        tile_info(t0 + 3, off, n);
        float* t = base + off;
        for (int kk = 0; kk < 128; ++kk)
          for (int c = 0; c < 128; ++c) at(t, 128, kk, c) = w2->data[static_cast<size_t>(c) * 512 + rank * 128 + kk];
        std::memcpy(t + 128 * 128, b2->data.data(), 128 * 4);
        std::memcpy(t + 128 * 128 + 128, n2w->data.data(), 128 * 4);
        std::memcpy(t + 128 * 128 + 256, n2b->data.data(), 128 * 4);
      }
    }
  }
  // ---- the tensor-core loop kernel's streams (denoise_tc.cuh): per rank, tiles in consumption order, each tile the
  //      fp16 hi / lo' planes of a [128 features x K] A operand in the producers' read order; + bias / LayerNorm vectors
  std::vector<uint32_t> blob2(static_cast<size_t>(dn2::kRanks) * dn2::kRankVec4 * 4, 0u);
  std::vector<float> vecs2(static_cast<size_t>(dn2::kRanks) * dn2::kRankVecFloats, 0.f);
  {
    auto put_tile = [](uint32_t* dst, int kind, auto&& Wfk) {   // Wfk(f, k): weight of output feature f, input feature k
      const int K = dn2::tile_K(kind), quads = dn2::tile_quads(kind), units = dn2::tile_units(kind);
      for (int u = 0; u < units; ++u)
        for (int q = 0; q < quads; ++q)
          for (int i = 0; i < 4; ++i)
            for (int lane = 0; lane < 32; ++lane)
              for (int w = 0; w < 4; ++w) {
                const int col = 16 * u + 4 * i + w, f = q * 32 + lane;
                const bool lo = col >= K / 2;
                const int k0 = 2 * (lo ? col - K / 2 : col);
                uint32_t word = 0;
                for (int e = 0; e < 2; ++e) {
                  uint16_t h, l;
                  dn2::split_fp16(Wfk(f, k0 + e), h, l);
                  word |= static_cast<uint32_t>(lo ? l : h) << (16 * e);
                }
                dst[((((static_cast<size_t>(u) * quads + q) * 4 + i) * 32) + lane) * 4 + w] = word;
              }
    };
    for (int l = 0; l < 9; ++l) {
      const std::string b = P + "encoder." + kBlocks[l];
      const HostTensor* inw = find(ctx, b + ".self_attn.in_proj_weight");
      const HostTensor* inb = find(ctx, b + ".self_attn.in_proj_bias");
      const HostTensor* ow = find(ctx, b + ".self_attn.out_proj.weight");
      const HostTensor* ob = find(ctx, b + ".self_attn.out_proj.bias");
      const HostTensor* w1 = find(ctx, b + ".linear1.weight");
      const HostTensor* b1 = find(ctx, b + ".linear1.bias");
      const HostTensor* w2 = find(ctx, b + ".linear2.weight");
      const HostTensor* b2 = find(ctx, b + ".linear2.bias");
      const HostTensor* n1w = find(ctx, b + ".norm1.weight");
      const HostTensor* n1b = find(ctx, b + ".norm1.bias");
      const HostTensor* n2w = find(ctx, b + ".norm2.weight");
      const HostTensor* n2b = find(ctx, b + ".norm2.bias");
      const HostTensor *skw = nullptr, *skb = nullptr;
      if (l >= 5) {
        skw = find(ctx, P + "encoder.linear_blocks." + std::to_string(l - 5) + ".weight");
        skb = find(ctx, P + "encoder.linear_blocks." + std::to_string(l - 5) + ".bias");
      }
      for (int rank = 0; rank < dn2::kRanks; ++rank) {
        // CTA rank / 2 streams its two weight ranks interleaved per stage: [stage][rank % 2] tiles, consumption order
        uint32_t* const cta = blob2.data() + static_cast<size_t>(rank / 2) * 2 * dn2::kRankVec4 * 4;
        auto wb_tile = [&](int kind_, int off_) {
          return cta + (static_cast<size_t>(2) * off_ + static_cast<size_t>(rank % 2) * dn2::tile_vec4(kind_)) * 4;
        };
        float* vb = vecs2.data() + static_cast<size_t>(rank) * dn2::kRankVecFloats;
        const int t0 = (l < 5) ? 4 * l : 20 + 5 * (l - 5) + 1;   // index of this layer's QKV tile
        int kind, off;
        if (l >= 5) {   // skip-linear, K-split: input features [64 rank, +64) of cat(x, skip)
          dn2::tile_info(t0 - 1, kind, off);
          put_tile(wb_tile(kind, off), kind,
                   [&](int f, int k) { return skw->data[static_cast<size_t>(f) * 256 + rank * 64 + k]; });
          std::memcpy(vb + (t0 - 1) * 384, skb->data.data(), 128 * 4);
        }
        dn2::tile_info(t0, kind, off);       // q | k | v rows of head `rank` (96 features)
        put_tile(wb_tile(kind, off), kind, [&](int f, int k) {
          return inw->data[static_cast<size_t>((f / 32) * 128 + rank * 32 + (f % 32)) * 128 + k];
        });
        for (int f = 0; f < 96; ++f) vb[t0 * 384 + f] = inb->data[(f / 32) * 128 + rank * 32 + (f % 32)];
        dn2::tile_info(t0 + 1, kind, off);   // out_proj, K-split by head
        put_tile(wb_tile(kind, off), kind,
                 [&](int f, int k) { return ow->data[static_cast<size_t>(f) * 128 + rank * 32 + k]; });
        std::memcpy(vb + (t0 + 1) * 384, ob->data.data(), 128 * 4);
        std::memcpy(vb + (t0 + 1) * 384 + 128, n1w->data.data(), 128 * 4);
        std::memcpy(vb + (t0 + 1) * 384 + 256, n1b->data.data(), 128 * 4);
        dn2::tile_info(t0 + 2, kind, off);   // linear1 rows [128 rank, +128)
        put_tile(wb_tile(kind, off), kind,
                 [&](int f, int k) { return w1->data[static_cast<size_t>(rank * 128 + f) * 128 + k]; });
        std::memcpy(vb + (t0 + 2) * 384, b1->data.data() + rank * 128, 128 * 4);
        dn2::tile_info(t0 + 3, kind, off);   // linear2, K-split over the rank's 128 hidden units
        put_tile(wb_tile(kind, off), kind,
                 [&](int f, int k) { return w2->data[static_cast<size_t>(f) * 512 + rank * 128 + k]; });
        std::memcpy(vb + (t0 + 3) * 384, b2->data.data(), 128 * 4);
        std::memcpy(vb + (t0 + 3) * 384 + 128, n2w->data.data(), 128 * 4);
        std::memcpy(vb + (t0 + 3) * 384 + 256, n2b->data.data(), 128 * 4);
      }
    }
  }
  // misc: time embedding (K-major), condition projections (K-major), PE, final norm, freqs
  NEED(tw1, P + "time_embedding.linear_1.weight", 128, 256);
  NEED(tb1, P + "time_embedding.linear_1.bias", 128);
  NEED(tw2, P + "time_embedding.linear_2.weight", 128, 128);
  NEED(tb2, P + "time_embedding.linear_2.bias", 128);
  NEED(pe, P + "query_pos.pe", 500, 1, 128);
  NEED(fnw, P + "encoder.norm.weight", 128);
  NEED(fnb, P + "encoder.norm.bias", 128);
  std::vector<float> misc;
  auto put = [&](size_t n) {
    size_t o = misc.size();
    misc.resize(o + ((n + 3) & ~size_t(3)), 0.f);
    return o;
  };
  DenW& d = ctx->den;
  d.freqs = put(128);
  if (const HostTensor* f = find(ctx, P + "time_proj.freqs")) {
    if (f->numel() != 128) return fail(ctx, AMUSE_E_INVALID, "denoiser.time_proj.freqs must have 128 entries");
    std::memcpy(&misc[d.freqs], f->data.data(), 128 * 4);
  } else {
    for (int i = 0; i < 128; ++i)   // exp(-ln(10000) * i / 128), embeddings.py:264-270 (freq_shift 0)
      misc[d.freqs + i] = static_cast<float>(std::exp(static_cast<double>(static_cast<float>(-std::log(10000.0)) *
                                                                         static_cast<float>(i) / 128.0f)));
  }
  d.w1t = put(256 * 128);
  transpose_into(&misc[d.w1t], tw1->data.data(), 128, 256);
  d.b1 = put(128);
  std::memcpy(&misc[d.b1], tb1->data.data(), 128 * 4);
  d.w2t = put(128 * 128);
  transpose_into(&misc[d.w2t], tw2->data.data(), 128, 128);
  d.b2 = put(128);
  std::memcpy(&misc[d.b2], tb2->data.data(), 128 * 4);
  const char* cn[3] = {"con", "emo", "sty"};
  for (int i = 0; i < 3; ++i) {
    NEED(cw, P + "emb_proj_" + cn[i] + ".1.weight", 128, 256);
    NEED(cb, P + "emb_proj_" + cn[i] + ".1.bias", 128);
    d.condw[i] = put(256 * 128);
    transpose_into(&misc[d.condw[i]], cw->data.data(), 128, 256);
    d.condb[i] = put(128);
    std::memcpy(&misc[d.condb[i]], cb->data.data(), 128 * 4);
  }
  d.pe = put(500 * 128);
  std::memcpy(&misc[d.pe], pe->data.data(), 500 * 128 * 4);
  d.fnorm = put(256);
  std::memcpy(&misc[d.fnorm], fnw->data.data(), 128 * 4);
  std::memcpy(&misc[d.fnorm + 128], fnb->data.data(), 128 * 4);

  CU(d.blob.ensure(blob.size()));
  CU(d.misc.ensure(misc.size()));
  CU(cudaMemcpyAsync(d.blob.p, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice, st));
  CU(d.blob2.ensure(blob2.size()));
  CU(d.vecs2.ensure(vecs2.size()));
  CU(cudaMemcpyAsync(d.blob2.p, blob2.data(), blob2.size() * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d.vecs2.p, vecs2.data(), vecs2.size() * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d.misc.p, misc.data(), misc.size() * 4, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  d.ready = true;
  drop_schedules(ctx);   // time tables depend on the weights
  return AMUSE_OK;
}

// -------------------------------------------------------------------- decoder packing
int pack_decoder(amuse_ctx* ctx, cudaStream_t st) {
  const std::string P = "vae.";
  std::vector<float> ar;
  auto put = [&](size_t n) {
    size_t o = ar.size();
    ar.resize(o + ((n + 3) & ~size_t(3)), 0.f);
    return o;
  };
  DecW& d = ctx->dec;
  d.wv_t = put(9 * 128 * 128);
  d.bv = put(9 * 128);
  d.wco_t = put(9 * 128 * 128);
  d.bco = put(9 * 128);
  for (int l = 0; l < 9; ++l) {
    const std::string b = P + "decoder." + kBlocks[l];
    NEED(inw, b + ".self_attn.in_proj_weight", 384, 128);
    NEED(inb, b + ".self_attn.in_proj_bias", 384);
    NEED(ow, b + ".self_attn.out_proj.weight", 128, 128);
    NEED(ob, b + ".self_attn.out_proj.bias", 128);
    NEED(cw, b + ".multihead_attn.in_proj_weight", 384, 128);
    NEED(cb, b + ".multihead_attn.in_proj_bias", 384);
    NEED(cow, b + ".multihead_attn.out_proj.weight", 128, 128);
    NEED(cob, b + ".multihead_attn.out_proj.bias", 128);
    NEED(w1, b + ".linear1.weight", 512, 128);
    NEED(b1, b + ".linear1.bias", 512);
    NEED(w2, b + ".linear2.weight", 128, 512);
    NEED(b2, b + ".linear2.bias", 128);
    DecLayerW& L = d.L[l];
    auto put_planes = [&](const float* w, size_t n) {
      const size_t o = put(2 * n);
      tc::split_host(w, &ar[o], &ar[o + n], n);
      return o;
    };
    L.p_qkv = put_planes(inw->data.data(), 384 * 128);
    L.p_wo = put_planes(ow->data.data(), 128 * 128);
    L.p_w1 = put_planes(w1->data.data(), 512 * 128);
    L.p_w2 = put_planes(w2->data.data(), 128 * 512);
    L.wqkv_t = put(128 * 384);
    transpose_into(&ar[L.wqkv_t], inw->data.data(), 384, 128);
    L.bqkv = put(384);
    std::memcpy(&ar[L.bqkv], inb->data.data(), 384 * 4);
    L.wo_t = put(128 * 128);
    transpose_into(&ar[L.wo_t], ow->data.data(), 128, 128);
    L.bo = put(128);
    std::memcpy(&ar[L.bo], ob->data.data(), 128 * 4);
    L.w1_t = put(128 * 512);
    transpose_into(&ar[L.w1_t], w1->data.data(), 512, 128);
    L.b1 = put(512);
    std::memcpy(&ar[L.b1], b1->data.data(), 512 * 4);
    L.w2_t = put(512 * 128);
    transpose_into(&ar[L.w2_t], w2->data.data(), 128, 512);
    L.b2 = put(128);
    std::memcpy(&ar[L.b2], b2->data.data(), 128 * 4);
    size_t* lns[3] = {&L.ln1, &L.ln2, &L.ln3};
    for (int q = 0; q < 3; ++q) {
      NEED(nw, b + ".norm" + std::to_string(q + 1) + ".weight", 128);
      NEED(nb, b + ".norm" + std::to_string(q + 1) + ".bias", 128);
      *lns[q] = put(256);
      std::memcpy(&ar[*lns[q]], nw->data.data(), 128 * 4);
      std::memcpy(&ar[*lns[q] + 128], nb->data.data(), 128 * 4);
    }
    // cross attention over a 1-token memory: only W_v (rows 256..383) and out_proj matter
    transpose_into(&ar[d.wv_t + static_cast<size_t>(l) * 128 * 128], cw->data.data() + 256 * 128, 128, 128);
    std::memcpy(&ar[d.bv + l * 128], cb->data.data() + 256, 128 * 4);
    transpose_into(&ar[d.wco_t + static_cast<size_t>(l) * 128 * 128], cow->data.data(), 128, 128);
    std::memcpy(&ar[d.bco + l * 128], cob->data.data(), 128 * 4);
  }
  for (int i = 0; i < 4; ++i) {
    const std::string s = P + "decoder.linear_blocks." + std::to_string(i);
    NEED(sw, s + ".weight", 128, 256);
    NEED(sb, s + ".bias", 128);
    d.p_skip[i] = put(2 * 128 * 256);
    tc::split_host(sw->data.data(), &ar[d.p_skip[i]], &ar[d.p_skip[i] + 128 * 256], 128 * 256);
    d.skip_t[i] = put(256 * 128);
    transpose_into(&ar[d.skip_t[i]], sw->data.data(), 128, 256);
    d.bskip[i] = put(128);
    std::memcpy(&ar[d.bskip[i]], sb->data.data(), 128 * 4);
  }
  NEED(nw, P + "decoder.norm.weight", 128);
  NEED(nb, P + "decoder.norm.bias", 128);
  d.norm = put(256);
  std::memcpy(&ar[d.norm], nw->data.data(), 128 * 4);
  std::memcpy(&ar[d.norm + 128], nb->data.data(), 128 * 4);
  NEED(fw, P + "final_layer.weight", kFeats, 128);
  NEED(fb, P + "final_layer.bias", kFeats);
  d.final_t = put(128 * 336);   // [128][336], columns 333..335 zero
  for (int k = 0; k < 128; ++k)
    for (int n = 0; n < kFeats; ++n) ar[d.final_t + static_cast<size_t>(k) * 336 + n] = fw->data[static_cast<size_t>(n) * 128 + k];
  d.p_final = put(2 * kFeats * 128);
  tc::split_host(fw->data.data(), &ar[d.p_final], &ar[d.p_final + kFeats * 128], static_cast<size_t>(kFeats) * 128);
  d.bfinal = put(336);
  std::memcpy(&ar[d.bfinal], fb->data.data(), kFeats * 4);
  NEED(pe, P + "query_pos_decoder.pe", 500, 1, 128);
  d.pe = put(500 * 128);
  std::memcpy(&ar[d.pe], pe->data.data(), 500 * 128 * 4);
  d.p_pe = put(2 * 500 * 128);
  tc::split_host(pe->data.data(), &ar[d.p_pe], &ar[d.p_pe + 500 * 128], 500 * 128);
  CU(d.arena.ensure(ar.size()));
  CU(cudaMemcpyAsync(d.arena.p, ar.data(), ar.size() * 4, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  d.ready = true;
  return AMUSE_OK;
}

// -------------------------------------------------------------------- encoder packing
int pack_encoder(amuse_ctx* ctx, cudaStream_t st) {
  const std::string P = "vae.";
  std::vector<float> ar;
  auto put = [&](size_t n) {
    size_t o = ar.size();
    ar.resize(o + ((n + 3) & ~size_t(3)), 0.f);
    return o;
  };
  auto put_planes = [&](const float* w, size_t n) {
    const size_t o = put(2 * n);
    tc::split_host(w, &ar[o], &ar[o + n], n);
    return o;
  };
  auto put_vec = [&](const float* v, size_t n) {
    const size_t o = put(n);
    std::memcpy(&ar[o], v, n * 4);
    return o;
  };
  EncW& d = ctx->enc;
  for (int l = 0; l < 9; ++l) {
    const std::string b = P + "encoder." + kBlocks[l];
    NEED(inw, b + ".self_attn.in_proj_weight", 384, 128);
    NEED(inb, b + ".self_attn.in_proj_bias", 384);
    NEED(ow, b + ".self_attn.out_proj.weight", 128, 128);
    NEED(ob, b + ".self_attn.out_proj.bias", 128);
    NEED(w1, b + ".linear1.weight", 512, 128);
    NEED(b1, b + ".linear1.bias", 512);
    NEED(w2, b + ".linear2.weight", 128, 512);
    NEED(b2, b + ".linear2.bias", 128);
    EncLayerW& L = d.L[l];
    L.p_qkv = put_planes(inw->data.data(), 384 * 128);
    L.bqkv = put_vec(inb->data.data(), 384);
    L.p_wo = put_planes(ow->data.data(), 128 * 128);
    L.bo = put_vec(ob->data.data(), 128);
    L.p_w1 = put_planes(w1->data.data(), 512 * 128);
    L.b1 = put_vec(b1->data.data(), 512);
    L.p_w2 = put_planes(w2->data.data(), 128 * 512);
    L.b2 = put_vec(b2->data.data(), 128);
    size_t* lns[2] = {&L.ln1, &L.ln2};
    for (int q = 0; q < 2; ++q) {
      NEED(nw, b + ".norm" + std::to_string(q + 1) + ".weight", 128);
      NEED(nb, b + ".norm" + std::to_string(q + 1) + ".bias", 128);
      *lns[q] = put(256);
      std::memcpy(&ar[*lns[q]], nw->data.data(), 128 * 4);
      std::memcpy(&ar[*lns[q] + 128], nb->data.data(), 128 * 4);
    }
  }
  for (int i = 0; i < 4; ++i) {
    const std::string sk = P + "encoder.linear_blocks." + std::to_string(i);
    NEED(sw, sk + ".weight", 128, 256);
    NEED(sb, sk + ".bias", 128);
    d.p_skip[i] = put_planes(sw->data.data(), 128 * 256);
    d.bskip[i] = put_vec(sb->data.data(), 128);
  }
  NEED(nw, P + "encoder.norm.weight", 128);
  NEED(nb, P + "encoder.norm.bias", 128);
  d.norm = put(256);
  std::memcpy(&ar[d.norm], nw->data.data(), 128 * 4);
  std::memcpy(&ar[d.norm + 128], nb->data.data(), 128 * 4);
  NEED(ew, P + "skel_embedding.weight", 128, kFeats);
  NEED(eb, P + "skel_embedding.bias", 128);
  {   // [128][333] -> [128][352], zero-padded K, then planes
    std::vector<float> wp(static_cast<size_t>(128) * 352, 0.f);
    for (int n = 0; n < 128; ++n)
      std::memcpy(&wp[static_cast<size_t>(n) * 352], ew->data.data() + static_cast<size_t>(n) * kFeats, kFeats * 4);
    d.p_embed = put_planes(wp.data(), wp.size());
  }
  d.bembed = put_vec(eb->data.data(), 128);
  NEED(gt, P + "global_motion_token", 2, 128);
  d.gtok = put_vec(gt->data.data(), 256);
  NEED(pe, P + "query_pos_encoder.pe", 500, 1, 128);
  d.pe = put_vec(pe->data.data(), 500 * 128);
  CU(d.arena.ensure(ar.size()));
  CU(cudaMemcpyAsync(d.arena.p, ar.data(), ar.size() * 4, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  d.ready = true;
  return AMUSE_OK;
}

bool any_with_prefix(amuse_ctx* ctx, const char* pre) {
  const size_t n = std::strlen(pre);
  for (auto& kv : ctx->raw)
    if (kv.first.compare(0, n, pre) == 0) return true;
  return false;
}

int choose_S(int B) {   // clips per 4-CTA cluster: one clip per cluster while all clusters stay co-resident
  return (B <= dn::kMaxClusters) ? 1 : dn::kSMax;
}

int run_denoise(amuse_ctx* ctx, int B, Schedule& sc, int clip, const float* latents0, const float* z_con,
                const float* z_emo, const float* z_sty, const float* step_noise, uint64_t seed,
                uint64_t elem_base, float* latents_out, cudaStream_t st) {
  DenW& d = ctx->den;
  CU(ctx->cond.ensure(static_cast<size_t>(B) * 3 * 128));
  CondArgs ca{};
  int nc = 0;
  const float* zs[3] = {z_con, z_emo, z_sty};
  for (int i = 0; i < 3; ++i) {
    if (!zs[i]) continue;
    ca.z[nc] = zs[i];
    ca.wt[nc] = d.misc.p + d.condw[i];
    ca.bias[nc] = d.misc.p + d.condb[i];
    ++nc;
  }
  ca.pe = d.misc.p + d.pe;
  ca.out = ctx->cond.p;
  CU(launch_cond_tokens(ca, B, nc, st));   // a4 (+a5): once per call, step-invariant
  ctx->launches++;

  if (!ctx->den_ffma) {
    dn2::Params p{};
    p.blob = reinterpret_cast<const uint4*>(d.blob2.p);
    p.vecs = d.vecs2.p;
    p.temb = sc.d_temb;
    p.cond = ctx->cond.p;
    p.pe01 = d.misc.p + d.pe;
    p.final_norm = d.misc.p + d.fnorm;
    p.latents0 = latents0;
    p.step_noise = step_noise;
    p.coef = sc.d_coef;
    p.latents_out = latents_out;
    p.prof = (ctx->prof_step >= 0) ? ctx->d_prof : nullptr;
    p.prof_step = ctx->prof_step;
    p.status = nullptr;
    p.B = B;
    p.T = 2 + nc;
    p.n_steps = sc.n_steps;
    p.dir_uses_eps = sc.dir_uses_eps ? 1 : 0;
    p.clip = clip;
    p.seed = seed;
    p.seed_elem_base = elem_base;
    p.prune_last = ctx->prune_last;
    if (const char* e = getenv("AMUSE_DN2_DEBUG")) p.debug_flags = atoi(e);
    CU(dn2::launch(p, st));
    ctx->launches++;
    ctx->prof_step = -1;
    return AMUSE_OK;
  }
  dn::Params p{};
  p.blob = d.blob.p;
  p.temb = sc.d_temb;
  p.cond = ctx->cond.p;
  p.pe01 = d.misc.p + d.pe;
  p.final_norm = d.misc.p + d.fnorm;
  p.latents0 = latents0;
  p.step_noise = step_noise;
  p.coef = sc.d_coef;
  p.latents_out = latents_out;
  p.prof = (ctx->prof_step >= 0) ? ctx->d_prof : nullptr;
  p.prof_step = ctx->prof_step;
  p.B = B;
  p.S = choose_S(B);
  p.T = 2 + nc;
  p.n_steps = sc.n_steps;
  p.dir_uses_eps = sc.dir_uses_eps ? 1 : 0;
  p.clip = clip;
  p.seed = seed;
  p.seed_elem_base = elem_base;
  p.prune_last = ctx->prune_last;
  p.wide_rows = ctx->wide_rows;
  CU(dn::launch(p, st));
  ctx->launches++;
  ctx->prof_step = -1;
  return AMUSE_OK;
}

int reserve_decode(amuse_ctx* ctx, int B) {
  const size_t Mc = static_cast<size_t>(std::min(B, ctx->dec_chunk)) * kFrames;
  CU(ctx->dXA.ensure(Mc * 128));
  CU(ctx->dXB.ensure(Mc * 128));
  CU(ctx->dXC.ensure(Mc * 128));
  CU(ctx->dQKV.ensure(Mc * 384));
  CU(ctx->dO.ensure(Mc * 128));
  CU(ctx->dH.ensure(Mc * 512));
  CU(ctx->dSkip.ensure(Mc * 128 * 4));
  CU(ctx->dFeats.ensure(Mc * kFeats));
  CU(ctx->dCvec.ensure(static_cast<size_t>(9) * B * 128));
  if (!ctx->dZero.p) {
    CU(ctx->dZero.ensure(128));
    CU(cudaMemset(ctx->dZero.p, 0, 128 * sizeof(float)));
  }
  return AMUSE_OK;
}

// MotionPrior.decode (vae.py:216-278) + 6D -> axis-angle (infer_ldm.py:165-174).
// Buffer plan per chunk of clips (no kernel ever reads and writes the same buffer):
//   layer input  cur : XB (l = 0: broadcast PE) | SK[l-1] (l = 1..4) | XC (l >= 5, skip-linear output)
//   after attention y: XA        layer output: SK[l] (l < 4, the skip stack) | XB (l >= 4)
int run_decode(amuse_ctx* ctx, int B, const float* latents, float* feats6d, float* poses, float* trans,
               cudaStream_t st) {
  using namespace dec;
  DecW& d = ctx->dec;
  const float* W = d.arena.p;
  if (int rc = reserve_decode(ctx, B)) return rc;
  const int chunk = ctx->dec_chunk;
  const size_t Mc = static_cast<size_t>(std::min(B, chunk)) * kFrames;
  float* XA = ctx->dXA.p;
  float* XB = ctx->dXB.p;
  float* XC = ctx->dXC.p;
  auto SK = [&](int i) { return ctx->dSkip.p + static_cast<size_t>(i) * Mc * 128; };

  // collapsed 1-key cross attention of all 9 layers for the whole batch
  CU(launch_cross_vectors(latents, W + d.wv_t, W + d.bv, W + d.wco_t, W + d.bco, ctx->dCvec.p, B, st));
  ctx->launches++;

  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    const int M = nb * kFrames;
    CU(launch_broadcast_rows(W + d.pe, XB, nb, kFrames, st));   // queries = zeros + pe[:300] (vae.py:221,253)
    ctx->launches++;
    const float* cur = XB;
    for (int l = 0; l < 9; ++l) {
      const DecLayerW& L = d.L[l];
      GemmArgs g{};
      if (l >= 5) {   // x = Linear(cat(x, xs.pop()))  (cross_attention.py:115-117)
        g.A = cur; g.lda = 128;
        g.A2 = SK(8 - l); g.lda2 = 128;
        g.Wt = W + d.skip_t[l - 5]; g.ldw = 128; g.bias = W + d.bskip[l - 5];
        g.C = XC; g.ldc = 128; g.M = M; g.N = 128; g.K = 256;
        CU(launch_gemm(EPI_BIAS, g, st));
        ctx->launches++;
        cur = XC;
      }
      // self attention: packed in_proj (q pre-scaled), per-head softmax(q k^T) v
      g = GemmArgs{};
      g.A = cur; g.lda = 128; g.Wt = W + L.wqkv_t; g.ldw = 384; g.bias = W + L.bqkv;
      g.C = ctx->dQKV.p; g.ldc = 384; g.M = M; g.N = 384; g.K = 128;
      CU(launch_gemm(EPI_QKV, g, st));
      CU(launch_self_attention(ctx->dQKV.p, ctx->dO.p, nb, kFrames, st));
      // y = norm2(norm1(x + out_proj(o)) + cross_vector)
      g = GemmArgs{};
      g.A = ctx->dO.p; g.lda = 128; g.Wt = W + L.wo_t; g.ldw = 128; g.bias = W + L.bo;
      g.C = XA; g.ldc = 128; g.M = M; g.N = 128; g.K = 128;
      g.R = cur; g.ldr = 128; g.ln_g = W + L.ln1; g.ln_b = W + L.ln1 + 128;
      g.cvec = ctx->dCvec.p + (static_cast<size_t>(l) * B + b0) * 128;
      g.ln2_g = W + L.ln2; g.ln2_b = W + L.ln2 + 128; g.rows_per_clip = kFrames;
      CU(launch_gemm(EPI_RES_LN_CROSS_LN, g, st));
      // h = gelu(linear1(y));  out = norm3(y + linear2(h))   [+ decoder.norm after the last block]
      g = GemmArgs{};
      g.A = XA; g.lda = 128; g.Wt = W + L.w1_t; g.ldw = 512; g.bias = W + L.b1;
      g.C = ctx->dH.p; g.ldc = 512; g.M = M; g.N = 512; g.K = 128;
      CU(launch_gemm(EPI_GELU, g, st));
      float* out = (l < 4) ? SK(l) : XB;
      g = GemmArgs{};
      g.A = ctx->dH.p; g.lda = 512; g.Wt = W + L.w2_t; g.ldw = 128; g.bias = W + L.b2;
      g.C = out; g.ldc = 128; g.M = M; g.N = 128; g.K = 512;
      g.R = XA; g.ldr = 128; g.ln_g = W + L.ln3; g.ln_b = W + L.ln3 + 128;
      if (l == 8) {   // SkipTransformerDecoder.norm (cross_attention.py:122-123) folded in: LN(LN3(..) + 0)
        g.cvec = ctx->dZero.p; g.rows_per_clip = 1 << 30;
        g.ln2_g = W + d.norm; g.ln2_b = W + d.norm + 128;
        CU(launch_gemm(EPI_RES_LN_CROSS_LN, g, st));
      } else {
        CU(launch_gemm(EPI_RES_LN, g, st));
      }
      ctx->launches += 5;
      cur = out;
    }
    // final_layer 128 -> 333 (vae.py:272), then 6D -> axis-angle + trans
    float* feats = feats6d ? feats6d + static_cast<size_t>(b0) * kFrames * kFeats : ctx->dFeats.p;
    GemmArgs g{};
    g.A = cur; g.lda = 128; g.Wt = W + d.final_t; g.ldw = 336; g.bias = W + d.bfinal;
    g.C = feats; g.ldc = kFeats; g.M = M; g.N = kFeats; g.K = 128;
    CU(launch_gemm(EPI_BIAS, g, st));
    ctx->launches++;
    if (poses) {
      CU(launch_rot6d(feats, kFeats, static_cast<long long>(M), poses + static_cast<size_t>(b0) * kFrames * 165,
                      trans ? trans + static_cast<size_t>(b0) * kFrames * 3 : nullptr, st));
      ctx->launches++;
    }
  }
  return AMUSE_OK;
}

// MotionPrior.decode on the tensor cores: every GEMM is a tcgen05 3xTF32 launch (tc_gemm.cu) with
// the bias / q-scale / GELU / residual + LayerNorm (+ collapsed cross-attention + LayerNorm)
// epilogues fused; activations travel between kernels as TF32 hi/lo planes (value = hi + lo).
// Same buffer plan as run_decode: layer input cur = XB | SK[l-1] | XC, after attention XA, output SK[l] | XB.
int run_decode_tc_body(amuse_ctx* ctx, int B, const float* latents, float* feats6d, float* poses, float* trans,
                       cudaStream_t st, bool prepare_only) {
  using namespace dec;
  DecW& d = ctx->dec;
  const float* W = d.arena.p;
  const int chunk = ctx->dec_chunk;
  const size_t Mc = static_cast<size_t>(std::min(B, chunk)) * kFrames;
  const size_t n128 = Mc * 128;
  // lanes: passes in flight at once, each with its own workspace
  const int n_pass = (B + chunk - 1) / chunk;
  const int NL = std::max(1, std::min(std::min(ctx->dec_lanes, n_pass), static_cast<int>(amuse_ctx::kMaxLanes)));
  for (int i = 0; i < 3; ++i) CU(ctx->tX[i].ensure(NL * 2 * n128));
  CU(ctx->tSkip.ensure(NL * 2 * n128 * 4));
  CU(ctx->tO.ensure(NL * 2 * n128));
  CU(ctx->tH.ensure(NL * 2 * Mc * 512));
  CU(ctx->dQKV.ensure(NL * Mc * 384));
  CU(ctx->dFeats.ensure(NL * Mc * kFeats));
  if (NL > 1 && !ctx->ev_fork) {
    for (int i = 0; i < amuse_ctx::kMaxLanes; ++i) {
      CU(cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  }
  CU(ctx->dCvec.ensure(static_cast<size_t>(9) * B * 128));
  if (!ctx->dZero.p) {
    CU(ctx->dZero.ensure(128));
    CU(cudaMemset(ctx->dZero.p, 0, 128 * sizeof(float)));
  }
  if (prepare_only) return AMUSE_OK;   // every allocation / stream / event exists: the rest only enqueues work
  struct P {
    float* hi;
    float* lo;
  };
  auto planes = [&](DevBuf& b, size_t n, size_t idx = 0) { return P{b.p + idx * 2 * n, b.p + idx * 2 * n + n}; };

  cudaStream_t const caller = st;
  CU(launch_cross_vectors(latents, W + d.wv_t, W + d.bv, W + d.wco_t, W + d.bco, ctx->dCvec.p, B, caller));
  ctx->launches++;
  if (NL > 1) {   // fork: the side streams start after everything queued on the caller's stream so far
    CU(cudaEventRecord(ctx->ev_fork, caller));
    for (int i = 0; i < NL; ++i) CU(cudaStreamWaitEvent(ctx->side[i], ctx->ev_fork, 0));
  }

  int pass = 0;
  for (int b0 = 0; b0 < B; b0 += chunk, ++pass) {
    const int nb = std::min(chunk, B - b0);
    const int M = nb * kFrames;
    const size_t lane = static_cast<size_t>(pass % NL);
    st = (NL > 1) ? ctx->side[lane] : caller;
    const P XA = planes(ctx->tX[0], n128, lane), XB = planes(ctx->tX[1], n128, lane), XC = planes(ctx->tX[2], n128, lane);
    const P O = planes(ctx->tO, n128, lane), H = planes(ctx->tH, Mc * 512, lane);
    auto SK = [&](int i) { return planes(ctx->tSkip, n128, lane * 4 + static_cast<size_t>(i)); };
    float* const qkv_ws = ctx->dQKV.p + lane * Mc * 384;
    // queries = zeros + pe[:300] (vae.py:221,253), as planes
    CU(launch_broadcast_rows(W + d.p_pe, XB.hi, nb, kFrames, st));
    CU(launch_broadcast_rows(W + d.p_pe + 500 * 128, XB.lo, nb, kFrames, st));
    ctx->launches += 2;
    P cur = XB;
    for (int l = 0; l < 9; ++l) {
      const DecLayerW& L = d.L[l];
      tc::GemmDesc g{};
      if (l >= 5) {   // x = Linear(cat(x, xs.pop()))
        const P sk = SK(8 - l);
        g.A_hi = cur.hi; g.A_lo = cur.lo; g.lda = 128;
        g.A2_hi = sk.hi; g.A2_lo = sk.lo; g.lda2 = 128; g.k_split = 128;
        g.W_hi = W + d.p_skip[l - 5]; g.W_lo = g.W_hi + 128 * 256; g.ldw = 256;
        g.M = M; g.N = 128; g.K = 256; g.bias = W + d.bskip[l - 5];
        g.C_hi = XC.hi; g.C_lo = XC.lo; g.ldc = 128;
        CU(tc::gemm(tc::EPI_PLANES, g, st));
        ctx->launches++;
        cur = XC;
      }
      // packed in_proj -> q (pre-scaled) | k | v, plain fp32 for the attention kernel
      g = tc::GemmDesc{};
      g.A_hi = cur.hi; g.A_lo = cur.lo; g.lda = 128;
      g.W_hi = W + L.p_qkv; g.W_lo = g.W_hi + 384 * 128; g.ldw = 128;
      g.M = M; g.N = 384; g.K = 128; g.bias = W + L.bqkv;
      g.C = qkv_ws; g.ldc = 384; g.q_cols = 128; g.q_scale = 0.17677669529663687f;
      CU(tc::gemm(tc::EPI_QKV, g, st));
      if (ctx->attn_ffma) CU(launch_self_attention_planes(qkv_ws, O.hi, O.lo, nb, kFrames, st));
      else CU(launch_self_attention_tc(qkv_ws, O.hi, O.lo, nb, kFrames, st));
      // y = norm2(norm1(x + out_proj(o)) + cross_vector)
      g = tc::GemmDesc{};
      g.A_hi = O.hi; g.A_lo = O.lo; g.lda = 128;
      g.W_hi = W + L.p_wo; g.W_lo = g.W_hi + 128 * 128; g.ldw = 128;
      g.M = M; g.N = 128; g.K = 128; g.bias = W + L.bo;
      g.R_hi = cur.hi; g.R_lo = cur.lo; g.ldr = 128;
      g.ln_g = W + L.ln1; g.ln_b = W + L.ln1 + 128; g.ln2_g = W + L.ln2; g.ln2_b = W + L.ln2 + 128;
      g.cvec = ctx->dCvec.p + (static_cast<size_t>(l) * B + b0) * 128; g.rows_per_clip = kFrames;
      g.C_hi = XA.hi; g.C_lo = XA.lo; g.ldc = 128;
      CU(tc::gemm(tc::EPI_RES_LN_CROSS_LN_PLANES, g, st));
      // h = gelu(linear1(y))
      g = tc::GemmDesc{};
      g.A_hi = XA.hi; g.A_lo = XA.lo; g.lda = 128;
      g.W_hi = W + L.p_w1; g.W_lo = g.W_hi + 512 * 128; g.ldw = 128;
      g.M = M; g.N = 512; g.K = 128; g.bias = W + L.b1;
      g.C_hi = H.hi; g.C_lo = H.lo; g.ldc = 512;
      CU(tc::gemm(tc::EPI_GELU_PLANES, g, st));
      // out = norm3(y + linear2(h))  [+ decoder.norm after the last block]
      const P out = (l < 4) ? SK(l) : XB;
      g = tc::GemmDesc{};
      g.A_hi = H.hi; g.A_lo = H.lo; g.lda = 512;
      g.W_hi = W + L.p_w2; g.W_lo = g.W_hi + 128 * 512; g.ldw = 512;
      g.M = M; g.N = 128; g.K = 512; g.bias = W + L.b2;
      g.R_hi = XA.hi; g.R_lo = XA.lo; g.ldr = 128;
      g.ln_g = W + L.ln3; g.ln_b = W + L.ln3 + 128;
      g.C_hi = out.hi; g.C_lo = out.lo; g.ldc = 128;
      if (l == 8) {
        g.cvec = ctx->dZero.p; g.rows_per_clip = 1 << 30;
        g.ln2_g = W + d.norm; g.ln2_b = W + d.norm + 128;
        CU(tc::gemm(tc::EPI_RES_LN_CROSS_LN_PLANES, g, st));
      } else {
        CU(tc::gemm(tc::EPI_RES_LN_PLANES, g, st));
      }
      ctx->launches += 5;
      cur = out;
    }
    float* feats = feats6d ? feats6d + static_cast<size_t>(b0) * kFrames * kFeats : ctx->dFeats.p + lane * Mc * kFeats;
    tc::GemmDesc g{};
    g.A_hi = cur.hi; g.A_lo = cur.lo; g.lda = 128;
    g.W_hi = W + d.p_final; g.W_lo = g.W_hi + kFeats * 128; g.ldw = 128;
    g.M = M; g.N = kFeats; g.K = 128; g.bias = W + d.bfinal;
    g.C = feats; g.ldc = kFeats;
    CU(tc::gemm(tc::EPI_PLAIN, g, st));
    ctx->launches++;
    if (poses) {
      CU(launch_rot6d(feats, kFeats, static_cast<long long>(M), poses + static_cast<size_t>(b0) * kFrames * 165,
                      trans ? trans + static_cast<size_t>(b0) * kFrames * 3 : nullptr, st));
      ctx->launches++;
    }
  }
  if (NL > 1) {   // join: the caller's stream continues after every lane
    for (int i = 0; i < NL; ++i) {
      CU(cudaEventRecord(ctx->ev_join[i], ctx->side[i]));
      CU(cudaStreamWaitEvent(caller, ctx->ev_join[i], 0));
    }
  }
  return AMUSE_OK;
}

int run_decode_tc(amuse_ctx* ctx, int B, const float* latents, float* feats6d, float* poses, float* trans,
                  cudaStream_t st) {
  if (int rc = run_decode_tc_body(ctx, B, latents, feats6d, poses, trans, st, true)) return rc;
  if (!ctx->dec_graph) return run_decode_tc_body(ctx, B, latents, feats6d, poses, trans, st, false);
  // the graph bakes every pointer in: key = batch + plan + caller buffers + workspace / weight base addresses
  auto u = [](const void* p) { return static_cast<uint64_t>(reinterpret_cast<uintptr_t>(p)); };
  const std::vector<uint64_t> key = {static_cast<uint64_t>(B), static_cast<uint64_t>(ctx->dec_chunk), static_cast<uint64_t>(ctx->dec_lanes),
                                     u(latents), u(feats6d), u(poses), u(trans), u(ctx->dec.arena.p), u(ctx->tX[0].p),
                                     u(ctx->tX[1].p), u(ctx->tX[2].p), u(ctx->tSkip.p), u(ctx->tO.p), u(ctx->tH.p),
                                     u(ctx->dQKV.p), u(ctx->dFeats.p), u(ctx->dCvec.p), u(ctx->dZero.p)};
  for (auto& g : ctx->dec_graphs)
    if (g.key == key) {
      CU(cudaGraphLaunch(g.exec, st));
      ctx->launches += g.launches;
      return AMUSE_OK;
    }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (cs != cudaStreamCaptureStatusNone)   // the caller is capturing already: just enqueue into its graph
    return run_decode_tc_body(ctx, B, latents, feats6d, poses, trans, st, false);
  const int64_t l0 = ctx->launches;
  // captured on a stream of our own (the caller's may be the legacy default stream, which cannot be captured); the
  // instantiated graph is then launched on the caller's stream
  if (!ctx->cap_stream) CU(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
  CU(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal));
  const int rc = run_decode_tc_body(ctx, B, latents, feats6d, poses, trans, ctx->cap_stream, false);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &graph);
  if (rc != AMUSE_OK || ce != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    ctx->launches = l0;
    if (rc != AMUSE_OK) return rc;
    ctx->dec_graph = false;   // capture is unavailable on this stream: enqueue the launches directly from now on
    return run_decode_tc_body(ctx, B, latents, feats6d, poses, trans, st, false);
  }
  amuse_ctx::DecGraph g;
  g.key = key;
  g.launches = ctx->launches - l0;
  const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    cudaGetLastError();
    ctx->dec_graph = false;
    ctx->launches = l0;
    return run_decode_tc_body(ctx, B, latents, feats6d, poses, trans, st, false);
  }
  if (ctx->dec_graphs.size() >= 8) {   // callers cycle through a handful of buffer sets at most
    cudaGraphExecDestroy(ctx->dec_graphs.front().exec);
    ctx->dec_graphs.erase(ctx->dec_graphs.begin());
  }
  ctx->dec_graphs.push_back(g);
  CU(cudaGraphLaunch(g.exec, st));
  return AMUSE_OK;
}

// MotionPrior.encode (vae.py:154-214) on the same tcgen05 GEMMs as the decoder: skel_embedding
// (K padded 333 -> 352), [2 global tokens | 300 frames] + learned PE, 9 post-LN encoder layers with
// U-Net skips over 302 tokens, encoder.norm, then mu = token 0, logvar = token 1.
int run_encode_tc(amuse_ctx* ctx, int B, const float* feats, float* mu, float* logvar, cudaStream_t st) {
  using namespace dec;
  constexpr int T = 302;
  EncW& d = ctx->enc;
  const float* W = d.arena.p;
  const int chunk = ctx->dec_chunk;
  const size_t Mc = static_cast<size_t>(std::min(B, chunk)) * T;
  const size_t n128 = Mc * 128;
  for (int i = 0; i < 3; ++i) CU(ctx->tX[i].ensure(2 * n128));
  CU(ctx->tSkip.ensure(2 * n128 * 4));
  CU(ctx->tO.ensure(2 * n128));
  CU(ctx->tH.ensure(2 * Mc * 512));
  CU(ctx->dQKV.ensure(Mc * 384));
  CU(ctx->eFeat.ensure(2 * Mc * 352));
  CU(ctx->eEmb.ensure(n128));
  if (!ctx->dZero.p) {
    CU(ctx->dZero.ensure(128));
    CU(cudaMemset(ctx->dZero.p, 0, 128 * sizeof(float)));
  }
  struct P {
    float* hi;
    float* lo;
  };
  auto planes = [&](DevBuf& b, size_t n, size_t idx = 0) { return P{b.p + idx * 2 * n, b.p + idx * 2 * n + n}; };
  const P XA = planes(ctx->tX[0], n128), XB = planes(ctx->tX[1], n128), XC = planes(ctx->tX[2], n128);
  const P O = planes(ctx->tO, n128), H = planes(ctx->tH, Mc * 512);
  auto SK = [&](int i) { return planes(ctx->tSkip, n128, static_cast<size_t>(i)); };

  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    const int Mf = nb * kFrames, M = nb * T;
    // skel_embedding on the 300 frames of every clip (vae.py:169)
    float* fh = ctx->eFeat.p;
    float* fl = ctx->eFeat.p + static_cast<size_t>(Mf) * 352;
    CU(launch_pack_feats(feats + static_cast<size_t>(b0) * kFrames * kFeats, Mf, fh, fl, st));
    tc::GemmDesc g{};
    g.A_hi = fh; g.A_lo = fl; g.lda = 352;
    g.W_hi = W + d.p_embed; g.W_lo = g.W_hi + 128 * 352; g.ldw = 352;
    g.M = Mf; g.N = 128; g.K = 352; g.bias = W + d.bembed;
    g.C = ctx->eEmb.p; g.ldc = 128;
    CU(tc::gemm(tc::EPI_PLAIN, g, st));
    // xseq = cat(global_motion_token, x) + query_pos_encoder.pe[:302]  (vae.py:176-191)
    CU(launch_encoder_tokens(ctx->eEmb.p, W + d.gtok, W + d.pe, nb, XB.hi, XB.lo, st));
    ctx->launches += 3;
    P cur = XB;
    for (int l = 0; l < 9; ++l) {
      const EncLayerW& L = d.L[l];
      if (l >= 5) {   // x = Linear(cat(x, xs.pop()))   (cross_attention.py:54-57)
        const P sk = SK(8 - l);
        g = tc::GemmDesc{};
        g.A_hi = cur.hi; g.A_lo = cur.lo; g.lda = 128;
        g.A2_hi = sk.hi; g.A2_lo = sk.lo; g.lda2 = 128; g.k_split = 128;
        g.W_hi = W + d.p_skip[l - 5]; g.W_lo = g.W_hi + 128 * 256; g.ldw = 256;
        g.M = M; g.N = 128; g.K = 256; g.bias = W + d.bskip[l - 5];
        g.C_hi = XC.hi; g.C_lo = XC.lo; g.ldc = 128;
        CU(tc::gemm(tc::EPI_PLANES, g, st));
        ctx->launches++;
        cur = XC;
      }
      g = tc::GemmDesc{};
      g.A_hi = cur.hi; g.A_lo = cur.lo; g.lda = 128;
      g.W_hi = W + L.p_qkv; g.W_lo = g.W_hi + 384 * 128; g.ldw = 128;
      g.M = M; g.N = 384; g.K = 128; g.bias = W + L.bqkv;
      g.C = ctx->dQKV.p; g.ldc = 384; g.q_cols = 128; g.q_scale = 0.17677669529663687f;
      CU(tc::gemm(tc::EPI_QKV, g, st));
      if (ctx->attn_ffma) CU(launch_self_attention_planes(ctx->dQKV.p, O.hi, O.lo, nb, T, st));
      else CU(launch_self_attention_tc(ctx->dQKV.p, O.hi, O.lo, nb, T, st));
      // y = norm1(x + out_proj(o))
      g = tc::GemmDesc{};
      g.A_hi = O.hi; g.A_lo = O.lo; g.lda = 128;
      g.W_hi = W + L.p_wo; g.W_lo = g.W_hi + 128 * 128; g.ldw = 128;
      g.M = M; g.N = 128; g.K = 128; g.bias = W + L.bo;
      g.R_hi = cur.hi; g.R_lo = cur.lo; g.ldr = 128;
      g.ln_g = W + L.ln1; g.ln_b = W + L.ln1 + 128;
      g.C_hi = XA.hi; g.C_lo = XA.lo; g.ldc = 128;
      CU(tc::gemm(tc::EPI_RES_LN_PLANES, g, st));
      // h = gelu(linear1(y))
      g = tc::GemmDesc{};
      g.A_hi = XA.hi; g.A_lo = XA.lo; g.lda = 128;
      g.W_hi = W + L.p_w1; g.W_lo = g.W_hi + 512 * 128; g.ldw = 128;
      g.M = M; g.N = 512; g.K = 128; g.bias = W + L.b1;
      g.C_hi = H.hi; g.C_lo = H.lo; g.ldc = 512;
      CU(tc::gemm(tc::EPI_GELU_PLANES, g, st));
      // out = norm2(y + linear2(h))  [+ encoder.norm after the last block]
      const P out = (l < 4) ? SK(l) : XB;
      g = tc::GemmDesc{};
      g.A_hi = H.hi; g.A_lo = H.lo; g.lda = 512;
      g.W_hi = W + L.p_w2; g.W_lo = g.W_hi + 128 * 512; g.ldw = 512;
      g.M = M; g.N = 128; g.K = 512; g.bias = W + L.b2;
      g.R_hi = XA.hi; g.R_lo = XA.lo; g.ldr = 128;
      g.ln_g = W + L.ln2; g.ln_b = W + L.ln2 + 128;
      g.C_hi = out.hi; g.C_lo = out.lo; g.ldc = 128;
      if (l == 8) {
        g.cvec = ctx->dZero.p; g.rows_per_clip = 1 << 30;
        g.ln2_g = W + d.norm; g.ln2_b = W + d.norm + 128;
        CU(tc::gemm(tc::EPI_RES_LN_CROSS_LN_PLANES, g, st));
      } else {
        CU(tc::gemm(tc::EPI_RES_LN_PLANES, g, st));
      }
      ctx->launches += 5;
      cur = out;
    }
    CU(launch_encoder_dist(cur.hi, cur.lo, nb, mu + static_cast<size_t>(b0) * 128,
                           logvar + static_cast<size_t>(b0) * 128, st));
    ctx->launches++;
  }
  return AMUSE_OK;
}

int run_decode_any(amuse_ctx* ctx, int B, const float* latents, float* feats6d, float* poses, float* trans,
                   cudaStream_t st) {
  return ctx->dec_use_tc ? run_decode_tc(ctx, B, latents, feats6d, poses, trans, st)
                         : run_decode(ctx, B, latents, feats6d, poses, trans, st);
}

int check_ready(amuse_ctx* ctx, bool den, bool dec) {
  if (!ctx) return AMUSE_E_INVALID;
  if (den && !ctx->den.ready) return fail(ctx, AMUSE_E_STATE, "denoiser weights not finalized");
  if (dec && !ctx->dec.ready) return fail(ctx, AMUSE_E_STATE, "vae decoder weights not finalized");
  return AMUSE_OK;
}

}  // namespace

// =========================================================================== extern "C"
extern "C" {

const char* amuse_version(void) { return "amuse_b200 0.1 (sm_100a)"; }

int amuse_create(amuse_ctx** out, int device_ordinal) {
  if (!out) return AMUSE_E_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device_ordinal < 0 || device_ordinal >= n) return AMUSE_E_CUDA;
  if (cudaSetDevice(device_ordinal) != cudaSuccess) return AMUSE_E_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) return AMUSE_E_CUDA;
  if (prop.major != 10) return AMUSE_E_UNSUPPORTED;   // sm_100a code only
  amuse_ctx* c = new amuse_ctx();
  c->device = device_ordinal;
  default_alphas(c->alphas_cumprod);
  if (const char* e = getenv("AMUSE_DECODE_FFMA")) c->dec_use_tc = !(e[0] == '1');   // A/B switch for measurements
  if (const char* e = getenv("AMUSE_PRUNE_LAST")) c->prune_last = (e[0] != '0');
  if (const char* e = getenv("AMUSE_DENOISE_FFMA")) c->den_ffma = (e[0] == '1');
  if (const char* e = getenv("AMUSE_ATTN_FFMA")) c->attn_ffma = (e[0] == '1');
  if (const char* e = getenv("AMUSE_DECODE_GRAPH")) c->dec_graph = (e[0] != '0');
  if (const char* e = getenv("AMUSE_WIDE_ROWS")) c->wide_rows = (e[0] != '0');
  if (cudaMalloc(&c->d_prof, 512 * sizeof(long long)) != cudaSuccess) {
    delete c;
    return AMUSE_E_CUDA;
  }
  cudaMemset(c->d_prof, 0, 512 * sizeof(long long));
  *out = c;
  return AMUSE_OK;
}

void amuse_destroy(amuse_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  drop_schedules(ctx);
  for (auto& g : ctx->dec_graphs) cudaGraphExecDestroy(g.exec);
  ctx->dec_graphs.clear();
  if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
  DevBuf* bufs[] = {&ctx->den.blob, &ctx->den.misc, &ctx->den.blob2, &ctx->den.vecs2, &ctx->dec.arena, &ctx->cond, &ctx->lat_tmp, &ctx->lat_out,
                    &ctx->one_coef, &ctx->dXA, &ctx->dXB, &ctx->dXC, &ctx->dQKV, &ctx->dO, &ctx->dH, &ctx->dSkip,
                    &ctx->dFeats, &ctx->dCvec, &ctx->dZero, &ctx->h2d, &ctx->tX[0], &ctx->tX[1], &ctx->tX[2],
                    &ctx->tSkip, &ctx->tO, &ctx->tH, &ctx->mel_t, &ctx->enc.arena, &ctx->eFeat, &ctx->eEmb};
  for (DevBuf* b : bufs) b->release();
  ast::release(ctx->astw);
  for (int i = 0; i < amuse_ctx::kMaxLanes; ++i) {
    if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
    if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->d_prof) cudaFree(ctx->d_prof);
  delete ctx;
}

const char* amuse_last_error(amuse_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int amuse_load_weights(amuse_ctx* ctx, const char* name, const void* data, const int64_t* shape, int ndim,
                       int dtype) {
  if (!ctx || !name || !data || !shape || ndim < 0 || ndim > 8) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  if (dtype != AMUSE_DTYPE_F32) return fail(ctx, AMUSE_E_UNSUPPORTED, "only fp32 weights are supported");
  cudaSetDevice(ctx->device);
  std::string key(name);
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  if (key.compare(0, 4, "ast.") == 0) {   // 1 GB of encoder weights stays on the device
    int rc = ast::stage(ctx->astw, key.substr(4), data, shape, ndim);
    if (rc) return fail(ctx, rc, "ast weight '%s' rejected", name);
    return AMUSE_OK;
  }
  if (key == "scheduler.alphas_cumprod") {
    ctx->alphas_cumprod.resize(static_cast<size_t>(n));
    CU(cudaMemcpy(ctx->alphas_cumprod.data(), data, static_cast<size_t>(n) * 4, cudaMemcpyDefault));
    drop_schedules(ctx);
    return AMUSE_OK;
  }
  if (key.compare(0, 9, "denoiser.") != 0 && key.compare(0, 4, "vae.") != 0)
    return fail(ctx, AMUSE_E_INVALID, "unknown weight namespace in '%s'", name);
  if (key == "denoiser.mem_pos.pe") return AMUSE_OK;   // never read by the reference's forward (denoiser.py:174-188)
  (key[0] == 'd' ? ctx->dirty_den : ctx->dirty_vae) = true;
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  t.data.resize(static_cast<size_t>(n));
  CU(cudaMemcpy(t.data.data(), data, static_cast<size_t>(n) * 4, cudaMemcpyDefault));
  ctx->raw[key] = std::move(t);
  return AMUSE_OK;
}

int amuse_finalize_weights(amuse_ctx* ctx, void* stream) {
  NvtxRange nvtx_("amuse.finalize_weights");
  if (!ctx) return AMUSE_E_INVALID;
  cudaSetDevice(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool any = false;
  if (find(ctx, "denoiser.encoder.norm.weight")) {   // (a bare "denoiser.time_proj.freqs" table does not count)
    if (ctx->dirty_den || !ctx->den.ready)
      if (int rc = pack_denoiser(ctx, st)) return rc;
    ctx->dirty_den = false;
    any = true;
  }
  if (find(ctx, "vae.decoder.norm.weight")) {
    if (ctx->dirty_vae || !ctx->dec.ready)
      if (int rc = pack_decoder(ctx, st)) return rc;
    any = true;
  }
  if (find(ctx, "vae.encoder.norm.weight")) {   // optional: only the edit path (MotionPrior.encode) needs it
    if (ctx->dirty_vae || !ctx->enc.ready)
      if (int rc = pack_encoder(ctx, st)) return rc;
    any = true;
  }
  ctx->dirty_vae = false;
  if (ast::staged(ctx->astw)) {
    int rc = ast::finalize(ctx->astw, st);
    if (rc) return fail(ctx, rc, "ast finalize: %s", ast::last_error(ctx->astw));
    any = true;
  }
  if (!any) return fail(ctx, AMUSE_E_MISSING, "no weights loaded");
  return AMUSE_OK;
}

int amuse_reserve(amuse_ctx* ctx, int max_clips, int max_steps) {
  if (!ctx || max_clips < 1) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  (void)max_steps;
  cudaSetDevice(ctx->device);
  CU(ctx->cond.ensure(static_cast<size_t>(max_clips) * 3 * 128));
  CU(ctx->lat_out.ensure(static_cast<size_t>(max_clips) * 128));
  if (ctx->dec.ready)
    if (int rc = reserve_decode(ctx, max_clips)) return rc;
  return AMUSE_OK;
}

int amuse_schedule(amuse_ctx* ctx, int n_steps, int sampler, float eta, int32_t* timesteps, float* coef) {
  if (!ctx) return AMUSE_E_INVALID;
  Schedule sc;
  if (int rc = build_schedule(ctx, n_steps, sampler, eta, sc)) return rc;
  if (timesteps) std::memcpy(timesteps, sc.timesteps.data(), sizeof(int32_t) * n_steps);
  if (coef) std::memcpy(coef, sc.coef.data(), sizeof(float) * 5 * n_steps);
  return AMUSE_OK;
}

int amuse_denoise(amuse_ctx* ctx, int B, int n_steps, int sampler, float eta, int clip_sample, const float* latents0,
                  const float* z_con, const float* z_emo, const float* z_sty, const float* step_noise,
                  uint64_t seed, uint64_t clip_offset, float* latents_out, void* stream) {
  NvtxRange nvtx_("amuse.denoise");
  if (int rc = check_ready(ctx, true, false)) return rc;
  if (B < 1 || !latents0 || !z_con || !latents_out) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  cudaSetDevice(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Schedule* sc = nullptr;
  if (int rc = get_schedule(ctx, n_steps, sampler, eta, st, &sc)) return rc;
  const int clip = (clip_sample < 0) ? (sampler == AMUSE_SAMPLER_DDIM ? 1 : 0) : (clip_sample != 0);
  return run_denoise(ctx, B, *sc, clip, latents0, z_con, z_emo, z_sty, step_noise, seed, clip_offset * 128ull,
                     latents_out, st);
}

int amuse_denoiser_eps(amuse_ctx* ctx, int B, int timestep, const float* sample, const float* z_con,
                       const float* z_emo, const float* z_sty, float* eps_out, void* stream) {
  if (int rc = check_ready(ctx, true, false)) return rc;
  if (B < 1 || !sample || !z_con || !eps_out) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  cudaSetDevice(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // a one-step "schedule" whose update returns eps itself: x0 = x, x' = 0*x0 + 1*eps
  Schedule sc;
  sc.n_steps = 1;
  sc.timesteps = {timestep};
  sc.coef = {1.f, 0.f, 0.f, 1.f, 0.f};
  sc.dir_uses_eps = true;
  CU(cudaMalloc(&sc.d_timesteps, sizeof(int)));
  CU(cudaMalloc(&sc.d_coef, 5 * sizeof(float)));
  CU(cudaMalloc(&sc.d_temb, 128 * sizeof(float)));
  CU(cudaMemcpyAsync(sc.d_timesteps, sc.timesteps.data(), sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(sc.d_coef, sc.coef.data(), 5 * sizeof(float), cudaMemcpyHostToDevice, st));
  const float* m = ctx->den.misc.p;
  CU(launch_time_table(sc.d_timesteps, 1, m + ctx->den.freqs, m + ctx->den.w1t, m + ctx->den.b1, m + ctx->den.w2t,
                       m + ctx->den.b2, sc.d_temb, st));
  ctx->launches++;
  int rc = run_denoise(ctx, B, sc, 0, sample, z_con, z_emo, z_sty, nullptr, 0, 0, eps_out, st);
  cudaStreamSynchronize(st);
  cudaFree(sc.d_timesteps);
  cudaFree(sc.d_coef);
  cudaFree(sc.d_temb);
  return rc;
}

int amuse_decode(amuse_ctx* ctx, int B, const float* latents, float* feats6d, float* poses, float* trans,
                 void* stream) {
  NvtxRange nvtx_("amuse.decode");
  if (int rc = check_ready(ctx, false, true)) return rc;
  if (B < 1 || !latents || (!feats6d && !poses)) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  cudaSetDevice(ctx->device);
  return run_decode_any(ctx, B, latents, feats6d, poses, trans, static_cast<cudaStream_t>(stream));
}

int amuse_encode(amuse_ctx* ctx, int B, const float* feats, float* mu, float* logvar, void* stream) {
  NvtxRange nvtx_("amuse.encode");
  if (!ctx) return AMUSE_E_INVALID;
  if (!ctx->enc.ready) return fail(ctx, AMUSE_E_STATE, "vae encoder weights not finalized");
  if (B < 1 || !feats || !mu || !logvar) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  cudaSetDevice(ctx->device);
  return run_encode_tc(ctx, B, feats, mu, logvar, static_cast<cudaStream_t>(stream));
}

int amuse_motion_to_feats(amuse_ctx* ctx, int64_t n_frames, const float* poses, const float* trans, float* feats,
                          void* stream) {
  if (!ctx || n_frames < 0 || !poses || !trans || !feats) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  if (n_frames == 0) return AMUSE_OK;
  cudaSetDevice(ctx->device);
  CU(launch_motion_to_feats(poses, trans, n_frames, feats, static_cast<cudaStream_t>(stream)));
  ctx->launches++;
  return AMUSE_OK;
}

int amuse_rot6d_to_axis_angle(amuse_ctx* ctx, int64_t n, const float* d6, float* axis_angle, void* stream) {
  if (!ctx || n < 0 || !d6 || !axis_angle) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  if (n == 0) return AMUSE_OK;
  cudaSetDevice(ctx->device);
  CU(launch_rot6d_flat(d6, n, axis_angle, static_cast<cudaStream_t>(stream)));
  ctx->launches++;
  return AMUSE_OK;
}

int amuse_diffusion_backward(amuse_ctx* ctx, int B, int n_steps, int sampler, float eta, int clip_sample,
                             const float* latents0, const float* z_con, const float* z_emo, const float* z_sty,
                             const float* step_noise, uint64_t seed, uint64_t clip_offset, float* latents_out,
                             float* feats6d, float* poses, float* trans, void* stream) {
  NvtxRange nvtx_("amuse.diffusion_backward");
  if (int rc = check_ready(ctx, true, true)) return rc;
  if (!poses) return fail(ctx, AMUSE_E_INVALID, "poses must not be NULL");
  cudaSetDevice(ctx->device);
  float* z = latents_out;
  if (!z) {
    CU(ctx->lat_out.ensure(static_cast<size_t>(B) * 128));
    z = ctx->lat_out.p;
  }
  if (int rc = amuse_denoise(ctx, B, n_steps, sampler, eta, clip_sample, latents0, z_con, z_emo, z_sty, step_noise,
                             seed, clip_offset, z, stream))
    return rc;
  return run_decode_any(ctx, B, z, feats6d, poses, trans, static_cast<cudaStream_t>(stream));
}

int amuse_diffusion_backward_host(amuse_ctx* ctx, int B, int n_steps, int sampler, float eta, int clip_sample,
                                  const float* latents0, const float* z_con, const float* z_emo,
                                  const float* z_sty, const float* step_noise, uint64_t seed,
                                  uint64_t clip_offset, float* poses, float* trans, void* stream) {
  NvtxRange nvtx_("amuse.diffusion_backward_host");
  if (int rc = check_ready(ctx, true, true)) return rc;
  if (B < 1 || !latents0 || !z_con || !poses) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  cudaSetDevice(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nl = static_cast<size_t>(B) * 128, nz = static_cast<size_t>(B) * 256;
  const size_t nn = step_noise ? static_cast<size_t>(n_steps) * B * 128 : 0;
  const size_t np = static_cast<size_t>(B) * kFrames * 165, nt = static_cast<size_t>(B) * kFrames * 3;
  CU(ctx->h2d.ensure(nl + 3 * nz + nn + np + nt));
  float* d_l0 = ctx->h2d.p;
  float* d_con = d_l0 + nl;
  float* d_emo = d_con + nz;
  float* d_sty = d_emo + nz;
  float* d_noise = d_sty + nz;
  float* d_poses = d_noise + nn;
  float* d_trans = d_poses + np;
  CU(cudaMemcpyAsync(d_l0, latents0, nl * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_con, z_con, nz * 4, cudaMemcpyHostToDevice, st));
  if (z_emo) CU(cudaMemcpyAsync(d_emo, z_emo, nz * 4, cudaMemcpyHostToDevice, st));
  if (z_sty) CU(cudaMemcpyAsync(d_sty, z_sty, nz * 4, cudaMemcpyHostToDevice, st));
  if (step_noise) CU(cudaMemcpyAsync(d_noise, step_noise, nn * 4, cudaMemcpyHostToDevice, st));
  if (int rc = amuse_diffusion_backward(ctx, B, n_steps, sampler, eta, clip_sample, d_l0, d_con,
                                        z_emo ? d_emo : nullptr, z_sty ? d_sty : nullptr,
                                        step_noise ? d_noise : nullptr, seed, clip_offset, nullptr, nullptr, d_poses,
                                        d_trans,
                                        stream))
    return rc;
  CU(cudaMemcpyAsync(poses, d_poses, np * 4, cudaMemcpyDeviceToHost, st));
  if (trans) CU(cudaMemcpyAsync(trans, d_trans, nt * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return AMUSE_OK;
}

int amuse_ast_features(amuse_ctx* ctx, int B, const float* fbank, float* con, float* emo, float* sty, void* stream) {
  NvtxRange nvtx_("amuse.ast_features");
  if (!ctx || B < 1 || !fbank || !con || !emo || !sty) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  if (!ast::ready(ctx->astw)) return fail(ctx, AMUSE_E_STATE, "ast weights not finalized");
  cudaSetDevice(ctx->device);
  int64_t launches = 0;
  int rc = ast::forward(ctx->astw, B, fbank, con, emo, sty, static_cast<cudaStream_t>(stream), &launches);
  ctx->launches += launches;
  if (rc) return fail(ctx, rc, "ast forward: %s", ast::last_error(ctx->astw));
  return AMUSE_OK;
}

int amuse_debug_tc_gemm(amuse_ctx* ctx, int epi, int M, int N, int K, const float* A, const float* A2, int k_split,
                        const float* W, const float* bias, const float* R, const float* ln, const float* cvec,
                        int rows_per_clip, float* C, void* stream) {
  if (!ctx || !A || !W || !bias || !C) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  cudaSetDevice(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool pair = (epi & 0x100) != 0;   // route to the CTA-pair kernel (tc_gemm2.cu); residual is [M][N] there
  epi &= 0xff;
  const int K1 = A2 ? k_split : K, K2 = K - K1;
  const int ldr = pair ? N : 128;
  const size_t nA = static_cast<size_t>(M) * K1, nA2 = static_cast<size_t>(M) * K2, nW = static_cast<size_t>(N) * K;
  const size_t nC = static_cast<size_t>(M) * N, nR = R ? static_cast<size_t>(M) * ldr : 0;
  DevBuf buf;
  CU(buf.ensure(2 * (nA + nA2 + nW + nC + nR)));
  float* p = buf.p;
  auto take = [&](size_t n) { float* r = p; p += n; return r; };
  float *a_hi = take(nA), *a_lo = take(nA), *a2_hi = take(nA2), *a2_lo = take(nA2), *w_hi = take(nW), *w_lo = take(nW);
  float *c_hi = take(nC), *c_lo = take(nC), *r_hi = take(nR), *r_lo = take(nR);
  CU(tc::split_planes(A, a_hi, a_lo, nA, st));
  if (A2) CU(tc::split_planes(A2, a2_hi, a2_lo, nA2, st));
  CU(tc::split_planes(W, w_hi, w_lo, nW, st));
  if (R) CU(tc::split_planes(R, r_hi, r_lo, nR, st));
  tc::GemmDesc d{};
  d.A_hi = a_hi; d.A_lo = a_lo; d.lda = K1;
  if (A2) { d.A2_hi = a2_hi; d.A2_lo = a2_lo; d.lda2 = K2; d.k_split = K1; }
  d.W_hi = w_hi; d.W_lo = w_lo; d.ldw = K;
  d.M = M; d.N = N; d.K = K; d.bias = bias;
  d.C = C; d.C_hi = c_hi; d.C_lo = c_lo; d.ldc = N;
  d.R_hi = r_hi; d.R_lo = r_lo; d.ldr = ldr;
  if (ln) { d.ln_g = ln; d.ln_b = ln + 128; d.ln2_g = ln + 256; d.ln2_b = ln + 384; }
  d.cvec = cvec; d.rows_per_clip = rows_per_clip > 0 ? rows_per_clip : 1;
  d.q_cols = 128; d.q_scale = 0.17677669529663687f;
  CU(pair ? tc::gemm2(epi, d, st) : tc::gemm(epi, d, st));
  ctx->launches++;
  if (epi != tc::EPI_PLAIN && epi != tc::EPI_QKV) {   // recombine the planes for the caller
    CU(launch_add_planes(c_hi, c_lo, C, nC, st));
  }
  CU(cudaStreamSynchronize(st));
  buf.release();
  return AMUSE_OK;
}

int amuse_fbank(amuse_ctx* ctx, int B, int n_samples, const float* wave, float norm_mean, float norm_std,
                float* fbank, void* stream) {
  NvtxRange nvtx_("amuse.fbank");
  if (!ctx || B < 1 || !wave || !fbank) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  if (n_samples < 400) return fail(ctx, AMUSE_E_INVALID, "need at least one 25 ms frame (400 samples)");
  cudaSetDevice(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!ctx->mel_t.p) {
    std::vector<float> m(257 * 128);
    fb::mel_banks_host(m.data());
    CU(ctx->mel_t.ensure(m.size()));
    CU(cudaMemcpy(ctx->mel_t.p, m.data(), m.size() * 4, cudaMemcpyHostToDevice));
    CU(fb::upload_tables());
  }
  CU(fb::launch(wave, B, n_samples, ctx->mel_t.p, norm_mean, norm_std, fbank, st));
  ctx->launches++;
  return AMUSE_OK;
}

int amuse_debug_philox_normals(amuse_ctx* ctx, uint64_t seed, uint64_t clip_offset, int B, int n_steps, float* out,
                               void* stream) {
  if (!ctx || B < 1 || n_steps < 1 || n_steps > 65535 || !out) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  cudaSetDevice(ctx->device);
  CU(launch_philox_export(seed, clip_offset, B, n_steps, out, static_cast<cudaStream_t>(stream)));
  ctx->launches++;
  return AMUSE_OK;
}

int amuse_debug_set_decode_plan(amuse_ctx* ctx, int clips_per_pass, int lanes) {
  if (!ctx || clips_per_pass < 1 || lanes < 1 || lanes > amuse_ctx::kMaxLanes) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  ctx->dec_chunk = clips_per_pass;
  ctx->dec_lanes = lanes;
  return AMUSE_OK;
}

int amuse_debug_attn_profile(amuse_ctx* ctx, int enable, int64_t* stamps, int n) {
  if (!ctx) return AMUSE_E_INVALID;
  cudaSetDevice(ctx->device);
  CU(cudaDeviceSynchronize());
  static_assert(sizeof(long long) == sizeof(int64_t), "stamp type");
  CU(attn::debug_profile(enable, reinterpret_cast<long long*>(stamps), n));
  return AMUSE_OK;
}

int64_t amuse_launch_count(amuse_ctx* ctx) { return ctx ? ctx->launches : 0; }

int amuse_profile_arm(amuse_ctx* ctx, int step) {
  if (!ctx) return AMUSE_E_INVALID;
  ctx->prof_step = step;
  return AMUSE_OK;
}
int amuse_profile_read(amuse_ctx* ctx, int64_t* stamps, int n) {
  if (!ctx || !stamps || n < 1 || n > 512) return fail(ctx, AMUSE_E_INVALID, "bad argument");
  cudaSetDevice(ctx->device);
  CU(cudaMemcpy(stamps, ctx->d_prof, sizeof(long long) * n, cudaMemcpyDeviceToHost));
  return AMUSE_OK;
}

}  // extern "C"
