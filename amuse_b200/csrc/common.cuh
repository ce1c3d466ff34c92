// Shared device helpers for the amuse_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "amuse_b200 kernels are written for sm_100a (B200) only"
#endif

#include <mutex>

namespace amuse {

// Host: one-time initialisation that is PER DEVICE (cudaFuncSetAttribute, __constant__ uploads): a process may hold
// contexts on several devices (amuse_create takes a device ordinal), possibly created from different threads.
struct PerDeviceOnce {
  std::mutex mu;
  bool done[64] = {};
  template <class F>
  cudaError_t run(F&& init) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev & 63]) return cudaSuccess;
    if ((e = init()) != cudaSuccess) return e;
    done[dev & 63] = true;
    return cudaSuccess;
  }
};

// ---- model constants (configs/diff_latent_v2.json:23-47, prior_emotional_fing.json:6-20)
constexpr int kD = 128;        // latent / model width
constexpr int kFF = 512;       // feed-forward width
constexpr int kHeads = 4;
constexpr int kHeadDim = 32;
constexpr int kLayers = 9;     // 4 input + middle + 4 output blocks
constexpr int kCond = 256;     // cond_dim
constexpr int kFrames = 300;   // train_pose_framelen
constexpr int kFeats = 333;    // 55 joints x 6D + 3 trans
constexpr int kJoints = 55;
constexpr float kLnEps = 1e-5f;

// ---- mbarrier / bulk-copy (TMA engine, 1-D) primitives -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy executed by the TMA engine; completion posted on `bar` (bytes).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// The dynamic shared-memory window rounded up to 1024 bytes (SWIZZLE_128B operands).  Pointer arithmetic on the
// __shared__ array itself: rounding through uintptr_t turns every later access into a GENERIC load / store (LD.E / ST.E
// with 64-bit addresses -- the tcgen05 sampler loop had 120 of them and not one LDS / STS).
__device__ __forceinline__ uint8_t* align_smem_1024(uint8_t* smem_raw) {
  return smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
}

// ---- thread-block-cluster primitives ------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
// address of `local_smem_ptr`'s twin in CTA `rank` of this cluster (shared::cluster window)
__device__ __forceinline__ uint32_t map_to_rank(const void* local_smem_ptr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
  return out;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// ---- math ---------------------------------------------------------------------------------
// Branch-free single-precision erf, < 1 ulp (max abs error 5.8e-8 against float64 erf over
// [-6, 6]; checked in tests/test_host_math.py against scipy).  Two minimax polynomials (the
// |x| > 0.927734375 one in the exponent of 1 - exp(.)); both are evaluated and selected, so a warp
// never diverges.  CUDA's erff has the same accuracy class but costs ~150 cycles per call here.
__device__ __forceinline__ float erf_fast(float a) {
  const float t = fabsf(a), s = a * a;
  float r = fmaf(-1.72853470e-5f, t, 3.83197126e-4f);
  const float u = fmaf(-3.88396438e-3f, t, 2.42546219e-2f);
  r = fmaf(r, s, u);
  r = fmaf(r, t, -1.06777877e-1f);
  r = fmaf(r, t, -6.34846687e-1f);
  r = fmaf(r, t, -1.28717512e-1f);
  r = fmaf(r, t, -t);
  const float big = copysignf(1.0f - __expf(r), a);   // exp(r) < 0.2 here: ex2.approx error << 1 ulp of the result
  float q = -5.96761703e-4f;
  q = fmaf(q, s, 4.99119423e-3f);
  q = fmaf(q, s, -2.67681349e-2f);
  q = fmaf(q, s, 1.12819925e-1f);
  q = fmaf(q, s, -3.76125336e-1f);
  q = fmaf(q, s, 1.28379166e-1f);
  const float small = fmaf(q, a, a);
  return (t > 0.927734375f) ? big : small;
}
__device__ __forceinline__ float gelu_erf(float x) {   // F.gelu default (exact erf form)
  return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// LayerNorm over one 128-wide row held as one float4 per lane of a full warp
// (two-pass mean / biased variance, eps 1e-5 -- nn.LayerNorm semantics).
__device__ __forceinline__ float4 warp_layernorm128(float4 v, const float4 g, const float4 b) {
  float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / 128.0f);
  float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
  float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / 128.0f);
  float rstd = 1.0f / sqrtf(var + kLnEps);
  return make_float4(dx * rstd * g.x + b.x, dy * rstd * g.y + b.y, dz * rstd * g.z + b.z, dw * rstd * g.w + b.w);
}

}  // namespace amuse
