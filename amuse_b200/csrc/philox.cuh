// Stateless Philox4x32-10 (Salmon et al., SC'11) + Box-Muller: the N(0,1) draw of latent element `elem` at sampler
// step `step` for a given 64-bit seed.  Counter = (step, 0, elem_lo, elem_hi), key = (seed_lo, seed_hi): the value
// depends on nothing else, so the noise a clip sees is independent of how clips are packed into CTAs, of the batch
// size and of the GPU count (the reference draws i.i.d. noise per batch row, infer_ldm.py:137-141).
// Both denoise-loop kernels and the debug export (amuse_debug_philox_normals) call this one function; every
// floating-point operation is an explicit single-rounding intrinsic so that all call sites produce the same bits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace amuse {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long elem, uint32_t step) {
  uint32_t r[4];
  philox4x32_10(step, 0u, static_cast<uint32_t>(elem), static_cast<uint32_t>(elem >> 32),
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  // u1 in (0, 1]: 24 random bits + half an ulp, never 0;  u2 in [0, 1)
  const float u1 = __fmaf_rn(static_cast<float>(r[0] >> 8), 5.9604644775390625e-8f, 2.98023223876953125e-8f);
  const float u2 = __fmul_rn(static_cast<float>(r[1] >> 8), 5.9604644775390625e-8f);
  const float rad = __fsqrt_rn(__fmul_rn(-2.0f, logf(u1)));
  return __fmul_rn(rad, cospif(__fmul_rn(2.0f, u2)));
}

}  // namespace amuse
