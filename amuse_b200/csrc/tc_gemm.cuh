// tcgen05 (5th-gen tensor core) GEMM with fp32-class accuracy: C = epi(A . W^T + bias)
//   A [M][K] and W [N][K] are both K-major fp32 and are supplied as TWO planes each
//   (hi = round-to-nearest TF32 of the value, lo = value - hi, both stored as fp32), and the kernel
//   issues 3 MMAs per k-step (hi.hi + lo.hi + hi.lo, "3xTF32") with fp32 accumulation in TMEM.
// Used by the VAE decoder (M = clips*300) and the AST encoders (M = clips*1214).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace amuse {
namespace tc {

enum Epi {
  EPI_PLAIN = 0,        // C fp32 = acc + bias                       (ldc, bounds-checked)
  EPI_QKV = 1,          // plain, columns [0, q_cols) scaled by q_scale (nn.MultiheadAttention q scaling)
  EPI_PLANES = 2,       // C_hi/C_lo = split(acc + bias)
  EPI_GELU_PLANES = 3,  // C_hi/C_lo = split(gelu_erf(acc + bias))
  EPI_RES_LN_PLANES = 4,          // N == 128: split(LN(acc + bias + (R_hi + R_lo)))
  EPI_RES_LN_CROSS_LN_PLANES = 5, // N == 128: split(LN2(LN(acc + bias + R) + cvec[row / rows_per_clip]))
  EPI_RES_PLANES = 6,   // C_hi/C_lo = split(acc + bias + (R_hi + R_lo))   (pre-norm residual, AST)
  EPI_QKV_HEADS = 7,    // AST qkv (N = 3*heads*64): per-head operand planes for ast_attn.cu --
                        //   q (scaled by q_scale), k -> [clip][head][tokp][64];  v -> transposed [clip][head][64][tokp]
};

struct Planes {
  float* hi;
  float* lo;
};

struct GemmDesc {
  // operands (device pointers, row-major, K contiguous); A2 = optional second K half (skip concat)
  const float *A_hi, *A_lo;
  int lda;
  const float *A2_hi, *A2_lo;
  int lda2;
  int k_split;          // columns of K taken from A (rest from A2); == K when A2 is null
  const float *W_hi, *W_lo;
  int ldw;
  int M, N, K;
  const float* bias;    // [N]
  // outputs
  float* C;             // plain fp32 output (EPI_PLAIN / EPI_QKV)
  float *C_hi, *C_lo;   // plane outputs
  int ldc;
  // epilogue extras
  const float *R_hi, *R_lo;
  int ldr;
  const float *ln_g, *ln_b, *ln2_g, *ln2_b, *cvec;
  int rows_per_clip;
  int q_cols;
  float q_scale;
  float ln_eps;
  // EPI_QKV_HEADS: row m = clip * tok + token
  float *q_hi, *q_lo, *k_hi, *k_lo, *vt_hi, *vt_lo;
  int tok, tokp, heads;
};

// Builds the TMA tensor maps for `d` and launches.  Returns cudaSuccess or the failing status.
cudaError_t gemm(int epi, const GemmDesc& d, cudaStream_t st);

// The same contract on the CTA-pair kernel (tc_gemm2.cu: 256x256 tiles on two SMs, persistent, epilogue
// overlapped with the next tile): N % 256 == 0, single A source, EPI_PLAIN / EPI_GELU_PLANES /
// EPI_RES_PLANES / EPI_QKV_HEADS.  Used for the AST GEMMs.
cudaError_t gemm2(int epi, const GemmDesc& d, cudaStream_t st);

// TMA tensor map over a row-major fp32 [rows][cols] matrix (leading dimension ld floats): boxes of
// 32 columns (one 128-B swizzle row) x box_rows rows, SWIZZLE_128B -- the UMMA K-major operand layout.
cudaError_t make_map_2d(CUtensorMap* tm, const float* ptr, int rows, int cols, int ld, int box_rows);

// elementwise helper: split an fp32 tensor into hi/lo planes on the device
cudaError_t split_planes(const float* src, float* hi, float* lo, size_t n, cudaStream_t st);

// host-side split (weights at finalize): hi = RN-even to 10 mantissa bits, lo = v - hi
void split_host(const float* src, float* hi, float* lo, size_t n);

}  // namespace tc
}  // namespace amuse
