#pragma once
#include <cuda_runtime.h>
namespace amuse {
namespace fb {
void mel_banks_host(float* mel_t /*[257][128]*/);
cudaError_t upload_tables();
// wave [B][n_samples] (16 kHz, one channel) -> out [B][1024][128] normalised log-mel filterbank
cudaError_t launch(const float* wave, int B, int n_samples, const float* mel_t, float norm_mean, float norm_std,
                   float* out, cudaStream_t st);
}  // namespace fb
}  // namespace amuse
