// K6: Kaldi-compatible log-mel filterbank on the device, fused with the
// pad / truncate to 1024 frames and the dataset normalisation of PretrainedLPDM_v1.process_single_seq
// (reference infer_ldm.py:182-190).  Follows torchaudio.compliance.kaldi.fbank as the reference calls
// it: sample_frequency 16000, frame_length 25 ms (400), frame_shift 10 ms (160), snip_edges,
// remove_dc_offset, preemphasis 0.97 (replicate-padded), hanning window, round_to_power_of_two (512),
// power spectrum, 128 mel bins over [20 Hz, Nyquist] (htk_compat, no energy), log(max(., eps)).
// One block per output frame: 400 samples -> shared memory -> 512-point radix-2 FFT -> 257 powers ->
// 128 triangular mel sums (weights K-major so the 128 threads read coalesced) -> log -> normalise.
#include "fbank_kernel.cuh"

#include <cmath>
#include <vector>

#include "common.cuh"

namespace amuse {
namespace fb {

namespace {
constexpr int kWin = 400, kShift = 160, kFft = 512, kBins = 257, kMel = 128, kTarget = 1024;
__constant__ float2 c_twiddle[256];   // exp(-2 pi i k / 512)
__constant__ float c_window[kWin];    // hann, periodic = False

__global__ void __launch_bounds__(256) fbank_kernel(const float* __restrict__ wave, int n_samples, int n_frames,
                                                    const float* __restrict__ mel_t /*[257][128]*/, float norm_mean,
                                                    float norm_std2, float* __restrict__ out) {
  __shared__ float re[kFft], im[kFft];
  __shared__ float red[8];
  const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* dst = out + (static_cast<size_t>(b) * kTarget + f) * kMel;
  if (f >= n_frames) {   // zero-padded rows are normalised too: (0 - mean) / (2 std)   (infer_ldm.py:185-190)
    if (tid < kMel) dst[tid] = (0.f - norm_mean) / norm_std2;
    return;
  }
  const float* x = wave + static_cast<size_t>(b) * n_samples + static_cast<size_t>(f) * kShift;
  // remove_dc_offset: subtract the frame mean
  const float v0 = (tid < kWin) ? x[tid] : 0.f;
  const float v1 = (tid + 256 < kWin) ? x[tid + 256] : 0.f;
  float s = warp_sum(v0 + v1);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float mean = tot / kWin;
  // pre-emphasis on the mean-removed frame with replicate padding, then the window; zero pad to 512;
  // store bit-reversed for the in-place decimation-in-time FFT
  auto sample = [&](int i) { return x[i] - mean; };
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = tid + q * 256;
    float v = 0.f;
    if (i < kWin) v = (sample(i) - 0.97f * sample(i > 0 ? i - 1 : 0)) * c_window[i];
    const int r = __brev(static_cast<unsigned>(i)) >> 23;   // 9-bit reversal
    re[r] = v;
    im[r] = 0.f;
  }
  __syncthreads();
#pragma unroll 1
  for (int len = 2; len <= kFft; len <<= 1) {
    const int half = len >> 1;
    const int j = tid & (half - 1);
    const int base = ((tid / half) * len) + j;
    const float2 w = c_twiddle[j * (kFft / len)];
    const float ar = re[base], ai = im[base], br = re[base + half], bi = im[base + half];
    const float tr = br * w.x - bi * w.y, ti = br * w.y + bi * w.x;
    re[base] = ar + tr;
    im[base] = ai + ti;
    re[base + half] = ar - tr;
    im[base + half] = ai - ti;
    __syncthreads();
  }
  // power spectrum, bins 0..256 (reuse re[] as the power array)
  const float p0 = re[tid] * re[tid] + im[tid] * im[tid];
  const float p256 = (tid == 0) ? (re[256] * re[256] + im[256] * im[256]) : 0.f;
  __syncthreads();
  re[tid] = p0;
  if (tid == 0) re[256] = p256;
  __syncthreads();
  if (tid < kMel) {
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < kBins; ++k) acc = fmaf(re[k], mel_t[k * kMel + tid], acc);
    const float e = logf(fmaxf(acc, 1.1920928955078125e-07f));   // torch.finfo(float32).eps
    dst[tid] = (e - norm_mean) / norm_std2;
  }
}

double mel_scale(double f) { return 1127.0 * std::log(1.0 + f / 700.0); }
PerDeviceOnce g_tables;   // the __constant__ tables are per device
}  // namespace

// mel_t [257][128]: torchaudio.compliance.kaldi.get_mel_banks(128, 512, 16000, 20, 0, 100, -500, 1.0),
// zero-padded to 257 bins and transposed.
void mel_banks_host(float* mel_t) {
  const double nyq = 8000.0, low = 20.0, high = nyq, bw = 16000.0 / kFft;
  const double ml = mel_scale(low), mh = mel_scale(high), delta = (mh - ml) / (kMel + 1);
  for (int k = 0; k < kBins; ++k)
    for (int m = 0; m < kMel; ++m) {
      float w = 0.f;
      if (k < 256) {
        const double left = ml + m * delta, center = ml + (m + 1) * delta, right = ml + (m + 2) * delta;
        const double mel = mel_scale(bw * k);
        const double up = (mel - left) / (center - left), down = (right - mel) / (right - center);
        w = static_cast<float>(std::fmax(0.0, std::fmin(up, down)));
      }
      mel_t[k * kMel + m] = w;
    }
}

cudaError_t upload_tables_once() {
  std::vector<float2> tw(256);
  std::vector<float> win(kWin);
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < 256; ++k) tw[k] = make_float2(static_cast<float>(std::cos(-2.0 * pi * k / kFft)),
                                                    static_cast<float>(std::sin(-2.0 * pi * k / kFft)));
  for (int i = 0; i < kWin; ++i) win[i] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * pi * i / (kWin - 1)));
  cudaError_t e = cudaMemcpyToSymbol(c_twiddle, tw.data(), sizeof(float2) * 256);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(c_window, win.data(), sizeof(float) * kWin);
  return e;
}
cudaError_t upload_tables() { return g_tables.run(upload_tables_once); }

cudaError_t launch(const float* wave, int B, int n_samples, const float* mel_t, float norm_mean, float norm_std,
                   float* out, cudaStream_t st) {
  if (n_samples < kWin) return cudaErrorInvalidValue;
  int n_frames = 1 + (n_samples - kWin) / kShift;
  if (n_frames > kTarget) n_frames = kTarget;
  fbank_kernel<<<dim3(kTarget, B), 256, 0, st>>>(wave, n_samples, n_frames, mel_t, norm_mean, norm_std * 2.0f, out);
  return cudaGetLastError();
}

}  // namespace fb
}  // namespace amuse
