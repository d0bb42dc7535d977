// Audio front-end on device: waveform -> log-mel "image" of the audio backbone.
//
// Reference: CAVP_TRAINER.preprocess_audio (trainer/trainer_cavp_vpo_mono.py:43-53,61-71) =
//   torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=512, win_length=400, hop_length=160, n_mels=64,
//   f_min=125, f_max=3800)  [defaults: power=2, center=True, pad_mode="reflect", periodic Hann window zero-padded to
//   n_fft, HTK mel scale, no filterbank norm]  -> first 96 / 300 frames -> transpose -> utils/sourcesep.py:23-47
//   db_from_amp: 20*log10(max(1e-5, x)) -> normalize_spec: 2*(x - spec_min)/(spec_max - spec_min) - 1.
// Here: (1) mel_frames_kernel writes the windowed, reflect-padded frames [rows*T][n_fft]; (2) the real DFT is ONE
// tensor-core GEMM against a constant [2*(n_fft/2+1)][n_fft] cos / -sin basis (cavp_igemm, fp32-parity mode: the path
// has no FFT butterfly to port, and 3.2 GFLOP per batch is ~20 us on the tile kernels); (3) mel_power_db_kernel turns
// re/im into power, applies the triangular filterbank from shared memory and the dB / range normalisation, and writes
// [rows][T][n_mels] directly (the reference's transpose(-1, -2) is a layout choice of the store).
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/cavp_b200.h"
#include "common.cuh"

namespace cavp {

// frames[(r*T + t)][j] = window[j] * padded_r[t*hop + j],  padded = reflect-pad(wave_r, n_fft/2)  (torch.stft center=True)
__global__ void mel_frames_kernel(const float* __restrict__ wave, long long ldw, int A, int rows, int T, int n_fft,
                                  int hop, const float* __restrict__ window, float* __restrict__ frames) {
  const int pad = n_fft / 2;
  const int q4 = n_fft / 4;
  const long long total = static_cast<long long>(rows) * T * q4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j0 = static_cast<int>(i % q4) * 4;
    const long long ft = i / q4;
    const int t = static_cast<int>(ft % T);
    const long long r = ft / T;
    const float* w = wave + r * ldw;
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int s = t * hop + j0 + u - pad;  // index into the unpadded waveform
      if (s < 0) s = -s;               // reflect (no edge repeat): x[-k] = x[k]
      if (s >= A) s = 2 * (A - 1) - s; //                            x[A-1+k] = x[A-1-k]
      v[u] = window[j0 + u] * w[s];
    }
    *reinterpret_cast<float4*>(frames + ft * n_fft + j0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// spec[(frame)][0..nf) = Re, [nf..2nf) = Im (row stride lds).  One block per frame: power -> mel -> dB -> range.
__global__ void mel_power_db_kernel(const float* __restrict__ spec, int lds, int nf, const float* __restrict__ fb,
                                    int n_mels, long long frames, float amin, float db_scale, float norm_a,
                                    float norm_b, float* __restrict__ out) {
  extern __shared__ float pw[];  // [nf]
  for (long long f = blockIdx.x; f < frames; f += gridDim.x) {
    const float* s = spec + f * lds;
    __syncthreads();
    for (int k = threadIdx.x; k < nf; k += blockDim.x) {
      const float re = s[k], im = s[nf + k];
      pw[k] = __fmaf_rn(re, re, __fmul_rn(im, im));  // |X|^2 (power = 2)
    }
    __syncthreads();
    for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < nf; ++k) acc = __fmaf_rn(pw[k], __ldg(fb + static_cast<size_t>(k) * n_mels + m), acc);
      const float db = db_scale * log10f(fmaxf(amin, acc));     // 20*log10(max(1e-5, x))
      out[f * n_mels + m] = __fmaf_rn(norm_a, db, norm_b);       // 2*(db - min)/(max - min) - 1
    }
  }
}

}  // namespace cavp

using namespace cavp;

extern "C" int cavp_mel_frames(const float* wave, long long ldw, int A, int rows, int T, int n_fft, int hop,
                               const float* window, float* frames, void* stream) {
  if (!wave || !window || !frames) return CAVP_ERR_NULL;
  if ((n_fft & 3) || (reinterpret_cast<uintptr_t>(frames) & 15)) return CAVP_ERR_ALIGN;
  // reflect padding needs n_fft/2 < A, and the last frame must end inside the padded signal
  if (rows <= 0 || T <= 0 || hop <= 0 || n_fft / 2 >= A || static_cast<long long>(T - 1) * hop + n_fft > A + n_fft)
    return CAVP_ERR_ARG;
  const long long work = static_cast<long long>(rows) * T * (n_fft / 4);
  mel_frames_kernel<<<grid_for(work, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(wave, ldw, A, rows, T, n_fft, hop,
                                                                                     window, frames);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int cavp_mel_power_db(const float* spec, int lds, int nf, const float* fb, int n_mels, long long frames,
                                 float amin, float db_scale, float spec_min, float spec_max, float* out, void* stream) {
  if (!spec || !fb || !out) return CAVP_ERR_NULL;
  if (nf <= 0 || n_mels <= 0 || frames <= 0 || lds < 2 * nf || !(spec_max > spec_min)) return CAVP_ERR_ARG;
  const float a = 2.0f / (spec_max - spec_min);
  const float b = -2.0f * spec_min / (spec_max - spec_min) - 1.0f;
  const int blocks = frames < NUM_SMS * 16 ? static_cast<int>(frames) : NUM_SMS * 16;
  mel_power_db_kernel<<<blocks, 64, nf * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      spec, lds, nf, fb, n_mels, frames, amin, db_scale, a, b, out);
  return static_cast<int>(cudaGetLastError());
}
