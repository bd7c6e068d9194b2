// pwv_mel.cuh -- mel front end on the GPU: waveform -> normalised mel-dB spectrogram, the input of the generation
// path (reference data_load.py:37-56 -> audio.py:327-356 `wav2melspec_db`, which calls librosa: centred STFT
// n_fft / win / hop with a periodic hann window zero-padded to n_fft and reflect padding, magnitude, slaney mel
// basis, amplitude_to_db with amin 1e-5 and top_db 80 relative to the utterance's maximum, then
// clip((db - min_db) / (max_db - min_db), 0, 1) * 2 - 1 (audio.py:278-286)).
//
//   k_mel_power : CTA = (utterance, MEL_FB frames). The frames' samples (reflect-padded at both ends of the
//                 utterance) are staged once in shared memory; thread (frame, bin) evaluates the DFT bin directly
//                 against a cos/sin table in shared memory -- the window has win_length <= n_fft non-zero taps, the
//                 table index is (bin * n) mod n_fft, computed in double on the host; the transform is 0.2 MFLOP per
//                 frame, far too small for an FFT to matter next to the vocoder's 7.7 MFLOP per SAMPLE -- then the
//                 frame's |X| row is folded with the mel basis (n_mels x bins, only each band's non-zero span),
//                 converted to dB and written as [N][t_mel][n_mels]; the utterance's maximum dB is tracked with an
//                 ordered-integer atomicMax.
//   k_mel_finish: top_db floor relative to that maximum + normalisation, in place.
// Bandwidth-trivial (12 B per audio sample in, 4 * n_mels / hop out); the only goal is that real-audio generation
// needs neither librosa nor a host round trip.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pwv {

constexpr int MEL_FB = 8;          // frames per CTA
constexpr int MEL_THREADS = 256;

struct MelParams {
  const float* wav;        // [N][T]
  float* out;              // [N][t_mel][n_mels]
  const float* window;     // [n_fft]: hann(win_length) centred in n_fft, zeros outside (torch.stft / librosa padding)
  const float2* twiddle;   // [n_fft]: (cos, -sin)(2 pi i / n_fft)
  const float* basis;      // [n_mels][bins]
  const int* band_lo;      // [n_mels]: first bin with a non-zero weight
  const int* band_hi;      // [n_mels]: one past the last
  int* utt_max;            // [N]: ordered-int encoding of the utterance's maximum dB (initialised to INT_MIN)
  int N, T, t_mel, n_fft, hop, n_mels, bins, win_lo, win_hi;   // window taps win_lo .. win_hi-1 are non-zero
  float amin, top_db, min_db, max_db;
  int normalise;
};

__device__ __forceinline__ int mel_float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float mel_ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// dynamic shared memory: samples [MEL_FB * hop + n_fft] | twiddle [n_fft] float2 | mag [MEL_FB][bins + 1]
__global__ void __launch_bounds__(MEL_THREADS) k_mel_power(MelParams p) {
  extern __shared__ __align__(16) unsigned char mel_smem[];
  const int span = (MEL_FB - 1) * p.hop + p.n_fft;
  float* xs = reinterpret_cast<float*>(mel_smem);
  float2* tw = reinterpret_cast<float2*>(xs + ((span + 3) & ~3));
  float* mag = reinterpret_cast<float*>(tw + p.n_fft);
  const int ldm = p.bins + 1;
  const int n = blockIdx.y, f0 = blockIdx.x * MEL_FB;
  const float* w = p.wav + (size_t)n * p.T;
  // centred frames: frame f covers padded samples f*hop .. f*hop + n_fft - 1, padded = reflect(wav, n_fft/2)
  const int pad = p.n_fft / 2, start = f0 * p.hop - pad;
  for (int i = threadIdx.x; i < span; i += MEL_THREADS) {
    int s = start + i;
    if (s < 0) s = -s;                               // reflect (no edge repeat), as numpy 'reflect'
    if (s >= p.T) s = 2 * (p.T - 1) - s;
    xs[i] = (s >= 0 && s < p.T) ? w[s] : 0.f;        // (utterances shorter than the padding: zeros beyond one reflection)
  }
  for (int i = threadIdx.x; i < p.n_fft; i += MEL_THREADS) tw[i] = p.twiddle[i];
  __syncthreads();
  const int mask = p.n_fft - 1;                      // n_fft is a power of two (checked on the host)
  for (int e = threadIdx.x; e < MEL_FB * p.bins; e += MEL_THREADS) {
    const int fr = e / p.bins, k = e % p.bins;
    const float* x = xs + fr * p.hop;
    float re = 0.f, im = 0.f;
    int idx = (k * p.win_lo) & mask;
    for (int i = p.win_lo; i < p.win_hi; ++i) {
      const float v = x[i] * p.window[i];
      const float2 c = tw[idx];
      re = fmaf(v, c.x, re);
      im = fmaf(v, c.y, im);
      idx = (idx + k) & mask;
    }
    mag[fr * ldm + k] = sqrtf(re * re + im * im);
  }
  __syncthreads();
  float local_max = -3.0e38f;
  for (int e = threadIdx.x; e < MEL_FB * p.n_mels; e += MEL_THREADS) {
    const int fr = e / p.n_mels, m = e % p.n_mels;
    if (f0 + fr >= p.t_mel) continue;
    const float* b = p.basis + (size_t)m * p.bins;
    float acc = 0.f;
    for (int k = p.band_lo[m]; k < p.band_hi[m]; ++k) acc = fmaf(b[k], mag[fr * ldm + k], acc);
    const float db = 20.f * log10f(fmaxf(acc, p.amin));
    p.out[((size_t)n * p.t_mel + f0 + fr) * p.n_mels + m] = db;
    local_max = fmaxf(local_max, db);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, off));
  if ((threadIdx.x & 31) == 0 && local_max > -3.0e38f) atomicMax(p.utt_max + n, mel_float_to_ordered(local_max));
}

__global__ void __launch_bounds__(256) k_mel_finish(MelParams p) {
  const size_t per_utt = (size_t)p.t_mel * p.n_mels;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_utt * p.N) return;
  const int n = (int)(idx / per_utt);
  float db = fmaxf(p.out[idx], mel_ordered_to_float(p.utt_max[n]) - p.top_db);
  if (p.normalise) db = (fminf(fmaxf((db - p.min_db) / (p.max_db - p.min_db), 0.f), 1.f) - 0.5f) * 2.f;
  p.out[idx] = db;
}

__global__ void k_mel_init(int* utt_max, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) utt_max[i] = (int)0x80000000;
}

}  // namespace pwv
