// pwv_norm.cuh -- the reference's 'in' normaliser (instance normalisation over TIME, reference modules.py:274-284)
// for the generation path, fp32:
//     mean, variance = tf.nn.moments(x, [1], keep_dims=True)           per (utterance, channel), population variance
//     y = gamma * ((x - mean) / (variance + 1e-8) ** .5) + beta
// It sits at every call site the reference has (models.py:27-29,70,121-122; modules.py:149-151,158-160,181-182,
// 230-234,253-257). A statistic over the whole time axis between two stages of a layer cannot live inside one tile
// kernel, so with a normaliser switched on the layer runs un-fused: pre-activations -> (stats, apply) -> gate + dense
// -> (stats, apply) ..., each statistic a two-kernel pass over the tensor (non-default option; no hparams case of the
// reference enables it, so this path is written for exactness, not speed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pwv {

// x [UB][T][Cn] -> stats [UB][Cn] = (mean, sqrt(variance + 1e-8)); sums in double, so the fp32 result is the
// correctly rounded statistic whatever T is. grid = (ceil(Cn / 32), UB), 256 threads: lane = channel, warp = time slice.
template <bool PRE_RELU>
__global__ void __launch_bounds__(256) k_in_stats(const float* __restrict__ x, float2* __restrict__ stats, int T, int Cn) {
  __shared__ double s_sum[8][33], s_sq[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane, ub = blockIdx.y;
  const float* xb = x + (size_t)ub * T * Cn;
  double sum = 0.0, sq = 0.0;
  if (c < Cn)
    for (int t = warp; t < T; t += 8) {
      float v = xb[(size_t)t * Cn + c];
      if (PRE_RELU) v = fmaxf(v, 0.f);
      sum += (double)v;
      sq += (double)v * (double)v;
    }
  s_sum[warp][lane] = sum;
  s_sq[warp][lane] = sq;
  __syncthreads();
  if (warp == 0 && c < Cn) {
    for (int w = 1; w < 8; ++w) { sum += s_sum[w][lane]; sq += s_sq[w][lane]; }
    const double mean = sum / T;
    double var = sq / T - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[(size_t)ub * Cn + c] = make_float2((float)mean, (float)sqrt(var + 1e-8));
  }
}

// in place: x = gamma * ((x - mean) / denom) + beta; gamma / beta of group ub / ub_per_group (the two WaveNet bodies
// of a flow are normalised in one launch with their own variables)
template <bool PRE_RELU>
__global__ void __launch_bounds__(256) k_in_apply(float* __restrict__ x, const float2* __restrict__ stats, const float* gamma0,
                                                  const float* beta0, const float* gamma1, const float* beta1, int ub_per_group,
                                                  size_t per_ub, int Cn, size_t total) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ub = (int)(idx / per_ub), c = (int)(idx % Cn);
  const bool g1 = ub >= ub_per_group;
  const float2 st = stats[(size_t)ub * Cn + c];
  float v = x[idx];
  if (PRE_RELU) v = fmaxf(v, 0.f);
  const float normalized = (v - st.x) / st.y;
  x[idx] = (g1 ? gamma1 : gamma0)[c] * normalized + (g1 ? beta1 : beta0)[c];
}

// cond at sample rate from mel-rate rows: out[n][s][:] = in[n][(s + hop/2) / hop][:]  (reference models.py:131-133:
// every frame repeated hop times, cropped by hop/2 at both ends) -- only materialised when the conditioning is normalised
__global__ void __launch_bounds__(256) k_repeat_crop(const float* __restrict__ in, float* __restrict__ out, int N, int T, int t_mel, int hop,
                                                     int Cc) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)N * T * Cc;
  if (idx >= total) return;
  const int c = (int)(idx % Cc);
  const size_t row = idx / Cc;
  const int s = (int)(row % T), n = (int)(row / T);
  out[idx] = in[((size_t)n * t_mel + (s + hop / 2) / hop) * Cc + c];
}

// out[row][c] = v[c]: the reference normalises the 4-D transposed-conv tensor (n, 1, len, C) over its size-1 axis
// (modules.py:277 with the call at models.py:121-122), which collapses every stage to its beta
__global__ void __launch_bounds__(256) k_fill_rows(float* __restrict__ out, const float* __restrict__ v, size_t rows, int Cn) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < rows * Cn) out[idx] = v[idx % Cn];
}

__global__ void __launch_bounds__(256) k_add_inplace(float* __restrict__ a, const float* __restrict__ b, size_t n) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) a[idx] += b[idx];
}

}  // namespace pwv
