// pwv_gen.cuh -- the general-shape fp32 path: any residual / dilation / skip channel counts and any filter_width
// (reference modules.py:210-244 takes them as free parameters; hparams/default.yaml:22-26). The fused kernels of
// pwv_simt.cuh / pwv_tc*.cuh cover R = D in {64, 128, 256}, S = 2R, filter_width 2; everything else runs here as
// an un-fused chain per layer -- pre-activation GEMM over the gathered taps, gate, dense GEMM + residual, skip GEMM --
// on ONE tiled FFMA GEMM kernel with the layer's gathers and adds folded into its loads and its epilogue.
// Exact fp32 arithmetic (FFMA chains, tanhf / expf), both bodies of a flow per launch (blockIdx.z).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace pwv {

// out[b][row][:Nc] (+)= act_in( gather(A[b]) ) . B[b] + bias[b] + cond[b][frame(row)] + resid[b][row]
//   row = n * T + t. gather: K = taps * lda, element k = tap * lda + ci reads A[row - (taps-1-tap) * dilation][ci],
//   zero when that sample lies before the utterance's start (causal_conv, reference modules.py:11-43).
struct GenGemm {
  const float* A[2];
  const float* B[2];      // [K][Nc]
  const float* bias[2];   // [Nc] or nullptr
  const float* cond[2];   // [N][crows][Nc] or nullptr: conditioning rows, frame = (t + hop/2) / hop (models.py:131-133)
  const float* resid[2];  // [rows][Nc] or nullptr
  float* out[2];          // [rows][Nc]
  int lda, taps, dilation, T, crows, hop;
  int M, K, Nc;           // M = N * T rows per body
  int relu_in, relu_out, accumulate;
};

__global__ void __launch_bounds__(256) k_gen_gemm(GenGemm p) {
  __shared__ float As[16][64 + 4];   // [k][m]
  __shared__ float Bs[16][64 + 4];   // [k][n]
  const int b = blockIdx.z;
  const float* __restrict__ A = p.A[b];
  const float* __restrict__ B = p.B[b];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int m = e / 16, k = e % 16;
      const int row = m0 + m, kk = k0 + k;
      float v = 0.f;
      if (row < p.M && kk < p.K) {
        const int tap = kk / p.lda, ci = kk - tap * p.lda;
        const long long shift = (long long)(p.taps - 1 - tap) * p.dilation;
        const int t = row % p.T;
        if ((long long)t >= shift) {
          v = A[((size_t)row - (size_t)shift) * p.lda + ci];
          if (p.relu_in) v = fmaxf(v, 0.f);
        }
      }
      As[k][m] = v;
    }
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      const int k = e / 64, n = e % 64;
      float v = 0.f;
      if (k0 + k < p.K && n0 + n < p.Nc) v = B[(size_t)(k0 + k) * p.Nc + n0 + n];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float* __restrict__ bias = p.bias[b];
  const float* __restrict__ cond = p.cond[b];
  const float* __restrict__ resid = p.resid[b];
  float* __restrict__ out = p.out[b];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= p.M) continue;
    const float* crow = nullptr;
    if (cond) {
      const int n = row / p.T, t = row - n * p.T;
      const int frame = (t + p.hop / 2) / p.hop;
      crow = cond + ((size_t)n * p.crows + frame) * p.Nc;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= p.Nc) continue;
      float v = acc[i][j];
      if (bias) v += bias[col];
      if (crow) v += crow[col];
      if (resid) v = resid[(size_t)row * p.Nc + col] + v;      // dense_output = input + transformed (modules.py:251)
      if (p.accumulate) v = out[(size_t)row * p.Nc + col] + v;  // sum(outputs) in layer order (modules.py:147)
      if (p.relu_out) v = fmaxf(v, 0.f);
      out[(size_t)row * p.Nc + col] = v;
    }
  }
}

// z[row][c] = tanh(fg[row][c]) * sigmoid(fg[row][D + c])      (modules.py:236); rows = both bodies
__global__ void __launch_bounds__(256) k_gen_gate(const float* __restrict__ fg, float* __restrict__ z, int D, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t row = i / D;
  const int c = (int)(i - row * D);
  const float f = fg[row * 2 * D + c], g = fg[row * 2 * D + D + c];
  z[i] = tanhf(f) * (1.0f / (1.0f + expf(-g)));
}

// Flow front for any residual_channels / filter_width: IAF combine of the previous flow (modules.py:57-59) + the
// causal layer of both bodies (modules.py:174-183: causal_conv with dilation 1, no bias):
//   cur_b[t][r] = sum_k Wc_b[k][r] * x_new[t - (taps-1-k)]
struct GenFront {
  const float* x_prev;    // [N][T]
  const float* scale;     // [N][T] or nullptr (first flow, or x already combined)
  const float* shift;
  float* x_new;           // [N][T]
  const float* wc[2];     // per body [taps][R]
  float* act;             // [2][N][T][R]
  int N, T, R, taps;
};

__global__ void __launch_bounds__(256) k_gen_front(GenFront p) {
  const size_t total = (size_t)p.N * p.T * p.R;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int r = (int)(idx % p.R);
  const size_t row = idx / p.R;
  const int t = (int)(row % p.T);
  float o0 = 0.f, o1 = 0.f;
  for (int k = 0; k < p.taps; ++k) {
    const int shift = p.taps - 1 - k;
    if (t < shift) continue;
    float x = p.x_prev[row - shift];
    if (p.scale) x = x * p.scale[row - shift] + p.shift[row - shift];
    if (shift == 0 && r == 0) p.x_new[row] = x;
    o0 = fmaf(p.wc[0][(size_t)k * p.R + r], x, o0);
    o1 = fmaf(p.wc[1][(size_t)k * p.R + r], x, o1);
  }
  p.act[idx] = o0;
  p.act[total + idx] = o1;
}

}  // namespace pwv
