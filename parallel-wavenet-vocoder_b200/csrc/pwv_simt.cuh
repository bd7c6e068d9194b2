// pwv_simt.cuh -- fp32 CUDA-core kernels of the IAF-vocoder generation path (PWV_PREC_FP32).
//
// These are the exact-arithmetic kernels: every contraction is an fp32 FFMA chain, tanh/sigmoid
// use the accurate libdevice routines. They are the parity anchor for the tcgen05 kernels and the
// path for channel counts the tensor-core kernels do not cover.
//
// Data layout in HBM (all fp32, channels-last so that one time step is one contiguous row):
//   act   [2 bodies][N][T][C]   gated-layer activations ("cur", reference modules.py:134,251)
//   cbias [2 bodies][L][N][t_mel][2C]  per-frame conditioning term of a layer, filter|gate halves,
//         = relu(mel.Wc)[frame] . [gc_filter|gc_gate] + [filter_bias|gate_bias]
//         (reference modules.py:216-228 evaluated at mel rate: the reference repeats each frame
//          hop times and then projects; projecting first gives the same dot products)
//   x / scale / shift [N][T]
// Sample s of an utterance uses mel frame (s + hop/2) / hop (reference models.py:131-133).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pwv {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int bytes = valid ? 16 : 0;   // src-size 0 => 16 bytes of zeros are written
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------------
// Generic row GEMM used for the two conditioning projections (1.2 % of the path's FLOPs):
//   out[z][m][n] = act( sum_k A[m][k] * B_z[k][n] + bias_z[n] )
// 64x64 output tile per CTA, 16x16 threads, 4x4 outputs per thread, K staged 16 at a time.
// B / bias / out of batch entry z sit at fixed strides so one launch covers every (body, layer)
// of a flow.
// ------------------------------------------------------------------------------------------------
struct RowGemmBatch {
  const float* B;      size_t strideB;      // entry z: [K][Nc] at B + z*strideB
  const float* bias;   size_t strideBias;   // entry z: [Nc] (nullptr: no bias)
  float* out;          size_t strideOut;    // entry z: [M][Nc]
  const float* colscale;                    // [Nc] applied after the bias (nullptr: none)
};

template <bool RELU>
__global__ void __launch_bounds__(256) k_row_gemm(const float* __restrict__ A, RowGemmBatch batch,
                                                   int M, int K, int Nc) {
  __shared__ float As[16][64 + 4];   // [k][m]
  __shared__ float Bs[16][64 + 4];   // [k][n]
  const int z = blockIdx.z;
  const float* __restrict__ B = batch.B + z * batch.strideB;
  const float* __restrict__ bias = batch.bias ? batch.bias + z * batch.strideBias : nullptr;
  float* __restrict__ out = batch.out + z * batch.strideOut;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    // A tile: 64 rows x 16 k; thread loads 4 elements
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      int m = e / 16, k = e % 16;
      float v = 0.f;
      if (m0 + m < M && k0 + k < K) v = A[(size_t)(m0 + m) * K + k0 + k];
      As[k][m] = v;
    }
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      int k = e / 64, n = e % 64;
      float v = 0.f;
      if (k0 + k < K && n0 + n < Nc) v = B[(size_t)(k0 + k) * Nc + n0 + n];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= Nc) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (batch.colscale) v *= batch.colscale[n];
      if (RELU) v = fmaxf(v, 0.f);
      out[(size_t)m * Nc + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Flow front: IAF combine of the previous flow + the causal layer of both bodies.
//   x_new[t] = x[t] * scale[t] + shift[t]        (reference modules.py:57-59; flow 0: x_new = noise)
//   cur_b[t][r] = Wc_b[0][r] * x_new[t-1] + Wc_b[1][r] * x_new[t]   (reference modules.py:174-183,
//                                                                    k=2, dilation 1, no bias)
// One thread per (sample, 4 channels); both bodies are written by the same thread.
// ------------------------------------------------------------------------------------------------
struct FrontParams {
  const float* x_prev;    // [N][T]
  const float* scale;     // [N][T] or nullptr (first flow)
  const float* shift;     // [N][T]
  float* x_new;           // [N][T]
  const float* wc[2];     // per body: [2][C]  (tap t-1, tap t)
  float* act;             // [2][N][T][C]
  int N, T, C;
};

__global__ void __launch_bounds__(256) k_front(FrontParams p) {
  const int cg_per_row = p.C / 4;
  const size_t total = (size_t)p.N * p.T * cg_per_row;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = (int)(idx % cg_per_row);
  const size_t row = idx / cg_per_row;          // n*T + t
  const int t = (int)(row % p.T);
  float xc = p.x_prev[row];
  float xp = (t > 0) ? p.x_prev[row - 1] : 0.f;
  if (p.scale) {
    xc = xc * p.scale[row] + p.shift[row];
    if (t > 0) xp = xp * p.scale[row - 1] + p.shift[row - 1];
  }
  if (cg == 0) p.x_new[row] = xc;
  const size_t body_stride = (size_t)p.N * p.T * p.C;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    const float4 w0 = *reinterpret_cast<const float4*>(p.wc[b] + cg * 4);
    const float4 w1 = *reinterpret_cast<const float4*>(p.wc[b] + p.C + cg * 4);
    float4 o;
    o.x = w0.x * xp + w1.x * xc;
    o.y = w0.y * xp + w1.y * xc;
    o.z = w0.z * xp + w1.z * xc;
    o.w = w0.w * xp + w1.w * xc;
    *reinterpret_cast<float4*>(p.act + b * body_stride + row * p.C + cg * 4) = o;
  }
}

// wav = x * scale + shift for the last flow
__global__ void __launch_bounds__(256) k_iaf_combine(const float* __restrict__ x, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, float* __restrict__ out,
                                                      size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] * scale[i] + shift[i];
}

// ------------------------------------------------------------------------------------------------
// Tile GEMM building block shared by the gated-layer and post-net kernels.
//
// CTA tile: TM rows (time steps) x ncols output columns; thread (tx, ty) of an NTX x NTY grid owns
// rows {ty + NTY*i, i<8} and columns {4tx..4tx+3} (+ {ncols/2 + 4tx ..+3} when NCT == 8, so the
// filter and gate pre-activations of a channel land in the same thread). A is resident in shared
// memory (row-major, leading dimension lda, read as float4 along k); B[K][ncols] streams from
// global/L2 through a double-buffered cp.async ring of KC-row chunks.
// ------------------------------------------------------------------------------------------------
constexpr int KC = 16;

template <int NTHREADS>
__device__ __forceinline__ void load_b_chunk(float* Bs, const float* __restrict__ Bg, int ncols, int k0) {
  const int n4 = KC * ncols / 4;
  const float4* src = reinterpret_cast<const float4*>(Bg + (size_t)k0 * ncols);
  for (int e = threadIdx.x; e < n4; e += NTHREADS) cp_async16(Bs + e * 4, src + e, true);
}

template <int NTX, int NTY, int NCT>
__device__ __forceinline__ void tile_gemm(const float* As, int lda, int K, const float* __restrict__ Bg,
                                          int ncols, float* Bs /* 2*KC*ncols floats */, float (&acc)[8][NCT]) {
  constexpr int NT = NTX * NTY;
  const int tx = threadIdx.x % NTX, ty = threadIdx.x / NTX;
  const int nchunks = K / KC;
  load_b_chunk<NT>(Bs, Bg, ncols, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    float* cur = Bs + (c & 1) * KC * ncols;
    if (c + 1 < nchunks) {
      load_b_chunk<NT>(Bs + ((c + 1) & 1) * KC * ncols, Bg, ncols, (c + 1) * KC);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // chunk c (and, for c == 0, the caller's A tile) visible to everyone
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        a[i] = *reinterpret_cast<const float4*>(As + (size_t)(ty + NTY * i) * lda + c * KC + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float b[NCT];
        const float4 b0 = *reinterpret_cast<const float4*>(cur + (kk + q) * ncols + 4 * tx);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        if (NCT == 8) {
          const float4 b1 = *reinterpret_cast<const float4*>(cur + (kk + q) * ncols + ncols / 2 + 4 * tx);
          b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float av = q == 0 ? a[i].x : q == 1 ? a[i].y : q == 2 ? a[i].z : a[i].w;
#pragma unroll
          for (int j = 0; j < NCT; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
        }
      }
    }
    __syncthreads();   // everyone done with `cur` before it is refilled two iterations later
  }
}

template <int C>
struct TileCfg {
  static constexpr int TM = (C <= 128) ? 64 : 32;
  static constexpr int NTX = C / 4;
  static constexpr int NTY = TM / 8;
  static constexpr int NT = NTX * NTY;
  static constexpr int LDA = 2 * C + 4;
  // A tile (TM x 2C, padded) + double-buffered B chunk of the widest GEMM (ncols = 2C)
  static constexpr size_t SMEM = sizeof(float) * ((size_t)TM * LDA + 2 * KC * 2 * C);
};

// ------------------------------------------------------------------------------------------------
// Conditioning projections of one flow, all (body, layer) entries in one launch:
//   out[z][m][:] = (cproj[m][:] . Wgc_z + bias_z) * colscale        m = (utterance, mel frame)
// 64-row x 2C-column tile per CTA on the tile_gemm building block (A resident, B streamed).
// grid = (ceil(M/64), 1, Z); K (= condition_channels) must be a multiple of 16 -- other values use
// k_row_gemm.
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(TileCfg<C>::NT) k_cond_gemm(const float* __restrict__ A, RowGemmBatch batch, int M, int K) {
  using Cfg = TileCfg<C>;
  constexpr int TM = Cfg::TM, NTX = Cfg::NTX, NTY = Cfg::NTY, NT = Cfg::NT;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                       // [TM][K + 4]
  const int lda = K + 4;
  float* Bs = smem + TM * Cfg::LDA;       // same carve-up as the layer kernel (K <= 2C)
  const int z = blockIdx.z, m0 = blockIdx.x * TM;
  const int tx = threadIdx.x % NTX, ty = threadIdx.x / NTX;
  const float* __restrict__ B = batch.B + z * batch.strideB;
  const float* __restrict__ bias = batch.bias ? batch.bias + z * batch.strideBias : nullptr;
  float* __restrict__ out = batch.out + z * batch.strideOut;
  {
    const int V = K / 4;
    for (int e = threadIdx.x; e < TM * V; e += NT) {
      const int m = e / V, v = e % V;
      const bool ok = m0 + m < M;
      cp_async16(As + (size_t)m * lda + v * 4, A + (size_t)(ok ? m0 + m : 0) * K + v * 4, ok);
    }
    cp_async_commit();
  }
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  tile_gemm<NTX, NTY, 8>(As, lda, K, B, 2 * C, Bs, acc);
  float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0, s0 = make_float4(1.f, 1.f, 1.f, 1.f), s1 = s0;
  if (bias) {
    b0 = *reinterpret_cast<const float4*>(bias + 4 * tx);
    b1 = *reinterpret_cast<const float4*>(bias + C + 4 * tx);
  }
  if (batch.colscale) {
    s0 = *reinterpret_cast<const float4*>(batch.colscale + 4 * tx);
    s1 = *reinterpret_cast<const float4*>(batch.colscale + C + 4 * tx);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty + NTY * i;
    if (m >= M) continue;
    float4 u, v;
    u.x = (acc[i][0] + b0.x) * s0.x; u.y = (acc[i][1] + b0.y) * s0.y; u.z = (acc[i][2] + b0.z) * s0.z; u.w = (acc[i][3] + b0.w) * s0.w;
    v.x = (acc[i][4] + b1.x) * s1.x; v.y = (acc[i][5] + b1.y) * s1.y; v.z = (acc[i][6] + b1.z) * s1.z; v.w = (acc[i][7] + b1.w) * s1.w;
    *reinterpret_cast<float4*>(out + (size_t)m * 2 * C + 4 * tx) = u;
    *reinterpret_cast<float4*>(out + (size_t)m * 2 * C + C + 4 * tx) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// One gated dilated layer (reference modules.py:185-259, k=2):
//   [f|g] = [x[t-d] | x[t]] . Wfg + cbias[frame(t)]
//   z     = tanh(f) * sigmoid(g)
//   mode 0: out[t] = x[t] + z . Wd + bd          (dense_output; the layer's skip output is dead
//                                                 unless it is the last layer)
//   mode 1: out[t] = z                           (last layer: its dense_output is dead, the
//                                                 post-net kernel consumes z)
//   mode 2: mode 0 + z_out[t] = z                (use_skip_connection: every layer's skip output
//                                                 is needed, k_skip_simt consumes z)
//   mode 3: fg_out[t] = [f|g] and stop           (normalize_wavenet: the pre-activations are normalised over the
//                                                 whole time axis before the gate, pwv_norm.cuh)
//   mode 4: [f|g] = fg_in[t] (already normalised), then as mode 2
// grid = (ceil(T/TM), N, 2 bodies)
// ------------------------------------------------------------------------------------------------
struct LayerParams {
  const float* x_in;      // [2][N][T][C]
  float* x_out;           // [2][N][T][C]
  const float* wfg[2];    // [2C][2C]  rows: tap(t-d) channels then tap(t) channels; cols: filter | gate
  const float* wd[2];     // [C][C]
  const float* bd[2];     // [C]
  const float* cbias[2];  // [N][t_mel][2C]
  float* z_out;           // [2][N][T][C], modes 2 and 4
  float* fg_out;          // [2][N][T][2C], mode 3
  const float* fg_in;     // [2][N][T][2C], mode 4
  int N, T, t_mel, hop, dilation, mode;
};

template <int C>
__global__ void __launch_bounds__(TileCfg<C>::NT) k_layer_simt(LayerParams p) {
  using Cfg = TileCfg<C>;
  constexpr int TM = Cfg::TM, NTX = Cfg::NTX, NTY = Cfg::NTY, NT = Cfg::NT, LDA = Cfg::LDA;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                 // [TM][LDA]
  float* Bs = smem + TM * LDA;      // [2][KC][2C]
  const int body = blockIdx.z, n = blockIdx.y, t0 = blockIdx.x * TM;
  const int tx = threadIdx.x % NTX, ty = threadIdx.x / NTX;
  const size_t body_stride = (size_t)p.N * p.T * C;
  const float* __restrict__ xin = p.x_in + body * body_stride + (size_t)n * p.T * C;
  float* __restrict__ xout = p.x_out + body * body_stride + (size_t)n * p.T * C;

  // A tile: columns [0,C) = x[t-d], [C,2C) = x[t]; rows past T or before 0 are zero.
  if (p.mode != 4) {
    constexpr int V = 2 * C / 4;   // float4 per row
    for (int e = threadIdx.x; e < TM * V; e += NT) {
      const int m = e / V, v = e % V;
      const int t = t0 + m;
      const bool delayed = v < C / 4;
      const int ts = delayed ? t - p.dilation : t;
      const int ch = (delayed ? v : v - C / 4) * 4;
      const bool ok = (t < p.T) && (ts >= 0);
      cp_async16(As + (size_t)m * LDA + v * 4, xin + (size_t)(ok ? ts : 0) * C + ch, ok);
    }
    cp_async_commit();   // joins the first B chunk's wait inside tile_gemm (wait_group counts groups)
  }

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (p.mode != 4) tile_gemm<NTX, NTY, 8>(As, LDA, 2 * C, p.wfg[body], 2 * C, Bs, acc);

  // conditioning term + gate; z tile aliases the A tile (tile_gemm ended with a barrier)
  float* Zs = As;                   // [TM][C+4]
  constexpr int LDZ = C + 4;
  const float* __restrict__ cb = p.cbias[body] + (size_t)n * p.t_mel * 2 * C;
  const size_t fg_base = ((size_t)body * p.N + n) * p.T * 2 * C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = ty + NTY * i;
    const int t = min(t0 + m, p.T - 1);
    float4 f, g;
    if (p.mode == 4) {              // normalised pre-activations from the previous pass
      f = *reinterpret_cast<const float4*>(p.fg_in + fg_base + (size_t)t * 2 * C + 4 * tx);
      g = *reinterpret_cast<const float4*>(p.fg_in + fg_base + (size_t)t * 2 * C + C + 4 * tx);
    } else {
      const int frame = (t + p.hop / 2) / p.hop;
      const float4 cf = *reinterpret_cast<const float4*>(cb + (size_t)frame * 2 * C + 4 * tx);
      const float4 cg = *reinterpret_cast<const float4*>(cb + (size_t)frame * 2 * C + C + 4 * tx);
      f = make_float4(acc[i][0] + cf.x, acc[i][1] + cf.y, acc[i][2] + cf.z, acc[i][3] + cf.w);
      g = make_float4(acc[i][4] + cg.x, acc[i][5] + cg.y, acc[i][6] + cg.z, acc[i][7] + cg.w);
    }
    if (p.mode == 3) {
      if (t0 + m < p.T) {
        *reinterpret_cast<float4*>(p.fg_out + fg_base + (size_t)(t0 + m) * 2 * C + 4 * tx) = f;
        *reinterpret_cast<float4*>(p.fg_out + fg_base + (size_t)(t0 + m) * 2 * C + C + 4 * tx) = g;
      }
      continue;
    }
    float4 z;
    z.x = tanhf(f.x) * sigmoid_exact(g.x);
    z.y = tanhf(f.y) * sigmoid_exact(g.y);
    z.z = tanhf(f.z) * sigmoid_exact(g.z);
    z.w = tanhf(f.w) * sigmoid_exact(g.w);
    if (p.mode == 1) {
      if (t0 + m < p.T) *reinterpret_cast<float4*>(xout + (size_t)(t0 + m) * C + 4 * tx) = z;
    } else {
      *reinterpret_cast<float4*>(Zs + (size_t)m * LDZ + 4 * tx) = z;
      if ((p.mode == 2 || p.mode == 4) && t0 + m < p.T)
        *reinterpret_cast<float4*>(p.z_out + body * body_stride + ((size_t)n * p.T + t0 + m) * C + 4 * tx) = z;
    }
  }
  if (p.mode == 1 || p.mode == 3) return;

  float acc2[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = 0.f;
  // (tile_gemm's first barrier makes the z tile visible)
  tile_gemm<NTX, NTY, 4>(Zs, LDZ, C, p.wd[body], C, Bs, acc2);

  const float4 bd = *reinterpret_cast<const float4*>(p.bd[body] + 4 * tx);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = t0 + ty + NTY * i;
    if (t >= p.T) continue;
    const float4 xr = *reinterpret_cast<const float4*>(xin + (size_t)t * C + 4 * tx);
    float4 o;
    o.x = xr.x + (acc2[i][0] + bd.x);
    o.y = xr.y + (acc2[i][1] + bd.y);
    o.z = xr.z + (acc2[i][2] + bd.z);
    o.w = xr.w + (acc2[i][3] + bd.w);
    *reinterpret_cast<float4*>(xout + (size_t)t * C + 4 * tx) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// use_skip_connection (reference modules.py:147: total = sum(outputs)): after every gated layer
//   skip_sum[t] (+)= z[t] . Ws + bs          (the layer's skip_output, modules.py:243-250)
// accumulated in layer order, as Python's left-to-right sum does. grid = (ceil(T/TM), N, 2 bodies)
// ------------------------------------------------------------------------------------------------
struct SkipParams {
  const float* z;         // [2][N][T][C]
  const float* ws[2];     // [C][2C]
  const float* bs[2];     // [2C]
  float* skip_sum;        // [2][N][T][2C]
  int N, T, first;        // first: assign instead of accumulate
};

template <int C>
__global__ void __launch_bounds__(TileCfg<C>::NT) k_skip_simt(SkipParams p) {
  using Cfg = TileCfg<C>;
  constexpr int TM = Cfg::TM, NTX = Cfg::NTX, NTY = Cfg::NTY, NT = Cfg::NT, LDA = Cfg::LDA;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + TM * LDA;
  const int body = blockIdx.z, n = blockIdx.y, t0 = blockIdx.x * TM;
  const int tx = threadIdx.x % NTX, ty = threadIdx.x / NTX;
  const float* __restrict__ zin = p.z + ((size_t)body * p.N + n) * p.T * C;
  float* __restrict__ sk = p.skip_sum + ((size_t)body * p.N + n) * p.T * 2 * C;
  constexpr int LDZ = C + 4;
  {
    constexpr int V = C / 4;
    for (int e = threadIdx.x; e < TM * V; e += NT) {
      const int m = e / V, v = e % V;
      const bool ok = t0 + m < p.T;
      cp_async16(As + (size_t)m * LDZ + v * 4, zin + (size_t)(ok ? t0 + m : 0) * C + v * 4, ok);
    }
    cp_async_commit();
  }
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  tile_gemm<NTX, NTY, 8>(As, LDZ, C, p.ws[body], 2 * C, Bs, acc);
  const float4 ba = *reinterpret_cast<const float4*>(p.bs[body] + 4 * tx);
  const float4 bb = *reinterpret_cast<const float4*>(p.bs[body] + C + 4 * tx);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = t0 + ty + NTY * i;
    if (t >= p.T) continue;
    float4* pa = reinterpret_cast<float4*>(sk + (size_t)t * 2 * C + 4 * tx);
    float4* pb = reinterpret_cast<float4*>(sk + (size_t)t * 2 * C + C + 4 * tx);
    float4 u = make_float4(acc[i][0] + ba.x, acc[i][1] + ba.y, acc[i][2] + ba.z, acc[i][3] + ba.w);
    float4 v = make_float4(acc[i][4] + bb.x, acc[i][5] + bb.y, acc[i][6] + bb.z, acc[i][7] + bb.w);
    if (!p.first) {
      const float4 su = *pa, sv = *pb;
      u.x += su.x; u.y += su.y; u.z += su.z; u.w += su.w;
      v.x += sv.x; v.y += sv.y; v.z += sv.z; v.w += sv.w;
    }
    *pa = u;
    *pb = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Post-net of a body (reference modules.py:145-165 with use_skip_connection False):
//   skip = z . Ws + bs ; h = relu(skip) . W1 + b1 ; y = relu(h) . W2 + b2      (S = 2C)
// z is the last layer's gate output. The final 2C -> 1 contraction is a warp-shuffle reduction
// across the threads that share a row. grid = (ceil(T/TM), N, 2 bodies)
// ------------------------------------------------------------------------------------------------
struct PostParams {
  const float* z;         // [2][N][T][C]
  const float* ws[2];     // [C][2C]
  const float* bs[2];     // [2C]
  const float* w1[2];     // [2C][2C]
  const float* b1[2];     // [2C]
  const float* w2[2];     // [2C]
  const float* b2[2];     // [1]
  float* y;               // [2][N][T]   (body 0 = scale, body 1 = shift)
  const float* skip_sum;  // use_skip_connection: [2][N][T][2C] = sum over layers of (z . Ws + bs); z / ws / bs unused
  int N, T;
};

template <int C>
__global__ void __launch_bounds__(TileCfg<C>::NT) k_post_simt(PostParams p) {
  using Cfg = TileCfg<C>;
  constexpr int TM = Cfg::TM, NTX = Cfg::NTX, NTY = Cfg::NTY, NT = Cfg::NT, LDA = Cfg::LDA;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                 // z tile [TM][C+4], later h tile [TM][2C+4]
  float* Bs = smem + TM * LDA;
  __shared__ float ysum[TM];
  const int body = blockIdx.z, n = blockIdx.y, t0 = blockIdx.x * TM;
  const int tx = threadIdx.x % NTX, ty = threadIdx.x / NTX;
  constexpr int LDZ = C + 4;
  if (threadIdx.x < TM) ysum[threadIdx.x] = 0.f;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (p.skip_sum) {
    // h0 = relu(sum of the layers' skip outputs), straight into the A tile of the postprocess1 GEMM
    const float* __restrict__ sk = p.skip_sum + ((size_t)body * p.N + n) * p.T * 2 * C;
    constexpr int V = 2 * C / 4;
    for (int e = threadIdx.x; e < TM * V; e += NT) {
      const int m = e / V, v = e % V;
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t0 + m < p.T) u = *reinterpret_cast<const float4*>(sk + (size_t)(t0 + m) * 2 * C + v * 4);
      u.x = fmaxf(u.x, 0.f); u.y = fmaxf(u.y, 0.f); u.z = fmaxf(u.z, 0.f); u.w = fmaxf(u.w, 0.f);
      *reinterpret_cast<float4*>(As + (size_t)m * LDA + v * 4) = u;
    }
  } else {
    const float* __restrict__ zin = p.z + ((size_t)body * p.N + n) * p.T * C;
    {
      constexpr int V = C / 4;
      for (int e = threadIdx.x; e < TM * V; e += NT) {
        const int m = e / V, v = e % V;
        const bool ok = t0 + m < p.T;
        cp_async16(As + (size_t)m * LDZ + v * 4, zin + (size_t)(ok ? t0 + m : 0) * C + v * 4, ok);
      }
      cp_async_commit();
    }
    tile_gemm<NTX, NTY, 8>(As, LDZ, C, p.ws[body], 2 * C, Bs, acc);

    // h0 = relu(skip) -> shared (the z tile is dead: tile_gemm ended with a barrier)
    const float4 ba = *reinterpret_cast<const float4*>(p.bs[body] + 4 * tx);
    const float4 bb = *reinterpret_cast<const float4*>(p.bs[body] + C + 4 * tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = ty + NTY * i;
      float4 u, v;
      u.x = fmaxf(acc[i][0] + ba.x, 0.f); u.y = fmaxf(acc[i][1] + ba.y, 0.f);
      u.z = fmaxf(acc[i][2] + ba.z, 0.f); u.w = fmaxf(acc[i][3] + ba.w, 0.f);
      v.x = fmaxf(acc[i][4] + bb.x, 0.f); v.y = fmaxf(acc[i][5] + bb.y, 0.f);
      v.z = fmaxf(acc[i][6] + bb.z, 0.f); v.w = fmaxf(acc[i][7] + bb.w, 0.f);
      *reinterpret_cast<float4*>(As + (size_t)m * LDA + 4 * tx) = u;
      *reinterpret_cast<float4*>(As + (size_t)m * LDA + C + 4 * tx) = v;
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
  }
  tile_gemm<NTX, NTY, 8>(As, LDA, 2 * C, p.w1[body], 2 * C, Bs, acc);

  {
    const float4 ba = *reinterpret_cast<const float4*>(p.b1[body] + 4 * tx);
    const float4 bb = *reinterpret_cast<const float4*>(p.b1[body] + C + 4 * tx);
    const float4 wa = *reinterpret_cast<const float4*>(p.w2[body] + 4 * tx);
    const float4 wb = *reinterpret_cast<const float4*>(p.w2[body] + C + 4 * tx);
    constexpr int W = NTX < 32 ? NTX : 32;    // lanes of a warp that share a row
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
      s = fmaf(fmaxf(acc[i][0] + ba.x, 0.f), wa.x, s);
      s = fmaf(fmaxf(acc[i][1] + ba.y, 0.f), wa.y, s);
      s = fmaf(fmaxf(acc[i][2] + ba.z, 0.f), wa.z, s);
      s = fmaf(fmaxf(acc[i][3] + ba.w, 0.f), wa.w, s);
      s = fmaf(fmaxf(acc[i][4] + bb.x, 0.f), wb.x, s);
      s = fmaf(fmaxf(acc[i][5] + bb.y, 0.f), wb.y, s);
      s = fmaf(fmaxf(acc[i][6] + bb.z, 0.f), wb.z, s);
      s = fmaf(fmaxf(acc[i][7] + bb.w, 0.f), wb.w, s);
#pragma unroll
      for (int off = W / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (tx % W == 0) {
        if (NTX <= 32) ysum[ty + NTY * i] = s;       // one writer per row
        else atomicAdd(&ysum[ty + NTY * i], s);      // two warps share a row (C = 256)
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < TM && t0 + threadIdx.x < p.T)
    p.y[((size_t)body * p.N + n) * p.T + t0 + threadIdx.x] = ysum[threadIdx.x] + p.b2[body][0];
}

}  // namespace pwv
