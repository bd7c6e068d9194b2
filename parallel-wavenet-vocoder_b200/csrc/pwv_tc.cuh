// pwv_tc.cuh -- tcgen05 (5th-gen tensor core) kernel of the gated dilated layer, sm_100a.
//
// One launch = one gated layer (reference modules.py:185-259) of BOTH WaveNet bodies of a flow.
// Persistent CTAs (one per SM); every CTA keeps its body's layer weights resident in shared memory
// and walks 128-row time tiles:
//
//   A1[128 x 128] = [ x[t-d] | x[t] ]                   (rows = time steps of the tile)
//   D1[128 x 128] = A1 . W1          W1 = [tap(t-d); tap(t)] x [filter | gate]      (tcgen05.mma)
//   z [128 x  64] = tanh(D1_f + c_f) * sigmoid(D1_g + c_g)                          (epilogue 1)
//   D2[128 x  64] = z . W2           W2 = dense 1x1                                  (tcgen05.mma)
//   out           = x[t] + D2 + b_dense                                              (epilogue 2)
//
// Arithmetic (PWV_PREC_F16X3): every fp32 operand is split v = hi + lo with hi = fp16(v),
// lo = fp16(v - hi) (22 significant bits) and each contraction is three kind::f16 MMAs
// (lo.hi + hi.lo + hi.hi) accumulated in fp32 in TMEM -- fp32-level accuracy at 1/3 of the fp16
// tensor rate, i.e. 1.5x the rate of a 3xTF32 scheme and half its operand footprint. Weights are
// pre-scaled by a power of two so that their lo parts stay in fp16's normal range; the inverse
// scale rides along in the epilogue FMAs. PWV_PREC_BF16 runs the single hi.hi pass on bf16.
//
// Operand placement: A operands (activations, z) live in TMEM (written by the epilogue threads
// with tcgen05.st, two 16-bit elements per column), B operands (weights) in shared memory in the
// K-major no-swizzle core-matrix layout, loaded once per CTA by 1-D bulk copies (TMA unit).
// Activations come in as TMA tensor boxes (128 rows x 32 channels, 128B swizzle) of a 3-D map
// [2N utterance-bodies][T][64]: rows before an utterance's start or past its end are zero-filled by
// the TMA unit, so the causal zero history and ragged last tiles need no masks on the way in. The
// output is written in place into the x[t] boxes and copied out by the same warps with full-line
// stores, after which the boxes are refilled at once. (Round-1 measurement that forced this: per-row 128/256-byte bulk copies
// cost ~12 cycles of TMA issue each -- 9k cycles per tile, profiles/r1_tc_trace_v3_bulk_rows.txt.)
//
// Warp roles (640 threads): 16 worker warps = 2 tile slots x 2 channel halves x 4 lane quarters.
// A tile slot owns 256 TMEM columns (D1 128 | A1hi 64 | A1lo 64; D2 and z alias D1 / A1) and a
// staging slot; the two slots hold alternate tiles, so one tile's epilogue overlaps the other
// tile's MMAs. Within a slot a thread owns one row and 32 of its 64 channels (the two warps that
// share a lane quarter split the columns), which doubles the warps available to hide the MUFU /
// TMEM / mbarrier latencies of the epilogues. Warp 16 allocates TMEM, loads the weights and issues
// the slot's MMAs (one elected thread each: warps 16, 17, blocking on their slot's mbarriers). Warps
// 18, 19 are the slots' TMA producers: each refills its slot's x[t-d] boxes as soon as the workers
// have converted them and the x[t] boxes (and conditioning rows) once the output has been copied out. The kernel is launched with programmatic stream serialization: its
// prologue (barriers, TMEM, weight image) overlaps the previous layer's tail, and only the
// producers' first activation load waits for the previous layer (griddepcontrol.wait).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "pwv_ptx.cuh"

namespace pwv {

// ------------------------------------------------------------------------------------------------
// weight image of one (flow, body, layer), laid out exactly as it sits in shared memory
// ------------------------------------------------------------------------------------------------
constexpr int TC_C = 64;                       // residual = dilation channels
constexpr int TC_TM = 128;                     // rows (time steps) per tile
constexpr int TC_W1_BYTES = 128 * 128 * 2;     // [k/8][n=128][8] 16-bit
constexpr int TC_W2_BYTES = 64 * 64 * 2;       // [k/8][n=64][8]
constexpr int TC_OFF_W1HI = 0;
constexpr int TC_OFF_W1LO = TC_OFF_W1HI + TC_W1_BYTES;
constexpr int TC_OFF_W2HI = TC_OFF_W1LO + TC_W1_BYTES;
constexpr int TC_OFF_W2LO = TC_OFF_W2HI + TC_W2_BYTES;
constexpr int TC_OFF_BD = TC_OFF_W2LO + TC_W2_BYTES;      // 64 floats
constexpr int TC_OFF_SCAL = TC_OFF_BD + 256;              // 4 floats: sf, sg, s2, unused
constexpr int TC_IMAGE_BYTES = TC_OFF_SCAL + 256;         // 82,432 (multiple of 256)

constexpr float TC_KF = -2.8853900817779268f;  // -2*log2(e): a = 2^(KF*f) = e^(-2f)
constexpr float TC_KG = -1.4426950408889634f;  // -log2(e):   b = 2^(KG*g) = e^(-g)

// Staging of one tile slot: four TMA boxes of 128 rows x 32 channels (128 B rows, 128B-swizzled):
// x[t-d] channels 0-31 | x[t-d] channels 32-63 | x[t] channels 0-31 | x[t] channels 32-63
constexpr int TC_BOX_BYTES = TC_TM * 128;                 // 16 KB
constexpr int TC_STAGE_BYTES = 4 * TC_BOX_BYTES;          // 64 KB
constexpr int TC_SMEM_STAGE0 = ((TC_IMAGE_BYTES + 1023) / 1024) * 1024;
constexpr int TC_CB_FRAMES = 4;                           // conditioning rows (mel frames) staged per tile slot
constexpr int TC_CB_BYTES = TC_CB_FRAMES * 128 * 4;
constexpr int TC_SMEM_CB0 = TC_SMEM_STAGE0 + 2 * TC_STAGE_BYTES;
constexpr int TC_SMEM_BARS = TC_SMEM_CB0 + 2 * TC_CB_BYTES;
constexpr int TC_SMEM_BYTES = TC_SMEM_BARS + 256;         // + barriers / tmem slot

struct TcLayerSrc {
  const float* wfg;   // host, [2C][2C] packed fp32 (tap rows, filter|gate cols)
  const float* wd;    // host, [C][C]
  const float* bd;    // host, [C]
};

struct TcModel {
  uint8_t* d_images = nullptr;   // [n_layers_total] x TC_IMAGE_BYTES, order = (flow, body, layer)
  uint8_t* d_post = nullptr;     // [n_iaf * 2] x TCP_IMAGE_BYTES, order = (flow, body)
  uint8_t* d_cond = nullptr;     // [n_layers_total] x cond_image_bytes, order = (flow, body, layer); may stay null
  int cond_image_bytes = 0;
  size_t bytes = 0;
  int precision = 0;
};

// post-net image of one (flow, body): skip 1x1 (64 -> 128), postprocess1 (128 -> 128), vectors
constexpr int TCP_WS_BYTES = 64 * 128 * 2;      // [k/8][n=128][8]
constexpr int TCP_W1_BYTES = 128 * 128 * 2;
constexpr int TCP_OFF_WSHI = 0;
constexpr int TCP_OFF_WSLO = TCP_OFF_WSHI + TCP_WS_BYTES;
constexpr int TCP_OFF_W1HI = TCP_OFF_WSLO + TCP_WS_BYTES;
constexpr int TCP_OFF_W1LO = TCP_OFF_W1HI + TCP_W1_BYTES;
constexpr int TCP_OFF_BS = TCP_OFF_W1LO + TCP_W1_BYTES;   // 128 floats: skip bias
constexpr int TCP_OFF_B1 = TCP_OFF_BS + 512;              // 128 floats: postprocess1 bias
constexpr int TCP_OFF_W2 = TCP_OFF_B1 + 512;              // 128 floats: postprocess2 weights
constexpr int TCP_OFF_SCAL = TCP_OFF_W2 + 512;            // 1/s_skip, 1/s_1, b2
constexpr int TCP_IMAGE_BYTES = TCP_OFF_SCAL + 256;       // 100,096
constexpr int TCP_STAGE_BYTES = 2 * TC_BOX_BYTES;         // z rows of one tile slot (two channel-half boxes)
constexpr int TCP_SMEM_STAGE0 = ((TCP_IMAGE_BYTES + 1023) / 1024) * 1024;
constexpr int TCP_SMEM_BYTES = TCP_SMEM_STAGE0 + 2 * TCP_STAGE_BYTES + 256;

struct TcPostSrc {
  const float* ws;   // host [64][128]
  const float* bs;   // [128]
  const float* w1;   // [128][128]
  const float* b1;   // [128]
  const float* w2;   // [128]
  const float* b2;   // [1]
};

inline uint16_t tc_to16(float v, bool bf16) {
  if (bf16) return __nv_bfloat16_raw(__float2bfloat16_rn(v)).x;
  return __half_raw(__float2half_rn(v)).x;
}
inline float tc_from16(uint16_t u, bool bf16) {
  if (bf16) { __nv_bfloat16_raw r; r.x = u; return __bfloat162float(__nv_bfloat16(r)); }
  __half_raw r; r.x = u; return __half2float(__half(r));
}

// power-of-two scale that brings max|w| into [2^12, 2^13) (fp16 max is 65504)
inline float tc_pow2_scale(const float* w, size_t n) {
  float mx = 0.f;
  for (size_t i = 0; i < n; ++i) mx = std::fmax(mx, std::fabs(w[i]));
  if (!(mx > 0.f) || !std::isfinite(mx)) return 1.f;
  int e;
  std::frexp(mx, &e);            // mx = m * 2^e, m in [0.5, 1)
  return std::ldexp(1.f, 13 - e);
}

// K-major no-swizzle image of B[n][k] = w[k][n] (w row-major [K][Ncols]): chunk (k/8) is a block of
// Ncols rows x 16 bytes.
inline void tc_pack_b(uint8_t* hi, uint8_t* lo, const float* w, int K, int Ncols, float scale, bool bf16, bool split) {
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < Ncols; ++n) {
      const float v = w[(size_t)k * Ncols + n] * scale;
      const uint16_t h = tc_to16(v, bf16);
      const uint16_t l = split ? tc_to16(v - tc_from16(h, bf16), bf16) : 0;
      const size_t off = (size_t)(k / 8) * (Ncols * 16) + (size_t)n * 16 + (k % 8) * 2;
      std::memcpy(hi + off, &h, 2);
      std::memcpy(lo + off, &l, 2);
    }
}

inline void tc_model_free(TcModel& t) {
  if (t.d_images) cudaFree(t.d_images);
  if (t.d_post) cudaFree(t.d_post);
  if (t.d_cond) cudaFree(t.d_cond);
  t.d_images = nullptr;
  t.d_post = nullptr;
  t.d_cond = nullptr;
  t.cond_image_bytes = 0;
  t.bytes = 0;
}

// Conditioning projection image of one (flow, body, layer): B[n = 128][k = Cc] = [gc_filter | gc_gate]
// as fp16 hi / lo K-major chunks, then colmul[128] = colscale / 2^s and coladd[128] = bias * colscale
// (out = acc * colmul + coladd = (cproj . Wgc + bias) * colscale).
constexpr int TCC_MAX_CC = 96;            // two weight images + the cproj tile + the output tile must fit 227 KB
inline int tcc_image_bytes(int Cc) { return ((2 * (Cc / 8) * 2048 + 1024) + 1023) / 1024 * 1024; }
inline int tcc_a_pitch(int Cc) { return Cc * 4 + 16; }          // staged cproj row + 16 B (bank spread)
inline int tcc_smem_bytes(int Cc) { return 2 * tcc_image_bytes(Cc) + ((128 * tcc_a_pitch(Cc) + 1023) / 1024 * 1024) + 128 * 512 + 256; }
struct TcCondSrc {
  const float* wgc;       // host [Cc][128]
  const float* bias;      // host [128]
};
inline const char* tc_cond_build(TcModel& t, int precision, int Cc, const float* colscale, const std::vector<TcCondSrc>& src) {
  if (Cc % 16 != 0 || Cc > TCC_MAX_CC) return nullptr;   // the FFMA conditioning kernel is used instead
  const bool bf16 = precision == 2, split = precision == 1;
  const int img_bytes = tcc_image_bytes(Cc), half_bytes = (Cc / 8) * 2048;
  std::vector<uint8_t> host(src.size() * (size_t)img_bytes, 0);
  for (size_t i = 0; i < src.size(); ++i) {
    uint8_t* img = host.data() + i * (size_t)img_bytes;
    const float sc = bf16 ? 1.f : tc_pow2_scale(src[i].wgc, (size_t)Cc * 128);
    tc_pack_b(img, img + half_bytes, src[i].wgc, Cc, 128, sc, bf16, split);
    float* vec = reinterpret_cast<float*>(img + 2 * half_bytes);
    for (int n = 0; n < 128; ++n) {
      vec[n] = colscale[n] / sc;
      vec[128 + n] = src[i].bias[n] * colscale[n];
    }
  }
  if (t.d_cond) cudaFree(t.d_cond);
  t.d_cond = nullptr;
  if (cudaMalloc(&t.d_cond, host.size()) != cudaSuccess) return "cudaMalloc of the conditioning weight images failed";
  if (cudaMemcpy(t.d_cond, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) return "upload of the conditioning weight images failed";
  t.cond_image_bytes = img_bytes;
  t.bytes += host.size();
  return nullptr;
}

// precision: 1 = f16x3, 2 = bf16. Returns nullptr on success or a static error string.
inline const char* tc_model_build(TcModel& t, int precision, int C, const std::vector<TcLayerSrc>& layers,
                                  const std::vector<TcPostSrc>& posts) {
  if (C != TC_C) return "tensor-core kernels need residual_channels = 64";
  const bool bf16 = precision == 2, split = precision == 1;
  std::vector<uint8_t> host(layers.size() * (size_t)TC_IMAGE_BYTES, 0);
  for (size_t i = 0; i < layers.size(); ++i) {
    uint8_t* img = host.data() + i * (size_t)TC_IMAGE_BYTES;
    const float s1 = bf16 ? 1.f : tc_pow2_scale(layers[i].wfg, (size_t)128 * 128);
    const float s2 = bf16 ? 1.f : tc_pow2_scale(layers[i].wd, (size_t)64 * 64);
    tc_pack_b(img + TC_OFF_W1HI, img + TC_OFF_W1LO, layers[i].wfg, 128, 128, s1, bf16, split);
    tc_pack_b(img + TC_OFF_W2HI, img + TC_OFF_W2LO, layers[i].wd, 64, 64, s2, bf16, split);
    std::memcpy(img + TC_OFF_BD, layers[i].bd, 64 * sizeof(float));
    const float scal[4] = {TC_KF / s1, TC_KG / s1, 1.f / s2, 0.f};
    std::memcpy(img + TC_OFF_SCAL, scal, sizeof(scal));
  }
  tc_model_free(t);
  if (cudaMalloc(&t.d_images, host.size()) != cudaSuccess) return "cudaMalloc of the tensor-core weight images failed";
  if (cudaMemcpy(t.d_images, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) return "upload of the tensor-core weight images failed";
  std::vector<uint8_t> hpost(posts.size() * (size_t)TCP_IMAGE_BYTES, 0);
  for (size_t i = 0; i < posts.size(); ++i) {
    uint8_t* img = hpost.data() + i * (size_t)TCP_IMAGE_BYTES;
    const float ss = bf16 ? 1.f : tc_pow2_scale(posts[i].ws, (size_t)64 * 128);
    const float s1 = bf16 ? 1.f : tc_pow2_scale(posts[i].w1, (size_t)128 * 128);
    tc_pack_b(img + TCP_OFF_WSHI, img + TCP_OFF_WSLO, posts[i].ws, 64, 128, ss, bf16, split);
    tc_pack_b(img + TCP_OFF_W1HI, img + TCP_OFF_W1LO, posts[i].w1, 128, 128, s1, bf16, split);
    std::memcpy(img + TCP_OFF_BS, posts[i].bs, 128 * sizeof(float));
    std::memcpy(img + TCP_OFF_B1, posts[i].b1, 128 * sizeof(float));
    std::memcpy(img + TCP_OFF_W2, posts[i].w2, 128 * sizeof(float));
    const float scal[4] = {1.f / ss, 1.f / s1, posts[i].b2[0], 0.f};
    std::memcpy(img + TCP_OFF_SCAL, scal, sizeof(scal));
  }
  if (cudaMalloc(&t.d_post, hpost.size()) != cudaSuccess) return "cudaMalloc of the post-net weight images failed";
  if (cudaMemcpy(t.d_post, hpost.data(), hpost.size(), cudaMemcpyHostToDevice) != cudaSuccess) return "upload of the post-net weight images failed";
  t.bytes = host.size() + hpost.size();
  t.precision = precision;
  return nullptr;
}

// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
struct TcLayerParams {
  float* x_out;             // [2][N][T][64]: the layer's output (mode 1: z)
  const uint8_t* image[2];  // per body
  const float* cbias[2];    // per body [N][t_mel][128], PRE-SCALED: filter half by KF, gate half by KG
  int N, T, t_mel, hop, dilation, mode;
  int tiles_per_utt;        // ceil(T / 128)
  int cb_in_smem;           // 1: a tile spans <= TC_CB_FRAMES mel frames, its conditioning rows are staged by TMA
  // Tile handshake between consecutive gated layers (both kernels resident under programmatic dependent launch):
  // flags_in  = the previous layer's per-tile "output stored" flags [2 bodies][N * tiles_per_utt], or nullptr:
  //             wait for the whole previous kernel (griddepcontrol.wait) -- first layer of a flow;
  // flags_out = this layer's flags (zeroed at the start of the forward), or nullptr.
  const int* flags_in;
  int* flags_out;
  int prev_dilation;        // dilation of the previous layer: its x[t-d] reads of the rows this layer overwrites
  long long* trace;         // debug: [4 roles][16 tiles][16 events] clock64 stamps of CTA 0 (or nullptr)
};

// debug timeline of CTA 0: role 0/1 = first thread of tile slot 0/1, role 2 = MMA thread, 3 = producer
#define TC_TRACE(role, j, k)                                                         \
  do {                                                                               \
    if (p.trace && blockIdx.x == 0 && (j) < 16) p.trace[((role) * 16 + (j)) * 16 + (k)] = clock64(); \
  } while (0)

template <bool BF16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {   // a -> low half (even k), b -> high half
  if (BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}
template <bool BF16>
__device__ __forceinline__ float2 unpack16(uint32_t u) {
  if (BF16) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  } else {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
}

// hi/lo split of 8 consecutive elements -> 4 packed hi columns + 4 packed lo columns
template <bool BF16, bool SPLIT>
__device__ __forceinline__ void split8(const float (&v)[8], uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t h = pack16<BF16>(v[2 * j], v[2 * j + 1]);
    hi[j] = h;
    if (SPLIT) {
      const float2 hf = unpack16<BF16>(h);
      lo[j] = pack16<BF16>(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
    }
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2: one issue slot for two IEEE fp32
// operations, bit-identical to the scalar instructions). The epilogues are issue-bound, so halving the
// instruction count of their FMA-pipe work is worth the register-pair plumbing.
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// hi/lo split of 8 consecutive elements with packed arithmetic for the residual (v - hi)
template <bool BF16, bool SPLIT>
__device__ __forceinline__ void split8p(const float (&v)[8], uint32_t* hi, uint32_t* lo) {
  const uint64_t M1 = pk2(-1.f, -1.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t h = pack16<BF16>(v[2 * j], v[2 * j + 1]);
    hi[j] = h;
    if (SPLIT) {
      const float2 hf = unpack16<BF16>(h);
      float l0, l1;
      upk2(fma2(pk2(hf.x, hf.y), M1, pk2(v[2 * j], v[2 * j + 1])), l0, l1);      // v - hi, one rounding
      lo[j] = pack16<BF16>(l0, l1);
    }
  }
}
template <bool BF16, bool SPLIT, bool PK>
__device__ __forceinline__ void split8x(const float (&v)[8], uint32_t* hi, uint32_t* lo) {
  if (PK) split8p<BF16, SPLIT>(v, hi, lo);
  else split8<BF16, SPLIT>(v, hi, lo);
}

// Gate of 16 channels: z = tanh(f) * sigmoid(g) from the raw accumulators fr / gr, the epilogue scales
// sf / sg and the pre-scaled conditioning rows cbf[0..3] / cbg[0..3] (float4 each). See the comments in
// k_layer_tc's epilogue 1 for the arithmetic; PK = packed fp32x2 version of the same operations.
template <bool BF16, bool PK, int NCH, bool TANH = BF16>
__device__ __forceinline__ void tc_gate(const uint32_t (&fr)[NCH], const uint32_t (&gr)[NCH], const float4* cbf, const float4* cbg,
                                        float sf, float sg, float (&z)[NCH]) {
#pragma unroll
  for (int q = 0; q < NCH / 4; ++q) {  // 4 channels per step; conditioning rows straight from shared memory
    const float4 ca = cbf[q], cb4 = cbg[q];
    const float cf[4] = {ca.x, ca.y, ca.z, ca.w}, cg[4] = {cb4.x, cb4.y, cb4.z, cb4.w};
    if (!PK) {
      if (TANH) {
        // bf16 mode has no 1e-4 bar (operands carry 2^-9 relative error): one MUFU per transcendental,
        // tanh(f) * (0.5 + 0.5 tanh(g/2)). cbias/scales hold fe = -2 log2e f, ge = -log2e g.
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float fe = fmaf(__uint_as_float(fr[4 * q + e]), sf, cf[e]);
          const float ge = fmaf(__uint_as_float(gr[4 * q + e]), sg, cg[e]);
          const float th = tanh_approx(fe * (-0.34657359f));         // f   = fe / (-2 log2e)
          const float tg = tanh_approx(ge * (-0.34657359f));         // g/2 = ge / (-2 log2e)
          z[4 * q + e] = th * fmaf(tg, 0.5f, 0.5f);
        }
      } else {
        // a = e^(-2f) = 2^fe, b = e^(-g) = 2^ge; z = (1 - a) / ((1 + a)(1 + b)). Two channels share one
        // reciprocal: 1/(d0 d1) * d1 = 1/d0. Clamps keep d0 d1 finite (< 2e33): tanh(10) = 1 - 4e-9,
        // sigmoid(-18) = 1.5e-8, far below the 1e-4 bar; a, b -> 0 on the other side is exact.
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          const float fe0 = fminf(fmaf(__uint_as_float(fr[4 * q + e]), sf, cf[e]), 28.853901f);
          const float ge0 = fminf(fmaf(__uint_as_float(gr[4 * q + e]), sg, cg[e]), 25.968511f);
          const float fe1 = fminf(fmaf(__uint_as_float(fr[4 * q + e + 1]), sf, cf[e + 1]), 28.853901f);
          const float ge1 = fminf(fmaf(__uint_as_float(gr[4 * q + e + 1]), sg, cg[e + 1]), 25.968511f);
          const float a0 = ex2_approx(fe0), b0 = ex2_approx(ge0), a1 = ex2_approx(fe1), b1 = ex2_approx(ge1);
          const float d0 = (1.f + a0) * (1.f + b0), d1 = (1.f + a1) * (1.f + b1);
          const float rinv = rcp_approx(d0 * d1);
          z[4 * q + e] = (1.f - a0) * d1 * rinv;
          z[4 * q + e + 1] = (1.f - a1) * d0 * rinv;
        }
      }
    } else {
      const uint64_t SF = pk2(sf, sf), SG = pk2(sg, sg), ONE = pk2(1.f, 1.f), M1 = pk2(-1.f, -1.f);
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        const uint64_t FE = fma2(pk2(__uint_as_float(fr[4 * q + e]), __uint_as_float(fr[4 * q + e + 1])), SF, pk2(cf[e], cf[e + 1]));
        const uint64_t GE = fma2(pk2(__uint_as_float(gr[4 * q + e]), __uint_as_float(gr[4 * q + e + 1])), SG, pk2(cg[e], cg[e + 1]));
        if (TANH) {
          const uint64_t K = pk2(-0.34657359f, -0.34657359f), HALF = pk2(0.5f, 0.5f);
          float f0, f1, g0, g1;
          upk2(mul2(FE, K), f0, f1);
          upk2(mul2(GE, K), g0, g1);
          const uint64_t TH = pk2(tanh_approx(f0), tanh_approx(f1)), TG = pk2(tanh_approx(g0), tanh_approx(g1));
          upk2(mul2(TH, fma2(TG, HALF, HALF)), z[4 * q + e], z[4 * q + e + 1]);
        } else {
          float fe0, fe1, ge0, ge1;
          upk2(FE, fe0, fe1);
          upk2(GE, ge0, ge1);
          fe0 = fminf(fe0, 28.853901f); fe1 = fminf(fe1, 28.853901f);
          ge0 = fminf(ge0, 25.968511f); ge1 = fminf(ge1, 25.968511f);
          const uint64_t A = pk2(ex2_approx(fe0), ex2_approx(fe1)), B = pk2(ex2_approx(ge0), ex2_approx(ge1));
          float d0, d1;
          upk2(mul2(add2(A, ONE), add2(B, ONE)), d0, d1);
          const float rinv = rcp_approx(d0 * d1);
          // (1 - a0) * d1 * rinv, (1 - a1) * d0 * rinv  (same association as the scalar form)
          upk2(mul2(mul2(fma2(A, M1, ONE), pk2(d1, d0)), pk2(rinv, rinv)), z[4 * q + e], z[4 * q + e + 1]);
        }
      }
    }
  }
}

// Logical 16-byte chunk `c` (0..7) of row `r` of a 128B-swizzled TMA box (rows are 128 B)
__device__ __forceinline__ float4* box_chunk(uint8_t* box_row, int r, int c) {
  return reinterpret_cast<float4*>(box_row + (((c ^ r) & 7) << 4));
}

// 32 staged floats of my box row -> 16 packed hi columns (+ 16 lo) at TMEM column taddr_hi / taddr_lo
template <bool BF16, bool SPLIT, bool PK = false>
__device__ __forceinline__ void tc_prep(uint8_t* box_row, int r, uint32_t taddr_hi, uint32_t taddr_lo) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 a = *box_chunk(box_row, r, c * 4 + q * 2), b = *box_chunk(box_row, r, c * 4 + q * 2 + 1);
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      split8x<BF16, SPLIT, PK>(v, hi + q * 4, lo + q * 4);
    }
    ptx::tmem_st8(taddr_hi + c * 8, hi);
    if (SPLIT) ptx::tmem_st8(taddr_lo + c * 8, lo);
  }
}
// Register re-partition between the warp roles (setmaxnreg works per warpgroup = 4 consecutive warps): the kernels
// launch with 640 threads x 96 registers = 61,440, which is the CTA's register pool -- setmaxnreg only moves registers
// INSIDE that pool (SASS: USETMAXREG.*.CTAPOOL); the SM's unallocated remainder is not available (round-2 measurement:
// a split that summed to 65,536 spun forever in TRY_ALLOC). The helper warpgroup (warps 16-19: MMA issuers, TMA
// producers) hands registers back, the four worker warpgroups take them: 512 x 112 + 128 x 32 = 61,440.
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
constexpr int TC_WORKER_REGS = 112, TC_HELPER_REGS = 32;
static_assert(TC_WORKER_REGS * 512 + TC_HELPER_REGS * 128 <= 96 * 640, "setmaxnreg redistributes the CTA's launch allocation only");
constexpr int TC_WORKER_WARPS = 16;                       // 2 tile slots x 2 channel halves x 4 lane quarters
constexpr int TC_MMA_WARP = 16;                           // 16, 17: MMA issuer of tile slot 0, 1
constexpr int TC_TMA_WARP = 18;                           // 18, 19: TMA producer of tile slot 0, 1
constexpr int TC_THREADS = (TC_WORKER_WARPS + 4) * 32;

// barrier block at the end of dynamic shared memory
struct TcBarriers {
  uint64_t w_ready;
  uint64_t x_full[2], y_full[2], c_full[2], x_free[2], y_free[2], a_ready[2], d1_ready[2], z_ready[2], d2_ready[2];
  uint32_t tmem_base;
  int mma_lock;             // the two MMA issuers take turns per GEMM (keeps the slots' phases staggered)
};

// The tensor pipe executes MMAs in issue order. If both slots' issuers interleave their GEMMs, both
// accumulators complete late and together, the two tiles' epilogues then collide on the MUFU/issue
// ports while the tensor pipe idles (measured: +8 % kernel time). A GEMM-granular lock restores the
// ping-pong: one slot's GEMM runs while the other slot is in its epilogue.
// (Only worth it when a GEMM is long: the single-pass bf16 mode runs without it.)
template <bool ENABLE>
__device__ __forceinline__ void tc_lock(int* lock) {
  if (ENABLE) {
    while (atomicCAS(lock, 0, 1) != 0) {
    }
  }
}
template <bool ENABLE>
__device__ __forceinline__ void tc_unlock(int* lock) {
  if (ENABLE) {
    __threadfence_block();
    atomicExch(lock, 0);
  }
}

// PK: packed fp32x2 epilogue arithmetic (bit-identical results, ~15 % fewer worker instructions, same speed:
// the worker phases are latency-bound, not issue-bound -- kept selectable for A/B runs).
template <bool BF16, bool SPLIT, bool PK = false, bool RG = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_layer_tc(const __grid_constant__ CUtensorMap map_in, TcLayerParams p) {
  using namespace ptx;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* smem = tc_smem;
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(smem + TC_SMEM_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int body = blockIdx.x & 1;
  const int cta_in_body = blockIdx.x >> 1, ctas_per_body = (gridDim.x + 1 - body) >> 1;
  const int tiles_body = p.N * p.tiles_per_utt;
  // this CTA's tiles: cta_in_body, +ctas_per_body, ...; local index; tile slot s takes local % 2 == s
  const int n_local = (tiles_body > cta_in_body) ? (tiles_body - cta_in_body + ctas_per_body - 1) / ctas_per_body : 0;

  pdl_launch_dependents();       // the next layer's CTAs may take over SMs as ours exit
  if (warp == TC_MMA_WARP) {
    if (lane == 0) {
      mbar_init(&bars->w_ready, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&bars->x_full[s], 1);
        mbar_init(&bars->y_full[s], 1);
        mbar_init(&bars->c_full[s], 1);
        constexpr uint32_t NW = 256;                    // worker threads that serve one tile slot
        mbar_init(&bars->x_free[s], NW);
        mbar_init(&bars->y_free[s], NW);
        mbar_init(&bars->a_ready[s], NW);
        mbar_init(&bars->d1_ready[s], 1);
        mbar_init(&bars->z_ready[s], NW);
        mbar_init(&bars->d2_ready[s], 1);
      }
      bars->mma_lock = 0;
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  // (RG: each role's code sits inside its own branch after the setmaxnreg -- code after a merge of the helper and worker
  //  paths would be allocated with the smaller of the two budgets)
  if (warp >= TC_MMA_WARP) {
   if (RG) reg_dec<TC_HELPER_REGS>();
   if (warp < TC_TMA_WARP) {
    // ======================= MMA issuers (one per tile slot); the first also loads the weights =======================
    if (elect_one()) {
      const int s = warp - TC_MMA_WARP;
      if (s == 0) {
        const uint8_t* img = body ? p.image[1] : p.image[0];
        mbar_arrive_expect_tx(&bars->w_ready, TC_IMAGE_BYTES);
        for (int off = 0; off < TC_IMAGE_BYTES; off += 16384) {
          const int n = min(16384, TC_IMAGE_BYTES - off);
          bulk_g2s(smem + off, img + off, n, &bars->w_ready);
        }
      }
      mbar_wait(&bars->w_ready, 0);
      const uint32_t w1hi = smem_u32(smem + TC_OFF_W1HI), w1lo = smem_u32(smem + TC_OFF_W1LO);
      const uint32_t w2hi = smem_u32(smem + TC_OFF_W2HI), w2lo = smem_u32(smem + TC_OFF_W2LO);
      constexpr uint32_t ID1 = idesc_f16(128, 128, BF16), ID2 = idesc_f16(128, 64, BF16);
      const uint32_t tD = tmem + s * 256;
      const uint32_t tAhi = tD + 128, tAlo = tD + 192;
      const int tiles_s = (n_local + 1 - s) / 2;
      for (int j = 0; j < tiles_s; ++j) {
        // D1 = A1lo.W1hi + A1hi.W1lo + A1hi.W1hi   (K = 128: 8 steps of 16; a step = 8 TMEM columns of A
        // and 2 K-chunks of 128 rows x 16 B of B)
        mbar_wait(&bars->a_ready[s], j & 1);
        tc_lock<SPLIT>(&bars->mma_lock);
        tc_fence_after_sync();
        TC_TRACE(2, j, s * 8 + 0);
        uint32_t acc = 0;
        if (SPLIT) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks, acc = 1)
            mma_f16_ts(tD, tAlo + ks * 8, smem_desc_kmajor_noswizzle(w1hi + ks * 2 * 2048, 2048, 128), ID1, acc);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(w1lo + ks * 2 * 2048, 2048, 128), ID1, 1);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks, acc = 1)
          mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(w1hi + ks * 2 * 2048, 2048, 128), ID1, acc);
        mma_commit(&bars->d1_ready[s]);
        tc_unlock<SPLIT>(&bars->mma_lock);
        TC_TRACE(2, j, s * 8 + 1);
        if (p.mode == 1) continue;
        // D2 = z.W2 (K = 64: 4 steps; chunk = 64 rows x 16 B)
        mbar_wait(&bars->z_ready[s], j & 1);
        tc_lock<SPLIT>(&bars->mma_lock);
        tc_fence_after_sync();
        TC_TRACE(2, j, s * 8 + 2);
        acc = 0;
        if (SPLIT) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks, acc = 1)
            mma_f16_ts(tD, tAlo + ks * 8, smem_desc_kmajor_noswizzle(w2hi + ks * 2 * 1024, 1024, 128), ID2, acc);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(w2lo + ks * 2 * 1024, 1024, 128), ID2, 1);
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks, acc = 1)
          mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(w2hi + ks * 2 * 1024, 1024, 128), ID2, acc);
        mma_commit(&bars->d2_ready[s]);
        tc_unlock<SPLIT>(&bars->mma_lock);
        TC_TRACE(2, j, s * 8 + 3);
      }
    }
    __syncwarp();
   } else {
    // ======================= TMA producers (one per tile slot): boxes in, output boxes out =======================
    if (elect_one()) {
      const int s = warp - TC_TMA_WARP;
      tma_prefetch_desc(&map_in);
      uint8_t* st = smem + TC_SMEM_STAGE0 + s * TC_STAGE_BYTES;
      const float* cbias = body ? p.cbias[1] : p.cbias[0];
      auto coords = [&](int j, int& n, int& t0) {
        const int tile = cta_in_body + (s + 2 * j) * ctas_per_body;
        n = tile / p.tiles_per_utt;
        t0 = (tile % p.tiles_per_utt) * TC_TM;
      };
      auto issue_x = [&](int j) {
        int n, t0;
        coords(j, n, t0);
        mbar_arrive_expect_tx(&bars->x_full[s], 2 * TC_BOX_BYTES);
        tma_load_3d(st, &map_in, 0, t0 - p.dilation, body * p.N + n, &bars->x_full[s]);
        tma_load_3d(st + TC_BOX_BYTES, &map_in, 32, t0 - p.dilation, body * p.N + n, &bars->x_full[s]);
      };
      auto issue_c = [&](int j) {     // conditioning rows of the frames the tile touches (contiguous)
        if (!p.cb_in_smem) return;
        int n, t0;
        coords(j, n, t0);
        const int f0 = (t0 + p.hop / 2) / p.hop, f1 = (min(t0 + TC_TM - 1, p.T - 1) + p.hop / 2) / p.hop;
        const uint32_t bytes = (uint32_t)(f1 - f0 + 1) * 512;
        mbar_arrive_expect_tx(&bars->c_full[s], bytes);
        bulk_g2s(smem + TC_SMEM_CB0 + s * TC_CB_BYTES, cbias + ((size_t)n * p.t_mel + f0) * 128, bytes, &bars->c_full[s]);
      };
      const int tiles_s = (n_local + 1 - s) / 2;
      // Tile handshake: tile (n, k) of this layer needs the previous layer's tiles k (its x[t] rows) and the one or
      // two tiles that hold rows t0-d .. t0+127-d; and it overwrites (ping-pong buffers) rows that the previous
      // layer's tiles k + d'/128 (+1) still read as THEIR x[t-d'] window. All of them processed by CTAs of the
      // previous kernel, which are running or done (1 CTA per SM: ours got its SM from one of them).
      auto wait_tiles = [&](int j) {
        if (!p.flags_in) return;
        int n, t0;
        coords(j, n, t0);
        const int k = t0 / TC_TM, last = p.tiles_per_utt - 1;
        const int* f = p.flags_in + (size_t)body * tiles_body + (size_t)n * p.tiles_per_utt;
        const int hi = t0 + TC_TM - 1 - p.dilation, lo = max(t0 - p.dilation, 0);
        const int k1 = hi >= 0 ? lo / TC_TM : k, k2 = hi >= 0 ? hi / TC_TM : k;
        const int k3 = min(k + p.prev_dilation / TC_TM, last), k4 = min(k + (p.prev_dilation + TC_TM - 1) / TC_TM, last);
        for (;;) {
          const int a = ld_relaxed_gpu(f + k), b = ld_relaxed_gpu(f + k1), c = ld_relaxed_gpu(f + k2);
          const int d = ld_relaxed_gpu(f + k3), e = ld_relaxed_gpu(f + k4);
          if (a & b & c & d & e) break;
        }
        fence_acq_rel_gpu();            // (acquire: the flagged tiles' rows are visible ...)
        fence_proxy_async_global();     // (... to the TMA loads issued next)
      };
      auto signal_tile = [&](int j) {   // after y_free[s]: all 256 workers of the slot have stored the tile's output
        if (!p.flags_out) return;
        const int tile = cta_in_body + (s + 2 * j) * ctas_per_body;
        fence_acq_rel_gpu();            // (release, cumulative over the workers' stores observed through y_free)
        st_relaxed_gpu(p.flags_out + (size_t)body * tiles_body + tile, 1);
      };
      if (!p.flags_in) pdl_wait_prior_grid();      // the previous kernel's output (and everything before it) is complete
      // Slot 1 lets slot 0's first boxes land before asking for its own: the TMA unit moves ~40 B/clk, so
      // interleaving both slots' 64 KB would deliver both at ~4k cycles; this way slot 0 starts at ~2k and
      // the two tiles begin half a phase apart, which is where the ping-pong wants them anyway.
      if (s == 1 && n_local > 0) mbar_wait(&bars->y_full[0], 0);
      if (tiles_s > 0) {
        int n, t0;
        coords(0, n, t0);
        wait_tiles(0);
        issue_x(0);
        mbar_arrive_expect_tx(&bars->y_full[s], 2 * TC_BOX_BYTES);
        tma_load_3d(st + 2 * TC_BOX_BYTES, &map_in, 0, t0, body * p.N + n, &bars->y_full[s]);
        tma_load_3d(st + 3 * TC_BOX_BYTES, &map_in, 32, t0, body * p.N + n, &bars->y_full[s]);
        issue_c(0);
      }
      for (int j = 0; j < tiles_s; ++j) {
        const bool more = j + 1 < tiles_s;
        if (more) {
          // the slot's x[t-d] boxes have been converted: refill them for the slot's next tile (whose inputs are
          // checked here, early in this tile's chain, so that the x[t] refill below never has to wait for them)
          mbar_wait(&bars->x_free[s], j & 1);
          wait_tiles(j + 1);
          issue_x(j + 1);
          TC_TRACE(3, j, s * 8 + 0);
        } else if (!p.flags_out) {
          break;
        }
        // the workers have copied the tile's output out of the x[t] boxes and are done with the
        // conditioning rows: refill both for the next tile, then publish the tile
        mbar_wait(&bars->y_free[s], j & 1);
        if (more) {
          int n, t0;
          coords(j + 1, n, t0);
          mbar_arrive_expect_tx(&bars->y_full[s], 2 * TC_BOX_BYTES);
          tma_load_3d(st + 2 * TC_BOX_BYTES, &map_in, 0, t0, body * p.N + n, &bars->y_full[s]);
          tma_load_3d(st + 3 * TC_BOX_BYTES, &map_in, 32, t0, body * p.N + n, &bars->y_full[s]);
          issue_c(j + 1);
          TC_TRACE(3, j, s * 8 + 2);
        }
        signal_tile(j);
      }
    }
    __syncwarp();
   }
  } else {
    if (RG) reg_inc<TC_WORKER_REGS>();
    // ======================= workers: operand prep, epilogues =======================
    // warp -> (tile slot, channel half, lane quarter); thread -> (row of the tile, 32 of the 64 channels)
    const int slot = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t tD = tmem + slot * 256 + lane_base;
    const uint32_t tAhi = tD + 128, tAlo = tD + 192;
    uint8_t* stage = smem + TC_SMEM_STAGE0 + slot * TC_STAGE_BYTES;
    uint8_t* my_x = stage + half * TC_BOX_BYTES + r * 128;                       // x[t-d], my 32 channels
    uint8_t* my_y = stage + 2 * TC_BOX_BYTES + half * TC_BOX_BYTES + r * 128;    // x[t]; later the output
    const float* bd_s = reinterpret_cast<const float*>(smem + TC_OFF_BD) + half * 32;
    const float* scal = reinterpret_cast<const float*>(smem + TC_OFF_SCAL);
    const bool tracer = (warp & 7) == 0 && lane == 0;
    bool weights_seen = false;
    float sf = 0.f, sg = 0.f, s2 = 0.f;

    int j = 0;
    for (int local = slot; local < n_local; local += 2, ++j) {
      const int tile = cta_in_body + local * ctas_per_body;
      const int n = tile / p.tiles_per_utt, t = (tile % p.tiles_per_utt) * TC_TM + r;
      const uint32_t par = j & 1;

      // ---- operand prep: A1 = [x[t-d] | x[t]] -> fp16 hi/lo -> TMEM (k = channel, +64 for the t tap)
      if (tracer) TC_TRACE(slot, j, 0);
      mbar_wait(&bars->x_full[slot], par);
      if (tracer) TC_TRACE(slot, j, 1);
      tc_prep<BF16, SPLIT, PK>(my_x, r, tAhi + half * 16, tAlo + half * 16);
      mbar_arrive(&bars->x_free[slot]);          // (release: my reads of the boxes are done)
      if (tracer) TC_TRACE(slot, j, 2);
      mbar_wait(&bars->y_full[slot], par);
      if (tracer) TC_TRACE(slot, j, 3);
      tc_prep<BF16, SPLIT, PK>(my_y, r, tAhi + 32 + half * 16, tAlo + 32 + half * 16);
      tmem_wait_st();
      tc_fence_before_sync();
      mbar_arrive(&bars->a_ready[slot]);
      if (tracer) TC_TRACE(slot, j, 4);

      if (!weights_seen) {              // scalars / bias live in the weight image
        mbar_wait(&bars->w_ready, 0);
        sf = scal[0]; sg = scal[1]; s2 = scal[2];
        weights_seen = true;
      }
      const int frame = (min(t, p.T - 1) + p.hop / 2) / p.hop;
      const float4* cb;
      if (p.cb_in_smem) {               // staged by the producer: row (frame - first frame of the tile)
        const int f0 = ((tile % p.tiles_per_utt) * TC_TM + p.hop / 2) / p.hop;
        cb = reinterpret_cast<const float4*>(smem + TC_SMEM_CB0 + slot * TC_CB_BYTES) + (frame - f0) * 32 + half * 8;
      } else {
        cb = reinterpret_cast<const float4*>((body ? p.cbias[1] : p.cbias[0]) + ((size_t)n * p.t_mel + frame) * 128) + half * 8;
      }

      // ---- epilogue 1: z = tanh(f) * sigmoid(g) on my 32 channels
      mbar_wait(&bars->d1_ready[slot], par);
      tc_fence_after_sync();
      if (p.cb_in_smem) mbar_wait(&bars->c_full[slot], par);
      if (tracer) TC_TRACE(slot, j, 5);
#pragma unroll
      for (int c = 0; c < 2; ++c) {          // 16 channels per TMEM round trip (32 + 32 at once spills at 96 registers)
        uint32_t fr[16], gr[16];
        tmem_ld16(tD + half * 32 + c * 16, fr);
        tmem_ld16(tD + 64 + half * 32 + c * 16, gr);
        tmem_wait_ld();
        float z[16];
        tc_gate<BF16, PK, 16>(fr, gr, cb + c * 4, cb + 16 + c * 4, sf, sg, z);
        if (p.mode == 1) {              // last layer: z itself is the output (x[t] is dead)
#pragma unroll
          for (int q = 0; q < 4; ++q) *box_chunk(my_y, r, c * 4 + q) = make_float4(z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
        } else {
          uint32_t hi[8], lo[8];
          float v0[8], v1[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { v0[e] = z[e]; v1[e] = z[8 + e]; }
          split8x<BF16, SPLIT, PK>(v0, hi, lo);
          split8x<BF16, SPLIT, PK>(v1, hi + 4, lo + 4);
          tmem_st8(tAhi + half * 16 + c * 8, hi);
          if (SPLIT) tmem_st8(tAlo + half * 16 + c * 8, lo);
        }
      }
      if (p.mode != 1) {
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->z_ready[slot]);
        if (tracer) TC_TRACE(slot, j, 6);

        // ---- epilogue 2: out = x[t] + D2 + b_dense (in place in my staged x[t] half row)
        mbar_wait(&bars->d2_ready[slot], par);
        tc_fence_after_sync();
        if (tracer) TC_TRACE(slot, j, 7);
        uint32_t dr[2][16];
        tmem_ld16(tD + half * 32, dr[0]);
        tmem_ld16(tD + half * 32 + 16, dr[1]);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(bd_s + q * 4);
          const float4 xv = *box_chunk(my_y, r, q);
          const uint32_t* d = &dr[q >> 2][(q & 3) * 4];
          float4 o;
          if (PK) {
            const uint64_t S2 = pk2(s2, s2);
            upk2(add2(pk2(xv.x, xv.y), fma2(pk2(__uint_as_float(d[0]), __uint_as_float(d[1])), S2, pk2(b.x, b.y))), o.x, o.y);
            upk2(add2(pk2(xv.z, xv.w), fma2(pk2(__uint_as_float(d[2]), __uint_as_float(d[3])), S2, pk2(b.z, b.w))), o.z, o.w);
          } else {
            o.x = xv.x + fmaf(__uint_as_float(d[0]), s2, b.x);
            o.y = xv.y + fmaf(__uint_as_float(d[1]), s2, b.y);
            o.z = xv.z + fmaf(__uint_as_float(d[2]), s2, b.z);
            o.w = xv.w + fmaf(__uint_as_float(d[3]), s2, b.w);
          }
          *box_chunk(my_y, r, q) = o;
        }
      }
      // ---- the tile's output sits in the x[t] boxes (my half: 128 rows x 128 B, swizzled). The four warps of
      //      this (slot, half) copy it out with full-line stores: a warp instruction writes 4 rows x 128 B, and each
      //      warp copies exactly the 32 rows its own lanes wrote.
      //      (Per-thread row stores and a TMA store were both measured slower: the former issues 32 partial
      //      lines per instruction, the latter holds the boxes ~1.5k cycles while the TMA unit drains them.)
      __syncwarp();     // rows 32*quarter .. +31 of the box were written by this warp's own lanes: no wider barrier needed
      {
        uint8_t* box = stage + 2 * TC_BOX_BYTES + half * TC_BOX_BYTES;
        const int t_first = (tile % p.tiles_per_utt) * TC_TM;
        float* out_tile = p.x_out + (((size_t)body * p.N + n) * p.T + t_first) * TC_C + half * 32;
        const int chunk = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = quarter * 32 + i * 4 + (lane >> 3);
          const float4 v = *box_chunk(box + row * 128, row, chunk);
          if (t_first + row < p.T) *reinterpret_cast<float4*>(out_tile + (size_t)row * TC_C + chunk * 4) = v;
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&bars->y_free[slot]);          // (release: my reads of the x[t] boxes and conditioning rows are done)
      if (tracer) TC_TRACE(slot, j, 8);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == TC_MMA_WARP) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// k_flow_tc: a RUN of consecutive gated layers of one flow (both bodies) in one persistent launch --
// the whole flow, or any segment of it down to a single layer (TcFlowParams::l0 / L).
//
// Same tile pipeline as k_layer_tc (tile slots, TMEM layout, staging, arithmetic -- the results are
// bit-identical), with the layer loop inside the kernel and NO grid-wide synchronisation between the
// layers of a launch: a tile of layer l+1 starts as soon as the few tiles of layer l it depends on have
// been published (per-tile flags, see k_layer_tc's tile handshake), whichever CTAs produced them.
// Measured reason (c2, profiles/r1_tc_trace_v14_layer2.txt): of the 83k cycles a layer takes as a
// kernel of its own only 63k are the steady-state pipeline; the rest is CTA exit -> launch -> prologue
// (barriers, TMEM, 80 KB of weights) -> first loads, and the half-phase stagger of the two slots at both
// ends. Here a CTA keeps walking its tile list layer after layer and the stagger survives the layer
// change. The first layer of a launch waits for the previous kernel as a whole (griddepcontrol.wait).
//
// Weights are single-buffered (80 KB; two layers do not fit next to 128 KB of staging). They are
// replaced in two parts by slot 0's helper: W1 (64 KB) + the bias/scale tail as soon as BOTH slots'
// last GEMM1 of the layer has completed (tcgen05.commit on w1_free), W2 (16 KB) after both slots' last
// GEMM2 -- by the time a slot's first tile of the next layer has been converted (1k cycles) W1 is
// there. The 512-byte tail (dense bias, epilogue scales) is double-buffered by layer parity because the
// workers still read it in the last tiles' residual step.
//
// Deadlock freedom: grid <= #SMs with one CTA per SM, so all CTAs are resident; every wait points at an
// earlier (layer, tile) of some CTA or at this CTA's own other roles, never forward.
//
// Two forms (template flag QUIET):
//   false  640 threads: 16 worker warps + per slot an MMA-issuer warp and a TMA-producer warp; every hand-off is
//          an mbarrier that the waiting warps poll (the form of k_layer_tc).
//   true   512 threads: the 16 worker warps only. The first warp of a slot (its "head" warp) also does the slot's
//          helper work with one thread, at the points where the slot's workers would be waiting anyway -- a slot's
//          events are strictly sequential (operands converted -> GEMM1 -> gate -> GEMM2 -> output copied out):
//          it issues the slot's MMAs and TMA loads, loads the weights (slot 0) and publishes the tiles. Warp-to-warp
//          hand-offs are hardware named barriers (the 7 other warps bar.arrive, the head warp bar.sync; nobody
//          polls); only completions of the async units (TMA bytes landed, tcgen05.commit) are mbarriers; the commit
//          barriers have ONE polling thread (the head warp's), which then joins the slot's bar.sync.
//          Measured reasons: (1) ncu source counters of the polled form (profiles/r1_flow_kernel.txt): 40 % of
//          all executed warp instructions were try_wait / branch / yield of polling warps (2,200 polls per tile);
//          (2) ptxas wants 122-128 registers for these kernels and gets 96 at 640 threads (the register file is
//          allocated per 128 threads: 544..640 threads all cap at 96); 512 threads lift the cap to 128
//          (profiles/r1_experiments_after_flow_kernel.txt).
//          Off the critical path by construction: the next tile's flags are LOADED before the GEMM1 issue loop and
//          examined after it; a tile's publication (gpu-scope release fence) is deferred to the next tile's GEMM1
//          window unless the head thread is about to block on other CTAs' tiles.
// ------------------------------------------------------------------------------------------------
constexpr int TCF_THREADS = TC_THREADS;                    // polled form (a 21st warp would cap the kernel at 80 registers)
constexpr int TCF_THREADS_QUIET = TC_WORKER_WARPS * 32;
constexpr int tcf_threads(bool quiet) { return quiet ? TCF_THREADS_QUIET : TCF_THREADS; }
constexpr int TCF_MAIN_BYTES = TC_OFF_BD;                  // W1hi | W1lo | W2hi | W2lo
constexpr int TCF_TAIL_BYTES = 512;                        // dense bias (256 B) + scales (256 B)
constexpr int TCF_SMEM_TAIL0 = TC_SMEM_CB0 + 2 * TC_CB_BYTES;
constexpr int TCF_SMEM_BARS = TCF_SMEM_TAIL0 + 2 * TCF_TAIL_BYTES;
constexpr int TCF_SMEM_BYTES = TCF_SMEM_BARS + 512;
constexpr int TCF_MAX_LAYERS = 64;

struct TcFlowParams {
  float* act[2];            // ping/pong [2][N][T][64]; layer l of the launch reads act[(cur0 + l) & 1] through map[(cur0 + l) & 1], writes the other
  const uint8_t* images;    // image of (flow, body 0, layer 0); (body b, flow layer g) at + (b * L_total + g) * TC_IMAGE_BYTES
  const float* cbias;       // [2][L_total][N][t_mel][128] (pre-scaled, as for k_layer_tc)
  int* flags;               // [L_total][2][N * tiles_per_utt], zeroed before the forward
  int N, T, t_mel, hop, cur0, tiles_per_utt, cb_in_smem;
  int L_total;              // gated layers of the flow
  int l0, L;                // this launch runs the flow's layers l0 .. l0 + L - 1
  int final_layer;          // 1: layer l0 + L - 1 is the flow's last gated layer (its output is z, no dense GEMM)
  int dilation[TCF_MAX_LAYERS];   // of the flow's layers
  long long* trace;         // debug timeline of CTA 0 for flow layer trace_layer (or nullptr)
  int trace_layer;
  int stagger;              // see the producers
  int rotate;               // 1: rotate the tile-to-CTA assignment from layer to layer (see cta_of in the kernel)
};

struct TcFlowBarriers {
  uint64_t w1_ready, w2_ready, w1_free, w2_free;
  uint64_t x_full[2], y_full[2], c_full[2], x_free[2], y_free[2], a_ready[2], d1_ready[2], z_ready[2], d2_ready[2];
  uint32_t tmem_base;
  int mma_lock;
};

#define TCF_TRACE(role, l, j, k)                                                                    \
  do {                                                                                              \
    if (p.trace && blockIdx.x == 0 && p.l0 + (l) == p.trace_layer && (j) < 16) p.trace[((role) * 16 + (j)) * 16 + (k)] = clock64(); \
  } while (0)

// named barriers of the QUIET form (+ slot): a barrier is reused tile after tile -- every arrival of tile j+1 is
// causally after the completion of tile j's barrier
constexpr int TCF_NB_AREADY = 1, TCF_NB_ZREADY = 3, TCF_NB_D1 = 5, TCF_NB_D2 = 7, TCF_NB_YFREE = 9;
constexpr int TCF_NB_COUNT = 256;              // the slot's 8 worker warps

template <bool BF16, bool SPLIT, bool PK = false, bool QUIET = false, bool RG = false>
__global__ void __launch_bounds__(tcf_threads(QUIET), 1)
k_flow_tc(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const __grid_constant__ TcFlowParams p) {
  using namespace ptx;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* smem = tc_smem;
  TcFlowBarriers* bars = reinterpret_cast<TcFlowBarriers*>(smem + TCF_SMEM_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int body = blockIdx.x & 1;
  const int cta_in_body = blockIdx.x >> 1, ctas_per_body = (gridDim.x + 1 - body) >> 1;
  const int tiles_body = p.N * p.tiles_per_utt;
  const int n_local = (tiles_body > cta_in_body) ? (tiles_body - cta_in_body + ctas_per_body - 1) / ctas_per_body : 0;
  const int L = p.L;
  // Tile-to-CTA assignment. Layer l's tiles c, c + P, c + 2P, ... (P = CTAs per body) belong to the CTA whose
  // virtual index for that layer is c. With tiles_body = a P + r the r lowest virtual indices carry a + 1 tiles;
  // a fixed assignment makes the same r CTAs the critical path of EVERY layer (c2: 14 vs 13 tiles, 3.6 %).
  // Rotating the virtual index by r per layer slides the window of long CTAs once around the grid, so that over
  // the flow every CTA walks the same number of tiles (+-1). The per-tile flags make any assignment correct.
  // Only when every CTA keeps >= 1 tile per slot in every layer (the weight hand-over counts on both slots).
  const int rot = (p.rotate && tiles_body >= 2 * ctas_per_body) ? tiles_body % ctas_per_body : 0;
  auto cta_of = [&](int l) { return rot ? (cta_in_body + (p.l0 + l) * rot) % ctas_per_body : cta_in_body; };
  auto tiles_in = [&](int l, int s) {            // tiles of slot s of this CTA in layer l (of the launch)
    const int c = cta_of(l);
    const int nl = (tiles_body > c) ? (tiles_body - c + ctas_per_body - 1) / ctas_per_body : 0;
    return (nl + 1 - s) / 2;
  };
  auto is_last = [&](int l) { return p.final_layer && l == L - 1; };

  // ---- TMA side of a slot (one elected thread of its producer / helper warp)
  auto coords = [&](int s, int l, int j, int& n, int& t0) {
    const int tile = cta_of(l) + (s + 2 * j) * ctas_per_body;
    n = tile / p.tiles_per_utt;
    t0 = (tile % p.tiles_per_utt) * TC_TM;
  };
  auto map_of = [&](int l) { return ((p.cur0 + l) & 1) ? &map1 : &map0; };
  auto issue_x = [&](int s, int l, int j) {
    uint8_t* st = smem + TC_SMEM_STAGE0 + s * TC_STAGE_BYTES;
    int n, t0;
    coords(s, l, j, n, t0);
    const int d = p.dilation[p.l0 + l];
    mbar_arrive_expect_tx(&bars->x_full[s], 2 * TC_BOX_BYTES);
    tma_load_3d(st, map_of(l), 0, t0 - d, body * p.N + n, &bars->x_full[s]);
    tma_load_3d(st + TC_BOX_BYTES, map_of(l), 32, t0 - d, body * p.N + n, &bars->x_full[s]);
  };
  auto issue_y = [&](int s, int l, int j) {     // x[t] boxes + the conditioning rows of the frames the tile touches
    uint8_t* st = smem + TC_SMEM_STAGE0 + s * TC_STAGE_BYTES;
    int n, t0;
    coords(s, l, j, n, t0);
    mbar_arrive_expect_tx(&bars->y_full[s], 2 * TC_BOX_BYTES);
    tma_load_3d(st + 2 * TC_BOX_BYTES, map_of(l), 0, t0, body * p.N + n, &bars->y_full[s]);
    tma_load_3d(st + 3 * TC_BOX_BYTES, map_of(l), 32, t0, body * p.N + n, &bars->y_full[s]);
    if (p.cb_in_smem) {
      const float* cbias = p.cbias + ((size_t)body * p.L_total + p.l0 + l) * p.N * p.t_mel * 128;
      const int f0 = (t0 + p.hop / 2) / p.hop, f1 = (min(t0 + TC_TM - 1, p.T - 1) + p.hop / 2) / p.hop;
      const uint32_t bytes = (uint32_t)(f1 - f0 + 1) * 512;
      mbar_arrive_expect_tx(&bars->c_full[s], bytes);
      bulk_g2s(smem + TC_SMEM_CB0 + s * TC_CB_BYTES, cbias + ((size_t)n * p.t_mel + f0) * 128, bytes, &bars->c_full[s]);
    }
  };
  // the tiles of layer l-1 that tile j of layer l reads or whose reads it overwrites (see k_layer_tc): published?
  // (layer 0 of the launch follows a kernel boundary: everything before it is complete)
  struct Probe { int a, b, c, d, e; };
  auto probe_load = [&](int s, int l, int j) -> Probe {      // relaxed loads of the (up to) five flags
    Probe f = {1, 1, 1, 1, 1};
    if (l == 0) return f;
    int n, t0;
    coords(s, l, j, n, t0);
    const int d = p.dilation[p.l0 + l], dp = p.dilation[p.l0 + l - 1];
    const int k = t0 / TC_TM, last = p.tiles_per_utt - 1;
    const int* fl = p.flags + ((size_t)(p.l0 + l - 1) * 2 + body) * tiles_body + (size_t)n * p.tiles_per_utt;
    const int hi = t0 + TC_TM - 1 - d, lo = max(t0 - d, 0);
    const int k1 = hi >= 0 ? lo / TC_TM : k, k2 = hi >= 0 ? hi / TC_TM : k;
    const int k3 = min(k + dp / TC_TM, last), k4 = min(k + (dp + TC_TM - 1) / TC_TM, last);
    f.a = ld_relaxed_gpu(fl + k); f.b = ld_relaxed_gpu(fl + k1); f.c = ld_relaxed_gpu(fl + k2);
    f.d = ld_relaxed_gpu(fl + k3); f.e = ld_relaxed_gpu(fl + k4);
    return f;
  };
  auto probe_test = [&](int l, const Probe& f) -> bool {
    if (l == 0) return true;
    if (!(f.a & f.b & f.c & f.d & f.e)) return false;
    fence_acq_rel_gpu();            // (acquire: the published tiles' rows are visible ...)
    fence_proxy_async_global();     // (... to the TMA loads issued next)
    return true;
  };
  auto tiles_ready = [&](int s, int l, int j) -> bool { return probe_test(l, probe_load(s, l, j)); };
  auto publish = [&](int s, int l, int j) {   // after the slot's 256 workers have stored the tile's output (observed through a barrier)
    const int tile = cta_of(l) + (s + 2 * j) * ctas_per_body;
    fence_acq_rel_gpu();               // (release, cumulative over the workers' stores)
    st_relaxed_gpu(p.flags + ((size_t)(p.l0 + l) * 2 + body) * tiles_body + tile, 1);
  };
  // ---- tensor-core side of a slot (one elected thread of its issuer / helper warp)
  const uint32_t w1hi = smem_u32(smem + TC_OFF_W1HI), w1lo = smem_u32(smem + TC_OFF_W1LO);
  const uint32_t w2hi = smem_u32(smem + TC_OFF_W2HI), w2lo = smem_u32(smem + TC_OFF_W2LO);
  constexpr uint32_t ID1 = idesc_f16(128, 128, BF16), ID2 = idesc_f16(128, 64, BF16);
  auto load_w1 = [&](int l) {     // W1 + bias/scale tail of layer l, once BOTH slots' last GEMM1 of layer l-1 has completed
    const uint8_t* img = p.images + ((size_t)body * p.L_total + p.l0 + l) * TC_IMAGE_BYTES;
    if (l > 0) mbar_wait(&bars->w1_free, (l - 1) & 1);
    mbar_arrive_expect_tx(&bars->w1_ready, 2 * TC_W1_BYTES + TCF_TAIL_BYTES);
    for (int off = 0; off < 2 * TC_W1_BYTES; off += 16384) bulk_g2s(smem + off, img + off, 16384, &bars->w1_ready);
    bulk_g2s(smem + TCF_SMEM_TAIL0 + (l & 1) * TCF_TAIL_BYTES, img + TC_OFF_BD, TCF_TAIL_BYTES, &bars->w1_ready);
  };
  auto load_w2 = [&](int l) {     // W2 of layer l, once both slots' last GEMM2 of layer l-1 has completed
    const uint8_t* img = p.images + ((size_t)body * p.L_total + p.l0 + l) * TC_IMAGE_BYTES;
    if (l > 0) mbar_wait(&bars->w2_free, (l - 1) & 1);
    mbar_arrive_expect_tx(&bars->w2_ready, 2 * TC_W2_BYTES);
    bulk_g2s(smem + TC_OFF_W2HI, img + TC_OFF_W2HI, 2 * TC_W2_BYTES, &bars->w2_ready);
  };
  auto idle_layer = [&](int l) {  // a slot without tiles only keeps the weight hand-over going: one arrival per barrier phase
    mbar_wait(&bars->w1_ready, l & 1);
    mbar_arrive(&bars->w1_free);
    if (!is_last(l)) {
      mbar_wait(&bars->w2_ready, l & 1);
      mbar_arrive(&bars->w2_free);
    }
  };
  auto gemm1 = [&](int s, bool last_tile) {
    const uint32_t tD = bars->tmem_base + s * 256, tAhi = tD + 128, tAlo = tD + 192;
    tc_lock<SPLIT>(&bars->mma_lock);
    tc_fence_after_sync();
    uint32_t acc = 0;
    if (SPLIT) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks, acc = 1)
        mma_f16_ts(tD, tAlo + ks * 8, smem_desc_kmajor_noswizzle(w1hi + ks * 2 * 2048, 2048, 128), ID1, acc);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(w1lo + ks * 2 * 2048, 2048, 128), ID1, 1);
    }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks, acc = 1)
      mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(w1hi + ks * 2 * 2048, 2048, 128), ID1, acc);
    mma_commit(&bars->d1_ready[s]);
    if (last_tile) mma_commit(&bars->w1_free);     // this slot is done with W1 of the layer
    tc_unlock<SPLIT>(&bars->mma_lock);
  };
  auto gemm2 = [&](int s, bool last_tile) {
    const uint32_t tD = bars->tmem_base + s * 256, tAhi = tD + 128, tAlo = tD + 192;
    tc_lock<SPLIT>(&bars->mma_lock);
    tc_fence_after_sync();
    uint32_t acc = 0;
    if (SPLIT) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks, acc = 1)
        mma_f16_ts(tD, tAlo + ks * 8, smem_desc_kmajor_noswizzle(w2hi + ks * 2 * 1024, 1024, 128), ID2, acc);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(w2lo + ks * 2 * 1024, 1024, 128), ID2, 1);
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks, acc = 1)
      mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(w2hi + ks * 2 * 1024, 1024, 128), ID2, acc);
    mma_commit(&bars->d2_ready[s]);
    if (last_tile) mma_commit(&bars->w2_free);
    tc_unlock<SPLIT>(&bars->mma_lock);
  };

  constexpr int SETUP_WARP = QUIET ? 0 : TC_MMA_WARP;
  pdl_launch_dependents();
  if (warp == SETUP_WARP) {
    if (lane == 0) {
      mbar_init(&bars->w1_ready, 1);
      mbar_init(&bars->w2_ready, 1);
      mbar_init(&bars->w1_free, 2);
      mbar_init(&bars->w2_free, 2);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&bars->x_full[s], 1);
        mbar_init(&bars->y_full[s], 1);
        mbar_init(&bars->c_full[s], 1);
        mbar_init(&bars->x_free[s], 256);
        mbar_init(&bars->y_free[s], 256);
        mbar_init(&bars->a_ready[s], 256);
        mbar_init(&bars->d1_ready[s], 1);
        mbar_init(&bars->z_ready[s], 256);
        mbar_init(&bars->d2_ready[s], 1);
      }
      bars->mma_lock = 0;
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  if (!QUIET && warp >= TC_MMA_WARP) {
   if (RG) reg_dec<TC_HELPER_REGS>();
   if (warp < TC_TMA_WARP) {
    // ======================= MMA issuers (one per tile slot), polled form; slot 0's also hands the weights over =======================
    if (elect_one() && n_local > 0) {
      const int s = warp - TC_MMA_WARP;
      uint32_t it = 0;                                  // tiles of this slot so far (barrier phase)
      for (int l = 0; l < L; ++l) {
        const bool last_layer = is_last(l);
        const int tiles_s = tiles_in(l, s);
        if (s == 0) load_w1(l);
        if (tiles_s == 0) {
          idle_layer(l);
          continue;
        }
        for (int j = 0; j < tiles_s; ++j, ++it) {
          mbar_wait(&bars->a_ready[s], it & 1);
          if (j == 0) mbar_wait(&bars->w1_ready, l & 1);
          TCF_TRACE(2, l, j, s * 8 + 0);
          gemm1(s, j == tiles_s - 1);
          TCF_TRACE(2, l, j, s * 8 + 1);
          if (last_layer) continue;
          if (s == 0 && j == 0) load_w2(l);
          mbar_wait(&bars->z_ready[s], it & 1);
          if (j == 0) mbar_wait(&bars->w2_ready, l & 1);
          TCF_TRACE(2, l, j, s * 8 + 2);
          gemm2(s, j == tiles_s - 1);
          TCF_TRACE(2, l, j, s * 8 + 3);
        }
      }
    }
    __syncwarp();
   } else {
    // ======================= TMA producers (one per tile slot), polled form =======================
    if (elect_one()) {
      const int s = warp - TC_TMA_WARP;
      tma_prefetch_desc(&map0);
      tma_prefetch_desc(&map1);
      pdl_wait_prior_grid();      // the launch's input (k_front / the previous layers' kernel) is complete
      // half-phase stagger of the two slots: slot 1 starts loading when slot 0's first boxes have landed (0), when
      // its first GEMM1 has completed (1) or when its first gate phase is through (2)
      if (s == 1 && n_local > 0) {
        if (p.stagger == 1) mbar_wait(&bars->d1_ready[0], 0);
        else if (p.stagger == 2) mbar_wait(&bars->z_ready[0], 0);
        else mbar_wait(&bars->y_full[0], 0);
      }
      // this slot's tile list: layer after layer, tiles_in(l, s) tiles each (the same count in every layer unless
      // the assignment rotates; a slot without tiles in layer 0 has none at all)
      int l = 0, j = 0, tiles_l = tiles_in(0, s);       // the tile in flight and its layer's tile count
      if (tiles_l > 0) {
        issue_x(s, 0, 0);
        issue_y(s, 0, 0);
      }
      // A tile is published as soon as it is stored, NEVER after a wait for other CTAs' tiles (two CTAs whose next
      // tiles need each other's current tiles would wait forever): the inputs of the slot's next tile are only
      // PROBED early (to prefetch its x[t-d] boxes, the normal case inside a layer); if they are not all there yet
      // the blocking wait comes after this tile's publication.
      for (uint32_t q = 0; tiles_l > 0; ++q) {          // q: tiles of this slot so far (barrier phase)
        const int ln = (j + 1 == tiles_l) ? l + 1 : l, jn = (j + 1 == tiles_l) ? 0 : j + 1;       // the slot's next tile
        if (ln == L) break;                             // (the launch's last tile: nothing to refill; published by the kernel's end)
        mbar_wait(&bars->x_free[s], q & 1);             // the slot's x[t-d] boxes have been converted
        const bool early = tiles_ready(s, ln, jn);
        if (early) {
          issue_x(s, ln, jn);
          TCF_TRACE(3, l, j, s * 8 + 0);
        }
        mbar_wait(&bars->y_free[s], q & 1);             // output copied out of the x[t] boxes, conditioning rows read
        if (early) {
          issue_y(s, ln, jn);
          if (l < L - 1) publish(s, l, j);
        } else {
          if (l < L - 1) publish(s, l, j);
          while (!tiles_ready(s, ln, jn)) {
          }
          issue_x(s, ln, jn);
          issue_y(s, ln, jn);
        }
        TCF_TRACE(3, l, j, s * 8 + 2);
        if (ln != l) tiles_l = tiles_in(ln, s);
        l = ln;
        j = jn;
      }
    }
    __syncwarp();
   }
  } else {
    if (RG && !QUIET) reg_inc<TC_WORKER_REGS>();
    // ======================= workers: operand prep, epilogues =======================
    const int slot = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t tD = tmem + slot * 256 + lane_base;
    const uint32_t tAhi = tD + 128, tAlo = tD + 192;
    uint8_t* stage = smem + TC_SMEM_STAGE0 + slot * TC_STAGE_BYTES;
    uint8_t* my_x = stage + half * TC_BOX_BYTES + r * 128;
    uint8_t* my_y = stage + 2 * TC_BOX_BYTES + half * TC_BOX_BYTES + r * 128;
    const bool tracer = (warp & 7) == 0 && lane == 0;
    // QUIET form: the slot's head warp (its first one) does the slot's helper work with its lane 0
    const bool head_warp = QUIET && (warp & 7) == 0, head = head_warp && lane == 0;
    int pub_l = -1, pub_j = 0;                          // head: tile whose publication is pending
    if (head) {
      tma_prefetch_desc(&map0);
      tma_prefetch_desc(&map1);
      pdl_wait_prior_grid();      // the launch's input (k_front / the previous layers' kernel) is complete
      // half-phase stagger of the two slots: slot 1 starts loading when slot 0's first boxes have landed (0) or
      // when its first GEMM1 has completed (1, 2)
      if (slot == 1 && n_local > 0) {
        if (p.stagger >= 1) mbar_wait(&bars->d1_ready[0], 0);
        else mbar_wait(&bars->y_full[0], 0);
      }
      if (tiles_in(0, slot) > 0) {
        issue_x(slot, 0, 0);
        issue_y(slot, 0, 0);
      }
    }
    if (QUIET) __syncwarp();
    uint32_t it = 0;
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const bool last_layer = is_last(l);
      const int tiles_s = tiles_in(l, slot), cta_l = cta_of(l);
      if (head && n_local > 0) {        // weights (both W1 barriers' previous phases are normally long complete here)
        if (slot == 0) load_w1(l);
        if (tiles_s == 0) idle_layer(l);
      }
      if (QUIET) __syncwarp();
      const float* tail = reinterpret_cast<const float*>(smem + TCF_SMEM_TAIL0 + (l & 1) * TCF_TAIL_BYTES);
      const float* bd_s = tail + half * 32;
      const float* cbias_l = p.cbias + ((size_t)body * p.L_total + p.l0 + l) * p.N * p.t_mel * 128;
      float* x_out = p.act[(p.cur0 + l + 1) & 1];
      float sf = 0.f, sg = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int j = 0; j < tiles_s; ++j, ++it) {
        const int tile = cta_l + (slot + 2 * j) * ctas_per_body;
        const int n = tile / p.tiles_per_utt, t = (tile % p.tiles_per_utt) * TC_TM + r;
        const uint32_t par = it & 1;

        if (tracer) TCF_TRACE(slot, l, j, 0);
        mbar_wait(&bars->x_full[slot], par);
        if (tracer) TCF_TRACE(slot, l, j, 1);
        tc_prep<BF16, SPLIT, PK>(my_x, r, tAhi + half * 16, tAlo + half * 16);
        if (!QUIET) mbar_arrive(&bars->x_free[slot]);   // (QUIET: the helper refills these boxes after GEMM1, which implies this)
        if (tracer) TCF_TRACE(slot, l, j, 2);
        mbar_wait(&bars->y_full[slot], par);
        if (tracer) TCF_TRACE(slot, l, j, 3);
        tc_prep<BF16, SPLIT, PK>(my_y, r, tAhi + 32 + half * 16, tAlo + 32 + half * 16);
        tmem_wait_st();
        tc_fence_before_sync();
        // the slot's next tile (l, j) -> (ln, jn) (head only; a layer's tile count never drops to 0 once the slot has tiles)
        const int ln = (j + 1 == tiles_s) ? l + 1 : l, jn = (j + 1 == tiles_s) ? 0 : j + 1;
        const bool more = ln < L;
        bool early = false;
        if (!QUIET) {
          mbar_arrive(&bars->a_ready[slot]);
        } else if (!head_warp) {
          named_bar_arrive(TCF_NB_AREADY + slot, TCF_NB_COUNT);
        } else {
          named_bar_sync(TCF_NB_AREADY + slot, TCF_NB_COUNT);      // the slot's A1 operands are in TMEM (x[t-d] boxes free)
          if (head) {
            if (j == 0) mbar_wait(&bars->w1_ready, l & 1);
            // the next tile's flags: loaded now, examined after the MMA issue loop (which the tensor pipe throttles)
            Probe pr = {1, 1, 1, 1, 1};
            if (more) pr = probe_load(slot, ln, jn);
            TCF_TRACE(2, l, j, slot * 8 + 0);
            gemm1(slot, j == tiles_s - 1);
            TCF_TRACE(2, l, j, slot * 8 + 1);
            if (!last_layer && slot == 0 && j == 0) load_w2(l);
            // prefetch the next tile's x[t-d] boxes if its inputs are published (only PROBED: a tile is published NEVER
            // after a wait for other CTAs' tiles -- two CTAs whose next tiles need each other's current tiles would
            // wait forever -- so the blocking wait comes after this tile's publication, below)
            if (more) {
              early = probe_test(ln, pr);
              if (early) {
                issue_x(slot, ln, jn);
                TCF_TRACE(3, l, j, slot * 8 + 0);
              }
            }
            if (pub_l >= 0) {                                      // the previous tile's deferred publication
              publish(slot, pub_l, pub_j);
              pub_l = -1;
            }
            mbar_wait(&bars->d1_ready[slot], par);                 // GEMM1 complete
          }
          __syncwarp();
        }
        if (tracer) TCF_TRACE(slot, l, j, 4);

        if (j == 0) {                     // the layer's scales / dense bias arrive with W1
          mbar_wait(&bars->w1_ready, l & 1);
          sf = tail[64]; sg = tail[65]; s2 = tail[66];
        }
        const int frame = (min(t, p.T - 1) + p.hop / 2) / p.hop;
        const float4* cb;
        if (p.cb_in_smem) {
          const int f0 = ((tile % p.tiles_per_utt) * TC_TM + p.hop / 2) / p.hop;
          cb = reinterpret_cast<const float4*>(smem + TC_SMEM_CB0 + slot * TC_CB_BYTES) + (frame - f0) * 32 + half * 8;
        } else {
          cb = reinterpret_cast<const float4*>(cbias_l + ((size_t)n * p.t_mel + frame) * 128) + half * 8;
        }

        // ---- epilogue 1: z = tanh(f) * sigmoid(g) on my 32 channels
        if (QUIET) named_bar_sync(TCF_NB_D1 + slot, TCF_NB_COUNT);     // (the slot's issuer saw GEMM1's commit)
        else mbar_wait(&bars->d1_ready[slot], par);
        tc_fence_after_sync();
        if (p.cb_in_smem) mbar_wait(&bars->c_full[slot], par);
        if (tracer) TCF_TRACE(slot, l, j, 5);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t fr[16], gr[16];
          tmem_ld16(tD + half * 32 + c * 16, fr);
          tmem_ld16(tD + 64 + half * 32 + c * 16, gr);
          tmem_wait_ld();
          if (tracer) TCF_TRACE(slot, l, j, 12 + c);
          float z[16];
          tc_gate<BF16, PK, 16>(fr, gr, cb + c * 4, cb + 16 + c * 4, sf, sg, z);
          if (last_layer) {               // z itself is the output (x[t] is dead)
#pragma unroll
            for (int q = 0; q < 4; ++q) *box_chunk(my_y, r, c * 4 + q) = make_float4(z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
          } else {
            uint32_t hi[8], lo[8];
            float v0[8], v1[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { v0[e] = z[e]; v1[e] = z[8 + e]; }
            split8x<BF16, SPLIT, PK>(v0, hi, lo);
            split8x<BF16, SPLIT, PK>(v1, hi + 4, lo + 4);
            tmem_st8(tAhi + half * 16 + c * 8, hi);
            if (SPLIT) tmem_st8(tAlo + half * 16 + c * 8, lo);
          }
        }
        if (!last_layer) {
          tmem_wait_st();
          tc_fence_before_sync();
          if (!QUIET) {
            mbar_arrive(&bars->z_ready[slot]);
          } else if (!head_warp) {
            named_bar_arrive(TCF_NB_ZREADY + slot, TCF_NB_COUNT);
          } else {
            named_bar_sync(TCF_NB_ZREADY + slot, TCF_NB_COUNT);    // z is in TMEM
            if (head) {
              if (j == 0) mbar_wait(&bars->w2_ready, l & 1);
              TCF_TRACE(2, l, j, slot * 8 + 2);
              gemm2(slot, j == tiles_s - 1);
              TCF_TRACE(2, l, j, slot * 8 + 3);
              mbar_wait(&bars->d2_ready[slot], par);
            }
            __syncwarp();
          }
          if (tracer) TCF_TRACE(slot, l, j, 6);

          // ---- epilogue 2: out = x[t] + D2 + b_dense (in place in my staged x[t] half row)
          if (QUIET) named_bar_sync(TCF_NB_D2 + slot, TCF_NB_COUNT);
          else mbar_wait(&bars->d2_ready[slot], par);
          tc_fence_after_sync();
          if (tracer) TCF_TRACE(slot, l, j, 7);
          uint32_t dr[2][16];
          tmem_ld16(tD + half * 32, dr[0]);
          tmem_ld16(tD + half * 32 + 16, dr[1]);
          tmem_wait_ld();
          if (tracer) TCF_TRACE(slot, l, j, 9);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(bd_s + q * 4);
            const float4 xv = *box_chunk(my_y, r, q);
            const uint32_t* d = &dr[q >> 2][(q & 3) * 4];
            float4 o;
            if (PK) {
              const uint64_t S2 = pk2(s2, s2);
              upk2(add2(pk2(xv.x, xv.y), fma2(pk2(__uint_as_float(d[0]), __uint_as_float(d[1])), S2, pk2(b.x, b.y))), o.x, o.y);
              upk2(add2(pk2(xv.z, xv.w), fma2(pk2(__uint_as_float(d[2]), __uint_as_float(d[3])), S2, pk2(b.z, b.w))), o.z, o.w);
            } else {
              o.x = xv.x + fmaf(__uint_as_float(d[0]), s2, b.x);
              o.y = xv.y + fmaf(__uint_as_float(d[1]), s2, b.y);
              o.z = xv.z + fmaf(__uint_as_float(d[2]), s2, b.z);
              o.w = xv.w + fmaf(__uint_as_float(d[3]), s2, b.w);
            }
            *box_chunk(my_y, r, q) = o;
          }
        }
        // ---- copy-out (see k_layer_tc)
        if (tracer) TCF_TRACE(slot, l, j, 10);
        __syncwarp();
        {
          uint8_t* box = stage + 2 * TC_BOX_BYTES + half * TC_BOX_BYTES;
          const int t_first = (tile % p.tiles_per_utt) * TC_TM;
          float* out_tile = x_out + (((size_t)body * p.N + n) * p.T + t_first) * TC_C + half * 32;
          const int chunk = lane & 7;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = quarter * 32 + i * 4 + (lane >> 3);
            const float4 v = *box_chunk(box + row * 128, row, chunk);
            if (t_first + row < p.T) *reinterpret_cast<float4*>(out_tile + (size_t)row * TC_C + chunk * 4) = v;
          }
        }
        if (tracer) TCF_TRACE(slot, l, j, 11);
        tc_fence_before_sync();
        if (!QUIET) {
          mbar_arrive(&bars->y_free[slot]);
        } else if (!head_warp) {
          named_bar_arrive(TCF_NB_YFREE + slot, TCF_NB_COUNT);
        } else {
          named_bar_sync(TCF_NB_YFREE + slot, TCF_NB_COUNT);       // output copied out of the x[t] boxes, conditioning rows read
          if (head) {
            if (more && early) issue_y(slot, ln, jn);
            if (l < L - 1) {                  // a later layer of this launch reads the tile
              if (more && !early) publish(slot, l, j);             // about to block on other CTAs' tiles: publish first
              else { pub_l = l; pub_j = j; }                       // otherwise in the next tile's GEMM1 window
            }
            if (more && !early) {
              while (!tiles_ready(slot, ln, jn)) {
              }
              issue_x(slot, ln, jn);
              issue_y(slot, ln, jn);
            }
            TCF_TRACE(3, l, j, slot * 8 + 2);
          }
          __syncwarp();
        }
        if (tracer) TCF_TRACE(slot, l, j, 8);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == SETUP_WARP) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// Post-net of both bodies on the tensor cores (reference modules.py:145-165, use_skip_connection
// False): per 128-row tile
//   D [128 x 128] = z . Ws                     (the last layer's skip 1x1; z from k_layer_tc mode 1)
//   h             = relu(D + bs)          -> A operand of the next GEMM (fp16 hi/lo in TMEM)
//   D'[128 x 128] = h . W1                     (postprocess1)
//   y             = relu(D' + b1) . w2 + b2    (postprocess2: 128 -> 1, in-thread dot product; the two
//                                               column halves of a row meet in one atomicAdd each on a
//                                               zeroed output -- two addends, so the sum is order-free)
// Same roles, TMEM slots and staging scheme as k_layer_tc.
// ------------------------------------------------------------------------------------------------
struct TcPostParams {
  const uint8_t* image[2];
  float* y;                 // [2][N][T], zeroed before the launch
  int N, T, tiles_per_utt;
};

struct TcPostBarriers {
  uint64_t w_ready;
  uint64_t z_full[2], z_free[2], a_ready[2], ds_ready[2], h_ready[2], d1_ready[2];
  uint32_t tmem_base;
  int mma_lock;
};

template <bool BF16, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1) k_post_tc(const __grid_constant__ CUtensorMap map_z, TcPostParams p) {
  using namespace ptx;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* smem = tc_smem;
  TcPostBarriers* bars = reinterpret_cast<TcPostBarriers*>(smem + TCP_SMEM_STAGE0 + 2 * TCP_STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int body = blockIdx.x & 1;
  const int cta_in_body = blockIdx.x >> 1, ctas_per_body = (gridDim.x + 1 - body) >> 1;
  const int tiles_body = p.N * p.tiles_per_utt;
  const int n_local = (tiles_body > cta_in_body) ? (tiles_body - cta_in_body + ctas_per_body - 1) / ctas_per_body : 0;

  if (warp == TC_MMA_WARP) {
    if (lane == 0) {
      mbar_init(&bars->w_ready, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&bars->z_full[s], 1);
        mbar_init(&bars->z_free[s], 256);
        mbar_init(&bars->a_ready[s], 256);
        mbar_init(&bars->ds_ready[s], 1);
        mbar_init(&bars->h_ready[s], 256);
        mbar_init(&bars->d1_ready[s], 1);
      }
      bars->mma_lock = 0;
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;
  const size_t body_off = (size_t)body * p.N * p.T;

  if (warp == TC_MMA_WARP || warp == TC_MMA_WARP + 1) {
    if (elect_one()) {
      const int s = warp - TC_MMA_WARP;
      if (s == 0) {
        const uint8_t* img = body ? p.image[1] : p.image[0];
        mbar_arrive_expect_tx(&bars->w_ready, TCP_IMAGE_BYTES);
        for (int off = 0; off < TCP_IMAGE_BYTES; off += 16384) {
          const int n = min(16384, TCP_IMAGE_BYTES - off);
          bulk_g2s(smem + off, img + off, n, &bars->w_ready);
        }
      }
      mbar_wait(&bars->w_ready, 0);
      const uint32_t wshi = smem_u32(smem + TCP_OFF_WSHI), wslo = smem_u32(smem + TCP_OFF_WSLO);
      const uint32_t w1hi = smem_u32(smem + TCP_OFF_W1HI), w1lo = smem_u32(smem + TCP_OFF_W1LO);
      constexpr uint32_t ID = idesc_f16(128, 128, BF16);
      const uint32_t tD = tmem + s * 256;
      const uint32_t tAhi = tD + 128, tAlo = tD + 192;
      const int tiles_s = (n_local + 1 - s) / 2;
      for (int j = 0; j < tiles_s; ++j) {
#pragma unroll
        for (int phase = 0; phase < 2; ++phase) {
          mbar_wait(phase == 0 ? &bars->a_ready[s] : &bars->h_ready[s], j & 1);
          tc_lock<SPLIT>(&bars->mma_lock);
          tc_fence_after_sync();
          const int ksteps = phase == 0 ? 4 : 8;          // K = 64 (z) / 128 (h)
          const uint32_t bhi = phase == 0 ? wshi : w1hi, blo = phase == 0 ? wslo : w1lo;
          uint32_t acc = 0;
          if (SPLIT) {
            for (int ks = 0; ks < ksteps; ++ks, acc = 1)
              mma_f16_ts(tD, tAlo + ks * 8, smem_desc_kmajor_noswizzle(bhi + ks * 2 * 2048, 2048, 128), ID, acc);
            for (int ks = 0; ks < ksteps; ++ks)
              mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(blo + ks * 2 * 2048, 2048, 128), ID, 1);
          }
          for (int ks = 0; ks < ksteps; ++ks, acc = 1)
            mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(bhi + ks * 2 * 2048, 2048, 128), ID, acc);
          mma_commit(phase == 0 ? &bars->ds_ready[s] : &bars->d1_ready[s]);
          tc_unlock<SPLIT>(&bars->mma_lock);
        }
      }
    }
    __syncwarp();
  } else if (warp >= TC_TMA_WARP) {
    if (elect_one()) {
      const int s = warp - TC_TMA_WARP;
      tma_prefetch_desc(&map_z);
      auto issue_z = [&](int j) {
        const int tile = cta_in_body + (s + 2 * j) * ctas_per_body;
        const int ub = body * p.N + tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * TC_TM;
        uint8_t* st = smem + TCP_SMEM_STAGE0 + s * TCP_STAGE_BYTES;
        mbar_arrive_expect_tx(&bars->z_full[s], 2 * TC_BOX_BYTES);
        tma_load_3d(st, &map_z, 0, t0, ub, &bars->z_full[s]);
        tma_load_3d(st + TC_BOX_BYTES, &map_z, 32, t0, ub, &bars->z_full[s]);
      };
      const int tiles_s = (n_local + 1 - s) / 2;
      if (tiles_s > 0) issue_z(0);
      for (int j = 0; j + 1 < tiles_s; ++j) {
        mbar_wait(&bars->z_free[s], j & 1);
        issue_z(j + 1);
      }
    }
    __syncwarp();
  } else {
    const int slot = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t tD = tmem + slot * 256 + lane_base;
    const uint32_t tAhi = tD + 128, tAlo = tD + 192;
    uint8_t* my_z = smem + TCP_SMEM_STAGE0 + slot * TCP_STAGE_BYTES + half * TC_BOX_BYTES + r * 128;
    const float* bs_s = reinterpret_cast<const float*>(smem + TCP_OFF_BS) + half * 64;
    const float* b1_s = reinterpret_cast<const float*>(smem + TCP_OFF_B1) + half * 64;
    const float* w2_s = reinterpret_cast<const float*>(smem + TCP_OFF_W2) + half * 64;
    const float* scal = reinterpret_cast<const float*>(smem + TCP_OFF_SCAL);
    bool weights_seen = false;
    float ss = 0.f, s1 = 0.f, b2 = 0.f;

    int j = 0;
    for (int local = slot; local < n_local; local += 2, ++j) {
      const int tile = cta_in_body + local * ctas_per_body;
      const int n = tile / p.tiles_per_utt, t = (tile % p.tiles_per_utt) * TC_TM + r;
      const bool row_ok = t < p.T;
      const uint32_t par = j & 1;

      // ---- A = z (my 32 channels: k = 32*half ..) -> fp16 hi/lo -> TMEM columns 16*half ..
      mbar_wait(&bars->z_full[slot], par);
      tc_prep<BF16, SPLIT>(my_z, r, tAhi + half * 16, tAlo + half * 16);
      mbar_arrive(&bars->z_free[slot]);
      tmem_wait_st();
      tc_fence_before_sync();
      mbar_arrive(&bars->a_ready[slot]);

      if (!weights_seen) {
        mbar_wait(&bars->w_ready, 0);
        ss = scal[0]; s1 = scal[1]; b2 = scal[2];
        weights_seen = true;
      }

      // ---- h = relu(skip) on my 64 of the 128 columns -> A' (k = 64*half .., columns 32*half ..);
      //      A' lives in the slot's A columns, disjoint from D
      mbar_wait(&bars->ds_ready[slot], par);
      tc_fence_after_sync();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t dr[16];
        tmem_ld16(tD + half * 64 + c * 16, dr);
        tmem_wait_ld();
        float v0[8], v1[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v0[e] = fmaxf(fmaf(__uint_as_float(dr[e]), ss, bs_s[c * 16 + e]), 0.f);
          v1[e] = fmaxf(fmaf(__uint_as_float(dr[8 + e]), ss, bs_s[c * 16 + 8 + e]), 0.f);
        }
        uint32_t hi[8], lo[8];
        split8<BF16, SPLIT>(v0, hi, lo);
        split8<BF16, SPLIT>(v1, hi + 4, lo + 4);
        tmem_st8(tAhi + half * 32 + c * 8, hi);
        if (SPLIT) tmem_st8(tAlo + half * 32 + c * 8, lo);
      }
      tmem_wait_st();
      tc_fence_before_sync();
      mbar_arrive(&bars->h_ready[slot]);

      // ---- y = relu(D' + b1) . w2 (+ b2 once per row)
      mbar_wait(&bars->d1_ready[slot], par);
      tc_fence_after_sync();
      float acc = half == 0 ? b2 : 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t dr[16];
        tmem_ld16(tD + half * 64 + c * 16, dr);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e)
          acc = fmaf(fmaxf(fmaf(__uint_as_float(dr[e]), s1, b1_s[c * 16 + e]), 0.f), w2_s[c * 16 + e], acc);
      }
      tc_fence_before_sync();
      if (row_ok) atomicAdd(p.y + body_off + (size_t)n * p.T + t, acc);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == TC_MMA_WARP) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// Conditioning rows of one flow on the tensor cores: for every (body, layer) z of the flow
//   out_z[m][0:128] = (cproj[m][0:Cc] . Wgc_z + bias_z) * colscale          m = (utterance, mel frame)
// (reference modules.py:216-228 at mel rate; 1.3 % of the path's FLOPs, 4 % of its time on CUDA cores).
// A CTA owns one 128-row tile of cproj -- split once into fp16 hi/lo and parked in TMEM -- and walks a
// strided subset of the z entries, streaming their weight images through a 2-deep shared-memory ring;
// accumulators are double-buffered in TMEM so the epilogue of entry b overlaps the MMAs of b+1. The
// output tile (128 x 128 fp32) goes through a swizzled staging tile and leaves as full 512-byte rows.
// grid = (ceil(M/128), zsplit); 320 threads: 8 worker warps, 1 MMA warp, 1 producer warp.
// ------------------------------------------------------------------------------------------------
struct TcCondParams {
  const float* cproj;       // [M][Cc]
  const uint8_t* images;    // entry z at images + z * image_bytes
  float* out;               // entry z at out + z * out_stride
  size_t out_stride;        // floats (= M * 128)
  int M, Cc, Z, image_bytes;
};

constexpr int TCC_THREADS = 320;
// shared memory: [2 weight images][cproj tile, rows of Cc*4+16 B][output tile 128 x 512 B, 16-byte chunks
// XOR-swizzled by row][barriers]; sizes depend on Cc (tcc_smem_bytes)

struct TcCondBarriers {
  uint64_t a_full, a_ready, w_full[2], w_free[2], d_ready[2], d_free[2];
  uint32_t tmem_base;
};

template <bool BF16, bool SPLIT>
__global__ void __launch_bounds__(TCC_THREADS, 1) k_cbias_tc(TcCondParams p) {
  using namespace ptx;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* smem = tc_smem;
  const int a_pitch = p.Cc * 4 + 16;
  const int TCC_W_BYTES = p.image_bytes;
  const int TCC_SMEM_W0 = 0;
  const int TCC_SMEM_A = 2 * TCC_W_BYTES;
  const int TCC_SMEM_OUT = TCC_SMEM_A + (128 * a_pitch + 1023) / 1024 * 1024;
  const int TCC_SMEM_BARS = TCC_SMEM_OUT + 128 * 512;
  TcCondBarriers* bars = reinterpret_cast<TcCondBarriers*>(smem + TCC_SMEM_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_TM;
  const int zsplit = gridDim.y, z0 = blockIdx.y;
  const int nb = (p.Z > z0) ? (p.Z - z0 + zsplit - 1) / zsplit : 0;      // entries z0, z0 + zsplit, ...
  const int ksteps = p.Cc / 16, half_bytes = (p.Cc / 8) * 2048;

  if (warp == 8) {
    if (lane == 0) {
      mbar_init(&bars->a_full, 1);
      mbar_init(&bars->a_ready, 256);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->w_full[i], 1);
        mbar_init(&bars->w_free[i], 1);
        mbar_init(&bars->d_ready[i], 1);
        mbar_init(&bars->d_free[i], 256);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;      // D0 [0,128) | D1 [128,256) | Ahi [256,320) | Alo [320,384)

  if (warp == 8) {
    // ---------------- MMA issuer
    if (elect_one()) {
      constexpr uint32_t ID = idesc_f16(128, 128, BF16);
      mbar_wait(&bars->a_ready, 0);
      for (int b = 0; b < nb; ++b) {
        const int buf = b & 1;
        const uint32_t use = (uint32_t)(b >> 1) & 1;
        mbar_wait(&bars->w_full[buf], use);
        mbar_wait(&bars->d_free[buf], use ^ 1);        // (first use of each buffer passes immediately)
        tc_fence_after_sync();
        const uint32_t whi = smem_u32(smem + TCC_SMEM_W0 + buf * TCC_W_BYTES), wlo = whi + half_bytes;
        const uint32_t tD = tmem + buf * 128, tAhi = tmem + 256, tAlo = tmem + 320;
        uint32_t acc = 0;
        if (SPLIT) {
          for (int ks = 0; ks < ksteps; ++ks, acc = 1)
            mma_f16_ts(tD, tAlo + ks * 8, smem_desc_kmajor_noswizzle(whi + ks * 2 * 2048, 2048, 128), ID, acc);
          for (int ks = 0; ks < ksteps; ++ks)
            mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(wlo + ks * 2 * 2048, 2048, 128), ID, 1);
        }
        for (int ks = 0; ks < ksteps; ++ks, acc = 1)
          mma_f16_ts(tD, tAhi + ks * 8, smem_desc_kmajor_noswizzle(whi + ks * 2 * 2048, 2048, 128), ID, acc);
        mma_commit(&bars->d_ready[buf]);
        mma_commit(&bars->w_free[buf]);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---------------- producer: the cproj tile once, then the weight images
    if (elect_one()) {
      const int rows = min(TC_TM, p.M - m0);
      mbar_arrive_expect_tx(&bars->a_full, (uint32_t)rows * p.Cc * 4);
      for (int r = 0; r < rows; ++r)
        bulk_g2s(smem + TCC_SMEM_A + r * a_pitch, p.cproj + (size_t)(m0 + r) * p.Cc, p.Cc * 4, &bars->a_full);
      for (int b = 0; b < nb; ++b) {
        const int buf = b & 1;
        const uint32_t use = (uint32_t)(b >> 1) & 1;
        if (b >= 2) {
          mbar_wait(&bars->w_free[buf], use ^ 1);      // the MMAs of entry b-2 have read the buffer ...
          mbar_wait(&bars->d_free[buf], use ^ 1);      // ... and the workers its colmul / coladd vectors
        }
        const int z = z0 + b * zsplit;
        mbar_arrive_expect_tx(&bars->w_full[buf], (uint32_t)p.image_bytes);
        bulk_g2s(smem + TCC_SMEM_W0 + buf * TCC_W_BYTES, p.images + (size_t)z * p.image_bytes, p.image_bytes, &bars->w_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ---------------- workers: thread = (row, 64 of the 128 output columns)
    const int half = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const bool row_ok = m0 + r < p.M;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    // A: k range of this half: [0, kh) / [kh, Cc) with kh a multiple of 16 so that tcgen05.st widths fit
    mbar_wait(&bars->a_full, 0);
    {
      const float4* arow = reinterpret_cast<const float4*>(smem + TCC_SMEM_A + r * a_pitch);
      const int k_per_half = p.Cc / 2;               // Cc % 16 == 0 -> multiple of 8
      const int kb = half * k_per_half;
      for (int k = kb; k < kb + k_per_half; k += 8) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a;
        if (row_ok) { a = arow[k / 4]; b4 = arow[k / 4 + 1]; }
        const float v[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
        uint32_t hi[4], lo[4];
        split8<BF16, SPLIT>(v, hi, lo);
        tmem_st4(tmem + 256 + lane_base + k / 2, hi);
        if (SPLIT) tmem_st4(tmem + 320 + lane_base + k / 2, lo);
      }
    }
    tmem_wait_st();
    tc_fence_before_sync();
    mbar_arrive(&bars->a_ready);

    uint8_t* stage_row = smem + TCC_SMEM_OUT + r * 512;
    for (int b = 0; b < nb; ++b) {
      const int buf = b & 1;
      const uint32_t use = (uint32_t)(b >> 1) & 1;
      const int z = z0 + b * zsplit;
      const float* vec = reinterpret_cast<const float*>(smem + TCC_SMEM_W0 + buf * TCC_W_BYTES + 2 * half_bytes);
      mbar_wait(&bars->d_ready[buf], use);
      tc_fence_after_sync();
      const uint32_t tD = tmem + buf * 128 + lane_base + half * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t dr[16];
        tmem_ld16(tD + c * 16, dr);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = half * 64 + c * 16 + q * 4;
          const float4 mul = *reinterpret_cast<const float4*>(vec + col);
          const float4 add = *reinterpret_cast<const float4*>(vec + 128 + col);
          float4 o;
          o.x = fmaf(__uint_as_float(dr[4 * q + 0]), mul.x, add.x);
          o.y = fmaf(__uint_as_float(dr[4 * q + 1]), mul.y, add.y);
          o.z = fmaf(__uint_as_float(dr[4 * q + 2]), mul.z, add.z);
          o.w = fmaf(__uint_as_float(dr[4 * q + 3]), mul.w, add.w);
          const int chunk = col >> 2;                       // 16-byte chunk 0..31 of the row
          *reinterpret_cast<float4*>(stage_row + (((chunk & ~7) | ((chunk ^ r) & 7)) << 4)) = o;
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&bars->d_free[buf]);                      // accumulator and weight vectors of entry b consumed
      named_bar_sync(1, 256);                               // the whole tile is staged
      float* out_tile = p.out + (size_t)z * p.out_stride + (size_t)m0 * 128;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {                        // a warp instruction writes one full 512-byte row
        const int row = warp * 16 + i;
        const float4 v = *reinterpret_cast<const float4*>(smem + TCC_SMEM_OUT + row * 512 + (((lane & ~7) | ((lane ^ row) & 7)) << 4));
        if (m0 + row < p.M) *reinterpret_cast<float4*>(out_tile + (size_t)row * 128 + lane * 4) = v;
      }
      named_bar_sync(1, 256);                               // staging tile free for the next entry
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

}  // namespace pwv
