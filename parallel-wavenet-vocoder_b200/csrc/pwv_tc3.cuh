// pwv_tc3.cuh -- k_wide_h: the gated dilated layer on tcgen05 for channel counts beyond 64 (BASELINE config c5:
// residual = dilation channels 128 / 256; in fact any multiple of 64), reference modules.py:185-259.
//
// At 64 channels the layer's weights (80 KB) stay resident in shared memory and a whole tile lives in tensor memory
// (pwv_tc2.cuh). At C = 128 the fp16 hi/lo weight images are 320 KB, at C = 256 1.25 MB, and one 128-row tile's
// operands plus accumulators (K = 2C, N = 2C) exceed the 512 TMEM columns -- so here the layer is two streamed-K GEMMs
// with fused epilogues, the classic warp-specialised tcgen05 GEMM shape, on the same 16-bit plane layout:
//
//   pass G (EPI_GATE):  D[128 x 128] = [x[t-d] | x[t]](128 x 2R) . W1[:, block]      block = 64 gate channels: f | g columns
//                       z = tanh(f) * sigmoid(g)  -> z planes (HBM)
//   pass D (EPI_DENSE): D[128 x  64] = z(128 x D) . W2[:, block]                      block = 64 output channels
//                       out = x[t] + D + b_dense  -> activation planes (HBM)
//
// A work item = (128-row tile, 64-channel output block); a persistent CTA walks items. Per item the K dimension streams
// in 64-channel blocks: the activation block arrives as TMA boxes (hi / lo planes, 128B swizzle) in a 2-deep landing
// ring, four copy warps move it verbatim into a 2-deep ring of TMEM A columns (the tensor core then reads A from TMEM,
// B only from shared memory); the matching 64-row slice of the weight block (fp16 hi / lo, K-major core-matrix layout,
// packed at finalize time in exactly this stage order) streams through a 3-deep bulk-copy ring; one thread issues the
// 3 x 4 MMAs of the stage (f16x3: lo.hi + hi.lo + hi.hi) into one of two accumulators, so the epilogue of item i
// overlaps the MMAs of item i+1. z costs one extra round trip through HBM (8 B per element and layer) -- irrelevant
// next to 5 C^2 x 3 MACs per row at these widths.
// Warps: 0-7 epilogue (thread = row x 32 of the block's 64 channels), 8-11 operand copy (thread = row), 12 MMA issuer,
// 13 loader (TMA boxes + weight stages), 14 storer (residual tile in, staged output tile out).
#pragma once
#include "pwv_tc2.cuh"

namespace pwv {

constexpr int TW_SA = 2, TW_SB = 3;
constexpr int TW_THREADS = 15 * 32;
constexpr int TW_EPI_GATE = 0, TW_EPI_DENSE = 1;
constexpr int TW_A_STAGE = 2 * TH_BOX_BYTES;                 // hi + lo box of one 64-channel block (bf16: first half used)
constexpr int TW_B_STAGE_MAX = 2 * 64 * 128 * 2;             // hi + lo slice: 64 K-rows x 128 columns x 2 B
constexpr int TW_SMEM_A0 = 0;
constexpr int TW_SMEM_B0 = TW_SMEM_A0 + TW_SA * TW_A_STAGE;  // 64 KB
constexpr int TW_SMEM_OUT = TW_SMEM_B0 + TW_SB * TW_B_STAGE_MAX;   // + 96 KB
constexpr int TW_SMEM_BARS = TW_SMEM_OUT + 2 * TH_BOX_BYTES;       // + 32 KB
constexpr int TW_SMEM_BYTES = TW_SMEM_BARS + 512;

struct TwParams {
  const uint8_t* wimg[2];   // per body: weight stream of this pass: item block nb, K block kb at ((nb * KB + kb) * stage_bytes)
  const float* vec[2];      // per body: [sf, sg, s2, 0 | dense bias (C_out floats)]
  const float* cbias[2];    // EPI_GATE: per body [N][t_mel][2 * C_out], pre-scaled (filter half by KF, gate half by KG)
  int N, T, t_mel, hop, dilation, tiles_per_utt;
  int KB;                   // K blocks of 64 channels (gate: 2 R / 64, dense: D / 64)
  int KB_tap;               // gate: blocks kb < KB_tap are x[t-d] (rows shifted by the dilation), the others x[t]; dense: = KB
  int NB;                   // output blocks of 64 channels
  int C_out;                // gate: D (dilation channels), dense: R
};

struct TwBarriers {
  uint64_t a_full[TW_SA], a_free[TW_SA], at_full[2], at_free[2], b_full[TW_SB], b_free[TW_SB], d_full[2], d_free[2];
  uint64_t x_full, out_ready, out_free;
  uint32_t tmem_base;
};

// host side: one stage = [hi | lo] K-major core-matrix images of W[k0 .. k0+63][cols], cols = the block's columns
inline void tw_pack_stage(uint8_t* dst, const float* w, int ld, int k0, const int* cols, int ncols, float scale, bool bf16, bool split) {
  const size_t half = (size_t)64 * ncols * 2;
  for (int k = 0; k < 64; ++k)
    for (int n = 0; n < ncols; ++n) {
      const float v = w[(size_t)(k0 + k) * ld + cols[n]] * scale;
      const uint16_t h = tc_to16(v, bf16);
      const uint16_t l = split ? tc_to16(v - tc_from16(h, bf16), bf16) : 0;
      const size_t off = (size_t)(k / 8) * (ncols * 16) + (size_t)n * 16 + (k % 8) * 2;
      std::memcpy(dst + off, &h, 2);
      std::memcpy(dst + half + off, &l, 2);
    }
}

// Weight streams of a model with C = residual = dilation channels (multiple of 64), per (flow, body, layer) in that order:
//   gate  stream: for every output block nb (64 gate channels: columns [f block | g block] of the packed [2C][2C] matrix) the
//                 2C/64 K-stages in K order (x[t-d] channels, then x[t] channels), each [hi | lo] = 2 x 64 x 128 x 2 B
//   dense stream: for every output block nb (64 residual channels) the C/64 K-stages, each 2 x 64 x 64 x 2 B
//   vec         : [KF / s1, KG / s1, 1 / s2, 0 | dense bias (C floats)]
struct TwModel {
  uint8_t* d_gate = nullptr;
  uint8_t* d_dense = nullptr;
  float* d_vec = nullptr;
  size_t gate_bytes = 0, dense_bytes = 0;     // per (flow, body, layer)
  int vec_floats = 0;
  int C = 0;
};

inline void tw_model_free(TwModel& t) {
  if (t.d_gate) cudaFree(t.d_gate);
  if (t.d_dense) cudaFree(t.d_dense);
  if (t.d_vec) cudaFree(t.d_vec);
  t = TwModel();
}

// precision: 1 = f16x3, 2 = bf16; layers in (flow, body, layer) order, sources as for tc_model_build (packed fp32 arena)
inline const char* tw_model_build(TwModel& t, int precision, int C, const std::vector<TcLayerSrc>& layers) {
  if (C % 64 != 0 || C < 64) return "wide tensor-core kernels need a channel count that is a multiple of 64";
  const bool bf16 = precision == 2, split = precision == 1;
  const int NB = C / 64, KBg = 2 * C / 64, KBd = C / 64;
  const size_t gstage = 2 * 64 * 128 * 2, dstage = 2 * 64 * 64 * 2;
  tw_model_free(t);
  t.C = C;
  t.gate_bytes = (size_t)NB * KBg * gstage;
  t.dense_bytes = (size_t)NB * KBd * dstage;
  t.vec_floats = 4 + C;
  std::vector<uint8_t> hg(layers.size() * t.gate_bytes, 0), hd(layers.size() * t.dense_bytes, 0);
  std::vector<float> hv(layers.size() * (size_t)t.vec_floats, 0.f);
  std::vector<int> cols(128);
  for (size_t i = 0; i < layers.size(); ++i) {
    const float s1 = bf16 ? 1.f : tc_pow2_scale(layers[i].wfg, (size_t)2 * C * 2 * C);
    const float s2 = bf16 ? 1.f : tc_pow2_scale(layers[i].wd, (size_t)C * C);
    for (int nb = 0; nb < NB; ++nb) {
      for (int c = 0; c < 64; ++c) { cols[c] = nb * 64 + c; cols[64 + c] = C + nb * 64 + c; }
      for (int kb = 0; kb < KBg; ++kb)
        tw_pack_stage(hg.data() + i * t.gate_bytes + ((size_t)nb * KBg + kb) * gstage, layers[i].wfg, 2 * C, kb * 64, cols.data(), 128, s1, bf16, split);
      for (int kb = 0; kb < KBd; ++kb)
        tw_pack_stage(hd.data() + i * t.dense_bytes + ((size_t)nb * KBd + kb) * dstage, layers[i].wd, C, kb * 64, cols.data(), 64, s2, bf16, split);
    }
    float* v = hv.data() + i * (size_t)t.vec_floats;
    v[0] = TC_KF / s1; v[1] = TC_KG / s1; v[2] = 1.f / s2; v[3] = 0.f;
    std::memcpy(v + 4, layers[i].bd, C * sizeof(float));
  }
  if (cudaMalloc(&t.d_gate, hg.size()) != cudaSuccess || cudaMalloc(&t.d_dense, hd.size()) != cudaSuccess ||
      cudaMalloc(&t.d_vec, hv.size() * sizeof(float)) != cudaSuccess)
    return "cudaMalloc of the wide tensor-core weight streams failed";
  if (cudaMemcpy(t.d_gate, hg.data(), hg.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(t.d_dense, hd.data(), hd.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(t.d_vec, hv.data(), hv.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
    return "upload of the wide tensor-core weight streams failed";
  return nullptr;
}

template <bool BF16, int EPI>
__global__ void __launch_bounds__(TW_THREADS, 1)
k_wide_h(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out, TwParams p) {
  using namespace ptx;
  constexpr bool SPLIT = !BF16;
  constexpr int P = BF16 ? 1 : 2;
  constexpr int NBC = EPI == TW_EPI_GATE ? 128 : 64;          // accumulator columns of an item
  constexpr int B_HALF = 64 * NBC * 2, B_STAGE = 2 * B_HALF;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* smem = tc_smem;
  TwBarriers* bars = reinterpret_cast<TwBarriers*>(smem + TW_SMEM_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_body = p.N * p.tiles_per_utt;
  const int items = 2 * tiles_body * p.NB;                    // (body, tile, block), block fastest
  const int n_items = (items > (int)blockIdx.x) ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto decode = [&](int it, int& body, int& n, int& t0, int& nb) {
    const int g = blockIdx.x + it * gridDim.x;
    nb = g % p.NB;
    const int tile = (g / p.NB) % tiles_body;
    body = g / (p.NB * tiles_body);
    n = tile / p.tiles_per_utt;
    t0 = (tile % p.tiles_per_utt) * TC_TM;
  };

  if (warp == 12) {
    if (lane == 0) {
      for (int i = 0; i < TW_SA; ++i) { mbar_init(&bars->a_full[i], 1); mbar_init(&bars->a_free[i], 128); }
      for (int i = 0; i < 2; ++i) { mbar_init(&bars->at_full[i], 128); mbar_init(&bars->at_free[i], 1); mbar_init(&bars->d_full[i], 1); mbar_init(&bars->d_free[i], 256); }
      for (int i = 0; i < TW_SB; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_free[i], 1); }
      mbar_init(&bars->x_full, 1);
      mbar_init(&bars->out_ready, 256);
      mbar_init(&bars->out_free, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;     // D0 [0,128) | D1 [128,256) | A ring: stage s at 256 + 64 s: hi 32 cols | lo 32 cols

  if (warp == 13) {
    // ======================= loader: activation boxes and weight stages, K block by K block =======================
    if (elect_one()) {
      tma_prefetch_desc(&map_a);
      uint32_t step = 0;
      for (int it = 0; it < n_items; ++it) {
        int body, n, t0, nb;
        decode(it, body, n, t0, nb);
        const uint8_t* wsrc = (body ? p.wimg[1] : p.wimg[0]) + (size_t)nb * p.KB * B_STAGE;
        for (int kb = 0; kb < p.KB; ++kb, ++step) {
          const int sa = step % TW_SA, sb = step % TW_SB;
          const uint32_t ua = step / TW_SA, ub = step / TW_SB;
          if (ua > 0) mbar_wait(&bars->a_free[sa], (ua - 1) & 1);
          const bool tap = kb < p.KB_tap && p.KB_tap < p.KB;     // gate pass: the x[t-d] half of K
          const int cblk = kb < p.KB_tap ? kb : kb - p.KB_tap;
          mbar_arrive_expect_tx(&bars->a_full[sa], P * TH_BOX_BYTES);
#pragma unroll
          for (int q = 0; q < P; ++q)
            tma_load_3d(smem + TW_SMEM_A0 + sa * TW_A_STAGE + q * TH_BOX_BYTES, &map_a, cblk * 64, tap ? t0 - p.dilation : t0,
                        q * 2 * p.N + body * p.N + n, &bars->a_full[sa]);
          if (ub > 0) mbar_wait(&bars->b_free[sb], (ub - 1) & 1);
          mbar_arrive_expect_tx(&bars->b_full[sb], B_STAGE);
          for (int off = 0; off < B_STAGE; off += 16384)
            bulk_g2s(smem + TW_SMEM_B0 + sb * TW_B_STAGE_MAX + off, wsrc + (size_t)kb * B_STAGE + off, 16384, &bars->b_full[sb]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 14) {
    // ======================= storer: residual tile in (dense pass), staged output tile out =======================
    if (elect_one()) {
      tma_prefetch_desc(&map_out);
      uint8_t* ob = smem + TW_SMEM_OUT;
      for (int it = 0; it < n_items; ++it) {
        int body, n, t0, nb;
        decode(it, body, n, t0, nb);
        const int ub = body * p.N + n;
        if (EPI == TW_EPI_DENSE) {         // (the buffer is free: this thread waited for the previous store's reads)
          mbar_arrive_expect_tx(&bars->x_full, P * TH_BOX_BYTES);
#pragma unroll
          for (int q = 0; q < P; ++q) tma_load_3d(ob + q * TH_BOX_BYTES, &map_x, nb * 64, t0, q * 2 * p.N + ub, &bars->x_full);
        }
        mbar_wait(&bars->out_ready, it & 1);
#pragma unroll
        for (int q = 0; q < P; ++q) tma_store_3d(&map_out, nb * 64, t0, q * 2 * p.N + ub, ob + q * TH_BOX_BYTES);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(&bars->out_free);
      }
      bulk_wait0();
    }
    __syncwarp();
  } else if (warp == 12) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      constexpr uint32_t ID = idesc_f16(128, NBC, BF16);
      uint32_t step = 0;
      for (int it = 0; it < n_items; ++it) {
        const int db = it & 1;
        const uint32_t ud = it >> 1;
        if (ud > 0) mbar_wait(&bars->d_free[db], (ud - 1) & 1);
        const uint32_t tD = tmem + db * 128;
        for (int kb = 0; kb < p.KB; ++kb, ++step) {
          const int s = step & 1, sb = step % TW_SB;
          mbar_wait(&bars->at_full[s], (step >> 1) & 1);
          mbar_wait(&bars->b_full[sb], (step / TW_SB) & 1);
          tc_fence_after_sync();
          const uint32_t tAhi = tmem + 256 + s * 64, tAlo = tAhi + 32;
          const uint64_t dhi = smem_desc_kmajor_noswizzle(smem_u32(smem + TW_SMEM_B0 + sb * TW_B_STAGE_MAX), NBC * 16, 128);
          const uint64_t dlo = smem_desc_kmajor_noswizzle(smem_u32(smem + TW_SMEM_B0 + sb * TW_B_STAGE_MAX + B_HALF), NBC * 16, 128);
          uint32_t acc = kb > 0 ? 1u : 0u;
          if (SPLIT) {
#pragma unroll 1
            for (int ks = 0; ks < 4; ++ks, acc = 1) mma_f16_ts(tD, tAlo + ks * 8, dhi + (uint64_t)(ks * NBC * 2), ID, acc);
#pragma unroll 1
            for (int ks = 0; ks < 4; ++ks) mma_f16_ts(tD, tAhi + ks * 8, dlo + (uint64_t)(ks * NBC * 2), ID, 1);
          }
#pragma unroll 1
          for (int ks = 0; ks < 4; ++ks, acc = 1) mma_f16_ts(tD, tAhi + ks * 8, dhi + (uint64_t)(ks * NBC * 2), ID, acc);
          mma_commit(&bars->at_free[s]);
          mma_commit(&bars->b_free[sb]);
        }
        mma_commit(&bars->d_full[db]);
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ======================= operand copy: landing boxes -> TMEM A ring (thread = row) =======================
    const int quarter = warp & 3, r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint32_t step = 0;
    for (int it = 0; it < n_items; ++it) {
      for (int kb = 0; kb < p.KB; ++kb, ++step) {
        const int s = step & 1, sa = step % TW_SA;
        mbar_wait(&bars->a_full[sa], (step / TW_SA) & 1);
        if (step >= 2) mbar_wait(&bars->at_free[s], ((step >> 1) - 1) & 1);
        tc_fence_after_sync();
        const uint8_t* box = smem + TW_SMEM_A0 + sa * TW_A_STAGE;
#pragma unroll
        for (int q = 0; q < P; ++q) {
          uint32_t v[32];
          const uint8_t* row = box + q * TH_BOX_BYTES + r * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(row + (((c ^ r) & 7) << 4));
            v[4 * c] = u.x; v[4 * c + 1] = u.y; v[4 * c + 2] = u.z; v[4 * c + 3] = u.w;
          }
          tmem_st32(tmem + 256 + s * 64 + q * 32 + lane_base, v);
        }
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->at_full[s]);
        mbar_arrive(&bars->a_free[sa]);
      }
    }
  } else {
    // ======================= epilogue: thread = (row, 32 of the block's 64 channels) =======================
    const int half = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint8_t* ob = smem + TW_SMEM_OUT;
    for (int it = 0; it < n_items; ++it) {
      int body, n, t0, nb;
      decode(it, body, n, t0, nb);
      const int db = it & 1;
      const uint32_t tD = tmem + db * 128 + lane_base;
      const float* vec = body ? p.vec[1] : p.vec[0];
      mbar_wait(&bars->d_full[db], (it >> 1) & 1);
      tc_fence_after_sync();
      uint32_t oh[16], ol[16];
      if (EPI == TW_EPI_GATE) {
        const float sf = vec[0], sg = vec[1];
        const int t = t0 + r;
        const int frame = (min(t, p.T - 1) + p.hop / 2) / p.hop;
        const float* cbrow = (body ? p.cbias[1] : p.cbias[0]) + ((size_t)n * p.t_mel + frame) * 2 * p.C_out + nb * 64 + half * 32;
        const float4* cbf = reinterpret_cast<const float4*>(cbrow);
        const float4* cbg = reinterpret_cast<const float4*>(cbrow + p.C_out);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t fr[16], gr[16];
          tmem_ld16(tD + half * 32 + c * 16, fr);
          tmem_ld16(tD + 64 + half * 32 + c * 16, gr);
          tmem_wait_ld();
          if (c == 1) {                    // the accumulator is in registers: the MMAs of item it+2 may overwrite it
            tc_fence_before_sync();
            mbar_arrive(&bars->d_free[db]);
          }
          float z[16];
          tc_gate<BF16, false, 16>(fr, gr, cbf + c * 4, cbg + c * 4, sf, sg, z);
          float v0[8], v1[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { v0[e] = z[e]; v1[e] = z[8 + e]; }
          split8<BF16, SPLIT>(v0, oh + c * 8, ol + c * 8);
          split8<BF16, SPLIT>(v1, oh + c * 8 + 4, ol + c * 8 + 4);
        }
        if (it > 0) mbar_wait(&bars->out_free, (it - 1) & 1);      // the previous item's tile has left the staging boxes
      } else {
        const float s2 = vec[2];
        const float* bd = vec + 4 + nb * 64 + half * 32;
        uint32_t dr[32];
        tmem_ld32(tD + half * 32, dr);
        tmem_wait_ld();
        tc_fence_before_sync();
        mbar_arrive(&bars->d_free[db]);
        mbar_wait(&bars->x_full, it & 1);                           // x[t] of the block (hi / lo boxes) sits in the staging boxes
        uint32_t xh[16], xl[16];
        th_ld_row64(ob, r, half * 4, xh);
        if (SPLIT) th_ld_row64(ob + TH_BOX_BYTES, r, half * 4, xl);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 h2 = unpack16<BF16>(xh[q * 4 + e]);
            float x0 = h2.x, x1 = h2.y;
            if (SPLIT) {
              const float2 l2 = unpack16<BF16>(xl[q * 4 + e]);
              x0 += l2.x;
              x1 += l2.y;
            }
            o[2 * e] = x0 + fmaf(__uint_as_float(dr[q * 8 + 2 * e]), s2, bd[q * 8 + 2 * e]);
            o[2 * e + 1] = x1 + fmaf(__uint_as_float(dr[q * 8 + 2 * e + 1]), s2, bd[q * 8 + 2 * e + 1]);
          }
          split8<BF16, SPLIT>(o, oh + q * 4, ol + q * 4);
        }
      }
      th_st_row64(ob, r, half * 4, oh);                             // (dense: in place over my own x chunks)
      if (SPLIT) th_st_row64(ob + TH_BOX_BYTES, r, half * 4, ol);
      fence_proxy_async_smem();
      mbar_arrive(&bars->out_ready);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem, 512);
}

// flow front for the plane layout at any channel count (multiple of 8): see k_front_h
struct FrontWParams {
  const float* x_prev;
  const float* scale;
  const float* shift;
  float* x_new;
  const float* wc[2];     // per body: [2][C]
  uint16_t* act;          // [planes][2N][T][C]
  int N, T, C;
};

template <bool BF16>
__global__ void __launch_bounds__(256) k_front_w(FrontWParams p) {
  const int C = p.C;
  const size_t total = (size_t)p.N * p.T * (C / 8);
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = (int)(idx % (C / 8));
  const size_t row = idx / (C / 8);
  const int t = (int)(row % p.T);
  float xc = p.x_prev[row];
  float xp = (t > 0) ? p.x_prev[row - 1] : 0.f;
  if (p.scale) {
    xc = xc * p.scale[row] + p.shift[row];
    if (t > 0) xp = xp * p.scale[row - 1] + p.shift[row - 1];
  }
  if (cg == 0) p.x_new[row] = xc;
  const size_t plane = (size_t)2 * p.N * p.T * C;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = p.wc[b][cg * 8 + e] * xp + p.wc[b][C + cg * 8 + e] * xc;
    uint32_t hi[4], lo[4];
    split8<BF16, !BF16>(v, hi, lo);
    const size_t off = ((size_t)b * p.N * p.T + row) * C + cg * 8;
    *reinterpret_cast<uint4*>(p.act + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (!BF16) *reinterpret_cast<uint4*>(p.act + plane + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// planes [.][rows][C] -> fp32 rows (post-net input / debug tap); pairs = rows * C / 2
template <bool BF16>
__global__ void __launch_bounds__(256) k_planes_to_f32_n(const uint16_t* __restrict__ act, float* __restrict__ out, size_t pairs, size_t plane_elems) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= pairs) return;
  float2 v = unpack16<BF16>(reinterpret_cast<const uint32_t*>(act)[idx]);
  if (!BF16) {
    const float2 l = unpack16<BF16>(reinterpret_cast<const uint32_t*>(act + plane_elems)[idx]);
    v.x += l.x;
    v.y += l.y;
  }
  reinterpret_cast<float2*>(out)[idx] = v;
}

}  // namespace pwv
