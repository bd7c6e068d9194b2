// pwv_tc2.cuh -- k_layer_h: the gated dilated layer (reference modules.py:185-259) on tcgen05 with the
// activations kept in HBM as 16-bit PLANES instead of fp32 rows (round 2).
//
// Why (round-1 evidence, profiles/r1_tc_trace_v16_flow_kernel_layer2.txt): k_layer_tc / k_flow_tc keep two tiles
// per SM in flight, each a serial chain of ~10.5k cycles
//     boxes landed -> fp32->fp16 hi/lo conversion (1.7k) -> GEMM1 (2.3k) -> gate (2.4k) -> GEMM2 (0.9k)
//     -> residual (1.6k) -> copy-out (0.8k) -> refill of the x[t] boxes (0.7k)
// of which only GEMM1 -> gate -> GEMM2 -> accumulator read-out needs the tile slot's tensor memory. No unit was
// saturated (tensor pipe 36 %); the kernel was bound by the length of that chain. Here the chain a TMEM slot sees is
//     operand copy (0.4k) -> GEMM1 -> gate -> GEMM2 -> read-out (0.3k)
// and everything else happens beside it:
//   * Layout. A layer's output is stored the way the next layer's tensor-core operand wants it: f16x3 mode keeps
//     two fp16 planes hi = fp16(v), lo = fp16(v - hi) (4 bytes per element, the same HBM bytes as fp32, 22 significant
//     bits; the round-1 kernels rounded every GEMM operand to exactly these 22 bits already, now the residual stream
//     carries them too: emulated drift 6e-7 on the default graph), bf16 mode ONE bf16 plane (half the bytes:
//     BASELINE config c3, SURVEY 8d's 30,732 B/sample model). act = [planes][2N utterance-bodies][T][64] 16-bit,
//     one time step of one plane = one 128-byte row; a TMA box = 128 rows x 64 channels = 16 KB, 128B-swizzled.
//     The conversion phase of every consumer (x[t-d] AND x[t]: twice per element per layer) becomes ONE split in
//     the producer's epilogue.
//   * Operand copy. The boxes are copied into the slot's TMEM A columns verbatim (4 x ld.shared.v4 + 1 tcgen05.st
//     per box and thread), after which the landing boxes are free: the x[t-d] boxes are refilled at once, the x[t]
//     boxes become the staging of the PREVIOUS tile's output.
//   * Residual from TMEM. x[t] = hi + lo is read back from the A columns together with D2 (the gate output z
//     overwrites only the x[t-d] columns), so the x[t] boxes are not needed after the copy.
//   * Software pipeline per slot: ... gate(j) -> [D2(j), x(j) -> registers] -> operand copy(j+1) -> (GEMM1(j+1)
//     runs) -> residual(j) + hi/lo split -> staged in the x[t] boxes -> TMA store by the producer warp -> x[t] boxes
//     refilled for j+2. The residual arithmetic and the store of tile j overlap GEMM1 of tile j+1.
//   * Registers: 640 threads launch with 96 registers; the helper warpgroup (MMA issuers, producers) drops to
//     TH_HELPER_REGS and the four worker warpgroups rise to TH_WORKER_REGS (setmaxnreg moves registers inside the CTA's
//     launch allocation of 640 x 96; ptxas allocates each role's branch with its own budget).
// Weights, TMEM slot layout (D1 128 | Ahi 64 | Alo 64), arithmetic (f16x3 / bf16, gate formulas, epilogue scales),
// tile-to-CTA assignment, programmatic dependent launch and the per-tile flag handshake between consecutive layers
// are those of k_layer_tc (pwv_tc.cuh).
#pragma once
#include "pwv_tc.cuh"

namespace pwv {

constexpr int TH_BOX_BYTES = TC_TM * 128;                 // 128 rows x 64 channels x 2 B
constexpr int TH_STAGE_BYTES = 4 * TH_BOX_BYTES;          // per slot: X boxes (<= 2 planes) | Y boxes (<= 2 planes)
constexpr int TH_SMEM_STAGE0 = ((TC_IMAGE_BYTES + 1023) / 1024) * 1024;
constexpr int TH_SMEM_CB0 = TH_SMEM_STAGE0 + 2 * TH_STAGE_BYTES;      // [slot][tile parity] conditioning rows
constexpr int TH_SMEM_BARS = TH_SMEM_CB0 + 4 * TC_CB_BYTES;
constexpr int TH_SMEM_BYTES = TH_SMEM_BARS + 256;
constexpr int TH_WORKER_REGS = 112, TH_HELPER_REGS = 32;   // 512 x 112 + 128 x 32 = 61,440 = 640 x 96, the CTA's launch allocation

static_assert(TH_WORKER_REGS * 512 + TH_HELPER_REGS * 128 <= 96 * 640, "setmaxnreg redistributes the CTA's launch allocation only (pwv_tc.cuh)");

struct ThLayerParams {
  const uint8_t* image[2];  // per body: the layer's weight image (TC_IMAGE_BYTES, as for k_layer_tc)
  const float* cbias[2];    // per body [N][t_mel][128], pre-scaled (filter half by KF, gate half by KG)
  int N, T, t_mel, hop, dilation;
  // template flag LAST = false: out = x + z.Wd + bd as planes through map_out; true: the flow's last gated layer,
  // out = z as fp32 rows (map_out = the fp32 map k_post_tc reads: two boxes of 32 channels)
  int tiles_per_utt, cb_in_smem;
  float* z_out;             // use_skip_connection: the gate output z of a non-last layer as fp32 rows [2][N][T][64] (else nullptr)
  const int* flags_in;      // tile handshake with the previous gated layer (see TcLayerParams)
  int* flags_out;
  int prev_dilation;
  // Whole-layer completion counters (one int per layer, zeroed with the flags): a CTA adds 1 when all its tiles are
  // stored. Once a producer has seen the previous layer's counter at `done_target` every tile it could wait for is
  // published, and the per-tile probes (5 gpu-scope loads + an acquire fence + a proxy fence: ~3k cycles per tile on
  // the producer's serial path, profiles/r2_trace_layer_h_v1.txt) are skipped for the rest of the launch.
  const int* done_in;
  int* done_out;
  int done_target;
  // Measured and left off (profiles/r2_ab_layer_h.txt: no gain; with the MMA lock released between the halves the two
  // slots' GEMMs interleave and the slots fall into lock-step, the effect DESIGN 4.1 describes for round 1):
  int use_cp;               // 1: the MMA issuer moves the landed boxes into TMEM itself (tcgen05.cp) instead of the workers
  int split1;               // 1: GEMM1 starts on the x[t-d] half of K as soon as those columns are copied (the x[t] boxes land later)
  int z_in_d;               // f16x3: the gate output goes into the dead gate-accumulator columns instead of the x[t-d] operand
                            //   columns, so the A columns are free right after GEMM1 and the next tile is copied while GEMM2 runs (default 1)
  int double_a;             // bf16: A operand double-buffered by tile parity, next tile copied while GEMM2 runs (default 1)
  int split2;               // 1: GEMM2 starts on the first 16-channel chunk of each half of z while the gate computes the second
  long long* trace;
};

struct ThBarriers {
  uint64_t w_ready;
  uint64_t x_full[2], y_full[2], ax_ready[2], ay_ready[2], d1_ready[2], za_ready[2], zb_ready[2], d2_ready[2], out_ready[2];
  uint64_t a_free[2], boxes_free[2];      // use_cp: slot's TMEM free for the next tile (8 warp arrivals) / its boxes copied (commit)
  uint32_t tmem_base;
  int mma_lock;
};

// four consecutive logical 16-byte chunks c0 .. c0+3 of row r of a 128B-swizzled box -> 16 registers
__device__ __forceinline__ void th_ld_row64(const uint8_t* box, int r, int c0, uint32_t (&v)[16]) {
  const uint8_t* row = box + r * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 q = *reinterpret_cast<const uint4*>(row + ((((c0 + c) ^ r) & 7) << 4));
    v[4 * c + 0] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
  }
}
__device__ __forceinline__ void th_st_row64(uint8_t* box, int r, int c0, const uint32_t (&v)[16]) {
  uint8_t* row = box + r * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(row + ((((c0 + c) ^ r) & 7) << 4)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// TANH: the gate uses tanh.approx (one MUFU per transcendental, 2^-11 accuracy: bf16 mode) instead of the ex2 / rcp form
template <bool BF16, bool LAST, bool PK = false, bool TANH = BF16>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_layer_h(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out, ThLayerParams p) {
  using namespace ptx;
  constexpr bool SPLIT = !BF16;
  constexpr int P = BF16 ? 1 : 2;                         // planes
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* smem = tc_smem;
  ThBarriers* bars = reinterpret_cast<ThBarriers*>(smem + TH_SMEM_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int body = blockIdx.x & 1;
  const int cta_in_body = blockIdx.x >> 1, ctas_per_body = (gridDim.x + 1 - body) >> 1;
  const int tiles_body = p.N * p.tiles_per_utt;
  const int n_local = (tiles_body > cta_in_body) ? (tiles_body - cta_in_body + ctas_per_body - 1) / ctas_per_body : 0;

  pdl_launch_dependents();
  if (warp == TC_MMA_WARP) {
    if (lane == 0) {
      mbar_init(&bars->w_ready, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&bars->x_full[s], 1);
        mbar_init(&bars->y_full[s], 1);
        // (worker hand-offs: ONE arrival per warp -- lane 0, after the warp's lanes have fenced and met at __syncwarp.
        //  256 per-thread arrivals on one mbarrier serialise in the shared-memory atomic unit, ~32 cycles per warp, and
        //  that sat on the critical path of every hand-off: profiles/r2_trace_layer_h_v4*.txt)
        mbar_init(&bars->ax_ready[s], 8);
        mbar_init(&bars->ay_ready[s], 8);
        mbar_init(&bars->d1_ready[s], 1);
        mbar_init(&bars->za_ready[s], 8);
        mbar_init(&bars->zb_ready[s], 8);
        mbar_init(&bars->d2_ready[s], 1);
        mbar_init(&bars->out_ready[s], 8);
        mbar_init(&bars->a_free[s], 8);
        mbar_init(&bars->boxes_free[s], 1);
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

  // Register re-partition: each role's code sits entirely inside its own branch after the setmaxnreg, so ptxas
  // allocates the helper branch with TH_HELPER_REGS and the worker branch with TH_WORKER_REGS (code after a merge of
  // the two would be held to the smaller budget).
  if (warp >= TC_MMA_WARP) {
   reg_dec<TH_HELPER_REGS>();
   if (warp < TC_TMA_WARP) {
    // ======================= MMA issuers (one per tile slot); slot 0's also loads the weights =======================
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
      // descriptors: K-step ks of W1 (128 cols) advances 4096 B, of W2 (64 cols) 2048 B -> +256 / +128 in the address field
      const uint64_t dW1hi = smem_desc_kmajor_noswizzle(smem_u32(smem + TC_OFF_W1HI), 2048, 128);
      const uint64_t dW1lo = smem_desc_kmajor_noswizzle(smem_u32(smem + TC_OFF_W1LO), 2048, 128);
      const uint64_t dW2hi = smem_desc_kmajor_noswizzle(smem_u32(smem + TC_OFF_W2HI), 1024, 128);
      const uint64_t dW2lo = smem_desc_kmajor_noswizzle(smem_u32(smem + TC_OFF_W2LO), 1024, 128);
      constexpr uint32_t ID1 = idesc_f16(128, 128, BF16), ID2 = idesc_f16(128, 64, BF16);
      const uint32_t tD = tmem + s * 256;
      const uint32_t tAlo = tD + 192;
      // bf16 has one operand plane, so columns [192, 256) of the slot are free: the A operand is double-buffered by tile
      // parity there, and the workers copy tile j+1 while GEMM2 of tile j runs (see the worker loop)
      const bool db = BF16 && !p.use_cp && p.double_a;
      // f16x3 has no free columns, but the gate accumulators [64, 128) are dead once a worker has read them: z (hi + lo,
      // 16 columns per 16-channel chunk) goes THERE, D2 into [0, 64), and the A columns belong to the next tile as soon as
      // GEMM1 is through and the workers have taken x[t] into registers
      const bool zd = !BF16 && !p.use_cp && p.z_in_d;
      const bool split1 = p.split1 && !db && !zd;
      auto tA_of = [&](int j) { return tD + 128 + (db ? (uint32_t)(j & 1) * 64 : 0u); };
      const int tiles_s = (n_local + 1 - s) / 2;
      // D1 = A1lo.W1hi + A1hi.W1lo + A1hi.W1hi over K steps [k0, k1) (a step = 16 channels = 8 TMEM columns of A and two
      // 16-byte K-chunks of B); D2 likewise over the listed steps of z.
      auto gemm1_part = [&](uint32_t tAhi, int k0, int k1, uint32_t acc) {
        if (SPLIT) {
#pragma unroll 1
          for (int ks = k0; ks < k1; ++ks, acc = 1) mma_f16_ts(tD, tAlo + ks * 8, dW1hi + (uint64_t)(ks * 256), ID1, acc);
#pragma unroll 1
          for (int ks = k0; ks < k1; ++ks) mma_f16_ts(tD, tAhi + ks * 8, dW1lo + (uint64_t)(ks * 256), ID1, 1);
        }
#pragma unroll 1
        for (int ks = k0; ks < k1; ++ks, acc = 1) mma_f16_ts(tD, tAhi + ks * 8, dW1hi + (uint64_t)(ks * 256), ID1, acc);
      };
      auto gemm2_part = [&](uint32_t tAhi, int k0, int kstep, uint32_t acc) {      // steps k0, k0 + kstep, ... < 4
        // z of step ks: columns ks * 8 of the hi / lo operand planes, or (zd) hi at 64 + ks * 16, lo 8 columns further
        const uint32_t zhi = zd ? tD + 64 : tAhi, zlo = zd ? tD + 72 : tAlo, zstep = zd ? 16 : 8;
        if (SPLIT) {
#pragma unroll 1
          for (int ks = k0; ks < 4; ks += kstep, acc = 1) mma_f16_ts(tD, zlo + ks * zstep, dW2hi + (uint64_t)(ks * 128), ID2, acc);
#pragma unroll 1
          for (int ks = k0; ks < 4; ks += kstep) mma_f16_ts(tD, zhi + ks * zstep, dW2lo + (uint64_t)(ks * 128), ID2, 1);
        }
#pragma unroll 1
        for (int ks = k0; ks < 4; ks += kstep, acc = 1) mma_f16_ts(tD, zhi + ks * zstep, dW2hi + (uint64_t)(ks * 128), ID2, acc);
      };
      const uint32_t box0 = smem_u32(smem + TH_SMEM_STAGE0 + s * TH_STAGE_BYTES);
      for (int j = 0; j < tiles_s; ++j) {
        const uint32_t par = j & 1;
        const uint32_t tAhi = tA_of(j);
        if (p.use_cp) {
          // the landed boxes go into the slot's A columns by tcgen05.cp, in order with the MMAs that read them
          if (j > 0) mbar_wait(&bars->a_free[s], (j - 1) & 1);      // the workers have read D2 and x of the slot's previous tile
          mbar_wait(&bars->x_full[s], par);
          mbar_wait(&bars->y_full[s], par);
          tc_lock<SPLIT>(&bars->mma_lock);
          tc_fence_after_sync();
          TC_TRACE(2, j, s * 8 + 0);
#pragma unroll 1
          for (int b = 0; b < 2; ++b)            // b = 0: x[t-d] boxes -> K 0..63, b = 1: x[t] boxes -> K 64..127
#pragma unroll 1
            for (int q = 0; q < P; ++q)
#pragma unroll 1
              for (int k = 0; k < 4; ++k)
                tmem_cp_128x256b((q ? tAlo : tAhi) + b * 32 + k * 8, smem_desc_kmajor_sw128(box0 + (2 * b + q) * TH_BOX_BYTES + k * 32));
          mma_commit(&bars->boxes_free[s]);
          gemm1_part(tAhi, 0, 8, 0);
          mma_commit(&bars->d1_ready[s]);
          tc_unlock<SPLIT>(&bars->mma_lock);
          TC_TRACE(2, j, s * 8 + 1);
        } else {
        mbar_wait(&bars->ax_ready[s], par);
        if (split1) {                      // the x[t-d] half of K (columns copied first) ...
          tc_lock<SPLIT>(&bars->mma_lock);
          tc_fence_after_sync();
          TC_TRACE(2, j, s * 8 + 0);
          gemm1_part(tAhi, 0, 4, 0);
          tc_unlock<SPLIT>(&bars->mma_lock);
        }
        mbar_wait(&bars->ay_ready[s], par);
        tc_lock<SPLIT>(&bars->mma_lock);
        tc_fence_after_sync();
        if (split1) {                      // ... then the x[t] half
          gemm1_part(tAhi, 4, 8, 1);
        } else {
          TC_TRACE(2, j, s * 8 + 0);
          gemm1_part(tAhi, 0, 8, 0);
        }
        mma_commit(&bars->d1_ready[s]);
        tc_unlock<SPLIT>(&bars->mma_lock);
        TC_TRACE(2, j, s * 8 + 1);
        }
        if (LAST) continue;
        // z columns: step 0 / 1 = first / second 16-channel chunk of the workers' half 0, steps 2 / 3 of half 1
        mbar_wait(&bars->za_ready[s], par);
        if (p.split2) {
          tc_lock<SPLIT>(&bars->mma_lock);
          tc_fence_after_sync();
          TC_TRACE(2, j, s * 8 + 2);
          gemm2_part(tAhi, 0, 2, 0);
          tc_unlock<SPLIT>(&bars->mma_lock);
        }
        mbar_wait(&bars->zb_ready[s], par);
        tc_lock<SPLIT>(&bars->mma_lock);
        tc_fence_after_sync();
        if (p.split2) {
          gemm2_part(tAhi, 1, 2, 1);
        } else {
          TC_TRACE(2, j, s * 8 + 2);
          gemm2_part(tAhi, 0, 1, 0);
        }
        mma_commit(&bars->d2_ready[s]);
        tc_unlock<SPLIT>(&bars->mma_lock);
        TC_TRACE(2, j, s * 8 + 3);
      }
    }
    __syncwarp();
   } else {
    // ======================= producers (one per tile slot): boxes in, staged output out =======================
    if (elect_one()) {
      const int s = warp - TC_TMA_WARP;
      tma_prefetch_desc(&map_in);
      tma_prefetch_desc(&map_out);
      uint8_t* st = smem + TH_SMEM_STAGE0 + s * TH_STAGE_BYTES;
      const float* cbias = body ? p.cbias[1] : p.cbias[0];
      const int tiles_s = (n_local + 1 - s) / 2;
      const int ub0 = body * p.N;                       // utterance-body index of utterance 0 of this body
      auto tile_of = [&](int j) { return cta_in_body + (s + 2 * j) * ctas_per_body; };
      auto issue_x = [&](int j) {                       // x[t-d] boxes + the conditioning rows of the tile's frames
        const int tile = tile_of(j);
        const int n = tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * TC_TM;
        uint32_t bytes = P * TH_BOX_BYTES, cbytes = 0;
        int f0 = 0;
        if (p.cb_in_smem) {
          f0 = (t0 + p.hop / 2) / p.hop;
          const int f1 = (min(t0 + TC_TM - 1, p.T - 1) + p.hop / 2) / p.hop;
          cbytes = (uint32_t)(f1 - f0 + 1) * 512;
        }
        mbar_arrive_expect_tx(&bars->x_full[s], bytes + cbytes);
#pragma unroll
        for (int q = 0; q < P; ++q) tma_load_3d(st + q * TH_BOX_BYTES, &map_in, 0, t0 - p.dilation, q * 2 * p.N + ub0 + n, &bars->x_full[s]);
        if (cbytes)
          bulk_g2s(smem + TH_SMEM_CB0 + (s * 2 + (j & 1)) * TC_CB_BYTES, cbias + ((size_t)n * p.t_mel + f0) * 128, cbytes, &bars->x_full[s]);
      };
      auto issue_y = [&](int j) {                       // x[t] boxes
        const int tile = tile_of(j);
        const int n = tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * TC_TM;
        mbar_arrive_expect_tx(&bars->y_full[s], P * TH_BOX_BYTES);
#pragma unroll
        for (int q = 0; q < P; ++q) tma_load_3d(st + (2 + q) * TH_BOX_BYTES, &map_in, 0, t0, q * 2 * p.N + ub0 + n, &bars->y_full[s]);
      };
      bool prev_done = p.flags_in == nullptr;           // no handshake: the whole previous kernel is waited for below
      auto wait_tiles = [&](int j) {                    // see k_layer_tc: the previous layer's tiles this tile reads / overwrites
        if (prev_done) return;
        if (ld_relaxed_gpu(p.done_in) >= p.done_target) {      // the previous layer has stored every tile
          fence_acq_rel_gpu();
          fence_proxy_async_global();
          prev_done = true;
          return;
        }
        const int tile = tile_of(j);
        const int n = tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * TC_TM;
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
        fence_acq_rel_gpu();
        fence_proxy_async_global();
      };
      auto publish = [&](int j) {                       // tile j's rows are in global memory: let the next layer read them
        bulk_wait0();
        fence_proxy_async_global();
        fence_acq_rel_gpu();
        st_relaxed_gpu(p.flags_out + (size_t)body * tiles_body + tile_of(j), 1);
      };
      if (!p.flags_in) pdl_wait_prior_grid();
      if (s == 1 && n_local > 0) mbar_wait(&bars->y_full[0], 0);       // half-phase stagger of the two slots
      if (tiles_s > 0) {
        wait_tiles(0);
        issue_x(0);
        issue_y(0);
      }
      uint64_t* const boxes_copied = p.use_cp ? bars->boxes_free : bars->ay_ready;     // tile's boxes are in TMEM
      uint64_t* const xboxes_copied = p.use_cp ? bars->boxes_free : bars->ax_ready;
      if (tiles_s > 1) {                                // both landing areas are free once tile 0 sits in TMEM
        wait_tiles(1);
        mbar_wait(&boxes_copied[s], 0);
        issue_x(1);
        issue_y(1);
      }
      if (tiles_s > 2) wait_tiles(2);                   // the probes run one tile ahead, in the producer's idle time
      for (int j = 0; j < tiles_s; ++j) {
        if (j + 1 < tiles_s) {
          mbar_wait(&xboxes_copied[s], (j + 1) & 1);    // tile j+1's x[t-d] columns are in TMEM: the X boxes are free
          if (j + 2 < tiles_s) {
            issue_x(j + 2);
            TC_TRACE(3, j, s * 8 + 0);
          }
        }
        mbar_wait(&bars->out_ready[s], j & 1);          // tile j's output is staged in the Y boxes (writers fenced)
        {
          const int tile = tile_of(j);
          const int n = tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * TC_TM;
          if (LAST) {                                   // z, fp32 rows: two boxes of 32 channels
            tma_store_3d(&map_out, 0, t0, ub0 + n, st + 2 * TH_BOX_BYTES);
            tma_store_3d(&map_out, 32, t0, ub0 + n, st + 3 * TH_BOX_BYTES);
          } else {
#pragma unroll
            for (int q = 0; q < P; ++q) tma_store_3d(&map_out, 0, t0, q * 2 * p.N + ub0 + n, st + (2 + q) * TH_BOX_BYTES);
          }
          bulk_commit();
          bulk_wait_read0();                            // the TMA unit has read the staging boxes
        }
        if (j + 2 < tiles_s) issue_y(j + 2);
        else if (j + 1 < tiles_s) mbar_arrive(&bars->y_full[s]);   // phase tiles_s: the slot's last tile waits for it before staging
        TC_TRACE(3, j, s * 8 + 2);
        if (p.flags_out) publish(j);
        if (j + 3 < tiles_s) wait_tiles(j + 3);
      }
      bulk_wait0();
      if (p.done_out) {                                 // this slot's share of the CTA is stored (2 arrivals per CTA)
        fence_proxy_async_global();
        fence_acq_rel_gpu();
        atomicAdd(p.done_out, 1);
      }
    }
    __syncwarp();
   }
  } else {
    reg_inc<TH_WORKER_REGS>();
    // ======================= workers: operand copy, gate, residual =======================
    const int slot = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t tD = tmem + slot * 256 + lane_base;
    const uint32_t tAlo = tD + 192;
    const bool db = BF16 && !p.use_cp && p.double_a;          // bf16: A operand double-buffered by tile parity (columns [192, 256) are free)
    const bool zd = !BF16 && !p.use_cp && p.z_in_d;           // f16x3: z into the dead gate-accumulator columns (see the MMA issuer)
    const bool split1 = p.split1 && !db && !zd;
    auto tA_of = [&](int j) { return tD + 128 + (db ? (uint32_t)(j & 1) * 64 : 0u); };
    uint8_t* stage = smem + TH_SMEM_STAGE0 + slot * TH_STAGE_BYTES;
    const float* bd_s = reinterpret_cast<const float*>(smem + TC_OFF_BD) + half * 32;
    const float* scal = reinterpret_cast<const float*>(smem + TC_OFF_SCAL);
    const bool tracer = (warp & 7) == 0 && lane == 0;
    const int n_s = (n_local + 1 - slot) / 2;

    auto warp_arrive = [&](uint64_t* bar) {     // every lane has fenced its own writes; the warp arrives once
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // boxes -> the slot's A columns, verbatim: K = channel (x[t-d]) / 64 + channel (x[t]), two 16-bit elements per column
    // `hand_over` false: the columns are written but the tile is not announced yet (a_copy_announce does that later)
    auto a_copy = [&](int jn, bool hand_over) {
      const uint32_t par = jn & 1, tAhi = tA_of(jn);
      uint32_t v[16];
      mbar_wait(&bars->x_full[slot], par);
#pragma unroll
      for (int q = 0; q < P; ++q) {
        th_ld_row64(stage + q * TH_BOX_BYTES, r, half * 4, v);
        tmem_st16((q ? tAlo : tAhi) + half * 16, v);
      }
      if (split1) {                             // the x[t-d] half of K is handed over on its own: GEMM1 may start on it
        tmem_wait_st();
        tc_fence_before_sync();
        warp_arrive(&bars->ax_ready[slot]);
      }
      mbar_wait(&bars->y_full[slot], par);
#pragma unroll
      for (int q = 0; q < P; ++q) {
        th_ld_row64(stage + (2 + q) * TH_BOX_BYTES, r, half * 4, v);
        tmem_st16((q ? tAlo : tAhi) + 32 + half * 16, v);
      }
      tmem_wait_st();
      if (!hand_over) return;
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (!split1) mbar_arrive(&bars->ax_ready[slot]);
        mbar_arrive(&bars->ay_ready[slot]);
      }
    };
    // One arrival says two things to the MMA issuer: the next tile's operand is in its A columns, and this warp has read
    // its part of the accumulator the next GEMM1 overwrites.
    auto a_copy_announce = [&]() {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars->ax_ready[slot]);
        mbar_arrive(&bars->ay_ready[slot]);
      }
    };

    if (n_s > 0) {
      if (tracer) TC_TRACE(slot, 0, 0);
      if (!p.use_cp) a_copy(0, true);
      if (tracer) TC_TRACE(slot, 0, 4);
      mbar_wait(&bars->w_ready, 0);
    }
    const float sf = scal[0], sg = scal[1], s2 = scal[2];

#pragma unroll 1
    for (int j = 0; j < n_s; ++j) {
      const int tile = cta_in_body + (slot + 2 * j) * ctas_per_body;
      const int n = tile / p.tiles_per_utt, t_first = (tile % p.tiles_per_utt) * TC_TM, t = t_first + r;
      const uint32_t par = j & 1;
      const uint32_t tAhi = tA_of(j);
      const int frame = (min(t, p.T - 1) + p.hop / 2) / p.hop;
      const float4* cb;
      if (p.cb_in_smem) {               // staged with the tile's x[t-d] boxes: row (frame - first frame of the tile)
        const int f0 = (t_first + p.hop / 2) / p.hop;
        cb = reinterpret_cast<const float4*>(smem + TH_SMEM_CB0 + (slot * 2 + par) * TC_CB_BYTES) + (frame - f0) * 32 + half * 8;
      } else {
        cb = reinterpret_cast<const float4*>((body ? p.cbias[1] : p.cbias[0]) + ((size_t)n * p.t_mel + frame) * 128) + half * 8;
      }

      // ---- gate: z = tanh(f) * sigmoid(g) on my 32 channels
      // (use_cp: the conditioning rows rode on x_full, which the MMA issuer waited for before the GEMM whose completion is
      //  awaited here -- the workers must NOT wait on x_full themselves: by now it may be two phases ahead and a parity
      //  wait would alias, measured as a hang)
      mbar_wait_sleepy(&bars->d1_ready[slot], par);
      tc_fence_after_sync();
      if (tracer) TC_TRACE(slot, j, 5);
      float zz[LAST ? 32 : 1];                      // LAST: z kept until it is staged
      uint32_t dr[LAST ? 1 : 32], xh[LAST ? 1 : 16], xl[LAST ? 1 : 16];   // otherwise: D2 and the x[t] operand columns
      {
        // the second 16-channel chunk's accumulators are requested before the first chunk is computed: its TMEM round
        // trip hides behind the first chunk's arithmetic (64 registers in flight)
        uint32_t f0r[16], g0r[16], f1r[16], g1r[16];
        tmem_ld16(tD + half * 32, f0r);
        tmem_ld16(tD + 64 + half * 32, g0r);
        tmem_wait_ld();
        tmem_ld16(tD + half * 32 + 16, f1r);
        tmem_ld16(tD + 64 + half * 32 + 16, g1r);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float z[16];
          if (c == 1) tmem_wait_ld();
          if (c == 0) tc_gate<BF16, PK, 16, TANH>(f0r, g0r, cb, cb + 16, sf, sg, z);
          else tc_gate<BF16, PK, 16, TANH>(f1r, g1r, cb + 4, cb + 20, sf, sg, z);
          if constexpr (LAST) {
#pragma unroll
            for (int e = 0; e < 16; ++e) zz[c * 16 + e] = z[e];
          } else {
            uint32_t hi[8], lo[8];
            float v0[8], v1[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { v0[e] = z[e]; v1[e] = z[8 + e]; }
            split8x<BF16, SPLIT, PK>(v0, hi, lo);
            split8x<BF16, SPLIT, PK>(v1, hi + 4, lo + 4);
            // (zd: chunk c of my half lands on the 16 gate-accumulator columns I have just consumed -- g0r / g1r)
            const uint32_t zc = tD + 64 + half * 32 + c * 16;
            tmem_st8(zd ? zc : tAhi + half * 16 + c * 8, hi);
            if (SPLIT) tmem_st8(zd ? zc + 8 : tAlo + half * 16 + c * 8, lo);
            if (c == 1 || p.split2) {          // (without the GEMM2 split both halves are handed over together)
              tmem_wait_st();
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) {
                if (c == 0 || !p.split2) mbar_arrive(&bars->za_ready[slot]);
                if (c == 1) mbar_arrive(&bars->zb_ready[slot]);
              }
            }
            if (p.z_out && t < p.T) {    // use_skip_connection (non-default): every layer's z feeds the skip sum (k_skip_simt)
              float4* zo = reinterpret_cast<float4*>(p.z_out + (((size_t)body * p.N + n) * p.T + t) * TC_C + half * 32 + c * 16);
#pragma unroll
              for (int q = 0; q < 4; ++q) zo[q] = make_float4(z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
            }
          }
        }
      }
      // bf16 (double-buffered A): the next tile's boxes go into the other A buffer NOW, while GEMM2 of this tile runs; the
      // copy leaves the slot's critical chain (gate -> GEMM2 -> read-out -> GEMM1 of the next tile)
      // f16x3 (zd): GEMM1 is through and z sits in the accumulator columns, so once x[t] of this tile is in registers the
      // A columns are free and take the next tile at the same point.
      const bool early = !LAST && (db || zd) && j + 1 < n_s;
      if constexpr (!LAST) {
        if (tracer) TC_TRACE(slot, j, 6);
        if (zd) {
          tmem_ld16(tAhi + 32 + half * 16, xh);
          if (SPLIT) tmem_ld16(tAlo + 32 + half * 16, xl);
          tmem_wait_ld();
        }
        if (early) a_copy(j + 1, false);
        // ---- D2 and x[t] (hi, lo) of my 32 channels -> registers; after this the slot's TMEM belongs to the next tile
        mbar_wait_sleepy(&bars->d2_ready[slot], par);
        tc_fence_after_sync();
        if (tracer) TC_TRACE(slot, j, 7);
        tmem_ld32(tD + half * 32, dr);
        if (!zd) {
          tmem_ld16(tAhi + 32 + half * 16, xh);
          if (SPLIT) tmem_ld16(tAlo + 32 + half * 16, xl);
        }
        tmem_wait_ld();
      }
      if (tracer) TC_TRACE(slot, j, 9);
      // ---- the slot's next tile goes into TMEM (its boxes landed long ago); GEMM1(j+1) starts when all 256 are through
      if (p.use_cp) {
        tc_fence_before_sync();
        warp_arrive(&bars->a_free[slot]);            // my part of D2 / x is in registers: the slot's TMEM may take the next tile
        if (j + 1 < n_s) mbar_wait(&bars->boxes_free[slot], (j + 1) & 1);   // ... whose boxes must have left before they become the staging
        else if (j >= 1) mbar_wait(&bars->y_full[slot], (j + 1) & 1);
      } else if (early) {
        a_copy_announce();
      } else if (j + 1 < n_s) {
        a_copy(j + 1, true);
      } else {
        tc_fence_before_sync();
        if (j >= 1) mbar_wait(&bars->y_full[slot], (j + 1) & 1);   // the store of tile j-1 has read the staging boxes
      }
      if (tracer) TC_TRACE(slot, j, 10);
      // The Y boxes become the staging of this tile's output. A thread's staging bytes (row r, chunks 4 half .. 4 half + 3
      // of each plane) are exactly the bytes it read itself in a_copy, so no barrier is needed -- except in the flow's last
      // layer, whose fp32 rows are twice as wide and overlap the other half's reads.
      if constexpr (LAST) named_bar_sync(1 + slot, 256);
      // ---- residual + split (mode 0) / z (mode 1) -> staging boxes
      if constexpr (LAST) {
        uint8_t* row = stage + (2 + half) * TH_BOX_BYTES + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(row + (((c ^ r) & 7) << 4)) = make_float4(zz[4 * c], zz[4 * c + 1], zz[4 * c + 2], zz[4 * c + 3]);
      } else {
        uint8_t* row_h = stage + 2 * TH_BOX_BYTES + r * 128;
        uint8_t* row_l = stage + 3 * TH_BOX_BYTES + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // 8 channels = one 16-byte chunk of each plane per step
          const float4 b0 = *reinterpret_cast<const float4*>(bd_s + q * 8), b1 = *reinterpret_cast<const float4*>(bd_s + q * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float o[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 h2 = unpack16<BF16>(xh[q * 4 + e]);
            float x0 = h2.x, x1 = h2.y;
            if (SPLIT) {
              const float2 l2 = unpack16<BF16>(xl[q * 4 + e]);
              x0 += l2.x;               // exact: hi + lo has <= 23 significant bits
              x1 += l2.y;
            }
            o[2 * e] = x0 + fmaf(__uint_as_float(dr[q * 8 + 2 * e]), s2, bb[2 * e]);
            o[2 * e + 1] = x1 + fmaf(__uint_as_float(dr[q * 8 + 2 * e + 1]), s2, bb[2 * e + 1]);
          }
          uint32_t oh[4], ol[4];
          split8x<BF16, SPLIT, PK>(o, oh, ol);
          const int sw = (((half * 4 + q) ^ r) & 7) << 4;
          *reinterpret_cast<uint4*>(row_h + sw) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
          if (SPLIT) *reinterpret_cast<uint4*>(row_l + sw) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
        }
      }
      fence_proxy_async_smem();         // my staging writes -> visible to the TMA store the producer issues
      warp_arrive(&bars->out_ready[slot]);
      if (tracer) TC_TRACE(slot, j, 8);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == TC_MMA_WARP) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// Flow front for the plane layout: IAF combine of the previous flow + causal layer of both bodies
// (reference modules.py:57-59, 174-183), written as 16-bit planes. One thread per (sample, 8 channels).
// ------------------------------------------------------------------------------------------------
struct FrontHParams {
  const float* x_prev;    // [N][T]
  const float* scale;     // [N][T] or nullptr (first flow)
  const float* shift;
  float* x_new;           // [N][T]
  const float* wc[2];     // per body: [2][64]
  uint16_t* act;          // [planes][2N][T][64]
  int N, T;
};

template <bool BF16>
__global__ void __launch_bounds__(256) k_front_h(FrontHParams p) {
  constexpr int C = TC_C;
  const size_t total = (size_t)p.N * p.T * (C / 8);
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = (int)(idx % (C / 8));
  const size_t row = idx / (C / 8);             // n*T + t
  const int t = (int)(row % p.T);
  float xc = p.x_prev[row];
  float xp = (t > 0) ? p.x_prev[row - 1] : 0.f;
  if (p.scale) {
    xc = xc * p.scale[row] + p.shift[row];
    if (t > 0) xp = xp * p.scale[row - 1] + p.shift[row - 1];
  }
  if (cg == 0) p.x_new[row] = xc;
  const size_t plane = (size_t)2 * p.N * p.T * C;       // elements per plane
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

// debug tap: planes of one body -> fp32 rows [N][T][64]
template <bool BF16>
__global__ void __launch_bounds__(256) k_planes_to_f32(const uint16_t* __restrict__ act, float* __restrict__ out, size_t rows, size_t plane_elems) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // (row, pair of channels)
  if (idx >= rows * 32) return;
  const uint32_t h = reinterpret_cast<const uint32_t*>(act)[idx];
  float2 v = unpack16<BF16>(h);
  if (!BF16) {
    const float2 l = unpack16<BF16>(reinterpret_cast<const uint32_t*>(act + plane_elems)[idx]);
    v.x += l.x;
    v.y += l.y;
  }
  reinterpret_cast<float2*>(out)[idx] = v;
}

}  // namespace pwv
