// tc_probe.cu -- hardware probes that pin down the tcgen05 conventions the layer kernel relies on
// (run on a B200 through gpurun; results are recorded in profiles/tc_probe_r1.txt):
//   * kind::f16 MMA with A in TMEM (two 16-bit elements per 32-bit column: which half is k even?)
//   * K-major no-swizzle shared-memory descriptor for B (and A): LBO / SBO roles
//   * M=128 x N in {128, 64} accumulator layout read back with tcgen05.ld.32x32b
//   * 1-D bulk copies global -> padded shared rows -> global
//   * issue-to-completion cycles of back-to-back MMAs (TS and SS operand modes)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tools/tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../parallel-wavenet-vocoder_b200/csrc/pwv_ptx.cuh"

using namespace pwv::ptx;

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e = (x);                                                         \
    if (e != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

struct ProbeArgs {
  const __half* A;   // [128][K] row-major
  const __half* B;   // [N][K] row-major (so D = A . B^T)
  float* D;          // [128][N]
  int N, K;
  int a_in_tmem;     // 1: TS mode, 0: SS mode
  int pack_swap;     // TS: 1 = put k even in the HIGH half of the column
  int swap_lbo_sbo;  // descriptor experiment
  int reps;          // timing: repeat the whole K loop this many times (accumulating)
  long long* cycles;
};

// interleaved K-major layout: element (r, k) of an R-row operand
__device__ __forceinline__ uint32_t il_off(int r, int k, int R) { return (uint32_t)((k / 8) * (R * 16) + r * 16 + (k % 8) * 2); }

__global__ void __launch_bounds__(160, 1) k_probe(ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  uint8_t* sB = smem;                       // N x K fp16 interleaved
  uint8_t* sA = smem + 64 * 1024;           // 128 x K fp16 interleaved (SS mode)
  const int N = p.N, K = p.K;

  if (warp == 4) {
    tmem_alloc(&tmem_slot, 512);
    if (lane == 0) {
      mbar_init(&bar_done, 1);
      fence_mbar_init();
    }
  }
  // operands -> shared memory (generic proxy), all 160 threads
  for (int e = threadIdx.x; e < N * K; e += blockDim.x) {
    int n = e / K, k = e % K;
    *reinterpret_cast<__half*>(sB + il_off(n, k, N)) = p.B[e];
  }
  for (int e = threadIdx.x; e < 128 * K; e += blockDim.x) {
    int m = e / K, k = e % K;
    *reinterpret_cast<__half*>(sA + il_off(m, k, 128)) = p.A[e];
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t tD = tbase;            // columns [0, N)
  const uint32_t tA = tbase + 256;      // columns [256, 256 + K/2)

  if (warp < 4 && p.a_in_tmem) {
    // thread <-> row (lane of TMEM); pack two k per column
    const int m = threadIdx.x;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        __half lo = p.A[m * K + 2 * (c0 + j)], hi = p.A[m * K + 2 * (c0 + j) + 1];
        if (p.pack_swap) { __half t = lo; lo = hi; hi = t; }
        v[j] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
      }
      tmem_st8(tA + lane_base + c0, v);
    }
    tmem_wait_st();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  if (warp == 4) {
    if (elect_one()) {
      const uint32_t idesc = idesc_f16(128, N, false);
      uint32_t lboB = N * 16, sboB = 128, lboA = 128 * 16, sboA = 128;
      if (p.swap_lbo_sbo) { uint32_t t = lboB; lboB = sboB; sboB = t; t = lboA; lboA = sboA; sboA = t; }
      long long t0 = clock64();
      for (int rep = 0; rep < p.reps; ++rep) {
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint64_t bdesc = smem_desc_kmajor_noswizzle(smem_u32(sB) + ks * 2 * (N * 16), lboB, sboB);
          const uint32_t acc = (rep > 0 || ks > 0) ? 1u : 0u;
          if (p.a_in_tmem) {
            mma_f16_ts(tD, tA + ks * 8, bdesc, idesc, acc);
          } else {
            const uint64_t adesc = smem_desc_kmajor_noswizzle(smem_u32(sA) + ks * 2 * (128 * 16), lboA, sboA);
            mma_f16_ss(tD, adesc, bdesc, idesc, acc);
          }
        }
      }
      mma_commit(&bar_done);
      mbar_wait(&bar_done, 0);
      long long t1 = clock64();
      if (p.cycles) *p.cycles = t1 - t0;
    }
    __syncwarp();
  }
  __syncthreads();
  tc_fence_after_sync();
  if (warp < 4) {
    mbar_wait(&bar_done, 0);
    tc_fence_after_sync();
    const int m = threadIdx.x;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tD + lane_base + c0, v);
      tmem_wait_ld();
      for (int j = 0; j < 16; ++j) p.D[m * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tbase, 512);
}

// bulk copy probe: 128 rows of 256 B: global -> shared rows with a 272-byte pitch -> global
__global__ void __launch_bounds__(128, 1) k_bulk_probe(const float* src, float* dst, int rows_valid) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 128);
    fence_mbar_init();
  }
  __syncthreads();
  const int r = threadIdx.x;
  uint8_t* row = smem + r * 272;
  if (r < rows_valid) {
    mbar_arrive_expect_tx(&bar, 256);
    bulk_g2s(row, src + r * 64, 256, &bar);
  } else {
    mbar_arrive(&bar);
  }
  mbar_wait(&bar, 0);
  float4* rp = reinterpret_cast<float4*>(row);
  for (int c = 0; c < 16; ++c) {
    float4 v = (r < rows_valid) ? rp[c] : make_float4(0, 0, 0, 0);
    v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
    rp[c] = v;
  }
  fence_proxy_async_smem();
  bulk_s2g(dst + r * 64, row, 256);
  bulk_commit();
  bulk_wait_read0();
}

static float run_case(int N, int K, int ts, int pack_swap, int swap_ls, int reps, long long* cyc_out, const char* label) {
  std::vector<__half> hA(128 * K), hB(N * K);
  std::vector<float> fA(128 * K), fB(N * K);
  srand(1234);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
  __half *dA, *dB; float* dD; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 128 * N * 4));
  ProbeArgs p{dA, dB, dD, N, K, ts, pack_swap, swap_ls, reps, dC};
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  k_probe<<<1, 160, 128 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-40s : KERNEL ERROR %s\n", label, cudaGetErrorString(e)); exit(2); }
  std::vector<float> hD(128 * N);
  long long cyc = 0;
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * fB[n * K + k];
      ref *= reps;
      maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
    }
  printf("%-40s : N=%3d K=%3d reps=%4d  max|err|=%.3e  cycles=%lld (%.1f per MMA)\n", label, N, K, reps, maxerr, cyc,
         (double)cyc / (reps * (K / 16)));
  if (cyc_out) *cyc_out = cyc;
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return (float)maxerr;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, SMs %d, smem/block optin %zu\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  // correctness, one K loop
  run_case(128, 64, 1, 0, 0, 1, nullptr, "TS  pack(k even = low half)");
  run_case(128, 64, 1, 1, 0, 1, nullptr, "TS  pack(k even = HIGH half)");
  run_case(128, 64, 0, 0, 0, 1, nullptr, "SS  lbo=K-chunk stride, sbo=8-row stride");
  run_case(128, 64, 0, 0, 1, 1, nullptr, "SS  lbo/sbo swapped");
  run_case(64, 64, 1, 0, 0, 1, nullptr, "TS  N=64");
  run_case(128, 128, 1, 0, 0, 1, nullptr, "TS  K=128");
  run_case(64, 64, 0, 0, 0, 1, nullptr, "SS  N=64");
  // timing (errors grow with reps because the reference is scaled; only cycles matter)
  run_case(128, 128, 1, 0, 0, 64, nullptr, "TS  N=128 timing");
  run_case(128, 128, 0, 0, 0, 64, nullptr, "SS  N=128 timing");
  run_case(64, 64, 1, 0, 0, 128, nullptr, "TS  N=64 timing");
  run_case(64, 64, 0, 0, 0, 128, nullptr, "SS  N=64 timing");

  // bulk copies
  {
    std::vector<float> h(128 * 64);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *s, *d;
    CK(cudaMalloc(&s, h.size() * 4)); CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(s, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d, 0, h.size() * 4));
    CK(cudaFuncSetAttribute(k_bulk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 272));
    k_bulk_probe<<<1, 128, 128 * 272>>>(s, d, 100);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("bulk probe: KERNEL ERROR %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> o(h.size());
    CK(cudaMemcpy(o.data(), d, o.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 64; ++c) {
        float want = r < 100 ? h[r * 64 + c] + 1.f : 1.f;
        if (o[r * 64 + c] != want) ++bad;
      }
    printf("bulk g2s/s2g rows with 272-byte pitch: %d mismatches\n", bad);
  }
  return 0;
}
