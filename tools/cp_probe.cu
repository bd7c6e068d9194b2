// cp_probe.cu -- does tcgen05.cp (shared memory -> tensor memory, issued by one thread, executed by the tensor
// pipe) move a 128B-swizzled K-major TMA box [128 rows][64 x 16 bit] into the TMEM A-operand layout of kind::f16
// (lane = row, column j = elements 2j, 2j+1), and what does it cost? Round-2 question: k_layer_h's workers copy the
// landed boxes into TMEM by ld.shared + tcgen05.st (64 KB per tile through the LSU, ~1.0-1.4k cycles of the per-slot
// chain); tcgen05.cp would take that off the workers. Run on a B200 through gpurun; results in profiles/r2_cp_probe.txt.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/cp_probe tools/cp_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

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

// shared-memory matrix descriptor, K-major, 128B swizzle: [0,14) addr>>4, [16,30) LBO>>4 (ignored for swizzled K-major;
// 1), [32,46) SBO>>4 (8 rows x 128 B = 1024), [46,48) version 1, [61,64) layout: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}

struct Args {
  int layout;        // descriptor layout-type field
  int reps;          // timing: copies of the whole 64 KB set
  int* mismatches;   // [2]: mismatching elements / total
  long long* cycles; // [2]: issue, issue -> complete
  unsigned* sample;  // first 8 words read back by thread 1
};

// box: 128 rows x 128 B, 16-byte chunk c of row r at ((c ^ (r & 7)) << 4)
__global__ void __launch_bounds__(160, 1) k_cp_probe(Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];      // 4 boxes of 16 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 4) {
    tmem_alloc(&tmem_slot, 512);
    if (lane == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
  }
  // element (box b, row r, channel c) = a small integer exactly representable in fp16
  for (int e = threadIdx.x; e < 4 * 128 * 64; e += blockDim.x) {
    const int b = e / (128 * 64), r = (e / 64) % 128, c = e % 64;
    const int chunk = c / 8, within = c % 8;
    const float v = (float)((b * 577 + r * 13 + c * 3) % 2048);
    *reinterpret_cast<__half*>(smem + b * 16384 + r * 128 + ((chunk ^ (r & 7)) << 4) + within * 2) = __float2half(v);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_slot;
  if (warp == 4 && elect_one()) {
    long long t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep)
      for (int b = 0; b < 4; ++b)
        for (int k = 0; k < 4; ++k)          // K step of 16 elements = 32 bytes along the row = 8 TMEM columns
          tmem_cp_128x256b(tbase + b * 32 + k * 8, desc_sw128(smem_u32(smem + b * 16384) + k * 32, 1024, (uint32_t)p.layout));
    long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    p.cycles[0] = t1 - t0;
    p.cycles[1] = t2 - t0;
  }
  __syncthreads();
  if (warp < 4) {
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    const int r = threadIdx.x;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    int bad = 0;
    for (int b = 0; b < 4; ++b) {
      uint32_t v[32];
      tmem_ld32(tbase + lane_base + b * 32, v);
      tmem_wait_ld();
      for (int j = 0; j < 32; ++j) {
        const __half2 h = *reinterpret_cast<__half2*>(&v[j]);
        const float lo = __low2float(h), hi = __high2float(h);
        const float e0 = (float)((b * 577 + r * 13 + (2 * j) * 3) % 2048), e1 = (float)((b * 577 + r * 13 + (2 * j + 1) * 3) % 2048);
        bad += (lo != e0) + (hi != e1);
      }
      if (b == 0 && r == 1)
        for (int j = 0; j < 8; ++j) p.sample[j] = v[j];
    }
    atomicAdd(&p.mismatches[0], bad);
    if (r == 0) p.mismatches[1] = 4 * 128 * 64;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tbase, 512);
}

int main() {
  int* d_mis;
  long long* d_cyc;
  unsigned* d_s;
  CK(cudaMalloc(&d_mis, 8));
  CK(cudaMalloc(&d_cyc, 16));
  CK(cudaMalloc(&d_s, 32));
  CK(cudaFuncSetAttribute(k_cp_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int layouts[] = {2, 1, 4, 6, 0};     // 2 = SWIZZLE_128B (expected), others for the record
  for (int layout : layouts)
    for (int reps : {1, 8}) {
      CK(cudaMemset(d_mis, 0, 8));
      Args a{layout, reps, d_mis, d_cyc, d_s};
      k_cp_probe<<<1, 160, 65536>>>(a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("layout %d reps %d: %s\n", layout, reps, cudaGetErrorString(e));
        return 1;
      }
      int mis[2];
      long long cyc[2];
      unsigned s[8];
      CK(cudaMemcpy(mis, d_mis, 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(s, d_s, 32, cudaMemcpyDeviceToHost));
      printf("tcgen05.cp 128x256b, descriptor layout %d, %d x 64 KB: %d of %d elements wrong; issue %lld cycles, issue->complete %lld cycles (%.1f per 64 KB); row 1 words %08x %08x %08x %08x\n",
             layout, reps, mis[0], mis[1], cyc[0], cyc[1], (double)cyc[1] / reps, s[0], s[1], s[2], s[3]);
    }
  return 0;
}
