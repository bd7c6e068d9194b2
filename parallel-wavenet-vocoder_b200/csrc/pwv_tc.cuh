// pwv_tc.cuh -- tcgen05 (5th-gen tensor core) kernels of the gated dilated layers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace pwv {

struct TcLayerSrc {
  const float* wfg;   // host, [2C][2C] packed fp32 (tap rows, filter|gate cols)
  const float* wd;    // host, [C][C]
  const float* bd;    // host, [C]
};

struct TcModel {
  void* d_images = nullptr;
  size_t bytes = 0;
};

inline const char* tc_model_build(TcModel&, int, int, const std::vector<TcLayerSrc>&) {
  return "tensor-core kernels are not built into this library yet";
}
inline void tc_model_free(TcModel& t) {
  if (t.d_images) cudaFree(t.d_images);
  t.d_images = nullptr;
}

}  // namespace pwv
