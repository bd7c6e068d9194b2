// pwv_api.cu -- C-ABI (include/pwv.h) of the B200-native IAF-vocoder generation path:
// model/weight management and the launch sequence of one forward pass
// (reference models.py:23-78 -> modules.py:53-60 -> modules.py:129-166).
#include "../../include/pwv.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "pwv_simt.cuh"
#include "pwv_tc.cuh"
#include "pwv_tc2.cuh"
#include "pwv_tc3.cuh"
#include "pwv_mel.cuh"
#include "pwv_norm.cuh"
#include "pwv_gen.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
namespace {
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define PWV_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(PWV_ECUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                  cudaGetErrorString(e_));                                                    \
  } while (0)

struct VarSpec {
  std::string name;
  int64_t shape[4];
  int ndim;
  size_t numel() const {
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i];
    return n;
  }
};

const char* kBodies[2] = {"scalar", "shifter"};   // reference models.py:48,63 (sic)
}  // namespace

// per (flow, body) offsets (in floats) into the device weight arena
struct NormOff {                 // instance-normalisation variables of one call site (gamma then beta), or npos when off
  size_t gamma = (size_t)-1, beta = (size_t)-1;
};
struct LayerOff {
  size_t wfg, wd, bd, ws, bs;
  NormOff n_fg;                  // [2C]: normalize_filter | normalize_gate (reference modules.py:230-234)
  NormOff n_skip, n_dense;       // [S], [R] (modules.py:253-257)
};
struct BodyOff {
  size_t causal;                 // [2][C]
  std::vector<LayerOff> layers;
  size_t w1, b1, w2, b2;
  NormOff n_causal, n_pp1, n_pp2;   // [R], [S], [S] (modules.py:181-182,149-151,158-160)
};

#ifndef PWV_TC_VARIANT_DEFAULT
#define PWV_TC_VARIANT_DEFAULT 1
#endif

struct pwv_model {
  pwv_hparams hp;
  int C, S, Cc;                  // C = residual_channels
  int D = 0, taps = 2;           // dilation_channels, filter_width
  bool general = false;          // shapes outside the fused kernels' coverage: the un-fused fp32 chain of pwv_gen.cuh
  int total_layers;              // sum over flows
  int max_layers;                // max over flows
  std::vector<VarSpec> vars;
  std::map<std::string, int> var_index;
  std::vector<std::vector<float>> staged;
  std::vector<char> loaded;
  bool finalized = false;
  int device = 0;

  size_t off_wc = 0;             // cond dense [n_mels][Cc]
  size_t off_wup[PWV_MAX_UPSAMPLE] = {0, 0, 0, 0};   // transposed_conv stage i: [Cin][stride_i * Cc]
  std::vector<BodyOff> bodies;   // [flow*2 + body]
  float* d_arena = nullptr;
  size_t arena_floats = 0;

  // conditioning projections of a flow, contiguous so one batched GEMM covers them:
  // off_wgc[flow] -> [2 bodies][L][Cc][2C]  ([gc_filter | gc_gate]),  off_bfg[flow] -> [2][L][2C]
  NormOff n_cond;                // [Cc] (models.py:27-29)
  NormOff n_up[PWV_MAX_UPSAMPLE];   // [Cc] per transposed-conv stage (models.py:121-122)
  std::vector<NormOff> n_flow;   // [1] per flow (models.py:70)
  std::vector<size_t> off_wgc, off_bfg;
  size_t off_colscale = 0;       // [2C]: -2log2e (filter half), -log2e (gate half) for the tensor-core epilogue
  int num_sms = 148;
  size_t dev_total_bytes = 0;      // cudaMemGetInfo at finalize (a driver query: never on the per-call path)

  pwv::TcModel tc;               // tensor-core weight images (empty in fp32 mode)
  pwv::TwModel tw;               // weight streams of the wide tensor-core kernels (C > 64)

  // host-buffer entry point staging
  float* h_dev = nullptr;
  size_t h_dev_bytes = 0;
  int last_launches = 0;

  // profiling (pwv_set_profiling): event pairs around the gated-layer launches of the last forward
  long long* trace = nullptr;    // pwv_debug_set_trace
  // Debug / A-B switches, settable only through pwv_debug_set (tests, tools); the product path never reads the environment.
  bool use_pdl = true;           // "pdl": programmatic dependent launch between the gated-layer launches
  bool use_flags = true;         // "tile_flags": tile handshake between consecutive gated layers
  int tc_path = 1;               // "path": 1 = 16-bit activation planes (k_layer_h, round 2), 0 = fp32 rows (k_layer_tc / k_flow_tc, round 1)
  int tc_stagger = 0;            // "stagger": how far slot 1 starts behind slot 0 in k_flow_tc
  bool use_flow = true;          // "flow" (path 0): one persistent launch per flow (k_flow_tc) inside its job-size window
  bool tc_rotate = false;        // "rotate" (path 0): k_flow_tc tile-to-CTA assignment rotates from layer to layer
  int tc_seg = 0;                // "seg" (path 0): k_flow_tc gated layers per launch (0 = by job size)
  int tc_variant = PWV_TC_VARIANT_DEFAULT;   // "variant": 0 scalar epilogue arithmetic, 1 packed fp32x2, 2 (path 0) setmaxnreg register re-partition
  bool trace_flow = false;       // "trace_flow" (path 0): the phase trace follows k_flow_tc instead of forcing per-layer launches
  bool use_cp = false;           // "cp" (path 1): boxes -> TMEM by tcgen05.cp from the MMA issuer instead of the workers' ld.shared + tcgen05.st
  bool split1 = false;           // "split1" (path 1): GEMM1 starts on the x[t-d] half of K before the x[t] boxes are copied
  bool z_in_d = true;            // "z_in_d" (path 1, f16x3): gate output into the dead gate-accumulator columns, next tile copied while GEMM2 runs
  bool double_a = true;          // "double_a" (path 1, bf16): A operand double-buffered in TMEM, next tile copied while GEMM2 of this one runs
  bool split2 = false;           // "split2" (path 1): GEMM2 starts on the first half of the z chunks
  int trace_launch = -1;         // index of the gated layer to trace (0 .. total layers - 1, flows concatenated)
  int profiling = 0;             // 1: event pair around every gated-layer launch (serialised, no PDL);
                                 // 2: one pair around each flow's chain of gated-layer launches (as in production)
  std::vector<cudaEvent_t> ev;   // [0],[1] = whole forward; then the pairs
  int ev_used = 0;
  int prof_launches = 0;         // gated-layer launches covered by the pairs
};

static cudaEvent_t prof_event(pwv_model* m) {
  if (m->ev_used == (int)m->ev.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    m->ev.push_back(e);
  }
  return m->ev[m->ev_used++];
}
#define PWV_PROF_MARK(m, st)                                  \
  do {                                                        \
    if ((m)->profiling) {                                     \
      cudaEvent_t e_ = prof_event(m);                         \
      if (e_) cudaEventRecord(e_, st);                        \
    }                                                         \
  } while (0)

// ------------------------------------------------------------------------------------------------
// variable list (the reference's creation order; shapes in TF layout)
// ------------------------------------------------------------------------------------------------
static void add_var(pwv_model* m, const std::string& name, std::initializer_list<int64_t> shape) {
  VarSpec v;
  v.name = name;
  v.ndim = (int)shape.size();
  int i = 0;
  for (auto s : shape) v.shape[i++] = s;
  for (; i < 4; ++i) v.shape[i] = 1;
  m->var_index[name] = (int)m->vars.size();
  m->vars.push_back(v);
}

static void build_var_list(pwv_model* m) {
  const pwv_hparams& hp = m->hp;
  const int64_t k = hp.filter_width, R = hp.residual_channels, D = hp.dilation_channels,
                S = hp.skip_channels, Cc = hp.condition_channels;
  // instance_normalization creates beta before gamma (reference modules.py:279-280)
  auto norm = [&](const std::string& scope, int64_t channels, bool on) {
    if (!on) return;
    add_var(m, scope + "/beta", {channels});
    add_var(m, scope + "/gamma", {channels});
  };
  const bool nc = hp.normalize_cond == PWV_NORM_IN, nw = hp.normalize_wavenet == PWV_NORM_IN, nf = hp.normalize == PWV_NORM_IN;
  if (hp.cond_upsample == PWV_UPSAMPLE_TRANSPOSED_CONV) {
    // reference models.py:109-124: w_i [1, stride_i, Cc (out), Cin]; Cin = n_mels for stage 0, Cc afterwards
    int64_t cin = hp.n_mels;
    for (int i = 0; i < hp.n_upsample; ++i) {
      add_var(m, "iaf_vocoder/cond/transposed_conv_" + std::to_string(i) + "_weights", {1, hp.upsample_strides[i], Cc, cin});
      norm("iaf_vocoder/cond/normalize_transposed_conv_" + std::to_string(i), Cc, nc);
      cin = Cc;
    }
  } else if (hp.cond_upsample == PWV_UPSAMPLE_REPEAT) {
    add_var(m, "iaf_vocoder/cond/dense", {1, hp.n_mels, Cc});
  }
  const bool conditioned = hp.cond_upsample != PWV_UPSAMPLE_NONE;     // reference models.py:134-135 / modules.py:216
  norm("iaf_vocoder/cond/normalize/normalize", Cc, nc);
  for (int i = 0; i < hp.n_iaf; ++i) {
    for (int b = 0; b < 2; ++b) {
      std::string p = "iaf_vocoder/iaf" + std::to_string(i) + "/" + kBodies[b];
      add_var(m, p + "/causal_layer/filter", {k, 1, R});
      norm(p + "/causal_layer/normalize", R, nw);
      for (int j = 0; j < hp.n_layers[i]; ++j) {
        std::string q = p + "/dilated_stack/layer" + std::to_string(j);
        add_var(m, q + "/filter", {k, R, D});
        add_var(m, q + "/gate", {k, R, D});
        if (conditioned) {
          add_var(m, q + "/gc_filter", {1, Cc, D});
          add_var(m, q + "/gc_gate", {1, Cc, D});
        }
        if (hp.use_biases) {
          add_var(m, q + "/filter_bias", {D});
          add_var(m, q + "/gate_bias", {D});
        }
        norm(q + "/normalize_filter", D, nw);
        norm(q + "/normalize_gate", D, nw);
        add_var(m, q + "/dense", {1, D, R});
        add_var(m, q + "/skip", {1, D, S});
        if (hp.use_biases) {
          add_var(m, q + "/dense_bias", {R});
          add_var(m, q + "/skip_bias", {S});
        }
        norm(q + "/normalize_skip_output", S, nw);
        norm(q + "/normalize_dense_output", R, nw);
      }
      std::string q = p + "/postprocessing";
      norm(q + "/normalize_postprocess1", S, nw);
      add_var(m, q + "/postprocess1", {1, S, S});
      if (hp.use_biases) add_var(m, q + "/postprocess1_bias", {S});
      norm(q + "/normalize_postprocess2", S, nw);
      add_var(m, q + "/postprocess2", {1, S, 1});
      if (hp.use_biases) add_var(m, q + "/postprocess2_bias", {1});
    }
    norm("iaf_vocoder/normalize" + std::to_string(i), 1, nf);
  }
  m->staged.resize(m->vars.size());
  m->loaded.assign(m->vars.size(), 0);
}

// Opt the kernels this model launches into their dynamic shared-memory sizes on the model's device
// (per device, so it is done at finalize time with that device current).
template <int C>
static int configure_simt_kernels() {
  using Cfg = pwv::TileCfg<C>;
  PWV_CUDA(cudaFuncSetAttribute(pwv::k_cond_gemm<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_simt<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  PWV_CUDA(cudaFuncSetAttribute(pwv::k_post_simt<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  PWV_CUDA(cudaFuncSetAttribute(pwv::k_skip_simt<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  return PWV_OK;
}
static int configure_kernels(const pwv_model* m) {
  if (m->general) return PWV_OK;      // k_gen_* use static shared memory only
  int rc = m->C == 64 ? configure_simt_kernels<64>() : m->C == 128 ? configure_simt_kernels<128>() : configure_simt_kernels<256>();
  if (rc) return rc;
  if (m->hp.precision != PWV_PREC_FP32) {
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TC_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TC_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_tc<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TC_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_tc<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TC_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_flow_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TCF_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_flow_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TCF_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_tc<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TC_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_tc<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TC_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_flow_tc<true, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TCF_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_flow_tc<false, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TCF_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_h<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TH_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_h<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TH_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_h<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TH_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_h<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TH_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_h<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TH_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_h<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TH_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_h<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TH_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_h<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TH_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_h<true, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TH_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_layer_h<true, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TH_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_wide_h<false, pwv::TW_EPI_GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TW_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_wide_h<false, pwv::TW_EPI_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TW_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_wide_h<true, pwv::TW_EPI_GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TW_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_wide_h<true, pwv::TW_EPI_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TW_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_post_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TCP_SMEM_BYTES));
    PWV_CUDA(cudaFuncSetAttribute(pwv::k_post_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::TCP_SMEM_BYTES));
    if (m->tc.d_cond) {
      PWV_CUDA(cudaFuncSetAttribute(pwv::k_cbias_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::tcc_smem_bytes(m->Cc)));
      PWV_CUDA(cudaFuncSetAttribute(pwv::k_cbias_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pwv::tcc_smem_bytes(m->Cc)));
    }
  }
  return PWV_OK;
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int pwv_version(void) { return PWV_VERSION; }
const char* pwv_last_error(void) { return g_err; }

int pwv_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(PWV_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

int pwv_model_create(const pwv_hparams* hp, pwv_model** out) {
  if (!hp || !out) return fail(PWV_EINVAL, "null argument");
  *out = nullptr;
  if (hp->n_iaf < 1 || hp->n_iaf > PWV_MAX_FLOWS) return fail(PWV_EINVAL, "n_iaf=%d out of range [1,%d]", hp->n_iaf, PWV_MAX_FLOWS);
  if (hp->filter_width < 1 || hp->filter_width > PWV_MAX_FILTER_WIDTH)
    return fail(PWV_EINVAL, "filter_width=%d out of range [1,%d]", hp->filter_width, PWV_MAX_FILTER_WIDTH);
  if (hp->residual_channels < 1 || hp->dilation_channels < 1 || hp->skip_channels < 1)
    return fail(PWV_EINVAL, "residual/dilation/skip channels (%d/%d/%d) must be positive", hp->residual_channels, hp->dilation_channels, hp->skip_channels);
  const int C = hp->residual_channels;
  // The fused kernels cover R = D in {64, 128, 256}, S = 2R, filter_width 2 (the reference's defaults and BASELINE's
  // sweep); any other shape (reference modules.py:210-244 takes them as free parameters) runs the general fp32 chain.
  const bool general = hp->filter_width != 2 || C != hp->dilation_channels || hp->skip_channels != 2 * C || (C != 64 && C != 128 && C != 256);
  if (general && hp->precision != PWV_PREC_FP32)
    return fail(PWV_EINVAL, "filter_width=%d, residual/dilation/skip channels %d/%d/%d run on the general fp32 path only (precision fp32); "
                "the tensor-core kernels cover filter_width 2, R = D in {64,128,256}, S = 2R",
                hp->filter_width, C, hp->dilation_channels, hp->skip_channels);
  if (hp->condition_channels < 1 || hp->n_mels < 1 || hp->hop_length < 1)
    return fail(PWV_EINVAL, "bad condition_channels/n_mels/hop_length (%d/%d/%d)", hp->condition_channels, hp->n_mels, hp->hop_length);
  if (hp->cond_upsample != PWV_UPSAMPLE_REPEAT && hp->cond_upsample != PWV_UPSAMPLE_TRANSPOSED_CONV && hp->cond_upsample != PWV_UPSAMPLE_NONE)
    return fail(PWV_EINVAL, "unknown cond_upsample %d", hp->cond_upsample);
  if (hp->cond_upsample == PWV_UPSAMPLE_NONE && hp->normalize_cond)
    return fail(PWV_EINVAL, "normalize_cond needs a conditioned graph (cond_upsample PWV_UPSAMPLE_NONE leaves none to normalise)");
  if (hp->cond_upsample == PWV_UPSAMPLE_TRANSPOSED_CONV) {
    if (hp->n_upsample < 1 || hp->n_upsample > PWV_MAX_UPSAMPLE) return fail(PWV_EINVAL, "n_upsample=%d out of range [1,%d]", hp->n_upsample, PWV_MAX_UPSAMPLE);
    long long prod = 1;
    for (int i = 0; i < hp->n_upsample; ++i) {
      if (hp->upsample_strides[i] < 1) return fail(PWV_EINVAL, "upsample stride %d is %d", i, hp->upsample_strides[i]);
      prod *= hp->upsample_strides[i];
    }
    if (prod != hp->hop_length)   // reference models.py:106
      return fail(PWV_EINVAL, "product of the upsample strides (%lld) must equal hop_length (%d)", prod, hp->hop_length);
  }
  if (hp->precision != PWV_PREC_FP32 && hp->precision != PWV_PREC_F16X3 && hp->precision != PWV_PREC_BF16)
    return fail(PWV_EINVAL, "unknown precision %d", hp->precision);
  for (int v : {hp->normalize, hp->normalize_cond, hp->normalize_wavenet})
    if (v != PWV_NORM_NONE && v != PWV_NORM_IN) return fail(PWV_EINVAL, "unknown normaliser %d (PWV_NORM_NONE or PWV_NORM_IN)", v);
  if ((hp->normalize || hp->normalize_cond || hp->normalize_wavenet) && hp->precision != PWV_PREC_FP32)
    return fail(PWV_EINVAL, "the 'in' normalisers run on the fp32 path only (precision fp32): a statistic over the whole time axis sits between the stages of a layer");
  if (hp->precision != PWV_PREC_FP32 && C != 64 && hp->use_skip_connection)
    return fail(PWV_EINVAL, "use_skip_connection=True at %d channels runs on the fp32 path only (precision fp32)", C);
  int total = 0, mx = 0;
  for (int i = 0; i < hp->n_iaf; ++i) {
    if (hp->n_layers[i] < 1 || hp->n_layers[i] > PWV_MAX_LAYERS) return fail(PWV_EINVAL, "flow %d: %d layers out of range [1,%d]", i, hp->n_layers[i], PWV_MAX_LAYERS);
    for (int j = 0; j < hp->n_layers[i]; ++j)
      if (hp->dilations[i][j] < 1) return fail(PWV_EINVAL, "flow %d layer %d: dilation %d < 1", i, j, hp->dilations[i][j]);
    total += hp->n_layers[i];
    mx = hp->n_layers[i] > mx ? hp->n_layers[i] : mx;
  }
  pwv_model* m = new (std::nothrow) pwv_model();
  if (!m) return fail(PWV_ENOMEM, "out of host memory");
  m->hp = *hp;
  m->C = C;
  m->D = hp->dilation_channels;
  m->taps = hp->filter_width;
  m->general = general;
  m->S = hp->skip_channels;
  m->Cc = hp->condition_channels;
  m->total_layers = total;
  m->max_layers = mx;
  build_var_list(m);
  *out = m;
  return PWV_OK;
}

int pwv_model_destroy(pwv_model* m) {
  if (!m) return PWV_OK;
  if (m->d_arena) cudaFree(m->d_arena);
  if (m->h_dev) cudaFree(m->h_dev);
  for (auto e : m->ev) cudaEventDestroy(e);
  pwv::tc_model_free(m->tc);
  pwv::tw_model_free(m->tw);
  delete m;
  return PWV_OK;
}

int pwv_model_num_variables(const pwv_model* m) {
  if (!m) return fail(PWV_EINVAL, "null model");
  return (int)m->vars.size();
}

int pwv_model_variable(const pwv_model* m, int index, const char** name, int64_t shape[4], int* ndim) {
  if (!m || index < 0 || index >= (int)m->vars.size()) return fail(PWV_EINVAL, "bad model/index");
  const VarSpec& v = m->vars[index];
  if (name) *name = v.name.c_str();
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = v.shape[i];
  if (ndim) *ndim = v.ndim;
  return PWV_OK;
}

int pwv_model_load_weight(pwv_model* m, const char* tf_name, const float* host_data, const int64_t* shape, int ndim) {
  if (!m || !tf_name || !host_data || !shape) return fail(PWV_EINVAL, "null argument");
  auto it = m->var_index.find(tf_name);
  if (it == m->var_index.end()) return fail(PWV_ENAME, "unknown variable '%s'", tf_name);
  const VarSpec& v = m->vars[it->second];
  if (ndim != v.ndim) return fail(PWV_ENAME, "%s: ndim %d, expected %d", tf_name, ndim, v.ndim);
  for (int i = 0; i < ndim; ++i)
    if (shape[i] != v.shape[i]) return fail(PWV_ENAME, "%s: dim %d is %lld, expected %lld", tf_name, i, (long long)shape[i], (long long)v.shape[i]);
  m->staged[it->second].assign(host_data, host_data + v.numel());
  m->loaded[it->second] = 1;
  m->finalized = false;
  return PWV_OK;
}

static const std::vector<float>& var(const pwv_model* m, const std::string& name) {
  return m->staged[m->var_index.at(name)];
}

int pwv_model_finalize(pwv_model* m) {
  if (!m) return fail(PWV_EINVAL, "null model");
  for (size_t i = 0; i < m->vars.size(); ++i)
    if (!m->loaded[i]) return fail(PWV_ESTATE, "variable '%s' was not loaded", m->vars[i].name.c_str());
  const pwv_hparams& hp = m->hp;
  const int C = m->C, S = m->S, Cc = m->Cc;
  const int R = m->C, D = m->D, FW = m->taps;      // (R = D = C, FW = 2 on the fused paths)
  std::vector<float> arena;
  auto put = [&](size_t n) {   // 16-float (64 B) aligned blocks
    size_t off = (arena.size() + 15) / 16 * 16;
    arena.resize(off + n, 0.f);
    return off;
  };
  if (hp.cond_upsample == PWV_UPSAMPLE_TRANSPOSED_CONV) {
    // stage i as a row GEMM: out[row][j*Cc + co] = relu(sum_ci in[row][ci] * w[0][j][co][ci]), i.e. B[ci][j*Cc + co];
    // the output viewed as [rows * stride][Cc] is the upsampled sequence (kernel width == stride: no overlap)
    int cin = hp.n_mels;
    for (int i = 0; i < hp.n_upsample; ++i) {
      const int sN = hp.upsample_strides[i];
      m->off_wup[i] = put((size_t)cin * sN * Cc);
      const auto& w = var(m, "iaf_vocoder/cond/transposed_conv_" + std::to_string(i) + "_weights");
      for (int j = 0; j < sN; ++j)
        for (int co = 0; co < Cc; ++co)
          for (int ci = 0; ci < cin; ++ci)
            arena[m->off_wup[i] + (size_t)ci * sN * Cc + (size_t)j * Cc + co] = w[((size_t)j * Cc + co) * cin + ci];
      cin = Cc;
    }
  } else {
    m->off_wc = put((size_t)hp.n_mels * Cc);      // (zeros in an unconditional graph)
    if (hp.cond_upsample == PWV_UPSAMPLE_REPEAT) {
      const auto& w = var(m, "iaf_vocoder/cond/dense");
      std::copy(w.begin(), w.end(), arena.begin() + m->off_wc);
    }
  }
  const bool conditioned = hp.cond_upsample != PWV_UPSAMPLE_NONE;
  m->off_colscale = put(2 * C);
  for (int c = 0; c < C; ++c) {
    arena[m->off_colscale + c] = pwv::TC_KF;
    arena[m->off_colscale + C + c] = pwv::TC_KG;
  }
  auto put_norm = [&](const std::string& scope, size_t n) {     // gamma then beta, n floats each
    NormOff o;
    o.gamma = put(n);
    o.beta = put(n);
    const auto& g = var(m, scope + "/gamma");
    const auto& b = var(m, scope + "/beta");
    std::copy(g.begin(), g.end(), arena.begin() + o.gamma);
    std::copy(b.begin(), b.end(), arena.begin() + o.beta);
    return o;
  };
  const bool nrm_c = hp.normalize_cond == PWV_NORM_IN, nrm_w = hp.normalize_wavenet == PWV_NORM_IN, nrm_f = hp.normalize == PWV_NORM_IN;
  if (nrm_c) {
    m->n_cond = put_norm("iaf_vocoder/cond/normalize/normalize", Cc);
    if (hp.cond_upsample == PWV_UPSAMPLE_TRANSPOSED_CONV)
      for (int i = 0; i < hp.n_upsample; ++i) m->n_up[i] = put_norm("iaf_vocoder/cond/normalize_transposed_conv_" + std::to_string(i), Cc);
  }
  m->n_flow.assign(hp.n_iaf, NormOff());
  if (nrm_f)
    for (int i = 0; i < hp.n_iaf; ++i) m->n_flow[i] = put_norm("iaf_vocoder/normalize" + std::to_string(i), 1);
  m->bodies.assign(hp.n_iaf * 2, BodyOff());
  m->off_wgc.assign(hp.n_iaf, 0);
  m->off_bfg.assign(hp.n_iaf, 0);
  for (int i = 0; i < hp.n_iaf; ++i) {
    m->off_wgc[i] = put((size_t)2 * hp.n_layers[i] * Cc * 2 * D);
    m->off_bfg[i] = put((size_t)2 * hp.n_layers[i] * 2 * D);
    for (int b = 0; b < 2; ++b) {
      BodyOff& bo = m->bodies[i * 2 + b];
      std::string p = "iaf_vocoder/iaf" + std::to_string(i) + "/" + kBodies[b];
      bo.causal = put((size_t)FW * R);
      {
        const auto& w = var(m, p + "/causal_layer/filter");   // [FW][1][R]
        std::copy(w.begin(), w.end(), arena.begin() + bo.causal);
      }
      if (nrm_w) bo.n_causal = put_norm(p + "/causal_layer/normalize", C);
      bo.layers.resize(hp.n_layers[i]);
      for (int j = 0; j < hp.n_layers[i]; ++j) {
        LayerOff& lo = bo.layers[j];
        std::string q = p + "/dilated_stack/layer" + std::to_string(j);
        if (nrm_w) {
          lo.n_fg.gamma = put(2 * D);          // [filter | gate] halves side by side, like the pre-activation rows
          lo.n_fg.beta = put(2 * D);
          const char* part[2] = {"/normalize_filter", "/normalize_gate"};
          for (int h2 = 0; h2 < 2; ++h2) {
            const auto& g = var(m, q + part[h2] + "/gamma");
            const auto& bt = var(m, q + part[h2] + "/beta");
            std::copy(g.begin(), g.end(), arena.begin() + lo.n_fg.gamma + h2 * D);
            std::copy(bt.begin(), bt.end(), arena.begin() + lo.n_fg.beta + h2 * D);
          }
          lo.n_skip = put_norm(q + "/normalize_skip_output", S);
          lo.n_dense = put_norm(q + "/normalize_dense_output", C);
        }
        const auto& wf = var(m, q + "/filter");      // [FW][R][D]
        const auto& wg = var(m, q + "/gate");
        lo.wfg = put((size_t)FW * R * 2 * D);
        for (int tap = 0; tap < FW; ++tap)
          for (int ci = 0; ci < R; ++ci)
            for (int co = 0; co < D; ++co) {
              size_t row = (size_t)tap * R + ci;   // tap k multiplies x[t - (FW-1-k) d]: FW = 2: tap 0 x[t-d], tap 1 x[t]
              arena[lo.wfg + row * 2 * D + co] = wf[((size_t)tap * R + ci) * D + co];
              arena[lo.wfg + row * 2 * D + D + co] = wg[((size_t)tap * R + ci) * D + co];
            }
        // (unconditional graph: the conditioning projections stay zero, so the per-frame conditioning rows the
        //  kernels add are just the filter / gate biases)
        const size_t wgc = m->off_wgc[i] + ((size_t)b * hp.n_layers[i] + j) * Cc * 2 * D;
        const size_t bfg = m->off_bfg[i] + ((size_t)b * hp.n_layers[i] + j) * 2 * D;
        if (conditioned) {
          const auto& gf = var(m, q + "/gc_filter");   // [1][Cc][D]
          const auto& gg = var(m, q + "/gc_gate");
          for (int c = 0; c < Cc; ++c)
            for (int co = 0; co < D; ++co) {
              arena[wgc + (size_t)c * 2 * D + co] = gf[(size_t)c * D + co];
              arena[wgc + (size_t)c * 2 * D + D + co] = gg[(size_t)c * D + co];
            }
        }
        lo.wd = put((size_t)D * R);
        lo.bd = put(R);
        lo.ws = put((size_t)D * S);
        lo.bs = put(S);
        {
          const auto& w = var(m, q + "/dense");
          std::copy(w.begin(), w.end(), arena.begin() + lo.wd);
          const auto& s = var(m, q + "/skip");
          std::copy(s.begin(), s.end(), arena.begin() + lo.ws);
        }
        if (hp.use_biases) {
          const auto& bf = var(m, q + "/filter_bias");
          const auto& bg = var(m, q + "/gate_bias");
          std::copy(bf.begin(), bf.end(), arena.begin() + bfg);
          std::copy(bg.begin(), bg.end(), arena.begin() + bfg + D);
          const auto& bd = var(m, q + "/dense_bias");
          std::copy(bd.begin(), bd.end(), arena.begin() + lo.bd);
          const auto& bs = var(m, q + "/skip_bias");
          std::copy(bs.begin(), bs.end(), arena.begin() + lo.bs);
        }
      }
      std::string q = p + "/postprocessing";
      if (nrm_w) {
        bo.n_pp1 = put_norm(q + "/normalize_postprocess1", S);
        bo.n_pp2 = put_norm(q + "/normalize_postprocess2", S);
      }
      bo.w1 = put((size_t)S * S);
      bo.b1 = put(S);
      bo.w2 = put(S);
      bo.b2 = put(1);
      {
        const auto& w1 = var(m, q + "/postprocess1");
        std::copy(w1.begin(), w1.end(), arena.begin() + bo.w1);
        const auto& w2 = var(m, q + "/postprocess2");
        std::copy(w2.begin(), w2.end(), arena.begin() + bo.w2);
        if (hp.use_biases) {
          const auto& b1 = var(m, q + "/postprocess1_bias");
          std::copy(b1.begin(), b1.end(), arena.begin() + bo.b1);
          const auto& b2 = var(m, q + "/postprocess2_bias");
          std::copy(b2.begin(), b2.end(), arena.begin() + bo.b2);
        }
      }
    }
  }
  PWV_CUDA(cudaGetDevice(&m->device));
  PWV_CUDA(cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, m->device));
  {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) m->dev_total_bytes = total_b;
  }
  if (m->d_arena) { cudaFree(m->d_arena); m->d_arena = nullptr; }
  m->arena_floats = arena.size();
  PWV_CUDA(cudaMalloc(&m->d_arena, arena.size() * sizeof(float)));
  PWV_CUDA(cudaMemcpy(m->d_arena, arena.data(), arena.size() * sizeof(float), cudaMemcpyHostToDevice));

  if (hp.precision != PWV_PREC_FP32) {
    // tensor-core operand images are built from the same packed fp32 arena
    std::vector<pwv::TcLayerSrc> src;
    for (int i = 0; i < hp.n_iaf; ++i)
      for (int b = 0; b < 2; ++b)
        for (int j = 0; j < hp.n_layers[i]; ++j) {
          const LayerOff& lo = m->bodies[i * 2 + b].layers[j];
          pwv::TcLayerSrc s;
          s.wfg = arena.data() + lo.wfg;
          s.wd = arena.data() + lo.wd;
          s.bd = arena.data() + lo.bd;
          src.push_back(s);
        }
    std::vector<pwv::TcPostSrc> posts;
    for (int i = 0; i < hp.n_iaf; ++i)
      for (int b = 0; b < 2; ++b) {
        const BodyOff& bo = m->bodies[i * 2 + b];
        const LayerOff& lo = bo.layers[hp.n_layers[i] - 1];
        pwv::TcPostSrc q;
        q.ws = arena.data() + lo.ws; q.bs = arena.data() + lo.bs;
        q.w1 = arena.data() + bo.w1; q.b1 = arena.data() + bo.b1;
        q.w2 = arena.data() + bo.w2; q.b2 = arena.data() + bo.b2;
        posts.push_back(q);
      }
    if (C != pwv::TC_C) {       // wide channel counts: streamed-K kernels (pwv_tc3.cuh); conditioning and post-net stay on the fp32 kernels
      const char* werr = pwv::tw_model_build(m->tw, hp.precision, C, src);
      if (werr) return fail(PWV_ECUDA, "wide tensor-core weight streams: %s", werr);
      const int rcw = configure_kernels(m);
      if (rcw) return rcw;
      m->finalized = true;
      return PWV_OK;
    }
    const char* err = pwv::tc_model_build(m->tc, hp.precision, C, src, posts);
    if (err) return fail(PWV_ECUDA, "tensor-core weight images: %s", err);
    // conditioning projections on the tensor cores too (when Cc allows it)
    std::vector<pwv::TcCondSrc> csrc;
    for (int i = 0; i < hp.n_iaf; ++i)
      for (int e = 0; e < 2 * hp.n_layers[i]; ++e) {
        pwv::TcCondSrc q;
        q.wgc = arena.data() + m->off_wgc[i] + (size_t)e * Cc * 2 * C;
        q.bias = arena.data() + m->off_bfg[i] + (size_t)e * 2 * C;
        csrc.push_back(q);
      }
    err = pwv::tc_cond_build(m->tc, hp.precision, Cc, arena.data() + m->off_colscale, csrc);
    if (err) return fail(PWV_ECUDA, "tensor-core conditioning images: %s", err);
  }
  {
    const int rc = configure_kernels(m);
    if (rc) return rc;
  }
  m->finalized = true;
  return PWV_OK;
}

// ------------------------------------------------------------------------------------------------
// workspace carving (all blocks 256-byte aligned)
// ------------------------------------------------------------------------------------------------
struct Workspace {
  float* cproj;     // [N][cond_rows][Cc]   cond_rows = t_mel ('repeat') or T ('transposed_conv')
  float* cbias;     // [2][Lmax][N][cond_rows][2C]
  float* up[2];     // transposed_conv: stage outputs (ping/pong)
  float* act[2];    // ping/pong, each [2][N][T][C]
  float* ss;        // [2][N][T] scale, shift
  float* x[2];      // [N][T] ping/pong
  float* zbuf;      // use_skip_connection: [2][N][T][C] gate output of the current layer
  float* skip;      // use_skip_connection: [2][N][T][2C] running sum of the skip outputs
  float* fg;        // normalize_wavenet: [2][N][T][2C] pre-activations of the current layer
  float* hbuf;      // normalize_wavenet: [2][N][T][2C] post-net hidden layer
  float* total;     // normalize_wavenet + use_skip_connection: [2][N][T][2C] running sum of the normalised skip outputs
  float2* stats;    // any normaliser: (mean, sqrt(var + eps)) per (utterance-body, channel)
  int* flags;       // tensor-core path: [total gated layers][2][N * ceil(T/128)] per-tile "output stored" flags
  size_t flags_bytes;
  size_t bytes;
};

// Conditioning rows per utterance and samples per row: 'repeat' keeps the per-layer conditioning terms at mel
// rate (row = frame (s + hop/2) / hop); 'transposed_conv' produces a different vector for every sample.
// (a normalised conditioning is a per-sample tensor too: its statistics run over the repeated, cropped frames)
static bool cond_full_rate(const pwv_model* m) {
  return m->hp.cond_upsample == PWV_UPSAMPLE_TRANSPOSED_CONV || m->hp.normalize_cond == PWV_NORM_IN;
}
static int cond_rows(const pwv_model* m, int T) { return cond_full_rate(m) ? T : 1 + T / m->hp.hop_length; }
static int cond_hop(const pwv_model* m) { return cond_full_rate(m) ? 1 : m->hp.hop_length; }

static void carve(const pwv_model* m, int N, int T, char* base, Workspace* w) {
  const int C = m->C, D = m->D, t_mel = 1 + T / m->hp.hop_length;      // (D = C on the fused paths)
  const bool tconv = m->hp.cond_upsample == PWV_UPSAMPLE_TRANSPOSED_CONV;
  const size_t crows = (size_t)cond_rows(m, T);               // conditioning rows per utterance
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += (bytes + 255) / 256 * 256;
    return p;
  };
  w->cproj = (float*)take(sizeof(float) * (size_t)N * crows * m->Cc);
  w->cbias = (float*)take(sizeof(float) * (size_t)2 * m->max_layers * N * crows * 2 * D);
  w->up[0] = w->up[1] = nullptr;
  if (tconv) {      // ping/pong stage outputs; the last stage has N * t_mel * hop rows
    size_t rows = (size_t)N * t_mel, big[2] = {0, 0};
    for (int i = 0; i < m->hp.n_upsample; ++i) {
      rows *= m->hp.upsample_strides[i];
      if (rows > big[i & 1]) big[i & 1] = rows;
    }
    w->up[0] = (float*)take(sizeof(float) * big[0] * m->Cc);
    w->up[1] = (float*)take(sizeof(float) * (big[1] ? big[1] : 1) * m->Cc);
  } else if (m->hp.normalize_cond == PWV_NORM_IN) {
    w->up[0] = (float*)take(sizeof(float) * (size_t)N * t_mel * m->Cc);   // mel-rate rows before the repeat is materialised
  }
  w->act[0] = (float*)take(sizeof(float) * (size_t)2 * N * T * C);
  w->act[1] = (float*)take(sizeof(float) * (size_t)2 * N * T * C);
  w->ss = (float*)take(sizeof(float) * (size_t)2 * N * T);
  w->x[0] = (float*)take(sizeof(float) * (size_t)N * T);
  w->x[1] = (float*)take(sizeof(float) * (size_t)N * T);
  w->zbuf = w->skip = w->fg = w->hbuf = w->total = nullptr;
  w->stats = nullptr;
  if (m->hp.normalize_wavenet == PWV_NORM_IN || m->general) {           // the un-fused chains keep every stage of a layer
    const size_t S = (size_t)m->S;
    w->zbuf = (float*)take(sizeof(float) * (size_t)2 * N * T * D);
    w->skip = (float*)take(sizeof(float) * (size_t)2 * N * T * S);
    w->fg = (float*)take(sizeof(float) * (size_t)2 * N * T * 2 * D);
    w->hbuf = (float*)take(sizeof(float) * (size_t)2 * N * T * S);
    if (m->hp.use_skip_connection && m->hp.normalize_wavenet == PWV_NORM_IN) w->total = (float*)take(sizeof(float) * (size_t)2 * N * T * S);
  } else if (m->hp.use_skip_connection) {
    w->zbuf = (float*)take(sizeof(float) * (size_t)2 * N * T * C);
    w->skip = (float*)take(sizeof(float) * (size_t)2 * N * T * 2 * C);
  } else if (m->hp.precision != PWV_PREC_FP32 && C != pwv::TC_C) {
    w->zbuf = (float*)take(sizeof(float) * (size_t)2 * N * T * C);     // z planes between the gate and the dense pass (k_wide_h)
  }
  if (m->hp.normalize || m->hp.normalize_cond || m->hp.normalize_wavenet) {
    size_t widest = (size_t)2 * N * (2 * D > m->S ? 2 * D : m->S);
    if ((size_t)2 * N * C > widest) widest = (size_t)2 * N * C;
    if ((size_t)N * m->Cc > widest) widest = (size_t)N * m->Cc;
    w->stats = (float2*)take(sizeof(float2) * widest);
  }
  w->flags = nullptr;
  w->flags_bytes = 0;
  if (m->hp.precision != PWV_PREC_FP32) {
    const size_t tiles_body = (size_t)N * ((T + pwv::TC_TM - 1) / pwv::TC_TM);
    w->flags_bytes = sizeof(int) * ((size_t)m->total_layers * 2 * tiles_body + m->total_layers);   // + a completion counter per layer
    w->flags = (int*)take(w->flags_bytes);
  }
  w->bytes = off;
}

static int check_shape(const pwv_model* m, int N, int T) {
  if (!m) return fail(PWV_EINVAL, "null model");
  if (N < 1 || T < 1) return fail(PWV_EINVAL, "N=%d, T=%d must be positive", N, T);
  if (T % m->hp.hop_length != 0)
    return fail(PWV_EINVAL, "length %d is not a multiple of hop_length %d (the reference's cond crop needs it, models.py:131-133)", T, m->hp.hop_length);
  if (N > 65535) return fail(PWV_EINVAL, "N=%d exceeds the grid limit 65535", N);
  return PWV_OK;
}

int pwv_workspace_bytes(const pwv_model* m, int N, int T, size_t* bytes) {
  int rc = check_shape(m, N, T);
  if (rc) return rc;
  if (!bytes) return fail(PWV_EINVAL, "null bytes");
  Workspace w;
  carve(m, N, T, nullptr, &w);
  *bytes = w.bytes;
  // A full-rate conditioning (cond_upsample_method 'transposed_conv', or normalize_cond) materialises the per-layer
  // conditioning terms for every SAMPLE: [2][L][N][T][2C] floats. Say so instead of letting the caller's allocator fail.
  // (the device's size was read once at finalize: with 8 ranks calling at the same moment a cudaMemGetInfo per call
  //  measured as milliseconds per pwv_forward_host)
  const size_t total_b = m->dev_total_bytes;
  if (total_b && w.bytes > total_b)
    return fail(PWV_ENOMEM, "workspace for N=%d, T=%d is %zu bytes, the device has %zu%s", N, T, w.bytes, total_b,
                cond_full_rate(m) ? " (full-rate conditioning: 2 x layers x N x T x 2C floats of per-layer conditioning terms; split the batch)" : "");
  return PWV_OK;
}

}  // extern "C"

template <int C>
static int launch_cond_gemm_c(const float* A, const pwv::RowGemmBatch& rb, int M, int K, int Z, cudaStream_t st) {
  using Cfg = pwv::TileCfg<C>;
  dim3 grid((M + Cfg::TM - 1) / Cfg::TM, 1, Z);
  pwv::k_cond_gemm<C><<<grid, Cfg::NT, Cfg::SMEM, st>>>(A, rb, M, K);
  return PWV_OK;
}
static int launch_cond_gemm(int C, const float* A, const pwv::RowGemmBatch& rb, int M, int K, int Z, cudaStream_t st) {
  if (C == 64) return launch_cond_gemm_c<64>(A, rb, M, K, Z, st);
  if (C == 128) return launch_cond_gemm_c<128>(A, rb, M, K, Z, st);
  return launch_cond_gemm_c<256>(A, rb, M, K, Z, st);
}


// x [UB][T][Cn] normalised in place over time (pwv_norm.cuh); groups of `ub_per_group` utterance-bodies use (g0, b0), the rest (g1, b1)
static int launch_in_norm(const pwv_model* m, const Workspace& w, float* x, int UB, int T, int Cn, const NormOff& a, const NormOff& b,
                          int ub_per_group, bool pre_relu, cudaStream_t st, int* launches) {
  const dim3 sgrid((Cn + 31) / 32, UB);
  const size_t total = (size_t)UB * T * Cn;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  const float *g0 = m->d_arena + a.gamma, *b0 = m->d_arena + a.beta, *g1 = m->d_arena + b.gamma, *b1 = m->d_arena + b.beta;
  if (pre_relu) {
    pwv::k_in_stats<true><<<sgrid, 256, 0, st>>>(x, w.stats, T, Cn);
    pwv::k_in_apply<true><<<blocks, 256, 0, st>>>(x, w.stats, g0, b0, g1, b1, ub_per_group, (size_t)T * Cn, Cn, total);
  } else {
    pwv::k_in_stats<false><<<sgrid, 256, 0, st>>>(x, w.stats, T, Cn);
    pwv::k_in_apply<false><<<blocks, 256, 0, st>>>(x, w.stats, g0, b0, g1, b1, ub_per_group, (size_t)T * Cn, Cn, total);
  }
  *launches += 2;
  PWV_CUDA(cudaGetLastError());
  return PWV_OK;
}

// One flow's WaveNet bodies with normalize_wavenet = 'in' (reference modules.py:129-259 with self.normalize set): every
// normalised tensor is materialised, so a layer is pre-activations -> norm -> gate + dense (+ z) -> norm (-> skip -> norm).
template <int C>
static int launch_layers_in(pwv_model* m, const Workspace& w, int flow, int N, int T, cudaStream_t st, const pwv_taps* taps,
                            int* cur_buf, int* launches) {
  using Cfg = pwv::TileCfg<C>;
  const pwv_hparams& hp = m->hp;
  const int L = hp.n_layers[flow], t_mel = cond_rows(m, T), c_hop = cond_hop(m), S = m->S;
  const dim3 grid((T + Cfg::TM - 1) / Cfg::TM, N, 2);
  const size_t rows = (size_t)N * T;
  const BodyOff& b0 = m->bodies[flow * 2 + 0];
  const BodyOff& b1 = m->bodies[flow * 2 + 1];
  int cur = *cur_buf, rc;
  rc = launch_in_norm(m, w, w.act[cur], 2 * N, T, C, b0.n_causal, b1.n_causal, N, false, st, launches);   // modules.py:181-182
  if (rc) return rc;
  for (int j = 0; j < L; ++j) {
    const bool last = j == L - 1;
    const LayerOff& l0 = b0.layers[j];
    const LayerOff& l1 = b1.layers[j];
    pwv::LayerParams p;
    p.x_in = w.act[cur];
    p.x_out = w.act[cur ^ 1];
    for (int b = 0; b < 2; ++b) {
      const LayerOff& lo = m->bodies[flow * 2 + b].layers[j];
      p.wfg[b] = m->d_arena + lo.wfg;
      p.wd[b] = m->d_arena + lo.wd;
      p.bd[b] = m->d_arena + lo.bd;
      p.cbias[b] = w.cbias + ((size_t)b * L + j) * N * t_mel * 2 * C;
    }
    p.N = N; p.T = T; p.t_mel = t_mel; p.hop = c_hop; p.dilation = hp.dilations[flow][j];
    p.z_out = w.zbuf; p.fg_out = w.fg; p.fg_in = w.fg;
    PWV_PROF_MARK(m, st);
    p.mode = 3;                                                                   // [f|g] incl. conditioning and biases
    pwv::k_layer_simt<C><<<grid, Cfg::NT, Cfg::SMEM, st>>>(p);
    ++*launches;
    rc = launch_in_norm(m, w, w.fg, 2 * N, T, 2 * C, l0.n_fg, l1.n_fg, N, false, st, launches);            // modules.py:230-234
    if (rc) return rc;
    p.mode = 4;                                                                   // gate, dense 1x1, residual; z kept for the skip
    pwv::k_layer_simt<C><<<grid, Cfg::NT, Cfg::SMEM, st>>>(p);
    ++*launches;
    cur ^= 1;
    rc = launch_in_norm(m, w, w.act[cur], 2 * N, T, C, l0.n_dense, l1.n_dense, N, false, st, launches);    // modules.py:256-257
    if (rc) return rc;
    if (hp.use_skip_connection || last) {                                          // skip_output (dead otherwise), modules.py:243-255
      pwv::SkipParams sp;
      sp.z = w.zbuf;
      sp.ws[0] = m->d_arena + l0.ws; sp.bs[0] = m->d_arena + l0.bs;
      sp.ws[1] = m->d_arena + l1.ws; sp.bs[1] = m->d_arena + l1.bs;
      sp.skip_sum = w.skip; sp.N = N; sp.T = T; sp.first = 1;
      pwv::k_skip_simt<C><<<grid, Cfg::NT, Cfg::SMEM, st>>>(sp);
      ++*launches;
      rc = launch_in_norm(m, w, w.skip, 2 * N, T, S, l0.n_skip, l1.n_skip, N, false, st, launches);
      if (rc) return rc;
      if (hp.use_skip_connection) {                                                // sum(outputs), modules.py:147
        const size_t n = 2 * rows * S;
        if (j == 0) PWV_CUDA(cudaMemcpyAsync(w.total, w.skip, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
        else { pwv::k_add_inplace<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.total, w.skip, n); ++*launches; }
      }
    }
    PWV_PROF_MARK(m, st);
    if (m->profiling) ++m->prof_launches;
    if (taps && taps->layer_out && taps->layer_flow == flow && taps->layer_index == j && (taps->layer_body == 0 || taps->layer_body == 1))
      PWV_CUDA(cudaMemcpyAsync(taps->layer_out, w.act[cur] + (size_t)taps->layer_body * rows * C, sizeof(float) * rows * C, cudaMemcpyDeviceToDevice, st));
  }
  // post-net: relu -> norm -> 1x1 + bias -> relu -> norm -> 1x1 + bias (modules.py:148-165)
  float* src = hp.use_skip_connection ? w.total : w.skip;
  rc = launch_in_norm(m, w, src, 2 * N, T, S, b0.n_pp1, b1.n_pp1, N, true, st, launches);
  if (rc) return rc;
  for (int b = 0; b < 2; ++b) {
    const BodyOff& bo = m->bodies[flow * 2 + b];
    pwv::RowGemmBatch rb{m->d_arena + bo.w1, 0, m->d_arena + bo.b1, 0, w.hbuf + (size_t)b * rows * S, 0, nullptr};
    dim3 g((S + 63) / 64, (unsigned)((rows + 63) / 64), 1);
    pwv::k_row_gemm<true><<<g, 256, 0, st>>>(src + (size_t)b * rows * S, rb, (int)rows, S, S);
    ++*launches;
  }
  rc = launch_in_norm(m, w, w.hbuf, 2 * N, T, S, b0.n_pp2, b1.n_pp2, N, false, st, launches);
  if (rc) return rc;
  for (int b = 0; b < 2; ++b) {
    const BodyOff& bo = m->bodies[flow * 2 + b];
    pwv::RowGemmBatch rb{m->d_arena + bo.w2, 0, m->d_arena + bo.b2, 0, w.ss + (size_t)b * rows, 0, nullptr};
    dim3 g(1, (unsigned)((rows + 63) / 64), 1);
    pwv::k_row_gemm<false><<<g, 256, 0, st>>>(w.hbuf + (size_t)b * rows * S, rb, (int)rows, S, 1);
    ++*launches;
  }
  *cur_buf = cur;
  PWV_CUDA(cudaGetLastError());
  return PWV_OK;
}

// The general-shape chain (pwv_gen.cuh): any R / D / S / filter_width, with or without the 'in' normalisers and the
// skip sum; one GEMM launch per stage covers both bodies. Reference modules.py:129-259, stage by stage.
static void gen_gemm(const pwv::GenGemm& g, cudaStream_t st, int* launches) {
  const dim3 grid((g.Nc + 63) / 64, (g.M + 63) / 64, 2);
  pwv::k_gen_gemm<<<grid, 256, 0, st>>>(g);
  ++*launches;
}
static int launch_layers_gen(pwv_model* m, const Workspace& w, int flow, int N, int T, cudaStream_t st, const pwv_taps* taps,
                             int* cur_buf, int* launches) {
  const pwv_hparams& hp = m->hp;
  const int L = hp.n_layers[flow], crows = cond_rows(m, T), c_hop = cond_hop(m);
  const int R = m->C, D = m->D, S = m->S, FW = m->taps;
  const size_t rows = (size_t)N * T;
  const bool nrm = hp.normalize_wavenet == PWV_NORM_IN;
  const BodyOff& b0 = m->bodies[flow * 2 + 0];
  const BodyOff& b1 = m->bodies[flow * 2 + 1];
  int cur = *cur_buf, rc;
  pwv::GenGemm base;
  memset(&base, 0, sizeof(base));
  base.lda = 1; base.taps = 1; base.dilation = 1; base.T = T; base.crows = crows; base.hop = c_hop; base.M = (int)rows;
  if (nrm) {
    rc = launch_in_norm(m, w, w.act[cur], 2 * N, T, R, b0.n_causal, b1.n_causal, N, false, st, launches);    // modules.py:181-182
    if (rc) return rc;
  }
  for (int j = 0; j < L; ++j) {
    const bool last = j == L - 1;
    const LayerOff* lo[2] = {&b0.layers[j], &b1.layers[j]};
    PWV_PROF_MARK(m, st);
    {   // [filter | gate] pre-activations: causal convs over the FW taps + conditioning + biases (modules.py:210-228)
      pwv::GenGemm g = base;
      for (int b = 0; b < 2; ++b) {
        g.A[b] = w.act[cur] + (size_t)b * rows * R;
        g.B[b] = m->d_arena + lo[b]->wfg;
        g.cond[b] = w.cbias + ((size_t)b * L + j) * N * crows * 2 * D;      // (carries filter_bias | gate_bias)
        g.out[b] = w.fg + (size_t)b * rows * 2 * D;
      }
      g.lda = R; g.taps = FW; g.dilation = hp.dilations[flow][j]; g.K = FW * R; g.Nc = 2 * D;
      gen_gemm(g, st, launches);
    }
    if (nrm) {
      rc = launch_in_norm(m, w, w.fg, 2 * N, T, 2 * D, lo[0]->n_fg, lo[1]->n_fg, N, false, st, launches);       // modules.py:230-234
      if (rc) return rc;
    }
    {
      const size_t n = 2 * rows * D;
      pwv::k_gen_gate<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.fg, w.zbuf, D, n);                         // modules.py:236
      ++*launches;
    }
    {   // dense 1x1 + bias + residual (modules.py:239-251)
      pwv::GenGemm g = base;
      for (int b = 0; b < 2; ++b) {
        g.A[b] = w.zbuf + (size_t)b * rows * D;
        g.B[b] = m->d_arena + lo[b]->wd;
        g.bias[b] = m->d_arena + lo[b]->bd;
        g.resid[b] = w.act[cur] + (size_t)b * rows * R;
        g.out[b] = w.act[cur ^ 1] + (size_t)b * rows * R;
      }
      g.lda = D; g.K = D; g.Nc = R;
      gen_gemm(g, st, launches);
    }
    cur ^= 1;
    if (nrm) {
      rc = launch_in_norm(m, w, w.act[cur], 2 * N, T, R, lo[0]->n_dense, lo[1]->n_dense, N, false, st, launches);   // modules.py:256-257
      if (rc) return rc;
    }
    if (hp.use_skip_connection || last) {       // skip_output: summed over the layers, or the last one alone (modules.py:147,243-255)
      pwv::GenGemm g = base;
      for (int b = 0; b < 2; ++b) {
        g.A[b] = w.zbuf + (size_t)b * rows * D;
        g.B[b] = m->d_arena + lo[b]->ws;
        g.bias[b] = m->d_arena + lo[b]->bs;
        g.out[b] = w.skip + (size_t)b * rows * S;
      }
      g.lda = D; g.K = D; g.Nc = S;
      g.accumulate = (!nrm && hp.use_skip_connection && j > 0) ? 1 : 0;
      gen_gemm(g, st, launches);
      if (nrm) {
        rc = launch_in_norm(m, w, w.skip, 2 * N, T, S, lo[0]->n_skip, lo[1]->n_skip, N, false, st, launches);
        if (rc) return rc;
        if (hp.use_skip_connection) {
          const size_t n = 2 * rows * S;
          if (j == 0) PWV_CUDA(cudaMemcpyAsync(w.total, w.skip, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
          else { pwv::k_add_inplace<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.total, w.skip, n); ++*launches; }
        }
      }
    }
    PWV_PROF_MARK(m, st);
    if (m->profiling) ++m->prof_launches;
    if (taps && taps->layer_out && taps->layer_flow == flow && taps->layer_index == j && (taps->layer_body == 0 || taps->layer_body == 1))
      PWV_CUDA(cudaMemcpyAsync(taps->layer_out, w.act[cur] + (size_t)taps->layer_body * rows * R, sizeof(float) * rows * R, cudaMemcpyDeviceToDevice, st));
  }
  // post-net: relu -> [norm] -> 1x1 + bias -> relu -> [norm] -> 1x1 + bias (modules.py:148-165)
  float* src = (nrm && hp.use_skip_connection) ? w.total : w.skip;
  if (nrm) {
    rc = launch_in_norm(m, w, src, 2 * N, T, S, b0.n_pp1, b1.n_pp1, N, true, st, launches);      // (applies the relu first)
    if (rc) return rc;
  }
  {
    pwv::GenGemm g = base;
    const BodyOff* bo[2] = {&b0, &b1};
    for (int b = 0; b < 2; ++b) {
      g.A[b] = src + (size_t)b * rows * S;
      g.B[b] = m->d_arena + bo[b]->w1;
      g.bias[b] = m->d_arena + bo[b]->b1;
      g.out[b] = w.hbuf + (size_t)b * rows * S;
    }
    g.lda = S; g.K = S; g.Nc = S; g.relu_in = nrm ? 0 : 1; g.relu_out = 1;
    gen_gemm(g, st, launches);
    if (nrm) {
      rc = launch_in_norm(m, w, w.hbuf, 2 * N, T, S, b0.n_pp2, b1.n_pp2, N, false, st, launches);
      if (rc) return rc;
    }
    pwv::GenGemm h = base;
    for (int b = 0; b < 2; ++b) {
      h.A[b] = w.hbuf + (size_t)b * rows * S;
      h.B[b] = m->d_arena + bo[b]->w2;
      h.bias[b] = m->d_arena + bo[b]->b2;
      h.out[b] = w.ss + (size_t)b * rows;
    }
    h.lda = S; h.K = S; h.Nc = 1;
    gen_gemm(h, st, launches);
  }
  *cur_buf = cur;
  PWV_CUDA(cudaGetLastError());
  return PWV_OK;
}

template <int C>
static int launch_layers_simt(pwv_model* m, const Workspace& w, int flow, int N, int T, cudaStream_t st,
                              const pwv_taps* taps, int* cur_buf, int* launches) {
  using Cfg = pwv::TileCfg<C>;
  const pwv_hparams& hp = m->hp;
  const int L = hp.n_layers[flow], t_mel = cond_rows(m, T), c_hop = cond_hop(m);
  dim3 grid((T + Cfg::TM - 1) / Cfg::TM, N, 2);
  int cur = *cur_buf;
  for (int j = 0; j < L; ++j) {
    pwv::LayerParams p;
    p.x_in = w.act[cur];
    p.x_out = w.act[cur ^ 1];
    for (int b = 0; b < 2; ++b) {
      const LayerOff& lo = m->bodies[flow * 2 + b].layers[j];
      p.wfg[b] = m->d_arena + lo.wfg;
      p.wd[b] = m->d_arena + lo.wd;
      p.bd[b] = m->d_arena + lo.bd;
      p.cbias[b] = w.cbias + ((size_t)b * L + j) * N * t_mel * 2 * C;
    }
    p.N = N; p.T = T; p.t_mel = t_mel; p.hop = c_hop; p.dilation = hp.dilations[flow][j];
    p.mode = (j == L - 1) ? 1 : (hp.use_skip_connection ? 2 : 0);
    p.z_out = w.zbuf;
    PWV_PROF_MARK(m, st);
    pwv::k_layer_simt<C><<<grid, Cfg::NT, Cfg::SMEM, st>>>(p);
    PWV_PROF_MARK(m, st);
    ++*launches;
    cur ^= 1;
    if (hp.use_skip_connection) {     // skip_sum (+)= z_j . Ws_j + bs_j, in layer order (reference modules.py:147)
      pwv::SkipParams sp;
      sp.z = (j == L - 1) ? w.act[cur] : w.zbuf;
      for (int b = 0; b < 2; ++b) {
        const LayerOff& lo = m->bodies[flow * 2 + b].layers[j];
        sp.ws[b] = m->d_arena + lo.ws;
        sp.bs[b] = m->d_arena + lo.bs;
      }
      sp.skip_sum = w.skip; sp.N = N; sp.T = T; sp.first = j == 0 ? 1 : 0;
      pwv::k_skip_simt<C><<<grid, Cfg::NT, Cfg::SMEM, st>>>(sp);
      ++*launches;
    }
    if (taps && taps->layer_out && taps->layer_flow == flow && taps->layer_index == j && (taps->layer_body == 0 || taps->layer_body == 1))
      PWV_CUDA(cudaMemcpyAsync(taps->layer_out, w.act[cur] + (size_t)taps->layer_body * N * T * C,
                               sizeof(float) * (size_t)N * T * C, cudaMemcpyDeviceToDevice, st));
  }
  // post-net on z (in act[cur])
  pwv::PostParams q;
  q.z = w.act[cur];
  for (int b = 0; b < 2; ++b) {
    const BodyOff& bo = m->bodies[flow * 2 + b];
    const LayerOff& lo = bo.layers[L - 1];
    q.ws[b] = m->d_arena + lo.ws; q.bs[b] = m->d_arena + lo.bs;
    q.w1[b] = m->d_arena + bo.w1; q.b1[b] = m->d_arena + bo.b1;
    q.w2[b] = m->d_arena + bo.w2; q.b2[b] = m->d_arena + bo.b2;
  }
  q.y = w.ss; q.N = N; q.T = T;
  q.skip_sum = hp.use_skip_connection ? w.skip : nullptr;
  pwv::k_post_simt<C><<<grid, Cfg::NT, Cfg::SMEM, st>>>(q);
  ++*launches;
  *cur_buf = cur;
  PWV_CUDA(cudaGetLastError());
  return PWV_OK;
}


// 3-D TMA map of an activation buffer [2N utterance-bodies][T][64] fp32, box = 32 channels x 128
// rows x 1, 128B swizzle, zero OOB fill (rows before/after an utterance read as zeros and are not
// written). cuTensorMapEncodeTiled is fetched through the runtime so libcuda is not a link dependency.
static int encode_act_map(CUtensorMap* map, float* base, int N, int T) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PWV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(PWV_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (EncodeFn)fn;
  }
  const cuuint64_t dims[3] = {64, (cuuint64_t)T, (cuuint64_t)2 * N};
  const cuuint64_t strides[2] = {64 * sizeof(float), (cuuint64_t)T * 64 * sizeof(float)};
  const cuuint32_t box[3] = {32, (cuuint32_t)pwv::TC_TM, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PWV_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (N=%d, T=%d)", (int)r, N, T);
  return PWV_OK;
}

// gated layers and post-net of one flow on the tensor cores
static int launch_layers_tc(pwv_model* m, const Workspace& w, const CUtensorMap* maps, int flow, int N, int T, cudaStream_t st,
                            const pwv_taps* taps, int* cur_buf, int* launches) {
  constexpr int C = pwv::TC_C;
  const pwv_hparams& hp = m->hp;
  const int L = hp.n_layers[flow], t_mel = cond_rows(m, T), c_hop = cond_hop(m);
  const bool bf16 = hp.precision == PWV_PREC_BF16;
  // layer-kernel variant (PWV_TC_VARIANT, read at pwv_model_create): 0 = scalar epilogue arithmetic,
  // 1 = packed fp32x2 epilogue arithmetic (bit-identical). Two more variants were measured and removed
  // (DESIGN.md 4.1): all 16 worker warps serving both slots in a static phase order, and a 1024-thread kernel.
  auto kern = bf16 ? pwv::k_layer_tc<true, false> : pwv::k_layer_tc<false, true>;
  if (m->tc_variant == 1) kern = bf16 ? pwv::k_layer_tc<true, false, true> : pwv::k_layer_tc<false, true, true>;
  if (m->tc_variant == 2) kern = bf16 ? pwv::k_layer_tc<true, false, false, true> : pwv::k_layer_tc<false, true, false, true>;
  const int block = pwv::TC_THREADS;
  const int tiles_per_utt = (T + pwv::TC_TM - 1) / pwv::TC_TM;
  const int tiles_body = N * tiles_per_utt;
  int grid = 2 * tiles_body < m->num_sms ? 2 * tiles_body : m->num_sms;
  grid &= ~1;   // the two bodies get the same number of CTAs
  if (grid < 2) grid = 2;
  size_t layer_base = 0;   // index of (flow, body 0, layer 0) in the image array
  for (int i = 0; i < flow; ++i) layer_base += 2 * (size_t)hp.n_layers[i];
  int cur = *cur_buf;
  // One persistent launch for all gated layers of the flow (k_flow_tc). The per-layer launches below remain for
  // per-launch profiling (mode 1), the layer tap of the parity tests, the phase trace and A/B runs (PWV_TC_FLOW=0).
  const bool tap_layer = taps && taps->layer_out && taps->layer_flow == flow;
  // Launch form of the flow's gated layers, by job size (measured: profiles/r1_experiments_after_flow_kernel.txt,
  // tools/sweep_modes.py). One persistent k_flow_tc launch for the whole flow saves the per-layer exit -> launch ->
  // prologue -> refill (~5 us of a 41 us layer at c2), but its steady-state tile is 7-10 % slower than k_layer_tc's (more
  // loop state under the 96-register cap) and a one-tile-per-CTA job pays a flag round trip per layer: it wins between
  // ~2.5 and ~17 tiles per CTA per layer (2 .. 8 utterances of 1 s) and loses outside (c1: -13..-20 %, c3: -8 %).
  // Outside the window: one k_layer_tc launch per layer, chained by programmatic dependent launch, no tile flags.
  // PWV_TC_SEG = n forces k_flow_tc with n layers per launch (>= L: the whole flow).
  const double tiles_per_cta = (double)tiles_body / (grid / 2);
  const bool in_window = tiles_per_cta >= 2.5 && tiles_per_cta <= 17.0;
  const bool flow_kernel = m->use_flow && m->use_flags && (m->tc_variant == 0 || m->tc_variant == 2) && m->profiling != 1 && !tap_layer && (!m->trace || m->trace_flow) &&
                           L <= pwv::TCF_MAX_LAYERS && grid <= m->num_sms && (m->tc_seg > 0 || (in_window && c_hop > 1));
  // (c_hop == 1: full-rate conditioning rows, cond_upsample_method 'transposed_conv' -- read from global memory per row;
  //  that combination is verified on the per-layer kernels only)
  // tile flags between per-layer launches only pay in the same window (the A/B form PWV_TC_FLOW=0)
  const bool layer_flags = m->use_flags && (in_window || m->tc_seg > 0);
  if (flow_kernel) {
    // layers per launch: the whole flow unless PWV_TC_SEG says otherwise (A/B runs, tests)
    int seg = m->tc_seg > 0 ? m->tc_seg : L;
    if (seg > L) seg = L;
    for (int l0 = 0; l0 < L; l0 += seg) {
      const int Ls = (L - l0 < seg) ? L - l0 : seg;
      pwv::TcFlowParams q;
      q.act[0] = w.act[0]; q.act[1] = w.act[1];
      q.images = m->tc.d_images + layer_base * pwv::TC_IMAGE_BYTES;
      q.cbias = w.cbias;
      q.flags = w.flags + (layer_base / 2) * 2 * (size_t)tiles_body;
      q.N = N; q.T = T; q.t_mel = t_mel; q.hop = c_hop; q.cur0 = cur; q.tiles_per_utt = tiles_per_utt;
      q.L_total = L; q.l0 = l0; q.L = Ls; q.final_layer = (l0 + Ls == L) ? 1 : 0;
      q.cb_in_smem = ((pwv::TC_TM - 1) / c_hop + 2 <= pwv::TC_CB_FRAMES) ? 1 : 0;
      for (int j = 0; j < L; ++j) q.dilation[j] = hp.dilations[flow][j];
      q.stagger = m->tc_stagger;
      q.rotate = m->tc_rotate ? 1 : 0;
      q.trace = m->trace; q.trace_layer = m->trace ? m->trace_launch - (int)(layer_base / 2) : -1;
      if (m->profiling == 2 && l0 == 0) PWV_PROF_MARK(m, st);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(pwv::tcf_threads(false));
      cfg.dynamicSmemBytes = pwv::TCF_SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = m->use_pdl ? 1 : 0;
      if (m->tc_variant == 2) {
        if (bf16) PWV_CUDA(cudaLaunchKernelEx(&cfg, pwv::k_flow_tc<true, false, false, false, true>, maps[0], maps[1], q));
        else PWV_CUDA(cudaLaunchKernelEx(&cfg, pwv::k_flow_tc<false, true, false, false, true>, maps[0], maps[1], q));
      } else if (bf16) PWV_CUDA(cudaLaunchKernelEx(&cfg, pwv::k_flow_tc<true, false>, maps[0], maps[1], q));
      else PWV_CUDA(cudaLaunchKernelEx(&cfg, pwv::k_flow_tc<false, true>, maps[0], maps[1], q));
      if (m->profiling == 2 && l0 + Ls == L) PWV_PROF_MARK(m, st);
      ++*launches;
      cur ^= (Ls & 1);
    }
    if (m->profiling) m->prof_launches += L;
  }
  for (int j = 0; j < L && !flow_kernel; ++j) {
    pwv::TcLayerParams p;
    p.x_out = w.act[cur ^ 1];
    for (int b = 0; b < 2; ++b) {
      p.image[b] = m->tc.d_images + (layer_base + (size_t)b * L + j) * pwv::TC_IMAGE_BYTES;
      p.cbias[b] = w.cbias + ((size_t)b * L + j) * N * t_mel * 2 * C;
    }
    p.N = N; p.T = T; p.t_mel = t_mel; p.hop = c_hop; p.dilation = hp.dilations[flow][j];
    p.mode = (j == L - 1) ? 1 : 0;
    p.tiles_per_utt = tiles_per_utt;
    {
      // tile handshake with the previous gated layer of the flow (see TcLayerParams); the first layer of a flow
      // follows k_front and waits for it as a whole. PWV_NO_TILE_FLAGS=1 restores whole-kernel waits everywhere.
      int* fl = w.flags + (layer_base / 2 + (size_t)j) * 2 * tiles_body;
      p.flags_out = (layer_flags && j + 1 < L) ? fl : nullptr;
      p.flags_in = (layer_flags && j > 0) ? fl - 2 * (size_t)tiles_body : nullptr;
      p.prev_dilation = j > 0 ? hp.dilations[flow][j - 1] : 0;
    }
    p.cb_in_smem = ((pwv::TC_TM - 1) / c_hop + 2 <= pwv::TC_CB_FRAMES) ? 1 : 0;
    p.trace = (m->trace && m->trace_launch == (int)(layer_base / 2) + j) ? m->trace : nullptr;
    if (m->profiling == 1 || (m->profiling == 2 && j == 0)) PWV_PROF_MARK(m, st);
    {
      // programmatic dependent launch: this layer's prologue overlaps the previous kernel's tail
      // (not while profiling: the events between the launches would serialise them anyway)
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(block);
      cfg.dynamicSmemBytes = pwv::TC_SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = (m->profiling == 1 || !m->use_pdl) ? 0 : 1;
      PWV_CUDA(cudaLaunchKernelEx(&cfg, kern, maps[cur], p));
    }
    if (m->profiling == 1 || (m->profiling == 2 && j == L - 1)) PWV_PROF_MARK(m, st);
    if (m->profiling) ++m->prof_launches;
    ++*launches;
    cur ^= 1;
    if (taps && taps->layer_out && taps->layer_flow == flow && taps->layer_index == j && (taps->layer_body == 0 || taps->layer_body == 1))
      PWV_CUDA(cudaMemcpyAsync(taps->layer_out, w.act[cur] + (size_t)taps->layer_body * N * T * C,
                               sizeof(float) * (size_t)N * T * C, cudaMemcpyDeviceToDevice, st));
  }
  // post-net of both bodies on the tensor cores; the two channel halves of a row accumulate into y
  PWV_CUDA(cudaMemsetAsync(w.ss, 0, sizeof(float) * 2 * (size_t)N * T, st));
  pwv::TcPostParams q;
  for (int b = 0; b < 2; ++b) q.image[b] = m->tc.d_post + ((size_t)flow * 2 + b) * pwv::TCP_IMAGE_BYTES;
  q.y = w.ss; q.N = N; q.T = T; q.tiles_per_utt = tiles_per_utt;
  if (bf16) pwv::k_post_tc<true, false><<<grid, pwv::TC_THREADS, pwv::TCP_SMEM_BYTES, st>>>(maps[cur], q);
  else pwv::k_post_tc<false, true><<<grid, pwv::TC_THREADS, pwv::TCP_SMEM_BYTES, st>>>(maps[cur], q);
  ++*launches;
  *cur_buf = cur;
  PWV_CUDA(cudaGetLastError());
  return PWV_OK;
}


// 3-D TMA map of an activation buffer in the plane layout [planes * 2N utterance-bodies][T][64] 16-bit (fp16 hi / lo
// planes, or one bf16 plane): box = 64 channels x 128 rows x 1 (16 KB), 128B swizzle, zero OOB fill.
static int encode_plane_map(CUtensorMap* map, void* base, int N, int T, int planes, bool bf16, int C = 64) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PWV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(PWV_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (EncodeFn)fn;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)planes * 2 * N};
  const cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)pwv::TC_TM, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PWV_ECUDA, "cuTensorMapEncodeTiled (planes) failed with CUresult %d (N=%d, T=%d)", (int)r, N, T);
  return PWV_OK;
}

template <bool BF16, bool LAST>
static cudaError_t launch_layer_h(const cudaLaunchConfig_t* cfg, int variant, const CUtensorMap& in, const CUtensorMap& out, const pwv::ThLayerParams& p) {
  if (variant == 2 && BF16) return cudaLaunchKernelEx(cfg, pwv::k_layer_h<BF16, LAST, true, false>, in, out, p);   // A/B: ex2 / rcp gate in bf16 mode
  if (variant != 0) return cudaLaunchKernelEx(cfg, pwv::k_layer_h<BF16, LAST, true>, in, out, p);
  return cudaLaunchKernelEx(cfg, pwv::k_layer_h<BF16, LAST, false>, in, out, p);
}

// gated layers and post-net of one flow on the tensor cores, activations as 16-bit planes (k_layer_h, pwv_tc2.cuh):
// one launch per gated layer, chained by programmatic dependent launch (+ the per-tile flag handshake inside the
// job-size window where it pays); the flow's last layer writes z as fp32 rows for k_post_tc.
//   maps_h[b] = plane map of act[b], maps_f[b] = fp32 row map of act[b]
static int launch_layers_h(pwv_model* m, const Workspace& w, const CUtensorMap* maps_h, const CUtensorMap* maps_f, int flow, int N, int T,
                           cudaStream_t st, const pwv_taps* taps, int* cur_buf, int* launches) {
  constexpr int C = pwv::TC_C;
  const pwv_hparams& hp = m->hp;
  const int L = hp.n_layers[flow], t_mel = cond_rows(m, T), c_hop = cond_hop(m);
  const bool bf16 = hp.precision == PWV_PREC_BF16;
  const int tiles_per_utt = (T + pwv::TC_TM - 1) / pwv::TC_TM;
  const int tiles_body = N * tiles_per_utt;
  int grid = 2 * tiles_body < m->num_sms ? 2 * tiles_body : m->num_sms;
  grid &= ~1;
  if (grid < 2) grid = 2;
  size_t layer_base = 0;
  for (int i = 0; i < flow; ++i) layer_base += 2 * (size_t)hp.n_layers[i];
  int cur = *cur_buf;
  // Tile flags between consecutive layers let a layer's tiles start while the previous layer's tail is still running;
  // they pay while a CTA owns a handful of tiles per layer (round-1 sweep: 2.5 .. 17; below that a tile-level flag round trip per layer costs more than the whole-kernel wait: c1), beyond that the whole-kernel wait of the
  // programmatic dependent launch costs less than the flag traffic.
  const double tiles_per_cta = (double)tiles_body / (grid / 2);
  // (use_skip_connection puts a skip-sum kernel between consecutive layers: no tile handshake across it)
  const bool layer_flags = m->use_flags && m->use_pdl && m->profiling != 1 && !hp.use_skip_connection && tiles_per_cta >= 2.5 && tiles_per_cta <= 17.0;
  const size_t plane_elems = (size_t)2 * N * T * C;
  using SCfg = pwv::TileCfg<64>;
  const dim3 sgrid((T + SCfg::TM - 1) / SCfg::TM, N, 2);
  for (int j = 0; j < L; ++j) {
    const bool last = j == L - 1;
    pwv::ThLayerParams p;
    for (int b = 0; b < 2; ++b) {
      p.image[b] = m->tc.d_images + (layer_base + (size_t)b * L + j) * pwv::TC_IMAGE_BYTES;
      p.cbias[b] = w.cbias + ((size_t)b * L + j) * N * t_mel * 2 * C;
    }
    p.N = N; p.T = T; p.t_mel = t_mel; p.hop = c_hop; p.dilation = hp.dilations[flow][j];
    p.tiles_per_utt = tiles_per_utt;
    p.cb_in_smem = ((pwv::TC_TM - 1) / c_hop + 2 <= pwv::TC_CB_FRAMES) ? 1 : 0;
    int* fl = w.flags + (layer_base / 2 + (size_t)j) * 2 * tiles_body;
    p.flags_out = (layer_flags && j + 1 < L) ? fl : nullptr;
    p.flags_in = (layer_flags && j > 0) ? fl - 2 * (size_t)tiles_body : nullptr;
    p.prev_dilation = j > 0 ? hp.dilations[flow][j - 1] : 0;
    {
      int* done = w.flags + (size_t)m->total_layers * 2 * tiles_body + layer_base / 2 + j;
      p.done_out = p.flags_out ? done : nullptr;
      p.done_in = p.flags_in ? done - 1 : nullptr;
      p.done_target = 2 * grid;          // both producers of every CTA of the previous layer's grid (same grid for the whole flow)
    }
    p.use_cp = m->use_cp ? 1 : 0;
    p.split1 = m->split1 ? 1 : 0;
    p.split2 = m->split2 ? 1 : 0;
    p.double_a = m->double_a ? 1 : 0;
    p.z_in_d = m->z_in_d ? 1 : 0;
    p.z_out = (hp.use_skip_connection && !last) ? w.zbuf : nullptr;
    p.trace = (m->trace && m->trace_launch == (int)(layer_base / 2) + j) ? m->trace : nullptr;
    if (m->profiling == 1 || (m->profiling == 2 && j == 0)) PWV_PROF_MARK(m, st);
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(pwv::TC_THREADS);
      cfg.dynamicSmemBytes = pwv::TH_SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = (m->profiling == 1 || !m->use_pdl) ? 0 : 1;
      const int pk = m->tc_variant;
      const CUtensorMap& in = maps_h[cur];
      const CUtensorMap& out = last ? maps_f[cur ^ 1] : maps_h[cur ^ 1];
      cudaError_t e;
      if (bf16) e = last ? launch_layer_h<true, true>(&cfg, pk, in, out, p) : launch_layer_h<true, false>(&cfg, pk, in, out, p);
      else e = last ? launch_layer_h<false, true>(&cfg, pk, in, out, p) : launch_layer_h<false, false>(&cfg, pk, in, out, p);
      PWV_CUDA(e);
    }
    if (m->profiling == 1 || (m->profiling == 2 && last)) PWV_PROF_MARK(m, st);
    if (m->profiling) ++m->prof_launches;
    ++*launches;
    cur ^= 1;
    if (hp.use_skip_connection) {     // skip_sum (+)= z_j . Ws_j + bs_j in layer order (reference modules.py:147), exact fp32 FFMA kernel
      pwv::SkipParams sp;
      sp.z = last ? w.act[cur] : w.zbuf;
      for (int b = 0; b < 2; ++b) {
        const LayerOff& lo = m->bodies[flow * 2 + b].layers[j];
        sp.ws[b] = m->d_arena + lo.ws;
        sp.bs[b] = m->d_arena + lo.bs;
      }
      sp.skip_sum = w.skip; sp.N = N; sp.T = T; sp.first = j == 0 ? 1 : 0;
      pwv::k_skip_simt<64><<<sgrid, SCfg::NT, SCfg::SMEM, st>>>(sp);
      ++*launches;
    }
    if (taps && taps->layer_out && taps->layer_flow == flow && taps->layer_index == j && (taps->layer_body == 0 || taps->layer_body == 1)) {
      const size_t rows = (size_t)N * T;
      if (last) {       // (the last layer's output buffer holds z as fp32 rows)
        PWV_CUDA(cudaMemcpyAsync(taps->layer_out, w.act[cur] + (size_t)taps->layer_body * rows * C, sizeof(float) * rows * C, cudaMemcpyDeviceToDevice, st));
      } else {
        const uint16_t* src = reinterpret_cast<const uint16_t*>(w.act[cur]) + (size_t)taps->layer_body * rows * C;
        const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
        if (bf16) pwv::k_planes_to_f32<true><<<blocks, 256, 0, st>>>(src, taps->layer_out, rows, plane_elems);
        else pwv::k_planes_to_f32<false><<<blocks, 256, 0, st>>>(src, taps->layer_out, rows, plane_elems);
        ++*launches;
      }
    }
  }
  if (hp.use_skip_connection) {       // post-net on relu(sum of the skip outputs): the fp32 kernel
    pwv::PostParams q;
    q.z = w.act[cur];
    for (int b = 0; b < 2; ++b) {
      const BodyOff& bo = m->bodies[flow * 2 + b];
      const LayerOff& lo = bo.layers[L - 1];
      q.ws[b] = m->d_arena + lo.ws; q.bs[b] = m->d_arena + lo.bs;
      q.w1[b] = m->d_arena + bo.w1; q.b1[b] = m->d_arena + bo.b1;
      q.w2[b] = m->d_arena + bo.w2; q.b2[b] = m->d_arena + bo.b2;
    }
    q.y = w.ss; q.N = N; q.T = T;
    q.skip_sum = w.skip;
    pwv::k_post_simt<64><<<sgrid, SCfg::NT, SCfg::SMEM, st>>>(q);
    ++*launches;
    *cur_buf = cur;
    PWV_CUDA(cudaGetLastError());
    return PWV_OK;
  }
  PWV_CUDA(cudaMemsetAsync(w.ss, 0, sizeof(float) * 2 * (size_t)N * T, st));
  pwv::TcPostParams q;
  for (int b = 0; b < 2; ++b) q.image[b] = m->tc.d_post + ((size_t)flow * 2 + b) * pwv::TCP_IMAGE_BYTES;
  q.y = w.ss; q.N = N; q.T = T; q.tiles_per_utt = tiles_per_utt;
  if (bf16) pwv::k_post_tc<true, false><<<grid, pwv::TC_THREADS, pwv::TCP_SMEM_BYTES, st>>>(maps_f[cur], q);
  else pwv::k_post_tc<false, true><<<grid, pwv::TC_THREADS, pwv::TCP_SMEM_BYTES, st>>>(maps_f[cur], q);
  ++*launches;
  *cur_buf = cur;
  PWV_CUDA(cudaGetLastError());
  return PWV_OK;
}


// gated layers of one flow at C > 64 channels on the tensor cores (k_wide_h, pwv_tc3.cuh): per layer a gate pass
// (x planes -> z planes) and a dense pass (z planes + x planes -> next x planes); the flow's last layer has no dense
// pass. The post-net runs on the fp32 kernel from z converted back to fp32 rows.
template <int C>
static int launch_layers_w(pwv_model* m, const Workspace& w, const CUtensorMap* maps_h, const CUtensorMap& map_z, int flow, int N, int T,
                           cudaStream_t st, const pwv_taps* taps, int* cur_buf, int* launches) {
  const pwv_hparams& hp = m->hp;
  const int L = hp.n_layers[flow], t_mel = cond_rows(m, T), c_hop = cond_hop(m);
  const bool bf16 = hp.precision == PWV_PREC_BF16;
  const int tiles_per_utt = (T + pwv::TC_TM - 1) / pwv::TC_TM;
  const int NB = C / 64;
  const long long items = 2LL * N * tiles_per_utt * NB;
  const int grid = items < m->num_sms ? (int)items : m->num_sms;
  size_t layer_base = 0;
  for (int i = 0; i < flow; ++i) layer_base += 2 * (size_t)hp.n_layers[i];
  const size_t plane_elems = (size_t)2 * N * T * C;
  int cur = *cur_buf;
  for (int j = 0; j < L; ++j) {
    const bool last = j == L - 1;
    pwv::TwParams p;
    p.N = N; p.T = T; p.t_mel = t_mel; p.hop = c_hop; p.dilation = hp.dilations[flow][j]; p.tiles_per_utt = tiles_per_utt;
    p.NB = NB; p.C_out = C;
    for (int b = 0; b < 2; ++b) {
      const size_t li = layer_base + (size_t)b * L + j;
      p.wimg[b] = m->tw.d_gate + li * m->tw.gate_bytes;
      p.vec[b] = m->tw.d_vec + li * m->tw.vec_floats;
      p.cbias[b] = w.cbias + ((size_t)b * L + j) * N * t_mel * 2 * C;
    }
    p.KB = 2 * C / 64; p.KB_tap = C / 64;
    if (m->profiling == 1 || (m->profiling == 2 && j == 0)) PWV_PROF_MARK(m, st);
    if (bf16) pwv::k_wide_h<true, pwv::TW_EPI_GATE><<<grid, pwv::TW_THREADS, pwv::TW_SMEM_BYTES, st>>>(maps_h[cur], maps_h[cur], map_z, p);
    else pwv::k_wide_h<false, pwv::TW_EPI_GATE><<<grid, pwv::TW_THREADS, pwv::TW_SMEM_BYTES, st>>>(maps_h[cur], maps_h[cur], map_z, p);
    ++*launches;
    if (!last) {
      for (int b = 0; b < 2; ++b) p.wimg[b] = m->tw.d_dense + (layer_base + (size_t)b * L + j) * m->tw.dense_bytes;
      p.KB = C / 64; p.KB_tap = p.KB; p.dilation = 0;
      if (bf16) pwv::k_wide_h<true, pwv::TW_EPI_DENSE><<<grid, pwv::TW_THREADS, pwv::TW_SMEM_BYTES, st>>>(map_z, maps_h[cur], maps_h[cur ^ 1], p);
      else pwv::k_wide_h<false, pwv::TW_EPI_DENSE><<<grid, pwv::TW_THREADS, pwv::TW_SMEM_BYTES, st>>>(map_z, maps_h[cur], maps_h[cur ^ 1], p);
      ++*launches;
    } else {      // z planes -> fp32 rows in the other activation buffer: the post-net kernel's input
      const size_t pairs = plane_elems / 2;
      const unsigned blocks = (unsigned)((pairs + 255) / 256);
      if (bf16) pwv::k_planes_to_f32_n<true><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(w.zbuf), w.act[cur ^ 1], pairs, plane_elems);
      else pwv::k_planes_to_f32_n<false><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(w.zbuf), w.act[cur ^ 1], pairs, plane_elems);
      ++*launches;
    }
    if (m->profiling == 1 || (m->profiling == 2 && last)) PWV_PROF_MARK(m, st);
    if (m->profiling) ++m->prof_launches;
    cur ^= 1;
    if (taps && taps->layer_out && taps->layer_flow == flow && taps->layer_index == j && (taps->layer_body == 0 || taps->layer_body == 1)) {
      const size_t rows = (size_t)N * T;
      if (last) {
        PWV_CUDA(cudaMemcpyAsync(taps->layer_out, w.act[cur] + (size_t)taps->layer_body * rows * C, sizeof(float) * rows * C, cudaMemcpyDeviceToDevice, st));
      } else {
        const uint16_t* src = reinterpret_cast<const uint16_t*>(w.act[cur]) + (size_t)taps->layer_body * rows * C;
        const size_t pairs = rows * C / 2;
        const unsigned blocks = (unsigned)((pairs + 255) / 256);
        if (bf16) pwv::k_planes_to_f32_n<true><<<blocks, 256, 0, st>>>(src, taps->layer_out, pairs, plane_elems);
        else pwv::k_planes_to_f32_n<false><<<blocks, 256, 0, st>>>(src, taps->layer_out, pairs, plane_elems);
        ++*launches;
      }
    }
  }
  {
    using Cfg = pwv::TileCfg<C>;
    const dim3 sgrid((T + Cfg::TM - 1) / Cfg::TM, N, 2);
    pwv::PostParams q;
    q.z = w.act[cur];
    for (int b = 0; b < 2; ++b) {
      const BodyOff& bo = m->bodies[flow * 2 + b];
      const LayerOff& lo = bo.layers[L - 1];
      q.ws[b] = m->d_arena + lo.ws; q.bs[b] = m->d_arena + lo.bs;
      q.w1[b] = m->d_arena + bo.w1; q.b1[b] = m->d_arena + bo.b1;
      q.w2[b] = m->d_arena + bo.w2; q.b2[b] = m->d_arena + bo.b2;
    }
    q.y = w.ss; q.N = N; q.T = T;
    q.skip_sum = nullptr;
    pwv::k_post_simt<C><<<sgrid, Cfg::NT, Cfg::SMEM, st>>>(q);
    ++*launches;
  }
  *cur_buf = cur;
  PWV_CUDA(cudaGetLastError());
  return PWV_OK;
}

extern "C" {

int pwv_forward(pwv_model* m, const float* noise, const float* mel, float* wav, void* workspace,
                size_t workspace_bytes, int N, int T, pwv_stream stream, const pwv_taps* taps) {
  int rc = check_shape(m, N, T);
  if (rc) return rc;
  if (!m->finalized) return fail(PWV_ESTATE, "pwv_model_finalize has not been called");
  if (!noise || !mel || !wav || !workspace) return fail(PWV_EINVAL, "null buffer");
  Workspace w;
  carve(m, N, T, (char*)workspace, &w);
  if (w.bytes > workspace_bytes) return fail(PWV_ENOMEM, "workspace has %zu bytes, %zu needed", workspace_bytes, w.bytes);
  if (((uintptr_t)workspace & 255) != 0) return fail(PWV_EINVAL, "workspace must be 256-byte aligned");
  if (m->hp.use_skip_connection && m->hp.precision != PWV_PREC_FP32 && m->tc_path == 0)
    return fail(PWV_EINVAL, "use_skip_connection=True is not implemented on the round-1 tensor-core kernels (debug path 0)");
  const pwv_hparams& hp = m->hp;
  const int C = m->C, Cc = m->Cc, t_mel = 1 + T / hp.hop_length;
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  m->ev_used = 0;
  m->prof_launches = 0;
  PWV_PROF_MARK(m, st);   // [0] forward start
  PWV_PROF_MARK(m, st);   // [1] placeholder, re-recorded at the end
  if (w.flags) PWV_CUDA(cudaMemsetAsync(w.flags, 0, w.flags_bytes, st));

  const int crows = cond_rows(m, T);     // conditioning rows per utterance (t_mel, or T for transposed_conv)
  if (hp.cond_upsample == PWV_UPSAMPLE_TRANSPOSED_CONV) {
    // reference models.py:109-124: stacked conv2d_transpose (kernel width = stride, so every output position has
    // exactly one source frame) + relu per stage, then the crop [hop/2 : -hop/2]. Stage i is a row GEMM
    // [rows][Cin] x [Cin][stride * Cc] whose output, read as [rows * stride][Cc], is the upsampled sequence.
    const float* src = mel;
    int rows = N * t_mel, cin = hp.n_mels;
    for (int i = 0; i < hp.n_upsample; ++i) {
      const int nc = hp.upsample_strides[i] * Cc;
      float* dst = w.up[i & 1];
      pwv::RowGemmBatch rb{m->d_arena + m->off_wup[i], 0, nullptr, 0, dst, 0, nullptr};
      dim3 grid((nc + 63) / 64, (rows + 63) / 64, 1);
      pwv::k_row_gemm<true><<<grid, 256, 0, st>>>(src, rb, rows, cin, nc);
      ++launches;
      src = dst;
      rows *= hp.upsample_strides[i];
      cin = Cc;
      if (hp.normalize_cond == PWV_NORM_IN) {
        // reference models.py:121-122 normalises the 4-D tensor (n, 1, len, C) over its size-1 axis (modules.py:277):
        // mean = x, variance = 0, so the stage's output is its beta, whatever the input (replayed literally)
        const size_t n_el = (size_t)rows * Cc;
        pwv::k_fill_rows<<<(unsigned)((n_el + 255) / 256), 256, 0, st>>>(dst, m->d_arena + m->n_up[i].beta, (size_t)rows, Cc);
        ++launches;
      }
    }
    // crop: utterance n keeps rows hop/2 .. hop/2 + T - 1 of its t_mel * hop upsampled rows
    PWV_CUDA(cudaMemcpy2DAsync(w.cproj, sizeof(float) * (size_t)T * Cc, src + (size_t)(hp.hop_length / 2) * Cc,
                               sizeof(float) * (size_t)t_mel * hp.hop_length * Cc, sizeof(float) * (size_t)T * Cc, N,
                               cudaMemcpyDeviceToDevice, st));
  } else {
    // conditioning: cproj = relu(mel . Wc)   (reference models.py:128-130, at mel rate)
    if (hp.cond_upsample == PWV_UPSAMPLE_NONE) {              // reference models.py:134-135: no conditioning; `mel` is not read
      PWV_CUDA(cudaMemsetAsync(w.cproj, 0, sizeof(float) * (size_t)N * t_mel * Cc, st));
    } else {
    const bool full = hp.normalize_cond == PWV_NORM_IN;       // then the repeated, cropped rows are materialised for the statistics
    pwv::RowGemmBatch rb{m->d_arena + m->off_wc, 0, nullptr, 0, full ? w.up[0] : w.cproj, 0, nullptr};
    const int M = N * t_mel;
    dim3 grid((Cc + 63) / 64, (M + 63) / 64, 1);
    pwv::k_row_gemm<true><<<grid, 256, 0, st>>>(mel, rb, M, hp.n_mels, Cc);
    ++launches;
    if (full) {
      const size_t n_el = (size_t)N * T * Cc;
      pwv::k_repeat_crop<<<(unsigned)((n_el + 255) / 256), 256, 0, st>>>(w.up[0], w.cproj, N, T, t_mel, hp.hop_length, Cc);
      ++launches;
    }
    }
  }
  if (hp.normalize_cond == PWV_NORM_IN) {                       // models.py:27-29, after the crop
    rc = launch_in_norm(m, w, w.cproj, N, T, Cc, m->n_cond, m->n_cond, N, false, st, &launches);
    if (rc) return rc;
  }

  CUtensorMap maps[2], maps_h[2], map_z;
  const bool wide = hp.precision != PWV_PREC_FP32 && C != pwv::TC_C;
  const bool planes = hp.precision != PWV_PREC_FP32 && (m->tc_path == 1 || wide);
  if (hp.precision != PWV_PREC_FP32) {
    const bool bf = hp.precision == PWV_PREC_BF16;
    for (int b = 0; b < 2; ++b) {
      if (!wide) {
        rc = encode_act_map(&maps[b], w.act[b], N, T);
        if (rc) return rc;
      }
      if (planes) {
        rc = encode_plane_map(&maps_h[b], w.act[b], N, T, bf ? 1 : 2, bf, C);
        if (rc) return rc;
      }
    }
    if (wide) {
      rc = encode_plane_map(&map_z, w.zbuf, N, T, bf ? 1 : 2, bf, C);
      if (rc) return rc;
    }
  }

  int cur = 0, xcur = 0;
  const float* x_prev = noise;
  const bool nrm_f = hp.normalize == PWV_NORM_IN;     // x is combined and normalised explicitly after every flow
  if ((hp.normalize_wavenet == PWV_NORM_IN || m->general) && ((size_t)N * T + 63) / 64 > 65535)
    return fail(PWV_EINVAL, "%s: N*T = %zu exceeds the un-fused chain's grid (4,194,240 samples per call)",
                m->general ? "general-shape path" : "normalize_wavenet", (size_t)N * T);
  for (int i = 0; i < hp.n_iaf; ++i) {
    const int L = hp.n_layers[i];
    // per-layer conditioning terms of this flow: cbias[b][j] = cproj . [gc_filter|gc_gate] + [bf|bg]
    {
      const int D = m->D;      // (= C on the fused paths)
      pwv::RowGemmBatch rb{m->d_arena + m->off_wgc[i], (size_t)Cc * 2 * D, m->d_arena + m->off_bfg[i], (size_t)2 * D,
                           w.cbias, (size_t)N * crows * 2 * D,
                           hp.precision == PWV_PREC_FP32 ? nullptr : m->d_arena + m->off_colscale};
      const int M = N * crows;
      if (hp.precision != PWV_PREC_FP32 && m->tc.d_cond) {
        pwv::TcCondParams q;
        size_t zbase = 0;
        for (int f = 0; f < i; ++f) zbase += 2 * (size_t)hp.n_layers[f];
        q.cproj = w.cproj;
        q.images = m->tc.d_cond + zbase * m->tc.cond_image_bytes;
        q.out = w.cbias;
        q.out_stride = (size_t)M * 2 * C;
        q.M = M; q.Cc = Cc; q.Z = 2 * L; q.image_bytes = m->tc.cond_image_bytes;
        const int m_tiles = (M + pwv::TC_TM - 1) / pwv::TC_TM;
        int zsplit = m->num_sms / m_tiles;
        zsplit = zsplit < 1 ? 1 : (zsplit > 2 * L ? 2 * L : zsplit);
        dim3 grid(m_tiles, zsplit);
        if (hp.precision == PWV_PREC_BF16) pwv::k_cbias_tc<true, false><<<grid, pwv::TCC_THREADS, pwv::tcc_smem_bytes(Cc), st>>>(q);
        else pwv::k_cbias_tc<false, true><<<grid, pwv::TCC_THREADS, pwv::tcc_smem_bytes(Cc), st>>>(q);
      } else if (!m->general && Cc % 16 == 0 && Cc <= 2 * C) {
        rc = launch_cond_gemm(C, w.cproj, rb, M, Cc, 2 * L, st);
        if (rc) return rc;
      } else {
        dim3 grid((2 * D + 63) / 64, (M + 63) / 64, 2 * L);
        pwv::k_row_gemm<false><<<grid, 256, 0, st>>>(w.cproj, rb, M, Cc, 2 * D);
      }
      ++launches;
    }
    // front: IAF combine of the previous flow + causal layers
    if (m->general) {
      pwv::GenFront f;
      f.x_prev = x_prev;
      f.scale = (i == 0 || nrm_f) ? nullptr : w.ss;
      f.shift = (i == 0 || nrm_f) ? nullptr : w.ss + (size_t)N * T;
      f.x_new = w.x[xcur];
      f.wc[0] = m->d_arena + m->bodies[i * 2 + 0].causal;
      f.wc[1] = m->d_arena + m->bodies[i * 2 + 1].causal;
      f.act = w.act[cur];
      f.N = N; f.T = T; f.R = C; f.taps = m->taps;
      const size_t total = (size_t)N * T * C;
      pwv::k_gen_front<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(f);
      ++launches;
      x_prev = w.x[xcur];
      xcur ^= 1;
    } else if (wide) {
      pwv::FrontWParams f;
      f.x_prev = x_prev;
      f.scale = i == 0 ? nullptr : w.ss;
      f.shift = i == 0 ? nullptr : w.ss + (size_t)N * T;
      f.x_new = w.x[xcur];
      f.wc[0] = m->d_arena + m->bodies[i * 2 + 0].causal;
      f.wc[1] = m->d_arena + m->bodies[i * 2 + 1].causal;
      f.act = reinterpret_cast<uint16_t*>(w.act[cur]);
      f.N = N; f.T = T; f.C = C;
      const size_t total = (size_t)N * T * (C / 8);
      if (hp.precision == PWV_PREC_BF16) pwv::k_front_w<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(f);
      else pwv::k_front_w<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(f);
      ++launches;
      x_prev = w.x[xcur];
      xcur ^= 1;
    } else if (planes) {
      pwv::FrontHParams f;
      f.x_prev = x_prev;
      f.scale = i == 0 ? nullptr : w.ss;
      f.shift = i == 0 ? nullptr : w.ss + (size_t)N * T;
      f.x_new = w.x[xcur];
      f.wc[0] = m->d_arena + m->bodies[i * 2 + 0].causal;
      f.wc[1] = m->d_arena + m->bodies[i * 2 + 1].causal;
      f.act = reinterpret_cast<uint16_t*>(w.act[cur]);
      f.N = N; f.T = T;
      const size_t total = (size_t)N * T * (C / 8);
      if (hp.precision == PWV_PREC_BF16) pwv::k_front_h<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(f);
      else pwv::k_front_h<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(f);
      ++launches;
      x_prev = w.x[xcur];
      xcur ^= 1;
    } else {
      pwv::FrontParams f;
      f.x_prev = x_prev;
      f.scale = (i == 0 || nrm_f) ? nullptr : w.ss;
      f.shift = (i == 0 || nrm_f) ? nullptr : w.ss + (size_t)N * T;
      f.x_new = w.x[xcur];
      f.wc[0] = m->d_arena + m->bodies[i * 2 + 0].causal;
      f.wc[1] = m->d_arena + m->bodies[i * 2 + 1].causal;
      f.act = w.act[cur];
      f.N = N; f.T = T; f.C = C;
      const size_t total = (size_t)N * T * (C / 4);
      pwv::k_front<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(f);
      ++launches;
      x_prev = w.x[xcur];
      xcur ^= 1;
    }
    if (m->general) {
      rc = launch_layers_gen(m, w, i, N, T, st, taps, &cur, &launches);
    } else if (hp.normalize_wavenet == PWV_NORM_IN) {
      if (C == 64) rc = launch_layers_in<64>(m, w, i, N, T, st, taps, &cur, &launches);
      else if (C == 128) rc = launch_layers_in<128>(m, w, i, N, T, st, taps, &cur, &launches);
      else rc = launch_layers_in<256>(m, w, i, N, T, st, taps, &cur, &launches);
    } else if (hp.precision == PWV_PREC_FP32) {
      if (C == 64) rc = launch_layers_simt<64>(m, w, i, N, T, st, taps, &cur, &launches);
      else if (C == 128) rc = launch_layers_simt<128>(m, w, i, N, T, st, taps, &cur, &launches);
      else rc = launch_layers_simt<256>(m, w, i, N, T, st, taps, &cur, &launches);
    } else if (wide) {
      if (C == 128) rc = launch_layers_w<128>(m, w, maps_h, map_z, i, N, T, st, taps, &cur, &launches);
      else rc = launch_layers_w<256>(m, w, maps_h, map_z, i, N, T, st, taps, &cur, &launches);
    } else if (planes) {
      rc = launch_layers_h(m, w, maps_h, maps, i, N, T, st, taps, &cur, &launches);
    } else {
      rc = launch_layers_tc(m, w, maps, i, N, T, st, taps, &cur, &launches);
    }
    if (rc) return rc;
    if (taps && taps->scale_shift)
      PWV_CUDA(cudaMemcpyAsync(taps->scale_shift + (size_t)i * 2 * N * T, w.ss, sizeof(float) * 2 * (size_t)N * T, cudaMemcpyDeviceToDevice, st));
    if (nrm_f) {        // x = x * scale + shift (modules.py:57-59), then normalize{i} over time (models.py:70)
      const size_t n = (size_t)N * T;
      float* xn = w.x[xcur];
      pwv::k_iaf_combine<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x_prev, w.ss, w.ss + n, xn, n);
      ++launches;
      rc = launch_in_norm(m, w, xn, N, T, 1, m->n_flow[i], m->n_flow[i], N, false, st, &launches);
      if (rc) return rc;
      x_prev = xn;
      xcur ^= 1;
      if (taps && taps->flow_out) PWV_CUDA(cudaMemcpyAsync(taps->flow_out + (size_t)i * n, xn, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    } else if (taps && taps->flow_out) {
      const size_t n = (size_t)N * T;
      pwv::k_iaf_combine<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x_prev, w.ss, w.ss + n, taps->flow_out + (size_t)i * n, n);
      ++launches;
    }
  }
  if (nrm_f) {
    PWV_CUDA(cudaMemcpyAsync(wav, x_prev, sizeof(float) * (size_t)N * T, cudaMemcpyDeviceToDevice, st));
  } else {
    const size_t n = (size_t)N * T;
    pwv::k_iaf_combine<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x_prev, w.ss, w.ss + n, wav, n);
    ++launches;
  }
  if (m->profiling && m->ev.size() >= 2) cudaEventRecord(m->ev[1], st);
  PWV_CUDA(cudaGetLastError());
  m->last_launches = launches;
  return PWV_OK;
}

int pwv_forward_host(pwv_model* m, const float* noise, const float* mel, float* wav, int N, int T, pwv_stream stream) {
  int rc = check_shape(m, N, T);
  if (rc) return rc;
  if (!noise || !mel || !wav) return fail(PWV_EINVAL, "null buffer");
  const int t_mel = 1 + T / m->hp.hop_length;
  size_t ws = 0;
  rc = pwv_workspace_bytes(m, N, T, &ws);
  if (rc) return rc;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t b_noise = al(sizeof(float) * (size_t)N * T), b_mel = al(sizeof(float) * (size_t)N * t_mel * m->hp.n_mels);
  const size_t need = ws + 2 * b_noise + b_mel;
  if (need > m->h_dev_bytes) {
    if (m->h_dev) cudaFree(m->h_dev);
    m->h_dev = nullptr;
    m->h_dev_bytes = 0;
    PWV_CUDA(cudaMalloc(&m->h_dev, need));
    m->h_dev_bytes = need;
  }
  char* base = (char*)m->h_dev;
  float* d_noise = (float*)(base + ws);
  float* d_wav = (float*)(base + ws + b_noise);
  float* d_mel = (float*)(base + ws + 2 * b_noise);
  cudaStream_t st = (cudaStream_t)stream;
  PWV_CUDA(cudaMemcpyAsync(d_noise, noise, sizeof(float) * (size_t)N * T, cudaMemcpyHostToDevice, st));
  PWV_CUDA(cudaMemcpyAsync(d_mel, mel, sizeof(float) * (size_t)N * t_mel * m->hp.n_mels, cudaMemcpyHostToDevice, st));
  rc = pwv_forward(m, d_noise, d_mel, d_wav, base, ws, N, T, stream, nullptr);
  if (rc) return rc;
  PWV_CUDA(cudaMemcpyAsync(wav, d_wav, sizeof(float) * (size_t)N * T, cudaMemcpyDeviceToHost, st));
  PWV_CUDA(cudaStreamSynchronize(st));
  return PWV_OK;
}

int pwv_last_launch_count(const pwv_model* m) { return m ? m->last_launches : fail(PWV_EINVAL, "null model"); }

int pwv_debug_set_trace(pwv_model* m, long long* device_buffer, int launch_index) {
  if (!m) return fail(PWV_EINVAL, "null model");
  m->trace = device_buffer;
  m->trace_launch = launch_index;
  return PWV_OK;
}

int pwv_debug_set(pwv_model* m, const char* key, int value) {
  if (!m || !key) return fail(PWV_EINVAL, "null argument");
  const std::string k(key);
  if (k == "pdl") m->use_pdl = value != 0;
  else if (k == "tile_flags") m->use_flags = value != 0;
  else if (k == "path") { if (value != 0 && value != 1) return fail(PWV_EINVAL, "path must be 0 or 1"); m->tc_path = value; }
  else if (k == "flow") m->use_flow = value != 0;
  else if (k == "stagger") m->tc_stagger = value;
  else if (k == "rotate") m->tc_rotate = value != 0;
  else if (k == "seg") m->tc_seg = value;
  else if (k == "variant") { if (value < 0 || value > 2) return fail(PWV_EINVAL, "variant must be 0, 1 or 2"); m->tc_variant = value; }
  else if (k == "trace_flow") m->trace_flow = value != 0;
  else if (k == "cp") m->use_cp = value != 0;
  else if (k == "split1") m->split1 = value != 0;
  else if (k == "split2") m->split2 = value != 0;
  else if (k == "double_a") m->double_a = value != 0;
  else if (k == "z_in_d") m->z_in_d = value != 0;
  else return fail(PWV_EINVAL, "unknown debug switch '%s'", key);
  return PWV_OK;
}

int pwv_set_profiling(pwv_model* m, int enable) {
  if (!m) return fail(PWV_EINVAL, "null model");
  m->profiling = enable == 2 ? 2 : (enable != 0 ? 1 : 0);
  m->ev_used = 0;
  return PWV_OK;
}

int pwv_profile_read(pwv_model* m, double* layer_ms, int* layer_launches, double* forward_ms) {
  if (!m) return fail(PWV_EINVAL, "null model");
  if (!m->profiling || m->ev_used < 2) return fail(PWV_ESTATE, "profiling is off or no forward has run since it was enabled");
  PWV_CUDA(cudaEventSynchronize(m->ev[1]));
  float ms = 0.f;
  PWV_CUDA(cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]));
  if (forward_ms) *forward_ms = ms;
  double sum = 0.0;
  int n = 0;
  for (int i = 2; i + 1 < m->ev_used; i += 2) {
    PWV_CUDA(cudaEventElapsedTime(&ms, m->ev[i], m->ev[i + 1]));
    sum += ms;
    ++n;
  }
  if (layer_ms) *layer_ms = sum;
  if (layer_launches) *layer_launches = m->prof_launches > 0 ? m->prof_launches : n;
  return PWV_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// mel front end (reference data_load.py:37-56 / audio.py:327-356), see pwv_mel.cuh
// ------------------------------------------------------------------------------------------------
struct pwv_melspec {
  pwv_mel_config cfg;
  int bins = 0, win_lo = 0, win_hi = 0;
  float* d_window = nullptr;
  float2* d_twiddle = nullptr;
  float* d_basis = nullptr;
  int* d_band = nullptr;       // [2][n_mels]
  int* d_max = nullptr;
  int max_cap = 0;
};

extern "C" {

int pwv_melspec_create(const pwv_mel_config* cfg, const float* mel_basis, pwv_melspec** out) {
  if (!cfg || !mel_basis || !out) return fail(PWV_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->n_fft < 2 || (cfg->n_fft & (cfg->n_fft - 1)) != 0) return fail(PWV_EINVAL, "n_fft=%d must be a power of two", cfg->n_fft);
  if (cfg->win_length < 1 || cfg->win_length > cfg->n_fft) return fail(PWV_EINVAL, "win_length=%d must be in [1, n_fft=%d]", cfg->win_length, cfg->n_fft);
  if (cfg->hop_length < 1 || cfg->n_mels < 1) return fail(PWV_EINVAL, "bad hop_length/n_mels (%d/%d)", cfg->hop_length, cfg->n_mels);
  if (cfg->normalise && !(cfg->max_db > cfg->min_db)) return fail(PWV_EINVAL, "normalisation needs max_db > min_db");
  pwv_melspec* h = new (std::nothrow) pwv_melspec();
  if (!h) return fail(PWV_ENOMEM, "out of host memory");
  h->cfg = *cfg;
  const int n_fft = cfg->n_fft, win = cfg->win_length, n_mels = cfg->n_mels;
  h->bins = 1 + n_fft / 2;
  h->win_lo = (n_fft - win) / 2;            // the window is centred in the frame (torch.stft / librosa pad_center)
  h->win_hi = h->win_lo + win;
  std::vector<float> window(n_fft, 0.f);
  std::vector<float2> tw(n_fft);
  const double PI = 3.14159265358979323846;
  for (int i = 0; i < win; ++i) window[h->win_lo + i] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * i / win));   // periodic hann
  for (int i = 0; i < n_fft; ++i) tw[i] = make_float2((float)std::cos(2.0 * PI * i / n_fft), (float)(-std::sin(2.0 * PI * i / n_fft)));
  std::vector<int> band(2 * n_mels);
  for (int m = 0; m < n_mels; ++m) {
    int lo = h->bins, hi = 0;
    for (int k = 0; k < h->bins; ++k)
      if (mel_basis[(size_t)m * h->bins + k] != 0.f) { lo = k < lo ? k : lo; hi = k + 1; }
    if (hi <= lo) lo = hi = 0;
    band[m] = lo;
    band[n_mels + m] = hi;
  }
  auto up = [&](void** dst, const void* src, size_t bytes) {
    if (cudaMalloc(dst, bytes) != cudaSuccess) return false;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
  };
  const bool ok = up((void**)&h->d_window, window.data(), sizeof(float) * n_fft) && up((void**)&h->d_twiddle, tw.data(), sizeof(float2) * n_fft) &&
                  up((void**)&h->d_basis, mel_basis, sizeof(float) * (size_t)n_mels * h->bins) && up((void**)&h->d_band, band.data(), sizeof(int) * band.size());
  if (!ok) {
    cudaError_t e = cudaGetLastError();
    pwv_melspec_destroy(h);
    return fail(PWV_ECUDA, "mel front end: device upload failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return PWV_OK;
}

int pwv_melspec_destroy(pwv_melspec* h) {
  if (!h) return PWV_OK;
  cudaFree(h->d_window); cudaFree(h->d_twiddle); cudaFree(h->d_basis); cudaFree(h->d_band); cudaFree(h->d_max);
  delete h;
  return PWV_OK;
}

int pwv_melspec_forward(pwv_melspec* h, const float* wav, float* mel, int N, int T, pwv_stream stream) {
  if (!h || !wav || !mel) return fail(PWV_EINVAL, "null argument");
  if (N < 1 || T < 1 || N > 65535) return fail(PWV_EINVAL, "N=%d, T=%d out of range", N, T);
  if (T <= h->cfg.n_fft / 2) return fail(PWV_EINVAL, "T=%d: reflect padding needs more than n_fft/2 = %d samples", T, h->cfg.n_fft / 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (N > h->max_cap) {                 // per-utterance maxima: grown on demand (the only allocation of this entry point)
    if (h->d_max) cudaFree(h->d_max);
    h->d_max = nullptr;
    h->max_cap = 0;
    PWV_CUDA(cudaMalloc(&h->d_max, sizeof(int) * (size_t)N));
    h->max_cap = N;
  }
  pwv::MelParams p;
  p.wav = wav; p.out = mel; p.window = h->d_window; p.twiddle = h->d_twiddle; p.basis = h->d_basis;
  p.band_lo = h->d_band; p.band_hi = h->d_band + h->cfg.n_mels; p.utt_max = h->d_max;
  p.N = N; p.T = T; p.t_mel = 1 + T / h->cfg.hop_length; p.n_fft = h->cfg.n_fft; p.hop = h->cfg.hop_length; p.n_mels = h->cfg.n_mels;
  p.bins = h->bins; p.win_lo = h->win_lo; p.win_hi = h->win_hi;
  p.amin = 1e-5f; p.top_db = 80.f;      // librosa.amplitude_to_db defaults, what reference audio.py:254-262 relies on
  p.min_db = h->cfg.min_db; p.max_db = h->cfg.max_db; p.normalise = h->cfg.normalise;
  const int span = (pwv::MEL_FB - 1) * p.hop + p.n_fft;
  const size_t smem = sizeof(float) * ((span + 3) & ~3) + sizeof(float2) * p.n_fft + sizeof(float) * pwv::MEL_FB * (p.bins + 1);
  if (smem > 48 * 1024) return fail(PWV_EINVAL, "mel front end: n_fft=%d / hop=%d need %zu bytes of shared memory (> 48 KB)", p.n_fft, p.hop, smem);
  pwv::k_mel_init<<<(N + 255) / 256, 256, 0, st>>>(h->d_max, N);
  dim3 grid((p.t_mel + pwv::MEL_FB - 1) / pwv::MEL_FB, N);
  pwv::k_mel_power<<<grid, pwv::MEL_THREADS, smem, st>>>(p);
  const size_t total = (size_t)N * p.t_mel * p.n_mels;
  pwv::k_mel_finish<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
  PWV_CUDA(cudaGetLastError());
  return PWV_OK;
}

}  // extern "C"
