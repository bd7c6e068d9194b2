/*
 * pwv.h -- C-ABI of the B200-native IAF-vocoder generation path.
 *
 * The reference (andabi/parallel-wavenet-vocoder @ 6c2fa069) has no FFI: its boundary for this
 * path is Python-level -- `IAFVocoder(batch_size, length)(wav, melspec, is_training)` building a
 * TF graph (reference models.py:18-78) that one `sess.run` executes (reference generate.py:68).
 * This header is what a binding for that path would bind instead of the TF session:
 *
 *   reference interface (file:line)                         replaced by
 *   ------------------------------------------------------  --------------------------------------
 *   hp.model.* / hp.signal.* read at graph build             pwv_hparams + pwv_model_create
 *     (models.py:20,26-70,106-133)
 *   tf.get_variable(name, shape) per variable                pwv_model_load_weight (same names, same
 *     (modules.py:152-164,179,210-248; models.py:128)          [k,Cin,Cout] layouts)
 *   tf.train.Saver(...).restore / global_variables_init      pwv_model_finalize (packs + uploads)
 *     (generate.py:55-66)
 *   sess.run(pred_wav_op)  (generate.py:68) ==               pwv_forward (device buffers, async) /
 *     IAFVocoder.__call__ (models.py:23-78) ->                 pwv_forward_host (host buffers, sync)
 *     LinearIAFLayer (modules.py:53-60) ->
 *     WaveNet.__call__ (modules.py:129-166) ->
 *     causal_conv (modules.py:11-43)
 *   Logistic(0,1).sample in-graph (models.py:32-33)          the caller passes `noise` explicitly
 *
 * Conventions: plain C, no C++/torch types. Every function returns 0 on success or a negative
 * PWV_E* code and never throws, exits or prints; pwv_last_error() returns a thread-local message
 * for the last failure on the calling thread. Buffers are contiguous row-major float32:
 * noise/wav [N][T], mel [N][1+T/hop][n_mels]. The library owns only the packed weights; the caller
 * owns inputs, outputs and workspace. pwv_forward enqueues on `stream` and returns without
 * synchronising; it allocates nothing. Threading: ONE forward in flight per model -- pwv_forward records its launch
 * count and profiling events in the model and the kernels of consecutive layers hand tiles to each other through
 * per-tile flags in the caller's workspace; use one model (the packed weights are 20 MB) and one workspace per host
 * thread / stream. Several models may run concurrently on one device: every wait inside the kernels targets a kernel
 * launched earlier on the same stream, never a CTA of the same grid, so no co-residency is assumed (the round-1
 * whole-flow kernel, which did assume it, is reachable only through pwv_debug_set("path", 0)).
 * Range: the tensor-core precisions hold activations as fp16 hi + lo (bf16 in PWV_PREC_BF16): |activation| > 65504
 * overflows where the reference's fp32 would not -- not reachable with trained or Glorot-initialised weights, but a
 * caller with exotic checkpoints should use PWV_PREC_FP32. There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with PWV_ECUDA.
 */
#ifndef PWV_H_
#define PWV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PWV_VERSION 201          /* major*10000 + minor*100 + patch */
#define PWV_MAX_FLOWS 8
#define PWV_MAX_LAYERS 64
#define PWV_MAX_UPSAMPLE 4
#define PWV_MAX_FILTER_WIDTH 32

/* mel -> sample-rate conditioning (reference models.py:105-136) */
#define PWV_UPSAMPLE_REPEAT 0            /* 1x1 conv + relu, every frame repeated hop times (default)       */
#define PWV_UPSAMPLE_TRANSPOSED_CONV 1   /* stacked conv2d_transpose (kernel = stride) + relu per stage      */
#define PWV_UPSAMPLE_NONE 2              /* any other method: no conditioning at all (models.py:134-135)   */
#define PWV_NORM_NONE 0
#define PWV_NORM_IN 1

/* error codes */
#define PWV_OK 0
#define PWV_EINVAL (-1)          /* bad argument / unsupported hyper-parameter                  */
#define PWV_ESTATE (-2)          /* call sequence (e.g. forward before finalize)                */
#define PWV_ECUDA (-3)           /* CUDA runtime / driver error (message has the CUDA string)   */
#define PWV_ENOMEM (-4)          /* workspace too small / allocation failed                     */
#define PWV_ENAME (-5)           /* unknown variable name or wrong shape in load_weight         */

/* arithmetic of the gated-layer contractions */
#define PWV_PREC_FP32 0          /* fp32 FFMA on CUDA cores; bit-for-bit IEEE fp32 accumulate    */
#define PWV_PREC_F16X3 1         /* tcgen05 kind::f16, fp16 hi/lo 3-term split, fp32 accumulate: */
                                 /*   22-bit operands, fp32-level parity                        */
#define PWV_PREC_BF16 2          /* tcgen05 kind::f16, single pass on bf16 operands, fp32 acc.  */

typedef struct pwv_model pwv_model;   /* opaque */
typedef void* pwv_stream;             /* a cudaStream_t / CUstream (0 = legacy default stream) */

/* POD mirror of the hparams the path reads (reference hparams/default.yaml:11-33). */
typedef struct pwv_hparams {
  int32_t n_iaf;                  /* model.n_iaf                                              */
  /* Shapes: the fused kernels (all three precisions) cover filter_width 2, R = D in {64,128,256}, S = 2R -- the
   * reference's defaults and BASELINE's channel sweep. Any other positive values (free parameters in the reference,
   * modules.py:210-244) run an un-fused general chain that needs PWV_PREC_FP32. */
  int32_t filter_width;           /* model.filter_width (1 .. PWV_MAX_FILTER_WIDTH)            */
  int32_t residual_channels;      /* model.residual_channels (R)                               */
  int32_t dilation_channels;      /* model.dilation_channels (D)                               */
  int32_t skip_channels;          /* model.skip_channels (S)                                   */
  int32_t condition_channels;     /* model.condition_channels (Cc)                             */
  int32_t n_mels;                 /* signal.n_mels                                             */
  int32_t hop_length;             /* signal.hop_length                                         */
  int32_t use_biases;             /* model.use_biases                                          */
  int32_t use_skip_connection;    /* model.use_skip_connection                                    */
  int32_t precision;              /* PWV_PREC_*                                                */
  int32_t n_layers[PWV_MAX_FLOWS];                     /* len(model.dilations[i])              */
  int32_t dilations[PWV_MAX_FLOWS][PWV_MAX_LAYERS];    /* model.dilations[i][j]                */
  int32_t cond_upsample;          /* model.cond_upsample_method: PWV_UPSAMPLE_REPEAT | _TRANSPOSED_CONV | _NONE */
  int32_t n_upsample;             /* transposed_conv: number of stages (reference models.py:23: 3) ...   */
  int32_t upsample_strides[PWV_MAX_UPSAMPLE];  /* ... and their strides ([4,4,5]); product == hop_length  */
  /* normalisers (reference modules.py:263-284): PWV_NORM_NONE ('' / None) or PWV_NORM_IN ('in': instance normalisation
   * over time); 'bn' is tf.layers.batch_normalization and is not implemented. Any normaliser needs PWV_PREC_FP32. */
  int32_t normalize;              /* model.normalize: x after every flow (models.py:70)                           */
  int32_t normalize_cond;         /* model.normalize_cond: the conditioning (models.py:27-29,121-122)             */
  int32_t normalize_wavenet;      /* model.normalize_wavenet: inside the WaveNet bodies (modules.py:149-257)      */
} pwv_hparams;

/* Optional debug taps for parity tests (all device pointers, any may be NULL). */
typedef struct pwv_taps {
  float* flow_out;                /* [n_iaf][N][T]: x after each flow (reference models.py:67)  */
  int32_t layer_flow, layer_body, layer_index;  /* which gated layer to capture ...            */
  float* layer_out;               /* ... [N][T][R]: its dense_output (reference modules.py:251) */
  float* scale_shift;             /* [n_iaf][2][N][T]: WaveNet outputs (reference modules.py:56-57) */
} pwv_taps;

int pwv_version(void);
const char* pwv_last_error(void);
/* Number of CUDA devices visible, or a negative PWV_ECUDA. */
int pwv_device_count(void);

/* Validate hparams and create an empty model on the current CUDA device. */
int pwv_model_create(const pwv_hparams* hp, pwv_model** out);
int pwv_model_destroy(pwv_model* m);

/* Number of variables the model expects, and the i-th one's TF name / shape (ndim <= 4). */
int pwv_model_num_variables(const pwv_model* m);
int pwv_model_variable(const pwv_model* m, int index, const char** name, int64_t shape[4], int* ndim);

/* Copy one variable (HOST pointer, float32, TF layout) into the model's staging area. */
int pwv_model_load_weight(pwv_model* m, const char* tf_name, const float* host_data,
                          const int64_t* shape, int ndim);
/* Check that every variable was loaded, repack into the kernels' device layouts, upload. */
int pwv_model_finalize(pwv_model* m);

/* Bytes of device workspace pwv_forward needs for a batch of N utterances of T samples. */
int pwv_workspace_bytes(const pwv_model* m, int N, int T, size_t* bytes);

/* noise [N][T], mel [N][1+T/hop][n_mels] -> wav [N][T]; all DEVICE pointers; T % hop == 0.
 * Asynchronous on `stream`. `taps` may be NULL. */
int pwv_forward(pwv_model* m, const float* noise, const float* mel, float* wav,
                void* workspace, size_t workspace_bytes, int N, int T,
                pwv_stream stream, const pwv_taps* taps);

/* Same computation with HOST buffers (pinned or pageable): H2D copies, pwv_forward, D2H copy and
 * a stream synchronise inside the call. Device staging is cached inside the model and grown on
 * demand (the only entry point that allocates). */
int pwv_forward_host(pwv_model* m, const float* noise, const float* mel, float* wav,
                     int N, int T, pwv_stream stream);

/* Kernel launches enqueued by the most recent pwv_forward on this model (for bench accounting). */
int pwv_last_launch_count(const pwv_model* m);

/* Per-kernel timing for the roofline report. enable = 1: pwv_forward brackets EVERY gated-layer kernel
 * launch (the dominant kernel) with CUDA events on `stream`; the events serialise the launches, so the
 * programmatic-dependent-launch overlap of production runs is off and each launch is timed in isolation.
 * enable = 2: one event pair around each flow's chain of gated-layer launches, launched exactly as in
 * production (overlap on); the average launch duration is chain time / launches. pwv_profile_read waits
 * for the most recent forward and returns the summed device time of the bracketed launches, their count,
 * and the device time of the whole forward. 0 switches profiling off. */
int pwv_set_profiling(pwv_model* m, int enable);
/* Debug: device buffer of 4*16*16 int64 that CTA 0 fills with clock64() stamps of its phases while it
 * runs ONE gated layer on the tensor cores (NULL switches it off). `layer_index` counts the gated
 * layers of a forward from 0 with the flows concatenated (default hparams: 0..59). */
int pwv_debug_set_trace(pwv_model* m, long long* device_buffer, int layer_index);
int pwv_profile_read(pwv_model* m, double* layer_ms, int* layer_launches, double* forward_ms);
/* Debug / A-B switches for tests and measurement tools (the library never reads the environment): key one of
 * "path" (1 = 16-bit activation planes, the default; 0 = the round-1 fp32-row kernels), "pdl", "tile_flags", "flow",
 * "seg", "rotate", "stagger", "variant", "trace_flow", "split1", "split2", "cp", "double_a", "z_in_d" -- see csrc/pwv_api.cu. Unknown keys fail with PWV_EINVAL. */
int pwv_debug_set(pwv_model* m, const char* key, int value);

/* ---- mel front end: the step immediately upstream of the path.
 * Replaces reference audio.py:327-356 `wav2melspec_db` (librosa.stft -> |.| -> librosa.filters.mel -> amplitude_to_db
 * -> normalize_db, called from data_load.py:37-56): centred STFT (periodic hann of win_length centred in n_fft,
 * reflect padding), magnitude, mel basis, 20 log10(max(., 1e-5)), floor at (utterance max - 80 dB), and, when
 * `normalise`, clip((db - min_db) / (max_db - min_db), 0, 1) * 2 - 1. The caller supplies the mel basis
 * [n_mels][1 + n_fft/2] (HOST pointer; librosa.filters.mel(sr, n_fft, n_mels): melspec.mel_basis). */
typedef struct pwv_mel_config {
  int32_t n_fft, win_length, hop_length, n_mels;   /* signal.n_fft / win_length / hop_length / n_mels; n_fft a power of two */
  float min_db, max_db;                            /* signal.min_db / max_db                                              */
  int32_t normalise;                               /* 0: raw dB (after the top_db floor)                                    */
} pwv_mel_config;
typedef struct pwv_melspec pwv_melspec;            /* opaque */
int pwv_melspec_create(const pwv_mel_config* cfg, const float* mel_basis, pwv_melspec** out);
int pwv_melspec_destroy(pwv_melspec* h);
/* wav [N][T] -> mel [N][1 + T/hop][n_mels], DEVICE pointers, asynchronous on `stream`; T > n_fft/2. */
int pwv_melspec_forward(pwv_melspec* h, const float* wav, float* mel, int N, int T, pwv_stream stream);

#ifdef __cplusplus
}
#endif
#endif  /* PWV_H_ */
