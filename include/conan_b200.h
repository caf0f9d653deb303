/*
 * conan_b200.h -- C ABI of the B200-native Conan chunkwise online-inference hot path.
 *
 * The reference (User-tian/Conan) is pure Python and has no FFI for this path; the
 * functions below are what a reference-side binding (ctypes)
 * would call instead of the PyTorch modules the streaming loop uses today.  Each
 * entry point names the reference interface it replaces (file:line relative to the
 * reference tree; TA = torchaudio/models/emformer.py, torchaudio 2.11).
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a CUDA device pointer owned by
 *     the caller, every pointer named *_host is host memory owned by the caller;
 *   - all calls are stream-ordered on `stream` (a cudaStream_t passed as void*) and do
 *     not synchronise unless the name ends in _host (those copy results back and wait);
 *   - return value 0 = success, non-zero = error; conan_last_error() returns a
 *     thread-local human readable message.  No C++ exception crosses this boundary;
 *   - a "slot" is the resident state of one voice stream (Emformer K/V ring + past
 *     length, Conan causal rings + per-session style cache, vocoder rings);
 *   - there is no CPU fallback: every function fails if no sm_100 device is present.
 */
#ifndef CONAN_B200_H_
#define CONAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CONAN_B200_ABI_VERSION 1

/* The library is built with -fvisibility=hidden: only the entry points declared here are exported. */
#if defined(__GNUC__)
#define CONAN_API __attribute__((visibility("default")))
#else
#define CONAN_API
#endif

typedef struct conan_engine conan_engine_t;

/* Hyper-parameters of the path.  Field names follow the reference's hparams keys
 * (egs/conan_emformer.yaml, egs/hifi_16k320_shuffle.yaml and their base chain). */
typedef struct conan_config {
  int32_t abi_version;            /* must be CONAN_B200_ABI_VERSION */
  int32_t device;                 /* CUDA device ordinal */
  int32_t max_slots;              /* number of resident stream slots */
  int32_t max_ref_frames;         /* longest reference-speech mel (frames) a session may use */
  /* Emformer (modules/Emformer/emformer.py:14-22) */
  int32_t emformer_layers;        /* 6 */
  int32_t emformer_dim;           /* 80  (input_dim) */
  int32_t emformer_heads;         /* 8 */
  int32_t emformer_ffn;           /* 2048 */
  int32_t segment;                /* chunk_size // 20 = 4 */
  int32_t right_context;          /* 2 */
  int32_t left_context;           /* 50 */
  int32_t emformer_output_dim;    /* 100 */
  /* Conan main model (modules/Conan/Conan.py:46-113) */
  int32_t hidden_size;            /* 256 */
  int32_t content_kernel;         /* kernel_size = 3 */
  int32_t dec_blocks;             /* len(dec_dilations) = 4 */
  int32_t dec_kernel;             /* dec_kernel_size = 5 */
  int32_t dec_post_kernel;        /* dec_post_net_kernel = 3 */
  int32_t predictor_kernel;       /* 5 */
  int32_t n_vq;                   /* nVQ = 512 */
  int32_t silent_token;           /* 57 */
  int32_t n_mels;                 /* audio_num_mel_bins = 80 */
  /* vocoder (modules/vocoder/hifigan/hifigan_causal.py:272-312) */
  int32_t voc_initial_channel;    /* upsample_initial_channel = 512 */
  int32_t voc_n_ups;              /* 4 */
  int32_t voc_rates[8];           /* upsample_rates        = 8,5,4,2 */
  int32_t voc_up_kernels[8];      /* upsample_kernel_sizes = 16,10,8,4 */
  int32_t voc_n_res;              /* 3 */
  int32_t voc_res_kernels[8];     /* resblock_kernel_sizes = 3,7,11 */
  int32_t voc_res_dilations[8];   /* resblock_dilation_sizes[*] = 1,3,5 (same for every kernel size) */
  int32_t voc_n_dil;              /* 3 */
  /* numerics / engine selection */
  int32_t voc_precision;          /* 0: fp32 operands (FFMA);  1: fp16 operands, fp32 accumulate (tcgen05 or FFMA);
                                     2: split-fp16 operands (x_hi*W_hi + x_hi*W_lo + x_lo*W_hi), fp32 accumulate, tcgen05 only:
                                        fp32-grade results (the reference arithmetic is fp32), needs voc_residual_from_ctx = 0 */
  int32_t voc_use_tensor_cores;   /* 1: tcgen05 implicit-GEMM kernels where eligible (needs voc_precision 1 or 2) */
  int32_t voc_group;              /* streams per vocoder pass (L2 blocking); 0 = all at once */
  int32_t voc_residual_from_ctx;  /* 1: the vocoder's resblock residual x_j is recovered from the activated copy lrelu(x_j) that is the
                                     next conv's input anyway (inverse LeakyReLU), so no separate fp32 residual stream is written or
                                     read; 0: keep an fp32 residual stream (reference-grade path) */
  int32_t lin_use_tensor_cores;   /* 1: Emformer / Conan linear + conv contractions on tcgen05 with split-fp16 operands
                                     (x_hi*W_hi + x_hi*W_lo + x_lo*W_hi, fp32 accumulate: fp32-grade results); 0: fp32 FFMA */
  int32_t voc_fuse_resblocks;     /* 1: at the scales with 32 / 64 channels a whole residual block (six convs) runs as one tcgen05 kernel
                                     with the activations kept in shared memory (needs tensor cores + voc_residual_from_ctx) */
  int32_t lin_fuse_ffn;           /* 1: the Emformer position-wise FFN (80 -> 2048 -> 80) runs as one tcgen05 kernel, the hidden
                                     activation stays in shared memory (needs lin_use_tensor_cores) */
  int32_t ses_use_tensor_cores;   /* 1: session setup runs the style encoder's ConvBlocks (conv k31 256 -> 512, 1x1 512 -> 256: 95 % of the
                                     12.8 GFLOP per session) on tcgen05 with split-fp16 operands; 0: fp32 FFMA */
  int32_t emformer_memory_size;   /* torchaudio Emformer max_memory_size M (0 in the reference config, modules/Emformer/emformer.py:14-22):
                                     M > 0 keeps a bank of the last M memory vectors per layer and stream and adds the summary query */
  int32_t step_graphs;            /* 1: conan_step / conan_step_host* replay a CUDA graph of the whole chunk step (captured once per distinct
                                     ready count and buffer set; the slot ids stay an indirection read from the device buffer) */
  int32_t lin_fuse_blocks;        /* 1: the decoder's residual block body (conv k5 -> GELU -> 1x1) and the aligner's feed-forward run as one
                                     tcgen05 kernel each, hidden activation in shared memory (needs lin_use_tensor_cores) */
} conan_config_t;

/* dtype codes for conan_engine_bind_weight */
#define CONAN_DTYPE_F32 0
#define CONAN_DTYPE_F16 1
#define CONAN_DTYPE_I32 2

CONAN_API const char* conan_last_error(void);
CONAN_API int conan_abi_version(void);
/* sizeof(conan_config_t) / sizeof(conan_conv_params_t) as compiled, so a foreign-language binding can
 * verify its struct layout before the first call */
CONAN_API size_t conan_sizeof_config(void);
CONAN_API size_t conan_sizeof_conv_params(void);

/* Replaces StreamingVoiceConversion.__init__ / _build_model / _build_vocoder /
 * _build_emformer (inference/Conan.py:26-52): create, bind every tensor the path
 * needs (names listed by conan_engine_weight_name), then finalize. */
CONAN_API int conan_engine_create(const conan_config_t* cfg, conan_engine_t** out);
CONAN_API void conan_engine_destroy(conan_engine_t* eng);
CONAN_API int conan_engine_num_weights(const conan_engine_t* eng);
/* name / expected element count / dtype of weight #idx (for the host-side packer) */
CONAN_API int conan_engine_weight_info(const conan_engine_t* eng, int idx, const char** name, size_t* numel, int* dtype);
/* The engine keeps the pointer (no copy); the caller keeps the allocation alive. */
CONAN_API int conan_engine_bind_weight(conan_engine_t* eng, const char* name, const void* data_dev, size_t numel, int dtype);
CONAN_API int conan_engine_finalize(conan_engine_t* eng);
/* bytes of device memory held by the state slab + scratch */
CONAN_API size_t conan_engine_state_bytes(const conan_engine_t* eng);

/* Zero the resident state of `n` slots (stream start).  parts: bit0 Emformer, bit1 Conan
 * rings, bit2 vocoder rings.  Replaces `state = None` (inference/Conan.py:92) and the
 * zero left-padding every causal conv of the reference starts from.  Stream-ordered; slots_host is read before the call
 * returns when it is pageable memory (page-locked memory must stay valid until the stream reaches this call). */
CONAN_API int conan_slots_reset(conan_engine_t* eng, int n, const int32_t* slots_host, int parts, void* stream);

/* Once per session: the reference-speech branch of Conan.forward
 * (modules/Conan/Conan.py:157-159,200-219 encode_spk_embed; :221-249 get_prosody up to
 * the aligner's keys; modules/Conan/prosody_util.py:183-200 LocalStyleAdaptor) for `n`
 * sessions whose reference mels all have `ref_frames` frames.  Caches style_embed and
 * the aligner K/V per slot.  ref_mel_dev: [n, ref_frames, n_mels] fp32.  Stream-ordered, does not synchronise: ref_mel_dev
 * must stay valid until the stream has passed this call (slots_host as for conan_slots_reset). */
CONAN_API int conan_session_open(conan_engine_t* eng, int n, const int32_t* slots_host, const float* ref_mel_dev,
                       int ref_frames, void* stream);

/* One streaming step of torchaudio Emformer.infer + proj + argmax for n streams
 * (inference/Conan.py:113-127, TA:745-803).  chunk_dev: [n, segment+right_context, dim]
 * laid out as the reference passes it (utterance rows first, look-ahead rows last).
 * Any of the outputs may be NULL.  enc [n,segment,dim], logits [n,segment,out_dim],
 * tokens [n,segment] int32. */
CONAN_API int conan_emformer_step(conan_engine_t* eng, int n, const int32_t* slot_ids_dev, const float* chunk_dev,
                        float* enc_out_dev, float* logits_out_dev, int32_t* tokens_out_dev, void* stream);

/* Full-utterance EmformerDistillModel.forward (modules/Emformer/emformer.py:31-47 -> torchaudio Emformer.forward, TA:709-743):
 * input_dev [n, frames, dim] is the utterance right-padded with right_context frames, so frames - right_context utterance frames
 * are encoded.  Runs as ceil((frames - rc) / segment) streaming steps over the slots' state (which is reset first) -- the block
 * attention mask of the reference's forward expresses exactly the per-segment visibility of the streaming steps; a partial last
 * segment and the memory bank are handled by the generic step.  enc [n, frames - rc, dim], logits [n, frames - rc, out_dim],
 * tokens [n, frames - rc] int32; any output may be NULL. */
CONAN_API int conan_emformer_forward(conan_engine_t* eng, int n, const int32_t* slots_host, const float* input_dev, int frames,
                                     float* enc_out_dev, float* logits_out_dev, int32_t* tokens_out_dev, void* stream);

/* Incremental Conan.forward(infer=True) on the newest `segment` tokens of each stream
 * (inference/Conan.py:131-145; modules/Conan/Conan.py:115-198).  tokens [n,segment] int32,
 * mel_out [n,segment,n_mels] fp32. */
CONAN_API int conan_decoder_step(conan_engine_t* eng, int n, const int32_t* slot_ids_dev, const int32_t* tokens_dev,
                       float* mel_out_dev, void* stream);

/* Incremental HifiGanGenerator.forward on the newest `segment` mel frames
 * (tasks/tts/vocoder_infer/hifigan.py:23-31, hifigan_causal.py:314-333).
 * mel [n,segment,n_mels] fp32 -> wav [n, segment*hop] fp32. */
CONAN_API int conan_vocoder_step(conan_engine_t* eng, int n, const int32_t* slot_ids_dev, const float* mel_dev,
                       float* wav_out_dev, void* stream);

/* The whole chunk step (one iteration of the loop at inference/Conan.py:95-156) for n
 * ready streams packed into one launch sequence.  Outputs may be NULL except wav. */
CONAN_API int conan_step(conan_engine_t* eng, int n, const int32_t* slot_ids_dev, const float* chunk_dev,
               float* wav_out_dev, float* mel_out_dev, int32_t* tokens_out_dev, void* stream);

/* Same, with HOST buffers (pinned or pageable): copies slot ids + mel chunks to the
 * device, runs the step, copies wav (and mel/tokens if non-NULL) back, and waits.
 * This is the call the reference-facing plugin makes per chunk step. */
CONAN_API int conan_step_host(conan_engine_t* eng, int n, const int32_t* slot_ids_host, const float* chunk_host,
                    float* wav_out_host, float* mel_out_host, int32_t* tokens_out_host, void* stream);

/* Pipelined variant for a serving loop: submit enqueues the input copies and the step on `stream`, the result copies on an
 * engine-owned copy stream, and returns a ticket; the result copies of step i overlap the compute of step i+1.  At most two
 * steps may be in flight.  Every host buffer passed to a submit (pinned memory) must stay valid and untouched until its ticket
 * has been waited for: use two sets of chunk / result buffers, alternating. */
CONAN_API int conan_step_host_submit(conan_engine_t* eng, int n, const int32_t* slot_ids_host, const float* chunk_host, float* wav_out_host,
                           float* mel_out_host, int32_t* tokens_out_host, void* stream, int* ticket);
CONAN_API int conan_step_host_wait(conan_engine_t* eng, int ticket);

/* kernels launched by this engine since creation (bench.py's gpu_launches claim) */
CONAN_API uint64_t conan_engine_launch_count(const conan_engine_t* eng);

/* chunk steps served by replaying a captured CUDA graph since creation (0 when step_graphs is off) */
CONAN_API uint64_t conan_engine_graph_replays(const conan_engine_t* eng);

/* Per-launch CUDA-event timing of the conv engines (measurement only: events are recorded on the
 * launching stream around every conv launch while enabled).  category 0 = FFMA, 1 = tcgen05 ring kernel
 * (fp16 operands), 2 = tcgen05 window kernel, 3 = tcgen05 ring kernel with split-fp16 operands, 4 = fused
 * residual-block kernel (six convs per launch), 5 = fused Emformer feed-forward kernel, 6 = fused Conan block kernel
 * (decoder residual block body / aligner feed-forward: two GEMMs per launch).
 * profile_read synchronises the device and returns the summed kernel time, launch count, algorithmic
 * FLOPs (2*M*N*K) and algorithmic HBM bytes since profiling was (re-)enabled. */
CONAN_API int conan_engine_set_profiling(conan_engine_t* eng, int enabled);
CONAN_API int conan_engine_profile_read(conan_engine_t* eng, int category, double* ms, uint64_t* launches, double* flops, double* bytes);

/* Debug/test access: copy a named internal per-slot tensor (fp32) of one slot to the
 * device buffer.  Returns the element count through *numel. Names: "style", "kv_cache",
 * "kpm", "emformer_past_len", "vq_index". */
CONAN_API int conan_debug_read(conan_engine_t* eng, const char* name, int slot, float* out_dev, size_t capacity,
                     size_t* numel, void* stream);

/* ------------------------------------------------------------------------------------
 * Stand-alone operator: implicit-GEMM 1-D convolution over per-slot context buffers with
 * the fused epilogue used everywhere on the path (tests drive both engines through it).
 *   out[i,t,n] = epilogue( bias[n] + sum_{j<k} sum_{c<Cin} X[slot_i, row0 + t + j*dil, c] * W[n, j*Cin + c] )
 * ---------------------------------------------------------------------------------- */
typedef struct conan_conv_params {
  const void* x;            /* context buffer, fp32 or fp16, [slots, x_rows, cin] */
  int64_t x_slot_stride;    /* elements */
  int32_t x_row_stride;     /* elements (>= cin) */
  int32_t x_rows;           /* rows per slot in the buffer (TMA bound) */
  int32_t x_is_half;
  int32_t row0;             /* first tap row for output t = 0 */
  int32_t L;                /* output rows per stream */
  int32_t cin, k, dil, cout;
  const void* w;            /* packed [cout, k*cin], same dtype as x */
  const float* bias;        /* [cout] or NULL */
  int32_t n_streams;
  const int32_t* slot_ids;  /* device, [n_streams]; NULL = identity */
  int32_t n_slots;          /* slots in the buffers (TMA bound) */
  float scale;              /* v = (acc + bias) * scale */
  int32_t act;              /* 0 none 1 relu 2 leaky(slope) 3 gelu(erf) 4 tanh */
  float slope;
  const float* res;         /* v += res[slot*res_slot_stride + t*res_row_stride + n]  (NULL = none) */
  int64_t res_slot_stride;
  int32_t res_row_stride;
  const float* rowmask;     /* v *= rowmask[slot*mask_slot_stride + t] (NULL = none) */
  int32_t mask_slot_stride;
  float out_scale;          /* v *= out_scale */
  float* y;                 /* fp32 output (NULL = none): y[slot*y_slot_stride + (y_row0+t)*y_row_stride + n] */
  int64_t y_slot_stride;
  int32_t y_row_stride, y_row0;
  int32_t accumulate;       /* v += old y before storing */
  void* y2;                 /* second output = act2(v), fp32 or fp16 (NULL = none) */
  int64_t y2_slot_stride;
  int32_t y2_row_stride, y2_row0;
  int32_t y2_is_half;
  int32_t act2;
  float slope2;
  /* fp32-grade tensor-core mode (tcgen05 engine only): operands are split fp16 pairs v = hi + lo.
   * x_split = 1: the hi plane of x is slot s, the lo plane slot s + x_lo_slot_off of the same buffer, and w is
   * packed [cout, 3*k*cin] = [W_hi | W_lo | W_hi] (optionally pre-scaled by 1/acc_scale); the kernel accumulates
   * x_hi*W_hi + x_hi*W_lo + x_lo*W_hi in fp32.  y2_split = 1: y2 is written as such a pair (lo at +y2_lo_off elements). */
  int32_t x_split;
  int64_t x_lo_slot_off;
  float acc_scale;          /* v = (acc * acc_scale + bias) * scale ; 0 means 1 */
  int32_t y2_split;
  int64_t y2_lo_off;
  /* residual read from an activated fp16/fp32 context buffer instead of an fp32 stream: res_is_half selects the
   * element type of `res`; res_inv_slope != 0 undoes the LeakyReLU the producer applied (r = h >= 0 ? h : h * res_inv_slope). */
  int32_t res_is_half;
  float res_inv_slope;
  /* second residual (fp16 or fp32, added as is) and an fp16 primary output: the running MRF sum of the vocoder's
   * three resblocks is carried as an fp16 tensor -- v += res2[slot*res2_slot_stride + t*res2_row_stride + n]. */
  const void* res2;
  int64_t res2_slot_stride;
  int32_t res2_row_stride;
  int32_t res2_is_half;
  int32_t y_is_half;        /* y is __half* (no accumulate in that case: use res2) */
} conan_conv_params_t;

/* Log-mel front-end on the device (replaces the host librosa path of inference/Conan.py:58-70 and
 * utils/audio/__init__.py:36-80 for streamed PCM).  wav_rows: the centre-padded signal of every stream as rows of `hop`
 * samples, [n_streams][rows_per_stream][hop] fp32 (fft_size/2 zeros in front, zeros behind); frame f covers rows
 * row0 + f .. row0 + f + taps - 1 (taps = ceil(fft_size / hop); the basis is zero beyond fft_size).  dft_w: window-weighted
 * DFT basis [2*bins][taps*hop] fp32 (rows 0..bins-1 cos, bins..2*bins-1 -sin); mel_basis_t: [bins][n_mels] fp32.
 * mel_out [n_streams][n_frames][n_mels] = clip(log10(max(mel_basis . |DFT|, eps)), vmin, vmax).
 * spec_scratch: n_streams*n_frames*2*bins floats. */
CONAN_API int conan_logmel(const float* wav_rows, int n_streams, int rows_per_stream, int hop, int taps, int row0, int n_frames,
                 const float* dft_w, int bins, const float* mel_basis_t, int n_mels, float eps, float vmin, float vmax,
                 float* spec_scratch, float* mel_out, void* stream);

/* engine: 0 = FFMA (fp32 accumulate on CUDA cores, fp32 or fp16 operands),
 *         1 = tcgen05 tensor cores (fp16 operands, fp32 accumulate in TMEM). */
CONAN_API int conan_conv_gemm(const conan_conv_params_t* p, int engine, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CONAN_B200_H_ */
