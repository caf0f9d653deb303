// Launchers of the non-GEMM kernels of the path (norm / softmax / gather / ring kernels).
// All of them are HBM/L2-bound, coalesced, vectorised where the layout allows, and use
// warp-shuffle reductions; none of them goes near the tensor cores.
#pragma once
#include "common.cuh"

namespace conan {

// A strided view of per-slot rows: element (slot, t, c) lives at
//   base[slot * slot_stride + (row0 + t) * row_stride + c].
//   is_half: 0 = fp32, 1 = fp16, 2 = split fp16 pair: hi = fp16(v) at base, lo = fp16(v - hi) at base + lo_off
//   (the operand format of the fp32-grade tensor-core GEMMs, see conv_gemm_tc.cu)
struct RowView {
  void* base = nullptr;
  long long slot_stride = 0;
  int row_stride = 0;
  int row0 = 0;
  int is_half = 0;
  long long lo_off = 0;
};

__device__ __forceinline__ void store_view(const RowView& v, long long off, float x) {
  if (v.is_half == 0) { reinterpret_cast<float*>(v.base)[off] = x; return; }
  __half h = __float2half_rn(x);
  reinterpret_cast<__half*>(v.base)[off] = h;
  if (v.is_half == 2) reinterpret_cast<__half*>(v.base)[off + v.lo_off] = __float2half_rn(x - __half2float(h));
}
inline RowView view_f32(float* p, long long ss, int rs, int r0 = 0) { return RowView{p, ss, rs, r0, 0}; }

struct LnArgs {
  RowView in;            // fp32
  RowView out;           // fp32, fp16 or split fp16
  RowView out2;          // optional second copy (base == nullptr: none), e.g. fp32 residual next to a split GEMM operand
  const float* gamma; const float* beta; float eps;
  int C, L, n;
  const int* slot_ids;
  const float* premask; int premask_slot_stride;    // x *= premask[slot, t] before the statistics
  const float* postmask; int postmask_slot_stride;  // y *= postmask[slot, t]
  float* write_mask; int write_mask_slot_stride;     // mask[slot, t] = (sum_c |x| > 0) of the raw input
  float* write_mask2;                                // optional second copy (same stride)
  // optional: the input row is  sum_p part[p][row][c] + part_bias[c] + part_res[row][c]  instead of `in` (the fused FFN kernel
  // and the fused residual-block kernel leave one partial per hidden-dimension slice; compact rows, C <= 256)
  const float* part = nullptr; int n_part = 0; long long part_stride = 0; int part_ld = 0;
  const float* part_bias = nullptr; const float* part_res = nullptr; int part_res_ld = 0;
  // (C <= 256) optionally the assembled row is multiplied by part_mask[row] (the residual block's nonpadding mask) and written
  // back as the new fp32 residual stream part_out[row][c] before it is normalised
  const float* part_mask = nullptr; float* part_out = nullptr; int part_out_ld = 0;
  // optional chained second LayerNorm over the first one's output y (C <= 128): out3 = LN2(y) -- the next Emformer layer's
  // input norm applied in the same pass as this layer's output norm
  const float* gamma2 = nullptr; const float* beta2 = nullptr; RowView out3;
};
int launch_layernorm(const LnArgs& a, cudaStream_t st);

// Emformer ------------------------------------------------------------------------------
// chunk [n, seg+rc, D] (utterance rows first) -> X[slot] rows ordered [rc | utt] (TA:430)
int launch_emformer_assemble(const float* chunk, float* X, int ldx, int n, const int* slot_ids, int seg, int rc, int D, cudaStream_t st);
// per stream: append utterance K/V rows to the ring, softmax(QK^T) V over [rc | left ctx | utt].
// qkv / att are compact (index i, row stride ld); kv_ring / past_len are resident state indexed by slot_ids[i].
// optional fused tail of the attention kernel: r1 = att . Wout^T + b + x_res (fp32, TA:416-425), fn = LayerNorm(r1) (the FFN's
// input norm) written as a GEMM operand; compact rows [i*rows + t], row stride ld
struct EmfAttnEpilogue {
  const float* wt; const float* bias;        // out_proj weight transposed [D(k)][D(c)] and bias [D], fp32
  const float* x_res; int ld; float* r1;     // residual in / out
  const float* ln_g; const float* ln_b; float eps; RowView fn;
};
int launch_emformer_attention(const float* qkv, float* kv_ring, const int* past_len, RowView att, int n,
                              const int* slot_ids, int seg, int rc, int lc, int ring_rows, int D, int heads, int ld_qkv,
                              cudaStream_t st, const EmfAttnEpilogue* ep = nullptr);
int launch_advance_past_len(int* past_len, int n, const int* slot_ids, int seg, cudaStream_t st);
// generic step (emformer_mem.cu): memory bank M > 0, summary query, partial segments.  Work buffers hold seg + rc + 1 rows
// per stream: [rc | utt | summary].
int launch_emformer_assemble_generic(const float* src, long long stream_stride, int utt_row0, int rc_row0, float* X, int ldx,
                                     float* mem0, int n, int seg, int n_utt, int rc, int D, cudaStream_t st);
int launch_emformer_mem_prepare(const float* xnf, int ld, RowView xn, const float* bank, RowView mb, int n, const int* slot_ids,
                                int seg, int n_utt, int rc, int D, int M, cudaStream_t st);
int launch_emformer_mem_update(const float* r1, int ld, const float* mem_in, float* mem_out, float* bank, int n, const int* slot_ids,
                               int seg, int rc, int D, int M, cudaStream_t st);
int launch_emformer_attention_mem(const float* qkv, const float* memkv, float* ring, const int* past_len, RowView att, int n,
                                  const int* slot_ids, int seg, int n_utt, int rc, int lc, int ring_rows, int D, int heads, int ldq,
                                  int M, cudaStream_t st);
int launch_advance_past_len_by(int* past_len, int n, const int* slot_ids, int by, cudaStream_t st);
int launch_copy_rows_strided(const float* src, long long src_slot_stride, int src_ld, int src_row0, float* dst, long long dst_slot_stride,
                             int dst_ld, int dst_row0, int n, int rows, int C, cudaStream_t st);
int launch_argmax_rows(const float* logits, int ld, int* tokens_a, int* tokens_b, int n, int rows, int C, cudaStream_t st);
int launch_copy_rows_out(const float* src_slot, long long slot_stride, int row_stride, int row0, float* dst, int n,
                         const int* slot_ids, int rows, int C, cudaStream_t st);
int launch_copy_rows_in(const void* src, int src_is_int, void* dst_slot, long long slot_stride_elems, int n,
                        const int* slot_ids, int elems, cudaStream_t st);

// Conan chunk path -------------------------------------------------------------------------
int launch_embedding_rows(const int* tokens_slot, const float* table, int vocab, RowView out, int n, const int* slot_ids,
                          int rows, int C, cudaStream_t st);
// nn.MultiheadAttention(256, 2) over the session-cached K/V (prosody_util.py:108-127)
// q / out compact; kv_cache, kpm, n_keys are session state indexed by slot_ids[i]
int launch_cross_attention(const float* q, const float* kv_cache, const float* kpm, const int* n_keys, RowView out, int n,
                           const int* slot_ids, int rows, int H, int heads, int layer, int n_layers, int tp_max, cudaStream_t st);
// out1 = a + b (fp32), optional second copy into a context buffer
int launch_add_rows(const float* a, const float* b, float* out1, RowView out2, int n, const int* slot_ids, int rows, int C, cudaStream_t st);
// uv_predictor tail + pitch embedding (nar_tts_modules.py:142-146, Conan.py:324-351, pitch/utils.py:17-28,71-82)
int launch_pitch(const float* h, const float* ln_g, const float* ln_b, const float* lin_w, const float* lin_b,
                 const int* tokens_slot, int silent_token, const float* pitch_table, const float* pitch_inp,
                 float* dec_inp, float* uv_pred_out, int n, const int* slot_ids, int rows, int Cuv, int H, cudaStream_t st);

// fp32 rows [slot, rows, C] -> a (possibly fp16) context-buffer view
int launch_rows_to_view(const float* src_slot, RowView out, int n, const int* slot_ids, int rows, int C, cudaStream_t st);

// vocoder ------------------------------------------------------------------------------------
// taps_host: optional HOST copy of [k*C weights | bias] (enables the fast path whose taps travel as a kernel parameter)
int launch_conv_post_tanh(const void* x, int x_is_half, long long slot_stride, int row_stride, int row0, int L, int C, int k,
                          const float* w, const float* bias, float* wav_out, int n, const int* slot_ids, cudaStream_t st,
                          const float* taps_host = nullptr, long long lo_off = 0);   // lo_off: lo plane of a split fp16 input (elements)

// fused HiFi-GAN residual block on tcgen05 (resblock_fused.cu) -------------------------------------------------
struct ResblockFusedParams {
  const void* x; long long x_slot_stride; int x_rows, x_hist_rows, n_slots;   // compact fp16 input context [i][x_rows][C] = lrelu(x)
  int C, L, k; int dil[3]; int n_streams;
  const int* slot_ids;                              // slot of stream i (history block index)
  const void* w; const float* bias;                 // packed [w_copies][6*C][k*C] fp16 (c1.0, c2.0, c1.1, c2.1, c1.2, c2.2) and [6*C] fp32
  int w_copies;                                     // identical copies of the packed weights (spreads the L2 load of lockstep CTAs)
  void* hist; long long hist_slot_stride;           // resident fp16 history [slot][resblock_fused_hist_rows][C] (elements)
  void* hist_out = nullptr;                         // two-lane kernel: compact [stream][hist rows][C] staging for the NEW history (the
                                                    // caller scatters it to the slots afterwards); lets lanes cut streams between tiles
  const void* sum_in; void* sum_out;                // running MRF sum, compact fp16 [i][L][C] (either may be null)
  void* next; long long next_slot_stride; int next_row0;    // following layer's context rows <- lrelu(out_scale * (x_out + sum_in))
  float out_scale, slope;
};
bool resblock_fused_eligible(int C, int L, int k, const int* dil);
int resblock_fused_hist_rows(int k, const int* dil);
int launch_resblock_fused(const ResblockFusedParams& p, cudaStream_t st);
// two streams in flight per CTA (resblock_fused2.cu): same parameters, same resident history format
size_t resblock_fused2_smem(int C, int k, const int* dil);       // 0: does not fit
int launch_resblock_fused2(const ResblockFusedParams& p, cudaStream_t st);

// fused position-wise FFN (ffn_fused.cu): y partials = W2 . relu(W1 . x + b1), split-fp16 operands
struct FfnFusedParams {
  const void* x; long long x_lo_off; int x_rows;     // split fp16 planes [x_rows][96]: hi at x, lo at x + x_lo_off (elements)
  int M, K, hidden, N;                               // valid rows, padded model dim (96), hidden width, padded output dim (96)
  const void* w1; const float* b1;                   // [hidden][3*K] fp16 (W_hi | W_lo | W_hi, x 2^10), [hidden] fp32
  const void* w2;                                    // [N][3*hidden] fp16
  float* partials; int FS;                           // [FS][M][N] fp32
  float acc_scale;
};
bool ffn_fused_eligible(int K, int hidden, int N);
int ffn_fused_split(int M);
int launch_ffn_fused(const FfnFusedParams& p, cudaStream_t st);

// fused two-GEMM residual block (block_fused.cu): partial y = W2 . act(scale * (conv_k(x) + b1)), split-fp16 operands
struct BlockFusedParams {
  const void* x; long long x_slot_stride; int x_rows; int n_slots; long long lo_slot_off;   // split context [2 planes][slots][x_rows][C1], compact
  int C1, k, row0, L, n_streams;                     // taps read rows row0 + t + j, j < k
  const void* w1; const float* b1; int hidden;       // [hidden][3*k*C1] fp16 (x 2^10), [hidden] fp32
  float scale1; int act;                             // ACT_GELU (exact erf) or ACT_RELU
  const void* w2; int N2;                            // [256][3*hidden] fp16 (x 2^10)
  float* partials; int FS;                           // [FS][n_streams*L][256] fp32
  float acc_scale;
  // cluster reduction (out != nullptr, FS == 4): the four slices of a row tile are summed through distributed shared memory and the
  // kernel writes finished rows  out[row][c] = (y + b2[c] + res[row][c]) * mask[row]  (row stride ld; mask may be null; out may be res)
  float* out; const float* res; const float* b2; const float* mask; int ld;
  // optional (cluster mode): the LayerNorm that consumes the finished rows, applied by the reducing warp:
  //   ln_wmask[row] = (sum_c |x| > 0);  y = ((x * ln_pre[row] - mean) * rstd * ln_g + ln_b) * ln_post[row]  ->  ln_out [, ln_out2 fp32]
  const float* ln_g; const float* ln_b; float ln_eps;
  RowView ln_out; float* ln_out2; int ln_out2_ld;
  const float* ln_pre; const float* ln_post; float* ln_wmask;
  const float* ln_add;                               // optional fp32 rows (stride ld) added to y before both stores
};
bool block_fused_eligible(int C1, int k, int hidden, int N2, int L);
int block_fused_split(int n_streams, int L, int hidden, long long max_partial_rows);
int launch_block_fused(const BlockFusedParams& p, cudaStream_t st);

// log-mel front-end: |DFT| -> Slaney mel -> log10 -> clip, one CTA per frame (spec from the conv-GEMM engine)
int launch_logmel(const float* spec, int ld, int bins, const float* basis_t, int n_mels, float eps, float vmin, float vmax,
                  float* mel, long long n_frames, cudaStream_t st);

// state maintenance --------------------------------------------------------------------------
// Resident history of a context buffer lives per slot in `hist` [slot, hist_bytes]; the step works on a compact
// buffer `work` [i, hist_bytes + new_bytes].  gather: hist[slot_i] -> work[i][0 : hist);  scatter: the last
// hist_bytes of work[i] -> hist[slot_i].  scatter_back = 0 marks read-only session state (style vector).
struct HistDesc { void* work; long long work_stride_bytes; void* hist; int hist_bytes; int new_bytes; int scatter_back; };   // scatter_back 2: scatter only (never gathered)
int launch_hist_gather(const HistDesc* descs_dev, int n_descs, int n, const int* slot_ids, cudaStream_t st);
int launch_hist_scatter(const HistDesc* descs_dev, int n_descs, int n, const int* slot_ids, cudaStream_t st);
struct ZeroDesc { void* base; long long slot_stride_bytes; long long bytes; };
int launch_zero_slots(const ZeroDesc* descs_dev, int n_descs, int n, const int* slot_ids, cudaStream_t st);

// session setup ------------------------------------------------------------------------------
// mask[i,t] = (sum_c |x[i,t,c]| > 0)   /   mask[i,t] = (x[i,t,0] != 0)
int launch_row_masks(const float* ref, float* mask_abs, float* mask_first, int n, int T, int C, cudaStream_t st);
int launch_gated_tanh_sigmoid(const float* in, float* out, long long rows, int C, cudaStream_t st);   // in [rows,2C] -> out [rows,C]
// WN residual/skip update (wavenet.py:79-85): x = (x + rs[:, :C]) * mask ; skip += rs[:, C:]   (last: skip += rs)
int launch_wn_update(const float* rs, float* x, RowView x_ctx, float* skip, const float* mask, long long rows, int T, int C,
                     int last, cudaStream_t st);
// mean over groups of 4 frames of (skip * mask)  (seq_utils.py:307-325) -> [n, Tp, C]
int launch_group_mean4(const float* skip, const float* mask, float* out, int n, int T, int Tp, int C, cudaStream_t st);
// VQ (prosody_util.py:34-46, 88): idx = argmin_e |x|^2 + |e|^2 - 2 x.e ; z = x + (E[idx] - x); zcat = [z | sinusoid(pos)]
int launch_vq_quantize(const float* x, const float* xe, const float* E, const float* e2, const float* pos_table, float* zcat, int* idx_out,
                       int n, int Tp, int H, int n_codes, cudaStream_t st);
// key padding mask of the aligner (Conan.py:249): kpm[slot, p] = (pe[i, p, 0] == 0); also n_keys[slot] = Tp
int launch_kpm(const float* pe, float* kpm, int* n_keys, const int* slots_dev, int n, int Tp, int H, int tp_max, cudaStream_t st);
// masked temporal mean (Conan.py:214-219): style[slot, c] = sum_t x*mask / sum_t mask
int launch_masked_time_mean(const float* x, const float* mask, float* style, const int* slots_dev, int n, int T, int TS, int C, cudaStream_t st);
// scatter session K/V [n, Tp, 2H] -> cache[slot, layer, tp_max, 2H]
int launch_scatter_kv(const float* kv, float* cache, const int* slots_dev, int n, int Tp, int H2, int layer, int n_layers,
                      int tp_max, cudaStream_t st);

}  // namespace conan
