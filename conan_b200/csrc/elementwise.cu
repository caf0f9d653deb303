// Norm / gather / ring / small-reduction kernels of the path.  One warp per row wherever a
// row (<= 512 channels) is the unit of work, so every reduction is a warp shuffle and every
// global access is a coalesced 128-byte line.
#include <cstring>

#include "kernels.cuh"

namespace conan {

namespace {

constexpr int WARPS_PER_CTA = 8;

__device__ __forceinline__ int slot_of(const int* slot_ids, int i) { return slot_ids ? slot_ids[i] : i; }

// ------------------------------------------------------------------ LayerNorm over channels
__global__ void layernorm_kernel(LnArgs a) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  long long rows = (long long)a.n * a.L;
  if (warp >= rows) return;
  int i = warp / a.L, t = warp - i * a.L;
  int slot = slot_of(a.slot_ids, i);
  const float* x = reinterpret_cast<const float*>(a.in.base) + (long long)slot * a.in.slot_stride +
                   (long long)(a.in.row0 + t) * a.in.row_stride;
  float xs[8];                                    // partial-sum mode: the row (C <= 256) is assembled once, in registers
  if (a.part) {
    const long long row = (long long)i * a.L + t;
    const float pmask = a.part_mask ? a.part_mask[row] : 1.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = lane + 32 * q;
      float v = 0.f;
      if (c < a.C) {
        for (int p = 0; p < a.n_part; ++p) v += a.part[p * a.part_stride + row * a.part_ld + c];
        v = (v + a.part_bias[c] + a.part_res[row * a.part_res_ld + c]) * pmask;
        if (a.part_out) a.part_out[row * a.part_out_ld + c] = v;
      }
      xs[q] = v;
    }
  }
  auto val = [&](int c) { return a.part ? xs[c >> 5] : x[c]; };
  float pm = a.premask ? a.premask[(long long)slot * a.premask_slot_stride + t] : 1.f;
  float s = 0.f, sabs = 0.f;
  for (int c = lane; c < a.C; c += 32) { float v = val(c); sabs += fabsf(v); s += v * pm; }
  s = warp_sum(s);
  if (a.write_mask) {
    sabs = warp_sum(sabs);
    if (lane == 0) {
      float mk = sabs > 0.f ? 1.f : 0.f;
      a.write_mask[(long long)slot * a.write_mask_slot_stride + t] = mk;
      if (a.write_mask2) a.write_mask2[(long long)slot * a.write_mask_slot_stride + t] = mk;
    }
  }
  float mean = s / a.C;
  float q = 0.f;
  for (int c = lane; c < a.C; c += 32) { float d = val(c) * pm - mean; q += d * d; }
  q = warp_sum(q);
  float rstd = 1.f / sqrtf(q / a.C + a.eps);
  float post = a.postmask ? a.postmask[(long long)slot * a.postmask_slot_stride + t] : 1.f;
  long long o = (long long)slot * a.out.slot_stride + (long long)(a.out.row0 + t) * a.out.row_stride;
  long long o2 = (long long)slot * a.out2.slot_stride + (long long)(a.out2.row0 + t) * a.out2.row_stride;
  float ys[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = lane; c < a.C; c += 32) {
    float y = ((val(c) * pm - mean) * rstd * a.gamma[c] + a.beta[c]) * post;
    store_view(a.out, o + c, y);
    if (a.out2.base) store_view(a.out2, o2 + c, y);
    if (a.gamma2) ys[(c >> 5) & 3] = y;
  }
  if (a.gamma2) {                                 // chained LayerNorm on y (same eps), C <= 128
    float s2 = 0.f;
#pragma unroll
    for (int q2 = 0; q2 < 4; ++q2) if (lane + 32 * q2 < a.C) s2 += ys[q2];
    const float mean2 = warp_sum(s2) / a.C;
    float v2 = 0.f;
#pragma unroll
    for (int q2 = 0; q2 < 4; ++q2) if (lane + 32 * q2 < a.C) { const float d = ys[q2] - mean2; v2 += d * d; }
    const float rstd2 = 1.f / sqrtf(warp_sum(v2) / a.C + a.eps);
    const long long o3 = (long long)slot * a.out3.slot_stride + (long long)(a.out3.row0 + t) * a.out3.row_stride;
#pragma unroll
    for (int q2 = 0; q2 < 4; ++q2) {
      const int c = lane + 32 * q2;
      if (c < a.C) store_view(a.out3, o3 + c, (ys[q2] - mean2) * rstd2 * a.gamma2[c] + a.beta2[c]);
    }
  }
}

// ------------------------------------------------------------------ Emformer chunk assembly
__global__ void emformer_assemble_kernel(const float* __restrict__ chunk, float* __restrict__ X, int ldx, int n,
                                         const int* __restrict__ slot_ids, int seg, int rc, int D) {
  int rows = seg + rc;
  long long total = (long long)n * rows * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = idx % D; long long r = idx / D; int row = r % rows; int i = r / rows;
    // internal order [rc | utt]: internal row q <- chunk row (q < rc ? seg + q : q - rc)
    int src = row < rc ? seg + row : row - rc;
    X[((long long)slot_of(slot_ids, i) * rows + row) * ldx + c] = chunk[((long long)i * rows + src) * D + c];
  }
}

__global__ void advance_past_len_kernel(int* past_len, int n, const int* slot_ids, int seg) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) past_len[slot_of(slot_ids, i)] += seg;
}

__global__ void argmax_rows_kernel(const float* __restrict__ logits, int ld, int* tokens_a, int* tokens_b, int n, int rows, int C) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n * rows) return;
  const float* x = logits + (long long)warp * ld;
  float best = -INFINITY; int bi = 0x7fffffff;
  for (int c = lane; c < C; c += 32) { float v = x[c]; if (v > best) { best = v; bi = c; } }   // first max within a lane
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }                        // torch.argmax: first occurrence
  }
  if (lane == 0) { if (tokens_a) tokens_a[warp] = bi; if (tokens_b) tokens_b[warp] = bi; }
}

__global__ void copy_rows_out_kernel(const float* __restrict__ src, long long slot_stride, int row_stride, int row0,
                                     float* __restrict__ dst, int n, const int* slot_ids, int rows, int C) {
  long long total = (long long)n * rows * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = idx % C; long long r = idx / C; int t = r % rows; int i = r / rows;
    dst[idx] = src[(long long)slot_of(slot_ids, i) * slot_stride + (long long)(row0 + t) * row_stride + c];
  }
}

__global__ void copy_rows_in_kernel(const int* __restrict__ src, int* __restrict__ dst, long long slot_stride, int n,
                                    const int* slot_ids, int elems) {
  long long total = (long long)n * elems;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int e = idx % elems; int i = idx / elems;
    dst[(long long)slot_of(slot_ids, i) * slot_stride + e] = src[idx];     // 4-byte payload, int or float alike
  }
}

// ------------------------------------------------------------------ Conan chunk path
__global__ void embedding_rows_kernel(const int* __restrict__ tokens_slot, const float* __restrict__ table, int vocab,
                                      RowView out, int n, const int* slot_ids, int rows, int C) {
  long long total = (long long)n * rows * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = idx % C; long long r = idx / C; int t = r % rows; int i = r / rows;
    int slot = slot_of(slot_ids, i);
    int tok = tokens_slot[slot * rows + t];
    tok = min(max(tok, 0), vocab - 1);
    store_view(out, (long long)slot * out.slot_stride + (long long)(out.row0 + t) * out.row_stride + c, table[(long long)tok * C + c]);
  }
}

__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out1,
                                RowView out2, int n, const int* slot_ids, int rows, int C) {
  long long total = (long long)n * rows * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = idx % C; long long r = idx / C; int t = r % rows; int i = r / rows;
    int slot = slot_of(slot_ids, i);
    long long o = ((long long)slot * rows + t) * C + c;
    float v = a[o] + b[o];
    if (out1) out1[o] = v;
    if (out2.base) store_view(out2, (long long)slot * out2.slot_stride + (long long)(out2.row0 + t) * out2.row_stride + c, v);
  }
}

__global__ void rows_to_view_kernel(const float* __restrict__ src, RowView out, int n, const int* slot_ids, int rows, int C) {
  long long total = (long long)n * rows * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = idx % C; long long r = idx / C; int t = r % rows; int i = r / rows;
    int slot = slot_of(slot_ids, i);
    float v = src[((long long)slot * rows + t) * C + c];
    store_view(out, (long long)slot * out.slot_stride + (long long)(out.row0 + t) * out.row_stride + c, v);
  }
}

// uv_predictor: LayerNorm(128) -> Linear(128,2); uv = ch0 > 0 | token == silent; f0 = clamp(2^ch1, 50, 900), 0 if uv;
// coarse mel-scale bucket; dec_inp = pitch_inp + pitch_embed[bucket].  One warp per frame.
__global__ void pitch_kernel(const float* __restrict__ h, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                             const float* __restrict__ lin_w, const float* __restrict__ lin_b,
                             const int* __restrict__ tokens_slot, int silent_token, const float* __restrict__ pitch_table,
                             const float* __restrict__ pitch_inp, float* __restrict__ dec_inp, float* __restrict__ uv_pred_out,
                             int n, const int* slot_ids, int rows, int Cuv, int H) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n * rows) return;
  int i = warp / rows, t = warp - i * rows, slot = slot_of(slot_ids, i);
  const float* x = h + ((long long)slot * rows + t) * Cuv;
  float s = 0.f;
  for (int c = lane; c < Cuv; c += 32) s += x[c];
  float mean = warp_sum(s) / Cuv;
  float q = 0.f;
  for (int c = lane; c < Cuv; c += 32) { float d = x[c] - mean; q += d * d; }
  float rstd = 1.f / sqrtf(warp_sum(q) / Cuv + 1e-5f);
  float d0 = 0.f, d1 = 0.f;
  for (int c = lane; c < Cuv; c += 32) {
    float y = (x[c] - mean) * rstd * ln_g[c] + ln_b[c];
    d0 = fmaf(y, lin_w[c], d0); d1 = fmaf(y, lin_w[Cuv + c], d1);
  }
  d0 = warp_sum(d0) + lin_b[0]; d1 = warp_sum(d1) + lin_b[1];
  bool uv = (d0 > 0.f) || (tokens_slot[slot * rows + t] == silent_token);
  float f0 = exp2f(d1);
  f0 = fminf(fmaxf(f0, 50.f), 900.f);
  if (uv) f0 = 0.f;
  // f0_to_coarse, fp32 op order of the torch branch.  The reference's constants are float64
  // numpy scalars that torch narrows to fp32 at each tensor-scalar op:
  //   fp32(1127*ln(1+50/700)) = 0x1.370516p+6,  fp32(mel_max - mel_min) = 0x1.aaf4b6p+9
  const float kMelMin = 0x1.370516p+6f, kMelSpan = 0x1.aaf4b6p+9f;
  float f0_mel = 1127.f * logf(1.f + f0 / 700.f);
  if (f0_mel > 0.f) f0_mel = (f0_mel - kMelMin) * 254.f / kMelSpan + 1.f;
  if (f0_mel <= 1.f) f0_mel = 1.f;
  if (f0_mel > 255.f) f0_mel = 255.f;
  int bucket = (int)(f0_mel + 0.5f);
  if (lane == 0 && uv_pred_out) {
    uv_pred_out[((long long)slot * rows + t) * 4 + 0] = d0; uv_pred_out[((long long)slot * rows + t) * 4 + 1] = d1;
    uv_pred_out[((long long)slot * rows + t) * 4 + 2] = f0; uv_pred_out[((long long)slot * rows + t) * 4 + 3] = (float)bucket;
  }
  long long o = ((long long)slot * rows + t) * H;
  for (int c = lane; c < H; c += 32) dec_inp[o + c] = pitch_inp[o + c] + pitch_table[(long long)bucket * H + c];
}

// ------------------------------------------------------------------ conv_post (C -> 1) + tanh
// One thread per output sample: k*C MACs over k contiguous context rows.  (N = 1 is a dot
// product, not a tensor-core shape.)
template <typename T>
__global__ void conv_post_tanh_kernel(const T* __restrict__ x, long long slot_stride, int row_stride, int row0, int L, int C,
                                      int k, const float* __restrict__ w, const float* __restrict__ bias,
                                      float* __restrict__ wav, int n, const int* slot_ids, long long lo_off) {
  // lo_off != 0 (T = __half): x is a split fp16 pair, value = x[..] + x[.. + lo_off]
  extern __shared__ float ws[];                // [k*C]
  for (int q = threadIdx.x; q < k * C; q += blockDim.x) ws[q] = w[q];
  __syncthreads();
  constexpr int VEC = 16 / sizeof(T);          // elements per 16-byte load (C % VEC == 0, rows 16-byte aligned)
  long long total = (long long)n * L;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int t = idx % L; int i = idx / L;
    const T* xp = x + (long long)slot_of(slot_ids, i) * slot_stride + (long long)(row0 + t) * row_stride;
    float acc = bias[0];
    for (int j = 0; j < k; ++j) {
      const uint4* r = reinterpret_cast<const uint4*>(xp + (long long)j * row_stride);
      const float* wj = ws + j * C;
      for (int v = 0; v < C / VEC; ++v) {
        uint4 u = r[v];
        if (sizeof(T) == 2) {
          const __half2* h = reinterpret_cast<const __half2*>(&u);
          uint4 ul = make_uint4(0, 0, 0, 0);
          if (lo_off) ul = *(reinterpret_cast<const uint4*>(xp + lo_off + (long long)j * row_stride) + v);
          const __half2* hl = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 f = __half22float2(h[q]);
            const float2 g = __half22float2(hl[q]);
            f.x += g.x; f.y += g.y;
            acc = fmaf(f.x, wj[v * 8 + 2 * q], acc);
            acc = fmaf(f.y, wj[v * 8 + 2 * q + 1], acc);
          }
        } else {
          const float* f = reinterpret_cast<const float*>(&u);
#pragma unroll
          for (int q = 0; q < 4; ++q) acc = fmaf(f[q], wj[v * 4 + q], acc);
        }
      }
    }
    wav[idx] = tanhf(acc);
  }
}

// fp16 fast path: a CTA stages the (TB + k - 1) x C input rows of TB consecutive output samples in shared memory with
// coalesced 16-byte loads (rows padded by 16 bytes: conflict-free per-thread row reads), then one thread per sample.
// The K*C filter taps travel as a kernel parameter: every FFMA takes its weight straight from the constant bank (the index is
// compile-time after unrolling), instead of one shared-memory load per multiply.
template <int C, int K>
struct PostTaps { float w[K * C]; float bias; };

template <int C, int K, int TB>
__global__ void __launch_bounds__(TB)
conv_post_tanh_tiled_kernel(const __half* __restrict__ x, long long slot_stride, int row0, int L, const __grid_constant__ PostTaps<C, K> taps,
                            float* __restrict__ wav, const int* slot_ids) {
  constexpr int ROWH = C + 8;                    // padded row, in halfs
  constexpr int V = C / 8;                       // 16-byte vectors per row
  __shared__ __align__(16) __half sx[(TB + K - 1) * ROWH];
  const int tiles = L / TB;
  const int i = blockIdx.x / tiles, t0 = (blockIdx.x - i * tiles) * TB;
  const __half* xp = x + (long long)slot_of(slot_ids, i) * slot_stride + (long long)(row0 + t0) * C;
  for (int q = threadIdx.x; q < (TB + K - 1) * V; q += TB) {
    const int r = q / V, v = q - r * V;
    *reinterpret_cast<uint4*>(&sx[r * ROWH + v * 8]) = *(reinterpret_cast<const uint4*>(xp + (long long)r * C) + v);
  }
  __syncthreads();
  float acc = taps.bias;
#pragma unroll
  for (int j = 0; j < K; ++j) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const uint4 u = *reinterpret_cast<const uint4*>(&sx[(threadIdx.x + j) * ROWH + v * 8]);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __half22float2(h[q]);
        acc = fmaf(f.x, taps.w[j * C + v * 8 + 2 * q], acc);
        acc = fmaf(f.y, taps.w[j * C + v * 8 + 2 * q + 1], acc);
      }
    }
  }
  wav[(long long)i * L + t0 + threadIdx.x] = tanhf(acc);
}

// ------------------------------------------------------------------ log-mel front-end
// spec [n*frames, ld] = windowed DFT of each frame: columns [0, bins) real parts, [bins, 2*bins) imaginary parts (produced by
// the conv-GEMM engine: a frame is 4 taps over 320-sample rows of the centre-padded signal).  One CTA per frame:
// magnitude -> Slaney mel (basis transposed [bins][n_mels]: coalesced over mel channels) -> log10(max(., eps)) -> clip.
// utils/audio/__init__.py:62-72, inference/Conan.py:58-70.
__global__ void __launch_bounds__(128)
logmel_kernel(const float* __restrict__ spec, int ld, int bins, const float* __restrict__ basis_t, int n_mels, float eps,
              float vmin, float vmax, float* __restrict__ mel) {
  extern __shared__ float mag[];
  const float* sp = spec + (long long)blockIdx.x * ld;
  for (int b = threadIdx.x; b < bins; b += blockDim.x) {
    const float re = sp[b], im = sp[bins + b];
    mag[b] = sqrtf(re * re + im * im);
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    float acc = 0.f;
#pragma unroll 1                                   // (an unrolled version issued shared loads past `bins`: compute-sanitizer)
    for (int b = 0; b < bins; ++b) acc = fmaf(basis_t[(long long)b * n_mels + m], mag[b], acc);
    const float v = log10f(fmaxf(acc, eps));
    mel[(long long)blockIdx.x * n_mels + m] = fminf(fmaxf(v, vmin), vmax);
  }
}

// ------------------------------------------------------------------ ring maintenance
__global__ void hist_move_kernel(const HistDesc* __restrict__ descs, const int* slot_ids, int scatter) {
  HistDesc d = descs[blockIdx.x];
  if (scatter ? !d.scatter_back : d.scatter_back == 2) return;
  int slot = slot_of(slot_ids, blockIdx.y);
  unsigned char* w = reinterpret_cast<unsigned char*>(d.work) + (long long)blockIdx.y * d.work_stride_bytes + (scatter ? d.new_bytes : 0);
  unsigned char* h = reinterpret_cast<unsigned char*>(d.hist) + (long long)slot * d.hist_bytes;
  const uint4* src = reinterpret_cast<const uint4*>(scatter ? w : h);
  uint4* dst = reinterpret_cast<uint4*>(scatter ? h : w);
  int nv = d.hist_bytes >> 4;                              // sizes are multiples of 16 bytes
  for (int q = threadIdx.x; q < nv; q += blockDim.x) dst[q] = src[q];
}

__global__ void zero_slots_kernel(const ZeroDesc* __restrict__ descs, int n, const int* slot_ids) {
  ZeroDesc d = descs[blockIdx.x];
  int slot = slot_of(slot_ids, blockIdx.y);
  uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(d.base) + (long long)slot * d.slot_stride_bytes);
  long long nv = d.bytes >> 4;
  uint4 z = make_uint4(0, 0, 0, 0);
  for (long long q = blockIdx.z * (long long)blockDim.x + threadIdx.x; q < nv; q += (long long)gridDim.z * blockDim.x) p[q] = z;
  if (blockIdx.z == 0) {      // tail (< 16 bytes), sizes are multiples of 4
    int* tail = reinterpret_cast<int*>(p + nv);
    int rem = (int)((d.bytes & 15) >> 2);
    if ((int)threadIdx.x < rem) tail[threadIdx.x] = 0;
  }
}

// ------------------------------------------------------------------ session-setup kernels
__global__ void row_masks_kernel(const float* __restrict__ ref, float* mask_abs, float* mask_first, long long rows, int C) {
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* x = ref + warp * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += fabsf(x[c]);
  s = warp_sum(s);
  if (lane == 0) { mask_abs[warp] = s > 0.f ? 1.f : 0.f; mask_first[warp] = x[0] != 0.f ? 1.f : 0.f; }
}

__global__ void gated_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int C) {
  long long total = rows * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / C; int c = idx % C;
    float a = in[r * 2 * C + c], b = in[r * 2 * C + C + c];
    out[idx] = tanhf(a) * (1.f / (1.f + expf(-b)));
  }
}

__global__ void wn_update_kernel(const float* __restrict__ rs, float* __restrict__ x, RowView x_ctx, float* __restrict__ skip,
                                 const float* __restrict__ mask, long long rows, int T, int C, int last) {
  long long total = rows * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / C; int c = idx % C;
    if (last) { skip[idx] += rs[r * C + c]; continue; }
    float xv = (x[idx] + rs[r * 2 * C + c]) * mask[r];
    x[idx] = xv;
    long long i = r / T; int t = r % T;
    reinterpret_cast<float*>(x_ctx.base)[i * x_ctx.slot_stride + (long long)(x_ctx.row0 + t) * x_ctx.row_stride + c] = xv;
    skip[idx] += rs[r * 2 * C + C + c];
  }
}

__global__ void group_mean4_kernel(const float* __restrict__ skip, const float* __restrict__ mask, float* __restrict__ out,
                                   int n, int T, int Tp, int C) {
  long long total = (long long)n * Tp * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = idx % C; long long r = idx / C; int p = r % Tp; int i = r / Tp;
    float s = 0.f; int cnt = 0;
    for (int q = 0; q < 4; ++q) {
      int t = p * 4 + q;
      if (t < T) { s += skip[((long long)i * T + t) * C + c] * mask[(long long)i * T + t]; ++cnt; }
    }
    out[idx] = s / (float)max(cnt, 1);
  }
}

// One CTA per session: argmin over codes per token (warp per token), then the sequential
// position count along the tokens, then zcat = [z | sinusoid(pos)].
__global__ void vq_quantize_kernel(const float* __restrict__ x, const float* __restrict__ xe, const float* __restrict__ E,
                                   const float* __restrict__ e2, const float* __restrict__ pos_table, float* __restrict__ zcat,
                                   int* __restrict__ idx_out, int Tp, int H, int n_codes) {
  extern __shared__ int sh_i[];                       // [Tp] code index, then [Tp] nonzero flag
  int* s_idx = sh_i; int* s_nz = sh_i + Tp; int* s_pos = sh_i + 2 * Tp;
  int i = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int p = warp; p < Tp; p += nw) {
    const float* xr = x + ((long long)i * Tp + p) * H;
    float x2 = 0.f;
    for (int c = lane; c < H; c += 32) x2 += xr[c] * xr[c];
    x2 = warp_sum(x2);
    const float* xer = xe + ((long long)i * Tp + p) * n_codes;
    float best = INFINITY; int bi = 0x7fffffff;
    for (int j = lane; j < n_codes; j += 32) {
      float d = (e2[j] + x2) + (-2.f) * xer[j];       // addmm(beta=1 * (e2 + x2), alpha=-2 * x.E^T)
      if (d < best) { best = d; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    float z0 = xr[0] + (E[(long long)bi * H] - xr[0]);
    if (lane == 0) { s_idx[p] = bi; s_nz[p] = z0 != 0.f ? 1 : 0; if (idx_out) idx_out[(long long)i * Tp + p] = bi; }
  }
  __syncthreads();
  if (threadIdx.x == 0) { int run = 0; for (int p = 0; p < Tp; ++p) { run += s_nz[p]; s_pos[p] = s_nz[p] ? run : 0; } }
  __syncthreads();
  for (long long q = threadIdx.x; q < (long long)Tp * H; q += blockDim.x) {
    int p = q / H, c = q % H;
    float xv = x[((long long)i * Tp + p) * H + c];
    float z = xv + (E[(long long)s_idx[p] * H + c] - xv);                      // x + (q - x), prosody_util.py:88
    zcat[((long long)i * Tp + p) * 2 * H + c] = z;
    zcat[((long long)i * Tp + p) * 2 * H + H + c] = pos_table[(long long)s_pos[p] * H + c];
  }
}

__global__ void kpm_kernel(const float* __restrict__ pe, float* __restrict__ kpm, int* __restrict__ n_keys,
                           const int* __restrict__ slots, int n, int Tp, int H, int tp_max) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * tp_max) return;
  int i = idx / tp_max, p = idx % tp_max, slot = slots[i];
  kpm[(long long)slot * tp_max + p] = (p < Tp) ? (pe[((long long)i * Tp + p) * H] == 0.f ? 1.f : 0.f) : 1.f;
  if (p == 0) n_keys[slot] = Tp;
}

__global__ void masked_time_mean_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ style,
                                        const int* __restrict__ slots, int T, int TS, int C) {
  // TS: rows per session in x / mask (>= T)
  int i = blockIdx.x; int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f, cnt = 0.f;
  for (int t = 0; t < T; ++t) { float m = mask[(long long)i * TS + t]; s += x[((long long)i * TS + t) * C + c] * m; cnt += m; }
  style[(long long)slots[i] * C + c] = s / cnt;
}

__global__ void scatter_kv_kernel(const float* __restrict__ kv, float* __restrict__ cache, const int* __restrict__ slots,
                                  int n, int Tp, int H2, int layer, int n_layers, int tp_max) {
  long long total = (long long)n * Tp * H2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = idx % H2; long long r = idx / H2; int p = r % Tp; int i = r / Tp;
    cache[(((long long)slots[i] * n_layers + layer) * tp_max + p) * H2 + c] = kv[idx];
  }
}

inline unsigned grid_for(long long total, int block, unsigned cap = 148u * 16u) {
  long long g = (total + block - 1) / block;
  if (g < 1) g = 1;
  return (unsigned)(g > cap ? cap : g);
}

}  // namespace

// ===================================================================== launchers
int launch_layernorm(const LnArgs& a, cudaStream_t st) {
  long long rows = (long long)a.n * a.L;
  if (rows <= 0) return 0;
  int block = 32 * WARPS_PER_CTA;
  layernorm_kernel<<<(unsigned)((rows + WARPS_PER_CTA - 1) / WARPS_PER_CTA), block, 0, st>>>(a);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_emformer_assemble(const float* chunk, float* X, int ldx, int n, const int* slot_ids, int seg, int rc, int D, cudaStream_t st) {
  if (n <= 0) return 0;
  emformer_assemble_kernel<<<grid_for((long long)n * (seg + rc) * D, 256), 256, 0, st>>>(chunk, X, ldx, n, slot_ids, seg, rc, D);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_advance_past_len(int* past_len, int n, const int* slot_ids, int seg, cudaStream_t st) {
  if (n <= 0) return 0;
  advance_past_len_kernel<<<(n + 255) / 256, 256, 0, st>>>(past_len, n, slot_ids, seg);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_argmax_rows(const float* logits, int ld, int* tokens_a, int* tokens_b, int n, int rows, int C, cudaStream_t st) {
  if (n <= 0) return 0;
  long long warps = (long long)n * rows;
  argmax_rows_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(logits, ld, tokens_a, tokens_b, n, rows, C);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_copy_rows_out(const float* src_slot, long long slot_stride, int row_stride, int row0, float* dst, int n,
                         const int* slot_ids, int rows, int C, cudaStream_t st) {
  if (n <= 0) return 0;
  copy_rows_out_kernel<<<grid_for((long long)n * rows * C, 256), 256, 0, st>>>(src_slot, slot_stride, row_stride, row0, dst, n, slot_ids, rows, C);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_copy_rows_in(const void* src, int, void* dst_slot, long long slot_stride_elems, int n, const int* slot_ids, int elems, cudaStream_t st) {
  if (n <= 0) return 0;
  copy_rows_in_kernel<<<grid_for((long long)n * elems, 256), 256, 0, st>>>((const int*)src, (int*)dst_slot, slot_stride_elems, n, slot_ids, elems);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_embedding_rows(const int* tokens_slot, const float* table, int vocab, RowView out, int n, const int* slot_ids,
                          int rows, int C, cudaStream_t st) {
  if (n <= 0) return 0;
  embedding_rows_kernel<<<grid_for((long long)n * rows * C, 256), 256, 0, st>>>(tokens_slot, table, vocab, out, n, slot_ids, rows, C);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_add_rows(const float* a, const float* b, float* out1, RowView out2, int n, const int* slot_ids, int rows, int C, cudaStream_t st) {
  if (n <= 0) return 0;
  add_rows_kernel<<<grid_for((long long)n * rows * C, 256), 256, 0, st>>>(a, b, out1, out2, n, slot_ids, rows, C);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_rows_to_view(const float* src_slot, RowView out, int n, const int* slot_ids, int rows, int C, cudaStream_t st) {
  if (n <= 0) return 0;
  rows_to_view_kernel<<<grid_for((long long)n * rows * C, 256), 256, 0, st>>>(src_slot, out, n, slot_ids, rows, C);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_pitch(const float* h, const float* ln_g, const float* ln_b, const float* lin_w, const float* lin_b,
                 const int* tokens_slot, int silent_token, const float* pitch_table, const float* pitch_inp,
                 float* dec_inp, float* uv_pred_out, int n, const int* slot_ids, int rows, int Cuv, int H, cudaStream_t st) {
  if (n <= 0) return 0;
  long long warps = (long long)n * rows;
  pitch_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(h, ln_g, ln_b, lin_w, lin_b, tokens_slot, silent_token, pitch_table,
                                                           pitch_inp, dec_inp, uv_pred_out, n, slot_ids, rows, Cuv, H);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_conv_post_tanh(const void* x, int x_is_half, long long slot_stride, int row_stride, int row0, int L, int C, int k,
                          const float* w, const float* bias, float* wav_out, int n, const int* slot_ids, cudaStream_t st, const float* taps_host,
                          long long lo_off) {
  if (n <= 0) return 0;
  unsigned grid = grid_for((long long)n * L, 256, 148u * 32u);
  size_t sh = (size_t)k * C * sizeof(float);
  if (taps_host && x_is_half && !lo_off && C == 32 && k == 7 && row_stride == C && L % 256 == 0 && slot_stride % 8 == 0 && ((uintptr_t)x) % 16 == 0) {
    PostTaps<32, 7> taps;
    memcpy(taps.w, taps_host, sizeof(taps.w));
    taps.bias = taps_host[7 * 32];
    conv_post_tanh_tiled_kernel<32, 7, 256><<<(unsigned)(n * (L / 256)), 256, 0, st>>>((const __half*)x, slot_stride, row0, L, taps, wav_out,
                                                                                      slot_ids);
    CONAN_CHECK_LAUNCH();
    return 0;
  }
  if (x_is_half)
    conv_post_tanh_kernel<__half><<<grid, 256, sh, st>>>((const __half*)x, slot_stride, row_stride, row0, L, C, k, w, bias, wav_out, n, slot_ids, lo_off);
  else
    conv_post_tanh_kernel<float><<<grid, 256, sh, st>>>((const float*)x, slot_stride, row_stride, row0, L, C, k, w, bias, wav_out, n, slot_ids, 0);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_logmel(const float* spec, int ld, int bins, const float* basis_t, int n_mels, float eps, float vmin, float vmax,
                  float* mel, long long n_frames, cudaStream_t st) {
  if (n_frames <= 0) return 0;
  logmel_kernel<<<(unsigned)n_frames, 128, ((size_t)bins + 64) * sizeof(float), st>>>(spec, ld, bins, basis_t, n_mels, eps, vmin, vmax, mel);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_hist_gather(const HistDesc* descs_dev, int n_descs, int n, const int* slot_ids, cudaStream_t st) {
  if (n <= 0 || n_descs <= 0) return 0;
  hist_move_kernel<<<dim3(n_descs, n), 128, 0, st>>>(descs_dev, slot_ids, 0);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_hist_scatter(const HistDesc* descs_dev, int n_descs, int n, const int* slot_ids, cudaStream_t st) {
  if (n <= 0 || n_descs <= 0) return 0;
  hist_move_kernel<<<dim3(n_descs, n), 128, 0, st>>>(descs_dev, slot_ids, 1);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_zero_slots(const ZeroDesc* descs_dev, int n_descs, int n, const int* slot_ids, cudaStream_t st) {
  if (n <= 0 || n_descs <= 0) return 0;
  zero_slots_kernel<<<dim3(n_descs, n, 4), 256, 0, st>>>(descs_dev, n, slot_ids);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_row_masks(const float* ref, float* mask_abs, float* mask_first, int n, int T, int C, cudaStream_t st) {
  long long rows = (long long)n * T;
  if (rows <= 0) return 0;
  row_masks_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(ref, mask_abs, mask_first, rows, C);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_gated_tanh_sigmoid(const float* in, float* out, long long rows, int C, cudaStream_t st) {
  if (rows <= 0) return 0;
  gated_kernel<<<grid_for(rows * C, 256), 256, 0, st>>>(in, out, rows, C);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_wn_update(const float* rs, float* x, RowView x_ctx, float* skip, const float* mask, long long rows, int T, int C,
                     int last, cudaStream_t st) {
  if (rows <= 0) return 0;
  wn_update_kernel<<<grid_for(rows * C, 256), 256, 0, st>>>(rs, x, x_ctx, skip, mask, rows, T, C, last);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_group_mean4(const float* skip, const float* mask, float* out, int n, int T, int Tp, int C, cudaStream_t st) {
  if (n <= 0) return 0;
  group_mean4_kernel<<<grid_for((long long)n * Tp * C, 256), 256, 0, st>>>(skip, mask, out, n, T, Tp, C);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_vq_quantize(const float* x, const float* xe, const float* E, const float* e2, const float* pos_table,
                            float* zcat, int* idx_out, int n, int Tp, int H, int n_codes, cudaStream_t st) {
  if (n <= 0) return 0;
  vq_quantize_kernel<<<n, 256, (size_t)3 * Tp * sizeof(int), st>>>(x, xe, E, e2, pos_table, zcat, idx_out, Tp, H, n_codes);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_kpm(const float* pe, float* kpm, int* n_keys, const int* slots_dev, int n, int Tp, int H, int tp_max, cudaStream_t st) {
  if (n <= 0) return 0;
  kpm_kernel<<<(n * tp_max + 255) / 256, 256, 0, st>>>(pe, kpm, n_keys, slots_dev, n, Tp, H, tp_max);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_masked_time_mean(const float* x, const float* mask, float* style, const int* slots_dev, int n, int T, int TS, int C, cudaStream_t st) {
  if (n <= 0) return 0;
  masked_time_mean_kernel<<<dim3(n, (C + 127) / 128), 128, 0, st>>>(x, mask, style, slots_dev, T, TS, C);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_scatter_kv(const float* kv, float* cache, const int* slots_dev, int n, int Tp, int H2, int layer, int n_layers,
                      int tp_max, cudaStream_t st) {
  if (n <= 0) return 0;
  scatter_kv_kernel<<<grid_for((long long)n * Tp * H2, 256), 256, 0, st>>>(kv, cache, slots_dev, n, Tp, H2, layer, n_layers, tp_max);
  CONAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace conan
