// Generic Emformer streaming step: memory bank (max_memory_size M > 0), summary query and partial segments.
//
// The reference configuration runs M = 0 (modules/Emformer/emformer.py:14-22 never passes max_memory_size) on whole
// segments; that case keeps its own fused kernels (attention.cu, ffn_fused.cu).  The kernels here cover what torchaudio's
// Emformer does beyond it (torchaudio/models/emformer.py, "TA"):
//   * memory bank: every layer keeps its last M input memory vectors per stream (state[0], TA:384-414).  A step projects the
//     m = min(M, ceil(past_len / seg)) newest of them to keys / values (TA:164), appends one SUMMARY query = mean of the
//     layer-normed utterance rows (AvgPool1d(seg), TA:478-480), masks the memory keys for that query (TA:297-300, -1e8 before an
//     fp32 softmax == excluded), clamps the summary's output row to [-10, 10] as the next layer's memory input (TA:211-215) and
//     pushes the layer's INPUT memory into the bank (TA:409, `mems`, not the new one);
//   * partial segments: the last segment of a full-utterance `forward` (TA:709-743) may hold fewer than seg frames; rows
//     n_utt..seg-1 are neither keys nor appended to the K/V ring, and past_len advances by n_utt.
// Row layout of every per-stream work buffer on this path: [rc | utt (seg slots) | summary] = seg + rc + 1 rows.
#include <algorithm>

#include "kernels.cuh"

namespace conan {

namespace {

constexpr int EMG_MAX_KEYS = 64;
constexpr int EMG_MAX_M = 8;

__device__ __forceinline__ int slot_of(const int* slot_ids, int i) { return slot_ids ? slot_ids[i] : i; }

// chunk rows of stream i: utterance rows at src[i*stride + (utt_row0 + t)*D], look-ahead rows at src[i*stride + (rc_row0 + q)*D]
// -> X[i] rows [rc | utt | 0] (row stride ldx, erows rows); mem0[i] = mean of the n_utt raw utterance rows (TA:786-789)
__global__ void emformer_assemble_generic_kernel(const float* __restrict__ src, long long stream_stride, int utt_row0, int rc_row0,
                                                 float* __restrict__ X, int ldx, float* __restrict__ mem0, int seg, int n_utt, int rc, int D) {
  const int i = blockIdx.x, erows = seg + rc + 1;
  const float* s = src + (long long)i * stream_stride;
  float* x = X + (long long)i * erows * ldx;
  for (int idx = threadIdx.x; idx < erows * D; idx += blockDim.x) {
    const int row = idx / D, c = idx - row * D;
    float v = 0.f;
    if (row < rc) v = s[(long long)(rc_row0 + row) * D + c];
    else if (row < rc + n_utt) v = s[(long long)(utt_row0 + row - rc) * D + c];
    x[(long long)row * ldx + c] = v;
  }
  if (mem0) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float a = 0.f;
      for (int t = 0; t < n_utt; ++t) a += s[(long long)(utt_row0 + t) * D + c];
      mem0[(long long)i * D + c] = a / (float)n_utt;
    }
  }
}

// summary row of the QKV operand <- mean of the layer-normed utterance rows; memory-bank rows -> GEMM operand rows
__global__ void emformer_mem_prepare_kernel(const float* __restrict__ xnf, int ld, RowView xn, const float* __restrict__ bank,
                                            RowView mb, const int* __restrict__ slot_ids, int seg, int n_utt, int rc, int D, int M) {
  const int i = blockIdx.x, erows = seg + rc + 1;
  const float* x = xnf + (long long)i * erows * ld;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f;
    for (int t = 0; t < n_utt; ++t) a += x[(long long)(rc + t) * ld + c];
    store_view(xn, (long long)i * xn.slot_stride + (long long)(xn.row0 + rc + seg) * xn.row_stride + c, a / (float)n_utt);
  }
  const float* b = bank + (long long)slot_of(slot_ids, i) * M * D;
  for (int idx = threadIdx.x; idx < M * D; idx += blockDim.x) {
    const int j = idx / D, c = idx - j * D;
    store_view(mb, (long long)i * mb.slot_stride + (long long)(mb.row0 + j) * mb.row_stride + c, b[idx]);
  }
}

// next layer's memory input <- clamp(summary output row); bank <- push(this layer's input memory)
__global__ void emformer_mem_update_kernel(const float* __restrict__ r1, int ld, const float* __restrict__ mem_in, float* __restrict__ mem_out,
                                           float* __restrict__ bank, const int* __restrict__ slot_ids, int seg, int rc, int D, int M) {
  const int i = blockIdx.x, erows = seg + rc + 1;
  float* b = bank + (long long)slot_of(slot_ids, i) * M * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = r1[((long long)i * erows + rc + seg) * ld + c];
    mem_out[(long long)i * D + c] = fminf(fmaxf(v, -10.f), 10.f);
    for (int j = 0; j + 1 < M; ++j) b[(long long)j * D + c] = b[(long long)(j + 1) * D + c];     // one thread owns column c: in-order shift
    b[(long long)(M - 1) * D + c] = mem_in[(long long)i * D + c];
  }
}

// keys in the reference's order [mems (m) | rc | left context (oldest first) | utt (n_utt)]; queries [rc | utt | summary]
__global__ void __launch_bounds__(256)
emformer_attention_mem_kernel(const float* __restrict__ qkv, const float* __restrict__ memkv, float* __restrict__ ring,
                              const int* __restrict__ past_len, RowView att, const int* __restrict__ slot_ids, int seg, int n_utt,
                              int rc, int lc, int ring_rows, int D, int heads, int ldq, int M) {
  extern __shared__ float sm[];
  const int erows = seg + rc + 1;
  const int slot = slot_of(slot_ids, blockIdx.x);
  const int past = past_len[slot];
  const int lc_len = min(lc, past);
  const int m = M > 0 ? min(M, (past + seg - 1) / seg) : 0;
  const int nkeys = m + rc + lc_len + n_utt;
  const int DS = D + 1;
  float* sK = sm;                                     // [EMG_MAX_KEYS][DS]
  float* sV = sK + (size_t)EMG_MAX_KEYS * DS;
  float* sQ = sV + (size_t)EMG_MAX_KEYS * DS;        // [erows][D], scaled
  float* sP = sQ + (size_t)erows * D;                // [warps][EMG_MAX_KEYS]
  const float* q_in = qkv + (long long)blockIdx.x * erows * ldq;
  const float* mk = memkv ? memkv + (long long)blockIdx.x * M * ldq : nullptr;
  float* rg = ring + (long long)slot * ring_rows * 2 * D;
  const int tid = threadIdx.x, hd = D / heads;
  const float scaling = rsqrtf((float)hd);
  for (int idx = tid; idx < erows * D; idx += blockDim.x) {
    const int r = idx / D, c = idx - r * D;
    sQ[idx] = q_in[(long long)r * ldq + c] * scaling;
  }
  for (int idx = tid; idx < nkeys * 2 * D; idx += blockDim.x) {
    const int key = idx / (2 * D), c = idx - key * 2 * D;
    const float* src;
    if (key < m) src = mk + (long long)(M - m + key) * ldq + D;                                            // newest m bank rows
    else if (key < m + rc) src = q_in + (long long)(key - m) * ldq + D;                                    // look-ahead rows
    else if (key < m + rc + lc_len) src = rg + (long long)((past - lc_len + (key - m - rc)) % ring_rows) * 2 * D;
    else src = q_in + (long long)(rc + (key - m - rc - lc_len)) * ldq + D;                                 // this chunk's utterance rows
    const float v = src[c];
    if (c < D) sK[key * DS + c] = v; else sV[key * DS + (c - D)] = v;
  }
  __syncthreads();
  // state update (_pack_state, TA:400-414): only the n_utt real utterance rows enter the ring
  for (int idx = tid; idx < n_utt * 2 * D; idx += blockDim.x) {
    const int t = idx / (2 * D), c = idx - t * 2 * D;
    rg[(long long)((past + t) % ring_rows) * 2 * D + c] = q_in[(long long)(rc + t) * ldq + D + c];
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  float* p = sP + warp * EMG_MAX_KEYS;
  for (int h = warp; h < heads; h += nwarps) {
    const int c0 = h * hd;
    for (int r = 0; r < erows; ++r) {
      const bool summary = r == rc + seg;
      const bool live = r < rc + n_utt || (summary && M > 0);
      if (!live) {                                   // padding rows of a partial segment / no summary at M = 0
        for (int d = lane; d < hd; d += 32) store_view(att, (long long)blockIdx.x * att.slot_stride + (long long)(att.row0 + r) * att.row_stride + c0 + d, 0.f);
        continue;
      }
      float s[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int key = lane + 32 * u;
        float a = -INFINITY;
        if (key < nkeys && !(summary && key < m)) {  // the summary query does not see the memory keys (TA:297-300)
          a = 0.f;
          for (int d = 0; d < hd; ++d) a = fmaf(sQ[r * D + c0 + d], sK[key * DS + c0 + d], a);
        }
        s[u] = a;
      }
      const float mx = warp_max(fmaxf(s[0], s[1]));
      const float e0 = s[0] == -INFINITY ? 0.f : expf(s[0] - mx), e1 = s[1] == -INFINITY ? 0.f : expf(s[1] - mx);
      const float inv = 1.f / warp_sum(e0 + e1);
      p[lane] = e0 * inv; p[lane + 32] = e1 * inv;
      __syncwarp();
      for (int d = lane; d < hd; d += 32) {
        float v = 0.f;
        for (int key = 0; key < nkeys; ++key) v = fmaf(p[key], sV[key * DS + c0 + d], v);
        store_view(att, (long long)blockIdx.x * att.slot_stride + (long long)(att.row0 + r) * att.row_stride + c0 + d, v);
      }
      __syncwarp();
    }
  }
}

// rows [src_row0, src_row0 + rows) x C columns of every stream, between two strided per-stream layouts
__global__ void copy_rows_strided_kernel(const float* __restrict__ src, long long src_slot_stride, int src_ld, int src_row0,
                                         float* __restrict__ dst, long long dst_slot_stride, int dst_ld, int dst_row0, int n, int rows, int C) {
  const long long total = (long long)n * rows * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % C; const long long r = idx / C; const int t = r % rows; const long long i = r / rows;
    dst[i * dst_slot_stride + (long long)(dst_row0 + t) * dst_ld + c] = src[i * src_slot_stride + (long long)(src_row0 + t) * src_ld + c];
  }
}

__global__ void advance_past_len_by_kernel(int* past_len, int n, const int* slot_ids, int by) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) past_len[slot_of(slot_ids, i)] += by;
}

}  // namespace

int launch_emformer_assemble_generic(const float* src, long long stream_stride, int utt_row0, int rc_row0, float* X, int ldx,
                                     float* mem0, int n, int seg, int n_utt, int rc, int D, cudaStream_t st) {
  if (n <= 0) return 0;
  emformer_assemble_generic_kernel<<<n, 128, 0, st>>>(src, stream_stride, utt_row0, rc_row0, X, ldx, mem0, seg, n_utt, rc, D);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_emformer_mem_prepare(const float* xnf, int ld, RowView xn, const float* bank, RowView mb, int n, const int* slot_ids,
                                int seg, int n_utt, int rc, int D, int M, cudaStream_t st) {
  if (n <= 0) return 0;
  emformer_mem_prepare_kernel<<<n, 128, 0, st>>>(xnf, ld, xn, bank, mb, slot_ids, seg, n_utt, rc, D, M);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_emformer_mem_update(const float* r1, int ld, const float* mem_in, float* mem_out, float* bank, int n, const int* slot_ids,
                               int seg, int rc, int D, int M, cudaStream_t st) {
  if (n <= 0) return 0;
  emformer_mem_update_kernel<<<n, 128, 0, st>>>(r1, ld, mem_in, mem_out, bank, slot_ids, seg, rc, D, M);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_emformer_attention_mem(const float* qkv, const float* memkv, float* ring, const int* past_len, RowView att, int n,
                                  const int* slot_ids, int seg, int n_utt, int rc, int lc, int ring_rows, int D, int heads, int ldq,
                                  int M, cudaStream_t st) {
  if (n <= 0) return 0;
  if (M > EMG_MAX_M || M + rc + lc + seg > EMG_MAX_KEYS) { set_error("emformer_attention_mem: memory + contexts + segment above 64 keys (or M > 8)"); return 1; }
  if (n_utt < 1 || n_utt > seg || D % heads != 0) { set_error("emformer_attention_mem: bad n_utt / head split"); return 1; }
  if (ring_rows < lc + seg) { set_error("emformer_attention_mem: ring too short"); return 1; }
  const size_t sh = ((size_t)2 * EMG_MAX_KEYS * (D + 1) + (size_t)(seg + rc + 1) * D + (size_t)8 * EMG_MAX_KEYS) * sizeof(float);
  if (sh > 96 * 1024) { set_error("emformer_attention_mem: shared memory above 96 KB"); return 1; }
  static DeviceOnce once;
  if (device_once(once, nullptr, [&](int*) {
        cudaError_t e = cudaFuncSetAttribute(emformer_attention_mem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
        return 0;
      }))
    return 1;
  emformer_attention_mem_kernel<<<n, 256, sh, st>>>(qkv, memkv, ring, past_len, att, slot_ids, seg, n_utt, rc, lc, ring_rows, D, heads, ldq, M);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_copy_rows_strided(const float* src, long long src_slot_stride, int src_ld, int src_row0, float* dst, long long dst_slot_stride,
                             int dst_ld, int dst_row0, int n, int rows, int C, cudaStream_t st) {
  if (n <= 0 || rows <= 0) return 0;
  const long long total = (long long)n * rows * C;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
  copy_rows_strided_kernel<<<grid, 256, 0, st>>>(src, src_slot_stride, src_ld, src_row0, dst, dst_slot_stride, dst_ld, dst_row0, n, rows, C);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_advance_past_len_by(int* past_len, int n, const int* slot_ids, int by, cudaStream_t st) {
  if (n <= 0) return 0;
  advance_past_len_by_kernel<<<(n + 255) / 256, 256, 0, st>>>(past_len, n, slot_ids, by);
  CONAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace conan
