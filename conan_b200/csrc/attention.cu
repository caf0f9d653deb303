// Attention kernels of the path.  Both problems are far below tensor-core tile sizes
// (Emformer: 8 heads x 6 queries x <=56 keys x head_dim 10; aligner: 2 heads x 4 queries x
// <=250 keys x head_dim 128), so they are warp-level fp32 kernels with shuffle reductions and
// an fp32 softmax, batched over streams: one CTA per stream (x head for the aligner).
#include "kernels.cuh"

namespace conan {

namespace {

__device__ __forceinline__ int slot_of(const int* slot_ids, int i) { return slot_ids ? slot_ids[i] : i; }

// ---------------------------------------------------------------------------------------
// Emformer layer attention for one streaming step (TA:257-316 + 146-217 + 391-414).
//   qkv    [slot, seg+rc, 3D]  rows ordered [rc | utt], columns [Q | K | V]  (Q not yet scaled)
//   ring   [slot, ring_rows, 2D]  K|V of the last utterance rows; row (past_len + t) % ring_rows
//   keys are visited in the reference's order [rc | left context (oldest first) | utt]
// The kernel first appends this chunk's utterance K/V rows to the ring (the state update of
// _pack_state), then attends.  past_len itself is advanced once per step after all layers.
// One warp per head; lanes own keys; per-stream valid left context = min(lc, past_len[slot]),
// which removes the reference's batch-element-0 limitation (TA:392).
// ---------------------------------------------------------------------------------------
constexpr int EMF_MAX_KEYS = 64;   // rc + lc + seg = 56 at the reference config
constexpr int EMF_MAX_HD = 16;

__global__ void __launch_bounds__(256)
emformer_attention_kernel(const float* __restrict__ qkv, float* __restrict__ ring, const int* __restrict__ past_len,
                          RowView att, const int* __restrict__ slot_ids, int seg, int rc, int lc,
                          int ring_rows, int D, int heads, int ldq) {
  extern __shared__ float sm[];
  const int rows = seg + rc;
  const int slot = slot_of(slot_ids, blockIdx.x);
  const int past = past_len[slot];
  const int lc_len = min(lc, past);
  const int nkeys = rc + lc_len + seg;
  // rows are padded to D + 1 floats: lanes own keys, so an unpadded stride of 80 words would put every
  // lane on one of two banks (16-way conflicts on every K/V read)
  const int DS = D + 1;
  float* sK = sm;                         // [nkeys][DS]
  float* sV = sK + (size_t)(rc + lc + seg) * DS;
  float* sQ = sV + (size_t)(rc + lc + seg) * DS;   // [rows][D]
  const float* q_in = qkv + (long long)blockIdx.x * rows * ldq;      // compact scratch: index i, row stride ldq
  float* rg = ring + (long long)slot * ring_rows * 2 * D;
  const int tid = threadIdx.x;

  // stage Q, and K/V in key order
  for (int idx = tid; idx < rows * D; idx += blockDim.x) {
    int r = idx / D, c = idx % D;
    sQ[idx] = q_in[(long long)r * ldq + c];
  }
  for (int idx = tid; idx < nkeys * D; idx += blockDim.x) {
    int key = idx / D, c = idx % D;
    float kval, vval;
    if (key < rc) {                                   // look-ahead rows
      kval = q_in[(long long)key * ldq + D + c]; vval = q_in[(long long)key * ldq + 2 * D + c];
    } else if (key < rc + lc_len) {                   // cached left context, oldest first
      int logical = past - lc_len + (key - rc);
      int rr = logical % ring_rows;
      kval = rg[(long long)rr * 2 * D + c]; vval = rg[(long long)rr * 2 * D + D + c];
    } else {                                          // this chunk's utterance rows
      int r = rc + (key - rc - lc_len);
      kval = q_in[(long long)r * ldq + D + c]; vval = q_in[(long long)r * ldq + 2 * D + c];
    }
    sK[key * DS + c] = kval; sV[key * DS + c] = vval;
  }
  __syncthreads();
  // state update: ring rows (past + t) % ring_rows <- utterance K/V.  ring_rows >= lc + seg, so the
  // rows overwritten are older than the left context that was just staged.
  for (int idx = tid; idx < seg * 2 * D; idx += blockDim.x) {
    int t = idx / (2 * D), c = idx % (2 * D);
    int rr = (past + t) % ring_rows;
    rg[(long long)rr * 2 * D + c] = q_in[(long long)(rc + t) * ldq + D + c];
  }

  const int hd = D / heads;
  const float scaling = rsqrtf((float)hd);           // (input_dim // num_heads) ** -0.5
  const int warp = tid >> 5, lane = tid & 31;
  for (int h = warp; h < heads; h += (blockDim.x >> 5)) {
    for (int qr = 0; qr < rows; ++qr) {
      float qv[EMF_MAX_HD];
#pragma unroll
      for (int d = 0; d < EMF_MAX_HD; ++d) qv[d] = d < hd ? sQ[qr * D + h * hd + d] * scaling : 0.f;
      float sc[EMF_MAX_KEYS / 32];
      float mx = -INFINITY;
#pragma unroll
      for (int kk = 0; kk < EMF_MAX_KEYS / 32; ++kk) {
        int key = lane + kk * 32;
        float s = -INFINITY;
        if (key < nkeys) {
          s = 0.f;
#pragma unroll
          for (int d = 0; d < EMF_MAX_HD; ++d) if (d < hd) s = fmaf(qv[d], sK[key * DS + h * hd + d], s);
        }
        sc[kk] = s; mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      float den = 0.f;
#pragma unroll
      for (int kk = 0; kk < EMF_MAX_KEYS / 32; ++kk) {
        int key = lane + kk * 32;
        sc[kk] = key < nkeys ? expf(sc[kk] - mx) : 0.f;
        den += sc[kk];
      }
      den = warp_sum(den);
      float inv = 1.f / den;
      float o[EMF_MAX_HD];
#pragma unroll
      for (int d = 0; d < EMF_MAX_HD; ++d) o[d] = 0.f;
#pragma unroll
      for (int kk = 0; kk < EMF_MAX_KEYS / 32; ++kk) {
        int key = lane + kk * 32;
        if (key < nkeys) {
          float p = sc[kk] * inv;
#pragma unroll
          for (int d = 0; d < EMF_MAX_HD; ++d) if (d < hd) o[d] = fmaf(p, sV[key * DS + h * hd + d], o[d]);
        }
      }
#pragma unroll
      for (int d = 0; d < EMF_MAX_HD; ++d) {
        if (d < hd) {
          float v = warp_sum(o[d]);
          if (lane == 0) store_view(att, (long long)blockIdx.x * att.slot_stride + (long long)qr * att.row_stride + h * hd + d, v);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Aligner cross-attention (nn.MultiheadAttention, prosody_util.py:108-127): queries are this
// chunk's frames, keys/values are the session-cached projections of the prosody tokens.
//   q [slot, rows, H] (unscaled), cache [slot, layer, tp_max, 2H] (K | V), kpm [slot, tp_max]
// grid (stream, head); one warp per query row; scores staged in shared memory.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
cross_attention_kernel(const float* __restrict__ q, const float* __restrict__ cache, const float* __restrict__ kpm,
                       const int* __restrict__ n_keys, RowView out, const int* __restrict__ slot_ids, int rows,
                       int H, int heads, int layer, int n_layers, int tp_max) {
  extern __shared__ float sc[];                      // [rows][tp_max]
  const int slot = slot_of(slot_ids, blockIdx.x), h = blockIdx.y;
  const int hd = H / heads;
  const int Tp = n_keys[slot];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* kv = cache + ((long long)slot * n_layers + layer) * tp_max * 2 * H;
  const float* pm = kpm + (long long)slot * tp_max;
  const float scaling = sqrtf(1.0f / (float)hd);     // q * math.sqrt(1.0 / head_dim)
  for (int r = warp; r < rows; r += (blockDim.x >> 5)) {
    float* s = sc + (size_t)r * tp_max;
    const float* qr = q + ((long long)blockIdx.x * rows + r) * H + h * hd;
    float qv[4];                                      // hd = 128 -> 4 per lane
    for (int j = 0; j < 4; ++j) { int d = lane + 32 * j; qv[j] = d < hd ? qr[d] * scaling : 0.f; }
    float mx = -INFINITY;
    for (int key = 0; key < Tp; ++key) {
      const float* kr = kv + (long long)key * 2 * H + h * hd;
      float a = 0.f;
      for (int j = 0; j < 4; ++j) { int d = lane + 32 * j; if (d < hd) a = fmaf(qv[j], kr[d], a); }
      a = warp_sum(a);
      if (pm[key] != 0.f) a = -INFINITY;              // key_padding_mask -> -inf
      if (lane == 0) s[key] = a;
      mx = fmaxf(mx, a);
    }
    __syncwarp();
    float den = 0.f;
    for (int key = lane; key < Tp; key += 32) { float e = expf(s[key] - mx); s[key] = e; den += e; }
    den = warp_sum(den);
    __syncwarp();
    float inv = 1.f / den;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    for (int key = 0; key < Tp; ++key) {
      float p = s[key] * inv;
      const float* vr = kv + (long long)key * 2 * H + H + h * hd;
      for (int j = 0; j < 4; ++j) { int d = lane + 32 * j; if (d < hd) o[j] = fmaf(p, vr[d], o[j]); }
    }
    const long long orow = (long long)blockIdx.x * out.slot_stride + (long long)r * out.row_stride + h * hd;
    for (int j = 0; j < 4; ++j) { int d = lane + 32 * j; if (d < hd) store_view(out, orow + d, o[j]); }
  }
}

}  // namespace

int launch_emformer_attention(const float* qkv, float* kv_ring, const int* past_len, RowView att, int n,
                              const int* slot_ids, int seg, int rc, int lc, int ring_rows, int D, int heads, int ld_qkv,
                              cudaStream_t st) {
  if (n <= 0) return 0;
  if (rc + lc + seg > EMF_MAX_KEYS || D / heads > EMF_MAX_HD) { set_error("emformer_attention: key count or head_dim above compiled limits"); return 1; }
  if (ring_rows < lc + seg) { set_error("emformer_attention: ring too short"); return 1; }
  size_t sh = ((size_t)2 * (rc + lc + seg) * (D + 1) + (size_t)(seg + rc) * D) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(emformer_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr_set = true; }
  if (sh > 96 * 1024) { set_error("emformer_attention: shared memory above 96 KB"); return 1; }
  emformer_attention_kernel<<<n, 256, sh, st>>>(qkv, kv_ring, past_len, att, slot_ids, seg, rc, lc, ring_rows, D, heads, ld_qkv);
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_cross_attention(const float* q, const float* kv_cache, const float* kpm, const int* n_keys, RowView out, int n,
                           const int* slot_ids, int rows, int H, int heads, int layer, int n_layers, int tp_max, cudaStream_t st) {
  if (n <= 0) return 0;
  if (H / heads > 128) { set_error("cross_attention: head_dim above 128"); return 1; }
  size_t sh = (size_t)rows * tp_max * sizeof(float);
  if (sh > 48 * 1024) { set_error("cross_attention: too many keys for the score buffer"); return 1; }
  cross_attention_kernel<<<dim3(n, heads), 128, sh, st>>>(q, kv_cache, kpm, n_keys, out, slot_ids, rows, H, heads, layer, n_layers, tp_max);
  CONAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace conan
