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
constexpr int EMF_MAX_ROWS = 8;

// FIXED: the reference configuration (segment 4, right context 2, left context 50, ring 56 rows, D = 80, 8 heads; LDQ = row stride of
// qkv) as compile-time constants.  With run-time shapes 45 % of this kernel's instructions were index arithmetic (IMAD / IADD3 / LEA /
// ISETP / IABS: strides, divisions and modulos by parameters) and it ran issue-bound; the generic instantiation stays for other configs.
template <bool FIXED, int LDQ>
__global__ void __launch_bounds__(256, 4)
emformer_attention_kernel(const float* __restrict__ qkv, float* __restrict__ ring, const int* __restrict__ past_len,
                          RowView att, const int* __restrict__ slot_ids, int seg_, int rc_, int lc_,
                          int ring_rows_, int D_, int heads_, int ldq_, EmfAttnEpilogue ep, int fused) {
  extern __shared__ float sm[];
  const int seg = FIXED ? 4 : seg_, rc = FIXED ? 2 : rc_, lc = FIXED ? 50 : lc_, ring_rows = FIXED ? 56 : ring_rows_;
  const int D = FIXED ? 80 : D_, heads = FIXED ? 8 : heads_, ldq = FIXED ? LDQ : ldq_;
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

  // stage Q, and K|V in key order.  Every key's K|V is one contiguous run of 2D floats (columns [D, 3D) of a
  // qkv row, or a ring row), moved as float4 with all loads of a thread issued before the first store.
  const int hd = D / heads;
  const float scaling = rsqrtf((float)hd);           // (input_dim // num_heads) ** -0.5, applied to Q before Q.K^T
  for (int idx = tid; idx < rows * D; idx += blockDim.x) {
    int r = idx / D, c = idx - r * D;
    sQ[idx] = q_in[(long long)r * ldq + c] * scaling;
  }
  {
    const int v4_per_key = (2 * D) / 4;                 // 40
    const int items = nkeys * v4_per_key;
    constexpr int UNR = 10;                             // 56 keys * 40 / 256 threads = 8.75 items per thread
    float4 buf[UNR];
    int kk[UNR], qq[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int item = tid + u * blockDim.x;
      kk[u] = -1;
      if (item < items) {
        const int key = item / v4_per_key, q4 = item - key * v4_per_key;
        const float* src;
        if (key < rc) src = q_in + (long long)key * ldq + D;                                   // look-ahead rows
        else if (key < rc + lc_len) src = rg + (long long)((past - lc_len + (key - rc)) % ring_rows) * 2 * D;   // left context, oldest first
        else src = q_in + (long long)(rc + (key - rc - lc_len)) * ldq + D;                     // this chunk's utterance rows
        buf[u] = *reinterpret_cast<const float4*>(src + q4 * 4);
        kk[u] = key; qq[u] = q4;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (kk[u] >= 0) {
        const int c = qq[u] * 4;
        float* dst = (c < D ? sK + kk[u] * DS + c : sV + kk[u] * DS + (c - D));
        dst[0] = buf[u].x; dst[1] = buf[u].y; dst[2] = buf[u].z; dst[3] = buf[u].w;
      }
    }
  }
  __syncthreads();
  // state update: ring rows (past + t) % ring_rows <- utterance K|V.  ring_rows >= lc + seg, so the
  // rows overwritten are older than the left context that was just staged.
  for (int idx = tid; idx < seg * (2 * D) / 4; idx += blockDim.x) {
    const int t = idx / ((2 * D) / 4), q4 = idx - t * ((2 * D) / 4);
    const int rr = (past + t) % ring_rows;
    *reinterpret_cast<float4*>(rg + (long long)rr * 2 * D + q4 * 4) =
        *reinterpret_cast<const float4*>(q_in + (long long)(rc + t) * ldq + D + q4 * 4);
  }

  // One warp per head.  Lanes own keys for the scores (each K element is read once for all query rows); the
  // probabilities then overwrite this head's K columns (private to the warp), and P.V runs with lanes over
  // the rows x head_dim outputs.
  const int warp = tid >> 5, lane = tid & 31;
  for (int h = warp; h < heads; h += (blockDim.x >> 5)) {
    const int c0 = h * hd;
    float acc[EMF_MAX_ROWS][EMF_MAX_KEYS / 32];
#pragma unroll
    for (int r = 0; r < EMF_MAX_ROWS; ++r)
#pragma unroll
      for (int kk = 0; kk < EMF_MAX_KEYS / 32; ++kk) acc[r][kk] = 0.f;
    const int key0 = min(lane, nkeys - 1), key1 = min(lane + 32, nkeys - 1);
    for (int d = 0; d < hd; ++d) {
      const float k0 = sK[key0 * DS + c0 + d], k1 = sK[key1 * DS + c0 + d];
#pragma unroll
      for (int r = 0; r < EMF_MAX_ROWS; ++r) {
        if (r < rows) {
          const float qv = sQ[r * D + c0 + d];
          acc[r][0] = fmaf(qv, k0, acc[r][0]);
          acc[r][1] = fmaf(qv, k1, acc[r][1]);
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < EMF_MAX_ROWS; ++r) {
      if (r < rows) {
        const float s0 = lane < nkeys ? acc[r][0] : -INFINITY, s1 = lane + 32 < nkeys ? acc[r][1] : -INFINITY;
        const float mx = warp_max(fmaxf(s0, s1));
        const float e0 = lane < nkeys ? expf(s0 - mx) : 0.f, e1 = lane + 32 < nkeys ? expf(s1 - mx) : 0.f;
        const float inv = 1.f / warp_sum(e0 + e1);
        if (lane < nkeys) sK[lane * DS + c0 + r] = e0 * inv;
        if (lane + 32 < nkeys) sK[(lane + 32) * DS + c0 + r] = e1 * inv;
      }
    }
    __syncwarp();
    for (int o = lane; o < rows * hd; o += 32) {
      const int r = o / hd, d = o - r * hd;
      float v = 0.f;
      for (int key = 0; key < nkeys; ++key) v = fmaf(sK[key * DS + c0 + r], sV[key * DS + c0 + d], v);
      if (fused) sQ[r * D + c0 + d] = v;          // this head's query columns are dead: the attention output takes their place
      else store_view(att, (long long)blockIdx.x * att.slot_stride + (long long)r * att.row_stride + c0 + d, v);
    }
  }
  if (!fused) return;
  // ---- fused tail: out_proj (fp32) + bias + residual -> r1; LayerNorm(r1) -> fn  (rows <= 8 warps)
  __syncthreads();
  // K / V are dead: their space takes the transposed out_proj weight [D][D] (one coalesced pass) and the rows of r1
  float* sW = sK;
  float* sR = sK + D * D;
  for (int idx = tid * 4; idx < D * D; idx += blockDim.x * 4)
    *reinterpret_cast<float4*>(sW + idx) = __ldg(reinterpret_cast<const float4*>(ep.wt + idx));
  __syncthreads();
  const long long row0 = (long long)blockIdx.x * rows;
  for (int o = tid; o < rows * D; o += blockDim.x) {
    const int r = o / D, c = o - r * D;
    float acc = ep.bias[c];
#pragma unroll 8
    for (int k = 0; k < D; ++k) acc = fmaf(sQ[r * D + k], sW[k * D + c], acc);      // sQ broadcast, sW consecutive over c
    acc += ep.x_res[(row0 + r) * ep.ld + c];
    ep.r1[(row0 + r) * ep.ld + c] = acc;
    sR[r * D + c] = acc;
  }
  __syncthreads();
  for (int r = warp; r < rows; r += (blockDim.x >> 5)) {
    float s1 = 0.f;
    for (int c = lane; c < D; c += 32) s1 += sR[r * D + c];
    const float mean = warp_sum(s1) / D;
    float s2 = 0.f;
    for (int c = lane; c < D; c += 32) { const float d = sR[r * D + c] - mean; s2 += d * d; }
    const float rstd = 1.f / sqrtf(warp_sum(s2) / D + ep.eps);
    const long long o = (long long)blockIdx.x * ep.fn.slot_stride + (long long)(ep.fn.row0 + r) * ep.fn.row_stride;
    for (int c = lane; c < D; c += 32) store_view(ep.fn, o + c, (sR[r * D + c] - mean) * rstd * ep.ln_g[c] + ep.ln_b[c]);
  }
}

// ---------------------------------------------------------------------------------------
// Aligner cross-attention (nn.MultiheadAttention, prosody_util.py:108-127): queries are this
// chunk's frames, keys/values are the session-cached projections of the prosody tokens.
//   q [i, rows, H] (unscaled, compact), cache [slot, layer, tp_max, 2H] (K | V), kpm [slot, tp_max]
// One warp per (stream, head) handles all query rows: scores with lanes over keys (each K row is read once
// for the 4 queries), fp32 softmax by warp shuffles, then P.V with lanes over the head dimension (coalesced
// V rows).  8 warps per CTA.
// ---------------------------------------------------------------------------------------
constexpr int XA_MAX_ROWS = 8;
constexpr int XA_WARPS = 8;

__global__ void __launch_bounds__(XA_WARPS * 32)
cross_attention_kernel(const float* __restrict__ q, const float* __restrict__ cache, const float* __restrict__ kpm,
                       const int* __restrict__ n_keys, RowView out, const int* __restrict__ slot_ids, int n, int rows,
                       int H, int heads, int layer, int n_layers, int tp_max) {
  extern __shared__ float xsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd = H / heads;                                          // 128
  float* sq = xsm + (size_t)warp * (XA_MAX_ROWS * hd + XA_MAX_ROWS * tp_max);   // [rows][hd] scaled queries
  float* sc = sq + XA_MAX_ROWS * hd;                                  // [rows][tp_max] scores -> probabilities
  const int pair = blockIdx.x * XA_WARPS + warp;
  if (pair >= n * heads) return;
  const int i = pair / heads, h = pair - i * heads;
  const int slot = slot_of(slot_ids, i);
  const int Tp = n_keys[slot];
  const float* kv = cache + ((long long)slot * n_layers + layer) * tp_max * 2 * H;
  const float* pm = kpm + (long long)slot * tp_max;
  const float scaling = sqrtf(1.0f / (float)hd);                      // q * math.sqrt(1.0 / head_dim)
  for (int idx = lane; idx < rows * hd; idx += 32) {
    const int r = idx / hd, d = idx - r * hd;
    sq[r * hd + d] = q[((long long)i * rows + r) * H + h * hd + d] * scaling;
  }
  __syncwarp();
  // ---- scores: lane <- key
  for (int key = lane; key < Tp; key += 32) {
    const float4* kr = reinterpret_cast<const float4*>(kv + (long long)key * 2 * H + h * hd);
    float acc[XA_MAX_ROWS];
#pragma unroll
    for (int r = 0; r < XA_MAX_ROWS; ++r) acc[r] = 0.f;
    for (int d4 = 0; d4 < hd / 4; ++d4) {
      const float4 k4 = kr[d4];
#pragma unroll
      for (int r = 0; r < XA_MAX_ROWS; ++r) {
        if (r < rows) {
          const float4 q4 = *reinterpret_cast<const float4*>(&sq[r * hd + d4 * 4]);
          acc[r] = fmaf(q4.x, k4.x, fmaf(q4.y, k4.y, fmaf(q4.z, k4.z, fmaf(q4.w, k4.w, acc[r]))));
        }
      }
    }
    const bool masked = pm[key] != 0.f;                               // key_padding_mask -> -inf
#pragma unroll
    for (int r = 0; r < XA_MAX_ROWS; ++r)
      if (r < rows) sc[r * tp_max + key] = masked ? -INFINITY : acc[r];
  }
  __syncwarp();
  // ---- softmax per query row
  for (int r = 0; r < rows; ++r) {
    float mx = -INFINITY;
    for (int key = lane; key < Tp; key += 32) mx = fmaxf(mx, sc[r * tp_max + key]);
    mx = warp_max(mx);
    float den = 0.f;
    for (int key = lane; key < Tp; key += 32) { const float e = expf(sc[r * tp_max + key] - mx); sc[r * tp_max + key] = e; den += e; }
    den = warp_sum(den);
    const float inv = 1.f / den;
    for (int key = lane; key < Tp; key += 32) sc[r * tp_max + key] *= inv;
  }
  __syncwarp();
  // ---- P.V: lane <- 4 consecutive head dims
  for (int d0 = lane * 4; d0 < hd; d0 += 128) {
    float4 o[XA_MAX_ROWS];
#pragma unroll
    for (int r = 0; r < XA_MAX_ROWS; ++r) o[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int key = 0; key < Tp; ++key) {
      const float4 v4 = *reinterpret_cast<const float4*>(kv + (long long)key * 2 * H + H + h * hd + d0);
#pragma unroll
      for (int r = 0; r < XA_MAX_ROWS; ++r) {
        if (r < rows) {
          const float p = sc[r * tp_max + key];
          o[r].x = fmaf(p, v4.x, o[r].x); o[r].y = fmaf(p, v4.y, o[r].y); o[r].z = fmaf(p, v4.z, o[r].z); o[r].w = fmaf(p, v4.w, o[r].w);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < XA_MAX_ROWS; ++r) {
      if (r < rows) {
        const long long orow = (long long)i * out.slot_stride + (long long)r * out.row_stride + h * hd + d0;
        store_view(out, orow, o[r].x); store_view(out, orow + 1, o[r].y); store_view(out, orow + 2, o[r].z); store_view(out, orow + 3, o[r].w);
      }
    }
  }
}

// Same operator with this (stream, head)'s K and V rows staged in shared memory by cp.async (every 16-byte copy of
// the CTA is in flight at once, which is what the HBM-resident cache needs), used while 2 * keys * (hd + 4) floats
// fit.  One CTA of 4 warps per (stream, head); a warp owns a query row.  K/V rows are padded by 4 floats so the
// per-lane float4 row reads of the score loop are conflict-free.
constexpr int XS_THREADS = 128;

__global__ void __launch_bounds__(XS_THREADS)
cross_attention_staged_kernel(const float* __restrict__ q, const float* __restrict__ cache, const float* __restrict__ kpm,
                              const int* __restrict__ n_keys, RowView out, const int* __restrict__ slot_ids, int rows,
                              int H, int heads, int layer, int n_layers, int tp_max) {
  extern __shared__ __align__(16) float xs[];
  const int hd = H / heads, HS = hd + 4;
  float* sK = xs;                                  // [tp_max][HS]
  float* sV = sK + (size_t)tp_max * HS;            // [tp_max][HS]
  float* sq = sV + (size_t)tp_max * HS;            // [rows][hd]
  float* sp = sq + (size_t)rows * hd;              // [rows][tp_max]
  const int i = blockIdx.x / heads, h = blockIdx.x - i * heads;
  const int slot = slot_of(slot_ids, i);
  const int Tp = n_keys[slot];
  const float* kv = cache + ((long long)slot * n_layers + layer) * tp_max * 2 * H;
  const float* pm = kpm + (long long)slot * tp_max;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int v4 = hd / 4;
  for (int idx = tid; idx < Tp * 2 * v4; idx += XS_THREADS) {
    const int key = idx / (2 * v4), rem = idx - key * 2 * v4;
    const int isv = rem >= v4, c4 = rem - isv * v4;
    const float* src = kv + (long long)key * 2 * H + isv * H + h * hd + c4 * 4;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared((isv ? sV : sK) + key * HS + c4 * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  }
  const float scaling = sqrtf(1.0f / (float)hd);                      // q * math.sqrt(1.0 / head_dim)
  for (int idx = tid; idx < rows * hd; idx += XS_THREADS) {
    const int r = idx / hd, d = idx - r * hd;
    sq[idx] = q[((long long)i * rows + r) * H + h * hd + d] * scaling;
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  for (int r = warp; r < rows; r += XS_THREADS / 32) {
    float* p = sp + (size_t)r * tp_max;
    const float4* q4 = reinterpret_cast<const float4*>(sq + (size_t)r * hd);
    float mx = -INFINITY;
    for (int key = lane; key < Tp; key += 32) {
      const float4* k4 = reinterpret_cast<const float4*>(sK + (size_t)key * HS);
      float a = 0.f;
#pragma unroll 8
      for (int d = 0; d < v4; ++d) {
        const float4 kk = k4[d], qq = q4[d];
        a = fmaf(qq.x, kk.x, fmaf(qq.y, kk.y, fmaf(qq.z, kk.z, fmaf(qq.w, kk.w, a))));
      }
      if (pm[key] != 0.f) a = -INFINITY;                              // key_padding_mask
      p[key] = a; mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float den = 0.f;
    for (int key = lane; key < Tp; key += 32) { const float e = expf(p[key] - mx); p[key] = e; den += e; }
    const float inv = 1.f / warp_sum(den);
    __syncwarp();
    for (int d0 = lane * 4; d0 < hd; d0 += 128) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int key = 0; key < Tp; ++key) {
        const float pk = p[key];
        const float4 v = *reinterpret_cast<const float4*>(sV + (size_t)key * HS + d0);
        o.x = fmaf(pk, v.x, o.x); o.y = fmaf(pk, v.y, o.y); o.z = fmaf(pk, v.z, o.z); o.w = fmaf(pk, v.w, o.w);
      }
      const long long orow = (long long)i * out.slot_stride + (long long)r * out.row_stride + h * hd + d0;
      store_view(out, orow, o.x * inv); store_view(out, orow + 1, o.y * inv);
      store_view(out, orow + 2, o.z * inv); store_view(out, orow + 3, o.w * inv);
    }
  }
}

}  // namespace

int launch_emformer_attention(const float* qkv, float* kv_ring, const int* past_len, RowView att, int n,
                              const int* slot_ids, int seg, int rc, int lc, int ring_rows, int D, int heads, int ld_qkv,
                              cudaStream_t st, const EmfAttnEpilogue* ep) {
  if (n <= 0) return 0;
  if (rc + lc + seg > EMF_MAX_KEYS || seg + rc > EMF_MAX_ROWS || seg + rc > D / heads) { set_error("emformer_attention: key count / query rows above compiled limits (rows <= min(8, head_dim))"); return 1; }
  if (D % 4 != 0 || ld_qkv % 4 != 0 || (rc + lc + seg) * ((2 * D) / 4) > 10 * 256) { set_error("emformer_attention: D / ld must be multiples of 4 and keys*2D/4 <= 2560"); return 1; }
  if (ring_rows < lc + seg) { set_error("emformer_attention: ring too short"); return 1; }
  if (ep && (size_t)D * D + (size_t)(seg + rc) * D > (size_t)2 * (rc + lc + seg) * (D + 1)) { set_error("emformer_attention: fused tail does not fit in the K/V staging area"); return 1; }
  size_t sh = ((size_t)2 * (rc + lc + seg) * (D + 1) + (size_t)(seg + rc) * D) * sizeof(float);
  if (sh > 96 * 1024) { set_error("emformer_attention: shared memory above 96 KB"); return 1; }
  const bool fixed = seg == 4 && rc == 2 && lc == 50 && ring_rows == 56 && D == 80 && heads == 8 && (ld_qkv == 256 || ld_qkv == 240);
  auto launch = [&](auto kern) -> int {
    static DeviceOnce once;                       // (one per instantiation: the lambda's operator() is a template)
    if (device_once(once, nullptr, [&](int*) {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
          if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
          return 0;
        }))
      return 1;
    kern<<<n, 256, sh, st>>>(qkv, kv_ring, past_len, att, slot_ids, seg, rc, lc, ring_rows, D, heads, ld_qkv, ep ? *ep : EmfAttnEpilogue{}, ep ? 1 : 0);
    return 0;
  };
  int rc_l;
  if (fixed && ld_qkv == 256) rc_l = launch(emformer_attention_kernel<true, 256>);
  else if (fixed) rc_l = launch(emformer_attention_kernel<true, 240>);
  else rc_l = launch(emformer_attention_kernel<false, 0>);
  if (rc_l) return 1;
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_cross_attention(const float* q, const float* kv_cache, const float* kpm, const int* n_keys, RowView out, int n,
                           const int* slot_ids, int rows, int H, int heads, int layer, int n_layers, int tp_max, cudaStream_t st) {
  if (n <= 0) return 0;
  const int hd = H / heads;
  if (rows > XA_MAX_ROWS || hd % 4 != 0) { set_error("cross_attention: rows above 8 or head_dim not a multiple of 4"); return 1; }
  const size_t sh_staged = ((size_t)2 * tp_max * (hd + 4) + (size_t)rows * hd + (size_t)rows * tp_max) * sizeof(float);
  if (sh_staged <= 100 * 1024 && H % 4 == 0) {
    static DeviceOnce once;
    if (device_once(once, nullptr, [&](int*) {
          cudaError_t e = cudaFuncSetAttribute(cross_attention_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
          if (e == cudaSuccess) e = cudaFuncSetAttribute(cross_attention_staged_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
          return 0;
        }))
      return 1;
    cross_attention_staged_kernel<<<n * heads, XS_THREADS, sh_staged, st>>>(q, kv_cache, kpm, n_keys, out, slot_ids, rows, H, heads, layer,
                                                                            n_layers, tp_max);
    CONAN_CHECK_LAUNCH();
    return 0;
  }
  size_t sh = (size_t)XA_WARPS * (XA_MAX_ROWS * hd + XA_MAX_ROWS * tp_max) * sizeof(float);
  if (sh > 200 * 1024) { set_error("cross_attention: too many keys for the score buffer"); return 1; }
  static DeviceOnce once_general;
  if (device_once(once_general, nullptr, [&](int*) {
        cudaError_t e = cudaFuncSetAttribute(cross_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
        return 0;
      }))
    return 1;
  const int pairs = n * heads;
  cross_attention_kernel<<<(pairs + XA_WARPS - 1) / XA_WARPS, XA_WARPS * 32, sh, st>>>(q, kv_cache, kpm, n_keys, out, slot_ids, n, rows, H,
                                                                                       heads, layer, n_layers, tp_max);
  CONAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace conan
