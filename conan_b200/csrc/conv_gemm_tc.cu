// Implicit-GEMM causal Conv1d on the 5th-generation tensor cores (sm_100a: tcgen05 + TMEM + TMA).
//
//   D[m, n] = sum_kk A[m, kk] * W[n, kk],  M = n_streams * L, N = cout, K = k * cin (tap-major)
//   A[m, j*cin + c] = X[slot(m), row0 + t(m) + j*dil, c]
//
// The vocoder keeps the input of every causal conv in a per-slot fp16 context buffer
// [slot, H + L, cin] (H history rows + the L rows of this chunk), so for one tap j and one
// 64-channel slice the 128 A rows of a tile are contiguous row ranges of that buffer: the
// producer warp fetches them with ONE TMA box {BK channels, TT rows, 128/TT streams} per K-block
// (TT = largest power of two <= 128 dividing L; the engine's working buffers are compact over the
// ready list, so consecutive streams are adjacent) straight into the swizzled K-major layout
// tcgen05.mma reads.  No im2col
// buffer exists anywhere; dilation is a row offset of the box.  Weights are pre-packed
// [cout, k*cin] fp16 and fetched as {BK, BN} boxes.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA
// issuer, warps 2..5 = epilogue (one TMEM lane quarter each).  The fp32 accumulator tile
// 128 x BN lives in TMEM; the epilogue reads it with tcgen05.ld 32x32b and applies the fused
// chain of the path: bias, scale, activation, residual add (fp32 stream), row mask, MRF
// 1/3 scale + accumulate, then writes the fp32 stream and/or the LeakyReLU'd fp16 copy that is
// the next conv's context buffer (pixel shuffle is folded into the weight row order, so an
// upsampling conv is just this kernel with a wider output row).
#include <cuda.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace conan {

namespace {


struct TcEpi {
  const float* bias; float scale; int act; float slope;
  const void* res; long long res_slot_stride; int res_row_stride;      // fp32 stream, or fp16 activated context rows (res_is_half)
  const float* rowmask; int mask_slot_stride; float out_scale;
  void* y; long long y_slot_stride; int y_row_stride, y_row0; int accumulate;
  __half* y2; long long y2_slot_stride; int y2_row_stride, y2_row0; int act2; float slope2;
  float acc_scale; long long y2_lo_off;      // y2_lo_off != 0: y2 is a split fp16 pair (hi, lo = fp16(v - hi))
  int res_is_half; float res_inv_slope;      // residual = inverse-LeakyReLU of the fp16 rows the producer wrote for the next conv
  const __half* res2; long long res2_slot_stride; int res2_row_stride;   // second fp16 residual (running MRF sum), RES_F16 path only
  int y_is_half;
};

struct TcArgs {
  int n_streams, L, TT, cin, k, dil, cout, row0;
  int kblocks;               // nseg * k * cin / BK
  int n_tiles;               // cout / BN
  int m_tiles;               // 128-row tiles over (stream, time)
  int num_tiles;             // CTA tiles: ceil(m_tiles / MT) * n_tiles
  int nseg;                  // 1, or 3 for split operands: K' = [x_hi*W_hi | x_hi*W_lo | x_lo*W_hi]
  int lo_slot_off;           // slot offset of the lo plane of a split x
  TcEpi e;
};

// Fused epilogue of one accumulator row (thread = tile row = TMEM lane): waits for the accumulator,
// then per 16-column chunk: bias, scale, activation, residual, mask, 1/3-scale (+ old output), fp32
// store and/or activated fp16 store.  Activations on this engine are none / relu / leaky (branch-free:
// v > 0 ? v : v * slope with slope 1, 0 or the leaky slope) or exact-erf GELU (warp-uniform branch).
// Latency hiding: the bias comes from shared memory (staged once per CTA); the residual (fp32 stream)
// does not depend on the accumulator, so PF chunks of it are fetched BEFORE the accumulator wait and the
// window is kept PF chunks ahead -- for BN <= 64 that is the whole row, i.e. every HBM load of the tile is
// in flight at once.
enum ResKind : int { RES_NONE = 0, RES_F32 = 1, RES_F16 = 2 };

template <int BN, int RK>
__device__ __forceinline__ void epilogue_rows_k(const TcEpi& e, const float* __restrict__ s_bias, uint32_t tmem_lane_base, int nbase,
                                                bool valid, int slot, int t, uint64_t* acc_full_bar, uint32_t parity) {
  constexpr int NCH = BN / 16;
  constexpr int PF = NCH < 4 ? NCH : 4;                // chunks of residual in flight (4 x 4 float4 = 64 registers at most)
  const float rm = (e.rowmask && valid) ? e.rowmask[(long long)slot * e.mask_slot_stride + t] : 1.f;
  const long long res_off = (long long)slot * e.res_slot_stride + (long long)t * e.res_row_stride + nbase;
  const float* resp = (RK == RES_F32 && valid) ? reinterpret_cast<const float*>(e.res) + res_off : nullptr;
  const __half* resh = (RK == RES_F16 && valid) ? reinterpret_cast<const __half*>(e.res) + res_off : nullptr;
  const long long y_off = (long long)slot * e.y_slot_stride + (long long)(e.y_row0 + t) * e.y_row_stride + nbase;
  float* yp = (e.y && valid && !e.y_is_half) ? reinterpret_cast<float*>(e.y) + y_off : nullptr;
  __half* yh = (e.y && valid && e.y_is_half) ? reinterpret_cast<__half*>(e.y) + y_off : nullptr;
  const __half* res2p = (RK == RES_F16 && e.res2 && valid)
                            ? e.res2 + (long long)slot * e.res2_slot_stride + (long long)t * e.res2_row_stride + nbase : nullptr;
  __half* y2p = (e.y2 && valid) ? e.y2 + (long long)slot * e.y2_slot_stride + (long long)(e.y2_row0 + t) * e.y2_row_stride + nbase : nullptr;
  const bool acc_old = e.accumulate && yp;
  const bool gelu = e.act == ACT_GELU;
  const float s1 = e.act == ACT_NONE ? 1.f : (e.act == ACT_RELU ? 0.f : e.slope);
  const float s2 = e.act2 == ACT_NONE ? 1.f : (e.act2 == ACT_RELU ? 0.f : e.slope2);
  const float f = rm * e.out_scale;
  constexpr int RW = RK == RES_F32 ? 4 : (RK == RES_F16 ? 2 : 1);     // 16-byte registers per chunk of residual
  float4 rbuf[PF][RW];
  float4 r2buf[RK == RES_F16 ? PF : 1][2];                   // second residual (fp16 running sum), same look-ahead
  // 16 residual values of a chunk: four float4 (fp32 stream) or two 16-byte loads of halfs
  auto fetch_res = [&](float4 (&dst)[RW], int c0) {
    if (RK == RES_F16) {
#pragma unroll
      for (int i = 0; i < RW; ++i)
        dst[i] = resh ? *(reinterpret_cast<const float4*>(resh + c0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (RK == RES_F32) {
#pragma unroll
      for (int i = 0; i < RW; ++i)
        dst[i] = resp ? *(reinterpret_cast<const float4*>(resp + c0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto fetch_res2 = [&](float4 (&dst)[2], int c0) {
    dst[0] = *(reinterpret_cast<const float4*>(res2p + c0));
    dst[1] = *(reinterpret_cast<const float4*>(res2p + c0) + 1);
  };
  if (RK != RES_NONE) {
#pragma unroll
    for (int c = 0; c < PF; ++c) fetch_res(rbuf[c], c * 16);
    if (RK == RES_F16 && res2p) {
#pragma unroll
      for (int c = 0; c < PF; ++c) fetch_res2(r2buf[c], c * 16);
    }
  }
  mbar_wait(acc_full_bar, parity);
  tc_fence_after();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int c0 = ch * 16;
    uint32_t acc[16];
    tc_ld_32x32b_x16(tmem_lane_base + (uint32_t)c0, acc);
    float v[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(s_bias + nbase + c0 + 4 * i);
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
      float rr[4] = {0.f, 0.f, 0.f, 0.f};
      if (RK == RES_F16) {                                   // halfs 4i .. 4i+3 of the chunk, LeakyReLU undone
        const __half2* hp = reinterpret_cast<const __half2*>(&rbuf[ch % PF][(i >> 1) % RW]) + (i & 1) * 2;
        const float2 a = __half22float2(hp[0]), b = __half22float2(hp[1]);
        rr[0] = a.x; rr[1] = a.y; rr[2] = b.x; rr[3] = b.y;
        const float inv = e.res_inv_slope != 0.f ? e.res_inv_slope : 1.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) rr[u] = rr[u] < 0.f ? rr[u] * inv : rr[u];
        if (res2p) {
          const __half2* sp = reinterpret_cast<const __half2*>(&r2buf[ch % PF][i >> 1]) + (i & 1) * 2;
          const float2 c = __half22float2(sp[0]), d = __half22float2(sp[1]);
          rr[0] += c.x; rr[1] += c.y; rr[2] += d.x; rr[3] += d.y;
        }
      } else if (RK == RES_F32) {
        const float4 r4 = rbuf[ch % PF][i % RW];
        rr[0] = r4.x; rr[1] = r4.y; rr[2] = r4.z; rr[3] = r4.w;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float x = fmaf(__uint_as_float(acc[4 * i + u]), e.acc_scale, bb[u]) * e.scale;
        if (gelu) x = 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));     // warp-uniform branch (exact-erf GELU)
        else x = x > 0.f ? x : x * s1;
        v[4 * i + u] = (x + rr[u]) * f;
      }
    }
    if (RK != RES_NONE && ch + PF < NCH) {
      fetch_res(rbuf[ch % PF], (ch + PF) * 16);
      if (RK == RES_F16 && res2p) fetch_res2(r2buf[ch % PF], (ch + PF) * 16);
    }
    if (yh) {                                                // fp16 primary output (running MRF sum)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        __half2 h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) h[u] = __floats2half2_rn(v[8 * i + 2 * u], v[8 * i + 2 * u + 1]);
        *(reinterpret_cast<uint4*>(yh + c0) + i) = *reinterpret_cast<uint4*>(h);
      }
    }
    if (yp) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4* dst = reinterpret_cast<float4*>(yp + c0) + i;
        if (acc_old) { const float4 o = *dst; v[4 * i] += o.x; v[4 * i + 1] += o.y; v[4 * i + 2] += o.z; v[4 * i + 3] += o.w; }
        *dst = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
    if (y2p) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        __half2 h[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float x0 = v[8 * i + 2 * u], x1 = v[8 * i + 2 * u + 1];
          x0 = x0 > 0.f ? x0 : x0 * s2; x1 = x1 > 0.f ? x1 : x1 * s2;
          h[u] = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(h[u]);
          l[u] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
        }
        *(reinterpret_cast<uint4*>(y2p + c0) + i) = *reinterpret_cast<uint4*>(h);
        if (e.y2_lo_off) *(reinterpret_cast<uint4*>(y2p + e.y2_lo_off + c0) + i) = *reinterpret_cast<uint4*>(l);
      }
    }
  }
}

// Lean epilogue for the vocoder layers (fp16 context rows in and out, no scale / mask / fp32 stream):
//   v = acc + bias [+ inverse-LeakyReLU(res)] [+ res2];  v *= out_scale;  y = fp16(v) and/or y2 = fp16(LeakyReLU(v))
// with the (inverse) LeakyReLU as min / max (slopes in (0, 1)).  The short-K layers are bound by the SM's instruction issue
// (a 128 x 128 tile is ~16 K elements per 1.5 K cycles of MMA), so this path spends ~4-8 instructions per element instead of
// the general path's 10-14.
template <int BN, int RK>
__device__ __forceinline__ void epilogue_rows_lean(const TcEpi& e, const float* __restrict__ s_bias, uint32_t tmem_lane_base, int nbase,
                                                   bool valid, int slot, int t, uint64_t* acc_full_bar, uint32_t parity) {
  constexpr int NCH = BN / 16;
  constexpr int PF = NCH < 4 ? NCH : 4;
  const long long res_off = (long long)slot * e.res_slot_stride + (long long)t * e.res_row_stride + nbase;
  const __half* resh = (RK == RES_F16 && valid) ? reinterpret_cast<const __half*>(e.res) + res_off : nullptr;
  const __half* res2p = (RK == RES_F16 && e.res2 && valid)
                            ? e.res2 + (long long)slot * e.res2_slot_stride + (long long)t * e.res2_row_stride + nbase : nullptr;
  __half* yh = (e.y && valid) ? reinterpret_cast<__half*>(e.y) + (long long)slot * e.y_slot_stride + (long long)(e.y_row0 + t) * e.y_row_stride + nbase
                              : nullptr;
  __half* y2p = (e.y2 && valid) ? e.y2 + (long long)slot * e.y2_slot_stride + (long long)(e.y2_row0 + t) * e.y2_row_stride + nbase : nullptr;
  const float inv = e.res_inv_slope != 0.f ? e.res_inv_slope : 1.f;
  const float s2 = e.act2 == ACT_NONE ? 1.f : e.slope2;
  const bool scaled = e.out_scale != 1.f;
  uint4 rbuf[RK == RES_F16 ? PF : 1][2], r2buf[RK == RES_F16 ? PF : 1][2];
  if (RK == RES_F16) {
#pragma unroll
    for (int c = 0; c < PF; ++c) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        rbuf[c][i] = resh ? *(reinterpret_cast<const uint4*>(resh + c * 16) + i) : make_uint4(0, 0, 0, 0);
        r2buf[c][i] = res2p ? *(reinterpret_cast<const uint4*>(res2p + c * 16) + i) : make_uint4(0, 0, 0, 0);
      }
    }
  }
  mbar_wait(acc_full_bar, parity);
  tc_fence_after();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int c0 = ch * 16;
    uint32_t acc[16];
    tc_ld_32x32b_x16(tmem_lane_base + (uint32_t)c0, acc);
    float v[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(s_bias + nbase + c0 + 4 * i);
      v[4 * i] = __uint_as_float(acc[4 * i]) + b4.x;
      v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + b4.y;
      v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + b4.z;
      v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + b4.w;
    }
    if (RK == RES_F16) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const __half2* hp = reinterpret_cast<const __half2*>(&rbuf[ch % PF][i]);
        const __half2* sp = reinterpret_cast<const __half2*>(&r2buf[ch % PF][i]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 a = __half22float2(hp[u]);
          v[8 * i + 2 * u] += fminf(a.x, a.x * inv);
          v[8 * i + 2 * u + 1] += fminf(a.y, a.y * inv);
          if (res2p) {                                          // warp-uniform
            const float2 c = __half22float2(sp[u]);
            v[8 * i + 2 * u] += c.x; v[8 * i + 2 * u + 1] += c.y;
          }
        }
      }
      if (ch + PF < NCH) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          rbuf[ch % PF][i] = resh ? *(reinterpret_cast<const uint4*>(resh + (ch + PF) * 16) + i) : make_uint4(0, 0, 0, 0);
          if (res2p) r2buf[ch % PF][i] = *(reinterpret_cast<const uint4*>(res2p + (ch + PF) * 16) + i);
        }
      }
    }
    if (scaled) {
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] *= e.out_scale;
    }
    if (yh) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        __half2 h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) h[u] = __floats2half2_rn(v[8 * i + 2 * u], v[8 * i + 2 * u + 1]);
        *(reinterpret_cast<uint4*>(yh + c0) + i) = *reinterpret_cast<uint4*>(h);
      }
    }
    if (y2p) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        __half2 h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float x0 = v[8 * i + 2 * u], x1 = v[8 * i + 2 * u + 1];
          h[u] = __floats2half2_rn(fmaxf(x0, x0 * s2), fmaxf(x1, x1 * s2));
        }
        *(reinterpret_cast<uint4*>(y2p + c0) + i) = *reinterpret_cast<uint4*>(h);
      }
    }
  }
}

// host-side mirror of the condition under which the lean epilogue applies
inline bool epilogue_is_lean(const TcEpi& e) {
  return e.scale == 1.f && e.acc_scale == 1.f && !e.rowmask && e.act == ACT_NONE && !e.accumulate && !e.y2_lo_off &&
         (!e.y || e.y_is_half) && (!e.res || e.res_is_half) && (e.act2 == ACT_NONE || (e.act2 == ACT_LRELU && e.slope2 > 0.f && e.slope2 < 1.f)) &&
         (e.res_inv_slope == 0.f || e.res_inv_slope > 1.f);
}

// LEAN_ONLY: the launcher has checked epilogue_is_lean(); the general path is not even compiled into that kernel variant
// (half the SASS: these kernels were stalling on instruction fetch)
template <int BN, bool LEAN_ONLY = false>
__device__ __forceinline__ void epilogue_rows(const TcEpi& e, const float* __restrict__ s_bias, uint32_t tmem_lane_base, int nbase,
                                              bool valid, int slot, int t, uint64_t* acc_full_bar, uint32_t parity) {
  if (LEAN_ONLY) {
    if (!e.res) epilogue_rows_lean<BN, RES_NONE>(e, s_bias, tmem_lane_base, nbase, valid, slot, t, acc_full_bar, parity);
    else epilogue_rows_lean<BN, RES_F16>(e, s_bias, tmem_lane_base, nbase, valid, slot, t, acc_full_bar, parity);
    return;
  }
  // vocoder layers: fp16 rows in and out, no scale / mask / activation before the residual, slopes in (0, 1)
  const bool lean = e.scale == 1.f && e.acc_scale == 1.f && !e.rowmask && e.act == ACT_NONE && !e.accumulate && !e.y2_lo_off &&
                    (!e.y || e.y_is_half) && (!e.res || e.res_is_half) && (e.act2 == ACT_NONE || (e.act2 == ACT_LRELU && e.slope2 > 0.f && e.slope2 < 1.f)) &&
                    (e.res_inv_slope == 0.f || e.res_inv_slope > 1.f);
  if (lean) {
    if (!e.res) epilogue_rows_lean<BN, RES_NONE>(e, s_bias, tmem_lane_base, nbase, valid, slot, t, acc_full_bar, parity);
    else epilogue_rows_lean<BN, RES_F16>(e, s_bias, tmem_lane_base, nbase, valid, slot, t, acc_full_bar, parity);
    return;
  }
  // one lean instantiation per residual kind (warp-uniform dispatch)
  if (!e.res) epilogue_rows_k<BN, RES_NONE>(e, s_bias, tmem_lane_base, nbase, valid, slot, t, acc_full_bar, parity);
  else if (e.res_is_half) epilogue_rows_k<BN, RES_F16>(e, s_bias, tmem_lane_base, nbase, valid, slot, t, acc_full_bar, parity);
  else epilogue_rows_k<BN, RES_F32>(e, s_bias, tmem_lane_base, nbase, valid, slot, t, acc_full_bar, parity);
}

// Epilogue of a SUMMED group (the last convs of the MRF branches of a vocoder scale, accumulated into one TMEM tile):
//   v = acc + sum_p bias_p + sum_p inverse-LeakyReLU(res_p rows)   (s_bias holds the summed bias; res_p = the fp16 context rows
//   branch p's conv c1 consumed, i.e. lrelu(x_p));  v *= out_scale;  y2 = lrelu(v) as fp16 rows of the next layer's context
// (hifigan_causal.py:324-329: the MRF average).  Residual chunks are double-buffered one chunk ahead.
struct SumRes { const __half* p[3]; float inv[3]; long long ss[3]; int rs[3]; };
template <int BN>
__device__ __forceinline__ void epilogue_rows_sum(const TcEpi& eo, const SumRes& sr, const float* __restrict__ s_bias, uint32_t tmem_lane_base,
                                                  int nbase, bool valid, int slot, int t, uint64_t* acc_full_bar, uint32_t parity) {
  constexpr int NCH = BN / 16;
  __half* y2p = (eo.y2 && valid) ? eo.y2 + (long long)slot * eo.y2_slot_stride + (long long)(eo.y2_row0 + t) * eo.y2_row_stride + nbase : nullptr;
  const float s2 = eo.act2 == ACT_NONE ? 1.f : eo.slope2;
  const __half* rp[3];
#pragma unroll
  for (int p = 0; p < 3; ++p)
    rp[p] = (sr.p[p] && valid) ? sr.p[p] + (long long)slot * sr.ss[p] + (long long)t * sr.rs[p] + nbase : nullptr;
  uint4 rb[2][3][2];
  auto fetch = [&](uint4 (&dst)[3][2], int c0) {
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int i = 0; i < 2; ++i) dst[p][i] = rp[p] ? *(reinterpret_cast<const uint4*>(rp[p] + c0) + i) : make_uint4(0, 0, 0, 0);
  };
  fetch(rb[0], 0);
  mbar_wait(acc_full_bar, parity);
  tc_fence_after();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int c0 = ch * 16;
    if (ch + 1 < NCH) fetch(rb[(ch + 1) & 1], c0 + 16);
    uint32_t acc[16];
    tc_ld_32x32b_x16(tmem_lane_base + (uint32_t)c0, acc);
    float v[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(s_bias + nbase + c0 + 4 * i);
      v[4 * i] = __uint_as_float(acc[4 * i]) + b4.x;
      v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + b4.y;
      v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + b4.z;
      v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + b4.w;
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const float inv = sr.inv[p];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const __half2* hp = reinterpret_cast<const __half2*>(&rb[ch & 1][p][i]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 a = __half22float2(hp[u]);
          v[8 * i + 2 * u] += fminf(a.x, a.x * inv);
          v[8 * i + 2 * u + 1] += fminf(a.y, a.y * inv);
        }
      }
    }
    if (y2p) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        __half2 h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float x0 = v[8 * i + 2 * u] * eo.out_scale, x1 = v[8 * i + 2 * u + 1] * eo.out_scale;
          h[u] = __floats2half2_rn(fmaxf(x0, x0 * s2), fmaxf(x1, x1 * s2));
        }
        *(reinterpret_cast<uint4*>(y2p + c0) + i) = *reinterpret_cast<uint4*>(h);
      }
    }
  }
}

// stage the bias (or zeros) of all `cout` output channels in shared memory, once per CTA
__device__ __forceinline__ void stage_bias(float* s_bias, const float* bias, int cout) {
  for (int i = threadIdx.x; i < cout; i += blockDim.x) s_bias[i] = bias ? bias[i] : 0.f;
}

// SP ("split, four-tile stages"): a stage holds A_hi, A_lo, W_hi, W_lo of one k-block and feeds three MMA groups
// (A_hi W_hi + A_hi W_lo + A_lo W_hi): 4 tiles of L2 -> SM traffic per product instead of the 6 of three plain k-blocks.
template <int BN, int BK, int STAGES, int MT, bool SP = false>
struct SmemLayout {
  static constexpr int A1_BYTES = TILE_M * BK * 2;      // one 128-row A tile
  static constexpr int A_BYTES = MT * A1_BYTES * (SP ? 2 : 1);
  static constexpr int B1_BYTES = BN * BK * 2;
  static constexpr int B_BYTES = B1_BYTES * (SP ? 2 : 1);
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BIAS_BYTES = 2048 * 4;           // bias of up to 2048 output channels
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers etc.*/ + BIAS_BYTES;
};

// MT = m-tiles per CTA tile.  The vocoder layers run at the L2 -> SM bandwidth limit (every k-block brings 16 KB of A and
// 16 KB of B for four 128x128x16 MMAs), so MT = 2 lets two 128-row A tiles share each B tile: 48 KB per eight MMAs instead
// of 64 KB.  Each m-tile of the pair has its own double-buffered TMEM accumulator (4 x 128 columns = all of TMEM, hence
// one CTA per SM) and its own epilogue warpgroup.
// ES = epilogue warpgroups per m-tile, each draining BN / ES accumulator columns (the short-K layers are bound by the
// epilogue's latency, not by the MMAs: more warps in flight per tile).
template <int BN, int BK, int STAGES, int MT, int ES, bool LEAN, bool SP = false>
__global__ void __launch_bounds__(64 + 128 * MT * ES, (MT == 2 || SP) ? 1 : (BN >= 128 ? 2 : (BN >= 64 ? 3 : 4)))
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, TcArgs a) {
  static_assert(!SP || MT == 1, "four-tile split stages take one m-tile per CTA");
  using SL = SmemLayout<BN, BK, STAGES, MT, SP>;
  constexpr int SWZ = BK * 2;                 // bytes per tile row = swizzle span (128 or 64)
  constexpr int TMEM_COLS = 2 * MT * BN < 32 ? 32 : 2 * MT * BN;     // two accumulators per m-tile: MMAs of tile i+1 overlap the epilogue of tile i
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * SL::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * SL::STAGE_BYTES + 256);
  stage_bias(s_bias, a.e.bias, a.cout);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: provably warp-uniform role branches
  // a tile is NS = 128/TT consecutive streams x TT consecutive time steps: one rectangular TMA box.
  // Persistent CTAs walk the tiles round-robin; consecutive tile ids share the A rows (nt fastest).  With MT = 2 a CTA tile is
  // the pair of m-tiles (2q, 2q+1); the second one may lie past the end (odd count): it is computed on stale operands and dropped.
  const int NS = TILE_M / a.TT, TPS = a.L / a.TT;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128 * MT * ES); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();
  pdl_wait();                                  // activations of the previous layer are complete and visible from here on

  if (warp == 0) {
    // ===================================================================== TMA producer (warp-uniform loop, one elected lane issues)
    {
      int s = 0;                                             // smem ring position / phase carried across tiles (the ring never drains)
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int nt = tile % a.n_tiles, mt = (tile / a.n_tiles) * MT;
        int stream0m[MT], t0m[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) { stream0m[m] = ((mt + m) / TPS) * NS; t0m[m] = ((mt + m) % TPS) * a.TT; }
        // k-blocks in order: segment (split operands) -> tap -> channel block; counters instead of divisions (this loop paces the TMA issue)
        int seg = 0, j = 0, c0 = 0;
        for (int kb = 0; kb < a.kblocks; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * SL::STAGE_BYTES;
          uint8_t* sb = sa + SL::A_BYTES;
          if (SP) {
            if (elect_one_sync()) {                          // A_hi, A_lo, W_hi, W_lo of this (tap, channel block)
              mbar_expect_tx(&full_bar[s], SL::A_BYTES + SL::B_BYTES);
              tma_load_3d(sa, &tmA, &full_bar[s], c0, a.row0 + t0m[0] + j * a.dil, stream0m[0]);
              tma_load_3d(sa + SL::A1_BYTES, &tmA, &full_bar[s], c0, a.row0 + t0m[0] + j * a.dil, stream0m[0] + a.lo_slot_off);
              tma_load_2d(sb, &tmW, &full_bar[s], kb * BK, nt * BN);
              tma_load_2d(sb + SL::B1_BYTES, &tmW, &full_bar[s], a.k * a.cin + kb * BK, nt * BN);
            }
          } else if (elect_one_sync()) {
            const bool second = MT == 2 && mt + 1 < a.m_tiles;
            mbar_expect_tx(&full_bar[s], SL::B_BYTES + (second ? 2 : 1) * SL::A1_BYTES);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              if (m == 1 && !second) break;
              tma_load_3d(sa + m * SL::A1_BYTES, &tmA, &full_bar[s], c0, a.row0 + t0m[m] + j * a.dil,
                          stream0m[m] + (seg == 2 ? a.lo_slot_off : 0));   // box {BK, TT, NS}
            }
            tma_load_2d(sb, &tmW, &full_bar[s], kb * BK, nt * BN);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
          c0 += BK;
          if (c0 == a.cin) { c0 = 0; if (++j == a.k) { j = 0; ++seg; } }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp-uniform loop, one elected lane issues)
    {
      constexpr uint32_t idesc = make_idesc<BN>();
      int it = 0, s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&acc_empty[ab], aph ^ 1);                  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(ab * MT * BN);
        for (int kb = 0; kb < a.kblocks; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * SL::STAGE_BYTES);
          const uint32_t sb = sa + SL::A_BYTES;
          const uint64_t bdesc = make_smem_desc<SWZ>(sb);
          if (SP) {
            // three products per stage: A_hi W_hi, A_hi W_lo, A_lo W_hi (the dropped A_lo W_lo term is below fp32 resolution)
#pragma unroll
            for (int g = 0; g < 3; ++g) {
              const uint64_t ad = make_smem_desc<SWZ>(sa + (g == 2 ? SL::A1_BYTES : 0));
              const uint64_t bd = make_smem_desc<SWZ>(sb + (g == 1 ? SL::B1_BYTES : 0));
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk)
                if (elect_one_sync())
                  tc_mma_f16(tacc, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, (kb | g | kk) != 0 ? 1u : 0u);
            }
          } else {
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const uint64_t adesc = make_smem_desc<SWZ>(sa + m * SL::A1_BYTES);
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              // advance 16 halfs = 32 bytes along K inside the swizzle atom: +2 in the (>>4) address field
              if (elect_one_sync())
                tc_mma_f16(tacc + (uint32_t)(m * BN), adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (kb | kk) != 0 ? 1u : 0u);
            }
          }
          }
          if (elect_one_sync()) tc_commit(&empty_bar[s]);          // frees the smem stage when these MMAs have read it
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one_sync()) tc_commit(&acc_full[ab]);            // accumulators complete
      }
    }
  } else {
    // ===================================================================== epilogue (one warpgroup per m-tile of the CTA tile)
    const int wgi = (warp - 2) >> 2;         // warpgroup index: m-tile of the pair (slow) x column part (fast)
    const int wg = wgi / ES, cpart = wgi - wg * ES;
    constexpr int BNE = BN / ES;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;       // tile row == TMEM lane
    const int q = r / a.TT, tt = r - q * a.TT;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int nt = tile % a.n_tiles, mt = (tile / a.n_tiles) * MT + wg;
      const int stream = (mt / TPS) * NS + q, t = (mt % TPS) * a.TT + tt;
      epilogue_rows<BNE, LEAN>(a.e, s_bias, tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((ab * MT + wg) * BN + cpart * BNE),
                         nt * BN + cpart * BNE, mt < a.m_tiles && stream < a.n_streams, stream, t, &acc_full[ab], aph);
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}


// ==============================================================================================
// CTA-pair variant of the ring kernel (tcgen05.mma.cta_group::2, M = 256): the vocoder layers with 128 / 256 output
// channels ran at the L2 -> SM operand limit with 128 x 128 tiles (16 KB of A + 16 KB of B per four MMAs = 128 B per MMA
// cycle against ~64 B/cycle of TMA delivery per SM).  A pair of CTAs on the two SMs of a TPC shares one 256 x BN tile: each
// CTA fetches its own 128 A rows and HALF of the B rows per k-block, the leader issues M = 256 MMAs that read both halves, and
// each CTA drains the accumulator rows that live in its own TMEM.  Per CTA and k-block: 16 KB + BN/2 x 128 B for 2 x BN MMA
// cycles -- 64 B per cycle at BN = 256.
//   full[s]   (leader's copy): leader's producer arrives with the byte count of BOTH CTAs; both CTAs' TMA loads complete on it
//   empty[s]  (each CTA's own): the leader's commit arrives on both copies (multicast) when the MMAs have read the stage
//   acc_full  (each CTA's own): multicast commit after a tile's last MMA;  acc_empty (leader's): every epilogue thread of both CTAs
// ==============================================================================================
template <int BN, int STAGES>
struct Smem2Layout {
  static constexpr int A_BYTES = TILE_M * 64 * 2;
  static constexpr int B_BYTES = (BN / 2) * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BIAS_BYTES = 2048 * 4;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256 + BIAS_BYTES;
};

// Up to kMaxGroup INDEPENDENT convolutions of the same shape (the three MRF branches of a vocoder scale: kernel sizes 3 / 7 / 11 on
// the same rows) run as one launch: the tile space is the concatenation of the problems' tile spaces, longest K first, so a
// launch's prologue, cluster barriers, first TMA round trip and last-tile epilogue are paid once per group instead of once per conv
// and the short-K problem's tiles fill the tail of the long one's.
constexpr int kMaxGroup = 3;
struct TcProb { int k, dil, row0, kblocks; TcEpi e; };
struct TcGroup {
  CUtensorMap tmA[kMaxGroup], tmW[kMaxGroup];
  TcProb prob[kMaxGroup];
  int n_prob;
  TcArgs a;                  // shared shape: n_streams, L, TT, cin, cout, n_tiles, m_tiles, num_tiles (per problem)
};

// SUM: the problems are not independent tiles but terms of ONE output (same rows, same output channels): every tile runs the k-blocks
// of all problems into the same accumulator and one epilogue adds the summed bias and every problem's residual.
template <int BN, int STAGES, int ES, int OCC, bool SUM = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * ES, OCC)
conv_gemm_tc2_kernel(const __grid_constant__ TcGroup g) {
  const TcArgs& a = g.a;
  using SL = Smem2Layout<BN, STAGES>;
  constexpr int BK = 64, SWZ = 128;
  constexpr int TMEM_COLS = 2 * BN;                    // two accumulators: MMAs of tile i+1 overlap the epilogue of tile i
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * SL::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * SL::STAGE_BYTES + 256);
  if (SUM) {
    for (int i = threadIdx.x; i < a.cout; i += blockDim.x) {
      float b = 0.f;
      for (int pi = 0; pi < g.n_prob; ++pi) b += g.prob[pi].e.bias ? g.prob[pi].e.bias[i] : 0.f;
      s_bias[i] = b;
    }
  } else {
    for (int pi = 0; pi < g.n_prob; ++pi) stage_bias(s_bias + pi * a.cout, g.prob[pi].e.bias, a.cout);
  }
  const int total_tiles = SUM ? a.num_tiles : g.n_prob * a.num_tiles;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int NS = TILE_M / a.TT, TPS = a.L / a.TT;

  if (warp == 0 && lane == 0) {
    for (int pi = 0; pi < g.n_prob; ++pi) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmA[pi]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmW[pi]) : "memory");
    }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * 128 * ES); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                          // both CTAs' barriers are initialised before any remote arrival / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs: own A rows, own half of B)
    int s = 0;
    uint32_t ph = 0;
    for (int gt = cluster_id; gt < total_tiles; gt += n_clusters) {
      const int p_lo = SUM ? 0 : gt / a.num_tiles, p_hi = SUM ? g.n_prob : p_lo + 1;
      const int tile = SUM ? gt : gt - p_lo * a.num_tiles;
      const int nt = tile % a.n_tiles;
      int mt = (tile / a.n_tiles) * 2 + rank;
      if (mt >= a.m_tiles) mt = a.m_tiles - 1;             // odd tile count: the peer computes a duplicate that its epilogue drops
      const int stream0 = (mt / TPS) * NS, t0 = (mt % TPS) * a.TT;
      for (int pi = p_lo; pi < p_hi; ++pi) {
        const TcProb& pr = g.prob[pi];
        int j = 0, c0 = 0;
        for (int kb = 0; kb < pr.kblocks; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * SL::STAGE_BYTES;
          uint8_t* sb = sa + SL::A_BYTES;
          if (elect_one_sync()) {
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * SL::STAGE_BYTES);
            tma_load_3d_2sm(sa, &g.tmA[pi], &full_bar[s], c0, pr.row0 + t0 + j * pr.dil, stream0);     // box {64, TT, NS}
            tma_load_2d_2sm(sb, &g.tmW[pi], &full_bar[s], kb * BK, nt * BN + rank * (BN / 2));           // box {64, BN / 2}
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
          c0 += BK;
          if (c0 == a.cin) { c0 = 0; ++j; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc<BN, 256>();
      int it = 0, s = 0;
      uint32_t ph = 0;
      for (int gt = cluster_id; gt < total_tiles; gt += n_clusters, ++it) {
        int kblocks = 0;
        if (SUM) { for (int pi = 0; pi < g.n_prob; ++pi) kblocks += g.prob[pi].kblocks; }
        else kblocks = g.prob[gt / a.num_tiles].kblocks;
        const int ab = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&acc_empty[ab], aph ^ 1);                  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(ab * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * SL::STAGE_BYTES);
          const uint64_t adesc = make_smem_desc<SWZ>(sa), bdesc = make_smem_desc<SWZ>(sa + SL::A_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            if (elect_one_sync())
              tc_mma_f16_2sm(tacc, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (kb | kk) != 0 ? 1u : 0u);
          if (elect_one_sync()) tc_commit_2sm(&empty_bar[s]);          // frees the stage in both CTAs
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one_sync()) tc_commit_2sm(&acc_full[ab]);            // accumulator rows complete in both CTAs
      }
    }
  } else {
    // ===================================================================== epilogue (both CTAs, own 128 rows; ES column parts)
    const int cpart = (warp - 2) >> 2;
    constexpr int BNE = BN / ES;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int q = r / a.TT, tt = r - q * a.TT;
    int it = 0;
    SumRes sr;
    if (SUM) {
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        const bool on = p < g.n_prob && g.prob[p].e.res;
        sr.p[p] = on ? reinterpret_cast<const __half*>(g.prob[p].e.res) : nullptr;
        sr.inv[p] = on && g.prob[p].e.res_inv_slope != 0.f ? g.prob[p].e.res_inv_slope : 1.f;
        sr.ss[p] = on ? g.prob[p].e.res_slot_stride : 0; sr.rs[p] = on ? g.prob[p].e.res_row_stride : 0;
      }
    }
    for (int gt = cluster_id; gt < total_tiles; gt += n_clusters, ++it) {
      const int pi = SUM ? g.n_prob - 1 : gt / a.num_tiles, tile = SUM ? gt : gt - pi * a.num_tiles;
      const int ab = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int nt = tile % a.n_tiles, mt = (tile / a.n_tiles) * 2 + rank;
      const int stream = (mt / TPS) * NS + q, t = (mt % TPS) * a.TT + tt;
      if (SUM)
        epilogue_rows_sum<BNE>(g.prob[pi].e, sr, s_bias, tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * BN + cpart * BNE),
                               nt * BN + cpart * BNE, mt < a.m_tiles && stream < a.n_streams, stream, t, &acc_full[ab], aph);
      else
      epilogue_rows<BNE, false>(g.prob[pi].e, s_bias + pi * a.cout, tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * BN + cpart * BNE),
                                nt * BN + cpart * BNE, mt < a.m_tiles && stream < a.n_streams, stream, t, &acc_full[ab], aph);
      tc_fence_before();
      mbar_arrive_leader(&acc_empty[ab]);
    }
  }
  tc_fence_before();
  cluster_sync_all();                          // no CTA may free TMEM / exit while its partner still uses the pair's resources
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ==============================================================================================
// "Window" variant for the long narrow layers (L % 128 == 0, cout = BN in {32, 64}: vocoder scales
// 2 and 3 and the last upsampling conv), which are HBM-bound, not tensor-bound:
//   * persistent CTAs, static round-robin over the 128-row tiles of (stream, time);
//   * the whole packed weight matrix (k taps x [BN, C]) is loaded into shared memory ONCE per CTA;
//   * per tile ONE TMA box brings the input window of 128 + (k-1)*dil rows; every tap's A operand is
//     the same window at a row offset (the UMMA descriptor start address moves by j*dil rows), so the
//     input is read from L2/HBM once instead of k times;
//   * two TMEM accumulators: the MMAs of tile i+1 run while the epilogue warps drain tile i; with one
//     CTA per SM two epilogue warpgroups alternate tiles so enough loads/stores are in flight.
// ==============================================================================================
struct WinArgs {
  int L, k, dil, row0, win_rows, num_tiles, tiles_per_stream;
  TcEpi e;
};

// A descriptor whose start address is offset by whole rows inside a swizzle atom (tap j starts
// j*dil rows into the window).  Measured on B200: the swizzle XOR is applied on absolute shared-memory
// address bits, so the descriptor's base_offset field must stay 0 for such starts (setting it to
// (start >> 7) & 7 produces wrong operands) -- tests/test_gpu_parity.py covers odd offsets in both
// the 128-byte and the 64-byte swizzle.

template <int C, int BN, int NBUF, int NEPI>
__global__ void __launch_bounds__(64 + 128 * NEPI, NEPI == 1 ? (C == 32 ? 4 : 2) : 1)
conv_window_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, WinArgs a) {
  constexpr int ROWB = C * 2;                       // bytes per row = swizzle span (64 or 128)
  constexpr int TAPB = BN * ROWB;                   // one tap of the weight matrix
  constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  const int winb = (a.win_rows * ROWB + 1023) & ~1023;
  uint8_t* sW = smem;
  uint8_t* sA = smem + ((a.k * TAPB + 1023) & ~1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + NBUF * winb);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = a_full + NBUF;
  uint64_t* acc_full = a_empty + NBUF;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);
  stage_bias(s_bias, a.e.bias, BN);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: provably warp-uniform role branches
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    mbar_init(w_full, 1);
    for (int s = 0; s < NBUF; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================================================== TMA producer (warp-uniform loop, one elected lane issues)
    {
      if (elect_one_sync()) {
        mbar_expect_tx(w_full, (uint32_t)(a.k * TAPB));                   // weights are constants: fetched before the dependency wait
        for (int j = 0; j < a.k; ++j) tma_load_2d(sW + j * TAPB, &tmW, w_full, j * C, 0);
      }
      pdl_wait();
      int it = 0;
      for (int g = blockIdx.x; g < a.num_tiles; g += gridDim.x, ++it) {
        const int buf = it % NBUF;
        const uint32_t ph = (it / NBUF) & 1;
        mbar_wait(&a_empty[buf], ph ^ 1);
        const int si = g / a.tiles_per_stream, t0 = (g - si * a.tiles_per_stream) * TILE_M;
        if (elect_one_sync()) {
          mbar_expect_tx(&a_full[buf], (uint32_t)(a.win_rows * ROWB));
          tma_load_3d(sA + buf * winb, &tmA, &a_full[buf], 0, a.row0 + t0, si);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp-uniform loop, one elected lane issues)
    {
      constexpr uint32_t idesc = make_idesc<BN>();
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint32_t sW32 = smem_u32(sW), sA32 = smem_u32(sA);
      int it = 0;
      for (int g = blockIdx.x; g < a.num_tiles; g += gridDim.x, ++it) {
        const int buf = it % NBUF, ab = it & 1;
        const uint32_t ph = (it / NBUF) & 1, aph = (it >> 1) & 1;
        mbar_wait(&acc_empty[ab], aph ^ 1);
        mbar_wait(&a_full[buf], ph);
        tc_fence_after();
        const uint32_t win = sA32 + buf * winb;
        for (int j = 0; j < a.k; ++j) {
          const uint32_t arow = win + (uint32_t)(j * a.dil * ROWB);
          const uint32_t brow = sW32 + (uint32_t)(j * TAPB);
#pragma unroll
          for (int kk = 0; kk < C / 16; ++kk)
            if (elect_one_sync())
              tc_mma_f16(tmem_base + (uint32_t)(ab * BN), make_smem_desc<ROWB>(arow + kk * 32),
                         make_smem_desc<ROWB>(brow + kk * 32), idesc, (j | kk) != 0 ? 1u : 0u);
        }
        if (elect_one_sync()) { tc_commit(&a_empty[buf]); tc_commit(&acc_full[ab]); }
      }
    }
  } else {
    // ===================================================================== epilogue warpgroups
    pdl_wait();
    const int wg = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    int it = 0;
    for (int g = blockIdx.x; g < a.num_tiles; g += gridDim.x, ++it) {
      if (NEPI == 2 && (it & 1) != wg) continue;
      const int ab = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int si = g / a.tiles_per_stream, t0 = (g - si * a.tiles_per_stream) * TILE_M;
      epilogue_rows<BN>(a.e, s_bias, tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * BN), 0, true, si, t0 + r,
                        &acc_full[ab], aph);
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct MapKey {
  const void* ptr; unsigned long long d0, d1, d2, s1, s2; unsigned b0, b1, b2, swz;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 && b0 == o.b0 && b1 == o.b1 &&
           b2 == o.b2 && swz == o.swz;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    for (unsigned long long v : {k.d0, k.d1, k.d2, k.s1, k.s2, (unsigned long long)k.b0, (unsigned long long)k.b1,
                                 (unsigned long long)k.b2, (unsigned long long)k.swz})
      h = h * 1000003ull ^ (size_t)v;
    return h;
  }
};

}  // namespace

int get_tensor_map(CUtensorMap* out, const void* ptr, int rank, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                   unsigned long long s1_bytes, unsigned long long s2_bytes, unsigned b0, unsigned b1, unsigned b2, int swz_bytes) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  MapKey key{ptr, d0, d1, d2, s1_bytes, s2_bytes, b0, b1, b2, (unsigned)swz_bytes};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return 0; }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)); return 1; }
  cache.emplace(key, m);
  *out = m;
  return 0;
}

// CTAs of a kernel that fit on one SM: registers, shared memory (with the carve-out preference set to
// "max shared", which the launchers request), warps and the 512 TMEM columns.  The CUDA occupancy query is
// not used: it answers for the *default* carve-out and returns 1 for these kernels.
int resident_ctas(const void* func, int threads, size_t dyn_smem, int tmem_cols) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, func) != cudaSuccess) { (void)cudaGetLastError(); return 1; }
  const int regs_per_cta = ((fa.numRegs * 32 + 255) / 256 * 256) * (threads / 32);      // allocation granularity: 256 regs per warp
  int n = 65536 / std::max(regs_per_cta, 1);
  n = std::min<int>(n, (int)((227 * 1024) / (dyn_smem + fa.sharedSizeBytes + 1024)));
  n = std::min(n, 64 / (threads / 32));
  n = std::min(n, 512 / std::max(tmem_cols, 32));
  return std::max(n, 1);
}

int num_sms() {
  static DeviceOnce once;
  int n = 148;
  device_once(once, &n, [](int* v) {
    int dev = 0;
    cudaGetDevice(&dev);
    return cudaDeviceGetAttribute(v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess ? 0 : 1;
  });
  return n;
}

namespace {

// Launch with programmatic stream serialisation (PDL) when CONAN_TC_PDL=1.  Measured on B200 (S = 1024): no gain
// (7.89 ms with, 7.83 ms without) -- the persistent grids fill every SM until their last tile, so the next kernel's
// prologue has nowhere to overlap; kept off by default.
template <typename K, typename A>
int launch_pdl(K kern, int grid, int threads, size_t smem, cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmW, const A& a) {
  static const bool pdl = [] { const char* v = getenv("CONAN_TC_PDL"); return v && atoi(v) != 0; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmW, a);
  if (e != cudaSuccess) { set_error(std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(e)); return 1; }
  return 0;
}

int pick_tt(int L) {
  for (int tt = 128; tt >= 1; tt >>= 1)
    if (L % tt == 0) return tt;
  return 0;
}
int pick_bn(int cout) {
  if (cout % 128 == 0) return 128;
  if (cout % 64 == 0) return 64;
  if (cout % 32 == 0) return 32;
  return 0;
}

template <int BN, int BK, int STAGES, int MT = 1, int ES = 1, bool LEAN = false, bool SP = false>
int launch_variant(const CUtensorMap& tmA, const CUtensorMap& tmW, TcArgs a, long long m_tiles, cudaStream_t st) {
  using SL = SmemLayout<BN, BK, STAGES, MT, SP>;
  auto kern = conv_gemm_tc_kernel<BN, BK, STAGES, MT, ES, LEAN, SP>;
  if (SP) a.kblocks /= 3;                      // one stage per (tap, channel block): the three products share it
  constexpr int threads = 64 + 128 * MT * ES;
  static DeviceOnce once;
  int per_sm = 1;
  if (device_once(once, &per_sm, [&](int* v) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
        *v = resident_ctas((const void*)kern, threads, SL::TOTAL, 2 * MT * BN);
        return 0;
      }))
    return 1;
  a.m_tiles = (int)m_tiles;
  a.num_tiles = (int)(((m_tiles + MT - 1) / MT) * a.n_tiles);
  const int grid = std::min(a.num_tiles, num_sms() * per_sm);           // persistent: exactly the co-resident CTAs
  if (getenv("CONAN_TC_VERBOSE")) fprintf(stderr, "ring<%d,%d,%d,%d,%d,sp%d> tiles %d per_sm %d grid %d\n", BN, BK, STAGES, MT, ES, (int)SP, a.num_tiles, per_sm, grid);
  if (launch_pdl(kern, grid, threads, SL::TOTAL, st, tmA, tmW, a)) return 1;
  CONAN_CHECK_LAUNCH();
  return 0;
}

template <int BN, int STAGES, int ES, int OCC = 1, bool SUM = false>
int launch_pair_variant(const TcGroup& g0, long long m_tiles, cudaStream_t st) {
  using SL = Smem2Layout<BN, STAGES>;
  static_assert(OCC * 2 * BN <= 512, "TMEM columns of the co-resident CTAs");
  auto kern = conv_gemm_tc2_kernel<BN, STAGES, ES, OCC, SUM>;
  constexpr int threads = 64 + 128 * ES;
  static DeviceOnce once;
  if (device_once(once, nullptr, [&](int*) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
        return 0;
      }))
    return 1;
  TcGroup g = g0;
  g.a.m_tiles = (int)m_tiles;
  g.a.num_tiles = (int)(((m_tiles + 1) / 2) * g.a.n_tiles);              // tiles of 256 rows x BN columns, per problem
  const int clusters = std::min((SUM ? 1 : g.n_prob) * g.a.num_tiles, (num_sms() / 2) * OCC);     // persistent: OCC CTA pairs per TPC
  if (getenv("CONAN_TC_VERBOSE")) fprintf(stderr, "pair<%d,%d,%d,%d> problems %d tiles %d clusters %d\n", BN, STAGES, ES, OCC, g.n_prob, g.a.num_tiles, clusters);
  kern<<<2 * clusters, threads, SL::TOTAL, st>>>(g);                     // cluster dims (2, 1, 1) are compiled into the kernel
  CONAN_CHECK_LAUNCH();
  return 0;
}

int pair_mode();
int launch_pair_group(const TcGroup& g, int bn2, long long m_tiles, cudaStream_t st, bool sum = false) {
  if (sum) {
    if (bn2 == 256) return launch_pair_variant<256, 5, 2, 1, true>(g, m_tiles, st);
    return launch_pair_variant<128, 3, 2, 2, true>(g, m_tiles, st);
  }
  if (bn2 == 256) return pair_mode() == 4 ? launch_pair_variant<256, 5, 4>(g, m_tiles, st) : launch_pair_variant<256, 5, 2>(g, m_tiles, st);
  // 128 output channels: two pairs per TPC (3 stages each), so one pair's epilogue runs under the other's MMAs
  return pair_mode() == 3 ? launch_pair_variant<128, 6, 2>(g, m_tiles, st) : launch_pair_variant<128, 3, 2, 2>(g, m_tiles, st);
}

int pair_mode() {
  // CONAN_TC_2CTA: 0 = never, 1 = layers with cout % 256 == 0 only, 2 (default) = also cout % 128 == 0
  static int v = [] { const char* e = getenv("CONAN_TC_2CTA"); return e ? atoi(e) : 2; }();
  return v;
}

int es_min_tiles() {
  static int v = [] { const char* e = getenv("CONAN_TC_ES_MIN"); return e ? atoi(e) : 2; }();
  return v;
}

int window_mode() {
  // CONAN_TC_WINDOW=0 routes every layer through the ring kernel (A/B comparisons, debugging)
  static int mode = [] { const char* v = getenv("CONAN_TC_WINDOW"); return v ? atoi(v) : 1; }();
  return mode;
}

constexpr int WIN_NBUF = 3;
size_t window_smem_bytes(const conan_conv_params_t& p) {
  const int rowb = p.cin * 2, tapb = p.cout * rowb;
  const int win_rows = TILE_M + (p.k - 1) * p.dil;
  const size_t winb = ((size_t)win_rows * rowb + 1023) & ~(size_t)1023;
  return (((size_t)p.k * tapb + 1023) & ~(size_t)1023) + WIN_NBUF * winb + 1024 + 512;      // + alignment slack, barriers, bias
}

bool window_eligible(const conan_conv_params_t& p) {
  if (window_mode() == 0 || p.x_split) return false;
  if (p.L % TILE_M != 0) return false;
  if (!((p.cin == 32 && p.cout == 32) || (p.cin == 64 && p.cout == 64))) return false;
  if (TILE_M + (p.k - 1) * p.dil > 256) return false;
  return window_smem_bytes(p) <= 200 * 1024;
}

// Persistent launch: the grid is exactly the number of CTAs that can be co-resident (queried from the
// occupancy calculator for this kernel / block size / dynamic smem), so the static round-robin tile
// schedule never leaves a second, partial wave.
template <int C, int BN, int NEPI>
int launch_window_variant(const CUtensorMap& tmA, const CUtensorMap& tmW, const WinArgs& a, size_t smem, cudaStream_t st) {
  auto kern = conv_window_tc_kernel<C, BN, WIN_NBUF, NEPI>;
  static DeviceOnce once;
  if (device_once(once, nullptr, [&](int*) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
        return 0;
      }))
    return 1;
  const int per_sm = resident_ctas((const void*)kern, 64 + 128 * NEPI, smem, 2 * BN);
  const int grid = std::min(a.num_tiles, num_sms() * per_sm);
  if (getenv("CONAN_TC_VERBOSE")) fprintf(stderr, "window<%d,%d,%d> k %d tiles %d smem %zu per_sm %d grid %d\n", C, BN, NEPI, a.k, a.num_tiles, smem, per_sm, grid);
  if (launch_pdl(kern, grid, 64 + 128 * NEPI, smem, st, tmA, tmW, a)) return 1;
  CONAN_CHECK_LAUNCH();
  return 0;
}

int launch_conv_window_tc(const conan_conv_params_t& p, cudaStream_t st) {
  const int C = p.cin, Ktot = p.k * p.cin;
  const int win_rows = TILE_M + (p.k - 1) * p.dil;
  CUtensorMap tmA, tmW;
  if (get_tensor_map(&tmA, p.x, 3, (unsigned long long)p.cin, (unsigned long long)p.x_rows, (unsigned long long)p.n_slots,
                     (unsigned long long)p.x_row_stride * 2, (unsigned long long)p.x_slot_stride * 2, C, win_rows, 1, C * 2))
    return 1;
  if (get_tensor_map(&tmW, p.w, 2, (unsigned long long)Ktot, (unsigned long long)p.cout, 1, (unsigned long long)Ktot * 2, 0, C, p.cout, 1, C * 2))
    return 1;
  WinArgs a;
  a.L = p.L; a.k = p.k; a.dil = p.dil; a.row0 = p.row0; a.win_rows = win_rows;
  a.tiles_per_stream = p.L / TILE_M; a.num_tiles = p.n_streams * a.tiles_per_stream;
  a.e = TcEpi{p.bias, p.scale, p.act, p.slope, p.res, p.res_slot_stride, p.res_row_stride, p.rowmask, p.mask_slot_stride,
              p.out_scale, p.y, p.y_slot_stride, p.y_row_stride, p.y_row0, p.accumulate, (__half*)p.y2, p.y2_slot_stride,
              p.y2_row_stride, p.y2_row0, p.act2, p.slope2, p.acc_scale == 0.f ? 1.f : p.acc_scale, p.y2_split ? p.y2_lo_off : 0,
              p.res_is_half, p.res_inv_slope, (const __half*)p.res2, p.res2_slot_stride, p.res2_row_stride, p.y_is_half};
  const size_t smem = window_smem_bytes(p);
  // one epilogue warpgroup per CTA when several CTAs fit on an SM, two when the resident weights leave room for one
  if ((227 * 1024) / (smem + 1024) >= 2) {
    if (C == 32) return launch_window_variant<32, 32, 1>(tmA, tmW, a, smem, st);
    return launch_window_variant<64, 64, 1>(tmA, tmW, a, smem, st);
  }
  if (C == 32) return launch_window_variant<32, 32, 2>(tmA, tmW, a, smem, st);
  return launch_window_variant<64, 64, 2>(tmA, tmW, a, smem, st);
}

}  // namespace

bool conv_gemm_tc_uses_window(const conan_conv_params_t& p) { return conv_gemm_tc_eligible(p) && window_eligible(p); }

bool conv_gemm_tc_eligible(const conan_conv_params_t& p) {
  if (!p.x_is_half) return false;
  if (p.slot_ids) return false;                   // compact operands only: a tile's input must be one rectangular TMA box
  if (p.y2 && !p.y2_is_half) return false;
  if (p.cin % 32 != 0 || p.x_row_stride != p.cin) return false;
  if (pick_bn(p.cout) == 0 || pick_tt(p.L) == 0) return false;
  if (p.row0 < 0) return false;
  if (p.act == ACT_TANH || p.act2 > ACT_LRELU) return false;          // none / relu / leaky / gelu (first), none / relu / leaky (second)
  if (p.y2_split && (!p.y2 || p.y2_lo_off % 8)) return false;
  if (p.y && (p.y_slot_stride % 8 || p.y_row_stride % 8)) return false;
  if (p.res2 && (!p.res || !p.res_is_half || !p.res2_is_half || p.res2_slot_stride % 8 || p.res2_row_stride % 8)) return false;
  if (p.y_is_half && p.accumulate) return false;
  if (p.res && (p.res_slot_stride % 8 || p.res_row_stride % 8 || ((uintptr_t)p.res) % 16)) return false;
  if (p.y2 && (p.y2_slot_stride % 8 || p.y2_row_stride % 8)) return false;
  if (((uintptr_t)p.x) % 128 || ((uintptr_t)p.w) % 128) return false;
  return true;
}

int launch_conv_gemm_tc(const conan_conv_params_t& p, cudaStream_t st) {
  if (!conv_gemm_tc_eligible(p)) { set_error("conv_gemm_tc: not eligible"); return 1; }
  if (p.n_streams <= 0) return 0;
  if (window_eligible(p)) return launch_conv_window_tc(p, st);
  const int BK = (p.cin % 64 == 0) ? 64 : 32;
  const int BN = pick_bn(p.cout);
  const int TT = pick_tt(p.L);
  const int nseg = p.x_split ? 3 : 1;
  const int Ktot = nseg * p.k * p.cin;
  CUtensorMap tmA, tmW;
  if (get_tensor_map(&tmA, p.x, 3, (unsigned long long)p.cin, (unsigned long long)p.x_rows,
                     (unsigned long long)(p.x_split ? p.x_lo_slot_off + p.n_slots : p.n_slots),
                     (unsigned long long)p.x_row_stride * 2, (unsigned long long)p.x_slot_stride * 2, BK, TT, TILE_M / TT, BK * 2))
    return 1;
  if (get_tensor_map(&tmW, p.w, 2, (unsigned long long)Ktot, (unsigned long long)p.cout, 1, (unsigned long long)Ktot * 2, 0, BK, BN, 1,
                     BK * 2))
    return 1;
  TcArgs a;
  a.n_streams = p.n_streams; a.L = p.L; a.TT = TT; a.cin = p.cin; a.k = p.k; a.dil = p.dil; a.cout = p.cout; a.row0 = p.row0;
  a.kblocks = Ktot / BK; a.n_tiles = p.cout / BN; a.nseg = nseg; a.lo_slot_off = (int)p.x_lo_slot_off;
  a.e = TcEpi{p.bias, p.scale, p.act, p.slope, p.res, p.res_slot_stride, p.res_row_stride, p.rowmask, p.mask_slot_stride,
              p.out_scale, p.y, p.y_slot_stride, p.y_row_stride, p.y_row0, p.accumulate, (__half*)p.y2, p.y2_slot_stride,
              p.y2_row_stride, p.y2_row0, p.act2, p.slope2, p.acc_scale == 0.f ? 1.f : p.acc_scale, p.y2_split ? p.y2_lo_off : 0,
              p.res_is_half, p.res_inv_slope, (const __half*)p.res2, p.res2_slot_stride, p.res2_row_stride, p.y_is_half};
  const int NS = TILE_M / TT;
  const long long m_tiles = (long long)((p.n_streams + NS - 1) / NS) * (p.L / TT);
  // Few CTAs and a long K loop (the Emformer / Conan GEMMs: M = 4..6 rows x streams): one CTA per SM anyway, so
  // spend the shared memory on pipeline depth instead of co-residency.
  const bool deep = m_tiles * a.n_tiles <= 2 * num_sms() && a.kblocks >= 12;
  // split operands: four-tile stages (CONAN_TC_SPLIT4=0 falls back to three plain k-block segments)
  static const int split4 = [] { const char* v = getenv("CONAN_TC_SPLIT4"); return v ? atoi(v) : 1; }();
  if (nseg == 3 && split4) {
    static const int narrow4 = [] { const char* v = getenv("CONAN_TC_NARROW"); return v ? atoi(v) : 1; }();
    int bn = BN;
    if (bn == 128 && deep && narrow4 && m_tiles * a.n_tiles * 2 <= num_sms())
      bn = (m_tiles * (p.cout / 64) * 4 >= 3LL * num_sms() || p.cout % 32 != 0) ? 64 : 32;
    if (bn != BN) {
      a.n_tiles = p.cout / bn;
      if (get_tensor_map(&tmW, p.w, 2, (unsigned long long)Ktot, (unsigned long long)p.cout, 1, (unsigned long long)Ktot * 2, 0, BK, bn, 1, BK * 2))
        return 1;
    }
    if (BK == 64) {
      if (bn == 128) return launch_variant<128, 64, 3, 1, 1, false, true>(tmA, tmW, a, m_tiles, st);
      if (bn == 64) return launch_variant<64, 64, 4, 1, 1, false, true>(tmA, tmW, a, m_tiles, st);
      return launch_variant<32, 64, 4, 1, 1, false, true>(tmA, tmW, a, m_tiles, st);
    }
    if (bn == 128) return launch_variant<128, 32, 4, 1, 1, false, true>(tmA, tmW, a, m_tiles, st);
    if (bn == 64) return launch_variant<64, 32, 4, 1, 1, false, true>(tmA, tmW, a, m_tiles, st);
    return launch_variant<32, 32, 4, 1, 1, false, true>(tmA, tmW, a, m_tiles, st);
  }
  // CTA pairs (M = 256 MMAs) for the wide fp16 layers with enough tiles to fill the machine
  if (nseg == 1 && BK == 64 && pair_mode() && p.cout % 128 == 0) {
    const int bn2 = (p.cout % 256 == 0) ? 256 : (pair_mode() >= 2 ? 128 : 0);
    // at least one 256-row tile per CTA pair (CONAN_TC_2CTA_MIN overrides the tile count from which pairs are used)
    static const long long min_tiles = [] { const char* e = getenv("CONAN_TC_2CTA_MIN"); return e ? atoll(e) : -1LL; }();
    const long long pair_tiles = bn2 ? ((m_tiles + 1) / 2) * (p.cout / bn2) : 0;
    if (bn2 && pair_tiles >= (min_tiles >= 0 ? min_tiles : (long long)num_sms() / 2)) {
      a.n_tiles = p.cout / bn2;
      if (get_tensor_map(&tmW, p.w, 2, (unsigned long long)Ktot, (unsigned long long)p.cout, 1, (unsigned long long)Ktot * 2, 0, BK, bn2 / 2, 1, BK * 2))
        return 1;
      TcGroup g;
      g.n_prob = 1; g.tmA[0] = tmA; g.tmW[0] = tmW; g.a = a;
      g.prob[0] = TcProb{p.k, p.dil, p.row0, a.kblocks, a.e};
      return launch_pair_group(g, bn2, m_tiles, st);
    }
  }
  if (BK == 64) {
    if (BN == 128) {
      // too few 128-wide tiles to fill the SMs: narrow the tile (more CTAs, each with the same K loop) until ~3/4 of them have one
      static const int narrow = [] { const char* v = getenv("CONAN_TC_NARROW"); return v ? atoi(v) : 1; }();
      if (deep && narrow && m_tiles * a.n_tiles * 2 <= num_sms()) {
        const int bn = (m_tiles * (p.cout / 64) * 4 >= 3LL * num_sms() || p.cout % 32 != 0) ? 64 : 32;
        a.n_tiles = p.cout / bn;
        if (get_tensor_map(&tmW, p.w, 2, (unsigned long long)Ktot, (unsigned long long)p.cout, 1, (unsigned long long)Ktot * 2, 0, BK, bn, 1, BK * 2))
          return 1;
        return bn == 64 ? launch_variant<64, 64, 8>(tmA, tmW, a, m_tiles, st) : launch_variant<32, 64, 8>(tmA, tmW, a, m_tiles, st);
      }
      if (deep) return launch_variant<128, 64, 6>(tmA, tmW, a, m_tiles, st);
      // enough tiles to fill the machine with pairs: two m-tiles share every B tile (the layer is L2 -> SM bandwidth bound)
      static const int pair_mode = [] { const char* v = getenv("CONAN_TC_PAIR"); return v ? atoi(v) : 2; }();
      if (nseg == 1 && m_tiles * a.n_tiles >= (long long)es_min_tiles() * num_sms()) {
        if (pair_mode == 1) return launch_variant<128, 64, 4, 2, 1>(tmA, tmW, a, m_tiles, st);
        static const int lean_only = [] { const char* v = getenv("CONAN_TC_LEANONLY"); return v ? atoi(v) : 0; }();   // measured: 1.99 vs 1.91 ms with the shared variant
        if (pair_mode == 2) return (lean_only && epilogue_is_lean(a.e)) ? launch_variant<128, 64, 3, 1, 2, true>(tmA, tmW, a, m_tiles, st)
                                                         : launch_variant<128, 64, 3, 1, 2>(tmA, tmW, a, m_tiles, st);
        if (pair_mode == 3) return launch_variant<128, 64, 4, 2, 2>(tmA, tmW, a, m_tiles, st);
      }
      return launch_variant<128, 64, 3>(tmA, tmW, a, m_tiles, st);
    }
    if (BN == 64) return deep ? launch_variant<64, 64, 8>(tmA, tmW, a, m_tiles, st) : launch_variant<64, 64, 2>(tmA, tmW, a, m_tiles, st);
    return deep ? launch_variant<32, 64, 8>(tmA, tmW, a, m_tiles, st) : launch_variant<32, 64, 2>(tmA, tmW, a, m_tiles, st);
  }
  if (BN == 128) return launch_variant<128, 32, 3>(tmA, tmW, a, m_tiles, st);
  if (BN == 64) return launch_variant<64, 32, 3>(tmA, tmW, a, m_tiles, st);
  return launch_variant<32, 32, 3>(tmA, tmW, a, m_tiles, st);
}

// Up to three independent convs of one shape as ONE launch of the CTA-pair kernel.  Returns -1 (nothing launched) when the group
// does not qualify: the caller then launches the convs one by one.
int launch_conv_gemm_tc_group(const conan_conv_params_t* ps, int G, cudaStream_t st, bool sum) {
  if (G < 2 || G > kMaxGroup || !pair_mode()) return -1;
  const conan_conv_params_t& p0 = ps[0];
  if (sum) {
    // terms of one output: fp16 context residuals only, no per-term outputs; ps[G - 1] carries the output (y2, out_scale, slope2)
    for (int i = 0; i < G; ++i) {
      const conan_conv_params_t& p = ps[i];
      if (p.y || p.res2 || p.rowmask || p.accumulate || p.act != ACT_NONE || p.scale != 1.f || (p.res && !p.res_is_half) ||
          (p.acc_scale != 0.f && p.acc_scale != 1.f) || (i < G - 1 && p.y2))
        return -1;
    }
    const conan_conv_params_t& pl = ps[G - 1];
    if (!pl.y2 || !pl.y2_is_half || pl.y2_split || !(pl.act2 == ACT_NONE || (pl.act2 == ACT_LRELU && pl.slope2 > 0.f && pl.slope2 < 1.f))) return -1;
  }
  for (int i = 0; i < G; ++i) {
    const conan_conv_params_t& p = ps[i];
    if (!conv_gemm_tc_eligible(p) || window_eligible(p) || p.x_split || p.cin % 64 != 0 || p.cout % 128 != 0) return -1;
    if (p.cin != p0.cin || p.cout != p0.cout || p.L != p0.L || p.n_streams != p0.n_streams || p.n_slots != p0.n_slots) return -1;
  }
  if (p0.n_streams <= 0) return 0;
  if (G * p0.cout > 2048) return -1;                                      // bias staging area
  const int BK = 64, TT = pick_tt(p0.L), NS = TILE_M / TT;
  const int bn2 = (p0.cout % 256 == 0) ? 256 : (pair_mode() >= 2 ? 128 : 0);
  if (!bn2) return -1;
  const long long m_tiles = (long long)((p0.n_streams + NS - 1) / NS) * (p0.L / TT);
  static const long long min_tiles = [] { const char* e = getenv("CONAN_TC_2CTA_MIN"); return e ? atoll(e) : -1LL; }();
  // (a summed group takes this kernel at every size: the individual launches round the running sum to fp16 between the terms, so
  // switching by stream count would make a stream's bits depend on how many others are ready)
  if (!sum && ((m_tiles + 1) / 2) * (p0.cout / bn2) < (min_tiles >= 0 ? min_tiles : (long long)num_sms() / 2)) return -1;
  // longest K first: the short problem's tiles fill the tail of the long one's
  int order[kMaxGroup] = {0, 1, 2};
  if (!sum) std::sort(order, order + G, [&](int x, int y) { return ps[x].k > ps[y].k; });
  TcGroup g;
  g.n_prob = G;
  for (int i = 0; i < G; ++i) {
    const conan_conv_params_t& p = ps[order[i]];
    const int Ktot = p.k * p.cin;
    if (get_tensor_map(&g.tmA[i], p.x, 3, (unsigned long long)p.cin, (unsigned long long)p.x_rows, (unsigned long long)p.n_slots,
                       (unsigned long long)p.x_row_stride * 2, (unsigned long long)p.x_slot_stride * 2, BK, TT, TILE_M / TT, BK * 2))
      return 1;
    if (get_tensor_map(&g.tmW[i], p.w, 2, (unsigned long long)Ktot, (unsigned long long)p.cout, 1, (unsigned long long)Ktot * 2, 0, BK, bn2 / 2, 1, BK * 2))
      return 1;
    TcEpi e{p.bias, p.scale, p.act, p.slope, p.res, p.res_slot_stride, p.res_row_stride, p.rowmask, p.mask_slot_stride,
            p.out_scale, p.y, p.y_slot_stride, p.y_row_stride, p.y_row0, p.accumulate, (__half*)p.y2, p.y2_slot_stride,
            p.y2_row_stride, p.y2_row0, p.act2, p.slope2, p.acc_scale == 0.f ? 1.f : p.acc_scale, p.y2_split ? p.y2_lo_off : 0,
            p.res_is_half, p.res_inv_slope, (const __half*)p.res2, p.res2_slot_stride, p.res2_row_stride, p.y_is_half};
    g.prob[i] = TcProb{p.k, p.dil, p.row0, Ktot / BK, e};
  }
  TcArgs& a = g.a;
  a.n_streams = p0.n_streams; a.L = p0.L; a.TT = TT; a.cin = p0.cin; a.k = 0; a.dil = 0; a.cout = p0.cout; a.row0 = 0;
  a.kblocks = 0; a.n_tiles = p0.cout / bn2; a.nseg = 1; a.lo_slot_off = 0; a.e = g.prob[0].e;
  return launch_pair_group(g, bn2, m_tiles, st, sum);
}

}  // namespace conan
