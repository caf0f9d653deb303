// Two-GEMM residual block of the Conan chunk path as ONE tcgen05 kernel with fp32-grade split-fp16 operands:
//
//   h = act(scale * (conv_k(x) + b1))        GEMM1: K1 = k taps x C1 channels  ->  hidden (chunks of 128)
//   y = W2 . h                               GEMM2: hidden -> N2 = 256, accumulated in TMEM across the hidden chunks of the CTA
//
// covering (a) the decoder's CausalResidualBlock body: causal conv k5 256 -> 512, x 5^-1/2, erf-GELU, 1x1 512 -> 256
// (modules/commons/conv.py:127-178) and (b) the aligner layer's feed-forward 256 -> 2048 ReLU -> 256
// (modules/Conan/prosody_util.py:108-127).  As separate launches the hidden tensor makes a round trip through HBM as a split
// fp16 pair and the second GEMM is one more deep, launch-bound kernel; here it stays in shared memory in exactly the swizzled
// K-major layout GEMM2 reads (same construction as ffn_fused.cu, which does this for the Emformer's 96-wide rows).
//
// A CTA owns a 128-row tile of (stream, time) rows and a slice of the hidden dimension.  GEMM1's operands are streamed: one ring
// slot (32 KB) per (tap, 32-channel block) holds A_hi, A_lo (TMA boxes {32 ch, TT rows, 128/TT streams} of the compact context
// buffer at row offset = tap) and W1_hi, W1_lo, and feeds the three MMA groups A_hi W_hi + A_hi W_lo + A_lo W_hi.  GEMM2's weight
// tiles (256 x 64, hi and lo) go through the same ring.  acc1 is double-buffered in TMEM (2 x 128 columns) next to acc2 (256
// columns).  Each CTA writes its partial y * 2^-10 to P[fs][row][256]; the LayerNorm that follows sums the partials with b2 and
// the residual (deterministic, no atomics).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace conan {

namespace {

constexpr int BF_THREADS = 384;        // warp 0: TMA producer, warp 1: MMA issuer (+TMEM alloc), warps 4..11: epilogue
constexpr int BF_EPI = 256;
constexpr int BF_N2 = 256;
constexpr int BF_CH = 128;             // hidden chunk
constexpr int BF_SLOTS = 4;
constexpr int BF_SLOT_BYTES = 32768;   // GEMM1: A_hi | A_lo | W1_hi | W1_lo (8 KB each, 32-wide k-block); GEMM2: one 256 x 64 W2 tile
constexpr int BF_T8 = 8192;
constexpr int BF_H_TILE = TILE_M * 128;                        // 128 rows x 64 halfs (128-byte swizzle)
constexpr int BF_OFF_W = 0;
constexpr int BF_OFF_H = BF_SLOTS * BF_SLOT_BYTES;             // 131072
constexpr int BF_OFF_BAR = BF_OFF_H + 4 * BF_H_TILE;           // 196608
constexpr int BF_SMEM = BF_OFF_BAR + 1024 + 1024;

struct BfArgs {
  int n_streams, L, TT, C1, k, row0, lo_slot_off;
  int hidden, n_chunks, FS;
  int K1;                       // k * C1
  const float* b1;
  float scale1; int act;
  float* P; long long M;        // [FS][M][256]
  float acc_scale;
  // cluster reduction (CL = true): out[row][c] = (sum over the 4 hidden slices + b2[c] + res[row][c]) * mask[row]
  float* out; const float* res; const float* b2; const float* mask; int ld;
  // optional LayerNorm of the finished rows, fused into the cluster epilogue (the reducing warp holds whole rows):
  //   wmask[row] = (sum_c |x| > 0);  y = ((x * pre - mean) * rstd * g + b) * post  ->  ln_out (GEMM operand rows) [, ln_out2 fp32]
  const float* ln_g; const float* ln_b; float ln_eps;
  RowView ln_out; float* ln_out2; int ln_out2_ld;
  const float* ln_pre; const float* ln_post; float* ln_wmask; const float* ln_add;
};

// CL: the 4 CTAs that hold the hidden slices of one row tile form a thread-block cluster; each parks its partial y tile in its own
// shared memory (the operand ring is dead by then), and CTA q of the cluster sums rows [32 q, 32 q + 32) of the four tiles through
// distributed shared memory, in slice order (deterministic), adds bias and residual, applies the row mask and writes the finished rows:
// no partial tensors in HBM and no partial-summing pass in the LayerNorm that follows.
constexpr int BF_YLD = 260;            // floats per parked row (256 + 4: conflict-free float4 rows)

template <bool CL>
__global__ void __launch_bounds__(BF_THREADS, 1)
block_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                   const __grid_constant__ CUtensorMap tmW2, BfArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BF_OFF_BAR);
  uint64_t* acc1_full = bars;              // [2]
  uint64_t* acc1_empty = bars + 2;         // [2]
  uint64_t* h_full = bars + 4;             // [1]
  uint64_t* h_empty = bars + 5;            // [1]
  uint64_t* acc2_full = bars + 6;          // [1]
  uint64_t* w_full = bars + 8;             // [SLOTS]
  uint64_t* w_empty = w_full + BF_SLOTS;   // [SLOTS]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + BF_SLOTS);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int mt = blockIdx.x / a.FS, fs = blockIdx.x - mt * a.FS;
  const int c_begin = (a.n_chunks * fs) / a.FS, c_end = (a.n_chunks * (fs + 1)) / a.FS, nc = c_end - c_begin;
  const int NS = TILE_M / a.TT, TPS = a.L / a.TT;
  const int stream0 = (mt / TPS) * NS, t0 = (mt % TPS) * a.TT;
  const int kb1 = a.K1 / 32;                // GEMM1 k-blocks per chunk

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    mbar_init(h_full, BF_EPI); mbar_init(h_empty, 1); mbar_init(acc2_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc1_full[s], 1); mbar_init(&acc1_empty[s], BF_EPI); }
    for (int s = 0; s < BF_SLOTS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer, in the order the MMA warp consumes:
    // G1(0), then per chunk [G1(c+1)], G2(c)
    int s = 0;
    uint32_t ph = 1;
    auto g1 = [&](int c) {
      const int f0 = (c_begin + c) * BF_CH;
      int j = 0, c0 = 0;
      for (int kb = 0; kb < kb1; ++kb) {
        mbar_wait_warp(&w_empty[s], ph);
        if (elect_one_sync()) {
          uint8_t* d = smem + BF_OFF_W + s * BF_SLOT_BYTES;
          mbar_expect_tx(&w_full[s], 4 * BF_T8);
          tma_load_3d(d, &tmX, &w_full[s], c0, a.row0 + t0 + j, stream0);                          // A_hi: box {32, TT, NS}
          tma_load_3d(d + BF_T8, &tmX, &w_full[s], c0, a.row0 + t0 + j, stream0 + a.lo_slot_off);  // A_lo
          tma_load_2d(d + 2 * BF_T8, &tmW1, &w_full[s], kb * 32, f0);                              // W1_hi: box {32, 128}
          tma_load_2d(d + 3 * BF_T8, &tmW1, &w_full[s], a.K1 + kb * 32, f0);                       // W1_lo
        }
        if (++s == BF_SLOTS) { s = 0; ph ^= 1; }
        c0 += 32;
        if (c0 == a.C1) { c0 = 0; ++j; }
      }
    };
    auto g2 = [&](int c) {
      const int f0 = (c_begin + c) * BF_CH;
      for (int t = 0; t < 4; ++t) {                          // (kb2, hi | lo): W2 tiles of 256 output rows x 64 hidden columns
        const int kb2 = t >> 1, lo = t & 1;
        mbar_wait_warp(&w_empty[s], ph);
        if (elect_one_sync()) {
          mbar_expect_tx(&w_full[s], BF_SLOT_BYTES);
          tma_load_2d(smem + BF_OFF_W + s * BF_SLOT_BYTES, &tmW2, &w_full[s], lo * a.hidden + f0 + kb2 * 64, 0);
        }
        if (++s == BF_SLOTS) { s = 0; ph ^= 1; }
      }
    };
    g1(0);
    for (int c = 0; c < nc; ++c) {
      if (c + 1 < nc) g1(c + 1);
      g2(c);
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp-uniform, one elected lane issues)
    constexpr uint32_t idesc1 = make_idesc<BF_CH>();
    constexpr uint32_t idesc2 = make_idesc<BF_N2>();
    const uint32_t s32 = smem_u32(smem);
    int s = 0;
    uint32_t ph = 0;
    auto g1 = [&](int c) {
      const int b = c & 1;
      mbar_wait_warp(&acc1_empty[b], ((c >> 1) & 1) ^ 1);           // epilogue 1 of chunk c-2 has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(b * BF_CH);
      for (int kb = 0; kb < kb1; ++kb) {
        mbar_wait_warp(&w_full[s], ph);
        tc_fence_after();
        const uint32_t d = s32 + BF_OFF_W + s * BF_SLOT_BYTES;
        const uint64_t a_hi = make_smem_desc<64>(d), a_lo = make_smem_desc<64>(d + BF_T8);
        const uint64_t w_hi = make_smem_desc<64>(d + 2 * BF_T8), w_lo = make_smem_desc<64>(d + 3 * BF_T8);
        tc_mma_f16_tap<2>(tacc, a_hi, w_hi, idesc1, kb == 0 ? 1u : 0u);
        tc_mma_f16_tap<2>(tacc, a_hi, w_lo, idesc1, 0u);
        tc_mma_f16_tap<2>(tacc, a_lo, w_hi, idesc1, 0u);
        if (elect_one_sync()) tc_commit(&w_empty[s]);
        if (++s == BF_SLOTS) { s = 0; ph ^= 1; }
      }
      if (elect_one_sync()) tc_commit(&acc1_full[b]);
    };
    auto g2 = [&](int c) {
      mbar_wait_warp(h_full, c & 1);                                 // epilogue 1 of chunk c has written h (hi, lo)
      tc_fence_after();
      const uint32_t tacc = tmem_base + 256u;
#pragma unroll
      for (int kb2 = 0; kb2 < 2; ++kb2) {
        const uint64_t h_hi = make_smem_desc<128>(s32 + BF_OFF_H + (0 * 2 + kb2) * BF_H_TILE);
        const uint64_t h_lo = make_smem_desc<128>(s32 + BF_OFF_H + (1 * 2 + kb2) * BF_H_TILE);
        // slot t = 2 kb2: W2_hi, slot 2 kb2 + 1: W2_lo.  Products: h_hi W_hi, h_lo W_hi (same slot), then h_hi W_lo.
        mbar_wait_warp(&w_full[s], ph);
        tc_fence_after();
        const uint64_t w_hi = make_smem_desc<128>(s32 + BF_OFF_W + s * BF_SLOT_BYTES);
        tc_mma_f16_tap<4>(tacc, h_hi, w_hi, idesc2, (c == 0 && kb2 == 0) ? 1u : 0u);
        tc_mma_f16_tap<4>(tacc, h_lo, w_hi, idesc2, 0u);
        if (elect_one_sync()) tc_commit(&w_empty[s]);
        if (++s == BF_SLOTS) { s = 0; ph ^= 1; }
        mbar_wait_warp(&w_full[s], ph);
        tc_fence_after();
        const uint64_t w_lo = make_smem_desc<128>(s32 + BF_OFF_W + s * BF_SLOT_BYTES);
        tc_mma_f16_tap<4>(tacc, h_hi, w_lo, idesc2, 0u);
        if (elect_one_sync()) tc_commit(&w_empty[s]);
        if (++s == BF_SLOTS) { s = 0; ph ^= 1; }
      }
      if (elect_one_sync()) tc_commit(h_empty);                      // h may be overwritten once these MMAs have read it
    };
    g1(0);
    for (int c = 0; c < nc; ++c) {
      if (c + 1 < nc) g1(c + 1);
      g2(c);
    }
    if (elect_one_sync()) tc_commit(acc2_full);
  } else if (warp >= 4) {
    // ===================================================================== epilogue warps: wg 0 / 1 = hidden columns 0-63 / 64-127
    const int wg = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const bool gelu = a.act == ACT_GELU;
    for (int c = 0; c < nc; ++c) {
      const int b = c & 1;
      const float* b1 = a.b1 + (long long)(c_begin + c) * BF_CH + wg * 64;
      mbar_wait_lane0(&acc1_full[b], (c >> 1) & 1, 0);
      mbar_wait_lane0(h_empty, (c & 1) ^ 1, 0);                      // GEMM2 of chunk c-1 has finished reading h
      tc_fence_after();
      uint8_t* hhi = smem + BF_OFF_H + (0 * 2 + wg) * BF_H_TILE + r * 128;
      uint8_t* hlo = smem + BF_OFF_H + (1 * 2 + wg) * BF_H_TILE + r * 128;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t acc[16];
        tc_ld_32x32b_x16(tmem_base + lane_base + (uint32_t)(b * BF_CH + wg * 64 + ch * 16), acc);
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          __half2 hi[4], lo[4];
          const float4 ba = *reinterpret_cast<const float4*>(b1 + ch * 16 + h8 * 8);
          const float4 bb = *reinterpret_cast<const float4*>(b1 + ch * 16 + h8 * 8 + 4);
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float x = fmaf(__uint_as_float(acc[h8 * 8 + u]), a.acc_scale, bv[u]) * a.scale1;
            v[u] = gelu ? 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)) : fmaxf(x, 0.f);     // warp-uniform branch
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            hi[u] = __floats2half2_rn(v[2 * u], v[2 * u + 1]);
            const float2 hf = __half22float2(hi[u]);
            lo[u] = __floats2half2_rn(v[2 * u] - hf.x, v[2 * u + 1] - hf.y);
          }
          const uint32_t chunk = (uint32_t)((ch * 2 + h8) ^ (r & 7)) << 4;      // 128-byte swizzle on a 1024-aligned tile
          *reinterpret_cast<uint4*>(hhi + chunk) = *reinterpret_cast<uint4*>(hi);
          *reinterpret_cast<uint4*>(hlo + chunk) = *reinterpret_cast<uint4*>(lo);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(h_full);
      mbar_arrive(&acc1_empty[b]);
    }
    // ---- final: partial y rows of this hidden slice (wg 0 / 1 = output columns 0-127 / 128-255)
    mbar_wait_lane0(acc2_full, 0, 0);
    tc_fence_after();
    const int stream = stream0 + r / a.TT, t = t0 + r % a.TT;
    const bool valid = stream < a.n_streams;
    float* prow = a.P + ((long long)fs * a.M + (long long)stream * a.L + t) * BF_N2 + wg * 128;
    float* yrow = reinterpret_cast<float*>(smem) + r * BF_YLD + wg * 128;      // CL: parked tile [128][BF_YLD] over the dead ring / h area
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      uint32_t acc[16];
      tc_ld_32x32b_x16(tmem_base + lane_base + (uint32_t)(256 + wg * 128 + ch * 16), acc);
      if (CL) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<float4*>(yrow + ch * 16 + 4 * u) =
              make_float4(__uint_as_float(acc[4 * u]) * a.acc_scale, __uint_as_float(acc[4 * u + 1]) * a.acc_scale,
                          __uint_as_float(acc[4 * u + 2]) * a.acc_scale, __uint_as_float(acc[4 * u + 3]) * a.acc_scale);
      } else if (valid) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<float4*>(prow + ch * 16 + 4 * u) =
              make_float4(__uint_as_float(acc[4 * u]) * a.acc_scale, __uint_as_float(acc[4 * u + 1]) * a.acc_scale,
                          __uint_as_float(acc[4 * u + 2]) * a.acc_scale, __uint_as_float(acc[4 * u + 3]) * a.acc_scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
  if (CL) {
    cluster_sync_all();                          // every slice's tile is parked (release / acquire over the cluster)
    if (warp >= 4) {
      // CTA `fs` of the cluster finishes rows [32 fs, 32 fs + 32): warp w -> 4 rows, lane -> 8 columns
      const int w = warp - 4;
      const float* base = reinterpret_cast<const float*>(smem);
      uint32_t src[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) src[q] = dsmem_addr(base, (uint32_t)q);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = fs * 32 + w * 4 + i;
        const int stream = stream0 + r / a.TT, t = t0 + r % a.TT;
        if (stream >= a.n_streams) continue;                        // warp-uniform
        const long long row = (long long)stream * a.L + t;
        const uint32_t off = (uint32_t)((r * BF_YLD + lane * 8) * 4);
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {                               // slice order: the sum does not depend on which CTA finishes when
          const float4 u0 = dsmem_ld_f4(src[q] + off), u1 = dsmem_ld_f4(src[q] + off + 16);
          s0.x += u0.x; s0.y += u0.y; s0.z += u0.z; s0.w += u0.w;
          s1.x += u1.x; s1.y += u1.y; s1.z += u1.z; s1.w += u1.w;
        }
        const float m = a.mask ? a.mask[row] : 1.f;
        const float4 b0 = *reinterpret_cast<const float4*>(a.b2 + lane * 8), b1 = *reinterpret_cast<const float4*>(a.b2 + lane * 8 + 4);
        const float4 r0 = *reinterpret_cast<const float4*>(a.res + row * a.ld + lane * 8);
        const float4 r1 = *reinterpret_cast<const float4*>(a.res + row * a.ld + lane * 8 + 4);
        float* o = a.out + row * a.ld + lane * 8;
        float v[8] = {(s0.x + b0.x + r0.x) * m, (s0.y + b0.y + r0.y) * m, (s0.z + b0.z + r0.z) * m, (s0.w + b0.w + r0.w) * m,
                      (s1.x + b1.x + r1.x) * m, (s1.y + b1.y + r1.y) * m, (s1.z + b1.z + r1.z) * m, (s1.w + b1.w + r1.w) * m};
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        if (a.ln_g) {                                               // warp-uniform: the LayerNorm that consumes these rows
          if (a.ln_wmask) {
            float sabs = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) sabs += fabsf(v[u]);
            sabs = warp_sum(sabs);
            if (lane == 0) a.ln_wmask[row] = sabs > 0.f ? 1.f : 0.f;
          }
          const float pm = a.ln_pre ? a.ln_pre[row] : 1.f;
          float sm1 = 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) { v[u] *= pm; sm1 += v[u]; }
          const float mean = warp_sum(sm1) * (1.f / BF_N2);
          float sq = 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) { const float d = v[u] - mean; sq += d * d; }
          const float rstd = 1.f / sqrtf(warp_sum(sq) * (1.f / BF_N2) + a.ln_eps);
          const float post = a.ln_post ? a.ln_post[row] : 1.f;
          const long long oo = (long long)stream * a.ln_out.slot_stride + (long long)(a.ln_out.row0 + t) * a.ln_out.row_stride + lane * 8;
          float y[8];
          const float4 g0 = *reinterpret_cast<const float4*>(a.ln_g + lane * 8), g1 = *reinterpret_cast<const float4*>(a.ln_g + lane * 8 + 4);
          const float4 e0 = *reinterpret_cast<const float4*>(a.ln_b + lane * 8), e1 = *reinterpret_cast<const float4*>(a.ln_b + lane * 8 + 4);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
          for (int u = 0; u < 8; ++u) y[u] = ((v[u] - mean) * rstd * gg[u] + bb[u]) * post;
          if (a.ln_add) {
            const float4 a0 = *reinterpret_cast<const float4*>(a.ln_add + row * a.ld + lane * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(a.ln_add + row * a.ld + lane * 8 + 4);
            y[0] += a0.x; y[1] += a0.y; y[2] += a0.z; y[3] += a0.w; y[4] += a1.x; y[5] += a1.y; y[6] += a1.z; y[7] += a1.w;
          }
          if (a.ln_out.is_half == 2 && (oo & 7) == 0 && (a.ln_out.lo_off & 7) == 0) {
            // split-fp16 operand rows: eight values = one 16-byte store per plane
            __half2 hi[4], lo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const __half h0 = __float2half_rn(y[2 * u]), h1 = __float2half_rn(y[2 * u + 1]);
              hi[u] = __halves2half2(h0, h1);
              lo[u] = __halves2half2(__float2half_rn(y[2 * u] - __half2float(h0)), __float2half_rn(y[2 * u + 1] - __half2float(h1)));
            }
            __half* ob = reinterpret_cast<__half*>(a.ln_out.base) + oo;
            *reinterpret_cast<uint4*>(ob) = *reinterpret_cast<const uint4*>(hi);
            *reinterpret_cast<uint4*>(ob + a.ln_out.lo_off) = *reinterpret_cast<const uint4*>(lo);
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) store_view(a.ln_out, oo + u, y[u]);
          }
          if (a.ln_out2) {
            float* o2 = a.ln_out2 + row * a.ln_out2_ld + lane * 8;
            *reinterpret_cast<float4*>(o2) = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4*>(o2 + 4) = make_float4(y[4], y[5], y[6], y[7]);
          }
        }
      }
    }
    cluster_sync_all();                          // no CTA may exit while a partner still reads its tile
  }
}

int pick_tt_bf(int L) {
  for (int tt = 128; tt >= 1; tt >>= 1)
    if (L % tt == 0) return tt;
  return 0;
}

}  // namespace

bool block_fused_eligible(int C1, int k, int hidden, int N2, int L) {
  return C1 % 32 == 0 && k >= 1 && hidden % BF_CH == 0 && hidden >= BF_CH && N2 == BF_N2 && pick_tt_bf(L) > 0;
}

int block_fused_split(int n_streams, int L, int hidden, long long max_partial_rows) {
  // The hidden-dimension split is a CONSTANT (4 slices, or one per chunk when there are fewer): which partial sums exist and the
  // order in which the LayerNorm adds them must not depend on how many streams happen to be ready, or a stream's mel would differ
  // in the last bits between a step it shares with 1023 others and one it runs alone (tests pin that equality bit for bit).
  // At 1024 streams 32 row tiles x 4 slices = 128 CTAs: one wave.
  const int fs = std::min(4, hidden / BF_CH);
  return ((long long)fs * n_streams * L <= max_partial_rows) ? fs : 1;
}

int launch_block_fused(const BlockFusedParams& p, cudaStream_t st) {
  if (!block_fused_eligible(p.C1, p.k, p.hidden, p.N2, p.L)) { set_error("block_fused: shape not eligible"); return 1; }
  if (p.n_streams <= 0) return 0;
  const int TT = pick_tt_bf(p.L), NS = TILE_M / TT;
  const int mt = ((p.n_streams + NS - 1) / NS) * (p.L / TT);
  if (p.FS < 1 || p.FS > p.hidden / BF_CH) { set_error("block_fused: bad hidden split"); return 1; }
  const int K1 = p.k * p.C1;
  CUtensorMap tmX, tmW1, tmW2;
  if (get_tensor_map(&tmX, p.x, 3, (unsigned long long)p.C1, (unsigned long long)p.x_rows, (unsigned long long)(p.lo_slot_off + p.n_slots),
                     (unsigned long long)p.C1 * 2, (unsigned long long)p.x_slot_stride * 2, 32, TT, NS, 64))
    return 1;
  if (get_tensor_map(&tmW1, p.w1, 2, (unsigned long long)3 * K1, (unsigned long long)p.hidden, 1, (unsigned long long)3 * K1 * 2, 0, 32, BF_CH, 1, 64)) return 1;
  if (get_tensor_map(&tmW2, p.w2, 2, (unsigned long long)3 * p.hidden, BF_N2, 1, (unsigned long long)3 * p.hidden * 2, 0, 64, BF_N2, 1, 128)) return 1;
  BfArgs a;
  a.n_streams = p.n_streams; a.L = p.L; a.TT = TT; a.C1 = p.C1; a.k = p.k; a.row0 = p.row0; a.lo_slot_off = (int)p.lo_slot_off;
  a.hidden = p.hidden; a.n_chunks = p.hidden / BF_CH; a.FS = p.FS; a.K1 = K1; a.b1 = p.b1; a.scale1 = p.scale1; a.act = p.act;
  a.P = p.partials; a.M = (long long)p.n_streams * p.L; a.acc_scale = p.acc_scale;
  a.out = p.out; a.res = p.res; a.b2 = p.b2; a.mask = p.mask; a.ld = p.ld;
  a.ln_g = p.ln_g; a.ln_b = p.ln_b; a.ln_eps = p.ln_eps; a.ln_out = p.ln_out; a.ln_out2 = p.ln_out2; a.ln_out2_ld = p.ln_out2_ld;
  a.ln_pre = p.ln_pre; a.ln_post = p.ln_post; a.ln_wmask = p.ln_wmask; a.ln_add = p.ln_add;
  if (p.ln_g && !p.out) { set_error("block_fused: the fused LayerNorm needs the cluster reduction"); return 1; }
  const bool cl = p.out != nullptr;
  if (cl && (p.FS != 4 || !p.res || !p.b2 || p.ld % 4 != 0)) { set_error("block_fused: the cluster reduction needs 4 hidden slices, a residual and a bias"); return 1; }
  static_assert(TILE_M * BF_YLD * 4 <= BF_OFF_BAR, "parked tile must fit below the barriers");
  auto setup = [&](auto kern, DeviceOnce& once) -> int {
    return device_once(once, nullptr, [&](int*) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
      return 0;
    });
  };
  static DeviceOnce once_plain, once_cl;
  if (!cl) {
    if (setup(block_fused_kernel<false>, once_plain)) return 1;
    block_fused_kernel<false><<<mt * p.FS, BF_THREADS, BF_SMEM, st>>>(tmX, tmW1, tmW2, a);
  } else {
    if (setup(block_fused_kernel<true>, once_cl)) return 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(mt * p.FS)); cfg.blockDim = dim3(BF_THREADS); cfg.dynamicSmemBytes = BF_SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, block_fused_kernel<true>, tmX, tmW1, tmW2, a);
    if (e != cudaSuccess) { set_error(std::string("cudaLaunchKernelEx(block_fused, cluster 4): ") + cudaGetErrorString(e)); return 1; }
  }
  CONAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace conan
