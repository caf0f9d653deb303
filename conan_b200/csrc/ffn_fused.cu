// Position-wise feed-forward block  y = W2 . relu(W1 . x + b1)  (Emformer `pos_ff`, TA:467-493 / _apply_post_attention_ffn
// TA:416-440) as ONE tcgen05 kernel with fp32-grade split-fp16 operands: the 2048-wide hidden activation never leaves the SM.
//
// Launched as two GEMMs the hidden tensor [rows, 2048] costs 50 MB written and read again per layer (split hi/lo fp16), which
// is what those two launches spend their time on.  Here a CTA owns a 128-row tile of x (both fp16 planes resident in shared
// memory, 48 KB) and a slice of the hidden dimension; per 128-wide hidden chunk:
//   GEMM1  acc1[128 x 128] = x_hi W1_hi + x_hi W1_lo + x_lo W1_hi          (18 MMAs, N = 128; W1 tiles stream through a TMA ring)
//   epilogue 1  h = relu(acc1 * 2^-10 + b1) -> hi = fp16(h), lo = fp16(h - hi) written into shared memory as the swizzled
//               K-major A operand of GEMM2 (two 64-column blocks per plane)
//   GEMM2  acc2[128 x 96] += h_hi W2_hi + h_hi W2_lo + h_lo W2_hi            (24 MMAs, N = 96; W2 tiles through the same ring)
// acc1 is double-buffered in TMEM so GEMM1 of chunk c+1 runs under epilogue 1 of chunk c.  The hidden dimension is split over
// FS CTAs per row tile (grid = row tiles x FS ~ one wave); each writes its partial acc2 * 2^-10 to P[fs][row][96] and the
// LayerNorm that follows sums the partials with b2 and the residual (deterministic: no atomics).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace conan {

namespace {

constexpr int FF_THREADS = 384;        // warp 0: TMA producer, warp 1: MMA issuer (+TMEM alloc), warps 4..11: epilogue
constexpr int FF_EPI = 256;
constexpr int FF_KP = 96;              // padded model dim (K of GEMM1, N of GEMM2)
constexpr int FF_CH = 128;             // hidden chunk
constexpr int FF_STAGES = 6;
constexpr int FF_STAGE_BYTES = 12288;  // max(W1 tile 128 x 32 halfs, W2 tile 96 x 64 halfs)
constexpr int FF_A_TILE = TILE_M * 64;         // 128 rows x 32 halfs (64-byte swizzle)
constexpr int FF_H_TILE = TILE_M * 128;        // 128 rows x 64 halfs (128-byte swizzle)
constexpr int FF_OFF_A = 0;
constexpr int FF_OFF_H = 6 * FF_A_TILE;                        // 49152
constexpr int FF_OFF_W = FF_OFF_H + 4 * FF_H_TILE;             // 114688
constexpr int FF_OFF_BAR = FF_OFF_W + FF_STAGES * FF_STAGE_BYTES;   // 188416
constexpr int FF_SMEM = FF_OFF_BAR + 1024 + 1024;              // + barriers + alignment slack

struct FfnArgs {
  int M;                      // valid rows
  int n_chunks;               // hidden / 128
  int FS;                     // hidden split
  int hidden;                 // F
  const float* b1;            // [F]
  float* P;                   // [FS][M][96] fp32 partial outputs
  float acc_scale;            // 2^-10 (both weight matrices are packed pre-scaled by 2^10)
};

__global__ void __launch_bounds__(FF_THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, FfnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FF_OFF_BAR);
  uint64_t* a_full = bars;                 // [1]
  uint64_t* acc1_full = bars + 1;          // [2]
  uint64_t* acc1_empty = bars + 3;         // [2]
  uint64_t* h_full = bars + 5;             // [1]
  uint64_t* h_empty = bars + 6;            // [1]
  uint64_t* acc2_full = bars + 7;          // [1]
  uint64_t* w_full = bars + 8;             // [STAGES]
  uint64_t* w_empty = w_full + FF_STAGES;  // [STAGES]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + FF_STAGES);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int mt = blockIdx.x / a.FS, fs = blockIdx.x - mt * a.FS;
  const int c_begin = (a.n_chunks * fs) / a.FS, c_end = (a.n_chunks * (fs + 1)) / a.FS, nc = c_end - c_begin;
  const int m0 = mt * TILE_M;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    mbar_init(a_full, 1); mbar_init(h_full, FF_EPI); mbar_init(h_empty, 1); mbar_init(acc2_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc1_full[s], 1); mbar_init(&acc1_empty[s], FF_EPI); }
    for (int s = 0; s < FF_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
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
    // ===================================================================== TMA producer: x tiles once, then the weight tiles in
    // exactly the order the MMA warp consumes them: G1(0), then per chunk [G1(c+1)], G2(c)
    if (elect_one_sync()) {
      mbar_expect_tx(a_full, 6 * FF_A_TILE);
      for (int pl = 0; pl < 2; ++pl)
        for (int kb = 0; kb < 3; ++kb) tma_load_3d(smem + FF_OFF_A + (pl * 3 + kb) * FF_A_TILE, &tmX, a_full, kb * 32, m0, pl);
    }
    int s = 0;
    uint32_t ph = 1;
    auto g1 = [&](int c) {
      const int f0 = (c_begin + c) * FF_CH;
      for (int kb = 0; kb < 9; ++kb) {
        mbar_wait_warp(&w_empty[s], ph);
        if (elect_one_sync()) {
          mbar_expect_tx(&w_full[s], TILE_M * 64);
          tma_load_2d(smem + FF_OFF_W + s * FF_STAGE_BYTES, &tmW1, &w_full[s], kb * 32, f0);
        }
        if (++s == FF_STAGES) { s = 0; ph ^= 1; }
      }
    };
    auto g2 = [&](int c) {
      const int f0 = (c_begin + c) * FF_CH;
      for (int t = 0; t < 6; ++t) {
        const int seg = t >> 1, kb2 = t & 1;
        mbar_wait_warp(&w_empty[s], ph);
        if (elect_one_sync()) {
          mbar_expect_tx(&w_full[s], FF_KP * 128);
          tma_load_2d(smem + FF_OFF_W + s * FF_STAGE_BYTES, &tmW2, &w_full[s], seg * a.hidden + f0 + kb2 * 64, 0);
        }
        if (++s == FF_STAGES) { s = 0; ph ^= 1; }
      }
    };
    g1(0);
    for (int c = 0; c < nc; ++c) {
      if (c + 1 < nc) g1(c + 1);
      g2(c);
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp-uniform, one elected lane issues)
    constexpr uint32_t idesc1 = make_idesc<FF_CH>();
    constexpr uint32_t idesc2 = make_idesc<FF_KP>();
    const uint32_t s32 = smem_u32(smem);
    int s = 0;
    uint32_t ph = 0;
    mbar_wait_warp(a_full, 0);
    tc_fence_after();
    auto g1 = [&](int c) {
      const int b = c & 1;
      mbar_wait_warp(&acc1_empty[b], ((c >> 1) & 1) ^ 1);           // epilogue 1 of chunk c-2 has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(b * FF_CH);
#pragma unroll
      for (int kb = 0; kb < 9; ++kb) {
        mbar_wait_warp(&w_full[s], ph);
        tc_fence_after();
        const int pl = kb >= 6 ? 1 : 0, akb = kb % 3;               // segments: x_hi W_hi | x_hi W_lo | x_lo W_hi
        const uint64_t ad = make_smem_desc<64>(s32 + FF_OFF_A + (pl * 3 + akb) * FF_A_TILE);
        const uint64_t bd = make_smem_desc<64>(s32 + FF_OFF_W + s * FF_STAGE_BYTES);
        tc_mma_f16_tap<2>(tacc, ad, bd, idesc1, kb == 0 ? 1u : 0u);
        if (elect_one_sync()) tc_commit(&w_empty[s]);
        if (++s == FF_STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one_sync()) tc_commit(&acc1_full[b]);
    };
    auto g2 = [&](int c) {
      mbar_wait_warp(h_full, c & 1);                                 // epilogue 1 of chunk c has written h (hi, lo)
      tc_fence_after();
      const uint32_t tacc = tmem_base + 256u;
#pragma unroll
      for (int t = 0; t < 6; ++t) {
        const int seg = t >> 1, kb2 = t & 1;
        mbar_wait_warp(&w_full[s], ph);
        tc_fence_after();
        const int pl = seg == 2 ? 1 : 0;                             // segments: h_hi W2_hi | h_hi W2_lo | h_lo W2_hi
        const uint64_t ad = make_smem_desc<128>(s32 + FF_OFF_H + (pl * 2 + kb2) * FF_H_TILE);
        const uint64_t bd = make_smem_desc<128>(s32 + FF_OFF_W + s * FF_STAGE_BYTES);
        tc_mma_f16_tap<4>(tacc, ad, bd, idesc2, (c == 0 && t == 0) ? 1u : 0u);
        if (elect_one_sync()) tc_commit(&w_empty[s]);
        if (++s == FF_STAGES) { s = 0; ph ^= 1; }
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
    for (int c = 0; c < nc; ++c) {
      const int b = c & 1;
      const float* b1 = a.b1 + (long long)(c_begin + c) * FF_CH + wg * 64;
      mbar_wait_lane0(&acc1_full[b], (c >> 1) & 1, 0);
      mbar_wait_lane0(h_empty, (c & 1) ^ 1, 0);                      // GEMM2 of chunk c-1 has finished reading h
      tc_fence_after();
      uint8_t* hhi = smem + FF_OFF_H + (0 * 2 + wg) * FF_H_TILE + r * 128;
      uint8_t* hlo = smem + FF_OFF_H + (1 * 2 + wg) * FF_H_TILE + r * 128;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t acc[16];
        tc_ld_32x32b_x16(tmem_base + lane_base + (uint32_t)(b * FF_CH + wg * 64 + ch * 16), acc);
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          __half2 hi[4], lo[4];
          const float4 ba = *reinterpret_cast<const float4*>(b1 + ch * 16 + h8 * 8);
          const float4 bb = *reinterpret_cast<const float4*>(b1 + ch * 16 + h8 * 8 + 4);
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float x0 = fmaxf(fmaf(__uint_as_float(acc[h8 * 8 + 2 * u]), a.acc_scale, bv[2 * u]), 0.f);
            const float x1 = fmaxf(fmaf(__uint_as_float(acc[h8 * 8 + 2 * u + 1]), a.acc_scale, bv[2 * u + 1]), 0.f);
            hi[u] = __floats2half2_rn(x0, x1);
            const float2 hf = __half22float2(hi[u]);
            lo[u] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
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
    // ---- final: partial y rows of this hidden slice
    mbar_wait_lane0(acc2_full, 0, 0);
    tc_fence_after();
    const int row = m0 + r;
    float* prow = a.P + ((long long)fs * a.M + row) * FF_KP + wg * 48;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      uint32_t acc[16];
      tc_ld_32x32b_x16(tmem_base + lane_base + (uint32_t)(256 + wg * 48 + ch * 16), acc);
      if (row < a.M) {
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
}

}  // namespace

int ffn_fused_split(int M) {
  // hidden-dimension split: a CONSTANT, so that the partial sums (and with them the last bits of the encoder output, and in a
  // near-tie the argmax token) do not depend on how many streams share the step.  3 slices: 48 row tiles x 3 = 144 CTAs = one wave
  // at 1024 streams (one CTA per SM: the kernel uses all of TMEM).
  (void)M;
  return 3;
}

bool ffn_fused_eligible(int K, int hidden, int N) {
  return K == FF_KP && N == FF_KP && hidden % FF_CH == 0 && hidden >= FF_CH;
}

int launch_ffn_fused(const FfnFusedParams& p, cudaStream_t st) {
  if (!ffn_fused_eligible(p.K, p.hidden, p.N)) { set_error("ffn_fused: shape not eligible (K = N = 96, hidden % 128 == 0)"); return 1; }
  if (p.M <= 0) return 0;
  const int mt = (p.M + TILE_M - 1) / TILE_M;
  if (p.FS < 1 || p.FS > p.hidden / FF_CH) { set_error("ffn_fused: bad hidden split"); return 1; }
  CUtensorMap tmX, tmW1, tmW2;
  // x: two fp16 planes [M_alloc][96]; rows past M are never stored (the tile may read rows of the allocation beyond M)
  if (get_tensor_map(&tmX, p.x, 3, FF_KP, (unsigned long long)p.x_rows, 2, (unsigned long long)FF_KP * 2,
                     (unsigned long long)p.x_lo_off * 2, 32, TILE_M, 1, 64))
    return 1;
  if (get_tensor_map(&tmW1, p.w1, 2, 3 * FF_KP, (unsigned long long)p.hidden, 1, (unsigned long long)3 * FF_KP * 2, 0, 32, FF_CH, 1, 64)) return 1;
  if (get_tensor_map(&tmW2, p.w2, 2, (unsigned long long)3 * p.hidden, FF_KP, 1, (unsigned long long)3 * p.hidden * 2, 0, 64, FF_KP, 1, 128)) return 1;
  FfnArgs a;
  a.M = p.M; a.n_chunks = p.hidden / FF_CH; a.FS = p.FS; a.hidden = p.hidden; a.b1 = p.b1; a.P = p.partials; a.acc_scale = p.acc_scale;
  static DeviceOnce once;
  if (device_once(once, nullptr, [&](int*) {
        cudaError_t e = cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
        return 0;
      }))
    return 1;
  ffn_fused_kernel<<<mt * p.FS, FF_THREADS, FF_SMEM, st>>>(tmX, tmW1, tmW2, a);
  CONAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace conan
