// One HiFi-GAN residual block (3 x [LeakyReLU -> dilated causal conv -> LeakyReLU -> causal conv -> + x],
// hifigan_causal.py:66-120) as ONE persistent tcgen05 kernel for the long narrow scales (C = 32 / 64).
//
// Launched conv by conv these layers are HBM-bound: every conv reads and writes the whole [streams, L, C]
// activation.  Here a CTA owns a stream and walks its L rows in 128-row tiles; per tile the six convs run back to
// back and the activation never leaves the SM:
//   * the input window (128 + halo rows of lrelu(x), fp16) arrives by one TMA box, double-buffered;
//   * conv c accumulates k taps x C/16 MMAs (M = 128, N = C) into a TMEM accumulator; every tap's A operand is the
//     window at a row offset (same single-window implicit GEMM as conv_window_tc_kernel);
//   * the epilogue warps read the accumulator, add bias (+ the residual x, recovered from the fp16 rows the
//     previous conv1 consumed), apply LeakyReLU and write fp16 rows straight into the NEXT conv's window in shared
//     memory, in the swizzled K-major layout the MMA reads -- that window's leading rows are the history of the
//     previous tile (moved down by exactly 128 rows, which preserves the swizzle phase) or, for the first tile of
//     the step, the slot's resident history block in HBM;
//   * weights stream through a ring of one-tap stages (they are L2-resident constants);
//   * the last conv adds the running MRF sum and writes either the new running sum or lrelu(sum / n_res) into the
//     following layer's context rows.
// HBM traffic per block drops from ~11 activation passes to input + sum in/out.  Numerics are those of the
// per-conv path (fp16 operands, fp32 accumulation, activations stored as fp16 at the same points).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace conan {

namespace {

constexpr int RF_CONVS = 6;
// warp 0: input-window TMA, warp 1: MMA issuer (+TMEM alloc), warp 2: weight TMA, warps 4..: epilogue warpgroups, one per
// 16 output columns (C/16 warpgroups: every epilogue thread owns one row x 16 columns = one tcgen05.ld.x16)
__host__ __device__ constexpr int rf_epi_threads(int C) { return 128 * (C / 16); }
__host__ __device__ constexpr int rf_threads(int C) { return 128 + rf_epi_threads(C); }
constexpr int RF_MAX_STAGES = 16;

struct FusedArgs {
  int n_streams, L, k, tiles;
  int dil[3];
  int in_row0;                 // first row of tile 0's window in the input context
  int H[RF_CONVS];             // halo rows of the window conv c reads ((k-1) * its dilation)
  int win_off[RF_CONVS];       // byte offset of window c in shared memory (c = 0: input buffer 0)
  int hist_off[RF_CONVS];      // first row of window c's history inside the slot's history block
  int in_winb, wt_off, bar_off, bias_off, stages, group, w_copies;
  int in_single;               // 1: one input-window buffer (C = 64: the shared memory goes to a deeper weight ring instead)
  long long* ts;               // optional timeline buffer (CONAN_FUSED_TIMELINE): CTA 0 stamps clock64 at each hand-over
  int dbg;                     // timing experiments only (CONAN_FUSED_DEBUG): 1 = weights fetched once, 2 = epilogue math skipped      // group: taps per weight stage
  const int* slot_ids;
  __half* hist; long long hist_slot_stride;                            // [slot][hist rows][C]
  const float* bias;                                                   // [6][C]
  const __half* sum_in; __half* sum_out; long long sum_slot_stride;    // compact [i][L][C]
  __half* next; long long next_slot_stride; int next_row0;             // compact context rows of the following layer
  float out_scale, slope;
};

template <int ROWB>
__device__ __forceinline__ uint32_t swz(uint32_t off) {      // byte offset inside a 1024-aligned window -> swizzled offset
  return off ^ (((off >> 7) & (ROWB == 128 ? 7u : 3u)) << 4);
}

// KT: kernel size known at compile time (3 / 7 / 11: the MMA issue loop is then fully unrolled, which matters -- the issuing
// warp, not the tensor pipe, paces a conv of k x C/16 short MMAs), or 0 for any k.
template <int C, int KT>
__global__ void __launch_bounds__(rf_threads(C), C == 32 ? 2 : 1)
resblock_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, FusedArgs a) {
  constexpr int ROWB = C * 2;
  constexpr int CH = ROWB / 16;
  constexpr int TAPB = C * ROWB;
  constexpr int TMEM_COLS = 2 * C < 32 ? 32 : 2 * C;
  constexpr int HALF = 16;                          // columns per epilogue warpgroup
  constexpr int RF_EPI = rf_epi_threads(C);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.bar_off);
  uint64_t* a_full = bars;                           // [2]
  uint64_t* a_empty = bars + 2;                      // [2]
  uint64_t* acc_full = bars + 4;                     // [2]
  uint64_t* win_ready = bars + 6;                    // [6] (1..5 used)
  uint64_t* w_full = bars + 12;                      // [stages]
  uint64_t* w_empty = w_full + RF_MAX_STAGES;        // [stages]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + RF_MAX_STAGES);
  float* s_bias = reinterpret_cast<float*>(smem + a.bias_off);
  for (int i = threadIdx.x; i < RF_CONVS * C; i += blockDim.x) s_bias[i] = a.bias[i];

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: provably warp-uniform role branches
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], RF_EPI); mbar_init(&acc_full[s], 1); }
    for (int c = 0; c < RF_CONVS; ++c) mbar_init(&win_ready[c], RF_EPI);
    for (int s = 0; s < a.stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
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
  const int in_rows = TILE_M + a.H[0];

  if (warp == 0) {
    // ===================================================================== input-window producer
    {
      int it = 0;
      for (int i = blockIdx.x; i < a.n_streams; i += gridDim.x)
        for (int t = 0; t < a.tiles; ++t, ++it) {
          const int buf = a.in_single ? 0 : (it & 1);
          mbar_wait_lane0(&a_empty[buf], ((a.in_single ? it : (it >> 1)) & 1) ^ 1, 64);
          if (elect_one_sync()) {
            mbar_expect_tx(&a_full[buf], (uint32_t)(in_rows * ROWB));
            tma_load_3d(smem + buf * a.in_winb, &tmA, &a_full[buf], 0, a.in_row0 + t * TILE_M, i);
          }
        }
    }
  } else if (warp == 2) {
    // ===================================================================== weight producer (`group` taps per stage)
    {
      int s = 0, cnt = 0;
      uint32_t ph = 1;                                      // parity that lets the first pass through the ring proceed
      const int wcopy = blockIdx.x % a.w_copies;
      for (int i = blockIdx.x; i < a.n_streams; i += gridDim.x)
        for (int t = 0; t < a.tiles; ++t)
          for (int c = 0; c < RF_CONVS; ++c)
            for (int j0 = 0; j0 < a.k; j0 += a.group, ++cnt) {
              const int nt = min(a.group, a.k - j0);
              if ((a.dbg & 1) && cnt >= a.stages) continue;
              mbar_wait_lane0(&w_empty[s], ph, 32);
              if (elect_one_sync()) {
                mbar_expect_tx(&w_full[s], (uint32_t)(nt * TAPB));
                uint8_t* dst = smem + a.wt_off + s * a.group * TAPB;
                for (int j = 0; j < nt; ++j) tma_load_3d(dst + j * TAPB, &tmW, &w_full[s], (j0 + j) * C, c * C, wcopy);
              }
              if (++s == a.stages) { s = 0; ph ^= 1; }
            }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // (the whole warp runs the loop so that descriptors stay in uniform registers; one elected lane issues.  The loop is
    // kept free of divisions and descriptor rebuilds: at k taps x C/16 MMAs of 16-48 cycles each its issue rate feeds the tensor pipe)
    {
      constexpr uint32_t idesc = make_idesc<C>();
      const uint32_t s32 = smem_u32(smem);
      const uint64_t desc0 = make_smem_desc<ROWB>(s32);                  // descriptor of the (1024-aligned) buffer base
      const uint64_t wdesc0 = desc0 + (uint64_t)(a.wt_off >> 4);
      const uint32_t stage_step = (uint32_t)((a.group * TAPB) >> 4);
      int it = 0, s = 0, cnt = 0;
      uint32_t n = 0, wph = 0;
      for (int i = blockIdx.x; i < a.n_streams; i += gridDim.x)
        for (int t = 0; t < a.tiles; ++t, ++it)
#pragma unroll
          for (int c = 0; c < RF_CONVS; ++c, ++n) {        // unrolled: window offsets / dilations become uniform constant loads
            const int ibuf = a.in_single ? 0 : (it & 1);
            if (c == 0) mbar_wait_warp(&a_full[ibuf], (a.in_single ? it : (it >> 1)) & 1);
            else mbar_wait_warp(&win_ready[c], it & 1);
            tc_fence_after();
            if (a.ts && blockIdx.x == 0 && n < 64 && elect_one_sync()) a.ts[n * 8 + 0] = clock64();      // window ready seen by the MMA warp
            const uint64_t adesc = desc0 + (uint64_t)((c == 0 ? ibuf * a.in_winb : a.win_off[c]) >> 4);
            const uint32_t tap_step = (uint32_t)((((c & 1) ? 1 : a.dil[c >> 1]) * ROWB) >> 4);
            const uint32_t tacc = tmem_base + (uint32_t)((n & 1) * C);
            uint64_t ad = adesc;
            constexpr int GROUP = C == 64 ? 2 : 4;            // taps per weight stage (the host sets a.group to the same value)
            const int kk_taps = KT > 0 ? KT : a.k;
#pragma unroll
            for (int j0 = 0; j0 < (KT > 0 ? KT : 64); j0 += GROUP) {
              if (KT == 0 && j0 >= kk_taps) break;
              if (!((a.dbg & 1) && cnt >= a.stages)) mbar_wait_warp(&w_full[s], wph);
              tc_fence_after();
              if (a.ts && blockIdx.x == 0 && n < 64 && j0 == 0 && elect_one_sync()) a.ts[n * 8 + 6] = clock64();   // first weight group landed
              uint64_t bd = wdesc0 + (uint64_t)(s * stage_step);
#pragma unroll
              for (int j = 0; j < GROUP; ++j) {
                if (j0 + j < kk_taps) {
                  tc_mma_f16_tap<C / 16>(tacc, ad, bd, idesc, (j0 | j) == 0 ? 1u : 0u);
                  ad += tap_step; bd += (TAPB >> 4);
                }
              }
              if (!(a.dbg & 1) && elect_one_sync()) tc_commit(&w_empty[s]);
              if (a.ts && blockIdx.x == 0 && n < 64 && j0 == 0 && elect_one_sync()) a.ts[n * 8 + 7] = clock64();   // first group issued
              if (++s == a.stages) { s = 0; wph ^= 1; }
              ++cnt;
            }
            if (elect_one_sync()) tc_commit(&acc_full[n & 1]);
            if (a.ts && blockIdx.x == 0 && n < 64 && elect_one_sync()) a.ts[n * 8 + 1] = clock64();      // last MMA + commit issued
          }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue (two warpgroups, half the columns each)
    const int etid = threadIdx.x - 128;
    const int wg = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const float inv_slope = 1.f / a.slope;
    int it = 0;
    uint32_t n = 0;
    for (int i = blockIdx.x; i < a.n_streams; i += gridDim.x) {
      const int slot = a.slot_ids ? a.slot_ids[i] : i;
      __half* hist = a.hist + (long long)slot * a.hist_slot_stride;
      // the slot's history rows become the leading rows of windows 1..5 (every MMA of the previous stream is complete:
      // these threads have passed its last accumulator barrier)
      for (int w = 1; w < RF_CONVS; ++w)
        for (int q = etid; q < a.H[w] * CH; q += RF_EPI) {
          const int row = q / CH, ch = q - row * CH;
          const uint4 v = *reinterpret_cast<const uint4*>(hist + (long long)(a.hist_off[w] + row) * C + ch * 8);
          *reinterpret_cast<uint4*>(smem + a.win_off[w] + swz<ROWB>((uint32_t)(row * ROWB + ch * 16))) = v;
        }
      for (int t = 0; t < a.tiles; ++t, ++it) {
        const bool last = t == a.tiles - 1;
        const long long grow = (long long)t * TILE_M + r;            // row of this thread inside the stream's L rows
#pragma unroll 1
        for (int c = 0; c < RF_CONVS; ++c, ++n) {
          const int ab = n & 1;
          uint4 sprev[HALF / 8];
          if (c == RF_CONVS - 1 && a.sum_in) {                       // running-sum rows do not depend on the accumulator
            const __half* sp = a.sum_in + (long long)i * a.sum_slot_stride + grow * C + wg * HALF;
#pragma unroll
            for (int u = 0; u < HALF / 8; ++u) sprev[u] = *(reinterpret_cast<const uint4*>(sp) + u);
          }
          const uint8_t* resw = smem + (c == 1 ? (a.in_single ? 0 : (it & 1)) * a.in_winb : a.win_off[c > 0 ? c - 1 : 0]);
          const uint32_t resrow = (uint32_t)((a.H[c > 0 ? c - 1 : 0] + r) * ROWB);
          uint8_t* dstw = smem + a.win_off[c < RF_CONVS - 1 ? c + 1 : 0];
          const uint32_t dstrow = (uint32_t)((a.H[c < RF_CONVS - 1 ? c + 1 : 0] + r) * ROWB);
          // x (residual of the second conv of a pair): the rows conv c-1 consumed, already in shared memory -- fetched
          // before the accumulator wait
          uint4 rv[HALF / 8];
          if (c & 1) {
#pragma unroll
            for (int u = 0; u < HALF / 8; ++u) rv[u] = *reinterpret_cast<const uint4*>(resw + swz<ROWB>(resrow + (uint32_t)(wg * HALF * 2 + u * 16)));
          }
          mbar_wait_lane0(&acc_full[ab], (n >> 1) & 1, a.dbg & 4 ? 32 : 0);
          tc_fence_after();
          const bool stamp = a.ts && blockIdx.x == 0 && n < 64 && etid == 0;
          if (a.ts) { if (stamp) a.ts[n * 8 + 2] = clock64(); __syncwarp(); }                                                   // accumulator complete seen by the epilogue
          const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * C + wg * HALF);
          uint32_t acc[HALF];
#pragma unroll
          for (int ch16 = 0; ch16 < HALF / 16; ++ch16) tc_ld_32x32b_x16_nowait(tl + (uint32_t)(ch16 * 16), &acc[ch16 * 16]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (a.ts) { if (stamp) a.ts[n * 8 + 3] = clock64(); __syncwarp(); }                                                   // TMEM read done
#pragma unroll
          for (int ch16 = 0; ch16 < ((a.dbg & 2) ? 0 : HALF / 16); ++ch16) {
            const int col0 = wg * HALF + ch16 * 16;
            float v[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c * C + col0 + 4 * u]);
              v[4 * u] = __uint_as_float(acc[ch16 * 16 + 4 * u]) + b4.x;
              v[4 * u + 1] = __uint_as_float(acc[ch16 * 16 + 4 * u + 1]) + b4.y;
              v[4 * u + 2] = __uint_as_float(acc[ch16 * 16 + 4 * u + 2]) + b4.z;
              v[4 * u + 3] = __uint_as_float(acc[ch16 * 16 + 4 * u + 3]) + b4.w;
            }
            if (c & 1) {                                              // + x: inverse LeakyReLU of the rows conv c-1 consumed
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {
                const __half2* hp = reinterpret_cast<const __half2*>(&rv[ch16 * 2 + h8]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float2 f = __half22float2(hp[u]);
                  v[h8 * 8 + 2 * u] += fminf(f.x, f.x * inv_slope);          // inverse LeakyReLU (slope < 1): min(h, h / slope)
                  v[h8 * 8 + 2 * u + 1] += fminf(f.y, f.y * inv_slope);
                }
              }
            }
            if (c < RF_CONVS - 1) {
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {
                __half2 h[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float x0 = v[h8 * 8 + 2 * u], x1 = v[h8 * 8 + 2 * u + 1];
                  x0 = fmaxf(x0, x0 * a.slope); x1 = fmaxf(x1, x1 * a.slope);      // LeakyReLU (slope < 1): max(x, slope x)
                  h[u] = __floats2half2_rn(x0, x1);
                }
                *reinterpret_cast<uint4*>(dstw + swz<ROWB>(dstrow + (uint32_t)(col0 * 2 + h8 * 16))) = *reinterpret_cast<uint4*>(h);
              }
            } else {
              if (a.sum_in) {
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                  const __half2* hp = reinterpret_cast<const __half2*>(&sprev[ch16 * 2 + h8]);
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const float2 f = __half22float2(hp[u]);
                    v[h8 * 8 + 2 * u] += f.x; v[h8 * 8 + 2 * u + 1] += f.y;
                  }
                }
              }
              if (a.sum_out) {
                __half* so = a.sum_out + (long long)i * a.sum_slot_stride + grow * C + col0;
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                  __half2 h[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) h[u] = __floats2half2_rn(v[h8 * 8 + 2 * u], v[h8 * 8 + 2 * u + 1]);
                  *(reinterpret_cast<uint4*>(so) + h8) = *reinterpret_cast<uint4*>(h);
                }
              }
              if (a.next) {
                __half* nx = a.next + (long long)i * a.next_slot_stride + ((long long)a.next_row0 + grow) * C + col0;
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                  __half2 h[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    float x0 = v[h8 * 8 + 2 * u] * a.out_scale, x1 = v[h8 * 8 + 2 * u + 1] * a.out_scale;
                    x0 = fmaxf(x0, x0 * a.slope); x1 = fmaxf(x1, x1 * a.slope);      // LeakyReLU (slope < 1): max(x, slope x)
                    h[u] = __floats2half2_rn(x0, x1);
                  }
                  *(reinterpret_cast<uint4*>(nx) + h8) = *reinterpret_cast<uint4*>(h);
                }
              }
            }
          }
          if (c == 1) mbar_arrive(&a_empty[a.in_single ? 0 : (it & 1)]);                 // the input window (A operand of conv 0, residual of conv 1) is free
          tc_fence_before();
          if (a.ts) { if (stamp) a.ts[n * 8 + 4] = clock64(); __syncwarp(); }                                                   // math + stores done
          if (c < RF_CONVS - 1) {
            fence_proxy_async_smem();                                 // rows written above are read by tcgen05.mma
            mbar_arrive(&win_ready[c + 1]);
          }
          if (a.ts) { if (stamp) a.ts[n * 8 + 5] = clock64(); __syncwarp(); }                                                   // hand-over signalled
          if (c >= 1) {
            // window c has been consumed (its accumulator is complete): its newest H rows are the next tile's history, or the
            // slot's after the last tile.  Off the critical path: the next reader of these rows is conv c of the NEXT tile,
            // released by a later win_ready[c] arrive of these same threads.
            for (int q = etid; q < a.H[c] * CH; q += RF_EPI) {
              const int row = q / CH, ch = q - row * CH;
              const uint4 v = *reinterpret_cast<const uint4*>(smem + a.win_off[c] + swz<ROWB>((uint32_t)((TILE_M + row) * ROWB + ch * 16)));
              if (last) *reinterpret_cast<uint4*>(hist + (long long)(a.hist_off[c] + row) * C + ch * 8) = v;
              else *reinterpret_cast<uint4*>(smem + a.win_off[c] + swz<ROWB>((uint32_t)(row * ROWB + ch * 16))) = v;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

inline int align1k(int x) { return (x + 1023) & ~1023; }

template <int C, int KT>
int launch_fused_variant(const CUtensorMap& tmA, const CUtensorMap& tmW, const FusedArgs& a, size_t smem, cudaStream_t st) {
  auto kern = resblock_fused_kernel<C, KT>;
  static DeviceOnce once;
  if (device_once(once, nullptr, [&](int*) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
        return 0;
      }))
    return 1;
  const int per_sm = resident_ctas((const void*)kern, rf_threads(C), smem, 2 * C < 32 ? 32 : 2 * C);
  const int grid = std::min(a.n_streams, num_sms() * per_sm);
  if (getenv("CONAN_TC_VERBOSE")) fprintf(stderr, "resblock_fused<%d,%d> k %d tiles/stream %d smem %zu per_sm %d grid %d\n", C, KT, a.k, a.tiles, smem, per_sm, grid);
  kern<<<grid, rf_threads(C), smem, st>>>(tmA, tmW, a);
  CONAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace

namespace {
template <int C>
int launch_fused_k(const CUtensorMap& tmA, const CUtensorMap& tmW, const FusedArgs& a, size_t smem, cudaStream_t st) {
  switch (a.k) {
    case 3: return launch_fused_variant<C, 3>(tmA, tmW, a, smem, st);
    case 7: return launch_fused_variant<C, 7>(tmA, tmW, a, smem, st);
    case 11: return launch_fused_variant<C, 11>(tmA, tmW, a, smem, st);
    default: return launch_fused_variant<C, 0>(tmA, tmW, a, smem, st);
  }
}
}  // namespace

int resblock_fused_hist_rows(int k, const int* dil) {
  return (k - 1) * (1 + dil[1] + 1 + dil[2] + 1);
}

bool resblock_fused_eligible(int C, int L, int k, const int* dil) {
  if (!(C == 32 || C == 64) || L % TILE_M != 0 || k < 1) return false;
  for (int j = 0; j < 3; ++j)
    if (TILE_M + (k - 1) * dil[j] > 256) return false;               // TMA box / window rows
  return true;
}

int launch_resblock_fused(const ResblockFusedParams& p, cudaStream_t st) {
  if (!resblock_fused_eligible(p.C, p.L, p.k, p.dil)) { set_error("resblock_fused: shape not eligible"); return 1; }
  if (!(p.slope > 0.f && p.slope < 1.f)) { set_error("resblock_fused: LeakyReLU slope must be in (0, 1)"); return 1; }
  if (p.n_streams <= 0) return 0;
  const int C = p.C, ROWB = C * 2, TAPB = C * ROWB;
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.n_streams = p.n_streams; a.L = p.L; a.k = p.k; a.tiles = p.L / TILE_M;
  for (int j = 0; j < 3; ++j) a.dil[j] = p.dil[j];
  for (int c = 0; c < RF_CONVS; ++c) a.H[c] = (p.k - 1) * ((c & 1) ? 1 : p.dil[c >> 1]);
  if (p.x_hist_rows < a.H[0]) { set_error("resblock_fused: input context keeps too little history"); return 1; }
  a.in_row0 = p.x_hist_rows - a.H[0];
  a.in_winb = align1k((TILE_M + a.H[0]) * ROWB);
  // TMA round trips are ~1 us under load and a weight stage is only recycled when its MMAs have completed, so weight streaming
  // is bound by the bytes in flight: at C = 64 (8 KB per tap, consumed in ~200 cycles) the ring gets the second input buffer's
  // space -- the next tile's input is not needed before four more convs have run
  { static int single = [] { const char* v = getenv("CONAN_FUSED_SINGLE"); return v ? atoi(v) : 1; }(); a.in_single = (C == 64 && single) ? 1 : 0; }
  int off = (a.in_single ? 1 : 2) * a.in_winb, hrow = 0;
  a.win_off[0] = 0; a.hist_off[0] = 0;
  for (int c = 1; c < RF_CONVS; ++c) {
    a.win_off[c] = off; off += align1k((TILE_M + a.H[c]) * ROWB);
    a.hist_off[c] = hrow; hrow += a.H[c];
  }
  a.wt_off = off;
  a.group = C == 64 ? 2 : 4;                 // 16 KB / 8 KB weight stages
  a.stages = a.in_single ? 6 : 4;
  off += a.stages * a.group * TAPB;
  a.bar_off = off; off += 512;
  a.bias_off = off; off += RF_CONVS * C * 4;
  const size_t smem = (size_t)off + 1024;
  if (smem > 227 * 1024 - 1024) { set_error("resblock_fused: windows do not fit in shared memory"); return 1; }
  a.slot_ids = p.slot_ids; a.hist = (__half*)p.hist; a.hist_slot_stride = p.hist_slot_stride; a.bias = p.bias;
  a.sum_in = (const __half*)p.sum_in; a.sum_out = (__half*)p.sum_out; a.sum_slot_stride = (long long)p.L * C;
  a.next = (__half*)p.next; a.next_slot_stride = p.next_slot_stride; a.next_row0 = p.next_row0;
  a.out_scale = p.out_scale; a.slope = p.slope;
  CUtensorMap tmA, tmW;
  if (get_tensor_map(&tmA, p.x, 3, (unsigned long long)C, (unsigned long long)p.x_rows, (unsigned long long)p.n_slots,
                     (unsigned long long)ROWB, (unsigned long long)p.x_slot_stride * 2, C, TILE_M + a.H[0], 1, ROWB))
    return 1;
  const unsigned long long Ktot = (unsigned long long)p.k * C;
  a.w_copies = p.w_copies > 0 ? p.w_copies : 1;
  { static int dbg = [] { const char* v = getenv("CONAN_FUSED_DEBUG"); return v ? atoi(v) : 0; }(); a.dbg = dbg; }

  if (get_tensor_map(&tmW, p.w, 3, Ktot, (unsigned long long)RF_CONVS * C, (unsigned long long)a.w_copies, Ktot * 2,
                     Ktot * 2 * RF_CONVS * C, C, C, 1, ROWB))
    return 1;
  if (getenv("CONAN_FUSED_TIMELINE")) {
    // developer aid: CTA 0 stamps the first 64 conv steps; printed after the launch (synchronises the stream)
    static long long* ts = nullptr;
    if (!ts) cudaMalloc(&ts, 64 * 8 * sizeof(long long));
    cudaMemsetAsync(ts, 0, 64 * 8 * sizeof(long long), st);
    a.ts = ts;
    int rc = C == 32 ? launch_fused_k<32>(tmA, tmW, a, smem, st) : launch_fused_k<64>(tmA, tmW, a, smem, st);
    long long h[64 * 8];
    cudaMemcpyAsync(h, ts, sizeof(h), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    fprintf(stderr, "timeline C=%d k=%d: step: mma_start  [w0_wait g0_issue rest] mma_issue_span  issue_end->acc_seen  tmem_ld  math  signal  | epilogue_signal->next_mma_start\n", C, p.k);
    for (int n = 6; n < 30; ++n)
      fprintf(stderr, "  %2d c=%d: %6lld [%5lld %5lld %5lld] %6lld %6lld %6lld %6lld %6lld | %6lld\n", n, n % 6, h[n * 8] - h[6 * 8], h[n * 8 + 6] - h[n * 8], h[n * 8 + 7] - h[n * 8 + 6],
              h[n * 8 + 1] - h[n * 8 + 7], h[n * 8 + 1] - h[n * 8], h[n * 8 + 2] - h[n * 8 + 1],
              h[n * 8 + 3] - h[n * 8 + 2], h[n * 8 + 4] - h[n * 8 + 3], h[n * 8 + 5] - h[n * 8 + 4], h[(n + 1) * 8] - h[n * 8 + 5]);
    return rc;
  }
  if (C == 32) return launch_fused_k<32>(tmA, tmW, a, smem, st);
  return launch_fused_k<64>(tmA, tmW, a, smem, st);
}

}  // namespace conan
