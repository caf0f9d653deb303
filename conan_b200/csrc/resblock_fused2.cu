// Two-lane version of the fused residual-block kernel (resblock_fused.cu describes the data flow of one lane).
//
// One lane is latency-bound: between two convs of a tile lie ~1000 cycles of MMA drain (last issue -> accumulator readable)
// and ~700 cycles of epilogue, during which the tensor pipe idles.  Here a CTA works on TWO streams at once ("lanes") and the
// MMA warp alternates between them conv by conv, so one lane's drain + epilogue runs under the other lane's MMAs.
//
// What makes two lanes fit in shared memory (C = 64: 2 x 83 KB + 48 KB weight ring) is a ROTATION of three window buffers per
// lane instead of five windows + two input buffers:  window c of a tile lives in buffer c mod 3 -- conv c reads buffer c % 3,
// its epilogue writes the next window into buffer (c+1) % 3 and takes the residual from buffer (c-1) % 3; the TMA input box of
// the next tile lands in buffer 0 as soon as conv 3 (the last user of buffer 0) is done.  Because a buffer is reused by another
// window before the next tile comes round, the newest halo rows of every window are parked in a small per-lane halo store
// (and in the slot's resident history block after the last tile) and put back in front of the window body by the epilogue
// that writes that body.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace conan {

namespace {

constexpr int R2_CONVS = 6;
constexpr int R2_MAX_STAGES = 8;
__host__ __device__ constexpr int r2_epi_threads(int C) { return 128 * (C / 16); }      // one warpgroup per 16 output columns
__host__ __device__ constexpr int r2_threads(int C) { return 128 + r2_epi_threads(C); }

struct Fused2Args {
  int n_streams, L, k, tiles;
  int dil[3];
  int in_row0;
  int H[R2_CONVS];             // halo rows of window c
  int hs_row[R2_CONVS];        // first row of window c's halo inside the halo store / the slot's history block (c >= 1)
  int bufb;                    // bytes of one window buffer (1024-aligned)
  int lane_bytes;              // 3 * bufb + halo store
  int hs_off;                  // halo store offset inside a lane
  int wt_off, bar_off, bias_off, stages, group;
  const int* slot_ids;
  __half* hist; long long hist_slot_stride;
  __half* hist_out;            // compact [stream][hist rows][C]: the new history (scattered to the slots after the launch); null: in place
  int share_w;                 // a conv's taps fit the weight ring: while both lanes are active they use ONE copy of the stream
  int split;                   // lanes take balanced contiguous TILE ranges (a range may start inside a stream); 0: whole streams
  int n_lanes;
  const float* bias;
  const __half* sum_in; __half* sum_out; long long sum_slot_stride;
  __half* next; long long next_slot_stride; int next_row0;
  float out_scale, slope;
  long long* ts;               // optional timeline (CONAN_FUSED_TIMELINE): CTA 0 stamps clock64 per lane and conv step
};

template <int ROWB>
__device__ __forceinline__ uint32_t swz2(uint32_t off) {
  return off ^ (((off >> 7) & (ROWB == 128 ? 7u : 3u)) << 4);
}

// position of one lane in its sequence of (stream, tile, conv) steps; every warp role advances identical copies.
// A lane owns a contiguous range of the launch's tiles in (stream, tile) order.  With a.split the ranges are balanced to +-1 tile and
// may start inside a stream: the lane then first runs the tile BEFORE its range as a warm-up (outputs suppressed).  The six window
// halos it leaves depend on at most sum_c H[c] <= 128 input rows, all inside that tile, so the range continues with exactly the
// state a lane that had run the stream from its first tile would hold -- same bits, no dependence on where the cut falls.
struct LaneIter {
  int i, t, c, left, tiles, stride;
  bool warm;
  uint32_t steps, tiles_done;     // conv steps / tiles completed so far (barrier parities)
  __device__ __forceinline__ bool done() const { return left <= 0; }
  __device__ __forceinline__ void advance() {
    ++steps;
    if (++c == R2_CONVS) { c = 0; ++tiles_done; --left; warm = false; if (++t == tiles) { t = 0; i += stride; } }
  }
  __device__ __forceinline__ void init(int g, int n_lanes, int n_streams, int tiles_, int split) {
    tiles = tiles_; c = 0; steps = 0; tiles_done = 0; warm = false; stride = 1;
    if (split) {
      const long long T = (long long)n_streams * tiles_;
      const long long b = T * g / n_lanes, e = T * (g + 1) / n_lanes;
      i = (int)(b / tiles_); t = (int)(b - (long long)i * tiles_); left = (int)(e - b);
      if (left > 0 && t > 0) { --t; ++left; warm = true; }
    } else {                                      // whole streams, round robin: lane g takes streams g, g + n_lanes, ...
      i = g; t = 0; stride = n_lanes;
      left = g < n_streams ? (n_streams - g + n_lanes - 1) / n_lanes * tiles_ : 0;
    }
  }
};

template <int C, int KT>
__global__ void __launch_bounds__(r2_threads(C), C == 32 ? 2 : 1)
resblock_fused2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, Fused2Args a) {
  constexpr int ROWB = C * 2;
  constexpr int CH = ROWB / 16;
  constexpr int TAPB = C * ROWB;
  constexpr int TMEM_COLS = 4 * C;                  // 2 lanes x 2 accumulators
  constexpr int EPI = r2_epi_threads(C);
  constexpr int GROUP = C == 64 ? 2 : 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.bar_off);
  uint64_t* a_full = bars;                           // [lane]
  uint64_t* in_free = bars + 2;                      // [lane]  buffer 0 may take the next tile's input
  uint64_t* acc_full = bars + 4;                     // [lane][2]
  uint64_t* win_ready = bars + 8;                    // [lane][6]
  uint64_t* w_full = bars + 20;                      // [stages]
  uint64_t* w_empty = w_full + R2_MAX_STAGES;
  uint64_t* tile_done = w_empty + R2_MAX_STAGES;     // [lane]  every epilogue thread has finished the lane's previous tile
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tile_done + 2);
  float* s_bias = reinterpret_cast<float*>(smem + a.bias_off);
  for (int i = threadIdx.x; i < R2_CONVS * C; i += blockDim.x) s_bias[i] = a.bias[i];

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int l = 0; l < 2; ++l) {
      mbar_init(&a_full[l], 1); mbar_init(&in_free[l], EPI); mbar_init(&tile_done[l], EPI);
      mbar_init(&acc_full[l * 2], 1); mbar_init(&acc_full[l * 2 + 1], 1);
      for (int c = 0; c < R2_CONVS; ++c) mbar_init(&win_ready[l * R2_CONVS + c], EPI);
    }
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

  LaneIter it[2];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const int g = (int)blockIdx.x + l * (int)gridDim.x;
    if (g < a.n_lanes) it[l].init(g, a.n_lanes, a.n_streams, a.tiles, a.split);
    else { it[l].init(0, 1, 0, a.tiles, 0); }
  }

  if (warp == 0) {
    // ===================================================================== input-window producer (both lanes, tile by tile)
    const int in_rows = TILE_M + a.H[0];
    while (!it[0].done() || !it[1].done()) {
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        if (it[l].done()) continue;
        if (it[l].c == 0) {
          mbar_wait_lane0(&in_free[l], (it[l].tiles_done & 1) ^ 1, 32);
          if (elect_one_sync()) {
            mbar_expect_tx(&a_full[l], (uint32_t)(in_rows * ROWB));
            tma_load_3d(smem + l * a.lane_bytes, &tmA, &a_full[l], 0, a.in_row0 + it[l].t * TILE_M, it[l].i);
          }
        }
        it[l].advance();
      }
    }
  } else if (warp == 2) {
    // ===================================================================== weight producer: taps in the MMA warp's global order
    // (share_w: while both lanes are active they run the same conv index back to back, so lane 1 re-uses the stages lane 0 consumed)
    int s = 0;
    uint32_t ph = 1;
    while (!it[0].done() || !it[1].done()) {
      const bool both = a.share_w && !it[0].done() && !it[1].done();
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        if (it[l].done()) continue;
        const int c = it[l].c;
        if (!(both && l == 1)) {
          for (int j0 = 0; j0 < a.k; j0 += GROUP) {
            const int nt = min(GROUP, a.k - j0);
            mbar_wait_lane0(&w_empty[s], ph, 32);
            if (elect_one_sync()) {
              mbar_expect_tx(&w_full[s], (uint32_t)(nt * TAPB));
              uint8_t* dst = smem + a.wt_off + s * GROUP * TAPB;
              for (int j = 0; j < nt; ++j) tma_load_2d(dst + j * TAPB, &tmW, &w_full[s], (j0 + j) * C, c * C);
            }
            if (++s == a.stages) { s = 0; ph ^= 1; }
          }
        }
        it[l].advance();
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer: alternates between the lanes conv by conv
    constexpr uint32_t idesc = make_idesc<C>();
    const uint32_t s32 = smem_u32(smem);
    const uint64_t desc0 = make_smem_desc<ROWB>(s32);
    const uint64_t wdesc0 = desc0 + (uint64_t)(a.wt_off >> 4);
    constexpr uint32_t stage_step = (uint32_t)((GROUP * TAPB) >> 4);
    int s = 0;
    uint32_t wph = 0;
    const int kk_taps = KT > 0 ? KT : a.k;
    while (!it[0].done() || !it[1].done()) {
      const bool both = a.share_w && !it[0].done() && !it[1].done();
      const int s_round = s;
      const uint32_t wph_round = wph;
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        if (it[l].done()) continue;
        const int c = it[l].c;
        const bool keep = both && l == 0;          // lane 1 runs the same conv next: leave the stages full
        if (both && l == 1) { s = s_round; wph = wph_round; }
        if (c == 0) mbar_wait_warp(&a_full[l], it[l].tiles_done & 1);
        else mbar_wait_warp(&win_ready[l * R2_CONVS + c], it[l].tiles_done & 1);
        tc_fence_after();
        if (a.ts && blockIdx.x == 0 && it[l].steps < 48 && elect_one_sync()) a.ts[(l * 48 + it[l].steps) * 8 + 0] = clock64();
        const int cm3 = c >= 3 ? c - 3 : c;                                   // c % 3
        uint64_t ad = desc0 + (uint64_t)((l * a.lane_bytes + cm3 * a.bufb) >> 4);
        const uint32_t tap_step = (uint32_t)((((c & 1) ? 1 : a.dil[c >> 1]) * ROWB) >> 4);
        const uint32_t ab = it[l].steps & 1;
        const uint32_t tacc = tmem_base + (uint32_t)((l * 2 + ab) * C);
#pragma unroll
        for (int j0 = 0; j0 < (KT > 0 ? KT : 64); j0 += GROUP) {
          if (KT == 0 && j0 >= kk_taps) break;
          mbar_wait_warp(&w_full[s], wph);
          tc_fence_after();
          uint64_t bd = wdesc0 + (uint64_t)(s * stage_step);
#pragma unroll
          for (int j = 0; j < GROUP; ++j) {
            if (j0 + j < kk_taps) {
              tc_mma_f16_tap<C / 16>(tacc, ad, bd, idesc, (j0 | j) == 0 ? 1u : 0u);
              ad += tap_step; bd += (TAPB >> 4);
            }
          }
          if (!keep && elect_one_sync()) tc_commit(&w_empty[s]);
          if (++s == a.stages) { s = 0; wph ^= 1; }
        }
        if (elect_one_sync()) tc_commit(&acc_full[l * 2 + ab]);
        if (a.ts && blockIdx.x == 0 && it[l].steps < 48 && elect_one_sync()) a.ts[(l * 48 + it[l].steps) * 8 + 1] = clock64();
        it[l].advance();
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue warpgroups (shared by the two lanes, same order)
    const int etid = threadIdx.x - 128;
    const int wg = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const float inv_slope = 1.f / a.slope;
    const int col0 = wg * 16;
    __half* hist_l[2] = {nullptr, nullptr};          // the lane's current stream's history block (slot id read once per stream)
    while (!it[0].done() || !it[1].done()) {
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        if (it[l].done()) continue;
        const int c = it[l].c, t = it[l].t, i = it[l].i;
        const bool last = t == a.tiles - 1, warm = it[l].warm;
        uint8_t* lbase = smem + l * a.lane_bytes;
        uint8_t* hs = lbase + a.hs_off;
        // conv 0 of a tile depends only on the TMA box, so a fast thread could get here while a slow one still reads window 4
        // (residual of conv 5 of the previous tile) from the buffer that window 1 is about to be written into: wait until every
        // epilogue thread has finished that tile (found with compute-sanitizer's timing, invisible at full speed)
        if (c == 0) mbar_wait_lane0(&tile_done[l], (it[l].tiles_done & 1) ^ 1, 0);
        if ((t == 0 || warm) && c == 0)
          hist_l[l] = a.hist_out ? a.hist_out + (long long)i * a.hist_slot_stride
                                 : a.hist + (long long)(a.slot_ids ? a.slot_ids[i] : i) * a.hist_slot_stride;
        __half* hist = hist_l[l];                     // where the stream's NEW history goes (its last tile)
        if (warm && c == 0) {
          // warm-up tile of a range that starts inside a stream: nothing the lane keeps depends on the halos (see LaneIter); zeros
          const int hrows = a.hs_row[R2_CONVS - 1] + a.H[R2_CONVS - 1];
          for (int q = etid; q < hrows * CH; q += EPI) *reinterpret_cast<uint4*>(hs + q * 16) = make_uint4(0, 0, 0, 0);
          asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
        }
        if (t == 0 && c == 0) {
          const __half* hin = a.hist + (long long)(a.slot_ids ? a.slot_ids[i] : i) * a.hist_slot_stride;
          // new stream on this lane: its resident history becomes the halo store (every MMA of the lane's previous stream is
          // complete: these threads have passed its last accumulator barrier)
          const int hrows = a.hs_row[R2_CONVS - 1] + a.H[R2_CONVS - 1];
          for (int q = etid; q < hrows * CH; q += EPI)
            *reinterpret_cast<uint4*>(hs + q * 16) = *reinterpret_cast<const uint4*>(hin + (long long)q * 8);
          asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");             // halo store complete before anyone restores from it
        }
        const int cp = c > 0 ? c - 1 : 0, cn = c < R2_CONVS - 1 ? c + 1 : 0;
        const int cm3 = c >= 3 ? c - 3 : c, cp3 = cp >= 3 ? cp - 3 : cp, cn3 = cn >= 3 ? cn - 3 : cn;
        const long long grow = (long long)t * TILE_M + r;
        uint4 sprev[2];
        if (c == R2_CONVS - 1 && a.sum_in && !warm) {
          const __half* sp = a.sum_in + (long long)i * a.sum_slot_stride + grow * C + col0;
          sprev[0] = *reinterpret_cast<const uint4*>(sp); sprev[1] = *(reinterpret_cast<const uint4*>(sp) + 1);
        }
        const uint8_t* resw = lbase + cp3 * a.bufb;
        const uint32_t resrow = (uint32_t)((a.H[cp] + r) * ROWB);
        uint8_t* dstw = lbase + cn3 * a.bufb;
        const uint32_t dstrow = (uint32_t)((a.H[cn] + r) * ROWB);
        uint4 rv[2];
        if (c & 1) {
          rv[0] = *reinterpret_cast<const uint4*>(resw + swz2<ROWB>(resrow + (uint32_t)(col0 * 2)));
          rv[1] = *reinterpret_cast<const uint4*>(resw + swz2<ROWB>(resrow + (uint32_t)(col0 * 2 + 16)));
        }
        const uint32_t ab = it[l].steps & 1;
        mbar_wait_lane0(&acc_full[l * 2 + ab], (it[l].steps >> 1) & 1, 0);
        tc_fence_after();
        const bool stamp = a.ts && blockIdx.x == 0 && it[l].steps < 48 && etid == 0;
        if (a.ts) { if (stamp) a.ts[(l * 48 + it[l].steps) * 8 + 2] = clock64(); __syncwarp(); }
        uint32_t acc[16];
        tc_ld_32x32b_x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((l * 2 + ab) * C + col0), acc);
        if (a.ts) { if (stamp) a.ts[(l * 48 + it[l].steps) * 8 + 3] = clock64(); __syncwarp(); }
        float v[16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c * C + col0 + 4 * u]);
          v[4 * u] = __uint_as_float(acc[4 * u]) + b4.x;
          v[4 * u + 1] = __uint_as_float(acc[4 * u + 1]) + b4.y;
          v[4 * u + 2] = __uint_as_float(acc[4 * u + 2]) + b4.z;
          v[4 * u + 3] = __uint_as_float(acc[4 * u + 3]) + b4.w;
        }
        if (c & 1) {
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const __half2* hp = reinterpret_cast<const __half2*>(&rv[h8]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 f = __half22float2(hp[u]);
              v[h8 * 8 + 2 * u] += fminf(f.x, f.x * inv_slope);
              v[h8 * 8 + 2 * u + 1] += fminf(f.y, f.y * inv_slope);
            }
          }
        }
        if (c < R2_CONVS - 1) {
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            __half2 h[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float x0 = v[h8 * 8 + 2 * u], x1 = v[h8 * 8 + 2 * u + 1];
              h[u] = __floats2half2_rn(fmaxf(x0, x0 * a.slope), fmaxf(x1, x1 * a.slope));
            }
            *reinterpret_cast<uint4*>(dstw + swz2<ROWB>(dstrow + (uint32_t)(col0 * 2 + h8 * 16))) = *reinterpret_cast<uint4*>(h);
          }
          // the halo of window c+1 (newest rows of the previous tile, or the slot's history) goes in front of the body just written
          for (int q = etid; q < a.H[cn] * CH; q += EPI) {
            const int row = q / CH, ch = q - row * CH;
            *reinterpret_cast<uint4*>(dstw + swz2<ROWB>((uint32_t)(row * ROWB + ch * 16))) =
                *reinterpret_cast<const uint4*>(hs + ((a.hs_row[cn] + row) * CH + ch) * 16);
          }
        } else if (!warm) {
          if (a.sum_in) {
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              const __half2* hp = reinterpret_cast<const __half2*>(&sprev[h8]);
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
                const float x0 = v[h8 * 8 + 2 * u] * a.out_scale, x1 = v[h8 * 8 + 2 * u + 1] * a.out_scale;
                h[u] = __floats2half2_rn(fmaxf(x0, x0 * a.slope), fmaxf(x1, x1 * a.slope));
              }
              *(reinterpret_cast<uint4*>(nx) + h8) = *reinterpret_cast<uint4*>(h);
            }
          }
        }
        tc_fence_before();
        if (a.ts) { if (stamp) a.ts[(l * 48 + it[l].steps) * 8 + 4] = clock64(); __syncwarp(); }
        if (c < R2_CONVS - 1) {
          fence_proxy_async_smem();
          mbar_arrive(&win_ready[l * R2_CONVS + c + 1]);
        }
        if (c >= 1) {
          // window c has been consumed: park its newest halo rows (next tile) / write them to the slot's history (last tile)
          const uint8_t* srcw = lbase + cm3 * a.bufb;
          for (int q = etid; q < a.H[c] * CH; q += EPI) {
            const int row = q / CH, ch = q - row * CH;
            const uint4 hv = *reinterpret_cast<const uint4*>(srcw + swz2<ROWB>((uint32_t)((TILE_M + row) * ROWB + ch * 16)));
            if (last) *reinterpret_cast<uint4*>(hist + (long long)(a.hs_row[c] + row) * C + ch * 8) = hv;
            else *reinterpret_cast<uint4*>(hs + ((a.hs_row[c] + row) * CH + ch) * 16) = hv;
          }
        }
        if (c == 3) mbar_arrive(&in_free[l]);       // buffer 0 (input window, then window 3) is free for the next tile's input box
        if (c == R2_CONVS - 1) mbar_arrive(&tile_done[l]);
        it[l].advance();
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

inline int align1k2(int x) { return (x + 1023) & ~1023; }

// CONAN_FUSED_SPLIT: 0 = lanes always take whole streams, 1 (default) = tile ranges when there are fewer streams than lanes, 2 = always
inline int fused_split_mode() {
  static const int v = [] { const char* e = getenv("CONAN_FUSED_SPLIT"); return e ? atoi(e) : 1; }();
  return v;
}

template <int C, int KT>
int launch_fused2_variant(const CUtensorMap& tmA, const CUtensorMap& tmW, const Fused2Args& a, size_t smem, cudaStream_t st) {
  auto kern = resblock_fused2_kernel<C, KT>;
  static DeviceOnce once;
  if (device_once(once, nullptr, [&](int*) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return 1; }
        return 0;
      }))
    return 1;
  // one CTA per SM at C = 64, two at C = 32; two streams in flight per CTA
  // With at least one stream per lane the lanes take whole streams: lanes that run out leave the SM's operand bandwidth to the others,
  // so the uneven last round costs little (measured: cutting 1024 streams into 296 equal ranges is 5 % SLOWER, every cut pays a
  // warm-up tile).  With fewer streams than lanes the launch is cut into tile ranges so the idle lanes work: as many lanes as the
  // device holds, unless that leaves under ~3 tiles per lane.
  const int max_lanes = 2 * num_sms() * (C == 32 ? 2 : 1);
  Fused2Args b = a;
  b.split = a.split && (a.n_streams < max_lanes || fused_split_mode() == 2);
  if (b.split) b.n_lanes = (int)std::max(1LL, std::min((long long)max_lanes, (long long)a.n_streams * a.tiles / 3));
  else b.n_lanes = std::min(a.n_streams, max_lanes);
  const int grid = (b.n_lanes + 1) / 2;
  if (getenv("CONAN_TC_VERBOSE")) fprintf(stderr, "resblock_fused2<%d,%d> k %d tiles/stream %d smem %zu grid %d lanes %d split %d\n", C, KT, a.k, a.tiles, smem, grid, b.n_lanes, b.split);
  kern<<<grid, r2_threads(C), smem, st>>>(tmA, tmW, b);
  CONAN_CHECK_LAUNCH();
  return 0;
}

template <int C>
int launch_fused2_k(const CUtensorMap& tmA, const CUtensorMap& tmW, const Fused2Args& a, size_t smem, cudaStream_t st) {
  switch (a.k) {
    case 3: return launch_fused2_variant<C, 3>(tmA, tmW, a, smem, st);
    case 7: return launch_fused2_variant<C, 7>(tmA, tmW, a, smem, st);
    case 11: return launch_fused2_variant<C, 11>(tmA, tmW, a, smem, st);
    default: return launch_fused2_variant<C, 0>(tmA, tmW, a, smem, st);
  }
}

}  // namespace

int r2_stages32() {
  static int v = [] { const char* e = getenv("CONAN_FUSED2_STAGES32"); int x = e ? atoi(e) : 3; return x < 2 ? 2 : (x > 4 ? 4 : x); }();   // 3: k = 11 launch 366 -> 338 us, others unchanged; still two CTAs per SM
  return v;
}

// shared-memory bytes the two-lane kernel needs, or 0 if the shape does not fit
size_t resblock_fused2_smem(int C, int k, const int* dil) {
  const int ROWB = C * 2, TAPB = C * ROWB;
  int hmax = 0, hsum = 0;
  for (int c = 0; c < R2_CONVS; ++c) {
    const int h = (k - 1) * ((c & 1) ? 1 : dil[c >> 1]);
    hmax = std::max(hmax, h);
    if (c >= 1) hsum += h;
  }
  const int bufb = align1k2((TILE_M + hmax) * ROWB);
  const int lane = 3 * bufb + align1k2(hsum * ROWB);
  const int stages = C == 64 ? 3 : r2_stages32(), group = C == 64 ? 2 : 4;
  const size_t total = (size_t)2 * lane + (size_t)stages * group * TAPB + 512 + R2_CONVS * C * 4 + 1024;
  return total <= 227 * 1024 ? total : 0;
}

int launch_resblock_fused2(const ResblockFusedParams& p, cudaStream_t st) {
  if (!resblock_fused_eligible(p.C, p.L, p.k, p.dil) || resblock_fused2_smem(p.C, p.k, p.dil) == 0) {
    set_error("resblock_fused2: shape not eligible");
    return 1;
  }
  if (!(p.slope > 0.f && p.slope < 1.f)) { set_error("resblock_fused2: LeakyReLU slope must be in (0, 1)"); return 1; }
  if (p.n_streams <= 0) return 0;
  const int C = p.C, ROWB = C * 2, TAPB = C * ROWB;
  Fused2Args a;
  memset(&a, 0, sizeof(a));
  a.n_streams = p.n_streams; a.L = p.L; a.k = p.k; a.tiles = p.L / TILE_M;
  int hmax = 0, hrow = 0;
  for (int j = 0; j < 3; ++j) a.dil[j] = p.dil[j];
  for (int c = 0; c < R2_CONVS; ++c) {
    a.H[c] = (p.k - 1) * ((c & 1) ? 1 : p.dil[c >> 1]);
    hmax = std::max(hmax, a.H[c]);
    a.hs_row[c] = c >= 1 ? hrow : 0;
    if (c >= 1) hrow += a.H[c];
  }
  if (p.x_hist_rows < a.H[0]) { set_error("resblock_fused2: input context keeps too little history"); return 1; }
  a.in_row0 = p.x_hist_rows - a.H[0];
  a.bufb = align1k2((TILE_M + hmax) * ROWB);
  a.hs_off = 3 * a.bufb;
  a.lane_bytes = 3 * a.bufb + align1k2(hrow * ROWB);
  a.group = C == 64 ? 2 : 4;
  a.stages = C == 64 ? 3 : r2_stages32();
  int off = 2 * a.lane_bytes;
  a.wt_off = off; off += a.stages * a.group * TAPB;
  a.bar_off = off; off += 512;
  a.bias_off = off; off += R2_CONVS * C * 4;
  const size_t smem = (size_t)off + 1024;
  a.slot_ids = p.slot_ids; a.hist = (__half*)p.hist; a.hist_slot_stride = p.hist_slot_stride; a.bias = p.bias;
  a.hist_out = (__half*)p.hist_out;
  {
    // one copy of the weight stream for both lanes when a conv's tap groups fit the ring with a stage to spare for the prefetch
    static const int share_env = [] { const char* v = getenv("CONAN_FUSED_SHARE_W"); return v ? atoi(v) : 1; }();
    a.share_w = share_env && (p.k + a.group - 1) / a.group < a.stages;
  }
  {
    // a range may start inside a stream only if (1) the new history does not land where another lane still reads the old one and
    // (2) one warm-up tile rebuilds every window halo: sum of the six halos <= 128 rows
    const int want = fused_split_mode();
    int hsum = 0;
    for (int c = 0; c < R2_CONVS; ++c) hsum += a.H[c];
    a.split = want && a.hist_out && hsum <= TILE_M && a.tiles > 1;
  }
  a.sum_in = (const __half*)p.sum_in; a.sum_out = (__half*)p.sum_out; a.sum_slot_stride = (long long)p.L * C;
  a.next = (__half*)p.next; a.next_slot_stride = p.next_slot_stride; a.next_row0 = p.next_row0;
  a.out_scale = p.out_scale; a.slope = p.slope;
  CUtensorMap tmA, tmW;
  if (get_tensor_map(&tmA, p.x, 3, (unsigned long long)C, (unsigned long long)p.x_rows, (unsigned long long)p.n_slots,
                     (unsigned long long)ROWB, (unsigned long long)p.x_slot_stride * 2, C, TILE_M + a.H[0], 1, ROWB))
    return 1;
  const unsigned long long Ktot = (unsigned long long)p.k * C;
  if (get_tensor_map(&tmW, p.w, 2, Ktot, (unsigned long long)R2_CONVS * C, 1, Ktot * 2, 0, C, C, 1, ROWB)) return 1;
  if (getenv("CONAN_FUSED_TIMELINE")) {
    static long long* ts = nullptr;
    if (!ts) cudaMalloc(&ts, 2 * 48 * 8 * sizeof(long long));
    cudaMemsetAsync(ts, 0, 2 * 48 * 8 * sizeof(long long), st);
    a.ts = ts;
    int rc = C == 32 ? launch_fused2_k<32>(tmA, tmW, a, smem, st) : launch_fused2_k<64>(tmA, tmW, a, smem, st);
    static long long h[2 * 48 * 8];
    cudaMemcpyAsync(h, ts, sizeof(h), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    const long long t0 = h[(0 * 48 + 6) * 8];
    fprintf(stderr, "timeline2 C=%d k=%d: lane step c: mma_start mma_issue_end | acc_seen tmem_ld_done math_done   (cycles since lane 0 step 6)\n", C, p.k);
    for (int n = 6; n < 24; ++n)
      for (int l = 0; l < 2; ++l) {
        const long long* q = h + (l * 48 + n) * 8;
        fprintf(stderr, "  L%d %2d c=%d: %7lld %7lld | %7lld %7lld %7lld\n", l, n, n % 6, q[0] - t0, q[1] - t0, q[2] - t0, q[3] - t0, q[4] - t0);
      }
    return rc;
  }
  if (C == 32) return launch_fused2_k<32>(tmA, tmW, a, smem, st);
  return launch_fused2_k<64>(tmA, tmW, a, smem, st);
}

}  // namespace conan
