// Implicit-GEMM 1-D convolution over per-slot context buffers, FP32 accumulate on the
// CUDA cores (FFMA).  This is the exact-fp32 engine of the path: every Emformer / Conan
// contraction runs through it (their discrete decisions -- argmax over 100 logits, VQ
// argmin, uv > 0, f0 bucket -- need fp32-grade operands), and it is
// the numerical cross-check for the tcgen05 engine in conv_gemm_tc.cu.
//
//   M = n_streams * L (stream-major, then time), N = cout, K = k * cin traversed tap-major.
//   A[m, j*cin + c] = X[slot(m), row0 + t(m) + j*dil, c]   gathered straight from the context
//   buffer (no im2col materialisation); W is pre-packed [cout, k*cin].
//
// Tile 128 x 64 x 16, 256 threads, 8 x 4 outputs per thread, register-prefetch double
// buffering through shared memory.  Operands may be fp32 or fp16 (converted on the way
// into shared memory); accumulation and the fused epilogue are always fp32.
#include "common.cuh"

namespace conan {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int APAD = 4, BPAD = 4;

template <typename T> struct LoadVec;
template <> struct LoadVec<float> {
  static constexpr int kPerRow = BK / 4;   // float4 loads per tile row
  static constexpr int kElems = 4;
  using V = float4;
  __device__ static void unpack(const V& v, float* f) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
};
template <> struct LoadVec<__half> {
  static constexpr int kPerRow = BK / 8;   // 16-byte loads (8 halfs) per tile row
  static constexpr int kElems = 8;
  using V = uint4;
  __device__ static void unpack(const V& v, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
};

struct EpiArgs {
  const float* bias; float scale; int act; float slope;
  const void* res; long long res_slot_stride; int res_row_stride;
  const float* rowmask; int mask_slot_stride; float out_scale;
  void* y; long long y_slot_stride; int y_row_stride, y_row0; int accumulate;
  void* y2; long long y2_slot_stride; int y2_row_stride, y2_row0; int y2_is_half; int act2; float slope2;
  int res_is_half; float res_inv_slope;
  const void* res2; long long res2_slot_stride; int res2_row_stride; int res2_is_half; int y_is_half;
};

template <typename T>
__global__ void __launch_bounds__(NT)
conv_gemm_ffma_kernel(const T* __restrict__ X, long long x_slot_stride, int x_row_stride, int row0, int L,
                      int cin, int k, int dil, int cout, const T* __restrict__ W, int n_streams,
                      const int* __restrict__ slot_ids, EpiArgs e) {
  using LV = LoadVec<T>;
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN + BPAD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int M = n_streams * L;
  const int Ktot = k * cin;
  const int nslices = Ktot / BK;

  // ---- per-thread global->smem load assignment
  constexpr int A_LOADS = (BM * LV::kPerRow + NT - 1) / NT;     // 2 (fp32) or 1 (fp16)
  constexpr int B_LOADS = (BN * LV::kPerRow + NT - 1) / NT;     // 1
  long long a_base[A_LOADS]; bool a_ok[A_LOADS]; int a_row[A_LOADS], a_kq[A_LOADS];
#pragma unroll
  for (int r = 0; r < A_LOADS; ++r) {
    int idx = tid + r * NT;
    a_row[r] = idx / LV::kPerRow; a_kq[r] = idx % LV::kPerRow;
    int m = m0 + a_row[r];
    a_ok[r] = (a_row[r] < BM) && (m < M);
    int i = a_ok[r] ? m / L : 0, t = a_ok[r] ? m - i * L : 0;
    int slot = slot_ids ? slot_ids[i] : i;
    a_base[r] = (long long)slot * x_slot_stride + (long long)(row0 + t) * x_row_stride + a_kq[r] * LV::kElems;
  }
  long long b_base[B_LOADS]; bool b_ok[B_LOADS]; int b_row[B_LOADS], b_kq[B_LOADS];
#pragma unroll
  for (int r = 0; r < B_LOADS; ++r) {
    int idx = tid + r * NT;
    b_row[r] = idx / LV::kPerRow; b_kq[r] = idx % LV::kPerRow;
    b_ok[r] = (b_row[r] < BN) && (n0 + b_row[r] < cout);
    b_base[r] = (long long)(n0 + b_row[r]) * Ktot + b_kq[r] * LV::kElems;
  }

  typename LV::V a_reg[A_LOADS], b_reg[B_LOADS];
  auto gload = [&](int s) {
    int kk0 = s * BK;
    int j = kk0 / cin, c0 = kk0 - j * cin;
    long long aoff = (long long)j * dil * x_row_stride + c0;
#pragma unroll
    for (int r = 0; r < A_LOADS; ++r) {
      if (a_ok[r]) a_reg[r] = *reinterpret_cast<const typename LV::V*>(X + a_base[r] + aoff);
      else a_reg[r] = typename LV::V{};
    }
#pragma unroll
    for (int r = 0; r < B_LOADS; ++r) {
      if (b_ok[r]) b_reg[r] = *reinterpret_cast<const typename LV::V*>(W + b_base[r] + kk0);
      else b_reg[r] = typename LV::V{};
    }
  };
  auto sstore = [&](int buf) {
    float f[LV::kElems];
#pragma unroll
    for (int r = 0; r < A_LOADS; ++r) {
      if (a_row[r] < BM) {
        LV::unpack(a_reg[r], f);
#pragma unroll
        for (int q = 0; q < LV::kElems; ++q) As[buf][a_kq[r] * LV::kElems + q][a_row[r]] = f[q];
      }
    }
#pragma unroll
    for (int r = 0; r < B_LOADS; ++r) {
      if (b_row[r] < BN) {
        LV::unpack(b_reg[r], f);
#pragma unroll
        for (int q = 0; q < LV::kElems; ++q) Bs[buf][b_kq[r] * LV::kElems + q][b_row[r]] = f[q];
      }
    }
  };

  const int tx = tid % 16, ty = tid / 16;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  for (int s = 0; s < nslices; ++s) {
    int buf = s & 1;
    if (s + 1 < nslices) gload(s + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (s + 1 < nslices) sstore(buf ^ 1);
    __syncthreads();
  }

  // ---- fused epilogue.  The activation is applied tile-wide first (one switch per thread, not per
  // element), then the residual / mask / output chain.
  switch (e.act) {
    case ACT_RELU:
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { int n = n0 + tx * 4 + j; float b = (e.bias && n < cout) ? e.bias[n] : 0.f; acc[i][j] = fmaxf((acc[i][j] + b) * e.scale, 0.f); }
      break;
    case ACT_LRELU:
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { int n = n0 + tx * 4 + j; float b = (e.bias && n < cout) ? e.bias[n] : 0.f; float v = (acc[i][j] + b) * e.scale; acc[i][j] = v > 0.f ? v : v * e.slope; }
      break;
    case ACT_GELU:
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { int n = n0 + tx * 4 + j; float b = (e.bias && n < cout) ? e.bias[n] : 0.f; float v = (acc[i][j] + b) * e.scale; acc[i][j] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
      break;
    case ACT_TANH:
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { int n = n0 + tx * 4 + j; float b = (e.bias && n < cout) ? e.bias[n] : 0.f; acc[i][j] = tanhf((acc[i][j] + b) * e.scale); }
      break;
    default:
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { int n = n0 + tx * 4 + j; float b = (e.bias && n < cout) ? e.bias[n] : 0.f; acc[i][j] = (acc[i][j] + b) * e.scale; }
  }
  const float s2 = e.act2 == ACT_NONE ? 1.f : (e.act2 == ACT_RELU ? 0.f : e.slope2);   // second output: none / relu / leaky
  const int nb = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    int si = m / L, t = m - si * L;
    int slot = slot_ids ? slot_ids[si] : si;
    float rm = e.rowmask ? e.rowmask[(long long)slot * e.mask_slot_stride + t] : 1.f;
    const float fsc = rm * e.out_scale;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = nb + j;
      if (n >= cout) continue;
      float v = acc[i][j];
      if (e.res) {
        const long long ro = (long long)slot * e.res_slot_stride + (long long)t * e.res_row_stride + n;
        float r = e.res_is_half ? __half2float(reinterpret_cast<const __half*>(e.res)[ro]) : reinterpret_cast<const float*>(e.res)[ro];
        if (e.res_inv_slope != 0.f && r < 0.f) r *= e.res_inv_slope;        // undo the producer's LeakyReLU
        v += r;
      }
      if (e.res2) {
        const long long ro = (long long)slot * e.res2_slot_stride + (long long)t * e.res2_row_stride + n;
        v += e.res2_is_half ? __half2float(reinterpret_cast<const __half*>(e.res2)[ro]) : reinterpret_cast<const float*>(e.res2)[ro];
      }
      v *= fsc;
      if (e.y) {
        const long long yo = (long long)slot * e.y_slot_stride + (long long)(e.y_row0 + t) * e.y_row_stride + n;
        if (e.y_is_half) {
          reinterpret_cast<__half*>(e.y)[yo] = __float2half_rn(v);
        } else {
          float* yp = reinterpret_cast<float*>(e.y) + yo;
          if (e.accumulate) v += *yp;
          *yp = v;
        }
      }
      if (e.y2) {
        float v2 = v > 0.f ? v : v * s2;
        long long o = (long long)slot * e.y2_slot_stride + (long long)(e.y2_row0 + t) * e.y2_row_stride + n;
        if (e.y2_is_half) reinterpret_cast<__half*>(e.y2)[o] = __float2half_rn(v2);
        else reinterpret_cast<float*>(e.y2)[o] = v2;
      }
    }
  }
}

}  // namespace

int launch_conv_gemm_ffma(const conan_conv_params_t& p, cudaStream_t st) {
  if (p.cin % BK != 0) { set_error("conv_gemm_ffma: cin must be a multiple of 16"); return 1; }
  if (p.x_row_stride % 8 != 0 || p.x_slot_stride % 8 != 0) { set_error("conv_gemm_ffma: x strides must be multiples of 8 elements"); return 1; }
  if (p.act2 > ACT_LRELU) { set_error("conv_gemm: act2 must be none, relu or leaky"); return 1; }
  if (p.n_streams <= 0) return 0;
  EpiArgs e{p.bias, p.scale, p.act, p.slope, p.res, p.res_slot_stride, p.res_row_stride, p.rowmask,
            p.mask_slot_stride, p.out_scale, p.y, p.y_slot_stride, p.y_row_stride, p.y_row0, p.accumulate,
            p.y2, p.y2_slot_stride, p.y2_row_stride, p.y2_row0, p.y2_is_half, p.act2, p.slope2, p.res_is_half, p.res_inv_slope,
            p.res2, p.res2_slot_stride, p.res2_row_stride, p.res2_is_half, p.y_is_half};
  long long M = (long long)p.n_streams * p.L;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((p.cout + BN - 1) / BN));
  if (p.x_is_half)
    conv_gemm_ffma_kernel<__half><<<grid, NT, 0, st>>>((const __half*)p.x, p.x_slot_stride, p.x_row_stride, p.row0, p.L,
                                                       p.cin, p.k, p.dil, p.cout, (const __half*)p.w, p.n_streams,
                                                       p.slot_ids, e);
  else
    conv_gemm_ffma_kernel<float><<<grid, NT, 0, st>>>((const float*)p.x, p.x_slot_stride, p.x_row_stride, p.row0, p.L,
                                                      p.cin, p.k, p.dil, p.cout, (const float*)p.w, p.n_streams,
                                                      p.slot_ids, e);
  CONAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace conan
