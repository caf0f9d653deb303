// Micro-benchmark: L2 -> shared-memory throughput of TMA tile loads per SM (all SMs loading at once), for the box shapes
// the conv kernels use.  A ring of STAGES 16 KB stages per CTA; a consumer warp releases each stage as soon as it lands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I conan_b200/csrc tools/bench/tma_rate.cu -o tools/bench/tma_rate -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace conan;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int STAGES>
__global__ void __launch_bounds__(64) tma_kernel(const __grid_constant__ CUtensorMap tm, int iters, int rows_per_box, int box_bytes,
                                                 int same_addr, int n_boxes_total, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&full[s], (uint32_t)box_bytes);
        const int box = same_addr ? (it % 64) : (int)(((long long)blockIdx.x * 977 + it) % n_boxes_total);
        tma_load_2d(smem + s * 16384, &tm, &full[s], 0, box * rows_per_box);
      }
    }
  } else {
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (it / STAGES) & 1);
      if (elect_one_sync()) mbar_arrive(&empty[s]);
    }
    long long t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) *out = t1 - t0;
  }
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const size_t total_rows = 1 << 19;                      // 2^19 rows x 128 B = 64 MiB (L2-resident after the first pass)
  __half* buf; cudaMalloc(&buf, total_rows * 128); cudaMemset(buf, 0, total_rows * 128);
  long long* d; cudaMalloc(&d, 8);
  struct Cfg { const char* name; int inner_halfs; int rows; CUtensorMapSwizzle swz; };
  Cfg cfgs[] = {{"box {64 halfs, 128 rows} SW128 (ring A/B tile)", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B},
                {"box {64 halfs, 64 rows} SW128 (one C=64 tap)", 64, 64, CU_TENSOR_MAP_SWIZZLE_128B},
                {"box {32 halfs, 256 rows} SW64", 32, 256, CU_TENSOR_MAP_SWIZZLE_64B},
                {"box {32 halfs, 32 rows} SW64 (one C=32 tap)", 32, 32, CU_TENSOR_MAP_SWIZZLE_64B}};
  for (auto& c : cfgs) {
    const size_t rowb = (size_t)c.inner_halfs * 2;
    const size_t nrows = total_rows * 128 / rowb;
    cuuint64_t dims[2] = {(cuuint64_t)c.inner_halfs, (cuuint64_t)nrows};
    cuuint64_t strides[1] = {(cuuint64_t)rowb};
    cuuint32_t box[2] = {(cuuint32_t)c.inner_halfs, (cuuint32_t)c.rows};
    cuuint32_t es[2] = {1, 1};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    const int box_bytes = (int)(rowb * c.rows);
    const int n_boxes = (int)(nrows / c.rows);
    for (int same : {0, 1})
      for (int ctas_per_sm : {1, 2}) {
        auto k = tma_kernel<8>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 1024);
        const int iters = 4000, grid = 148 * ctas_per_sm;
        for (int rep = 0; rep < 2; ++rep) k<<<grid, 64, 8 * 16384 + 1024>>>(tm, iters, c.rows, box_bytes, same, n_boxes, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%-48s %s  %d CTA/SM x 8 stages: %6.1f B/clk/CTA  %6.1f B/clk/SM  (%5.0f clk per box of %5d B) %s\n", c.name,
               same ? "same lines on all SMs" : "distinct lines       ", ctas_per_sm, (double)box_bytes * iters / h,
               (double)box_bytes * iters / h * ctas_per_sm, (double)h / iters, box_bytes, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  }
  return 0;
}
