// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, M = 128, fp32 accumulate, SS mode) on one SM as a function
// of N, of the swizzle span, and of the A start-row offset inside the swizzle atom (the "taps are row offsets" trick
// of the conv kernels).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I conan_b200/csrc tools/bench/mma_rate.cu -o /tmp/mma_rate && /tmp/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace conan;

template <int N, int ROWB, int NACC, int CE, int BG>
__global__ void __launch_bounds__(128) rate_kernel(int iters, int a_row_step, const uint8_t* gsrc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2, bgbar[4];
  __shared__ volatile int stop_flag;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { stop_flag = 0; mbar_init(&bar, 1); mbar_init(&bar2, 1u << 20); for (int i = 0; i < 4; ++i) mbar_init(&bgbar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  constexpr int KK = ROWB / 32, TAPS = 8;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc<N>();
    const uint32_t s32 = smem_u32(smem);
    const uint64_t adesc0 = make_smem_desc<ROWB>(s32), bdesc0 = make_smem_desc<ROWB>(s32 + 48 * 1024);
    const uint64_t step = (uint64_t)((a_row_step * ROWB) >> 4);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
        for (int kk = 0; kk < KK; ++kk)
          if (elect_one_sync())
            tc_mma_f16(tmem_base + (uint32_t)((tap % NACC) * N), adesc0 + tap * step + (uint64_t)(kk * 2), bdesc0 + (uint64_t)(tap * 64 + kk * 2), idesc, 1u);
        if (CE > 0 && (tap + 1) % (CE > 0 ? CE : 1) == 0 && elect_one_sync()) tc_commit(&bar2);   // per-stage "smem slot free" commit of a real pipeline
      }
    }
    if (elect_one_sync()) tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
    stop_flag = 1;
  } else if (BG > 0 && warp == 1) {
    // background: a ring of 4 x BG-KB bulk copies global(L2) -> shared, as a weight / activation producer would run
    uint8_t* dst = smem + 64 * 1024;
    int it = 0;
    while (!stop_flag) {
      const int s = it & 3;
      if (it >= 4) mbar_wait(&bgbar[s], ((it >> 2) - 1) & 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&bgbar[s], BG * 1024);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst + s * BG * 1024)), "l"(gsrc + (size_t)((blockIdx.x * 64 + (it & 63)) * BG * 1024)), "r"(BG * 1024), "r"(smem_u32(&bgbar[s])) : "memory");
      }
      ++it;
    }
    for (int k2 = (it > 4 ? it - 4 : 0); k2 < it; ++k2) mbar_wait(&bgbar[k2 & 3], (k2 >> 2) & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

template <int N, int ROWB, int NACC = 1, int CE = 0, int BG = 0>
void run(const char* name, int a_row_step, int grid) {
  static uint8_t* gsrc = nullptr; if (!gsrc) { cudaMalloc(&gsrc, (size_t)296 * 64 * 8 * 1024); cudaMemset(gsrc, 0, (size_t)296 * 64 * 8 * 1024); }
  const int commit_every = CE;
  long long* d; cudaMalloc(&d, 8);
  auto k = rate_kernel<N, ROWB, NACC, CE, BG>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 500, per_it = 8 * (ROWB / 32);
  k<<<grid, 128, 100 * 1024>>>(iters, a_row_step, gsrc, d);
  k<<<grid, 128, 100 * 1024>>>(iters, a_row_step, gsrc, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  if (commit_every) printf("[commit every %d taps] ", commit_every);
  if (BG) printf("[background %d KB bulk copies] ", BG);
  printf("%-34s N=%3d span=%3dB a_row_step=%2d grid=%3d accs=%d : %7.1f cycles / MMA (floor N/2 = %d)  %s\n", name, N, ROWB, a_row_step, grid, NACC,
         (double)h / (iters * per_it), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

// Does a tcgen05.commit fire when ITS batch of MMAs is complete, or only once later-issued MMAs have drained too?
// warp 0: batch A (NA MMAs -> accumulator 0), commit(barA), wait `gap` cycles, batch B (NB MMAs -> accumulator 1), commit(barB).
// warp 1 polls barA then barB and records when it saw them.
template <int N, int ROWB>
__global__ void __launch_bounds__(64) commit_kernel(int NA, int NB, int gap, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t barA, barB;
  __shared__ uint32_t tmem_ptr;
  __shared__ long long t_issueA, t_issueB;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&barA, 1); mbar_init(&barB, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  constexpr uint32_t idesc = make_idesc<N>();
  const uint32_t s32 = smem_u32(smem);
  const uint64_t ad = make_smem_desc<ROWB>(s32), bd = make_smem_desc<ROWB>(s32 + 32 * 1024);
  if (warp == 0) {
    for (int i = 0; i < NA; ++i) if (elect_one_sync()) tc_mma_f16(tmem_base, ad + (uint64_t)((i & 3) * 2), bd + (uint64_t)((i & 3) * 2), idesc, 1u);
    if (elect_one_sync()) tc_commit(&barA);
    const long long t1 = clock64();
    if (threadIdx.x == 0) t_issueA = t1;
    while (clock64() - t1 < gap) {}
    for (int i = 0; i < NB; ++i) if (elect_one_sync()) tc_mma_f16(tmem_base + N, ad + (uint64_t)((i & 3) * 2), bd + (uint64_t)((i & 3) * 2), idesc, 1u);
    if (elect_one_sync()) tc_commit(&barB);
    if (threadIdx.x == 0) t_issueB = clock64();
  } else {
    mbar_wait(&barA, 0);
    const long long ta = clock64();
    mbar_wait(&barB, 0);
    const long long tb = clock64();
    __syncwarp();
    if (threadIdx.x == 32) { out[0] = ta; out[1] = tb; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) { out[2] = t_issueA; out[3] = t_issueB; }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

void run_commit(int NA, int NB, int gap) {
  long long* d; cudaMalloc(&d, 32);
  auto k = commit_kernel<64, 128>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 64, 80 * 1024>>>(NA, NB, gap, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("commit test N=64: A=%2d MMAs, gap %4d cycles, B=%2d MMAs : barA seen %5lld cycles after A's commit was issued; barB seen %5lld after B's commit issued %s\n",
         NA, gap, NB, h[0] - h[2], h[1] - h[3], e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<32, 64>("C=32 window conv", 0, grid);   run<32, 64>("C=32 window conv, taps", 1, grid);  run<32, 64>("C=32 window conv, taps dil 5", 5, grid);
    run<64, 128>("C=64 window conv", 0, grid);  run<64, 128>("C=64 window conv, taps", 1, grid); run<64, 128>("C=64 window conv, taps dil 5", 5, grid);
    run<64, 128>("C=64, atom-aligned taps", 8, grid);
    run<128, 128>("ring 128x128", 0, grid);     run<128, 128>("ring 128x128, taps", 1, grid);
    run<256, 128>("N=256", 0, grid);
  }
  // independent accumulators: is the minimum a dependent-accumulate latency or an operand-read throughput limit?
  run<32, 64, 2>("C=32, 2 accumulators", 1, 148);  run<32, 64, 4>("C=32, 4 accumulators", 1, 148);
  run<32, 128, 1>("N=32 on 128B rows", 1, 148);    run<32, 128, 2>("N=32 on 128B rows", 1, 148);
  run<64, 128, 2>("C=64, 2 accumulators", 1, 148); run<64, 128, 4>("C=64, 4 accumulators", 1, 148);
  run<128, 128, 2>("N=128, 2 accumulators", 1, 148);
  run<16, 128, 1>("N=16 on 128B rows", 1, 148);
  // a tcgen05.commit after every tap / every second tap (what a real smem pipeline does)
  run<64, 128, 1, 1>("C=64", 1, 148); run<64, 128, 1, 2>("C=64", 1, 148); run<32, 64, 1, 1>("C=32", 1, 148); run<32, 64, 1, 4>("C=32", 1, 148);
  run<128, 128, 1, 1>("N=128", 0, 148);
  // concurrent bulk copies into shared memory (the weight / activation producers of a real kernel)
  run<64, 128, 1, 2, 8>("C=64", 1, 148); run<32, 64, 1, 4, 2>("C=32", 1, 148); run<32, 64, 1, 4, 8>("C=32", 1, 148); run<128, 128, 1, 1, 8>("N=128", 0, 148);
  // two CTAs per SM, each with its own accumulator
  run<32, 64>("C=32, 2 CTAs/SM", 1, 296); run<64, 128>("C=64, 2 CTAs/SM", 1, 296); run<128, 128>("N=128, 2 CTAs/SM", 1, 296);
  run_commit(12, 0, 0); run_commit(12, 12, 0); run_commit(12, 12, 700); run_commit(12, 44, 700); run_commit(44, 0, 0); run_commit(44, 44, 0);
  return 0;
}
