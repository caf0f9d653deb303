// Shared device/host helpers for the conan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <mutex>
#include <string>

#include "../../include/conan_b200.h"

namespace conan {

void set_error(const std::string& msg);
void count_launch(int n = 1);

#define CONAN_CUDA_OK(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      conan::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));          \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

#define CONAN_CHECK_LAUNCH()                                                         \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      conan::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) + \
                       " at " + __FILE__ + ":" + std::to_string(__LINE__));          \
      return 1;                                                                      \
    }                                                                                \
    conan::count_launch();                                                           \
  } while (0)

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_GELU = 3, ACT_TANH = 4 };

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_LRELU: return v > 0.f ? v : v * slope;
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));  // exact-erf GELU (nn.GELU default)
    case ACT_TANH: return tanhf(v);
    default: return v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Per-device one-time setup.  cudaFuncSetAttribute (the > 48 KB dynamic shared memory opt-in, the carve-out preference) and the
// occupancy figures derived from it apply to the CURRENT device only, so launchers cache them per device ordinal, under a mutex
// (two host threads may drive engines on different GPUs of one process).  `setup(int* value)` returns 0 on success.
struct DeviceOnce {
  static constexpr int kMaxDevices = 64;
  std::mutex mu;
  bool done[kMaxDevices] = {};
  int value[kMaxDevices] = {};
};
template <typename F>
inline int device_once(DeviceOnce& o, int* value, F&& setup) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= DeviceOnce::kMaxDevices) { set_error("device_once: bad current device"); return 1; }
  std::lock_guard<std::mutex> lk(o.mu);
  if (!o.done[dev]) {
    int v = 0;
    if (setup(&v)) return 1;
    o.value[dev] = v;
    o.done[dev] = true;
  }
  if (value) *value = o.value[dev];
  return 0;
}

// launchers implemented in the .cu files
int launch_conv_gemm_ffma(const conan_conv_params_t& p, cudaStream_t st);
int launch_conv_gemm_tc(const conan_conv_params_t& p, cudaStream_t st);
bool conv_gemm_tc_eligible(const conan_conv_params_t& p);
bool conv_gemm_tc_uses_window(const conan_conv_params_t& p);
// G <= 3 independent convs of one shape as one launch of the CTA-pair kernel; -1 = the group does not qualify (nothing launched)
int launch_conv_gemm_tc_group(const conan_conv_params_t* ps, int G, cudaStream_t st, bool sum = false);

}  // namespace conan
