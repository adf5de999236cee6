// Shared host/device helpers for libdeepxi_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/deepxi_b200.h"

namespace dxi {

constexpr int N_D = 512;      // window duration (samples)   main.py:33
constexpr int N_S = 256;      // window shift (samples)      main.py:34
constexpr int NFFT = 512;     // K                           main.py:35
constexpr int NBINS = 257;    // K/2 + 1                     inp_tgt.py:156

void set_error(const char* fmt, ...);
int check_device();
extern thread_local int64_t g_launches;

inline int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return DXI_E_CUDA;
}

#define DXI_CUDA(expr)                                           \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return ::dxi::cuda_fail(_e, #expr);   \
  } while (0)

#define DXI_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) { ::dxi::set_error("%s", msg); return DXI_E_INVALID; } \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define DXI_LAUNCHED(name)                                       \
  do {                                                           \
    ++::dxi::g_launches;                                         \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return ::dxi::cuda_fail(_e, name);    \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Optional per-kernel timing (dxi_profile_enable): CUDA events recorded on the launching stream around a
// named group of launches; dxi_profile_read sums the elapsed times.  Off by default (no events recorded).
struct ProfScope {
  ProfScope(const char* key, cudaStream_t st, int launches);
  ~ProfScope();
  void* span_;
  cudaStream_t st_;
  const char* key_ = nullptr;
};

}  // namespace dxi
