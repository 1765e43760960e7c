// Shared helpers for libpodsdf (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/podsdf.h"

namespace podsdf {

// error codes (PODSDF_OK, PODSDF_ERR_*) come from include/podsdf.h

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define PODSDF_CUDA(call)                                                      \
  do {                                                                         \
    cudaError_t _e = (call);                                                   \
    if (_e != cudaSuccess) return ::podsdf::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define PODSDF_TRY(call)                 \
  do {                                   \
    int _rc = (call);                    \
    if (_rc != 0) return _rc;            \
  } while (0)

// PODSDF_OK if `device` is a usable sm_100 device, else PODSDF_ERR_NO_DEVICE / _INVALID
int check_device(int device);

// number of kernels launched by this library since load (bench.py's gpu_launches)
void count_launch(int n = 1);

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// RAII device switcher
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Filter taps travel as a kernel parameter (constant bank, static index => DFMA constant operand).
constexpr int kMaxTaps = 129;  // nf <= 64 (config 5 sweeps nf = 4..64)
struct Taps {
  double b[kMaxTaps];
};

}  // namespace podsdf
