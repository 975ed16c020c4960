// Shared host/device helpers for the r2l_b200 C-ABI library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace r2l {

// error codes returned across the C ABI (0 = ok)
enum : int {
  R2L_OK = 0,
  R2L_ERR_INVALID = 1,   // bad argument (null pointer, size, unsupported configuration)
  R2L_ERR_CUDA = 2,      // a CUDA runtime call / kernel launch failed
  R2L_ERR_UNSUPPORTED = 3,
  R2L_ERR_DEVICE_TRAP = 4,  // kernel watchdog fired (see DebugBuf)
  R2L_ERR_RANGE = 5,        // a weight or an activation left the range of the 16-bit operand type (fp16: 65504)
};

void set_last_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

#define R2L_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return ::r2l::fail(::r2l::R2L_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define R2L_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::r2l::fail(::r2l::R2L_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                                       \
  } while (0)

// every kernel launch of the library is followed by one of these: it also counts the launch (r2l_kernel_launches)
#define R2L_LAUNCH_CHECK()                                                                          \
  do {                                                                                              \
    ::r2l::count_launch();                                                                          \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess)                                                                          \
      return ::r2l::fail(::r2l::R2L_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                   \
                         cudaGetErrorString(_e), __FILE__, __LINE__);                               \
  } while (0)

static inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

int sm_count();
void count_launch();

}  // namespace r2l
