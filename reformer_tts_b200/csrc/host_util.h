// Host-side helpers shared by the C-ABI entry points: error reporting and argument checks.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>

namespace rtts {

enum : int { kOk = 0, kErrBadArg = -1, kErrCuda = -2, kErrUnsupported = -3 };

char* error_buffer();  // thread-local, 512 bytes (api.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(kErrCuda, "%s: %s", what, cudaGetErrorString(e));
  return kOk;
}

#define RTTS_REQUIRE(cond, ...) \
  do {                          \
    if (!(cond)) return ::rtts::fail(::rtts::kErrBadArg, __VA_ARGS__); \
  } while (0)

}  // namespace rtts
