// Internal host-side helpers shared by the .cu translation units of libmtp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

#include "../../include/mtp_b200.h"

namespace mtp {

int set_error(int code, const char* fmt, ...);   // records the message for mtp_last_error(), returns code

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return MTP_OK;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

int num_sms();

}  // namespace mtp

#define MTP_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return mtp::set_error(MTP_ERR_INVALID, __VA_ARGS__); \
  } while (0)
