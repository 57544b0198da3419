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
int sampling_fused_mask();      // A/B switch (MTP_SAMPLING_FUSED=0 restores the separate pool / GEMV / wgrad launches of the RVSA sampling heads)

// Programmatic dependent launch (mtp_set_pdl): every kernel of this library begins with MTP_PDL_ENTRY() (griddepcontrol.wait, then
// griddepcontrol.launch_dependents), so a kernel launched through launch_k may be scheduled while its stream predecessor drains;
// its blocks set up (barriers, TMEM, descriptor prefetch) and then hold at the wait until the predecessor's memory is visible.
// This removes most of the 3.5-8 us launch gap between dependent kernels (tools/gemm_gaps.py).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace mtp

// first statements of every kernel: wait for the stream predecessor's results, then let the successor start its own set-up
#define MTP_PDL_ENTRY()                                              \
  do {                                                               \
    asm volatile("griddepcontrol.wait;" ::: "memory");               \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  \
  } while (0)

#define MTP_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return mtp::set_error(MTP_ERR_INVALID, __VA_ARGS__); \
  } while (0)
