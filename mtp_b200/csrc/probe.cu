// Measurement aid (tools/turnaround_probe.py): a do-nothing kernel with a configurable resource footprint -- dynamic shared memory,
// TMEM allocation, a busy-wait -- stamped with the global timer, to find out what makes the SM turnaround between two dependent
// heavy launches (5-7 us measured between tcgen05 GEMM launches vs 0.8 us between trivial kernels).
#include "common.h"
#include "ptx.cuh"

namespace mtp {

__global__ void __launch_bounds__(320, 1)
probe_kernel(long long* stamps, int tmem_cols, int spin_ns, int pdl_early) {
  extern __shared__ uint8_t sm[];
  __shared__ uint32_t slot;
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (pdl_early) MTP_PDL_ENTRY();
  if (tmem_cols > 0 && threadIdx.x < 32) tmem_alloc(&slot, tmem_cols);
  if (threadIdx.x == 0) sm[0] = 1;          // touch the dynamic allocation
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (!pdl_early) MTP_PDL_ENTRY();
  long long t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  long long t = t1;
  while (t - t1 < spin_ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  __syncthreads();
  if (tmem_cols > 0 && threadIdx.x < 32) tmem_dealloc(slot, tmem_cols);
  if (threadIdx.x == 0) {
    long long t2;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
    stamps[blockIdx.x * 4 + 0] = t0;
    stamps[blockIdx.x * 4 + 1] = t1;
    stamps[blockIdx.x * 4 + 2] = t2;
  }
}

}  // namespace mtp

extern "C" int mtp_probe_launch(long long* stamps, int grid, int threads, int smem_bytes, int tmem_cols, int spin_ns, int pdl_early,
                                mtp_stream_t stream) {
  MTP_REQUIRE(stamps && grid > 0 && threads >= 32 && threads <= 320 && smem_bytes >= 16, "mtp_probe_launch: bad args");
  static int attr = 0;
  if (smem_bytes > attr) {
    cudaError_t e = cudaFuncSetAttribute(mtp::probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return mtp::set_error(MTP_ERR_CUDA, "probe smem attr: %s", cudaGetErrorString(e));
    attr = smem_bytes;
  }
  (void)mtp::launch_k(mtp::probe_kernel, grid, threads, smem_bytes, reinterpret_cast<cudaStream_t>(stream), stamps, tmem_cols, spin_ns, pdl_early);
  return mtp::check_launch("probe_kernel");
}
