// Measurement aid (tools/turnaround_probe.py): a do-nothing kernel with a configurable resource footprint -- dynamic shared memory,
// TMEM allocation, a busy-wait, end-of-kernel accumulator reads / global reads / stores -- stamped with the global timer: how long after one
// launch's last CTA is the next launch ready?  (0.9-1.0 us whatever the footprint and the end work: the "5-7 us after a GEMM" this was built to
// explain turned out to be the GEMM's own exposed last-tile epilogue)
#include "common.h"
#include "ptx.cuh"

namespace mtp {

__global__ void __launch_bounds__(320, 1)
probe_kernel(long long* stamps, int tmem_cols, int spin_ns, int pdl_early, uint8_t* buf, int store_bytes, int store_pattern, int ldo,
             int n_tmem_ld, int n_loads) {
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
  // optional end-of-kernel work of a GEMM epilogue: accumulator reads, then the tile's output stores in one of three access patterns
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint32_t keep = 0;
  if (n_loads > 0 && buf != nullptr) {           // n_loads x 512 B per warp of (L2-resident) global reads: the aux operand of an epilogue
    const uint4* src = reinterpret_cast<const uint4*>(buf) + ((size_t)blockIdx.x * nw + warp) * n_loads * 32 + lane;
    for (int j = 0; j < n_loads; ++j) { const uint4 u = src[j * 32]; keep ^= u.x ^ u.w; }
  }
  if (n_tmem_ld > 0 && tmem_cols >= 32) {
    const uint32_t taddr = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[32];
    for (int j = 0; j < n_tmem_ld; ++j) {
      tmem_ld_32x32(taddr + (j * 32) % tmem_cols, r);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) keep ^= r[k];
    }
  }
  if (store_bytes > 0 && buf != nullptr) {
    const int n_instr = store_bytes / 512 / nw;           // 512 B per warp instruction
    const int row_bytes = store_bytes / 128;              // the CTA's tile: 128 rows x row_bytes
    const int tiles_n = ldo / row_bytes;
    uint8_t* tile = buf + (size_t)(blockIdx.x / tiles_n) * 128 * ldo + (size_t)(blockIdx.x % tiles_n) * row_bytes;
    const uint4 val = make_uint4(keep, threadIdx.x, blockIdx.x, 1u);
    for (int j = 0; j < n_instr; ++j) {
      uint8_t* a;
      if (store_pattern == 0) {                           // fully coalesced: 512 contiguous bytes per instruction
        a = buf + ((size_t)blockIdx.x * nw + warp) * n_instr * 512 + (size_t)j * 512 + lane * 16;
      } else if (store_pattern == 1) {                    // 8 rows x 64 B per instruction (the GEMM epilogue after its 4-lane transpose)
        const int jj = warp * n_instr + j;                // 16 instructions cover 128 rows of one 64-byte column chunk
        a = tile + (size_t)((jj % 16) * 8 + (lane >> 2)) * ldo + (jj / 16) * 64 + (lane & 3) * 16;
      } else if (store_pattern == 2) {                    // 32 rows x 16 B per instruction (row per lane)
        const int jj = warp * n_instr + j;                // 4 instructions cover 128 rows of one 16-byte column chunk
        a = tile + (size_t)((jj % 4) * 32 + lane) * ldo + (jj / 4) * 16;
      } else {                                            // 2 rows x 256 B per instruction
        const int jj = warp * n_instr + j;                // 64 instructions cover 128 rows of one 256-byte column chunk
        a = tile + (size_t)((jj % 64) * 2 + (lane >> 4)) * ldo + (jj / 64) * 256 + (lane & 15) * 16;
      }
      *reinterpret_cast<uint4*>(a) = val;
    }
  } else if (keep == 0x9e3779b9u && stamps != nullptr) {
    stamps[blockIdx.x * 4 + 3] = keep;
  }
  tc_fence_before();
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

extern "C" int mtp_probe_launch2(long long* stamps, int grid, int threads, int smem_bytes, int tmem_cols, int spin_ns, int pdl_early, void* buf,
                                 int store_bytes, int store_pattern, int ldo, int n_tmem_ld, int n_loads, mtp_stream_t stream) {
  MTP_REQUIRE(stamps && grid > 0 && threads >= 32 && threads <= 320 && threads % 32 == 0 && smem_bytes >= 16, "mtp_probe_launch: bad args");
  MTP_REQUIRE(store_bytes == 0 || (buf != nullptr && ldo > 0 && store_bytes % (512 * (threads / 32)) == 0 && ldo % (store_bytes / 128) == 0),
              "mtp_probe_launch: store_bytes must be a multiple of 512 B per warp and tile the row pitch");
  static int attr = 0;
  if (smem_bytes > attr) {
    cudaError_t e = cudaFuncSetAttribute(mtp::probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return mtp::set_error(MTP_ERR_CUDA, "probe smem attr: %s", cudaGetErrorString(e));
    attr = smem_bytes;
  }
  (void)mtp::launch_k(mtp::probe_kernel, grid, threads, smem_bytes, reinterpret_cast<cudaStream_t>(stream), stamps, tmem_cols, spin_ns, pdl_early,
                      reinterpret_cast<uint8_t*>(buf), store_bytes, store_pattern, ldo, n_tmem_ld, n_loads);
  return mtp::check_launch("probe_kernel");
}

extern "C" int mtp_probe_launch(long long* stamps, int grid, int threads, int smem_bytes, int tmem_cols, int spin_ns, int pdl_early,
                                mtp_stream_t stream) {
  return mtp_probe_launch2(stamps, grid, threads, smem_bytes, tmem_cols, spin_ns, pdl_early, nullptr, 0, 0, 0, 0, 0, stream);
}
