// Helpers for CTA-local tcgen05 matmuls on shared-memory tiles written by ordinary threads (attention kernels).
//
// A "tile" is R rows x 128 bytes (64 bf16) with the 128-byte swizzle the UMMA descriptors expect (the layout TMA's
// SWIZZLE_128B would produce): 16-byte chunk c of row r lives at  r*128 + ((c ^ (r & 7)) << 4).  Base 1024-byte aligned.
// The same physical tile is read by the tensor core either K-major (row = M/N index, the 64 elements = K) or MN-major
// (row = K index, the 64 elements = M/N), selected in the instruction descriptor.  Wider matrices are several tiles
// ("atoms") side by side, `atom_stride` bytes apart.
#pragma once
#include "ptx.cuh"

namespace mtp {

__device__ __forceinline__ uint32_t tile_chunk_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

// D[tmem] (+)= A * B^T-style product with M x N x K (K multiple of 16) from swizzled smem tiles.
//   K-major operand : element (i, k) in atom k/64, row i;           atom_stride = bytes between 64-wide K atoms
//   MN-major operand: element (i, k) in atom i/64, row k (K index);  atom_stride = bytes between 64-wide MN atoms
// One thread calls this.
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void tc_mma_tiles(uint32_t d_tmem, uint32_t a_base, uint32_t a_atom_stride, uint32_t b_base,
                                             uint32_t b_atom_stride, int M, int N, int K, bool accumulate) {
  const uint32_t idesc = make_idesc_bf16(M, N, A_MN, B_MN);
  for (int k = 0; k < K / 16; ++k) {
    uint64_t ad, bd;
    if (A_MN) ad = make_smem_desc(a_base + k * 2048, a_atom_stride, 1024);
    else ad = make_smem_desc(a_base + (k >> 2) * a_atom_stride + (k & 3) * 32, 16, 1024);
    if (B_MN) bd = make_smem_desc(b_base + k * 2048, b_atom_stride, 1024);
    else bd = make_smem_desc(b_base + (k >> 2) * b_atom_stride + (k & 3) * 32, 16, 1024);
    umma_bf16(d_tmem, ad, bd, idesc, accumulate || k > 0);
  }
}


// fp32 [rows][64] tables in shared memory read as float4 by lanes that index DIFFERENT rows (rel-pos tables: the row depends on the lane's
// token): with a 256-byte row pitch every row starts in bank 0 and a quarter warp serialises.  16-byte chunk c of row r is stored at chunk
// c ^ (r & 7) instead (no padding: the RVSA backward has 104 bytes of shared memory to spare).  Returns the float offset of chunk c of row r.
__device__ __forceinline__ int tab_chunk_off(int r, int c) { return r * 64 + ((c ^ (r & 7)) << 2); }

// zero a region of shared memory cooperatively (bytes multiple of 16)
__device__ __forceinline__ void smem_zero(void* p, int bytes, int tid, int nthreads) {
  uint4* q = reinterpret_cast<uint4*>(p);
  for (int i = tid; i < bytes / 16; i += nthreads) q[i] = make_uint4(0, 0, 0, 0);
}

}  // namespace mtp
