// Dense attention with decomposed rel-pos bias on tcgen05 tensor cores for LONG sequences (N = gh*gw > 256: 512^2 and 1024^2 inputs),
// forward.  Flash-style: K / V stream through shared memory in blocks of 128 keys (cp.async, double-buffered), the N x N scores never
// leave the SM.                                                           [V]:90-111, 142-193; SURVEY.md K6
//
// One CTA (128 threads) per (image, head, 128-query tile); thread r owns query row r.
//   prologue : the rel-pos tables of the head go through shared memory once; thread r builds its row's factorised bias
//              rh[r][jy] = q.Rh[qy - jy + gh - 1], rw[r][jx] = q.Rw[qx - jx + gw - 1] (fp32, N*(gh+gw)*64 MACs instead of an N x N table)
//   per block: S = Q K_blk^T (UMMA 128x128x64 -> TMEM) -> row r: running max / sum, P = exp(scale (s - m)) as bf16 into a swizzled tile
//              -> PV = P V_blk (UMMA 128x64x128 -> TMEM) -> o[r][:] = o * exp(scale (m_old - m)) + PV in registers
//   epilogue : o / l -> bf16, LSE saved for the backward.
#include "common.h"
#include "ptx.cuh"
#include "tc_tile.cuh"

namespace mtp {

constexpr int FS_THREADS = 128;
constexpr int FS_TILE = 128 * 128;       // bytes of a 128-row x 64-bf16 tile
constexpr int FS_TLD = 65;               // padded row stride (floats) of the rel-pos tables in smem: conflict-free for per-thread rows

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;           // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

__host__ __device__ inline int fs_ldr(int g) { return g | 1; }      // odd row stride of the per-row rel terms
__host__ inline int fs_smem_bytes(int gh, int gw) {
  const int rel = 128 * (fs_ldr(gh) + fs_ldr(gw)) * 4;
  return FS_TILE /*Q*/ + 4 * FS_TILE /*K0 K1 V0 V1*/ + 2 * FS_TILE /*P*/ + rel + 64;
}

__global__ void __launch_bounds__(FS_THREADS)
full_attn_fwd_stream_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                               __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N, int gh, int gw, int C, int nH, int use_rel) {
  MTP_PDL_ENTRY();
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* Qs = sm;
  uint8_t* KV = Qs + FS_TILE;              // K0 | K1 | V0 | V1
  uint8_t* Pt = KV + 4 * FS_TILE;          // 2 atoms (keys 0..63 | 64..127 of the block)
  const int ldh = fs_ldr(gh), ldw = fs_ldr(gw);
  float* rh_s = reinterpret_cast<float*>(Pt + 2 * FS_TILE);      // [128][ldh]
  float* rw_s = rh_s + 128 * ldh;                                 // [128][ldw]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(rw_s + 128 * ldw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  float* tab = reinterpret_cast<float*>(KV);                      // prologue only: rel-pos tables, rows padded to FS_TLD floats

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, n = blockIdx.y, b = blockIdx.z;
  const int C3 = 3 * C;
  const float scale = 0.125f;
  const __nv_bfloat16* base = qkv + (size_t)b * N * C3 + n * 64;

  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (tid == 32) { mbar_init(mbar, 1); fence_barrier_init(); }
  for (int i = tid; i < 128 * 8; i += FS_THREADS) {             // Q tile (rows beyond N are zero)
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q0 + r < N) v = *reinterpret_cast<const uint4*>(base + (size_t)(q0 + r) * C3 + c * 8);
    *reinterpret_cast<uint4*>(Qs + tile_chunk_off(r, c)) = v;
  }
  const int rows_h = 2 * gh - 1, rows_w = 2 * gw - 1;
  if (use_rel) {
    for (int i = tid; i < rows_h * 64; i += FS_THREADS) tab[(i >> 6) * FS_TLD + (i & 63)] = rel_h[i];
    for (int i = tid; i < rows_w * 64; i += FS_THREADS) tab[(rows_h + (i >> 6)) * FS_TLD + (i & 63)] = rel_w[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t T_S = tmem, T_PV = tmem + 128;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const int q = q0 + tid;
  const bool qvalid = q < N;
  float* rh = rh_s + tid * ldh;
  float* rw = rw_s + tid * ldw;
  if (use_rel) {          // this row's factorised bias terms (UNscaled q; the scale multiplies score + bias together below)
    const int qy = qvalid ? q / gw : 0, qx = qvalid ? q % gw : 0;
    float qv[64];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = *reinterpret_cast<const uint4*>(Qs + tile_chunk_off(tid, c));
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) { const float2 f = unpack_bf16x2(w4[t]); qv[c * 8 + 2 * t] = f.x; qv[c * 8 + 2 * t + 1] = f.y; }
    }
    for (int k = 0; k < gh; ++k) {
      const float* th = tab + (qy - k + gh - 1) * FS_TLD;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 64; ++d) s += qv[d] * th[d];
      rh[k] = s;
    }
    for (int k = 0; k < gw; ++k) {
      const float* tw = tab + (rows_h + qx - k + gw - 1) * FS_TLD;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 64; ++d) s += qv[d] * tw[d];
      rw[k] = s;
    }
  }
  __syncthreads();                          // the table region is free: it becomes the K / V ring

  auto load_block = [&](int kb) {           // 128 keys x (K, V) rows of 128 B -> swizzled tiles of buffer kb & 1 (zero rows beyond N)
    uint8_t* Kd = KV + (kb & 1) * FS_TILE;
    uint8_t* Vd = KV + (2 + (kb & 1)) * FS_TILE;
    const int k0 = kb * 128;
    for (int i = tid; i < 128 * 8; i += FS_THREADS) {
      const int r = i >> 3, c = i & 7;
      const bool ok = k0 + r < N;
      const __nv_bfloat16* src = base + (size_t)(ok ? k0 + r : 0) * C3 + c * 8;
      cp_async16(smem_u32(Kd + tile_chunk_off(r, c)), src + C, ok);
      cp_async16(smem_u32(Vd + tile_chunk_off(r, c)), src + 2 * C, ok);
    }
    cp_async_commit();
  };

  const int nkb = (N + 127) / 128;
  float o[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) o[d] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  uint32_t phase = 0;
  load_block(0);
  for (int kb = 0; kb < nkb; ++kb) {
    const int k0 = kb * 128;
    if (kb + 1 < nkb) { load_block(kb + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    fence_proxy_async_smem();
    __syncthreads();
    const uint32_t Kb = smem_u32(KV + (kb & 1) * FS_TILE), Vb = smem_u32(KV + (2 + (kb & 1)) * FS_TILE);
    if (warp == 0) {      // warp-uniform issue (elected lane)
      tc_fence_after();
      if (elect_one()) {
        tc_mma_tiles<false, false>(T_S, smem_u32(Qs), 0, Kb, 0, 128, 128, 64, false);
        umma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- pass 1: block maximum of this row (raw scores + bias; the scale is applied inside the exponent)
    float mx = -INFINITY;
    {
      int jy = k0 / gw, jx = k0 % gw;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(T_S + lane_base + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = k0 + c * 32 + e;
          if (j < N) {
            float s = __uint_as_float(r[e]);
            if (use_rel) s += rh[jy] + rw[jx];
            mx = fmaxf(mx, s);
            if (++jx == gw) { jx = 0; ++jy; }
          }
        }
      }
    }
    const float m_new = fmaxf(m_run, mx);
    const float corr = __expf(scale * (m_run - m_new));          // exp(-inf) = 0 on the first block
    float l_blk = 0.f;
    {
      int jy = k0 / gw, jx = k0 % gw;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(T_S + lane_base + c * 32, r);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = k0 + c * 32 + e;
          float p = 0.f;
          if (j < N) {
            float s = __uint_as_float(r[e]);
            if (use_rel) s += rh[jy] + rw[jx];
            p = qvalid ? __expf(scale * (s - m_new)) : 0.f;
            if (++jx == gw) { jx = 0; ++jy; }
          }
          pv[e] = p;
          l_blk += p;
        }
        uint8_t* atom = Pt + (c >> 1) * FS_TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 u;
          u.x = pack_bf16x2(pv[8 * k], pv[8 * k + 1]); u.y = pack_bf16x2(pv[8 * k + 2], pv[8 * k + 3]);
          u.z = pack_bf16x2(pv[8 * k + 4], pv[8 * k + 5]); u.w = pack_bf16x2(pv[8 * k + 6], pv[8 * k + 7]);
          *reinterpret_cast<uint4*>(atom + tile_chunk_off(tid, (c & 1) * 4 + k)) = u;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        tc_mma_tiles<false, true>(T_PV, smem_u32(Pt), FS_TILE, Vb, 0, 128, 64, 128, false);
        umma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(T_PV + lane_base, r0);
      tmem_ld_32x32(T_PV + lane_base + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        o[d] = o[d] * corr + __uint_as_float(r0[d]);
        o[32 + d] = o[32 + d] * corr + __uint_as_float(r1[d]);
      }
    }
    l_run = l_run * corr + l_blk;
    m_run = m_new;
    tc_fence_before();
    __syncthreads();          // T_S / T_PV, the P tile and this block's K / V buffer may be overwritten
    tc_fence_after();
  }
  if (qvalid) {
    const float inv = 1.0f / l_run;
    __nv_bfloat16* dst = out + ((size_t)b * N + q) * C + n * 64;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 u;
      u.x = pack_bf16x2(o[8 * c] * inv, o[8 * c + 1] * inv); u.y = pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
      u.z = pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv); u.w = pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
      *reinterpret_cast<uint4*>(dst + 8 * c) = u;
    }
    if (lse) lse[((size_t)b * nH + n) * N + q] = scale * m_run + __logf(l_run);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int launch_full_attn_fwd_stream_tc(const void* qkv, const float* rel_h, const float* rel_w, void* out, float* lse, int B, int gh, int gw,
                                   int C, int nH, cudaStream_t st) {
  const int N = gh * gw;
  const int smem = fs_smem_bytes(gh, gw);
  // the prologue parks both rel-pos tables (rows padded to FS_TLD floats) in the K / V ring + P tile region
  MTP_REQUIRE((2 * gh - 1 + 2 * gw - 1) * FS_TLD * 4 <= 6 * FS_TILE && smem <= 227 * 1024,
              "mtp_full_attn_fwd: grid %dx%d too large for the streaming tensor-core kernel", gh, gw);
  static int attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(full_attn_fwd_stream_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "full_attn_fwd_stream_tc smem attr: %s", cudaGetErrorString(e));
    attr = smem;
  }
  (void)launch_k(full_attn_fwd_stream_tc_kernel, dim3(ceil_div(N, 128), nH, B), FS_THREADS, smem, st, reinterpret_cast<const __nv_bfloat16*>(qkv),
                 rel_h, rel_w, reinterpret_cast<__nv_bfloat16*>(out), lse, N, gh, gw, C, nH, rel_h != nullptr);
  return check_launch("full_attn_fwd_stream_tc_kernel");
}

}  // namespace mtp
