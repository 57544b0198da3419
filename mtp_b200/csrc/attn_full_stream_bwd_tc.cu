// Dense attention with decomposed rel-pos bias, BACKWARD, on tcgen05 tensor cores for long sequences (N = gh*gw > 256).
// Autograd of attn_full_stream_tc.cu; [V]:90-111, 142-193.
//
//   prep   (thread = query row)         : D[q] = dO[q].O[q];  rel[q][0:gh] = q.Rh[qy-k+gh-1], rel[q][gh:gh+gw] = q.Rw[qx-k+gw-1]  (fp32)
//   main   (CTA = image, head, 128 KEYS): K / V of the block stay in smem, dK / dV accumulate in TMEM over all query tiles;
//            per query tile:  S = Q K^T, dP = dO V^T (UMMA)  ->  row r: P = exp(scale s - lse), dS = P (dP - D), row sums of dS per key row /
//            key column  ->  dQ_part = dS K, dK += dS^T Q, dV += P^T dO (UMMA)  ->  dQ_part and the dS row sums are added to fp32 scratch
//            with red.global (every key-block CTA contributes to every query row)
//   finish (thread = query row)         : dq = scale (dQ + sum_k dSh[k] Rh[..] + sum_k dSw[k] Rw[..]) -> bf16;  d rel tables = W^T Q as two more UMMAs
// Workspace: D [B nH N] | rel [B nH N (gh+gw)] | dsrow [B nH N (gh+gw)] | dq [B N C]   (fp32; the last two are zeroed by the launcher).
#include "common.h"
#include "ptx.cuh"
#include "tc_tile.cuh"

namespace mtp {

constexpr int SB_THREADS = 128;
constexpr int SBT = 128 * 128;         // bytes of a 128-row x 64-bf16 tile
constexpr int SB_TLD = 65;             // padded row stride of the rel-pos tables in smem

__device__ __forceinline__ void sb_cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void sb_red_add4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sb_load_q64(const uint8_t* tile, int row, float (&qv)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(tile + tile_chunk_off(row, c));
    const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) { const float2 f = unpack_bf16x2(w4[t]); qv[c * 8 + 2 * t] = f.x; qv[c * 8 + 2 * t + 1] = f.y; }
  }
}

// ------------------------------------------------------------------------------------------------------------ prep
__global__ void __launch_bounds__(SB_THREADS)
dense_bwd_prep_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                      const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout, float* __restrict__ Dbuf,
                      float* __restrict__ relbuf, int N, int gh, int gw, int C, int nH, int use_rel) {
  MTP_PDL_ENTRY();
  extern __shared__ float tab[];          // [(2gh-1) + (2gw-1)][SB_TLD]
  const int tid = threadIdx.x, n = blockIdx.y, b = blockIdx.z;
  const int q = blockIdx.x * 128 + tid;
  const int rows_h = 2 * gh - 1, rows_w = 2 * gw - 1;
  if (use_rel) {
    for (int i = tid; i < rows_h * 64; i += SB_THREADS) tab[(i >> 6) * SB_TLD + (i & 63)] = rel_h[i];
    for (int i = tid; i < rows_w * 64; i += SB_THREADS) tab[(rows_h + (i >> 6)) * SB_TLD + (i & 63)] = rel_w[i];
  }
  __syncthreads();
  if (q >= N) return;
  const size_t tok = (size_t)b * N + q;
  float qv[64];
  float dsum = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 qu = *reinterpret_cast<const uint4*>(qkv + tok * 3 * C + n * 64 + c * 8);
    const uint4 ou = *reinterpret_cast<const uint4*>(out + tok * C + n * 64 + c * 8);
    const uint4 gu = *reinterpret_cast<const uint4*>(dout + tok * C + n * 64 + c * 8);
    const uint32_t qw[4] = {qu.x, qu.y, qu.z, qu.w}, ow[4] = {ou.x, ou.y, ou.z, ou.w}, gw4[4] = {gu.x, gu.y, gu.z, gu.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16x2(qw[t]), a = unpack_bf16x2(ow[t]), g2 = unpack_bf16x2(gw4[t]);
      qv[c * 8 + 2 * t] = f.x; qv[c * 8 + 2 * t + 1] = f.y;
      dsum += a.x * g2.x + a.y * g2.y;
    }
  }
  const size_t row = ((size_t)b * nH + n) * N + q;
  Dbuf[row] = dsum;
  if (use_rel) {
    float* dst = relbuf + row * (gh + gw);
    const int qy = q / gw, qx = q % gw;
    for (int k = 0; k < gh; ++k) {
      const float* th = tab + (qy - k + gh - 1) * SB_TLD;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 64; ++d) s += qv[d] * th[d];
      dst[k] = s;
    }
    for (int k = 0; k < gw; ++k) {
      const float* tw = tab + (rows_h + qx - k + gw - 1) * SB_TLD;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 64; ++d) s += qv[d] * tw[d];
      dst[gh + k] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------ main
// smem: K | V | Q | dO | P (2 atoms) | dS (2 atoms) | rel_s [128][R] | acc_s [128][R] | mbar | slot,   R = odd(nyb + gw)
__global__ void __launch_bounds__(SB_THREADS)
full_attn_bwd_stream_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ lse, const float* __restrict__ Dbuf,
                               const float* __restrict__ relbuf, const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqkv,
                               float* __restrict__ dq_acc, float* __restrict__ dsrow, int N, int gh, int gw, int C, int nH, int use_rel,
                               int R) {
  MTP_PDL_ENTRY();
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* Ks = sm;
  uint8_t* Vs = Ks + SBT;
  uint8_t* Qs = Vs + SBT;
  uint8_t* Gs = Qs + SBT;
  uint8_t* Pt = Gs + SBT;                // 2 atoms: key columns 0..63 | 64..127 of the block
  uint8_t* St = Pt + 2 * SBT;
  float* rel_s = reinterpret_cast<float*>(St + 2 * SBT);      // this query tile's bias terms for the block's key rows / all key columns
  float* acc_s = rel_s + 128 * R;                              // row sums of dS, same layout
  uint64_t* mbar = reinterpret_cast<uint64_t*>(acc_s + 128 * R);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int kbk = blockIdx.x, n = blockIdx.y, b = blockIdx.z;
  const int k0 = kbk * 128;
  const int C3 = 3 * C;
  const float scale = 0.125f;
  const __nv_bfloat16* base = qkv + (size_t)b * N * C3 + n * 64;
  const __nv_bfloat16* gbase = dout + (size_t)b * N * C + n * 64;
  const int jy0 = k0 / gw;                                   // first key row of the block
  const int nyb = (min(N, k0 + 128) - 1) / gw - jy0 + 1;     // key rows touched by the block
  const int GH = gh + gw;

  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 32) { mbar_init(mbar, 1); fence_barrier_init(); }
  for (int i = tid; i < 128 * 8; i += SB_THREADS) {        // resident K / V block (rows beyond N are zero)
    const int r = i >> 3, c = i & 7;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (k0 + r < N) {
      kv = *reinterpret_cast<const uint4*>(base + (size_t)(k0 + r) * C3 + C + c * 8);
      vv = *reinterpret_cast<const uint4*>(base + (size_t)(k0 + r) * C3 + 2 * C + c * 8);
    }
    *reinterpret_cast<uint4*>(Ks + tile_chunk_off(r, c)) = kv;
    *reinterpret_cast<uint4*>(Vs + tile_chunk_off(r, c)) = vv;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t T_S = tmem, T_DP = tmem + 128, T_DK = tmem + 256, T_DV = tmem + 320, T_DQ = tmem + 384;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;
  const int nqt = (N + 127) / 128;

  for (int qt = 0; qt < nqt; ++qt) {
    const int q0 = qt * 128;
    const int q = q0 + tid;
    const bool qvalid = q < N;
    for (int i = tid; i < 128 * 8; i += SB_THREADS) {       // Q and dO tiles of this query tile
      const int r = i >> 3, c = i & 7;
      const bool ok = q0 + r < N;
      sb_cp_async16(smem_u32(Qs + tile_chunk_off(r, c)), base + (size_t)(ok ? q0 + r : 0) * C3 + c * 8, ok);
      sb_cp_async16(smem_u32(Gs + tile_chunk_off(r, c)), gbase + (size_t)(ok ? q0 + r : 0) * C + c * 8, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const size_t rowbase = ((size_t)b * nH + n) * N;
    if (use_rel) {          // bias terms of the tile's rows: [key rows jy0 .. jy0+nyb) | [all key columns), coalesced along the row
      for (int i = tid; i < 128 * (nyb + gw); i += SB_THREADS) {
        const int r = i / (nyb + gw), cidx = i % (nyb + gw);
        float v = 0.f;
        if (q0 + r < N) v = relbuf[(rowbase + q0 + r) * GH + (cidx < nyb ? jy0 + cidx : gh + (cidx - nyb))];
        rel_s[r * R + cidx] = v;
        acc_s[r * R + cidx] = 0.f;
      }
    }
    const float l = qvalid ? lse[rowbase + q] : 0.f;
    const float D = qvalid ? Dbuf[rowbase + q] : 0.f;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        tc_mma_tiles<false, false>(T_S, smem_u32(Qs), 0, smem_u32(Ks), 0, 128, 128, 64, false);
        tc_mma_tiles<false, false>(T_DP, smem_u32(Gs), 0, smem_u32(Vs), 0, 128, 128, 64, false);
        umma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      float* rrow = rel_s + tid * R;
      float* arow = acc_s + tid * R;
      int jy = 0, jx = k0 % gw;                         // jy relative to jy0
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(T_S + lane_base + c * 32, r0);
        tmem_ld_32x32(T_DP + lane_base + c * 32, r1);
        tmem_ld_wait();
        float pv[32], dv[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = k0 + c * 32 + e;
          float p = 0.f, ds = 0.f;
          if (j < N) {
            float s = __uint_as_float(r0[e]);
            if (use_rel) s += rrow[jy] + rrow[nyb + jx];
            p = qvalid ? __expf(scale * s - l) : 0.f;
            ds = p * (__uint_as_float(r1[e]) - D);
            if (use_rel) { arow[jy] += ds; arow[nyb + jx] += ds; }
            if (++jx == gw) { jx = 0; ++jy; }
          }
          pv[e] = p;
          dv[e] = ds;
        }
        uint8_t* pa = Pt + (c >> 1) * SBT;
        uint8_t* sa = St + (c >> 1) * SBT;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 u;
          u.x = pack_bf16x2(pv[8 * k], pv[8 * k + 1]); u.y = pack_bf16x2(pv[8 * k + 2], pv[8 * k + 3]);
          u.z = pack_bf16x2(pv[8 * k + 4], pv[8 * k + 5]); u.w = pack_bf16x2(pv[8 * k + 6], pv[8 * k + 7]);
          *reinterpret_cast<uint4*>(pa + tile_chunk_off(tid, (c & 1) * 4 + k)) = u;
          u.x = pack_bf16x2(dv[8 * k], dv[8 * k + 1]); u.y = pack_bf16x2(dv[8 * k + 2], dv[8 * k + 3]);
          u.z = pack_bf16x2(dv[8 * k + 4], dv[8 * k + 5]); u.w = pack_bf16x2(dv[8 * k + 6], dv[8 * k + 7]);
          *reinterpret_cast<uint4*>(sa + tile_chunk_off(tid, (c & 1) * 4 + k)) = u;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        // dQ_part = dS K (fresh) ; dK += dS^T Q ; dV += P^T dO     (transposes through MN-major A operands)
        tc_mma_tiles<false, true>(T_DQ, smem_u32(St), SBT, smem_u32(Ks), 0, 128, 64, 128, false);
        tc_mma_tiles<true, true>(T_DK, smem_u32(St), SBT, smem_u32(Qs), 0, 128, 64, 128, qt > 0);
        tc_mma_tiles<true, true>(T_DV, smem_u32(Pt), SBT, smem_u32(Gs), 0, 128, 64, 128, qt > 0);
        umma_commit(mbar);
      }
      __syncwarp();
    }
    // meanwhile: this tile's dS row sums join the global accumulator (coalesced along the row)
    if (use_rel) {
      for (int i = tid; i < 128 * (nyb + gw); i += SB_THREADS) {
        const int r = i / (nyb + gw), cidx = i % (nyb + gw);
        if (q0 + r < N) {
          const float v = acc_s[r * R + cidx];
          if (v != 0.f) atomicAdd(dsrow + (rowbase + q0 + r) * GH + (cidx < nyb ? jy0 + cidx : gh + (cidx - nyb)), v);
        }
      }
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(T_DQ + lane_base, r0);
      tmem_ld_32x32(T_DQ + lane_base + 32, r1);
      tmem_ld_wait();
      if (qvalid) {
        float* dst = dq_acc + ((size_t)b * N + q) * C + n * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          sb_red_add4(dst + 4 * c, __uint_as_float(r0[4 * c]), __uint_as_float(r0[4 * c + 1]), __uint_as_float(r0[4 * c + 2]), __uint_as_float(r0[4 * c + 3]));
          sb_red_add4(dst + 32 + 4 * c, __uint_as_float(r1[4 * c]), __uint_as_float(r1[4 * c + 1]), __uint_as_float(r1[4 * c + 2]), __uint_as_float(r1[4 * c + 3]));
        }
      }
    }
    tc_fence_before();
    __syncthreads();          // Q / dO / P / dS tiles, T_S / T_DP / T_DQ and the rel / acc rows are rewritten by the next query tile
    tc_fence_after();
  }

  // ---- dK (x scale) and dV rows of this key block: thread r = key k0 + r (each key belongs to exactly one CTA: plain stores)
  {
    const int j = k0 + tid;
    uint32_t r0[32], r1[32];
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      const uint32_t t0 = (part == 0 ? T_DK : T_DV) + lane_base;
      tmem_ld_32x32(t0, r0);
      tmem_ld_32x32(t0 + 32, r1);
      tmem_ld_wait();
      if (j < N) {
        const float f = part == 0 ? scale : 1.0f;
        __nv_bfloat16* o = dqkv + ((size_t)b * N + j) * C3 + n * 64 + (part == 0 ? C : 2 * C);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(f * __uint_as_float(r0[8 * c]), f * __uint_as_float(r0[8 * c + 1]));
          u.y = pack_bf16x2(f * __uint_as_float(r0[8 * c + 2]), f * __uint_as_float(r0[8 * c + 3]));
          u.z = pack_bf16x2(f * __uint_as_float(r0[8 * c + 4]), f * __uint_as_float(r0[8 * c + 5]));
          u.w = pack_bf16x2(f * __uint_as_float(r0[8 * c + 6]), f * __uint_as_float(r0[8 * c + 7]));
          *reinterpret_cast<uint4*>(o + 8 * c) = u;
          u.x = pack_bf16x2(f * __uint_as_float(r1[8 * c]), f * __uint_as_float(r1[8 * c + 1]));
          u.y = pack_bf16x2(f * __uint_as_float(r1[8 * c + 2]), f * __uint_as_float(r1[8 * c + 3]));
          u.z = pack_bf16x2(f * __uint_as_float(r1[8 * c + 4]), f * __uint_as_float(r1[8 * c + 5]));
          u.w = pack_bf16x2(f * __uint_as_float(r1[8 * c + 6]), f * __uint_as_float(r1[8 * c + 7]));
          *reinterpret_cast<uint4*>(o + 32 + 8 * c) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------------------ finish
// smem: Q tile | W_h (2 atoms) | W_w (2 atoms) | tables [(2gh-1)+(2gw-1)][SB_TLD] | mbar | slot
__global__ void __launch_bounds__(SB_THREADS)
dense_bwd_finish_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                        const float* __restrict__ dq_acc, const float* __restrict__ dsrow, __nv_bfloat16* __restrict__ dqkv,
                        float* __restrict__ d_rel_h, float* __restrict__ d_rel_w, int N, int gh, int gw, int C, int nH, int use_rel) {
  MTP_PDL_ENTRY();
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* Qs = sm;
  uint8_t* Wh = Qs + SBT;               // [128 q][128 r] bf16 as 2 atoms (r 0..63 | 64..127)
  uint8_t* Ww = Wh + 2 * SBT;
  float* tab = reinterpret_cast<float*>(Ww + 2 * SBT);
  const int rows_h = 2 * gh - 1, rows_w = 2 * gw - 1;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tab + (rows_h + rows_w) * SB_TLD);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, n = blockIdx.y, b = blockIdx.z;
  const int q = q0 + tid;
  const bool qvalid = q < N;
  const int C3 = 3 * C;
  const float scale = 0.125f;
  const size_t tok = (size_t)b * N + (qvalid ? q : 0);

  float dq[64];
  if (qvalid) {
    const float* src = dq_acc + tok * C + n * 64;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(src + 4 * c);
      dq[4 * c] = v.x; dq[4 * c + 1] = v.y; dq[4 * c + 2] = v.z; dq[4 * c + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int d = 0; d < 64; ++d) dq[d] = 0.f;
  }
  if (!use_rel) {
    if (qvalid) {
      __nv_bfloat16* dst = dqkv + tok * C3 + n * 64;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u;
        u.x = pack_bf16x2(scale * dq[8 * c], scale * dq[8 * c + 1]); u.y = pack_bf16x2(scale * dq[8 * c + 2], scale * dq[8 * c + 3]);
        u.z = pack_bf16x2(scale * dq[8 * c + 4], scale * dq[8 * c + 5]); u.w = pack_bf16x2(scale * dq[8 * c + 6], scale * dq[8 * c + 7]);
        *reinterpret_cast<uint4*>(dst + 8 * c) = u;
      }
    }
    return;
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  if (tid == 32) { mbar_init(mbar, 1); fence_barrier_init(); }
  for (int i = tid; i < rows_h * 64; i += SB_THREADS) tab[(i >> 6) * SB_TLD + (i & 63)] = rel_h[i];
  for (int i = tid; i < rows_w * 64; i += SB_THREADS) tab[(rows_h + (i >> 6)) * SB_TLD + (i & 63)] = rel_w[i];
  for (int i = tid; i < 128 * 8; i += SB_THREADS) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q0 + r < N) v = *reinterpret_cast<const uint4*>(qkv + ((size_t)b * N + q0 + r) * C3 + n * 64 + c * 8);
    *reinterpret_cast<uint4*>(Qs + tile_chunk_off(r, c)) = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const int qy = qvalid ? q / gw : 0, qx = qvalid ? q % gw : 0;
  const float* srow = dsrow + (((size_t)b * nH + n) * N + (qvalid ? q : 0)) * (gh + gw);
  // ---- dq += sum_k dSh[k] Rh[qy-k+gh-1] + sum_k dSw[k] Rw[qx-k+gw-1] ; and row q of the weight tiles W_h[q][r] = dSh[q][qy-r+gh-1], W_w likewise
  for (int a = 0; a < 2; ++a) {
    uint8_t* Wt = a == 0 ? Wh : Ww;
    const int G = a == 0 ? gh : gw, qq = a == 0 ? qy : qx, rows = a == 0 ? rows_h : rows_w;
    const float* tb = tab + (a == 0 ? 0 : rows_h) * SB_TLD;
    const float* sv = srow + (a == 0 ? 0 : gh);
    for (int c = 0; c < 16; ++c) {          // 16 chunks of 8 table rows r
      float wv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int r = c * 8 + e;
        const int k = qq - r + G - 1;
        float v = 0.f;
        if (qvalid && r < rows && k >= 0 && k < G) {
          v = sv[k];
          const float* trow = tb + r * SB_TLD;
#pragma unroll
          for (int d = 0; d < 64; ++d) dq[d] += v * trow[d];
        }
        wv[e] = v;
      }
      uint4 u;
      u.x = pack_bf16x2(wv[0], wv[1]); u.y = pack_bf16x2(wv[2], wv[3]); u.z = pack_bf16x2(wv[4], wv[5]); u.w = pack_bf16x2(wv[6], wv[7]);
      *reinterpret_cast<uint4*>(Wt + (c >> 3) * SBT + tile_chunk_off(tid, c & 7)) = u;
    }
  }
  if (qvalid) {
    __nv_bfloat16* dst = dqkv + tok * C3 + n * 64;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 u;
      u.x = pack_bf16x2(scale * dq[8 * c], scale * dq[8 * c + 1]); u.y = pack_bf16x2(scale * dq[8 * c + 2], scale * dq[8 * c + 3]);
      u.z = pack_bf16x2(scale * dq[8 * c + 4], scale * dq[8 * c + 5]); u.w = pack_bf16x2(scale * dq[8 * c + 6], scale * dq[8 * c + 7]);
      *reinterpret_cast<uint4*>(dst + 8 * c) = u;
    }
  }
  tc_fence_before();
  fence_proxy_async_smem();
  __syncthreads();
  // ---- d rel tables of this query tile: dR[r][:] = sum_q W[q][r] q[q][:]   (M = 128 table rows, N = 64, K = 128 queries)
  if (warp == 0) {
    tc_fence_after();
    if (elect_one()) {
      tc_mma_tiles<true, true>(tmem, smem_u32(Wh), SBT, smem_u32(Qs), 0, 128, 64, 128, false);
      tc_mma_tiles<true, true>(tmem + 64, smem_u32(Ww), SBT, smem_u32(Qs), 0, 128, 64, 128, false);
      umma_commit(mbar);
    }
    __syncwarp();
  }
  mbar_wait(mbar, 0);
  tc_fence_after();
#pragma unroll 1
  for (int a = 0; a < 2; ++a) {
    uint32_t r0[32], r1[32];
    tmem_ld_32x32(tmem + 64 * a + lane_base, r0);
    tmem_ld_32x32(tmem + 64 * a + lane_base + 32, r1);
    tmem_ld_wait();
    const int rows = a == 0 ? rows_h : rows_w;
    if (tid < rows) {
      float* dst = (a == 0 ? d_rel_h : d_rel_w) + (size_t)tid * 64;
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        atomicAdd(dst + d, scale * __uint_as_float(r0[d]));
        atomicAdd(dst + 32 + d, scale * __uint_as_float(r1[d]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

static inline size_t sb_pad(size_t n) { return (n + 15) / 16 * 16; }      // regions start on 64-byte boundaries (float4 / red.v4 accesses)
size_t full_attn_bwd_stream_workspace_bytes(int B, int gh, int gw, int nH) {
  const size_t N = (size_t)gh * gw, rows = (size_t)B * nH * N;
  return (sb_pad(rows) + 2 * sb_pad(rows * (gh + gw)) + (size_t)B * N * nH * 64) * sizeof(float);
}

int launch_full_attn_bwd_stream_tc(const void* qkv, const float* rel_h, const float* rel_w, const float* lse, const void* out, const void* dout,
                                   void* dqkv, float* d_rel_h, float* d_rel_w, void* workspace, int B, int gh, int gw, int C, int nH,
                                   cudaStream_t st) {
  const int N = gh * gw, GH = gh + gw;
  const int use_rel = rel_h != nullptr;
  const size_t rows = (size_t)B * nH * N;
  float* Dbuf = reinterpret_cast<float*>(workspace);
  float* relbuf = Dbuf + sb_pad(rows);
  float* dsrow = relbuf + sb_pad(rows * GH);
  float* dq_acc = dsrow + sb_pad(rows * GH);
  const int tab_bytes = (2 * gh - 1 + 2 * gw - 1) * SB_TLD * 4;
  const int nyb_max = std::min(gh, 128 / gw + 2);
  const int R = (nyb_max + gw) | 1;
  const int smem_main = 8 * SBT + 2 * 128 * R * 4 + 64;
  const int smem_fin = 5 * SBT + tab_bytes + 64;
  MTP_REQUIRE(tab_bytes <= 200 * 1024 && smem_main <= 227 * 1024 && smem_fin <= 227 * 1024 && 2 * gh - 1 <= 128 && 2 * gw - 1 <= 128,
              "mtp_full_attn_bwd: grid %dx%d too large for the streaming tensor-core kernels", gh, gw);
  static int a_prep = 0, a_main = 0, a_fin = 0;
  cudaError_t e = cudaSuccess;
  if (tab_bytes > a_prep) { e = cudaFuncSetAttribute(dense_bwd_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tab_bytes); a_prep = tab_bytes; }
  if (e == cudaSuccess && smem_main > a_main) { e = cudaFuncSetAttribute(full_attn_bwd_stream_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_main); a_main = smem_main; }
  if (e == cudaSuccess && smem_fin > a_fin) { e = cudaFuncSetAttribute(dense_bwd_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fin); a_fin = smem_fin; }
  if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "full_attn_bwd_stream smem attr: %s", cudaGetErrorString(e));
  e = cudaMemsetAsync(dsrow, 0, (sb_pad(rows * GH) + (size_t)B * N * C) * sizeof(float), st);      // dsrow | dq_acc are contiguous
  if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "full_attn_bwd_stream memset: %s", cudaGetErrorString(e));
  const dim3 grid(ceil_div(N, 128), nH, B);
  const __nv_bfloat16* q16 = reinterpret_cast<const __nv_bfloat16*>(qkv);
  (void)launch_k(dense_bwd_prep_kernel, grid, SB_THREADS, tab_bytes, st, q16, rel_h, rel_w, reinterpret_cast<const __nv_bfloat16*>(out),
                 reinterpret_cast<const __nv_bfloat16*>(dout), Dbuf, relbuf, N, gh, gw, C, nH, use_rel);
  int rc = check_launch("dense_bwd_prep_kernel");
  if (rc) return rc;
  (void)launch_k(full_attn_bwd_stream_tc_kernel, grid, SB_THREADS, smem_main, st, q16, lse, Dbuf, relbuf, reinterpret_cast<const __nv_bfloat16*>(dout),
                 reinterpret_cast<__nv_bfloat16*>(dqkv), dq_acc, dsrow, N, gh, gw, C, nH, use_rel, R);
  rc = check_launch("full_attn_bwd_stream_tc_kernel");
  if (rc) return rc;
  (void)launch_k(dense_bwd_finish_kernel, grid, SB_THREADS, smem_fin, st, q16, rel_h, rel_w, dq_acc, dsrow, reinterpret_cast<__nv_bfloat16*>(dqkv),
                 d_rel_h, d_rel_w, N, gh, gw, C, nH, use_rel);
  return check_launch("dense_bwd_finish_kernel");
}

}  // namespace mtp
