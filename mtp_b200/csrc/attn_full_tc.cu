// Dense attention with decomposed rel-pos bias on tcgen05 tensor cores, forward and backward, for token grids with
// N = gh*gw <= 256 (224^2 and 256^2 inputs: the dense blocks of the headline configuration).  Larger grids use the
// streaming tensor-core kernels of attn_full_stream_tc.cu / attn_full_stream_bwd_tc.cu.       [V]:90-111, 142-193; SURVEY.md K6
//
// One CTA (128 threads) per (image, head).  K and V of the head live in shared memory for the whole CTA as 128B-swizzled
// bf16 tiles ([256 rows] x 128 B); query tiles of 128 rows are processed in turn:
//   forward : S = Q K^T (UMMA M=128, N=N16, K=64) -> thread r owns row r: two passes over its TMEM row (max, then
//             exp / sum / bf16 P into a [128 x 256] tile) with the rel-pos bias  scale * (q.Rh[qy-jy+gh-1] + q.Rw[qx-jx+gw-1])
//             added in fp32 -> O = P V (K = N16) -> O / sum -> bf16; LSE saved.
//   backward: per (query tile, key half of 128): S then dP into TMEM; rows give P = exp(S - lse), dS = P (dP - D);
//             dQ += dS K (registers), dK_half += dS^T Q and dV_half += P^T dO accumulate in TMEM across query tiles
//             (transposes via MN-major A operands); rel-pos table gradients reduced per CTA, then atomics.
#include "common.h"
#include "ptx.cuh"
#include "tc_tile.cuh"

namespace mtp {

constexpr int FT_THREADS = 128;
constexpr int FT_TILE = 128 * 128;       // bytes of a 128-row tile
constexpr int FT_MAXG = 16;              // max grid side (N <= 256)
constexpr int FT_TS = 68;               // row pitch (floats) of the rel-pos tables in shared memory: rows read by the 8 lanes of a quarter warp
                                        // (consecutive qx) fall into different banks (pitch 64: every row in the same banks, ncu: 76 % of all
                                        // shared-memory wavefronts of the forward kernel were conflicts)
constexpr int FT_TAB = (2 * FT_MAXG - 1) * FT_TS;   // floats per rel-pos table

// copy rows [row0, row0+nrows) of a head slice (64 bf16 per row, row pitch ld) into a swizzled tile; rows >= nvalid are zero
__device__ __forceinline__ void ft_load_rows(uint8_t* tile, const __nv_bfloat16* src, size_t ld, int row0, int nrows, int nvalid) {
  for (int i = threadIdx.x; i < nrows * 8; i += FT_THREADS) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row0 + r < nvalid) v = *reinterpret_cast<const uint4*>(src + (size_t)(row0 + r) * ld + c * 8);
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(r, c)) = v;
  }
}

// this thread's rel-pos terms: rh[k] = q . Rh[qy - k + gh - 1] (k < gh), rw[k] = q . Rw[qx - k + gw - 1] (k < gw), UNscaled q
__device__ __forceinline__ void ft_rel_terms(const uint8_t* Qs, const float* relh_t, const float* relw_t, int row, int qy, int qx, int gh,
                                             int gw, float* rh_s, float* rw_s) {
  float qv[64];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(Qs + tile_chunk_off(row, c));
    const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) { const float2 f = unpack_bf16x2(w4[t]); qv[c * 8 + 2 * t] = f.x; qv[c * 8 + 2 * t + 1] = f.y; }
  }
  for (int k = 0; k < gh; ++k) {
    const float* th = relh_t + (qy - k + gh - 1) * FT_TS;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) s += qv[d] * th[d];
    rh_s[k] = s;
  }
  for (int k = 0; k < gw; ++k) {
    const float* tw = relw_t + (qx - k + gw - 1) * FT_TS;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) s += qv[d] * tw[d];
    rw_s[k] = s;
  }
}

// Grid side known at compile time (G = 14: the 224^2 headline case): the rel terms stay in registers and every key index of the
// softmax / dS loops is a constant, so the per-element shared-memory traffic of the generic path disappears.
template <int G>
__device__ __forceinline__ void ft_rel_terms_fixed(const uint8_t* Qs, const float* relh_t, const float* relw_t, int row, int qy, int qx,
                                                   float (&rh)[G], float (&rw)[G]) {
#pragma unroll
  for (int k = 0; k < G; ++k) { rh[k] = 0.f; rw[k] = 0.f; }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(Qs + tile_chunk_off(row, c));
    const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
    float qv[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) { const float2 f = unpack_bf16x2(w4[t]); qv[2 * t] = f.x; qv[2 * t + 1] = f.y; }
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const float4* th = reinterpret_cast<const float4*>(relh_t + (qy - k + G - 1) * FT_TS + c * 8);
      const float4* tw = reinterpret_cast<const float4*>(relw_t + (qx - k + G - 1) * FT_TS + c * 8);
      const float4 h0 = th[0], h1 = th[1], w0 = tw[0], w1 = tw[1];
      rh[k] += qv[0] * h0.x + qv[1] * h0.y + qv[2] * h0.z + qv[3] * h0.w + qv[4] * h1.x + qv[5] * h1.y + qv[6] * h1.z + qv[7] * h1.w;
      rw[k] += qv[0] * w0.x + qv[1] * w0.y + qv[2] * w0.z + qv[3] * w0.w + qv[4] * w1.x + qv[5] * w1.y + qv[6] * w1.z + qv[7] * w1.w;
    }
  }
}

// ---- rel-pos terms on the tensor core (grid side G known at compile time) ---------------------------------------------------------------
// REL[q][t] = q . R[t]  for the stacked table R = [Rh rows 0 .. 2G-2 | zero | Rw rows at 32 .. 32+2G-2 | zero] (64 rows) is one more
// 128 x 64 x 64 MMA on the Q tile instead of 2 * G dot products of length 64 per thread (1800 FMAs + 900 shared-memory loads per row: 40 %
// of the forward kernel's instructions).  The fp32 tables are split into bf16 hi + lo words (two accumulating MMAs): exact to 2^-17.
// The same tile, read MN-major, is the B operand of the backward's  dq += W R  (W = the dS row / column sums).
template <int G>
__device__ __forceinline__ void ft_build_rel_tiles(uint8_t* hi, uint8_t* lo, const float* __restrict__ rel_h, const float* __restrict__ rel_w) {
  static_assert(2 * G - 1 <= 32, "table rows");
  static_assert(64 * 8 % FT_THREADS == 0, "items per thread");
  constexpr int NI = 64 * 8 / FT_THREADS;      // 4 (row, 8-column chunk) items per thread: all their loads are issued before the first conversion
  float4 va[NI], vb[NI];
#pragma unroll
  for (int k = 0; k < NI; ++k) {
    const int i = threadIdx.x + k * FT_THREADS, r = i >> 3, c = i & 7;
    const int t = r < 32 ? r : r - 32;
    va[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    vb[k] = va[k];
    if (t < 2 * G - 1) {
      const float* src = (r < 32 ? rel_h : rel_w) + t * 64 + c * 8;
      va[k] = __ldg(reinterpret_cast<const float4*>(src));
      vb[k] = __ldg(reinterpret_cast<const float4*>(src + 4));
    }
  }
#pragma unroll
  for (int k = 0; k < NI; ++k) {
    const int i = threadIdx.x + k * FT_THREADS, r = i >> 3, c = i & 7;
    const float v[8] = {va[k].x, va[k].y, va[k].z, va[k].w, vb[k].x, vb[k].y, vb[k].z, vb[k].w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float h0 = __bfloat162float(__float2bfloat16_rn(v[2 * e])), h1 = __bfloat162float(__float2bfloat16_rn(v[2 * e + 1]));
      h[e] = pack_bf16x2(h0, h1);
      l[e] = pack_bf16x2(v[2 * e] - h0, v[2 * e + 1] - h1);
    }
    *reinterpret_cast<uint4*>(hi + tile_chunk_off(r, c)) = make_uint4(h[0], h[1], h[2], h[3]);      // rows outside the tables: zeros
    *reinterpret_cast<uint4*>(lo + tile_chunk_off(r, c)) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// out[k] = v[base + s + (G - 1) - k], k < G, for a run-time shift s in [0, G): a 4-stage barrel shifter on registers (no dynamic indexing)
template <int G>
__device__ __forceinline__ void ft_pick_terms(const uint32_t (&v)[32], int s, float (&out)[G]) {
  float a[2 * G - 1];
#pragma unroll
  for (int j = 0; j < 2 * G - 1; ++j) a[j] = __uint_as_float(v[j]);
#pragma unroll
  for (int st = 8; st >= 1; st >>= 1) {
    const bool on = (s & st) != 0;
#pragma unroll
    for (int j = 0; j + st < 2 * G - 1; ++j) a[j] = on ? a[j + st] : a[j];
  }
#pragma unroll
  for (int k = 0; k < G; ++k) out[k] = a[G - 1 - k];
}

// the inverse: out[t] = in[s + (G - 1) - t] for t - s in [0, G), else 0  (t < 32): a dS row / column sum lands on the table row it multiplies
template <int G>
__device__ __forceinline__ void ft_spread_terms(const float (&in)[G], int s, float (&out)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) out[j] = j < G ? in[G - 1 - j] : 0.f;
#pragma unroll
  for (int st = 8; st >= 1; st >>= 1) {
    const bool on = (s & st) != 0;
#pragma unroll
    for (int j = 31; j >= 0; --j) out[j] = on ? (j >= st ? out[j - st] : 0.f) : out[j];
  }
}

// ================================================================================================== forward
constexpr int FTF_SMEM = FT_TILE /*Q*/ + 2 * 2 * FT_TILE /*K,V*/ + 4 * FT_TILE /*P*/ + (2 * FT_TAB + 2 * 128 * 17) * 4 + 64;
constexpr int ftf_smem_duo(int G) { return 4 * FT_TILE /*P over Q,K*/ + ((G * G + 15) & ~15) * 128 /*V*/ + 2 * 64 * 128 /*rel-pos table tiles*/ + 64; }
static_assert(2 * (ftf_smem_duo(14) + 1024) <= 228 * 1024, "two forward CTAs of the 14 x 14 case must fit one SM");

template <int G>      // G > 0: gh == gw == G known at compile time (and rel-pos in use); G == 0: generic
__global__ void __launch_bounds__(FT_THREADS)
full_attn_fwd_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                        __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N, int gh, int gw, int C, int nH, int use_rel) {
  MTP_PDL_ENTRY();
  extern __shared__ __align__(1024) uint8_t sm[];
  // G > 0 (the 14 x 14 grid of the headline configuration): a layout small enough for TWO CTAs per SM (105 KB, 256 TMEM columns, 165
  // registers x 128 threads).  One CTA per SM = one warp per scheduler ran at 13 cycles per instruction (ncu: long / short scoreboard and
  // instruction-fetch stalls with nothing to switch to), and 256 CTAs took two rounds on 148 SMs.  The P tile (4 atoms, 64 KB) lies over Q and
  // K, which are dead once S = Q K^T is complete and every thread has its rel-pos terms (barrier below); O reuses the first TMEM columns of S.
  constexpr bool DUO = G > 0;
  constexpr int KROWS = DUO ? ((G * G + 15) & ~15) : 256;        // rows of the K and V tiles
  constexpr int TROWS = DUO ? 2 * G - 1 : 2 * FT_MAXG - 1;        // rows of each rel-pos table
  uint8_t* Pt = sm;                                               // 4 atoms of 128 rows
  uint8_t* Qs = DUO ? sm : sm + 4 * FT_TILE;
  uint8_t* Ks = Qs + FT_TILE;
  uint8_t* Vs = DUO ? sm + 4 * FT_TILE : Ks + 2 * FT_TILE;
  // DUO: the stacked rel-pos table as two bf16 B-operand tiles (hi, lo words) of 64 rows; generic: fp32 tables + per-row terms in smem
  uint8_t* Rhi = Vs + KROWS * 128;         // KROWS * 128 is a multiple of 1024 for G = 14 (208 rows)
  uint8_t* Rlo = Rhi + 64 * 128;
  float* relh_t = reinterpret_cast<float*>(DUO ? Rlo + 64 * 128 : Vs + KROWS * 128);
  float* relw_t = relh_t + (DUO ? 0 : TROWS * FT_TS);
  float* rh_s = relw_t + (DUO ? 0 : TROWS * FT_TS);    // [128][17] per-row rel terms (generic path only)
  float* rw_s = rh_s + (DUO ? 0 : 128 * 17);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(rw_s + (DUO ? 0 : 128 * 17));
  static_assert(!DUO || (KROWS * 128) % 1024 == 0, "tile alignment");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  constexpr int TMEM_COLS = DUO ? 256 : 512;
  constexpr int GD = G > 0 ? G : 1;

  const int tid = threadIdx.x, warp = tid >> 5;
  // one CTA per (image, head, 128-query tile): K / V of the head are re-staged by each of the (1 or 2) query-tile CTAs, which doubles the
  // CTA count at 224^2 (2 x 128 = 256 CTAs instead of 128 on 148 SMs) for 32 KB of extra L2 reads per CTA
  const int nqt = (N + 127) / 128;
  const int qt_idx = blockIdx.x % nqt;
  const int n = (blockIdx.x / nqt) % nH, b = blockIdx.x / (nqt * nH);
  const int C3 = 3 * C;
  const float scale = 0.125f;
  const __nv_bfloat16* base = qkv + (size_t)b * N * C3 + n * 64;
  const int N16 = (N + 15) & ~15;

  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  if (tid == 32) { mbar_init(mbar, 1); fence_barrier_init(); }
  ft_load_rows(Ks, base + C, C3, 0, KROWS, N);
  ft_load_rows(Vs, base + 2 * C, C3, 0, KROWS, N);
  if constexpr (DUO) {
    ft_build_rel_tiles<GD>(Rhi, Rlo, rel_h, rel_w);
  } else if (use_rel) {
    for (int i = tid; i < (2 * gh - 1) * 64; i += FT_THREADS) relh_t[(i >> 6) * FT_TS + (i & 63)] = rel_h[i];
    for (int i = tid; i < (2 * gw - 1) * 64; i += FT_THREADS) relw_t[(i >> 6) * FT_TS + (i & 63)] = rel_w[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t T_S = tmem, T_O = DUO ? tmem : tmem + 256;      // DUO: O is written after every row of S has been read
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;

  for (int q0 = qt_idx * 128; q0 < min(N, (qt_idx + 1) * 128); q0 += 128) {
    ft_load_rows(Qs, base, C3, q0, 128, N);
    fence_proxy_async_smem();
    __syncthreads();
    const int q = q0 + tid;
    const bool qvalid = q < N;
    const int qy = qvalid ? q / gw : 0, qx = qvalid ? q % gw : 0;
    float* rh = rh_s + tid * 17;
    float* rw = rw_s + tid * 17;
    constexpr int GG = G > 0 ? G : 1;
    float rhr[GG], rwr[GG];
    if constexpr (DUO) {
      // rel-pos terms first, into the TMEM columns S will overwrite (256 columns per CTA leave no room for both)
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          tc_mma_tiles<false, false>(T_S, smem_u32(Qs), 0, smem_u32(Rhi), 0, 128, 64, 64, false);
          tc_mma_tiles<false, false>(T_S, smem_u32(Qs), 0, smem_u32(Rlo), 0, 128, 64, 64, true);
          umma_commit(mbar);
        }
        __syncwarp();
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      tc_fence_after();
      {
        uint32_t vh[32], vw[32];
        tmem_ld_32x32(T_S + lane_base, vh);
        tmem_ld_32x32(T_S + lane_base + 32, vw);
        tmem_ld_wait();
        ft_pick_terms<GD>(vh, qy, rhr);      // rhr[k] = q . Rh[qy - k + G - 1]
        ft_pick_terms<GD>(vw, qx, rwr);
      }
      tc_fence_before();
      __syncthreads();                       // every row of REL has been read: S may overwrite it
    }
    if (warp == 0) {      // warp-uniform issue: descriptors stay in uniform registers (no ELECT/R2UR waterfall per MMA)
      tc_fence_after();
      if (elect_one()) {
        tc_mma_tiles<false, false>(T_S, smem_u32(Qs), 0, smem_u32(Ks), 0, 128, N16, 64, false);
        umma_commit(mbar);
      }
      __syncwarp();
    }
    if constexpr (!DUO) {
      if (use_rel) ft_rel_terms(Qs, relh_t, relw_t, tid, qy, qx, gh, gw, rh, rw);
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();

    const int n_chunks = (N + 31) / 32;
    float m = -INFINITY;
    float sum = 0.f;
    if constexpr (G > 0) {
      constexpr int NF = G * G, NCH = (NF + 31) / 32;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {          // pass 1: row maximum
        uint32_t r[32];
        tmem_ld_32x32(T_S + lane_base + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = c * 32 + e;
          if (j < NF) m = fmaxf(m, __uint_as_float(r[e]) + rhr[j / G] + rwr[j % G]);
        }
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) {          // pass 2: exp, sum, P (bf16, unnormalised) -> smem
        uint32_t r[32];
        tmem_ld_32x32(T_S + lane_base + c * 32, r);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = c * 32 + e;
          float p = 0.f;
          if (j < NF) p = qvalid ? __expf(scale * (__uint_as_float(r[e]) + rhr[j / G] + rwr[j % G] - m)) : 0.f;
          pv[e] = p;
          sum += p;
        }
        uint8_t* atom = Pt + (c >> 1) * FT_TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 u;
          u.x = pack_bf16x2(pv[8 * k], pv[8 * k + 1]); u.y = pack_bf16x2(pv[8 * k + 2], pv[8 * k + 3]);
          u.z = pack_bf16x2(pv[8 * k + 4], pv[8 * k + 5]); u.w = pack_bf16x2(pv[8 * k + 6], pv[8 * k + 7]);
          *reinterpret_cast<uint4*>(atom + tile_chunk_off(tid, (c & 1) * 4 + k)) = u;
        }
      }
    } else {
    // pass 1: row maximum
    {
      int jy = 0, jx = 0;
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(T_S + lane_base + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = c * 32 + e;
          if (j < N) {
            float s = __uint_as_float(r[e]);
            if (use_rel) s += rh[jy] + rw[jx];
            m = fmaxf(m, s);
            if (++jx == gw) { jx = 0; ++jy; }
          }
        }
      }
    }
    // pass 2: exp, sum, P (bf16, unnormalised) -> smem
    {
      int jy = 0, jx = 0;
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(T_S + lane_base + c * 32, r);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = c * 32 + e;
          float p = 0.f;
          if (j < N) {
            float s = __uint_as_float(r[e]);
            if (use_rel) s += rh[jy] + rw[jx];
            p = qvalid ? __expf(scale * (s - m)) : 0.f;
            if (++jx == gw) { jx = 0; ++jy; }
          }
          pv[e] = p;
          sum += p;
        }
        uint8_t* atom = Pt + (c >> 1) * FT_TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 u;
          u.x = pack_bf16x2(pv[8 * k], pv[8 * k + 1]); u.y = pack_bf16x2(pv[8 * k + 2], pv[8 * k + 3]);
          u.z = pack_bf16x2(pv[8 * k + 4], pv[8 * k + 5]); u.w = pack_bf16x2(pv[8 * k + 6], pv[8 * k + 7]);
          *reinterpret_cast<uint4*>(atom + tile_chunk_off(tid, (c & 1) * 4 + k)) = u;
        }
      }
    }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {      // warp-uniform issue: descriptors stay in uniform registers (no ELECT/R2UR waterfall per MMA)
      tc_fence_after();
      if (elect_one()) {
        tc_mma_tiles<false, true>(T_O, smem_u32(Pt), FT_TILE, smem_u32(Vs), 0, 128, 64, N16, false);
        umma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(T_O + lane_base, r0);
      tmem_ld_32x32(T_O + lane_base + 32, r1);
      tmem_ld_wait();
      if (qvalid) {
        const float inv = 1.0f / sum;
        __nv_bfloat16* dst = out + ((size_t)b * N + q) * C + n * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r0[8 * c]) * inv, __uint_as_float(r0[8 * c + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(r0[8 * c + 2]) * inv, __uint_as_float(r0[8 * c + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(r0[8 * c + 4]) * inv, __uint_as_float(r0[8 * c + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(r0[8 * c + 6]) * inv, __uint_as_float(r0[8 * c + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * c) = u;
          u.x = pack_bf16x2(__uint_as_float(r1[8 * c]) * inv, __uint_as_float(r1[8 * c + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(r1[8 * c + 2]) * inv, __uint_as_float(r1[8 * c + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(r1[8 * c + 4]) * inv, __uint_as_float(r1[8 * c + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(r1[8 * c + 6]) * inv, __uint_as_float(r1[8 * c + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 32 + 8 * c) = u;
        }
        if (lse) lse[((size_t)b * nH + n) * N + q] = scale * m + __logf(sum);
      }
    }
    tc_fence_before();
    __syncthreads();          // TMEM rows and the Q / P tiles are free for the next query tile
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

int launch_full_attn_fwd_tc(const void* qkv, const float* rel_h, const float* rel_w, void* out, float* lse, int B, int gh, int gw, int C,
                            int nH, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(full_attn_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FTF_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(full_attn_fwd_tc_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, ftf_smem_duo(14));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(full_attn_fwd_tc_kernel<14>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "full_attn_fwd_tc smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int nqt = (gh * gw + 127) / 128;
  if (gh == 14 && gw == 14 && rel_h != nullptr)
    (void)launch_k(full_attn_fwd_tc_kernel<14>, B * nH * nqt, FT_THREADS, ftf_smem_duo(14), st, reinterpret_cast<const __nv_bfloat16*>(qkv), rel_h, rel_w,
                   reinterpret_cast<__nv_bfloat16*>(out), lse, gh * gw, gh, gw, C, nH, 1);
  else
    (void)launch_k(full_attn_fwd_tc_kernel<0>, B * nH * nqt, FT_THREADS, FTF_SMEM, st, reinterpret_cast<const __nv_bfloat16*>(qkv), rel_h, rel_w,
                   reinterpret_cast<__nv_bfloat16*>(out), lse, gh * gw, gh, gw, C, nH, rel_h != nullptr);
  return check_launch("full_attn_fwd_tc_kernel");
}

// ================================================================================================== backward
// smem: Q | dO | K (256 rows) | V (256 rows) | P (2 atoms) | dS (2 atoms) | tables | per-row rel terms | dSh, dSw | D, lse
constexpr int FTB_SMEM = 2 * FT_TILE + 4 * FT_TILE + 4 * FT_TILE + (2 * FT_TAB + 4 * 128 * 17 + 2 * 128) * 4 + 64;

template <int G>      // G > 0: gh == gw == G known at compile time (and rel-pos in use); G == 0: generic
__global__ void __launch_bounds__(FT_THREADS)
full_attn_bwd_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                        const float* __restrict__ lse, const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                        __nv_bfloat16* __restrict__ dqkv, float* __restrict__ d_rel_h, float* __restrict__ d_rel_w, int N, int gh, int gw,
                        int C, int nH, int use_rel) {
  MTP_PDL_ENTRY();
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* Qs = sm;
  uint8_t* Gs = Qs + FT_TILE;
  uint8_t* Ks = Gs + FT_TILE;
  uint8_t* Vs = Ks + 2 * FT_TILE;
  uint8_t* Pt = Vs + 2 * FT_TILE;          // [128 q x 128 keys of the current half]: 2 atoms
  uint8_t* St = Pt + 2 * FT_TILE;
  float* relh_t = reinterpret_cast<float*>(St + 2 * FT_TILE);
  float* relw_t = relh_t + FT_TAB;
  float* rh_s = relw_t + FT_TAB;           // [128][17]
  float* rw_s = rh_s + 128 * 17;
  float* dSh = rw_s + 128 * 17;            // [128][17] row sums of dS per key row
  float* dSw = dSh + 128 * 17;
  float* D_s = dSw + 128 * 17;             // [128]
  float* lse_s = D_s + 128;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(lse_s + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int n = blockIdx.x % nH, b = blockIdx.x / nH;
  const int C3 = 3 * C;
  const float scale = 0.125f;
  const __nv_bfloat16* base = qkv + (size_t)b * N * C3 + n * 64;
  const int N16 = G > 0 ? ((G * G + 15) & ~15) : ((N + 15) & ~15);
  const int n_halves = G > 0 ? (G * G + 127) / 128 : (N + 127) / 128;
  constexpr int GG = G > 0 ? G : 1;

  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 32) { mbar_init(mbar, 1); fence_barrier_init(); }
  ft_load_rows(Ks, base + C, C3, 0, 256, N);
  ft_load_rows(Vs, base + 2 * C, C3, 0, 256, N);
  // G > 0: the stacked rel-pos table as bf16 hi / lo operand tiles (in the space of the fp32 tables), see ft_build_rel_tiles
  uint8_t* Rhi = reinterpret_cast<uint8_t*>(relh_t);
  uint8_t* Rlo = Rhi + 64 * 128;
  static_assert(2 * FT_TAB * 4 >= 2 * 64 * 128, "the table tiles fit the fp32 table region");
  if constexpr (G > 0) {
    ft_build_rel_tiles<GG>(Rhi, Rlo, rel_h, rel_w);
  } else if (use_rel) {
    for (int i = tid; i < (2 * gh - 1) * 64; i += FT_THREADS) relh_t[(i >> 6) * FT_TS + (i & 63)] = rel_h[i];
    for (int i = tid; i < (2 * gw - 1) * 64; i += FT_THREADS) relw_t[(i >> 6) * FT_TS + (i & 63)] = rel_w[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t T_S = tmem, T_DP = tmem + 128, T_DK = tmem + 256, T_DV = tmem + 384;     // dQ temp reuses T_S
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;

  // per-thread partial sums of the rel-pos table gradients: outputs e = tid, tid+128, ... over [(2gh-1) + (2gw-1)] x 64
  const int rows_h = 2 * gh - 1, rows_w = 2 * gw - 1;
  float relacc[64];            // thread r < rows_h (+32 for w): its row of the rel-pos table gradients, accumulated over query tiles
#pragma unroll
  for (int i = 0; i < 64; ++i) relacc[i] = 0.f;

  for (int q0 = 0, qt = 0; q0 < N; q0 += 128, ++qt) {
    ft_load_rows(Qs, base, C3, q0, 128, N);
    ft_load_rows(Gs, dout + (size_t)b * N * C + n * 64, C, q0, 128, N);
    const int q = q0 + tid;
    const bool qvalid = q < N;
    const int qy = qvalid ? q / gw : 0, qx = qvalid ? q % gw : 0;
    {   // D = dO . O, lse
      float dsum = 0.f;
      if (qvalid) {
        const size_t off = ((size_t)b * N + q) * C + n * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 ou = *reinterpret_cast<const uint4*>(out + off + 8 * c);
          const uint4 gu = *reinterpret_cast<const uint4*>(dout + off + 8 * c);
          const uint32_t ow[4] = {ou.x, ou.y, ou.z, ou.w}, gw4[4] = {gu.x, gu.y, gu.z, gu.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 a = unpack_bf16x2(ow[t]), g2 = unpack_bf16x2(gw4[t]);
            dsum += a.x * g2.x + a.y * g2.y;
          }
        }
      }
      D_s[tid] = dsum;
      lse_s[tid] = qvalid ? lse[((size_t)b * nH + n) * N + q] : 0.f;
    }
    for (int k = 0; k < 17; ++k) { dSh[tid * 17 + k] = 0.f; dSw[tid * 17 + k] = 0.f; }
    fence_proxy_async_smem();
    __syncthreads();
    float* rh = rh_s + tid * 17;
    float* rw = rw_s + tid * 17;
    float rhr[GG], rwr[GG], dShr[GG], dSwr[GG];
#pragma unroll
    for (int k = 0; k < GG; ++k) { dShr[k] = 0.f; dSwr[k] = 0.f; }
    if constexpr (G > 0) {      // rel-pos terms of the tile's rows as an MMA into the columns S will use (see the forward kernel)
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          tc_mma_tiles<false, false>(T_S, smem_u32(Qs), 0, smem_u32(Rhi), 0, 128, 64, 64, false);
          tc_mma_tiles<false, false>(T_S, smem_u32(Qs), 0, smem_u32(Rlo), 0, 128, 64, 64, true);
          umma_commit(mbar);
        }
        __syncwarp();
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      tc_fence_after();
      {
        uint32_t vh[32], vw[32];
        tmem_ld_32x32(T_S + lane_base, vh);
        tmem_ld_32x32(T_S + lane_base + 32, vw);
        tmem_ld_wait();
        ft_pick_terms<GG>(vh, qy, rhr);
        ft_pick_terms<GG>(vw, qx, rwr);
      }
      tc_fence_before();
      __syncthreads();
    } else if (use_rel) {
      ft_rel_terms(Qs, relh_t, relw_t, tid, qy, qx, gh, gw, rh, rw);
    }
    float dq[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) dq[d] = 0.f;

#pragma unroll
    for (int h = 0; h < n_halves; ++h) {
      const int k0 = h * 128;
      const int nk16 = min(128, N16 - k0);                 // keys of this half, multiple of 16
      if (warp == 0) {      // warp-uniform issue: descriptors stay in uniform registers (no ELECT/R2UR waterfall per MMA)
        tc_fence_after();
        if (elect_one()) {
          tc_mma_tiles<false, false>(T_S, smem_u32(Qs), 0, smem_u32(Ks) + h * FT_TILE, 0, 128, nk16, 64, false);
          tc_mma_tiles<false, false>(T_DP, smem_u32(Gs), 0, smem_u32(Vs) + h * FT_TILE, 0, 128, nk16, 64, false);
          umma_commit(mbar);
        }
        __syncwarp();
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      tc_fence_after();
      {
        const float l = lse_s[tid], D = D_s[tid];
        int jy = k0 / gw, jx = k0 % gw;
        const int n_chunks = (nk16 + 31) / 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float pv[32], dv[32];
          if (c < n_chunks) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(T_S + lane_base + c * 32, r0);
            tmem_ld_32x32(T_DP + lane_base + c * 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int j = k0 + c * 32 + e;
              float p = 0.f, ds = 0.f;
              if constexpr (G > 0) {
                if (j < G * G) {          // j is a compile-time constant here (h, c, e unrolled)
                  const float s = __uint_as_float(r0[e]) + rhr[(j / GG) % GG] + rwr[j % GG];
                  p = qvalid ? __expf(scale * s - l) : 0.f;
                  ds = p * (__uint_as_float(r1[e]) - D);
                  dShr[(j / GG) % GG] += ds;
                  dSwr[j % GG] += ds;
                }
              } else if (j < N && c * 32 + e < nk16) {
                float s = __uint_as_float(r0[e]);
                if (use_rel) s += rh[jy] + rw[jx];
                p = qvalid ? __expf(scale * s - l) : 0.f;
                ds = p * (__uint_as_float(r1[e]) - D);
                if (use_rel) { dSh[tid * 17 + jy] += ds; dSw[tid * 17 + jx] += ds; }
                if (++jx == gw) { jx = 0; ++jy; }
              }
              pv[e] = p;
              dv[e] = ds;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) { pv[e] = 0.f; dv[e] = 0.f; }
          }
          uint8_t* pa = Pt + (c >> 1) * FT_TILE;
          uint8_t* sa = St + (c >> 1) * FT_TILE;
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
      if (warp == 0) {      // warp-uniform issue: descriptors stay in uniform registers (no ELECT/R2UR waterfall per MMA)
        tc_fence_after();
        if (elect_one()) {
          // dQ_half = dS K_h (temp in T_S) ; dK_h += dS^T Q ; dV_h += P^T dO     (M = 128 keys/queries, N = 64, K = 128)
          tc_mma_tiles<false, true>(T_S, smem_u32(St), FT_TILE, smem_u32(Ks) + h * FT_TILE, 0, 128, 64, 128, false);
          tc_mma_tiles<true, true>(T_DK + 64 * h, smem_u32(St), FT_TILE, smem_u32(Qs), 0, 128, 64, 128, qt > 0);
          tc_mma_tiles<true, true>(T_DV + 64 * h, smem_u32(Pt), FT_TILE, smem_u32(Gs), 0, 128, 64, 128, qt > 0);
          umma_commit(mbar);
        }
        __syncwarp();
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      tc_fence_after();
      {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(T_S + lane_base, r0);
        tmem_ld_32x32(T_S + lane_base + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) { dq[d] += __uint_as_float(r0[d]); dq[32 + d] += __uint_as_float(r1[d]); }
      }
      tc_fence_before();
      __syncthreads();          // T_S, P and dS tiles are reused by the next half
      tc_fence_after();
    }

    if constexpr (G > 0) {
      // W[q][t] = dSh[q][qy - t + G - 1] (t < 2G-1), W[q][32 + t] = dSw[q][qx - t + G - 1]: the dS row / column sums spread over the table rows
      // they multiply.  Two products on the tensor core:  dq += W R  (W as bf16 hi + lo tiles, R as hi + lo: exact to 2^-17) and the table
      // gradients  dR[t][:] = sum_q W[q][t] q[q][:]  (W hi only, as before).
      {
        float wh[32], ww[32];
        ft_spread_terms<GG>(dShr, qy, wh);
        ft_spread_terms<GG>(dSwr, qx, ww);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int t = 8 * c + 2 * k;
            const float v0 = qvalid ? (t < 32 ? wh[t] : ww[t - 32]) : 0.f, v1 = qvalid ? (t + 1 < 32 ? wh[t + 1] : ww[t + 1 - 32]) : 0.f;
            const float h0 = __bfloat162float(__float2bfloat16_rn(v0)), h1 = __bfloat162float(__float2bfloat16_rn(v1));
            h[k] = pack_bf16x2(h0, h1);
            l[k] = pack_bf16x2(v0 - h0, v1 - h1);
          }
          *reinterpret_cast<uint4*>(Pt + tile_chunk_off(tid, c)) = make_uint4(h[0], h[1], h[2], h[3]);                 // P tile is free
          *reinterpret_cast<uint4*>(Pt + FT_TILE + tile_chunk_off(tid, c)) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          // table gradients: rows 64..127 of the result (the lo words as extra "table rows") are never read
          tc_mma_tiles<true, true>(T_S + 64, smem_u32(Pt), FT_TILE, smem_u32(Qs), 0, 128, 64, 128, false);
          tc_mma_tiles<false, true>(T_S, smem_u32(Pt), 0, smem_u32(Rhi), 0, 128, 64, 64, false);
          tc_mma_tiles<false, true>(T_S, smem_u32(Pt), 0, smem_u32(Rlo), 0, 128, 64, 64, true);
          tc_mma_tiles<false, true>(T_S, smem_u32(Pt) + FT_TILE, 0, smem_u32(Rhi), 0, 128, 64, 64, true);
          umma_commit(mbar);
        }
        __syncwarp();
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      tc_fence_after();
      {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(T_S + lane_base, r0);
        tmem_ld_32x32(T_S + lane_base + 32, r1);
        tmem_ld_wait();
        if (qvalid) {
          __nv_bfloat16* dst = dqkv + ((size_t)b * N + q) * C3 + n * 64;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u;
            u.x = pack_bf16x2(scale * (dq[8 * c] + __uint_as_float(r0[8 * c])), scale * (dq[8 * c + 1] + __uint_as_float(r0[8 * c + 1])));
            u.y = pack_bf16x2(scale * (dq[8 * c + 2] + __uint_as_float(r0[8 * c + 2])), scale * (dq[8 * c + 3] + __uint_as_float(r0[8 * c + 3])));
            u.z = pack_bf16x2(scale * (dq[8 * c + 4] + __uint_as_float(r0[8 * c + 4])), scale * (dq[8 * c + 5] + __uint_as_float(r0[8 * c + 5])));
            u.w = pack_bf16x2(scale * (dq[8 * c + 6] + __uint_as_float(r0[8 * c + 6])), scale * (dq[8 * c + 7] + __uint_as_float(r0[8 * c + 7])));
            *reinterpret_cast<uint4*>(dst + 8 * c) = u;
            u.x = pack_bf16x2(scale * (dq[32 + 8 * c] + __uint_as_float(r1[8 * c])), scale * (dq[32 + 8 * c + 1] + __uint_as_float(r1[8 * c + 1])));
            u.y = pack_bf16x2(scale * (dq[32 + 8 * c + 2] + __uint_as_float(r1[8 * c + 2])), scale * (dq[32 + 8 * c + 3] + __uint_as_float(r1[8 * c + 3])));
            u.z = pack_bf16x2(scale * (dq[32 + 8 * c + 4] + __uint_as_float(r1[8 * c + 4])), scale * (dq[32 + 8 * c + 5] + __uint_as_float(r1[8 * c + 5])));
            u.w = pack_bf16x2(scale * (dq[32 + 8 * c + 6] + __uint_as_float(r1[8 * c + 6])), scale * (dq[32 + 8 * c + 7] + __uint_as_float(r1[8 * c + 7])));
            *reinterpret_cast<uint4*>(dst + 32 + 8 * c) = u;
          }
        }
      }
      if (warp < 2) {           // rows 0..63 of the table-gradient product
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(T_S + 64 + lane_base, r0);
        tmem_ld_32x32(T_S + 64 + lane_base + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) { relacc[d] += __uint_as_float(r0[d]); relacc[32 + d] += __uint_as_float(r1[d]); }
      }
      tc_fence_before();
    } else {
    // dq = scale * (dS K + sum_k dSh[k] Rh[qy-k+gh-1] + sum_k dSw[k] Rw[qx-k+gw-1])
    if (qvalid) {
      if (use_rel) {
        for (int k = 0; k < gh; ++k) {
          const float ch = dSh[tid * 17 + k];
          const float* th = relh_t + (qy - k + gh - 1) * FT_TS;
#pragma unroll
          for (int d = 0; d < 64; ++d) dq[d] += ch * th[d];
        }
        for (int k = 0; k < gw; ++k) {
          const float cw = dSw[tid * 17 + k];
          const float* tw = relw_t + (qx - k + gw - 1) * FT_TS;
#pragma unroll
          for (int d = 0; d < 64; ++d) dq[d] += cw * tw[d];
        }
      }
      __nv_bfloat16* dst = dqkv + ((size_t)b * N + q) * C3 + n * 64;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u;
        u.x = pack_bf16x2(scale * dq[8 * c], scale * dq[8 * c + 1]); u.y = pack_bf16x2(scale * dq[8 * c + 2], scale * dq[8 * c + 3]);
        u.z = pack_bf16x2(scale * dq[8 * c + 4], scale * dq[8 * c + 5]); u.w = pack_bf16x2(scale * dq[8 * c + 6], scale * dq[8 * c + 7]);
        *reinterpret_cast<uint4*>(dst + 8 * c) = u;
      }
    }
    // rel-pos table gradients of this query tile as one more MMA:  dR[r][:] = sum_q W[q][r] q[q][:]  with
    // W[q][r] = dSh[q][qy - r + gh - 1] for r < 2gh-1 and W[q][32 + r] = dSw[q][qx - r + gw - 1] (zero when out of range).
    if (use_rel) {
      float wv[64];
#pragma unroll
      for (int r = 0; r < 64; ++r) {
        float v = 0.f;
        if (qvalid) {
          if (r < 32) { const int k = qy - r + gh - 1; if (r < rows_h && k >= 0 && k < gh) v = dSh[tid * 17 + k]; }
          else { const int k = qx - (r - 32) + gw - 1; if (r - 32 < rows_w && k >= 0 && k < gw) v = dSw[tid * 17 + k]; }
        }
        wv[r] = v;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u;
        u.x = pack_bf16x2(wv[8 * c], wv[8 * c + 1]); u.y = pack_bf16x2(wv[8 * c + 2], wv[8 * c + 3]);
        u.z = pack_bf16x2(wv[8 * c + 4], wv[8 * c + 5]); u.w = pack_bf16x2(wv[8 * c + 6], wv[8 * c + 7]);
        *reinterpret_cast<uint4*>(Pt + tile_chunk_off(tid, c)) = u;                       // P tile is free: reuse it for W
        *reinterpret_cast<uint4*>(Pt + FT_TILE + tile_chunk_off(tid, c)) = make_uint4(0, 0, 0, 0);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncthreads();
      if (warp == 0) {      // warp-uniform issue: descriptors stay in uniform registers (no ELECT/R2UR waterfall per MMA)
        tc_fence_after();
        if (elect_one()) {
          tc_mma_tiles<true, true>(T_S + 64, smem_u32(Pt), FT_TILE, smem_u32(Qs), 0, 128, 64, 128, false);
          umma_commit(mbar);
        }
        __syncwarp();
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      tc_fence_after();
      if (warp < 2) {           // rows 0..63 of the result
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(T_S + 64 + lane_base, r0);
        tmem_ld_32x32(T_S + 64 + lane_base + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) { relacc[d] += __uint_as_float(r0[d]); relacc[32 + d] += __uint_as_float(r1[d]); }
      }
      tc_fence_before();
    }
    }
    __syncthreads();            // Q / dO tiles and dSh / dSw are rewritten by the next query tile
  }

  // ---- dK, dV rows (thread r = key 128 h + r)
  for (int h = 0; h < n_halves; ++h) {
    const int j = 128 * h + tid;
    uint32_t r0[32], r1[32];
    __nv_bfloat16* dst = dqkv + ((size_t)b * N + j) * C3 + n * 64;
#pragma unroll
    for (int part = 0; part < 2; ++part) {      // 0: dK (x scale), 1: dV
      const uint32_t t0 = (part == 0 ? T_DK : T_DV) + 64 * h + lane_base;
      tmem_ld_32x32(t0, r0);
      tmem_ld_32x32(t0 + 32, r1);
      tmem_ld_wait();
      if (j < N) {
        const float f = part == 0 ? scale : 1.0f;
        __nv_bfloat16* o = dst + (part == 0 ? C : 2 * C);
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
  if (use_rel && tid < 64) {
    float* dstp = nullptr;
    if (tid < rows_h) dstp = d_rel_h + (size_t)tid * 64;
    else if (tid >= 32 && tid - 32 < rows_w) dstp = d_rel_w + (size_t)(tid - 32) * 64;
    if (dstp != nullptr) {
#pragma unroll
      for (int d = 0; d < 64; ++d) atomicAdd(dstp + d, scale * relacc[d]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int launch_full_attn_bwd_tc(const void* qkv, const float* rel_h, const float* rel_w, const float* lse, const void* out, const void* dout,
                            void* dqkv, float* d_rel_h, float* d_rel_w, int B, int gh, int gw, int C, int nH, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(full_attn_bwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FTB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(full_attn_bwd_tc_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, FTB_SMEM);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "full_attn_bwd_tc smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  auto kern = (gh == 14 && gw == 14 && rel_h != nullptr) ? full_attn_bwd_tc_kernel<14> : full_attn_bwd_tc_kernel<0>;
  (void)launch_k(kern, B * nH, FT_THREADS, FTB_SMEM, st,
      reinterpret_cast<const __nv_bfloat16*>(qkv), rel_h, rel_w, lse, reinterpret_cast<const __nv_bfloat16*>(out),
      reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<__nv_bfloat16*>(dqkv), d_rel_h, d_rel_w, gh * gw, gh, gw, C, nH,
      rel_h != nullptr);
  return check_launch("full_attn_bwd_tc_kernel");
}

}  // namespace mtp
