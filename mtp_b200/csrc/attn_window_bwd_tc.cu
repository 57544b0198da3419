// RVSA window attention backward on tcgen05 tensor cores (autograd of attn_window_tc.cu; [V]:372-428).
//
// Same tile shape as the forward: two (window, head) problems stacked on the 128 rows of one UMMA tile.  256 threads: thread
// `row` (0..127) owns accumulator row `row`; its helper `row + 128` (same TMEM lane quarter) takes the second half of the
// row's columns in the phases that split (rel-pos terms, dq, TMEM read-out) and joins the cooperative phases.
//   gather : q, dO rows and the bilinearly blended k~, v~ rows -> bf16 swizzled tiles
//   MMA 1  : S  = Q K~^T , dP = dO V~^T                              (M = N = 128, K = 64)
//   rows   : thread r: P = exp(S - lse), D = sum P dP, dS = P (dP - D); P and dS written block-diagonally (bf16);
//            row sums of dS per key row / key column (rel-pos), bias-table partials via shared-memory atomics
//   MMA 2  : dQ = dS K~ ,  dK~ = dS^T Q ,  dV~ = P^T dO              (M = 128, N = 64, K = 128; transposes via MN-major A)
//   tail   : dq rows (+ rel-pos terms) -> dqkv; dk~ / dv~ rows are staged TMEM -> smem, then 8 lanes per sample row (16 B of
//            k and of v each, all 4 taps' loads in flight at once) scatter them through the bilinear taps with coalesced
//            red.global.add.v4.f32 into the fp32 scratch; tap dot-products -> d(coords) -> d(ox, oy, sx, sy, theta)
// Partials for rel_pos_h/w are written per CTA (both heads summed), bias-table partials per (window, head).
#include "common.h"
#include "ptx.cuh"
#include "rvsa_geom.cuh"
#include "tc_tile.cuh"

namespace mtp {

constexpr int WB_THREADS = 256;
constexpr int WB_TILE = 128 * 128;
// Q | K~ | V~ | dO | PS (2 atoms: P first, then dS) | rel tables | bias tables | coords | dSh,dSw | rw | gxy | red | mbar | slot
// 110 KB and 256 TMEM columns per CTA, 128 registers per thread: TWO CTAs fit one SM, so the latency chains of one (gather -> MMA -> row
// phase -> MMA -> scatter) are covered by the other (r1: 185 KB / 512 columns, one CTA per SM, 256 CTAs = 1.73 waves).
// W (the rel-table weights) is written over the dead V~ tile; its second atom (rows 64..127 of the product, never read) aliases dO.
// (dyn + 1 KB reserve) x 2 <= 228 KB  =>  dyn <= 115,712 B: the helper hand-over buffer aliases dSh, the coordinate gradients alias dSw
constexpr int WB_SMEM = 6 * WB_TILE + (2 * 13 * 64 + 2 * 169 + 2 * 98 + 2 * 128 * 8 + 64) * 4 + 64;
static_assert(2 * (WB_SMEM + 1024) <= 228 * 1024, "two CTAs of the RVSA backward must fit one SM");

__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float tile_read(const uint8_t* tile, int row, int d) {     // bf16 element (row, d) of a swizzled tile
  return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(tile + tile_chunk_off(row, d >> 3) + (d & 7) * 2));
}

__global__ void __launch_bounds__(WB_THREADS, 2)
rvsa_attn_bwd_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ params, const float* __restrict__ rel_h,
                        const float* __restrict__ rel_w, const float* __restrict__ bias_table, const float* __restrict__ lse,
                        const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dkv,
                        float* __restrict__ dparams, float* __restrict__ part_rel, float* __restrict__ part_table, const RvsaGeom g) {
  MTP_PDL_ENTRY();
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* Qs = sm;
  uint8_t* Ks = Qs + WB_TILE;
  uint8_t* Vs = Ks + WB_TILE;
  uint8_t* Gs = Vs + WB_TILE;
  uint8_t* Pt = Gs + WB_TILE;                 // 2 atoms: P (block-diagonal) for dV~ = P^T dO, then overwritten by dS for dQ / dK~ / dR
  uint8_t* St = Pt;
  uint8_t* Wt = Vs;                           // row q, column r = weight of q in d rel table row r; written after MMA 1 has consumed V~
  float* relt = reinterpret_cast<float*>(Pt + 2 * WB_TILE);     // [2][13][64]
  float* tabs = relt + 2 * 13 * 64;           // [2][169]
  float* cpx = tabs + 2 * 169;                // [98]
  float* cpy = cpx + 98;
  float* dSh = cpy + 98;                      // [128][8]
  float* dSw = dSh + 128 * 8;
  float* rwS = dSh;                           // [128][8] rel-pos column terms handed from helper to owner (read back before dSh is written)
  float* gxy = dSw;                           // [128][2] d(sample coords): written by the scatter, after the last reader of dSw
  float* red = dSw + 128 * 8;                 // [64]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(red + 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half_heads = g.nH >> 1;
  const int hp = blockIdx.x % half_heads;
  const int bw = blockIdx.x / half_heads;
  const int nwin = g.nh * g.nw;
  const int b = bw / nwin, win = bw % nwin;
  const int wy = win / g.nw, wx = win % g.nw;
  const int C = g.C, C3 = 3 * g.C;
  const float scale = 0.125f;

  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (tid == 32) {
    mbar_init(mbar, 1);
    fence_barrier_init();
  }
  // padding rows (49..63 of each problem) of the four operand tiles; every other row is written by the gather.
  for (int i = tid; i < 4 * 30 * 8; i += WB_THREADS) {
    const int tile = i / 240, rr = (i % 240) >> 3, c = i & 7;
    const int row = rr < 15 ? NTOK + rr : 64 + NTOK + (rr - 15);
    *reinterpret_cast<uint4*>(Qs + tile * WB_TILE + tile_chunk_off(row, c)) = make_uint4(0, 0, 0, 0);
  }
  {
    // Prologue loads (rel-pos tables, the two heads' bias-table columns, the sampling parameters) are issued as ONE batch and stored afterwards:
    // written as `smem[i] = global[i]` loops they stayed in program order -- seven dependent L2 round trips on every CTA's critical path
    // (ncu source view of the backward: the table store waiting on its load was the kernel's top stall).
    static_assert(WB_THREADS == 256, "batch sizes below assume 256 threads");
    constexpr int NT = (2 * 13 * 64 + 255) / 256;      // 7 table elements per thread
    float tv[NT], bv[2], pr[5];
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const int i = tid + k * 256;
      tv[k] = i < 2 * 13 * 64 ? (i < 13 * 64 ? __ldg(rel_h + i) : __ldg(rel_w + i - 13 * 64)) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = tid + k * 256;
      bv[k] = i < 2 * 169 ? __ldg(bias_table + (i % 169) * g.nH + 2 * hp + i / 169) : 0.f;
    }
    if (tid < 98) {
      const float* prm = params + ((size_t)bw * g.nH + 2 * hp + tid / NTOK) * 8;
#pragma unroll
      for (int k = 0; k < 5; ++k) pr[k] = prm[k];
    }
#pragma unroll
    for (int k = 0; k < NT; ++k) {      // chunk-swizzled rows (tab_chunk_off): lanes of a warp read different rows
      const int i = tid + k * 256;
      if (i < 2 * 13 * 64) {
        const int t = i / (13 * 64), r = (i % (13 * 64)) >> 6, d = i & 63;
        relt[t * 13 * 64 + tab_chunk_off(r, d >> 2) + (d & 3)] = tv[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (tid + k * 256 < 2 * 169) tabs[tid + k * 256] = bv[k];
    if (tid < 98) {
      const int j = tid % NTOK;
      float px, py;
      rvsa_sample_coord(g, wy, wx, j / WS, j % WS, pr[0], pr[1], pr[2], pr[3], pr[4], px, py);
      cpx[tid] = px;
      cpy[tid] = py;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // S and dP are consumed into registers by the row phase; the four products of the second round reuse their columns
  const uint32_t T_S = tmem, T_DP = tmem + 128, T_DV = tmem, T_DQ = tmem + 64, T_DK = tmem + 128, T_DR = tmem + 192;

  // ---- gather: 8 lanes per row, 16 B per lane (see the forward kernel)
  for (int i = tid >> 3; i < 2 * NTOK; i += WB_THREADS / 8) {
    const int c = tid & 7;
    const int p = i / NTOK, j = i % NTOK, r = 64 * p + j;
    const int head = 2 * hp + p;
    const __nv_bfloat16* qkv_b = qkv + (size_t)b * g.h * g.w * C3 + head * HD + c * 8;
    const uint32_t soff = tile_chunk_off(r, c);
    const int y = wy * WS + j / WS - g.pt, x = wx * WS + j % WS - g.pl;
    if (y >= 0 && y < g.h && x >= 0 && x < g.w) {
      const size_t t = (size_t)(b * g.h + y) * g.w + x;
      *reinterpret_cast<uint4*>(Qs + soff) = *reinterpret_cast<const uint4*>(qkv + t * C3 + head * HD + c * 8);
      *reinterpret_cast<uint4*>(Gs + soff) = *reinterpret_cast<const uint4*>(dout + t * C + head * HD + c * 8);
    } else {
      *reinterpret_cast<uint4*>(Qs + soff) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(Gs + soff) = make_uint4(0, 0, 0, 0);
    }
    const float px = cpx[i], py = cpy[i];
    const float fx0 = floorf(px), fy0 = floorf(py);
    const float ax = px - fx0, ay = py - fy0;
    const int x0 = (int)fx0 - g.pl, y0 = (int)fy0 - g.pt;
    uint4 kt[4], vt[4];
    float wgt[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
      const bool ok = xx >= 0 && xx < g.w && yy >= 0 && yy < g.h;
      wgt[t] = ok ? ((t & 1) ? ax : 1.f - ax) * ((t >> 1) ? ay : 1.f - ay) : 0.f;
      const __nv_bfloat16* src = qkv_b + (size_t)(ok ? yy * g.w + xx : 0) * C3;
      kt[t] = *reinterpret_cast<const uint4*>(src + C);
      vt[t] = *reinterpret_cast<const uint4*>(src + 2 * C);
    }
    float ka[8], va[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { ka[e] = 0.f; va[e] = 0.f; }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const uint32_t kw4[4] = {kt[t].x, kt[t].y, kt[t].z, kt[t].w}, vw4[4] = {vt[t].x, vt[t].y, vt[t].z, vt[t].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 kf = unpack_bf16x2(kw4[e]), vf = unpack_bf16x2(vw4[e]);
        ka[2 * e] += wgt[t] * kf.x; ka[2 * e + 1] += wgt[t] * kf.y;
        va[2 * e] += wgt[t] * vf.x; va[2 * e + 1] += wgt[t] * vf.y;
      }
    }
    uint4 u;
    u.x = pack_bf16x2(ka[0], ka[1]); u.y = pack_bf16x2(ka[2], ka[3]); u.z = pack_bf16x2(ka[4], ka[5]); u.w = pack_bf16x2(ka[6], ka[7]);
    *reinterpret_cast<uint4*>(Ks + soff) = u;
    u.x = pack_bf16x2(va[0], va[1]); u.y = pack_bf16x2(va[2], va[3]); u.z = pack_bf16x2(va[4], va[5]); u.w = pack_bf16x2(va[6], va[7]);
    *reinterpret_cast<uint4*>(Vs + soff) = u;
  }
  fence_proxy_async_smem();
  __syncthreads();

  // ---- S = Q K~^T, dP = dO V~^T
  if (warp == 0) {      // warp-uniform issue: descriptors stay in uniform registers (no ELECT/R2UR waterfall per MMA)
    tc_fence_after();
    if (elect_one()) {
      tc_mma_tiles<false, false>(T_S, smem_u32(Qs), 0, smem_u32(Ks), 0, 128, 128, 64, false);
      tc_mma_tiles<false, false>(T_DP, smem_u32(Gs), 0, smem_u32(Vs), 0, 128, 128, 64, false);
      umma_commit(mbar);
    }
    __syncwarp();
  }
  const int row = tid & 127;
  const bool helper = tid >= 128;
  const int p = row >> 6, q = row & 63;
  const bool qvalid = q < NTOK;
  const int qy = q / WS, qx = q % WS;
  const int head = 2 * hp + p;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;      // TMEM lane quarter of this thread's row
  // rel-pos terms of the row: the owner computes the height terms, its helper the width terms (handed over through smem)
  float rh[WS], rw[WS];
#pragma unroll
  for (int k = 0; k < WS; ++k) { rh[k] = 0.f; rw[k] = 0.f; }
  if (qvalid) {
    const float* tbase = relt + (helper ? 13 * 64 : 0);
    const int qq = helper ? qx : qy;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = *reinterpret_cast<const uint4*>(Qs + tile_chunk_off(row, c));
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
      float qv[8];
#pragma unroll
      for (int t = 0; t < 4; ++t) { const float2 f = unpack_bf16x2(w4[t]); qv[2 * t] = f.x; qv[2 * t + 1] = f.y; }
#pragma unroll
      for (int k = 0; k < WS; ++k) {
        const int tr = qq - k + WS - 1;
        const float4 h0 = *reinterpret_cast<const float4*>(tbase + tab_chunk_off(tr, 2 * c));
        const float4 h1 = *reinterpret_cast<const float4*>(tbase + tab_chunk_off(tr, 2 * c + 1));
        rh[k] += qv[0] * h0.x + qv[1] * h0.y + qv[2] * h0.z + qv[3] * h0.w + qv[4] * h1.x + qv[5] * h1.y + qv[6] * h1.z + qv[7] * h1.w;
      }
    }
  }
  if (helper) {
#pragma unroll
    for (int k = 0; k < WS; ++k) rwS[row * 8 + k] = rh[k];
  }
  __syncthreads();
  if (!helper) {
#pragma unroll
    for (int k = 0; k < WS; ++k) rw[k] = rwS[row * 8 + k];
  }
  mbar_wait(mbar, 0);
  tc_fence_after();

  // ---- P, dS of this thread's row
  float sh[WS], sw[WS];
#pragma unroll
  for (int k = 0; k < WS; ++k) { sh[k] = 0.f; sw[k] = 0.f; }
  uint4 dsp[8];                           // this row's dS, packed bf16 (owner threads)
#pragma unroll
  for (int c = 0; c < 8; ++c) dsp[c] = make_uint4(0, 0, 0, 0);
  if (!helper) {
    uint32_t r0[32], r1[32];
    tmem_ld_32x32(T_S + lane_base + 64 * p, r0);
    tmem_ld_32x32(T_S + lane_base + 64 * p + 32, r1);
    tmem_ld_wait();
    float pr[NTOK];
    const float l = qvalid ? lse[((size_t)bw * g.nH + head) * NTOK + q] : 0.f;
    const float* tab = tabs + p * 169;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      const float acc = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]);
      const int jy = j / WS, jx = j % WS;
      const float s = scale * acc + rh[jy] + rw[jx] + (qvalid ? tab[(qy - jy + WS - 1) * (2 * WS - 1) + (qx - jx + WS - 1)] : 0.f);
      pr[j] = qvalid ? __expf(s - l) : 0.f;
    }
    tmem_ld_32x32(T_DP + lane_base + 64 * p, r0);
    tmem_ld_32x32(T_DP + lane_base + 64 * p + 32, r1);
    tmem_ld_wait();
    float D = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) D += pr[j] * __uint_as_float(j < 32 ? r0[j] : r1[j - 32]);
    float ds[NTOK];
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      ds[j] = pr[j] * (__uint_as_float(j < 32 ? r0[j] : r1[j - 32]) - D);
      sh[j / WS] += ds[j];
      sw[j % WS] += ds[j];
    }
#pragma unroll
    for (int k = 0; k < WS; ++k) { dSh[row * 8 + k] = sh[k]; dSw[row * 8 + k] = sw[k]; }
    {   // row q of W: column r < 13 -> dSh[q][qy - r + 6], column 13 + r -> dSw[q][qx - r + 6] (0 when out of range)
      float wv[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        float v = 0.f;
        if (r < 13) { const int k = qy - r + WS - 1; if (qvalid && k >= 0 && k < WS) v = dSh[row * 8 + k]; }
        else if (r < 26) { const int k = qx - (r - 13) + WS - 1; if (qvalid && k >= 0 && k < WS) v = dSw[row * 8 + k]; }
        wv[r] = v;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (c < 4) {
          u.x = pack_bf16x2(wv[8 * c], wv[8 * c + 1]); u.y = pack_bf16x2(wv[8 * c + 2], wv[8 * c + 3]);
          u.z = pack_bf16x2(wv[8 * c + 4], wv[8 * c + 5]); u.w = pack_bf16x2(wv[8 * c + 6], wv[8 * c + 7]);
        }
        *reinterpret_cast<uint4*>(Wt + tile_chunk_off(row, c)) = u;
      }
    }
    uint8_t* p_mine = Pt + p * WB_TILE;
    uint8_t* p_other = Pt + (1 - p) * WB_TILE;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float vp[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = c * 8 + e;
        vp[e] = j < NTOK ? pr[j < NTOK ? j : 0] : 0.f;
      }
      uint4 u;
      u.x = pack_bf16x2(vp[0], vp[1]); u.y = pack_bf16x2(vp[2], vp[3]); u.z = pack_bf16x2(vp[4], vp[5]); u.w = pack_bf16x2(vp[6], vp[7]);
      *reinterpret_cast<uint4*>(p_mine + tile_chunk_off(row, c)) = u;
      *reinterpret_cast<uint4*>(p_other + tile_chunk_off(row, c)) = make_uint4(0, 0, 0, 0);
    }
    // dS of this row waits in packed form until the P tile has been consumed
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float vs[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = c * 8 + e;
        vs[e] = j < NTOK ? ds[j < NTOK ? j : 0] : 0.f;
      }
      dsp[c].x = pack_bf16x2(vs[0], vs[1]); dsp[c].y = pack_bf16x2(vs[2], vs[3]); dsp[c].z = pack_bf16x2(vs[4], vs[5]); dsp[c].w = pack_bf16x2(vs[6], vs[7]);
    }
  }
  tc_fence_before();
  fence_proxy_async_smem();
  __syncthreads();

  // ---- dV~ = P^T dO  (the P tile is then handed over to dS)
  if (warp == 0) {      // warp-uniform issue: descriptors stay in uniform registers (no ELECT/R2UR waterfall per MMA)
    tc_fence_after();
    if (elect_one()) {
      tc_mma_tiles<true, true>(T_DV, smem_u32(Pt), WB_TILE, smem_u32(Gs), 0, 128, 64, 128, false);
      umma_commit(mbar);
    }
    __syncwarp();
  }
  mbar_wait(mbar, 1);
  tc_fence_after();
  if (!helper) {
    uint8_t* s_mine = St + p * WB_TILE;
    uint8_t* s_other = St + (1 - p) * WB_TILE;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      *reinterpret_cast<uint4*>(s_mine + tile_chunk_off(row, c)) = dsp[c];
      *reinterpret_cast<uint4*>(s_other + tile_chunk_off(row, c)) = make_uint4(0, 0, 0, 0);
    }
  }
  tc_fence_before();
  fence_proxy_async_smem();
  __syncthreads();

  // ---- dQ = dS K~ ; dK~ = dS^T Q ; d rel tables = W^T Q
  if (warp == 0) {      // warp-uniform issue: descriptors stay in uniform registers (no ELECT/R2UR waterfall per MMA)
    tc_fence_after();
    if (elect_one()) {
      tc_mma_tiles<false, true>(T_DQ, smem_u32(St), WB_TILE, smem_u32(Ks), 0, 128, 64, 128, false);
      tc_mma_tiles<true, true>(T_DK, smem_u32(St), WB_TILE, smem_u32(Qs), 0, 128, 64, 128, false);
      tc_mma_tiles<true, true>(T_DR, smem_u32(Wt), WB_TILE, smem_u32(Qs), 0, 128, 64, 128, false);     // d rel tables = W^T Q (both heads)
      umma_commit(mbar);
    }
    __syncwarp();
  }
  // meanwhile: bias-table partials, one output per (head, displacement): sum of dS over the pairs with that displacement
  for (int i = tid; i < 2 * 169; i += WB_THREADS) {
    const int pp = i / 169, idx = i % 169;
    const int dy = idx / 13 - (WS - 1), dx = idx % 13 - (WS - 1);
    const uint8_t* atom = St + pp * WB_TILE;
    float s = 0.f;
    for (int qy2 = max(0, dy); qy2 < min(WS, WS + dy); ++qy2)
      for (int qx2 = max(0, dx); qx2 < min(WS, WS + dx); ++qx2)
        s += tile_read(atom, 64 * pp + qy2 * WS + qx2, (qy2 - dy) * WS + (qx2 - dx));
    part_table[((size_t)bw * g.nH + 2 * hp + pp) * 169 + idx] = s;
  }
  mbar_wait(mbar, 0);
  tc_fence_after();

  if (warp == 0) {          // rows 0..12 = d rel_pos_h partial, rows 13..25 = d rel_pos_w partial (both heads of this CTA)
    uint32_t r0[32], r1[32];
    tmem_ld_32x32(T_DR + lane_base, r0);
    tmem_ld_32x32(T_DR + lane_base + 32, r1);
    tmem_ld_wait();
    if (tid < 26) {
      float* dst = part_rel + (size_t)blockIdx.x * (2 * (2 * WS - 1) * HD) + tid * HD;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(__uint_as_float(r0[4 * c]), __uint_as_float(r0[4 * c + 1]), __uint_as_float(r0[4 * c + 2]), __uint_as_float(r0[4 * c + 3]));
        *reinterpret_cast<float4*>(dst + 32 + 4 * c) = make_float4(__uint_as_float(r1[4 * c]), __uint_as_float(r1[4 * c + 1]), __uint_as_float(r1[4 * c + 2]), __uint_as_float(r1[4 * c + 3]));
      }
    }
  }
  const int hb = helper ? 1 : 0;          // which 32 of the row's 64 head dims this thread reads out of TMEM
  // ---- dq row -> dqkv
  {
    const int y = wy * WS + qy - g.pt, x = wx * WS + qx - g.pl;
    const bool tok_ok = qvalid && y >= 0 && y < g.h && x >= 0 && x < g.w;
    uint32_t r0[32];
    tmem_ld_32x32(T_DQ + lane_base + 32 * hb, r0);
    tmem_ld_wait();
    if (tok_ok) {
      float o[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = scale * __uint_as_float(r0[e]);
#pragma unroll
      for (int k = 0; k < WS; ++k) {
        const float shk = dSh[row * 8 + k], swk = dSw[row * 8 + k];
        const int trh = qy - k + WS - 1, trw = qx - k + WS - 1;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 a = *reinterpret_cast<const float4*>(relt + tab_chunk_off(trh, 8 * hb + e));
          const float4 c = *reinterpret_cast<const float4*>(relt + 13 * 64 + tab_chunk_off(trw, 8 * hb + e));
          o[4 * e] += shk * a.x + swk * c.x;
          o[4 * e + 1] += shk * a.y + swk * c.y;
          o[4 * e + 2] += shk * a.z + swk * c.z;
          o[4 * e + 3] += shk * a.w + swk * c.w;
        }
      }
      __nv_bfloat16* dst = dqkv + ((size_t)(b * g.h + y) * g.w + x) * C3 + head * HD + 32 * hb;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 u;
        u.x = pack_bf16x2(o[8 * c], o[8 * c + 1]); u.y = pack_bf16x2(o[8 * c + 2], o[8 * c + 3]);
        u.z = pack_bf16x2(o[8 * c + 4], o[8 * c + 5]); u.w = pack_bf16x2(o[8 * c + 6], o[8 * c + 7]);
        *reinterpret_cast<uint4*>(dst + 8 * c) = u;
      }
    }
  }

  // ---- dk~ (scaled) / dv~ rows: TMEM -> fp32 staging over the dead operand tiles, 256 B per row, 16-byte chunk c of row r at
  //      chunk (c & 8) | ((c ^ r) & 7) so that both the row-per-lane writes and the 8-lanes-per-row reads are conflict-free
  uint8_t* DKs = Qs;                     // 32 KB (Q and K~ tiles)
  uint8_t* DVs = Vs;                     // 32 KB (V~ and dO tiles)
  {
    uint32_t rk[32], rv[32];
    tmem_ld_32x32(T_DK + lane_base + 32 * hb, rk);
    tmem_ld_32x32(T_DV + lane_base + 32 * hb, rv);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t off = (uint32_t)row * 256u + (uint32_t)(((8 * hb) | ((c ^ row) & 7)) << 4);
      *reinterpret_cast<float4*>(DKs + off) = make_float4(scale * __uint_as_float(rk[4 * c]), scale * __uint_as_float(rk[4 * c + 1]),
                                                          scale * __uint_as_float(rk[4 * c + 2]), scale * __uint_as_float(rk[4 * c + 3]));
      *reinterpret_cast<float4*>(DVs + off) = make_float4(__uint_as_float(rv[4 * c]), __uint_as_float(rv[4 * c + 1]),
                                                          __uint_as_float(rv[4 * c + 2]), __uint_as_float(rv[4 * c + 3]));
    }
  }
  __syncthreads();

  // ---- scatter through the bilinear taps + coordinate gradients: 8 lanes per sample row; lane c8 owns head dims
  //      [4 c8, 4 c8 + 4) and [32 + 4 c8, 32 + 4 c8 + 4), so every warp-wide access covers whole 128-byte lines
  {
    const int c8 = tid & 7;
    constexpr int ROWS_PER_PASS = WB_THREADS / 8;
#pragma unroll 1
    for (int pass = 0; pass < (2 * NTOK + ROWS_PER_PASS - 1) / ROWS_PER_PASS; ++pass) {
      const int i = pass * ROWS_PER_PASS + (tid >> 3);
      const bool valid = i < 2 * NTOK;
      const int ii = valid ? i : 0;
      const int pp = ii / NTOK, j = ii % NTOK, r = 64 * pp + j, hd = 2 * hp + pp;
      const float px = cpx[ii], py = cpy[ii];
      const float fx0 = floorf(px), fy0 = floorf(py);
      const float ax = px - fx0, ay = py - fy0;
      const int x0 = (int)fx0 - g.pl, y0 = (int)fy0 - g.pt;
      const uint32_t offA = (uint32_t)r * 256u + (uint32_t)(((c8 ^ r) & 7) << 4), offB = offA + 128u;
      const float4 gkA = *reinterpret_cast<const float4*>(DKs + offA), gkB = *reinterpret_cast<const float4*>(DKs + offB);
      const float4 gvA = *reinterpret_cast<const float4*>(DVs + offA), gvB = *reinterpret_cast<const float4*>(DVs + offB);
      const __nv_bfloat16* qkv_b = qkv + (size_t)b * g.h * g.w * C3 + hd * HD + 4 * c8;
      float* dk_b = dkv + (size_t)b * g.h * g.w * 2 * C + hd * HD + 4 * c8;
      uint2 kA[4], kB[4], vA[4], vB[4];
      bool ok[4];
      size_t pix[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
        ok[t] = valid && xx >= 0 && xx < g.w && yy >= 0 && yy < g.h;
        pix[t] = ok[t] ? (size_t)(yy * g.w + xx) : 0;
        const __nv_bfloat16* src = qkv_b + pix[t] * C3;
        kA[t] = *reinterpret_cast<const uint2*>(src + C);
        kB[t] = *reinterpret_cast<const uint2*>(src + C + 32);
        vA[t] = *reinterpret_cast<const uint2*>(src + 2 * C);
        vB[t] = *reinterpret_cast<const uint2*>(src + 2 * C + 32);
      }
      float td[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 a0 = unpack_bf16x2(kA[t].x), a1 = unpack_bf16x2(kA[t].y), b0 = unpack_bf16x2(kB[t].x), b1 = unpack_bf16x2(kB[t].y);
        const float2 c0 = unpack_bf16x2(vA[t].x), c1 = unpack_bf16x2(vA[t].y), d0 = unpack_bf16x2(vB[t].x), d1 = unpack_bf16x2(vB[t].y);
        float d = gkA.x * a0.x + gkA.y * a0.y + gkA.z * a1.x + gkA.w * a1.y + gkB.x * b0.x + gkB.y * b0.y + gkB.z * b1.x + gkB.w * b1.y;
        d += gvA.x * c0.x + gvA.y * c0.y + gvA.z * c1.x + gvA.w * c1.y + gvB.x * d0.x + gvB.y * d0.y + gvB.z * d1.x + gvB.w * d1.y;
        td[t] = ok[t] ? d : 0.f;
        if (ok[t]) {
          const float wgt = ((t & 1) ? ax : 1.f - ax) * ((t >> 1) ? ay : 1.f - ay);
          float* dst = dk_b + pix[t] * 2 * C;
          red_add_f32x4(dst, wgt * gkA.x, wgt * gkA.y, wgt * gkA.z, wgt * gkA.w);
          red_add_f32x4(dst + 32, wgt * gkB.x, wgt * gkB.y, wgt * gkB.z, wgt * gkB.w);
          red_add_f32x4(dst + C, wgt * gvA.x, wgt * gvA.y, wgt * gvA.z, wgt * gvA.w);
          red_add_f32x4(dst + C + 32, wgt * gvB.x, wgt * gvB.y, wgt * gvB.z, wgt * gvB.w);
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        td[t] += __shfl_xor_sync(0xffffffffu, td[t], 1);
        td[t] += __shfl_xor_sync(0xffffffffu, td[t], 2);
        td[t] += __shfl_xor_sync(0xffffffffu, td[t], 4);
      }
      if (c8 == 0 && valid) {
        gxy[2 * r] = (1.f - ay) * (td[1] - td[0]) + ay * (td[3] - td[2]);
        gxy[2 * r + 1] = (1.f - ax) * (td[2] - td[0]) + ax * (td[3] - td[1]);
      }
    }
  }
  __syncthreads();
  const float gx = (qvalid && !helper) ? gxy[2 * row] : 0.f, gy = (qvalid && !helper) ? gxy[2 * row + 1] : 0.f;
  // ---- chain to (ox, oy, sx, sy, theta) and reduce over the 49 samples of each problem (2 warps per problem)
  float v5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (qvalid && !helper) {
    const float* prm = params + ((size_t)bw * g.nH + head) * 8;
    const float inv_w = 2.0f / (float)(g.Wq - 1), inv_h = 2.0f / (float)(g.Hq - 1);
    const float bx = (float)(qx - WS / 2) * inv_w, by = (float)(qy - WS / 2) * inv_h;
    const float X = (1.f + prm[2]) * bx, Y = (1.f + prm[3]) * by;
    float s, c;
    sincosf(prm[4], &s, &c);
    const float dcx = gx * 0.5f * (float)(g.Wq - 1), dcy = gy * 0.5f * (float)(g.Hq - 1);
    v5[0] = dcx;
    v5[1] = dcy;
    v5[2] = (dcx * c + dcy * s) * bx;
    v5[3] = (-dcx * s + dcy * c) * by;
    v5[4] = dcx * (-X * s - Y * c) + dcy * (-Y * s + X * c);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float r = warp_sum(v5[k]);
    if (lane == 0 && !helper) red[warp * 8 + k] = r;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 16) {          // slots 5..7 of a parameter row are unused: written as zeros
    const int pp = tid >> 3, k = tid & 7;
    dparams[((size_t)bw * g.nH + 2 * hp + pp) * 8 + k] = k < 5 ? red[(2 * pp) * 8 + k] + red[(2 * pp + 1) * 8 + k] : 0.f;
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

int launch_rvsa_attn_bwd_tc(const void* qkv, const float* params, const float* rel_h, const float* rel_w, const float* table,
                            const float* lse, const void* dout, void* dqkv, float* dkv, float* dparams, float* part_rel,
                            float* part_table, const RvsaGeom& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(rvsa_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rvsa_attn_bwd_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "rvsa_attn_bwd_tc smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  (void)launch_k(rvsa_attn_bwd_tc_kernel, g.B * g.nh * g.nw * (g.nH / 2), WB_THREADS, WB_SMEM, st, 
      reinterpret_cast<const __nv_bfloat16*>(qkv), params, rel_h, rel_w, table, lse, reinterpret_cast<const __nv_bfloat16*>(dout),
      reinterpret_cast<__nv_bfloat16*>(dqkv), dkv, dparams, part_rel, part_table, g);
  return check_launch("rvsa_attn_bwd_tc_kernel");
}

}  // namespace mtp
