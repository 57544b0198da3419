// RVSA window attention forward on tcgen05 tensor cores.                    [V]:372-428, SURVEY.md K5
//
// One CTA handles TWO (window, head) problems of the same window stacked on the 128 rows of one UMMA tile (rows 0..48 = head 2p,
// rows 64..112 = head 2p+1; 15 padding rows each).  256 threads: thread `row` owns accumulator row `row`, its helper `row + 128`
// (same TMEM lane quarter) computes the width rel-pos terms, stores the second half of the output row and joins the gather:
//   gather : threads blend the 4 bilinear taps of every sampled K / V row in fp32 and write bf16 rows straight into
//            128B-swizzled shared-memory tiles (the layout the UMMA descriptors read) together with the Q rows
//   S      : tcgen05.mma  S[128x128] = Q K~^T (K = 64) into TMEM; only the two diagonal 64x64 blocks are meaningful
//   softmax: thread r owns row r: tcgen05.ld its diagonal block, add scale / decomposed rel-pos (fp32, unscaled q) /
//            bias table, softmax over the 49 keys, write P (bf16) block-diagonally into a [128 x 128] tile
//   O      : tcgen05.mma  O[128x64] = P V~ (K = 128; the zero off-diagonal blocks keep the two problems apart)
//   store  : tcgen05.ld O row -> bf16 -> token-major output (padding tokens are never written: the crop is free)
#include "common.h"
#include "ptx.cuh"
#include "rvsa_geom.cuh"
#include "tc_tile.cuh"

namespace mtp {

constexpr int WTC_THREADS = 256;
constexpr int WTC_TILE = 128 * 128;                       // bytes of one 128-row tile
// smem: Q | K~ | V~ | P (2 atoms) | rel_h,rel_w fp32 [2][13][64] | table [2][169] | coords [2][98] | rw [128][8] | mbar | tmem slot
constexpr int WTC_SMEM = 5 * WTC_TILE + 2 * 13 * 64 * 4 + 2 * 169 * 4 + 2 * 98 * 4 + 128 * 8 * 4 + 64;

__global__ void __launch_bounds__(WTC_THREADS, 2)
rvsa_attn_fwd_tc_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ params, const float* __restrict__ rel_h,
                        const float* __restrict__ rel_w, const float* __restrict__ bias_table, __nv_bfloat16* __restrict__ out,
                        float* __restrict__ lse, const RvsaGeom g) {
  MTP_PDL_ENTRY();
  extern __shared__ __align__(1024) uint8_t sm[];      // indexed directly so every access stays in the shared state space
  uint8_t* Qs = sm;
  uint8_t* Ks = Qs + WTC_TILE;
  uint8_t* Vs = Ks + WTC_TILE;
  uint8_t* Ps = Vs + WTC_TILE;                            // two atoms of 128 rows x 128 B
  float* relt = reinterpret_cast<float*>(Ps + 2 * WTC_TILE);   // [2][13][64]: rel_pos_h then rel_pos_w
  float* tabs = relt + 2 * 13 * 64;                       // [2 heads][169]
  float* cpx = tabs + 2 * 169;                            // [98]
  float* cpy = cpx + 98;
  float* rwS = cpy + 98;                                  // [128][8]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(rwS + 128 * 8);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half_heads = g.nH >> 1;
  const int hp = blockIdx.x % half_heads;
  const int bw = blockIdx.x / half_heads;
  const int nwin = g.nh * g.nw;
  const int b = bw / nwin, win = bw % nwin;
  const int wy = win / g.nw, wx = win % g.nw;
  const int C = g.C, C3 = 3 * g.C;

  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (tid == 32) {
    mbar_init(mbar, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 3 * 30 * 8; i += WTC_THREADS) {   // padding rows of Q / K~ / V~ must be zero (the gather writes the rest)
    const int tile = i / 240, rr = (i % 240) >> 3, c = i & 7;
    const int row = rr < 15 ? NTOK + rr : 64 + NTOK + (rr - 15);
    *reinterpret_cast<uint4*>(Qs + tile * WTC_TILE + tile_chunk_off(row, c)) = make_uint4(0, 0, 0, 0);
  }
  {
    // Prologue loads (rel-pos tables, the two heads' bias-table columns, the sampling parameters) are issued as ONE batch and stored afterwards:
    // written as `smem[i] = global[i]` loops they stayed in program order -- seven dependent L2 round trips on every CTA's critical path
    // (ncu source view of the backward: the table store waiting on its load was the kernel's top stall).
    static_assert(WTC_THREADS == 256, "batch sizes below assume 256 threads");
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

  // ---- gather: 8 lanes per (problem, token) row, 16 B (8 head dims) per lane -> 16 rows per pass, 9 x 128-bit loads in flight
  for (int i = tid >> 3; i < 2 * NTOK; i += WTC_THREADS / 8) {
    const int c = tid & 7;
    const int p = i / NTOK, j = i % NTOK, r = 64 * p + j;
    const __nv_bfloat16* qkv_b = qkv + (size_t)b * g.h * g.w * C3 + (2 * hp + p) * HD + c * 8;
    const uint32_t soff = tile_chunk_off(r, c);
    const int y = wy * WS + j / WS - g.pt, x = wx * WS + j % WS - g.pl;
    if (y >= 0 && y < g.h && x >= 0 && x < g.w)
      *reinterpret_cast<uint4*>(Qs + soff) = *reinterpret_cast<const uint4*>(qkv_b + (size_t)(y * g.w + x) * C3);
    else
      *reinterpret_cast<uint4*>(Qs + soff) = make_uint4(0, 0, 0, 0);
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
  fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
  __syncthreads();

  // ---- S = Q K~^T  (M = 128, N = 128, K = 64) -> TMEM columns [0, 128)
  if (warp == 0) {      // warp-uniform issue: descriptors stay in uniform registers (no ELECT/R2UR waterfall per MMA)
    tc_fence_after();
    if (elect_one()) {
      tc_mma_tiles<false, false>(tmem, smem_u32(Qs), 0, smem_u32(Ks), 0, 128, 128, 64, false);
      umma_commit(mbar);
    }
    __syncwarp();
  }
  // meanwhile: decomposed rel-pos terms of this thread's row (fp32, UNscaled q)
  const int row = tid & 127;
  const bool helper = tid >= 128;
  const int p = row >> 6, q = row & 63;
  const bool qvalid = q < NTOK;
  const int qy = q / WS, qx = q % WS;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  float rh[WS], rw[WS];
#pragma unroll
  for (int k = 0; k < WS; ++k) { rh[k] = 0.f; rw[k] = 0.f; }
  if (qvalid) {          // the owner computes the height terms, the helper the width terms
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

  // ---- softmax of this thread's row over its problem's 49 keys
  if (!helper) {
    uint32_t r0[32], r1[32];
    const uint32_t taddr = tmem + lane_base + 64 * p;
    tmem_ld_32x32(taddr, r0);
    tmem_ld_32x32(taddr + 32, r1);
    tmem_ld_wait();
    float s[NTOK];
    float m = -INFINITY;
    const float* tab = tabs + p * 169;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      const float acc = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]);
      const int jy = j / WS, jx = j % WS;
      s[j] = 0.125f * acc + rh[jy] + rw[jx] + (qvalid ? tab[(qy - jy + WS - 1) * (2 * WS - 1) + (qx - jx + WS - 1)] : 0.f);
      m = fmaxf(m, s[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) { s[j] = __expf(s[j] - m); sum += s[j]; }
    const float inv = qvalid ? 1.0f / sum : 0.f;          // padding rows: P = 0
    if (qvalid && lse) lse[((size_t)bw * g.nH + 2 * hp + p) * NTOK + q] = m + __logf(sum);
    uint8_t* prow_mine = Ps + p * WTC_TILE;
    uint8_t* prow_other = Ps + (1 - p) * WTC_TILE;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (c * 8 + e < NTOK) ? s[(c * 8 + e < NTOK) ? c * 8 + e : 0] * inv : 0.f;
      uint4 u;
      u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(prow_mine + tile_chunk_off(row, c)) = u;
      *reinterpret_cast<uint4*>(prow_other + tile_chunk_off(row, c)) = make_uint4(0, 0, 0, 0);
    }
  }
  tc_fence_before();
  fence_proxy_async_smem();
  __syncthreads();

  // ---- O = P V~  (M = 128, N = 64, K = 128; V~ read MN-major: rows = key index) -> TMEM columns [128, 192)
  if (warp == 0) {      // warp-uniform issue: descriptors stay in uniform registers (no ELECT/R2UR waterfall per MMA)
    tc_fence_after();
    if (elect_one()) {
      tc_mma_tiles<false, true>(tmem + 128, smem_u32(Ps), WTC_TILE, smem_u32(Vs), 0, 128, 64, 128, false);
      umma_commit(mbar);
    }
    __syncwarp();
  }
  mbar_wait(mbar, 1);
  tc_fence_after();
  {
    const int hb = helper ? 1 : 0;          // owner stores head dims 0..31 of its row, the helper 32..63
    uint32_t r0[32];
    tmem_ld_32x32(tmem + lane_base + 128 + 32 * hb, r0);
    tmem_ld_wait();
    const int y = wy * WS + qy - g.pt, x = wx * WS + qx - g.pl;
    if (qvalid && y >= 0 && y < g.h && x >= 0 && x < g.w) {
      __nv_bfloat16* dst = out + ((size_t)(b * g.h + y) * g.w + x) * C + (2 * hp + p) * HD + 32 * hb;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(r0[8 * c]), __uint_as_float(r0[8 * c + 1]));
        u.y = pack_bf16x2(__uint_as_float(r0[8 * c + 2]), __uint_as_float(r0[8 * c + 3]));
        u.z = pack_bf16x2(__uint_as_float(r0[8 * c + 4]), __uint_as_float(r0[8 * c + 5]));
        u.w = pack_bf16x2(__uint_as_float(r0[8 * c + 6]), __uint_as_float(r0[8 * c + 7]));
        *reinterpret_cast<uint4*>(dst + 8 * c) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

int launch_rvsa_attn_fwd_tc(const void* qkv, const float* params, const float* rel_h, const float* rel_w, const float* table, void* out,
                            float* lse, const RvsaGeom& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(rvsa_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WTC_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rvsa_attn_fwd_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "rvsa_attn_fwd_tc smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  (void)launch_k(rvsa_attn_fwd_tc_kernel, g.B * g.nh * g.nw * (g.nH / 2), WTC_THREADS, WTC_SMEM, st, 
      reinterpret_cast<const __nv_bfloat16*>(qkv), params, rel_h, rel_w, table, reinterpret_cast<__nv_bfloat16*>(out), lse, g);
  return check_launch("rvsa_attn_fwd_tc_kernel");
}

}  // namespace mtp
