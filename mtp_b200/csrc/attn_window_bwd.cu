// Rotated varied-size window attention (RVSA), backward.   Autograd of [V]:287-433 restated per (image, window, head).
//
//   rvsa_attn_bwd      : recompute coords / k~ / v~ / P from (qkv, params, lse); dV~ = P^T dO, dP = dO V~^T,
//                        dS = P o (dP - rowsum(P o dP)), dQ = scale dS K~ + rel-pos terms, dK~ = scale dS^T Q;
//                        dq is stored straight into dqkv (each query token belongs to exactly one window);
//                        dk/dv are scattered with the bilinear tap weights into an fp32 scratch (red.global.add);
//                        d(coords) -> d(ox, oy, sx, sy, theta); per-CTA partials for rel_pos_h/w and the bias table.
//   rvsa_partials_reduce: deterministic reduction of the per-CTA partials.
//   rvsa_kv_finalize   : fp32 dk/dv scratch -> bf16 k|v slots of dqkv.
//   rvsa_sampling_bwd_*: backward of AvgPool -> LeakyReLU -> 1x1 convs (weights, biases, and the pooled-path gradient
//                        that is added to the LN1-output cotangent).
#include <algorithm>

#include "common.h"
#include "ptx.cuh"
#include "rvsa_geom.cuh"

namespace mtp {

constexpr int BW_LD = 68;     // fp32 row stride of the 49 x 64 tiles
constexpr int BW_LDP = 52;    // row stride of the 49 x 49 tiles
constexpr int BW_THREADS = 64;
constexpr int RVSA_BWD_SMEM_FLOATS = 4 * NTOK * BW_LD + 2 * NTOK * BW_LDP + 4 * NTOK * 8 + 4 * NTOK + 64;

// C[i][j] (7x7 micro-tile per thread, i in [7ti,7ti+7), j in [7tj,7tj+7)) = sum_d A[i][d] * B[j][d]   (both tiles [49][BW_LD])
__device__ __forceinline__ void mm_abt_7x7(const float* __restrict__ A, const float* __restrict__ B, int ti, int tj, float (&acc)[WS][WS]) {
#pragma unroll
  for (int a = 0; a < WS; ++a)
#pragma unroll
    for (int c = 0; c < WS; ++c) acc[a][c] = 0.f;
  for (int d = 0; d < HD; d += 4) {
    float4 av[WS], bv[WS];
#pragma unroll
    for (int a = 0; a < WS; ++a) av[a] = *reinterpret_cast<const float4*>(A + (ti * WS + a) * BW_LD + d);
#pragma unroll
    for (int c = 0; c < WS; ++c) bv[c] = *reinterpret_cast<const float4*>(B + (tj * WS + c) * BW_LD + d);
#pragma unroll
    for (int a = 0; a < WS; ++a)
#pragma unroll
      for (int c = 0; c < WS; ++c)
        acc[a][c] += av[a].x * bv[c].x + av[a].y * bv[c].y + av[a].z * bv[c].z + av[a].w * bv[c].w;
  }
}

// out[i][d] (i in [7ti,7ti+7), d in {4td..4td+3, 32+4td..}) = sum_j M[i][j] * B[j][d]      (TRANS: M[j][i] instead)
template <bool TRANS>
__device__ __forceinline__ void mm_ab_7x8(const float* __restrict__ M, const float* __restrict__ B, int ti, int td, float (&acc)[WS][8]) {
#pragma unroll
  for (int a = 0; a < WS; ++a)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
  for (int j = 0; j < NTOK; ++j) {
    const float4 b0 = *reinterpret_cast<const float4*>(B + j * BW_LD + td * 4);
    const float4 b1 = *reinterpret_cast<const float4*>(B + j * BW_LD + 32 + td * 4);
#pragma unroll
    for (int a = 0; a < WS; ++a) {
      const float m = TRANS ? M[j * BW_LDP + ti * WS + a] : M[(ti * WS + a) * BW_LDP + j];
      acc[a][0] += m * b0.x; acc[a][1] += m * b0.y; acc[a][2] += m * b0.z; acc[a][3] += m * b0.w;
      acc[a][4] += m * b1.x; acc[a][5] += m * b1.y; acc[a][6] += m * b1.z; acc[a][7] += m * b1.w;
    }
  }
}

__device__ __forceinline__ void red_add_f32x2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

__global__ void __launch_bounds__(BW_THREADS)
rvsa_attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ params, const float* __restrict__ rel_h,
                     const float* __restrict__ rel_w, const float* __restrict__ bias_table, const float* __restrict__ lse,
                     const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dkv,
                     float* __restrict__ dparams, float* __restrict__ part_rel, float* __restrict__ part_table, const RvsaGeom g) {
  MTP_PDL_ENTRY();
  extern __shared__ float sm[];
  float* Qs = sm;                          // [49][68]
  float* Ks = Qs + NTOK * BW_LD;           // k~, later dk~
  float* Vs = Ks + NTOK * BW_LD;           // v~, later dv~
  float* Gs = Vs + NTOK * BW_LD;           // dO
  float* Ps = Gs + NTOK * BW_LD;           // [49][52] P
  float* Ds = Ps + NTOK * BW_LDP;          // [49][52] dP, then dS
  float* relh = Ds + NTOK * BW_LDP;        // [49][8]
  float* relw = relh + NTOK * 8;
  float* dSh = relw + NTOK * 8;            // [49][8] row sums of dS over key columns (per key row)
  float* dSw = dSh + NTOK * 8;
  float* cpx = dSw + NTOK * 8;             // [49]
  float* cpy = cpx + NTOK;
  float* gpx = cpy + NTOK;                 // [49] d loss / d px
  float* gpy = gpx + NTOK;
  float* red = gpy + NTOK;                 // [64] scratch

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x % g.nH;
  const int bw = blockIdx.x / g.nH;
  const int nwin = g.nh * g.nw;
  const int b = bw / nwin, win = bw % nwin;
  const int wy = win / g.nw, wx = win % g.nw;
  const int C = g.C, C3 = 3 * g.C;
  const float scale = 0.125f;
  const float* prm = params + ((size_t)bw * g.nH + n) * 8;

  if (tid < NTOK) {
    float px, py;
    rvsa_sample_coord(g, wy, wx, tid / WS, tid % WS, prm[0], prm[1], prm[2], prm[3], prm[4], px, py);
    cpx[tid] = px;
    cpy[tid] = py;
  }
  __syncthreads();

  // ---- stage q, dO (zero rows for padding positions) and the gathered k~, v~
  const __nv_bfloat16* qkv_b = qkv + (size_t)b * g.h * g.w * C3 + n * HD;
  for (int j = warp; j < NTOK; j += 2) {
    const int y = wy * WS + j / WS - g.pt, x = wx * WS + j % WS - g.pl;
    float2 qv = make_float2(0.f, 0.f), gv = make_float2(0.f, 0.f);
    if (y >= 0 && y < g.h && x >= 0 && x < g.w) {
      const size_t t = (size_t)(b * g.h + y) * g.w + x;
      qv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(qkv + t * C3 + n * HD + lane * 2));
      gv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + t * C + n * HD + lane * 2));
    }
    *reinterpret_cast<float2*>(Qs + j * BW_LD + lane * 2) = qv;
    *reinterpret_cast<float2*>(Gs + j * BW_LD + lane * 2) = gv;
    const float px = cpx[j], py = cpy[j];
    const float fx0 = floorf(px), fy0 = floorf(py);
    const float ax = px - fx0, ay = py - fy0;
    const int x0 = (int)fx0 - g.pl, y0 = (int)fy0 - g.pt;
    float2 ka = make_float2(0.f, 0.f), va = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
      const float wgt = ((t & 1) ? ax : 1.f - ax) * ((t >> 1) ? ay : 1.f - ay);
      if (xx >= 0 && xx < g.w && yy >= 0 && yy < g.h) {
        const __nv_bfloat16* src = qkv_b + (size_t)(yy * g.w + xx) * C3 + lane * 2;
        const float2 kv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + C));
        const float2 vv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + 2 * C));
        ka.x += wgt * kv.x; ka.y += wgt * kv.y;
        va.x += wgt * vv.x; va.y += wgt * vv.y;
      }
    }
    *reinterpret_cast<float2*>(Ks + j * BW_LD + lane * 2) = ka;
    *reinterpret_cast<float2*>(Vs + j * BW_LD + lane * 2) = va;
  }
  __syncthreads();

  // ---- rel-pos tables of the forward
  for (int e = tid; e < NTOK * 14; e += BW_THREADS) {
    const int q = e / 14, r = e % 14, kk = r % 7;
    const float* tab = (r < 7 ? rel_h : rel_w) + ((r < 7 ? q / WS : q % WS) - kk + WS - 1) * HD;
    const float* qr = Qs + q * BW_LD;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 tv = __ldg(reinterpret_cast<const float4*>(tab + d));
      const float4 qv = *reinterpret_cast<const float4*>(qr + d);
      s += qv.x * tv.x + qv.y * tv.y + qv.z * tv.z + qv.w * tv.w;
    }
    (r < 7 ? relh : relw)[q * 8 + kk] = s;
  }
  __syncthreads();

  // ---- P = exp(S - lse), dP = dO v~^T   (thread (tq, tj): 7 x 7 micro-tiles)
  if (tid < NTOK) {
    const int tq = tid / WS, tj = tid % WS;
    float acc[WS][WS];
    mm_abt_7x7(Qs, Ks, tq, tj, acc);
    const float* lrow = lse + (size_t)blockIdx.x * NTOK;
#pragma unroll
    for (int a = 0; a < WS; ++a) {
      const int q = tq * WS + a;
      const float l = lrow[q];
#pragma unroll
      for (int c = 0; c < WS; ++c) {
        const int idx = (tq - tj + WS - 1) * (2 * WS - 1) + (a - c + WS - 1);
        const float s = scale * acc[a][c] + relh[q * 8 + tj] + relw[q * 8 + c] + __ldg(bias_table + idx * g.nH + n);
        Ps[q * BW_LDP + tj * WS + c] = __expf(s - l);
      }
    }
    mm_abt_7x7(Gs, Vs, tq, tj, acc);
#pragma unroll
    for (int a = 0; a < WS; ++a)
#pragma unroll
      for (int c = 0; c < WS; ++c) Ds[(tq * WS + a) * BW_LDP + tj * WS + c] = acc[a][c];
  }
  __syncthreads();

  // ---- dS = P o (dP - D), D = rowsum(P o dP); row sums of dS per key row / key column for the rel-pos terms
  if (tid < NTOK) {
    const float* p = Ps + tid * BW_LDP;
    float* dp = Ds + tid * BW_LDP;
    float D = 0.f;
    for (int j = 0; j < NTOK; ++j) D += p[j] * dp[j];
    float sh[WS], sw[WS];
#pragma unroll
    for (int k = 0; k < WS; ++k) { sh[k] = 0.f; sw[k] = 0.f; }
#pragma unroll
    for (int kh = 0; kh < WS; ++kh)
#pragma unroll
      for (int kw = 0; kw < WS; ++kw) {
        const float ds = p[kh * WS + kw] * (dp[kh * WS + kw] - D);
        dp[kh * WS + kw] = ds;
        sh[kh] += ds;
        sw[kw] += ds;
      }
#pragma unroll
    for (int k = 0; k < WS; ++k) { dSh[tid * 8 + k] = sh[k]; dSw[tid * 8 + k] = sw[k]; }
  }
  __syncthreads();

  // ---- dQ = scale dS k~ + sum_kh dSh Rh + sum_kw dSw Rw   -> dqkv (q slot), valid tokens only
  if (tid < 56) {
    const int tq = tid >> 3, td = tid & 7;
    float acc[WS][8];
    mm_ab_7x8<false>(Ds, Ks, tq, td, acc);
#pragma unroll
    for (int a = 0; a < WS; ++a) {
      const int q = tq * WS + a;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = scale * acc[a][e];
#pragma unroll
      for (int k = 0; k < WS; ++k) {
        const float ch = dSh[q * 8 + k], cw = dSw[q * 8 + k];
        const float* th = rel_h + (tq - k + WS - 1) * HD;     // qy = tq
        const float* tw = rel_w + (a - k + WS - 1) * HD;      // qx = a
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(th + td * 4)), h1 = __ldg(reinterpret_cast<const float4*>(th + 32 + td * 4));
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(tw + td * 4)), w1 = __ldg(reinterpret_cast<const float4*>(tw + 32 + td * 4));
        o[0] += ch * h0.x + cw * w0.x; o[1] += ch * h0.y + cw * w0.y; o[2] += ch * h0.z + cw * w0.z; o[3] += ch * h0.w + cw * w0.w;
        o[4] += ch * h1.x + cw * w1.x; o[5] += ch * h1.y + cw * w1.y; o[6] += ch * h1.z + cw * w1.z; o[7] += ch * h1.w + cw * w1.w;
      }
      const int y = wy * WS + tq - g.pt, x = wx * WS + a - g.pl;
      if (y >= 0 && y < g.h && x >= 0 && x < g.w) {
        __nv_bfloat16* dst = dqkv + ((size_t)(b * g.h + y) * g.w + x) * C3 + n * HD;
        uint2 u0, u1;
        u0.x = pack_bf16x2(o[0], o[1]); u0.y = pack_bf16x2(o[2], o[3]);
        u1.x = pack_bf16x2(o[4], o[5]); u1.y = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint2*>(dst + td * 4) = u0;
        *reinterpret_cast<uint2*>(dst + 32 + td * 4) = u1;
      }
    }
  }
  __syncthreads();      // Ks, Vs are overwritten next

  // ---- dk~ = scale dS^T Q -> Ks ;  dv~ = P^T dO -> Vs
  if (tid < 56) {
    const int tj = tid >> 3, td = tid & 7;
    float acc[WS][8];
    mm_ab_7x8<true>(Ds, Qs, tj, td, acc);
#pragma unroll
    for (int a = 0; a < WS; ++a) {
      float* r = Ks + (tj * WS + a) * BW_LD;
      *reinterpret_cast<float4*>(r + td * 4) = make_float4(scale * acc[a][0], scale * acc[a][1], scale * acc[a][2], scale * acc[a][3]);
      *reinterpret_cast<float4*>(r + 32 + td * 4) = make_float4(scale * acc[a][4], scale * acc[a][5], scale * acc[a][6], scale * acc[a][7]);
    }
    mm_ab_7x8<true>(Ps, Gs, tj, td, acc);
#pragma unroll
    for (int a = 0; a < WS; ++a) {
      float* r = Vs + (tj * WS + a) * BW_LD;
      *reinterpret_cast<float4*>(r + td * 4) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      *reinterpret_cast<float4*>(r + 32 + td * 4) = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
    }
  }
  // ---- per-CTA partials of d rel_pos_h / d rel_pos_w:  dR[r][d] = sum over (q, k) with q_axis - k + 6 == r of dSx[q][k] * Q[q][d]
  for (int e = tid; e < 2 * (2 * WS - 1) * HD; e += BW_THREADS) {
    const int d = e % HD, r = (e / HD) % (2 * WS - 1), which = e / (HD * (2 * WS - 1));
    float s = 0.f;
    for (int qa = 0; qa < WS; ++qa) {          // qa = query coordinate along the axis (row for h, column for w)
      const int k = qa - (r - (WS - 1));
      if (k < 0 || k >= WS) continue;
      for (int qb = 0; qb < WS; ++qb) {
        const int q = which == 0 ? qa * WS + qb : qb * WS + qa;
        s += (which == 0 ? dSh : dSw)[q * 8 + k] * Qs[q * BW_LD + d];
      }
    }
    part_rel[(size_t)blockIdx.x * (2 * (2 * WS - 1) * HD) + e] = s;
  }
  // ---- per-CTA partial of d bias table: index (dy+6)*13 + (dx+6) collects dS over all pairs with that displacement
  for (int idx = tid; idx < 169; idx += BW_THREADS) {
    const int dy = idx / 13 - (WS - 1), dx = idx % 13 - (WS - 1);
    float s = 0.f;
    for (int qy = max(0, dy); qy < min(WS, WS + dy); ++qy)
      for (int qx = max(0, dx); qx < min(WS, WS + dx); ++qx)
        s += Ds[(qy * WS + qx) * BW_LDP + (qy - dy) * WS + (qx - dx)];
    part_table[(size_t)blockIdx.x * 169 + idx] = s;
  }
  __syncthreads();

  // ---- scatter dk~, dv~ through the bilinear taps; coordinate gradients
  float* dk_b = dkv + (size_t)b * g.h * g.w * 2 * C + n * HD;      // scratch rows are [k (C) | v (C)]
  for (int j = warp; j < NTOK; j += 2) {
    const float px = cpx[j], py = cpy[j];
    const float fx0 = floorf(px), fy0 = floorf(py);
    const float ax = px - fx0, ay = py - fy0;
    const int x0 = (int)fx0 - g.pl, y0 = (int)fy0 - g.pt;
    const float2 gk = *reinterpret_cast<const float2*>(Ks + j * BW_LD + lane * 2);
    const float2 gv = *reinterpret_cast<const float2*>(Vs + j * BW_LD + lane * 2);
    float tapdot[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
      const float wgt = ((t & 1) ? ax : 1.f - ax) * ((t >> 1) ? ay : 1.f - ay);
      tapdot[t] = 0.f;
      if (xx >= 0 && xx < g.w && yy >= 0 && yy < g.h) {
        const size_t p = (size_t)(yy * g.w + xx);
        const __nv_bfloat16* src = qkv_b + p * C3 + lane * 2;
        const float2 kv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + C));
        const float2 vv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + 2 * C));
        tapdot[t] = gk.x * kv.x + gk.y * kv.y + gv.x * vv.x + gv.y * vv.y;
        float* dst = dk_b + p * 2 * C + lane * 2;
        red_add_f32x2(dst, wgt * gk.x, wgt * gk.y);
        red_add_f32x2(dst + C, wgt * gv.x, wgt * gv.y);
      }
    }
    // d/dpx = (1-ay)(t1 - t0) + ay(t3 - t2) ; d/dpy = (1-ax)(t2 - t0) + ax(t3 - t1)   (out-of-range taps are zeros)
    float gx = (1.f - ay) * (tapdot[1] - tapdot[0]) + ay * (tapdot[3] - tapdot[2]);
    float gy = (1.f - ax) * (tapdot[2] - tapdot[0]) + ax * (tapdot[3] - tapdot[1]);
    gx = warp_sum(gx);
    gy = warp_sum(gy);
    if (lane == 0) { gpx[j] = gx; gpy[j] = gy; }
  }
  __syncthreads();

  // ---- chain to (ox, oy, sx, sy, theta):  cx = refx + X c - Y s + ox, cy = refy + Y c + X s + oy, X = (1+sx) bx, Y = (1+sy) by
  float v5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (tid < NTOK) {
    const int iy = tid / WS, ix = tid % WS;
    const float inv_w = 2.0f / (float)(g.Wq - 1), inv_h = 2.0f / (float)(g.Hq - 1);
    const float bx = (float)(ix - WS / 2) * inv_w, by = (float)(iy - WS / 2) * inv_h;
    const float X = (1.f + prm[2]) * bx, Y = (1.f + prm[3]) * by;
    float s, c;
    sincosf(prm[4], &s, &c);
    const float dcx = gpx[tid] * 0.5f * (float)(g.Wq - 1), dcy = gpy[tid] * 0.5f * (float)(g.Hq - 1);
    v5[0] = dcx;
    v5[1] = dcy;
    v5[2] = (dcx * c + dcy * s) * bx;
    v5[3] = (-dcx * s + dcy * c) * by;
    v5[4] = dcx * (-X * s - Y * c) + dcy * (-Y * s + X * c);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float r = warp_sum(v5[k]);
    if (lane == 0) red[warp * 8 + k] = r;
  }
  __syncthreads();
  if (tid < 8) dparams[((size_t)bw * g.nH + n) * 8 + tid] = tid < 5 ? red[tid] + red[8 + tid] : 0.f;      // slots 5..7 unused
}

// ------------------------------------------------------------------------------------------------ reductions
// d_rel[2][13][64] += sum over CTAs ; d_table[169][nH] += sum over (image, window).
// A warp owns 32 consecutive outputs (coalesced rows of the partial arrays) and one slice of the partials, all of its loads in
// flight at once; the 8 warps of a CTA hold the 8 slices of the same outputs and combine through shared memory.
constexpr int PR_SLICES = 8;
__device__ __forceinline__ void rvsa_partials_reduce_body(const float* __restrict__ part_rel, const float* __restrict__ part_table,
                                                          float* __restrict__ d_rel_h, float* __restrict__ d_rel_w, float* __restrict__ d_table,
                                                          int n_rel_parts, int n_bw, int nH, int bx) {
  __shared__ float red[PR_SLICES][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rel = 2 * (2 * WS - 1) * HD;
  const int rel_blocks = n_rel / 32;                       // 52
  const int tab_blocks = (169 + 31) / 32;                  // per head
  float s = 0.f;
  float* dst = nullptr;
  if (bx < rel_blocks) {
    const int i = bx * 32 + lane;
    const int per = (n_rel_parts + PR_SLICES - 1) / PR_SLICES;
    const int c0 = warp * per, c1 = min(n_rel_parts, c0 + per);
#pragma unroll 8
    for (int c = c0; c < c1; ++c) s += part_rel[(size_t)c * n_rel + i];
    dst = i < n_rel / 2 ? d_rel_h + i : d_rel_w + (i - n_rel / 2);
  } else {
    const int e = bx - rel_blocks;                 // (head, 32-wide displacement block)
    const int n = e / tab_blocks, idx = (e % tab_blocks) * 32 + lane;
    if (n < nH && idx < 169) {
      const int per = (n_bw + PR_SLICES - 1) / PR_SLICES;
      const int c0 = warp * per, c1 = min(n_bw, c0 + per);
#pragma unroll 8
      for (int bw = c0; bw < c1; ++bw) s += part_table[((size_t)bw * nH + n) * 169 + idx];
      dst = d_table + idx * nH + n;
    }
  }
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && dst != nullptr) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < PR_SLICES; ++w) t += red[w][lane];
    *dst += t;
  }
}

__global__ void __launch_bounds__(256)
rvsa_partials_reduce_kernel(const float* __restrict__ part_rel, const float* __restrict__ part_table, float* __restrict__ d_rel_h,
                            float* __restrict__ d_rel_w, float* __restrict__ d_table, int n_rel_parts, int n_bw, int nH) {
  MTP_PDL_ENTRY();
  rvsa_partials_reduce_body(part_rel, part_table, d_rel_h, d_rel_w, d_table, n_rel_parts, n_bw, nH, (int)blockIdx.x);
}

// dqkv[t, C + c] = bf16(dkv[t, c]) for c in [0, 2C); optionally colsum[0, 3C) += column sums of the finished bf16 dqkv (the qkv
// bias gradient), reading the dq part the attention kernel wrote.  Thread = 4 consecutive columns, CTA = a band of rows.
__device__ __forceinline__ void rvsa_kv_finalize_body(float* __restrict__ dkv, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ colsum,
                                                      int T, int C, int rows_per_cta, int rezero, int bx, int by) {
  const int c = (bx * 256 + threadIdx.x) * 4;               // column of dqkv
  if (c >= 3 * C) return;
  if (c < C && colsum == nullptr) return;
  const int row0 = by * rows_per_cta, row1 = min(T, row0 + rows_per_cta);
  float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
  for (int t = row0; t < row1; ++t) {
    uint2 u;
    if (c < C) {
      u = *reinterpret_cast<const uint2*>(dqkv + (size_t)t * 3 * C + c);
    } else {
      float4* src = reinterpret_cast<float4*>(dkv + (size_t)t * 2 * C + (c - C));
      const float4 v = *src;
      if (rezero) *src = make_float4(0.f, 0.f, 0.f, 0.f);      // hand the scatter scratch back all-zero (no memset next time)
      u.x = pack_bf16x2(v.x, v.y);
      u.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(dqkv + (size_t)t * 3 * C + c) = u;
    }
    const float2 a = unpack_bf16x2(u.x), d = unpack_bf16x2(u.y);
    acc.x += a.x; acc.y += a.y; acc.z += d.x; acc.w += d.y;
  }
  if (colsum != nullptr)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum + c), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
}
__global__ void __launch_bounds__(256)
rvsa_kv_finalize_kernel(float* __restrict__ dkv, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ colsum, int T, int C,
                        int rows_per_cta, int rezero) {
  MTP_PDL_ENTRY();
  rvsa_kv_finalize_body(dkv, dqkv, colsum, T, C, rows_per_cta, rezero, (int)blockIdx.x, (int)blockIdx.y);
}

// ------------------------------------------------------------------------------------------------ sampling heads bwd
// (1) per (image, window): gradient w.r.t. the 5nH raw conv outputs (g_out), and the pooled-path gradient
//     dpooled[c] = leaky'(pooled[c]) * sum_o g_o W[o][c]
// gradient w.r.t. raw conv output o of window bw, from the gradient of the sampling parameters (offsets are divided by the window count)
__device__ __forceinline__ float rvsa_sampling_gout(const float* __restrict__ dparams, int bw, int o, int nH, const RvsaGeom& g) {
  if (o < 2 * nH) return dparams[((size_t)bw * nH + (o >> 1)) * 8 + (o & 1)] / (float)(((o & 1) == 0 ? g.h : g.w) / WS);
  if (o < 4 * nH) return dparams[((size_t)bw * nH + ((o - 2 * nH) >> 1)) * 8 + 2 + ((o - 2 * nH) & 1)];
  return dparams[((size_t)bw * nH + (o - 4 * nH)) * 8 + 4];
}

__device__ __forceinline__ void rvsa_sampling_bwd_body(const float* __restrict__ dparams, const float* __restrict__ pooled,
                                                       const float* __restrict__ w_off, const float* __restrict__ w_sc,
                                                       const float* __restrict__ w_ang, float* __restrict__ g_out, float* __restrict__ dpooled,
                                                       const RvsaGeom& g, int bx, int by) {
  __shared__ float gs[5 * 64];       // nH <= 64
  const int bw = bx, nH = g.nH, C = g.C;
  for (int o = threadIdx.x; o < 5 * nH; o += 256) {
    const float v = rvsa_sampling_gout(dparams, bw, o, nH, g);
    gs[o] = v;
    if (by == 0 && g_out != nullptr) g_out[(size_t)bw * 5 * nH + o] = v;
  }
  __syncthreads();
  const int c = by * 256 + threadIdx.x;                  // CTA = (image-window, 256-channel slab)
  if (c < C) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int o = 0; o < 2 * nH; ++o) {
      s0 += gs[o] * __ldg(w_off + (size_t)o * C + c);
      s1 += gs[2 * nH + o] * __ldg(w_sc + (size_t)o * C + c);
    }
#pragma unroll 8
    for (int o = 0; o < nH; ++o) s0 += gs[4 * nH + o] * __ldg(w_ang + (size_t)o * C + c);
    const float p = pooled[(size_t)bw * C + c];
    dpooled[(size_t)bw * C + c] = (p >= 0.f ? 1.0f : 0.01f) * (s0 + s1);
  }
}
__global__ void __launch_bounds__(256)
rvsa_sampling_bwd_kernel(const float* __restrict__ dparams, const float* __restrict__ pooled, const float* __restrict__ w_off,
                         const float* __restrict__ w_sc, const float* __restrict__ w_ang, float* __restrict__ g_out,
                         float* __restrict__ dpooled, const RvsaGeom g) {
  MTP_PDL_ENTRY();
  rvsa_sampling_bwd_body(dparams, pooled, w_off, w_sc, w_ang, g_out, dpooled, g, (int)blockIdx.x, (int)blockIdx.y);
}

// (2) weight / bias gradients: dW[o][c] += sum_bw g[bw][o] * leaky(pooled[bw][c]) ; db[o] += sum_bw g[bw][o], with g recomputed from dparams
//     (so that this does not depend on kernel (1) and both can share a launch)
__device__ __forceinline__ void rvsa_sampling_wgrad_body(const float* __restrict__ dparams, const float* __restrict__ pooled,
                                                         float* __restrict__ dw_off, float* __restrict__ db_off, float* __restrict__ dw_sc,
                                                         float* __restrict__ db_sc, float* __restrict__ dw_ang, float* __restrict__ db_ang,
                                                         int n_bw, const RvsaGeom& g, int bx, int by) {
  __shared__ float gw[256];
  const int nH = g.nH, C = g.C;
  const int o = by;                              // 0 .. 5nH-1
  const int c = bx * 256 + threadIdx.x;
  float* dw; float* db; int oo;
  if (o < 2 * nH) { dw = dw_off; db = db_off; oo = o; }
  else if (o < 4 * nH) { dw = dw_sc; db = db_sc; oo = o - 2 * nH; }
  else { dw = dw_ang; db = db_ang; oo = o - 4 * nH; }
  float s = 0.f, sb = 0.f;
  for (int bw0 = 0; bw0 < n_bw; bw0 += 256) {
    const int nb = min(256, n_bw - bw0);
    __syncthreads();
    if ((int)threadIdx.x < nb) gw[threadIdx.x] = rvsa_sampling_gout(dparams, bw0 + threadIdx.x, o, nH, g);
    __syncthreads();
    if (c < C) {
#pragma unroll 8
      for (int i = 0; i < nb; ++i) {
        const float p = pooled[(size_t)(bw0 + i) * C + c];
        s += gw[i] * (p >= 0.f ? p : 0.01f * p);
      }
    }
    if (bx == 0 && threadIdx.x == 0)
      for (int i = 0; i < nb; ++i) sb += gw[i];
  }
  if (c < C) dw[(size_t)oo * C + c] += s;
  if (bx == 0 && threadIdx.x == 0) db[oo] += sb;
}
__global__ void __launch_bounds__(256)
rvsa_sampling_wgrad_kernel(const float* __restrict__ dparams, const float* __restrict__ pooled, float* __restrict__ dw_off,
                           float* __restrict__ db_off, float* __restrict__ dw_sc, float* __restrict__ db_sc, float* __restrict__ dw_ang,
                           float* __restrict__ db_ang, int n_bw, const RvsaGeom g) {
  MTP_PDL_ENTRY();
  rvsa_sampling_wgrad_body(dparams, pooled, dw_off, db_off, dw_sc, db_sc, dw_ang, db_ang, n_bw, g, (int)blockIdx.x, (int)blockIdx.y);
}

// The four kernels that follow the attention backward -- finishing dK / dV (+ the qkv bias gradient), reducing the rel-pos / bias-table
// partials, the sampling heads' input gradient and their weight gradients -- only depend on that kernel's outputs, not on each other, and
// each is a short latency-bound launch (128 - 588 CTAs, 6 - 12 us): ONE launch, role by block range, lets them share the SMs and saves three
// kernel boundaries per window block.
struct RvsaTailArgs {
  // kv finalize
  float* dkv; __nv_bfloat16* dqkv; float* colsum; int T, C, rows_per_cta, rezero, kv_gx, kv_gy;
  // partials reduce
  const float* part_rel; const float* part_table; float* d_rel_h; float* d_rel_w; float* d_table; int n_rel_parts, n_bw, nH, n_pr;
  // sampling heads
  const float* dparams; const float* pooled; const float* w_off; const float* w_sc; const float* w_ang; float* dpooled;
  float* dw_off; float* db_off; float* dw_sc; float* db_sc; float* dw_ang; float* db_ang; int sb_gy, wg_gx;
  RvsaGeom g;
};
__global__ void __launch_bounds__(256)
rvsa_bwd_tail_kernel(const __grid_constant__ RvsaTailArgs a) {
  MTP_PDL_ENTRY();
  int b = (int)blockIdx.x;
  const int n_kv = a.kv_gx * a.kv_gy;
  if (b < n_kv) { rvsa_kv_finalize_body(a.dkv, a.dqkv, a.colsum, a.T, a.C, a.rows_per_cta, a.rezero, b % a.kv_gx, b / a.kv_gx); return; }
  b -= n_kv;
  const int n_sb = a.n_bw * a.sb_gy;
  if (b < n_sb) { rvsa_sampling_bwd_body(a.dparams, a.pooled, a.w_off, a.w_sc, a.w_ang, nullptr, a.dpooled, a.g, b % a.n_bw, b / a.n_bw); return; }
  b -= n_sb;
  const int n_wg = a.wg_gx * 5 * a.nH;
  if (b < n_wg) { rvsa_sampling_wgrad_body(a.dparams, a.pooled, a.dw_off, a.db_off, a.dw_sc, a.db_sc, a.dw_ang, a.db_ang, a.n_bw, a.g, b % a.wg_gx, b / a.wg_gx); return; }
  b -= n_wg;
  rvsa_partials_reduce_body(a.part_rel, a.part_table, a.d_rel_h, a.d_rel_w, a.d_table, a.n_rel_parts, a.n_bw, a.nH, b);
}

// (1)+(2) in ONE launch: CTA = slab of 32 channels.  g (the gradient w.r.t. the 5nH raw conv outputs of every window), the slab of
// LeakyReLU(pooled) / its derivative and the slab of the conv weights sit in shared memory; thread (c, o-group) accumulates
// dW[o][c] over the windows, thread (c, window-group) dpooled[window][c] over the outputs.  Replaces two launches (128 + 320 CTAs).
constexpr int SB_SLAB = 32;
__global__ void __launch_bounds__(256)
rvsa_sampling_fused_bwd_kernel(const float* __restrict__ dparams, const float* __restrict__ pooled, const float* __restrict__ w_off,
                               const float* __restrict__ w_sc, const float* __restrict__ w_ang, float* __restrict__ dw_off,
                               float* __restrict__ db_off, float* __restrict__ dw_sc, float* __restrict__ db_sc, float* __restrict__ dw_ang,
                               float* __restrict__ db_ang, float* __restrict__ dpooled, int n_bw, const RvsaGeom g) {
  MTP_PDL_ENTRY();
  extern __shared__ float sbm[];
  const int nH = g.nH, C = g.C, NO = 5 * nH;
  float* g_s = sbm;                          // [n_bw][NO]
  float* a_s = g_s + n_bw * NO;              // [n_bw][32]  LeakyReLU(pooled)
  float* k_s = a_s + n_bw * SB_SLAB;         // [n_bw][32]  LeakyReLU'(pooled)
  float* w_s = k_s + n_bw * SB_SLAB;         // [NO][32]
  const int tid = threadIdx.x, c0 = blockIdx.x * SB_SLAB;
  for (int i = tid; i < n_bw * NO; i += 256) {
    const int bw = i / NO, o = i % NO;
    float v;
    if (o < 2 * nH) v = dparams[((size_t)bw * nH + (o >> 1)) * 8 + (o & 1)] / (float)(((o & 1) == 0 ? g.h : g.w) / WS);
    else if (o < 4 * nH) v = dparams[((size_t)bw * nH + ((o - 2 * nH) >> 1)) * 8 + 2 + ((o - 2 * nH) & 1)];
    else v = dparams[((size_t)bw * nH + (o - 4 * nH)) * 8 + 4];
    g_s[i] = v;
  }
  for (int i = tid; i < n_bw * SB_SLAB; i += 256) {
    const int bw = i / SB_SLAB, c = i % SB_SLAB;
    const float p = (c0 + c < C) ? pooled[(size_t)bw * C + c0 + c] : 0.f;
    a_s[i] = p >= 0.f ? p : 0.01f * p;
    k_s[i] = p >= 0.f ? 1.0f : 0.01f;
  }
  for (int i = tid; i < NO * SB_SLAB; i += 256) {
    const int o = i / SB_SLAB, c = i % SB_SLAB;
    const float* wr = o < 2 * nH ? w_off + (size_t)o * C : o < 4 * nH ? w_sc + (size_t)(o - 2 * nH) * C : w_ang + (size_t)(o - 4 * nH) * C;
    w_s[i] = (c0 + c < C) ? wr[c0 + c] : 0.f;
  }
  __syncthreads();
  const int c = tid & 31, grp = tid >> 5;
  if (c0 + c < C) {
    for (int o = grp; o < NO; o += 8) {            // dW[o][c] += sum_bw g[bw][o] * a[bw][c]
      float acc = 0.f;
      for (int bw = 0; bw < n_bw; ++bw) acc += g_s[bw * NO + o] * a_s[bw * SB_SLAB + c];
      float* dw = o < 2 * nH ? dw_off + (size_t)o * C : o < 4 * nH ? dw_sc + (size_t)(o - 2 * nH) * C : dw_ang + (size_t)(o - 4 * nH) * C;
      dw[c0 + c] += acc;
    }
    for (int bw = grp; bw < n_bw; bw += 8) {       // dpooled[bw][c] = leaky'(pooled) * sum_o g[bw][o] W[o][c]
      float acc = 0.f;
      for (int o = 0; o < NO; ++o) acc += g_s[bw * NO + o] * w_s[o * SB_SLAB + c];
      dpooled[(size_t)bw * C + c0 + c] = k_s[bw * SB_SLAB + c] * acc;
    }
  }
  if (blockIdx.x == 0) {                          // bias gradients
    for (int o = tid; o < NO; o += 256) {
      float acc = 0.f;
      for (int bw = 0; bw < n_bw; ++bw) acc += g_s[bw * NO + o];
      float* db = o < 2 * nH ? db_off + o : o < 4 * nH ? db_sc + (o - 2 * nH) : db_ang + (o - 4 * nH);
      *db += acc;
    }
  }
}

// (3) dyn[t][c] += dpooled[window(t)][c] / 49   (AvgPool backward; only real tokens receive it)
__global__ void __launch_bounds__(256)
rvsa_pool_bwd_add_kernel(const float* __restrict__ dpooled, __nv_bfloat16* __restrict__ dyn, const RvsaGeom g) {
  MTP_PDL_ENTRY();
  const int c4n = g.C / 4;
  const size_t total = (size_t)g.B * g.h * g.w * c4n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const size_t t = i / c4n;
    const int x = (int)(t % g.w), y = (int)((t / g.w) % g.h), b = (int)(t / ((size_t)g.w * g.h));
    const int bw = (b * g.nh + (y + g.pt) / WS) * g.nw + (x + g.pl) / WS;
    const float4 dp = *reinterpret_cast<const float4*>(dpooled + (size_t)bw * g.C + c);
    uint2 u = *reinterpret_cast<uint2*>(dyn + t * g.C + c);
    float2 a = unpack_bf16x2(u.x), d = unpack_bf16x2(u.y);
    const float inv = 1.0f / (WS * WS);
    a.x += dp.x * inv; a.y += dp.y * inv; d.x += dp.z * inv; d.y += dp.w * inv;
    u.x = pack_bf16x2(a.x, a.y);
    u.y = pack_bf16x2(d.x, d.y);
    *reinterpret_cast<uint2*>(dyn + t * g.C + c) = u;
  }
}

int launch_rvsa_attn_bwd_tc(const void* qkv, const float* params, const float* rel_h, const float* rel_w, const float* table,
                            const float* lse, const void* dout, void* dqkv, float* dkv, float* dparams, float* part_rel,
                            float* part_table, const RvsaGeom& g, cudaStream_t st);     // attn_window_bwd_tc.cu

}  // namespace mtp

using namespace mtp;

extern "C" size_t mtp_rvsa_bwd_workspace_bytes(int B, int h, int w, int C, int nH) {
  const RvsaGeom g = make_rvsa_geom(B, h, w, C, nH);
  const size_t n_cta = (size_t)B * g.nh * g.nw * nH;
  // [part_rel | part_table | dkv scratch]
  return (n_cta * (2 * (2 * WS - 1) * HD) + n_cta * 169 + (size_t)B * h * w * 2 * C) * sizeof(float);
}

extern "C" int mtp_rvsa_attn_bwd(const void* qkv_bf16, const float* params, const float* rel_pos_h, const float* rel_pos_w,
                                 const float* bias_table, const float* lse, const void* dout_bf16, void* dqkv_bf16, float* dparams,
                                 float* d_rel_pos_h, float* d_rel_pos_w, float* d_bias_table, float* d_qkv_bias, void* workspace,
                                 int scratch_zeroed, int B, int h, int w, int C, int nH, mtp_stream_t stream) {
  MTP_REQUIRE(qkv_bf16 && params && rel_pos_h && rel_pos_w && bias_table && lse && dout_bf16 && dqkv_bf16 && dparams &&
                  d_rel_pos_h && d_rel_pos_w && d_bias_table && workspace, "mtp_rvsa_attn_bwd: null pointer");
  MTP_REQUIRE(B > 0 && h >= WS && w >= WS && C == nH * HD, "mtp_rvsa_attn_bwd: B=%d h=%d w=%d C=%d nH=%d unsupported", B, h, w, C, nH);
  const RvsaGeom g = make_rvsa_geom(B, h, w, C, nH);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n_cta = B * g.nh * g.nw * nH;
  float* part_rel = reinterpret_cast<float*>(workspace);
  float* part_table = part_rel + (size_t)n_cta * (2 * (2 * WS - 1) * HD);
  float* dkv = part_table + (size_t)n_cta * 169;
  const size_t T = (size_t)B * h * w;
  cudaError_t e = cudaSuccess;
  if (!scratch_zeroed) {
    e = cudaMemsetAsync(dkv, 0, T * 2 * C * sizeof(float), st);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "rvsa bwd memset: %s", cudaGetErrorString(e));
  }
  int rc, n_rel_parts;
  if (nH % 2 == 0) {       // tensor-core path (two heads per UMMA tile); rel-pos partials are per CTA = per head pair
    rc = launch_rvsa_attn_bwd_tc(qkv_bf16, params, rel_pos_h, rel_pos_w, bias_table, lse, dout_bf16, dqkv_bf16, dkv, dparams, part_rel,
                                 part_table, g, st);
    n_rel_parts = n_cta / 2;
  } else {
    static bool attr = false;
    const int smem = RVSA_BWD_SMEM_FLOATS * sizeof(float);
    if (!attr) {
      e = cudaFuncSetAttribute(rvsa_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "rvsa_attn_bwd smem attr: %s", cudaGetErrorString(e));
      attr = true;
    }
    (void)launch_k(rvsa_attn_bwd_kernel, n_cta, BW_THREADS, smem, st, 
        reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), params, rel_pos_h, rel_pos_w, bias_table, lse,
        reinterpret_cast<const __nv_bfloat16*>(dout_bf16), reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), dkv, dparams, part_rel, part_table, g);
    rc = check_launch("rvsa_attn_bwd_kernel");
    n_rel_parts = n_cta;
  }
  if (rc) return rc;
  const int n_red_blocks = 2 * (2 * WS - 1) * HD / 32 + nH * ((169 + 31) / 32);
  (void)launch_k(rvsa_partials_reduce_kernel, n_red_blocks, 256, 0, st, part_rel, part_table, d_rel_pos_h, d_rel_pos_w, d_bias_table,
                 n_rel_parts, n_cta / nH, nH);
  rc = check_launch("rvsa_partials_reduce_kernel");
  if (rc) return rc;
  const int gx = ceil_div(3 * C, 1024);
  const int gy = std::max(1, std::min(ceil_div((int)T, 8), 4 * num_sms() / gx));
  const int rpc = ceil_div((int)T, gy);
  (void)launch_k(rvsa_kv_finalize_kernel, dim3(gx, ceil_div((int)T, rpc)), 256, 0, st, dkv, reinterpret_cast<__nv_bfloat16*>(dqkv_bf16),
                 d_qkv_bias, (int)T, C, rpc, scratch_zeroed);
  return check_launch("rvsa_kv_finalize_kernel");
}

// Attention backward of a window block AND the sampling heads' backward in three launches (scratch memset, the attention kernel, the fused tail):
// the caller gets dqkv (complete, incl. d_qkv_bias), dparams, the table / weight gradients accumulated, and dpooled [n_bw][C] (fp32) in
// sampling_workspace at offset n_bw * 5 * nH floats for mtp_layernorm_bwd(pool_add).  Same arithmetic as mtp_rvsa_attn_bwd + mtp_rvsa_sampling_bwd.
extern "C" int mtp_rvsa_attn_bwd_fused(const void* qkv_bf16, const float* params, const float* rel_pos_h, const float* rel_pos_w,
                                       const float* bias_table, const float* lse, const void* dout_bf16, void* dqkv_bf16, float* dparams,
                                       float* d_rel_pos_h, float* d_rel_pos_w, float* d_bias_table, float* d_qkv_bias, void* workspace,
                                       const float* pooled, const float* w_off, const float* w_scale, const float* w_angle, float* dw_off,
                                       float* db_off, float* dw_scale, float* db_scale, float* dw_angle, float* db_angle,
                                       void* sampling_workspace, int B, int h, int w, int C, int nH, mtp_stream_t stream) {
  MTP_REQUIRE(qkv_bf16 && params && rel_pos_h && rel_pos_w && bias_table && lse && dout_bf16 && dqkv_bf16 && dparams && d_rel_pos_h &&
                  d_rel_pos_w && d_bias_table && workspace && pooled && w_off && w_scale && w_angle && dw_off && db_off && dw_scale &&
                  db_scale && dw_angle && db_angle && sampling_workspace, "mtp_rvsa_attn_bwd_fused: null pointer");
  MTP_REQUIRE(B > 0 && h >= WS && w >= WS && C == nH * HD && nH % 2 == 0 && nH <= 64, "mtp_rvsa_attn_bwd_fused: B=%d h=%d w=%d C=%d nH=%d unsupported",
              B, h, w, C, nH);
  const RvsaGeom g = make_rvsa_geom(B, h, w, C, nH);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n_cta = B * g.nh * g.nw * nH, n_bw = n_cta / nH;
  float* part_rel = reinterpret_cast<float*>(workspace);
  float* part_table = part_rel + (size_t)n_cta * (2 * (2 * WS - 1) * HD);
  float* dkv = part_table + (size_t)n_cta * 169;
  const size_t T = (size_t)B * h * w;
  cudaError_t e = cudaMemsetAsync(dkv, 0, T * 2 * C * sizeof(float), st);
  if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "rvsa bwd memset: %s", cudaGetErrorString(e));
  int rc = launch_rvsa_attn_bwd_tc(qkv_bf16, params, rel_pos_h, rel_pos_w, bias_table, lse, dout_bf16, dqkv_bf16, dkv, dparams, part_rel, part_table,
                                   g, st);
  if (rc) return rc;
  RvsaTailArgs a;
  a.dkv = dkv; a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv_bf16); a.colsum = d_qkv_bias; a.T = (int)T; a.C = C; a.rezero = 0;
  a.kv_gx = ceil_div(3 * C, 1024);
  const int gy = std::max(1, std::min(ceil_div((int)T, 8), 4 * num_sms() / a.kv_gx));
  a.rows_per_cta = ceil_div((int)T, gy);
  a.kv_gy = ceil_div((int)T, a.rows_per_cta);
  a.part_rel = part_rel; a.part_table = part_table; a.d_rel_h = d_rel_pos_h; a.d_rel_w = d_rel_pos_w; a.d_table = d_bias_table;
  a.n_rel_parts = n_cta / 2; a.n_bw = n_bw; a.nH = nH;
  a.n_pr = 2 * (2 * WS - 1) * HD / 32 + nH * ((169 + 31) / 32);
  a.dparams = dparams; a.pooled = pooled; a.w_off = w_off; a.w_sc = w_scale; a.w_ang = w_angle;
  a.dpooled = reinterpret_cast<float*>(sampling_workspace) + (size_t)n_bw * 5 * nH;
  a.dw_off = dw_off; a.db_off = db_off; a.dw_sc = dw_scale; a.db_sc = db_scale; a.dw_ang = dw_angle; a.db_ang = db_angle;
  a.sb_gy = ceil_div(C, 256); a.wg_gx = ceil_div(C, 256);
  a.g = g;
  const int grid = a.kv_gx * a.kv_gy + n_bw * a.sb_gy + a.wg_gx * 5 * nH + a.n_pr;
  (void)launch_k(rvsa_bwd_tail_kernel, grid, 256, 0, st, a);
  return check_launch("rvsa_bwd_tail_kernel");
}

extern "C" int mtp_rvsa_sampling_bwd(const float* dparams, const float* pooled, const float* w_off, const float* w_scale,
                                     const float* w_angle, float* dw_off, float* db_off, float* dw_scale, float* db_scale,
                                     float* dw_angle, float* db_angle, void* dyn_bf16, void* workspace, int B, int h, int w, int C,
                                     int nH, mtp_stream_t stream) {
  MTP_REQUIRE(dparams && pooled && w_off && w_scale && w_angle && dw_off && db_off && dw_scale && db_scale && dw_angle && db_angle &&
                  workspace, "mtp_rvsa_sampling_bwd: null pointer");
  MTP_REQUIRE(B > 0 && h >= WS && w >= WS && C == nH * HD && nH <= 64, "mtp_rvsa_sampling_bwd: unsupported geometry");
  const RvsaGeom g = make_rvsa_geom(B, h, w, C, nH);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n_bw = B * g.nh * g.nw;
  float* g_out = reinterpret_cast<float*>(workspace);            // [n_bw][5nH]
  float* dpooled = g_out + (size_t)n_bw * 5 * nH;                // [n_bw][C]
  const int fused_smem = (n_bw * (5 * nH + 2 * SB_SLAB) + 5 * nH * SB_SLAB) * (int)sizeof(float);
  if ((sampling_fused_mask() & 2) && fused_smem <= 160 * 1024) {
    static int attr = 0;
    if (fused_smem > attr) {
      cudaError_t e = cudaFuncSetAttribute(rvsa_sampling_fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem);
      if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "rvsa_sampling_fused_bwd smem attr: %s", cudaGetErrorString(e));
      attr = fused_smem;
    }
    (void)launch_k(rvsa_sampling_fused_bwd_kernel, ceil_div(C, SB_SLAB), 256, fused_smem, st, dparams, pooled, w_off, w_scale, w_angle, dw_off, db_off,
                   dw_scale, db_scale, dw_angle, db_angle, dpooled, n_bw, g);
    int rc = check_launch("rvsa_sampling_fused_bwd_kernel");
    if (rc || dyn_bf16 == nullptr) return rc;
    const size_t total = (size_t)B * h * w * (C / 4);
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 8);
    (void)launch_k(rvsa_pool_bwd_add_kernel, grid, 256, 0, st, dpooled, reinterpret_cast<__nv_bfloat16*>(dyn_bf16), g);
    return check_launch("rvsa_pool_bwd_add_kernel");
  }
  (void)launch_k(rvsa_sampling_bwd_kernel, dim3(n_bw, ceil_div(C, 256)), 256, 0, st, dparams, pooled, w_off, w_scale, w_angle, g_out, dpooled, g);
  int rc = check_launch("rvsa_sampling_bwd_kernel");
  if (rc) return rc;
  (void)launch_k(rvsa_sampling_wgrad_kernel, dim3(ceil_div(C, 256), 5 * nH), 256, 0, st, dparams, pooled, dw_off, db_off, dw_scale, db_scale, dw_angle,
                 db_angle, n_bw, g);
  rc = check_launch("rvsa_sampling_wgrad_kernel");
  if (rc || dyn_bf16 == nullptr) return rc;      // no dyn: the caller hands dpooled (workspace + n_bw*5*nH floats) to mtp_layernorm_bwd
  const size_t total = (size_t)B * h * w * (C / 4);
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 8);
  (void)launch_k(rvsa_pool_bwd_add_kernel, grid, 256, 0, st, dpooled, reinterpret_cast<__nv_bfloat16*>(dyn_bf16), g);
  return check_launch("rvsa_pool_bwd_add_kernel");
}

extern "C" size_t mtp_rvsa_sampling_bwd_workspace_bytes(int B, int h, int w, int C, int nH) {
  const RvsaGeom g = make_rvsa_geom(B, h, w, C, nH);
  return (size_t)B * g.nh * g.nw * (5 * nH + C) * sizeof(float);
}
