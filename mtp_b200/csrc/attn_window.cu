// Rotated varied-size window attention (RVSA), forward.      [V]:287-433, SURVEY.md Appendix A.1
//
//   rvsa_sampling_fwd : per (image, window): zero-padded 7x7 mean of the LN'd tokens -> LeakyReLU(0.01) ->
//                       three 1x1 convs (GEMVs) -> (ox, oy, sx, sy, theta) per head                     [V]:228-243,354-368
//   rvsa_attn_fwd     : per (image, window, head): sampling coordinates in registers -> 4-tap bilinear gather of the
//                       K and V rows (one 128-byte line per tap from the token-major qkv matrix) blended in fp32 into
//                       shared memory -> S = scale q.k~ + q.Rh + q.Rw + bias-table -> softmax -> P v~ -> token-major
//                       bf16 output (padding rows are never written: crop is free).                      [V]:372-428
//
// Tokens live in the [T, 3C] bf16 output of the qkv GEMM (q | k | v, each head-major with hd = 64), so a K or V tap of one
// head is one contiguous 128-byte line and no window partition / pad / permute copy is ever materialised.
#include "common.h"
#include "ptx.cuh"
#include "rvsa_geom.cuh"

namespace mtp {

// ------------------------------------------------------------------------------------------------ sampling params
// (1) zero-padded 7x7 mean of the LN'd tokens: CTA = (image-window, 256-channel slab); thread = (token group of 4, 4 channels):
//     the <= 13 token loads of a thread are all in flight at once, the four groups combine through shared memory
__global__ void __launch_bounds__(256)
rvsa_pool_fwd_kernel(const __nv_bfloat16* __restrict__ yn, float* __restrict__ pooled, const RvsaGeom g, int ld, int lo) {
  MTP_PDL_ENTRY();
  __shared__ float4 part[4][64];
  const int bw = blockIdx.x;
  const int b = bw / (g.nh * g.nw), win = bw % (g.nh * g.nw);
  const int wy = win / g.nw, wx = win % g.nw;
  const int tg = threadIdx.x >> 6, cq = threadIdx.x & 63;
  const int c = blockIdx.y * 256 + cq * 4;
  float4 s = make_float4(0, 0, 0, 0);
  if (c < g.C) {
#pragma unroll
    for (int k = 0; k < (WS * WS + 3) / 4; ++k) {
      const int i = tg + 4 * k;
      const int y = wy * WS + i / WS - g.pt, x = wx * WS + i % WS - g.pl;
      if (i < WS * WS && y >= 0 && y < g.h && x >= 0 && x < g.w) {
        const __nv_bfloat16* src = yn + ((size_t)(b * g.h + y) * g.w + x) * ld + c;      // ld = C, or 2C with the lo words at +lo (fp32-class mode)
        const uint2 u = *reinterpret_cast<const uint2*>(src);
        const float2 a = unpack_bf16x2(u.x), d = unpack_bf16x2(u.y);
        s.x += a.x; s.y += a.y; s.z += d.x; s.w += d.y;
        if (lo > 0) {
          const uint2 ul = *reinterpret_cast<const uint2*>(src + lo);
          const float2 al = unpack_bf16x2(ul.x), dl = unpack_bf16x2(ul.y);
          s.x += al.x; s.y += al.y; s.z += dl.x; s.w += dl.y;
        }
      }
    }
  }
  part[tg][cq] = s;
  __syncthreads();
  if (tg == 0 && c < g.C) {
    const float4 p1 = part[1][cq], p2 = part[2][cq], p3 = part[3][cq];
    const float inv = 1.0f / (WS * WS);                    // zeros of the padding are part of the mean ([V]:347,354)
    s.x = (s.x + p1.x + p2.x + p3.x) * inv; s.y = (s.y + p1.y + p2.y + p3.y) * inv;
    s.z = (s.z + p1.z + p2.z + p3.z) * inv; s.w = (s.w + p1.w + p2.w + p3.w) * inv;
    *reinterpret_cast<float4*>(pooled + (size_t)bw * g.C + c) = s;
  }
}

// (2) the three 1x1 convs on LeakyReLU(pooled): CTA = one of the 5nH output channels, its weight row kept in registers,
//     warps stride over the (image, window) rows
__global__ void __launch_bounds__(256)
rvsa_heads_fwd_kernel(const float* __restrict__ pooled, const float* __restrict__ w_off, const float* __restrict__ b_off,
                      const float* __restrict__ w_sc, const float* __restrict__ b_sc, const float* __restrict__ w_ang,
                      const float* __restrict__ b_ang, float* __restrict__ params, int n_bw, const RvsaGeom g) {
  MTP_PDL_ENTRY();
  const int o = blockIdx.x, nH = g.nH, C = g.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // output order: [0,2nH) offsets (head-major, x then y), [2nH,4nH) scales, [4nH,5nH) angle
  const float* wrow;
  float bias;
  int n, slot;
  float div = 1.0f;
  if (o < 2 * nH) { wrow = w_off + (size_t)o * C; bias = b_off[o]; n = o >> 1; slot = o & 1; div = (float)((slot == 0 ? g.h : g.w) / WS); }   // sic: x by h//7, y by w//7
  else if (o < 4 * nH) { wrow = w_sc + (size_t)(o - 2 * nH) * C; bias = b_sc[o - 2 * nH]; n = (o - 2 * nH) >> 1; slot = 2 + ((o - 2 * nH) & 1); }
  else { wrow = w_ang + (size_t)(o - 4 * nH) * C; bias = b_ang[o - 4 * nH]; n = o - 4 * nH; slot = 4; }
  float4 wv[8];                                           // C <= 1024: 8 float4 per lane
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] = (i * 128 + lane * 4 < C) ? __ldg(reinterpret_cast<const float4*>(wrow + i * 128 + lane * 4)) : make_float4(0, 0, 0, 0);
  for (int bw = warp; bw < n_bw; bw += 8) {
    const float* pr = pooled + (size_t)bw * C;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i * 128 + lane * 4 < C) {
        float4 a = *reinterpret_cast<const float4*>(pr + i * 128 + lane * 4);
        a.x = a.x >= 0 ? a.x : 0.01f * a.x; a.y = a.y >= 0 ? a.y : 0.01f * a.y;
        a.z = a.z >= 0 ? a.z : 0.01f * a.z; a.w = a.w >= 0 ? a.w : 0.01f * a.w;
        s += wv[i].x * a.x + wv[i].y * a.y + wv[i].z * a.z + wv[i].w * a.w;
      }
    }
    s = warp_sum(s) + bias;
    if (lane == 0) params[((size_t)bw * nH + n) * 8 + slot] = s / div;
    if (slot == 4 && lane >= 1 && lane < 4) params[((size_t)bw * nH + n) * 8 + 4 + lane] = 0.f;      // unused slots 5..7
  }
}

// (1)+(2) in ONE launch: CTA = (image-window, group of 16 output channels), 1024 threads.  Every CTA of a window recomputes the window's
// pooled vector (49 token rows of <= 2 KB, all 13 loads of a thread in flight: the whole phase is one memory round trip), then two warps
// per output channel take half of the dot product each.  Replaces a 128-CTA pooling launch + an 80-CTA GEMV launch whose in-situ cost was
// 23 us per block (tools/step_breakdown.py) for ~0.2 MFLOP.
constexpr int SF_THREADS = 1024, SF_OUT = 16;
__global__ void __launch_bounds__(SF_THREADS)
rvsa_sampling_fused_fwd_kernel(const __nv_bfloat16* __restrict__ yn, const float* __restrict__ w_off, const float* __restrict__ b_off,
                               const float* __restrict__ w_sc, const float* __restrict__ b_sc, const float* __restrict__ w_ang,
                               const float* __restrict__ b_ang, float* __restrict__ pooled, float* __restrict__ params, const RvsaGeom g,
                               int ld, int lo) {
  MTP_PDL_ENTRY();
  __shared__ float4 part[4][256];
  __shared__ float pooled_s[1024];
  __shared__ float half_s[SF_OUT];
  const int bw = blockIdx.x;
  const int b = bw / (g.nh * g.nw), win = bw % (g.nh * g.nw);
  const int wy = win / g.nw, wx = win % g.nw;
  const int tid = threadIdx.x, tg = tid >> 8, cq = tid & 255;
  const int c = cq * 4, C = g.C, nH = g.nH;
  // ---- the three 1x1 convs: output order [0,2nH) offsets (head-major, x then y), [2nH,4nH) scales, [4nH,5nH) angle
  const int warp = tid >> 5, lane = tid & 31;
  const int k = warp >> 1, hf = warp & 1;
  const int o = blockIdx.y * SF_OUT + k;
  const bool live = o < 5 * nH;
  const float* wrow = nullptr;
  float bias = 0.f, div = 1.0f;
  int n = 0, slot = 0;
  if (live) {
    if (o < 2 * nH) { wrow = w_off + (size_t)o * C; bias = b_off[o]; n = o >> 1; slot = o & 1; div = (float)((slot == 0 ? g.h : g.w) / WS); }   // sic: x by h//7, y by w//7
    else if (o < 4 * nH) { wrow = w_sc + (size_t)(o - 2 * nH) * C; bias = b_sc[o - 2 * nH]; n = (o - 2 * nH) >> 1; slot = 2 + ((o - 2 * nH) & 1); }
    else { wrow = w_ang + (size_t)(o - 4 * nH) * C; bias = b_ang[o - 4 * nH]; n = o - 4 * nH; slot = 4; }
  }
  // element offsets of the window's 49 token rows (-1: padding), computed once per CTA: the per-row index arithmetic (two divisions by 7, bounds,
  // a 64-bit multiply) was most of the kernel's instructions (ncu: 860 per warp, issue slots 56 % busy)
  __shared__ int rowoff_s[52];
  if (tid < 52) {
    const int y = wy * WS + tid / WS - g.pt, x = wx * WS + tid % WS - g.pl;
    const bool ok = tid < WS * WS && y >= 0 && y < g.h && x >= 0 && x < g.w;
    rowoff_s[tid] = ok ? ((b * g.h + y) * g.w + x) * ld : -1;      // < 2^31 elements (B * h * w * ld)
  }
  __syncthreads();
  float4 s = make_float4(0, 0, 0, 0);
  if (c < C) {
    // all 13 row loads of this thread are issued before the first one is consumed (written as load-then-accumulate per row the compiler kept
    // them in program order: 13 serialised L2 round trips)
    constexpr int NR = (WS * WS + 3) / 4;
    const __nv_bfloat16* org = yn + c;
#define MTP_POOL_PASS(OFF)                                                                                        \
    {                                                                                                               \
      uint2 v[NR];                                                                                                  \
      _Pragma("unroll") for (int kk = 0; kk < NR; ++kk) {                                                           \
        const int off = rowoff_s[tg + 4 * kk];                                                                      \
        v[kk] = off >= 0 ? __ldg(reinterpret_cast<const uint2*>(org + off + (OFF))) : make_uint2(0u, 0u);           \
      }                                                                                                             \
      _Pragma("unroll") for (int kk = 0; kk < NR; ++kk) {                                                           \
        const float2 a = unpack_bf16x2(v[kk].x), d = unpack_bf16x2(v[kk].y);                                        \
        s.x += a.x; s.y += a.y; s.z += d.x; s.w += d.y;                                                             \
      }                                                                                                             \
    }
    MTP_POOL_PASS(0)
    if (lo > 0) MTP_POOL_PASS(lo)      // the lo words of the fp32-class mode
#undef MTP_POOL_PASS
  }
  part[tg][cq] = s;
  __syncthreads();
  if (tg == 0 && c < C) {
    const float4 p1 = part[1][cq], p2 = part[2][cq], p3 = part[3][cq];
    const float inv = 1.0f / (WS * WS);                    // zeros of the padding are part of the mean ([V]:347,354)
    s.x = (s.x + p1.x + p2.x + p3.x) * inv; s.y = (s.y + p1.y + p2.y + p3.y) * inv;
    s.z = (s.z + p1.z + p2.z + p3.z) * inv; s.w = (s.w + p1.w + p2.w + p3.w) * inv;
    if (blockIdx.y == 0 && pooled != nullptr) *reinterpret_cast<float4*>(pooled + (size_t)bw * C + c) = s;      // saved for the backward
    *reinterpret_cast<float4*>(pooled_s + c) = s;
  }
  __syncthreads();
  // ---- the three 1x1 convs on LeakyReLU(pooled)
  float acc = 0.f;
  if (live) {
    const int cbeg = hf * 512;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cc = cbeg + i * 128 + lane * 4;
      if (cc < C) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + cc));
        float4 a = *reinterpret_cast<const float4*>(pooled_s + cc);
        a.x = a.x >= 0 ? a.x : 0.01f * a.x; a.y = a.y >= 0 ? a.y : 0.01f * a.y;
        a.z = a.z >= 0 ? a.z : 0.01f * a.z; a.w = a.w >= 0 ? a.w : 0.01f * a.w;
        acc += w4.x * a.x + w4.y * a.y + w4.z * a.z + w4.w * a.w;
      }
    }
    acc = warp_sum(acc);
    if (hf == 1 && lane == 0) half_s[k] = acc;
  }
  __syncthreads();
  if (live && hf == 0) {
    if (lane == 0) params[((size_t)bw * nH + n) * 8 + slot] = (acc + half_s[k] + bias) / div;
    if (slot == 4 && lane >= 1 && lane < 4) params[((size_t)bw * nH + n) * 8 + 4 + lane] = 0.f;      // unused slots 5..7
  }
}

static int launch_rvsa_sampling_fused_fwd(const void* yn, const float* w_off, const float* b_off, const float* w_sc, const float* b_sc,
                                          const float* w_ang, const float* b_ang, float* pooled, float* params, const RvsaGeom& g, int ld, int lo,
                                          cudaStream_t st) {
  const int n_bw = g.B * g.nh * g.nw;
  (void)launch_k(rvsa_sampling_fused_fwd_kernel, dim3(n_bw, ceil_div(5 * g.nH, SF_OUT)), SF_THREADS, 0, st, reinterpret_cast<const __nv_bfloat16*>(yn),
                 w_off, b_off, w_sc, b_sc, w_ang, b_ang, pooled, params, g, ld, lo);
  return check_launch("rvsa_sampling_fused_fwd_kernel");
}

// ------------------------------------------------------------------------------------------------ attention fwd
constexpr int LDS_ROW = 68;      // padded fp32 row stride (floats) for Q/K/V tiles: conflict-free float4 row access
constexpr int LDP = 52;          // row stride of the 49x49 score tile
constexpr int RVSA_SMEM_FLOATS = 3 * NTOK * LDS_ROW + NTOK * LDP + 2 * NTOK * 8 + 2 * NTOK;

// HILO (fp32-class mode): every qkv / out value is a pair of bf16 words, hi at column c and lo at column c + 3C (qkv, row pitch 6C) /
// c + C (out, row pitch 2C); the arithmetic below is fp32 throughout, so this kernel is the exact-precision attention of that mode.
template <bool HILO>
__global__ void __launch_bounds__(64)
rvsa_attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ params, const float* __restrict__ rel_h,
                     const float* __restrict__ rel_w, const float* __restrict__ bias_table, __nv_bfloat16* __restrict__ out,
                     float* __restrict__ lse, const RvsaGeom g) {
  MTP_PDL_ENTRY();
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + NTOK * LDS_ROW;
  float* Vs = Ks + NTOK * LDS_ROW;
  float* Ps = Vs + NTOK * LDS_ROW;
  float* relh = Ps + NTOK * LDP;       // [49][8]
  float* relw = relh + NTOK * 8;
  float* cpx = relw + NTOK * 8;        // [49]
  float* cpy = cpx + NTOK;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x % g.nH;
  const int bw = blockIdx.x / g.nH;
  const int nwin = g.nh * g.nw;
  const int b = bw / nwin, win = bw % nwin;
  const int wy = win / g.nw, wx = win % g.nw;
  const int C = g.C, C3 = (HILO ? 6 : 3) * g.C;      // row pitch of qkv
  const int LOQ = 3 * g.C;                             // offset of the lo words (HILO)
  const float scale = 0.125f;          // hd^-0.5, hd = 64

  if (tid < NTOK) {
    const float* p = params + ((size_t)bw * g.nH + n) * 8;
    float px, py;
    rvsa_sample_coord(g, wy, wx, tid / WS, tid % WS, p[0], p[1], p[2], p[3], p[4], px, py);
    cpx[tid] = px;
    cpy[tid] = py;
  }
  __syncthreads();

  const __nv_bfloat16* qkv_b = qkv + (size_t)b * g.h * g.w * C3 + n * HD;
  for (int j = warp; j < NTOK; j += 2) {
    // ---- q row of window token j (zero for padding positions)
    {
      const int y = wy * WS + j / WS - g.pt, x = wx * WS + j % WS - g.pl;
      float2 qv = make_float2(0.f, 0.f);
      if (y >= 0 && y < g.h && x >= 0 && x < g.w) {
        const __nv_bfloat16* src = qkv_b + (size_t)(y * g.w + x) * C3 + lane * 2;
        qv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src));
        if (HILO) { const float2 l = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + LOQ)); qv.x += l.x; qv.y += l.y; }
      }
      *reinterpret_cast<float2*>(Qs + j * LDS_ROW + lane * 2) = qv;
    }
    // ---- bilinear gather of k~, v~ at sample j (grid_sample: bilinear, zeros, align_corners=True)
    const float px = cpx[j], py = cpy[j];
    const float fx0 = floorf(px), fy0 = floorf(py);
    const float ax = px - fx0, ay = py - fy0;
    const int x0 = (int)fx0 - g.pl, y0 = (int)fy0 - g.pt;        // tap coords in the un-padded grid
    float2 ka = make_float2(0.f, 0.f), va = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
      const float wgt = ((t & 1) ? ax : 1.f - ax) * ((t >> 1) ? ay : 1.f - ay);
      // taps in the zero padding or outside the padded grid contribute 0 (pad is applied after the bias, [V]:392)
      if (xx >= 0 && xx < g.w && yy >= 0 && yy < g.h) {
        const __nv_bfloat16* src = qkv_b + (size_t)(yy * g.w + xx) * C3 + lane * 2;
        float2 kv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + C));
        float2 vv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + 2 * C));
        if (HILO) {
          const float2 kl = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + C + LOQ));
          const float2 vl = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + 2 * C + LOQ));
          kv.x += kl.x; kv.y += kl.y; vv.x += vl.x; vv.y += vl.y;
        }
        ka.x += wgt * kv.x; ka.y += wgt * kv.y;
        va.x += wgt * vv.x; va.y += wgt * vv.y;
      }
    }
    *reinterpret_cast<float2*>(Ks + j * LDS_ROW + lane * 2) = ka;
    *reinterpret_cast<float2*>(Vs + j * LDS_ROW + lane * 2) = va;
  }
  __syncthreads();

  // ---- decomposed rel-pos: relh[q][kh] = q . rel_pos_h[qy - kh + 6], relw[q][kw] = q . rel_pos_w[qx - kw + 6]  (q UNscaled)
  for (int e = tid; e < NTOK * 14; e += 64) {
    const int q = e / 14, r = e % 14;
    const int kk = r % 7;
    const float* tab = (r < 7 ? rel_h : rel_w) + ((r < 7 ? q / WS : q % WS) - kk + WS - 1) * HD;
    const float* qr = Qs + q * LDS_ROW;
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

  // ---- scores: thread (tq, tj) owns the 7x7 micro-tile q in [7tq, 7tq+7), j in [7tj, 7tj+7)
  if (tid < NTOK) {
    const int tq = tid / WS, tj = tid % WS;
    float acc[WS][WS];
#pragma unroll
    for (int a = 0; a < WS; ++a)
#pragma unroll
      for (int c = 0; c < WS; ++c) acc[a][c] = 0.f;
    for (int d = 0; d < HD; d += 4) {
      float4 qv[WS], kv[WS];
#pragma unroll
      for (int a = 0; a < WS; ++a) qv[a] = *reinterpret_cast<const float4*>(Qs + (tq * WS + a) * LDS_ROW + d);
#pragma unroll
      for (int c = 0; c < WS; ++c) kv[c] = *reinterpret_cast<const float4*>(Ks + (tj * WS + c) * LDS_ROW + d);
#pragma unroll
      for (int a = 0; a < WS; ++a)
#pragma unroll
        for (int c = 0; c < WS; ++c)
          acc[a][c] += qv[a].x * kv[c].x + qv[a].y * kv[c].y + qv[a].z * kv[c].z + qv[a].w * kv[c].w;
    }
    // q = (qy = tq, qx = a), key j = (jy = tj, jx = c)
#pragma unroll
    for (int a = 0; a < WS; ++a) {
      const int q = tq * WS + a;
#pragma unroll
      for (int c = 0; c < WS; ++c) {
        const int idx = (tq - tj + WS - 1) * (2 * WS - 1) + (a - c + WS - 1);
        Ps[q * LDP + tj * WS + c] = scale * acc[a][c] + relh[q * 8 + tj] + relw[q * 8 + c] + __ldg(bias_table + idx * g.nH + n);
      }
    }
  }
  __syncthreads();

  // ---- softmax over keys, one thread per query row
  if (tid < NTOK) {
    float* row = Ps + tid * LDP;
    float m = -INFINITY;
    for (int j = 0; j < NTOK; ++j) m = fmaxf(m, row[j]);
    float s = 0.f;
    for (int j = 0; j < NTOK; ++j) {
      const float e = __expf(row[j] - m);
      row[j] = e;
      s += e;
    }
    const float inv = 1.0f / s;
    for (int j = 0; j < NTOK; ++j) row[j] *= inv;
    if (lse) lse[(size_t)blockIdx.x * NTOK + tid] = m + __logf(s);
  }
  __syncthreads();

  // ---- O = P v~ : thread (tq, td) owns rows [7tq, 7tq+7) x dims [4td, 4td+4) and [32+4td, 32+4td+4)
  if (tid < 56) {
    const int tq = tid >> 3, td = tid & 7;
    float acc[WS][8];
#pragma unroll
    for (int a = 0; a < WS; ++a)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
    for (int j = 0; j < NTOK; ++j) {
      const float4 v0 = *reinterpret_cast<const float4*>(Vs + j * LDS_ROW + td * 4);
      const float4 v1 = *reinterpret_cast<const float4*>(Vs + j * LDS_ROW + 32 + td * 4);
#pragma unroll
      for (int a = 0; a < WS; ++a) {
        const float p = Ps[(tq * WS + a) * LDP + j];
        acc[a][0] += p * v0.x; acc[a][1] += p * v0.y; acc[a][2] += p * v0.z; acc[a][3] += p * v0.w;
        acc[a][4] += p * v1.x; acc[a][5] += p * v1.y; acc[a][6] += p * v1.z; acc[a][7] += p * v1.w;
      }
    }
#pragma unroll
    for (int a = 0; a < WS; ++a) {
      const int y = wy * WS + tq - g.pt, x = wx * WS + a - g.pl;
      if (y >= 0 && y < g.h && x >= 0 && x < g.w) {
        uint2 u0, u1;
        u0.x = pack_bf16x2(acc[a][0], acc[a][1]);
        u0.y = pack_bf16x2(acc[a][2], acc[a][3]);
        u1.x = pack_bf16x2(acc[a][4], acc[a][5]);
        u1.y = pack_bf16x2(acc[a][6], acc[a][7]);
        __nv_bfloat16* orow = out + ((size_t)(b * g.h + y) * g.w + x) * (HILO ? 2 * C : C) + n * HD;
        *reinterpret_cast<uint2*>(orow + td * 4) = u0;
        *reinterpret_cast<uint2*>(orow + 32 + td * 4) = u1;
        if (HILO) {
          float r[8];
          const uint32_t w4[4] = {u0.x, u0.y, u1.x, u1.y};
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float2 h2 = unpack_bf16x2(w4[e]); r[2 * e] = acc[a][2 * e] - h2.x; r[2 * e + 1] = acc[a][2 * e + 1] - h2.y; }
          uint2 l0, l1;
          l0.x = pack_bf16x2(r[0], r[1]); l0.y = pack_bf16x2(r[2], r[3]);
          l1.x = pack_bf16x2(r[4], r[5]); l1.y = pack_bf16x2(r[6], r[7]);
          *reinterpret_cast<uint2*>(orow + C + td * 4) = l0;
          *reinterpret_cast<uint2*>(orow + C + 32 + td * 4) = l1;
        }
      }
    }
  }
}

int launch_rvsa_attn_fwd_tc(const void* qkv, const float* params, const float* rel_h, const float* rel_w, const float* table, void* out,
                            float* lse, const RvsaGeom& g, cudaStream_t st);     // attn_window_tc.cu

template <bool HILO>
static int launch_rvsa_attn_fwd_simt(const void* qkv, const float* params, const float* rel_h, const float* rel_w, const float* table, void* out,
                                     float* lse, const RvsaGeom& g, cudaStream_t st) {
  static bool attr = false;
  const int smem = RVSA_SMEM_FLOATS * sizeof(float);
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(rvsa_attn_fwd_kernel<HILO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "rvsa_attn_fwd smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  (void)launch_k(rvsa_attn_fwd_kernel<HILO>, g.B * g.nh * g.nw * g.nH, 64, smem, st, reinterpret_cast<const __nv_bfloat16*>(qkv), params, rel_h,
                 rel_w, table, reinterpret_cast<__nv_bfloat16*>(out), lse, g);
  return check_launch("rvsa_attn_fwd_kernel");
}

}  // namespace mtp

using namespace mtp;

extern "C" int mtp_rvsa_sampling_fwd(const void* yn_bf16, const float* w_off, const float* b_off, const float* w_scale,
                                     const float* b_scale, const float* w_angle, const float* b_angle, float* pooled, float* params,
                                     int B, int h, int w, int C, int nH, mtp_stream_t stream) {
  MTP_REQUIRE(yn_bf16 && w_off && b_off && w_scale && b_scale && w_angle && b_angle && pooled && params, "mtp_rvsa_sampling_fwd: null pointer");
  MTP_REQUIRE(B > 0 && h >= WS && w >= WS && C == nH * HD && C <= 1024, "mtp_rvsa_sampling_fwd: B=%d h=%d w=%d C=%d nH=%d unsupported (need h,w>=7, C==64*nH<=1024)", B, h, w, C, nH);
  const RvsaGeom g = make_rvsa_geom(B, h, w, C, nH);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((sampling_fused_mask() & 1) && (size_t)B * h * w * C < ((size_t)1 << 31))      // the fused kernel keeps 32-bit row offsets
    return launch_rvsa_sampling_fused_fwd(yn_bf16, w_off, b_off, w_scale, b_scale, w_angle, b_angle, pooled, params, g, C, 0, st);
  const int n_bw = B * g.nh * g.nw;
  (void)launch_k(rvsa_pool_fwd_kernel, dim3(n_bw, ceil_div(C, 256)), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(yn_bf16), pooled, g, C, 0);
  int rc = check_launch("rvsa_pool_fwd_kernel");
  if (rc) return rc;
  (void)launch_k(rvsa_heads_fwd_kernel, 5 * nH, 256, 0, st, pooled, w_off, b_off, w_scale, b_scale, w_angle, b_angle, params, n_bw, g);
  return check_launch("rvsa_heads_fwd_kernel");
}

extern "C" int mtp_rvsa_attn_fwd(const void* qkv_bf16, const float* params, const float* rel_pos_h, const float* rel_pos_w,
                                 const float* bias_table, void* out_bf16, float* lse, int B, int h, int w, int C, int nH,
                                 mtp_stream_t stream) {
  MTP_REQUIRE(qkv_bf16 && params && rel_pos_h && rel_pos_w && bias_table && out_bf16, "mtp_rvsa_attn_fwd: null pointer");
  MTP_REQUIRE(B > 0 && h >= WS && w >= WS && C == nH * HD, "mtp_rvsa_attn_fwd: B=%d h=%d w=%d C=%d nH=%d unsupported", B, h, w, C, nH);
  const RvsaGeom g = make_rvsa_geom(B, h, w, C, nH);
  if (nH % 2 == 0)       // tensor-core path: two heads of a window per 128-row UMMA tile
    return launch_rvsa_attn_fwd_tc(qkv_bf16, params, rel_pos_h, rel_pos_w, bias_table, out_bf16, lse, g, reinterpret_cast<cudaStream_t>(stream));
  return launch_rvsa_attn_fwd_simt<false>(qkv_bf16, params, rel_pos_h, rel_pos_w, bias_table, out_bf16, lse, g, reinterpret_cast<cudaStream_t>(stream));
}

/* fp32-class mode ("fp32x3"): the same operator on hi | lo word pairs (qkv [T, 6C], out [T, 2C]), fp32 arithmetic throughout;
 * yn of the sampling heads likewise ([T, 2C]). */
extern "C" int mtp_rvsa_sampling_fwd_hilo(const void* yn_hilo, const float* w_off, const float* b_off, const float* w_scale,
                                          const float* b_scale, const float* w_angle, const float* b_angle, float* pooled, float* params,
                                          int B, int h, int w, int C, int nH, mtp_stream_t stream) {
  MTP_REQUIRE(yn_hilo && w_off && b_off && w_scale && b_scale && w_angle && b_angle && pooled && params, "mtp_rvsa_sampling_fwd_hilo: null pointer");
  MTP_REQUIRE(B > 0 && h >= WS && w >= WS && C == nH * HD && C <= 1024, "mtp_rvsa_sampling_fwd_hilo: unsupported geometry");
  const RvsaGeom g = make_rvsa_geom(B, h, w, C, nH);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((sampling_fused_mask() & 1) && (size_t)B * h * w * 2 * C < ((size_t)1 << 31))
    return launch_rvsa_sampling_fused_fwd(yn_hilo, w_off, b_off, w_scale, b_scale, w_angle, b_angle, pooled, params, g, 2 * C, C, st);
  const int n_bw = B * g.nh * g.nw;
  (void)launch_k(rvsa_pool_fwd_kernel, dim3(n_bw, ceil_div(C, 256)), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(yn_hilo), pooled, g, 2 * C, C);
  int rc = check_launch("rvsa_pool_fwd_kernel");
  if (rc) return rc;
  (void)launch_k(rvsa_heads_fwd_kernel, 5 * nH, 256, 0, st, pooled, w_off, b_off, w_scale, b_scale, w_angle, b_angle, params, n_bw, g);
  return check_launch("rvsa_heads_fwd_kernel");
}

extern "C" int mtp_rvsa_attn_fwd_hilo(const void* qkv_hilo, const float* params, const float* rel_pos_h, const float* rel_pos_w,
                                      const float* bias_table, void* out_hilo, int B, int h, int w, int C, int nH, mtp_stream_t stream) {
  MTP_REQUIRE(qkv_hilo && params && rel_pos_h && rel_pos_w && bias_table && out_hilo, "mtp_rvsa_attn_fwd_hilo: null pointer");
  MTP_REQUIRE(B > 0 && h >= WS && w >= WS && C == nH * HD, "mtp_rvsa_attn_fwd_hilo: unsupported geometry");
  const RvsaGeom g = make_rvsa_geom(B, h, w, C, nH);
  return launch_rvsa_attn_fwd_simt<true>(qkv_hilo, params, rel_pos_h, rel_pos_w, bias_table, out_hilo, nullptr, g, reinterpret_cast<cudaStream_t>(stream));
}
