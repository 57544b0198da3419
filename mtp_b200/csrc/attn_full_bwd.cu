// Dense attention with decomposed rel-pos bias, backward (flash-style recompute; nothing N x N touches HBM).
// Autograd of Attention.forward [V]:90-111 + calc_rel_pos_spatial [V]:142-193.
//
//   full_attn_bwd_dq  : CTA per (image, head, 64-query tile).  D = rowsum(dO o O); loops over key tiles recomputing
//                       P = exp(S - lse); dS = P o (dO V^T - D); dq^ += dS K; accumulates the per-key-row / per-key-column
//                       sums of dS that feed the rel-pos terms; finally dq = scale (dq^ + dSh Rh + dSw Rw) and
//                       atomically accumulates d full_attn_rel_pos_{h,w} (already reduced over the tile).
//   full_attn_bwd_dkv : CTA per (image, head, 64-key tile).  Loops over query tiles: dV += P^T dO, dK += dS^T q^.
#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace mtp {

constexpr int FB_T = 64, FB_HD = 64, FB_LD = 68, FB_THREADS = 128;

// 64 x 64 bf16 tile (rows row0.., head slice) -> fp32 smem [64][FB_LD]; rows >= nvalid are zero; optional scale
__device__ __forceinline__ void fb_load_tile(float* dst, const __nv_bfloat16* src, int ld, int row0, int nrows, float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < FB_T; r += FB_THREADS / 32) {
    float2 v = make_float2(0.f, 0.f);
    if (row0 + r < nrows) {
      v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + (size_t)(row0 + r) * ld + lane * 2));
      v.x *= scale; v.y *= scale;
    }
    *reinterpret_cast<float2*>(dst + r * FB_LD + lane * 2) = v;
  }
}

// relh[r][c] = q^[r] . Rh[qy - c + gh - 1],  relw[r][c] = q^[r] . Rw[qx - c + gw - 1]
__device__ __forceinline__ void fb_relbias(const float* Qs, const float* rel_h, const float* rel_w, int q0, int N, int gh, int gw,
                                           float* relh, float* relw) {
  for (int e = threadIdx.x; e < FB_T * (gh + gw); e += FB_THREADS) {
    const int r = e / (gh + gw), c = e % (gh + gw);
    const int q = min(q0 + r, N - 1);
    const float* tab = c < gh ? rel_h + (size_t)(q / gw - c + gh - 1) * FB_HD : rel_w + (size_t)(q % gw - (c - gh) + gw - 1) * FB_HD;
    const float* qr = Qs + r * FB_LD;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < FB_HD; d += 4) {
      const float4 tv = __ldg(reinterpret_cast<const float4*>(tab + d));
      const float4 qv = *reinterpret_cast<const float4*>(qr + d);
      s += qv.x * tv.x + qv.y * tv.y + qv.z * tv.z + qv.w * tv.w;
    }
    if (c < gh) relh[r * gh + c] = s; else relw[r * gw + (c - gh)] = s;
  }
}

// s[a][c] = sum_d A[4tq+a][d] * B[tj+8c][d]
__device__ __forceinline__ void fb_mm_abt(const float* A, const float* B, int tq, int tj, float (&s)[4][8]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) s[a][c] = 0.f;
  for (int d = 0; d < FB_HD; d += 4) {
    float4 av[4], bv[8];
#pragma unroll
    for (int a = 0; a < 4; ++a) av[a] = *reinterpret_cast<const float4*>(A + (tq * 4 + a) * FB_LD + d);
#pragma unroll
    for (int c = 0; c < 8; ++c) bv[c] = *reinterpret_cast<const float4*>(B + (tj + 8 * c) * FB_LD + d);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 8; ++c) s[a][c] += av[a].x * bv[c].x + av[a].y * bv[c].y + av[a].z * bv[c].z + av[a].w * bv[c].w;
  }
}

// P and dS for one (query tile, key tile) pair, left in registers in the (tq, tj) micro-tile mapping
__device__ __forceinline__ void fb_p_ds(const float* Qs, const float* Ks, const float* Vs, const float* Gs, const float* relh,
                                        const float* relw, const float* lse_s, const float* D_s, int k0, int N, int gh, int gw,
                                        int use_rel, int tq, int tj, float (&p)[4][8], float (&ds)[4][8]) {
  fb_mm_abt(Qs, Ks, tq, tj, p);
  fb_mm_abt(Gs, Vs, tq, tj, ds);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = tq * 4 + a;
    const float l = lse_s[r], D = D_s[r];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = k0 + tj + 8 * c;
      if (j < N) {
        float s = p[a][c];
        if (use_rel) s += relh[r * gh + j / gw] + relw[r * gw + j % gw];
        const float pv = __expf(s - l);
        p[a][c] = pv;
        ds[a][c] = pv * (ds[a][c] - D);
      } else {
        p[a][c] = 0.f;
        ds[a][c] = 0.f;
      }
    }
  }
}

__global__ void __launch_bounds__(FB_THREADS)
full_attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                        const float* __restrict__ lse, const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                        __nv_bfloat16* __restrict__ dqkv, float* __restrict__ Dbuf, float* __restrict__ d_rel_h,
                        float* __restrict__ d_rel_w, int N, int gh, int gw, int C, int nH, int use_rel) {
  MTP_PDL_ENTRY();
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + FB_T * FB_LD;
  float* Vs = Ks + FB_T * FB_LD;
  float* Gs = Vs + FB_T * FB_LD;
  float* Ss = Gs + FB_T * FB_LD;               // dS tile
  float* lse_s = Ss + FB_T * FB_LD;            // [64]
  float* D_s = lse_s + FB_T;                   // [64]
  float* relh = D_s + FB_T;                    // [64][gh]
  float* relw = relh + FB_T * gh;              // [64][gw]
  float* dSh = relw + FB_T * gw;               // [64][gh]
  float* dSw = dSh + FB_T * gh;                // [64][gw]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * FB_T, n = blockIdx.y, b = blockIdx.z;
  const int C3 = 3 * C;
  const float scale = 0.125f;
  const __nv_bfloat16* base = qkv + (size_t)b * N * C3 + n * FB_HD;

  fb_load_tile(Qs, base, C3, q0, N, scale);
  fb_load_tile(Gs, dout + (size_t)b * N * C + n * FB_HD, C, q0, N, 1.0f);
  for (int e = tid; e < FB_T * (gh + gw); e += FB_THREADS) dSh[e] = 0.f;       // dSh and dSw are contiguous
  // D[q] = dO[q] . O[q]
  for (int r = warp; r < FB_T; r += FB_THREADS / 32) {
    float s = 0.f;
    if (q0 + r < N) {
      const size_t off = ((size_t)b * N + q0 + r) * C + n * FB_HD + lane * 2;
      const float2 o = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(out + off));
      const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + off));
      s = o.x * g.x + o.y * g.y;
    }
    s = warp_sum(s);
    if (lane == 0) {
      D_s[r] = s;
      lse_s[r] = q0 + r < N ? lse[((size_t)b * nH + n) * N + q0 + r] : 0.f;
      if (q0 + r < N) Dbuf[((size_t)b * nH + n) * N + q0 + r] = s;
    }
  }
  __syncthreads();
  if (use_rel) fb_relbias(Qs, rel_h, rel_w, q0, N, gh, gw, relh, relw);

  const int tq = tid >> 3, tj = tid & 7;
  float dq[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < 8; ++e) dq[a][e] = 0.f;

  for (int k0 = 0; k0 < N; k0 += FB_T) {
    __syncthreads();
    fb_load_tile(Ks, base + C, C3, k0, N, 1.0f);
    fb_load_tile(Vs, base + 2 * C, C3, k0, N, 1.0f);
    __syncthreads();
    float p[4][8], ds[4][8];
    fb_p_ds(Qs, Ks, Vs, Gs, relh, relw, lse_s, D_s, k0, N, gh, gw, use_rel, tq, tj, p, ds);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = tq * 4 + a;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int j = k0 + tj + 8 * c;
        Ss[r * FB_LD + tj + 8 * c] = ds[a][c];
        if (use_rel && j < N) {
          atomicAdd(&dSh[r * gh + j / gw], ds[a][c]);
          atomicAdd(&dSw[r * gw + j % gw], ds[a][c]);
        }
      }
    }
    __syncwarp();          // dS rows of this warp are produced and consumed by this warp
    for (int j = 0; j < FB_T; ++j) {
      const float4 k0v = *reinterpret_cast<const float4*>(Ks + j * FB_LD + tj * 4);
      const float4 k1v = *reinterpret_cast<const float4*>(Ks + j * FB_LD + 32 + tj * 4);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float s = Ss[(tq * 4 + a) * FB_LD + j];
        dq[a][0] += s * k0v.x; dq[a][1] += s * k0v.y; dq[a][2] += s * k0v.z; dq[a][3] += s * k0v.w;
        dq[a][4] += s * k1v.x; dq[a][5] += s * k1v.y; dq[a][6] += s * k1v.z; dq[a][7] += s * k1v.w;
      }
    }
  }
  __syncthreads();

  // rel-pos contribution to dq^, then dq = scale * dq^
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = tq * 4 + a, q = q0 + r;
    if (q >= N) continue;
    if (use_rel) {
      const int qy = q / gw, qx = q % gw;
      for (int k = 0; k < gh; ++k) {
        const float ch = dSh[r * gh + k];
        const float* th = rel_h + (size_t)(qy - k + gh - 1) * FB_HD;
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(th + tj * 4)), h1 = __ldg(reinterpret_cast<const float4*>(th + 32 + tj * 4));
        dq[a][0] += ch * h0.x; dq[a][1] += ch * h0.y; dq[a][2] += ch * h0.z; dq[a][3] += ch * h0.w;
        dq[a][4] += ch * h1.x; dq[a][5] += ch * h1.y; dq[a][6] += ch * h1.z; dq[a][7] += ch * h1.w;
      }
      for (int k = 0; k < gw; ++k) {
        const float cw = dSw[r * gw + k];
        const float* tw = rel_w + (size_t)(qx - k + gw - 1) * FB_HD;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(tw + tj * 4)), w1 = __ldg(reinterpret_cast<const float4*>(tw + 32 + tj * 4));
        dq[a][0] += cw * w0.x; dq[a][1] += cw * w0.y; dq[a][2] += cw * w0.z; dq[a][3] += cw * w0.w;
        dq[a][4] += cw * w1.x; dq[a][5] += cw * w1.y; dq[a][6] += cw * w1.z; dq[a][7] += cw * w1.w;
      }
    }
    __nv_bfloat16* dst = dqkv + ((size_t)b * N + q) * C3 + n * FB_HD;
    uint2 u0, u1;
    u0.x = pack_bf16x2(scale * dq[a][0], scale * dq[a][1]); u0.y = pack_bf16x2(scale * dq[a][2], scale * dq[a][3]);
    u1.x = pack_bf16x2(scale * dq[a][4], scale * dq[a][5]); u1.y = pack_bf16x2(scale * dq[a][6], scale * dq[a][7]);
    *reinterpret_cast<uint2*>(dst + tj * 4) = u0;
    *reinterpret_cast<uint2*>(dst + 32 + tj * 4) = u1;
  }

  // d rel tables: dR[r][d] += sum over (q in tile, k) with q_axis - k + g - 1 == r of dSx[q][k] * q^[q][d]
  if (use_rel) {
    const int nq = min(FB_T, N - q0);
    const int rows_h = 2 * gh - 1, rows_w = 2 * gw - 1;
    for (int e = tid; e < (rows_h + rows_w) * FB_HD; e += FB_THREADS) {
      const int d = e % FB_HD, rr = e / FB_HD;
      const bool is_h = rr < rows_h;
      const int r = is_h ? rr : rr - rows_h;
      float s = 0.f;
      bool any = false;
      for (int i = 0; i < nq; ++i) {
        const int q = q0 + i;
        const int k = is_h ? q / gw - (r - (gh - 1)) : q % gw - (r - (gw - 1));
        if (k >= 0 && k < (is_h ? gh : gw)) {
          s += (is_h ? dSh[i * gh + k] : dSw[i * gw + k]) * Qs[i * FB_LD + d];
          any = true;
        }
      }
      if (any) atomicAdd((is_h ? d_rel_h : d_rel_w) + (size_t)r * FB_HD + d, s);
    }
  }
}

__global__ void __launch_bounds__(FB_THREADS)
full_attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                         const float* __restrict__ lse, const float* __restrict__ Dbuf, const __nv_bfloat16* __restrict__ dout,
                         __nv_bfloat16* __restrict__ dqkv, int N, int gh, int gw, int C, int nH, int use_rel) {
  MTP_PDL_ENTRY();
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + FB_T * FB_LD;
  float* Vs = Ks + FB_T * FB_LD;
  float* Gs = Vs + FB_T * FB_LD;
  float* Ps = Gs + FB_T * FB_LD;
  float* Ss = Ps + FB_T * FB_LD;
  float* lse_s = Ss + FB_T * FB_LD;
  float* D_s = lse_s + FB_T;
  float* relh = D_s + FB_T;
  float* relw = relh + FB_T * gh;

  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * FB_T, n = blockIdx.y, b = blockIdx.z;
  const int C3 = 3 * C;
  const float scale = 0.125f;
  const __nv_bfloat16* base = qkv + (size_t)b * N * C3 + n * FB_HD;
  fb_load_tile(Ks, base + C, C3, k0, N, 1.0f);
  fb_load_tile(Vs, base + 2 * C, C3, k0, N, 1.0f);

  const int tq = tid >> 3, tj = tid & 7;       // also (key-row group, dim group) for the transposed products
  float dk[4][8], dv[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < 8; ++e) { dk[a][e] = 0.f; dv[a][e] = 0.f; }

  for (int q0 = 0; q0 < N; q0 += FB_T) {
    __syncthreads();
    fb_load_tile(Qs, base, C3, q0, N, scale);
    fb_load_tile(Gs, dout + (size_t)b * N * C + n * FB_HD, C, q0, N, 1.0f);
    if (tid < FB_T) {
      const bool ok = q0 + tid < N;
      lse_s[tid] = ok ? lse[((size_t)b * nH + n) * N + q0 + tid] : 0.f;
      D_s[tid] = ok ? Dbuf[((size_t)b * nH + n) * N + q0 + tid] : 0.f;
    }
    __syncthreads();
    if (use_rel) {
      fb_relbias(Qs, rel_h, rel_w, q0, N, gh, gw, relh, relw);
      __syncthreads();
    }
    float p[4][8], ds[4][8];
    fb_p_ds(Qs, Ks, Vs, Gs, relh, relw, lse_s, D_s, k0, N, gh, gw, use_rel, tq, tj, p, ds);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = tq * 4 + a;
      const bool rok = q0 + r < N;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        Ps[r * FB_LD + tj + 8 * c] = rok ? p[a][c] : 0.f;
        Ss[r * FB_LD + tj + 8 * c] = rok ? ds[a][c] : 0.f;
      }
    }
    __syncthreads();
    // dV[j][:] += sum_q P[q][j] dO[q][:] ; dK[j][:] += sum_q dS[q][j] q^[q][:]     (j = 4tq+a, dims of tj)
    for (int q = 0; q < FB_T; ++q) {
      const float4 pj = *reinterpret_cast<const float4*>(Ps + q * FB_LD + tq * 4);
      const float4 sj = *reinterpret_cast<const float4*>(Ss + q * FB_LD + tq * 4);
      const float4 g0 = *reinterpret_cast<const float4*>(Gs + q * FB_LD + tj * 4), g1 = *reinterpret_cast<const float4*>(Gs + q * FB_LD + 32 + tj * 4);
      const float4 x0 = *reinterpret_cast<const float4*>(Qs + q * FB_LD + tj * 4), x1 = *reinterpret_cast<const float4*>(Qs + q * FB_LD + 32 + tj * 4);
      const float pa[4] = {pj.x, pj.y, pj.z, pj.w}, sa[4] = {sj.x, sj.y, sj.z, sj.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        dv[a][0] += pa[a] * g0.x; dv[a][1] += pa[a] * g0.y; dv[a][2] += pa[a] * g0.z; dv[a][3] += pa[a] * g0.w;
        dv[a][4] += pa[a] * g1.x; dv[a][5] += pa[a] * g1.y; dv[a][6] += pa[a] * g1.z; dv[a][7] += pa[a] * g1.w;
        dk[a][0] += sa[a] * x0.x; dk[a][1] += sa[a] * x0.y; dk[a][2] += sa[a] * x0.z; dk[a][3] += sa[a] * x0.w;
        dk[a][4] += sa[a] * x1.x; dk[a][5] += sa[a] * x1.y; dk[a][6] += sa[a] * x1.z; dk[a][7] += sa[a] * x1.w;
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = k0 + tq * 4 + a;
    if (j >= N) continue;
    __nv_bfloat16* dst = dqkv + ((size_t)b * N + j) * C3 + n * FB_HD;
    uint2 u0, u1;
    u0.x = pack_bf16x2(dk[a][0], dk[a][1]); u0.y = pack_bf16x2(dk[a][2], dk[a][3]);
    u1.x = pack_bf16x2(dk[a][4], dk[a][5]); u1.y = pack_bf16x2(dk[a][6], dk[a][7]);
    *reinterpret_cast<uint2*>(dst + C + tj * 4) = u0;
    *reinterpret_cast<uint2*>(dst + C + 32 + tj * 4) = u1;
    u0.x = pack_bf16x2(dv[a][0], dv[a][1]); u0.y = pack_bf16x2(dv[a][2], dv[a][3]);
    u1.x = pack_bf16x2(dv[a][4], dv[a][5]); u1.y = pack_bf16x2(dv[a][6], dv[a][7]);
    *reinterpret_cast<uint2*>(dst + 2 * C + tj * 4) = u0;
    *reinterpret_cast<uint2*>(dst + 2 * C + 32 + tj * 4) = u1;
  }
}

int launch_full_attn_bwd_tc(const void* qkv, const float* rel_h, const float* rel_w, const float* lse, const void* out, const void* dout,
                            void* dqkv, float* d_rel_h, float* d_rel_w, int B, int gh, int gw, int C, int nH, cudaStream_t st);   // attn_full_tc.cu

}  // namespace mtp

using namespace mtp;

namespace mtp {
size_t full_attn_bwd_stream_workspace_bytes(int B, int gh, int gw, int nH);      // attn_full_stream_bwd_tc.cu
int launch_full_attn_bwd_stream_tc(const void* qkv, const float* rel_h, const float* rel_w, const float* lse, const void* out, const void* dout,
                                   void* dqkv, float* d_rel_h, float* d_rel_w, void* workspace, int B, int gh, int gw, int C, int nH,
                                   cudaStream_t st);
static bool dense_stream_all() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTP_DENSE_STREAM_ALL"); v = (e != nullptr && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
static bool dense_stream_bwd_ok(int gh, int gw) {
  static int simt = -1;
  if (simt < 0) { const char* e = getenv("MTP_DENSE_SIMT"); simt = (e != nullptr && e[0] == '1') ? 1 : 0; }      // A/B switch
  return !simt && 2 * gh - 1 <= 128 && 2 * gw - 1 <= 128 && (2 * gh - 1 + 2 * gw - 1) * 65 * 4 + 5 * 16384 + 64 <= 227 * 1024;
}
}  // namespace mtp

extern "C" size_t mtp_full_attn_bwd_workspace_bytes(int B, int gh, int gw, int nH) {
  const size_t simt = (size_t)B * nH * gh * gw * sizeof(float);
  const bool resident = gh * gw <= 256 && gh <= 16 && gw <= 16 && !mtp::dense_stream_all();      // must mirror the dispatch in mtp_full_attn_bwd
  return !resident && mtp::dense_stream_bwd_ok(gh, gw) ? std::max(simt, mtp::full_attn_bwd_stream_workspace_bytes(B, gh, gw, nH)) : simt;
}

extern "C" int mtp_full_attn_bwd(const void* qkv_bf16, const float* rel_pos_h, const float* rel_pos_w, const float* lse,
                                 const void* out_bf16, const void* dout_bf16, void* dqkv_bf16, float* d_rel_pos_h,
                                 float* d_rel_pos_w, void* workspace, int B, int gh, int gw, int C, int nH, mtp_stream_t stream) {
  MTP_REQUIRE(qkv_bf16 && lse && out_bf16 && dout_bf16 && dqkv_bf16 && workspace, "mtp_full_attn_bwd: null pointer");
  const bool use_rel = rel_pos_h != nullptr;
  MTP_REQUIRE((rel_pos_h == nullptr) == (rel_pos_w == nullptr), "mtp_full_attn_bwd: give both rel-pos tables or neither");
  MTP_REQUIRE(!use_rel || (d_rel_pos_h && d_rel_pos_w), "mtp_full_attn_bwd: rel-pos gradients requested without buffers");
  MTP_REQUIRE(B > 0 && gh > 0 && gw > 0 && C == nH * FB_HD, "mtp_full_attn_bwd: unsupported geometry");
  const int N = gh * gw;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N <= 256 && gh <= 16 && gw <= 16 && !dense_stream_all())       // tensor-core path (one CTA per image-head, K/V resident, dK/dV accumulated in TMEM)
    return launch_full_attn_bwd_tc(qkv_bf16, rel_pos_h, rel_pos_w, lse, out_bf16, dout_bf16, dqkv_bf16, d_rel_pos_h, d_rel_pos_w, B, gh, gw,
                                   C, nH, st);
  if (dense_stream_bwd_ok(gh, gw))      // long sequences on the tensor cores: K / V block resident, dK / dV in TMEM, dQ through red.global
    return launch_full_attn_bwd_stream_tc(qkv_bf16, rel_pos_h, rel_pos_w, lse, out_bf16, dout_bf16, dqkv_bf16, d_rel_pos_h, d_rel_pos_w, workspace, B, gh,
                                          gw, C, nH, st);
  float* Dbuf = reinterpret_cast<float*>(workspace);
  const int smem_dq = (5 * FB_T * FB_LD + 2 * FB_T + 2 * FB_T * (gh + gw)) * (int)sizeof(float);
  const int smem_dkv = (6 * FB_T * FB_LD + 2 * FB_T + FB_T * (gh + gw)) * (int)sizeof(float);
  MTP_REQUIRE(smem_dq <= 220 * 1024 && smem_dkv <= 220 * 1024, "mtp_full_attn_bwd: grid %dx%d too large for shared memory", gh, gw);
  static int attr_dq = 0, attr_dkv = 0;
  if (smem_dq > attr_dq) {
    cudaError_t e = cudaFuncSetAttribute(full_attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "full_attn_bwd_dq smem attr: %s", cudaGetErrorString(e));
    attr_dq = smem_dq;
  }
  if (smem_dkv > attr_dkv) {
    cudaError_t e = cudaFuncSetAttribute(full_attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "full_attn_bwd_dkv smem attr: %s", cudaGetErrorString(e));
    attr_dkv = smem_dkv;
  }
  const dim3 grid(ceil_div(N, FB_T), nH, B);
  (void)launch_k(full_attn_bwd_dq_kernel, grid, FB_THREADS, smem_dq, st, 
      reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), rel_pos_h, rel_pos_w, lse, reinterpret_cast<const __nv_bfloat16*>(out_bf16),
      reinterpret_cast<const __nv_bfloat16*>(dout_bf16), reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), Dbuf, d_rel_pos_h, d_rel_pos_w, N,
      gh, gw, C, nH, use_rel);
  int rc = check_launch("full_attn_bwd_dq_kernel");
  if (rc) return rc;
  (void)launch_k(full_attn_bwd_dkv_kernel, grid, FB_THREADS, smem_dkv, st, 
      reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), rel_pos_h, rel_pos_w, lse, Dbuf, reinterpret_cast<const __nv_bfloat16*>(dout_bf16),
      reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), N, gh, gw, C, nH, use_rel);
  return check_launch("full_attn_bwd_dkv_kernel");
}
