// HBM-bound row kernels: LayerNorm forward/backward (+fused residual-gradient add), casts, column sums.
// One warp per token row, 128-bit coalesced accesses, fp32 statistics.   [V]:484,496,508-509,576-584,596
#include <algorithm>

#include "common.h"
#include "ptx.cuh"
#include "rvsa_geom.cuh"

namespace mtp {

constexpr int ROW_WARPS = 8;   // warps (= rows in flight) per CTA

template <typename T> struct Vec4;   // 4 consecutive elements as fp32
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// ------------------------------------------------------------------------------------------------ LayerNorm fwd
// y = LN(x) * gamma + beta (optionally GELU'd), stats saved for backward.  NV = C / 128 float4 chunks per lane.
template <typename TIn, int NV, bool GELU>
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int rows, float eps) {
  MTP_PDL_ENTRY();
  constexpr int C = NV * 128;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const TIn* xr = x + (size_t)row * C;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = Vec4<TIn>::load(xr + (i * 32 + lane) * 4);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mu = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
    q += a * a + b * b + c * c + d * d;
  }
  const float rs = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
  __nv_bfloat16* yr = y + (size_t)row * C;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 o;
    o.x = (v[i].x - mu) * rs * g.x + b.x;
    o.y = (v[i].y - mu) * rs * g.y + b.y;
    o.z = (v[i].z - mu) * rs * g.z + b.z;
    o.w = (v[i].w - mu) * rs * g.w + b.w;
    if (GELU) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
    Vec4<__nv_bfloat16>::store(yr + c, o);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm bwd
// dx_out = dres (optional fp32 residual-path gradient) + LN'(dy)   where dy is the gradient w.r.t. the LN output
// (for the GELU variant: w.r.t. the GELU output; the pre-GELU value is recomputed from x).
// dgamma/dbeta: per-CTA partial sums over its rows (each lane owns fixed columns), cross-warp reduce in smem, one
// atomicAdd per column per CTA.
// CAST: additionally emits g16 = bf16(row_scale[row / rows_per_group] * dx_out) -- the A operand of the next dgrad / wgrad GEMMs
// (DropPath backward of the next residual branch) -- and accumulates its column sums (that branch's last bias gradient), which
// saves the separate scale_cast pass over the fp32 gradient.
// The CTA has blockDim.x / 32 warps, one row each per iteration; the host sizes it so that the rows fill the SMs in ONE wave.
constexpr int LN_BWD_MAX_WARPS = 12;

template <typename TIn, typename TDx, int NV, bool GELU, bool CAST>
__global__ void __launch_bounds__(LN_BWD_MAX_WARPS * 32)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const TIn* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
              const float* __restrict__ dres, TDx* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
              const float* __restrict__ row_scale, int rows_per_group, __nv_bfloat16* __restrict__ g16, float* __restrict__ colsum16,
              const float* __restrict__ pool_add, const RvsaGeom pg, int rows, int rows_per_cta) {
  constexpr int C = NV * 128;
  extern __shared__ float red[];          // [warps][C]
  MTP_PDL_ENTRY();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  float4 ag[NV], ab[NV], ac[CAST ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) { ag[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); }
#pragma unroll
  for (int i = 0; i < (CAST ? NV : 1); ++i) ac[i] = make_float4(0, 0, 0, 0);
  const int row0 = blockIdx.x * rows_per_cta;
  const int row1 = min(rows, row0 + rows_per_cta);
  for (int row = row0 + warp; row < row1; row += n_warps) {
    const float mu = mean[row], rs = rstd[row];
    const TIn* xr = x + (size_t)row * C;
    const __nv_bfloat16* dyr = dy + (size_t)row * C;
    // AvgPool backward of the RVSA sampling heads: every real token of a window receives dpooled[window] / 49 on top of dy
    const float* pa = nullptr;
    if (pool_add != nullptr) {
      const int xw = row % pg.w, yh = (row / pg.w) % pg.h, b = row / (pg.w * pg.h);
      pa = pool_add + (size_t)((b * pg.nh + (yh + pg.pt) / WS) * pg.nw + (xw + pg.pl) / WS) * C;
    }
    float4 xh[NV], g[NV], rres[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {             // the residual-path cotangent is fetched with the other operands (one latency round)
      if (dres != nullptr) rres[i] = *reinterpret_cast<const float4*>(dres + (size_t)row * C + (i * 32 + lane) * 4);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 xv = Vec4<TIn>::load(xr + c);
      float4 d = Vec4<__nv_bfloat16>::load(dyr + c);
      if (pa != nullptr) {
        const float4 pv = *reinterpret_cast<const float4*>(pa + c);
        const float inv = 1.0f / (WS * WS);
        d.x += pv.x * inv; d.y += pv.y * inv; d.z += pv.z * inv; d.w += pv.w * inv;
      }
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      if (GELU) {
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
        d.x *= gelu_erf_grad(xh[i].x * gm.x + bt.x);
        d.y *= gelu_erf_grad(xh[i].y * gm.y + bt.y);
        d.z *= gelu_erf_grad(xh[i].z * gm.z + bt.z);
        d.w *= gelu_erf_grad(xh[i].w * gm.w + bt.w);
      }
      ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += g[i].x + g[i].y + g[i].z + g[i].w;
      s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
    }
    s1 = warp_sum(s1) * (1.0f / C);
    s2 = warp_sum(s2) * (1.0f / C);
    TDx* dxr = dx + (size_t)row * C;
    const float sc = (CAST && row_scale != nullptr) ? __ldg(row_scale + row / rows_per_group) : 1.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 o;
      o.x = rs * (g[i].x - s1 - xh[i].x * s2);
      o.y = rs * (g[i].y - s1 - xh[i].y * s2);
      o.z = rs * (g[i].z - s1 - xh[i].z * s2);
      o.w = rs * (g[i].w - s1 - xh[i].w * s2);
      if (dres != nullptr) {
        const float4 r = rres[i];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      Vec4<TDx>::store(dxr + c, o);
      if (CAST) {
        o.x *= sc; o.y *= sc; o.z *= sc; o.w *= sc;
        Vec4<__nv_bfloat16>::store(g16 + (size_t)row * C + c, o);
        ac[i].x += o.x; ac[i].y += o.y; ac[i].z += o.z; ac[i].w += o.w;
      }
    }
  }
  // cross-warp reduction of the column sums: all (2 or 3) arrays go through shared memory at once ([array][warp][C]); a thread
  // then owns 4 consecutive columns of one array and issues ONE vector atomic for them
  constexpr int NARR = CAST ? 3 : 2;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    *reinterpret_cast<float4*>(&red[(0 * n_warps + warp) * C + c]) = ag[i];
    *reinterpret_cast<float4*>(&red[(1 * n_warps + warp) * C + c]) = ab[i];
    if (CAST) *reinterpret_cast<float4*>(&red[(2 * n_warps + warp) * C + c]) = ac[CAST ? i : 0];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NARR * (C / 4); e += blockDim.x) {
    const int arr = e / (C / 4), c = (e % (C / 4)) * 4;
    float* dst = arr == 0 ? dgamma : arr == 1 ? dbeta : colsum16;
    if (dst == nullptr) continue;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < n_warps; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(&red[(arr * n_warps + w) * C + c]);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ casts / sums
// out_bf16[r, c] = in_f32[r, c] * row_scale[r / rows_per_group];  optional column sums of the scaled values.
__global__ void __launch_bounds__(256)
scale_cast_kernel(const float* __restrict__ in, const float* __restrict__ row_scale, int rows_per_group,
                  __nv_bfloat16* __restrict__ out, float* __restrict__ colsum, int rows, int C, int rows_per_cta) {
  MTP_PDL_ENTRY();
  // thread owns 4 consecutive columns; CTA walks rows [row0, row1)
  const int row0 = blockIdx.y * rows_per_cta, row1 = min(rows, row0 + rows_per_cta);
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= C) return;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int r = row0; r < row1; ++r) {
    const float s = row_scale ? __ldg(row_scale + r / rows_per_group) : 1.0f;
    float4 v = *reinterpret_cast<const float4*>(in + (size_t)r * C + c);
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    Vec4<__nv_bfloat16>::store(out + (size_t)r * C + c, v);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (colsum)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum + c), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
}

// colsum[c] += sum_r in_bf16[r, c]
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ colsum, int rows, int C, int ld, int rows_per_cta) {
  MTP_PDL_ENTRY();
  const int row0 = blockIdx.y * rows_per_cta, row1 = min(rows, row0 + rows_per_cta);
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= C) return;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int r = row0; r < row1; ++r) {
    const float4 v = Vec4<__nv_bfloat16>::load(in + (size_t)r * ld + c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum + c), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n4) {
  MTP_PDL_ENTRY();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    Vec4<__nv_bfloat16>::store(out + i * 4, *reinterpret_cast<const float4*>(in + i * 4));
}

// out_f32 += in_bf16  (feature-map gradient joining the residual-stream gradient)
__global__ void __launch_bounds__(256) add_bf16_into_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n4) {
  MTP_PDL_ENTRY();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = Vec4<__nv_bfloat16>::load(in + i * 4);
    float4 o = *reinterpret_cast<float4*>(out + i * 4);
    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    *reinterpret_cast<float4*>(out + i * 4) = o;
  }
}

template <typename TIn, bool GELU>
static int ln_fwd_dispatch(const void* x, const float* g, const float* b, void* y, float* mean, float* rstd, int rows, int C,
                           float eps, cudaStream_t st) {
  const dim3 grid(ceil_div(rows, ROW_WARPS)), block(ROW_WARPS * 32);
  const TIn* xp = reinterpret_cast<const TIn*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  switch (C / 128) {
#define LN_CASE(NV) case NV: (void)launch_k(ln_fwd_kernel<TIn, NV, GELU>, grid, block, 0, st, xp, g, b, yp, mean, rstd, rows, eps); break;
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(6) LN_CASE(8)
#undef LN_CASE
    default: return set_error(MTP_ERR_INVALID, "layernorm: unsupported C=%d", C);
  }
  return check_launch("ln_fwd_kernel");
}

template <typename TIn, typename TDx, bool GELU, bool CAST>
static int ln_bwd_dispatch(const void* dy, const void* x, const float* mean, const float* rstd, const float* g, const float* b,
                           const float* dres, void* dx, float* dgamma, float* dbeta, const float* row_scale, int rows_per_group,
                           void* g16, float* colsum16, const float* pool_add, const RvsaGeom& pg, int rows, int C, cudaStream_t st) {
  // one wave: every SM gets one CTA whose warps take one row each (rows <= 12 * SMs); beyond that CTAs loop over rows
  const int per_sm = ceil_div(rows, num_sms());
  const int warps = std::max(4, std::min(per_sm, LN_BWD_MAX_WARPS));
  const int rows_per_cta = std::max(per_sm, warps);
  const dim3 grid(ceil_div(rows, rows_per_cta)), block(warps * 32);
  const size_t smem = (size_t)(CAST ? 3 : 2) * warps * C * sizeof(float);
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
  const TIn* xp = reinterpret_cast<const TIn*>(x);
  TDx* dxp = reinterpret_cast<TDx*>(dx);
  __nv_bfloat16* g16p = reinterpret_cast<__nv_bfloat16*>(g16);
  cudaError_t e = cudaSuccess;
  switch (C / 128) {
#define LN_CASE(NV)                                                                                                        \
  case NV: {                                                                                                               \
    auto kern = ln_bwd_kernel<TIn, TDx, NV, GELU, CAST>;                                                                   \
    static bool attr = false;                                                                                              \
    if (!attr && smem > 48 * 1024) {                                                                                       \
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * LN_BWD_MAX_WARPS * NV * 128 * 4);    \
      if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "ln_bwd smem attr: %s", cudaGetErrorString(e));                 \
      attr = true;                                                                                                         \
    }                                                                                                                      \
    e = launch_k(kern, grid, block, smem, st, dyp, xp, mean, rstd, g, b, dres, dxp, dgamma, dbeta, row_scale,              \
                 rows_per_group, g16p, colsum16, pool_add, pg, rows, rows_per_cta);                                        \
  } break;
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(6) LN_CASE(8)
#undef LN_CASE
    default: return set_error(MTP_ERR_INVALID, "layernorm bwd: unsupported C=%d", C);
  }
  if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "ln_bwd_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("ln_bwd_kernel");
}

}  // namespace mtp

using namespace mtp;

extern "C" int mtp_layernorm_fwd(const void* x, int x_is_bf16, const float* gamma, const float* beta, void* y_bf16, float* mean,
                                 float* rstd, int rows, int C, float eps, int fuse_gelu, mtp_stream_t stream) {
  MTP_REQUIRE(x && gamma && beta && y_bf16, "mtp_layernorm_fwd: null pointer");
  MTP_REQUIRE(rows > 0 && C % 128 == 0 && C <= 1024, "mtp_layernorm_fwd: rows=%d C=%d unsupported (C%%128==0, C<=1024)", rows, C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x_is_bf16) {
    return fuse_gelu ? ln_fwd_dispatch<__nv_bfloat16, true>(x, gamma, beta, y_bf16, mean, rstd, rows, C, eps, st)
                     : ln_fwd_dispatch<__nv_bfloat16, false>(x, gamma, beta, y_bf16, mean, rstd, rows, C, eps, st);
  }
  return fuse_gelu ? ln_fwd_dispatch<float, true>(x, gamma, beta, y_bf16, mean, rstd, rows, C, eps, st)
                   : ln_fwd_dispatch<float, false>(x, gamma, beta, y_bf16, mean, rstd, rows, C, eps, st);
}

extern "C" int mtp_layernorm_bwd(const void* dy_bf16, const void* x, int x_is_bf16, const float* mean, const float* rstd,
                                 const float* gamma, const float* beta, const float* dres_f32, void* dx, int dx_is_bf16,
                                 float* dgamma, float* dbeta, const float* cast_row_scale, int cast_rows_per_group,
                                 void* cast_out_bf16, float* cast_colsum, const float* pool_add, int pool_h, int pool_w, int rows, int C,
                                 int fused_gelu, mtp_stream_t stream) {
  MTP_REQUIRE(dy_bf16 && x && mean && rstd && gamma && dx && dgamma && dbeta, "mtp_layernorm_bwd: null pointer");
  MTP_REQUIRE(rows > 0 && C % 128 == 0 && C <= 1024, "mtp_layernorm_bwd: rows=%d C=%d unsupported", rows, C);
  MTP_REQUIRE(!fused_gelu || beta, "mtp_layernorm_bwd: GELU variant needs beta");
  MTP_REQUIRE(!cast_row_scale || cast_rows_per_group > 0, "mtp_layernorm_bwd: cast_rows_per_group");
  MTP_REQUIRE(cast_out_bf16 || !cast_colsum, "mtp_layernorm_bwd: cast_colsum needs cast_out_bf16");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  RvsaGeom pg = make_rvsa_geom(1, WS, WS, C, 1);
  if (pool_add != nullptr) {
    MTP_REQUIRE(pool_h >= WS && pool_w >= WS && rows % (pool_h * pool_w) == 0, "mtp_layernorm_bwd: pool_add needs rows = B * pool_h * pool_w");
    pg = make_rvsa_geom(rows / (pool_h * pool_w), pool_h, pool_w, C, 1);
  }
  if (!x_is_bf16 && !dx_is_bf16 && !fused_gelu) {
    if (cast_out_bf16)
      return ln_bwd_dispatch<float, float, false, true>(dy_bf16, x, mean, rstd, gamma, beta, dres_f32, dx, dgamma, dbeta, cast_row_scale,
                                                        cast_rows_per_group, cast_out_bf16, cast_colsum, pool_add, pg, rows, C, st);
    return ln_bwd_dispatch<float, float, false, false>(dy_bf16, x, mean, rstd, gamma, beta, dres_f32, dx, dgamma, dbeta, nullptr, 0,
                                                       nullptr, nullptr, pool_add, pg, rows, C, st);
  }
  MTP_REQUIRE(!cast_out_bf16 && !pool_add, "mtp_layernorm_bwd: the fused cast / pool_add are only available for the (f32 x, f32 dx, no gelu) variant");
  if (x_is_bf16 && dx_is_bf16 && fused_gelu)
    return ln_bwd_dispatch<__nv_bfloat16, __nv_bfloat16, true, false>(dy_bf16, x, mean, rstd, gamma, beta, dres_f32, dx, dgamma, dbeta,
                                                                      nullptr, 0, nullptr, nullptr, nullptr, pg, rows, C, st);
  return set_error(MTP_ERR_INVALID, "mtp_layernorm_bwd: unsupported variant (x_bf16=%d dx_bf16=%d gelu=%d)", x_is_bf16, dx_is_bf16, fused_gelu);
}

extern "C" int mtp_scale_cast_bf16(const float* in, const float* row_scale, int rows_per_group, void* out_bf16, float* colsum,
                                   int rows, int C, mtp_stream_t stream) {
  MTP_REQUIRE(in && out_bf16, "mtp_scale_cast_bf16: null pointer");
  MTP_REQUIRE(rows > 0 && C > 0 && C % 4 == 0, "mtp_scale_cast_bf16: rows=%d C=%d", rows, C);
  MTP_REQUIRE(!row_scale || rows_per_group > 0, "mtp_scale_cast_bf16: rows_per_group");
  const int gx = ceil_div(C, 1024);
  const int gy = max(1, min(ceil_div(rows, 8), 4 * num_sms() / gx));
  const int rpc = ceil_div(rows, gy);
  (void)launch_k(scale_cast_kernel, dim3(gx, ceil_div(rows, rpc)), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      in, row_scale, rows_per_group, reinterpret_cast<__nv_bfloat16*>(out_bf16), colsum, rows, C, rpc);
  return check_launch("scale_cast_kernel");
}

extern "C" int mtp_colsum_bf16(const void* in_bf16, int ld, float* colsum, int rows, int C, mtp_stream_t stream) {
  MTP_REQUIRE(in_bf16 && colsum, "mtp_colsum_bf16: null pointer");
  MTP_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && ld % 4 == 0, "mtp_colsum_bf16: rows=%d C=%d ld=%d", rows, C, ld);
  const int gx = ceil_div(C, 1024);
  const int gy = max(1, min(ceil_div(rows, 8), 4 * num_sms() / gx));
  const int rpc = ceil_div(rows, gy);
  (void)launch_k(colsum_bf16_kernel, dim3(gx, ceil_div(rows, rpc)), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(in_bf16), colsum, rows, C, ld, rpc);
  return check_launch("colsum_bf16_kernel");
}

extern "C" int mtp_cast_f32_bf16(const float* in, void* out_bf16, size_t n, mtp_stream_t stream) {
  MTP_REQUIRE(in && out_bf16 && n % 4 == 0, "mtp_cast_f32_bf16: bad args");
  if (n == 0) return MTP_OK;
  const size_t n4 = n / 4;
  const int grid = (int)std::min<size_t>((n4 + 255) / 256, (size_t)num_sms() * 8);
  (void)launch_k(cast_f32_bf16_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), in, reinterpret_cast<__nv_bfloat16*>(out_bf16), n4);
  return check_launch("cast_f32_bf16_kernel");
}

namespace mtp {
__global__ void __launch_bounds__(256) add_f32_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n4) {
  MTP_PDL_ENTRY();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(in + i * 4);
    float4 o = *reinterpret_cast<float4*>(out + i * 4);
    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    *reinterpret_cast<float4*>(out + i * 4) = o;
  }
}
}  // namespace mtp

/* out_f32 += in_f32 (pyramid-tap cotangents computed ahead of the block loop joining the residual-stream gradient) */
extern "C" int mtp_add_f32(const float* in, float* out, size_t n, mtp_stream_t stream) {
  MTP_REQUIRE(in && out && n % 4 == 0, "mtp_add_f32: bad args");
  if (n == 0) return MTP_OK;
  const size_t n4 = n / 4;
  const int grid = (int)std::min<size_t>((n4 + 255) / 256, (size_t)num_sms() * 8);
  (void)launch_k(mtp::add_f32_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), in, out, n4);
  return check_launch("add_f32_kernel");
}

extern "C" int mtp_add_bf16_into_f32(const void* in_bf16, float* out, size_t n, mtp_stream_t stream) {
  MTP_REQUIRE(in_bf16 && out && n % 4 == 0, "mtp_add_bf16_into_f32: bad args");
  if (n == 0) return MTP_OK;
  const size_t n4 = n / 4;
  const int grid = (int)std::min<size_t>((n4 + 255) / 256, (size_t)num_sms() * 8);
  (void)launch_k(add_bf16_into_f32_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(in_bf16), out, n4);
  return check_launch("add_bf16_into_f32_kernel");
}
