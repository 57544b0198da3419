// fp32-class forward mode ("fp32x3"): row kernels that produce / consume values stored as TWO bf16 words, hi = bf16(v) and
// lo = bf16(v - hi), laid out side by side ([rows, 2K]: hi in columns [0, K), lo in [K, 2K)).  The tcgen05 GEMM multiplies such
// operands in three passes (hi*hi + hi*lo + lo*hi, fp32 accumulate: mtp_epilogue.hilo), the attention runs in fp32 (attn_window.cu /
// attn_full.cu, HILO instantiations), and the kernels here cover the rest of the forward: weight / activation splitting, the patch
// gather, LayerNorm (+GELU) and the token -> NCHW scatter.  Purpose: the north-star accuracy target (forward within 1e-3 of the fp32
// reference at ViT-L depth) and a full-depth logic check that bf16 rounding noise cannot mask.           [V]:787-817
#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace mtp {

__device__ __forceinline__ void hilo4(const float4 v, uint2& hi, uint2& lo) {
  hi.x = pack_bf16x2(v.x, v.y);
  hi.y = pack_bf16x2(v.z, v.w);
  const float2 a = unpack_bf16x2(hi.x), b = unpack_bf16x2(hi.y);
  lo.x = pack_bf16x2(v.x - a.x, v.y - a.y);
  lo.y = pack_bf16x2(v.z - b.x, v.w - b.y);
}
__device__ __forceinline__ float4 load_hilo4(const __nv_bfloat16* p, int lo_off) {
  const uint2 h = *reinterpret_cast<const uint2*>(p), l = *reinterpret_cast<const uint2*>(p + lo_off);
  const float2 a = unpack_bf16x2(h.x), b = unpack_bf16x2(h.y), c = unpack_bf16x2(l.x), d = unpack_bf16x2(l.y);
  return make_float4(a.x + c.x, a.y + c.y, b.x + d.x, b.y + d.y);
}

// in fp32 [rows, K] (row pitch ld_in) -> out [rows, 2K]
__global__ void __launch_bounds__(256)
split_hilo_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t rows, int K, int ld_in) {
  MTP_PDL_ENTRY();
  const int k4 = K / 4;
  const size_t total = rows * (size_t)k4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / k4;
    const int c = (int)(i % k4) * 4;
    uint2 hi, lo;
    hilo4(*reinterpret_cast<const float4*>(in + r * ld_in + c), hi, lo);
    *reinterpret_cast<uint2*>(out + r * 2 * K + c) = hi;
    *reinterpret_cast<uint2*>(out + r * 2 * K + K + c) = lo;
  }
}

// image fp32 (B, cin, H, W) -> patch rows [T, 2 * cin*256] (hi | lo)        [V]:536-539
__global__ void __launch_bounds__(256)
patchify_hilo_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int cin, int H, int W, int gh, int gw) {
  MTP_PDL_ENTRY();
  const int K0 = cin * 256;
  const int chunks_per_tok = cin * 16 * 4;
  const size_t total = (size_t)B * gh * gw * chunks_per_tok;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks_per_tok);
    const size_t tok = i / chunks_per_tok;
    const int kx4 = ch & 3, ky = (ch >> 2) & 15, c = ch >> 6;
    const int px = (int)(tok % gw), py = (int)((tok / gw) % gh), b = (int)(tok / ((size_t)gw * gh));
    const float4 v = *reinterpret_cast<const float4*>(img + (((size_t)b * cin + c) * H + py * 16 + ky) * W + px * 16 + kx4 * 4);
    uint2 hi, lo;
    hilo4(v, hi, lo);
    __nv_bfloat16* o = out + tok * (size_t)(2 * K0) + c * 256 + ky * 16 + kx4 * 4;
    *reinterpret_cast<uint2*>(o) = hi;
    *reinterpret_cast<uint2*>(o + K0) = lo;
  }
}

// y (hi | lo, [rows, 2C]) = LN(x) * gamma + beta, optionally GELU'd; x fp32 [rows, C] or hi | lo [rows, 2C].  One warp per row.
template <bool XHILO, int NV, bool GELU>
__global__ void __launch_bounds__(256)
ln_fwd_hilo_kernel(const void* __restrict__ x_, const float* __restrict__ gamma, const float* __restrict__ beta,
                   __nv_bfloat16* __restrict__ y, int rows, float eps, int x_ld, int x_lo, int x_sub) {
  MTP_PDL_ENTRY();
  constexpr int C = NV * 128;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    // hi | lo input: logical row r lives in physical row r / x_sub at column (r % x_sub) * C, lo words x_lo further on (the [4T, C] view
    // of a ConvTranspose2d GEMM output [T, 4C | 4C] has x_sub = 4, x_ld = 8C, x_lo = 4C; a plain [rows, C | C] matrix 1, 2C, C)
    if (XHILO) v[i] = load_hilo4(reinterpret_cast<const __nv_bfloat16*>(x_) + (size_t)(row / x_sub) * x_ld + (row % x_sub) * C + c, x_lo);
    else v[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x_) + (size_t)row * C + c);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mu = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
    q += a * a + b * b + c * c + d * d;
  }
  const float rs = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + eps);
  __nv_bfloat16* yr = y + (size_t)row * 2 * C;
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
    uint2 hi, lo;
    hilo4(o, hi, lo);
    *reinterpret_cast<uint2*>(yr + c) = hi;
    *reinterpret_cast<uint2*>(yr + C + c) = lo;
  }
}

template <bool XHILO, bool GELU>
static int ln_hilo_dispatch(const void* x, const float* gamma, const float* beta, void* y, int rows, int C, float eps, int x_ld, int x_lo,
                            int x_sub, cudaStream_t st) {
  const int grid = ceil_div(rows, 8);
  __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(y);
#define MTP_LNH(NV_) case NV_: (void)launch_k(ln_fwd_hilo_kernel<XHILO, NV_, GELU>, grid, 256, 0, st, x, gamma, beta, yy, rows, eps, x_ld, x_lo, x_sub); break;
  switch (C / 128) {
    MTP_LNH(1) MTP_LNH(2) MTP_LNH(3) MTP_LNH(4) MTP_LNH(5) MTP_LNH(6) MTP_LNH(7) MTP_LNH(8)
    default: return set_error(MTP_ERR_INVALID, "mtp_layernorm_fwd_hilo: C=%d unsupported", C);
  }
#undef MTP_LNH
  return check_launch("ln_fwd_hilo_kernel");
}

}  // namespace mtp

using namespace mtp;

extern "C" int mtp_split_hilo(const float* in, int ld_in, void* out_hilo, size_t rows, int K, mtp_stream_t stream) {
  MTP_REQUIRE(in && out_hilo && rows > 0 && K > 0 && K % 8 == 0 && ld_in >= K && ld_in % 4 == 0, "mtp_split_hilo: bad args (K %% 8 == 0)");
  const size_t total = rows * (size_t)(K / 4);
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 8);
  (void)launch_k(split_hilo_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), in, reinterpret_cast<__nv_bfloat16*>(out_hilo), rows, K, ld_in);
  return check_launch("split_hilo_kernel");
}

extern "C" int mtp_patchify_hilo(const float* img, void* out_hilo, int B, int cin, int H, int W, mtp_stream_t stream) {
  MTP_REQUIRE(img && out_hilo, "mtp_patchify_hilo: null pointer");
  MTP_REQUIRE(B > 0 && cin > 0 && H >= 16 && W >= 16 && W % 4 == 0, "mtp_patchify_hilo: B=%d cin=%d H=%d W=%d unsupported", B, cin, H, W);
  const int gh = H / 16, gw = W / 16;
  const size_t total = (size_t)B * gh * gw * cin * 64;
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 8);
  (void)launch_k(patchify_hilo_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), img, reinterpret_cast<__nv_bfloat16*>(out_hilo), B, cin, H,
                 W, gh, gw);
  return check_launch("patchify_hilo_kernel");
}

extern "C" int mtp_layernorm_fwd_hilo(const void* x, int x_is_hilo, int x_ld, int x_lo_offset, int x_sub, const float* gamma,
                                      const float* beta, void* y_hilo, int rows, int C, float eps, int fuse_gelu, mtp_stream_t stream) {
  MTP_REQUIRE(x && gamma && beta && y_hilo, "mtp_layernorm_fwd_hilo: null pointer");
  MTP_REQUIRE(rows > 0 && C % 128 == 0 && C <= 1024, "mtp_layernorm_fwd_hilo: rows=%d C=%d unsupported", rows, C);
  MTP_REQUIRE(!x_is_hilo || (x_sub >= 1 && x_lo_offset >= x_sub * C && x_ld >= x_lo_offset + x_sub * C && rows % x_sub == 0),
              "mtp_layernorm_fwd_hilo: bad hi|lo input geometry (ld=%d lo=%d sub=%d)", x_ld, x_lo_offset, x_sub);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x_is_hilo) return fuse_gelu ? ln_hilo_dispatch<true, true>(x, gamma, beta, y_hilo, rows, C, eps, x_ld, x_lo_offset, x_sub, st)
                                  : ln_hilo_dispatch<true, false>(x, gamma, beta, y_hilo, rows, C, eps, x_ld, x_lo_offset, x_sub, st);
  return fuse_gelu ? ln_hilo_dispatch<false, true>(x, gamma, beta, y_hilo, rows, C, eps, 0, 0, 1, st)
                   : ln_hilo_dispatch<false, false>(x, gamma, beta, y_hilo, rows, C, eps, 0, 0, 1, st);
}
