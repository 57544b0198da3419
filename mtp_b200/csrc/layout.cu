// Layout kernels at the two ends of the backbone: image -> patch rows (the A operand of the patch-embed GEMM), and
// token-major feature matrices <-> the NCHW maps the decoders consume, including the pixel-shuffle of the k2/s2
// transposed convolutions (a ConvTranspose2d(k2,s2) is the GEMM [T,Cin] x [Cin,4Cout] whose output row (b,y,x) holds the
// four sub-pixels (dy,dx) side by side; nesting two of them gives level 2).        [V]:529-540, 640-654, 807-811
#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace mtp {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------------- patchify
// out[(b, py, px), c*256 + ky*16 + kx] = img[b, c, py*16+ky, px*16+kx]   (Conv2d weight (C, cin, 16, 16) flattens the same way)
template <typename TIn>
__global__ void __launch_bounds__(256)
patchify_kernel(const TIn* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int cin, int H, int W, int gh, int gw) {
  MTP_PDL_ENTRY();
  const int chunks_per_tok = cin * 16 * 4;                 // 4-element chunks per token row
  const size_t total = (size_t)B * gh * gw * chunks_per_tok;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks_per_tok);
    const size_t tok = i / chunks_per_tok;
    const int kx4 = ch & 3, ky = (ch >> 2) & 15, c = ch >> 6;
    const int px = (int)(tok % gw), py = (int)((tok / gw) % gh), b = (int)(tok / ((size_t)gw * gh));
    const TIn* src = img + (((size_t)b * cin + c) * H + py * 16 + ky) * W + px * 16 + kx4 * 4;
    uint2 u;
    u.x = pack_bf16x2(to_f32(src[0]), to_f32(src[1]));
    u.y = pack_bf16x2(to_f32(src[2]), to_f32(src[3]));
    *reinterpret_cast<uint2*>(out + tok * (size_t)(cin * 256) + c * 256 + ky * 16 + kx4 * 4) = u;
  }
}

// uint8 images straight from the loader, with MTP_DataPreprocessor's arithmetic folded in (Multi-Task_Pretrain/preprocessing.py:145-187
// -> mmengine ImgDataPreprocessor: optional BGR<->RGB channel flip, .float(), (x - mean[c]) / std[c]); the padding step is the identity
// because the backbone requires H, W == img_size.  out[(b,py,px), c*256 + ky*16 + kx] = bf16((img[b, src(c), y, x] - mean[c]) / std[c]),
// src(c) = cin-1-c when `flip`.  Layout CHW (what PackDetInputs hands to the preprocessor) or HWC (a decoded image as it lies in memory).
struct PreNorm { float mean[4], stdv[4]; };

template <bool HWC>
__global__ void __launch_bounds__(256)
patchify_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int cin, int H, int W, int gh, int gw, int flip,
                   const PreNorm pn) {
  MTP_PDL_ENTRY();
  const int chunks_per_tok = cin * 16 * 4;                 // 4-pixel chunks per token row
  const size_t total = (size_t)B * gh * gw * chunks_per_tok;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks_per_tok);
    const size_t tok = i / chunks_per_tok;
    const int kx4 = ch & 3, ky = (ch >> 2) & 15, c = ch >> 6;
    const int px = (int)(tok % gw), py = (int)((tok / gw) % gh), b = (int)(tok / ((size_t)gw * gh));
    const int cs = flip ? cin - 1 - c : c;
    const int y = py * 16 + ky, x = px * 16 + kx4 * 4;
    float v[4];
    if (HWC) {
      const uint8_t* src = img + (((size_t)b * H + y) * W + x) * cin + cs;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (float)src[k * cin];
    } else {
      const uchar4 u = *reinterpret_cast<const uchar4*>(img + (((size_t)b * cin + cs) * H + y) * W + x);
      v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
    }
    const float m = pn.mean[c], sd = pn.stdv[c];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __fdiv_rn(v[k] - m, sd);      // IEEE division: identical to the reference's fp32 (x - mean) / std
    uint2 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(out + tok * (size_t)(cin * 256) + c * 256 + ky * 16 + kx4 * 4) = o;
  }
}

// ---------------------------------------------------------------------------------------------------- tok <-> NCHW
struct MapGeom { int B, h, w, C, L; };      // base grid h x w, channels C, pixel-shuffle level L (0, 1, 2)

// (row, column base) of output pixel (b, Y, X) of the level-L map inside the token-major matrix
__device__ __forceinline__ void map_index(const MapGeom& g, int b, int Y, int X, size_t& row, int& colbase) {
  if (g.L == 0) { row = ((size_t)b * g.h + Y) * g.w + X; colbase = 0; }
  else if (g.L == 1) { row = ((size_t)b * g.h + (Y >> 1)) * g.w + (X >> 1); colbase = (((Y & 1) << 1) | (X & 1)) * g.C; }
  else {
    row = (((size_t)b * g.h + (Y >> 2)) * g.w + (X >> 2)) * 4 + ((((Y >> 1) & 1) << 1) | ((X >> 1) & 1));
    colbase = (((Y & 1) << 1) | (X & 1)) * g.C;
  }
}

// NCHW out[b, c, Y, X] = tok[row(b,Y,X), colbase + c];   32(X) x 32(c) tiles through shared memory
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
tok_to_nchw_kernel(const TIn* __restrict__ tok, int ld, TOut* __restrict__ out, const MapGeom g, int lo) {
  MTP_PDL_ENTRY();
  __shared__ float tile[32][33];
  const int Ho = g.h << g.L, Wo = g.w << g.L;
  const int x_tiles = ceil_div(Wo, 32);
  const int xt = blockIdx.x % x_tiles, Y = blockIdx.x / x_tiles;
  const int c0 = blockIdx.y * 32, b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (int i = ty; i < 32; i += 8) {                            // i = X within tile, tx = channel
    const int X = xt * 32 + i;
    if (X < Wo && c0 + tx < g.C) {
      size_t row; int cb;
      map_index(g, b, Y, X, row, cb);
      float v = to_f32(tok[row * ld + cb + c0 + tx]);
      if (lo > 0) v += to_f32(tok[row * ld + cb + c0 + tx + lo]);      // fp32-class mode: hi | lo word pairs (lo at +lo)
      tile[i][tx] = v;
    }
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {                            // i = channel within tile, tx = X
    const int X = xt * 32 + tx, c = c0 + i;
    if (X < Wo && c < g.C) out[(((size_t)b * g.C + c) * Ho + Y) * Wo + X] = from_f32<TOut>(tile[tx][i]);
  }
}

// tok[row(b,Y,X), colbase + c] (+)= in[b, c, Y, X]
template <typename TIn, typename TOut, bool ACC>
__global__ void __launch_bounds__(256)
nchw_to_tok_kernel(const TIn* __restrict__ in, TOut* __restrict__ tok, int ld, const MapGeom g) {
  MTP_PDL_ENTRY();
  __shared__ float tile[32][33];
  const int Ho = g.h << g.L, Wo = g.w << g.L;
  const int x_tiles = ceil_div(Wo, 32);
  const int xt = blockIdx.x % x_tiles, Y = blockIdx.x / x_tiles;
  const int c0 = blockIdx.y * 32, b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {                            // i = channel, tx = X
    const int X = xt * 32 + tx, c = c0 + i;
    if (X < Wo && c < g.C) tile[tx][i] = to_f32(in[(((size_t)b * g.C + c) * Ho + Y) * Wo + X]);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {                            // i = X, tx = channel
    const int X = xt * 32 + i;
    if (X < Wo && c0 + tx < g.C) {
      size_t row; int cb;
      map_index(g, b, Y, X, row, cb);
      TOut* dst = tok + row * ld + cb + c0 + tx;
      if (ACC) *dst = from_f32<TOut>(to_f32(*dst) + tile[i][tx]);
      else *dst = from_f32<TOut>(tile[i][tx]);
    }
  }
}

// bf16 -> bf16, 16-byte accesses on both sides (the 56 x 56 pyramid maps are 51 MB each way; the scalar 32 x 32 transposes above spend 22
// thread instructions per element and run at 1.8 TB/s).  CTA = (b, Y, 32 X, 64 channels), 256 threads: the token side moves 8 channels per
// thread (one X), the NCHW side 8 X per thread (one channel); the tile goes through shared memory as 32-bit words [64 c][33]: the token-side
// accesses are 2-way, the NCHW-side accesses conflict-free.  Needs Wo % 8 == 0 and C % 64 == 0.
__global__ void __launch_bounds__(256)
tok_to_nchw_vec_kernel(const __nv_bfloat16* __restrict__ tok, int ld, __nv_bfloat16* __restrict__ out, const MapGeom g) {
  MTP_PDL_ENTRY();
  __shared__ uint32_t tile[64][33];
  const int Ho = g.h << g.L, Wo = g.w << g.L;
  const int x_tiles = ceil_div(Wo, 32);
  const int xt = blockIdx.x % x_tiles, Y = blockIdx.x / x_tiles;
  const int c0 = blockIdx.y * 64, b = blockIdx.z;
  {
    const int xi = threadIdx.x >> 3, ch = threadIdx.x & 7;      // X within the tile, 8-channel chunk
    const int X = xt * 32 + xi;
    if (X < Wo) {
      size_t row; int cb;
      map_index(g, b, Y, X, row, cb);
      const uint4 u = *reinterpret_cast<const uint4*>(tok + row * ld + cb + c0 + ch * 8);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tile[ch * 8 + 2 * k][xi] = w4[k] & 0xffffu;
        tile[ch * 8 + 2 * k + 1][xi] = w4[k] >> 16;
      }
    }
  }
  __syncthreads();
  {
    const int c = threadIdx.x >> 2, xq = threadIdx.x & 3;        // channel within the tile, group of 8 X
    const int X0 = xt * 32 + xq * 8;
    if (X0 < Wo) {                                               // Wo % 8 == 0: the group is all in or all out
      uint32_t w4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) w4[k] = tile[c][xq * 8 + 2 * k] | (tile[c][xq * 8 + 2 * k + 1] << 16);
      *reinterpret_cast<uint4*>(out + (((size_t)b * g.C + c0 + c) * Ho + Y) * Wo + X0) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
  }
}

__global__ void __launch_bounds__(256)
nchw_to_tok_vec_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ tok, int ld, const MapGeom g) {
  MTP_PDL_ENTRY();
  __shared__ uint32_t tile[64][33];
  const int Ho = g.h << g.L, Wo = g.w << g.L;
  const int x_tiles = ceil_div(Wo, 32);
  const int xt = blockIdx.x % x_tiles, Y = blockIdx.x / x_tiles;
  const int c0 = blockIdx.y * 64, b = blockIdx.z;
  {
    const int c = threadIdx.x >> 2, xq = threadIdx.x & 3;
    const int X0 = xt * 32 + xq * 8;
    if (X0 < Wo) {
      const uint4 u = *reinterpret_cast<const uint4*>(in + (((size_t)b * g.C + c0 + c) * Ho + Y) * Wo + X0);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tile[c][xq * 8 + 2 * k] = w4[k] & 0xffffu;
        tile[c][xq * 8 + 2 * k + 1] = w4[k] >> 16;
      }
    }
  }
  __syncthreads();
  {
    const int xi = threadIdx.x >> 3, ch = threadIdx.x & 7;
    const int X = xt * 32 + xi;
    if (X < Wo) {
      size_t row; int cb;
      map_index(g, b, Y, X, row, cb);
      uint32_t w4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) w4[k] = tile[ch * 8 + 2 * k][xi] | (tile[ch * 8 + 2 * k + 1][xi] << 16);
      *reinterpret_cast<uint4*>(tok + row * ld + cb + c0 + ch * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------- 2x2 max pool (fpn4)
// token-major f32 [B, h, w, C] -> token-major f32 [B, h/2, w/2, C]                                   [V]:654
__global__ void __launch_bounds__(256)
maxpool2_tok_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int h, int w, int C) {
  MTP_PDL_ENTRY();
  const int ho = h / 2, wo = w / 2, c4n = C / 4;
  const size_t total = (size_t)B * ho * wo * c4n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const size_t p = i / c4n;
    const int xo = (int)(p % wo), yo = (int)((p / wo) % ho), b = (int)(p / ((size_t)wo * ho));
    const float* s = x + (((size_t)b * h + 2 * yo) * w + 2 * xo) * C + c;
    const float4 a = *reinterpret_cast<const float4*>(s), bq = *reinterpret_cast<const float4*>(s + C);
    const float4 cq = *reinterpret_cast<const float4*>(s + (size_t)w * C), d = *reinterpret_cast<const float4*>(s + (size_t)w * C + C);
    float4 m;
    m.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(cq.x, d.x));
    m.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(cq.y, d.y));
    m.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(cq.z, d.z));
    m.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(cq.w, d.w));
    *reinterpret_cast<float4*>(y + p * C + c) = m;
  }
}

// dx[argmax position] += dy   (first maximum in scan order wins, as in ATen's max_pool2d backward)
__global__ void __launch_bounds__(256)
maxpool2_tok_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int B, int h, int w, int C) {
  MTP_PDL_ENTRY();
  const int ho = h / 2, wo = w / 2;
  const size_t total = (size_t)B * ho * wo * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t p = i / C;
    const int xo = (int)(p % wo), yo = (int)((p / wo) % ho), b = (int)(p / ((size_t)wo * ho));
    const size_t base = (((size_t)b * h + 2 * yo) * w + 2 * xo) * C + c;
    const size_t offs[4] = {0, (size_t)C, (size_t)w * C, (size_t)w * C + C};
    int best = 0;
    float bv = x[base];
#pragma unroll
    for (int t = 1; t < 4; ++t) {
      const float v = x[base + offs[t]];
      if (v > bv) { bv = v; best = t; }
    }
    dx[base + offs[best]] += dy[i];
  }
}

static inline int grid_for(size_t n, int block) { return (int)std::min<size_t>((n + block - 1) / block, (size_t)num_sms() * 16); }

// Stand-in objective of the pretraining step when no decoder heads are attached (trainer.synthetic_heads, bench.py):
// loss += 0.5 * mean(f^2), d loss / d f = f / n, one pass over the bf16 feature map.
__global__ void __launch_bounds__(256)
sqloss_kernel(const __nv_bfloat16* __restrict__ f, __nv_bfloat16* __restrict__ g, float* __restrict__ loss, size_t n8, float inv_n) {
  MTP_PDL_ENTRY();
  __shared__ float red[8];
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(f + i * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 v = unpack_bf16x2(w[t]);
      s += v.x * v.x + v.y * v.y;
      o[t] = pack_bf16x2(v.x * inv_n, v.y * inv_n);
    }
    *reinterpret_cast<uint4*>(g + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = red[threadIdx.x];
    t += __shfl_xor_sync(0xffu, t, 4);
    t += __shfl_xor_sync(0xffu, t, 2);
    t += __shfl_xor_sync(0xffu, t, 1);
    if (threadIdx.x == 0) atomicAdd(loss, 0.5f * inv_n * t);
  }
}

}  // namespace mtp

using namespace mtp;

extern "C" int mtp_patchify(const void* img, int img_is_bf16, void* out_bf16, int B, int cin, int H, int W, mtp_stream_t stream) {
  MTP_REQUIRE(img && out_bf16, "mtp_patchify: null pointer");
  MTP_REQUIRE(B > 0 && cin > 0 && H >= 16 && W >= 16 && W % 4 == 0, "mtp_patchify: B=%d cin=%d H=%d W=%d unsupported", B, cin, H, W);
  const int gh = H / 16, gw = W / 16;
  const size_t total = (size_t)B * gh * gw * cin * 64;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (img_is_bf16)
    (void)launch_k(patchify_kernel<__nv_bfloat16>, grid_for(total, 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(img),
                                                                         reinterpret_cast<__nv_bfloat16*>(out_bf16), B, cin, H, W, gh, gw);
  else
    (void)launch_k(patchify_kernel<float>, grid_for(total, 256), 256, 0, st, reinterpret_cast<const float*>(img),
                                                                 reinterpret_cast<__nv_bfloat16*>(out_bf16), B, cin, H, W, gh, gw);
  return check_launch("patchify_kernel");
}

extern "C" int mtp_patchify_u8(const void* img_u8, int hwc, int flip_channels, const float* mean, const float* stdv, void* out_bf16, int B,
                               int cin, int H, int W, mtp_stream_t stream) {
  MTP_REQUIRE(img_u8 && out_bf16 && mean && stdv, "mtp_patchify_u8: null pointer");
  MTP_REQUIRE(B > 0 && cin > 0 && cin <= 4 && H >= 16 && W >= 16 && W % 4 == 0, "mtp_patchify_u8: B=%d cin=%d H=%d W=%d unsupported", B, cin, H, W);
  MTP_REQUIRE(hwc || ((uintptr_t)img_u8 & 3) == 0, "mtp_patchify_u8: CHW images must be 4-byte aligned");
  PreNorm pn;
  for (int c = 0; c < 4; ++c) {
    pn.mean[c] = c < cin ? mean[c] : 0.f;                 // host arrays, passed by value
    pn.stdv[c] = c < cin ? stdv[c] : 1.f;
    MTP_REQUIRE(pn.stdv[c] != 0.f, "mtp_patchify_u8: std[%d] == 0", c);
  }
  const int gh = H / 16, gw = W / 16;
  const size_t total = (size_t)B * gh * gw * cin * 64;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (hwc)
    (void)launch_k(patchify_u8_kernel<true>, grid_for(total, 256), 256, 0, st, reinterpret_cast<const uint8_t*>(img_u8),
                   reinterpret_cast<__nv_bfloat16*>(out_bf16), B, cin, H, W, gh, gw, flip_channels, pn);
  else
    (void)launch_k(patchify_u8_kernel<false>, grid_for(total, 256), 256, 0, st, reinterpret_cast<const uint8_t*>(img_u8),
                   reinterpret_cast<__nv_bfloat16*>(out_bf16), B, cin, H, W, gh, gw, flip_channels, pn);
  return check_launch("patchify_u8_kernel");
}

extern "C" int mtp_tok_to_nchw(const void* tok, int tok_is_bf16, int ld, void* out, int out_is_bf16, int B, int h, int w, int C,
                               int level, mtp_stream_t stream) {
  MTP_REQUIRE(tok && out, "mtp_tok_to_nchw: null pointer");
  MTP_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && level >= 0 && level <= 2 && ld >= (level ? 4 * C : C), "mtp_tok_to_nchw: bad geometry");
  const MapGeom g{B, h, w, C, level};
  const dim3 grid(ceil_div(w << level, 32) * (h << level), ceil_div(C, 32), B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define T2N(TI, TO) (void)launch_k(tok_to_nchw_kernel<TI, TO>, grid, 256, 0, st, reinterpret_cast<const TI*>(tok), ld, reinterpret_cast<TO*>(out), g, 0)
  if (tok_is_bf16 && out_is_bf16 && (w << level) % 8 == 0 && C % 64 == 0 && ld % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(tok) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const dim3 gridv(ceil_div(w << level, 32) * (h << level), C / 64, B);
    (void)launch_k(tok_to_nchw_vec_kernel, gridv, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(tok), ld, reinterpret_cast<__nv_bfloat16*>(out), g);
    return check_launch("tok_to_nchw_vec_kernel");
  }
  if (tok_is_bf16 && out_is_bf16) T2N(__nv_bfloat16, __nv_bfloat16);
  else if (tok_is_bf16) T2N(__nv_bfloat16, float);
  else if (out_is_bf16) T2N(float, __nv_bfloat16);
  else T2N(float, float);
#undef T2N
  return check_launch("tok_to_nchw_kernel");
}

extern "C" int mtp_tok_to_nchw_hilo(const void* tok_hilo, int ld, int lo_offset, float* out, int B, int h, int w, int C, int level,
                                    mtp_stream_t stream) {
  MTP_REQUIRE(tok_hilo && out, "mtp_tok_to_nchw_hilo: null pointer");
  MTP_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && level >= 0 && level <= 2 && lo_offset > 0 && ld >= lo_offset + (level ? 4 * C : C),
              "mtp_tok_to_nchw_hilo: bad geometry");
  const MapGeom g{B, h, w, C, level};
  const dim3 grid(ceil_div(w << level, 32) * (h << level), ceil_div(C, 32), B);
  (void)launch_k(tok_to_nchw_kernel<__nv_bfloat16, float>, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream),
                 reinterpret_cast<const __nv_bfloat16*>(tok_hilo), ld, out, g, lo_offset);
  return check_launch("tok_to_nchw_kernel");
}

extern "C" int mtp_nchw_to_tok(const void* in, int in_is_bf16, void* tok, int tok_is_bf16, int ld, int accumulate, int B, int h, int w,
                               int C, int level, mtp_stream_t stream) {
  MTP_REQUIRE(in && tok, "mtp_nchw_to_tok: null pointer");
  MTP_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && level >= 0 && level <= 2 && ld >= (level ? 4 * C : C), "mtp_nchw_to_tok: bad geometry");
  const MapGeom g{B, h, w, C, level};
  const dim3 grid(ceil_div(w << level, 32) * (h << level), ceil_div(C, 32), B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define N2T(TI, TO, A) (void)launch_k(nchw_to_tok_kernel<TI, TO, A>, grid, 256, 0, st, reinterpret_cast<const TI*>(in), reinterpret_cast<TO*>(tok), ld, g)
  if (accumulate) {
    MTP_REQUIRE(!tok_is_bf16, "mtp_nchw_to_tok: accumulate needs an f32 destination");
    if (in_is_bf16) N2T(__nv_bfloat16, float, true); else N2T(float, float, true);
  } else if (tok_is_bf16) {
    if (in_is_bf16 && (w << level) % 8 == 0 && C % 64 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(tok) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
      const dim3 gridv(ceil_div(w << level, 32) * (h << level), C / 64, B);
      (void)launch_k(nchw_to_tok_vec_kernel, gridv, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(tok), ld, g);
      return check_launch("nchw_to_tok_vec_kernel");
    }
    if (in_is_bf16) N2T(__nv_bfloat16, __nv_bfloat16, false); else N2T(float, __nv_bfloat16, false);
  } else {
    if (in_is_bf16) N2T(__nv_bfloat16, float, false); else N2T(float, float, false);
  }
#undef N2T
  return check_launch("nchw_to_tok_kernel");
}

extern "C" int mtp_maxpool2_tok_fwd(const float* x, float* y, int B, int h, int w, int C, mtp_stream_t stream) {
  MTP_REQUIRE(x && y && B > 0 && h >= 2 && w >= 2 && C % 4 == 0, "mtp_maxpool2_tok_fwd: bad args");
  const size_t total = (size_t)B * (h / 2) * (w / 2) * (C / 4);
  (void)launch_k(maxpool2_tok_fwd_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), x, y, B, h, w, C);
  return check_launch("maxpool2_tok_fwd_kernel");
}

extern "C" int mtp_maxpool2_tok_bwd(const float* x, const float* dy, float* dx, int B, int h, int w, int C, mtp_stream_t stream) {
  MTP_REQUIRE(x && dy && dx && B > 0 && h >= 2 && w >= 2, "mtp_maxpool2_tok_bwd: bad args");
  const size_t total = (size_t)B * (h / 2) * (w / 2) * C;
  (void)launch_k(maxpool2_tok_bwd_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), x, dy, dx, B, h, w, C);
  return check_launch("maxpool2_tok_bwd_kernel");
}

extern "C" int mtp_sqloss_fwd_bwd_w(const void* feat_bf16, void* grad_bf16, float* loss, size_t n, float weight, mtp_stream_t stream) {
  MTP_REQUIRE(feat_bf16 && grad_bf16 && loss && n > 0 && n % 8 == 0, "mtp_sqloss_fwd_bwd: bad args (n must be a positive multiple of 8)");
  MTP_REQUIRE((((uintptr_t)feat_bf16 | (uintptr_t)grad_bf16) & 15) == 0, "mtp_sqloss_fwd_bwd: pointers must be 16-byte aligned");
  const size_t n8 = n / 8;
  const int grid = (int)std::min<size_t>((n8 + 255) / 256, (size_t)num_sms() * 8);
  (void)launch_k(sqloss_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(feat_bf16),
                 reinterpret_cast<__nv_bfloat16*>(grad_bf16), loss, n8, weight / (float)n);
  return check_launch("sqloss_kernel");
}

extern "C" int mtp_sqloss_fwd_bwd(const void* feat_bf16, void* grad_bf16, float* loss, size_t n, mtp_stream_t stream) {
  return mtp_sqloss_fwd_bwd_w(feat_bf16, grad_bf16, loss, n, 1.0f, stream);
}
