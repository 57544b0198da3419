// Fused optimizer tail of the pretrain step over FLAT parameter / gradient buffers (one launch each):
//   mtp_sumsq_f32  : global gradient sum of squares (for clip_grad_norm_, main_pretrain.py:786)
//   mtp_adamw_step : AdamW (torch.optim.AdamW semantics, main_pretrain.py:787) with per-parameter-group lr scale and
//                    weight decay (layer-decay constructor groups, mmcv_custom/layer_decay_optimizer_constructor_vit.py),
//                    gradient averaging over ranks, norm clipping, cosine schedule (main_pretrain.py:832), and the bf16
//                    mirror of the weights (the next step's GEMM operands) written in the same pass.
#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace mtp {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, size_t n4, float* __restrict__ out) {
  MTP_PDL_ENTRY();
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

__global__ void __launch_bounds__(256) sumsq_bf16_kernel(const __nv_bfloat16* __restrict__ x, size_t n8, float* __restrict__ out) {
  MTP_PDL_ENTRY();
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + i * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) { const float2 f = unpack_bf16x2(w[t]); s += f.x * f.x + f.y * f.y; }
  }
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

struct AdamWArgs {
  float lr0, eta_min, beta1, beta2, eps, max_norm, grad_scale;
  int t_max;
};

// state[0] = step counter (as float), state[1] = sum of squares of the (summed over ranks) gradients
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ p16, const uint8_t* __restrict__ chunk_group, const float* __restrict__ group_lr,
             const float* __restrict__ group_wd, const float* __restrict__ state, size_t n4, AdamWArgs a,
             const __nv_bfloat16* __restrict__ g16, size_t g16_from4) {
  MTP_PDL_ENTRY();
  const float step = state[0];
  float lr = a.lr0;
  if (a.t_max > 0) lr = a.eta_min + (a.lr0 - a.eta_min) * 0.5f * (1.0f + cospif(fminf(step - 1.0f, (float)a.t_max) / (float)a.t_max));
  const float bc1 = 1.0f - powf(a.beta1, step), bc2 = 1.0f - powf(a.beta2, step);
  float gs = a.grad_scale;
  if (a.max_norm > 0.f) {
    const float total = sqrtf(state[1]) * a.grad_scale;
    gs *= fminf(1.0f, a.max_norm / (total + 1e-6f));
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int grp = chunk_group[i >> 4];                 // 64-element chunks = 16 float4
    const float lr_g = lr * __ldg(group_lr + grp), wd = __ldg(group_wd + grp);
    float4 pv = *reinterpret_cast<float4*>(p + i * 4);
    float4 gv;
    if (g16 != nullptr && i >= g16_from4) {      // gradients of the GEMM-weight region were all-reduced as bf16 (data-parallel trainer)
      const uint2 u = *reinterpret_cast<const uint2*>(g16 + i * 4);
      const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y);
      gv = make_float4(a0.x, a0.y, a1.x, a1.y);
    } else {
      gv = *reinterpret_cast<const float4*>(g + i * 4);
    }
    float4 mv = *reinterpret_cast<float4*>(m + i * 4), vv = *reinterpret_cast<float4*>(v + i * 4);
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gp[k] * gs;
      pp[k] *= 1.0f - lr_g * wd;
      mp[k] = a.beta1 * mp[k] + (1.0f - a.beta1) * gk;
      vp[k] = a.beta2 * vp[k] + (1.0f - a.beta2) * gk * gk;
      pp[k] -= lr_g * (mp[k] / bc1) / (sqrtf(vp[k] / bc2) + a.eps);
    }
    *reinterpret_cast<float4*>(p + i * 4) = pv;
    *reinterpret_cast<float4*>(m + i * 4) = mv;
    *reinterpret_cast<float4*>(v + i * 4) = vv;
    if (p16) {
      uint2 u;
      u.x = pack_bf16x2(pv.x, pv.y);
      u.y = pack_bf16x2(pv.z, pv.w);
      *reinterpret_cast<uint2*>(p16 + i * 4) = u;
    }
  }
}

__global__ void step_begin_kernel(float* state) {
  MTP_PDL_ENTRY();
  state[0] += 1.0f;
  state[1] = 0.0f;
}

}  // namespace mtp

using namespace mtp;

extern "C" int mtp_optim_step_begin(float* state, mtp_stream_t stream) {
  MTP_REQUIRE(state, "mtp_optim_step_begin: null pointer");
  (void)launch_k(step_begin_kernel, 1, 1, 0, reinterpret_cast<cudaStream_t>(stream), state);
  return check_launch("step_begin_kernel");
}

extern "C" int mtp_sumsq_f32(const float* x, size_t n, float* out, mtp_stream_t stream) {
  MTP_REQUIRE(x && out && n % 4 == 0, "mtp_sumsq_f32: bad args");
  if (n == 0) return MTP_OK;
  const size_t n4 = n / 4;
  const int grid = (int)std::min<size_t>((n4 + 255) / 256, (size_t)num_sms() * 8);
  (void)launch_k(sumsq_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), x, n4, out);
  return check_launch("sumsq_kernel");
}

extern "C" int mtp_sumsq_bf16(const void* x_bf16, size_t n, float* out, mtp_stream_t stream) {
  MTP_REQUIRE(x_bf16 && out && n % 8 == 0 && ((uintptr_t)x_bf16 & 15) == 0, "mtp_sumsq_bf16: bad args");
  if (n == 0) return MTP_OK;
  const size_t n8 = n / 8;
  const int grid = (int)std::min<size_t>((n8 + 255) / 256, (size_t)num_sms() * 8);
  (void)launch_k(sumsq_bf16_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(x_bf16), n8, out);
  return check_launch("sumsq_bf16_kernel");
}

extern "C" int mtp_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, const uint8_t* chunk_group,
                              const float* group_lr_scale, const float* group_weight_decay, const float* state, size_t n, float lr0,
                              float eta_min, int t_max, float beta1, float beta2, float eps, float max_norm, float grad_scale,
                              mtp_stream_t stream) {
  return mtp_adamw_step_mixed(p, g, nullptr, 0, m, v, p_bf16, chunk_group, group_lr_scale, group_weight_decay, state, n, lr0, eta_min, t_max,
                              beta1, beta2, eps, max_norm, grad_scale, stream);
}

extern "C" int mtp_adamw_step_mixed(float* p, const float* g, const void* g_bf16, size_t bf16_from, float* m, float* v, void* p_bf16,
                                    const uint8_t* chunk_group, const float* group_lr_scale, const float* group_weight_decay,
                                    const float* state, size_t n, float lr0, float eta_min, int t_max, float beta1, float beta2, float eps,
                                    float max_norm, float grad_scale, mtp_stream_t stream) {
  MTP_REQUIRE(p && g && m && v && chunk_group && group_lr_scale && group_weight_decay && state, "mtp_adamw_step: null pointer");
  MTP_REQUIRE(bf16_from % 4 == 0, "mtp_adamw_step_mixed: bf16_from must be a multiple of 4");
  MTP_REQUIRE(n % 64 == 0, "mtp_adamw_step: n=%zu must be a multiple of 64 (chunked group table)", n);
  if (n == 0) return MTP_OK;
  AdamWArgs a{lr0, eta_min, beta1, beta2, eps, max_norm, grad_scale, t_max};
  const size_t n4 = n / 4;
  const int grid = (int)std::min<size_t>((n4 + 255) / 256, (size_t)num_sms() * 16);
  (void)launch_k(adamw_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), chunk_group,
                                                                      group_lr_scale, group_weight_decay, state, n4, a,
                                                                      reinterpret_cast<const __nv_bfloat16*>(g_bf16), bf16_from / 4);
  return check_launch("adamw_kernel");
}
