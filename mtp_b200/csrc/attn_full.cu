// Dense multi-head attention with decomposed relative-position bias, forward (flash-style: no N x N in HBM).
//   S[q,j] = q^.k_j + q^.Rh[qy - jy + Hp - 1] + q^.Rw[qx - jx + Wp - 1],  q^ = q * hd^-0.5 ;  O = softmax(S) V
// [V]:90-111 (Attention.forward), [V]:142-193 (calc_rel_pos_spatial); SURVEY.md Appendix A.2.
//
// One CTA per (image, head, 64-query tile); keys/values stream through shared memory in tiles of 64 with an online
// softmax.  The rel-pos term factorises into two small per-query tables relh[q][jy] and relw[q][jx] that are built once
// per query tile, so the bias costs N*(Hp+Wp)*hd MACs instead of an N x N table.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace mtp {

constexpr int FA_BQ = 64, FA_BK = 64, FA_HD = 64, FA_LD = 68, FA_THREADS = 128;

// smem floats: Q, K, V, P tiles (64 x 68) + relh [64][gh] + relw [64][gw]
__host__ __device__ inline int fa_smem_floats(int gh, int gw) { return 4 * FA_BQ * FA_LD + FA_BQ * (gh + gw); }

// HILO (fp32-class mode): qkv [T, 6C] / out [T, 2C] hold hi | lo bf16 word pairs (lo at +3C / +C); arithmetic is fp32 throughout.
template <bool HILO>
__global__ void __launch_bounds__(FA_THREADS)
full_attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N, int gh, int gw, int C, int nH, int use_rel) {
  MTP_PDL_ENTRY();
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + FA_BQ * FA_LD;
  float* Vs = Ks + FA_BK * FA_LD;
  float* Ps = Vs + FA_BK * FA_LD;
  float* relh = Ps + FA_BQ * FA_LD;     // [64][gh]
  float* relw = relh + FA_BQ * gh;      // [64][gw]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * FA_BQ, n = blockIdx.y, b = blockIdx.z;
  const int C3 = (HILO ? 6 : 3) * C, LOQ = 3 * C;
  const float scale = 0.125f;
  const __nv_bfloat16* base = qkv + (size_t)b * N * C3 + n * FA_HD;

  // ---- load the (pre-scaled) query tile: warp handles rows, lane handles 2 dims
  for (int r = warp; r < FA_BQ; r += FA_THREADS / 32) {
    float2 v = make_float2(0.f, 0.f);
    if (q0 + r < N) {
      const __nv_bfloat16* src = base + (size_t)(q0 + r) * C3 + lane * 2;
      v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src));
      if (HILO) { const float2 l = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + LOQ)); v.x += l.x; v.y += l.y; }
      v.x *= scale; v.y *= scale;
    }
    *reinterpret_cast<float2*>(Qs + r * FA_LD + lane * 2) = v;
  }
  __syncthreads();
  if (use_rel) {
    for (int e = tid; e < FA_BQ * (gh + gw); e += FA_THREADS) {
      const int r = e / (gh + gw), c = e % (gh + gw);
      const int q = min(q0 + r, N - 1);
      const float* tab = c < gh ? rel_h + (size_t)(q / gw - c + gh - 1) * FA_HD : rel_w + (size_t)(q % gw - (c - gh) + gw - 1) * FA_HD;
      const float* qr = Qs + r * FA_LD;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < FA_HD; d += 4) {
        const float4 tv = __ldg(reinterpret_cast<const float4*>(tab + d));
        const float4 qv = *reinterpret_cast<const float4*>(qr + d);
        s += qv.x * tv.x + qv.y * tv.y + qv.z * tv.z + qv.w * tv.w;
      }
      if (c < gh) relh[r * gh + c] = s; else relw[r * gw + (c - gh)] = s;
    }
  }

  // thread (tq, tj): score micro-tile rows [4tq, 4tq+4) x cols [8tj, 8tj+8); output rows [4tq,4tq+4) x dims [4tj, 4tj+4) and [32+4tj, 32+4tj+4)
  const int tq = tid >> 3, tj = tid & 7;
  float m_run[4], l_run[4], o[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = -INFINITY; l_run[a] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[a][e] = 0.f;
  }

  for (int k0 = 0; k0 < N; k0 += FA_BK) {
    __syncthreads();                     // previous tile fully consumed (and relh/relw visible on the first pass)
    for (int r = warp; r < FA_BK; r += FA_THREADS / 32) {
      float2 kv = make_float2(0.f, 0.f), vv = make_float2(0.f, 0.f);
      if (k0 + r < N) {
        const __nv_bfloat16* src = base + (size_t)(k0 + r) * C3 + lane * 2;
        kv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + C));
        vv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + 2 * C));
        if (HILO) {
          const float2 kl = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + C + LOQ));
          const float2 vl = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + 2 * C + LOQ));
          kv.x += kl.x; kv.y += kl.y; vv.x += vl.x; vv.y += vl.y;
        }
      }
      *reinterpret_cast<float2*>(Ks + r * FA_LD + lane * 2) = kv;
      *reinterpret_cast<float2*>(Vs + r * FA_LD + lane * 2) = vv;
    }
    __syncthreads();

    float s[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 8; ++c) s[a][c] = 0.f;
    for (int d = 0; d < FA_HD; d += 4) {
      float4 qv[4], kv[8];
#pragma unroll
      for (int a = 0; a < 4; ++a) qv[a] = *reinterpret_cast<const float4*>(Qs + (tq * 4 + a) * FA_LD + d);
#pragma unroll
      for (int c = 0; c < 8; ++c) kv[c] = *reinterpret_cast<const float4*>(Ks + (tj + 8 * c) * FA_LD + d);   // cols tj, tj+8, ...
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          s[a][c] += qv[a].x * kv[c].x + qv[a].y * kv[c].y + qv[a].z * kv[c].z + qv[a].w * kv[c].w;
    }
    // bias + mask + online softmax (row statistics shared by the 8 threads tj = 0..7 of a row group: same warp)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = tq * 4 + a;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int j = k0 + tj + 8 * c;
        if (j < N) {
          if (use_rel) s[a][c] += relh[r * gh + j / gw] + relw[r * gw + j % gw];
          mx = fmaxf(mx, s[a][c]);
        } else {
          s[a][c] = -INFINITY;
        }
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      const float m_new = fmaxf(m_run[a], mx);
      const float corr = __expf(m_run[a] - m_new);          // exp(-inf) = 0 on the first tile
      float ps = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float p = __expf(s[a][c] - m_new);
        Ps[r * FA_LD + tj + 8 * c] = p;
        ps += p;
      }
      ps += __shfl_xor_sync(0xffffffffu, ps, 1);
      ps += __shfl_xor_sync(0xffffffffu, ps, 2);
      ps += __shfl_xor_sync(0xffffffffu, ps, 4);
      l_run[a] = l_run[a] * corr + ps;
      m_run[a] = m_new;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[a][e] *= corr;
    }
    __syncwarp();      // a row group's P values are produced and consumed by the same warp (rows 4tq.. of warp = tid>>5)
    for (int j = 0; j < FA_BK; ++j) {
      const float4 v0 = *reinterpret_cast<const float4*>(Vs + j * FA_LD + tj * 4);
      const float4 v1 = *reinterpret_cast<const float4*>(Vs + j * FA_LD + 32 + tj * 4);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float p = Ps[(tq * 4 + a) * FA_LD + j];
        o[a][0] += p * v0.x; o[a][1] += p * v0.y; o[a][2] += p * v0.z; o[a][3] += p * v0.w;
        o[a][4] += p * v1.x; o[a][5] += p * v1.y; o[a][6] += p * v1.z; o[a][7] += p * v1.w;
      }
    }
  }

#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int q = q0 + tq * 4 + a;
    if (q < N) {
      const float inv = 1.0f / l_run[a];
      uint2 u0, u1;
      u0.x = pack_bf16x2(o[a][0] * inv, o[a][1] * inv);
      u0.y = pack_bf16x2(o[a][2] * inv, o[a][3] * inv);
      u1.x = pack_bf16x2(o[a][4] * inv, o[a][5] * inv);
      u1.y = pack_bf16x2(o[a][6] * inv, o[a][7] * inv);
      __nv_bfloat16* orow = out + ((size_t)b * N + q) * (HILO ? 2 * C : C) + n * FA_HD;
      *reinterpret_cast<uint2*>(orow + tj * 4) = u0;
      *reinterpret_cast<uint2*>(orow + 32 + tj * 4) = u1;
      if (HILO) {
        float r[8];
        const uint32_t w4[4] = {u0.x, u0.y, u1.x, u1.y};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 h2 = unpack_bf16x2(w4[e]); r[2 * e] = o[a][2 * e] * inv - h2.x; r[2 * e + 1] = o[a][2 * e + 1] * inv - h2.y; }
        uint2 l0, l1;
        l0.x = pack_bf16x2(r[0], r[1]); l0.y = pack_bf16x2(r[2], r[3]);
        l1.x = pack_bf16x2(r[4], r[5]); l1.y = pack_bf16x2(r[6], r[7]);
        *reinterpret_cast<uint2*>(orow + C + tj * 4) = l0;
        *reinterpret_cast<uint2*>(orow + C + 32 + tj * 4) = l1;
      }
      if (lse && tj == 0) lse[((size_t)b * nH + n) * N + q] = m_run[a] + __logf(l_run[a]);
    }
  }
}

int launch_full_attn_fwd_tc(const void* qkv, const float* rel_h, const float* rel_w, void* out, float* lse, int B, int gh, int gw, int C,
                            int nH, cudaStream_t st);      // attn_full_tc.cu
int launch_full_attn_fwd_stream_tc(const void* qkv, const float* rel_h, const float* rel_w, void* out, float* lse, int B, int gh, int gw,
                                   int C, int nH, cudaStream_t st);      // attn_full_stream_tc.cu

template <bool HILO>
static int launch_full_attn_fwd_simt(const void* qkv, const float* rel_h, const float* rel_w, void* out, float* lse, int B, int gh, int gw,
                                     int C, int nH, cudaStream_t st) {
  const int N = gh * gw;
  const int smem = fa_smem_floats(gh, gw) * (int)sizeof(float);
  MTP_REQUIRE(smem <= 220 * 1024, "mtp_full_attn_fwd: grid %dx%d too large for the rel-pos tables in shared memory", gh, gw);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(full_attn_fwd_kernel<HILO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "full_attn_fwd smem attr: %s", cudaGetErrorString(e));
    attr_smem = smem;
  }
  const dim3 grid(ceil_div(N, FA_BQ), nH, B);
  (void)launch_k(full_attn_fwd_kernel<HILO>, grid, FA_THREADS, smem, st, reinterpret_cast<const __nv_bfloat16*>(qkv), rel_h, rel_w,
                 reinterpret_cast<__nv_bfloat16*>(out), lse, N, gh, gw, C, nH, rel_h != nullptr);
  return check_launch("full_attn_fwd_kernel");
}

}  // namespace mtp

using namespace mtp;

extern "C" int mtp_full_attn_fwd(const void* qkv_bf16, const float* rel_pos_h, const float* rel_pos_w, void* out_bf16, float* lse,
                                 int B, int gh, int gw, int C, int nH, mtp_stream_t stream) {
  MTP_REQUIRE(qkv_bf16 && out_bf16, "mtp_full_attn_fwd: null pointer");
  MTP_REQUIRE((rel_pos_h == nullptr) == (rel_pos_w == nullptr), "mtp_full_attn_fwd: give both rel-pos tables or neither");
  MTP_REQUIRE(B > 0 && gh > 0 && gw > 0 && C == nH * FA_HD, "mtp_full_attn_fwd: B=%d grid=%dx%d C=%d nH=%d unsupported (hd must be 64)", B, gh, gw, C, nH);
  const int N = gh * gw;
  static int stream_all = -1;      // A/B switch: MTP_DENSE_STREAM_ALL=1 sends short sequences through the streaming kernels too
  if (stream_all < 0) { const char* e = getenv("MTP_DENSE_STREAM_ALL"); stream_all = (e != nullptr && e[0] == '1') ? 1 : 0; }
  if (N <= 256 && gh <= 16 && gw <= 16 && !stream_all)       // tensor-core path: K/V of a head resident in shared memory
    return launch_full_attn_fwd_tc(qkv_bf16, rel_pos_h, rel_pos_w, out_bf16, lse, B, gh, gw, C, nH, reinterpret_cast<cudaStream_t>(stream));
  {       // long sequences: K / V streamed in 128-key blocks with an online softmax, still on the tensor cores
    static int simt = -1;
    if (simt < 0) { const char* e = getenv("MTP_DENSE_SIMT"); simt = (e != nullptr && e[0] == '1') ? 1 : 0; }      // A/B switch
    if (!simt && (2 * gh - 1 + 2 * gw - 1) * 65 * 4 <= 6 * 16384)      // both rel-pos tables fit the kernel's prologue staging area
      return launch_full_attn_fwd_stream_tc(qkv_bf16, rel_pos_h, rel_pos_w, out_bf16, lse, B, gh, gw, C, nH, reinterpret_cast<cudaStream_t>(stream));
  }
  return launch_full_attn_fwd_simt<false>(qkv_bf16, rel_pos_h, rel_pos_w, out_bf16, lse, B, gh, gw, C, nH, reinterpret_cast<cudaStream_t>(stream));
}

/* fp32-class mode ("fp32x3"): qkv [T, 6C] / out [T, 2C] as hi | lo word pairs, fp32 arithmetic (streaming SIMT kernel at every N) */
extern "C" int mtp_full_attn_fwd_hilo(const void* qkv_hilo, const float* rel_pos_h, const float* rel_pos_w, void* out_hilo, int B, int gh,
                                      int gw, int C, int nH, mtp_stream_t stream) {
  MTP_REQUIRE(qkv_hilo && out_hilo, "mtp_full_attn_fwd_hilo: null pointer");
  MTP_REQUIRE((rel_pos_h == nullptr) == (rel_pos_w == nullptr), "mtp_full_attn_fwd_hilo: give both rel-pos tables or neither");
  MTP_REQUIRE(B > 0 && gh > 0 && gw > 0 && C == nH * FA_HD, "mtp_full_attn_fwd_hilo: unsupported geometry");
  return launch_full_attn_fwd_simt<true>(qkv_hilo, rel_pos_h, rel_pos_w, out_hilo, nullptr, B, gh, gw, C, nH, reinterpret_cast<cudaStream_t>(stream));
}
