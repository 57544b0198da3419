// Persistent, warp-specialised bf16 GEMM for sm_100a:  TMA -> 128B-swizzled smem ring -> tcgen05.mma (cta_group::1,
// M=128, N=BN) -> fp32 accumulators in TMEM (double-buffered) -> fused epilogue.
//
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane)
//   warps 2..9  : epilogue (two warps per TMEM lane quarter taking alternate 32-column chunks; tcgen05.ld pipelined
//                 one chunk ahead, bias staged in smem per tile, aux operands loaded before the TMEM wait)
//
// Both operands may be K-major (row = m or n, k contiguous) or MN-major (row = k, m/n contiguous), which covers
// forward (K,K), dgrad (K,MN) and wgrad (MN,MN) without materialising any transpose.  See include/mtp_b200.h.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <queue>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace mtp {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int EPI_THREADS = 256;
constexpr int SMEM_BUDGET = 200 * 1024;

template <int BN, bool PAIR = false>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;      // PAIR: each CTA stages half of the pair's B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BN * 4 /*bias*/;
};

struct EpiParams {
  int mode, ldo;
  const float* bias;
  void* out;
  void* out2;
  const void* aux;
  const float* row_scale;
  int rows_per_group, pos_rows, accumulate, ps_h, ps_w, ps_cout;
  float* colsum;
  int b_static;
  float* sumsq;
  int hilo, out_lo;      // fp32-class mode (see mtp_epilogue.hilo)
};

// ---- epilogue ---------------------------------------------------------------------------------------------------------
// tcgen05.ld (32x32b) hands every lane 32 consecutive columns of ONE accumulator row, so storing straight from that layout makes
// each warp-wide store touch 32 different rows with 16 bytes apiece: partial-sector requests that throttle L2 (measured: the
// 9.6 MB output of the qkv GEMM cost 5.3 us of a 19 us launch).  Each group of four lanes therefore first transposes its 4 rows x
// 4 column blocks through warp shuffles: afterwards a lane owns four PIECES, piece p = 8 columns of row (lane & ~3) | ((lane & 3)
// ^ {0,2,1,3}[p]), and the four lanes of a group cover 64 contiguous bytes of a row per 16-byte access (bf16 outputs: columns
// i*8 + k; fp32 outputs: two runs i*4 + {0,16} + k so that every access still fills whole 32-byte sectors).  All epilogue
// arithmetic, the aux loads and the stores run in that layout.
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ bool mode_is_f32(int mode) { return mode == MTP_EPI_F32_RESID || mode == MTP_EPI_F32_POS || mode == MTP_EPI_F32; }

// column (within the 32-column chunk) of element k of this lane's pieces
template <bool F32>
__device__ __forceinline__ int piece_col(int lane, int k) {
  const int i = lane & 3;
  return F32 ? ((k >> 2) * 16 + i * 4 + (k & 3)) : (i * 8 + k);
}
// row (within the warp's 32 rows) of piece p
__device__ __forceinline__ int piece_row(int lane, int p) {
  const int perm = (p == 0) ? 0 : (p == 1) ? 2 : (p == 2) ? 1 : 3;
  return (lane & ~3) | ((lane & 3) ^ perm);
}

template <bool F32>
__device__ __forceinline__ void lane_transpose(const float (&v)[32], float (&t)[4][8], int lane) {
  const bool b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
#define MTP_BLK(c, k) v[F32 ? (((k) >> 2) * 16 + (c) * 4 + ((k) & 3)) : ((c) * 8 + (k))]
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float x0 = MTP_BLK(0, k), x1 = MTP_BLK(1, k), x2 = MTP_BLK(2, k), x3 = MTP_BLK(3, k);
    // stage 1 (partner lane ^ 2): keep the block pair {2*b1, 2*b1+1} of my row, fetch the same pair of the partner's row
    const float keep_lo = b1 ? x2 : x0, keep_hi = b1 ? x3 : x1;
    const float recv_lo = __shfl_xor_sync(0xffffffffu, b1 ? x0 : x2, 2);
    const float recv_hi = __shfl_xor_sync(0xffffffffu, b1 ? x1 : x3, 2);
    // stage 2 (partner lane ^ 1): keep block (lane & 3) of both rows, fetch it from the partner's two rows
    t[0][k] = b0 ? keep_hi : keep_lo;
    t[1][k] = b0 ? recv_hi : recv_lo;
    t[2][k] = __shfl_xor_sync(0xffffffffu, b0 ? keep_lo : keep_hi, 1);
    t[3][k] = __shfl_xor_sync(0xffffffffu, b0 ? recv_lo : recv_hi, 1);
  }
#undef MTP_BLK
}

// aux operand of this lane's four pieces, fetched BEFORE waiting on the TMEM load so the two latencies overlap
struct AuxRegs { float4 f[8]; };

// m[p] = global row of piece p (callers pass ok[p] = row is in range); n = first column of the 32-column chunk
__device__ __forceinline__ void load_aux(const EpiParams& ep, AuxRegs& a, const int (&m)[4], const bool (&ok)[4], int n, int N, int lane) {
  const int i = lane & 3;
  if (ep.mode == MTP_EPI_F32_RESID || ep.mode == MTP_EPI_F32_POS || (ep.mode == MTP_EPI_F32 && ep.accumulate)) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (!ok[p]) continue;
      const float* r = ep.mode == MTP_EPI_F32_RESID ? reinterpret_cast<const float*>(ep.aux) + (size_t)m[p] * ep.ldo
                       : ep.mode == MTP_EPI_F32_POS ? reinterpret_cast<const float*>(ep.aux) + (size_t)(m[p] % ep.pos_rows) * N
                                                    : reinterpret_cast<const float*>(ep.out) + (size_t)m[p] * ep.ldo;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = n + h * 16 + i * 4;
        if (col < N) a.f[p * 2 + h] = *reinterpret_cast<const float4*>(r + col);
      }
    }
  } else if (ep.mode == MTP_EPI_BF16_DGELU) {
    const int col = n + i * 8;
#pragma unroll
    for (int p = 0; p < 4; ++p)
      if (ok[p] && col < N)
        a.f[p] = *reinterpret_cast<const float4*>(reinterpret_cast<const __nv_bfloat16*>(ep.aux) + (size_t)m[p] * ep.ldo + col);
  }
}

template <bool HILO>
__device__ __forceinline__ void epilogue_pieces(const EpiParams& ep, float (&t)[4][8], const AuxRegs& a, const float* bias_s,
                                                const int (&m)[4], const bool (&ok)[4], int n, int N, int lane, float& sq) {
  const int i = lane & 3;
  const bool f32 = mode_is_f32(ep.mode);
  if (bias_s != nullptr) {             // smem; columns >= N hold zeros
    float b[8];
    if (f32) {
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(bias_s + i * 4);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(bias_s + 16 + i * 4);
    } else {
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(bias_s + i * 8);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(bias_s + i * 8 + 4);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int k = 0; k < 8; ++k) t[p][k] += b[k];
  }
  if (f32) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (!ok[p]) continue;
      float* o = reinterpret_cast<float*>(ep.out) + (size_t)m[p] * ep.ldo;
      const float s = (ep.mode == MTP_EPI_F32_RESID && ep.row_scale) ? __ldg(ep.row_scale + m[p] / ep.rows_per_group) : 1.0f;
      const bool add = ep.mode != MTP_EPI_F32 || ep.accumulate;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = n + h * 16 + i * 4;
        if (col < N) {
          float4 x = make_float4(s * t[p][h * 4], s * t[p][h * 4 + 1], s * t[p][h * 4 + 2], s * t[p][h * 4 + 3]);
          if (add) {
            const float4 y = a.f[p * 2 + h];
            x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
          }
          *reinterpret_cast<float4*>(o + col) = x;
          sq += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;      // sum of squares of the stored values (ep.sumsq: gradient norm)
        }
      }
    }
    return;
  }
  const int col = n + i * 8;
  const bool col_ok = col < N;
  float cs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) cs[k] = 0.f;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (!ok[p] || !col_ok) continue;
    size_t row = (size_t)m[p];
    int ocol = col;
    if (ep.mode == MTP_EPI_BF16_PIXSHUF) {
      const int g = col / ep.ps_cout;
      ocol = col % ep.ps_cout;
      const int dy = g >> 1, dx = g & 1;
      const int hw = ep.ps_h * ep.ps_w;
      const int b = m[p] / hw, rem = m[p] % hw;
      const int y = rem / ep.ps_w, x = rem % ep.ps_w;
      row = ((size_t)b * 2 * ep.ps_h + 2 * y + dy) * (2 * ep.ps_w) + 2 * x + dx;
    }
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + row * ep.ldo + ocol;
    if (ep.mode == MTP_EPI_BF16_GELU) {
      if (ep.out2 != nullptr) store_bf16x8(reinterpret_cast<__nv_bfloat16*>(ep.out2) + row * ep.ldo + ocol, t[p]);
#pragma unroll
      for (int k = 0; k < 8; ++k) t[p][k] = gelu_erf(t[p][k]);
    } else if (ep.mode == MTP_EPI_BF16_DGELU) {
      const float4 hv = a.f[p];
      const uint32_t w[4] = {__float_as_uint(hv.x), __float_as_uint(hv.y), __float_as_uint(hv.z), __float_as_uint(hv.w)};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack_bf16x2(w[q]);
        t[p][2 * q] *= gelu_erf_grad(f.x);
        t[p][2 * q + 1] *= gelu_erf_grad(f.y);
      }
    }
    store_bf16x8(o, t[p]);
    if (HILO && ep.out_lo > 0) {          // fp32-class mode: second bf16 word of each value, lo = bf16(v - float(bf16(v)))
      float lo[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) lo[k] = t[p][k] - __bfloat162float(__float2bfloat16_rn(t[p][k]));
      store_bf16x8(o + ep.out_lo, lo);
    }
    if (ep.colsum != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) cs[k] += __bfloat162float(__float2bfloat16_rn(t[p][k]));      // sums of the STORED values
    }
  }
  if (ep.colsum != nullptr) {       // column sums over the warp's 32 rows: the 8 lane groups hold the same columns
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 4);
      cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);
      cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16);
    }
    if (lane < 4 && col_ok) {
      float* dst = ep.colsum + col;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
    }
  }
}

// ---- specialised epilogue of one tile -----------------------------------------------------------------------------------------
// Same piece layout as above, for the modes that make up the training step (bf16 / GELU / GELU' / fp32 / fp32 + residual; N a multiple of
// 32).  The generic code resolves the mode, the piece addresses (64-bit multiplies), the DropPath row scale (an integer division) and the
// bounds for every piece of every chunk: ncu's source view shows ~600 warp instructions per 32 x 32 chunk without the GELU math, two
// thirds of them IMAD / ISETP / MOV / LDC, and with two epilogue warps per scheduler the tile's epilogue is issue-bound (1.3 us per pair of
// chunk columns; the last tile's epilogue is the exposed tail of every launch, tools/gemm_gaps.py).  Here everything that does not depend
// on the chunk is resolved once per tile and the mode is a template parameter.
__device__ __forceinline__ long long gtime();

// 16 bytes of the aux operand for each of the lane's four pieces (same byte offset from the piece's output address)
__device__ __forceinline__ void load_aux4(float4 (&a)[4], char* const (&o)[4], const bool (&ok)[4], ptrdiff_t d) {
#pragma unroll
  for (int p = 0; p < 4; ++p)
    if (ok[p]) a[p] = *reinterpret_cast<const float4*>(o[p] + d);
}

// bf16 outputs: plain / GELU (+ pre-activation copy) / GELU' (times the saved pre-activation)
// (row0 = first row of this warp's TMEM lane quarter; rows >= m_end are not stored)
template <int MODE>
__device__ __forceinline__ void epilogue_tile_bf16(const EpiParams* epp, uint32_t taddr, int n_chunks, int hsel, int row0, int m_end, int n0,
                                                const float* bsm, uint64_t* acc_bar, uint32_t acc_phase, long long* stamp) {
  const EpiParams& ep = *epp;
  const int lane = threadIdx.x & 31;
  int m[4];
  bool ok[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    m[p] = row0 + piece_row(lane, p);
    ok[p] = m[p] < m_end;
  }
  const int i = lane & 3;
  char* const out = reinterpret_cast<char*>(ep.out);
  char* o[4];                                                          // this lane's first element of piece p in chunk 0
#pragma unroll
  for (int p = 0; p < 4; ++p) o[p] = out + ((size_t)m[p] * ep.ldo + n0 + i * 8) * 2;
  // the GELU' pre-activation and the GELU pre-activation copy have the output's shape and pitch
  const ptrdiff_t aux_d = MODE == MTP_EPI_BF16_DGELU ? reinterpret_cast<const char*>(ep.aux) - out : 0;
  const bool has_out2 = MODE == MTP_EPI_BF16_GELU && ep.out2 != nullptr;
  const ptrdiff_t out2_d = has_out2 ? reinterpret_cast<char*>(ep.out2) - out : 0;
  float* const colsum = ep.colsum;
  float4 aux[4];
  int c = hsel;
  if (MODE == MTP_EPI_BF16_DGELU && c < n_chunks) load_aux4(aux, o, ok, aux_d + c * 64);      // requested before the accumulator is complete
  mbar_wait(acc_bar, acc_phase);
  if (stamp != nullptr) *stamp = gtime();
  tc_fence_after();
  uint32_t r[32];
  if (c < n_chunks) tmem_ld_32x32(taddr + c * 32, r);
  for (; c < n_chunks; c += 2) {
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (c + 2 < n_chunks) tmem_ld_32x32(taddr + (c + 2) * 32, r);      // next chunk streams in while this one is processed
    float t[4][8];
    lane_transpose<false>(v, t, lane);
    if (bsm != nullptr) {                                              // smem
      const float* bs = bsm + c * 32 + i * 8;
      float b[8];
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(bs);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(bs + 4);
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int k = 0; k < 8; ++k) t[p][k] += b[k];
    }
    const int cb = c * 64;
    float cs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cs[k] = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (!ok[p]) continue;
      if (MODE == MTP_EPI_BF16_GELU) {
        if (has_out2) store_bf16x8(reinterpret_cast<__nv_bfloat16*>(o[p] + out2_d + cb), t[p]);
#pragma unroll
        for (int k = 0; k < 8; ++k) t[p][k] = gelu_erf(t[p][k]);
      } else if (MODE == MTP_EPI_BF16_DGELU) {
        const float4 hv = aux[p];
        const uint32_t w[4] = {__float_as_uint(hv.x), __float_as_uint(hv.y), __float_as_uint(hv.z), __float_as_uint(hv.w)};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16x2(w[q]);
          t[p][2 * q] *= gelu_erf_grad(f.x);
          t[p][2 * q + 1] *= gelu_erf_grad(f.y);
        }
      }
      store_bf16x8(reinterpret_cast<__nv_bfloat16*>(o[p] + cb), t[p]);
      if (colsum != nullptr) {
#pragma unroll
        for (int k = 0; k < 8; ++k) cs[k] += __bfloat162float(__float2bfloat16_rn(t[p][k]));      // sums of the STORED values
      }
    }
    if (MODE == MTP_EPI_BF16_DGELU && c + 2 < n_chunks) load_aux4(aux, o, ok, aux_d + (c + 2) * 64);      // into the registers just consumed
    if (colsum != nullptr) {       // column sums over the warp's 32 rows: the 8 lane groups hold the same columns
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 4);
        cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);
        cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16);
      }
      if (lane < 4) {
        float* dst = colsum + n0 + c * 32 + i * 8;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
      }
    }
  }
}

// 4 rows x 4 column blocks of 4 fp32 values: after the exchange the lane holds block (lane & 3) of the group's four rows (piece_row order)
__device__ __forceinline__ void lane_transpose16(const float (&v)[16], float (&t)[4][4], int lane) {
  const bool b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float x0 = v[k], x1 = v[4 + k], x2 = v[8 + k], x3 = v[12 + k];
    const float keep_lo = b1 ? x2 : x0, keep_hi = b1 ? x3 : x1;
    const float recv_lo = __shfl_xor_sync(0xffffffffu, b1 ? x0 : x2, 2);
    const float recv_hi = __shfl_xor_sync(0xffffffffu, b1 ? x1 : x3, 2);
    t[0][k] = b0 ? keep_hi : keep_lo;
    t[1][k] = b0 ? recv_hi : recv_lo;
    t[2][k] = __shfl_xor_sync(0xffffffffu, b0 ? keep_lo : keep_hi, 1);
    t[3][k] = __shfl_xor_sync(0xffffffffu, b0 ? recv_lo : recv_hi, 1);
  }
}

// fp32 outputs (weight gradients; the residual stream): units of 16 columns -- the four lanes of a group write 64 contiguous bytes of a
// row -- so that the aux operand (residual / accumulate target) can be double-buffered in registers: unit j+1's is requested before
// unit j's arithmetic (32 aux + 32 transposed + 32 prefetched accumulator registers per 32-column chunk did not fit beside the rest).
template <bool RESID, bool ADD>
__device__ __forceinline__ float epilogue_tile_f32(const EpiParams* epp, uint32_t taddr, int n_chunks, int hsel, int row0, int m_end, int n0,
                                                const float* bsm, uint64_t* acc_bar, uint32_t acc_phase, long long* stamp) {
  const EpiParams& ep = *epp;
  const int lane = threadIdx.x & 31;
  int m[4];
  bool ok[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    m[p] = row0 + piece_row(lane, p);
    ok[p] = m[p] < m_end;
  }
  float sq = 0.f;      // sum of squares of the values this thread stored
  const int i = lane & 3;
  char* const out = reinterpret_cast<char*>(ep.out);
  char* o[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) o[p] = out + ((size_t)m[p] * ep.ldo + n0 + i * 4) * 4;
  const ptrdiff_t aux_d = RESID ? reinterpret_cast<const char*>(ep.aux) - out : 0;      // accumulate: the output itself
  float rs[4] = {1.f, 1.f, 1.f, 1.f};                                  // DropPath scale of each piece's row
  if (RESID && ep.row_scale != nullptr) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
      if (ok[p]) rs[p] = __ldg(ep.row_scale + m[p] / ep.rows_per_group);
  }
  const bool want_sq = ep.sumsq != nullptr;
  const int n_units = hsel < n_chunks ? 2 * ((n_chunks - hsel + 1) / 2) : 0;      // this warp's chunks hsel, hsel + 2, ...: two units each
#define MTP_UNIT_COL(j) ((hsel + 2 * ((j) >> 1)) * 32 + ((j) & 1) * 16)
  float4 auxa[4], auxb[4];
  if (ADD && n_units > 0) {                  // requested before the accumulator is complete
#pragma unroll
    for (int p = 0; p < 4; ++p)
      if (ok[p]) auxa[p] = *reinterpret_cast<const float4*>(o[p] + aux_d + MTP_UNIT_COL(0) * 4);
  }
  mbar_wait(acc_bar, acc_phase);
  if (stamp != nullptr) *stamp = gtime();
  tc_fence_after();
  uint32_t r[16];
  if (n_units > 0) tmem_ld_32x16(taddr + MTP_UNIT_COL(0), r);
  // one unit = 16 columns of the lane quarter's 32 rows: wait for its accumulators (already requested into r), request the next unit's
  // accumulators and aux operand, then bias / row scale / + aux / store.  (A macro, not a function taking the two aux buffers by reference:
  // that left both buffers in local memory.)
#ifndef MTP_AUX_DOUBLE
#define MTP_AUX_DOUBLE 1
#endif
#define MTP_F32_UNIT(ACUR, ANEXT, HAS_NEXT, NEXT_COL, COL)                                                          \
  {                                                                                                               \
    tmem_ld_wait();                                                                                               \
    float v[16];                                                                                                  \
    _Pragma("unroll") for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]);                                  \
    if (HAS_NEXT) tmem_ld_32x16(taddr + (NEXT_COL), r);                                                           \
    if (MTP_AUX_DOUBLE && ADD && (HAS_NEXT)) {                                                                    \
      _Pragma("unroll") for (int p = 0; p < 4; ++p)                                                               \
        if (ok[p]) ANEXT[p] = *reinterpret_cast<const float4*>(o[p] + aux_d + (NEXT_COL) * 4);                    \
    }                                                                                                             \
    float t[4][4];                                                                                                \
    lane_transpose16(v, t, lane);                                                                                 \
    if (bsm != nullptr) {                                                                                         \
      const float4 b = *reinterpret_cast<const float4*>(bsm + (COL) + i * 4);                                     \
      _Pragma("unroll") for (int p = 0; p < 4; ++p) { t[p][0] += b.x; t[p][1] += b.y; t[p][2] += b.z; t[p][3] += b.w; } \
    }                                                                                                             \
    _Pragma("unroll") for (int p = 0; p < 4; ++p) {                                                               \
      if (!ok[p]) continue;                                                                                       \
      float4 x = make_float4(rs[p] * t[p][0], rs[p] * t[p][1], rs[p] * t[p][2], rs[p] * t[p][3]);                \
      if (ADD) { x.x += ACUR[p].x; x.y += ACUR[p].y; x.z += ACUR[p].z; x.w += ACUR[p].w; }                        \
      *reinterpret_cast<float4*>(o[p] + (COL) * 4) = x;                                                           \
      if (want_sq) sq += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;                                           \
    }                                                                                                             \
    if (!MTP_AUX_DOUBLE && ADD && (HAS_NEXT)) {                                                                   \
      _Pragma("unroll") for (int p = 0; p < 4; ++p)                                                               \
        if (ok[p]) ANEXT[p] = *reinterpret_cast<const float4*>(o[p] + aux_d + (NEXT_COL) * 4);                    \
    }                                                                                                             \
  }
  for (int j = 0; j < n_units; j += 2) {      // n_units is even
    const int col0 = MTP_UNIT_COL(j);
    const bool more = j + 2 < n_units;
#if MTP_AUX_DOUBLE
    MTP_F32_UNIT(auxa, auxb, true, col0 + 16, col0)
    MTP_F32_UNIT(auxb, auxa, more, col0 + 64, col0 + 16)
#else
    MTP_F32_UNIT(auxa, auxa, true, col0 + 16, col0)
    MTP_F32_UNIT(auxa, auxa, more, col0 + 64, col0 + 16)
#endif
  }
#undef MTP_F32_UNIT
#undef MTP_UNIT_COL
  return sq;
}

// -------------------------------------------------------------------------------------------------------------------
// Grouped persistent kernel: up to two independent GEMM problems (e.g. the dgrad and the wgrad of one Linear) share one
// launch.  Work items (tiles, or 256-row tile pairs in CL2 mode) of both problems are assigned to the CTAs by a
// longest-processing-time schedule computed on the host and passed BY VALUE, so all three warp roles walk the same list
// without any device-side scheduler traffic, idle SMs of one problem are filled by the other, and one prologue/epilogue
// tail is paid instead of two.
//
// CL2 (cta_group::2): the two CTAs of a cluster compute one 256 x BN tile with a single MMA stream issued by the leader
// (even rank).  Each CTA stages its 128 rows of A and HALF of B (BN/2 columns) and accumulates its 128 rows in its own TMEM;
// the tensor cores read the other half of B from the peer's shared memory.
constexpr int MAX_SLOTS = 148;
constexpr int MAX_ITEMS = 8;

struct GemmProblem {
  CUtensorMap tmA, tmB;
  EpiParams ep;
  int M, N, K, tiles_m, tiles_n, a_mn, b_mn;
};

struct Sched {
  uint16_t count[MAX_SLOTS];
  uint16_t item[MAX_SLOTS][MAX_ITEMS];
  int strided_total;   // > 0: the lists above are unused; slot s walks items s, s + slots, s + 2*slots, ... < strided_total
  int n_stages;        // depth of the smem ring actually used (<= GemmCfg::STAGES): fewer stages = less dynamic smem, so that the CTA of
                       // the NEXT launch can become resident beside this one (co-residency experiments, tools/gemm_gaps.py)
  int dbg_mode;        // tuning aid: 0 normal, 1 = skip the TMA loads (MMA pipeline only), 2 = skip the MMAs (TMA pipeline only)
  long long* dbg;      // optional: [gridDim.x][8] globaltimer stamps of the pipeline phases (tuning aid, mtp_gemm_set_debug)
};

__device__ __forceinline__ long long gtime() {
  long long t;
  // "memory": without it the read may be scheduled ahead of a preceding bar.sync (the end-of-CTA stamp then records when the FIRST warp
  // reached the final barrier, which read as a 5 us "turnaround" after the CTA's end for a long time)
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
  return t;
}
#define MTP_STAMP(i) do { if (sched.dbg) sched.dbg[blockIdx.x * 8 + (i)] = gtime(); } while (0)

#ifndef MTP_GEMM_MINBLOCKS
#define MTP_GEMM_MINBLOCKS 1      // 2: cap registers so that two CTAs (this launch's and the next one's) fit one SM -- co-residency experiments
#endif
// ESET: which epilogue code the kernel carries.  0 = the generic epilogue (every mode, any N);  1 = GELU, 2 = fp32 + residual,
// 3 = bf16 | fp32 (a dgrad and / or a wgrad), 4 = GELU' | fp32 -- the specialised per-tile epilogues above, N a multiple of 32, picked by
// the host per launch (pick_eset).  (All variants inside one kernel made ptxas spill inside the chunk loops: 168 registers.)
template <int BN, bool CL2, bool HILO, int ESET>
__global__ void __launch_bounds__(GEMM_THREADS, MTP_GEMM_MINBLOCKS)
gemm_bf16_kernel(const __grid_constant__ GemmProblem p0, const __grid_constant__ GemmProblem p1, const __grid_constant__ Sched sched) {
  using Cfg = GemmCfg<BN, CL2>;
  const int STAGES = sched.n_stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);   // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) MTP_STAMP(0);
  if (sched.dbg_mode == 4) return;      // measurement aid: an empty launch (step time without the GEMM work; results are garbage)
  const int crank = CL2 ? (int)cluster_ctarank() : 0;
  const int slot = CL2 ? blockIdx.x / 2 : blockIdx.x;
  const int n_slots = CL2 ? gridDim.x / 2 : gridDim.x;
  const int n_items = sched.strided_total > 0 ? (sched.strided_total - slot + n_slots - 1) / n_slots : sched.count[slot];
  const int groups0 = CL2 ? (p0.tiles_m + 1) / 2 : p0.tiles_m;
  const int items0 = groups0 * p0.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p0.tmA);
    tma_prefetch_desc(&p0.tmB);
    if (p1.tiles_m > 0) {
      tma_prefetch_desc(&p1.tmA);
      tma_prefetch_desc(&p1.tmB);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], (CL2 ? 2 : 1) * EPI_THREADS / 32);   // one arrive per epilogue warp (CL2: of both CTAs, on the leader)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CL2) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();       // peer barriers are initialised before any remote signal can arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above overlapped the previous kernel's tail (programmatic dependent launch).  Each warp role passes the
  // dependency wait itself: the producer first prefetches B tiles that do not depend on the predecessor (weights).
  if (threadIdx.x == 0) MTP_STAMP(1);

#define MTP_DECODE_ITEM(IT)                                                                      \
  const int item_ = sched.strided_total > 0 ? slot + (IT) * n_slots : sched.item[slot][IT];      \
  const GemmProblem& P = item_ < items0 ? p0 : p1;                                               \
  const int local_ = item_ < items0 ? item_ : item_ - items0;                                    \
  const int mg_ = CL2 ? (P.tiles_m + 1) / 2 : P.tiles_m;                                         \
  const int m0 = ((local_ % mg_) * (CL2 ? 2 : 1) + crank) * BM;                                  \
  const int n0 = (local_ / mg_) * BN;                                                            \
  const int kb1_ = (P.K + BK - 1) / BK;                                                          \
  const int k_blocks = HILO ? 3 * kb1_ : kb1_;
  // fp32-class mode (operands stored as [rows, 2K] = hi | lo bf16 words): the k loop runs three passes over K,
  // A_hi B_hi + A_hi B_lo + A_lo B_hi, by moving the k coordinate of the TMA boxes; everything downstream is unchanged
#define MTP_KA(kb) (!HILO || (kb) < kb1_ ? (kb) : (kb) - kb1_)                      /* hi, hi, lo (lo blocks start at kb1_) */
#define MTP_KB(kb) (!HILO || (kb) < 2 * kb1_ ? (kb) : (kb) - 2 * kb1_)              /* hi, lo, hi */

  if (sched.dbg_mode == 5) {
    // measurement aid: prologue + teardown only (barriers, TMEM alloc / dealloc, descriptor prefetch), no work
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp walks the loop (warp-uniform control flow keeps descriptors / coordinates in uniform registers);
    // one elected lane issues.  Issuing from inside a divergent `if (lane == 0)` makes the compiler wrap every UTMALDG /
    // UTCHMMA in an ELECT + R2UR.BROADCAST waterfall loop (measured: ~175 cycles per MMA instead of its N/2).
    {
      // B tiles of the first item's first k-blocks that are independent of the stream predecessor (ep.b_static: weights, saved
      // activations) are requested BEFORE the dependency wait, so their HBM latency overlaps the predecessor's tail
      int pre = 0;
      if (n_items > 0 && (sched.dbg_mode == 0 || sched.dbg_mode == 30)) {
        MTP_DECODE_ITEM(0)
        if (P.ep.b_static) {
          pre = min(STAGES, k_blocks);
          for (int kb = 0; kb < pre; ++kb) {
            uint8_t* sb = smem + kb * Cfg::STAGE_BYTES + Cfg::A_BYTES;
            if (elect_one()) {
              if (!CL2) {
                mbar_arrive_expect_tx(&full_bar[kb], Cfg::STAGE_BYTES);
                if (!P.b_mn) {
                  tma_load_2d(sb, &P.tmB, &full_bar[kb], MTP_KB(kb) * BK, n0);
                } else {
#pragma unroll
                  for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &P.tmB, &full_bar[kb], n0 + j * 64, kb * BK);
                }
              } else {
                if (crank == 0) mbar_arrive_expect_tx(&full_bar[kb], 2 * Cfg::STAGE_BYTES);
                if (!P.b_mn) {
                  tma_load_2d_2sm(sb, &P.tmB, &full_bar[kb], kb * BK, n0 + crank * (BN / 2));
                } else {
#pragma unroll
                  for (int j = 0; j < BN / 128; ++j) tma_load_2d_2sm(sb + j * 8192, &P.tmB, &full_bar[kb], n0 + crank * (BN / 2) + j * 64, kb * BK);
                }
              }
            }
            __syncwarp();
          }
        }
      }
      MTP_PDL_ENTRY();
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_items; ++it) {
        MTP_DECODE_ITEM(it)
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const bool b_done = it == 0 && kb < pre;      // this stage's B half (and its expect_tx) were issued above
          if (elect_one()) {
            if (sched.dbg_mode == 1) {
              if (!CL2 || crank == 0) mbar_arrive(&full_bar[stage]);
            } else if (!CL2) {
              if (!b_done) mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              if (!P.a_mn) {
                tma_load_2d(sa, &P.tmA, &full_bar[stage], MTP_KA(kb) * BK, m0);
              } else {
#pragma unroll
                for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &P.tmA, &full_bar[stage], m0 + j * 64, kb * BK);
              }
              if (b_done) {
              } else if (!P.b_mn) {
                tma_load_2d(sb, &P.tmB, &full_bar[stage], MTP_KB(kb) * BK, n0);
              } else {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &P.tmB, &full_bar[stage], n0 + j * 64, kb * BK);
              }
            } else {
              // both CTAs' bytes are credited to the LEADER's full barrier (only the leader waits on it and issues the MMAs)
              if (crank == 0 && !b_done) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
              if (!P.a_mn) {
                tma_load_2d_2sm(sa, &P.tmA, &full_bar[stage], kb * BK, m0);
              } else {
#pragma unroll
                for (int j = 0; j < BM / 64; ++j) tma_load_2d_2sm(sa + j * 8192, &P.tmA, &full_bar[stage], m0 + j * 64, kb * BK);
              }
              if (b_done) {
              } else if (!P.b_mn) {       // my half of the pair's B tile: columns [n0 + crank*BN/2, +BN/2)
                tma_load_2d_2sm(sb, &P.tmB, &full_bar[stage], kb * BK, n0 + crank * (BN / 2));
              } else {
#pragma unroll
                for (int j = 0; j < BN / 128; ++j) tma_load_2d_2sm(sb + j * 8192, &P.tmB, &full_bar[stage], n0 + crank * (BN / 2) + j * 64, kb * BK);
              }
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, elected lane issues)
    MTP_PDL_ENTRY();
    if (!CL2 || crank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = 0; it < n_items; ++it) {
        MTP_DECODE_ITEM(it)
        (void)m0; (void)n0;
        const uint32_t idesc = make_idesc_bf16(CL2 ? 2 * BM : BM, BN, P.a_mn != 0, P.b_mn != 0);
        // K-major: 16 bf16 = 32 B inside the 128 B swizzle row; 8-row groups are 1024 B apart (SBO).
        // MN-major: 16 k-rows = 2 swizzle atoms of 8 rows x 128 B = 2048 B; 64-wide MN atoms are 8192 B apart (LBO).
        const uint32_t a_lbo = P.a_mn ? 8192 : 16, b_lbo = P.b_mn ? 8192 : 16;
        const uint64_t a_step = P.a_mn ? 128 : 2, b_step = P.b_mn ? 128 : 2;      // start-address units of 16 B per k-step
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (it == 0 && kb == 0 && lane == 0) MTP_STAMP(2);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t a_desc0 = make_smem_desc(sa, a_lbo, 1024);
          const uint64_t b_desc0 = make_smem_desc(sb, b_lbo, 1024);
          if (elect_one()) {
            if (sched.dbg_mode == 2 && !CL2) {
              mbar_arrive(&empty_bar[stage]);
            } else {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                if (CL2) umma_bf16_2sm(d_tmem, a_desc0 + k * a_step, b_desc0 + k * b_step, idesc, (kb | k) != 0);
                else umma_bf16(d_tmem, a_desc0 + k * a_step, b_desc0 + k * b_step, idesc, (kb | k) != 0);
              }
              if (CL2) umma_commit_2sm_mcast(&empty_bar[stage], 0x3);   // releases the stage in BOTH CTAs
              else umma_commit(&empty_bar[stage]);                      // smem slot reusable once these MMAs retire
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
          if (CL2) umma_commit_2sm_mcast(&tmem_full[acc], 0x3);       // accumulators complete in both CTAs' TMEM
          else umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (lane == 0) {
          if (it == 0) MTP_STAMP(3);
          if (it == n_items - 1) MTP_STAMP(4);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    MTP_PDL_ENTRY();
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int hsel = (warp - 2) >> 2;       // which alternate chunks this warp takes
    const int et = threadIdx.x - 64;        // 0..255 within the epilogue group
    int acc = 0;
    uint32_t acc_phase = 0;
    float sq0 = 0.f, sq1 = 0.f;             // per problem: sum of squares of the fp32 outputs this thread stored
    for (int it = 0; it < n_items; ++it) {
      MTP_DECODE_ITEM(it)
      (void)k_blocks;
      const EpiParams& ep = P.ep;
      const int M = P.M, N = P.N;
      float* bsm = bias_s + acc * BN;
      if (ep.bias != nullptr) {             // stage this tile's bias slice once (zeros beyond N)
        for (int i = et; i < BN; i += EPI_THREADS) {
          const int n = n0 + i;
          bsm[i] = n < N ? __ldg(ep.bias + (ep.ps_cout > 0 ? n % ep.ps_cout : n)) : 0.f;      // ps_cout: period of the bias vector
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");      // bias visible; also keeps the 8 warps on the same item
      int pm[4];
      bool pok[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        pm[p] = m0 + q * 32 + piece_row(lane, p);
        pok[p] = pm[p] < M && sched.dbg_mode != 3;      // dbg 3: no epilogue stores (isolates the store drain at kernel end)
      }
      const bool f32 = mode_is_f32(ep.mode);
      const int n_chunks = sched.dbg_mode == 6 ? 0 : (min(BN, N - n0) + 31) / 32;      // dbg 6: accumulators are never read out
      // The aux operand (residual / GELU' pre-activation / accumulate target: global memory, ~1 us away) is requested early: the first
      // chunk's before the accumulator is complete, each following one as soon as its registers are free (a full double buffer spills: the
      // kernel sits at the 168-register cap of 10 warps).  Measured r2: the GELU' epilogue of the fc2 dgrad cost 8.5 us per 128 x 256 tile
      // (1.7 us for the GELU of fc1) because every chunk waited for its own aux loads.
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      float& sq_item = item_ < items0 ? sq0 : sq1;
      const int emode = ep.mode;
      if constexpr (ESET != 0) {
        const float* bsf = ep.bias != nullptr ? bsm : nullptr;
        long long* stamp5 = (sched.dbg != nullptr && it == n_items - 1 && threadIdx.x == 64) ? sched.dbg + blockIdx.x * 8 + 5 : nullptr;
        const int row0 = m0 + q * 32, m_end = sched.dbg_mode == 3 ? 0 : M;
        const EpiParams* epp = &P.ep;
        if constexpr (ESET == 1) {
          epilogue_tile_bf16<MTP_EPI_BF16_GELU>(epp, taddr, n_chunks, hsel, row0, m_end, n0, bsf, &tmem_full[acc], acc_phase, stamp5);
        } else if constexpr (ESET == 2) {
          sq_item += epilogue_tile_f32<true, true>(epp, taddr, n_chunks, hsel, row0, m_end, n0, bsf, &tmem_full[acc], acc_phase, stamp5);
        } else {
          if (emode == MTP_EPI_F32) sq_item += epilogue_tile_f32<false, false>(epp, taddr, n_chunks, hsel, row0, m_end, n0, bsf, &tmem_full[acc], acc_phase, stamp5);
          else epilogue_tile_bf16<ESET == 3 ? MTP_EPI_BF16 : MTP_EPI_BF16_DGELU>(epp, taddr, n_chunks, hsel, row0, m_end, n0, bsf, &tmem_full[acc], acc_phase, stamp5);
        }
      } else {
      AuxRegs aux_cur;
      int c = hsel;
      if (c < n_chunks) load_aux(ep, aux_cur, pm, pok, n0 + c * 32, N, lane);
      mbar_wait(&tmem_full[acc], acc_phase);
      if (it == n_items - 1 && threadIdx.x == 64) MTP_STAMP(5);
      tc_fence_after();
      uint32_t r[32];
      if (c < n_chunks) tmem_ld_32x32(taddr + c * 32, r);
      for (; c < n_chunks; c += 2) {
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (c + 2 < n_chunks) tmem_ld_32x32(taddr + (c + 2) * 32, r);      // next chunk streams in while this one is processed
        float t[4][8];
        if (f32) lane_transpose<true>(v, t, lane);
        else lane_transpose<false>(v, t, lane);
        epilogue_pieces<HILO>(ep, t, aux_cur, ep.bias != nullptr ? bsm + c * 32 : nullptr, pm, pok, n0 + c * 32, N, lane, sq_item);
        // the next chunk's aux operand goes into the registers just consumed; its latency overlaps the next TMEM wait and lane transpose
        if (c + 2 < n_chunks) load_aux(ep, aux_cur, pm, pok, n0 + (c + 2) * 32, N, lane);
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL2) mbar_arrive_remote(&tmem_empty[acc], 0);      // the leader's MMA thread waits for both CTAs' epilogues
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p0.ep.sumsq != nullptr) {           // one atomic per warp and launch
      sq0 = warp_sum(sq0);
      if (lane == 0 && sq0 != 0.f) atomicAdd(p0.ep.sumsq, sq0);
    }
    if (p1.tiles_m > 0 && p1.ep.sumsq != nullptr) {
      sq1 = warp_sum(sq1);
      if (lane == 0 && sq1 != 0.f) atomicAdd(p1.ep.sumsq, sq1);
    }
  }
#undef MTP_DECODE_ITEM
#undef MTP_KA
#undef MTP_KB

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) MTP_STAMP(6);
  if (CL2) cluster_sync_all();       // no CTA leaves while its peer may still signal its barriers
  // dbg 20..22 (tools/gemm_gaps.py): where the teardown time goes -- extra stamps / dealloc from the producer warp / no fence
  if (warp == (sched.dbg_mode == 21 ? 0 : 1)) {
    if (sched.dbg_mode == 20 && lane == 0) MTP_STAMP(2);
    if (sched.dbg_mode != 22) tc_fence_after();
    if (sched.dbg_mode == 20 && lane == 0) MTP_STAMP(3);
    if (CL2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    if (lane == 0) MTP_STAMP(7);
  }
}

// ===================================================================================================================
// Variant 2: ONE tile per CTA, TWO CTAs per SM (mtp_gemm_set_variant(2) / MTP_GEMM_VARIANT=2).
//
// Motivation (profiles/r2_summary.md section 4): behind the persistent kernel above the epilogue of every SM's last tile is exposed, and a
// CTA of 200 KB smem / 512 TMEM columns must fully retire before its successor can become resident.  (An experiment, not the default: the step
// was slower with it -- the shallow operand ring costs more than the overlap gains)  Here a CTA owns 98 KB smem (2-3
// stage ring), one 128 x BN accumulator (<= 256 TMEM columns) and 192 threads (TMA warp, MMA warp, 4 epilogue warps), so two fit one SM:
// while one CTA runs its epilogue the other one's mainloop owns the tensor core, and the NEXT launch's CTAs move into the slots freed by
// early finishers, run their prologue and wait at griddepcontrol.wait before their predecessor grid has drained.  Tiles are handed to the SMs
// by the hardware block scheduler in blockIdx order (problem 0 first), which also balances the tail dynamically.
constexpr int G2_THREADS = 192, G2_EPI_THREADS = 128;
template <int BN>
struct Gemm2Cfg {
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN <= 128 ? 3 : 2;
  static constexpr int TMEM_COLS = BN <= 128 ? 128 : 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 128 /*barriers*/ + BN * 4 /*bias*/;
  static_assert(2 * (SMEM_BYTES + 1024) <= 228 * 1024, "two CTAs must fit one SM");
};

template <int BN>
__global__ void __launch_bounds__(G2_THREADS, 2)
gemm2_bf16_kernel(const __grid_constant__ GemmProblem p0, const __grid_constant__ GemmProblem p1, int items0, int dbg_mode) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (dbg_mode == 4) return;
  const int item = blockIdx.x;
  const GemmProblem& P = item < items0 ? p0 : p1;
  const int local = item < items0 ? item : item - items0;
  const int m0 = (local % P.tiles_m) * BM, n0 = (local / P.tiles_m) * BN;
  const int k_blocks = (P.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform loop, elected lane issues)
    int pre = 0;
    if (P.ep.b_static) {          // B tiles that do not depend on the stream predecessor: requested before the dependency wait
      pre = min(STAGES, k_blocks);
      for (int kb = 0; kb < pre; ++kb) {
        uint8_t* sb = smem + kb * Cfg::STAGE_BYTES + Cfg::A_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[kb], Cfg::STAGE_BYTES);
          if (!P.b_mn) {
            tma_load_2d(sb, &P.tmB, &full_bar[kb], kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &P.tmB, &full_bar[kb], n0 + j * 64, kb * BK);
          }
        }
        __syncwarp();
      }
    }
    MTP_PDL_ENTRY();
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < k_blocks; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
      uint8_t* sb = sa + Cfg::A_BYTES;
      const bool b_done = kb < pre;
      if (elect_one()) {
        if (!b_done) mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        if (!P.a_mn) {
          tma_load_2d(sa, &P.tmA, &full_bar[stage], kb * BK, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &P.tmA, &full_bar[stage], m0 + j * 64, kb * BK);
        }
        if (b_done) {
        } else if (!P.b_mn) {
          tma_load_2d(sb, &P.tmB, &full_bar[stage], kb * BK, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &P.tmB, &full_bar[stage], n0 + j * 64, kb * BK);
        }
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    MTP_PDL_ENTRY();
    const uint32_t idesc = make_idesc_bf16(BM, BN, P.a_mn != 0, P.b_mn != 0);
    const uint32_t a_lbo = P.a_mn ? 8192 : 16, b_lbo = P.b_mn ? 8192 : 16;
    const uint64_t a_step = P.a_mn ? 128 : 2, b_step = P.b_mn ? 128 : 2;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < k_blocks; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
      const uint32_t sb = sa + Cfg::A_BYTES;
      const uint64_t a_desc0 = make_smem_desc(sa, a_lbo, 1024);
      const uint64_t b_desc0 = make_smem_desc(sb, b_lbo, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, a_desc0 + k * a_step, b_desc0 + k * b_step, idesc, (kb | k) != 0);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(tmem_full);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: warp w owns TMEM lane quarter w & 3, all column chunks
    MTP_PDL_ENTRY();
    const int q = warp & 3;
    const int et = threadIdx.x - 64;
    const EpiParams& ep = P.ep;
    const int M = P.M, N = P.N;
    if (ep.bias != nullptr) {
      for (int i = et; i < BN; i += G2_EPI_THREADS) {
        const int n = n0 + i;
        bias_s[i] = n < N ? __ldg(ep.bias + (ep.ps_cout > 0 ? n % ep.ps_cout : n)) : 0.f;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(G2_EPI_THREADS) : "memory");
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    int pm[4];
    bool pok[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      pm[p] = m0 + q * 32 + piece_row(lane, p);
      pok[p] = pm[p] < M && dbg_mode != 3;
    }
    const bool f32 = mode_is_f32(ep.mode);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_chunks = (min(BN, N - n0) + 31) / 32;
    float sq = 0.f;
    uint32_t r[32];
    if (n_chunks > 0) tmem_ld_32x32(taddr, r);
    for (int c = 0; c < n_chunks; ++c) {
      AuxRegs aux;
      load_aux(ep, aux, pm, pok, n0 + c * 32, N, lane);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (c + 1 < n_chunks) tmem_ld_32x32(taddr + (c + 1) * 32, r);
      float t[4][8];
      if (f32) lane_transpose<true>(v, t, lane);
      else lane_transpose<false>(v, t, lane);
      epilogue_pieces<false>(ep, t, aux, ep.bias != nullptr ? bias_s + c * 32 : nullptr, pm, pok, n0 + c * 32, N, lane, sq);
    }
    if (ep.sumsq != nullptr) {
      sq = warp_sum(sq);
      if (lane == 0 && sq != 0.f) atomicAdd(ep.sumsq, sq);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, ld] matrix with `cols` valid columns; box = [box_rows, 64 cols], 128B swizzle.
static int make_tmap(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return set_error(MTP_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MTP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
  return MTP_OK;
}

struct HostProblem {
  const void* A; int lda, a_mn;
  const void* B; int ldb, b_mn;
  int M, N, K;
  EpiParams ep;
};

// Cycles of one 64-deep k-block of a BN-wide tile (fit to the in-kernel phase stamps, tools/gemm_phases.py: 0.27 / 0.27 / 0.31 /
// 0.34 us for BN = 64 / 128 / 192 / 256, 0.27 us for a 256-wide pair): the MMAs take 2*BN cycles, the operand fill of the SM
// ~400 cycles for the A box plus ~1 cycle per 128-byte row of B, and a cta_group::2 pair halves the B rows each SM stages.
static double kblock_cycles(int bn, bool cl2) { return std::max(2.0 * bn, 400.0 + 1.05 * (cl2 ? bn / 2 : bn)); }
static const double kTileFixedCycles = 2500.0;      // epilogue / pipeline fill per tile
static const double kClusterLaunchCycles = 5000.0;  // a cluster launch starts / retires ~2.5 us later than a plain one (tools/gemm_gaps.py)

// Longest-processing-time schedule of the work items of up to two problems over the slots; returns the makespan (cycles).
static double build_schedule(const HostProblem* pr, int np, int bn, bool cl2, Sched* out) {
  const int slots = cl2 ? num_sms() / 2 : num_sms();
  struct Item { int id; double cost; };
  std::vector<Item> items;
  int base = 0;
  for (int p = 0; p < np; ++p) {
    const int tiles_m = ceil_div(pr[p].M, BM), tiles_n = ceil_div(pr[p].N, bn);
    const int n = (cl2 ? (tiles_m + 1) / 2 : tiles_m) * tiles_n;
    const double c = ceil_div(pr[p].K, BK) * (pr[p].ep.hilo ? 3 : 1) * kblock_cycles(bn, cl2) + kTileFixedCycles;
    for (int i = 0; i < n; ++i) items.push_back({base + i, c});
    base += n;
  }
  auto strided = [&]() {
    // too many items for the by-value lists: every slot strides through the item range (items of one problem cost the same,
    // so this is the same balance as LPT up to one item at the boundary between the problems)
    std::vector<double> load(slots, 0.0);
    for (size_t i = 0; i < items.size(); ++i) load[i % slots] += items[i].cost;
    if (out) {
      out->strided_total = base;
      for (int s = 0; s < MAX_SLOTS; ++s) out->count[s] = s < slots ? 1 : 0;     // all slots are launched
    }
    return *std::max_element(load.begin(), load.end());
  };
  if (out) out->strided_total = 0;
  if (base > MAX_ITEMS * slots) return strided();
  std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.cost > b.cost; });
  std::vector<double> load(slots, 0.0);
  std::vector<int> cnt(slots, 0);
  // min-heap over (load, slot)
  std::priority_queue<std::pair<double, int>, std::vector<std::pair<double, int>>, std::greater<std::pair<double, int>>> pq;
  for (int s = 0; s < slots; ++s) pq.push({0.0, s});
  for (const Item& it : items) {
    auto top = pq.top();
    pq.pop();
    const int s = top.second;
    if (cnt[s] >= MAX_ITEMS) {
      std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.id < b.id; });
      return strided();
    }
    if (out) out->item[s][cnt[s]] = (uint16_t)it.id;
    ++cnt[s];
    load[s] = top.first + it.cost;
    pq.push({load[s], s});
  }
  if (out) {
    for (int s = 0; s < MAX_SLOTS; ++s) out->count[s] = s < slots ? (uint16_t)cnt[s] : 0;
    // Order inside a slot: the epilogue of a CTA's LAST item cannot hide behind a following mainloop, so items with an expensive
    // epilogue (GELU / GELU': ~20 ALU instructions per element) go first and the cheapest epilogue (plain fp32 / bf16 store) last.
    if (np == 2) {
      auto heavy = [&](int id) {
        const int items0 = (cl2 ? (ceil_div(pr[0].M, BM) + 1) / 2 : ceil_div(pr[0].M, BM)) * ceil_div(pr[0].N, bn);
        const int m = pr[id < items0 ? 0 : 1].ep.mode;
        return (m == MTP_EPI_BF16_GELU || m == MTP_EPI_BF16_DGELU) ? 0 : 1;      // sort key: heavy epilogues first
      };
      for (int s = 0; s < slots; ++s)
        std::stable_sort(out->item[s], out->item[s] + cnt[s], [&](uint16_t a, uint16_t b) { return heavy(a) < heavy(b); });
    }
  }
  return *std::max_element(load.begin(), load.end());
}

// config choice (cached per shape signature): minimise the LPT makespan over tile widths and single / paired CTAs
struct Config { int bn; bool cl2; Sched sched; };

static const Config* get_config(const HostProblem* pr, int np, int force_bn) {
  static std::map<std::vector<int>, Config> cache;      // node-based: returned pointers stay valid
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  std::vector<int> key = {force_bn, np, num_sms()};
  for (int p = 0; p < np; ++p) { key.push_back(pr[p].M); key.push_back(pr[p].N); key.push_back(pr[p].K); key.push_back(pr[p].b_mn + 2 * pr[p].ep.hilo); }
  auto it = cache.find(key);
  if (it != cache.end()) return &it->second;
  bool any_bmn = false, all_pairable = true;
  for (int p = 0; p < np; ++p) { any_bmn |= pr[p].b_mn != 0; all_pairable &= ceil_div(pr[p].M, BM) >= 2 && !pr[p].ep.hilo; }
  Config best;
  best.bn = 0;
  double best_cost = 1e300;
  const int cand[4] = {256, 192, 128, 64};
  for (int cl = 0; cl < 2; ++cl) {
    for (int i = 0; i < 4; ++i) {
      const int bn = cand[i];
      if (force_bn) { if (cl != (force_bn >= 1000) || bn != force_bn % 1000) continue; if (cl == 1 && !all_pairable) continue; }
      else if (cl == 1 && !all_pairable) continue;
      if (cl == 1 && any_bmn && bn % 128 != 0) continue;     // MN-major B is fetched in 64-column boxes: a pair needs an even count
      double c = build_schedule(pr, np, bn, cl == 1, nullptr);
      if (c >= 0 && cl == 1) c += kClusterLaunchCycles;
      if (c >= 0 && c < best_cost) { best_cost = c; best.bn = bn; best.cl2 = cl == 1; }
    }
  }
  if (best.bn == 0) return nullptr;
  build_schedule(pr, np, best.bn, best.cl2, &best.sched);
  return &cache.emplace(key, best).first->second;
}

static int g_last_config = 0;
static long long* g_gemm_dbg = nullptr;
static int g_gemm_dbg_mode = 0;
static int g_gemm_max_stages = 0;      // 0 = as many as fit the smem budget

// Which specialised epilogue set covers this launch (gemm_bf16_kernel's ESET); 0 = none, use the generic epilogue.
static int pick_eset(const HostProblem* pr, int np) {
  if (g_gemm_dbg_mode == 30) return 0;      // tuning aid: the generic epilogue everywhere (A/B)
  bool gelu = false, resid = false, dgelu = false, plain = false;
  for (int p = 0; p < np; ++p) {
    const mtp::EpiParams& ep = pr[p].ep;
    if (ep.hilo || (pr[p].N & 31) != 0) return 0;
    switch (ep.mode) {
      case MTP_EPI_BF16: plain = true; break;
      case MTP_EPI_BF16_GELU: gelu = true; break;
      case MTP_EPI_BF16_DGELU: dgelu = true; break;
      case MTP_EPI_F32_RESID: resid = true; break;
      case MTP_EPI_F32: if (ep.accumulate) return 0; break;
      default: return 0;
    }
  }
  if (gelu) return (np == 1) ? 1 : 0;
  if (resid) return (np == 1) ? 2 : 0;
  if (dgelu) return plain ? 0 : 4;
  return 3;
}

template <int BN, bool CL2, bool HILO = false, int ESET = 0>
static int launch_grouped(const HostProblem* pr, int np, const Sched& sched_in, cudaStream_t stream) {
  if constexpr (!HILO && !CL2) {          // the fp32-class (hi | lo, three-pass) instantiation exists for single CTAs only
    if (pr[0].ep.hilo) return launch_grouped<BN, CL2, true>(pr, np, sched_in, stream);
  }
  if constexpr (!HILO && ESET == 0) {
    switch (pick_eset(pr, np)) {
      case 1: return launch_grouped<BN, CL2, false, 1>(pr, np, sched_in, stream);
      case 2: return launch_grouped<BN, CL2, false, 2>(pr, np, sched_in, stream);
      case 3: return launch_grouped<BN, CL2, false, 3>(pr, np, sched_in, stream);
      case 4: return launch_grouped<BN, CL2, false, 4>(pr, np, sched_in, stream);
      default: break;
    }
  }
  Sched sched = sched_in;
  sched.dbg = g_gemm_dbg;
  sched.dbg_mode = g_gemm_dbg_mode;
  using Cfg = GemmCfg<BN, CL2>;
  sched.n_stages = g_gemm_max_stages > 0 ? std::max(2, std::min(Cfg::STAGES, g_gemm_max_stages)) : Cfg::STAGES;
  const int smem_bytes = Cfg::SMEM_BYTES - (Cfg::STAGES - sched.n_stages) * Cfg::STAGE_BYTES;
  GemmProblem gp[2];
  memset(gp, 0, sizeof(gp));
  for (int p = 0; p < np; ++p) {
    const HostProblem& h = pr[p];
    const int kcols = h.ep.hilo ? 2 * h.K : h.K;      // hi | lo words side by side
    int rc = h.a_mn ? make_tmap(&gp[p].tmA, h.A, h.K, h.M, h.lda, BK) : make_tmap(&gp[p].tmA, h.A, h.M, kcols, h.lda, BM);
    if (rc) return rc;
    rc = h.b_mn ? make_tmap(&gp[p].tmB, h.B, h.K, h.N, h.ldb, BK) : make_tmap(&gp[p].tmB, h.B, h.N, kcols, h.ldb, CL2 ? BN / 2 : BN);
    if (rc) return rc;
    gp[p].ep = h.ep;
    gp[p].M = h.M; gp[p].N = h.N; gp[p].K = h.K;
    gp[p].tiles_m = ceil_div(h.M, BM); gp[p].tiles_n = ceil_div(h.N, BN);
    gp[p].a_mn = h.a_mn; gp[p].b_mn = h.b_mn;
  }
  static bool attr_set_dev[64] = {};      // the attribute is per device (function handles are per context)
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool& attr_set = attr_set_dev[dev_ & 63];
  auto kern = gemm_bf16_kernel<BN, CL2, HILO, ESET>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "cudaFuncSetAttribute(gemm BN=%d): %s", BN, cudaGetErrorString(e));
    attr_set = true;
  }
  const int slots = CL2 ? num_sms() / 2 : num_sms();
  int used = 0;
  for (int s = 0; s < slots; ++s) if (sched.count[s] > 0) used = s + 1;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  cfg.attrs = attr;
  if (CL2) {
    cfg.gridDim = dim3(2 * used);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3(used);
  }
  if (pdl_enabled()) {
    attr[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
    ++cfg.numAttrs;
  }
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, gp[0], gp[1], sched);
  if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "gemm_bf16_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("gemm_bf16_kernel");
}

static int validate_problem(const HostProblem& h) {
  MTP_REQUIRE(h.A && h.B && h.ep.out, "mtp_gemm_bf16: null pointer");
  MTP_REQUIRE(h.M > 0 && h.N > 0 && h.K > 0, "mtp_gemm_bf16: empty problem M=%d N=%d K=%d", h.M, h.N, h.K);
  MTP_REQUIRE(h.N % 8 == 0, "mtp_gemm_bf16: N=%d must be a multiple of 8", h.N);
  MTP_REQUIRE(h.lda % 8 == 0 && h.ldb % 8 == 0, "mtp_gemm_bf16: lda/ldb must be multiples of 8 (got %d, %d)", h.lda, h.ldb);
  MTP_REQUIRE(((uintptr_t)h.A & 15) == 0 && ((uintptr_t)h.B & 15) == 0 && ((uintptr_t)h.ep.out & 15) == 0,
              "mtp_gemm_bf16: pointers must be 16-byte aligned");
  const EpiParams& ep = h.ep;
  MTP_REQUIRE(ep.mode >= MTP_EPI_BF16 && ep.mode <= MTP_EPI_BF16_PIXSHUF, "mtp_gemm_bf16: bad epilogue mode %d", ep.mode);
  MTP_REQUIRE(ep.ldo % 8 == 0 && ep.ldo > 0, "mtp_gemm_bf16: ldo=%d must be a positive multiple of 8", ep.ldo);
  if (ep.mode == MTP_EPI_F32_RESID || ep.mode == MTP_EPI_F32_POS || ep.mode == MTP_EPI_BF16_DGELU)
    MTP_REQUIRE(ep.aux != nullptr, "mtp_gemm_bf16: epilogue mode %d needs aux", ep.mode);
  if (ep.mode == MTP_EPI_F32_RESID && ep.row_scale) MTP_REQUIRE(ep.rows_per_group > 0, "mtp_gemm_bf16: rows_per_group");
  if (ep.mode == MTP_EPI_F32_POS) MTP_REQUIRE(ep.pos_rows > 0, "mtp_gemm_bf16: pos_rows");
  if (ep.colsum != nullptr)
    MTP_REQUIRE((ep.mode == MTP_EPI_BF16 || ep.mode == MTP_EPI_BF16_DGELU) && ((uintptr_t)ep.colsum & 15) == 0,
                "mtp_gemm_bf16: colsum needs a BF16 / BF16_DGELU epilogue and a 16-byte aligned pointer");
  if (ep.sumsq != nullptr) MTP_REQUIRE(ep.mode == MTP_EPI_F32, "mtp_gemm_bf16: sumsq needs the F32 epilogue");
  if (ep.hilo) {
    MTP_REQUIRE(!h.a_mn && !h.b_mn && h.K % BK == 0 && h.lda >= 2 * h.K && h.ldb >= 2 * h.K,
                "mtp_gemm_bf16: hilo mode needs K-major operands stored as [rows, 2K] (hi | lo) and K %% 64 == 0 (K=%d)", h.K);
    MTP_REQUIRE(ep.mode != MTP_EPI_BF16_PIXSHUF && ep.mode != MTP_EPI_BF16_DGELU && ep.out2 == nullptr, "mtp_gemm_bf16: hilo mode is forward-only");
  }
  MTP_REQUIRE(ep.out_lo >= 0 && ep.out_lo % 8 == 0, "mtp_gemm_bf16: out_lo_offset must be a non-negative multiple of 8");
  if (ep.mode == MTP_EPI_BF16_PIXSHUF)
    MTP_REQUIRE(ep.ps_h > 0 && ep.ps_w > 0 && ep.ps_cout > 0 && ep.ps_cout % 32 == 0 && h.N == 4 * ep.ps_cout,
                "mtp_gemm_bf16: bad pixel-shuffle geometry");
  return MTP_OK;
}

static EpiParams to_epi(const mtp_epilogue* ep) {
  EpiParams p;
  p.mode = ep->mode; p.ldo = ep->ldo; p.bias = ep->bias; p.out = ep->out; p.out2 = ep->out2; p.aux = ep->aux;
  p.row_scale = ep->row_scale; p.rows_per_group = ep->rows_per_group; p.pos_rows = ep->pos_rows;
  p.accumulate = ep->accumulate; p.ps_h = ep->ps_h; p.ps_w = ep->ps_w; p.ps_cout = ep->ps_cout;
  p.colsum = ep->colsum;
  p.b_static = ep->b_static;
  p.sumsq = ep->sumsq;
  p.hilo = ep->hilo;
  p.out_lo = ep->out_lo_offset;
  return p;
}

static int g_gemm_variant = -1;      // 1: persistent kernel, 2: one tile per CTA with two CTAs per SM; -1: read MTP_GEMM_VARIANT once
static int gemm_variant() {
  if (g_gemm_variant < 0) {
    const char* e = getenv("MTP_GEMM_VARIANT");
    g_gemm_variant = (e != nullptr && (e[0] == '2' || e[0] == '3')) ? e[0] - '0' : 1;
  }
  return g_gemm_variant;
}

template <int BN>
static int launch_gemm2(const HostProblem* pr, int np, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  GemmProblem gp[2];
  memset(gp, 0, sizeof(gp));
  int items[2] = {0, 0};
  for (int p = 0; p < np; ++p) {
    const HostProblem& h = pr[p];
    int rc = h.a_mn ? make_tmap(&gp[p].tmA, h.A, h.K, h.M, h.lda, BK) : make_tmap(&gp[p].tmA, h.A, h.M, h.K, h.lda, BM);
    if (rc) return rc;
    rc = h.b_mn ? make_tmap(&gp[p].tmB, h.B, h.K, h.N, h.ldb, BK) : make_tmap(&gp[p].tmB, h.B, h.N, h.K, h.ldb, BN);
    if (rc) return rc;
    gp[p].ep = h.ep;
    gp[p].M = h.M; gp[p].N = h.N; gp[p].K = h.K;
    gp[p].tiles_m = ceil_div(h.M, BM); gp[p].tiles_n = ceil_div(h.N, BN);
    gp[p].a_mn = h.a_mn; gp[p].b_mn = h.b_mn;
    items[p] = gp[p].tiles_m * gp[p].tiles_n;
  }
  static bool attr_set_dev[64] = {};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool& attr_set = attr_set_dev[dev_ & 63];
  auto kern = gemm2_bf16_kernel<BN>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "cudaFuncSetAttribute(gemm2 BN=%d): %s", BN, cudaGetErrorString(e));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.gridDim = dim3(items[0] + items[1]);
  if (pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
  }
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, gp[0], gp[1], items[0], g_gemm_dbg_mode);
  if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "gemm2_bf16_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("gemm2_bf16_kernel");
}

// tile width of variant 2: per-SM work when the hardware deals the tiles to 2 slots per SM (both slots share one tensor core)
static int gemm2_pick_bn(const HostProblem* pr, int np, int force_bn) {
  if (force_bn) return force_bn % 1000;
  const int cand[3] = {256, 192, 128};
  int best = 256;
  double best_c = 1e300;
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    double c = 0.0;
    long long tiles = 0;
    for (int p = 0; p < np; ++p) {
      const long long t = (long long)ceil_div(pr[p].M, BM) * ceil_div(pr[p].N, bn);
      tiles += t;
      c += (double)t * (ceil_div(pr[p].K, BK) * kblock_cycles(bn, false) + kTileFixedCycles);
    }
    const int P = num_sms();
    const double per_sm = std::ceil((double)tiles / P) / ((double)tiles / P);      // quantisation of whole tiles per SM
    c = c / P * per_sm;
    if (c < best_c) { best_c = c; best = bn; }
  }
  return best;
}

static int run_grouped(const HostProblem* pr, int np, int force_bn, cudaStream_t stream) {
  for (int p = 0; p < np; ++p) {
    int rc = validate_problem(pr[p]);
    if (rc) return rc;
    MTP_REQUIRE(!pr[p].ep.hilo || np == 1, "mtp_gemm_bf16_dual: hilo mode is not available in grouped launches");
  }
  // variant 3: the co-resident kernel for single launches only (measured r2: single forward launches 0.4-3.4 us faster, the grouped
  // dgrad+wgrad launches 9-12 us slower than with the persistent kernel's LPT schedule)
  if ((gemm_variant() == 2 || (gemm_variant() == 3 && np == 1)) && force_bn < 1000 && !pr[0].ep.hilo && (force_bn == 0 || force_bn >= 128)) {
    const int bn = gemm2_pick_bn(pr, np, force_bn);
    g_last_config = 3000 + bn;
    switch (bn) {
      case 128: return launch_gemm2<128>(pr, np, stream);
      case 192: return launch_gemm2<192>(pr, np, stream);
      case 256: return launch_gemm2<256>(pr, np, stream);
      default: break;
    }
  }
  if (force_bn >= 1000) {
    for (int p = 0; p < np; ++p)
      MTP_REQUIRE(!pr[p].b_mn || (force_bn % 1000) % 128 == 0, "mtp_gemm_bf16: paired CTAs with MN-major B need a tile width of 128 or 256");
  }
  const Config* cfg = get_config(pr, np, force_bn);
  if (cfg == nullptr) return set_error(MTP_ERR_INVALID, "mtp_gemm_bf16: no valid tile configuration (force_bn=%d)", force_bn);
  g_last_config = cfg->bn + (cfg->cl2 ? 1000 : 0);
#define MTP_LAUNCH(BN_)                                                             \
  case BN_:                                                                         \
    return cfg->cl2 ? launch_grouped<BN_, true>(pr, np, cfg->sched, stream) : launch_grouped<BN_, false>(pr, np, cfg->sched, stream);
  switch (cfg->bn) {
    MTP_LAUNCH(64)
    MTP_LAUNCH(128)
    MTP_LAUNCH(192)
    MTP_LAUNCH(256)
    default: break;
  }
#undef MTP_LAUNCH
  return set_error(MTP_ERR_INVALID, "mtp_gemm_bf16: unsupported tile width %d", cfg->bn);
}

// Host-only planning (no CUDA call): the tile configuration the heuristic picks and a self-check of the work schedule --
// every tile of every problem is assigned to exactly one slot, no slot exceeds the by-value list.
static int plan_grouped(const HostProblem* pr, int np, int force_bn, int* out_config, int* out_ctas, double* out_cycles) {
  for (int p = 0; p < np; ++p) {
    int rc = validate_problem(pr[p]);
    if (rc) return rc;
  }
  const Config* cfg = get_config(pr, np, force_bn);
  if (cfg == nullptr) return set_error(MTP_ERR_INVALID, "mtp_gemm_plan: no valid tile configuration (force_bn=%d)", force_bn);
  const int slots = cfg->cl2 ? num_sms() / 2 : num_sms();
  int total = 0;
  for (int p = 0; p < np; ++p) {
    const int tiles_m = ceil_div(pr[p].M, BM), tiles_n = ceil_div(pr[p].N, cfg->bn);
    total += (cfg->cl2 ? (tiles_m + 1) / 2 : tiles_m) * tiles_n;
  }
  std::vector<int> seen(total, 0);
  int used = 0;
  const Sched& sc = cfg->sched;
  if (sc.strided_total > 0) {
    MTP_REQUIRE(sc.strided_total == total, "mtp_gemm_plan: strided schedule covers %d of %d items", sc.strided_total, total);
    for (int s = 0; s < slots; ++s)
      for (int i = s; i < total; i += slots) ++seen[i];
    used = std::min(slots, total);
  } else {
    for (int s = 0; s < slots; ++s) {
      MTP_REQUIRE(sc.count[s] <= MAX_ITEMS, "mtp_gemm_plan: slot %d holds %d items", s, (int)sc.count[s]);
      for (int i = 0; i < sc.count[s]; ++i) {
        MTP_REQUIRE(sc.item[s][i] < total, "mtp_gemm_plan: item id %d out of range", (int)sc.item[s][i]);
        ++seen[sc.item[s][i]];
      }
      if (sc.count[s] > 0) used = s + 1;
    }
  }
  for (int i = 0; i < total; ++i) MTP_REQUIRE(seen[i] == 1, "mtp_gemm_plan: item %d scheduled %d times", i, seen[i]);
  if (out_config) *out_config = cfg->bn + (cfg->cl2 ? 1000 : 0);
  if (out_ctas) *out_ctas = cfg->cl2 ? 2 * used : used;
  if (out_cycles) *out_cycles = build_schedule(pr, np, cfg->bn, cfg->cl2, nullptr) + (cfg->cl2 ? kClusterLaunchCycles : 0.0);
  return MTP_OK;
}

}  // namespace mtp

using namespace mtp;

static EpiParams plan_epi(int n) {      // a valid epilogue for planning (never dereferenced)
  EpiParams e;
  memset(&e, 0, sizeof(e));
  e.mode = MTP_EPI_BF16;
  e.hilo = 0;
  e.out_lo = 0;
  e.ldo = (n + 7) / 8 * 8;
  e.out = reinterpret_cast<void*>(uintptr_t(16));
  return e;
}

extern "C" int mtp_gemm_plan(int M0, int N0, int K0, int b0_mn_major, int M1, int N1, int K1, int b1_mn_major, int force_bn,
                             int* out_config, int* out_ctas, double* out_cycles) {
  void* dummy = reinterpret_cast<void*>(uintptr_t(16));
  HostProblem h[2] = {{dummy, 8, 0, dummy, 8, b0_mn_major, M0, N0, K0, plan_epi(N0)}, {dummy, 8, 0, dummy, 8, b1_mn_major, M1, N1, K1, plan_epi(N1)}};
  return plan_grouped(h, M1 > 0 ? 2 : 1, force_bn, out_config, out_ctas, out_cycles);
}

extern "C" int mtp_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M, int N,
                             int K, const mtp_epilogue* ep, int force_bn, mtp_stream_t stream_) {
  MTP_REQUIRE(ep != nullptr, "mtp_gemm_bf16: null epilogue");
  HostProblem h = {A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, to_epi(ep)};
  return run_grouped(&h, 1, force_bn, reinterpret_cast<cudaStream_t>(stream_));
}

/* tuning aid: when set, every GEMM launch writes per-CTA globaltimer stamps [grid][8]: 0 start, 1 prologue done, 2 first operands
 * landed, 3 first item's MMAs issued, 4 last item's MMAs issued, 5 last accumulator complete (epilogue starts), 6 CTA done */
extern "C" int mtp_gemm_set_debug(void* device_buffer) {
  g_gemm_dbg = reinterpret_cast<long long*>(device_buffer);
  return MTP_OK;
}
extern "C" int mtp_gemm_last_config(void) { return g_last_config; }
/* 1: persistent warp-specialised kernel (default); 2: one tile per CTA, two CTAs per SM (the next launch's CTAs become resident while
 * this launch's last epilogues run).  mtp_gemm_last_config() reports 3000 + BN for variant 2. */
extern "C" int mtp_gemm_set_variant(int v) {
  g_gemm_variant = (v == 2 || v == 3) ? v : 1;
  return MTP_OK;
}
/* tuning aid: cap the depth of the operand ring (0 = fill the smem budget).  A shallow ring leaves room for the next launch's CTA on the
 * same SM (co-residency, at the price of less latency cover in the mainloop). */
extern "C" int mtp_gemm_set_max_stages(int n) {
  g_gemm_max_stages = n > 0 ? n : 0;
  return MTP_OK;
}
extern "C" int mtp_gemm_set_debug_mode(int mode) {
  g_gemm_dbg_mode = mode;
  return MTP_OK;
}

extern "C" int mtp_gemm_bf16_dual(const mtp_gemm_desc* g0, const mtp_gemm_desc* g1, int force_bn, mtp_stream_t stream_) {
  MTP_REQUIRE(g0 && g1 && g0->ep && g1->ep, "mtp_gemm_bf16_dual: null descriptor");
  HostProblem h[2] = {{g0->A, g0->lda, g0->a_mn_major, g0->B, g0->ldb, g0->b_mn_major, g0->M, g0->N, g0->K, to_epi(g0->ep)},
                      {g1->A, g1->lda, g1->a_mn_major, g1->B, g1->ldb, g1->b_mn_major, g1->M, g1->N, g1->K, to_epi(g1->ep)}};
  return run_grouped(h, 2, force_bn, reinterpret_cast<cudaStream_t>(stream_));
}
