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
};

// ---- epilogue for one thread: 32 consecutive columns [n, n+32) of row m -------------------------------------------
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// aux operand of one 32-column chunk, fetched BEFORE waiting on the TMEM load so the two latencies overlap
struct AuxRegs { float4 f[8]; };

__device__ __forceinline__ void load_aux(const EpiParams& ep, AuxRegs& a, int m, int n, int N) {
  if (ep.mode == MTP_EPI_F32_RESID) {
    const float* r = reinterpret_cast<const float*>(ep.aux) + (size_t)m * ep.ldo + n;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (n + 4 * j < N) a.f[j] = *reinterpret_cast<const float4*>(r + 4 * j);
  } else if (ep.mode == MTP_EPI_F32_POS) {
    const float* r = reinterpret_cast<const float*>(ep.aux) + (size_t)(m % ep.pos_rows) * N + n;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (n + 4 * j < N) a.f[j] = __ldg(reinterpret_cast<const float4*>(r + 4 * j));
  } else if (ep.mode == MTP_EPI_BF16_DGELU) {
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(ep.aux) + (size_t)m * ep.ldo + n;
#pragma unroll
    for (int j = 0; j < 4; ++j) if (n + 8 * j < N) a.f[j] = *reinterpret_cast<const float4*>(h + 8 * j);
  } else if (ep.mode == MTP_EPI_F32 && ep.accumulate) {
    const float* o = reinterpret_cast<const float*>(ep.out) + (size_t)m * ep.ldo + n;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (n + 4 * j < N) a.f[j] = *reinterpret_cast<const float4*>(o + 4 * j);
  }
}

__device__ __forceinline__ void epilogue_chunk(const EpiParams& ep, float (&v)[32], const AuxRegs& a, const float* bias_s, int m, int n, int N) {
  if (bias_s != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bias_s + j);      // smem broadcast; columns >= N hold zeros
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  switch (ep.mode) {
    case MTP_EPI_BF16: {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)m * ep.ldo + n;
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        if (n + j < N) store_bf16x8(o + j, v + j);
    } break;
    case MTP_EPI_BF16_GELU: {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)m * ep.ldo + n;
      if (ep.out2 != nullptr) {
        __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(ep.out2) + (size_t)m * ep.ldo + n;
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          if (n + j < N) store_bf16x8(o2 + j, v + j);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        if (n + j < N) store_bf16x8(o + j, v + j);
    } break;
    case MTP_EPI_F32_RESID: {
      const float s = ep.row_scale ? __ldg(ep.row_scale + m / ep.rows_per_group) : 1.0f;
      float* o = reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldo + n;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (n + j < N) {
          float4 x = a.f[j >> 2];
          x.x += s * v[j]; x.y += s * v[j + 1]; x.z += s * v[j + 2]; x.w += s * v[j + 3];
          *reinterpret_cast<float4*>(o + j) = x;
        }
      }
    } break;
    case MTP_EPI_F32_POS: {
      float* o = reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldo + n;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (n + j < N) {
          float4 x = a.f[j >> 2];
          x.x += v[j]; x.y += v[j + 1]; x.z += v[j + 2]; x.w += v[j + 3];
          *reinterpret_cast<float4*>(o + j) = x;
        }
      }
    } break;
    case MTP_EPI_F32: {
      float* o = reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldo + n;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (n + j < N) {
          float4 x = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (ep.accumulate) {
            const float4 y = a.f[j >> 2];
            x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
          }
          *reinterpret_cast<float4*>(o + j) = x;
        }
      }
    } break;
    case MTP_EPI_BF16_DGELU: {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)m * ep.ldo + n;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (n + j < N) {
          const float4 hv = a.f[j >> 3];
          const uint32_t w[4] = {__float_as_uint(hv.x), __float_as_uint(hv.y), __float_as_uint(hv.z), __float_as_uint(hv.w)};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack_bf16x2(w[t]);
            v[j + 2 * t] *= gelu_erf_grad(f.x);
            v[j + 2 * t + 1] *= gelu_erf_grad(f.y);
          }
          store_bf16x8(o + j, v + j);
        }
      }
    } break;
    case MTP_EPI_BF16_PIXSHUF: {
      const int g = n / ep.ps_cout, co = n % ep.ps_cout;
      const int dy = g >> 1, dx = g & 1;
      const int hw = ep.ps_h * ep.ps_w;
      const int b = m / hw, rem = m % hw;
      const int y = rem / ep.ps_w, x = rem % ep.ps_w;
      const size_t row = ((size_t)b * 2 * ep.ps_h + 2 * y + dy) * (2 * ep.ps_w) + 2 * x + dx;
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + row * ep.ldo + co;
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        if (n + j < N) store_bf16x8(o + j, v + j);
    } break;
    default: break;
  }
}

// -------------------------------------------------------------------------------------------------------------------
// CL2 (cta_group::2): the two CTAs of a cluster compute one 256 x BN tile with a single MMA stream issued by the leader
// (even rank).  Each CTA stages its 128 rows of A and HALF of B (BN/2 columns) and accumulates its 128 rows in its own TMEM;
// the tensor cores read the other half of B from the peer's shared memory.  This cuts the bytes each SM must pull in per MMA
// by a third, which is what bounds the 1-CTA kernel (measured: ~70 B/clk/SM of operand ingress regardless of L2 multicast).
template <int BN, bool A_MN, bool B_MN, bool CL2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiParams ep,
                 const int M, const int N, const int K, const int tiles_m, const int tiles_n) {
  using Cfg = GemmCfg<BN, CL2>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);   // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k_blocks = (K + BK - 1) / BK;
  // work items: single tiles, or (CL2) pairs of m-tiles handled by the two CTAs of a cluster in lockstep
  const int crank = CL2 ? (int)cluster_ctarank() : 0;
  const int m_groups = CL2 ? (tiles_m + 1) / 2 : tiles_m;
  const int num_items = m_groups * tiles_n;
  const int item0 = CL2 ? blockIdx.x / 2 : blockIdx.x;
  const int item_stride = CL2 ? gridDim.x / 2 : gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
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
  if (CL2) cluster_sync_all();       // peer barriers are initialised before any multicast can arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = item0; item < num_items; item += item_stride) {
        const int m0 = ((item % m_groups) * (CL2 ? 2 : 1) + crank) * BM;
        const int n0 = (item / m_groups) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (!CL2) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if (!A_MN) {
              tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m0 + j * 64, kb * BK);
            }
          } else {
            // both CTAs' bytes are credited to the LEADER's full barrier (only the leader waits on it and issues the MMAs)
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            if (!A_MN) {
              tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d_2sm(sa + j * 8192, &tmA, &full_bar[stage], m0 + j * 64, kb * BK);
            }
          }
          if (!CL2) {
            if (!B_MN) {
              tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, kb * BK);
            }
          } else {                       // my half of the pair's B tile (columns [n0 + crank*BN/2, +BN/2))
            if (!B_MN) {
              tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * BK, n0 + crank * (BN / 2));
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) tma_load_2d_2sm(sb + j * 8192, &tmB, &full_bar[stage], n0 + crank * (BN / 2) + j * 64, kb * BK);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && (!CL2 || crank == 0)) {
      constexpr uint32_t idesc = make_idesc_bf16(CL2 ? 2 * BM : BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = item0; item < num_items; item += item_stride) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          // K-major: 16 bf16 = 32 B inside the 128 B swizzle row; 8-row groups are 1024 B apart (SBO).
          // MN-major: 16 k-rows = 2 swizzle atoms of 8 rows x 128 B = 2048 B; 64-wide MN atoms are 8192 B apart (LBO).
          // Descriptors are built once per stage; a k-step only bumps the 14-bit start-address field (units of 16 B).
          const uint64_t a_desc0 = A_MN ? make_smem_desc(sa, 8192, 1024) : make_smem_desc(sa, 16, 1024);
          const uint64_t b_desc0 = B_MN ? make_smem_desc(sb, 8192, 1024) : make_smem_desc(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t a_desc = a_desc0 + (uint64_t)(k * (A_MN ? 2048 : 32) >> 4);
            const uint64_t b_desc = b_desc0 + (uint64_t)(k * (B_MN ? 2048 : 32) >> 4);
            if (CL2) umma_bf16_2sm(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0);
            else umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0);
          }
          if (CL2) umma_commit_2sm_mcast(&empty_bar[stage], 0x3);   // releases the stage in BOTH CTAs
          else umma_commit(&empty_bar[stage]);                      // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (CL2) umma_commit_2sm_mcast(&tmem_full[acc], 0x3);       // accumulators complete in both CTAs' TMEM
        else umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int hsel = (warp - 2) >> 2;       // which alternate chunks this warp takes
    const int et = threadIdx.x - 64;        // 0..255 within the epilogue group
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = item0; item < num_items; item += item_stride) {
      const int m0 = ((item % m_groups) * (CL2 ? 2 : 1) + crank) * BM;
      const int n0 = (item / m_groups) * BN;
      float* bsm = bias_s + acc * BN;
      if (ep.bias != nullptr) {             // stage this tile's bias slice once (zeros beyond N)
        for (int i = et; i < BN; i += EPI_THREADS) {
          const int n = n0 + i;
          bsm[i] = n < N ? __ldg(ep.bias + (ep.mode == MTP_EPI_BF16_PIXSHUF ? n % ep.ps_cout : n)) : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < M;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const int n_chunks = (min(BN, N - n0) + 31) / 32;
      uint32_t r[32];
      int c = hsel;
      if (c < n_chunks) tmem_ld_32x32(taddr + c * 32, r);
      for (; c < n_chunks; c += 2) {
        AuxRegs aux;
        if (row_ok) load_aux(ep, aux, m, n0 + c * 32, N);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (c + 2 < n_chunks) tmem_ld_32x32(taddr + (c + 2) * 32, r);      // next chunk streams in while this one is processed
        if (row_ok) epilogue_chunk(ep, v, aux, ep.bias != nullptr ? bsm + c * 32 : nullptr, m, n0 + c * 32, N);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL2) mbar_arrive_remote(&tmem_empty[acc], 0);      // the leader's MMA thread waits for both CTAs' epilogues
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();       // no CTA leaves while its peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if (CL2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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

template <int BN, bool A_MN, bool B_MN, bool CL2>
static int launch_gemm(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const EpiParams& ep,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CL2>;
  CUtensorMap tmA, tmB;
  int rc;
  // K-major: matrix [M rows, K cols], box rows = BM.  MN-major: matrix [K rows, M cols], box = 64 k-rows x 64 cols.
  rc = A_MN ? make_tmap(&tmA, A, K, M, lda, BK) : make_tmap(&tmA, A, M, K, lda, BM);
  if (rc) return rc;
  rc = B_MN ? make_tmap(&tmB, B, K, N, ldb, BK) : make_tmap(&tmB, B, N, K, ldb, CL2 ? BN / 2 : BN);
  if (rc) return rc;
  static bool attr_set = false;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, CL2>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "cudaFuncSetAttribute(gemm BN=%d): %s", BN, cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles_m = ceil_div(M, BM), tiles_n = ceil_div(N, BN);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (CL2) {
    const int items = ((tiles_m + 1) / 2) * tiles_n;
    cfg.gridDim = dim3(2 * std::min(items, num_sms() / 2));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3(std::min(tiles_m * tiles_n, num_sms()));
  }
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, ep, M, N, K, tiles_m, tiles_n);
  if (e != cudaSuccess) return set_error(MTP_ERR_CUDA, "gemm_bf16_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("gemm_bf16_kernel");
}

// Tile width and pairing by a wave-quantisation cost model.  Per k-block a CTA needs max(MMA time, operand-ingress time):
// the MMAs take 2*BN cycles; ingress is bounded by the measured ~70 B/clk/SM, and a cta_group::2 pair halves the B bytes each
// SM stages.  Cost = waves x (per-k-block time + amortised fixed overhead).
static void pick_config(int M, int N, bool b_mn, int& bn_out, bool& cl2_out) {
  const int sms = num_sms();
  const int tiles_m = ceil_div(M, BM);
  const int cand[4] = {256, 192, 128, 64};
  double best_cost = 1e30;
  bn_out = 128;
  cl2_out = false;
  for (int cl = 0; cl < 2; ++cl) {
    if (cl == 1 && tiles_m < 2) continue;
    for (int i = 0; i < 4; ++i) {
      const int bn = cand[i];
      if (cl == 1 && b_mn && bn % 128 != 0) continue;          // MN-major B is fetched in 64-column boxes: need an even count
      const long items = (long)(cl ? (tiles_m + 1) / 2 : tiles_m) * ceil_div(N, bn);
      const long slots = cl ? sms / 2 : sms;
      const long waves = (items + slots - 1) / slots;
      const double mma = 2.0 * bn;                               // cycles per 64-deep k-block
      const double fill = (16384.0 + (cl ? 64.0 : 128.0) * bn) / 70.0;
      const double cost = (double)waves * (std::max(mma, fill) + 90.0);     // +90: amortised prologue / epilogue tail per k-block scale
      if (cost < best_cost) { best_cost = cost; bn_out = bn; cl2_out = cl == 1; }
    }
  }
}

}  // namespace mtp

using namespace mtp;

#define DISPATCH_LAYOUT(BN_, CL_)                                                                                   \
  do {                                                                                                              \
    if (!a_mn_major && !b_mn_major) return launch_gemm<BN_, false, false, CL_>(A, lda, B, ldb, M, N, K, p, stream); \
    if (!a_mn_major && b_mn_major) return launch_gemm<BN_, false, true, CL_>(A, lda, B, ldb, M, N, K, p, stream);   \
    if (a_mn_major && b_mn_major) return launch_gemm<BN_, true, true, CL_>(A, lda, B, ldb, M, N, K, p, stream);     \
    return launch_gemm<BN_, true, false, CL_>(A, lda, B, ldb, M, N, K, p, stream);                                  \
  } while (0)

extern "C" int mtp_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M, int N,
                             int K, const mtp_epilogue* ep, int force_bn, mtp_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTP_REQUIRE(A && B && ep && ep->out, "mtp_gemm_bf16: null pointer");
  MTP_REQUIRE(M > 0 && N > 0 && K > 0, "mtp_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  MTP_REQUIRE(N % 8 == 0, "mtp_gemm_bf16: N=%d must be a multiple of 8", N);
  MTP_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "mtp_gemm_bf16: lda/ldb must be multiples of 8 (got %d, %d)", lda, ldb);
  MTP_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)ep->out & 15) == 0,
              "mtp_gemm_bf16: pointers must be 16-byte aligned");
  MTP_REQUIRE(ep->mode >= MTP_EPI_BF16 && ep->mode <= MTP_EPI_BF16_PIXSHUF, "mtp_gemm_bf16: bad epilogue mode %d", ep->mode);
  MTP_REQUIRE(ep->ldo % 8 == 0 && ep->ldo > 0, "mtp_gemm_bf16: ldo=%d must be a positive multiple of 8", ep->ldo);
  if (ep->mode == MTP_EPI_F32_RESID || ep->mode == MTP_EPI_F32_POS || ep->mode == MTP_EPI_BF16_DGELU)
    MTP_REQUIRE(ep->aux != nullptr, "mtp_gemm_bf16: epilogue mode %d needs aux", ep->mode);
  if (ep->mode == MTP_EPI_F32_RESID && ep->row_scale) MTP_REQUIRE(ep->rows_per_group > 0, "mtp_gemm_bf16: rows_per_group");
  if (ep->mode == MTP_EPI_F32_POS) MTP_REQUIRE(ep->pos_rows > 0, "mtp_gemm_bf16: pos_rows");
  if (ep->mode == MTP_EPI_BF16_PIXSHUF)
    MTP_REQUIRE(ep->ps_h > 0 && ep->ps_w > 0 && ep->ps_cout > 0 && ep->ps_cout % 32 == 0 && N == 4 * ep->ps_cout,
                "mtp_gemm_bf16: bad pixel-shuffle geometry");
  EpiParams p;
  p.mode = ep->mode; p.ldo = ep->ldo; p.bias = ep->bias; p.out = ep->out; p.out2 = ep->out2; p.aux = ep->aux;
  p.row_scale = ep->row_scale; p.rows_per_group = ep->rows_per_group; p.pos_rows = ep->pos_rows;
  p.accumulate = ep->accumulate; p.ps_h = ep->ps_h; p.ps_w = ep->ps_w; p.ps_cout = ep->ps_cout;
  // force_bn: 0 = heuristic; otherwise tile width, +1000 to force the 2-CTA multicast cluster variant (tests / tuning)
  int bn;
  bool cl2;
  if (force_bn == 0) {
    pick_config(M, N, b_mn_major != 0, bn, cl2);
  } else {
    cl2 = force_bn >= 1000;
    bn = force_bn % 1000;
    MTP_REQUIRE(!cl2 || !b_mn_major || bn % 128 == 0, "mtp_gemm_bf16: clustered MN-major B needs a tile width of 128 or 256");
  }
  if (cl2) {
    switch (bn) {
      case 64: DISPATCH_LAYOUT(64, true);
      case 128: DISPATCH_LAYOUT(128, true);
      case 192: DISPATCH_LAYOUT(192, true);
      case 256: DISPATCH_LAYOUT(256, true);
      default: break;
    }
  } else {
    switch (bn) {
      case 64: DISPATCH_LAYOUT(64, false);
      case 128: DISPATCH_LAYOUT(128, false);
      case 192: DISPATCH_LAYOUT(192, false);
      case 256: DISPATCH_LAYOUT(256, false);
      default: break;
    }
  }
  return set_error(MTP_ERR_INVALID, "mtp_gemm_bf16: unsupported tile width %d", bn);
}
