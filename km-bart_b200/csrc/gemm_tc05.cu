// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = epilogue(sum_k A(m,k) * B(n,k)).
//
// One persistent CTA per SM, 384 threads, warp-specialised:
//   warp 0      TMA producer   (one elected lane; cp.async.bulk.tensor 2D, SWIZZLE_128B)
//   warp 1      MMA issuer     (one elected lane; tcgen05.mma cta_group::1, M=128, N=BN, K=16|8)
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4..11 epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> swizzled smem
//                               transpose -> coalesced global I/O)
// Three mbarrier pipelines: smem full/empty (TMA<->MMA), tmem full/empty (MMA<->epilogue),
// and a static round-robin tile schedule.  Operands may be K-major or MN-major (the
// backward GEMMs contract over the token dimension, which is the slow dimension of
// every activation), bf16 or tf32.
//
// Replaces the cuBLAS calls the reference issues through nn.Linear / F.linear
// (SURVEY.md §2.3(b) K2,K5,K7,K8,K10) — see include/kmbart.h for the call-site map.
#include <cuda.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

constexpr int BM = 128;
constexpr int TILE_BYTES_ROW = 128;  // one swizzle span: 64 bf16 or 32 tf32
constexpr int A_TILE_BYTES = BM * TILE_BYTES_ROW;

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, k_blocks;
  KmbGemmEpilogue e;
  int vec_ok;  // all leading dims / pointers allow 16-byte row-segment access
  int tma_out; // bf16 outputs leave through smem + TMA bulk stores (tmOut / tmPre valid)
  int split_k, kb_per_split;  // split-K work items; partial sums are reduced with fp32 red.global.add
  uint32_t drop_thresh16;
  float drop_scale;
};

template <int BN>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * TILE_BYTES_ROW;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int EPI_STAGING = 8 * 4096;  // one 32x32 fp32 transpose tile per epilogue warp
  static constexpr int STAGES_RAW = (227 * 1024 - 1024 - 256 - EPI_STAGING) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_STAGING;
};

// Instruction descriptor, field layout per cute/arch/mma_sm100_desc.hpp InstrDescriptor.
__device__ __forceinline__ uint32_t make_idesc(int elt, int a_mn, int b_mn, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                             // c_format = F32
  const uint32_t fmt = elt == 0 ? 1u : 2u;  // BF16 : TF32
  d |= fmt << 7;                            // a_format
  d |= fmt << 10;                           // b_format
  d |= (uint32_t)a_mn << 15;                // a_major (0 = K, 1 = MN)
  d |= (uint32_t)b_mn << 16;                // b_major
  d |= (uint32_t)(n >> 3) << 17;            // n_dim
  d |= (uint32_t)(BM >> 4) << 24;           // m_dim
  return d;
}

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ------------------------------------------------------------------ epilogue staging
// Each epilogue warp owns a 32x32 fp32 staging tile in shared memory (4 KB, float4 slots,
// XOR-swizzled so both access patterns below are bank-conflict free).  tcgen05.ld hands a
// thread one accumulator ROW (32 consecutive columns); global memory wants a warp to touch
// one row segment per instruction.  The tile converts between the two, so every global
// load/store of the epilogue is a full 128-byte (fp32) / 64-byte (bf16) row segment.
__device__ __forceinline__ void tile_put_row(float4* st, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) st[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void tile_get_row(const float4* st, int lane, float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = st[lane * 8 + (j ^ (lane & 7))];
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
// g points at (first row of this warp's 32-row group, first column of the chunk)
__device__ __forceinline__ void tile_load_f32(float4* st, int lane, const float* g, int64_t ld, int rows_valid, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3), grp = lane & 7;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows_valid) x = __ldg(reinterpret_cast<const float4*>(g + (int64_t)row * ld + 4 * grp));
    st[row * 8 + (grp ^ (row & 7))] = x;
  }
  __syncwarp();
  tile_get_row(st, lane, v);
  __syncwarp();
}
__device__ __forceinline__ void tile_load_bf16(float4* st, int lane, const bf16* g, int64_t ld, int rows_valid, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + (lane >> 2), piece = lane & 3;
    uint4 x = make_uint4(0, 0, 0, 0);
    if (row < rows_valid) x = __ldg(reinterpret_cast<const uint4*>(g + (int64_t)row * ld + 8 * piece));
    st[row * 8 + ((2 * piece) ^ (row & 7))] = make_float4(bf16lo(x.x), bf16hi(x.x), bf16lo(x.y), bf16hi(x.y));
    st[row * 8 + ((2 * piece + 1) ^ (row & 7))] = make_float4(bf16lo(x.z), bf16hi(x.z), bf16lo(x.w), bf16hi(x.w));
  }
  __syncwarp();
  tile_get_row(st, lane, v);
  __syncwarp();
}
__device__ __forceinline__ void tile_store_f32(float4* st, int lane, float* g, int64_t ld, int rows_valid, const float (&v)[32]) {
  tile_put_row(st, lane, v);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3), grp = lane & 7;
    if (row < rows_valid) *reinterpret_cast<float4*>(g + (int64_t)row * ld + 4 * grp) = st[row * 8 + (grp ^ (row & 7))];
  }
  __syncwarp();
}
__device__ __forceinline__ void tile_store_bf16(float4* st, int lane, bf16* g, int64_t ld, int rows_valid, const float (&v)[32]) {
  tile_put_row(st, lane, v);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + (lane >> 2), piece = lane & 3;
    if (row < rows_valid) {
      const float4 a = st[row * 8 + ((2 * piece) ^ (row & 7))];
      const float4 c = st[row * 8 + ((2 * piece + 1) ^ (row & 7))];
      *reinterpret_cast<uint4*>(g + (int64_t)row * ld + 8 * piece) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------ epilogue: one 32x32 chunk per warp
// v: this thread's accumulator row (row = row0 + lane), columns [col0, col0+32).
// Warp-uniform: col0, ncols, row0.  `full` = chunk fully inside N and 16-byte vector access legal.
__device__ __forceinline__ void epilogue_linear(const GemmParams& p, float4* st, int lane, float (&v)[32], int row0,
                                                int col0, int ncols, uint32_t drop_key) {
  const KmbGemmEpilogue& e = p.e;
  const bool full = (ncols == 32) && p.vec_ok;
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  int rows_valid = p.M - row0;
  rows_valid = rows_valid > 32 ? 32 : rows_valid;
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
  if (e.bias) {
    if (full) {
      const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 t = __ldg(b4 + j);
        v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) v[j] += __ldg(e.bias + col0 + j);
    }
  }
  if (e.act == KMB_ACT_GELU) {
    if (e.out_preact) {
      bf16* pp = reinterpret_cast<bf16*>(e.out_preact) + (int64_t)row0 * e.ld_bf16 + col0;
      if (full) {
        tile_store_bf16(st, lane, pp, e.ld_bf16, rows_valid, v);
      } else if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) pp[(int64_t)lane * e.ld_bf16 + j] = __float2bfloat16(v[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (e.act == KMB_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
  } else if (e.act == KMB_ACT_GELU_GRAD || e.act == KMB_ACT_TANH_GRAD) {
    const bf16* ap = reinterpret_cast<const bf16*>(e.aux) + (int64_t)row0 * e.ld_aux + col0;
    float a[32];
    if (full) {
      tile_load_bf16(st, lane, ap, e.ld_aux, rows_valid, a);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = (row_ok && j < ncols) ? __bfloat162float(ap[(int64_t)lane * e.ld_aux + j]) : 0.f;
    }
    if (e.act == KMB_ACT_GELU_GRAD) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= gelu_erf_grad(a[j]);
    } else {  // aux holds tanh output y: d/dx = 1 - y^2
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= (1.f - a[j] * a[j]);
    }
  }
  if (p.drop_thresh16) {
    const uint64_t base = (uint64_t)row * (uint64_t)p.N + (uint64_t)col0;  // col0 % 32 == 0, N % 4 == 0
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint64_t bits = dropout_bits4_k(drop_key, (base >> 2) + j);
#pragma unroll
      for (int l = 0; l < 4; ++l)
        v[4 * j + l] = dropout_keep(bits, l, p.drop_thresh16) ? v[4 * j + l] * p.drop_scale : 0.f;
    }
  }
  if (e.residual) {
    const float* rp = e.residual + (int64_t)row0 * e.ld_res + col0;
    if (full) {
      float r[32];
      tile_load_f32(st, lane, rp, e.ld_res, rows_valid, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += r[j];
    } else if (row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) v[j] += rp[(int64_t)lane * e.ld_res + j];
    }
  }
  if (e.out_f32 && p.split_k > 1) {
    // split-K partial: reduce into the (pre-zeroed or accumulating) fp32 output with vector reds
    float* op = e.out_f32 + (int64_t)row0 * e.ld_f32 + col0;
    if (full) {
      tile_put_row(st, lane, v);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3), grp = lane & 7;
        if (rr < rows_valid) {
          const float4 x = st[rr * 8 + (grp ^ (rr & 7))];
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + (int64_t)rr * e.ld_f32 + 4 * grp), "f"(x.x),
                       "f"(x.y), "f"(x.z), "f"(x.w)
                       : "memory");
        }
      }
      __syncwarp();
    } else if (row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) atomicAdd(op + (int64_t)lane * e.ld_f32 + j, v[j]);
    }
  } else if (e.out_f32) {
    float* op = e.out_f32 + (int64_t)row0 * e.ld_f32 + col0;
    if (full) {
      if (e.accumulate) {
        float r[32];
        tile_load_f32(st, lane, op, e.ld_f32, rows_valid, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += r[j];
      }
      tile_store_f32(st, lane, op, e.ld_f32, rows_valid, v);
    } else if (row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) {
          float* q = op + (int64_t)lane * e.ld_f32 + j;
          if (e.accumulate) v[j] += *q;
          *q = v[j];
        }
    }
  }
  if (e.out_bf16) {
    bf16* op = reinterpret_cast<bf16*>(e.out_bf16) + (int64_t)row0 * e.ld_bf16 + col0;
    if (full) {
      tile_store_bf16(st, lane, op, e.ld_bf16, rows_valid, v);
    } else if (row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) op[(int64_t)lane * e.ld_bf16 + j] = __float2bfloat16(v[j]);
    }
  }
}

// ------------------------------------------------------------------ epilogue fast path: bf16 via TMA store
// 32 rows x 64 columns (two tcgen05.ld chunks) are packed to bf16 and written into the warp's 4 KB
// staging tile in the SWIZZLE_128B pattern (16-byte piece p of row r at r*128 + ((p ^ (r & 7)) << 4)),
// which is both bank-conflict free for thread-per-row stores and the layout a 128B-swizzled tensor
// map expects; one elected lane then issues an asynchronous bulk store.  Rows / columns outside
// [M, N] are clipped by the TMA unit.
__device__ __forceinline__ void stage_bf16_pair_and_store(float4* st, int lane, const float (&v)[64], const void* tmap,
                                                          int col0, int row0) {
  if (lane == 0) tma_store_wait_read();  // previous bulk store has drained the tile
  __syncwarp();
  uint4* sp = reinterpret_cast<uint4*>(st);
#pragma unroll
  for (int pc = 0; pc < 8; ++pc)
    sp[lane * 8 + (pc ^ (lane & 7))] = make_uint4(pack_bf16(v[8 * pc], v[8 * pc + 1]), pack_bf16(v[8 * pc + 2], v[8 * pc + 3]),
                                                  pack_bf16(v[8 * pc + 4], v[8 * pc + 5]), pack_bf16(v[8 * pc + 6], v[8 * pc + 7]));
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmap, st, col0, row0);
    tma_store_commit();
  }
}

template <int BN, int ELT, int A_MN, int B_MN>
__global__ void __launch_bounds__(384, 1)
gemm_tc05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmPre,
                 const GemmParams p) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int ELT_BYTES = ELT == 0 ? 2 : 4;
  constexpr int BK = TILE_BYTES_ROW / ELT_BYTES;  // 64 bf16 / 32 tf32 per k-block
  constexpr int UK = 32 / ELT_BYTES;              // UMMA K: 16 bf16 / 8 tf32
  constexpr int KSTEPS = BK / UK;                 // 4
  constexpr int MN_CHUNK = BK;                    // elements per 128-byte row of an MN-major tile
  static_assert(!B_MN || BN >= MN_CHUNK, "MN-major B needs BN >= one 128-byte chunk");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // [pipeline stages | 8 x 4 KB epilogue staging tiles (1024-byte aligned: TMA swizzle is address based) | barriers]
  float4* stage_tiles = reinterpret_cast<float4*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + C::EPI_STAGING);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 256);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles * p.split_k;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m0 = (t % p.m_tiles) * BM;
        const int n0 = ((t / p.m_tiles) % p.n_tiles) * BN;
        const int kb0 = (t / (p.m_tiles * p.n_tiles)) * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          const int k0 = kb * BK;
          if (A_MN == 0) {
            tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / MN_CHUNK; ++c)
              tma_load_2d(sa + c * BK * TILE_BYTES_ROW, &tmA, &full_bar[stage], m0 + c * MN_CHUNK, k0);
          }
          if (B_MN == 0) {
            tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / MN_CHUNK; ++c)
              tma_load_2d(sb + c * BK * TILE_BYTES_ROW, &tmB, &full_bar[stage], n0 + c * MN_CHUNK, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(ELT, A_MN, B_MN, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        const int kb_lo = (t / (p.m_tiles * p.n_tiles)) * p.kb_per_split;
        const int kb_n = min(p.k_blocks, kb_lo + p.kb_per_split) - kb_lo;
        for (int kb = 0; kb < kb_n; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk) {
            uint64_t da, db;
            if (A_MN == 0) da = make_smem_desc_sw128(sa + kk * 32, 16, 1024);
            else da = make_smem_desc_sw128(sa + kk * UK * TILE_BYTES_ROW, BK * TILE_BYTES_ROW, 1024);
            if (B_MN == 0) db = make_smem_desc_sw128(sb + kk * 32, 16, 1024);
            else db = make_smem_desc_sw128(sb + kk * UK * TILE_BYTES_ROW, BK * TILE_BYTES_ROW, 1024);
            if (ELT == 0) umma_f16(tmem_d, da, db, idesc, (kb | kk) != 0);
            else umma_tf32(tmem_d, da, db, idesc, (kb | kk) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // warp w may only touch TMEM lanes [32*(w%4), +32); warps 4..7 take even 32-column chunks,
    // warps 8..11 the odd ones.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    float4* st = stage_tiles + (warp - 4) * 256;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t drop_key = 0;
    if (p.drop_thresh16 && p.e.dropout_seed) drop_key = dropout_key(*p.e.dropout_seed, p.e.dropout_tag);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m_blk = t % p.m_tiles, n_blk = (t / p.m_tiles) % p.n_tiles;
      const int row0 = m_blk * BM + q * 32;
      const int row = row0 + lane;
      const int n0 = n_blk * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      const bool row_ok = row < p.M;
      int rows_valid = p.M - row0;
      rows_valid = rows_valid > 32 ? 32 : rows_valid;

      if (p.e.mode == KMB_EPI_LINEAR && p.tma_out && BN >= 64) {
        // ---- fast path: bias / activation -> bf16 -> TMA store, 64 columns at a time
        const KmbGemmEpilogue& e = p.e;
#pragma unroll 1
        for (int pr = half; pr < BN / 64; pr += 2) {
          uint32_t r0[32], r1[32];
          tmem_ld32(taddr + pr * 64, r0);
          tmem_ld32(taddr + pr * 64 + 32, r1);
          tmem_ld_wait();
          const int col0 = n0 + pr * 64;
          if (rows_valid > 0 && col0 < p.N) {
            float v[64];
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
            if (e.alpha != 1.0f) {
#pragma unroll
              for (int j = 0; j < 64; ++j) v[j] *= e.alpha;
            }
            if (e.bias) {
              if (col0 + 64 <= p.N) {
                const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float4 t4 = __ldg(b4 + j);
                  v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 64; ++j)
                  if (col0 + j < p.N) v[j] += __ldg(e.bias + col0 + j);
              }
            }
            if (e.act == KMB_ACT_GELU) {
              if (e.out_preact) stage_bf16_pair_and_store(st, lane, v, &tmPre, col0, row0);
#pragma unroll
              for (int j = 0; j < 64; ++j) v[j] = gelu_erf(v[j]);
            } else if (e.act == KMB_ACT_TANH) {
#pragma unroll
              for (int j = 0; j < 64; ++j) v[j] = tanhf(v[j]);
            } else if (e.act == KMB_ACT_GELU_GRAD || e.act == KMB_ACT_TANH_GRAD) {
              if (lane == 0) tma_store_wait_read();  // the staging tile doubles as the aux transpose buffer
              __syncwarp();
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                float a[32];
                const int cc = col0 + hh * 32;
                const bf16* ap = reinterpret_cast<const bf16*>(e.aux) + (int64_t)row0 * e.ld_aux + cc;
                if (cc + 32 <= p.N && p.vec_ok) {
                  tile_load_bf16(st, lane, ap, e.ld_aux, rows_valid, a);
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    a[j] = (row_ok && cc + j < p.N) ? __bfloat162float(ap[(int64_t)lane * e.ld_aux + j]) : 0.f;
                }
                if (e.act == KMB_ACT_GELU_GRAD) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[hh * 32 + j] *= gelu_erf_grad(a[j]);
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[hh * 32 + j] *= (1.f - a[j] * a[j]);
                }
              }
            }
            stage_bf16_pair_and_store(st, lane, v, &tmOut, col0, row0);
          }
        }
      } else if (p.e.mode == KMB_EPI_LINEAR) {
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          const int col0 = n0 + c * 32;
          int ncols = p.N - col0;
          ncols = ncols > 32 ? 32 : ncols;
          if (rows_valid > 0 && ncols > 0) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            epilogue_linear(p, st, lane, v, row0, col0, ncols, drop_key);
          }
        }
      } else if (p.e.mode == KMB_EPI_CE_STATS) {
        // online softmax partial over this warp's columns of the tile (+ final_logits_bias);
        // partial index = n_blk * 2 + half
        float mx = -INFINITY, sm = 0.f;
        const int64_t label = row_ok ? p.e.labels[row] : -100;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          const int col0 = n0 + c * 32;
          int ncols = p.N - col0;
          ncols = ncols > 32 ? 32 : ncols;
          if (row_ok && ncols > 0) {
            float v[32];
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = __uint_as_float(r[j]) * p.e.alpha;
              if (j < ncols) {
                if (p.e.bias) v[j] += __ldg(p.e.bias + col0 + j);
                cm = fmaxf(cm, v[j]);
              }
            }
            const float nm = fmaxf(mx, cm);
            float cs = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) cs += __expf(v[j] - nm);
            sm = sm * __expf(mx - nm) + cs;
            mx = nm;
            if (label >= col0 && label < col0 + ncols) {
              float lv = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j == (int)label) lv = v[j];
              p.e.ce_label_logit[row] = lv;
            }
          }
        }
        if (row_ok) {
          p.e.ce_max[(int64_t)row * (2 * p.n_tiles) + 2 * n_blk + half] = mx;
          p.e.ce_sum[(int64_t)row * (2 * p.n_tiles) + 2 * n_blk + half] = sm;
        }
      } else {  // KMB_EPI_CE_GRAD: dlogits = (softmax - onehot) * gscale, bf16, through the TMA store path
        const int64_t label = row_ok ? p.e.labels[row] : -100;
        const float lse = row_ok ? p.e.ce_lse[row] : 0.f;
        const float gs = (label >= 0) ? *p.e.ce_gscale : 0.f;
#pragma unroll 1
        for (int pr = half; pr < BN / 64; pr += 2) {
          uint32_t r0[32], r1[32];
          tmem_ld32(taddr + pr * 64, r0);
          tmem_ld32(taddr + pr * 64 + 32, r1);
          tmem_ld_wait();
          const int col0 = n0 + pr * 64;
          if (rows_valid > 0 && col0 < p.N) {
            float v[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              float x = __uint_as_float(j < 32 ? r0[j & 31] : r1[j & 31]) * p.e.alpha;
              if (p.e.bias && col0 + j < p.N) x += __ldg(p.e.bias + col0 + j);
              float pv = __expf(x - lse);
              if (col0 + j == (int)label) pv -= 1.f;
              v[j] = pv * gs;
            }
            stage_bf16_pair_and_store(st, lane, v, &tmOut, col0, row0);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_read();  // staging smem must outlive the last bulk store's read
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// 2-D map over a row-major [rows, ld] matrix with `cols` valid columns, box = (box_cols, box_rows)
static int make_tmap(CUtensorMap* tm, const void* ptr, int elt, uint64_t rows, uint64_t cols,
                     uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    kmb_set_last_error("cuTensorMapEncodeTiled entry point unavailable", __FILE__, __LINE__);
    return KMB_ERR_TMAP;
  }
  const uint64_t esz = elt == 0 ? 2 : 4;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, elt == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%ux%u",
             (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
    kmb_set_last_error(msg, __FILE__, __LINE__);
    return KMB_ERR_TMAP;
  }
  return KMB_OK;
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int ELT, int A_MN, int B_MN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmPre,
                  const GemmParams& p, cudaStream_t st) {
  using C = Cfg<BN>;
  auto kern = gemm_tc05_kernel<BN, ELT, A_MN, B_MN>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      kmb_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return KMB_ERR_CUDA;
    }
    attr_set = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.split_k;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, 384, C::SMEM_BYTES, st>>>(tmA, tmB, tmOut, tmPre, p);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

template <int BN, int ELT>
static int launch_major(int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                        const CUtensorMap& tmPre, const GemmParams& p, cudaStream_t st) {
  // an MN-major B tile is built from 128-byte-wide chunks, so BN must cover one chunk
  constexpr bool kBmnOk = BN >= (ELT == 0 ? 64 : 32);
  if (!a_mn && !b_mn) return launch<BN, ELT, 0, 0>(tmA, tmB, tmOut, tmPre, p, st);
  if (a_mn && !b_mn) return launch<BN, ELT, 1, 0>(tmA, tmB, tmOut, tmPre, p, st);
  if constexpr (kBmnOk) {
    if (!a_mn && b_mn) return launch<BN, ELT, 0, 1>(tmA, tmB, tmOut, tmPre, p, st);
    return launch<BN, ELT, 1, 1>(tmA, tmB, tmOut, tmPre, p, st);
  }
  kmb_set_last_error("kmb_gemm: tile_n too narrow for an MN-major B operand", __FILE__, __LINE__);
  return KMB_ERR_ARG;
}

}  // namespace kmb

extern "C" int kmb_gemm_pick_tile_n(int M, int N) {
  // Wide tiles halve the operand traffic per flop (measured: 128-wide tiles saturate near 0.75 PFLOP/s
  // on L2 bandwidth, 256-wide reach ~1.3); narrow tiles only win when they are needed to fill the SMs
  // (small-M decode problems streaming the weights).
  const int mt = (M + kmb::BM - 1) / kmb::BM;
  const int cands[4] = {256, 128, 64, 32};
  const double mainloop[4] = {1.0, 0.6, 0.4, 0.25};
  int best = 32;
  double best_score = -1.0;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (bn > 32 && bn / 2 >= N) continue;  // do not pad N by more than 2x
    const int tiles = mt * ((N + bn - 1) / bn);
    const int waves = (tiles + 147) / 148;
    const double eff = (double)tiles / (waves * 148.0);           // SM fill
    const double npad = (double)N / (((N + bn - 1) / bn) * bn);   // useful columns
    const double score = eff * npad * mainloop[i];
    if (score > best_score + 1e-9) { best_score = score; best = bn; }
  }
  return best;
}

// number of (max, sumexp) partials per row that KMB_EPI_CE_STATS writes: two per n-tile
// (one per epilogue column half)
extern "C" int kmb_gemm_n_tiles(int N, int tile_n) {
  if (tile_n != 32 && tile_n != 64 && tile_n != 128 && tile_n != 256) return KMB_ERR_ARG;
  return 2 * ((N + tile_n - 1) / tile_n);
}

extern "C" int kmb_gemm(const void* A, const void* B, int M, int N, int K, int64_t lda, int64_t ldb,
                        int a_mn, int b_mn, int elt, const KmbGemmEpilogue* epi, int tile_n,
                        kmb_stream_t stream) {
  using namespace kmb;
  if (!A || !B || !epi || M <= 0 || N <= 0 || K <= 0 || (elt != 0 && elt != 1)) {
    kmb_set_last_error("kmb_gemm: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if (elt == 1 && (a_mn || b_mn)) {
    // only the forward (parity-mode) GEMMs run in tf32 and both of their operands are K-major
    kmb_set_last_error("kmb_gemm: tf32 operands must be K-major (a_mn = b_mn = 0)", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const int esz = elt == 0 ? 2 : 4;
  if ((lda * esz) % 16 || (ldb * esz) % 16 || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) {
    kmb_set_last_error("kmb_gemm: operands must be 16-byte aligned with 16-byte row pitch", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const bool splitk_ok = epi->mode == KMB_EPI_LINEAR && epi->accumulate && epi->out_f32 && !epi->out_bf16 &&
                         !epi->residual && !epi->bias && epi->act == KMB_ACT_NONE && epi->dropout_p <= 0.f;
  if (tile_n == 0) {
    // split-K fills the SMs for weight-gradient shapes, so they can always use the widest tile
    if (splitk_ok) tile_n = N > 128 ? 256 : (N > 64 ? 128 : 64);
    else tile_n = kmb_gemm_pick_tile_n(M, N);
  }
  if (tile_n != 32 && tile_n != 64 && tile_n != 128 && tile_n != 256) {
    kmb_set_last_error("kmb_gemm: tile_n must be 32/64/128/256", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const int bk = TILE_BYTES_ROW / esz;
  if (b_mn && tile_n < bk) tile_n = bk;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + tile_n - 1) / tile_n;
  p.k_blocks = (K + bk - 1) / bk;
  p.e = *epi;
  p.drop_thresh16 = 0;
  p.drop_scale = 1.f;
  if (epi->dropout_p > 0.f) {
    if (!epi->dropout_seed) {
      kmb_set_last_error("kmb_gemm: dropout needs a device seed pointer", __FILE__, __LINE__);
      return KMB_ERR_ARG;
    }
    p.drop_thresh16 = (uint32_t)(epi->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 1.0f / (1.0f - epi->dropout_p);
  }
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  p.vec_ok = 1;
  if (epi->bias && !al16(epi->bias)) p.vec_ok = 0;
  if (epi->residual && (!al16(epi->residual) || (epi->ld_res % 4))) p.vec_ok = 0;
  if (epi->aux && (!al16(epi->aux) || (epi->ld_aux % 8))) p.vec_ok = 0;
  if (epi->out_f32 && (!al16(epi->out_f32) || (epi->ld_f32 % 4))) p.vec_ok = 0;
  if ((epi->out_bf16 || epi->out_preact) && (epi->ld_bf16 % 8)) p.vec_ok = 0;
  if (epi->out_bf16 && !al16(epi->out_bf16)) p.vec_ok = 0;
  if (epi->out_preact && !al16(epi->out_preact)) p.vec_ok = 0;
  if (epi->mode == KMB_EPI_CE_STATS && (!epi->labels || !epi->ce_max || !epi->ce_sum || !epi->ce_label_logit)) {
    kmb_set_last_error("kmb_gemm: CE_STATS needs labels/ce_max/ce_sum/ce_label_logit", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if (epi->mode == KMB_EPI_CE_GRAD && (!epi->labels || !epi->ce_lse || !epi->ce_gscale || !epi->out_bf16)) {
    kmb_set_last_error("kmb_gemm: CE_GRAD needs labels/ce_lse/ce_gscale/out_bf16", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }

  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap(&tmA, A, elt, (uint64_t)M, (uint64_t)K, (uint64_t)lda, bk, BM);
  else rc = make_tmap(&tmA, A, elt, (uint64_t)K, (uint64_t)M, (uint64_t)lda, bk, bk);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap(&tmB, B, elt, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, bk, tile_n);
  else rc = make_tmap(&tmB, B, elt, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, bk, bk);
  if (rc) return rc;

  // split-K: weight-gradient shaped problems (few output tiles, long K) that accumulate into fp32
  p.split_k = 1;
  p.kb_per_split = p.k_blocks;
  if (splitk_ok) {
    const int tiles = p.m_tiles * p.n_tiles;
    const int sms = num_sms();
    int best = 1;
    double best_eff = 0.0;
    for (int sk = 1; sk <= 16; ++sk) {
      if (p.k_blocks / sk < 8 && sk > 1) break;
      const int items = tiles * sk;
      const double eff = (double)items / (((items + sms - 1) / sms) * (double)sms);
      if (eff > best_eff + 0.02) { best_eff = eff; best = sk; }
    }
    if (best > 1) {
      p.kb_per_split = (p.k_blocks + best - 1) / best;
      p.split_k = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    }
  }

  // bf16 outputs of plain Linear / CE-grad epilogues leave through TMA bulk stores
  CUtensorMap tmOut, tmPre;
  memset(&tmOut, 0, sizeof tmOut);
  memset(&tmPre, 0, sizeof tmPre);
  p.tma_out = 0;
  const bool lin_fast = epi->mode == KMB_EPI_LINEAR && epi->out_bf16 && !epi->out_f32 && !epi->residual &&
                        epi->dropout_p <= 0.f;
  if ((lin_fast || epi->mode == KMB_EPI_CE_GRAD) && tile_n >= 64 && (epi->ld_bf16 % 8) == 0 && al16(epi->out_bf16) &&
      (!epi->out_preact || al16(epi->out_preact))) {
    rc = make_tmap(&tmOut, epi->out_bf16, 0, (uint64_t)M, (uint64_t)N, (uint64_t)epi->ld_bf16, 64, 32);
    if (rc) return rc;
    if (epi->out_preact) {
      rc = make_tmap(&tmPre, epi->out_preact, 0, (uint64_t)M, (uint64_t)N, (uint64_t)epi->ld_bf16, 64, 32);
      if (rc) return rc;
    }
    p.tma_out = 1;
  }
  if (epi->mode == KMB_EPI_CE_GRAD && !p.tma_out) {
    kmb_set_last_error("kmb_gemm: CE_GRAD needs tile_n >= 64 and a 16-byte aligned bf16 output", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define KMB_DISPATCH_BN(BNV)                                                        \
  case BNV:                                                                         \
    return elt == 0 ? launch_major<BNV, 0>(a_mn, b_mn, tmA, tmB, tmOut, tmPre, p, st) \
                    : launch_major<BNV, 1>(a_mn, b_mn, tmA, tmB, tmOut, tmPre, p, st);
  switch (tile_n) {
    KMB_DISPATCH_BN(32)
    KMB_DISPATCH_BN(64)
    KMB_DISPATCH_BN(128)
    KMB_DISPATCH_BN(256)
  }
#undef KMB_DISPATCH_BN
  return KMB_ERR_ARG;
}
