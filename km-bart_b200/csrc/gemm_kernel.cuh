// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = epilogue(sum_k A(m,k) * B(n,k)).
//
// One persistent CTA per SM, 128 + 128 * ES threads (384 or 640), warp-specialised.  CG = 1: every CTA owns 128 x BN tiles
// (tcgen05.mma cta_group::1).  CG = 2: the two CTAs of a 2-CTA cluster own one 256 x BN tile
// (cta_group::2): each loads its own 128 rows of A and HALF of the B rows, which cuts the
// bytes an SM has to ingest per flop by a third -- the measured limiter of the single-CTA
// kernel (profiles/r01_*: ~47 B/clk/SM of TMA ingest, tensor pipe 45 % busy).
//   warp 0      TMA producer   (one elected lane; cp.async.bulk.tensor 2D, SWIZZLE_128B)
//   warp 1      MMA issuer     (one elected lane of the pair's leader CTA; M=128*CG, N=BN, K=16|8)
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4..19 epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> swizzled smem
//                               transpose -> coalesced global I/O)
// Three mbarrier pipelines: smem full/empty (TMA<->MMA), tmem full/empty (MMA<->epilogue),
// and a static round-robin tile schedule.  Operands may be K-major or MN-major (the
// backward GEMMs contract over the token dimension, which is the slow dimension of
// every activation), bf16 or tf32.
//
// Replaces the cuBLAS calls the reference issues through nn.Linear / F.linear
// (SURVEY.md §2.3(b) K2,K5,K7,K8,K10) — see include/kmbart.h for the call-site map.
#pragma once
#include <cuda.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

constexpr int BM = 128;
// Epilogue warps: ES ("epilogue slices") warps per TMEM lane quadrant, each taking every ES-th 32-column chunk of the
// accumulator; a template parameter chosen per launch.  Measured (profiles/r01f_*, build/gemm_test bench):
//   ES = 2 (8 warps, 384 threads, 168 registers, one more pipeline stage): best for plain bias / bf16-out epilogues,
//          which hide behind the next tile's MMAs anyway (fc1-shape 1198 vs 1103 TFLOP/s, 8192^3 1330 vs 1198);
//   ES = 4 (16 warps, 640 threads, 96 registers): activation / activation-gradient / cross-entropy epilogues issue
//          at ~0.4 IPC with two warps per scheduler and pace the tensor pipe (fc1+GELU: tensor 34 % active); four
//          per scheduler fill the issue slots (fc1+GELU 84 -> 68 us, GELU-grad 107 -> 75 us).
constexpr int EPI_SLICES_MAX = 4;   // also the number of cross-entropy partials per n-tile (CE epilogues use ES = 4)
constexpr int TILE_BYTES_ROW = 128;  // one swizzle span: 64 bf16 or 32 tf32
constexpr int A_TILE_BYTES = BM * TILE_BYTES_ROW;

// optional in-kernel timeline (block 0 only), enabled by kmb_gemm_debug_timeline(1): %globaltimer ns at
// [0] entry, [1] setup done, [2] first stage landed, [3] first accumulator complete, [4] first epilogue done,
// [5] last epilogue done, [6] exit
static __device__ unsigned long long g_gemm_timeline[8];   // one copy per translation unit
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, k_blocks;
  KmbGemmEpilogue e;
  int vec_ok;  // all leading dims / pointers allow 16-byte row-segment access
  int tma_out; // bf16 outputs leave through smem + TMA bulk stores (tmOut / tmPre valid)
  int split_k, kb_per_split;  // split-K work items; partial sums are reduced with fp32 red.global.add
  uint32_t drop_thresh16;
  float drop_scale;
  int timeline;
};

template <int BN, int CG, int ES>
struct Cfg {
  static constexpr int BN_CTA = BN / CG;  // B rows resident in one CTA
  static constexpr int B_TILE_BYTES = BN_CTA * TILE_BYTES_ROW;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int EPI_STAGING = 4 * ES * 4096;  // one 32x32 fp32 transpose tile per epilogue warp
  static constexpr int STAGES_RAW = (227 * 1024 - 1024 - 512 - EPI_STAGING) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN) <= 32 ? 32 : (2 * BN) <= 64 ? 64 : (2 * BN) <= 128 ? 128 : (2 * BN) <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/ + EPI_STAGING;
};

// Instruction descriptor, field layout per cute/arch/mma_sm100_desc.hpp InstrDescriptor.
__device__ __forceinline__ uint32_t make_idesc(int elt, int a_mn, int b_mn, int n, int m) {
  uint32_t d = 0;
  d |= 1u << 4;                             // c_format = F32
  const uint32_t fmt = elt == 0 ? 1u : 2u;  // BF16 : TF32
  d |= fmt << 7;                            // a_format
  d |= fmt << 10;                           // b_format
  d |= (uint32_t)a_mn << 15;                // a_major (0 = K, 1 = MN)
  d |= (uint32_t)b_mn << 16;                // b_major
  d |= (uint32_t)(n >> 3) << 17;            // n_dim
  d |= (uint32_t)(m >> 4) << 24;            // m_dim
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
// 32 rows x 32 columns (one tcgen05.ld chunk) are packed to bf16 and written into one of the warp's two
// 2 KB staging buffers in the SWIZZLE_64B pattern (16-byte piece p of row r at r*64 + ((p ^ ((r >> 1) & 3)) << 4)),
// which is bank-conflict free for thread-per-row stores and is the layout a 64B-swizzled tensor map
// expects; one elected lane then issues an asynchronous bulk store.  The buffers ping-pong, so a store
// is still draining while the next chunk is converted.  Rows / columns outside [M, N] are clipped by
// the TMA unit.
__device__ __forceinline__ void stage_bf16_chunk_and_store(float4* st, int& pp, int lane, const float (&v)[32], const void* tmap,
                                                           int col0, int row0) {
  if (lane == 0) tma_store_wait_read1();  // the store issued two chunks ago has drained this buffer
  __syncwarp();
  uint4* sp = reinterpret_cast<uint4*>(st) + pp * 128;
#pragma unroll
  for (int pc = 0; pc < 4; ++pc)
    sp[lane * 4 + (pc ^ ((lane >> 1) & 3))] = make_uint4(pack_bf16(v[8 * pc], v[8 * pc + 1]), pack_bf16(v[8 * pc + 2], v[8 * pc + 3]),
                                                         pack_bf16(v[8 * pc + 4], v[8 * pc + 5]), pack_bf16(v[8 * pc + 6], v[8 * pc + 7]));
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmap, sp, col0, row0);
    tma_store_commit();
  }
  pp ^= 1;
}

// 32 consecutive bias values (uniform across the warp) as independent 16-byte loads; zeros when
// bias is null; clamped scalar loads for the ragged last chunk.
__device__ __forceinline__ void load_bias32(const float* bias, int col0, int N, float (&b)[32]) {
  if (!bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = 0.f;
  } else if (col0 + 32 <= N && (reinterpret_cast<uintptr_t>(bias + col0) & 15) == 0) {
    const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t4 = __ldg(b4 + j);
      b[4 * j] = t4.x; b[4 * j + 1] = t4.y; b[4 * j + 2] = t4.z; b[4 * j + 3] = t4.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int cidx = col0 + j < N ? col0 + j : N - 1;
      b[j] = __ldg(bias + (cidx < 0 ? 0 : cidx));
    }
  }
}

template <int BN, int ELT, int A_MN, int B_MN, int CG, int ES>
__global__ void __launch_bounds__(128 + 128 * ES, 1)
gemm_tc05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmPre,
                 const GemmParams p) {
  using C = Cfg<BN, CG, ES>;
  constexpr int STAGES = C::STAGES;
  constexpr int EPI_SLICES = ES;
  constexpr int EPI_WARPS = 4 * ES;
  constexpr int BN_CTA = C::BN_CTA;
  constexpr int ELT_BYTES = ELT == 0 ? 2 : 4;
  constexpr int BK = TILE_BYTES_ROW / ELT_BYTES;  // 64 bf16 / 32 tf32 per k-block
  constexpr int UK = 32 / ELT_BYTES;              // UMMA K: 16 bf16 / 8 tf32
  constexpr int KSTEPS = BK / UK;                 // 4
  constexpr int MN_CHUNK = BK;                    // elements per 128-byte row of an MN-major tile
  static_assert(!B_MN || (BN_CTA % MN_CHUNK) == 0, "MN-major B is built from 128-byte-wide chunks");
  static_assert(CG == 1 || ELT == 0, "CTA pairs are instantiated for bf16 only");

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
  uint64_t* aux_bars = bars + 2 * STAGES + 5;   // one per epilogue warp: activation-gradient operand tiles (TMA)

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool tl = p.timeline && blockIdx.x == 0;
  if (tl && threadIdx.x == 0) g_gemm_timeline[0] = gtimer();
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;   // 0 = leader of the pair
  const int tile_id0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_stride = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < EPI_WARPS; ++s) mbar_init(&aux_bars[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS * CG);   // one arrival per epilogue warp of every CTA of the group
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    else tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // peer barriers are initialised before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // barriers, TMEM and descriptors are set up while the previous kernel drains; no global access before here
  if (tl && threadIdx.x == 0) g_gemm_timeline[1] = gtimer();

  const int total_tiles = p.m_tiles * p.n_tiles * p.split_k;

  if (warp < 4) {
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // CG == 2: both CTAs load their own A rows / B half; every byte is credited to the LEADER's full barrier
      auto load = [&](void* dst, const CUtensorMap* tm, int stage_, int c0, int c1) {
        if (CG == 2) tma_load_2d_pair(dst, tm, mapa_shared(smem_u32(&full_bar[stage_]), 0), c0, c1);
        else tma_load_2d(dst, tm, &full_bar[stage_], c0, c1);
      };
      for (int t = tile_id0; t < total_tiles; t += tile_stride) {
        const int m0 = (t % p.m_tiles) * (BM * CG) + (int)cta_rank * BM;
        const int n0 = ((t / p.m_tiles) % p.n_tiles) * BN + (int)cta_rank * BN_CTA;
        const int kb0 = (t / (p.m_tiles * p.n_tiles)) * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_TILE_BYTES;
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], CG * C::STAGE_BYTES);
          const int k0 = kb * BK;
          if (A_MN == 0) {
            load(sa, &tmA, stage, k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / MN_CHUNK; ++c)
              load(sa + c * BK * TILE_BYTES_ROW, &tmA, stage, m0 + c * MN_CHUNK, k0);
          }
          if (B_MN == 0) {
            load(sb, &tmB, stage, k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN_CTA / MN_CHUNK; ++c)
              load(sb + c * BK * TILE_BYTES_ROW, &tmB, stage, n0 + c * MN_CHUNK, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && cta_rank == 0) {
      const uint32_t idesc = make_idesc(ELT, A_MN, B_MN, BN, BM * CG);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = tile_id0; t < total_tiles; t += tile_stride) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        const int kb_lo = (t / (p.m_tiles * p.n_tiles)) * p.kb_per_split;
        const int kb_n = min(p.k_blocks, kb_lo + p.kb_per_split) - kb_lo;
        for (int kb = 0; kb < kb_n; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (tl && kb == 0 && t == tile_id0) g_gemm_timeline[2] = gtimer();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk) {
            uint64_t da, db;
            if (A_MN == 0) da = make_smem_desc_sw128(sa + kk * 32, 16, 1024);
            else da = make_smem_desc_sw128(sa + kk * UK * TILE_BYTES_ROW, BK * TILE_BYTES_ROW, 1024);
            if (B_MN == 0) db = make_smem_desc_sw128(sb + kk * 32, 16, 1024);
            else db = make_smem_desc_sw128(sb + kk * UK * TILE_BYTES_ROW, BK * TILE_BYTES_ROW, 1024);
            if (CG == 2) umma_f16_pair(tmem_d, da, db, idesc, (kb | kk) != 0);
            else if (ELT == 0) umma_f16(tmem_d, da, db, idesc, (kb | kk) != 0);
            else umma_tf32(tmem_d, da, db, idesc, (kb | kk) != 0);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if (CG == 2) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (CG == 2) umma_commit_pair(&tfull_bar[acc]);  // accumulator complete (both CTAs' epilogues wake)
        else umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }
  } else {
    // ===================== epilogue (EPI_WARPS warps) =====================
    // warp w may only touch TMEM lanes [32*(w%4), +32); warps 4..7 take 32-column chunks 0, 4, ...,
    // warps 8..11 chunks 1, 5, ... and so on.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;   // column slice of this warp: chunks half, half + EPI_SLICES, ...
    float4* st = stage_tiles + (warp - 4) * 256;
    int pp = 0;  // ping-pong index of the TMA-store staging buffers
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t drop_key = 0;
    if (p.drop_thresh16 && p.e.dropout_seed) drop_key = dropout_key(*p.e.dropout_seed, p.e.dropout_tag);
    const uint32_t tempty_leader0 = CG == 2 ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : 0u;
    // Activation-gradient epilogues (dX = dY W o act'(aux)): the 32x32 bf16 tile of `aux` that belongs to a chunk is
    // fetched by TMA (64-byte swizzle, tmPre is encoded over aux in this mode) into the second half of the warp's
    // staging buffer one chunk ahead -- across tile boundaries too -- so its latency never sits between the TMEM
    // load and the math; the bf16 output then leaves through the first half only.
    const bool grad_act = p.e.mode == KMB_EPI_LINEAR && p.tma_out && (p.e.act == KMB_ACT_GELU_GRAD || p.e.act == KMB_ACT_TANH_GRAD);
    uint64_t* aux_bar = &aux_bars[warp - 4];
    uint32_t aux_phase = 0;
    uint4* aux_buf = reinterpret_cast<uint4*>(st) + 128;
    auto aux_issue = [&](int tt, int cc) {   // lane 0
      const int mb = tt % p.m_tiles, nb = (tt / p.m_tiles) % p.n_tiles;
      mbar_arrive_expect_tx(aux_bar, 2048);
      tma_load_2d(aux_buf, &tmPre, aux_bar, nb * BN + cc * 32, mb * (BM * CG) + (int)cta_rank * BM + q * 32);
    };
    if (grad_act && half < BN / 32 && lane == 0 && tile_id0 < total_tiles) aux_issue(tile_id0, half);
    for (int t = tile_id0; t < total_tiles; t += tile_stride) {
      const int m_blk = t % p.m_tiles, n_blk = (t / p.m_tiles) % p.n_tiles;
      const int row0 = m_blk * (BM * CG) + (int)cta_rank * BM + q * 32;
      const int row = row0 + lane;
      const int n0 = n_blk * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (tl && warp == 4 && lane == 0 && t == tile_id0) g_gemm_timeline[3] = gtimer();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      const bool row_ok = row < p.M;
      int rows_valid = p.M - row0;
      rows_valid = rows_valid > 32 ? 32 : rows_valid;

      if (p.e.mode == KMB_EPI_LINEAR && p.tma_out) {
        // ---- fast path: bias / activation -> bf16 -> TMA store, 32 columns at a time.  Operands that do not
        // depend on the accumulator are requested before the TMEM load is awaited: the bias slice (one value per
        // lane) and, for activation gradients, the aux tile that TMA parked in shared memory a chunk earlier.
        const KmbGemmEpilogue& e = p.e;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += EPI_SLICES) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          const int col0 = n0 + c * 32;
          uint4 ax[4];
          if (grad_act) {
            mbar_wait(aux_bar, aux_phase);
            aux_phase ^= 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) ax[j] = aux_buf[lane * 4 + (j ^ ((lane >> 1) & 3))];
            __syncwarp();
            if (lane == 0) {   // the buffer is free again: request the next chunk of this warp
              const int cn = c + EPI_SLICES;
              if (cn < BN / 32) aux_issue(t, cn);
              else if (t + tile_stride < total_tiles) aux_issue(t + tile_stride, half);
            }
          }
          // lane j keeps bias[col0 + j] (one register); it is broadcast with shuffles after the TMEM load has landed
          float bias_lane = 0.f;
          if (e.bias) {
            const int cb = col0 + lane;
            bias_lane = __ldg(e.bias + (cb < p.N ? cb : p.N - 1));
          }
          tmem_ld_wait();
          if (rows_valid > 0 && col0 < p.N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (e.alpha != 1.0f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
            }
            if (e.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bias_lane, j);
            }
            if (e.act == KMB_ACT_GELU) {
              if (e.out_preact) stage_bf16_chunk_and_store(st, pp, lane, v, &tmPre, col0, row0);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
            } else if (e.act == KMB_ACT_TANH) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
            } else if (grad_act) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t w4[4] = {ax[j].x, ax[j].y, ax[j].z, ax[j].w};
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                  const int jj = 8 * j + 2 * l;
                  const float a0 = __uint_as_float(w4[l] << 16), a1 = __uint_as_float(w4[l] & 0xFFFF0000u);
                  if (e.act == KMB_ACT_GELU_GRAD) {
                    v[jj] *= gelu_erf_grad(a0);
                    v[jj + 1] *= gelu_erf_grad(a1);
                  } else {
                    v[jj] *= (1.f - a0 * a0);
                    v[jj + 1] *= (1.f - a1 * a1);
                  }
                }
              }
            }
            if (grad_act) {
              if (lane == 0) tma_store_wait_read();   // single staging buffer in this mode (the other half holds aux)
              pp = 0;
            }
            stage_bf16_chunk_and_store(st, pp, lane, v, &tmOut, col0, row0);
          }
        }
      } else if (p.e.mode == KMB_EPI_LINEAR) {
#pragma unroll 1
        for (int c = half; c < BN / 32; c += EPI_SLICES) {
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
        // partial index = n_blk * EPI_SLICES + slice.  The bias slice of a chunk is fetched up front as
        // independent vector loads (a per-column load -> add -> max chain serialises 32 L2 latencies
        // per chunk and made this epilogue 4x longer than the tile's MMA time).
        float mx = -INFINITY, sm = 0.f;
        const int64_t label = row_ok ? p.e.labels[row] : -100;
        constexpr float LOG2E = 1.4426950408889634f;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += EPI_SLICES) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          const int col0 = n0 + c * 32;
          int ncols = p.N - col0;
          ncols = ncols > 32 ? 32 : ncols;
          float v[32];
          load_bias32(p.e.bias, col0, p.N, v);
          tmem_ld_wait();
          if (row_ok && ncols > 0) {
            float cm = -INFINITY;
            if (ncols == 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v[j] = fmaf(__uint_as_float(r[j]), p.e.alpha, v[j]);
                cm = fmaxf(cm, v[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v[j] = j < ncols ? fmaf(__uint_as_float(r[j]), p.e.alpha, v[j]) : -INFINITY;
                cm = fmaxf(cm, v[j]);
              }
            }
            const float nm = fmaxf(mx, cm);
            const float nm2 = -nm * LOG2E;
            float cs = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) cs += fast_ex2(fmaf(v[j], LOG2E, nm2));   // exp(-inf) = 0 past ncols
            sm = sm * fast_ex2(fmaf(mx, LOG2E, nm2)) + cs;
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
          p.e.ce_max[(int64_t)row * (EPI_SLICES * p.n_tiles) + EPI_SLICES * n_blk + half] = mx;
          p.e.ce_sum[(int64_t)row * (EPI_SLICES * p.n_tiles) + EPI_SLICES * n_blk + half] = sm;
        }
      } else {  // KMB_EPI_CE_GRAD: dlogits = (softmax - onehot) * gscale, bf16, through the TMA store path
        const int64_t label = row_ok ? p.e.labels[row] : -100;
        const float lse = row_ok ? p.e.ce_lse[row] : 0.f;
        const float gs = (label >= 0) ? *p.e.ce_gscale : 0.f;
        constexpr float LOG2E = 1.4426950408889634f;
        const float a2 = p.e.alpha * LOG2E, nl2 = -lse * LOG2E;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += EPI_SLICES) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          const int col0 = n0 + c * 32;
          float v[32];
          load_bias32(p.e.bias, col0, p.N, v);
          tmem_ld_wait();
          if (rows_valid > 0 && col0 < p.N) {
            const int lj = (int)label - col0;   // column of the label inside this chunk (or out of range)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              // exp(alpha*acc + bias - lse) in the log2 domain: one FFMA pair + ex2
              const float pv = fast_ex2(fmaf(__uint_as_float(r[j]), a2, fmaf(v[j], LOG2E, nl2)));
              v[j] = (pv - (j == lj ? 1.f : 0.f)) * gs;
            }
            stage_bf16_chunk_and_store(st, pp, lane, v, &tmOut, col0, row0);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty_leader0 + acc * 8);   // the MMA issuer lives in the leader CTA
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (tl && warp == 4 && lane == 0) { if (t == tile_id0) g_gemm_timeline[4] = gtimer(); g_gemm_timeline[5] = gtimer(); }
    }
    if (lane == 0) tma_store_wait_read();  // staging smem must outlive the last bulk store's read
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // no CTA exits (or frees TMEM) while its peer can still signal it
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
  if (tl && threadIdx.x == 0) g_gemm_timeline[6] = gtimer();
}

}  // namespace kmb
