// tcgen05 / TMEM / TMA fused multi-head attention for sm_100a, head_dim 64, S_q, S_k <= 128 (every attention of the base
// model: encoder 100 x 100, decoder self 48 x 48 causal, cross 48 x 100), forward and backward.
// replaces: HF-3.0.2 SelfAttention.forward — bmm(q, k^T), additive causal mask (src/model/model.py:63-70), key-padding
//   masked_fill (src/model/modules.py:130-131), softmax, bmm(p, v) — and its autograd backward, instantiated at
//   src/model/modules.py:84 and src/model/model.py:35.
//
// One (batch, head) pair is ONE tile: S = Q K^T is a single 128 x 128 x 64 tcgen05.mma chain into tensor memory, the
// softmax runs on accumulator rows read back with tcgen05.ld (one thread per query row, no cross-lane reductions), P goes to
// shared memory once in the 128-byte-swizzled layout and feeds the second MMA.  Backward keeps S, dP, dV, dK, dQ in tensor
// memory (448 of the 512 columns); P and dS are written once and consumed both K-major (dQ = dS K) and MN-major
// (dV = P^T dO, dK = dS^T Q) — the same bytes serve both descriptors.  Q / K / V / dO tiles arrive by TMA two pairs ahead.
//   warps 0-3  softmax / gradient math + global stores (TMEM lane quadrant = warp)
//   warp  4    MMA issuer (one elected lane) + TMEM allocation
//   warp  5    TMA producer (one elected lane)
// Rows / keys beyond S_q / S_k (the 128-row TMA boxes overhang into the next batch element or are zero-filled) are masked
// to P = dS = 0, so they contribute nothing; their outputs are never stored.
#include <cuda.h>
#include <mutex>
#include <string.h>
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

constexpr int AT_CW = 16;                          // compute warps: 4 TMEM lane quadrants x 4 column groups of 32
constexpr int AT_CT = AT_CW * 32;                  // compute threads
constexpr int AT_THREADS = AT_CT + 64;             // + MMA warp + TMA warp
constexpr int AT_TILE = 128 * 128;                 // bytes of one [128 rows x 64 bf16] tile
constexpr float AT_LOG2E = 1.4426950408889634f;

struct AtParams {
  const bf16* o; int64_t ldo;                      // bwd: forward output (for D = rowsum(dO o O))
  bf16* out; int64_t ldout;                        // fwd: O
  float* lse;                                      // [B, H, Sq] natural-log lse of the scaled scores (fwd: written; bwd: read)
  const uint8_t* key_pad;                          // [B, Sk] or null
  bf16 *dq, *dk, *dv; int64_t lddq, lddk, lddv;
  int B, H, Sq, Sk, causal;
  float scale;
};

__device__ __forceinline__ uint32_t at_idesc(int a_mn, int b_mn, int n, int m) {
  uint32_t d = 0;
  d |= 1u << 4;                 // c_format = F32
  d |= 1u << 7;                 // a_format = BF16
  d |= 1u << 10;                // b_format = BF16
  d |= (uint32_t)a_mn << 15;    // a_major (0 = K, 1 = MN)
  d |= (uint32_t)b_mn << 16;    // b_major
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}
// K-major operand: [rows][128 B = 64 k], k-step kk (16 elements) of k-block kb (tiles 16 KB apart)
__device__ __forceinline__ uint64_t at_desc_k(uint32_t tile, int kk) {
  return make_smem_desc_sw128(tile + (kk >> 2) * AT_TILE + (kk & 3) * 32, 16, 1024);
}
// MN-major operand: [k rows][128 B = 64 mn], 64-wide mn chunks 16 KB apart, k-step kk = 16 rows
__device__ __forceinline__ uint64_t at_desc_mn(uint32_t tile, int kk) {
  return make_smem_desc_sw128(tile + kk * 2048, AT_TILE, 1024);
}
__device__ __forceinline__ void at_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void at_bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(AT_CT) : "memory"); }
__device__ __forceinline__ float at_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t at_pack(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// 16-byte chunk c (8 bf16) of row r inside a [128 rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t at_sw(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// key mask words of one pair: bit j of word w set = key 32 w + j is masked for EVERY query (padding or >= Sk)
__device__ __forceinline__ void at_key_mask(const AtParams& p, int b, uint32_t* words) {   // threads 0..127
  const int j = threadIdx.x;
  bool masked = j >= p.Sk;
  if (!masked && p.key_pad) masked = p.key_pad[(int64_t)b * p.Sk + j] != 0;
  const uint32_t w = __ballot_sync(0xffffffffu, masked);
  if ((threadIdx.x & 31) == 0) words[threadIdx.x >> 5] = w;
}
// mask word of query row r for the 32 keys of column group cg
__device__ __forceinline__ uint32_t at_row_mask(uint32_t w, int causal, int r, int cg) {
  if (causal) {
    const int lim = r - cg * 32;          // keys 32 cg + j with j > lim lie in the future
    if (lim < 31) w |= lim < 0 ? 0xffffffffu : (0xfffffffeu << lim);
  }
  return w;
}
__device__ __forceinline__ void at_store_row32(bf16* dst, const uint32_t (&v)[32], float mul) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(dst + q * 8) =
        make_uint4(at_pack(__uint_as_float(v[8 * q]) * mul, __uint_as_float(v[8 * q + 1]) * mul),
                   at_pack(__uint_as_float(v[8 * q + 2]) * mul, __uint_as_float(v[8 * q + 3]) * mul),
                   at_pack(__uint_as_float(v[8 * q + 4]) * mul, __uint_as_float(v[8 * q + 5]) * mul),
                   at_pack(__uint_as_float(v[8 * q + 6]) * mul, __uint_as_float(v[8 * q + 7]) * mul));
}

struct AtSmem {
  uint64_t full[2], empty[2], s_full, p_full, o_full;
  uint32_t tmem_slot;
  uint32_t kmask[2][4];
  float red[2][4][128];        // fwd: partial row max / row sum of the four column groups; bwd: red[0][0] = D
};

// ------------------------------------------------------------------ forward
constexpr int AT_FWD_STAGE = 3 * AT_TILE;                          // Q, K, V
constexpr int AT_FWD_SMEM = 1024 + 2 * AT_FWD_STAGE + 2 * AT_TILE + 5120;   // align + stages + P + barriers / scratch

__global__ void __launch_bounds__(AT_THREADS, 1) attn_fwd_tc05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV, const AtParams p) {
  extern __shared__ uint8_t at_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage0 = smem;
  uint8_t* pbuf = smem + 2 * AT_FWD_STAGE;
  AtSmem& sh = *reinterpret_cast<AtSmem*>(smem + 2 * AT_FWD_STAGE + 2 * AT_TILE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 1); }
    mbar_init(&sh.s_full, 1); mbar_init(&sh.p_full, AT_CW); mbar_init(&sh.o_full, 1);
    fence_mbar_init();
  }
  if (warp == AT_CW) tmem_alloc(&sh.tmem_slot, 256);
  if (warp == AT_CW + 1 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_slot;
  pdl_wait();
  const int n_pairs = p.B * p.H;
  const int SkR = (p.Sk + 15) & ~15, SqR = (p.Sq + 15) & ~15;     // MMA extents: N of S, K of P V
  if (warp == AT_CW + 1) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
        const int st = it & 1, b = pair / p.H, h = pair % p.H;
        mbar_wait(&sh.empty[st], ((it >> 1) & 1) ^ 1);
        uint8_t* s = stage0 + st * AT_FWD_STAGE;
        mbar_arrive_expect_tx(&sh.full[st], 3 * AT_TILE);
        tma_load_2d(s, &tmQ, &sh.full[st], h * 64, b * p.Sq);
        tma_load_2d(s + AT_TILE, &tmK, &sh.full[st], h * 64, b * p.Sk);
        tma_load_2d(s + 2 * AT_TILE, &tmV, &sh.full[st], h * 64, b * p.Sk);
      }
    }
  } else if (warp == AT_CW) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id_s = at_idesc(0, 0, SkR, 128), id_o = at_idesc(0, 1, 64, 128);
      const uint32_t pb = smem_u32(pbuf);
      int it = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t sq = smem_u32(stage0 + st * AT_FWD_STAGE), sk = sq + AT_TILE, sv = sq + 2 * AT_TILE;
        mbar_wait(&sh.full[st], (it >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, at_desc_k(sq, kk), at_desc_k(sk, kk), id_s, kk != 0);   // S = Q K^T
        umma_commit(&sh.s_full);
        mbar_wait(&sh.p_full, it & 1);
        tc_fence_after();
        for (int kk = 0; kk < SkR / 16; ++kk) umma_f16(tmem + 128, at_desc_k(pb, kk), at_desc_mn(sv, kk), id_o, kk != 0);   // O = P V
        umma_commit(&sh.o_full);
        umma_commit(&sh.empty[st]);
      }
    }
  } else {
    // ===================== softmax / epilogue: thread = (query row, 32-key column group) =====================
    const int qd = warp & 3, cg = warp >> 2;
    const int r = qd * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16);
    const float sc2 = p.scale * AT_LOG2E;
    const bool act = qd * 32 < SqR && cg * 32 < SkR;      // warp-uniform (tcgen05.ld is .sync.aligned): this warp's 32 x 32 scores exist
    int it = 0;
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
      const int b = pair / p.H, h = pair % p.H;
      uint32_t* km = sh.kmask[it & 1];
      if (threadIdx.x < 128) at_key_mask(p, b, km);
      at_bar_compute();
      mbar_wait(&sh.s_full, it & 1);
      tc_fence_after();
      const bool row_ok = r < p.Sq;
      uint32_t v[32];
      uint32_t w = 0xffffffffu;
      float m2 = -INFINITY;
      if (act) {
        tmem_ld32(trow + cg * 32, v);
        tmem_ld_wait();
        w = row_ok ? at_row_mask(km[cg], p.causal, r, cg) : 0xffffffffu;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (!((w >> j) & 1)) m2 = fmaxf(m2, __uint_as_float(v[j]) * sc2);
      }
      sh.red[0][cg][r] = m2;
      at_bar_compute();
      m2 = fmaxf(fmaxf(sh.red[0][0][r], sh.red[0][1][r]), fmaxf(sh.red[0][2][r], sh.red[0][3][r]));
      float l = 0.f;
      if (act) {
        const float nm = (m2 == -INFINITY) ? 0.f : -m2;
        uint32_t e2[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = ((w >> j) & 1) ? 0.f : at_ex2(fmaf(__uint_as_float(v[j]), sc2, nm));
          const float e1 = ((w >> (j + 1)) & 1) ? 0.f : at_ex2(fmaf(__uint_as_float(v[j + 1]), sc2, nm));
          l += e0 + e1;
          e2[j >> 1] = at_pack(e0, e1);
        }
        uint8_t* blk = pbuf + (cg >> 1) * AT_TILE;      // P (bf16, unnormalised) in the swizzled K-major layout
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(blk + at_sw(r, (cg & 1) * 4 + q)) = make_uint4(e2[4 * q], e2[4 * q + 1], e2[4 * q + 2], e2[4 * q + 3]);
      }
      sh.red[1][cg][r] = l;
      tc_fence_before();
      at_fence_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.p_full);
      at_bar_compute();
      l = (sh.red[1][0][r] + sh.red[1][1][r]) + (sh.red[1][2][r] + sh.red[1][3][r]);
      if (cg == 0 && row_ok && p.lse) p.lse[((int64_t)b * p.H + h) * p.Sq + r] = (l > 0.f) ? (m2 + log2f(l)) * (1.f / AT_LOG2E) : -INFINITY;
      mbar_wait(&sh.o_full, it & 1);
      tc_fence_after();
      if (cg < 2) {      // O: 64 columns, column groups 0 and 1
        // fully masked row: 0 * inf = NaN, like softmax over an all -inf row in the reference
        tmem_ld32(trow + 128 + cg * 32, v);
        tmem_ld_wait();
        if (row_ok) at_store_row32(p.out + ((int64_t)b * p.Sq + r) * p.ldout + h * 64 + cg * 32, v, 1.f / l);
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_CW) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ------------------------------------------------------------------ backward
constexpr int AT_BWD_STAGE = 4 * AT_TILE;                          // Q, K, V, dO
constexpr int AT_OROW = 144;                                      // padded row pitch of the prefetched O tile (conflict-free row reads)
constexpr int AT_BWD_SMEM = 1024 + 2 * AT_BWD_STAGE + 4 * AT_TILE + 128 * AT_OROW + 5120;   // align + stages + P + dS + O rows + barriers / scratch

__global__ void __launch_bounds__(AT_THREADS, 1) attn_bwd_tc05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                                                                     const AtParams p) {
  extern __shared__ uint8_t at_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage0 = smem;
  uint8_t* pbuf = smem + 2 * AT_BWD_STAGE;          // P  [sq][sk], two 64-key blocks
  uint8_t* dsbuf = pbuf + 2 * AT_TILE;              // dS [sq][sk]
  uint8_t* obuf = dsbuf + 2 * AT_TILE;              // O rows of the NEXT pair (cp.async, one pair ahead)
  AtSmem& sh = *reinterpret_cast<AtSmem*>(obuf + 128 * AT_OROW);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 1); }
    mbar_init(&sh.s_full, 1); mbar_init(&sh.p_full, AT_CW); mbar_init(&sh.o_full, 1);
    fence_mbar_init();
  }
  if (warp == AT_CW) tmem_alloc(&sh.tmem_slot, 512);
  if (warp == AT_CW + 1 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_slot;
  // tensor-memory columns: S 0..127 | dP 128..255 | dV 256..319 | dK 320..383 | dQ 384..447
  pdl_wait();
  const int n_pairs = p.B * p.H;
  const int SkR = (p.Sk + 15) & ~15, SqR = (p.Sq + 15) & ~15;
  if (warp == AT_CW + 1) {
    if (lane == 0) {
      int it = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
        const int st = it & 1, b = pair / p.H, h = pair % p.H;
        mbar_wait(&sh.empty[st], ((it >> 1) & 1) ^ 1);
        uint8_t* s = stage0 + st * AT_BWD_STAGE;
        mbar_arrive_expect_tx(&sh.full[st], 4 * AT_TILE);
        tma_load_2d(s, &tmQ, &sh.full[st], h * 64, b * p.Sq);
        tma_load_2d(s + AT_TILE, &tmK, &sh.full[st], h * 64, b * p.Sk);
        tma_load_2d(s + 2 * AT_TILE, &tmV, &sh.full[st], h * 64, b * p.Sk);
        tma_load_2d(s + 3 * AT_TILE, &tmDO, &sh.full[st], h * 64, b * p.Sq);
      }
    }
  } else if (warp == AT_CW) {
    if (lane == 0) {
      const uint32_t id_s = at_idesc(0, 0, SkR, 128);      // S = Q K^T, dP = dO V^T
      const uint32_t id_t = at_idesc(1, 1, 64, 128);       // dV = P^T dO, dK = dS^T Q  (K extent = query rows)
      const uint32_t id_q = at_idesc(0, 1, 64, 128);       // dQ = dS K                 (K extent = keys)
      const uint32_t pb = smem_u32(pbuf), db = smem_u32(dsbuf);
      int it = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t sq = smem_u32(stage0 + st * AT_BWD_STAGE), sk = sq + AT_TILE, sv = sq + 2 * AT_TILE, sdo = sq + 3 * AT_TILE;
        mbar_wait(&sh.full[st], (it >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, at_desc_k(sq, kk), at_desc_k(sk, kk), id_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16(tmem + 128, at_desc_k(sdo, kk), at_desc_k(sv, kk), id_s, kk != 0);
        umma_commit(&sh.s_full);
        mbar_wait(&sh.p_full, it & 1);
        tc_fence_after();
        for (int kk = 0; kk < SqR / 16; ++kk) umma_f16(tmem + 256, at_desc_mn(pb, kk), at_desc_mn(sdo, kk), id_t, kk != 0);
        for (int kk = 0; kk < SqR / 16; ++kk) umma_f16(tmem + 320, at_desc_mn(db, kk), at_desc_mn(sq, kk), id_t, kk != 0);
        for (int kk = 0; kk < SkR / 16; ++kk) umma_f16(tmem + 384, at_desc_k(db, kk), at_desc_mn(sk, kk), id_q, kk != 0);
        umma_commit(&sh.o_full);
        umma_commit(&sh.empty[st]);
      }
    }
  } else {
    const int qd = warp & 3, cg = warp >> 2;
    const int r = qd * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16);
    const float sc2 = p.scale * AT_LOG2E;
    const bool act = qd * 32 < SqR && cg * 32 < SkR;      // warp-uniform
    // column group 0 prepares the per-row constants of a pair: key mask, -lse * log2(e), D = rowsum(dO o O).  The O rows and
    // lse of pair i+1 are requested while pair i is processed (cp.async into obuf / a register): a dependent global load at
    // the head of every pair was the longest link of the per-pair chain
    auto o_prefetch = [&](int pair) {
      const int b = pair / p.H, h = pair % p.H;
      if (r < p.Sq) {
        const bf16* op = p.o + ((int64_t)b * p.Sq + r) * p.ldo + h * 64;
        const uint32_t dst = smem_u32(obuf + r * AT_OROW);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * 16), "l"(op + q * 8) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float lse_next = 0.f;
    if (cg == 0 && (int)blockIdx.x < n_pairs) {
      o_prefetch(blockIdx.x);
      if (r < p.Sq) lse_next = p.lse[((int64_t)(blockIdx.x / p.H) * p.H + blockIdx.x % p.H) * p.Sq + r];
    }
    int it = 0;
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
      const int b = pair / p.H, h = pair % p.H, st = it & 1;
      uint32_t* km = sh.kmask[it & 1];
      float* Dsh = sh.red[it & 1][0];
      float* Lsh = sh.red[it & 1][1];
      const bool row_ok = r < p.Sq;
      if (cg == 0) {
        at_key_mask(p, b, km);
        const float nl2 = row_ok ? -lse_next * AT_LOG2E : INFINITY;       // absent / fully masked row -> P = dS = 0
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        mbar_wait(&sh.full[st], (it >> 1) & 1);
        float D = 0.f;
        if (row_ok) {
          const uint8_t* dot = stage0 + st * AT_BWD_STAGE + 3 * AT_TILE;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint4 a = *reinterpret_cast<const uint4*>(dot + at_sw(r, q));
            const uint4 o4 = *reinterpret_cast<const uint4*>(obuf + r * AT_OROW + q * 16);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              D = fmaf(__uint_as_float(aw[j] << 16), __uint_as_float(bw[j] << 16), D);
              D = fmaf(__uint_as_float(aw[j] & 0xffff0000u), __uint_as_float(bw[j] & 0xffff0000u), D);
            }
          }
        }
        Dsh[r] = D;
        Lsh[r] = nl2;
        // this thread's O row has been consumed (each thread reads only the row it copied): request the next pair's
        const int nxt = pair + gridDim.x;
        if (nxt < n_pairs) {
          o_prefetch(nxt);
          if (row_ok) lse_next = p.lse[((int64_t)(nxt / p.H) * p.H + nxt % p.H) * p.Sq + r];
        }
      }
      at_bar_compute();
      mbar_wait(&sh.s_full, it & 1);
      tc_fence_after();
      if (act) {
        const float D = Dsh[r], nl2 = Lsh[r];
        uint32_t s[32], dp[32];
        tmem_ld32(trow + cg * 32, s);
        tmem_ld32(trow + 128 + cg * 32, dp);
        tmem_ld_wait();
        uint32_t w = at_row_mask(km[cg], p.causal, r, cg);
        if (!(nl2 < INFINITY)) w = 0xffffffffu;
        uint32_t pw[16], dw[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float p0 = ((w >> j) & 1) ? 0.f : at_ex2(fmaf(__uint_as_float(s[j]), sc2, nl2));
          const float p1 = ((w >> (j + 1)) & 1) ? 0.f : at_ex2(fmaf(__uint_as_float(s[j + 1]), sc2, nl2));
          const float d0 = p0 * (__uint_as_float(dp[j]) - D) * p.scale;
          const float d1 = p1 * (__uint_as_float(dp[j + 1]) - D) * p.scale;
          pw[j >> 1] = at_pack(p0, p1);
          dw[j >> 1] = at_pack(d0, d1);
        }
        uint8_t* pblk = pbuf + (cg >> 1) * AT_TILE;
        uint8_t* dblk = dsbuf + (cg >> 1) * AT_TILE;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<uint4*>(pblk + at_sw(r, (cg & 1) * 4 + q)) = make_uint4(pw[4 * q], pw[4 * q + 1], pw[4 * q + 2], pw[4 * q + 3]);
          *reinterpret_cast<uint4*>(dblk + at_sw(r, (cg & 1) * 4 + q)) = make_uint4(dw[4 * q], dw[4 * q + 1], dw[4 * q + 2], dw[4 * q + 3]);
        }
      }
      tc_fence_before();
      at_fence_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.p_full);
      mbar_wait(&sh.o_full, it & 1);
      tc_fence_after();
      // outputs: six 32-column chunks over the four column groups: 0 -> dV[0:32], dQ[0:32]; 1 -> dV[32:64], dQ[32:64]; 2 -> dK[0:32]; 3 -> dK[32:64]
      {
        uint32_t v[32];
        const int half = cg & 1;
        if (cg < 2) {
          tmem_ld32(trow + 256 + half * 32, v);      // dV: thread = key row
          tmem_ld_wait();
          if (r < p.Sk) at_store_row32(p.dv + ((int64_t)b * p.Sk + r) * p.lddv + h * 64 + half * 32, v, 1.f);
          tmem_ld32(trow + 384 + half * 32, v);      // dQ: thread = query row
          tmem_ld_wait();
          if (row_ok) at_store_row32(p.dq + ((int64_t)b * p.Sq + r) * p.lddq + h * 64 + half * 32, v, 1.f);
        } else {
          tmem_ld32(trow + 320 + half * 32, v);      // dK: thread = key row
          tmem_ld_wait();
          if (r < p.Sk) at_store_row32(p.dk + ((int64_t)b * p.Sk + r) * p.lddk + h * 64 + half * 32, v, 1.f);
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_CW) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*AtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static AtEncodeFn at_encode_fn() {
  static AtEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<AtEncodeFn>(f);
  });
  return fn;
}
// [rows, cols] bf16 matrix with row pitch ld (elements), box = 64 columns (one head) x 128 rows, 128-byte swizzle
static int at_tmap(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  AtEncodeFn enc = at_encode_fn();
  if (!enc) { kmb_set_last_error("cuTensorMapEncodeTiled entry point unavailable", __FILE__, __LINE__); return KMB_ERR_TMAP; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "attention: cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu", (int)r, (unsigned long long)rows,
             (unsigned long long)cols, (unsigned long long)ld);
    kmb_set_last_error(msg, __FILE__, __LINE__);
    return KMB_ERR_TMAP;
  }
  return KMB_OK;
}

static int at_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Which attention launches take the tcgen05 kernels (both are parity-tested for every S <= 128):
//   KMBART_ATTN_TC05=1 all of them, =0 none; default: where they win on B200 (profiles/r02_attention_tc05.md) — the backward of
//   large tiles (encoder self-attention, 100 x 100: 78 vs 98 us).  One pair in flight per SM costs ~5 us of hand-offs
//   (TMA -> MMA -> tcgen05.ld -> smem -> MMA -> tcgen05.ld), which the mma.sync kernels beat on the small decoder shapes and
//   match in the forward pass.
bool attn_tc05_enabled(int is_bwd, int Sq, int Sk) {
  const char* e = getenv("KMBART_ATTN_TC05");
  if (e && e[0] == '0') return false;
  if (e && e[0] == '1') return true;
  return is_bwd && Sq >= 96 && Sk >= 96;
}

// token-major q / k / v / o ([B*S, ld], head h in columns [64h, 64h+64)); returns KMB_OK or an error (caller falls back on none)
int attn_fwd_tc05(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* o, int64_t ldo, float* lse,
                  const uint8_t* key_pad, int B, int H, int Sq, int Sk, int causal, float scale, cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = at_tmap(&tq, q, (uint64_t)B * Sq, (uint64_t)H * 64, ldq))) return rc;
  if ((rc = at_tmap(&tk, k, (uint64_t)B * Sk, (uint64_t)H * 64, ldk))) return rc;
  if ((rc = at_tmap(&tv, v, (uint64_t)B * Sk, (uint64_t)H * 64, ldv))) return rc;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_fwd_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_FWD_SMEM) != cudaSuccess) {
      kmb_set_last_error("attn_fwd_tc05_kernel: cannot reserve shared memory", __FILE__, __LINE__);
      return KMB_ERR_CUDA;
    }
    attr = true;
  }
  AtParams p = {};
  p.out = (bf16*)o; p.ldout = ldo; p.lse = lse; p.key_pad = key_pad; p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = causal; p.scale = scale;
  const int pairs = B * H, sms = at_num_sms();
  launch_pdl(attn_fwd_tc05_kernel, dim3(pairs < sms ? pairs : sms), dim3(AT_THREADS), (size_t)AT_FWD_SMEM, st, tq, tk, tv, p);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

int attn_bwd_tc05(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const void* o, int64_t ldo, const void* d_o,
                  int64_t lddo, const float* lse, const uint8_t* key_pad, void* dq, void* dk, void* dv, int64_t lddq, int64_t lddk, int64_t lddv,
                  int B, int H, int Sq, int Sk, int causal, float scale, cudaStream_t st) {
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = at_tmap(&tq, q, (uint64_t)B * Sq, (uint64_t)H * 64, ldq))) return rc;
  if ((rc = at_tmap(&tk, k, (uint64_t)B * Sk, (uint64_t)H * 64, ldk))) return rc;
  if ((rc = at_tmap(&tv, v, (uint64_t)B * Sk, (uint64_t)H * 64, ldv))) return rc;
  if ((rc = at_tmap(&tdo, d_o, (uint64_t)B * Sq, (uint64_t)H * 64, lddo))) return rc;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_bwd_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_BWD_SMEM) != cudaSuccess) {
      kmb_set_last_error("attn_bwd_tc05_kernel: cannot reserve shared memory", __FILE__, __LINE__);
      return KMB_ERR_CUDA;
    }
    attr = true;
  }
  AtParams p = {};
  p.o = (const bf16*)o; p.ldo = ldo; p.lse = const_cast<float*>(lse); p.key_pad = key_pad;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = causal; p.scale = scale;
  const int pairs = B * H, sms = at_num_sms();
  launch_pdl(attn_bwd_tc05_kernel, dim3(pairs < sms ? pairs : sms), dim3(AT_THREADS), (size_t)AT_BWD_SMEM, st, tq, tk, tv, tdo, p);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

}  // namespace kmb
