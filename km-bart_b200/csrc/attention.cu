// Fused multi-head attention (training / prefill), head_dim 64, bf16 in, fp32 softmax.
//
// Replaces HF-3.0.2 SelfAttention.forward as instantiated by the reference
// (src/model/modules.py:84 encoder self-attention; src/model/model.py:35 decoder self- and
// cross-attention; masks built at src/model/modules.py:130-131 and src/model/model.py:63-70):
//   w = (q*dh^-0.5) k^T ; w += causal ; w.masked_fill(key_padding, -inf) ; softmax ; w v
// The [B*h, S, S] score tensor of the reference never reaches HBM: scores live in registers
// (flash-style online softmax over 64-key blocks), only O and the row log-sum-exp are written.
// S is ~100 / ~48 on this path and attention is 1.7 % of the FLOPs, so the
// contractions run on mma.sync m16n8k16 (one warp = 16 query rows); the kernel's job is to
// remove the score round-trips, not to chase tcgen05 utilisation.
//
// Backward is two kernels that both recompute the scores: one owns query rows (dQ), the
// other owns key rows and works on the transposed problem (dK, dV); no atomics.
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

constexpr int TQ = 64;   // rows owned by a CTA (4 warps x 16)
constexpr int TK = 64;   // key / query block of one MMA pass
constexpr int SB = 128;  // rows of the streamed operand resident in shared memory at a time
constexpr int DH = 64;
constexpr int LDS = 72;  // padded smem row (elements) -> conflict-free ldmatrix

struct AttnParams {
  const bf16 *q, *k, *v;
  int64_t ldq, ldk, ldv;
  bf16* o;
  int64_t ldo;
  // element strides between batches / heads: token-major [B*S, H*64] uses (S*ld, 64),
  // the legacy cache layout [B, H, S, 64] uses (H*S*64, S*64) with ld = 64
  int64_t sbq, shq, sbk, shk, sbv, shv, sbo, sho;
  float* lse;              // [B, H, Sq]
  const uint8_t* key_pad;  // [B, Sk], 1 = padding, or null
  int B, H, Sq, Sk, causal;
  float scale;
  // backward
  const bf16* dO;
  int64_t lddo;
  float* D;  // [B, H, Sq]
  bf16 *dq, *dk, *dv;
  int64_t lddq, lddk, lddv;
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx_fwd(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// 64 x 64 bf16 tile, rows [r0, r0+64) of a [nrows, ld] matrix, zero-filled past nrows
__device__ __forceinline__ void load_tile(bf16* s, const bf16* g, int64_t ld, int r0, int nrows) {
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r0 + r < nrows) v = *reinterpret_cast<const uint4*>(g + (int64_t)(r0 + r) * ld + c);
    *reinterpret_cast<uint4*>(s + r * LDS + c) = v;
  }
}

// Asynchronous variant (cp.async, 16 B per request, zero-fill past nrows): `rows` rows starting at r0.
// All tiles a CTA needs are requested up front and awaited once, so a CTA has a single exposed
// global-memory latency instead of one per tile.
__device__ __forceinline__ void load_tile_async(bf16* s, const bf16* g, int64_t ld, int r0, int nrows, int rows) {
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    const bool ok = (r0 + r) < nrows;
    const bf16* src = ok ? g + (int64_t)(r0 + r) * ld + c : g;
    const uint32_t dst = smem_u32(s + r * LDS + c);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
  }
}
// Same copy with the per-thread (row, 16-byte chunk) pattern hoisted out of the loop: thread t owns chunk t & 7 of
// rows t >> 3, + nthreads / 8, ... so each request costs two pointer increments instead of 64-bit index arithmetic
// (the index math of load_tile_async was ~20 % of all instructions of the persistent kernels).
__device__ __forceinline__ void load_rows_async(bf16* s, const bf16* g, int64_t ld, int nrows, int rows, int tid, int nthreads) {
  const int c = (tid & 7) * 8, rstep = nthreads >> 3;
  int r = tid >> 3;
  const bf16* src = g + (int64_t)r * ld + c;
  uint32_t dst = smem_u32(s + r * LDS + c);
  const int64_t sstep = (int64_t)rstep * ld;
  const uint32_t dstep = (uint32_t)rstep * LDS * 2;
  for (; r < rows; r += rstep, src += sstep, dst += dstep) {
    const bool ok = r < nrows;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ok ? src : g), "r"(ok ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// A-operand fragments (16 rows x 64) for this warp's rows
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], const bf16* s, int warp, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldsm_x4(f[ks], s + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + ks * 16 + (lane >> 4) * 8);
}

// acc[16 x 64 cols] = A(16 x 64) * Y^T where Y is a [64 cols][64 k] smem tile (k contiguous)
__device__ __forceinline__ void mma_a_yt(float (&acc)[8][4], const uint32_t (&a)[4][4], const bf16* y, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(b, y + (np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8);
      mma16816(acc[2 * np], a[ks], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
}

// acc[16 x 64 dh] += P(16 x 64 k, C-fragment layout) * Y where Y is a [64 k][64 dh] smem tile
__device__ __forceinline__ void mma_p_y(float (&acc)[8][4], const float (&pm)[8][4], const bf16* y, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    a[0] = pack2(pm[2 * ks][0], pm[2 * ks][1]);
    a[1] = pack2(pm[2 * ks][2], pm[2 * ks][3]);
    a[2] = pack2(pm[2 * ks + 1][0], pm[2 * ks + 1][1]);
    a[3] = pack2(pm[2 * ks + 1][2], pm[2 * ks + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, y + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

__device__ __forceinline__ void store_rows(bf16* g, int64_t ld, int r_lo, int nrows, const float (&acc)[8][4], int lane,
                                           float s_lo, float s_hi) {
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + 2 * t;
    if (r_lo < nrows) *reinterpret_cast<uint32_t*>(g + (int64_t)r_lo * ld + c) = pack2(acc[nt][0] * s_lo, acc[nt][1] * s_lo);
    if (r_lo + 8 < nrows) *reinterpret_cast<uint32_t*>(g + (int64_t)(r_lo + 8) * ld + c) = pack2(acc[nt][2] * s_hi, acc[nt][3] * s_hi);
  }
}

// ------------------------------------------------------------------ forward
// NW warps per CTA = 16 * NW query rows: the launcher picks NW = ceil(Sq / 16) when one CTA can own every query row
// of a (batch, head) pair (Sq <= 128: K / V are then read once per pair and no warp idles on padding rows:
// Sq = 100 -> 7 warps, Sq = 48 -> 3), else 4-warp tiles of 64 rows.
template <int NW>
__global__ void __launch_bounds__(32 * NW) attn_fwd_kernel(const AttnParams p) {
  constexpr int TQ = 16 * NW;
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t dsm[];
  bf16* sQ = reinterpret_cast<bf16*>(dsm);
  bf16* sK = sQ + TQ * LDS;
  bf16* sV = sK + SB * LDS;
  uint32_t* sBits = reinterpret_cast<uint32_t*>(sV + SB * LDS);   // 4 words: bit k set = resident key k is padding / beyond Sk
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const bf16* qg = p.q + b * p.sbq + h * p.shq;
  const bf16* kg = p.k + b * p.sbk + h * p.shk;
  const bf16* vg = p.v + b * p.sbv + h * p.shv;
  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  // running row maximum of the RAW scores (scale > 0 is folded into the exponent: one FFMA + ex2 per element)
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float sl2 = p.scale * 1.4426950408889634f;
  const int r_lo = q0 + warp * 16 + g;
  const int w_row0 = q0 + warp * 16;     // first query row of this warp (causal tile test)
  int kend = p.Sk;
  if (p.causal && q0 + TQ < kend) kend = q0 + TQ;
  for (int kk0 = 0; kk0 < kend; kk0 += SB) {
    __syncthreads();
    if (kk0 == 0) load_rows_async(sQ, qg + (int64_t)q0 * p.ldq, p.ldq, p.Sq - q0, TQ, threadIdx.x, 32 * NW);
    load_rows_async(sK, kg + (int64_t)kk0 * p.ldk, p.ldk, p.Sk - kk0, SB, threadIdx.x, 32 * NW);
    load_rows_async(sV, vg + (int64_t)kk0 * p.ldv, p.ldv, p.Sk - kk0, SB, threadIdx.x, 32 * NW);
    for (int w = warp; w < SB / 32; w += NW) {   // one ballot per 32 resident keys
      const int c = kk0 + w * 32 + lane;
      const bool masked = c >= p.Sk || (p.key_pad && p.key_pad[(int64_t)b * p.Sk + c]);
      const uint32_t bits = __ballot_sync(0xffffffffu, masked);
      if (lane == 0) sBits[w] = bits;
    }
    cp_async_wait_all();
    __syncthreads();
    if (kk0 == 0) load_a_frags(qf, sQ, warp, lane);
    const int kstop = min(kend, kk0 + SB);
    for (int k0 = kk0; k0 < kstop; k0 += TK) {
      const bf16* sKb = sK + (k0 - kk0) * LDS;
      const bf16* sVb = sV + (k0 - kk0) * LDS;
      const uint32_t wbits[2] = {sBits[(k0 - kk0) >> 5], sBits[((k0 - kk0) >> 5) + 1]};
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      mma_a_yt(s, qf, sKb, lane);
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t byte = (wbits[nt >> 2] >> ((nt & 3) * 8)) & 0xFFu;
        if (byte != 0u || (p.causal && k0 + nt * 8 + 7 > w_row0)) {   // warp-uniform: clean tiles take no mask work
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int kl = 2 * t + (e & 1);
            const bool masked = ((byte >> kl) & 1u) || (p.causal && k0 + nt * 8 + kl > r_lo + (e >> 1) * 8);
            if (masked) s[nt][e] = -INFINITY;
          }
        }
        mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
      }
      float corr[2], nm2[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
        const float mnew = fmaxf(mrow[hh], mx[hh]);
        const bool dead = mnew == -INFINITY;          // every key so far masked: p = 0, nothing to rescale
        corr[hh] = dead ? 1.f : ex2_approx_fwd((mrow[hh] - mnew) * sl2);
        nm2[hh] = dead ? 0.f : -mnew * sl2;
        mrow[hh] = mnew;
        lrow[hh] *= corr[hh];
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int hh = e >> 1;
          const float pv = ex2_approx_fwd(fmaf(s[nt][e], sl2, nm2[hh]));   // s = -inf -> 0
          s[nt][e] = pv;
          lrow[hh] += pv;
          o[nt][e] *= corr[hh];
        }
      mma_p_y(o, s, sVb, lane);
    }
  }
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    lrow[hh] += __shfl_xor_sync(0xffffffffu, lrow[hh], 1);
    lrow[hh] += __shfl_xor_sync(0xffffffffu, lrow[hh], 2);
  }
  // a fully masked row yields NaN like the reference's softmax over all -inf
  const float inv_lo = 1.f / lrow[0], inv_hi = 1.f / lrow[1];
  bf16* og = p.o + b * p.sbo + h * p.sho;
  store_rows(og, p.ldo, r_lo, p.Sq, o, lane, inv_lo, inv_hi);
  if (p.lse && t == 0) {
    float* l = p.lse + ((int64_t)b * p.H + h) * p.Sq;
    if (r_lo < p.Sq) l[r_lo] = (mrow[0] * sl2 + log2f(lrow[0])) * 0.6931471805599453f;
    if (r_lo + 8 < p.Sq) l[r_lo + 8] = (mrow[1] * sl2 + log2f(lrow[1])) * 0.6931471805599453f;
  }
}

// ------------------------------------------------------------------ forward, persistent (Sq, Sk <= 128)
// One persistent 16-warp CTA per SM walks groups of G (batch, head) pairs.  Q, K, V of a group are fetched with
// cp.async into one of two input buffers while the previous group is computed, each pair's K / V is read from HBM
// exactly once (the tiled kernel above re-reads them per 64-row query tile and exposes one full load latency per
// CTA: 21 % of HBM peak at S = 100).  One warp owns 16 query rows of one pair: S = Q K^T for all keys stays in
// registers (<= 128 keys -> plain two-pass softmax, no online rescale), O = P V, O is staged through the warp's own
// (already consumed) Q rows in shared memory and leaves as full 128-byte lines.
constexpr int AF_THREADS = 512;
constexpr int AF_WARPS = AF_THREADS / 32;
constexpr int AF_MAXG = 5;

__host__ __device__ inline int af_unit_bytes(int SqP, int SkP) { return (SqP + 2 * SkP) * LDS * 2 + 16; }

template <int NKT>   // 16-key tiles
__device__ __forceinline__ void af_compute(const AttnParams& p, bf16* sQ, const bf16* sK, const bf16* sV, const uint32_t* bits,
                                           int qt, int b, int h, int lane) {
  const int g = lane >> 2, t = lane & 3;
  uint32_t qf[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldsm_x4(qf[ks], sQ + (qt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + ks * 16 + (lane >> 4) * 8);
  float s[2 * NKT][4];
#pragma unroll
  for (int i = 0; i < 2 * NKT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
  for (int np = 0; np < NKT; ++np)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t bk[4];
      ldsm_x4(bk, sK + (np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8);
      mma16816(s[2 * np], qf[ks], bk[0], bk[1]);
      mma16816(s[2 * np + 1], qf[ks], bk[2], bk[3]);
    }
  // masks: padding / beyond Sk from the bit words, causal from the row index.  An 8-key tile whose mask byte is
  // clear (and that lies at or below the diagonal for causal attention) takes no per-element work at all.
  const int r_lo = qt * 16 + g;
  const float sl2 = p.scale * 1.4426950408889634f;   // scale > 0 (checked on the host)
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < 2 * NKT; ++nt) {
    const uint32_t byte = (bits[nt >> 2] >> ((nt & 3) * 8)) & 0xFFu;
    if (byte != 0u || (p.causal && nt * 8 + 7 > qt * 16)) {   // warp-uniform
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kl = 2 * t + (e & 1);
        const bool masked = ((byte >> kl) & 1u) || (p.causal && nt * 8 + kl > r_lo + (e >> 1) * 8);
        if (masked) s[nt][e] = -INFINITY;
      }
    }
    mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
    mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
  }
  float lsum[2] = {0.f, 0.f}, nm2[2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
    mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
    nm2[hh] = (mx[hh] == -INFINITY) ? 0.f : -mx[hh] * sl2;   // fully masked row: every p = ex2(-inf) = 0
  }
#pragma unroll
  for (int nt = 0; nt < 2 * NKT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float pv = ex2_approx_fwd(fmaf(s[nt][e], sl2, nm2[e >> 1]));
      s[nt][e] = pv;
      lsum[e >> 1] += pv;
    }
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    lsum[hh] += __shfl_xor_sync(0xffffffffu, lsum[hh], 1);
    lsum[hh] += __shfl_xor_sync(0xffffffffu, lsum[hh], 2);
  }
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < NKT; ++ks) {
    uint32_t a[4];
    a[0] = pack2(s[2 * ks][0], s[2 * ks][1]);
    a[1] = pack2(s[2 * ks][2], s[2 * ks][3]);
    a[2] = pack2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
    a[3] = pack2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bv[4];
      ldsm_x4_t(bv, sV + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8);
      mma16816(o[2 * np], a, bv[0], bv[1]);
      mma16816(o[2 * np + 1], a, bv[2], bv[3]);
    }
  }
  // a fully masked row yields NaN like the reference's softmax over all -inf (0 * inf)
  const float inv[2] = {1.f / lsum[0], 1.f / lsum[1]};
  __syncwarp();   // every lane has fetched its Q fragments: the warp's 16 Q rows become the O staging tile
  bf16* so = sQ + qt * 16 * LDS;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(so + g * LDS + nt * 8 + 2 * t) = pack2(o[nt][0] * inv[0], o[nt][1] * inv[0]);
    *reinterpret_cast<uint32_t*>(so + (g + 8) * LDS + nt * 8 + 2 * t) = pack2(o[nt][2] * inv[1], o[nt][3] * inv[1]);
  }
  __syncwarp();
  bf16* og = p.o + b * p.sbo + h * p.sho;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 4 + (lane >> 3), c = (lane & 7) * 8;
    const int row = qt * 16 + r;
    if (row < p.Sq) *reinterpret_cast<uint4*>(og + (int64_t)row * p.ldo + c) = *reinterpret_cast<const uint4*>(so + r * LDS + c);
  }
  if (p.lse && t == 0) {
    float* l = p.lse + ((int64_t)b * p.H + h) * p.Sq;
    // natural-log lse of the scaled scores: (max + log2(sum)) / log2(e)
    if (r_lo < p.Sq) l[r_lo] = (mx[0] * sl2 + log2f(lsum[0])) * 0.6931471805599453f;
    if (r_lo + 8 < p.Sq) l[r_lo + 8] = (mx[1] * sl2 + log2f(lsum[1])) * 0.6931471805599453f;
  }
}

__global__ void __launch_bounds__(AF_THREADS, 1) attn_fwd_persist_kernel(const AttnParams p, int SqP, int SkP, int G) {
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t dsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit_bytes = af_unit_bytes(SqP, SkP);
  const int buf_bytes = G * unit_bytes;
  const int n_tiles = p.B * p.H;
  const int n_groups = (n_tiles + G - 1) / G;
  const int nQ = SqP >> 4, nKT = SkP >> 4;
  pdl_wait();

  auto issue = [&](int group, int buf) {
    for (int gi = 0; gi < G; ++gi) {
      const int tile = group * G + gi;
      if (tile >= n_tiles) break;
      const int b = tile / p.H, h = tile % p.H;
      bf16* sQ = reinterpret_cast<bf16*>(dsm + buf * buf_bytes + gi * unit_bytes);
      bf16* sK = sQ + SqP * LDS;
      bf16* sV = sK + SkP * LDS;
      load_rows_async(sQ, p.q + b * p.sbq + h * p.shq, p.ldq, p.Sq, SqP, threadIdx.x, AF_THREADS);
      load_rows_async(sK, p.k + b * p.sbk + h * p.shk, p.ldk, p.Sk, SkP, threadIdx.x, AF_THREADS);
      load_rows_async(sV, p.v + b * p.sbv + h * p.shv, p.ldv, p.Sk, SkP, threadIdx.x, AF_THREADS);
    }
    for (int w = warp; w < 4 * G; w += AF_WARPS) {   // key-mask words: bit k of word k/32 set = key k is padding / beyond Sk
      const int gi = w >> 2, key = (w & 3) * 32 + lane;
      const int tile = group * G + gi;
      bool masked = true;
      if (tile < n_tiles) masked = key >= p.Sk || (p.key_pad && p.key_pad[(int64_t)(tile / p.H) * p.Sk + key]);
      const uint32_t bits = __ballot_sync(0xffffffffu, masked);
      if (lane == 0) reinterpret_cast<uint32_t*>(dsm + buf * buf_bytes + gi * unit_bytes + (SqP + 2 * SkP) * LDS * 2)[w & 3] = bits;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if ((int)blockIdx.x < n_groups) issue(blockIdx.x, 0);
  int it = 0;
  for (int group = blockIdx.x; group < n_groups; group += gridDim.x, ++it) {
    const int buf = it & 1;
    const bool has_next = group + (int)gridDim.x < n_groups;
    if (has_next) {
      issue(group + gridDim.x, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int gi = warp / nQ, qt = warp - gi * nQ;
    const int tile = group * G + gi;
    if (gi < G && tile < n_tiles) {
      bf16* sQ = reinterpret_cast<bf16*>(dsm + buf * buf_bytes + gi * unit_bytes);
      const bf16* sK = sQ + SqP * LDS;
      const bf16* sV = sK + SkP * LDS;
      const uint32_t* bits = reinterpret_cast<const uint32_t*>(dsm + buf * buf_bytes + gi * unit_bytes + (SqP + 2 * SkP) * LDS * 2);
      const int b = tile / p.H, h = tile % p.H;
      switch (nKT) {
        case 1: af_compute<1>(p, sQ, sK, sV, bits, qt, b, h, lane); break;
        case 2: af_compute<2>(p, sQ, sK, sV, bits, qt, b, h, lane); break;
        case 3: af_compute<3>(p, sQ, sK, sV, bits, qt, b, h, lane); break;
        case 4: af_compute<4>(p, sQ, sK, sV, bits, qt, b, h, lane); break;
        case 5: af_compute<5>(p, sQ, sK, sV, bits, qt, b, h, lane); break;
        case 6: af_compute<6>(p, sQ, sK, sV, bits, qt, b, h, lane); break;
        case 7: af_compute<7>(p, sQ, sK, sV, bits, qt, b, h, lane); break;
        default: af_compute<8>(p, sQ, sK, sV, bits, qt, b, h, lane); break;
      }
    }
    __syncthreads();   // the buffer is refilled by the issue() of the next iteration
  }
}

// ------------------------------------------------------------------ backward prep: D = rowsum(dO * O)
__global__ void attn_bwd_prep_kernel(const AttnParams p) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = p.B * p.H * p.Sq;
  if (warp >= total) return;
  const int i = warp % p.Sq, h = (warp / p.Sq) % p.H, b = warp / (p.Sq * p.H);
  const bf16* o = p.o + ((int64_t)b * p.Sq + i) * p.ldo + h * DH;
  const bf16* d = p.dO + ((int64_t)b * p.Sq + i) * p.lddo + h * DH;
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(o + 2 * lane);
  const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(d + 2 * lane);
  float v = __bfloat162float(a.x) * __bfloat162float(c.x) + __bfloat162float(a.y) * __bfloat162float(c.y);
  v = warp_sum(v);
  if (lane == 0) p.D[warp] = v;
}

// ------------------------------------------------------------------ backward, query-row owner: dQ
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t dsm[];
  bf16* sQ = reinterpret_cast<bf16*>(dsm);
  bf16* sdO = sQ + TQ * LDS;
  bf16* sK = sdO + TQ * LDS;
  bf16* sV = sK + SB * LDS;
  uint32_t* sBits = reinterpret_cast<uint32_t*>(sV + SB * LDS);   // bit k set = resident key k is padding / beyond Sk
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const bf16* qg = p.q + b * p.sbq + h * p.shq;
  const bf16* kg = p.k + b * p.sbk + h * p.shk;
  const bf16* vg = p.v + b * p.sbv + h * p.shv;
  const bf16* dog = p.dO + (int64_t)b * p.Sq * p.lddo + h * DH;
  uint32_t qf[4][4], dof[4][4];
  const int r_lo = q0 + warp * 16 + g;
  const int w_row0 = q0 + warp * 16;
  const float* lse = p.lse + ((int64_t)b * p.H + h) * p.Sq;
  const float* Dg = p.D + ((int64_t)b * p.H + h) * p.Sq;
  // per-row constants in the log2 domain: P = ex2(s*scale*log2e + nl); a dead row (lse = -inf) gets nl = -inf -> P = 0
  const float sl2 = p.scale * 1.4426950408889634f;
  float nl[2], dsc[2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int r = r_lo + hh * 8;
    const float l = r < p.Sq ? lse[r] : -INFINITY;
    nl[hh] = l == -INFINITY ? -INFINITY : -l * 1.4426950408889634f;
    dsc[hh] = (r < p.Sq ? Dg[r] : 0.f) * p.scale;
  }
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  int kend = p.Sk;
  if (p.causal && q0 + TQ < kend) kend = q0 + TQ;
  for (int kk0 = 0; kk0 < kend; kk0 += SB) {
    __syncthreads();
    if (kk0 == 0) {
      load_rows_async(sQ, qg + (int64_t)q0 * p.ldq, p.ldq, p.Sq - q0, TQ, threadIdx.x, 128);
      load_rows_async(sdO, dog + (int64_t)q0 * p.lddo, p.lddo, p.Sq - q0, TQ, threadIdx.x, 128);
    }
    load_rows_async(sK, kg + (int64_t)kk0 * p.ldk, p.ldk, p.Sk - kk0, SB, threadIdx.x, 128);
    load_rows_async(sV, vg + (int64_t)kk0 * p.ldv, p.ldv, p.Sk - kk0, SB, threadIdx.x, 128);
    {
      const int c = kk0 + threadIdx.x;
      const bool masked = c >= p.Sk || (p.key_pad && p.key_pad[(int64_t)b * p.Sk + c]);
      const uint32_t bits = __ballot_sync(0xffffffffu, masked);
      if (lane == 0) sBits[warp] = bits;
    }
    cp_async_wait_all();
    __syncthreads();
    if (kk0 == 0) {
      load_a_frags(qf, sQ, warp, lane);
      load_a_frags(dof, sdO, warp, lane);
    }
    const int kstop = min(kend, kk0 + SB);
    for (int k0 = kk0; k0 < kstop; k0 += TK) {
      const bf16* sKb = sK + (k0 - kk0) * LDS;
      const bf16* sVb = sV + (k0 - kk0) * LDS;
      const uint32_t wbits[2] = {sBits[(k0 - kk0) >> 5], sBits[((k0 - kk0) >> 5) + 1]};
      float s[8][4], dp[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      }
      mma_a_yt(s, qf, sKb, lane);
      mma_a_yt(dp, dof, sVb, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t byte = (wbits[nt >> 2] >> ((nt & 3) * 8)) & 0xFFu;
        if (byte != 0u || (p.causal && k0 + nt * 8 + 7 > w_row0)) {   // warp-uniform
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int kl = 2 * t + (e & 1);
            if (((byte >> kl) & 1u) || (p.causal && k0 + nt * 8 + kl > r_lo + (e >> 1) * 8)) s[nt][e] = -INFINITY;
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int hh = e >> 1;
          const float pv = ex2_approx_fwd(fmaf(s[nt][e], sl2, nl[hh]));
          s[nt][e] = pv * fmaf(dp[nt][e], p.scale, -dsc[hh]);   // dS
        }
      }
      mma_p_y(dq, s, sKb, lane);
    }
  }
  bf16* dqg = p.dq + (int64_t)b * p.Sq * p.lddq + h * DH;
  store_rows(dqg, p.lddq, r_lo, p.Sq, dq, lane, 1.f, 1.f);
}

// ------------------------------------------------------------------ backward, key-row owner: dK, dV
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t dsm[];
  bf16* sKr = reinterpret_cast<bf16*>(dsm);  // owned key rows: K and V
  bf16* sVr = sKr + TK * LDS;
  bf16* sQ = sVr + TK * LDS;                  // streamed query rows: Q and dO
  bf16* sdO = sQ + SB * LDS;
  float* sL = reinterpret_cast<float*>(sdO + SB * LDS);
  float* sD = sL + SB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int k0 = blockIdx.x * TK, h = blockIdx.y, b = blockIdx.z;
  const bf16* qg = p.q + b * p.sbq + h * p.shq;
  const bf16* kg = p.k + b * p.sbk + h * p.shk;
  const bf16* vg = p.v + b * p.sbv + h * p.shv;
  const bf16* dog = p.dO + (int64_t)b * p.Sq * p.lddo + h * DH;
  const float* lse = p.lse + ((int64_t)b * p.H + h) * p.Sq;
  const float* Dg = p.D + ((int64_t)b * p.H + h) * p.Sq;
  uint32_t kf[4][4], vf[4][4];
  const int r_lo = k0 + warp * 16 + g;  // key index of this thread's rows
  const float sl2 = p.scale * 1.4426950408889634f;
  bool rpad[2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int r = r_lo + hh * 8;
    rpad[hh] = (r >= p.Sk) || (p.key_pad && p.key_pad[(int64_t)b * p.Sk + r]);
  }
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  const int qstart = p.causal ? (k0 / TQ) * TQ : 0;  // queries before the key block see none of it
  for (int qq0 = qstart; qq0 < p.Sq; qq0 += SB) {
    __syncthreads();
    if (qq0 == qstart) {
      load_rows_async(sKr, kg + (int64_t)k0 * p.ldk, p.ldk, p.Sk - k0, TK, threadIdx.x, 128);
      load_rows_async(sVr, vg + (int64_t)k0 * p.ldv, p.ldv, p.Sk - k0, TK, threadIdx.x, 128);
    }
    load_rows_async(sQ, qg + (int64_t)qq0 * p.ldq, p.ldq, p.Sq - qq0, SB, threadIdx.x, 128);
    load_rows_async(sdO, dog + (int64_t)qq0 * p.lddo, p.lddo, p.Sq - qq0, SB, threadIdx.x, 128);
    if (threadIdx.x < SB) {   // per-query constants in the log2 domain: -lse*log2e (-inf for dead / absent rows), D*scale
      const int r = qq0 + threadIdx.x;
      const float l = r < p.Sq ? lse[r] : -INFINITY;
      sL[threadIdx.x] = l == -INFINITY ? -INFINITY : -l * 1.4426950408889634f;
      sD[threadIdx.x] = (r < p.Sq ? Dg[r] : 0.f) * p.scale;
    }
    cp_async_wait_all();
    __syncthreads();
    if (qq0 == qstart) {
      load_a_frags(kf, sKr, warp, lane);
      load_a_frags(vf, sVr, warp, lane);
    }
    const int qstop = min(p.Sq, qq0 + SB);
    for (int q0 = qq0; q0 < qstop; q0 += TQ) {
      const bf16* sQb = sQ + (q0 - qq0) * LDS;
      const bf16* sdOb = sdO + (q0 - qq0) * LDS;
      const float* sLb = sL + (q0 - qq0);
      const float* sDb = sD + (q0 - qq0);
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
      }
      mma_a_yt(st, kf, sQb, lane);     // S^T = K Q^T
      mma_a_yt(dpt, vf, sdOb, lane);   // dP^T = V dO^T
      float ds[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 nl2 = *reinterpret_cast<const float2*>(sLb + nt * 8 + 2 * t);   // queries nt*8 + 2t, +1
        const float2 dd2 = *reinterpret_cast<const float2*>(sDb + nt * 8 + 2 * t);
        const bool diag = p.causal && k0 + warp * 16 + 15 > q0 + nt * 8;             // warp-uniform: tile touches the diagonal
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int hh = e >> 1;
          const float nlq = (e & 1) ? nl2.y : nl2.x;
          const float ddq = (e & 1) ? dd2.y : dd2.x;
          bool masked = rpad[hh];
          if (diag) masked = masked || (r_lo + hh * 8 > q0 + nt * 8 + 2 * t + (e & 1));
          const float pv = masked ? 0.f : ex2_approx_fwd(fmaf(st[nt][e], sl2, nlq));
          st[nt][e] = pv;                                              // P^T
          ds[nt][e] = pv * fmaf(dpt[nt][e], p.scale, -ddq);            // dS^T
        }
      }
      mma_p_y(dv, st, sdOb, lane);
      mma_p_y(dk, ds, sQb, lane);
    }
  }
  bf16* dkg = p.dk + (int64_t)b * p.Sk * p.lddk + h * DH;
  bf16* dvg = p.dv + (int64_t)b * p.Sk * p.lddv + h * DH;
  store_rows(dkg, p.lddk, r_lo, p.Sk, dk, lane, 1.f, 1.f);
  store_rows(dvg, p.lddv, r_lo, p.Sk, dv, lane, 1.f, 1.f);
}

// ------------------------------------------------------------------ backward, fused (Sq, Sk <= 128)
// One persistent CTA per SM walks the (batch, head) pairs.  Q, dO, K, V of a pair (<= 72 KB) are fetched with
// cp.async into one of two input buffers while the previous pair is being computed, so HBM latency is hidden
// without relying on co-resident CTAs.  Per pair:
//   phase 1  units (16 query rows x 32 keys):  S = QK^T, dP = dO V^T, P = exp(S*scale - lse),
//            dS = P (dP - D) * scale with D = rowsum(dO o O); P and dS are parked in shared memory as bf16
//   phase 2  units: key-row owners  dV = P^T dO, dK = dS^T Q  (A operands read transposed with ldmatrix.trans);
//            query-row owners  dQ = dS K
// Scores are computed once (the split kernels below recompute them for dQ and again for dK/dV) and no operand is
// loaded twice.  Same rounding points as the split path: P and dS are rounded to bf16 before the second GEMMs.
constexpr int FB_THREADS = 512;
constexpr int FB_WARPS = FB_THREADS / 32;
constexpr int FB_MAXS = 128;
constexpr int FB_MAXG = 3;            // (batch, head) pairs processed together when the shapes are small
static_assert(FB_MAXG <= 3, "the unit -> pair split in attn_bwd_fused_kernel compares against U and 2U");
constexpr int FB_SMEM_MAX = 232448;   // 227 KB opt-in limit

__host__ __device__ inline int fb_tile_in_bytes(int SqP, int SkP) { return (2 * SqP + 2 * SkP) * LDS * 2; }
// one input buffer: G x (Q, dO, K, V) | G x 4 key-mask words | G x SqP D | G x SqP lse
__host__ __device__ inline int fb_buf_bytes(int SqP, int SkP, int G) { return G * (fb_tile_in_bytes(SqP, SkP) + 16 + SqP * 8); }
__host__ __device__ inline int fb_smem_bytes(int SqP, int SkP, int G) {
  return 2 * fb_buf_bytes(SqP, SkP, G) + G * 2 * SqP * (SkP + 8) * 2;
}

__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b) {
  const uint32_t x[4] = {a.x, a.y, a.z, a.w}, y[4] = {b.x, b.y, b.z, b.w};
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v = fmaf(__uint_as_float(x[i] << 16), __uint_as_float(y[i] << 16), v);
    v = fmaf(__uint_as_float(x[i] & 0xFFFF0000u), __uint_as_float(y[i] & 0xFFFF0000u), v);
  }
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// out[16 rows x 32 dh-columns (half `hf`)] = A(16 x 16*nks) * Y, A fragments fetched by `afrag(ks, a)`,
// Y a [k][64] shared-memory tile (row pitch LDS); result stored as bf16 rows r_lo / r_lo + 8 of g
template <typename AFrag>
__device__ __forceinline__ void fb_gemm_half(AFrag afrag, int nks, const bf16* y, int hf, bf16* g, int64_t ld, int r_lo,
                                             int nrows, int lane) {
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int ks = 0; ks < nks; ++ks) {
    uint32_t a[4];
    afrag(ks, a);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t bb[4];
      ldsm_x4_t(bb, y + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + hf * 32 + np * 16 + (lane >> 4) * 8);
      mma16816(acc[2 * np], a, bb[0], bb[1]);
      mma16816(acc[2 * np + 1], a, bb[2], bb[3]);
    }
  }
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int c = hf * 32 + nt * 8 + 2 * t;
    if (r_lo < nrows) *reinterpret_cast<uint32_t*>(g + (int64_t)r_lo * ld + c) = pack2(acc[nt][0], acc[nt][1]);
    if (r_lo + 8 < nrows) *reinterpret_cast<uint32_t*>(g + (int64_t)(r_lo + 8) * ld + c) = pack2(acc[nt][2], acc[nt][3]);
  }
}

// D = rowsum(dO o O) and lse of one group, staged through registers: the loads of the NEXT group are issued before
// phase 2 of the current one and consumed after it, so their latency never sits on the critical path.
struct FbRowPrefetch {
  uint4 o[2][2], d[2][2];
  float l[2];
};

// SQP / SKP / GT: padded sequence lengths and pairs per iteration as compile-time constants (0 = take the run-time
// arguments).  The unit loops are dominated by ldmatrix address arithmetic on SqP, SkP + 8 and the unit splits
// (ncu: IMAD + IADD3 + LOP3 = 39 % of executed instructions, HMMA 6 %); the model's three shapes are instantiated.
template <int SQP, int SKP, int GT>
__global__ void __launch_bounds__(FB_THREADS, 1) attn_bwd_fused_kernel(const AttnParams p, int SqP_rt, int SkP_rt, int G_rt) {
  const int SqP = SQP ? SQP : SqP_rt, SkP = SKP ? SKP : SkP_rt, G = GT ? GT : G_rt;
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t dsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int LDP = SkP + 8;
  const int tile_in = fb_tile_in_bytes(SqP, SkP);
  const int buf_bytes = fb_buf_bytes(SqP, SkP, G);
  bf16* sPall = reinterpret_cast<bf16*>(dsm + 2 * buf_bytes);
  const int n_tiles = p.B * p.H;
  const int n_groups = (n_tiles + G - 1) / G;
  const int nQ = SqP >> 4, nK = SkP >> 4, nKC = (SkP + 31) >> 5;
  const int nRG = SqP >> 3;   // 8-row groups per tile
  // reciprocals for the per-unit index splits (exact for the small operands used here: uu < 2^10, tile < 2^31 / H)
  const uint32_t inv_nKC = (65536u + nKC - 1) / nKC;
  const uint64_t inv_H = ((1ull << 32) + p.H - 1) / p.H;
  pdl_wait();

  auto buf_bits = [&](int buf) { return reinterpret_cast<uint32_t*>(dsm + buf * buf_bytes + G * tile_in); };
  auto buf_D = [&](int buf) { return reinterpret_cast<float*>(dsm + buf * buf_bytes + G * (tile_in + 16)); };
  auto buf_L = [&](int buf) { return buf_D(buf) + G * SqP; };

  auto issue = [&](int group, int buf) {
    for (int gi = 0; gi < G; ++gi) {
      const int tile = group * G + gi;
      if (tile >= n_tiles) break;
      const int b = tile / p.H, h = tile % p.H;
      bf16* sQ = reinterpret_cast<bf16*>(dsm + buf * buf_bytes + gi * tile_in);
      bf16* sdO = sQ + SqP * LDS;
      bf16* sK = sdO + SqP * LDS;
      bf16* sV = sK + SkP * LDS;
      load_rows_async(sQ, p.q + b * p.sbq + h * p.shq, p.ldq, p.Sq, SqP, threadIdx.x, FB_THREADS);
      load_rows_async(sdO, p.dO + (int64_t)b * p.Sq * p.lddo + h * DH, p.lddo, p.Sq, SqP, threadIdx.x, FB_THREADS);
      load_rows_async(sK, p.k + b * p.sbk + h * p.shk, p.ldk, p.Sk, SkP, threadIdx.x, FB_THREADS);
      load_rows_async(sV, p.v + b * p.sbv + h * p.shv, p.ldv, p.Sk, SkP, threadIdx.x, FB_THREADS);
    }
    if (warp < 4 * G) {   // key-mask words: bit k of word k/32 set = key k is padding / beyond Sk
      const int gi = warp >> 2, key = (warp & 3) * 32 + lane;
      const int tile = group * G + gi;
      bool masked = true;
      if (tile < n_tiles) masked = key >= p.Sk || (p.key_pad && p.key_pad[(int64_t)(tile / p.H) * p.Sk + key]);
      const uint32_t bits = __ballot_sync(0xffffffffu, masked);
      if (lane == 0) buf_bits(buf)[warp] = bits;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  auto rows_load = [&](int group, FbRowPrefetch& r) {
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const int rg = warp + ps * FB_WARPS;
      const int gi = rg / nRG, row = (rg % nRG) * 8 + g;
      const int tile = group * G + gi;
      r.o[ps][0] = r.o[ps][1] = r.d[ps][0] = r.d[ps][1] = make_uint4(0, 0, 0, 0);
      r.l[ps] = -INFINITY;
      if (rg < G * nRG && tile < n_tiles && row < p.Sq) {
        const int b = tile / p.H, h = tile % p.H;
        const bf16* op = p.o + b * p.sbo + h * p.sho + (int64_t)row * p.ldo + t * 16;
        const bf16* dp = p.dO + ((int64_t)b * p.Sq + row) * p.lddo + h * DH + t * 16;
        r.o[ps][0] = ldg_nc16(op);
        r.o[ps][1] = ldg_nc16(op + 8);
        r.d[ps][0] = ldg_nc16(dp);
        r.d[ps][1] = ldg_nc16(dp + 8);
        r.l[ps] = __ldg(p.lse + ((int64_t)b * p.H + h) * p.Sq + row);
      }
    }
  };
  auto rows_store = [&](const FbRowPrefetch& r, int buf) {
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const int rg = warp + ps * FB_WARPS;
      float v = dot8_bf16(r.o[ps][0], r.d[ps][0]) + dot8_bf16(r.o[ps][1], r.d[ps][1]);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (rg < G * nRG && t == 0) {
        buf_D(buf)[rg * 8 + g] = v;          // rg * 8 + g == gi * SqP + row
        buf_L(buf)[rg * 8 + g] = r.l[ps];
      }
    }
  };

  const float sl2 = p.scale * 1.4426950408889634f;
  FbRowPrefetch pre;
  if ((int)blockIdx.x < n_groups) {
    issue(blockIdx.x, 0);
    rows_load(blockIdx.x, pre);
    rows_store(pre, 0);
  }
  int it = 0;
  for (int group = blockIdx.x; group < n_groups; group += gridDim.x, ++it) {
    const int buf = it & 1;
    const bool has_next = group + (int)gridDim.x < n_groups;
    if (has_next) {
      issue(group + gridDim.x, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint8_t* bbase = dsm + buf * buf_bytes;
    const uint32_t* sBits = buf_bits(buf);
    const float* sD = buf_D(buf);
    const float* sL = buf_L(buf);

    // ---------------- phase 1: P and dS tiles (16 query rows x 32 keys per unit)
    const int U1 = nQ * nKC;
    for (int u = warp; u < G * U1; u += FB_WARPS) {
      const int gi = (u >= U1) + (u >= 2 * U1), uu = u - gi * U1;   // G <= FB_MAXG = 3: no integer division
      if (group * G + gi >= n_tiles) break;
      const bf16* sQ = reinterpret_cast<const bf16*>(bbase + gi * tile_in);
      const bf16* sdO = sQ + SqP * LDS;
      const bf16* sK = sdO + SqP * LDS;
      const bf16* sV = sK + SkP * LDS;
      bf16* sP = sPall + gi * 2 * SqP * LDP;
      bf16* sdS = sP + SqP * LDP;
      const int qt = (int)(((uint32_t)uu * inv_nKC) >> 16), kc = uu - qt * nKC, k0 = kc * 32;
      const int npmax = min(2, (SkP - k0) >> 4);
      const int r_lo = qt * 16 + g;
      float sc[4][4], dp[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t qf[4], dof[4];
        const int aoff = (qt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + ks * 16 + (lane >> 4) * 8;
        ldsm_x4(qf, sQ + aoff);
        ldsm_x4(dof, sdO + aoff);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          if (np < npmax) {
            const int boff = (k0 + np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8;
            uint32_t bk[4], bv[4];
            ldsm_x4(bk, sK + boff);
            ldsm_x4(bv, sV + boff);
            mma16816(sc[2 * np], qf, bk[0], bk[1]);
            mma16816(sc[2 * np + 1], qf, bk[2], bk[3]);
            mma16816(dp[2 * np], dof, bv[0], bv[1]);
            mma16816(dp[2 * np + 1], dof, bv[2], bv[3]);
          }
        }
      }
      // per-row constants: -lse*log2(e), D*scale, and a 32-key mask word (padding | causal | dead row) shifted by 2t
      const uint32_t kmask = sBits[gi * 4 + kc];
      float nl[2], dsc[2];
      uint32_t rm[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int row = r_lo + hh * 8;
        const float l = sL[gi * SqP + row];
        const bool dead = l == -INFINITY;
        nl[hh] = dead ? 0.f : -l * 1.4426950408889634f;
        dsc[hh] = sD[gi * SqP + row] * p.scale;
        uint32_t m = dead ? 0xFFFFFFFFu : kmask;
        if (p.causal) {
          const int sft = row - k0 + 1;   // keys k0 + kl with kl >= sft lie in the future of `row`
          m |= sft <= 0 ? 0xFFFFFFFFu : (sft >= 32 ? 0u : (0xFFFFFFFFu << sft));
        }
        rm[hh] = m >> (2 * t);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (nt < 2 * npmax) {
          float pv[4], dsv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int hh = e >> 1;
            const bool masked = (rm[hh] >> (nt * 8 + (e & 1))) & 1u;
            pv[e] = masked ? 0.f : ex2_approx(fmaf(sc[nt][e], sl2, nl[hh]));
            dsv[e] = pv[e] * fmaf(dp[nt][e], p.scale, -dsc[hh]);
          }
          const int off = r_lo * LDP + k0 + nt * 8 + 2 * t;
          *reinterpret_cast<uint32_t*>(sP + off) = pack2(pv[0], pv[1]);
          *reinterpret_cast<uint32_t*>(sP + off + 8 * LDP) = pack2(pv[2], pv[3]);
          *reinterpret_cast<uint32_t*>(sdS + off) = pack2(dsv[0], dsv[1]);
          *reinterpret_cast<uint32_t*>(sdS + off + 8 * LDP) = pack2(dsv[2], dsv[3]);
        }
      }
    }
    if (has_next) rows_load(group + gridDim.x, pre);
    __syncthreads();

    // ---------------- phase 2: units of 16 rows x 32 head-dim columns: dV, dK per key tile, dQ per query tile
    const int U2 = 2 * (2 * nK + nQ);
    for (int u = warp; u < G * U2; u += FB_WARPS) {
      const int gi = (u >= U2) + (u >= 2 * U2), uu = u - gi * U2;
      const int tile = group * G + gi;
      if (tile >= n_tiles) break;
      const int b = (int)(((uint64_t)(uint32_t)tile * inv_H) >> 32), h = tile - b * p.H;
      const bf16* sQ = reinterpret_cast<const bf16*>(bbase + gi * tile_in);
      const bf16* sdO = sQ + SqP * LDS;
      const bf16* sK = sdO + SqP * LDS;
      const bf16* sP = sPall + gi * 2 * SqP * LDP;
      const bf16* sdS = sP + SqP * LDP;
      const int hf = uu & 1, v = uu >> 1;
      if (v < 2 * nK) {
        const int j0 = (v >> 1) * 16;
        const bf16* sA = (v & 1) ? sdS : sP;            // dK = dS^T Q ; dV = P^T dO
        const bf16* sY = (v & 1) ? sQ : sdO;
        bf16* dst = (v & 1) ? p.dk + (int64_t)b * p.Sk * p.lddk + h * DH : p.dv + (int64_t)b * p.Sk * p.lddv + h * DH;
        const int64_t ld = (v & 1) ? p.lddk : p.lddv;
        const bf16* abase = sA + ((lane & 7) + (lane >> 4) * 8) * LDP + j0 + ((lane >> 3) & 1) * 8;
        fb_gemm_half([&](int ks, uint32_t (&a)[4]) { ldsm_x4_t(a, abase + ks * 16 * LDP); }, nQ, sY, hf, dst, ld, j0 + g, p.Sk, lane);
      } else {
        const int qt = v - 2 * nK;
        const bf16* abase = sdS + (qt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDP + (lane >> 4) * 8;
        fb_gemm_half([&](int ks, uint32_t (&a)[4]) { ldsm_x4(a, abase + ks * 16); }, nK, sK, hf,
                     p.dq + (int64_t)b * p.Sq * p.lddq + h * DH, p.lddq, qt * 16 + g, p.Sq, lane);
      }
    }
    if (has_next) rows_store(pre, buf ^ 1);
    __syncthreads();
  }
}

constexpr int SMEM_FWD_MAX = (128 + 2 * SB) * LDS * 2 + SB;   // 8-warp instantiation
constexpr int smem_fwd(int nw) { return (16 * nw + 2 * SB) * LDS * 2 + SB; }
constexpr int SMEM_DQ = (2 * TQ + 2 * SB) * LDS * 2 + SB;
constexpr int SMEM_DKV = (2 * TK + 2 * SB) * LDS * 2 + 2 * SB * 4;

static int set_attn_smem_attrs() {
  static bool done = false;
  if (done) return KMB_OK;
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd(4));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd(1));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd(2));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd(3));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd(5));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd(6));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd(7));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fwd(8));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV);
  if (e != cudaSuccess) {
    kmb_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
    return KMB_ERR_CUDA;
  }
  done = true;
  return KMB_OK;
}

// KMBART_ATTN_BWD_SPLIT=1 keeps the three-kernel backward for every shape (A/B timing, tests)
static bool attn_bwd_fused_enabled() {
  const char* e = getenv("KMBART_ATTN_BWD_SPLIT");
  return !(e && e[0] == '1');
}

// The persistent forward is opt-in (KMBART_ATTN_FWD_PERSIST=1).  Timed alone it beats the tiled kernel (enc 44 vs 58 us,
// rotating buffers), but inside the train step the tiled kernel's 3072 small CTAs overlap with the neighbouring
// kernels under programmatic dependent launch while a 193 KB-smem persistent CTA cannot co-reside with a draining GEMM:
// A/B on the same box, 30 steps: 12.58 / 12.66 ms persistent vs 12.55 / 12.61 ms tiled.  Both paths are parity-tested.
static bool attn_fwd_persist_enabled() {
  const char* e = getenv("KMBART_ATTN_FWD_PERSIST");
  return e && e[0] == '1';
}
static int attn_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

static int check_attn_args(const AttnParams& p) {
  if (!p.q || !p.k || !p.v || p.B <= 0 || p.H <= 0 || p.Sq <= 0 || p.Sk <= 0) return KMB_ERR_ARG;
  if (!(p.scale > 0.f)) return KMB_ERR_ARG;   // the kernels track the maximum of the raw scores and fold the scale into the exponent
  if ((p.ldq % 8) || (p.ldk % 8) || (p.ldv % 8)) return KMB_ERR_ARG;
  return KMB_OK;
}

// tiled forward launch: one CTA of ceil(Sq/16) warps per (batch, head) when Sq <= 128, else 64-row tiles
static void launch_attn_fwd_tiled(const AttnParams& p, cudaStream_t st) {
  const int nw = p.Sq <= 128 ? (p.Sq + 15) / 16 : 4;
  const dim3 grid((p.Sq + 16 * nw - 1) / (16 * nw), p.H, p.B);
  switch (nw) {
    case 1: launch_pdl(attn_fwd_kernel<1>, grid, dim3(32), (size_t)smem_fwd(1), st, p); break;
    case 2: launch_pdl(attn_fwd_kernel<2>, grid, dim3(64), (size_t)smem_fwd(2), st, p); break;
    case 3: launch_pdl(attn_fwd_kernel<3>, grid, dim3(96), (size_t)smem_fwd(3), st, p); break;
    case 5: launch_pdl(attn_fwd_kernel<5>, grid, dim3(160), (size_t)smem_fwd(5), st, p); break;
    case 6: launch_pdl(attn_fwd_kernel<6>, grid, dim3(192), (size_t)smem_fwd(6), st, p); break;
    case 7: launch_pdl(attn_fwd_kernel<7>, grid, dim3(224), (size_t)smem_fwd(7), st, p); break;
    case 8: launch_pdl(attn_fwd_kernel<8>, grid, dim3(256), (size_t)smem_fwd(8), st, p); break;
    default: launch_pdl(attn_fwd_kernel<4>, grid, dim3(128), (size_t)smem_fwd(4), st, p); break;
  }
}

}  // namespace kmb

static void token_major_strides(kmb::AttnParams& p) {
  p.sbq = (int64_t)p.Sq * p.ldq; p.shq = kmb::DH;
  p.sbk = (int64_t)p.Sk * p.ldk; p.shk = kmb::DH;
  p.sbv = (int64_t)p.Sk * p.ldv; p.shv = kmb::DH;
  p.sbo = (int64_t)p.Sq * p.ldo; p.sho = kmb::DH;
}

namespace kmb {   // attention_tc05.cu
bool attn_tc05_enabled(int is_bwd, int Sq, int Sk);
int attn_fwd_tc05(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* o, int64_t ldo, float* lse,
                  const uint8_t* key_pad, int B, int H, int Sq, int Sk, int causal, float scale, cudaStream_t st);
int attn_bwd_tc05(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const void* o, int64_t ldo, const void* d_o,
                  int64_t lddo, const float* lse, const uint8_t* key_pad, void* dq, void* dk, void* dv, int64_t lddq, int64_t lddk, int64_t lddv,
                  int B, int H, int Sq, int Sk, int causal, float scale, cudaStream_t st);
static bool tc05_aligned(const void* a, const void* b, const void* c, int64_t la, int64_t lb, int64_t lc) {
  return !(((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) && !((la | lb | lc) % 8);
}
}  // namespace kmb

extern "C" int kmb_attn_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                            void* o, int64_t ldo, float* lse, const uint8_t* key_pad, int B, int H, int Sq, int Sk,
                            int head_dim, int causal, float scale, kmb_stream_t stream) {
  using namespace kmb;
  AttnParams p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.o = (bf16*)o; p.ldo = ldo; p.lse = lse; p.key_pad = key_pad;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = causal; p.scale = scale;
  token_major_strides(p);
  if (head_dim != DH || !o || (ldo % 8) || check_attn_args(p)) {
    kmb_set_last_error("kmb_attn_fwd: bad argument (head_dim must be 64, strides multiples of 8)", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  // every attention of the base model (S <= 128) is one tcgen05 tile per (batch, head): attention_tc05.cu
  if (Sq <= 128 && Sk <= 128 && attn_tc05_enabled(0, Sq, Sk) && !attn_fwd_persist_enabled() && tc05_aligned(q, k, v, ldq, ldk, ldv) &&
      !((uintptr_t)o & 15))
    return attn_fwd_tc05(q, k, v, ldq, ldk, ldv, o, ldo, lse, key_pad, B, H, Sq, Sk, causal, scale, (cudaStream_t)stream);
  if (set_attn_smem_attrs()) return KMB_ERR_CUDA;
  if (Sq <= 128 && Sk <= 128 && scale > 0.f && attn_fwd_persist_enabled()) {
    const int SqP = (Sq + 15) & ~15, SkP = (Sk + 15) & ~15;
    const int nQ = SqP >> 4;
    int G = AF_WARPS / nQ;
    G = G > AF_MAXG ? AF_MAXG : G;
    const int fit = (FB_SMEM_MAX / 2) / af_unit_bytes(SqP, SkP);
    G = G > fit ? fit : G;
    if (G >= 1) {
      static bool attr = false;
      if (!attr) {
        if (cudaFuncSetAttribute(attn_fwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_MAX) != cudaSuccess) {
          kmb_set_last_error("attn_fwd_persist_kernel: cannot opt in to 227 KB shared memory", __FILE__, __LINE__);
          return KMB_ERR_CUDA;
        }
        attr = true;
      }
      const int n_groups = (B * H + G - 1) / G;
      const int sms = attn_num_sms();
      launch_pdl(attn_fwd_persist_kernel, dim3(n_groups < sms ? n_groups : sms), dim3(AF_THREADS), (size_t)2 * G * af_unit_bytes(SqP, SkP),
                 (cudaStream_t)stream, p, SqP, SkP, G);
      KMB_CHECK_LAUNCH();
      return KMB_OK;
    }
  }
  launch_attn_fwd_tiled(p, (cudaStream_t)stream);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

// general-stride forward (no lse): q/o token-major or cache layout, see AttnParams
extern "C" int kmb_attn_fwd_strided(const void* q, const void* k, const void* v, void* o, const int64_t* strides12,
                                    const uint8_t* key_pad, int B, int H, int Sq, int Sk, int head_dim, int causal,
                                    float scale, kmb_stream_t stream) {
  using namespace kmb;
  AttnParams p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (bf16*)o;
  if (!strides12) { kmb_set_last_error("kmb_attn_fwd_strided: strides missing", __FILE__, __LINE__); return KMB_ERR_ARG; }
  p.sbq = strides12[0]; p.shq = strides12[1]; p.ldq = strides12[2];
  p.sbk = strides12[3]; p.shk = strides12[4]; p.ldk = strides12[5];
  p.sbv = strides12[6]; p.shv = strides12[7]; p.ldv = strides12[8];
  p.sbo = strides12[9]; p.sho = strides12[10]; p.ldo = strides12[11];
  p.key_pad = key_pad; p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = causal; p.scale = scale;
  bool bad = head_dim != DH || !o || check_attn_args(p);
  for (int i = 0; i < 12; ++i) bad = bad || (strides12[i] % 8);
  if (bad) { kmb_set_last_error("kmb_attn_fwd_strided: bad argument", __FILE__, __LINE__); return KMB_ERR_ARG; }
  if (set_attn_smem_attrs()) return KMB_ERR_CUDA;
  launch_attn_fwd_tiled(p, (cudaStream_t)stream);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_attn_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                            const void* o, int64_t ldo, const void* d_o, int64_t lddo, const float* lse,
                            float* d_scratch, const uint8_t* key_pad, void* dq, void* dk, void* dv, int64_t lddq,
                            int64_t lddk, int64_t lddv, int B, int H, int Sq, int Sk, int head_dim, int causal,
                            float scale, kmb_stream_t stream) {
  using namespace kmb;
  AttnParams p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.o = (bf16*)o; p.ldo = ldo; p.lse = (float*)lse; p.key_pad = key_pad;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.causal = causal; p.scale = scale;
  p.dO = (const bf16*)d_o; p.lddo = lddo; p.D = d_scratch;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  token_major_strides(p);
  if (head_dim != DH || !o || !d_o || !lse || !d_scratch || !dq || !dk || !dv || (ldo % 8) || (lddo % 8) ||
      (lddq % 2) || (lddk % 2) || (lddv % 2) || check_attn_args(p)) {
    kmb_set_last_error("kmb_attn_bwd: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (Sq <= 128 && Sk <= 128 && attn_tc05_enabled(1, Sq, Sk) && attn_bwd_fused_enabled() && tc05_aligned(q, k, v, ldq, ldk, ldv) &&
      tc05_aligned(d_o, o, dq, lddo, ldo, lddq) && tc05_aligned(dk, dv, dq, lddk, lddv, lddq))
    return attn_bwd_tc05(q, k, v, ldq, ldk, ldv, o, ldo, d_o, lddo, lse, key_pad, dq, dk, dv, lddq, lddk, lddv, B, H, Sq, Sk, causal, scale, st);
  if (Sq <= FB_MAXS && Sk <= FB_MAXS && attn_bwd_fused_enabled()) {
    const int SqP = (Sq + 15) & ~15, SkP = (Sk + 15) & ~15;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_fused_kernel<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_MAX);
      if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(attn_bwd_fused_kernel<112, 112, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_MAX);
      if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(attn_bwd_fused_kernel<48, 48, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_MAX);
      if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(attn_bwd_fused_kernel<48, 112, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_MAX);
      if (e1 != cudaSuccess) {
        kmb_set_last_error("kmb_attn_bwd: cannot reserve shared memory", __FILE__, __LINE__);
        return KMB_ERR_CUDA;
      }
      attr_set = true;
    }
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (sms <= 0) sms = 148;
    }
    int G = FB_MAXG;   // small shapes: several (batch, head) pairs per iteration keep all 16 warps busy
    while (G > 1 && (fb_smem_bytes(SqP, SkP, G) > FB_SMEM_MAX || G * SqP > 2 * FB_WARPS * 8)) --G;
    const int groups = (B * H + G - 1) / G;
    const int grid = groups < sms ? groups : sms;
    const size_t smem = (size_t)fb_smem_bytes(SqP, SkP, G);
    // the shapes of the configured workloads (S_e = 100 -> 112, S_d = 48) get constant-folded index arithmetic
    if (SqP == 112 && SkP == 112 && G == 1) launch_pdl(attn_bwd_fused_kernel<112, 112, 1>, dim3(grid), dim3(FB_THREADS), smem, st, p, SqP, SkP, G);
    else if (SqP == 48 && SkP == 48 && G == 3) launch_pdl(attn_bwd_fused_kernel<48, 48, 3>, dim3(grid), dim3(FB_THREADS), smem, st, p, SqP, SkP, G);
    else if (SqP == 48 && SkP == 112 && G == 2) launch_pdl(attn_bwd_fused_kernel<48, 112, 2>, dim3(grid), dim3(FB_THREADS), smem, st, p, SqP, SkP, G);
    else launch_pdl(attn_bwd_fused_kernel<0, 0, 0>, dim3(grid), dim3(FB_THREADS), smem, st, p, SqP, SkP, G);
    KMB_CHECK_LAUNCH();
    return KMB_OK;
  }
  const int rows = B * H * Sq;
  launch_pdl(attn_bwd_prep_kernel, dim3((rows * 32 + 255) / 256), dim3(256), 0, st, p);
  KMB_CHECK_LAUNCH();
  if (set_attn_smem_attrs()) return KMB_ERR_CUDA;
  launch_pdl(attn_bwd_dq_kernel, dim3(dim3((Sq + TQ - 1) / TQ, H, B)), dim3(128), SMEM_DQ, st, p);
  KMB_CHECK_LAUNCH();
  launch_pdl(attn_bwd_dkv_kernel, dim3(dim3((Sk + TK - 1) / TK, H, B)), dim3(128), SMEM_DKV, st, p);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

// ------------------------------------------------------------------ decode-step attention (one query token per row)
// replaces: HF-3.0.2 SelfAttention.forward on the cached path (use_cache=True): the reference grows
//   prev_key / prev_value with torch.cat every step and re-orders every cached tensor (cross-attention K/V
//   included) with index_select after each beam step (src/model/mixins.py:419-434).
// Here the caches are preallocated and never moved: row j reads the key/value of position p from slot
//   slot_tbl[j, p] (beam ancestry table, self-attention) or slot j / row_div (cross-attention: K/V exist
//   once per SAMPLE, not once per beam).  One warp per (row, head); scores live in registers.
namespace kmb {
struct DecAttnParams {
  const void* q; int64_t q_rs;                        // q of (row, head) at q + row*q_rs + head*64 (elements)
  const void* k; const void* v; int64_t kv_ss, kv_ps;  // K/V of (slot, pos, head) at k + slot*kv_ss + pos*kv_ps + head*64
  const int* slot_tbl; int64_t tbl_ld;                // [rows, tbl_ld] or null
  int row_div;                                         // slot = row / row_div when slot_tbl is null
  const uint8_t* key_pad; int64_t pad_ld;             // [n_slots, pad_ld] (1 = padding) indexed by row / row_div, or null
  void* o; int64_t o_rs;
  int rows, H, T;
  int causal_mod;                                      // > 0: row is query (row % causal_mod) of its slot and sees keys <= it
  float scale;
};

template <typename T> struct Ld16;   // 16 consecutive elements -> 16 floats
template <> struct Ld16<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&f)[16]) {
    const uint4* q4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint4 u = __ldg(q4 + i);
      const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[8 * i + 2 * j] = __uint_as_float(ww[j] << 16);
        f[8 * i + 2 * j + 1] = __uint_as_float(ww[j] & 0xFFFF0000u);
      }
    }
  }
  static __device__ __forceinline__ float2 load2(const bf16* p) {
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
  }
  static __device__ __forceinline__ void store2(bf16* p, float a, float b) { *reinterpret_cast<uint32_t*>(p) = pack2(a, b); }
};
template <> struct Ld16<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[16]) {
    const float4* q4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 u = __ldg(q4 + i);
      f[4 * i] = u.x; f[4 * i + 1] = u.y; f[4 * i + 2] = u.z; f[4 * i + 3] = u.w;
    }
  }
  static __device__ __forceinline__ float2 load2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};

// Four lanes share one key: each lane holds 16 of the 64 query dims and reads 16 contiguous elements of the key, so
// a quad reads the key's whole row (full sectors) and a warp scores 8 keys per pass with 2 shuffles.
// ET = bf16: the decode chain.  ET = float: the fp32 parity mode, where the same kernel also serves full-sequence
// attention (one "row" per query token, causal_mod = S_q).
template <typename ET>
__global__ void __launch_bounds__(128) decode_attn_kernel(const DecAttnParams p) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= p.rows * p.H) return;
  const int row = w / p.H, h = w % p.H;
  const int sub = lane & 3, kq = lane >> 2;   // dims [16*sub, 16*sub+16) of key (8*pass + kq)
  const ET* qb = reinterpret_cast<const ET*>(p.q);
  const ET* kb = reinterpret_cast<const ET*>(p.k);
  const ET* vb = reinterpret_cast<const ET*>(p.v);
  float qf[16];
  Ld16<ET>::load(qb + (int64_t)row * p.q_rs + h * 64 + sub * 16, qf);
  const int bslot = row / p.row_div;
  const int T = p.causal_mod > 0 ? min(p.T, row % p.causal_mod + 1) : p.T;
  __shared__ float s_sc[4][1024];   // scores / probabilities of this warp's keys (T <= 1024)
  __shared__ int s_slot[4][1024];
  float* sc = s_sc[threadIdx.x >> 5];
  int* sl = s_slot[threadIdx.x >> 5];
  float mx = -INFINITY;
  for (int pos0 = 0; pos0 < T; pos0 += 8) {
    const int pos = pos0 + kq;
    float acc = 0.f;
    bool ok = pos < T;
    int slot = bslot;
    if (ok) {
      if (p.slot_tbl) slot = p.slot_tbl[(int64_t)row * p.tbl_ld + pos];
      if (p.key_pad && p.key_pad[(int64_t)bslot * p.pad_ld + pos]) ok = false;
    }
    if (ok) {
      float kf[16];
      Ld16<ET>::load(kb + (int64_t)slot * p.kv_ss + (int64_t)pos * p.kv_ps + h * 64 + sub * 16, kf);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc = fmaf(qf[j], kf[j], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    const float v = ok ? acc * p.scale : -INFINITY;
    if (sub == 0 && pos < T) { sc[pos] = v; sl[pos] = slot; }
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  __syncwarp();
  float sum = 0.f;
  for (int pos = lane; pos < T; pos += 32) {
    const float v = sc[pos];
    const float e = (v == -INFINITY) ? 0.f : __expf(v - mx);   // a fully masked row gives 0/0 = NaN like the reference
    sc[pos] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  // output: lane owns dims (2*lane, 2*lane+1); one coalesced row read of V per key
  float o0 = 0.f, o1 = 0.f;
#pragma unroll 4
  for (int pos = 0; pos < T; ++pos) {
    const float pr = sc[pos];
    if (pr != 0.f) {   // warp-uniform
      const float2 u = Ld16<ET>::load2(vb + (int64_t)sl[pos] * p.kv_ss + (int64_t)pos * p.kv_ps + h * 64 + 2 * lane);
      o0 = fmaf(pr, u.x, o0);
      o1 = fmaf(pr, u.y, o1);
    }
  }
  Ld16<ET>::store2(reinterpret_cast<ET*>(p.o) + (int64_t)row * p.o_rs + h * 64 + 2 * lane, o0 * inv, o1 * inv);
}
}  // namespace kmb

static int decode_attn_common(int elt, const void* q, int64_t q_row_stride, const void* k, const void* v, int64_t kv_slot_stride,
                              int64_t kv_pos_stride, const int* slot_tbl, int64_t tbl_ld, int row_div, const uint8_t* key_pad,
                              int64_t pad_ld, void* o, int64_t o_row_stride, int rows, int H, int T, int head_dim, int causal_mod,
                              float scale, kmb_stream_t stream) {
  using namespace kmb;
  const int al = elt == 0 ? 8 : 4;   // 16-byte vector loads
  if (!q || !k || !v || !o || rows <= 0 || H <= 0 || T <= 0 || T > 1024 || head_dim != DH || row_div <= 0 || (q_row_stride % al) ||
      (kv_slot_stride % al) || (kv_pos_stride % al) || (o_row_stride % 2) || causal_mod < 0) {
    kmb_set_last_error("kmb_decode_attn: bad argument (head_dim 64, T <= 1024, 16-byte aligned strides)", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  DecAttnParams p;
  p.q = q; p.q_rs = q_row_stride; p.k = k; p.v = v; p.kv_ss = kv_slot_stride;
  p.kv_ps = kv_pos_stride; p.slot_tbl = slot_tbl; p.tbl_ld = tbl_ld; p.row_div = row_div; p.key_pad = key_pad; p.pad_ld = pad_ld;
  p.o = o; p.o_rs = o_row_stride; p.rows = rows; p.H = H; p.T = T; p.causal_mod = causal_mod; p.scale = scale;
  const int64_t warps = (int64_t)rows * H;
  const unsigned blocks = (unsigned)((warps * 32 + 127) / 128);
  if (elt == 0) decode_attn_kernel<bf16><<<blocks, 128, 0, (cudaStream_t)stream>>>(p);
  else decode_attn_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>(p);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_decode_attn(const void* q, int64_t q_row_stride, const void* k, const void* v, int64_t kv_slot_stride,
                               int64_t kv_pos_stride, const int* slot_tbl, int64_t tbl_ld, int row_div, const uint8_t* key_pad,
                               int64_t pad_ld, void* o, int64_t o_row_stride, int rows, int H, int T, int head_dim, float scale,
                               kmb_stream_t stream) {
  return decode_attn_common(0, q, q_row_stride, k, v, kv_slot_stride, kv_pos_stride, slot_tbl, tbl_ld, row_div, key_pad, pad_ld, o,
                            o_row_stride, rows, H, T, head_dim, 0, scale, stream);
}

// fp32 parity mode: same kernel on fp32 q/k/v/o; causal_mod = S_q turns it into full-sequence causal attention
extern "C" int kmb_attn_f32(const float* q, int64_t q_row_stride, const float* k, const float* v, int64_t kv_slot_stride,
                            int64_t kv_pos_stride, int row_div, const uint8_t* key_pad, int64_t pad_ld, float* o,
                            int64_t o_row_stride, int rows, int H, int T, int head_dim, int causal_mod, float scale,
                            kmb_stream_t stream) {
  return decode_attn_common(1, q, q_row_stride, k, v, kv_slot_stride, kv_pos_stride, nullptr, 0, row_div, key_pad, pad_ld, o,
                            o_row_stride, rows, H, T, head_dim, causal_mod, scale, stream);
}
