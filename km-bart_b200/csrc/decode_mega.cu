// Persistent decode-step kernel: ONE cooperative launch runs the whole cached decoder step
//   embed+LN -> L x { QKV -> self-attention -> out-proj -> LN -> cross-q -> cross-attention -> out-proj -> LN
//                     -> fc1+GELU -> fc2 -> LN } [-> LM head (+ greedy arg-max)]
// for all rows (batch x beams) of one generation step.
// replaces: one `self(**model_inputs)` call of the HF-3.0.2 generation loops reached from
//   src/model/mixins.py:336-382 (decoder forward with use_cache=True: HF-3.0.2 BartDecoder.forward,
//   DecoderLayer.forward, SelfAttention.forward cached branches; instantiated at src/model/model.py:35) and the
//   LM head of src/model/model.py:397.
//
// Why one kernel: at rows <= 320 and d = 768 every sub-step is a 1-3 us latency chain; 68 dependent launches
// per token cost 750 us against 47 us of HBM time (profiles/r01c_decode_analysis.md).  Here the dependent steps
// are separated by grid-wide barriers (~1 us) instead of launches, and — the part a launch chain cannot do — the
// weight stream is decoupled from the dependency chain: each CTA knows the sequence of weight slices it will
// need (it owns fixed output columns of every Linear), so an elected warp keeps a ring of shared-memory slots
// filled with cp.async.bulk copies running up to NSLOT GEMM units (about half a layer) ahead of the barriers.
// HBM therefore streams continuously while the activations (<= 0.5 MB, L2 resident) take the latency path.
//
// Work decomposition.  A GEMM "unit" is  out[rows, 8*NT columns] (+)= A[rows, KC] . W[8*NT rows, KC]^T  with
// KC = d: the Linear's output columns are cut into strips of 8/16/24 and fc2's K = ffn into ffn/d chunks (partial
// sums combined with red.global.add.f32).  Unit u of a Linear belongs to CTA u % grid.  Inside a unit the 16 warps
// tile 64 rows x 4 K-groups; A fragments are read straight from L2 into registers (ld.global.cg, 16 B per lane,
// with a K permutation shared by A and W so that one 16-byte load feeds two mma.sync.m16n8k16), W fragments come
// from the prefetched slot (rows padded by 64 B: conflict-free LDS.128).  Tensor throughput is irrelevant here
// (0.2 GFLOP per Linear); mma.sync is used because it needs no TMEM/descriptor setup inside a latency chain.
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

constexpr int MG_THREADS = 512;
constexpr int MG_WARPS = MG_THREADS / 32;
constexpr int MG_MAXT = 512;        // keys per attention call (self: max_len, cross: S_e)
constexpr int MG_MAX_NT = 3;        // unit width = 8 * NT columns
constexpr int MG_ROW_PAD = 64;      // bytes of padding after each weight row in a slot

template <int D> struct MegaCfg {
  static constexpr int KC = D;                                   // K chunk of every unit
  static constexpr int ROW_BYTES = KC * 2 + MG_ROW_PAD;
  static constexpr int SLOT_BYTES = 8 * MG_MAX_NT * ROW_BYTES;
  static constexpr int SCRATCH_BYTES = MG_WARPS * MG_MAXT * 8;   // attention scores+slots / GEMM K-group partials
  static constexpr int NSLOT = (D <= 768) ? 4 : 3;
  static constexpr int SMEM_BYTES = NSLOT * SLOT_BYTES + SCRATCH_BYTES + 256;
  static_assert(4 * 64 * 8 * MG_MAX_NT * 4 <= SCRATCH_BYTES, "partials must fit the scratch area");
};

// ------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ uint4 ldcg16(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldcg_f4(const void* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ldcg_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// ------------------------------------------------------------------ grid barrier
// Monotonic 64-bit arrival counter shared by every launch of a session (never reset): a launch starts from the
// largest multiple of gridDim.x not above the value it first observes (a late CTA can see at most gridDim.x - 1
// early arrivals of barrier 0).  Bounded spin: a scheduling accident traps instead of hanging the box.
__device__ __forceinline__ unsigned long long mg_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct GridBar {
  unsigned long long* ctr;
  unsigned long long target;
  unsigned long long* trace;   // optional {arrive, release} timestamps per barrier and CTA
  int n;
  __device__ void init(unsigned long long* c, unsigned long long* tr) {
    ctr = c;
    trace = tr;
    n = 0;
    const unsigned long long v = ld_acquire_u64(c);
    target = v - v % gridDim.x;
  }
  __device__ void sync() {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      if (trace) trace[((size_t)n * gridDim.x + blockIdx.x) * 2] = mg_gtimer();
      __threadfence();
      atomicAdd(ctr, 1ULL);
      const long long t0 = clock64();
      while (ld_acquire_u64(ctr) < target) {
        if (clock64() - t0 > 4000000000LL) {
          printf("kmbart decode_mega: grid barrier timeout (block %d)\n", blockIdx.x);
          __trap();
        }
      }
      if (trace) trace[((size_t)n * gridDim.x + blockIdx.x) * 2 + 1] = mg_gtimer();
    }
    ++n;
    __syncthreads();
  }
};

// ------------------------------------------------------------------ GEMM units
enum { MG_OUT_BF16 = 0, MG_OUT_BF16_GELU = 1, MG_OUT_F32 = 2, MG_OUT_F32_RED = 3, MG_OUT_LOGITS = 4 };

struct MgGemm {
  const bf16* W;      // [N, ldw], k contiguous
  const float* bias;  // [N]
  const bf16* A;      // [rows, lda]
  void* out;          // bf16 or fp32 [rows, ldo]
  int64_t lda, ldo, ldw;
  int N, NT, kchunks, mode;
  __device__ int units() const { return (N + 8 * NT - 1) / (8 * NT) * kchunks; }
};

template <int D>
struct WeightRing {
  using C = MegaCfg<D>;
  uint8_t* slots;
  uint64_t* full;
  uint32_t cons;   // units consumed by this CTA so far (all threads track it)
  // producer cursor (warp 0 only)
  uint32_t prod;
  int pg, pu;      // next unit to fetch: Linear index pg in the step's sequence, unit pu
};

template <int D>
__device__ __forceinline__ void ring_issue(WeightRing<D>& r, const MgGemm& g, int u) {
  using C = MegaCfg<D>;
  const int lane = threadIdx.x & 31;
  const int slot = r.prod % C::NSLOT;
  const int n_strips = (g.N + 8 * g.NT - 1) / (8 * g.NT);
  const int strip = u % n_strips, kc = u / n_strips;
  const int n0 = strip * 8 * g.NT;
  const int nrows = min(8 * g.NT, g.N - n0);
  if (lane == 0) mbar_arrive_expect_tx(&r.full[slot], (uint32_t)nrows * C::KC * 2);
  __syncwarp();
  if (lane < nrows)
    bulk_g2s(r.slots + (size_t)slot * C::SLOT_BYTES + (size_t)lane * C::ROW_BYTES,
             g.W + (int64_t)(n0 + lane) * g.ldw + (int64_t)kc * C::KC, C::KC * 2, &r.full[slot]);
  r.prod++;
}

// One unit: out[rows, n0 .. n0+8NT) (+)= A[:, kc*KC .. +KC) . Wslot^T, 64 rows at a time.
template <int D, int NT>
__device__ void gemm_unit(const MgGemm& g, int u, int rows, const uint8_t* wslot, float* scratch) {
  using C = MegaCfg<D>;
  constexpr int KG = C::KC / 4;        // K range of one warp
  constexpr int KB = KG / 32;          // 32-wide K blocks per warp (one 16-byte load each)
  static_assert(KG % 32 == 0, "d must be a multiple of 128");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 3, kg = warp >> 2, gq = lane >> 2, tg = lane & 3;
  const int n_strips = (g.N + 8 * NT - 1) / (8 * NT);
  const int strip = u % n_strips, kc = u / n_strips;
  const int n0 = strip * 8 * NT;
  const bf16* abase = g.A + (int64_t)kc * C::KC + kg * KG + tg * 8;
  uint4 alo[KB], ahi[KB];
  auto load_a = [&](int m0) {
    const int ra = m0 + mt * 16 + gq, rb = ra + 8;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      alo[kb] = ra < rows ? ldcg16(abase + (int64_t)ra * g.lda + kb * 32) : make_uint4(0, 0, 0, 0);
      ahi[kb] = rb < rows ? ldcg16(abase + (int64_t)rb * g.lda + kb * 32) : make_uint4(0, 0, 0, 0);
    }
  };
  load_a(0);
  for (int m0 = 0; m0 < rows; m0 += 64) {
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const uint8_t* wb = wslot + (size_t)gq * C::ROW_BYTES + (size_t)(kg * KG + tg * 8) * 2;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const uint4 b = *reinterpret_cast<const uint4*>(wb + (size_t)j * 8 * C::ROW_BYTES + kb * 64);
        mma_bf16(acc[j], alo[kb].x, ahi[kb].x, alo[kb].y, ahi[kb].y, b.x, b.y);
        mma_bf16(acc[j], alo[kb].z, ahi[kb].z, alo[kb].w, ahi[kb].w, b.z, b.w);
      }
    }
    // the next 64 rows of A travel from L2 while the partials are reduced and stored (rows > 64: beam search; with
    // the loads at the top of the iteration every 64-row block paid one exposed L2 round trip)
    if (m0 + 64 < rows) load_a(m0 + 64);
    // K-group partials -> scratch[kg][64][8NT], then every thread finishes a few (row, column-pair) outputs
    float* part = scratch + (size_t)kg * 64 * 8 * NT;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      *reinterpret_cast<float2*>(part + (mt * 16 + gq) * 8 * NT + j * 8 + 2 * tg) = make_float2(acc[j][0], acc[j][1]);
      *reinterpret_cast<float2*>(part + (mt * 16 + gq + 8) * 8 * NT + j * 8 + 2 * tg) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncthreads();
    constexpr int PAIRS = 64 * 4 * NT;
    for (int i = threadIdx.x; i < PAIRS; i += MG_THREADS) {
      const int r = i / (4 * NT), cp = i % (4 * NT);
      const int row = m0 + r, col = n0 + 2 * cp;
      if (row < rows && col < g.N) {
        float2 v = *reinterpret_cast<const float2*>(scratch + r * 8 * NT + 2 * cp);
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          const float2 w = *reinterpret_cast<const float2*>(scratch + (size_t)q * 64 * 8 * NT + r * 8 * NT + 2 * cp);
          v.x += w.x; v.y += w.y;
        }
        if (kc == 0 && g.bias) { v.x += __ldg(g.bias + col); v.y += __ldg(g.bias + col + 1); }
        if (g.mode == MG_OUT_BF16 || g.mode == MG_OUT_BF16_GELU) {
          if (g.mode == MG_OUT_BF16_GELU) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); }
          *reinterpret_cast<uint32_t*>(reinterpret_cast<bf16*>(g.out) + (int64_t)row * g.ldo + col) = pack_bf16x2(v.x, v.y);
        } else if (g.mode == MG_OUT_F32_RED) {
          float* o = reinterpret_cast<float*>(g.out) + (int64_t)row * g.ldo + col;
          red_add_f32(o, v.x);
          red_add_f32(o + 1, v.y);
        } else {
          *reinterpret_cast<float2*>(reinterpret_cast<float*>(g.out) + (int64_t)row * g.ldo + col) = v;
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ LayerNorm rows (warp per row)
// y = LN(x (+ add)) -> out_f32, out_bf16; `add` (the Linear output, bias included) is optionally cleared
// afterwards so that the split-K fc2 can accumulate into it with reductions.
template <int D>
__device__ __forceinline__ void ln_row(const float4 (&xin)[D / 128], const float* gamma, const float* beta, float* out_f32,
                                       bf16* out_b16, int lane) {
  constexpr int NV = D / 128;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) sum += xin[i].x + xin[i].y + xin[i].z + xin[i].w;
  const float mean = warp_sum(sum) * (1.f / D);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = xin[i].x - mean, b = xin[i].y - mean, c = xin[i].z - mean, e = xin[i].w - mean;
    var += a * a + b * b + c * c + e * e;
  }
  const float rstd = rsqrtf(warp_sum(var) * (1.f / D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y;
    y.x = (xin[i].x - mean) * rstd * g.x + b.x;
    y.y = (xin[i].y - mean) * rstd * g.y + b.y;
    y.z = (xin[i].z - mean) * rstd * g.z + b.z;
    y.w = (xin[i].w - mean) * rstd * g.w + b.w;
    *reinterpret_cast<float4*>(out_f32 + c) = y;
    uint2 pk;
    pk.x = pack_bf16x2(y.x, y.y);
    pk.y = pack_bf16x2(y.z, y.w);
    *reinterpret_cast<uint2*>(out_b16 + c) = pk;
  }
}

template <int D>
__device__ void ln_phase(int rows, float* xf, bf16* xb, float* add, bool clear_add, const float* gamma, const float* beta) {
  constexpr int NV = D / 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x + gridDim.x * warp; row < rows; row += gridDim.x * MG_WARPS) {
    float4 x[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      x[i] = ldcg_f4(xf + (int64_t)row * D + c);
      const float4 a = ldcg_f4(add + (int64_t)row * D + c);
      x[i].x += a.x; x[i].y += a.y; x[i].z += a.z; x[i].w += a.w;
      if (clear_add) *reinterpret_cast<float4*>(add + (int64_t)row * D + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ln_row<D>(x, gamma, beta, xf + (int64_t)row * D, xb + (int64_t)row * D, lane);
  }
}

// ------------------------------------------------------------------ attention rows (warp per (row, head))
// Same arithmetic as decode_attn_kernel (attention.cu): four lanes share a key, fp32 softmax, lane-owns-2-dims PV.
struct MgAttn {
  const bf16 *q, *k, *v;
  int64_t q_rs, kv_ss, kv_ps;
  const int* slot_tbl;
  int64_t tbl_ld;
  int row_div;
  const uint8_t* key_pad;
  int64_t pad_ld;
  bf16* o;
  int64_t o_rs;
  int T;
  float scale;
  bool kv_mutable;   // keys written earlier in this launch: read through L2
};

__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float (&f)[16]) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[2 * i] = bf_lo(w[i]); f[2 * i + 1] = bf_hi(w[i]); }
}

// Memory-level parallelism is the whole game here (one warp, one (row, head), a ~0.7 us L2/HBM round trip per
// dependent load): both loops request 32 keys per iteration with every load of the batch issued before the first use.
//   scores: lane = (key phase kq = lane >> 2, 16-dim slice sub = lane & 3): 4 keys x 8 phases in flight
//   P V   : lane = (key phase kg = lane >> 3, 8-dim slice dc = lane & 7): 16-byte loads, 8 keys x 4 phases in flight,
//           the four key phases are combined with two shuffles at the end.
__device__ __forceinline__ uint4 mg_ld16(const bf16* p, bool mut) {
  return mut ? ldcg16(p) : __ldg(reinterpret_cast<const uint4*>(p));
}

__device__ void attn_phase(const MgAttn& p, int rows, int H, float* scratch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc = scratch + (size_t)warp * MG_MAXT * 2;
  int* sl = reinterpret_cast<int*>(sc + MG_MAXT);
  const int sub = lane & 3, kq = lane >> 2;
  const int kg = lane >> 3, dc = lane & 7;
  const int units = rows * H;
  const bool mut = p.kv_mutable;
  for (int u = blockIdx.x + gridDim.x * warp; u < units; u += gridDim.x * MG_WARPS) {
    const int row = u / H, h = u % H;
    float qf[16];
    {
      const bf16* qp = p.q + (int64_t)row * p.q_rs + h * 64 + sub * 16;
      unpack16(ldcg16(qp), ldcg16(qp + 8), qf);
    }
    const int bslot = row / p.row_div;
    const int T = p.T;
    float mx = -INFINITY;
    for (int pos0 = 0; pos0 < T; pos0 += 32) {
      bool ok[4];
      int slot[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pos = pos0 + 8 * j + kq;
        ok[j] = pos < T;
        slot[j] = bslot;
        if (ok[j] && p.slot_tbl) slot[j] = __ldg(p.slot_tbl + (int64_t)row * p.tbl_ld + pos);
      }
      if (p.key_pad) {
        uint8_t pd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pd[j] = ok[j] ? __ldg(p.key_pad + (int64_t)bslot * p.pad_ld + pos0 + 8 * j + kq) : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) ok[j] = ok[j] && !pd[j];
      }
      uint4 ka[4], kb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pos = pos0 + 8 * j + kq;
        const bf16* kp = p.k + (int64_t)slot[j] * p.kv_ss + (int64_t)pos * p.kv_ps + h * 64 + sub * 16;
        ka[j] = kb[j] = make_uint4(0, 0, 0, 0);
        if (ok[j]) { ka[j] = mg_ld16(kp, mut); kb[j] = mg_ld16(kp + 8, mut); }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pos = pos0 + 8 * j + kq;
        float kf[16];
        unpack16(ka[j], kb[j], kf);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc = fmaf(qf[i], kf[i], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        const float v = ok[j] ? acc * p.scale : -INFINITY;
        if (sub == 0 && pos < T) { sc[pos] = v; sl[pos] = slot[j]; }
        mx = fmaxf(mx, v);
      }
    }
    mx = warp_max(mx);
    __syncwarp();
    float sum = 0.f;
    for (int pos = lane; pos < T; pos += 32) {
      const float v = sc[pos];
      const float e = (v == -INFINITY) ? 0.f : __expf(v - mx);   // fully masked row: 0/0 = NaN like the reference
      sc[pos] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    for (int pos0 = 0; pos0 < T; pos0 += 32) {
      float pr[8];
      uint4 vv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int pos = pos0 + 4 * j + kg;
        pr[j] = pos < T ? sc[pos] : 0.f;
        vv[j] = make_uint4(0, 0, 0, 0);
        if (pr[j] != 0.f) vv[j] = mg_ld16(p.v + (int64_t)sl[pos] * p.kv_ss + (int64_t)pos * p.kv_ps + h * 64 + dc * 8, mut);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t w[4] = {vv[j].x, vv[j].y, vv[j].z, vv[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[2 * i] = fmaf(pr[j], bf_lo(w[i]), o[2 * i]);
          o[2 * i + 1] = fmaf(pr[j], bf_hi(w[i]), o[2 * i + 1]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
      o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
    }
    if (kg == 0) {
      uint4 pk;
      pk.x = pack_bf16x2(o[0] * inv, o[1] * inv);
      pk.y = pack_bf16x2(o[2] * inv, o[3] * inv);
      pk.z = pack_bf16x2(o[4] * inv, o[5] * inv);
      pk.w = pack_bf16x2(o[6] * inv, o[7] * inv);
      *reinterpret_cast<uint4*>(p.o + (int64_t)row * p.o_rs + h * 64 + dc * 8) = pk;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ the step
__device__ __forceinline__ MgGemm layer_gemm(const KmbDecodeStep& p, int l, int kind) {
  const KmbDecodeLayer& L = p.layers[l];
  const int d = p.d;
  MgGemm g;
  g.lda = d; g.ldw = d; g.kchunks = 1; g.ldo = d; g.N = d; g.NT = p.nt[kind];
  switch (kind) {
    case 0:   // fused q|k|v projection of the new token, written into the cache at position t
      g.W = (const bf16*)L.w_qkv; g.bias = L.b_qkv; g.A = (const bf16*)p.x_b16; g.N = 3 * d;
      g.out = (bf16*)L.cache + (int64_t)p.t * 3 * d; g.ldo = (int64_t)p.max_len * 3 * d; g.mode = MG_OUT_BF16;
      break;
    case 1:
      g.W = (const bf16*)L.w_o; g.bias = L.b_o; g.A = (const bf16*)p.ctx; g.out = p.lin; g.mode = MG_OUT_F32;
      break;
    case 2:
      g.W = (const bf16*)L.w_cq; g.bias = L.b_cq; g.A = (const bf16*)p.x_b16; g.out = p.q2; g.mode = MG_OUT_BF16;
      break;
    case 3:
      g.W = (const bf16*)L.w_co; g.bias = L.b_co; g.A = (const bf16*)p.ctx; g.out = p.lin; g.mode = MG_OUT_F32;
      break;
    case 4:
      g.W = (const bf16*)L.w_fc1; g.bias = L.b_fc1; g.A = (const bf16*)p.x_b16; g.N = p.F; g.out = p.h; g.ldo = p.F;
      g.mode = MG_OUT_BF16_GELU;
      break;
    default:  // fc2: K = ffn cut into ffn/d chunks, partial sums reduced into the (cleared) fp32 buffer
      g.W = (const bf16*)L.w_fc2; g.bias = L.b_fc2; g.A = (const bf16*)p.h; g.lda = p.F; g.ldw = p.F; g.kchunks = p.F / d;
      g.out = p.lin; g.mode = MG_OUT_F32_RED;
      break;
  }
  return g;
}

template <int D>
__device__ __forceinline__ void ring_advance(const KmbDecodeStep& p, WeightRing<D>& r) {
  // warp 0: fetch the next unit of this CTA's sequence (if any) into the slot that was just released
  const int total = p.L * 6;
  while (r.pg < total) {
    const MgGemm g = layer_gemm(p, r.pg / 6, r.pg % 6);
    if (r.pu < g.units()) {
      ring_issue<D>(r, g, r.pu);
      r.pu += gridDim.x;
      return;
    }
    r.pg++;
    r.pu = blockIdx.x;
  }
}

template <int D>
__device__ void gemm_phase(const KmbDecodeStep& p, WeightRing<D>& r, int l, int kind, float* scratch) {
  using C = MegaCfg<D>;
  const MgGemm g = layer_gemm(p, l, kind);
  const int U = g.units();
  for (int u = blockIdx.x; u < U; u += gridDim.x) {
    const int slot = r.cons % C::NSLOT;
    mbar_wait(&r.full[slot], (r.cons / C::NSLOT) & 1);
    const uint8_t* ws = r.slots + (size_t)slot * C::SLOT_BYTES;
    if (g.NT == 1) gemm_unit<D, 1>(g, u, p.rows, ws, scratch);
    else if (g.NT == 2) gemm_unit<D, 2>(g, u, p.rows, ws, scratch);
    else gemm_unit<D, 3>(g, u, p.rows, ws, scratch);
    r.cons++;
    // gemm_unit ends with __syncthreads(): every read of the slot is done, refill it
    if (threadIdx.x < 32) ring_advance<D>(p, r);
  }
}

template <int D>
__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(const __grid_constant__ KmbDecodeStep p) {
  using C = MegaCfg<D>;
  extern __shared__ __align__(128) uint8_t smem[];
  WeightRing<D> ring;
  ring.slots = smem;
  float* scratch = reinterpret_cast<float*>(smem + (size_t)C::NSLOT * C::SLOT_BYTES);
  ring.full = reinterpret_cast<uint64_t*>(smem + (size_t)C::NSLOT * C::SLOT_BYTES + C::SCRATCH_BYTES);
  ring.cons = 0; ring.prod = 0; ring.pg = 0; ring.pu = blockIdx.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NSLOT; ++i) mbar_init(&ring.full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x < 32)
    for (int i = 0; i < C::NSLOT; ++i) ring_advance<D>(p, ring);   // the weight stream starts before anything else
  GridBar bar;
  bar.init(p.barrier, p.trace);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xf = p.x_f32;
  bf16* xb = (bf16*)p.x_b16;
  // phase 0: token embedding * scale + learned position (offset already folded into pos_row) -> LayerNorm
  for (int row = blockIdx.x + gridDim.x * warp; row < p.rows; row += gridDim.x * MG_WARPS) {
    const float* te = p.tok_emb + (int64_t)p.ids[row] * D;
    const float* pe = p.pos_emb + (int64_t)p.pos_row * D;
    float4 x[D / 128];
#pragma unroll
    for (int i = 0; i < D / 128; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 a = __ldg(reinterpret_cast<const float4*>(te + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(pe + c));
      x[i] = make_float4(a.x * p.embed_scale + b.x, a.y * p.embed_scale + b.y, a.z * p.embed_scale + b.z, a.w * p.embed_scale + b.w);
    }
    ln_row<D>(x, p.lne_g, p.lne_b, xf + (int64_t)row * D, xb + (int64_t)row * D, lane);
  }
  bar.sync();

  for (int l = 0; l < p.L; ++l) {
    const KmbDecodeLayer& L = p.layers[l];
    gemm_phase<D>(p, ring, l, 0, scratch);
    bar.sync();
    {
      MgAttn a;
      const bf16* base = (const bf16*)L.cache;
      a.q = base + (int64_t)p.t * 3 * D; a.q_rs = (int64_t)p.max_len * 3 * D;
      a.k = base + D; a.v = base + 2 * D; a.kv_ss = (int64_t)p.max_len * 3 * D; a.kv_ps = 3 * D;
      a.slot_tbl = p.slot_tbl; a.tbl_ld = p.max_len; a.row_div = 1; a.key_pad = nullptr; a.pad_ld = 0;
      a.o = (bf16*)p.ctx; a.o_rs = D; a.T = p.t + 1; a.scale = p.attn_scale; a.kv_mutable = true;
      attn_phase(a, p.rows, p.H, scratch);
    }
    bar.sync();
    gemm_phase<D>(p, ring, l, 1, scratch);
    bar.sync();
    ln_phase<D>(p.rows, xf, xb, p.lin, false, L.ln1_g, L.ln1_b);
    bar.sync();
    gemm_phase<D>(p, ring, l, 2, scratch);
    bar.sync();
    {
      MgAttn a;
      const bf16* kv = (const bf16*)L.cross_kv;
      a.q = (const bf16*)p.q2; a.q_rs = D;
      a.k = kv; a.v = kv + D; a.kv_ss = (int64_t)p.Se * 2 * D; a.kv_ps = 2 * D;
      a.slot_tbl = nullptr; a.tbl_ld = 0; a.row_div = p.row_div; a.key_pad = p.key_pad; a.pad_ld = p.Se;
      a.o = (bf16*)p.ctx; a.o_rs = D; a.T = p.Se; a.scale = p.attn_scale; a.kv_mutable = false;
      attn_phase(a, p.rows, p.H, scratch);
    }
    bar.sync();
    gemm_phase<D>(p, ring, l, 3, scratch);
    bar.sync();
    ln_phase<D>(p.rows, xf, xb, p.lin, true, L.ln2_g, L.ln2_b);
    bar.sync();
    gemm_phase<D>(p, ring, l, 4, scratch);
    bar.sync();
    gemm_phase<D>(p, ring, l, 5, scratch);
    bar.sync();
    ln_phase<D>(p.rows, xf, xb, p.lin, false, L.ln3_g, L.ln3_b);
    if (l + 1 < p.L) bar.sync();
  }
}

}  // namespace kmb

// strip width (8 * nt columns): the narrowest strip whose unit count still fits one wave of the grid (most CTAs
// busy, one unit each); if none fits, the widest one (fewest units)
static int mega_nt_for(int N, int kchunks, int grid) {
  int widest = 1;
  for (int nt = 1; nt <= kmb::MG_MAX_NT; ++nt) {
    if (N % (8 * nt)) continue;
    widest = nt;
    if (N / (8 * nt) * kchunks <= grid) return nt;
  }
  return widest;
}

extern "C" int kmb_decode_step_grid(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return KMB_ERR_CUDA;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <int D>
static int launch_mega(const KmbDecodeStep& p, int grid, cudaStream_t stream) {
  using namespace kmb;
  using C = MegaCfg<D>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(decode_mega_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) {
      kmb_set_last_error("kmb_decode_step: cannot reserve shared memory", __FILE__, __LINE__);
      return KMB_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;   // co-residency of the whole grid is what makes the barriers legal
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, decode_mega_kernel<D>, p) != cudaSuccess) {
    kmb_set_last_error(cudaGetErrorString(cudaGetLastError()), __FILE__, __LINE__);
    return KMB_ERR_CUDA;
  }
  return KMB_OK;
}

extern "C" int kmb_decode_step(const KmbDecodeStep* step, kmb_stream_t stream) {
  using namespace kmb;
  if (!step) { kmb_set_last_error("kmb_decode_step: null argument", __FILE__, __LINE__); return KMB_ERR_ARG; }
  KmbDecodeStep p = *step;
  if ((p.d != 128 && p.d != 768 && p.d != 1024) || p.H * 64 != p.d || p.F % p.d || p.L < 1 || p.L > KMB_DECODE_MAX_LAYERS ||
      p.rows < 1 || p.t < 0 || p.t >= p.max_len || p.max_len > MG_MAXT || p.Se < 1 || p.Se > MG_MAXT || p.row_div < 1 ||
      !p.barrier || !p.ids || !p.x_f32 || !p.x_b16 || !p.ctx || !p.lin || !p.q2 || !p.h) {
    kmb_set_last_error("kmb_decode_step: unsupported shape (d in {128, 768, 1024}, head_dim 64, max_len / S_e <= 512) or null buffer",
                       __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const int grid = kmb_decode_step_grid();
  if (grid <= 0) { kmb_set_last_error("kmb_decode_step: no device", __FILE__, __LINE__); return KMB_ERR_CUDA; }
  p.nt[0] = mega_nt_for(3 * p.d, 1, grid);
  p.nt[1] = p.nt[2] = p.nt[3] = mega_nt_for(p.d, 1, grid);
  p.nt[4] = mega_nt_for(p.F, 1, grid);
  p.nt[5] = mega_nt_for(p.d, p.F / p.d, grid);
  if (p.d == 128) return launch_mega<128>(p, grid, (cudaStream_t)stream);
  if (p.d == 768) return launch_mega<768>(p, grid, (cudaStream_t)stream);
  return launch_mega<1024>(p, grid, (cudaStream_t)stream);
}
