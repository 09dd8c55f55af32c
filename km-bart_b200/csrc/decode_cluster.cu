// Persistent decode-step kernel, second generation: ONE launch runs the whole KV-cached decoder step
//   embed -> L x { [LN -> QKV -> self-attention] | out-proj | [LN -> cross-q -> cross-attention] | out-proj | LN -> fc1+GELU | fc2 } -> LN
// for all rows (batch x beams) of one generation step: 6 grid-barrier-separated phases per layer (the first
// generation needed 11).
// replaces: one `self(**model_inputs)` call of the HF-3.0.2 generation loops reached from src/model/mixins.py:336-382
//   (BartDecoder.forward / DecoderLayer.forward / SelfAttention.forward cached branches, instantiated at
//   src/model/model.py:35).
//
// What bounded the first generation (profiles/r02a_decode_trace.txt): every CTA owned output columns of every Linear
// and therefore pulled the WHOLE [rows, K] activation from L2 after every barrier (148 CTAs reading the same 98 KB in
// the same order: 4.5 us per GEMM phase at rows 64, 13-19 us at rows 320), plus three LayerNorm phases and two
// attention phases per layer whose only content was a barrier and a latency chain.  Here:
//   * 4-CTA clusters.  A cluster owns output-column strips; its CTAs split K four ways, so each CTA stages only a
//     [rows, K/4] slice of the activation (4x fewer bytes per SM, 4.6x fewer readers per L2 line), multiplies it with
//     its K-slice of the prefetched weight strips (mma.sync m16n8k16, operands via ldmatrix) and the four partial tiles
//     are summed through distributed shared memory (one cluster barrier), each CTA finishing a quarter of the rows.
//   * LayerNorm never is a phase: the producing epilogue writes the fp32 pre-LN sum and per-(row, strip) partial
//     (sum, sum of squares) with plain stores (fixed summation order => bit-reproducible); consumers normalise on load.
//   * Attention is not a phase either: a cluster owns (head, row group); it projects q|k|v (or the cross-attention q)
//     for exactly the rows it then attends, so projection -> softmax(QK^T)V needs only the cluster barrier.  Two warps
//     share a (row, head) unit (split keys, flash-style merge); cross K/V lines are L2-prefetched a phase ahead.
//   * the weight stream stays decoupled from the dependency chain: every CTA knows the weight slices it will need and
//     keeps a shared-memory ring of cp.async.bulk copies running ~half a layer ahead of the barriers.
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

constexpr int DS_THREADS = 512;
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int DS_CL = 4;              // CTAs per cluster = K split
constexpr int DS_MAX_CLUSTERS = 32;
constexpr int DS_MAXT = 512;          // keys per attention call (self: max_len, cross: S_e)
constexpr int DS_QKV_LD = 192;        // floats per row of the q|k|v staging tile

template <int D> struct DsCfg {
  static constexpr int KS = D / DS_CL;                       // K elements per CTA per chunk
  static constexpr int KSTEPS = KS / 16;
  static constexpr int ROWB = KS * 2 + 16;                   // smem row pitch of A tiles and weight slots (odd multiple of 16 B)
  static constexpr int NTG = D == 768 ? 3 : (D == 1024 ? 4 : 1);   // generic strip = 8 * NTG columns
  static constexpr int SLOT_ROWS = NTG * 8 > 24 ? NTG * 8 : 24;
  static constexpr int SLOT_BYTES = SLOT_ROWS * ROWB;
  static constexpr int ATILE = 64 * ROWB;
  static constexpr int PARTMAX = (NTG == 4) ? 32768 : 24576;
  static constexpr int PARTB = ATILE > PARTMAX ? ATILE : PARTMAX;
  static constexpr int OFF_A0 = 0;
  static constexpr int OFF_A1 = OFF_A0 + ATILE;              // second A buffer, aliased with the partial tiles
  static constexpr int OFF_QKV = OFF_A1 + PARTB;             // [8][192] fp32
  static constexpr int OFF_MG = OFF_QKV + 8 * DS_QKV_LD * 4; // [8][68] fp32 merge buffers
  static constexpr int OFF_RS = OFF_MG + 8 * 68 * 4;         // [64][2] fp32 row mean / rstd
  static constexpr int OFF_ST = OFF_RS + 64 * 2 * 4;         // [16 rows][8][2] fp32 stats scratch (y-epilogues: one strip per round)
  static constexpr int OFF_GB = OFF_ST + 16 * 16 * 4;        // [2][KS] fp32 LayerNorm weight / bias slice of this CTA's K range
  static constexpr int OFF_BAR = OFF_GB + 2 * KS * 4;
  static constexpr int TAB_CAP = KMB_DECODE_MAX_LAYERS * 32;  // weight slots of one CTA and step (22 per layer with 32 clusters)
  static constexpr int OFF_TAB = OFF_BAR + 256;              // [TAB_CAP] source address | rows / 8
  static constexpr int OFF_RING = OFF_TAB + TAB_CAP * 8;
  static constexpr int NSLOT_RAW = (232448 - 1024 - OFF_RING) / SLOT_BYTES;
  static constexpr int NSLOT = NSLOT_RAW > 16 ? 16 : NSLOT_RAW;
  static constexpr int SMEM_BYTES = OFF_RING + NSLOT * SLOT_BYTES;
  static_assert(NSLOT >= 8, "a phase may need 8 resident weight slots");
  static_assert(KSTEPS % 2 == 0 || KSTEPS == 1, "K split");
};

// ------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ uint4 ds_ldcg16(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ds_ldcg_f4(const void* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ds_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ds_ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ds_fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void ds_bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ds_red_release(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ds_bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void ds_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void ds_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ds_ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ds_ldsm2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ float4 ds_ld_cluster_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr));
  return v;
}
__device__ __forceinline__ uint32_t ds_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ds_round_bf16(float a) { return __bfloat162float(__float2bfloat16(a)); }
__device__ __forceinline__ float ds_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float ds_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ unsigned long long ds_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// optional intra-phase event stamps (profiling aid): clock64 of thread 0, layer 1 only, behind the barrier table
#define DS_EV(id) do { if (ev && threadIdx.x == 0) ev[id] = (unsigned long long)clock64(); } while (0)

// ------------------------------------------------------------------ grid barrier
// Monotonic 64-bit arrival counter shared by every launch of a session (never reset): a launch starts from the
// largest multiple of gridDim.x not above the value it first observes.  One release-add per CTA, acquire polling.
struct DsGridBar {
  unsigned long long* ctr;
  unsigned long long target;
  unsigned long long* trace;
  int n;
  __device__ void init(unsigned long long* c, unsigned long long* tr) {
    ctr = c; trace = tr; n = 0;
    const unsigned long long v = ds_ld_acquire(c);
    target = v - v % gridDim.x;
  }
  __device__ void sync() {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      if (trace) trace[((size_t)n * gridDim.x + blockIdx.x) * 2] = ds_gtimer();
      ds_red_release(ctr, 1ULL);
      const long long t0 = clock64();
      while (ds_ld_relaxed(ctr) < target) {
        if (clock64() - t0 > 4000000000LL) {
          printf("kmbart decode_step: grid barrier timeout (block %d, barrier %d)\n", blockIdx.x, n);
          __trap();
        }
      }
      // no acquire fence: every load of data another CTA wrote in this launch is an L2 load (ld.global.cg / ld.relaxed),
      // issued after this loop exits; L1 only ever holds launch-invariant data (weights, biases, cross K/V, tables)
      if (trace) trace[((size_t)n * gridDim.x + blockIdx.x) * 2 + 1] = ds_gtimer();
    }
    ++n;
    __syncthreads();
  }
};

// ------------------------------------------------------------------ per-phase weight geometry
// Strip s, slot row r of a phase maps to virtual column vc = vc0 + s * strip_stride + r and weight row / output column
//   row_base + (vc / seg_len) * seg_stride + vc % seg_len.
struct DsPhaseW {
  const uint8_t* W;      // packed slots of this phase (kmb_decode_pack_weights): one contiguous sw x ROWB image per slot
  int slot0, slot_strip; // packed slot index of (strip s, chunk c) for this CTA = slot0 + s * slot_strip + c
  int n_strips, sw, n_chunks, sp;
  int vc0, strip_stride, seg_shift, seg_stride, row_base, kbase;
  __device__ __forceinline__ int col(int s, int r) const {
    const int vc = vc0 + s * strip_stride + r;
    return row_base + (vc >> seg_shift) * seg_stride + (vc & ((1 << seg_shift) - 1));
  }
  __device__ __forceinline__ int n_slots() const { return n_strips * n_chunks; }
};

struct DsGeom {   // row groups of the attention blocks
  int ncl, cid, kq, H, G, RG, h, g;
  bool owns_head;
};

template <int D>
__device__ __forceinline__ DsPhaseW ds_phase_w(const KmbDecodeStepC& p, const DsGeom& ge, int l, int kind) {
  using C = DsCfg<D>;
  const KmbDecodeLayer& L = p.layers[l];
  DsPhaseW w;
  w.n_chunks = 1; w.seg_shift = 30; w.seg_stride = 0; w.row_base = 0; w.kbase = ge.kq * C::KS;
  w.W = reinterpret_cast<const uint8_t*>(L.packed[kind]);
  if (kind == 0 || kind == 2) {   // head-owning clusters: q|k|v (192 virtual columns) or cross q (64 columns) of head h
    w.sw = kind == 0 ? 24 : 16;
    w.sp = kind == 0 ? 8 : 4;
    w.n_strips = ge.owns_head ? w.sp : 0;
    w.vc0 = 0; w.strip_stride = w.sw; w.seg_shift = 6; w.seg_stride = D; w.row_base = ge.h * 64;
    w.slot0 = ge.h * w.sp * DS_CL + ge.kq; w.slot_strip = DS_CL;             // packed order [head][strip][kq]
  } else {                        // generic: strips of the output columns dealt round-robin to the clusters
    const int N = kind == 4 ? p.F : D;
    w.sw = C::NTG * 8;
    w.sp = (kind == 4 && D != 128) ? 4 : 1;
    const int U = N / w.sw;
    w.n_strips = ge.cid < U ? (U - ge.cid + ge.ncl - 1) / ge.ncl : 0;
    w.vc0 = ge.cid * w.sw; w.strip_stride = ge.ncl * w.sw;
    if (kind == 5) { w.n_chunks = p.F / D; w.kbase = ge.kq * (p.F / DS_CL); }
    w.slot0 = (ge.cid * DS_CL + ge.kq) * w.n_chunks; w.slot_strip = ge.ncl * DS_CL * w.n_chunks;   // packed order [unit][kq][chunk]
  }
  return w;
}

// ------------------------------------------------------------------ weight ring (producer = warp 0)
// Every CTA knows the whole sequence of weight slots it will consume in the step.  The sequence ({source, bytes} per
// slot, consumption order) is tabulated in shared memory once at kernel start, so refilling the ring is one table read
// and one bulk copy per slot, issued by up to NSLOT lanes at once.
typedef unsigned long long DsSlotSrc;   // source address (64-byte aligned) | slot rows / 8
template <int D>
struct DsRing {
  uint32_t slots;      // shared address of slot 0
  uint64_t* full;
  const DsSlotSrc* tab;
  uint32_t total;      // slots of the whole step
  uint32_t cons_g;     // first slot of the current phase (all threads)
  uint32_t prod_g;     // next slot to issue            (warp 0)
};

template <int D>
__device__ void ds_ring_build(const KmbDecodeStepC& p, const DsGeom& ge, DsRing<D>& r, DsSlotSrc* tab) {
  using C = DsCfg<D>;
  int pre[7];
  pre[0] = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) pre[k + 1] = pre[k] + ds_phase_w<D>(p, ge, 0, k).n_slots();
  const int per_layer = pre[6];
  r.total = (uint32_t)(per_layer * p.L);
  if (r.total > (uint32_t)C::TAB_CAP) {
    if (threadIdx.x == 0) printf("kmbart decode_step: %u weight slots per CTA exceed the table (too few clusters: %d)\n", r.total, ge.ncl);
    __trap();
  }
  for (int i = threadIdx.x; i < (int)r.total; i += DS_THREADS) {
    const int l = i / per_layer, rem = i - l * per_layer;
    int kind = 0;
    while (rem >= pre[kind + 1]) ++kind;
    const int j = rem - pre[kind];
    const DsPhaseW w = ds_phase_w<D>(p, ge, l, kind);
    // consumption order: rounds of `sp` strips; inside a round chunk-major
    const int per_round = w.sp * w.n_chunks;
    const int s0 = (j / per_round) * w.sp;
    const int rj = j - s0 * w.n_chunks;
    const int nr = min(w.sp, w.n_strips - s0);
    const int c = rj / nr, strip = s0 + rj % nr;
    const uint32_t bytes = (uint32_t)w.sw * C::ROWB;
    tab[i] = (unsigned long long)(uintptr_t)(w.W + (int64_t)(w.slot0 + strip * w.slot_strip + c) * bytes) | (unsigned)(w.sw >> 3);
  }
  r.tab = tab;
}

template <int D>
__device__ __forceinline__ void ds_ring_advance(DsRing<D>& r) {   // warp 0
  using C = DsCfg<D>;
  const uint32_t lim = min(r.total, r.cons_g + (uint32_t)C::NSLOT);
  const uint32_t g = r.prod_g + (threadIdx.x & 31);
  if (g < lim) {   // one bulk copy per slot: the packed image already has the shared-memory row pitch
    const DsSlotSrc e = r.tab[g];
    const uint32_t slot = g % C::NSLOT, bytes = (uint32_t)(e & 63) * 8 * C::ROWB;
    mbar_arrive_expect_tx(&r.full[slot], bytes);
    ds_bulk_g2s(r.slots + slot * C::SLOT_BYTES, reinterpret_cast<const void*>((uintptr_t)(e & ~63ULL)), bytes, &r.full[slot]);
  }
  if (lim > r.prod_g) r.prod_g = lim;
}

// ------------------------------------------------------------------ row statistics -> mean / rstd
// stats[s][row][NP][2]: per-strip partial (sum, sum of squares) of the fp32 pre-LayerNorm row, written with plain stores.
// Four threads per row add the NP partials in a fixed order.
template <int D>
__device__ __forceinline__ void ds_row_stats(const float* st, int NP, int row0, int nrows, int rows_total, float* rs) {
  const int tid = threadIdx.x;
  const int r = tid >> 2, q = tid & 3;
  if (r < nrows) {   // nrows <= 64 -> threads 0..255, whole warps
    float s = 0.f, ss = 0.f;
    const int row = row0 + r;
    if (row < rows_total) {
      const float* sp = st + (int64_t)row * NP * 2;
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {             // float4 = two partials; NP <= 32: all loads in flight together
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((q + 4 * i) * 2 < NP) v[i] = ds_ldcg_f4(sp + (q + 4 * i) * 4);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { s += v[i].x + v[i].z; ss += v[i].y + v[i].w; }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2); ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    if (q == 0) {
      const float mean = s * (1.f / D);
      const float var = fmaxf(ss * (1.f / D) - mean * mean, 0.f);
      rs[r * 2] = mean;
      rs[r * 2 + 1] = rsqrtf(var + 1e-5f);
    }
  }
}

// ------------------------------------------------------------------ A-operand staging
struct DsA {
  const void* src;       // bf16 [rows, ld] or fp32 [rows, ld] (pre-LayerNorm)
  int64_t ld;
  int ln;                // 1: normalise on load with (rs, gamma, beta)
  const float *gamma, *beta;
};

// [MB rows x KS] slice starting at (m0, k0) -> bf16 tile (row pitch ROWB); rows >= rows_hi are zero.  Two halves so that
// the global loads of a slice are in flight while something else runs (row statistics, the MMAs of the previous K chunk).
template <int D, int MB>
struct DsStage {
  using C = DsCfg<D>;
  static constexpr int PPR = C::KS / 8;            // 16-byte pieces per row
  static constexpr int TOTAL = MB * PPR;
  static constexpr int NPT = (TOTAL + DS_THREADS - 1) / DS_THREADS;
  uint4 raw[NPT][2];
  __device__ __forceinline__ void load(const DsA& a, int m0, int rows_hi, int k0, int rot, bool skip = false) {
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      int idx = threadIdx.x + i * DS_THREADS + rot;
      if (idx >= TOTAL) idx -= TOTAL;
      const int r = idx / PPR, kc = idx % PPR;
      const int row = m0 + r;
      raw[i][0] = raw[i][1] = make_uint4(0, 0, 0, 0);
      if (threadIdx.x + i * DS_THREADS < TOTAL && row < rows_hi && !skip) {
        if (!a.ln) {
          raw[i][0] = ds_ldcg16(reinterpret_cast<const bf16*>(a.src) + (int64_t)row * a.ld + k0 + kc * 8);
        } else {
          const float* yp = reinterpret_cast<const float*>(a.src) + (int64_t)row * a.ld + k0 + kc * 8;
          raw[i][0] = ds_ldcg16(yp);
          raw[i][1] = ds_ldcg16(yp + 4);
        }
      }
    }
  }
  __device__ __forceinline__ void store(const DsA& a, int m0, int rows_hi, int rot, uint8_t* dst, const float* rs, const float* gb) const {
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      int idx = threadIdx.x + i * DS_THREADS + rot;
      if (idx >= TOTAL) idx -= TOTAL;
      const int r = idx / PPR, kc = idx % PPR;
      if (threadIdx.x + i * DS_THREADS < TOTAL) {
        uint4 out = raw[i][0];
        if (a.ln) {
          out = make_uint4(0, 0, 0, 0);
          if (m0 + r < rows_hi) {
            const float4 g0 = *reinterpret_cast<const float4*>(gb + kc * 8);
            const float4 g1 = *reinterpret_cast<const float4*>(gb + kc * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(gb + C::KS + kc * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(gb + C::KS + kc * 8 + 4);
            const float mean = rs[r * 2], rstd = rs[r * 2 + 1];
            const float y[8] = {__uint_as_float(raw[i][0].x), __uint_as_float(raw[i][0].y), __uint_as_float(raw[i][0].z), __uint_as_float(raw[i][0].w),
                                __uint_as_float(raw[i][1].x), __uint_as_float(raw[i][1].y), __uint_as_float(raw[i][1].z), __uint_as_float(raw[i][1].w)};
            out.x = ds_pack((y[0] - mean) * rstd * g0.x + b0.x, (y[1] - mean) * rstd * g0.y + b0.y);
            out.y = ds_pack((y[2] - mean) * rstd * g0.z + b0.z, (y[3] - mean) * rstd * g0.w + b0.w);
            out.z = ds_pack((y[4] - mean) * rstd * g1.x + b1.x, (y[5] - mean) * rstd * g1.y + b1.y);
            out.w = ds_pack((y[6] - mean) * rstd * g1.z + b1.z, (y[7] - mean) * rstd * g1.w + b1.w);
          }
        }
        *reinterpret_cast<uint4*>(dst + r * C::ROWB + kc * 16) = out;
      }
    }
  }
};

// ------------------------------------------------------------------ epilogue descriptors
enum { DS_EPI_Y = 0, DS_EPI_GELU = 1, DS_EPI_QKV = 2, DS_EPI_Q2 = 3 };
struct DsEpi {
  int kind;
  const float* bias;
  // DS_EPI_Y: y_new = acc + bias + LN(y_prev); partial row statistics of y_new
  const float* y_prev; float* y_new; const float *g_prev, *b_prev; float* stats_new; int NP;
  // DS_EPI_GELU
  bf16* h; int64_t ldh;
  // DS_EPI_QKV: k|v of the new token into the cache at position t; q|k|v (bf16-rounded) into the smem tile
  bf16* cache; int64_t cache_rs; int head;
  float* qkv_s;
};

// ------------------------------------------------------------------ the cluster GEMM
// out[rows row_lo..row_hi, strips of this cluster] = A[rows, K] . W^T with K split over the 4 CTAs of the cluster.
// NT: n8 tiles per strip; MT m16 tiles per row block (MB = 16 MT rows); SP strips in flight; KP k-parts inside a CTA.
// After the DSMEM reduction CTA kq owns rows [m0 + kq*MB/4, +MB/4) of the block; `post` runs right after the epilogue of
// every row block (attention of the head-owning blocks).
template <int D, int NT, int MT, int SP, int KP, typename Mid, typename Post>
__device__ void ds_gemm(const KmbDecodeStepC& p, const DsGeom& ge, DsRing<D>& ring, uint8_t* smem, const DsPhaseW& w, const DsA& a,
                        const float* stats_in, int NP_in, const DsEpi& e, const float* stats_res, int row_lo, int row_hi, Mid mid, Post post,
                        unsigned long long* ev) {
  using C = DsCfg<D>;
  constexpr int MB = MT * 16, RPC = MB / DS_CL, SW = NT * 8;
  constexpr int KPS = C::KSTEPS / KP;       // k16 steps per warp and chunk
  constexpr int C4 = SW / 4;                // float4 column groups per strip (<= 8); a row's groups sit in 8 consecutive lanes
  static_assert(C::KSTEPS % KP == 0, "K parts");
  static_assert(KP * SP * MB * SW * 4 <= C::PARTB, "partial tiles");
  static_assert(RPC * SP * 8 <= DS_THREADS, "one epilogue item per thread");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp % MT, sp = (warp / MT) % SP, kp = warp / (MT * SP);
  const bool mma_warp = warp < MT * SP * KP;
  const int gq = lane >> 2, tg = lane & 3;
  uint8_t* abuf0 = smem + C::OFF_A0;
  uint8_t* abuf1 = smem + C::OFF_A1;
  float* part = reinterpret_cast<float*>(smem + C::OFF_A1);
  float* rs = reinterpret_cast<float*>(smem + C::OFF_RS);
  float* stsc = reinterpret_cast<float*>(smem + C::OFF_ST);
  float* gb = reinterpret_cast<float*>(smem + C::OFF_GB);
  const uint32_t part_u32 = smem_u32(part);
  const int rot = (ge.cid * 37) % (MB * C::KS / 8);
  if (w.n_strips == 0) {
    if (warp == 0) ds_ring_advance<D>(ring);
    return;
  }
  DS_EV(0);
  const int rbase = ge.kq * RPC;            // after the reduction this CTA finishes rows [m0 + rbase, +RPC) of a row block
  const int n_mblocks = (row_hi - row_lo + MB - 1) / MB;
  DsStage<D, MB> st;
  for (int mb = 0; mb < n_mblocks; ++mb) {
    const int m0 = row_lo + mb * MB;
    // everything that does not depend on this block's accumulators is requested first: the A slice, the row statistics
    // (LayerNorm on load, or the residual rows of a y-epilogue: never both in one phase)
    const bool xskip = (p.flags & 8) || ((p.flags & 16) && ge.cid >= 8);   // timing experiments only (wrong results)
    st.load(a, m0, row_hi, w.kbase, rot, xskip);
    if (a.ln) {
      if (mb == 0 && threadIdx.x < C::KS / 2) {   // LayerNorm weight / bias of this CTA's K range (chunked K never uses LN)
        const int j = threadIdx.x % (C::KS / 4), sel = threadIdx.x / (C::KS / 4);
        reinterpret_cast<float4*>(gb)[sel * (C::KS / 4) + j] = __ldg(reinterpret_cast<const float4*>((sel ? a.beta : a.gamma) + w.kbase) + j);
      }
      ds_row_stats<D>(stats_in, NP_in, m0, MB, row_hi, rs);
    } else if (e.kind == DS_EPI_Y) ds_row_stats<D>(stats_res, e.NP, m0 + rbase, RPC, row_hi, rs);
    __syncthreads();
    DS_EV(1);
    st.store(a, m0, row_hi, rot, abuf0, rs, gb);
    // Refill the weight slots the PREVIOUS phase released only now, when this phase's first operand loads have returned:
    // a burst of bulk copies issued at the end of a phase sits in front of the next phase's demand loads in the memory
    // system (up to 10 MB chip-wide = 1.5 us of HBM time) and was the largest term of every phase (profiles/r02f).
    if (mb == 0) {
      if (warp == 0) ds_ring_advance<D>(ring);
      mid();   // L2 prefetches for later phases: same reasoning
    }
    for (int s0 = 0; s0 < w.n_strips; s0 += SP) {
      const int nr = min(SP, w.n_strips - s0);
      // epilogue operands of this thread's output item (bias, residual row, LayerNorm weights): in flight during the MMAs
      const int items = RPC * nr * 8;
      const int it = threadIdx.x;
      const int c4 = it & 7;
      int es, er;
      if ((nr & (nr - 1)) == 0) { const int sh = 31 - __clz(nr); es = (it >> 3) & (nr - 1); er = it >> (3 + sh); }
      else { es = (it >> 3) % nr; er = (it >> 3) / nr; }
      const bool ok = it < items && c4 < C4;
      const int erow = m0 + rbase + er;
      const bool live = ok && erow < row_hi;
      const int ecol = live ? w.col(s0 + es, c4 * 4) : 0;
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), yp4 = bias4, g4 = bias4, b4 = bias4;
      if (live) {
        bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + ecol));
        if (e.kind == DS_EPI_Y) {
          yp4 = ds_ldcg_f4(e.y_prev + (int64_t)erow * D + ecol);
          g4 = __ldg(reinterpret_cast<const float4*>(e.g_prev + ecol));
          b4 = __ldg(reinterpret_cast<const float4*>(e.b_prev + ecol));
        }
      }
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      const bool active = mma_warp && sp < nr;
      for (int c = 0; c < w.n_chunks; ++c) {
        uint8_t* ab = (c & 1) ? abuf1 : abuf0;
        __syncthreads();                     // A chunk c is in shared memory; chunk c-1's MMAs are done
        DS_EV(2);
        if (c + 1 < w.n_chunks) st.load(a, m0, row_hi, w.kbase + (c + 1) * C::KS, rot, xskip);
        if (active) {
          const uint32_t g = ring.cons_g + s0 * w.n_chunks + c * nr + sp;
          mbar_wait(&ring.full[g % C::NSLOT], (g / C::NSLOT) & 1);
          DS_EV(3);
          const uint32_t wb = ring.slots + (g % C::NSLOT) * C::SLOT_BYTES;
          const uint32_t a_addr = smem_u32(ab) + (mt * 16 + (lane & 15)) * C::ROWB + (lane >> 4) * 16;
          const uint32_t b_addr = wb + ((lane & 7) + (lane >> 4) * 8) * C::ROWB + ((lane >> 3) & 1) * 16;
          // fragments of k-step k+1 are requested before the MMAs of step k (the asm statements keep their order)
          uint32_t af[2][4], bfr[2][NT][2];
          auto frags = [&](int k, int b) {
            const int ks = kp * KPS + k;
            ds_ldsm4(a_addr + ks * 32, af[b][0], af[b][1], af[b][2], af[b][3]);
#pragma unroll
            for (int j = 0; j + 1 < NT; j += 2) ds_ldsm4(b_addr + j * 8 * C::ROWB + ks * 32, bfr[b][j][0], bfr[b][j][1], bfr[b][j + 1][0], bfr[b][j + 1][1]);
            if (NT & 1) ds_ldsm2(wb + ((NT - 1) * 8 + (lane & 7)) * C::ROWB + ((lane >> 3) & 1) * 16 + ks * 32, bfr[b][NT - 1][0], bfr[b][NT - 1][1]);
          };
          frags(0, 0);
#pragma unroll
          for (int k = 0; k < KPS; ++k) {
            if (k + 1 < KPS) frags(k + 1, (k + 1) & 1);
#pragma unroll
            for (int j = 0; j < NT; ++j) ds_mma(acc[j], af[k & 1], bfr[k & 1][j][0], bfr[k & 1][j][1]);
          }
        }
        // chunk c+1 goes into the other buffer: its last readers (the MMAs of chunk c-1) passed the barrier above
        if (c + 1 < w.n_chunks) st.store(a, m0, row_hi, rot, ((c + 1) & 1) ? abuf1 : abuf0, rs, gb);
      }
      DS_EV(4);
      if (w.n_chunks > 1) __syncthreads();   // the partial tiles alias the second A buffer
      // partial tile [kp][sp][row][SW]
      if (active) {
        float* pt = part + ((kp * SP + sp) * MB + mt * 16) * SW;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          *reinterpret_cast<float2*>(pt + gq * SW + j * 8 + 2 * tg) = make_float2(acc[j][0], acc[j][1]);
          *reinterpret_cast<float2*>(pt + (gq + 8) * SW + j * 8 + 2 * tg) = make_float2(acc[j][2], acc[j][3]);
        }
      }
      if (KP > 1) {
        __syncthreads();
        constexpr int V4 = SP * MB * SW / 4;
        for (int i = threadIdx.x; i < V4; i += DS_THREADS) {
          float4 v = reinterpret_cast<float4*>(part)[i];
#pragma unroll
          for (int q = 1; q < KP; ++q) {
            const float4 u = reinterpret_cast<float4*>(part)[q * V4 + i];
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
          }
          reinterpret_cast<float4*>(part)[i] = v;
        }
      }
      DS_EV(5);
      cluster_sync_all();
      DS_EV(6);
      // ---- reduction over the cluster + epilogue: one output item (4 columns of one row) per thread
      if (it < items) {   // whole warps: items is a multiple of 64
        float4 v = bias4;
        if (ok) {
          const uint32_t off = part_u32 + ((es * MB + rbase + er) * SW + c4 * 4) * 4;
#pragma unroll
          for (int q = 0; q < DS_CL; ++q) {
            const float4 u = ds_ld_cluster_f4(mapa_shared(off, q));
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
          }
        }
        if (e.kind == DS_EPI_Y) {
          float s1 = 0.f, s2 = 0.f;
          if (live) {
            const float mean = rs[er * 2], rstd = rs[er * 2 + 1];
            v.x += (yp4.x - mean) * rstd * g4.x + b4.x;
            v.y += (yp4.y - mean) * rstd * g4.y + b4.y;
            v.z += (yp4.z - mean) * rstd * g4.z + b4.z;
            v.w += (yp4.w - mean) * rstd * g4.w + b4.w;
            *reinterpret_cast<float4*>(e.y_new + (int64_t)erow * D + ecol) = v;
            s1 = v.x + v.y + v.z + v.w;
            s2 = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          }
          // per-(row, strip) partial statistics, summed in a fixed order by the first lane of the row's 8-lane group
          stsc[(it >> 3) * 16 + c4 * 2] = s1;
          stsc[(it >> 3) * 16 + c4 * 2 + 1] = s2;
          __syncwarp();
          if (live && c4 == 0) {
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int q = 0; q < C4; ++q) { t1 += stsc[(it >> 3) * 16 + q * 2]; t2 += stsc[(it >> 3) * 16 + q * 2 + 1]; }
            *reinterpret_cast<float2*>(e.stats_new + ((int64_t)erow * e.NP + ecol / SW) * 2) = make_float2(t1, t2);
          }
        } else if (e.kind == DS_EPI_GELU) {
          if (live) {
            uint2 o;
            o.x = ds_pack(gelu_erf(v.x), gelu_erf(v.y));
            o.y = ds_pack(gelu_erf(v.z), gelu_erf(v.w));
            *reinterpret_cast<uint2*>(e.h + (int64_t)erow * e.ldh + ecol) = o;
          }
        } else if (ok) {   // DS_EPI_QKV / DS_EPI_Q2: the projections of this CTA's rows stay in shared memory for the attention
          const int vc = w.vc0 + (s0 + es) * w.strip_stride + c4 * 4;
          const float4 o = make_float4(ds_round_bf16(v.x), ds_round_bf16(v.y), ds_round_bf16(v.z), ds_round_bf16(v.w));
          *reinterpret_cast<float4*>(e.qkv_s + er * DS_QKV_LD + vc) = o;
          if (live && e.kind == DS_EPI_QKV && vc >= 64) {
            uint2 pk;
            pk.x = ds_pack(o.x, o.y);
            pk.y = ds_pack(o.z, o.w);
            *reinterpret_cast<uint2*>(e.cache + (int64_t)erow * e.cache_rs + (vc / 64) * D + e.head * 64 + vc % 64) = pk;
          }
        }
      }
      const bool more = (s0 + SP < w.n_strips) || (mb + 1 < n_mblocks);
      __syncthreads();
      DS_EV(8);
      post(m0 + rbase, RPC, ev);
      DS_EV(9);
      if (more) cluster_sync_all();   // peers are done reading this CTA's partial tiles before they are overwritten
    }
  }
}

struct DsNoMid {
  __device__ __forceinline__ void operator()() const {}
};
struct DsNoPost {
  __device__ __forceinline__ void operator()(int, int, unsigned long long*) const {}
};

// ------------------------------------------------------------------ attention of 8 rows x one head (two warps per row)
struct DsAttn {
  const bf16 *k, *v;          // global K / V of head 0 (element pointers); head h at + h*64
  int64_t kv_ss, kv_ps;       // slot (cache row / sample) stride, position stride
  const int* slot_tbl; int64_t tbl_ld;
  int row_div;
  const uint8_t* key_pad; int64_t pad_ld;
  int T;                      // keys; self-attention: position T-1 is the new token (from the smem tile)
  int self;
  float scale;
  bf16* o; int64_t o_rs;
  int head;
};

__device__ __forceinline__ void ds_unpack16(const uint4& a, const uint4& b, float (&f)[16]) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[2 * i] = ds_lo(w[i]); f[2 * i + 1] = ds_hi(w[i]); }
}

template <int D>
__device__ void ds_attend(const DsAttn& p, uint8_t* smem, int row0, int nrows, int row_hi, unsigned long long* ev) {
  // Two warps per (row, head): each takes half of the keys, 32 keys per iteration with the K AND V lines of the block
  // requested together (one memory round trip per block), online softmax across blocks, flash-style merge of the halves.
  //   scores: lane = (key phase kq = lane >> 2, 16-dim slice sub = lane & 3), keys pos0 + 8 j + kq, j < 4
  //   P V   : lane = (key group kg = lane >> 3, 8-dim slice dc = lane & 7), keys pos0 + 4 j + kg, j < 8
  using C = DsCfg<D>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = warp >> 1, half = warp & 1;
  const float* qkv = reinterpret_cast<const float*>(smem + C::OFF_QKV) + u * DS_QKV_LD;
  float* mg = reinterpret_cast<float*>(smem + C::OFF_MG) + u * 68;
  const int row = row0 + u;
  const bool live = u < nrows && row < row_hi;
  const int sub = lane & 3, kq = lane >> 2;
  const int kg = lane >> 3, dc = lane & 7;
  const int T = p.T, Th = (T + 1) >> 1;
  const int k_lo = half ? Th : 0, k_hi = half ? T : Th;
  const int t_new = p.self ? T - 1 : -1;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  if (live) {
    const int bslot = row / p.row_div;
    const bf16* kb = p.k + p.head * 64 + sub * 16;
    const bf16* vb = p.v + p.head * 64 + dc * 8;
    for (int pos0 = k_lo; pos0 < k_hi; pos0 += 32) {
      bool ok[4];
      uint4 ka[4], kb2[4], vv[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pos = pos0 + 8 * j + kq;
        ok[j] = pos < k_hi;
        ka[j] = kb2[j] = make_uint4(0, 0, 0, 0);
        if (ok[j] && pos != t_new) {
          const int slot = p.slot_tbl ? __ldg(p.slot_tbl + (int64_t)row * p.tbl_ld + pos) : bslot;
          const bf16* kp = kb + (int64_t)slot * p.kv_ss + (int64_t)pos * p.kv_ps;
          if (p.self) { ka[j] = ds_ldcg16(kp); kb2[j] = ds_ldcg16(kp + 8); }
          else { ka[j] = __ldg(reinterpret_cast<const uint4*>(kp)); kb2[j] = __ldg(reinterpret_cast<const uint4*>(kp + 8)); }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int pos = pos0 + 4 * j + kg;
        vv[j] = make_uint4(0, 0, 0, 0);
        if (pos < k_hi && pos != t_new) {
          const int slot = p.slot_tbl ? __ldg(p.slot_tbl + (int64_t)row * p.tbl_ld + pos) : bslot;
          const bf16* vp = vb + (int64_t)slot * p.kv_ss + (int64_t)pos * p.kv_ps;
          vv[j] = p.self ? ds_ldcg16(vp) : __ldg(reinterpret_cast<const uint4*>(vp));
        }
      }
      if (p.key_pad) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ok[j] && __ldg(p.key_pad + (int64_t)bslot * p.pad_ld + pos0 + 8 * j + kq)) ok[j] = false;
      }
      float qf[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) qf[i] = qkv[sub * 16 + i];
      float sc[4];
      float bm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pos = pos0 + 8 * j + kq;
        float kf[16];
        if (pos == t_new) {
#pragma unroll
          for (int i = 0; i < 16; ++i) kf[i] = qkv[64 + sub * 16 + i];
        } else {
          ds_unpack16(ka[j], kb2[j], kf);
        }
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc = fmaf(qf[i], kf[i], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        sc[j] = ok[j] ? acc * p.scale : -INFINITY;
        bm = fmaxf(bm, sc[j]);
      }
      bm = warp_max(bm);
      const float m_new = fmaxf(m_run, bm);
      const float corr = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_new);
      float pj[4], bs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pj[j] = (sc[j] == -INFINITY) ? 0.f : __expf(sc[j] - m_new);
        bs += pj[j];
      }
      bs = warp_sum(sub == 0 ? bs : 0.f);        // every key's probability is replicated in the 4 lanes of its phase
      l_run = l_run * corr + bs;
      m_run = m_new;
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] *= corr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // key index 4 j + kg inside the block = 8 (j >> 1) + (4 (j & 1) + kg): register j >> 1 of the lanes of phase 4 (j & 1) + kg
        const float pr = __shfl_sync(0xffffffffu, pj[j >> 1], (4 * (j & 1) + kg) * 4);
        const int pos = pos0 + 4 * j + kg;
        if (pos == t_new) {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = fmaf(pr, qkv[128 + dc * 8 + i], o[i]);
        } else {
          const uint32_t w4[4] = {vv[j].x, vv[j].y, vv[j].z, vv[j].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            o[2 * i] = fmaf(pr, ds_lo(w4[i]), o[2 * i]);
            o[2 * i + 1] = fmaf(pr, ds_hi(w4[i]), o[2 * i + 1]);
          }
        }
      }
    }
    DS_EV(12);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
      o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
    }
    if (half == 1) {
      if (lane == 0) { mg[64] = m_run; mg[65] = l_run; }
      if (kg == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mg[dc * 8 + i] = o[i];
      }
    }
  }
  __syncthreads();
  DS_EV(13);
  if (live && half == 0) {
    const float m1 = mg[64], l1 = mg[65];
    const float m = fmaxf(m_run, m1);
    // a half with no unmasked key has max = -inf and sum = 0: weight 0; both empty -> 0/0 = NaN like the reference
    const float w0 = (m_run == -INFINITY) ? 0.f : __expf(m_run - m);
    const float w1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - m);
    const float inv = 1.f / (l_run * w0 + l1 * w1);
    if (kg == 0) {
      float r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = (o[i] * w0 + mg[dc * 8 + i] * w1) * inv;
      uint4 pk;
      pk.x = ds_pack(r[0], r[1]); pk.y = ds_pack(r[2], r[3]); pk.z = ds_pack(r[4], r[5]); pk.w = ds_pack(r[6], r[7]);
      *reinterpret_cast<uint4*>(p.o + (int64_t)row * p.o_rs + p.head * 64 + dc * 8) = pk;
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ LayerNorm parameters of residual-stream state s
__device__ __forceinline__ void ds_ln_params(const KmbDecodeStepC& p, int s, const float*& g, const float*& b) {
  if (s == 0) { g = p.lne_g; b = p.lne_b; return; }
  const KmbDecodeLayer& L = p.layers[(s - 1) / 3];
  const int w = (s - 1) % 3;
  g = w == 0 ? L.ln1_g : w == 1 ? L.ln2_g : L.ln3_g;
  b = w == 0 ? L.ln1_b : w == 1 ? L.ln2_b : L.ln3_b;
}

template <int D>
struct DsAttnPost {
  DsAttn at;
  uint8_t* smem;
  int row_hi;
  __device__ __forceinline__ void operator()(int row0, int nrows, unsigned long long* ev) const { ds_attend<D>(at, smem, row0, nrows, row_hi, ev); }
};


// ------------------------------------------------------------------ weight packing (once per weight version)
// Slot images in exactly the layout the ring wants: sw rows of ROWB bytes (KS bf16 + 16 bytes of padding), one image per
// (strip / unit, K quarter, K chunk), so that a slot is ONE cp.async.bulk (per-row copies of 384 bytes were bound by the
// bulk-copy issue rate: 528 copies per layer and CTA, profiles/r02b_decode_trace.txt).
//   kind 0 (q|k|v of a head): [H][8 strips of 24 virtual columns][4 kq]      kind 2 (cross q): [H][4 strips of 16][4 kq]
//   kinds 1, 3, 4, 5 (out-proj, cross out-proj, fc1, fc2): [N / sw units][4 kq][K / d chunks]
template <int D>
__global__ void decode_pack_kernel(KmbDecodeLayer L, int H, int F, uint8_t* out, int64_t off0, int64_t off1, int64_t off2, int64_t off3,
                                   int64_t off4, int64_t off5) {
  using C = DsCfg<D>;
  const int kind = blockIdx.y;
  const int64_t offs[6] = {off0, off1, off2, off3, off4, off5};
  const bf16* W = (const bf16*)(kind == 0 ? L.w_qkv : kind == 1 ? L.w_o : kind == 2 ? L.w_cq : kind == 3 ? L.w_co : kind == 4 ? L.w_fc1 : L.w_fc2);
  const bool head = kind == 0 || kind == 2;
  const int sw = kind == 0 ? 24 : kind == 2 ? 16 : C::NTG * 8;
  const int K = kind == 5 ? F : D, N = kind == 4 ? F : D;
  const int n_chunks = K / D;
  const int n_slots = head ? H * (kind == 0 ? 8 : 4) * DS_CL : (N / sw) * DS_CL * n_chunks;
  constexpr int PPR = C::ROWB / 16;            // 16-byte pieces per padded row
  for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
    int kq, chunk = 0, row_first, strip = 0, h = 0;
    if (head) { kq = slot % DS_CL; strip = (slot / DS_CL) % (kind == 0 ? 8 : 4); h = slot / (DS_CL * (kind == 0 ? 8 : 4)); row_first = 0; }
    else { chunk = slot % n_chunks; kq = (slot / n_chunks) % DS_CL; row_first = (slot / (n_chunks * DS_CL)) * sw; }
    uint8_t* dst = out + offs[kind] + (int64_t)slot * sw * C::ROWB;
    for (int i = threadIdx.x; i < sw * PPR; i += blockDim.x) {
      const int r = i / PPR, pc = i % PPR;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (pc < C::KS / 8) {
        int wrow;
        if (head) { const int vc = strip * sw + r; wrow = h * 64 + (vc / 64) * D + vc % 64; }
        else wrow = row_first + r;
        const int k = kq * (K / DS_CL) + chunk * C::KS + pc * 8;
        v = *reinterpret_cast<const uint4*>(W + (int64_t)wrow * K + k);
      }
      *reinterpret_cast<uint4*>(dst + r * C::ROWB + pc * 16) = v;
    }
  }
}

// ------------------------------------------------------------------ the step
template <int D>
__global__ void __launch_bounds__(DS_THREADS, 1) decode_step_kernel(const __grid_constant__ KmbDecodeStepC p) {
  using C = DsCfg<D>;
  constexpr int NTG = C::NTG;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DsGeom ge;
  ge.ncl = gridDim.x / DS_CL;
  ge.cid = blockIdx.x / DS_CL;
  ge.kq = (int)cluster_ctarank();
  ge.H = p.H;
  {
    const int RS = max(1, ge.ncl / p.H);
    int rg = (p.rows + RS - 1) / RS;
    rg = max(8, (rg + 7) & ~7);
    ge.RG = rg;
    ge.G = (p.rows + rg - 1) / rg;
    ge.h = ge.cid % p.H;
    ge.g = ge.cid / p.H;
    ge.owns_head = ge.g < ge.G;
  }
  const int NP = D / (NTG * 8);     // statistics partials per row (= generic strips of a d-wide output)
  DsRing<D> ring;
  ring.slots = smem_u32(smem + C::OFF_RING);
  ring.full = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  ring.cons_g = 0; ring.prod_g = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NSLOT; ++i) mbar_init(&ring.full[i], 1);
    fence_mbar_init();
  }
  ds_ring_build<D>(p, ge, ring, reinterpret_cast<DsSlotSrc*>(smem + C::OFF_TAB));
  __syncthreads();
  if (warp == 0) ds_ring_advance<D>(ring);   // the weight stream starts before anything else
  DsGridBar bar;
  bar.init(p.barrier, p.trace);

  const int gw = blockIdx.x * DS_WARPS + warp, gws = gridDim.x * DS_WARPS;
  // phase 0: y0 = token embedding * scale + learned position (pre-LayerNorm), statistics of state 0
  for (int row = gw; row < p.rows; row += gws) {
    const float* te = p.tok_emb + (int64_t)p.ids[row] * D;
    const float* pe = p.pos_emb + (int64_t)p.pos_row * D;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < D / 128; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 a = __ldg(reinterpret_cast<const float4*>(te + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(pe + c));
      const float4 x = make_float4(a.x * p.embed_scale + b.x, a.y * p.embed_scale + b.y, a.z * p.embed_scale + b.z, a.w * p.embed_scale + b.w);
      *reinterpret_cast<float4*>(p.y0 + (int64_t)row * D + c) = x;
      s1 += x.x + x.y + x.z + x.w;
      s2 += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    float* st = p.stats + (int64_t)row * NP * 2;
    for (int i = lane; i < NP; i += 32) *reinterpret_cast<float2*>(st + i * 2) = i == 0 ? make_float2(s1, s2) : make_float2(0.f, 0.f);
  }
  bar.sync();

  const int64_t stat_stride = (int64_t)p.rows * NP * 2;
  float* qkv_s = reinterpret_cast<float*>(smem + C::OFF_QKV);
  unsigned long long* evbase = p.trace ? p.trace + (size_t)(1 + 6 * p.L) * gridDim.x * 2 + (size_t)blockIdx.x * 96 : nullptr;
#define EVP(k) ((evbase && l == 1) ? evbase + (k) * 16 : nullptr)
  for (int l = 0; l < p.L; ++l) {
    const KmbDecodeLayer& L = p.layers[l];
    const int s_in = 3 * l;                       // residual-stream state entering the layer
    float* ybuf[2] = {p.y0, p.y1};
    const int row_lo = ge.g * ge.RG, row_hi = min(p.rows, row_lo + ge.RG);
    // ---------------- phase A: LN(s_in) -> q|k|v of (head, row group) -> self-attention -> ctx
    {
      const DsPhaseW w = ds_phase_w<D>(p, ge, l, 0);
      if (ge.owns_head) {
        auto prefetch_kv = [&]() {
        // cross K/V of the rows this CTA attends in phase C: pull the lines into L2 now
          const bf16* kv = (const bf16*)L.cross_kv;
          for (int mb = row_lo; mb < row_hi && !(p.flags & KMB_DECODE_NO_CROSS_PREFETCH); mb += 32) {
            const int r0 = mb + ge.kq * 8, r1 = min(row_hi, r0 + 8);
            if (r0 < r1) {
              const int s_lo = r0 / p.row_div, s_hi = (r1 - 1) / p.row_div;
              const int n = (s_hi - s_lo + 1) * p.Se * 2;
              for (int i = threadIdx.x; i < n; i += DS_THREADS) {
                const int kvsel = i & 1, pos = (i >> 1) % p.Se, smp = s_lo + (i >> 1) / p.Se;
                ds_prefetch_l2(kv + ((int64_t)smp * p.Se + pos) * 2 * D + kvsel * D + ge.h * 64);
              }
            }
          }
          if (p.t > 0 && !(p.flags & KMB_DECODE_NO_SELF_PREFETCH)) {   // ... and the cached self-attention K/V of the same rows
            const bf16* base = (const bf16*)L.cache;
            const int nblk = (row_hi - row_lo + 31) / 32;
            const int n = nblk * 8 * p.t * 2;
            for (int i = threadIdx.x; i < n; i += DS_THREADS) {
              const int kvsel = i & 1, pos = (i >> 1) % p.t, rr = (i >> 1) / p.t;
              const int row = row_lo + (rr >> 3) * 32 + ge.kq * 8 + (rr & 7);
              if (row < row_hi) {
                const int slot = p.slot_tbl ? __ldg(p.slot_tbl + (int64_t)row * p.max_len + pos) : row;
                ds_prefetch_l2(base + (int64_t)slot * p.max_len * 3 * D + (int64_t)pos * 3 * D + (1 + kvsel) * D + ge.h * 64);
              }
            }
          }
        };
        DsA a;
        a.src = ybuf[s_in & 1]; a.ld = D; a.ln = 1;
        ds_ln_params(p, s_in, a.gamma, a.beta);
        DsEpi e = {};
        e.kind = DS_EPI_QKV; e.bias = L.b_qkv; e.qkv_s = qkv_s; e.head = ge.h;
        e.cache = (bf16*)L.cache + (int64_t)p.t * 3 * D; e.cache_rs = (int64_t)p.max_len * 3 * D;
        DsAttnPost<D> post;
        post.smem = smem; post.row_hi = row_hi;
        const bf16* base = (const bf16*)L.cache;
        post.at.k = base + D; post.at.v = base + 2 * D; post.at.kv_ss = (int64_t)p.max_len * 3 * D; post.at.kv_ps = 3 * D;
        post.at.slot_tbl = p.slot_tbl; post.at.tbl_ld = p.max_len; post.at.row_div = 1; post.at.key_pad = nullptr; post.at.pad_ld = 0;
        post.at.T = p.t + 1; post.at.self = 1; post.at.scale = p.attn_scale; post.at.o = (bf16*)p.ctx; post.at.o_rs = D; post.at.head = ge.h;
        ds_gemm<D, 3, 2, 8, 1>(p, ge, ring, smem, w, a, p.stats + s_in * stat_stride, NP, e, nullptr, row_lo, row_hi, prefetch_kv, post, EVP(0));
      } else if (warp == 0) {
        ds_ring_advance<D>(ring);
      }
      ring.cons_g += w.n_slots();   // released; refilled inside the next phase
    }
    { unsigned long long* ev = EVP(0); DS_EV(10); }
    bar.sync();
    { unsigned long long* ev = EVP(0); DS_EV(11); }
    // ---------------- phase B: y(s_in+1) = LN(s_in) + out_proj(ctx) (+ statistics)
    {
      const DsPhaseW w = ds_phase_w<D>(p, ge, l, 1);
      DsA a; a.src = p.ctx; a.ld = D; a.ln = 0; a.gamma = a.beta = nullptr;
      DsEpi e = {};
      e.kind = DS_EPI_Y; e.bias = L.b_o; e.y_prev = ybuf[s_in & 1]; e.y_new = ybuf[(s_in + 1) & 1];
      ds_ln_params(p, s_in, e.g_prev, e.b_prev);
      e.stats_new = p.stats + (s_in + 1) * stat_stride; e.NP = NP;
      ds_gemm<D, NTG, 4, 1, (D == 128 ? 2 : 4)>(p, ge, ring, smem, w, a, nullptr, NP, e, p.stats + s_in * stat_stride, 0, p.rows, DsNoMid(), DsNoPost(), EVP(1));
      if (p.flags & 4) {   // experiment: the same phase again with warm instruction / constant caches
        __syncthreads();
        cluster_sync_all();
        ds_gemm<D, NTG, 4, 1, (D == 128 ? 2 : 4)>(p, ge, ring, smem, w, a, nullptr, NP, e, p.stats + s_in * stat_stride, 0, p.rows, DsNoMid(), DsNoPost(), EVP(1));
      }
      ring.cons_g += w.n_slots();   // released; refilled inside the next phase
    }
    { unsigned long long* ev = EVP(1); DS_EV(10); }
    bar.sync();
    { unsigned long long* ev = EVP(1); DS_EV(11); }
    // ---------------- phase C: LN(s_in+1) -> cross q of (head, row group) -> cross-attention -> ctx
    {
      const DsPhaseW w = ds_phase_w<D>(p, ge, l, 2);
      if (ge.owns_head) {
        DsA a;
        a.src = ybuf[(s_in + 1) & 1]; a.ld = D; a.ln = 1;
        ds_ln_params(p, s_in + 1, a.gamma, a.beta);
        DsEpi e = {};
        e.kind = DS_EPI_Q2; e.bias = L.b_cq; e.qkv_s = qkv_s; e.head = ge.h;
        DsAttnPost<D> post;
        post.smem = smem; post.row_hi = row_hi;
        const bf16* kv = (const bf16*)L.cross_kv;
        post.at.k = kv; post.at.v = kv + D; post.at.kv_ss = (int64_t)p.Se * 2 * D; post.at.kv_ps = 2 * D;
        post.at.slot_tbl = nullptr; post.at.tbl_ld = 0; post.at.row_div = p.row_div; post.at.key_pad = p.key_pad; post.at.pad_ld = p.Se;
        post.at.T = p.Se; post.at.self = 0; post.at.scale = p.attn_scale; post.at.o = (bf16*)p.ctx; post.at.o_rs = D; post.at.head = ge.h;
        ds_gemm<D, 2, 2, 4, 2>(p, ge, ring, smem, w, a, p.stats + (s_in + 1) * stat_stride, NP, e, nullptr, row_lo, row_hi, DsNoMid(), post, EVP(2));
      } else if (warp == 0) {
        ds_ring_advance<D>(ring);
      }
      ring.cons_g += w.n_slots();   // released; refilled inside the next phase
    }
    { unsigned long long* ev = EVP(2); DS_EV(10); }
    bar.sync();
    { unsigned long long* ev = EVP(2); DS_EV(11); }
    // ---------------- phase D: y(s_in+2) = LN(s_in+1) + out_proj(ctx)
    {
      const DsPhaseW w = ds_phase_w<D>(p, ge, l, 3);
      DsA a; a.src = p.ctx; a.ld = D; a.ln = 0; a.gamma = a.beta = nullptr;
      DsEpi e = {};
      e.kind = DS_EPI_Y; e.bias = L.b_co; e.y_prev = ybuf[(s_in + 1) & 1]; e.y_new = ybuf[(s_in + 2) & 1];
      ds_ln_params(p, s_in + 1, e.g_prev, e.b_prev);
      e.stats_new = p.stats + (s_in + 2) * stat_stride; e.NP = NP;
      ds_gemm<D, NTG, 4, 1, (D == 128 ? 2 : 4)>(p, ge, ring, smem, w, a, nullptr, NP, e, p.stats + (s_in + 1) * stat_stride, 0, p.rows, DsNoMid(), DsNoPost(), EVP(3));
      ring.cons_g += w.n_slots();   // released; refilled inside the next phase
    }
    { unsigned long long* ev = EVP(3); DS_EV(10); }
    bar.sync();
    { unsigned long long* ev = EVP(3); DS_EV(11); }
    // ---------------- phase E: h = gelu(fc1(LN(s_in+2)))
    {
      const DsPhaseW w = ds_phase_w<D>(p, ge, l, 4);
      DsA a;
      a.src = ybuf[(s_in + 2) & 1]; a.ld = D; a.ln = 1;
      ds_ln_params(p, s_in + 2, a.gamma, a.beta);
      DsEpi e = {};
      e.kind = DS_EPI_GELU; e.bias = L.b_fc1; e.h = (bf16*)p.h; e.ldh = p.F;
      if (D == 128) ds_gemm<D, NTG, 4, 1, 2>(p, ge, ring, smem, w, a, p.stats + (s_in + 2) * stat_stride, NP, e, nullptr, 0, p.rows, DsNoMid(), DsNoPost(), EVP(4));
      else ds_gemm<D, NTG, 4, 4, 1>(p, ge, ring, smem, w, a, p.stats + (s_in + 2) * stat_stride, NP, e, nullptr, 0, p.rows, DsNoMid(), DsNoPost(), EVP(4));
      ring.cons_g += w.n_slots();   // released; refilled inside the next phase
    }
    { unsigned long long* ev = EVP(4); DS_EV(10); }
    bar.sync();
    { unsigned long long* ev = EVP(4); DS_EV(11); }
    // ---------------- phase F: y(s_in+3) = LN(s_in+2) + fc2(h)
    {
      const DsPhaseW w = ds_phase_w<D>(p, ge, l, 5);
      DsA a; a.src = p.h; a.ld = p.F; a.ln = 0; a.gamma = a.beta = nullptr;
      DsEpi e = {};
      e.kind = DS_EPI_Y; e.bias = L.b_fc2; e.y_prev = ybuf[(s_in + 2) & 1]; e.y_new = ybuf[(s_in + 3) & 1];
      ds_ln_params(p, s_in + 2, e.g_prev, e.b_prev);
      e.stats_new = p.stats + (s_in + 3) * stat_stride; e.NP = NP;
      ds_gemm<D, NTG, 4, 1, (D == 128 ? 2 : 4)>(p, ge, ring, smem, w, a, nullptr, NP, e, p.stats + (s_in + 2) * stat_stride, 0, p.rows, DsNoMid(), DsNoPost(), EVP(5));
      ring.cons_g += w.n_slots();   // released; refilled inside the next phase
    }
    { unsigned long long* ev = EVP(5); DS_EV(10); }
    bar.sync();
    { unsigned long long* ev = EVP(5); DS_EV(11); }
  }
  // final: x = LN(state 3L) -> fp32 + bf16 (operand of the LM head)
  {
    const int s = 3 * p.L;
    const float* y = (s & 1) ? p.y1 : p.y0;
    const float *g, *b;
    ds_ln_params(p, s, g, b);
    const float* st = p.stats + s * stat_stride;
    for (int row = gw; row < p.rows; row += gws) {
      float s1 = 0.f, s2 = 0.f;
      for (int i = lane; i < NP; i += 32) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(st + ((int64_t)row * NP + i) * 2));
        s1 += v.x; s2 += v.y;
      }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      const float mean = s1 * (1.f / D);
      const float rstd = rsqrtf(fmaxf(s2 * (1.f / D) - mean * mean, 0.f) + 1e-5f);
#pragma unroll
      for (int i = 0; i < D / 128; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 v = ds_ldcg_f4(y + (int64_t)row * D + c);
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
        float4 x;
        x.x = (v.x - mean) * rstd * gg.x + bb.x; x.y = (v.y - mean) * rstd * gg.y + bb.y;
        x.z = (v.z - mean) * rstd * gg.z + bb.z; x.w = (v.w - mean) * rstd * gg.w + bb.w;
        *reinterpret_cast<float4*>(p.x_f32 + (int64_t)row * D + c) = x;
        uint2 pk;
        pk.x = ds_pack(x.x, x.y); pk.y = ds_pack(x.z, x.w);
        *reinterpret_cast<uint2*>((bf16*)p.x_b16 + (int64_t)row * D + c) = pk;
      }
    }
  }
  // nothing may still be in flight into this CTA's shared memory when it exits, and no peer may still read it
  cluster_sync_all();
}

}  // namespace kmb

// ------------------------------------------------------------------ host side
template <int D>
static int ds_clusters(void) {
  using namespace kmb;
  using C = DsCfg<D>;
  static int cached = -1;
  if (cached >= 0) return cached;
  if (cudaFuncSetAttribute(decode_step_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) {
    kmb_set_last_error("kmb_decode_step_cluster: cannot reserve shared memory", __FILE__, __LINE__);
    return -1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(DS_MAX_CLUSTERS * DS_CL);
  cfg.blockDim = dim3(DS_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = DS_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, decode_step_kernel<D>, &cfg) != cudaSuccess || n < 1) {
    kmb_set_last_error("kmb_decode_step_cluster: cudaOccupancyMaxActiveClusters failed", __FILE__, __LINE__);
    (void)cudaGetLastError();
    return -1;
  }
  cached = n < DS_MAX_CLUSTERS ? n : DS_MAX_CLUSTERS;
  return cached;
}


template <int D>
static void ds_pack_offsets(int H, int F, int64_t* off7) {
  using C = kmb::DsCfg<D>;
  const int64_t swg = C::NTG * 8;
  const int64_t bytes[6] = {(int64_t)H * 8 * 4 * 24 * C::ROWB, (D / swg) * 4 * swg * C::ROWB, (int64_t)H * 4 * 4 * 16 * C::ROWB,
                            (D / swg) * 4 * swg * C::ROWB, (F / swg) * 4 * swg * C::ROWB, (D / swg) * 4 * (F / D) * swg * C::ROWB};
  off7[0] = 0;
  for (int i = 0; i < 6; ++i) off7[i + 1] = off7[i] + ((bytes[i] + 255) & ~(int64_t)255);
}

/* byte offsets of the six packed weight sections of one decoder layer (off7[6] = bytes per layer) */
extern "C" int kmb_decode_pack_offsets(int d, int H, int F, int64_t* off7) {
  if (!off7 || (d != 128 && d != 768 && d != 1024) || H * 64 != d || F % d) {
    kmb_set_last_error("kmb_decode_pack_offsets: unsupported shape", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if (d == 128) ds_pack_offsets<128>(H, F, off7);
  else if (d == 768) ds_pack_offsets<768>(H, F, off7);
  else ds_pack_offsets<1024>(H, F, off7);
  return KMB_OK;
}

/* packs the six weight matrices of one decoder layer (layer->w_*) into `out` (kmb_decode_pack_offsets()[6] bytes) */
extern "C" int kmb_decode_pack_weights(const KmbDecodeLayer* layer, int d, int H, int F, void* out, kmb_stream_t stream) {
  int64_t off[7];
  if (!layer || !out || kmb_decode_pack_offsets(d, H, F, off) != KMB_OK) return KMB_ERR_ARG;
  dim3 grid(256, 6);
  if (d == 128) kmb::decode_pack_kernel<128><<<grid, 256, 0, (cudaStream_t)stream>>>(*layer, H, F, (uint8_t*)out, off[0], off[1], off[2], off[3], off[4], off[5]);
  else if (d == 768) kmb::decode_pack_kernel<768><<<grid, 256, 0, (cudaStream_t)stream>>>(*layer, H, F, (uint8_t*)out, off[0], off[1], off[2], off[3], off[4], off[5]);
  else kmb::decode_pack_kernel<1024><<<grid, 256, 0, (cudaStream_t)stream>>>(*layer, H, F, (uint8_t*)out, off[0], off[1], off[2], off[3], off[4], off[5]);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_decode_cluster_grid(int d) {
  const int n = d == 128 ? ds_clusters<128>() : d == 1024 ? ds_clusters<1024>() : ds_clusters<768>();
  return n < 0 ? KMB_ERR_CUDA : n * kmb::DS_CL;
}

/* number of grid barriers of one step (size of the optional trace buffer) */
extern "C" int kmb_decode_cluster_barriers(int n_layers) { return 1 + 6 * n_layers; }

template <int D>
static int launch_step(const KmbDecodeStepC& p, cudaStream_t stream) {
  using namespace kmb;
  using C = DsCfg<D>;
  const int ncl = ds_clusters<D>();
  if (ncl < 1) return KMB_ERR_CUDA;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ncl * DS_CL));
  cfg.blockDim = dim3(DS_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = DS_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;   // co-residency of the whole grid is what makes the grid barriers legal
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  static int coop = 1;
  cfg.numAttrs = coop ? 2 : 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, decode_step_kernel<D>, p);
  if (err != cudaSuccess && coop) {
    // cooperative + cluster launch refused by this driver: the grid (<= 32 clusters, one CTA per SM, sized by
    // cudaOccupancyMaxActiveClusters) is still co-resident on an otherwise idle device; barriers are spin-bounded
    (void)cudaGetLastError();
    coop = 0;
    cfg.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg, decode_step_kernel<D>, p);
  }
  if (err != cudaSuccess) {
    kmb_set_last_error(cudaGetErrorString(err), __FILE__, __LINE__);
    (void)cudaGetLastError();
    return KMB_ERR_CUDA;
  }
  return KMB_OK;
}

extern "C" int kmb_decode_step_cluster(const KmbDecodeStepC* step, kmb_stream_t stream) {
  using namespace kmb;
  if (!step) { kmb_set_last_error("kmb_decode_step_cluster: null argument", __FILE__, __LINE__); return KMB_ERR_ARG; }
  const KmbDecodeStepC& p = *step;
  if ((p.d != 128 && p.d != 768 && p.d != 1024) || p.H * 64 != p.d || p.F % p.d || p.L < 1 || p.L > KMB_DECODE_MAX_LAYERS ||
      p.rows < 1 || p.t < 0 || p.t >= p.max_len || p.max_len > DS_MAXT || p.Se < 1 || p.Se > DS_MAXT || p.row_div < 1 ||
      !p.barrier || !p.layers[0].packed[0] || !p.ids || !p.y0 || !p.y1 || !p.stats || !p.x_f32 || !p.x_b16 || !p.ctx || !p.h) {
    kmb_set_last_error("kmb_decode_step_cluster: unsupported shape (d in {128, 768, 1024}, head_dim 64, max_len / S_e <= 512) or null buffer",
                       __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if (p.d == 128) return launch_step<128>(p, (cudaStream_t)stream);
  if (p.d == 768) return launch_step<768>(p, (cudaStream_t)stream);
  return launch_step<1024>(p, (cudaStream_t)stream);
}
