// Device-side decode controllers: top-k sampling and beam search bookkeeping, one launch (or two) per generated token,
// no host round trip inside the loop.
// replaces: the loop bodies of HF-3.0.2 generation_utils._generate_no_beam_search (do_sample branch: temperature,
//   top_k_top_p_filtering, softmax, multinomial) and _generate_beam_search (log_softmax, postprocess_next_token_scores,
//   topk(2 * num_beams) over beams x vocab, the per-candidate Python loop with BeamHypotheses.add / is_done, the beam
//   re-ordering of input_ids and — through _reorder_cache, src/model/mixins.py:419-434 — of every cached tensor), which the
//   reference reaches from src/model/mixins.py:336-382; adjust_logits_during_generation / _force_token_ids_generation
//   (src/model/mixins.py:400-417) are the `force_tok` argument.
//
// One CTA per row keeps the whole logits row (V fp32, 201 KB for V = 50 320) in shared memory: one global read, then
// every pass (max, sum-exp, radix select of the k-th largest value, ordered prefix scan for the multinomial draw) runs
// out of shared memory.  The k-th largest value comes from a 4 x 8-bit radix select on order-preserving keys with
// warp-aggregated histogram updates (logits share their leading bytes, plain shared-memory atomics would serialise).
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

constexpr int DC_THREADS = 1024;
constexpr int DC_WARPS = DC_THREADS / 32;

__device__ __forceinline__ uint32_t dc_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct DcScratch {
  float red[DC_WARPS];
  uint32_t hist[256];
  uint32_t bc[4];
  int ibox[4];
};

__device__ __forceinline__ float dc_block_max(float v, DcScratch& s) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s.red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s.red[threadIdx.x & 31];
  r = warp_max(r);
  return r;
}
// sum in a fixed order (lane tree, then warp 0 tree): bit-reproducible
__device__ __forceinline__ float dc_block_sum(float v, DcScratch& s) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s.red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s.red[threadIdx.x & 31];
  r = warp_sum(r);
  return r;
}

// Row of logits -> shared memory (fp32), returns the thread's running maximum.  16-byte loads, four per thread in flight:
// with one 4-byte load per thread per iteration the copy is latency bound (1024 x 4 B in flight per SM = 40 us for a
// 50 k vocabulary, profiles/r02_decode_analysis.md); `f(i, v)` is applied to every element before it is stored.
template <typename F>
__device__ __forceinline__ float dc_stage_row(const float* __restrict__ x, float* xs, int V, F f) {
  float mx = -INFINITY;
  const int V4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? (V >> 2) : 0;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int i = threadIdx.x; i < V4; i += 4 * DC_THREADS) {
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * DC_THREADS < V4) a[u] = __ldg(x4 + i + u * DC_THREADS);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = i + u * DC_THREADS;
      if (j < V4) {
        float4 v = a[u];
        v.x = f(4 * j, v.x); v.y = f(4 * j + 1, v.y); v.z = f(4 * j + 2, v.z); v.w = f(4 * j + 3, v.w);
        reinterpret_cast<float4*>(xs)[j] = v;
        mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
      }
    }
  }
  for (int i = 4 * V4 + threadIdx.x; i < V; i += DC_THREADS) {
    const float v = f(i, x[i]);
    xs[i] = v;
    mx = fmaxf(mx, v);
  }
  return mx;
}

// key of the k-th largest element of x[0, V) (1 <= k <= V)
__device__ uint32_t dc_radix_kth(const float* x, int V, int k, DcScratch& s) {
  const int lane = threadIdx.x & 31;
  uint32_t prefix = 0, mask = 0;
  int krem = k;
  const int Vp = (V + 31) & ~31;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += DC_THREADS) s.hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < Vp; i += DC_THREADS) {
      const uint32_t key = i < V ? dc_key(x[i]) : 0u;
      const bool in = i < V && (key & mask) == prefix;
      const unsigned act = __ballot_sync(0xffffffffu, in);
      if (in) {
        const uint32_t bin = (key >> shift) & 255u;
        const unsigned peers = __match_any_sync(act, bin);
        if (lane == __ffs(peers) - 1) atomicAdd(&s.hist[bin], (uint32_t)__popc(peers));
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // warp 0: bins from the top, 8 per lane
      uint32_t c[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = s.hist[255 - (lane * 8 + j)]; tot += c[j]; }
      uint32_t incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      uint32_t cum = incl - tot;
      if (cum < (uint32_t)krem && (uint32_t)krem <= incl) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (cum + c[j] >= (uint32_t)krem) { s.bc[0] = 255 - (lane * 8 + j); s.bc[1] = krem - cum; break; }
          cum += c[j];
        }
      }
    }
    __syncthreads();
    prefix |= s.bc[0] << shift;
    mask |= 255u << shift;
    krem = (int)s.bc[1];
  }
  return prefix;
}

// Nucleus threshold on the (already top-k filtered, exponentiated) row e[0, V): HF-3.0.2 top_k_top_p_filtering keeps a token
// iff the probability mass of the tokens sorted strictly before it is <= top_p (the first token always survives).  With
// G(x) = mass of entries > x that is "keep iff G(e_i) <= top_p * total": returns the key of the smallest kept value, found by
// the same 4 x 8-bit radix descent with per-bin MASS histograms (e >= 0, so the float bit pattern orders like the value).
__device__ uint32_t dc_radix_mass(const float* e, int V, float limit, DcScratch& s) {
  const int lane = threadIdx.x & 31;
  float* fh = reinterpret_cast<float*>(s.hist);
  uint32_t prefix = 0, mask = 0;
  float above = 0.f;                         // mass of the bins already known to lie above the threshold
  const int Vp = (V + 31) & ~31;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += DC_THREADS) fh[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < Vp; i += DC_THREADS) {
      const float v = i < V ? e[i] : 0.f;
      const uint32_t key = __float_as_uint(v);
      const bool in = i < V && v > 0.f && (key & mask) == prefix;
      const unsigned act = __ballot_sync(0xffffffffu, in);
      if (in) {
        const uint32_t bin = (key >> shift) & 255u;
        const unsigned peers = __match_any_sync(act, bin);
        float acc = 0.f;                     // mass of the lanes that share this bin, summed in lane order
        for (unsigned q = peers; q; q &= q - 1) acc += __shfl_sync(peers, v, __ffs(q) - 1);
        if (lane == __ffs(peers) - 1) atomicAdd(&fh[bin], acc);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // bins from the top, 8 per lane: first bin whose inclusive mass exceeds the limit
      float c[8], tot = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = fh[255 - (lane * 8 + j)]; tot += c[j]; }
      float incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      float cum = above + incl - tot;
      const bool mine = cum <= limit && above + incl > limit;
      const unsigned who = __ballot_sync(0xffffffffu, mine);
      const float all_bins = __shfl_sync(0xffffffffu, incl, 31);   // by every lane: a shuffle under `lane == 0` never returns
      if (who == 0) {                         // rounding: everything fits under the limit -> lowest populated bin
        if (lane == 0) { s.bc[0] = 0; s.bc[2] = __float_as_uint(above + all_bins); }
      } else if (lane == __ffs(who) - 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (cum + c[j] > limit) { s.bc[0] = 255 - (lane * 8 + j); s.bc[2] = __float_as_uint(cum); break; }
          cum += c[j];
        }
      }
    }
    __syncthreads();
    prefix |= s.bc[0] << shift;
    mask |= 255u << shift;
    above = __uint_as_float(s.bc[2]);
    __syncthreads();
  }
  return prefix;
}

// k-th largest key of a staged row for SMALL k (beam search: 2 * num_beams; top-k sampling: 50) without the four
// histogram passes over the whole vocabulary: the k-th largest of the 1024 per-thread maxima is a lower bound t0 of the
// answer (there are at least k elements >= t0), one pass collects the few elements >= t0, and the exact k-th largest is
// selected among those.  Returns false (caller runs dc_radix_kth over the row) when k is large or the candidates overflow
// (heavy ties, e.g. a forced-token step where all but one logit are -inf).
constexpr int DC_CAND_CAP = 1024;
struct DcCand {
  float v[DC_CAND_CAP];
  float tmax[DC_THREADS];
  int n;
};
// `slack`: how many of the elements behind the thread maxima were removed from xs after the maxima were taken.
__device__ bool dc_kth_filtered(const float* xs, int V, int k, int slack, float tmax, DcScratch& s, DcCand& cd, uint32_t* thr) {
  if (k + slack > DC_THREADS / 8 || V < DC_THREADS) return false;
  cd.tmax[threadIdx.x] = tmax;
  if (threadIdx.x == 0) cd.n = 0;
  __syncthreads();
  const uint32_t t0 = dc_radix_kth(cd.tmax, DC_THREADS, k + slack, s);
  for (int i = threadIdx.x; i < V; i += DC_THREADS) {
    const float v = xs[i];
    if (dc_key(v) >= t0) {
      const int slot = atomicAdd(&cd.n, 1);
      if (slot < DC_CAND_CAP) cd.v[slot] = v;
    }
  }
  __syncthreads();
  const int n = cd.n;
  if (n > DC_CAND_CAP) return false;
  *thr = dc_radix_kth(cd.v, n, k, s);
  return true;
}

// ------------------------------------------------------------------ top-k / top-p sampling (num_beams == 1, do_sample)
__global__ void __launch_bounds__(DC_THREADS) sample_select_kernel(const float* logits, int64_t ld, int V, float inv_temp, int top_k, float top_p, int eos, int pad,
                                                                  int ban_eos, int cur_len, const unsigned long long* seed, int64_t* unfinished,
                                                                  int64_t* sent_len, int64_t* out, int64_t out_ld, int64_t* ids_next) {
  extern __shared__ __align__(16) uint8_t dc_smem[];
  float* xs = reinterpret_cast<float*>(dc_smem);
  DcScratch& sc = *reinterpret_cast<DcScratch*>(dc_smem + (((size_t)V * 4 + 15) & ~(size_t)15));
  const int row = blockIdx.x;
  const float* x = logits + (int64_t)row * ld;
  __shared__ DcCand cd;
  const float tmax = dc_stage_row(x, xs, V, [&](int i, float v) { return ((ban_eos && i == eos) ? -INFINITY : v) * inv_temp; });
  const float mx = dc_block_max(tmax, sc);
  __syncthreads();
  // HF-3.0.2 top_k_top_p_filtering: top_k > 0 removes every logit below the k-th largest VALUE (ties survive); 0 = no filter
  uint32_t thr = 0;
  if (top_k > 0 && top_k < V && !dc_kth_filtered(xs, V, top_k, 0, tmax, sc, cd, &thr)) {
    __syncthreads();
    thr = dc_radix_kth(xs, V, top_k, sc);
  }
  // probabilities in place, contiguous chunks per thread so that the scan below walks the vocabulary in index order
  const int chunk = (V + DC_THREADS - 1) / DC_THREADS;
  const int i0 = threadIdx.x * chunk, i1 = min(V, i0 + chunk);
  float loc = 0.f;
  int last_kept = -1;
  for (int i = i0; i < i1; ++i) {
    const float v = xs[i];
    const bool keep = v > -INFINITY && dc_key(v) >= thr;
    const float e = keep ? __expf(v - mx) : 0.f;
    xs[i] = e;
    loc += e;
    if (keep) last_kept = i;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // Inverse-CDF draw over the surviving probabilities in index order (exclusive scan of the per-thread chunk sums, fixed
  // order).  Nucleus (top_p < 1, HF-3.0.2 top_k_top_p_filtering after the top-k filter: a token survives iff the mass of
  // the strictly more probable tokens is <= top_p): sampling from the renormalised nucleus is sampling from the whole
  // distribution conditioned on landing inside it, so the kernel draws, measures the mass above the drawn token with one
  // pass, and redraws (next counter value) when it fell outside — acceptance probability >= top_p, no sort and no
  // histogram.  After DC_NUCLEUS_TRIES rejections the exact threshold (dc_radix_mass) filters the row and one more draw decides.
  constexpr int DC_NUCLEUS_TRIES = 24;
  const bool nucleus = top_p < 1.0f;
  for (int attempt = 0;; ++attempt) {
    const bool exact = nucleus && attempt == DC_NUCLEUS_TRIES;
    if (exact) {
      const float tot0 = dc_block_sum(loc, sc);
      __syncthreads();
      const uint32_t pthr = dc_radix_mass(xs, V, top_p * tot0, sc);
      loc = 0.f;
      last_kept = -1;
      for (int i = i0; i < i1; ++i) {
        const float e = xs[i];
        const bool keep = e > 0.f && __float_as_uint(e) >= pthr;
        if (!keep) xs[i] = 0.f;
        else { loc += e; last_kept = i; }
      }
    }
    float incl = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) sc.red[wid] = incl;
    if (threadIdx.x == 0) { sc.ibox[0] = -1; sc.ibox[2] = 0; }
    __syncthreads();
    float wbase = 0.f, total = 0.f;
    for (int w = 0; w < DC_WARPS; ++w) {
      const float t = sc.red[w];
      if (w < wid) wbase += t;
      total += t;
    }
    atomicMax(&sc.ibox[0], last_kept);   // fallback: the last surviving token (rounding can push the target past the total)
    const float excl = wbase + incl - loc;
    // one uniform draw per (seed, row, step, attempt)
    const unsigned long long bits = mix64(seed[0] ^ mix64(((unsigned long long)row << 20) + (unsigned long long)cur_len + 0x9E3779B97F4A7C15ULL +
                                                          (unsigned long long)attempt * 0xD1B54A32D192ED03ULL));
    const float u = (float)(bits >> 40) * (1.0f / 16777216.0f);
    const float target = u * total;
    __syncthreads();
    if (loc > 0.f && excl <= target && target < excl + loc) {
      float run = excl;
      int pick = last_kept;
      for (int i = i0; i < i1; ++i) {
        run += xs[i];
        if (xs[i] > 0.f && target < run) { pick = i; break; }
      }
      sc.ibox[1] = pick;
      sc.ibox[2] = 1;
    }
    __syncthreads();
    if (!nucleus || exact) break;
    // mass of the tokens strictly more probable than the drawn one
    const int tokc = (sc.ibox[2] == 1) ? sc.ibox[1] : sc.ibox[0];
    const float ep = tokc >= 0 ? xs[tokc] : 0.f;
    float above = 0.f;
    for (int i = i0; i < i1; ++i) {
      const float e = xs[i];
      above += e > ep ? e : 0.f;
    }
    above = dc_block_sum(above, sc);
    if (above <= top_p * total) break;       // block-uniform: inside the nucleus
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tokv = (sc.ibox[2] == 1) ? sc.ibox[1] : sc.ibox[0];
    if (tokv < 0) tokv = 0;
    int64_t tok = tokv;
    if (eos >= 0) {
      const int64_t uf = unfinished[row];
      tok = tok * uf + (int64_t)pad * (1 - uf);
      if (uf && tok == eos) {
        sent_len[row] = cur_len + 1;
        unfinished[row] = 0;
      }
    }
    out[(int64_t)row * out_ld + cur_len] = tok;
    ids_next[row] = tok;
  }
}

// ------------------------------------------------------------------ beam search, step 1 of 2: per-row candidates
// scores = log_softmax(force(logits)); EOS banned below min_length; the K = 2 * num_beams best (score + beam_score, token)
// of the row.  The per-sample top-K over beams x vocab is a subset of the union of the per-row top-K lists.
__global__ void __launch_bounds__(DC_THREADS) beam_topk_kernel(const float* logits, int64_t ld, int V, int K, int eos, int force_tok, int ban_eos,
                                                              const float* beam_scores, float* cand_val, int* cand_tok) {
  extern __shared__ __align__(16) uint8_t dc_smem[];
  float* xs = reinterpret_cast<float*>(dc_smem);
  DcScratch& sc = *reinterpret_cast<DcScratch*>(dc_smem + (((size_t)V * 4 + 15) & ~(size_t)15));
  const int row = blockIdx.x;
  const float* x = logits + (int64_t)row * ld;
  if (force_tok >= 0 && force_tok < V) {
    // Forced step (HF-3.0.2 adjust_logits_during_generation: BOS at the first step, EOS at max_length - 1): every other
    // logit is -inf, so log_softmax gives the forced token 0 (when its logit is finite) and -inf elsewhere; the K best are the
    // forced token followed by the lowest token indices (ties at -inf in index order).  No pass over the vocabulary.
    if (threadIdx.x == 0) {
      const float xf = x[force_tok];
      const bool dead = (ban_eos && force_tok == eos) || !(xf > -INFINITY && xf < INFINITY);
      const float bsc = beam_scores[row];
      int out = 0;
      if (!dead) {
        cand_val[(int64_t)row * K] = ((xf - xf) - 0.f) + bsc;
        cand_tok[(int64_t)row * K] = force_tok;
        out = 1;
      }
      for (int i = 0; out < K && i < V; ++i) {
        if (!dead && i == force_tok) continue;
        cand_val[(int64_t)row * K + out] = -INFINITY;
        cand_tok[(int64_t)row * K + out] = i;
        ++out;
      }
    }
    return;
  }
  __shared__ DcCand cd;
  float tmax = dc_stage_row(x, xs, V, [&](int i, float v) { return v; });
  const float mx = dc_block_max(tmax, sc);
  float se = 0.f;
  for (int i = threadIdx.x; i < V; i += DC_THREADS) se += __expf(xs[i] - mx);
  se = dc_block_sum(se, sc);
  const float lse = logf(se);
  __syncthreads();
  const bool banned = ban_eos && eos >= 0 && eos < V;
  if (banned && threadIdx.x == 0) xs[eos] = -INFINITY;
  __syncthreads();
  // Fast path for K + 1 <= 32 (beam search at the reference's num_beams = 5 needs K = 10): the (K + slack)-th largest of
  // the 32 WARP maxima bounds the K-th largest logit from below, one pass collects the ~K..2K logits above the bound
  // with their indices, and each collected logit finds its rank by counting — no histogram passes, no second scan of the
  // row for the output.  The maxima were taken before the ban, hence slack = 1 (the log-sum-exp above includes the banned
  // logit, like HF-3.0.2).  Heavy ties (forced-token steps) overflow the list and take the general path below.
  {
    __shared__ int cidx[DC_CAND_CAP];
    __shared__ float wmax[DC_WARPS];
    __shared__ uint32_t t0s;
    const int lane_ = threadIdx.x & 31, wid_ = threadIdx.x >> 5;
    const int kk = K + (banned ? 1 : 0);
    const float wm = warp_max(tmax);
    if (lane_ == 0) wmax[wid_] = wm;
    if (threadIdx.x == 0) cd.n = 0;
    __syncthreads();
    if (kk <= DC_WARPS && V >= DC_THREADS) {
      if (wid_ == 0) {
        const float mine = wmax[lane_];
        int r = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float o = __shfl_sync(0xffffffffu, mine, j);
          r += (o > mine) || (o == mine && j < lane_);
        }
        if (r == kk - 1) t0s = dc_key(mine);
      }
      __syncthreads();
      const uint32_t t0 = t0s;
      for (int i = threadIdx.x; i < V; i += DC_THREADS) {
        const float v = xs[i];
        if (dc_key(v) >= t0) {
          const int slot = atomicAdd(&cd.n, 1);
          if (slot < DC_CAND_CAP) { cd.v[slot] = v; cidx[slot] = i; }
        }
      }
      __syncthreads();
      const int n = cd.n;
      if (n <= DC_CAND_CAP) {                     // n >= K by construction
        const float bsc_ = beam_scores[row];
        for (int c = threadIdx.x; c < n; c += DC_THREADS) {
          const float v = cd.v[c];
          const int i = cidx[c];
          int r = 0;
          for (int j = 0; j < n; ++j) {
            const float o = cd.v[j];
            r += (o > v) || (o == v && cidx[j] < i);
          }
          if (r < K) {
            cand_val[(int64_t)row * K + r] = ((v - mx) - lse) + bsc_;
            cand_tok[(int64_t)row * K + r] = i;
          }
        }
        return;
      }
      __syncthreads();
    }
  }
  uint32_t thr = 0;
  if (!dc_kth_filtered(xs, V, K, banned ? 1 : 0, tmax, sc, cd, &thr)) {
    __syncthreads();
    thr = dc_radix_kth(xs, V, K, sc);
  }
  // strictly greater first (any order), then ties with the threshold in index order until K candidates are out
  if (threadIdx.x == 0) { sc.ibox[0] = 0; sc.ibox[1] = 0; }
  __syncthreads();
  const float bsc = beam_scores[row];
  const int chunk = (V + DC_THREADS - 1) / DC_THREADS;
  const int i0 = threadIdx.x * chunk, i1 = min(V, i0 + chunk);
  int n_eq = 0;
  for (int i = i0; i < i1; ++i) {
    const uint32_t key = dc_key(xs[i]);
    if (key > thr) {
      const int slot = atomicAdd(&sc.ibox[0], 1);
      cand_val[(int64_t)row * K + slot] = ((xs[i] - mx) - lse) + bsc;
      cand_tok[(int64_t)row * K + slot] = i;
    } else if (key == thr) {
      ++n_eq;
    }
  }
  // exclusive scan of the tie counts (index order)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = n_eq;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __shared__ int wtot[DC_WARPS];
  if (lane == 31) wtot[wid] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < wid; ++w) base += wtot[w];
  int rank = base + incl - n_eq;
  const int n_gt = sc.ibox[0];
  const int need = K - n_gt;
  if (n_eq > 0 && rank < need) {
    for (int i = i0; i < i1 && rank < need; ++i) {
      if (dc_key(xs[i]) == thr) {
        cand_val[(int64_t)row * K + n_gt + rank] = ((xs[i] - mx) - lse) + bsc;
        cand_tok[(int64_t)row * K + n_gt + rank] = i;
        ++rank;
      }
    }
  }
}

// ------------------------------------------------------------------ beam search, step 2 of 2: per-sample bookkeeping
// HF-3.0.2 _generate_beam_search loop body after the topk, one CTA per batch element.
__global__ void __launch_bounds__(256) beam_update_kernel(const KmbBeamState st, int cur_len) {
  extern __shared__ __align__(16) uint8_t bu_smem[];
  const int b = blockIdx.x, nb = st.num_beams, K = st.K, n = nb * K;
  float* cv = reinterpret_cast<float*>(bu_smem);              // [n] candidate scores
  long long* cf = reinterpret_cast<long long*>(cv + ((n + 1) & ~1));   // [n] flat index beam * V + token
  int* order = reinterpret_cast<int*>(cf + n);                // [K] candidates by rank
  int* nxt_beam = order + K;                                  // [nb]
  int* nxt_tok = nxt_beam + nb;                               // [nb]
  float* nxt_sc = reinterpret_cast<float*>(nxt_tok + nb);     // [nb]
  int* tmp = reinterpret_cast<int*>(nxt_sc + nb);             // [nb * max_len] history / ancestry staging
  const int row0 = b * nb;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int beam = i / K;
    cv[i] = st.cand_val[(int64_t)(row0 + beam) * K + (i % K)];
    cf[i] = (long long)beam * st.V + st.cand_tok[(int64_t)(row0 + beam) * K + (i % K)];
  }
  __syncthreads();
  // rank by counting: larger score first, lower flat index first among equals (sorted=True order of torch.topk)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = cv[i];
    const long long f = cf[i];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (cv[j] > v) || (cv[j] == v && cf[j] < f);
    if (r < K) order[r] = i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int done = st.done[b];
    if (done) {
      for (int i = 0; i < nb; ++i) { nxt_sc[i] = 0.f; nxt_tok[i] = st.pad; nxt_beam[i] = 0; }
    } else {
      int cnt = 0;
      int hn = st.hyp_n[b];
      double worst = st.worst[b];
      double* hs = st.hyp_score + (int64_t)b * (nb + 1);
      int* hl = st.hyp_len + (int64_t)b * (nb + 1);
      int* ht = st.hyp_tok + (int64_t)b * (nb + 1) * st.max_len;
      const double lenpow = pow((double)cur_len, st.length_penalty);
      for (int r = 0; r < K && cnt < nb; ++r) {
        const int i = order[r];
        const int beam = (int)(cf[i] / st.V), tok = (int)(cf[i] % st.V);
        if (st.eos >= 0 && tok == st.eos) {
          if (r >= nb) continue;
          // BeamHypotheses.add(input_ids[effective_beam_id].clone(), score)
          const double score = (double)cv[i] / lenpow;
          if (hn < nb || score > worst) {
            hs[hn] = score; hl[hn] = cur_len;
            const int* src = st.hist + (int64_t)(row0 + beam) * st.max_len;
            for (int q = 0; q < cur_len; ++q) ht[(int64_t)hn * st.max_len + q] = src[q];
            ++hn;
            if (hn > nb) {
              // drop the worst (lowest score, first inserted among equals); worst_score = the next lowest
              int w0 = 0;
              for (int q = 1; q < hn; ++q) if (hs[q] < hs[w0]) w0 = q;
              for (int q = w0; q + 1 < hn; ++q) {
                hs[q] = hs[q + 1]; hl[q] = hl[q + 1];
                for (int z = 0; z < st.max_len; ++z) ht[(int64_t)q * st.max_len + z] = ht[(int64_t)(q + 1) * st.max_len + z];
              }
              --hn;
              double w1 = hs[0];
              for (int q = 1; q < hn; ++q) w1 = fmin(w1, hs[q]);
              worst = w1;
            } else {
              worst = fmin(score, worst);
            }
          }
        } else {
          nxt_sc[cnt] = cv[i]; nxt_tok[cnt] = tok; nxt_beam[cnt] = beam;
          ++cnt;
        }
      }
      // is_done(best_sum_logprobs = max of the sample's candidate scores, cur_len)
      if (hn >= nb) {
        if (st.early_stopping) done = 1;
        else done = worst >= (double)cv[order[0]] / lenpow;
      }
      st.hyp_n[b] = hn;
      st.worst[b] = worst;
      if (done) { st.done[b] = 1; atomicAdd(st.done_count, 1); }
      for (; cnt < nb; ++cnt) { nxt_sc[cnt] = -1e9f; nxt_tok[cnt] = st.pad; nxt_beam[cnt] = 0; }   // unreachable (HF asserts a full beam)
    }
  }
  __syncthreads();
  // new beam i continues old beam nxt_beam[i]: token history, ancestry table, scores, next input token
  const int L = cur_len;
  for (int i = threadIdx.x; i < nb * L; i += blockDim.x) tmp[i] = st.hist[(int64_t)(row0 + nxt_beam[i / L]) * st.max_len + (i % L)];
  __syncthreads();
  for (int i = threadIdx.x; i < nb * L; i += blockDim.x) st.hist[(int64_t)(row0 + i / L) * st.max_len + (i % L)] = tmp[i];
  __syncthreads();
  if (st.slot_tbl) {
    for (int i = threadIdx.x; i < nb * L; i += blockDim.x) tmp[i] = st.slot_tbl[(int64_t)(row0 + nxt_beam[i / L]) * st.max_len + (i % L)];
    __syncthreads();
    for (int i = threadIdx.x; i < nb * L; i += blockDim.x) st.slot_tbl[(int64_t)(row0 + i / L) * st.max_len + (i % L)] = tmp[i];
  }
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    const int row = row0 + i;
    if (L < st.max_len) st.hist[(int64_t)row * st.max_len + L] = nxt_tok[i];
    st.beam_scores[row] = nxt_sc[i];
    st.ids_next[row] = nxt_tok[i];
    st.beam_idx[row] = row0 + nxt_beam[i];
  }
}

}  // namespace kmb

static const int DC_MAX_DYN_SMEM = 227 * 1024 - 14 * 1024;   // static shared memory of the kernels comes out of the same 227 KB
static int dc_row_smem(int V) { return (int)((((size_t)V * 4 + 15) & ~(size_t)15) + sizeof(kmb::DcScratch) + 64); }

extern "C" int kmb_select_max_vocab(void) { return (DC_MAX_DYN_SMEM - (int)sizeof(kmb::DcScratch) - 128) / 4; }

extern "C" int kmb_sample_select(const float* logits, int64_t ld, int rows, int V, float temperature, int top_k, float top_p, int eos_token_id,
                                 int pad_token_id, int ban_eos, int cur_len, const uint64_t* seed, int64_t* unfinished, int64_t* sent_len,
                                 int64_t* out_tokens, int64_t out_ld, int64_t* ids_next, kmb_stream_t stream) {
  if (!logits || rows <= 0 || V <= 0 || V > kmb_select_max_vocab() || !(temperature > 0.f) || top_k < 0 || !(top_p >= 0.f) || !seed || !unfinished || !sent_len ||
      !out_tokens || !ids_next || cur_len < 0 || cur_len >= out_ld) {
    kmb_set_last_error("kmb_sample_select: bad argument (vocabulary must fit one CTA's shared memory)", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kmb::sample_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DC_MAX_DYN_SMEM) != cudaSuccess) {
      kmb_set_last_error("kmb_sample_select: cannot reserve shared memory", __FILE__, __LINE__);
      return KMB_ERR_CUDA;
    }
    attr = true;
  }
  kmb::sample_select_kernel<<<rows, kmb::DC_THREADS, dc_row_smem(V), (cudaStream_t)stream>>>(
      logits, ld, V, 1.0f / temperature, top_k, top_p, eos_token_id, pad_token_id, ban_eos, cur_len, (const unsigned long long*)seed, unfinished, sent_len,
      out_tokens, out_ld, ids_next);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_beam_step(const float* logits, int64_t ld, const KmbBeamState* state, int cur_len, int force_token, int ban_eos,
                             kmb_stream_t stream) {
  if (!logits || !state) { kmb_set_last_error("kmb_beam_step: null argument", __FILE__, __LINE__); return KMB_ERR_ARG; }
  const KmbBeamState& st = *state;
  const int rows = st.batch * st.num_beams;
  if (st.batch <= 0 || st.num_beams < 1 || st.K != 2 * st.num_beams || st.K > st.V || st.V > kmb_select_max_vocab() || cur_len < 1 ||
      cur_len >= st.max_len || !st.cand_val || !st.cand_tok || !st.beam_scores || !st.hist || !st.ids_next || !st.beam_idx || !st.done || !st.hyp_n ||
      !st.hyp_score || !st.hyp_len || !st.hyp_tok || !st.worst || !st.done_count) {
    kmb_set_last_error("kmb_beam_step: bad state", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const size_t usm = (size_t)((st.num_beams * st.K + 1) & ~1) * 4 + (size_t)st.num_beams * st.K * 8 + (size_t)st.K * 4 + (size_t)st.num_beams * 12 +
                     (size_t)st.num_beams * st.max_len * 4 + 64;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kmb::beam_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DC_MAX_DYN_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(kmb::beam_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      kmb_set_last_error("kmb_beam_step: cannot reserve shared memory", __FILE__, __LINE__);
      (void)cudaGetLastError();
      return KMB_ERR_CUDA;
    }
    attr = true;
  }
  if (usm > 200 * 1024) { kmb_set_last_error("kmb_beam_step: num_beams * max_len too large", __FILE__, __LINE__); return KMB_ERR_ARG; }
  kmb::beam_topk_kernel<<<rows, kmb::DC_THREADS, dc_row_smem(st.V), (cudaStream_t)stream>>>(logits, ld, st.V, st.K, st.eos, force_token, ban_eos,
                                                                                          st.beam_scores, st.cand_val, st.cand_tok);
  KMB_CHECK_LAUNCH();
  kmb::beam_update_kernel<<<st.batch, 256, usm, (cudaStream_t)stream>>>(st, cur_len);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}
