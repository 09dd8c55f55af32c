// Device-side token selection for the greedy decode loop.
// replaces: the body of HF-3.0.2 _generate_no_beam_search that the reference reaches from
//   src/model/mixins.py:368-382 (postprocess_next_token_scores min_length EOS ban, argmax, pad for finished
//   rows, append to input_ids, sent_lengths / unfinished_sents bookkeeping) — ~15 tiny torch kernels and, in the
//   reference, a host sync per step (`unfinished_sents.max() == 0`).
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

// one CTA per row: argmax over V fp32 logits (first index on ties, like torch.argmax), then the bookkeeping
__global__ void __launch_bounds__(1024) greedy_select_kernel(const float* logits, int64_t ld, int V, int eos, int pad, int ban_eos,
                                                            int cur_len, int64_t* unfinished, int64_t* sent_len, int64_t* out,
                                                            int64_t out_ld, int64_t* ids_next) {
  const int row = blockIdx.x;
  const float* x = logits + (int64_t)row * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  auto consume = [&](const float (&v)[4], int i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float val = v[j];
      if (ban_eos && i + j == eos) val = -INFINITY;
      if (i + j < V && (val > best || (val == best && i + j < bi))) { best = val; bi = i + j; }
    }
  };
  const int stride = blockDim.x * 4;
  int i = threadIdx.x * 4;
  if ((((uintptr_t)x) & 15) == 0) {
    // four independent 16-byte loads in flight per thread (the row is a 200 KB latency-bound stream for one CTA)
    for (; i + 3 * stride + 4 <= V; i += 4 * stride) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = *reinterpret_cast<const float4*>(x + i + u * stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float v[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
        consume(v, i + u * stride);
      }
    }
  }
  for (; i < V; i += stride) {
    float v[4];
    if (i + 4 <= V && ((((uintptr_t)(x + i)) & 15) == 0)) {
      const float4 t = *reinterpret_cast<const float4*>(x + i);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (i + j < V) ? x[i + j] : -INFINITY;
    }
    consume(v, i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  __shared__ float sb[32];
  __shared__ int si[32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (lane == 0) { sb[wib] = best; si[wib] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (sb[w] > best || (sb[w] == best && si[w] < bi)) { best = sb[w]; bi = si[w]; }
    if (bi == 0x7fffffff) bi = 0;   // all -inf / NaN row: torch.argmax returns 0
    int64_t tok = bi;
    if (eos >= 0) {
      const int64_t u = unfinished[row];
      tok = tok * u + (int64_t)pad * (1 - u);
      if (u && tok == eos) {
        sent_len[row] = cur_len + 1;
        unfinished[row] = 0;
      }
    }
    out[(int64_t)row * out_ld + cur_len] = tok;
    ids_next[row] = tok;
  }
}

}  // namespace kmb

extern "C" int kmb_greedy_select(const float* logits, int64_t ld, int rows, int V, int eos_token_id, int pad_token_id,
                                 int ban_eos, int cur_len, int64_t* unfinished, int64_t* sent_len, int64_t* out_tokens,
                                 int64_t out_ld, int64_t* ids_next, kmb_stream_t stream) {
  if (!logits || rows <= 0 || V <= 0 || !unfinished || !sent_len || !out_tokens || !ids_next || cur_len < 0 || cur_len >= out_ld) {
    kmb_set_last_error("kmb_greedy_select: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  kmb::greedy_select_kernel<<<rows, 1024, 0, (cudaStream_t)stream>>>(logits, ld, V, eos_token_id, pad_token_id, ban_eos, cur_len,
                                                                    unfinished, sent_len, out_tokens, out_ld, ids_next);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}
