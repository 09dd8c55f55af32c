// HBM-bound kernels around the GEMMs: RoI-feature packing, the multimodal embedding
// (token gather | visual row + box projection) + position + LayerNorm (+dropout),
// post-residual LayerNorm forward/backward with fused dropout re-masking and the
// column reductions that produce LayerNorm / bias gradients, and the embedding backward.
//
// Reference call sites replaced:
//   ImageEmbedding.forward cat + Linear ........ src/model/modules.py:24-41
//   _embed_multi_modal gather / overwrite ...... src/model/modules.py:89-102
//   + embed_positions + layernorm_embedding .... src/model/modules.py:133-137
//   decoder embed + pos + LN ................... HF-3.0.2 BartDecoder.forward (src/model/model.py:87-97)
//   residual + LayerNorm (post-LN) ............. HF-3.0.2 EncoderLayer / DecoderLayer
// One warp owns one token row; a row (d <= 1024 fp32) stays in registers between the
// statistics pass and the normalise pass, so each element is read once and written once.
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

constexpr int FEAT = 2048;  // RoI feature width; the 4 box columns follow (src/data/dataset.py:44-47)
constexpr int MAXV = 8;     // float4 per lane: d <= 1024

struct DropCfg {
  uint32_t thresh16;  // 0 = off
  float scale;
  uint32_t tag;
  const uint64_t* seed;
};

__device__ __forceinline__ float4 drop4(float4 v, uint64_t seed, const DropCfg& dc, uint64_t q) {
  const uint64_t bits = dropout_bits4(seed, dc.tag, q);
  v.x = dropout_keep(bits, 0, dc.thresh16) ? v.x * dc.scale : 0.f;
  v.y = dropout_keep(bits, 1, dc.thresh16) ? v.y * dc.scale : 0.f;
  v.z = dropout_keep(bits, 2, dc.thresh16) ? v.z * dc.scale : 0.f;
  v.w = dropout_keep(bits, 3, dc.thresh16) ? v.w * dc.scale : 0.f;
  return v;
}

__device__ __forceinline__ uint2 pack4(float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// ------------------------------------------------------------------ RoI feature packing
// list-of-tensors [n_i, 2052] fp32 (pointer table) or one packed [R, 2052] fp32  ->
// feats bf16 [R, 2048] (tensor-core operand) + boxes fp32 [R, 4] (kept exact: raw pixels)
__global__ void __launch_bounds__(256) pack_features_kernel(const float* const* ptrs, const int* row_off, int B,
                                                            const float* packed, bf16* feats, float* boxes, int R) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  if (r >= R) return;
  if (row_off && r >= row_off[B]) {
    // R is the CAPACITY of the packed buffers; rows past the batch's region count stay zero (GEMM rows / K-reductions /
    // column sums over them contribute nothing)
    *reinterpret_cast<uint4*>(feats + (int64_t)r * FEAT + threadIdx.x * 8) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) *reinterpret_cast<float4*>(boxes + (int64_t)r * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float* src;
  if (ptrs) {
    int lo = 0, hi = B;  // largest b with row_off[b] <= r
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (row_off[mid] <= r) lo = mid; else hi = mid;
    }
    src = ptrs[lo] + (int64_t)(r - row_off[lo]) * (FEAT + 4);
  } else {
    src = packed + (int64_t)r * (FEAT + 4);
  }
  const int c = threadIdx.x * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(src + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(src + c + 4));
  const uint2 pa = pack4(a), pb = pack4(b);
  *reinterpret_cast<uint4*>(feats + (int64_t)r * FEAT + c) = make_uint4(pa.x, pa.y, pb.x, pb.y);
  if (threadIdx.x == 0)
    *reinterpret_cast<float4*>(boxes + (int64_t)r * 4) = __ldg(reinterpret_cast<const float4*>(src + FEAT));
}

// slot_idx[b, s] = packed RoI row that overwrites token (b, s), or -1 (src/model/modules.py:91,98-100)
__global__ void slot_index_kernel(const int64_t* ids, const int* row_off, int S, int img_feat_id, int cls_id,
                                  int* slot_idx) {
  const int b = blockIdx.x;
  __shared__ int warp_cnt[32];
  __shared__ int running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int s0 = 0; s0 < S; s0 += blockDim.x) {
    const int s = s0 + threadIdx.x;
    int flag = 0;
    if (s < S) {
      const int64_t t = ids[(int64_t)b * S + s];
      flag = (t == img_feat_id || t == cls_id);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[w] = __popc(bal);
    __syncthreads();
    int before = running;
    for (int i = 0; i < w; ++i) before += warp_cnt[i];
    const int rank = before + __popc(bal & ((1u << lane) - 1));
    if (s < S) {
      const int n_b = row_off[b + 1] - row_off[b];
      slot_idx[(int64_t)b * S + s] = (flag && rank < n_b) ? row_off[b] + rank : -1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += warp_cnt[i];
      running += tot;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ embedding + LN forward
struct EmbedParams {
  const int64_t* ids;     // [M]
  const int* slot_idx;    // [M] or null (decoder)
  const float* tok_emb;   // [V, d] fp32 master
  const float* pos_emb;   // [npos, d]
  const float* vis_acc;   // [R, d] fp32: feats . Wfeat^T (no bias)
  const float* boxes;     // [R, 4]
  const float* w_box;     // [d, 4]
  const float* b_img;     // [d]
  const float* gamma;
  const float* beta;
  float* pre;             // [M, d] LN input (kept for backward) or null
  float* out_f32;         // [M, d]
  bf16* out_bf16;         // [M, d]
  float* mean;            // [M]
  float* rstd;            // [M]
  int M, S, d, pos_offset;
  const int* pos_index;   // device scalar: explicit position for every row (cached decode) or null
  float embed_scale;
  DropCfg drop;
};

template <int NV>
__global__ void __launch_bounds__(256) embed_ln_fwd_kernel(const EmbedParams p) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= p.M) return;
  const int d = p.d;
  const int s = p.pos_index ? *p.pos_index : row % p.S;
  const int slot = p.slot_idx ? p.slot_idx[row] : -1;
  const float* pos = p.pos_emb + (int64_t)(s + p.pos_offset) * d;
  float4 x[NV];
  if (slot >= 0) {
    const float* va = p.vis_acc + (int64_t)slot * d;
    const float4 bx = __ldg(reinterpret_cast<const float4*>(p.boxes + (int64_t)slot * 4));
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        float4 v = __ldg(reinterpret_cast<const float4*>(va + c));
        const float4 bi = __ldg(reinterpret_cast<const float4*>(p.b_img + c));
        float o[4] = {v.x + bi.x, v.y + bi.y, v.z + bi.z, v.w + bi.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(p.w_box + (int64_t)(c + j) * 4));
          o[j] += bx.x * w.x + bx.y * w.y + bx.z * w.z + bx.w * w.w;
        }
        x[i] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  } else {
    const float* te = p.tok_emb + (int64_t)p.ids[row] * d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) x[i] = __ldg(reinterpret_cast<const float4*>(te + c));
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + c));
      x[i].x = x[i].x * p.embed_scale + pe.x;
      x[i].y = x[i].y * p.embed_scale + pe.y;
      x[i].z = x[i].z * p.embed_scale + pe.z;
      x[i].w = x[i].w * p.embed_scale + pe.w;
      sum += x[i].x + x[i].y + x[i].z + x[i].w;
      if (p.pre) *reinterpret_cast<float4*>(p.pre + (int64_t)row * d + c) = x[i];
    }
  }
  const float mean = warp_sum(sum) / d;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      const float a = x[i].x - mean, b = x[i].y - mean, cc = x[i].z - mean, e = x[i].w - mean;
      var += a * a + b * b + cc * cc + e * e;
    }
  }
  const float rstd = rsqrtf(warp_sum(var) / d + 1e-5f);
  if (lane == 0) {
    if (p.mean) p.mean[row] = mean;
    if (p.rstd) p.rstd[row] = rstd;
  }
  const uint64_t seed = p.drop.thresh16 ? *p.drop.seed : 0ull;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + c));
      float4 y;
      y.x = (x[i].x - mean) * rstd * g.x + b.x;
      y.y = (x[i].y - mean) * rstd * g.y + b.y;
      y.z = (x[i].z - mean) * rstd * g.z + b.z;
      y.w = (x[i].w - mean) * rstd * g.w + b.w;
      if (p.drop.thresh16) y = drop4(y, seed, p.drop, ((uint64_t)row * d + c) >> 2);
      if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + (int64_t)row * d + c) = y;
      if (p.out_bf16) *reinterpret_cast<uint2*>(p.out_bf16 + (int64_t)row * d + c) = pack4(y);
    }
  }
}

// ------------------------------------------------------------------ residual + dropout + LayerNorm forward
// pre = residual + dropout(z)   (z = bf16 output of the preceding Linear, bias included)
// y   = LN(pre)                 written as fp32 (residual stream) and bf16 (next GEMM operand)
// With z == nullptr the kernel is a plain LayerNorm of `pre_in`.
// HF-3.0.2 EncoderLayer/DecoderLayer: x = LN(residual + F.dropout(sublayer(x))).
struct LnFwdParams {
  const bf16* z;          // [M, d] or null
  const float* residual;  // [M, d] (or the LN input itself when z is null)
  const float* gamma;
  const float* beta;
  float* pre;             // [M, d] fp32 LN input, stored for backward (may be null)
  float* out_f32;
  bf16* out_bf16;
  float* mean;
  float* rstd;
  int M, d;
  DropCfg drop;
};

template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const LnFwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= p.M) return;
  const int d = p.d;
  float4 x[NV];
  float sum = 0.f;
  const uint32_t key = p.drop.thresh16 ? dropout_key(*p.drop.seed, p.drop.tag) : 0u;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      x[i] = *reinterpret_cast<const float4*>(p.residual + (int64_t)row * d + c);
      if (p.z) {
        const uint2 zz = *reinterpret_cast<const uint2*>(p.z + (int64_t)row * d + c);
        float4 zv = make_float4(__uint_as_float(zz.x << 16), __uint_as_float(zz.x & 0xFFFF0000u),
                                __uint_as_float(zz.y << 16), __uint_as_float(zz.y & 0xFFFF0000u));
        if (p.drop.thresh16) {
          const uint64_t bits = dropout_bits4_k(key, ((uint64_t)row * d + c) >> 2);
          zv.x = dropout_keep(bits, 0, p.drop.thresh16) ? zv.x * p.drop.scale : 0.f;
          zv.y = dropout_keep(bits, 1, p.drop.thresh16) ? zv.y * p.drop.scale : 0.f;
          zv.z = dropout_keep(bits, 2, p.drop.thresh16) ? zv.z * p.drop.scale : 0.f;
          zv.w = dropout_keep(bits, 3, p.drop.thresh16) ? zv.w * p.drop.scale : 0.f;
        }
        x[i].x += zv.x; x[i].y += zv.y; x[i].z += zv.z; x[i].w += zv.w;
        if (p.pre) *reinterpret_cast<float4*>(p.pre + (int64_t)row * d + c) = x[i];
      }
      sum += x[i].x + x[i].y + x[i].z + x[i].w;
    }
  }
  const float mean = warp_sum(sum) / d;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      const float a = x[i].x - mean, b = x[i].y - mean, cc = x[i].z - mean, e = x[i].w - mean;
      var += a * a + b * b + cc * cc + e * e;
    }
  }
  const float rstd = rsqrtf(warp_sum(var) / d + 1e-5f);
  if (lane == 0) {
    if (p.mean) p.mean[row] = mean;
    if (p.rstd) p.rstd[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + c));
      float4 y;
      y.x = (x[i].x - mean) * rstd * g.x + b.x;
      y.y = (x[i].y - mean) * rstd * g.y + b.y;
      y.z = (x[i].z - mean) * rstd * g.z + b.z;
      y.w = (x[i].w - mean) * rstd * g.w + b.w;
      if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + (int64_t)row * d + c) = y;
      if (p.out_bf16) *reinterpret_cast<uint2*>(p.out_bf16 + (int64_t)row * d + c) = pack4(y);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// dy  : grad w.r.t. the LN output (fp32).  If drop_in is on, the forward applied dropout to
//       the LN output (embedding LN), so dy is re-masked on read.
// dpre: grad w.r.t. the LN input (fp32) — the residual-stream gradient.
// dz  : bf16 copy of dpre, re-masked with drop_out (the dropout that sat between the preceding
//       Linear and the residual add) — the dY operand of that Linear's dgrad / wgrad GEMMs.
// dgamma/dbeta/dbias (+=): column sums of dy*xhat, dy, dz.  Zero them before the first call.
struct LnBwdParams {
  const float* dy;    // fp32 part of the incoming gradient (residual path) or null
  const bf16* dy_b;   // bf16 part (dgrad GEMM output of the consumer Linear) or null
  const float* pre;
  const float* mean;
  const float* rstd;
  const float* gamma;
  float* dpre;
  bf16* dz;
  float* dgamma;
  float* dbeta;
  float* dbias;
  int M, d;
  DropCfg drop_in, drop_out;
};

// Column-owner layout: thread t owns the float4 column group t of every row the block visits, so
// the dgamma / dbeta / dbias column sums are three float4 registers per thread (no big per-warp
// accumulator arrays -> high occupancy), and a row is read by d/4 consecutive threads (coalesced).
// Row statistics (mean of g, mean of g*xhat) are block-reduced for LN_RB rows at a time.
constexpr int LN_RB = 4;
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdParams p, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  const int d = p.d;
  const int c = threadIdx.x * 4;
  const bool active = c < d;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __shared__ float red[2][LN_RB][2][8];
  const uint32_t key_in = p.drop_in.thresh16 ? dropout_key(*p.drop_in.seed, p.drop_in.tag) : 0u;
  const uint32_t key_out = p.drop_out.thresh16 ? dropout_key(*p.drop_out.seed, p.drop_out.tag) : 0u;
  float4 gm = make_float4(0, 0, 0, 0);
  if (active) gm = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
  float4 ag = make_float4(0, 0, 0, 0), ab = ag, az = ag;
  const int r_begin = blockIdx.x * rows_per_block;
  const int r_end = min(p.M, r_begin + rows_per_block);
  int buf = 0;
  for (int r0 = r_begin; r0 < r_end; r0 += LN_RB, buf ^= 1) {
    float4 g[LN_RB], xh[LN_RB], dyv[LN_RB];
    float rs[LN_RB];
#pragma unroll
    for (int i = 0; i < LN_RB; ++i) {
      const int row = r0 + i;
      g[i] = xh[i] = dyv[i] = make_float4(0, 0, 0, 0);
      rs[i] = 0.f;
      float s1 = 0.f, s2 = 0.f;
      if (row < r_end && active) {
        float4 dy = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.dy) dy = *reinterpret_cast<const float4*>(p.dy + (int64_t)row * d + c);
        if (p.dy_b) {
          const uint2 t = *reinterpret_cast<const uint2*>(p.dy_b + (int64_t)row * d + c);
          dy.x += __uint_as_float(t.x << 16); dy.y += __uint_as_float(t.x & 0xFFFF0000u);
          dy.z += __uint_as_float(t.y << 16); dy.w += __uint_as_float(t.y & 0xFFFF0000u);
        }
        if (p.drop_in.thresh16) {
          const uint64_t bits = dropout_bits4_k(key_in, ((uint64_t)row * d + c) >> 2);
          dy.x = dropout_keep(bits, 0, p.drop_in.thresh16) ? dy.x * p.drop_in.scale : 0.f;
          dy.y = dropout_keep(bits, 1, p.drop_in.thresh16) ? dy.y * p.drop_in.scale : 0.f;
          dy.z = dropout_keep(bits, 2, p.drop_in.thresh16) ? dy.z * p.drop_in.scale : 0.f;
          dy.w = dropout_keep(bits, 3, p.drop_in.thresh16) ? dy.w * p.drop_in.scale : 0.f;
        }
        const float4 x = *reinterpret_cast<const float4*>(p.pre + (int64_t)row * d + c);
        const float mean = __ldg(p.mean + row), rstd = __ldg(p.rstd + row);
        rs[i] = rstd;
        dyv[i] = dy;
        xh[i] = make_float4((x.x - mean) * rstd, (x.y - mean) * rstd, (x.z - mean) * rstd, (x.w - mean) * rstd);
        g[i] = make_float4(dy.x * gm.x, dy.y * gm.y, dy.z * gm.z, dy.w * gm.w);
        s1 = g[i].x + g[i].y + g[i].z + g[i].w;
        s2 = g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) { red[buf][i][0][wib] = s1; red[buf][i][1][wib] = s2; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < LN_RB; ++i) {
      const int row = r0 + i;
      if (row < r_end && active) {
        float s1 = 0.f, s2 = 0.f;
        for (int w = 0; w < nw; ++w) { s1 += red[buf][i][0][w]; s2 += red[buf][i][1][w]; }
        s1 /= d; s2 /= d;
        float4 dx;
        dx.x = rs[i] * (g[i].x - s1 - xh[i].x * s2);
        dx.y = rs[i] * (g[i].y - s1 - xh[i].y * s2);
        dx.z = rs[i] * (g[i].z - s1 - xh[i].z * s2);
        dx.w = rs[i] * (g[i].w - s1 - xh[i].w * s2);
        if (p.dpre) *reinterpret_cast<float4*>(p.dpre + (int64_t)row * d + c) = dx;
        ag.x += dyv[i].x * xh[i].x; ag.y += dyv[i].y * xh[i].y; ag.z += dyv[i].z * xh[i].z; ag.w += dyv[i].w * xh[i].w;
        ab.x += dyv[i].x; ab.y += dyv[i].y; ab.z += dyv[i].z; ab.w += dyv[i].w;
        if (p.dz || p.dbias) {
          float4 z = dx;
          if (p.drop_out.thresh16) {
            const uint64_t bits = dropout_bits4_k(key_out, ((uint64_t)row * d + c) >> 2);
            z.x = dropout_keep(bits, 0, p.drop_out.thresh16) ? z.x * p.drop_out.scale : 0.f;
            z.y = dropout_keep(bits, 1, p.drop_out.thresh16) ? z.y * p.drop_out.scale : 0.f;
            z.z = dropout_keep(bits, 2, p.drop_out.thresh16) ? z.z * p.drop_out.scale : 0.f;
            z.w = dropout_keep(bits, 3, p.drop_out.thresh16) ? z.w * p.drop_out.scale : 0.f;
          }
          if (p.dz) *reinterpret_cast<uint2*>(p.dz + (int64_t)row * d + c) = pack4(z);
          az.x += z.x; az.y += z.y; az.z += z.z; az.w += z.w;
        }
      }
    }
  }
  if (active) {
    if (p.dgamma) { atomicAdd(p.dgamma + c, ag.x); atomicAdd(p.dgamma + c + 1, ag.y); atomicAdd(p.dgamma + c + 2, ag.z); atomicAdd(p.dgamma + c + 3, ag.w); }
    if (p.dbeta) { atomicAdd(p.dbeta + c, ab.x); atomicAdd(p.dbeta + c + 1, ab.y); atomicAdd(p.dbeta + c + 2, ab.z); atomicAdd(p.dbeta + c + 3, ab.w); }
    if (p.dbias) { atomicAdd(p.dbias + c, az.x); atomicAdd(p.dbias + c + 1, az.y); atomicAdd(p.dbias + c + 2, az.z); atomicAdd(p.dbias + c + 3, az.w); }
  }
}

// ------------------------------------------------------------------ LayerNorm backward, bulk-copy pipelined
// Same arithmetic as ln_bwd_kernel, restructured for HBM bandwidth: one persistent CTA per SM owns a
// contiguous row range and streams it through shared memory in LNP_ROWS-row tiles with 1-D bulk copies
// (cp.async.bulk -> mbarrier complete_tx), LNP stages deep, so 100+ KB per SM are in flight without
// holding registers.  Per tile: phase A = one warp per row computes the two row statistics from smem;
// phase B = thread t owns float4 column group t, computes dpre / dz for every row of the tile and keeps
// the dgamma / dbeta / dbias column sums in 12 registers for the CTA's whole range; they are flushed once
// with red.global.add.v4.f32 (148 CTAs x d/4 vector reductions instead of ~400 x 3d scalar atomics, whose
// per-address serialisation in L2 was a ~15 us tail on every launch).
constexpr int LNP_ROWS = 4;      // rows per tile; 8 / LNP_ROWS warps share a row in phase A
constexpr int LNP_THREADS = 256;
constexpr int LNP_WPR = (LNP_THREADS / 32) / LNP_ROWS;
constexpr int LNP_CTAS_PER_SM = 2;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 lnp_load_dy(const float* sdy, const bf16* sdyb, int idx4, const DropCfg& din, uint32_t key_in,
                                              uint64_t q) {
  float4 dy = make_float4(0.f, 0.f, 0.f, 0.f);
  if (sdy) dy = *reinterpret_cast<const float4*>(sdy + 4 * idx4);
  if (sdyb) {
    const uint2 t = *reinterpret_cast<const uint2*>(sdyb + 4 * idx4);
    dy.x += __uint_as_float(t.x << 16); dy.y += __uint_as_float(t.x & 0xFFFF0000u);
    dy.z += __uint_as_float(t.y << 16); dy.w += __uint_as_float(t.y & 0xFFFF0000u);
  }
  if (din.thresh16) {
    const uint64_t bits = dropout_bits4_k(key_in, q);
    dy.x = dropout_keep(bits, 0, din.thresh16) ? dy.x * din.scale : 0.f;
    dy.y = dropout_keep(bits, 1, din.thresh16) ? dy.y * din.scale : 0.f;
    dy.z = dropout_keep(bits, 2, din.thresh16) ? dy.z * din.scale : 0.f;
    dy.w = dropout_keep(bits, 3, din.thresh16) ? dy.w * din.scale : 0.f;
  }
  return dy;
}

__global__ void __launch_bounds__(LNP_THREADS, LNP_CTAS_PER_SM) ln_bwd_pipe_kernel(const LnBwdParams p, int rows_per_block, int stages, int vec_red) {
  extern __shared__ __align__(128) uint8_t lnp_smem[];
  pdl_trigger();
  const int d = p.d, d4 = d >> 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* full = reinterpret_cast<uint64_t*>(lnp_smem);                  // [stages] (<= 8)
  float* s_stat = reinterpret_cast<float*>(lnp_smem + 64);                 // [LNP_ROWS][LNP_WPR][4]: s1, s2, mean, rstd partials
  float* s_gamma = reinterpret_cast<float*>(lnp_smem + 256);               // [d]
  uint8_t* stage0 = lnp_smem + 256 + ((d * 4 + 127) & ~127);
  const uint32_t off_dyb = p.dy ? LNP_ROWS * d * 4 : 0;
  const uint32_t off_pre = off_dyb + (p.dy_b ? LNP_ROWS * d * 2 : 0);
  const uint32_t stage_bytes = off_pre + LNP_ROWS * d * 4;
  const int r_begin = blockIdx.x * rows_per_block;
  const int r_end = min(p.M, r_begin + rows_per_block);
  const int ntiles = r_end > r_begin ? (r_end - r_begin + LNP_ROWS - 1) / LNP_ROWS : 0;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  const uint32_t key_in = p.drop_in.thresh16 ? dropout_key(*p.drop_in.seed, p.drop_in.tag) : 0u;
  const uint32_t key_out = p.drop_out.thresh16 ? dropout_key(*p.drop_out.seed, p.drop_out.tag) : 0u;
  pdl_wait();
  for (int i = tid; i < d; i += LNP_THREADS) s_gamma[i] = __ldg(p.gamma + i);
  __syncthreads();

  auto issue = [&](int tile) {
    const int s = tile % stages;
    const int row0 = r_begin + tile * LNP_ROWS;
    const uint32_t nr = (uint32_t)min(LNP_ROWS, r_end - row0);
    uint8_t* st = stage0 + (size_t)s * stage_bytes;
    const uint32_t b4 = nr * d * 4, b2 = nr * d * 2;
    mbar_arrive_expect_tx(&full[s], (p.dy ? b4 : 0) + (p.dy_b ? b2 : 0) + b4);
    if (p.dy) bulk_load_1d(st, p.dy + (int64_t)row0 * d, b4, &full[s]);
    if (p.dy_b) bulk_load_1d(st + off_dyb, p.dy_b + (int64_t)row0 * d, b2, &full[s]);
    bulk_load_1d(st + off_pre, p.pre + (int64_t)row0 * d, b4, &full[s]);
  };
  if (tid == 0)
    for (int t = 0; t < stages && t < ntiles; ++t) issue(t);

  float4 ag = make_float4(0, 0, 0, 0), ab = ag, az = ag;
  const bool owner = tid < d4;
  float4 gm = make_float4(0, 0, 0, 0);
  if (owner) gm = *reinterpret_cast<const float4*>(s_gamma + 4 * tid);
  const float inv_d = 1.f / d;

  for (int tile = 0; tile < ntiles; ++tile) {
    const int s = tile % stages;
    const uint32_t parity = (uint32_t)(tile / stages) & 1u;
    const int row0 = r_begin + tile * LNP_ROWS;
    const int nr = min(LNP_ROWS, r_end - row0);
    const uint8_t* st = stage0 + (size_t)s * stage_bytes;
    const float* sdy = p.dy ? reinterpret_cast<const float*>(st) : nullptr;
    const bf16* sdyb = p.dy_b ? reinterpret_cast<const bf16*>(st + off_dyb) : nullptr;
    const float* spre = reinterpret_cast<const float*>(st + off_pre);
    // row statistics of this warp's row are independent of the tile data: fetch them while the copy lands
    const int ar = warp / LNP_WPR, apart = warp % LNP_WPR;
    float mean = 0.f, rstd = 0.f;
    if (ar < nr) { mean = __ldg(p.mean + row0 + ar); rstd = __ldg(p.rstd + row0 + ar); }
    mbar_wait(&full[s], parity);
    // ---- phase A: LNP_WPR warps reduce one row (interleaved 32-float4 slices)
    if (ar < nr) {
      const int row = row0 + ar;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 3
      for (int f = apart * 32 + lane; f < d4; f += 32 * LNP_WPR) {
        const float4 dy = lnp_load_dy(sdy ? sdy + ar * d : nullptr, sdyb ? sdyb + ar * d : nullptr, f, p.drop_in, key_in,
                                      (uint64_t)row * d4 + f);
        const float4 x = *reinterpret_cast<const float4*>(spre + ar * d + 4 * f);
        const float4 g4 = *reinterpret_cast<const float4*>(s_gamma + 4 * f);
        const float gx = dy.x * g4.x, gy = dy.y * g4.y, gz = dy.z * g4.z, gw = dy.w * g4.w;
        s1 += gx + gy + gz + gw;
        s2 += gx * ((x.x - mean) * rstd) + gy * ((x.y - mean) * rstd) + gz * ((x.z - mean) * rstd) + gw * ((x.w - mean) * rstd);
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) *reinterpret_cast<float4*>(s_stat + 4 * warp) = make_float4(s1, s2, mean, rstd);
    }
    __syncthreads();
    // ---- phase B: thread t owns float4 column group t of every row
    if (owner) {
#pragma unroll
      for (int r = 0; r < LNP_ROWS; ++r) {
        if (r >= nr) break;
        const int row = row0 + r;
        float4 stt = *reinterpret_cast<const float4*>(s_stat + 4 * r * LNP_WPR);
#pragma unroll
        for (int w = 1; w < LNP_WPR; ++w) {
          const float4 t2 = *reinterpret_cast<const float4*>(s_stat + 4 * (r * LNP_WPR + w));
          stt.x += t2.x; stt.y += t2.y;
        }
        stt.x *= inv_d; stt.y *= inv_d;
        const float4 dy = lnp_load_dy(sdy ? sdy + r * d : nullptr, sdyb ? sdyb + r * d : nullptr, tid, p.drop_in, key_in,
                                      (uint64_t)row * d4 + tid);
        const float4 x = *reinterpret_cast<const float4*>(spre + r * d + 4 * tid);
        const float4 xh = make_float4((x.x - stt.z) * stt.w, (x.y - stt.z) * stt.w, (x.z - stt.z) * stt.w, (x.w - stt.z) * stt.w);
        float4 dx;
        dx.x = stt.w * (dy.x * gm.x - stt.x - xh.x * stt.y);
        dx.y = stt.w * (dy.y * gm.y - stt.x - xh.y * stt.y);
        dx.z = stt.w * (dy.z * gm.z - stt.x - xh.z * stt.y);
        dx.w = stt.w * (dy.w * gm.w - stt.x - xh.w * stt.y);
        if (p.dpre) *reinterpret_cast<float4*>(p.dpre + (int64_t)row * d + 4 * tid) = dx;
        ag.x += dy.x * xh.x; ag.y += dy.y * xh.y; ag.z += dy.z * xh.z; ag.w += dy.w * xh.w;
        ab.x += dy.x; ab.y += dy.y; ab.z += dy.z; ab.w += dy.w;
        if (p.dz || p.dbias) {
          float4 z = dx;
          if (p.drop_out.thresh16) {
            const uint64_t bits = dropout_bits4_k(key_out, (uint64_t)row * d4 + tid);
            z.x = dropout_keep(bits, 0, p.drop_out.thresh16) ? z.x * p.drop_out.scale : 0.f;
            z.y = dropout_keep(bits, 1, p.drop_out.thresh16) ? z.y * p.drop_out.scale : 0.f;
            z.z = dropout_keep(bits, 2, p.drop_out.thresh16) ? z.z * p.drop_out.scale : 0.f;
            z.w = dropout_keep(bits, 3, p.drop_out.thresh16) ? z.w * p.drop_out.scale : 0.f;
          }
          if (p.dz) *reinterpret_cast<uint2*>(p.dz + (int64_t)row * d + 4 * tid) = pack4(z);
          az.x += z.x; az.y += z.y; az.z += z.z; az.w += z.w;
        }
      }
    }
    __syncthreads();   // every generic read of this stage is done
    if (tid == 0 && tile + stages < ntiles) {
      fence_proxy_async_smem();
      issue(tile + stages);
    }
  }
  if (owner) {
    const int c = 4 * tid;
    if (vec_red) {
      if (p.dgamma) red_add_v4(p.dgamma + c, ag);
      if (p.dbeta) red_add_v4(p.dbeta + c, ab);
      if (p.dbias) red_add_v4(p.dbias + c, az);
    } else {
      if (p.dgamma) { atomicAdd(p.dgamma + c, ag.x); atomicAdd(p.dgamma + c + 1, ag.y); atomicAdd(p.dgamma + c + 2, ag.z); atomicAdd(p.dgamma + c + 3, ag.w); }
      if (p.dbeta) { atomicAdd(p.dbeta + c, ab.x); atomicAdd(p.dbeta + c + 1, ab.y); atomicAdd(p.dbeta + c + 2, ab.z); atomicAdd(p.dbeta + c + 3, ab.w); }
      if (p.dbias) { atomicAdd(p.dbias + c, az.x); atomicAdd(p.dbias + c + 1, az.y); atomicAdd(p.dbias + c + 2, az.z); atomicAdd(p.dbias + c + 3, az.w); }
    }
  }
}

// ------------------------------------------------------------------ embedding backward
// demb fp32 [M, d] = grad w.r.t. (tok|vis)*scale + pos.  Token rows scatter-add into the shared
// embedding gradient (skipping padding_idx like nn.Embedding), visual rows are written to
// dvis (bf16 for the wgrad GEMM); positions are reduced over the batch without atomics.
__global__ void __launch_bounds__(256) embed_bwd_scatter_kernel(const float* demb, const int64_t* ids, const int* slot_idx,
                                                                float* d_tok, bf16* dvis, int M, int d, int pad_id,
                                                                float embed_scale) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const int slot = slot_idx ? slot_idx[row] : -1;
  const int64_t tok = ids[row];
  for (int c = lane * 4; c < d; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(demb + (int64_t)row * d + c);
    v.x *= embed_scale; v.y *= embed_scale; v.z *= embed_scale; v.w *= embed_scale;
    if (slot >= 0) {
      *reinterpret_cast<uint2*>(dvis + (int64_t)slot * d + c) = pack4(v);
    } else if (tok != pad_id) {
      float* dst = d_tok + tok * d + c;
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        red_add_v4(dst, v);   // one 16-byte reduction instead of four scalar atomics
      } else {
        atomicAdd(dst, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
      }
    }
  }
}

// dpos[s + off, :] (+)= sum_b demb[b, s, :]
// grid (S, nsplit): block (s, y) sums batch elements y, y + nsplit, ... of position s; with nsplit > 1 the partial sums are
// combined with atomics onto the (pre-cleared / accumulating) gradient, so a 100-position launch is not limited to 100 CTAs
__global__ void pos_grad_kernel(const float* demb, float* dpos, int B, int S, int d, int pos_offset, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int s = blockIdx.x, nsplit = gridDim.y;
  for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 4
    for (int b = blockIdx.y; b < B; b += nsplit) {
      const float4 v = *reinterpret_cast<const float4*>(demb + ((int64_t)b * S + s) * d + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* dst = dpos + (int64_t)(s + pos_offset) * d + c;
    if (nsplit > 1) {
      atomicAdd(dst, acc.x); atomicAdd(dst + 1, acc.y); atomicAdd(dst + 2, acc.z); atomicAdd(dst + 3, acc.w);
    } else {
      float4* d4 = reinterpret_cast<float4*>(dst);
      if (accumulate) {
        const float4 o = *d4;
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
      *d4 = acc;
    }
  }
}
__global__ void box_wgrad_kernel(const bf16* dvis, const float* boxes, float* dw_img, int R, int d, int ld_w,
                                 int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 8
  for (int r = r0; r < r1; ++r) {
    const float g = __bfloat162float(dvis[(int64_t)r * d + c]);
    const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes + (int64_t)r * 4));
    a0 += g * bx.x; a1 += g * bx.y; a2 += g * bx.z; a3 += g * bx.w;
  }
  float* dst = dw_img + (int64_t)c * ld_w + FEAT;
  atomicAdd(dst, a0); atomicAdd(dst + 1, a1); atomicAdd(dst + 2, a2); atomicAdd(dst + 3, a3);
}

// out[n] += sum_m x[m, n]   (bf16 in, fp32 out) — bias gradients of qkv / fc1 / heads
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* x, int64_t ld, float* out, int M, int N,
                                                          int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (c >= N) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float a = 0.f, b = 0.f;
  for (int r = r0; r < r1; ++r) {
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(x + (int64_t)r * ld + c);
    a += __bfloat162float(v.x);
    b += __bfloat162float(v.y);
  }
  atomicAdd(out + c, a);
  if (c + 1 < N) atomicAdd(out + c + 1, b);
}

// gather rows: out[i, :] = src[idx[i], :] (bf16), used by the pretraining heads
// Wide variant: a block covers 256 columns x rows_per_block rows; thread = (8-column group, row phase 0..7), 16-byte
// loads with eight rows in flight per thread, row phases combined through shared memory, one atomic per column.
__global__ void __launch_bounds__(256) colsum_bf16_wide_kernel(const bf16* x, int64_t ld, float* out, int M, int N,
                                                               int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  __shared__ float part[8][256 + 8];
  const int cg = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < N) {
    const bf16* px = x + c;
    for (int r = r0 + rg; r < r1; r += 64) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int rr = r + 8 * u;
        v[u] = rr < r1 ? __ldg(reinterpret_cast<const uint4*>(px + (int64_t)rr * ld)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(w[j] << 16);
          acc[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[rg][cg * 8 + j] = acc[j];
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < N) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += part[g][threadIdx.x];
    atomicAdd(out + col, t);
  }
}

__global__ void gather_rows_bf16_kernel(const bf16* src, int64_t ld_src, const int* idx, bf16* out, int64_t ld_out,
                                        int n, int d) {
  const int row = blockIdx.x;
  if (row >= n) return;
  const int s = idx[row];
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8)
    *reinterpret_cast<uint4*>(out + (int64_t)row * ld_out + c) =
        *reinterpret_cast<const uint4*>(src + (int64_t)s * ld_src + c);
}

// scatter-add rows: dst[idx[i], :] += src[i, :] (bf16 src, fp32 dst) — head gradients back into dH
__global__ void scatter_add_rows_kernel(const bf16* src, int64_t ld_src, const int* idx, float* dst, int64_t ld_dst,
                                        int n, int d) {
  const int row = blockIdx.x;
  if (row >= n) return;
  const int t = idx[row];
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    atomicAdd(dst + (int64_t)t * ld_dst + c, __bfloat162float(src[(int64_t)row * ld_src + c]));
}

static DropCfg make_drop(float p, uint32_t tag, const uint64_t* seed) {
  DropCfg dc;
  dc.thresh16 = (p > 0.f && seed) ? (uint32_t)(p * 65536.0f + 0.5f) : 0;
  dc.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  dc.tag = tag;
  dc.seed = seed;
  return dc;
}

}  // namespace kmb

using namespace kmb;

extern "C" int kmb_pack_features(const float* const* feat_ptrs, const int* row_offsets, int B, const float* packed,
                                 void* feats_bf16, float* boxes, int R, kmb_stream_t stream) {
  if ((!feat_ptrs && !packed) || (feat_ptrs && !row_offsets) || !feats_bf16 || !boxes || R < 0) {
    kmb_set_last_error("kmb_pack_features: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if (R == 0) return KMB_OK;
  launch_pdl(pack_features_kernel, dim3(R), dim3(256), 0, (cudaStream_t)stream, feat_ptrs, row_offsets, B, packed, (bf16*)feats_bf16, boxes, R);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_slot_index(const int64_t* input_ids, const int* row_offsets, int B, int S, int img_feat_id,
                              int cls_token_id, int* slot_idx, kmb_stream_t stream) {
  if (!input_ids || !row_offsets || !slot_idx || B <= 0 || S <= 0) {
    kmb_set_last_error("kmb_slot_index: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  slot_index_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(input_ids, row_offsets, S, img_feat_id, cls_token_id, slot_idx);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

#define KMB_DISPATCH_NV(d, CALL6, CALL8)                                              \
  if ((d) % 128 == 0 && (d) <= 768) { CALL6; }                                        \
  else if ((d) % 4 == 0 && (d) <= 1024) { CALL8; }                                    \
  else { kmb_set_last_error("d_model must be a multiple of 4 and <= 1024", __FILE__, __LINE__); return KMB_ERR_ARG; }

extern "C" int kmb_embed_ln_fwd(const int64_t* ids, const int* slot_idx, const float* tok_emb, const float* pos_emb,
                                const float* vis_acc, const float* boxes, const float* w_box, const float* b_img,
                                const float* gamma, const float* beta, float* pre, float* out_f32, void* out_bf16,
                                float* mean, float* rstd, int M, int S, int d, int pos_offset, const int* pos_index,
                                float embed_scale, float dropout_p, uint32_t dropout_tag, const uint64_t* dropout_seed,
                                kmb_stream_t stream) {
  if (!ids || !tok_emb || !pos_emb || !gamma || !beta || M <= 0 || S <= 0) {
    kmb_set_last_error("kmb_embed_ln_fwd: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  EmbedParams p;
  p.ids = ids; p.slot_idx = slot_idx; p.tok_emb = tok_emb; p.pos_emb = pos_emb; p.vis_acc = vis_acc; p.boxes = boxes;
  p.w_box = w_box; p.b_img = b_img; p.gamma = gamma; p.beta = beta; p.pre = pre; p.out_f32 = out_f32;
  p.out_bf16 = (bf16*)out_bf16; p.mean = mean; p.rstd = rstd; p.M = M; p.S = S; p.d = d; p.pos_offset = pos_offset;
  p.pos_index = pos_index; p.embed_scale = embed_scale; p.drop = make_drop(dropout_p, dropout_tag, dropout_seed);
  const int blocks = (M * 32 + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  KMB_DISPATCH_NV(d, (launch_pdl(embed_ln_fwd_kernel<6>, dim3(blocks), dim3(256), 0, st, p)), (launch_pdl(embed_ln_fwd_kernel<8>, dim3(blocks), dim3(256), 0, st, p)));
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_layernorm_fwd(const void* z_bf16, const float* residual, const float* gamma, const float* beta,
                                 float* pre, float* out_f32, void* out_bf16, float* mean, float* rstd, int M, int d,
                                 float dropout_p, uint32_t dropout_tag, const uint64_t* dropout_seed, kmb_stream_t stream) {
  if (!residual || !gamma || !beta || M <= 0) {
    kmb_set_last_error("kmb_layernorm_fwd: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  LnFwdParams p;
  p.z = (const bf16*)z_bf16; p.residual = residual; p.gamma = gamma; p.beta = beta; p.pre = pre; p.out_f32 = out_f32;
  p.out_bf16 = (bf16*)out_bf16; p.mean = mean; p.rstd = rstd; p.M = M; p.d = d;
  p.drop = make_drop(z_bf16 ? dropout_p : 0.f, dropout_tag, dropout_seed);
  const int blocks = (M * 32 + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  KMB_DISPATCH_NV(d, (launch_pdl(ln_fwd_kernel<6>, dim3(blocks), dim3(256), 0, st, p)), (launch_pdl(ln_fwd_kernel<8>, dim3(blocks), dim3(256), 0, st, p)));
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_layernorm_bwd(const float* dy, const void* dy_bf16, const float* pre, const float* mean, const float* rstd,
                                 const float* gamma, float* dpre, void* dz_bf16, float* dgamma, float* dbeta,
                                 float* dbias, int M, int d, float drop_in_p, uint32_t drop_in_tag, float drop_out_p,
                                 uint32_t drop_out_tag, const uint64_t* dropout_seed, kmb_stream_t stream) {
  if ((!dy && !dy_bf16) || !pre || !mean || !rstd || !gamma || M <= 0) {
    kmb_set_last_error("kmb_layernorm_bwd: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  LnBwdParams p;
  p.dy = dy; p.dy_b = (const bf16*)dy_bf16; p.pre = pre; p.mean = mean; p.rstd = rstd; p.gamma = gamma; p.dpre = dpre; p.dz = (bf16*)dz_bf16;
  p.dgamma = dgamma; p.dbeta = dbeta; p.dbias = dbias; p.M = M; p.d = d;
  p.drop_in = make_drop(drop_in_p, drop_in_tag, dropout_seed);
  p.drop_out = make_drop(drop_out_p, drop_out_tag, dropout_seed);
  if ((d % 4) || d > 1024) {
    kmb_set_last_error("kmb_layernorm_bwd: d_model must be a multiple of 4 and <= 1024", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  // bulk-copy pipelined kernel: needs 16-byte aligned operands and rows that are a multiple of 16 bytes in bf16
  {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    static int sm_count = 0, smem_optin = 0;
    if (!sm_count) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    const size_t stage_bytes = (size_t)LNP_ROWS * d * ((dy ? 4 : 0) + (dy_bf16 ? 2 : 0) + 4);
    const size_t fixed = 256 + (((size_t)d * 4 + 127) & ~(size_t)127);
    int stages = (int)((((size_t)smem_optin + 1024) / LNP_CTAS_PER_SM - 2048 - fixed) / stage_bytes);
    stages = stages > 4 ? 4 : stages;
    if ((d % 8) == 0 && stages >= 2 && al16(dy) && al16(dy_bf16) && al16(pre) && al16(dpre) && al16(dz_bf16) && al16(gamma)) {
      const int vec_red = al16(dgamma) && al16(dbeta) && al16(dbias);
      const int nctas = sm_count * LNP_CTAS_PER_SM;
      int rpb = (M + nctas - 1) / nctas;
      const int nblk = (M + rpb - 1) / rpb;
      const size_t smem = fixed + stages * stage_bytes;
      static size_t smem_set = 0;
      if (smem > smem_set) {
        cudaFuncSetAttribute(ln_bwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
        smem_set = smem_optin;
      }
      launch_pdl(ln_bwd_pipe_kernel, dim3(nblk), dim3(LNP_THREADS), smem, (cudaStream_t)stream, p, rpb, stages, vec_red);
      KMB_CHECK_LAUNCH();
      return KMB_OK;
    }
  }
  // fallback (unaligned views): ~3 blocks per SM, scalar atomics
  int rows_per_block = (M + 148 * 3 - 1) / (148 * 3);
  rows_per_block = (rows_per_block + LN_RB - 1) / LN_RB * LN_RB;
  if (rows_per_block < 2 * LN_RB) rows_per_block = 2 * LN_RB;
  const int blocks = (M + rows_per_block - 1) / rows_per_block;
  const int threads = ((d / 4) + 31) / 32 * 32;
  launch_pdl(ln_bwd_kernel, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, p, rows_per_block);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_embed_bwd(const float* demb, const int64_t* ids, const int* slot_idx, float* d_tok, void* dvis_bf16,
                             float* dpos, int B, int S, int d, int pos_offset, int pad_id, float embed_scale,
                             int accumulate_pos, kmb_stream_t stream) {
  if (!demb || !ids || !d_tok || !dpos || B <= 0 || S <= 0 || (d % 4)) {
    kmb_set_last_error("kmb_embed_bwd: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int M = B * S;
  launch_pdl(embed_bwd_scatter_kernel, dim3((M * 32 + 255) / 256), dim3(256), 0, st, demb, ids, slot_idx, d_tok, (bf16*)dvis_bf16, M, d, pad_id, embed_scale);
  KMB_CHECK_LAUNCH();
  // accumulating launches (the training backward: the gradient buffer was cleared at the start of the sweep) split the batch
  const int nsplit = (accumulate_pos && B >= 16) ? 8 : 1;
  launch_pdl(pos_grad_kernel, dim3(S, nsplit), dim3(192), 0, st, demb, dpos, B, S, d, pos_offset, accumulate_pos);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_box_wgrad(const void* dvis_bf16, const float* boxes, float* dw_img, int R, int d, int ld_w,
                             kmb_stream_t stream) {
  if (!dvis_bf16 || !boxes || !dw_img || R <= 0) {
    kmb_set_last_error("kmb_box_wgrad: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const int rpb = 32;
  launch_pdl(box_wgrad_kernel, dim3(dim3((d + 127) / 128, (R + rpb - 1) / rpb)), dim3(128), 0, (cudaStream_t)stream, (const bf16*)dvis_bf16, boxes, dw_img, R, d, ld_w, rpb);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_colsum_bf16(const void* x, int64_t ld, float* out, int M, int N, kmb_stream_t stream) {
  if (!x || !out || M <= 0 || N <= 0 || (ld % 2) || (N % 2)) {
    kmb_set_last_error("kmb_colsum_bf16: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if ((N % 8) == 0 && (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // 16-byte loads, 8 rows in flight per thread; enough blocks for ~2 per SM
    const int cblocks = (N + 255) / 256;
    int rpb = 256;
    while (rpb > 32 && cblocks * ((M + rpb - 1) / rpb) < 296) rpb >>= 1;
    launch_pdl(colsum_bf16_wide_kernel, dim3(cblocks, (M + rpb - 1) / rpb), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, ld, out, M, N, rpb);
    KMB_CHECK_LAUNCH();
    return KMB_OK;
  }
  const int rpb = 64;
  launch_pdl(colsum_bf16_kernel, dim3(dim3((N / 2 + 255) / 256, (M + rpb - 1) / rpb)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, ld, out, M, N, rpb);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_gather_rows_bf16(const void* src, int64_t ld_src, const int* idx, void* out, int64_t ld_out, int n,
                                    int d, kmb_stream_t stream) {
  if (!src || !idx || !out || n < 0 || (d % 8) || (ld_src % 8) || (ld_out % 8)) {
    kmb_set_last_error("kmb_gather_rows_bf16: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if (n == 0) return KMB_OK;
  gather_rows_bf16_kernel<<<n, 128, 0, (cudaStream_t)stream>>>((const bf16*)src, ld_src, idx, (bf16*)out, ld_out, n, d);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_scatter_add_rows(const void* src_bf16, int64_t ld_src, const int* idx, float* dst, int64_t ld_dst,
                                    int n, int d, kmb_stream_t stream) {
  if (!src_bf16 || !idx || !dst || n < 0) {
    kmb_set_last_error("kmb_scatter_add_rows: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if (n == 0) return KMB_OK;
  scatter_add_rows_kernel<<<n, 256, 0, (cudaStream_t)stream>>>((const bf16*)src_bf16, ld_src, idx, dst, ld_dst, n, d);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}
