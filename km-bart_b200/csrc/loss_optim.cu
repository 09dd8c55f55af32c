// Loss reductions and the optimizer:
//   * the row-combine half of the fused LM-head + CrossEntropyLoss (partials come from the
//     GEMM's KMB_EPI_CE_STATS epilogue)             — src/model/model.py:397-402, :296-301
//   * small-class softmax losses of the pretraining heads (KL-div batchmean / CE mean) with
//     their gradients                                 — src/model/model.py:248-289
//   * fused multi-tensor AdamW with HF-3.0.2 semantics (eps added to the un-corrected
//     sqrt(v), bias correction folded into the step size, decoupled decay after the update)
//     that also refreshes the bf16 shadow weights     — vcg_train.py:100, src/training.py:136-143
//   * fp32 -> bf16 shadow cast, image-projection weight repack, mask inversion.
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

// ------------------------------------------------------------------ LM-head CE: combine partials
// acc[0] += sum of row losses, acc[1] += number of valid rows
__global__ void __launch_bounds__(256) ce_combine_kernel(const float* pmax, const float* psum, const float* label_logit,
                                                         const int64_t* labels, int M, int n_tiles, float* lse_out,
                                                         float* row_loss, float* acc) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  float loss = 0.f, cnt = 0.f;
  if (row < M) {
    float mx = -INFINITY;
    for (int i = lane; i < n_tiles; i += 32) mx = fmaxf(mx, pmax[(int64_t)row * n_tiles + i]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < n_tiles; i += 32)
      s += psum[(int64_t)row * n_tiles + i] * __expf(pmax[(int64_t)row * n_tiles + i] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    if (lane == 0) {
      lse_out[row] = lse;
      const int64_t lab = labels[row];
      if (lab >= 0) {
        loss = lse - label_logit[row];
        cnt = 1.f;
      }
      if (row_loss) row_loss[row] = loss;
    }
  }
  __shared__ float sl[8], sc[8];
  const int wib = threadIdx.x >> 5;
  if (lane == 0) { sl[wib] = loss; sc[wib] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sl[i]; b += sc[i]; }
    if (b > 0.f) { atomicAdd(acc, a); atomicAdd(acc + 1, b); }
  }
}

// loss = factor * acc[0] / acc[1]  (mean over labels != -100, nn.CrossEntropyLoss default)
__global__ void ce_finalize_kernel(const float* acc, float factor, float* loss_out, float* loss_total, int add_total) {
  const float l = factor * acc[0] / acc[1];
  *loss_out = l;
  if (loss_total) *loss_total = add_total ? *loss_total + l : l;
}

// gscale = upstream * factor / n_valid
__global__ void ce_gscale_kernel(const float* acc, const float* upstream, float factor, float* gscale) {
  *gscale = (upstream ? *upstream : 1.f) * factor / acc[1];
}

// ------------------------------------------------------------------ small-class softmax losses
// mode 0: CE with int64 labels, mean over n rows;  mode 1: KL-div(log_softmax(x), soft target)
// "batchmean" = sum / n.  One warp per row.  Adds factor*loss to loss_accum[0] (forward) and/or
// writes dlogits = gscale * d(loss)/d(logits) as bf16 (backward).
__global__ void __launch_bounds__(256) small_xent_kernel(const float* logits, int64_t ld, int n, int C, int mode,
                                                         const int64_t* labels, const float* soft, int64_t ld_soft,
                                                         float factor, float* loss_accum, bf16* dlogits, int64_t ld_d,
                                                         const float* upstream) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* x = logits + (int64_t)row * ld;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(x[c] - mx);
  s = warp_sum(s);
  const float lse = mx + logf(s);
  float loss = 0.f, tsum = 1.f;
  if (mode == 0) {
    const int64_t lab = labels[row];
    if (lane == 0) loss = lse - x[lab];
  } else {
    const float* t = soft + (int64_t)row * ld_soft;
    float a = 0.f, ts = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float tv = t[c];
      ts += tv;
      if (tv > 0.f) a += tv * (logf(tv) - (x[c] - lse));
    }
    loss = warp_sum(a);
    tsum = warp_sum(ts);
  }
  if (loss_accum && lane == 0) atomicAdd(loss_accum, factor * loss / n);
  if (dlogits) {
    const float gs = (upstream ? *upstream : 1.f) * factor / n;
    bf16* dr = dlogits + (int64_t)row * ld_d;
    for (int c = lane; c < ld_d; c += 32) {
      float g = 0.f;
      if (c < C) {
        const float pr = __expf(x[c] - lse);
        if (mode == 0) g = pr - (c == (int)labels[row] ? 1.f : 0.f);
        else g = pr * tsum - soft[(int64_t)row * ld_soft + c];
      }
      dr[c] = __float2bfloat16(g * gs);
    }
  }
}

// ------------------------------------------------------------------ AdamW (HF-3.0.2 semantics)
struct AdamHyper {
  float lr, beta1, beta2, omb1, omb2, eps, weight_decay, lr_wd;  // omb = 1 - beta, rounded from double like the reference
  const int* step;          // device {int step; float step_size}: written by step_incr_kernel for this step
  const float* inv_scale;   // optional device scalar multiplying the gradients (GradScaler), or null
};

__device__ __forceinline__ void adam_update4(float4& p, float4 g, float4& m, float4& v, const AdamHyper& h,
                                             float step_size, float gs) {
  float* pp = &p.x; float* gg = &g.x; float* mm = &m.x; float* vv = &v.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float gi = gg[i] * gs;
    mm[i] = h.beta1 * mm[i] + h.omb1 * gi;
    vv[i] = h.beta2 * vv[i] + h.omb2 * gi * gi;
    pp[i] -= step_size * mm[i] / (sqrtf(vv[i]) + h.eps);
    if (h.weight_decay > 0.f) pp[i] -= h.lr_wd * pp[i];
  }
}

__device__ __forceinline__ float adam_step_size(const AdamHyper& h) { return __int_as_float(h.step[1]); }

// step += 1; step_size = lr * sqrt(1 - beta2^t) / (1 - beta1^t) in double, like the Python reference
__global__ void step_incr_kernel(int* step, double lr, double beta1, double beta2, int correct_bias) {
  const int t = *step + 1;
  *step = t;
  double ss = lr;
  if (correct_bias) ss = lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t));
  step[1] = __float_as_int((float)ss);
}

struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  bf16* p16;
  int64_t n;
};

constexpr int ADAM_CHUNK = 8192;  // elements per block

// chunk_map[i] = (tensor index, chunk index within tensor)
__global__ void __launch_bounds__(256) adamw_multi_kernel(const AdamTensor* table, const int2* chunk_map, AdamHyper h) {
  pdl_trigger();
  pdl_wait();
  const int2 cm = chunk_map[blockIdx.x];
  const AdamTensor t = table[cm.x];
  const int64_t base = (int64_t)cm.y * ADAM_CHUNK;
  const float ss = adam_step_size(h);
  const float gs = h.inv_scale ? *h.inv_scale : 1.f;
  const bool aligned = (((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0 &&
                       (!t.p16 || ((uintptr_t)t.p16 & 7) == 0);
  for (int i = threadIdx.x * 4; i < ADAM_CHUNK; i += blockDim.x * 4) {
    const int64_t e = base + i;
    if (e >= t.n) break;
    if (aligned && e + 4 <= t.n) {
      float4 p = *reinterpret_cast<float4*>(t.p + e);
      const float4 g = *reinterpret_cast<const float4*>(t.g + e);
      float4 m = *reinterpret_cast<float4*>(t.m + e);
      float4 v = *reinterpret_cast<float4*>(t.v + e);
      adam_update4(p, g, m, v, h, ss, gs);
      *reinterpret_cast<float4*>(t.p + e) = p;
      *reinterpret_cast<float4*>(t.m + e) = m;
      *reinterpret_cast<float4*>(t.v + e) = v;
      if (t.p16) {
        __nv_bfloat162 a = __floats2bfloat162_rn(p.x, p.y), b = __floats2bfloat162_rn(p.z, p.w);
        *reinterpret_cast<uint2*>(t.p16 + e) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
      }
    } else {
      for (int j = 0; j < 4 && e + j < t.n; ++j) {
        float4 p = make_float4(t.p[e + j], 0, 0, 0), g = make_float4(t.g[e + j], 0, 0, 0);
        float4 m = make_float4(t.m[e + j], 0, 0, 0), v = make_float4(t.v[e + j], 0, 0, 0);
        adam_update4(p, g, m, v, h, ss, gs);
        t.p[e + j] = p.x; t.m[e + j] = m.x; t.v[e + j] = v.x;
        if (t.p16) t.p16[e + j] = __float2bfloat16(p.x);
      }
    }
  }
}

// ------------------------------------------------------------------ casts / repacks / masks
__global__ void cast_bf16_kernel(const float* src, bf16* dst, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + i) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  } else {
    for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
  }
}

// W_img fp32 [d, 2052] -> W_feat bf16 [d, 2048] (TMA-legal pitch) + W_box fp32 [d, 4]
__global__ void repack_img_weight_kernel(const float* w, bf16* w_feat, float* w_box, int d, int fin) {
  const int r = blockIdx.x;
  const int feat = fin - 4;
  for (int c = threadIdx.x; c < fin; c += blockDim.x) {
    const float v = w[(int64_t)r * fin + c];
    if (c < feat) w_feat[(int64_t)r * feat + c] = __float2bfloat16(v);
    else w_box[r * 4 + (c - feat)] = v;
  }
}

// attention_mask (1 = keep, int64) -> key padding bytes (1 = pad); HF-3.0.2 invert_mask
__global__ void invert_mask_kernel(const int64_t* mask, uint8_t* pad, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pad[i] = mask[i] == 0;
}

// split an fp32 matrix [rows, K] into the 3xTF32 operand [rows, 3K]:
// side 0 (activations): [hi | hi | lo]   side 1 (weights): [hi | lo | hi]
// so that A'.B'^T = hi.hi + hi.lo + lo.hi  (fp32-accurate product on tf32 tensor cores)
__global__ void split_tf32_kernel(const float* src, int64_t ld_src, float* dst, int rows, int K, int side) {
  const int r = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows || c >= K) return;
  const float x = src[(int64_t)r * ld_src + c];
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  const float lo = x - hi;
  float* d = dst + (int64_t)r * 3 * K;
  d[c] = hi;
  d[K + c] = side == 0 ? hi : lo;
  d[2 * K + c] = side == 0 ? lo : hi;
}

}  // namespace kmb

using namespace kmb;

extern "C" int kmb_ce_combine(const float* ce_max, const float* ce_sum, const float* label_logit, const int64_t* labels,
                              int M, int n_tiles, float* lse, float* row_loss, float* acc2, float factor,
                              float* loss_out, float* loss_total, int add_total, kmb_stream_t stream) {
  if (!ce_max || !ce_sum || !label_logit || !labels || !lse || !acc2 || !loss_out || M <= 0 || n_tiles <= 0) {
    kmb_set_last_error("kmb_ce_combine: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(acc2, 0, 2 * sizeof(float), st);
  launch_pdl(ce_combine_kernel, dim3((M * 32 + 255) / 256), dim3(256), 0, st, ce_max, ce_sum, label_logit, labels, M, n_tiles, lse, row_loss, acc2);
  KMB_CHECK_LAUNCH();
  ce_finalize_kernel<<<1, 1, 0, st>>>(acc2, factor, loss_out, loss_total, add_total);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_ce_gscale(const float* acc2, const float* upstream, float factor, float* gscale, kmb_stream_t stream) {
  if (!acc2 || !gscale) {
    kmb_set_last_error("kmb_ce_gscale: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  ce_gscale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(acc2, upstream, factor, gscale);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_small_xent(const float* logits, int64_t ld, int n, int C, int mode, const int64_t* labels,
                              const float* soft, int64_t ld_soft, float factor, float* loss_accum, void* dlogits_bf16,
                              int64_t ld_d, const float* upstream, kmb_stream_t stream) {
  if (!logits || n < 0 || C <= 0 || (mode == 0 && !labels) || (mode == 1 && !soft)) {
    kmb_set_last_error("kmb_small_xent: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if (n == 0) return KMB_OK;
  small_xent_kernel<<<(n * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(logits, ld, n, C, mode, labels, soft, ld_soft, factor,
                                                                         loss_accum, (bf16*)dlogits_bf16, ld_d, upstream);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_adamw_chunk_elems(void) { return ADAM_CHUNK; }

extern "C" int kmb_adamw_multi_part(const void* table_dev, const void* chunk_map_dev, int n_chunks, int* step_dev, double lr,
                                    double beta1, double beta2, double eps, double weight_decay, int correct_bias,
                                    const float* inv_scale_dev, int advance_step, kmb_stream_t stream) {
  if (!table_dev || !chunk_map_dev || n_chunks <= 0 || !step_dev) {
    kmb_set_last_error("kmb_adamw_multi: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (advance_step) step_incr_kernel<<<1, 1, 0, st>>>(step_dev, lr, beta1, beta2, correct_bias);
  AdamHyper h;
  h.lr = (float)lr; h.beta1 = (float)beta1; h.beta2 = (float)beta2; h.omb1 = (float)(1.0 - beta1); h.omb2 = (float)(1.0 - beta2);
  h.eps = (float)eps; h.weight_decay = (float)weight_decay; h.lr_wd = (float)(lr * weight_decay);
  h.step = step_dev; h.inv_scale = inv_scale_dev;
  launch_pdl(adamw_multi_kernel, dim3(n_chunks), dim3(256), 0, st, (const AdamTensor*)table_dev, (const int2*)chunk_map_dev, h);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_adamw_multi(const void* table_dev, const void* chunk_map_dev, int n_chunks, int* step_dev, double lr,
                               double beta1, double beta2, double eps, double weight_decay, int correct_bias,
                               const float* inv_scale_dev, kmb_stream_t stream) {
  return kmb_adamw_multi_part(table_dev, chunk_map_dev, n_chunks, step_dev, lr, beta1, beta2, eps, weight_decay, correct_bias,
                              inv_scale_dev, 1, stream);
}

extern "C" int kmb_cast_bf16(const float* src, void* dst, int64_t n, kmb_stream_t stream) {
  if (!src || !dst || n <= 0 || ((uintptr_t)src & 15) || ((uintptr_t)dst & 7)) {
    kmb_set_last_error("kmb_cast_bf16: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const int64_t groups = (n + 3) / 4;
  cast_bf16_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_repack_img_weight(const float* w, void* w_feat_bf16, float* w_box, int d, int fin, kmb_stream_t stream) {
  if (!w || !w_feat_bf16 || !w_box || d <= 0 || fin <= 4) {
    kmb_set_last_error("kmb_repack_img_weight: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  repack_img_weight_kernel<<<d, 256, 0, (cudaStream_t)stream>>>(w, (bf16*)w_feat_bf16, w_box, d, fin);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_invert_mask(const int64_t* mask, uint8_t* pad, int64_t n, kmb_stream_t stream) {
  if (!mask || !pad || n <= 0) {
    kmb_set_last_error("kmb_invert_mask: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  invert_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mask, pad, n);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_split_tf32(const float* src, int64_t ld_src, float* dst, int rows, int K, int side, kmb_stream_t stream) {
  if (!src || !dst || rows <= 0 || K <= 0) {
    kmb_set_last_error("kmb_split_tf32: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  split_tf32_kernel<<<dim3((K + 255) / 256, rows), 256, 0, (cudaStream_t)stream>>>(src, ld_src, dst, rows, K, side);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

// device-side dropout seed stream: state = mix64(state + golden); *out = state (graph-replay safe)
__global__ void next_seed_kernel(uint64_t* state, uint64_t* out) {
  const uint64_t s = kmb::mix64(*state + 0x9E3779B97F4A7C15ULL);
  *state = s;
  *out = s;
}

extern "C" int kmb_next_seed(uint64_t* state, uint64_t* out, kmb_stream_t stream) {
  if (!state || !out) {
    kmb_set_last_error("kmb_next_seed: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  next_seed_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, out);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}
