/* kmbart.h — C-ABI of libkmbart_sm100.so, the B200-native kernel library behind the
 * KM-BART `src.model` API.
 *
 * The reference (fomalhautb/KM-BART) has no FFI layer: its "operator API" is the
 * Python classes in src/model/{model,modules,mixins}.py, whose arithmetic is issued
 * as PyTorch library calls (SURVEY.md §2.3(b), rows K1–K14).  Each entry point below
 * replaces one of those call sites; the citation after "replaces:" is the reference
 * file:line (paths relative to the reference root; "HF-3.0.2" = the un-vendored
 * transformers==3.0.2 dependency, environment.yaml:159).
 *
 * Conventions: plain pointers and sizes, no torch types; every function enqueues on
 * `stream` and never synchronises the device or allocates; return 0 on success or a
 * negative KMB_ERR_* code (kmb_last_error() gives the text).  All device pointers
 * must be 16-byte aligned.  Activations are row-major [tokens, features] with the
 * batch-major token order (b * S + s).
 */
#ifndef KMBART_H_
#define KMBART_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* kmb_stream_t; /* cudaStream_t */

#define KMB_OK 0
#define KMB_ERR_ARG (-1)
#define KMB_ERR_CUDA (-2)
#define KMB_ERR_ARCH (-3)
#define KMB_ERR_TMAP (-4)

/* library / device checks ------------------------------------------------------ */
int kmb_version(void);
/* 0 iff the current device is sm_100 (B200); KMB_ERR_ARCH otherwise.  There is no
 * CPU or other-arch fallback. */
int kmb_arch_check(void);
const char* kmb_last_error(void);

/* ------------------------------------------------------------------------------
 * tcgen05/TMEM GEMM:  D[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
 * replaces: every nn.Linear / F.linear on the path — q/k/v/out_proj, fc1, fc2
 *   (HF-3.0.2 EncoderLayer/DecoderLayer/SelfAttention, instantiated at
 *   src/model/modules.py:84 and src/model/model.py:35), ImageEmbedding.linear
 *   (src/model/modules.py:30-31), the LM head F.linear (src/model/model.py:397,
 *   :291), BartClassificationHead (src/model/model.py:133-158), and their autograd
 *   backward forms (dgrad / wgrad).
 *
 * Operand storage: a_mn = 0 -> A is stored [M, lda] with k contiguous ("K-major");
 *                  a_mn = 1 -> A is stored [K, lda] with m contiguous ("MN-major").
 *                  Same for B with n.  elt = 0: bf16 operands; elt = 1: fp32 operands
 *                  consumed as tf32 (K-major only; with kmb_split_tf32 operands this is the
 *                  3xTF32 fp32-parity mode).  Accumulation is fp32 in tensor memory.
 * Epilogue modes:
 *   KMB_EPI_LINEAR   v = alpha*acc (+bias[n]) ; act ; dropout ; (+residual) ; (+= out_f32)
 *                    then written to out_f32 and/or out_bf16.
 *   KMB_EPI_CE_STATS per (row, n-tile) online-softmax partial (max, sumexp) and the
 *                    label logit — the forward half of the fused LM-head +
 *                    CrossEntropyLoss (src/model/model.py:397-402); logits are never
 *                    written to HBM.
 *   KMB_EPI_CE_GRAD  dlogits = (exp(v - lse[row]) - [n == label]) * gscale, written
 *                    as bf16 — the backward half of the same loss.
 */
enum { KMB_EPI_LINEAR = 0, KMB_EPI_CE_STATS = 1, KMB_EPI_CE_GRAD = 2 };
enum { KMB_ACT_NONE = 0, KMB_ACT_GELU = 1, KMB_ACT_GELU_GRAD = 2, KMB_ACT_TANH = 3,
       KMB_ACT_TANH_GRAD = 4 };

typedef struct KmbGemmEpilogue {
  int32_t mode;          /* KMB_EPI_* */
  int32_t act;           /* KMB_ACT_* */
  float alpha;           /* scales the accumulator */
  int32_t accumulate;    /* 1: out_f32 += result */
  const float* bias;     /* [N] fp32 or NULL */
  const float* residual; /* fp32 [M, ld_res] or NULL (added after act+dropout) */
  int64_t ld_res;
  const void* aux;       /* bf16 [M, ld_aux]: GELU_GRAD/TANH_GRAD input (pre-act / act) */
  int64_t ld_aux;
  float* out_f32;        /* fp32 [M, ld_f32] or NULL */
  int64_t ld_f32;
  void* out_bf16;        /* bf16 [M, ld_bf16] or NULL */
  int64_t ld_bf16;
  void* out_preact;      /* bf16 [M, ld_bf16]: pre-activation copy (ACT_GELU) or NULL */
  /* dropout on the activation output, mask = f(seed, tag, m*N+n) */
  float dropout_p;
  uint32_t dropout_tag;
  const uint64_t* dropout_seed; /* device pointer (graph-replay safe) or NULL */
  /* cross-entropy modes */
  const int64_t* labels; /* [M] (-100 = ignore) */
  float* ce_max;         /* [M, n_tiles] */
  float* ce_sum;         /* [M, n_tiles] */
  float* ce_label_logit; /* [M] */
  const float* ce_lse;   /* [M] */
  const float* ce_gscale;/* device scalar: upstream_grad / n_valid */
} KmbGemmEpilogue;

/* tile_n: 0 = choose from (M, N) with the ingest/wave cost model; 32/64/128/256 = single-CTA 128 x tile_n
 * tiles (tcgen05 cta_group::1); 1128/1192/1256 = CTA-pair 256 x (tile_n - 1000) tiles (cta_group::2,
 * bf16 only; 192 needs a K-major B). */
int kmb_gemm(const void* A, const void* B, int M, int N, int K, int64_t lda, int64_t ldb,
             int a_mn, int b_mn, int elt, const KmbGemmEpilogue* epi, int tile_n,
             kmb_stream_t stream);
/* number of n-tiles kmb_gemm will use for (N, tile_n) — sizes ce_max / ce_sum */
int kmb_gemm_n_tiles(int N, int tile_n);
int kmb_gemm_pick_tile_n(int M, int N);
/* debug aid: enable the block-0 in-kernel timeline (globaltimer ns at entry / setup done / first operands landed /
 * first accumulator complete / first epilogue done / last epilogue done / exit) and read it back */
int kmb_gemm_debug_timeline(int enable, int pair, unsigned long long* out7);

/* ------------------------------------------------------------------------------
 * Fused multi-head attention, head_dim = 64, bf16 in/out, fp32 online softmax.
 * replaces: HF-3.0.2 SelfAttention.forward — bmm(q,k^T), additive causal mask
 *   (src/model/model.py:63-70), key-padding masked_fill (src/model/modules.py:130-131),
 *   softmax, bmm(p,v) — instantiated at src/model/modules.py:84 and src/model/model.py:35.
 * q/k/v/o are token-major with element row strides ld*; head h lives in columns
 * [64h, 64h+64).  key_pad: [B, Sk] bytes, 1 = padding (or NULL).  lse: [B, H, Sq].
 * `scale` multiplies q.k (dh^-0.5; the reference scales q instead, same product).
 */
int kmb_attn_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                 void* o, int64_t ldo, float* lse, const uint8_t* key_pad, int B, int H, int Sq, int Sk,
                 int head_dim, int causal, float scale, kmb_stream_t stream);
/* Same kernel with explicit element strides {batch, head, row} for q, k, v, o (12 values), so
 * the legacy KV-cache layout [B, H, T, 64] of HF-3.0.2 layer_state (prev_key / prev_value,
 * reordered by src/model/mixins.py:419-434) can be consumed in place. */
int kmb_attn_fwd_strided(const void* q, const void* k, const void* v, void* o, const int64_t* strides12,
                         const uint8_t* key_pad, int B, int H, int Sq, int Sk, int head_dim, int causal,
                         float scale, kmb_stream_t stream);
/* Decode-step attention: one query token per row over a preallocated, never-moved cache.
 * replaces: the cached branch of HF-3.0.2 SelfAttention.forward (torch.cat growth of prev_key / prev_value) and the
 *   per-step index_select of every cached tensor in _reorder_cache (src/model/mixins.py:419-434).
 * K/V of (slot, pos, head) live at k + slot*kv_slot_stride + pos*kv_pos_stride + head*64 (elements).  Row j reads
 * position p from slot slot_tbl[j*tbl_ld + p] (beam ancestry table) or, when slot_tbl is NULL, from slot j / row_div
 * (cross-attention K/V stored once per sample, shared by its beams).  key_pad: [n_samples, pad_ld] bytes (1 = pad),
 * indexed by j / row_div, or NULL.  T <= 1024 keys. */
int kmb_decode_attn(const void* q, int64_t q_row_stride, const void* k, const void* v, int64_t kv_slot_stride,
                    int64_t kv_pos_stride, const int* slot_tbl, int64_t tbl_ld, int row_div, const uint8_t* key_pad,
                    int64_t pad_ld, void* o, int64_t o_row_stride, int rows, int H, int T, int head_dim, float scale,
                    kmb_stream_t stream);
/* fp32 parity mode attention (north_star "fp32 mode": logits within 1e-4 relative, greedy decode token-identical):
 * the decode-attention kernel on fp32 q / k / v / o.  Row r belongs to slot r / row_div; with causal_mod = S_q the
 * rows of a slot are its S_q query positions and row r sees keys <= r % S_q, which makes this the full-sequence
 * (self, causal or not, and cross) attention of HF-3.0.2 SelfAttention.forward in fp32.  T <= 1024 keys. */
int kmb_attn_f32(const float* q, int64_t q_row_stride, const float* k, const float* v, int64_t kv_slot_stride,
                 int64_t kv_pos_stride, int row_div, const uint8_t* key_pad, int64_t pad_ld, float* o,
                 int64_t o_row_stride, int rows, int H, int T, int head_dim, int causal_mod, float scale,
                 kmb_stream_t stream);
/* Greedy token selection + finished-sentence bookkeeping of one decode step, on device.
 * replaces: HF-3.0.2 _generate_no_beam_search loop body reached from src/model/mixins.py:368-382 (EOS ban below
 *   min_length, argmax, pad for finished rows, append, sent_lengths / unfinished_sents update).
 * eos_token_id < 0 = no EOS handling.  out_tokens [rows, out_ld] receives the token at column cur_len. */
int kmb_greedy_select(const float* logits, int64_t ld, int rows, int V, int eos_token_id, int pad_token_id,
                      int ban_eos, int cur_len, int64_t* unfinished, int64_t* sent_len, int64_t* out_tokens,
                      int64_t out_ld, int64_t* ids_next, kmb_stream_t stream);
/* Sampling step of the no-beam loop, on device: temperature, top-k filter (0 = none; ties with the k-th value survive),
 * nucleus (top-p) filter on the renormalised softmax (1.0 = none; keep a token iff the mass sorted strictly before it is <=
 * top_p), one multinomial draw per row from hash(seed[0], row, cur_len), then the same bookkeeping as
 * kmb_greedy_select.  One CTA per row holds the logits row in shared memory (V <= kmb_select_max_vocab()).
 * replaces: HF-3.0.2 _generate_no_beam_search do_sample branch (top_k_top_p_filtering + softmax + multinomial) reached
 *   from src/model/mixins.py:368-382 with the vcg_generate.py defaults (src/generation.py:22-32). */
int kmb_select_max_vocab(void);
int kmb_sample_select(const float* logits, int64_t ld, int rows, int V, float temperature, int top_k, float top_p, int eos_token_id,
                      int pad_token_id, int ban_eos, int cur_len, const uint64_t* seed, int64_t* unfinished,
                      int64_t* sent_len, int64_t* out_tokens, int64_t out_ld, int64_t* ids_next, kmb_stream_t stream);
/* Beam-search step on device (num_beams > 1, do_sample = False): log_softmax of the (optionally forced) logits, EOS ban
 * below min_length, + beam_scores, top 2 * num_beams over beams x vocab per batch element, then the HF-3.0.2 candidate
 * loop: EOS candidates ranked < num_beams become finished hypotheses (BeamHypotheses.add with length_penalty, worst-score
 * eviction), the first num_beams others become the next beams, is_done / early_stopping; beams are re-ordered by
 * permuting the token history and the cache ancestry table (slot_tbl, may be NULL), never the cache.
 * replaces: the body of HF-3.0.2 _generate_beam_search reached from src/model/mixins.py:336-366, including
 *   adjust_logits_during_generation / _force_token_ids_generation (src/model/mixins.py:400-417: force_token >= 0) and
 *   _reorder_cache (src/model/mixins.py:419-434).  Hypotheses are read back and finalised on the host once, at the end.
 * State (all device memory, rows = batch * num_beams): cand_* [rows, K] scratch; beam_scores [rows]; hist [rows, max_len]
 * int32 tokens so far (column 0 = decoder start); ids_next [rows] int64 token fed to the next model step; beam_idx [rows]
 * the parent row of every new beam (for callers that re-order a legacy cache); done / hyp_n [batch]; hyp_score
 * [batch, num_beams + 1] double; hyp_len [batch, num_beams + 1]; hyp_tok [batch, num_beams + 1, max_len]; worst [batch]
 * double (initialised to 1e9); done_count [1]. */
typedef struct KmbBeamState {
  int32_t batch, num_beams, K, V, eos, pad, max_len, early_stopping;
  double length_penalty;
  float* cand_val; int32_t* cand_tok;
  float* beam_scores; int32_t* hist; int32_t* slot_tbl; int64_t* ids_next; int32_t* beam_idx;
  int32_t* done; int32_t* hyp_n; double* hyp_score; int32_t* hyp_len; int32_t* hyp_tok; double* worst; int32_t* done_count;
} KmbBeamState;
int kmb_beam_step(const float* logits, int64_t ld, const KmbBeamState* state, int cur_len, int force_token, int ban_eos,
                  kmb_stream_t stream);

/* ------------------------------------------------------------------------------
 * Persistent decode step, cluster variant (opt-in, KMBART_DECODE_CLUSTER=1): the whole cached decoder forward of ONE generation step in one launch of 4-CTA clusters
 * (csrc/decode_step.cu): embedding, L decoder layers (LayerNorm-on-load, q|k|v projection fused with self-attention,
 * out-proj, cross-q projection fused with cross-attention, out-proj, fc1+GELU, fc2 — six grid-barrier-separated phases
 * per layer, K split four ways inside a cluster and reduced through distributed shared memory) and the final
 * LayerNorm; each CTA prefetches its weight slices through a shared-memory ring ahead of the barriers.
 * replaces: one `self(**model_inputs)` iteration of HF-3.0.2 _generate_no_beam_search / _generate_beam_search reached
 *   from src/model/mixins.py:336-382 — BartDecoder.forward(use_cache=True) with DecoderLayer / SelfAttention cached
 *   branches (instantiated src/model/model.py:35), i.e. the 67 kernels of the launch-chain version of the same step.
 * Buffers: y0 / y1 fp32 [rows, d] hold the pre-LayerNorm residual stream (ping-pong), stats fp32
 * [1 + 3L, rows, NP, 2] the per-strip (sum, sum of squares) partials of every residual state (NP = d / strip width:
 * 32 for d = 768 / 1024, 16 for d = 128); x_f32 / x_b16 [rows, d] receive the final decoder state (LM-head operand);
 * ctx [rows, d] and h [rows, F] bf16 are scratch; barrier points at ONE 64-bit counter zeroed once per session and used
 * by every launch of that session.  Layer caches are [rows, max_len, 3d] (q|k|v per position; the step writes k|v of
 * position t); row j reads position p < t of the self cache from row slot_tbl[j * max_len + p] (beam ancestry, or NULL =
 * own row); cross K/V are [n_samples * Se, 2d], sample = row / row_div; key_pad [n_samples, Se] bytes (1 = pad) or NULL.
 * d in {128, 768, 1024}, head_dim 64, max_len and Se <= 512.
 */
#define KMB_DECODE_MAX_LAYERS 12
#define KMB_DECODE_NO_CROSS_PREFETCH 1
#define KMB_DECODE_NO_SELF_PREFETCH 2
typedef struct KmbDecodeLayer {
  const void *w_qkv, *w_o, *w_cq, *w_co, *w_fc1, *w_fc2;      /* bf16 [3d,d] [d,d] [d,d] [d,d] [F,d] [d,F] */
  const float *b_qkv, *b_o, *b_cq, *b_co, *b_fc1, *b_fc2;
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b; /* self_attn / encoder_attn / final layer norms */
  void* cache;                                                /* bf16 [rows, max_len, 3d] */
  const void* cross_kv;                                       /* bf16 [n_samples * Se, 2d] */
  const void* packed[6];                                      /* kmb_decode_pack_weights images of w_qkv, w_o, w_cq, w_co, w_fc1, w_fc2 */
} KmbDecodeLayer;
typedef struct KmbDecodeStepC {
  int32_t rows, d, H, F, L, t, max_len, Se, row_div, pos_row; /* pos_row = t + position offset (2) */
  float embed_scale, attn_scale;
  int32_t flags, reserved;                                    /* tuning switches (KMB_DECODE_*), 0 = defaults */
  const int64_t* ids;                                         /* [rows] token fed to this step */
  const float *tok_emb, *pos_emb, *lne_g, *lne_b;             /* fp32 [V,d], [npos,d], layernorm_embedding */
  const int32_t* slot_tbl;
  const uint8_t* key_pad;
  float *y0, *y1, *stats;
  float* x_f32; void* x_b16; void* ctx; void* h;
  unsigned long long* barrier;
  unsigned long long* trace;   /* NULL, or [n_barriers][grid][2] globaltimer ns: {arrive, release} of every grid barrier (profiling aid) */
  KmbDecodeLayer layers[KMB_DECODE_MAX_LAYERS];
} KmbDecodeStepC;
int kmb_decode_step_cluster(const KmbDecodeStepC* step, kmb_stream_t stream);
/* Weight images for the step's shared-memory ring: every (column strip, K quarter, K chunk) slice of a weight matrix
 * stored contiguously with the ring's row pitch, so a ring slot is filled by ONE bulk copy.  Pack once per weight
 * version; kmb_decode_pack_offsets gives the byte offsets of the six sections inside one layer's image (off7[6] = bytes
 * per layer), kmb_decode_pack_weights fills one layer's image from layer->w_*. */
int kmb_decode_pack_offsets(int d, int H, int F, int64_t* off7);
int kmb_decode_pack_weights(const KmbDecodeLayer* layer, int d, int H, int F, void* out, kmb_stream_t stream);
/* CTAs of the persistent grid for model width d (4 x the number of co-resident clusters, at most 128) */
int kmb_decode_cluster_grid(int d);
/* grid barriers of one step (first dimension of the trace buffer) */
int kmb_decode_cluster_barriers(int n_layers);

/* ------------------------------------------------------------------------------
 * Persistent decode step: the whole cached decoder forward of ONE generation step in one cooperative launch
 * (embedding + LayerNorm, L decoder layers with self / cross attention over the preallocated caches, FFN), grid-wide
 * barriers between the dependent sub-steps and a shared-memory weight ring that prefetches each CTA's weight slices
 * ahead of the barriers.
 * replaces: one `self(**model_inputs)` iteration of HF-3.0.2 _generate_no_beam_search / _generate_beam_search reached
 *   from src/model/mixins.py:336-382 — BartDecoder.forward(use_cache=True) with DecoderLayer / SelfAttention cached
 *   branches (instantiated src/model/model.py:35), i.e. the 67 kernels of the launch-chain version of the same step.
 * Buffers: x_f32/x_b16 [rows, d] receive the final decoder state (LM-head operand); ctx, q2 [rows, d] bf16, lin
 * [rows, d] fp32 (MUST be zero on entry the first time; the kernel leaves it consumed-and-cleared as it needs),
 * h [rows, F] bf16 are scratch; barrier points at ONE 64-bit counter zeroed once per session and used by every launch
 * of that session.  Layer caches are [rows, max_len, 3d] (q|k|v per position); row j reads position p of the self
 * cache from row slot_tbl[j * max_len + p] (beam ancestry, or NULL = own row); cross K/V are [n_samples * Se, 2d],
 * sample = row / row_div; key_pad [n_samples, Se] bytes (1 = pad) or NULL.  d = 768, head_dim 64, max_len and Se <= 512.
 */
typedef struct KmbDecodeStep {
  int32_t rows, d, H, F, L, t, max_len, Se, row_div, pos_row; /* pos_row = t + position offset (2) */
  float embed_scale, attn_scale;
  int32_t nt[6];                                              /* filled by the library */
  const int64_t* ids;                                         /* [rows] token fed to this step */
  const float *tok_emb, *pos_emb, *lne_g, *lne_b;             /* fp32 [V,d], [npos,d], layernorm_embedding */
  const int32_t* slot_tbl;
  const uint8_t* key_pad;
  float* x_f32; void* x_b16; void* ctx; float* lin; void* q2; void* h;
  unsigned long long* barrier;
  unsigned long long* trace;   /* NULL, or [n_barriers][grid][2] globaltimer ns: {arrive, release} of every grid barrier (profiling aid) */
  KmbDecodeLayer layers[KMB_DECODE_MAX_LAYERS];
} KmbDecodeStep;
int kmb_decode_step(const KmbDecodeStep* step, kmb_stream_t stream);
/* CTAs of the persistent grid (= SMs of the current device) */
int kmb_decode_step_grid(void);

/* autograd backward of kmb_attn_fwd; d_scratch: [B, H, Sq] floats. */
int kmb_attn_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                 const void* o, int64_t ldo, const void* d_o, int64_t lddo, const float* lse,
                 float* d_scratch, const uint8_t* key_pad, void* dq, void* dk, void* dv, int64_t lddq,
                 int64_t lddk, int64_t lddv, int B, int H, int Sq, int Sk, int head_dim, int causal,
                 float scale, kmb_stream_t stream);

/* ------------------------------------------------------------------------------
 * Visual-token embedding path.
 * replaces: ImageEmbedding.forward (src/model/modules.py:24-41: torch.cat of the ragged
 *   list + Linear(2052->d) + Python split) and _embed_multi_modal (src/model/modules.py:89-102).
 * kmb_pack_features: list of B fp32 [n_i, 2052] tensors (device pointer table + row offsets
 *   [B+1]) or one packed [R, 2052] buffer -> bf16 RoI features [R, 2048] + fp32 boxes [R, 4].
 * kmb_slot_index: slot_idx[b, s] = packed RoI row that overwrites token (b, s), else -1.
 */
int kmb_pack_features(const float* const* feat_ptrs, const int* row_offsets, int B, const float* packed,
                      void* feats_bf16, float* boxes, int R, kmb_stream_t stream);
int kmb_slot_index(const int64_t* input_ids, const int* row_offsets, int B, int S, int img_feat_id,
                   int cls_token_id, int* slot_idx, kmb_stream_t stream);
/* (token gather | visual GEMM row + bias + box projection) * embed_scale + learned position
 * (offset 2) -> LayerNorm(eps 1e-5) -> dropout.
 * replaces: src/model/modules.py:93-100, :133-137 (encoder) and HF-3.0.2 BartDecoder.forward
 *   embed+pos+LN (decoder; pass slot_idx = NULL).  pos_index != NULL: every row uses position
 *   *pos_index (cached decode, LearnedPositionalEmbedding(use_cache=True)). */
int kmb_embed_ln_fwd(const int64_t* ids, const int* slot_idx, const float* tok_emb, const float* pos_emb,
                     const float* vis_acc, const float* boxes, const float* w_box, const float* b_img,
                     const float* gamma, const float* beta, float* pre, float* out_f32, void* out_bf16,
                     float* mean, float* rstd, int M, int S, int d, int pos_offset, const int* pos_index,
                     float embed_scale, float dropout_p, uint32_t dropout_tag, const uint64_t* dropout_seed,
                     kmb_stream_t stream);
int kmb_embed_bwd(const float* demb, const int64_t* ids, const int* slot_idx, float* d_tok, void* dvis_bf16,
                  float* dpos, int B, int S, int d, int pos_offset, int pad_id, float embed_scale,
                  int accumulate_pos, kmb_stream_t stream);
int kmb_box_wgrad(const void* dvis_bf16, const float* boxes, float* dw_img, int R, int d, int ld_w,
                  kmb_stream_t stream);

/* ------------------------------------------------------------------------------
 * LayerNorm (eps 1e-5) on the fp32 residual stream.
 * replaces: HF-3.0.2 LayerNorm = torch.nn.LayerNorm in EncoderLayer/DecoderLayer (post-LN,
 *   normalize_before=False) and layernorm_embedding (src/model/modules.py:85,136).
 * fwd fuses the residual add and the dropout that sit between a sublayer's last Linear and its
 * LayerNorm: pre = residual + dropout(z), y = LN(pre); z = NULL gives a plain LN of `residual`.
 * bwd takes the incoming gradient as an fp32 part (residual path) plus an optional bf16 part
 * (dgrad GEMM output), produces dpre, dz = dropout-re-masked bf16 copy of dpre (the dY operand
 * of the Linear that fed the residual add) and accumulates dgamma/dbeta/dbias column sums.
 */
int kmb_layernorm_fwd(const void* z_bf16, const float* residual, const float* gamma, const float* beta,
                      float* pre, float* out_f32, void* out_bf16, float* mean, float* rstd, int M, int d,
                      float dropout_p, uint32_t dropout_tag, const uint64_t* dropout_seed, kmb_stream_t stream);
int kmb_layernorm_bwd(const float* dy, const void* dy_bf16, const float* pre, const float* mean, const float* rstd,
                      const float* gamma, float* dpre, void* dz_bf16, float* dgamma, float* dbeta,
                      float* dbias, int M, int d, float drop_in_p, uint32_t drop_in_tag, float drop_out_p,
                      uint32_t drop_out_tag, const uint64_t* dropout_seed, kmb_stream_t stream);
int kmb_colsum_bf16(const void* x, int64_t ld, float* out, int M, int N, kmb_stream_t stream);
int kmb_gather_rows_bf16(const void* src, int64_t ld_src, const int* idx, void* out, int64_t ld_out, int n,
                         int d, kmb_stream_t stream);
int kmb_scatter_add_rows(const void* src_bf16, int64_t ld_src, const int* idx, float* dst, int64_t ld_dst,
                         int n, int d, kmb_stream_t stream);

/* ------------------------------------------------------------------------------
 * Losses.
 * kmb_ce_combine: reduces the KMB_EPI_CE_STATS partials to per-row log-sum-exp and the mean
 *   cross-entropy over labels != -100 — nn.CrossEntropyLoss() at src/model/model.py:401-402
 *   (:299-301 with lm_loss_factor).  acc2 = {sum, count} scratch; loss_total (+)= loss.
 * kmb_ce_gscale: gscale = upstream * factor / count for the KMB_EPI_CE_GRAD epilogue.
 * kmb_small_xent: pretraining heads — mode 0 CE mean (src/model/model.py:264-266, :285-287),
 *   mode 1 KL-div batchmean on log_softmax (src/model/model.py:253-255).
 */
int kmb_ce_combine(const float* ce_max, const float* ce_sum, const float* label_logit, const int64_t* labels,
                   int M, int n_tiles, float* lse, float* row_loss, float* acc2, float factor,
                   float* loss_out, float* loss_total, int add_total, kmb_stream_t stream);
int kmb_ce_gscale(const float* acc2, const float* upstream, float factor, float* gscale, kmb_stream_t stream);
int kmb_small_xent(const float* logits, int64_t ld, int n, int C, int mode, const int64_t* labels,
                   const float* soft, int64_t ld_soft, float factor, float* loss_accum, void* dlogits_bf16,
                   int64_t ld_d, const float* upstream, kmb_stream_t stream);

/* ------------------------------------------------------------------------------
 * Optimizer: one launch over every parameter tensor.
 * replaces: transformers.AdamW.step (HF-3.0.2 optimization.py; constructed at
 *   vcg_train.py:100 / pretrain.py:100, stepped at src/training.py:136-143).
 * table_dev: device array of {float* p; const float* g; float* m; float* v; bf16* p16;
 * int64 n}; chunk_map_dev: device array of int2 {tensor, chunk}; chunks are
 * kmb_adamw_chunk_elems() elements.  step_dev points at TWO 32-bit words {int step; float
 * step_size}: the step is incremented and the bias-corrected step size is recomputed (in
 * double, like the Python reference) on device, so the call is graph-replay safe.  p16
 * (optional) receives the refreshed bf16 shadow weights.  Hyper-parameters are doubles
 * because the reference derives 1-beta and beta^t in Python floats.
 */
int kmb_adamw_chunk_elems(void);
int kmb_adamw_multi(const void* table_dev, const void* chunk_map_dev, int n_chunks, int* step_dev, double lr,
                    double beta1, double beta2, double eps, double weight_decay, int correct_bias,
                    const float* inv_scale_dev, kmb_stream_t stream);
/* the same update over a sub-range of the chunk map; advance_step = 0 re-uses the step / step size of the
 * preceding call of the same optimizer step (the data-parallel path updates the parameters whose gradient
 * exchange has finished while the last all-reduce is still in flight, then the rest: kmbart/optim.py) */
int kmb_adamw_multi_part(const void* table_dev, const void* chunk_map_dev, int n_chunks, int* step_dev, double lr,
                         double beta1, double beta2, double eps, double weight_decay, int correct_bias,
                         const float* inv_scale_dev, int advance_step, kmb_stream_t stream);
/* ---- data-parallel gradient exchange over NVLink peer memory (csrc/peer_exchange.cu).
 * Replaces, on one NVSwitch node, the bucketed NCCL all-reduce that torch DDP performs for the reference
 * (vcg_train.py:96-98, pretrain.py:96-98: DDP(model, find_unused_parameters=True)); kmbart/parallel.py falls back to
 * NCCL when the ranks are not all peers of each other.
 * kmb_ipc_export / kmb_ipc_open: CUDA IPC handle (64 bytes) of the allocation that contains `ptr`, plus the offset of
 * `ptr` inside it; the opener maps the allocation and adds the offset itself. */
int kmb_ipc_export(const void* ptr, unsigned char* handle64, unsigned long long* offset);
int kmb_ipc_open(const unsigned char* handle64, void** base);
int kmb_ipc_close(void* base);
int kmb_peer_can_access(int dev, int peer_dev);   /* 1 / 0 */
/* g: this rank's flat fp32 gradient buffer; staging: [2 lanes][world][slot_elems] fp32; flags: [2][n_regions][world] u32, zeroed;
 * peer_*[r]: rank r's buffers mapped into this process (entry `rank` is ignored). */
int kmb_peer_ctx_create(int rank, int world, float* g, void* const* peer_g, float* staging, void* const* peer_staging,
                        unsigned* flags, void* const* peer_flags, size_t slot_elems, int n_regions, void** ctx_out);
int kmb_peer_ctx_destroy(void* ctx);
/* g[start, end) is final on `compute`: average it over the ranks on the exchange's own streams.  mode 0: copy engines
 * + one small reduction kernel (no SM taken from a running sweep); mode 1: one kernel that loads the slice from every
 * rank and stores the average to every rank (for regions exchanged after the sweep).  `value` is the same on every
 * rank and grows with every exchange of this region; so must `mode` be. */
int kmb_peer_exchange_region(void* ctx, int region, size_t start, size_t end, unsigned value, int mode, kmb_stream_t compute);
/* orders `compute` after everything enqueued so far on the exchange's stream */
int kmb_peer_join(void* ctx, kmb_stream_t compute);
/* kmb_peer_mark remembers the current end of the exchange's streams; kmb_peer_join_mark orders `compute` after that
 * point only (regions enqueued after the mark stay in flight: the optimizer joins them later, kmbart/optim.py) */
int kmb_peer_mark(void* ctx);
int kmb_peer_join_mark(void* ctx, kmb_stream_t compute);
int kmb_cast_bf16(const float* src, void* dst, int64_t n, kmb_stream_t stream);
int kmb_repack_img_weight(const float* w, void* w_feat_bf16, float* w_box, int d, int fin, kmb_stream_t stream);
/* attention_mask (int64, 1 = keep) -> padding bytes (1 = pad): HF-3.0.2 invert_mask
 * (src/model/modules.py:130-131) */
int kmb_invert_mask(const int64_t* mask, uint8_t* pad, int64_t n, kmb_stream_t stream);
/* advances the device-resident dropout seed stream (no host RNG in the step, graph-replay safe) */
int kmb_next_seed(uint64_t* state, uint64_t* out, kmb_stream_t stream);
/* fp32 [rows, K] -> 3xTF32 operand [rows, 3K] (side 0: hi|hi|lo, side 1: hi|lo|hi) for the
 * fp32-parity mode of kmb_gemm(elt = 1) */
int kmb_split_tf32(const float* src, int64_t ld_src, float* dst, int rows, int K, int side, kmb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KMBART_H_ */
