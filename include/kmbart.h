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
 *                  consumed as tf32.  Accumulation is fp32 in tensor memory.
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

int kmb_gemm(const void* A, const void* B, int M, int N, int K, int64_t lda, int64_t ldb,
             int a_mn, int b_mn, int elt, const KmbGemmEpilogue* epi, int tile_n,
             kmb_stream_t stream);
/* number of n-tiles kmb_gemm will use for (N, tile_n) — sizes ce_max / ce_sum */
int kmb_gemm_n_tiles(int N, int tile_n);
int kmb_gemm_pick_tile_n(int M, int N);

#ifdef __cplusplus
}
#endif
#endif /* KMBART_H_ */
