// Host side of kmb_gemm (C-ABI, include/kmbart.h): tensor-map construction, tile-shape selection,
// split-K planning and dispatch to the gemm_tc05_kernel instantiations (device code: gemm_kernel.cuh).
#include "gemm_launch.cuh"

namespace kmb {

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// 2-D map over a row-major [rows, ld] matrix with `cols` valid columns, box = (box_cols, box_rows)
static int make_tmap(CUtensorMap* tm, const void* ptr, int elt, uint64_t rows, uint64_t cols,
                     uint64_t ld, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    kmb_set_last_error("cuTensorMapEncodeTiled entry point unavailable", __FILE__, __LINE__);
    return KMB_ERR_TMAP;
  }
  const uint64_t esz = elt == 0 ? 2 : 4;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, elt == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%ux%u",
             (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
    kmb_set_last_error(msg, __FILE__, __LINE__);
    return KMB_ERR_TMAP;
  }
  return KMB_OK;
}


template <int BN, int ELT>
static int launch_major(int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                        const CUtensorMap& tmPre, const GemmParams& p, cudaStream_t st) {
  // an MN-major B tile is built from 128-byte-wide chunks, so BN must cover one chunk
  constexpr bool kBmnOk = BN >= 64 && ELT == 0;
  if (!a_mn && !b_mn) return gemm_launch<BN, ELT, 0, 0, 1>(tmA, tmB, tmOut, tmPre, p, st);
  if constexpr (ELT == 0) {
    if (a_mn && !b_mn) return gemm_launch<BN, ELT, 1, 0, 1>(tmA, tmB, tmOut, tmPre, p, st);
    if constexpr (kBmnOk) {
      if (!a_mn && b_mn) return gemm_launch<BN, ELT, 0, 1, 1>(tmA, tmB, tmOut, tmPre, p, st);
      return gemm_launch<BN, ELT, 1, 1, 1>(tmA, tmB, tmOut, tmPre, p, st);
    }
  }
  kmb_set_last_error("kmb_gemm: tile_n too narrow for an MN-major B operand", __FILE__, __LINE__);
  return KMB_ERR_ARG;
}

// ---- tile selection -------------------------------------------------------------------------------
// Cost model from profiles/r01_*: a k-block of a tile costs max(MMA issue, operand ingest) and the
// measured limiter is ingest (~47 B/clk/SM): per CTA (128 rows of A + BN/CG rows of B) * 128 B.
// A problem costs waves * per-tile cost (+ a per-tile epilogue/drain term); pick the cheapest shape.
struct TileChoice { int cg, bn; };
static TileChoice pick_tile(int M, int N, int b_mn, int elt, bool splitk_ok) {
  const int sms = gemm_num_sms();
  TileChoice best = {1, 32};
  double best_cost = 1e30;
  auto consider = [&](int cg, int bn) {
    if (bn / 2 >= N && bn > 32) return;                 // do not pad N by more than 2x
    if (cg == 2 && (elt != 0 || M <= 128)) return;
    if (cg == 2 && b_mn && ((bn / 2) % 64)) return;
    if (cg == 1 && b_mn && bn < 64) return;
    const int mt = (M + 128 * cg - 1) / (128 * cg), nt = (N + bn - 1) / bn;
    const int groups = sms / cg;
    const long tiles = (long)mt * nt;
    const long waves = splitk_ok ? 1 : (tiles + groups - 1) / groups;   // split-K fills partial waves
    const double fill = splitk_ok ? (double)tiles / groups : (double)waves;
    const double ingest = 128.0 + (double)bn / cg;       // rows of operands per CTA per k-block
    const double mma = 0.735 * bn;                        // 2*bn clk of tcgen05.mma per k-block, in the same row units
    const double per_tile = (ingest > mma ? ingest : mma) + 24.0;
    const double cost = fill * per_tile;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = {cg, bn}; }
  };
  consider(2, 256); consider(2, 192); consider(2, 128);
  consider(1, 256); consider(1, 128); consider(1, 64); consider(1, 32);
  return best;
}

}  // namespace kmb

static int g_timeline_on = 0;
// debug: enable the block-0 timeline and read it back (ns since the kernel's first timestamp)
extern "C" int kmb_gemm_debug_timeline(int enable, int pair, unsigned long long* out7) {
  g_timeline_on = enable;
  if (out7) {
    unsigned long long h[8];
    if (pair) { if (kmb::gemm_pair_read_timeline(h)) return KMB_ERR_CUDA; }
    else if (cudaMemcpyFromSymbol(h, kmb::g_gemm_timeline, sizeof h) != cudaSuccess) return KMB_ERR_CUDA;
    for (int i = 0; i < 7; ++i) out7[i] = h[i];
  }
  return KMB_OK;
}

extern "C" int kmb_gemm_pick_tile_n(int M, int N) {
  const kmb::TileChoice c = kmb::pick_tile(M, N, 0, 0, false);
  return c.cg == 2 ? 1000 + c.bn : c.bn;
}


// number of (max, sumexp) partials per row that KMB_EPI_CE_STATS writes: EPI_SLICES_MAX per n-tile
// (one per epilogue column slice; the CE epilogues always run with 16 epilogue warps)
extern "C" int kmb_gemm_n_tiles(int N, int tile_n) {
  if (tile_n >= 1000) tile_n -= 1000;
  if (tile_n != 32 && tile_n != 64 && tile_n != 128 && tile_n != 192 && tile_n != 256) return KMB_ERR_ARG;
  return kmb::EPI_SLICES_MAX * ((N + tile_n - 1) / tile_n);
}

extern "C" int kmb_gemm(const void* A, const void* B, int M, int N, int K, int64_t lda, int64_t ldb,
                        int a_mn, int b_mn, int elt, const KmbGemmEpilogue* epi, int tile_n,
                        kmb_stream_t stream) {
  using namespace kmb;
  if (!A || !B || !epi || M <= 0 || N <= 0 || K <= 0 || (elt != 0 && elt != 1)) {
    kmb_set_last_error("kmb_gemm: bad argument", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if (elt == 1 && (a_mn || b_mn)) {
    // only the forward (parity-mode) GEMMs run in tf32 and both of their operands are K-major
    kmb_set_last_error("kmb_gemm: tf32 operands must be K-major (a_mn = b_mn = 0)", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const int esz = elt == 0 ? 2 : 4;
  if ((lda * esz) % 16 || (ldb * esz) % 16 || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) {
    kmb_set_last_error("kmb_gemm: operands must be 16-byte aligned with 16-byte row pitch", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const bool splitk_ok = epi->mode == KMB_EPI_LINEAR && epi->accumulate && epi->out_f32 && !epi->out_bf16 &&
                         !epi->residual && !epi->bias && epi->act == KMB_ACT_NONE && epi->dropout_p <= 0.f;
  int cg = 1;
  if (tile_n == 0) {
    const TileChoice c = pick_tile(M, N, b_mn, elt, splitk_ok);
    cg = c.cg;
    tile_n = c.bn;
  } else if (tile_n >= 1000) {
    cg = 2;
    tile_n -= 1000;
  }
  const bool bn_ok = cg == 1 ? (tile_n == 32 || tile_n == 64 || tile_n == 128 || tile_n == 256)
                             : (tile_n == 128 || tile_n == 192 || tile_n == 256);
  if (!bn_ok || (cg == 2 && elt != 0)) {
    kmb_set_last_error("kmb_gemm: tile_n must be 32/64/128/256 (single CTA) or 1128/1192/1256 (bf16 CTA pair)", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const int bk = TILE_BYTES_ROW / esz;
  if (b_mn && cg == 1 && tile_n < bk) tile_n = bk;
  if (b_mn && cg == 2 && ((tile_n / 2) % bk)) tile_n = 256;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + BM * cg - 1) / (BM * cg);
  p.n_tiles = (N + tile_n - 1) / tile_n;
  p.k_blocks = (K + bk - 1) / bk;
  p.e = *epi;
  p.drop_thresh16 = 0;
  p.drop_scale = 1.f;
  p.timeline = g_timeline_on;
  if (epi->dropout_p > 0.f) {
    if (!epi->dropout_seed) {
      kmb_set_last_error("kmb_gemm: dropout needs a device seed pointer", __FILE__, __LINE__);
      return KMB_ERR_ARG;
    }
    p.drop_thresh16 = (uint32_t)(epi->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 1.0f / (1.0f - epi->dropout_p);
  }
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  p.vec_ok = 1;
  if (epi->bias && !al16(epi->bias)) p.vec_ok = 0;
  if (epi->residual && (!al16(epi->residual) || (epi->ld_res % 4))) p.vec_ok = 0;
  if (epi->aux && (!al16(epi->aux) || (epi->ld_aux % 8))) p.vec_ok = 0;
  if (epi->out_f32 && (!al16(epi->out_f32) || (epi->ld_f32 % 4))) p.vec_ok = 0;
  if ((epi->out_bf16 || epi->out_preact) && (epi->ld_bf16 % 8)) p.vec_ok = 0;
  if (epi->out_bf16 && !al16(epi->out_bf16)) p.vec_ok = 0;
  if (epi->out_preact && !al16(epi->out_preact)) p.vec_ok = 0;
  if (epi->mode == KMB_EPI_CE_STATS && (!epi->labels || !epi->ce_max || !epi->ce_sum || !epi->ce_label_logit)) {
    kmb_set_last_error("kmb_gemm: CE_STATS needs labels/ce_max/ce_sum/ce_label_logit", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  if (epi->mode == KMB_EPI_CE_GRAD && (!epi->labels || !epi->ce_lse || !epi->ce_gscale || !epi->out_bf16)) {
    kmb_set_last_error("kmb_gemm: CE_GRAD needs labels/ce_lse/ce_gscale/out_bf16", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }

  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap(&tmA, A, elt, (uint64_t)M, (uint64_t)K, (uint64_t)lda, bk, BM);
  else rc = make_tmap(&tmA, A, elt, (uint64_t)K, (uint64_t)M, (uint64_t)lda, bk, bk);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap(&tmB, B, elt, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, bk, tile_n / cg);
  else rc = make_tmap(&tmB, B, elt, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, bk, bk);
  if (rc) return rc;

  // split-K: weight-gradient shaped problems (few output tiles, long K) that accumulate into fp32
  p.split_k = 1;
  p.kb_per_split = p.k_blocks;
  if (splitk_ok) {
    const int tiles = p.m_tiles * p.n_tiles;
    const int sms = gemm_num_sms() / cg;
    int best = 1;
    double best_eff = 0.0;
    for (int sk = 1; sk <= 16; ++sk) {
      if (p.k_blocks / sk < 8 && sk > 1) break;
      const int items = tiles * sk;
      const double eff = (double)items / (((items + sms - 1) / sms) * (double)sms);
      if (eff > best_eff + 0.02) { best_eff = eff; best = sk; }
    }
    if (best > 1) {
      p.kb_per_split = (p.k_blocks + best - 1) / best;
      p.split_k = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    }
  }

  // bf16 outputs of plain Linear / CE-grad epilogues leave through TMA bulk stores
  CUtensorMap tmOut, tmPre;
  memset(&tmOut, 0, sizeof tmOut);
  memset(&tmPre, 0, sizeof tmPre);
  p.tma_out = 0;
  const bool lin_fast = epi->mode == KMB_EPI_LINEAR && epi->out_bf16 && !epi->out_f32 && !epi->residual &&
                        epi->dropout_p <= 0.f;
  const bool grad_act = epi->mode == KMB_EPI_LINEAR && (epi->act == KMB_ACT_GELU_GRAD || epi->act == KMB_ACT_TANH_GRAD);
  const bool aux_ok = !grad_act || (epi->aux && al16(epi->aux) && (epi->ld_aux % 8) == 0 && !epi->out_preact);
  if ((lin_fast || epi->mode == KMB_EPI_CE_GRAD) && (epi->ld_bf16 % 8) == 0 && al16(epi->out_bf16) &&
      (!epi->out_preact || al16(epi->out_preact)) && aux_ok) {
    rc = make_tmap(&tmOut, epi->out_bf16, 0, (uint64_t)M, (uint64_t)N, (uint64_t)epi->ld_bf16, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    if (epi->out_preact) {
      rc = make_tmap(&tmPre, epi->out_preact, 0, (uint64_t)M, (uint64_t)N, (uint64_t)epi->ld_bf16, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
    if (grad_act) {   // the fast path streams the activation-gradient operand through the tmPre slot
      rc = make_tmap(&tmPre, epi->aux, 0, (uint64_t)M, (uint64_t)N, (uint64_t)epi->ld_aux, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
    p.tma_out = 1;
  }
  if (epi->mode == KMB_EPI_CE_GRAD && !p.tma_out) {
    kmb_set_last_error("kmb_gemm: CE_GRAD needs a 16-byte aligned bf16 output with a 16-byte row pitch", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cg == 2) return gemm_launch_pair(tile_n, a_mn, b_mn, tmA, tmB, tmOut, tmPre, p, st);
#define KMB_DISPATCH_BN(BNV)                                                        \
  case BNV:                                                                         \
    return elt == 0 ? launch_major<BNV, 0>(a_mn, b_mn, tmA, tmB, tmOut, tmPre, p, st) \
                    : launch_major<BNV, 1>(a_mn, b_mn, tmA, tmB, tmOut, tmPre, p, st);
  switch (tile_n) {
    KMB_DISPATCH_BN(32)
    KMB_DISPATCH_BN(64)
    KMB_DISPATCH_BN(128)
    KMB_DISPATCH_BN(256)
  }
#undef KMB_DISPATCH_BN
  return KMB_ERR_ARG;
}
