// CTA-pair (tcgen05 cta_group::2) instantiations of the GEMM kernel: 256 x {128, 192, 256} output tiles
// computed by the two SMs of a 2-CTA cluster (see gemm_kernel.cuh).
#include "gemm_launch.cuh"

namespace kmb {

template <int BN>
static int pair_major(int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                      const CUtensorMap& tmPre, const GemmParams& p, cudaStream_t st) {
  // an MN-major B tile is built from 64-element chunks, so each CTA's half of BN must be a multiple of 64
  constexpr bool kBmnOk = ((BN / 2) % 64) == 0;
  if (!a_mn && !b_mn) return gemm_launch<BN, 0, 0, 0, 2>(tmA, tmB, tmOut, tmPre, p, st);
  if (a_mn && !b_mn) return gemm_launch<BN, 0, 1, 0, 2>(tmA, tmB, tmOut, tmPre, p, st);
  if constexpr (kBmnOk) {
    if (!a_mn && b_mn) return gemm_launch<BN, 0, 0, 1, 2>(tmA, tmB, tmOut, tmPre, p, st);
    return gemm_launch<BN, 0, 1, 1, 2>(tmA, tmB, tmOut, tmPre, p, st);
  }
  kmb_set_last_error("kmb_gemm: pair tile width not usable with an MN-major B operand", __FILE__, __LINE__);
  return KMB_ERR_ARG;
}

int gemm_launch_pair(int bn, int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                     const CUtensorMap& tmPre, const GemmParams& p, cudaStream_t st) {
  switch (bn) {
    case 128: return pair_major<128>(a_mn, b_mn, tmA, tmB, tmOut, tmPre, p, st);
    case 192: return pair_major<192>(a_mn, b_mn, tmA, tmB, tmOut, tmPre, p, st);
    case 256: return pair_major<256>(a_mn, b_mn, tmA, tmB, tmOut, tmPre, p, st);
  }
  kmb_set_last_error("kmb_gemm: pair tile width must be 128, 192 or 256", __FILE__, __LINE__);
  return KMB_ERR_ARG;
}

int gemm_pair_read_timeline(unsigned long long* h8) {
  return cudaMemcpyFromSymbol(h8, g_gemm_timeline, 8 * sizeof(unsigned long long)) == cudaSuccess ? KMB_OK : KMB_ERR_CUDA;
}

}  // namespace kmb
