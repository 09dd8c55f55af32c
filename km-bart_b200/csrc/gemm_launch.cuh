// Host-side launch helper shared by the two translation units that instantiate gemm_tc05_kernel
// (single-CTA tiles in gemm_tc05.cu, CTA-pair tiles in gemm_pair.cu; split so they compile in parallel).
#pragma once
#include "gemm_kernel.cuh"

namespace kmb {

// SMs the persistent GEMM grids (and the tile / split-K cost models) plan for.  KMBART_GEMM_SMS caps it: under data
// parallelism NCCL's all-reduce CTAs occupy a few SMs while the backward GEMMs run; a statically scheduled persistent grid
// that assumes all 148 then has CTA pairs that cannot be resident and run as a second wave (kmbart/parallel.py sets the cap
// together with NCCL_MAX_CTAS).
inline int gemm_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    const char* e = getenv("KMBART_GEMM_SMS");
    if (e && atoi(e) >= 2 && atoi(e) < n) n = atoi(e) & ~1;
  }
  return n;
}

template <int BN, int ELT, int A_MN, int B_MN, int CG, int ES>
static int gemm_launch_es(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmPre,
                       const GemmParams& p, cudaStream_t st) {
  using C = Cfg<BN, CG, ES>;
  auto kern = gemm_tc05_kernel<BN, ELT, A_MN, B_MN, CG, ES>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      kmb_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return KMB_ERR_CUDA;
    }
    attr_set = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.split_k;
  const int groups = gemm_num_sms() / CG;
  const int grid = (tiles < groups ? tiles : groups) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128 + 128 * ES);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CG;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {   // the kernel executes griddepcontrol.wait before its first global access
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmPre, p);
  if (e != cudaSuccess) {
    kmb_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
    return KMB_ERR_CUDA;
  }
  return KMB_OK;
}

// epilogue-warp count by epilogue weight (see gemm_kernel.cuh): activations, activation gradients and the
// cross-entropy epilogues take 16 warps, plain linear epilogues 8
template <int BN, int ELT, int A_MN, int B_MN, int CG>
static int gemm_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmPre,
                       const GemmParams& p, cudaStream_t st) {
  const bool heavy = p.e.mode != KMB_EPI_LINEAR || p.e.act != KMB_ACT_NONE;
  if (heavy) return gemm_launch_es<BN, ELT, A_MN, B_MN, CG, 4>(tmA, tmB, tmOut, tmPre, p, st);
  return gemm_launch_es<BN, ELT, A_MN, B_MN, CG, 2>(tmA, tmB, tmOut, tmPre, p, st);
}

// defined in gemm_pair.cu: CTA-pair (cta_group::2) tiles of 256 x bn, bf16 operands
int gemm_launch_pair(int bn, int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                     const CUtensorMap& tmPre, const GemmParams& p, cudaStream_t st);

int gemm_pair_read_timeline(unsigned long long* h8);

}  // namespace kmb
