// Standalone GPU check of kmb_gemm (tcgen05 path) against a naive fp32 CUDA kernel on the
// same bf16/fp32 inputs.  Run on the B200 box:  build/gemm_test
// Prints one line per case: max |err|, max |ref|, and the first mismatches (for
// descriptor debugging).  Exit code = number of failing cases.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/kmbart.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(99); } } while (0)

// A(m,k): a_mn ? A[k*lda+m] : A[m*lda+k]
template <typename T>
__global__ void ref_gemm(const T* A, const T* B, float* D, int M, int N, int K, int lda, int ldb, int a_mn, int b_mn, int tf32) {
  int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float a = (float)(a_mn ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]);
    float b = (float)(b_mn ? B[(size_t)k * ldb + n] : B[(size_t)n * ldb + k]);
    if (tf32) { a = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u); b = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u); }
    acc += a * b;
  }
  D[(size_t)m * N + n] = acc;
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

static int run_case(int M, int N, int K, int a_mn, int b_mn, int elt, int tile_n, int with_epi) {
  const int esz = elt == 0 ? 2 : 4;
  const int pad = 16 / esz;
  auto rup = [&](int x) { return (x + pad - 1) / pad * pad; };
  const int lda = a_mn ? rup(M) : rup(K), ldb = b_mn ? rup(N) : rup(K);
  const size_t a_rows = a_mn ? K : M, b_rows = b_mn ? K : N;
  std::vector<float> hA(a_rows * lda), hB(b_rows * ldb);
  for (auto& x : hA) x = frand();
  for (auto& x : hB) x = frand();
  void *dA, *dB; float *dRef, *dOut, *dBias = nullptr, *dRes = nullptr; __nv_bfloat16* dOutB;
  CK(cudaMalloc(&dA, hA.size() * esz)); CK(cudaMalloc(&dB, hB.size() * esz));
  if (elt == 0) {
    std::vector<__nv_bfloat16> t(hA.size()); for (size_t i = 0; i < t.size(); ++i) t[i] = __float2bfloat16(hA[i]);
    CK(cudaMemcpy(dA, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
    t.resize(hB.size()); for (size_t i = 0; i < t.size(); ++i) t[i] = __float2bfloat16(hB[i]);
    CK(cudaMemcpy(dB, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
  } else {
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&dRef, (size_t)M * N * 4)); CK(cudaMalloc(&dOut, (size_t)M * N * 4)); CK(cudaMalloc(&dOutB, (size_t)M * N * 2));
  CK(cudaMemset(dOut, 0xFF, (size_t)M * N * 4));
  std::vector<float> hBias(N), hRes((size_t)M * N);
  if (with_epi) {
    for (auto& x : hBias) x = frand();
    for (auto& x : hRes) x = frand();
    CK(cudaMalloc(&dBias, N * 4)); CK(cudaMalloc(&dRes, (size_t)M * N * 4));
    CK(cudaMemcpy(dBias, hBias.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dRes, hRes.data(), (size_t)M * N * 4, cudaMemcpyHostToDevice));
  }
  dim3 g((N + 127) / 128, M);
  if (elt == 0) ref_gemm<__nv_bfloat16><<<g, 128>>>((__nv_bfloat16*)dA, (__nv_bfloat16*)dB, dRef, M, N, K, lda, ldb, a_mn, b_mn, 0);
  else ref_gemm<float><<<g, 128>>>((float*)dA, (float*)dB, dRef, M, N, K, lda, ldb, a_mn, b_mn, 1);
  CK(cudaDeviceSynchronize());

  KmbGemmEpilogue e; memset(&e, 0, sizeof e);
  e.mode = KMB_EPI_LINEAR; e.alpha = 1.f; e.out_f32 = dOut; e.ld_f32 = N;
  if (with_epi == 1) { e.bias = dBias; e.residual = dRes; e.ld_res = N; e.alpha = 0.5f; if (N % 8 == 0) { e.out_bf16 = dOutB; e.ld_bf16 = N; } }
  if (with_epi == 2) { e.accumulate = 1; CK(cudaMemcpy(dOut, hRes.data(), (size_t)M * N * 4, cudaMemcpyHostToDevice)); }   // split-K path
  if (with_epi == 3) { e.out_f32 = nullptr; e.out_bf16 = dOutB; e.ld_bf16 = N; e.bias = dBias; }                          // TMA-store path
  int rc = kmb_gemm(dA, dB, M, N, K, lda, ldb, a_mn, b_mn, elt, &e, tile_n, 0);
  cudaError_t se = cudaDeviceSynchronize();
  int fail = 0;
  if (rc != 0 || se != cudaSuccess) {
    printf("CASE M=%d N=%d K=%d a_mn=%d b_mn=%d elt=%d tile_n=%d epi=%d: rc=%d (%s) cuda=%s\n", M, N, K, a_mn, b_mn, elt, tile_n, with_epi, rc, kmb_last_error(), cudaGetErrorString(se));
    if (se != cudaSuccess) exit(98);
    fail = 1;
  } else {
    std::vector<float> ref((size_t)M * N), out((size_t)M * N);
    CK(cudaMemcpy(ref.data(), dRef, ref.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
    if (with_epi == 3) {
      std::vector<__nv_bfloat16> ob((size_t)M * N);
      CK(cudaMemcpy(ob.data(), dOutB, ob.size() * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < ob.size(); ++i) out[i] = __bfloat162float(ob[i]);
    }
    double maxerr = 0, maxref = 0; int nbad = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      float r = ref[(size_t)m * N + n];
      if (with_epi == 1) r = 0.5f * r + hBias[n] + hRes[(size_t)m * N + n];
      if (with_epi == 2) r = r + hRes[(size_t)m * N + n];
      if (with_epi == 3) r = r + hBias[n];
      float o = out[(size_t)m * N + n];
      double err = fabs((double)r - o);
      if (!(err <= 1e30)) err = 1e30;
      if (err > maxerr) maxerr = err;
      if (fabs(r) > maxref) maxref = fabs(r);
      const double tol = (elt == 0 ? 2e-3 : 2e-3) * sqrt((double)K) + 1e-3 + (with_epi == 3 ? 0.01 * fabs(r) : 0.0);
      if (err > tol) { if (nbad < 6) printf("   bad (m=%d,n=%d): ref=%g out=%g\n", m, n, r, o); ++nbad; }
    }
    fail = nbad > 0;
    printf("CASE M=%d N=%d K=%d a_mn=%d b_mn=%d elt=%d tile_n=%d epi=%d: maxerr=%.4g maxref=%.4g bad=%d %s\n", M, N, K, a_mn, b_mn, elt, tile_n, with_epi, maxerr, maxref, nbad, fail ? "FAIL" : "ok");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dRef); cudaFree(dOut); cudaFree(dOutB); if (dBias) cudaFree(dBias); if (dRes) cudaFree(dRes);
  return fail;
}

__global__ void fill_rand(__nv_bfloat16* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) { unsigned x = (unsigned)i * 2654435761u + seed; x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; p[i] = __float2bfloat16(((x & 0xFFFF) / 32768.f - 1.f) * 0.5f); }
}
static int g_noout = 0, g_gelu = 0;
static void bench_case(int M, int N, int K, int a_mn, int b_mn, int tile_n, const char* name) {
  size_t na = (size_t)M * K, nb = (size_t)N * K;
  __nv_bfloat16 *dA, *dB, *dO, *dP;
  CK(cudaMalloc(&dA, na * 2)); CK(cudaMalloc(&dB, nb * 2)); CK(cudaMalloc(&dO, (size_t)M * N * 2)); CK(cudaMalloc(&dP, (size_t)M * N * 2));
  fill_rand<<<(na + 255) / 256, 256>>>(dA, na, 1); fill_rand<<<(nb + 255) / 256, 256>>>(dB, nb, 2);
  KmbGemmEpilogue e; memset(&e, 0, sizeof e); e.alpha = 1.f; e.out_bf16 = g_noout ? nullptr : dO; e.ld_bf16 = N;
  if (g_gelu) { e.act = KMB_ACT_GELU; e.out_preact = dP; }
  const int lda = a_mn ? M : K, ldb = b_mn ? N : K;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) kmb_gemm(dA, dB, M, N, K, lda, ldb, a_mn, b_mn, 0, &e, tile_n, 0);
  cudaEventRecord(e0);
  const int iters = 20;
  for (int i = 0; i < iters; ++i) kmb_gemm(dA, dB, M, N, K, lda, ldb, a_mn, b_mn, 0, &e, tile_n, 0);
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  printf("BENCH %-12s M=%d N=%d K=%d mn=%d%d tile_n=%d: %.3f ms  %.1f TFLOP/s\n", name, M, N, K, a_mn, b_mn, tile_n, ms, 2.0 * M * N * K / ms / 1e9);
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dP);
}

int main(int argc, char** argv) {
  if (kmb_arch_check() != 0) { printf("arch check failed: %s\n", kmb_last_error()); return 97; }
  if (argc > 2 && !strcmp(argv[1], "one")) {   // single shape for ncu: one <mode 0|1|2>
    const int mode = atoi(argv[2]);
    g_noout = mode == 0; g_gelu = mode == 2;
    bench_case(12800, 3072, 768, 0, 0, 256, "fc1");
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "timeline")) {   // where does a launch spend its time? (block 0, ns)
    for (int mode = 0; mode < 2; ++mode) {
      g_noout = mode == 0;
      for (int tn : {256, 1256}) {
        int shapes[4][3] = {{18944, 256, 768}, {18944, 2048, 768}, {18944, 256, 3072}, {12800, 2304, 768}};
        for (auto& sh : shapes) {
          kmb_gemm_debug_timeline(1, 0, nullptr);
          bench_case(sh[0], sh[1], sh[2], 0, 0, tn, "tl");
          CK(cudaDeviceSynchronize());
          unsigned long long t[7];
          kmb_gemm_debug_timeline(0, tn > 1000, t);
          printf("   timeline ns: setup %llu  first-load %llu  first-acc %llu  first-epi %llu  last-epi %llu  exit %llu\n", t[1] - t[0],
                 t[2] - t[0], t[3] - t[0], t[4] - t[0], t[5] - t[0], t[6] - t[0]);
        }
      }
    }
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "sweep")) {   // mainloop / epilogue decomposition (profiles/)
    for (int mode = 0; mode < 3; ++mode) {
      g_noout = mode == 0; g_gelu = mode == 2;
      printf("--- mode %s\n", mode == 0 ? "no output (mainloop bound)" : mode == 1 ? "bf16 out" : "gelu + preact + bf16 out");
      for (int tn : {256, 1256}) {
        bench_case(8192, 8192, 8192, 0, 0, tn, "big");
        bench_case(18944, 2048, 8192, 0, 0, tn, "1wave-longK");   // 148 single tiles x 8 = 8 waves exactly / 74 pairs x 8
        bench_case(18944, 2048, 768, 0, 0, tn, "8wave-K768");
        bench_case(18944, 256, 768, 0, 0, tn, "1wave-K768");
        bench_case(18944, 256, 3072, 0, 0, tn, "1wave-K3072");
        bench_case(12800, 2304, 768, 0, 0, tn, "qkv");
        bench_case(12800, 3072, 768, 0, 0, tn, "fc1");
        bench_case(12800, 768, 3072, 0, 0, tn, "fc2");
      }
    }
    return 0;
  }
  int fails = 0;
  srand(1);
  // smallest possible first: one tile, one k-block, K-major both
  fails += run_case(128, 64, 64, 0, 0, 0, 64, 0);
  fails += run_case(128, 128, 256, 0, 0, 0, 128, 0);
  fails += run_case(128, 64, 64, 0, 1, 0, 64, 0);
  fails += run_case(128, 64, 64, 1, 0, 0, 64, 0);
  fails += run_case(128, 128, 128, 1, 1, 0, 128, 0);
  for (int tn : {32, 64, 128, 256}) fails += run_case(300, 520, 200, 0, 0, 0, tn, 1);
  for (int amn = 0; amn < 2; ++amn) for (int bmn = 0; bmn < 2; ++bmn) {
    fails += run_case(384, 768, 1000, amn, bmn, 0, 0, 1);
    fails += run_case(1000, 264, 328, amn, bmn, 0, 256, 0);
    if (!amn && !bmn) {   // tf32 (fp32-parity mode) is K-major only
      fails += run_case(256, 256, 96, 0, 0, 1, 128, 0);
      fails += run_case(200, 136, 100, 0, 0, 1, 0, 1);
    }
  }
  fails += run_case(256, 512, 8192, 1, 1, 0, 0, 2);    // split-K reds
  fails += run_case(768, 768, 12800, 1, 1, 0, 0, 2);
  fails += run_case(200, 328, 4000, 1, 1, 0, 128, 2);
  fails += run_case(1000, 520, 200, 0, 0, 0, 256, 3);  // TMA-store epilogue with ragged M / N
  fails += run_case(1000, 520, 200, 0, 1, 0, 128, 3);
  fails += run_case(300, 72, 200, 0, 0, 0, 64, 3);
  fails += run_case(4096, 768, 768, 0, 0, 0, 0, 1);   // many tiles per CTA (pipeline wrap)
  fails += run_case(20000, 256, 64, 0, 0, 0, 128, 0); // > 148 tiles with 1 k-block
  printf("gemm_test: %d failing case(s)\n", fails);
  if (argc > 1 && !strcmp(argv[1], "bench2")) {
    for (int mode = 0; mode < 3; ++mode) {
      g_noout = mode == 0; g_gelu = mode == 2;
      printf("--- mode %s\n", mode == 0 ? "no output (mainloop bound)" : mode == 1 ? "bf16 out" : "gelu + preact + bf16 out");
      for (int tn : {128, 256}) {
        bench_case(12800, 768, 768, 0, 0, tn, "proj");
        bench_case(12800, 2304, 768, 0, 0, tn, "qkv");
        bench_case(12800, 3072, 768, 0, 0, tn, "fc1");
        bench_case(12800, 768, 3072, 0, 0, tn, "fc2");
        bench_case(12800, 768, 3072, 0, 1, tn, "dgrad_fc1");
        bench_case(768, 3072, 12800, 1, 1, tn, "wgrad_fc2");
      }
    }
    bench_case(8192, 8192, 8192, 0, 0, 256, "big");
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "bench3")) {   // CTA-pair tile widths per shape, plain and GELU epilogues
    for (int mode = 1; mode < 3; ++mode) {
      g_noout = 0; g_gelu = mode == 2;
      printf("--- mode %s\n", mode == 1 ? "bf16 out" : "gelu + preact + bf16 out");
      for (int tn : {0, 1128, 1192, 1256}) {
        bench_case(12800, 3072, 768, 0, 0, tn, "fc1");
        bench_case(6144, 3072, 768, 0, 0, tn, "fc1_dec");
        if (mode == 1) {
          bench_case(12800, 2304, 768, 0, 0, tn, "qkv");
          bench_case(12800, 768, 3072, 0, 0, tn, "fc2");
          bench_case(6144, 768, 768, 0, 0, tn, "proj_dec");
          bench_case(12800, 768, 768, 0, 0, tn, "proj_enc");
        }
      }
    }
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "bench")) {
    bench_case(12800, 768, 768, 0, 0, 0, "proj");
    bench_case(12800, 768, 768, 0, 0, 128, "proj128");
    bench_case(12800, 768, 768, 0, 0, 256, "proj256");
    bench_case(12800, 2304, 768, 0, 0, 0, "qkv");
    bench_case(12800, 3072, 768, 0, 0, 0, "fc1");
    bench_case(12800, 3072, 768, 0, 0, 256, "fc1_256");
    bench_case(12800, 768, 3072, 0, 0, 0, "fc2");
    bench_case(768, 3072, 12800, 1, 1, 0, "wgrad_fc2");
    bench_case(12800, 768, 3072, 0, 1, 0, "dgrad_mn");
    bench_case(6144, 50320, 768, 0, 0, 0, "lmhead");
    bench_case(6144, 50320, 768, 0, 0, 256, "lmhead256");
    bench_case(8192, 8192, 8192, 0, 0, 256, "big");
  }
  return fails;
}
