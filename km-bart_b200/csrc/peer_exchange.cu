// Data-parallel gradient exchange over NVLink peer memory, driven by the copy engines.
//
// The reference wraps the model in torch DDP (vcg_train.py:96-98, pretrain.py:96-98): bucketed NCCL all-reduce of the
// gradients.  NCCL's reduction kernels need whole SMs (640 threads x 96 registers per CTA do not fit beside a CTA of the
// persistent tcgen05 GEMM), so every all-reduce that overlaps the backward sweep pushes GEMMs that are exactly one wave
// of tiles into a second wave (profiles/r01k_analysis.md §8).  Here the bytes move by cudaMemcpyAsync between peer
// mappings of the flat gradient buffer (copy engines, no SM), and the only SM work is a 32-register elementwise sum
// that co-resides with any compute CTA:
//
//   region [a, b) of the flat fp32 gradient buffer becomes final on every rank at the same point of the sweep;
//   it is cut into `world` slices, rank r owns slice r.
//   1. push   my copy of slice p  -> staging[me] on rank p            (world-1 peer copies)
//   2. signal PUSHED[region][me] on every peer, wait for PUSHED[region][p] from every peer
//   3. reduce my slice:  g = (sum over ranks, in rank order) / world    (one small kernel, deterministic)
//   4. push   the reduced slice   -> the same slice of g on every peer (world-1 peer copies)
//   5. signal DONE[region][me] on every peer, wait for DONE[region][p] from every peer
//
// Every rank ends with bit-identical averaged gradients.  Flags are step counters (monotonic, never reset) in a
// peer-mapped array, written and awaited by stream memory operations (cuStreamBatchMemOp: the front end executes them, no
// CTA has to find a slot beside the compute stream's grids); KMBART_PEER_SIGNAL=kernel selects one-warp kernels instead
// (the waiter polls LOCAL memory, the signaller writes across NVLink; 20 s watchdog that traps instead of hanging).
// All of it runs on private, highest-priority, non-blocking streams — two lanes (regions alternate between them, each
// lane has its own staging slots) so that step 4 of one region overlaps step 1 of the next; the compute stream only
// records "region ready" events and joins at the end of the sweep.  Regions that are exchanged after the sweep use,
// at more than two ranks, one kernel that loads the slice from every rank and stores the average to every rank
// (px_twoshot_kernel): `world - 1` copy-engine transfers per phase are executed one at a time and lose to NVLink there.
#include <cuda.h>
#include <string.h>
#include "common.cuh"
#include "../../include/kmbart.h"

namespace kmb {

constexpr int PX_MAX_WORLD = 8;
constexpr int PX_COPY_STREAMS = 4;
constexpr int PX_LANES = 2;

struct PxPeers {
  unsigned* p[PX_MAX_WORLD];
};

__global__ void px_signal_kernel(PxPeers peers, int world, int me, size_t idx, unsigned value) {
  int t = threadIdx.x;
  __threadfence_system();
  if (t < world && t != me)
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.p[t] + idx), "r"(value) : "memory");
}

__global__ void px_wait_kernel(const unsigned* flags, size_t base, int world, int me, unsigned value) {
  int t = threadIdx.x;
  if (t < world && t != me) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned x;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(x) : "l"(flags + base + t) : "memory");
      if ((int)(x - value) >= 0) break;
      __nanosleep(100);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) {
        printf("kmbart peer exchange: rank %d waited 20 s for rank %d (flag %llu: %u < %u)\n", me, t, (unsigned long long)base, x, value);
        __trap();
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

// g[i] = scale * sum_r (r == me ? g[i] : staging[r * slot + i]), rank order
__global__ void __launch_bounds__(256) px_reduce_kernel(float* __restrict__ g, const float* __restrict__ staging, size_t slot, size_t n,
                                                         int world, int me, float scale) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    size_t n4 = n >> 2;
    float4* g4 = reinterpret_cast<float4*>(g);
    for (size_t i = tid; i < n4; i += 2 * nth) {          // two independent 16-byte columns per thread in flight
      const size_t j = i + nth;
      const bool two = j < n4;
      float4 v[PX_MAX_WORLD], w[PX_MAX_WORLD];
#pragma unroll
      for (int r = 0; r < PX_MAX_WORLD; ++r)
        if (r < world) {
          const float4* src = (r == me) ? g4 : reinterpret_cast<const float4*>(staging + (size_t)r * slot);
          v[r] = __ldcs(src + i);
          if (two) w[r] = __ldcs(src + j);
        }
      float4 a = v[0], b = w[0];
#pragma unroll
      for (int r = 1; r < PX_MAX_WORLD; ++r)
        if (r < world) {
          a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w;
          if (two) { b.x += w[r].x; b.y += w[r].y; b.z += w[r].z; b.w += w[r].w; }
        }
      a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
      g4[i] = a;
      if (two) {
        b.x *= scale; b.y *= scale; b.z *= scale; b.w *= scale;
        g4[j] = b;
      }
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += nth) {
      float a = 0.f;
      for (int r = 0; r < world; ++r) a += (r == me) ? g[i] : staging[(size_t)r * slot + i];
      g[i] = a * scale;
    }
  } else {
    for (size_t i = tid; i < n; i += nth) {
      float a = 0.f;
      for (int r = 0; r < world; ++r) a += (r == me) ? g[i] : staging[(size_t)r * slot + i];
      g[i] = a * scale;
    }
  }
}

struct PxBufs {
  float* p[PX_MAX_WORLD];
};

// Two-shot exchange of one slice in ONE kernel, for the regions that are exchanged after the sweep (nothing left to
// protect from SM contention, and `world - 1` serialized copy-engine transfers per phase are slower than NVLink):
// every thread loads its 16-byte column of the slice from all ranks' gradient buffers (peer loads), sums in rank
// order, and stores the average into every rank's buffer (peer stores).  g.p[r] points at the region start on rank r.
__global__ void __launch_bounds__(256) px_twoshot_kernel(PxBufs g, size_t lo, size_t n, int world, float scale) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  bool aligned = true;
#pragma unroll
  for (int r = 0; r < PX_MAX_WORLD; ++r)
    if (r < world) aligned = aligned && ((reinterpret_cast<uintptr_t>(g.p[r] + lo) & 15) == 0);
  size_t n4 = aligned ? (n >> 2) : 0;
  for (size_t i = tid; i < n4; i += 2 * nth) {
    const size_t j = i + nth;
    const bool two = j < n4;
    float4 v[PX_MAX_WORLD], w[PX_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < PX_MAX_WORLD; ++r)
      if (r < world) {
        const float4* src = reinterpret_cast<const float4*>(g.p[r] + lo);
        v[r] = __ldcs(src + i);
        if (two) w[r] = __ldcs(src + j);
      }
    float4 a = v[0], b = w[0];
#pragma unroll
    for (int r = 1; r < PX_MAX_WORLD; ++r)
      if (r < world) {
        a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w;
        if (two) { b.x += w[r].x; b.y += w[r].y; b.z += w[r].z; b.w += w[r].w; }
      }
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    b.x *= scale; b.y *= scale; b.z *= scale; b.w *= scale;
#pragma unroll
    for (int r = 0; r < PX_MAX_WORLD; ++r)
      if (r < world) {
        float4* dst = reinterpret_cast<float4*>(g.p[r] + lo);
        __stcs(dst + i, a);
        if (two) __stcs(dst + j, b);
      }
  }
  for (size_t i = (n4 << 2) + tid; i < n; i += nth) {
    float a = 0.f;
    for (int r = 0; r < world; ++r) a += g.p[r][lo + i];
    a *= scale;
    for (int r = 0; r < world; ++r) g.p[r][lo + i] = a;
  }
}

struct PxCtx {
  int rank, world, n_regions;
  float* g;
  float* peer_g[PX_MAX_WORLD];
  float* staging;
  float* peer_staging[PX_MAX_WORLD];
  unsigned* flags;
  PxPeers peer_flags;
  size_t slot_elems;
  cudaStream_t comm[PX_LANES];
  cudaStream_t cp[PX_LANES][PX_COPY_STREAMS];
  cudaEvent_t ev_ready, ev_join[PX_LANES], ev_mark[PX_LANES], ev_fork[PX_LANES], ev_cp[PX_LANES][PX_COPY_STREAMS];
};

// KMBART_PEER_DEBUG (timing experiments only, results are wrong): 1 = no copies, 2 = no signal / wait kernels, 4 = no reduction
static inline int px_debug() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KMBART_PEER_DEBUG");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static inline size_t px_chunk(size_t len, int world) { return (((len + world - 1) / world) + 3) & ~(size_t)3; }

}  // namespace kmb

using namespace kmb;

#define PX_TRY(x)                                                   \
  do {                                                              \
    cudaError_t e__ = (x);                                          \
    if (e__ != cudaSuccess) {                                       \
      kmb_set_last_error(cudaGetErrorString(e__), __FILE__, __LINE__); \
      return KMB_ERR_CUDA;                                          \
    }                                                               \
  } while (0)

extern "C" int kmb_ipc_export(const void* ptr, unsigned char* handle64, unsigned long long* offset) {
  typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
  static GetRange get_range = nullptr;
  if (!get_range) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) {
      kmb_set_last_error("cuMemGetAddressRange not available", __FILE__, __LINE__);
      return KMB_ERR_CUDA;
    }
    get_range = reinterpret_cast<GetRange>(f);
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (get_range(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0) {
    kmb_set_last_error("cuMemGetAddressRange failed", __FILE__, __LINE__);
    return KMB_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  PX_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *offset = (unsigned long long)(uintptr_t)ptr - base;
  return KMB_OK;
}

extern "C" int kmb_ipc_open(const unsigned char* handle64, void** base) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  PX_TRY(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess));
  return KMB_OK;
}

extern "C" int kmb_ipc_close(void* base) {
  PX_TRY(cudaIpcCloseMemHandle(base));
  return KMB_OK;
}

extern "C" int kmb_peer_can_access(int dev, int peer_dev) {
  int ok = 0;
  if (cudaDeviceCanAccessPeer(&ok, dev, peer_dev) != cudaSuccess) return 0;
  return ok;
}

extern "C" int kmb_peer_ctx_create(int rank, int world, float* g, void* const* peer_g, float* staging, void* const* peer_staging,
                                   unsigned* flags, void* const* peer_flags, size_t slot_elems, int n_regions, void** ctx_out) {
  if (world < 2 || world > PX_MAX_WORLD || rank < 0 || rank >= world || (slot_elems & 3)) {
    kmb_set_last_error("kmb_peer_ctx_create: world must be 2..8, slot_elems a multiple of 4", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  PxCtx* c = new PxCtx();
  c->rank = rank; c->world = world; c->n_regions = n_regions;
  c->g = g; c->staging = staging; c->flags = flags; c->slot_elems = slot_elems;
  for (int p = 0; p < PX_MAX_WORLD; ++p) {
    c->peer_g[p] = p < world ? static_cast<float*>(peer_g[p]) : nullptr;
    c->peer_staging[p] = p < world ? static_cast<float*>(peer_staging[p]) : nullptr;
    c->peer_flags.p[p] = p < world ? static_cast<unsigned*>(peer_flags[p]) : nullptr;
  }
  // highest priority: the block scheduler places the exchange's few small CTAs ahead of the queued CTAs of whatever
  // large kernel the compute stream is running (without it the signal / wait / reduce kernels of the last regions sit
  // behind every CTA of the AdamW kernel: profiles/r02_dp_timeline.md)
  int prio_least = 0, prio_greatest = 0;
  PX_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  for (int l = 0; l < PX_LANES; ++l) {
    PX_TRY(cudaStreamCreateWithPriority(&c->comm[l], cudaStreamNonBlocking, prio_greatest));
    for (int i = 0; i < PX_COPY_STREAMS; ++i) {
      PX_TRY(cudaStreamCreateWithPriority(&c->cp[l][i], cudaStreamNonBlocking, prio_greatest));
      PX_TRY(cudaEventCreateWithFlags(&c->ev_cp[l][i], cudaEventDisableTiming));
    }
    PX_TRY(cudaEventCreateWithFlags(&c->ev_join[l], cudaEventDisableTiming));
    PX_TRY(cudaEventCreateWithFlags(&c->ev_mark[l], cudaEventDisableTiming));
    PX_TRY(cudaEventCreateWithFlags(&c->ev_fork[l], cudaEventDisableTiming));
  }
  PX_TRY(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  *ctx_out = c;
  return KMB_OK;
}

extern "C" int kmb_peer_ctx_destroy(void* ctx) {
  PxCtx* c = static_cast<PxCtx*>(ctx);
  if (!c) return KMB_OK;
  for (int l = 0; l < PX_LANES; ++l) {
    cudaStreamSynchronize(c->comm[l]);
    for (int i = 0; i < PX_COPY_STREAMS; ++i) {
      cudaStreamDestroy(c->cp[l][i]);
      cudaEventDestroy(c->ev_cp[l][i]);
    }
    cudaStreamDestroy(c->comm[l]);
    cudaEventDestroy(c->ev_join[l]);
    cudaEventDestroy(c->ev_mark[l]);
    cudaEventDestroy(c->ev_fork[l]);
  }
  cudaEventDestroy(c->ev_ready);
  delete c;
  return KMB_OK;
}

// Flags move by stream memory operations (cuStreamBatchMemOp: executed by the front end, no CTA to schedule behind the
// compute stream's kernels); KMBART_PEER_SIGNAL=kernel, or a driver that refuses them, selects the one-warp kernels.
typedef CUresult (*PxBatchMemOp)(CUstream, unsigned int, CUstreamBatchMemOpParams*, unsigned int);
static PxBatchMemOp px_memop() {
  static int state = -1;
  static PxBatchMemOp fn = nullptr;
  if (state < 0) {
    state = 0;
    const char* e = getenv("KMBART_PEER_SIGNAL");
    if (!(e && e[0] == 'k')) {
      void* f = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuStreamBatchMemOp", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && f) {
        fn = reinterpret_cast<PxBatchMemOp>(f);
        state = 1;
      }
    }
  }
  return state == 1 ? fn : nullptr;
}
static bool g_px_memop_failed = false;

static int px_signal(PxCtx* c, cudaStream_t st, size_t idx, unsigned value) {
  const int W = c->world, me = c->rank;
  if (px_debug() & 2) return KMB_OK;
  PxBatchMemOp op = g_px_memop_failed ? nullptr : px_memop();
  if (op) {
    CUstreamBatchMemOpParams prm[PX_MAX_WORLD];
    memset(prm, 0, sizeof prm);
    int n = 0;
    for (int p = 0; p < W; ++p)
      if (p != me) {
        prm[n].writeValue.operation = CU_STREAM_MEM_OP_WRITE_VALUE_32;
        prm[n].writeValue.address = (CUdeviceptr)(uintptr_t)(c->peer_flags.p[p] + idx);
        prm[n].writeValue.value = value;
        prm[n].writeValue.flags = CU_STREAM_WRITE_VALUE_DEFAULT;
        ++n;
      }
    if (op((CUstream)st, n, prm, 0) == CUDA_SUCCESS) return KMB_OK;
    g_px_memop_failed = true;
  }
  px_signal_kernel<<<1, 32, 0, st>>>(c->peer_flags, W, me, idx, value);
  return KMB_OK;
}

static int px_wait(PxCtx* c, cudaStream_t st, size_t base, unsigned value) {
  const int W = c->world, me = c->rank;
  if (px_debug() & 2) return KMB_OK;
  PxBatchMemOp op = g_px_memop_failed ? nullptr : px_memop();
  if (op) {
    CUstreamBatchMemOpParams prm[PX_MAX_WORLD];
    memset(prm, 0, sizeof prm);
    int n = 0;
    for (int p = 0; p < W; ++p)
      if (p != me) {
        prm[n].waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_32;
        prm[n].waitValue.address = (CUdeviceptr)(uintptr_t)(c->flags + base + p);
        prm[n].waitValue.value = value;
        prm[n].waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
        ++n;
      }
    if (op((CUstream)st, n, prm, 0) == CUDA_SUCCESS) return KMB_OK;
    g_px_memop_failed = true;
  }
  px_wait_kernel<<<1, 32, 0, st>>>(c->flags, base, W, me, value);
  return KMB_OK;
}

// world-1 peer copies fanned out over the copy streams (several copy engines at once), joined back into `comm`
static int px_push_all(PxCtx* c, int lane, bool phase2, size_t a, size_t len, size_t chunk) {
  const int W = c->world, me = c->rank;
  cudaStream_t comm = c->comm[lane];
  PX_TRY(cudaEventRecord(c->ev_fork[lane], comm));
  int used = 0;
  for (int k = 1; k < W; ++k) {
    int p = (me + k) % W;                        // staggered so that no two ranks start on the same target
    size_t lo, n;
    const float* src;
    float* dst;
    if (!phase2) {                               // my copy of slice p -> staging[me] on rank p
      lo = (size_t)p * chunk;
      if (lo >= len) continue;
      n = (lo + chunk <= len ? chunk : len - lo);
      src = c->g + a + lo;
      dst = c->peer_staging[p] + ((size_t)lane * W + me) * c->slot_elems;
    } else {                                     // my reduced slice -> the same slice of g on rank p
      lo = (size_t)me * chunk;
      if (lo >= len) continue;
      n = (lo + chunk <= len ? chunk : len - lo);
      src = c->g + a + lo;
      dst = c->peer_g[p] + a + lo;
    }
    int s = used % PX_COPY_STREAMS;
    if (used < PX_COPY_STREAMS) PX_TRY(cudaStreamWaitEvent(c->cp[lane][s], c->ev_fork[lane], 0));
    if (!(px_debug() & 1)) PX_TRY(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, c->cp[lane][s]));
    ++used;
  }
  for (int s = 0; s < used && s < PX_COPY_STREAMS; ++s) {
    PX_TRY(cudaEventRecord(c->ev_cp[lane][s], c->cp[lane][s]));
    PX_TRY(cudaStreamWaitEvent(comm, c->ev_cp[lane][s], 0));
  }
  return KMB_OK;
}

extern "C" int kmb_peer_exchange_region(void* ctx, int region, size_t start, size_t end, unsigned value, int mode, kmb_stream_t compute_) {
  cudaStream_t compute = static_cast<cudaStream_t>(compute_);
  PxCtx* c = static_cast<PxCtx*>(ctx);
  if (!c || region < 0 || region >= c->n_regions || end <= start) {
    kmb_set_last_error("kmb_peer_exchange_region: bad region", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const int W = c->world, me = c->rank;
  const size_t len = end - start, chunk = px_chunk(len, W);
  if (chunk > c->slot_elems) {
    kmb_set_last_error("kmb_peer_exchange_region: region larger than the staging slots", __FILE__, __LINE__);
    return KMB_ERR_ARG;
  }
  const int lane = region % PX_LANES;
  cudaStream_t comm = c->comm[lane];
  PX_TRY(cudaEventRecord(c->ev_ready, compute));
  PX_TRY(cudaStreamWaitEvent(comm, c->ev_ready, 0));
  const size_t lo_ = (size_t)me * chunk;
  if (mode == 1) {             // READY flags, one kernel that loads from and stores to every rank, DONE flags
    const size_t f_ready = ((size_t)0 * c->n_regions + region) * W, f_fin = ((size_t)1 * c->n_regions + region) * W;
    px_signal(c, comm, f_ready + me, value);
    px_wait(c, comm, f_ready, value);
    if (lo_ < len && !(px_debug() & 4)) {
      size_t n = (lo_ + chunk <= len ? chunk : len - lo_);
      PxBufs bufs;
      for (int r = 0; r < PX_MAX_WORLD; ++r) bufs.p[r] = r == me ? c->g + start : (r < W ? c->peer_g[r] + start : nullptr);
      // One CTA per SM: 256 threads x 2 x 16 B x `world` loads in flight per SM is far more than NVLink needs, and every
      // further resident CTA (94 registers x 256 threads) takes occupancy from the AdamW kernel that runs beside it
      // (4 GPUs, ms per step: cap 592 12.53, 296 12.51, 148 12.45).  KMBART_PEER_TWOSHOT_CTAS overrides (A/B timing).
      static int cap = 0;
      if (!cap) {
        const char* e = getenv("KMBART_PEER_TWOSHOT_CTAS");
        cap = e ? atoi(e) : 0;
        if (cap < 1) {
          int dev = 0, sms = 0;
          cudaGetDevice(&dev);
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
          cap = sms > 0 ? sms : 148;
        }
      }
      int blocks = (int)((n / 8 + 255) / 256);
      if (blocks < 1) blocks = 1;
      if (blocks > cap) blocks = cap;
      px_twoshot_kernel<<<blocks, 256, 0, comm>>>(bufs, lo_, n, W, 1.0f / (float)W);
    }
    px_signal(c, comm, f_fin + me, value);
    px_wait(c, comm, f_fin, value);
    KMB_CHECK_LAUNCH();
    return KMB_OK;
  }
  int rc = px_push_all(c, lane, false, start, len, chunk);
  if (rc) return rc;
  const size_t f_pushed = ((size_t)0 * c->n_regions + region) * W, f_done = ((size_t)1 * c->n_regions + region) * W;
  px_signal(c, comm, f_pushed + me, value);
  px_wait(c, comm, f_pushed, value);
  const size_t lo = (size_t)me * chunk;
  if (lo < len && !(px_debug() & 4)) {
    size_t n = (lo + chunk <= len ? chunk : len - lo);
    int blocks = (int)((n / 8 + 255) / 256);
    if (blocks < 1) blocks = 1;
    if (blocks > 296) blocks = 296;
    px_reduce_kernel<<<blocks, 256, 0, comm>>>(c->g + start + lo, c->staging + (size_t)lane * W * c->slot_elems, c->slot_elems, n, W, me, 1.0f / (float)W);
  }
  rc = px_push_all(c, lane, true, start, len, chunk);
  if (rc) return rc;
  px_signal(c, comm, f_done + me, value);
  px_wait(c, comm, f_done, value);
  KMB_CHECK_LAUNCH();
  return KMB_OK;
}

extern "C" int kmb_peer_join(void* ctx, kmb_stream_t compute_) {
  cudaStream_t compute = static_cast<cudaStream_t>(compute_);
  PxCtx* c = static_cast<PxCtx*>(ctx);
  for (int l = 0; l < PX_LANES; ++l) {
    PX_TRY(cudaEventRecord(c->ev_join[l], c->comm[l]));
    PX_TRY(cudaStreamWaitEvent(compute, c->ev_join[l], 0));
  }
  return KMB_OK;
}

extern "C" int kmb_peer_mark(void* ctx) {
  PxCtx* c = static_cast<PxCtx*>(ctx);
  for (int l = 0; l < PX_LANES; ++l) PX_TRY(cudaEventRecord(c->ev_mark[l], c->comm[l]));
  return KMB_OK;
}

extern "C" int kmb_peer_join_mark(void* ctx, kmb_stream_t compute_) {
  cudaStream_t compute = static_cast<cudaStream_t>(compute_);
  PxCtx* c = static_cast<PxCtx*>(ctx);
  for (int l = 0; l < PX_LANES; ++l) PX_TRY(cudaStreamWaitEvent(compute, c->ev_mark[l], 0));
  return KMB_OK;
}
