// common.cuh — shared device/host helpers of the B200 retrieval kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "../../include/trueno_rag_b200.h"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
void trr_set_error(const std::string& msg);
int trr_fail(int status, const std::string& msg);

#define TRR_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      return trr_fail(_e == cudaErrorMemoryAllocation ? TRR_ERR_OOM : TRR_ERR_CUDA,                 \
                      std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" +  \
                          std::to_string(__LINE__));                                                \
    }                                                                                               \
  } while (0)

#define TRR_CHECK(expr)        \
  do {                         \
    int _s = (expr);           \
    if (_s != TRR_OK) return _s; \
  } while (0)

// Triage knobs (environment variables read by the orchestration) exist only in -DTRR_TRIAGE builds, which
// tools/gpu_probe.py uses; the default build ignores the environment, so a stray variable cannot change a result.
#ifdef TRR_TRIAGE
#include <stdlib.h>
#define TRR_KNOB(name) getenv(name)
#else
#define TRR_KNOB(name) (static_cast<const char*>(nullptr))
#endif

// ------------------------------------------------------------------------------------------------
// context: one GPU, one stream, a reusable workspace, launch counters
// ------------------------------------------------------------------------------------------------
struct trr_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;        // stream every entry point enqueues on (may be caller-owned)
  cudaStream_t owned_stream = nullptr;  // stream created with the context
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::mutex mu;             // serialises search calls on this context (re-entrancy)
  void* ws = nullptr;        // CtxExtra (growable device buffers), see capi.cu
  void* pin = nullptr;       // pinned host staging (grown on demand)
  size_t pin_bytes = 0;
  uint64_t launches = 0;     // kernels of this library launched on the context
};

int trr_ctx_reserve_ws(trr_ctx* ctx, size_t bytes);
int trr_ctx_reserve_pin(trr_ctx* ctx, size_t bytes);

// carve aligned sub-buffers out of the workspace
struct WsCarver {
  char* base;
  size_t off = 0;
  explicit WsCarver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* r = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return r;
  }
  static size_t need(std::initializer_list<size_t> sizes) {
    size_t t = 0;
    for (size_t s : sizes) t = ((t + 255) & ~size_t(255)) + s;
    return t + 256;
  }
};

// ------------------------------------------------------------------------------------------------
// canonical ordering keys: larger key == better hit.
//   high 32 bits: order-preserving image of the f32 score (-0.0 folded onto +0.0, because the
//   reference's partial_cmp treats them as equal); low 32 bits: ~ordinal, so that among equal
//   scores the smaller ordinal wins.  (score desc, ordinal asc) == key desc.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t trr_f32_orderable(float f) {
  f = f + 0.0f;  // -0.0 -> +0.0 (exact for every other value)
#if defined(__CUDA_ARCH__)
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } v; v.f = f; uint32_t u = v.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__host__ __device__ __forceinline__ float trr_orderable_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } v; v.u = u; return v.f;
#endif
}

__host__ __device__ __forceinline__ uint64_t trr_make_key(float score, uint32_t ord) {
  return (static_cast<uint64_t>(trr_f32_orderable(score)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - ord);
}
__host__ __device__ __forceinline__ float trr_key_score(uint64_t key) {
  return trr_orderable_f32(static_cast<uint32_t>(key >> 32));
}
__host__ __device__ __forceinline__ uint32_t trr_key_ord(uint64_t key) {
  return 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFu);
}
// key 0 is below every real key (a real key has a non-zero high word: orderable(x) >= 0x00800000 for
// every finite x and -inf maps to 0x007FFFFF) and is used as "empty".
#define TRR_KEY_EMPTY 0ull

__host__ __device__ __forceinline__ uint32_t trr_pow2_ceil(uint32_t x) {
  uint32_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------
// bitonic sort (descending) of n = power-of-two u64 keys in shared memory.
// `tid`/`nthreads` describe the cooperating group; SYNC is __syncthreads() or __syncwarp().
// ------------------------------------------------------------------------------------------------
template <typename SyncFn>
__device__ __forceinline__ void trr_bitonic_sort_desc(uint64_t* keys, uint32_t n, uint32_t tid, uint32_t nthreads,
                                                      SyncFn sync) {
  for (uint32_t size = 2; size <= n; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      sync();
      for (uint32_t i = tid; i < (n >> 1); i += nthreads) {
        uint32_t lo = 2 * i - (i & (stride - 1));  // index with bit `stride` cleared
        uint32_t hi = lo + stride;
        bool desc = ((lo & size) == 0);
        uint64_t a = keys[lo], b = keys[hi];
        bool swap = desc ? (a < b) : (a > b);
        if (swap) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  sync();
}

struct BlockSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
struct WarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, bulk async copy (TMA 1-D), TMA 2-D tensor tiles, tcgen05
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t trr_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void trr_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(trr_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void trr_fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void trr_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void trr_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(trr_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void trr_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(trr_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool trr_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(trr_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void trr_mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!trr_mbar_try_wait(bar, parity)) {
  }
}
// bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU
__device__ __forceinline__ void trr_mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  if (trr_mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!trr_mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes multiple of 16, 16-B aligned)
__device__ __forceinline__ void trr_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   trr_smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(trr_smem_u32(bar))
               : "memory");
}
// 2-D TMA tile load global -> shared (tensor map in kernel parameter space), completion on an mbarrier
__device__ __forceinline__ void trr_tma_load_2d(const void* map, void* smem_dst, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(trr_smem_u32(smem_dst)), "l"(map), "r"(trr_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
#endif  // __CUDACC__
