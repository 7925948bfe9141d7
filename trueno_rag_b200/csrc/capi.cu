// capi.cu — the C ABI of include/trueno_rag_b200.h: handle management and kernel orchestration.
// No arithmetic of the retrieval path lives here; it sequences the kernels of dense_scan.cu (K1, merge,
// rescoring), dense_gemm.cu (K2), bm25.cu (K3) and fusion.cu (K4) on the context stream.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler is attached

#include "bm25.cuh"
#include "common.cuh"
#include "dense.cuh"
#include "fusion.cuh"

// NVTX range around the host side of a library step (Nsight Systems shows which entry point enqueued which kernels)
struct TrrRange {
  explicit TrrRange(const char* name) { nvtxRangePushA(name); }
  ~TrrRange() { nvtxRangePop(); }
  TrrRange(const TrrRange&) = delete;
  TrrRange& operator=(const TrrRange&) = delete;
};

cudaError_t trr_launch_synth_rows(uint64_t seed, uint64_t first_row, uint64_t n, uint32_t dim, int dups, int to_bf16,
                                  void* out, cudaStream_t st);
cudaError_t trr_launch_flush(void* p, size_t bytes, cudaStream_t st);
cudaError_t trr_launch_gemm_topk_dump(const GemmTopkArgs& a, const void* map_q128, const void* map_d128, unsigned grid,
                                      float* dump, uint32_t dump_ld, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
void trr_set_error(const std::string& msg) { g_last_error = msg; }
int trr_fail(int status, const std::string& msg) {
  g_last_error = msg;
  return status;
}

extern "C" const char* trr_last_error(void) { return g_last_error.c_str(); }
extern "C" int trr_version(void) { return 100; }
extern "C" int trr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ------------------------------------------------------------------------------------------------
// growable device buffers
// ------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t n) {
    if (n <= bytes) return TRR_OK;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    n = (n + 4095) & ~size_t(4095);
    TRR_CUDA(cudaMalloc(&p, n));
    bytes = n;
    return TRR_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct CtxExtra {
  uint32_t* dbg_host = nullptr;  // pinned + mapped: kernels record the site of a barrier timeout before trapping
  uint32_t* dbg_dev = nullptr;
  DevBuf scratch;  // kernel-internal scratch of dense / bm25 searches
  DevBuf io;       // device copies of host inputs / outputs of the host-buffer entry points
  DevBuf hy;       // hybrid exchange record (single-GPU path)
  DevBuf flush;
};

static CtxExtra* extra(trr_ctx* c) { return reinterpret_cast<CtxExtra*>(c->ws); }

int trr_ctx_reserve_ws(trr_ctx* ctx, size_t bytes) { return extra(ctx)->scratch.reserve(bytes); }
int trr_ctx_reserve_pin(trr_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pin_bytes) return TRR_OK;
  if (ctx->pin) cudaFreeHost(ctx->pin);
  ctx->pin = nullptr; ctx->pin_bytes = 0;
  TRR_CUDA(cudaMallocHost(&ctx->pin, bytes));
  ctx->pin_bytes = bytes;
  return TRR_OK;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" int trr_ctx_create(int device, trr_ctx** out) {
  if (!out) return trr_fail(TRR_ERR_INVALID_ARG, "trr_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return trr_fail(TRR_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  }
  if (device < 0 || device >= n) return trr_fail(TRR_ERR_INVALID_ARG, "trr_ctx_create: bad device index");
  cudaDeviceProp prop;
  TRR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return trr_fail(TRR_ERR_UNSUPPORTED, "this library is built for sm_100a (B200) only");
  TRR_CUDA(cudaSetDevice(device));
  trr_ctx* c = new trr_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->ws = new CtxExtra();
  {
    CtxExtra* x = extra(c);
    if (cudaHostAlloc(reinterpret_cast<void**>(&x->dbg_host), 64, cudaHostAllocMapped) == cudaSuccess) {
      memset(x->dbg_host, 0, 64);
      cudaHostGetDevicePointer(reinterpret_cast<void**>(&x->dbg_dev), x->dbg_host, 0);
    } else {
      cudaGetLastError();
    }
  }
  TRR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->owned_stream = c->stream;
  for (auto& e : c->ev) TRR_CUDA(cudaEventCreate(&e));
  *out = c;
  return TRR_OK;
}

extern "C" int trr_ctx_destroy(trr_ctx* c) {
  if (!c) return TRR_OK;
  DeviceGuard g(c->device);
  cudaStreamSynchronize(c->stream);
  CtxExtra* x = extra(c);
  x->scratch.release(); x->io.release(); x->hy.release(); x->flush.release();
  if (x->dbg_host) cudaFreeHost(x->dbg_host);
  delete x;
  if (c->pin) cudaFreeHost(c->pin);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  cudaStreamDestroy(c->owned_stream);
  delete c;
  return TRR_OK;
}

extern "C" int trr_ctx_sync(trr_ctx* c) {
  if (!c) return trr_fail(TRR_ERR_INVALID_ARG, "ctx is NULL");
  DeviceGuard g(c->device);
  TRR_CUDA(cudaStreamSynchronize(c->stream));
  return TRR_OK;
}
extern "C" int trr_ctx_stream(trr_ctx* c, void** out) {
  if (!c || !out) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  *out = c->stream;
  return TRR_OK;
}
extern "C" int trr_ctx_sm_count(trr_ctx* c, int* out) {
  if (!c || !out) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  *out = c->sm_count;
  return TRR_OK;
}
extern "C" int trr_ctx_set_stream(trr_ctx* c, void* stream) {
  if (!c) return trr_fail(TRR_ERR_INVALID_ARG, "ctx is NULL");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  TRR_CUDA(cudaStreamSynchronize(c->stream));
  c->stream = reinterpret_cast<cudaStream_t>(stream);  // the original stream is kept alive until destroy
  return TRR_OK;
}
extern "C" int trr_ctx_launch_count(trr_ctx* c, uint64_t* out) {
  if (!c || !out) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  *out = c->launches;
  return TRR_OK;
}
extern "C" int trr_ctx_flush_l2(trr_ctx* c, size_t bytes) {
  if (!c) return trr_fail(TRR_ERR_INVALID_ARG, "ctx is NULL");
  DeviceGuard g(c->device);
  TRR_CHECK(extra(c)->flush.reserve(bytes));
  TRR_CUDA(trr_launch_flush(extra(c)->flush.p, bytes, c->stream));
  return TRR_OK;
}

// ------------------------------------------------------------------------------------------------
// dense store
// ------------------------------------------------------------------------------------------------
struct trr_dense {
  trr_ctx* ctx = nullptr;
  uint32_t dim = 0;
  int metric = 0, dtype = 0;
  uint32_t elem = 4, row_bytes = 0;
  uint64_t n = 0, cap = 0, n_dead = 0;
  uint32_t base = 0;
  uint8_t* rows = nullptr;
  float* norms = nullptr;
  uint8_t* dead = nullptr;
  uint64_t frozen_n = 0;
  // tensor-core path operands
  bool gemm_ready = false;
  DevBuf shadow, scale_bias, max_norm, qbuf;
  uint32_t dim_pad = 0;
  uint64_t n_tiles = 0;
  alignas(64) uint8_t map_d[128];       // documents, 256-row box (1-CTA kernel)
  alignas(64) uint8_t map_d_half[128];  // documents, 128-row box (2-CTA kernel: each CTA loads half a tile)
  alignas(64) uint8_t map_scan[128];    // slab in its own dtype, 32-row x 128-byte box (K1 TMA ring)
  uint64_t map_scan_n = 0;              // rows covered by map_scan (0 = not built)
  const void* map_scan_base = nullptr;
  uint32_t* stat_dev = nullptr;         // [2] device copy of {n_flagged, max_gap bits} of the last GEMM search
  bool stat_pending = false;
  // Width of the first exact re-scoring pass, adapted to the data: the candidate proof needs (width - k) ranks of score
  // spacing to exceed the a-priori error bound, and the spacing depends on the corpus (dimension, size, distribution).
  // After every GEMM search the number of queries that failed the first proof lands in page-locked host memory
  // (asynchronous copy); the next search doubles the width while more than 1/64 of the batch failed.  Results are exact at
  // every width - only the time moves.
  uint32_t* scan_done = nullptr;        // [256] arrival counters of the scan kernel's fused merge epilogue (kept zero)
  uint32_t cp_level = 0;
  uint32_t* feedback_host = nullptr;    // [2] {queries of the last search, queries that failed the first proof}
  uint32_t feedback_B = 0;
  int mode = TRR_DENSE_AUTO;
  trr_stats stats{};
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // [0,1] whole call, [2,3] dominant kernel
};

static int dense_grow(trr_dense* h, uint64_t need) {
  if (need <= h->cap) return TRR_OK;
  uint64_t ncap = std::max<uint64_t>(need, h->cap ? h->cap * 2 : 1024);
  uint8_t* nrows = nullptr; float* nnorms = nullptr; uint8_t* ndead = nullptr;
  if (cudaMalloc(&nrows, ncap * h->row_bytes + 256) != cudaSuccess || cudaMalloc(&nnorms, ncap * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&ndead, ncap) != cudaSuccess) {
    cudaGetLastError();
    if (nrows) cudaFree(nrows);
    if (nnorms) cudaFree(nnorms);
    return trr_fail(TRR_ERR_OOM, "dense store: out of device memory");
  }
  cudaStream_t st = h->ctx->stream;
  TRR_CUDA(cudaMemsetAsync(ndead, 0, ncap, st));
  if (h->n) {
    TRR_CUDA(cudaMemcpyAsync(nrows, h->rows, h->n * h->row_bytes, cudaMemcpyDeviceToDevice, st));
    TRR_CUDA(cudaMemcpyAsync(nnorms, h->norms, h->n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRR_CUDA(cudaMemcpyAsync(ndead, h->dead, h->n, cudaMemcpyDeviceToDevice, st));
  }
  TRR_CUDA(cudaStreamSynchronize(st));
  if (h->rows) cudaFree(h->rows);
  if (h->norms) cudaFree(h->norms);
  if (h->dead) cudaFree(h->dead);
  h->rows = nrows; h->norms = nnorms; h->dead = ndead; h->cap = ncap;
  return TRR_OK;
}

extern "C" int trr_dense_create(trr_ctx* ctx, uint32_t dim, int metric, int dtype, uint64_t capacity_hint,
                                trr_dense** out) {
  if (!ctx || !out) return trr_fail(TRR_ERR_INVALID_ARG, "trr_dense_create: NULL argument");
  *out = nullptr;
  if (dim == 0) return trr_fail(TRR_ERR_INVALID_ARG, "trr_dense_create: dimension must be > 0");
  if (metric < 0 || metric > 2) return trr_fail(TRR_ERR_INVALID_ARG, "trr_dense_create: bad metric");
  if (dtype != TRR_DTYPE_F32 && dtype != TRR_DTYPE_BF16) return trr_fail(TRR_ERR_INVALID_ARG, "bad dtype");
  DeviceGuard g(ctx->device);
  trr_dense* h = new trr_dense();
  h->ctx = ctx; h->dim = dim; h->metric = metric; h->dtype = dtype;
  h->elem = dtype == TRR_DTYPE_BF16 ? 2 : 4;
  h->row_bytes = dim * h->elem;
  for (auto& e : h->ev) cudaEventCreate(&e);
  if (capacity_hint) {
    int s = dense_grow(h, capacity_hint);
    if (s != TRR_OK) { delete h; return s; }
  }
  *out = h;
  return TRR_OK;
}

extern "C" int trr_dense_destroy(trr_dense* h) {
  if (!h) return TRR_OK;
  DeviceGuard g(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  if (h->rows) cudaFree(h->rows);
  if (h->norms) cudaFree(h->norms);
  if (h->dead) cudaFree(h->dead);
  if (h->stat_dev) cudaFree(h->stat_dev);
  if (h->feedback_host) cudaFreeHost(h->feedback_host);
  if (h->scan_done) cudaFree(h->scan_done);
  h->shadow.release(); h->scale_bias.release(); h->max_norm.release(); h->qbuf.release();
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  delete h;
  return TRR_OK;
}

extern "C" int trr_dense_set_base(trr_dense* h, uint32_t base) {
  if (!h) return trr_fail(TRR_ERR_INVALID_ARG, "handle is NULL");
  h->base = base;
  return TRR_OK;
}
extern "C" int trr_dense_set_mode(trr_dense* h, int mode) {
  if (!h || mode < 0 || mode > 2) return trr_fail(TRR_ERR_INVALID_ARG, "bad mode");
  h->mode = mode;
  return TRR_OK;
}
extern "C" int trr_dense_len(trr_dense* h, uint64_t* out) {
  if (!h || !out) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  *out = h->n - h->n_dead;
  return TRR_OK;
}

static int dense_append_common(trr_dense* h, const void* src, uint64_t n, int src_kind /*0 f32 host,1 bf16 host,2 dev*/) {
  if (!h || (!src && n)) return trr_fail(TRR_ERR_INVALID_ARG, "trr_dense_append: NULL argument");
  if (n == 0) return TRR_OK;
  if (h->n + n + h->base > 0xFFFFFFFEull) return trr_fail(TRR_ERR_UNSUPPORTED, "more than 2^32-2 ordinals");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  TRR_CHECK(dense_grow(h, h->n + n));
  cudaStream_t st = h->ctx->stream;
  uint8_t* dst = h->rows + h->n * h->row_bytes;
  if (src_kind == 2) {
    TRR_CUDA(cudaMemcpyAsync(dst, src, n * h->row_bytes, cudaMemcpyDeviceToDevice, st));
  } else if ((src_kind == 0 && h->dtype == TRR_DTYPE_F32) || (src_kind == 1 && h->dtype == TRR_DTYPE_BF16)) {
    TRR_CUDA(cudaMemcpyAsync(dst, src, n * h->row_bytes, cudaMemcpyHostToDevice, st));
  } else if (src_kind == 0 && h->dtype == TRR_DTYPE_BF16) {
    // f32 host rows into a bf16 store: upload in chunks, round to nearest even on the device
    const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / ((uint64_t)h->dim * 4));
    TRR_CHECK(extra(h->ctx)->io.reserve(chunk * h->dim * 4));
    for (uint64_t r0 = 0; r0 < n; r0 += chunk) {
      const uint64_t m = std::min(chunk, n - r0);
      TRR_CUDA(cudaMemcpyAsync(extra(h->ctx)->io.p, static_cast<const float*>(src) + r0 * h->dim, m * h->dim * 4,
                               cudaMemcpyHostToDevice, st));
      trr_launch_shadow(extra(h->ctx)->io.p, 0, h->dim, h->dim, 0, m,
                        reinterpret_cast<uint16_t*>(dst + r0 * h->row_bytes), st);
      h->ctx->launches++;
      TRR_CUDA(cudaStreamSynchronize(st));
    }
  } else {
    return trr_fail(TRR_ERR_INVALID_ARG, "bf16 rows can only be appended to a bf16 store");
  }
  TRR_CUDA(cudaStreamSynchronize(st));
  h->n += n;
  h->gemm_ready = false;
  return TRR_OK;
}

extern "C" int trr_dense_append(trr_dense* h, const float* rows, uint64_t n) { return dense_append_common(h, rows, n, 0); }
extern "C" int trr_dense_append_bf16(trr_dense* h, const uint16_t* rows, uint64_t n) {
  return dense_append_common(h, rows, n, 1);
}
extern "C" int trr_dense_append_device(trr_dense* h, const void* d_rows, uint64_t n) {
  return dense_append_common(h, d_rows, n, 2);
}

extern "C" int trr_dense_append_synth(trr_dense* h, uint64_t seed, uint64_t first_row, uint64_t n, int dups) {
  if (!h) return trr_fail(TRR_ERR_INVALID_ARG, "handle is NULL");
  if (n == 0) return TRR_OK;
  if (h->n + n + h->base > 0xFFFFFFFEull) return trr_fail(TRR_ERR_UNSUPPORTED, "more than 2^32-2 ordinals");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  TRR_CHECK(dense_grow(h, h->n + n));
  TRR_CUDA(trr_launch_synth_rows(seed, first_row, n, h->dim, dups, h->dtype == TRR_DTYPE_BF16,
                                 h->rows + h->n * h->row_bytes, h->ctx->stream));
  h->ctx->launches++;
  TRR_CUDA(cudaStreamSynchronize(h->ctx->stream));
  h->n += n;
  h->gemm_ready = false;
  return TRR_OK;
}

extern "C" int trr_dense_remove(trr_dense* h, uint32_t ordinal) {
  if (!h) return trr_fail(TRR_ERR_INVALID_ARG, "handle is NULL");
  if (ordinal >= h->n) return trr_fail(TRR_ERR_INVALID_ARG, "trr_dense_remove: ordinal out of range");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  uint8_t was = 0;
  TRR_CUDA(cudaMemcpy(&was, h->dead + ordinal, 1, cudaMemcpyDeviceToHost));
  if (!was) {
    const uint8_t one = 1;
    TRR_CUDA(cudaMemcpy(h->dead + ordinal, &one, 1, cudaMemcpyHostToDevice));
    h->n_dead++;
    h->gemm_ready = false;
  }
  return TRR_OK;
}

static int dense_freeze_locked(trr_dense* h) {
  if (h->frozen_n < h->n) {
    trr_launch_norms(h->dtype == TRR_DTYPE_BF16, h->rows, h->dim, h->frozen_n, h->n - h->frozen_n, h->norms,
                     h->ctx->stream);
    h->ctx->launches++;
    TRR_CUDA(cudaGetLastError());
    h->frozen_n = h->n;
  }
  return TRR_OK;
}

extern "C" int trr_dense_freeze(trr_dense* h) {
  if (!h) return trr_fail(TRR_ERR_INVALID_ARG, "handle is NULL");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  TRR_CHECK(dense_freeze_locked(h));
  TRR_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return TRR_OK;
}

extern "C" int trr_dense_copy_norms(trr_dense* h, float* out, uint64_t n) {
  if (!h || !out) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  TRR_CHECK(dense_freeze_locked(h));
  TRR_CUDA(cudaStreamSynchronize(h->ctx->stream));
  TRR_CUDA(cudaMemcpy(out, h->norms, std::min<uint64_t>(n, h->n) * sizeof(float), cudaMemcpyDeviceToHost));
  return TRR_OK;
}

extern "C" int trr_dense_copy_rows(trr_dense* h, const uint32_t* ordinals, uint64_t n, void* out_rows) {
  if (!h || (n && (!ordinals || !out_rows))) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  TRR_CUDA(cudaStreamSynchronize(h->ctx->stream));
  for (uint64_t i = 0; i < n; ++i) {
    if (ordinals[i] >= h->n) return trr_fail(TRR_ERR_INVALID_ARG, "ordinal out of range");
    TRR_CUDA(cudaMemcpy(static_cast<char*>(out_rows) + i * h->row_bytes, h->rows + (uint64_t)ordinals[i] * h->row_bytes,
                        h->row_bytes, cudaMemcpyDeviceToHost));
  }
  return TRR_OK;
}

// builds the operands of the tensor-core pass: bf16 shadow (if the slab cannot be used in place),
// per-document scale/bias, max norm, TMA descriptor
static int dense_prepare_gemm(trr_dense* h) {
  if (h->gemm_ready) return TRR_OK;
  cudaStream_t st = h->ctx->stream;
  h->n_tiles = (h->n + TRR_GEMM_TILE_N - 1) / TRR_GEMM_TILE_N;
  const uint64_t n_padded = h->n_tiles * TRR_GEMM_TILE_N;
  const void* operand = h->rows;
  uint64_t cols = h->dim;
  if (h->dtype != TRR_DTYPE_BF16 || (h->dim % 8) != 0) {
    h->dim_pad = (h->dim + 63) / 64 * 64;
    TRR_CHECK(h->shadow.reserve(h->n * (uint64_t)h->dim_pad * 2));
    trr_launch_shadow(h->rows, h->dtype == TRR_DTYPE_BF16, h->dim, h->dim_pad, 0, h->n,
                      reinterpret_cast<uint16_t*>(h->shadow.p), st);
    h->ctx->launches++;
    operand = h->shadow.p;
    cols = h->dim_pad;
  } else {
    h->dim_pad = h->dim;
  }
  TRR_CHECK(h->scale_bias.reserve(n_padded * sizeof(float2)));
  TRR_CHECK(h->max_norm.reserve(256));
  TRR_CUDA(cudaMemsetAsync(h->max_norm.p, 0, 4, st));
  trr_launch_gemm_operands(h->norms, h->dead, h->n, n_padded, h->metric, reinterpret_cast<float2*>(h->scale_bias.p),
                           reinterpret_cast<float*>(h->max_norm.p), st);
  h->ctx->launches++;
  TRR_CUDA(cudaGetLastError());
  TRR_CHECK(trr_make_tensor_map(h->map_d, operand, h->n, cols, TRR_GEMM_TILE_N));
  TRR_CHECK(trr_make_tensor_map(h->map_d_half, operand, h->n, cols, TRR_GEMM_TILE_N / 2));
  h->gemm_ready = true;
  return TRR_OK;
}

struct ScanPlan {
  bool tma;   // 2-D TMA ring kernel (rows of a multiple of 16 bytes, >= 4096 rows)
  uint32_t n_slots;
  uint32_t nq;  // queries per pass over the slab (TMA kernel: 1 or 4)
  bool bulk;
  unsigned grid;
  uint32_t warps, cap, ch_bytes, n_chunks;
  size_t smem;
};

static int plan_scan(trr_dense* h, uint32_t k, ScanPlan* p, uint32_t n_queries = 1) {
  p->nq = 1;
  p->cap = trr_pow2_ceil(k + 32);
  if (p->cap < 64) p->cap = 64;
  const uint32_t q_bytes = (h->dim * 4 + 127) & ~127u;
  const size_t optin = h->ctx->smem_optin;
  p->tma = false; p->n_slots = 0;
  p->bulk = (h->row_bytes % 16 == 0) && h->n >= 4096;
  if (p->bulk && !TRR_KNOB("TRR_SCAN_NO_TMA")) {
    // 16 warps per CTA hide the per-box serial overheads (measured 6.44 TB/s vs 5.70 TB/s with 4 warps on 1M x 384 f32);
    // ring slots per warp come from what is left of the shared memory (4 KB per slot, at least 3)
    uint32_t nw_first = 16;
    if (const char* e = TRR_KNOB("TRR_SCAN_WARPS")) nw_first = (uint32_t)std::min(16, std::max(1, atoi(e)));
    // several queries on the exact path share each pass over the slab four at a time (if the buffers fit with >= 8 warps)
    uint32_t nq = (n_queries >= 2 && !TRR_KNOB("TRR_SCAN_NQ1")) ? 4u : 1u;
    if (nq == 4 && trr_scan_tma_smem(h->dim, p->cap, 3, 8, 4) > optin) nq = 1;
    p->nq = nq;
    for (uint32_t nw = nw_first; nw >= 1; nw >>= 1) {
      const size_t fixed = trr_scan_tma_smem(h->dim, p->cap, 0, nw, nq);
      if (fixed + (size_t)nw * 3 * 4104 > optin) continue;
      uint32_t slots = (uint32_t)((optin - fixed) / ((size_t)nw * (4096 + 8)));
      if (slots > 12) slots = 12;
      if (const char* e = TRR_KNOB("TRR_SCAN_SLOTS")) slots = std::min<uint32_t>(slots, (uint32_t)std::max(3, atoi(e)));
      p->tma = true; p->bulk = false; p->n_slots = slots; p->warps = nw;
      p->grid = (unsigned)h->ctx->sm_count;
      p->ch_bytes = 0; p->n_chunks = 0;
      p->smem = trr_scan_tma_smem(h->dim, p->cap, slots, nw, nq);
      return TRR_OK;
    }
  }
  if (p->bulk) {
    const size_t fixed = q_bytes + 128 + (size_t)4 * p->cap * 8;
    if (fixed + 128 * 32 > optin) p->bulk = false;
    else {
      size_t per_row = (optin - fixed) / 128;                 // bytes of pitch available per staged row
      uint32_t max_ch = (uint32_t)((per_row - 16) & ~size_t(15));
      if (max_ch > 4096) max_ch = 4096;
      if (max_ch < 64) p->bulk = false;
      else {
        p->n_chunks = (h->row_bytes + max_ch - 1) / max_ch;
        p->ch_bytes = ((h->row_bytes + p->n_chunks - 1) / p->n_chunks + 15) & ~15u;
        p->n_chunks = (h->row_bytes + p->ch_bytes - 1) / p->ch_bytes;
        p->warps = 4;
        p->grid = (unsigned)h->ctx->sm_count;
        p->smem = fixed + (size_t)128 * (p->ch_bytes + 16);
      }
    }
  }
  if (!p->bulk) {
    p->warps = 8;
    p->ch_bytes = 0; p->n_chunks = 0;
    p->smem = q_bytes + (size_t)8 * p->cap * 8;
    if (p->smem > optin) return trr_fail(TRR_ERR_UNSUPPORTED, "dimension / k too large for the scan kernel");
    const uint64_t groups = (h->n + 255) / 256;
    p->grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(groups, (uint64_t)h->ctx->sm_count * 4));
  }
  return TRR_OK;
}

// exact scan (K1) of the queries selected by (d_sel, n_sel) or all B; results into d_ord/d_score/d_n rows
// d_n_sel (nullable): DEVICE count of selected queries (<= n_sel, which then only sizes the buffers and the grids)
static int dense_scan_locked(trr_dense* h, const float* d_q, const float* d_qn, uint32_t n_sel, const uint32_t* d_sel,
                             uint32_t k, uint32_t* d_ord, float* d_score, uint32_t* d_n, uint64_t* d_keys,
                             size_t scratch_off, const uint32_t* d_n_sel = nullptr, bool record_events = true,
                             bool fused = false, const float* h_q = nullptr, float h_qn = 0.0f) {
  TrrRange nvtx_range("trr:dense_scan");
  ScanPlan p;
  TRR_CHECK(plan_scan(h, k, &p, n_sel));
  const uint64_t lists = (uint64_t)p.grid;  // one merged list per CTA
  const size_t need = scratch_off + WsCarver::need({(size_t)n_sel * lists * k * 8, (size_t)n_sel * lists * 4});
  TRR_CHECK(extra(h->ctx)->scratch.reserve(need));
  WsCarver ws(static_cast<char*>(extra(h->ctx)->scratch.p) + scratch_off);
  uint64_t* partial = ws.take<uint64_t>((size_t)n_sel * lists * k);
  uint32_t* partial_n = ws.take<uint32_t>((size_t)n_sel * lists);
  DenseScanArgs a{};
  a.rows = h->rows; a.row_bytes = h->row_bytes; a.dim = h->dim; a.n_rows = h->n;
  a.norms = h->norms; a.dead = h->n_dead ? h->dead : nullptr;
  a.q = d_q; a.q_norms = d_qn; a.sel = d_sel; a.n_sel_ptr = d_n_sel; a.n_sel = n_sel;
  a.ch_bytes = p.ch_bytes; a.n_chunks = p.n_chunks; a.n_slots = p.n_slots; a.k = k; a.cap = p.cap; a.base_ord = h->base;
  a.partial = partial; a.partial_n = partial_n;
  cudaStream_t st = h->ctx->stream;
  // fused: the single-launch form (merge by the last CTA; d_qn == NULL: query norm in the kernel's prologue)
  if (fused) {
    if (!h->scan_done) {
      TRR_CUDA(cudaMalloc(&h->scan_done, 256 * 4));
      TRR_CUDA(cudaMemsetAsync(h->scan_done, 0, 256 * 4, st));
    }
    a.done = h->scan_done; a.out_keys = d_keys; a.out_ord = d_ord; a.out_score = d_score; a.out_n = d_n;
  }
  if (p.tma && (h->map_scan_n != h->n || h->map_scan_base != h->rows)) {
    TRR_CHECK(trr_make_tensor_map_ex(h->map_scan, h->rows, h->n, h->dim, h->elem, 128 / h->elem, 32));
    h->map_scan_n = h->n; h->map_scan_base = h->rows;
  }
  if (record_events) TRR_CUDA(cudaEventRecord(h->ev[2], st));
  if (p.tma) TRR_CUDA(trr_launch_scan_tma(a, h->map_scan, h->dtype == TRR_DTYPE_BF16, h->metric, p.grid, p.warps, p.nq, p.smem, st, h_q, h_qn));
  else TRR_CUDA(trr_launch_scan(a, h->dtype == TRR_DTYPE_BF16, h->metric, p.bulk, p.grid, p.smem, st));
  if (record_events) TRR_CUDA(cudaEventRecord(h->ev[3], st));
  h->ctx->launches++;
  if (fused) return TRR_OK;
  TopkMergeArgs m{};
  m.lists = partial; m.list_n = partial_n; m.n_lists = (uint32_t)lists; m.list_stride = k;
  m.n_rows = n_sel; m.n_rows_ptr = d_n_sel; m.row_map = d_sel; m.k = k; m.k2 = trr_pow2_ceil(k);
  m.out_keys = d_keys; m.out_ord = d_ord; m.out_score = d_score; m.out_n = d_n;
  TRR_CUDA(trr_launch_topk_merge(m, d_n_sel ? std::min<uint32_t>(n_sel, 4u * (uint32_t)h->ctx->sm_count) : n_sel, st));
  h->ctx->launches++;
  return TRR_OK;
}

// reads the fallback count / max gap of the last GEMM search (the stream must be idle)
static void dense_resolve_stats(trr_dense* h) {
  if (!h->stat_pending || !h->stat_dev) return;
  uint32_t hc[2] = {0, 0};
  if (cudaMemcpy(hc, h->stat_dev, 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
    h->stats.n_guard_fallbacks = hc[0];
    memcpy(&h->stats.max_fast_exact_gap, &hc[1], 4);
  }
  h->stat_pending = false;
}

static int dense_search_locked(trr_dense* h, const float* d_q, uint32_t B, uint32_t k, uint32_t* d_ord, float* d_score,
                               uint32_t* d_n, bool sync_stats, const float* d_qn_pre = nullptr, const float* h_q1 = nullptr,
                               float h_qn1 = 0.0f) {
  TrrRange nvtx_range("trr:dense_search");
  trr_ctx* c = h->ctx;
  cudaStream_t st = c->stream;
  // very large batches are served in pieces (scratch and the query-block grid scale with B)
  constexpr uint32_t kMaxBatch = 4096;
  if (B > kMaxBatch) {
    for (uint32_t b0 = 0; b0 < B; b0 += kMaxBatch) {
      const uint32_t nb = std::min(kMaxBatch, B - b0);
      TRR_CHECK(dense_search_locked(h, d_q + (size_t)b0 * h->dim, nb, k, d_ord + (size_t)b0 * k, d_score + (size_t)b0 * k,
                                    d_n + b0, sync_stats, d_qn_pre ? d_qn_pre + b0 : nullptr));
    }
    h->stats.n_queries = B;
    return TRR_OK;
  }
  const uint64_t launches0 = c->launches;
  h->stats = trr_stats{};
  h->stats.n_queries = B;
  if (B == 0) return TRR_OK;
  if (k == 0 || h->n == h->n_dead) {
    TRR_CUDA(cudaMemsetAsync(d_n, 0, (size_t)B * 4, st));
    return TRR_OK;
  }
  if (k > 1024) return trr_fail(TRR_ERR_UNSUPPORTED, "k > 1024 is not supported yet");
  TRR_CHECK(dense_freeze_locked(h));
  TRR_CUDA(cudaEventRecord(h->ev[0], st));

  const uint64_t n_live = h->n - h->n_dead;
  bool use_gemm = false;
  if (h->mode == TRR_DENSE_GEMM) use_gemm = true;
  else if (h->mode == TRR_DENSE_AUTO) use_gemm = B >= 2 && h->n >= 16384;  // K1 re-streams the slab per query: 8 queries over 10M x 768 take 19 ms through K1, 2.6 ms through K2
  // feedback of the previous GEMM search on this store (see trr_dense::cp_level)
  if (h->feedback_host && h->feedback_B) {
    const uint32_t failed = *reinterpret_cast<volatile uint32_t*>(h->feedback_host + 1);
    if (failed != 0xFFFFFFFFu) {
      if (failed * 64u > h->feedback_B && h->cp_level < 4) ++h->cp_level;  // (a failed query costs a wide pass; one exact scan costs a pass over the slab)
      h->feedback_host[1] = 0xFFFFFFFFu;  // consumed
    }
  }
  // base width: k plus a margin of ranks, as a power of two (64 up to k = 50, 128 up to k = 100, ... 2048 for k = 1024)
  uint32_t CP = std::min<uint32_t>(std::max<uint32_t>(TRR_GEMM_CP, trr_pow2_ceil(k + std::max<uint32_t>(14, k / 4))) << h->cp_level, 2048u);
  if (use_gemm) {
    // the width cannot exceed what the half-slice lists of the tensor-core pass hold (2 * slices * 32 entries)
    const uint32_t n_qb = (B + TRR_GEMM_TILE_M - 1) / TRR_GEMM_TILE_M;
    const uint64_t tiles = (h->n + TRR_GEMM_TILE_N - 1) / TRR_GEMM_TILE_N;
    const uint64_t n_sl = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c->sm_count / std::max(n_qb, 1u), tiles));
    while (CP > TRR_GEMM_CP && CP / 2 >= k + 14 && 2 * n_sl * TRR_GEMM_CPS_MAX < CP) CP >>= 1;
    if (2 * n_sl * TRR_GEMM_CPS_MAX < CP || CP < k) {
      // (large k with many query blocks: too few document slices per block to hold k candidates per query)
      if (h->mode == TRR_DENSE_GEMM) return trr_fail(TRR_ERR_UNSUPPORTED, "GEMM mode: batch too large for this k");
      use_gemm = false;
    }
  }
  // scratch: query norms first
  TRR_CHECK(extra(c)->scratch.reserve(WsCarver::need({(size_t)B * 4})));
  float* d_qn = nullptr;
  size_t scratch_off = 0;
  {
    WsCarver ws(extra(c)->scratch.p);
    d_qn = ws.take<float>(B);
    scratch_off = (ws.off + 255) & ~size_t(255);
  }

  if (!use_gemm) {
    // the scratch buffer may be re-allocated by dense_scan_locked: compute the norms after reserving there
    ScanPlan p;
    TRR_CHECK(plan_scan(h, k, &p, B));
    const uint64_t lists = (uint64_t)p.grid;
    TRR_CHECK(extra(c)->scratch.reserve(scratch_off + WsCarver::need({(size_t)B * lists * k * 8, (size_t)B * lists * 4})));
    d_qn = reinterpret_cast<float*>(extra(c)->scratch.p);
    // few queries on the TMA kernel: ONE launch (|q| in the kernel's prologue, the merge of the per-CTA lists by the last
    // CTA, inside the ring memory) instead of norm kernel + scan + merge kernel
    const bool fused = p.tma && (B + p.nq - 1) / p.nq <= 256 && !TRR_KNOB("TRR_SCAN_NOFUSE");
    const float* qn = d_qn_pre;  // norms computed by the caller (host-buffer entry point: on the host, in reference order)
    if (!qn && !fused) {
      trr_launch_query_norms(d_q, h->dim, B, d_qn, st);
      c->launches++;
      qn = d_qn;
    }
    TRR_CHECK(dense_scan_locked(h, d_q, qn, B, nullptr, k, d_ord, d_score, d_n, nullptr, scratch_off, nullptr, true, fused,
                                B == 1 ? h_q1 : nullptr, h_qn1));
    h->stats.mode_used = TRR_DENSE_SCAN;
  } else {
    TRR_CHECK(dense_prepare_gemm(h));
    uint32_t n_qblocks = (B + TRR_GEMM_TILE_M - 1) / TRR_GEMM_TILE_M;
    // The 2-CTA kernel (cta_group::2: each CTA of a pair loads half of the document tile) moves a third less operand
    // data into shared memory per FLOP; the B200 runs this GEMM against its power cap (SM clock 1.3-1.6 GHz of 1.965), so
    // the saving is time: 12.9 ms vs 14.1 ms at 10M x 768, B = 1024 under back-to-back launches.  It needs an even number
    // of query blocks; an odd count is padded only when the padding costs less than the gain (TRR_GEMM_PAIR=0/1 overrides).
    int pair_mode = n_qblocks >= 2 && (n_qblocks % 2 == 0 || n_qblocks >= 9) ? 1 : 0;
    if (const char* e = TRR_KNOB("TRR_GEMM_PAIR")) pair_mode = atoi(e) ? 1 : 0;
    if (pair_mode) n_qblocks = (n_qblocks + 1) & ~1u;
    if (n_qblocks > (uint32_t)c->sm_count)
      return trr_fail(TRR_ERR_UNSUPPORTED, "batch larger than 128 x SM count; split the batch");
    uint32_t n_slices = (uint32_t)c->sm_count / n_qblocks;
    if (n_slices > h->n_tiles) n_slices = (uint32_t)h->n_tiles;
    if (n_slices == 0) n_slices = 1;
    const uint32_t B_pad = n_qblocks * TRR_GEMM_TILE_M;
    // exact re-scoring width per query: k plus a margin of ranks for the candidate proof
    // (TRR_GEMM_CP128=1 forces the wide width for every k: to be measured - with f32 queries the quantisation term of the
    // proof is worth about 10 ranks of score spacing, so k = 50 of 64 often falls through to the second pass)
    // list length per (query, slice): the global top-CP by fast score is spread over the slices (about CP / n_slices
    // per slice), so short lists suffice when there are many slices (Poisson tail < 1e-6 for the choices below); the
    // candidate proof (rescore_select_kernel) catches the data sets where they do not
    // The kernels keep one list per HALF slice (the two epilogue warp groups of a CTA split the columns of every tile), so
    // the re-scoring kernel sees 2 * n_slices "virtual" slices of at most TRR_GEMM_CPS_MAX entries.
    const uint32_t vslices = 2 * n_slices;
    const uint32_t cps_max = TRR_GEMM_CPS_MAX;
    const float per_slice = (float)CP / (float)vslices;
    // (measured at cfg5, width 256 = 7 per half slice: lists of 16 make the tensor-core pass 2 ms faster than lists of 32, but
    // a few queries per batch then fail even the wide proof - a full list bounds what it dropped by its own minimum - and
    // one exact scan of the 20 GB shard costs 5 ms)
    uint32_t cps = per_slice <= 1.0f ? 8u : (per_slice <= 4.0f ? 16u : 32u);
    if (const char* e = TRR_KNOB("TRR_GEMM_CPS")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) cps = (uint32_t)v; }
    if ((uint64_t)vslices * cps < CP) cps = cps_max;  // (vslices * cps_max >= CP is checked before taking this path)
    const size_t n_cand = (size_t)vslices * n_qblocks * TRR_GEMM_TILE_M * cps;
    const uint32_t cap2 = std::max<uint32_t>(trr_pow2_ceil(vslices * cps), 2 * CP);
    // scratch layout (single reservation so that pointers stay valid)
    ScanPlan p;
    TRR_CHECK(plan_scan(h, k, &p, B));
    const uint64_t lists = (uint64_t)p.grid;
    // queries whose candidate proof fails are re-run through the exact scan.  Normally the fallback is DEVICE-DRIVEN (the
    // scan and merge kernels read the number of flagged queries from device memory and return at once when it is zero), so
    // the whole search is enqueued without a host round trip; very large B x k falls back to a host-read count and chunks.
    const bool dev_fallback = (size_t)B * lists * k * 8 <= ((size_t)256 << 20) && !TRR_KNOB("TRR_GEMM_HOST_FALLBACK");
    const uint32_t fb_chunk = dev_fallback ? B : 64;  // fallback queries per scan launch
    // merged thresholds (helper warps of the 2-CTA kernel): every list publishes its pub_rank-th best score
    // (j = the smallest power of two with ceil(CP / j) lists available; s = how many lists may fall short of it)
    uint32_t pub_rank = 0, pub_pick = 0;
    if (pair_mode && !TRR_KNOB("TRR_GEMM_NOMERGE")) {
      for (uint32_t j = 1; j <= std::min<uint32_t>(cps, 8u); j <<= 1) {
        const uint32_t lists_needed = (CP + j - 1) / j;
        if (lists_needed <= vslices) { pub_rank = j; pub_pick = std::min<uint32_t>(vslices - lists_needed + 1, 8u); break; }
      }
    }
    const size_t n_pub = pub_rank ? (size_t)vslices * B_pad : 0;
    const size_t need = WsCarver::need({(size_t)B * 4, (size_t)B * 4, n_cand * 4, n_cand * 4, (size_t)(B_pad + n_pub) * 4,
                                        (size_t)B * 4, (size_t)B * 4, (size_t)B * 4, 256, (size_t)B_pad * h->dim_pad * 2,
                                        (size_t)fb_chunk * lists * k * 8 + 512, (size_t)fb_chunk * lists * 4 + 512}) +
                        4096;
    TRR_CHECK(extra(c)->scratch.reserve(need));
    WsCarver ws(extra(c)->scratch.p);
    d_qn = ws.take<float>(B);
    float* d_qdelta = ws.take<float>(B);
    float* cand_score = ws.take<float>(n_cand);
    uint32_t* cand_ord = ws.take<uint32_t>(n_cand);
    uint32_t* gthr = ws.take<uint32_t>(B_pad + n_pub);  // [B_pad] shared thresholds, then [vslices][B_pad] published list ranks
    uint32_t* flags = ws.take<uint32_t>(B);
    uint32_t* flagged = ws.take<uint32_t>(B);
    uint32_t* flagged2 = ws.take<uint32_t>(B);
    uint32_t* counters = ws.take<uint32_t>(64);  // [0] n_flagged (first pass), [1] max_gap (float bits), [4] n_flagged after the wide pass, [5] resolved by it
    uint16_t* q_bf16 = ws.take<uint16_t>((size_t)B_pad * h->dim_pad);
    const size_t fb_off = (ws.off + 255) & ~size_t(255);

    trr_launch_query_prep(d_q, h->dim, h->dim_pad, B, B_pad, q_bf16, d_qdelta, d_qn, st);  // also ||q|| (reference order)
    c->launches += 1;
    TRR_CUDA(cudaMemsetAsync(gthr, 0, (size_t)(B_pad + n_pub) * 4, st));
    TRR_CUDA(cudaMemsetAsync(counters, 0, 256, st));
    alignas(64) uint8_t map_q[128];
    TRR_CHECK(trr_make_tensor_map(map_q, q_bf16, B_pad, h->dim_pad, TRR_GEMM_TILE_M));
    GemmTopkArgs ga{};
    ga.n_qblocks = n_qblocks; ga.n_slices = n_slices; ga.n_tiles = (uint32_t)h->n_tiles;
    ga.k_blocks = (h->dim_pad + 63) / 64; ga.base_ord = h->base;
    ga.scale_bias = reinterpret_cast<const float2*>(h->scale_bias.p);
    ga.cand_score = cand_score; ga.cand_ord = cand_ord; ga.gthr = gthr; ga.share_thresholds = 1;
    ga.pair_mode = pair_mode; ga.cps = cps;
    ga.pub = pub_rank ? gthr + B_pad : nullptr; ga.pub_rank = pub_rank; ga.pub_pick = pub_pick;
    ga.dbg = extra(c)->dbg_dev;
    if (const char* e = TRR_KNOB("TRR_GEMM_DEBUG")) ga.debug_mode = atoi(e);      // perf triage only; results are wrong
    if (const char* e = TRR_KNOB("TRR_GEMM_NOSHARE")) ga.share_thresholds = atoi(e) ? 0 : 1;
    TRR_CUDA(cudaEventRecord(h->ev[2], st));
    TRR_CUDA(trr_launch_gemm_topk(ga, map_q, pair_mode ? h->map_d_half : h->map_d, n_slices * n_qblocks, st));
    TRR_CUDA(cudaEventRecord(h->ev[3], st));
    c->launches++;

    RescoreArgs ra{};
    ra.cand_score = cand_score; ra.cand_ord = cand_ord; ra.n_slices = vslices; ra.n_qblocks = n_qblocks;
    ra.cps = cps; ra.cp = CP; ra.cap2 = cap2; ra.gthr = gthr;
    ra.rows = h->rows; ra.dim = h->dim; ra.norms = h->norms; ra.base_ord = h->base; ra.n_live = n_live;
    ra.q = d_q; ra.q_norms = d_qn; ra.q_norms_out = nullptr; ra.q_delta = d_qdelta; ra.max_norm = reinterpret_cast<const float*>(h->max_norm.p);
    ra.B = B; ra.k = k; ra.metric = h->metric;
    // candidate rows are staged in shared memory in chunks (28 KB per CTA keeps six CTAs per SM)
    uint32_t stage_budget = 28u << 10;
    if (const char* e = TRR_KNOB("TRR_RESCORE_STAGE_KB")) stage_budget = (uint32_t)std::max(1, atoi(e)) << 10;
    const int per_cand = (int)(stage_budget / CP) - 16;  // bytes of a candidate row that fit the staging budget
    ra.stage_chunk = (h->row_bytes % 16 == 0 && per_cand >= 64 && !TRR_KNOB("TRR_RESCORE_NO_STAGE"))
                         ? std::min<uint32_t>(h->row_bytes, (uint32_t)per_cand & ~15u) : 0u;
    // |fast - exact| <= eps_rel * |q||d|: products of bf16 values are exact in f32; the tensor-core sum and the
    // reference's sequential sum each carry at most D roundings of relative size 2^-23 on partial sums bounded by
    // sum|q_i d_i| <= |q||d|; the scale multiply, the division and the norm product add a few more ulps.
    // A f32 store scored through its bf16 shadow adds the quantisation term 2^-8 * |q||d|.
    float eps_rel = (2.0f * (float)h->dim + 8.0f) * 1.1920929e-07f;
    if (h->dtype != TRR_DTYPE_BF16) eps_rel += 0.00390625f;
    ra.eps_rel = eps_rel;
    ra.out_keys = nullptr; ra.out_ord = d_ord; ra.out_score = d_score; ra.out_n = d_n;
    ra.flags = flags; ra.flagged = flagged; ra.n_flagged = counters; ra.max_gap = reinterpret_cast<float*>(counters + 1);
    TRR_CUDA(trr_launch_rescore(ra, h->dtype == TRR_DTYPE_BF16, st));
    c->launches++;
    if (!h->feedback_host) {
      TRR_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->feedback_host), 64));
      h->feedback_host[0] = 0; h->feedback_host[1] = 0xFFFFFFFFu;
    }
    h->feedback_B = B;
    TRR_CUDA(cudaMemcpyAsync(h->feedback_host + 1, counters, 4, cudaMemcpyDeviceToHost, st));
    // Second level: a query whose proof failed (rank spacing tighter than the a-priori error bound: large dimensions, large
    // k) is re-scored over EVERYTHING the slices kept (n_slices x cps candidates instead of the best CP), so that the only
    // documents left out are the ones the slices dropped, which sit far further down the ranking.  Only the queries that
    // still fail go to the exact scan.  CTAs beyond the device-side count of flagged queries return at once.
    const uint32_t cp_wide = trr_pow2_ceil(vslices * cps);
    uint32_t* flagged_final = flagged;
    uint32_t* counters_final = counters;
    if (cp_wide > CP && cp_wide <= 2048 && !TRR_KNOB("TRR_GEMM_NO_WIDE")) {
      RescoreArgs rw = ra;
      rw.cp = cp_wide; rw.cap2 = std::max(cap2, cp_wide); rw.stage_chunk = 0;
      rw.sel = flagged; rw.sel_n = counters;
      rw.flagged = flagged2; rw.n_flagged = counters + 4; rw.n_resolved = counters + 5;
      rw.max_gap = reinterpret_cast<float*>(counters + 1);
      TRR_CUDA(trr_launch_rescore(rw, h->dtype == TRR_DTYPE_BF16, st));
      c->launches++;
      flagged_final = flagged2;
      counters_final = counters + 4;
    }
    h->stats.eps_bound = eps_rel;
    h->stats.rescore_width = CP;
    if (dev_fallback) {
      if (!h->stat_dev) TRR_CUDA(cudaMalloc(&h->stat_dev, 64));
      TRR_CUDA(cudaMemcpyAsync(h->stat_dev, counters_final, 4, cudaMemcpyDeviceToDevice, st));
      TRR_CUDA(cudaMemcpyAsync(h->stat_dev + 1, counters + 1, 4, cudaMemcpyDeviceToDevice, st));
      h->stat_pending = true;
      if (!(ga.debug_mode & 7))
        TRR_CHECK(dense_scan_locked(h, d_q, d_qn, B, flagged_final, k, d_ord, d_score, d_n, nullptr, fb_off, counters_final, false));
    } else {
      uint32_t hc[2] = {0, 0};
      TRR_CUDA(cudaMemcpyAsync(&hc[0], counters_final, 4, cudaMemcpyDeviceToHost, st));
      TRR_CUDA(cudaMemcpyAsync(&hc[1], counters + 1, 4, cudaMemcpyDeviceToHost, st));
      cudaError_t se = cudaStreamSynchronize(st);
      if (se != cudaSuccess) {
        const uint32_t w = extra(c)->dbg_host ? extra(c)->dbg_host[0] : 0;
        return trr_fail(TRR_ERR_CUDA, std::string("GEMM path failed: ") + cudaGetErrorString(se) +
                                          " (barrier-timeout word 0x" + [](uint32_t v) { char b[16]; snprintf(b, 16, "%x", v); return std::string(b); }(w) + ")");
      }
      h->stats.n_guard_fallbacks = hc[0];
      memcpy(&h->stats.max_fast_exact_gap, &hc[1], 4);
      if (ga.debug_mode & 7) hc[0] = 0;  // perf triage: results are meaningless, do not time the fallback
      for (uint32_t f0 = 0; f0 < hc[0]; f0 += fb_chunk) {
        const uint32_t m = std::min(fb_chunk, hc[0] - f0);
        TRR_CHECK(dense_scan_locked(h, d_q, d_qn, m, flagged_final + f0, k, d_ord, d_score, d_n, nullptr, fb_off, nullptr, false));
      }
    }
    h->stats.mode_used = TRR_DENSE_GEMM;
  }
  TRR_CUDA(cudaEventRecord(h->ev[1], st));
  h->stats.n_kernel_launches = (uint32_t)(c->launches - launches0);
  if (sync_stats) {
    cudaError_t se = cudaStreamSynchronize(st);
    if (se != cudaSuccess) {
      const uint32_t w = extra(c)->dbg_host ? extra(c)->dbg_host[0] : 0;
      char wb[16]; snprintf(wb, 16, "%x", w);
      return trr_fail(TRR_ERR_CUDA, std::string("dense search failed: ") + cudaGetErrorString(se) +
                                        " (barrier-timeout word 0x" + wb + ")");
    }
    cudaEventElapsedTime(&h->stats.ms_total, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->stats.ms_main_kernel, h->ev[2], h->ev[3]);
    dense_resolve_stats(h);
    if (TRR_KNOB("TRR_GEMM_DEBUG") && (atoi(TRR_KNOB("TRR_GEMM_DEBUG")) & 8) && extra(c)->dbg_host)
      fprintf(stderr, "[trr] K2 CTA 0: %u cycles in %u ns = %.0f MHz; MMA warp waited %u on operands, %u on the epilogue; "
                      "producer(s) waited %u / %u on free stages\n", extra(c)->dbg_host[8], extra(c)->dbg_host[9],
              1000.0 * extra(c)->dbg_host[8] / std::max(1u, extra(c)->dbg_host[9]), extra(c)->dbg_host[10],
              extra(c)->dbg_host[11], extra(c)->dbg_host[12], extra(c)->dbg_host[13]);
  }
  return TRR_OK;
}

extern "C" int trr_dense_search_device(trr_dense* h, const float* d_q, uint32_t B, uint32_t k, uint32_t* d_ord,
                                       float* d_score, uint32_t* d_n) {
  if (!h || (B && (!d_q || !d_ord || !d_score || !d_n))) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  return dense_search_locked(h, d_q, B, k, d_ord, d_score, d_n, false);
}

// |q| exactly as the reference computes it (src/index.rs:442-443): sequential f32 sum of squares, then sqrt.  volatile keeps
// the host compiler from contracting the multiply-add or re-associating the sum.
static void host_query_norms(const float* q, uint32_t dim, uint32_t B, float* out) {
  for (uint32_t b = 0; b < B; ++b) {
    const float* p = q + (size_t)b * dim;
    volatile float s = 0.0f;
    for (uint32_t j = 0; j < dim; ++j) {
      volatile float sq = p[j] * p[j];
      s = s + sq;
    }
    out[b] = sqrtf(s);
  }
}

extern "C" int trr_dense_search(trr_dense* h, const float* q, uint32_t B, uint32_t k, uint32_t* out_ord,
                                float* out_score, uint32_t* out_n) {
  if (!h || (B && (!q || !out_n)) || (B && k && (!out_ord || !out_score)))
    return trr_fail(TRR_ERR_INVALID_ARG, "trr_dense_search: NULL argument");
  if (B == 0) return TRR_OK;
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  trr_ctx* c = h->ctx;
  cudaStream_t st = c->stream;
  const size_t kk = std::max<uint32_t>(k, 1);
  const size_t q_bytes = (size_t)B * h->dim * 4;
  // outputs in one contiguous span (ord | score | n), so that a small search needs ONE device -> host copy
  const size_t ord_off = 0, score_off = ((size_t)B * kk * 4 + 255) & ~size_t(255);
  const size_t n_off = score_off + (((size_t)B * kk * 4 + 255) & ~size_t(255));
  const size_t out_bytes = n_off + (size_t)B * 4;
  TRR_CHECK(extra(c)->io.reserve(WsCarver::need({q_bytes + (size_t)B * 4, out_bytes})));
  WsCarver io(extra(c)->io.p);
  float* d_q = io.take<float>((size_t)B * h->dim + B);  // (+ B query norms behind the queries)
  uint8_t* d_out = io.take<uint8_t>(out_bytes);
  uint32_t* d_ord = reinterpret_cast<uint32_t*>(d_out + ord_off);
  float* d_score = reinterpret_cast<float*>(d_out + score_off);
  uint32_t* d_n = reinterpret_cast<uint32_t*>(d_out + n_off);
  // Small searches (the single-query call of VectorStore::search) go through the context's page-locked staging buffer:
  // a copy from / to pageable memory is a synchronous driver staging of its own, three of them cost more than the merge
  // kernel.  Large batches copy straight from the caller's buffers.
  const bool staged = q_bytes + out_bytes <= ((size_t)256 << 10);
  if (staged) {
    // |q| travels with the query: the reference's sequential f32 sum of squares + sqrt (src/index.rs:442) is a fraction of
    // a microsecond on the host, a dependent chain of `dim` steps in front of the scan on the device
    const size_t pin_q = (q_bytes + (size_t)B * 4 + 255) & ~size_t(255);
    TRR_CHECK(trr_ctx_reserve_pin(c, pin_q + out_bytes));
    uint8_t* pin = static_cast<uint8_t*>(c->pin);
    memcpy(pin, q, q_bytes);
    host_query_norms(q, h->dim, B, reinterpret_cast<float*>(pin + q_bytes));
    if (B == 1 && h->dim <= TRR_SCAN_PARAM_DIM && h->mode != TRR_DENSE_GEMM) {
      // VectorStore::search is ONE query: no copy operations at all.  The query and its norm travel in the kernel's
      // parameter space (every CTA reads them through the constant cache), and the last CTA writes the k results straight
      // into the page-locked staging buffer, which is mapped into the device's address space (unified addressing): a
      // kernel launch and a stream wait are all that is left around the scan (two copy operations cost ~15 us on a 240 us
      // kernel; letting 148 CTAs read the query from host memory instead cost 24 us).
      float* m_q = reinterpret_cast<float*>(pin);
      TRR_CHECK(dense_search_locked(h, m_q, B, k, reinterpret_cast<uint32_t*>(pin + pin_q + ord_off),
                                    reinterpret_cast<float*>(pin + pin_q + score_off),
                                    reinterpret_cast<uint32_t*>(pin + pin_q + n_off), false, m_q + (size_t)B * h->dim, m_q,
                                    m_q[(size_t)B * h->dim]));
    } else {
      TRR_CUDA(cudaMemcpyAsync(d_q, pin, q_bytes + (size_t)B * 4, cudaMemcpyHostToDevice, st));
      TRR_CHECK(dense_search_locked(h, d_q, B, k, d_ord, d_score, d_n, false, d_q + (size_t)B * h->dim));
      TRR_CUDA(cudaMemcpyAsync(pin + pin_q, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    }
    cudaError_t se = cudaStreamSynchronize(st);
    if (se != cudaSuccess) return trr_fail(TRR_ERR_CUDA, std::string("dense search failed: ") + cudaGetErrorString(se));
    if (k) {
      memcpy(out_ord, pin + pin_q + ord_off, (size_t)B * k * 4);
      memcpy(out_score, pin + pin_q + score_off, (size_t)B * k * 4);
    }
    memcpy(out_n, pin + pin_q + n_off, (size_t)B * 4);
    return TRR_OK;
  }
  TRR_CUDA(cudaMemcpyAsync(d_q, q, q_bytes, cudaMemcpyHostToDevice, st));
  TRR_CHECK(dense_search_locked(h, d_q, B, k, d_ord, d_score, d_n, true));
  if (k) {
    TRR_CUDA(cudaMemcpyAsync(out_ord, d_ord, (size_t)B * k * 4, cudaMemcpyDeviceToHost, st));
    TRR_CUDA(cudaMemcpyAsync(out_score, d_score, (size_t)B * k * 4, cudaMemcpyDeviceToHost, st));
  }
  TRR_CUDA(cudaMemcpyAsync(out_n, d_n, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  TRR_CUDA(cudaStreamSynchronize(st));
  return TRR_OK;
}

extern "C" int trr_dense_last_stats(trr_dense* h, trr_stats* out) {
  if (!h || !out) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  if (h->stats.mode_used) {
    cudaEventElapsedTime(&h->stats.ms_total, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->stats.ms_main_kernel, h->ev[2], h->ev[3]);
    cudaGetLastError();
  }
  dense_resolve_stats(h);
  *out = h->stats;
  return TRR_OK;
}

// debug / test hook: raw fast scores of the tensor-core pass for a small problem (B <= 128 * SMs, N small)
extern "C" TRR_API int trr_debug_gemm_scores(trr_dense* h, const float* q, uint32_t B, float* out, uint32_t out_ld) {
  if (!h || !q || !out) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  trr_ctx* c = h->ctx;
  cudaStream_t st = c->stream;
  TRR_CHECK(dense_freeze_locked(h));
  TRR_CHECK(dense_prepare_gemm(h));
  uint32_t n_qblocks = (B + 127) / 128;
  int pair_mode = 0;
  if (const char* e = TRR_KNOB("TRR_GEMM_PAIR")) pair_mode = atoi(e) ? 1 : 0;
  if (pair_mode) n_qblocks = (n_qblocks + 1) & ~1u;
  const uint32_t B_pad = n_qblocks * 128;
  const uint64_t n_pad = h->n_tiles * TRR_GEMM_TILE_N;
  if (out_ld < n_pad) return trr_fail(TRR_ERR_INVALID_ARG, "out_ld must be >= padded document count");
  uint32_t n_slices = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)c->sm_count / n_qblocks, (uint32_t)h->n_tiles));
  const size_t n_cand = (size_t)2 * n_slices * n_qblocks * 128 * TRR_GEMM_CP;
  const size_t need = WsCarver::need({(size_t)B * h->dim * 4, (size_t)B * 4, n_cand * 4, n_cand * 4, (size_t)B_pad * 4,
                                      (size_t)B_pad * h->dim_pad * 2, (size_t)B_pad * out_ld * 4});
  TRR_CHECK(extra(c)->scratch.reserve(need));
  WsCarver ws(extra(c)->scratch.p);
  float* d_q = ws.take<float>((size_t)B * h->dim);
  float* d_qdelta = ws.take<float>(B);
  float* cand_score = ws.take<float>(n_cand);
  uint32_t* cand_ord = ws.take<uint32_t>(n_cand);
  uint32_t* gthr = ws.take<uint32_t>(B_pad);
  uint16_t* q_bf16 = ws.take<uint16_t>((size_t)B_pad * h->dim_pad);
  float* dump = ws.take<float>((size_t)B_pad * out_ld);
  TRR_CUDA(cudaMemcpyAsync(d_q, q, (size_t)B * h->dim * 4, cudaMemcpyHostToDevice, st));
  trr_launch_query_prep(d_q, h->dim, h->dim_pad, B, B_pad, q_bf16, d_qdelta, nullptr, st);
  TRR_CUDA(cudaMemsetAsync(gthr, 0, (size_t)B_pad * 4, st));
  TRR_CUDA(cudaMemsetAsync(dump, 0, (size_t)B_pad * out_ld * 4, st));
  alignas(64) uint8_t map_q[128];
  TRR_CHECK(trr_make_tensor_map(map_q, q_bf16, B_pad, h->dim_pad, TRR_GEMM_TILE_M));
  GemmTopkArgs ga{};
  ga.n_qblocks = n_qblocks; ga.n_slices = n_slices; ga.n_tiles = (uint32_t)h->n_tiles;
  ga.k_blocks = (h->dim_pad + 63) / 64; ga.base_ord = h->base;
  ga.scale_bias = reinterpret_cast<const float2*>(h->scale_bias.p);
  ga.cand_score = cand_score; ga.cand_ord = cand_ord; ga.gthr = gthr; ga.share_thresholds = 0;
  ga.pair_mode = pair_mode; ga.cps = TRR_GEMM_CPS_MAX;
  ga.dbg = extra(c)->dbg_dev;
  TRR_CUDA(trr_launch_gemm_topk_dump(ga, map_q, pair_mode ? h->map_d_half : h->map_d, n_slices * n_qblocks, dump, out_ld,
                                     st));
  TRR_CUDA(cudaMemcpyAsync(out, dump, (size_t)B * out_ld * 4, cudaMemcpyDeviceToHost, st));
  TRR_CUDA(cudaStreamSynchronize(st));
  return TRR_OK;
}

// ------------------------------------------------------------------------------------------------
// BM25
// ------------------------------------------------------------------------------------------------
struct trr_bm25 {
  trr_ctx* ctx = nullptr;
  uint32_t n_docs = 0, n_terms = 0, doc_base = 0;
  uint64_t n_postings = 0;
  uint2* post = nullptr;
  uint32_t* skip = nullptr;
  uint32_t* term_min = nullptr;  // [n_terms + 1]: per-term minimum impact, then one flag word
  uint32_t* term_max = nullptr;  // [n_terms]: per-term maximum impact (scale of the integer fast pass)
  bool impacts_ok = true;        // every live impact is a finite value > 0 (the integer fast pass relies on it)
  uint32_t* stat_dev = nullptr;  // device copy of the number of queries the last search sent to the exact fallback
  bool stat_pending = false;
  // raw index kept for trr_bm25_append (every impact depends on the global N / df / avgdl, so an append re-weights all)
  uint64_t* d_term_off = nullptr;  // [n_terms + 1]
  uint32_t* tf = nullptr;          // [n_postings] term frequencies
  uint32_t* doc_len = nullptr;     // [n_docs]
  std::vector<uint64_t> h_term_off;
  uint64_t n_dead_postings = 0;    // postings of removed documents still in place (tf == 0)
  uint32_t range_shift = TRR_BM25_MAX_RANGE_SHIFT, n_ranges = 0, skip_ld = 0;
  trr_stats stats{};
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

static uint32_t bm25_pick_shift(uint32_t n_docs) {
  // documents per range: 32768 (128 KB of f32 accumulators in shared memory, one CTA per SM), fewer for small indexes
  uint32_t shift = TRR_BM25_MAX_RANGE_SHIFT;
  if (const char* e = TRR_KNOB("TRR_BM25_RANGE_SHIFT")) shift = (uint32_t)atoi(e);
  shift = std::min(std::max(shift, TRR_BM25_MIN_RANGE_SHIFT), TRR_BM25_MAX_RANGE_SHIFT);
  while (shift > TRR_BM25_MIN_RANGE_SHIFT && (1u << (shift - 1)) >= std::max<uint32_t>(n_docs, 1)) --shift;
  return shift;
}

#define BM_TRY(e)                                                                                          \
  do {                                                                                                     \
    cudaError_t _e = (e);                                                                                  \
    if (_e != cudaSuccess)                                                                                 \
      return trr_fail(_e == cudaErrorMemoryAllocation ? TRR_ERR_OOM : TRR_ERR_CUDA,                        \
                      std::string(#e) + ": " + cudaGetErrorString(_e));                                    \
  } while (0)

// (Re)computes everything derived from the raw index held by the handle — h->post[].x (doc ids, or d_post_doc when the
// ids still sit in a separate array), h->tf, h->doc_len, h->d_term_off — for the statistics given: impacts, skip table,
// per-term minimum impacts.  Allocates h->skip / h->term_min for the current n_terms / n_docs.
static int bm25_weight_locked(trr_bm25* h, const uint32_t* d_post_doc, float avgdl, float k1, float b, const float* idf_host) {
  trr_ctx* ctx = h->ctx;
  cudaStream_t st = ctx->stream;
  h->range_shift = bm25_pick_shift(h->n_docs);
  h->n_ranges = h->n_docs ? (uint32_t)(((uint64_t)h->n_docs + (1u << h->range_shift) - 1) >> h->range_shift) : 0;
  h->skip_ld = h->n_ranges + 1;
  if (h->skip) { cudaFree(h->skip); h->skip = nullptr; }
  if (h->term_min) { cudaFree(h->term_min); h->term_min = nullptr; }
  if (h->term_max) { cudaFree(h->term_max); h->term_max = nullptr; }
  BM_TRY(cudaMalloc(&h->skip, std::max<uint64_t>((uint64_t)h->n_terms * h->skip_ld, 1) * 4));
  BM_TRY(cudaMalloc(&h->term_min, ((uint64_t)h->n_terms + 1) * 4));
  BM_TRY(cudaMalloc(&h->term_max, std::max<uint64_t>(h->n_terms, 1) * 4));
  BM_TRY(cudaMemsetAsync(h->term_min, 0xFF, (uint64_t)h->n_terms * 4, st));
  BM_TRY(cudaMemsetAsync(h->term_min + h->n_terms, 0, 4, st));
  BM_TRY(cudaMemsetAsync(h->term_max, 0, std::max<uint64_t>(h->n_terms, 1) * 4, st));
  float* d_idf = nullptr;
  BM_TRY(cudaMalloc(&d_idf, std::max<uint64_t>(h->n_terms, 1) * 4));
  if (h->n_terms && idf_host) BM_TRY(cudaMemcpyAsync(d_idf, idf_host, (uint64_t)h->n_terms * 4, cudaMemcpyHostToDevice, st));
  Bm25BuildArgs a{};
  a.n_postings = h->n_postings; a.n_terms = h->n_terms; a.n_docs = h->n_docs; a.term_off = h->d_term_off;
  a.post_doc = d_post_doc; a.post_tf = h->tf; a.doc_len = h->doc_len; a.idf = d_idf; a.avgdl = avgdl; a.k1 = k1; a.b = b;
  a.range_shift = h->range_shift; a.n_ranges = h->n_ranges; a.skip_ld = h->skip_ld; a.post = h->post; a.skip = h->skip;
  a.term_min = h->term_min; a.term_max = h->term_max; a.flags = h->term_min + h->n_terms;
  cudaError_t e = trr_launch_bm25_build(a, st);
  ctx->launches += 2;
  uint32_t flag_word = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flag_word, h->term_min + h->n_terms, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  h->impacts_ok = flag_word == 0;
  // dead postings make the posting count of a term an over-estimate of its live documents: no threshold bootstrap then
  if (e == cudaSuccess && h->n_dead_postings) {
    flag_word |= 2u;
    e = cudaMemcpy(h->term_min + h->n_terms, &flag_word, 4, cudaMemcpyHostToDevice);
  }
  cudaFree(d_idf);
  BM_TRY(e);
  return TRR_OK;
}

extern "C" int trr_bm25_build(trr_ctx* ctx, uint32_t n_docs, uint32_t n_terms, const uint64_t* term_off,
                              const uint32_t* post_doc, const uint32_t* post_tf, const uint32_t* doc_len, float avgdl,
                              float k1, float b, const float* idf, uint32_t doc_base, trr_bm25** out) {
  if (!ctx || !out || !term_off) return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_build: NULL argument");
  *out = nullptr;
  const uint64_t P = term_off[n_terms];
  if (P && (!post_doc || !post_tf || !doc_len || !idf)) return trr_fail(TRR_ERR_INVALID_ARG, "NULL postings");
  if (P >= 0xFFFFFFFFull) return trr_fail(TRR_ERR_UNSUPPORTED, "more than 2^32-1 postings per shard");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  trr_bm25* h = new trr_bm25();
  h->ctx = ctx; h->n_docs = n_docs; h->n_terms = n_terms; h->doc_base = doc_base; h->n_postings = P;
  for (auto& e : h->ev) cudaEventCreate(&e);
  cudaStream_t st = ctx->stream;
  uint32_t* d_pd = nullptr;
  auto body = [&]() -> int {
    BM_TRY(cudaMalloc(&h->post, (P + 2) * sizeof(uint2)));  // +2: 16-byte aligned bulk copies may read one posting past the end
    BM_TRY(cudaMemsetAsync(h->post, 0xFF, (P + 2) * sizeof(uint2), st));
    BM_TRY(cudaMalloc(&h->d_term_off, ((uint64_t)n_terms + 1) * 8));
    BM_TRY(cudaMalloc(&h->tf, std::max<uint64_t>(P, 1) * 4));
    BM_TRY(cudaMalloc(&h->doc_len, std::max<uint64_t>(n_docs, 1) * 4));
    BM_TRY(cudaMalloc(&d_pd, std::max<uint64_t>(P, 1) * 4));
    BM_TRY(cudaMemcpyAsync(h->d_term_off, term_off, ((uint64_t)n_terms + 1) * 8, cudaMemcpyHostToDevice, st));
    if (P) {
      BM_TRY(cudaMemcpyAsync(d_pd, post_doc, P * 4, cudaMemcpyHostToDevice, st));
      BM_TRY(cudaMemcpyAsync(h->tf, post_tf, P * 4, cudaMemcpyHostToDevice, st));
    }
    if (n_docs) BM_TRY(cudaMemcpyAsync(h->doc_len, doc_len, (uint64_t)n_docs * 4, cudaMemcpyHostToDevice, st));
    h->h_term_off.assign(term_off, term_off + n_terms + 1);
    return bm25_weight_locked(h, d_pd, avgdl, k1, b, idf);
  };
  const int s = body();
  if (d_pd) cudaFree(d_pd);
  if (s != TRR_OK) { trr_bm25_destroy(h); return s; }
  *out = h;
  return TRR_OK;
}

// BM25Index::add for documents arriving after the index was built (reference src/index.rs:176-204).  Every impact
// depends on the global N, df and avgdl, so an append re-weights the whole index; what it saves is the host-side merge and
// the upload of the existing postings: only the CSR of the NEW documents crosses PCIe, the old and new postings are merged
// per term on the device, and the weighting pass (the same kernel as trr_bm25_build) runs at HBM speed.
extern "C" int trr_bm25_append(trr_bm25* h, uint32_t n_new_docs, uint32_t n_terms_new, const uint64_t* delta_term_off,
                               const uint32_t* delta_post_doc, const uint32_t* delta_post_tf, const uint32_t* delta_doc_len,
                               float avgdl, float k1, float b, const float* idf) {
  if (!h || !delta_term_off || (n_terms_new && !idf)) return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_append: NULL argument");
  if (n_terms_new < h->n_terms) return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_append: the vocabulary cannot shrink");
  if (!h->tf || !h->d_term_off || h->h_term_off.size() != (size_t)h->n_terms + 1)
    return trr_fail(TRR_ERR_UNSUPPORTED, "trr_bm25_append: this index holds no raw postings");
  const uint64_t dP = delta_term_off[n_terms_new];
  if (dP && (!delta_post_doc || !delta_post_tf)) return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_append: NULL postings");
  if (n_new_docs && !delta_doc_len) return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_append: NULL doc lengths");
  const uint64_t P_new = h->n_postings + dP;
  if (P_new >= 0xFFFFFFFFull || (uint64_t)h->n_docs + n_new_docs >= 0xFFFFFFFFull)
    return trr_fail(TRR_ERR_UNSUPPORTED, "more than 2^32-1 postings or documents per shard");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  cudaStream_t st = h->ctx->stream;
  // new per-term offsets (host prefix sum over old length + appended length)
  std::vector<uint64_t> new_off((size_t)n_terms_new + 1, 0);
  for (uint32_t t = 0; t < n_terms_new; ++t) {
    const uint64_t old_len = t < h->n_terms ? h->h_term_off[t + 1] - h->h_term_off[t] : 0;
    new_off[t + 1] = new_off[t] + old_len + (delta_term_off[t + 1] - delta_term_off[t]);
  }
  uint64_t *d_new_off = nullptr, *d_delta_off = nullptr;
  uint32_t *d_dd = nullptr, *d_dtf = nullptr, *new_tf = nullptr, *new_dl = nullptr;
  uint2* new_post = nullptr;
  auto body = [&]() -> int {
    BM_TRY(cudaMalloc(&d_new_off, ((size_t)n_terms_new + 1) * 8));
    BM_TRY(cudaMalloc(&d_delta_off, ((size_t)n_terms_new + 1) * 8));
    BM_TRY(cudaMalloc(&d_dd, std::max<uint64_t>(dP, 1) * 4));
    BM_TRY(cudaMalloc(&d_dtf, std::max<uint64_t>(dP, 1) * 4));
    BM_TRY(cudaMalloc(&new_post, (P_new + 2) * sizeof(uint2)));
    BM_TRY(cudaMalloc(&new_tf, std::max<uint64_t>(P_new, 1) * 4));
    BM_TRY(cudaMalloc(&new_dl, std::max<uint64_t>((uint64_t)h->n_docs + n_new_docs, 1) * 4));
    BM_TRY(cudaMemsetAsync(new_post, 0xFF, (P_new + 2) * sizeof(uint2), st));
    BM_TRY(cudaMemcpyAsync(d_new_off, new_off.data(), new_off.size() * 8, cudaMemcpyHostToDevice, st));
    BM_TRY(cudaMemcpyAsync(d_delta_off, delta_term_off, ((size_t)n_terms_new + 1) * 8, cudaMemcpyHostToDevice, st));
    if (dP) {
      BM_TRY(cudaMemcpyAsync(d_dd, delta_post_doc, dP * 4, cudaMemcpyHostToDevice, st));
      BM_TRY(cudaMemcpyAsync(d_dtf, delta_post_tf, dP * 4, cudaMemcpyHostToDevice, st));
    }
    if (h->n_docs) BM_TRY(cudaMemcpyAsync(new_dl, h->doc_len, (uint64_t)h->n_docs * 4, cudaMemcpyDeviceToDevice, st));
    if (n_new_docs) BM_TRY(cudaMemcpyAsync(new_dl + h->n_docs, delta_doc_len, (uint64_t)n_new_docs * 4, cudaMemcpyHostToDevice, st));
    Bm25MergeArgs m{};
    m.n_postings_new = P_new; m.n_terms_new = n_terms_new; m.n_terms_old = h->n_terms; m.n_docs_old = h->n_docs;
    m.new_off = d_new_off; m.old_off = h->d_term_off; m.delta_off = d_delta_off; m.old_post = h->post; m.old_tf = h->tf;
    m.delta_doc = d_dd; m.delta_tf = d_dtf; m.new_post = new_post; m.new_tf = new_tf;
    BM_TRY(trr_launch_bm25_merge(m, st));
    h->ctx->launches++;
    BM_TRY(cudaStreamSynchronize(st));
    // swap the raw index, then re-weight
    cudaFree(h->post); cudaFree(h->tf); cudaFree(h->doc_len); cudaFree(h->d_term_off);
    h->post = new_post; h->tf = new_tf; h->doc_len = new_dl; h->d_term_off = d_new_off;
    new_post = nullptr; new_tf = nullptr; new_dl = nullptr; d_new_off = nullptr;
    h->n_docs += n_new_docs; h->n_terms = n_terms_new; h->n_postings = P_new;
    h->h_term_off.swap(new_off);
    return bm25_weight_locked(h, nullptr, avgdl, k1, b, idf);
  };
  const int s = body();
  for (void* p : {(void*)d_new_off, (void*)d_delta_off, (void*)d_dd, (void*)d_dtf, (void*)new_post, (void*)new_tf, (void*)new_dl})
    if (p) cudaFree(p);
  return s;
}
// BM25Index::remove (reference src/index.rs:245-275) without a rebuild: the postings of the removed documents stay in
// place with term frequency 0 (they weigh +0.0 and the document is dropped by `score > 0.0`), and every other posting is
// re-weighted with the new N / df / avgdl supplied by the host.  out_dead_postings (nullable) reports how many postings
// are dead in total, so that the host can decide when a compacting rebuild pays off.
extern "C" int trr_bm25_remove(trr_bm25* h, const uint32_t* ordinals, uint32_t n, float avgdl, float k1, float b,
                               const float* idf, uint64_t* out_dead_postings) {
  if (!h || (n && !ordinals) || (h->n_terms && !idf)) return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_remove: NULL argument");
  if (!h->tf) return trr_fail(TRR_ERR_UNSUPPORTED, "trr_bm25_remove: this index holds no raw postings");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  cudaStream_t st = h->ctx->stream;
  const size_t words = ((size_t)h->n_docs + 31) / 32 + 1;
  std::vector<uint32_t> bits(words, 0);
  for (uint32_t i = 0; i < n; ++i) {
    if (ordinals[i] >= h->n_docs) return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_remove: ordinal out of range");
    bits[ordinals[i] >> 5] |= 1u << (ordinals[i] & 31);
  }
  uint32_t* d_bits = nullptr;
  auto body = [&]() -> int {
    BM_TRY(cudaMalloc(&d_bits, (words + 1) * 4));
    BM_TRY(cudaMemcpyAsync(d_bits, bits.data(), words * 4, cudaMemcpyHostToDevice, st));
    BM_TRY(cudaMemsetAsync(d_bits + words, 0, 4, st));
    BM_TRY(trr_launch_bm25_kill(h->post, h->tf, h->n_postings, d_bits, d_bits + words, st));
    h->ctx->launches++;
    uint32_t killed = 0;
    BM_TRY(cudaMemcpyAsync(&killed, d_bits + words, 4, cudaMemcpyDeviceToHost, st));
    BM_TRY(cudaStreamSynchronize(st));
    h->n_dead_postings += killed;
    return bm25_weight_locked(h, nullptr, avgdl, k1, b, idf);
  };
  const int s = body();
  if (d_bits) cudaFree(d_bits);
  if (out_dead_postings) *out_dead_postings = h->n_dead_postings;
  return s;
}
#undef BM_TRY

extern "C" int trr_bm25_destroy(trr_bm25* h) {
  if (!h) return TRR_OK;
  DeviceGuard g(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  if (h->post) cudaFree(h->post);
  if (h->skip) cudaFree(h->skip);
  if (h->term_min) cudaFree(h->term_min);
  if (h->term_max) cudaFree(h->term_max);
  if (h->stat_dev) cudaFree(h->stat_dev);
  if (h->d_term_off) cudaFree(h->d_term_off);
  if (h->tf) cudaFree(h->tf);
  if (h->doc_len) cudaFree(h->doc_len);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  delete h;
  return TRR_OK;
}

extern "C" int trr_bm25_n_postings(trr_bm25* h, uint64_t* out) {
  if (!h || !out) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  *out = h->n_postings;
  return TRR_OK;
}

extern "C" int trr_bm25_copy_impacts(trr_bm25* h, float* out, uint64_t n) {
  if (!h || !out) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  n = std::min(n, h->n_postings);
  std::vector<uint2> tmp(n);
  TRR_CUDA(cudaMemcpy(tmp.data(), h->post, n * sizeof(uint2), cudaMemcpyDeviceToHost));
  for (uint64_t i = 0; i < n; ++i) memcpy(&out[i], &tmp[i].y, 4);
  return TRR_OK;
}

// stage capacity (postings) of a BM25 search kernel: what is left of the shared memory after the fixed part, split over
// n_stages buffers of 8 bytes per posting; ranges of <= 16K documents leave room for two CTAs per SM
#ifdef TRR_TRIAGE
static uint32_t* g_triage_out = nullptr;
#endif
static int bm25_stage_cap(trr_ctx* c, size_t fixed, uint32_t n_stages, uint32_t* ctas_per_sm, uint32_t* stage_cap) {
  size_t budget = c->smem_optin;
  if (*ctas_per_sm == 2) budget = (228 * 1024) / 2 - 1024 - 1024;  // 228 KB per SM, 1 KB reserved per CTA, 1 KB slack
  if (fixed + (size_t)n_stages * 8 * 256 > budget) { *ctas_per_sm = 1; budget = c->smem_optin; }
  if (fixed + (size_t)n_stages * 8 * 256 > budget) return trr_fail(TRR_ERR_UNSUPPORTED, "BM25 kernel does not fit shared memory");
  size_t cap = (budget - fixed) / ((size_t)8 * n_stages);
  cap = std::min<size_t>(cap, 8192) & ~size_t(1);
  *stage_cap = (uint32_t)cap;
  return TRR_OK;
}

static int bm25_search_locked(trr_bm25* h, const uint32_t* d_q_terms, const uint32_t* d_q_off, const uint32_t* h_q_off,
                              uint32_t B, uint32_t k, uint32_t* d_ord, float* d_score, uint32_t* d_n, size_t scratch_off) {
  TrrRange nvtx_range("trr:bm25_search");
  trr_ctx* c = h->ctx;
  cudaStream_t st = c->stream;
  h->stats = trr_stats{};
  h->stats.n_queries = B;
  if (B == 0) return TRR_OK;
  if (k == 0 || h->n_postings == 0) {
    TRR_CUDA(cudaMemsetAsync(d_n, 0, (size_t)B * 4, st));
    return TRR_OK;
  }
  if (k > 1024) return trr_fail(TRR_ERR_UNSUPPORTED, "k > 1024 is not supported yet");
  for (uint32_t b = 0; b < B; ++b)
    if (h_q_off[b + 1] - h_q_off[b] > TRR_BM25_MAX_QUERY_TERMS)
      return trr_fail(TRR_ERR_UNSUPPORTED, "more than 512 terms in one query");
  // Default path: integer fast pass with 16-bit cells (selection) + exact re-scoring with a proof; queries whose proof fails
  // repeat both with 32-bit cells, and what still fails goes to the exact kernel.  The fallback levels are device-driven and
  // split every flagged query into chunks of document ranges, so that a single failure does not serialise on one CTA.
  // The exact kernel alone serves indexes whose impacts are not all finite and positive (the fast pass's bounds rely on it).
  const bool fast = h->impacts_ok && !TRR_KNOB("TRR_BM25_EXACT");
  const uint32_t kf = k + std::max<uint32_t>(64, k / 2);  // candidates kept by the fast pass: k plus a margin for the proof
  Bm25SearchArgs a{};
  a.post = h->post; a.skip = h->skip; a.skip_ld = h->skip_ld; a.n_terms = h->n_terms; a.n_docs = h->n_docs;
  a.n_ranges = h->n_ranges; a.range_shift = h->range_shift; a.doc_base = h->doc_base;
  a.q_terms = d_q_terms; a.q_off = d_q_off; a.B = B; a.k = k; a.kf = kf;
  const uint32_t cand_exact = trr_pow2_ceil(k + 512), cand_fast = trr_pow2_ceil(kf + 512);
  uint32_t ctas_exact = a.range_shift <= 14 ? 2u : 1u, ctas16 = 1, ctas32 = 1;  // (the fast kernels fill an SM with one CTA)
  if (const char* e = TRR_KNOB("TRR_BM25_CTAS_PER_SM")) ctas_exact = std::min(2, std::max(1, atoi(e)));
  uint32_t stage_exact = 0, stage16 = 0, stage32 = 0;
  // accumulators of the fast pass: a ring of two wherever two fit next to the stages (16-bit cells always, 32-bit cells up
  // to 16K documents per range), so that one range is scanned while the next accumulates
  uint32_t rmul16 = a.range_shift == TRR_BM25_MAX_RANGE_SHIFT ? TRR_BM25_FAST_RMUL16 : 0u;
  if (const char* e = TRR_KNOB("TRR_BM25_RMUL")) rmul16 = (uint32_t)std::min(1, std::max(0, atoi(e)));
  const uint32_t nacc16 = rmul16 ? 1u : 2u, nacc32 = a.range_shift <= 14 ? 2u : 1u;
  TRR_CHECK(bm25_stage_cap(c, trr_bm25_search_smem(a.range_shift, 0, cand_exact) + 64, 2, &ctas_exact, &stage_exact));
  if (fast) {
    TRR_CHECK(bm25_stage_cap(c, trr_bm25_fast_smem(16, a.range_shift + rmul16, nacc16, 0, cand_fast) + 128, TRR_BM25_FAST_STAGES, &ctas16, &stage16));
    TRR_CHECK(bm25_stage_cap(c, trr_bm25_fast_smem(32, a.range_shift, nacc32, 0, cand_fast) + 128, TRR_BM25_FAST_STAGES, &ctas32, &stage32));
  }
  // few queries: split every query into chunks of document ranges so that all SMs have work
  const uint32_t slots = (uint32_t)c->sm_count * (fast ? ctas16 : ctas_exact);
  uint32_t n_chunks = 1;
  if (B < 2u * slots) n_chunks = std::min<uint32_t>(h->n_ranges, (2u * slots + B - 1) / B);
  if (const char* e = TRR_KNOB("TRR_BM25_CHUNKS")) n_chunks = std::min<uint32_t>(h->n_ranges, std::max(1, atoi(e)));
  n_chunks = std::max<uint32_t>(n_chunks, 1);
  a.n_chunks = n_chunks;
  const uint32_t nc_fb = std::max<uint32_t>(n_chunks, std::min<uint32_t>(h->n_ranges, 32));  // chunks per query of the fallback levels
  const size_t k_list = fast ? kf : k;
  const size_t list_items = (size_t)B * std::max(n_chunks, fast ? nc_fb : 1u);
  const size_t need = scratch_off + WsCarver::need({(size_t)B * 4, (size_t)B * 8, 64, list_items * k_list * 8 + 8,
                                                    (size_t)trr_pow2_ceil(B) * 8, (size_t)B * 16, (size_t)B * kf * 8 + 8,
                                                    (size_t)B * 8});
  TRR_CHECK(extra(c)->scratch.reserve(need));
  WsCarver ws(static_cast<char*>(extra(c)->scratch.p) + scratch_off);
  a.term_min = h->term_min; a.term_max = h->term_max; a.flags = h->term_min + h->n_terms;
  a.order = ws.take<uint32_t>(B);
  a.thr0 = ws.take<uint64_t>(B);
  a.queue = ws.take<uint32_t>(16);  // [0..2] work queues of the three levels, [3] / [4] queries flagged by the 16- / 32-bit proof
  uint64_t* lists = ws.take<uint64_t>(list_items * k_list + 1);  // per-chunk lists of the level that is running
  uint64_t* plan_keys = ws.take<uint64_t>(trr_pow2_ceil(B));
  uint32_t* per_query = ws.take<uint32_t>((size_t)B * 4);
  uint64_t* merged = ws.take<uint64_t>((size_t)B * kf + 1);
  uint32_t* flagged1 = ws.take<uint32_t>((size_t)B * 2);
  uint32_t* flagged2 = flagged1 + B;
  a.dbg = extra(c)->dbg_dev;
  if (const char* e = TRR_KNOB("TRR_BM25_TRIAGE")) a.triage = (uint32_t)atoi(e);
#ifdef TRR_TRIAGE
  if (!g_triage_out) { cudaMalloc(&g_triage_out, 256); }
  cudaMemsetAsync(g_triage_out, 0, 256, st);
  a.triage_out = g_triage_out;
#endif
  uint32_t launches = (B > 1 && B <= 4096) ? 2u : 1u;  // plan
  auto merge = [&](const uint64_t* src, uint32_t n_lists, uint32_t stride, uint32_t kk, const uint32_t* n_rows_ptr,
                   const uint32_t* row_map, uint64_t* out_keys, bool outputs) -> int {
    TopkMergeArgs m{};
    m.lists = src; m.list_n = nullptr; m.n_lists = n_lists; m.list_stride = stride;
    m.n_rows = B; m.n_rows_ptr = n_rows_ptr; m.row_map = row_map; m.k = kk; m.k2 = trr_pow2_ceil(kk);
    m.out_keys = out_keys;
    if (outputs) { m.out_ord = d_ord; m.out_score = d_score; m.out_n = d_n; }
    TRR_CUDA(trr_launch_topk_merge(m, n_rows_ptr ? std::min<uint32_t>(B, 4u * (uint32_t)c->sm_count) : B, st));
    ++launches;
    return TRR_OK;
  };
  // the exact kernel over the queries listed in sel (count on the device, or all B in plan order)
  auto run_exact = [&](const uint32_t* sel, const uint32_t* n_sel_ptr, uint32_t* queue, uint32_t nc) -> int {
    Bm25SearchArgs x = a;
    if (sel) x.order = const_cast<uint32_t*>(sel);
    x.n_sel_ptr = n_sel_ptr; x.queue = queue; x.n_chunks = nc;
    x.stage_cap = stage_exact; x.cand_cap = cand_exact; x.partial = lists;
    x.out_keys = nullptr; x.out_ord = d_ord; x.out_score = d_score; x.out_n = d_n;
    const unsigned grid_x = (unsigned)std::min<uint64_t>((uint64_t)B * nc, (uint64_t)c->sm_count * ctas_exact);
    TRR_CUDA(trr_launch_bm25_search(x, grid_x, st));
    ++launches;
    if (nc > 1) TRR_CHECK(merge(lists, nc, k, k, n_sel_ptr, x.order, nullptr, true));
    return TRR_OK;
  };
  if (!fast) {
    TRR_CUDA(trr_launch_bm25_plan(a, plan_keys, st));
    TRR_CUDA(cudaEventRecord(h->ev[2], st));
    TRR_CHECK(run_exact(nullptr, nullptr, a.queue, n_chunks));
    TRR_CUDA(cudaEventRecord(h->ev[3], st));
    h->stats.mode_used = 1;
  } else {
    a.qscale16 = reinterpret_cast<float*>(per_query); a.thr0f16 = per_query + B;
    a.qscale32 = reinterpret_cast<float*>(per_query + 2 * (size_t)B); a.thr0f32 = per_query + 3 * (size_t)B;
    TRR_CUDA(trr_launch_bm25_plan(a, plan_keys, st));
    TRR_CUDA(cudaEventRecord(h->ev[2], st));
    uint32_t margin = 0;
    if (const char* e = TRR_KNOB("TRR_BM25_PROOF_MARGIN")) margin = (uint32_t)atoi(e);  // tests: force the fallback levels
    // one level of the fast pass: fast kernel -> (merge of the chunk lists) -> exact re-scoring + proof
    auto level = [&](int bits, const uint32_t* sel, const uint32_t* n_sel_ptr, uint32_t* queue, uint32_t nc, uint32_t ctas,
                     uint32_t stage_cap, uint32_t* flagged, uint32_t* n_flagged, uint32_t margin_f) -> int {
      Bm25SearchArgs x = a;
      if (sel) x.order = const_cast<uint32_t*>(sel);
      x.n_sel_ptr = n_sel_ptr; x.queue = queue; x.n_chunks = nc;
      x.qscale = bits == 16 ? a.qscale16 : a.qscale32; x.thr0f = bits == 16 ? a.thr0f16 : a.thr0f32;
      x.fast_keys = lists; x.stage_cap = stage_cap; x.cand_cap = cand_fast; x.rmul_shift = bits == 16 ? rmul16 : 0u; x.n_acc = bits == 16 ? nacc16 : nacc32;
      const unsigned grid_x = (unsigned)std::min<uint64_t>((uint64_t)B * nc, (uint64_t)c->sm_count * ctas);
      TRR_CUDA(trr_launch_bm25_fast(x, bits, grid_x, st));
      ++launches;
      const uint64_t* fast_keys = lists;
      if (nc > 1) {
        TRR_CHECK(merge(lists, nc, kf, kf, n_sel_ptr, nullptr, merged, false));
        fast_keys = merged;
      }
      Bm25RescoreArgs r{};
      r.post = h->post; r.skip = h->skip; r.skip_ld = h->skip_ld; r.n_terms = h->n_terms; r.range_shift = h->range_shift;
      r.doc_base = h->doc_base; r.q_terms = d_q_terms; r.q_off = d_q_off; r.B = B; r.k = k; r.kf = kf;
      r.cap2 = trr_pow2_ceil(kf); r.sel = x.order; r.sel_n = n_sel_ptr; r.fast_keys = fast_keys; r.qscale = x.qscale;
      r.out_ord = d_ord; r.out_score = d_score; r.out_n = d_n; r.flagged = flagged; r.n_flagged = n_flagged;
      r.margin_f = margin_f;
      TRR_CUDA(trr_launch_bm25_rescore(r, st));
      ++launches;
      return TRR_OK;
    };
    TRR_CHECK(level(16, nullptr, nullptr, a.queue, n_chunks, ctas16, stage16, flagged1, a.queue + 3, margin));
    TRR_CHECK(level(32, flagged1, a.queue + 3, a.queue + 1, nc_fb, ctas32, stage32, flagged2, a.queue + 4, margin > 1 ? margin : 0));
    TRR_CHECK(run_exact(flagged2, a.queue + 4, a.queue + 2, nc_fb));
    TRR_CUDA(cudaEventRecord(h->ev[3], st));
    if (!h->stat_dev) TRR_CUDA(cudaMalloc(&h->stat_dev, 64));
    TRR_CUDA(cudaMemcpyAsync(h->stat_dev, a.queue + 3, 8, cudaMemcpyDeviceToDevice, st));
    h->stat_pending = true;
    h->stats.mode_used = 2;
  }
  c->launches += launches;
  h->stats.n_kernel_launches = launches;
  return TRR_OK;
}

extern "C" int trr_bm25_search_device(trr_bm25* h, const uint32_t* d_q_terms, const uint32_t* d_q_off,
                                      const uint32_t* h_q_off, uint32_t B, uint32_t k, uint32_t* d_ord, float* d_score,
                                      uint32_t* d_n) {
  if (!h || (B && (!d_q_off || !h_q_off || !d_ord || !d_score || !d_n)))
    return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  return bm25_search_locked(h, d_q_terms, d_q_off, h_q_off, B, k, d_ord, d_score, d_n, 0);
}

extern "C" int trr_bm25_search(trr_bm25* h, const uint32_t* q_terms, const uint32_t* q_off, uint32_t B, uint32_t k,
                               uint32_t* out_ord, float* out_score, uint32_t* out_n) {
  if (!h || (B && (!q_off || !out_n)) || (B && k && (!out_ord || !out_score)))
    return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_search: NULL argument");
  if (B == 0) return TRR_OK;
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  trr_ctx* c = h->ctx;
  cudaStream_t st = c->stream;
  const size_t nt = q_off[B], kk = std::max<uint32_t>(k, 1);
  const size_t need = WsCarver::need({std::max<size_t>(nt, 1) * 4, (size_t)(B + 1) * 4, (size_t)B * kk * 4,
                                      (size_t)B * kk * 4, (size_t)B * 4});
  TRR_CHECK(extra(c)->io.reserve(need));
  WsCarver io(extra(c)->io.p);
  uint32_t* d_terms = io.take<uint32_t>(std::max<size_t>(nt, 1));
  uint32_t* d_off = io.take<uint32_t>(B + 1);
  uint32_t* d_ord = io.take<uint32_t>((size_t)B * kk);
  float* d_score = io.take<float>((size_t)B * kk);
  uint32_t* d_n = io.take<uint32_t>(B);
  if (nt) TRR_CUDA(cudaMemcpyAsync(d_terms, q_terms, nt * 4, cudaMemcpyHostToDevice, st));
  TRR_CUDA(cudaMemcpyAsync(d_off, q_off, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
  TRR_CUDA(cudaEventRecord(h->ev[0], st));
  TRR_CHECK(bm25_search_locked(h, d_terms, d_off, q_off, B, k, d_ord, d_score, d_n, 0));
  TRR_CUDA(cudaEventRecord(h->ev[1], st));
  if (k) {
    TRR_CUDA(cudaMemcpyAsync(out_ord, d_ord, (size_t)B * k * 4, cudaMemcpyDeviceToHost, st));
    TRR_CUDA(cudaMemcpyAsync(out_score, d_score, (size_t)B * k * 4, cudaMemcpyDeviceToHost, st));
  }
  TRR_CUDA(cudaMemcpyAsync(out_n, d_n, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  TRR_CUDA(cudaStreamSynchronize(st));
  if (h->stats.mode_used) {
    cudaEventElapsedTime(&h->stats.ms_total, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->stats.ms_main_kernel, h->ev[2], h->ev[3]);
  }
  return TRR_OK;
}

extern "C" int trr_bm25_last_stats(trr_bm25* h, trr_stats* out) {
  if (!h || !out) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  if (h->stats.mode_used) {
    cudaEventElapsedTime(&h->stats.ms_main_kernel, h->ev[2], h->ev[3]);
    cudaGetLastError();
  }
#ifdef TRR_TRIAGE
  if (h->stats.mode_used == 2 && g_triage_out) {
    uint32_t pw[64] = {0};
    cudaMemcpy(pw, g_triage_out, 256, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[trr] K3 fast, CTA 0 (x16 cycles): accumulate warp 0: total %u = wait postings %u + walk %u + wait slot %u (%u passes); "
                    "scan warp 0: total %u = wait accumulator %u + scan %u + settle %u + end of item %u (%u hand-overs, %u compactions); "
                    "producer: total %u, waited %u for a free stage (%u passes)\n",
            pw[0], pw[1], pw[2], pw[3], pw[4], pw[8], pw[9], pw[10], pw[11], pw[12], pw[13], pw[14], pw[16], pw[17], pw[18]);
  }
#endif
  if (h->stat_pending && h->stat_dev) {
    uint32_t nf[2] = {0, 0};  // queries re-run with 32-bit cells / by the exact kernel
    if (cudaMemcpy(nf, h->stat_dev, 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
      h->stats.n_guard_fallbacks = nf[0];
      h->stats.n_exact_fallbacks = nf[1];
    }
    h->stat_pending = false;
  }
  *out = h->stats;
  return TRR_OK;
}

// ------------------------------------------------------------------------------------------------
// snapshots: device structures <-> flat little-endian files (chunked through pinned staging)
// ------------------------------------------------------------------------------------------------
namespace {
struct FileCloser {
  FILE* f;
  ~FileCloser() { if (f) fclose(f); }
};
constexpr size_t SNAP_CHUNK = (size_t)64 << 20;

int snap_write_dev(trr_ctx* c, FILE* f, const void* d_src, size_t bytes) {
  TRR_CHECK(trr_ctx_reserve_pin(c, SNAP_CHUNK));
  for (size_t off = 0; off < bytes; off += SNAP_CHUNK) {
    const size_t n = std::min(SNAP_CHUNK, bytes - off);
    TRR_CUDA(cudaMemcpy(c->pin, static_cast<const char*>(d_src) + off, n, cudaMemcpyDeviceToHost));
    if (fwrite(c->pin, 1, n, f) != n) return trr_fail(TRR_ERR_INVALID_ARG, "snapshot: short write");
  }
  return TRR_OK;
}
int snap_read_dev(trr_ctx* c, FILE* f, void* d_dst, size_t bytes) {
  TRR_CHECK(trr_ctx_reserve_pin(c, SNAP_CHUNK));
  for (size_t off = 0; off < bytes; off += SNAP_CHUNK) {
    const size_t n = std::min(SNAP_CHUNK, bytes - off);
    if (fread(c->pin, 1, n, f) != n) return trr_fail(TRR_ERR_INVALID_ARG, "snapshot: truncated file");
    TRR_CUDA(cudaMemcpy(static_cast<char*>(d_dst) + off, c->pin, n, cudaMemcpyHostToDevice));
  }
  return TRR_OK;
}
struct DenseSnapHeader {
  char magic[8];  // "TRRDNS01"
  uint32_t dim; int32_t metric, dtype; uint32_t base;
  uint64_t n, n_dead;
};
struct Bm25SnapHeader {
  char magic[8];  // "TRRBM253"
  uint32_t n_docs, n_terms, doc_base, range_shift, n_ranges, skip_ld;
  uint64_t n_postings;
  uint64_t n_dead_postings;  // postings of removed documents still in place (tf == 0)
  uint32_t impacts_ok, reserved;
};
}  // namespace

extern "C" int trr_dense_info(trr_dense* h, uint32_t* out_dim, int* out_metric, int* out_dtype, uint32_t* out_base) {
  if (!h) return trr_fail(TRR_ERR_INVALID_ARG, "handle is NULL");
  if (out_dim) *out_dim = h->dim;
  if (out_metric) *out_metric = h->metric;
  if (out_dtype) *out_dtype = h->dtype;
  if (out_base) *out_base = h->base;
  return TRR_OK;
}

extern "C" int trr_dense_save(trr_dense* h, const char* path) {
  if (!h || !path) return trr_fail(TRR_ERR_INVALID_ARG, "trr_dense_save: NULL argument");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  TRR_CUDA(cudaStreamSynchronize(h->ctx->stream));
  FileCloser fc{fopen(path, "wb")};
  if (!fc.f) return trr_fail(TRR_ERR_INVALID_ARG, std::string("trr_dense_save: cannot open ") + path);
  DenseSnapHeader hd{};
  memcpy(hd.magic, "TRRDNS01", 8);
  hd.dim = h->dim; hd.metric = h->metric; hd.dtype = h->dtype; hd.base = h->base; hd.n = h->n; hd.n_dead = h->n_dead;
  if (fwrite(&hd, sizeof(hd), 1, fc.f) != 1) return trr_fail(TRR_ERR_INVALID_ARG, "snapshot: short write");
  TRR_CHECK(snap_write_dev(h->ctx, fc.f, h->rows, h->n * h->row_bytes));
  TRR_CHECK(snap_write_dev(h->ctx, fc.f, h->dead, h->n));
  return TRR_OK;
}

extern "C" int trr_dense_load(trr_ctx* ctx, const char* path, trr_dense** out) {
  if (!ctx || !path || !out) return trr_fail(TRR_ERR_INVALID_ARG, "trr_dense_load: NULL argument");
  *out = nullptr;
  FileCloser fc{fopen(path, "rb")};
  if (!fc.f) return trr_fail(TRR_ERR_INVALID_ARG, std::string("trr_dense_load: cannot open ") + path);
  DenseSnapHeader hd{};
  if (fread(&hd, sizeof(hd), 1, fc.f) != 1 || memcmp(hd.magic, "TRRDNS01", 8) != 0)
    return trr_fail(TRR_ERR_INVALID_ARG, "trr_dense_load: not a dense snapshot");
  if (hd.n_dead > hd.n || hd.n + hd.base > 0xFFFFFFFEull) return trr_fail(TRR_ERR_INVALID_ARG, "trr_dense_load: corrupt header");
  trr_dense* h = nullptr;
  TRR_CHECK(trr_dense_create(ctx, hd.dim, hd.metric, hd.dtype, std::max<uint64_t>(hd.n, 1), &h));
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  h->base = hd.base;
  int s = snap_read_dev(ctx, fc.f, h->rows, hd.n * h->row_bytes);
  if (s == TRR_OK) s = snap_read_dev(ctx, fc.f, h->dead, hd.n);
  if (s != TRR_OK) { trr_dense_destroy(h); return s; }
  h->n = hd.n; h->n_dead = hd.n_dead; h->frozen_n = 0; h->gemm_ready = false;  // norms / GEMM operands rebuilt lazily
  *out = h;
  return TRR_OK;
}

extern "C" int trr_bm25_save(trr_bm25* h, const char* path) {
  if (!h || !path) return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_save: NULL argument");
  std::lock_guard<std::mutex> lk(h->ctx->mu);
  DeviceGuard g(h->ctx->device);
  TRR_CUDA(cudaStreamSynchronize(h->ctx->stream));
  FileCloser fc{fopen(path, "wb")};
  if (!fc.f) return trr_fail(TRR_ERR_INVALID_ARG, std::string("trr_bm25_save: cannot open ") + path);
  Bm25SnapHeader hd{};
  memcpy(hd.magic, "TRRBM253", 8);
  hd.n_docs = h->n_docs; hd.n_terms = h->n_terms; hd.doc_base = h->doc_base; hd.range_shift = h->range_shift;
  hd.n_ranges = h->n_ranges; hd.skip_ld = h->skip_ld; hd.n_postings = h->n_postings;
  hd.n_dead_postings = h->n_dead_postings; hd.impacts_ok = h->impacts_ok ? 1u : 0u;
  if (fwrite(&hd, sizeof(hd), 1, fc.f) != 1) return trr_fail(TRR_ERR_INVALID_ARG, "snapshot: short write");
  TRR_CHECK(snap_write_dev(h->ctx, fc.f, h->post, (h->n_postings + 2) * sizeof(uint2)));
  TRR_CHECK(snap_write_dev(h->ctx, fc.f, h->skip, (uint64_t)h->n_terms * h->skip_ld * 4));
  TRR_CHECK(snap_write_dev(h->ctx, fc.f, h->term_min, ((uint64_t)h->n_terms + 1) * 4));
  TRR_CHECK(snap_write_dev(h->ctx, fc.f, h->term_max, (uint64_t)h->n_terms * 4));
  TRR_CHECK(snap_write_dev(h->ctx, fc.f, h->d_term_off, ((uint64_t)h->n_terms + 1) * 8));
  TRR_CHECK(snap_write_dev(h->ctx, fc.f, h->tf, h->n_postings * 4));
  TRR_CHECK(snap_write_dev(h->ctx, fc.f, h->doc_len, (uint64_t)h->n_docs * 4));
  return TRR_OK;
}

extern "C" int trr_bm25_load(trr_ctx* ctx, const char* path, trr_bm25** out) {
  if (!ctx || !path || !out) return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_load: NULL argument");
  *out = nullptr;
  FileCloser fc{fopen(path, "rb")};
  if (!fc.f) return trr_fail(TRR_ERR_INVALID_ARG, std::string("trr_bm25_load: cannot open ") + path);
  Bm25SnapHeader hd{};
  if (fread(&hd, sizeof(hd), 1, fc.f) != 1 || memcmp(hd.magic, "TRRBM253", 8) != 0)
    return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_load: not a BM25 snapshot");
  if (hd.range_shift < TRR_BM25_MIN_RANGE_SHIFT || hd.range_shift > TRR_BM25_MAX_RANGE_SHIFT || hd.skip_ld != hd.n_ranges + 1 ||
      hd.n_postings >= 0xFFFFFFFFull || hd.n_dead_postings > hd.n_postings ||
      hd.n_ranges != (hd.n_docs ? (uint32_t)(((uint64_t)hd.n_docs + (1u << hd.range_shift) - 1) >> hd.range_shift) : 0u))
    return trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_load: corrupt header");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  trr_bm25* h = new trr_bm25();
  h->ctx = ctx; h->n_docs = hd.n_docs; h->n_terms = hd.n_terms; h->doc_base = hd.doc_base; h->n_postings = hd.n_postings;
  h->range_shift = hd.range_shift; h->n_ranges = hd.n_ranges; h->skip_ld = hd.skip_ld;
  h->n_dead_postings = hd.n_dead_postings; h->impacts_ok = hd.impacts_ok != 0;
  for (auto& e : h->ev) cudaEventCreate(&e);
  auto fail = [&](int s) { trr_bm25_destroy(h); return s; };
  if (cudaMalloc(&h->post, (hd.n_postings + 2) * sizeof(uint2)) != cudaSuccess ||
      cudaMalloc(&h->skip, std::max<uint64_t>((uint64_t)hd.n_terms * hd.skip_ld, 1) * 4) != cudaSuccess ||
      cudaMalloc(&h->term_min, ((uint64_t)hd.n_terms + 1) * 4) != cudaSuccess ||
      cudaMalloc(&h->term_max, std::max<uint64_t>(hd.n_terms, 1) * 4) != cudaSuccess ||
      cudaMalloc(&h->d_term_off, ((uint64_t)hd.n_terms + 1) * 8) != cudaSuccess ||
      cudaMalloc(&h->tf, std::max<uint64_t>(hd.n_postings, 1) * 4) != cudaSuccess ||
      cudaMalloc(&h->doc_len, std::max<uint64_t>(hd.n_docs, 1) * 4) != cudaSuccess) {
    cudaGetLastError();
    trr_fail(TRR_ERR_OOM, "trr_bm25_load: out of device memory");
    return fail(TRR_ERR_OOM);
  }
  int s = snap_read_dev(ctx, fc.f, h->post, (hd.n_postings + 2) * sizeof(uint2));
  if (s == TRR_OK) s = snap_read_dev(ctx, fc.f, h->skip, (uint64_t)hd.n_terms * hd.skip_ld * 4);
  if (s == TRR_OK) s = snap_read_dev(ctx, fc.f, h->term_min, ((uint64_t)hd.n_terms + 1) * 4);
  if (s == TRR_OK) s = snap_read_dev(ctx, fc.f, h->term_max, (uint64_t)hd.n_terms * 4);
  if (s == TRR_OK) s = snap_read_dev(ctx, fc.f, h->d_term_off, ((uint64_t)hd.n_terms + 1) * 8);
  if (s == TRR_OK) s = snap_read_dev(ctx, fc.f, h->tf, hd.n_postings * 4);
  if (s == TRR_OK) s = snap_read_dev(ctx, fc.f, h->doc_len, (uint64_t)hd.n_docs * 4);
  if (s == TRR_OK) {
    h->h_term_off.resize((size_t)hd.n_terms + 1);
    if (cudaMemcpy(h->h_term_off.data(), h->d_term_off, h->h_term_off.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
      s = trr_fail(TRR_ERR_CUDA, "trr_bm25_load: copy failed");
  }
  if (s == TRR_OK) {
    // the offsets index the posting arrays on the device: a file whose offsets are not a monotone partition of the postings
    // must be refused here, not discovered by a kernel
    bool ok = h->h_term_off.front() == 0 && h->h_term_off.back() == hd.n_postings;
    for (size_t t = 0; ok && t + 1 < h->h_term_off.size(); ++t) ok = h->h_term_off[t] <= h->h_term_off[t + 1];
    if (!ok) s = trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_load: corrupt term offsets");
  }
  if (s == TRR_OK) {
    // ... and the skip rows and document ids, which the search kernels use as indices (bm25_check_kernel)
    uint32_t* d_bad = nullptr;
    uint32_t bad = 1;
    if (cudaMalloc(&d_bad, 4) != cudaSuccess || cudaMemsetAsync(d_bad, 0, 4, ctx->stream) != cudaSuccess ||
        trr_launch_bm25_check(h->skip, h->skip_ld, h->n_terms, h->n_ranges, h->d_term_off, h->post, h->n_postings, h->n_docs,
                              d_bad, ctx->stream) != cudaSuccess ||
        cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      cudaGetLastError();
      s = trr_fail(TRR_ERR_CUDA, "trr_bm25_load: validation failed to run");
    } else if (bad) {
      s = trr_fail(TRR_ERR_INVALID_ARG, "trr_bm25_load: corrupt skip table or document ids");
    }
    if (d_bad) cudaFree(d_bad);
  }
  if (s != TRR_OK) return fail(s);
  *out = h;
  return TRR_OK;
}

// ------------------------------------------------------------------------------------------------
// fusion / hybrid
// ------------------------------------------------------------------------------------------------
extern "C" size_t trr_exchange_bytes(uint32_t B, uint32_t C) { return ((size_t)4 * B * C + (size_t)2 * B) * 4; }

struct ExchangeView {
  uint32_t* ord[2];
  float* score[2];
  uint32_t* n[2];
};
static ExchangeView exchange_view(void* rec, uint32_t B, uint32_t C) {
  ExchangeView v;
  uint32_t* w = static_cast<uint32_t*>(rec);
  const size_t bc = (size_t)B * C;
  v.ord[0] = w; v.ord[1] = w + bc;
  v.score[0] = reinterpret_cast<float*>(w + 2 * bc); v.score[1] = reinterpret_cast<float*>(w + 3 * bc);
  v.n[0] = w + 4 * bc; v.n[1] = w + 4 * bc + B;
  return v;
}

static int fuse_locked(trr_ctx* c, const ExchangeView& v, uint64_t shard_stride, uint32_t G, uint32_t B, uint32_t C,
                       int strategy, float param, uint32_t k, bool have_dense, bool have_sparse, uint32_t* d_out_ord,
                       float* d_out_fused, float* d_out_dense, float* d_out_sparse, uint32_t* d_out_n,
                       cudaStream_t st_override = nullptr) {
  TrrRange nvtx_range("trr:fuse");
  if (strategy < 0 || strategy > 5) return trr_fail(TRR_ERR_INVALID_ARG, "bad fusion strategy");
  if (C == 0 || C > 1024) return trr_fail(TRR_ERR_UNSUPPORTED, "candidates per source must be in 1..1024");
  if ((uint64_t)G * C > 8192) return trr_fail(TRR_ERR_UNSUPPORTED, "shards x candidates > 8192");
  FuseArgs a{};
  for (int s = 0; s < 2; ++s) { a.ord[s] = v.ord[s]; a.score[s] = v.score[s]; a.n[s] = v.n[s]; }
  if (!have_dense) a.n[0] = nullptr;
  if (!have_sparse) a.n[1] = nullptr;
  a.shard_stride = shard_stride; a.G = G; a.B = B; a.C = C; a.strategy = strategy; a.param = param; a.k = k;
  a.mcap = trr_pow2_ceil(std::max<uint32_t>(G * C, 2)); a.fcap = trr_pow2_ceil(2 * C);
  a.out_ord = d_out_ord; a.out_fused = d_out_fused; a.out_dense = d_out_dense; a.out_sparse = d_out_sparse;
  a.out_n = d_out_n;
  if (trr_fuse_smem(a) > c->smem_optin) return trr_fail(TRR_ERR_UNSUPPORTED, "fusion lists exceed shared memory");
  TRR_CUDA(trr_launch_fuse(a, st_override ? st_override : c->stream));
  c->launches++;
  return TRR_OK;
}

// device outputs -> host outputs helper
static int fuse_download(trr_ctx* c, uint32_t B, uint32_t k, const uint32_t* d_ord, const float* d_f, const float* d_d,
                         const float* d_s, const uint32_t* d_n, uint32_t* out_ord, float* out_fused, float* out_dense,
                         float* out_sparse, uint32_t* out_n) {
  cudaStream_t st = c->stream;
  const size_t bk = (size_t)B * k * 4;
  TRR_CUDA(cudaMemcpyAsync(out_ord, d_ord, bk, cudaMemcpyDeviceToHost, st));
  TRR_CUDA(cudaMemcpyAsync(out_fused, d_f, bk, cudaMemcpyDeviceToHost, st));
  if (out_dense) TRR_CUDA(cudaMemcpyAsync(out_dense, d_d, bk, cudaMemcpyDeviceToHost, st));
  if (out_sparse) TRR_CUDA(cudaMemcpyAsync(out_sparse, d_s, bk, cudaMemcpyDeviceToHost, st));
  TRR_CUDA(cudaMemcpyAsync(out_n, d_n, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  TRR_CUDA(cudaStreamSynchronize(st));
  return TRR_OK;
}

extern "C" int trr_fuse(trr_ctx* c, int strategy, float param, const uint32_t* d_ord, const float* d_score,
                        const uint32_t* d_n, const uint32_t* s_ord, const float* s_score, const uint32_t* s_n, uint32_t B,
                        uint32_t C, uint32_t k_out, uint32_t* out_ord, float* out_fused, float* out_dense,
                        float* out_sparse, uint32_t* out_n) {
  if (!c || (B && (!out_ord || !out_fused || !out_n))) return trr_fail(TRR_ERR_INVALID_ARG, "trr_fuse: NULL argument");
  if (B == 0) return TRR_OK;
  if (k_out == 0) return trr_fail(TRR_ERR_INVALID_ARG, "trr_fuse: k_out must be >= 1");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  cudaStream_t st = c->stream;
  const size_t rec = trr_exchange_bytes(B, C);
  const size_t bk = (size_t)B * k_out;
  TRR_CHECK(extra(c)->io.reserve(WsCarver::need({rec, bk * 4, bk * 4, bk * 4, bk * 4, (size_t)B * 4})));
  WsCarver io(extra(c)->io.p);
  uint8_t* d_rec = io.take<uint8_t>(rec);
  uint32_t* o_ord = io.take<uint32_t>(bk);
  float* o_f = io.take<float>(bk);
  float* o_d = io.take<float>(bk);
  float* o_s = io.take<float>(bk);
  uint32_t* o_n = io.take<uint32_t>(B);
  ExchangeView v = exchange_view(d_rec, B, C);
  const size_t bc = (size_t)B * C * 4;
  const bool have_d = d_ord && d_score && d_n, have_s = s_ord && s_score && s_n;
  if (have_d) {
    TRR_CUDA(cudaMemcpyAsync(v.ord[0], d_ord, bc, cudaMemcpyHostToDevice, st));
    TRR_CUDA(cudaMemcpyAsync(v.score[0], d_score, bc, cudaMemcpyHostToDevice, st));
    TRR_CUDA(cudaMemcpyAsync(v.n[0], d_n, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  }
  if (have_s) {
    TRR_CUDA(cudaMemcpyAsync(v.ord[1], s_ord, bc, cudaMemcpyHostToDevice, st));
    TRR_CUDA(cudaMemcpyAsync(v.score[1], s_score, bc, cudaMemcpyHostToDevice, st));
    TRR_CUDA(cudaMemcpyAsync(v.n[1], s_n, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  }
  TRR_CHECK(fuse_locked(c, v, 0, 1, B, C, strategy, param, k_out, have_d, have_s, o_ord, o_f, o_d, o_s, o_n));
  return fuse_download(c, B, k_out, o_ord, o_f, o_d, o_s, o_n, out_ord, out_fused, out_dense, out_sparse, out_n);
}

// shard-local stage into a device exchange record (locks taken by the caller)
static int hybrid_local_locked(trr_dense* dense, trr_bm25* bm25, trr_ctx* c, const float* q, const uint32_t* q_terms,
                               const uint32_t* q_off, uint32_t B, uint32_t C, bool use_dense, bool use_sparse,
                               void* d_exchange) {
  cudaStream_t st = c->stream;
  ExchangeView v = exchange_view(d_exchange, B, C);
  const uint32_t dim = dense ? dense->dim : 0;
  const size_t nt = (use_sparse && q_off) ? q_off[B] : 0;
  TRR_CHECK(extra(c)->io.reserve(WsCarver::need({(size_t)B * dim * 4, std::max<size_t>(nt, 1) * 4, (size_t)(B + 1) * 4})));
  WsCarver io(extra(c)->io.p);
  float* d_q = io.take<float>((size_t)B * dim);
  uint32_t* d_terms = io.take<uint32_t>(std::max<size_t>(nt, 1));
  uint32_t* d_off = io.take<uint32_t>(B + 1);
  if (use_dense) {
    TRR_CUDA(cudaMemcpyAsync(d_q, q, (size_t)B * dim * 4, cudaMemcpyHostToDevice, st));
    TRR_CHECK(dense_search_locked(dense, d_q, B, C, v.ord[0], v.score[0], v.n[0], false));
  } else {
    TRR_CUDA(cudaMemsetAsync(v.n[0], 0, (size_t)B * 4, st));
  }
  if (use_sparse) {
    if (nt) TRR_CUDA(cudaMemcpyAsync(d_terms, q_terms, nt * 4, cudaMemcpyHostToDevice, st));
    TRR_CUDA(cudaMemcpyAsync(d_off, q_off, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
    TRR_CHECK(bm25_search_locked(bm25, d_terms, d_off, q_off, B, C, v.ord[1], v.score[1], v.n[1], 0));
  } else {
    TRR_CUDA(cudaMemsetAsync(v.n[1], 0, (size_t)B * 4, st));
  }
  return TRR_OK;
}

static int hybrid_check(trr_dense* dense, trr_bm25* bm25, const float* q, const uint32_t* q_off, uint32_t B, uint32_t C,
                        int use_dense, int use_sparse, trr_ctx** c) {
  if (use_dense && (!dense || (B && !q))) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: dense store / queries missing");
  if (use_sparse && (!bm25 || (B && !q_off))) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: sparse index / query terms missing");
  if (!dense && !bm25) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: no index given");
  *c = dense ? dense->ctx : bm25->ctx;
  if (dense && bm25 && dense->ctx != bm25->ctx) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: indexes live on different contexts");
  if (C == 0 || C > 1024) return trr_fail(TRR_ERR_UNSUPPORTED, "candidates per source must be in 1..1024");
  return TRR_OK;
}

extern "C" int trr_hybrid_local_device(trr_dense* dense, trr_bm25* bm25, const float* d_q, const uint32_t* d_q_terms,
                                       const uint32_t* d_q_off, const uint32_t* h_q_off, uint32_t B, uint32_t C,
                                       int use_dense, int use_sparse, void* d_exchange) {
  trr_ctx* c = nullptr;
  TRR_CHECK(hybrid_check(dense, bm25, d_q, h_q_off, B, C, use_dense, use_sparse, &c));
  if (!d_exchange || (use_sparse && B && !d_q_off)) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: NULL device buffer");
  if (B == 0) return TRR_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  cudaStream_t st = c->stream;
  ExchangeView v = exchange_view(d_exchange, B, C);
  if (use_dense) TRR_CHECK(dense_search_locked(dense, d_q, B, C, v.ord[0], v.score[0], v.n[0], false));
  else TRR_CUDA(cudaMemsetAsync(v.n[0], 0, (size_t)B * 4, st));
  if (use_sparse) TRR_CHECK(bm25_search_locked(bm25, d_q_terms, d_q_off, h_q_off, B, C, v.ord[1], v.score[1], v.n[1], 0));
  else TRR_CUDA(cudaMemsetAsync(v.n[1], 0, (size_t)B * 4, st));
  return TRR_OK;
}

extern "C" int trr_hybrid_merge_device(trr_ctx* c, const void* d_gathered, uint32_t G, uint32_t B, uint32_t C, int strategy,
                                       float param, uint32_t k, uint32_t* d_out_ord, float* d_out_fused,
                                       float* d_out_dense, float* d_out_sparse, uint32_t* d_out_n) {
  if (!c || !d_gathered || (B && (!d_out_ord || !d_out_fused || !d_out_n))) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  if (B == 0) return TRR_OK;
  if (k == 0 || G == 0) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid merge: k and G must be >= 1");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  ExchangeView v = exchange_view(const_cast<void*>(d_gathered), B, C);
  return fuse_locked(c, v, trr_exchange_bytes(B, C) / 4, G, B, C, strategy, param, k, true, true, d_out_ord, d_out_fused,
                     d_out_dense, d_out_sparse, d_out_n);
}

extern "C" int trr_hybrid_local(trr_dense* dense, trr_bm25* bm25, const float* q, const uint32_t* q_terms,
                                const uint32_t* q_off, uint32_t B, uint32_t C, int use_dense, int use_sparse,
                                void* d_exchange) {
  trr_ctx* c = nullptr;
  TRR_CHECK(hybrid_check(dense, bm25, q, q_off, B, C, use_dense, use_sparse, &c));
  if (!d_exchange) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: exchange buffer is NULL");
  if (B == 0) return TRR_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  TRR_CHECK(hybrid_local_locked(dense, bm25, c, q, q_terms, q_off, B, C, use_dense != 0, use_sparse != 0, d_exchange));
  TRR_CUDA(cudaStreamSynchronize(c->stream));
  return TRR_OK;
}

static int hybrid_merge_locked(trr_ctx* c, const void* d_gathered, uint32_t G, uint32_t B, uint32_t C, int strategy,
                               float param, uint32_t k, uint32_t* out_ord, float* out_fused, float* out_dense,
                               float* out_sparse, uint32_t* out_n) {
  const size_t bk = (size_t)B * k;
  TRR_CHECK(extra(c)->scratch.reserve(WsCarver::need({bk * 4, bk * 4, bk * 4, bk * 4, (size_t)B * 4})));
  WsCarver ws(extra(c)->scratch.p);
  uint32_t* o_ord = ws.take<uint32_t>(bk);
  float* o_f = ws.take<float>(bk);
  float* o_d = ws.take<float>(bk);
  float* o_s = ws.take<float>(bk);
  uint32_t* o_n = ws.take<uint32_t>(B);
  ExchangeView v = exchange_view(const_cast<void*>(d_gathered), B, C);
  TRR_CHECK(fuse_locked(c, v, trr_exchange_bytes(B, C) / 4, G, B, C, strategy, param, k, true, true, o_ord, o_f, o_d, o_s,
                        o_n));
  return fuse_download(c, B, k, o_ord, o_f, o_d, o_s, o_n, out_ord, out_fused, out_dense, out_sparse, out_n);
}

extern "C" int trr_hybrid_merge(trr_ctx* c, const void* d_gathered, uint32_t G, uint32_t B, uint32_t C, int strategy,
                                float param, uint32_t k, uint32_t* out_ord, float* out_fused, float* out_dense,
                                float* out_sparse, uint32_t* out_n) {
  if (!c || !d_gathered || (B && (!out_ord || !out_fused || !out_n))) return trr_fail(TRR_ERR_INVALID_ARG, "NULL argument");
  if (B == 0) return TRR_OK;
  if (k == 0 || G == 0) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid merge: k and G must be >= 1");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  return hybrid_merge_locked(c, d_gathered, G, B, C, strategy, param, k, out_ord, out_fused, out_dense, out_sparse, out_n);
}

extern "C" int trr_hybrid_search(trr_dense* dense, trr_bm25* bm25, const float* q, const uint32_t* q_terms,
                                 const uint32_t* q_off, uint32_t B, uint32_t C, int strategy, float param, uint32_t k,
                                 int use_dense, int use_sparse, uint32_t* out_ord, float* out_fused, float* out_dense,
                                 float* out_sparse, uint32_t* out_n) {
  trr_ctx* c = nullptr;
  TRR_CHECK(hybrid_check(dense, bm25, q, q_off, B, C, use_dense, use_sparse, &c));
  if (B && (!out_ord || !out_fused || !out_n)) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: NULL output");
  if (B == 0) return TRR_OK;
  if (k == 0) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: k must be >= 1");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  TRR_CHECK(extra(c)->hy.reserve(trr_exchange_bytes(B, C)));
  TRR_CHECK(hybrid_local_locked(dense, bm25, c, q, q_terms, q_off, B, C, use_dense != 0, use_sparse != 0, extra(c)->hy.p));
  return hybrid_merge_locked(c, extra(c)->hy.p, 1, B, C, strategy, param, k, out_ord, out_fused, out_dense, out_sparse,
                             out_n);
}

// ------------------------------------------------------------------------------------------------
// sharded search: one process per GPU, the cross-GPU exchange inside the call
// ------------------------------------------------------------------------------------------------
// NCCL is resolved at run time (dlopen of libnccl.so.2, the library torch.distributed already maps into a Python host and
// a Rust host links): the library itself has no link-time dependency on it, and a single-GPU host never needs it.
#include <dlfcn.h>

namespace {
struct NcclId { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return TRR_OK;
  void* lib = nullptr;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return trr_fail(TRR_ERR_UNSUPPORTED, std::string("sharded search needs NCCL (libnccl.so.2 on the library path): ") + dlerror());
  NcclApi n;
  n.lib = lib;
  n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(lib, "ncclAllGather"));
  n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllGather || !n.AllReduce || !n.GetErrorString)
    return trr_fail(TRR_ERR_UNSUPPORTED, "libnccl lacks a required symbol");
  g_nccl = n;
  return TRR_OK;
}
constexpr int NCCL_U8 = 1, NCCL_U64 = 5, NCCL_SUM = 0, NCCL_MAX = 2;
}  // namespace

#define TRR_NCCL(expr)                                                                                     \
  do {                                                                                                     \
    int _r = (expr);                                                                                       \
    if (_r != 0) return trr_fail(TRR_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));    \
  } while (0)

struct trr_group {
  trr_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  void* comm = nullptr;
  int exchange = TRR_EXCHANGE_NCCL;   // what the calls use
  cudaStream_t xs = nullptr;          // exchange + merge stream (overlaps the shard-local kernels of the next call)
  cudaStream_t cs = nullptr;          // host-buffer calls: input copies (overlap the kernels of the previous call)
  cudaEvent_t ev_local[2] = {nullptr, nullptr}, ev_rec_free[2] = {nullptr, nullptr};
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  size_t io_set = 0;                  // bytes of one staging set of the host-buffer calls
  uint64_t calls = 0;
  DevBuf rec[2];                      // this rank's exchange record, by parity of the call
  DevBuf gath[2];                     // NCCL exchange: gathered records
  DevBuf stage;                       // all-reduce / handle staging
  DevBuf io;                          // device copies of the host-buffer entry point
  // peer exchange
  size_t max_rec = 0;                 // record slot size of the shared gather buffers
  uint8_t* shared = nullptr;          // own block: gather[2][world][max_rec] | flags[2][world]
  size_t flags_off = 0;
  uint8_t* peer_base[TRR_MAX_GROUP] = {};
  uint32_t* block_counter = nullptr;
};

extern "C" int trr_group_unique_id(void* out_id128) {
  if (!out_id128) return trr_fail(TRR_ERR_INVALID_ARG, "trr_group_unique_id: NULL argument");
  TRR_CHECK(nccl_load());
  NcclId id;
  TRR_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out_id128, &id, sizeof(id));
  return TRR_OK;
}

extern "C" int trr_group_destroy(trr_group* g) {
  if (!g) return TRR_OK;
  DeviceGuard dg(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  if (g->xs) cudaStreamSynchronize(g->xs);
  if (g->cs) cudaStreamSynchronize(g->cs);
  for (int p = 0; p < g->world && p < (int)TRR_MAX_GROUP; ++p)
    if (p != g->rank && g->peer_base[p]) cudaIpcCloseMemHandle(g->peer_base[p]);
  if (g->comm) g_nccl.CommDestroy(g->comm);
  if (g->shared) cudaFree(g->shared);
  if (g->block_counter) cudaFree(g->block_counter);
  for (auto& b : g->rec) b.release();
  for (auto& b : g->gath) b.release();
  g->stage.release(); g->io.release();
  for (auto& e : g->ev_local) if (e) cudaEventDestroy(e);
  for (auto& e : g->ev_rec_free) if (e) cudaEventDestroy(e);
  for (auto& e : g->ev_in) if (e) cudaEventDestroy(e);
  for (auto& e : g->ev_out) if (e) cudaEventDestroy(e);
  if (g->xs) cudaStreamDestroy(g->xs);
  if (g->cs) cudaStreamDestroy(g->cs);
  cudaGetLastError();
  delete g;
  return TRR_OK;
}

extern "C" int trr_group_create(trr_ctx* ctx, const void* id128, int rank, int world, int exchange, size_t max_record_bytes,
                                trr_group** out) {
  if (!ctx || !out || (world > 1 && !id128)) return trr_fail(TRR_ERR_INVALID_ARG, "trr_group_create: NULL argument");
  *out = nullptr;
  if (world < 1 || world > (int)TRR_MAX_GROUP || rank < 0 || rank >= world)
    return trr_fail(TRR_ERR_INVALID_ARG, "trr_group_create: rank / world out of range (at most 16 ranks)");
  if (exchange != TRR_EXCHANGE_NCCL && exchange != TRR_EXCHANGE_PEER) return trr_fail(TRR_ERR_INVALID_ARG, "bad exchange kind");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard dg(ctx->device);
  trr_group* g = new trr_group();
  g->ctx = ctx; g->rank = rank; g->world = world;
  auto fail = [&](int s) { trr_group_destroy(g); return s; };
  if (cudaStreamCreateWithFlags(&g->xs, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&g->cs, cudaStreamNonBlocking) != cudaSuccess)
    return fail(trr_fail(TRR_ERR_CUDA, "stream creation failed"));
  for (int p = 0; p < 2; ++p) {
    cudaEventCreateWithFlags(&g->ev_local[p], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&g->ev_rec_free[p], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&g->ev_in[p], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&g->ev_out[p], cudaEventDisableTiming);
  }
  if (world > 1) {
    int s = nccl_load();
    if (s != TRR_OK) return fail(s);
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    int r = g_nccl.CommInitRank(&g->comm, world, id, rank);
    if (r != 0) return fail(trr_fail(TRR_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)));
  }
  g->exchange = TRR_EXCHANGE_NCCL;
  if (exchange == TRR_EXCHANGE_PEER && world > 1 && max_record_bytes) {
    // one IPC-shared block per rank; the handles travel through an NCCL all-gather; any failure falls back to NCCL exchange
    g->max_rec = (max_record_bytes + 255) & ~size_t(255);
    g->flags_off = (size_t)2 * world * g->max_rec;
    const size_t total = g->flags_off + 4096;
    bool ok = cudaMalloc(&g->shared, total) == cudaSuccess && cudaMemset(g->shared, 0, total) == cudaSuccess &&
              cudaMalloc(&g->block_counter, 256) == cudaSuccess && cudaMemset(g->block_counter, 0, 256) == cudaSuccess;
    cudaIpcMemHandle_t mine;
    ok = ok && cudaIpcGetMemHandle(&mine, g->shared) == cudaSuccess;
    // every rank takes part in the all-gather (a rank whose allocation failed sends zeros and the group agrees on NCCL)
    uint64_t okflag = ok ? 1 : 0;
    if (g->stage.reserve((size_t)(world + 1) * 128) != TRR_OK) return fail(TRR_ERR_OOM);
    uint8_t sendbuf[128] = {0};
    if (ok) memcpy(sendbuf, &mine, sizeof(mine));
    memcpy(sendbuf + 64, &okflag, 8);
    uint8_t* d_send = static_cast<uint8_t*>(g->stage.p);
    uint8_t* d_recv = d_send + 128;
    cudaMemcpy(d_send, sendbuf, 128, cudaMemcpyHostToDevice);
    int r = g_nccl.AllGather(d_send, d_recv, 128, NCCL_U8, g->comm, g->xs);
    if (r != 0 || cudaStreamSynchronize(g->xs) != cudaSuccess) return fail(trr_fail(TRR_ERR_CUDA, "handle all-gather failed"));
    std::vector<uint8_t> all((size_t)world * 128);
    cudaMemcpy(all.data(), d_recv, all.size(), cudaMemcpyDeviceToHost);
    bool all_ok = true;
    for (int p = 0; p < world; ++p) { uint64_t f; memcpy(&f, &all[(size_t)p * 128 + 64], 8); all_ok = all_ok && f == 1; }
    if (all_ok) {
      for (int p = 0; p < world && all_ok; ++p) {
        if (p == rank) { g->peer_base[p] = g->shared; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, &all[(size_t)p * 128], sizeof(h));
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); all_ok = false; }
        g->peer_base[p] = static_cast<uint8_t*>(ptr);
      }
    }
    // second agreement round: a rank that could not open a peer's block must take everyone back to NCCL
    uint64_t v = all_ok ? 0 : 1;
    cudaMemcpy(d_send, &v, 8, cudaMemcpyHostToDevice);
    r = g_nccl.AllReduce(d_send, d_send, 1, NCCL_U64, NCCL_SUM, g->comm, g->xs);
    if (r != 0 || cudaStreamSynchronize(g->xs) != cudaSuccess) return fail(trr_fail(TRR_ERR_CUDA, "agreement all-reduce failed"));
    cudaMemcpy(&v, d_send, 8, cudaMemcpyDeviceToHost);
    if (v == 0) g->exchange = TRR_EXCHANGE_PEER;
    cudaGetLastError();
  }
  *out = g;
  return TRR_OK;
}

extern "C" int trr_group_info(trr_group* g, int* out_rank, int* out_world, int* out_exchange) {
  if (!g) return trr_fail(TRR_ERR_INVALID_ARG, "group is NULL");
  if (out_rank) *out_rank = g->rank;
  if (out_world) *out_world = g->world;
  if (out_exchange) *out_exchange = g->exchange;
  return TRR_OK;
}

extern "C" int trr_group_sync(trr_group* g) {
  if (!g) return trr_fail(TRR_ERR_INVALID_ARG, "group is NULL");
  DeviceGuard dg(g->ctx->device);
  TRR_CUDA(cudaStreamSynchronize(g->ctx->stream));
  TRR_CUDA(cudaStreamSynchronize(g->xs));
  TRR_CUDA(cudaStreamSynchronize(g->cs));
  return TRR_OK;
}

// element-wise sum / max over the ranks of n u64 values held in HOST memory (global BM25 statistics: df, total length;
// max-over-ranks timings).  Blocking.
extern "C" int trr_group_allreduce_u64(trr_group* g, uint64_t* inout, size_t n, int op_max) {
  if (!g || (n && !inout)) return trr_fail(TRR_ERR_INVALID_ARG, "trr_group_allreduce_u64: NULL argument");
  if (g->world == 1 || n == 0) return TRR_OK;
  std::lock_guard<std::mutex> lk(g->ctx->mu);
  DeviceGuard dg(g->ctx->device);
  TRR_CHECK(g->stage.reserve(n * 8));
  TRR_CUDA(cudaMemcpyAsync(g->stage.p, inout, n * 8, cudaMemcpyHostToDevice, g->xs));
  TRR_NCCL(g_nccl.AllReduce(g->stage.p, g->stage.p, n, NCCL_U64, op_max ? NCCL_MAX : NCCL_SUM, g->comm, g->xs));
  TRR_CUDA(cudaMemcpyAsync(inout, g->stage.p, n * 8, cudaMemcpyDeviceToHost, g->xs));
  TRR_CUDA(cudaStreamSynchronize(g->xs));
  return TRR_OK;
}

// One sharded hybrid step with DEVICE buffers.  The shard-local kernels are enqueued on the context stream; the exchange
// and the merge + fusion kernel on the group's second stream, behind an event, so that they overlap the shard-local
// kernels of the NEXT call (records and gather buffers alternate by the parity of the call).  Nothing waits on the host.
static int group_step_locked(trr_group* g, trr_dense* dense, trr_bm25* bm25, const float* d_q, const uint32_t* d_q_terms,
                             const uint32_t* d_q_off, const uint32_t* h_q_off, uint32_t B, uint32_t C, int strategy,
                             float param, uint32_t k, bool use_dense, bool use_sparse, uint32_t* d_out_ord, float* d_out_fused,
                             float* d_out_dense, float* d_out_sparse, uint32_t* d_out_n) {
  TrrRange nvtx_range("trr:sharded_step");
  trr_ctx* c = g->ctx;
  cudaStream_t st = c->stream;
  const uint32_t G = (uint32_t)g->world;
  const size_t rec_bytes = trr_exchange_bytes(B, C);
  const int p = (int)(g->calls & 1);
  const uint64_t seq = ++g->calls;
  TRR_CHECK(g->rec[p].reserve(rec_bytes));
  const bool peer = g->exchange == TRR_EXCHANGE_PEER && rec_bytes <= g->max_rec && (rec_bytes % 16) == 0;
  if (!peer && G > 1) TRR_CHECK(g->gath[p].reserve((size_t)G * rec_bytes));
  // the exchange of call seq-2 has consumed rec[p]
  TRR_CUDA(cudaStreamWaitEvent(st, g->ev_rec_free[p], 0));
  ExchangeView v = exchange_view(g->rec[p].p, B, C);
  if (use_dense) TRR_CHECK(dense_search_locked(dense, d_q, B, C, v.ord[0], v.score[0], v.n[0], false));
  else TRR_CUDA(cudaMemsetAsync(v.n[0], 0, (size_t)B * 4, st));
  if (use_sparse) TRR_CHECK(bm25_search_locked(bm25, d_q_terms, d_q_off, h_q_off, B, C, v.ord[1], v.score[1], v.n[1], 0));
  else TRR_CUDA(cudaMemsetAsync(v.n[1], 0, (size_t)B * 4, st));
  TRR_CUDA(cudaEventRecord(g->ev_local[p], st));
  TRR_CUDA(cudaStreamWaitEvent(g->xs, g->ev_local[p], 0));
  const void* gathered = g->rec[p].p;
  uint64_t stride_words = rec_bytes / 4;
  if (G > 1) {
    if (peer) {
      ExchangeScatterArgs a{};
      a.record = static_cast<const uint8_t*>(g->rec[p].p); a.record_bytes = rec_bytes;
      for (uint32_t r = 0; r < G; ++r) {
        a.peer_gath[r] = g->peer_base[r];
        a.peer_flags[r] = reinterpret_cast<uint32_t*>(g->peer_base[r] + g->flags_off);
      }
      a.slot_off = ((size_t)p * G + (size_t)g->rank) * g->max_rec;
      a.flag_index = (uint32_t)p * G + (uint32_t)g->rank;
      a.seq = (uint32_t)seq; a.block_counter = g->block_counter;
      TRR_CUDA(trr_launch_exchange_scatter(a, G, g->xs));
      TRR_CUDA(cudaEventRecord(g->ev_rec_free[p], g->xs));
      TRR_CUDA(trr_launch_exchange_wait(reinterpret_cast<const uint32_t*>(g->shared + g->flags_off) + (size_t)p * G, G, (uint32_t)seq,
                                        extra(c)->dbg_dev, g->xs));
      c->launches += 2;
      gathered = g->shared + (size_t)p * G * g->max_rec;
      stride_words = g->max_rec / 4;
    } else {
      TRR_NCCL(g_nccl.AllGather(g->rec[p].p, g->gath[p].p, rec_bytes, NCCL_U8, g->comm, g->xs));
      TRR_CUDA(cudaEventRecord(g->ev_rec_free[p], g->xs));
      gathered = g->gath[p].p;
    }
  }
  ExchangeView gv = exchange_view(const_cast<void*>(gathered), B, C);
  TRR_CHECK(fuse_locked(c, gv, stride_words, G, B, C, strategy, param, k, true, true, d_out_ord, d_out_fused, d_out_dense,
                        d_out_sparse, d_out_n, g->xs));
  if (G == 1) TRR_CUDA(cudaEventRecord(g->ev_rec_free[p], g->xs));  // (the merge reads the record itself)
  return TRR_OK;
}

extern "C" int trr_hybrid_search_sharded_device(trr_group* g, trr_dense* dense, trr_bm25* bm25, const float* d_q,
                                                const uint32_t* d_q_terms, const uint32_t* d_q_off, const uint32_t* h_q_off,
                                                uint32_t B, uint32_t C, int strategy, float param, uint32_t k, int use_dense,
                                                int use_sparse, uint32_t* d_out_ord, float* d_out_fused, float* d_out_dense,
                                                float* d_out_sparse, uint32_t* d_out_n) {
  if (!g) return trr_fail(TRR_ERR_INVALID_ARG, "group is NULL");
  trr_ctx* c = nullptr;
  TRR_CHECK(hybrid_check(dense, bm25, d_q, h_q_off, B, C, use_dense, use_sparse, &c));
  if (c != g->ctx) return trr_fail(TRR_ERR_INVALID_ARG, "sharded search: the indexes live on another context than the group");
  if (B && (!d_out_ord || !d_out_fused || !d_out_n || (use_sparse && !d_q_off))) return trr_fail(TRR_ERR_INVALID_ARG, "sharded search: NULL device buffer");
  if (B == 0) return TRR_OK;
  if (k == 0) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: k must be >= 1");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard dg(c->device);
  return group_step_locked(g, dense, bm25, d_q, d_q_terms, d_q_off, h_q_off, B, C, strategy, param, k, use_dense != 0,
                           use_sparse != 0, d_out_ord, d_out_fused, d_out_dense, d_out_sparse, d_out_n);
}

// HybridRetriever::retrieve over a corpus sharded by document across the ranks of the group: ONE call = shard-local
// dense + BM25 top-C -> exchange -> merge, fusion, top-k; every rank receives the full result.  HOST buffers.
// The inputs of a call travel on a copy stream into one of two device staging sets (by the parity of the call) and the
// results leave behind the merge kernel on the exchange stream, so consecutive asynchronous calls overlap their copies
// with each other's kernels; the blocking form is the asynchronous one followed by a wait.
static int group_search_host(trr_group* g, trr_dense* dense, trr_bm25* bm25, const float* q, const uint32_t* q_terms,
                             const uint32_t* q_off, uint32_t B, uint32_t C, int strategy, float param, uint32_t k,
                             int use_dense, int use_sparse, uint32_t* out_ord, float* out_fused, float* out_dense,
                             float* out_sparse, uint32_t* out_n, bool blocking) {
  TrrRange nvtx_range("trr:sharded_search_host");
  if (!g) return trr_fail(TRR_ERR_INVALID_ARG, "group is NULL");
  trr_ctx* c = nullptr;
  TRR_CHECK(hybrid_check(dense, bm25, q, q_off, B, C, use_dense, use_sparse, &c));
  if (c != g->ctx) return trr_fail(TRR_ERR_INVALID_ARG, "sharded search: the indexes live on another context than the group");
  if (B && (!out_ord || !out_fused || !out_n)) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: NULL output");
  if (B == 0) return TRR_OK;
  if (k == 0) return trr_fail(TRR_ERR_INVALID_ARG, "hybrid: k must be >= 1");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard dg(c->device);
  cudaStream_t st = c->stream;
  const uint32_t dim = dense ? dense->dim : 0;
  const size_t nt = (use_sparse && q_off) ? q_off[B] : 0;
  const size_t bk = (size_t)B * k;
  const int p = (int)(g->calls & 1);  // parity of the call that group_step_locked is about to run
  const size_t set_bytes = WsCarver::need({(size_t)B * dim * 4, std::max<size_t>(nt, 1) * 4, (size_t)(B + 1) * 4, bk * 4, bk * 4,
                                           bk * 4, bk * 4, (size_t)B * 4});
  if (g->io.bytes < 2 * set_bytes) {
    // (re)allocation: nothing may still be using the old buffers
    TRR_CUDA(cudaStreamSynchronize(st)); TRR_CUDA(cudaStreamSynchronize(g->xs)); TRR_CUDA(cudaStreamSynchronize(g->cs));
    TRR_CHECK(g->io.reserve(2 * set_bytes));
    g->io_set = set_bytes;
  }
  WsCarver io(static_cast<char*>(g->io.p) + (size_t)p * g->io_set);
  float* d_q = io.take<float>((size_t)B * dim);
  uint32_t* d_terms = io.take<uint32_t>(std::max<size_t>(nt, 1));
  uint32_t* d_off = io.take<uint32_t>(B + 1);
  uint32_t* o_ord = io.take<uint32_t>(bk);
  float* o_f = io.take<float>(bk);
  float* o_d = io.take<float>(bk);
  float* o_s = io.take<float>(bk);
  uint32_t* o_n = io.take<uint32_t>(B);
  // the call two steps back used this staging set: its kernels have read the inputs, its results have left
  cudaStream_t cs = g->cs;
  TRR_CUDA(cudaStreamWaitEvent(cs, g->ev_local[p], 0));
  TRR_CUDA(cudaStreamWaitEvent(cs, g->ev_out[p], 0));
  if (use_dense) TRR_CUDA(cudaMemcpyAsync(d_q, q, (size_t)B * dim * 4, cudaMemcpyHostToDevice, cs));
  if (use_sparse) {
    if (nt) TRR_CUDA(cudaMemcpyAsync(d_terms, q_terms, nt * 4, cudaMemcpyHostToDevice, cs));
    TRR_CUDA(cudaMemcpyAsync(d_off, q_off, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, cs));
  }
  TRR_CUDA(cudaEventRecord(g->ev_in[p], cs));
  TRR_CUDA(cudaStreamWaitEvent(st, g->ev_in[p], 0));
  TRR_CHECK(group_step_locked(g, dense, bm25, d_q, d_terms, d_off, q_off, B, C, strategy, param, k, use_dense != 0,
                              use_sparse != 0, o_ord, o_f, o_d, o_s, o_n));
  cudaStream_t xs = g->xs;
  TRR_CUDA(cudaMemcpyAsync(out_ord, o_ord, bk * 4, cudaMemcpyDeviceToHost, xs));
  TRR_CUDA(cudaMemcpyAsync(out_fused, o_f, bk * 4, cudaMemcpyDeviceToHost, xs));
  if (out_dense) TRR_CUDA(cudaMemcpyAsync(out_dense, o_d, bk * 4, cudaMemcpyDeviceToHost, xs));
  if (out_sparse) TRR_CUDA(cudaMemcpyAsync(out_sparse, o_s, bk * 4, cudaMemcpyDeviceToHost, xs));
  TRR_CUDA(cudaMemcpyAsync(out_n, o_n, (size_t)B * 4, cudaMemcpyDeviceToHost, xs));
  TRR_CUDA(cudaEventRecord(g->ev_out[p], xs));
  if (blocking) TRR_CUDA(cudaStreamSynchronize(xs));
  return TRR_OK;
}

extern "C" int trr_hybrid_search_sharded(trr_group* g, trr_dense* dense, trr_bm25* bm25, const float* q,
                                         const uint32_t* q_terms, const uint32_t* q_off, uint32_t B, uint32_t C, int strategy,
                                         float param, uint32_t k, int use_dense, int use_sparse, uint32_t* out_ord,
                                         float* out_fused, float* out_dense, float* out_sparse, uint32_t* out_n) {
  return group_search_host(g, dense, bm25, q, q_terms, q_off, B, C, strategy, param, k, use_dense, use_sparse, out_ord,
                           out_fused, out_dense, out_sparse, out_n, true);
}

extern "C" int trr_hybrid_search_sharded_async(trr_group* g, trr_dense* dense, trr_bm25* bm25, const float* q,
                                               const uint32_t* q_terms, const uint32_t* q_off, uint32_t B, uint32_t C,
                                               int strategy, float param, uint32_t k, int use_dense, int use_sparse,
                                               uint32_t* out_ord, float* out_fused, float* out_dense, float* out_sparse,
                                               uint32_t* out_n) {
  return group_search_host(g, dense, bm25, q, q_terms, q_off, B, C, strategy, param, k, use_dense, use_sparse, out_ord,
                           out_fused, out_dense, out_sparse, out_n, false);
}
