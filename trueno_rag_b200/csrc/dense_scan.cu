// dense_scan.cu — strict-f32 dense kernels (compiled with -fmad=false -prec-div=true -prec-sqrt=true).
//
// Replaces the exhaustive scan of VectorStore::search (reference src/index.rs:386-412) and the
// distance functions it calls (src/index.rs:440-462).  Everything in this translation unit
// reproduces the reference's f32 arithmetic bit for bit: sums are sequential in dimension order,
// starting from 0.0, with separate multiply and add (no FMA).
//
//   K1  dense_scan_bulk_kernel     HBM-bound scan, one row per lane, rows staged into shared memory by
//                                  the TMA engine (cp.async.bulk, one bulk copy per row and chunk)
//       dense_scan_generic_kernel  any dimension / tiny stores (rows not 16-byte multiples)
//       dense_norms_kernel         per-row sqrt(sum x*x) in reference order + GEMM epilogue operands
//       topk_merge_kernel          merges per-warp partial top-k lists (canonical order)
//       rescore_select_kernel      exact rescoring + candidate proof for the tensor-core fast pass
#include <cuda.h>
#include <math_constants.h>
#include <string.h>

#include "common.cuh"
#include "dense.cuh"

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// warp-private bounded top-k buffer in shared memory (capacity = power of two >= k + 32)
// ---------------------------------------------------------------------------------------------
struct WarpTopK {
  uint64_t* buf;
  uint32_t cap, k, cnt;
  uint64_t thr;  // k-th best key so far (TRR_KEY_EMPTY while fewer than k were seen)

  __device__ __forceinline__ void init(uint64_t* b, uint32_t cap_, uint32_t k_) {
    buf = b; cap = cap_; k = k_; cnt = 0; thr = TRR_KEY_EMPTY;
  }
  __device__ __forceinline__ void compact(uint32_t lane) {
    for (uint32_t i = cnt + lane; i < cap; i += 32) buf[i] = TRR_KEY_EMPTY;
    trr_bitonic_sort_desc(buf, cap, lane, 32u, WarpSync());
    if (cnt > k) cnt = k;
    if (cnt == k && k > 0) thr = buf[k - 1];
    __syncwarp();
  }
  // every lane of the warp must call this (valid == false for lanes without a candidate)
  __device__ __forceinline__ void push(uint64_t key, bool valid, uint32_t lane) {
    bool pass = valid && key > thr;
    uint32_t m = __ballot_sync(FULL, pass);
    if (m == 0) return;
    if (cnt + __popc(m) > cap) {
      compact(lane);
      pass = pass && key > thr;
      m = __ballot_sync(FULL, pass);
      if (m == 0) return;
    }
    if (pass) buf[cnt + __popc(m & ((1u << lane) - 1u))] = key;
    cnt += __popc(m);
    __syncwarp();
  }
};

// ---------------------------------------------------------------------------------------------
// element access
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

template <int METRIC>
__device__ __forceinline__ void acc_step(float& acc, float q, float x) {
  if (METRIC == TRR_METRIC_EUCLIDEAN) {
    float t = q - x;       // (x - y) with x = query, y = stored vector   (src/index.rs:455)
    acc = acc + t * t;     // .powi(2) == t*t, then sequential sum
  } else {
    acc = acc + q * x;     // src/index.rs:441,461
  }
}

// src/index.rs:398-402 + :445-449 — turns the accumulated sum into the ranking score
template <int METRIC>
__device__ __forceinline__ float finish_score(float acc, float q_norm, float d_norm) {
  if (METRIC == TRR_METRIC_COSINE) {
    if (q_norm == 0.0f || d_norm == 0.0f) return 0.0f;
    return acc / (q_norm * d_norm);
  } else if (METRIC == TRR_METRIC_EUCLIDEAN) {
    return -sqrtf(acc);
  }
  return acc;
}

// merges the per-warp lists of a CTA into warp 0's buffer and writes one list per CTA
__device__ __forceinline__ void cta_merge_and_store(WarpTopK& tk, uint64_t* tk_base, uint32_t cap, uint32_t n_warps,
                                                    uint32_t warp, uint32_t lane, uint32_t* warp_cnt, uint64_t* out,
                                                    uint32_t* out_n) {
  if (lane == 0) warp_cnt[warp] = tk.cnt;
  __syncthreads();
  if (warp == 0) {
    for (uint32_t w = 1; w < n_warps; ++w) {
      const uint64_t* src = tk_base + (size_t)w * cap;
      const uint32_t cnt = warp_cnt[w];
      for (uint32_t i = 0; i < cnt; i += 32) {
        const bool valid = i + lane < cnt;
        tk.push(valid ? src[i + lane] : TRR_KEY_EMPTY, valid, lane);
      }
    }
    tk.compact(lane);
    for (uint32_t i = lane; i < tk.cnt; i += 32) out[i] = tk.buf[i];
    if (lane == 0) *out_n = tk.cnt;
  }
  __syncthreads();
}

}  // namespace

// =============================================================================================
// norms + GEMM epilogue operands
// =============================================================================================
template <int IS_BF16>
__global__ void dense_norms_kernel(const void* __restrict__ rows, uint32_t dim, uint64_t row0, uint64_t n_rows,
                                   float* __restrict__ norms) {
  uint64_t r = row0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= row0 + n_rows) return;
  float s = 0.0f;
  if (IS_BF16) {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(rows) + r * dim;
    for (uint32_t j = 0; j < dim; ++j) {
      float x = __uint_as_float(((uint32_t)p[j]) << 16);
      s = s + x * x;  // src/index.rs:442-443
    }
  } else {
    const float* p = reinterpret_cast<const float*>(rows) + r * dim;
    for (uint32_t j = 0; j < dim; ++j) {
      float x = p[j];
      s = s + x * x;
    }
  }
  norms[r] = sqrtf(s);
}

// fast-pass operands of the tensor-core path: fast = dot * scale + bias (see dense_gemm.cu)
__global__ void dense_gemm_operands_kernel(const float* __restrict__ norms, const uint8_t* __restrict__ dead,
                                           uint64_t n_rows, uint64_t n_padded, int metric,
                                           float2* __restrict__ scale_bias, float* __restrict__ max_norm) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float nrm = 0.0f;
  if (r < n_padded) {
    float2 sb = make_float2(0.0f, -CUDART_INF_F);  // padding rows and tombstones can never be selected
    if (r < n_rows && !(dead && dead[r])) {
      nrm = norms[r];
      if (metric == TRR_METRIC_COSINE) sb = make_float2(nrm == 0.0f ? 0.0f : 1.0f / nrm, 0.0f);
      else if (metric == TRR_METRIC_DOT) sb = make_float2(1.0f, 0.0f);
      else sb = make_float2(2.0f, -(nrm * nrm));
    }
    scale_bias[r] = sb;
  }
  // block max of the norms -> global max (norms are >= 0, so the int ordering of the bits is the float ordering)
  for (int o = 16; o > 0; o >>= 1) nrm = fmaxf(nrm, __shfl_xor_sync(FULL, nrm, o));
  if ((threadIdx.x & 31) == 0 && nrm > 0.0f) atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(nrm));
}

// query norms in reference order (src/index.rs:442), one thread per query
__global__ void dense_query_norms_kernel(const float* __restrict__ q, uint32_t dim, uint32_t B, float* __restrict__ qn) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* p = q + (uint64_t)b * dim;
  float s = 0.0f;
  for (uint32_t j = 0; j < dim; ++j) s = s + p[j] * p[j];
  qn[b] = sqrtf(s);
}

// =============================================================================================
// K1: bulk scan.  128 threads = 4 independent warps; each warp owns one shared-memory stage of
// 32 rows x ch_bytes (row pitch ch_bytes + 16 so that the per-lane LDS.128 are conflict-free),
// refilled by 32 TMA bulk copies that complete on the warp's mbarrier.
// =============================================================================================
template <int IS_BF16, int METRIC>
__global__ void __launch_bounds__(128, 1)
dense_scan_bulk_kernel(DenseScanArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t pitch = a.ch_bytes + 16;
  // layout: [query f32 (q_bytes)] [4 mbarriers] [4 x topk cap*8] [4 x stage 32*pitch]
  float* qs = reinterpret_cast<float*>(smem);
  const uint32_t q_bytes = (a.dim * 4 + 127) & ~127u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + q_bytes);
  constexpr uint32_t NWARPS = 4;
  uint64_t* tk_base = reinterpret_cast<uint64_t*>(smem + q_bytes + 128);
  uint64_t* tk_buf = tk_base + (size_t)warp * a.cap;
  uint32_t* warp_cnt = reinterpret_cast<uint32_t*>(smem + q_bytes + 64);
  uint8_t* stage = smem + q_bytes + 128 + (size_t)4 * a.cap * 8 + (size_t)warp * 32 * pitch;

  if (lane == 0) trr_mbar_init(&bars[warp], 1);
  trr_fence_mbar_init();
  __syncthreads();

  const uint32_t n_sel = a.n_sel_ptr ? *a.n_sel_ptr : a.n_sel;
  const uint64_t n_groups = (a.n_rows + 31) / 32;
  const uint64_t gwarp = (uint64_t)blockIdx.x * 4 + warp, gstride = (uint64_t)gridDim.x * 4;
  uint32_t parity = 0;
  uint8_t* my_row = stage + (size_t)lane * pitch;

  for (uint32_t si = 0; si < n_sel; ++si) {
    const uint32_t qi = a.sel ? a.sel[si] : si;
    __syncthreads();  // previous query's reads of qs are done
    for (uint32_t j = threadIdx.x; j < a.dim; j += blockDim.x) qs[j] = a.q[(uint64_t)qi * a.dim + j];
    __syncthreads();
    const float q_norm = a.q_norms[qi];
    WarpTopK tk;
    tk.init(tk_buf, a.cap, a.k);

    for (uint64_t g = gwarp; g < n_groups; g += gstride) {
      const uint64_t row = g * 32 + lane;
      const bool in_range = row < a.n_rows;
      float acc = 0.0f;
      for (uint32_t c = 0; c < a.n_chunks; ++c) {
        const uint32_t off = c * a.ch_bytes;
        const uint32_t bytes = min(a.ch_bytes, a.row_bytes - off);
        const uint64_t rows_left = a.n_rows - g * 32;
        const uint32_t valid_rows = rows_left < 32 ? (uint32_t)rows_left : 32u;
        trr_fence_proxy_async();  // order this warp's earlier generic reads of the stage before the async writes
        if (lane == 0) trr_mbar_expect_tx(&bars[warp], bytes * valid_rows);
        __syncwarp();
        if (in_range) trr_bulk_g2s(my_row, a.rows + row * a.row_bytes + off, bytes, &bars[warp]);
        trr_mbar_wait(&bars[warp], parity);
        parity ^= 1;
        if (in_range) {
          const uint4* rp = reinterpret_cast<const uint4*>(my_row);
          const uint32_t nvec = bytes >> 4;
          if (IS_BF16) {
            const float4* qp = reinterpret_cast<const float4*>(qs + (off >> 1));
#pragma unroll 4
            for (uint32_t i = 0; i < nvec; ++i) {
              const uint4 v = rp[i];
              const float4 q0 = qp[2 * i], q1 = qp[2 * i + 1];
              acc_step<METRIC>(acc, q0.x, bf16lo(v.x)); acc_step<METRIC>(acc, q0.y, bf16hi(v.x));
              acc_step<METRIC>(acc, q0.z, bf16lo(v.y)); acc_step<METRIC>(acc, q0.w, bf16hi(v.y));
              acc_step<METRIC>(acc, q1.x, bf16lo(v.z)); acc_step<METRIC>(acc, q1.y, bf16hi(v.z));
              acc_step<METRIC>(acc, q1.z, bf16lo(v.w)); acc_step<METRIC>(acc, q1.w, bf16hi(v.w));
            }
          } else {
            const float4* qp = reinterpret_cast<const float4*>(qs + (off >> 2));
#pragma unroll 4
            for (uint32_t i = 0; i < nvec; ++i) {
              const uint4 v = rp[i];
              const float4 qq = qp[i];
              acc_step<METRIC>(acc, qq.x, __uint_as_float(v.x)); acc_step<METRIC>(acc, qq.y, __uint_as_float(v.y));
              acc_step<METRIC>(acc, qq.z, __uint_as_float(v.z)); acc_step<METRIC>(acc, qq.w, __uint_as_float(v.w));
            }
          }
        }
        __syncwarp();
      }
      bool valid = in_range && !(a.dead && a.dead[row]);
      float score = 0.0f;
      if (valid) score = finish_score<METRIC>(acc, q_norm, METRIC == TRR_METRIC_COSINE ? a.norms[row] : 0.0f);
      tk.push(trr_make_key(score, a.base_ord + (uint32_t)row), valid, lane);
    }
    tk.compact(lane);
    cta_merge_and_store(tk, tk_base, a.cap, NWARPS, warp, lane, warp_cnt,
                        a.partial + ((uint64_t)si * gridDim.x + blockIdx.x) * a.k,
                        a.partial_n + ((uint64_t)si * gridDim.x + blockIdx.x));
  }
}

// K1, TMA ring: every warp streams its row groups through a private ring of 4 KB slots.  A slot holds one TMA box of
// 32 rows x 128 bytes (cp.async.bulk.tensor.2d, 128-byte swizzle), so lane r reads row r with conflict-free 128-bit
// loads (chunk j of row r sits at 16-byte position j ^ (r & 7)); an elected lane refills a slot as soon as the warp has
// consumed it, which keeps (n_slots - 1) x 4 KB per warp in flight at all times.  Rows past the end of the slab are
// zero-filled by the TMA unit.  The arithmetic per row is the same strict sequence as everywhere else in this file.
// NQ queries share one pass over the slab (NQ = 1 for a single query; NQ = 4 when several queries take the exact path:
// Euclidean or k > 100 batches, queries whose candidate proof failed): every row element is loaded once and feeds NQ
// independent strict chains, so four queries cost about one pass instead of four.
template <int IS_BF16, int METRIC, int NQ>
__global__ void __launch_bounds__(512, 1)
dense_scan_tma_kernel(const __grid_constant__ CUtensorMap map, DenseScanArgs a, const __grid_constant__ ScanQueryParam qp) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t SLOT = 4096;
  const uint32_t NWARPS = blockDim.x >> 5;
  constexpr uint32_t BOX_ELEMS = IS_BF16 ? 64 : 32;
  const uint32_t S = a.n_slots;
  uint8_t* smem = smem_dyn + ((1024u - (trr_smem_u32(smem_dyn) & 1023u)) & 1023u);  // swizzle atoms need 1 KB alignment
  uint8_t* ring = smem + (size_t)warp * S * SLOT;
  float* qs = reinterpret_cast<float*>(smem + (size_t)NWARPS * S * SLOT);
  const uint32_t q_bytes = (a.dim * 4 + 127) & ~127u;
  const uint32_t q_floats = q_bytes >> 2;
  uint8_t* after_q = reinterpret_cast<uint8_t*>(qs) + (size_t)NQ * q_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(after_q) + (size_t)warp * S;
  uint32_t* warp_cnt = reinterpret_cast<uint32_t*>(after_q + (size_t)NWARPS * S * 8);
  float* s_qn = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(warp_cnt) + 64);      // [4] query norms (fused prologue)
  uint32_t& s_last = *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(warp_cnt) + 96);  // fused epilogue: this CTA merges
  uint64_t* tk_base = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(warp_cnt) + 128);  // [NQ][NWARPS][cap]

  if (lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
    for (uint32_t s = 0; s < S; ++s) trr_mbar_init(&bars[s], 1);
  }
  trr_fence_mbar_init();
  __syncthreads();

  const uint32_t n_sel = a.n_sel_ptr ? *a.n_sel_ptr : a.n_sel;
  const uint32_t n_boxes = (a.row_bytes + 127) >> 7;
  const uint64_t n_groups = (a.n_rows + 31) / 32;
  const uint64_t gwarp = (uint64_t)blockIdx.x * NWARPS + warp, gstride = (uint64_t)gridDim.x * NWARPS;
  const uint64_t n_my = n_groups > gwarp ? (n_groups - gwarp + gstride - 1) / gstride : 0;
  const uint32_t sw = lane & 7;
  const uint8_t* lane_base = ring + lane * 128;
  uint32_t c_slot = 0, c_par = 0;  // consumer position in the ring (persists across queries)
  uint32_t i_slot = 0;             // producer position

  for (uint32_t si = 0; si < n_sel; si += NQ) {
    const uint32_t nq = min((uint32_t)NQ, n_sel - si);
    __syncthreads();  // previous group's reads of qs are done
    float q_norm[NQ];
    WarpTopK tk[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const uint32_t qi = (uint32_t)q < nq ? (a.sel ? a.sel[si + q] : si + q) : 0u;
      if (a.q_in_param) {  // (one query, q == 0: it arrived with the launch)
        for (uint32_t j = threadIdx.x; j < a.dim; j += blockDim.x) qs[q * q_floats + j] = q == 0 ? qp.v[j] : 0.0f;
        q_norm[q] = q == 0 ? qp.norm : 0.0f;
      } else {
        for (uint32_t j = threadIdx.x; j < a.dim; j += blockDim.x)
          qs[q * q_floats + j] = (uint32_t)q < nq ? a.q[(uint64_t)qi * a.dim + j] : 0.0f;
        q_norm[q] = ((uint32_t)q < nq && a.q_norms) ? a.q_norms[qi] : 0.0f;
      }
      tk[q].init(tk_base + ((size_t)q * NWARPS + warp) * a.cap, a.cap, a.k);
    }
    __syncthreads();
    if (!a.q_norms && !a.q_in_param && METRIC == TRR_METRIC_COSINE) {
      // |q| in the reference's order (src/index.rs:442): one thread per query of the group, sequential f32 sum + sqrt
      if (lane == 0 && warp < NQ) {
        float sq = 0.0f;
        for (uint32_t j = 0; j < a.dim; ++j) sq = sq + qs[warp * q_floats + j] * qs[warp * q_floats + j];
        s_qn[warp] = sqrtf(sq);
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < NQ; ++q) q_norm[q] = s_qn[q];
    }

    // producer state: next box to request = (group i_k, box i_c)
    uint64_t i_k = 0;
    uint32_t i_c = 0;
    auto issue = [&]() {  // lane 0 only
      if (i_k < n_my) {
        const uint64_t row0 = (gwarp + i_k * gstride) * 32;
        trr_fence_proxy_async();  // the warp's generic reads of this slot (ordered by __syncwarp) precede the async write
        trr_mbar_expect_tx(&bars[i_slot], SLOT);
        trr_tma_load_2d(&map, ring + (size_t)i_slot * SLOT, &bars[i_slot], (int32_t)(i_c * BOX_ELEMS), (int32_t)row0);
        if (++i_slot == S) i_slot = 0;
        if (++i_c == n_boxes) { i_c = 0; ++i_k; }
      }
    };
    if (lane == 0) for (uint32_t s = 0; s < S; ++s) issue();

    for (uint64_t k = 0; k < n_my; ++k) {
      const uint64_t row = (gwarp + k * gstride) * 32 + lane;
      float acc[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[q] = 0.0f;
      for (uint32_t c = 0; c < n_boxes; ++c) {
        trr_mbar_wait_bounded(&bars[c_slot], c_par);
        const uint8_t* rp = lane_base + (size_t)c_slot * SLOT;
        const uint32_t left = a.row_bytes - (c << 7);
        const uint32_t nvec = left >= 128 ? 8u : (left >> 4);
        if (IS_BF16) {
          auto step = [&](uint32_t j) {
            const uint4 v = *reinterpret_cast<const uint4*>(rp + ((j ^ sw) << 4));
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              const float4* qp = reinterpret_cast<const float4*>(qs + q * q_floats + c * 64);
              const float4 q0 = qp[2 * j], q1 = qp[2 * j + 1];
              acc_step<METRIC>(acc[q], q0.x, bf16lo(v.x)); acc_step<METRIC>(acc[q], q0.y, bf16hi(v.x));
              acc_step<METRIC>(acc[q], q0.z, bf16lo(v.y)); acc_step<METRIC>(acc[q], q0.w, bf16hi(v.y));
              acc_step<METRIC>(acc[q], q1.x, bf16lo(v.z)); acc_step<METRIC>(acc[q], q1.y, bf16hi(v.z));
              acc_step<METRIC>(acc[q], q1.z, bf16lo(v.w)); acc_step<METRIC>(acc[q], q1.w, bf16hi(v.w));
            }
          };
          if (nvec == 8) {
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) step(j);
          } else {
            for (uint32_t j = 0; j < nvec; ++j) step(j);
          }
        } else {
          auto step = [&](uint32_t j) {
            const uint4 v = *reinterpret_cast<const uint4*>(rp + ((j ^ sw) << 4));
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              const float4 qq = reinterpret_cast<const float4*>(qs + q * q_floats + c * 32)[j];
              acc_step<METRIC>(acc[q], qq.x, __uint_as_float(v.x)); acc_step<METRIC>(acc[q], qq.y, __uint_as_float(v.y));
              acc_step<METRIC>(acc[q], qq.z, __uint_as_float(v.z)); acc_step<METRIC>(acc[q], qq.w, __uint_as_float(v.w));
            }
          };
          if (nvec == 8) {
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) step(j);
          } else {
            for (uint32_t j = 0; j < nvec; ++j) step(j);
          }
        }
        __syncwarp();
        if (lane == 0) issue();  // refill the slot just consumed
        if (++c_slot == S) { c_slot = 0; c_par ^= 1; }
      }
      const bool valid = row < a.n_rows && !(a.dead && a.dead[row]);
      const float d_norm = (valid && METRIC == TRR_METRIC_COSINE) ? a.norms[row] : 0.0f;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        if ((uint32_t)q < nq) {  // uniform across the warp
          float score = 0.0f;
          if (valid) score = finish_score<METRIC>(acc[q], q_norm[q], d_norm);
          tk[q].push(trr_make_key(score, a.base_ord + (uint32_t)row), valid, lane);
        }
      }
    }
    // ring is empty here: every box requested for this group has been consumed; i_slot == c_slot
    i_slot = __shfl_sync(FULL, i_slot, 0);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if ((uint32_t)q < nq) {  // uniform across the CTA
        tk[q].compact(lane);
        cta_merge_and_store(tk[q], tk_base + (size_t)q * NWARPS * a.cap, a.cap, NWARPS, warp, lane, warp_cnt,
                            a.partial + ((uint64_t)(si + q) * gridDim.x + blockIdx.x) * a.k,
                            a.partial_n + ((uint64_t)(si + q) * gridDim.x + blockIdx.x));
      }
    }
    if (a.done) {
      // fused epilogue: the CTA that arrives last merges the gridDim.x lists of every query of the group and writes the
      // results - no merge kernel, no second launch
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_last = atomicAdd(&a.done[si / NQ], 1u) == gridDim.x - 1 ? 1u : 0u;
      __syncthreads();
      if (s_last) {
        __threadfence();
        for (uint32_t q = 0; q < nq; ++q) {
          const uint32_t out_row = a.sel ? a.sel[si + q] : si + q;
          const uint64_t* lists = a.partial + (uint64_t)(si + q) * gridDim.x * a.k;
          const uint32_t* list_n = a.partial_n + (uint64_t)(si + q) * gridDim.x;
          // every warp folds its share of the per-CTA lists into its top-k buffer, then the CTA merge used above leaves
          // the final list in this CTA's own slot of `partial` (its content has been folded already)
          WarpTopK m;
          m.init(tk_base + ((size_t)q * NWARPS + warp) * a.cap, a.cap, a.k);
          // (flat walk over the gridDim.x * k slots, four independent loads per lane in flight: the lists sit in L2)
          const uint32_t total = gridDim.x * a.k;
          for (uint32_t eb = warp * 32; eb < total; eb += NWARPS * 32 * 4) {  // (warp-uniform bound: push is collective)
            uint64_t key[4];
            bool valid[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t e = eb + lane + u * NWARPS * 32;
              valid[u] = false; key[u] = TRR_KEY_EMPTY;
              if (e < total) {
                const uint32_t l = e / a.k;
                valid[u] = (e - l * a.k) < list_n[l];
                key[u] = lists[e];
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) m.push(valid[u] ? key[u] : TRR_KEY_EMPTY, valid[u], lane);
          }
          m.compact(lane);
          uint64_t* dst = a.partial + ((uint64_t)(si + q) * gridDim.x + blockIdx.x) * a.k;
          uint32_t* dst_n = a.partial_n + ((uint64_t)(si + q) * gridDim.x + blockIdx.x);
          cta_merge_and_store(m, tk_base + (size_t)q * NWARPS * a.cap, a.cap, NWARPS, warp, lane, warp_cnt, dst, dst_n);
          const uint32_t n_out = *dst_n;
          for (uint32_t i = threadIdx.x; i < a.k; i += blockDim.x) {
            const bool ok = i < n_out;
            const uint64_t key = ok ? dst[i] : TRR_KEY_EMPTY;
            if (a.out_keys) a.out_keys[(uint64_t)out_row * a.k + i] = key;
            if (a.out_ord) a.out_ord[(uint64_t)out_row * a.k + i] = ok ? trr_key_ord(key) : 0xFFFFFFFFu;
            if (a.out_score) a.out_score[(uint64_t)out_row * a.k + i] = ok ? trr_key_score(key) : 0.0f;
          }
          if (threadIdx.x == 0 && a.out_n) a.out_n[out_row] = n_out;
          __syncthreads();
        }
        if (threadIdx.x == 0) a.done[si / NQ] = 0;  // ready for the next launch
      }
    }
  }
}

// generic scan: one row per thread straight from global memory; any dimension.  256 threads.
template <int IS_BF16, int METRIC>
__global__ void __launch_bounds__(256)
dense_scan_generic_kernel(DenseScanArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = reinterpret_cast<float*>(smem);
  const uint32_t q_bytes = (a.dim * 4 + 127) & ~127u;
  constexpr uint32_t NWARPS = 8;
  uint64_t* tk_base = reinterpret_cast<uint64_t*>(smem + q_bytes);
  uint64_t* tk_buf = tk_base + (size_t)warp * a.cap;
  __shared__ uint32_t warp_cnt[8];
  const uint32_t n_sel = a.n_sel_ptr ? *a.n_sel_ptr : a.n_sel;
  const uint64_t n_groups = (a.n_rows + 31) / 32;
  const uint64_t gwarp = (uint64_t)blockIdx.x * 8 + warp, gstride = (uint64_t)gridDim.x * 8;

  for (uint32_t si = 0; si < n_sel; ++si) {
    const uint32_t qi = a.sel ? a.sel[si] : si;
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < a.dim; j += blockDim.x) qs[j] = a.q[(uint64_t)qi * a.dim + j];
    __syncthreads();
    const float q_norm = a.q_norms[qi];
    WarpTopK tk;
    tk.init(tk_buf, a.cap, a.k);
    for (uint64_t g = gwarp; g < n_groups; g += gstride) {
      const uint64_t row = g * 32 + lane;
      bool valid = row < a.n_rows && !(a.dead && a.dead[row]);
      float score = 0.0f;
      if (valid) {
        float acc = 0.0f;
        if (IS_BF16) {
          const uint16_t* p = reinterpret_cast<const uint16_t*>(a.rows) + row * a.dim;
          for (uint32_t j = 0; j < a.dim; ++j) acc_step<METRIC>(acc, qs[j], __uint_as_float(((uint32_t)p[j]) << 16));
        } else {
          const float* p = reinterpret_cast<const float*>(a.rows) + row * a.dim;
          for (uint32_t j = 0; j < a.dim; ++j) acc_step<METRIC>(acc, qs[j], p[j]);
        }
        score = finish_score<METRIC>(acc, q_norm, METRIC == TRR_METRIC_COSINE ? a.norms[row] : 0.0f);
      }
      tk.push(trr_make_key(score, a.base_ord + (uint32_t)row), valid, lane);
    }
    tk.compact(lane);
    cta_merge_and_store(tk, tk_base, a.cap, NWARPS, warp, lane, warp_cnt,
                        a.partial + ((uint64_t)si * gridDim.x + blockIdx.x) * a.k,
                        a.partial_n + ((uint64_t)si * gridDim.x + blockIdx.x));
  }
}

// =============================================================================================
// merge of partial lists: one CTA (1024 threads) per output row.  lists[row][l][0..n[row][l]) hold keys.
//   small inputs (<= 4096 keys): load everything, one bitonic sort;
//   large inputs: the best 2048 keys live in the first half of a 4096-key buffer; the remaining keys stream
//   through a filter (key > current k-th best) into the second half, which is merged by a sort when it fills up.
// Writes ordinals/scores in canonical order.
// =============================================================================================
constexpr uint32_t MERGE_THREADS = 1024;
constexpr uint32_t MERGE_HALF = 2048;

__global__ void __launch_bounds__(MERGE_THREADS)
topk_merge_kernel(TopkMergeArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);
  __shared__ uint32_t s_fill, s_cnt;
  __shared__ uint64_t s_thr;
  const uint32_t tid = threadIdx.x;
  const uint32_t n_rows = a.n_rows_ptr ? *a.n_rows_ptr : a.n_rows;
  const uint32_t total = a.n_lists * a.list_stride;
  for (uint32_t ri = blockIdx.x; ri < n_rows; ri += gridDim.x) {
    const uint32_t out_row = a.row_map ? a.row_map[ri] : ri;
    const uint64_t* lists = a.lists + (uint64_t)ri * total;
    const uint32_t* list_n = a.list_n ? a.list_n + (uint64_t)ri * a.n_lists : nullptr;
    auto load_key = [&](uint32_t e) -> uint64_t {
      if (e >= total) return TRR_KEY_EMPTY;
      const uint32_t l = e / a.list_stride, j = e - l * a.list_stride;
      if (list_n && j >= list_n[l]) return TRR_KEY_EMPTY;
      return lists[e];
    };
    __syncthreads();
    uint32_t sorted_len;
    if (total <= 2 * MERGE_HALF) {
      sorted_len = trr_pow2_ceil(total < 2 ? 2 : total);
      for (uint32_t e = tid; e < sorted_len; e += MERGE_THREADS) buf[e] = load_key(e);
      trr_bitonic_sort_desc(buf, sorted_len, tid, MERGE_THREADS, BlockSync());
    } else {
      sorted_len = 2 * MERGE_HALF;
      for (uint32_t e = tid; e < sorted_len; e += MERGE_THREADS) buf[e] = TRR_KEY_EMPTY;
      if (tid == 0) { s_fill = 0; s_thr = TRR_KEY_EMPTY; }
      __syncthreads();
      uint32_t fill = 0;  // register copy of s_fill, uniform across the CTA (advanced by __syncthreads_count)
      for (uint32_t base = 0; base < total; base += MERGE_THREADS) {
        if (fill > MERGE_HALF - MERGE_THREADS) {
          trr_bitonic_sort_desc(buf, sorted_len, tid, MERGE_THREADS, BlockSync());
          for (uint32_t e = tid; e < MERGE_HALF; e += MERGE_THREADS) buf[MERGE_HALF + e] = TRR_KEY_EMPTY;
          if (tid == 0) { s_fill = 0; s_thr = buf[a.k - 1]; }
          fill = 0;
          __syncthreads();
        }
        const uint64_t key = load_key(base + tid);
        int pushed = 0;
        if (key > s_thr) { buf[MERGE_HALF + atomicAdd(&s_fill, 1u)] = key; pushed = 1; }
        fill += (uint32_t)__syncthreads_count(pushed);
      }
      trr_bitonic_sort_desc(buf, sorted_len, tid, MERGE_THREADS, BlockSync());
    }
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    uint32_t local = 0;
    for (uint32_t i = tid; i < a.k; i += MERGE_THREADS) {
      const uint64_t key = i < sorted_len ? buf[i] : TRR_KEY_EMPTY;
      const bool ok = key != TRR_KEY_EMPTY;
      local += ok;
      if (a.out_keys) a.out_keys[(uint64_t)out_row * a.k + i] = key;
      if (a.out_ord) a.out_ord[(uint64_t)out_row * a.k + i] = ok ? trr_key_ord(key) : 0xFFFFFFFFu;
      if (a.out_score) a.out_score[(uint64_t)out_row * a.k + i] = ok ? trr_key_score(key) : 0.0f;
    }
    if (local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (tid == 0 && a.out_n) a.out_n[out_row] = s_cnt;
  }
}

// =============================================================================================
// exact rescoring + candidate proof for the tensor-core fast pass (dense_gemm.cu).
// One CTA (256 threads) per query:
//   1. gather the query's fast candidates from every slice, keep the best CP by (fast, ordinal);
//   2. recompute those CP scores exactly as the reference does (one thread per candidate, sequential);
//   3. sort by the canonical key and emit the top k;
//   4. proof: every document that was NOT rescored has fast score <= f_min (the smallest fast score kept
//      by any full per-slice list, or the CP-th best after the merge); with |fast - exact| <= eps for every
//      document, none of them can reach or tie the k-th exact score if  f_min + eps < exact_k.  If the
//      proof fails the query is flagged and re-run through the exact scan (K1).
// =============================================================================================
// one term of the reference's sum: dot / cosine `acc + q*x` (src/index.rs:441,461), Euclidean `acc + (q-x)^2` (:455)
template <bool EU>
__device__ __forceinline__ float rs_step(float acc, float q, float x) {
  if (EU) { const float t = q - x; return acc + t * t; }
  return acc + q * x;
}

template <int IS_BF16, bool EU>
__global__ void __launch_bounds__(256)
rescore_select_kernel(RescoreArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);            // cap2 fast keys
  uint64_t* ekeys = keys + a.cap2;                                     // CP exact keys (CP power of two)
  __shared__ float s_gap, s_qnorm;
  __shared__ uint32_t s_total, s_tmax;
  __shared__ uint32_t s_smin[320], s_scnt[320];  // per candidate list (at most two half slices per SM): list minimum, fill count
  // first pass: CTA == query.  Second (wide) pass: CTA i re-scores the i-th query flagged by the first pass
  if (a.sel_n && blockIdx.x >= *a.sel_n) return;
  const uint32_t b = a.sel ? a.sel[blockIdx.x] : blockIdx.x;
  if (b >= a.B) return;
  const uint32_t tid = threadIdx.x;
  if (tid == 0) { s_gap = 0.0f; s_total = 0; s_tmax = 0; }
  for (uint32_t i = tid; i < a.cap2; i += blockDim.x) keys[i] = TRR_KEY_EMPTY;
  __syncthreads();
  // 1. gather: slice s of query block qb = b / 128, row r = b % 128
  const uint32_t qb = b / 128, r = b % 128;
  uint32_t local_total = 0;
  // a slice whose list is full may have dropped documents: all of them scored at most the list minimum or at most the
  // largest threshold shared between the slices (a.gthr: list minima and the helper warps' merged thresholds).  s_tmax = the
  // largest such bound; the per-slice minimum and fill count are collected with shared-memory atomics while the candidates
  // are gathered.
  for (uint32_t s = tid; s < a.n_slices; s += blockDim.x) { s_smin[s] = 0xFFFFFFFFu; s_scnt[s] = 0; }
  __syncthreads();
  for (uint32_t i = tid; i < a.n_slices * a.cps; i += blockDim.x) {
    const uint32_t s = i / a.cps, j = i % a.cps;
    const uint64_t base = (((uint64_t)s * a.n_qblocks + qb) * 128 + r) * a.cps + j;
    const float sc = a.cand_score[base];
    const uint32_t od = a.cand_ord[base];
    if (od != 0xFFFFFFFFu) {
      keys[i] = trr_make_key(sc, od);
      ++local_total;
      atomicMin(&s_smin[s], trr_f32_orderable(sc));
      atomicAdd(&s_scnt[s], 1u);
    }
  }
  if (local_total) atomicAdd(&s_total, local_total);
  __syncthreads();
  for (uint32_t s = tid; s < a.n_slices; s += blockDim.x)
    if (s_scnt[s] == a.cps) atomicMax(&s_tmax, s_smin[s]);
  if (tid == 0 && a.gthr) atomicMax(&s_tmax, a.gthr[b]);
  __syncthreads();
  trr_bitonic_sort_desc(keys, a.cap2, tid, blockDim.x, BlockSync());
  const uint32_t n_cand = min(s_total, a.cp);
  // 2. exact rescoring, strict reference order (one candidate per thread: the sum is a sequential chain; the loads are
  //    128-bit and run ahead of it).  The query is staged in shared memory once per CTA.
  float* qs = reinterpret_cast<float*>(ekeys + a.cp);
  const float* qv = a.q + (uint64_t)b * a.dim;
  for (uint32_t j = tid; j < a.dim; j += blockDim.x) qs[j] = qv[j];
  __syncthreads();
  if (a.q_norms_out && tid == blockDim.x - 1) {  // src/index.rs:442: sequential f32 sum of squares, then sqrt
    float s = 0.0f;
    for (uint32_t j = 0; j < a.dim; ++j) s = s + qs[j] * qs[j];
    s_qnorm = sqrtf(s);
    a.q_norms_out[b] = s_qnorm;
  }
  if (a.q_norms_out) __syncthreads();
  const float q_norm = a.q_norms_out ? s_qnorm : a.q_norms[b];
  // Staged path (rows of a multiple of 16 bytes): the candidate rows are scattered over HBM, so all 256 threads copy them
  // into shared memory with 16-byte cp.async (row pitch + 16 bytes: lane == candidate reads are conflict-free), chunk by
  // chunk; then one thread per candidate runs the strict sequential sum out of shared memory.
  const uint32_t row_bytes = a.dim * (IS_BF16 ? 2u : 4u);
  const bool staged = a.stage_chunk != 0 && (row_bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(a.rows) & 15u) == 0;
  float acc_staged = 0.0f;
  if (staged) {
    uint8_t* rowbuf = reinterpret_cast<uint8_t*>(qs) + ((a.dim * 4 + 15) & ~15u);
    const uint32_t pitch = a.stage_chunk + 16;
    for (uint32_t c0 = 0; c0 < row_bytes; c0 += a.stage_chunk) {
      const uint32_t cb = min(a.stage_chunk, row_bytes - c0), pieces = cb >> 4;
      for (uint32_t idx = tid; idx < n_cand * pieces; idx += blockDim.x) {
        const uint32_t r = idx / pieces, pc = idx - r * pieces;
        const uint64_t row = trr_key_ord(keys[r]) - a.base_ord;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.rows) + row * row_bytes + c0 + (pc << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(trr_smem_u32(rowbuf + r * pitch + (pc << 4))), "l"(src)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      if (tid < n_cand) {
        const uint4* rp = reinterpret_cast<const uint4*>(rowbuf + tid * pitch);
        float acc = acc_staged;
        if (IS_BF16) {
          const float* qq = qs + (c0 >> 1);
#pragma unroll 4
          for (uint32_t i = 0; i < pieces; ++i) {
            const uint4 v = rp[i];
            const float4 q0 = *reinterpret_cast<const float4*>(qq + 8 * i), q1 = *reinterpret_cast<const float4*>(qq + 8 * i + 4);
            acc = rs_step<EU>(acc, q0.x, bf16lo(v.x)); acc = rs_step<EU>(acc, q0.y, bf16hi(v.x));
            acc = rs_step<EU>(acc, q0.z, bf16lo(v.y)); acc = rs_step<EU>(acc, q0.w, bf16hi(v.y));
            acc = rs_step<EU>(acc, q1.x, bf16lo(v.z)); acc = rs_step<EU>(acc, q1.y, bf16hi(v.z));
            acc = rs_step<EU>(acc, q1.z, bf16lo(v.w)); acc = rs_step<EU>(acc, q1.w, bf16hi(v.w));
          }
        } else {
          const float* qq = qs + (c0 >> 2);
#pragma unroll 4
          for (uint32_t i = 0; i < pieces; ++i) {
            const uint4 v = rp[i];
            const float4 q0 = *reinterpret_cast<const float4*>(qq + 4 * i);
            acc = rs_step<EU>(acc, q0.x, __uint_as_float(v.x)); acc = rs_step<EU>(acc, q0.y, __uint_as_float(v.y));
            acc = rs_step<EU>(acc, q0.z, __uint_as_float(v.z)); acc = rs_step<EU>(acc, q0.w, __uint_as_float(v.w));
          }
        }
        acc_staged = acc;
      }
      __syncthreads();
    }
  }
  for (uint32_t c = tid; c < a.cp; c += blockDim.x) {  // one iteration unless this is the wide pass (cp > blockDim.x)
    uint64_t ek = TRR_KEY_EMPTY;
    float fast = 0.0f;
    if (c < n_cand) {
      const uint64_t fk = keys[c];
      const uint32_t od = trr_key_ord(fk);
      fast = trr_key_score(fk);
      const uint64_t row = od - a.base_ord;
      float acc = acc_staged;
      if (staged) {
        // sum already complete
      } else if (IS_BF16) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(a.rows) + row * a.dim;
        uint32_t j = 0;
        if ((a.dim & 7u) == 0 && (reinterpret_cast<uintptr_t>(a.rows) & 15u) == 0) {
          const uint4* p4 = reinterpret_cast<const uint4*>(p);
          // batches of 8 x 128-bit loads in flight (the rows are scattered over HBM: latency-bound), then the strict chain
          for (; j + 64 <= a.dim; j += 64) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(p4 + (j >> 3) + u);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float* qq = qs + j + u * 8;
              acc = rs_step<EU>(acc, qq[0], bf16lo(v[u].x)); acc = rs_step<EU>(acc, qq[1], bf16hi(v[u].x));
              acc = rs_step<EU>(acc, qq[2], bf16lo(v[u].y)); acc = rs_step<EU>(acc, qq[3], bf16hi(v[u].y));
              acc = rs_step<EU>(acc, qq[4], bf16lo(v[u].z)); acc = rs_step<EU>(acc, qq[5], bf16hi(v[u].z));
              acc = rs_step<EU>(acc, qq[6], bf16lo(v[u].w)); acc = rs_step<EU>(acc, qq[7], bf16hi(v[u].w));
            }
          }
          for (; j < a.dim; j += 8) {
            const uint4 v = __ldg(p4 + (j >> 3));
            acc = rs_step<EU>(acc, qs[j], bf16lo(v.x));     acc = rs_step<EU>(acc, qs[j + 1], bf16hi(v.x));
            acc = rs_step<EU>(acc, qs[j + 2], bf16lo(v.y)); acc = rs_step<EU>(acc, qs[j + 3], bf16hi(v.y));
            acc = rs_step<EU>(acc, qs[j + 4], bf16lo(v.z)); acc = rs_step<EU>(acc, qs[j + 5], bf16hi(v.z));
            acc = rs_step<EU>(acc, qs[j + 6], bf16lo(v.w)); acc = rs_step<EU>(acc, qs[j + 7], bf16hi(v.w));
          }
        }
        for (; j < a.dim; ++j) acc = rs_step<EU>(acc, qs[j], __uint_as_float(((uint32_t)p[j]) << 16));
      } else {
        const float* p = reinterpret_cast<const float*>(a.rows) + row * a.dim;
        uint32_t j = 0;
        if ((a.dim & 3u) == 0 && (reinterpret_cast<uintptr_t>(a.rows) & 15u) == 0) {
          const float4* p4 = reinterpret_cast<const float4*>(p);
          for (; j + 32 <= a.dim; j += 32) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(p4 + (j >> 2) + u);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float* qq = qs + j + u * 4;
              acc = rs_step<EU>(acc, qq[0], v[u].x); acc = rs_step<EU>(acc, qq[1], v[u].y);
              acc = rs_step<EU>(acc, qq[2], v[u].z); acc = rs_step<EU>(acc, qq[3], v[u].w);
            }
          }
          for (; j < a.dim; j += 4) {
            const float4 v = __ldg(p4 + (j >> 2));
            acc = rs_step<EU>(acc, qs[j], v.x);     acc = rs_step<EU>(acc, qs[j + 1], v.y);
            acc = rs_step<EU>(acc, qs[j + 2], v.z); acc = rs_step<EU>(acc, qs[j + 3], v.w);
          }
        }
        for (; j < a.dim; ++j) acc = rs_step<EU>(acc, qs[j], p[j]);
      }
      float score;
      if (a.metric == TRR_METRIC_COSINE) {
        const float dn = a.norms[row];
        score = (q_norm == 0.0f || dn == 0.0f) ? 0.0f : acc / (q_norm * dn);
      } else if (EU) {
        score = -sqrtf(acc);  // -euclidean_distance (src/index.rs:400,452-458)
      } else {
        score = acc;
      }
      ek = trr_make_key(score, od);
      // fast score in score units
      float fs = (a.metric == TRR_METRIC_COSINE) ? (q_norm > 0.0f ? fast / q_norm : 0.0f) : fast;
      if (EU) fs = -sqrtf(fmaxf(q_norm * q_norm - fast, 0.0f));  // the fast pass scores 2 q.d - |d|^2 = |q|^2 - |q - d|^2
      atomicMax(reinterpret_cast<int*>(&s_gap), __float_as_int(fabsf(fs - score)));
    }
    ekeys[c] = ek;
  }
  __syncthreads();
  trr_bitonic_sort_desc(ekeys, a.cp, tid, blockDim.x, BlockSync());
  // 3. emit
  const uint32_t n_out = min(n_cand, a.k);
  for (uint32_t i = tid; i < a.k; i += blockDim.x) {
    const bool ok = i < n_out;
    const uint64_t key = ok ? ekeys[i] : TRR_KEY_EMPTY;
    if (a.out_keys) a.out_keys[(uint64_t)b * a.k + i] = key;
    if (a.out_ord) a.out_ord[(uint64_t)b * a.k + i] = ok ? trr_key_ord(key) : 0xFFFFFFFFu;
    if (a.out_score) a.out_score[(uint64_t)b * a.k + i] = ok ? trr_key_score(key) : 0.0f;
  }
  // 4. proof
  if (tid == 0) {
    if (a.out_n) a.out_n[b] = n_out;
    bool ok = true;
    const bool something_excluded = a.n_live > (uint64_t)n_cand;
    if (something_excluded) {
      const bool bounded = s_total > a.cp || s_tmax != 0u;
      if (n_out < a.k || !bounded) {
        ok = false;  // fewer results than asked for, or documents vanished without a bound (non-finite scores)
      } else {
        // largest fast score any excluded document can have: the best gathered candidate that was not re-scored, or
        // the largest bound on what a slice dropped
        float f_excl = -CUDART_INF_F;
        if (s_total > a.cp) f_excl = trr_key_score(keys[a.cp]);
        if (s_tmax != 0u) f_excl = fmaxf(f_excl, trr_orderable_f32(s_tmax));
        const float f_excl_s = (a.metric == TRR_METRIC_COSINE) ? (q_norm > 0.0f ? f_excl / q_norm : 0.0f) : f_excl;
        float eps = a.eps_rel;                                  // cosine units
        if (a.metric == TRR_METRIC_COSINE) eps += (q_norm > 0.0f ? a.q_delta[b] / q_norm : 0.0f);
        else eps = (a.eps_rel * q_norm + a.q_delta[b]) * (*a.max_norm);
        const float exact_k = trr_key_score(ekeys[a.k - 1]);
        if (EU) {
          // Euclidean: the fast pass ranks by f = 2 q~.d - |d|^2 (q~ = bf16(q)), i.e. by -|q - d|^2 up to the per-query
          // constant |q|^2.  For a document that was not re-scored f <= f_excl, hence its true squared distance is at least
          //   D_lo = |q|^2 - f_excl - eps_D,  eps_D = 2 (eps_gemm |q~| + |q - q~|) max|d| + delta max|d|^2
          // (tensor-core accumulation + query quantisation through Cauchy-Schwarz, and the stored f32 norm squared);
          // the reference's f32 sum of squares is within delta of the true one, sqrt is monotone, so no such document
          // can reach or tie the k-th exact score -sqrt(D_k) when D_lo (1 - delta) > D_k (1 + 2^-20).
          const double delta = ((double)a.dim + 8.0) * 1.1920928955078125e-07;
          const double mn = (double)(*a.max_norm), qn = (double)q_norm, qd = (double)a.q_delta[b];
          const double eps_d = 2.0 * ((double)a.eps_rel * (qn + qd) + qd) * mn + delta * mn * mn;
          const double d_lo = qn * qn * (1.0 - delta) - (double)f_excl - eps_d;
          const double d_k = (double)exact_k * (double)exact_k;
          ok = d_lo > 0.0 && d_lo * (1.0 - delta) > d_k * (1.0 + 9.5367431640625e-07);
        } else {
          ok = (q_norm > 0.0f || a.metric != TRR_METRIC_COSINE) && (f_excl_s + eps < exact_k);
        }
      }
    }
    a.flags[b] = ok ? 0u : 1u;
    if (!ok) {
      const uint32_t slot = atomicAdd(a.n_flagged, 1u);
      a.flagged[slot] = b;
    }
    if (a.n_resolved && ok) atomicAdd(a.n_resolved, 1u);
    atomicMax(reinterpret_cast<int*>(a.max_gap), __float_as_int(s_gap));
  }
}

// ---------------------------------------------------------------------------------------------
// launch helpers (host)
// ---------------------------------------------------------------------------------------------
template <int IS_BF16>
static void launch_norms(const void* rows, uint32_t dim, uint64_t row0, uint64_t n, float* norms, cudaStream_t st) {
  if (n == 0) return;
  dense_norms_kernel<IS_BF16><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(rows, dim, row0, n, norms);
}

void trr_launch_norms(int is_bf16, const void* rows, uint32_t dim, uint64_t row0, uint64_t n, float* norms,
                      cudaStream_t st) {
  if (is_bf16) launch_norms<1>(rows, dim, row0, n, norms, st);
  else launch_norms<0>(rows, dim, row0, n, norms, st);
}

void trr_launch_gemm_operands(const float* norms, const uint8_t* dead, uint64_t n_rows, uint64_t n_padded, int metric,
                              float2* scale_bias, float* max_norm, cudaStream_t st) {
  if (n_padded == 0) return;
  dense_gemm_operands_kernel<<<(unsigned)((n_padded + 255) / 256), 256, 0, st>>>(norms, dead, n_rows, n_padded, metric,
                                                                                 scale_bias, max_norm);
}

void trr_launch_query_norms(const float* q, uint32_t dim, uint32_t B, float* qn, cudaStream_t st) {
  if (B == 0) return;
  dense_query_norms_kernel<<<(B + 127) / 128, 128, 0, st>>>(q, dim, B, qn);
}

template <int IS_BF16, int METRIC>
static cudaError_t launch_scan_t(const DenseScanArgs& a, bool bulk, unsigned grid, size_t smem, cudaStream_t st) {
  if (bulk) {
    auto kern = dense_scan_bulk_kernel<IS_BF16, METRIC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 128, smem, st>>>(a);
  } else {
    auto kern = dense_scan_generic_kernel<IS_BF16, METRIC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 256, smem, st>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t trr_launch_scan(const DenseScanArgs& a, int is_bf16, int metric, bool bulk, unsigned grid, size_t smem,
                            cudaStream_t st) {
#define TRR_SCAN_CASE(B, M) \
  if (is_bf16 == B && metric == M) return launch_scan_t<B, M>(a, bulk, grid, smem, st)
  TRR_SCAN_CASE(0, TRR_METRIC_COSINE);
  TRR_SCAN_CASE(0, TRR_METRIC_EUCLIDEAN);
  TRR_SCAN_CASE(0, TRR_METRIC_DOT);
  TRR_SCAN_CASE(1, TRR_METRIC_COSINE);
  TRR_SCAN_CASE(1, TRR_METRIC_EUCLIDEAN);
  TRR_SCAN_CASE(1, TRR_METRIC_DOT);
#undef TRR_SCAN_CASE
  return cudaErrorInvalidValue;
}

size_t trr_scan_tma_smem(uint32_t dim, uint32_t cap, uint32_t n_slots, uint32_t n_warps, uint32_t nq) {
  return 1024 + (size_t)n_warps * n_slots * 4096 + (size_t)nq * ((dim * 4 + 127) & ~127u) + (size_t)n_warps * n_slots * 8 + 128 +
         (size_t)nq * n_warps * cap * 8;  // (+128: per-warp counts, query norms and the last-CTA flag in front of the top-k buffers)
}

template <int IS_BF16, int METRIC, int NQ>
static cudaError_t launch_scan_tma_t(const DenseScanArgs& a, const void* map128, unsigned grid, unsigned n_warps, size_t smem,
                                     cudaStream_t st, const ScanQueryParam& qp) {
  CUtensorMap m;
  memcpy(&m, map128, 128);
  auto kern = dense_scan_tma_kernel<IS_BF16, METRIC, NQ>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, 32 * n_warps, smem, st>>>(m, a, qp);
  return cudaGetLastError();
}

cudaError_t trr_launch_scan_tma(const DenseScanArgs& a_in, const void* map128, int is_bf16, int metric, unsigned grid,
                                unsigned n_warps, uint32_t nq, size_t smem, cudaStream_t st, const float* host_q,
                                float host_q_norm) {
  DenseScanArgs a = a_in;
  static thread_local ScanQueryParam qp;  // (3 KB: kept off the stack; unused entries are never read)
  a.q_in_param = 0;
  if (host_q && a.dim <= TRR_SCAN_PARAM_DIM && nq == 1 && a.n_sel == 1 && !a.n_sel_ptr && !a.sel) {
    memcpy(qp.v, host_q, (size_t)a.dim * 4);
    qp.norm = host_q_norm;
    a.q_in_param = 1;
  }
#define TRR_SCAN_CASE(B, M)                                                                          \
  if (is_bf16 == B && metric == M)                                                                   \
    return nq == 4 ? launch_scan_tma_t<B, M, 4>(a, map128, grid, n_warps, smem, st, qp)              \
                   : launch_scan_tma_t<B, M, 1>(a, map128, grid, n_warps, smem, st, qp)
  TRR_SCAN_CASE(0, TRR_METRIC_COSINE);
  TRR_SCAN_CASE(0, TRR_METRIC_EUCLIDEAN);
  TRR_SCAN_CASE(0, TRR_METRIC_DOT);
  TRR_SCAN_CASE(1, TRR_METRIC_COSINE);
  TRR_SCAN_CASE(1, TRR_METRIC_EUCLIDEAN);
  TRR_SCAN_CASE(1, TRR_METRIC_DOT);
#undef TRR_SCAN_CASE
  return cudaErrorInvalidValue;
}

cudaError_t trr_launch_topk_merge(const TopkMergeArgs& a, unsigned grid, cudaStream_t st) {
  if (grid == 0) return cudaSuccess;
  size_t smem = (size_t)2 * MERGE_HALF * sizeof(uint64_t);
  cudaError_t e = cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  topk_merge_kernel<<<grid, MERGE_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

template <int IS_BF16, bool EU>
static cudaError_t launch_rescore_t(const RescoreArgs& a, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(rescore_select_kernel<IS_BF16, EU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  rescore_select_kernel<IS_BF16, EU><<<a.B, 256, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t trr_launch_rescore(const RescoreArgs& a, int is_bf16, cudaStream_t st) {
  if (a.B == 0) return cudaSuccess;
  size_t smem = (size_t)a.cap2 * 8 + (size_t)a.cp * 8 + (((size_t)a.dim * 4 + 15) & ~(size_t)15) + 16 +
                (a.stage_chunk ? (size_t)a.cp * (a.stage_chunk + 16) : 0);
  const bool eu = a.metric == TRR_METRIC_EUCLIDEAN;
  if (is_bf16) return eu ? launch_rescore_t<1, true>(a, smem, st) : launch_rescore_t<1, false>(a, smem, st);
  return eu ? launch_rescore_t<0, true>(a, smem, st) : launch_rescore_t<0, false>(a, smem, st);
}
