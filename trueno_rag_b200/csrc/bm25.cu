// bm25.cu — K3: BM25 posting-list scoring (compiled with -fmad=false -prec-div=true).
//
// Replaces BM25Index::search / score_term / term_frequency (reference src/index.rs:127-154, 212-243).
// The reference scores every candidate document by summing, over the query terms in query order
// (duplicates included), idf(term) * tf_norm(term, doc).  Adding 0.0 for a non-matching term is a no-op,
// so a term-at-a-time accumulation in query-term order gives bit-identical f32 sums.
//
// Device index (built once by bm25_build_kernel from the host CSR):
//   post[p]  = { local doc id, impact }  with impact = idf * (tf*(k1+1)) / (tf + k1*(1 - b + b*dl/avgdl)),
//              evaluated in exactly the reference's operation order (src/index.rs:147-153);
//   skip[t][r] = index of the first posting of term t whose doc id is >= r * R (R = documents per range),
//              r = 0..n_ranges, so the postings of term t inside range r are [skip[t][r], skip[t][r+1]).
//
// Search (bm25_search_kernel): persistent CTAs, one per SM, fed from a dynamic queue of work items
// (query, chunk of document ranges) ordered by decreasing posting volume (bm25_plan_kernel).
//   producer warp   walks the ranges of the item; per range it turns the skip-table entries of the query terms
//                   into one "pass": one cp.async.bulk (TMA 1-D) per term segment into a 2-stage shared-memory
//                   ring, completion on an mbarrier, plus a small descriptor (segment bounds inside the stage).
//   16 consumer warps  each owns a 1/16 sub-range of the R-document f32 accumulator in shared memory.  A warp
//                   binary-searches every staged segment (sorted by document) for its sub-range and accumulates
//                   the terms IN QUERY ORDER with plain read-modify-write: a document occurs once per term and
//                   belongs to exactly one warp, so there are no atomics and no block barriers between terms,
//                   and the f32 sum order equals the reference's.
//   harvest         after the last pass of a range each warp scans its accumulators (128-bit loads), keeps the
//                   documents whose (score, ordinal) key beats the query's running k-th best, and re-zeroes
//                   them.  Candidates go to a CTA-wide buffer that is compacted by a bitonic sort (named barrier
//                   over the consumer warps, once per range).
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"
#include "bm25.cuh"

namespace {

__device__ __forceinline__ uint32_t term_of_posting(const uint64_t* __restrict__ term_off, uint32_t n_terms, uint64_t p) {
  // last t with term_off[t] <= p
  uint32_t lo = 0, hi = n_terms;  // invariant: term_off[lo] <= p < term_off[hi]
  while (hi - lo > 1) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (term_off[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

}  // namespace

__global__ void bm25_build_kernel(Bm25BuildArgs a) {
  const uint64_t total = a.n_postings;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t t = term_of_posting(a.term_off, a.n_terms, p);
    // document ids come from the host CSR (build) or are already in place in post[].x (append / re-weighting)
    const uint32_t doc = a.post_doc ? a.post_doc[p] : a.post[p].x;
    // src/index.rs:137-153
    const float tf = (float)a.post_tf[p];
    const float doc_len = (float)a.doc_len[doc];
    const float idf = a.idf[t];
    const float tf_norm = (tf * (a.k1 + 1.0f)) / (tf + a.k1 * (1.0f - a.b + a.b * doc_len / a.avgdl));
    // tf == 0 marks the posting of a removed document (trr_bm25_remove): it weighs exactly +0.0, so the document can only
    // score 0.0 and is dropped by `score > 0.0`; it stays out of the per-term minimum used by the threshold bootstrap
    const bool dead = a.post_tf[p] == 0u;
    const float impact = dead ? 0.0f : idf * tf_norm;
    a.post[p] = make_uint2(doc, __float_as_uint(impact));
    if (!dead) {
      atomicMin(&a.term_min[t], trr_f32_orderable(impact));
      atomicMax(&a.term_max[t], trr_f32_orderable(impact));
      if (!(impact > 0.0f && impact < CUDART_INF_F)) a.flags[0] = 1u;  // NaN compares false
    }
    // skip table
    const uint64_t t_begin = a.term_off[t], t_end = a.term_off[t + 1];
    const uint32_t r = doc >> a.range_shift;
    const int64_t r_prev = (p == t_begin) ? -1 : (int64_t)((a.post_doc ? a.post_doc[p - 1] : a.post[p - 1].x) >> a.range_shift);
    uint32_t* row = a.skip + (uint64_t)t * a.skip_ld;
    for (int64_t rr = r_prev + 1; rr <= (int64_t)r; ++rr) row[rr] = (uint32_t)p;
    if (p + 1 == t_end)
      for (uint32_t rr = r + 1; rr <= a.n_ranges; ++rr) row[rr] = (uint32_t)t_end;
  }
}

// append: postings of term t in the new index = its old postings followed by the postings of the appended documents
// (their ordinals are larger, so the per-term doc order is kept).  One thread per new posting slot.
__global__ void bm25_merge_kernel(Bm25MergeArgs a) {
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.n_postings_new;
       p += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t t = term_of_posting(a.new_off, a.n_terms_new, p);
    const uint64_t j = p - a.new_off[t];
    const uint64_t old_len = t < a.n_terms_old ? a.old_off[t + 1] - a.old_off[t] : 0;
    uint32_t doc, tf;
    if (j < old_len) {
      const uint64_t q = a.old_off[t] + j;
      doc = a.old_post[q].x;
      tf = a.old_tf[q];
    } else {
      const uint64_t q = a.delta_off[t] + (j - old_len);
      doc = a.n_docs_old + a.delta_doc[q];
      tf = a.delta_tf[q];
    }
    a.new_post[p] = make_uint2(doc, 0u);
    a.new_tf[p] = tf;
  }
}

// remove: zeroes the term frequency of every posting whose document is in the removed set (bitmap over local doc ids)
__global__ void bm25_kill_kernel(const uint2* __restrict__ post, uint32_t* __restrict__ tf, uint64_t n_postings,
                                 const uint32_t* __restrict__ dead_bits, uint32_t* __restrict__ n_killed) {
  uint32_t local = 0;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_postings; p += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t d = post[p].x;
    if (((dead_bits[d >> 5] >> (d & 31)) & 1u) && tf[p] != 0u) { tf[p] = 0u; ++local; }
  }
  if (local) atomicAdd(n_killed, local);
}

// trr_bm25_load: a snapshot's skip rows and document ids become device indices of the search kernels (bulk-copy sources,
// accumulator cells), so a file that passes the header checks is still verified here: every skip row is a monotone partition
// of its term's postings, every document id is inside the shard.  *bad counts the violations.
__global__ void bm25_check_kernel(const uint32_t* __restrict__ skip, uint32_t skip_ld, uint32_t n_terms, uint32_t n_ranges,
                                  const uint64_t* __restrict__ term_off, const uint2* __restrict__ post, uint64_t n_postings,
                                  uint32_t n_docs, uint32_t* __restrict__ bad) {
  const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t local = 0;
  for (uint64_t t = i0; t < n_terms; t += stride) {
    const uint32_t* row = skip + t * skip_ld;
    uint64_t prev = term_off[t];
    if (row[0] != prev || row[n_ranges] != term_off[t + 1]) ++local;
    for (uint32_t r = 1; r <= n_ranges; ++r) {
      const uint32_t v = row[r];
      if (v < prev || v > n_postings) { ++local; break; }
      prev = v;
    }
  }
  for (uint64_t p = i0; p < n_postings; p += stride)
    if (post[p].x >= n_docs) ++local;
  if (local) atomicAdd(bad, local);
}

// terms without postings: every boundary is the (empty) term's offset
__global__ void bm25_skip_empty_kernel(Bm25BuildArgs a) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n_terms) return;
  if (a.term_off[t] != a.term_off[t + 1]) return;
  uint32_t* row = a.skip + (uint64_t)t * a.skip_ld;
  for (uint32_t rr = 0; rr <= a.n_ranges; ++rr) row[rr] = (uint32_t)a.term_off[t];
}

// =============================================================================================
// planning: posting volume per query -> processing order (largest first), bootstrap thresholds, queue reset
// =============================================================================================
// one warp per query, one lane per query term (strided): the skip-row look-ups of all terms are in flight together
__global__ void __launch_bounds__(256)
bm25_cost_kernel(Bm25SearchArgs a, uint64_t* __restrict__ keys, uint32_t cap2) {
  const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x < 8) a.queue[threadIdx.x] = 0;  // work queues and flagged-query counters of the search
  const uint32_t n_loop = cap2 ? cap2 : a.B;
  if (b >= n_loop) return;
  if (b >= a.B) {  // padding entries of the sort
    if (lane == 0) keys[b] = 0;
    return;
  }
  // Threshold bootstrap.  Impacts are positive and f32 addition of non-negative values is monotone, so a document that
  // contains term t scores at least min_impact(t).  If t has >= k postings in this shard, at least k documents have a
  // key above K0 = key(min_impact(t), worst ordinal), hence the k-th best key is >= K0 and everything <= K0 - 1 can be
  // dropped from the first range on (instead of flooding the candidate buffer until the running top-k fills up).
  const bool boot = a.n_chunks == 1 && a.flags[0] == 0u;
  uint64_t cost = 0;  // postings the query touches in this shard
  uint32_t best = 0;  // largest min-impact among its terms with >= k postings
  float smax = 0.0f;  // sum over the term slots of the term's largest impact: bounds every document's score
  float mmax = 0.0f;  // largest single impact
  const uint32_t q0 = a.q_off[b], T = a.q_off[b + 1] - q0;
  for (uint32_t i = q0 + lane; i < q0 + T; i += 32) {
    const uint32_t t = a.q_terms[i];
    if (t < a.n_terms) {
      const uint32_t* row = a.skip + (uint64_t)t * a.skip_ld;
      const uint32_t cnt = row[a.n_ranges] - row[0];
      cost += cnt;
      if (boot && cnt >= a.k) best = max(best, a.term_min[t]);
      const uint32_t mx = a.term_max[t];
      if (mx) { const float m = trr_orderable_f32(mx); smax += m; mmax = fmaxf(mmax, m); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cost += __shfl_xor_sync(0xFFFFFFFFu, cost, o);
    best = max(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
    smax += __shfl_xor_sync(0xFFFFFFFFu, smax, o);
    mmax = fmaxf(mmax, __shfl_xor_sync(0xFFFFFFFFu, mmax, o));
  }
  // Fixed-point scales of the integer fast pass: every posting adds ceil(impact * scale) to a cell.  A cell never exceeds
  // scale * smax + T, which has to fit the cell (2^16 - 1 / 2^31), and a single term stays below 2^22 (the FFMA trick).
  float scale16 = 1.0f, scale32 = 1.0f;
  if (smax > 0.0f) {
    // (0.9999: smax is a float tree sum; 1e30: a denormal-sized smax must not turn the scale into infinity)
    scale16 = fminf((65535.0f - (float)T) / smax * 0.9999f, 1e30f);
    scale32 = fminf(fminf(4194304.0f / mmax, 2147483000.0f / smax) * 0.9999f, 1e30f);
  }
  uint32_t bestf16 = 0, bestf32 = 0;
  if (a.thr0f16 && boot) {  // the same bootstrap in the fixed-point domains, for the kf candidates of the fast pass
    for (uint32_t i = q0 + lane; i < q0 + T; i += 32) {
      const uint32_t t = a.q_terms[i];
      if (t < a.n_terms) {
        const uint32_t* row = a.skip + (uint64_t)t * a.skip_ld;
        if (row[a.n_ranges] - row[0] >= a.kf) {
          const float mn = trr_orderable_f32(a.term_min[t]);
          bestf16 = max(bestf16, __float_as_uint(__fmaf_ru(mn, scale16, 8388608.0f)) & 0x7FFFFFu);
          bestf32 = max(bestf32, __float_as_uint(__fmaf_ru(mn, scale32, 8388608.0f)) & 0x7FFFFFu);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bestf16 = max(bestf16, __shfl_xor_sync(0xFFFFFFFFu, bestf16, o));
      bestf32 = max(bestf32, __shfl_xor_sync(0xFFFFFFFFu, bestf32, o));
    }
  }
  if (lane == 0) {
    a.thr0[b] = best ? (((uint64_t)best << 32) - 1ull) : TRR_KEY_EMPTY;
    if (a.thr0f16) { a.qscale16[b] = scale16; a.thr0f16[b] = bestf16; a.qscale32[b] = scale32; a.thr0f32[b] = bestf32; }
    if (cost > 0xFFFFFFFFull) cost = 0xFFFFFFFFull;
    if (cap2) keys[b] = ((cost + 1) << 32) | (uint64_t)(0xFFFFFFFFu - b);  // never TRR_KEY_EMPTY; ties: smaller b first
    else a.order[b] = b;
  }
}

__global__ void __launch_bounds__(1024, 1)
bm25_order_kernel(Bm25SearchArgs a, const uint64_t* __restrict__ gkeys, uint32_t cap2) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // cap2 = power of two >= B
  const uint32_t tid = threadIdx.x;
  for (uint32_t i = tid; i < cap2; i += blockDim.x) keys[i] = gkeys[i];
  trr_bitonic_sort_desc(keys, cap2, tid, blockDim.x, BlockSync());
  for (uint32_t i = tid; i < a.B; i += blockDim.x) a.order[i] = 0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFu);
}

// =============================================================================================
// search
// =============================================================================================
namespace {

constexpr uint32_t FULLM = 0xFFFFFFFFu;
constexpr uint32_t CW = TRR_BM25_CONSUMER_WARPS;       // 16
constexpr uint32_t CT = CW * 32;                       // consumer threads
constexpr uint32_t F_HARVEST = 1u, F_END_ITEM = 2u, F_QUIT = 4u, F_HAS_POSTINGS = 8u;

struct PassDesc {
  uint32_t flags;
  uint32_t range_base;  // local id of the first document of the range
  uint32_t item;
  uint32_t pad;
  uint64_t thr0;           // the item's bootstrap threshold (0 = none)
  uint32_t seg_begin[32];  // per term slot of the pass: [begin, end) inside the stage buffer
  uint32_t seg_end[32];
};

template <bool BACKOFF = false>  // BACKOFF: a waiting producer warp sleeps between polls instead of taking issue slots
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity, uint32_t site, uint32_t* dbg) {
  if (trr_mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!trr_mbar_try_wait(bar, parity)) {
    if (BACKOFF) __nanosleep(100);
    if (clock64() - t0 > 4000000000LL) {  // ~2 s: a protocol bug must surface as an error, never as a hung GPU
      if (dbg) { *dbg = 0x40000000u | (site << 16) | (blockIdx.x & 0xFFFFu); __threadfence_system(); }
      __trap();
    }
  }
}

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory"); }
struct ConsumerSync { __device__ __forceinline__ void operator()() const { consumer_bar(); } };
// the same barrier with an OR-reduction of a predicate over the consumer threads
__device__ __forceinline__ bool consumer_bar_or(bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 p, %1, 0;\n\t"
      "barrier.red.or.pred q, 1, %2, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"((uint32_t)pred), "n"(CT)
      : "memory");
  return r != 0;
}

__device__ __forceinline__ uint32_t lower_bound_doc(const uint2* st, uint32_t lo, uint32_t hi, uint32_t doc) {
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (st[mid].x < doc) lo = mid + 1; else hi = mid;
  }
  return lo;
}

}  // namespace

#ifndef TRR_BM25_MIN_CTAS
#define TRR_BM25_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(TRR_BM25_THREADS, TRR_BM25_MIN_CTAS)
bm25_search_kernel(Bm25SearchArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t R = 1u << a.range_shift;
  const uint32_t SUB = R / CW;                                                   // documents owned by one consumer warp
  float* acc = reinterpret_cast<float*>(smem_raw);                                // R
  uint2* stage_buf = reinterpret_cast<uint2*>(acc + R);                           // 2 x stage_cap
  uint64_t* cand = reinterpret_cast<uint64_t*>(stage_buf + 2 * (size_t)a.stage_cap);  // cand_cap
  PassDesc* desc = reinterpret_cast<PassDesc*>(cand + a.cand_cap);                // 2
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(desc + 2);                     // 2
  uint64_t* empty_bar = full_bar + 2;                                             // 2
  uint32_t* bnd_all = reinterpret_cast<uint32_t*>(empty_bar + 2);                 // 2 x 32 x 17 sub-range boundaries
  __shared__ uint32_t s_cnt, s_overflow, s_ovf_latched;
  __shared__ uint64_t s_thr;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { trr_mbar_init(&full_bar[s], 1); trr_mbar_init(&empty_bar[s], CW); }
    trr_fence_mbar_init();
    s_cnt = 0; s_overflow = 0; s_thr = TRR_KEY_EMPTY;
  }
  if (warp < CW) for (uint32_t i = tid; i < R; i += CT) acc[i] = 0.0f;
  __syncthreads();

  if (warp == CW) {
    // ============================ producer warp ============================
    uint32_t stage = 0, phase = 0;
    // publishes one pass: descriptor + bulk copies.  Every lane passes its own slot (len == 0: not in the pass).
    uint64_t item_thr0 = TRR_KEY_EMPTY;
    auto emit = [&](uint32_t flags, uint32_t range_base, uint32_t item, uint32_t off, uint32_t src_al, uint32_t al,
                    uint32_t begin, uint32_t end, uint32_t total_al) {
      mbar_wait_or_trap(&empty_bar[stage], phase ^ 1, 1, a.dbg);
      PassDesc& d = desc[stage];
      d.seg_begin[lane] = begin;
      d.seg_end[lane] = end;
      if (lane == 0) {
        d.flags = flags | (total_al ? F_HAS_POSTINGS : 0u);
        d.range_base = range_base; d.item = item; d.thr0 = item_thr0;
      }
      __syncwarp();
      if (lane == 0) {
        if (total_al) trr_mbar_expect_tx(&full_bar[stage], total_al * 8u);
        else trr_mbar_arrive(&full_bar[stage]);
      }
      __syncwarp();
      if (al) trr_bulk_g2s(stage_buf + (size_t)stage * a.stage_cap + off, a.post + src_al, al * 8u, &full_bar[stage]);
      if (++stage == 2) { stage = 0; phase ^= 1; }
    };
    const uint32_t n_items = (a.n_sel_ptr ? *a.n_sel_ptr : a.B) * a.n_chunks;
    while (true) {
      uint32_t item = 0;
      if (lane == 0) item = atomicAdd(a.queue, 1u);
      item = __shfl_sync(FULLM, item, 0);
      if (item >= n_items) { emit(F_QUIT, 0, item, 0, 0, 0, 0, 0, 0); break; }
      const uint32_t b = a.order[item / a.n_chunks], c = item % a.n_chunks;
      item_thr0 = a.n_chunks == 1 ? a.thr0[b] : TRR_KEY_EMPTY;  // (the bootstrap counts the postings of the whole shard)
      const uint32_t r0 = (uint32_t)(((uint64_t)c * a.n_ranges) / a.n_chunks);
      const uint32_t r1 = (uint32_t)(((uint64_t)(c + 1) * a.n_ranges) / a.n_chunks);
      const uint32_t q0 = a.q_off[b];
      const uint32_t T = a.q_off[b + 1] - q0;
      const uint32_t G = (T + 31) >> 5;
      // G == 1: every lane keeps a cursor through the skip row of its term, fetched one range ahead
      const uint32_t* row1 = nullptr;
      uint32_t c_s = 0, c_e = 0, c_n = 0;
      if (G == 1 && lane < T) {
        const uint32_t term = a.q_terms[q0 + lane];
        if (term < a.n_terms) {
          row1 = a.skip + (uint64_t)term * a.skip_ld;
          c_s = row1[r0];
          c_e = r0 < r1 ? row1[r0 + 1] : c_s;
          c_n = r0 + 2 <= a.n_ranges ? row1[r0 + 2] : c_e;
        }
      }
      for (uint32_t r = r0; r < r1; ++r) {
        const uint32_t range_base = r << a.range_shift;
        bool pending_harvest = false;  // a pass of this range was emitted without the harvest flag
        for (uint32_t g = 0; g < G; ++g) {
          uint32_t s = 0, e = 0;
          if (G == 1) {
            s = c_s; e = c_e;
            c_s = c_e; c_e = c_n;
            if (row1 && r + 3 <= a.n_ranges) c_n = row1[r + 3];
          } else {
            const uint32_t ti = g * 32 + lane;
            if (ti < T) {
              const uint32_t term = a.q_terms[q0 + ti];
              if (term < a.n_terms) {
                const uint32_t* row = a.skip + (uint64_t)term * a.skip_ld;
                s = row[r]; e = row[r + 1];
              }
            }
          }
          uint32_t first = 0;  // slots below `first` are done
          while (first < 32) {
            const uint32_t len = lane >= first ? e - s : 0u;
            const uint32_t al = len ? (((s & 1u) + len + 1u) & ~1u) : 0u;  // postings copied: 16-byte aligned both ends
            uint32_t incl = al;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const uint32_t v = __shfl_up_sync(FULLM, incl, o);
              if ((int)lane >= o) incl += v;
            }
            const uint32_t total = __shfl_sync(FULLM, incl, 31);
            if (total == 0) break;
            const uint32_t fits = __ballot_sync(FULLM, incl <= a.stage_cap);
            const uint32_t n_fit = fits == FULLM ? 32u : (uint32_t)(__ffs(~fits) - 1);  // slots [first, n_fit) fit
            if (n_fit == first) {
              // slot `first` alone exceeds the stage: stream it in stage-sized pieces (same term: any order is exact)
              const bool me = lane == first;
              const uint32_t s_al = s & ~1u;
              const uint32_t end_idx = min(e, s_al + a.stage_cap);
              emit(0, range_base, item, 0, me ? s_al : 0u, me ? a.stage_cap : 0u, me ? (s & 1u) : 0u,
                   me ? end_idx - s_al : 0u, a.stage_cap);
              if (me) s = end_idx;
              pending_harvest = true;
            } else {
              const bool me = lane >= first && lane < n_fit;
              const uint32_t taken = __shfl_sync(FULLM, incl, n_fit - 1);
              const bool last = (G == 1) && (taken == total);
              const uint32_t off = incl - al;
              emit(last ? F_HARVEST : 0u, range_base, item, me ? off : 0u, me ? (s & ~1u) : 0u, me ? al : 0u,
                   me ? off + (s & 1u) : 0u, me ? off + (s & 1u) + len : 0u, taken);
              pending_harvest = !last;
              first = n_fit;
            }
          }
        }
        if (pending_harvest) emit(F_HARVEST, range_base, item, 0, 0, 0, 0, 0, 0);
      }
      emit(F_END_ITEM, 0, item, 0, 0, 0, 0, 0, 0);
    }
  } else {
    // ============================ consumer warps ============================
    uint32_t stage = 0, phase = 0;
    bool touched = false;  // this warp accumulated something in the current range
    float* my_acc = acc + warp * SUB;
    const uint32_t compact_at = a.k + ((a.cand_cap - a.k) >> 1);
    // block-wide (consumer warps) compaction: afterwards cand[0..s_cnt) is sorted descending and s_thr is the k-th best
    auto compact = [&]() {
      consumer_bar();
      const uint32_t cnt = min(s_cnt, a.cand_cap);
      for (uint32_t i = cnt + tid; i < a.cand_cap; i += CT) cand[i] = TRR_KEY_EMPTY;
      trr_bitonic_sort_desc(cand, a.cand_cap, tid, CT, ConsumerSync());
      if (tid == 0) {
        const uint32_t c2 = min(cnt, a.k);
        s_cnt = c2;
        s_ovf_latched = s_overflow;
        s_overflow = 0;
        if (c2 == a.k && a.k > 0) s_thr = cand[a.k - 1];
      }
      consumer_bar();
    };
    while (true) {
      mbar_wait_or_trap(&full_bar[stage], phase, 2, a.dbg);
      const PassDesc& d = desc[stage];
      const uint32_t flags = d.flags, range_base = d.range_base, item = d.item;
      const uint64_t thr0 = d.thr0;
      if (flags & F_QUIT) break;
      const uint2* st = stage_buf + (size_t)stage * a.stage_cap;
      {
        // sub-range boundaries of every staged segment, computed once by the 512 consumer threads together:
        // thread (term slot l = tid / 16, sub-range w = tid % 16) finds the first posting of slot l with doc >= start of
        // sub-range w; warp w then reads its [lo, hi) per slot from shared memory (one search per boundary instead of
        // two per (warp, slot)).  The table is double-buffered by stage.
        uint32_t* bnd = bnd_all + stage * (32 * 17);
        if (flags & F_HAS_POSTINGS) {
          const uint32_t l = tid >> 4, w = tid & 15;
          uint32_t lo_s = d.seg_begin[l], hi_s = d.seg_end[l];
          const uint32_t dw = range_base + w * SUB;
          if (w == 0) bnd[l * 17 + 16] = hi_s;
          while (lo_s < hi_s) {
            const uint32_t mid = (lo_s + hi_s) >> 1;
            if (st[mid].x < dw) lo_s = mid + 1; else hi_s = mid;
          }
          bnd[l * 17 + w] = lo_s;
          consumer_bar();
        }
        uint32_t lo = 0, hi = 0;
        if (flags & F_HAS_POSTINGS) { lo = bnd[lane * 17 + warp]; hi = bnd[lane * 17 + warp + 1]; }
        // Measured and dropped (4M documents, 1024 queries; this loop: 4.0 ms): a flattened walk (all slots of the warp as
        // one sequence, 32 postings per step, match.any for documents that occur in two slots of a step) 6.7 ms, and 4.1 ms
        // even without the conflict handling; four slots per round with the posting loads in flight together 4.7 ms.
        uint32_t m = __ballot_sync(FULLM, hi > lo);
        if (m) touched = true;
        while (m) {  // ascending slot == query-term order
          const uint32_t l = __ffs(m) - 1;
          m &= m - 1;
          const uint32_t lo_ = __shfl_sync(FULLM, lo, l), hi_ = __shfl_sync(FULLM, hi, l);
#pragma unroll 1
          for (uint32_t p = lo_ + lane; p < hi_; p += 32) {
            const uint2 e0 = st[p];
            float* p0 = acc + (e0.x - range_base);
            *p0 = *p0 + __uint_as_float(e0.y);
          }
          __syncwarp();
        }
      }
      __syncwarp();
      if (lane == 0) trr_mbar_arrive(&empty_bar[stage]);  // stage and descriptor are free again
      if (++stage == 2) { stage = 0; phase ^= 1; }

      if (flags & F_HARVEST) {
        while (true) {
          bool need_compact = false;
          const uint64_t thr = max(*reinterpret_cast<volatile uint64_t*>(&s_thr), thr0);
          const float thr_f = thr == TRR_KEY_EMPTY ? -CUDART_INF_F : trr_key_score(thr);
          if (touched) {
            uint4* a4 = reinterpret_cast<uint4*>(my_acc);
            const uint32_t ord0 = a.doc_base + range_base + warp * SUB;
            auto harvest4 = [&](uint32_t i, uint4 v) {
              if ((v.x | v.y | v.z | v.w) == 0u) return;
              const float mx = fmaxf(fmaxf(__uint_as_float(v.x), __uint_as_float(v.y)),
                                     fmaxf(__uint_as_float(v.z), __uint_as_float(v.w)));
              if (!(mx >= thr_f)) {  // nothing here can enter the top-k (NaN compares false): just re-zero
                a4[i] = make_uint4(0u, 0u, 0u, 0u);
                return;
              }
#pragma unroll 1
              for (int j = 0; j < 4; ++j) {
                const float f = my_acc[i * 4 + j];
                bool keep = false;
                if (f >= thr_f && f > 0.0f) {  // src/index.rs:236 keeps only score > 0.0
                  const uint64_t key = trr_make_key(f, ord0 + i * 4 + j);
                  if (key > thr) {
                    const uint32_t pos = atomicAdd(&s_cnt, 1u);
                    if (pos + 1u >= compact_at) need_compact = true;
                    if (pos < a.cand_cap) cand[pos] = key;
                    else { s_overflow = 1u; keep = true; }  // stays in the accumulator; retried after the compaction
                  }
                }
                if (!keep) my_acc[i * 4 + j] = 0.0f;
              }
            };
            const uint32_t n4 = SUB >> 2;  // multiple of 32 (SUB >= 128)
            // branch-light scan: re-zero the touched cells that cannot enter the top-k with a predicated store; the
            // per-element path runs only when some lane of the warp holds a cell at or above the threshold
            auto pre4 = [&](uint32_t i, const uint4& v) -> bool {
              const bool nz = (v.x | v.y | v.z | v.w) != 0u;
              const float mx = fmaxf(fmaxf(__uint_as_float(v.x), __uint_as_float(v.y)),
                                     fmaxf(__uint_as_float(v.z), __uint_as_float(v.w)));
              const bool hit = mx >= thr_f;  // NaN compares false
              if (nz && !hit) a4[i] = make_uint4(0u, 0u, 0u, 0u);
              return nz && hit;
            };
            uint32_t i = lane;
            for (; i + 96 < n4; i += 128) {
              const uint4 v0 = a4[i], v1 = a4[i + 32], v2 = a4[i + 64], v3 = a4[i + 96];
              const bool h0 = pre4(i, v0), h1 = pre4(i + 32, v1), h2 = pre4(i + 64, v2), h3 = pre4(i + 96, v3);
              if (__any_sync(FULLM, h0 | h1 | h2 | h3)) {
                if (h0) harvest4(i, v0);
                if (h1) harvest4(i + 32, v1);
                if (h2) harvest4(i + 64, v2);
                if (h3) harvest4(i + 96, v3);
              }
            }
            for (; i < n4; i += 32) harvest4(i, a4[i]);
          }
          // one barrier per harvest: it also tells every thread whether some push reached the compaction mark (or overflowed)
          if (!consumer_bar_or(need_compact)) break;
          compact();
          if (!s_ovf_latched) break;  // (stable until the next compaction, which is behind further barriers)
        }
        touched = false;
      }
      if (flags & F_END_ITEM) {
        compact();
        const uint32_t n_out = s_cnt;
        const uint32_t b = a.order[item / a.n_chunks], c = item % a.n_chunks;
        if (a.n_chunks == 1) {
          for (uint32_t i = tid; i < a.k; i += CT) {
            const bool ok = i < n_out;
            const uint64_t key = ok ? cand[i] : TRR_KEY_EMPTY;
            if (a.out_keys) a.out_keys[(uint64_t)b * a.k + i] = key;
            if (a.out_ord) a.out_ord[(uint64_t)b * a.k + i] = ok ? trr_key_ord(key) : 0xFFFFFFFFu;
            if (a.out_score) a.out_score[(uint64_t)b * a.k + i] = ok ? trr_key_score(key) : 0.0f;
          }
          if (tid == 0 && a.out_n) a.out_n[b] = n_out;
        } else {
          uint64_t* dst = a.partial + (uint64_t)item * a.k;  // indexed by position in `order` (x n_chunks + chunk)
          for (uint32_t i = tid; i < a.k; i += CT) dst[i] = i < n_out ? cand[i] : TRR_KEY_EMPTY;
        }
        consumer_bar();
        if (tid == 0) { s_cnt = 0; s_thr = TRR_KEY_EMPTY; }
        consumer_bar();
      }
    }
  }
}

// =============================================================================================
// integer fast pass (the default search path): order-free accumulation, then exact re-scoring of the survivors
// =============================================================================================
// The exact kernel above pays for the reference's summation order: every consumer warp owns a sub-range and walks its
// slice of every term segment on its own (a dozen postings per slice), and the f32 cells force plain read-modify-write.
// The fast pass only SELECTS.  Every posting adds ceil(impact * scale) to an integer cell with one native shared-memory
// atomic (ATOMS.ADD: measured 6.3 cycles per 32 random cells against 10.4 for a plain LDS/FADD/STS chain and 16.4 for
// the CAS loop an f32 atomic compiles to; tools/ubench/smem_atom.cu), so the order inside a pass is free and the
// accumulate threads walk the staged postings FLAT - no sub-range ownership, no boundary searches, no per-segment shuffles.
// ceil(impact * scale) is one FFMA with round-up onto 2^23 (the integer appears in the mantissa; F2I runs at a fraction of
// the FMA rate).  Integer sums are exact and never below scale * (real-number sum of the impacts), which bounds the
// reference's f32 score of every document from above:  score_f32(d) <= F(d) / scale * (1 + T * 2^-23)  (T term slots,
// each f32 add rounds by at most 2^-24 relative).  The kernel keeps the kf best documents per query by (F, ordinal);
// bm25_rescore_kernel recomputes their scores in the reference's order and proves that nothing that was dropped can reach
// the k-th exact score.
// Two cell widths (template BITS): 16-bit cells, two per word, halve the per-range scan of the accumulator - the largest
// shared-memory cost of the pass - and leave room for a second accumulator (below); their scale (65535 / sum of the query terms' largest
// impacts) makes the bound a few hundredths wide, far below the score gap between rank k and rank kf on ordinary data.
// A query whose 16-bit proof fails is re-run with 32-bit cells (bound ~1e-5 relative), and only if that proof fails too
// (dozens of documents within rounding distance of the k-th score, e.g. exact duplicates) by the exact kernel.  Both
// fallbacks are device-driven: they read the number of flagged queries from device memory and leave at once when it is 0.
//
// Three warp roles: 4 PRODUCER warps (staging), 8 ACCUMULATE warps (postings -> atomics) and 16 SCAN warps (harvest +
// re-zeroing), with a ring of two accumulators between the last two, so that range i is scanned while range i + 1
// accumulates.  (History of the shape, cfg4 batch on one B200: one group of 16 warps doing both in turn 5.08 ms; the same
// over two skip-table ranges per 128 KB accumulator, i.e. half the barrier intervals, 4.35 ms; two groups 4.39 ms with 8
// loads in flight per thread, 4.26 ms with 4 / 2 - the shared-memory pipe is the limit and deep queues only delay the
// other group; 12 + 12, 16 + 8, 4 + 20 ... warps all within 4.2 - 4.9 ms.  Also tried on this structure: FOLDING 2 or 4
// documents into one cell, which halves the scan per document - a cell still bounds each of its documents, selection and
// proof run on cells - but on Zipf text the sums of two mid-scoring documents crowd out the real top-k cells and 95 % of
// the queries fail the proof.)
// Staging is that of the exact kernel (one cp.async.bulk per term segment into a shared-memory ring, here of two ~45 KB
// stages), with one difference: a flat walk cannot mask the alignment padding of a copy by segment
// bounds.  Padding postings of the SAME term fall outside the document range and are dropped by the range check; the only
// paddings that could fall inside are the last posting of the previous term / the first of the next one, i.e. when an
// odd-aligned segment starts (ends) exactly at its term's first (last) posting.  Those segments are copied without that
// posting, which travels in the pass descriptor instead (FastDesc::single).
#ifdef TRR_TRIAGE  // cycle counters of CTA 0 (tools/gpu_probe.py): where the producer and consumer warp 0 spend their time
#define TRI(...) __VA_ARGS__
#else
#define TRI(...)
#endif

namespace {

constexpr uint32_t NS = TRR_BM25_FAST_STAGES;
constexpr uint32_t NONE32 = 0xFFFFFFFFu;
// A cp.async.bulk costs the issuing warp ~50 cycles whatever its size (tools/ubench/tma_small.cu: 16 copies from one warp
// take 1300 cycles per stage, from four warps 800), and a range needs one copy per query term with postings in it (~16).
// Four producer warps run the same plan in lockstep; each issues the copies of every fourth slot, warp 0 also publishes
// the descriptor.
#ifndef TRR_BM25_FAST_PRODUCER_WARPS
#define TRR_BM25_FAST_PRODUCER_WARPS 4
#endif
constexpr uint32_t PW = TRR_BM25_FAST_PRODUCER_WARPS;
// The consumer side is two groups of warps that work on different document ranges at the same time: ACCUMULATE warps
// walk the staged postings of range i + 1 (LDS + ATOMS) while SCAN warps harvest and re-zero the accumulator of range i
// (LDS.128 + STS.128).  With one group doing both in turn every phase was a few dependent shared-memory round trips with
// the pipe idle in between (49 % busy); two groups keep it fed.  Hand-over through named barriers (arrive / sync pairs).
#ifndef TRR_BM25_FAST_ACC_WARPS
#define TRR_BM25_FAST_ACC_WARPS 8
#endif
#ifndef TRR_BM25_FAST_SCAN_WARPS
#define TRR_BM25_FAST_SCAN_WARPS 16
#endif
constexpr uint32_t XW = TRR_BM25_FAST_ACC_WARPS, XT = XW * 32;    // accumulate warps
constexpr uint32_t YW = TRR_BM25_FAST_SCAN_WARPS, YT = YW * 32;   // scan warps (the first warps of the CTA)
constexpr uint32_t FAST_THREADS = (YW + XW + PW) * 32;
#ifndef TRR_BM25_FAST_FUX
#define TRR_BM25_FAST_FUX 4
#endif
#ifndef TRR_BM25_FAST_FUY
#define TRR_BM25_FAST_FUY 2
#endif
constexpr int FUX = TRR_BM25_FAST_FUX;  // posting loads in flight per accumulate thread
constexpr int FUY = TRR_BM25_FAST_FUY;  // 128-bit cell groups in flight per scan thread
constexpr uint32_t BAR_FULL = 3, BAR_EMPTY = 5;  // named barriers of the hand-over ring (+ slot, two slots at most)
__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 2, %0;" ::"n"(PW * 32) : "memory"); }
__device__ __forceinline__ void fast_bar() { asm volatile("bar.sync 1, %0;" ::"n"(YT) : "memory"); }
__device__ __forceinline__ void ring_arrive(uint32_t id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(XT + YT) : "memory"); }
__device__ __forceinline__ void ring_sync(uint32_t id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(XT + YT) : "memory"); }
struct FastSync { __device__ __forceinline__ void operator()() const { fast_bar(); } };
__device__ __forceinline__ bool fast_bar_or(bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 p, %1, 0;\n\t"
      "barrier.red.or.pred q, 1, %2, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"((uint32_t)pred), "n"(YT)
      : "memory");
  return r != 0;
}

struct FastDesc {
  uint32_t flags;
  uint32_t range_base;  // local id of the first document of the range
  uint32_t item;
  uint32_t total;       // staged postings (alignment padding included)
  uint32_t thr0f;       // the item's bootstrap threshold in fixed point (0 = none)
  float scale;          // 2^e of the item's query
  uint32_t pad0, pad1;
  uint2 single[64];     // [slot] first / [32 + slot] last posting of a term, when the aligned copy had to leave it out
};
struct FastHandoff {    // accumulate warps -> scan warps: one accumulator (or an end-of-item / quit notice)
  uint32_t flags, range_base, item, thr0f;
};

}  // namespace

template <int BITS>  // width of an accumulator cell: 16 (two cells per word) or 32
__global__ void __launch_bounds__(FAST_THREADS, 1)
bm25_fast_kernel(Bm25SearchArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // a pass range = 2^rmul_shift consecutive ranges of the skip table (their postings are contiguous per term): fewer, larger
  // barrier intervals per document
  const uint32_t R = 1u << (a.range_shift + a.rmul_shift);
  const uint32_t W = BITS == 16 ? R >> 1 : R;                                             // accumulator words
  const uint32_t NA = a.n_acc;                                                            // accumulators in the ring (1 or 2)
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                                  // NA x R fixed-point cells
  uint2* stage_buf = reinterpret_cast<uint2*>(acc + (size_t)NA * W);                      // NS x stage_cap
  uint64_t* cand = reinterpret_cast<uint64_t*>(stage_buf + (size_t)NS * a.stage_cap);     // cand_cap
  FastDesc* desc = reinterpret_cast<FastDesc*>(cand + a.cand_cap);                        // NS
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(desc + NS);                            // NS
  uint64_t* empty_bar = full_bar + NS;                                                    // NS
  __shared__ uint32_t s_cnt, s_overflow, s_ovf_latched, s_item;
  __shared__ uint64_t s_thr;
  __shared__ FastHandoff s_hand[2];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (uint32_t s = 0; s < NS; ++s) { trr_mbar_init(&full_bar[s], 1); trr_mbar_init(&empty_bar[s], XW); }
    trr_fence_mbar_init();
    s_cnt = 0; s_overflow = 0; s_thr = TRR_KEY_EMPTY;
  }
  if (warp < YW + XW) for (uint32_t i = tid; i < NA * W; i += XT + YT) acc[i] = 0u;
  __syncthreads();

  if (warp >= YW + XW) {
    // ============================ producer warps ============================
    const uint32_t pw = warp - (YW + XW);
    TRI(long long t_wait = 0; uint32_t t_n = 0; const long long t_begin = clock64();)
    uint32_t stage = 0, phase = 0;
    uint32_t item_thr0f = 0;
    float item_scale = 1.0f;
    const uint2 none2 = make_uint2(NONE32, 0u);
    // publishes one pass: descriptor + bulk copies.  Every lane passes its own slot (al == 0: no copy for the slot).
    auto emit = [&](uint32_t flags, uint32_t range_base, uint32_t item, uint32_t off, uint32_t src_al, uint32_t al,
                    uint32_t total_al, uint2 sf, uint2 sb) {
      TRI(const long long tw = clock64();)
      mbar_wait_or_trap<true>(&empty_bar[stage], phase ^ 1, 1, a.dbg);
      TRI(t_wait += clock64() - tw; ++t_n;)
      if (pw == 0) {
        FastDesc& d = desc[stage];
        d.single[lane] = sf;
        d.single[32 + lane] = sb;
        if (lane == 0) {
          d.flags = flags; d.range_base = range_base; d.item = item; d.total = total_al;
          d.thr0f = item_thr0f; d.scale = item_scale;
        }
        __syncwarp();
        if (lane == 0) {
          if (total_al) trr_mbar_expect_tx(&full_bar[stage], total_al * 8u);
          else trr_mbar_arrive(&full_bar[stage]);
        }
      }
      // (a copy may complete its bytes before warp 0 has announced them: the phase cannot end before warp 0's arrival)
      if (al && (lane & (PW - 1)) == pw)
        trr_bulk_g2s(stage_buf + (size_t)stage * a.stage_cap + off, a.post + src_al, al * 8u, &full_bar[stage]);
      if (++stage == NS) { stage = 0; phase ^= 1; }
    };
    const uint32_t n_items = (a.n_sel_ptr ? *a.n_sel_ptr : a.B) * a.n_chunks;
    while (true) {
      producer_bar();  // every producer warp has read the previous item id
      if (pw == 0 && lane == 0) s_item = atomicAdd(a.queue, 1u);
      producer_bar();
      const uint32_t item = s_item;
      if (item >= n_items) { emit(F_QUIT, 0, item, 0, 0, 0, 0, none2, none2); break; }
      const uint32_t b = a.order[item / a.n_chunks], c = item % a.n_chunks;
      item_thr0f = a.n_chunks == 1 ? a.thr0f[b] : 0u;  // (the bootstrap counts the postings of the whole shard)
      item_scale = a.qscale[b];
      const uint32_t m = a.rmul_shift;
      const uint32_t n_sr = (a.n_ranges + (1u << m) - 1u) >> m;  // pass ranges
      auto sk = [&](uint32_t r) { return min(r << m, a.n_ranges); };  // pass range -> column of the skip row
      const uint32_t r0 = (uint32_t)(((uint64_t)c * n_sr) / a.n_chunks);
      const uint32_t r1 = (uint32_t)(((uint64_t)(c + 1) * n_sr) / a.n_chunks);
      const uint32_t q0 = a.q_off[b];
      const uint32_t T = a.q_off[b + 1] - q0;
      const uint32_t G = (T + 31) >> 5;
      // G == 1: every lane keeps a cursor through the skip row of its term, fetched one range ahead, and knows the term's
      // first / last posting (the ones an aligned copy may have to leave out)
      const uint32_t* row1 = nullptr;
      uint32_t c_s = 0, c_e = 0, c_n = 0, ts1 = 0, te1 = 0;
      uint2 fp1 = none2, lp1 = none2;
      if (G == 1 && lane < T) {
        const uint32_t term = a.q_terms[q0 + lane];
        if (term < a.n_terms) {
          row1 = a.skip + (uint64_t)term * a.skip_ld;
          c_s = row1[sk(r0)];
          c_e = r0 < r1 ? row1[sk(r0 + 1)] : c_s;
          c_n = r0 + 2 <= n_sr ? row1[sk(r0 + 2)] : c_e;
          ts1 = row1[0];
          te1 = row1[a.n_ranges];
          if (te1 > ts1) {
            if (ts1 & 1u) fp1 = a.post[ts1];
            if (te1 & 1u) lp1 = a.post[te1 - 1];
          }
        }
      }
      for (uint32_t r = r0; r < r1; ++r) {
        const uint32_t range_base = r << (a.range_shift + m);
        bool pending_harvest = false;  // a pass of this range was emitted without the harvest flag
        for (uint32_t g = 0; g < G; ++g) {
          uint32_t s = 0, e = 0, ts = 0, te = 0;
          if (G == 1) {
            s = c_s; e = c_e; ts = ts1; te = te1;
            c_s = c_e; c_e = c_n;
            if (row1 && r + 3 <= n_sr) c_n = row1[sk(r + 3)];
          } else {
            const uint32_t ti = g * 32 + lane;
            if (ti < T) {
              const uint32_t term = a.q_terms[q0 + ti];
              if (term < a.n_terms) {
                const uint32_t* row = a.skip + (uint64_t)term * a.skip_ld;
                s = row[sk(r)]; e = row[sk(r + 1)]; ts = row[0]; te = row[a.n_ranges];
              }
            }
          }
          // the first / last posting of the term, when it sits on an odd index, leaves the copy and rides in the descriptor
          uint2 sf = none2, sb = none2;
          if (e > s && (s & 1u) && s == ts) { sf = G == 1 ? fp1 : a.post[s]; ++s; }
          if (e > s && (e & 1u) && e == te) { sb = G == 1 ? lp1 : a.post[e - 1]; --e; }
          bool singles = __any_sync(FULLM, sf.x != NONE32 || sb.x != NONE32);
          uint32_t first = 0;  // slots below `first` are done
          while (true) {
            const uint32_t len = lane >= first ? e - s : 0u;
            const uint32_t al = len ? (((s & 1u) + len + 1u) & ~1u) : 0u;  // postings copied: 16-byte aligned both ends
            uint32_t incl = al;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const uint32_t v = __shfl_up_sync(FULLM, incl, o);
              if ((int)lane >= o) incl += v;
            }
            const uint32_t total = __shfl_sync(FULLM, incl, 31);
            if (total == 0) {
              if (singles) {  // nothing to copy, but some slot's only postings ride in the descriptor
                emit(0u, range_base, item, 0, 0, 0, 0, sf, sb);
                pending_harvest = true;
              }
              break;
            }
            const uint32_t fits = __ballot_sync(FULLM, incl <= a.stage_cap);
            const uint32_t n_fit = fits == FULLM ? 32u : (uint32_t)(__ffs(~fits) - 1);  // slots [first, n_fit) fit
            if (n_fit == first) {
              // slot `first` alone exceeds the stage: stream it in stage-sized pieces that end on even indices, so that no
              // piece carries a padding posting of the same range
              const bool me = lane == first;
              const uint32_t s_al = s & ~1u;
              emit(0u, range_base, item, 0, me ? s_al : 0u, me ? a.stage_cap : 0u, a.stage_cap, singles ? sf : none2,
                   singles ? sb : none2);
              if (me) s = s_al + a.stage_cap;  // < e, because the slot did not fit
              pending_harvest = true;
            } else {
              const bool me = lane >= first && lane < n_fit;
              const uint32_t taken = __shfl_sync(FULLM, incl, n_fit - 1);
              const bool last = (G == 1) && (taken == total);
              emit(last ? F_HARVEST : 0u, range_base, item, me ? incl - al : 0u, me ? (s & ~1u) : 0u, me ? al : 0u, taken,
                   singles ? sf : none2, singles ? sb : none2);
              pending_harvest = !last;
              first = n_fit;
            }
            singles = false;
            sf = none2; sb = none2;
          }
        }
        if (pending_harvest) emit(F_HARVEST, range_base, item, 0, 0, 0, 0, none2, none2);
      }
      emit(F_END_ITEM, 0, item, 0, 0, 0, 0, none2, none2);
    }
    TRI(if (blockIdx.x == 0 && pw == 0 && lane == 0 && a.triage_out && t_n > 64) {
      a.triage_out[16] = (uint32_t)((clock64() - t_begin) >> 4); a.triage_out[17] = (uint32_t)(t_wait >> 4); a.triage_out[18] = t_n; })
  } else if (warp >= YW) {
    // ============================ accumulate warps ============================
    TRI(long long t_full = 0, t_walk = 0, t_empty = 0; uint32_t t_n = 0; const long long t_begin = clock64();)
    const uint32_t xt = tid - YT;
    uint32_t stage = 0, phase = 0, n_msg = 0, slot = 0;
    uint32_t* acc_cur = acc;
    while (true) {
      TRI(long long tc = clock64();)
      mbar_wait_or_trap(&full_bar[stage], phase, 2, a.dbg);
      TRI({ const long long t = clock64(); t_full += t - tc; tc = t; ++t_n; })
      const FastDesc& d = desc[stage];
      const uint32_t flags = d.flags, range_base = d.range_base, item = d.item, total = d.total, thr0f = d.thr0f;
      const float scale = d.scale;
      if (!(flags & F_QUIT)) {
        const uint2* st = stage_buf + (size_t)stage * a.stage_cap;
        // flat walk: posting p of the stage belongs to thread p mod XT; padding and foreign ranges fail the range check
        const uint32_t acc_s = trr_smem_u32(acc_cur);
        auto add1 = [&](const uint2 e) {
          const uint32_t dd = e.x - range_base;
          // ceil(impact * scale) in the low mantissa bits of RU(impact * scale + 2^23)  (impact * scale < 2^22)
          uint32_t q = __float_as_uint(__fmaf_ru(__uint_as_float(e.y), scale, 8388608.0f)) & 0x7FFFFFu;
          uint32_t addr;
          if (BITS == 16) { addr = acc_s + ((dd >> 1) << 2); q <<= (dd & 1u) << 4; }
          else addr = acc_s + (dd << 2);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, %1;\n\t@p red.shared.add.u32 [%2], %3;\n\t}"
                       ::"r"(dd), "r"(R), "r"(addr), "r"(q) : "memory");
        };
        // Every load of a round is issued before the first atomic.  Stage positions past `total` hold stale postings; they
        // are read (in bounds) and replaced by a no-op.
        const uint2 nop = make_uint2(NONE32, 0u);
        const uint2 sg = xt < 64 ? d.single[xt] : nop;
        for (uint32_t p = xt; p < total; p += FUX * XT) {
          uint2 e[FUX];
#pragma unroll
          for (int u = 0; u < FUX; ++u) { const uint32_t pu = p + u * XT; e[u] = st[min(pu, a.stage_cap - 1)]; if (pu >= total) e[u] = nop; }
#pragma unroll
          for (int u = 0; u < FUX; ++u) add1(e[u]);
        }
        add1(sg);
        __syncwarp();
        if (lane == 0) trr_mbar_arrive(&empty_bar[stage]);  // stage and descriptor are free again
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
      TRI({ const long long t = clock64(); t_walk += t - tc; tc = t; })
      if (flags & (F_HARVEST | F_END_ITEM | F_QUIT)) {
        // hand the accumulator (or the notice) to the scan warps and take the next slot of the ring
        if (xt == 0) s_hand[slot] = FastHandoff{flags, range_base, item, thr0f};
        ring_arrive(BAR_FULL + slot);  // (completes when the scan warps sync on it: every add of the range has landed)
        if (flags & F_QUIT) {
          // leave no barrier half-arrived: consume the scan warps' releases of the slots still in flight
          for (uint32_t j = n_msg >= NA - 1 ? n_msg - (NA - 1) : 0u; j < n_msg; ++j) ring_sync(BAR_EMPTY + j % NA);
          break;
        }
        ++n_msg;
        slot = n_msg % NA;
        acc_cur = acc + (size_t)slot * W;
        if (n_msg >= NA) ring_sync(BAR_EMPTY + slot);  // the scan of the range that used this slot before is done
        TRI({ const long long t = clock64(); t_empty += t - tc; tc = t; })
      }
    }
    TRI(if (blockIdx.x == 0 && xt == 0 && a.triage_out && t_n > 64) {
      a.triage_out[0] = (uint32_t)((clock64() - t_begin) >> 4); a.triage_out[1] = (uint32_t)(t_full >> 4);
      a.triage_out[2] = (uint32_t)(t_walk >> 4); a.triage_out[3] = (uint32_t)(t_empty >> 4); a.triage_out[4] = t_n; })
  } else {
    // ============================ scan warps ============================
    const uint32_t compact_at = a.kf + ((a.cand_cap - a.kf) >> 1);
    TRI(uint32_t t_ncomp = 0;)
    // block-wide (consumer warps) compaction: afterwards cand[0..s_cnt) is sorted descending and s_thr is the kf-th best
    auto compact = [&]() {
      fast_bar();
      const uint32_t cnt = min(s_cnt, a.cand_cap);
      for (uint32_t i = cnt + tid; i < a.cand_cap; i += YT) cand[i] = TRR_KEY_EMPTY;
      trr_bitonic_sort_desc(cand, a.cand_cap, tid, YT, FastSync());
      if (tid == 0) {
        const uint32_t c2 = min(cnt, a.kf);
        s_cnt = c2;
        s_ovf_latched = s_overflow;
        s_overflow = 0;
        if (c2 == a.kf) s_thr = cand[a.kf - 1];
      }
      fast_bar();
    };
    // Scan of one accumulator (all of its adds have landed): cells that beat the query's running kf-th best key go to the
    // candidate buffer, everything is re-zeroed.  Returns true when a push reached the compaction mark (or overflowed).
    // (Measured and dropped: a histogram over every thread's largest cell to install a threshold before the first scan of a
    // query, instead of letting ~4000 touched cells go through the 1024-entry candidate buffer in rounds of push / overflow
    // / sort.  A bin boundary of the high byte is too coarse - the bin that holds the kf-th maximum holds a thousand cells -
    // and the extra pass made the kernel 10 % slower at both 1.25M and 10M documents.)
    auto scan = [&](uint32_t* accx, uint32_t range_base, uint32_t thr0f) -> bool {
      bool need_compact = false;
      const uint64_t thr = *reinterpret_cast<volatile uint64_t*>(&s_thr);
      const uint32_t thr_hi = max(max((uint32_t)(thr >> 32), thr0f), 1u);
      uint4* a4 = reinterpret_cast<uint4*>(accx);
      const uint32_t ord0 = a.doc_base + range_base;
      // a 128-bit group with a candidate, processed from the registers of the scan: one atomic reserves the slots of all
      // its qualifying cells (a single shared-memory round trip - the warp that finds a candidate delays the barrier)
      auto harvest4 = [&](uint32_t i, const uint4& v) {
        constexpr int NC = BITS == 16 ? 8 : 4;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t f[NC];
        uint32_t qual = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          f[c] = BITS == 16 ? ((w[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu) : w[c];
          const uint64_t key = ((uint64_t)f[c] << 32) | (uint64_t)(0xFFFFFFFFu - (ord0 + i * NC + c));
          if (f[c] >= thr_hi && key > thr) qual |= 1u << c;
        }
        uint32_t nw[4] = {0u, 0u, 0u, 0u};
        if (qual) {
          const uint32_t n = __popc(qual);
          uint32_t pos = atomicAdd(&s_cnt, n);
          if (pos + n >= compact_at) need_compact = true;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if ((qual >> c) & 1u) {
              if (pos < a.cand_cap) {
                cand[pos] = ((uint64_t)f[c] << 32) | (uint64_t)(0xFFFFFFFFu - (ord0 + i * NC + c));
              } else {  // candidate buffer full: the cell stays in the accumulator and is retried after the compaction
                s_overflow = 1u;
                nw[BITS == 16 ? (c >> 1) : c] |= BITS == 16 ? f[c] << ((c & 1) * 16) : f[c];
              }
              ++pos;
            }
          }
        }
        a4[i] = make_uint4(nw[0], nw[1], nw[2], nw[3]);
      };
      auto pre4 = [&](uint32_t i, const uint4& v) -> bool {
        uint32_t mx;
        if (BITS == 16) {
          const uint32_t m2 = __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w));
          mx = max(m2 & 0xFFFFu, m2 >> 16);
        } else {
          mx = max(max(v.x, v.y), max(v.z, v.w));
        }
        const bool hit = mx >= thr_hi;
        if (mx != 0u && !hit) a4[i] = make_uint4(0u, 0u, 0u, 0u);
        return hit;
      };
      const uint32_t n4 = W >> 2;
      uint32_t i = tid;
      for (; i + (FUY - 1) * YT < n4; i += FUY * YT) {  // FUY 128-bit groups per round trip
        uint4 v[FUY];
#pragma unroll
        for (int u = 0; u < FUY; ++u) v[u] = a4[i + u * YT];
        uint32_t hits = 0;
#pragma unroll
        for (int u = 0; u < FUY; ++u) hits |= pre4(i + u * YT, v[u]) ? 1u << u : 0u;
        if (hits) {
#pragma unroll
          for (int u = 0; u < FUY; ++u)
            if ((hits >> u) & 1u) harvest4(i + u * YT, v[u]);
        }
      }
      for (; i < n4; i += YT) {
        const uint4 v = a4[i];
        if (pre4(i, v)) harvest4(i, v);
      }
      return need_compact;
    };
    // Closes a scan: ONE barrier, which also tells every thread whether some push reached the compaction mark; when the
    // candidate buffer overflowed, the cells that stayed behind are scanned again after the compaction.
    auto settle = [&](uint32_t* accx, uint32_t range_base, uint32_t thr0f, bool need) {
      while (fast_bar_or(need)) {
        TRI(++t_ncomp;)
        compact();
        if (!s_ovf_latched) break;  // (stable until the next compaction, which is behind further barriers)
        need = scan(accx, range_base, thr0f);
      }
    };
    // (Round-2 history: with ONE group of warps doing accumulate and scan in turn, a second accumulator that let scan(i)
    // share a barrier interval with accumulate(i + 1) was slower, 6.0 against 5.6 ms at cfg4 - all warps still moved through
    // the phases together.  Walking two skip-table ranges per 128 KB accumulator (rmul_shift = 1, n_acc = 1) gave 4.35 ms.)
    // (Measured and dropped: closing a scan behind the NEXT hand-over barrier instead of a barrier of its own - 4.50 against
    // 4.39 ms.  The time "in the barrier" is the scan warps' own loads and zeroing stores draining through the shared-memory
    // pipe, which they wait for wherever the next barrier instruction is.)
    uint32_t n = 0;
    TRI(long long t_full = 0, t_scan = 0, t_settle = 0, t_end = 0; const long long t_begin = clock64();)
    while (true) {
      const uint32_t slot = n % NA;
      TRI(long long tc = clock64();)
      ring_sync(BAR_FULL + slot);  // every add of the range has landed (or a notice is posted)
      TRI({ const long long t = clock64(); t_full += t - tc; tc = t; })
      const FastHandoff m = s_hand[slot];
      if (m.flags & F_QUIT) break;
      uint32_t* accx = acc + (size_t)slot * W;
      if (m.flags & F_HARVEST) {
        const bool need = scan(accx, m.range_base, m.thr0f);
        TRI({ const long long t = clock64(); t_scan += t - tc; tc = t; })
        settle(accx, m.range_base, m.thr0f, need);
        TRI({ const long long t = clock64(); t_settle += t - tc; tc = t; })
      }
      if (m.flags & F_END_ITEM) {
        compact();
        const uint32_t n_out = s_cnt;
        uint64_t* dst = a.fast_keys + (uint64_t)m.item * a.kf;  // indexed by position in `order` (x n_chunks + chunk)
        for (uint32_t i = tid; i < a.kf; i += YT) dst[i] = i < n_out ? cand[i] : TRR_KEY_EMPTY;
        fast_bar();
        if (tid == 0) { s_cnt = 0; s_thr = TRR_KEY_EMPTY; }
        fast_bar();
        TRI({ const long long t = clock64(); t_end += t - tc; tc = t; })
      }
      ring_arrive(BAR_EMPTY + slot);  // the slot (zeroed again) goes back to the accumulate warps
      ++n;
    }
    TRI(if (blockIdx.x == 0 && tid == 0 && a.triage_out && n > 64) {
      a.triage_out[8] = (uint32_t)((clock64() - t_begin) >> 4); a.triage_out[9] = (uint32_t)(t_full >> 4);
      a.triage_out[10] = (uint32_t)(t_scan >> 4); a.triage_out[11] = (uint32_t)(t_settle >> 4);
      a.triage_out[12] = (uint32_t)(t_end >> 4); a.triage_out[13] = n; a.triage_out[14] = t_ncomp; })
  }
}

// ---------------------------------------------------------------------------------------------
// exact re-scoring of the fast pass's candidates + proof.  One CTA per query.
//   1. impact of every (candidate, term slot) pair by binary search in the term's postings of the candidate's range
//      (32 slots at a time through shared memory);
//   2. per candidate the reference's sum: slots in query order, f32, sequential (src/index.rs:228-233; a slot whose term
//      does not contain the document adds +0.0, which leaves a non-negative partial sum unchanged);
//   3. canonical sort, top k;
//   4. proof: a document the fast pass dropped has F <= F_kf (the last kept fixed-point score), hence an f32 score
//      <= F_kf * 2^-e * (1 + T * 2^-23); if that is below the k-th exact score nothing dropped can enter or tie the
//      top k.  A list that is not full dropped nothing with a positive score.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr uint32_t RS_THREADS = 256;
constexpr uint32_t RS_CB = 64;   // candidates per block
constexpr uint32_t RS_TB = 32;   // term slots per block
}

__global__ void __launch_bounds__(RS_THREADS)
bm25_rescore_kernel(Bm25RescoreArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);            // cap2
  float* sums = reinterpret_cast<float*>(keys + a.cap2);             // kf
  float* w = sums + a.kf;                                            // RS_CB x (RS_TB + 1)
  __shared__ uint32_t s_n;
  const uint32_t tid = threadIdx.x;
  if (blockIdx.x >= (a.sel_n ? *a.sel_n : a.B)) return;
  const uint32_t b = a.sel[blockIdx.x];
  const uint64_t* fk = a.fast_keys + (uint64_t)blockIdx.x * a.kf;
  const uint32_t q0 = a.q_off[b], T = a.q_off[b + 1] - q0;
  if (tid == 0) s_n = 0;
  __syncthreads();
  {
    uint32_t local = 0;
    for (uint32_t i = tid; i < a.kf; i += RS_THREADS) { sums[i] = 0.0f; local += fk[i] != TRR_KEY_EMPTY; }
    if (local) atomicAdd(&s_n, local);
  }
  __syncthreads();
  const uint32_t n = s_n;  // candidates (the list is sorted: entries [0, n) are the non-empty ones)
  for (uint32_t c0 = 0; c0 < n; c0 += RS_CB) {
    const uint32_t nc = min(RS_CB, n - c0);
    for (uint32_t t0 = 0; t0 < T; t0 += RS_TB) {
      const uint32_t nt = min(RS_TB, T - t0);
      // every thread resolves its (candidate, slot) pairs TOGETHER: the look-ups are chains of dependent loads that miss
      // the L2 (the postings of a 10M-document index do not fit), so what matters is how many chains are in flight
      constexpr uint32_t NP = (RS_CB * RS_TB + RS_THREADS - 1) / RS_THREADS;  // 8 pairs per thread at most
      uint32_t lo[NP], hi[NP], end[NP], dc[NP];
#pragma unroll
      for (uint32_t j = 0; j < NP; ++j) {
        const uint32_t pi = tid + j * RS_THREADS;
        lo[j] = hi[j] = end[j] = 0; dc[j] = 0;
        if (pi < nc * nt) {
          const uint32_t ci = pi / nt, ti = pi - ci * nt;
          const uint32_t term = a.q_terms[q0 + t0 + ti];
          if (term < a.n_terms) {
            const uint32_t doc = trr_key_ord(fk[c0 + ci]) - a.doc_base;
            const uint32_t* row = a.skip + (uint64_t)term * a.skip_ld + (doc >> a.range_shift);
            const uint32_t b0 = row[0], e0 = row[1];
            dc[j] = doc; end[j] = e0; hi[j] = e0;
            // first probe by interpolation: the documents of a term are spread evenly over a range
            uint32_t g = b0;
            if (e0 > b0) {
              g = b0 + (uint32_t)(((uint64_t)(doc & ((1u << a.range_shift) - 1u)) * (e0 - b0)) >> a.range_shift);
              const uint32_t d = a.post[g].x;
              if (d < doc) { lo[j] = g + 1; } else { lo[j] = b0; hi[j] = g; }
              // (the bracket is halved by the binary search below; starting from the interpolated position saves the
              // first steps only when the list is long, which is where the chain is longest)
              const uint32_t span = 64;
              if (d < doc && g + span < e0 && a.post[g + span].x >= doc) hi[j] = g + span;
              else if (d >= doc && g >= b0 + span && a.post[g - span].x < doc) lo[j] = g - span + 1;
            } else {
              lo[j] = b0;
            }
          }
        }
      }
      while (true) {
        bool any = false;
#pragma unroll
        for (uint32_t j = 0; j < NP; ++j) {
          if (lo[j] < hi[j]) {
            const uint32_t mid = (lo[j] + hi[j]) >> 1;
            if (a.post[mid].x < dc[j]) lo[j] = mid + 1; else hi[j] = mid;
            any = true;
          }
        }
        if (!any) break;
      }
#pragma unroll
      for (uint32_t j = 0; j < NP; ++j) {
        const uint32_t pi = tid + j * RS_THREADS;
        if (pi < nc * nt) {
          const uint32_t ci = pi / nt, ti = pi - ci * nt;
          float v = 0.0f;
          if (lo[j] < end[j]) {
            const uint2 e = a.post[lo[j]];
            if (e.x == dc[j]) v = __uint_as_float(e.y);
          }
          w[ci * (RS_TB + 1) + ti] = v;
        }
      }
      __syncthreads();
      if (tid < nc) {
        float s = sums[c0 + tid];
        for (uint32_t ti = 0; ti < nt; ++ti) s = s + w[tid * (RS_TB + 1) + ti];
        sums[c0 + tid] = s;
      }
      __syncthreads();
    }
  }
  for (uint32_t i = tid; i < a.cap2; i += RS_THREADS) {
    uint64_t key = TRR_KEY_EMPTY;
    if (i < n) {
      const float s = sums[i];
      if (s > 0.0f) key = trr_make_key(s, trr_key_ord(fk[i]));  // src/index.rs:236 keeps only score > 0.0
    }
    keys[i] = key;
  }
  trr_bitonic_sort_desc(keys, a.cap2, tid, RS_THREADS, BlockSync());
  uint32_t local = 0;
  for (uint32_t i = tid; i < a.k; i += RS_THREADS) {
    const uint64_t key = i < a.cap2 ? keys[i] : TRR_KEY_EMPTY;
    const bool ok = key != TRR_KEY_EMPTY;
    local += ok;
    a.out_ord[(uint64_t)b * a.k + i] = ok ? trr_key_ord(key) : 0xFFFFFFFFu;
    a.out_score[(uint64_t)b * a.k + i] = ok ? trr_key_score(key) : 0.0f;
  }
  if (tid == 0) s_n = 0;
  __syncthreads();
  if (local) atomicAdd(&s_n, local);
  __syncthreads();
  if (tid == 0) {
    a.out_n[b] = s_n;
    bool proven = true;
    if (n == a.kf) {  // the list is full: documents were dropped
      const uint32_t f_excl = (uint32_t)(fk[a.kf - 1] >> 32) + a.margin_f;
      const double bound = (double)f_excl / (double)a.qscale[b] * (1.0 + (double)T * 1.1920928955078125e-07);
      const uint64_t kth = keys[a.k - 1];
      proven = kth != TRR_KEY_EMPTY && bound < (double)trr_key_score(kth);
    }
    if (!proven) a.flagged[atomicAdd(a.n_flagged, 1u)] = b;
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
cudaError_t trr_launch_bm25_kill(const uint2* post, uint32_t* tf, uint64_t n_postings, const uint32_t* dead_bits,
                                 uint32_t* n_killed, cudaStream_t st) {
  if (n_postings == 0) return cudaSuccess;
  unsigned grid = (unsigned)std::min<uint64_t>((n_postings + 255) / 256, 148u * 32u);
  bm25_kill_kernel<<<grid, 256, 0, st>>>(post, tf, n_postings, dead_bits, n_killed);
  return cudaGetLastError();
}

cudaError_t trr_launch_bm25_check(const uint32_t* skip, uint32_t skip_ld, uint32_t n_terms, uint32_t n_ranges,
                                  const uint64_t* term_off, const uint2* post, uint64_t n_postings, uint32_t n_docs,
                                  uint32_t* bad, cudaStream_t st) {
  const uint64_t work = std::max<uint64_t>(n_terms, n_postings);
  if (work == 0) return cudaSuccess;
  unsigned grid = (unsigned)std::min<uint64_t>((work + 255) / 256, 148u * 32u);
  bm25_check_kernel<<<grid, 256, 0, st>>>(skip, skip_ld, n_terms, n_ranges, term_off, post, n_postings, n_docs, bad);
  return cudaGetLastError();
}

cudaError_t trr_launch_bm25_merge(const Bm25MergeArgs& a, cudaStream_t st) {
  if (a.n_postings_new == 0) return cudaSuccess;
  unsigned grid = (unsigned)std::min<uint64_t>((a.n_postings_new + 255) / 256, 148u * 32u);
  bm25_merge_kernel<<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t trr_launch_bm25_build(const Bm25BuildArgs& a, cudaStream_t st) {
  if (a.n_terms) bm25_skip_empty_kernel<<<(a.n_terms + 255) / 256, 256, 0, st>>>(a);
  if (a.n_postings) {
    unsigned grid = (unsigned)((a.n_postings + 255) / 256);
    if (grid > 148u * 32u) grid = 148u * 32u;
    bm25_build_kernel<<<grid, 256, 0, st>>>(a);
  }
  return cudaGetLastError();
}

size_t trr_bm25_search_smem(uint32_t range_shift, uint32_t stage_cap, uint32_t cand_cap) {
  return ((size_t)4 << range_shift) + (size_t)2 * stage_cap * 8 + (size_t)cand_cap * 8 + 2 * sizeof(PassDesc) + 4 * 8 +
         (size_t)2 * 32 * 17 * 4;
}

cudaError_t trr_launch_bm25_plan(const Bm25SearchArgs& a, uint64_t* plan_keys, cudaStream_t st) {
  uint32_t cap2 = 0;
  if (a.B > 1 && a.B <= 4096) cap2 = trr_pow2_ceil(a.B);  // larger batches keep the submission order
  const uint32_t n = cap2 ? cap2 : a.B;
  bm25_cost_kernel<<<(n + 7) / 8, 256, 0, st>>>(a, plan_keys, cap2);
  if (cap2) bm25_order_kernel<<<1, 1024, (size_t)cap2 * 8, st>>>(a, plan_keys, cap2);
  return cudaGetLastError();
}

cudaError_t trr_launch_bm25_search(const Bm25SearchArgs& a, unsigned grid, cudaStream_t st) {
  if (a.B == 0 || grid == 0) return cudaSuccess;
  const size_t smem = trr_bm25_search_smem(a.range_shift, a.stage_cap, a.cand_cap);
  cudaError_t e = cudaFuncSetAttribute(bm25_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaFuncSetAttribute(bm25_search_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  bm25_search_kernel<<<grid, TRR_BM25_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}


size_t trr_bm25_fast_smem(int bits, uint32_t range_shift, uint32_t n_acc, uint32_t stage_cap, uint32_t cand_cap) {
  return n_acc * ((size_t)(bits == 16 ? 2 : 4) << range_shift) + (size_t)NS * stage_cap * 8 + (size_t)cand_cap * 8 + NS * sizeof(FastDesc) +
         2 * NS * 8;
}

cudaError_t trr_launch_bm25_fast(const Bm25SearchArgs& a, int bits, unsigned grid, cudaStream_t st) {
  if (a.B == 0 || grid == 0) return cudaSuccess;
  if (a.n_acc < 1 || a.n_acc > 2) return cudaErrorInvalidValue;
  const size_t smem = trr_bm25_fast_smem(bits, a.range_shift + a.rmul_shift, a.n_acc, a.stage_cap, a.cand_cap);
  auto kernel = bits == 16 ? bm25_fast_kernel<16> : bm25_fast_kernel<32>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  kernel<<<grid, FAST_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t trr_launch_bm25_rescore(const Bm25RescoreArgs& a, cudaStream_t st) {
  if (a.B == 0) return cudaSuccess;
  const size_t smem = (size_t)a.cap2 * 8 + (size_t)a.kf * 4 + (size_t)RS_CB * (RS_TB + 1) * 4;
  cudaError_t e = cudaFuncSetAttribute(bm25_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  bm25_rescore_kernel<<<a.B, RS_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}
