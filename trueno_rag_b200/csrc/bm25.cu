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
      if (!(impact > 0.0f)) a.flags[0] = 1u;
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
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.queue[0] = 0; a.queue[1] = 0; }
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
  for (uint32_t i = a.q_off[b] + lane; i < a.q_off[b + 1]; i += 32) {
    const uint32_t t = a.q_terms[i];
    if (t < a.n_terms) {
      const uint32_t* row = a.skip + (uint64_t)t * a.skip_ld;
      const uint32_t cnt = row[a.n_ranges] - row[0];
      cost += cnt;
      if (boot && cnt >= a.k) best = max(best, a.term_min[t]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cost += __shfl_xor_sync(0xFFFFFFFFu, cost, o);
    best = max(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
  }
  if (lane == 0) {
    a.thr0[b] = best ? (((uint64_t)best << 32) - 1ull) : TRR_KEY_EMPTY;
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

__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity, uint32_t site, uint32_t* dbg) {
  if (trr_mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!trr_mbar_try_wait(bar, parity)) {
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

  // perf triage (TRR_BM25_DEBUG=8): where CTA 0 spends its cycles
  const bool timed = (a.debug_mode & 8u) && blockIdx.x == 0 && a.dbg != nullptr;
  long long w_prod = 0, w_full = 0, w_bnd = 0, w_acc = 0, w_harv = 0;
  const long long t_begin = timed ? clock64() : 0;
  if (warp == CW) {
    // ============================ producer warp ============================
    uint32_t stage = 0, phase = 0;
    // publishes one pass: descriptor + bulk copies.  Every lane passes its own slot (len == 0: not in the pass).
    uint64_t item_thr0 = TRR_KEY_EMPTY;
    auto emit = [&](uint32_t flags, uint32_t range_base, uint32_t item, uint32_t off, uint32_t src_al, uint32_t al,
                    uint32_t begin, uint32_t end, uint32_t total_al) {
      const long long tw = timed ? clock64() : 0;
      mbar_wait_or_trap(&empty_bar[stage], phase ^ 1, 1, a.dbg);
      if (timed) w_prod += clock64() - tw;
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
    const uint32_t n_items = a.B * a.n_chunks;
    while (true) {
      uint32_t item = 0;
      if (lane == 0) item = atomicAdd(a.queue, 1u);
      item = __shfl_sync(FULLM, item, 0);
      if (item >= n_items) { emit(F_QUIT, 0, item, 0, 0, 0, 0, 0, 0); break; }
      const uint32_t b = a.order[item / a.n_chunks], c = item % a.n_chunks;
      item_thr0 = a.thr0[b];
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
    if (timed && lane == 0) a.dbg[9] = (uint32_t)(w_prod >> 4);
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
      long long tc = timed ? clock64() : 0;
      mbar_wait_or_trap(&full_bar[stage], phase, 2, a.dbg);
      if (timed) { const long long t = clock64(); w_full += t - tc; tc = t; }
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
        if (timed) { const long long t = clock64(); w_bnd += t - tc; tc = t; }
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
      if (timed) { const long long t = clock64(); w_acc += t - tc; tc = t; }
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
      if (timed) { const long long t = clock64(); w_harv += t - tc; tc = t; }
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
          uint64_t* dst = a.partial + ((uint64_t)b * a.n_chunks + c) * a.k;
          for (uint32_t i = tid; i < a.k; i += CT) dst[i] = i < n_out ? cand[i] : TRR_KEY_EMPTY;
        }
        consumer_bar();
        if (tid == 0) { s_cnt = 0; s_thr = TRR_KEY_EMPTY; }
        consumer_bar();
      }
    }
    if (timed && tid == 0) {
      a.dbg[8] = (uint32_t)((clock64() - t_begin) >> 4);
      a.dbg[10] = (uint32_t)(w_full >> 4); a.dbg[11] = (uint32_t)(w_bnd >> 4);
      a.dbg[12] = (uint32_t)(w_acc >> 4); a.dbg[13] = (uint32_t)(w_harv >> 4);
    }
  }
}

// =============================================================================================
// V2 search kernel (opt-in, TRR_BM25_V2=1): warp-autonomous sub-ranges
// =============================================================================================
// A CTA (8 warps) still owns one (query, chunk) item with its candidate buffer and threshold in shared memory, but the
// warps do not move in lockstep: each pulls 2048-document sub-ranges from a CTA-local counter, owns an 8 KB accumulator,
// reads the posting bounds of all query terms with one lane-parallel look-up (fine skip table for frequent terms, the
// coarse table + a document-range filter for the rest), loads the postings straight from global memory (four term slots
// in flight) and harvests its own cells.  CTA-wide barriers happen only when the candidate buffer has to be compacted
// and at the end of the item.  Sum order per document = query-term order, as in the V1 kernel.
__global__ void bm25_fine_kernel(const uint2* __restrict__ post, const uint64_t* __restrict__ term_off,
                                 const uint32_t* __restrict__ fine_terms, uint32_t n_fine, uint32_t* __restrict__ fine,
                                 uint32_t fine_ld, uint32_t n_sub) {
  const uint32_t f = blockIdx.x;
  if (f >= n_fine) return;
  const uint32_t t = fine_terms[f];
  const uint64_t begin = term_off[t], end = term_off[t + 1];
  uint32_t* row = fine + (uint64_t)f * fine_ld;
  if (begin == end) {
    for (uint32_t j = threadIdx.x; j <= n_sub; j += blockDim.x) row[j] = (uint32_t)begin;
    return;
  }
  for (uint64_t p = begin + threadIdx.x; p < end; p += blockDim.x) {
    const uint32_t cur = post[p].x >> TRR_BM25_SUB_SHIFT;
    const int64_t prev = p == begin ? -1 : (int64_t)(post[p - 1].x >> TRR_BM25_SUB_SHIFT);
    for (int64_t j = prev + 1; j <= (int64_t)cur; ++j) row[j] = (uint32_t)p;
    if (p + 1 == end)
      for (uint32_t j = cur + 1; j <= n_sub; ++j) row[j] = (uint32_t)end;
  }
}

// VARIANT 1 is the kernel measured in round 1.  VARIANT 2 (TRR_BM25_V2=2) applies the next steps listed in DESIGN.md
// section 7 - slot bounds compacted into shared memory instead of two shuffles + ffs per slot, posting loads of the next
// eight slots issued before the read-modify-writes of the current eight, four-fold unrolled branch-light harvest - and
// HAS NOT RUN ON HARDWARE YET (the round's GPU budget was spent); no test selects it.
template <int VARIANT>
__global__ void __launch_bounds__(TRR_BM25_V2_WARPS * 32, 3)
bm25_search_warp_kernel(Bm25SearchArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr uint32_t SUBN = 1u << TRR_BM25_SUB_SHIFT;
  constexpr uint32_t NT = TRR_BM25_V2_WARPS * 32;
  float* acc = reinterpret_cast<float*>(smem_raw);                                   // [warps][2048]
  uint64_t* cand = reinterpret_cast<uint64_t*>(acc + TRR_BM25_V2_WARPS * SUBN);      // cand_cap
  uint32_t* slot_tab = reinterpret_cast<uint32_t*>(cand + a.cand_cap);               // VARIANT 2: [warps][2][32] compacted slot bounds
  __shared__ uint32_t s_cnt, s_next, s_item, s_need;
  __shared__ uint64_t s_thr;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* my = acc + warp * SUBN;
  for (uint32_t i = tid; i < TRR_BM25_V2_WARPS * SUBN; i += NT) acc[i] = 0.0f;
  const uint32_t compact_at = a.k + ((a.cand_cap - a.k) >> 1);
  const uint32_t n_items = a.B * a.n_chunks;
  const uint32_t coarse_shift = a.range_shift - TRR_BM25_SUB_SHIFT;

  // candidate buffer -> sorted descending, s_cnt <= k, s_thr = k-th best; every thread of the CTA takes part
  auto compact = [&]() {
    const uint32_t cnt = min(s_cnt, a.cand_cap);
    for (uint32_t i = cnt + tid; i < a.cand_cap; i += NT) cand[i] = TRR_KEY_EMPTY;
    trr_bitonic_sort_desc(cand, a.cand_cap, tid, NT, BlockSync());
    if (tid == 0) {
      const uint32_t c2 = min(cnt, a.k);
      s_cnt = c2;
      if (c2 == a.k && a.k > 0) s_thr = cand[a.k - 1];
      s_need = 0;
    }
    __syncthreads();
  };

  while (true) {
    __syncthreads();  // the previous item is finished (outputs written, shared state free)
    if (tid == 0) s_item = atomicAdd(a.queue, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const uint32_t b = a.order[item / a.n_chunks], c = item % a.n_chunks;
    const uint32_t sub0 = (uint32_t)(((uint64_t)c * a.n_sub) / a.n_chunks);
    const uint32_t sub1 = (uint32_t)(((uint64_t)(c + 1) * a.n_sub) / a.n_chunks);
    const uint64_t thr0 = a.thr0[b];
    const uint32_t q0 = a.q_off[b];
    const uint32_t T = a.q_off[b + 1] - q0;
    const uint32_t G = (T + 31) >> 5;
    if (tid == 0) { s_next = sub0; s_cnt = 0; s_thr = TRR_KEY_EMPTY; s_need = 0; }
    __syncthreads();
    // T <= 32 (the usual case): every lane resolves its term's skip row once per item; a sub-range then costs one
    // dependent look-up instead of three (term id -> row id -> bounds)
    const uint32_t* rowp = nullptr;
    uint32_t rshift = 0;
    if (G == 1 && lane < T) {
      const uint32_t term = a.q_terms[q0 + lane];
      if (term < a.n_terms) {
        const uint32_t fr = a.fine_row[term];
        if (fr != 0xFFFFFFFFu) { rowp = a.fine + (uint64_t)fr * a.fine_ld; }
        else { rowp = a.skip + (uint64_t)term * a.skip_ld; rshift = coarse_shift; }
      }
    }

    bool redo = false;       // the last harvest left candidates in the accumulator (buffer full): harvest again after the compaction
    uint32_t redo_sub = 0;
    while (true) {  // rounds, separated by compactions
      bool exhausted = false;
      while (true) {
        uint32_t sub;
        bool any = true;
        if (redo) {
          sub = redo_sub;
        } else {
          if (*reinterpret_cast<volatile uint32_t*>(&s_need)) break;
          sub = 0;
          if (lane == 0) sub = atomicAdd(&s_next, 1u);
          sub = __shfl_sync(FULLM, sub, 0);
          if (sub >= sub1) { exhausted = true; break; }
          // ---- accumulate the sub-range, term slots in query order ----
          any = false;
          const uint32_t sub_base = sub << TRR_BM25_SUB_SHIFT;
          for (uint32_t g = 0; g < G; ++g) {
            const uint32_t ti = g * 32 + lane;
            uint32_t s = 0, e = 0;
            if (G == 1) {
              if (rowp) { const uint32_t* r2 = rowp + (sub >> rshift); s = r2[0]; e = r2[1]; }
            } else if (ti < T) {
              const uint32_t term = a.q_terms[q0 + ti];
              if (term < a.n_terms) {
                const uint32_t fr = a.fine_row[term];
                if (fr != 0xFFFFFFFFu) {
                  const uint32_t* row = a.fine + (uint64_t)fr * a.fine_ld;
                  s = row[sub];
                  e = row[sub + 1];
                } else {  // infrequent term: its postings of the whole coarse range, filtered by document below
                  const uint32_t* row = a.skip + (uint64_t)term * a.skip_ld + (sub >> coarse_shift);
                  s = row[0];
                  e = row[1];
                }
              }
            }
            if constexpr (VARIANT == 1) {
              uint32_t m = __ballot_sync(FULLM, e > s);
              while (m) {
                uint32_t lo_[4], n_[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const bool valid = m != 0u;
                  const uint32_t l = valid ? __ffs(m) - 1 : 0u;
                  m &= m - 1;
                  lo_[u] = __shfl_sync(FULLM, s, l);
                  const uint32_t hi = __shfl_sync(FULLM, e, l);
                  n_[u] = valid ? hi - lo_[u] : 0u;
                }
                uint2 pe[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) pe[u] = lane < n_[u] ? a.post[lo_[u] + lane] : make_uint2(0xFFFFFFFFu, 0u);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (n_[u] == 0) continue;  // warp-uniform
                  uint32_t d = pe[u].x - sub_base;
                  if (lane < n_[u] && d < SUBN) { my[d] = my[d] + __uint_as_float(pe[u].y); any = true; }
                  for (uint32_t p = 32 + lane; p < n_[u]; p += 32) {  // a segment longer than one load (order inside a term is free)
                    const uint2 x = a.post[lo_[u] + p];
                    d = x.x - sub_base;
                    if (d < SUBN) { my[d] = my[d] + __uint_as_float(x.y); any = true; }
                  }
                  __syncwarp();
                }
              }
            } else {
              // compacted (first posting, count) of the non-empty slots, in slot order
              uint32_t* w_lo = slot_tab + warp * 64;
              uint32_t* w_n = w_lo + 32;
              const uint32_t ne = __ballot_sync(FULLM, e > s);
              const uint32_t n_ne = (uint32_t)__popc(ne);
              if (e > s) {
                const uint32_t r = (uint32_t)__popc(ne & ((1u << lane) - 1u));
                w_lo[r] = s;
                w_n[r] = e - s;
              }
              __syncwarp();
              uint2 nxt[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                nxt[u] = make_uint2(0xFFFFFFFFu, 0u);
                if ((uint32_t)u < n_ne && lane < w_n[u]) nxt[u] = a.post[w_lo[u] + lane];
              }
              for (uint32_t base = 0; base < n_ne; base += 8) {
                uint2 cur[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
#pragma unroll
                for (int u = 0; u < 8; ++u) {  // the next eight slots' postings are in flight while these eight are added
                  const uint32_t slot = base + 8 + (uint32_t)u;
                  nxt[u] = make_uint2(0xFFFFFFFFu, 0u);
                  if (slot < n_ne && lane < w_n[slot]) nxt[u] = a.post[w_lo[slot] + lane];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const uint32_t slot = base + (uint32_t)u;
                  if (slot < n_ne) {  // warp-uniform
                    const uint32_t n = w_n[slot];
                    uint32_t d = cur[u].x - sub_base;
                    if (lane < n && d < SUBN) { my[d] = my[d] + __uint_as_float(cur[u].y); any = true; }
                    if (n > 32) {
                      const uint32_t lo = w_lo[slot];
                      for (uint32_t p = 32 + lane; p < n; p += 32) {
                        const uint2 x = a.post[lo + p];
                        d = x.x - sub_base;
                        if (d < SUBN) { my[d] = my[d] + __uint_as_float(x.y); any = true; }
                      }
                    }
                    __syncwarp();
                  }
                }
              }
              __syncwarp();  // the table is rewritten by the next group of terms
            }
          }
          any = __any_sync(FULLM, any);
        }
        // ---- harvest the warp's 2048 cells ----
        redo = false;
        if (any) {
          const uint64_t thr = max(*reinterpret_cast<volatile uint64_t*>(&s_thr), thr0);
          const float thr_f = thr == TRR_KEY_EMPTY ? -CUDART_INF_F : trr_key_score(thr);
          const uint32_t ord0 = a.doc_base + (sub << TRR_BM25_SUB_SHIFT);
          uint4* a4 = reinterpret_cast<uint4*>(my);
          bool kept = false;
          auto cells4 = [&](uint32_t i) {  // per-element pass over one 128-bit group that holds a candidate
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
              const float f = my[i * 4 + j];
              bool keep = false;
              if (f >= thr_f && f > 0.0f) {  // src/index.rs:236 keeps only score > 0.0
                const uint64_t key = trr_make_key(f, ord0 + i * 4 + j);
                if (key > thr) {
                  const uint32_t pos = atomicAdd(&s_cnt, 1u);
                  if (pos < a.cand_cap) cand[pos] = key;
                  else keep = true;  // stays in the accumulator; harvested again after the compaction
                }
              }
              if (keep) kept = true; else my[i * 4 + j] = 0.0f;
            }
          };
          if constexpr (VARIANT == 1) {
            for (uint32_t i = lane; i < SUBN / 4; i += 32) {
              const uint4 v = a4[i];
              const bool nz = (v.x | v.y | v.z | v.w) != 0u;
              const float mx = fmaxf(fmaxf(__uint_as_float(v.x), __uint_as_float(v.y)),
                                     fmaxf(__uint_as_float(v.z), __uint_as_float(v.w)));
              const bool hit = nz && mx >= thr_f;  // NaN compares false
              if (nz && !hit) a4[i] = make_uint4(0u, 0u, 0u, 0u);
              if (hit) cells4(i);
            }
          } else {
            auto pre4 = [&](uint32_t i, const uint4& v) -> bool {
              const bool nz = (v.x | v.y | v.z | v.w) != 0u;
              const float mx = fmaxf(fmaxf(__uint_as_float(v.x), __uint_as_float(v.y)),
                                     fmaxf(__uint_as_float(v.z), __uint_as_float(v.w)));
              const bool hit = nz && mx >= thr_f;
              if (nz && !hit) a4[i] = make_uint4(0u, 0u, 0u, 0u);
              return hit;
            };
            for (uint32_t i = lane; i < SUBN / 4; i += 128) {  // SUBN / 4 is a multiple of 128
              const uint4 v0 = a4[i], v1 = a4[i + 32], v2 = a4[i + 64], v3 = a4[i + 96];
              const bool h0 = pre4(i, v0), h1 = pre4(i + 32, v1), h2 = pre4(i + 64, v2), h3 = pre4(i + 96, v3);
              if (__any_sync(FULLM, h0 | h1 | h2 | h3)) {
                if (h0) cells4(i);
                if (h1) cells4(i + 32);
                if (h2) cells4(i + 64);
                if (h3) cells4(i + 96);
              }
            }
          }
          __syncwarp();
          redo = __any_sync(FULLM, kept);
        }
        if (redo) {
          redo_sub = sub;
          if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&s_need) = 1u;
          break;
        }
        if (lane == 0 && *reinterpret_cast<volatile uint32_t*>(&s_cnt) >= compact_at)
          *reinterpret_cast<volatile uint32_t*>(&s_need) = 1u;
      }
      // ---- end of a round: every warp is here (out of sub-ranges, or asked to stop for a compaction) ----
      const int all_done = __syncthreads_and(exhausted && !redo);
      compact();
      if (all_done) break;
    }
    // ---- outputs of the item ----
    const uint32_t n_out = s_cnt;
    if (a.n_chunks == 1) {
      for (uint32_t i = tid; i < a.k; i += NT) {
        const bool ok = i < n_out;
        const uint64_t key = ok ? cand[i] : TRR_KEY_EMPTY;
        if (a.out_keys) a.out_keys[(uint64_t)b * a.k + i] = key;
        if (a.out_ord) a.out_ord[(uint64_t)b * a.k + i] = ok ? trr_key_ord(key) : 0xFFFFFFFFu;
        if (a.out_score) a.out_score[(uint64_t)b * a.k + i] = ok ? trr_key_score(key) : 0.0f;
      }
      if (tid == 0 && a.out_n) a.out_n[b] = n_out;
    } else {
      uint64_t* dst = a.partial + ((uint64_t)b * a.n_chunks + c) * a.k;
      for (uint32_t i = tid; i < a.k; i += NT) dst[i] = i < n_out ? cand[i] : TRR_KEY_EMPTY;
    }
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

size_t trr_bm25_search_warp_smem(uint32_t cand_cap) {
  return (size_t)TRR_BM25_V2_WARPS * (4u << TRR_BM25_SUB_SHIFT) + (size_t)cand_cap * 8 + (size_t)TRR_BM25_V2_WARPS * 64 * 4;
}

cudaError_t trr_launch_bm25_search_warp(const Bm25SearchArgs& a, unsigned grid, int variant, cudaStream_t st) {
  const size_t smem = trr_bm25_search_warp_smem(a.cand_cap);
  auto kernel = variant == 2 ? bm25_search_warp_kernel<2> : bm25_search_warp_kernel<1>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  kernel<<<grid, TRR_BM25_V2_WARPS * 32, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t trr_launch_bm25_fine(const uint2* post, const uint64_t* term_off, const uint32_t* fine_terms, uint32_t n_fine,
                                 uint32_t* fine, uint32_t fine_ld, uint32_t n_sub, cudaStream_t st) {
  if (n_fine == 0) return cudaSuccess;
  bm25_fine_kernel<<<n_fine, 256, 0, st>>>(post, term_off, fine_terms, n_fine, fine, fine_ld, n_sub);
  return cudaGetLastError();
}
