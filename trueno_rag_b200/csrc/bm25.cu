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
// Search: one CTA per query (dynamic queue).  The CTA walks the document ranges in order; per range it
// stages the postings of all query terms into shared memory, then accumulates them term by term IN QUERY
// ORDER into an R-entry f32 accumulator in shared memory (a document occurs at most once per term, so
// plain read-modify-write without atomics is race-free inside a term; a barrier separates terms, which
// makes the sum order deterministic and equal to the reference's).  Touched documents are then harvested
// with an exchange-with-zero (which also re-zeroes the accumulator), filtered by the query's running k-th
// best key and appended to a candidate buffer that is compacted by a block-wide bitonic sort.
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
    const uint32_t doc = a.post_doc[p];
    // src/index.rs:137-153
    const float tf = (float)a.post_tf[p];
    const float doc_len = (float)a.doc_len[doc];
    const float idf = a.idf[t];
    const float tf_norm = (tf * (a.k1 + 1.0f)) / (tf + a.k1 * (1.0f - a.b + a.b * doc_len / a.avgdl));
    const float impact = idf * tf_norm;
    a.post[p] = make_uint2(doc, __float_as_uint(impact));
    // skip table
    const uint64_t t_begin = a.term_off[t], t_end = a.term_off[t + 1];
    const uint32_t r = doc >> a.range_shift;
    const int64_t r_prev = (p == t_begin) ? -1 : (int64_t)(a.post_doc[p - 1] >> a.range_shift);
    uint32_t* row = a.skip + (uint64_t)t * a.skip_ld;
    for (int64_t rr = r_prev + 1; rr <= (int64_t)r; ++rr) row[rr] = (uint32_t)p;
    if (p + 1 == t_end)
      for (uint32_t rr = r + 1; rr <= a.n_ranges; ++rr) row[rr] = (uint32_t)t_end;
  }
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
// search
// =============================================================================================
template <int NT>
__global__ void __launch_bounds__(NT, 2)
bm25_search_kernel(Bm25SearchArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t R = 1u << a.range_shift;
  float* acc = reinterpret_cast<float*>(smem_raw);                              // R
  uint2* st = reinterpret_cast<uint2*>(acc + R);                                // stage_cap
  uint64_t* cand = reinterpret_cast<uint64_t*>(st + a.stage_cap);               // cand_cap
  uint32_t* seg_s = reinterpret_cast<uint32_t*>(cand + a.cand_cap);             // NT
  uint32_t* seg_l = seg_s + NT;                                                 // NT
  __shared__ uint32_t s_q, s_cnt;
  __shared__ uint64_t s_thr;
  const uint32_t tid = threadIdx.x;

  for (uint32_t i = tid; i < R; i += NT) acc[i] = 0.0f;
  __syncthreads();

  // compaction of the candidate buffer (block-wide); afterwards cand[0..cnt) is sorted descending
  // Every thread carries the candidate count in a register (`n_cand`, uniform across the CTA): it is advanced by
  // __syncthreads_count at the end of each harvest round, so the decision to compact never races with the pushes
  // of the round in flight.
  auto compact = [&](uint32_t cnt) -> uint32_t {
    __syncthreads();
    for (uint32_t i = cnt + tid; i < a.cand_cap; i += NT) cand[i] = TRR_KEY_EMPTY;
    trr_bitonic_sort_desc(cand, a.cand_cap, tid, (uint32_t)NT, BlockSync());
    const uint32_t c2 = min(cnt, a.k);
    if (tid == 0) {
      s_cnt = c2;
      if (c2 == a.k && a.k > 0) s_thr = cand[a.k - 1];
    }
    __syncthreads();
    return c2;
  };

  while (true) {
    __syncthreads();
    if (tid == 0) s_q = atomicAdd(a.counter + 1, 1u);
    __syncthreads();
    const uint32_t bi = s_q;
    if (bi >= *a.n_slow) break;
    const uint32_t b = a.slow_list[bi];
    const uint32_t q0 = a.q_off[b];
    const uint32_t T = a.q_off[b + 1] - q0;   // host guarantees T <= NT
    if (tid == 0) { s_cnt = 0; s_thr = TRR_KEY_EMPTY; }
    uint32_t n_cand = 0;
    // per-thread cursor of "its" query term through the skip table
    uint32_t my_term = 0xFFFFFFFFu;
    const uint32_t* my_skip = nullptr;
    uint32_t cur = 0, nxt = 0;
    if (tid < T) {
      my_term = a.q_terms[q0 + tid];
      if (my_term < a.n_terms) {
        my_skip = a.skip + (uint64_t)my_term * a.skip_ld;
        cur = my_skip[0];
        nxt = my_skip[1];
      }
    }
    __syncthreads();

    for (uint32_t r = 0; r < a.n_ranges; ++r) {
      const uint32_t range_base = r << a.range_shift;
      if (tid < T) {
        seg_s[tid] = cur;
        seg_l[tid] = my_skip ? (nxt - cur) : 0u;
        cur = nxt;
        if (my_skip && r + 2 <= a.n_ranges) nxt = my_skip[r + 2];
      }
      __syncthreads();
      // ---- ordered accumulation, staged in batches of consecutive terms ----
      uint32_t i = 0, last_tot = 0, n_any = 0, n_staged_batches = 0, n_direct = 0;
      while (i < T) {
        const uint32_t len_i = seg_l[i];
        if (len_i > a.stage_cap) {
          // a single term larger than the stage: stream it in pieces (same term: no ordering inside)
          const uint32_t s0 = seg_s[i];
          for (uint32_t pos = 0; pos < len_i; pos += a.stage_cap) {
            const uint32_t take = min(a.stage_cap, len_i - pos);
            for (uint32_t e = tid; e < take; e += NT) {
              const uint2 p = a.post[s0 + pos + e];
              acc[p.x - range_base] = acc[p.x - range_base] + __uint_as_float(p.y);
            }
          }
          __syncthreads();
          n_any += len_i;
          ++n_direct;
          ++i;
          continue;
        }
        uint32_t j = i, tot = 0;
        while (j < T && tot + seg_l[j] <= a.stage_cap) { tot += seg_l[j]; ++j; }
        if (tot == 0) { i = j; continue; }
        // stage the batch
        uint32_t off = 0;
        for (uint32_t x = i; x < j; ++x) {
          const uint32_t len = seg_l[x], s0 = seg_s[x];
          for (uint32_t e = tid; e < len; e += NT) st[off + e] = a.post[s0 + e];
          off += len;
        }
        __syncthreads();
        // accumulate term by term, in query order
        off = 0;
        for (uint32_t x = i; x < j; ++x) {
          const uint32_t len = seg_l[x];
          if (len == 0) continue;
          for (uint32_t e = tid; e < len; e += NT) {
            const uint2 p = st[off + e];
            acc[p.x - range_base] = acc[p.x - range_base] + __uint_as_float(p.y);
          }
          off += len;
          __syncthreads();
        }
        ++n_staged_batches;
        last_tot = tot;
        n_any += tot;
        i = j;
      }
      if (n_any == 0) { __syncthreads(); continue; }
      // ---- harvest touched documents ----
      if (n_staged_batches == 1 && n_direct == 0) {  // every posting of this range is still in the stage
        for (uint32_t e0 = 0; e0 < last_tot; e0 += NT) {
          if (n_cand > a.cand_cap - NT) n_cand = compact(n_cand);
          const uint32_t e = e0 + tid;
          int pushed = 0;
          if (e < last_tot) {
            const uint32_t d = st[e].x;
            const float v = atomicExch(&acc[d - range_base], 0.0f);
            if (v > 0.0f) {  // src/index.rs:236 keeps only score > 0.0
              const uint64_t key = trr_make_key(v, a.doc_base + d);
              if (key > s_thr) { const uint32_t pos = atomicAdd(&s_cnt, 1u); cand[pos] = key; pushed = 1; }
            }
          }
          n_cand += (uint32_t)__syncthreads_count(pushed);
        }
      } else {
        for (uint32_t x = 0; x < T; ++x) {
          const uint32_t len = seg_l[x], s0 = seg_s[x];
          for (uint32_t e0 = 0; e0 < len; e0 += NT) {
            if (n_cand > a.cand_cap - NT) n_cand = compact(n_cand);
            const uint32_t e = e0 + tid;
            int pushed = 0;
            if (e < len) {
              const uint32_t d = a.post[s0 + e].x;
              const float v = atomicExch(&acc[d - range_base], 0.0f);
              if (v > 0.0f) {
                const uint64_t key = trr_make_key(v, a.doc_base + d);
                if (key > s_thr) { const uint32_t pos = atomicAdd(&s_cnt, 1u); cand[pos] = key; pushed = 1; }
              }
            }
            n_cand += (uint32_t)__syncthreads_count(pushed);
          }
        }
      }
    }
    // ---- emit the query's top-k ----
    n_cand = compact(n_cand);
    const uint32_t n_out = min(n_cand, a.k);
    for (uint32_t i = tid; i < a.k; i += NT) {
      const bool ok = i < n_out;
      const uint64_t key = ok ? cand[i] : TRR_KEY_EMPTY;
      if (a.out_keys) a.out_keys[(uint64_t)b * a.k + i] = key;
      if (a.out_ord) a.out_ord[(uint64_t)b * a.k + i] = ok ? trr_key_ord(key) : 0xFFFFFFFFu;
      if (a.out_score) a.out_score[(uint64_t)b * a.k + i] = ok ? trr_key_score(key) : 0.0f;
    }
    if (tid == 0 && a.out_n) a.out_n[b] = n_out;
  }
}

// =============================================================================================
// fast search kernel (queries with at most BM25_FAST_TMAX terms; everything else takes the kernel above)
//
// f32 addition is commutative, so the sum over the query terms of a document that matches ONE or TWO terms does
// not depend on the order of accumulation; only documents matching three or more terms need the reference's
// query-term order.  Per (query, document range):
//   phase 1  every posting of every query term bumps a packed 8-bit match counter of its document (shared-memory
//            atomics); postings stay cached in registers for the later phases;
//   phase 2  postings of documents with <= 2 matches are added with shared-memory float atomics (order-free, exact);
//            postings of documents with >= 3 matches are deferred and their term slots recorded in a bit mask;
//   phase 3  the deferred postings are replayed in query-term order (ascending term slot): normally by ONE warp from
//            a small shared-memory list (__syncwarp between slots); block-wide from the registers if the list overflows;
//   harvest  exchange-with-zero of the touched accumulators, threshold filter, candidate buffer.
// Queries with more than BM25_FAST_TMAX terms are routed to the general kernel by the host.
// =============================================================================================
constexpr int BM25_FAST_TMAX = 128;
constexpr int BM25_EPT = 8;          // postings cached per thread
constexpr int BM25_DEF_CAP = 1024;   // deferred postings replayed by a single warp (more: block-wide replay)

template <int NT>
__global__ void __launch_bounds__(NT, 2)
bm25_search_fast_kernel(Bm25SearchArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t R = 1u << a.range_shift;
  float* acc = reinterpret_cast<float*>(smem_raw);                        // R f32
  uint32_t* cnt32 = reinterpret_cast<uint32_t*>(acc + R);                 // R packed 8-bit counters
  uint64_t* cand = reinterpret_cast<uint64_t*>(cnt32 + (R >> 2));         // cand_cap keys
  uint32_t* seg_s = reinterpret_cast<uint32_t*>(cand + a.cand_cap);       // TMAX
  uint32_t* seg_off = seg_s + BM25_FAST_TMAX;                             // TMAX + 1 (exclusive prefix of lengths)
  uint32_t* def_doc = seg_off + BM25_FAST_TMAX + 4;                       // DEF_CAP: (local doc << 8) | term slot
  float* def_imp = reinterpret_cast<float*>(def_doc + BM25_DEF_CAP);      // DEF_CAP
  __shared__ uint32_t s_q, s_cnt, s_ndef, s_slot_mask[BM25_FAST_TMAX / 32], s_warp_tot[NT / 32];
  __shared__ uint64_t s_thr;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* cnt8 = reinterpret_cast<uint8_t*>(cnt32);

  for (uint32_t i = tid; i < R; i += NT) acc[i] = 0.0f;
  for (uint32_t i = tid; i < (R >> 2); i += NT) cnt32[i] = 0u;
  __syncthreads();

  auto compact = [&]() {
    __syncthreads();
    const uint32_t cnt = min(s_cnt, a.cand_cap);
    for (uint32_t i = cnt + tid; i < a.cand_cap; i += NT) cand[i] = TRR_KEY_EMPTY;
    trr_bitonic_sort_desc(cand, a.cand_cap, tid, (uint32_t)NT, BlockSync());
    if (tid == 0) {
      const uint32_t c2 = min(cnt, a.k);
      s_cnt = c2;
      if (c2 == a.k && a.k > 0) s_thr = cand[a.k - 1];
    }
    __syncthreads();
  };

  while (true) {
    __syncthreads();
    if (tid == 0) s_q = atomicAdd(a.counter, 1u);
    __syncthreads();
    const uint32_t bi = s_q;
    if (bi >= a.n_fast) break;
    const uint32_t b = a.fast_list[bi];
    const uint32_t q0 = a.q_off[b];
    const uint32_t T = a.q_off[b + 1] - q0;  // <= BM25_FAST_TMAX (host-side split)
    if (tid == 0) { s_cnt = 0; s_thr = TRR_KEY_EMPTY; }
    const uint32_t* my_skip = nullptr;
    uint32_t c_lo = 0, c_hi = 0, c_pre = 0;
    if (tid < T) {
      const uint32_t term = a.q_terms[q0 + tid];
      if (term < a.n_terms) {
        my_skip = a.skip + (uint64_t)term * a.skip_ld;
        c_lo = my_skip[0];
        c_hi = my_skip[1];
        c_pre = a.n_ranges >= 2 ? my_skip[2] : c_hi;
      }
    }
    __syncthreads();

    for (uint32_t r = 0; r < a.n_ranges; ++r) {
      const uint32_t range_base = r << a.range_shift;
      // ---- segment table of this range + exclusive prefix of the lengths (block scan over <= 128 values)
      uint32_t my_len = 0;
      if (tid < BM25_FAST_TMAX) {
        if (tid < T && my_skip) {
          seg_s[tid] = c_lo;
          my_len = c_hi - c_lo;
          c_lo = c_hi;
          c_hi = c_pre;
          if (r + 3 <= a.n_ranges) c_pre = my_skip[r + 3];
        }
        uint32_t incl = my_len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
          if ((int)lane >= o) incl += v;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        my_len = incl - my_len;  // exclusive prefix inside the warp
      }
      if (tid < BM25_FAST_TMAX / 32) s_slot_mask[tid] = 0;
      if (tid == 0) s_ndef = 0;
      __syncthreads();
      if (tid < BM25_FAST_TMAX) {
        uint32_t add = 0;
        for (uint32_t w = 0; w < warp; ++w) add += s_warp_tot[w];
        seg_off[tid] = my_len + add;
      }
      uint32_t tot = 0;
#pragma unroll
      for (uint32_t w = 0; w < BM25_FAST_TMAX / 32; ++w) tot += s_warp_tot[w];
      if (tid == 0) seg_off[BM25_FAST_TMAX] = tot;
      __syncthreads();
      if (tot == 0) continue;

      const bool single = tot <= (uint32_t)(NT * BM25_EPT);
      uint32_t e_doc[BM25_EPT];   // local doc id, 0xFFFFFFFF = no posting
      float e_imp[BM25_EPT];
      uint32_t e_slot[BM25_EPT];  // query term slot
      // Flattened posting index e -> (term slot, offset).  A warp's postings are contiguous in e, so the slot is found
      // once per warp by binary search over the 128 prefix sums and then advanced linearly per lane.
      auto load_chunk = [&](uint32_t base) {
        const uint32_t e_first = base + warp * 32u;  // posting of lane 0, j = 0
        uint32_t x = 0;
        if (e_first < tot) {
          uint32_t lo = 0, hi = BM25_FAST_TMAX;
#pragma unroll
          for (int it = 0; it < 7; ++it) {
            const uint32_t mid = (lo + hi) >> 1;
            if (seg_off[mid] <= e_first) lo = mid; else hi = mid;
          }
          x = lo;
        }
#pragma unroll
        for (int j = 0; j < BM25_EPT; ++j) {
          const uint32_t e = base + (uint32_t)j * NT + tid;
          e_doc[j] = 0xFFFFFFFFu;
          if (e < tot) {
            while (seg_off[x + 1] <= e) ++x;  // seg_off[TMAX] == tot > e terminates the walk
            const uint2 p = a.post[seg_s[x] + (e - seg_off[x])];
            e_doc[j] = p.x - range_base;
            e_imp[j] = __uint_as_float(p.y);
            e_slot[j] = x;
          }
        }
      };
      // ---- phase 1: match counters
      for (uint32_t base = 0; base < tot; base += NT * BM25_EPT) {
        load_chunk(base);
#pragma unroll
        for (int j = 0; j < BM25_EPT; ++j)
          if (e_doc[j] != 0xFFFFFFFFu) atomicAdd(&cnt32[e_doc[j] >> 2], 1u << (8 * (e_doc[j] & 3)));
      }
      __syncthreads();
      // ---- phase 2: order-free adds; postings of documents with >= 3 matches are deferred
      uint32_t deferred = 0;  // bit j: cached posting j belongs to a >= 3-match document (single-chunk case)
      for (uint32_t base = 0; base < tot; base += NT * BM25_EPT) {
        if (!single) load_chunk(base);
#pragma unroll
        for (int j = 0; j < BM25_EPT; ++j) {
          if (e_doc[j] == 0xFFFFFFFFu) continue;
          if (cnt8[e_doc[j]] <= 2) {
            atomicAdd(&acc[e_doc[j]], e_imp[j]);
          } else {
            deferred |= 1u << j;
            atomicOr(&s_slot_mask[e_slot[j] >> 5], 1u << (e_slot[j] & 31));
            const uint32_t pos = atomicAdd(&s_ndef, 1u);
            if (pos < (uint32_t)BM25_DEF_CAP) { def_doc[pos] = (e_doc[j] << 8) | e_slot[j]; def_imp[pos] = e_imp[j]; }
          }
        }
      }
      __syncthreads();
      // ---- phase 3: exact ordered replay of the deferred postings, ascending term slot == query order.  Inside a
      // slot every posting has a different document, so plain read-modify-write is race-free.
      const uint32_t n_def = s_ndef;
      if (n_def != 0 && n_def <= (uint32_t)BM25_DEF_CAP) {
        // common case: one warp walks the list once per slot, __syncwarp between slots
        if (warp == 0) {
#pragma unroll 1
          for (uint32_t w = 0; w < BM25_FAST_TMAX / 32; ++w) {
            uint32_t m = s_slot_mask[w];
            while (m) {
              const uint32_t x = w * 32 + (__ffs(m) - 1);
              m &= m - 1;
              for (uint32_t u = lane; u < n_def; u += 32) {
                const uint32_t e = def_doc[u];
                if ((e & 0xFFu) == x) acc[e >> 8] = acc[e >> 8] + def_imp[u];
              }
              __syncwarp();
            }
          }
        }
        __syncthreads();
      } else if (n_def != 0) {
        // heavy case (e.g. a frequent term repeated in the query): block-wide replay from the registers / postings
#pragma unroll 1
        for (uint32_t w = 0; w < BM25_FAST_TMAX / 32; ++w) {
          uint32_t m = s_slot_mask[w];
          while (m) {
            const uint32_t x = w * 32 + (__ffs(m) - 1);
            m &= m - 1;
            if (single) {
              if (deferred) {
#pragma unroll
                for (int j = 0; j < BM25_EPT; ++j)
                  if (((deferred >> j) & 1u) && e_slot[j] == x) acc[e_doc[j]] = acc[e_doc[j]] + e_imp[j];
              }
            } else {
              for (uint32_t base = 0; base < tot; base += NT * BM25_EPT) {
                load_chunk(base);
#pragma unroll
                for (int j = 0; j < BM25_EPT; ++j)
                  if (e_doc[j] != 0xFFFFFFFFu && e_slot[j] == x && cnt8[e_doc[j]] > 2)
                    acc[e_doc[j]] = acc[e_doc[j]] + e_imp[j];
              }
            }
            __syncthreads();
          }
        }
      }
      // ---- harvest (also clears the counters); a full candidate buffer triggers a compaction and a retry
      for (uint32_t base = 0; base < tot; base += NT * BM25_EPT) {
        if (!single) load_chunk(base);
        uint32_t done = 0;
        while (true) {
          int overflow = 0;
#pragma unroll
          for (int j = 0; j < BM25_EPT; ++j) {
            if (e_doc[j] == 0xFFFFFFFFu || (done >> j) & 1u) continue;
            const float v = atomicExch(&acc[e_doc[j]], 0.0f);
            cnt8[e_doc[j]] = 0;
            if (v > 0.0f) {  // src/index.rs:236 keeps only score > 0.0
              const uint64_t key = trr_make_key(v, a.doc_base + range_base + e_doc[j]);
              if (key > s_thr) {
                const uint32_t pos = atomicAdd(&s_cnt, 1u);
                if (pos < a.cand_cap) cand[pos] = key;
                else { acc[e_doc[j]] = v; overflow = 1; continue; }  // put it back, retry after the compaction
              }
            }
            done |= 1u << j;
          }
          if (!__syncthreads_or(overflow)) break;
          compact();
        }
      }
      __syncthreads();
      if (s_cnt > (a.cand_cap >> 1)) compact();  // uniform (read after a barrier): keeps the threshold tight
    }
    // ---- emit the query's top-k
    compact();
    const uint32_t n_out = min(s_cnt, a.k);
    for (uint32_t i = tid; i < a.k; i += NT) {
      const bool ok = i < n_out;
      const uint64_t key = ok ? cand[i] : TRR_KEY_EMPTY;
      if (a.out_keys) a.out_keys[(uint64_t)b * a.k + i] = key;
      if (a.out_ord) a.out_ord[(uint64_t)b * a.k + i] = ok ? trr_key_ord(key) : 0xFFFFFFFFu;
      if (a.out_score) a.out_score[(uint64_t)b * a.k + i] = ok ? trr_key_score(key) : 0.0f;
    }
    if (tid == 0 && a.out_n) a.out_n[b] = n_out;
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
cudaError_t trr_launch_bm25_build(const Bm25BuildArgs& a, cudaStream_t st) {
  if (a.n_terms) bm25_skip_empty_kernel<<<(a.n_terms + 255) / 256, 256, 0, st>>>(a);
  if (a.n_postings) {
    unsigned grid = (unsigned)((a.n_postings + 255) / 256);
    if (grid > 148u * 32u) grid = 148u * 32u;
    bm25_build_kernel<<<grid, 256, 0, st>>>(a);
  }
  return cudaGetLastError();
}

size_t trr_bm25_general_smem(const Bm25SearchArgs& a) {
  return ((size_t)4 << a.range_shift) + (size_t)a.stage_cap * 8 + (size_t)a.cand_cap * 8 + (size_t)TRR_BM25_THREADS * 8;
}
size_t trr_bm25_fast_smem(const Bm25SearchArgs& a) {
  return ((size_t)4 << a.range_shift) + ((size_t)1 << a.range_shift) + (size_t)a.cand_cap * 8 +
         (size_t)(2 * BM25_FAST_TMAX + 8) * 4 + (size_t)BM25_DEF_CAP * 8 + 64;
}

// fast kernel over a.fast_list, then the general kernel over a.slow_list (host-listed long queries + queries the
// fast kernel gave up on); the second launch reads its work count from device memory, so no host sync is needed
cudaError_t trr_launch_bm25_search(const Bm25SearchArgs& a, unsigned grid_fast, unsigned grid_slow, cudaStream_t st) {
  if (a.B == 0) return cudaSuccess;
  if (a.n_fast && grid_fast) {
    const size_t smem = trr_bm25_fast_smem(a);
    auto kern = bm25_search_fast_kernel<TRR_BM25_THREADS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid_fast, TRR_BM25_THREADS, smem, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (grid_slow) {
    const size_t smem = trr_bm25_general_smem(a);
    auto kern = bm25_search_kernel<TRR_BM25_THREADS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid_slow, TRR_BM25_THREADS, smem, st>>>(a);
  }
  return cudaGetLastError();
}
