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
    const uint32_t b = s_q;
    if (b >= a.B) break;
    const uint32_t q0 = a.q_off[b];
    const uint32_t T = a.q_off[b + 1] - q0;   // host guarantees T <= NT
    if (tid == 0) { s_cnt = 0; s_thr = TRR_KEY_EMPTY; }
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
          if (s_cnt > a.cand_cap - NT) compact();
          const uint32_t e = e0 + tid;
          if (e < last_tot) {
            const uint32_t d = st[e].x;
            const float v = atomicExch(&acc[d - range_base], 0.0f);
            if (v > 0.0f) {  // src/index.rs:236 keeps only score > 0.0
              const uint64_t key = trr_make_key(v, a.doc_base + d);
              if (key > s_thr) { const uint32_t pos = atomicAdd(&s_cnt, 1u); cand[pos] = key; }
            }
          }
          __syncthreads();
        }
      } else {
        for (uint32_t x = 0; x < T; ++x) {
          const uint32_t len = seg_l[x], s0 = seg_s[x];
          for (uint32_t e0 = 0; e0 < len; e0 += NT) {
            if (s_cnt > a.cand_cap - NT) compact();
            const uint32_t e = e0 + tid;
            if (e < len) {
              const uint32_t d = a.post[s0 + e].x;
              const float v = atomicExch(&acc[d - range_base], 0.0f);
              if (v > 0.0f) {
                const uint64_t key = trr_make_key(v, a.doc_base + d);
                if (key > s_thr) { const uint32_t pos = atomicAdd(&s_cnt, 1u); cand[pos] = key; }
              }
            }
            __syncthreads();
          }
        }
      }
    }
    // ---- emit the query's top-k ----
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

size_t trr_bm25_search_smem(const Bm25SearchArgs& a) {
  return ((size_t)4 << a.range_shift) + (size_t)a.stage_cap * 8 + (size_t)a.cand_cap * 8 + (size_t)TRR_BM25_THREADS * 8;
}

cudaError_t trr_launch_bm25_search(const Bm25SearchArgs& a, unsigned grid, cudaStream_t st) {
  if (grid == 0 || a.B == 0) return cudaSuccess;
  const size_t smem = trr_bm25_search_smem(a);
  auto kern = bm25_search_kernel<TRR_BM25_THREADS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, TRR_BM25_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}
