// fusion.cu — K4: merge of shard-local lists + score fusion + final top-k (compiled with -fmad=false).
//
// Replaces FusionStrategy::fuse and its helpers (reference src/fusion.rs:42-231) and the result assembly of
// HybridRetriever::retrieve (src/retrieve.rs:193-217).  One CTA per query:
//   1. per source (dense, sparse): merge the G shard-local top-C lists into the global top-C in canonical
//      order (G = 1 on a single GPU; the lists then pass through unchanged);
//   2. fuse.  The reference accumulates per ChunkId in a HashMap, dense list first and then the sparse list,
//      each in list order; here one thread owns each distinct id and replays exactly that sequence of f32
//      additions, so the fused scores are bit-identical (including for duplicated ids inside one list);
//   3. sort by (fused score desc, ordinal asc) — Union sorts by rank instead (src/fusion.rs:154-155) — and
//      emit the first k entries with the per-source scores (NaN when absent).
#include <math_constants.h>

#include "common.cuh"
#include "fusion.cuh"

namespace {

constexpr int FT = 128;  // threads per CTA
constexpr float F32_EPSILON = 1.1920929e-07f;

// min-max normalisation (src/fusion.rs:183-202); fminf/fmaxf folds are order-independent
__device__ void min_max_normalize(const float* s, uint32_t n, float* out, float* red) {
  if (n == 0) return;
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
  for (uint32_t i = threadIdx.x; i < n; i += FT) { mn = fminf(mn, s[i]); mx = fmaxf(mx, s[i]); }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = mn; red[4 + (threadIdx.x >> 5)] = mx; }
  __syncthreads();
  mn = fminf(fminf(red[0], red[1]), fminf(red[2], red[3]));
  mx = fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7]));
  const float range = mx - mn;
  if (fabsf(range) < F32_EPSILON) {
    for (uint32_t i = threadIdx.x; i < n; i += FT) out[i] = 1.0f;
  } else {
    for (uint32_t i = threadIdx.x; i < n; i += FT) out[i] = (s[i] - mn) / range;
  }
  __syncthreads();
}

// z-score normalisation (src/fusion.rs:205-224); the two sums are sequential, as in the reference
__device__ void z_score_normalize(const float* s, uint32_t n, float* out, float* red) {
  if (n == 0) return;
  if (threadIdx.x == 0) {
    const float nf = (float)n;
    float sum = 0.0f;
    for (uint32_t i = 0; i < n; ++i) sum = sum + s[i];
    const float mean = sum / nf;
    float vs = 0.0f;
    for (uint32_t i = 0; i < n; ++i) { const float t = s[i] - mean; vs = vs + t * t; }
    const float variance = vs / nf;
    red[0] = mean;
    red[1] = sqrtf(variance);
  }
  __syncthreads();
  const float mean = red[0], std_dev = red[1];
  if (fabsf(std_dev) < F32_EPSILON) {
    for (uint32_t i = threadIdx.x; i < n; i += FT) out[i] = 0.0f;
  } else {
    for (uint32_t i = threadIdx.x; i < n; i += FT) out[i] = (s[i] - mean) / std_dev;
  }
  __syncthreads();
}

}  // namespace

__global__ void __launch_bounds__(FT)
fuse_kernel(FuseArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // layout: keys[mcap] | id[2C] | sc[2C] | nrm[2C] | fkeys[fcap]
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* id = reinterpret_cast<uint32_t*>(keys + a.mcap);
  float* sc = reinterpret_cast<float*>(id + 2 * a.C);
  float* nrm = sc + 2 * a.C;
  uint64_t* fkeys = reinterpret_cast<uint64_t*>(nrm + 2 * a.C);
  __shared__ float red[8];
  __shared__ uint32_t s_n[2], s_nf;
  const uint32_t b = blockIdx.x, tid = threadIdx.x;
  if (b >= a.B) return;

  // ---- 1. per-source merge of the G shard lists ----
  for (uint32_t src = 0; src < 2; ++src) {
    uint32_t* lid = id + src * a.C;
    float* lsc = sc + src * a.C;
    if (a.G == 1) {
      const uint32_t n = a.n[src] ? min(a.n[src][b], a.C) : 0u;
      for (uint32_t i = tid; i < n; i += FT) {
        lid[i] = a.ord[src][(uint64_t)b * a.C + i];
        lsc[i] = a.score[src][(uint64_t)b * a.C + i];
      }
      if (tid == 0) s_n[src] = n;
      __syncthreads();
    } else {
      for (uint32_t i = tid; i < a.mcap; i += FT) keys[i] = TRR_KEY_EMPTY;
      __syncthreads();
      uint32_t local = 0;
      for (uint32_t i = tid; i < a.G * a.C; i += FT) {
        const uint32_t g = i / a.C, j = i % a.C;
        const uint32_t n = a.n[src] ? min(a.n[src][(uint64_t)g * a.shard_stride + b], a.C) : 0u;
        if (j < n) {
          const uint64_t o = (uint64_t)g * a.shard_stride + (uint64_t)b * a.C + j;
          keys[i] = trr_make_key(a.score[src][o], a.ord[src][o]);
          ++local;
        }
      }
      if (tid == 0) s_n[src] = 0;
      __syncthreads();
      if (local) atomicAdd(&s_n[src], local);
      trr_bitonic_sort_desc(keys, a.mcap, tid, (uint32_t)FT, BlockSync());
      const uint32_t n = min(s_n[src], a.C);
      __syncthreads();
      for (uint32_t i = tid; i < n; i += FT) {
        // the shard-local scores are the reference's exact f32 values, and the key keeps them bit for bit
        // (only -0.0 is folded onto +0.0)
        lid[i] = trr_key_ord(keys[i]);
        lsc[i] = trr_key_score(keys[i]);
      }
      if (tid == 0) s_n[src] = n;
      __syncthreads();
    }
  }
  const uint32_t nd = s_n[0], ns = s_n[1];
  const uint32_t* d_id = id;            const float* d_sc = sc;
  const uint32_t* s_id = id + a.C;      const float* s_sc = sc + a.C;
  float* dn = nrm;                      float* sn = nrm + a.C;

  // ---- 2. fuse ----
  if (a.strategy == TRR_FUSE_LINEAR || a.strategy == TRR_FUSE_CONVEX) {
    min_max_normalize(d_sc, nd, dn, red);
    min_max_normalize(s_sc, ns, sn, red);
  } else if (a.strategy == TRR_FUSE_DBSF) {
    z_score_normalize(d_sc, nd, dn, red);
    z_score_normalize(s_sc, ns, sn, red);
  }
  for (uint32_t i = tid; i < a.fcap; i += FT) fkeys[i] = TRR_KEY_EMPTY;
  if (tid == 0) s_nf = 0;
  __syncthreads();

  const uint32_t n_all = nd + ns;
  for (uint32_t e = tid; e < n_all; e += FT) {
    const bool from_dense = e < nd;
    const uint32_t my = from_dense ? d_id[e] : s_id[e - nd];
    // the owner of an id is its first occurrence in (dense list, then sparse list) order
    bool first = true;
    for (uint32_t j = 0; j < e && first; ++j) first = ((j < nd ? d_id[j] : s_id[j - nd]) != my);
    if (!first) continue;
    float fused = 0.0f;
    bool emit = true;
    uint32_t rank_key = 0;
    switch (a.strategy) {
      case TRR_FUSE_RRF: {  // src/fusion.rs:68-84
        const float k = a.param;
        for (uint32_t r = 0; r < nd; ++r) if (d_id[r] == my) fused = fused + 1.0f / (k + (float)r + 1.0f);
        for (uint32_t r = 0; r < ns; ++r) if (s_id[r] == my) fused = fused + 1.0f / (k + (float)r + 1.0f);
        break;
      }
      case TRR_FUSE_LINEAR:
      case TRR_FUSE_CONVEX: {  // src/fusion.rs:87-119
        const float dw = a.param, sw = 1.0f - a.param;
        for (uint32_t r = 0; r < nd; ++r) if (d_id[r] == my) fused = fused + dw * dn[r];
        for (uint32_t r = 0; r < ns; ++r) if (s_id[r] == my) fused = fused + sw * sn[r];
        break;
      }
      case TRR_FUSE_DBSF: {  // src/fusion.rs:122-138
        for (uint32_t r = 0; r < nd; ++r) if (d_id[r] == my) fused = fused + dn[r];
        for (uint32_t r = 0; r < ns; ++r) if (s_id[r] == my) fused = fused + sn[r];
        break;
      }
      case TRR_FUSE_UNION: {  // src/fusion.rs:141-160: dense insert overwrites, sparse or_insert
        bool in_dense = false;
        for (uint32_t r = 0; r < nd; ++r) if (d_id[r] == my) { fused = d_sc[r]; rank_key = r; in_dense = true; }
        if (!in_dense) {
          for (uint32_t r = 0; r < ns; ++r) if (s_id[r] == my) { fused = s_sc[r]; rank_key = nd + r; break; }
        }
        break;
      }
      default: {  // TRR_FUSE_INTERSECTION, src/fusion.rs:163-180: maps built by collect(), last occurrence wins
        bool in_d = false, in_s = false;
        float dv = 0.0f, sv = 0.0f;
        for (uint32_t r = 0; r < nd; ++r) if (d_id[r] == my) { dv = d_sc[r]; in_d = true; }
        for (uint32_t r = 0; r < ns; ++r) if (s_id[r] == my) { sv = s_sc[r]; in_s = true; }
        emit = in_d && in_s;
        fused = (dv + sv) / 2.0f;
        break;
      }
    }
    if (!emit) continue;
    // sort key: Union by rank ascending; otherwise (score desc, ordinal asc)
    fkeys[e] = (a.strategy == TRR_FUSE_UNION)
                   ? ((static_cast<uint64_t>(0xFFFFFFFFu - rank_key) << 32) | e)
                   : trr_make_key(fused, my);
  }
  __syncthreads();
  trr_bitonic_sort_desc(fkeys, a.fcap, tid, (uint32_t)FT, BlockSync());
  // count fused entries
  uint32_t cnt_local = 0;
  for (uint32_t i = tid; i < a.fcap; i += FT) cnt_local += (fkeys[i] != TRR_KEY_EMPTY);
  if (cnt_local) atomicAdd(&s_nf, cnt_local);
  __syncthreads();
  const uint32_t n_out = min(s_nf, a.k);

  // ---- 3. emit ----
  for (uint32_t i = tid; i < a.k; i += FT) {
    uint32_t oid = 0xFFFFFFFFu;
    float fs = 0.0f, ds = CUDART_NAN_F, ss = CUDART_NAN_F;
    if (i < n_out) {
      const uint64_t key = fkeys[i];
      if (a.strategy == TRR_FUSE_UNION) {
        const uint32_t e = (uint32_t)(key & 0xFFFFFFFFu);
        oid = e < nd ? d_id[e] : s_id[e - nd];
        // Union returns the ORIGINAL score of the winning occurrence (bit-exact, -0.0 preserved)
        bool in_dense = false;
        for (uint32_t r = 0; r < nd; ++r) if (d_id[r] == oid) { fs = d_sc[r]; in_dense = true; }
        if (!in_dense) for (uint32_t r = 0; r < ns; ++r) if (s_id[r] == oid) { fs = s_sc[r]; break; }
      } else {
        oid = trr_key_ord(key);
        fs = trr_key_score(key);
      }
      // src/retrieve.rs:197-213 — score maps built by collect(): last occurrence wins
      for (uint32_t r = 0; r < nd; ++r) if (d_id[r] == oid) ds = d_sc[r];
      for (uint32_t r = 0; r < ns; ++r) if (s_id[r] == oid) ss = s_sc[r];
    }
    const uint64_t o = (uint64_t)b * a.k + i;
    a.out_ord[o] = oid;
    a.out_fused[o] = fs;
    if (a.out_dense) a.out_dense[o] = ds;
    if (a.out_sparse) a.out_sparse[o] = ss;
  }
  if (tid == 0) a.out_n[b] = n_out;
}

size_t trr_fuse_smem(const FuseArgs& a) {
  return (size_t)a.mcap * 8 + (size_t)2 * a.C * 4 * 3 + (size_t)a.fcap * 8 + 64;
}

cudaError_t trr_launch_fuse(const FuseArgs& a, cudaStream_t st) {
  if (a.B == 0) return cudaSuccess;
  const size_t smem = trr_fuse_smem(a);
  cudaError_t e = cudaFuncSetAttribute(fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fuse_kernel<<<a.B, FT, smem, st>>>(a);
  return cudaGetLastError();
}
