// dense_gemm.cu — K2: batched query x corpus scoring on the 5th-generation tensor cores with a fused
// per-row top-k selection, so that the B x N score matrix never reaches HBM.
//
// Replaces, for batches of queries, the per-vector scoring loop of VectorStore::search (reference
// src/index.rs:394-405).  This is the FAST pass only: it selects candidates by
//     fast(q, d) = dot_bf16(q, d) * scale[d] + bias[d]
// (cosine: scale = 1/|d|; dot: scale = 1) with fp32 accumulation in TMEM.  Reported scores and the final
// order come from the exact rescoring + candidate proof in dense_scan.cu (rescore_select_kernel), which
// reproduces the reference arithmetic bit for bit.
//
// Kernel shape (sm_100a; dense_gemm_topk_kernel = cta_group::1, dense_gemm_topk_pair_kernel = cta_group::2, the default
// for two or more query blocks - see the comment above the 2-CTA kernel):
//   grid  = n_slices x n_qblocks persistent CTAs of 320 threads (<= one per SM); a CTA owns 128 queries (rows of A) and
//           a contiguous range of 256-document tiles (rows of B), so its per-query candidate lists live in shared
//           memory for the whole kernel.
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D tiles (128B swizzle) of Q [128 x 64] and docs [256 x 64] (2-CTA:
//            each CTA of the pair loads half of the document tile) into a 3-stage (2-CTA: 4-stage) ring, completion on
//            mbarriers; the scale/bias of a tile's documents arrive by a 2 KB bulk copy.
//   warp 1   TMEM allocator + MMA issuer: 4 x tcgen05.mma (M128 / M256, N256, K16, bf16 -> f32) per k-block, commits
//            release the smem stage and, after the last k-block, publish the accumulator stage.
//   warps 2-9 epilogue: the two warps that share a TMEM lane quarter split the 256 columns of a tile; tcgen05.ld of the
//            accumulator (lane == query row), fused scale/bias from shared memory, one 3-input max tree + compare per
//            32-column chunk against the row's threshold, lane-private replace-min insertion into the (row, half) list
//            (8 / 16 / 32 entries, chosen by the host) with a compare-tree rescan in the rare chunks that have a hit.
//            Two accumulator stages (2 x 256 TMEM columns) overlap epilogue(t) with MMA(t+1).
//   CTAs working on different document slices of the same queries share their thresholds through a global
//   atomicMax table.  Measured (10M x 768 bf16, B = 1024, inside the bench step): 11.9-12.2 ms = 1293-1317 TFLOP/s,
//   93-95 % of the measured sustained bf16 peak; the part runs this kernel against its power cap.
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"
#include "dense.cuh"

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr uint32_t BM = TRR_GEMM_TILE_M;  // 128 queries
constexpr uint32_t BN = TRR_GEMM_TILE_N;  // 256 documents
constexpr uint32_t BK = 64;               // bf16 elements per k-block = one 128-byte swizzle atom
constexpr uint32_t UMMA_K = 16;
#ifndef TRR_GEMM_STAGES
#define TRR_GEMM_STAGES 3
#endif
constexpr uint32_t STAGES = TRR_GEMM_STAGES;
constexpr uint32_t CP = TRR_GEMM_CP;
constexpr uint32_t A_BYTES = BM * BK * 2;  // 16 KB
constexpr uint32_t B_BYTES = BN * BK * 2;  // 32 KB
// Both kernels run 8 epilogue warps per CTA; the two warps that share a TMEM lane quarter split the 256 columns of a tile
// into halves, and every (row, half) owns a candidate list of up to L1MAX entries: [half][row][L1STRIDE] (odd stride:
// lanes of a warp are consecutive rows of one half, conflict-free).
constexpr uint32_t L1MAX = TRR_GEMM_CPS_MAX;
constexpr uint32_t L1STRIDE = L1MAX + 1;
constexpr uint32_t LIST1_BYTES = 2 * BM * L1STRIDE * 4;
constexpr uint32_t SB_BYTES = BN * 8;  // scale/bias of one document tile
constexpr uint32_t SMEM_A = 0;
constexpr uint32_t SMEM_B = SMEM_A + STAGES * A_BYTES;
constexpr uint32_t SMEM_LS = SMEM_B + STAGES * B_BYTES;
constexpr uint32_t SMEM_LO = SMEM_LS + LIST1_BYTES;
constexpr uint32_t SMEM_SB = SMEM_LO + LIST1_BYTES;
constexpr uint32_t SMEM_BAR = SMEM_SB + 2 * SB_BYTES;
constexpr uint32_t SMEM_TOTAL = SMEM_BAR + 256;
static_assert(SMEM_TOTAL <= 227 * 1024, "1-CTA GEMM shared memory");
constexpr uint32_t GEMM1_THREADS = 320;  // TMA producer, MMA issuer, 8 epilogue warps
constexpr uint32_t GEMM2_THREADS = 384;  // pair kernel: + 2 helper warps (merged thresholds)

// instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), both K-major
// (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);

__device__ uint32_t* g_gemm_dbg = nullptr;

// bounded mbarrier wait: a protocol bug must surface as an error, never as a hung GPU.  The site of the timeout
// (and the CTA) is written to a host-mapped word before trapping so that the host can report it.
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity, int site, uint32_t* dbg = nullptr) {
  if (trr_mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!trr_mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) {
      if (dbg) { *dbg = 0x80000000u | ((uint32_t)site << 16) | (blockIdx.x & 0xFFFFu); __threadfence_system(); }
      __trap();
    }
  }
}

// perf triage (debug_mode & 8): the wait with the cycles it took added to `acc`
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, int site, uint32_t* dbg, bool timed,
                                                long long& acc) {
  if (!timed) { mbar_wait_bounded(bar, parity, site, dbg); return; }
  const long long t = clock64();
  mbar_wait_bounded(bar, parity, site, dbg);
  acc += clock64() - t;
}

__device__ __forceinline__ void tma_load_2d(const void* map, void* smem_dst, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(trr_smem_u32(smem_dst)), "l"(map), "r"(trr_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor: start address >> 4 in bits 0-13, leading byte
// offset unused for a single swizzle atom along K, stride byte offset (8 rows x 128 B = 1024) >> 4 in bits
// 32-45, descriptor version 1 at bit 46, layout type SWIZZLE_128B (2) in bits 61-63.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(trr_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
// waits for the outstanding tcgen05.ld; the "+r" operands pin every consumer of v[] behind the wait
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
        "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
        "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
        "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
      :
      : "memory");
}

// per-row running state of the epilogue (lives in the registers of the lane that owns the row)
struct RowState {
  float thr;       // max(list_min, threshold shared by the other slices)
  float list_min;  // smallest fast score in the row's candidate list
  uint32_t minpos; // its position
};

// Lane-private replace-min insertion: every epilogue lane owns one query row and its candidate list of `cps` entries
// (16 / 32 / 64, chosen by the host from the number of document slices).  The new value overwrites the list minimum and
// the minimum is found again by a scan of the lane's own list; lanes of a warp insert independently (rows differ), so no
// warp-level cooperation or synchronisation is needed.  Kept out of line: rare in steady state.
// (value, position) minimum of N list entries by a compare tree of depth log2(N): the scan is on the critical path of
// every insertion (an epilogue warp is alone on its scheduler), so a sequential compare chain would cost ~10 cycles per entry
template <uint32_t N>
__device__ __forceinline__ void list_min_tree(const float* my_ls, uint32_t off, float& mn, uint32_t& pos) {
  float mv[N];
  uint32_t mp[N];
#pragma unroll
  for (uint32_t i = 0; i < N; ++i) { mv[i] = my_ls[off + i]; mp[i] = off + i; }
#pragma unroll
  for (uint32_t w = N / 2; w >= 1; w >>= 1) {
#pragma unroll
    for (uint32_t i = 0; i < w; ++i) {
      const bool hi = mv[i + w] < mv[i];
      mv[i] = hi ? mv[i + w] : mv[i];
      mp[i] = hi ? mp[i + w] : mp[i];
    }
  }
  mn = mv[0];
  pos = mp[0];
}

__device__ __noinline__ RowState insert_private(RowState st, float s, uint32_t doc, float* my_ls, uint32_t* my_lo,
                                                uint32_t cps) {
  if (!(s > st.thr)) return st;
  my_ls[st.minpos] = s;
  my_lo[st.minpos] = doc;
  float mn;
  uint32_t pos;
  if (cps == 8) {
    list_min_tree<8>(my_ls, 0, mn, pos);
  } else {
    list_min_tree<16>(my_ls, 0, mn, pos);
  }
  for (uint32_t off = 16; off < cps; off += 16) {
    float m2;
    uint32_t p2;
    list_min_tree<16>(my_ls, off, m2, p2);
    if (m2 < mn) { mn = m2; pos = p2; }
  }
  st.list_min = mn;
  st.minpos = pos;
  if (mn > st.thr) st.thr = mn;
  return st;
}


// One 32-column chunk of a finished accumulator stage, for one epilogue warp (lane == query row): read the 32 f32 sums
// of the lane's row from TMEM, turn them into fast scores with the per-document scale/bias staged in shared memory (sb16:
// 16 float4 = 32 x {scale, bias}, the same address in every lane: broadcast), and insert the ones above the row's
// threshold into the lane's candidate list.  PREFILTER: one 3-input max tree and a single compare per chunk instead of
// a compare + mask bit per element; the per-element tests run only in the (rare) chunks where some lane has a hit.
// `mode` (perf triage only): 2 no insertions, 3 TMEM reads only, 4 no TMEM reads.
template <bool PREFILTER>
__device__ __forceinline__ RowState epilogue_chunk(uint32_t taddr, const float4* sb16, uint32_t dbase, RowState st,
                                                   float* my_ls, uint32_t* my_lo, uint32_t cps, uint32_t mode,
                                                   float* dump_row) {
  uint32_t v[32];
  if (mode != 4) {
    tmem_ld32_issue(taddr, v);
    tmem_ld32_wait(v);
  } else {
#pragma unroll
    for (uint32_t j = 0; j < 32; ++j) v[j] = dbase + j;
  }
  if (mode == 3) {
    uint32_t x = 0;
#pragma unroll
    for (uint32_t j = 0; j < 32; ++j) x ^= v[j];
    if (x == 0x7FC12345u) st.thr = 0.0f;
    return st;
  }
  float sv[32];
  if (PREFILTER) {
#pragma unroll
    for (uint32_t j = 0; j < 32; j += 2) {
      const float4 sb = sb16[j >> 1];
      sv[j] = fmaf(__uint_as_float(v[j]), sb.x, sb.y);
      sv[j + 1] = fmaf(__uint_as_float(v[j + 1]), sb.z, sb.w);
    }
    float m[4];
#pragma unroll
    for (uint32_t g = 0; g < 4; ++g) {  // four independent chains of 3-input maxima (NaN never wins)
      m[g] = fmaxf(fmaxf(sv[8 * g], sv[8 * g + 1]), sv[8 * g + 2]);
      m[g] = fmaxf(fmaxf(m[g], sv[8 * g + 3]), sv[8 * g + 4]);
      m[g] = fmaxf(fmaxf(m[g], sv[8 * g + 5]), sv[8 * g + 6]);
      m[g] = fmaxf(m[g], sv[8 * g + 7]);
    }
    const float mx = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
    if (dump_row) {
#pragma unroll
      for (uint32_t j = 0; j < 32; ++j) dump_row[j] = sv[j];
    }
    const bool hit = mode != 2 && mx > st.thr;
    if (__any_sync(FULL, hit)) {
      if (hit) {
#pragma unroll
        for (uint32_t j = 0; j < 32; ++j)
          if (sv[j] > st.thr) st = insert_private(st, sv[j], dbase + j, my_ls, my_lo, cps);
      }
    }
  } else {
    uint32_t pmask = 0;
#pragma unroll
    for (uint32_t j = 0; j < 32; j += 2) {
      const float4 sb = sb16[j >> 1];
      sv[j] = fmaf(__uint_as_float(v[j]), sb.x, sb.y);
      sv[j + 1] = fmaf(__uint_as_float(v[j + 1]), sb.z, sb.w);
      pmask |= (sv[j] > st.thr ? 1u : 0u) << j;
      pmask |= (sv[j + 1] > st.thr ? 1u : 0u) << (j + 1);
    }
    if (dump_row) {
#pragma unroll
      for (uint32_t j = 0; j < 32; ++j) dump_row[j] = sv[j];
    }
    if (mode == 2) pmask = 0;
    // rare path: some lane has a value above its row's threshold (re-tested inside: the threshold rises as we insert)
    if (__any_sync(FULL, pmask != 0)) {
#pragma unroll
      for (uint32_t j = 0; j < 32; ++j)
        if (pmask & (1u << j)) st = insert_private(st, sv[j], dbase + j, my_ls, my_lo, cps);
    }
  }
  return st;
}

}  // namespace

__global__ void __launch_bounds__(GEMM1_THREADS, 1)
dense_gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
                       GemmTopkArgs a, float* __restrict__ dump, uint32_t dump_ld) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* sbfull_bar = tempty_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sbfull_bar + 2);

  const uint32_t qb = blockIdx.x % a.n_qblocks;
  const uint32_t slice = blockIdx.x / a.n_qblocks;
  const uint32_t t0 = (uint32_t)(((uint64_t)slice * a.n_tiles) / a.n_slices);
  const uint32_t t1 = (uint32_t)(((uint64_t)(slice + 1) * a.n_tiles) / a.n_slices);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_d) : "memory");
    for (uint32_t s = 0; s < STAGES; ++s) { trr_mbar_init(&full_bar[s], 1); trr_mbar_init(&empty_bar[s], 1); }
    for (uint32_t s = 0; s < 2; ++s) {
      trr_mbar_init(&tfull_bar[s], 1);
      trr_mbar_init(&tempty_bar[s], 8);  // one arrival per epilogue warp
      trr_mbar_init(&sbfull_bar[s], 1);
    }
    trr_fence_mbar_init();
  }
  if (warp == 1) {  // TMEM: 512 columns = 2 accumulator stages of 256 f32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(trr_smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // perf triage (debug_mode & 8): SM cycles and nanoseconds spent by CTA 0, i.e. the SM clock this kernel ran at
  long long dbg_c0 = 0;
  uint64_t dbg_t0 = 0;
  const bool timed = (a.debug_mode & 8) && blockIdx.x < 2 && a.dbg != nullptr;
  long long w_empty = 0, w_full = 0, w_tempty = 0;  // cycles the producer / the MMA warp spent waiting
  if ((a.debug_mode & 8) && blockIdx.x == 0 && threadIdx.x == 0) {
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t t = t0; t < t1; ++t) {
        for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait_timed(&empty_bar[stage], phase ^ 1, 1, a.dbg, timed, w_empty);
          trr_mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
          tma_load_2d(&map_q, smem + SMEM_A + stage * A_BYTES, &full_bar[stage], (int32_t)(kb * BK), (int32_t)(qb * BM));
          tma_load_2d(&map_d, smem + SMEM_B + stage * B_BYTES, &full_bar[stage], (int32_t)(kb * BK), (int32_t)(t * BN));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (timed && blockIdx.x == 0) a.dbg[12] = (uint32_t)w_empty;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t stage = 0, phase = 0;
    for (uint32_t t = t0, it = 0; t < t1; ++t, ++it) {
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      mbar_wait_timed(&tempty_bar[as], aphase ^ 1, 2, a.dbg, timed, w_tempty);
      tc_fence_after();
      // the epilogue of tile it-2 has let go of this stage's scale/bias buffer as well: refill it for tile t
      if (lane == 0) {
        trr_mbar_expect_tx(&sbfull_bar[as], SB_BYTES);
        trr_bulk_g2s(smem + SMEM_SB + as * SB_BYTES, a.scale_bias + (uint64_t)t * BN, SB_BYTES, &sbfull_bar[as]);
      }
      for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
        mbar_wait_timed(&full_bar[stage], phase, 3, a.dbg, timed, w_full);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = make_smem_desc(trr_smem_u32(smem + SMEM_A + stage * A_BYTES));
          const uint64_t db = make_smem_desc(trr_smem_u32(smem + SMEM_B + stage * B_BYTES));
#pragma unroll
          for (uint32_t k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(tmem_base + as * BN, da + 2 * k, db + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                        // frees the smem stage once the MMAs retire
          if (kb == a.k_blocks - 1) umma_commit(&tfull_bar[as]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (timed && blockIdx.x == 0 && lane == 0) { a.dbg[10] = (uint32_t)w_full; a.dbg[11] = (uint32_t)w_tempty; }
  } else {
    // ===================== epilogue (8 warps) =====================
    const uint32_t quarter = warp & 3;            // TMEM lane group this warp may access
    const uint32_t half = (warp - 2) >> 2;        // which 128 columns of every tile this warp reads
    const uint32_t row = quarter * 32 + lane;     // query row inside the block
    float* my_ls = reinterpret_cast<float*>(smem + SMEM_LS) + (half * BM + row) * L1STRIDE;
    uint32_t* my_lo = reinterpret_cast<uint32_t*>(smem + SMEM_LO) + (half * BM + row) * L1STRIDE;
    const uint32_t cps = a.cps;
    for (uint32_t j = 0; j < cps; ++j) { my_ls[j] = -CUDART_INF_F; my_lo[j] = 0xFFFFFFFFu; }
    RowState st;
    st.list_min = -CUDART_INF_F;      // smallest score in this row's list
    st.minpos = 0;
    st.thr = -CUDART_INF_F;           // max(list_min, shared threshold)
    uint32_t* gthr = a.gthr + (qb * BM + row);
    constexpr uint32_t CHUNKS = BN / 64;  // 32-column chunks per warp and tile

    for (uint32_t t = t0, it = 0; t < t1; ++t, ++it) {
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      if (a.share_thresholds) {
        const uint32_t g = *reinterpret_cast<volatile uint32_t*>(gthr);
        if (g > trr_f32_orderable(st.thr)) st.thr = trr_orderable_f32(g);
      }
      mbar_wait_bounded(&sbfull_bar[as], aphase, 5, a.dbg);
      mbar_wait_bounded(&tfull_bar[as], aphase, 4, a.dbg);
      tc_fence_after();
      const uint32_t doc0 = t * BN + half * (BN / 2);
      const float4* sbs = reinterpret_cast<const float4*>(smem + SMEM_SB + as * SB_BYTES) + half * (BN / 4);
      const uint32_t mode = a.debug_mode & 7;
      const bool prefilter = !(a.debug_mode & 16);
#pragma unroll 1
      for (uint32_t c = 0; c < (mode == 1 ? 0u : CHUNKS); ++c) {
        const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + as * BN + half * (BN / 2) + c * 32;
        float* dump_row = dump ? dump + (uint64_t)(qb * BM + row) * dump_ld + doc0 + c * 32 : nullptr;
        if (prefilter)
          st = epilogue_chunk<true>(taddr, sbs + c * 16, a.base_ord + doc0 + c * 32, st, my_ls, my_lo, cps, mode, dump_row);
        else
          st = epilogue_chunk<false>(taddr, sbs + c * 16, a.base_ord + doc0 + c * 32, st, my_ls, my_lo, cps, mode, dump_row);
      }
      // accumulator stage and scale/bias buffer drained: hand them back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) trr_mbar_arrive(&tempty_bar[as]);
      if (a.share_thresholds && st.list_min > -CUDART_INF_F) atomicMax(gthr, trr_f32_orderable(st.list_min));
    }
    // publish this half-slice's candidates: virtual slice 2 * slice + half
    const uint64_t base = (((uint64_t)(slice * 2 + half) * a.n_qblocks + qb) * BM + row) * cps;
    for (uint32_t j = 0; j < cps; ++j) { a.cand_score[base + j] = my_ls[j]; a.cand_ord[base + j] = my_lo[j]; }
  }

  tc_fence_before();
  __syncthreads();
  if ((a.debug_mode & 8) && blockIdx.x == 0 && threadIdx.x == 0 && a.dbg) {
    uint64_t t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    a.dbg[8] = (uint32_t)(clock64() - dbg_c0);
    a.dbg[9] = (uint32_t)(t1 - dbg_t0);
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =============================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs (two SMs of one TPC) computes a 256-query x 256-document
// tile.  Each CTA owns 128 query rows (its own A operand and its own 128 x 256 f32 accumulator in its TMEM) and
// loads HALF of the document tile (128 rows of B); the tensor cores of both SMs read both halves.  Compared with the
// 1-CTA kernel this cuts the L2 -> shared-memory operand traffic per CTA from 48 KB to 32 KB per k-block and leaves
// room for a deeper ring in bytes per SM.  Protocol (barriers at identical shared-memory offsets in both CTAs):
//   full[s]    lives in the leader CTA: 1 arrival (the leader's producer) + 64 KB of TMA transaction bytes from both CTAs
//              (a remote arrive per k-block from the peer's producer cost ~900 cycles each and halved the throughput);
//   empty[s]   per CTA, released by the leader's tcgen05.commit multicast to both CTAs;
//   tfull[a]   per CTA, same multicast commit after the last k-block of a tile;
//   tempty[a]  lives in the leader: 16 arrivals (one per epilogue warp of both CTAs);
//   sbfull[i] / sbempty[i]  per CTA: ring of three scale/bias buffers between the producer and the epilogue warps.
// =============================================================================================
namespace {
constexpr uint32_t STAGES2 = 4;
constexpr uint32_t B2_BYTES = (BN / 2) * BK * 2;  // 16 KB: this CTA's half of the document tile
constexpr uint32_t STAGE2_BYTES = A_BYTES + B2_BYTES;
constexpr uint32_t SB_RING = 3;                   // scale/bias buffers (filled by the producer a tile ahead)
constexpr uint32_t SMEM2_A = 0;
constexpr uint32_t SMEM2_B = SMEM2_A + STAGES2 * A_BYTES;
constexpr uint32_t SMEM2_LS = SMEM2_B + STAGES2 * B2_BYTES;
constexpr uint32_t SMEM2_LO = SMEM2_LS + LIST1_BYTES;
constexpr uint32_t SMEM2_SB = SMEM2_LO + LIST1_BYTES;
constexpr uint32_t SMEM2_BAR = SMEM2_SB + SB_RING * SB_BYTES;
constexpr uint32_t SMEM2_TOTAL = SMEM2_BAR + 256;
static_assert(SMEM2_TOTAL <= 227 * 1024, "2-CTA GEMM shared memory");
constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((256u >> 4) << 24);  // M = 256

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const void* map, void* smem_dst, uint32_t bar_cluster_addr, int32_t c0,
                                                 int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(trr_smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(0x1000000000000000ull)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(trr_smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
dense_gemm_topk_pair_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
                            GemmTopkArgs a, float* __restrict__ dump, uint32_t dump_ld) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM2_BAR);
  uint64_t* empty_bar = full_bar + STAGES2;
  uint64_t* tfull_bar = empty_bar + STAGES2;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* sbfull_bar = tempty_bar + 2;
  uint64_t* sbempty_bar = sbfull_bar + SB_RING;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sbempty_bar + SB_RING);
  uint32_t* helper_flags = reinterpret_cast<uint32_t*>(smem + SMEM2_BAR + 192);  // [0] epilogue warps whose lists are initialised, [1] ... that are done

  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const uint32_t qb = blockIdx.x % a.n_qblocks;     // n_qblocks is even; the pair owns query blocks (qb & ~1, qb | 1)
  const uint32_t slice = blockIdx.x / a.n_qblocks;
  const uint32_t t0 = (uint32_t)(((uint64_t)slice * a.n_tiles) / a.n_slices);
  const uint32_t t1 = (uint32_t)(((uint64_t)(slice + 1) * a.n_tiles) / a.n_slices);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_d) : "memory");
    for (uint32_t s = 0; s < STAGES2; ++s) { trr_mbar_init(&full_bar[s], 1); trr_mbar_init(&empty_bar[s], 1); }
    for (uint32_t s = 0; s < 2; ++s) { trr_mbar_init(&tfull_bar[s], 1); trr_mbar_init(&tempty_bar[s], 16); }  // one arrival per epilogue warp of both CTAs
    for (uint32_t s = 0; s < SB_RING; ++s) { trr_mbar_init(&sbfull_bar[s], 1); trr_mbar_init(&sbempty_bar[s], 8); }
    helper_flags[0] = 0; helper_flags[1] = 0;
    trr_fence_mbar_init();
  }
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / TMA / commit
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(trr_smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const bool timed = (a.debug_mode & 8) && blockIdx.x < 2 && a.dbg != nullptr;
  long long w_empty = 0, w_full = 0, w_tempty = 0;
  long long dbg_c0 = 0;
  uint64_t dbg_t0 = 0;
  if (timed && blockIdx.x == 0 && threadIdx.x == 0) {
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  }

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t t = t0, it = 0; t < t1; ++t, ++it) {
        {
          // scale/bias of this tile's 256 documents for this CTA's epilogue; the buffer was last read for tile it - 3,
          // whose epilogue finished before the MMAs of tile it - 1 could start, so this wait does not stall the ring
          const uint32_t sbuf = it % SB_RING, sphase = (it / SB_RING) & 1;
          mbar_wait_bounded(&sbempty_bar[sbuf], sphase ^ 1, 15, a.dbg);
          trr_mbar_expect_tx(&sbfull_bar[sbuf], SB_BYTES);
          trr_bulk_g2s(smem + SMEM2_SB + sbuf * SB_BYTES, a.scale_bias + (uint64_t)t * BN, SB_BYTES, &sbfull_bar[sbuf]);
        }
        for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait_timed(&empty_bar[stage], phase ^ 1, 11, a.dbg, timed, w_empty);
          const uint32_t leader_full = mapa_u32(trr_smem_u32(&full_bar[stage]), 0);
          // the leader alone arrives, expecting the bytes of both CTAs: the peer's copies may complete on the barrier before
          // the leader's expect_tx (the transaction count goes negative; the phase cannot complete without the arrival)
          if (rank == 0) trr_mbar_expect_tx(&full_bar[stage], 2 * STAGE2_BYTES);
          tma_load_2d_pair(&map_q, smem + SMEM2_A + stage * A_BYTES, leader_full, (int32_t)(kb * BK), (int32_t)(qb * BM));
          tma_load_2d_pair(&map_d, smem + SMEM2_B + stage * B2_BYTES, leader_full, (int32_t)(kb * BK),
                           (int32_t)(t * BN + rank * (BN / 2)));
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
      if (timed) a.dbg[12 + blockIdx.x] = (uint32_t)w_empty;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t t = t0, it = 0; t < t1; ++t, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait_timed(&tempty_bar[as], aphase ^ 1, 12, a.dbg, timed, w_tempty);
        tc_fence_after();
        for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait_timed(&full_bar[stage], phase, 13, a.dbg, timed, w_full);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t da = make_smem_desc(trr_smem_u32(smem + SMEM2_A + stage * A_BYTES));
            const uint64_t db = make_smem_desc(trr_smem_u32(smem + SMEM2_B + stage * B2_BYTES));
#pragma unroll
            for (uint32_t k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_pair(tmem_base + as * BN, da + 2 * k, db + 2 * k, IDESC2, (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
            if (kb == a.k_blocks - 1) umma_commit_pair(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
      if (timed && blockIdx.x == 0 && lane == 0) { a.dbg[10] = (uint32_t)w_full; a.dbg[11] = (uint32_t)w_tempty; }
    }
  } else if (warp < 10) {
    // ===================== epilogue (both CTAs, 8 warps each; as in the 1-CTA kernel) =====================
    const uint32_t quarter = warp & 3;
    const uint32_t half = (warp - 2) >> 2;
    const uint32_t row = quarter * 32 + lane;
    float* my_ls = reinterpret_cast<float*>(smem + SMEM2_LS) + (half * BM + row) * L1STRIDE;
    uint32_t* my_lo = reinterpret_cast<uint32_t*>(smem + SMEM2_LO) + (half * BM + row) * L1STRIDE;
    const uint32_t cps = a.cps;
    for (uint32_t j = 0; j < cps; ++j) { my_ls[j] = -CUDART_INF_F; my_lo[j] = 0xFFFFFFFFu; }
    __syncwarp();
    if (lane == 0) atomicAdd(&helper_flags[0], 1u);  // (the helper warps read the lists once all of them are initialised)
    RowState st;
    st.list_min = -CUDART_INF_F;
    st.minpos = 0;
    st.thr = -CUDART_INF_F;
    uint32_t* gthr = a.gthr + (qb * BM + row);
    const uint32_t leader_tempty[2] = {mapa_u32(trr_smem_u32(&tempty_bar[0]), 0), mapa_u32(trr_smem_u32(&tempty_bar[1]), 0)};
    constexpr uint32_t CHUNKS = BN / 64;
    const uint32_t mode = a.debug_mode & 7;
    const bool prefilter = !(a.debug_mode & 16);

    for (uint32_t t = t0, it = 0; t < t1; ++t, ++it) {
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const uint32_t sbuf = it % SB_RING, sphase = (it / SB_RING) & 1;
      if (a.share_thresholds) {
        const uint32_t g = *reinterpret_cast<volatile uint32_t*>(gthr);
        if (g > trr_f32_orderable(st.thr)) st.thr = trr_orderable_f32(g);
      }
      mbar_wait_bounded(&sbfull_bar[sbuf], sphase, 16, a.dbg);
      mbar_wait_bounded(&tfull_bar[as], aphase, 14, a.dbg);
      tc_fence_after();
      const uint32_t doc0 = t * BN + half * (BN / 2);
      const float4* sbs = reinterpret_cast<const float4*>(smem + SMEM2_SB + sbuf * SB_BYTES) + half * (BN / 4);
#pragma unroll 1
      for (uint32_t c = 0; c < (mode == 1 ? 0u : CHUNKS); ++c) {
        const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + as * BN + half * (BN / 2) + c * 32;
        float* dump_row = dump ? dump + (uint64_t)(qb * BM + row) * dump_ld + doc0 + c * 32 : nullptr;
        if (prefilter)
          st = epilogue_chunk<true>(taddr, sbs + c * 16, a.base_ord + doc0 + c * 32, st, my_ls, my_lo, cps, mode, dump_row);
        else
          st = epilogue_chunk<false>(taddr, sbs + c * 16, a.base_ord + doc0 + c * 32, st, my_ls, my_lo, cps, mode, dump_row);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(leader_tempty[as]);  // accumulator stage free (the leader's MMA warp waits for all 16 warps)
        trr_mbar_arrive(&sbempty_bar[sbuf]);     // scale/bias buffer free (this CTA's producer)
      }
      if (a.share_thresholds && st.list_min > -CUDART_INF_F) atomicMax(gthr, trr_f32_orderable(st.list_min));
    }
    // publish this half-slice's candidates: virtual slice 2 * slice + half
    const uint64_t base = (((uint64_t)(slice * 2 + half) * a.n_qblocks + qb) * BM + row) * cps;
    for (uint32_t j = 0; j < cps; ++j) { a.cand_score[base + j] = my_ls[j]; a.cand_ord[base + j] = my_lo[j]; }
    __syncwarp();
    if (lane == 0) atomicAdd(&helper_flags[1], 1u);
  } else if (a.pub != nullptr && a.pub_rank != 0) {
    // ===================== helper warps: merged thresholds =====================
    // A row's own list minimum is the cps-th best of what ITS slice has seen - with 2 * n_slices lists per query that is
    // about rank cps * 2 * n_slices of the documents seen so far, while the re-scoring needs rank CP.  Half of the 32-column
    // chunks of an epilogue warp then still have a lane above its threshold and take the insertion path.  These two warps
    // keep publishing every list's j-th best score (j = pub_rank) and fold, per query, the pub_pick-th SMALLEST of the
    // published values into the shared threshold: 2 * n_slices - pub_pick + 1 lists have j documents at or above it, and
    // the host picks j and pub_pick so that this is at least CP documents.  List entries only ever grow and
    // 32-bit shared-memory reads are single-copy atomic, so a racing read yields a value the list reached at some time:
    // the j-th best of such a snapshot never exceeds the true one.  Exactness does not depend on it: whatever threshold was
    // applied ends up in gthr, which the re-scoring kernel's proof treats as the bound of everything that was dropped.
    const uint32_t h = threadIdx.x - 320;  // 0..63: rows h and h + 64
    volatile uint32_t* hf = helper_flags;
    while (hf[0] < 8u) __nanosleep(200);
    const volatile float* ls = reinterpret_cast<const volatile float*>(smem + SMEM2_LS);
    const uint32_t B_pad = a.n_qblocks * BM, nv = 2 * a.n_slices, cps = a.cps;
    uint32_t last[2] = {0u, 0u};
    // a sweep costs these warps ~3 us of issue slots on two of the four schedulers, and thresholds improve like 1 / (documents
    // seen): the pause between sweeps doubles from 1 us to 64 us (polled in 1 us naps, so the kernel's end is not delayed)
    uint32_t naps = 1;
    while (hf[1] < 8u) {
#pragma unroll
      for (uint32_t rr = 0; rr < 2; ++rr) {
        const uint32_t row = h + rr * 64;
#pragma unroll
        for (uint32_t half = 0; half < 2; ++half) {
          const volatile float* l = ls + (half * BM + row) * L1STRIDE;
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = -CUDART_INF_F;
          for (uint32_t j = 0; j < cps; ++j) {
            float v = l[j];
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float hi = fmaxf(t[i], v); v = fminf(t[i], v); t[i] = hi; }
          }
          float val = t[0];
#pragma unroll
          for (int i = 1; i < 8; ++i) val = (uint32_t)i < a.pub_rank ? t[i] : val;
          const uint32_t enc = val > -CUDART_INF_F ? trr_f32_orderable(val) : 0u;
          __stcg(a.pub + (uint64_t)(slice * 2 + half) * B_pad + qb * BM + row, enc);
        }
      }
#pragma unroll
      for (uint32_t rr = 0; rr < 2; ++rr) {
        const uint32_t q = qb * BM + h + rr * 64;
        uint32_t sm[8];  // the 8 smallest published values, ascending
#pragma unroll
        for (int i = 0; i < 8; ++i) sm[i] = 0xFFFFFFFFu;
        for (uint32_t v = 0; v < nv; ++v) {
          uint32_t x = __ldcg(a.pub + (uint64_t)v * B_pad + q);
#pragma unroll
          for (int i = 0; i < 8; ++i) { const uint32_t lo = min(sm[i], x); x = max(sm[i], x); sm[i] = lo; }
        }
        uint32_t m = sm[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) m = (uint32_t)i < a.pub_pick ? sm[i] : m;
        if (sm[0] != 0u && m != 0xFFFFFFFFu && m > last[rr]) { atomicMax(a.gthr + q, m); last[rr] = m; }
      }
      for (uint32_t i = 0; i < naps && hf[1] < 8u; ++i) __nanosleep(1000);
      naps = min(naps * 2u, 64u);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal this CTA's barriers
  if (timed && blockIdx.x == 0 && threadIdx.x == 0) {
    uint64_t t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    a.dbg[8] = (uint32_t)(clock64() - dbg_c0);
    a.dbg[9] = (uint32_t)(t1 - dbg_t0);
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t f32_to_bf16_rne(float f) {
  uint32_t u = __float_as_uint(f);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// one CTA per (padded) query row: bf16 operand row, ||q - bf16(q)|| for the proof, and ||q|| in the reference's order
// (sequential f32 sum of squares, src/index.rs:442; explicit _rn intrinsics because this file allows FMA contraction)
__global__ void query_prep_kernel(const float* __restrict__ q, uint32_t dim, uint32_t dim_pad, uint32_t B,
                                  uint16_t* __restrict__ q_bf16, float* __restrict__ q_delta, float* __restrict__ q_norm) {
  extern __shared__ float qsh[];
  const uint32_t b = blockIdx.x;
  float d2 = 0.0f;
  for (uint32_t j = threadIdx.x; j < dim_pad; j += blockDim.x) {
    uint16_t h = 0;
    if (b < B && j < dim) {
      const float x = q[(uint64_t)b * dim + j];
      qsh[j] = x;
      h = f32_to_bf16_rne(x);
      const float r = x - __uint_as_float(((uint32_t)h) << 16);
      d2 += r * r;
    }
    q_bf16[(uint64_t)b * dim_pad + j] = h;
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(FULL, d2, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d2;
  __syncthreads();
  if (threadIdx.x == 0 && b < B) {
    float s = 0.0f;
    for (uint32_t w = 0; w < (blockDim.x + 31) / 32; ++w) s += red[w];
    q_delta[b] = sqrtf(s) * 1.0001f;  // upper bound on || q - bf16(q) ||
  }
  if (threadIdx.x == 32 && b < B && q_norm) {
    float s = 0.0f;
    for (uint32_t j = 0; j < dim; ++j) s = __fadd_rn(s, __fmul_rn(qsh[j], qsh[j]));
    q_norm[b] = __fsqrt_rn(s);
  }
}

__global__ void shadow_kernel(const void* __restrict__ rows, int is_bf16, uint32_t dim, uint32_t dim_pad,
                              uint64_t row0, uint64_t n, uint16_t* __restrict__ shadow) {
  const uint64_t total = n * dim_pad;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = row0 + i / dim_pad;
    const uint32_t j = (uint32_t)(i % dim_pad);
    uint16_t h = 0;
    if (j < dim) {
      h = is_bf16 ? reinterpret_cast<const uint16_t*>(rows)[r * dim + j]
                  : f32_to_bf16_rne(reinterpret_cast<const float*>(rows)[r * dim + j]);
    }
    shadow[r * dim_pad + j] = h;
  }
}

void trr_launch_query_prep(const float* q, uint32_t dim, uint32_t dim_pad, uint32_t B, uint32_t B_pad, uint16_t* q_bf16,
                           float* q_delta, float* q_norm, cudaStream_t st) {
  if (B_pad == 0) return;
  query_prep_kernel<<<B_pad, 256, (size_t)dim * 4, st>>>(q, dim, dim_pad, B, q_bf16, q_delta, q_norm);
}

void trr_launch_shadow(const void* rows, int is_bf16, uint32_t dim, uint32_t dim_pad, uint64_t row0, uint64_t n,
                       uint16_t* shadow, cudaStream_t st) {
  if (n == 0) return;
  shadow_kernel<<<1184, 256, 0, st>>>(rows, is_bf16, dim, dim_pad, row0, n, shadow);
}

// ---------------------------------------------------------------------------------------------
// TMA descriptors (driver entry point fetched through the runtime: no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int trr_make_tensor_map_ex(void* out_map128, const void* base, uint64_t rows, uint64_t cols, uint32_t elem_bytes,
                           uint32_t box_cols, uint32_t box_rows) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p)
      return trr_fail(TRR_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
  CUtensorMap* m = reinterpret_cast<CUtensorMap*>(out_map128);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * elem_bytes};  // bytes between rows
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return trr_fail(TRR_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return TRR_OK;
}

int trr_make_tensor_map(void* out_map128, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return trr_make_tensor_map_ex(out_map128, base, rows, cols, 2, BK, box_rows);
}

cudaError_t trr_launch_gemm_topk_dump(const GemmTopkArgs& a, const void* map_q128, const void* map_d128, unsigned grid,
                                      float* dump, uint32_t dump_ld, cudaStream_t st);
cudaError_t trr_launch_gemm_topk(const GemmTopkArgs& a, const void* map_q128, const void* map_d128, unsigned grid,
                                 cudaStream_t st) {
  return trr_launch_gemm_topk_dump(a, map_q128, map_d128, grid, nullptr, 0, st);
}

// debug: additionally dumps every fast score to dump[(q) * dump_ld + doc] (small problems only)
cudaError_t trr_launch_gemm_topk_dump(const GemmTopkArgs& a, const void* map_q128, const void* map_d128, unsigned grid,
                                      float* dump, uint32_t dump_ld, cudaStream_t st) {
  if (grid == 0) return cudaSuccess;
  CUtensorMap mq, md;
  memcpy(&mq, map_q128, 128);
  memcpy(&md, map_d128, 128);
  if (a.pair_mode) {  // grid and a.n_qblocks are even: consecutive CTAs form the 2-CTA clusters
    cudaError_t e = cudaFuncSetAttribute(dense_gemm_topk_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SMEM2_TOTAL);
    if (e != cudaSuccess) return e;
    dense_gemm_topk_pair_kernel<<<grid, GEMM2_THREADS, SMEM2_TOTAL, st>>>(mq, md, a, dump, dump_ld);
  } else {
    cudaError_t e = cudaFuncSetAttribute(dense_gemm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SMEM_TOTAL);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(dense_gemm_topk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    dense_gemm_topk_kernel<<<grid, GEMM1_THREADS, SMEM_TOTAL, st>>>(mq, md, a, dump, dump_ld);
  }
  return cudaGetLastError();
}
