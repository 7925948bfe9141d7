// dense.cuh — argument blocks and launchers shared by the dense translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// K1 scan (dense_scan.cu)
struct DenseScanArgs {
  const uint8_t* rows;       // slab, row-major, row_bytes per row
  uint32_t row_bytes;
  uint32_t dim;
  uint64_t n_rows;           // rows in the slab (including tombstoned ones)
  const float* norms;        // per-row norm (reference order)
  const uint8_t* dead;       // nullable tombstones
  const float* q;            // all queries, B x dim f32 (device)
  const float* q_norms;      // B
  const uint32_t* sel;       // nullable: indices of the queries to run
  const uint32_t* n_sel_ptr; // nullable: device count of `sel` (else n_sel)
  uint32_t n_sel;
  uint32_t ch_bytes;         // bulk kernel: bytes of a row staged per chunk (multiple of 16)
  uint32_t n_chunks;
  uint32_t n_slots;          // TMA kernel: 4 KB ring slots per warp
  uint32_t k;
  uint32_t cap;              // per-warp top-k buffer capacity (power of two >= k + 32)
  uint32_t base_ord;
  uint64_t* partial;         // [n_sel][n_warps_total][k] keys
  uint32_t* partial_n;       // [n_sel][n_warps_total]
  // TMA kernel, fused epilogue (single-call latency: VectorStore::search is one query): the last CTA to finish a query
  // group merges the per-CTA lists and writes the results, so the search is ONE launch; q_norms may then be NULL and the
  // kernel computes |q| itself (reference order).  Used when done != NULL.
  uint32_t q_in_param;       // 1: the (single) query and its norm come from the ScanQueryParam kernel parameter
  uint32_t* done;            // [ceil(n_sel / NQ)] arrival counters, zero before the launch; the kernel re-zeroes them
  uint64_t* out_keys;        // nullable [n_sel][k]
  uint32_t* out_ord;         // nullable [n_sel][k]
  float* out_score;          // nullable
  uint32_t* out_n;           // nullable [n_sel]
};

// A single query small enough for the kernel-parameter space (4 KB in all) travels with the launch: no host -> device copy
// operation, every CTA reads it through the constant cache.  DenseScanArgs::q_in_param selects it (one query, NQ = 1).
constexpr uint32_t TRR_SCAN_PARAM_DIM = 768;
struct ScanQueryParam {
  float norm;                      // |q| in reference order (computed on the host)
  float v[TRR_SCAN_PARAM_DIM];
};

// generic merge of key lists into a canonical top-k (dense_scan.cu)
struct TopkMergeArgs {
  const uint64_t* lists;     // [n_rows][n_lists][list_stride]
  const uint32_t* list_n;    // nullable [n_rows][n_lists]; else every list is full (empty keys are skipped)
  uint32_t n_lists;
  uint32_t list_stride;
  uint32_t n_rows;
  const uint32_t* n_rows_ptr;// nullable device override of n_rows
  const uint32_t* row_map;   // nullable: output row of input row i
  uint32_t k;
  uint32_t k2;               // power of two >= k
  uint64_t* out_keys;        // nullable [rows][k]
  uint32_t* out_ord;         // nullable
  float* out_score;          // nullable
  uint32_t* out_n;           // nullable
};

// exact rescoring + candidate proof (dense_scan.cu)
struct RescoreArgs {
  const float* cand_score;   // [n_slices][n_qblocks][128][cps] fast scores
  const uint32_t* cand_ord;  // same shape, 0xFFFFFFFF = empty
  uint32_t n_slices, n_qblocks;
  uint32_t cps;              // candidates kept per (query, slice) by the tensor-core pass
  uint32_t cp;               // candidates re-scored exactly per query (power of two, >= k)
  uint32_t cap2;             // power of two >= max(n_slices*cps, cp + 1)
  const uint32_t* gthr;      // nullable [n_qblocks*128]: the largest threshold the tensor-core pass applied to the query (orderable)
  const void* rows;          // slab in the store dtype
  uint32_t dim;
  const float* norms;
  uint32_t base_ord;
  uint64_t n_live;           // live rows in the store
  const float* q;            // B x dim f32
  const float* q_norms;      // B (ignored when q_norms_out is set)
  float* q_norms_out;        // nullable: the kernel computes ||q|| itself (reference order) and stores it here
  const float* q_delta;      // B: || q - bf16(q) ||  (0 when the query is bf16-exact)
  const float* max_norm;     // device scalar: max row norm
  uint32_t B, k;
  int metric;
  uint32_t stage_chunk;      // bytes of each candidate row staged in shared memory per pass (multiple of 16; 0 = read rows from global)
  float eps_rel;             // accumulation error bound relative to |q||d|
  uint64_t* out_keys;        // nullable [B][k]
  uint32_t* out_ord;         // nullable [B][k]
  float* out_score;          // nullable
  uint32_t* out_n;           // nullable
  const uint32_t* sel;       // nullable: wide pass, CTA i handles query sel[i] ...
  const uint32_t* sel_n;     // ... for i < *sel_n (device count)
  uint32_t* n_resolved;      // nullable: counts the queries whose proof holds in this pass (diagnostics)
  uint32_t* flags;           // [B] 1 = proof failed
  uint32_t* flagged;         // [B] compact list of flagged queries
  uint32_t* n_flagged;       // device counter (zeroed by the caller)
  float* max_gap;            // device scalar: max |fast - exact| seen (zeroed by the caller)
};

// K2 tensor-core pass (dense_gemm.cu)
struct GemmTopkArgs {
  uint32_t n_qblocks;        // ceil(B / 128)
  uint32_t n_slices;         // document slices (CTAs per query block)
  uint32_t n_tiles;          // 256-document tiles in the slab
  uint32_t k_blocks;         // ceil(dim / 64)
  uint32_t base_ord;
  const float2* scale_bias;  // [n_tiles*256]
  uint32_t cps;              // candidates kept per (query, half slice): 8, 16 or 32
  float* cand_score;         // [n_slices][n_qblocks][128][cps]
  uint32_t* cand_ord;
  uint32_t* gthr;            // [n_qblocks*128] shared running thresholds (orderable-encoded), zeroed by the caller
  int share_thresholds;
  uint32_t* pub;             // nullable [2 * n_slices][n_qblocks*128]: every list's pub_rank-th best fast score (orderable; 0 = none),
                             // zeroed by the caller; the pair kernel's helper warps turn it into merged thresholds (see there)
  uint32_t pub_rank;         // j: 1..8; 0 = off
  uint32_t pub_pick;         // s: 1..8, the threshold is the s-th smallest published value: (2 * n_slices - s + 1) * j >= re-scoring width
  uint32_t* dbg;             // nullable host-mapped word: site of a barrier timeout
  int pair_mode;             // 1 = 2-CTA kernel (cta_group::2); needs an even n_qblocks
  int debug_mode;            // 0 = normal; 1 = epilogue skips TMEM reads; 2 = reads but never inserts (perf triage only)
};

void trr_launch_norms(int is_bf16, const void* rows, uint32_t dim, uint64_t row0, uint64_t n, float* norms,
                      cudaStream_t st);
void trr_launch_gemm_operands(const float* norms, const uint8_t* dead, uint64_t n_rows, uint64_t n_padded, int metric,
                              float2* scale_bias, float* max_norm, cudaStream_t st);
void trr_launch_query_norms(const float* q, uint32_t dim, uint32_t B, float* qn, cudaStream_t st);
cudaError_t trr_launch_scan(const DenseScanArgs& a, int is_bf16, int metric, bool bulk, unsigned grid, size_t smem,
                            cudaStream_t st);
// K1 with a 2-D TMA ring (rows of a multiple of 16 bytes): map_rows128 = tensor map of the slab, box 32 rows x 128 bytes
// nq = 1 or 4: queries sharing one pass over the slab
// host_q (nullable): the single query (a.dim floats, host memory) + host_q_norm to pass as a kernel parameter (a.q_in_param)
cudaError_t trr_launch_scan_tma(const DenseScanArgs& a, const void* map_rows128, int is_bf16, int metric, unsigned grid,
                                unsigned n_warps, uint32_t nq, size_t smem, cudaStream_t st, const float* host_q = nullptr,
                                float host_q_norm = 0.0f);
size_t trr_scan_tma_smem(uint32_t dim, uint32_t cap, uint32_t n_slots, uint32_t n_warps, uint32_t nq);
cudaError_t trr_launch_topk_merge(const TopkMergeArgs& a, unsigned grid, cudaStream_t st);
cudaError_t trr_launch_rescore(const RescoreArgs& a, int is_bf16, cudaStream_t st);

// dense_gemm.cu
constexpr uint32_t TRR_GEMM_CP = 64;      // exact re-scoring width per query (doubled for k > 50)
constexpr uint32_t TRR_GEMM_CPS_MAX = 32; // maximum list length per (query, half slice)
constexpr uint32_t TRR_GEMM_TILE_N = 256; // documents per MMA tile
constexpr uint32_t TRR_GEMM_TILE_M = 128; // queries per CTA
// converts B x dim f32 queries to a zero-padded [n_qblocks*128][dim_pad] bf16 matrix and reports ||q - bf16(q)||
void trr_launch_query_prep(const float* q, uint32_t dim, uint32_t dim_pad, uint32_t B, uint32_t B_pad, uint16_t* q_bf16,
                           float* q_delta, float* q_norm, cudaStream_t st);
// f32 slab -> bf16 shadow (round to nearest even), zero-padded to dim_pad columns
void trr_launch_shadow(const void* rows, int is_bf16, uint32_t dim, uint32_t dim_pad, uint64_t row0, uint64_t n,
                       uint16_t* shadow, cudaStream_t st);
// encodes a 2-D K-major bf16 tensor map (rows x cols, box = box_rows x 64, 128-byte swizzle); 128-byte opaque blob
int trr_make_tensor_map(void* out_map128, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
// general form: element size 2 (bf16) or 4 (f32), box = box_rows x box_cols elements (box_cols * elem_bytes == 128)
int trr_make_tensor_map_ex(void* out_map128, const void* base, uint64_t rows, uint64_t cols, uint32_t elem_bytes,
                           uint32_t box_cols, uint32_t box_rows);
cudaError_t trr_launch_gemm_topk(const GemmTopkArgs& a, const void* map_q128, const void* map_d128, unsigned grid,
                                 cudaStream_t st);
