// fusion.cuh — argument block of the fusion kernel (fusion.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct FuseArgs {
  // source 0 = dense, 1 = sparse.  Shard g's list of query b starts at element g*shard_stride + b*C
  // (ord/score) and its length is n[src][g*shard_stride + b]; a NULL n[src] means "source absent".
  const uint32_t* ord[2];
  const float* score[2];
  const uint32_t* n[2];
  uint64_t shard_stride;  // in 4-byte elements; ignored when G == 1
  uint32_t G, B, C;
  int strategy;
  float param;
  uint32_t k;             // output row length
  uint32_t mcap;          // power of two >= G*C (merge buffer; unused when G == 1 but must be >= 1)
  uint32_t fcap;          // power of two >= 2*C
  uint32_t* out_ord;
  float* out_fused;
  float* out_dense;       // nullable
  float* out_sparse;      // nullable
  uint32_t* out_n;
};

size_t trr_fuse_smem(const FuseArgs& a);
cudaError_t trr_launch_fuse(const FuseArgs& a, cudaStream_t st);

// peer-memory exchange (exchange.cu)
constexpr uint32_t TRR_MAX_GROUP = 16;
struct ExchangeScatterArgs {
  const uint8_t* record;                 // this rank's exchange record (record_bytes, multiple of 16)
  uint64_t record_bytes;
  uint8_t* peer_gath[TRR_MAX_GROUP];     // base of every rank's gather buffer (own buffer for g == rank)
  uint32_t* peer_flags[TRR_MAX_GROUP];   // base of every rank's flag array
  uint64_t slot_off;                     // byte offset of (parity, this rank's slot) inside a gather buffer
  uint32_t flag_index;                   // parity * world + rank
  uint32_t seq;                          // sequence number of the call (> 0)
  uint32_t* block_counter;               // [world] zero-initialised device scratch
};
cudaError_t trr_launch_exchange_scatter(const ExchangeScatterArgs& a, uint32_t world, cudaStream_t st);
cudaError_t trr_launch_exchange_wait(const uint32_t* flags, uint32_t world, uint32_t seq, uint32_t* dbg, cudaStream_t st);
