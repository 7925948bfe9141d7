// exchange.cu — peer-memory exchange of the shard-local top-C records without a collective launch.
//
// Every rank owns a gather buffer (`world` record slots per parity) and a flag array, both mapped into every other
// process of the node through CUDA IPC.  After its shard-local kernels a rank runs ONE kernel that stores its record
// into slot `rank` of every peer's gather buffer over NVLink (st.global, 16 bytes per thread) and, once all of a
// peer's blocks are done, publishes the call's sequence number in that peer's flag array (release at system scope).
// The merge + fusion kernel of a rank is preceded by a wait kernel that spins (bounded) until all `world` flags of the
// parity carry the sequence number.  Reference: src/retrieve.rs:175-220 (the two searches, then the fusion).
#include "common.cuh"
#include <algorithm>

#include "fusion.cuh"

__global__ void __launch_bounds__(256)
exchange_scatter_kernel(ExchangeScatterArgs a) {
  const uint32_t g = blockIdx.y;  // destination rank
  uint4* dst = reinterpret_cast<uint4*>(a.peer_gath[g] + a.slot_off);
  const uint4* src = reinterpret_cast<const uint4*>(a.record);
  const uint64_t n16 = a.record_bytes >> 4;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
  __threadfence_system();  // this thread's stores are visible system-wide before the block counts itself done
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t done = atomicAdd(&a.block_counter[g], 1u) + 1u;
    if (done == gridDim.x) {
      a.block_counter[g] = 0;
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_flags[g] + a.flag_index), "r"(a.seq) : "memory");
    }
  }
}

__global__ void exchange_wait_kernel(const uint32_t* flags, uint32_t world, uint32_t seq, uint32_t* dbg) {
  const uint32_t g = threadIdx.x;
  if (g >= world) return;
  const long long t0 = clock64();
  while (true) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + g) : "memory");
    if (v == seq) break;
    __nanosleep(200);
    if (clock64() - t0 > 20000000000LL) {  // ~10 s: a missing peer must surface as an error, never as a hung GPU
      if (dbg) { *dbg = 0x50000000u | g; __threadfence_system(); }
      __trap();
    }
  }
}

cudaError_t trr_launch_exchange_scatter(const ExchangeScatterArgs& a, uint32_t world, cudaStream_t st) {
  const unsigned blocks = (unsigned)std::min<uint64_t>(64, ((a.record_bytes >> 4) + 255) / 256);
  exchange_scatter_kernel<<<dim3(std::max(blocks, 1u), world), 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t trr_launch_exchange_wait(const uint32_t* flags, uint32_t world, uint32_t seq, uint32_t* dbg, cudaStream_t st) {
  exchange_wait_kernel<<<1, 32 * ((world + 31) / 32), 0, st>>>(flags, world, seq, dbg);
  return cudaGetLastError();
}
