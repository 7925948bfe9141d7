#include <cstdint>
__global__ void k(uint32_t* out) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) asm volatile("st.bulk.weak.shared::cta [%0], %1, 0;" ::"r"(a), "l"((uint64_t)131072) : "memory");
  __syncthreads();
  out[threadIdx.x] = reinterpret_cast<uint32_t*>(smem)[threadIdx.x];
}
