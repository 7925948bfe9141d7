// Microbenchmark (triage only): cost of K small cp.async.bulk copies per stage (3-stage ring, consumers only wait/release).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NS = 3;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void __launch_bounds__(640, 1) k(const uint8_t* __restrict__ src, size_t src_bytes, uint32_t K, uint32_t S, uint32_t passes, uint32_t stride, long long* cyc, uint32_t PW) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_bar[NS], empty_bar[NS];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full_bar[s])), "r"(PW));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty_bar[s])), "r"(16));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  uint32_t stage = 0, phase = 0;
  if (warp >= 16) {
    if (warp - 16 >= PW) return;
    const uint32_t kw = K / PW, k0 = (warp - 16) * kw;
    size_t off = ((size_t)blockIdx.x * 7919u * 4096u) % (src_bytes / 2);
    for (uint32_t p = 0; p < passes; ++p) {
      while (!try_wait(&empty_bar[stage], phase ^ 1)) {}
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full_bar[stage])), "r"(kw * S) : "memory");
      __syncwarp();
      if (lane < kw) {
        const uint8_t* g = src + ((off + (size_t)(k0 + lane) * stride) % (src_bytes - 65536));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(smem + (size_t)stage * 32768 + (k0 + lane) * S)), "l"(g), "r"(S), "r"(s32(&full_bar[stage])) : "memory");
      }
      off += (size_t)K * S + 1048576;
      off &= ~(size_t)15;
      if (++stage == NS) { stage = 0; phase ^= 1; }
    }
  } else {
    for (uint32_t p = 0; p < passes; ++p) {
      while (!try_wait(&full_bar[stage], phase)) {}
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty_bar[stage])) : "memory");
      if (++stage == NS) { stage = 0; phase ^= 1; }
    }
  }
  if (tid == 0) cyc[blockIdx.x] = clock64() - t0;
}
int main() {
  const size_t bytes = (size_t)4 << 30;
  uint8_t* src; long long* cyc;
  cudaMalloc(&src, bytes); cudaMalloc(&cyc, 148 * 8);
  cudaMemset(src, 1, bytes);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768);
  const uint32_t passes = 2000;
  struct { uint32_t K, S, stride, PW; } cfg[] = {{16, 1024, 1 << 20, 1}, {16, 1024, 1 << 20, 2}, {16, 1024, 1 << 20, 4}, {32, 1024, 1 << 20, 4},
                                                 {32, 1024, 1 << 20, 2}, {16, 2048, 1 << 20, 4}, {4, 8192, 1 << 20, 4}, {4, 8192, 1 << 20, 1}};
  for (auto c : cfg) {
    k<<<148, 640, 3 * 32768>>>(src, bytes, c.K, c.S, passes, c.stride, cyc, c.PW);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
    printf("PW=%u K=%2u copies x %5u B (%6u B/pass): %7.0f cyc/pass  -> %6.1f B/cyc/SM  (%s)\n", c.PW, c.K, c.S, c.K * c.S, avg / passes, c.K * c.S / (avg / passes), cudaGetErrorString(e));
  }
  return 0;
}
