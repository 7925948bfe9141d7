// Microbenchmark (triage only): the consumer loop of bm25_fast_kernel in isolation (postings already in shared memory):
// flat atomic accumulate + named barrier + scan/zero + barrier.red.or, 16 consumer warps + 1 idle warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int CT = 512;
constexpr int CELLS = 32768;
constexpr int ITEMS = 256;
__host__ __device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory"); }
__device__ __forceinline__ bool cbar_or(bool pred) {
  uint32_t r;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbarrier.red.or.pred q, 1, %2, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
               : "=r"(r) : "r"((uint32_t)pred), "n"(CT) : "memory");
  return r != 0;
}
// MODE bit0: F2I path; bit1: barrier.red.or at the end (else bar.sync); bit2: skip scan; bit3: skip atomics
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void __launch_bounds__(CT + 32, 1) k(const uint2* __restrict__ post, uint32_t n_post, uint32_t* out, long long* cyc, float scale, const uint8_t* src, uint32_t K, uint32_t S) {
  __shared__ uint64_t tbar[3];
  __shared__ volatile uint32_t s_iter;
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem);
  uint2* st = reinterpret_cast<uint2*>(smem + CELLS * 4);
  const uint32_t tid = threadIdx.x;
  if (tid == 0) { for (int j = 0; j < 3; ++j) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&tbar[j])), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); s_iter = 0; }
  __syncthreads();
  if (tid >= CT) {
    // side warp: per consumer iteration, K bulk copies of S bytes into a ring nobody reads (3 in flight)
    if (K == 0) return;
    const uint32_t lane = tid & 31;
    uint8_t* ring = smem + CELLS * 4 + 8192 * 8;
    size_t off = (size_t)blockIdx.x * 33554432u;
    for (int it = 0; it < ITEMS; ++it) {
      const int st3 = it % 3;
      if (it >= 3) { uint32_t ok = 0; const uint32_t par = ((it / 3) - 1) & 1; while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&tbar[st3])), "r"(par) : "memory"); }
      while ((int)s_iter < it - 2) __nanosleep(50);
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&tbar[st3])), "r"(K * S) : "memory");
      __syncwarp();
      if (lane < K) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(ring + st3 * 32768 + lane * S)), "l"(src + ((off + (size_t)lane * 1048576u) & 0xFFFFFFF0ull) % ((size_t)3 << 30)), "r"(S), "r"(s32(&tbar[st3])) : "memory");
      off += 65536 + K * S;
    }
    return;
  }
  for (uint32_t i = tid; i < CELLS; i += CT) acc[i] = 0;
  for (uint32_t i = tid; i < n_post; i += CT) st[i] = post[(size_t)blockIdx.x * 8192 + i];
  cbar();
  const long long t0 = clock64();
  uint32_t sink = 0;
  const uint32_t range_base = 7, R = CELLS, thr_hi = 0x7fffff00u;
  for (int it = 0; it < ITEMS; ++it) {
    if (tid == 0) s_iter = it;
    if (!(MODE & 8)) {
      auto add1 = [&](const uint2 e) {
        const uint32_t dd = e.x - range_base;
        if (MODE & 1) { if (dd < R) atomicAdd(&acc[dd], __float2uint_ru(__uint_as_float(e.y) * scale)); }
        else { if (dd < R) atomicAdd(&acc[dd], e.y & 0xFFu); }
      };
      uint32_t p = tid;
      for (; p + 3 * CT < n_post; p += 4 * CT) {
        const uint2 e0 = st[p], e1 = st[p + CT], e2 = st[p + 2 * CT], e3 = st[p + 3 * CT];
        add1(e0); add1(e1); add1(e2); add1(e3);
      }
      for (; p < n_post; p += CT) add1(st[p]);
    }
    cbar();
    bool need = false;
    if (!(MODE & 4)) {
      uint4* a4 = reinterpret_cast<uint4*>(acc);
      auto pre4 = [&](uint32_t i, const uint4& v) -> bool {
        const uint32_t mx = max(max(v.x, v.y), max(v.z, v.w));
        const bool hit = mx >= thr_hi;
        if (mx != 0u && !hit) a4[i] = make_uint4(0u, 0u, 0u, 0u);
        return hit;
      };
      const uint32_t n4 = R >> 2;
      uint32_t i = tid;
      for (; i + 3 * CT < n4; i += 4 * CT) {
        const uint4 v0 = a4[i], v1 = a4[i + CT], v2 = a4[i + 2 * CT], v3 = a4[i + 3 * CT];
        const bool h0 = pre4(i, v0), h1 = pre4(i + CT, v1), h2 = pre4(i + 2 * CT, v2), h3 = pre4(i + 3 * CT, v3);
        if (h0 | h1 | h2 | h3) { sink += i; need = true; }
      }
    }
    if (MODE & 2) { if (cbar_or(need)) sink += 1; } else cbar();
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  uint32_t s2 = sink;
  for (uint32_t i = tid; i < CELLS; i += CT) s2 += acc[i];
  if (s2 == 0x12345678u) out[0] = s2;
}
template <int MODE>
void run(const char* name, const uint2* post, uint32_t n_post, uint32_t* out, long long* cyc, const uint8_t* src = nullptr, uint32_t K = 0, uint32_t S = 0) {
  const size_t smem = (size_t)CELLS * 4 + 8192 * 8 + (K ? 3 * 32768 : 0);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<148, CT + 32, smem>>>(post, n_post, out, cyc, 1024.0f, src, K, S);
  k<MODE><<<148, CT + 32, smem>>>(post, n_post, out, cyc, 1024.0f, src, K, S);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
  printf("%-64s n=%5u K=%2u S=%5u %8.0f cyc/item (%s)\n", name, n_post, K, S, avg / ITEMS, cudaGetErrorString(e));
}
int main() {
  uint2* hp = new uint2[(size_t)148 * 8192];
  uint32_t s = 12345;
  for (size_t i = 0; i < (size_t)148 * 8192; ++i) { hp[i].x = 7 + rng(s) % CELLS; float f = 0.5f + (rng(s) & 0xFF) / 64.0f; hp[i].y = *reinterpret_cast<uint32_t*>(&f); }
  uint2* post; uint32_t* out; long long* cyc;
  cudaMalloc(&post, (size_t)148 * 8192 * 8); cudaMalloc(&out, 8); cudaMalloc(&cyc, 148 * 8);
  cudaMemcpy(post, hp, (size_t)148 * 8192 * 8, cudaMemcpyHostToDevice);
  uint8_t* src; cudaMalloc(&src, ((size_t)3 << 30) + (64 << 20)); cudaMemset(src, 1, ((size_t)3 << 30) + (64 << 20));
  run<3>("RED(F2I) + bar + scan + bar.red.or", post, 4096, out, cyc);
  run<3>("same + side warp: TMA 16 x 2 KB per iteration", post, 4096, out, cyc, src, 16, 2048);
  run<3>("same + side warp: TMA 16 x 16 B per iteration", post, 4096, out, cyc, src, 16, 16);
  run<3>("same + side warp: TMA 2 x 16 KB per iteration", post, 4096, out, cyc, src, 2, 16384);
  run<3>("same + side warp: TMA 32 x 1 KB per iteration", post, 4096, out, cyc, src, 32, 1024);
  for (uint32_t n : {4096u}) {
    run<0>("RED + bar + scan + bar", post, n, out, cyc);
    run<1>("RED(F2I) + bar + scan + bar", post, n, out, cyc);
    run<3>("RED(F2I) + bar + scan + bar.red.or", post, n, out, cyc);
    run<7>("RED(F2I) + bar + [no scan] + bar.red.or", post, n, out, cyc);
    run<11>("[no atomics] + bar + scan + bar.red.or", post, n, out, cyc);
    run<15>("[nothing] bar + bar.red.or", post, n, out, cyc);
  }
  return 0;
}
