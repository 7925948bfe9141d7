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
template <int MODE>
__global__ void __launch_bounds__(CT + 32, 1) k(const uint2* __restrict__ post, uint32_t n_post, uint32_t* out, long long* cyc, float scale) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem);
  uint2* st = reinterpret_cast<uint2*>(smem + CELLS * 4);
  const uint32_t tid = threadIdx.x;
  if (tid >= CT) return;
  for (uint32_t i = tid; i < CELLS; i += CT) acc[i] = 0;
  for (uint32_t i = tid; i < n_post; i += CT) st[i] = post[(size_t)blockIdx.x * 8192 + i];
  cbar();
  const long long t0 = clock64();
  uint32_t sink = 0;
  const uint32_t range_base = 7, R = CELLS, thr_hi = 0x7fffff00u;
  for (int it = 0; it < ITEMS; ++it) {
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
void run(const char* name, const uint2* post, uint32_t n_post, uint32_t* out, long long* cyc) {
  const size_t smem = (size_t)CELLS * 4 + 8192 * 8;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<148, CT + 32, smem>>>(post, n_post, out, cyc, 1024.0f);
  k<MODE><<<148, CT + 32, smem>>>(post, n_post, out, cyc, 1024.0f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
  printf("%-64s n=%5u %8.0f cyc/item (%s)\n", name, n_post, avg / ITEMS, cudaGetErrorString(e));
}
int main() {
  uint2* hp = new uint2[(size_t)148 * 8192];
  uint32_t s = 12345;
  for (size_t i = 0; i < (size_t)148 * 8192; ++i) { hp[i].x = 7 + rng(s) % CELLS; float f = 0.5f + (rng(s) & 0xFF) / 64.0f; hp[i].y = *reinterpret_cast<uint32_t*>(&f); }
  uint2* post; uint32_t* out; long long* cyc;
  cudaMalloc(&post, (size_t)148 * 8192 * 8); cudaMalloc(&out, 8); cudaMalloc(&cyc, 148 * 8);
  cudaMemcpy(post, hp, (size_t)148 * 8192 * 8, cudaMemcpyHostToDevice);
  for (uint32_t n : {4096u, 2048u}) {
    run<0>("RED + bar + scan + bar", post, n, out, cyc);
    run<1>("RED(F2I) + bar + scan + bar", post, n, out, cyc);
    run<3>("RED(F2I) + bar + scan + bar.red.or", post, n, out, cyc);
    run<7>("RED(F2I) + bar + [no scan] + bar.red.or", post, n, out, cyc);
    run<11>("[no atomics] + bar + scan + bar.red.or", post, n, out, cyc);
    run<15>("[nothing] bar + bar.red.or", post, n, out, cyc);
  }
  return 0;
}
