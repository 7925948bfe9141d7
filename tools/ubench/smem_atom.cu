// Microbenchmark (triage only, not product): shared-memory update primitives on random cells of a 128 KB array.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atom smem_atom.cu ; run on one B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CELLS = 32768;
constexpr int THREADS = 512;
constexpr int N_POST = 4096;   // postings per "item"
constexpr int ITEMS = 256;

__host__ __device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k(const uint2* __restrict__ post, uint32_t* out, long long* cyc) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem);
  uint2* st = reinterpret_cast<uint2*>(smem + CELLS * 4);
  const int tid = threadIdx.x;
  for (int i = tid; i < CELLS; i += THREADS) acc[i] = 0;
  for (int i = tid; i < N_POST; i += THREADS) st[i] = post[(size_t)blockIdx.x * N_POST + i];
  __syncthreads();
  const long long t0 = clock64();
  uint32_t sink = 0;
  for (int it = 0; it < ITEMS; ++it) {
    if (MODE == 0) {  // RED.u32
#pragma unroll 4
      for (int p = tid; p < N_POST; p += THREADS) { const uint2 e = st[p]; atomicAdd(&acc[e.x], e.y); }
    } else if (MODE == 1) {  // ATOMS.u32 with return + crossing check
#pragma unroll 4
      for (int p = tid; p < N_POST; p += THREADS) {
        const uint2 e = st[p];
        const uint32_t old = atomicAdd(&acc[e.x], e.y);
        if ((uint32_t)(0x7fffff00u - old - 1u) < e.y) sink += e.x;
      }
    } else if (MODE == 2) {  // red.f32
      float* facc = reinterpret_cast<float*>(acc);
#pragma unroll 4
      for (int p = tid; p < N_POST; p += THREADS) { const uint2 e = st[p]; atomicAdd(&facc[e.x], __uint_as_float(e.y)); }
    } else if (MODE == 3) {  // plain RMW (racy; cost reference)
#pragma unroll 4
      for (int p = tid; p < N_POST; p += THREADS) { const uint2 e = st[p]; acc[e.x] = acc[e.x] + e.y; }
    } else if (MODE == 4) {  // ATOMS.EXCH
#pragma unroll 4
      for (int p = tid; p < N_POST; p += THREADS) { const uint2 e = st[p]; sink += atomicExch(&acc[e.x], 0u); }
    } else if (MODE == 5) {  // zero fill STS.128
      uint4* a4 = reinterpret_cast<uint4*>(acc);
#pragma unroll 4
      for (int i = tid; i < CELLS / 4; i += THREADS) a4[i] = make_uint4(0, 0, 0, 0);
    } else if (MODE == 6) {  // scan LDS.128 + max + predicated zero
      uint4* a4 = reinterpret_cast<uint4*>(acc);
#pragma unroll 4
      for (int i = tid; i < CELLS / 4; i += THREADS) {
        const uint4 v = a4[i];
        const uint32_t m = max(max(v.x, v.y), max(v.z, v.w));
        if (m >= 0x7fffff00u) sink += m;
        if (m) a4[i] = make_uint4(0, 0, 0, 0);
      }
    } else if (MODE == 7) {  // match.any per 32 postings + plain RMW by group leaders
#pragma unroll 2
      for (int p = tid; p < N_POST; p += THREADS) {
        const uint2 e = st[p];
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, e.x);
        if ((grp & ((1u << (tid & 31)) - 1u)) == 0) acc[e.x] = acc[e.x] + e.y;
      }
    } else if (MODE == 8) {  // 16-bit packed RED (two cells per word)
#pragma unroll 4
      for (int p = tid; p < N_POST; p += THREADS) { const uint2 e = st[p]; atomicAdd(&acc[e.x >> 1], (e.y & 0xFFu) << ((e.x & 1u) * 16u)); }
    } else if (MODE == 9) {  // RED.u32 with only 8 of 32 lanes active (does cost scale with active lanes?)
#pragma unroll 4
      for (int p = tid; p < N_POST; p += THREADS) { const uint2 e = st[p]; if ((tid & 3) == 0) atomicAdd(&acc[e.x], e.y); }
    } else if (MODE == 10) {  // RED.u32, conflict-free addresses (lane i -> bank i)
#pragma unroll 4
      for (int p = tid; p < N_POST; p += THREADS) { const uint2 e = st[p]; atomicAdd(&acc[(e.x & ~31u) | (tid & 31)], e.y); }
    } else if (MODE == 11) {  // plain RMW conflict-free addresses
#pragma unroll 4
      for (int p = tid; p < N_POST; p += THREADS) { const uint2 e = st[p]; const uint32_t a = (e.x & ~31u) | (tid & 31); acc[a] = acc[a] + e.y; }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  uint32_t s2 = sink;
  for (int i = tid; i < CELLS; i += THREADS) s2 += acc[i];
  if (s2 == 0x12345678u) out[0] = s2;
}

template <int MODE>
void run(const char* name, const uint2* post, uint32_t* out, long long* cyc) {
  const size_t smem = CELLS * 4 + N_POST * 8;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<148, THREADS, smem>>>(post, out, cyc);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<148, THREADS, smem>>>(post, out, cyc);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
  const double units = (MODE == 5 || MODE == 6) ? (double)CELLS : (double)N_POST;
  printf("%-44s %8.3f ms  %9.0f cyc/item  %7.3f cyc per 32 units  (%s)\n", name, ms, avg / ITEMS, avg / ITEMS / units * 32.0,
         cudaGetErrorString(e));
}

int main() {
  uint2* hp = new uint2[(size_t)148 * N_POST];
  uint32_t s = 12345;
  for (size_t i = 0; i < (size_t)148 * N_POST; ++i) { hp[i].x = rng(s) % CELLS; hp[i].y = 1 + (rng(s) & 0xFF); }
  uint2* post; uint32_t* out; long long* cyc;
  cudaMalloc(&post, (size_t)148 * N_POST * 8); cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
  cudaMemcpy(post, hp, (size_t)148 * N_POST * 8, cudaMemcpyHostToDevice);
  run<0>("RED.u32 random", post, out, cyc);
  run<1>("ATOMS.u32 ret + crossing check random", post, out, cyc);
  run<2>("red.f32 random", post, out, cyc);
  run<3>("plain RMW random (racy)", post, out, cyc);
  run<4>("ATOMS.EXCH random", post, out, cyc);
  run<5>("zero fill STS.128 (per 32 cells)", post, out, cyc);
  run<6>("scan LDS.128+max+pred zero (per 32 cells)", post, out, cyc);
  run<7>("match.any + leader RMW", post, out, cyc);
  run<8>("RED.u32 packed 2x16", post, out, cyc);
  run<9>("RED.u32 8/32 lanes active (per 32 slots)", post, out, cyc);
  run<10>("RED.u32 conflict-free banks", post, out, cyc);
  run<11>("plain RMW conflict-free banks", post, out, cyc);
  return 0;
}
