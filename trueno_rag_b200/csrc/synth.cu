// synth.cu — device-side synthetic corpus generator (tests / benches only; compiled with -fmad=false).
// Bit-identical to oracle/trr_oracle.c:orc_synth_corpus_rows; both follow csrc/synth_spec.h.  Not reference code.
#include "common.cuh"
#include "synth_spec.h"

template <int TO_BF16>
__global__ void synth_rows_kernel(uint64_t seed, uint64_t first_row, uint64_t n, uint32_t dim, int dups, void* out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t row = first_row + i;
  const uint64_t src = trr_dup_source(seed, row, dups);
  float s = 0.0f;
  for (uint32_t j = 0; j < dim; ++j) {
    const float x = trr_uniform_pm1(trr_hash4(seed, TRR_STREAM_CORPUS, src, j));
    s = s + x * x;
  }
  const float nrm = sqrtf(s);
  for (uint32_t j = 0; j < dim; ++j) {
    float x = trr_uniform_pm1(trr_hash4(seed, TRR_STREAM_CORPUS, src, j));
    if (nrm > 0.0f) x = x / nrm;
    if (TO_BF16) reinterpret_cast<uint16_t*>(out)[i * dim + j] = trr_f32_to_bf16_bits(x);
    else reinterpret_cast<float*>(out)[i * dim + j] = x;
  }
}

cudaError_t trr_launch_synth_rows(uint64_t seed, uint64_t first_row, uint64_t n, uint32_t dim, int dups, int to_bf16,
                                  void* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (to_bf16) synth_rows_kernel<1><<<grid, 128, 0, st>>>(seed, first_row, n, dim, dups, out);
  else synth_rows_kernel<0><<<grid, 128, 0, st>>>(seed, first_row, n, dim, dups, out);
  return cudaGetLastError();
}

__global__ void flush_kernel(uint32_t* p, size_t n_words, uint32_t v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

cudaError_t trr_launch_flush(void* p, size_t bytes, cudaStream_t st) {
  if (bytes < 4) return cudaSuccess;
  flush_kernel<<<1184, 256, 0, st>>>(reinterpret_cast<uint32_t*>(p), bytes / 4, 0x5EED5EEDu);
  return cudaGetLastError();
}
