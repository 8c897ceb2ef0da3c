// gather_bench2.cu -- the random-sector ceiling the presence-probe kernel (k_probe_singles) is judged
// against: every thread issues EIGHT independent 4-byte reads at hashed addresses of a table far
// larger than L2, with nothing else to do (one multiply-xorshift per address, power-of-two table).
// The first version of this tool (gather_bench.cu) took a 64-bit modulo per address and kept one read
// in flight per thread: it measured its own arithmetic (37 G gathers/s), not the memory system.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench2 gather_bench2.cu
//   ./gather_bench2 [log2 table bytes = 35]      (prints G gathers/s per load flavour)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z *= 0x9E3779B97F4A7C15ULL; z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 32;
  return z;
}
template <int MODE>
__device__ __forceinline__ uint32_t ld(const uint32_t* p) {
  uint32_t v;
  if (MODE == 0) asm("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (MODE == 1) asm("ld.global.nc.L2::128B.u32 %0, [%1];" : "=r"(v) : "l"(p));
  else asm("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
template <int MODE>
__global__ void __launch_bounds__(256) k_gather8(const uint32_t* __restrict__ tab, uint64_t wmask, uint64_t n8, uint64_t salt,
                                                 unsigned long long* sink) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n8) return;
  uint32_t v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = ld<MODE>(tab + (mix((t * 8 + k) ^ salt) & wmask));
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc += v[k];
  if (acc == 0x12345678u) atomicAdd(sink, 1ULL);
}
int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 35;
  const uint64_t bytes = 1ULL << lg;
  uint32_t* tab; unsigned long long* sink;
  if (cudaMalloc(&tab, bytes) != cudaSuccess) { printf("cannot allocate 2^%d bytes\n", lg); return 1; }
  cudaMemset(tab, 1, bytes); cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
  const uint64_t n8 = 1ULL << 25;   // threads: 2^28 gathers per launch
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("table 2^%d bytes, %llu gathers per launch, 8 independent reads per thread\n", lg, (unsigned long long)(n8 * 8));
  const char* names[] = {"ld.global.nc.L2::64B", "ld.global.nc.L2::128B", "ld.global.nc"};
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      const unsigned grid = (unsigned)((n8 + 255) / 256);
      if (mode == 0) k_gather8<0><<<grid, 256>>>(tab, bytes / 4 - 1, n8, rep * 7919 + 1, sink);
      else if (mode == 1) k_gather8<1><<<grid, 256>>>(tab, bytes / 4 - 1, n8, rep * 7919 + 1, sink);
      else k_gather8<2><<<grid, 256>>>(tab, bytes / 4 - 1, n8, rep * 7919 + 1, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    printf("  %-22s %.3f ms  %.1f G gathers/s\n", names[mode], best, n8 * 8 / best / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
