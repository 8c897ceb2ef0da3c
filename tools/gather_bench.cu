// gather_bench.cu -- the access pattern of the hot path in isolation: every thread reads one random,
// aligned element (4 / 16 / 32 bytes) of a table much larger than L2.  Gives the gather roofline the
// k_search kernels are judged against (DESIGN.md "Roofline"): elements/s and, under ncu
// (--metrics dram__bytes_read.sum), the DRAM bytes the chip really moves per element.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
//   ./gather_bench [table GiB = 8] [l2 fetch granularity = 0 (leave default)]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL; z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
  return z;
}
// the same 4-byte gather with an explicit PTX L2 prefetch-size hint / eviction policy
template <int MODE>
__global__ void k_gather_hint(const uint8_t* __restrict__ tab, uint64_t nelem, uint64_t n, uint32_t salt, unsigned long long* sink) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (uint64_t i = t; i < n; i += stride) {
    uint64_t e = mix(i * 0x100000001B3ULL + salt) % nelem;
    const uint8_t* p = tab + e * 4;
    uint32_t v;
    if (MODE == 0) asm volatile("ld.global.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.L2::128B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.L2::256B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 4) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    acc += v;
  }
  if (acc == 0x123456789ULL) atomicAdd(sink, 1ULL);
}
template <int BYTES>
__global__ void k_gather(const uint8_t* __restrict__ tab, uint64_t nelem, uint64_t n, uint32_t salt, unsigned long long* sink) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (uint64_t i = t; i < n; i += stride) {
    uint64_t e = mix(i * 0x100000001B3ULL + salt) % nelem;
    const uint8_t* p = tab + e * BYTES;
    if (BYTES == 4) acc += __ldg((const uint32_t*)p);
    else if (BYTES == 16) { uint4 v = __ldg((const uint4*)p); acc += v.x ^ v.w; }
    else { uint4 a = __ldg((const uint4*)p), b = __ldg((const uint4*)p + 1); acc += a.x ^ b.w; }
  }
  if (acc == 0x123456789ULL) atomicAdd(sink, 1ULL);
}
int main(int argc, char** argv) {
  double gib = argc > 1 ? atof(argv[1]) : 8.0;
  int gran = argc > 2 ? atoi(argv[2]) : 0;
  if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
  size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
  uint64_t bytes = (uint64_t)(gib * (1ULL << 30));
  uint8_t* tab; unsigned long long* sink;
  cudaMalloc(&tab, bytes); cudaMemset(tab, 1, bytes); cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
  uint64_t n = 1ULL << 28;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("table %.1f GiB, %llu gathers per launch, L2 fetch granularity limit %zu\n", gib, (unsigned long long)n, g);
  for (int bytes_per : {4, 16, 32}) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (bytes_per == 4) k_gather<4><<<148 * 8, 256>>>(tab, bytes / 4, n, rep, sink);
      else if (bytes_per == 16) k_gather<16><<<148 * 8, 256>>>(tab, bytes / 16, n, rep, sink);
      else k_gather<32><<<148 * 8, 256>>>(tab, bytes / 32, n, rep, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("  %2d-byte elements: %.3f ms  %.1f G gathers/s  (%.0f GB/s of 32-byte sectors)\n", bytes_per, ms, n / ms / 1e6, n * 32.0 / ms / 1e6);
    }
  }
  const char* names[] = {"L2::64B", "L2::128B", "L2::256B", "L1::no_allocate", "cv", "cg"};
  for (int mode = 0; mode < 6; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      switch (mode) {
        case 0: k_gather_hint<0><<<148 * 8, 256>>>(tab, bytes / 4, n, rep, sink); break;
        case 1: k_gather_hint<1><<<148 * 8, 256>>>(tab, bytes / 4, n, rep, sink); break;
        case 2: k_gather_hint<2><<<148 * 8, 256>>>(tab, bytes / 4, n, rep, sink); break;
        case 3: k_gather_hint<3><<<148 * 8, 256>>>(tab, bytes / 4, n, rep, sink); break;
        case 4: k_gather_hint<4><<<148 * 8, 256>>>(tab, bytes / 4, n, rep, sink); break;
        default: k_gather_hint<5><<<148 * 8, 256>>>(tab, bytes / 4, n, rep, sink); break;
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("  4-byte, ld.global.%s: %.3f ms  %.1f G gathers/s\n", names[mode], ms, n / ms / 1e6);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
