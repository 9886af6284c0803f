#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int COLS>
__global__ void __launch_bounds__(192) kt(float* p) {
  extern __shared__ float s[];
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  s[threadIdx.x] = 1.f;
  // hold the SM for a while so that co-residency shows up in %nsmid / timing
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {}
  if (p) atomicAdd(p, 1.f);
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(COLS) : "memory");
}
template <int COLS>
void probe(const char* name) {
  const int smem = 98304;
  cudaFuncSetAttribute(kt<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int n = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kt<COLS>, 192, smem);
  float* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // 296 CTAs each spinning 2M cycles: one wave if 2 CTAs/SM are really co-resident, two waves otherwise
  kt<COLS><<<296, 192, smem>>>(d); cudaDeviceSynchronize();
  cudaEventRecord(e0); kt<COLS><<<296, 192, smem>>>(d); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaEventRecord(e0); kt<COLS><<<148, 192, smem>>>(d); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms1; cudaEventElapsedTime(&ms1, e0, e1);
  printf("%s: occupancy API %d; 296 CTAs %.3f ms vs 148 CTAs %.3f ms (%s)\n", name, n, ms, ms1, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  probe<32>("tcgen05.alloc 32 cols");
  probe<128>("tcgen05.alloc 128 cols");
  probe<256>("tcgen05.alloc 256 cols");
  return 0;
}
