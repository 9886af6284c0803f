// Hardware-floor microbenchmarks that the attention kernel design depends on (one CTA per SM, clock64 deltas):
//   mufu:  ex2.approx.ftz.f32 throughput per SM with 4 / 8 warps
//   umma:  back-to-back tcgen05.mma (kind::f16, M=128) cycles per instruction for SS (A,B in smem) and TS (A in TMEM)
//          at N = 64 / 128 / 256
//   ldtm:  tcgen05.ld 32x32b.x32 cycles per instruction, 4 warps (one per lane quarter)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o sm_floor sm_floor.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "../../dynamic_llava_b200/csrc/common.cuh"

extern "C" int dl_set_error(int code, const char*) { return code; }

__global__ void mufu_kernel(float* out, long long* cyc, int iters) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void fma_kernel(float* out, long long* cyc, int iters) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

DL_DEVINL void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// MODE 0: SS, MODE 1: TS.  N template.  `n_mma` MMAs back to back, then one commit.  smem content is garbage (timing only).
template <int N, int MODE>
__global__ void __launch_bounds__(128) umma_kernel(long long* cyc, int n_mma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  // zero the operand tiles so that no NaN slow path can matter
  for (int i = threadIdx.x; i < (128 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  fence_proxy_async();
  if (warp == 0) tmem_alloc<512>(&tbase);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tbase;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t ID = umma_idesc_bf16(128, N, 0, 0);
      const uint32_t a = smem_u32(sm), b = smem_u32(sm) + 32768;
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t o = (kk >> 2) * 16384 + (kk & 3) * 32;
          if (MODE == 0)
            umma_bf16_ss(tmem, umma_desc_kmajor_sw128(a + o), umma_desc_kmajor_sw128(b + o), ID, 1u);
          else
            umma_ts(tmem, tmem + 256 + kk * 8, umma_desc_kmajor_sw128(b + o), ID, 1u);
        }
      }
      umma_commit(&bar);
      const long long t1 = clock64();
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      cyc[blockIdx.x * 2] = t1 - t0;
      cyc[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

__global__ void __launch_bounds__(256) ldtm_kernel(long long* cyc, float* out, int iters) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tbase);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32];
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    tmem_ld_32x32b_x32(t + (it & 7) * 32, r);
    tmem_ld_wait();
    acc += __uint_as_float(r[it & 31]);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int N, int MODE>
void run_umma(long long* d_cyc, const char* name) {
  const int smem = 128 * 1024 + 1024;
  CK(cudaFuncSetAttribute(umma_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int n_mma : {64, 512}) {
    umma_kernel<N, MODE><<<1, 128, smem>>>(d_cyc, n_mma);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"umma\", \"mode\": \"%s\", \"N\": %d, \"n_mma\": %d, \"issue_cycles_per_mma\": %.1f, \"complete_cycles_per_mma\": %.1f, \"floor\": %.1f}\n",
           name, N, n_mma, (double)h[0] / n_mma, (double)h[1] / n_mma, 128.0 * N / 256.0);
  }
}

int main() {
  long long* d_cyc;
  float* d_out;
  CK(cudaMalloc(&d_cyc, 4096 * sizeof(long long)));
  CK(cudaMalloc(&d_out, 1 << 22));
  for (int warps : {4, 8, 16}) {
    const int iters = 256;
    mufu_kernel<<<1, warps * 32>>>(d_out, d_cyc, iters);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"mufu_ex2\", \"warps\": %d, \"ex2_per_clk_per_sm\": %.2f}\n", warps, (double)warps * 32 * 32 * iters / h);
    fma_kernel<<<1, warps * 32>>>(d_out, d_cyc, iters);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"ffma\", \"warps\": %d, \"fma_per_clk_per_sm\": %.2f}\n", warps, (double)warps * 32 * 32 * iters / h);
  }
  run_umma<64, 0>(d_cyc, "SS");
  run_umma<128, 0>(d_cyc, "SS");
  run_umma<256, 0>(d_cyc, "SS");
  run_umma<64, 1>(d_cyc, "TS");
  run_umma<128, 1>(d_cyc, "TS");
  run_umma<256, 1>(d_cyc, "TS");
  for (int warps : {4, 8}) {
    ldtm_kernel<<<1, warps * 32>>>(d_cyc, d_out, 1024);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"ldtm_x32\", \"warps\": %d, \"cycles_per_ld_serial\": %.1f}\n", warps, (double)h / 1024);
  }
  return 0;
}
