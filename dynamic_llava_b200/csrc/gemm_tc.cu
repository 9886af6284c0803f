// K6: dense bf16 GEMM on tcgen05 tiles with a fused bias + activation epilogue:
//     out[M, N] = act( X[M, K] . W[N, K]^T + bias[N] )        (nn.Linear layout, fp32 accumulate, one bf16 rounding)
// Used for the CLIP patch embedding (HF CLIPVisionEmbeddings: Conv2d(3 -> C, k = s = 14, bias=False) == im2col GEMM
// [B*576, 588->640] x [640, 1024]) and available for the projector / predictor linears
// (multimodal_projector/builder.py:172-179).  The big decoder projections of the prefill stay on cuBLASLt (library GEMM).
//
// 128 x 128 output tile per CTA, K streamed in 64-column stages: warp 0 = TMA producer (X tile + W tile, 128B swizzle,
// 3-stage mbarrier ring), warp 1 = one elected thread issuing tcgen05.mma M=128 N=128 K=16 with the fp32 accumulator in
// TMEM (128 columns), warps 2-5 = epilogue: tcgen05.ld 64 columns -> +bias -> activation -> bf16 -> 32-row x 64-column box in
// shared memory (128B swizzle) -> one TMA store per box.  (A row-per-thread store straight from registers is 32 scattered
// 16-byte requests per instruction; the same change took the attention epilogue from 16 k to 2.6 k cycles.)  97 KB of shared
// memory (the staging boxes reuse pipeline stage 0) and 128 TMEM columns per CTA: two CTAs share an SM, one's epilogue runs under the other's main loop.
#include <cuda.h>

#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define GT_BM 128
#define GT_BN 128
#define GT_BK 64
#define GT_STAGES 3
#define GT_THREADS 192

struct GtSmem {
  alignas(1024) bf16 a[GT_STAGES][GT_BM * GT_BK];
  alignas(1024) bf16 b[GT_STAGES][GT_BN * GT_BK];
  alignas(8) uint64_t full[GT_STAGES];
  alignas(8) uint64_t empty[GT_STAGES];
  alignas(8) uint64_t tmem_full;
  uint32_t tmem_base;
};

DL_DEVINL void gt_tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(smem_u32(smem_src)),
               "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __launch_bounds__(GT_THREADS, 2) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                const __grid_constant__ CUtensorMap tmap_w,
                                                                const __grid_constant__ CUtensorMap tmap_o,
                                                                const bf16* __restrict__ bias, int M, int N, int KB, int act) {
  extern __shared__ uint8_t smem_raw[];
  GtSmem& S = *reinterpret_cast<GtSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * GT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_o);
    for (int s = 0; s < GT_STAGES; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1);
    }
    mbar_init(&S.tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(&S.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = S.tmem_base;

  if (warp == 0) {
    if (elect_one()) {  // converged warp + elect.sync keeps the UTCHMMA/UTMALDG issue on the uniform datapath
      constexpr uint32_t BYTES = (GT_BM + GT_BN) * GT_BK * 2;
      for (int i = 0; i < KB; ++i) {
        const int s = i % GT_STAGES;
        mbar_wait(&S.empty[s], ((i / GT_STAGES) & 1) ^ 1);
        mbar_expect_tx(&S.full[s], BYTES);
        tma_load_2d(S.a[s], &tmap_x, i * GT_BK, m0, &S.full[s]);
        tma_load_2d(S.b[s], &tmap_w, i * GT_BK, n0, &S.full[s]);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // converged warp + elect.sync keeps the UTCHMMA/UTMALDG issue on the uniform datapath
      constexpr uint32_t IDESC = umma_idesc_bf16(GT_BM, GT_BN);
      for (int i = 0; i < KB; ++i) {
        const int s = i % GT_STAGES;
        mbar_wait(&S.full[s], (i / GT_STAGES) & 1);
        tc_fence_after();
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(S.a[s]));
        const uint64_t db = umma_desc_kmajor_sw128(smem_u32(S.b[s]));
#pragma unroll
        for (int k = 0; k < GT_BK / 16; ++k)
          umma_bf16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&S.empty[s]);
      }
      umma_commit(&S.tmem_full);
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter of this warp = rows [32q, 32q+32) of the tile
    // staging tile: 4 KB of pipeline stage 0 (every stage has been consumed once tmem_full fires)
    uint8_t* stage = reinterpret_cast<uint8_t*>(S.a[0]) + q * 4096;
    mbar_wait(&S.tmem_full, 0);
    tc_fence_after();
#pragma unroll
    for (int hh = 0; hh < GT_BN / 64; ++hh) {
      uint32_t r[64];
      tmem_ld_32x32b_x32(tmem_d + ((uint32_t)(q * 32) << 16) + hh * 64, r);
      tmem_ld_32x32b_x32(tmem_d + ((uint32_t)(q * 32) << 16) + hh * 64 + 32, r + 32);
      tmem_ld_wait();
      bf16x8 pk[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = n0 + hh * 64 + c * 8;
        float f[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(r[c * 8 + u]);
        if (bias != nullptr && col < N) {  // N % 8 == 0
          float bv[8];
          unpack8(ld8(bias + col), bv);
#pragma unroll
          for (int u = 0; u < 8; ++u) f[u] += bv[u];
        }
        if (act == 1) {  // nn.GELU (erf) applied to the bf16-rounded linear output, as the reference's separate modules do
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float y = rbf(f[u]);
            f[u] = 0.5f * y * (1.0f + erff(y * 0.70710678118654752440f));
          }
        }
        pk[c] = pack8(f);
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous box has left the staging tile
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<bf16x8*>(stage + lane * 128 + ((c ^ (lane & 7)) << 4)) = pk[c];
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && m0 + q * 32 < M && n0 + hh * 64 < N) {  // the tensor map clips rows >= M and columns >= N
        gt_tma_store_2d(&tmap_o, stage, n0 + hh * 64, m0 + q * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_d);
}

extern "C" int dl_gemm_bias_act(const void* x, const void* w, const void* bias, void* out, int M, int N, int K, int act,
                                cudaStream_t stream) {
  DL_REQUIRE(M >= 1 && N % 8 == 0 && K % GT_BK == 0 && (act == 0 || act == 1), "dl_gemm_bias_act");
  alignas(64) CUtensorMap tx, tw;
  int rc = dl_make_tmap_2d(&tx, x, (uint64_t)M, (uint64_t)K, (uint64_t)K, GT_BM, GT_BK);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tw, w, (uint64_t)N, (uint64_t)K, (uint64_t)K, GT_BN, GT_BK);
  if (rc) return rc;
  alignas(64) CUtensorMap to;  // out [M, N], boxes of 32 rows x 64 columns
  rc = dl_make_tmap_2d(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)N, 32, 64);
  if (rc) return rc;
  const int smem = (int)sizeof(GtSmem) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_gemm_bias_act(smem attr)");
    attr_set = true;
  }
  dim3 grid((N + GT_BN - 1) / GT_BN, (M + GT_BM - 1) / GT_BM);
  gemm_tc_kernel<<<grid, GT_THREADS, smem, stream>>>(tx, tw, to, (const bf16*)bias, M, N, K / GT_BK, act);
  DL_CHECK_LAUNCH("dl_gemm_bias_act");
}
