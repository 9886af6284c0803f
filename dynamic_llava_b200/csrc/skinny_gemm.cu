// Weight-streaming "skinny" GEMM for the decode step:  Y[M, N] = X[M, K] . W[N, K]^T  with M = batch <= 64.
// (every nn.Linear of the decoder at q_len == 1: dynamic_modeling_llama.py:1011-1013, :1127, :328, :2709.)
//
// At M <= 64 the op is bound by reading W once from HBM (13.5 GB per 7B decode step), so the kernel is built as a
// streaming engine: a TMA producer warp pulls 128-row x 64-col W tiles (16 KB, 128B-swizzled) plus the matching X tile
// through a 5-stage mbarrier ring; ONE elected thread issues tcgen05.mma (A = W tile from smem, B = X tile from smem,
// D = 128 x NB fp32 accumulator in TMEM, NB = batch padded to 16/32/64); four epilogue warps read TMEM with tcgen05.ld.
// No weight byte ever passes through registers or the LSU.  Split-K across blockIdx.y fills the 148 SMs for the
// narrow-N projections; every split writes its own fp32 slice partial[split][m][n] (no atomics -> deterministic), and
// the fused consumer kernels in decode_fused.cu sum the slices in fixed order.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#ifdef DL_HALF
#define DL_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#else
#define DL_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#endif
#include "../../include/dynllava_b200.h"

#define SG_BM 128
#define SG_BK 64
#define SG_STAGES 5  // 6 stages (110 KB, still 2 CTAs/SM) measured no faster: profiles/r1_microbench_v8_skinny_6stage.json
#define SG_THREADS 192  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue

template <int NB>
struct SgSmem {
  alignas(1024) bf16 a[SG_STAGES][SG_BM * SG_BK];
  alignas(1024) bf16 b[SG_STAGES][(NB < 8 ? 8 : NB) * SG_BK];
  alignas(8) uint64_t full[SG_STAGES];
  alignas(8) uint64_t empty[SG_STAGES];
  alignas(8) uint64_t tmem_full;
  uint32_t tmem_base;
};

// EPI == 0: fp32 split-K slices.  EPI == 1 (gate|up projection, splits == 1): each 128-row A tile is assembled from TWO
// 64-row TMA boxes — gate rows [64j, 64j+64) and the matching up rows [I+64j, I+64j+64) of the same fused [2I, K] weight the
// prefill uses — so one tile holds both halves of 64 SwiGLU outputs; the epilogue exchanges the `up` half through shared
// memory and writes act = bf16(bf16(silu(bf16 g)) * bf16 u) directly (dynamic_modeling_llama.py:328): no slice round trip
// and no separate SwiGLU kernel in the decode step.
template <int NB, int EPI>
__global__ void __launch_bounds__(SG_THREADS) skinny_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                                 const __grid_constant__ CUtensorMap tmap_x,
                                                                 float* __restrict__ partial, int M, int N, int KB,
                                                                 int kb_per_split, int l2_pf, bf16* __restrict__ act_out) {
  constexpr int TM_COLS = NB < 32 ? 32 : NB;
  extern __shared__ uint8_t smem_raw[];
  SgSmem<NB>& S = *reinterpret_cast<SgSmem<NB>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * SG_BM;  // EPI == 1: tile j = blockIdx.x covers SwiGLU outputs [64j, 64j+64)
  const int split = blockIdx.y;
  // balanced K partition: split s owns k-blocks [s*KB/S, (s+1)*KB/S) (sizes differ by at most one block)
  const int n_splits = gridDim.y;
  const int kb0 = (int)(((int64_t)split * KB) / n_splits);
  const int kb1 = (int)(((int64_t)(split + 1) * KB) / n_splits);
  const int nkb = kb1 - kb0;
  (void)kb_per_split;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < SG_STAGES; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1);
    }
    mbar_init(&S.tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TM_COLS>(&S.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = S.tmem_base;

  if (warp == 0) {
    if (elect_one()) {  // converged warp + elect.sync keeps the UTCHMMA/UTMALDG issue on the uniform datapath
      constexpr uint32_t BYTES = (SG_BM + NB) * SG_BK * 2;
      // The weights do not depend on the previous kernel: start streaming W into every free stage right away, and only
      // then wait for the predecessor (PDL) before fetching the activation tiles X it produced.
      const int npre = min(nkb, SG_STAGES);
      const int I_rows = N >> 1, g0 = blockIdx.x * 64;
      auto load_w = [&](int stage, int kb) {
        if constexpr (EPI == 1) {  // tmap_w box = 64 rows: gate half then up half (8 KB apart keeps the 128B swizzle phase)
          tma_load_2d(S.a[stage], &tmap_w, kb * SG_BK, g0, &S.full[stage]);
          tma_load_2d(S.a[stage] + 64 * SG_BK, &tmap_w, kb * SG_BK, I_rows + g0, &S.full[stage]);
        } else {
          tma_load_2d(S.a[stage], &tmap_w, kb * SG_BK, n0, &S.full[stage]);
        }
      };
      for (int i = 0; i < npre; ++i) {
        mbar_expect_tx(&S.full[i], BYTES);
        load_w(i, kb0 + i);
      }
      // ...and ask the TMA unit to pull the next `l2_pf` W tiles of this CTA into the 126 MB L2 (no smem needed), so HBM
      // keeps streaming weights while the small predecessor kernel (norm / SwiGLU / attention) is still running.
      // l2_pf > 0: ahead of the wait (measured negative in round 1: competes with the predecessor's own stream);
      // l2_pf < 0: right AFTER the wait, i.e. exactly when the predecessor has finished and HBM would otherwise idle until the
      // first X tile arrives and the MMAs start freeing ring stages (round 2: also measured neutral, 2.73-2.78 ms per decode step for
      // 0 / 12 / 24 / 48 MB, profiles/r2_l2_prefetch_after_wait.md)
      const int npf = min(nkb, npre + (l2_pf > 0 ? l2_pf : -l2_pf));
      if (l2_pf > 0)
        for (int i = npre; i < npf; ++i) tma_prefetch_2d(&tmap_w, (kb0 + i) * SG_BK, EPI == 1 ? g0 : n0);
      pdl_launch_dependents();
      pdl_wait();
      if (l2_pf < 0) {
        for (int i = npre; i < npf; ++i) {
          tma_prefetch_2d(&tmap_w, (kb0 + i) * SG_BK, EPI == 1 ? g0 : n0);
          if constexpr (EPI == 1) tma_prefetch_2d(&tmap_w, (kb0 + i) * SG_BK, I_rows + g0);
        }
      }
      for (int i = 0; i < npre; ++i) tma_load_2d(S.b[i], &tmap_x, (kb0 + i) * SG_BK, 0, &S.full[i]);
      for (int i = npre; i < nkb; ++i) {
        const int s = i % SG_STAGES;
        const uint32_t ph = (i / SG_STAGES) & 1;
        mbar_wait(&S.empty[s], ph ^ 1);
        mbar_expect_tx(&S.full[s], BYTES);
        load_w(s, kb0 + i);
        tma_load_2d(S.b[s], &tmap_x, (kb0 + i) * SG_BK, 0, &S.full[s]);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // converged warp + elect.sync keeps the UTCHMMA/UTMALDG issue on the uniform datapath
      constexpr uint32_t IDESC = umma_idesc_bf16(SG_BM, NB);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % SG_STAGES;
        const uint32_t ph = (i / SG_STAGES) & 1;
        mbar_wait(&S.full[s], ph);
        tc_fence_after();
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(S.a[s]));
        const uint64_t db = umma_desc_kmajor_sw128(smem_u32(S.b[s]));
#pragma unroll
        for (int k = 0; k < SG_BK / 16; ++k)  // +32 B per K=16 step inside the 128 B swizzle atom (encoded >> 4)
          umma_bf16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&S.empty[s]);  // smem slot reusable once these MMAs have read it
      }
      umma_commit(&S.tmem_full);
    }
  } else {
    // epilogue: TMEM lane = W row inside the tile, TMEM column = batch row m
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int n = n0 + q * 32 + lane;
    float* dst = partial + (int64_t)split * M * N + n;
    if constexpr (EPI == 1) {
      float* xch = reinterpret_cast<float*>(S.a[0]);  // [64 rows][NB] floats; every stage has been consumed by now
      mbar_wait(&S.tmem_full, 0);
      tc_fence_after();
      uint32_t r[NB];
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += 16) tmem_ld_32x32b_x16(tmem_d + ((uint32_t)(q * 32) << 16) + c0, r + c0);
      tmem_ld_wait();
      const int row = (q & 1) * 32 + lane;  // index inside the 64-row half
      if (q >= 2) {
#pragma unroll
        for (int m = 0; m < NB; ++m) xch[row * NB + m] = __uint_as_float(r[m]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
      if (q < 2) {
        const int I = N >> 1;
        const int col = blockIdx.x * 64 + row;
        if (col < I) {
#pragma unroll
          for (int m = 0; m < NB; ++m) {
            if (m < M) {
              const float g = rbf(__uint_as_float(r[m])), u = rbf(xch[row * NB + m]);
              act_out[(int64_t)m * I + col] = f2bf(rbf(g / (1.0f + expf(-g))) * u);
            }
          }
        }
      }
    } else if (nkb > 0) {
      mbar_wait(&S.tmem_full, 0);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_d + ((uint32_t)(q * 32) << 16) + c0, r);
        tmem_ld_wait();
        if (n < N) {
#pragma unroll
          for (int m = 0; m < 16; ++m)
            if (c0 + m < M) dst[(int64_t)(c0 + m) * N] = __uint_as_float(r[m]);
        }
      }
    } else if (n < N) {
      for (int m = 0; m < M; ++m) dst[(int64_t)m * N] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TM_COLS>(tmem_d);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point resolved at run time: the .so must load on a CPU-only box)
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128-byte swizzle.
// Encoding costs ~1-2 us of host time per map and a prefill issues ~700 TMA launches over buffers whose addresses repeat (torch's
// caching allocator, the session's fixed decode buffers, the weights): a small per-thread direct-mapped cache keyed by every
// parameter of the encoding returns the finished 128-byte descriptor instead (a descriptor is a pure function of its key).
struct TmapKey {
  const void* base;
  uint64_t rows, cols, stride;
  uint32_t box_rows, box_cols;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && stride == o.stride && box_rows == o.box_rows && box_cols == o.box_cols;
  }
};
struct TmapEntry {
  TmapKey key;
  bool valid;
  alignas(64) CUtensorMap map;
};
#define TMAP_CACHE_SLOTS 2048
static thread_local TmapEntry* g_tmap_cache = nullptr;

extern "C" int dl_make_tmap_2d(void* out_map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                               uint32_t box_rows, uint32_t box_cols) {
  if (g_tmap_cache == nullptr) g_tmap_cache = (TmapEntry*)calloc(TMAP_CACHE_SLOTS, sizeof(TmapEntry));
  const TmapKey key = {base, rows, cols, row_stride_elems, box_rows, box_cols};
  uint64_t h = (uint64_t)(uintptr_t)base * 0x9E3779B97F4A7C15ull;
  h ^= (rows * 0xC2B2AE3D27D4EB4Full) ^ (cols * 0x165667B19E3779F9ull) ^ ((uint64_t)box_rows << 40) ^ ((uint64_t)box_cols << 20) ^ row_stride_elems;
  h ^= h >> 29;
  TmapEntry* e = g_tmap_cache ? &g_tmap_cache[h % TMAP_CACHE_SLOTS] : nullptr;
  if (e != nullptr && e->valid && e->key == key) {
    memcpy(out_map, &e->map, sizeof(CUtensorMap));
    return 0;
  }
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (enc == nullptr) return dl_set_error(-2, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc((CUtensorMap*)out_map, DL_TMAP_DTYPE, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return dl_set_error((int)r, "cuTensorMapEncodeTiled");
  if (e != nullptr) {
    e->key = key;
    memcpy(&e->map, out_map, sizeof(CUtensorMap));
    e->valid = true;
  }
  return 0;
}

template <int NB, int EPI>
static int launch_sg(const CUtensorMap& tw, const CUtensorMap& tx, float* partial, int M, int N, int KB, int splits, int kbps,
                     int l2_prefetch_mb, bf16* act_out, cudaStream_t stream) {
  const int smem = (int)sizeof(SgSmem<NB>) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(skinny_gemm_kernel<NB, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_skinny_gemm(smem attr)");
    attr_set = true;
  }
  dim3 grid((N + SG_BM - 1) / SG_BM, splits);
  // l2_prefetch_mb: 0 / -1 = off; > 0 = MB per launch prefetched into L2 BEFORE the dependency wait; < -1 = |mb| MB AFTER the wait
  const int64_t mb = l2_prefetch_mb == -1 ? 0 : l2_prefetch_mb;
  const int64_t budget = (mb < 0 ? -mb : mb) << 20;
  int l2_pf = (int)(budget / ((int64_t)grid.x * grid.y * SG_BM * SG_BK * 2));
  if (mb < 0) l2_pf = -l2_pf;
  cudaError_t le = dl_launch_pdl(skinny_gemm_kernel<NB, EPI>, grid, dim3(SG_THREADS), (size_t)smem, stream, tw, tx, partial, M, N, KB, kbps,
                                 l2_pf, act_out);
  if (le != cudaSuccess) return dl_set_error((int)le, "dl_skinny_gemm(launch)");
  DL_CHECK_LAUNCH("dl_skinny_gemm");
}

extern "C" int dl_skinny_gemm_splits(int N, int K, int target_ctas) {
  const int tiles = (N + SG_BM - 1) / SG_BM;
  const int KB = K / SG_BK;
  int s = (target_ctas + tiles / 2) / tiles;
  if (tiles >= 148) s = 1;  // one full wave of row tiles already: splitting K only adds slices (measured: profiles/r1_microbench_v4_skinny_splits.json)
  if (s < 1) s = 1;
  int max_s = KB / 4;  // keep >= 4 k-blocks (64 KB of W) per CTA
  if (max_s < 1) max_s = 1;
  if (s > max_s) s = max_s;
  if (s > 8) s = 8;  // consumers fetch the slices in one 8-deep group of loads
  return s;
}

extern "C" int dl_skinny_gemm(const void* x, const void* w, float* partial, int M, int N, int K, int splits, int l2_prefetch_mb,
                              cudaStream_t stream) {
  DL_REQUIRE(M >= 1 && M <= 64 && K % SG_BK == 0 && N % 8 == 0 && splits >= 1, "dl_skinny_gemm");
  const int KB = K / SG_BK;
  const int kbps = (KB + splits - 1) / splits;
  DL_REQUIRE(splits <= KB, "dl_skinny_gemm(empty split)");
  const int NB = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  alignas(64) CUtensorMap tw, tx;
  int rc = dl_make_tmap_2d(&tw, w, (uint64_t)N, (uint64_t)K, (uint64_t)K, SG_BM, SG_BK);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tx, x, (uint64_t)M, (uint64_t)K, (uint64_t)K, (uint32_t)NB, SG_BK);
  if (rc) return rc;
  if (NB == 16) return launch_sg<16, 0>(tw, tx, partial, M, N, KB, splits, kbps, l2_prefetch_mb, nullptr, stream);
  if (NB == 32) return launch_sg<32, 0>(tw, tx, partial, M, N, KB, splits, kbps, l2_prefetch_mb, nullptr, stream);
  return launch_sg<64, 0>(tw, tx, partial, M, N, KB, splits, kbps, l2_prefetch_mb, nullptr, stream);
}

// gate|up projection + SwiGLU in one launch.  w_gate_up [2I, K] = [gate rows | up rows]; act_out bf16 [M, I].
extern "C" int dl_skinny_gemm_swiglu(const void* x, const void* w_gate_up, void* act_out, int M, int I, int K,
                                     cudaStream_t stream) {
  DL_REQUIRE(M >= 1 && M <= 64 && K % SG_BK == 0 && I % 64 == 0, "dl_skinny_gemm_swiglu");
  const int N = 2 * I, KB = K / SG_BK;
  const int NB = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  alignas(64) CUtensorMap tw, tx;
  int rc = dl_make_tmap_2d(&tw, w_gate_up, (uint64_t)N, (uint64_t)K, (uint64_t)K, 64, SG_BK);  // 64-row boxes
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tx, x, (uint64_t)M, (uint64_t)K, (uint64_t)K, (uint32_t)NB, SG_BK);
  if (rc) return rc;
  if (NB == 16) return launch_sg<16, 1>(tw, tx, nullptr, M, N, KB, 1, KB, 0, (bf16*)act_out, stream);
  if (NB == 32) return launch_sg<32, 1>(tw, tx, nullptr, M, N, KB, 1, KB, 0, (bf16*)act_out, stream);
  return launch_sg<64, 1>(tw, tx, nullptr, M, N, KB, 1, KB, 0, (bf16*)act_out, stream);
}
