// Decode-step GEMMs with the neighbouring small ops folded in (q_len == 1, M = batch <= 64): every nn.Linear of a decoder layer
// streams its weight exactly once (TMA -> 5-stage smem ring -> tcgen05.mma with the accumulator in TMEM, as in skinny_gemm.cu) and
// the ops that used to be separate launches between two GEMMs live in the GEMM's prologue / epilogue:
//
//   PROLOGUE  DG_PRO_TMA   X tile by TMA (the attention output, the SwiGLU activations)
//             DG_PRO_NORM  RMSNorm ON LOAD (dynamic_modeling_llama.py:134-139): one warp builds every X tile in shared memory from the bf16
//                          residual stream x, the per-row sum of squares the previous GEMM's epilogue left as partials, and the norm
//                          weight: bf16(x * rstd) * w, written in the 128B-swizzled K-major layout the MMA descriptor expects
//   EPILOGUE  DG_EPI_RESID   split-K across a THREAD-BLOCK CLUSTER (cluster = the splits of one 128-row tile): partial accumulators
//                            meet in distributed shared memory, each rank sums a slice of the rows in fixed order (deterministic),
//                            x = bf16(x + bf16(sum)) (:1289/:1295) and the row's partial sum of squares for the next RMSNorm
//             DG_EPI_QKV     same cluster reduce, then RoPE on q and k (:253-285, reference rounding), q to a [B, H] buffer, k / v
//                            straight into the KV arena at slot kv_len[b] (whether the slot stays is decided by dl_kv_advance)
//             DG_EPI_SWIGLU  gate|up tile assembled from two 64-row boxes, act = bf16(bf16(silu(g)) * u) (:328)
//             DG_EPI_LOGITS  bf16 logits (lm_head, :2709)
//
// so a decoder layer is 5 launches (QKV -> attention -> O -> gate|up -> down) instead of 7, the split-K slices never travel through
// global memory, and the attention kernel no longer sums slices / rotates / appends before its first QK^T.
#include <cuda.h>

#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define DG_BM 128
#define DG_BK 64
#define DG_STAGES 5
#define DG_THREADS 192  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue (and, during the main loop, norm-on-load)
#define DG_MAX_NORM_K 5120  // widest hidden size a norm-on-load CTA keeps the norm weight of (13B: 5120)

enum { DG_PRO_TMA = 0, DG_PRO_NORM = 1 };
enum { DG_EPI_RESID = 0, DG_EPI_QKV = 1, DG_EPI_SWIGLU = 2, DG_EPI_LOGITS = 3 };

struct DgParams {
  // norm-on-load prologue
  const bf16* x;       // residual stream [M, K]
  const float* ssq;    // [M, n_parts] partial sums of squares of the rows of x
  int n_parts;
  const bf16* norm_w;  // [K]
  float eps;
  int K;
  // problem
  int M, N, KB;
  // DG_EPI_RESID
  bf16* xres;      // [M, N], updated in place
  float* ssq_out;  // [M, tiles * splits]
  // DG_EPI_QKV
  bf16* q_out;  // [M, heads * 128]
  bf16* k_arena;
  bf16* v_arena;
  const int32_t* kv_len;  // [M] slot of the new token
  const int32_t* pos;     // [M] RoPE position
  const bf16* cos_t;
  const bf16* sin_t;
  int heads, cap;
  // DG_EPI_SWIGLU / DG_EPI_LOGITS
  bf16* out;
};

template <int NB>
struct DgSmem {
  alignas(1024) bf16 a[DG_STAGES][DG_BM * DG_BK];
  alignas(1024) bf16 b[DG_STAGES][(NB < 8 ? 8 : NB) * DG_BK];
  alignas(16) bf16 wn[DG_MAX_NORM_K];  // norm-on-load: this CTA's K-range of the RMSNorm weight
  float rstd[64];  // norm-on-load: 1 / sqrt(mean(x^2) + eps) per batch row
  alignas(8) uint64_t full[DG_STAGES];
  alignas(8) uint64_t empty[DG_STAGES];
  alignas(8) uint64_t xfull[DG_STAGES];  // norm-on-load: the raw x tile of a stage has landed
  alignas(8) uint64_t wn_full;
  alignas(8) uint64_t tmem_full;
  uint32_t tmem_base;
};

DL_DEVINL void dg_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
DL_DEVINL float dg_ld_dsmem(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

template <int NB, int PRO, int EPI>
__global__ void __launch_bounds__(DG_THREADS) dec_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                              const __grid_constant__ CUtensorMap tmap_x, const DgParams p) {
  constexpr int TM_COLS = NB < 32 ? 32 : NB;
  constexpr bool CLUSTERED = (EPI == DG_EPI_RESID || EPI == DG_EPI_QKV);
  extern __shared__ uint8_t smem_raw[];
  DgSmem<NB>& S = *reinterpret_cast<DgSmem<NB>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int n0 = tile * DG_BM;
  const int split = blockIdx.y, n_splits = gridDim.y;  // CLUSTERED: the cluster is (1, n_splits, 1): rank in cluster == split
  const int kb0 = (int)(((int64_t)split * p.KB) / n_splits);
  const int kb1 = (int)(((int64_t)(split + 1) * p.KB) / n_splits);
  const int nkb = kb1 - kb0;
  const int M = p.M;
  // epilogue (part 2) operands that do not depend on the accumulator: fetched by the epilogue threads before they wait for the
  // main loop, so that their L2 latency is not in the chain behind the last MMA (two (row, column) assignments per thread are
  // covered; further ones are loaded on the fly)
  [[maybe_unused]] int e_pos[2] = {0, 0}, e_slot[2] = {0, 0};
  [[maybe_unused]] float e_c[2] = {1.f, 1.f}, e_s[2] = {0.f, 0.f}, e_x[2] = {0.f, 0.f};

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < DG_STAGES; ++s) {
      mbar_init(&S.full[s], PRO == DG_PRO_NORM ? 2 : 1);  // TMA producer (+ the norm-on-load warp)
      mbar_init(&S.empty[s], 1);
      mbar_init(&S.xfull[s], 1);
    }
    mbar_init(&S.wn_full, 1);
    mbar_init(&S.tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TM_COLS>(&S.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = S.tmem_base;

  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t BYTES = DG_BM * DG_BK * 2 + (PRO == DG_PRO_TMA ? NB * DG_BK * 2 : 0);
      constexpr uint32_t XBYTES = NB * DG_BK * 2;
      const int npre = min(nkb, DG_STAGES);
      const int I_rows = p.N >> 1, g0 = tile * 64;
      auto load_w = [&](int stage, int kb) {
        if constexpr (EPI == DG_EPI_SWIGLU) {  // 64 gate rows + the 64 matching up rows (8 KB apart keeps the swizzle phase)
          tma_load_2d(S.a[stage], &tmap_w, kb * DG_BK, g0, &S.full[stage]);
          tma_load_2d(S.a[stage] + 64 * DG_BK, &tmap_w, kb * DG_BK, I_rows + g0, &S.full[stage]);
        } else {
          tma_load_2d(S.a[stage], &tmap_w, kb * DG_BK, n0, &S.full[stage]);
        }
      };
      // weights (and the norm weight) do not depend on the predecessor: fill the ring first, then wait (PDL)
      if constexpr (PRO == DG_PRO_NORM) {
        mbar_expect_tx(&S.wn_full, (uint32_t)nkb * DG_BK * 2);
        tma_bulk_g2s(S.wn, p.norm_w + (int64_t)kb0 * DG_BK, (uint32_t)nkb * DG_BK * 2, &S.wn_full);
      }
      for (int i = 0; i < npre; ++i) {
        mbar_expect_tx(&S.full[i], BYTES);
        load_w(i, kb0 + i);
      }
      pdl_launch_dependents();
      pdl_wait();
      // PRO_TMA: the X tile completes the stage's `full` barrier; PRO_NORM: the RAW x tile lands on `xfull`, the norm-on-load warp
      // rewrites it in place and then arrives on `full`
      for (int i = 0; i < npre; ++i) {
        if constexpr (PRO == DG_PRO_TMA) {
          tma_load_2d(S.b[i], &tmap_x, (kb0 + i) * DG_BK, 0, &S.full[i]);
        } else {
          mbar_expect_tx(&S.xfull[i], XBYTES);
          tma_load_2d(S.b[i], &tmap_x, (kb0 + i) * DG_BK, 0, &S.xfull[i]);
        }
      }
      for (int i = npre; i < nkb; ++i) {
        const int s = i % DG_STAGES;
        mbar_wait(&S.empty[s], ((i / DG_STAGES) & 1) ^ 1);
        mbar_expect_tx(&S.full[s], BYTES);
        load_w(s, kb0 + i);
        if constexpr (PRO == DG_PRO_TMA) {
          tma_load_2d(S.b[s], &tmap_x, (kb0 + i) * DG_BK, 0, &S.full[s]);
        } else {
          mbar_expect_tx(&S.xfull[s], XBYTES);
          tma_load_2d(S.b[s], &tmap_x, (kb0 + i) * DG_BK, 0, &S.xfull[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t IDESC = umma_idesc_bf16(DG_BM, NB);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % DG_STAGES;
        mbar_wait(&S.full[s], (i / DG_STAGES) & 1);
        tc_fence_after();
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(S.a[s]));
        const uint64_t db = umma_desc_kmajor_sw128(smem_u32(S.b[s]));
#pragma unroll
        for (int k = 0; k < DG_BK / 16; ++k) umma_bf16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&S.empty[s]);
      }
      umma_commit(&S.tmem_full);
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter
    const int et = threadIdx.x - 64;  // 0..127
    if constexpr (EPI == DG_EPI_QKV) {
      pdl_wait();
      const int lo = (64 * split) / n_splits, cnt = (64 * (split + 1)) / n_splits - lo;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int idx = et + it * 128;
        if (idx < cnt * M) {
          const int m = idx / cnt, d = lo + idx % cnt;
          e_pos[it] = p.pos[m];
          e_slot[it] = p.kv_len[m];
          if (tile < 2 * p.heads) {
            e_c[it] = bf2f(p.cos_t[(int64_t)e_pos[it] * 128 + d]);
            e_s[it] = bf2f(p.sin_t[(int64_t)e_pos[it] * 128 + d]);
          }
        }
      }
    } else if constexpr (EPI == DG_EPI_RESID) {
      pdl_wait();
      const int lo = (DG_BM * split) / n_splits, cnt = (DG_BM * (split + 1)) / n_splits - lo;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int idx = et + it * 128;
        if (idx < cnt * M) {
          const int m = idx / cnt, n = n0 + lo + idx % cnt;
          if (n < p.N) e_x[it] = bf2f(p.xres[(int64_t)m * p.N + n]);
        }
      }
    }
    if (PRO == DG_PRO_NORM) {
      // ---------------- RMSNorm on load: the epilogue warps (idle during the main loop) turn each raw x tile (TMA, swizzled) into the
      // normalised B operand.  A STAGE belongs to one warp (stage s -> warp 2 + s % 4): its barrier phases are then observed strictly
      // in order by their only waiter, which can never be lapped (the stage is not refilled before that warp's own arrival let the
      // MMA consume it).  Dealing out k-blocks round-robin instead deadlocked in the real chain: a delayed warp was lapped by two
      // phases of a stage it merely observed, its parity wait aliased, and the MMA issuer waited for that warp's next arrival. ----------------
      constexpr int NBX = NB < 8 ? 8 : NB;
      constexpr int RPL = NBX / 4;  // rows per lane: row = (lane >> 3) + 4 j, physical 16-byte chunk pc = lane & 7
      const int sub = lane >> 3, pc = lane & 7;
      pdl_wait();
      // rstd: warp w sums the partials of rows m = w, w + 4, ...: up to 16 independent loads per lane go out together (row after row with
      // a dependent loop cost one L2 round trip per 4 partials: +8-10 us per launch at 256 partials), rows meet in shared memory
      for (int m = warp - 2; m < M; m += 4) {
        const float* src = p.ssq + (int64_t)m * p.n_parts;
        float sacc = 0.f;
        for (int p0 = 0; p0 < p.n_parts; p0 += 512) {
          float t[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) t[u] = (p0 + u * 32 + lane < p.n_parts) ? src[p0 + u * 32 + lane] : 0.f;
#pragma unroll
          for (int u = 0; u < 16; ++u) sacc += t[u];
        }
        sacc = warp_sum(sacc);
        if (lane == 0) S.rstd[m] = 1.0f / sqrtf(sacc / (float)p.K + p.eps);
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");  // the four epilogue warps
      float rstd[RPL];
#pragma unroll
      for (int j = 0; j < RPL; ++j) rstd[j] = (sub + 4 * j < M) ? S.rstd[sub + 4 * j] : 0.f;
      mbar_wait(&S.wn_full, 0);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % DG_STAGES;
        if ((s & 3) != warp - 2) continue;
        mbar_wait(&S.xfull[s], (i / DG_STAGES) & 1);
        bf16* tilep = S.b[s];
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          const int row = sub + 4 * j;
          if (row < M) {
            const int c = pc ^ (row & 7);  // logical chunk (8 columns) stored at physical chunk pc of this row (128B swizzle)
            float xv[8], wv[8];
            bf16* ptr = tilep + row * DG_BK + pc * 8;
            unpack8(*reinterpret_cast<const bf16x8*>(ptr), xv);
            unpack8(*reinterpret_cast<const bf16x8*>(S.wn + i * DG_BK + c * 8), wv);
#pragma unroll
            for (int e = 0; e < 8; ++e) xv[e] = wv[e] * rbf(xv[e] * rstd[j]);
            st8(ptr, pack8(xv));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.full[s]);
      }
    }

    // ---------------- epilogue, part 1: accumulator out of TMEM ----------------
    const int n_loc = q * 32 + lane;
    float* red = reinterpret_cast<float*>(S.a[0]);  // [128][NB] fp32: every stage has been consumed once tmem_full fires
    uint32_t r[NB];
    if (nkb > 0) {
      mbar_wait(&S.tmem_full, 0);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += 16) tmem_ld_32x32b_x16(tmem_d + ((uint32_t)(q * 32) << 16) + c0, r + c0);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int m = 0; m < NB; ++m) r[m] = 0u;
    }
    if constexpr (CLUSTERED) {
#pragma unroll
      for (int m = 0; m < NB; m += 4) *reinterpret_cast<uint4*>(red + n_loc * NB + m) = make_uint4(r[m], r[m + 1], r[m + 2], r[m + 3]);
    } else if constexpr (EPI == DG_EPI_LOGITS) {
      const int n = n0 + n_loc;
      if (n < p.N) {
#pragma unroll
        for (int m = 0; m < NB; ++m)
          if (m < M) p.out[(int64_t)m * p.N + n] = f2bf(__uint_as_float(r[m]));
      }
    } else if constexpr (EPI == DG_EPI_SWIGLU) {
      const int row = (q & 1) * 32 + lane;  // index inside the 64-row half
      if (q >= 2) {
#pragma unroll
        for (int m = 0; m < NB; ++m) red[row * NB + m] = __uint_as_float(r[m]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
      if (q < 2) {
        const int I = p.N >> 1;
        const int col = tile * 64 + row;
        if (col < I) {
#pragma unroll
          for (int m = 0; m < NB; ++m) {
            if (m < M) {
              const float g = rbf(__uint_as_float(r[m])), u = rbf(red[row * NB + m]);
              p.out[(int64_t)m * I + col] = f2bf(rbf(g / (1.0f + expf(-g))) * u);
            }
          }
        }
      }
    }
  }

  if constexpr (CLUSTERED) {
    tc_fence_before();
    dg_cluster_sync();  // every rank's partial tile is in its shared memory
    if (warp >= 2) {
      // ---------------- epilogue, part 2: fixed-order reduce over the cluster through DSMEM ----------------
      float* red = reinterpret_cast<float*>(S.a[0]);
      const uint32_t red_u32 = smem_u32(red);
      const int et = threadIdx.x - 64;
      if constexpr (EPI == DG_EPI_RESID) {
        float* sq = reinterpret_cast<float*>(S.a[2]);  // [cnt * M] squares of the updated residual (red may span a[0..1] at NB = 64)
        const int lo = (DG_BM * split) / n_splits, cnt = (DG_BM * (split + 1)) / n_splits - lo;
        int it = 0;
        for (int idx = et; idx < cnt * M; idx += 128, ++it) {
          const int m = idx / cnt, nn = lo + idx % cnt;
          float a = 0.f;
          for (int j = 0; j < n_splits; ++j) a += dg_ld_dsmem(red_u32 + (uint32_t)(nn * NB + m) * 4u, (uint32_t)j);
          const int n = n0 + nn;
          float xo = 0.f;
          if (n < p.N) {
            const float xin = it == 0 ? e_x[0] : (it == 1 ? e_x[1] : bf2f(p.xres[(int64_t)m * p.N + n]));
            xo = rbf(xin + rbf(a));
            p.xres[(int64_t)m * p.N + n] = f2bf(xo);
          }
          sq[idx] = xo * xo;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et < M) {
          float sacc = 0.f;
          for (int i = 0; i < cnt; ++i) sacc += sq[et * cnt + i];  // fixed order
          p.ssq_out[(int64_t)et * ((int64_t)gridDim.x * n_splits) + (int64_t)tile * n_splits + split] = sacc;
        }
      } else {  // DG_EPI_QKV
        const int item = tile / p.heads, h = tile % p.heads;  // 0 = q, 1 = k, 2 = v
        const int lo = (64 * split) / n_splits, cnt = (64 * (split + 1)) / n_splits - lo;
        int it = 0;
        for (int idx = et; idx < cnt * M; idx += 128, ++it) {
          const int m = idx / cnt, d = lo + idx % cnt;
          float x1 = 0.f, x2 = 0.f;
          for (int j = 0; j < n_splits; ++j) {
            x1 += dg_ld_dsmem(red_u32 + (uint32_t)(d * NB + m) * 4u, (uint32_t)j);
            x2 += dg_ld_dsmem(red_u32 + (uint32_t)((d + 64) * NB + m) * 4u, (uint32_t)j);
          }
          x1 = rbf(x1);
          x2 = rbf(x2);
          int slot;
          float c = 1.f, sn = 0.f;
          if (it < 2) {
            slot = it == 0 ? e_slot[0] : e_slot[1];
            c = it == 0 ? e_c[0] : e_c[1];
            sn = it == 0 ? e_s[0] : e_s[1];
          } else {
            slot = p.kv_len[m];
            if (item < 2) {
              c = bf2f(p.cos_t[(int64_t)p.pos[m] * 128 + d]);
              sn = bf2f(p.sin_t[(int64_t)p.pos[m] * 128 + d]);
            }
          }
          if (item < 2) {
            const float o1 = rbf(rbf(x1 * c) + rbf(-x2 * sn));
            const float o2 = rbf(rbf(x2 * c) + rbf(x1 * sn));
            x1 = o1;
            x2 = o2;
          }
          bf16* dst;
          if (item == 0)
            dst = p.q_out + (int64_t)m * p.heads * 128 + h * 128;
          else
            dst = (item == 1 ? p.k_arena : p.v_arena) + (((int64_t)m * p.heads + h) * p.cap + slot) * 128;
          dst[d] = f2bf(x1);
          dst[d + 64] = f2bf(x2);
        }
      }
    }
    dg_cluster_sync();  // no rank leaves (or frees its shared memory) while another may still read it
  } else {
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 1) tmem_dealloc<TM_COLS>(tmem_d);
}

// ------------------------------------------------------------------------------------------------------------------------------
template <int NB, int PRO, int EPI>
static int dg_launch(const void* w, int w_box_rows, const void* x_tma, const DgParams& p, int tiles, int splits, cudaStream_t stream) {
  alignas(64) CUtensorMap tw, tx;
  const int K = p.KB * DG_BK;
  int rc = dl_make_tmap_2d(&tw, w, (uint64_t)p.N, (uint64_t)K, (uint64_t)K, (uint32_t)w_box_rows, DG_BK);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tx, PRO == DG_PRO_TMA ? x_tma : (const void*)p.x, (uint64_t)p.M, (uint64_t)K, (uint64_t)K, (uint32_t)NB, DG_BK);
  if (rc) return rc;
  if (PRO == DG_PRO_NORM && K > DG_MAX_NORM_K) return dl_set_error(-1, "dl_dec_gemm: norm-on-load supports K <= 5120");
  const int smem = (int)sizeof(DgSmem<NB>) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dec_gemm_kernel<NB, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_dec_gemm(smem attr)");
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles, splits);
  cfg.blockDim = dim3(DG_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (EPI == DG_EPI_RESID || EPI == DG_EPI_QKV) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = splits;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le = cudaLaunchKernelEx(&cfg, dec_gemm_kernel<NB, PRO, EPI>, tw, tx, p);
  if (le != cudaSuccess) return dl_set_error((int)le, "dl_dec_gemm(launch)");
  DL_CHECK_LAUNCH("dl_dec_gemm");
}

#define DG_DISPATCH_NB(M, CALL)                \
  do {                                         \
    if ((M) <= 16) { constexpr int NB_ = 16; CALL; } \
    else if ((M) <= 32) { constexpr int NB_ = 32; CALL; } \
    else { constexpr int NB_ = 64; CALL; }     \
  } while (0)

// split-K factor = cluster size (<= 8, the portable maximum) that brings the grid near one wave of 2 CTAs per SM
static int dg_pick_splits(int tiles, int KB) {
  int s = (296 + tiles / 2) / tiles;
  if (tiles >= 148) s = 1;
  if (s > KB / 4) s = KB / 4;
  if (s > 8) s = 8;
  if (s < 1) s = 1;
  return s;
}

extern "C" int dl_dec_gemm_qkv(const void* x, const float* ssq, int n_parts, const void* norm_w, float eps, const void* w_qkv,
                               void* q_out, void* k_arena, void* v_arena, const int32_t* kv_len, const int32_t* pos,
                               const void* cos_table, const void* sin_table, int M, int heads, int head_dim, int cap, int K,
                               cudaStream_t stream) {
  DL_REQUIRE(M >= 1 && M <= 64 && head_dim == 128 && K % DG_BK == 0 && K % 8 == 0, "dl_dec_gemm_qkv");
  DgParams p = {};
  p.x = (const bf16*)x; p.ssq = ssq; p.n_parts = n_parts; p.norm_w = (const bf16*)norm_w; p.eps = eps; p.K = K;
  p.M = M; p.N = 3 * heads * 128; p.KB = K / DG_BK;
  p.q_out = (bf16*)q_out; p.k_arena = (bf16*)k_arena; p.v_arena = (bf16*)v_arena; p.kv_len = kv_len; p.pos = pos;
  p.cos_t = (const bf16*)cos_table; p.sin_t = (const bf16*)sin_table; p.heads = heads; p.cap = cap;
  const int tiles = 3 * heads;
  const int splits = dg_pick_splits(tiles, p.KB);
  DG_DISPATCH_NB(M, return (dg_launch<NB_, DG_PRO_NORM, DG_EPI_QKV>(w_qkv, DG_BM, nullptr, p, tiles, splits, stream)));
  return 0;
}

extern "C" int dl_dec_gemm_residual(const void* x_in, const void* w, void* xres, float* ssq_out, int M, int N, int K,
                                    cudaStream_t stream) {
  DL_REQUIRE(M >= 1 && M <= 64 && K % DG_BK == 0 && N % 8 == 0, "dl_dec_gemm_residual");
  DgParams p = {};
  p.M = M; p.N = N; p.KB = K / DG_BK; p.K = K;
  p.xres = (bf16*)xres; p.ssq_out = ssq_out;
  const int tiles = (N + DG_BM - 1) / DG_BM;
  const int splits = dg_pick_splits(tiles, p.KB);
  DG_DISPATCH_NB(M, return (dg_launch<NB_, DG_PRO_TMA, DG_EPI_RESID>(w, DG_BM, x_in, p, tiles, splits, stream)));
  return 0;
}
extern "C" int dl_dec_gemm_residual_parts(int N, int K) {
  const int tiles = (N + DG_BM - 1) / DG_BM;
  return tiles * dg_pick_splits(tiles, K / DG_BK);
}

extern "C" int dl_dec_gemm_swiglu(const void* x, const float* ssq, int n_parts, const void* norm_w, float eps, const void* w_gate_up,
                                  void* act_out, int M, int I, int K, cudaStream_t stream) {
  DL_REQUIRE(M >= 1 && M <= 64 && K % DG_BK == 0 && I % 64 == 0, "dl_dec_gemm_swiglu");
  DgParams p = {};
  p.x = (const bf16*)x; p.ssq = ssq; p.n_parts = n_parts; p.norm_w = (const bf16*)norm_w; p.eps = eps; p.K = K;
  p.M = M; p.N = 2 * I; p.KB = K / DG_BK;
  p.out = (bf16*)act_out;
  DG_DISPATCH_NB(M, return (dg_launch<NB_, DG_PRO_NORM, DG_EPI_SWIGLU>(w_gate_up, 64, nullptr, p, I / 64, 1, stream)));
  return 0;
}

extern "C" int dl_dec_gemm_logits(const void* x, const float* ssq, int n_parts, const void* norm_w, float eps, const void* w,
                                  void* logits, int M, int N, int K, cudaStream_t stream) {
  DL_REQUIRE(M >= 1 && M <= 64 && K % DG_BK == 0 && N % 8 == 0, "dl_dec_gemm_logits");
  DgParams p = {};
  p.x = (const bf16*)x; p.ssq = ssq; p.n_parts = n_parts; p.norm_w = (const bf16*)norm_w; p.eps = eps; p.K = K;
  p.M = M; p.N = N; p.KB = K / DG_BK;
  p.out = (bf16*)logits;
  DG_DISPATCH_NB(M, return (dg_launch<NB_, DG_PRO_NORM, DG_EPI_LOGITS>(w, DG_BM, nullptr, p, (N + DG_BM - 1) / DG_BM, 1, stream)));
  return 0;
}

// embed_tokens for the decode step + the row's sum of squares (one partial) for the first layer's RMSNorm-on-load
__global__ void __launch_bounds__(256) embed_rows_ssq_kernel(const int64_t* __restrict__ ids, const bf16* __restrict__ table,
                                                             bf16* __restrict__ out, float* __restrict__ ssq, int H) {
  __shared__ float redbuf[32];
  pdl_launch_dependents();
  pdl_wait();
  const int64_t id = ids[blockIdx.x];
  const bf16* src = table + id * H;
  bf16* dst = out + (int64_t)blockIdx.x * H;
  float s = 0.f;
  for (int c = threadIdx.x; c < (H >> 3); c += blockDim.x) {
    const bf16x8 v = ld8(src + c * 8);
    st8(dst + c * 8, v);
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += f[e] * f[e];
  }
  s = block_sum(s, redbuf);
  if (threadIdx.x == 0) ssq[blockIdx.x] = s;
}
extern "C" int dl_embed_rows_ssq(const int64_t* ids, const void* table, void* out, float* ssq, int n, int H, cudaStream_t stream) {
  DL_REQUIRE(H % 8 == 0, "dl_embed_rows_ssq");
  if (n == 0) return 0;
  dl_launch_pdl(embed_rows_ssq_kernel, dim3(n), dim3(256), 0, stream, ids, (const bf16*)table, (bf16*)out, ssq, H);
  DL_CHECK_LAUNCH("dl_embed_rows_ssq");
}
