// Dense bf16 GEMM for every nn.Linear at q_len > 1 (prefill q/k/v/o/gate/up/down, CLIP, predictor blocks, projector, lm_head over
// all positions):      out[M, N] = epilogue( X[M, K] . W[N, K]^T )        (nn.Linear layout, fp32 accumulate)
// Reference call sites: llava/model/language_model/dynamic_modeling_llama.py:294-330 (LlamaMLP), :377-392 + :1011-1013 + :1127
// (q/k/v/o_proj), :1324-1346 (VisionPredictor linears), :2709 (lm_head); custom_transformer_layer.py:86-180 (predictor blocks);
// multimodal_projector/builder.py:172-179; HF CLIPEncoderLayer / CLIPVisionEmbeddings (unvendored).
//
// B200 design: persistent kernel, one CTA per SM, CTAs paired into 2-CTA clusters that share one 256 x 256 output tile
// (`tcgen05.mma.cta_group::2`, M = 256, N = 256).  Per pair and 64-wide k-block each CTA TMA-loads ITS 128 rows of X and ITS 128 rows
// of W: every operand byte is fetched from L2 once per pair (64 B/clk/SM at the tensor peak — a 256 x 128 tile needs 96 B/clk/SM and
// measured L2-bound) and each SM's tensor core reads only half of B from its own shared memory.  The leader CTA's single elected
// thread issues the MMAs for both SMs; accumulators live in TMEM, double buffered (2 x 256 columns), so the eight epilogue warps of
// each CTA drain segment i (tcgen05.ld -> bias / activation / residual / SwiGLU -> bf16 -> swizzled smem -> TMA store) while the
// tensor cores already run segment i+1.
//   warp 0: TMA producer (6-stage mbarrier ring, both CTAs; completion bytes of BOTH CTAs land on the leader's `full` barrier)
//   warp 1: MMA issuer (leader CTA only) + TMEM allocation (both CTAs, cta_group::2)
//   warps 4-11: epilogue (TMEM lane quarter = warp % 4; the two warps of a quarter take alternate 64-column chunks); the register
//   file is rebalanced with setmaxnreg: 40 per thread for warp group 0, 232 for the two epilogue warp groups
// Barriers: full[s] (leader: 1 arrival + 2 x stage bytes), empty[s] (both CTAs, multicast tcgen05.commit), tmem_full[a] (both CTAs,
// multicast commit), tmem_empty[a] (leader: 16 arrivals = 2 CTAs x 8 epilogue warps, the peer's arrive remotely through DSMEM).
//
// Work distribution: whole tiles, round-robin over the pairs in lockstep waves (concurrent pairs sit at the same k of neighbouring
// tiles, so X / W tiles are shared through L2 — a fully desynchronised stream-K measured 11x the algorithmic DRAM traffic on the
// 4856 x 12288 x 4096 prefill GEMM), and STREAM-K ONLY FOR THE REMAINDER: the tiles of the last, partial wave (M = 8 x 146 prefill
// rows give 80 tiles for 74 pairs: 2 waves tile-wise, 1.08 here) are cut in the (tile, k-block) space into one equal range per
// pair.  A range that starts inside a tile computes that tile's k-blocks and parks the fp32 partial accumulator in the
// caller-provided workspace (one slot per pair, per-warp release flags); the pair whose range holds the tile's k = 0 block finishes
// it: it adds the parked partials of the following pairs in fixed order (deterministic) and runs the epilogue.  Writers never
// wait, so there is no cyclic dependency; every CTA triggers its PDL dependents only once it is resident, so a finisher never
// waits for a pair that cannot be scheduled.  Flags are reset by their single reader: the workspace is all-zero again after every
// launch (CUDA-graph replayable).  Tile order: bands of `band` m-tiles sized so that a band of X stays L2-resident, m fastest.
#include <cuda.h>

#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define G2_BK 64
#define G2_BN 256
#define G2_STAGES 6
#define G2_EPI_WARPS 8
#define G2_THREADS (32 * (4 + G2_EPI_WARPS))  // warp group 0: producer, MMA issuer, two idle warps; warp groups 1-2: epilogue
#define G2_SLOT_FLOATS (2 * 4 * 256 * 32)  // one pair's parked accumulator: 2 CTAs x 4 lane quarters x 256 columns x 32 lanes
#define G2_SLOT_FLAGS 16                   // 2 CTAs x 8 epilogue warps

enum { G2_EPI_NONE = 0, G2_EPI_GELU = 1, G2_EPI_QUICK_GELU = 2, G2_EPI_RESIDUAL = 3, G2_EPI_SWIGLU = 4 };

struct G2Smem {
  alignas(1024) bf16 a[G2_STAGES][128 * G2_BK];
  alignas(1024) bf16 b[G2_STAGES][128 * G2_BK];
  alignas(1024) uint8_t stage_out[G2_EPI_WARPS][4096];  // per epilogue warp: one 32-row x 64-column bf16 box (128B swizzle)
  alignas(8) uint64_t full[G2_STAGES];
  alignas(8) uint64_t empty[G2_STAGES];
  alignas(8) uint64_t tmem_full[2];
  alignas(8) uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

DL_DEVINL uint32_t g2_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
DL_DEVINL void g2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in the CTA of rank `rank`
DL_DEVINL uint32_t g2_mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
DL_DEVINL void g2_mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA load: data lands in THIS CTA's shared memory, the completion bytes on the barrier at `mbar_cluster_addr` (the leader's)
DL_DEVINL void g2_tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint32_t mbar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
DL_DEVINL void g2_umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive (once complete) on the barrier at this offset in BOTH CTAs of the pair
DL_DEVINL void g2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
template <int COLS>
DL_DEVINL void g2_tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
DL_DEVINL void g2_tmem_dealloc(uint32_t tmem_addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "n"(COLS) : "memory");
}
DL_DEVINL void g2_tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(smem_u32(smem_src)),
               "r"(c0), "r"(c1)
               : "memory");
}
DL_DEVINL int g2_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
DL_DEVINL void g2_st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// Work list of one pair: `dp_waves` whole tiles (tile = wave * n_pairs + pair), then — stream-K remainder — the range
// [sk_begin(pair), sk_begin(pair + 1)) of the (tile, k-block) space of the `sk_tiles` tiles behind the whole waves.
struct G2Sched {
  int tiles_m, tiles_n, band, KB, n_pairs, dp_waves, sk_tiles;
  DL_DEVINL int n_tiles() const { return tiles_m * tiles_n; }
  DL_DEVINL int sk_begin(int p) const { return (int)(((int64_t)p * sk_tiles * KB) / n_pairs); }
  // tile -> (m-tile, n-tile): bands of `band` m-tiles, m fastest inside a band
  DL_DEVINL void coords(int tile, int& mt, int& nt) const {
    const int band_sz = band * tiles_n;
    const int b = tile / band_sz, r = tile - b * band_sz;
    const int m_in_band = min(band, tiles_m - b * band);
    nt = r / m_in_band;
    mt = b * band + (r - nt * m_in_band);
  }
  // f(tile, kb0, kb1, sk_tile_local) for every segment of `pair`, in execution order (sk_tile_local < 0: a whole data-parallel tile)
  template <class F>
  DL_DEVINL void for_each_segment(int pair, F&& f) const {
    for (int w = 0; w < dp_waves; ++w) {
      const int tile = w * n_pairs + pair;
      if (tile < n_tiles()) f(tile, 0, KB, -1);
    }
    if (sk_tiles > 0) {
      const int base = dp_waves * n_pairs;
      const int g1 = sk_begin(pair + 1);
      for (int g = sk_begin(pair); g < g1;) {
        const int t = g / KB, kb0 = g - t * KB, kb1 = min(KB, kb0 + (g1 - g));
        f(base + t, kb0, kb1, t);
        g += kb1 - kb0;
      }
    }
  }
};

DL_DEVINL float g2_gelu_erf(float y) { return 0.5f * y * (1.0f + erff(y * 0.70710678118654752440f)); }
DL_DEVINL float g2_sigmoid(float t) { return __fdividef(1.0f, 1.0f + __expf(-t)); }

// EPI_SWIGLU: W = [gate rows (I) | up rows (I)]; the pair's B tile is gate rows [128 nt, +128) (CTA 0) and the matching up rows
// (CTA 1), so accumulator columns [0,128) / [128,256) hold gate / up of the SAME 128 outputs; out is [M, I], N_out = I.
template <int EPI>
__global__ void __launch_bounds__(G2_THREADS, 1) gemm2_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                              const __grid_constant__ CUtensorMap tmap_w,
                                                              const __grid_constant__ CUtensorMap tmap_o,
                                                              const bf16* __restrict__ bias, const bf16* residual, int M, int N_out,
                                                              const G2Sched sched, int up_row0, float* ws_partial, int* ws_flags) {
  constexpr int BN = G2_BN;
  constexpr int STAGES = G2_STAGES;
  constexpr int TM_COLS = 2 * BN;
  constexpr uint32_t STAGE_BYTES = (128 + BN / 2) * G2_BK * 2;
  constexpr int OUT_W = EPI == G2_EPI_SWIGLU ? BN / 2 : BN;  // output columns per tile
  extern __shared__ uint8_t smem_raw[];
  G2Smem& S = *reinterpret_cast<G2Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = g2_cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int KB = sched.KB;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_o);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&S.tmem_full[a], 1);
      mbar_init(&S.tmem_empty[a], 2 * G2_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) g2_tmem_alloc<TM_COLS>(&S.tmem_base);
  tc_fence_before();
  g2_cluster_sync();  // barrier inits and the TMEM allocation of BOTH CTAs are visible before any remote arrive / 2-SM MMA
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_base;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ------------------------------------------------ TMA producer (both CTAs) ------------------------------------------------
    if (elect_one()) {
      // every CTA signals at its start, i.e. once it is resident: the dependent grid can only take SMs after ALL pairs run (a pair
      // that waits for a parked partial must never wait for a pair that cannot be scheduled)
      pdl_launch_dependents();
      pdl_wait();  // X (and a residual) come from the predecessor kernel
      int it = 0;
      sched.for_each_segment(pair, [&](int tile, int kb0, int kb1, int) {
        int mt, nt;
        sched.coords(tile, mt, nt);
        const int m0 = mt * 256 + (int)rank * 128;
        const int n0 = EPI == G2_EPI_SWIGLU ? (rank == 0 ? nt * (BN / 2) : up_row0 + nt * (BN / 2)) : nt * BN + (int)rank * (BN / 2);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&S.empty[s], ph ^ 1);
          const uint32_t full_leader = g2_mapa(smem_u32(&S.full[s]), 0);
          if (rank == 0) mbar_expect_tx(&S.full[s], 2 * STAGE_BYTES);
          g2_tma_load_2d(S.a[s], &tmap_x, kb * G2_BK, m0, full_leader);
          g2_tma_load_2d(S.b[s], &tmap_w, kb * G2_BK, n0, full_leader);
        }
      });
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA) ------------------------------------------------
    if (rank == 0 && elect_one()) {
      constexpr uint32_t IDESC = umma_idesc_bf16(256, BN);
      int it = 0, sc = 0;
      sched.for_each_segment(pair, [&](int, int kb0, int kb1, int) {
        const int a = sc & 1;
        mbar_wait(&S.tmem_empty[a], ((sc >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(a * BN);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&S.full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint64_t da = umma_desc_kmajor_sw128(smem_u32(S.a[s]));
          const uint64_t db = umma_desc_kmajor_sw128(smem_u32(S.b[s]));
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k) g2_umma(d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
          g2_commit_both(&S.empty[s]);  // the stage is free in both CTAs once these MMAs have read it
        }
        g2_commit_both(&S.tmem_full[a]);
        ++sc;
      });
    }
  }
  } else {
    // ------------------------------------------------ epilogue (both CTAs) ------------------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // The stream-K workspace is shared by every dl_linear launch of the stream, and this kernel is launched with programmatic
    // serialization: its CTAs can be resident while the PREVIOUS GEMM's pairs are still parking and gathering through the same slots
    // and flags.  Every warp that touches global memory on its own initiative must therefore execute griddepcontrol.wait itself —
    // the producer's wait orders only what depends on its TMA loads.  (Round 2 bug: a tile finisher polled the flags before anything
    // ordered it behind the producer's wait and consumed the previous launch's flags; profiles/r2_stream_k_workspace_race.md.)
    pdl_wait();
    const int ew = warp - 4;
    const int q = warp & 3;   // TMEM lane quarter of this warp = rows [32q, 32q+32) of this CTA's half tile
    const int hh = ew >> 2;   // which of the two warps of the quarter: takes the 64-column chunks c with c % 2 == hh
    uint8_t* stage = S.stage_out[ew];
    const int64_t my_part = ((int64_t)rank * 4 + q) * (256 * 32);  // float offset of this warp pair's region inside a slot
    const int my_flag = (int)rank * 8 + ew;
    int sc = -1;
    sched.for_each_segment(pair, [&](int tile, int kb0, int kb1, int skt) {
      ++sc;
      int mt, nt;
      sched.coords(tile, mt, nt);
      const int a = sc & 1;
      const int row0 = mt * 256 + (int)rank * 128 + q * 32;
      const int col_t = nt * OUT_W;
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN);
      const uint32_t tmem_empty_leader = g2_mapa(smem_u32(&S.tmem_empty[a]), 0);

      if (kb0 > 0) {
        // ---- this range started inside a tile: park the raw fp32 accumulator for the pair that finishes the tile ----
        mbar_wait(&S.tmem_full[a], (sc >> 1) & 1);
        tc_fence_after();
        float* dst = ws_partial + (int64_t)pair * G2_SLOT_FLOATS + my_part;
#pragma unroll 1
        for (int rc = hh; rc < BN / 64; rc += 2) {
          uint32_t r[64];
          tmem_ld_32x32b_x32(tacc + rc * 64, r);
          tmem_ld_32x32b_x32(tacc + rc * 64 + 32, r + 32);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 16; ++u)  // unit = 4 columns; 32 lanes x 16 B contiguous per unit
            *reinterpret_cast<uint4*>(dst + ((rc * 16 + u) * 32 + lane) * 4) = make_uint4(r[u * 4], r[u * 4 + 1], r[u * 4 + 2], r[u * 4 + 3]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          g2_mbar_arrive_cluster(tmem_empty_leader);
          __threadfence();
          g2_st_release(ws_flags + pair * G2_SLOT_FLAGS + my_flag, 1);
        }
        return;
      }

      const bool gather = kb1 < KB;  // the tile's remaining k-blocks were computed (first thing) by the following pairs
      if (gather) {
        // wait until this warp's counterpart in every contributing pair has parked its slice: polled BEFORE this warp's own accumulator
        // is waited for, so the flag round trips through L2 hide under the pair's main loop.  Safe only because this warp executed
        // griddepcontrol.wait itself (top of the epilogue branch): the workspace is quiescent, see there.
        if (lane == 0) {
          const int tile_end = (skt + 1) * KB;
          for (int p2 = pair + 1;; ++p2) {
            const int* f = ws_flags + p2 * G2_SLOT_FLAGS + my_flag;
            while (g2_ld_acquire(f) == 0) {
            }
            if (sched.sk_begin(p2 + 1) >= tile_end) break;
          }
        }
        __syncwarp();
      }
      mbar_wait(&S.tmem_full[a], (sc >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = hh; c < OUT_W / 64; c += 2) {
        const int col0 = col_t + c * 64;
        // residual: the 32 x 64 box is fetched with 8 fully coalesced 16-byte loads per lane (4 rows x 128 B per instruction), issued
        // before the TMEM loads so that their latency hides, then transposed through the staging box so each lane gets its own row
        uint4 rres[8];
        if constexpr (EPI == G2_EPI_RESIDUAL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = row0 + i * 4 + (lane >> 3), cc = col0 + (lane & 7) * 8;
            rres[i] = (rr < M && cc < N_out) ? *reinterpret_cast<const uint4*>(residual + (int64_t)rr * N_out + cc) : make_uint4(0, 0, 0, 0);
          }
        }
        uint32_t r[64];
        uint32_t ru[EPI == G2_EPI_SWIGLU ? 64 : 1];
        tmem_ld_32x32b_x32(tacc + c * 64, r);
        tmem_ld_32x32b_x32(tacc + c * 64 + 32, r + 32);
        if constexpr (EPI == G2_EPI_SWIGLU) {
          tmem_ld_32x32b_x32(tacc + BN / 2 + c * 64, ru);
          tmem_ld_32x32b_x32(tacc + BN / 2 + c * 64 + 32, ru + 32);
        }
        tmem_ld_wait();
        if (c + 2 >= OUT_W / 64) {  // this warp has read its whole accumulator slice: hand the TMEM stage back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) g2_mbar_arrive_cluster(tmem_empty_leader);
        }
        if (gather) {
          const int tile_end = (skt + 1) * KB;
          for (int p2 = pair + 1;; ++p2) {  // fixed order: deterministic sums
            const float* src = ws_partial + (int64_t)p2 * G2_SLOT_FLOATS + my_part;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(src + ((c * 16 + u) * 32 + lane) * 4));
              r[u * 4] = __float_as_uint(__uint_as_float(r[u * 4]) + v.x);
              r[u * 4 + 1] = __float_as_uint(__uint_as_float(r[u * 4 + 1]) + v.y);
              r[u * 4 + 2] = __float_as_uint(__uint_as_float(r[u * 4 + 2]) + v.z);
              r[u * 4 + 3] = __float_as_uint(__uint_as_float(r[u * 4 + 3]) + v.w);
            }
            if constexpr (EPI == G2_EPI_SWIGLU) {
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (((c + 2) * 16 + u) * 32 + lane) * 4));
                ru[u * 4] = __float_as_uint(__uint_as_float(ru[u * 4]) + v.x);
                ru[u * 4 + 1] = __float_as_uint(__uint_as_float(ru[u * 4 + 1]) + v.y);
                ru[u * 4 + 2] = __float_as_uint(__uint_as_float(ru[u * 4 + 2]) + v.z);
                ru[u * 4 + 3] = __float_as_uint(__uint_as_float(ru[u * 4 + 3]) + v.w);
              }
            }
            if (sched.sk_begin(p2 + 1) >= tile_end) break;
          }
        }
        // the TMA store that last used this warp's staging box has read it
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        bf16x8 resrow[EPI == G2_EPI_RESIDUAL ? 8 : 1];
        if constexpr (EPI == G2_EPI_RESIDUAL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3), ch = lane & 7;
            *reinterpret_cast<uint4*>(stage + rr * 128 + ((ch ^ (rr & 7)) << 4)) = rres[i];
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) resrow[j] = *reinterpret_cast<const bf16x8*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4));
          __syncwarp();
        }
        bf16x8 pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = col0 + j * 8;
          float f[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(r[j * 8 + u]);
          if (bias != nullptr && col < N_out) {  // N_out % 8 == 0
            float bv[8];
            unpack8(ld8(bias + col), bv);
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] += bv[u];
          }
          if constexpr (EPI == G2_EPI_GELU) {  // nn.GELU (erf) on the bf16-rounded linear output, as the reference's separate modules do
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = g2_gelu_erf(rbf(f[u]));
          } else if constexpr (EPI == G2_EPI_QUICK_GELU) {  // HF quick_gelu on bf16 tensors: t = bf16(1.702 y); s = bf16(sigmoid t); y * s
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float y = rbf(f[u]);
              f[u] = y * rbf(g2_sigmoid(rbf(1.702f * y)));
            }
          } else if constexpr (EPI == G2_EPI_RESIDUAL) {  // x + linear(...): bf16 linear output, then a bf16 add (:1289, :1295)
            float rv[8];
            unpack8(resrow[j], rv);
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = rbf(f[u]) + rv[u];
          } else if constexpr (EPI == G2_EPI_SWIGLU) {  // bf16(bf16(silu(bf16 g)) * bf16 u)  (:328)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float gt = rbf(f[u]), up = rbf(__uint_as_float(ru[j * 8 + u]));
              f[u] = rbf(gt * g2_sigmoid(gt)) * up;
            }
          }
          pk[j] = pack8(f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<bf16x8*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) = pk[j];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (row0 < M && col0 < N_out) g2_tma_store_2d(&tmap_o, stage, col0, row0);  // the tensor map clips rows >= M, columns >= N_out
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (gather) {  // single reader of these flags: hand them back zeroed (the workspace is all-zero again after the launch)
        __syncwarp();
        if (lane == 0) {
          const int tile_end = (skt + 1) * KB;
          for (int p2 = pair + 1;; ++p2) {
            ws_flags[p2 * G2_SLOT_FLAGS + my_flag] = 0;
            if (sched.sk_begin(p2 + 1) >= tile_end) break;
          }
        }
      }
    });
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  g2_cluster_sync();  // neither CTA frees TMEM (or exits) while the pair's MMAs / remote arrives may still touch it
  if (warp == 1) g2_tmem_dealloc<TM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------------------
static int g2_num_sms() {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  return n_sm;
}

extern "C" int64_t dl_linear_workspace_bytes(void) {
  const int64_t slots = g2_num_sms() / 2 + 1;
  return slots * (int64_t)(G2_SLOT_FLOATS * sizeof(float) + G2_SLOT_FLAGS * sizeof(int));
}

template <int EPI>
static int g2_launch(const void* x, const void* w, const void* bias, const void* residual, void* out, int M, int N_w, int N_out, int K,
                     int up_row0, void* workspace, cudaStream_t stream) {
  alignas(64) CUtensorMap tx, tw, to;
  int rc = dl_make_tmap_2d(&tx, x, (uint64_t)M, (uint64_t)K, (uint64_t)K, 128, G2_BK);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tw, w, (uint64_t)N_w, (uint64_t)K, (uint64_t)K, G2_BN / 2, G2_BK);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&to, out, (uint64_t)M, (uint64_t)N_out, (uint64_t)N_out, 32, 64);
  if (rc) return rc;
  const int smem = (int)sizeof(G2Smem) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_linear(smem attr)");
    attr_set = true;
  }
  constexpr int OUT_W = EPI == G2_EPI_SWIGLU ? G2_BN / 2 : G2_BN;
  G2Sched sc;
  sc.tiles_m = (M + 255) / 256;
  sc.tiles_n = (N_out + OUT_W - 1) / OUT_W;
  sc.KB = K / G2_BK;
  const int n_tiles = sc.tiles_m * sc.tiles_n;
  // a band of X (band x 256 rows x K) should stay L2-resident while the pairs sweep the n-tiles: ~48 MB of the 126 MB L2
  sc.band = (int)(((int64_t)48 << 20) / ((int64_t)256 * K * 2));
  if (sc.band < 8) sc.band = 8;
  if (sc.band > sc.tiles_m) sc.band = sc.tiles_m;
  sc.n_pairs = g2_num_sms() / 2;
  if (sc.n_pairs > n_tiles && workspace == nullptr) sc.n_pairs = n_tiles;
  sc.dp_waves = (n_tiles + sc.n_pairs - 1) / sc.n_pairs;
  sc.sk_tiles = 0;
  float* ws_partial = nullptr;
  int* ws_flags = nullptr;
  const int rem = n_tiles % sc.n_pairs;
  if (workspace != nullptr && rem != 0) {
    // Stream-K over the LAST FULL WAVE + THE REMAINDER (every pair gets between one and two tiles' worth of k-blocks, so a tile is
    // cut at most once: one parked partial per pair, one gathered per finisher), everything before it in lockstep whole tiles.
    // With fewer tiles than pairs the whole problem is cut, as long as a tile ends up in at most ~4 pieces.  Worth it when the
    // k-block times saved exceed the fix-up (parking + gathering 256 KB per pair through L2 at the end: ~20 k-block times measured).
    int64_t gain;
    int sk_tiles, dp_waves;
    if (n_tiles > sc.n_pairs) {
      sk_tiles = rem + sc.n_pairs;
      dp_waves = n_tiles / sc.n_pairs - 1;
      gain = (int64_t)(sc.n_pairs - rem) * sc.KB / sc.n_pairs;
    } else {
      sk_tiles = n_tiles;
      dp_waves = 0;
      const int64_t sk_len = ((int64_t)n_tiles * sc.KB + sc.n_pairs - 1) / sc.n_pairs;
      gain = sk_len * 4 >= sc.KB ? sc.KB - sk_len : 0;
    }
    if (gain > 20) {
      sc.dp_waves = dp_waves;
      sc.sk_tiles = sk_tiles;
      ws_partial = (float*)workspace;
      ws_flags = (int*)((uint8_t*)workspace + (int64_t)(g2_num_sms() / 2 + 1) * G2_SLOT_FLOATS * sizeof(float));
    }
  }
  if (sc.sk_tiles == 0 && sc.n_pairs > n_tiles) {
    sc.n_pairs = n_tiles;
    sc.dp_waves = 1;
  }
  const int n_pairs = sc.n_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * n_pairs);
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = dl_pdl_enabled(2);
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm2_kernel<EPI>, tx, tw, to, (const bf16*)bias, (const bf16*)residual, M, N_out, sc, up_row0,
                                      ws_partial, ws_flags);
  if (le != cudaSuccess) return dl_set_error((int)le, "dl_linear(launch)");
  DL_CHECK_LAUNCH("dl_linear");
}

extern "C" int dl_linear(const void* x, const void* w, const void* bias, const void* residual, void* out, int M, int N, int K, int act,
                         void* workspace, cudaStream_t stream) {
  DL_REQUIRE(M >= 1 && N >= 8 && N % 8 == 0 && K >= G2_BK && K % G2_BK == 0 && act >= 0 && act <= 2 && (residual == nullptr || act == 0),
             "dl_linear");
  if (residual != nullptr) return g2_launch<G2_EPI_RESIDUAL>(x, w, bias, residual, out, M, N, N, K, 0, workspace, stream);
  if (act == 1) return g2_launch<G2_EPI_GELU>(x, w, bias, nullptr, out, M, N, N, K, 0, workspace, stream);
  if (act == 2) return g2_launch<G2_EPI_QUICK_GELU>(x, w, bias, nullptr, out, M, N, N, K, 0, workspace, stream);
  return g2_launch<G2_EPI_NONE>(x, w, bias, nullptr, out, M, N, N, K, 0, workspace, stream);
}

extern "C" int dl_linear_swiglu(const void* x, const void* w_gate_up, void* act_out, int M, int I, int K, void* workspace,
                                cudaStream_t stream) {
  DL_REQUIRE(M >= 1 && I % 8 == 0 && K >= G2_BK && K % G2_BK == 0, "dl_linear_swiglu");
  return g2_launch<G2_EPI_SWIGLU>(x, w_gate_up, nullptr, nullptr, act_out, M, 2 * I, I, K, I, workspace, stream);
}
