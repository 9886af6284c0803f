// K2 (v4): variable-length prefill attention — PERSISTENT tcgen05 pipeline.
// Contract: include/dynllava_b200.h (dl_attn_prefill_tc4 / dl_attn_packed_tc4); reference op:
// F.scaled_dot_product_attention, llava/model/language_model/dynamic_modeling_llama.py:1114-1122 (causal, head_dim 128)
// and custom_transformer_layer.py:153-180 / HF CLIPAttention (non-causal, head_dim 64).
//
// Work item = (256 query rows = two 128-row groups A and B, one head, one sample).  CTA = 352 threads, one per SM:
//   warp 0      lane 0: Q + K producer, lane 1: V producer (TMA, 128-key tiles in 2-stage rings, HD/64 SW128 boxes per tile);
//   warp 1, 10  MMA issuers, one elected thread per group.  Per 64-key half j:  PV(j) S(j+2);
//   warps 2-5   softmax group A, warps 6-9 group B: one thread per query row (TMEM lane == row).
// TMEM (all 512 columns): S_A0 [0,64) S_A1 [64,128) S_B0 [128,192) S_B1 [192,256) O_A [256,256+HD) O_B [256+HD,256+2HD).
//   * 64-key HALF tiles with two S buffers per group: S(j+1) is complete in TMEM when softmax(j) ends and S(j+2) is issued right
//     behind PV(j), so neither the softmax warps nor the tensor pipe wait for the other's latency (v2's serial chain);
//   * S is read ONCE into registers, P = exp2(S*scale - m) is written back as packed bf16 over the first 32 columns of the same
//     buffer and consumed from TMEM as the A operand of the PV MMA; O accumulates in TMEM, V is an MN-major B operand in place;
//   * the running max is only refreshed (and O rescaled) when it grows by more than 2^8 (exact after the final 1/l);
//   * EXP_POLY: one exp2 in four on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, rel. error 7.5e-5);
//   * O leaves through shared memory: 32-row x 64-column boxes (128B swizzle) written with TMA stores.  A row-per-thread store
//     straight from registers is 32 scattered 16-byte requests per instruction and cost ~16 k cycles per item.
// Persistence: grid = #SMs; all mbarriers live for the whole kernel and are indexed by running counters, so the roles drift across
// item boundaries:
//   * the producers load the next item's Q (per group, once that group's issuer has released it) and K/V tiles while the softmax
//     warps are still in the current item's last halves and epilogue;
//   * the issuers compute S(0), S(1) of the next item under the current epilogue; only PV(0) waits for o_free (O read out);
//   * every group releases every K/V stage of every item exactly once (tcgen05.commit, also for tiles it did not read), which
//     keeps the per-group empty barriers in step with the producers' running tile counter;
//   * every parity wait targets a phase that can be at most ONE ahead of the waiter (the anti-phase start of group B uses its own
//     one-shot barrier for that reason).
// Schedule (t4_pos_*): (sample, head) pairs in L2-sized windows, items sorted heaviest-first inside a window and dealt out as a snake.
// History and measurements: profiles/r1_ncu_attn_prefill_tc3.md (non-persistent v3, source in the git history),
// profiles/r1_ncu_attn_prefill_tc4.md.
#include <cuda.h>

#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define T4_THREADS 352

#ifdef DL_T4_TRACE
// debug build only (tools/trace_attn.py --tc4): per-item clock stamps of CTA 0: slot = group * 4 + kind, index = item ordinal
__device__ long long* g_t4_trace = nullptr;
#define T4_STAMP(slot, idx)                                                    \
  do {                                                                         \
    if (g_t4_trace && blockIdx.x == 0 && (idx) < 64 && (idx) >= 0) g_t4_trace[(slot) * 64 + (idx)] = clock64(); \
  } while (0)
extern "C" int dl_debug_t4_set_trace(long long* p) { return (int)cudaMemcpyToSymbol(g_t4_trace, &p, sizeof(p)); }
#else
#define T4_STAMP(slot, idx) \
  do {                      \
  } while (0)
#endif
#define T4_RESCALE_LOG2 8.0f

template <int HD>
struct T4Smem {
  static constexpr int TILE = 128 * HD * 2;  // one 128-row operand tile = HD/64 blocks of 16 KB (128B-swizzled)
  alignas(1024) uint8_t q[2][TILE];
  alignas(1024) uint8_t k[2][TILE];
  alignas(1024) uint8_t v[2][TILE];
  alignas(1024) uint8_t ostage[8][4096];  // per softmax warp: 32 rows x 64 bf16 of O, 128B-swizzled, source of the TMA store
  alignas(8) uint64_t q_full[2];      // per group: the groups drift apart by the causal imbalance (two halves per item)
  alignas(8) uint64_t q_empty[2];     // tcgen05.commit of the group's issuer behind its last S of the item
  alignas(8) uint64_t k_full[2];
  alignas(8) uint64_t v_full[2];
  alignas(8) uint64_t k_empty[2][2];  // [group][stage]
  alignas(8) uint64_t v_empty[2][2];
  alignas(8) uint64_t s_full[2][2];   // [group][running half & 1]
  alignas(8) uint64_t p_full[2][2];
  alignas(8) uint64_t o_done[2][2];
  alignas(8) uint64_t o_free[2];      // count 128: the group's softmax threads have read O out of TMEM
  alignas(8) uint64_t stagger;        // count 128, ONE phase only: group A has stored its first P of the kernel
  uint32_t tmem_base;
  int32_t cu_s[64];                   // CSR header and past lengths of up to 63 samples: every role re-derives every item, and
  int32_t past_s[64];                 // two dependent global loads per item per role were ~1 k cycles at each item boundary
};

DL_DEVINL uint64_t t4_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D[tmem] (+)= A[tmem] . B[smem]
DL_DEVINL void t4_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
DL_DEVINL void t4_tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
DL_DEVINL float t4_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA/ALU pipes: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial, n added to
// the exponent field.  x is clamped at -126 so the result never leaves the normal range (2^-126 rounds to 0 in the sum).
DL_DEVINL float t4_ex2_poly(float x) {
  const float xc = fmaxf(x, -126.f);
  const float t = xc + 12582912.f;  // 1.5 * 2^23: the low mantissa bits of t hold n
  const float f = xc - (t - 12582912.f);
  float p = fmaf(0.055171650f, f, 0.24261113f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992806f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}


// shared -> global tiled store through a tensor map (async proxy, bulk group of the issuing thread)
DL_DEVINL void t4_tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(smem_u32(smem_src)),
               "r"(c0), "r"(c1)
               : "memory");
}
DL_DEVINL void t4_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DL_DEVINL void t4_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// one work item = (256-row query block mt, head h, sample b); identical arithmetic in every role
struct T4Item {
  int h, q_start, q_len, m0, off, kv_total, n_a, n_b, n_h, n_t, n_ta, n_tb, kv_row0, kv_col0;
  bool b_active;
};
// Item schedule.  The (sample, head) pairs are taken in WINDOWS of Wn = 2*gridDim.x / n_qblk pairs (their K/V, Wn * 2 * len * HD * 2
// bytes, stay in the 126 MB L2 while all their query blocks are processed: a globally heaviest-first list re-read K/V from HBM
// 3x, profiles/r1_ncu_attn_prefill_tc4.md).  Inside a window the items are sorted heaviest block first and dealt out in a snake:
// CTA c takes positions c, 2G-1-c, 2G+c, ...; with two items per CTA the pair is (block n-1-i, block i): equal causal work.
struct T4Pos {
  int step, n_steps;  // step = window * slots_per_window + slot
  int wnd, slot;      // tracked incrementally (no division per step)
  int Wn, spw, HB;    // pairs per window, slots per window, heads * samples
  float inv_heads;
  int blk, b, h;      // decoded item: query block rank (0 = heaviest), sample, head
};
// floor(a / d) for 0 <= a < 2^22, d > 0 through the float reciprocal (exact: (a + 0.5) / d is >= 0.5 / d away from an integer)
DL_DEVINL int t4_div(int a, float inv_d) { return (int)(((float)a + 0.5f) * inv_d); }
DL_DEVINL void t4_pos_init(T4Pos& p, int heads, int Bn, int n_qblk) {
  const int G = gridDim.x;
  p.HB = heads * Bn;
  p.Wn = max(1, min(p.HB, (2 * G) / n_qblk));
  p.spw = (p.Wn * n_qblk + G - 1) / G;
  p.n_steps = ((p.HB + p.Wn - 1) / p.Wn) * p.spw;
  p.step = p.wnd = p.slot = 0;
  p.inv_heads = 1.0f / (float)heads;
}
DL_DEVINL void t4_pos_next(T4Pos& p) {
  ++p.step;
  if (++p.slot == p.spw) {
    p.slot = 0;
    ++p.wnd;
  }
}
// decode the current step for this CTA; false = no item in this slot
DL_DEVINL bool t4_pos_decode(T4Pos& p, int heads, int n_qblk) {
  const int G = gridDim.x, c = blockIdx.x;
  const int hb0 = p.wnd * p.Wn;
  const int wn = min(p.Wn, p.HB - hb0);  // pairs in this window
  const int pos = (p.slot & 1) ? (p.slot + 1) * G - 1 - c : p.slot * G + c;
  if (pos >= wn * n_qblk) return false;
  p.blk = t4_div(pos, 1.0f / (float)wn);
  const int hb = hb0 + (pos - p.blk * wn);
  p.b = t4_div(hb, p.inv_heads);
  p.h = hb - p.b * heads;
  return true;
}
template <int HD, bool CAUSAL, bool PACKED>
DL_DEVINL bool t4_item(const T4Pos& p, int n_qblk, int heads, int cap, const int32_t* __restrict__ cu_q,
                       const int32_t* __restrict__ kv_past, T4Item& I) {
  // rank 0 = the heaviest block: the last one when causal, any full one otherwise (the ragged last block is the lightest)
  const int mt = CAUSAL ? n_qblk - 1 - p.blk : p.blk;
  const int b = p.b;
  I.h = p.h;
  I.q_start = cu_q[b];
  I.q_len = cu_q[b + 1] - I.q_start;
  I.m0 = mt * 256;
  if (I.m0 >= I.q_len) return false;
  const int past = kv_past ? kv_past[b] : 0;
  I.off = past;
  I.kv_total = past + I.q_len;
  I.b_active = (I.m0 + 128) < I.q_len;
  I.n_a = ((CAUSAL ? min(I.kv_total, I.m0 + 128 + past) : I.kv_total) + 63) / 64;
  I.n_b = I.b_active ? ((CAUSAL ? min(I.kv_total, I.m0 + 256 + past) : I.kv_total) + 63) / 64 : 0;
  I.n_h = max(I.n_a, I.n_b);
  I.n_t = (I.n_h + 1) >> 1;
  I.n_ta = (I.n_a + 1) >> 1;
  I.n_tb = (I.n_b + 1) >> 1;
  I.kv_row0 = PACKED ? I.q_start : (b * heads + I.h) * cap;
  I.kv_col0 = PACKED ? I.h * HD : 0;
  return true;
}

template <int HD, bool CAUSAL, bool PACKED, bool EXP_POLY>
__global__ void __launch_bounds__(T4_THREADS, 1) attn_prefill_tc4_kernel(
    const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o, bf16* __restrict__ out,
    int64_t o_tok_stride, const int32_t* __restrict__ cu_q, const int32_t* __restrict__ kv_past, int heads, int Bn, int cap,
    int n_qblk, int n_items, float scale_log2) {
  constexpr int TILE_BYTES = T4Smem<HD>::TILE;
  constexpr int NBOX = HD / 64;    // 64-column TMA boxes per operand tile
  constexpr int KSTEPS = HD / 16;  // K = 16 MMAs of S = Q K^T
  constexpr uint32_t O_COL = 256;  // first O column
  extern __shared__ uint8_t smem_raw[];
  T4Smem<HD>& S = *reinterpret_cast<T4Smem<HD>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&S.q_full[s], 1);
      mbar_init(&S.q_empty[s], 1);
      mbar_init(&S.k_full[s], 1);
      mbar_init(&S.v_full[s], 1);
      mbar_init(&S.o_free[s], 128);
      if (s == 0) mbar_init(&S.stagger, 128);
      for (int u = 0; u < 2; ++u) {
        mbar_init(&S.k_empty[s][u], 1);
        mbar_init(&S.v_empty[s][u], 1);
        mbar_init(&S.s_full[s][u], 1);
        mbar_init(&S.p_full[s][u], 128);
        mbar_init(&S.o_done[s][u], 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&S.tmem_base);
  const bool hdr_in_smem = Bn < 64;
  if (hdr_in_smem && threadIdx.x >= 64 && threadIdx.x < 64 + 64) {
    const int i = threadIdx.x - 64;
    if (i <= Bn) S.cu_s[i] = cu_q[i];
    S.past_s[i] = (kv_past != nullptr && i < Bn) ? kv_past[i] : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  if (hdr_in_smem) {
    cu_q = S.cu_s;
    kv_past = S.past_s;
  }
  T4Item I;
  T4Pos pos;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------- Q + K producer -------------------------------
      int T = 0, ia[2] = {0, 0};  // 128-key tiles loaded so far; items each group has worked on
      for (t4_pos_init(pos, heads, Bn, n_qblk); pos.step < pos.n_steps; t4_pos_next(pos)) {
        if (!t4_pos_decode(pos, heads, n_qblk) || !t4_item<HD, CAUSAL, PACKED>(pos, n_qblk, heads, cap, cu_q, kv_past, I)) continue;
        auto load_q = [&](int g) {
          if (ia[g] >= 1) mbar_wait(&S.q_empty[g], (ia[g] - 1) & 1);  // the group's last S of its previous item has read the tile
          mbar_expect_tx(&S.q_full[g], TILE_BYTES);
#pragma unroll
          for (int x = 0; x < NBOX; ++x)
            tma_load_2d(S.q[g] + x * 16384, &tmap_q, I.h * HD + x * 64, I.q_start + I.m0 + g * 128, &S.q_full[g]);
          ++ia[g];
        };
        auto load_k = [&](int t) {
          const int s = T & 1;
          if (T >= 2) {  // stage s still holds tile T-2: both groups release every tile
            mbar_wait(&S.k_empty[0][s], ((T >> 1) & 1) ^ 1);
            mbar_wait(&S.k_empty[1][s], ((T >> 1) & 1) ^ 1);
          }
          mbar_expect_tx(&S.k_full[s], TILE_BYTES);
#pragma unroll
          for (int x = 0; x < NBOX; ++x)
            tma_load_2d(S.k[s] + x * 16384, &tmap_k, I.kv_col0 + x * 64, I.kv_row0 + t * 128, &S.k_full[s]);
          ++T;
        };
        // group A finishes an item two halves before group B: give it its Q tile and the first K tile before waiting for B
        load_q(0);
        load_k(0);
        if (I.b_active) load_q(1);
        for (int t = 1; t < I.n_t; ++t) load_k(t);
      }
    } else if (lane == 1) {
      // ------------------------------- V producer -------------------------------
      int T = 0;
      for (t4_pos_init(pos, heads, Bn, n_qblk); pos.step < pos.n_steps; t4_pos_next(pos)) {
        if (!t4_pos_decode(pos, heads, n_qblk) || !t4_item<HD, CAUSAL, PACKED>(pos, n_qblk, heads, cap, cu_q, kv_past, I)) continue;
        for (int t = 0; t < I.n_t; ++t, ++T) {
          const int s = T & 1;
          if (T >= 2) {
            mbar_wait(&S.v_empty[0][s], ((T >> 1) & 1) ^ 1);
            mbar_wait(&S.v_empty[1][s], ((T >> 1) & 1) ^ 1);
          }
          mbar_expect_tx(&S.v_full[s], TILE_BYTES);
#pragma unroll
          for (int x = 0; x < NBOX; ++x)
            tma_load_2d(S.v[s] + x * 16384, &tmap_v, I.kv_col0 + x * 64, I.kv_row0 + t * 128, &S.v_full[s]);
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ------------------------------- MMA issuers: warp 1 -> group A, warp 10 -> group B -------------------------------
    const int g = warp == 1 ? 0 : 1;
    if (elect_one()) {  // converged warp + elect.sync keeps the UTCHMMA issue on the uniform datapath
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, HD, 0, 1);  // N = head_dim; B (= V) MN-major
      const uint32_t qa = smem_u32(S.q[g]);
      const uint32_t ts = tmem + (uint32_t)g * 128, to = tmem + O_COL + (uint32_t)g * HD;
      int it = 0, T0 = 0, J = 0, ia = 0;  // items seen, tiles before this item, halves of this group so far, items this group worked on
      for (t4_pos_init(pos, heads, Bn, n_qblk); pos.step < pos.n_steps; t4_pos_next(pos)) {
        if (!t4_pos_decode(pos, heads, n_qblk) || !t4_item<HD, CAUSAL, PACKED>(pos, n_qblk, heads, cap, cu_q, kv_past, I)) continue;
        const int n_g = g == 0 ? I.n_a : I.n_b;
        const int n_tg = g == 0 ? I.n_ta : I.n_tb;
        if (n_g > 0) mbar_wait(&S.q_full[g], ia & 1);
        // S(j) = Q_g . K[64 keys of half j]^T  ->  S buffer (J + j) & 1
        auto issue_s = [&](int j) {
          const int T = T0 + (j >> 1), s = T & 1, jj = J + j;
          mbar_wait(&S.k_full[s], (T >> 1) & 1);
          tc_fence_after();
          const uint32_t ka = smem_u32(S.k[s]) + (uint32_t)(j & 1) * 8192u;  // rows 64.. of each box
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk) {
            const uint32_t o = (kk >> 2) * 16384 + (kk & 3) * 32;
            umma_bf16_ss(ts + (uint32_t)(jj & 1) * 64, umma_desc_kmajor_sw128(qa + o), umma_desc_kmajor_sw128(ka + o), IDESC_QK,
                         kk > 0 ? 1u : 0u);
          }
          umma_commit(&S.s_full[g][jj & 1]);
          if ((j & 1) || j == n_g - 1) umma_commit(&S.k_empty[g][s]);  // last half of this tile the group reads
          if (j == n_g - 1) umma_commit(&S.q_empty[g]);                // last read of this item's Q tile
        };
        // O_g += P(j) . V[64 keys of half j]
        auto issue_pv = [&](int j) {
          const int T = T0 + (j >> 1), s = T & 1, jj = J + j;
          mbar_wait(&S.p_full[g][jj & 1], (jj >> 1) & 1);
          mbar_wait(&S.v_full[s], (T >> 1) & 1);
          tc_fence_after();
          const uint32_t va = smem_u32(S.v[s]) + (uint32_t)(j & 1) * 8192u;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // 16 keys per MMA = 8 packed P columns and 2 KB of V rows
            t4_umma_ts(to, ts + (uint32_t)(jj & 1) * 64 + kk * 8, t4_desc_mn_sw128(va + kk * 2048, 16384), IDESC_PV,
                       (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&S.o_done[g][jj & 1]);
          if ((j & 1) || j == n_g - 1) umma_commit(&S.v_empty[g][s]);
        };
        if (n_g > 0) {
          // Anti-phase start (first item only): group B begins one softmax pass after group A.  This must be a one-shot barrier:
          // waiting on p_full[0][0] by parity deadlocks if this thread arrives after group A's THIRD half (phase 1 complete
          // looks like phase 0 pending, and group A cannot reach phase 2 without the K/V stages group B has yet to release) —
          // seen once in ~20 full-model prefills.
          if (it == 0 && g == 1) mbar_wait(&S.stagger, 0);
          issue_s(0);
          if (n_g > 1) issue_s(1);
          if (ia >= 1) {  // O of the previous item must have been read out before PV(0) overwrites it
            mbar_wait(&S.o_free[g], (ia - 1) & 1);
            tc_fence_after();
          }
          for (int j = 0; j < n_g; ++j) {
            issue_pv(j);
            if (j + 2 < n_g) issue_s(j + 2);
          }
          J += n_g;
          ++ia;
        }
        // tiles of this item the group did not read are released too (behind everything issued so far), so that the
        // per-group empty barriers advance exactly once per tile
        for (int t = n_tg; t < I.n_t; ++t) {
          const int T = T0 + t, s = T & 1;
          mbar_wait(&S.k_full[s], (T >> 1) & 1);
          umma_commit(&S.k_empty[g][s]);
          mbar_wait(&S.v_full[s], (T >> 1) & 1);
          umma_commit(&S.v_empty[g][s]);
        }
        T0 += I.n_t;
        ++it;
      }
    }
  } else {
    // ------------------------------- softmax groups -------------------------------
    const int g = (warp - 2) >> 2;
    const int r = (warp & 3) * 32 + lane;  // TMEM lane quarter of this warp -> row inside the tile
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t ts = tmem + (uint32_t)g * 128 + lane_addr;
    const uint32_t to = tmem + O_COL + (uint32_t)g * HD + lane_addr;
    int J = 0;
    [[maybe_unused]] int ord = 0;
    for (t4_pos_init(pos, heads, Bn, n_qblk); pos.step < pos.n_steps; t4_pos_next(pos)) {
      if (!t4_pos_decode(pos, heads, n_qblk) || !t4_item<HD, CAUSAL, PACKED>(pos, n_qblk, heads, cap, cu_q, kv_past, I)) continue;
      const int n_g = g == 0 ? I.n_a : I.n_b;
      if (n_g == 0) continue;
      if (r == 0) T4_STAMP(g * 4 + 0, ord);  // item entered
      const int qi = I.m0 + g * 128 + r;
      const int key_lim = CAUSAL ? min(I.kv_total - 1, qi + I.off) : I.kv_total - 1;
      float m_ref = -INFINITY, l_run = 0.f;

      for (int j = 0; j < n_g; ++j) {
        const int jj = J + j;
        const uint32_t tsj = ts + (uint32_t)(jj & 1) * 64;
        mbar_wait(&S.s_full[g][jj & 1], (jj >> 1) & 1);
        tc_fence_after();
        if (r == 0 && j == 0) T4_STAMP(g * 4 + 1, ord);  // first S of the item seen
        // S is read ONCE.  (Issuing the tcgen05.ld of S(j+1) early into registers this half has consumed was measured 4 % slower,
        // profiles/r1_ncu_attn_prefill_tc4.md: the ~230-cycle load is already hidden by the other group's warp on the same SMSP.)
        uint32_t v[64];
        tmem_ld_32x32b_x32(tsj, v);
        tmem_ld_32x32b_x32(tsj + 32, v + 32);
        tmem_ld_wait();
        const int kbase = j * 64;
        if (kbase + 63 > key_lim) {
#pragma unroll
          for (int e = 0; e < 64; ++e)
            if (kbase + e > key_lim) v[e] = 0xff800000u;  // -inf
        }
        // four independent max chains (a single fmaxf chain is 32 dependent FMNMX3 with nothing to overlap them)
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int e = 0; e < 64; e += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], __uint_as_float(v[e + u]));
        }
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        mx *= scale_log2;
        float alpha = 1.f;
        bool rescale = false;
        if (j == 0) {
          m_ref = (mx == -INFINITY) ? 0.f : mx;
        } else if (mx > m_ref + T4_RESCALE_LOG2) {
          alpha = t4_ex2(m_ref - mx);
          m_ref = mx;
          rescale = true;
        }
        float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t pk[16];
#pragma unroll
          for (int e = hh * 32; e < hh * 32 + 32; e += 2) {
            const float x0 = fmaf(__uint_as_float(v[e]), scale_log2, -m_ref);
            const float x1 = fmaf(__uint_as_float(v[e + 1]), scale_log2, -m_ref);
            const float p0 = (EXP_POLY && (e & 6) == 0) ? t4_ex2_poly(x0) : t4_ex2(x0);  // elements 0 of every 8 ...
            const float p1 = (EXP_POLY && (e & 6) == 4) ? t4_ex2_poly(x1) : t4_ex2(x1);  // ... and 5 of every 8: 1 in 4
            ls0 += p0;
            ls1 += p1;
            bf16x2 t = f2bf2(p0, p1);
            pk[(e >> 1) & 15] = *reinterpret_cast<uint32_t*>(&t);
          }
          tmem_st_32x32b_x16(tsj + hh * 16, pk);  // P overwrites S columns this thread has already read; the first store
        }                                         // drains while the second 32 exponentials are evaluated
        l_run = l_run * alpha + (ls0 + ls1);
        if (j > 0 && __any_sync(0xffffffffu, rescale)) {
          // rare: bring O to the new reference max before PV_j accumulates onto it
          mbar_wait(&S.o_done[g][(jj - 1) & 1], ((jj - 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(to + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
            t4_tmem_st_x32(to + c * 32, ov);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&S.p_full[g][jj & 1]);
        if (g == 0 && jj == 0) mbar_arrive(&S.stagger);  // group A's first P of the kernel
      }
      if (r == 0) T4_STAMP(g * 4 + 2, ord);  // last P of the item stored
      // epilogue (the previous phase of this barrier, PV(jl - 2), is known complete: S(jl) was issued behind it)
      const int jl = J + n_g - 1;
      mbar_wait(&S.o_done[g][jl & 1], (jl >> 1) & 1);
      tc_fence_after();
      const float inv = 1.f / l_run;
      // O leaves through shared memory: a row-per-thread store straight from registers is 32 scattered 16-byte requests per
      // instruction and cost ~16 k cycles per item (tools/trace_attn4.py); staged 32-row x 64-column boxes go out as one TMA
      // store each.  Warps whose 32 rows are not all inside the sample (ragged tail) keep the direct path: a TMA box would
      // clip at the end of the packed buffer, not at the end of the sample.
      const int wrow0 = I.m0 + g * 128 + (warp & 3) * 32;
      const bool warp_full = wrow0 + 32 <= I.q_len;
      uint8_t* stage = S.ostage[warp - 2];
      bf16* dst = out + (int64_t)(I.q_start + qi) * o_tok_stride + (int64_t)I.h * HD;
#pragma unroll
      for (int hh = 0; hh < HD / 64; ++hh) {
        uint32_t ov[64];
        tmem_ld_32x32b_x32(to + hh * 64, ov);
        tmem_ld_32x32b_x32(to + hh * 64 + 32, ov + 32);
        tmem_ld_wait();
        if (hh == HD / 64 - 1) {  // O is in registers: the issuer may start the next item's PV(0)
          tc_fence_before();
          mbar_arrive(&S.o_free[g]);
        }
        if (warp_full) {
          bf16x8 pk8[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(ov[c * 8 + u]) * inv;
            pk8[c] = pack8(f);
          }
          if (lane == 0) t4_bulk_wait_read0();  // the previous box has left the staging tile
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<bf16x8*>(stage + lane * 128 + ((c ^ (lane & 7)) << 4)) = pk8[c];
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            t4_tma_store_2d(&tmap_o, stage, I.h * HD + hh * 64, I.q_start + wrow0);
            t4_bulk_commit();
          }
        } else if (qi < I.q_len) {
#pragma unroll
          for (int e = 0; e < 64; e += 8) {
            float f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(ov[e + u]) * inv;
            st8(dst + hh * 64 + e, pack8(f));
          }
        }
      }
      if (r == 0) {
        T4_STAMP(g * 4 + 3, ord);  // epilogue done
        T4_STAMP(8 + g, ord) ;
#ifdef DL_T4_TRACE
        if (g_t4_trace && blockIdx.x == 0 && ord < 64) g_t4_trace[(8 + g) * 64 + ord] = n_g;
#endif
      }
      ++ord;
      J += n_g;
    }
    if (lane == 0) t4_bulk_wait_read0();  // staged boxes must have been read before the CTA's shared memory goes away
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

static int t4_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int HD, bool CAUSAL, bool PACKED, bool EXP_POLY>
static int launch_t4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, bf16* out, int64_t o_tok_stride, int n_tok,
                     const int32_t* cu_q, const int32_t* kv_past, int B, int heads, int cap, int max_q_len, float scale,
                     cudaStream_t stream) {
  alignas(64) CUtensorMap to;  // O as [n_tok, heads*HD], boxes of 32 rows x 64 columns
  int rc = dl_make_tmap_2d(&to, out, (uint64_t)n_tok, (uint64_t)heads * HD, (uint64_t)o_tok_stride, 32, 64);
  if (rc) return rc;
  const int smem = (int)sizeof(T4Smem<HD>) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_prefill_tc4_kernel<HD, CAUSAL, PACKED, EXP_POLY>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_attn_tc4(smem attr)");
    attr_set = true;
  }
  const int n_qblk = (max_q_len + 255) / 256;
  const int n_items = n_qblk * heads * B;
  const int grid = n_items < t4_sm_count() ? n_items : t4_sm_count();
  attn_prefill_tc4_kernel<HD, CAUSAL, PACKED, EXP_POLY><<<grid, T4_THREADS, smem, stream>>>(
      tq, tk, tv, to, out, o_tok_stride, cu_q, kv_past, heads, B, cap, n_qblk, n_items, scale * 1.4426950408889634f);
  DL_CHECK_LAUNCH("dl_attn_tc4");
}

// flags bit 0: evaluate one exp2 in four with the FMA-pipe polynomial
extern "C" int dl_attn_prefill_tc4(const void* q, int64_t q_tok_stride, const void* k_arena, const void* v_arena, void* out,
                                   int64_t o_tok_stride, const int32_t* cu_q, const int32_t* kv_past, int n_tok, int B, int heads,
                                   int cap, int max_q_len, float scale, int flags, cudaStream_t stream) {
  if (B == 0 || max_q_len == 0) return 0;
  alignas(64) CUtensorMap tq, tk, tv;
  int rc = dl_make_tmap_2d(&tq, q, (uint64_t)n_tok, (uint64_t)heads * 128, (uint64_t)q_tok_stride, 128, 64);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tk, k_arena, (uint64_t)B * heads * cap, 128, 128, 128, 64);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tv, v_arena, (uint64_t)B * heads * cap, 128, 128, 128, 64);
  if (rc) return rc;
  if (flags & 1)
    return launch_t4<128, true, false, true>(tq, tk, tv, (bf16*)out, o_tok_stride, n_tok, cu_q, kv_past, B, heads, cap, max_q_len, scale, stream);
  return launch_t4<128, true, false, false>(tq, tk, tv, (bf16*)out, o_tok_stride, n_tok, cu_q, kv_past, B, heads, cap, max_q_len, scale, stream);
}

// K6/K7: non-causal attention over a fused projection qkv [n_tok, 3*heads*64] = [q | k | v] (HF CLIPAttention;
// custom_transformer_layer.py:153-180), per-sample token ranges from the CSR header cu_q.  out [n_tok, heads*64].
extern "C" int dl_attn_packed_tc4(const void* qkv, void* out, const int32_t* cu_q, int n_tok, int B, int heads, int head_dim,
                                  int max_q_len, float scale, int flags, cudaStream_t stream) {
  DL_REQUIRE(head_dim == 64, "dl_attn_packed_tc4");
  if (B == 0 || max_q_len == 0) return 0;
  const int64_t C = (int64_t)heads * 64;
  const bf16* base = (const bf16*)qkv;
  alignas(64) CUtensorMap tq, tk, tv;
  int rc = dl_make_tmap_2d(&tq, base, (uint64_t)n_tok, (uint64_t)C, (uint64_t)(3 * C), 128, 64);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tk, base + C, (uint64_t)n_tok, (uint64_t)C, (uint64_t)(3 * C), 128, 64);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tv, base + 2 * C, (uint64_t)n_tok, (uint64_t)C, (uint64_t)(3 * C), 128, 64);
  if (rc) return rc;
  if (flags & 1)
    return launch_t4<64, false, true, true>(tq, tk, tv, (bf16*)out, C, n_tok, cu_q, nullptr, B, heads, 0, max_q_len, scale, stream);
  return launch_t4<64, false, true, false>(tq, tk, tv, (bf16*)out, C, n_tok, cu_q, nullptr, B, heads, 0, max_q_len, scale, stream);
}
