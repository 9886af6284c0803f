// K2 (v3): variable-length prefill attention — tcgen05 pipeline with DOUBLE-BUFFERED score tiles.
// Same contract as dl_attn_prefill_tc2 / dl_attn_packed_tc (attn_prefill_tc2.cu); reference op:
// F.scaled_dot_product_attention, llava/model/language_model/dynamic_modeling_llama.py:1114-1122 (causal, head_dim 128)
// and custom_transformer_layer.py:153-180 / HF CLIPAttention (non-causal, head_dim 64).
//
// v2's measured limit (profiles/r1_ncu_attn_prefill_tc2.md): each softmax group's loop was one serial chain
// softmax(j) -> PV(j) -> S(j+1) -> softmax(j+1): a group idled while the tensor pipe produced its next S, and the tensor
// pipe idled while the group digested it.  v3 breaks the chain by working on 64-key HALF tiles with two S buffers per
// group: S(j+1) is already complete in TMEM when softmax(j) ends, and S(j+2) is issued right behind PV(j).
//
// One CTA = (256 query rows = two 128-row tiles A and B, one head, one sample); 384 threads:
//   warp 0      K producer (+Q): 128-key K tiles by TMA into a 2-stage ring (HD/64 SW128 boxes each);
//   warp 10     V producer: same for V (K and V stages are released at different times, so they have their own rings);
//   warp 1, 11  MMA issuers, one thread per group (a single issuer would park group A's ready MMAs behind its wait for
//               group B's P).  Per 64-key half j:  PV(j) S(j+2);
//   warps 2-5   softmax group A, warps 6-9 softmax group B: one thread per query row (TMEM lane == row).
// TMEM: S_A0 [0,64) S_A1 [64,128) S_B0 [128,192) S_B1 [192,256) O_A [256,256+HD) O_B [256+HD,256+2HD).
//   * S(j) lives in buffer j&1; it is read ONCE into registers, P = exp2(S*scale - m) is written back as packed bf16 over
//     the first 32 columns of the same buffer and consumed from TMEM as the A operand of the PV MMA;
//   * O accumulates in TMEM across halves, V is consumed in place as an MN-major B operand;
//   * the running max is only refreshed (and O rescaled) when it grows by more than 2^8 (exact after the final 1/l);
//   * EXP_POLY: one exp2 in four is evaluated on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial,
//     relative error 7.5e-5 << bf16's 2^-9) so that the 16-lane/clk SFU is no longer the co-bottleneck of the tensor pipe.
#include <cuda.h>

#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define T3_THREADS 384
#define T3_RESCALE_LOG2 8.0f

#ifdef DL_T3_TRACE
// debug build only (tools/trace_attn.py): clock stamps of CTA (x=gridDim.x-1,0,0) into a global buffer
__device__ long long* g_t3_trace = nullptr;
#define T3_STAMP(slot, idx)                                                                        \
  do {                                                                                             \
    if (g_t3_trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (idx) < 64 && (idx) >= 0) \
      g_t3_trace[(slot) * 64 + (idx)] = clock64();                                                 \
  } while (0)
extern "C" int dl_debug_t3_set_trace(long long* p) {
  return (int)cudaMemcpyToSymbol(g_t3_trace, &p, sizeof(p));
}
#else
#define T3_STAMP(slot, idx) \
  do {                      \
  } while (0)
#endif

template <int HD>
struct T3Smem {
  static constexpr int TILE = 128 * HD * 2;  // one 128-row operand tile = HD/64 blocks of 16 KB (128B-swizzled)
  alignas(1024) uint8_t q[2][TILE];
  alignas(1024) uint8_t k[2][TILE];
  alignas(1024) uint8_t v[2][TILE];
  alignas(8) uint64_t q_full;
  alignas(8) uint64_t k_full[2];
  alignas(8) uint64_t v_full[2];
  alignas(8) uint64_t k_empty[2][2];  // [group][stage]: each group's issuer releases the stages it has read
  alignas(8) uint64_t v_empty[2][2];
  alignas(8) uint64_t s_full[2][2];  // [group][buffer]
  alignas(8) uint64_t p_full[2][2];
  alignas(8) uint64_t o_done[2][2];  // [group][half parity]: with the S look-ahead a single barrier's parity would be ambiguous
  uint32_t tmem_base;
};

DL_DEVINL uint64_t t3_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D[tmem] (+)= A[tmem] . B[smem]
DL_DEVINL void t3_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
DL_DEVINL void t3_tmem_st_x32(uint32_t taddr, const uint32_t* r) {
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
DL_DEVINL float t3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA/ALU pipes: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial, n added to
// the exponent field.  x is clamped at -126 so the result never leaves the normal range (2^-126 rounds to 0 in the sum).
DL_DEVINL float t3_ex2_poly(float x) {
  const float xc = fmaxf(x, -126.f);
  const float t = xc + 12582912.f;  // 1.5 * 2^23: the low mantissa bits of t hold n
  const float f = xc - (t - 12582912.f);
  float p = fmaf(0.055171650f, f, 0.24261113f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992806f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int HD, bool CAUSAL, bool PACKED, bool EXP_POLY, bool STAGGER>
__global__ void __launch_bounds__(T3_THREADS, 1) attn_prefill_tc3_kernel(
    const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
    const __grid_constant__ CUtensorMap tmap_v, bf16* __restrict__ out, int64_t o_tok_stride,
    const int32_t* __restrict__ cu_q, const int32_t* __restrict__ kv_past, int heads, int cap, float scale_log2) {
  constexpr int TILE_BYTES = T3Smem<HD>::TILE;
  constexpr int NBOX = HD / 64;    // 64-column TMA boxes per operand tile
  constexpr int KSTEPS = HD / 16;  // K = 16 MMAs of S = Q K^T
  constexpr uint32_t O_COL = 256;  // first O column
  extern __shared__ uint8_t smem_raw[];
  T3Smem<HD>& S = *reinterpret_cast<T3Smem<HD>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal: the last query tiles see the most keys -> schedule them first (longest-processing-time order)
  const int mt = gridDim.x - 1 - blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q_start = cu_q[b];
  const int q_len = cu_q[b + 1] - q_start;
  const int m0 = mt * 256;
  if (m0 >= q_len) return;
  const int past = kv_past ? kv_past[b] : 0;
  const int kv_total = past + q_len;
  const int off = past;
  const bool b_active = (m0 + 128) < q_len;
  // work in 64-key halves
  const int n_a = ((CAUSAL ? min(kv_total, m0 + 128 + off) : kv_total) + 63) / 64;
  const int n_b = b_active ? ((CAUSAL ? min(kv_total, m0 + 256 + off) : kv_total) + 63) / 64 : 0;
  const int n_h = max(n_a, n_b);
  const int n_t = (n_h + 1) >> 1;  // 128-key K/V tiles
  // K/V rows: arena slab of (sample, head), or — PACKED — the sample's own token rows of the fused q|k|v buffer
  const int kv_row0 = PACKED ? q_start : (b * heads + h) * cap;
  const int kv_col0 = PACKED ? h * HD : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(&S.q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&S.k_full[s], 1);
      mbar_init(&S.v_full[s], 1);
      for (int u = 0; u < 2; ++u) {
        mbar_init(&S.o_done[s][u], 1);
        mbar_init(&S.k_empty[s][u], 1);
        mbar_init(&S.v_empty[s][u], 1);
        mbar_init(&S.s_full[s][u], 1);
        mbar_init(&S.p_full[s][u], 128);
      }
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&S.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  // 128-key K/V tiles each group touches (causal: group A stops one tile earlier than B at most)
  const int n_ta = (n_a + 1) >> 1, n_tb = (n_b + 1) >> 1;
  if (warp == 0) {
    // ------------------------------- Q + K producer -------------------------------
    if (elect_one()) {  // converged warp + elect.sync keeps the UTMALDG issue on the uniform datapath
      mbar_expect_tx(&S.q_full, (b_active ? 2 : 1) * TILE_BYTES);
#pragma unroll
      for (int x = 0; x < NBOX; ++x) {
        tma_load_2d(S.q[0] + x * 16384, &tmap_q, h * HD + x * 64, q_start + m0, &S.q_full);
        if (b_active) tma_load_2d(S.q[1] + x * 16384, &tmap_q, h * HD + x * 64, q_start + m0 + 128, &S.q_full);
      }
      for (int t = 0; t < n_t; ++t) {
        const int s = t & 1;
        if (t >= 2) {  // stage s still holds tile t-2: wait until every group that read it has released it
          if (t - 2 < n_ta) mbar_wait(&S.k_empty[0][s], ((t - 2) >> 1) & 1);
          if (t - 2 < n_tb) mbar_wait(&S.k_empty[1][s], ((t - 2) >> 1) & 1);
        }
        mbar_expect_tx(&S.k_full[s], TILE_BYTES);
#pragma unroll
        for (int x = 0; x < NBOX; ++x) tma_load_2d(S.k[s] + x * 16384, &tmap_k, kv_col0 + x * 64, kv_row0 + t * 128, &S.k_full[s]);
      }
    }
  } else if (warp == 10) {
    // ------------------------------- V producer -------------------------------
    if (elect_one()) {
      for (int t = 0; t < n_t; ++t) {
        const int s = t & 1;
        if (t >= 2) {
          if (t - 2 < n_ta) mbar_wait(&S.v_empty[0][s], ((t - 2) >> 1) & 1);
          if (t - 2 < n_tb) mbar_wait(&S.v_empty[1][s], ((t - 2) >> 1) & 1);
        }
        mbar_expect_tx(&S.v_full[s], TILE_BYTES);
#pragma unroll
        for (int x = 0; x < NBOX; ++x) tma_load_2d(S.v[s] + x * 16384, &tmap_v, kv_col0 + x * 64, kv_row0 + t * 128, &S.v_full[s]);
      }
    }
  } else if (warp == 1 || warp == 11) {
    // ------------------------------- MMA issuers: warp 1 -> group A, warp 11 -> group B -------------------------------
    // One issuing thread per group: neither group's MMAs ever queue behind a wait for the other group's P.
    const int g = warp == 1 ? 0 : 1;
    const int n_g = g == 0 ? n_a : n_b;
    if (n_g > 0 && elect_one()) {  // converged warp + elect.sync keeps the UTCHMMA issue on the uniform datapath
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, HD, 0, 1);  // N = head_dim; B (= V) MN-major
      const uint32_t qa = smem_u32(S.q[g]);
      const uint32_t ts = tmem + (uint32_t)g * 128, to = tmem + O_COL + (uint32_t)g * HD;
      // a 128-key stage is released once per group, behind the last MMA of the group that reads it
      auto last_half_of_tile = [&](int j) { return (j & 1) || j == n_g - 1; };
      // S(j) = Q_g . K[64 keys of half j]^T  ->  S buffer j & 1
      auto issue_s = [&](int j) {
        const int t = j >> 1, s = t & 1;
        mbar_wait(&S.k_full[s], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t ka = smem_u32(S.k[s]) + (uint32_t)(j & 1) * 8192u;  // rows 64.. of each box
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
          const uint32_t o = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_bf16_ss(ts + (uint32_t)(j & 1) * 64, umma_desc_kmajor_sw128(qa + o), umma_desc_kmajor_sw128(ka + o), IDESC_QK,
                       kk > 0 ? 1u : 0u);
        }
        umma_commit(&S.s_full[g][j & 1]);
        if (last_half_of_tile(j)) umma_commit(&S.k_empty[g][s]);
      };
      // O_g += P(j) . V[64 keys of half j]
      auto issue_pv = [&](int j) {
        const int t = j >> 1, s = t & 1;
        mbar_wait(&S.p_full[g][j & 1], (j >> 1) & 1);
        mbar_wait(&S.v_full[s], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t va = smem_u32(S.v[s]) + (uint32_t)(j & 1) * 8192u;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // 16 keys per MMA = 8 packed P columns and 2 KB of V rows
          t3_umma_ts(to, ts + (uint32_t)(j & 1) * 64 + kk * 8, t3_desc_mn_sw128(va + kk * 2048, 16384), IDESC_PV,
                     (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&S.o_done[g][j & 1]);
        if (last_half_of_tile(j)) umma_commit(&S.v_empty[g][s]);
      };
      mbar_wait(&S.q_full, 0);
      // Anti-phase start: group B begins one softmax pass after group A.  Both groups share the four SFUs; started together
      // they stay in lockstep (tcgen05.ld, max, exp2, tcgen05.st phases coincide and the exp2 phases halve each other's
      // rate), started half a period apart one group's TMEM traffic and dependent chains hide under the other's exp2.
      if (STAGGER && g == 1) mbar_wait(&S.p_full[0][0], 0);
      issue_s(0);
      if (n_g > 1) issue_s(1);
      T3_STAMP(12, g);
      for (int j = 0; j < n_g; ++j) {
        issue_pv(j);
        T3_STAMP(8 + 2 * g, j);
        if (j + 2 < n_g) issue_s(j + 2);
        T3_STAMP(9 + 2 * g, j);
      }
    }
  } else {
    // ------------------------------- softmax groups -------------------------------
    const int g = (warp - 2) >> 2;
    const int n_g = g == 0 ? n_a : n_b;
    if (n_g > 0) {
      const int r = (warp & 3) * 32 + lane;  // TMEM lane quarter of this warp -> row inside the tile
      const int qi = m0 + g * 128 + r;
      const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t ts = tmem + (uint32_t)g * 128 + lane_addr;
      const uint32_t to = tmem + O_COL + (uint32_t)g * HD + lane_addr;
      const int key_lim = CAUSAL ? min(kv_total - 1, qi + off) : kv_total - 1;
      float m_ref = -INFINITY, l_run = 0.f;

      for (int j = 0; j < n_g; ++j) {
        const uint32_t tsj = ts + (uint32_t)(j & 1) * 64;
        mbar_wait(&S.s_full[g][j & 1], (j >> 1) & 1);
        tc_fence_after();
        if (r == 0) T3_STAMP(g * 4 + 0, j);
        uint32_t v[64];
        tmem_ld_32x32b_x32(tsj, v);
        tmem_ld_32x32b_x32(tsj + 32, v + 32);
        tmem_ld_wait();
        if (r == 0) T3_STAMP(g * 4 + 1, j);
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
        } else if (mx > m_ref + T3_RESCALE_LOG2) {
          alpha = t3_ex2(m_ref - mx);
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
            const float p0 = (EXP_POLY && (e & 6) == 0) ? t3_ex2_poly(x0) : t3_ex2(x0);  // elements 0 of every 8 ...
            const float p1 = (EXP_POLY && (e & 6) == 4) ? t3_ex2_poly(x1) : t3_ex2(x1);  // ... and 5 of every 8: 1 in 4
            ls0 += p0;
            ls1 += p1;
            __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
            pk[(e >> 1) & 15] = *reinterpret_cast<uint32_t*>(&t);
          }
          tmem_st_32x32b_x16(tsj + hh * 16, pk);  // P overwrites S columns this thread has already read; the first store
        }                                         // drains while the second 32 exponentials are evaluated
        l_run = l_run * alpha + (ls0 + ls1);
        if (j > 0 && __any_sync(0xffffffffu, rescale)) {
          // rare: bring O to the new reference max before PV_j accumulates onto it
          mbar_wait(&S.o_done[g][(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(to + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
            t3_tmem_st_x32(to + c * 32, ov);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&S.p_full[g][j & 1]);
        if (r == 0) T3_STAMP(g * 4 + 2, j);
      }
      // epilogue (PV(n_g - 3), the previous phase of this barrier, is known complete: S(n_g - 1) was issued behind it)
      mbar_wait(&S.o_done[g][(n_g - 1) & 1], ((n_g - 1) >> 1) & 1);
      tc_fence_after();
      const float inv = 1.f / l_run;
      bf16* dst = out + (int64_t)(q_start + qi) * o_tok_stride + (int64_t)h * HD;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(to + c * 32, ov);
        tmem_ld_wait();
        if (qi < q_len) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            float f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(ov[e + u]) * inv;
            st8(dst + c * 32 + e, pack8(f));
          }
        }
      }
      if (r == 0) T3_STAMP(g * 4 + 3, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

template <int HD, bool CAUSAL, bool PACKED, bool EXP_POLY, bool STAGGER>
static int launch_t3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, bf16* out, int64_t o_tok_stride,
                     const int32_t* cu_q, const int32_t* kv_past, int B, int heads, int cap, int max_q_len, float scale,
                     cudaStream_t stream) {
  const int smem = (int)sizeof(T3Smem<HD>) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_prefill_tc3_kernel<HD, CAUSAL, PACKED, EXP_POLY, STAGGER>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_attn_tc3(smem attr)");
    attr_set = true;
  }
  dim3 grid((max_q_len + 255) / 256, heads, B);
  attn_prefill_tc3_kernel<HD, CAUSAL, PACKED, EXP_POLY, STAGGER><<<grid, T3_THREADS, smem, stream>>>(
      tq, tk, tv, out, o_tok_stride, cu_q, kv_past, heads, cap, scale * 1.4426950408889634f);
  DL_CHECK_LAUNCH("dl_attn_tc3");
}

// flags bit 0: evaluate one exp2 in four with the FMA-pipe polynomial; bit 1: start both groups together (no anti-phase)
extern "C" int dl_attn_prefill_tc3(const void* q, int64_t q_tok_stride, const void* k_arena, const void* v_arena, void* out,
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
#define T3_GO(P, G) launch_t3<128, true, false, P, G>(tq, tk, tv, (bf16*)out, o_tok_stride, cu_q, kv_past, B, heads, cap, max_q_len, scale, stream)
  switch (flags & 3) {
    case 0: return T3_GO(false, true);
    case 1: return T3_GO(true, true);
    case 2: return T3_GO(false, false);
    default: return T3_GO(true, false);
  }
#undef T3_GO
}

// K6/K7: non-causal attention over a fused projection qkv [n_tok, 3*heads*64] = [q | k | v] (HF CLIPAttention;
// custom_transformer_layer.py:153-180), per-sample token ranges from the CSR header cu_q.  out [n_tok, heads*64].
extern "C" int dl_attn_packed_tc3(const void* qkv, void* out, const int32_t* cu_q, int n_tok, int B, int heads, int head_dim,
                                  int max_q_len, float scale, int flags, cudaStream_t stream) {
  DL_REQUIRE(head_dim == 64, "dl_attn_packed_tc3");
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
#define T3_GO(P, G) launch_t3<64, false, true, P, G>(tq, tk, tv, (bf16*)out, C, cu_q, nullptr, B, heads, 0, max_q_len, scale, stream)
  switch (flags & 3) {
    case 0: return T3_GO(false, true);
    case 1: return T3_GO(true, true);
    case 2: return T3_GO(false, false);
    default: return T3_GO(true, false);
  }
#undef T3_GO
}
