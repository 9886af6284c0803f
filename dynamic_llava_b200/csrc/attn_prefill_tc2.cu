// K2 (v2): variable-length causal prefill attention, head_dim 128 — ping-pong tcgen05 pipeline.
// (The non-persistent form; the default is the persistent v4, attn_prefill_tc4.cu.)  Reference op: F.scaled_dot_product_attention(is_causal=True),
// llava/model/language_model/dynamic_modeling_llama.py:1114-1122.
//
// One CTA = (256 query rows = two 128-row tiles A and B, one head, one sample); 320 threads:
//   warp 0      TMA producer: Q_A, Q_B once; K_j, V_j 128x128 tiles (two SW128 boxes each) into a 2-stage ring;
//   warp 1      MMA issuer (one thread).  Per KV tile j the tensor pipe runs  PV_A(j) S_A(j+1) PV_B(j) S_B(j+1):
//               while softmax group A works on S_A(j+1) the tensor core is busy with group B's MMAs, and vice versa,
//               and each K/V tile fetched from L2/HBM feeds two query tiles;
//   warps 2-5   softmax group A, warps 6-9 softmax group B: one thread per query row (TMEM lane == row).
// TMEM (all 512 columns): S_A [0,128) S_B [128,256) O_A [256,384) O_B [384,512).
//   * S is read ONCE into registers (tcgen05.ld), P = exp2(S*scale - m) is written back as packed bf16 over the first
//     64 columns of the same S region (tcgen05.st) and consumed from TMEM as the A operand of the PV MMA
//     (tcgen05.mma with A in tensor memory) — P never touches shared memory;
//   * O accumulates in TMEM across KV tiles (accumulate flag), V is consumed in place as an MN-major B operand;
//   * the running max is only refreshed (and O rescaled through tcgen05.ld/st) when it grows by more than 2^8 —
//     P stays bounded by 256 in bf16/fp32 and the final O / l is exact — so the correction path is rare.
#include <cuda.h>

#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define T2_THREADS 320
#define T2_RESCALE_LOG2 8.0f

// HD = 128: Llama prefill (causal, K/V in the arena).  HD = 64: CLIP ViT-L/14 and the VisionPredictor blocks (non-causal,
// q|k|v packed in one fused projection buffer) — K6/K7's attention on the same TMA + tcgen05 pipeline.
template <int HD>
struct T2Smem {
  static constexpr int TILE = 128 * HD * 2;  // one 128-row operand tile = HD/64 blocks of 16 KB (128B-swizzled)
  alignas(1024) uint8_t q[2][TILE];
  alignas(1024) uint8_t k[2][TILE];
  alignas(1024) uint8_t v[2][TILE];
  alignas(8) uint64_t q_full;
  alignas(8) uint64_t k_full[2];
  alignas(8) uint64_t v_full[2];
  alignas(8) uint64_t kv_empty[2];
  alignas(8) uint64_t s_full[2];
  alignas(8) uint64_t p_full[2];
  alignas(8) uint64_t o_done[2];
  uint32_t tmem_base;
};

DL_DEVINL uint64_t t2_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D[tmem] (+)= A[tmem] . B[smem]
DL_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
DL_DEVINL void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
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
DL_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD, bool CAUSAL, bool PACKED>
__global__ void __launch_bounds__(T2_THREADS, 1) attn_prefill_tc2_kernel(
    const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
    const __grid_constant__ CUtensorMap tmap_v, bf16* __restrict__ out, int64_t o_tok_stride,
    const int32_t* __restrict__ cu_q, const int32_t* __restrict__ kv_past, int heads, int cap, float scale_log2) {
  constexpr int T2_TILE_BYTES = T2Smem<HD>::TILE;
  constexpr int NBOX = HD / 64;   // 64-column TMA boxes per operand tile
  constexpr int KSTEPS = HD / 16;  // K = 16 MMAs of S = Q K^T
  extern __shared__ uint8_t smem_raw[];
  T2Smem<HD>& S = *reinterpret_cast<T2Smem<HD>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
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
  const int n_a = ((CAUSAL ? min(kv_total, m0 + 128 + off) : kv_total) + 127) / 128;
  const int n_b = b_active ? ((CAUSAL ? min(kv_total, m0 + 256 + off) : kv_total) + 127) / 128 : 0;
  const int n_tiles = max(n_a, n_b);
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
      mbar_init(&S.kv_empty[s], 1);
      mbar_init(&S.s_full[s], 1);
      mbar_init(&S.p_full[s], 128);
      mbar_init(&S.o_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&S.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (elect_one()) {  // converged warp + elect.sync keeps the UTCHMMA/UTMALDG issue on the uniform datapath
      mbar_expect_tx(&S.q_full, (b_active ? 2 : 1) * T2_TILE_BYTES);
#pragma unroll
      for (int x = 0; x < NBOX; ++x) {
        tma_load_2d(S.q[0] + x * 16384, &tmap_q, h * HD + x * 64, q_start + m0, &S.q_full);
        if (b_active) tma_load_2d(S.q[1] + x * 16384, &tmap_q, h * HD + x * 64, q_start + m0 + 128, &S.q_full);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        mbar_wait(&S.kv_empty[s], ((j >> 1) & 1) ^ 1);
        const int row = kv_row0 + j * 128;
        mbar_expect_tx(&S.k_full[s], T2_TILE_BYTES);
#pragma unroll
        for (int x = 0; x < NBOX; ++x) tma_load_2d(S.k[s] + x * 16384, &tmap_k, kv_col0 + x * 64, row, &S.k_full[s]);
        mbar_expect_tx(&S.v_full[s], T2_TILE_BYTES);
#pragma unroll
        for (int x = 0; x < NBOX; ++x) tma_load_2d(S.v[s] + x * 16384, &tmap_v, kv_col0 + x * 64, row, &S.v_full[s]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    if (elect_one()) {  // converged warp + elect.sync keeps the UTCHMMA/UTMALDG issue on the uniform datapath
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, HD, 0, 1);  // N = head_dim; B (= V) MN-major
      auto issue_s = [&](int g, int j) {
        const int s = j & 1;
        mbar_wait(&S.k_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(S.q[g]), ka = smem_u32(S.k[s]);
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
          const uint32_t o = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_bf16_ss(tmem + (uint32_t)g * 128, umma_desc_kmajor_sw128(qa + o), umma_desc_kmajor_sw128(ka + o), IDESC_QK, kk > 0 ? 1u : 0u);
        }
        umma_commit(&S.s_full[g]);
      };
      auto issue_pv = [&](int g, int j) {
        const int s = j & 1;
        mbar_wait(&S.p_full[g], j & 1);
        mbar_wait(&S.v_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t va = smem_u32(S.v[s]);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // 16 keys per MMA = 8 packed P columns and 2 KB of V rows
          umma_bf16_ts(tmem + 256 + (uint32_t)g * 128, tmem + (uint32_t)g * 128 + kk * 8, t2_desc_mn_sw128(va + kk * 2048, 16384),
                       IDESC_PV, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&S.o_done[g]);
      };
      mbar_wait(&S.q_full, 0);
      issue_s(0, 0);
      if (b_active) issue_s(1, 0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j < n_a) issue_pv(0, j);
        if (j + 1 < n_a) issue_s(0, j + 1);
        if (j < n_b) issue_pv(1, j);
        umma_commit(&S.kv_empty[j & 1]);  // every MMA that reads stage j&1 has been issued
        if (j + 1 < n_b) issue_s(1, j + 1);
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
      const uint32_t to = tmem + 256 + (uint32_t)g * 128 + lane_addr;
      const int key_lim = CAUSAL ? min(kv_total - 1, qi + off) : kv_total - 1;
      float m_ref = -INFINITY, l_run = 0.f;

      for (int j = 0; j < n_g; ++j) {
        mbar_wait(&S.s_full[g], j & 1);
        tc_fence_after();
        uint32_t v[128];
        tmem_ld_32x32b_x32(ts, v);
        tmem_ld_32x32b_x32(ts + 32, v + 32);
        tmem_ld_32x32b_x32(ts + 64, v + 64);
        tmem_ld_32x32b_x32(ts + 96, v + 96);
        tmem_ld_wait();
        const int kbase = j * 128;
        const bool need_mask = kbase + 127 > key_lim;
        float mx = -INFINITY;
        if (need_mask) {
#pragma unroll
          for (int e = 0; e < 128; ++e) {
            if (kbase + e > key_lim) v[e] = 0xff800000u;  // -inf
            mx = fmaxf(mx, __uint_as_float(v[e]));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 128; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
        }
        mx *= scale_log2;
        float alpha = 1.f;
        bool rescale = false;
        if (j == 0) {
          m_ref = (mx == -INFINITY) ? 0.f : mx;
        } else if (mx > m_ref + T2_RESCALE_LOG2) {
          alpha = ex2(m_ref - mx);
          m_ref = mx;
          rescale = true;
        }
        float ls = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t pk[32];
#pragma unroll
          for (int e = 0; e < 64; e += 2) {
            const float p0 = ex2(__uint_as_float(v[hh * 64 + e]) * scale_log2 - m_ref);
            const float p1 = ex2(__uint_as_float(v[hh * 64 + e + 1]) * scale_log2 - m_ref);
            ls += p0 + p1;
            __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&t);
          }
          tmem_st_32x32b_x32(ts + hh * 32, pk);  // P overwrites the S columns this thread has already read
        }
        l_run = l_run * alpha + ls;
        if (j > 0 && __any_sync(0xffffffffu, rescale)) {
          // rare: bring O to the new reference max before PV_j accumulates onto it
          mbar_wait(&S.o_done[g], (j - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(to + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
            tmem_st_32x32b_x32(to + c * 32, ov);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&S.p_full[g]);
      }
      // epilogue
      mbar_wait(&S.o_done[g], (n_g - 1) & 1);
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
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

template <int HD, bool CAUSAL, bool PACKED>
static int launch_t2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, bf16* out, int64_t o_tok_stride,
                     const int32_t* cu_q, const int32_t* kv_past, int B, int heads, int cap, int max_q_len, float scale,
                     cudaStream_t stream) {
  const int smem = (int)sizeof(T2Smem<HD>) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_prefill_tc2_kernel<HD, CAUSAL, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_attn_tc2(smem attr)");
    attr_set = true;
  }
  dim3 grid((max_q_len + 255) / 256, heads, B);
  attn_prefill_tc2_kernel<HD, CAUSAL, PACKED><<<grid, T2_THREADS, smem, stream>>>(tq, tk, tv, out, o_tok_stride, cu_q, kv_past, heads, cap,
                                                                                  scale * 1.4426950408889634f);
  DL_CHECK_LAUNCH("dl_attn_tc2");
}

extern "C" int dl_attn_prefill_tc2(const void* q, int64_t q_tok_stride, const void* k_arena, const void* v_arena, void* out,
                                   int64_t o_tok_stride, const int32_t* cu_q, const int32_t* kv_past, int n_tok, int B, int heads,
                                   int cap, int max_q_len, float scale, cudaStream_t stream) {
  if (B == 0 || max_q_len == 0) return 0;
  alignas(64) CUtensorMap tq, tk, tv;
  int rc = dl_make_tmap_2d(&tq, q, (uint64_t)n_tok, (uint64_t)heads * 128, (uint64_t)q_tok_stride, 128, 64);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tk, k_arena, (uint64_t)B * heads * cap, 128, 128, 128, 64);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tv, v_arena, (uint64_t)B * heads * cap, 128, 128, 128, 64);
  if (rc) return rc;
  return launch_t2<128, true, false>(tq, tk, tv, (bf16*)out, o_tok_stride, cu_q, kv_past, B, heads, cap, max_q_len, scale, stream);
}

// K6/K7: non-causal attention over a fused projection qkv [n_tok, 3*heads*64] = [q | k | v] (HF CLIPAttention;
// custom_transformer_layer.py:153-180), per-sample token ranges from the CSR header cu_q.  out [n_tok, heads*64].
extern "C" int dl_attn_packed_tc(const void* qkv, void* out, const int32_t* cu_q, int n_tok, int B, int heads, int head_dim,
                                 int max_q_len, float scale, cudaStream_t stream) {
  DL_REQUIRE(head_dim == 64, "dl_attn_packed_tc");
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
  return launch_t2<64, false, true>(tq, tk, tv, (bf16*)out, C, cu_q, nullptr, B, heads, 0, max_q_len, scale, stream);
}
