// K2 (v3): variable-length causal prefill attention, head_dim 128 — ping-pong tcgen05 pipeline, two threads per query row.
// Same contract and data flow as attn_prefill_tc2.cu (two 128-row query tiles A/B per CTA, S/P/O resident in TMEM, P fed to
// the PV MMA from tensor memory, lazy O rescale); reference op: F.scaled_dot_product_attention(is_causal=True),
// llava/model/language_model/dynamic_modeling_llama.py:1114-1122.
//
// v2's profile (profiles/r1_ncu_attn_prefill_tc2.md) showed the tensor pipe idle 2/3 of the time waiting for the softmax
// warps (one thread per 128-element row: ~700 dependent-latency-bound instructions per tile).  v3 doubles the softmax
// width: 576 threads = TMA warp + MMA warp + 2 groups x 8 warps; warps w and w+4 of a group share a TMEM lane quarter and
// split each row's 128 columns in halves, exchanging the half-row maxima through shared memory behind one 256-thread named
// barrier per tile, so the per-tile softmax latency that gates the next PV MMA is halved.
#include <cuda.h>

#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define T3_THREADS 576
#define T3_TILE_BYTES (128 * 128 * 2)
#define T3_RESCALE_LOG2 8.0f

struct T3Smem {
  alignas(1024) uint8_t q[2][T3_TILE_BYTES];
  alignas(1024) uint8_t k[2][T3_TILE_BYTES];
  alignas(1024) uint8_t v[2][T3_TILE_BYTES];
  alignas(8) uint64_t q_full;
  alignas(8) uint64_t k_full[2];
  alignas(8) uint64_t v_full[2];
  alignas(8) uint64_t kv_empty[2];
  alignas(8) uint64_t s_full[2];
  alignas(8) uint64_t p_full[2];
  alignas(8) uint64_t o_done[2];
  float xmax[2][2][128];  // [group][column half][row]: half-row maxima / sums exchanged between the two threads of a row
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

__global__ void __launch_bounds__(T3_THREADS, 1) attn_prefill_tc3_kernel(
    const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
    const __grid_constant__ CUtensorMap tmap_v, bf16* __restrict__ out, int64_t o_tok_stride,
    const int32_t* __restrict__ cu_q, const int32_t* __restrict__ kv_past, int heads, int cap, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  T3Smem& S = *reinterpret_cast<T3Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
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
  const int n_a = (min(kv_total, m0 + 128 + off) + 127) / 128;
  const int n_b = b_active ? (min(kv_total, m0 + 256 + off) + 127) / 128 : 0;
  const int n_tiles = max(n_a, n_b);
  const int kv_row0 = (b * heads + h) * cap;

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
      mbar_init(&S.p_full[s], 256);
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
    if (lane == 0) {
      mbar_expect_tx(&S.q_full, (b_active ? 2 : 1) * T3_TILE_BYTES);
      tma_load_2d(S.q[0], &tmap_q, h * 128, q_start + m0, &S.q_full);
      tma_load_2d(S.q[0] + T3_TILE_BYTES / 2, &tmap_q, h * 128 + 64, q_start + m0, &S.q_full);
      if (b_active) {
        tma_load_2d(S.q[1], &tmap_q, h * 128, q_start + m0 + 128, &S.q_full);
        tma_load_2d(S.q[1] + T3_TILE_BYTES / 2, &tmap_q, h * 128 + 64, q_start + m0 + 128, &S.q_full);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        mbar_wait(&S.kv_empty[s], ((j >> 1) & 1) ^ 1);
        const int row = kv_row0 + j * 128;
        mbar_expect_tx(&S.k_full[s], T3_TILE_BYTES);
        tma_load_2d(S.k[s], &tmap_k, 0, row, &S.k_full[s]);
        tma_load_2d(S.k[s] + T3_TILE_BYTES / 2, &tmap_k, 64, row, &S.k_full[s]);
        mbar_expect_tx(&S.v_full[s], T3_TILE_BYTES);
        tma_load_2d(S.v[s], &tmap_v, 0, row, &S.v_full[s]);
        tma_load_2d(S.v[s] + T3_TILE_BYTES / 2, &tmap_v, 64, row, &S.v_full[s]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, 128, 0, 1);  // B (= V) MN-major
      auto issue_s = [&](int g, int j) {
        const int s = j & 1;
        mbar_wait(&S.k_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(S.q[g]), ka = smem_u32(S.k[s]);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t o = (kk >> 2) * (T3_TILE_BYTES / 2) + (kk & 3) * 32;
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
          umma_bf16_ts(tmem + 256 + (uint32_t)g * 128, tmem + (uint32_t)g * 128 + kk * 8, t3_desc_mn_sw128(va + kk * 2048, T3_TILE_BYTES / 2),
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
    // ------------------------------- softmax groups (2 threads per query row) -------------------------------
    const int g = (warp - 2) >> 3;
    const int half = ((warp - 2) >> 2) & 1;  // which 64 of the row's 128 score columns this thread owns
    const int n_g = g == 0 ? n_a : n_b;
    if (n_g > 0) {
      const int r = (warp & 3) * 32 + lane;  // TMEM lane quarter of this warp -> row inside the tile
      const int qi = m0 + g * 128 + r;
      const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t ts = tmem + (uint32_t)g * 128 + lane_addr;
      const uint32_t to = tmem + 256 + (uint32_t)g * 128 + lane_addr + half * 64;
      const int key_lim = min(kv_total - 1, qi + off);
      const int bar_id = 1 + g;
      float m_ref = -INFINITY, l_run = 0.f;

      for (int j = 0; j < n_g; ++j) {
        mbar_wait(&S.s_full[g], j & 1);
        tc_fence_after();
        uint32_t v[64];
        tmem_ld_32x32b_x32(ts + half * 64, v);
        tmem_ld_32x32b_x32(ts + half * 64 + 32, v + 32);
        tmem_ld_wait();
        const int kbase = j * 128 + half * 64;
        const bool need_mask = kbase + 63 > key_lim;
        float mx = -INFINITY;
        if (need_mask) {
#pragma unroll
          for (int e = 0; e < 64; ++e) {
            if (kbase + e > key_lim) v[e] = 0xff800000u;  // -inf
            mx = fmaxf(mx, __uint_as_float(v[e]));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 64; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
        }
        S.xmax[g][half][r] = mx;
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
        mx = fmaxf(mx, S.xmax[g][half ^ 1][r]) * scale_log2;
        float alpha = 1.f;
        bool rescale = false;
        if (j == 0) {
          m_ref = (mx == -INFINITY) ? 0.f : mx;
        } else if (mx > m_ref + T3_RESCALE_LOG2) {
          alpha = ex2(m_ref - mx);
          m_ref = mx;
          rescale = true;
        }
        float ls = 0.f;
        {
          uint32_t pk[32];
#pragma unroll
          for (int e = 0; e < 64; e += 2) {
            const float p0 = ex2(__uint_as_float(v[e]) * scale_log2 - m_ref);
            const float p1 = ex2(__uint_as_float(v[e + 1]) * scale_log2 - m_ref);
            ls += p0 + p1;
            __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&t);
          }
          // P (packed bf16) overwrites S columns [0,64); the partner's S columns may live there, so everybody must have
          // finished READING S first: the exchange barrier above sits between the loads and this store.
          tmem_st_32x32b_x32(ts + half * 32, pk);
        }
        l_run = l_run * alpha + ls;
        if (j > 0 && __any_sync(0xffffffffu, rescale)) {
          mbar_wait(&S.o_done[g], (j - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
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
      // epilogue: combine the two half-row sums, then each thread normalises and stores its 64 output columns
      mbar_wait(&S.o_done[g], (n_g - 1) & 1);
      tc_fence_after();
      S.xmax[g][half][r] = l_run;
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      const float inv = 1.f / (l_run + S.xmax[g][half ^ 1][r]);
      bf16* dst = out + (int64_t)(q_start + qi) * o_tok_stride + (int64_t)h * 128 + half * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
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

extern "C" int dl_attn_prefill_tc3(const void* q, int64_t q_tok_stride, const void* k_arena, const void* v_arena, void* out,
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
  const int smem = (int)sizeof(T3Smem) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_prefill_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_attn_prefill_tc3(smem attr)");
    attr_set = true;
  }
  dim3 grid((max_q_len + 255) / 256, heads, B);
  attn_prefill_tc3_kernel<<<grid, T3_THREADS, smem, stream>>>(tq, tk, tv, (bf16*)out, o_tok_stride, cu_q, kv_past, heads, cap,
                                                              scale * 1.4426950408889634f);
  DL_CHECK_LAUNCH("dl_attn_prefill_tc3");
}
