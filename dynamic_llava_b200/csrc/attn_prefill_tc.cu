// K2: variable-length causal prefill attention on the 5th-gen tensor cores (TMA + tcgen05.mma + TMEM), head_dim 128.
// Reference op: DynamicLlamaSdpaAttention prefill, F.scaled_dot_product_attention(is_causal=True)
// (llava/model/language_model/dynamic_modeling_llama.py:1114-1122), per-sample extents from the CSR header written by
// dl_select_compact and the per-sample arena lengths — no padding is materialised and tiles above the causal diagonal
// or past a sample's end are never issued.
//
// One CTA = one (128-query tile, head, sample); 192 threads:
//   warps 0-3  softmax / accumulate: one thread per query row (TMEM lane == row, so no shuffles): S row via tcgen05.ld,
//              exp2 online softmax in fp32, P written to shared memory as the bf16 A operand of the PV MMA, PV tile
//              read back from TMEM and folded into the fp32 O row held in registers;
//   warp 4     TMA producer: Q once, then K and V 128x128 tiles (two 64-column 128B-swizzled boxes each) into a
//              2-stage mbarrier ring;
//   warp 5     MMA issuer (one elected thread): S_j = Q K_j^T (K-major x K-major, 8 x K16) into a double-buffered
//              TMEM accumulator so the tensor core runs tile j+1 while the softmax warps work on tile j, then
//              T_j = P_j V_j with V consumed in place as an MN-major operand (no transpose pass).
// TMEM: S0 cols [0,128), S1 [128,256), T [256,384)  (512-column allocation, 1 CTA/SM by shared-memory footprint).
#include <cuda.h>

#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define TC_BM 128
#define TC_BN 128
#define TC_HD 128
#define TC_THREADS 192
#define TC_TILE_BYTES (128 * 128 * 2)  // one 128x128 bf16 operand tile = two 16 KB swizzled blocks

struct TcSmem {
  alignas(1024) uint8_t q[TC_TILE_BYTES];
  alignas(1024) uint8_t k[2][TC_TILE_BYTES];
  alignas(1024) uint8_t v[2][TC_TILE_BYTES];
  alignas(1024) uint8_t p[TC_TILE_BYTES];
  alignas(8) uint64_t bar_q;
  alignas(8) uint64_t k_full[2];
  alignas(8) uint64_t v_full[2];
  alignas(8) uint64_t kv_empty[2];
  alignas(8) uint64_t bar_s[2];
  alignas(8) uint64_t bar_p;
  alignas(8) uint64_t bar_t;
  uint32_t tmem_base;
};

// MN-major SWIZZLE_128B operand (V as B of the PV MMA): 64-element N atoms `lbo` bytes apart, 8-row K groups 1024 B apart
DL_DEVINL uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(TC_THREADS, 1) attn_prefill_tc_kernel(
    const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
    const __grid_constant__ CUtensorMap tmap_v, bf16* __restrict__ out, int64_t o_tok_stride,
    const int32_t* __restrict__ cu_q, const int32_t* __restrict__ kv_past, int heads, int cap, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  TcSmem& S = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q_start = cu_q[b];
  const int q_len = cu_q[b + 1] - q_start;
  const int m0 = mt * TC_BM;
  if (m0 >= q_len) return;  // uniform per CTA, before any barrier / TMEM allocation
  const int past = kv_past ? kv_past[b] : 0;
  const int kv_total = past + q_len;
  const int off = past;
  const int kv_end = min(kv_total, m0 + TC_BM + off);
  const int n_tiles = (kv_end + TC_BN - 1) / TC_BN;
  const int kv_row0 = (b * heads + h) * cap;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(&S.bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&S.k_full[s], 1);
      mbar_init(&S.v_full[s], 1);
      mbar_init(&S.kv_empty[s], 1);
      mbar_init(&S.bar_s[s], 1);
    }
    mbar_init(&S.bar_p, 128);
    mbar_init(&S.bar_t, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<512>(&S.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const uint32_t tmem_T = tmem + 256;

  if (warp == 4) {
    // ------------------------------- TMA producer -------------------------------
    if (elect_one()) {  // converged warp + elect.sync keeps the UTCHMMA/UTMALDG issue on the uniform datapath
      mbar_expect_tx(&S.bar_q, TC_TILE_BYTES);
      tma_load_2d(S.q, &tmap_q, h * TC_HD, q_start + m0, &S.bar_q);
      tma_load_2d(S.q + TC_TILE_BYTES / 2, &tmap_q, h * TC_HD + 64, q_start + m0, &S.bar_q);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&S.kv_empty[s], ph ^ 1);
        const int row = kv_row0 + j * TC_BN;
        mbar_expect_tx(&S.k_full[s], TC_TILE_BYTES);
        tma_load_2d(S.k[s], &tmap_k, 0, row, &S.k_full[s]);
        tma_load_2d(S.k[s] + TC_TILE_BYTES / 2, &tmap_k, 64, row, &S.k_full[s]);
        mbar_expect_tx(&S.v_full[s], TC_TILE_BYTES);
        tma_load_2d(S.v[s], &tmap_v, 0, row, &S.v_full[s]);
        tma_load_2d(S.v[s] + TC_TILE_BYTES / 2, &tmap_v, 64, row, &S.v_full[s]);
      }
    }
  } else if (warp == 5) {
    // ------------------------------- MMA issuer -------------------------------
    if (elect_one()) {  // converged warp + elect.sync keeps the UTCHMMA/UTMALDG issue on the uniform datapath
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, 128, 0, 1);  // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(S.q), p_addr = smem_u32(S.p);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&S.k_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(S.k[s]);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t o = (kk >> 2) * (TC_TILE_BYTES / 2) + (kk & 3) * 32;
          umma_bf16_ss(tmem + (uint32_t)s * 128, umma_desc_kmajor_sw128(q_addr + o), umma_desc_kmajor_sw128(k_addr + o),
                       IDESC_QK, kk > 0 ? 1u : 0u);
        }
        umma_commit(&S.bar_s[s]);
      };
      mbar_wait(&S.bar_q, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        if (j + 1 < n_tiles) issue_s(j + 1);  // tensor core works on tile j+1 while the softmax warps are on tile j
        mbar_wait(&S.bar_p, j & 1);
        mbar_wait(&S.v_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(S.v[s]);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t po = (kk >> 2) * (TC_TILE_BYTES / 2) + (kk & 3) * 32;
          umma_bf16_ss(tmem_T, umma_desc_kmajor_sw128(p_addr + po), umma_desc_mnmajor_sw128(v_addr + kk * 2048, TC_TILE_BYTES / 2),
                       IDESC_PV, kk > 0 ? 1u : 0u);
        }
        umma_commit(&S.bar_t);
        umma_commit(&S.kv_empty[s]);
      }
    }
  } else {
    // ------------------------------- softmax / accumulate (one thread per query row) -------------------------------
    const int r = threadIdx.x;  // row in tile == TMEM lane
    const int qi = m0 + r;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const int key_lim = min(kv_total - 1, qi + off);  // last visible key of this row
    float o[TC_HD];
#pragma unroll
    for (int d = 0; d < TC_HD; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = S.p + r * 128;

    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ts = tmem + (uint32_t)(j & 1) * 128 + lane_addr;
      mbar_wait(&S.bar_s[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int kbase = j * TC_BN;
      const bool need_mask = kbase + TC_BN - 1 > key_lim;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(ts + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float x = __uint_as_float(v[e]);
          if (need_mask && kbase + c * 32 + e > key_lim) x = -INFINITY;
          mx = fmaxf(mx, x);
        }
      }
      float m_new = fmaxf(m_run, mx * scale_log2);
      if (m_new == -INFINITY) m_new = 0.f;
      const float alpha = exp2f(m_run - m_new);
      float ls = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(ts + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float x0 = __uint_as_float(v[e]) * scale_log2 - m_new;
          float x1 = __uint_as_float(v[e + 1]) * scale_log2 - m_new;
          float p0 = exp2f(x0), p1 = exp2f(x1);
          if (need_mask) {
            if (kbase + c * 32 + e > key_lim) p0 = 0.f;
            if (kbase + c * 32 + e + 1 > key_lim) p1 = 0.f;
          }
          ls += p0 + p1;
          __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
          pk[e >> 1] = *reinterpret_cast<uint32_t*>(&t);
        }
        // 32 keys = 4 chunks of 16 B in block (c >> 1); logical chunk (c & 1) * 4 + i, 128B-swizzled by row
        uint8_t* blk = prow + (c >> 1) * (TC_TILE_BYTES / 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int chunk = ((c & 1) * 4 + i) ^ (r & 7);
          *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(pk[i * 4], pk[i * 4 + 1], pk[i * 4 + 2], pk[i * 4 + 3]);
        }
      }
      l_run = l_run * alpha + ls;
      m_run = m_new;
      fence_proxy_async();  // generic-proxy smem writes of P -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(&S.bar_p);

      mbar_wait(&S.bar_t, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_T + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[c * 32 + e] = o[c * 32 + e] * alpha + __uint_as_float(v[e]);
      }
      tc_fence_before();
    }
    if (qi < q_len) {
      const float inv = 1.f / l_run;
      bf16* dst = out + (int64_t)(q_start + qi) * o_tok_stride + (int64_t)h * TC_HD;
#pragma unroll
      for (int c = 0; c < TC_HD / 8; ++c) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = o[c * 8 + e] * inv;
        st8(dst + c * 8, pack8(f));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem);
}

extern "C" int dl_attn_prefill_tc(const void* q, int64_t q_tok_stride, const void* k_arena, const void* v_arena, void* out,
                                  int64_t o_tok_stride, const int32_t* cu_q, const int32_t* kv_past, int n_tok, int B, int heads,
                                  int cap, int max_q_len, float scale, cudaStream_t stream) {
  if (B == 0 || max_q_len == 0) return 0;
  alignas(64) CUtensorMap tq, tk, tv;
  int rc = dl_make_tmap_2d(&tq, q, (uint64_t)n_tok, (uint64_t)heads * TC_HD, (uint64_t)q_tok_stride, TC_BM, 64);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tk, k_arena, (uint64_t)B * heads * cap, TC_HD, TC_HD, TC_BN, 64);
  if (rc) return rc;
  rc = dl_make_tmap_2d(&tv, v_arena, (uint64_t)B * heads * cap, TC_HD, TC_HD, TC_BN, 64);
  if (rc) return rc;
  const int smem = (int)sizeof(TcSmem) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_attn_prefill_tc(smem attr)");
    attr_set = true;
  }
  dim3 grid((max_q_len + TC_BM - 1) / TC_BM, heads, B);
  attn_prefill_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(tq, tk, tv, (bf16*)out, o_tok_stride, cu_q, kv_past, heads, cap,
                                                             scale * 1.4426950408889634f);
  DL_CHECK_LAUNCH("dl_attn_prefill_tc");
}
