// K2 (portable tile path) / K6 / K7: variable-length flash attention with warp-level bf16 MMA.
//
// Serves every attention on the path whose head_dim is 64 (CLIP ViT-L/14, 577 tokens, non-causal — HF CLIPAttention;
// VisionPredictor blocks, 576 tokens, non-causal — custom_transformer_layer.py:153-180) and is the small-sequence
// path for the Llama prefill (head_dim 128, causal — dynamic_modeling_llama.py:1114-1122).  The tcgen05/TMEM kernel in
// attn_prefill_tc4.cu (tcgen05, persistent) takes the causal head_dim-128 and the packed head_dim-64 cases.
//
// Per-sample extents come from a CSR header (`cu_q`, built by dl_select_compact) plus the arena's per-sample lengths,
// so no padding is ever materialised: grid = (q tiles, heads, samples), out-of-range tiles exit.
// 64 query rows per CTA (16 per warp), 64-key K/V tiles double-buffered with cp.async, XOR-swizzled shared memory,
// ldmatrix fragments, online softmax in fp32 (exp2 with the scale folded in), P rounded to bf16 for the PV MMA.
#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define FA_BM 64
#define FA_BN 64
#define FA_THREADS 128

struct FaParams {
  const bf16* q;
  const bf16* k;
  const bf16* v;
  bf16* o;
  const int32_t* cu_q;     // [B+1] CSR header over query tokens
  const int32_t* kv_past;  // [B] keys already in the arena before this pass, or null (=0)
  int64_t q_tok_stride, q_head_stride;
  int64_t k_batch_stride, k_head_stride, k_tok_stride;
  int64_t v_batch_stride, v_head_stride, v_tok_stride;
  int64_t o_tok_stride;
  int kv_packed;  // 1: K/V rows of sample b start at cu_q[b] (packed token-major); 0: at row 0 of the sample's arena slab
  int heads;
  float scale_log2;
};

DL_DEVINL void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = smem_u32(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
DL_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
DL_DEVINL void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
DL_DEVINL void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
DL_DEVINL void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
DL_DEVINL void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
DL_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int HD>
DL_DEVINL int swz(int row, int chunk) {  // element offset of 16-byte chunk `chunk` of row `row`
  return row * HD + ((chunk ^ (row & 7)) << 3);
}

template <int HD>
DL_DEVINL void load_tile(bf16* smem, const bf16* gbase, int64_t tok_stride, int row0, int n_valid_rows) {
  constexpr int CH = HD / 8;
  for (int idx = threadIdx.x; idx < FA_BN * CH; idx += FA_THREADS) {
    const int r = idx / CH, c = idx % CH;
    const bool ok = (row0 + r) < n_valid_rows;
    const bf16* src = gbase + (int64_t)(ok ? row0 + r : 0) * tok_stride + c * 8;
    cp_async16(smem + swz<HD>(r, c), src, ok);
  }
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(FA_THREADS) attn_prefill_mma_kernel(const FaParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + FA_BM * HD;           // [2][FA_BN * HD]
  bf16* sV = sK + 2 * FA_BN * HD;       // [2][FA_BN * HD]
  constexpr int KS = HD / 16;           // k-steps of the QK^T MMA
  constexpr int NT = HD / 8;            // n-tiles of the PV MMA

  const int b = blockIdx.z, h = blockIdx.y, mt = blockIdx.x;
  const int q_start = p.cu_q[b];
  const int q_len = p.cu_q[b + 1] - q_start;
  const int m0 = mt * FA_BM;
  if (m0 >= q_len) return;
  const int past = p.kv_past ? p.kv_past[b] : 0;
  const int kv_total = past + q_len;
  const int off = kv_total - q_len;  // causal: query i sees keys j <= i + off
  int kv_end = kv_total;
  if (CAUSAL) kv_end = min(kv_total, m0 + FA_BM + off);
  const int n_tiles = (kv_end + FA_BN - 1) / FA_BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;

  const bf16* qg = p.q + (int64_t)q_start * p.q_tok_stride + (int64_t)h * p.q_head_stride;
  const int kv_row0 = p.kv_packed ? q_start : 0;
  const bf16* kg = p.k + (int64_t)b * p.k_batch_stride + (int64_t)h * p.k_head_stride + (int64_t)kv_row0 * p.k_tok_stride;
  const bf16* vg = p.v + (int64_t)b * p.v_batch_stride + (int64_t)h * p.v_head_stride + (int64_t)kv_row0 * p.v_tok_stride;

  // prologue: Q tile + first K/V tile
  load_tile<HD>(sQ, qg, p.q_tok_stride, m0, q_len);
  load_tile<HD>(sK, kg, p.k_tok_stride, 0, kv_total);
  load_tile<HD>(sV, vg, p.v_tok_stride, 0, kv_total);
  cp_async_commit();

  uint32_t qf[KS][4];
  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      load_tile<HD>(sK + (buf ^ 1) * FA_BN * HD, kg, p.k_tok_stride, (t + 1) * FA_BN, kv_total);
      load_tile<HD>(sV + (buf ^ 1) * FA_BN * HD, vg, p.v_tok_stride, (t + 1) * FA_BN, kv_total);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int row = warp * 16 + (lane & 15);
        ldsm_x4(smem_u32(sQ + swz<HD>(row, ks * 2 + (lane >> 4))), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const bf16* Kt = sK + buf * FA_BN * HD;
    const bf16* Vt = sV + buf * FA_BN * HD;

    // S = Q K^T   (16 rows x 64 keys per warp)
    float s[FA_BN / 8][4];
#pragma unroll
    for (int j = 0; j < FA_BN / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int jp = 0; jp < FA_BN / 16; ++jp) {
        uint32_t b0, b1, b2, b3;
        const int row = jp * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
        ldsm_x4(smem_u32(Kt + swz<HD>(row, ks * 2 + ((lane >> 3) & 1))), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * jp], qf[ks], b0, b1);
        mma_bf16_16816(s[2 * jp + 1], qf[ks], b2, b3);
      }
    }

    // mask + online softmax (rows g and g+8 of this warp's 16)
    const int kbase = t * FA_BN;
    const int i0 = m0 + warp * 16 + g, i1 = i0 + 8;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < FA_BN / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kbase + j * 8 + t4 * 2 + (e & 1);
        const int qi = (e < 2) ? i0 : i1;
        bool ok = key < kv_total;
        if (CAUSAL) ok = ok && (key <= qi + off);
        float v = ok ? s[j][e] * p.scale_log2 : -INFINITY;
        s[j][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float alpha[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mnew[r] = fmaxf(m_run[r], mx[r]);
      if (mnew[r] == -INFINITY) mnew[r] = 0.f;  // fully masked so far: keep everything at zero weight
      alpha[r] = exp2f(m_run[r] - mnew[r]);
      m_run[r] = mnew[r];
    }
    float ls[2] = {0.f, 0.f};
    uint32_t pa[FA_BN / 16][4];
#pragma unroll
    for (int j = 0; j < FA_BN / 8; ++j) {
      const float p0 = exp2f(s[j][0] - mnew[0]), p1 = exp2f(s[j][1] - mnew[0]);
      const float p2 = exp2f(s[j][2] - mnew[1]), p3 = exp2f(s[j][3] - mnew[1]);
      ls[0] += p0 + p1;
      ls[1] += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      ls[r] += __shfl_xor_sync(0xffffffffu, ls[r], 1);
      ls[r] += __shfl_xor_sync(0xffffffffu, ls[r], 2);
      l_run[r] = l_run[r] * alpha[r] + ls[r];
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      o[j][0] *= alpha[0];
      o[j][1] *= alpha[0];
      o[j][2] *= alpha[1];
      o[j][3] *= alpha[1];
    }

    // O += P V
#pragma unroll
    for (int kk = 0; kk < FA_BN / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(smem_u32(Vt + swz<HD>(row, np * 2 + (lane >> 4))), b0, b1, b2, b3);
        mma_bf16_16816(o[2 * np], pa[kk], b0, b1);
        mma_bf16_16816(o[2 * np + 1], pa[kk], b2, b3);
      }
    }
    __syncthreads();  // everyone done with `buf` before the next iteration's prefetch overwrites it
  }

  // epilogue: normalise, stage through sQ (each warp touches only its own 16 rows), coalesced 128-bit stores
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int col = j * 8 + t4 * 2;
      *reinterpret_cast<uint32_t*>(sQ + swz<HD>(r0, col >> 3) + (col & 7)) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
      *reinterpret_cast<uint32_t*>(sQ + swz<HD>(r1, col >> 3) + (col & 7)) = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    }
  }
  __syncwarp();
  constexpr int CH = HD / 8;
  bf16* og = p.o + (int64_t)q_start * p.o_tok_stride + (int64_t)h * HD;
  for (int idx = lane; idx < 16 * CH; idx += 32) {
    const int r = warp * 16 + idx / CH, c = idx % CH;
    if (m0 + r < q_len)
      *reinterpret_cast<int4*>(og + (int64_t)(m0 + r) * p.o_tok_stride + c * 8) = *reinterpret_cast<const int4*>(sQ + swz<HD>(r, c));
  }
}

template <int HD, bool CAUSAL>
static int launch_fa(const FaParams& p, int B, int max_q_len, cudaStream_t stream) {
  const int smem = (FA_BM + 4 * FA_BN) * HD * (int)sizeof(bf16);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_prefill_mma_kernel<HD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_attn_varlen(smem attr)");
    attr_set = true;
  }
  dim3 grid((max_q_len + FA_BM - 1) / FA_BM, p.heads, B);
  attn_prefill_mma_kernel<HD, CAUSAL><<<grid, FA_THREADS, smem, stream>>>(p);
  DL_CHECK_LAUNCH("dl_attn_varlen");
}

extern "C" int dl_attn_varlen(const void* q, int64_t q_tok_stride, int64_t q_head_stride, const void* k, const void* v,
                              int64_t kv_batch_stride, int64_t kv_head_stride, int64_t kv_tok_stride, int kv_packed,
                              void* out, int64_t o_tok_stride, const int32_t* cu_q, const int32_t* kv_past, int B, int heads,
                              int head_dim, int max_q_len, int causal, float scale, cudaStream_t stream) {
  DL_REQUIRE(head_dim == 64 || head_dim == 128, "dl_attn_varlen");
  if (B == 0 || max_q_len == 0) return 0;
  FaParams p;
  p.q = (const bf16*)q;
  p.k = (const bf16*)k;
  p.v = (const bf16*)v;
  p.o = (bf16*)out;
  p.cu_q = cu_q;
  p.kv_past = kv_past;
  p.q_tok_stride = q_tok_stride;
  p.q_head_stride = q_head_stride;
  p.k_batch_stride = p.v_batch_stride = kv_batch_stride;
  p.k_head_stride = p.v_head_stride = kv_head_stride;
  p.k_tok_stride = p.v_tok_stride = kv_tok_stride;
  p.o_tok_stride = o_tok_stride;
  p.kv_packed = kv_packed;
  p.heads = heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (head_dim == 128) return causal ? launch_fa<128, true>(p, B, max_q_len, stream) : launch_fa<128, false>(p, B, max_q_len, stream);
  return causal ? launch_fa<64, true>(p, B, max_q_len, stream) : launch_fa<64, false>(p, B, max_q_len, stream);
}
