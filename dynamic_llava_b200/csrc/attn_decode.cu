// K3: ragged split-KV decode attention (q_len == 1) over the KV arena [B, heads, cap, 128].
// Reference: DynamicLlamaSdpaAttention with q_len == 1 (dynamic_modeling_llama.py:1061-1122): the new token attends to
// cache ++ itself, unmasked, scale 1/sqrt(128), fp32 softmax; whether its k/v are persisted is decided elsewhere.
//
// HBM-bound (bytes = len * 2 * 128 * 2 per (sample, head)); tensor cores have nothing to do at one query row per head.
// One CTA per (head, sample, split): a producer warp streams 64-key K and V tiles (16 KB each, contiguous in the arena)
// with 1-D TMA bulk copies into a 3-stage shared-memory ring guarded by mbarriers; two groups of four consumer warps
// take alternate tiles, each warp owns 16 keys of its tile and keeps a private online-softmax state, merged once at the
// end (short ragged caches are latency-bound: the two groups halve the dependent chain per CTA).  Per-sample lengths come from the device
// (`kv_len`), so a decode step needs no host sync and is CUDA-graph replayable; CTAs whose split is empty exit early.
//
// `quirk_pad` reproduces the reference's B>1 batching artefact: every layer's cache is zero-padded to
// the batch maximum and the pad is NOT masked, i.e. (max_b len - len_b) extra keys with score 0 and value 0.
#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define AD_HD 128
#define AD_TK 64
#define AD_STAGES 3
#define AD_GROUPS 2                 // consumer warp groups; group g takes tiles t = g, g + AD_GROUPS, ...
#define AD_CWARPS (4 * AD_GROUPS)   // 4 warps x 16 keys cover one 64-key tile
#define AD_THREADS ((AD_CWARPS + 1) * 32)
#define AD_STAGE_BYTES (2 * AD_TK * AD_HD * 2)

struct AdSmem {
  alignas(128) bf16 kv[AD_STAGES][2][AD_TK * AD_HD];  // [stage][K|V][key][dim]
  alignas(8) uint64_t full[AD_STAGES];
  alignas(8) uint64_t empty[AD_STAGES];
  float m[AD_CWARPS], l[AD_CWARPS];
  float acc[AD_CWARPS][AD_HD];
  float qs[AD_HD], kn[AD_HD], vn[AD_HD];  // fused mode: this (sample, head)'s rotated+scaled q and the new token's k, v
};

// Fused front end (decode step): the QKV projection arrives as the skinny GEMM's split-K slices; each (sample, head) CTA
// sums ITS 3 x 128 values in fixed order, rounds to bf16 (nn.Linear output), applies RoPE with the reference's rounding
// (dynamic_modeling_llama.py:253-285), keeps q/k_new/v_new in shared memory and — split 0 only — appends k/v to the arena
// when the token is kept (cache_utils.py:153-241).  This removes the separate rope/append kernel from the step.
struct AdFused {
  const float* partial;  // [n_slices][B][3H] or null
  int n_slices;
  const int32_t* pos;    // [B]
  const bf16* cos_t;
  const bf16* sin_t;
  const uint8_t* keep;   // [B] or null
  bf16* k_arena_w;
  bf16* v_arena_w;
  // optional: while this (latency-bound) kernel runs, HBM is mostly idle — every CTA asks for one slice of the NEXT GEMM's
  // weight (o_proj, 33 MB at 7B: fits the 126 MB L2) to be pulled into L2 with a non-blocking bulk prefetch
  const uint8_t* l2_prefetch;
  int64_t l2_prefetch_bytes;
  // host-known lower bound of every sample's cached length for this layer (the prefill length).  Rows below it were
  // written long before this decode step, so (n_split == 1 only) the producer may start their TMA loads BEFORE
  // griddepcontrol.wait, overlapping the K/V fetch with the tail of the QKV GEMM.
  int len_floor;
};

__global__ void __launch_bounds__(AD_THREADS) attn_decode_kernel(
    const bf16* __restrict__ q, int q_row_stride,  // q[b] at q + b*q_row_stride, [heads*128]
    const bf16* __restrict__ k_arena, const bf16* __restrict__ v_arena, const bf16* __restrict__ k_side,
    const bf16* __restrict__ v_side, const int32_t* __restrict__ kv_len, bf16* __restrict__ out, float* __restrict__ ws_acc,
    float* __restrict__ ws_ml, int B, int heads, int cap, int n_split, float scale, int quirk_pad, const AdFused fz) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  AdSmem& S = *reinterpret_cast<AdSmem*>(smem_raw);
  const int h = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int n_pre = 0;  // tiles whose loads were issued before the dependency wait (producer lane only)
  if (threadIdx.x == AD_CWARPS * 32) {  // the producer lane owns barrier init so that it can start loading immediately
    for (int s = 0; s < AD_STAGES; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], AD_CWARPS);
    }
    fence_barrier_init();
    if (n_split == 1) {
      n_pre = min(AD_STAGES, fz.len_floor / AD_TK);
      const bf16* kb = k_arena + ((int64_t)b * heads + h) * cap * AD_HD;
      const bf16* vb = v_arena + ((int64_t)b * heads + h) * cap * AD_HD;
      for (int t = 0; t < n_pre; ++t) {
        mbar_expect_tx(&S.full[t], 2 * AD_TK * AD_HD * 2);
        tma_bulk_g2s(S.kv[t][0], kb + (int64_t)t * AD_TK * AD_HD, AD_TK * AD_HD * 2, &S.full[t]);
        tma_bulk_g2s(S.kv[t][1], vb + (int64_t)t * AD_TK * AD_HD, AD_TK * AD_HD * 2, &S.full[t]);
      }
    }
  }
  pdl_launch_dependents();
  pdl_wait();  // q, side k/v, kv_len and the arena rows of this step come from predecessor kernels
  // every scalar this CTA needs goes out in ONE round trip (len, position, keep flag), and the cos/sin row that depends on the
  // position right behind it, so that the front end below costs two dependent L2 round trips instead of four
  const bool fused = fz.partial != nullptr;
  const int len = kv_len[b];
  const int pos_b = fused ? fz.pos[b] : 0;
  const bool persist = fused && (fz.keep == nullptr || fz.keep[b] != 0);
  int kps = (len + n_split - 1) / n_split;
  kps = ((kps + AD_TK - 1) / AD_TK) * AD_TK;
  const int k0 = split * kps;
  const int k1 = min(len, k0 + kps);
  const int n_keys = max(k1 - k0, 0);
  const int n_tiles = (n_keys + AD_TK - 1) / AD_TK;
  const bf16* kbase = k_arena + (((int64_t)b * heads + h) * cap + k0) * AD_HD;
  const bf16* vbase = v_arena + (((int64_t)b * heads + h) * cap + k0) * AD_HD;
  // The cached length is known now: every remaining K/V tile that still has a free ring stage goes out BEFORE the QKV reduce +
  // RoPE front end below (two dependent L2 round trips and a block barrier), not after it.  Rows below `len` were written by
  // earlier steps; this step's own row (slot `len`) is served from shared memory.
  int n_issued = n_pre;
  if (threadIdx.x == AD_CWARPS * 32) {
    for (; n_issued < n_tiles && n_issued < AD_STAGES; ++n_issued) {
      const int keys = min(AD_TK, n_keys - n_issued * AD_TK);
      const uint32_t bytes = keys * AD_HD * 2;
      mbar_expect_tx(&S.full[n_issued], 2 * bytes);
      tma_bulk_g2s(S.kv[n_issued][0], kbase + (int64_t)n_issued * AD_TK * AD_HD, bytes, &S.full[n_issued]);
      tma_bulk_g2s(S.kv[n_issued][1], vbase + (int64_t)n_issued * AD_TK * AD_HD, bytes, &S.full[n_issued]);
    }
  }
  if (fz.l2_prefetch != nullptr && threadIdx.x == 32) {
    const int64_t n_cta = (int64_t)gridDim.x * gridDim.y * gridDim.z;
    const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    int64_t chunk = ((fz.l2_prefetch_bytes + n_cta - 1) / n_cta + 15) & ~(int64_t)15;
    const int64_t lo = cta * chunk;
    if (lo < fz.l2_prefetch_bytes) {
      chunk = min(chunk, fz.l2_prefetch_bytes - lo) & ~(int64_t)15;
      if (chunk > 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(fz.l2_prefetch + lo), "r"((uint32_t)chunk) : "memory");
    }
  }

  if (fused && threadIdx.x < 192) {
    const int item = threadIdx.x >> 6, d = threadIdx.x & 63;  // 0 = q, 1 = k, 2 = v; dims d and d + 64
    if (item == 0 || split == 0) {
      const int H = heads * AD_HD;
      const int64_t slice = (int64_t)B * 3 * H;
      const float* base = fz.partial + (int64_t)b * 3 * H + (int64_t)item * H + h * AD_HD + d;
      float c = 1.f, sn = 0.f;
      if (item < 2) {  // issued before the slice loads are consumed: both latencies overlap
        c = bf2f(fz.cos_t[(int64_t)pos_b * AD_HD + d]);
        sn = bf2f(fz.sin_t[(int64_t)pos_b * AD_HD + d]);
      }
      float x1 = 0.f, x2 = 0.f;
      for (int s0 = 0; s0 < fz.n_slices; s0 += 8) {
        float t1[8], t2[8];
#pragma unroll
        for (int g = 0; g < 8; ++g)
          if (s0 + g < fz.n_slices) {
            t1[g] = base[(int64_t)(s0 + g) * slice];
            t2[g] = base[(int64_t)(s0 + g) * slice + 64];
          }
#pragma unroll
        for (int g = 0; g < 8; ++g)
          if (s0 + g < fz.n_slices) {
            x1 += t1[g];
            x2 += t2[g];
          }
      }
      x1 = rbf(x1);
      x2 = rbf(x2);
      const int64_t arow = (((int64_t)b * heads + h) * cap + len) * AD_HD;
      if (item < 2) {
        const float o1 = rbf(rbf(x1 * c) + rbf(-x2 * sn));
        const float o2 = rbf(rbf(x2 * c) + rbf(x1 * sn));
        if (item == 0) {
          S.qs[d] = o1 * scale;
          S.qs[d + 64] = o2 * scale;
        } else {
          S.kn[d] = o1;
          S.kn[d + 64] = o2;
          if (persist) {
            fz.k_arena_w[arow + d] = f2bf(o1);
            fz.k_arena_w[arow + d + 64] = f2bf(o2);
          }
        }
      } else {
        S.vn[d] = x1;
        S.vn[d + 64] = x2;
        if (persist) {
          fz.v_arena_w[arow + d] = f2bf(x1);
          fz.v_arena_w[arow + d + 64] = f2bf(x2);
        }
      }
    }
  }
  __syncthreads();

  if (warp == AD_CWARPS) {
    // ---------------- producer warp: TMA bulk copies ----------------
    if (lane == 0) {
      for (int t = n_issued; t < n_tiles; ++t) {
        const int s = t % AD_STAGES;
        const uint32_t ph = (t / AD_STAGES) & 1;
        mbar_wait(&S.empty[s], ph ^ 1);
        const int keys = min(AD_TK, n_keys - t * AD_TK);
        const uint32_t bytes = keys * AD_HD * 2;
        mbar_expect_tx(&S.full[s], 2 * bytes);
        tma_bulk_g2s(S.kv[s][0], kbase + (int64_t)t * AD_TK * AD_HD, bytes, &S.full[s]);
        tma_bulk_g2s(S.kv[s][1], vbase + (int64_t)t * AD_TK * AD_HD, bytes, &S.full[s]);
      }
      // a len_floor above the real length (a stale host constant baked into a replayed graph) leaves pre-issued tiles that no consumer
      // waits for: drain them here so that the CTA never exits with bulk copies in flight into its shared memory
      for (int t = n_tiles; t < n_pre; ++t) mbar_wait(&S.full[t], 0);
    }
  } else {
    // ---------------- consumer warps ----------------
    const int cidx = lane & 15, sub = lane >> 4;
    float qf[8];
    if (fused) {
#pragma unroll
      for (int e = 0; e < 8; ++e) qf[e] = S.qs[cidx * 8 + e];
    } else {
      unpack8(ld8(q + (int64_t)b * q_row_stride + h * AD_HD + cidx * 8), qf);
#pragma unroll
      for (int e = 0; e < 8; ++e) qf[e] *= scale;
    }
    float m_run = -INFINITY, l_run = 0.f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};

    // Every consumer warp observes EVERY phase of every stage barrier (parity waits alias if a waiter skips a phase);
    // only the owning group (t mod AD_GROUPS) does the math, the other group releases the stage at once.
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % AD_STAGES;
      const uint32_t ph = (t / AD_STAGES) & 1;
      mbar_wait(&S.full[s], ph);
      const int keys = min(AD_TK, n_keys - t * AD_TK);
      const int wk0 = (warp & 3) * 16;  // first key of this warp inside the tile
      if ((t % AD_GROUPS) == (warp >> 2) && wk0 < keys) {
        const bf16* Ks = S.kv[s][0];
        const bf16* Vs = S.kv[s][1];
        float sc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kk = wk0 + 2 * i + sub;
          float kf[8];
          unpack8(*reinterpret_cast<const bf16x8*>(Ks + kk * AD_HD + cidx * 8), kf);
          float d = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) d += qf[e] * kf[e];
          d += __shfl_xor_sync(0xffffffffu, d, 8);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          sc[i] = (kk < keys) ? d : -INFINITY;
        }
        float mt = sc[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) mt = fmaxf(mt, sc[i]);
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
        const float m_new = fmaxf(m_run, mt);  // finite: key wk0 is valid
        const float alpha = __expf(m_run - m_new);
        float pm[8], po[8], ls = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pm[i] = __expf(sc[i] - m_new);
          ls += pm[i];
        }
        ls += __shfl_xor_sync(0xffffffffu, ls, 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) po[i] = __shfl_xor_sync(0xffffffffu, pm[i], 16);
        l_run = l_run * alpha + ls;
        m_run = m_new;
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] *= alpha;
        const int nk = min(16, keys - wk0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int kk = 2 * i + u;
            if (kk < nk) {
              const float p = (u == sub) ? pm[i] : po[i];
              const uint2 raw = *reinterpret_cast<const uint2*>(Vs + (wk0 + kk) * AD_HD + lane * 4);
              const float2 v01 = bf2f2(*reinterpret_cast<const bf16x2*>(&raw.x));
              const float2 v23 = bf2f2(*reinterpret_cast<const bf16x2*>(&raw.y));
              acc[0] += p * v01.x;
              acc[1] += p * v01.y;
              acc[2] += p * v23.x;
              acc[3] += p * v23.y;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[s]);
    }

    // the new token itself (side buffer); handled once, by warp 0 of split 0
    if (warp == 0 && split == 0 && (fused || k_side != nullptr)) {
      float d = 0.f;
      if (sub == 0) {
        float kf[8];
        if (fused) {
#pragma unroll
          for (int e = 0; e < 8; ++e) kf[e] = S.kn[cidx * 8 + e];
        } else {
          unpack8(ld8(k_side + ((int64_t)b * heads + h) * AD_HD + cidx * 8), kf);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) d += qf[e] * kf[e];
      }
      d += __shfl_xor_sync(0xffffffffu, d, 8);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d = __shfl_sync(0xffffffffu, d, 0);
      const float m_new = fmaxf(m_run, d);
      const float alpha = __expf(m_run - m_new);
      const float p = __expf(d - m_new);
      float2 v01, v23;
      if (fused) {
        v01 = make_float2(S.vn[lane * 4], S.vn[lane * 4 + 1]);
        v23 = make_float2(S.vn[lane * 4 + 2], S.vn[lane * 4 + 3]);
      } else {
        const uint2 raw = *reinterpret_cast<const uint2*>(v_side + ((int64_t)b * heads + h) * AD_HD + lane * 4);
        v01 = bf2f2(*reinterpret_cast<const bf16x2*>(&raw.x));
        v23 = bf2f2(*reinterpret_cast<const bf16x2*>(&raw.y));
      }
      acc[0] = acc[0] * alpha + p * v01.x;
      acc[1] = acc[1] * alpha + p * v01.y;
      acc[2] = acc[2] * alpha + p * v23.x;
      acc[3] = acc[3] * alpha + p * v23.y;
      l_run = l_run * alpha + p;
      m_run = m_new;
    }
    if (lane == 0) {
      S.m[warp] = m_run;
      S.l[warp] = l_run;
    }
    *reinterpret_cast<float4*>(&S.acc[warp][lane * 4]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
  __syncthreads();

  // ---------------- merge the four warp states (+ the zero-pad pseudo keys) ----------------
  if (threadIdx.x < AD_HD) {
    const int d = threadIdx.x;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < AD_CWARPS; ++w) M = fmaxf(M, S.m[w]);
    int n_pad = 0;
    if (quirk_pad && split == 0) {
      int mx = 0;
      for (int i = 0; i < B; ++i) mx = max(mx, kv_len[i]);
      n_pad = mx - len;
      if (n_pad > 0) M = fmaxf(M, 0.f);
    }
    float L = 0.f, O = 0.f;
    if (M > -INFINITY) {
#pragma unroll
      for (int w = 0; w < AD_CWARPS; ++w) {
        const float f = (S.m[w] > -INFINITY) ? __expf(S.m[w] - M) : 0.f;
        L += S.l[w] * f;
        O += S.acc[w][d] * f;
      }
      if (n_pad > 0) L += (float)n_pad * __expf(0.f - M);
    }
    if (n_split == 1) {
      out[(int64_t)b * heads * AD_HD + h * AD_HD + d] = f2bf(O / L);
    } else {
      const int64_t slot = ((int64_t)b * heads + h) * n_split + split;
      ws_acc[slot * AD_HD + d] = O;
      if (d == 0) {
        ws_ml[slot * 2] = M;
        ws_ml[slot * 2 + 1] = L;
      }
    }
  }
}

__global__ void __launch_bounds__(AD_HD) attn_decode_combine_kernel(const float* __restrict__ ws_acc, const float* __restrict__ ws_ml,
                                                                    bf16* __restrict__ out, int heads, int n_split) {
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  const int64_t base = ((int64_t)b * heads + h) * n_split;
  float M = -INFINITY;
  for (int s = 0; s < n_split; ++s) M = fmaxf(M, ws_ml[(base + s) * 2]);
  float L = 0.f, O = 0.f;
  for (int s = 0; s < n_split; ++s) {
    const float m = ws_ml[(base + s) * 2];
    const float f = (m > -INFINITY) ? __expf(m - M) : 0.f;
    L += ws_ml[(base + s) * 2 + 1] * f;
    O += ws_acc[(base + s) * AD_HD + d] * f;
  }
  out[(int64_t)b * heads * AD_HD + h * AD_HD + d] = f2bf(O / L);
}

extern "C" int64_t dl_attn_decode_workspace_bytes(int B, int heads, int n_split) {
  return n_split <= 1 ? 0 : (int64_t)B * heads * n_split * (AD_HD + 2) * (int64_t)sizeof(float);
}

extern "C" int dl_attn_decode(const void* q, int q_row_stride, const void* k_arena, const void* v_arena, const void* k_side,
                              const void* v_side, const int32_t* kv_len, void* out, void* workspace, int B, int heads,
                              int head_dim, int cap, int n_split, float scale, int quirk_pad, cudaStream_t stream) {
  DL_REQUIRE(head_dim == AD_HD && n_split >= 1 && (n_split == 1 || workspace != nullptr), "dl_attn_decode");
  if (B == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AdSmem));
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_attn_decode(smem attr)");
    attr_set = true;
  }
  float* ws_acc = (float*)workspace;
  float* ws_ml = ws_acc ? ws_acc + (int64_t)B * heads * n_split * AD_HD : nullptr;
  AdFused fz = {};
  dl_launch_pdl(attn_decode_kernel, dim3(heads, B, n_split), dim3(AD_THREADS), sizeof(AdSmem), stream, (const bf16*)q, q_row_stride,
                (const bf16*)k_arena, (const bf16*)v_arena, (const bf16*)k_side, (const bf16*)v_side, kv_len, (bf16*)out, ws_acc, ws_ml,
                B, heads, cap, n_split, scale, quirk_pad, fz);
  if (n_split > 1)
    dl_launch_pdl(attn_decode_combine_kernel, dim3(heads, B), dim3(AD_HD), 0, stream, (const float*)ws_acc, (const float*)ws_ml,
                  (bf16*)out, heads, n_split);
  DL_CHECK_LAUNCH("dl_attn_decode");
}

// Decode step, fused: split-K QKV slices -> (reduce, RoPE, conditional append) -> ragged attention, one launch per layer.
extern "C" int dl_attn_decode_fused(const float* qkv_partial, int n_slices, const int32_t* pos, const uint8_t* keep,
                                    const void* cos_table, const void* sin_table, void* k_arena, void* v_arena,
                                    const int32_t* kv_len, void* out, void* workspace, int B, int heads, int head_dim, int cap,
                                    int n_split, float scale, int quirk_pad, const void* l2_prefetch, int64_t l2_prefetch_bytes,
                                    int len_floor, cudaStream_t stream) {
  DL_REQUIRE(head_dim == AD_HD && n_split >= 1 && (n_split == 1 || workspace != nullptr) && qkv_partial != nullptr,
             "dl_attn_decode_fused");
  if (B == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AdSmem));
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_attn_decode_fused(smem attr)");
    attr_set = true;
  }
  float* ws_acc = (float*)workspace;
  float* ws_ml = ws_acc ? ws_acc + (int64_t)B * heads * n_split * AD_HD : nullptr;
  AdFused fz;
  fz.partial = qkv_partial;
  fz.n_slices = n_slices;
  fz.pos = pos;
  fz.cos_t = (const bf16*)cos_table;
  fz.sin_t = (const bf16*)sin_table;
  fz.keep = keep;
  fz.k_arena_w = (bf16*)k_arena;
  fz.v_arena_w = (bf16*)v_arena;
  fz.l2_prefetch = (const uint8_t*)l2_prefetch;
  fz.l2_prefetch_bytes = l2_prefetch_bytes;
  fz.len_floor = len_floor > 0 ? len_floor : 0;
  dl_launch_pdl(attn_decode_kernel, dim3(heads, B, n_split), dim3(AD_THREADS), sizeof(AdSmem), stream, (const bf16*)nullptr, 0,
                (const bf16*)k_arena, (const bf16*)v_arena, (const bf16*)nullptr, (const bf16*)nullptr, kv_len, (bf16*)out, ws_acc, ws_ml,
                B, heads, cap, n_split, scale, quirk_pad, fz);
  if (n_split > 1)
    dl_launch_pdl(attn_decode_combine_kernel, dim3(heads, B), dim3(AD_HD), 0, stream, (const float*)ws_acc, (const float*)ws_ml,
                  (bf16*)out, heads, n_split);
  DL_CHECK_LAUNCH("dl_attn_decode_fused");
}

