// K4 (RoPE) + K5 (conditional KV append), fused: one pass over the fused QKV projection output.
//
//  * q, k rotated with the reference's half-split layout and bf16 rounding order
//    (dynamic_modeling_llama.py:253-285: bf16(q*cos) + bf16(rotate_half(q)*sin), cos/sin tables in bf16);
//  * rotated k and v written straight into the ragged KV arena [B, heads, cap, hd] at slot kv_len[b] + j
//    (cache_utils.py:139-149 prefill / :242-250 plain append), or only when keep[n] is set
//    (cache_utils.py:153-241: the decode-time `text_decision`); kv_len itself is advanced later by dl_kv_advance so
//    the whole decode step stays free of host syncs and is CUDA-graph replayable;
//  * the current token's k/v are ALSO written to a side buffer when requested, because the reference always attends
//    to the new token even when it is not persisted (cache_utils.py:256-268 get_cache).
//
// The reference grows each layer's cache with torch.cat — an O(len) copy of K and V per layer per step; this append
// moves 2*H*2 bytes per kept (sample, layer).
#include "common.cuh"
#include "../../include/dynllava_b200.h"

template <int HD>
__global__ void __launch_bounds__(256) rope_append_kernel(bf16* __restrict__ qkv, const int32_t* __restrict__ pos,
                                                          const int32_t* __restrict__ cu_seqlens,
                                                          const int32_t* __restrict__ kv_len, const uint8_t* __restrict__ keep,
                                                          const bf16* __restrict__ cos_t, const bf16* __restrict__ sin_t,
                                                          bf16* __restrict__ k_arena, bf16* __restrict__ v_arena,
                                                          bf16* __restrict__ k_side, bf16* __restrict__ v_side, int B,
                                                          int heads, int cap) {
  constexpr int HALF = HD / 2;
  constexpr int PAIRS = HALF / 8;   // 128-bit chunk pairs (d, d + HD/2) per head
  constexpr int VCH = HD / 8;       // 128-bit chunks per head of v
  const int n = blockIdx.x;
  const int H = heads * HD;
  // sample of this token (CSR header lookup; B is small)
  int b = 0;
  while (b + 1 < B && cu_seqlens[b + 1] <= n) ++b;
  const int j = n - cu_seqlens[b];
  const bool persist = keep == nullptr || keep[n] != 0;
  const int slot = kv_len[b] + j;
  const int p = pos[n];
  bf16* row = qkv + (int64_t)n * 3 * H;
  const bf16* cs = cos_t + (int64_t)p * HD;
  const bf16* sn = sin_t + (int64_t)p * HD;
  const int n_rot = heads * PAIRS;
  const int items = 2 * n_rot + heads * VCH;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    if (it < 2 * n_rot) {
      const int which = it / n_rot;  // 0 = q, 1 = k
      const int r = it - which * n_rot;
      const int h = r / PAIRS, c = r % PAIRS;
      bf16* base = row + which * H + h * HD + c * 8;
      float x1[8], x2[8], cc[8], ss[8], o1[8], o2[8];
      unpack8(ld8(base), x1);
      unpack8(ld8(base + HALF), x2);
      unpack8(ld8(cs + c * 8), cc);  // cos[d] == cos[d + HD/2] (emb = cat(freqs, freqs))
      unpack8(ld8(sn + c * 8), ss);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o1[e] = rbf(x1[e] * cc[e]) + rbf(-x2[e] * ss[e]);
        o2[e] = rbf(x2[e] * cc[e]) + rbf(x1[e] * ss[e]);
      }
      const bf16x8 p1 = pack8(o1), p2 = pack8(o2);
      if (which == 0) {
        st8(base, p1);
        st8(base + HALF, p2);
      } else {
        if (persist) {
          bf16* dst = k_arena + (((int64_t)b * heads + h) * cap + slot) * HD + c * 8;
          st8(dst, p1);
          st8(dst + HALF, p2);
        }
        if (k_side != nullptr) {
          bf16* dst = k_side + ((int64_t)n * heads + h) * HD + c * 8;
          st8(dst, p1);
          st8(dst + HALF, p2);
        }
      }
    } else {
      const int r = it - 2 * n_rot;
      const int h = r / VCH, c = r % VCH;
      const bf16x8 v = ld8(row + 2 * H + h * HD + c * 8);
      if (persist) st8(v_arena + (((int64_t)b * heads + h) * cap + slot) * HD + c * 8, v);
      if (v_side != nullptr) st8(v_side + ((int64_t)n * heads + h) * HD + c * 8, v);
    }
  }
}

extern "C" int dl_rope_append(void* qkv, const int32_t* pos, const int32_t* cu_seqlens, const int32_t* kv_len,
                              const uint8_t* keep, const void* cos_table, const void* sin_table, void* k_arena, void* v_arena,
                              void* k_side, void* v_side, int n_tok, int B, int heads, int head_dim, int cap,
                              cudaStream_t stream) {
  DL_REQUIRE(head_dim == 128, "dl_rope_append");
  DL_REQUIRE(keep == nullptr || n_tok == B, "dl_rope_append(keep needs one token per sample)");
  if (n_tok == 0) return 0;
  rope_append_kernel<128><<<n_tok, 256, 0, stream>>>((bf16*)qkv, pos, cu_seqlens, kv_len, keep, (const bf16*)cos_table,
                                                     (const bf16*)sin_table, (bf16*)k_arena, (bf16*)v_arena, (bf16*)k_side,
                                                     (bf16*)v_side, B, heads, cap);
  DL_CHECK_LAUNCH("dl_rope_append");
}

// Decode-step variant: the fused QKV projection arrives as the skinny GEMM's split-K fp32 slices
// partial[s][b][3H]; sum them in fixed order, round to bf16 (the nn.Linear output rounding), then rotate / append as above.
// One token per sample; q goes to a dense [B, H] buffer.
template <int HD>
__global__ void __launch_bounds__(256) dec_reduce_rope_append_kernel(const float* __restrict__ partial, int S,
                                                                     bf16* __restrict__ q_out, const int32_t* __restrict__ pos,
                                                                     const int32_t* __restrict__ kv_len,
                                                                     const uint8_t* __restrict__ keep,
                                                                     const bf16* __restrict__ cos_t, const bf16* __restrict__ sin_t,
                                                                     bf16* __restrict__ k_arena, bf16* __restrict__ v_arena,
                                                                     bf16* __restrict__ k_side, bf16* __restrict__ v_side, int B,
                                                                     int heads, int cap) {
  constexpr int HALF = HD / 2;
  constexpr int PAIRS = HALF / 8;
  constexpr int VCH = HD / 8;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int H = heads * HD;
  const int64_t slice = (int64_t)B * 3 * H;
  const bool persist = keep == nullptr || keep[b] != 0;
  const int slot = kv_len[b];
  const int p = pos[b];
  const float* row = partial + (int64_t)b * 3 * H;
  const bf16* cs = cos_t + (int64_t)p * HD;
  const bf16* sn = sin_t + (int64_t)p * HD;
  const int n_rot = heads * PAIRS;
  // blockIdx.y selects the part of the row this CTA handles: 0 = q pairs, 1 = k pairs, 2 = v chunks (3x the CTAs in
  // flight; with one token per sample the kernel is latency-, not bandwidth-bound)
  const int it0 = blockIdx.y == 0 ? 0 : (blockIdx.y == 1 ? n_rot : 2 * n_rot);
  const int it1 = blockIdx.y == 0 ? n_rot : (blockIdx.y == 1 ? 2 * n_rot : 2 * n_rot + heads * VCH);
  auto load8 = [&](const float* src, float* f) {
    const float4 a = sum_slices_f4(src, slice, S), c = sum_slices_f4(src + 4, slice, S);
    f[0] = rbf(a.x); f[1] = rbf(a.y); f[2] = rbf(a.z); f[3] = rbf(a.w);
    f[4] = rbf(c.x); f[5] = rbf(c.y); f[6] = rbf(c.z); f[7] = rbf(c.w);
  };
  for (int it = it0 + threadIdx.x; it < it1; it += blockDim.x) {
    if (it < 2 * n_rot) {
      const int which = it / n_rot;
      const int r = it - which * n_rot;
      const int h = r / PAIRS, c = r % PAIRS;
      const float* base = row + which * H + h * HD + c * 8;
      float x1[8], x2[8], cc[8], ss[8], o1[8], o2[8];
      load8(base, x1);
      load8(base + HALF, x2);
      unpack8(ld8(cs + c * 8), cc);
      unpack8(ld8(sn + c * 8), ss);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o1[e] = rbf(x1[e] * cc[e]) + rbf(-x2[e] * ss[e]);
        o2[e] = rbf(x2[e] * cc[e]) + rbf(x1[e] * ss[e]);
      }
      const bf16x8 p1 = pack8(o1), p2 = pack8(o2);
      if (which == 0) {
        bf16* dst = q_out + (int64_t)b * H + h * HD + c * 8;
        st8(dst, p1);
        st8(dst + HALF, p2);
      } else {
        if (persist) {
          bf16* dst = k_arena + (((int64_t)b * heads + h) * cap + slot) * HD + c * 8;
          st8(dst, p1);
          st8(dst + HALF, p2);
        }
        bf16* dst = k_side + ((int64_t)b * heads + h) * HD + c * 8;
        st8(dst, p1);
        st8(dst + HALF, p2);
      }
    } else {
      const int r = it - 2 * n_rot;
      const int h = r / VCH, c = r % VCH;
      float f[8];
      load8(row + 2 * H + h * HD + c * 8, f);
      const bf16x8 v = pack8(f);
      if (persist) st8(v_arena + (((int64_t)b * heads + h) * cap + slot) * HD + c * 8, v);
      st8(v_side + ((int64_t)b * heads + h) * HD + c * 8, v);
    }
  }
}

extern "C" int dl_dec_reduce_rope_append(const float* partial, int splits, void* q_out, const int32_t* pos,
                                         const int32_t* kv_len, const uint8_t* keep, const void* cos_table,
                                         const void* sin_table, void* k_arena, void* v_arena, void* k_side, void* v_side, int B,
                                         int heads, int head_dim, int cap, cudaStream_t stream) {
  DL_REQUIRE(head_dim == 128 && k_side != nullptr && v_side != nullptr, "dl_dec_reduce_rope_append");
  if (B == 0) return 0;
  dl_launch_pdl(dec_reduce_rope_append_kernel<128>, dim3(B, 3), dim3(256), 0, stream, partial, splits, (bf16*)q_out, pos, kv_len, keep,
                (const bf16*)cos_table, (const bf16*)sin_table, (bf16*)k_arena, (bf16*)v_arena, (bf16*)k_side, (bf16*)v_side, B,
                heads, cap);
  DL_CHECK_LAUNCH("dl_dec_reduce_rope_append");
}

// kv_len[l][b] += tokens persisted by this forward pass:
//   l <  sparse_layer                  : n_lo, always (layers below the sparse layer never drop)
//   l >= sparse_layer, keep == null    : n_hi (prefill: n_hi = P' < n_lo = P; plain decode: 1)
//   l >= sparse_layer, keep != null    : n_hi only where keep[b]
// also pos[b] += n_lo (decode position = full-length count, fact 8) and step_counter += 1.
__global__ void kv_advance_kernel(int32_t* __restrict__ kv_len, int32_t* __restrict__ pos, int32_t* __restrict__ step_counter,
                                  const uint8_t* __restrict__ keep, int L, int B, int sparse_layer, int n_lo, int n_hi) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < L * B) {
    const int l = idx / B, b = idx % B;
    if (l < sparse_layer)
      kv_len[idx] += n_lo;
    else if (keep == nullptr || keep[b] != 0)
      kv_len[idx] += n_hi;
  }
  if (idx < B && pos != nullptr) pos[idx] += n_lo;
  if (idx == 0 && step_counter != nullptr) step_counter[0] += 1;
}
extern "C" int dl_kv_advance(int32_t* kv_len, int32_t* pos, int32_t* step_counter, const uint8_t* keep, int L, int B,
                             int sparse_layer, int n_lo, int n_hi, cudaStream_t stream) {
  const int n = L * B > B ? L * B : B;
  if (n == 0) return 0;
  dl_launch_pdl(kv_advance_kernel, dim3((n + 127) / 128), dim3(128), 0, stream, kv_len, pos, step_counter, keep, L, B, sparse_layer, n_lo, n_hi);
  DL_CHECK_LAUNCH("dl_kv_advance");
}
