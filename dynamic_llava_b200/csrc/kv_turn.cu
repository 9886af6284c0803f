// Instruct-predictor bookkeeping (sparse_config["use_instruct_predictor"]):
//
//  * prefill: the tokens of the last instruct span whose predictor logit0 <= logit1 leave the residual stream at the
//    sparse layer (dynamic_modeling_llama.py:2261-2375) -> dl_gather_rows_pos moves the surviving rows and their position ids;
//  * a new user turn (q_len > 1 over a cache, :2506-2521): every new token is attended to, but layers >= sparse_layer
//    persist K/V only where the decision is set, the last token always (cache_utils.py:165-241, per sample: slice to the
//    true length, append the kept rows).  The reference rebuilds each layer's cache with torch.cat / zero-pad / stack; here
//    the new rows already sit in the arena behind the cached ones (dl_rope_append), so persisting "only the kept ones" is an
//    in-place left compaction of those Tq rows per (sample, head) plus a per-sample length update.
//
// Byte work, HBM-bound: <= 2 * Tq * H * 2 bytes read and written per layer.
#include "common.cuh"
#include "../../include/dynllava_b200.h"

// keep[n] = forced ? forced[n] : logit0 > logit1 (compared in bf16, like the reference's tensors); with force_last the
// last token of every sample is kept (text_decision[:, -1:] = True, :2520).
__global__ void turn_decision_kernel(const bf16* __restrict__ logits, const uint8_t* __restrict__ forced,
                                     uint8_t* __restrict__ keep, int B, int Tq, int force_last) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B * Tq) return;
  bool d = forced != nullptr ? forced[n] != 0
                             : bf2f(logits[2 * n]) > bf2f(logits[2 * n + 1]);
  if (force_last && (n % Tq) == Tq - 1) d = true;
  keep[n] = d ? 1 : 0;
}

extern "C" int dl_turn_decision(const void* logits, const uint8_t* forced, uint8_t* keep, int B, int Tq, int force_last,
                                cudaStream_t stream) {
  DL_REQUIRE(logits != nullptr || forced != nullptr, "dl_turn_decision");
  if (B * Tq == 0) return 0;
  turn_decision_kernel<<<(B * Tq + 127) / 128, 128, 0, stream>>>((const bf16*)logits, forced, keep, B, Tq, force_last);
  DL_CHECK_LAUNCH("dl_turn_decision");
}

// In-place left compaction of the Tq rows that dl_rope_append put at slots [kv_len[b], kv_len[b] + Tq).
// One CTA per (head, sample, K|V).  Rows move in chunks of CH: every thread first loads its share of the chunk's kept rows,
// the CTA synchronises, then stores — a chunk's destinations never reach past its own last row and later chunks only read
// rows behind it, so one barrier per chunk orders every read before the write that could clobber it.
template <int HD>
__global__ void __launch_bounds__(256) kv_compact_turn_kernel(bf16* __restrict__ k_arena, bf16* __restrict__ v_arena,
                                                              const int32_t* __restrict__ kv_len,
                                                              const uint8_t* __restrict__ keep, int Tq, int heads, int cap) {
  constexpr int VCH = HD / 8;   // 128-bit chunks per row
  constexpr int CH = 64;        // rows per pass
  constexpr int PER = CH * VCH / 256;
  const int h = blockIdx.x, b = blockIdx.y;
  bf16* base = (blockIdx.z == 0 ? k_arena : v_arena) + (((int64_t)b * heads + h) * cap + kv_len[b]) * HD;
  const uint8_t* kp = keep + (int64_t)b * Tq;
  __shared__ int dst[CH];
  int run = 0;  // rows persisted so far (thread 0's copy is the one that counts)
  for (int c0 = 0; c0 < Tq; c0 += CH) {
    const int n = min(CH, Tq - c0);
    if (threadIdx.x == 0) {
      for (int i = 0; i < n; ++i) dst[i] = kp[c0 + i] ? run++ : -1;
    }
    __syncthreads();
    bf16x8 reg[PER];
    int d[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int it = threadIdx.x + i * 256;
      const int r = it / VCH, c = it % VCH;
      d[i] = -1;
      if (r < n && dst[r] >= 0 && dst[r] != c0 + r) {  // a row already in place is not touched
        d[i] = dst[r] * HD + c * 8;
        reg[i] = ld8(base + (int64_t)(c0 + r) * HD + c * 8);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; ++i)
      if (d[i] >= 0) st8(base + d[i], reg[i]);
  }
  // rows [kept, Tq) are no longer part of the cache: zero them, so that "slots at or beyond kv_len are zero" (cache.py: the strided
  // view IS the reference's zero-padded tensor) keeps holding after a turn
  __shared__ int n_kept;
  __syncthreads();
  if (threadIdx.x == 0) n_kept = run;
  __syncthreads();
  bf16x8 z;
#pragma unroll
  for (int i = 0; i < 4; ++i) z.v[i] = f2bf2(0.f, 0.f);
  for (int it = n_kept * VCH + threadIdx.x; it < Tq * VCH; it += 256) st8(base + (int64_t)it * 8, z);
}

extern "C" int dl_kv_compact_turn(void* k_arena, void* v_arena, const int32_t* kv_len, const uint8_t* keep, int Tq, int B,
                                  int heads, int head_dim, int cap, cudaStream_t stream) {
  DL_REQUIRE(head_dim == 128 && keep != nullptr, "dl_kv_compact_turn");
  if (B * Tq == 0) return 0;
  kv_compact_turn_kernel<128><<<dim3(heads, B, 2), 256, 0, stream>>>((bf16*)k_arena, (bf16*)v_arena, kv_len, keep, Tq, heads, cap);
  DL_CHECK_LAUNCH("dl_kv_compact_turn");
}

// true_cache_length after a new turn (cache_utils.py:197-199 / :250): layers < sparse_layer += Tq, layers >= sparse_layer
// += the number of kept tokens of that sample; pos[b] += Tq (positions follow the layer-0 length, :1714-1722).
__global__ void kv_advance_turn_kernel(int32_t* __restrict__ kv_len, int32_t* __restrict__ pos, const uint8_t* __restrict__ keep,
                                       int L, int B, int sparse_layer, int Tq) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < L * B) {
    const int l = idx / B, b = idx % B;
    int n = Tq;
    if (l >= sparse_layer && keep != nullptr) {
      n = 0;
      for (int i = 0; i < Tq; ++i) n += keep[(int64_t)b * Tq + i] != 0;
    }
    kv_len[idx] += n;
  }
  if (idx < B && pos != nullptr) pos[idx] += Tq;
}

extern "C" int dl_kv_advance_turn(int32_t* kv_len, int32_t* pos, const uint8_t* keep, int L, int B, int sparse_layer, int Tq,
                                  cudaStream_t stream) {
  const int n = L * B > B ? L * B : B;
  if (n == 0) return 0;
  kv_advance_turn_kernel<<<(n + 127) / 128, 128, 0, stream>>>(kv_len, pos, keep, L, B, sparse_layer, Tq);
  DL_CHECK_LAUNCH("dl_kv_advance_turn");
}

// out[i] = x[idx[i]], pos_out[i] = pos[idx[i]]: the residual stream and its position ids after the instruct drop
// (dynamic_modeling_llama.py:2286-2337).  One CTA per output row; H % 8 == 0.
__global__ void __launch_bounds__(256) gather_rows_pos_kernel(const bf16* __restrict__ x, const int32_t* __restrict__ pos,
                                                              const int64_t* __restrict__ idx, bf16* __restrict__ out,
                                                              int32_t* __restrict__ pos_out, int H) {
  const int i = blockIdx.x;
  const int64_t src = idx[i];
  const bf16* s = x + src * H;
  bf16* d = out + (int64_t)i * H;
  for (int c = threadIdx.x * 8; c < H; c += blockDim.x * 8) st8(d + c, ld8(s + c));
  if (threadIdx.x == 0 && pos != nullptr) pos_out[i] = pos[src];
}

extern "C" int dl_gather_rows_pos(const void* x, const int32_t* pos, const int64_t* idx, void* out, int32_t* pos_out, int n,
                                  int H, cudaStream_t stream) {
  DL_REQUIRE(H % 8 == 0 && (pos == nullptr) == (pos_out == nullptr), "dl_gather_rows_pos");
  if (n == 0) return 0;
  gather_rows_pos_kernel<<<n, 256, 0, stream>>>((const bf16*)x, pos, idx, (bf16*)out, pos_out, H);
  DL_CHECK_LAUNCH("dl_gather_rows_pos");
}
