// Output-token thinning of the no-KV-cache mode (`use_cache=False`; dynamic_modeling_llama.py:2393-2504).
//
// Every call re-feeds the whole sequence.  At the sparse layer the generated tokens between the sticky `answer_indice` A and the
// last token (n = Lq - 1 - A of them) are scored by the output text predictor; every sample keeps the same number of them,
// n_keep = max_b #{logit0 > logit1} (:2413-2417), chosen by the largest channel-0 logits and kept in sequence order
// (`sort(argsort(score, descending)[:, :n_keep])`, :2420-2426).  The host needs n_keep for the shapes that follow (so does the
// reference: its slice by a 0-dim tensor synchronises), hence two small kernels around one B-int read:
//   dl_output_rank        rank[b,i] = #{j : s_j > s_i or (s_j == s_i and j < i)}  (a permutation of 0..n-1; ties: lower index first,
//                         torch's argsort is not stable there), count[b] = #{logit0 > logit1}
//   dl_output_keep_index  keep_index[b,:] = ascending {i : rank[b,i] < n_keep}; rows[b,:] = source rows of the thinned sequence
//                         [0, A) ++ A + keep_index ++ [Lq - 1] as global row numbers b*Lq + t for dl_gather_rows_pos
// Index work on <= a few thousand 2-byte scores per sample: latency-bound, one CTA per sample.
#include "common.cuh"
#include "../../include/dynllava_b200.h"

__global__ void __launch_bounds__(256) output_rank_kernel(const bf16* __restrict__ logits, int32_t* __restrict__ rank,
                                                          int32_t* __restrict__ count, int n) {
  extern __shared__ float s_score[];
  __shared__ int s_cnt[8];
  const int b = blockIdx.x;
  const bf16* lg = logits + (int64_t)b * n * 2;
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float l0 = bf2f(lg[2 * i]), l1 = bf2f(lg[2 * i + 1]);
    s_score[i] = l0;
    c += l0 > l1 ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_cnt[w];
    count[b] = t;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float si = s_score[i];
    int r = 0;
    for (int j = 0; j < n; ++j) {  // broadcast reads: every lane looks at the same j
      const float sj = s_score[j];
      r += (sj > si || (sj == si && j < i)) ? 1 : 0;
    }
    rank[(int64_t)b * n + i] = r;
  }
}

extern "C" int dl_output_rank(const void* logits, int32_t* rank, int32_t* count, int B, int n, cudaStream_t stream) {
  DL_REQUIRE(B > 0 && n > 0 && n <= 8192, "dl_output_rank");
  output_rank_kernel<<<B, 256, n * sizeof(float), stream>>>((const bf16*)logits, rank, count, n);
  DL_CHECK_LAUNCH("dl_output_rank");
}

__global__ void __launch_bounds__(256) output_keep_index_kernel(const int32_t* __restrict__ rank, int n, int n_keep, int A, int Lq,
                                                                int64_t* __restrict__ keep_index, int64_t* __restrict__ rows) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ln = A + n_keep + 1;
  int64_t* row_b = rows + (int64_t)b * Ln;
  const int64_t g0 = (int64_t)b * Lq;
  for (int t = threadIdx.x; t < A; t += blockDim.x) row_b[t] = g0 + t;
  if (threadIdx.x == 0) {
    row_b[Ln - 1] = g0 + Lq - 1;
    s_base = 0;
  }
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const bool k = i < n && rank[(int64_t)b * n + i] < n_keep;
    const unsigned m = __ballot_sync(0xffffffffu, k);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    off += __popc(m & ((1u << lane) - 1u));
    if (k) {
      if (keep_index != nullptr) keep_index[(int64_t)b * n_keep + off] = i;
      row_b[A + off] = g0 + A + i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
}

extern "C" int dl_output_keep_index(const int32_t* rank, int B, int n, int n_keep, int answer_index, int Lq, int64_t* keep_index,
                                    int64_t* rows, cudaStream_t stream) {
  DL_REQUIRE(B > 0 && n >= 0 && n_keep >= 0 && n_keep <= n && answer_index >= 0 && answer_index <= Lq && Lq >= 1 && rows != nullptr &&
                 (n == 0 || rank != nullptr),
             "dl_output_keep_index");
  output_keep_index_kernel<<<B, 256, 0, stream>>>(rank, n, n_keep, answer_index, Lq, keep_index, rows);
  DL_CHECK_LAUNCH("dl_output_keep_index");
}
