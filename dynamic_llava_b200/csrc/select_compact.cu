// K1: vision-token selection + gather/compact in ONE kernel
// (dynamic_modeling_llama.py:1867, :1897-1994): score = log_softmax(logits)[..., 0] (bf16 rounding),
// keep = ascending indices of the k best scores, hidden' = [left | image[keep] | right], RoPE positions stay the
// ORIGINAL (pre-drop) positions, and a CSR header (cu_seqlens) describes the packed result.
//
// Grid = (R, B).  Every CTA of sample b redoes the selection for its sample in shared memory — an O(n) two-pass radix
// select on the 16-bit score keys (256-bin histogram of the high byte, then of the low byte inside the boundary bin),
// followed by an ordered compaction scan; ~2 us, so replicating it per CTA is cheaper than a second kernel or a round
// trip of the index map through HBM — and then streams its share of the P' output rows with 128-bit no-allocate
// loads/stores.  Ties at the k-th score are broken towards the lower index (= stable descending sort).
// HBM-bound: algorithmic bytes = 2 * B * P' * H * 2.
#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define SC_THREADS 256
#define SC_MAX_IMG 1024

__global__ void __launch_bounds__(SC_THREADS) select_compact_kernel(
    const bf16* __restrict__ logits,         // [B, n_img, 2]
    const int64_t* __restrict__ forced_keep, // [B, k] or null
    const bf16* __restrict__ hid_in,         // [B, P, H]
    bf16* __restrict__ hid_out,              // [B, P', H]
    int64_t* __restrict__ keep_index,        // [B, k]
    int32_t* __restrict__ pos_out,           // [B, P']
    int32_t* __restrict__ cu_seqlens,        // [B + 1]
    int P, int H, int s0, int n_img, int k) {
  __shared__ uint16_t key[SC_MAX_IMG];
  __shared__ int16_t kidx[SC_MAX_IMG];
  __shared__ int hist[256];
  __shared__ int wsum[SC_THREADS / 32];
  __shared__ int sel[4];  // [0] boundary bin, [1] #keys above it
  const int b = blockIdx.y, tid = threadIdx.x;
  const int Pp = P - n_img + k;

  if (forced_keep != nullptr) {
    for (int i = tid; i < k; i += SC_THREADS) kidx[i] = (int16_t)forced_keep[(int64_t)b * k + i];
  } else if (k > 0) {
    // scores rounded where the reference's bf16 log_softmax rounds, mapped to order-preserving 16-bit keys
    for (int i = tid; i < n_img; i += SC_THREADS) {
      const __nv_bfloat162 l = *reinterpret_cast<const __nv_bfloat162*>(logits + ((int64_t)b * n_img + i) * 2);
      const float l0 = __low2float(l), l1 = __high2float(l);
      const float m = fmaxf(l0, l1);
      const bf16 sc = f2bf((l0 - m) - logf(expf(l0 - m) + expf(l1 - m)));
      uint16_t bits = *reinterpret_cast<const uint16_t*>(&sc);
      if (bits == 0x8000u) bits = 0;  // -0 == +0
      key[i] = (bits & 0x8000u) ? (uint16_t)~bits : (uint16_t)(bits | 0x8000u);
    }
    // two radix passes: find T = the k-th largest key, c_gt = #keys > T
    int prefix = 0, c_gt = 0, want = k;
    for (int pass = 0; pass < 2; ++pass) {
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      for (int i = tid; i < n_img; i += SC_THREADS) {
        const int kk = key[i];
        if (pass == 0)
          atomicAdd(&hist[kk >> 8], 1);
        else if ((kk >> 8) == prefix)
          atomicAdd(&hist[kk & 255], 1);
      }
      __syncthreads();
      if (tid < 32) {  // suffix scan over the 256 bins (8 per lane), locate the bin that holds the `want`-th largest
        int loc = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) loc += hist[tid * 8 + j];
        int inc = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_down_sync(0xffffffffu, inc, o);
          if (tid + o < 32) inc += t;
        }
        int run = inc - loc;  // keys in bins above this lane's 8
        for (int j = 7; j >= 0; --j) {
          const int h = hist[tid * 8 + j];
          if (run < want && want <= run + h) {
            sel[0] = tid * 8 + j;
            sel[1] = run;
          }
          run += h;
        }
      }
      __syncthreads();
      if (pass == 0) prefix = sel[0];
      c_gt += sel[1];
      want -= sel[1];
      if (pass == 1) prefix = (prefix << 8) | sel[0];
      __syncthreads();
    }
    const int T = prefix, need_eq = k - c_gt;
    // ordered compaction: contiguous chunk per thread, block exclusive scan of (#gt, #eq) packed in one int
    const int per = (n_img + SC_THREADS - 1) / SC_THREADS;
    const int lo = tid * per, hi = min(lo + per, n_img);
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += (key[i] > T) + ((key[i] == T) << 16);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((tid & 31) >= o) inc += t;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = inc;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < (tid >> 5); ++w) base += wsum[w];
    const int excl = base + inc - cnt;
    int gt_before = excl & 0xffff, eq_before = excl >> 16;
    for (int i = lo; i < hi; ++i) {
      const int kk = key[i];
      if (kk > T) {
        kidx[gt_before + min(eq_before, need_eq)] = (int16_t)i;
        ++gt_before;
      } else if (kk == T) {
        if (eq_before < need_eq) kidx[gt_before + eq_before] = (int16_t)i;
        ++eq_before;
      }
    }
  }
  __syncthreads();

  if (blockIdx.x == 0) {  // index map + CSR header
    for (int i = tid; i < k; i += SC_THREADS) keep_index[(int64_t)b * k + i] = kidx[i];
    for (int j = tid; j < Pp; j += SC_THREADS) {
      int p = j < s0 ? j : (j < s0 + k ? s0 + kidx[j - s0] : j + (n_img - k));
      pos_out[(int64_t)b * Pp + j] = p;
    }
    if (tid == 0) {
      cu_seqlens[b + 1] = (b + 1) * Pp;
      if (b == 0) cu_seqlens[0] = 0;
    }
  }

  // streamed row gather: 2 rows in flight per iteration
  const int nchunk = H >> 3;
  const bf16* in_b = hid_in + (int64_t)b * P * H;
  bf16* out_b = hid_out + (int64_t)b * Pp * H;
  for (int j = blockIdx.x; j < Pp; j += 2 * gridDim.x) {
    const int j2 = j + gridDim.x;
    const int sj = j < s0 ? j : (j < s0 + k ? s0 + kidx[j - s0] : j + (n_img - k));
    const bool has2 = j2 < Pp;
    const int sj2 = !has2 ? sj : (j2 < s0 ? j2 : (j2 < s0 + k ? s0 + kidx[j2 - s0] : j2 + (n_img - k)));
    const bf16* a = in_b + (int64_t)sj * H;
    const bf16* a2 = in_b + (int64_t)sj2 * H;
    bf16* o = out_b + (int64_t)j * H;
    bf16* o2 = out_b + (int64_t)j2 * H;
    for (int c = tid; c < nchunk; c += 2 * SC_THREADS) {
      const int c2 = c + SC_THREADS;
      int4 v0 = ld_stream(a + c * 8), v1, v2, v3;
      if (c2 < nchunk) v1 = ld_stream(a + c2 * 8);
      if (has2) {
        v2 = ld_stream(a2 + c * 8);
        if (c2 < nchunk) v3 = ld_stream(a2 + c2 * 8);
      }
      st_stream(o + c * 8, v0);
      if (c2 < nchunk) st_stream(o + c2 * 8, v1);
      if (has2) {
        st_stream(o2 + c * 8, v2);
        if (c2 < nchunk) st_stream(o2 + c2 * 8, v3);
      }
    }
  }
}

extern "C" int dl_select_compact(const void* logits, const int64_t* forced_keep, const void* hidden_in, void* hidden_out,
                                 int64_t* keep_index, int32_t* pos_out, int32_t* cu_seqlens, int B, int P, int H,
                                 int img_start, int n_img, int k, int ctas_per_sample, cudaStream_t stream) {
  DL_REQUIRE(H % 8 == 0 && n_img <= SC_MAX_IMG && k >= 0 && k <= n_img && img_start >= 0 && img_start + n_img <= P,
             "dl_select_compact");
  if (B == 0) return 0;
  const int Pp = P - n_img + k;
  int R = ctas_per_sample;
  if (R <= 0) R = (148 * 8 + B - 1) / B;  // ~8 resident CTAs per SM worth of rows in flight
  if (R > Pp) R = Pp;
  if (R < 1) R = 1;
  select_compact_kernel<<<dim3(R, B), SC_THREADS, 0, stream>>>((const bf16*)logits, forced_keep, (const bf16*)hidden_in,
                                                               (bf16*)hidden_out, keep_index, pos_out, cu_seqlens, P, H,
                                                               img_start, n_img, k);
  DL_CHECK_LAUNCH("dl_select_compact");
}
