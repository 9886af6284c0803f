// Multinomial next-token sampling on the device: the `do_sample=True` branch of HF 4.37.2
// GenerationMixin.sample as the reference reaches it from llava/dynamic_eval/model_vqa_loader.py:162-175
// (do_sample = temperature > 0, temperature, top_p, num_beams = 1; top_k comes from GenerationConfig, default 50):
//     scores = logits.float() / temperature                                  (TemperatureLogitsWarper)
//     scores[scores < kth_largest(scores, top_k)] = -inf                     (TopKLogitsWarper; ties at the k-th value are kept)
//     ascending cumsum of softmax(scores) <= 1 - top_p  ->  -inf, top-1 kept (TopPLogitsWarper, min_tokens_to_keep = 1)
//     next = multinomial(softmax(scores), 1);  finished rows emit pad;  finished |= (next == eos)
// HF 4.37.2 is not vendored in the reference tree: the warpers are restated here and pinned by tests against a torch
// restatement (tests/test_kernels_gpu.py::test_sample_next_*); the random STREAM (Philox4x32-10 keyed by a per-call seed, counter =
// (step, row)) is this library's own — torch.multinomial's stream is an implementation detail no caller can depend on.
//
// One CTA of 1024 threads per row; the row (V <= 32 * 1024 fp32 scores) lives in registers.  top_k > 0: a 4-pass radix select on the
// order-preserving 32-bit keys finds the k-th largest score, the <= 1024 candidates are compacted into shared memory, sorted
// (bitonic, score descending / index ascending), softmaxed, top-p filtered with a suffix scan and sampled by inverse CDF.
// top_k == 0 (top_p must be 1): plain temperature sampling over the whole row with a block-wide prefix scan.
// Everything that changes per call sits in device memory (seed, step counter): the launch is CUDA-graph replayable.
#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define SP_THREADS 1024
#define SP_PER 32  // scores per thread: V <= 32768
#define SP_CAND 1024

DL_DEVINL uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
DL_DEVINL uint32_t sp_key(float f) {  // order-preserving: larger float <-> larger key (-inf smallest)
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(SP_THREADS) sample_kernel(const bf16* __restrict__ logits, int V, float inv_temperature, int top_k,
                                                            float top_p, const int64_t* __restrict__ seed_ptr,
                                                            int64_t* __restrict__ next_ids, int64_t* __restrict__ tokens_out,
                                                            int out_stride, const int32_t* __restrict__ step_counter,
                                                            int32_t* __restrict__ finished, int eos, int pad,
                                                            float* __restrict__ probs_out) {
  __shared__ int hist[256];
  __shared__ int sel[2];
  __shared__ float cval[SP_CAND];
  __shared__ int cidx[SP_CAND];
  __shared__ float red[32];
  __shared__ float wtot[32];
  __shared__ int n_cand_s, token_s, last_kept_s;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bf16* row = logits + (int64_t)b * V;
  float s[SP_PER];
#pragma unroll
  for (int j = 0; j < SP_PER; ++j) {
    const int i = tid + j * SP_THREADS;
    s[j] = i < V ? bf2f(row[i]) * inv_temperature : -INFINITY;
  }
  const int step = step_counter[0];
  const int64_t seed = seed_ptr[0];
  const uint4 rnd = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)b, 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float u = ((rnd.x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
  if (probs_out != nullptr)
    for (int i = tid; i < V; i += SP_THREADS) probs_out[(int64_t)b * V + i] = 0.f;
  if (tid == 0) {
    token_s = -1;
    last_kept_s = 0;
  }

  if (top_k > 0) {
    // ---- k-th largest score: radix select, 4 passes of 8 bits ----
    uint32_t prefix = 0;
    int want = min(top_k, V);
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < SP_PER; ++j) {
        const int i = tid + j * SP_THREADS;
        if (i < V) {
          const uint32_t k = sp_key(s[j]);
          if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1);
        }
      }
      __syncthreads();
      if (tid == 0) {  // 256 bins, top down: a single thread is enough (4 x 256 steps per launch)
        int run = 0, bin = 255;
        for (; bin > 0; --bin) {
          if (run + hist[bin] >= want) break;
          run += hist[bin];
        }
        sel[0] = bin;
        sel[1] = run;
      }
      __syncthreads();
      prefix = (prefix << 8) | (uint32_t)sel[0];
      want -= sel[1];
      __syncthreads();
    }
    const uint32_t thr = prefix;  // key of the k-th largest score; everything >= thr is a candidate (ties kept, as HF does)
    if (tid == 0) n_cand_s = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SP_PER; ++j) {
      const int i = tid + j * SP_THREADS;
      if (i < V && sp_key(s[j]) >= thr) {
        const int slot = atomicAdd(&n_cand_s, 1);
        if (slot < SP_CAND) {
          cval[slot] = s[j];
          cidx[slot] = i;
        }
      }
    }
    __syncthreads();
    const int n_cand = min(n_cand_s, SP_CAND);
    if (tid >= n_cand) {
      cval[tid] = -INFINITY;
      cidx[tid] = 0x7fffffff;
    }
    __syncthreads();
    // ---- bitonic sort of 1024 (score desc, index asc) ----
    for (int k2 = 2; k2 <= SP_CAND; k2 <<= 1) {
      for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
        const int p = tid ^ j2;
        if (p > tid) {
          const float va = cval[tid], vb = cval[p];
          const int ia = cidx[tid], ib = cidx[p];
          const bool a_first = va > vb || (va == vb && ia < ib);  // a belongs before b in the final order
          const bool up = (tid & k2) == 0;
          if (up ? !a_first : a_first) {
            cval[tid] = vb;
            cval[p] = va;
            cidx[tid] = ib;
            cidx[p] = ia;
          }
        }
        __syncthreads();
      }
    }
    // ---- softmax over the candidates, top-p, inverse CDF ----
    const float vmax = cval[0];
    float e = tid < n_cand ? expf(cval[tid] - vmax) : 0.f;
    const float tot = block_sum(e, red);
    float p = e / tot;
    // inclusive SUFFIX sum of p in descending order == HF's ascending cumulative sum
    float inc = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_down_sync(0xffffffffu, inc, o);
      if (lane + o < 32) inc += t;
    }
    if (lane == 0) wtot[wid] = inc;
    __syncthreads();
    float after = 0.f;
    for (int w = wid + 1; w < 32; ++w) after += wtot[w];
    const float suffix = inc + after;
    const bool kept = tid < n_cand && (tid == 0 || !(suffix <= 1.0f - top_p));
    const float pk = kept ? p : 0.f;
    const float tot2 = block_sum(pk, red);  // (block_sum starts with a barrier: wtot is free to be reused after it)
    const float q = pk / tot2;
    if (probs_out != nullptr && kept) probs_out[(int64_t)b * V + cidx[tid]] = q;
    // inclusive PREFIX sum of q, then the first position whose prefix exceeds u
    float pre = q;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    if (lane == 31) wtot[wid] = pre;
    __syncthreads();
    float before = 0.f;
    for (int w = 0; w < wid; ++w) before += wtot[w];
    pre += before;
    const bool hit = kept && pre > u && (pre - q) <= u;
    if (hit) token_s = cidx[tid];
    if (kept) atomicMax(&last_kept_s, tid);
    __syncthreads();
    if (tid == 0 && token_s < 0) token_s = cidx[last_kept_s];  // u beyond the accumulated mass by rounding: the last kept candidate
  } else {
    // ---- plain temperature sampling over the whole row ----
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SP_PER; ++j) mx = fmaxf(mx, s[j]);
    mx = warp_max(mx);
    if (lane == 0) red[wid] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < 32; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    // thread-contiguous ownership for an ordered scan: thread t owns indices [t*per, (t+1)*per)
    const int per = (V + SP_THREADS - 1) / SP_THREADS;
    float loc = 0.f;
    for (int j = 0; j < per; ++j) {
      const int i = tid * per + j;
      if (i < V) loc += expf(bf2f(row[i]) * inv_temperature - mx);
    }
    const float tot = block_sum(loc, red);
    float pre = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    if (lane == 31) wtot[wid] = pre;
    __syncthreads();
    float before = 0.f;
    for (int w = 0; w < wid; ++w) before += wtot[w];
    pre += before;
    const float target = u * tot;
    if (pre > target && (pre - loc) <= target) {  // the sampled token is one of this thread's
      float acc = pre - loc;
      int pick = min(tid * per, V - 1);
      for (int j = 0; j < per; ++j) {
        const int i = tid * per + j;
        if (i < V) {
          acc += expf(bf2f(row[i]) * inv_temperature - mx);
          pick = i;
          if (acc > target) break;
        }
      }
      token_s = pick;
    }
    if (probs_out != nullptr)
      for (int i = tid; i < V; i += SP_THREADS) probs_out[(int64_t)b * V + i] = expf(bf2f(row[i]) * inv_temperature - mx) / tot;
    __syncthreads();
    if (tid == 0 && token_s < 0) token_s = V - 1;  // rounding: u * tot beyond the accumulated mass
  }
  __syncthreads();
  if (tid == 0) {
    int tok = token_s;
    if (eos >= 0) {
      if (finished[b]) tok = pad;
      if (tok == eos) finished[b] = 1;
    }
    next_ids[b] = tok;
    if (tokens_out != nullptr) tokens_out[(int64_t)b * out_stride + step] = tok;
  }
}

// logits bf16 [B, V] (V <= 32768); seed int64[1] and step_counter int32[1] on the device; probs_out fp32 [B, V] or NULL (the
// distribution the token is drawn from, for tests).  top_k == 0 disables the top-k filter and then requires top_p >= 1.
extern "C" int dl_sample_next(const void* logits, int B, int V, float temperature, int top_k, float top_p, const int64_t* seed,
                              int64_t* next_ids, int64_t* tokens_out, int out_stride, const int32_t* step_counter,
                              int32_t* finished, int eos_token_id, int pad_token_id, float* probs_out, cudaStream_t stream) {
  DL_REQUIRE(V >= 1 && V <= SP_THREADS * SP_PER && temperature > 0.f && top_k >= 0 && top_k <= SP_CAND && top_p > 0.f,
             "dl_sample_next");
  DL_REQUIRE(top_k > 0 || top_p >= 1.f, "dl_sample_next: top_p < 1 needs a top_k candidate set (<= 1024)");
  if (B == 0) return 0;
  dl_launch_pdl(sample_kernel, dim3(B), dim3(SP_THREADS), 0, stream, (const bf16*)logits, V, 1.0f / temperature, top_k,
                top_p > 1.f ? 1.f : top_p, seed, next_ids, tokens_out, out_stride, step_counter, finished, eos_token_id, pad_token_id,
                probs_out);
  DL_CHECK_LAUNCH("dl_sample_next");
}
