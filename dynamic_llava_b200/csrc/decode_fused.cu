// Fused consumers of the skinny GEMM's split-K slices for the decode step (q_len == 1).
// Each kernel first sums partial[s][m][n] over s in fixed order and rounds to bf16 — that is the rounding point of the
// reference's bf16 nn.Linear output — and then applies the op that follows the projection in the reference graph, so a
// decoder layer costs 4 GEMM launches + 4 of these + 1 attention launch with no other HBM round trips:
//   dl_dec_reduce_bf16           y = bf16(sum)                                          (lm_head, :2709)
//   dl_dec_reduce_residual_norm  x = bf16(x + bf16(sum)); xn = rmsnorm(x, w)            (:1289/:1295 + :134-139)
//   dl_dec_reduce_swiglu         a = bf16(bf16(silu(gate)) * up)                        (:328)
//   dl_text_decision             TextPredictor tail + `logit0 > logit1`                 (:1374-1387, :2385-2391)
// (line numbers: llava/model/language_model/dynamic_modeling_llama.py)
#include "common.cuh"
#include "../../include/dynllava_b200.h"

DL_DEVINL float sum_slices(const float* __restrict__ p, int64_t slice_stride, int S) {
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += p[(int64_t)s * slice_stride];
  return a;
}
DL_DEVINL float4 sum_slices4(const float* __restrict__ p, int64_t slice_stride, int S) { return sum_slices_f4(p, slice_stride, S); }

__global__ void __launch_bounds__(256) dec_reduce_bf16_kernel(const float* __restrict__ partial, bf16* __restrict__ y, int64_t MN,
                                                              int S) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= MN) return;
  const float4 a = sum_slices4(partial + i, MN, S);
  bf16x2 lo = f2bf2(a.x, a.y), hi = f2bf2(a.z, a.w);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&lo);
  o.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(y + i) = o;
}
extern "C" int dl_dec_reduce_bf16(const float* partial, void* y, int M, int N, int splits, cudaStream_t stream) {
  DL_REQUIRE(N % 4 == 0, "dl_dec_reduce_bf16");
  const int64_t MN = (int64_t)M * N;
  if (MN == 0) return 0;
  dl_launch_pdl(dec_reduce_bf16_kernel, dim3((unsigned)((MN / 4 + 255) / 256)), dim3(256), 0, stream, partial, (bf16*)y, MN, splits);
  DL_CHECK_LAUNCH("dl_dec_reduce_bf16");
}

// one CTA per row m:  x[m] += bf16(sum_s partial[s][m]);  xn[m] = w * bf16(x * rstd)
// 1024 threads; each thread owns up to RN_V float4 columns of the row and issues every split-K slice load up front
// (the slices are L2-resident: the GEMM epilogue has just written them), so the kernel is two L2 round trips + one
// block reduction deep.
#define RN_THREADS 1024
template <int RN_V>
__global__ void __launch_bounds__(RN_THREADS) dec_reduce_residual_norm_kernel(const float* __restrict__ partial, bf16* __restrict__ x,
                                                                              const bf16* __restrict__ w, bf16* __restrict__ xn, int M,
                                                                              int H, int S, float eps) {
  __shared__ float red[32];
  // the norm weight is a constant of the model: fetch it BEFORE the dependency wait, off the critical path
  uint2 wraw[RN_V];
#pragma unroll
  for (int v = 0; v < RN_V; ++v) {
    const int i = (threadIdx.x + v * RN_THREADS) * 4;
    wraw[v] = (w != nullptr && i < H) ? *reinterpret_cast<const uint2*>(w + i) : make_uint2(0u, 0u);
  }
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const int64_t slice = (int64_t)M * H;
  float r[RN_V][4];
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < RN_V; ++v) {
    const int i = (threadIdx.x + v * RN_THREADS) * 4;
    if (i < H) {
      const uint2 raw = *reinterpret_cast<const uint2*>(x + (int64_t)m * H + i);
      const float4 a = sum_slices_f4(partial + (int64_t)m * H + i, slice, S);
      const float2 x01 = bf2f2(*reinterpret_cast<const bf16x2*>(&raw.x));
      const float2 x23 = bf2f2(*reinterpret_cast<const bf16x2*>(&raw.y));
      r[v][0] = rbf(x01.x + rbf(a.x));
      r[v][1] = rbf(x01.y + rbf(a.y));
      r[v][2] = rbf(x23.x + rbf(a.z));
      r[v][3] = rbf(x23.y + rbf(a.w));
      bf16x2 lo = f2bf2(r[v][0], r[v][1]), hi = f2bf2(r[v][2], r[v][3]);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(x + (int64_t)m * H + i) = o;
#pragma unroll
      for (int e = 0; e < 4; ++e) ss += r[v][e] * r[v][e];
    }
  }
  ss = block_sum(ss, red);
  if (w == nullptr) return;
  const float rstd = 1.0f / sqrtf(ss / (float)H + eps);
#pragma unroll
  for (int v = 0; v < RN_V; ++v) {
    const int i = (threadIdx.x + v * RN_THREADS) * 4;
    if (i < H) {
      const uint2 raw = wraw[v];
      const float2 w01 = bf2f2(*reinterpret_cast<const bf16x2*>(&raw.x));
      const float2 w23 = bf2f2(*reinterpret_cast<const bf16x2*>(&raw.y));
      bf16x2 lo = f2bf2(w01.x * rbf(r[v][0] * rstd), w01.y * rbf(r[v][1] * rstd));
      bf16x2 hi = f2bf2(w23.x * rbf(r[v][2] * rstd), w23.y * rbf(r[v][3] * rstd));
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(xn + (int64_t)m * H + i) = o;
    }
  }
}
extern "C" int dl_dec_reduce_residual_norm(const float* partial, void* x, const void* norm_w, void* xn, int M, int H,
                                           int splits, float eps, cudaStream_t stream) {
  DL_REQUIRE(H % 4 == 0 && H <= RN_THREADS * 4 * 2, "dl_dec_reduce_residual_norm");
  if (M == 0) return 0;
  if (H <= RN_THREADS * 4)
    dl_launch_pdl(dec_reduce_residual_norm_kernel<1>, dim3(M), dim3(RN_THREADS), 0, stream, partial, (bf16*)x, (const bf16*)norm_w,
                  (bf16*)xn, M, H, splits, eps);
  else
    dl_launch_pdl(dec_reduce_residual_norm_kernel<2>, dim3(M), dim3(RN_THREADS), 0, stream, partial, (bf16*)x, (const bf16*)norm_w,
                  (bf16*)xn, M, H, splits, eps);
  DL_CHECK_LAUNCH("dl_dec_reduce_residual_norm");
}

// partial rows are [gate(I) | up(I)]
__global__ void __launch_bounds__(256) dec_reduce_swiglu_kernel(const float* __restrict__ partial, bf16* __restrict__ act, int M,
                                                                int I, int S) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (idx >= (int64_t)M * I) return;
  const int m = (int)(idx / I), i = (int)(idx % I);
  const int64_t slice = (int64_t)M * 2 * I;
  const float4 g = sum_slices4(partial + (int64_t)m * 2 * I + i, slice, S);
  const float4 u = sum_slices4(partial + (int64_t)m * 2 * I + I + i, slice, S);
  const float gg[4] = {rbf(g.x), rbf(g.y), rbf(g.z), rbf(g.w)};
  const float uu[4] = {rbf(u.x), rbf(u.y), rbf(u.z), rbf(u.w)};
  float o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) o[e] = rbf(gg[e] / (1.0f + expf(-gg[e]))) * uu[e];
  bf16x2 lo = f2bf2(o[0], o[1]), hi = f2bf2(o[2], o[3]);
  uint2 w;
  w.x = *reinterpret_cast<uint32_t*>(&lo);
  w.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(act + (int64_t)m * I + i) = w;
}
extern "C" int dl_dec_reduce_swiglu(const float* partial, void* act, int M, int I, int splits, cudaStream_t stream) {
  DL_REQUIRE(I % 4 == 0, "dl_dec_reduce_swiglu");
  const int64_t n = (int64_t)M * I / 4;
  if (n == 0) return 0;
  dl_launch_pdl(dec_reduce_swiglu_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, partial, (bf16*)act, M, I, splits);
  DL_CHECK_LAUNCH("dl_dec_reduce_swiglu");
}

// ---------------------------------------------------------------------------------------------
// Text predictor tail (one CTA per sample).  Input: split-K slices of Linear(H -> D) applied to LN(hidden)
// (the big first projection ran on the skinny GEMM).  Then +bias, GELU, Linear(D -> D/2), GELU, Linear(D/2 -> D/4),
// GELU, Linear(D/4 -> 2), each output rounded to bf16 like the reference's bf16 modules; keep[b] = logit0 > logit1.
// If `forced` is given the predictor result is overridden by forced[step_counter[0] * B + b].
// ---------------------------------------------------------------------------------------------
DL_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(1024) text_decision_kernel(const float* __restrict__ partial, int S, int B, int D,
                                                            const bf16* __restrict__ b1, const bf16* __restrict__ w3,
                                                            const bf16* __restrict__ b3, const bf16* __restrict__ w5,
                                                            const bf16* __restrict__ b5, const bf16* __restrict__ w7,
                                                            const bf16* __restrict__ b7, const uint8_t* __restrict__ forced,
                                                            const int32_t* __restrict__ step_counter, int forced_offset,
                                                            uint8_t* __restrict__ keep, bf16* __restrict__ logits_out) {
  extern __shared__ float sm[];
  float* h1 = sm;          // D
  float* h2 = h1 + D;      // D/2
  float* h3 = h2 + D / 2;  // D/4
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (partial != nullptr) {
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      const float a = sum_slices(partial + (int64_t)b * D + i, (int64_t)B * D, S);
      h1[i] = rbf(gelu_erf(rbf(a + bf2f(b1[i]))));
    }
    __syncthreads();
    for (int o = warp; o < D / 2; o += nw) {
      float a = 0.f;
      for (int i = lane * 8; i < D; i += 256) {  // D % 8 == 0: 128-bit weight loads
        float wv[8];
        unpack8(ld8(w3 + (int64_t)o * D + i), wv);
#pragma unroll
        for (int e = 0; e < 8; ++e) a += h1[i + e] * wv[e];
      }
      a = warp_sum(a);
      if (lane == 0) h2[o] = rbf(gelu_erf(rbf(a + bf2f(b3[o]))));
    }
    __syncthreads();
    for (int o = warp; o < D / 4; o += nw) {
      float a = 0.f;
      for (int i = lane; i < D / 2; i += 32) a += h2[i] * bf2f(w5[(int64_t)o * (D / 2) + i]);
      a = warp_sum(a);
      if (lane == 0) h3[o] = rbf(gelu_erf(rbf(a + bf2f(b5[o]))));
    }
    __syncthreads();
    if (warp == 0) {
      float l[2];
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        float a = 0.f;
        for (int i = lane; i < D / 4; i += 32) a += h3[i] * bf2f(w7[(int64_t)o * (D / 4) + i]);
        a = warp_sum(a);
        l[o] = rbf(a + bf2f(b7[o]));
      }
      if (lane == 0) {
        keep[b] = l[0] > l[1];
        if (logits_out != nullptr) {
          logits_out[b * 2] = f2bf(l[0]);
          logits_out[b * 2 + 1] = f2bf(l[1]);
        }
      }
    }
  }
  if (forced != nullptr && threadIdx.x == 0) {  // same thread as the natural write above: program order suffices
    keep[b] = forced[(int64_t)(step_counter[0] + forced_offset) * B + b];
  }
}
extern "C" int dl_text_decision(const float* partial, int splits, int B, int D, const void* b1, const void* w3, const void* b3,
                                const void* w5, const void* b5, const void* w7, const void* b7, const uint8_t* forced,
                                const int32_t* step_counter, int forced_offset, uint8_t* keep, void* logits_out,
                                cudaStream_t stream) {
  DL_REQUIRE(D % 8 == 0 && (partial != nullptr || forced != nullptr), "dl_text_decision");
  if (B == 0) return 0;
  const size_t smem = (size_t)(D + D / 2 + D / 4) * sizeof(float);
  dl_launch_pdl(text_decision_kernel, dim3(B), dim3(1024), smem, stream, partial, splits, B, D, (const bf16*)b1, (const bf16*)w3,
                (const bf16*)b3, (const bf16*)w5, (const bf16*)b5, (const bf16*)w7, (const bf16*)b7, forced, step_counter, forced_offset,
                keep, (bf16*)logits_out);
  DL_CHECK_LAUNCH("dl_text_decision");
}
