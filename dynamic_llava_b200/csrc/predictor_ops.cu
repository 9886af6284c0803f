// K7 helpers for the VisionPredictor (dynamic_modeling_llama.py:1348-1359):
//  * LayerNorm over the image rows of the residual stream WITHOUT first re-stacking them sample by sample
//    (the reference copies + zero-pads at :1832-1862): rows are addressed as x[b, row0 + i] through a batch stride;
//  * the local/global channel mix: z = cat[y[..., :C/2], mean_over_tokens(y[..., C/2:]) broadcast]  (:1353-1357),
//    with the reference's rounding: bf16(sum) then bf16(sum / N)  (policy is all ones at inference, :1779-1785).
#include "common.cuh"
#include "../../include/dynllava_b200.h"

__global__ void __launch_bounds__(512) layernorm_strided_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                const bf16* __restrict__ b, bf16* __restrict__ y,
                                                                int rows_per_batch, int64_t x_batch_stride, int H, float eps) {
  __shared__ float red[32];
  extern __shared__ float xs[];
  const int r = blockIdx.x;
  const int bb = r / rows_per_batch, i = r % rows_per_batch;
  const bf16* xr = x + (int64_t)bb * x_batch_stride + (int64_t)i * H;
  bf16* yr = y + (int64_t)r * H;
  float s = 0.f;
  for (int c = threadIdx.x; c < (H >> 3); c += blockDim.x) {
    float f[8];
    unpack8(ld8(xr + c * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xs[c * 8 + j] = f[j];
      s += f[j];
    }
  }
  const float mean = block_sum(s, red) / (float)H;
  float v = 0.f;
  for (int c = threadIdx.x; c < (H >> 3); c += blockDim.x) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = xs[c * 8 + j] - mean;
      v += d * d;
    }
  }
  const float rstd = 1.0f / sqrtf(block_sum(v, red) / (float)H + eps);
  for (int c = threadIdx.x; c < (H >> 3); c += blockDim.x) {
    float g[8], bv[8], o[8];
    unpack8(ld8(w + c * 8), g);
    unpack8(ld8(b + c * 8), bv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (xs[c * 8 + j] - mean) * rstd * g[j] + bv[j];
    st8(yr + c * 8, pack8(o));
  }
}
extern "C" int dl_layernorm_strided(const void* x, const void* w, const void* b, void* y, int B, int rows_per_batch,
                                    int64_t x_batch_stride, int H, float eps, cudaStream_t stream) {
  DL_REQUIRE(H % 8 == 0 && H * 4 <= 48 * 1024, "dl_layernorm_strided");
  if (B * rows_per_batch == 0) return 0;
  layernorm_strided_kernel<<<B * rows_per_batch, 256, H * sizeof(float), stream>>>(
      (const bf16*)x, (const bf16*)w, (const bf16*)b, (bf16*)y, rows_per_batch, x_batch_stride, H, eps);
  DL_CHECK_LAUNCH("dl_layernorm_strided");
}

// g[b, c] = bf16( bf16(sum_n y[b, n, C/2 + c]) / N )
__global__ void __launch_bounds__(256) token_mean_kernel(const bf16* __restrict__ y, bf16* __restrict__ g, int N, int C) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = C >> 1;
  if (c >= half) return;
  const bf16* p = y + (int64_t)b * N * C + half + c;
  float s = 0.f;
  for (int n = 0; n < N; ++n) s += bf2f(p[(int64_t)n * C]);
  g[(int64_t)b * half + c] = f2bf(rbf(s) / rbf((float)N));
}
__global__ void __launch_bounds__(128) predictor_mix_kernel(const bf16* __restrict__ y, const bf16* __restrict__ g,
                                                            bf16* __restrict__ z, int N, int C) {
  const int n = blockIdx.x, b = blockIdx.y;
  const int half = C >> 1;
  const bf16* yr = y + ((int64_t)b * N + n) * C;
  bf16* zr = z + ((int64_t)b * N + n) * C;
  for (int c = threadIdx.x; c < (C >> 3); c += blockDim.x) {
    const int e = c * 8;
    *reinterpret_cast<int4*>(zr + e) =
        (e < half) ? *reinterpret_cast<const int4*>(yr + e) : *reinterpret_cast<const int4*>(g + (int64_t)b * half + (e - half));
  }
}
extern "C" int dl_predictor_mix(const void* y, void* g_scratch, void* z, int B, int N, int C, cudaStream_t stream) {
  DL_REQUIRE(C % 16 == 0, "dl_predictor_mix");
  if (B * N == 0) return 0;
  token_mean_kernel<<<dim3((C / 2 + 255) / 256, B), 256, 0, stream>>>((const bf16*)y, (bf16*)g_scratch, N, C);
  predictor_mix_kernel<<<dim3(N, B), 128, 0, stream>>>((const bf16*)y, (const bf16*)g_scratch, (bf16*)z, N, C);
  DL_CHECK_LAUNCH("dl_predictor_mix");
}

// The predictors' 2-logit output heads (Linear(D/4 -> 2), DML:1345, :1383): N is far below a tensor-core tile and breaks the 16-byte row
// stride TMA needs, so one warp per row does the N dot products (fp32 accumulate, + bias, one bf16 rounding like the bf16 nn.Linear).
// out[m, n] = bf16(sum_k x[m, k] w[n, k] + b[n]);  N <= 8, K % 8 == 0.
__global__ void __launch_bounds__(256) linear_tiny_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                                                          bf16* __restrict__ out, int M, int N, int K) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  float acc[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    float xv[8];
    unpack8(ld8(x + (int64_t)m * K + k), xv);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if (n < N) {
        float wv[8];
        unpack8(ld8(w + (int64_t)n * K + k), wv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[n] += xv[e] * wv[e];
      }
    }
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    if (n < N) {
      const float s = warp_sum(acc[n]);
      if (lane == 0) out[(int64_t)m * N + n] = f2bf(s + (b != nullptr ? bf2f(b[n]) : 0.f));
    }
  }
}
extern "C" int dl_linear_tiny(const void* x, const void* w, const void* bias, void* out, int M, int N, int K, cudaStream_t stream) {
  DL_REQUIRE(N >= 1 && N <= 8 && K % 8 == 0, "dl_linear_tiny");
  if (M == 0) return 0;
  linear_tiny_kernel<<<(M + 7) / 8, 256, 0, stream>>>((const bf16*)x, (const bf16*)w, (const bf16*)bias, (bf16*)out, M, N, K);
  DL_CHECK_LAUNCH("dl_linear_tiny");
}
