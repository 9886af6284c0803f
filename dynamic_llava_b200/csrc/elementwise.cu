// K4: HBM-bound row kernels — RMSNorm, LayerNorm, SwiGLU, bias+activation, residual add, embedding/splice,
// CLIP patch im2col + token assembly, greedy argmax.  One 128-bit access per thread per step, fp32 math,
// rounding points placed exactly where the reference's eager bf16 graph rounds.
#include "common.cuh"
#include "../../include/dynllava_b200.h"

// ------------------------------------------------------------------------------------------------
// RMSNorm  (dynamic_modeling_llama.py:134-139): y = w * bf16(x * rsqrt(mean(x^2) + eps))
// ------------------------------------------------------------------------------------------------
template <int MAXV>  // 128-bit chunks cached per thread
__global__ void __launch_bounds__(512) rmsnorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                      bf16* __restrict__ y, int H, float eps) {
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = blockIdx.x;
  const bf16* xr = x + row * H;
  bf16* yr = y + row * H;
  const int nchunk = H >> 3;
  bf16x8 cache[MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * blockDim.x;
    if (c < nchunk) {
      cache[i] = ld8(xr + c * 8);
      float f[8];
      unpack8(cache[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = 1.0f / sqrtf(ss / (float)H + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * blockDim.x;
    if (c < nchunk) {
      float f[8], g[8];
      unpack8(cache[i], f);
      unpack8(ld8(w + c * 8), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g[j] * rbf(f[j] * rstd);
      st8(yr + c * 8, pack8(f));
    }
  }
}

// Many rows (prefill: B*P rows): ONE WARP per row, CH 128-bit chunks per lane all in flight at once, warp shuffles only — no block
// barrier and 8 rows per CTA, so 1168 rows are one wave of 146 CTAs instead of two waves of 512-thread CTAs serialised on two
// __syncthreads each (measured 9 us -> see profiles/r2_norm_warp_rows.md).
template <int CH>
__global__ void __launch_bounds__(256, 2) rmsnorm_rows_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                                                           int rows, int H, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const bf16* xr = x + row * H;
  bf16* yr = y + row * H;
  bf16x8 cache[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) cache[i] = ld8(xr + (lane + i * 32) * 8);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    float f[8];
    unpack8(cache[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
  }
  ss = warp_sum(ss);
  const float rstd = 1.0f / sqrtf(ss / (float)H + eps);
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    // keep the row PACKED between the two passes (CH * 4 registers): without this the compiler carries the 8 * CH unpacked floats
    // across the reduction and spills at CH >= 16
    uint32_t* cw = reinterpret_cast<uint32_t*>(&cache[i]);
    asm volatile("" : "+r"(cw[0]), "+r"(cw[1]), "+r"(cw[2]), "+r"(cw[3]));
    float f[8], g[8];
    unpack8(cache[i], f);
    unpack8(ld8(w + (lane + i * 32) * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = g[j] * rbf(f[j] * rstd);
    st8(yr + (lane + i * 32) * 8, pack8(f));
  }
}

extern "C" int dl_rmsnorm(const void* x, const void* w, void* y, int rows, int H, float eps, cudaStream_t stream) {
  DL_REQUIRE(H % 8 == 0 && H <= 512 * 8 * 4, "dl_rmsnorm");
  if (rows == 0) return 0;
  if (rows >= 64 && H % 256 == 0 && H / 256 <= 20) {
    const dim3 grid((rows + 7) / 8), block(256);
#define DL_RMS_ROWS(CH_)                                                                                                          \
  case CH_:                                                                                                                       \
    dl_pdl_site = 4;                                                                                                              \
    dl_launch_pdl(rmsnorm_rows_kernel<CH_>, grid, block, 0, stream, (const bf16*)x, (const bf16*)w, (bf16*)y, rows, H, eps); \
    DL_CHECK_LAUNCH("dl_rmsnorm");
    switch (H / 256) {
      DL_RMS_ROWS(1) DL_RMS_ROWS(2) DL_RMS_ROWS(4) DL_RMS_ROWS(8) DL_RMS_ROWS(16) DL_RMS_ROWS(20)
      default: break;
    }
#undef DL_RMS_ROWS
  }
  int threads = H / 8 >= 512 ? 512 : ((H / 8 + 31) / 32) * 32;
  int per = (H / 8 + threads - 1) / threads;
  if (per <= 1)
    dl_launch_pdl(rmsnorm_kernel<1>, dim3(rows), dim3(threads), 0, stream, (const bf16*)x, (const bf16*)w, (bf16*)y, H, eps);
  else if (per <= 2)
    dl_launch_pdl(rmsnorm_kernel<2>, dim3(rows), dim3(threads), 0, stream, (const bf16*)x, (const bf16*)w, (bf16*)y, H, eps);
  else
    dl_launch_pdl(rmsnorm_kernel<4>, dim3(rows), dim3(threads), 0, stream, (const bf16*)x, (const bf16*)w, (bf16*)y, H, eps);
  DL_CHECK_LAUNCH("dl_rmsnorm");
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (torch.nn.LayerNorm: fp32 statistics, one rounding at the end), eps 1e-5 in predictors and CLIP
// ------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(512) layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                        const bf16* __restrict__ b, bf16* __restrict__ y, int H, float eps) {
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = blockIdx.x;
  const bf16* xr = x + row * H;
  bf16* yr = y + row * H;
  const int nchunk = H >> 3;
  float f[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * blockDim.x;
    if (c < nchunk) {
      unpack8(ld8(xr + c * 8), f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
  const float mean = block_sum(s, red) / (float)H;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * blockDim.x;
    if (c < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = f[i][j] - mean;
        v += d * d;
      }
    }
  }
  const float rstd = 1.0f / sqrtf(block_sum(v, red) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = threadIdx.x + i * blockDim.x;
    if (c < nchunk) {
      float g[8], bb[8], o[8];
      unpack8(ld8(w + c * 8), g);
      unpack8(ld8(b + c * 8), bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * g[j] + bb[j];
      st8(yr + c * 8, pack8(o));
    }
  }
}

template <int CH>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                             const bf16* __restrict__ b, bf16* __restrict__ y, int rows, int H, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const bf16* xr = x + row * H;
  bf16* yr = y + row * H;
  float f[CH][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) unpack8(ld8(xr + (lane + i * 32) * 8), f[i]);
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[i][j];
  const float mean = warp_sum(s) / (float)H;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = f[i][j] - mean;
      v += d * d;
    }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    float g[8], bb[8], o[8];
    unpack8(ld8(w + (lane + i * 32) * 8), g);
    unpack8(ld8(b + (lane + i * 32) * 8), bb);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * g[j] + bb[j];
    st8(yr + (lane + i * 32) * 8, pack8(o));
  }
}

extern "C" int dl_layernorm(const void* x, const void* w, const void* b, void* y, int rows, int H, float eps,
                            cudaStream_t stream) {
  DL_REQUIRE(H % 8 == 0 && H <= 512 * 8 * 4, "dl_layernorm");
  if (rows == 0) return 0;
  if (rows >= 64 && H % 256 == 0 && H / 256 <= 8) {
    const dim3 grid((rows + 7) / 8), block(256);
#define DL_LN_ROWS(CH_)                                                                                                              \
  case CH_:                                                                                                                          \
    dl_pdl_site = 4;                                                                                                              \
    dl_launch_pdl(layernorm_rows_kernel<CH_>, grid, block, 0, stream, (const bf16*)x, (const bf16*)w, (const bf16*)b, (bf16*)y, rows, H, \
                  eps);                                                                                                              \
    DL_CHECK_LAUNCH("dl_layernorm");
    switch (H / 256) {
      DL_LN_ROWS(1) DL_LN_ROWS(2) DL_LN_ROWS(4) DL_LN_ROWS(8)
      default: break;
    }
#undef DL_LN_ROWS
  }
  int threads = H / 8 >= 512 ? 512 : ((H / 8 + 31) / 32) * 32;
  int per = (H / 8 + threads - 1) / threads;
  if (per <= 1)
    dl_launch_pdl(layernorm_kernel<1>, dim3(rows), dim3(threads), 0, stream, (const bf16*)x, (const bf16*)w, (const bf16*)b, (bf16*)y, H, eps);
  else if (per <= 2)
    dl_launch_pdl(layernorm_kernel<2>, dim3(rows), dim3(threads), 0, stream, (const bf16*)x, (const bf16*)w, (const bf16*)b, (bf16*)y, H, eps);
  else
    dl_launch_pdl(layernorm_kernel<4>, dim3(rows), dim3(threads), 0, stream, (const bf16*)x, (const bf16*)w, (const bf16*)b, (bf16*)y, H, eps);
  DL_CHECK_LAUNCH("dl_layernorm");
}

// ------------------------------------------------------------------------------------------------
// SwiGLU (dynamic_modeling_llama.py:328): out = bf16( bf16(silu(gate)) * up ), gate|up packed as [rows, 2I]
// ------------------------------------------------------------------------------------------------
DL_DEVINL float silu_f(float g) { return g / (1.0f + expf(-g)); }

__global__ void __launch_bounds__(256) swiglu_kernel(const bf16* __restrict__ gu, bf16* __restrict__ out, int64_t rows, int I) {
  const int cpr = I >> 3;
  const int64_t total = rows * cpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cpr;
    const int c = (int)(idx - r * cpr);
    const bf16* g = gu + r * 2 * I + c * 8;
    int4 gv = ld_stream(g), uv = ld_stream(g + I);
    float gf[8], uf[8], o[8];
    unpack8(*reinterpret_cast<bf16x8*>(&gv), gf);
    unpack8(*reinterpret_cast<bf16x8*>(&uv), uf);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = rbf(silu_f(gf[j])) * uf[j];
    bf16x8 p = pack8(o);
    st_stream(out + r * I + c * 8, *reinterpret_cast<int4*>(&p));
  }
}

extern "C" int dl_swiglu(const void* gate_up, void* out, int64_t rows, int I, cudaStream_t stream) {
  DL_REQUIRE(I % 8 == 0, "dl_swiglu");
  if (rows == 0) return 0;
  int64_t total = rows * (I / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  swiglu_kernel<<<blocks, 256, 0, stream>>>((const bf16*)gate_up, (bf16*)out, rows, I);
  DL_CHECK_LAUNCH("dl_swiglu");
}

// ------------------------------------------------------------------------------------------------
// bias + activation, in place.  act: 0 none, 1 GELU(erf) (nn.GELU default), 2 quick_gelu (CLIP).
// quick_gelu follows HF's eager graph on bf16 tensors: t = bf16(1.702*x); s = bf16(sigmoid(t)); y = bf16(x*s).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bias_act_kernel(bf16* __restrict__ x, const bf16* __restrict__ bias, int64_t rows, int N, int act) {
  const int cpr = N >> 3;
  const int64_t total = rows * cpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cpr);
    float f[8];
    unpack8(ld8(x + idx * 8), f);
    if (bias != nullptr) {
      float b[8];
      unpack8(ld8(bias + c * 8), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = rbf(f[j] + b[j]);
    }
    if (act == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.5f * f[j] * (1.0f + erff(f[j] * 0.70710678118654752440f));
    } else if (act == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = rbf(1.702f * f[j]);
        float s = rbf(1.0f / (1.0f + expf(-t)));
        f[j] = f[j] * s;
      }
    }
    st8(x + idx * 8, pack8(f));
  }
}

extern "C" int dl_bias_act(void* x, const void* bias, int64_t rows, int N, int act, cudaStream_t stream) {
  DL_REQUIRE(N % 8 == 0 && act >= 0 && act <= 2, "dl_bias_act");
  if (rows == 0) return 0;
  int64_t total = rows * (N / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bias_act_kernel<<<blocks, 256, 0, stream>>>((bf16*)x, (const bf16*)bias, rows, N, act);
  DL_CHECK_LAUNCH("dl_bias_act");
}

// residual: x = bf16(x + y)   (dynamic_modeling_llama.py:1289, :1295; custom_transformer_layer.py:321-322)
__global__ void __launch_bounds__(256) add_kernel(bf16* __restrict__ x, const bf16* __restrict__ y, int64_t nchunk) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < nchunk; idx += (int64_t)gridDim.x * blockDim.x) {
    float a[8], b[8];
    unpack8(ld8(x + idx * 8), a);
    unpack8(ld8(y + idx * 8), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    st8(x + idx * 8, pack8(a));
  }
}
extern "C" int dl_add(void* x, const void* y, int64_t n, cudaStream_t stream) {
  DL_REQUIRE(n % 8 == 0, "dl_add");
  if (n == 0) return 0;
  int64_t nchunk = n / 8;
  int blocks = (int)((nchunk + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  add_kernel<<<blocks, 256, 0, stream>>>((bf16*)x, (const bf16*)y, nchunk);
  DL_CHECK_LAUNCH("dl_add");
}

// ------------------------------------------------------------------------------------------------
// Multimodal splice (dynamic_llava_arch.py:364-405, :509-579): out[b, r] =
//   embed[ids[b, r]]                     r <  s_b                 (system)
//   image_feat[b, r - s_b]               s_b <= r < s_b + n_img   (image)
//   embed[ids[b, r - n_img + 1]]         r >= s_b + n_img         (instruct; the <image> id is skipped)
//   0                                    r >= T_b - 1 + n_img     (right padding of shorter samples)
// ids are int64 [B, T] padded with anything beyond seq_len[b]; img_pos[b] < 0 means "no image in this sample".
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_splice_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ seq_len,
                                                           const int32_t* __restrict__ img_pos, const bf16* __restrict__ table,
                                                           const bf16* __restrict__ feat, bf16* __restrict__ out, int T, int P,
                                                           int n_img, int H, int V) {
  const int r = blockIdx.x, b = blockIdx.y;
  const int s = img_pos[b];
  const int len = seq_len[b];
  const bf16* src = nullptr;
  if (s < 0) {
    if (r < len) src = table + (int64_t)ids[(int64_t)b * T + r] * H;
  } else if (r < s) {
    src = table + (int64_t)ids[(int64_t)b * T + r] * H;
  } else if (r < s + n_img) {
    src = feat + ((int64_t)b * n_img + (r - s)) * H;
  } else if (r < len - 1 + n_img) {
    int64_t id = ids[(int64_t)b * T + (r - n_img + 1)];
    src = table + id * H;
  }
  bf16* dst = out + ((int64_t)b * P + r) * H;
  const int nchunk = H >> 3;
  for (int c = threadIdx.x; c < nchunk; c += blockDim.x) {
    int4 v = make_int4(0, 0, 0, 0);
    if (src != nullptr) v = *reinterpret_cast<const int4*>(src + c * 8);
    *reinterpret_cast<int4*>(dst + c * 8) = v;
  }
}
extern "C" int dl_embed_splice(const int64_t* ids, const int32_t* seq_len, const int32_t* img_pos, const void* table,
                               const void* image_feat, void* out, int B, int T, int P, int n_img, int H, int V,
                               cudaStream_t stream) {
  DL_REQUIRE(H % 8 == 0, "dl_embed_splice");
  if (B == 0 || P == 0) return 0;
  embed_splice_kernel<<<dim3(P, B), 256, 0, stream>>>(ids, seq_len, img_pos, (const bf16*)table, (const bf16*)image_feat,
                                                      (bf16*)out, T, P, n_img, H, V);
  DL_CHECK_LAUNCH("dl_embed_splice");
}

// plain embedding gather for decode steps: out[b] = table[ids[b]]  (ids int64 on device)
__global__ void __launch_bounds__(256) embed_rows_kernel(const int64_t* __restrict__ ids, const bf16* __restrict__ table,
                                                         bf16* __restrict__ out, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t id = ids[blockIdx.x];
  const bf16* src = table + id * H;
  bf16* dst = out + (int64_t)blockIdx.x * H;
  for (int c = threadIdx.x; c < (H >> 3); c += blockDim.x)
    *reinterpret_cast<int4*>(dst + c * 8) = *reinterpret_cast<const int4*>(src + c * 8);
}
extern "C" int dl_embed_rows(const int64_t* ids, const void* table, void* out, int n, int H, cudaStream_t stream) {
  DL_REQUIRE(H % 8 == 0, "dl_embed_rows");
  if (n == 0) return 0;
  dl_launch_pdl(embed_rows_kernel, dim3(n), dim3(256), 0, stream, ids, (const bf16*)table, (bf16*)out, H);
  DL_CHECK_LAUNCH("dl_embed_rows");
}

// ------------------------------------------------------------------------------------------------
// CLIP patch embedding front end (HF CLIPVisionEmbeddings).
// im2col: pixels [B, C, S, S] -> patches [B * G*G, Kpad] with column order (c, ky, kx) == Conv2d weight flattening;
// Kpad >= C*ps*ps is zero padded so the following GEMM has 16-byte aligned rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_kernel(const bf16* __restrict__ pix, bf16* __restrict__ out, int C, int S, int ps,
                                                     int G, int K, int Kpad) {
  const int patch = blockIdx.x, b = blockIdx.y;
  const int py = patch / G, px = patch % G;
  bf16* dst = out + ((int64_t)b * G * G + patch) * Kpad;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    bf16 v = f2bf(0.f);
    if (k < K) {
      int c = k / (ps * ps), rem = k % (ps * ps), ky = rem / ps, kx = rem % ps;
      v = pix[(((int64_t)b * C + c) * S + (py * ps + ky)) * S + (px * ps + kx)];
    }
    dst[k] = v;
  }
}
extern "C" int dl_im2col_patches(const void* pixels, void* out, int B, int C, int S, int ps, int Kpad, cudaStream_t stream) {
  DL_REQUIRE(S % ps == 0 && Kpad >= C * ps * ps, "dl_im2col_patches");
  if (B == 0) return 0;
  const int G = S / ps;
  im2col_kernel<<<dim3(G * G, B), 256, 0, stream>>>((const bf16*)pixels, (bf16*)out, C, S, ps, G, C * ps * ps, Kpad);
  DL_CHECK_LAUNCH("dl_im2col_patches");
}

// tokens[b, 0] = bf16(cls + pos[0]);  tokens[b, 1+p] = bf16(patch_embed[b, p] + pos[1+p])
__global__ void __launch_bounds__(128) clip_assemble_kernel(const bf16* __restrict__ pe, const bf16* __restrict__ cls,
                                                            const bf16* __restrict__ pos, bf16* __restrict__ out, int NP, int C) {
  const int t = blockIdx.x, b = blockIdx.y;
  const bf16* src = (t == 0) ? cls : pe + ((int64_t)b * NP + (t - 1)) * C;
  bf16* dst = out + ((int64_t)b * (NP + 1) + t) * C;
  for (int c = threadIdx.x; c < (C >> 3); c += blockDim.x) {
    float a[8], p[8];
    unpack8(ld8(src + c * 8), a);
    unpack8(ld8(pos + (int64_t)t * C + c * 8), p);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += p[j];
    st8(dst + c * 8, pack8(a));
  }
}
extern "C" int dl_clip_assemble(const void* patch_embed, const void* cls, const void* pos, void* out, int B, int NP, int C,
                                cudaStream_t stream) {
  DL_REQUIRE(C % 8 == 0, "dl_clip_assemble");
  if (B == 0) return 0;
  clip_assemble_kernel<<<dim3(NP + 1, B), 128, 0, stream>>>((const bf16*)patch_embed, (const bf16*)cls, (const bf16*)pos,
                                                            (bf16*)out, NP, C);
  DL_CHECK_LAUNCH("dl_clip_assemble");
}

// drop CLS: out[b, p] = x[b, 1 + p]  (clip_encoder.py:43-51 feature_select "patch")
__global__ void __launch_bounds__(128) drop_cls_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int NP, int C) {
  const int p = blockIdx.x, b = blockIdx.y;
  const bf16* src = x + ((int64_t)b * (NP + 1) + p + 1) * C;
  bf16* dst = out + ((int64_t)b * NP + p) * C;
  for (int c = threadIdx.x; c < (C >> 3); c += blockDim.x)
    *reinterpret_cast<int4*>(dst + c * 8) = *reinterpret_cast<const int4*>(src + c * 8);
}
extern "C" int dl_drop_cls(const void* x, void* out, int B, int NP, int C, cudaStream_t stream) {
  DL_REQUIRE(C % 8 == 0, "dl_drop_cls");
  if (B == 0) return 0;
  drop_cls_kernel<<<dim3(NP, B), 128, 0, stream>>>((const bf16*)x, (bf16*)out, NP, C);
  DL_CHECK_LAUNCH("dl_drop_cls");
}

// ------------------------------------------------------------------------------------------------
// Greedy next token (HF 4.37.2 greedy search): argmax over fp32(bf16 logits), first index on ties;
// rows already finished emit pad; finished |= (tok == eos).  Writes tokens_out[b * out_stride + step_slot] and
// next_ids[b].  `step_counter` (device int32) supplies the slot so that the launch is CUDA-graph replayable.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) greedy_kernel(const bf16* __restrict__ logits, int V, int64_t* __restrict__ next_ids,
                                                      int64_t* __restrict__ tokens_out, int out_stride,
                                                      const int32_t* __restrict__ step_counter, int32_t* __restrict__ finished,
                                                      int eos, int pad) {
  __shared__ float sval[32];
  __shared__ int sidx[32];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const bf16* row = logits + (int64_t)b * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float v = bf2f(row[i]);
    if (v > best || (v == best && i < bi)) {
      best = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    sval[wid] = best;
    sidx[wid] = bi;
  }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    best = lane < nw ? sval[lane] : -INFINITY;
    bi = lane < nw ? sidx[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      int tok = bi;
      if (eos >= 0) {
        if (finished[b]) tok = pad;
        if (tok == eos) finished[b] = 1;
      }
      next_ids[b] = tok;
      if (tokens_out != nullptr) tokens_out[(int64_t)b * out_stride + step_counter[0]] = tok;
    }
  }
}
extern "C" int dl_greedy_next(const void* logits, int B, int V, int64_t* next_ids, int64_t* tokens_out, int out_stride,
                              const int32_t* step_counter, int32_t* finished, int eos_token_id, int pad_token_id,
                              cudaStream_t stream) {
  if (B == 0) return 0;
  dl_launch_pdl(greedy_kernel, dim3(B), dim3(1024), 0, stream, (const bf16*)logits, V, next_ids, tokens_out, out_stride, step_counter,
                finished, eos_token_id, pad_token_id);
  DL_CHECK_LAUNCH("dl_greedy_next");
}
