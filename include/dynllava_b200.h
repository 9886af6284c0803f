/* dynllava_b200.h — C ABI of libdynllava_b200.so (hand-written sm_100a kernels for Dynamic-LLaVA's sparsified
 * prefill/decode path).
 *
 * The reference (Osilly/dynamic_llava @ e1469bc) has no FFI/plugin boundary — it is pure PyTorch — so these entry points
 * are what a PyTorch-side binding of that path would call (ctypes stub in INTEGRATION.md); each one names the reference
 * code it replaces (paths relative to llava/model/; DML = language_model/dynamic_modeling_llama.py,
 * CU = language_model/cache_utils.py, ARCH = dynamic_llava_arch.py, CTL = language_model/custom_transformer_layer.py).
 *
 * Two libraries are built from the same sources and export the same symbols: libdynllava_b200.so computes in
 * bfloat16, libdynllava_b200_f16.so in float16 (the dtype the reference's eval drivers load,
 * dynamic_llava_builder.py:62); dl_element_type() says which one is loaded.  "bf16" below means "the library's
 * 16-bit element type".
 *
 * Conventions: every pointer is a DEVICE pointer owned by the caller (kernels never allocate or free); tensors are
 * bf16 unless stated; all calls are asynchronous on `stream`; return 0 on success, a cudaError_t value (>0) or a
 * negative argument error otherwise; dl_last_error_string() describes the last failure on the calling thread.
 * Nothing here synchronises with the host, so any sequence of calls can be captured into a CUDA graph.
 */
#ifndef DYNLLAVA_B200_H
#define DYNLLAVA_B200_H

#include <cuda_runtime_api.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DL_B200_ABI_VERSION 7  /* 7: dl_element_type (bf16 and fp16 builds); 6: superseded attention kernels (warp-MMA dl_attn_varlen, non-persistent dl_attn_prefill_tc2 / dl_attn_packed_tc) removed; 5: dl_linear / dl_linear_swiglu (2-CTA tcgen05 GEMM) replace dl_gemm_bias_act; 2: persistent attention (tc4) added, v1 tcgen05 attention removed; 3: instruct-predictor entry points;
                                 * 4: no-KV-cache output sparsification (dl_output_rank / dl_output_keep_index) */

int dl_version(void);
int dl_element_type(void); /* 0: bfloat16 (libdynllava_b200.so), 1: float16 (libdynllava_b200_f16.so) */
const char* dl_last_error_string(void);
int dl_device_is_sm100(void);
int dl_set_error(int code, const char* where); /* internal */

/* ---- K4: row kernels ------------------------------------------------------------------------------------------ */
/* LlamaRMSNorm.forward, DML:134-139.  x,y [rows,H]; w [H]. */
int dl_rmsnorm(const void* x, const void* w, void* y, int rows, int H, float eps, cudaStream_t stream);
/* nn.LayerNorm in the predictors (DML:1325, :1375; CTL:293,308) and HF CLIP (pre_layrnorm, layer_norm1/2). */
int dl_layernorm(const void* x, const void* w, const void* b, void* y, int rows, int H, float eps, cudaStream_t stream);
/* LlamaMLP: act_fn(gate) * up, DML:328.  gate_up [rows, 2I] = [gate | up]; out [rows, I]. */
int dl_swiglu(const void* gate_up, void* out, int64_t rows, int I, cudaStream_t stream);
/* optional bias + activation in place.  act 0: none, 1: nn.GELU (erf) (DML:1327,1341,1343; CTL:94; projector GELU),
 * 2: quick_gelu (HF CLIP MLP). */
int dl_bias_act(void* x, const void* bias, int64_t rows, int N, int act, cudaStream_t stream);
/* residual x += y (DML:1289,1295; CTL:321-322; HF CLIPEncoderLayer). */
int dl_add(void* x, const void* y, int64_t n, cudaStream_t stream);

/* ---- multimodal splice: prepare_inputs_labels_for_multimodal, ARCH:364-405, :509-579 -------------------------- */
/* ids int64 [B,T]; seq_len int32 [B] (#ids per sample); img_pos int32 [B] (index of the -200 id, <0: none);
 * table [V,H]; image_feat [B,n_img,H]; out [B,P,H], P = max_b(seq_len-1+n_img); rows past a sample's end are zeroed. */
int dl_embed_splice(const int64_t* ids, const int32_t* seq_len, const int32_t* img_pos, const void* table,
                    const void* image_feat, void* out, int B, int T, int P, int n_img, int H, int V, cudaStream_t stream);
/* embed_tokens for decode steps, DML:1725.  ids int64 [n] on device. */
int dl_embed_rows(const int64_t* ids, const void* table, void* out, int n, int H, cudaStream_t stream);

/* ---- CLIP front end: HF CLIPVisionEmbeddings (transformers, unvendored) + feature_select ---- */
int dl_im2col_patches(const void* pixels, void* out, int B, int C, int S, int ps, int Kpad, cudaStream_t stream);
int dl_clip_assemble(const void* patch_embed, const void* cls, const void* pos, void* out, int B, int NP, int C,
                     cudaStream_t stream);
/* multimodal_encoder/clip_encoder.py:43-51 (select_feature == "patch") */
int dl_drop_cls(const void* x, void* out, int B, int NP, int C, cudaStream_t stream);

/* Dense bf16 GEMM for every nn.Linear at q_len > 1 — prefill q/k/v/o/gate/up/down (DML:377-392, :1011-1013, :1127, :294-330), lm_head
 * over all positions (DML:2709), CLIP encoder linears + Conv2d patch embedding as an im2col GEMM (HF CLIPVisionEmbeddings /
 * CLIPEncoderLayer), the mlp2x_gelu projector (multimodal_projector/builder.py:172-179), the predictor linears (DML:1324-1346,
 * CTL:86-180): persistent 2-CTA (cta_group::2) tcgen05 tiles, TMA operand staging, double-buffered TMEM accumulators (gemm2.cu).
 *   out[M,N] = epi(X[M,K] . W[N,K]^T + bias[N]);  act 0: none, 1: nn.GELU (erf) on the bf16-rounded sum, 2: HF quick_gelu;
 *   residual != NULL (act must be 0): out = bf16(residual + bf16(X.W^T + bias)), residual [M,N] may alias out.
 * bias / residual may be NULL.  K % 64 == 0, N % 8 == 0.  x, w, out contiguous row-major.
 * workspace (or NULL): dl_linear_workspace_bytes() bytes of device memory, ZEROED once by the caller and then owned by the calls of
 * one stream: enables stream-K work distribution (equal (tile, k-block) ranges per CTA pair; partial accumulators of split tiles are
 * parked there and added in fixed order: deterministic); every launch leaves it zeroed again.  NULL: whole tiles per pair.
 * Launched with programmatic stream serialization: consecutive launches of the stream overlap their prologue with the predecessor's
 * tail and may share the workspace (each warp that touches it waits for the predecessor grid first); two STREAMS must not share one. */
int64_t dl_linear_workspace_bytes(void);
int dl_linear(const void* x, const void* w, const void* bias, const void* residual, void* out, int M, int N, int K, int act,
              void* workspace, cudaStream_t stream);
/* LlamaMLP's gate|up projection with act_fn(gate) * up fused into the epilogue (DML:328): w_gate_up [2I,K] = [gate rows | up rows],
 * act_out [M,I] = bf16(bf16(silu(bf16 g)) * bf16 u).  I % 128 == 0 recommended (7B: 11008 = 86*128, 13B: 13824 = 108*128). */
int dl_linear_swiglu(const void* x, const void* w_gate_up, void* act_out, int M, int I, int K, void* workspace, cudaStream_t stream);
/* The predictors' 2-logit heads (DML:1345, :1383): out[M,N] = bf16(X[M,K] . W[N,K]^T + bias), N <= 8, K % 8 == 0 (one warp per row). */
int dl_linear_tiny(const void* x, const void* w, const void* bias, void* out, int M, int N, int K, cudaStream_t stream);

/* ---- K1: predictor score + top-k + gather/compact + index map + CSR header, DML:1867, :1897-1994 ------------- */
/* logits [B,n_img,2]; forced_keep int64 [B,k] or NULL; hidden_in [B,P,H]; hidden_out [B,P',H], P' = P-n_img+k;
 * keep_index int64 [B,k] ascending; pos_out int32 [B,P'] original positions; cu_seqlens int32 [B+1].
 * ctas_per_sample <= 0 picks a grid that fills 148 SMs. */
int dl_select_compact(const void* logits, const int64_t* forced_keep, const void* hidden_in, void* hidden_out,
                      int64_t* keep_index, int32_t* pos_out, int32_t* cu_seqlens, int B, int P, int H, int img_start,
                      int n_img, int k, int ctas_per_sample, cudaStream_t stream);

/* ---- K4+K5: RoPE + (conditional) KV append, DML:253-285, :1043-1083; CU:109-268 ------------------------------- */
/* qkv [n_tok, 3H] = [q|k|v], q rotated in place; pos int32 [n_tok]; cu_seqlens int32 [B+1]; kv_len int32 [B] = slot
 * base of this layer; keep uint8 [n_tok] or NULL; cos/sin [max_pos, hd]; arenas [B, heads, cap, hd];
 * k_side/v_side [n_tok, heads, hd] or NULL. */
int dl_rope_append(void* qkv, const int32_t* pos, const int32_t* cu_seqlens, const int32_t* kv_len, const uint8_t* keep,
                   const void* cos_table, const void* sin_table, void* k_arena, void* v_arena, void* k_side, void* v_side,
                   int n_tok, int B, int heads, int head_dim, int cap, cudaStream_t stream);
/* same, reading the skinny GEMM's split-K slices partial[splits][B][3H] (decode step, one token per sample). */
int dl_dec_reduce_rope_append(const float* partial, int splits, void* q_out, const int32_t* pos, const int32_t* kv_len,
                              const uint8_t* keep, const void* cos_table, const void* sin_table, void* k_arena,
                              void* v_arena, void* k_side, void* v_side, int B, int heads, int head_dim, int cap,
                              cudaStream_t stream);
/* DynamicCachePlus.true_cache_length bookkeeping (CU:146-148, :162, :197-199, :250), on the device.
 * kv_len int32 [L,B]; layers < sparse_layer += n_lo; layers >= sparse_layer += n_hi where keep (NULL = all);
 * pos[b] += n_lo; step_counter[0] += 1 (either may be NULL). */
int dl_kv_advance(int32_t* kv_len, int32_t* pos, int32_t* step_counter, const uint8_t* keep, int L, int B, int sparse_layer,
                  int n_lo, int n_hi, cudaStream_t stream);

/* ---- K7: VisionPredictor helpers, DML:1348-1359 (+ the re-stacking at :1832-1862 made unnecessary) ------------- */
/* LayerNorm of rows x[b*x_batch_stride + i*H], i < rows_per_batch, into contiguous y [B*rows_per_batch, H]. */
int dl_layernorm_strided(const void* x, const void* w, const void* b, void* y, int B, int rows_per_batch,
                         int64_t x_batch_stride, int H, float eps, cudaStream_t stream);
/* z = cat[y[..., :C/2], token-mean(y[..., C/2:]) broadcast]; g_scratch bf16 [B, C/2]. */
int dl_predictor_mix(const void* y, void* g_scratch, void* z, int B, int N, int C, cudaStream_t stream);

/* ---- K3: ragged split-KV decode attention, DML:1061-1122 with q_len == 1 -------------------------------------- */
int64_t dl_attn_decode_workspace_bytes(int B, int heads, int n_split);
/* q [B, heads*hd] (row stride q_row_stride elems); k_side/v_side [B, heads, hd] = the new token (NULL: cache only);
 * kv_len int32 [B] cached keys (excl. the new token); out [B, heads*hd]; quirk_pad != 0 reproduces the reference's
 * unmasked zero padding for B > 1 (CU:165-241). */
int dl_attn_decode(const void* q, int q_row_stride, const void* k_arena, const void* v_arena, const void* k_side,
                   const void* v_side, const int32_t* kv_len, void* out, void* workspace, int B, int heads, int head_dim,
                   int cap, int n_split, float scale, int quirk_pad, cudaStream_t stream);

/* K3 + K4 + K5 fused for the decode step: reads the QKV projection as split-K slices partial[n_slices][B][3H], does the
 * reduce + bf16 rounding + RoPE + conditional append per (sample, head) inside the attention CTA (DML:1011-1083). */
int dl_attn_decode_fused(const float* qkv_partial, int n_slices, const int32_t* pos, const uint8_t* keep, const void* cos_table,
                         const void* sin_table, void* k_arena, void* v_arena, const int32_t* kv_len, void* out, void* workspace,
                         int B, int heads, int head_dim, int cap, int n_split, float scale, int quirk_pad, const void* l2_prefetch,
                         int64_t l2_prefetch_bytes, int len_floor, cudaStream_t stream);
/* l2_prefetch (or NULL): device range (the next GEMM's weight) the CTAs ask the TMA unit to pull into L2 while they run.
 * len_floor: host-known lower bound of every kv_len[b] (the prefill length; 0 = unknown): K/V tiles below it are fetched
 * before the kernel waits on its predecessor. */

/* ---- K2 / K6 / K7: prefill attention on the persistent TMA + tcgen05 pipeline (attn_prefill_tc4.cu), DML:1114-1122; CTL:164-169; HF
 * CLIPAttention.
 * dl_attn_prefill_tc4: causal, head_dim 128.  q [n_tok, heads*128] rows at q_tok_stride elements (the fused q|k|v projection);
 * K / V in the arena [B, heads, cap, 128] (this pass's keys already appended); per-sample query ranges from cu_q int32 [B+1];
 * kv_past int32 [B] or NULL: keys cached before this pass (query i attends keys j <= i + kv_past[b]); out [n_tok, heads*128] rows at
 * o_tok_stride elements.
 * dl_attn_packed_tc4: non-causal, head_dim 64, q|k|v packed in one fused projection buffer qkv [n_tok, 3*heads*64]; out [n_tok, heads*64]
 * (CLIP ViT-L/14, the VisionPredictor blocks CTL:153-180).
 * 64-key half tiles with double-buffered S in TMEM, one MMA issuer per softmax group, grid = #SMs, heaviest-first tile list, barriers
 * indexed by running counters so that the loads and S(0)/S(1) of the next tile hide under the current tile's tail, O written with TMA
 * stores.  flags bit 0: one exp2 in four on the FMA pipe (degree-3 polynomial, rel. error 7.5e-5) instead of the SFU. */
int dl_attn_prefill_tc4(const void* q, int64_t q_tok_stride, const void* k_arena, const void* v_arena, void* out,
                        int64_t o_tok_stride, const int32_t* cu_q, const int32_t* kv_past, int n_tok, int B, int heads,
                        int cap, int max_q_len, float scale, int flags, cudaStream_t stream);
int dl_attn_packed_tc4(const void* qkv, void* out, const int32_t* cu_q, int n_tok, int B, int heads, int head_dim,
                       int max_q_len, float scale, int flags, cudaStream_t stream);

/* ---- decode-step weight streaming: every nn.Linear at q_len == 1 (DML:1011-1013, :1127, :328, :2709) ----------- */
/* partial fp32 [splits, M, N] = per-split X[M,K] . W[N,K]^T; M <= 64, K % 64 == 0.  TMA + tcgen05.
 * l2_prefetch_mb: bytes of W (total over the grid) each launch asks the TMA unit to pull into L2 ahead of its smem ring
 * while it still waits (PDL) for its predecessor; <= 0 = off (default: measured slower on B200, profiles/r1_l2_prefetch_sweep.md). */
int dl_skinny_gemm_splits(int N, int K, int target_ctas);
int dl_skinny_gemm(const void* x, const void* w, float* partial, int M, int N, int K, int splits, int l2_prefetch_mb,
                   cudaStream_t stream);
/* gate|up projection with the SwiGLU fused into the TMEM epilogue (DML:328 at q_len == 1).  w_gate_up [2I,K] = [gate rows | up
 * rows] (the same fused weight the prefill uses): each CTA tile is assembled from 64 gate rows + the 64 matching up rows by
 * two TMA boxes; act_out bf16 [M, I]. */
int dl_skinny_gemm_swiglu(const void* x, const void* w_gate_up, void* act_out, int M, int I, int K, cudaStream_t stream);
int dl_make_tmap_2d(void* out_map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                    uint32_t box_rows, uint32_t box_cols);
int dl_dec_reduce_bf16(const float* partial, void* y, int M, int N, int splits, cudaStream_t stream);
int dl_dec_reduce_residual_norm(const float* partial, void* x, const void* norm_w, void* xn, int M, int H, int splits,
                                float eps, cudaStream_t stream);
int dl_dec_reduce_swiglu(const float* partial, void* act, int M, int I, int splits, cudaStream_t stream);
/* TextPredictor tail + decode decision, DML:1374-1387, :2385-2391.  partial = split-K slices of Linear(H->D)(LN(x));
 * forced uint8 [n_steps,B] or NULL overrides with forced[(step_counter[0]+forced_offset)*B + b]. */
int dl_text_decision(const float* partial, int splits, int B, int D, const void* b1, const void* w3, const void* b3,
                     const void* w5, const void* b5, const void* w7, const void* b7, const uint8_t* forced,
                     const int32_t* step_counter, int forced_offset, uint8_t* keep, void* logits_out, cudaStream_t stream);

/* ---- greedy next token: transformers-4.37.2 greedy search step ------------------------------ */
/* logits bf16 [B,V]; next_ids int64 [B]; tokens_out int64 [B,out_stride] slot step_counter[0] (NULL: skip);
 * finished int32 [B]; eos_token_id < 0 disables EOS tracking. */
int dl_greedy_next(const void* logits, int B, int V, int64_t* next_ids, int64_t* tokens_out, int out_stride,
                   const int32_t* step_counter, int32_t* finished, int eos_token_id, int pad_token_id, cudaStream_t stream);

/* HF 4.37.2 GenerationMixin.sample step (transformers, unvendored; reached from llava/dynamic_eval/model_vqa_loader.py:162-175 with
 * do_sample = temperature > 0): scores = logits / temperature -> top-k (ties at the k-th value kept) -> top-p (ascending cumulative
 * softmax <= 1 - top_p removed, top-1 always kept) -> multinomial draw (Philox4x32-10, key = seed[0], counter = (step, row)); finished
 * rows emit pad, finished |= (token == eos).  logits bf16 [B,V], V <= 32768; seed int64[1], step_counter int32[1] on the device;
 * probs_out fp32 [B,V] or NULL (the distribution drawn from).  top_k in [0, 1024]; top_k == 0 requires top_p >= 1. */
int dl_sample_next(const void* logits, int B, int V, float temperature, int top_k, float top_p, const int64_t* seed,
                   int64_t* next_ids, int64_t* tokens_out, int out_stride, const int32_t* step_counter, int32_t* finished,
                   int eos_token_id, int pad_token_id, float* probs_out, cudaStream_t stream);

/* ---- instruct predictor (sparse_config["use_instruct_predictor"]), DML:2261-2375 (prefill drop), :2506-2521 (new turn),
 * CU:165-241 (per-sample persistence of the kept rows) ---------------------------------------------------------------- */
/* keep[n] = forced ? forced[n] : logit0 > logit1 over logits bf16 [B*Tq, 2]; force_last: the last token of every sample is
 * kept (DML:2520).  logits or forced may be NULL, not both. */
int dl_turn_decision(const void* logits, const uint8_t* forced, uint8_t* keep, int B, int Tq, int force_last, cudaStream_t stream);
/* After a new-turn layer's attention: of the Tq rows dl_rope_append wrote at slots [kv_len[b], kv_len[b]+Tq) of one layer's
 * arenas [B, heads, cap, hd], only those with keep[b*Tq + j] stay, moved left in place.  kv_len int32 [B] of that layer. */
int dl_kv_compact_turn(void* k_arena, void* v_arena, const int32_t* kv_len, const uint8_t* keep, int Tq, int B, int heads,
                       int head_dim, int cap, cudaStream_t stream);
/* kv_len int32 [L,B]: layers < sparse_layer += Tq, layers >= sparse_layer += sum_j keep[b*Tq + j] (keep NULL: Tq);
 * pos[b] += Tq (NULL allowed). */
int dl_kv_advance_turn(int32_t* kv_len, int32_t* pos, const uint8_t* keep, int L, int B, int sparse_layer, int Tq,
                       cudaStream_t stream);
/* out[i] = x[idx[i]] (rows of H bf16), pos_out[i] = pos[idx[i]]; idx int64 [n] on the device; pos/pos_out both NULL or both set. */
int dl_gather_rows_pos(const void* x, const int32_t* pos, const int64_t* idx, void* out, int32_t* pos_out, int n, int H,
                       cudaStream_t stream);

/* ---- no-KV-cache output sparsification (`use_cache=False`), DML:2393-2504 ---------------------------------------------------
 * logits bf16 [B, n, 2] of the output text predictor over the n generated tokens between the sticky answer index and the last
 * token.  rank int32 [B, n]: position of token i in the descending order of logit0 (ties: lower index first);
 * count int32 [B] = #{logit0 > logit1}.  The caller reads count and takes n_keep = max_b count[b] (DML:2413-2417). n <= 8192. */
int dl_output_rank(const void* logits, int32_t* rank, int32_t* count, int B, int n, cudaStream_t stream);
/* keep_index int64 [B, n_keep] (NULL allowed) = ascending {i : rank[b,i] < n_keep} (DML:2420-2426); rows int64 [B, A+n_keep+1] =
 * global source rows b*Lq + t of the thinned sequence [0, A) ++ A + keep_index ++ [Lq-1] (DML:2429-2455), A = answer_index.
 * n may be 0 (rank NULL): the first call, where the reference's empty `[A:-1]` slice leaves left ++ last (DML:2405-2409). */
int dl_output_keep_index(const int32_t* rank, int B, int n, int n_keep, int answer_index, int Lq, int64_t* keep_index,
                         int64_t* rows, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNLLAVA_B200_H */
