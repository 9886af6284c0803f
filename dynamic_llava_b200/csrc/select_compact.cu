// K1: vision-token selection + gather/compact in ONE kernel
// (dynamic_modeling_llama.py:1867, :1897-1994): score = log_softmax(logits)[..., 0] (fp32 math, ONE rounding to bf16: torch's
// CUDA kernel; torch's CPU bf16 path rounds intermediates and is one ulp off on ~45 % of the scores),
// keep = ascending indices of the k best scores, hidden' = [left | image[keep] | right], RoPE positions stay the
// ORIGINAL (pre-drop) positions, and a CSR header (cu_seqlens) describes the packed result.
//
// Grid = (R, B).  Every CTA of sample b redoes the selection for its sample: ONE selection thread per image token holds that
// token's 16-bit score key in a register, the k-th largest key is found by a 16-step bitwise search in which every step is a
// single counting barrier (`bar.red.popc` on the selection threads' named barrier: count(key >= candidate) in one instruction, no
// shared-memory histogram, no atomics), one more count gives #keys above it, and an ordered compaction (warp ballots + the
// per-warp totals) writes the ascending index list.  ~20 barriers of ~50 cycles: well under a microsecond, so replicating it per
// CTA is cheaper than a second kernel or a round trip of the index map through HBM.
// The rows move through a shared-memory ring of whole-row slots driven by ONE thread (warp 0) with bulk async copies
// (cp.async.bulk global->shared on an mbarrier, shared->global in bulk groups): no register staging, 64 KB in flight per
// CTA.  Rows are visited text-first: the [left | right] text rows do not depend on the selection, so the copy thread streams
// them from the first cycle while warps 1-8 run the radix select on their own named barrier; it only waits (mbarrier
// `sel_done`) right before its first IMAGE row.
// Ties at the k-th score are broken towards the lower index (= stable descending sort).
// HBM-bound: algorithmic bytes = 2 * B * P' * H * 2.
#include "common.cuh"
#include "../../include/dynllava_b200.h"

#define SC_MAX_IMG 1024
#define SC_MAX_BLOCK (SC_MAX_IMG + 32)  // one selection thread per image token (rounded up to whole warps) + warp 0: the copy engine driver
#define SC_MAX_SLOTS 16
#define SC_RING_BYTES 65536

// shared -> global bulk copy (async proxy), tracked by bulk groups
DL_DEVINL void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
DL_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
DL_DEVINL void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// barrier 1 = the selection threads only (n_sel of them: whole warps).  sel_count: block-wide popcount of a predicate in ONE
// barrier instruction; every thread gets the same total.
DL_DEVINL void sel_sync(int n_sel) { asm volatile("bar.sync 1, %0;" ::"r"(n_sel) : "memory"); }
DL_DEVINL int sel_count(int n_sel, bool pred) {
  int c;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.s32 p, %2, 0;\n"
      "bar.red.popc.u32 %0, 1, %1, p;\n"
      "}"
      : "=r"(c)
      : "r"(n_sel), "r"((int)pred)
      : "memory");
  return c;
}

__global__ void __launch_bounds__(SC_MAX_BLOCK) select_compact_kernel(
    const bf16* __restrict__ logits,         // [B, n_img, 2]
    const int64_t* __restrict__ forced_keep, // [B, k] or null
    const bf16* __restrict__ hid_in,         // [B, P, H]
    bf16* __restrict__ hid_out,              // [B, P', H]
    int64_t* __restrict__ keep_index,        // [B, k]
    int32_t* __restrict__ pos_out,           // [B, P']
    int32_t* __restrict__ cu_seqlens,        // [B + 1]
    int P, int H, int s0, int n_img, int k, int n_slots) {
  extern __shared__ __align__(128) uint8_t ring[];  // n_slots whole rows
  __shared__ __align__(8) uint64_t full[SC_MAX_SLOTS];
  __shared__ __align__(8) uint64_t sel_done;  // count 1: kidx[] is final (one arrival after the selection threads' last barrier)
  __shared__ int16_t kidx[SC_MAX_IMG];
  __shared__ int wsum[SC_MAX_IMG / 32];
  const int b = blockIdx.y;
  const int tid = (int)threadIdx.x - 32;  // selection thread index == image token index; warp 0 (tid < 0) drives the copies
  const int n_sel = (int)blockDim.x - 32;  // n_img rounded up to whole warps
  const int R = gridDim.x, r = blockIdx.x;
  const int Pp = P - n_img + k;
  const int n_text = Pp - k;
  const uint32_t row_bytes = (uint32_t)H * 2u;
  const bf16* in_b = hid_in + (int64_t)b * P * H;
  bf16* out_b = hid_out + (int64_t)b * Pp * H;
  // this CTA's rows in text-first order: v = r, r + R, ...; v < n_text is text row v (left part, then right part),
  // v >= n_text is kept image row v - n_text
  const int n_mine = r < Pp ? (Pp - r + R - 1) / R : 0;
  int issued = 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_slots; ++s) mbar_init(&full[s], 1);
    mbar_init(&sel_done, 1);
    fence_barrier_init();
  }
  __syncthreads();
  // launched with the programmatic-serialization attribute: the block is resident and its barriers are initialised while the
  // producer of `logits` / `hid_in` (the predictor's last Linear) is still draining; nothing of the predecessor is read above
  pdl_wait();
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    // ---- copy engine: ring of whole-row slots, loads n_slots - 1 deep, stores in bulk groups.  It never takes part in the
    // selection: text rows flow from the first cycle, and it only waits for kidx[] right before its first IMAGE row. ----
    bool sel_ready = false;
    auto src_row = [&](int v) {
      if (v < n_text) return v < s0 ? v : v + n_img;
      if (!sel_ready) {
        mbar_wait(&sel_done, 0);
        sel_ready = true;
      }
      return s0 + (int)kidx[v - n_text];
    };
    auto dst_row = [&](int v) { return v < n_text ? (v < s0 ? v : v + k) : s0 + (v - n_text); };
    auto issue_load = [&](int i) {
      const int s = i % n_slots;
      const int sj = src_row(r + i * R);
      mbar_expect_tx(&full[s], row_bytes);
      tma_bulk_g2s(ring + (size_t)s * row_bytes, in_b + (int64_t)sj * H, row_bytes, &full[s]);
    };
    while (issued < n_mine && issued < n_slots) issue_load(issued++);
    for (int i = 0; i < n_mine; ++i) {
      const int s = i % n_slots;
      mbar_wait(&full[s], (uint32_t)(i / n_slots) & 1u);
      bulk_s2g(out_b + (int64_t)dst_row(r + i * R) * H, ring + (size_t)s * row_bytes, row_bytes);
      bulk_commit();
      if (i >= 1 && issued < n_mine) {  // slot of row i-1 is free once its store has read it
        bulk_wait_read<1>();
        issue_load(issued++);
      }
    }
    bulk_wait_read<0>();
    return;
  }
  if (tid < 0) return;  // the rest of warp 0

  if (forced_keep != nullptr) {
    for (int i = tid; i < k; i += n_sel) kidx[i] = (int16_t)forced_keep[(int64_t)b * k + i];
  } else if (k == n_img) {  // keep rate 1.0: the ascending indices of ALL tokens, whatever the scores
    for (int i = tid; i < k; i += n_sel) kidx[i] = (int16_t)i;
  } else if (k > 0) {
    // this thread's token: score rounded where the reference's 16-bit log_softmax rounds, mapped to an order-preserving 16-bit key
    const bool valid = tid < n_img;
    uint32_t key = 0;
    if (valid) {
      const bf16x2 l = *reinterpret_cast<const bf16x2*>(logits + ((int64_t)b * n_img + tid) * 2);
      const float2 lf = bf2f2(l);
      const float m = fmaxf(lf.x, lf.y);
      const bf16 sc = f2bf((lf.x - m) - logf(expf(lf.x - m) + expf(lf.y - m)));
      uint16_t bits = *reinterpret_cast<const uint16_t*>(&sc);
      if (bits == 0x8000u) bits = 0;  // -0 == +0
      key = (bits & 0x8000u) ? (uint16_t)~bits : (uint16_t)(bits | 0x8000u);
    }
    // T = the k-th largest key = the largest T with count(key >= T) >= k, built bit by bit from the top
    uint32_t T = 0;
#pragma unroll
    for (int bit = 15; bit >= 0; --bit) {
      const uint32_t cand = T | (1u << bit);
      if (sel_count(n_sel, valid && key >= cand) >= k) T = cand;
    }
    const bool gt = valid && key > T, eq = valid && key == T;
    const int need_eq = k - sel_count(n_sel, gt);
    // ordered compaction: exclusive (#gt, #eq) before this token = lanes below in the warp + the warps before
    const unsigned m_gt = __ballot_sync(0xffffffffu, gt), m_eq = __ballot_sync(0xffffffffu, eq);
    const unsigned below = (1u << (tid & 31)) - 1u;
    if ((tid & 31) == 0) wsum[tid >> 5] = __popc(m_gt) | (__popc(m_eq) << 16);
    sel_sync(n_sel);
    int base = 0;
    for (int w = 0; w < (tid >> 5); ++w) base += wsum[w];
    const int gt_before = (base & 0xffff) + __popc(m_gt & below), eq_before = (base >> 16) + __popc(m_eq & below);
    // ties at the k-th score go to the lower index (= stable descending sort): the first need_eq equal keys are kept
    if (gt)
      kidx[gt_before + min(eq_before, need_eq)] = (int16_t)tid;
    else if (eq && eq_before < need_eq)
      kidx[gt_before + eq_before] = (int16_t)tid;
  }
  sel_sync(n_sel);

  // ONE arrival (hundreds of arrivals on one mbarrier word would serialise): the named barrier above ordered every selection
  // thread's kidx[] writes before this release, the copy thread's wait is the acquire
  if (tid == 0) mbar_arrive(&sel_done);
  if (r == 0) {  // index map + CSR header
    for (int i = tid; i < k; i += n_sel) keep_index[(int64_t)b * k + i] = kidx[i];
    for (int j = tid; j < Pp; j += n_sel) {
      int p = j < s0 ? j : (j < s0 + k ? s0 + kidx[j - s0] : j + (n_img - k));
      pos_out[(int64_t)b * Pp + j] = p;
    }
    if (tid == 0) {
      cu_seqlens[b + 1] = (b + 1) * Pp;
      if (b == 0) cu_seqlens[0] = 0;
    }
  }
}

extern "C" int dl_select_compact(const void* logits, const int64_t* forced_keep, const void* hidden_in, void* hidden_out,
                                 int64_t* keep_index, int32_t* pos_out, int32_t* cu_seqlens, int B, int P, int H,
                                 int img_start, int n_img, int k, int ctas_per_sample, cudaStream_t stream) {
  DL_REQUIRE(H % 8 == 0 && H * 2 <= SC_RING_BYTES && n_img <= SC_MAX_IMG && k >= 0 && k <= n_img && img_start >= 0 &&
                 img_start + n_img <= P,
             "dl_select_compact");
  DL_REQUIRE((((uintptr_t)hidden_in | (uintptr_t)hidden_out) & 15) == 0, "dl_select_compact (16-byte aligned rows)");
  if (B == 0) return 0;
  const int Pp = P - n_img + k;
  if (Pp == 0) return 0;
  int n_slots = SC_RING_BYTES / (H * 2);
  if (n_slots > SC_MAX_SLOTS) n_slots = SC_MAX_SLOTS;
  const int smem = n_slots * H * 2;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(select_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_RING_BYTES);
    if (e != cudaSuccess) return dl_set_error((int)e, "dl_select_compact(smem attr)");
    attr_set = true;
  }
  int R = ctas_per_sample;
  if (R <= 0) {
    // measured (tools/k1_sweep.py): text rows stream best in short CTAs (~10 rows each; more, shorter CTAs also even out the
    // tail), image rows in long ones (every CTA repeats the ~1 us select before its first image row)
    const int n_free = (k == n_img && forced_keep == nullptr) ? Pp : Pp - k;  // rows whose source is known without the select
    R = (n_free + 9) / 10 + (Pp - n_free + 71) / 72;
    if (R * B < 148) R = B <= 148 ? 148 / B : 1;  // small problems: one CTA per SM (never a second, partial wave)
  }
  if (R > Pp) R = Pp;
  if (R < 1) R = 1;
  const int n_sel = n_img > 0 ? (n_img + 31) / 32 * 32 : 32;  // one selection thread per image token, whole warps
  dl_pdl_site = 8;
  cudaError_t le = dl_launch_pdl(select_compact_kernel, dim3(R, B), dim3(32 + n_sel), (size_t)smem, stream, (const bf16*)logits,
                                 forced_keep, (const bf16*)hidden_in, (bf16*)hidden_out, keep_index, pos_out, cu_seqlens, P, H,
                                 img_start, n_img, k, n_slots);
  if (le != cudaSuccess) return dl_set_error((int)le, "dl_select_compact(launch)");
  DL_CHECK_LAUNCH("dl_select_compact");
}
