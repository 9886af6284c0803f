// Shared device helpers for the dynamic_llava_b200 kernels (sm_100a only).
// PTX spellings follow the CUDA 12.9 PTX ISA (tcgen05 / cp.async.bulk / mbarrier).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

// Element type of this build.  The same sources are compiled twice: libdynllava_b200.so computes in bfloat16 (default)
// and libdynllava_b200_f16.so (-DDL_HALF) in float16, the dtype the reference's eval drivers load
// (llava/model/dynamic_llava_builder.py:62).  Every kernel spells the element type `bf16` and converts only through
// the helpers below, so the two builds differ in these typedefs, the UMMA operand format and the tensor-map data type.
#ifdef DL_HALF
typedef __half bf16;
typedef __half2 bf16x2;
#define DL_UMMA_AB_FORMAT 0u  // F16
#define DL_DTYPE_CODE 1
#else
typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf16x2;
#define DL_UMMA_AB_FORMAT 1u  // BF16
#define DL_DTYPE_CODE 0
#endif

#define DL_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// error plumbing (abi.cu owns the storage)
// ---------------------------------------------------------------------------------------------
extern "C" int dl_set_error(int code, const char* where);
#define DL_CHECK_LAUNCH(where)                                   \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return dl_set_error((int)e__, where); \
    return 0;                                                    \
  } while (0)
#define DL_REQUIRE(cond, where)                      \
  do {                                               \
    if (!(cond)) return dl_set_error(-1, where ": " #cond); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the decode step is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: it may start (prologue, barrier init, weight prefetch) while its
// predecessor is still draining, executes pdl_wait() before touching anything a predecessor wrote, and calls
// pdl_launch_dependents() early so that its successor can do the same.  EVERY kernel in the chain calls pdl_wait()
// (even if it would not need to) so that "kernel N complete" still implies "kernels < N complete".
// ---------------------------------------------------------------------------------------------
DL_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
DL_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// debugging aid: DLB200_NO_PDL=<bitmask> launches kernels fully serialised (their griddepcontrol instructions become no-ops):
// 1 = every dl_launch_pdl user, 2 = the prefill GEMM (gemm2), 4 = the many-row norm kernels, 8 = K1
static inline int dl_pdl_mask() {
  static int mask = -1;
  if (mask < 0) {
    const char* e = getenv("DLB200_NO_PDL");
    mask = e != nullptr ? atoi(e) : 0;
  }
  return mask;
}
static inline int dl_pdl_enabled(int site_bit = 0) { return (dl_pdl_mask() & (1 | site_bit)) == 0 || (site_bit == 2 && (dl_pdl_mask() & 2) == 0) ? 1 : 0; }
static thread_local int dl_pdl_site = 0;  // call sites that can be switched off separately set this right before dl_launch_pdl

template <typename... KArgs, typename... Args>
static inline cudaError_t dl_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                        Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = dl_pdl_enabled(dl_pdl_site);
  dl_pdl_site = 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// bf16 <-> fp32, 128-bit vectors
// ---------------------------------------------------------------------------------------------
struct __align__(16) bf16x8 {
  bf16x2 v[4];
};

#ifdef DL_HALF
DL_DEVINL float bf2f(bf16 x) { return __half2float(x); }
DL_DEVINL bf16 f2bf(float x) { return __float2half_rn(x); }
DL_DEVINL float2 bf2f2(bf16x2 v) { return __half22float2(v); }
DL_DEVINL bf16x2 f2bf2(float a, float b) { return __floats2half2_rn(a, b); }
#else
DL_DEVINL float bf2f(bf16 x) { return __bfloat162float(x); }
DL_DEVINL bf16 f2bf(float x) { return __float2bfloat16_rn(x); }
DL_DEVINL float2 bf2f2(bf16x2 v) { return __bfloat1622float2(v); }
DL_DEVINL bf16x2 f2bf2(float a, float b) { return __floats2bfloat162_rn(a, b); }
#endif
// round-trip through the element type: the "16-bit tensor op" rounding point of the reference's eager PyTorch graph
DL_DEVINL float rbf(float x) { return bf2f(f2bf(x)); }

DL_DEVINL void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = bf2f2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
DL_DEVINL bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = f2bf2(f[2 * i], f[2 * i + 1]);
  return p;
}
DL_DEVINL bf16x8 ld8(const bf16* p) { return *reinterpret_cast<const bf16x8*>(p); }
DL_DEVINL void st8(bf16* p, const bf16x8& v) { *reinterpret_cast<bf16x8*>(p) = v; }

// streaming (read-once / write-once) 128-bit accesses that do not pollute L1
DL_DEVINL int4 ld_stream(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
DL_DEVINL void st_stream(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// sum of the skinny GEMM's split-K slices p[s * stride], s < S, in FIXED order (deterministic); loads are issued in
// groups of 8 so that one L2 round trip covers S <= 8
DL_DEVINL float4 sum_slices_f4(const float* __restrict__ p, int64_t stride, int S) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s0 = 0; s0 < S; s0 += 8) {
    float4 t[8];
#pragma unroll
    for (int g = 0; g < 8; ++g)
      if (s0 + g < S) t[g] = *reinterpret_cast<const float4*>(p + (int64_t)(s0 + g) * stride);
#pragma unroll
    for (int g = 0; g < 8; ++g)
      if (s0 + g < S) {
        a.x += t[g].x;
        a.y += t[g].y;
        a.z += t[g].z;
        a.w += t[g].w;
      }
  }
  return a;
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
DL_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DL_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// blockDim.x <= 1024; `red` is a __shared__ float[32]; result broadcast to every thread
DL_DEVINL float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  return warp_sum(t);
}

// ---------------------------------------------------------------------------------------------
// shared-memory addresses, mbarrier, TMA
// ---------------------------------------------------------------------------------------------
DL_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DL_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
DL_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DL_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DL_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DL_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DL_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time when the phase is not complete)
DL_DEVINL bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
DL_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// 1-D bulk copy global -> shared, completion counted in bytes on `bar`
DL_DEVINL void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// tiled TMA loads through a CUtensorMap (passed as a __grid_constant__ kernel parameter)
DL_DEVINL void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
DL_DEVINL void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
DL_DEVINL void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// L2 prefetch of one tensor-map box (no shared memory, no barrier): starts the HBM -> L2 stream early
DL_DEVINL void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
DL_DEVINL void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

DL_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major operand tile stored with 128-byte swizzle:
// rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO), descriptor version 1 (sm_100), layout SWIZZLE_128B (=2).
DL_DEVINL uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // LBO (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;              // SBO, bits [32,46)
  d |= (uint64_t)1 << 46;                        // version = 1, bits [46,48)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B, bits [61,64)
  return d;
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, A and B K-major.
DL_DEVINL constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4)                       // c_format = F32
         | (DL_UMMA_AB_FORMAT << 7)      // a_format = BF16 (F16 in the -DDL_HALF build)
         | (DL_UMMA_AB_FORMAT << 10)     // b_format
         | ((uint32_t)a_mn_major << 15)  // a_major
         | ((uint32_t)b_mn_major << 16)  // b_major
         | ((uint32_t)(N >> 3) << 17)    // n_dim
         | ((uint32_t)(M >> 4) << 24);   // m_dim
}
DL_DEVINL void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
DL_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DL_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DL_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// whole warp; writes the TMEM base address to *smem_slot
template <int COLS>
DL_DEVINL void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
DL_DEVINL void tmem_dealloc(uint32_t tmem_addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "n"(COLS) : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns
DL_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
DL_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
DL_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM (used to stage the bf16 P operand of the PV GEMM)
DL_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
DL_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
