// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld), fences. No CUTLASS dependency.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hero {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 %%rx;\n"
      ".reg .pred %%px;\n"
      "elect.sync %%rx|%%px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, %%px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) after ~4 s of wall clock instead
// of hanging the GPU. try_wait itself suspends for a HW-defined slice, so count time, not spins.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0 && global_timer_ns() - t0 > 4000000000ull) __trap();
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane = row).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// one fp32 column: thread (lane) <- its own row
__device__ __forceinline__ uint32_t tmem_ld_32x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster (ranks 0/1 on one TPC) run ONE 256-row MMA: each holds half of A (its 128
// rows) and half of B (128 of the 256 columns) in its own smem, the leader (rank 0) issues the
// instruction, each CTA's TMEM receives its 128 accumulator rows. Halves the smem operand traffic
// per MAC, which is what bounds the single-CTA kernel.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// In a CTA pair the shared::cluster address of the leader's copy of a smem object is the local
// shared::cta address with bit 24 cleared (CUTLASS Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
// Relaxed: the only thing this arrival publishes is "my tcgen05.ld's of the accumulator have
// completed", which the preceding tcgen05.wait::ld + tcgen05.fence::before_thread_sync already
// order. The default .release at cluster scope compiles to MEMBAR.ALL.CTA + ERRBAR, which waits
// for every outstanding global store / fp32 atomic of the epilogue (13 % of the samples of a
// pair-mode GEMM in profiles/r01d).
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(
                   smem_u32(bar) & kPeerBitMask)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint64_t* bar,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const void* tmap, uint64_t* bar,
                                                int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once the issued MMAs retire) on the barrier at the same smem offset in BOTH CTAs.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major,
                                                       int b_mn_major) {
  return (1u << 4)                                   // D format: f32
         | (1u << 7)                                 // A format: bf16
         | (1u << 10)                                // B format: bf16
         | (static_cast<uint32_t>(a_mn_major) << 15) // A major (0 = K, 1 = MN)
         | (static_cast<uint32_t>(b_mn_major) << 16) // B major
         | (static_cast<uint32_t>(N >> 3) << 17)     // N / 8
         | (static_cast<uint32_t>(M >> 4) << 24);    // M / 16
}

// ---------------------------------------------------------------- programmatic dependent launch
// Kernels launched with the programmatic-stream-serialization attribute may start while the
// previous kernel in the stream is still draining: everything before pdl_wait() (barrier init,
// TMEM allocation, descriptor prefetch) overlaps that tail; pdl_wait() blocks until the previous
// grid has completed and its writes are visible. pdl_launch_dependents() lets the NEXT kernel
// begin its own prologue early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---------------------------------------------------------------- misc
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// Counter-based per-element hash used for dropout masks: fwd and bwd regenerate the same
// mask from (key, element index); nothing is stored.
__device__ __forceinline__ uint32_t hash_u32(uint32_t key, uint32_t idx) {
  uint32_t h = idx * 0x9E3779B1u ^ key;
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
// One 32-bit hash decides TWO consecutive elements (16 bits each): keep iff bits16 >= p * 2^16.
// `threshold` is passed as p * 2^32 (C-ABI) and narrowed here; p = 0.1 -> 6553 / 65536.
__device__ __forceinline__ bool dropout_keep(uint32_t key, uint32_t idx, uint32_t threshold) {
  const uint32_t h = hash_u32(key, idx >> 1);
  const uint32_t bits = (idx & 1u) ? (h >> 16) : (h & 0xFFFFu);
  return bits >= (threshold >> 16);
}
// 8 consecutive elements starting at an even index: 4 hashes. Applies v = keep ? v * scale : 0.
__device__ __forceinline__ void dropout_apply8(float (&v)[8], uint32_t key, uint32_t idx0,
                                               uint32_t threshold, float scale) {
  const uint32_t t16 = threshold >> 16;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t h = hash_u32(key, (idx0 >> 1) + j);
    v[2 * j] = ((h & 0xFFFFu) >= t16) ? v[2 * j] * scale : 0.0f;
    v[2 * j + 1] = ((h >> 16) >= t16) ? v[2 * j + 1] * scale : 0.0f;
  }
}

// ---- attention dropout words. One hash per (query token, head, group of 8 key columns); the
// four pair-words of the group (two 15-bit lanes each, bits [0,15) and [16,31)) are the hash and
// three multiply-xorshift derivations of it (3 instructions instead of 10 per pair: the hash was
// the largest single item of the softmax loops). A lane keeps its probability iff lane15 >= t15,
// t15 = p * 2^15 (p = 0.1 -> 3276 / 32768).
__device__ __forceinline__ uint32_t attn_drop_group(uint32_t key, int tok, int heads, int head,
                                                    int col_group) {
  return hash_u32(key, ((uint32_t)tok * (uint32_t)heads + (uint32_t)head) * 128u +
                           (uint32_t)col_group);
}
__device__ __forceinline__ uint32_t attn_drop_pair(uint32_t h0, int k) {   // k = pair in group
  if (k == 0) return h0;
  uint32_t w = h0 * (k == 1 ? 0x9E3779B1u : (k == 2 ? 0x85EBCA6Bu : 0xC2B2AE35u));
  return w ^ (w >> 15);
}
// both lanes at once: 0xFFFF in each 16-bit half whose lane is kept. `k2` = (0x8000 - t15) in
// both halves: bit 15 of (lane15 + 0x8000 - t15) is set iff lane15 >= t15; PRMT then replicates
// the sign bits of bytes 1 and 3 over their halves.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ uint32_t attn_keep_sum(uint32_t w, uint32_t k2) {
  return (w & 0x7FFF7FFFu) + k2;
}
__device__ __forceinline__ uint32_t attn_keep_mask2(uint32_t w, uint32_t k2) {
  return prmt(attn_keep_sum(w, k2), 0u, 0xBB99u);
}
__device__ __forceinline__ uint32_t attn_drop_k2(uint32_t threshold) {   // threshold = p * 2^32
  const uint32_t k = 0x8000u - (threshold >> 17);
  return k | (k << 16);
}
// scalar form (long-row kernels): is key column j of this query kept?
__device__ __forceinline__ bool attn_drop_keep(uint32_t key, uint32_t threshold, int tok, int heads,
                                               int head, int j) {
  const uint32_t w = attn_drop_pair(attn_drop_group(key, tok, heads, head, j >> 3), (j >> 1) & 3);
  const uint32_t lane = (j & 1) ? ((w >> 16) & 0x7FFFu) : (w & 0x7FFFu);
  return lane >= (threshold >> 17);
}

// exp2 on the MUFU unit (one op per element is the epilogue's throughput budget on sm_100).
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 0.5 * erfc(|x| / sqrt 2) = 2^(|x| P(|x|) - 1): P is the degree-5 polynomial fitted to
// -log2(erfc(t))/t (t = |x|/sqrt 2 in [0, 6]) with the 1/sqrt2 factors, the sign and the factor
// 0.5 folded into its coefficients. Weighted fit: error of gelu < 3e-7 absolute and < 0.4 %
// relative even in the far negative tail, i.e. below bf16 resolution everywhere. One MUFU op, no
// branches, 7 FP32 instructions — CUDA's erff() costs ~2.5x that and was the GEMM epilogue's
// bottleneck.
__device__ __forceinline__ float half_erfc_abs(float x) {
  const float ax = fminf(fabsf(x), 8.4852814f);
  float p = 3.1522031349595636e-05f;
  p = fmaf(p, ax, -0.000753118481952697f);
  p = fmaf(p, ax, 0.00801779329776764f);
  p = fmaf(p, ax, -0.05329384654760361f);
  p = fmaf(p, ax, -0.4588814675807953f);
  p = fmaf(p, ax, -1.1511543989181519f);
  return fast_ex2(fmaf(p, ax, -1.0f));
}

// gelu(x) = 0.5 x (1 + erf(x / sqrt 2)) = relu(x) - |x| * 0.5 erfc(|x| / sqrt 2)
// (model/layers.py:16-25 of the reference)
__device__ __forceinline__ float gelu_erf(float x) {
  return fmaxf(x, 0.f) - fabsf(x * half_erfc_abs(x));
}
// gelu(x) and its derivative together (the forward FFN-up epilogue saves the derivative so the
// backward epilogue is a plain multiply): shares the erfc evaluation, one extra MUFU for the pdf
// (1/sqrt(2 pi) folded into the exponent).
__device__ __forceinline__ float gelu_erf_with_grad(float x, float& dgelu) {
  const float e = half_erfc_abs(x);
  const float pdf = fast_ex2(fmaf(x * x, -0.72134752044448170f, -1.3257480647361592f));
  const float cdf = 0.5f + copysignf(0.5f - e, x);
  dgelu = fmaf(x, pdf, cdf);
  return fmaxf(x, 0.f) - fabsf(x * e);
}
// ---- packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2: one issue slot, two lanes) ----
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_splat(float v) { return f2_pack(v, v); }

// gelu_erf_with_grad on two values at once: the polynomial, the pdf exponent and the final
// combinations run as packed FFMA2 (25 issue slots per pair instead of 38).
__device__ __forceinline__ void gelu_erf_with_grad2(float& x0, float& x1, float& d0, float& d1) {
  const float a0 = fminf(fabsf(x0), 8.4852814f), a1 = fminf(fabsf(x1), 8.4852814f);
  const uint64_t ax = f2_pack(a0, a1);
  uint64_t p = f2_splat(3.1522031349595636e-05f);
  p = f2_fma(p, ax, f2_splat(-0.000753118481952697f));
  p = f2_fma(p, ax, f2_splat(0.00801779329776764f));
  p = f2_fma(p, ax, f2_splat(-0.05329384654760361f));
  p = f2_fma(p, ax, f2_splat(-0.4588814675807953f));
  p = f2_fma(p, ax, f2_splat(-1.1511543989181519f));
  p = f2_fma(p, ax, f2_splat(-1.0f));
  float p0, p1;
  f2_unpack(p, p0, p1);
  const float e0 = fast_ex2(p0), e1 = fast_ex2(p1);
  const uint64_t x = f2_pack(x0, x1);
  const uint64_t t = f2_fma(f2_mul(x, x), f2_splat(-0.72134752044448170f),
                            f2_splat(-1.3257480647361592f));
  float t0, t1;
  f2_unpack(t, t0, t1);
  const uint64_t pdf = f2_pack(fast_ex2(t0), fast_ex2(t1));
  const uint64_t cdf = f2_pack(0.5f + copysignf(0.5f - e0, x0), 0.5f + copysignf(0.5f - e1, x1));
  f2_unpack(f2_fma(x, pdf, cdf), d0, d1);
  float xe0, xe1;
  f2_unpack(f2_mul(x, f2_pack(e0, e1)), xe0, xe1);
  x0 = fmaxf(x0, 0.f) - fabsf(xe0);
  x1 = fmaxf(x1, 0.f) - fabsf(xe1);
}

__device__ __forceinline__ float gelu_erf_grad(float x) {
  float d;
  (void)gelu_erf_with_grad(x, d);
  return d;
}

// TMA store of a smem box (bulk async group) and its group bookkeeping.
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tmap)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace hero
