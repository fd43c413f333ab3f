// Variable-length multi-head self-attention on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
// HERO's sequences are short (cross-modal rows ~10-70 tokens, temporal rows <= 100 frames), so a
// 128-row MMA tile would be mostly empty if it held one sequence. Instead the host packs
// CONSECUTIVE sequences of the packed token stream into tiles of <= 128 tokens (a sequence never
// straddles tiles); one CTA handles one (tile, head):
//
//   forward   S = Q K^T (128x128x64)  -> per-row softmax restricted to the row's own sequence
//             (block-diagonal mask = the key-padding mask of model/layers.py:299-302 in the packed
//             layout) -> P (bf16, smem) -> O = P V (128x64x128) -> ctx
//   backward  S = Q K^T, dP = dO V^T -> P, dS rows -> dV = P^T dO, dK = dS^T Q, dQ = dS K
//
// Q/K/V/dO tiles arrive by TMA (128-byte swizzle); every contraction is a tcgen05.mma with fp32
// accumulators in TMEM. The transposed operands (P^T, dS^T, V/dO/Q/K as [k][n]) need no copies:
// the SAME smem tiles are addressed as MN-major operands through the UMMA descriptors.
// Softmax statistics, exp2, dropout masks and the P / dS tiles never touch HBM.
//
// Replaces model/layers.py:129-160 (BertSelfAttention after the QKV projections) and its autograd
// backward. D_i = sum_j P_ij dP_ij is taken as dO_i . O_i from the saved forward output.
#include <cuda.h>

#include "common.h"
#include "ptx.cuh"

namespace hero {

constexpr int AT_ROWS = 128;           // tokens per tile
constexpr int AT_D = 64;               // head dim
constexpr int AT_TILE_BYTES = AT_ROWS * AT_D * 2;   // 16 KB: one [128 x 64] bf16 operand tile
constexpr int AT_P_BYTES = AT_ROWS * AT_ROWS * 2;   // 32 KB: one [128 x 128] bf16 tile (2 chunks)

// Block-diagonal mask on the tensor core. A fifth K = 16 step of the S = Q K^T MMA multiplies the
// tile's sequence-membership matrix E [128 tokens x 16 sequence ordinals] (128.0 at a token's own
// sequence, else 0) with itself: S'_ij = S_ij + 16384 * [seq(i) == seq(j)]. Softmax is
// shift-invariant per row, so the row's own sequence sees its plain softmax while every other
// column (other sequences of the tile, rows past the tile's end) sits 16384 raw = 2955 log2 units
// lower and its exp2 is exactly 0: the softmax loops need no compares, selects or predicates.
// 16384 is exact in fp32 next to |S| < 2^10 up to an absolute error of 2^-9 (2.4e-4 relative in
// p, an eighth of P's bf16 rounding). The reference adds -10000 to the scaled scores
// (model/layers.py:299-302), i.e. 80000 raw: equally finite. The host plan keeps <= AT_MAX_SEQS
// sequences per tile.
constexpr int AT_MAX_SEQS = 16;
constexpr int AT_MASK_BYTES = 2 * AT_MAX_SEQS * 128;   // [k = ordinal][mn = token], 2 chunks of 64 tokens
constexpr float AT_MASK_BIG = 16384.0f;

// Column i of the membership operand (MN-major: one 128-byte row per ordinal, tokens contiguous,
// two 64-token chunks of 2 KB): thread i owns its token's 16 entries.
template <int K0 = 0, int K1 = AT_MAX_SEQS>
__device__ __forceinline__ void write_membership(uint8_t* sE, int i, int ord) {
  uint8_t* base = sE + (i >> 6) * (AT_MAX_SEQS * 128) + ((i & 7) << 1);
  const int u = (i & 63) >> 3;
#pragma unroll
  for (int k = K0; k < K1; ++k)
    *reinterpret_cast<uint16_t*>(base + k * 128 + ((u ^ (k & 7)) << 4)) =
        (k == ord) ? static_cast<uint16_t>(0x4300) : static_cast<uint16_t>(0);   // bf16 128.0
}

// Ordinal (0-based) of thread i's sequence within the tile, -1 for rows past the tile's end:
// the number of sequence starts at rows <= i, minus one. `warp_starts` is 4 ints of smem; the
// caller synchronises the CTA between the two halves.
__device__ __forceinline__ uint32_t seq_starts_ballot(bool valid, int lo, int i, int* warp_starts) {
  const uint32_t bal = __ballot_sync(0xffffffffu, valid && lo == i);
  // (a 256-thread CTA has two warps per row quadrant: both write the same count)
  if ((threadIdx.x & 31) == 0) warp_starts[(threadIdx.x >> 5) & 3] = __popc(bal);
  return bal;
}
__device__ __forceinline__ int seq_ordinal(bool valid, uint32_t bal, const int* warp_starts) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 3;
  int ord = __popc(bal & (0xffffffffu >> (31 - lane))) - 1;
  for (int w = 0; w < warp; ++w) ord += warp_starts[w];
  return valid ? ord : -1;
}

// element (row i, col j) of a [128 x 128] bf16 tile stored as two 64-column chunks of
// [128 rows x 128 B] with the 128-byte swizzle: byte offset of the 16-byte unit holding cols
// [8u', 8u'+8) where j = 64*chunk + 8u + (j % 8).
__device__ __forceinline__ uint32_t p_unit_offset(int row, int chunk, int u) {
  return chunk * (AT_ROWS * 128) + row * 128 + ((u ^ (row & 7)) << 4);
}

__device__ __forceinline__ float ex2(float x) { return fast_ex2(x); }

struct AttnTcArgs {
  const int32_t* tile_tok0;
  const int32_t* tile_ntok;
  const int32_t* seq_lo;
  const int32_t* seq_hi;
  int heads;
  int H;
  float scale_log2;      // log2(e) / sqrt(d)
  float scale;           // 1 / sqrt(d)
  uint32_t drop_thr, drop_key;
  float drop_scale;
};

// Each thread parks its 64-feature row (two 32-column TMEM fragments, optionally scaled) in a
// swizzled [128 x 128 B] smem tile ...
__device__ __forceinline__ void stage_row(uint8_t* tile, int row, const uint32_t (&r0)[32],
                                          const uint32_t (&r1)[32], float scale) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint32_t* r = g < 4 ? r0 : r1;
    const int b = (g & 3) * 8;
    uint4 v;
    v.x = pack_bf16x2(__uint_as_float(r[b]) * scale, __uint_as_float(r[b + 1]) * scale);
    v.y = pack_bf16x2(__uint_as_float(r[b + 2]) * scale, __uint_as_float(r[b + 3]) * scale);
    v.z = pack_bf16x2(__uint_as_float(r[b + 4]) * scale, __uint_as_float(r[b + 5]) * scale);
    v.w = pack_bf16x2(__uint_as_float(r[b + 6]) * scale, __uint_as_float(r[b + 7]) * scale);
    *reinterpret_cast<uint4*>(tile + row * 128 + ((g ^ (row & 7)) << 4)) = v;
  }
}
// 32 of the row's 64 features (one TMEM fragment): units 4 * half .. 4 * half + 3
__device__ __forceinline__ void stage_half_row(uint8_t* tile, int row, int half,
                                               const uint32_t (&r)[32], float scale) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int b = g * 8;
    uint4 v;
    v.x = pack_bf16x2(__uint_as_float(r[b]) * scale, __uint_as_float(r[b + 1]) * scale);
    v.y = pack_bf16x2(__uint_as_float(r[b + 2]) * scale, __uint_as_float(r[b + 3]) * scale);
    v.z = pack_bf16x2(__uint_as_float(r[b + 4]) * scale, __uint_as_float(r[b + 5]) * scale);
    v.w = pack_bf16x2(__uint_as_float(r[b + 6]) * scale, __uint_as_float(r[b + 7]) * scale);
    *reinterpret_cast<uint4*>(tile + row * 128 + (((half * 4 + g) ^ (row & 7)) << 4)) = v;
  }
}
// ... and the CTA then writes the tile(s) out with every warp covering 4 full 128-byte rows per
// instruction (a lane-per-row store would touch 32 different lines per instruction and choked the
// LSU: `lg_throttle` was the top stall of the first version). Rows >= ntok belong to the next tile.
__device__ __forceinline__ void store_tiles(const uint8_t* tiles, int n_parts, __nv_bfloat16* base,
                                            long long row_stride, long long part_stride, int ntok) {
  for (int idx = threadIdx.x; idx < n_parts * AT_ROWS * 8; idx += blockDim.x) {
    const int part = idx >> 10, row = (idx >> 3) & 127, u = idx & 7;
    if (row < ntok) {
      const uint4 v = *reinterpret_cast<const uint4*>(tiles + part * AT_TILE_BYTES + row * 128 +
                                                      ((u ^ (row & 7)) << 4));
      *reinterpret_cast<uint4*>(base + row * row_stride + part * part_stride + u * 8) = v;
    }
  }
}

// ------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(128)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnTcArgs a,
                   __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  // 48 KB of operand tiles; P (32 KB) is written over Q|K once S = QK^T has retired, and O reuses
  // S's TMEM columns, so 4 CTAs fit per SM (smem 49 KB, 128 TMEM columns each) and their TMA /
  // MMA / softmax phases overlap.
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE_BYTES;
  uint8_t* sV = smem + 2 * AT_TILE_BYTES;
  uint8_t* sP = smem;
  uint8_t* sE = smem + 3 * AT_TILE_BYTES;           // sequence-membership operand (4 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sE + AT_MASK_BYTES);
  uint64_t* tma_bar = bars;
  uint64_t* mma_bar = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  int* warp_starts = reinterpret_cast<int*>(bars + 3);

  const int tile = blockIdx.x, head = blockIdx.y;
  const int warp = threadIdx.x >> 5;
  const int tok0 = a.tile_tok0[tile];
  const int ntok = a.tile_ntok[tile];
  // (the plan arrays were uploaded long before the previous kernel started: safe before pdl_wait)
  const int i = threadIdx.x;
  const bool valid = i < ntok;
  int lo = 0, hi = 0;
  if (valid) {
    lo = a.seq_lo[tok0 + i] - tok0;
    hi = a.seq_hi[tok0 + i] - tok0;
  }
  const uint32_t starts = seq_starts_ballot(valid, lo, i, warp_starts);

  // The tile loads go out first: they need neither TMEM nor the other threads, and the wait for
  // them was the top stall of the kernel (ncu: 23 % of the samples on this barrier's spin loop)
  // while TMEM allocation and the membership writes sat in front of their issue.
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(tma_bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
    pdl_wait();
    mbar_arrive_expect_tx(tma_bar, 3 * AT_TILE_BYTES);
    tma_load_2d(sQ, &tmap_qkv, tma_bar, head * AT_D, tok0);
    tma_load_2d(sK, &tmap_qkv, tma_bar, a.H + head * AT_D, tok0);
    tma_load_2d(sV, &tmap_qkv, tma_bar, 2 * a.H + head * AT_D, tok0);
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  {
    const int ord = seq_ordinal(valid, starts, warp_starts);
    if (ord >= AT_MAX_SEQS) __trap();     // the plan packs <= AT_MAX_SEQS sequences per tile
    write_membership(sE, i, ord);
    fence_proxy_async();
  }
  __syncthreads();
  mbar_wait(tma_bar, 0);

  if (threadIdx.x == 0) {
    tc_fence_after_sync();
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
#pragma unroll
    for (int k = 0; k < AT_D / 16; ++k) {
      const uint64_t ad = make_sw128_desc(smem_u32(sQ) + k * 32, 16, 1024);
      const uint64_t bd = make_sw128_desc(smem_u32(sK) + k * 32, 16, 1024);
      umma_f16(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    // + 16384 where query and key token belong to the same sequence (E E^T, both MN-major)
    const uint64_t ed = make_sw128_desc(smem_u32(sE), AT_MAX_SEQS * 128, 1024);
    umma_f16(tmem, ed, ed, make_idesc_bf16(128, 128, 1, 1), 1u);
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after_sync();

  // ---- softmax on this thread's row: the other sequences' columns underflow to exactly 0 ----
  // warp-uniform column range (tcgen05.ld is warp-collective)
  int wlo = valid ? lo : AT_ROWS, whi = hi;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wlo = min(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
    whi = max(whi, __shfl_xor_sync(0xffffffffu, whi, o));
  }
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  float mx = -INFINITY;
  for (int c = 0; c < 4; ++c) {
    if (c * 32 >= whi || c * 32 + 32 <= wlo) continue;   // warp-uniform
    uint32_t r[32];
    tmem_ld_32x32(t_row + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
  }
  const float nmx = -mx * a.scale_log2;
  float sum0 = 0.f, sum1 = 0.f;
  const uint32_t k2 = attn_drop_k2(a.drop_thr);
  for (int c = 0; c < 4; ++c) {
    uint32_t pk[16];
    const bool touch = !(c * 32 >= whi || c * 32 + 32 <= wlo);   // warp-uniform
    if (touch) {
      uint32_t r[32];
      tmem_ld_32x32(t_row + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = ex2(fmaf(__uint_as_float(r[j]), a.scale_log2, nmx));
        const float p1 = ex2(fmaf(__uint_as_float(r[j + 1]), a.scale_log2, nmx));
        sum0 += p0;
        sum1 += p1;
        pk[j >> 1] = pack_bf16x2(p0, p1);
      }
      if (a.drop_thr != 0u) {     // the 1 / (1 - p) scale is folded into the output row scale
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t h0 = attn_drop_group(a.drop_key, tok0 + i, a.heads, head, c * 4 + g);
#pragma unroll
          for (int k = 0; k < 4; ++k) pk[g * 4 + k] &= attn_keep_mask2(attn_drop_pair(h0, k), k2);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = 0u;
    }
    // 32 columns = 4 units of 16 B inside 64-column chunk (c >> 1), units (c & 1) * 4 + g
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 v = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
      *reinterpret_cast<uint4*>(sP + p_unit_offset(i, c >> 1, (c & 1) * 4 + g)) = v;
    }
  }
  const float sum = sum0 + sum1;
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after_sync();
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);   // A = P (K-major), B = V (MN-major)
#pragma unroll
    for (int k = 0; k < AT_ROWS / 16; ++k) {
      const uint64_t ad =
          make_sw128_desc(smem_u32(sP) + (k >> 2) * (AT_ROWS * 128) + (k & 3) * 32, 16, 1024);
      const uint64_t bd = make_sw128_desc(smem_u32(sV) + k * 2048, AT_TILE_BYTES, 1024);
      umma_f16(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);   // O overwrites S's columns [0, 64)
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 1);
  tc_fence_after_sync();

  {
    const float inv = (sum > 0.f ? 1.0f / sum : 0.f) * (a.drop_thr != 0u ? a.drop_scale : 1.0f);
    uint32_t r0[32], r1[32];
    tmem_ld_32x32(t_row, r0);
    tmem_ld_32x32(t_row + 32, r1);
    tmem_ld_wait();
    if (valid && lse != nullptr)   // log2-domain log-sum-exp of the scaled scores, for the backward
      lse[(long long)(tok0 + i) * a.heads + head] = (mx - AT_MASK_BIG) * a.scale_log2 + log2f(sum);
    stage_row(sQ, i, r0, r1, inv);   // Q's tile is dead: P.V has retired
  }
  __syncthreads();
  store_tiles(sQ, 1, ctx + (long long)tok0 * a.H + head * AT_D, a.H, 0, ntok);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 128);
  }
}

// ------------------------------------------------------------------------------------ backward
// TMEM (256 columns): S [0,128) and dP [128,256); once every thread has turned them into P / dS,
// dQ [0,64), dK [64,128), dV [128,192) reuse the same columns. smem (112 KB): Q, K, V, dO tiles;
// P's first 64-column chunk is written over V (dead after dP = dO V^T), so two CTAs fit per SM and
// overlap each other's TMA / MMA / softmax phases. Probabilities are rebuilt from the forward's
// log-sum-exp in ONE pass over S.
// 256 threads: TWO threads per tile row (warps w and w + 4 share TMEM lane quadrant w), each owning
// 64 of the row's 128 score columns and 32 of the 64 features of every output row. With only two
// CTAs per SM (TMEM: 2 x 256 columns) the 128-thread version left the SM with 8 warps to hide the
// TMEM / MUFU / smem latencies of its longest phase.
constexpr int AT_BWD_THREADS = 256;
__global__ void __launch_bounds__(AT_BWD_THREADS)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                   const __grid_constant__ CUtensorMap tmap_do,
                   const __grid_constant__ CUtensorMap tmap_o, const AttnTcArgs a,
                   const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv,
                   float* __restrict__ dbias) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE_BYTES;
  uint8_t* sV = smem + 2 * AT_TILE_BYTES;
  uint8_t* sdO = smem + 3 * AT_TILE_BYTES;
  // P chunk 0 over V, chunk 1 after dO: chunk stride = 2 tiles
  uint8_t* sP = sV;
  constexpr uint32_t P_CHUNK = 2 * AT_TILE_BYTES;
  uint8_t* sdS = smem + 5 * AT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + AT_P_BYTES);
  uint64_t* tma_bar = bars;
  uint64_t* mma_bar = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  int* warp_starts = reinterpret_cast<int*>(bars + 3);
  uint8_t* sE = sdS;     // membership operand of the S MMA; dS is written after that MMA retired

  const int tile = blockIdx.x, head = blockIdx.y;
  const int warp = threadIdx.x >> 5;
  const int tok0 = a.tile_tok0[tile];
  const int ntok = a.tile_ntok[tile];
  const int i = threadIdx.x & 127;        // tile row
  const int half = threadIdx.x >> 7;      // which 64 score columns / 32 output features
  const bool valid = i < ntok;
  int lo = 0, hi = 0;
  if (valid) {
    lo = a.seq_lo[tok0 + i] - tok0;
    hi = a.seq_hi[tok0 + i] - tok0;
  }
  const uint32_t starts = seq_starts_ballot(valid, lo, i, warp_starts);

  // tile loads first (see the forward kernel)
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_o);
    mbar_init(tma_bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
    pdl_wait();
    mbar_arrive_expect_tx(tma_bar, 5 * AT_TILE_BYTES);
    tma_load_2d(sQ, &tmap_qkv, tma_bar, head * AT_D, tok0);
    tma_load_2d(sK, &tmap_qkv, tma_bar, a.H + head * AT_D, tok0);
    tma_load_2d(sV, &tmap_qkv, tma_bar, 2 * a.H + head * AT_D, tok0);
    tma_load_2d(sdO, &tmap_do, tma_bar, head * AT_D, tok0);
    // the forward output O lands where P's second chunk will be written later; every thread has
    // read its O row before the CTA barrier that precedes the first MMA
    tma_load_2d(sP + P_CHUNK, &tmap_o, tma_bar, head * AT_D, tok0);
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();


  {
    const int ord = seq_ordinal(valid, starts, warp_starts);
    if (ord >= AT_MAX_SEQS) __trap();
    if (half == 0) write_membership<0, AT_MAX_SEQS / 2>(sE, i, ord);
    else write_membership<AT_MAX_SEQS / 2, AT_MAX_SEQS>(sE, i, ord);
    fence_proxy_async();
  }
  __syncthreads();
  mbar_wait(tma_bar, 0);
  // S, dP go first; D_i is computed while they run
  if (threadIdx.x == 0) {
    tc_fence_after_sync();
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
#pragma unroll
    for (int k = 0; k < AT_D / 16; ++k) {   // S = Q K^T
      const uint64_t ad = make_sw128_desc(smem_u32(sQ) + k * 32, 16, 1024);
      const uint64_t bd = make_sw128_desc(smem_u32(sK) + k * 32, 16, 1024);
      umma_f16(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    {   // + 16384 on same-sequence pairs (see AT_MASK_BIG)
      const uint64_t ed = make_sw128_desc(smem_u32(sE), AT_MAX_SEQS * 128, 1024);
      umma_f16(tmem, ed, ed, make_idesc_bf16(128, 128, 1, 1), 1u);
    }
#pragma unroll
    for (int k = 0; k < AT_D / 16; ++k) {   // dP = dO V^T
      const uint64_t ad = make_sw128_desc(smem_u32(sdO) + k * 32, 16, 1024);
      const uint64_t bd = make_sw128_desc(smem_u32(sV) + k * 32, 16, 1024);
      umma_f16(tmem + 128, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(mma_bar);
  }
  // D_i = dO_i . O_i from the swizzled smem tiles (conflict-free 16-byte reads)
  float Di = 0.f;
  if (valid) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const uint32_t off = i * 128 + ((g ^ (i & 7)) << 4);
      const uint4 o = *reinterpret_cast<const uint4*>(sP + P_CHUNK + off);
      const uint4 d = *reinterpret_cast<const uint4*>(sdO + off);
      const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = unpack_bf16x2(ow[j]), y = unpack_bf16x2(dw[j]);
        Di = fmaf(x.x, y.x, Di);
        Di = fmaf(x.y, y.y, Di);
      }
    }
  }

  // both threads of a row have read its O row: its smem is P's second chunk from here on
  __syncthreads();
  mbar_wait(mma_bar, 0);
  tc_fence_after_sync();

  int wlo = valid ? lo : AT_ROWS, whi = hi;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wlo = min(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
    whi = max(whi, __shfl_xor_sync(0xffffffffu, whi, o));
  }
  const uint32_t t_row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  // exponent offset of this row: its log-sum-exp plus the membership shift; rows past the tile's
  // end get +inf so that their P and dS rows (which the transposed products sum over) are 0
  const float row_c = valid ? fmaf(AT_MASK_BIG, a.scale_log2,
                                   lse[(long long)(tok0 + i) * a.heads + head])
                            : INFINITY;
  const bool drop = a.drop_thr != 0u;
  const float keep_scale = drop ? a.drop_scale : 1.0f;
  const uint32_t ks_bits = __float_as_uint(keep_scale);
  const uint32_t k2 = attn_drop_k2(a.drop_thr);
  // Stored tiles: P = bf16(p) on kept lanes (dV = keep_scale * P^T dO, scaled when dV is staged),
  // dS = p * (dP * keep - D) (dQ, dK scaled by 1 / sqrt(d) when staged; a power of two).
  for (int c = 2 * half; c < 2 * half + 2; ++c) {
    uint32_t pk[16], dk[16];
    const bool touch = !(c * 32 >= whi || c * 32 + 32 <= wlo);
    if (touch) {
      uint32_t r[32], d[32];
      tmem_ld_32x32(t_row + c * 32, r);
      tmem_ld_32x32(t_row + 128 + c * 32, d);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t h0 = 0u;
        if (drop) h0 = attn_drop_group(a.drop_key, tok0 + i, a.heads, head, c * 4 + g);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = g * 8 + k * 2;
          uint32_t m2 = 0xFFFFFFFFu;
          float keep0 = 1.0f, keep1 = 1.0f;
          if (drop) {
            const uint32_t sg = attn_keep_sum(attn_drop_pair(h0, k), k2);
            m2 = prmt(sg, 0u, 0xBB99u);
            keep0 = __uint_as_float(prmt(sg, 0u, 0x9999u) & ks_bits);
            keep1 = __uint_as_float(prmt(sg, 0u, 0xBBBBu) & ks_bits);
          }
          const float p0 = ex2(fmaf(__uint_as_float(r[j]), a.scale_log2, -row_c));
          const float p1 = ex2(fmaf(__uint_as_float(r[j + 1]), a.scale_log2, -row_c));
          const float ds0 = p0 * fmaf(__uint_as_float(d[j]), keep0, -Di);
          const float ds1 = p1 * fmaf(__uint_as_float(d[j + 1]), keep1, -Di);
          pk[j >> 1] = pack_bf16x2(p0, p1) & m2;
          dk[j >> 1] = pack_bf16x2(ds0, ds1);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = dk[j] = 0u;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int u = (c & 1) * 4 + g;
      const uint32_t in_chunk = i * 128 + ((u ^ (i & 7)) << 4);
      *reinterpret_cast<uint4*>(sP + (c >> 1) * P_CHUNK + in_chunk) =
          make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
      *reinterpret_cast<uint4*>(sdS + (c >> 1) * (AT_ROWS * 128) + in_chunk) =
          make_uint4(dk[4 * g], dk[4 * g + 1], dk[4 * g + 2], dk[4 * g + 3]);
    }
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after_sync();
    // dQ[i][d] = sum_j dS[i][j] K[j][d] : A = dS K-major, B = K tile as [k=j][n=d] (MN-major)
    constexpr uint32_t id_q = make_idesc_bf16(128, 64, 0, 1);
    // dK[j][d] = sum_i dS[i][j] Q[i][d], dV[j][d] = sum_i P[i][j] dO[i][d]:
    //   A = dS^T / P^T = the same smem tiles read MN-major, B = Q / dO tiles MN-major
    constexpr uint32_t id_t = make_idesc_bf16(128, 64, 1, 1);
#pragma unroll
    for (int k = 0; k < AT_ROWS / 16; ++k) {
      const uint64_t ad =
          make_sw128_desc(smem_u32(sdS) + (k >> 2) * (AT_ROWS * 128) + (k & 3) * 32, 16, 1024);
      const uint64_t bd = make_sw128_desc(smem_u32(sK) + k * 2048, AT_TILE_BYTES, 1024);
      umma_f16(tmem, ad, bd, id_q, k > 0 ? 1u : 0u);
    }
#pragma unroll
    for (int k = 0; k < AT_ROWS / 16; ++k) {
      const uint64_t ad = make_sw128_desc(smem_u32(sdS) + k * 2048, AT_ROWS * 128, 1024);
      const uint64_t bd = make_sw128_desc(smem_u32(sQ) + k * 2048, AT_TILE_BYTES, 1024);
      umma_f16(tmem + 64, ad, bd, id_t, k > 0 ? 1u : 0u);
    }
#pragma unroll
    for (int k = 0; k < AT_ROWS / 16; ++k) {
      const uint64_t ad = make_sw128_desc(smem_u32(sP) + k * 2048, P_CHUNK, 1024);
      const uint64_t bd = make_sw128_desc(smem_u32(sdO) + k * 2048, AT_TILE_BYTES, 1024);
      umma_f16(tmem + 128, ad, bd, id_t, k > 0 ? 1u : 0u);
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 1);
  tc_fence_after_sync();

  // rows of dQ (query i), dK / dV (key i) -> dqkv[tok0 + i, {0, H, 2H} + head * 64 ...]: staged in the
  // (now dead) Q / K / V tiles, then written with coalesced row stores
#pragma unroll 1
  for (int part = 0; part < 3; ++part) {
    uint32_t r0[32];
    tmem_ld_32x32(t_row + part * 64 + half * 32, r0);
    tmem_ld_wait();
    stage_half_row(smem + part * AT_TILE_BYTES, i, half, r0, part == 2 ? keep_scale : a.scale);
  }
  // Bias gradient of the QKV projection = column sums of dqkv over the tile's tokens: one more
  // MMA over the staged tiles. A = [dQ | dK] (then [dV | -]) read MN-major (m = feature,
  // k = token), B = 16 identical rows of the token-validity vector (K-major) -> every column of
  // D holds the column sums, feature m in TMEM lane m: one fp32 atomic per thread replaces the
  // separate pass over dqkv (hero_colsum_bf16: 76 MB per layer).
  uint8_t* sB = sdS;
  if (dbias != nullptr) {
    {
      const int idx = threadIdx.x;                          // unit (n, chunk, u)
      const int n = idx >> 4, ch = (idx >> 3) & 1, u = idx & 7;
      const int k0 = ch * 64 + u * 8;
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        w[e] = ((k0 + 2 * e < ntok) ? 0x3F80u : 0u) | ((k0 + 2 * e + 1 < ntok) ? 0x3F800000u : 0u);
      *reinterpret_cast<uint4*>(sB + ch * 2048 + n * 128 + ((u ^ (n & 7)) << 4)) =
          make_uint4(w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (dbias != nullptr && threadIdx.x == 0) {
    tc_fence_after_sync();
    constexpr uint32_t id_c = make_idesc_bf16(128, 16, 1, 0);
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) {
#pragma unroll
      for (int k = 0; k < AT_ROWS / 16; ++k) {
        const uint64_t ad = make_sw128_desc(smem_u32(smem) + grp * 2 * AT_TILE_BYTES + k * 2048,
                                            AT_TILE_BYTES, 1024);
        const uint64_t bd = make_sw128_desc(smem_u32(sB) + (k >> 2) * 2048 + (k & 3) * 32, 16, 1024);
        umma_f16(tmem + 192 + grp * 16, ad, bd, id_c, k > 0 ? 1u : 0u);
      }
    }
    umma_commit(mma_bar);
  }
  store_tiles(smem, 3, dqkv + (long long)tok0 * (3 * a.H) + head * AT_D, 3LL * a.H, a.H, ntok);
  if (dbias != nullptr) {
    mbar_wait(mma_bar, 0);       // third phase of this barrier
    tc_fence_after_sync();
    const float cs = __uint_as_float(tmem_ld_32x1(t_row + 192 + half * 16));
    tmem_ld_wait();
    float* db = dbias + head * AT_D;
    if (half == 0) {
      if (i < AT_D) atomicAdd(db + i, cs);                 // dQ feature i
      else atomicAdd(db + a.H + (i - AT_D), cs);           // dK feature i - 64
    } else if (i < AT_D) {
      atomicAdd(db + 2 * a.H + i, cs);                     // dV feature i
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 256);
  }
}


// ------------------------------------------------------------------------------------ long rows
// Sequences longer than one 128-token tile (reference limit: max_position_embeddings = 514;
// cross-modal rows of max_txt_len subtitle tokens + matched frames, model/embed.py:33-41,
// config/train-tv*.json) are rare in HERO batches. They are handled by plain fp32 CUDA-core
// kernels, one CTA per (sequence, head), K / V (forward, dQ) or Q / dO (dK, dV) of the whole
// sequence staged in shared memory as bf16: exact same arithmetic contract as the tile kernels
// (softmax over the row's own sequence, the same dropout words, log2-domain LSE), ~20x slower per
// token, which is irrelevant at their frequency.
constexpr int AL_MAX = 768;            // tokens per long sequence (2 x 768 x 128 B = 192 KB smem)
constexpr int AL_THREADS = 256;

__device__ __forceinline__ void al_load_rows(uint8_t* dst, const __nv_bfloat16* src, long long ld,
                                             int n) {
  // n rows of 64 bf16 (128 B) -> dense [n][128 B] smem, 16-byte units
  for (int idx = threadIdx.x; idx < n * 8; idx += blockDim.x) {
    const int r = idx >> 3, u = idx & 7;
    *reinterpret_cast<uint4*>(dst + r * 128 + u * 16) =
        *reinterpret_cast<const uint4*>(src + (long long)r * ld + u * 8);
  }
}

__device__ __forceinline__ void al_row_f32(const uint8_t* row, float (&v)[64]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 q = *reinterpret_cast<const uint4*>(row + u * 16);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      v[u * 8 + 2 * j] = f.x;
      v[u * 8 + 2 * j + 1] = f.y;
    }
  }
}

__device__ __forceinline__ void al_global_row_f32(const __nv_bfloat16* p, float (&v)[64]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 q = *reinterpret_cast<const uint4*>(p + u * 8);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      v[u * 8 + 2 * j] = f.x;
      v[u * 8 + 2 * j + 1] = f.y;
    }
  }
}

__device__ __forceinline__ float al_dot_row(const float (&q)[64], const uint8_t* row) {
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 k = *reinterpret_cast<const uint4*>(row + u * 16);   // broadcast read
    const uint32_t w[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      s = fmaf(q[u * 8 + 2 * j], f.x, s);
      s = fmaf(q[u * 8 + 2 * j + 1], f.y, s);
    }
  }
  return s;
}

__device__ __forceinline__ void al_axpy_row(float (&acc)[64], float a, const uint8_t* row) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 k = *reinterpret_cast<const uint4*>(row + u * 16);
    const uint32_t w[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      acc[u * 8 + 2 * j] = fmaf(a, f.x, acc[u * 8 + 2 * j]);
      acc[u * 8 + 2 * j + 1] = fmaf(a, f.y, acc[u * 8 + 2 * j + 1]);
    }
  }
}

__device__ __forceinline__ void al_store_row(__nv_bfloat16* p, const float (&v)[64], float scale) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    uint4 q;
    q.x = pack_bf16x2(v[u * 8] * scale, v[u * 8 + 1] * scale);
    q.y = pack_bf16x2(v[u * 8 + 2] * scale, v[u * 8 + 3] * scale);
    q.z = pack_bf16x2(v[u * 8 + 4] * scale, v[u * 8 + 5] * scale);
    q.w = pack_bf16x2(v[u * 8 + 6] * scale, v[u * 8 + 7] * scale);
    *reinterpret_cast<uint4*>(p + u * 8) = q;
  }
}

// bias-gradient contribution of one stored row (the bf16-rounded values, like the tile kernels)
template <int N>
__device__ __forceinline__ void al_add_row(float* dst, const float (&v)[N]) {
#pragma unroll
  for (int d = 0; d < N; ++d) atomicAdd(dst + d, __bfloat162float(__float2bfloat16(v[d])));
}

// dropout keep factor of probability (query token tok, head, key column j RELATIVE to the tile
// start = the sequence start for a long tile): the same word layout as the tile kernels
__device__ __forceinline__ float al_keep(const AttnTcArgs& a, int tok, int head, int j) {
  if (a.drop_thr == 0u) return 1.0f;
  return attn_drop_keep(a.drop_key, a.drop_thr, tok, a.heads, head, j) ? a.drop_scale : 0.f;
}

// forward (MODE 0): ctx, lse.   dQ (MODE 1): K, V in smem, thread per query row.
template <int MODE>
__global__ void __launch_bounds__(AL_THREADS)
attn_long_q_kernel(const AttnTcArgs a, int first_tile, const __nv_bfloat16* __restrict__ qkv,
                   __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse,
                   const __nv_bfloat16* __restrict__ dctx, __nv_bfloat16* __restrict__ dqkv,
                   float* __restrict__ dbias) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_wait();
  pdl_launch_dependents();
  const int tile = first_tile + blockIdx.x, head = blockIdx.y;
  const int tok0 = a.tile_tok0[tile], n = a.tile_ntok[tile];
  uint8_t* sK = smem;
  uint8_t* sV = smem + (size_t)n * 128;
  const long long ld = 3LL * a.H;
  al_load_rows(sK, qkv + (long long)tok0 * ld + a.H + head * AT_D, ld, n);
  al_load_rows(sV, qkv + (long long)tok0 * ld + 2 * a.H + head * AT_D, ld, n);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float q[64];
    al_global_row_f32(qkv + (long long)(tok0 + i) * ld + head * AT_D, q);
    if (MODE == 0) {
      float mx = -INFINITY;
      for (int j = 0; j < n; ++j) mx = fmaxf(mx, al_dot_row(q, sK + j * 128));
      float sum = 0.f, o[64];
#pragma unroll
      for (int d = 0; d < 64; ++d) o[d] = 0.f;
      for (int j = 0; j < n; ++j) {
        const float p = ex2((al_dot_row(q, sK + j * 128) - mx) * a.scale_log2);
        sum += p;
        al_axpy_row(o, p * al_keep(a, tok0 + i, head, j), sV + j * 128);
      }
      if (lse != nullptr) lse[(long long)(tok0 + i) * a.heads + head] = mx * a.scale_log2 + log2f(sum);
      al_store_row(ctx + (long long)(tok0 + i) * a.H + head * AT_D, o, 1.0f / sum);
    } else {
      float dO[64], O[64], dq[64];
      al_global_row_f32(dctx + (long long)(tok0 + i) * a.H + head * AT_D, dO);
      al_global_row_f32(ctx + (long long)(tok0 + i) * a.H + head * AT_D, O);
      float Di = 0.f;
#pragma unroll
      for (int d = 0; d < 64; ++d) {
        Di = fmaf(dO[d], O[d], Di);
        dq[d] = 0.f;
      }
      const float row_lse = lse[(long long)(tok0 + i) * a.heads + head];
      for (int j = 0; j < n; ++j) {
        const float p = ex2(fmaf(al_dot_row(q, sK + j * 128), a.scale_log2, -row_lse));
        const float dp = al_dot_row(dO, sV + j * 128);
        const float ds = p * (dp * al_keep(a, tok0 + i, head, j) - Di) * a.scale;
        al_axpy_row(dq, ds, sK + j * 128);
      }
      al_store_row(dqkv + (long long)(tok0 + i) * ld + head * AT_D, dq, 1.0f);
      if (dbias != nullptr) al_add_row(dbias + head * AT_D, dq);
    }
  }
}

// dK, dV: Q, dO of the sequence in smem (+ per-query LSE and D); a PAIR of threads owns one key
// row, each half of the 64 features (partial dot products are summed with one shuffle).
__global__ void __launch_bounds__(AL_THREADS)
attn_long_kv_kernel(const AttnTcArgs a, int first_tile, const __nv_bfloat16* __restrict__ qkv,
                    const __nv_bfloat16* __restrict__ ctx, const float* __restrict__ lse,
                    const __nv_bfloat16* __restrict__ dctx, __nv_bfloat16* __restrict__ dqkv,
                    float* __restrict__ dbias) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_wait();
  pdl_launch_dependents();
  const int tile = first_tile + blockIdx.x, head = blockIdx.y;
  const int tok0 = a.tile_tok0[tile], n = a.tile_ntok[tile];
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + (size_t)n * 128;
  float* sL = reinterpret_cast<float*>(smem + (size_t)n * 256);
  float* sD = sL + n;
  const long long ld = 3LL * a.H;
  al_load_rows(sQ, qkv + (long long)tok0 * ld + head * AT_D, ld, n);
  al_load_rows(sdO, dctx + (long long)tok0 * a.H + head * AT_D, a.H, n);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float dO[64], O[64];
    al_row_f32(sdO + i * 128, dO);
    al_global_row_f32(ctx + (long long)(tok0 + i) * a.H + head * AT_D, O);
    float Di = 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) Di = fmaf(dO[d], O[d], Di);
    sD[i] = Di;
    sL[i] = lse[(long long)(tok0 + i) * a.heads + head];
  }
  __syncthreads();
  const int half = threadIdx.x & 1;
  for (int j0 = 0; j0 < n; j0 += AL_THREADS / 2) {
    const int j = j0 + (threadIdx.x >> 1);
    const bool ok = j < n;          // pairs stay together: the shuffle below needs both lanes
    float k[32], v[32], dk[32], dv[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) k[d] = v[d] = dk[d] = dv[d] = 0.f;
    if (ok) {
      const __nv_bfloat16* kp = qkv + (long long)(tok0 + j) * ld + a.H + head * AT_D + half * 32;
      const __nv_bfloat16* vp = kp + a.H;
#pragma unroll
      for (int d = 0; d < 32; d += 2) {
        const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(kp + d));
        const float2 y = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vp + d));
        k[d] = x.x; k[d + 1] = x.y; v[d] = y.x; v[d + 1] = y.y;
      }
    }
    for (int i = 0; i < n; ++i) {
      const uint8_t* qrow = sQ + i * 128 + half * 64;
      const uint8_t* drow = sdO + i * 128 + half * 64;
      float q[32], dO[32];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint4 a4 = *reinterpret_cast<const uint4*>(qrow + u * 16);
        const uint4 b4 = *reinterpret_cast<const uint4*>(drow + u * 16);
        const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 x = unpack_bf16x2(aw[t]), y = unpack_bf16x2(bw[t]);
          q[u * 8 + 2 * t] = x.x; q[u * 8 + 2 * t + 1] = x.y;
          dO[u * 8 + 2 * t] = y.x; dO[u * 8 + 2 * t + 1] = y.y;
        }
      }
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        s = fmaf(q[d], k[d], s);
        dp = fmaf(dO[d], v[d], dp);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      const float p = ex2(fmaf(s, a.scale_log2, -sL[i]));
      const float keep = al_keep(a, tok0 + i, head, j);
      const float pd = p * keep;
      const float ds = p * (dp * keep - sD[i]) * a.scale;
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        dv[d] = fmaf(pd, dO[d], dv[d]);
        dk[d] = fmaf(ds, q[d], dk[d]);
      }
    }
    if (ok) {
      __nv_bfloat16* dkp = dqkv + (long long)(tok0 + j) * ld + a.H + head * AT_D + half * 32;
      __nv_bfloat16* dvp = dkp + a.H;
#pragma unroll
      for (int d = 0; d < 32; d += 2) {
        *reinterpret_cast<uint32_t*>(dkp + d) = pack_bf16x2(dk[d], dk[d + 1]);
        *reinterpret_cast<uint32_t*>(dvp + d) = pack_bf16x2(dv[d], dv[d + 1]);
      }
      if (dbias != nullptr) {
        al_add_row(dbias + a.H + head * AT_D + half * 32, dk);
        al_add_row(dbias + 2 * a.H + head * AT_D + half * 32, dv);
      }
    }
  }
}

static int fill_args(AttnTcArgs* a, const int32_t* tile_tok0, const int32_t* tile_ntok,
                     const int32_t* seq_lo, const int32_t* seq_hi, int heads, int head_dim,
                     float scale, uint32_t thr, uint32_t key, float dscale) {
  HERO_REQUIRE(tile_tok0 && tile_ntok && seq_lo && seq_hi, "attn: null plan pointer");
  HERO_REQUIRE(head_dim == AT_D, "attn: head_dim must be 64 (got %d)", head_dim);
  HERO_REQUIRE(heads > 0, "attn: bad head count");
  a->tile_tok0 = tile_tok0;
  a->tile_ntok = tile_ntok;
  a->seq_lo = seq_lo;
  a->seq_hi = seq_hi;
  a->heads = heads;
  a->H = heads * head_dim;
  a->scale = scale;
  a->scale_log2 = scale * 1.4426950408889634f;
  a->drop_thr = thr;
  a->drop_key = key;
  a->drop_scale = dscale;
  return HERO_OK;
}

}  // namespace hero

using namespace hero;

static int long_smem_bytes(int max_long) { return max_long * 256 + max_long * 8 + 64; }

extern "C" int hero_attn_fwd(const void* qkv, const int32_t* tile_tok0, const int32_t* tile_ntok,
                                const int32_t* seq_lo, const int32_t* seq_hi, void* ctx, float* lse,
                                int32_t n_tok, int32_t n_tiles, int32_t n_long, int32_t max_long,
                                int32_t heads, int32_t head_dim,
                                float scale, uint32_t drop_threshold, uint32_t drop_key,
                                float drop_scale, void* stream) {
  HERO_REQUIRE(qkv && ctx, "attn_fwd: null pointer");
  HERO_REQUIRE(n_long >= 0 && n_long <= n_tiles && (n_long == 0 || (max_long > AT_ROWS && max_long <= AL_MAX)),
               "attn_fwd: bad long-sequence tiles (n_long %d, max_long %d; limit %d tokens)", n_long,
               max_long, AL_MAX);
  if (n_tiles <= 0 || n_tok <= 0) return HERO_OK;
  const int n_short = n_tiles - n_long;
  AttnTcArgs a;
  if (int rc = fill_args(&a, tile_tok0, tile_ntok, seq_lo, seq_hi, heads, head_dim, scale,
                         drop_threshold, drop_key, drop_scale))
    return rc;
  CUtensorMap tm;
  if (int rc = encode_tmap_2d_bf16(&tm, qkv, 3LL * a.H, n_tok, 3LL * a.H, AT_D, AT_ROWS)) return rc;
  const int smem = 3 * AT_TILE_BYTES + AT_MASK_BYTES + 64;
  static bool configured = false;
  if (!configured) {
    HERO_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_fwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (n_short > 0) {
    dim3 grid(n_short, heads);
    HERO_CUDA_CHECK(launch_pdl(attn_tc_fwd_kernel, grid, dim3(128), smem,
                               reinterpret_cast<cudaStream_t>(stream), tm, a,
                               reinterpret_cast<__nv_bfloat16*>(ctx), lse));
  }
  if (n_long > 0) {
    const int lsm = long_smem_bytes(max_long);
    HERO_CUDA_CHECK(cudaFuncSetAttribute(attn_long_q_kernel<0>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, lsm));
    HERO_CUDA_CHECK(launch_pdl(attn_long_q_kernel<0>, dim3(n_long, heads), dim3(AL_THREADS), lsm,
                               reinterpret_cast<cudaStream_t>(stream), a, n_short,
                               reinterpret_cast<const __nv_bfloat16*>(qkv),
                               reinterpret_cast<__nv_bfloat16*>(ctx), lse,
                               static_cast<const __nv_bfloat16*>(nullptr),
                               static_cast<__nv_bfloat16*>(nullptr), static_cast<float*>(nullptr)));
  }
  return HERO_OK;
}

extern "C" int hero_attn_bwd(const void* qkv, const int32_t* tile_tok0, const int32_t* tile_ntok,
                                const int32_t* seq_lo, const int32_t* seq_hi, const void* ctx,
                                const void* dctx, const float* lse, void* dqkv, float* dbias,
                                int32_t n_tok,
                                int32_t n_tiles, int32_t n_long, int32_t max_long,
                                int32_t heads, int32_t head_dim, float scale,
                                uint32_t drop_threshold, uint32_t drop_key, float drop_scale,
                                void* stream) {
  HERO_REQUIRE(qkv && ctx && dctx && dqkv && lse, "attn_bwd: null pointer");
  HERO_REQUIRE(n_long >= 0 && n_long <= n_tiles && (n_long == 0 || (max_long > AT_ROWS && max_long <= AL_MAX)),
               "attn_bwd: bad long-sequence tiles (n_long %d, max_long %d; limit %d tokens)", n_long,
               max_long, AL_MAX);
  if (n_tiles <= 0 || n_tok <= 0) return HERO_OK;
  const int n_short = n_tiles - n_long;
  AttnTcArgs a;
  if (int rc = fill_args(&a, tile_tok0, tile_ntok, seq_lo, seq_hi, heads, head_dim, scale,
                         drop_threshold, drop_key, drop_scale))
    return rc;
  CUtensorMap tq, td, to;
  if (int rc = encode_tmap_2d_bf16(&tq, qkv, 3LL * a.H, n_tok, 3LL * a.H, AT_D, AT_ROWS)) return rc;
  if (int rc = encode_tmap_2d_bf16(&td, dctx, a.H, n_tok, a.H, AT_D, AT_ROWS)) return rc;
  if (int rc = encode_tmap_2d_bf16(&to, ctx, a.H, n_tok, a.H, AT_D, AT_ROWS)) return rc;
  const int smem = 5 * AT_TILE_BYTES + AT_P_BYTES + 64;
  static bool configured = false;
  if (!configured) {
    HERO_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (n_short > 0) {
    dim3 grid(n_short, heads);
    HERO_CUDA_CHECK(launch_pdl(attn_tc_bwd_kernel, grid, dim3(AT_BWD_THREADS), smem,
                               reinterpret_cast<cudaStream_t>(stream), tq, td, to, a, lse,
                               reinterpret_cast<__nv_bfloat16*>(dqkv), dbias));
  }
  if (n_long > 0) {
    const int lsm = long_smem_bytes(max_long);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    HERO_CUDA_CHECK(cudaFuncSetAttribute(attn_long_q_kernel<1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, lsm));
    HERO_CUDA_CHECK(cudaFuncSetAttribute(attn_long_kv_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, lsm));
    HERO_CUDA_CHECK(launch_pdl(attn_long_q_kernel<1>, dim3(n_long, heads), dim3(AL_THREADS), lsm, st,
                               a, n_short, reinterpret_cast<const __nv_bfloat16*>(qkv),
                               const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(ctx)),
                               const_cast<float*>(lse),
                               reinterpret_cast<const __nv_bfloat16*>(dctx),
                               reinterpret_cast<__nv_bfloat16*>(dqkv), dbias));
    HERO_CUDA_CHECK(launch_pdl(attn_long_kv_kernel, dim3(n_long, heads), dim3(AL_THREADS), lsm, st, a,
                               n_short, reinterpret_cast<const __nv_bfloat16*>(qkv),
                               reinterpret_cast<const __nv_bfloat16*>(ctx), lse,
                               reinterpret_cast<const __nv_bfloat16*>(dctx),
                               reinterpret_cast<__nv_bfloat16*>(dqkv), dbias));
  }
  return HERO_OK;
}
