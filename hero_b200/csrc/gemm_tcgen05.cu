// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring ->
// tcgen05.mma (fp32 accumulators in TMEM, double buffered) -> fused epilogue.
//
// One CTA per SM (or one CTA pair per TPC), 10 warps:
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one lane issues tcgen05.mma / tcgen05.commit)
//   warps 2..9  epilogue: two warps per 32-lane TMEM quadrant, splitting the tile's column slabs.
//               Per slab (32 rows x 64 bf16 columns, or 32 rows x 32 fp32 columns when the
//               residual stream is involved — always 32 x 128 B = one 4 KB swizzled smem slab):
//                 tcgen05.ld -> bias -> activation -> dropout -> residual -> pack -> smem -> TMA store
//               Everything that comes from or goes to HBM moves by TMA: the residual /
//               saved-derivative slab is prefetched one slab ahead into a warp-private smem slab
//               (its own mbarrier), outputs leave through double-buffered slabs so a store never
//               waits for the previous one to drain. No CTA-wide barrier in the epilogue: each warp
//               keeps its own bias slice.
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty x2 (MMA <-> epilogue), so the
// epilogue of tile i overlaps the mainloop of tile i+1.
//
// Replaces the cuBLAS calls behind nn.Linear in model/layers.py:125-127,176,237,251,
// model/embed.py:112 and model/layers.py:82-90 (reference), forward and backward.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace hero {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle span
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;
constexpr int SLAB_BYTES = 32 * 128;   // 32 rows x 128 B, swizzle-128B
constexpr int SMEM_LIMIT = 232448;     // 227 KB opt-in dynamic shared memory per CTA

// ACT_CE / ACT_CE_GRAD: the LM-head of the MLM task (model/layers.py:330-354 + the cross entropy of
// model/encoder.py:370-372) without materialising fp32 logits: the forward epilogue reduces every
// 64-column slab of a row to (max, sum exp) partials + the label's logit, the backward epilogue
// turns the recomputed logits into g * (softmax - onehot) in bf16.
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_GELU_GRAD = 3, ACT_CE = 4, ACT_CE_GRAD = 5 };
enum { OUT_BF16 = 0, OUT_F32_ATOMIC = 1, OUT_F32 = 2, OUT_NONE = 3 };
// RES_F32_LN: the fp32 residual is given in its pre-LayerNorm form — the epilogue applies
// (r - mean[row]) * rstd[row] * gamma[col] + beta[col] itself, so the LayerNorm kernel that feeds the
// next GEMM never has to write an fp32 copy of its output
enum { RES_NONE = 0, RES_BF16 = 1, RES_F32 = 2, RES_F32_LN = 3 };

struct GemmShape {
  int M, N, K;
  int num_m_blocks, num_n_blocks, k_splits, k_blocks;
  // 1, or 3 = split-bf16 operands: A = A_hi + A_lo, B = B_hi + B_lo, accumulate A_hi B_hi +
  // A_lo B_hi + A_hi B_lo into the same fp32 accumulator (~16 mantissa bits per operand): the
  // frame_transform pre-activation, whose ReLU gate flips on bf16 operand rounding
  int k_passes;
};

struct GemmEpilogue {
  const float* bias;
  const float* ln_mean;    // RES_F32_LN: per-row statistics and per-column affine of the residual
  const float* ln_rstd;
  const float* ln_gamma;
  const float* ln_beta;
  // ACT_CE / ACT_CE_GRAD
  const int32_t* ce_label;   // [M] target column of each row
  float2* ce_partial;        // [ceil(N / 64)][ce_ld] (max, sum exp) per (slab, row)
  float* ce_lab;             // [M] logit of the label column
  const float* ce_lse;       // [M] log-sum-exp of the row (backward)
  const float* ce_g;         // [M] upstream gradient of the row's loss (backward)
  long long ce_ld;
  int ce_n_valid;            // columns >= this are vocabulary padding: excluded / zero gradient
  int has_aux;
  float* colsum;       // OUT_BF16: column sums of the stored rows accumulate here (may be NULL)
  void* out;           // OUT_F32_ATOMIC only
  long long ld_out;
  uint32_t drop_threshold;
  uint32_t drop_key;
  float drop_scale;
};

// CTA2 = 1: the CTA is half of a pair (cluster of 2) running cta_group::2 MMAs on a 256 x BLOCK_N
// tile; it stages its own 128 rows of A and BLOCK_N / 2 columns of B per pipeline stage.
template <int BLOCK_N, int CTA2, int ACT, int OUT, int RES>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_ROWS = CTA2 ? BLOCK_N / 2 : BLOCK_N;   // B columns staged by this CTA
  static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator buffers
  // slab width in columns: 128-byte rows of the element type that travels
  static constexpr int W = (OUT == OUT_F32 || RES == RES_F32 || RES == RES_F32_LN) ? 32 : 64;
  // staging slabs per epilogue warp: output (double buffered; the GELU / ReLU kernels, which also
  // store a second tensor, use one slab per tensor), residual operand
  static constexpr int AUX = (ACT == ACT_GELU || ACT == ACT_RELU) ? 1 : 0;
  static constexpr int N_RES_BUF = RES ? 1 : 0;
  // warp-private per-column slices: bias (64 floats) + residual-LayerNorm gamma, beta (32 + 32)
  static constexpr int BIAS_BYTES = NUM_EPI_WARPS * 128 * 4;
  static constexpr int BAR_BYTES = 512;
  static constexpr int stages_with(int slabs_per_warp) {
    return (SMEM_LIMIT - NUM_EPI_WARPS * slabs_per_warp * SLAB_BYTES - BIAS_BYTES - BAR_BYTES) /
           STAGE_BYTES;
  }
  // a second output slab only where it leaves the operand pipeline at least 4 stages deep
  static constexpr int N_OUT_BUF =
      (OUT == OUT_F32_ATOMIC || OUT == OUT_NONE) ? 0
                                                 : ((!AUX && stages_with(2 + N_RES_BUF) >= 4) ? 2 : 1);
  static constexpr int SLABS_PER_WARP = N_OUT_BUF + AUX + N_RES_BUF;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * SLABS_PER_WARP * SLAB_BYTES;
  static constexpr int MAX_STAGES = stages_with(SLABS_PER_WARP);
  static constexpr int STAGES = MAX_STAGES > 6 ? 6 : MAX_STAGES;
  static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = STAGING_OFFSET + STAGING_BYTES;
  static constexpr int BAR_OFFSET = BIAS_OFFSET + BIAS_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFFSET + BAR_BYTES;
  static_assert(STAGES >= 3, "not enough shared memory for a 3-stage pipeline");
  static_assert(SMEM_BYTES <= SMEM_LIMIT, "shared memory budget exceeded");
};

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

// Write W fp32 values of this lane's row into its 128-byte row of a swizzled slab, as bf16
// (W = 64) or fp32 (W = 32): 8 16-byte chunks, chunk g stored at position g ^ (row & 7).
template <int W, bool F32>
__device__ __forceinline__ void stage_row(uint8_t* slab, int lane, const float (&v)[W]) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    uint4 u;
    if (F32) {
      u.x = __float_as_uint(v[4 * g]); u.y = __float_as_uint(v[4 * g + 1]);
      u.z = __float_as_uint(v[4 * g + 2]); u.w = __float_as_uint(v[4 * g + 3]);
    } else {
      u = pack8(&v[8 * g]);
    }
    *reinterpret_cast<uint4*>(slab + lane * 128 + ((g ^ (lane & 7)) << 4)) = u;
  }
}

template <int BLOCK_N, int A_MN, int B_MN, int ACT, int OUT, int CTA2, int RES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out,
                    const __grid_constant__ CUtensorMap tmap_aux,
                    const __grid_constant__ CUtensorMap tmap_res,
                    const __grid_constant__ CUtensorMap tmap_a_lo,
                    const __grid_constant__ CUtensorMap tmap_b_lo, const GemmShape s,
                    const GemmEpilogue e) {
  using Cfg = GemmCfg<BLOCK_N, CTA2, ACT, OUT, RES>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int W = Cfg::W;
  constexpr int NSLAB = BLOCK_N / W;
  static_assert(!(ACT == ACT_GELU || ACT == ACT_RELU || ACT == ACT_GELU_GRAD) || W == 64,
                "activation epilogues use 64-column bf16 slabs");
  static_assert(ACT != ACT_GELU_GRAD || RES == RES_BF16, "GELU' multiplier arrives as a bf16 slab");
  static_assert(OUT != OUT_F32_ATOMIC || RES == RES_NONE, "split-K accumulation takes no residual");
  static_assert((ACT == ACT_CE) == (OUT == OUT_NONE), "the CE forward epilogue stores no tile");
  // pair rank (0 = leader: issues the MMAs and owns the pipeline "full" / TMEM "empty" barriers)
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0u);

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();  // swizzle-128B atoms need a 1024 B aligned base
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFFSET);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;
  uint64_t* res_bar = bars + 2 * STAGES + 4;                     // one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + NUM_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (OUT != OUT_F32_ATOMIC && OUT != OUT_NONE) tma_prefetch_desc(&tmap_out);
    if (RES) tma_prefetch_desc(&tmap_res);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      // one arrive per epilogue warp (of both CTAs on the leader's barrier in pair mode)
      mbar_init(&tmem_empty_bar[i], NUM_EPI_WARPS * (CTA2 ? 2 : 1));
    }
    for (int i = 0; i < NUM_EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTA2) {
      tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before_sync();
  if (CTA2) cluster_sync(); else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // everything above overlapped the tail of the previous kernel (PDL); from here on we touch its
  // outputs. Let the next kernel start its own prologue right away.
  pdl_wait();
  pdl_launch_dependents();

  // persistent schedule over tiles; in pair mode both CTAs of a cluster walk the same tile list
  // (num_m_blocks then counts 256-row blocks) and the CTA's rank selects its 128-row half
  const int total_tiles = s.num_m_blocks * s.num_n_blocks * s.k_splits;
  const int first_tile = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int ks = tile % s.k_splits;
        const int t2 = tile / s.k_splits;
        const int n_blk = t2 % s.num_n_blocks;
        const int m_blk = t2 / s.num_n_blocks;
        const int kb0 = (int)(((long long)ks * s.k_blocks) / s.k_splits);
        const int kb1 = (int)(((long long)(ks + 1) * s.k_blocks) / s.k_splits);
        for (int pass = 0; pass < s.k_passes; ++pass) {
          // split-bf16 passes: (A_hi, B_hi), (A_lo, B_hi), (A_hi, B_lo)
          const CUtensorMap* ma = (pass == 1) ? &tmap_a_lo : &tmap_a;
          const CUtensorMap* mb = (pass == 2) ? &tmap_b_lo : &tmap_b;
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const uint32_t stage = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(&empty_bar[stage], ph ^ 1u);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            if (CTA2) {
              // both CTAs' loads complete on the LEADER's barrier, which expects the pair's bytes
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
              const int m128 = m_blk * 2 + (int)rank;                       // this CTA's A rows
              const int n0 = n_blk * BLOCK_N + (int)rank * (BLOCK_N / 2);   // this CTA's B columns
              if (A_MN)
                tma_load_3d_2sm(sa, ma, &full_bar[stage], 0, kb * BLOCK_K, m128 * (BLOCK_M / 64));
              else
                tma_load_2d_2sm(sa, ma, &full_bar[stage], kb * BLOCK_K, m128 * BLOCK_M);
              if (B_MN)
                tma_load_3d_2sm(sb, mb, &full_bar[stage], 0, kb * BLOCK_K, n0 / 64);
              else
                tma_load_2d_2sm(sb, mb, &full_bar[stage], kb * BLOCK_K, n0);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              if (A_MN)
                tma_load_3d(sa, ma, &full_bar[stage], 0, kb * BLOCK_K, m_blk * (BLOCK_M / 64));
              else
                tma_load_2d(sa, ma, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
              if (B_MN)
                tma_load_3d(sb, mb, &full_bar[stage], 0, kb * BLOCK_K, n_blk * (BLOCK_N / 64));
              else
                tma_load_2d(sb, mb, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (pair: leader only)
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16(CTA2 ? 2 * BLOCK_M : BLOCK_M, BLOCK_N, A_MN, B_MN);
      // K-major: 8-row groups 1024 B apart (SBO), LBO unused (encoded 1 like CUTLASS).
      // MN-major: 64-element MN chunks BLOCK_K*128 B apart (LBO), 8-k groups 1024 B apart (SBO).
      constexpr uint32_t A_LBO = A_MN ? BLOCK_K * 128 : 16;
      constexpr uint32_t B_LBO = B_MN ? BLOCK_K * 128 : 16;
      constexpr uint32_t A_KSTEP = A_MN ? UMMA_K * 128 : UMMA_K * 2;
      constexpr uint32_t B_KSTEP = B_MN ? UMMA_K * 128 : UMMA_K * 2;
      uint32_t it = 0;
      uint32_t local_tile = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++local_tile) {
        const int ks = tile % s.k_splits;
        const int kb0 = (int)(((long long)ks * s.k_blocks) / s.k_splits);
        const int kb1 = (int)(((long long)(ks + 1) * s.k_blocks) / s.k_splits);
        const uint32_t acc = local_tile & 1u;
        const uint32_t acc_ph = (local_tile >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        const int n_kb = (kb1 - kb0) * s.k_passes;
        for (int kbi = 0; kbi < n_kb; ++kbi, ++it) {
          const uint32_t stage = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(&full_bar[stage], ph);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = make_sw128_desc(sa + k * A_KSTEP, A_LBO, 1024);
            const uint64_t bdesc = make_sw128_desc(sb + k * B_KSTEP, B_LBO, 1024);
            if (CTA2)
              umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kbi > 0 || k > 0) ? 1u : 0u);
            else
              umma_f16(d_tmem, adesc, bdesc, idesc, (kbi > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if (CTA2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (CTA2) umma_commit_2sm(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue warps
    // Two warps share each TMEM lane quadrant (hardware: warp w may touch lanes 32*(w%4)..+31) and
    // split the tile's column slabs between them (slab c -> warp half c & 1).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ew = warp - 2;
    uint8_t* stage0 = smem + Cfg::STAGING_OFFSET + ew * Cfg::SLABS_PER_WARP * SLAB_BYTES;
    uint8_t* out_slab = stage0;                                            // N_OUT_BUF slabs
    uint8_t* aux_slab = stage0 + Cfg::N_OUT_BUF * SLAB_BYTES;             // AUX slab
    uint8_t* res_slab = stage0 + (Cfg::N_OUT_BUF + Cfg::AUX) * SLAB_BYTES;
    float* bias_w = reinterpret_cast<float*>(smem + Cfg::BIAS_OFFSET) + ew * 128;
    float* gamma_w = bias_w + 64;
    float* beta_w = bias_w + 96;
    uint64_t* my_res_bar = &res_bar[ew];
    const bool has_aux = Cfg::AUX && e.has_aux;
    const bool has_bias = (e.bias != nullptr);
    const int row_off = q * 32;

    // (tile, slab) sequence of this warp, one step ahead of the slab being processed: its bias
    // slice is fetched into a register pair and its residual slab is requested by TMA
    auto slab_col0 = [&](int tile, int c) {
      return ((tile / s.k_splits) % s.num_n_blocks) * BLOCK_N + c * W;
    };
    auto slab_row0 = [&](int tile) {
      const int m_blk = (tile / s.k_splits) / s.num_n_blocks;
      return (CTA2 ? m_blk * 2 + (int)rank : m_blk) * BLOCK_M + row_off;
    };
    int ntile = first_tile, nc = half;
    auto skip_invalid = [&]() {
      while (ntile < total_tiles && (nc >= NSLAB || slab_col0(ntile, nc) >= s.N)) {
        ntile += tile_step;
        nc = half;
        if (half >= NSLAB) { ntile = total_tiles; break; }
      }
    };
    auto fetch_bias = [&]() -> float2 {
      float2 b = make_float2(0.f, 0.f);
      if (has_bias && ntile < total_tiles && W == 64) {
        const int col = slab_col0(ntile, nc) + 2 * lane;
        if (col < s.N) b = __ldg(reinterpret_cast<const float2*>(e.bias + col));
      } else if (has_bias && ntile < total_tiles) {   // W == 32: one column per lane
        const int col = slab_col0(ntile, nc) + lane;
        if (col < s.N) b.x = __ldg(e.bias + col);
      }
      return b;
    };
    auto request_res = [&]() {
      if (RES && ntile < total_tiles && lane == 0) {
        mbar_arrive_expect_tx(my_res_bar, SLAB_BYTES);
        tma_load_2d(res_slab, &tmap_res, my_res_bar, slab_col0(ntile, nc), slab_row0(ntile));
      }
    };
    auto fetch_affine = [&]() -> float2 {      // (gamma, beta) of this lane's column of the next slab
      float2 gb = make_float2(0.f, 0.f);
      if (RES == RES_F32_LN && ntile < total_tiles) {
        const int col = slab_col0(ntile, nc) + lane;
        if (col < s.N) gb = make_float2(__ldg(e.ln_gamma + col), __ldg(e.ln_beta + col));
      }
      return gb;
    };
    skip_invalid();
    float2 bias_next = fetch_bias();
    float2 affine_next = fetch_affine();
    request_res();
    uint32_t res_phase = 0;
    uint32_t out_it = 0;
    uint32_t local_tile = 0;

    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++local_tile) {
      const uint32_t acc = local_tile & 1u;
      const uint32_t acc_ph = (local_tile >> 1) & 1u;
      const int row0 = slab_row0(tile);
      const int row = row0 + lane;
      const bool row_ok = row < s.M;
      int ce_label = -1;
      float ce_lse = 0.f, ce_g = 0.f;
      if ((ACT == ACT_CE || ACT == ACT_CE_GRAD) && row_ok) {
        ce_label = __ldg(e.ce_label + row);
        if (ACT == ACT_CE_GRAD) {
          ce_lse = __ldg(e.ce_lse + row);
          ce_g = __ldg(e.ce_g + row);
        }
      }
      float ln_scale = 0.f, ln_shift = 0.f;      // (r - mean) * rstd = r * rstd - mean * rstd
      if (RES == RES_F32_LN && row_ok) {
        ln_scale = __ldg(e.ln_rstd + row);
        ln_shift = -__ldg(e.ln_mean + row) * ln_scale;
      }
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(row_off) << 16);
#pragma unroll 1
      for (int c = half; c < NSLAB; c += 2) {
        const int col0 = slab_col0(tile, c);
        if (col0 >= s.N) break;  // warp-uniform
        // accumulator slab -> registers (asynchronous until tmem_ld_wait)
        uint32_t r[W / 32][32];
        tmem_ld_32x32(t_addr + c * W, r[0]);
        if (W == 64) tmem_ld_32x32(t_addr + c * W + 32, r[W / 32 - 1]);
        // this slab's bias slice -> warp-private smem; then look one slab ahead
        if (has_bias || RES == RES_F32_LN) {
          __syncwarp();     // the previous slab's broadcast reads of the slices are done
          if (W == 64) *reinterpret_cast<float2*>(bias_w + 2 * lane) = bias_next;
          else bias_w[lane] = bias_next.x;
          if (RES == RES_F32_LN) {
            gamma_w[lane] = affine_next.x;
            beta_w[lane] = affine_next.y;
          }
          __syncwarp();
        }
        nc += 2;
        skip_invalid();
        bias_next = fetch_bias();
        affine_next = fetch_affine();
        // residual / saved-derivative slab (requested one slab ago) -> registers; its smem slab is
        // then free for the next request
        uint4 opnd[8];
        if (RES) {
          mbar_wait(my_res_bar, res_phase);
          res_phase ^= 1u;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            opnd[g] = *reinterpret_cast<const uint4*>(res_slab + lane * 128 + ((g ^ (lane & 7)) << 4));
          __syncwarp();
          request_res();
        }
        tmem_ld_wait();
        float v[W];
#pragma unroll
        for (int j = 0; j < W; ++j) v[j] = __uint_as_float(r[j >> 5][j & 31]);
        if (has_bias) {
#pragma unroll
          for (int g = 0; g < W / 4; ++g) {
            const float4 b = *reinterpret_cast<const float4*>(bias_w + g * 4);
            v[4 * g] += b.x; v[4 * g + 1] += b.y; v[4 * g + 2] += b.z; v[4 * g + 3] += b.w;
          }
        }
        // activation (compile-time); the training FFN-up also emits gelu'(x) through its own slab
        if constexpr (ACT == ACT_GELU) {
          if (has_aux) {
            if (lane == 0) bulk_wait_read<1>();   // the aux slab's previous store (two groups back)
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float d[8];
#pragma unroll
              for (int j = 0; j < 8; j += 2)
                gelu_erf_with_grad2(v[8 * g + j], v[8 * g + j + 1], d[j], d[j + 1]);
              *reinterpret_cast<uint4*>(aux_slab + lane * 128 + ((g ^ (lane & 7)) << 4)) = pack8(d);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_aux, aux_slab, col0, row0);
              bulk_commit();
            }
          } else {
#pragma unroll
            for (int j = 0; j < W; ++j) v[j] = gelu_erf(v[j]);
          }
        } else if constexpr (ACT == ACT_RELU) {
          if (has_aux) {   // pre-activation copy (the ReLU backward needs its sign)
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            stage_row<W, false>(aux_slab, lane, v);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_aux, aux_slab, col0, row0);
              bulk_commit();
            }
          }
#pragma unroll
          for (int j = 0; j < W; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else if constexpr (ACT == ACT_CE) {
          // online-softmax partials of this row over the slab's valid columns
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < W; ++j)
            if (col0 + j < e.ce_n_valid) m = fmaxf(m, v[j]);
          float ssum = 0.f;
#pragma unroll
          for (int j = 0; j < W; ++j)
            if (col0 + j < e.ce_n_valid) ssum += fast_ex2((v[j] - m) * 1.4426950408889634f);
          if (row_ok) {
            e.ce_partial[(long long)(col0 / W) * e.ce_ld + row] = make_float2(m, ssum);
            const int rel = ce_label - col0;
            if (rel >= 0 && rel < W) {
              float lab = 0.f;
#pragma unroll
              for (int j = 0; j < W; ++j) lab = (j == rel) ? v[j] : lab;
              e.ce_lab[row] = lab;
            }
          }
        } else if constexpr (ACT == ACT_CE_GRAD) {
          const int rel = ce_label - col0;
#pragma unroll
          for (int j = 0; j < W; ++j) {
            float p = fast_ex2((v[j] - ce_lse) * 1.4426950408889634f);
            p = (j == rel) ? p - 1.0f : p;
            v[j] = (col0 + j < e.ce_n_valid) ? p * ce_g : 0.f;
          }
        } else if constexpr (ACT == ACT_GELU_GRAD) {   // multiply by the saved activation derivative
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint32_t pw[4] = {opnd[g].x, opnd[g].y, opnd[g].z, opnd[g].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 x = unpack_bf16x2(pw[j]);
              v[8 * g + 2 * j] *= x.x;
              v[8 * g + 2 * j + 1] *= x.y;
            }
          }
        }
        if (e.drop_threshold != 0u) {
#pragma unroll
          for (int g = 0; g < W / 8; ++g) {
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = v[8 * g + j];
            dropout_apply8(t, e.drop_key, (uint32_t)row * (uint32_t)s.N + (uint32_t)(col0 + g * 8),
                           e.drop_threshold, e.drop_scale);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 * g + j] = t[j];
          }
        }
        if constexpr (RES == RES_F32) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            v[4 * g] += __uint_as_float(opnd[g].x); v[4 * g + 1] += __uint_as_float(opnd[g].y);
            v[4 * g + 2] += __uint_as_float(opnd[g].z); v[4 * g + 3] += __uint_as_float(opnd[g].w);
          }
        } else if constexpr (RES == RES_F32_LN) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 ga = *reinterpret_cast<const float4*>(gamma_w + g * 4);
            const float4 be = *reinterpret_cast<const float4*>(beta_w + g * 4);
            v[4 * g] += fmaf(fmaf(__uint_as_float(opnd[g].x), ln_scale, ln_shift), ga.x, be.x);
            v[4 * g + 1] += fmaf(fmaf(__uint_as_float(opnd[g].y), ln_scale, ln_shift), ga.y, be.y);
            v[4 * g + 2] += fmaf(fmaf(__uint_as_float(opnd[g].z), ln_scale, ln_shift), ga.z, be.z);
            v[4 * g + 3] += fmaf(fmaf(__uint_as_float(opnd[g].w), ln_scale, ln_shift), ga.w, be.w);
          }
        } else if constexpr (RES == RES_BF16 && ACT != ACT_GELU_GRAD) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint32_t pw[4] = {opnd[g].x, opnd[g].y, opnd[g].z, opnd[g].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 x = unpack_bf16x2(pw[j]);
              v[8 * g + 2 * j] += x.x;
              v[8 * g + 2 * j + 1] += x.y;
            }
          }
        }
        if constexpr (OUT == OUT_NONE) {
          // nothing to store
        } else if constexpr (OUT == OUT_F32_ATOMIC) {
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < W / 4; ++g) {
              const int col = col0 + g * 4;
              if (col < s.N) {
                float* o = reinterpret_cast<float*>(e.out) + (long long)row * e.ld_out + col;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o),
                             "f"(v[4 * g]), "f"(v[4 * g + 1]), "f"(v[4 * g + 2]), "f"(v[4 * g + 3])
                             : "memory");
              }
            }
          }
        } else {
          // stage the slab in swizzled smem and let TMA write full 128 B rows. Output slabs are
          // double buffered: only the store issued two slabs ago must have finished READING smem
          // (bulk groups complete in order; with an aux store per slab the distance is the same).
          // (a GELU / ReLU kernel that stores no second tensor uses that tensor's slab as its
          // second output buffer: the slabs are adjacent)
          const bool two_bufs = (Cfg::N_OUT_BUF == 2) || (Cfg::AUX && !has_aux);
          uint8_t* slab = out_slab + (two_bufs ? (out_it & 1u) * SLAB_BYTES : 0);
          if (lane == 0) {
            // one group may stay in flight when it reads another slab (the other output buffer,
            // or this slab's derivative store); a lone slab must have drained
            if (two_bufs || has_aux) bulk_wait_read<1>(); else bulk_wait_read<0>();
          }
          __syncwarp();
          stage_row<W, OUT == OUT_F32>(slab, lane, v);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, slab, col0, row0);
            bulk_commit();
          }
          if constexpr (OUT == OUT_BF16 && W == 64) {
            // column sums of the slab just staged (bf16, as stored): lane l owns columns
            // 2l, 2l + 1 = word l of every 128-byte row (conflict-free), rows past M excluded
            if (e.colsum != nullptr) {
              const int rmax = min(32, s.M - row0);       // warp-uniform
              float sx = 0.f, sy = 0.f;
              const int u = lane >> 2, w4 = (lane & 3) * 4;
#pragma unroll 8
              for (int r = 0; r < rmax; ++r) {
                const float2 x = unpack_bf16x2(
                    *reinterpret_cast<const uint32_t*>(slab + r * 128 + ((u ^ (r & 7)) << 4) + w4));
                sx += x.x;
                sy += x.y;
              }
              const int col = col0 + 2 * lane;
              if (col < s.N)
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(e.colsum + col), "f"(sx),
                             "f"(sy)
                             : "memory");
            }
          }
          ++out_it;
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]);
      }
    }
    // the slabs must not go away under stores still reading them; their global writes are
    // complete when the grid is (a dependent kernel's griddepcontrol.wait covers them)
    if (OUT != OUT_F32_ATOMIC && OUT != OUT_NONE && lane == 0) bulk_wait_read<0>();
  }

  tc_fence_before_sync();
  if (CTA2) cluster_sync(); else __syncthreads();   // pair: no CTA may exit while its peer can
                                                     // still signal its barriers / read its smem
  if (warp == 1) {
    tc_fence_after_sync();
    if (CTA2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side: tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}

// Tensor maps are pure functions of (pointer, extents, leading dimension, box, kind), and the
// training loop presents the same few hundred combinations every step (weights never move;
// activation workspaces come back from the caching allocator at the same addresses): a small
// direct-mapped cache replaces ~400 cuTensorMapEncodeTiled calls per step with lookups.
enum { TM_KMAJOR = 0, TM_MNMAJOR = 1, TM_SLAB_BF16 = 2, TM_SLAB_F32 = 3 };
struct TmapKey {
  const void* ptr;
  long long ld;
  int d0, d1, box, kind;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && ld == o.ld && d0 == o.d0 && d1 == o.d1 && box == o.box && kind == o.kind;
  }
};
struct TmapEntry {
  TmapKey key;
  bool valid;
  CUtensorMap map;
};
constexpr int TMAP_CACHE_SIZE = 4096;   // entries (power of two)

static int encode_uncached(CUtensorMap* map, const TmapKey& k) {
  auto fn = get_encode_fn();
  if (!fn) return set_error(HERO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUresult r;
  if (k.kind == TM_KMAJOR) {
    // global [rows = d0, k = d1] row-major (ld elements); box = 64 k x box rows
    cuuint64_t dims[2] = {(cuuint64_t)k.d1, (cuuint64_t)k.d0};
    cuuint64_t strides[1] = {(cuuint64_t)k.ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)k.box};
    cuuint32_t estr[2] = {1, 1};
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(k.ptr), dims, strides, box,
           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (k.kind == TM_MNMAJOR) {
    // global [k = d0, mn = d1] row-major (ld elements), viewed as (64, k, mn/64) so one box
    // (64, BLOCK_K, box/64) lands in smem as consecutive [BLOCK_K x 128 B] swizzle-128B chunks
    cuuint64_t dims[3] = {64, (cuuint64_t)k.d0, (cuuint64_t)(k.d1 / 64)};
    cuuint64_t strides[2] = {(cuuint64_t)k.ld * 2, 128};
    cuuint32_t box[3] = {64, (cuuint32_t)BLOCK_K, (cuuint32_t)(k.box / 64)};
    cuuint32_t estr[3] = {1, 1, 1};
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(k.ptr), dims, strides, box,
           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    // epilogue slab over a row-major [rows = d0, cols = d1] matrix: 32 rows x 128 bytes
    const bool f32 = (k.kind == TM_SLAB_F32);
    cuuint64_t dims[2] = {(cuuint64_t)k.d1, (cuuint64_t)k.d0};
    cuuint64_t strides[1] = {(cuuint64_t)k.ld * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), 32};
    cuuint32_t estr[2] = {1, 1};
    r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
           const_cast<void*>(k.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS)
    return set_error(HERO_ERR_CUDA, "cuTensorMapEncodeTiled(kind %d, %d x %d, ld %lld) failed: %d",
                     k.kind, k.d0, k.d1, k.ld, (int)r);
  return HERO_OK;
}

static int get_tmap(CUtensorMap* out, const void* ptr, int d0, int d1, long long ld, int box,
                    int kind) {
  static thread_local std::vector<TmapEntry> cache;
  if (cache.empty()) {
    cache.resize(TMAP_CACHE_SIZE);
    for (auto& en : cache) en.valid = false;
  }
  TmapKey k;
  memset(&k, 0, sizeof(k));
  k.ptr = ptr; k.ld = ld; k.d0 = d0; k.d1 = d1; k.box = box; k.kind = kind;
  uint64_t h = reinterpret_cast<uintptr_t>(ptr) >> 4;
  h ^= (uint64_t)d0 * 0x9E3779B97F4A7C15ull;
  h ^= ((uint64_t)d1 << 21) ^ ((uint64_t)box << 7) ^ (uint64_t)kind ^ ((uint64_t)ld << 40);
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  TmapEntry& en = cache[h & (TMAP_CACHE_SIZE - 1)];
  if (!(en.valid && en.key == k)) {
    if (int rc = encode_uncached(&en.map, k)) {
      en.valid = false;
      return rc;
    }
    en.key = k;
    en.valid = true;
  }
  *out = en.map;
  return HERO_OK;
}

struct GemmMaps {
  CUtensorMap a, b, out, aux, res, a_lo, b_lo;
};

template <int BLOCK_N, int A_MN, int B_MN, int ACT, int OUT, int CTA2, int RES>
static int launch(const GemmMaps& tm, const GemmShape& s, const GemmEpilogue& e,
                  cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, CTA2, ACT, OUT, RES>;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, ACT, OUT, CTA2, RES>;
  static bool attr_set = false;
  if (!attr_set) {
    HERO_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int total = s.num_m_blocks * s.num_n_blocks * s.k_splits;
  const int sms = sm_count();
  if (sms <= 0) return set_error(HERO_ERR_NO_DEVICE, "no CUDA device");
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = serial_profiling() ? 0 : 1;
  cfg.numAttrs = 1;
  if (CTA2) {
    const int clusters = total < sms / 2 ? total : sms / 2;   // one CTA pair per TPC
    cfg.gridDim = dim3(2 * clusters);
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  } else {
    cfg.gridDim = dim3(total < sms ? total : sms);
  }
  cfg.attrs = attr;
  HERO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tm.a, tm.b, tm.out, tm.aux, tm.res, tm.a_lo, tm.b_lo, s, e));
  return HERO_OK;
}

template <int BLOCK_N, int CTA2>
static int dispatch(const hero_gemm_args* g, const GemmMaps& tm, const GemmShape& s,
                    const GemmEpilogue& e, cudaStream_t st) {
  const int layout = g->a_mn_major * 2 + g->b_mn_major;
  const bool res = g->resid != nullptr;
  if (g->out_f32_accumulate) {
    if (layout == 3) return launch<BLOCK_N, 1, 1, ACT_NONE, OUT_F32_ATOMIC, CTA2, RES_NONE>(tm, s, e, st);
    if (layout == 0) return launch<BLOCK_N, 0, 0, ACT_NONE, OUT_F32_ATOMIC, CTA2, RES_NONE>(tm, s, e, st);
    return set_error(HERO_ERR_INVALID, "fp32-accumulate supports layouts (0,0) and (1,1)");
  }
  if (g->out_f32_store) {
    // the residual stream: fp32 pre-LayerNorm sums (forward out-projection / FFN-down)
    if (layout == 0 && g->act == ACT_NONE) {
      if (res && g->resid_ln_mean)
        return launch<BLOCK_N, 0, 0, ACT_NONE, OUT_F32, CTA2, RES_F32_LN>(tm, s, e, st);
      if (res) return launch<BLOCK_N, 0, 0, ACT_NONE, OUT_F32, CTA2, RES_F32>(tm, s, e, st);
      return launch<BLOCK_N, 0, 0, ACT_NONE, OUT_F32, CTA2, RES_NONE>(tm, s, e, st);
    }
    return set_error(HERO_ERR_INVALID, "fp32-store output supports layout (0,0), act 0 only");
  }
  if (layout == 0) {
    switch (g->act) {
      case ACT_CE:
        if (!res) return launch<BLOCK_N, 0, 0, ACT_CE, OUT_NONE, CTA2, RES_NONE>(tm, s, e, st);
        break;
      case ACT_CE_GRAD:
        if (!res) return launch<BLOCK_N, 0, 0, ACT_CE_GRAD, OUT_BF16, CTA2, RES_NONE>(tm, s, e, st);
        break;
      case ACT_NONE:
        if (res) return launch<BLOCK_N, 0, 0, ACT_NONE, OUT_BF16, CTA2, RES_BF16>(tm, s, e, st);
        return launch<BLOCK_N, 0, 0, ACT_NONE, OUT_BF16, CTA2, RES_NONE>(tm, s, e, st);
      case ACT_GELU:
        if (!res) return launch<BLOCK_N, 0, 0, ACT_GELU, OUT_BF16, CTA2, RES_NONE>(tm, s, e, st);
        break;
      case ACT_RELU:
        if (res) {
          // three staging slabs per warp (output, pre-activation copy, residual): the single-CTA
          // 128 x 256 tile has no room left for a pipeline (the host picks 128-wide tiles or pairs)
          if constexpr (BLOCK_N == 256 && !CTA2)
            return set_error(HERO_ERR_INVALID, "relu + residual needs block_n 128 or CTA pairs");
          else
            return launch<BLOCK_N, 0, 0, ACT_RELU, OUT_BF16, CTA2, RES_BF16>(tm, s, e, st);
        }
        return launch<BLOCK_N, 0, 0, ACT_RELU, OUT_BF16, CTA2, RES_NONE>(tm, s, e, st);
      default: break;
    }
  } else if (layout == 1) {
    switch (g->act) {
      case ACT_NONE:
        if (res) return launch<BLOCK_N, 0, 1, ACT_NONE, OUT_BF16, CTA2, RES_BF16>(tm, s, e, st);
        return launch<BLOCK_N, 0, 1, ACT_NONE, OUT_BF16, CTA2, RES_NONE>(tm, s, e, st);
      case ACT_GELU_GRAD:
        return launch<BLOCK_N, 0, 1, ACT_GELU_GRAD, OUT_BF16, CTA2, RES_BF16>(tm, s, e, st);
      default: break;
    }
  }
  return set_error(HERO_ERR_INVALID,
                   "unsupported gemm variant: a_mn=%d b_mn=%d act=%d f32acc=%d f32store=%d resid=%d",
                   g->a_mn_major, g->b_mn_major, g->act, g->out_f32_accumulate, g->out_f32_store,
                   (int)res);
}

// Optional per-launch timing (bench.py roofline): CUDA events on the launching stream around every
// GEMM launch between hero_gemm_profile_begin() and hero_gemm_profile_end().
struct GemmProfileRec {
  int m, n, k, a_mn, b_mn, act, f32;
};
struct GemmProfile {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // pairs
  std::vector<cudaEvent_t> pool; // recycled events
  std::vector<GemmProfileRec> rec;
  double flops = 0.0;
};
static GemmProfile g_prof;

bool gemm_profile_active() { return g_prof.on; }

}  // namespace hero

extern "C" int hero_gemm_profile_begin(void) {
  using namespace hero;
  for (cudaEvent_t e : g_prof.ev) g_prof.pool.push_back(e);
  g_prof.ev.clear();
  g_prof.rec.clear();
  g_prof.flops = 0.0;
  g_prof.on = true;
  return HERO_OK;
}

extern "C" int hero_gemm_profile_end(double* ms, double* flops, int64_t* launches) {
  using namespace hero;
  g_prof.on = false;
  double total = 0.0;
  // HERO_GEMM_PROFILE_DUMP=<path>: per-launch CSV (shape, variant, ms) for tuning
  const char* path = getenv("HERO_GEMM_PROFILE_DUMP");
  FILE* f = path ? fopen(path, "w") : nullptr;
  if (f) fprintf(f, "m,n,k,a_mn,b_mn,act,f32,ms\n");
  for (size_t i = 0; i + 1 < g_prof.ev.size(); i += 2) {
    HERO_CUDA_CHECK(cudaEventSynchronize(g_prof.ev[i + 1]));
    float t = 0.f;
    HERO_CUDA_CHECK(cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]));
    total += t;
    if (f && i / 2 < g_prof.rec.size()) {
      const GemmProfileRec& r = g_prof.rec[i / 2];
      fprintf(f, "%d,%d,%d,%d,%d,%d,%d,%.5f\n", r.m, r.n, r.k, r.a_mn, r.b_mn, r.act, r.f32, t);
    }
  }
  if (f) fclose(f);
  if (ms) *ms = total;
  if (flops) *flops = g_prof.flops;
  if (launches) *launches = (int64_t)(g_prof.ev.size() / 2);
  for (cudaEvent_t e : g_prof.ev) g_prof.pool.push_back(e);
  g_prof.ev.clear();
  return HERO_OK;
}

static int hero_gemm_bf16_impl(const hero_gemm_args* g, void* stream);

extern "C" int hero_gemm_bf16(const hero_gemm_args* g, void* stream) {
  using namespace hero;
  if (!g_prof.on) return hero_gemm_bf16_impl(g, stream);
  cudaEvent_t ev[2];
  for (int i = 0; i < 2; ++i) {
    if (!g_prof.pool.empty()) {
      ev[i] = g_prof.pool.back();
      g_prof.pool.pop_back();
    } else {
      HERO_CUDA_CHECK(cudaEventCreate(&ev[i]));
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HERO_CUDA_CHECK(cudaEventRecord(ev[0], st));
  const int rc = hero_gemm_bf16_impl(g, stream);
  HERO_CUDA_CHECK(cudaEventRecord(ev[1], st));
  g_prof.ev.push_back(ev[0]);
  g_prof.ev.push_back(ev[1]);
  if (g)
    g_prof.rec.push_back(GemmProfileRec{g->m, g->n, g->k, g->a_mn_major, g->b_mn_major, g->act,
                                        g->out_f32_accumulate + 2 * g->out_f32_store});
  // (a split-bf16 GEMM issues 3x the MMAs; the profile counts its algorithmic 2MNK once)
  if (rc == HERO_OK && g) g_prof.flops += 2.0 * (double)g->m * (double)g->n * (double)g->k;
  return rc;
}

static int hero_gemm_bf16_impl(const hero_gemm_args* g, void* stream) {
  using namespace hero;
  HERO_REQUIRE(g != nullptr, "null args");
  HERO_REQUIRE(g->a && g->b && (g->out || g->act == ACT_CE), "null operand pointer");
  HERO_REQUIRE(g->m > 0 && g->n > 0 && g->k > 0, "empty gemm %dx%dx%d", g->m, g->n, g->k);
  HERO_REQUIRE(g->n % 8 == 0, "n must be a multiple of 8 (n=%d)", g->n);
  HERO_REQUIRE(g->lda % 8 == 0 && g->ldb % 8 == 0 && g->ld_out % 4 == 0, "unaligned leading dim");
  HERO_REQUIRE(g->act != ACT_GELU_GRAD || g->aux_in != nullptr, "act=3 needs aux_in");
  HERO_REQUIRE(g->act != ACT_GELU_GRAD || g->resid == nullptr, "act=3 cannot take a residual");
  HERO_REQUIRE(!(g->out_f32_accumulate && g->out_f32_store), "fp32 store and accumulate exclude each other");
  HERO_REQUIRE(!g->out_f32_accumulate || (g->act == ACT_NONE && !g->resid && !g->aux_out),
               "fp32-accumulate output supports act=0 without residual only");
  HERO_REQUIRE(!g->resid || (g->resid_f32 != 0) == (g->out_f32_store != 0),
               "an fp32 residual goes with an fp32 stored output (and a bf16 one with bf16)");
  // (an MN-major A whose row count is not a multiple of 64 is accepted when its storage is: the
  // tail chunk then reads the zero padding, and rows >= m are never written)
  if (g->a_mn_major)
    HERO_REQUIRE(g->m % 64 == 0 || g->lda >= (g->m + 63) / 64 * 64,
                 "MN-major A needs m %% 64 == 0 or lda >= round_up(m, 64) (m=%d lda=%lld)", g->m,
                 (long long)g->lda);
  if (g->b_mn_major) HERO_REQUIRE(g->n % 64 == 0, "MN-major B needs n %% 64 == 0 (n=%d)", g->n);

  int block_n = g->block_n;
  const int sms = sm_count();
  if (sms <= 0) return set_error(HERO_ERR_NO_DEVICE, "no CUDA device");
  const int m_blocks = ceil_div(g->m, BLOCK_M);
  if (block_n == 0) {
    // Prefer 256-wide tiles (less smem traffic per MAC) unless that leaves most SMs idle.
    const int t256 = m_blocks * ceil_div(g->n, 256);
    block_n = (g->n % 256 == 0 || g->n > 1024) ? 256 : 128;
    if (block_n == 256 && t256 < 2 * sms && !g->out_f32_accumulate) {
      const int t128 = m_blocks * ceil_div(g->n, 128);
      const double eff256 = (double)t256 / (ceil_div(t256, sms) * (double)sms);
      const double eff128 = (double)t128 / (ceil_div(t128, sms) * (double)sms);
      if (eff128 > eff256 * 1.15) block_n = 128;
    }
  }
  HERO_REQUIRE(block_n == 128 || block_n == 256, "block_n must be 128 or 256");
  const bool relu_res = g->act == ACT_RELU && g->resid != nullptr;
  // CTA pairs (cta_group::2, 256-row tiles) whenever the tile is 256 wide and there is more than
  // one 128-row block; cta_pair: 0 auto, 1 never, 2 force.
  // A single-CTA 128x256 tile needs more L2->SM bandwidth than the fabric delivers at the tensor
  // peak ("Why pairs"), so every GEMM with enough row blocks to fill the machine runs
  // as CTA pairs.
  const bool pair_auto = g->m > 128 &&
                         (g->out_f32_accumulate || m_blocks * ceil_div(g->n, 256) >= sms);
  bool pair = (block_n == 256) && (g->cta_pair == 2 || (g->cta_pair == 0 && pair_auto));
  // relu + residual + saved pre-activation stages three slabs per warp: no room for a pipeline in
  // the single-CTA 128 x 256 configuration -> CTA pairs (32 KB stages) when there are two row
  // blocks to pair and pairs were not ruled out, else 128-wide tiles
  if (relu_res && block_n == 256 && !pair) {
    if (g->m > 128 && g->cta_pair != 1) pair = true; else block_n = 128;
  }

  GemmShape s;
  s.M = g->m; s.N = g->n; s.K = g->k;
  s.num_m_blocks = pair ? ceil_div(g->m, 2 * BLOCK_M) : m_blocks;
  s.num_n_blocks = ceil_div(g->n, block_n);
  s.k_blocks = ceil_div(g->k, BLOCK_K);
  int k_splits = g->k_splits;
  if (!g->out_f32_accumulate) {
    k_splits = 1;
  } else if (k_splits <= 0) {
    // Split K so that the (tiles x splits) work items fill whole waves of the resident CTAs
    // (pairs): minimise  waves(s) * (k-blocks per split + epilogue)  over s. The epilogue of a
    // split-K tile (128 x BLOCK_N fp32 atomics, not overlapped with anything when a CTA owns a
    // single tile) costs about as much as 10 k-blocks of main loop. Keeps >= 4 k-blocks per
    // split so the TMA pipeline has something to overlap; ties go to the smaller split count.
    const int tiles = s.num_m_blocks * s.num_n_blocks;
    const int slots = pair ? sms / 2 : sms;   // concurrently resident tiles
    const int max_s = s.k_blocks / 4 > 1 ? s.k_blocks / 4 : 1;
    const int epi = 10;
    long long best_cost = -1;
    k_splits = 1;
    for (int sp = 1; sp <= max_s && sp <= 64; ++sp) {
      const long long waves = ceil_div(tiles * sp, slots);
      const long long cost = waves * (ceil_div(s.k_blocks, sp) + epi);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        k_splits = sp;
      }
    }
  }
  if (k_splits > s.k_blocks) k_splits = s.k_blocks;
  s.k_splits = k_splits;
  HERO_REQUIRE((g->a_lo == nullptr) == (g->b_lo == nullptr), "a_lo and b_lo go together");
  s.k_passes = g->a_lo ? 3 : 1;

  GemmEpilogue e;
  e.bias = g->bias;
  e.ln_mean = g->resid_ln_mean;
  e.ln_rstd = g->resid_ln_rstd;
  e.ln_gamma = g->resid_ln_gamma;
  e.ln_beta = g->resid_ln_beta;
  e.ce_label = g->ce_label;
  e.ce_partial = reinterpret_cast<float2*>(g->ce_partial);
  e.ce_lab = g->ce_label_logit;
  e.ce_lse = g->ce_lse;
  e.ce_g = g->ce_grad;
  e.colsum = g->out_colsum;
  if (g->out_colsum != nullptr)
    HERO_REQUIRE(!g->out_f32_accumulate && !g->out_f32_store && !g->resid_f32 && g->out != nullptr &&
                     g->n % 2 == 0 && g->act != ACT_CE,
                 "out_colsum goes with a stored bf16 output of even width");
  e.ce_ld = g->ce_ld_partial;
  e.ce_n_valid = g->ce_n_valid > 0 ? g->ce_n_valid : g->n;
  if (g->act == ACT_CE)
    HERO_REQUIRE(g->ce_label && g->ce_partial && g->ce_label_logit && g->ce_ld_partial >= g->m &&
                     !g->out_f32_accumulate && !g->out_f32_store,
                 "act 4 (cross-entropy partials) needs ce_label, ce_partial, ce_label_logit");
  if (g->act == ACT_CE_GRAD)
    HERO_REQUIRE(g->ce_label && g->ce_lse && g->ce_grad && !g->out_f32_accumulate && !g->out_f32_store,
                 "act 5 (cross-entropy gradient) needs ce_label, ce_lse, ce_grad");
  if (g->resid_ln_mean || g->resid_ln_rstd || g->resid_ln_gamma || g->resid_ln_beta)
    HERO_REQUIRE(g->resid && g->resid_f32 && g->resid_ln_mean && g->resid_ln_rstd &&
                     g->resid_ln_gamma && g->resid_ln_beta,
                 "a LayerNorm-form residual needs an fp32 resid and all four of mean/rstd/gamma/beta");
  e.has_aux = g->aux_out != nullptr;
  e.out = g->out;
  e.ld_out = g->ld_out;
  e.drop_threshold = g->drop_threshold;
  e.drop_key = g->drop_key;
  e.drop_scale = g->drop_scale;

  GemmMaps tm;
  int rc;
  if (g->a_mn_major)
    rc = get_tmap(&tm.a, g->a, g->k, (g->m + 63) / 64 * 64, g->lda, BLOCK_M, TM_MNMAJOR);
  else
    rc = get_tmap(&tm.a, g->a, g->m, g->k, g->lda, BLOCK_M, TM_KMAJOR);
  if (rc) return rc;
  const int b_box = pair ? block_n / 2 : block_n;   // a CTA of a pair stages half of the B tile
  if (g->b_mn_major)
    rc = get_tmap(&tm.b, g->b, g->k, g->n, g->ldb, b_box, TM_MNMAJOR);
  else
    rc = get_tmap(&tm.b, g->b, g->n, g->k, g->ldb, b_box, TM_KMAJOR);
  if (rc) return rc;
  tm.a_lo = tm.a;
  tm.b_lo = tm.b;
  if (g->a_lo) {     // same shapes / leading dimensions as the hi parts
    if (g->a_mn_major)
      rc = get_tmap(&tm.a_lo, g->a_lo, g->k, g->m, g->lda, BLOCK_M, TM_MNMAJOR);
    else
      rc = get_tmap(&tm.a_lo, g->a_lo, g->m, g->k, g->lda, BLOCK_M, TM_KMAJOR);
    if (rc) return rc;
    if (g->b_mn_major)
      rc = get_tmap(&tm.b_lo, g->b_lo, g->k, g->n, g->ldb, b_box, TM_MNMAJOR);
    else
      rc = get_tmap(&tm.b_lo, g->b_lo, g->n, g->k, g->ldb, b_box, TM_KMAJOR);
    if (rc) return rc;
  }
  tm.out = tm.a;
  tm.aux = tm.a;
  tm.res = tm.a;
  if (!g->out_f32_accumulate && g->act != ACT_CE) {
    if (g->out_f32_store) {
      if ((rc = get_tmap(&tm.out, g->out, g->m, g->n, g->ld_out, 32, TM_SLAB_F32))) return rc;
    } else {
      HERO_REQUIRE(g->ld_out % 8 == 0, "bf16 output needs ld_out %% 8 == 0");
      if ((rc = get_tmap(&tm.out, g->out, g->m, g->n, g->ld_out, 64, TM_SLAB_BF16))) return rc;
    }
    if (g->aux_out) {
      HERO_REQUIRE(g->ld_aux_out % 8 == 0, "aux_out needs ld %% 8 == 0");
      HERO_REQUIRE(g->act == ACT_GELU || g->act == ACT_RELU, "aux_out needs act 1 or 2");
      if ((rc = get_tmap(&tm.aux, g->aux_out, g->m, g->n, g->ld_aux_out, 64, TM_SLAB_BF16)))
        return rc;
    }
    if (g->act == ACT_GELU_GRAD) {
      HERO_REQUIRE(g->ld_aux_in % 8 == 0, "aux_in needs ld %% 8 == 0");
      if ((rc = get_tmap(&tm.res, g->aux_in, g->m, g->n, g->ld_aux_in, 64, TM_SLAB_BF16)))
        return rc;
    } else if (g->resid) {
      if (g->resid_f32) {
        HERO_REQUIRE(g->ld_resid % 4 == 0, "fp32 residual needs ld %% 4 == 0");
        if ((rc = get_tmap(&tm.res, g->resid, g->m, g->n, g->ld_resid, 32, TM_SLAB_F32))) return rc;
      } else {
        HERO_REQUIRE(g->ld_resid % 8 == 0, "bf16 residual needs ld %% 8 == 0");
        if ((rc = get_tmap(&tm.res, g->resid, g->m, g->n, g->ld_resid, 64, TM_SLAB_BF16)))
          return rc;
      }
    }
  }

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pair) return dispatch<256, 1>(g, tm, s, e, st);
  if (block_n == 256) return dispatch<256, 0>(g, tm, s, e, st);
  return dispatch<128, 0>(g, tm, s, e, st);
}
