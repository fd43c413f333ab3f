// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring ->
// tcgen05.mma (fp32 accumulators in TMEM, double buffered) -> fused epilogue.
//
// One CTA per SM, 6 warps:
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one lane issues tcgen05.mma / tcgen05.commit)
//   warps 2..5  epilogue: tcgen05.ld their 32-lane TMEM quadrant, bias/act/dropout/residual,
//               bf16 store or fp32 atomic accumulate (split-K wgrad)
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty x2 (MMA <-> epilogue), so the
// epilogue of tile i overlaps the mainloop of tile i+1.
//
// Replaces the cuBLAS calls behind nn.Linear in model/layers.py:125-127,176,237,251,
// model/embed.py:112 and model/layers.py:82-90 (reference), forward and backward.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace hero {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle span
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_GELU_GRAD = 3 };

struct GemmShape {
  int M, N, K;
  int num_m_blocks, num_n_blocks, k_splits, k_blocks;
};

struct GemmEpilogue {
  const float* bias;
  const __nv_bfloat16* resid;
  long long ld_resid;
  const __nv_bfloat16* aux_in;
  long long ld_aux_in;
  __nv_bfloat16* aux_out;
  long long ld_aux_out;
  void* out;
  long long ld_out;
  uint32_t drop_threshold;
  uint32_t drop_key;
  float drop_scale;
};

// CTA2 = 1: the CTA is half of a pair (cluster of 2) running cta_group::2 MMAs on a 256 x BLOCK_N
// tile; it stages its own 128 rows of A and BLOCK_N / 2 columns of B per pipeline stage.
// SLABS = 2 (GELU kernels, which store two tensors per tile: the activation and its derivative):
// a second staging slab per epilogue warp so the two TMA stores of a slab never wait for each
// other; paid for with one pipeline stage (these are K = hidden tiles, epilogue- not load-bound).
template <int BLOCK_N, int CTA2 = 0, int SLABS = 1>
struct GemmCfg {
  static constexpr int STAGES = ((BLOCK_N == 256 && !CTA2) ? 4 : 6) - (SLABS - 1);
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_ROWS = CTA2 ? BLOCK_N / 2 : BLOCK_N;   // B columns staged by this CTA
  static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator buffers
  // epilogue staging: one slab of 32 rows x 64 bf16 (128 B rows, swizzled) per epilogue warp
  static constexpr int SLAB_BYTES = 32 * 128;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * SLAB_BYTES * SLABS;
  static constexpr int BIAS_OFFSET = STAGES * STAGE_BYTES + STAGING_BYTES;
  static constexpr int BIAS_BYTES = 2 * BLOCK_N * 4;  // per-tile bias slice, double buffered
  static constexpr int BAR_OFFSET = BIAS_OFFSET + BIAS_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 256 /*barriers*/;
};

// Epilogue arithmetic on one 64-column slab of one row (v in/out), organised in phases so that
// every warp-uniform option (bias / dropout / residual / derivative output) is tested once per slab
// and the arithmetic inside a phase is straight-line code. Operands that come from memory arrive
// already loaded: bias from the per-tile smem slice, residual / saved-derivative rows prefetched
// into registers one slab ahead.
__device__ __forceinline__ void slab_add_bias(float (&v)[64], const float* bias_s) {
#pragma unroll
  for (int g = 0; g < 16; ++g) {
    const float4 b = *reinterpret_cast<const float4*>(bias_s + g * 4);
    v[4 * g] += b.x; v[4 * g + 1] += b.y; v[4 * g + 2] += b.z; v[4 * g + 3] += b.w;
  }
}
__device__ __forceinline__ void slab_mul_bf16(float (&v)[64], const uint4 (&m)[8]) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint32_t pw[4] = {m[g].x, m[g].y, m[g].z, m[g].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16x2(pw[j]);
      v[8 * g + 2 * j] *= x.x;
      v[8 * g + 2 * j + 1] *= x.y;
    }
  }
}
__device__ __forceinline__ void slab_add_bf16(float (&v)[64], const uint4 (&m)[8]) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint32_t pw[4] = {m[g].x, m[g].y, m[g].z, m[g].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16x2(pw[j]);
      v[8 * g + 2 * j] += x.x;
      v[8 * g + 2 * j + 1] += x.y;
    }
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

template <int BLOCK_N, int A_MN, int B_MN, int ACT, int OUT_F32, int CTA2>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out,
                    const __grid_constant__ CUtensorMap tmap_aux, const GemmShape s,
                    const GemmEpilogue e) {
  using Cfg = GemmCfg<BLOCK_N, CTA2, (ACT == ACT_GELU) ? 2 : 1>;
  constexpr int STAGES = Cfg::STAGES;
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (!OUT_F32) tma_prefetch_desc(&tmap_out);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      // one arrive per epilogue warp (of both CTAs on the leader's barrier in pair mode)
      mbar_init(&tmem_empty_bar[i], NUM_EPI_WARPS * (CTA2 ? 2 : 1));
    }
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
              tma_load_3d_2sm(sa, &tmap_a, &full_bar[stage], 0, kb * BLOCK_K, m128 * (BLOCK_M / 64));
            else
              tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m128 * BLOCK_M);
            if (B_MN)
              tma_load_3d_2sm(sb, &tmap_b, &full_bar[stage], 0, kb * BLOCK_K, n0 / 64);
            else
              tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if (A_MN)
              tma_load_3d(sa, &tmap_a, &full_bar[stage], 0, kb * BLOCK_K, m_blk * (BLOCK_M / 64));
            else
              tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
            if (B_MN)
              tma_load_3d(sb, &tmap_b, &full_bar[stage], 0, kb * BLOCK_K, n_blk * (BLOCK_N / 64));
            else
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
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
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
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
              umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else
              umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
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
    // split the tile's 64-column slabs between them; 8 warps = 2 per SM sub-partition, which hides
    // the ALU/MUFU latency of the GELU / dropout / pack arithmetic.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;     // 0 or 1: which slabs of the tile this warp handles
    const int et = threadIdx.x - 64;      // 0..255 within the epilogue group
    uint8_t* slab = smem + STAGES * Cfg::STAGE_BYTES + (warp - 2) * Cfg::SLAB_BYTES;
    // second slab (GELU kernels): staging of the derivative tensor
    uint8_t* slab_aux = (ACT == ACT_GELU) ? slab + NUM_EPI_WARPS * Cfg::SLAB_BYTES : slab;
    float* bias_all = reinterpret_cast<float*>(smem + Cfg::BIAS_OFFSET);
    const bool has_aux = (e.aux_out != nullptr);
    const bool has_bias = (e.bias != nullptr);
    const bool has_resid = (ACT != ACT_GELU_GRAD) && (e.resid != nullptr);
    uint32_t slab_it = 0;
    uint32_t local_tile = 0;
    // The bias slice of tile i+1 is read (into a register) before tile i is processed, so its
    // latency never sits on the critical path; the residual / saved-derivative rows of a slab are
    // requested together with the slab's TMEM load. (Requesting them one slab ahead was measured
    // slower: +32 live registers push the kernel over the 168-register budget of a 10-warp CTA.)
    const bool has_opnd = has_resid || ACT == ACT_GELU_GRAD;
    const __nv_bfloat16* opnd_base = (ACT == ACT_GELU_GRAD) ? e.aux_in : e.resid;
    const long long ld_opnd = (ACT == ACT_GELU_GRAD) ? e.ld_aux_in : e.ld_resid;
    auto tile_bias = [&](int t) -> float {
      if (!has_bias || et >= BLOCK_N || t >= total_tiles) return 0.0f;
      const int col = ((t / s.k_splits) % s.num_n_blocks) * BLOCK_N + et;
      return (col < s.N) ? __ldg(e.bias + col) : 0.0f;
    };
    auto load_opnd = [&](uint4 (&dst)[8], int row, bool row_ok, int col0) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int col = col0 + g * 8;
        dst[g] = (row_ok && col < s.N)
                     ? *reinterpret_cast<const uint4*>(opnd_base + (long long)row * ld_opnd + col)
                     : make_uint4(0, 0, 0, 0);
      }
    };
    float bias_next = tile_bias(first_tile);
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++local_tile) {
      const int t2 = tile / s.k_splits;
      const int n_blk = t2 % s.num_n_blocks;
      const int m_blk = t2 / s.num_n_blocks;
      const uint32_t acc = local_tile & 1u;
      const uint32_t acc_ph = (local_tile >> 1) & 1u;
      float* bias_s = bias_all + acc * BLOCK_N;
      const int row0 = (CTA2 ? m_blk * 2 + (int)rank : m_blk) * BLOCK_M + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < s.M;
      if (has_bias) {
        // this tile's bias slice (fetched one tile ago) -> smem; the slot was last read two tiles
        // back, and every epilogue warp has passed this barrier once since then
        if (et < BLOCK_N) bias_s[et] = bias_next;
        bias_next = tile_bias(tile + tile_step);
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
      }
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = half; c < BLOCK_N / 64; c += 2) {
        const int col0 = n_blk * BLOCK_N + c * 64;
        if (col0 >= s.N) break;  // warp-uniform
        uint32_t r[2][32];
        tmem_ld_32x32(t_addr + c * 64, r[0]);
        tmem_ld_32x32(t_addr + c * 64 + 32, r[1]);
        uint4 opnd[8];
        if (has_opnd) load_opnd(opnd, row, row_ok, col0);
        tmem_ld_wait();
        float v[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j >> 5][j & 31]);
        if (has_bias) slab_add_bias(v, bias_s + c * 64);
        // activation (compile-time); the training FFN-up also emits gelu'(x) through its own slab
        if (ACT == ACT_GELU) {
          if (has_aux) {
            if (lane == 0) bulk_wait_read<1>();   // the aux slab's previous store (two groups back)
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float d[8];
#pragma unroll
              for (int j = 0; j < 8; j += 2)
                gelu_erf_with_grad2(v[8 * g + j], v[8 * g + j + 1], d[j], d[j + 1]);
              *reinterpret_cast<uint4*>(slab_aux + lane * 128 + ((g ^ (lane & 7)) << 4)) = pack8(d);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_aux, slab_aux, col0, row0);
              bulk_commit();
            }
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = gelu_erf(v[j]);
          }
        } else if (ACT == ACT_RELU) {
          if (has_aux) {   // pre-activation copy (the ReLU backward needs its sign)
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<uint4*>(slab_aux + lane * 128 + ((g ^ (lane & 7)) << 4)) =
                  pack8(v + 8 * g);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_aux, slab_aux, col0, row0);
              bulk_commit();
            }
          }
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else if (ACT == ACT_GELU_GRAD) {   // multiply by the saved activation derivative
          slab_mul_bf16(v, opnd);
        }
        if (e.drop_threshold != 0u) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = v[8 * g + j];
            dropout_apply8(t, e.drop_key, (uint32_t)row * (uint32_t)s.N + (uint32_t)(col0 + g * 8),
                           e.drop_threshold, e.drop_scale);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 * g + j] = t[j];
          }
        }
        if (has_resid) slab_add_bf16(v, opnd);
        if (OUT_F32) {
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 16; ++g) {
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
          // stage the 32 x 64 bf16 slab in swizzled smem and let TMA write full 128 B rows; the
          // previous store from this slab must have drained (a derivative store issued just above
          // from the second slab may stay in flight)
          uint4 outp[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) outp[g] = pack8(v + 8 * g);
          if (lane == 0) {
            if (ACT == ACT_GELU && has_aux) bulk_wait_read<1>(); else bulk_wait_read<0>();
          }
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(slab + lane * 128 + ((g ^ (lane & 7)) << 4)) = outp[g];
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, slab, col0, row0);
            bulk_commit();
          }
          ++slab_it;
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]);
      }
    }
    if (!OUT_F32 && lane == 0) bulk_wait_all();
  }

  tc_fence_before_sync();
  if (CTA2) cluster_sync(); else __syncthreads();   // pair: no CTA may exit while its peer can
                                                     // still signal its barriers / read its smem
  if (warp == 1) {
    tc_fence_after_sync();
    if (CTA2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
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

// K-major operand: global [rows, k] row-major (ld elements); box = 64 k x box_rows rows.
static int encode_kmajor(CUtensorMap* map, const void* ptr, int rows, int k, long long ld,
                         int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) return set_error(HERO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(HERO_ERR_CUDA, "cuTensorMapEncodeTiled(K-major %dx%d ld %lld) failed: %d", rows,
                     k, ld, (int)r);
  return HERO_OK;
}

// MN-major operand: global [k, mn] row-major (ld elements), viewed as (64, k, mn/64) so one box
// (64, BLOCK_K, box_mn/64) lands in smem as consecutive [BLOCK_K x 128 B] swizzle-128B chunks.
static int encode_mnmajor(CUtensorMap* map, const void* ptr, int k, int mn, long long ld,
                          int box_mn) {
  auto fn = get_encode_fn();
  if (!fn) return set_error(HERO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {64, (cuuint64_t)k, (cuuint64_t)(mn / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)BLOCK_K, (cuuint32_t)(box_mn / 64)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(HERO_ERR_CUDA, "cuTensorMapEncodeTiled(MN-major %dx%d ld %lld) failed: %d", k,
                     mn, ld, (int)r);
  return HERO_OK;
}

// bf16 row-major output [rows, cols]: box = 64 cols x 32 rows (one epilogue-warp slab).
static int encode_out(CUtensorMap* map, const void* ptr, int rows, int cols, long long ld) {
  auto fn = get_encode_fn();
  if (!fn) return set_error(HERO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(HERO_ERR_CUDA, "cuTensorMapEncodeTiled(out %dx%d ld %lld) failed: %d", rows,
                     cols, ld, (int)r);
  return HERO_OK;
}

struct GemmMaps {
  CUtensorMap a, b, out, aux;
};

template <int BLOCK_N, int A_MN, int B_MN, int ACT, int OUT_F32, int CTA2>
static int launch(const GemmMaps& tm, const GemmShape& s, const GemmEpilogue& e,
                  cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, CTA2, (ACT == ACT_GELU) ? 2 : 1>;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, ACT, OUT_F32, CTA2>;
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
  HERO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tm.a, tm.b, tm.out, tm.aux, s, e));
  return HERO_OK;
}

template <int BLOCK_N, int CTA2>
static int dispatch(const hero_gemm_args* g, const GemmMaps& tm, const GemmShape& s,
                    const GemmEpilogue& e, cudaStream_t st) {
  const int layout = g->a_mn_major * 2 + g->b_mn_major;
  if (g->out_f32_accumulate) {
    HERO_REQUIRE(g->act == ACT_NONE, "fp32-accumulate output supports act=0 only");
    if (layout == 3) return launch<BLOCK_N, 1, 1, ACT_NONE, 1, CTA2>(tm, s, e, st);
    if (layout == 0) return launch<BLOCK_N, 0, 0, ACT_NONE, 1, CTA2>(tm, s, e, st);
    return set_error(HERO_ERR_INVALID, "fp32-accumulate supports layouts (0,0) and (1,1)");
  }
  if (layout == 0) {
    switch (g->act) {
      case ACT_NONE: return launch<BLOCK_N, 0, 0, ACT_NONE, 0, CTA2>(tm, s, e, st);
      case ACT_GELU: return launch<BLOCK_N, 0, 0, ACT_GELU, 0, CTA2>(tm, s, e, st);
      case ACT_RELU: return launch<BLOCK_N, 0, 0, ACT_RELU, 0, CTA2>(tm, s, e, st);
      default: break;
    }
  } else if (layout == 1) {
    switch (g->act) {
      case ACT_NONE: return launch<BLOCK_N, 0, 1, ACT_NONE, 0, CTA2>(tm, s, e, st);
      case ACT_GELU_GRAD: return launch<BLOCK_N, 0, 1, ACT_GELU_GRAD, 0, CTA2>(tm, s, e, st);
      default: break;
    }
  }
  return set_error(HERO_ERR_INVALID, "unsupported gemm variant: a_mn=%d b_mn=%d act=%d f32=%d",
                   g->a_mn_major, g->b_mn_major, g->act, g->out_f32_accumulate);
}

// Optional per-launch timing (bench.py roofline): CUDA events on the launching stream around every
// GEMM launch between hero_gemm_profile_begin() and hero_gemm_profile_end().
struct GemmProfileRec {
  int m, n, k, a_mn, b_mn, act, f32;
};
struct GemmProfile {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // pairs
  std::vector<GemmProfileRec> rec;
  double flops = 0.0;
  const char* dump_path = nullptr;
};
static GemmProfile g_prof;

bool gemm_profile_active() { return g_prof.on; }

}  // namespace hero

extern "C" int hero_gemm_profile_begin(void) {
  using namespace hero;
  for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
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
  for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
  g_prof.ev.clear();
  return HERO_OK;
}

static int hero_gemm_bf16_impl(const hero_gemm_args* g, void* stream);

extern "C" int hero_gemm_bf16(const hero_gemm_args* g, void* stream) {
  using namespace hero;
  if (!g_prof.on) return hero_gemm_bf16_impl(g, stream);
  cudaEvent_t e0, e1;
  HERO_CUDA_CHECK(cudaEventCreate(&e0));
  HERO_CUDA_CHECK(cudaEventCreate(&e1));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HERO_CUDA_CHECK(cudaEventRecord(e0, st));
  const int rc = hero_gemm_bf16_impl(g, stream);
  HERO_CUDA_CHECK(cudaEventRecord(e1, st));
  g_prof.ev.push_back(e0);
  g_prof.ev.push_back(e1);
  if (g)
    g_prof.rec.push_back(GemmProfileRec{g->m, g->n, g->k, g->a_mn_major, g->b_mn_major, g->act,
                                        g->out_f32_accumulate});
  if (rc == HERO_OK && g) g_prof.flops += 2.0 * (double)g->m * (double)g->n * (double)g->k;
  return rc;
}

static int hero_gemm_bf16_impl(const hero_gemm_args* g, void* stream) {
  using namespace hero;
  HERO_REQUIRE(g != nullptr, "null args");
  HERO_REQUIRE(g->a && g->b && g->out, "null operand pointer");
  HERO_REQUIRE(g->m > 0 && g->n > 0 && g->k > 0, "empty gemm %dx%dx%d", g->m, g->n, g->k);
  HERO_REQUIRE(g->n % 8 == 0, "n must be a multiple of 8 (n=%d)", g->n);
  HERO_REQUIRE(g->lda % 8 == 0 && g->ldb % 8 == 0 && g->ld_out % 4 == 0, "unaligned leading dim");
  HERO_REQUIRE(g->act != ACT_GELU_GRAD || g->aux_in != nullptr, "act=3 needs aux_in");
  HERO_REQUIRE(g->act != ACT_GELU_GRAD || g->resid == nullptr, "act=3 cannot take a residual");
  if (g->a_mn_major) HERO_REQUIRE(g->m % 64 == 0, "MN-major A needs m %% 64 == 0 (m=%d)", g->m);
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
  // CTA pairs (cta_group::2, 256-row tiles) whenever the tile is 256 wide and there is more than
  // one 128-row block; cta_pair: 0 auto, 1 never, 2 force.
  // A single-CTA 128x256 tile needs more L2->SM bandwidth than the fabric delivers at the tensor
  // peak ("Why pairs"), so every GEMM with enough row blocks to fill the machine runs
  // as CTA pairs. (Until the remote accumulator-release arrive was made .relaxed, pairs lost ~10 %
  // on short-K tiles with heavy epilogues: each release compiled to MEMBAR + ERRBAR.)
  const bool pair_auto = g->m > 128 &&
                         (g->out_f32_accumulate || m_blocks * ceil_div(g->n, 256) >= sms);
  const bool pair = (block_n == 256) && (g->cta_pair == 2 || (g->cta_pair == 0 && pair_auto));

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

  GemmEpilogue e;
  e.bias = g->bias;
  e.resid = reinterpret_cast<const __nv_bfloat16*>(g->resid);
  e.ld_resid = g->ld_resid;
  e.aux_in = reinterpret_cast<const __nv_bfloat16*>(g->aux_in);
  e.ld_aux_in = g->ld_aux_in;
  e.aux_out = reinterpret_cast<__nv_bfloat16*>(g->aux_out);
  e.ld_aux_out = g->ld_aux_out;
  e.out = g->out;
  e.ld_out = g->ld_out;
  e.drop_threshold = g->drop_threshold;
  e.drop_key = g->drop_key;
  e.drop_scale = g->drop_scale;

  GemmMaps tm;
  int rc;
  if (g->a_mn_major)
    rc = encode_mnmajor(&tm.a, g->a, g->k, g->m, g->lda, BLOCK_M);
  else
    rc = encode_kmajor(&tm.a, g->a, g->m, g->k, g->lda, BLOCK_M);
  if (rc) return rc;
  const int b_box = pair ? block_n / 2 : block_n;   // a CTA of a pair stages half of the B tile
  if (g->b_mn_major)
    rc = encode_mnmajor(&tm.b, g->b, g->k, g->n, g->ldb, b_box);
  else
    rc = encode_kmajor(&tm.b, g->b, g->n, g->k, g->ldb, b_box);
  if (rc) return rc;
  if (!g->out_f32_accumulate) {
    HERO_REQUIRE(g->ld_out % 8 == 0, "bf16 output needs ld_out %% 8 == 0");
    if ((rc = encode_out(&tm.out, g->out, g->m, g->n, g->ld_out))) return rc;
    if (g->aux_out) {
      HERO_REQUIRE(g->ld_aux_out % 8 == 0, "aux_out needs ld %% 8 == 0");
      if ((rc = encode_out(&tm.aux, g->aux_out, g->m, g->n, g->ld_aux_out))) return rc;
    } else {
      tm.aux = tm.out;
    }
  } else {
    tm.out = tm.a;
    tm.aux = tm.a;
  }

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pair) return dispatch<256, 1>(g, tm, s, e, st);
  if (block_n == 256) return dispatch<256, 0>(g, tm, s, e, st);
  return dispatch<128, 0>(g, tm, s, e, st);
}
