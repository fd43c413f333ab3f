// Variable-length multi-head self-attention over packed sequences (forward + backward).
//
// HERO's sequences are short (cross-modal rows: ~10-70 tokens, temporal rows: <= 100 frames), and
// QK^T/PV are ~1-2 % of the layer FLOPs, so this path is bound by HBM/latency, not tensor
// throughput. v1: one CTA per (sequence, head); Q/K/V staged once in shared memory as bf16x2 words
// with a 33-word row stride (conflict-free for both the key-parallel score loop and the
// feature-parallel PV loop); warp-level softmax in fp32; probabilities never touch HBM (the
// backward recomputes them).
//
// Replaces model/layers.py:129-160 of the reference (BertSelfAttention.forward after the QKV
// projections) and its autograd backward.
#include "common.h"
#include "ptx.cuh"

namespace hero {

constexpr int ATT_THREADS = 128;
constexpr int ATT_WARPS = 4;
constexpr int ATT_MAX_LEN = 128;
constexpr int ROW_WORDS = 33;  // 32 bf16x2 words (64 features) + 1 pad word

// Stage `n` rows of 64 bf16 features (global row stride `ld` elements) into smem words.
__device__ __forceinline__ void stage_rows(uint32_t* dst, const __nv_bfloat16* src, long long ld,
                                           int n) {
  for (int t = threadIdx.x; t < n * 8; t += ATT_THREADS) {
    const int r = t >> 3, c = t & 7;
    const uint4 u = *reinterpret_cast<const uint4*>(src + (long long)r * ld + c * 8);
    uint32_t* d = dst + r * ROW_WORDS + c * 4;
    d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w;
  }
}

__device__ __forceinline__ float dot64(const uint32_t* a_row, const uint32_t* b_row) {
  float acc = 0.f;
#pragma unroll
  for (int w = 0; w < 32; ++w) {
    const float2 a = unpack_bf16x2(a_row[w]);
    const float2 b = unpack_bf16x2(b_row[w]);
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ cu,
                __nv_bfloat16* __restrict__ ctx, int heads, int max_len, float scale,
                uint32_t drop_thr, uint32_t drop_key, float drop_scale) {
  extern __shared__ uint32_t att_smem[];
  const int seq = blockIdx.x, head = blockIdx.y;
  const int tok0 = cu[seq];
  const int n = cu[seq + 1] - tok0;
  if (n <= 0) return;
  const int H = heads * 64;
  const long long ld = 3LL * H;
  uint32_t* sq = att_smem;
  uint32_t* sk = sq + max_len * ROW_WORDS;
  uint32_t* sv = sk + max_len * ROW_WORDS;
  float* sp = reinterpret_cast<float*>(sv + max_len * ROW_WORDS);  // [ATT_WARPS][max_len]

  const __nv_bfloat16* base = qkv + (long long)tok0 * ld + head * 64;
  stage_rows(sq, base, ld, n);
  stage_rows(sk, base + H, ld, n);
  stage_rows(sv, base + 2 * H, ld, n);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* myp = sp + warp * max_len;
  const int nchunk = (n + 31) >> 5;
  for (int i = warp; i < n; i += ATT_WARPS) {
    float s[4];
    float mx = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      s[kk] = -INFINITY;
      const int j = kk * 32 + lane;
      if (kk < nchunk && j < n) {
        s[kk] = dot64(sq + i * ROW_WORDS, sk + j * ROW_WORDS) * scale;
        mx = fmaxf(mx, s[kk]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int j = kk * 32 + lane;
      s[kk] = (kk < nchunk && j < n) ? __expf(s[kk] - mx) : 0.f;
      sum += s[kk];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int j = kk * 32 + lane;
      if (kk < nchunk && j < n) {
        float p = s[kk] * inv;
        if (drop_thr != 0u) {
          const uint32_t idx =
              (((uint32_t)seq * heads + head) * max_len + i) * (uint32_t)max_len + j;
          p = dropout_keep(drop_key, idx, drop_thr) ? p * drop_scale : 0.f;
        }
        myp[j] = p;
      }
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < n; ++j) {
      const float p = myp[j];
      const float2 v = unpack_bf16x2(sv[j * ROW_WORDS + lane]);
      o0 = fmaf(p, v.x, o0);
      o1 = fmaf(p, v.y, o1);
    }
    *reinterpret_cast<uint32_t*>(ctx + (long long)(tok0 + i) * H + head * 64 + 2 * lane) =
        pack_bf16x2(o0, o1);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ cu,
                const __nv_bfloat16* __restrict__ dctx, __nv_bfloat16* __restrict__ dqkv, int heads,
                int max_len, float scale, uint32_t drop_thr, uint32_t drop_key, float drop_scale) {
  extern __shared__ uint32_t att_smem[];
  const int seq = blockIdx.x, head = blockIdx.y;
  const int tok0 = cu[seq];
  const int n = cu[seq + 1] - tok0;
  if (n <= 0) return;
  const int H = heads * 64;
  const long long ld = 3LL * H;
  const int pstride = max_len + 1;
  uint32_t* sq = att_smem;
  uint32_t* sk = sq + max_len * ROW_WORDS;
  uint32_t* sv = sk + max_len * ROW_WORDS;
  uint32_t* sdo = sv + max_len * ROW_WORDS;
  float* sP = reinterpret_cast<float*>(sdo + max_len * ROW_WORDS);  // dropped probabilities
  float* sdS = sP + max_len * pstride;                              // d(scores) incl. 1/sqrt(d)

  const __nv_bfloat16* base = qkv + (long long)tok0 * ld + head * 64;
  stage_rows(sq, base, ld, n);
  stage_rows(sk, base + H, ld, n);
  stage_rows(sv, base + 2 * H, ld, n);
  stage_rows(sdo, dctx + (long long)tok0 * H + head * 64, H, n);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = (n + 31) >> 5;
  __nv_bfloat16* dbase = dqkv + (long long)tok0 * ld + head * 64;

  // phase 1: per query row -> P, dS rows in smem, dQ row to HBM
  for (int i = warp; i < n; i += ATT_WARPS) {
    float s[4], dp[4];
    float mx = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      s[kk] = -INFINITY;
      dp[kk] = 0.f;
      const int j = kk * 32 + lane;
      if (kk < nchunk && j < n) {
        s[kk] = dot64(sq + i * ROW_WORDS, sk + j * ROW_WORDS) * scale;
        dp[kk] = dot64(sdo + i * ROW_WORDS, sv + j * ROW_WORDS);
        mx = fmaxf(mx, s[kk]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int j = kk * 32 + lane;
      s[kk] = (kk < nchunk && j < n) ? __expf(s[kk] - mx) : 0.f;
      sum += s[kk];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float dsum = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int j = kk * 32 + lane;
      float p = s[kk] * inv;
      float keep_scale = 1.0f;
      if (drop_thr != 0u && kk < nchunk && j < n) {
        const uint32_t idx =
            (((uint32_t)seq * heads + head) * max_len + i) * (uint32_t)max_len + j;
        keep_scale = dropout_keep(drop_key, idx, drop_thr) ? drop_scale : 0.f;
      }
      s[kk] = p;                    // softmax probability
      dp[kk] = dp[kk] * keep_scale; // gradient wrt the pre-dropout probability
      dsum += p * dp[kk];
      if (kk < nchunk && j < n) sP[i * pstride + j] = p * keep_scale;
    }
    dsum = warp_sum(dsum);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int j = kk * 32 + lane;
      if (kk < nchunk && j < n) sdS[i * pstride + j] = s[kk] * (dp[kk] - dsum) * scale;
    }
    __syncwarp();
    float q0 = 0.f, q1 = 0.f;
    for (int j = 0; j < n; ++j) {
      const float ds = sdS[i * pstride + j];
      const float2 k = unpack_bf16x2(sk[j * ROW_WORDS + lane]);
      q0 = fmaf(ds, k.x, q0);
      q1 = fmaf(ds, k.y, q1);
    }
    *reinterpret_cast<uint32_t*>(dbase + (long long)i * ld + 2 * lane) = pack_bf16x2(q0, q1);
  }
  __syncthreads();
  // phase 2: per key row -> dK, dV
  for (int j = warp; j < n; j += ATT_WARPS) {
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int i = 0; i < n; ++i) {
      const float ds = sdS[i * pstride + j];
      const float p = sP[i * pstride + j];
      const float2 q = unpack_bf16x2(sq[i * ROW_WORDS + lane]);
      const float2 d = unpack_bf16x2(sdo[i * ROW_WORDS + lane]);
      k0 = fmaf(ds, q.x, k0);
      k1 = fmaf(ds, q.y, k1);
      v0 = fmaf(p, d.x, v0);
      v1 = fmaf(p, d.y, v1);
    }
    *reinterpret_cast<uint32_t*>(dbase + (long long)j * ld + H + 2 * lane) = pack_bf16x2(k0, k1);
    *reinterpret_cast<uint32_t*>(dbase + (long long)j * ld + 2 * H + 2 * lane) = pack_bf16x2(v0, v1);
  }
}

static int check_attn(const void* qkv, const int32_t* cu, const void* io, int n_seq, int max_len,
                      int heads, int head_dim) {
  HERO_REQUIRE(qkv && cu && io, "attn: null pointer");
  HERO_REQUIRE(head_dim == 64, "attn: head_dim must be 64 (got %d)", head_dim);
  HERO_REQUIRE(heads > 0 && n_seq >= 0, "attn: bad heads/n_seq");
  HERO_REQUIRE(max_len > 0 && max_len <= ATT_MAX_LEN,
               "attn: max sequence length %d exceeds the supported %d", max_len, ATT_MAX_LEN);
  return HERO_OK;
}

}  // namespace hero

using namespace hero;

extern "C" int hero_attn_fwd(const void* qkv, const int32_t* cu_seqlens, void* ctx, int32_t n_seq,
                             int32_t max_len, int32_t heads, int32_t head_dim, float scale,
                             uint32_t drop_threshold, uint32_t drop_key, float drop_scale,
                             void* stream) {
  if (int rc = check_attn(qkv, cu_seqlens, ctx, n_seq, max_len, heads, head_dim)) return rc;
  if (n_seq == 0) return HERO_OK;
  const int smem = 3 * max_len * ROW_WORDS * 4 + ATT_WARPS * max_len * 4;
  static int configured = 0;
  if (smem > configured) {
    HERO_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid(n_seq, heads);
  attn_fwd_kernel<<<grid, ATT_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), cu_seqlens, reinterpret_cast<__nv_bfloat16*>(ctx),
      heads, max_len, scale, drop_threshold, drop_key, drop_scale);
  HERO_LAUNCH_CHECK();
  return HERO_OK;
}

extern "C" int hero_attn_bwd(const void* qkv, const int32_t* cu_seqlens, const void* dctx,
                             void* dqkv, int32_t n_seq, int32_t max_len, int32_t heads,
                             int32_t head_dim, float scale, uint32_t drop_threshold,
                             uint32_t drop_key, float drop_scale, void* stream) {
  if (int rc = check_attn(qkv, cu_seqlens, dctx, n_seq, max_len, heads, head_dim)) return rc;
  HERO_REQUIRE(dqkv != nullptr, "attn_bwd: null dqkv");
  if (n_seq == 0) return HERO_OK;
  const int smem = 4 * max_len * ROW_WORDS * 4 + 2 * max_len * (max_len + 1) * 4;
  static int configured = 0;
  if (smem > configured) {
    HERO_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid(n_seq, heads);
  attn_bwd_kernel<<<grid, ATT_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), cu_seqlens,
      reinterpret_cast<const __nv_bfloat16*>(dctx), reinterpret_cast<__nv_bfloat16*>(dqkv), heads,
      max_len, scale, drop_threshold, drop_key, drop_scale);
  HERO_LAUNCH_CHECK();
  return HERO_OK;
}
