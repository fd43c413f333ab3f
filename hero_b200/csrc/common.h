// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/hero_b200.h"

namespace hero {

// Thread-local error text returned by hero_last_error().
char* error_buffer();
int set_error(int status, const char* fmt, ...);
int sm_count();

#define HERO_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::hero::set_error(HERO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,              \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                \
  } while (0)

#define HERO_REQUIRE(cond, ...)                                           \
  do {                                                                    \
    if (!(cond)) return ::hero::set_error(HERO_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define HERO_LAUNCH_CHECK()                                                                  \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return ::hero::set_error(HERO_ERR_CUDA, "kernel launch failed: %s (%s:%d)",            \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// 2-D bf16 tensor map over a row-major [outer, inner] matrix (leading dim `ld` elements) with a
// 128-byte-swizzled box of box_inner (<= 64) x box_outer elements. `map` is a CUtensorMap.
int encode_tmap_2d_bf16(void* map, const void* ptr, long long inner, long long outer, long long ld,
                        int box_inner, int box_outer);

}  // namespace hero
