// Shared helpers for libdiqt_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <utility>

#include "../../include/diqt.h"

namespace diqt {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return DIQT_ECUDA;
  }
  return DIQT_OK;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// One sampler iteration is a chain of ~175 dependent kernels, many of them a few microseconds long.  Every hot-path
// kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization: it may become resident while its
// predecessor is still running, does its private prologue (barrier init, TMEM allocation, weight / bias loads) and then
// blocks in pdl_wait() until the predecessor grid has completed and flushed.  Because EVERY such kernel executes
// pdl_wait(), completion is transitive: after pdl_wait() all earlier kernels of the stream are complete.  Kernels that
// do not call pdl_wait() must be launched with plain <<<>>>.  Works under stream capture (programmatic graph edges).
bool pdl_enabled();  // DIQT_DISABLE_PDL=1 switches the attribute off (A/B measurements)

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);  // errors surface in check_launch()
}

#define DIQT_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::diqt::set_error(__VA_ARGS__);    \
      return DIQT_EINVAL;                \
    }                                    \
  } while (0)

#define DIQT_CUDA(call)                                                             \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::diqt::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
      return DIQT_ECUDA;                                                            \
    }                                                                               \
  } while (0)

// ---- 16-byte vector access over fp32 (4 lanes) or bf16 (8 lanes) ----------------------------
template <typename T>
struct Vec;

template <>
struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ float to_float(float x) { return x; }
__device__ __forceinline__ float to_float(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T>
__device__ __forceinline__ T from_float(float x);
template <>
__device__ __forceinline__ float from_float<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// Row addressing of per-volume kernels.  f <= 1: volume i owns rows [i*voxels, (i+1)*voxels).  f > 1 ("boundary" mode,
// imagen_pytorch3D.py:37-46): the f^3 sub-volumes of side h live merged in ONE volume of side f*h so that convolutions see
// their neighbours; sub-volume b = zb + f*yb + f*f*xb is block (zb, yb, xb) along (d0, d1, d2) (utils_mine.py:25-67) and its
// local voxel l = (zl*h + yl)*h + xl sits at merged row ((zb*h+zl)*f*h + yb*h+yl)*f*h + xb*h+xl.
struct SubGeom {
  int f, h;
};
__device__ __forceinline__ int64_t sub_row(const SubGeom g, int64_t voxels, int b, int64_t l) {
  if (g.f <= 1) return (int64_t)b * voxels + l;
  const int h = g.h, fh = g.f * g.h;
  const int xl = (int)(l % h), yl = (int)((l / h) % h), zl = (int)(l / ((int64_t)h * h));
  const int zb = b % g.f, yb = (b / g.f) % g.f, xb = b / (g.f * g.f);
  return ((int64_t)(zb * h + zl) * fh + (yb * h + yl)) * fh + xb * h + xl;
}

// Mish(x) = x * tanh(softplus(x)) = x * n / (n + 2),  n = e^x (e^x + 2)       (nn.Mish)
// Branch-free: the exponent is clamped at 20 (the softplus threshold of the reference op), where n/(n+2) == 1 in fp32.
// kFast uses the approximate SFU ops with flush-to-zero (2 MUFU + 7 FP32 ops per element, no range fix-up code).
template <bool kFast>
__device__ __forceinline__ float mish(float x) {
  if (kFast) {
    float t = fminf(x * 1.4426950408889634f, 28.853900817779268f), u, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(t));
    const float n = u * (u + 2.f);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n + 2.f));
    return x * (n * r);
  }
  const float u = expf(fminf(x, 20.f));
  const float n = u * (u + 2.f);
  return x * (n / (n + 2.f));
}

}  // namespace diqt
