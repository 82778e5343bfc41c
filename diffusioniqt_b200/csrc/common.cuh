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

// Build-time variants for A/B measurements (tools/gpu_ab.sh, profiles/r1m_summary.md):
//   DIQT_PDL_LATE_TRIGGER=1  (default) let the dependents launch only AFTER this kernel's own wait: at most two kernels of the chain
//                            are resident at once (K+1 becomes resident when K-1 has completed).  0: trigger at kernel entry, an
//                            unbounded chain of blocked kernels may pile up on the SMs; measured 0.5-0.7 % slower.
//   DIQT_LOAD_NC=1           Vec<T>::load through the non-coherent path (LDG.CONSTANT) instead of ld.global.cg; measured ~1 % faster,
//                            not the default because its lines are only guaranteed fresh for data that is read-only over the whole
//                            lifetime of the grid (see Vec<T>::load).
#ifndef DIQT_PDL_LATE_TRIGGER
#define DIQT_PDL_LATE_TRIGGER 1
#endif
#ifndef DIQT_LOAD_NC
#define DIQT_LOAD_NC 0
#endif
__device__ __forceinline__ void pdl_sync() {
#if DIQT_PDL_LATE_TRIGGER
  pdl_wait();
  pdl_launch_dependents();
#else
  pdl_launch_dependents();
  pdl_wait();
#endif
}

// `allow` = false launches the same kernel without the attribute (full stream serialisation): used for the first and last
// kernels of a sampler step, which sit next to launches this library does not control (graph boundaries, torch kernels).
template <bool kAllow = true, typename... KArgs, typename... Args>
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
  cfg.numAttrs = (kAllow && pdl_enabled()) ? 1 : 0;
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
// Loads go through ld.global.cg (L2 only), never the non-coherent path: a kernel launched with programmatic dependent launch is
// resident (and its SM's L1 / read-only cache alive) while its predecessors are still writing the buffers it will read after
// pdl_wait(); `const __restrict__` + plain dereference compiles to LDG.CONSTANT, whose lines are only guaranteed fresh for data that
// is read-only over the whole lifetime of the grid -- which, with PDL, starts before the producer has finished.  Streaming data has
// no L1 reuse anyway, so this costs nothing.
template <typename T>
struct Vec;

template <>
struct Vec<float> {
  static constexpr int N = 4;
  using Raw = float4;
  float v[4];
  static __device__ __forceinline__ Raw load_raw(const float* p) {
#if DIQT_LOAD_NC
    return __ldg(reinterpret_cast<const float4*>(p));
#else
    return __ldcg(reinterpret_cast<const float4*>(p));
#endif
  }
  __device__ __forceinline__ void unpack(const Raw& t) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ __forceinline__ void load(const float* p) { unpack(load_raw(p)); }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  using Raw = uint4;
  float v[8];
  static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) {
#if DIQT_LOAD_NC
    return __ldg(reinterpret_cast<const uint4*>(p));
#else
    return __ldcg(reinterpret_cast<const uint4*>(p));
#endif
  }
  __device__ __forceinline__ void unpack(const Raw& t) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { unpack(load_raw(p)); }
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

// ---- grouped channel statistics -----------------------------------------------------------------------
// Producers (convs, scale_residual, channel_stats) write one partial row [c][2] (sum, sum of squares) per CTA slot:
// partial[n][nblk][c][2].  With a StatsGroups sink the LAST CTA of every group of `gsize` consecutive rows to finish
// (atomic ticket) also sums its group's rows in index order into group[n][ngroups][c][2], ngroups <= 16.  Consumers then
// fold the GroupNorm / SE finalisation into their own prologue (ngroups short rows) instead of a separate single-CTA kernel
// over hundreds of rows.  Summation order is fixed at both levels, so results stay bitwise reproducible.
struct StatsGroups {
  float* group;           // NULL = off
  unsigned int* tickets;  // [ngroups], zero before first use; the last CTA of a group resets its ticket
  int gsize, ngroups;
};

__host__ __device__ inline int stats_group_size(int nblk, int rows_per_cta) {
  int g = (nblk + 15) / 16;
  g = (g + rows_per_cta - 1) / rows_per_cta * rows_per_cta;
  return g < rows_per_cta ? rows_per_cta : g;
}

// Called by `nthreads` threads (tid 0..nthreads-1) of a CTA that owns rows [row0, row0 + nrows) of `partial`, after those rows
// have been stored; `sync` is a barrier over exactly these threads, s_flag a shared int.
template <typename Sync>
__device__ __forceinline__ void stats_group_tail(const StatsGroups& g, const float* partial, int n, int nblk, int c, int row0,
                                                 int nrows, int tid, int nthreads, int* s_flag, Sync sync) {
  if (!g.group) return;
  __threadfence();  // this thread's partial stores are visible device-wide before the ticket is taken
  sync();
  const int grp = row0 / g.gsize;
  const int r0 = grp * g.gsize, r1 = min(nblk, r0 + g.gsize);
  if (tid == 0) {
    const unsigned members = (unsigned)((r1 - r0) / nrows);
    const unsigned t = atomicAdd(&g.tickets[grp], 1u);
    *s_flag = (t == members - 1u);
    if (t == members - 1u) g.tickets[grp] = 0u;
  }
  sync();
  if (!*s_flag) return;
  __threadfence();
  const int width = c * 2;
  for (int idx = tid; idx < n * width; idx += nthreads) {
    const int nv = idx / width, j = idx - nv * width;
    const float* src = partial + ((size_t)nv * nblk + r0) * width + j;
    // the last CTA of a group finishes after everybody else: its row loop is on the critical path of the whole grid (the in-graph
    // stamps of profiles/r3t_zm_cta_graph.txt show those CTAs ~2 us behind the median), so eight rows are in flight at once; the order
    // of the additions stays fixed
    float acc = 0.f;
    for (int r = r0; r < r1; r += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = r + u < r1 ? __ldcg(src + (size_t)u * width) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
      src += (size_t)8 * width;
    }
    g.group[((size_t)nv * g.ngroups + grp) * width + j] = acc;
  }
}

// Consumer side: total (sum, sumsq) of channel ch of volume nv over the group rows, in double.
__device__ __forceinline__ void stats_group_total(const float* __restrict__ group, int ngroups, int c, int nv, int ch, double& s, double& q) {
  const float* p = group + ((size_t)nv * ngroups) * c * 2 + ch * 2;
  s = 0.0; q = 0.0;
  for (int g = 0; g < ngroups; ++g) {
    const float2 v = __ldcg(reinterpret_cast<const float2*>(p + (size_t)g * c * 2));
    s += (double)v.x;
    q += (double)v.y;
  }
}

// Per-channel totals of volume nv from the group rows, by all `nthreads` threads of a CTA: `slices` threads per channel each sum a
// contiguous run of group rows (independent loads, one L2 round trip), then one thread per channel adds the slices.  Fixed order at
// both levels.  part: [slices][c][2] doubles with slices = max(1, nthreads / c); tot: [c][2] doubles.  Ends with a barrier.
struct SyncThreads {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
// `sync` must be a barrier over exactly the `nthreads` calling threads (default: the whole CTA)
template <typename Sync = SyncThreads>
__device__ __forceinline__ void group_channel_totals(const float* __restrict__ group, int ngroups, int c, int nv, int tid, int nthreads,
                                                     double* part, double* tot, Sync sync = Sync()) {
  const int slices = max(1, nthreads / c);
  const int per = (ngroups + slices - 1) / slices;
  const float* base = group + (size_t)nv * ngroups * c * 2;
  for (int item = tid; item < slices * c; item += nthreads) {
    const int ch = item % c, sl = item / c;
    const int g0 = sl * per, g1 = min(ngroups, g0 + per);
    double s = 0.0, q = 0.0;
    for (int g = g0; g < g1; g += 4) {
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = (g + u < g1) ? __ldcg(reinterpret_cast<const float2*>(base + ((size_t)(g + u) * c + ch) * 2)) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) { s += (double)v[u].x; q += (double)v[u].y; }
    }
    part[(sl * c + ch) * 2] = s;
    part[(sl * c + ch) * 2 + 1] = q;
  }
  sync();
  for (int ch = tid; ch < c; ch += nthreads) {
    double s = 0.0, q = 0.0;
    for (int sl = 0; sl < slices; ++sl) { s += part[(sl * c + ch) * 2]; q += part[(sl * c + ch) * 2 + 1]; }
    tot[2 * ch] = s;
    tot[2 * ch + 1] = q;
  }
  sync();
}

// GroupNorm(+FiLM) folded into y = a*x + b for every channel of volume nv, from grouped statistics, into shared a_s[c], b_s[c].
// All `nthreads` threads of the CTA call both halves.   (imagen_pytorch3D.py:546, :559-561)
struct GnParams {
  const float* group; int ngroups; long long voxels; int c, groups; float eps;
  const float* gamma; const float* beta; const float* film; int film_ld; const int* film_row; int film_row_stride_n;
};
// doubles: part[slices*c*2] + tot[c*2] + gstat[groups*2]; floats: a_s[c], b_s[c], cst[4c]
__host__ __device__ inline size_t gn_scratch_bytes(int c, int groups, int nthreads) {
  const int slices = nthreads / c > 0 ? nthreads / c : 1;
  return ((size_t)slices * c * 2 + (size_t)c * 2 + (size_t)groups * 2) * sizeof(double) + (size_t)6 * c * sizeof(float);
}
struct GnScratch {
  double *part, *tot, *gstat;
  float *a_s, *b_s, *cst;
};
__device__ __forceinline__ GnScratch gn_scratch_layout(void* smem, int c, int groups, int nthreads) {
  const int slices = max(1, nthreads / c);
  GnScratch g;
  g.part = reinterpret_cast<double*>(smem);
  g.tot = g.part + (size_t)slices * c * 2;
  g.gstat = g.tot + (size_t)c * 2;
  g.a_s = reinterpret_cast<float*>(g.gstat + (size_t)groups * 2);
  g.b_s = g.a_s + c;
  g.cst = g.b_s + c;
  return g;
}
// Half 1, BEFORE pdl_wait(): the constants (gamma, beta, FiLM row of this sampler step) go to shared memory while the producer kernel
// is still draining.  Legal ahead of the wait: they were written before the first kernel of this forward, which is never launched with
// the PDL attribute (engine: init_conv / edm_prepare), so it started only after they were complete.
__device__ __forceinline__ void gn_prefetch_constants(const GnParams& gp, int nv, int tid, int nthreads, const GnScratch& sc) {
  const float* fr = nullptr;
  if (gp.film) fr = gp.film + (long long)((gp.film_row ? __ldcg(gp.film_row) : 0) + nv * gp.film_row_stride_n) * gp.film_ld;
  for (int ch = tid; ch < gp.c; ch += nthreads) {
    sc.cst[ch] = gp.gamma[ch];
    sc.cst[gp.c + ch] = gp.beta[ch];
    sc.cst[2 * gp.c + ch] = fr ? fr[ch] + 1.f : 1.f;  // x * (scale + 1) + shift
    sc.cst[3 * gp.c + ch] = fr ? fr[gp.c + ch] : 0.f;
  }
}
// Half 2, after pdl_wait(): one L2 round trip for the group rows, then shared-memory arithmetic.  This sits on the critical path of
// every consumer (the fused conv cannot normalise its first plane before it: profiles/r2v_zm_timeline.txt showed 3.5 us from grid
// dependency to first plane with the round-1 version: five barriers, a separate pass per reduction level), so it is two phases now:
//   1. one thread per channel sums that channel's <= 16 group rows (independent loads: one round trip) in double;
//   2. one thread per channel adds the totals of its GroupNorm group (redundantly per channel: a handful of shared-memory reads) and
//      forms (a, b).
template <typename Sync = SyncThreads>
__device__ __forceinline__ void gn_affine_from_groups(const GnParams& gp, int nv, int tid, int nthreads, const GnScratch& sc, Sync sync = Sync()) {
  const float* base = gp.group + (size_t)nv * gp.ngroups * gp.c * 2;
  for (int ch = tid; ch < gp.c; ch += nthreads) {
    double s = 0.0, q = 0.0;
    for (int g = 0; g < gp.ngroups; g += 8) {
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = (g + u < gp.ngroups) ? __ldcg(reinterpret_cast<const float2*>(base + ((size_t)(g + u) * gp.c + ch) * 2)) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) { s += (double)v[u].x; q += (double)v[u].y; }  // fixed order
    }
    sc.tot[2 * ch] = s;
    sc.tot[2 * ch + 1] = q;
  }
  sync();
  const int cpg = gp.c / gp.groups;
  for (int ch = tid; ch < gp.c; ch += nthreads) {
    const int g = ch / cpg;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < cpg; ++i) { s += sc.tot[(g * cpg + i) * 2]; q += sc.tot[(g * cpg + i) * 2 + 1]; }
    const double cnt = (double)gp.voxels * cpg, mean_d = s / cnt;
    double var = q / cnt - mean_d * mean_d;
    if (var < 0.0) var = 0.0;
    const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var + (double)gp.eps));
    float a = rstd * sc.cst[ch];
    float b = sc.cst[gp.c + ch] - mean * a;
    const float fs = sc.cst[2 * gp.c + ch], fh = sc.cst[3 * gp.c + ch];
    a *= fs;
    b = fmaf(b, fs, fh);
    sc.a_s[ch] = a;
    sc.b_s[ch] = b;
  }
  sync();
}

// Mish(x) = x * tanh(softplus(x)) = x * n / (n + 2),  n = e^x (e^x + 2)       (nn.Mish)
// Branch-free: the exponent is clamped at 20 (the softplus threshold of the reference op), where n/(n+2) == 1 in fp32.
// kFast uses the approximate SFU ops with flush-to-zero (2 MUFU + 7 FP32 ops per element, no range fix-up code).
//   DIQT_MISH_V2=1   kFast: x - 2x / (u^2 + 2u + 2), u = e^x: two instructions less per element and no clamp (u = inf gives 1/inf = 0,
//                    i.e. mish = x); loses RELATIVE accuracy in the far negative tail (|mish| < 1e-4), which bf16 storage cannot see.
#ifndef DIQT_MISH_V2
#define DIQT_MISH_V2 1
#endif
template <bool kFast>
__device__ __forceinline__ float mish(float x) {
#if DIQT_MISH_V2
  if (kFast) {
    float u, w;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(x * 1.4426950408889634f));
    const float d = fmaf(u, u + 2.f, 2.f);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(d));
    return fmaf(x * w, -2.f, x);
  }
#endif
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

// mish<true> on two values at once: the non-MUFU arithmetic as packed fp32x2 instructions (FMUL2 / FADD2 / FFMA2, sm_100); per lane the same
// operations in the same order, so the results are bit-identical to mish<true> (DIQT_MISH_V2 form)
__device__ __forceinline__ float2 mish2_fast(float2 x) {
#if DIQT_MISH_V2
  const float2 t = __fmul2_rn(x, make_float2(1.4426950408889634f, 1.4426950408889634f));
  float2 u, w;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u.x) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u.y) : "f"(t.y));
  const float2 d = __ffma2_rn(u, __fadd2_rn(u, make_float2(2.f, 2.f)), make_float2(2.f, 2.f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(w.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(w.y) : "f"(d.y));
  return __ffma2_rn(__fmul2_rn(x, w), make_float2(-2.f, -2.f), x);
#else
  return make_float2(mish<true>(x.x), mish<true>(x.y));
#endif
}

}  // namespace diqt
