// Bandwidth-bound kernels of the ResnetBlock: channel statistics, GroupNorm(+FiLM)+Mish,
// squeeze-excitation gate, gated residual add.  All operate on channels-last activations
// [n][voxels][c] with row pitch ld, 16-byte vector accesses, fp32 math.
//
// Reference call sites replaced: nn.GroupNorm (imagen_pytorch3D.py:546, 557), FiLM (:559-561),
// nn.Mish (:547, 563), SE3D (:617-632), residual add (:612).
#include <algorithm>

#include <stdlib.h>

#include "tc_common.cuh"

namespace diqt {

// squeeze-excitation gate computed in the consumer's prologue from grouped statistics (group == NULL: read `gate` instead)
struct SeParams {
  const float* group;
  int ngroups, hidden;
  const float* w1;
  const float* w2;
  int stage_w;  // 1: W1 / W2 are staged in shared memory ahead of the PDL wait (they fit), 0: read from global
};

// thread mapping shared by every kernel in this file:
//   nvec = c / VEC vectors per voxel row; thread -> (col = tid % nvec, lane = tid / nvec);
//   a block owns voxels [blk*vpb, (blk+1)*vpb) of volume blockIdx.y and strides them by `lanes`.
struct RowMap {
  int nvec, lanes, threads;
};

static inline RowMap make_rowmap(int c, int vec, int max_threads) {
  RowMap m;
  m.nvec = c / vec;
  m.lanes = max_threads / m.nvec;
  if (m.lanes < 1) m.lanes = 1;
  m.threads = m.nvec * m.lanes;
  return m;
}

// ---- per-block reduction of per-thread channel sums -> partial[n][blk][c][2] ------------------
template <int VEC>
__device__ __forceinline__ void block_channel_reduce(const float (&s)[VEC], const float (&q)[VEC], int col, int lane,
                                                     int lanes, int c, float* __restrict__ partial_out,
                                                     float* smem /* 2 * lanes * c floats */) {
  float* ss = smem;
  float* sq = smem + (size_t)lanes * c;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    ss[lane * c + col * VEC + i] = s[i];
    sq[lane * c + col * VEC + i] = q[i];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) {  // fixed order: deterministic
      a += ss[l * c + ch];
      b += sq[l * c + ch];
    }
    partial_out[2 * ch] = a;
    partial_out[2 * ch + 1] = b;
  }
}

template <typename T>
__global__ void channel_stats_kernel(const T* __restrict__ x, int64_t voxels, int c, int ld, int nvec, int lanes,
                                     int64_t vpb, float* __restrict__ partial, SubGeom sg, StatsGroups og) {
  constexpr int VEC = Vec<T>::N;
  extern __shared__ float smem[];
  const int col = threadIdx.x % nvec, lane = threadIdx.x / nvec;
  const int blk = blockIdx.x, n = blockIdx.y, nblk = gridDim.x;
  const int64_t v0 = (int64_t)blk * vpb;
  const int64_t v1 = min(voxels, v0 + vpb);
  pdl_sync();
  const T* base = x + col * VEC;
  float s[VEC], q[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) s[i] = q[i] = 0.f;
  int64_t v = v0 + lane;
  for (; v + 3 * (int64_t)lanes < v1; v += 4 * (int64_t)lanes) {
    Vec<T> r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u].load(base + sub_row(sg, voxels, n, v + (int64_t)u * lanes) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        s[i] += r[u].v[i];
        q[i] = fmaf(r[u].v[i], r[u].v[i], q[i]);
      }
  }
  for (; v < v1; v += lanes) {
    Vec<T> r;
    r.load(base + sub_row(sg, voxels, n, v) * ld);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      s[i] += r.v[i];
      q[i] = fmaf(r.v[i], r.v[i], q[i]);
    }
  }
  block_channel_reduce<VEC>(s, q, col, lane, lanes, c, partial + ((int64_t)n * nblk + blk) * c * 2, smem);
  // grouped sink (single-volume launches only: the group row of volume n is built from this volume's partial rows)
  StatsGroups ov = og;  // this volume's slice of the sink (tickets[n][ngroups], group[n][ngroups][c][2])
  if (ov.group) { ov.group += (int64_t)n * og.ngroups * c * 2; ov.tickets += n * og.ngroups; }
  stats_group_tail(ov, partial + (int64_t)n * nblk * c * 2, 1, nblk, c, blk, 1, threadIdx.x, blockDim.x, reinterpret_cast<int*>(smem),
                   [] { __syncthreads(); });
}

// ---- GroupNorm finalize: partial sums -> per-(n,c) affine (a, b), with FiLM folded in -----------
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int nblk, int64_t voxels, int c, int groups,
                                   float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ film, int film_ld, const int32_t* __restrict__ film_row,
                                   int film_row_stride_n, float* __restrict__ a_out, float* __restrict__ b_out) {
  extern __shared__ double sm[];  // acc[parts][c][2], tot[c][2], gstat[groups][2]
  pdl_sync();
  const int n = blockIdx.x;
  const int parts = max(1, (int)blockDim.x / c);
  double* acc = sm;
  double* tot = sm + (size_t)parts * c * 2;
  double* gstat = tot + (size_t)c * 2;
  const float* p = partial + (int64_t)n * nblk * c * 2;
  for (int idx = threadIdx.x; idx < parts * c; idx += blockDim.x) {
    const int ch = idx % c, part = idx / c;
    double s = 0.0, q = 0.0;
    int b = part;
    for (; b + 3 * parts < nblk; b += 4 * parts) {  // four independent loads in flight; summation order stays fixed
      const float2 v0 = __ldcg(reinterpret_cast<const float2*>(p + ((int64_t)b * c + ch) * 2));
      const float2 v1 = __ldcg(reinterpret_cast<const float2*>(p + ((int64_t)(b + parts) * c + ch) * 2));
      const float2 v2 = __ldcg(reinterpret_cast<const float2*>(p + ((int64_t)(b + 2 * parts) * c + ch) * 2));
      const float2 v3 = __ldcg(reinterpret_cast<const float2*>(p + ((int64_t)(b + 3 * parts) * c + ch) * 2));
      s += (double)v0.x; q += (double)v0.y;
      s += (double)v1.x; q += (double)v1.y;
      s += (double)v2.x; q += (double)v2.y;
      s += (double)v3.x; q += (double)v3.y;
    }
    for (; b < nblk; b += parts) {
      const float2 v = __ldcg(reinterpret_cast<const float2*>(p + ((int64_t)b * c + ch) * 2));
      s += (double)v.x; q += (double)v.y;
    }
    acc[(part * c + ch) * 2] = s;
    acc[(part * c + ch) * 2 + 1] = q;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int part = 0; part < parts; ++part) {
      s += acc[(part * c + ch) * 2];
      q += acc[(part * c + ch) * 2 + 1];
    }
    tot[ch * 2] = s;
    tot[ch * 2 + 1] = q;
  }
  __syncthreads();
  const int cpg = c / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < cpg; ++i) {
      s += tot[(g * cpg + i) * 2];
      q += tot[(g * cpg + i) * 2 + 1];
    }
    const double cnt = (double)voxels * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    gstat[g * 2] = mean;
    gstat[g * 2 + 1] = 1.0 / sqrt(var + (double)eps);
  }
  __syncthreads();
  const float* fr = nullptr;
  if (film) {
    const int row = (film_row ? __ldcg(film_row) : 0) + n * film_row_stride_n;
    fr = film + (int64_t)row * film_ld;
  }
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const int g = ch / cpg;
    const float mean = (float)gstat[g * 2], rstd = (float)gstat[g * 2 + 1];
    float a = rstd * gamma[ch];
    float b = beta[ch] - mean * a;
    if (fr) {  // x * (scale + 1) + shift   (imagen_pytorch3D.py:559-561)
      const float sc = fr[ch] + 1.f, sh = fr[c + ch];
      a *= sc;
      b = fmaf(b, sc, sh);
    }
    a_out[(int64_t)n * c + ch] = a;
    b_out[(int64_t)n * c + ch] = b;
  }
}

// ---- y = mish(a * x + b) -----------------------------------------------------------------------
template <typename T, bool kFast>
__global__ void __launch_bounds__(256, 4) affine_mish_kernel(const T* __restrict__ x, int ld_x, T* __restrict__ y, int ld_y, int64_t voxels,
                                                             int c, int nvec, int lanes, int64_t vpb, const float* __restrict__ a,
                                                             const float* __restrict__ b, SubGeom sg, GnParams gp) {
  constexpr int VEC = Vec<T>::N;
  using Raw = typename Vec<T>::Raw;
  extern __shared__ double gn_scratch[];  // grouped mode: gn_scratch_bytes()
  const int col = threadIdx.x % nvec, lane = threadIdx.x / nvec;
  const int n = blockIdx.y;
  const int64_t v0 = (int64_t)blockIdx.x * vpb;
  const int64_t v1 = min(voxels, v0 + vpb);
  float av[VEC], bv[VEC];
  if (gp.group) {  // GroupNorm (+FiLM) finalised here from the producer's grouped statistics: no separate finalize kernel
    const GnScratch sc = gn_scratch_layout(gn_scratch, c, gp.groups, blockDim.x);
    gn_prefetch_constants(gp, n, threadIdx.x, blockDim.x, sc);  // overlaps the tail of the producer kernel
    pdl_sync();
    gn_affine_from_groups(gp, n, threadIdx.x, blockDim.x, sc);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      av[i] = sc.a_s[col * VEC + i];
      bv[i] = sc.b_s[col * VEC + i];
    }
  } else {
    pdl_sync();
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      av[i] = __ldcg(a + (int64_t)n * c + col * VEC + i);  // written by the previous kernel (gn_finalize): L2 only, see Vec<T>::load
      bv[i] = __ldcg(b + (int64_t)n * c + col * VEC + i);
    }
  }
  const T* xb = x + col * VEC;
  T* yb = y + col * VEC;
  int64_t v = v0 + lane;
  // four 16-byte loads in flight per thread, kept packed until they are used (register budget: 4 CTAs of 256 threads per SM)
  for (; v + 3 * (int64_t)lanes < v1; v += 4 * (int64_t)lanes) {
    Raw raw[4];
    int64_t row[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      row[u] = sub_row(sg, voxels, n, v + (int64_t)u * lanes);
      raw[u] = Vec<T>::load_raw(xb + row[u] * ld_x);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      Vec<T> r;
      r.unpack(raw[u]);
#pragma unroll
      for (int i = 0; i < VEC; i += 2) {
        if (kFast) {   // channel pairs per FFMA2 / FMUL2 / FADD2: the same IEEE operations per lane (mish2_fast)
          const float2 o = mish2_fast(__ffma2_rn(make_float2(av[i], av[i + 1]), make_float2(r.v[i], r.v[i + 1]), make_float2(bv[i], bv[i + 1])));
          r.v[i] = o.x; r.v[i + 1] = o.y;
        } else {
          r.v[i] = mish<kFast>(fmaf(av[i], r.v[i], bv[i]));
          r.v[i + 1] = mish<kFast>(fmaf(av[i + 1], r.v[i + 1], bv[i + 1]));
        }
      }
      r.store(yb + row[u] * ld_y);
    }
  }
  for (; v < v1; v += lanes) {
    Vec<T> r;
    const int64_t row = sub_row(sg, voxels, n, v);
    r.load(xb + row * ld_x);
#pragma unroll
    for (int i = 0; i < VEC; i += 2) {
      if (kFast) {
        const float2 o = mish2_fast(__ffma2_rn(make_float2(av[i], av[i + 1]), make_float2(r.v[i], r.v[i + 1]), make_float2(bv[i], bv[i + 1])));
        r.v[i] = o.x; r.v[i + 1] = o.y;
      } else {
        r.v[i] = mish<kFast>(fmaf(av[i], r.v[i], bv[i]));
        r.v[i + 1] = mish<kFast>(fmaf(av[i + 1], r.v[i + 1], bv[i + 1]));
      }
    }
    r.store(yb + row * ld_y);
  }
}

// ---- SE gate from channel partial sums -----------------------------------------------------------
__global__ void se_gate_kernel(const float* __restrict__ partial, int nblk, int64_t voxels, int c, int hidden,
                               const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ gate) {
  extern __shared__ double sd[];  // acc[parts][c] doubles, then mean[c] + hid[hidden] floats
  pdl_sync();
  const int n = blockIdx.x;
  const int parts = max(1, (int)blockDim.x / c);
  double* acc = sd;
  float* mean = reinterpret_cast<float*>(sd + (size_t)parts * c);
  float* hid = mean + c;
  const float* p = partial + (int64_t)n * nblk * c * 2;
  for (int idx = threadIdx.x; idx < parts * c; idx += blockDim.x) {
    const int ch = idx % c, part = idx / c;
    double s = 0.0;
    int b = part;
    for (; b + 3 * parts < nblk; b += 4 * parts) {
      const float v0 = __ldcg(p + ((int64_t)b * c + ch) * 2), v1 = __ldcg(p + ((int64_t)(b + parts) * c + ch) * 2);
      const float v2 = __ldcg(p + ((int64_t)(b + 2 * parts) * c + ch) * 2), v3 = __ldcg(p + ((int64_t)(b + 3 * parts) * c + ch) * 2);
      s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
    }
    for (; b < nblk; b += parts) s += (double)__ldcg(p + ((int64_t)b * c + ch) * 2);
    acc[part * c + ch] = s;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    double s = 0.0;
    for (int part = 0; part < parts; ++part) s += acc[part * c + ch];
    mean[ch] = (float)(s / (double)voxels);
  }
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarps = blockDim.x / 32;
  for (int j = warp; j < hidden; j += nwarps) {
    float a = 0.f;
    for (int k = lane; k < c; k += 32) a = fmaf(w1[(int64_t)j * c + k], mean[k], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) hid[j] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < hidden; ++j) a = fmaf(w2[(int64_t)ch * hidden + j], hid[j], a);
    gate[(int64_t)n * c + ch] = 1.f / (1.f + expf(-a));
  }
}

// ---- out = h * gate + res, plus channel stats of out ---------------------------------------------
template <typename T>
__global__ void scale_residual_kernel(const T* __restrict__ h, int ld_h, const T* __restrict__ res, int ld_res,
                                      T* __restrict__ out, int ld_out, int64_t voxels, int c, int nvec, int lanes,
                                      int64_t vpb, const float* __restrict__ gate, float* __restrict__ partial, SubGeom sg,
                                      SeParams se, StatsGroups og) {
  constexpr int VEC = Vec<T>::N;
  extern __shared__ float smem[];
  const int col = threadIdx.x % nvec, lane = threadIdx.x / nvec;
  const int blk = blockIdx.x, n = blockIdx.y, nblk = gridDim.x;
  const int64_t v0 = (int64_t)blk * vpb;
  const int64_t v1 = min(voxels, v0 + vpb);
  using Raw = typename Vec<T>::Raw;
  float g[VEC], s[VEC], q[VEC];
  const T* hb = h + col * VEC;
  const T* rb = res + col * VEC;
  T* ob = out + col * VEC;
  // The first batch of streaming loads (8 x 16 bytes per thread) is issued right after the grid dependency resolves and BEFORE the gate
  // is worked out: the gate needs an L2 round trip and five barriers, during which the memory system would otherwise sit idle.
  Raw ra0[4], rr0[4];
  int64_t row0[4];
  int64_t v = v0 + lane;
  bool pre = false;
  auto preload = [&] {
    pre = v + 3 * (int64_t)lanes < v1;
    if (pre) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        row0[u] = sub_row(sg, voxels, n, v + (int64_t)u * lanes);
        ra0[u] = Vec<T>::load_raw(hb + row0[u] * ld_h);
        rr0[u] = Vec<T>::load_raw(rb + row0[u] * ld_res);
      }
    }
  };
  if (se.group) {
    // squeeze-excitation gate (imagen_pytorch3D.py:617-632) from the grouped statistics of conv2's output, recomputed by every CTA.
    // Shared layout (se_scratch_bytes): doubles part[slices*c*2], tot[c*2]; floats mean[c], hid[hidden], gs[c], w1[hidden*c], w2[c*hidden]
    const int slices = max(1, (int)blockDim.x / c);
    double* part = reinterpret_cast<double*>(smem);
    double* tot = part + (size_t)slices * c * 2;
    float* mean = reinterpret_cast<float*>(tot + (size_t)c * 2);
    float* hid = mean + c;
    float* gs = hid + se.hidden;
    const float* w1s = se.w1;
    const float* w2s = se.w2;
    if (se.stage_w) {  // constant weights: staged ahead of the wait
      float* s1 = gs + c;
      float* s2 = s1 + (size_t)se.hidden * c;
      for (int idx = threadIdx.x; idx < se.hidden * c; idx += blockDim.x) {
        s1[idx] = se.w1[idx];
        s2[idx] = se.w2[idx];
      }
      w1s = s1;
      w2s = s2;
    }
    pdl_sync();
    preload();
    group_channel_totals(se.group, se.ngroups, c, n, threadIdx.x, blockDim.x, part, tot);
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) mean[ch] = (float)(tot[2 * ch] / (double)voxels);
    __syncthreads();
    const int warp = threadIdx.x / 32, wl = threadIdx.x % 32, nwarps = blockDim.x / 32;
    for (int j = warp; j < se.hidden; j += nwarps) {
      float acc = 0.f;
      for (int k = wl; k < c; k += 32) acc = fmaf(w1s[(int64_t)j * c + k], mean[k], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (wl == 0) hid[j] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      float acc = 0.f;
      for (int j = 0; j < se.hidden; ++j) acc = fmaf(w2s[(int64_t)ch * se.hidden + j], hid[j], acc);
      gs[ch] = 1.f / (1.f + expf(-acc));
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] = gs[col * VEC + i];
    __syncthreads();  // smem is reused by the block reduction below
  } else {
    pdl_sync();
    preload();
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] = gate ? __ldcg(gate + (int64_t)n * c + col * VEC + i) : 1.f;
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) s[i] = q[i] = 0.f;
  auto emit = [&](const Raw& ra, const Raw& rr, int64_t row) {
    Vec<T> a, r, o;
    a.unpack(ra);
    r.unpack(rr);
#pragma unroll
    for (int i = 0; i < VEC; ++i) o.v[i] = fmaf(a.v[i], g[i], r.v[i]);
    o.store(ob + row * ld_out);
    if (partial) {
      // statistics of what the next GroupNorm will actually read (the stored, rounded value)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float w = to_float(from_float<T>(o.v[i]));
        s[i] += w;
        q[i] = fmaf(w, w, q[i]);
      }
    }
  };
  if (pre) {
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(ra0[u], rr0[u], row0[u]);
    v += 4 * (int64_t)lanes;
  }
  // eight 16-byte loads in flight per thread (4 voxels x 2 streams), kept packed until used; voxel order per thread is unchanged
  for (; v + 3 * (int64_t)lanes < v1; v += 4 * (int64_t)lanes) {
    Raw ra[4], rr[4];
    int64_t row[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      row[u] = sub_row(sg, voxels, n, v + (int64_t)u * lanes);
      ra[u] = Vec<T>::load_raw(hb + row[u] * ld_h);
      rr[u] = Vec<T>::load_raw(rb + row[u] * ld_res);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(ra[u], rr[u], row[u]);
  }
  for (; v < v1; v += lanes) {
    const int64_t row = sub_row(sg, voxels, n, v);
    const Raw ra = Vec<T>::load_raw(hb + row * ld_h), rr = Vec<T>::load_raw(rb + row * ld_res);
    emit(ra, rr, row);
  }
  if (partial) {
    block_channel_reduce<VEC>(s, q, col, lane, lanes, c, partial + ((int64_t)n * nblk + blk) * c * 2, smem);
    StatsGroups ov = og;
    if (ov.group) { ov.group += (int64_t)n * og.ngroups * c * 2; ov.tickets += n * og.ngroups; }
    stats_group_tail(ov, partial + (int64_t)n * nblk * c * 2, 1, nblk, c, blk, 1, threadIdx.x, blockDim.x, reinterpret_cast<int*>(smem),
                     [] { __syncthreads(); });
  }
}


// ---- out = h * gate + res through a shared-memory ring fed by bulk copies --------------------------------------------------------
// The register-staged kernel above keeps 8 x 16 bytes per thread in flight only while its threads are not computing or storing:
// 100 MB (64^3 x 64 channels: two reads + one write) took 24.7 us = 0.62 of the measured HBM peak (profiles/bench_residual_r2s.jsonl).
// Here one producer lane streams 16 KB tiles of h and res into a ring of RS_STAGES stages with cp.async.bulk (up to 160 KB in flight
// per SM, independent of what the other warps do), eight consumer warps multiply, store and accumulate the statistics, and the
// squeeze-excitation gate is worked out by the consumers while the first stages are already landing.  bf16, contiguous h / res rows.
constexpr int RS_STAGES = 5, RS_TILE_BYTES = 16384, RS_CONSUMERS = 256;

__global__ void __launch_bounds__(RS_CONSUMERS + 32, 1)
scale_residual_ring_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ res, __nv_bfloat16* __restrict__ out, int ld_out,
                           int64_t voxels, int c, int64_t vpb, const float* __restrict__ gate, float* __restrict__ partial, SeParams se, StatsGroups og) {
  extern __shared__ __align__(1024) uint8_t ring_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>(((uintptr_t)ring_raw + 127) & ~(uintptr_t)127);   // [RS_STAGES][2][RS_TILE_BYTES]
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)RS_STAGES * 2 * RS_TILE_BYTES);
  uint64_t* empty = full + RS_STAGES;
  float* scratch = reinterpret_cast<float*>(empty + RS_STAGES);   // SE scratch, later the block reduction
  const int warp = threadIdx.x >> 5, lane_id = threadIdx.x & 31;
  const int blk = blockIdx.x, n = blockIdx.y, nblk = gridDim.x;
  const int64_t v0 = (int64_t)blk * vpb, v1 = min(voxels, v0 + vpb);
  const int tile_rows = RS_TILE_BYTES / (c * 2);
  const int ntiles = v1 > v0 ? (int)((v1 - v0 + tile_rows - 1) / tile_rows) : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < RS_STAGES; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), RS_CONSUMERS / 32); }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == RS_CONSUMERS / 32) {
    // ===================== producer =====================
    if (lane_id == 0) {
      pdl_sync();
      const __nv_bfloat16* hb = h + ((int64_t)n * voxels + v0) * c;
      const __nv_bfloat16* rb = res + ((int64_t)n * voxels + v0) * c;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int64_t r0 = (int64_t)t * tile_rows;
        const uint32_t bytes = (uint32_t)(min((int64_t)tile_rows, (v1 - v0) - r0) * c * 2);
        mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
        const uint32_t bar = smem_u32(&full[stage]);
        mbar_expect_tx(bar, 2 * bytes);
        bulk_load(smem_u32(ring + (size_t)(stage * 2) * RS_TILE_BYTES), hb + r0 * c, bytes, bar);
        bulk_load(smem_u32(ring + (size_t)(stage * 2 + 1) * RS_TILE_BYTES), rb + r0 * c, bytes, bar);
        if (++stage == RS_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    return;   // (no block-wide barrier below: the consumers synchronise among themselves)
  }
  // ===================== consumers (warps 0..7) =====================
  const int tid = threadIdx.x;
  auto csync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
  const int nvec = c / 8, lanes = RS_CONSUMERS / nvec;
  const int col = tid % nvec, lane = tid / nvec;
  float g[8], s[8], q[8];
  if (se.group) {
    const int slices = max(1, RS_CONSUMERS / c);
    double* part = reinterpret_cast<double*>(scratch);
    double* tot = part + (size_t)slices * c * 2;
    float* mean = reinterpret_cast<float*>(tot + (size_t)c * 2);
    float* hid = mean + c;
    float* gs = hid + se.hidden;
    pdl_wait();
    group_channel_totals(se.group, se.ngroups, c, n, tid, RS_CONSUMERS, part, tot, csync);
    for (int ch = tid; ch < c; ch += RS_CONSUMERS) mean[ch] = (float)(tot[2 * ch] / (double)voxels);
    csync();
    const int nwarps = RS_CONSUMERS / 32;
    for (int j = warp; j < se.hidden; j += nwarps) {
      float acc = 0.f;
      for (int k = lane_id; k < c; k += 32) acc = fmaf(se.w1[(int64_t)j * c + k], mean[k], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane_id == 0) hid[j] = fmaxf(acc, 0.f);
    }
    csync();
    for (int ch = tid; ch < c; ch += RS_CONSUMERS) {
      float acc = 0.f;
      for (int j = 0; j < se.hidden; ++j) acc = fmaf(se.w2[(int64_t)ch * se.hidden + j], hid[j], acc);
      gs[ch] = 1.f / (1.f + expf(-acc));
    }
    csync();
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = gs[col * 8 + i];
    csync();  // the scratch is reused by the block reduction
  } else {
    pdl_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = gate ? __ldcg(gate + (int64_t)n * c + col * 8 + i) : 1.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  __nv_bfloat16* ob = out + ((int64_t)n * voxels + v0) * ld_out + col * 8;
  int stage = 0;
  uint32_t phase = 0;
  for (int t = 0; t < ntiles; ++t) {
    const int64_t r0 = (int64_t)t * tile_rows;
    const int rows = (int)min((int64_t)tile_rows, (v1 - v0) - r0);
    mbar_wait(smem_u32(&full[stage]), phase);
    // (32-bit shared-window loads and channel pairs per FFMA2 / FADD2: the same IEEE operations per lane with about two thirds of the
    // instructions; one warp issues an instruction every ~5 cycles, so the instruction count is part of this kernel's time)
    const uint32_t hs = smem_u32(ring + (size_t)(stage * 2) * RS_TILE_BYTES) + (uint32_t)col * 16;
    const uint32_t rs = hs + RS_TILE_BYTES;
    for (int r = lane; r < rows; r += lanes) {
      Vec<__nv_bfloat16> a, x;
      a.unpack(lds_128(hs + (uint32_t)(r * c * 2)));
      x.unpack(lds_128(rs + (uint32_t)(r * c * 2)));
      uint32_t w4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 o = __ffma2_rn(make_float2(a.v[2 * i], a.v[2 * i + 1]), make_float2(g[2 * i], g[2 * i + 1]), make_float2(x.v[2 * i], x.v[2 * i + 1]));
        __nv_bfloat162 hh = __floats2bfloat162_rn(o.x, o.y);
        w4[i] = *reinterpret_cast<uint32_t*>(&hh);
      }
      *reinterpret_cast<uint4*>(ob + (r0 + r) * ld_out) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      if (partial) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // what the next GroupNorm will read: the stored, rounded values
          const float2 w = make_float2(__uint_as_float(w4[i] << 16), __uint_as_float(w4[i] & 0xffff0000u));
          const float2 ns = __fadd2_rn(make_float2(s[2 * i], s[2 * i + 1]), w);
          const float2 nq = __ffma2_rn(w, w, make_float2(q[2 * i], q[2 * i + 1]));
          s[2 * i] = ns.x; s[2 * i + 1] = ns.y; q[2 * i] = nq.x; q[2 * i + 1] = nq.y;
        }
      }
    }
    __syncwarp();
    if (lane_id == 0) mbar_arrive(smem_u32(&empty[stage]));
    if (++stage == RS_STAGES) { stage = 0; phase ^= 1; }
  }
  if (partial) {
    // per-CTA reduction of the per-thread channel sums, fixed order (as block_channel_reduce, over the consumer threads only)
    float* ss = scratch;
    float* sq = scratch + (size_t)lanes * c;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ss[lane * c + col * 8 + i] = s[i];
      sq[lane * c + col * 8 + i] = q[i];
    }
    csync();
    float* dst = partial + ((int64_t)n * nblk + blk) * c * 2;
    for (int ch = tid; ch < c; ch += RS_CONSUMERS) {
      float a = 0.f, b = 0.f;
      for (int l = 0; l < lanes; ++l) { a += ss[l * c + ch]; b += sq[l * c + ch]; }
      dst[2 * ch] = a;
      dst[2 * ch + 1] = b;
    }
    StatsGroups ov = og;
    if (ov.group) { ov.group += (int64_t)n * og.ngroups * c * 2; ov.tickets += n * og.ngroups; }
    stats_group_tail(ov, partial + (int64_t)n * nblk * c * 2, 1, nblk, c, blk, 1, tid, RS_CONSUMERS, reinterpret_cast<int*>(scratch + 2 * (size_t)lanes * c), csync);
  }
}

// ---- dst = src * scale  (row-pitched copy; the scaled skip connection, imagen_pytorch3D.py:1346, 1653) ----------
template <typename T>
__global__ void scale_copy_kernel(const T* __restrict__ src, int ld_src, T* __restrict__ dst, int ld_dst, int64_t rows,
                                  int nvec, float scale) {
  const int64_t total = rows * nvec;
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nvec;
    const int col = (int)(i - r * nvec);
    Vec<T> v;
    v.load(src + r * ld_src + col * Vec<T>::N);
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) v.v[k] *= scale;
    v.store(dst + r * ld_dst + col * Vec<T>::N);
  }
}

static inline int64_t vox_per_block(int64_t voxels, int nblk) { return (voxels + nblk - 1) / nblk; }

}  // namespace diqt

using namespace diqt;

static StatsGroups make_sink(float* group, uint32_t* tickets, int nblk, int rows_per_cta) {
  StatsGroups g;
  g.group = group;
  g.tickets = tickets;
  g.gsize = stats_group_size(nblk, rows_per_cta);
  g.ngroups = (nblk + g.gsize - 1) / g.gsize;
  return g;
}

extern "C" int diqt_stats_groups(int nblk, int rows_per_cta, int* ngroups) {
  DIQT_REQUIRE(nblk > 0 && rows_per_cta > 0 && ngroups, "stats_groups: bad arguments");
  const int gs = stats_group_size(nblk, rows_per_cta);
  *ngroups = (nblk + gs - 1) / gs;
  return DIQT_OK;
}

static int channel_stats_impl(const void* x, int dtype, int n, int64_t voxels, int c, int ld, int nblk,
                              float* partial, float* group, uint32_t* tickets, int sub_f, int sub_h, void* stream) {
  const SubGeom sg{sub_f, sub_h};
  const StatsGroups og = make_sink(group, tickets, nblk, 1);
  if (group) DIQT_REQUIRE(tickets && sub_f <= 1, "channel_stats: grouped sink needs tickets and plain volumes");
  if (sub_f > 1) DIQT_REQUIRE(n == sub_f * sub_f * sub_f && voxels == (int64_t)sub_h * sub_h * sub_h, "channel_stats: sub-volume geometry mismatch");
  const int vec = dtype == DIQT_BF16 ? 8 : 4;
  DIQT_REQUIRE(x && partial && n > 0 && voxels > 0 && nblk > 0, "channel_stats: bad arguments");
  DIQT_REQUIRE(c % vec == 0 && ld % vec == 0 && c / vec <= 256, "channel_stats: c=%d ld=%d not a multiple of %d", c, ld, vec);
  RowMap m = make_rowmap(c, vec, 512);
  dim3 grid(nblk, n);
  size_t sh = (size_t)2 * m.lanes * c * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DIQT_BF16)
    launch_pdl(channel_stats_kernel<__nv_bfloat16>, grid, m.threads, sh, st, (const __nv_bfloat16*)x, voxels, c, ld, m.nvec, m.lanes,
               vox_per_block(voxels, nblk), partial, sg, og);
  else
    launch_pdl(channel_stats_kernel<float>, grid, m.threads, sh, st, (const float*)x, voxels, c, ld, m.nvec, m.lanes,
               vox_per_block(voxels, nblk), partial, sg, og);
  return check_launch("channel_stats");
}

extern "C" int diqt_channel_stats(const void* x, int dtype, int n, int64_t voxels, int c, int ld, int nblk,
                                  float* partial, int sub_f, int sub_h, void* stream) {
  return channel_stats_impl(x, dtype, n, voxels, c, ld, nblk, partial, nullptr, nullptr, sub_f, sub_h, stream);
}

extern "C" int diqt_channel_stats_g(const void* x, int dtype, int n, int64_t voxels, int c, int ld, int nblk, float* partial,
                                    float* group, uint32_t* tickets, void* stream) {
  DIQT_REQUIRE(group && tickets, "channel_stats_g: null sink");
  return channel_stats_impl(x, dtype, n, voxels, c, ld, nblk, partial, group, tickets, 0, 0, stream);
}

extern "C" int diqt_gn_finalize(const float* partial, int n, int nblk, int64_t voxels, int c, int groups, float eps,
                                const float* gamma, const float* beta, const float* film, int film_ld,
                                const int32_t* film_row, int film_row_stride_n, float* a, float* b, void* stream) {
  DIQT_REQUIRE(partial && gamma && beta && a && b, "gn_finalize: null pointer");
  DIQT_REQUIRE(groups > 0 && c % groups == 0 && c <= 4096, "gn_finalize: c=%d groups=%d", c, groups);
  const int threads = 1024;
  const int parts = threads / c > 0 ? threads / c : 1;
  size_t sh = ((size_t)parts * c * 2 + (size_t)c * 2 + (size_t)groups * 2) * sizeof(double);
  static bool attr_set = false;
  if (sh > 48 * 1024 && !attr_set) {
    DIQT_CUDA(cudaFuncSetAttribute(gn_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  launch_pdl(gn_finalize_kernel, n, threads, sh, (cudaStream_t)stream, partial, nblk, voxels, c, groups, eps, gamma, beta, film,
                                                               film_ld, film_row, film_row_stride_n, a, b);
  return check_launch("gn_finalize");
}

static int affine_mish_impl(const void* x, int ld_x, void* y, int ld_y, int dtype, int n, int64_t voxels, int c,
                            const float* a, const float* b, int nblk, int sub_f, int sub_h, const GnParams& gp, void* stream) {
  const SubGeom sg{sub_f, sub_h};
  if (sub_f > 1) DIQT_REQUIRE(n == sub_f * sub_f * sub_f && voxels == (int64_t)sub_h * sub_h * sub_h, "affine_mish: sub-volume geometry mismatch");
  const int vec = dtype == DIQT_BF16 ? 8 : 4;
  DIQT_REQUIRE(x && y && ((a && b) || gp.group) && nblk > 0, "affine_mish: bad arguments");
  RowMap m = make_rowmap(c, vec, 256);
  const size_t gsh = gp.group ? gn_scratch_bytes(c, gp.groups, m.threads) : 0;
  DIQT_REQUIRE(c % vec == 0 && ld_x % vec == 0 && ld_y % vec == 0 && c / vec <= 256, "affine_mish: c=%d not a multiple of %d", c, vec);
  DIQT_REQUIRE(gsh <= 160 * 1024, "affine_mish: c=%d needs %zu bytes of shared memory", c, gsh);
  static bool big_smem = false;
  if (gsh > 48 * 1024 && !big_smem) {  // very wide levels only (c >= 1024)
    DIQT_CUDA(cudaFuncSetAttribute(affine_mish_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DIQT_CUDA(cudaFuncSetAttribute(affine_mish_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    big_smem = true;
  }
  dim3 grid(nblk, n);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DIQT_BF16)
    launch_pdl(affine_mish_kernel<__nv_bfloat16, true>, grid, m.threads, gsh, st, (const __nv_bfloat16*)x, ld_x, (__nv_bfloat16*)y, ld_y,
               voxels, c, m.nvec, m.lanes, vox_per_block(voxels, nblk), a, b, sg, gp);
  else
    launch_pdl(affine_mish_kernel<float, false>, grid, m.threads, gsh, st, (const float*)x, ld_x, (float*)y, ld_y, voxels, c, m.nvec,
               m.lanes, vox_per_block(voxels, nblk), a, b, sg, gp);
  return check_launch("affine_mish");
}

extern "C" int diqt_affine_mish(const void* x, int ld_x, void* y, int ld_y, int dtype, int n, int64_t voxels, int c,
                                const float* a, const float* b, int nblk, int sub_f, int sub_h, void* stream) {
  GnParams gp = {};
  return affine_mish_impl(x, ld_x, y, ld_y, dtype, n, voxels, c, a, b, nblk, sub_f, sub_h, gp, stream);
}

extern "C" int diqt_gn_mish_g(const void* x, int ld_x, void* y, int ld_y, int dtype, int n, int64_t voxels, int c, const float* group,
                              int ngroups, int groups, float eps, const float* gamma, const float* beta, const float* film, int film_ld,
                              const int32_t* film_row, int film_row_stride_n, int nblk, void* stream) {
  DIQT_REQUIRE(group && ngroups > 0 && gamma && beta, "gn_mish_g: null pointer");
  DIQT_REQUIRE(groups > 0 && c % groups == 0 && c <= 2048, "gn_mish_g: c=%d groups=%d", c, groups);
  GnParams gp = {group, ngroups, (long long)voxels, c, groups, eps, gamma, beta, film, film_ld, film_row, film_row_stride_n};
  return affine_mish_impl(x, ld_x, y, ld_y, dtype, n, voxels, c, nullptr, nullptr, nblk, 0, 0, gp, stream);
}

extern "C" int diqt_se_gate(const float* partial, int n, int nblk, int64_t voxels, int c, int hidden, const float* w1,
                            const float* w2, float* gate, void* stream) {
  DIQT_REQUIRE(partial && w1 && w2 && gate && hidden > 0, "se_gate: bad arguments (hidden=%d)", hidden);
  DIQT_REQUIRE(c <= 2048, "se_gate: c=%d too large", c);
  const int threads = 1024;
  const int parts = threads / c > 0 ? threads / c : 1;
  size_t sh = (size_t)parts * c * sizeof(double) + (size_t)(c + hidden) * sizeof(float);
  launch_pdl(se_gate_kernel, n, threads, sh, (cudaStream_t)stream, partial, nblk, voxels, c, hidden, w1, w2, gate);
  return check_launch("se_gate");
}

static int scale_residual_impl(const void* h, int ld_h, const void* res, int ld_res, void* out, int ld_out, int dtype,
                               int n, int64_t voxels, int c, const float* gate, int nblk, float* partial,
                               int sub_f, int sub_h, const SeParams& se, float* group_out, uint32_t* tickets, void* stream) {
  const SubGeom sg{sub_f, sub_h};
  const StatsGroups og = make_sink(group_out, tickets, nblk, 1);
  if (group_out) DIQT_REQUIRE(partial && tickets && sub_f <= 1, "scale_residual: grouped sink needs partial rows, tickets and plain volumes");
  if (sub_f > 1) DIQT_REQUIRE(n == sub_f * sub_f * sub_f && voxels == (int64_t)sub_h * sub_h * sub_h, "scale_residual: sub-volume geometry mismatch");
  const int vec = dtype == DIQT_BF16 ? 8 : 4;
  DIQT_REQUIRE(h && res && out && nblk > 0, "scale_residual: bad arguments");
  DIQT_REQUIRE(c % vec == 0 && ld_h % vec == 0 && ld_res % vec == 0 && ld_out % vec == 0 && c / vec <= 256,
               "scale_residual: c=%d not a multiple of %d", c, vec);
  cudaStream_t st0 = (cudaStream_t)stream;
  static int ring_env = -1;
  if (ring_env < 0) { const char* e = getenv("DIQT_DISABLE_RESIDUAL_RING"); ring_env = (e && e[0] == '1') ? 0 : 1; }
  // bf16, plain volumes, contiguous input rows, channel counts whose 16 KB tiles split evenly over the 256 consumer threads
  if (ring_env && dtype == DIQT_BF16 && sub_f <= 1 && ld_h == c && ld_res == c && c % 8 == 0 && (c == 64 || c == 128 || c == 256) && voxels / nblk >= 128) {
    const int lanes = RS_CONSUMERS / (c / 8);
    size_t scratch = (size_t)2 * lanes * c * sizeof(float) + 64;
    if (se.group) {
      const int slices = RS_CONSUMERS / c > 0 ? RS_CONSUMERS / c : 1;
      scratch = std::max(scratch, ((size_t)slices * c * 2 + (size_t)c * 2) * sizeof(double) + ((size_t)2 * c + se.hidden) * sizeof(float));
    }
    const size_t sh = (size_t)RS_STAGES * 2 * RS_TILE_BYTES + 2 * RS_STAGES * sizeof(uint64_t) + scratch + 256;
    static bool attr_done = false;
    if (!attr_done) {
      DIQT_CUDA(cudaFuncSetAttribute(scale_residual_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_done = true;
    }
    launch_pdl(scale_residual_ring_kernel, dim3(nblk, n), RS_CONSUMERS + 32, sh, st0, (const __nv_bfloat16*)h, (const __nv_bfloat16*)res, (__nv_bfloat16*)out,
               ld_out, voxels, c, vox_per_block(voxels, nblk), gate, partial, se, og);
    return check_launch("scale_residual");
  }
  RowMap m = make_rowmap(c, vec, 512);  // one 512-thread CTA per SM; 2 x 256 or 4 x 256 threads per SM measured no better (r1s A/B)
  dim3 grid(nblk, n);
  size_t sh = partial ? (size_t)2 * m.lanes * c * sizeof(float) : 0;
  SeParams sev = se;
  if (se.group) {
    const int slices = m.threads / c > 0 ? m.threads / c : 1;
    const size_t base = ((size_t)slices * c * 2 + (size_t)c * 2) * sizeof(double) + ((size_t)2 * c + se.hidden) * sizeof(float);
    const size_t wbytes = (size_t)2 * se.hidden * c * sizeof(float);
    sev.stage_w = base + wbytes <= 40 * 1024 ? 1 : 0;
    sh = std::max(sh, base + (sev.stage_w ? wbytes : 0));
  }
  DIQT_REQUIRE(sh <= 48 * 1024, "scale_residual: c=%d hidden=%d needs %zu bytes of shared memory", c, se.hidden, sh);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DIQT_BF16)
    launch_pdl(scale_residual_kernel<__nv_bfloat16>, grid, m.threads, sh, st, 
        (const __nv_bfloat16*)h, ld_h, (const __nv_bfloat16*)res, ld_res, (__nv_bfloat16*)out, ld_out, voxels, c, m.nvec,
        m.lanes, vox_per_block(voxels, nblk), gate, partial, sg, sev, og);
  else
    launch_pdl(scale_residual_kernel<float>, grid, m.threads, sh, st, (const float*)h, ld_h, (const float*)res, ld_res, (float*)out,
               ld_out, voxels, c, m.nvec, m.lanes, vox_per_block(voxels, nblk), gate, partial, sg, sev, og);
  return check_launch("scale_residual");
}

extern "C" int diqt_scale_residual(const void* h, int ld_h, const void* res, int ld_res, void* out, int ld_out, int dtype,
                                   int n, int64_t voxels, int c, const float* gate, int nblk, float* partial,
                                   int sub_f, int sub_h, void* stream) {
  SeParams se = {};
  return scale_residual_impl(h, ld_h, res, ld_res, out, ld_out, dtype, n, voxels, c, gate, nblk, partial, sub_f, sub_h, se, nullptr, nullptr,
                             stream);
}

extern "C" int diqt_scale_residual_g(const void* h, int ld_h, const void* res, int ld_res, void* out, int ld_out, int dtype, int n,
                                     int64_t voxels, int c, const float* se_group, int se_ngroups, int hidden, const float* w1,
                                     const float* w2, int nblk, float* partial, float* group_out, uint32_t* tickets, void* stream) {
  SeParams se = {se_group, se_ngroups, hidden, w1, w2, 0};
  if (se_group) DIQT_REQUIRE(se_ngroups > 0 && hidden > 0 && w1 && w2, "scale_residual_g: incomplete SE description");
  return scale_residual_impl(h, ld_h, res, ld_res, out, ld_out, dtype, n, voxels, c, nullptr, nblk, partial, 0, 0, se, group_out, tickets,
                             stream);
}

extern "C" int diqt_scale_copy(const void* src, int ld_src, void* dst, int ld_dst, int dtype, int64_t rows, int c, float scale,
                               void* stream) {
  const int vec = dtype == DIQT_BF16 ? 8 : 4;
  DIQT_REQUIRE(src && dst && rows > 0, "scale_copy: bad arguments");
  DIQT_REQUIRE(c % vec == 0 && ld_src % vec == 0 && ld_dst % vec == 0, "scale_copy: c=%d not a multiple of %d", c, vec);
  const int nvec = c / vec;
  int64_t blocks = (rows * nvec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DIQT_BF16)
    launch_pdl(scale_copy_kernel<__nv_bfloat16>, (unsigned)blocks, 256, 0, st, (const __nv_bfloat16*)src, ld_src, (__nv_bfloat16*)dst, ld_dst,
                                                                     rows, nvec, scale);
  else
    launch_pdl(scale_copy_kernel<float>, (unsigned)blocks, 256, 0, st, (const float*)src, ld_src, (float*)dst, ld_dst, rows, nvec, scale);
  return check_launch("scale_copy");
}
