// Backward kernels of the training step (Imagen.p_losses -> loss.backward() -> optimizer step; imagen_pytorch3D.py:2277-2387 and the
// autograd graph PyTorch builds for Unet.forward :1554-1684).  Channels-last activations [n][voxels][c] with row pitch ld, fp32 or bf16
// storage, fp32 math, every reduction in a fixed order (bitwise reproducible gradients).
//
// The reverse of one `Block` (GroupNorm -> FiLM -> Mish -> conv, :546-565) and of the SE gate / residual join (:612-632) needs exactly
// two shapes of pass over an activation tensor:
//   reduce : per (n, channel)   S1 = sum_v t ,  S2 = sum_v t * x       t = dz * mish'(a_c x + b_c)   (GroupNorm / FiLM / Mish)
//                                                                      t = dz                        (SE gate: x = the gated tensor)
//   apply  : out = c1 * t + c2 * x + c3 (+ acc)      with per-(n, channel) coefficients: the GroupNorm input gradient
//            r (1+s) gamma dw - r m1 - r^2 m2 (x - mu), or  gate * dz + dmean / V  for the SE join; `acc` adds the residual branch.
// The coefficients are a few (n, c) vectors derived from S1 / S2 on the host side of the C ABI (diffusioniqt_b200/train.py).
// Weight gradients: conv_wgrad_simt (any shape, fp32 accumulation, per-chunk partials summed in a fixed order); the tcgen05 version for
// channel counts that are multiples of 64 lives in wgrad_tc.cu.  Data gradients of the convolutions are convolutions with the flipped,
// transposed weights and run through the forward kernels.
#include <algorithm>

#include "common.cuh"

namespace diqt {

namespace {

struct BwMap {
  int nvec, lanes, threads;
};
inline BwMap bw_map(int c, int vec, int max_threads) {
  BwMap m;
  m.nvec = c / vec;
  m.lanes = max_threads / m.nvec;
  if (m.lanes < 1) m.lanes = 1;
  m.threads = m.nvec * m.lanes;
  return m;
}
inline int64_t bw_vpb(int64_t voxels, int nblk) { return (voxels + nblk - 1) / nblk; }

// d/dw [ w tanh(softplus(w)) ] = T + w sigma(w) (1 - T^2),  T = n / (n + 2),  n = u (u + 2),  u = e^w ;  1 - T^2 = 4 (n + 1) / (n + 2)^2
template <bool kFast>
__device__ __forceinline__ float mish_grad(float w) {
  if (kFast) {   // bf16 storage: ex2.approx + two rcp.approx (relative error ~1e-6, far below the bf16 rounding of the result)
    float u, d, e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(fminf(w, 20.f) * 1.4426950408889634f));
    const float n = u * (u + 2.f);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(n + 2.f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(1.f + u));
    return fmaf(w * u * e, 4.f * (n + 1.f) * d * d, n * d);
  }
  const float u = expf(fminf(w, 20.f));      // w > 20: T = 1 and the second term vanishes in fp32
  const float n = u * (u + 2.f);
  const float d = 1.f / (n + 2.f);
  const float T = n * d;
  const float sig = u / (1.f + u);
  return fmaf(w * sig, 4.f * (n + 1.f) * d * d, T);
}
template <typename T>
struct FastMath { static constexpr bool value = false; };
template <>
struct FastMath<__nv_bfloat16> { static constexpr bool value = true; };

}  // namespace

// ---- reduce: partial[n][blk][c][2] = (sum t, sum t * x) over the block's voxels ------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(256) bwd_reduce_kernel(const T* x, int ld_x, const T* dz, int ld_dz, int64_t voxels, int c, int nvec, int lanes,
                                                         int64_t vpb, const float* a, const float* b, float* partial) {
  constexpr int VEC = Vec<T>::N;
  extern __shared__ float bw_smem[];
  const int col = threadIdx.x % nvec, lane = threadIdx.x / nvec;
  const int blk = blockIdx.x, n = blockIdx.y, nblk = gridDim.x;
  const int64_t v0 = (int64_t)blk * vpb, v1 = min(voxels, v0 + vpb);
  pdl_sync();
  float av[VEC], bv[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    av[i] = MODE ? __ldcg(a + (int64_t)n * c + col * VEC + i) : 1.f;
    bv[i] = MODE ? __ldcg(b + (int64_t)n * c + col * VEC + i) : 0.f;
    s1[i] = s2[i] = 0.f;
  }
  const T* xb = x + (int64_t)n * voxels * ld_x + col * VEC;
  const T* gb = dz + (int64_t)n * voxels * ld_dz + col * VEC;
  for (int64_t v = v0 + lane; v < v1; v += 2 * (int64_t)lanes) {
    Vec<T> xr[2], gr[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t vv = v + (int64_t)u * lanes;
      if (vv < v1) {
        xr[u].load(xb + vv * ld_x);
        gr[u].load(gb + vv * ld_dz);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) xr[u].v[i] = gr[u].v[i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float t = MODE ? gr[u].v[i] * mish_grad<FastMath<T>::value>(fmaf(av[i], xr[u].v[i], bv[i])) : gr[u].v[i];
        s1[i] += t;
        s2[i] = fmaf(t, xr[u].v[i], s2[i]);
      }
  }
  float* ss = bw_smem;
  float* sq = bw_smem + (size_t)lanes * c;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    ss[lane * c + col * VEC + i] = s1[i];
    sq[lane * c + col * VEC + i] = s2[i];
  }
  __syncthreads();
  float* out = partial + ((int64_t)n * nblk + blk) * c * 2;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float p = 0.f, q = 0.f;
    for (int l = 0; l < lanes; ++l) {  // fixed order
      p += ss[l * c + ch];
      q += sq[l * c + ch];
    }
    out[2 * ch] = p;
    out[2 * ch + 1] = q;
  }
}

// ---- apply: out = c1 * t + c2 * x + c3 (+ acc) ------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(256) bwd_apply_kernel(const T* x, int ld_x, const T* dz, int ld_dz, const T* acc, int ld_acc, T* out, int ld_out,
                                                        int64_t voxels, int c, int nvec, int lanes, int64_t vpb, const float* a, const float* b,
                                                        const float* c1, const float* c2, const float* c3) {
  constexpr int VEC = Vec<T>::N;
  const int col = threadIdx.x % nvec, lane = threadIdx.x / nvec;
  const int n = blockIdx.y;
  const int64_t v0 = (int64_t)blockIdx.x * vpb, v1 = min(voxels, v0 + vpb);
  pdl_sync();
  float av[VEC], bv[VEC], k1[VEC], k2[VEC], k3[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int64_t o = (int64_t)n * c + col * VEC + i;
    av[i] = MODE ? __ldcg(a + o) : 1.f;
    bv[i] = MODE ? __ldcg(b + o) : 0.f;
    k1[i] = __ldcg(c1 + o);
    k2[i] = c2 ? __ldcg(c2 + o) : 0.f;
    k3[i] = c3 ? __ldcg(c3 + o) : 0.f;
  }
  const int64_t base = (int64_t)n * voxels;
  const bool need_x = MODE || c2;
  for (int64_t v = v0 + lane; v < v1; v += 2 * (int64_t)lanes) {   // two rows in flight per thread
    Vec<T> xr[2], gr[2], ar[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t vv = v + (int64_t)u * lanes;
      if (vv < v1) {
        gr[u].load(dz + (base + vv) * ld_dz + col * VEC);
        if (need_x) xr[u].load(x + (base + vv) * ld_x + col * VEC);
        if (acc) ar[u].load(acc + (base + vv) * ld_acc + col * VEC);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t vv = v + (int64_t)u * lanes;
      if (vv >= v1) break;
      Vec<T> o;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float xv = need_x ? xr[u].v[i] : 0.f;
        const float t = MODE ? gr[u].v[i] * mish_grad<FastMath<T>::value>(fmaf(av[i], xv, bv[i])) : gr[u].v[i];
        float r = fmaf(k1[i], t, fmaf(k2[i], xv, k3[i]));
        if (acc) r += ar[u].v[i];
        o.v[i] = r;
      }
      o.store(out + (base + vv) * ld_out + col * VEC);
    }
  }
}

// ---- the (n, c)-sized algebra between reduce and apply for GroupNorm -> FiLM -> Mish (see gn_backward_coefficients in train.py for the
// derivation): one CTA per volume, fp64, fixed summation order; d gamma / d beta leave as per-volume rows [n][c].  fpart: the forward statistics (sum x, sum x^2), bpart: (sum dw, sum dw x).
__global__ void __launch_bounds__(1024) gn_bwd_finalize_kernel(const float* fpart, int nblk_f, const float* bpart, int nblk_b, int n, int64_t voxels, int c,
                                                               int groups, float eps, const float* gamma, const float* beta, const float* film,
                                                               float* c1, float* c2, float* c3, float* dgamma, float* dbeta, float* dfilm) {
  extern __shared__ double fin_sm[];
  const int parts = max(1, (int)blockDim.x / c), cpg = c / groups;
  double* acc = fin_sm;                          // [parts][c][4]
  double* tot = acc + (size_t)parts * c * 4;     // [c][4]: sum x, sum x^2, S1, S2x   (later [c][2]: kg S1, kg S2)
  double* gst = tot + (size_t)c * 4;             // [groups][4]: mean, rstd, m1, m2
  pdl_sync();
  const double cnt = (double)voxels * cpg;
  {
    const int nv = blockIdx.x;      // one CTA per volume; d gamma / d beta leave as per-volume rows [n][c] (summed over n by the caller)
    for (int idx = threadIdx.x; idx < parts * c; idx += blockDim.x) {
      const int ch = idx % c, part = idx / c;
      double s = 0, q = 0, s1 = 0, s2 = 0;
      auto sum_rows = [&](const float* p, int nblk, double& o0, double& o1) {   // four loads in flight, summation order unchanged
        const float2 zero = make_float2(0.f, 0.f);
        for (int b = part; b < nblk; b += 4 * parts) {
          float2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            v[u] = b + u * parts < nblk ? __ldcg(reinterpret_cast<const float2*>(p + (((int64_t)nv * nblk + b + u * parts) * c + ch) * 2)) : zero;
#pragma unroll
          for (int u = 0; u < 4; ++u) { o0 += v[u].x; o1 += v[u].y; }
        }
      };
      sum_rows(fpart, nblk_f, s, q);
      sum_rows(bpart, nblk_b, s1, s2);
      double* a = acc + ((size_t)part * c + ch) * 4;
      a[0] = s; a[1] = q; a[2] = s1; a[3] = s2;
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      double t[4] = {0, 0, 0, 0};
      for (int pz = 0; pz < parts; ++pz)
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] += acc[((size_t)pz * c + ch) * 4 + j];
#pragma unroll
      for (int j = 0; j < 4; ++j) tot[ch * 4 + j] = t[j];
    }
    __syncthreads();
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      double s = 0, q = 0;
      for (int i = 0; i < cpg; ++i) { s += tot[(g * cpg + i) * 4]; q += tot[(g * cpg + i) * 4 + 1]; }
      const double mean = s / cnt;
      double var = q / cnt - mean * mean;
      if (var < 0) var = 0;
      gst[g * 4] = mean;
      gst[g * 4 + 1] = 1.0 / sqrt(var + (double)eps);
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const int g = ch / cpg;
      const double mu = gst[g * 4], r = gst[g * 4 + 1];
      const double S1 = tot[ch * 4 + 2], S2 = r * (tot[ch * 4 + 3] - mu * S1);
      const double k = film ? 1.0 + (double)film[(int64_t)nv * 2 * c + ch] : 1.0;
      const double ga = gamma[ch], be = beta[ch];
      dbeta[(int64_t)nv * c + ch] = (float)(k * S1);
      dgamma[(int64_t)nv * c + ch] = (float)(k * S2);
      if (dfilm) {
        dfilm[(int64_t)nv * 2 * c + ch] = (float)(ga * S2 + be * S1);
        dfilm[(int64_t)nv * 2 * c + c + ch] = (float)S1;
      }
      tot[ch * 4] = k * ga * S1;      // (sum x / sum x^2 are consumed: reuse the slots)
      tot[ch * 4 + 1] = k * ga * S2;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      double a1 = 0, a2 = 0;
      for (int i = 0; i < cpg; ++i) { a1 += tot[(g * cpg + i) * 4]; a2 += tot[(g * cpg + i) * 4 + 1]; }
      gst[g * 4 + 2] = a1 / cnt;
      gst[g * 4 + 3] = a2 / cnt;
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const int g = ch / cpg;
      const double mu = gst[g * 4], r = gst[g * 4 + 1], m1 = gst[g * 4 + 2], m2 = gst[g * 4 + 3];
      const double k = film ? 1.0 + (double)film[(int64_t)nv * 2 * c + ch] : 1.0;
      const int64_t o = (int64_t)nv * c + ch;
      c1[o] = (float)(r * k * (double)gamma[ch]);
      c2[o] = (float)(-r * r * m2);
      c3[o] = (float)(r * (r * m2 * mu - m1));
    }
  }
}

// ---- reverse of the SE gate MLP (SE3D :617-632: gate = sigmoid(W2 relu(W1 mean_v h))), one CTA, fp32, fixed order.
// fpart: forward statistics of h ([n][nblk_f][c][2], sum h first); bpart: diqt_bwd_reduce mode 0 ([n][nblk_b][c][2], sum d_out * h second).
// Outputs: c3[n][c] = d loss / d mean / V (the additive term of diqt_bwd_apply) and per-volume dw1[n][hid][c], dw2[n][c][hid]; one CTA per volume.
__global__ void __launch_bounds__(512) se_bwd_kernel(const float* fpart, int nblk_f, const float* bpart, int nblk_b, int n, int64_t voxels, int c, int hid,
                                                     const float* w1, const float* w2, const float* gate, float* c3, float* dw1, float* dw2) {
  extern __shared__ float se_sm[];
  float* mean = se_sm;            // [c]
  float* dy = mean + c;           // [c]
  float* z = dy + c;              // [hid]
  float* dz = z + hid;            // [hid]
  pdl_sync();
  const int tid = threadIdx.x, nt = blockDim.x;
  {
    const int nv = blockIdx.x;      // one CTA per volume; dw1 / dw2 leave as per-volume slabs [n][hid * c] (summed over n by the caller)
    dw1 += (int64_t)nv * hid * c;
    dw2 += (int64_t)nv * hid * c;
    // row sums: blockDim / c slices per channel, four loads in flight each, slices added in a fixed order
    {
      const int parts = max(1, nt / c);
      float* ps = dz + hid;          // [parts][c][2]
      for (int idx = tid; idx < parts * c; idx += nt) {
        const int ch = idx % c, part = idx / c;
        float s = 0.f, d = 0.f;
        for (int b = part; b < nblk_f; b += 4 * parts) {
          float v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = b + u * parts < nblk_f ? __ldcg(fpart + (((int64_t)nv * nblk_f + b + u * parts) * c + ch) * 2) : 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u) s += v[u];
        }
        for (int b = part; b < nblk_b; b += 4 * parts) {
          float v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = b + u * parts < nblk_b ? __ldcg(bpart + (((int64_t)nv * nblk_b + b + u * parts) * c + ch) * 2 + 1) : 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u) d += v[u];
        }
        ps[(part * c + ch) * 2] = s;
        ps[(part * c + ch) * 2 + 1] = d;
      }
      __syncthreads();
      for (int ch = tid; ch < c; ch += nt) {
        float s = 0.f, d = 0.f;
        for (int pz = 0; pz < parts; ++pz) { s += ps[(pz * c + ch) * 2]; d += ps[(pz * c + ch) * 2 + 1]; }
        const float g = gate[(int64_t)nv * c + ch];
        mean[ch] = s / (float)voxels;
        dy[ch] = d * g * (1.f - g);          // through the sigmoid
      }
    }
    __syncthreads();
    for (int j = tid; j < hid; j += nt) {
      float a = 0.f, b = 0.f;
      for (int ch = 0; ch < c; ++ch) {
        a = fmaf(w1[j * c + ch], mean[ch], a);        // z = W1 mean
        b = fmaf(w2[ch * hid + j], dy[ch], b);        // d relu-out = W2^T dy
      }
      z[j] = a;
      dz[j] = a > 0.f ? b : 0.f;                      // through the ReLU
    }
    __syncthreads();
    for (int i = tid; i < hid * c; i += nt) {
      const int j1 = i / c, c1i = i - j1 * c;         // dw1[j][ch] += dz[j] mean[ch]
      dw1[i] = dz[j1] * mean[c1i];
      const int c2i = i / hid, j2 = i - c2i * hid;    // dw2[ch][j] = dy[ch] relu(z[j])
      dw2[i] = dy[c2i] * fmaxf(z[j2], 0.f);
    }
    for (int ch = tid; ch < c; ch += nt) {
      float d = 0.f;
      for (int j = 0; j < hid; ++j) d = fmaf(w1[j * c + ch], dz[j], d);   // d mean = W1^T dz
      c3[(int64_t)nv * c + ch] = d / (float)voxels;
    }
  }
}

// ---- weight gradient, CUDA cores: partial[chunk][tap][c_out][c_in] over the chunk's voxels ---------------------------------------
constexpr int WG_VT = 32;   // voxels per shared-memory tile
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const T* x, int ld_x, const T* dy, int ld_dy, int n, int d0, int d1, int d2, int c_in,
                                                              int c_out, int taps, int ci_tiles, int64_t vox_per_chunk, float* partial) {
  __shared__ float dys[WG_VT][65], xs[WG_VT][65];
  const int chunk = blockIdx.x, tap = blockIdx.y;
  const int co0 = (blockIdx.z / ci_tiles) * 64, ci0 = (blockIdx.z % ci_tiles) * 64;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int oz = taps == 27 ? tap / 9 - 1 : 0, oy = taps == 27 ? (tap / 3) % 3 - 1 : 0, ox = taps == 27 ? tap % 3 - 1 : 0;
  const int64_t vol = (int64_t)d0 * d1 * d2, total = vol * n;
  const int64_t r0 = (int64_t)chunk * vox_per_chunk, r1 = min(total, r0 + vox_per_chunk);
  pdl_sync();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t base = r0; base < r1; base += WG_VT) {
    for (int idx = threadIdx.x; idx < WG_VT * 64; idx += 256) {
      const int r = idx >> 6, ch = idx & 63;
      const int64_t row = base + r;
      float gv = 0.f, xv = 0.f;
      if (row < r1) {
        if (co0 + ch < c_out) gv = to_float(dy[row * ld_dy + co0 + ch]);
        if (ci0 + ch < c_in) {
          const int64_t vv = row % vol;
          const int z = (int)(vv / ((int64_t)d1 * d2)), y = (int)((vv / d2) % d1), xx = (int)(vv % d2);
          const int zz = z + oz, yy = y + oy, xq = xx + ox;
          if ((unsigned)zz < (unsigned)d0 && (unsigned)yy < (unsigned)d1 && (unsigned)xq < (unsigned)d2)
            xv = to_float(x[(row + ((int64_t)oz * d1 + oy) * d2 + ox) * ld_x + ci0 + ch]);
        }
      }
      dys[r][ch] = gv;
      xs[r][ch] = xv;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < WG_VT; ++r) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = dys[r][ty * 4 + i];
        bv[i] = xs[r][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = partial + ((int64_t)chunk * taps + tap) * c_out * c_in;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + ty * 4 + i, ci = ci0 + tx * 4 + j;
      if (co < c_out && ci < c_in) dst[(int64_t)co * c_in + ci] = acc[i][j];
    }
}

// Narrow inputs (c_in <= 16: init_conv :1291 sees the 2-channel concat of x_t and the low-field patch): thread = (output channel, voxel
// lane), the c_in input values of a voxel are a broadcast load, c_in FMAs per dy value; same partial layout as the tiled kernel.
constexpr int WGS_LANES = 4;
template <typename T, int CI>
__global__ void __launch_bounds__(64 * WGS_LANES) conv_wgrad_narrow_kernel(const T* x, int ld_x, const T* dy, int ld_dy, int n, int d0, int d1, int d2, int c_in,
                                                                           int c_out, int taps, int64_t vox_per_chunk, float* partial) {
  __shared__ float red[WGS_LANES][64][CI];
  const int chunk = blockIdx.x, tap = blockIdx.y, co0 = blockIdx.z * 64;
  const int co = co0 + (threadIdx.x & 63), lane = threadIdx.x >> 6;
  const int oz = taps == 27 ? tap / 9 - 1 : 0, oy = taps == 27 ? (tap / 3) % 3 - 1 : 0, ox = taps == 27 ? tap % 3 - 1 : 0;
  const int64_t vol = (int64_t)d0 * d1 * d2, total = vol * n;
  const int64_t r0 = (int64_t)chunk * vox_per_chunk, r1 = min(total, r0 + vox_per_chunk);
  const int64_t shift = ((int64_t)oz * d1 + oy) * d2 + ox;
  pdl_sync();
  float acc[CI];
#pragma unroll
  for (int i = 0; i < CI; ++i) acc[i] = 0.f;
  // voxel coordinates are advanced incrementally (a 64-bit division per voxel made the first version of this kernel 20x slower), and four
  // voxels are in flight per thread: the loop is a chain of L2 round trips otherwise
  int cz, cy, cx;
  {
    const int64_t vv = (r0 + lane) % vol;
    cz = (int)(vv / ((int64_t)d1 * d2)); cy = (int)((vv / d2) % d1); cx = (int)(vv % d2);
  }
  constexpr int U = 4;
  for (int64_t row = r0 + lane; row < r1; row += WGS_LANES * U) {
    float g[U], xv[U][CI];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = row + (int64_t)u * WGS_LANES;
      const int z = cz + oz, y = cy + oy, xx = cx + ox;
      cx += WGS_LANES;
      while (cx >= d2) {
        cx -= d2;
        if (++cy == d1) { cy = 0; if (++cz == d0) cz = 0; }
      }
      const bool ok = rr < r1 && (unsigned)z < (unsigned)d0 && (unsigned)y < (unsigned)d1 && (unsigned)xx < (unsigned)d2;   // warp-uniform
      g[u] = (ok && co < c_out) ? to_float(dy[rr * ld_dy + co]) : 0.f;
      const T* xr = x + (rr + shift) * ld_x;
#pragma unroll
      for (int i = 0; i < CI; ++i) xv[u][i] = (ok && i < c_in) ? to_float(xr[i]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int i = 0; i < CI; ++i) acc[i] = fmaf(g[u], xv[u][i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < CI; ++i) red[lane][threadIdx.x & 63][i] = acc[i];
  __syncthreads();
  if (lane == 0 && co < c_out) {
    float* dst = partial + (((int64_t)chunk * taps + tap) * c_out + co) * c_in;
    for (int i = 0; i < c_in; ++i) {
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < WGS_LANES; ++l) s += red[l][threadIdx.x & 63][i];   // fixed order
      dst[i] = s;
    }
  }
}

// dw[c_out][c_in][taps] = sum_chunks partial[chunk][tap][c_out][c_in]  (fixed order; the layout of nn.Conv3d.weight)
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* partial, int nchunks, int taps, int c_out, int c_in, float* dw) {
  pdl_sync();
  const int64_t per = (int64_t)c_out * c_in, count = per * taps;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < count; idx += (int64_t)gridDim.x * 256) {
    const int tap = (int)(idx / per);
    const int64_t cc = idx - (int64_t)tap * per;
    float s = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) s += __ldcg(partial + ((int64_t)ch * taps + tap) * per + cc);
    dw[cc * taps + tap] = s;
  }
}

// ---- loss and its gradient (p_losses :2355-2365): per-sample mean of l1 / l2 / smooth-l1, x_start objective clamps pred from below ----
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* pred, const float* target, int64_t count, int kind, int clamp_lo, float lo,
                                                        const float* sample_weight, float* dpred, float* loss_partial) {
  __shared__ float red[256];
  const int n = blockIdx.y, nblk = gridDim.x;
  const int64_t per = (count + nblk - 1) / nblk, i0 = (int64_t)blockIdx.x * per, i1 = min(count, i0 + per);
  const float w = sample_weight[n];   // loss weight of the sample / (batch * count)
  float s = 0.f;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
    const float p = pred[(int64_t)n * count + i], t = target[(int64_t)n * count + i];
    const bool clamped = clamp_lo && p < lo;
    const float d = (clamped ? lo : p) - t;
    float l, g;
    if (kind == 0) { l = fabsf(d); g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
    else if (kind == 1) { l = d * d; g = 2.f * d; }
    else { const float ad = fabsf(d); l = ad < 1.f ? 0.5f * d * d : ad - 0.5f; g = ad < 1.f ? d : (d > 0.f ? 1.f : -1.f); }
    s += l;
    dpred[(int64_t)n * count + i] = clamped ? 0.f : g * w;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_partial[n * nblk + blockIdx.x] = red[0];
}

// ---- Adam (torch.optim.Adam, no amsgrad; trainer.py's optimizer) + optional EMA of the parameter -----------------------------------
__global__ void __launch_bounds__(256) adam_step_kernel(float* p, const float* g, float* m, float* v, int64_t count, float lr, float b1, float b2,
                                                        float eps, float wd, float bc1, float bc2_sqrt, float grad_scale, float* ema, float ema_decay) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (int64_t)gridDim.x * 256) {
    float gi = g[i] * grad_scale;
    float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
    if (ema) ema[i] = fmaf(ema_decay, ema[i] - pi, pi);   // ema * decay + p * (1 - decay)
  }
}

}  // namespace diqt

using namespace diqt;

extern "C" int diqt_bwd_reduce(const void* x, int ld_x, const void* dz, int ld_dz, int dtype, int n, int64_t voxels, int c, const float* a, const float* b,
                               int mode, int nblk, float* partial, void* stream) {
  DIQT_REQUIRE(x && dz && partial && n > 0 && voxels > 0 && nblk > 0, "bwd_reduce: bad arguments");
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "bwd_reduce: bad dtype %d", dtype);
  DIQT_REQUIRE(mode == 0 || (mode == 1 && a && b), "bwd_reduce: mode %d (0 plain, 1 through mish(a x + b))", mode);
  const int vec = dtype == DIQT_BF16 ? 8 : 4;
  DIQT_REQUIRE(c % vec == 0 && ld_x % vec == 0 && ld_dz % vec == 0 && c / vec <= 256, "bwd_reduce: c=%d not a multiple of %d", c, vec);
  const BwMap m = bw_map(c, vec, 256);
  const size_t sh = (size_t)2 * m.lanes * c * sizeof(float);
  DIQT_REQUIRE(sh <= 48 * 1024, "bwd_reduce: c=%d too wide", c);
  const dim3 grid(nblk, n);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t vpb = bw_vpb(voxels, nblk);
#define DIQT_BW_REDUCE(T, M) \
  launch_pdl(bwd_reduce_kernel<T, M>, grid, m.threads, sh, st, (const T*)x, ld_x, (const T*)dz, ld_dz, voxels, c, m.nvec, m.lanes, vpb, a, b, partial)
  if (dtype == DIQT_BF16) { if (mode) DIQT_BW_REDUCE(__nv_bfloat16, 1); else DIQT_BW_REDUCE(__nv_bfloat16, 0); }
  else { if (mode) DIQT_BW_REDUCE(float, 1); else DIQT_BW_REDUCE(float, 0); }
#undef DIQT_BW_REDUCE
  return check_launch("bwd_reduce");
}

extern "C" int diqt_bwd_apply(const void* x, int ld_x, const void* dz, int ld_dz, const void* acc, int ld_acc, void* out, int ld_out, int dtype, int n,
                              int64_t voxels, int c, const float* a, const float* b, const float* c1, const float* c2, const float* c3, int mode,
                              int nblk, void* stream) {
  DIQT_REQUIRE(dz && out && c1 && n > 0 && voxels > 0 && nblk > 0, "bwd_apply: bad arguments");
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "bwd_apply: bad dtype %d", dtype);
  DIQT_REQUIRE(mode == 0 || (mode == 1 && a && b && x), "bwd_apply: mode %d (0 plain, 1 through mish(a x + b))", mode);
  DIQT_REQUIRE(x || !c2, "bwd_apply: c2 needs x");
  const int vec = dtype == DIQT_BF16 ? 8 : 4;
  DIQT_REQUIRE(c % vec == 0 && ld_x % vec == 0 && ld_dz % vec == 0 && ld_acc % vec == 0 && ld_out % vec == 0 && c / vec <= 256,
               "bwd_apply: c=%d not a multiple of %d", c, vec);
  const BwMap m = bw_map(c, vec, 256);
  const dim3 grid(nblk, n);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t vpb = bw_vpb(voxels, nblk);
#define DIQT_BW_APPLY(T, M)                                                                                                                        \
  launch_pdl(bwd_apply_kernel<T, M>, grid, m.threads, 0, st, (const T*)x, ld_x, (const T*)dz, ld_dz, (const T*)acc, ld_acc, (T*)out, ld_out, voxels, c, \
             m.nvec, m.lanes, vpb, a, b, c1, c2, c3)
  if (dtype == DIQT_BF16) { if (mode) DIQT_BW_APPLY(__nv_bfloat16, 1); else DIQT_BW_APPLY(__nv_bfloat16, 0); }
  else { if (mode) DIQT_BW_APPLY(float, 1); else DIQT_BW_APPLY(float, 0); }
#undef DIQT_BW_APPLY
  return check_launch("bwd_apply");
}

extern "C" int diqt_gn_bwd_finalize(const float* fwd_partial, int nblk_f, const float* bwd_partial, int nblk_b, int n, int64_t voxels, int c, int groups,
                                    float eps, const float* gamma, const float* beta, const float* film, float* c1, float* c2, float* c3, float* dgamma,
                                    float* dbeta, float* dfilm, void* stream) {
  DIQT_REQUIRE(fwd_partial && bwd_partial && gamma && beta && c1 && c2 && c3 && dgamma && dbeta && n > 0 && nblk_f > 0 && nblk_b > 0, "gn_bwd_finalize: bad arguments");
  DIQT_REQUIRE(groups > 0 && c % groups == 0 && c <= 4096, "gn_bwd_finalize: c=%d groups=%d", c, groups);
  DIQT_REQUIRE(!dfilm || film, "gn_bwd_finalize: dfilm without film");
  const int threads = 1024, parts = threads / c > 0 ? threads / c : 1;
  const size_t sh = ((size_t)parts * c * 4 + (size_t)c * 4 + (size_t)groups * 4) * sizeof(double);
  DIQT_REQUIRE(sh <= 200 * 1024, "gn_bwd_finalize: c=%d too wide", c);
  static bool attr_set = false;
  if (sh > 48 * 1024 && !attr_set) {
    DIQT_CUDA(cudaFuncSetAttribute(gn_bwd_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  launch_pdl(gn_bwd_finalize_kernel, dim3(n), dim3(threads), sh, (cudaStream_t)stream, fwd_partial, nblk_f, bwd_partial, nblk_b, n, voxels, c, groups, eps, gamma,
             beta, film, c1, c2, c3, dgamma, dbeta, dfilm);
  return check_launch("gn_bwd_finalize");
}

extern "C" int diqt_se_bwd(const float* fwd_partial, int nblk_f, const float* bwd_partial, int nblk_b, int n, int64_t voxels, int c, int hidden,
                           const float* w1, const float* w2, const float* gate, float* c3, float* dw1, float* dw2, void* stream) {
  DIQT_REQUIRE(fwd_partial && bwd_partial && w1 && w2 && gate && c3 && dw1 && dw2 && n > 0 && c > 0 && hidden > 0 && nblk_f > 0 && nblk_b > 0,
               "se_bwd: bad arguments");
  const size_t sh = ((size_t)2 * c + 2 * hidden + (size_t)2 * (512 / c > 0 ? 512 / c : 1) * c) * sizeof(float);
  DIQT_REQUIRE(sh <= 48 * 1024, "se_bwd: c=%d too wide", c);
  launch_pdl(se_bwd_kernel, dim3(n), dim3(512), sh, (cudaStream_t)stream, fwd_partial, nblk_f, bwd_partial, nblk_b, n, voxels, c, hidden, w1, w2, gate, c3, dw1,
             dw2);
  return check_launch("se_bwd");
}

static int wgrad_chunks(int64_t total_vox, int taps, int tiles) {
  int64_t ch = 1184 / ((int64_t)taps * tiles);
  if (ch > 64) ch = 64;
  const int64_t max_ch = (total_vox + WG_VT - 1) / WG_VT;
  if (ch > max_ch) ch = max_ch;
  if (ch < 1) ch = 1;
  return (int)ch;
}

namespace diqt {   // wgrad_tc.cu
bool wgrad_tc_supported(int dtype, int c_in, int c_out, int ld_x, int ld_dy, int taps);
size_t wgrad_tc_workspace_bytes(int n, int d0, int d1, int d2, int c_in, int c_out, int taps);
int wgrad_tc_run(const void* x, int ld_x, const void* dy, int ld_dy, int n, int d0, int d1, int d2, int c_in, int c_out, int taps, float* dw,
                 float* workspace, cudaStream_t st);
}  // namespace diqt

extern "C" int diqt_conv_wgrad_workspace_bytes(int n, int d0, int d1, int d2, int c_in, int c_out, int taps, size_t* bytes) {
  DIQT_REQUIRE(bytes && n > 0 && d0 > 0 && d1 > 0 && d2 > 0 && c_in > 0 && c_out > 0 && (taps == 1 || taps == 27), "conv_wgrad_workspace_bytes: bad arguments");
  const int tiles = ((c_out + 63) / 64) * ((c_in + 63) / 64);
  size_t b = (size_t)wgrad_chunks((int64_t)n * d0 * d1 * d2, taps, tiles) * taps * c_out * c_in * sizeof(float);
  if (wgrad_tc_supported(DIQT_BF16, c_in, c_out, 8, 8, taps)) b = std::max(b, wgrad_tc_workspace_bytes(n, d0, d1, d2, c_in, c_out, taps));
  *bytes = b;
  return DIQT_OK;
}

extern "C" int diqt_conv_wgrad_resolved_impl(int dtype, int c_in, int c_out, int ld_x, int ld_dy, int taps, int impl, int* resolved) {
  DIQT_REQUIRE(resolved && (impl == DIQT_IMPL_AUTO || impl == DIQT_IMPL_SIMT || impl == DIQT_IMPL_TC), "conv_wgrad_resolved_impl: bad arguments");
  const bool tc = wgrad_tc_supported(dtype, c_in, c_out, ld_x, ld_dy, taps);
  if (impl == DIQT_IMPL_TC && !tc) {
    set_error("conv_wgrad: the tcgen05 kernel needs bf16 and channel counts that are multiples of 64 (got dtype=%d c_in=%d c_out=%d)", dtype, c_in, c_out);
    return DIQT_EUNSUPPORTED;
  }
  *resolved = (impl == DIQT_IMPL_SIMT || !tc) ? DIQT_IMPL_SIMT : DIQT_IMPL_TC;
  return DIQT_OK;
}

extern "C" int diqt_conv_wgrad(const void* x, int ld_x, const void* dy, int ld_dy, int dtype, int n, int d0, int d1, int d2, int c_in, int c_out, int taps,
                               int impl, float* dw, float* workspace, void* stream) {
  DIQT_REQUIRE(x && dy && dw && workspace && n > 0 && d0 > 0 && d1 > 0 && d2 > 0 && c_in > 0 && c_out > 0, "conv_wgrad: bad arguments");
  DIQT_REQUIRE(taps == 1 || taps == 27, "conv_wgrad: taps %d (1: 1x1x1, 27: 3x3x3 with padding 1)", taps);
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "conv_wgrad: bad dtype %d", dtype);
  int resolved = 0;
  const int rc = diqt_conv_wgrad_resolved_impl(dtype, c_in, c_out, ld_x, ld_dy, taps, impl, &resolved);
  if (rc) return rc;
  if (resolved == DIQT_IMPL_TC) return wgrad_tc_run(x, ld_x, dy, ld_dy, n, d0, d1, d2, c_in, c_out, taps, dw, workspace, (cudaStream_t)stream);
  const int ci_tiles = (c_in + 63) / 64, tiles = ((c_out + 63) / 64) * ci_tiles;
  const int64_t total = (int64_t)n * d0 * d1 * d2;
  const int nchunks = wgrad_chunks(total, taps, tiles);
  int64_t vpc = (total + nchunks - 1) / nchunks;
  vpc = (vpc + WG_VT - 1) / WG_VT * WG_VT;
  const dim3 grid(nchunks, taps, tiles);
  cudaStream_t st = (cudaStream_t)stream;
  if (c_in <= 16) {
    const dim3 ngrid(nchunks, taps, (c_out + 63) / 64);
#define DIQT_WG_NARROW(T, CI)                                                                                                                         \
  launch_pdl(conv_wgrad_narrow_kernel<T, CI>, ngrid, 64 * WGS_LANES, 0, st, (const T*)x, ld_x, (const T*)dy, ld_dy, n, d0, d1, d2, c_in, c_out, taps, vpc, \
             workspace)
    if (dtype == DIQT_BF16) {
      if (c_in <= 2) DIQT_WG_NARROW(__nv_bfloat16, 2); else if (c_in <= 4) DIQT_WG_NARROW(__nv_bfloat16, 4);
      else if (c_in <= 8) DIQT_WG_NARROW(__nv_bfloat16, 8); else DIQT_WG_NARROW(__nv_bfloat16, 16);
    } else {
      if (c_in <= 2) DIQT_WG_NARROW(float, 2); else if (c_in <= 4) DIQT_WG_NARROW(float, 4);
      else if (c_in <= 8) DIQT_WG_NARROW(float, 8); else DIQT_WG_NARROW(float, 16);
    }
#undef DIQT_WG_NARROW
  } else if (dtype == DIQT_BF16)
    launch_pdl(conv_wgrad_simt_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)x, ld_x, (const __nv_bfloat16*)dy, ld_dy, n, d0, d1, d2, c_in,
               c_out, taps, ci_tiles, vpc, workspace);
  else
    launch_pdl(conv_wgrad_simt_kernel<float>, grid, 256, 0, st, (const float*)x, ld_x, (const float*)dy, ld_dy, n, d0, d1, d2, c_in, c_out, taps, ci_tiles,
               vpc, workspace);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const int64_t count = (int64_t)c_out * c_in * taps;
  launch_pdl(conv_wgrad_reduce_kernel, dim3((unsigned)((count + 255) / 256 > 1184 ? 1184 : (count + 255) / 256)), 256, 0, st, (const float*)workspace,
             nchunks, taps, c_out, c_in, dw);
  return check_launch("conv_wgrad");
}

extern "C" int diqt_loss_grad(const float* pred, const float* target, int n, int64_t count, int kind, int clamp_lo, float lo, const float* sample_weight,
                              float* dpred, float* loss_partial, int nblk, void* stream) {
  DIQT_REQUIRE(pred && target && sample_weight && dpred && loss_partial && n > 0 && count > 0 && nblk > 0, "loss_grad: bad arguments");
  DIQT_REQUIRE(kind >= 0 && kind <= 2, "loss_grad: kind %d (0 l1, 1 l2, 2 huber)", kind);
  loss_grad_kernel<<<dim3(nblk, n), 256, 0, (cudaStream_t)stream>>>(pred, target, count, kind, clamp_lo, lo, sample_weight, dpred, loss_partial);
  return check_launch("loss_grad");
}

extern "C" int diqt_adam_step(float* p, const float* g, float* m, float* v, int64_t count, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int step, float grad_scale, float* ema, float ema_decay, void* stream) {
  DIQT_REQUIRE(p && g && m && v && count > 0 && step > 0, "adam_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  const int64_t blocks = (count + 255) / 256;
  adam_step_kernel<<<(unsigned)(blocks > 1184 ? 1184 : blocks), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, count, lr, beta1, beta2, eps, weight_decay, bc1,
                                                                                                 sqrtf(bc2), grad_scale, ema, ema_decay);
  return check_launch("adam_step");
}
