// Attention blocks of the 3-D U-Net: the kernels around the 1x1x1 convolutions (which run through the
// tcgen05 / SIMT conv families) of LinearAttention / SoftMaxAttention / ChanFeedForward / ViT3D,
// /root/reference/imagen_pytorch3D.py:361-382 (LayerNorm), :858-869 (depthwise_separable_conv3d), :913-924 (Patchify),
// :926-1016 (LinearAttention), :1018-1106 (SoftMaxAttention), :1108-1116 (ChanFeedForward), :811-838 (MultiHeadAttention).
//
// Layouts.  Activations are channels-last rows [row][ld] in the engine dtype (bf16 / fp32); all arithmetic is fp32.
// "native" rows follow the engine (n sub-volumes of h^3 voxels, or already merged in boundary mode); the attention blocks of the
// reference see the f^3 sub-volumes MERGED into one (f*h)^3 volume (utils_mine.py:44-67, imagen_pytorch3D.py:1613-1617), so the two
// kernels at the edge of a block (patchify on the way in, the last ChanLayerNorm on the way out) translate rows with SubGeom.
// Token tensors ([tokens][channels], token = (tz*g + ty)*g + tx over the merged volume) always use the merged order.
//
// The N x N softmax attention has two kernels: the tcgen05 one in attn_tc.cu (bf16, head dim 64) and the fp32 CUDA-core one below
// (flash-style, no N x N tensor in memory) for the other head dims and the fp32 exact mode.
#include "common.cuh"

namespace diqt {

// merged row (x fastest over a (f*h)^3 volume) -> native row (sub-volume b = zb + f*yb + f*f*xb, utils_mine.py:25-67)
__device__ __forceinline__ int64_t merged_to_native(const SubGeom g, int64_t m) {
  if (g.f <= 1) return m;
  const int h = g.h, fh = g.f * g.h;
  const int X = (int)(m % fh), Y = (int)((m / fh) % fh), Z = (int)(m / ((int64_t)fh * fh));
  const int zb = Z / h, yb = Y / h, xb = X / h;
  const int b = zb + g.f * yb + g.f * g.f * xb;
  return (int64_t)b * h * h * h + ((int64_t)(Z - zb * h) * h + (Y - yb * h)) * h + (X - xb * h);
}
// native row -> merged row
__device__ __forceinline__ int64_t native_to_merged(const SubGeom g, int64_t r) {
  if (g.f <= 1) return r;
  const int64_t vox = (int64_t)g.h * g.h * g.h;
  return sub_row(g, vox, (int)(r / vox), r % vox);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <int kAct>
__device__ __forceinline__ float apply_act(float x) {
  if (kAct == 1) return mish<false>(x);
  if (kAct == 2) return gelu_erf(x);
  return x;
}
__device__ __forceinline__ float apply_act_rt(float x, int act) { return act == 1 ? mish<false>(x) : act == 2 ? gelu_erf(x) : x; }

// ---------------------------------------------------------------------------------------------------------------------
// Channel LayerNorm over the C entries of every row (LayerNorm(dim=-4) :361-382: biased variance, eps inside the sqrt, scale g, no
// bias; nn.LayerNorm for the ViT with beta).  out[r] = LN(act(x[xrow(r)])) * g (+ beta) (+ res1[r]) (+ res2[r]).  One warp per row.
// x_merged: x is stored in merged order while out / res are native (the last norm of a block) -> xrow = native_to_merged.
template <typename T>
__global__ void __launch_bounds__(256) chan_ln_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo, int64_t rows, int c,
                                                      const float* __restrict__ g, const float* __restrict__ beta, float eps, int pre_act,
                                                      const T* __restrict__ res1, int ld1, const T* __restrict__ res2, int ld2, SubGeom xmap) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < rows; r += nwarps) {
    const T* xr = x + native_to_merged(xmap, r) * ldx;
    float s = 0.f;
    for (int ch = lane; ch < c; ch += 32) s += apply_act_rt(to_float(xr[ch]), pre_act);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)c;
    float q = 0.f;
    for (int ch = lane; ch < c; ch += 32) {
      const float d = apply_act_rt(to_float(xr[ch]), pre_act) - mean;
      q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.f / sqrtf(q / (float)c + eps);
    __syncwarp();
    T* orow = out + r * ldo;
    for (int ch = lane; ch < c; ch += 32) {
      float y = (apply_act_rt(to_float(xr[ch]), pre_act) - mean) * rstd * g[ch];
      if (beta) y += beta[ch];
      if (res1) y += to_float(res1[r * ld1 + ch]);
      if (res2) y += to_float(res2[r * ld2 + ch]);
      orow[ch] = from_float<T>(y);
    }
  }
}

// out[r][ch] = act(a[r][ch]) (+ b[r][ch]) (+ c2[r][ch]): the residual adds around attention / feed-forward (:1148-1149, :1622) and
// the activations that have no producer kernel to live in.
template <typename T>
__global__ void __launch_bounds__(256) rows_combine_kernel(const T* __restrict__ a, int lda, int act, const T* __restrict__ b, int ldb,
                                                           const T* __restrict__ c2, int ldc, T* __restrict__ out, int ldo, int64_t rows, int c) {
  const int64_t total = rows * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c;
    const int ch = (int)(i - r * c);
    float y = apply_act_rt(to_float(a[r * lda + ch]), act);
    if (b) y += to_float(b[r * ldb + ch]);
    if (c2) y += to_float(c2[r * ldc + ch]);
    out[r * ldo + ch] = from_float<T>(y);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Depthwise p^3 / stride p convolution of Patchify / PatchEmbedding (:847, :919): tokens[t][ch] = bias[ch] + sum_taps w[tap][ch] *
// x[voxel(t, tap)][ch].  One CTA per token: 32 channel lanes x 8 tap slices, slices reduced in a fixed order through shared memory.
// x is native (xmap translates), w is [p^3][c] fp32 (tap = (dz*p + dy)*p + dx).
template <typename T>
__global__ void __launch_bounds__(256) dw_patchify_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo, int gdim, int p, int c,
                                                          const float* __restrict__ w, const float* __restrict__ bias, SubGeom xmap) {
  __shared__ float red[8][33];
  const int t = blockIdx.x;
  const int tx = t % gdim, ty = (t / gdim) % gdim, tz = t / (gdim * gdim);
  const int G = gdim * p, taps = p * p * p;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  for (int c0 = 0; c0 < c; c0 += 32) {
    const int ch = c0 + lane;
    float acc = 0.f;
    if (ch < c) {
      for (int tap = slice; tap < taps; tap += 8) {
        const int dx = tap % p, dy = (tap / p) % p, dz = tap / (p * p);
        const int64_t m = ((int64_t)(tz * p + dz) * G + (ty * p + dy)) * G + (tx * p + dx);
        acc = fmaf(w[(size_t)tap * c + ch], to_float(x[merged_to_native(xmap, m) * ldx + ch]), acc);
      }
    }
    red[slice][lane] = acc;
    __syncthreads();
    if (slice == 0 && ch < c) {
      float s = bias ? bias[ch] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i][lane];
      out[(size_t)t * ldo + ch] = from_float<T>(s);
    }
    __syncthreads();
  }
}

// Depthwise 3x3x3, stride 1, zero padding 1 over one channels-last volume (d0, d1, d2) (:963, :969, :975 to_q/k/v.2 without bias;
// :955 reconstruct.1.depthwise with bias).  w: [27][c] fp32, tap = (dz*3 + dy)*3 + dx.
template <typename T>
__global__ void __launch_bounds__(256) dw_conv3_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo, int d0, int d1, int d2, int c,
                                                       const float* __restrict__ w, const float* __restrict__ bias) {
  const int64_t total = (int64_t)d0 * d1 * d2 * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / c;
    const int ch = (int)(i - v * c);
    const int X = (int)(v % d2), Y = (int)((v / d2) % d1), Z = (int)(v / ((int64_t)d2 * d1));
    float acc = bias ? bias[ch] : 0.f;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
      const int z = Z + dz;
      if (z < 0 || z >= d0) continue;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = Y + dy;
        if (y < 0 || y >= d1) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = X + dx;
          if (xx < 0 || xx >= d2) continue;
          const int tap = ((dz + 1) * 3 + (dy + 1)) * 3 + (dx + 1);
          acc = fmaf(w[tap * c + ch], to_float(x[(((int64_t)z * d1 + y) * d2 + xx) * ldx + ch]), acc);
        }
      }
    }
    out[v * ldo + ch] = from_float<T>(acc);
  }
}

// nn.Upsample(scale_factor=p, mode='trilinear', align_corners=True) (:900, :954) of a token volume g^3 to (g*p)^3, channels-last.
// Source index = dst * (g-1)/(G-1) in fp32, i1 = i0 + (i0 < g-1), weights (1-l, l): ATen's area_pixel_compute_source_index.
template <typename T>
__global__ void __launch_bounds__(256) upsample_trilinear_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo, int gdim, int p, int c) {
  const int G = gdim * p;
  const float scale = G > 1 ? (float)(gdim - 1) / (float)(G - 1) : 0.f;
  const int64_t total = (int64_t)G * G * G * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / c;
    const int ch = (int)(i - v * c);
    const int X = (int)(v % G), Y = (int)((v / G) % G), Z = (int)(v / ((int64_t)G * G));
    int i0[3], i1[3];
    float l0[3], l1[3];
    const int pos[3] = {Z, Y, X};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float src = scale * (float)pos[a];
      int k = (int)src;
      if (k > gdim - 1) k = gdim - 1;
      i0[a] = k;
      i1[a] = k + (k < gdim - 1 ? 1 : 0);
      float l = src - (float)k;
      l = fminf(fmaxf(l, 0.f), 1.f);
      l1[a] = l;
      l0[a] = 1.f - l;
    }
    auto at = [&](int z, int y, int xx) { return to_float(x[(((int64_t)z * gdim + y) * gdim + xx) * ldx + ch]); };
    const float a00 = l0[2] * at(i0[0], i0[1], i0[2]) + l1[2] * at(i0[0], i0[1], i1[2]);
    const float a01 = l0[2] * at(i0[0], i1[1], i0[2]) + l1[2] * at(i0[0], i1[1], i1[2]);
    const float a10 = l0[2] * at(i1[0], i0[1], i0[2]) + l1[2] * at(i1[0], i0[1], i1[2]);
    const float a11 = l0[2] * at(i1[0], i1[1], i0[2]) + l1[2] * at(i1[0], i1[1], i1[2]);
    const float y = l0[0] * (l0[1] * a00 + l1[1] * a01) + l1[0] * (l0[1] * a10 + l1[1] * a11);
    out[v * ldo + ch] = from_float<T>(y);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Linear attention (:1001-1008): q = softmax_d(q) * scale, k = softmax_n(k), ctx = k^T v per head, out = q ctx, then Mish (:1011).
// (1) column statistics of k over the tokens: stat[col] = (max_n k, sum_n exp(k - max)).  CTA = 32 columns x 32 row slices.
template <typename T>
__global__ void __launch_bounds__(1024) col_softmax_stats_kernel(const T* __restrict__ k, int ld, int ntok, int cols, float* __restrict__ stat) {
  __shared__ float red[32][33];
  const int lane = threadIdx.x, slice = threadIdx.y;
  const int col = blockIdx.x * 32 + lane;
  float m = -INFINITY;
  if (col < cols)
    for (int n = slice; n < ntok; n += 32) m = fmaxf(m, to_float(k[(size_t)n * ld + col]));
  red[slice][lane] = m;
  __syncthreads();
  m = red[0][lane];
#pragma unroll
  for (int i = 1; i < 32; ++i) m = fmaxf(m, red[i][lane]);
  __syncthreads();
  float s = 0.f;
  if (col < cols)
    for (int n = slice; n < ntok; n += 32) s += expf(to_float(k[(size_t)n * ld + col]) - m);
  red[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][lane];
    stat[2 * col] = m;
    stat[2 * col + 1] = t;
  }
}

// (2) partial[chunk][head][d][e] = sum_{n in chunk} exp(k[n][head,d] - max[head,d]) * v[n][head,e]   (normalised by the consumer).
// grid (chunks, heads), 256 threads; dh in {16, 32, 64}; thread (td, te) owns outputs d = td + 16 i, e = te + 16 j.
constexpr int kCtxTile = 32;
template <typename T>
__global__ void __launch_bounds__(256) linattn_ctx_kernel(const T* __restrict__ k, const T* __restrict__ v, int ld, int ntok, int dh,
                                                          const float* __restrict__ stat, float* __restrict__ partial, int heads) {
  __shared__ float ks[kCtxTile][64], vs[kCtxTile][64];
  const int chunk = blockIdx.x, head = blockIdx.y, nchunks = gridDim.x;
  const int per = (ntok + nchunks - 1) / nchunks;
  const int n0 = chunk * per, n1 = min(ntok, n0 + per);
  const int td = threadIdx.x >> 4, te = threadIdx.x & 15, reps = dh >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int base = n0; base < n1; base += kCtxTile) {
    const int cnt = min(kCtxTile, n1 - base);
    for (int idx = threadIdx.x; idx < kCtxTile * dh; idx += 256) {
      const int r = idx / dh, d = idx - r * dh;
      float kv = 0.f, vv = 0.f;
      if (r < cnt) {
        const size_t off = (size_t)(base + r) * ld + head * dh + d;
        kv = expf(to_float(k[off]) - stat[2 * (head * dh + d)]);
        vv = to_float(v[off]);
      }
      ks[r][d] = kv;
      vs[r][d] = vv;
    }
    __syncthreads();
    for (int r = 0; r < kCtxTile; ++r) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < reps) {
          const float kd = ks[r][td + 16 * i];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < reps) acc[i][j] = fmaf(kd, vs[r][te + 16 * j], acc[i][j]);
        }
      }
    }
    __syncthreads();
  }
  float* dst = partial + ((size_t)chunk * heads + head) * dh * dh;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i < reps && j < reps) dst[(td + 16 * i) * dh + te + 16 * j] = acc[i][j];
}

// (3) out[n][head,e] = mish( sum_d softmax_d(q[n][head,:])[d] * scale * ctx[head][d][e] ), ctx = sum_chunks partial / sum_n exp(k).
// grid (token tiles of 64, heads), 256 threads.
template <typename T>
__global__ void __launch_bounds__(256) linattn_out_kernel(const T* __restrict__ q, int ldq, T* __restrict__ out, int ldo, int ntok, int dh, int heads,
                                                          const float* __restrict__ stat, const float* __restrict__ partial, int nchunks, float scale,
                                                          int act) {
  __shared__ float ctx[64][65];
  __shared__ float qs[64][65];
  const int tile = blockIdx.x, head = blockIdx.y;
  for (int idx = threadIdx.x; idx < dh * dh; idx += 256) {
    const int d = idx / dh, e = idx - d * dh;
    float s = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) s += partial[((size_t)ch * heads + head) * dh * dh + idx];  // fixed order
    ctx[d][e] = s / stat[2 * (head * dh + d) + 1];
  }
  const int n0 = tile * 64;
  for (int idx = threadIdx.x; idx < 64 * dh; idx += 256) {
    const int r = idx / dh, d = idx - r * dh;
    qs[r][d] = (n0 + r < ntok) ? to_float(q[(size_t)(n0 + r) * ldq + head * dh + d]) : 0.f;
  }
  __syncthreads();
  if (threadIdx.x < 64) {  // softmax over the head dimension, one token per thread
    const int r = threadIdx.x;
    float m = -INFINITY;
    for (int d = 0; d < dh; ++d) m = fmaxf(m, qs[r][d]);
    float s = 0.f;
    for (int d = 0; d < dh; ++d) {
      const float e = expf(qs[r][d] - m);
      qs[r][d] = e;
      s += e;
    }
    const float inv = scale / s;
    for (int d = 0; d < dh; ++d) qs[r][d] *= inv;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 64 * dh; idx += 256) {
    const int r = idx / dh, e = idx - r * dh;
    if (n0 + r >= ntok) continue;
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(qs[r][d], ctx[d][e], acc);
    out[(size_t)(n0 + r) * ldo + head * dh + e] = from_float<T>(apply_act_rt(acc, act));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Softmax attention (:1087-1097, MultiHeadAttention :826-836): out = softmax_k(q k^T * scale) v per head, flash style (online softmax
// over key tiles of 32, nothing N x N in memory), optional Mish (:1100).  CTA = 32 queries x 4 lanes; lane l of a quad owns head
// dims l, l+4, ...   grid (ceil(N/32), heads).
template <typename T, int DH>
__global__ void __launch_bounds__(128) softmax_attn_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int ldq, int ldk,
                                                           int ldv, T* __restrict__ out, int ldo, int ntok, float scale, int act) {
  constexpr int PER = DH / 4, KT = 32;
  __shared__ float ks[KT][DH], vs[KT][DH];
  const int head = blockIdx.y;
  const int qi = blockIdx.x * 32 + (threadIdx.x >> 2), ql = threadIdx.x & 3;
  const bool live = qi < ntok;
  float qr[PER], acc[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    qr[i] = live ? to_float(q[(size_t)qi * ldq + head * DH + ql + 4 * i]) * scale : 0.f;
    acc[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int base = 0; base < ntok; base += KT) {
    const int cnt = min(KT, ntok - base);
    __syncthreads();
    for (int idx = threadIdx.x; idx < KT * DH; idx += 128) {
      const int r = idx / DH, d = idx - r * DH;
      const bool ok = r < cnt;
      ks[r][d] = ok ? to_float(k[(size_t)(base + r) * ldk + head * DH + d]) : 0.f;
      vs[r][d] = ok ? to_float(v[(size_t)(base + r) * ldv + head * DH + d]) : 0.f;
    }
    __syncthreads();
    float s[KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) d = fmaf(qr[i], ks[j][ql + 4 * i], d);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      s[j] = j < cnt ? d : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float mn = fmaxf(m, tmax);
    const float corr = expf(m - mn);  // first tile: exp(-inf) = 0
    l *= corr;
#pragma unroll
    for (int i = 0; i < PER; ++i) acc[i] *= corr;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float pj = expf(s[j] - mn);  // masked keys: exp(-inf) = 0
      l += pj;
#pragma unroll
      for (int i = 0; i < PER; ++i) acc[i] = fmaf(pj, vs[j][ql + 4 * i], acc[i]);
    }
    m = mn;
  }
  if (live) {
    const float inv = 1.f / l;
#pragma unroll
    for (int i = 0; i < PER; ++i) out[(size_t)qi * ldo + head * DH + ql + 4 * i] = from_float<T>(apply_act_rt(acc[i] * inv, act));
  }
}

static inline unsigned grid_for(int64_t work, int threads) {
  int64_t b = (work + threads - 1) / threads;
  if (b > 148 * 32) b = 148 * 32;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace diqt

using namespace diqt;

extern "C" int diqt_chan_layernorm(const void* x, int ld_x, void* out, int ld_out, int dtype, int64_t rows, int c, const float* g,
                                   const float* beta, float eps, int pre_act, const void* res1, int ld_res1, const void* res2, int ld_res2,
                                   int x_sub_f, int x_sub_h, void* stream) {
  DIQT_REQUIRE(x && out && g && rows > 0 && c > 0, "chan_layernorm: bad arguments");
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "chan_layernorm: bad dtype %d", dtype);
  DIQT_REQUIRE(pre_act >= 0 && pre_act <= 2, "chan_layernorm: bad activation %d", pre_act);
  DIQT_REQUIRE(ld_x >= c && ld_out >= c, "chan_layernorm: pitch smaller than c=%d", c);
  if (x_sub_f > 1) DIQT_REQUIRE(rows == (int64_t)x_sub_f * x_sub_f * x_sub_f * x_sub_h * x_sub_h * x_sub_h, "chan_layernorm: rows do not match the sub-volume geometry");
  cudaStream_t st = (cudaStream_t)stream;
  const SubGeom xm = {x_sub_f, x_sub_h};
  const unsigned grid = grid_for(rows, 8);
  if (dtype == DIQT_BF16)
    chan_ln_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ld_x, (__nv_bfloat16*)out, ld_out, rows, c, g, beta, eps, pre_act,
                                                       (const __nv_bfloat16*)res1, ld_res1, (const __nv_bfloat16*)res2, ld_res2, xm);
  else
    chan_ln_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ld_x, (float*)out, ld_out, rows, c, g, beta, eps, pre_act, (const float*)res1, ld_res1,
                                               (const float*)res2, ld_res2, xm);
  return check_launch("chan_layernorm");
}

extern "C" int diqt_rows_combine(const void* a, int ld_a, int act, const void* b, int ld_b, const void* c2, int ld_c, void* out, int ld_out, int dtype,
                                 int64_t rows, int c, void* stream) {
  DIQT_REQUIRE(a && out && rows > 0 && c > 0, "rows_combine: bad arguments");
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "rows_combine: bad dtype %d", dtype);
  DIQT_REQUIRE(act >= 0 && act <= 2, "rows_combine: bad activation %d", act);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(rows * c, 256);
  if (dtype == DIQT_BF16)
    rows_combine_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, ld_a, act, (const __nv_bfloat16*)b, ld_b, (const __nv_bfloat16*)c2,
                                                            ld_c, (__nv_bfloat16*)out, ld_out, rows, c);
  else
    rows_combine_kernel<float><<<grid, 256, 0, st>>>((const float*)a, ld_a, act, (const float*)b, ld_b, (const float*)c2, ld_c, (float*)out, ld_out, rows, c);
  return check_launch("rows_combine");
}

extern "C" int diqt_dw_patchify(const void* x, int ld_x, void* tokens, int ld_t, int dtype, int grid_dim, int patch, int c, const float* w,
                                const float* bias, int x_sub_f, int x_sub_h, void* stream) {
  DIQT_REQUIRE(x && tokens && w && grid_dim > 0 && patch > 0 && c > 0, "dw_patchify: bad arguments");
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "dw_patchify: bad dtype %d", dtype);
  if (x_sub_f > 1)
    DIQT_REQUIRE(x_sub_f * x_sub_h == grid_dim * patch && x_sub_h % patch == 0, "dw_patchify: sub-volume side %d x %d does not tile into %d patches of %d", x_sub_f,
                 x_sub_h, grid_dim, patch);
  cudaStream_t st = (cudaStream_t)stream;
  const SubGeom xm = {x_sub_f, x_sub_h};
  const unsigned grid = (unsigned)(grid_dim * grid_dim * grid_dim);
  if (dtype == DIQT_BF16)
    dw_patchify_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ld_x, (__nv_bfloat16*)tokens, ld_t, grid_dim, patch, c, w, bias, xm);
  else
    dw_patchify_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ld_x, (float*)tokens, ld_t, grid_dim, patch, c, w, bias, xm);
  return check_launch("dw_patchify");
}

extern "C" int diqt_dw_conv3(const void* x, int ld_x, void* out, int ld_out, int dtype, int d0, int d1, int d2, int c, const float* w, const float* bias,
                             void* stream) {
  DIQT_REQUIRE(x && out && w && d0 > 0 && d1 > 0 && d2 > 0 && c > 0 && x != out, "dw_conv3: bad arguments");
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "dw_conv3: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for((int64_t)d0 * d1 * d2 * c, 256);
  if (dtype == DIQT_BF16)
    dw_conv3_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ld_x, (__nv_bfloat16*)out, ld_out, d0, d1, d2, c, w, bias);
  else
    dw_conv3_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ld_x, (float*)out, ld_out, d0, d1, d2, c, w, bias);
  return check_launch("dw_conv3");
}

extern "C" int diqt_upsample_trilinear(const void* tokens, int ld_t, void* out, int ld_out, int dtype, int grid_dim, int factor, int c, void* stream) {
  DIQT_REQUIRE(tokens && out && grid_dim > 0 && factor > 0 && c > 0, "upsample_trilinear: bad arguments");
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "upsample_trilinear: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t G = (int64_t)grid_dim * factor;
  const unsigned grid = grid_for(G * G * G * c, 256);
  if (dtype == DIQT_BF16)
    upsample_trilinear_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)tokens, ld_t, (__nv_bfloat16*)out, ld_out, grid_dim, factor, c);
  else
    upsample_trilinear_kernel<float><<<grid, 256, 0, st>>>((const float*)tokens, ld_t, (float*)out, ld_out, grid_dim, factor, c);
  return check_launch("upsample_trilinear");
}

extern "C" int diqt_linear_attention_chunks(int tokens, int* chunks) {
  DIQT_REQUIRE(chunks && tokens > 0, "linear_attention_chunks: bad arguments");
  int c = (tokens + 127) / 128;
  *chunks = c > 32 ? 32 : c;
  return DIQT_OK;
}

extern "C" int diqt_linear_attention(const void* q, const void* k, const void* v, int ld_qkv, void* out, int ld_out, int dtype, int tokens, int heads,
                                     int dim_head, float scale, int act, float* col_stat, float* partial, void* stream) {
  DIQT_REQUIRE(q && k && v && out && col_stat && partial && tokens > 0 && heads > 0, "linear_attention: bad arguments");
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "linear_attention: bad dtype %d", dtype);
  DIQT_REQUIRE(dim_head == 16 || dim_head == 32 || dim_head == 64, "linear_attention: dim_head=%d (16, 32 or 64)", dim_head);
  cudaStream_t st = (cudaStream_t)stream;
  const int cols = heads * dim_head;
  int chunks = 0;
  diqt_linear_attention_chunks(tokens, &chunks);
  const dim3 sblock(32, 32), cgrid(chunks, heads), ogrid((tokens + 63) / 64, heads);
  if (dtype == DIQT_BF16) {
    col_softmax_stats_kernel<__nv_bfloat16><<<(cols + 31) / 32, sblock, 0, st>>>((const __nv_bfloat16*)k, ld_qkv, tokens, cols, col_stat);
    linattn_ctx_kernel<__nv_bfloat16><<<cgrid, 256, 0, st>>>((const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ld_qkv, tokens, dim_head, col_stat, partial, heads);
    linattn_out_kernel<__nv_bfloat16><<<ogrid, 256, 0, st>>>((const __nv_bfloat16*)q, ld_qkv, (__nv_bfloat16*)out, ld_out, tokens, dim_head, heads, col_stat,
                                                            partial, chunks, scale, act);
  } else {
    col_softmax_stats_kernel<float><<<(cols + 31) / 32, sblock, 0, st>>>((const float*)k, ld_qkv, tokens, cols, col_stat);
    linattn_ctx_kernel<float><<<cgrid, 256, 0, st>>>((const float*)k, (const float*)v, ld_qkv, tokens, dim_head, col_stat, partial, heads);
    linattn_out_kernel<float><<<ogrid, 256, 0, st>>>((const float*)q, ld_qkv, (float*)out, ld_out, tokens, dim_head, heads, col_stat, partial, chunks, scale,
                                                    act);
  }
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return check_launch("linear_attention");
}

template <typename T>
static void launch_softmax_attn(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, void* out, int ldo, int tokens, int heads, int dh,
                                float scale, int act, cudaStream_t st) {
  const dim3 grid((tokens + 31) / 32, heads);
  if (dh == 16)
    softmax_attn_kernel<T, 16><<<grid, 128, 0, st>>>((const T*)q, (const T*)k, (const T*)v, ldq, ldk, ldv, (T*)out, ldo, tokens, scale, act);
  else if (dh == 32)
    softmax_attn_kernel<T, 32><<<grid, 128, 0, st>>>((const T*)q, (const T*)k, (const T*)v, ldq, ldk, ldv, (T*)out, ldo, tokens, scale, act);
  else
    softmax_attn_kernel<T, 64><<<grid, 128, 0, st>>>((const T*)q, (const T*)k, (const T*)v, ldq, ldk, ldv, (T*)out, ldo, tokens, scale, act);
}

extern "C" int diqt_softmax_attention(const void* q, const void* k, const void* v, int ld_q, int ld_k, int ld_v, void* out, int ld_out, int dtype, int tokens,
                                      int heads, int dim_head, float scale, int act, void* stream) {
  DIQT_REQUIRE(q && k && v && out && tokens > 0 && heads > 0, "softmax_attention: bad arguments");
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "softmax_attention: bad dtype %d", dtype);
  DIQT_REQUIRE(dim_head == 16 || dim_head == 32 || dim_head == 64, "softmax_attention: dim_head=%d (16, 32 or 64)", dim_head);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DIQT_BF16)
    launch_softmax_attn<__nv_bfloat16>(q, k, v, ld_q, ld_k, ld_v, out, ld_out, tokens, heads, dim_head, scale, act, st);
  else
    launch_softmax_attn<float>(q, k, v, ld_q, ld_k, ld_v, out, ld_out, tokens, heads, dim_head, scale, act, st);
  return check_launch("softmax_attention");
}
