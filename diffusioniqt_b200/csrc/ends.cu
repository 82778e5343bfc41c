// Network ends and sampler glue: init_conv over fp32 planes, final 1x1x1 conv fused with the
// DDPM posterior update, the time-conditioning dense layers, and the step counter.
//
// Reference call sites replaced: init_conv (imagen_pytorch3D.py:1291, 1576-1589), final_conv
// (:1477, 1682), p_mean_variance / p_sample / q_posterior (:1976-2056, :290-309),
// LearnedSinusoidalPosEmb + to_time_hiddens + to_time_cond + ResnetBlock.time_mlp
// (:518-533, :1305-1316, :586-589, :603-605).
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace diqt {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("DIQT_DISABLE_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- init conv: up to 8 single-channel fp32 planes -> channels-last activations -----------------
constexpr int kInitMaxCin = 8;
struct InitPlanes {
  const float* p[kInitMaxCin];
  long long stride[kInitMaxCin];
};

// element offset of merged-volume voxel (b, zz, yy, xx) inside a (n, d0, d1, d2) fp32 plane; in boundary mode (sg.f > 1) the
// planes hold f^3 separate sub-volumes of side h while the conv runs over the merged volume
__device__ __forceinline__ int64_t plane_offset(const SubGeom sg, int b, int zz, int yy, int xx, int d1, int d2, long long stride) {
  if (sg.f <= 1) return (int64_t)b * stride + ((int64_t)zz * d1 + yy) * d2 + xx;
  const int h = sg.h;
  const int sb = zz / h + sg.f * (yy / h) + sg.f * sg.f * (xx / h);
  return (int64_t)sb * stride + ((int64_t)(zz % h) * h + yy % h) * h + xx % h;
}

// one thread = one output voxel; CO_T output channels at a time are held in registers
template <typename T, int CO_T>
__global__ void __launch_bounds__(128)
init_conv_kernel(InitPlanes planes, int c_in, const float* __restrict__ w /*[27][c_in][c_out]*/,
                 const float* __restrict__ bias, T* __restrict__ out, int ld_out, int n, int d0, int d1, int d2,
                 int c_out, SubGeom sg) {
  extern __shared__ float sw[];  // weights [27*c_in][c_out] + bias[c_out]
  const int wcount = 27 * c_in * c_out;
  for (int i = threadIdx.x; i < wcount; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < c_out; i += blockDim.x) sw[wcount + i] = bias[i];
  __syncthreads();
  const int64_t vol = (int64_t)d0 * d1 * d2;
  const int64_t total = (int64_t)n * vol;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int b = (int)(idx / vol);
  int64_t r = idx - (int64_t)b * vol;
  const int z = (int)(r / ((int64_t)d1 * d2));
  r -= (int64_t)z * d1 * d2;
  const int y = (int)(r / d2);
  const int x = (int)(r - (int64_t)y * d2);
  T* orow = out + idx * ld_out;
  for (int co0 = 0; co0 < c_out; co0 += CO_T) {
    float acc[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[j] = sw[wcount + co0 + j];
#pragma unroll 1
    for (int t = 0; t < 27; ++t) {
      const int zz = z + t / 9 - 1, yy = y + (t / 3) % 3 - 1, xx = x + t % 3 - 1;
      if (zz < 0 || zz >= d0 || yy < 0 || yy >= d1 || xx < 0 || xx >= d2) continue;  // zero padding
      for (int ci = 0; ci < c_in; ++ci) {
        const float v = __ldg(planes.p[ci] + plane_offset(sg, b, zz, yy, xx, d1, d2, planes.stride[ci]));
        const float* wr = sw + (t * c_in + ci) * c_out + co0;
#pragma unroll
        for (int j = 0; j < CO_T; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
      }
    }
    constexpr int VN = Vec<T>::N;
#pragma unroll
    for (int j = 0; j < CO_T; j += VN) {
      Vec<T> o;
#pragma unroll
      for (int i = 0; i < VN; ++i) o.v[i] = acc[j + i];
      o.store(orow + co0 + j);
    }
  }
}

// four consecutive voxels along d2 per thread, CO_T output channels per pass: every weight vector read from shared memory
// feeds 4 x as many FMAs as in the one-voxel kernel (which was bound by broadcast LDS traffic)
template <typename T, int CO_T>
__global__ void __launch_bounds__(128)
init_conv_x4_kernel(InitPlanes planes, int c_in, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out,
                    int ld_out, int n, int d0, int d1, int d2, int c_out, SubGeom sg) {
  extern __shared__ float sw[];
  const int wcount = 27 * c_in * c_out;
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < wcount; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < c_out; i += blockDim.x) sw[wcount + i] = bias[i];
  __syncthreads();
  pdl_wait();
  const int xq = d2 / 4;
  const int64_t total = (int64_t)n * d0 * d1 * xq;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int64_t r = idx;
  const int x = (int)(r % xq) * 4; r /= xq;
  const int y = (int)(r % d1); r /= d1;
  const int z = (int)(r % d0);
  const int b = (int)(r / d0);
  const int64_t vox = (((int64_t)b * d0 + z) * d1 + y) * d2 + x;
  for (int co0 = 0; co0 < c_out; co0 += CO_T) {
    float acc[4][CO_T];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int j = 0; j < CO_T; ++j) acc[v][j] = sw[wcount + co0 + j];
#pragma unroll 1
    for (int t9 = 0; t9 < 9; ++t9) {
      const int zz = z + t9 / 3 - 1, yy = y + t9 % 3 - 1;
      if (zz < 0 || zz >= d0 || yy < 0 || yy >= d1) continue;
      for (int ci = 0; ci < c_in; ++ci) {
        float in[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int xx = x - 1 + k;
          in[k] = (xx >= 0 && xx < d2) ? __ldg(planes.p[ci] + plane_offset(sg, b, zz, yy, xx, d1, d2, planes.stride[ci])) : 0.f;
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float* wr = sw + ((t9 * 3 + dx) * c_in + ci) * c_out + co0;
#pragma unroll
          for (int j = 0; j < CO_T; j += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(wr + j);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              acc[v][j] = fmaf(in[v + dx], wv.x, acc[v][j]);
              acc[v][j + 1] = fmaf(in[v + dx], wv.y, acc[v][j + 1]);
              acc[v][j + 2] = fmaf(in[v + dx], wv.z, acc[v][j + 2]);
              acc[v][j + 3] = fmaf(in[v + dx], wv.w, acc[v][j + 3]);
            }
          }
        }
      }
    }
    constexpr int VN = Vec<T>::N;
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int j = 0; j < CO_T; j += VN) {
        Vec<T> o;
#pragma unroll
        for (int i = 0; i < VN; ++i) o.v[i] = acc[v][j + i];
        o.store(out + (vox + v) * ld_out + co0 + j);
      }
  }
}

__global__ void init_conv_pack_kernel(const float* __restrict__ w, int c_out, int c_in, float* __restrict__ packed) {
  // (c_out, c_in, 3,3,3) -> [tap][c_in][c_out]
  const int total = 27 * c_in * c_out;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % c_out, ci = (i / c_out) % c_in, t = i / (c_out * c_in);
    packed[i] = w[((int64_t)co * c_in + ci) * 27 + t];
  }
}

// ---- final 1x1x1 conv (+ fused DDPM update) -------------------------------------------------------
// sched row: (alpha, sigma, c, alpha_next, noise_scale, lo, hi, objective)
struct StepConsts {
  float alpha, sigma, c, alpha_next, noise_scale, lo, hi;
  int objective;  // 0 x_start, 1 noise, 2 v
};

__device__ __forceinline__ StepConsts load_step(const float* __restrict__ sched, const int32_t* __restrict__ step) {
  const float* r = sched + (int64_t)(step ? __ldcg(step) : 0) * 8;
  StepConsts s;
  s.alpha = r[0]; s.sigma = r[1]; s.c = r[2]; s.alpha_next = r[3]; s.noise_scale = r[4]; s.lo = r[5]; s.hi = r[6];
  s.objective = (int)r[7];
  return s;
}

// torch.clamp propagates NaN (fminf / fmaxf drop it): a diverged prediction must stay visible in x0 and in the stitched volume
__device__ __forceinline__ float clamp_nan(float x, float lo, float hi) { return x != x ? x : fminf(fmaxf(x, lo), hi); }

// x_start from the network output (imagen_pytorch3D.py:1996-2004), static clamp (:2022-2026), posterior mean
// (:301-302) and the sampling line (:2055), in the reference's operation order.
__device__ __forceinline__ void ddpm_point(const StepConsts& s, float pred, float xt, float eps, float& x_next, float& x0) {
  float xs = pred;
  if (s.objective == 1) xs = (xt - s.sigma * pred) / fmaxf(s.alpha, 1e-8f);
  else if (s.objective == 2) xs = s.alpha * xt - s.sigma * pred;
  xs = clamp_nan(xs, s.lo, s.hi);
  const float mean = s.alpha_next * (xt * (1.f - s.c) / s.alpha + s.c * xs);
  x_next = mean + s.noise_scale * eps;
  x0 = xs;
}

// ---- Elucidated (Karras et al.) sampler step, elucidated_imagen.py:329-358 (preconditioned network) and :468-519 -----
// table row per U-Net forward (16 floats): 0 S_noise, 1 sqrt(sigma_hat^2 - sigma^2), 2 c_in(sigma of this forward), 3 c_skip,
// 4 c_out, 5 divisor sigma (sigma_hat for the Euler pass, sigma_next for the Heun pass), 6 (sigma_next - sigma_hat),
// 7 0.5 * (sigma_next - sigma_hat), 8 c_in(sigma_next), 9 clamp lo, 10 clamp hi
constexpr int kEdmRow = 16;
struct EdmConsts {
  float s_noise, noise_k, c_in, c_skip, c_out, sigma, dt, half_dt, c_in_next, lo, hi;
};
__device__ __forceinline__ EdmConsts load_edm(const float* __restrict__ table, const int32_t* __restrict__ step) {
  const float* r = table + (int64_t)(step ? __ldcg(step) : 0) * kEdmRow;
  EdmConsts e;
  e.s_noise = r[0]; e.noise_k = r[1]; e.c_in = r[2]; e.c_skip = r[3]; e.c_out = r[4]; e.sigma = r[5]; e.dt = r[6];
  e.half_dt = r[7]; e.c_in_next = r[8]; e.lo = r[9]; e.hi = r[10];
  return e;
}
// model_output = clamp(c_skip * x + c_out * net)   (:352-358; products and sum rounded separately like the reference's tensor ops)
__device__ __forceinline__ float edm_denoised(const EdmConsts& e, float x, float net) {
  const float out = __fadd_rn(__fmul_rn(e.c_skip, x), __fmul_rn(e.c_out, net));
  return clamp_nan(out, e.lo, e.hi);
}
// Euler pass (:488-498): x_hat -> (d = (x_hat - D(x_hat)) / sigma_hat, x_euler = x_hat + (sigma_next - sigma_hat) d)
__device__ __forceinline__ void edm_euler_point(const EdmConsts& e, float net, float x_hat, float& d, float& x_euler, float& x0) {
  x0 = edm_denoised(e, x_hat, net);
  d = __fdiv_rn(__fsub_rn(x_hat, x0), e.sigma);
  x_euler = __fadd_rn(x_hat, __fmul_rn(e.dt, d));
}
// Heun pass (:502-516): x = x_hat + 0.5 (sigma_next - sigma_hat) (d + d'),  d' = (x_euler - D(x_euler)) / sigma_next
__device__ __forceinline__ void edm_heun_point(const EdmConsts& e, float net, float x_hat, float x_euler, float d, float& x_new, float& x0) {
  x0 = edm_denoised(e, x_euler, net);
  const float dp = __fdiv_rn(__fsub_rn(x_euler, x0), e.sigma);
  x_new = __fadd_rn(x_hat, __fmul_rn(e.half_dt, __fadd_rn(d, dp)));
}

// x_hat = x + sqrt(sigma_hat^2 - sigma^2) * (S_noise * eps);  x_in = c_in(sigma_hat) * x_hat     (:476-481, :349)
__global__ void edm_prepare_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ table,
                                   const int32_t* __restrict__ step, float* __restrict__ x_hat, float* __restrict__ x_in, int64_t count) {
  pdl_sync();
  const EdmConsts e = load_edm(table, step);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float xh = __fadd_rn(x[i], __fmul_rn(e.noise_k, __fmul_rn(e.s_noise, eps[i])));
    x_hat[i] = xh;
    x_in[i] = __fmul_rn(e.c_in, xh);
  }
}

template <typename T, int MAXCO>
__global__ void final_conv_kernel(const T* __restrict__ x, int ld, int64_t voxels, int c, int c_out, int nvec,
                                  const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ pred,
                                  int step_mode, const float* __restrict__ sched, const int32_t* __restrict__ step,
                                  const float* x_t /* may alias x_next (in-place update) */, const float* __restrict__ noise,
                                  float* x_next, float* __restrict__ x0, float* __restrict__ aux_out,
                                  int64_t total_rows, SubGeom sg) {
  constexpr int VEC = Vec<T>::N;
  // nvec (power of two <= 32) consecutive threads share one voxel row
  const int col = threadIdx.x % nvec;
  const int64_t rows_per_block = blockDim.x / nvec;
  float wv[MAXCO][VEC];
#pragma unroll
  for (int co = 0; co < MAXCO; ++co)
#pragma unroll
    for (int i = 0; i < VEC; ++i) wv[co][i] = co < c_out ? w[co * c + col * VEC + i] : 0.f;
  pdl_sync();
  // step_mode: 0 store the prediction, 1 DDPM update, 2 Elucidated Euler pass, 3 Elucidated Heun pass.  In the Elucidated modes
  // `noise` carries x_hat and `pred` the slope buffer d (written by 2, read by 3); aux_out receives the next network input.
  StepConsts sc;
  EdmConsts ec;
  if (step_mode == 1) sc = load_step(sched, step);
  if (step_mode >= 2) ec = load_edm(sched, step);
  // A group of nvec threads owns UNR CONSECUTIVE voxel rows: the row loads are coalesced 16-byte chunks, the dot products are
  // reduced with butterfly shuffles (every lane ends up with every sum), and lane u of the group finishes row u - so the fp32
  // sampler state (x_t, noise, x_next, x0) is read and written as contiguous runs instead of one float per 8 lanes.
  // Block-uniform trip count so the full-mask shuffles are always converged.
  constexpr int UNR = 8;
  using Raw = typename Vec<T>::Raw;
  const int R = min(UNR, nvec);  // rows per group: one finishing lane per row
  const int grp = threadIdx.x / nvec;
  for (int64_t base = (int64_t)blockIdx.x * rows_per_block * R; base < total_rows; base += (int64_t)gridDim.x * rows_per_block * R) {
    const int64_t row0 = base + (int64_t)grp * R;
    Raw raw[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      raw[u] = Raw();
      if (u < R && row0 + u < total_rows) raw[u] = Vec<T>::load_raw(x + (row0 + u) * ld + col * VEC);
    }
    float mine[MAXCO];
#pragma unroll
    for (int co = 0; co < MAXCO; ++co) mine[co] = 0.f;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (u >= R) break;  // block-uniform
      Vec<T> r;
      r.unpack(raw[u]);
#pragma unroll
      for (int co = 0; co < MAXCO; ++co) {
        if (co < c_out) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc = fmaf(r.v[i], wv[co][i], acc);
          for (int o = nvec >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          if (col == u) mine[co] = acc;
        }
      }
    }
    const int64_t row = row0 + col;
    if (col < R && row < total_rows) {
      int64_t b = row / voxels, v = row - b * voxels, ovox = voxels;
      if (sg.f > 1) {  // merged row -> (sub-volume, local voxel); the fp32 state tensors stay in sub-volume layout
        const int h = sg.h, fh = sg.f * sg.h;
        const int xx = (int)(row % fh), yy = (int)((row / fh) % fh), zz = (int)(row / ((int64_t)fh * fh));
        b = zz / h + sg.f * (yy / h) + sg.f * sg.f * (xx / h);
        v = ((int64_t)(zz % h) * h + yy % h) * h + xx % h;
        ovox = (int64_t)h * h * h;
      }
#pragma unroll
      for (int co = 0; co < MAXCO; ++co) {
        if (co >= c_out) break;
        const float p = mine[co] + bias[co];
        const int64_t o = (b * c_out + co) * ovox + v;  // NCDHW fp32
        if (!step_mode) {
          pred[o] = p;
        } else if (step_mode == 1) {
          float xn, xs;
          ddpm_point(sc, p, __ldcg(x_t + o), __ldcg(noise + o), xn, xs);
          x_next[o] = xn;
          x0[o] = xs;
        } else if (step_mode == 2) {
          float d, xe, xs;
          edm_euler_point(ec, p, __ldcg(noise + o), d, xe, xs);
          pred[o] = d;
          x_next[o] = xe;
          aux_out[o] = __fmul_rn(ec.c_in_next, xe);
          x0[o] = xs;
        } else {
          float xn, xs;
          edm_heun_point(ec, p, __ldcg(noise + o), __ldcg(x_t + o), __ldcg(pred + o), xn, xs);
          x_next[o] = xn;
          x0[o] = xs;
        }
      }
    }
  }
}

__global__ void ddpm_update_kernel(const float* pred /* may alias x0 */, const float* __restrict__ sched,
                                   const int32_t* __restrict__ step, const float* x_t /* may alias x_next */,
                                   const float* __restrict__ noise, float* x_next, float* x0, int64_t count) {
  pdl_sync();
  StepConsts sc = load_step(sched, step);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float xn, xs;
    ddpm_point(sc, __ldcg(pred + i), __ldcg(x_t + i), __ldcg(noise + i), xn, xs);
    x_next[i] = xn;
    if (x0) x0[i] = xs;
  }
}

// ---- time conditioning ------------------------------------------------------------------------------
__global__ void fourier_kernel(const float* __restrict__ t, int rows, const float* __restrict__ w, int half,
                               float* __restrict__ out) {
  const int r = blockIdx.x;
  const int width = 1 + 2 * half;
  const float tv = t[r];
  for (int j = threadIdx.x; j < width; j += blockDim.x) {
    float v;
    if (j == 0) v = tv;
    else {
      const int k = (j - 1) % half;
      const float f = tv * w[k] * 2.f * 3.14159265358979323846f;  // x * weights * 2 * pi  (:530)
      v = (j - 1) < half ? sinf(f) : cosf(f);
    }
    out[(int64_t)r * width + j] = v;
  }
}

__global__ void linear_kernel(const float* __restrict__ x, int ldx, int k, const float* __restrict__ w,
                              const float* __restrict__ b, int out_features, float* __restrict__ y, int ldy, int act_in,
                              int act_out) {
  extern __shared__ float xs[];
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    float v = x[(int64_t)r * ldx + i];
    xs[i] = act_in ? mish<false>(v) : v;
  }
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarps = blockDim.x / 32;
  for (int o = blockIdx.y * nwarps + warp; o < out_features; o += gridDim.y * nwarps) {
    float acc = 0.f;
    for (int i = lane; i < k; i += 32) acc = fmaf(xs[i], w[(int64_t)o * k + i], acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
      acc += b ? b[o] : 0.f;
      y[(int64_t)r * ldy + o] = act_out ? mish<false>(acc) : acc;
    }
  }
}

__global__ void advance_step_kernel(int32_t* step) {
  pdl_sync();
  *step += 1;
}

__global__ void clamp_kernel(float* __restrict__ x, int64_t count, float lo, float hi) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = clamp_nan(x[i], lo, hi);
}


// ---- init conv on the tensor cores: im2col of the fp32 planes into K = 64 bf16 columns --------------------------------------------
// init_conv (:1291) has K = 27 * c_in = 54 and is compute bound on CUDA cores (0.9 GMAC at 64^3: 108 us).  For 27 * c_in <= 64 the
// engine instead writes col[row][k = tap * c_in + ci] = plane_ci[voxel + tap] (zero outside the volume = the conv's zero padding,
// zero for k >= 27 * c_in) and runs a 1x1x1 tcgen05 convolution over it with the weights laid out to match.  One thread = one
// 16-byte chunk (8 columns) of one row; this is the first kernel of a step, launched without the PDL attribute.
// grid (ceil(d1 / 8), d0, n): a CTA stages the haloed input rows of an 8 (y) x d2 (x) tile of one z-plane in shared memory
// (c_in x 3 x 10 x (d2 + 2) floats, zero outside the volume = the conv's zero padding) and writes the 8 * d2 rows of 64 bf16 columns.
__global__ void __launch_bounds__(256) init_im2col_kernel(InitPlanes planes, int c_in, __nv_bfloat16* __restrict__ col, int n, int d0, int d1, int d2) {
  extern __shared__ float tile[];  // [c_in][3][10][d2 + 2], then int koff[64]
  const int W = d2 + 2, plane = 3 * 10 * W;
  int* koff = reinterpret_cast<int*>(tile + (size_t)c_in * plane);
  const int y0 = blockIdx.x * 8, z = blockIdx.y, b = blockIdx.z;
  if (threadIdx.x < 64) {
    const int k = threadIdx.x;
    int o = -1;
    if (k < 27 * c_in) {
      const int tap = k / c_in, ci = k - tap * c_in;
      o = ci * plane + ((tap / 9) * 10 + (tap / 3) % 3) * W + tap % 3;
    }
    koff[k] = o;
  }
  for (int idx = threadIdx.x; idx < c_in * plane; idx += blockDim.x) {
    const int ci = idx / plane;
    int r = idx - ci * plane;
    const int dz = r / (10 * W);
    r -= dz * 10 * W;
    const int yy = r / W, xx = r - yy * W;
    const int zz = z + dz - 1, y = y0 + yy - 1, x = xx - 1;
    float v = 0.f;
    if (zz >= 0 && zz < d0 && y >= 0 && y < d1 && x >= 0 && x < d2) {
      const float* pl = planes.p[0];
      long long ps = planes.stride[0];
#pragma unroll
      for (int q = 1; q < kInitMaxCin; ++q)  // select chain: no dynamic indexing of the parameter struct (would go through local memory)
        if (ci == q) { pl = planes.p[q]; ps = planes.stride[q]; }
      v = __ldg(pl + (int64_t)b * ps + ((int64_t)zz * d1 + y) * d2 + x);
    }
    tile[idx] = v;
  }
  __syncthreads();
  // a thread keeps the same 16-byte chunk (8 columns) for all its tasks (blockDim is a multiple of 8): its 8 tile offsets live in registers
  const int chunk = threadIdx.x & 7;
  int off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) off[j] = koff[chunk * 8 + j];
  const int nvox = 8 * d2;
  for (int vox = threadIdx.x >> 3; vox < nvox; vox += blockDim.x >> 3) {
    const int yl = vox / d2, x = vox - yl * d2;
    if (y0 + yl >= d1) break;  // later tasks of this thread have an even larger yl
    const int base = yl * W + x;
    uint32_t w[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float v0 = off[2 * jj] >= 0 ? tile[off[2 * jj] + base] : 0.f, v1 = off[2 * jj + 1] >= 0 ? tile[off[2 * jj + 1] + base] : 0.f;
      __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
      w[jj] = *reinterpret_cast<uint32_t*>(&h);
    }
    const int64_t row = (((int64_t)b * d0 + z) * d1 + (y0 + yl)) * d2 + x;
    *reinterpret_cast<uint4*>(col + row * 64 + chunk * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}


// ---- CrossEmbedLayer as init conv (imagen_pytorch3D.py:661-686, :1289-1291; the constructor default, init_cross_embed=True): several
// Conv3d(c_in, dim_scale, k, stride 1, padding (k-1)/2) with k in e.g. (3, 7, 15) over the same input, concatenated along channels.
// One launch per kernel size: one thread = one voxel x 16 output channels; the weights of one dz-slab ([k*k][c_in][16] floats) are
// staged in shared memory per step, every input value is read once per slab (L1-cached fp32 planes) and feeds 16 FMAs.
// 15^3 x 2 x 16 MACs per voxel is CUDA-core work by construction (K = 2 input channels); it is what the reference computes.
template <typename T>
__global__ void __launch_bounds__(128) init_conv_k_kernel(InitPlanes planes, int c_in, int k, const float* __restrict__ w /*[k^3][c_in][nco]*/,
                                                         const float* __restrict__ bias, T* __restrict__ out, int ld_out, int co_off, int nco,
                                                         int n, int d0, int d1, int d2) {
  extern __shared__ float sw[];  // [k*k][c_in][16]
  const int64_t vol = (int64_t)d0 * d1 * d2, total = (int64_t)n * vol;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = idx < total;
  const int pad = (k - 1) / 2, kk = k * k;
  int b = 0, z = 0, y = 0, x = 0;
  if (live) {
    b = (int)(idx / vol);
    int64_t r = idx - (int64_t)b * vol;
    z = (int)(r / ((int64_t)d1 * d2));
    r -= (int64_t)z * d1 * d2;
    y = (int)(r / d2);
    x = (int)(r - (int64_t)y * d2);
  }
  for (int co0 = 0; co0 < nco; co0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = (co0 + j < nco) ? bias[co0 + j] : 0.f;
    for (int dz = 0; dz < k; ++dz) {
      __syncthreads();
      for (int i = threadIdx.x; i < kk * c_in * 16; i += blockDim.x) {
        const int j = i & 15, tc = i >> 4;  // tc = tap_in_slab * c_in + ci
        sw[i] = (co0 + j < nco) ? w[((size_t)dz * kk * c_in + tc) * nco + co0 + j] : 0.f;
      }
      __syncthreads();
      const int zz = z + dz - pad;
      if (!live || zz < 0 || zz >= d0) continue;
      for (int dy = 0; dy < k; ++dy) {
        const int yy = y + dy - pad;
        if (yy < 0 || yy >= d1) continue;
        for (int dx = 0; dx < k; ++dx) {
          const int xx = x + dx - pad;
          if (xx < 0 || xx >= d2) continue;
          const int64_t off = ((int64_t)zz * d1 + yy) * d2 + xx;
          const float* wt = sw + (size_t)(dy * k + dx) * c_in * 16;
          for (int ci = 0; ci < c_in; ++ci) {
            const float* pl = planes.p[0];
            long long ps = planes.stride[0];
#pragma unroll
            for (int q = 1; q < kInitMaxCin; ++q)
              if (ci == q) { pl = planes.p[q]; ps = planes.stride[q]; }
            const float v = __ldg(pl + (int64_t)b * ps + off);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = fmaf(v, wt[ci * 16 + j], acc[j]);
          }
        }
      }
    }
    if (live) {
      T* orow = out + idx * ld_out + co_off + co0;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (co0 + j < nco) orow[j] = from_float<T>(acc[j]);
    }
  }
}
}  // namespace diqt

using namespace diqt;

extern "C" int diqt_abi_version(void) { return DIQT_ABI_VERSION; }
extern "C" const char* diqt_last_error(void) { return g_err; }
extern "C" uint64_t diqt_launch_count(void) { return g_launches.load(); }

extern "C" int diqt_init_conv_k(const float* const* planes, const int64_t* plane_stride, int c_in, int k, const float* w, const float* bias, void* out,
                                int ld_out, int co_off, int nco, int dtype, int n, int d0, int d1, int d2, void* stream) {
  DIQT_REQUIRE(planes && plane_stride && w && bias && out && n > 0 && d0 > 0 && d1 > 0 && d2 > 0, "init_conv_k: bad arguments");
  DIQT_REQUIRE(c_in > 0 && c_in <= kInitMaxCin, "init_conv_k: c_in=%d (max %d)", c_in, kInitMaxCin);
  DIQT_REQUIRE(k >= 1 && k % 2 == 1 && k <= 31, "init_conv_k: kernel size %d (odd, <= 31)", k);
  DIQT_REQUIRE(nco > 0 && co_off >= 0 && co_off + nco <= ld_out, "init_conv_k: channel slice [%d, %d) does not fit pitch %d", co_off, co_off + nco, ld_out);
  DIQT_REQUIRE(dtype == DIQT_F32 || dtype == DIQT_BF16, "init_conv_k: bad dtype %d", dtype);
  InitPlanes ip;
  for (int i = 0; i < kInitMaxCin; ++i) {
    ip.p[i] = i < c_in ? planes[i] : nullptr;
    ip.stride[i] = i < c_in ? plane_stride[i] : 0;
  }
  const size_t sh = (size_t)k * k * c_in * 16 * sizeof(float);
  DIQT_REQUIRE(sh <= 160 * 1024, "init_conv_k: weight slab of %zu bytes does not fit shared memory", sh);
  const int64_t total = (int64_t)n * d0 * d1 * d2;
  const unsigned blocks = (unsigned)((total + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DIQT_BF16) {
    auto kern = init_conv_k_kernel<__nv_bfloat16>;
    if (sh > 48 * 1024) DIQT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    kern<<<blocks, 128, sh, st>>>(ip, c_in, k, w, bias, (__nv_bfloat16*)out, ld_out, co_off, nco, n, d0, d1, d2);
  } else {
    auto kern = init_conv_k_kernel<float>;
    if (sh > 48 * 1024) DIQT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    kern<<<blocks, 128, sh, st>>>(ip, c_in, k, w, bias, (float*)out, ld_out, co_off, nco, n, d0, d1, d2);
  }
  return check_launch("init_conv_k");
}

extern "C" int diqt_init_im2col(const float* const* planes, const int64_t* plane_stride, int c_in, void* col, int n, int d0, int d1, int d2,
                                void* stream) {
  DIQT_REQUIRE(planes && plane_stride && col && n > 0 && d0 > 0 && d1 > 0 && d2 > 0, "init_im2col: bad arguments");
  DIQT_REQUIRE(c_in > 0 && c_in <= kInitMaxCin && 27 * c_in <= 64, "init_im2col: 27 * c_in = %d columns do not fit K = 64", 27 * c_in);
  InitPlanes ip;
  for (int i = 0; i < kInitMaxCin; ++i) {
    ip.p[i] = i < c_in ? planes[i] : nullptr;
    ip.stride[i] = i < c_in ? plane_stride[i] : 0;
  }
  const size_t sh = (size_t)c_in * 3 * 10 * (d2 + 2) * sizeof(float) + 64 * sizeof(int);
  DIQT_REQUIRE(sh <= 48 * 1024 && d0 <= 65535 && n <= 65535, "init_im2col: d2=%d too wide for the shared-memory tile", d2);
  const dim3 grid((d1 + 7) / 8, d0, n);
  init_im2col_kernel<<<grid, 256, sh, (cudaStream_t)stream>>>(ip, c_in, (__nv_bfloat16*)col, n, d0, d1, d2);
  return check_launch("init_im2col");
}

extern "C" int diqt_init_conv_pack(const float* w, int c_out, int c_in, float* packed, void* stream) {
  DIQT_REQUIRE(w && packed && c_in > 0 && c_in <= kInitMaxCin, "init_conv_pack: c_in=%d (max %d)", c_in, kInitMaxCin);
  init_conv_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(w, c_out, c_in, packed);
  return check_launch("init_conv_pack");
}

extern "C" int diqt_init_conv(const float* const* planes, const int64_t* plane_stride, int c_in, const float* w_packed,
                              const float* bias, void* out, int ld_out, int dtype, int n, int d0, int d1, int d2, int c_out,
                              int sub_f, int sub_h, void* stream) {
  const SubGeom sg{sub_f, sub_h};
  if (sub_f > 1) DIQT_REQUIRE(n == 1 && d0 == sub_f * sub_h && d1 == d0 && d2 == d0, "init_conv: boundary mode runs over ONE merged volume of side f*h");
  DIQT_REQUIRE(planes && plane_stride && w_packed && bias && out, "init_conv: null pointer");
  DIQT_REQUIRE(c_in > 0 && c_in <= kInitMaxCin, "init_conv: c_in=%d (max %d)", c_in, kInitMaxCin);
  DIQT_REQUIRE(c_out % 16 == 0, "init_conv: c_out=%d must be a multiple of 16", c_out);
  InitPlanes ip;
  for (int i = 0; i < kInitMaxCin; ++i) {
    ip.p[i] = i < c_in ? planes[i] : nullptr;
    ip.stride[i] = i < c_in ? plane_stride[i] : 0;
  }
  const int64_t total = (int64_t)n * d0 * d1 * d2;
  const int threads = 128;
  const size_t sh = ((size_t)27 * c_in * c_out + c_out) * sizeof(float);
  DIQT_REQUIRE(sh <= 200 * 1024, "init_conv: weights do not fit shared memory");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (d2 % 4 == 0 && c_out % 32 == 0) {
    const int64_t total4 = total / 4;
    const unsigned blocks4 = (unsigned)((total4 + threads - 1) / threads);
    if (dtype == DIQT_BF16) {
      auto k = init_conv_x4_kernel<__nv_bfloat16, 32>;
      if (sh > 48 * 1024) DIQT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
      launch_pdl<false>(k, blocks4, threads, sh, st, ip, c_in, w_packed, bias, (__nv_bfloat16*)out, ld_out, n, d0, d1, d2, c_out, sg);
    } else {
      auto k = init_conv_x4_kernel<float, 32>;
      if (sh > 48 * 1024) DIQT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
      launch_pdl<false>(k, blocks4, threads, sh, st, ip, c_in, w_packed, bias, (float*)out, ld_out, n, d0, d1, d2, c_out, sg);
    }
    return check_launch("init_conv_x4");
  }
#define DIQT_INIT_LAUNCH(T, COT)                                                                                  \
  do {                                                                                                            \
    auto k = init_conv_kernel<T, COT>;                                                                            \
    if (sh > 48 * 1024) DIQT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh)); \
    k<<<blocks, threads, sh, st>>>(ip, c_in, w_packed, bias, (T*)out, ld_out, n, d0, d1, d2, c_out, sg);          \
  } while (0)
  if (dtype == DIQT_BF16) {
    if (c_out % 64 == 0) DIQT_INIT_LAUNCH(__nv_bfloat16, 64);
    else if (c_out % 32 == 0) DIQT_INIT_LAUNCH(__nv_bfloat16, 32);
    else DIQT_INIT_LAUNCH(__nv_bfloat16, 16);
  } else {
    if (c_out % 64 == 0) DIQT_INIT_LAUNCH(float, 64);
    else if (c_out % 32 == 0) DIQT_INIT_LAUNCH(float, 32);
    else DIQT_INIT_LAUNCH(float, 16);
  }
#undef DIQT_INIT_LAUNCH
  return check_launch("init_conv");
}

static int final_conv_impl(const void* x, int ld, int dtype, int n, int64_t voxels, int c, int c_out, const float* w,
                           const float* bias, float* pred, int step_mode, const float* sched, const int32_t* step,
                           const float* x_t, const float* noise, float* x_next, float* x0, float* aux_out, int sub_f, int sub_h, void* stream) {
  const SubGeom sg{sub_f, sub_h};
  if (sub_f > 1) DIQT_REQUIRE(n == 1 && voxels == (int64_t)sub_f * sub_f * sub_f * sub_h * sub_h * sub_h, "final_conv: boundary mode runs over ONE merged volume");
  const int vec = dtype == DIQT_BF16 ? 8 : 4;
  DIQT_REQUIRE(x && w && bias, "final_conv: null pointer");
  DIQT_REQUIRE(c_out >= 1 && c_out <= 4, "final_conv: c_out=%d (1..4 supported)", c_out);
  DIQT_REQUIRE(c % vec == 0 && ld % vec == 0, "final_conv: c=%d not a multiple of %d", c, vec);
  const int nvec = c / vec;
  DIQT_REQUIRE(nvec <= 32 && (nvec & (nvec - 1)) == 0, "final_conv: c/%d=%d must be a power of two <= 32", vec, nvec);
  DIQT_REQUIRE(step_mode >= 0 && step_mode <= 3, "final_conv: bad step_mode %d", step_mode);
  if (step_mode) DIQT_REQUIRE(sched && x_t && noise && x_next && x0, "final_conv: fused step needs sched/x_t/noise/x_next/x0");
  if (step_mode != 1) DIQT_REQUIRE(pred, "final_conv: pred is null");
  if (step_mode == 2) DIQT_REQUIRE(aux_out, "final_conv: the Euler pass needs the next-input buffer");
  const int64_t rows = (int64_t)n * voxels;
  const int threads = 256;
  const int64_t rpb = threads / nvec;
  const int64_t rgrp = nvec < 8 ? nvec : 8;  // rows per thread group (final_conv_kernel: R)
  int64_t blocks = (rows + rpb * rgrp - 1) / (rpb * rgrp);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DIQT_BF16)
    launch_pdl<false>(final_conv_kernel<__nv_bfloat16, 4>, (unsigned)blocks, threads, 0, st, (const __nv_bfloat16*)x, ld, voxels, c, c_out, nvec, w,
               bias, pred, step_mode, sched, step, x_t, noise, x_next, x0, aux_out, rows, sg);
  else
    launch_pdl<false>(final_conv_kernel<float, 4>, (unsigned)blocks, threads, 0, st, (const float*)x, ld, voxels, c, c_out, nvec, w, bias, pred,
               step_mode, sched, step, x_t, noise, x_next, x0, aux_out, rows, sg);
  return check_launch("final_conv");
}

extern "C" int diqt_final_conv(const void* x, int ld, int dtype, int n, int64_t voxels, int c, int c_out, const float* w,
                               const float* bias, float* pred, int step_mode, const float* sched, const int32_t* step,
                               const float* x_t, const float* noise, float* x_next, float* x0, int sub_f, int sub_h, void* stream) {
  DIQT_REQUIRE(step_mode == 0 || step_mode == 1, "final_conv: step_mode %d (0 or 1; the Elucidated passes go through diqt_final_conv_edm)", step_mode);
  return final_conv_impl(x, ld, dtype, n, voxels, c, c_out, w, bias, pred, step_mode, sched, step, x_t, noise, x_next, x0, nullptr, sub_f, sub_h, stream);
}

extern "C" int diqt_final_conv_edm(const void* x, int ld, int dtype, int n, int64_t voxels, int c, int c_out, const float* w,
                                   const float* bias, int pass, const float* table, const int32_t* step, const float* x_hat,
                                   float* slope, float* state, float* x0, float* next_input, int sub_f, int sub_h, void* stream) {
  DIQT_REQUIRE(pass == 0 || pass == 1, "final_conv_edm: pass %d (0 Euler, 1 Heun)", pass);
  return final_conv_impl(x, ld, dtype, n, voxels, c, c_out, w, bias, slope, 2 + pass, table, step, state, x_hat, state, x0, next_input, sub_f,
                         sub_h, stream);
}

extern "C" int diqt_edm_prepare(const float* x, const float* eps, const float* table, const int32_t* step, float* x_hat, float* x_in,
                                int64_t count, void* stream) {
  DIQT_REQUIRE(x && eps && table && x_hat && x_in && count > 0, "edm_prepare: bad arguments");
  int64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl<false>(edm_prepare_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, x, eps, table, step, x_hat, x_in, count);
  return check_launch("edm_prepare");
}

// the same two updates on an existing network output (dynamic thresholding runs its quantile in torch between them)
__global__ void edm_update_kernel(const float* __restrict__ denoised, int pass, const float* __restrict__ table, const int32_t* __restrict__ step,
                                  const float* __restrict__ x_hat, float* __restrict__ slope, float* __restrict__ state,
                                  float* __restrict__ next_input, int64_t count) {
  pdl_sync();
  const EdmConsts e = load_edm(table, step);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float x0 = denoised[i];
    if (pass == 0) {
      const float d = __fdiv_rn(__fsub_rn(x_hat[i], x0), e.sigma);
      const float xe = __fadd_rn(x_hat[i], __fmul_rn(e.dt, d));
      slope[i] = d;
      state[i] = xe;
      next_input[i] = __fmul_rn(e.c_in_next, xe);
    } else {
      const float xe = state[i];
      const float dp = __fdiv_rn(__fsub_rn(xe, x0), e.sigma);
      state[i] = __fadd_rn(x_hat[i], __fmul_rn(e.half_dt, __fadd_rn(slope[i], dp)));
    }
  }
}

extern "C" int diqt_edm_update(const float* denoised, int pass, const float* table, const int32_t* step, const float* x_hat, float* slope,
                               float* state, float* next_input, int64_t count, void* stream) {
  DIQT_REQUIRE(denoised && table && x_hat && slope && state && count > 0 && (pass == 0 || pass == 1), "edm_update: bad arguments");
  if (pass == 0) DIQT_REQUIRE(next_input, "edm_update: the Euler pass needs the next-input buffer");
  int64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl<false>(edm_update_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, denoised, pass, table, step, x_hat, slope, state, next_input, count);
  return check_launch("edm_update");
}

extern "C" int diqt_ddpm_update(const float* pred, const float* sched, const int32_t* step, const float* x_t,
                                const float* noise, float* x_next, float* x0, int64_t count, void* stream) {
  DIQT_REQUIRE(pred && sched && x_t && noise && x_next, "ddpm_update: null pointer");
  int64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl<false>(ddpm_update_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, pred, sched, step, x_t, noise, x_next, x0, count);
  return check_launch("ddpm_update");
}

extern "C" int diqt_fourier_features(const float* t, int rows, const float* w, int half, float* out, void* stream) {
  DIQT_REQUIRE(t && w && out && rows > 0 && half > 0, "fourier_features: bad arguments");
  fourier_kernel<<<rows, 64, 0, (cudaStream_t)stream>>>(t, rows, w, half, out);
  return check_launch("fourier_features");
}

extern "C" int diqt_linear(const float* x, int ldx, int rows, int k, const float* w, const float* b, int out_features,
                           float* y, int ldy, int act_in, int act_out, void* stream) {
  DIQT_REQUIRE(x && w && y && rows > 0 && k > 0 && out_features > 0, "linear: bad arguments");
  DIQT_REQUIRE((size_t)k * sizeof(float) <= 48 * 1024, "linear: k=%d too large", k);
  dim3 grid(rows, (out_features + 63) / 64);
  linear_kernel<<<grid, 256, (size_t)k * sizeof(float), (cudaStream_t)stream>>>(x, ldx, k, w, b, out_features, y, ldy, act_in,
                                                                               act_out);
  return check_launch("linear");
}

extern "C" int diqt_advance_step(int32_t* step, void* stream) {
  DIQT_REQUIRE(step, "advance_step: null pointer");
  launch_pdl<false>(advance_step_kernel, 1, 1, 0, (cudaStream_t)stream, step);
  return check_launch("advance_step");
}

extern "C" int diqt_clamp(float* x, int64_t count, float lo, float hi, void* stream) {
  DIQT_REQUIRE(x && count > 0, "clamp: bad arguments");
  int64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  clamp_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, count, lo, hi);
  return check_launch("clamp");
}
