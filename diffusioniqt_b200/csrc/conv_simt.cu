// CUDA-core implicit-GEMM convolution (fp32 accumulate) for the exact fp32 mode and for channel
// counts the tcgen05 kernel does not take.  One kernel covers the four conv shapes on the path
// through a small tap table:
//   K3   : 27 taps, offsets -1..1, stride 1      Block.project           imagen_pytorch3D.py:550-553
//   K1   : 1 tap                                 res_conv :597, :1388
//   DOWN : 8 taps (s1,s2,s3) in {0,1}^3, stride 2  pixel-unshuffle + 1x1   :489-496
//   UP   : 1 tap, N = 8*C, Mish + pixel-shuffle in the epilogue            :459-487, :416-439
//
// GEMM view: M = n*od0*od1*od2 output voxels (linear order), N = c_out, K = taps * c_in.
// Tile 64 x 64 x 16, 256 threads, 4x4 register micro-tile, register-prefetch double buffering.
// Weights are packed [tap][c_in][c_out] fp32 (UP: output channel order re-arranged to
// [sub-position][c] so that shuffled stores are contiguous).
#include "common.cuh"

namespace diqt {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int AS_LD = BM + 4;

struct SimtConvParams {
  const void* in;
  void* out;
  const float* w;     // packed
  const float* bias;  // [c_out] in GEMM-N order
  int n, id0, id1, id2;  // input dims
  int od0, od1, od2;     // GEMM output voxel grid
  int c_in, ld_in, c_out, ld_out;
  int taps, stride, mode;
  int64_t m_total;
};

__device__ __forceinline__ void tap_offset(int mode, int t, int& dz, int& dy, int& dx) {
  if (mode == DIQT_CONV_K3) { dz = t / 9 - 1; dy = (t / 3) % 3 - 1; dx = t % 3 - 1; }
  else if (mode == DIQT_CONV_DOWN) { dz = (t >> 2) & 1; dy = (t >> 1) & 1; dx = t & 1; }
  else { dz = dy = dx = 0; }
}

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(SimtConvParams p) {
  pdl_sync();
  __shared__ float As[2][BK][AS_LD];
  __shared__ float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- loader roles: A: voxel lm = tid/4, channel quad lq = tid%4 ; B: row bk = tid/16, col quad bq = tid%16
  const int lm = tid >> 2, lq = tid & 3;
  const int64_t gm = m0 + lm;
  const bool m_ok = gm < p.m_total;
  int vb = 0, vz = 0, vy = 0, vx = 0;
  if (m_ok) {
    int64_t r = gm;
    const int64_t ovol = (int64_t)p.od0 * p.od1 * p.od2;
    vb = (int)(r / ovol); r -= (int64_t)vb * ovol;
    vz = (int)(r / ((int64_t)p.od1 * p.od2)); r -= (int64_t)vz * p.od1 * p.od2;
    vy = (int)(r / p.od2); vx = (int)(r - (int64_t)vy * p.od2);
  }
  const int bk = tid >> 4, bq = tid & 15;
  const T* in = reinterpret_cast<const T*>(p.in);

  const int kchunks = p.c_in / BK;
  const int ksteps = p.taps * kchunks;

  float a_reg[4], b_reg[4];
  auto fetch = [&](int ks) {
    const int t = ks / kchunks, c0 = (ks - t * kchunks) * BK;
    int dz, dy, dx;
    tap_offset(p.mode, t, dz, dy, dx);
    const int iz = vz * p.stride + dz, iy = vy * p.stride + dy, ix = vx * p.stride + dx;
    const bool ok = m_ok && iz >= 0 && iz < p.id0 && iy >= 0 && iy < p.id1 && ix >= 0 && ix < p.id2;
    if (ok) {
      const T* src = in + ((((int64_t)vb * p.id0 + iz) * p.id1 + iy) * p.id2 + ix) * p.ld_in + c0 + lq * 4;
      if constexpr (sizeof(T) == 4) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src));  // L2 only: see Vec<T>::load
        a_reg[0] = v.x; a_reg[1] = v.y; a_reg[2] = v.z; a_reg[3] = v.w;
      } else {
        const uint2 v = __ldcg(reinterpret_cast<const uint2*>(src));
        a_reg[0] = __uint_as_float(v.x << 16); a_reg[1] = __uint_as_float(v.x & 0xffff0000u);
        a_reg[2] = __uint_as_float(v.y << 16); a_reg[3] = __uint_as_float(v.y & 0xffff0000u);
      }
    } else {
      a_reg[0] = a_reg[1] = a_reg[2] = a_reg[3] = 0.f;
    }
    const int nn = n0 + bq * 4;
    if (nn < p.c_out) {
      const float4 v = *reinterpret_cast<const float4*>(p.w + ((int64_t)t * p.c_in + c0 + bk) * p.c_out + nn);
      b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
    } else {
      b_reg[0] = b_reg[1] = b_reg[2] = b_reg[3] = 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) As[buf][lq * 4 + i][lm] = a_reg[i];
    *reinterpret_cast<float4*>(&Bs[buf][bk][bq * 4]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
  };

  // ---- compute roles: 16 x 16 threads, each a 4 (m) x 4 (n) micro-tile
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  fetch(0);
  stash(0);
  __syncthreads();
  for (int ks = 0; ks < ksteps; ++ks) {
    const int buf = ks & 1;
    if (ks + 1 < ksteps) fetch(ks + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][tm]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tn]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (ks + 1 < ksteps) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue
  T* out = reinterpret_cast<T*>(p.out);
  const int nn = n0 + tn;
  if (nn >= p.c_out) return;
  float bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bias[j] = p.bias ? p.bias[nn + j] : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + tm + i;
    if (m >= p.m_total) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    T* dst;
    if (p.mode == DIQT_CONV_UP) {
      // GEMM column nn = sub * C + c (packed order); scatter to voxel (2z+i, 2y+j, 2x+k), channel c
      const int C = p.c_out / 8;
      const int sub = nn / C, c = nn - sub * C;
      int64_t r = m;
      const int64_t ovol = (int64_t)p.od0 * p.od1 * p.od2;
      const int b = (int)(r / ovol); r -= (int64_t)b * ovol;
      const int z = (int)(r / ((int64_t)p.od1 * p.od2)); r -= (int64_t)z * p.od1 * p.od2;
      const int y = (int)(r / p.od2), x = (int)(r - (int64_t)y * p.od2);
      const int Z = 2 * z + ((sub >> 2) & 1), Y = 2 * y + ((sub >> 1) & 1), X = 2 * x + (sub & 1);
      dst = out + ((((int64_t)b * (2 * p.od0) + Z) * (2 * p.od1) + Y) * (2 * p.od2) + X) * p.ld_out + c;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = sizeof(T) == 4 ? mish<false>(v[j]) : mish<true>(v[j]);
    } else {
      dst = out + m * p.ld_out + nn;
    }
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(dst) = u;
    }
  }
}

// (c_out, c_in_total, k,k,k) fp32 -> [tap][c_in][c_out'] fp32 ; also permutes the bias for UP
__global__ void conv_pack_simt_kernel(const float* __restrict__ w, int mode, int c_in, int c_out, int taps,
                                      float* __restrict__ packed) {
  const int64_t total = (int64_t)taps * c_in * c_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int nn = (int)(i % c_out), ci = (int)((i / c_out) % c_in), t = (int)(i / ((int64_t)c_out * c_in));
    float v;
    if (mode == DIQT_CONV_K3) v = w[((int64_t)nn * c_in + ci) * 27 + t];
    else if (mode == DIQT_CONV_K1) v = w[(int64_t)nn * c_in + ci];
    else if (mode == DIQT_CONV_DOWN) v = w[(int64_t)nn * (c_in * 8) + ci * 8 + t];  // (c s1 s2 s3) channel order
    else {  // UP: packed column nn = sub*C + c  <-  reference output channel c*8 + sub
      const int C = c_out / 8, sub = nn / C, c = nn - sub * C;
      v = w[(int64_t)(c * 8 + sub) * c_in + ci];
    }
    packed[i] = v;
  }
}

__global__ void bias_permute_up_kernel(const float* __restrict__ b, int c_out, float* __restrict__ out) {
  const int C = c_out / 8;
  for (int nn = blockIdx.x * blockDim.x + threadIdx.x; nn < c_out; nn += gridDim.x * blockDim.x) {
    const int sub = nn / C, c = nn - sub * C;
    out[nn] = b[c * 8 + sub];
  }
}

int simt_taps(int mode) { return mode == DIQT_CONV_K3 ? 27 : mode == DIQT_CONV_DOWN ? 8 : 1; }

int conv_simt_pack(const diqt_conv_desc* d, const float* w, void* packed, cudaStream_t st) {
  conv_pack_simt_kernel<<<256, 256, 0, st>>>(w, d->mode, d->c_in, d->c_out, simt_taps(d->mode), (float*)packed);
  return check_launch("conv_pack_simt");
}

int conv_simt_run(const diqt_conv_desc* d, const void* in, void* out, const void* packed, const float* bias,
                  cudaStream_t st) {
  DIQT_REQUIRE(d->c_in % BK == 0, "conv(simt): c_in=%d must be a multiple of %d", d->c_in, BK);
  DIQT_REQUIRE(d->c_out % 4 == 0 && d->ld_in % 4 == 0 && d->ld_out % 4 == 0, "conv(simt): channel counts must be multiples of 4");
  SimtConvParams p;
  p.in = in; p.out = out; p.w = (const float*)packed; p.bias = bias;
  p.n = d->n; p.id0 = d->d0; p.id1 = d->d1; p.id2 = d->d2;
  p.stride = d->mode == DIQT_CONV_DOWN ? 2 : 1;
  if (d->mode == DIQT_CONV_DOWN) {
    DIQT_REQUIRE(d->d0 % 2 == 0 && d->d1 % 2 == 0 && d->d2 % 2 == 0, "conv(down): odd input dims");
    p.od0 = d->d0 / 2; p.od1 = d->d1 / 2; p.od2 = d->d2 / 2;
  } else {
    p.od0 = d->d0; p.od1 = d->d1; p.od2 = d->d2;
  }
  if (d->mode == DIQT_CONV_UP) DIQT_REQUIRE(d->c_out % 32 == 0, "conv(up): c_out=%d must be a multiple of 32", d->c_out);
  p.c_in = d->c_in; p.ld_in = d->ld_in; p.c_out = d->c_out; p.ld_out = d->ld_out;
  p.taps = simt_taps(d->mode); p.mode = d->mode;
  p.m_total = (int64_t)d->n * p.od0 * p.od1 * p.od2;
  dim3 grid((unsigned)((p.m_total + BM - 1) / BM), (unsigned)((d->c_out + BN - 1) / BN));
  if (d->dtype == DIQT_BF16) launch_pdl(conv_simt_kernel<__nv_bfloat16>, grid, 256, 0, st, p);
  else launch_pdl(conv_simt_kernel<float>, grid, 256, 0, st, p);
  return check_launch("conv_simt");
}

int conv_bias_permute_up(const float* b, int c_out, float* out, cudaStream_t st) {
  bias_permute_up_kernel<<<4, 256, 0, st>>>(b, c_out, out);
  return check_launch("bias_permute_up");
}

}  // namespace diqt
