// init_conv on the tensor cores in ONE kernel (Unet.init_conv, imagen_pytorch3D.py:1291: Conv3d(c_in, init_dim, 3, padding 1) over the
// concatenation [cond_images, x, lowres_cond_img] of :1569-1584).
//
// K = 27 * c_in (54 for the driver's two input channels) is not a tensor-core shape as a 3x3x3 convolution, but it is as a GEMM over
// im2col rows of K = 64 bf16 columns.  Round 1 materialised those rows in global memory (128 bytes per voxel written and read back:
// init_im2col_kernel 29 us + the 1x1x1 tcgen05 conv 26 us at 64^3).  Here the rows never leave the SM: a persistent CTA
//   1. stages the haloed fp32 input of an 8 (y) x d2 (x) tile of one z-plane in shared memory (zero outside the volume = the padding),
//   2. writes the 8 * d2 im2col rows as bf16 straight into 128B-swizzled A tiles (128 rows x 64 columns each),
//   3. issues tcgen05.mma (M128, N = c_out, K64) per A tile against the packed weights into tensor memory,
//   4. drains TMEM: bias, bf16, 128-byte row stores (consecutive rows are consecutive in memory), and the channel sums / sums of squares
//      the first GroupNorm needs, accumulated over the CTA's whole life and written as ONE partial row per CTA (+ grouped sink).
// HBM traffic: 8 bytes in + 128 bytes out per voxel instead of 8 + 128 + 128 + 128.
#include "tc_common.cuh"

namespace diqt {

constexpr int IT_MAX_CIN = 8;
constexpr int IT_THREADS = 256;
constexpr int IT_TY = 8;  // y rows per work item

struct InitTcParams {
  const float* plane[IT_MAX_CIN];
  long long stride[IT_MAX_CIN];
  const uint8_t* w;    // [c_out][64] bf16, rows pre-swizzled (16-byte chunk ^ (row & 7)); column k = tap * c_in + ci, zero padded to 64
  const float* bias;   // [c_out]
  __nv_bfloat16* out;  // channels-last rows, pitch ld_out
  float* stats;        // NULL or partial[n][gridDim.x][c_out][2]
  StatsGroups sink;
  int c_in, c_out, ld_out, n, d0, d1, d2, ytiles, items;
  uint32_t idesc;
};

__global__ void __launch_bounds__(IT_THREADS, 2) init_conv_tc_kernel(const __grid_constant__ InitTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int W = p.d2 + 2, plane = 3 * (IT_TY + 2) * W;
  const int rows = IT_TY * p.d2, mtiles = rows / 128;
  uint8_t* a_s = smem;                                              // [mtiles][128 rows][128 B]; reused as the bf16 staging tile of the epilogue
  uint8_t* w_s = a_s + (size_t)mtiles * 16384;                      // [c_out][128 B]
  float* tile = reinterpret_cast<float*>(w_s + (size_t)p.c_out * 128);  // [c_in][3][10][W]
  int* koff = reinterpret_cast<int*>(tile + (size_t)p.c_in * plane);    // [64]
  float* s_bias = reinterpret_cast<float*>(koff + 64);              // [c_out]
  float* s_red = s_bias + p.c_out;                                  // [8 warps][c_out][2]
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(s_red + 8 * p.c_out * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) {
    const int k = threadIdx.x;
    int o = -1;
    if (k < 27 * p.c_in) {
      const int tap = k / p.c_in, ci = k - tap * p.c_in;
      o = ci * plane + ((tap / 9) * (IT_TY + 2) + (tap / 3) % 3) * W + tap % 3;
    }
    koff[k] = o;
  }
  for (int i = threadIdx.x; i < p.c_out; i += IT_THREADS) s_bias[i] = p.bias[i];
  for (int i = threadIdx.x; i < p.c_out * 8; i += IT_THREADS) reinterpret_cast<uint4*>(w_s)[i] = reinterpret_cast<const uint4*>(p.w)[i];
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(mma_done), 1);
    fence_barrier_init();
  }
  const uint32_t ncols = (uint32_t)(mtiles * p.c_out <= 128 ? 128 : mtiles * p.c_out <= 256 ? 256 : 512);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();  // the weights were written with generic stores and are read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int chunk = threadIdx.x & 7;  // a thread always builds the same 16-byte chunk (8 im2col columns)
  int off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) off[j] = koff[chunk * 8 + j];
  // epilogue mapping: warp w drains TMEM lane quarter (w & 3) of the A tiles (w >> 2), (w >> 2) + 2, ...
  const int quarter = warp & 3;
  float st_s[4], st_q[4];  // this thread's channels: 2 * lane, 2 * lane + 1 (+ 64 for c_out = 128)
#pragma unroll
  for (int i = 0; i < 4; ++i) st_s[i] = st_q[i] = 0.f;
  int cur_b = -1;
  const int nblk = (int)gridDim.x;
  uint32_t phase = 0;

  auto flush = [&](int b) {
    // per-warp sums -> shared -> one partial row of this CTA for volume b
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int ch = (h >> 1) * 64 + 2 * lane + (h & 1);
      if (ch < p.c_out) {
        s_red[(warp * p.c_out + ch) * 2] = st_s[h];
        s_red[(warp * p.c_out + ch) * 2 + 1] = st_q[h];
      }
      st_s[h] = st_q[h] = 0.f;
    }
    __syncthreads();
    float* dst = p.stats + ((size_t)b * nblk + blockIdx.x) * p.c_out * 2;
    for (int ch = threadIdx.x; ch < p.c_out; ch += IT_THREADS) {
      float a = 0.f, q = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) { a += s_red[(w8 * p.c_out + ch) * 2]; q += s_red[(w8 * p.c_out + ch) * 2 + 1]; }
      dst[ch * 2] = a;
      dst[ch * 2 + 1] = q;
    }
    __syncthreads();
  };

  if (p.stats) {  // every (volume, channel) entry of this CTA's row must be defined even if the CTA never sees that volume
    for (int b = 0; b < p.n; ++b)
      for (int idx = threadIdx.x; idx < p.c_out * 2; idx += IT_THREADS) p.stats[((size_t)b * nblk + blockIdx.x) * p.c_out * 2 + idx] = 0.f;
    __syncthreads();
  }

  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int yt = item % p.ytiles;
    int t = item / p.ytiles;
    const int z = t % p.d0, b = t / p.d0;
    const int y0 = yt * IT_TY;
    if (p.stats && b != cur_b) {
      if (cur_b >= 0) flush(cur_b);
      cur_b = b;
    }
    // ---- 1. haloed fp32 input tile
    for (int idx = threadIdx.x; idx < p.c_in * plane; idx += IT_THREADS) {
      const int ci = idx / plane;
      int r = idx - ci * plane;
      const int dz = r / ((IT_TY + 2) * W);
      r -= dz * (IT_TY + 2) * W;
      const int yy = r / W, xx = r - yy * W;
      const int zz = z + dz - 1, y = y0 + yy - 1, x = xx - 1;
      float v = 0.f;
      if (zz >= 0 && zz < p.d0 && y >= 0 && y < p.d1 && x >= 0 && x < p.d2) {
        const float* pl = p.plane[0];
        long long ps = p.stride[0];
#pragma unroll
        for (int q = 1; q < IT_MAX_CIN; ++q)  // select chain: no dynamic indexing of the parameter struct
          if (ci == q) { pl = p.plane[q]; ps = p.stride[q]; }
        v = __ldg(pl + (int64_t)b * ps + ((int64_t)zz * p.d1 + y) * p.d2 + x);
      }
      tile[idx] = v;
    }
    __syncthreads();
    // ---- 2. im2col rows -> swizzled A tiles
    for (int vox = threadIdx.x >> 3; vox < rows; vox += IT_THREADS >> 3) {
      const int yl = vox / p.d2, x = vox - yl * p.d2;
      const int base = yl * W + x;
      uint32_t wv[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float v0 = off[2 * jj] >= 0 ? tile[off[2 * jj] + base] : 0.f, v1 = off[2 * jj + 1] >= 0 ? tile[off[2 * jj + 1] + base] : 0.f;
        __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
        wv[jj] = *reinterpret_cast<uint32_t*>(&h);
      }
      const int mt = vox >> 7, r = vox & 127;
      *reinterpret_cast<uint4*>(a_s + (size_t)mt * 16384 + (size_t)r * 128 + ((chunk ^ (r & 7)) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    fence_proxy_async();
    __syncthreads();
    // ---- 3. MMAs (warp-uniform code in warp 0; the issuing lane is elected inside the wrappers)
    if (warp == 0) {
      tc_fence_after();
      const uint64_t bdesc = make_sw128_desc(smem_u32(w_s));
      for (int mt = 0; mt < mtiles; ++mt) {
        const uint64_t adesc = make_sw128_desc(smem_u32(a_s + (size_t)mt * 16384));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + (uint32_t)(mt * p.c_out), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc, k != 0);
      }
      umma_commit(smem_u32(mma_done));
    }
    mbar_wait(smem_u32(mma_done), phase);
    phase ^= 1;
    tc_fence_after();
    // ---- 4. epilogue: the A tiles are free again (the MMAs have completed): tile mt doubles as the bf16 staging tile of its own output
    for (int mt = warp >> 2; mt < mtiles; mt += 2) {
      const int row = quarter * 32 + lane;
      const int vox = mt * 128 + row;
      const int yl = vox / p.d2, x = vox - yl * p.d2;
      const bool live = y0 + yl < p.d1;
      const int64_t grow = (((int64_t)b * p.d0 + z) * p.d1 + (y0 + yl)) * p.d2 + x;
      uint8_t* stage = a_s + (size_t)mt * 16384;
      for (int c64 = 0; c64 < p.c_out; c64 += 64) {
#pragma unroll
        for (int c32 = 0; c32 < 2; ++c32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * p.c_out + c64 + c32 * 32), r);
          tmem_ld_wait();
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[2 * j]) + s_bias[c64 + c32 * 32 + 2 * j],
                                                      __uint_as_float(r[2 * j + 1]) + s_bias[c64 + c32 * 32 + 2 * j + 1]);
            packed[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          if (live) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + grow * p.ld_out + c64 + c32 * 32);
#pragma unroll
            for (int u = 0; u < 4; ++u) dst[u] = make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
          }
          if (p.stats) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int ch16 = (c32 * 4 + u) ^ (row & 7);
              *reinterpret_cast<uint4*>(stage + (size_t)row * 128 + (ch16 << 4)) =
                  live ? make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]) : make_uint4(0, 0, 0, 0);
            }
          }
        }
        if (p.stats) {
          // this warp's 32 rows x 64 channels, read back transposed: lane -> channels 2 * lane, 2 * lane + 1 (what the next GroupNorm
          // will read: the stored, rounded values)
          __syncwarp();
          const int h0 = (c64 >> 6) * 2;
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const int r2 = quarter * 32 + rr;
            const uint32_t v = *reinterpret_cast<const uint32_t*>(stage + (size_t)r2 * 128 + ((((lane >> 2) ^ (r2 & 7))) << 4) + ((lane & 3) << 2));
            const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
            st_s[h0] += lo; st_q[h0] = fmaf(lo, lo, st_q[h0]);
            st_s[h0 + 1] += hi; st_q[h0 + 1] = fmaf(hi, hi, st_q[h0 + 1]);
          }
          __syncwarp();
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM drained and the staging tiles read before the next item overwrites them
    tc_fence_after();
  }
  if (p.stats) {
    if (cur_b >= 0) flush(cur_b);
    stats_group_tail(p.sink, p.stats, p.n, nblk, p.c_out, (int)blockIdx.x, 1, (int)threadIdx.x, IT_THREADS, reinterpret_cast<int*>(s_red),
                     [] { __syncthreads(); });
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

}  // namespace diqt

using namespace diqt;

extern "C" int diqt_init_conv_tc_supported(int c_in, int c_out, int d1, int d2) {
  return c_in > 0 && c_in <= IT_MAX_CIN && 27 * c_in <= 64 && (c_out == 64 || c_out == 128) && (IT_TY * d2) % 128 == 0 && IT_TY * d2 * c_out / 128 <= 512 &&
         d2 <= 128;
}

extern "C" int diqt_init_conv_tc_blocks(int n, int d0, int d1, int* nblk) {
  DIQT_REQUIRE(nblk, "init_conv_tc_blocks: null output");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t items = (int64_t)n * d0 * ((d1 + IT_TY - 1) / IT_TY);
  *nblk = (int)(items < 2 * sms ? items : 2 * sms);
  return DIQT_OK;
}

extern "C" int diqt_init_conv_tc(const float* const* planes, const int64_t* plane_stride, int c_in, const void* w_packed, const float* bias, void* out,
                                 int ld_out, int n, int d0, int d1, int d2, int c_out, float* partial, float* group, uint32_t* tickets, void* stream) {
  DIQT_REQUIRE(planes && plane_stride && w_packed && bias && out, "init_conv_tc: null pointer");
  DIQT_REQUIRE(diqt_init_conv_tc_supported(c_in, c_out, d1, d2), "init_conv_tc: needs 27 * c_in <= 64, c_out in {64, 128}, d2 %% 16 == 0 (c_in=%d c_out=%d d2=%d)",
               c_in, c_out, d2);
  DIQT_REQUIRE(ld_out % 8 == 0 && ld_out >= c_out, "init_conv_tc: ld_out=%d", ld_out);
  DIQT_REQUIRE(!group || (partial && tickets), "init_conv_tc: the grouped sink needs partial rows and tickets");
  InitTcParams p = {};
  for (int i = 0; i < IT_MAX_CIN; ++i) {
    p.plane[i] = i < c_in ? planes[i] : nullptr;
    p.stride[i] = i < c_in ? plane_stride[i] : 0;
  }
  p.w = (const uint8_t*)w_packed;
  p.bias = bias;
  p.out = (__nv_bfloat16*)out;
  p.stats = partial;
  p.c_in = c_in; p.c_out = c_out; p.ld_out = ld_out; p.n = n; p.d0 = d0; p.d1 = d1; p.d2 = d2;
  p.ytiles = (d1 + IT_TY - 1) / IT_TY;
  p.items = n * d0 * p.ytiles;
  p.idesc = make_idesc_bf16(128, c_out);
  int nblk = 0;
  int rc = diqt_init_conv_tc_blocks(n, d0, d1, &nblk);
  if (rc) return rc;
  p.sink.group = group;
  p.sink.tickets = tickets;
  p.sink.gsize = stats_group_size(nblk, 1);
  p.sink.ngroups = (nblk + p.sink.gsize - 1) / p.sink.gsize;
  const int mtiles = IT_TY * d2 / 128;
  const size_t smem = (size_t)mtiles * 16384 + (size_t)c_out * 128 + (size_t)c_in * 3 * (IT_TY + 2) * (d2 + 2) * 4 + 64 * 4 + (size_t)c_out * 4 +
                      (size_t)8 * c_out * 2 * 4 + 64 + 1024;
  DIQT_REQUIRE(smem <= 227 * 1024, "init_conv_tc: %zu bytes of shared memory", smem);
  static bool attr_done = false;
  if (!attr_done) {
    DIQT_CUDA(cudaFuncSetAttribute(init_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  // first kernel of a forward: its inputs come from launches this library does not control, so no programmatic dependent launch
  init_conv_tc_kernel<<<nblk, IT_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("init_conv_tc");
}
