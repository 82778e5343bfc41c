// init_conv on the tensor cores in ONE kernel (Unet.init_conv, imagen_pytorch3D.py:1291: Conv3d(c_in, init_dim, 3, padding 1) over the
// concatenation [cond_images, x, lowres_cond_img] of :1569-1584).
//
// K = 27 * c_in (54 for the driver's two input channels) is not a tensor-core shape as a 3x3x3 convolution, but it is as a GEMM over
// im2col rows of K = 64 bf16 columns.  Round 1 materialised those rows in global memory (128 bytes per voxel written and read back:
// init_im2col_kernel 29 us + the 1x1x1 tcgen05 conv 26 us at 64^3).  Here the rows never leave the SM.  A persistent CTA (one per SM)
// walks work items of ty (y) x d2 (x) voxels of one z-plane through a three-role pipeline with double-buffered A tiles and accumulators:
//   producers (4 warps)  cp.async the haloed fp32 input tile of item i+1 (zero fill outside the volume = the padding) while they write the
//                        im2col rows of item i as bf16 straight into 128B-swizzled A tiles (128 rows x 64 columns each),
//   issuer (1 warp)      tcgen05.mma (M128, N = c_out, K64) per A tile against the packed weights into tensor memory,
//   epilogue (4 warps)   drains TMEM: bias, bf16, a swizzled staging tile per warp, 512-byte coalesced row stores, and the channel sums /
//                        sums of squares the first GroupNorm needs, accumulated over the CTA's whole life and written as ONE partial row
//                        per CTA and volume (+ grouped sink).
// The first version ran the four phases one after the other behind __syncthreads (57 us at 64^3 for 33 MB of traffic); the roles now overlap.
// HBM traffic: 8 bytes in + 128 bytes out per voxel instead of 8 + 128 + 128 + 128.
#include <stdlib.h>

#include "tc_common.cuh"

namespace diqt {

constexpr int IT_MAX_CIN = 2;                   // 27 * c_in <= 64 im2col columns
constexpr int IT_MAX_SEG = 6;                   // 16-byte input segments a producer thread copies per work item
// One warp issues an instruction every ~6 cycles here (profiles/r5b_ncu_init_conv.md): the first pipelined version, with four producer and
// four epilogue warps, was bound by exactly that (56 us).  Seven and eight of them give every scheduler four busy warps.
constexpr int IT_PROD_WARPS = 7;                 // warps 0-6: input tile + im2col rows (16 warps in all: ptxas sizes the register file for blocks of 128 threads)
constexpr int IT_ISSUE_WARP = IT_PROD_WARPS;     // warp 7: TMEM owner + MMA issuer
constexpr int IT_EPI_WARP0 = IT_PROD_WARPS + 1;  // warps 8-15: epilogue (TMEM lane quarter = warp & 3; two warps per quarter split the A tiles)
constexpr int IT_EPI_WARPS = 8;
constexpr int IT_THREADS = (IT_PROD_WARPS + 1 + IT_EPI_WARPS) * 32;
constexpr int IT_PROD_THREADS = IT_PROD_WARPS * 32;
constexpr int IT_EPI_THREADS = IT_EPI_WARPS * 32;

struct InitTcParams {
  const float* plane[IT_MAX_CIN];
  long long stride[IT_MAX_CIN];
  const uint8_t* w;    // [c_out][64] bf16, rows pre-swizzled (16-byte chunk ^ (row & 7)); column k = tap * c_in + ci, zero padded to 64
  const float* bias;   // [c_out]
  __nv_bfloat16* out;  // channels-last rows, pitch ld_out
  float* stats;        // NULL or partial[n][gridDim.x][c_out][2]
  StatsGroups sink;
  int c_in, c_out, ld_out, n, d0, d1, d2, ty, ytiles, items;
  int d2_shift;        // log2(d2), or -1 when d2 is not a power of two
  uint32_t idesc, tmem_cols;
};

// y rows per work item: the largest of 8, 4, 2, 1 whose ty * d2 voxels are whole 128-row A tiles, whose two accumulator sets fit the 512
// TMEM columns and whose buffers fit shared memory; 0 = shape not supported
__host__ __device__ inline size_t init_tc_smem_bytes(int c_in, int c_out, int d2, int ty) {
  const int mtiles = ty * d2 / 128;
  return (size_t)2 * mtiles * 16384 + (size_t)c_out * 128 + (size_t)IT_EPI_WARPS * 4096 + (size_t)2 * c_in * 3 * (ty + 2) * (d2 + 8) * 4 + 64 * 4 +
         (size_t)c_out * 4 + (size_t)IT_EPI_WARPS * c_out * 2 * 4 + 128 + 1024;
}
static int init_tc_ty(int c_in, int c_out, int d2) {
  static int forced = -1;   // DIQT_INIT_TY: A/B measurements
  if (forced < 0) {
    const char* e = getenv("DIQT_INIT_TY");
    forced = e ? atoi(e) : 0;
  }
  for (int ty = 8; ty >= 1; ty >>= 1) {
    if (forced > 0 && ty > forced) continue;
    const int rows = ty * d2;
    if (rows % 128 != 0) continue;
    if (rows / 128 * c_out > 256) continue;
    if (init_tc_smem_bytes(c_in, c_out, d2, ty) > (size_t)227 * 1024) continue;
    if (c_in * 3 * (ty + 2) * (d2 / 4) > IT_MAX_SEG * IT_PROD_THREADS) continue;
    return ty;
  }
  return 0;
}

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const float* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;   // src-size 0: nothing is read, the sixteen bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(IT_THREADS, 1) init_conv_tc_kernel(const __grid_constant__ InitTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // input tile rows: pitch W = d2 + 8 floats, voxel x at index x + 4 (16-byte aligned for the cp.async segments), the x halo at 3 and d2 + 4
  const int W = p.d2 + 8, TYH = p.ty + 2, plane = 3 * TYH * W;
  const int rows = p.ty * p.d2, mtiles = rows / 128;
  const int tile_elems = p.c_in * plane;
  uint8_t* a_s = smem;                                                  // [2][mtiles][128 rows][128 B]
  uint8_t* w_s = a_s + (size_t)2 * mtiles * 16384;                      // [c_out][128 B]
  uint8_t* stage = w_s + (size_t)p.c_out * 128;                         // [8 epilogue warps][32 rows][128 B]
  float* tile = reinterpret_cast<float*>(stage + IT_EPI_WARPS * 4096);  // [2][c_in][3][ty + 2][W]
  int* koff = reinterpret_cast<int*>(tile + (size_t)2 * tile_elems);    // [64]
  float* s_bias = reinterpret_cast<float*>(koff + 64);                  // [c_out]
  float* s_red = s_bias + p.c_out;                                      // [8 warps][c_out][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + IT_EPI_WARPS * p.c_out * 2);
  uint64_t* a_full = bars;        // [2] im2col rows written (one arrival per producer warp)
  uint64_t* a_empty = bars + 2;   // [2] the MMAs have read the A tiles (tcgen05.commit)
  uint64_t* acc_full = bars + 4;  // [2] accumulators complete (tcgen05.commit)
  uint64_t* acc_empty = bars + 6; // [2] accumulators drained (one arrival per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) {
    const int k = threadIdx.x;
    int o = -1;
    if (k < 27 * p.c_in) {
      const int tap = k / p.c_in, ci = k - tap * p.c_in;
      o = ci * plane + ((tap / 9) * TYH + (tap / 3) % 3) * W + tap % 3 + 3;
    }
    koff[k] = o;
  }
  for (int i = threadIdx.x; i < p.c_out; i += IT_THREADS) s_bias[i] = p.bias[i];
  for (int i = threadIdx.x; i < p.c_out * 8; i += IT_THREADS) reinterpret_cast<uint4*>(w_s)[i] = reinterpret_cast<const uint4*>(p.w)[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&a_full[i]), IT_PROD_WARPS);
      mbar_init(smem_u32(&a_empty[i]), 1);
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), IT_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == IT_ISSUE_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();  // the weights were written with generic stores and are read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int acc_cols = mtiles * p.c_out;   // columns of one accumulator set
  const int nblk = (int)gridDim.x;

  auto decode = [&](int item, int& b, int& z, int& y0) {
    const int yt = item % p.ytiles;
    const int t = item / p.ytiles;
    z = t % p.d0;
    b = t / p.d0;
    y0 = yt * p.ty;
  };

  if (warp < IT_PROD_WARPS) {
    // ===================== producers: haloed input tile (cp.async, one item ahead) + im2col rows =====================
    const int pt = threadIdx.x;  // 0..223
    auto sync_prod = [] { asm volatile("bar.sync 2, %0;" ::"n"(IT_PROD_THREADS) : "memory"); };
    // The tile spans whole x rows, so its x halo is always outside the volume: both tile buffers are zeroed once and only the interior
    // is rewritten per item.  The im2col columns k >= 27 * c_in are zero padding of K: the A buffers are zeroed once as well and whole
    // padding chunks are never written again.
    for (int i = pt; i < 2 * tile_elems; i += IT_PROD_THREADS) tile[i] = 0.f;
    for (int i = pt; i < 2 * mtiles * 1024; i += IT_PROD_THREADS) sts_128(smem_u32(a_s) + (uint32_t)i * 16, 0u, 0u, 0u, 0u);
    // 16-byte segments of the tile interior this thread copies for every item: (row r = (ci, dz, yy), four voxels at x4 * 4)
    const int q = p.d2 >> 2, nseg = p.c_in * 3 * TYH * q;
    int seg_src[IT_MAX_SEG];       // element offset from the item's (z, y0) row of plane ci
    uint32_t seg_dst[IT_MAX_SEG];  // byte offset inside a tile buffer
    int seg_key[IT_MAX_SEG];       // ci << 16 | dz << 8 | yy;  -1 = no segment
#pragma unroll
    for (int sidx = 0; sidx < IT_MAX_SEG; ++sidx) {
      const int idx = pt + sidx * IT_PROD_THREADS;
      seg_key[sidx] = -1; seg_src[sidx] = 0; seg_dst[sidx] = 0;
      if (idx < nseg) {
        const int r = idx / q, x4 = idx - r * q;
        const int ci = r / (3 * TYH), rr = r - ci * 3 * TYH;
        const int dz = rr / TYH, yy = rr - dz * TYH;
        seg_key[sidx] = ci << 16 | dz << 8 | yy;
        seg_src[sidx] = ((dz - 1) * p.d1 + (yy - 1)) * p.d2 + x4 * 4;
        seg_dst[sidx] = (uint32_t)(r * W + 4 + x4 * 4) * 4;
      }
    }
    auto stage_in = [&](int item, int buf) {
      int b, z, y0;
      decode(item, b, z, y0);
      const uint32_t dst = smem_u32(tile + (size_t)buf * tile_elems);
      const int64_t row = ((int64_t)z * p.d1 + y0) * p.d2;
      const float* base0 = p.plane[0] + (int64_t)b * p.stride[0] + row;
      const float* base1 = p.plane[1] + (int64_t)b * p.stride[1] + row;   // (unused when c_in = 1: no segment has ci = 1)
#pragma unroll
      for (int sidx = 0; sidx < IT_MAX_SEG; ++sidx) {
        const int key = seg_key[sidx];
        if (key < 0) continue;
        const int zz = z + ((key >> 8) & 0xff) - 1, y = y0 + (key & 0xff) - 1;
        const bool ok = zz >= 0 && zz < p.d0 && y >= 0 && y < p.d1;
        const float* src = ((key >> 16) ? base1 : base0) + seg_src[sidx];
        cp_async16_zfill(dst + seg_dst[sidx], ok ? src : p.plane[0], ok);
      }
    };
    // im2col: a warp builds ONE 16-byte chunk (eight columns k) of 32 consecutive voxels per step, so the eight tile offsets are the same
    // for all lanes (conflict-free shared loads) and, with as many producer warps as non-padding chunks, loop invariant
    const int nchunks = (27 * p.c_in + 7) >> 3;
    const int nunits = nchunks * (rows >> 5);
    int cur_chunk = -1;
    uint32_t offb[8];   // byte offsets of the chunk's eight columns inside the input tile
    uint32_t vmask = 0; // columns that exist (k < 27 * c_in)
    sync_prod();        // the zero fill above is complete before the first cp.async lands in the tile
    int k = 0;
    stage_in((int)blockIdx.x, 0);
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++k) {
      const int buf = k & 1;
      cp_async_wait_all();
      sync_prod();  // tile k complete for every producer; everybody is done reading tile k - 1
      if (item + (int)gridDim.x < p.items) stage_in(item + (int)gridDim.x, buf ^ 1);
      mbar_wait(smem_u32(&a_empty[buf]), (uint32_t)(((k >> 1) & 1) ^ 1));  // the MMAs of item k - 2 have read this A buffer
      const uint32_t src = smem_u32(tile + (size_t)buf * tile_elems);
      const uint32_t a_buf = smem_u32(a_s + (size_t)buf * mtiles * 16384);
      int chunk = warp, vb = 0;
      while (chunk >= nchunks) { chunk -= nchunks; ++vb; }
      for (int u = warp; u < nunits; u += IT_PROD_WARPS) {
        if (chunk != cur_chunk) {
          cur_chunk = chunk;
          vmask = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int o = koff[chunk * 8 + j];
            offb[j] = o >= 0 ? (uint32_t)o * 4 : 0u;
            if (o >= 0) vmask |= 1u << j;
          }
        }
        const int vox = vb * 32 + lane;
        int yl, x;
        if (p.d2_shift >= 0) { yl = vox >> p.d2_shift; x = vox & (p.d2 - 1); }
        else { yl = vox / p.d2; x = vox - yl * p.d2; }
        const uint32_t base = src + (uint32_t)(yl * W + x) * 4;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        if (vmask == 0xffu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = lds_f32(base + offb[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if ((vmask >> j) & 1u) v[j] = lds_f32(base + offb[j]);
        }
        // row r of A tile mt, 16-byte chunk XOR-swizzled by (r & 7): (vox & 127) * 128 + mt * 16384 = vox * 128
        sts_128(a_buf + (uint32_t)vox * 128 + (uint32_t)((chunk ^ (vox & 7)) << 4), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                pack_bf16x2(v[6], v[7]));
        chunk += IT_PROD_WARPS;
        while (chunk >= nchunks) { chunk -= nchunks; ++vb; }
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&a_full[buf]));
    }
    cp_async_wait_all();
  } else if (warp == IT_ISSUE_WARP) {
    // ===================== MMA issuer (warp-uniform code; the issuing lane is elected inside the wrappers) =====================
    const uint64_t bdesc = make_sw128_desc(smem_u32(w_s));
    int k = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++k) {
      const int buf = k & 1;
      const uint32_t ph = (uint32_t)((k >> 1) & 1);
      mbar_wait(smem_u32(&acc_empty[buf]), ph ^ 1);  // accumulator set drained (item k - 2)
      mbar_wait(smem_u32(&a_full[buf]), ph);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < mtiles; ++mt) {
        const uint64_t adesc = make_sw128_desc(smem_u32(a_s + ((size_t)buf * mtiles + mt) * 16384));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_base + (uint32_t)(buf * acc_cols + mt * p.c_out), adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), p.idesc, kk != 0);
      }
      umma_commit(smem_u32(&a_empty[buf]));
      umma_commit(smem_u32(&acc_full[buf]));
    }
  } else {
    // ===================== epilogue (eight warps; warp e drains TMEM lane quarter (warp & 3) of the A tiles e / 4, e / 4 + 2, ...) ==========
    const int quarter = warp & 3;
    const int et = threadIdx.x - IT_EPI_WARP0 * 32;  // 0..255
    const int ew = et >> 5, half = ew >> 2;
    auto sync_epi = [] { asm volatile("bar.sync 1, %0;" ::"n"(IT_EPI_THREADS) : "memory"); };
    const uint32_t my_stage = smem_u32(stage + (size_t)ew * 4096);  // this warp's 32 rows x 128 B
    const uint32_t bias_a = smem_u32(s_bias);
    // statistics of the rows this warp drains, in the copy-out mapping: lane -> rows (lane >> 3) + 4 i, channels (lane & 7) * 8 + e (+ 64 g)
    float st_s[2][8], st_q[2][8];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int e = 0; e < 8; ++e) st_s[g][e] = st_q[g][e] = 0.f;
    int cur_b = -1;

    auto flush = [&](int b) {
      // every lane parks its sums in its warp's staging rows (idle between items) as [row group = lane >> 3][channel][2]; then one thread
      // per channel adds the 8 warps x 4 row groups in index order -> one partial row of this CTA for volume b
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g * 64 >= p.c_out) break;
        const uint32_t dst = my_stage + (uint32_t)(((lane >> 3) * p.c_out + g * 64 + (lane & 7) * 8) * 8);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          sts_128(dst + (uint32_t)e * 8, __float_as_uint(st_s[g][e]), __float_as_uint(st_q[g][e]), __float_as_uint(st_s[g][e + 1]), __float_as_uint(st_q[g][e + 1]));
          st_s[g][e] = st_q[g][e] = st_s[g][e + 1] = st_q[g][e + 1] = 0.f;
        }
      }
      sync_epi();
      float* dst = p.stats + ((size_t)b * nblk + blockIdx.x) * p.c_out * 2;
      for (int ch = et; ch < p.c_out; ch += IT_EPI_THREADS) {
        float a = 0.f, q = 0.f;
#pragma unroll 4
        for (int wr = 0; wr < IT_EPI_WARPS * 4; ++wr) {
          const float2 v = *reinterpret_cast<const float2*>(stage + (size_t)(wr >> 2) * 4096 + (size_t)(((wr & 3) * p.c_out + ch) * 8));
          a += v.x; q += v.y;
        }
        dst[ch * 2] = a;
        dst[ch * 2 + 1] = q;
      }
      sync_epi();
    };

    if (p.stats) {  // every (volume, channel) entry of this CTA's row must be defined even if the CTA never sees that volume
      for (int b = 0; b < p.n; ++b)
        for (int idx = et; idx < p.c_out * 2; idx += IT_EPI_THREADS) p.stats[((size_t)b * nblk + blockIdx.x) * p.c_out * 2 + idx] = 0.f;
      sync_epi();
    }
    // staging addresses of this lane: its own row when draining TMEM; (row, chunk) = (4 i + lane / 8, lane & 7) when copying out
    const uint32_t st_row = my_stage + (uint32_t)lane * 128, sw = (uint32_t)(lane & 7);
    const uint32_t cp_addr = my_stage + (uint32_t)(lane >> 3) * 128 + ((sw ^ (uint32_t)(lane >> 3)) << 4);
    int k = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++k) {
      const int buf = k & 1;
      int b, z, y0;
      decode(item, b, z, y0);
      if (p.stats && b != cur_b) {
        if (cur_b >= 0) flush(cur_b);
        cur_b = b;
      }
      const int live_rows = min(p.d1 - y0, p.ty) * p.d2;                     // rows of this item inside the volume (whole x rows)
      const int64_t row0 = (((int64_t)b * p.d0 + z) * p.d1 + y0) * p.d2;       // global voxel row of the item's row 0
      mbar_wait(smem_u32(&acc_full[buf]), (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      for (int mt = half; mt < mtiles; mt += 2) {
        const int wrow0 = mt * 128 + quarter * 32;   // first item row of this warp's 32
        if (wrow0 >= live_rows) continue;             // (whole warp outside the volume: nothing to store, nothing to count)
        const bool all_live = wrow0 + 32 <= live_rows, live = wrow0 + lane < live_rows;
        __nv_bfloat16* orow = p.out + (row0 + wrow0 + (lane >> 3)) * p.ld_out + (lane & 7) * 8;
#pragma unroll
        for (int g64 = 0; g64 < 2; ++g64) {   // (static indices into the statistics registers)
          const int c64 = g64 * 64;
          if (c64 >= p.c_out) break;
#pragma unroll
          for (int c32 = 0; c32 < 2; ++c32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * acc_cols + mt * p.c_out + c64 + c32 * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 b0 = lds_128(bias_a + (uint32_t)(c64 + c32 * 32 + u * 8) * 4), b1 = lds_128(bias_a + (uint32_t)(c64 + c32 * 32 + u * 8 + 4) * 4);
              // (channel pairs per FADD2: the same IEEE additions per lane)
              const float2 s0 = __fadd2_rn(make_float2(__uint_as_float(r[u * 8]), __uint_as_float(r[u * 8 + 1])), make_float2(__uint_as_float(b0.x), __uint_as_float(b0.y)));
              const float2 s1 = __fadd2_rn(make_float2(__uint_as_float(r[u * 8 + 2]), __uint_as_float(r[u * 8 + 3])), make_float2(__uint_as_float(b0.z), __uint_as_float(b0.w)));
              const float2 s2 = __fadd2_rn(make_float2(__uint_as_float(r[u * 8 + 4]), __uint_as_float(r[u * 8 + 5])), make_float2(__uint_as_float(b1.x), __uint_as_float(b1.y)));
              const float2 s3 = __fadd2_rn(make_float2(__uint_as_float(r[u * 8 + 6]), __uint_as_float(r[u * 8 + 7])), make_float2(__uint_as_float(b1.z), __uint_as_float(b1.w)));
              uint32_t q0 = pack_bf16x2(s0.x, s0.y), q1 = pack_bf16x2(s1.x, s1.y), q2 = pack_bf16x2(s2.x, s2.y), q3 = pack_bf16x2(s3.x, s3.y);
              if (!all_live) {   // warp-uniform; rows outside the volume: zeros, so the statistics may read them
                if (!live) q0 = q1 = q2 = q3 = 0u;
              }
              sts_128(st_row + ((((uint32_t)(c32 * 4 + u)) ^ sw) << 4), q0, q1, q2, q3);
            }
          }
          __syncwarp();
          // 32 rows x 128 B of this warp -> global memory, four whole rows (512 contiguous bytes when ld_out = 64) per instruction; the
          // same registers feed the channel sums (what the next GroupNorm will read: the stored, rounded values)
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            // row it * 4 + (lane >> 3): its swizzle term is ((it & 1) * 4 + (lane >> 3)) & 7 = (lane >> 3) ^ ((it & 1) * 4)
            const uint4 v = lds_128((cp_addr + (uint32_t)it * 512) ^ (uint32_t)((it & 1) << 6));
            if (all_live || wrow0 + it * 4 + (lane >> 3) < live_rows)
              *reinterpret_cast<uint4*>(orow + (int64_t)(it * 4) * p.ld_out + c64) = v;
            if (p.stats) {
              const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) {
                const float2 lh = make_float2(__uint_as_float(w4[e2] << 16), __uint_as_float(w4[e2] & 0xffff0000u));
                const float2 ns = __fadd2_rn(make_float2(st_s[g64][2 * e2], st_s[g64][2 * e2 + 1]), lh);
                const float2 nq = __ffma2_rn(lh, lh, make_float2(st_q[g64][2 * e2], st_q[g64][2 * e2 + 1]));
                st_s[g64][2 * e2] = ns.x; st_s[g64][2 * e2 + 1] = ns.y; st_q[g64][2 * e2] = nq.x; st_q[g64][2 * e2 + 1] = nq.y;
              }
            }
          }
          __syncwarp();  // the staging rows are free again
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
    }
    if (p.stats) {
      if (cur_b >= 0) flush(cur_b);
      stats_group_tail(p.sink, p.stats, p.n, nblk, p.c_out, (int)blockIdx.x, 1, et, IT_EPI_THREADS, reinterpret_cast<int*>(s_red), sync_epi);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == IT_ISSUE_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace diqt

using namespace diqt;

extern "C" int diqt_init_conv_tc_supported(int c_in, int c_out, int d1, int d2) {
  (void)d1;
  return c_in > 0 && c_in <= IT_MAX_CIN && 27 * c_in <= 64 && (c_out == 64 || c_out == 128) && d2 <= 128 && init_tc_ty(c_in, c_out, d2) > 0;
}

// rows of the statistics partial = CTAs of the launch: one per SM, never more than the work items of the coarsest tiling (8 y rows)
extern "C" int diqt_init_conv_tc_blocks(int n, int d0, int d1, int* nblk) {
  DIQT_REQUIRE(nblk, "init_conv_tc_blocks: null output");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t items = (int64_t)n * d0 * ((d1 + 7) / 8);
  *nblk = (int)(items < sms ? items : sms);
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
  p.ty = init_tc_ty(c_in, c_out, d2);
  p.d2_shift = -1;
  for (int sh = 0; sh < 16; ++sh)
    if ((1 << sh) == d2) p.d2_shift = sh;
  for (int i = 0; i < c_in; ++i)
    DIQT_REQUIRE(((uintptr_t)planes[i] & 15) == 0 && plane_stride[i] % 4 == 0, "init_conv_tc: input planes must be 16-byte aligned (plane %d)", i);
  p.ytiles = (d1 + p.ty - 1) / p.ty;
  p.items = n * d0 * p.ytiles;
  p.idesc = make_idesc_bf16(128, c_out);
  const int mtiles = p.ty * d2 / 128;
  const int cols = 2 * mtiles * c_out;
  p.tmem_cols = cols <= 32 ? 32u : cols <= 64 ? 64u : cols <= 128 ? 128u : cols <= 256 ? 256u : 512u;
  int nblk = 0;
  int rc = diqt_init_conv_tc_blocks(n, d0, d1, &nblk);
  if (rc) return rc;
  p.sink.group = group;
  p.sink.tickets = tickets;
  p.sink.gsize = stats_group_size(nblk, 1);
  p.sink.ngroups = (nblk + p.sink.gsize - 1) / p.sink.gsize;
  const size_t smem = init_tc_smem_bytes(c_in, c_out, d2, p.ty);
  static bool attr_done = false;
  if (!attr_done) {
    DIQT_CUDA(cudaFuncSetAttribute(init_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  // first kernel of a forward: its inputs come from launches this library does not control, so no programmatic dependent launch
  init_conv_tc_kernel<<<nblk, IT_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("init_conv_tc");
}
