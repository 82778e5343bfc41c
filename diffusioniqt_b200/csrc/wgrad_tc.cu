// Convolution weight gradient on the 5th-gen tensor cores:  dw[co][ci][tap] = sum_voxels dy[v][co] * x[v + tap][ci]
// (the reverse of nn.Conv3d in Block.project imagen_pytorch3D.py:550, res_conv :597, the 1x1x1 convs :1388 / :459-496), bf16 operands,
// fp32 accumulation in TMEM, channel counts that are multiples of 64.
//
// The reduction runs over VOXELS, and both operands are stored voxel-row by voxel-row ([voxels][channels], channels contiguous): exactly
// the MN-major SWIZZLE_128B operand layout (K = voxel rows in 8-row groups 1024 B apart, 64 channels = one 128-byte row; see
// linattn_tc.cu).  So a [128 voxels][64 channels] TMA box of dy is the B operand (N = c_out) and a box of x at the TAP-SHIFTED
// coordinate is the A operand (M = c_in; out-of-volume rows arrive as zeros = the conv's zero padding), with no transposes anywhere.
// M = 128 stacks two (tap, 64-channel chunk) units; a CTA owns up to two such M-blocks (four x boxes per voxel tile against one load
// of the dy boxes), a chunk of the voxel tiles and an N tile of <= 128 output channels:
//     D_mb[128 x N] += X_mb^T dY          (8 tcgen05.mma of K = 16 voxels per 128-voxel tile and M-block)
// Warp roles (192 threads): 0-3 epilogue (TMEM -> per-chunk partials), 4 TMA producer, 5 MMA issuer.  The partials of the chunks are
// summed in a fixed order by wgrad_tc_reduce_kernel (bitwise reproducible), which also writes the nn.Conv3d.weight layout.
#include <string.h>

#include <algorithm>
#include <new>

#include "tc_common.cuh"

namespace diqt {

constexpr int WT_BOX = 128 * 128;   // [128 voxels][64 bf16]
constexpr int WT_STAGES = 2;
constexpr int WT_THREADS = 192;

struct WgTcParams {
  CUtensorMap x_map, dy_map;
  float* partial;   // [chunk][unit][64 ci][c_out]
  int taps, KC, units, nmb, NB, N, c_in, c_out;
  int bx, by, bz, bn, tiles_x, tiles_y, tiles_z, tiles_n, m_tiles, nchunks;
  uint32_t idesc;
};

__device__ __forceinline__ uint64_t wt_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = (p.NB + 4) * WT_BOX;          // dy boxes, then four x boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WT_STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + WT_STAGES;
  uint64_t* d_full = bars + 2 * WT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x, mb0 = blockIdx.y * 2, ntile = blockIdx.z;
  const int nmb_here = min(2, p.nmb - mb0);              // M-blocks of this CTA
  const int t0 = (int)((int64_t)chunk * p.m_tiles / p.nchunks), t1 = (int)((int64_t)(chunk + 1) * p.m_tiles / p.nchunks);
  const uint32_t tmem_cols = 2 * p.N <= 128 ? 128u : (2 * p.N <= 256 ? 256u : 512u);

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < WT_STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    mbar_init(smem_u32(d_full), 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      pdl_sync();   // x and dy come from the predecessors
      // the (tap, channel chunk) units of this CTA's x boxes; a unit past the end repeats the last one (its rows are never stored)
      int ucx[4], ux[4], uy[4], uz[4];
      for (int j = 0; j < 4; ++j) {
        const int u = min(mb0 * 2 + j, p.units - 1);
        const int tap = u / p.KC, kc = u - tap * p.KC;
        ucx[j] = kc * 64;
        uz[j] = p.taps == 27 ? tap / 9 - 1 : 0;
        uy[j] = p.taps == 27 ? (tap / 3) % 3 - 1 : 0;
        ux[j] = p.taps == 27 ? tap % 3 - 1 : 0;
      }
      const uint32_t tx_bytes = (uint32_t)(p.NB + 2 * nmb_here) * WT_BOX;
      for (int t = t0; t < t1; ++t) {
        const int it = t - t0, s = it % WT_STAGES;
        const uint32_t ph = (uint32_t)((it / WT_STAGES) & 1);
        int r = t;
        const int tx = r % p.tiles_x; r /= p.tiles_x;
        const int ty = r % p.tiles_y; r /= p.tiles_y;
        const int tz = r % p.tiles_z; r /= p.tiles_z;
        const int x0 = tx * p.bx, y0 = ty * p.by, z0 = tz * p.bz, n0 = r * p.bn;
        mbar_wait(smem_u32(&empty[s]), ph ^ 1);
        const uint32_t bar = smem_u32(&full[s]), dst = smem_u32(smem + s * stage_bytes);
        mbar_expect_tx(bar, tx_bytes);
        for (int nb = 0; nb < p.NB; ++nb) tma_load_5d(dst + nb * WT_BOX, &p.dy_map, bar, (ntile * p.NB + nb) * 64, x0, y0, z0, n0);
        for (int j = 0; j < 2 * nmb_here; ++j)
          tma_load_5d(dst + (p.NB + j) * WT_BOX, &p.x_map, bar, ucx[j], x0 + ux[j], y0 + uy[j], z0 + uz[j], n0);
      }
    }
  } else if (warp == 5) {
    for (int t = t0; t < t1; ++t) {
      const int it = t - t0, s = it % WT_STAGES;
      const uint32_t ph = (uint32_t)((it / WT_STAGES) & 1);
      mbar_wait(smem_u32(&full[s]), ph);
      tc_fence_after();
      const uint32_t dyb = smem_u32(smem + s * stage_bytes), xb = dyb + p.NB * WT_BOX;
      for (int m = 0; m < nmb_here; ++m) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {   // K = 16 voxels per instruction
          const uint64_t adesc = wt_mn_desc(xb + m * 2 * WT_BOX + kk * 2048, WT_BOX, 1024);
          const uint64_t bdesc = wt_mn_desc(dyb + kk * 2048, WT_BOX, 1024);
          umma_bf16(tmem_base + (uint32_t)(m * p.N), adesc, bdesc, p.idesc, (uint32_t)((it | kk) != 0));
        }
      }
      umma_commit(smem_u32(&empty[s]));
    }
    umma_commit(smem_u32(d_full));
  } else {
    // epilogue: lane = row (unit within the M-block, input channel), columns = output channels of the N tile
    const int row = threadIdx.x;
    mbar_wait(smem_u32(d_full), 0);
    tc_fence_after();
    for (int m = 0; m < nmb_here; ++m) {
      const int u = (mb0 + m) * 2 + (row >> 6);
      float* dst = p.partial + (((size_t)chunk * p.units + min(u, p.units - 1)) * 64 + (row & 63)) * p.c_out + ntile * p.N;
      for (int c32 = 0; c32 < p.N / 32; ++c32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(m * p.N + c32 * 32), r);
        tmem_ld_wait();
        if (u < p.units) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            reinterpret_cast<float4*>(dst + c32 * 32)[q] =
                make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// dw[co][ci][tap] = sum_chunks partial[chunk][tap * KC + ci / 64][ci % 64][co]   (fixed order)
__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const float* partial, int nchunks, int taps, int KC, int c_in, int c_out, float* dw) {
  pdl_sync();
  const int64_t count = (int64_t)taps * c_in * c_out;
  const int64_t per_chunk = count;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < count; idx += (int64_t)gridDim.x * 256) {
    // idx runs over partial's own order (unit, ci_local, co): coalesced reads
    const int co = (int)(idx % c_out);
    const int64_t r = idx / c_out;
    const int cil = (int)(r % 64), u = (int)(r / 64);
    const int tap = u / KC, ci = (u - tap * KC) * 64 + cil;
    float s = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) s += __ldcg(partial + (int64_t)ch * per_chunk + idx);
    dw[((int64_t)co * c_in + ci) * taps + tap] = s;
  }
}

struct WgTcGeom {
  int bx, by, bz, bn, tiles_x, tiles_y, tiles_z, tiles_n, m_tiles, units, nmb, NB, N, ntile_n, nchunks;
};

static int pow2floor_i(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}

static WgTcGeom wgrad_tc_geom(int n, int d0, int d1, int d2, int c_in, int c_out, int taps) {
  WgTcGeom g;
  int rem = 128;
  g.bx = std::min(8, pow2floor_i(d2)); rem /= g.bx;
  g.by = std::min(std::min(4, rem), pow2floor_i(d1)); rem /= g.by;
  g.bz = std::min(rem, pow2floor_i(d0)); rem /= g.bz;
  g.bn = rem;
  g.tiles_x = (d2 + g.bx - 1) / g.bx;
  g.tiles_y = (d1 + g.by - 1) / g.by;
  g.tiles_z = (d0 + g.bz - 1) / g.bz;
  g.tiles_n = (n + g.bn - 1) / g.bn;
  g.m_tiles = g.tiles_x * g.tiles_y * g.tiles_z * g.tiles_n;
  g.units = taps * (c_in / 64);
  g.nmb = (g.units + 1) / 2;
  g.NB = (c_out % 128 == 0) ? 2 : 1;
  g.N = 64 * g.NB;
  g.ntile_n = c_out / g.N;
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  (void)cudaGetLastError();
  if (sms <= 0) sms = 148;
  const int per_chunk = ((g.nmb + 1) / 2) * g.ntile_n;
  int nc = sms / per_chunk;   // one wave: 22 x 7 = 154 CTAs on 148 SMs ran as two waves and took twice as long (profiles/r3j_ncu_wgrad_tc.csv)
  nc = std::max(1, std::min(nc, g.m_tiles));
  nc = std::min(nc, 64);
  g.nchunks = nc;
  return g;
}

bool wgrad_tc_supported(int dtype, int c_in, int c_out, int ld_x, int ld_dy, int taps) {
  return dtype == DIQT_BF16 && c_in % 64 == 0 && c_out % 64 == 0 && ld_x % 8 == 0 && ld_dy % 8 == 0 && (taps == 1 || taps == 27);
}

size_t wgrad_tc_workspace_bytes(int n, int d0, int d1, int d2, int c_in, int c_out, int taps) {
  const WgTcGeom g = wgrad_tc_geom(n, d0, d1, d2, c_in, c_out, taps);
  return (size_t)g.nchunks * g.units * 64 * c_out * sizeof(float);
}

int wgrad_tc_run(const void* x, int ld_x, const void* dy, int ld_dy, int n, int d0, int d1, int d2, int c_in, int c_out, int taps, float* dw,
                 float* workspace, cudaStream_t st) {
  const WgTcGeom g = wgrad_tc_geom(n, d0, d1, d2, c_in, c_out, taps);
  WgTcParams p;
  memset(&p, 0, sizeof(p));
  p.partial = workspace;
  p.taps = taps; p.KC = c_in / 64; p.units = g.units; p.nmb = g.nmb; p.NB = g.NB; p.N = g.N; p.c_in = c_in; p.c_out = c_out;
  p.bx = g.bx; p.by = g.by; p.bz = g.bz; p.bn = g.bn;
  p.tiles_x = g.tiles_x; p.tiles_y = g.tiles_y; p.tiles_z = g.tiles_z; p.tiles_n = g.tiles_n; p.m_tiles = g.m_tiles; p.nchunks = g.nchunks;
  p.idesc = make_idesc_bf16(128, g.N) | (1u << 15) | (1u << 16);   // A and B MN-major
  int rc = encode_volume_map(&p.x_map, x, c_in, d2, d1, d0, n, ld_x, (int64_t)d2 * ld_x, (int64_t)d1 * d2 * ld_x, (int64_t)d0 * d1 * d2 * ld_x, g.bx, g.by,
                             g.bz, g.bn);
  if (rc == DIQT_OK)
    rc = encode_volume_map(&p.dy_map, dy, c_out, d2, d1, d0, n, ld_dy, (int64_t)d2 * ld_dy, (int64_t)d1 * d2 * ld_dy, (int64_t)d0 * d1 * d2 * ld_dy, g.bx,
                           g.by, g.bz, g.bn);
  if (rc != DIQT_OK) return rc;
  const size_t smem = (size_t)WT_STAGES * (g.NB + 4) * WT_BOX + 128 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    DIQT_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done = true;
  }
  launch_pdl(wgrad_tc_kernel, dim3(g.nchunks, (g.nmb + 1) / 2, g.ntile_n), dim3(WT_THREADS), smem, st, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const int64_t count = (int64_t)taps * c_in * c_out;
  launch_pdl(wgrad_tc_reduce_kernel, dim3((unsigned)std::min<int64_t>((count + 255) / 256, 2368)), dim3(256), 0, st, (const float*)workspace, g.nchunks,
             taps, c_in / 64, c_in, c_out, dw);
  return check_launch("conv_wgrad_tc");
}

}  // namespace diqt
