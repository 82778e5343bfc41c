// tcgen05 implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// Replaces nn.Conv3d at imagen_pytorch3D.py:551-553 (3x3x3), :597 / :1388 (1x1x1), :495
// (pixel-unshuffle + 1x1x1) and :467 (+ Mish + PixelShuffle3D, :416-439) when channel counts are
// multiples of 64.
//
// GEMM view.  M tile = 128 output voxels forming a box (bx,by,bz,bn) of the channels-last volume,
// N tile = block_n <= 256 output channels, K = taps x c_in walked in K-blocks of 64 channels of
// one tap.  For every K-block
//   A [128 voxels x 64 ch] is ONE 5-D TMA box load of the input volume at the tap-shifted
//     coordinate; out-of-volume voxels are zero-filled by TMA, which IS the conv's zero padding;
//     the box lands in shared memory as 128 rows of 128 bytes in the 128B-swizzle K-major layout
//     tcgen05.mma wants;
//   B [block_n x 64] is one bulk copy of weights pre-swizzled at pack time.
// One elected thread issues 4 tcgen05.mma (K=16 each) per K-block into a TMEM accumulator;
// accumulators are double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
// Epilogue warps read TMEM (tcgen05.ld), add bias, (UP: Mish), round to bf16, stage the tile in
// swizzled shared memory and write it with TMA stores - for UP through per-sub-position tensor
// maps so that PixelShuffle3D is folded into the store coordinates; for DOWN the pixel-unshuffle
// is folded into per-tap load tensor maps.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-9 epilogue.  One warp issues an instruction
// every ~5 cycles here, so the epilogue of a wide tile (128 x 256 with Mish: 4 k instructions per thread) on four warps bounded the
// pixel-shuffle up-conv at 31 us for 40 MB of traffic (profiles/r5_init_conv.md has the same finding for init_conv): eight warps, the two of a
// lane quarter splitting the columns.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "tc_common.cuh"

namespace diqt {

// ------------------------------------------------------------------------------------------------
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;  // one A stage: 128 rows x 64 bf16
constexpr int kMaxMaps = 8;
constexpr int kEpiWarps = 8;              // two warps per TMEM lane quarter, each drains half of the accumulator's columns
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;

struct TcParams {
  CUtensorMap in_map[kMaxMaps];
  CUtensorMap out_map[kMaxMaps];
  const uint8_t* w;   // [n_tile][kblock][block_n][128 B], pre-swizzled
  const float* bias;  // GEMM column order, n_tiles * block_n entries
  int mode, taps, kchunks;
  int block_n, n_tiles;
  int up_c;  // UP: channels stored per sub-position
  int bx, by, bz, bn;
  int tiles_x, tiles_y, tiles_z, tiles_n, m_tiles;
  int stages;
  uint32_t idesc;
  uint32_t tmem_cols;
  // fused channel statistics of the stored output (for the next GroupNorm / the SE pool); NULL = off
  float* stats;          // [n][gridDim.x][c_out][2]
  StatsGroups sink;      // optional grouped reduction of the statistics rows (common.cuh)
  int n_batch;           // volumes
  int ox, oy, oz;        // output voxel grid (to mask rows of edge tiles)
  int edge_tiles;        // 1 if some tile sticks out of the volume
  // split-K (small volumes: fewer output tiles than SMs): the K-blocks of a tile are dealt to `ksplit` CTAs, each leaves its fp32
  // partial tile in `ws`; the LAST of them to finish (ticket) sums all partials in split order -- not arrival order: results stay
  // bitwise reproducible -- and runs the normal epilogue.  ksplit = 1: off.
  int ksplit, kbper;
  float* ws;                 // [tile][ksplit][128 rows][block_n] fp32
  unsigned int* tickets;     // [tile], zero before first use, self-resetting
};

__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up: [stages x (A | B)] [out staging block_n/64 x 16 KB] [bias] [barriers] [tmem ptr]
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.block_n * 128;
  const int stage_bytes = kABytes + b_bytes;
  uint8_t* stage_base = smem;
  uint8_t* out_stage = stage_base + (size_t)p.stages * stage_bytes;
  float* s_bias = reinterpret_cast<float*>(out_stage + (size_t)(p.block_n / 64) * kABytes);
  float* s_red = s_bias + p.n_tiles * p.block_n;  // [4][block_n][2] statistics scratch
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 4 * p.block_n * 2);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tfull_bar = bars + 2 * p.stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = p.taps * p.kchunks;
  const int total_work = p.m_tiles * p.n_tiles * p.ksplit;

  for (int i = threadIdx.x; i < p.n_tiles * p.block_n; i += blockDim.x) s_bias[i] = p.bias[i];
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // prologue above is private to this CTA; the previous kernel's activations are read only below

  auto decode_tile = [&](int work, int& n_tile, int& x0, int& y0, int& z0, int& b0) {
    work /= p.ksplit;  // the splits of one tile are neighbouring work units: they run at the same time on different CTAs
    n_tile = work % p.n_tiles;
    int m = work / p.n_tiles;
    const int tx = m % p.tiles_x; m /= p.tiles_x;
    const int ty = m % p.tiles_y; m /= p.tiles_y;
    const int tz = m % p.tiles_z; m /= p.tiles_z;
    x0 = tx * p.bx; y0 = ty * p.by; z0 = tz * p.bz; b0 = m * p.bn;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        int n_tile, x0, y0, z0, b0;
        decode_tile(work, n_tile, x0, y0, z0, b0);
        const uint8_t* wsrc = p.w + (size_t)n_tile * kblocks * b_bytes;
        const int kb0 = (work % p.ksplit) * p.kbper, kb1 = min(kblocks, kb0 + p.kbper);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int t = kb / p.kchunks, q = kb - t * p.kchunks;
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t bar = smem_u32(&full_bar[stage]);
          mbar_expect_tx(bar, (uint32_t)stage_bytes);
          uint8_t* a_dst = stage_base + (size_t)stage * stage_bytes;
          int dx = 0, dy = 0, dz = 0, mi = 0;
          if (p.mode == DIQT_CONV_K3) { dz = t / 9 - 1; dy = (t / 3) % 3 - 1; dx = t % 3 - 1; }
          else if (p.mode == DIQT_CONV_DOWN) { mi = t; }
          tma_load_5d(smem_u32(a_dst), &p.in_map[mi], bar, q * 64, x0 + dx, y0 + dy, z0 + dz, b0);
          bulk_load(smem_u32(a_dst + kABytes), wsrc + (size_t)kb * b_bytes, (uint32_t)b_bytes, bar);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
      const int kb0 = (work % p.ksplit) * p.kbper, kb1 = min(kblocks, kb0 + p.kbper);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        {  // warp-uniform issue code; the issuing lane is elected inside umma_bf16 / umma_commit
          const uint32_t a_addr = smem_u32(stage_base + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_sw128_desc(a_addr);
          const uint64_t bdesc = make_sw128_desc(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // +32 B (16 bf16) along K inside the swizzle atom = +2 in the address field
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc, ((kb - kb0) | k) != 0);
          umma_commit(smem_u32(&empty_bar[stage]));
          if (kb == kb1 - 1) umma_commit(smem_u32(&tfull_bar[acc]));
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quarter = warp & 3;           // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;    // tile row = voxel index inside the box
    const int et = threadIdx.x - 64;        // 0..255
    const int half = et >> 7;               // which half of the tile's 32-column chunks this warp drains
    const int nc32 = p.block_n / 32, c32_lo = half * (nc32 >> 1), c32_hi = c32_lo + (nc32 >> 1);
    int acc = 0;
    uint32_t acc_phase = 0;
    const int ngroups = p.block_n / 64;
    // fused statistics: thread -> (column pair cp of a 64-column group, row quarter rq); sums live in registers
    const int cp = et & 31, rq = (et >> 5) & 3;   // statistics: rows rq * 32 .. + 31 of the 64-column groups g with (g & 1) == half
    float st_s[4][2], st_q[4][2];
#pragma unroll
    for (int g = 0; g < 4; ++g) st_s[g][0] = st_s[g][1] = st_q[g][0] = st_q[g][1] = 0.f;
    int st_n = -1, st_first = -1;
    const uint32_t stage_a = smem_u32(out_stage);
    auto flush_stats = [&](int nvol) {
      // combine the four row quarters in a fixed order -> one partial per (volume, CTA): deterministic
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g < ngroups && (g & 1) == half) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            s_red[(rq * p.block_n + g * 64 + cp * 2 + h) * 2] = st_s[g][h];
            s_red[(rq * p.block_n + g * 64 + cp * 2 + h) * 2 + 1] = st_q[g][h];
            st_s[g][h] = st_q[g][h] = 0.f;
          }
        }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      float* dst = p.stats + ((size_t)nvol * gridDim.x + blockIdx.x) * p.block_n * 2;
      for (int col = et; col < p.block_n; col += kEpiThreads) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          a += s_red[(r4 * p.block_n + col) * 2];
          b += s_red[(r4 * p.block_n + col) * 2 + 1];
        }
        dst[col * 2] = a;
        dst[col * 2 + 1] = b;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    };
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      int n_tile, x0, y0, z0, b0;
      decode_tile(work, n_tile, x0, y0, z0, b0);
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      const float* wtile = nullptr;  // split-K: this tile's partials [ksplit][128][block_n]
      if (p.ksplit > 1) {
        const int tile = work / p.ksplit, ks = work % p.ksplit;
        // partial tile as [16-byte column chunk][row]: the 32 lanes of a warp (consecutive rows) touch consecutive 16-byte words
        // (the row-major layout of the first version made every access 32 separate sectors: 38 us instead of 21 at 16^3 x 128)
        uint4* mine = reinterpret_cast<uint4*>(p.ws + ((size_t)tile * p.ksplit + ks) * kTileM * p.block_n) + row;
        for (int c32 = c32_lo; c32 < c32_hi; ++c32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.block_n + c32 * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) __stcg(mine + (size_t)(c32 * 8 + j) * kTileM, make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));  // the accumulator is free: the next unit's MMAs may start
        __threadfence();  // the partial is visible device-wide before the ticket is taken
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        int* s_flag = reinterpret_cast<int*>(s_red);
        if (et == 0) {
          const unsigned t = atomicAdd(&p.tickets[tile], 1u);
          *s_flag = (t == (unsigned)p.ksplit - 1u);
          if (t == (unsigned)p.ksplit - 1u) p.tickets[tile] = 0u;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        const bool last = *s_flag != 0;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // s_red is reused by the statistics flush
        if (!last) {
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          continue;
        }
        __threadfence();
        wtile = p.ws + (size_t)tile * p.ksplit * kTileM * p.block_n + (size_t)row * 4;
      }
      if (p.stats) {
        if (st_n >= 0 && b0 != st_n) flush_stats(st_n);
        if (st_first < 0) st_first = b0;
        st_n = b0;
      }
      if (et == 0) bulk_wait_read0();  // previous tile's TMA stores have finished reading the staging tile
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      const uint32_t bias_a = smem_u32(s_bias + n_tile * p.block_n);
      for (int c32 = c32_lo; c32 < c32_hi; ++c32) {
        uint32_t r[32];
        if (wtile) {  // sum of the partials in split order
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
          for (int ks = 0; ks < p.ksplit; ++ks) {
            const uint4* src = reinterpret_cast<const uint4*>(wtile + (size_t)ks * kTileM * p.block_n) + (size_t)c32 * 8 * kTileM;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 v = __ldcg(src + (size_t)j * kTileM);
              r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + __uint_as_float(v.x));
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + __uint_as_float(v.y));
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + __uint_as_float(v.z));
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + __uint_as_float(v.w));
            }
          }
        } else {
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.block_n + c32 * 32), r);
          tmem_ld_wait();
        }
        // (bias, staging tile and statistics through 32-bit shared-window addresses, the bias in 16-byte loads: the generic pointers cost
        // 32 scalar loads per chunk and a 64-bit address computation per access)
        uint32_t packed[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const uint4 bq = lds_128(bias_a + (uint32_t)(c32 * 32 + j4 * 4) * 4);
          // two channels per FADD2 / FFMA2 / FMUL2 (per lane the same IEEE operations as the scalar code and as mish<true>)
          float2 v01 = __fadd2_rn(make_float2(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1])), make_float2(__uint_as_float(bq.x), __uint_as_float(bq.y)));
          float2 v23 = __fadd2_rn(make_float2(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])), make_float2(__uint_as_float(bq.z), __uint_as_float(bq.w)));
          if (p.mode == DIQT_CONV_UP) { v01 = mish2_fast(v01); v23 = mish2_fast(v23); }
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v01.x, v01.y), h1 = __floats2bfloat162_rn(v23.x, v23.y);
          packed[2 * j4] = *reinterpret_cast<uint32_t*>(&h0);
          packed[2 * j4 + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        // 64 B of this row -> four 16 B chunks of a 128 B swizzled staging row
        const uint32_t rowp = stage_a + (uint32_t)((c32 >> 1) * kABytes + row * 128);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = ((c32 & 1) * 4 + j) ^ (row & 7);
          sts_128(rowp + (uint32_t)(chunk * 16), packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && !wtile) mbar_arrive(smem_u32(&tempty_bar[acc]));  // (split-K released it before the ticket)
      fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (et == 0) {
        for (int g = 0; g < ngroups; ++g) {
          const int col = n_tile * p.block_n + g * 64;
          if (p.mode == DIQT_CONV_UP) {
            const int sub = col / p.up_c, c0 = col - sub * p.up_c;
            tma_store_5d(&p.out_map[sub], smem_u32(out_stage + (size_t)g * kABytes), c0, x0, y0, z0, b0);
          } else {
            tma_store_5d(&p.out_map[0], smem_u32(out_stage + (size_t)g * kABytes), col, x0, y0, z0, b0);
          }
        }
        bulk_commit();
      }
      if (p.stats) {
        // column sums of the staged (bf16-rounded) tile: a warp reads one 128 B row per step -> conflict free
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (g < ngroups && (g & 1) == half) {
            const uint32_t gb = stage_a + (uint32_t)(g * kABytes);
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
              const int rr = rq * 32 + r;
              if (p.edge_tiles) {
                const int lx = rr % p.bx, ly = (rr / p.bx) % p.by, lz = rr / (p.bx * p.by);
                if (x0 + lx >= p.ox || y0 + ly >= p.oy || z0 + lz >= p.oz) continue;
              }
              const uint32_t v = lds_u32(gb + (uint32_t)(rr * 128 + (((cp >> 2) ^ (rr & 7)) << 4) + ((cp & 3) << 2)));
              const float2 lh = make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
              const float2 ns = __fadd2_rn(make_float2(st_s[g][0], st_s[g][1]), lh);
              const float2 nq = __ffma2_rn(lh, lh, make_float2(st_q[g][0], st_q[g][1]));
              st_s[g][0] = ns.x; st_s[g][1] = ns.y; st_q[g][0] = nq.x; st_q[g][1] = nq.y;
            }
          }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.stats) {
      if (st_n >= 0) flush_stats(st_n);
      // volumes this CTA never touched still need a (zero) partial
      for (int nv = 0; nv < p.n_batch; ++nv) {
        if (st_first >= 0 && nv >= st_first && nv <= st_n) continue;
        float* dst = p.stats + ((size_t)nv * gridDim.x + blockIdx.x) * p.block_n * 2;
        for (int col = et; col < p.block_n * 2; col += kEpiThreads) dst[col] = 0.f;
      }
      stats_group_tail(p.sink, p.stats, p.n_batch, (int)gridDim.x, p.block_n, (int)blockIdx.x, 1, et, kEpiThreads, reinterpret_cast<int*>(s_red),
                       [] { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); });
    }
    if (et == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: (c_out, c_in*, k,k,k) fp32 -> [n_tile][kblock = tap*kchunks + q][block_n][64] bf16, 16-byte chunks
// XOR-swizzled by (row & 7) (what TMA SWIZZLE_128B would have produced).
// ------------------------------------------------------------------------------------------------
__global__ void conv_pack_tc_kernel(const float* __restrict__ w, int mode, int c_in, int c_out, int taps, int block_n,
                                    __nv_bfloat16* __restrict__ packed) {
  const int kchunks = c_in / 64, kblocks = taps * kchunks, n_tiles = c_out / block_n;
  const int64_t total = (int64_t)n_tiles * kblocks * block_n * 64;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % 64);
    const int r = (int)((i / 64) % block_n);
    const int kb = (int)((i / (64 * (int64_t)block_n)) % kblocks);
    const int nt = (int)(i / (64 * (int64_t)block_n * kblocks));
    const int t = kb / kchunks, q = kb - t * kchunks;
    const int ci = q * 64 + e;
    const int nn = nt * block_n + r;  // GEMM column
    float v;
    if (mode == DIQT_CONV_K3) v = w[((int64_t)nn * c_in + ci) * 27 + t];
    else if (mode == DIQT_CONV_K1) v = w[(int64_t)nn * c_in + ci];
    else if (mode == DIQT_CONV_DOWN) v = w[(int64_t)nn * (c_in * 8) + ci * 8 + t];
    else {
      const int C = c_out / 8, sub = nn / C, c = nn - sub * C;
      v = w[(int64_t)(c * 8 + sub) * c_in + ci];
    }
    const int chunk = (e >> 3) ^ (r & 7);
    const int64_t dst = (((int64_t)nt * kblocks + kb) * block_n + r) * 64 + chunk * 8 + (e & 7);
    packed[dst] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)ptr;
  return fn;
}

static int pow2floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

int encode_volume_map(CUtensorMap* map, const void* base, int c, int x, int y, int z, int n, int64_t sx, int64_t sy,
                             int64_t sz, int64_t sn, int bx, int by, int bz, int bn) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available (is a CUDA driver loaded?)");
    return DIQT_ECUDA;
  }
  cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)x, (cuuint64_t)y, (cuuint64_t)z, (cuuint64_t)n};
  cuuint64_t strides[4] = {(cuuint64_t)sx * 2, (cuuint64_t)sy * 2, (cuuint64_t)sz * 2, (cuuint64_t)sn * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)bn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims c=%d x=%d y=%d z=%d n=%d box %d,%d,%d,%d", (int)r, c, x, y, z, n, bx, by,
              bz, bn);
    return DIQT_ECUDA;
  }
  return DIQT_OK;
}

struct TcPlan {
  TcParams p;
  int grid;
  size_t smem;
};

int tc_taps(int mode) { return mode == DIQT_CONV_K3 ? 27 : mode == DIQT_CONV_DOWN ? 8 : 1; }

int tc_block_n(const diqt_conv_desc* d) {
  if (d->c_out % 256 == 0) return 256;
  if (d->c_out % 128 == 0) return 128;
  return 64;
}

bool conv_tc_supported(const diqt_conv_desc* d) {
  if (d->dtype != DIQT_BF16) return false;
  if (d->c_in % 64 != 0 || d->c_out % 64 != 0) return false;
  if (d->ld_in % 8 != 0 || d->ld_out % 8 != 0) return false;
  if (d->mode == DIQT_CONV_UP && (d->c_out / 8) % 64 != 0) return false;
  if (d->mode == DIQT_CONV_DOWN && (d->d0 % 2 || d->d1 % 2 || d->d2 % 2)) return false;
  return true;
}

size_t conv_tc_packed_bytes(const diqt_conv_desc* d) { return (size_t)tc_taps(d->mode) * d->c_in * d->c_out * 2; }

int conv_tc_pack(const diqt_conv_desc* d, const float* w, void* packed, cudaStream_t st) {
  conv_pack_tc_kernel<<<512, 256, 0, st>>>(w, d->mode, d->c_in, d->c_out, tc_taps(d->mode), tc_block_n(d), (__nv_bfloat16*)packed);
  return check_launch("conv_pack_tc");
}

int conv_tc_plan(const diqt_conv_desc* d, const void* in, void* out, const void* packed, const float* bias, TcPlan** out_plan) {
  DIQT_REQUIRE(conv_tc_supported(d), "conv(tc): unsupported shape c_in=%d c_out=%d ld_in=%d ld_out=%d mode=%d", d->c_in, d->c_out,
               d->ld_in, d->ld_out, d->mode);
  TcPlan* plan = new TcPlan();
  TcParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.mode = d->mode;
  p.taps = tc_taps(d->mode);
  p.kchunks = d->c_in / 64;
  p.block_n = tc_block_n(d);
  p.n_tiles = d->c_out / p.block_n;
  p.up_c = d->c_out / 8;
  p.w = (const uint8_t*)packed;
  p.bias = bias;
  // GEMM output voxel grid
  int od0 = d->d0, od1 = d->d1, od2 = d->d2;
  if (d->mode == DIQT_CONV_DOWN) { od0 /= 2; od1 /= 2; od2 /= 2; }
  int rem = kTileM;
  p.bx = std::min(8, pow2floor(od2)); rem /= p.bx;
  p.by = std::min(std::min(4, rem), pow2floor(od1)); rem /= p.by;
  p.bz = std::min(rem, pow2floor(od0)); rem /= p.bz;
  p.bn = rem;
  p.tiles_x = (od2 + p.bx - 1) / p.bx;
  p.tiles_y = (od1 + p.by - 1) / p.by;
  p.tiles_z = (od0 + p.bz - 1) / p.bz;
  p.tiles_n = (d->n + p.bn - 1) / p.bn;
  p.m_tiles = p.tiles_x * p.tiles_y * p.tiles_z * p.tiles_n;
  p.idesc = make_idesc_bf16(kTileM, p.block_n);
  p.stats = nullptr;
  p.n_batch = d->n;
  p.ox = od2; p.oy = od1; p.oz = od0;
  p.edge_tiles = (od2 % p.bx || od1 % p.by || od0 % p.bz) ? 1 : 0;
  p.tmem_cols = 2 * p.block_n < 32 ? 32 : 2 * p.block_n;  // 128 / 256 / 512: powers of two
  p.ksplit = 1;
  p.kbper = p.taps * p.kchunks;

  const __nv_bfloat16* inb = (const __nv_bfloat16*)in;
  __nv_bfloat16* outb = (__nv_bfloat16*)out;
  const int64_t ld = d->ld_in;
  int rc = DIQT_OK;
  if (d->mode == DIQT_CONV_DOWN) {
    // tap (s1,s2,s3): input voxel (2z+s1, 2y+s2, 2x+s3)  -> a view with doubled strides and an offset base
    for (int t = 0; t < 8 && rc == DIQT_OK; ++t) {
      const int s1 = (t >> 2) & 1, s2 = (t >> 1) & 1, s3 = t & 1;
      const __nv_bfloat16* base = inb + (((int64_t)s1 * d->d1 + s2) * d->d2 + s3) * ld;
      rc = encode_volume_map(&p.in_map[t], base, d->c_in, od2, od1, od0, d->n, 2 * ld, 2 * (int64_t)d->d2 * ld,
                             2 * (int64_t)d->d1 * d->d2 * ld, (int64_t)d->d0 * d->d1 * d->d2 * ld, p.bx, p.by, p.bz, p.bn);
    }
  } else {
    rc = encode_volume_map(&p.in_map[0], inb, d->c_in, d->d2, d->d1, d->d0, d->n, ld, (int64_t)d->d2 * ld,
                           (int64_t)d->d1 * d->d2 * ld, (int64_t)d->d0 * d->d1 * d->d2 * ld, p.bx, p.by, p.bz, p.bn);
  }
  const int64_t lo = d->ld_out;
  if (rc == DIQT_OK) {
    if (d->mode == DIQT_CONV_UP) {
      const int D0 = 2 * d->d0, D1 = 2 * d->d1, D2 = 2 * d->d2;
      for (int s = 0; s < 8 && rc == DIQT_OK; ++s) {
        const int i = (s >> 2) & 1, j = (s >> 1) & 1, k = s & 1;
        __nv_bfloat16* base = outb + (((int64_t)i * D1 + j) * D2 + k) * lo;
        rc = encode_volume_map(&p.out_map[s], base, p.up_c, d->d2, d->d1, d->d0, d->n, 2 * lo, 2 * (int64_t)D2 * lo,
                               2 * (int64_t)D1 * D2 * lo, (int64_t)D0 * D1 * D2 * lo, p.bx, p.by, p.bz, p.bn);
      }
    } else {
      rc = encode_volume_map(&p.out_map[0], outb, d->c_out, od2, od1, od0, d->n, lo, (int64_t)od2 * lo, (int64_t)od1 * od2 * lo,
                             (int64_t)od0 * od1 * od2 * lo, p.bx, p.by, p.bz, p.bn);
    }
  }
  if (rc != DIQT_OK) {
    delete plan;
    return rc;
  }
  // shared memory budget
  const int stage_bytes = kABytes + p.block_n * 128;
  const size_t fixed = (size_t)(p.block_n / 64) * kABytes + (size_t)p.n_tiles * p.block_n * 4 + (size_t)p.block_n * 32 + 256 + 1024;
  int stages = (int)((220 * 1024 - fixed) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) {
    set_error("conv(tc): not enough shared memory for block_n=%d", p.block_n);
    delete plan;
    return DIQT_EINVAL;
  }
  p.stages = stages;
  plan->smem = fixed + (size_t)stages * stage_bytes;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int total = p.m_tiles * p.n_tiles;
  plan->grid = total < sms ? total : sms;
  static bool attr_done = false;
  if (!attr_done) {
    DIQT_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  *out_plan = plan;
  return DIQT_OK;
}

int conv_tc_run(const TcPlan* plan, cudaStream_t st) {
  launch_pdl(conv_tc_kernel, plan->grid, kThreads, plan->smem, st, plan->p);
  return check_launch("conv_tc");
}

void conv_tc_destroy(TcPlan* plan) { delete plan; }

// ---- split-K for small volumes ---------------------------------------------------------------------------------------------
static int tc_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
// splits that bring the number of CTAs close to the SM count (0 / 1: not worth it)
static int tc_ksplit_for(const TcParams& p, int* kbper_out) {
  const int kblocks = p.taps * p.kchunks, tiles = p.m_tiles * p.n_tiles, sms = tc_sms();
  if (p.mode != DIQT_CONV_K3 || tiles * 2 > sms || kblocks < 8) return 1;
  int want = sms / tiles;
  if (want > 8) want = 8;
  const int kbper = (kblocks + want - 1) / want;
  if (kbper_out) *kbper_out = kbper;
  return (kblocks + kbper - 1) / kbper;
}
bool conv_tc_prefers_small(const diqt_conv_desc* d) {  // fewer output tiles than half the SMs: the per-tap kernel with split-K wins over the z-march
  static int env = -1;   // DIQT_TC_SMALL=0: A/B of the z-march kernel (with its fused GroupNorm) on small volumes
  if (env < 0) { const char* e = getenv("DIQT_TC_SMALL"); env = (e && e[0] == '0') ? 0 : 1; }
  if (!env) return false;
  if (d->mode != DIQT_CONV_K3 || !conv_tc_supported(d)) return false;
  const int64_t rows = (int64_t)d->n * d->d0 * d->d1 * d->d2;
  const int64_t tiles = (rows + kTileM - 1) / kTileM * (d->c_out / tc_block_n(d));
  return tiles * 2 <= tc_sms();
}
size_t conv_tc_workspace_bytes(const TcPlan* plan) {
  int kbper = 0;
  const int ks = tc_ksplit_for(plan->p, &kbper);
  if (ks <= 1) return 0;
  const size_t tiles = (size_t)plan->p.m_tiles * plan->p.n_tiles;
  return ((tiles * 4 + 255) / 256) * 256 + tiles * ks * kTileM * plan->p.block_n * sizeof(float);
}
// workspace: [tickets (zeroed by the caller once; self-resetting)] [partials]; must be called before conv_tc_set_stats (the grid changes)
int conv_tc_set_workspace(TcPlan* plan, void* ws, size_t bytes) {
  const size_t need = conv_tc_workspace_bytes(plan);
  DIQT_REQUIRE(need > 0, "conv(tc): this plan does not use a split-K workspace");
  DIQT_REQUIRE(ws && bytes >= need && (uintptr_t)ws % 256 == 0, "conv(tc): split-K workspace of %zu bytes (256-byte aligned) needed, got %zu", need, bytes);
  DIQT_REQUIRE(!plan->p.stats, "conv(tc): set the workspace before the statistics sink");
  TcParams& p = plan->p;
  p.ksplit = tc_ksplit_for(p, &p.kbper);
  const size_t tiles = (size_t)p.m_tiles * p.n_tiles;
  p.tickets = (unsigned int*)ws;
  p.ws = (float*)((uint8_t*)ws + ((tiles * 4 + 255) / 256) * 256);
  const int total = (int)tiles * p.ksplit, sms = tc_sms();
  plan->grid = total < sms ? total : sms;
  return DIQT_OK;
}

// Fused output statistics are available when one CTA tile never mixes volumes and N fits one tile.
int conv_tc_set_stats(TcPlan* plan, float* partial, float* group, unsigned int* tickets, int* ngroups) {
  TcParams& p = plan->p;
  if (ngroups) *ngroups = 0;
  if (p.n_tiles != 1 || p.bn != 1 || p.mode == DIQT_CONV_UP) return 0;
  p.stats = partial;
  StatsGroups& g = p.sink;
  g.group = group;
  g.tickets = tickets;
  g.gsize = stats_group_size(plan->grid, 1);
  g.ngroups = (plan->grid + g.gsize - 1) / g.gsize;
  if (ngroups) *ngroups = group ? g.ngroups : 0;
  return plan->grid;
}

}  // namespace diqt
