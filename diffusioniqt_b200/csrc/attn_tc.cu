// Softmax attention on the 5th-gen tensor cores (SoftMaxAttention.forward imagen_pytorch3D.py:1087-1100, MultiHeadAttention.forward
// :826-836): out = act( softmax_k(q k^T * scale) v ) per head, dim_head = 64, bf16 operands, fp32 accumulation in TMEM, nothing N x N in
// memory.
//
// One CTA = 128 queries of one head.  Two passes over the key tiles (128 keys each), so the output accumulator never has to be rescaled:
//   pass 1:  S = Q K_j^T (tcgen05.mma M128 N128 K64 -> TMEM) ; the softmax warps only take the row maximum m (no exponentials)
//   pass 2:  S = Q K_j^T again ; P = exp(scale (s - m)) as bf16 into a 128B-swizzled shared tile, l += P ; O += P V_j (M128 N64 K128
//            -> TMEM) ; the epilogue scales O by 1 / l
// Q K^T is computed twice (tensor time is cheap) in exchange for no tcgen05.ld / st round trips of O between tiles.  Two CTAs fit an SM
// (100 KB of shared memory, 256 TMEM columns each), so one CTA's softmax overlaps the other's MMA / TMA latency.
// Operands are all K-major SWIZZLE_128B tiles written by TMA: Q and K straight from the [tokens][channels] q | k | v buffer, V from a
// transposed copy V^T [channels][tokens] (written by attn_transpose_kernel; its padding columns must be zero).
// Warp roles (192 threads): 0 TMA producer, 1 MMA issuer (owns TMEM), 2-5 softmax / epilogue (one query row per thread).
#include <stdlib.h>
#include <string.h>

#include <new>

#include "tc_common.cuh"

namespace diqt {

constexpr int AT_Q = 128, AT_K = 128, AT_D = 64;        // queries per CTA, keys per tile, head dim
constexpr int AT_TILE = AT_Q * 128;                     // one [128 rows][64 bf16] tile: 16 KB
constexpr int AT_VCHUNK = 64 * 128;                     // one [64 d rows][64 keys] tile of V^T: 8 KB
constexpr int AT_THREADS = 192;

struct AttnTcParams {
  CUtensorMap q_map, k_map, vt_map, v_map;   // vt_map: transposed copy (two-pass kernel); v_map: V as it lies (single-pass kernel)
  __nv_bfloat16* out;
  int ld_out, ntok, heads, q_col0, k_col0, act;
  float c;  // scale * log2(e)
  uint32_t idesc_s, idesc_o, idesc_o_mn;   // idesc_o_mn: B operand MN-major
};

__device__ __forceinline__ float ex2_approx(float x) {  // arguments are <= 0 here
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes (Cody-Waite split + cubic): the softmax is bound by the 16 / clk / SM special-function unit, so a share of
// the exponentials is computed here instead (the trick FlashAttention-4 uses on this architecture).  x <= ~8; relative error < 7e-4,
// far below the bf16 rounding of P; x = -inf (masked keys) gives 2^-126 instead of 0, which no sum notices.
__device__ __forceinline__ float ex2_poly(float x) {
  const float xr = fmaxf(x, -126.f);
  const float t = xr + 12582912.f;           // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = xr - (t - 12582912.f);     // [-0.5, 0.5]
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
#ifndef DIQT_ATTN_POLY
#define DIQT_ATTN_POLY 0   // 1: every fourth exponential of the single-pass kernel goes through ex2_poly; 0: all through MUFU (measured: 0.580 ms
                           // against 0.594 with the polynomial share at 13 824 tokens: the kernel is not MUFU-saturated, the extra issue slots cost more)
#endif

__device__ __forceinline__ void tma_load_2d_as5(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int row0) {
  tma_load_5d(dst, map, bar, c0, row0, 0, 0, 0);
}

__global__ void __launch_bounds__(AT_THREADS, 2) softmax_attn_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;                       // 16 KB
  uint8_t* k_s = q_s + AT_TILE;              // 2 stages x 16 KB
  uint8_t* v_s = k_s + 2 * AT_TILE;          // 2 key chunks x 8 KB (one stage: V_j is only needed once P_j exists)
  uint8_t* p_s = v_s + 2 * AT_VCHUNK;        // 2 key chunks x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 2 * AT_TILE);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // 2
  uint64_t* k_empty = bars + 3;       // 2
  uint64_t* v_full = bars + 5;        // 2
  uint64_t* v_empty = bars + 7;       // 2
  uint64_t* s_full = bars + 9;        // 1
  uint64_t* s_free = bars + 10;       // 1 (4 arrivals)
  uint64_t* p_full = bars + 11;       // 1 (4 arrivals)
  uint64_t* p_free = bars + 12;       // 1
  uint64_t* o_full = bars + 13;       // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = blockIdx.x, head = blockIdx.y;
  const int nkt = (p.ntok + AT_K - 1) / AT_K;

  if (warp == 0 && lane == 0) {
    mbar_init(smem_u32(q_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&k_full[s]), 1);
      mbar_init(smem_u32(&k_empty[s]), 1);
      mbar_init(smem_u32(&v_full[s]), 1);
      mbar_init(smem_u32(&v_empty[s]), 1);
    }
    mbar_init(smem_u32(s_full), 1);
    mbar_init(smem_u32(s_free), 4);
    mbar_init(smem_u32(p_full), 4);
    mbar_init(smem_u32(p_free), 1);
    mbar_init(smem_u32(o_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;          // columns [0, 128): S
  const uint32_t tmem_o = tmem_base + 128;    // columns [128, 192): O

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(smem_u32(q_full), AT_TILE);
      tma_load_2d_as5(smem_u32(q_s), &p.q_map, smem_u32(q_full), p.q_col0 + head * AT_D, qtile * AT_Q);
      int ks = 0;
      const int vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < nkt; ++j) {
          mbar_wait(smem_u32(&k_empty[ks]), kph ^ 1);
          mbar_expect_tx(smem_u32(&k_full[ks]), AT_TILE);
          tma_load_2d_as5(smem_u32(k_s + ks * AT_TILE), &p.k_map, smem_u32(&k_full[ks]), p.k_col0 + head * AT_D, j * AT_K);
          if (++ks == 2) { ks = 0; kph ^= 1; }
          if (pass == 1) {
            mbar_wait(smem_u32(&v_empty[vs]), vph ^ 1);
            mbar_expect_tx(smem_u32(&v_full[vs]), 2 * AT_VCHUNK);
            uint8_t* dst = v_s + vs * 2 * AT_VCHUNK;
            tma_load_2d_as5(smem_u32(dst), &p.vt_map, smem_u32(&v_full[vs]), j * AT_K, head * AT_D);
            tma_load_2d_as5(smem_u32(dst + AT_VCHUNK), &p.vt_map, smem_u32(&v_full[vs]), j * AT_K + 64, head * AT_D);
            vph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform; the issuing lane is elected inside umma_bf16 / umma_commit) =====================
    mbar_wait(smem_u32(q_full), 0);
    tc_fence_after();
    const uint64_t qdesc = make_sw128_desc(smem_u32(q_s));
    int ks = 0, it = 0, pit = 0;
    const int vs = 0;
    uint32_t kph = 0, vph = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int j = 0; j < nkt; ++j) {
        mbar_wait(smem_u32(&k_full[ks]), kph);
        mbar_wait(smem_u32(s_free), (uint32_t)((it & 1) ^ 1));  // the softmax warps have read the previous S
        tc_fence_after();
        {
          const uint64_t kdesc = make_sw128_desc(smem_u32(k_s + ks * AT_TILE));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), p.idesc_s, k != 0);
          umma_commit(smem_u32(&k_empty[ks]));
          umma_commit(smem_u32(s_full));
        }
        if (++ks == 2) { ks = 0; kph ^= 1; }
        ++it;
        if (pass == 1) {
          mbar_wait(smem_u32(&v_full[vs]), vph);
          mbar_wait(smem_u32(p_full), (uint32_t)(pit & 1));
          tc_fence_after();
          {
            const uint32_t vb = smem_u32(v_s + vs * 2 * AT_VCHUNK);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const uint64_t pdesc = make_sw128_desc(smem_u32(p_s + c * AT_TILE));
              const uint64_t vdesc = make_sw128_desc(vb + c * AT_VCHUNK);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_o, pdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), p.idesc_o, (j | c | k) != 0);
            }
            umma_commit(smem_u32(&v_empty[vs]));
            umma_commit(smem_u32(p_free));
          }
          vph ^= 1;
          ++pit;
        }
      }
    }
    umma_commit(smem_u32(o_full));
  } else {
    // ===================== softmax / epilogue (warps 2..5): thread = one query row =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    float m = -INFINITY, l = 0.f;  // row maximum of the raw scores, sum of exp(scale (s - m))
    int it = 0;
    // ---- pass 1: row maximum
    for (int j = 0; j < nkt; ++j, ++it) {
      mbar_wait(smem_u32(s_full), (uint32_t)(it & 1));
      tc_fence_after();
#pragma unroll 1
      for (int c32 = 0; c32 < 4; ++c32) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_addr + (uint32_t)(c32 * 32), r);
        tmem_ld_wait();
        const int key0 = j * AT_K + c32 * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (key0 + i < p.ntok) m = fmaxf(m, __uint_as_float(r[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(s_free));
    }
    const float mc = m * p.c;  // key 0 always exists: m is finite
    // ---- pass 2: probabilities -> shared memory (A operand of P V)
    int pit = 0;
    for (int j = 0; j < nkt; ++j, ++it, ++pit) {
      mbar_wait(smem_u32(s_full), (uint32_t)(it & 1));
      mbar_wait(smem_u32(p_free), (uint32_t)((pit & 1) ^ 1));  // the previous P V has finished reading the P tile
      tc_fence_after();
#pragma unroll 1
      for (int c32 = 0; c32 < 4; ++c32) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_addr + (uint32_t)(c32 * 32), r);
        tmem_ld_wait();
        const int key0 = j * AT_K + c32 * 32;
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = key0 + 2 * i < p.ntok ? ex2_approx(fmaf(__uint_as_float(r[2 * i]), p.c, -mc)) : 0.f;
          const float p1 = key0 + 2 * i + 1 < p.ntok ? ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), p.c, -mc)) : 0.f;
          l += p0 + p1;
          __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
          packed[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        // 32 keys = 64 B of this row: four 16-byte units of the 128-byte swizzled row of key chunk (c32 >> 1)
        uint8_t* rowp = p_s + (size_t)(c32 >> 1) * AT_TILE + (size_t)row * 128;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int unit = ((c32 & 1) * 4 + u) ^ (row & 7);
          *reinterpret_cast<uint4*>(rowp + unit * 16) = make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async();  // generic-proxy writes of P become visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(s_free));
        mbar_arrive(smem_u32(p_full));
      }
    }
    // ---- epilogue: O / l -> activation -> bf16 rows
    const float inv_l = 1.f / l;
    mbar_wait(smem_u32(o_full), 0);
    tc_fence_after();
    const int q = qtile * AT_Q + row;
#pragma unroll 1
    for (int c32 = 0; c32 < 2; ++c32) {
      uint32_t r[32];
      tmem_ld32(tmem_o + lane_addr + (uint32_t)(c32 * 32), r);
      tmem_ld_wait();
      if (q < p.ntok) {
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float v0 = __uint_as_float(r[2 * i]) * inv_l, v1 = __uint_as_float(r[2 * i + 1]) * inv_l;
          if (p.act == 1) { v0 = mish<false>(v0); v1 = mish<false>(v1); }
          __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
          packed[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)q * p.ld_out + head * AT_D + c32 * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[u] = make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------------------------
// Version 2: ONE pass over the keys (online softmax), NQ query tiles of 128 rows per CTA in ping-pong, probabilities handed to the
// tensor core through TENSOR MEMORY (tcgen05.mma with the A operand in TMEM), never through shared memory.
//
// Why (profiles/r1s_ncu_softmax_attn_tc.csv, profiles/sweep_r2a.jsonl): the two-pass kernel above ran the chain Q K^T -> tcgen05.ld ->
// exp -> st.shared -> P V serially per CTA (tensor pipe 14 % active) and computed Q K^T twice; cuDNN's fused attention on the same GPU
// was 1.9x (1728 tokens) and 2.7x (13 824 tokens) faster.  Per (query tile, key tile) the tensor pipe needs 256 (Q K^T, N = 128) + 512
// (P V: N = 64 runs at the 64-cycle-per-instruction floor) cycles, the 128 x 128 exponentials need 1024 MUFU cycles: the kernel is bound
// by the SFU, so everything else has to overlap with it:
//   * two query tiles per CTA, each with its own softmax warpgroup (thread = one query row, no shuffles), S / P / O of both in TMEM
//     (2 x 128 + 2 x 64 + 2 x 64 = 512 columns): while one warpgroup exponentiates, the tensor pipe fills the other's S;
//   * the issuer queues Q_i K_{j+1}^T as soon as warpgroup i has READ S_i(j) into registers (s_free), then P_i(j) V_j when P_i(j) exists;
//   * online softmax with LAZY rescaling: the running reference maximum only moves when a row's new maximum exceeds it by more than
//     2^8 (probabilities up to 256 are harmless in bf16 / fp32), so O is rescaled in TMEM (tcgen05.ld / st by the warpgroup that owns
//     the rows) a handful of times per row instead of once per key tile; the final O / l is exact either way.
// Warp roles (128 NQ + 64 threads): warps 0 .. 4 NQ - 1 softmax / epilogue, then the TMA producer, then the MMA issuer (TMEM owner);
// the single-thread roles have the highest warp ids (scheduler priority, see conv_zm.cu).
#ifndef DIQT_ATTN_LOADS
#define DIQT_ATTN_LOADS 3   // 1: whole row in registers, then maximum, then exponentials; 2: two TMEM reads (measured 1.6x slower); 3: speculative maximum
#endif
constexpr int AT2_KSTAGES = 3, AT2_VSTAGES = 2;
constexpr float AT2_TAU = 8.f;  // log2 units

template <int NQ>
__global__ void __launch_bounds__(128 * NQ + 64, 1) softmax_attn_tc2_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;                                   // NQ x 16 KB
  uint8_t* k_s = q_s + NQ * AT_TILE;                     // AT2_KSTAGES x 16 KB
  uint8_t* v_s = k_s + AT2_KSTAGES * AT_TILE;            // AT2_VSTAGES x (2 key chunks x 8 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + AT2_VSTAGES * 2 * AT_VCHUNK);
  uint64_t* q_full = bars;                               // 1
  uint64_t* k_full = q_full + 1;                         // [AT2_KSTAGES]
  uint64_t* k_empty = k_full + AT2_KSTAGES;
  uint64_t* v_full = k_empty + AT2_KSTAGES;              // [AT2_VSTAGES]
  uint64_t* v_empty = v_full + AT2_VSTAGES;
  uint64_t* s_full = v_empty + AT2_VSTAGES;              // [NQ]  Q_i K_j^T complete
  uint64_t* s_free = s_full + 2;                         // [NQ]  S_i read into registers (4 warp arrivals)
  uint64_t* p_full = s_free + 2;                         // [NQ]  P_i(j) stored in TMEM, O_i rescaled if needed (4 warp arrivals)
  uint64_t* pv_done = p_full + 2;                        // [NQ]  P_i(j) V_j complete: P_i may be overwritten, O_i is up to date
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = 4 * NQ, W_MMA = 4 * NQ + 1;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * (AT_Q * NQ);
  const int nkt = (p.ntok + AT_K - 1) / AT_K;

  if (warp == W_TMA && lane == 0) {
    mbar_init(smem_u32(q_full), 1);
    for (int s = 0; s < AT2_KSTAGES; ++s) { mbar_init(smem_u32(&k_full[s]), 1); mbar_init(smem_u32(&k_empty[s]), 1); }
    for (int s = 0; s < AT2_VSTAGES; ++s) { mbar_init(smem_u32(&v_full[s]), 1); mbar_init(smem_u32(&v_empty[s]), 1); }
    for (int i = 0; i < NQ; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_free[i]), 4);
      mbar_init(smem_u32(&p_full[i]), 4);
      mbar_init(smem_u32(&pv_done[i]), 1);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // columns: S_i at 128 i, O_i at 256 + 64 i, P_i at 384 + 64 i
  if (warp == W_TMA) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(smem_u32(q_full), NQ * AT_TILE);
      for (int i = 0; i < NQ; ++i) tma_load_2d_as5(smem_u32(q_s + i * AT_TILE), &p.q_map, smem_u32(q_full), p.q_col0 + head * AT_D, q0 + i * AT_Q);
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < nkt; ++j) {
        mbar_wait(smem_u32(&k_empty[ks]), kph ^ 1);
        mbar_expect_tx(smem_u32(&k_full[ks]), AT_TILE);
        tma_load_2d_as5(smem_u32(k_s + ks * AT_TILE), &p.k_map, smem_u32(&k_full[ks]), p.k_col0 + head * AT_D, j * AT_K);
        if (++ks == AT2_KSTAGES) { ks = 0; kph ^= 1; }
        mbar_wait(smem_u32(&v_empty[vs]), vph ^ 1);
        mbar_expect_tx(smem_u32(&v_full[vs]), 2 * AT_VCHUNK);
        // V_j as it lies in memory: a [128 keys][64 d] box IS the MN-major SWIZZLE_128B B operand (N = d contiguous, eight-key groups
        // 1024 B apart; see linattn_tc.cu), so the single-pass kernel needs no transposed copy; keys past the end arrive as zeros
        tma_load_2d_as5(smem_u32(v_s + vs * 2 * AT_VCHUNK), &p.v_map, smem_u32(&v_full[vs]), head * AT_D, j * AT_K);
        if (++vs == AT2_VSTAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (warp-uniform code; the issuing lane is elected inside the wrappers) =====================
    mbar_wait(smem_u32(q_full), 0);
    tc_fence_after();
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    auto issue_qk = [&](int i, int kstage) {
      const uint64_t qdesc = make_sw128_desc(smem_u32(q_s + i * AT_TILE));
      const uint64_t kdesc = make_sw128_desc(smem_u32(k_s + kstage * AT_TILE));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + (uint32_t)(i * 128), qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), p.idesc_s, k != 0);
      umma_commit(smem_u32(&s_full[i]));
    };
    // S_i(0)
    mbar_wait(smem_u32(&k_full[0]), 0);
    tc_fence_after();
    for (int i = 0; i < NQ; ++i) issue_qk(i, 0);
    umma_commit(smem_u32(&k_empty[0]));
    ks = 1 % AT2_KSTAGES;
    if (ks == 0) kph ^= 1;
    for (int j = 0; j < nkt; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      if (j + 1 < nkt) {  // S_i(j + 1) as soon as warpgroup i holds S_i(j) in registers
        mbar_wait(smem_u32(&k_full[ks]), kph);
        for (int i = 0; i < NQ; ++i) {
          mbar_wait(smem_u32(&s_free[i]), ph);
          tc_fence_after();
          issue_qk(i, ks);
        }
        umma_commit(smem_u32(&k_empty[ks]));
        if (++ks == AT2_KSTAGES) { ks = 0; kph ^= 1; }
      }
      mbar_wait(smem_u32(&v_full[vs]), vph);
      const uint32_t vb = smem_u32(v_s + vs * 2 * AT_VCHUNK);
      for (int i = 0; i < NQ; ++i) {
        mbar_wait(smem_u32(&p_full[i]), ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // K = 16 keys per instruction: 8 TMEM columns of P, 32 bytes along the V^T rows
          const uint64_t vdesc = make_sw128_desc_sbo(vb + kk * 2048, 1024);   // 16 keys = two 8-key groups 1024 B apart
          umma_bf16_ts(tmem_base + (uint32_t)(256 + i * 64), tmem_base + (uint32_t)(384 + i * 64 + kk * 8), vdesc, p.idesc_o_mn, (j | kk) != 0);
        }
        umma_commit(smem_u32(&pv_done[i]));
      }
      umma_commit(smem_u32(&v_empty[vs]));
      if (++vs == AT2_VSTAGES) { vs = 0; vph ^= 1; }
    }
  } else {
    // ===================== softmax / epilogue: warpgroup i = query tile i, thread = one query row =====================
    const int i = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr + (uint32_t)(i * 128);
    const uint32_t t_o = tmem_base + lane_addr + (uint32_t)(256 + i * 64);
    const uint32_t t_p = tmem_base + lane_addr + (uint32_t)(384 + i * 64);
    float m_used = -INFINITY, l = 0.f;  // reference maximum (log2 units, already scaled) and sum of 2^(s c - m_used)
    for (int j = 0; j < nkt; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      mbar_wait(smem_u32(&s_full[i]), ph);
      tc_fence_after();
      const int nvalid = p.ntok - j * AT_K;  // keys of this tile that exist (the TMA zero-fills the rest)
#if DIQT_ATTN_LOADS == 3
      // Speculative reference maximum: tensor memory delivers S at ~64 B/clk per SM (measured: a second read of S costs as much as all
      // the exponentials), so the loads must overlap the MUFU work of the SAME warp, which the row maximum normally forbids (it needs the
      // whole row first).  With lazy rescaling the reference maximum of a row rarely moves: exponentiate chunk c against the CURRENT
      // reference while chunk c + 1 is in flight, track the row maximum on the side, and only if some row of the warp outgrew its
      // reference by more than 2^8 redo the tile for the warp (first tile always; afterwards a handful of tiles per row at most).
      uint32_t pk[64];
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      auto exp_chunk = [&](const uint32_t (&r)[32], int c, const bool mask) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int k0 = c * 32 + 2 * k;
          const float v0 = (!mask || k0 < nvalid) ? __uint_as_float(r[2 * k]) : -INFINITY;
          const float v1 = (!mask || k0 + 1 < nvalid) ? __uint_as_float(r[2 * k + 1]) : -INFINITY;
          mx4[k & 3] = fmaxf(mx4[k & 3], fmaxf(v0, v1));
          const float p0 = ex2_approx(fmaf(v0, p.c, -m_used));
          const float p1 = (DIQT_ATTN_POLY && (k & 1)) ? ex2_poly(fmaf(v1, p.c, -m_used)) : ex2_approx(fmaf(v1, p.c, -m_used));
          sum4[k & 3] += p0 + p1;
          __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
          pk[c * 16 + k] = *reinterpret_cast<uint32_t*>(&h);
        }
      };
      auto max_chunk = [&](const uint32_t (&r)[32], int c) {
#pragma unroll
        for (int k = 0; k < 32; ++k) mx4[k & 3] = fmaxf(mx4[k & 3], c * 32 + k < nvalid ? __uint_as_float(r[k]) : -INFINITY);
      };
      auto exp_tile = [&](const bool mask) {  // four chunks, the load of chunk c + 1 in flight under the exponentials of chunk c
        uint32_t ra[32], rb[32];
        tmem_ld32(t_s, ra);
        tmem_ld_wait();
        tmem_ld32(t_s + 32, rb); exp_chunk(ra, 0, mask); tmem_ld_wait();
        tmem_ld32(t_s + 64, ra); exp_chunk(rb, 1, mask); tmem_ld_wait();
        tmem_ld32(t_s + 96, rb); exp_chunk(ra, 2, mask); tmem_ld_wait();
        exp_chunk(rb, 3, mask);
      };
      const bool full = nvalid >= AT_K;  // every tile but possibly the last
      if (j > 0) {
        if (full) exp_tile(false); else exp_tile(true);
      } else {  // no reference yet: only the maximum
        uint32_t ra[32], rb[32];
        tmem_ld32(t_s, ra);
        tmem_ld_wait();
        tmem_ld32(t_s + 32, rb); max_chunk(ra, 0); tmem_ld_wait();
        tmem_ld32(t_s + 64, ra); max_chunk(rb, 1); tmem_ld_wait();
        tmem_ld32(t_s + 96, rb); max_chunk(ra, 2); tmem_ld_wait();
        max_chunk(rb, 3);
      }
      const float mnew = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.c;
      float alpha = 1.f;
      const bool grow = mnew > m_used + AT2_TAU;
      const bool redo = __any_sync(0xffffffffu, grow);
      if (redo) {
        if (grow) {
          alpha = ex2_approx(m_used - mnew);  // 0 on the first tile (m_used = -inf)
          m_used = mnew;
          l *= alpha;
        }
        sum4[0] = sum4[1] = sum4[2] = sum4[3] = 0.f;
        exp_tile(true);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_free[i]));
      l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      if (j > 0) {
        mbar_wait(smem_u32(&pv_done[i]), ph ^ 1);  // P_i(j - 1) V_{j-1} complete: P_i free, O_i consistent
        tc_fence_after();
        if (redo) {  // rescale this warp's 32 rows of O_i (alpha = 1 for the rows that did not move)
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(t_o + (uint32_t)(c * 32), o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tmem_st32(t_o + (uint32_t)(c * 32), o);
          }
        }
      }
      tmem_st32(t_p, *reinterpret_cast<uint32_t(*)[32]>(pk));
      tmem_st32(t_p + 32, *reinterpret_cast<uint32_t(*)[32]>(pk + 32));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_full[i]));
    }
#elif DIQT_ATTN_LOADS == 2
      // Two reads of S from tensor memory (64 columns at a time): the first for the row maximum, the second to exponentiate.  A 320-thread
      // block is allocated registers for 12 warps (168 per thread); holding all 128 scores of a row plus the packed probabilities does
      // not fit without spills, while TMEM reads are cheap (16 KB per warp and pass at 64 B/clk, hidden behind the other warpgroup's MUFU).
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t r[64];
        tmem_ld32(t_s + (uint32_t)(half * 64), *reinterpret_cast<uint32_t(*)[32]>(r));
        tmem_ld32(t_s + (uint32_t)(half * 64 + 32), *reinterpret_cast<uint32_t(*)[32]>(r + 32));
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          const float v = (half * 64 + k < nvalid) ? __uint_as_float(r[k]) : -INFINITY;
          mx4[k & 3] = fmaxf(mx4[k & 3], v);
        }
      }
      const float mnew = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.c;
      float alpha = 1.f;
      const bool grow = mnew > m_used + AT2_TAU;  // key 0 of tile 0 always exists: the first tile always sets the reference
      if (grow) {
        alpha = ex2_approx(m_used - mnew);  // 0 on the first tile (m_used = -inf)
        m_used = mnew;
        l *= alpha;
      }
      if (j > 0) {
        mbar_wait(smem_u32(&pv_done[i]), ph ^ 1);  // P_i(j - 1) V_{j-1} complete: P_i free, O_i consistent
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {  // rescale this warp's 32 rows of O_i (alpha = 1 for the rows that did not move)
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(t_o + (uint32_t)(c * 32), o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tmem_st32(t_o + (uint32_t)(c * 32), o);
          }
        }
      }
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t r[64];
        tmem_ld32(t_s + (uint32_t)(half * 64), *reinterpret_cast<uint32_t(*)[32]>(r));
        tmem_ld32(t_s + (uint32_t)(half * 64 + 32), *reinterpret_cast<uint32_t(*)[32]>(r + 32));
        tmem_ld_wait();
        if (half == 1) {  // S_i is in registers for the last time: the issuer may overwrite it with Q_i K_{j+1}^T
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&s_free[i]));
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int k0 = half * 64 + 2 * k;
          const float p0 = k0 < nvalid ? ex2_approx(fmaf(__uint_as_float(r[2 * k]), p.c, -m_used)) : 0.f;
          const float p1 = k0 + 1 < nvalid ? ex2_approx(fmaf(__uint_as_float(r[2 * k + 1]), p.c, -m_used)) : 0.f;
          sum4[k & 3] += p0 + p1;
          __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
          r[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        tmem_st32(t_p + (uint32_t)(half * 32), *reinterpret_cast<uint32_t(*)[32]>(r));
      }
      l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_full[i]));
    }
#else
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(t_s + (uint32_t)(c * 32), *reinterpret_cast<uint32_t(*)[32]>(s + 32 * c));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_free[i]));
      if (nvalid < AT_K) {
#pragma unroll
        for (int k = 0; k < 128; ++k)
          if (k >= nvalid) s[k] = 0xff800000u;  // -inf
      }
      // eight independent chains (a single running maximum is a 128-deep dependent chain: ~500 cycles of pure latency per tile with
      // only two warps per scheduler to hide it)
      float mx8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) mx8[u] = __uint_as_float(s[u]);
#pragma unroll
      for (int k = 8; k < 128; k += 8)
#pragma unroll
        for (int u = 0; u < 8; ++u) mx8[u] = fmaxf(mx8[u], __uint_as_float(s[k + u]));
      const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      const float mnew = mx * p.c;
      float alpha = 1.f;
      const bool grow = mnew > m_used + AT2_TAU;  // key 0 of tile 0 always exists: the first tile always sets the reference
      if (grow) {
        alpha = ex2_approx(m_used - mnew);  // 0 on the first tile (m_used = -inf)
        m_used = mnew;
        l *= alpha;
      }
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * k]), p.c, -m_used));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * k + 1]), p.c, -m_used));
        sum4[k & 3] += p0 + p1;
        __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
        s[k] = *reinterpret_cast<uint32_t*>(&h);
      }
      l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      if (j > 0) {
        mbar_wait(smem_u32(&pv_done[i]), ph ^ 1);  // P_i(j - 1) V_{j-1} complete: P_i free, O_i consistent
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {  // rescale this warp's 32 rows of O_i (alpha = 1 for the rows that did not move)
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(t_o + (uint32_t)(c * 32), o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tmem_st32(t_o + (uint32_t)(c * 32), o);
          }
        }
      }
      tmem_st32(t_p, *reinterpret_cast<uint32_t(*)[32]>(s));
      tmem_st32(t_p + 32, *reinterpret_cast<uint32_t(*)[32]>(s + 32));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_full[i]));
    }
#endif
    // ---- epilogue: O / l -> activation -> bf16 rows
    const float inv_l = 1.f / l;
    mbar_wait(smem_u32(&pv_done[i]), (uint32_t)((nkt - 1) & 1));
    tc_fence_after();
    const int q = q0 + i * AT_Q + row;
#pragma unroll 1
    for (int c32 = 0; c32 < 2; ++c32) {
      uint32_t r[32];
      tmem_ld32(t_o + (uint32_t)(c32 * 32), r);
      tmem_ld_wait();
      if (q < p.ntok) {
        uint32_t packed[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float v0 = __uint_as_float(r[2 * k]) * inv_l, v1 = __uint_as_float(r[2 * k + 1]) * inv_l;
          if (p.act == 1) { v0 = mish<false>(v0); v1 = mish<false>(v1); }
          __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
          packed[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)q * p.ld_out + head * AT_D + c32 * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[u] = make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// vt[ch][tok] = v[tok][ch] for ch < channels, tok < ntok (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256) attn_transpose_kernel(const __nv_bfloat16* __restrict__ v, int ld, int ntok, int channels,
                                                             __nv_bfloat16* __restrict__ vt, int ldt) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int tok = t0 + r, ch = c0 + tx;
    tile[r][tx] = (tok < ntok && ch < channels) ? v[(size_t)tok * ld + ch] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ch = c0 + r, tok = t0 + tx;
    if (ch < channels && tok < ntok) vt[(size_t)ch * ldt + tok] = tile[tx][r];
  }
}

struct AttnTcPlan {
  AttnTcParams p;
  const __nv_bfloat16* v;
  __nv_bfloat16* vt;
  int ld_v, npad, inner;
  dim3 grid;
  size_t smem;
  int version, nq;   // 1: two-pass kernel; 2: single-pass kernel with nq query tiles per CTA
};

}  // namespace diqt

using namespace diqt;

struct diqt_attn_plan {
  AttnTcPlan a;
};

extern "C" int diqt_attn_tc_supported(int dtype, int dim_head, int ld_q, int ld_k, int ld_v, int ld_out) {
  return dtype == DIQT_BF16 && dim_head == AT_D && ld_q % 8 == 0 && ld_k % 8 == 0 && ld_v % 8 == 0 && ld_out % 8 == 0;
}

extern "C" int diqt_attn_tc_workspace_bytes(int tokens, int heads, size_t* bytes) {
  DIQT_REQUIRE(bytes && tokens > 0 && heads > 0, "attn_tc_workspace_bytes: bad arguments");
  const size_t npad = ((size_t)tokens + 127) / 128 * 128;
  *bytes = (size_t)heads * AT_D * npad * 2;
  return DIQT_OK;
}

extern "C" int diqt_attn_tc_plan_create(const void* q, const void* k, const void* v, int ld_q, int ld_k, int ld_v, void* out, int ld_out, int tokens,
                                        int heads, float scale, int act, void* workspace, diqt_attn_plan** plan) {
  DIQT_REQUIRE(q && k && v && out && workspace && plan && tokens > 0 && heads > 0, "attn_tc_plan_create: bad arguments");
  DIQT_REQUIRE(diqt_attn_tc_supported(DIQT_BF16, AT_D, ld_q, ld_k, ld_v, ld_out), "attn_tc_plan_create: pitches must be multiples of 8");
  DIQT_REQUIRE(act == 0 || act == 1, "attn_tc_plan_create: act %d (0 none, 1 Mish)", act);
  DIQT_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out | (uintptr_t)workspace) % 16 == 0, "attn_tc_plan_create: pointers must be 16-byte aligned");
  diqt_attn_plan* pl = new (std::nothrow) diqt_attn_plan();
  DIQT_REQUIRE(pl, "attn_tc_plan_create: out of host memory");
  AttnTcPlan& a = pl->a;
  memset(&a.p, 0, sizeof(a.p));
  const int inner = heads * AT_D;
  a.inner = inner;
  a.npad = (tokens + 127) / 128 * 128;
  a.v = (const __nv_bfloat16*)v;
  a.vt = (__nv_bfloat16*)workspace;
  a.ld_v = ld_v;
  AttnTcParams& p = a.p;
  p.out = (__nv_bfloat16*)out;
  p.ld_out = ld_out; p.ntok = tokens; p.heads = heads; p.q_col0 = 0; p.k_col0 = 0; p.act = act;
  p.c = scale * 1.4426950408889634f;
  p.idesc_s = make_idesc_bf16(AT_Q, AT_K);
  p.idesc_o = make_idesc_bf16(AT_Q, AT_D);
  p.idesc_o_mn = p.idesc_o | (1u << 16);
  // q / k: [tokens][ld] rows, the head's 64 channels are one 128-byte box row; v^T: [inner][npad]
  int rc = encode_volume_map(&p.q_map, q, inner, tokens, 1, 1, 1, ld_q, (int64_t)tokens * ld_q, (int64_t)tokens * ld_q, (int64_t)tokens * ld_q, AT_Q, 1, 1, 1);
  if (rc == DIQT_OK)
    rc = encode_volume_map(&p.k_map, k, inner, tokens, 1, 1, 1, ld_k, (int64_t)tokens * ld_k, (int64_t)tokens * ld_k, (int64_t)tokens * ld_k, AT_K, 1, 1, 1);
  if (rc == DIQT_OK)
    rc = encode_volume_map(&p.vt_map, workspace, a.npad, inner, 1, 1, 1, a.npad, (int64_t)inner * a.npad, (int64_t)inner * a.npad,
                           (int64_t)inner * a.npad, AT_D, 1, 1, 1);
  if (rc == DIQT_OK)
    rc = encode_volume_map(&p.v_map, v, inner, tokens, 1, 1, 1, ld_v, (int64_t)tokens * ld_v, (int64_t)tokens * ld_v, (int64_t)tokens * ld_v, AT_K, 1, 1, 1);
  if (rc != DIQT_OK) {
    delete pl;
    return rc;
  }
  const char* ev = getenv("DIQT_ATTN_TC_VERSION");   // variable: A/B against the two-pass kernel
  a.version = (ev && ev[0] == '1') ? 1 : 2;
  if (a.version == 1) {
    a.nq = 1;
    a.grid = dim3((tokens + AT_Q - 1) / AT_Q, heads);
    a.smem = (size_t)AT_TILE * 6 + 256 + 1024;  // Q + 2 K + 2 V^T chunks (= 1 tile) + 2 P, barriers, alignment slack: two CTAs per SM
  } else {
    // two query tiles per CTA (ping-pong) when that still leaves at least one CTA per SM, else one
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    a.nq = ((tokens + 2 * AT_Q - 1) / (2 * AT_Q)) * heads >= sms ? 2 : 1;
    a.grid = dim3((tokens + a.nq * AT_Q - 1) / (a.nq * AT_Q), heads);
    a.smem = (size_t)AT_TILE * (a.nq + AT2_KSTAGES + AT2_VSTAGES) + 256 + 1024;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(softmax_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(softmax_attn_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(softmax_attn_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) {
      set_error("attn_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      delete pl;
      return DIQT_ECUDA;
    }
    attr_done = true;
  }
  *plan = pl;
  return DIQT_OK;
}

extern "C" void diqt_attn_tc_plan_destroy(diqt_attn_plan* plan) { delete plan; }

extern "C" int diqt_attn_tc_run(const diqt_attn_plan* plan, void* stream) {
  DIQT_REQUIRE(plan, "attn_tc_run: null plan");
  const AttnTcPlan& a = plan->a;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 tgrid((a.p.ntok + 31) / 32, (a.inner + 31) / 32);
  if (a.version == 1) {   // only the two-pass kernel reads the transposed copy
    attn_transpose_kernel<<<tgrid, 256, 0, st>>>(a.v, a.ld_v, a.p.ntok, a.inner, a.vt, a.npad);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (a.version == 1) softmax_attn_tc_kernel<<<a.grid, AT_THREADS, a.smem, st>>>(a.p);
  else if (a.nq == 2) softmax_attn_tc2_kernel<2><<<a.grid, 128 * 2 + 64, a.smem, st>>>(a.p);
  else softmax_attn_tc2_kernel<1><<<a.grid, 128 + 64, a.smem, st>>>(a.p);
  return check_launch("softmax_attention_tc");
}
