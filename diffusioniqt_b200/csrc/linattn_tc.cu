// Linear attention on the 5th-gen tensor cores (LinearAttention.forward imagen_pytorch3D.py:1001-1011):
//   q = softmax_d(q) * scale ; k = softmax_n(k) ; ctx = k^T v (per head, d x d) ; out = Mish(q ctx)
// bf16, dim_head = 64, an even number of heads.  The op is bandwidth work (4 N d^2 FLOP per head against 4 N d values: 32 FLOP per
// byte), but 32 FLOP / byte at HBM speed is ~200 TFLOP/s: more than the CUDA cores deliver, so both products run as tcgen05.mma and
// q, k, v are read exactly ONCE (k a second time, from L2, for its column maximum), the output written once.
//
//   1. linattn_colmax_kernel    column maxima of k over the tokens (the softmax over n needs a global reference), 8 rows in flight per thread
//   2. linattn_ctx_tc_kernel    CTA = (token chunk, head pair).  TMA lands [128 tokens][64 ch] boxes of k and v (two heads each) in a
//                               three-stage ring; eight warps rewrite the k boxes in place as P = exp(k - max) (bf16) and keep the
//                               column sums; one thread issues  D[128 x 128] += P^T V  with BOTH operands MN-major: a SWIZZLE_128B
//                               box of [tokens][channels] rows IS the canonical MN-major operand layout (K = tokens, 8-token groups
//                               1024 B apart, the second head's box one leading-byte-offset away), so nothing is transposed
//                               anywhere.  M = 128 stacks two heads; the two diagonal 64 x 64 blocks of D are the contexts.
//   3. linattn_combine_kernel   partial contexts / column sums of the chunks summed in a fixed order (bitwise reproducible), scaled,
//                               and written as the bf16 K-major SWIZZLE_128B shared-memory image of ctx^T the last kernel bulk-copies
//   4. linattn_out_tc_kernel    CTA = (128-token tile, head pair): thread = token row rewrites q in place as exp(q - max_d), one
//                               tcgen05.mma group per head (M128 N64 K64) against the context image, epilogue scales by 1 / sum,
//                               Mish, bf16 rows.
#include <string.h>

#include <new>

#include "tc_common.cuh"

namespace diqt {

constexpr int LT_TOK = 128;               // tokens per tile
constexpr int LT_BOX = LT_TOK * 128;      // one [128 tokens][64 bf16] box: 16 KB
constexpr int LT_STAGE = 4 * LT_BOX;      // k (2 heads) + v (2 heads)
constexpr int LT_STAGES = 3;
constexpr int LT_CTX_THREADS = 320;       // warps 0-7 transform (0-3 also drain D), 8 TMA producer, 9 MMA issuer
constexpr int LT_OUT_THREADS = 288;       // warps 0-7 softmax / epilogue (thread = token row x head), 8 TMA + MMA issuer
constexpr int LT_MAXCHUNK = 128;
constexpr float LT_LOG2E = 1.4426950408889634f;

struct LinAttnParams {
  CUtensorMap q_map, k_map, v_map;
  const __nv_bfloat16* k;
  __nv_bfloat16* out;
  float* colmax;          // [inner]
  float* part;            // [nchunks][hp][128][64]
  float* spart;           // [nchunks][hp][128]
  __nv_bfloat16* ctxn;    // [hp][2][64 e][64 d]: shared-memory image (K-major, 128B swizzle) of scale * ctx^T / column sum
  int ld_qkv, ld_out, ntok, ntiles, heads, inner, hp, nmax, nchunks, act;
  float scale;
  uint32_t idesc_ctx, idesc_out;
};

// MN-major SWIZZLE_128B operand (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>): 64 MN-elements are contiguous (128 B), the
// next 64 are `lbo` bytes away; eight K-rows are 128 B apart, the next eight `sbo` bytes away.
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ float lt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void lt_unpack(const uint4& raw, float (&x)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[2 * i] = __uint_as_float(w[i] << 16);
    x[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
// two floats -> packed bf16 pair (round to nearest even) ; the rounded values come back in a, b
__device__ __forceinline__ uint32_t lt_pack(float& a, float& b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xFFFF0000u);
  return u;
}
__device__ __forceinline__ void bar_sync_256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------------------------
// 1. colmax[col] = max over the tokens of k[token][col].  A maximum does not depend on the order of its operands, so the CTAs combine
//    through atomics and the result is still bitwise reproducible.  colmax must hold a lower bound of the result on entry (-inf): the
//    combine kernel of the previous run re-arms it, plan_create arms the first run.  (A stale, larger value would only change the
//    reference point of the softmax, which cancels.)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
constexpr int LT_CM_THREADS = 1024;
__global__ void __launch_bounds__(LT_CM_THREADS) linattn_colmax_kernel(const __nv_bfloat16* k, int ld, int ntok, int inner, int nmax, float* colmax) {
  __shared__ float red[LT_CM_THREADS * 8];
  pdl_sync();
  const int ncc = inner >> 3;
  int ncc_pad = 1;
  while (ncc_pad < ncc) ncc_pad <<= 1;
  const int lanes_r = LT_CM_THREADS / ncc_pad, pitch = ncc_pad * 8;
  const int cc = threadIdx.x % ncc_pad, rl = threadIdx.x / ncc_pad;
  const int per = (ntok + nmax - 1) / nmax;
  const int r0 = blockIdx.x * per, r1 = min(ntok, r0 + per);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
  if (cc < ncc) {
    const __nv_bfloat16* src = k + cc * 8;
    int r = r0 + rl;
    for (; r + 7 * lanes_r < r1; r += 8 * lanes_r) {
      uint4 raw[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) raw[u] = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)(r + u * lanes_r) * ld));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float x[8];
        lt_unpack(raw[u], x);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], x[j]);
      }
    }
    uint4 raw[8];   // tail: up to seven more rows, still issued together
#pragma unroll
    for (int u = 0; u < 8; ++u)
      raw[u] = r + u * lanes_r < r1 ? __ldcg(reinterpret_cast<const uint4*>(src + (size_t)(r + u * lanes_r) * ld)) : make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float x[8];
      lt_unpack(raw[u], x);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], x[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl * pitch + cc * 8 + j] = m[j];
  __syncthreads();
  for (int col = threadIdx.x; col < inner; col += LT_CM_THREADS) {
    float mm = -INFINITY;
    for (int i = 0; i < lanes_r; ++i) mm = fmaxf(mm, red[i * pitch + col]);
    if (mm > -INFINITY) atomic_max_float(colmax + col, mm);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// 2. partial contexts of one token chunk for one head pair
__global__ void __launch_bounds__(LT_CTX_THREADS, 1) linattn_ctx_tc_kernel(const __grid_constant__ LinAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = smem;                                                   // LT_STAGES x 64 KB
  float* mref = reinterpret_cast<float*>(stages + LT_STAGES * LT_STAGE);    // [128]  -max * log2(e)
  float* red = mref + 128;                                                  // [16][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 16 * 128);
  uint64_t* full = bars;                   // [3] TMA landed
  uint64_t* empty = full + LT_STAGES;      // [3] the stage's MMAs completed
  uint64_t* ready = empty + LT_STAGES;     // [3] k rewritten as P (8 warp arrivals)
  uint64_t* d_full = ready + LT_STAGES;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x, hp = blockIdx.y;
  const int t0 = (int)((int64_t)chunk * p.ntiles / p.nchunks), t1 = (int)((int64_t)(chunk + 1) * p.ntiles / p.nchunks);
  const int niter = t1 - t0;

  if (warp == 8 && lane == 0) {
    for (int s = 0; s < LT_STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
      mbar_init(smem_u32(&ready[s]), 8);
    }
    mbar_init(smem_u32(d_full), 1);
    fence_barrier_init();
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===================== TMA producer: k and v were complete before this grid could start (the predecessor waited for them) ======
    if (lane == 0) {
      for (int it = 0; it < niter; ++it) {
        const int s = it % LT_STAGES;
        const uint32_t ph = (uint32_t)((it / LT_STAGES) & 1);
        mbar_wait(smem_u32(&empty[s]), ph ^ 1);
        mbar_expect_tx(smem_u32(&full[s]), LT_STAGE);
        const uint32_t dst = smem_u32(stages + s * LT_STAGE), bar = smem_u32(&full[s]);
        const int row0 = (t0 + it) * LT_TOK, c0 = hp * 128;
        tma_load_5d(dst, &p.k_map, bar, c0, row0, 0, 0, 0);
        tma_load_5d(dst + LT_BOX, &p.k_map, bar, c0 + 64, row0, 0, 0, 0);
        tma_load_5d(dst + 2 * LT_BOX, &p.v_map, bar, c0, row0, 0, 0, 0);
        tma_load_5d(dst + 3 * LT_BOX, &p.v_map, bar, c0 + 64, row0, 0, 0, 0);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (warp-uniform; the issuing lane is elected inside the wrappers) =====================
    for (int it = 0; it < niter; ++it) {
      const int s = it % LT_STAGES;
      const uint32_t ph = (uint32_t)((it / LT_STAGES) & 1);
      mbar_wait(smem_u32(&ready[s]), ph);
      tc_fence_after();
      const uint32_t kb = smem_u32(stages + s * LT_STAGE), vb = kb + 2 * LT_BOX;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {   // K = 16 tokens per instruction: two 8-token groups, 2048 B
        const uint64_t adesc = make_sw128_mn_desc(kb + kk * 2048, LT_BOX, 1024);
        const uint64_t bdesc = make_sw128_mn_desc(vb + kk * 2048, LT_BOX, 1024);
        umma_bf16(tmem_base, adesc, bdesc, p.idesc_ctx, (uint32_t)((it | kk) != 0));
      }
      umma_commit(smem_u32(&empty[s]));
    }
    umma_commit(smem_u32(d_full));
  } else {
    // ===================== transform: k -> P = exp(k - max) in place =====================
    const int t = threadIdx.x;
    pdl_sync();   // the column maxima come from the predecessor
    if (t < 128) mref[t] = -__ldcg(&p.colmax[hp * 128 + t]) * LT_LOG2E;
    bar_sync_256();
    // thread -> (box, physical 16-byte chunk pc, rows rb + 16 i): the chunk holds logical chunk pc ^ (row & 7) and (rb + 16 i) & 7 == rb & 7,
    // so a thread always works on the same eight channels
    const int box = t >> 7, tt = t & 127, pc = tt & 7, rb = tt >> 3, lc = pc ^ (rb & 7);
    float nm[8], cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      nm[j] = mref[box * 64 + lc * 8 + j];
      cs[j] = 0.f;
    }
    for (int it = 0; it < niter; ++it) {
      const int s = it % LT_STAGES;
      const uint32_t ph = (uint32_t)((it / LT_STAGES) & 1);
      mbar_wait(smem_u32(&full[s]), ph);
      uint8_t* kb = stages + s * LT_STAGE + box * LT_BOX + rb * 128 + pc * 16;
      const int tok0 = (t0 + it) * LT_TOK + rb;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4* cell = reinterpret_cast<uint4*>(kb + i * 2048);
        float x[8];
        lt_unpack(*cell, x);
        const bool valid = tok0 + 16 * i < p.ntok;   // rows past the end (TMA zero fill) must not count
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = valid ? lt_ex2(fmaf(x[j], LT_LOG2E, nm[j])) : 0.f;
        uint4 o;
        o.x = lt_pack(x[0], x[1]);
        o.y = lt_pack(x[2], x[3]);
        o.z = lt_pack(x[4], x[5]);
        o.w = lt_pack(x[6], x[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[j] += x[j];   // sums of the ROUNDED probabilities: consistent with the numerator
        *cell = o;
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ready[s]));
    }
    // column sums: 16 row groups in a fixed order
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rb * 128 + box * 64 + lc * 8 + j] = cs[j];
    bar_sync_256();
    const size_t slot = (size_t)chunk * p.hp + hp;
    if (t < 128) {
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) sum += red[g * 128 + t];
      p.spart[slot * 128 + t] = sum;
      // D: lane = row d' (head d' >> 6); the context of that head sits in columns [64 (d' >> 6), +64)
      mbar_wait(smem_u32(d_full), 0);
      tc_fence_after();
      const int hb = t >> 6;
      float* dst = p.part + (slot * 128 + t) * 64;
#pragma unroll 1
      for (int c32 = 0; c32 < 2; ++c32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(hb * 64 + c32 * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 8; ++u)
          reinterpret_cast<float4*>(dst + c32 * 32)[u] =
              make_float4(__uint_as_float(r[4 * u]), __uint_as_float(r[4 * u + 1]), __uint_as_float(r[4 * u + 2]), __uint_as_float(r[4 * u + 3]));
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// 3. ctxn = scale * sum_chunks part / sum_chunks spart, as the shared-memory image of ctx^T ([e][d] rows of 128 B, 16-byte chunks XOR-ed
//    with e & 7)
__global__ void __launch_bounds__(256) linattn_combine_kernel(const float* part, const float* spart, __nv_bfloat16* ctxn, int nchunks, int hp_count,
                                                              float scale, float* colmax, int inner) {
  // CTA = 16 items (hp, row, e4) x 16 chunk slices: slice q adds the chunks c = q, q + 16, ... in order, then the slice sums are added
  // in order: a fixed summation tree (bitwise reproducible), and the loads are spread over threads instead of queued in one
  __shared__ float red[15][16][5];
  pdl_sync();
  const int gid = blockIdx.x * 256 + threadIdx.x;
  if (gid < inner) colmax[gid] = -INFINITY;   // consumed by the predecessor: re-armed for the next run
  const int slice = threadIdx.x >> 4, li = threadIdx.x & 15;
  const int item = blockIdx.x * 16 + li;       // hp_count * 2048 items: a multiple of 16
  const int e4 = item & 15, row = (item >> 4) & 127, hp = item >> 11;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float s = 0.f;
  for (int c = slice; c < nchunks; c += 16) {
    const size_t slot = (size_t)c * hp_count + hp;
    const float4 v = __ldcg(reinterpret_cast<const float4*>(part + (slot * 128 + row) * 64 + e4 * 4));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    s += __ldcg(spart + slot * 128 + row);
  }
  if (slice > 0) {
    float* r = red[slice - 1][li];
    r[0] = acc.x; r[1] = acc.y; r[2] = acc.z; r[3] = acc.w; r[4] = s;
  }
  __syncthreads();
  if (slice > 0) return;
#pragma unroll
  for (int q = 0; q < 15; ++q) {
    const float* r = red[q][li];
    acc.x += r[0]; acc.y += r[1]; acc.z += r[2]; acc.w += r[3]; s += r[4];
  }
  const float f = scale / s;
  const int hb = row >> 6, d = row & 63;
  __nv_bfloat16* base = ctxn + (size_t)(hp * 2 + hb) * 4096;
  const float vals[4] = {acc.x * f, acc.y * f, acc.z * f, acc.w * f};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = e4 * 4 + u;
    base[e * 64 + (((d >> 3) ^ (e & 7)) << 3) + (d & 7)] = __float2bfloat16(vals[u]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// 4. out = act( softmax_d(q) (scale ctx) ) for one 128-token tile of one head pair
__global__ void __launch_bounds__(LT_OUT_THREADS, 4) linattn_out_tc_kernel(const __grid_constant__ LinAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;                       // 2 boxes
  uint8_t* c_s = q_s + 2 * LT_BOX;           // 2 x 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(c_s + 2 * 8192);
  uint64_t* q_full = bars;
  uint64_t* c_full = bars + 1;
  uint64_t* ready = bars + 2;                // [2] per head: 4 warp arrivals
  uint64_t* d_full = bars + 4;               // [2] per head
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, hp = blockIdx.y;

  if (warp == 8) {
    if (lane == 0) {
      mbar_init(smem_u32(q_full), 1);
      mbar_init(smem_u32(c_full), 1);
      for (int hb = 0; hb < 2; ++hb) {
        mbar_init(smem_u32(&ready[hb]), 4);
        mbar_init(smem_u32(&d_full[hb]), 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(q_full), 2 * LT_BOX);   // q was complete before this grid could start
      tma_load_5d(smem_u32(q_s), &p.q_map, smem_u32(q_full), hp * 128, tile * LT_TOK, 0, 0, 0);
      tma_load_5d(smem_u32(q_s + LT_BOX), &p.q_map, smem_u32(q_full), hp * 128 + 64, tile * LT_TOK, 0, 0, 0);
      pdl_sync();                                     // the context image comes from the predecessor
      mbar_expect_tx(smem_u32(c_full), 2 * 8192);
      bulk_load(smem_u32(c_s), p.ctxn + (size_t)hp * 8192, 2 * 8192, smem_u32(c_full));
    }
    __syncwarp();
    mbar_wait(smem_u32(c_full), 0);
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      mbar_wait(smem_u32(&ready[hb]), 0);
      tc_fence_after();
      const uint64_t adesc = make_sw128_desc(smem_u32(q_s + hb * LT_BOX));
      const uint64_t bdesc = make_sw128_desc(smem_u32(c_s + hb * 8192));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + (uint32_t)(hb * 64), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc_out, (uint32_t)(k != 0));
      umma_commit(smem_u32(&d_full[hb]));
    }
  } else {
    const int hb = warp >> 2, row = threadIdx.x & 127;   // warps 4 hb .. 4 hb + 3 own head hb; token row = TMEM lane, quarter = warp & 3
    mbar_wait(smem_u32(q_full), 0);
    uint8_t* base = q_s + hb * LT_BOX + row * 128;
    // chunk order rotated by the lane: the eight lanes of a quarter warp touch eight different bank groups (rows are 128 B apart);
    // the row is read twice from shared memory (maximum, then exponentials) instead of being held in 32 registers
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x[8];
      lt_unpack(*reinterpret_cast<const uint4*>(base + (((j + lane) & 7) << 4)), x);
#pragma unroll
      for (int u = 0; u < 8; ++u) m = fmaxf(m, x[u]);
    }
    const float nmx = -m * LT_LOG2E;
    float sum = 0.f;
#pragma unroll 2
    for (int j = 0; j < 8; ++j) {
      float x[8];
      lt_unpack(*reinterpret_cast<const uint4*>(base + (((j + lane) & 7) << 4)), x);
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = lt_ex2(fmaf(x[u], LT_LOG2E, nmx));
      uint4 o;
      o.x = lt_pack(x[0], x[1]);
      o.y = lt_pack(x[2], x[3]);
      o.z = lt_pack(x[4], x[5]);
      o.w = lt_pack(x[6], x[7]);
#pragma unroll
      for (int u = 0; u < 8; ++u) sum += x[u];
      *reinterpret_cast<uint4*>(base + (((j + lane) & 7) << 4)) = o;
    }
    const float inv = 1.f / sum;
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&ready[hb]));
    mbar_wait(smem_u32(&d_full[hb]), 0);
    tc_fence_after();
    const int n = tile * LT_TOK + row;
#pragma unroll 1
    for (int c32 = 0; c32 < 2; ++c32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(hb * 64 + c32 * 32), r);
      tmem_ld_wait();
      if (n < p.ntok) {
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float v0 = __uint_as_float(r[2 * i]) * inv, v1 = __uint_as_float(r[2 * i + 1]) * inv;
          if (p.act == 1) { v0 = mish<false>(v0); v1 = mish<false>(v1); }
          const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
          packed[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)n * p.ld_out + hp * 128 + hb * 64 + c32 * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[u] = make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

struct LinAttnPlan {
  LinAttnParams p;
  size_t smem_ctx, smem_out;
};

static void linattn_layout(int tokens, int heads, int sms, int* ntiles, int* nmax, int* nchunks, size_t* off_part, size_t* off_spart, size_t* off_ctxn,
                           size_t* total) {
  const int hp = heads / 2, inner = heads * 64;
  *ntiles = (tokens + LT_TOK - 1) / LT_TOK;
  *nmax = *ntiles < LT_MAXCHUNK ? *ntiles : LT_MAXCHUNK;
  int nc = sms / hp;
  if (nc < 1) nc = 1;
  if (nc > *ntiles) nc = *ntiles;
  *nchunks = nc;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t off = up((size_t)inner * 4);
  *off_part = off;
  off += up((size_t)nc * hp * 128 * 64 * 4);
  *off_spart = off;
  off += up((size_t)nc * hp * 128 * 4);
  *off_ctxn = off;
  off += up((size_t)hp * 2 * 4096 * 2);
  *total = off;
}

static int linattn_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  (void)cudaGetLastError();
  return sms > 0 ? sms : 148;
}

}  // namespace diqt

using namespace diqt;

struct diqt_linattn_plan {
  LinAttnPlan a;
};

extern "C" int diqt_linattn_tc_supported(int dtype, int dim_head, int heads, int ld_qkv, int ld_out) {
  return dtype == DIQT_BF16 && dim_head == 64 && heads > 0 && heads % 2 == 0 && heads <= 32 && ld_qkv % 8 == 0 && ld_out % 8 == 0;
}

extern "C" int diqt_linattn_tc_workspace_bytes(int tokens, int heads, size_t* bytes) {
  DIQT_REQUIRE(bytes && tokens > 0 && heads > 0 && heads % 2 == 0, "linattn_tc_workspace_bytes: bad arguments");
  int ntiles, nmax, nchunks;
  size_t o1, o2, o3;
  linattn_layout(tokens, heads, linattn_sms(), &ntiles, &nmax, &nchunks, &o1, &o2, &o3, bytes);
  return DIQT_OK;
}

extern "C" int diqt_linattn_tc_plan_create(const void* q, const void* k, const void* v, int ld_qkv, void* out, int ld_out, int tokens, int heads,
                                           float scale, int act, void* workspace, diqt_linattn_plan** plan) {
  DIQT_REQUIRE(q && k && v && out && workspace && plan && tokens > 0, "linattn_tc_plan_create: bad arguments");
  DIQT_REQUIRE(diqt_linattn_tc_supported(DIQT_BF16, 64, heads, ld_qkv, ld_out), "linattn_tc_plan_create: needs an even number of heads (<= 32) and pitches that are multiples of 8");
  DIQT_REQUIRE(act == 0 || act == 1, "linattn_tc_plan_create: act %d (0 none, 1 Mish)", act);
  DIQT_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0 && (uintptr_t)workspace % 256 == 0,
               "linattn_tc_plan_create: q / k / v / out must be 16-byte aligned, the workspace 256-byte aligned");
  diqt_linattn_plan* pl = new (std::nothrow) diqt_linattn_plan();
  DIQT_REQUIRE(pl, "linattn_tc_plan_create: out of host memory");
  LinAttnParams& p = pl->a.p;
  memset(&p, 0, sizeof(p));
  size_t o_part, o_spart, o_ctxn, total;
  linattn_layout(tokens, heads, linattn_sms(), &p.ntiles, &p.nmax, &p.nchunks, &o_part, &o_spart, &o_ctxn, &total);
  uint8_t* ws = (uint8_t*)workspace;
  p.k = (const __nv_bfloat16*)k;
  p.out = (__nv_bfloat16*)out;
  p.colmax = (float*)ws;
  p.part = (float*)(ws + o_part);
  p.spart = (float*)(ws + o_spart);
  p.ctxn = (__nv_bfloat16*)(ws + o_ctxn);
  p.ld_qkv = ld_qkv; p.ld_out = ld_out; p.ntok = tokens; p.heads = heads; p.inner = heads * 64; p.hp = heads / 2; p.act = act;
  p.scale = scale;
  p.idesc_ctx = make_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);   // A and B MN-major
  p.idesc_out = make_idesc_bf16(128, 64);
  const int64_t all = (int64_t)tokens * ld_qkv;
  int rc = encode_volume_map(&p.q_map, q, p.inner, tokens, 1, 1, 1, ld_qkv, all, all, all, LT_TOK, 1, 1, 1);
  if (rc == DIQT_OK) rc = encode_volume_map(&p.k_map, k, p.inner, tokens, 1, 1, 1, ld_qkv, all, all, all, LT_TOK, 1, 1, 1);
  if (rc == DIQT_OK) rc = encode_volume_map(&p.v_map, v, p.inner, tokens, 1, 1, 1, ld_qkv, all, all, all, LT_TOK, 1, 1, 1);
  if (rc != DIQT_OK) {
    delete pl;
    return rc;
  }
  {   // arm the column maxima (see linattn_colmax_kernel); plans are created outside stream capture
    float minus_inf[2048];
    for (int i = 0; i < p.inner; ++i) minus_inf[i] = -INFINITY;
    cudaError_t e = cudaMemcpy(p.colmax, minus_inf, (size_t)p.inner * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("linattn_tc_plan_create: cudaMemcpy failed: %s", cudaGetErrorString(e));
      delete pl;
      return DIQT_ECUDA;
    }
  }
  pl->a.smem_ctx = (size_t)LT_STAGES * LT_STAGE + 128 * 4 + 16 * 128 * 4 + 128 + 1024;
  pl->a.smem_out = (size_t)2 * LT_BOX + 2 * 8192 + 64 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(linattn_ctx_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(linattn_out_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) {
      set_error("linattn_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      delete pl;
      return DIQT_ECUDA;
    }
    attr_done = true;
  }
  *plan = pl;
  return DIQT_OK;
}

extern "C" void diqt_linattn_tc_plan_destroy(diqt_linattn_plan* plan) { delete plan; }

extern "C" int diqt_linattn_tc_run(const diqt_linattn_plan* plan, void* stream) {
  DIQT_REQUIRE(plan, "linattn_tc_run: null plan");
  const LinAttnPlan& a = plan->a;
  const LinAttnParams& p = a.p;
  cudaStream_t st = (cudaStream_t)stream;
  launch_pdl(linattn_colmax_kernel, dim3(p.nmax), dim3(LT_CM_THREADS), 0, st, p.k, p.ld_qkv, p.ntok, p.inner, p.nmax, p.colmax);
  launch_pdl(linattn_ctx_tc_kernel, dim3(p.nchunks, p.hp), dim3(LT_CTX_THREADS), a.smem_ctx, st, p);
  launch_pdl(linattn_combine_kernel, dim3(p.hp * 128), dim3(256), 0, st, (const float*)p.part, (const float*)p.spart, p.ctxn, p.nchunks,
             p.hp, p.scale, p.colmax, p.inner);
  launch_pdl(linattn_out_tc_kernel, dim3(p.ntiles, p.hp), dim3(LT_OUT_THREADS), a.smem_out, st, p);
  g_launches.fetch_add(3, std::memory_order_relaxed);
  return check_launch("linear_attention_tc");
}
