// "z-march" tcgen05 convolution: 3x3x3, C_in = 64*KC -> C_out = 64*NH channels, bf16: every Block.project of the U-Net
// (imagen_pytorch3D.py:550-553; 64->64 at full resolution alone carries 77 % of the FLOPs).
//
// Two measurements on B200 shape this kernel (tools/umma_probe.cu, profiles/umma_probe_r1.log):
//   1. tcgen05.mma M=128 takes 64 cycles per K=16 step for N=64 AND for N=128: an N=64 GEMM can only reach half of
//      the tensor peak.  So the three dz taps that read the SAME input voxels are stacked along N: one MMA with
//      N = 192 multiplies an input z-plane by [W(dz=+1) | W(dz=0) | W(dz=-1)] and accumulates into the three output
//      planes p-1, p, p+1 at once (full rate).
//   2. The 128B-swizzle XOR is taken from absolute shared-memory address bits, so a K-major descriptor may start at
//      any 128-byte row of a TMA-written tile (base_offset = 0).  So the nine (dy,dx) taps are nine row-shifted views
//      (start = (kh*10+kw) rows, 8-row groups 10 rows apart) of ONE haloed input plane of 18 x 10 voxels.
//
// A CTA marches along z through a column of 16(y) x 8(x) voxels: every input plane is loaded once (1.4x halo
// redundancy instead of 27x) and contributes 9 MMAs of N=192.  Output plane z lives in TMEM block (z - z0) mod 4 of
// the slot's 256 columns; it is complete after input plane z+1, is drained / stored / re-zeroed by the epilogue warps
// while the tensor core already works on the next planes.  Two such columns ("slots") are processed in lockstep so
// that each 24 KB weight stage streamed from L2 is used twice.
//
// Wider layers: the input plane is walked in KC chunks of 64 channels (one TMA box per chunk, 9 weight stages each, all
// accumulating into the same TMEM blocks) and the output channels in NH groups of 64 that are separate work items (the
// per-tap kernel this replaces re-read every input voxel 27 times and was L2->SM bandwidth bound, profiles/r1f_*).
//
// ncu (profiles/r1f_ncu_conv_summary.md) showed ONE issuing warp could not keep the tensor pipe fed (~104 cycles of
// descriptor arithmetic per UTCHMMA against 64-96 cycles of pipe time), so each slot has its own issuing warp.
//
// Warp roles (256 threads): four epilogue warps, plane TMA producer, weight producer, two MMA issuers (slot 0 / 1; the first owns
// TMEM); see the role constants in the kernel for the order and why it matters.
//
// Fused GroupNorm + FiLM + Mish (kGN = true, 512 threads): Block.forward is GroupNorm -> FiLM -> Mish -> conv
// (imagen_pytorch3D.py:555-565) and the normalisation needs the statistics of the whole tensor, so it cannot ride on the PRODUCER's
// epilogue; it rides on the CONSUMER's load path instead.  The TMA lands the raw plane, eight more warps rewrite it in
// place as mish(a_c * x + b_c) (zero padding rows stay zero), fence it towards the async proxy and only then hand it to the MMA
// issuer (in the kGN instantiation warps 0-7 do this and the other roles follow).  Per plane and slot that is 11 520 elements = 2 MUFU + ~12 issue slots each against 9 x 4 x 96 = 3456 tensor-pipe cycles,
// i.e. ~40 % of the MUFU and ~30 % of the issue budget of the SM, and it removes one full read + write of the activation tensor and
// one kernel launch per convolution (38 per U-Net forward at the driver config).  The per-channel (a, b) come from the producer's
// grouped statistics and are finalised in this kernel's prologue exactly like affine_mish_kernel does (common.cuh), so the values
// entering the tensor cores are bit-identical to the two-kernel path.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "tc_common.cuh"

namespace diqt {

// DIQT_XF_MODE (build-time, timing experiments only; results are wrong unless 0): 1 = the transform warps only hand the plane on,
// 2 = affine without Mish.  Separates the cost of the extra pipeline stage from the cost of its arithmetic (profiles/r2f_xf_modes.md).
#ifndef DIQT_XF_MODE
#define DIQT_XF_MODE 0
#endif

// DIQT_ZM_TRACE (build-time, diagnostics only): CTA 0 records clock64() at the hand-over points of its pipeline into g_zm_trace
// [event][slot][plane iteration]; tools/zm_trace.py prints the timeline (profiles/r2_zm_timeline.md).
#ifndef DIQT_ZM_TRACE
#define DIQT_ZM_TRACE 0
#endif
// DIQT_XF_PACKED (build-time, A/B): the transform's affine + Mish on packed fp32 pairs
#ifndef DIQT_XF_PACKED
#define DIQT_XF_PACKED 1
#endif
#if DIQT_ZM_TRACE
__device__ long long g_zm_trace[8 * 2 * 16];
#define ZM_TRACE(ev, slot, it) do { if (blockIdx.x == 0 && (it) < 16) g_zm_trace[((ev) * 2 + (slot)) * 16 + (it)] = clock64(); } while (0)
#else
#define ZM_TRACE(ev, slot, it) do { } while (0)
#endif

constexpr int ZM_TX = 8, ZM_TY = 16;                    // output tile of one plane: 16 (y) x 8 (x) = 128 GEMM rows
constexpr int ZM_PLANE_BYTES = (ZM_TY + 2) * (ZM_TX + 2) * 128;  // 180 haloed rows x 64 bf16 = 23040
constexpr int ZM_PLANE_STRIDE = 23552;                  // next multiple of 1024
#ifndef DIQT_ZM_RING
#define DIQT_ZM_RING 2
#endif
#ifndef DIQT_ZM_WSTAGES
#define DIQT_ZM_WSTAGES 4
#endif
constexpr int ZM_RING = DIQT_ZM_RING;                   // plane buffers per slot
constexpr int ZM_WBLOCK = 64 * 128;                     // one (dz,dy,dx) weight block: 64 c_out rows x 64 c_in
constexpr int ZM_WSTAGE = 3 * ZM_WBLOCK;
constexpr int ZM_WSTAGES = DIQT_ZM_WSTAGES;
constexpr int ZM_THREADS = 256;
constexpr int ZM_THREADS_GN = 512;                      // + 8 warps that apply GroupNorm + FiLM + Mish to the landed planes
constexpr int ZM_GN_MAX_CIN = 256, ZM_GN_MAX_N = 2;      // (a, b) of every (volume, channel) live in shared memory
constexpr int ZM_PLANE_ROWS = (ZM_TY + 2) * (ZM_TX + 2);
constexpr int ZM_OUT_BYTES = 128 * 128;
constexpr int ZM_MAX_COUT = 512;                        // BASELINE config 5 sweeps up to 512 channels
constexpr int ZM_MAX_SEG = 64;                          // z-segments per column (boundaries live in the kernel parameters)

struct ZmParams {
  CUtensorMap in_map, out_map;
  const uint8_t* w;   // [nh][kc][kb = kh*3+kw][j: 0 -> kd=2 (dz=+1), 1 -> kd=1, 2 -> kd=0 (dz=-1)][64 c_out][64 c_in], pre-swizzled
  const float* bias;  // [c_out]
  float* stats;       // NULL or [n][2*gridDim.x][c_out][2]
  StatsGroups sink;   // optional grouped reduction of the statistics rows (common.cuh)
  GnParams gn;        // kGN: GroupNorm (+FiLM) of the INPUT, from the grouped statistics of its producer (gn.group != NULL)
  const float* aff_a; // kGN, alternative: per-(volume, channel) affine y = mish(a * x + b) already finalised by diqt_gn_finalize
  const float* aff_b; //      ([n][c_in] fp32, any batch size / width)
  int n, D, H, W;
  int KC, NH, c_out;  // c_in / 64, c_out / 64
  int tiles_x, tiles_y, nseg;
  short zs[ZM_MAX_SEG + 1];  // segment s covers output planes [zs[s], zs[s+1])
  int ipn, ipn_pad;   // work items per output-channel group (padded to even so that a slot pair shares its weights)
  int items, pairs;
  uint32_t idesc[3];  // N = 64, 128, 192 (M = 128, or M = 256 for the CTA-pair kernel)
  // k2 (CTA pair, tcgen05.mma.cta_group::2): a work item is a PAIR of x-adjacent columns (one per CTA of the pair) over one z-segment
  int ncp, ncp_pad;   // column pairs (n * tiles_y * tiles_x / 2), padded to even so that both slots always share their z-segment
  CUtensorMap w_map;  // the packed weights as rows of 64 channels (128 bytes), box = 32 rows, no TMA swizzle (they are pre-swizzled)
};

struct ZmItem {
  int valid, b, nh, x0, y0, z0, z1, p_lo, niter;
};

template <bool k2 = false>
__device__ __forceinline__ ZmItem zm_item(const ZmParams& p, int item, int rank = 0) {
  ZmItem it;
  it.nh = item / p.ipn_pad;
  const int r = item - it.nh * p.ipn_pad;
  int seg, t;
  if (k2) {  // column pair fastest: the two slots of a CTA pair (items 2q, 2q + 1) always share their z-segment
    seg = r / p.ncp_pad;
    t = r - seg * p.ncp_pad;
    it.valid = item < p.items && t < p.ncp;
  } else {
    seg = r % p.nseg;
    t = r / p.nseg;
    it.valid = item < p.items && r < p.ipn;
  }
  if (!it.valid) { it.b = it.x0 = it.y0 = it.z0 = it.z1 = it.p_lo = 0; it.niter = 0; return it; }
  const int ntx = k2 ? p.tiles_x / 2 : p.tiles_x;
  int tx = t % ntx; t /= ntx;
  if (k2) tx = 2 * tx + rank;
  const int ty = t % p.tiles_y;
  it.b = t / p.tiles_y;
  it.x0 = tx * ZM_TX; it.y0 = ty * ZM_TY;
  it.z0 = p.zs[seg]; it.z1 = p.zs[seg + 1];
  it.p_lo = max(it.z0 - 1, 0);
  it.niter = min(it.z1, p.D - 1) - it.p_lo + 1;
  return it;
}

// weight sub-blocks j (output plane p-1+j) needed when input plane pl feeds outputs [z0, z1)
__device__ __forceinline__ void zm_jrange(int pl, int z0, int z1, int& jlo, int& jhi) {
  jlo = max(0, z0 - (pl - 1));
  jhi = min(2, (z1 - 1) - (pl - 1));
}

// ---- CTA-pair helpers (k2) --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Remote arrive with the DEFAULT semantics (.release at CTA scope), as production 2-SM GEMMs do: what the arrival publishes is consumed
// by tcgen05 instructions (already ordered by fence.proxy.async / tcgen05.fence in the signalling thread), not by ld/st of the peer.  The
// first version used .release.cluster, which ptxas turns into a full ERRBAR fence per arrival: 16 % of the kernel's stall samples
// (profiles/r2d_pair_stalls.md).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // (arrivals come from the peer CTA too; default semantics, see above)
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      ".reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t"
      "}" ::"r"(bar)
      : "memory");
}
// TMA loads of the pair: data lands in THIS CTA's shared memory, the bytes are counted on a barrier of the leader CTA
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1)
               : "memory");
}

// k2: two CTAs on the SMs of one TPC run ONE tcgen05.mma.cta_group::2 per step (M = 256: 128 voxel rows from each CTA's own plane, N
// split between the CTAs' weight halves).  Shared memory serves ~135 B/clk per SM and the tensor pipe's operand fetch has priority
// (tools/umma_probe.cu E4, profiles/umma_probe_e4_r2c.log): at M = 128, N = 192 the MMAs alone fetch 107 B/clk, leaving a quarter of
// what the plane / weight TMA writes, the plane transform and the epilogue need at full tensor rate -- the single-CTA kernel is bound by
// shared-memory bandwidth at ~70-80 % tensor utilisation.  In the pair each CTA fetches its 128 A rows but only HALF of B (73 B/clk) and
// lands half of every weight stage.  Barriers that gate the single issuing CTA (weights landed, planes ready, accumulators drained) live
// in the LEADER (cluster rank 0) and are signalled from both CTAs; barriers that gate per-CTA roles (stage / plane / accumulator
// consumed by the MMAs) exist in both CTAs and are signalled by the leader's multicast tcgen05.commit.
template <bool kGN, bool k2>
__global__ void __launch_bounds__(kGN ? ZM_THREADS_GN : ZM_THREADS, 1) conv_zm_kernel(const __grid_constant__ ZmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* planes = smem;                                               // [2 slots][ZM_RING][ZM_PLANE_STRIDE]
  uint8_t* wst = planes + 2 * ZM_RING * ZM_PLANE_STRIDE;                // [ZM_WSTAGES][ZM_WSTAGE]
  uint8_t* out_stage = wst + ZM_WSTAGES * ZM_WSTAGE;                    // 16 KB
  float* s_bias = reinterpret_cast<float*>(out_stage + ZM_OUT_BYTES);   // [ZM_MAX_COUT]
  float* s_red = s_bias + ZM_MAX_COUT;                                  // [4][64][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 4 * 64 * 2);
  uint64_t* pl_full = bars;                       // [2][ZM_RING]
  uint64_t* pl_empty = pl_full + 2 * ZM_RING;     // [2][ZM_RING]
  uint64_t* w_full = pl_empty + 2 * ZM_RING;      // [ZM_WSTAGES]
  uint64_t* w_empty = w_full + ZM_WSTAGES;        // [ZM_WSTAGES]  (one arrival per issuing warp)
  uint64_t* acc_full = w_empty + ZM_WSTAGES;      // [2][2]
  uint64_t* acc_free = acc_full + 4;              // [2][2]
  uint64_t* pl_ready = acc_free + 4;              // [2][ZM_RING]  kGN: plane normalised (one arrival per transform warp of the slot)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pl_ready + 2 * ZM_RING);
  float* aff = reinterpret_cast<float*>(bars + 64);  // kGN: a[n][c_in] then b[n][c_in], n <= ZM_GN_MAX_N, c_in <= ZM_GN_MAX_CIN

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = k2 ? (int)cluster_ctarank() : 0;                      // leader = 0
  const int unit0 = k2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // work units (slot pairs) are dealt to CTAs / CTA pairs round robin
  const int unit_stride = k2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto to_leader = [&](uint64_t* bar) -> uint32_t { return k2 ? map_to_cta(smem_u32(bar), 0) : smem_u32(bar); };
  // Warp roles.  The SM's warp scheduler prefers the HIGHEST warp id among the eligible warps of a sub-partition (measured,
  // B300_MICROARCH "arbiter priority"), so the latency-critical single-thread roles get the highest ids: the two MMA issuers, then the
  // two TMA producers, then the epilogue (its four warps must cover the four TMEM lane quarters: warp & 3), and the throughput-bound
  // plane-transform warps (kGN) the lowest.
  constexpr int W_XF = 0;                     // kGN: warps 0..7
  constexpr int W_EPI = kGN ? 8 : 0;          // 4 warps
  constexpr int W_PLANE = W_EPI + 4, W_WEIGHT = W_EPI + 5, W_ISSUE = W_EPI + 6;   // W_ISSUE, W_ISSUE + 1 (slot 0 / 1)

  for (int i = threadIdx.x; i < p.c_out; i += blockDim.x) s_bias[i] = p.bias[i];
  if (warp == W_PLANE && lane == 0) {
    for (int i = 0; i < 2 * ZM_RING; ++i) {
      mbar_init(smem_u32(&pl_full[i]), 1);
      mbar_init(smem_u32(&pl_empty[i]), 1);
      mbar_init(smem_u32(&pl_ready[i]), k2 ? 16 : 8);   // one arrival per transform warp (of both CTAs)
    }
    for (int i = 0; i < ZM_WSTAGES; ++i) { mbar_init(smem_u32(&w_full[i]), 1); mbar_init(smem_u32(&w_empty[i]), 2); }
    for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_free[i]), k2 ? 8 : 4); }   // one per epilogue warp
    fence_barrier_init();
  }
  if (k2) cluster_sync_all();  // both CTAs' barriers exist before anyone signals across the pair
  if (warp == W_ISSUE) {
    if (k2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (k2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= W_EPI && warp < W_EPI + 4) {  // all accumulator blocks start at zero: every MMA accumulates
    const uint32_t q = (uint32_t)(warp & 3) * 32;
    for (int c = 0; c < 16; ++c) tmem_st32_zero(tmem_base + (q << 16) + (uint32_t)c * 32);
    tmem_st_wait();
  }
  tc_fence_before();
  if (k2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
#if !DIQT_PDL_LATE_TRIGGER
  pdl_launch_dependents();
#endif
  // Everything above is private to this CTA.  Only the roles that touch activations / statistics of earlier kernels block in
  // pdl_wait(): the plane producer (reads the input) and the epilogue (writes the output and the statistics rows).  The weight
  // producer streams constant weights and runs ahead, so the first weight stages are already in flight when the predecessor
  // kernel drains; the issuers only wait on mbarriers fed by those two.

  if (warp == W_PLANE) {
    // ===================== input plane producer =====================
    if (lane == 0) {
      pdl_wait();
#if DIQT_PDL_LATE_TRIGGER
      pdl_launch_dependents();
#endif
      int ring[2] = {0, 0};
      uint32_t phase[2] = {0, 0};
      for (int pair = unit0; pair < p.pairs; pair += unit_stride) {
        const ZmItem it0 = zm_item<k2>(p, 2 * pair, rank), it1 = zm_item<k2>(p, 2 * pair + 1, rank);
        const int niter = max(it0.niter, it1.niter);
        for (int i = 0; i < niter; ++i) {
          // chunk-major, slot-minor: both issuers consume the weight stages of chunk kc in lockstep, so the producer must never
          // block on one slot's ring while the other slot still lacks an earlier chunk (KC > ZM_RING would deadlock)
          for (int kc = 0; kc < p.KC; ++kc) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              const ZmItem& it = s ? it1 : it0;
              if (i >= it.niter) continue;
              const int b = s * ZM_RING + ring[s];
              mbar_wait(smem_u32(&pl_empty[b]), phase[s] ^ 1);
              ZM_TRACE(0, s, i);
              if (k2 && !kGN) {
                // the issuer (leader CTA) waits for BOTH CTAs' planes on its own barrier; with kGN the transform warps of each CTA wait
                // on their local barrier instead and report to the leader's pl_ready
                if (rank == 0) mbar_expect_tx(smem_u32(&pl_full[b]), 2 * ZM_PLANE_BYTES);
                tma_load_5d_pair(smem_u32(planes + (size_t)b * ZM_PLANE_STRIDE), &p.in_map, to_leader(&pl_full[b]), kc * 64, it.x0 - 1, it.y0 - 1,
                                 it.p_lo + i, it.b);
              } else {
                const uint32_t bar = smem_u32(&pl_full[b]);
                mbar_expect_tx(bar, ZM_PLANE_BYTES);
                tma_load_5d(smem_u32(planes + (size_t)b * ZM_PLANE_STRIDE), &p.in_map, bar, kc * 64, it.x0 - 1, it.y0 - 1, it.p_lo + i, it.b);
              }
              if (++ring[s] == ZM_RING) { ring[s] = 0; phase[s] ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == W_WEIGHT) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pair = unit0; pair < p.pairs; pair += unit_stride) {
        const ZmItem it0 = zm_item<k2>(p, 2 * pair, rank), it1 = zm_item<k2>(p, 2 * pair + 1, rank);
        const int niter = max(it0.niter, it1.niter);
        const uint8_t* wnh = p.w + (size_t)it0.nh * p.KC * 27 * ZM_WBLOCK;
        for (int i = 0; i < niter; ++i) {
          int jlo = 2, jhi = 0;
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const ZmItem& it = s ? it1 : it0;
            if (i >= it.niter) continue;
            int a, b;
            zm_jrange(it.p_lo + i, it.z0, it.z1, a, b);
            jlo = min(jlo, a); jhi = max(jhi, b);
          }
          const uint32_t bytes = (uint32_t)(jhi - jlo + 1) * ZM_WBLOCK;
          // k2: an MMA over weight rows [a, a + N) takes rows [a, a + N/2) from the leader's shared memory and rows [a + N/2, a + N) from
          // the peer's, both AT THE POSITION of row a (one descriptor serves both CTAs).  So each CTA lands its half of every run at the
          // run's natural position: two runs at most (the accumulator window may wrap around the 4-block TMEM ring, see the issuer).
          int ra[2] = {0, 0}, rn[2] = {0, 0};
          if (k2) {
            const int nb = jhi - jlo + 1, blk = (it0.p_lo + i - 1 + jlo - it0.z0) & 3;
            const int len0 = min(nb, 4 - blk);
            ra[0] = jlo * 64; rn[0] = len0 * 64;
            ra[1] = (jlo + len0) * 64; rn[1] = (nb - len0) * 64;
          }
          for (int kk = 0; kk < 9 * p.KC; ++kk) {  // kk = kc * 9 + kb
            mbar_wait(smem_u32(&w_empty[stage]), phase ^ 1);
            if (k2) {
              if (rank == 0) mbar_expect_tx(smem_u32(&w_full[stage]), bytes);  // both halves count on the leader's barrier
              const uint32_t bar = to_leader(&w_full[stage]);
              const int row0 = ((it0.nh * p.KC * 9 + kk) * 3) * 64;            // first row of this (chunk, tap) in the packed weights
#pragma unroll
              for (int r = 0; r < 2; ++r)
                for (int q = 0; q < rn[r] / 2; q += 32)
                  tma_load_2d_pair(smem_u32(wst + (size_t)stage * ZM_WSTAGE + (size_t)(ra[r] + q) * 128), &p.w_map, bar, 0,
                                   row0 + ra[r] + rank * (rn[r] / 2) + q);
            } else {
              const uint32_t bar = smem_u32(&w_full[stage]);
              mbar_expect_tx(bar, bytes);
              bulk_load(smem_u32(wst + (size_t)stage * ZM_WSTAGE + (size_t)jlo * ZM_WBLOCK), wnh + ((size_t)kk * 3 + jlo) * ZM_WBLOCK, bytes, bar);
            }
            if (++stage == ZM_WSTAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= W_ISSUE) {
    // ===================== MMA issuer of slot s =====================
    const int s = warp - W_ISSUE;
    auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { if (k2) umma_bf16_2cta(d, a, b, idesc, acc); else umma_bf16(d, a, b, idesc, acc); };
    auto commit = [](uint32_t bar) { if (k2) umma_commit_pair(bar); else umma_commit(bar); };
    auto mbar_wait_any = [](uint32_t bar, uint32_t parity) { if (k2) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity); };
    int stage = 0;
    uint32_t wphase = 0;
    int ring = 0;
    uint32_t rphase = 0;
    int kcount = 0;  // plane iterations issued so far for this slot (pairs up with the epilogue's counter)
    for (int pair = unit0; pair < p.pairs && rank == 0; pair += unit_stride) {  // k2: only the leader CTA issues (for both)
      const ZmItem it = zm_item<k2>(p, 2 * pair + s, 0), ot = zm_item<k2>(p, 2 * pair + 1 - s, 0);
      const int niter = max(it.niter, ot.niter);
      for (int i = 0; i < niter; ++i) {
        const bool act = i < it.niter;
        // the needed accumulator blocks, split into (at most two) runs that do not wrap around the 4-block ring of TMEM columns
        uint32_t b_off0 = 0, b_off1 = 0, d0 = 0, d1 = 0, idesc0 = 0, idesc1 = 0;
        bool run1 = false;
        if (act) {
          const int pl = it.p_lo + i;
          int jlo, jhi;
          zm_jrange(pl, it.z0, it.z1, jlo, jhi);
          const int nb = jhi - jlo + 1, blk = (pl - 1 + jlo - it.z0) & 3;  // TMEM block of the first needed output plane
          const int len0 = min(nb, 4 - blk);
          b_off0 = (uint32_t)jlo * ZM_WBLOCK;
          d0 = tmem_base + (uint32_t)(s * 256 + blk * 64);
          idesc0 = p.idesc[len0 - 1];
          run1 = len0 < nb;  // the rest of the window starts again at block 0
          b_off1 = (uint32_t)(jlo + len0) * ZM_WBLOCK;
          d1 = tmem_base + (uint32_t)(s * 256);
          idesc1 = p.idesc[run1 ? nb - len0 - 1 : 0];
          const int k = kcount;
          // the block written for the first time in this iteration was drained two iterations ago; a new item needs
          // every block of the slot drained
          if (k >= 2) mbar_wait_any(smem_u32(&acc_free[s * 2 + (k & 1)]), (uint32_t)(((k - 2) >> 1) & 1));
          if (i == 0 && k >= 1) mbar_wait_any(smem_u32(&acc_free[s * 2 + ((k - 1) & 1)]), (uint32_t)(((k - 1) >> 1) & 1));
        }
        for (int kc = 0; kc < p.KC; ++kc) {
          uint32_t a_base = 0;
          if (act) {
            mbar_wait_any(smem_u32(kGN ? &pl_ready[s * ZM_RING + ring] : &pl_full[s * ZM_RING + ring]), rphase);
            if (lane == 0) ZM_TRACE(3, s, i);
            a_base = smem_u32(planes + (size_t)(s * ZM_RING + ring) * ZM_PLANE_STRIDE);
          }
          tc_fence_after();
          for (int kb = 0; kb < 9; ++kb) {
            mbar_wait_any(smem_u32(&w_full[stage]), wphase);
            tc_fence_after();
            if (act) {  // warp-uniform issue code; the single issuing lane is elected inside umma_bf16 / umma_commit
              const uint32_t w_addr = smem_u32(wst + (size_t)stage * ZM_WSTAGE);
              const uint64_t adesc = make_sw128_desc_sbo(a_base + (uint32_t)((kb / 3) * (ZM_TX + 2) + (kb % 3)) * 128, (ZM_TX + 2) * 128);
              {
                const uint64_t bdesc = make_sw128_desc(w_addr + b_off0);
#pragma unroll
                for (int k = 0; k < 4; ++k) mma(d0, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc0, 1u);
              }
              if (run1) {
                const uint64_t bdesc = make_sw128_desc(w_addr + b_off1);
#pragma unroll
                for (int k = 0; k < 4; ++k) mma(d1, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc1, 1u);
              }
              commit(smem_u32(&w_empty[stage]));
            } else {
              // idle slot (odd item count / shorter z-segment): stay in lockstep with the weight ring (of both CTAs)
              if (lane == 0) {
                mbar_arrive(smem_u32(&w_empty[stage]));
                if (k2) mbar_arrive_cluster(map_to_cta(smem_u32(&w_empty[stage]), 1));
              }
              __syncwarp();
            }
            if (++stage == ZM_WSTAGES) { stage = 0; wphase ^= 1; }
          }
          if (act) {
            if (lane == 0) ZM_TRACE(4, s, i);
            commit(smem_u32(&pl_empty[s * ZM_RING + ring]));
            if (++ring == ZM_RING) { ring = 0; rphase ^= 1; }
          }
        }
        if (act) {
          commit(smem_u32(&acc_full[s * 2 + (kcount & 1)]));
          ++kcount;
        }
      }
    }
  } else if (warp >= W_EPI) {
    // ===================== epilogue (four warps) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // GEMM row = y * 8 + x inside the tile
    const int et = threadIdx.x - W_EPI * 32;  // 0..127
    const int cp = et & 31, rq = et >> 5;
    const int nblk = 2 * (int)gridDim.x;
    const uint32_t stage_a = smem_u32(out_stage);
    pdl_wait();
    int kcount[2] = {0, 0};
    float st_s[2][2], st_q[2][2];
    int st_key[2] = {-1, -1};  // (volume, output-channel group) the running sums of a slot belong to
    st_s[0][0] = st_s[0][1] = st_s[1][0] = st_s[1][1] = 0.f;
    st_q[0][0] = st_q[0][1] = st_q[1][0] = st_q[1][1] = 0.f;
    if (p.stats) {
      // every (volume, channel) entry of this CTA's two partial rows must be defined: zero them, then overwrite what we produce
      for (int nv = 0; nv < p.n; ++nv)
        for (int s = 0; s < 2; ++s) {
          float* dst = p.stats + ((size_t)nv * nblk + blockIdx.x * 2 + s) * p.c_out * 2;
          for (int idx = et; idx < p.c_out * 2; idx += 128) dst[idx] = 0.f;
        }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    auto flush_stats = [&](int s, int key) {
      const int nvol = key / p.NH, nh = key - nvol * p.NH;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        s_red[(rq * 64 + cp * 2 + h) * 2] = st_s[s][h];
        s_red[(rq * 64 + cp * 2 + h) * 2 + 1] = st_q[s][h];
        st_s[s][h] = st_q[s][h] = 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float* dst = p.stats + (((size_t)nvol * nblk + blockIdx.x * 2 + s) * p.c_out + nh * 64) * 2;
      if (et < 64) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) { a += s_red[(r4 * 64 + et) * 2]; b += s_red[(r4 * 64 + et) * 2 + 1]; }
        dst[et * 2] = a;
        dst[et * 2 + 1] = b;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };

    for (int pair = unit0; pair < p.pairs; pair += unit_stride) {
      const ZmItem it0 = zm_item<k2>(p, 2 * pair, rank), it1 = zm_item<k2>(p, 2 * pair + 1, rank);
      const int niter = max(it0.niter, it1.niter);
      if (p.stats) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const ZmItem& it = s ? it1 : it0;
          if (!it.valid) continue;
          const int key = it.b * p.NH + it.nh;
          if (st_key[s] >= 0 && key != st_key[s]) flush_stats(s, st_key[s]);
          st_key[s] = key;
        }
      }
      for (int i = 0; i < niter; ++i) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const ZmItem& it = s ? it1 : it0;
          if (i >= it.niter) continue;
          const int pl = it.p_lo + i;
          const int k = kcount[s];
          const uint32_t bias_a = smem_u32(s_bias + it.nh * 64);
          mbar_wait(smem_u32(&acc_full[s * 2 + (k & 1)]), (uint32_t)((k >> 1) & 1));
          if (et == 0) ZM_TRACE(5, s, i);
          tc_fence_after();
          // outputs whose last contributing input plane is pl:  z = pl-1, and z = pl at the top face of the volume
          for (int which = 0; which < 2; ++which) {
            const int z = which == 0 ? pl - 1 : pl;
            if (which == 0 && z < it.z0) continue;
            if (which == 1 && !(pl == p.D - 1 && pl < it.z1)) continue;
            const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * 256 + ((z - it.z0) & 3) * 64);
            if (et == 0) bulk_wait_read0();  // staging tile free again
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int c32 = 0; c32 < 2; ++c32) {
              uint32_t r[32];
              tmem_ld32(tcol + c32 * 32, r);
              tmem_ld_wait();
              tmem_st32_zero(tcol + c32 * 32);  // the block is reused for output plane z+4
              // (bias, staging tile and statistics through 32-bit shared-window addresses and 16-byte bias loads: the generic pointers
              // cost 64 scalar bias loads and a 64-bit address computation per access)
              uint32_t packed[16];
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const uint4 bq = lds_128(bias_a + (uint32_t)(c32 * 32 + j4 * 4) * 4);
                // (two channels per FADD2: the same IEEE additions per lane)
                const float2 v01 = __fadd2_rn(make_float2(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1])), make_float2(__uint_as_float(bq.x), __uint_as_float(bq.y)));
                const float2 v23 = __fadd2_rn(make_float2(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])), make_float2(__uint_as_float(bq.z), __uint_as_float(bq.w)));
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v01.x, v01.y);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(v23.x, v23.y);
                packed[2 * j4] = *reinterpret_cast<uint32_t*>(&h0);
                packed[2 * j4 + 1] = *reinterpret_cast<uint32_t*>(&h1);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int chunk = (c32 * 4 + j) ^ (row & 7);
                sts_128(stage_a + (uint32_t)(row * 128 + chunk * 16), packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
              }
            }
            fence_proxy_async();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) {
              tma_store_5d(&p.out_map, smem_u32(out_stage), it.nh * 64, it.x0, it.y0, z, it.b);
              bulk_commit();
            }
            if (p.stats) {
#pragma unroll 4
              for (int r = 0; r < 32; ++r) {
                const int rr = rq * 32 + r;
                const uint32_t v = lds_u32(stage_a + (uint32_t)(rr * 128 + (((cp >> 2) ^ (rr & 7)) << 4) + ((cp & 3) << 2)));
                const float2 lh = make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
                const float2 ns = __fadd2_rn(make_float2(st_s[s][0], st_s[s][1]), lh);          // (channel pair per instruction, same sums per lane)
                const float2 nq = __ffma2_rn(lh, lh, make_float2(st_q[s][0], st_q[s][1]));
                st_s[s][0] = ns.x; st_s[s][1] = ns.y; st_q[s][0] = nq.x; st_q[s][1] = nq.y;
              }
            }
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (et == 0) ZM_TRACE(6, s, i);
          if (lane == 0) {
            if (k2) mbar_arrive_cluster(to_leader(&acc_free[s * 2 + (k & 1)]));
            else mbar_arrive(smem_u32(&acc_free[s * 2 + (k & 1)]));
          }
          ++kcount[s];
        }
      }
    }
    if (p.stats) {
#pragma unroll
      for (int s = 0; s < 2; ++s)
        if (st_key[s] >= 0) flush_stats(s, st_key[s]);
      stats_group_tail(p.sink, p.stats, p.n, nblk, p.c_out, (int)blockIdx.x * 2, 2, et, 128, reinterpret_cast<int*>(s_red),
                       [] { asm volatile("bar.sync 1, 128;" ::: "memory"); });
    }
    if (et == 0) bulk_wait0();
  } else if (kGN) {
    // ===================== GroupNorm + FiLM + Mish of the landed input planes (warps 0..7) =====================
    // All eight warps work on ONE plane at a time, in the order the producer issues them (slot 0, slot 1 alternating): the latency from
    // "plane landed" to "plane ready" is what the two-deep plane ring has to hide behind one plane's worth of MMAs.
    const int tt = threadIdx.x - W_XF * 32;  // 0..255
    const int c_in = p.KC * 64;
    auto sync256 = [] { asm volatile("bar.sync 2, 256;" ::: "memory"); };
#if DIQT_XF_MODE == 3   // timing experiment: no finalisation prologue (identity affine)
    if (true) {
      pdl_wait();
      for (int ch = tt; ch < c_in * p.n; ch += 256) { aff[ch] = 1.f; aff[p.n * c_in + ch] = 0.f; }
      sync256();
    } else
#endif
    if (!p.gn.group) {
      pdl_wait();  // (a, b) were written by the finalize kernel in front of this one
    } else {
      // (a, b) per (volume, channel) from the producer's grouped statistics.  The scratch aliases the output staging tile, which the
      // epilogue cannot touch before the first accumulator is complete, i.e. not before these warps have released a plane.
      const GnScratch sc = gn_scratch_layout(out_stage, c_in, p.gn.groups, 256);
      gn_prefetch_constants(p.gn, 0, tt, 256, sc);  // constants of volume 0 while the producer kernel is still draining
      pdl_wait();
      for (int nv = 0; nv < p.n; ++nv) {
        if (nv > 0) gn_prefetch_constants(p.gn, nv, tt, 256, sc);   // (every thread reads back only the constants it wrote itself)
        gn_affine_from_groups(p.gn, nv, tt, 256, sc, sync256);
        for (int ch = tt; ch < c_in; ch += 256) {                    // same thread -> channel mapping as inside: no barrier in between
          aff[nv * c_in + ch] = sc.a_s[ch];
          aff[(p.n + nv) * c_in + ch] = sc.b_s[ch];
        }
      }
      sync256();
    }
    // thread -> (physical 16-byte chunk pc, rows rbase + 32 k): the swizzled chunk holds logical chunk pc ^ (row & 7), and
    // (rbase + 32 k) & 7 == rbase & 7, so one thread always works on the same eight channels of a 64-channel chunk
    const int pc = tt & 7, rbase = tt >> 3;
    const int ch0 = (pc ^ (rbase & 7)) << 3;
    const uint32_t aff_a = smem_u32(aff);
    constexpr int XF_ROWS = (ZM_PLANE_ROWS + 31) / 32;  // 6 row groups of 32
    int ring[2] = {0, 0};
    uint32_t phase[2] = {0, 0};
    for (int pair = unit0; pair < p.pairs; pair += unit_stride) {
      const ZmItem it0 = zm_item<k2>(p, 2 * pair, rank), it1 = zm_item<k2>(p, 2 * pair + 1, rank);
      const int niter = max(it0.niter, it1.niter);
      uint32_t vmask[2] = {0, 0};  // rows of this thread that lie inside the volume (the others are the zero padding: left untouched)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const ZmItem& it = s ? it1 : it0;
#pragma unroll
        for (int k = 0; k < XF_ROWS; ++k) {
          const int r = rbase + 32 * k;
          const int ry = r / (ZM_TX + 2), rx = r - ry * (ZM_TX + 2);
          const int y = it.y0 - 1 + ry, x = it.x0 - 1 + rx;
          if (r < ZM_PLANE_ROWS && (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W) vmask[s] |= 1u << k;
        }
      }
      for (int i = 0; i < niter; ++i) {
        for (int kc = 0; kc < p.KC; ++kc) {
          // (runtime loops, small unrolled bodies: the five roles of this kernel share the SM's instruction cache; the first version of
          // this loop, fully unrolled over slots and rows, spent a fifth of its samples waiting for instructions)
#pragma unroll 1
          for (int s = 0; s < 2; ++s) {
            const ZmItem& it = s ? it1 : it0;
            if (i >= it.niter) continue;
            float av[8], bv[8];
            {
              float4 a0, a1, b0, b1;
              if (p.gn.group) {
                const uint32_t ap = aff_a + (uint32_t)(it.b * c_in + kc * 64 + ch0) * 4, bp = aff_a + (uint32_t)((p.n + it.b) * c_in + kc * 64 + ch0) * 4;
                const uint4 qa0 = lds_128(ap), qa1 = lds_128(ap + 16), qb0 = lds_128(bp), qb1 = lds_128(bp + 16);
                a0 = make_float4(__uint_as_float(qa0.x), __uint_as_float(qa0.y), __uint_as_float(qa0.z), __uint_as_float(qa0.w));
                a1 = make_float4(__uint_as_float(qa1.x), __uint_as_float(qa1.y), __uint_as_float(qa1.z), __uint_as_float(qa1.w));
                b0 = make_float4(__uint_as_float(qb0.x), __uint_as_float(qb0.y), __uint_as_float(qb0.z), __uint_as_float(qb0.w));
                b1 = make_float4(__uint_as_float(qb1.x), __uint_as_float(qb1.y), __uint_as_float(qb1.z), __uint_as_float(qb1.w));
              } else {  // L2 round trip, issued before (and hidden behind) the wait for the plane
                const float4* ap = reinterpret_cast<const float4*>(p.aff_a + (size_t)it.b * c_in + kc * 64 + ch0);
                const float4* bp = reinterpret_cast<const float4*>(p.aff_b + (size_t)it.b * c_in + kc * 64 + ch0);
                a0 = __ldcg(ap); a1 = __ldcg(ap + 1); b0 = __ldcg(bp); b1 = __ldcg(bp + 1);
              }
              av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
              bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
            }
            const int b = s * ZM_RING + ring[s];
            const uint32_t vm = s ? vmask[1] : vmask[0];
            mbar_wait(smem_u32(&pl_full[b]), s ? phase[1] : phase[0]);
            if (tt == 0) ZM_TRACE(1, s, i);
            // (32-bit shared-window addresses: through the generic pointer every access cost a 64-bit address computation)
            const uint32_t base = smem_u32(planes + (size_t)b * ZM_PLANE_STRIDE + rbase * 128 + pc * 16);
#if DIQT_XF_MODE == 1   // timing experiment: handshake only, the plane is handed on untouched
            if (false)
#endif
#pragma unroll 1
            for (int k0 = 0; k0 < XF_ROWS; k0 += 3) {
              uint4 raw[3];
#pragma unroll
              for (int u = 0; u < 3; ++u)
                if ((vm >> (k0 + u)) & 1u) raw[u] = lds_128(base + (uint32_t)((k0 + u) * 32 * 128));
#pragma unroll
              for (int u = 0; u < 3; ++u) {
                if (!((vm >> (k0 + u)) & 1u)) continue;
                Vec<__nv_bfloat16> r;
                r.unpack(raw[u]);
#if DIQT_XF_MODE == 2   // timing experiment: affine only (no MUFU)
#pragma unroll
                for (int e = 0; e < 8; ++e) r.v[e] = fmaf(av[e], r.v[e], bv[e]);
#elif DIQT_XF_PACKED
                // two channels per instruction on the FMA pipe (FFMA2 / FMUL2 / FADD2, sm_100): per lane the same IEEE operations in the same
                // order as mish<true>(fmaf(a, x, b)), so the values entering the tensor cores stay bit-identical to the two-kernel path
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                  const float2 y = __ffma2_rn(make_float2(av[e], av[e + 1]), make_float2(r.v[e], r.v[e + 1]), make_float2(bv[e], bv[e + 1]));
                  const float2 t = __fmul2_rn(y, make_float2(1.4426950408889634f, 1.4426950408889634f));
                  float2 u, w;
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u.x) : "f"(t.x));
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u.y) : "f"(t.y));
                  const float2 d = __ffma2_rn(u, __fadd2_rn(u, make_float2(2.f, 2.f)), make_float2(2.f, 2.f));
                  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(w.x) : "f"(d.x));
                  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(w.y) : "f"(d.y));
                  const float2 o = __ffma2_rn(__fmul2_rn(y, w), make_float2(-2.f, -2.f), y);
                  r.v[e] = o.x;
                  r.v[e + 1] = o.y;
                }
#else
#pragma unroll
                for (int e = 0; e < 8; ++e) r.v[e] = mish<true>(fmaf(av[e], r.v[e], bv[e]));
#endif
                {
                  uint32_t w[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(r.v[2 * e], r.v[2 * e + 1]);
                    w[e] = *reinterpret_cast<uint32_t*>(&h);
                  }
                  sts_128(base + (uint32_t)((k0 + u) * 32 * 128), w[0], w[1], w[2], w[3]);
                }
              }
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (tt == 0) ZM_TRACE(2, s, i);
            if (lane == 0) {
              if (k2) mbar_arrive_cluster(to_leader(&pl_ready[b]));
              else mbar_arrive(smem_u32(&pl_ready[b]));
            }
            if (s) { if (++ring[1] == ZM_RING) { ring[1] = 0; phase[1] ^= 1; } }
            else   { if (++ring[0] == ZM_RING) { ring[0] = 0; phase[0] ^= 1; } }
          }
        }
      }
    }
  }

  tc_fence_before();
  if (k2) cluster_sync_all(); else __syncthreads();  // k2: neither CTA may leave while the other can still signal its barriers
  if (warp == W_ISSUE) {
    tc_fence_after();
    if (k2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// (c_out, c_in, 3,3,3) fp32 -> [nh][kc][kb = kh*3+kw][j][64 c_out][64 c_in] bf16 with 16-byte chunks XOR-swizzled by (row & 7)
__global__ void conv_pack_zm_kernel(const float* __restrict__ w, int c_in, int c_out, __nv_bfloat16* __restrict__ packed) {
  const int KC = c_in / 64;
  const int64_t total = (int64_t)27 * c_in * c_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % 64), r = (int)((i / 64) % 64), j = (int)((i / 4096) % 3), kb = (int)((i / 12288) % 9);
    const int kc = (int)((i / (12288 * 9)) % KC), nh = (int)(i / ((int64_t)12288 * 9 * KC));
    const int kd = 2 - j, kh = kb / 3, kw = kb % 3;
    const float v = w[((int64_t)(nh * 64 + r) * c_in + kc * 64 + e) * 27 + kd * 9 + kh * 3 + kw];
    const int chunk = (e >> 3) ^ (r & 7);
    packed[(i / 64) * 64 + chunk * 8 + (e & 7)] = __float2bfloat16_rn(v);
  }
}

struct ZmPlan {
  ZmParams p;
  int grid;
  size_t smem;
  bool pair;   // CTA-pair kernel (k2)
};

// The z-march kernel as a CTA pair: needs an even number of x tiles (the two CTAs of a pair take x-adjacent columns) and at least two
// column pairs per launch so that both slots of the pair have work.  DIQT_ZM_2CTA=0 switches it off (A/B measurements).
static bool zm_use_pair(const diqt_conv_desc* d) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("DIQT_ZM_2CTA");
    env = (e && e[0] == '0') ? 0 : 1;
  }
  const int tiles_x = d->d2 / ZM_TX, tiles_y = d->d1 / ZM_TY;
  // measured (profiles/sweep_conv_r2e.jsonl): the pair wins from 64^3 up (41.3 vs 42.9 us at 64 channels, 2.36 vs 2.44 ms at 512) and
  // loses a few percent on 32^3 volumes, where its two cluster barriers and cross-CTA signalling are a larger share of a 15-30 us kernel
  const char* mp = getenv("DIQT_ZM_2CTA_MIN_PAIRS");   // read per plan: the parity tests lower it to exercise the pair kernel on small volumes
  const int min_pairs = mp ? atoi(mp) : 16;
  return env == 1 && !(d->flags & DIQT_CONV_FLAG_NO_CTA_PAIR) && tiles_x % 2 == 0 && (int64_t)d->n * tiles_y * (tiles_x / 2) >= min_pairs;
}

bool conv_zm_supported(const diqt_conv_desc* d) {
  if (d->mode != DIQT_CONV_K3 || d->dtype != DIQT_BF16) return false;
  if (d->c_in % 64 != 0 || d->c_out % 64 != 0 || d->c_out > ZM_MAX_COUT) return false;
  if (d->ld_in % 8 != 0 || d->ld_out % 8 != 0) return false;
  if (d->d2 % ZM_TX != 0 || d->d1 % ZM_TY != 0 || d->d0 < 2) return false;
  return true;
}

// The per-tap kernel re-reads the input 27 times from L2; the z-march kernel wins whenever its tile shape fits.
bool conv_zm_profitable(const diqt_conv_desc* d) { return conv_zm_supported(d); }

size_t conv_zm_packed_bytes(const diqt_conv_desc* d) { return (size_t)27 * d->c_in * d->c_out * 2; }

int conv_zm_pack(const diqt_conv_desc* d, const float* w, void* packed, cudaStream_t st) {
  conv_pack_zm_kernel<<<256, 256, 0, st>>>(w, d->c_in, d->c_out, (__nv_bfloat16*)packed);
  return check_launch("conv_pack_zm");
}

// tensor-pipe cycles (per K=16 step) one z-segment [z0, z1) of a column costs: N=64 and N=128 take 64 cycles, N=192 takes 96,
// and a three-block window that wraps around the 4-block TMEM ring is issued as N=128 + N=64 (profiles/umma_probe_r1.log)
static int zm_segment_cycles(int z0, int z1, int D) {
  const int p_lo = std::max(z0 - 1, 0), niter = std::min(z1, D - 1) - p_lo + 1;
  int cyc = 0;
  for (int i = 0; i < niter; ++i) {
    const int pl = p_lo + i;
    const int jlo = std::max(0, z0 - (pl - 1)), jhi = std::min(2, (z1 - 1) - (pl - 1));
    const int nb = jhi - jlo + 1, blk = (pl - 1 + jlo - z0) & 3;
    cyc += nb <= 2 ? (blk + nb > 4 ? 128 : 64) : (blk + nb > 4 ? 128 : 96);
  }
  return cyc;
}

int conv_zm_plan(const diqt_conv_desc* d, const void* in, void* out, const void* packed, const float* bias, ZmPlan** out_plan) {
  DIQT_REQUIRE(conv_zm_supported(d), "conv(zm): needs 3x3x3 bf16, c_in %% 64 == 0, c_out %% 64 == 0 (<= %d), d2 %% 8 == 0 and d1 %% 16 == 0", ZM_MAX_COUT);
  ZmPlan* plan = new ZmPlan();
  ZmParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.w = (const uint8_t*)packed;
  p.bias = bias;
  p.stats = nullptr;
  p.n = d->n; p.D = d->d0; p.H = d->d1; p.W = d->d2;
  p.KC = d->c_in / 64; p.NH = d->c_out / 64; p.c_out = d->c_out;
  p.tiles_x = d->d2 / ZM_TX;
  p.tiles_y = d->d1 / ZM_TY;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // pick the z-segmentation: minimise  rounds x (tensor cycles of the slowest slot pair + fixed per-CTA cost); ties go to fewer
  // segments (fewer halo planes streamed from L2).  Segments are balanced by COST, not length: an interior segment of n planes
  // reads n + 2 input planes, the two at the ends of the volume n + 1, so the end segments get one more output plane
  // (64 planes in 9 segments = 8,7,7,7,7,7,7,7,7: nine input planes each, 144 CTAs, instead of 8 x 8: ten input planes, 128 CTAs).
  plan->pair = zm_use_pair(d);
  const bool k2 = plan->pair;
  // units of the column dimension: columns, or x-adjacent column pairs (padded to even) for the CTA-pair kernel
  p.ncp = (int)((int64_t)d->n * p.tiles_y * (p.tiles_x / 2));
  p.ncp_pad = p.ncp + (p.ncp & 1);
  const int64_t cols = k2 ? p.ncp_pad : (int64_t)d->n * p.tiles_x * p.tiles_y;
  const int workers = k2 ? sms / 2 : sms;   // CTAs or CTA pairs
  const int D = d->d0;
  auto split = [&](int nseg, short* zs) {
    // lengths: T - 1 at both ends, T - 2 inside, with T the smallest per-segment plane budget that covers D; surplus removed from the back
    int T = 3;
    while (true) {
      const int cap = nseg == 1 ? T : 2 * (T - 1) + (nseg - 2) * (T - 2);
      if (cap >= D) break;
      ++T;
    }
    int len[ZM_MAX_SEG];
    int total = 0;
    for (int i = 0; i < nseg; ++i) { len[i] = (nseg == 1) ? D : ((i == 0 || i == nseg - 1) ? T - 1 : T - 2); total += len[i]; }
    for (int i = nseg - 1; total > D; i = (i == 0 ? nseg - 1 : i - 1))
      if (len[i] > 1) { --len[i]; --total; }
    zs[0] = 0;
    for (int i = 0; i < nseg; ++i) zs[i + 1] = (short)(zs[i] + len[i]);
  };
  double best = 1e30;
  int best_nseg = 1;
  for (int nseg = 1; nseg <= std::min(D, ZM_MAX_SEG); ++nseg) {
    short zs[ZM_MAX_SEG + 1];
    split(nseg, zs);
    if (zs[nseg] != D) continue;
    const int64_t ipn = cols * nseg, ipn_pad = ipn + (ipn & 1);
    const int64_t pairs = (int64_t)p.NH * ipn_pad / 2;
    const int64_t rounds = (pairs + workers - 1) / workers;
    int worst = 0;
    for (int sgi = 0; sgi < nseg; ++sgi) worst = std::max(worst, zm_segment_cycles(zs[sgi], zs[sgi + 1], D));
    const double cost = (double)rounds * (2.0 * worst * 36.0 * p.KC + 6000.0);
    if (cost < best * 0.99) { best = cost; best_nseg = nseg; }
  }
  p.nseg = best_nseg;
  split(p.nseg, p.zs);
  p.ipn = (int)(cols * p.nseg);
  p.ipn_pad = p.ipn + (p.ipn & 1);
  p.items = p.NH * p.ipn_pad;
  p.pairs = p.items / 2;
  for (int i = 0; i < 3; ++i) p.idesc[i] = make_idesc_bf16(k2 ? 256 : 128, 64 * (i + 1));
  const int64_t ld = d->ld_in, lo = d->ld_out;
  int rc = encode_volume_map(&p.in_map, in, d->c_in, d->d2, d->d1, d->d0, d->n, ld, (int64_t)d->d2 * ld, (int64_t)d->d1 * d->d2 * ld,
                             (int64_t)d->d0 * d->d1 * d->d2 * ld, ZM_TX + 2, ZM_TY + 2, 1, 1);
  if (rc == DIQT_OK)
    rc = encode_volume_map(&p.out_map, out, d->c_out, d->d2, d->d1, d->d0, d->n, lo, (int64_t)d->d2 * lo, (int64_t)d->d1 * d->d2 * lo,
                           (int64_t)d->d0 * d->d1 * d->d2 * lo, ZM_TX, ZM_TY, 1, 1);
  if (rc != DIQT_OK) {
    delete plan;
    return rc;
  }
  if (k2 && rc == DIQT_OK) {
    // packed weights as a 2-D tensor of 128-byte rows; 32-row boxes; the rows are already XOR-swizzled in global memory
    const cuuint64_t gdim[2] = {64, (cuuint64_t)27 * d->c_in * d->c_out / 64};
    const cuuint64_t gstr[1] = {128};
    const cuuint32_t box[2] = {64, 32}, estr[2] = {1, 1};
    const CUresult cr = get_encode_tiled()(&p.w_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(packed), gdim, gstr, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("conv(zm): cuTensorMapEncodeTiled(weights) failed (%d)", (int)cr);
      rc = DIQT_ECUDA;
    }
  }
  if (rc != DIQT_OK) {
    delete plan;
    return rc;
  }
  plan->grid = k2 ? 2 * (p.pairs < workers ? p.pairs : workers) : (p.pairs < sms ? p.pairs : sms);
  plan->smem = (size_t)2 * ZM_RING * ZM_PLANE_STRIDE + (size_t)ZM_WSTAGES * ZM_WSTAGE + ZM_OUT_BYTES + ZM_MAX_COUT * 4 + 4 * 64 * 2 * 4 + 512 + 1024 +
               (size_t)2 * ZM_GN_MAX_N * ZM_GN_MAX_CIN * 4;
  static bool attr_done = false;
  if (!attr_done) {
    DIQT_CUDA(cudaFuncSetAttribute(conv_zm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DIQT_CUDA(cudaFuncSetAttribute(conv_zm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DIQT_CUDA(cudaFuncSetAttribute(conv_zm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DIQT_CUDA(cudaFuncSetAttribute(conv_zm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  *out_plan = plan;
  return DIQT_OK;
}

// PDL launch of a kernel whose CTAs come in pairs (thread-block cluster of two: the unit tcgen05.mma.cta_group::2 works on)
template <typename K>
static void launch_pair(K kernel, int grid, int threads, size_t smem, cudaStream_t st, const ZmParams& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, p);  // errors surface in check_launch()
}

int conv_zm_run(const ZmPlan* plan, cudaStream_t st) {
  const bool gn = plan->p.gn.group || plan->p.aff_a;
  if (plan->pair) {
    if (gn) launch_pair(conv_zm_kernel<true, true>, plan->grid, ZM_THREADS_GN, plan->smem, st, plan->p);
    else launch_pair(conv_zm_kernel<false, true>, plan->grid, ZM_THREADS, plan->smem, st, plan->p);
  } else if (gn) {
    launch_pdl(conv_zm_kernel<true, false>, plan->grid, ZM_THREADS_GN, plan->smem, st, plan->p);
  } else {
    launch_pdl(conv_zm_kernel<false, false>, plan->grid, ZM_THREADS, plan->smem, st, plan->p);
  }
  return check_launch("conv_zm");
}

// GroupNorm (+FiLM) + Mish of the input folded into the plane path.  The statistics must be the grouped kind (<= 16 rows).
bool conv_zm_gn_supported(const diqt_conv_desc* d) {
  return conv_zm_supported(d) && d->c_in <= ZM_GN_MAX_CIN && d->n <= ZM_GN_MAX_N;
}

int conv_zm_set_gn(ZmPlan* plan, const GnParams& gn) {
  const int c_in = plan->p.KC * 64;
  DIQT_REQUIRE(c_in <= ZM_GN_MAX_CIN && plan->p.n <= ZM_GN_MAX_N, "conv(zm) fused GroupNorm: c_in=%d (<= %d), n=%d (<= %d)", c_in, ZM_GN_MAX_CIN,
               plan->p.n, ZM_GN_MAX_N);
  DIQT_REQUIRE(gn.group && gn.ngroups > 0 && gn.gamma && gn.beta && gn.groups > 0 && c_in % gn.groups == 0 && gn.c == c_in,
               "conv(zm) fused GroupNorm: bad description (c=%d, c_in=%d, groups=%d)", gn.c, c_in, gn.groups);
  DIQT_REQUIRE(gn_scratch_bytes(c_in, gn.groups, 256) <= (size_t)ZM_OUT_BYTES, "conv(zm) fused GroupNorm: finalisation scratch of %zu bytes does not fit",
               gn_scratch_bytes(c_in, gn.groups, 256));
  plan->p.gn = gn;
  plan->p.aff_a = plan->p.aff_b = nullptr;
  return DIQT_OK;
}

int conv_zm_set_gn_affine(ZmPlan* plan, const float* a, const float* b) {
  DIQT_REQUIRE(a && b, "conv(zm) fused GroupNorm: null affine");
  plan->p.gn = GnParams{};
  plan->p.aff_a = a;
  plan->p.aff_b = b;
  return DIQT_OK;
}

void conv_zm_set_film(ZmPlan* plan, const float* film, int film_ld, const int* film_row, int film_row_stride_n) {
  plan->p.gn.film = film;
  plan->p.gn.film_ld = film_ld;
  plan->p.gn.film_row = film_row;
  plan->p.gn.film_row_stride_n = film_row_stride_n;
}

int conv_zm_set_stats(ZmPlan* plan, float* partial, float* group, unsigned int* tickets, int* ngroups) {
  plan->p.stats = partial;
  const int nblk = 2 * plan->grid;
  StatsGroups& g = plan->p.sink;
  g.group = group;
  g.tickets = tickets;
  g.gsize = stats_group_size(nblk, 2);
  g.ngroups = (nblk + g.gsize - 1) / g.gsize;
  if (ngroups) *ngroups = group ? g.ngroups : 0;
  return nblk;
}

void conv_zm_destroy(ZmPlan* plan) { delete plan; }

}  // namespace diqt

#if DIQT_ZM_TRACE
extern "C" int diqt_debug_zm_trace(long long* out) {
  return cudaMemcpyFromSymbol(out, diqt::g_zm_trace, sizeof(diqt::g_zm_trace)) == cudaSuccess ? 0 : -1;
}
#endif
