// Patch extraction and stitching around the sampler, on the device.
// Reference (host Python, one patch at a time): /root/reference/data.py:139-202 (supervisedIQT_INF: patch grid, slicing),
// /root/reference/utils_mine.py:25-67 (96^3 <-> 27 x 32^3 re-tiling), /root/reference/test_all.py:239-300 (stitch with centre crops,
// later patches overwrite earlier ones, background mask).  Pure data movement: one gather kernel for a whole batch of patches, one
// stitch kernel for the whole volume (each output voxel looks up the LAST patch in grid order whose crop covers it).
#include "common.cuh"

namespace diqt {

// local voxel (a0, a1, a2) of a patch of side P -> element offset inside the patch's block of the batch tensor.
// f <= 1: (P, P, P) row-major.  f > 1: the patch is stored as f^3 sub-volumes of side h = P / f in the reference's order
// (sub-volume b = b0 + f*b1 + f*f*b2 is block (b0, b1, b2) along dims (0, 1, 2); convertVolume2subVolume, utils_mine.py:25-42).
__device__ __forceinline__ int64_t patch_offset(int a0, int a1, int a2, int P, int f) {
  if (f <= 1) return ((int64_t)a0 * P + a1) * P + a2;
  const int h = P / f;
  const int b = a0 / h + f * (a1 / h) + f * f * (a2 / h);
  return (int64_t)b * h * h * h + ((int64_t)(a0 % h) * h + a1 % h) * h + a2 % h;
}

__global__ void __launch_bounds__(256) gather_patches_kernel(const float* __restrict__ vol, int d1, int d2, const int* __restrict__ origins, int nb,
                                                             int P, int f, float* __restrict__ out) {
  const int64_t per = (int64_t)P * P * P;
  const int64_t total = per * nb;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per);
    int64_t r = i - (int64_t)b * per;
    const int a0 = (int)(r / ((int64_t)P * P));
    r -= (int64_t)a0 * P * P;
    const int a1 = (int)(r / P), a2 = (int)(r - (int64_t)a1 * P);
    const int o0 = origins[3 * b], o1 = origins[3 * b + 1], o2 = origins[3 * b + 2];
    out[(int64_t)b * per + patch_offset(a0, a1, a2, P, f)] = vol[((int64_t)(o0 + a0) * d1 + (o1 + a1)) * d2 + (o2 + a2)];
  }
}

struct StitchParams {
  const float* patches;  // [kept][P^3] in patch_offset() layout
  const int* slot;       // [g0*g1*g2]: slot of grid patch (i, j, k) in `patches`, or -1 if it was skipped (data.py:192-196)
  float* pred;           // (d0, d1, d2), pre-filled
  const float* lowres;   // optional background mask source (test_all.py:300)
  float min_val;
  int d0, d1, d2, g0, g1, g2, stride, P, f, op, batch_sample, vol;
};

// does the crop of the patch with origin o cover coordinate v along one axis?   (test_all.py:244-263 plain, :270-293 batch_sample)
__device__ __forceinline__ bool crop_covers(const StitchParams& p, int o, int v) {
  int ms = 0, me = 0;
  if (p.op >= 0) {  // op < 0 encodes overlap >= patch: no cropping (:264-265, :297-298)
    ms = o == 0 ? 0 : p.op;
    if (p.batch_sample) me = (p.vol == o + p.P || p.vol - p.P <= o) ? 0 : p.op;
    else me = (p.vol - p.P <= o + p.P) ? 0 : p.op;
  }
  return v >= o + ms && v < o + p.P - me;
}

__global__ void __launch_bounds__(256) stitch_patches_kernel(StitchParams p) {
  const int64_t total = (int64_t)p.d0 * p.d1 * p.d2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (p.lowres && p.lowres[i] == p.min_val) {
      p.pred[i] = p.min_val;
      continue;
    }
    const int v2 = (int)(i % p.d2), v1 = (int)((i / p.d2) % p.d1), v0 = (int)(i / ((int64_t)p.d2 * p.d1));
    const int v[3] = {v0, v1, v2};
    const int g[3] = {p.g0, p.g1, p.g2};
    int lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      hi[a] = min(g[a] - 1, v[a] / p.stride);
      const int t = v[a] - p.P + 1;
      lo[a] = t <= 0 ? 0 : (t + p.stride - 1) / p.stride;
    }
    bool done = false;
    // patches are applied in grid order (i outermost, k fastest): the last writer is the lexicographically largest covering patch
    for (int a = hi[0]; a >= lo[0] && !done; --a) {
      if (!crop_covers(p, a * p.stride, v0)) continue;
      for (int b = hi[1]; b >= lo[1] && !done; --b) {
        if (!crop_covers(p, b * p.stride, v1)) continue;
        for (int c = hi[2]; c >= lo[2]; --c) {
          if (!crop_covers(p, c * p.stride, v2)) continue;
          const int s = p.slot[(a * p.g1 + b) * p.g2 + c];
          if (s < 0) continue;
          const int64_t per = (int64_t)p.P * p.P * p.P;
          p.pred[i] = p.patches[(int64_t)s * per + patch_offset(v0 - a * p.stride, v1 - b * p.stride, v2 - c * p.stride, p.P, p.f)];
          done = true;
          break;
        }
      }
    }
  }
}

}  // namespace diqt

using namespace diqt;

extern "C" int diqt_gather_patches(const float* volume, int d0, int d1, int d2, const int32_t* origins, int n_patches, int patch, int sub_f,
                                   float* out, void* stream) {
  DIQT_REQUIRE(volume && origins && out && n_patches > 0 && patch > 0, "gather_patches: bad arguments");
  DIQT_REQUIRE(patch <= d0 && patch <= d1 && patch <= d2, "gather_patches: patch %d larger than the volume (%d, %d, %d)", patch, d0, d1, d2);
  DIQT_REQUIRE(sub_f <= 1 || patch % sub_f == 0, "gather_patches: patch %d is not a multiple of the sub-volume factor %d", patch, sub_f);
  const int64_t total = (int64_t)n_patches * patch * patch * patch;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_patches_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(volume, d1, d2, origins, n_patches, patch, sub_f, out);
  return check_launch("gather_patches");
}

extern "C" int diqt_stitch_patches(const float* patches, const int32_t* slot_of_grid, int g0, int g1, int g2, int stride, int patch, int overlap,
                                   int batch_sample, int sub_f, float* pred, int d0, int d1, int d2, const float* lowres, float min_val,
                                   void* stream) {
  DIQT_REQUIRE(patches && slot_of_grid && pred && g0 > 0 && g1 > 0 && g2 > 0 && stride > 0 && patch > 0, "stitch_patches: bad arguments");
  DIQT_REQUIRE((g0 - 1) * stride + patch <= d0 && (g1 - 1) * stride + patch <= d1 && (g2 - 1) * stride + patch <= d2,
               "stitch_patches: the patch grid does not fit the volume");
  DIQT_REQUIRE(sub_f <= 1 || patch % sub_f == 0, "stitch_patches: patch %d is not a multiple of the sub-volume factor %d", patch, sub_f);
  StitchParams p;
  p.patches = patches; p.slot = slot_of_grid; p.pred = pred; p.lowres = lowres; p.min_val = min_val;
  p.d0 = d0; p.d1 = d1; p.d2 = d2; p.g0 = g0; p.g1 = g1; p.g2 = g2; p.stride = stride; p.P = patch; p.f = sub_f;
  p.op = overlap < patch ? overlap / 2 : -1;
  p.batch_sample = batch_sample;
  p.vol = d2;  // the reference uses pred_ary.shape[-1] for every axis (test_all.py:249-261)
  const int64_t total = (int64_t)d0 * d1 * d2;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  stitch_patches_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("stitch_patches");
}
