// C-ABI entry points for the convolutions: choose a kernel family, pack weights, own the plan.
#include <new>

#include "common.cuh"

namespace diqt {
struct TcPlan;
int simt_taps(int mode);
int conv_simt_pack(const diqt_conv_desc* d, const float* w, void* packed, cudaStream_t st);
int conv_simt_run(const diqt_conv_desc* d, const void* in, void* out, const void* packed, const float* bias, cudaStream_t st);
int conv_bias_permute_up(const float* b, int c_out, float* out, cudaStream_t st);
bool conv_tc_supported(const diqt_conv_desc* d);
size_t conv_tc_packed_bytes(const diqt_conv_desc* d);
int conv_tc_pack(const diqt_conv_desc* d, const float* w, void* packed, cudaStream_t st);
int conv_tc_plan(const diqt_conv_desc* d, const void* in, void* out, const void* packed, const float* bias, TcPlan** plan);
int conv_tc_run(const TcPlan* plan, cudaStream_t st);
void conv_tc_destroy(TcPlan* plan);
int conv_tc_set_stats(TcPlan* plan, float* partial, float* group, unsigned int* tickets, int* ngroups);
bool conv_tc_prefers_small(const diqt_conv_desc* d);
size_t conv_tc_workspace_bytes(const TcPlan* plan);
int conv_tc_set_workspace(TcPlan* plan, void* ws, size_t bytes);
struct ZmPlan;
bool conv_zm_supported(const diqt_conv_desc* d);
bool conv_zm_profitable(const diqt_conv_desc* d);
size_t conv_zm_packed_bytes(const diqt_conv_desc* d);
int conv_zm_pack(const diqt_conv_desc* d, const float* w, void* packed, cudaStream_t st);
int conv_zm_plan(const diqt_conv_desc* d, const void* in, void* out, const void* packed, const float* bias, ZmPlan** plan);
int conv_zm_run(const ZmPlan* plan, cudaStream_t st);
int conv_zm_set_stats(ZmPlan* plan, float* partial, float* group, unsigned int* tickets, int* ngroups);
void conv_zm_destroy(ZmPlan* plan);
bool conv_zm_gn_supported(const diqt_conv_desc* d);
int conv_zm_set_gn(ZmPlan* plan, const GnParams& gn);
int conv_zm_set_gn_affine(ZmPlan* plan, const float* a, const float* b);
void conv_zm_set_film(ZmPlan* plan, const float* film, int film_ld, const int* film_row, int film_row_stride_n);
}  // namespace diqt

using namespace diqt;

struct diqt_conv_plan {
  diqt_conv_desc d;
  int impl;
  const void* in;
  void* out;
  const void* packed;
  const float* bias;
  TcPlan* tc;
  ZmPlan* zm;
};

static int check_desc(const diqt_conv_desc* d) {
  DIQT_REQUIRE(d, "conv: null descriptor");
  DIQT_REQUIRE(d->mode >= DIQT_CONV_K3 && d->mode <= DIQT_CONV_UP, "conv: bad mode %d", d->mode);
  DIQT_REQUIRE(d->dtype == DIQT_F32 || d->dtype == DIQT_BF16, "conv: bad dtype %d", d->dtype);
  DIQT_REQUIRE(d->n > 0 && d->d0 > 0 && d->d1 > 0 && d->d2 > 0 && d->c_in > 0 && d->c_out > 0, "conv: non-positive dimension");
  DIQT_REQUIRE(d->ld_in >= d->c_in, "conv: ld_in=%d < c_in=%d", d->ld_in, d->c_in);
  if (d->mode == DIQT_CONV_UP) {
    DIQT_REQUIRE(d->c_out % 8 == 0 && d->ld_out >= d->c_out / 8, "conv(up): c_out=%d ld_out=%d", d->c_out, d->ld_out);
  } else {
    DIQT_REQUIRE(d->ld_out >= d->c_out, "conv: ld_out=%d < c_out=%d", d->ld_out, d->c_out);
  }
  return DIQT_OK;
}

static int resolve_impl(const diqt_conv_desc* d, int* impl) {
  int want = d->impl;
  if (want == DIQT_IMPL_AUTO)
    want = (conv_zm_supported(d) && conv_zm_profitable(d) && !conv_tc_prefers_small(d)) ? DIQT_IMPL_ZM
           : conv_tc_supported(d)                                                          ? DIQT_IMPL_TC
                                                                                           : DIQT_IMPL_SIMT;
  if (want == DIQT_IMPL_ZM && !conv_zm_supported(d)) {
    set_error("conv: z-march kernel needs 3x3x3 bf16, c_in %% 64 == 0, c_out %% 64 == 0 (<= 256), d2 %% 8 == 0, d1 %% 16 == 0 (got c_in=%d c_out=%d dims %d,%d,%d)", d->c_in,
              d->c_out, d->d0, d->d1, d->d2);
    return DIQT_EUNSUPPORTED;
  }
  if (want == DIQT_IMPL_TC && !conv_tc_supported(d)) {
    set_error("conv: tcgen05 kernel does not take dtype=%d c_in=%d c_out=%d ld_in=%d ld_out=%d", d->dtype, d->c_in, d->c_out,
              d->ld_in, d->ld_out);
    return DIQT_EUNSUPPORTED;
  }
  if (want != DIQT_IMPL_TC && want != DIQT_IMPL_SIMT && want != DIQT_IMPL_ZM) {
    set_error("conv: bad impl %d", d->impl);
    return DIQT_EINVAL;
  }
  *impl = want;
  return DIQT_OK;
}

extern "C" int diqt_conv_resolved_impl(const diqt_conv_desc* d, int* impl) {
  int rc = check_desc(d);
  if (rc) return rc;
  return resolve_impl(d, impl);
}

extern "C" int diqt_conv_packed_bytes(const diqt_conv_desc* d, size_t* bytes) {
  int rc = check_desc(d), impl = 0;
  if (rc) return rc;
  if ((rc = resolve_impl(d, &impl))) return rc;
  DIQT_REQUIRE(bytes, "conv_packed_bytes: null output");
  *bytes = impl == DIQT_IMPL_ZM   ? conv_zm_packed_bytes(d)
           : impl == DIQT_IMPL_TC ? conv_tc_packed_bytes(d)
                                  : (size_t)simt_taps(d->mode) * d->c_in * d->c_out * sizeof(float);
  return DIQT_OK;
}

extern "C" int diqt_conv_pack(const diqt_conv_desc* d, const float* w, const float* bias, void* packed_w, float* packed_bias,
                              void* stream) {
  int rc = check_desc(d), impl = 0;
  if (rc) return rc;
  if ((rc = resolve_impl(d, &impl))) return rc;
  DIQT_REQUIRE(w && packed_w && packed_bias, "conv_pack: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  rc = impl == DIQT_IMPL_ZM ? conv_zm_pack(d, w, packed_w, st) : impl == DIQT_IMPL_TC ? conv_tc_pack(d, w, packed_w, st) : conv_simt_pack(d, w, packed_w, st);
  if (rc) return rc;
  if (!bias) {
    DIQT_CUDA(cudaMemsetAsync(packed_bias, 0, sizeof(float) * d->c_out, st));
    return DIQT_OK;
  }
  if (d->mode == DIQT_CONV_UP) return conv_bias_permute_up(bias, d->c_out, packed_bias, st);
  DIQT_CUDA(cudaMemcpyAsync(packed_bias, bias, sizeof(float) * d->c_out, cudaMemcpyDeviceToDevice, st));
  return DIQT_OK;
}

extern "C" int diqt_conv_plan_create(const diqt_conv_desc* d, const void* in, void* out, const void* packed_w,
                                     const float* packed_bias, diqt_conv_plan** plan) {
  int rc = check_desc(d), impl = 0;
  if (rc) return rc;
  if ((rc = resolve_impl(d, &impl))) return rc;
  DIQT_REQUIRE(in && out && packed_w && packed_bias && plan, "conv_plan_create: null pointer");
  diqt_conv_plan* pl = new (std::nothrow) diqt_conv_plan();
  DIQT_REQUIRE(pl, "conv_plan_create: out of host memory");
  pl->d = *d;
  pl->impl = impl;
  pl->in = in;
  pl->out = out;
  pl->packed = packed_w;
  pl->bias = packed_bias;
  pl->tc = nullptr;
  pl->zm = nullptr;
  if (impl == DIQT_IMPL_TC) rc = conv_tc_plan(d, in, out, packed_w, packed_bias, &pl->tc);
  if (impl == DIQT_IMPL_ZM) rc = conv_zm_plan(d, in, out, packed_w, packed_bias, &pl->zm);
  if (rc) {
    delete pl;
    return rc;
  }
  *plan = pl;
  return DIQT_OK;
}

extern "C" void diqt_conv_plan_destroy(diqt_conv_plan* plan) {
  if (!plan) return;
  if (plan->tc) conv_tc_destroy(plan->tc);
  if (plan->zm) conv_zm_destroy(plan->zm);
  delete plan;
}

extern "C" int diqt_conv_run(const diqt_conv_plan* plan, void* stream) {
  DIQT_REQUIRE(plan, "conv_run: null plan");
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->impl == DIQT_IMPL_ZM) return conv_zm_run(plan->zm, st);
  if (plan->impl == DIQT_IMPL_TC) return conv_tc_run(plan->tc, st);
  return conv_simt_run(&plan->d, plan->in, plan->out, plan->packed, plan->bias, st);
}

extern "C" int diqt_conv_plan_set_stats(diqt_conv_plan* plan, float* partial, int* nblk) {
  DIQT_REQUIRE(plan && partial && nblk, "conv_plan_set_stats: null pointer");
  *nblk = plan->impl == DIQT_IMPL_ZM   ? conv_zm_set_stats(plan->zm, partial, nullptr, nullptr, nullptr)
          : plan->impl == DIQT_IMPL_TC ? conv_tc_set_stats(plan->tc, partial, nullptr, nullptr, nullptr)
                                       : 0;
  return DIQT_OK;
}

extern "C" int diqt_conv_plan_set_stats_g(diqt_conv_plan* plan, float* partial, float* group, uint32_t* tickets, int* nblk, int* ngroups) {
  DIQT_REQUIRE(plan && partial && group && tickets && nblk && ngroups, "conv_plan_set_stats_g: null pointer");
  *nblk = plan->impl == DIQT_IMPL_ZM   ? conv_zm_set_stats(plan->zm, partial, group, tickets, ngroups)
          : plan->impl == DIQT_IMPL_TC ? conv_tc_set_stats(plan->tc, partial, group, tickets, ngroups)
                                       : 0;
  if (*nblk == 0) *ngroups = 0;
  return DIQT_OK;
}

// ---- GroupNorm (+FiLM) + Mish of the conv INPUT, fused into the conv's load path (z-march family) ------------------------------
extern "C" int diqt_conv_gn_fusable(const diqt_conv_desc* d) {
  int impl = 0;
  if (check_desc(d) || resolve_impl(d, &impl)) return 0;
  return impl == DIQT_IMPL_ZM && conv_zm_gn_supported(d) ? 1 : 0;
}

extern "C" int diqt_conv_plan_set_gn(diqt_conv_plan* plan, const float* group, int ngroups, int64_t voxels, int groups, float eps,
                                     const float* gamma, const float* beta) {
  DIQT_REQUIRE(plan && group && gamma && beta, "conv_plan_set_gn: null pointer");
  DIQT_REQUIRE(plan->impl == DIQT_IMPL_ZM && plan->zm, "conv_plan_set_gn: only the z-march family fuses the input GroupNorm (ask diqt_conv_gn_fusable first)");
  GnParams gn = {group, ngroups, (long long)voxels, plan->d.c_in, groups, eps, gamma, beta, nullptr, 0, nullptr, 0};
  return conv_zm_set_gn(plan->zm, gn);
}

extern "C" int diqt_conv_plan_set_gn_affine(diqt_conv_plan* plan, const float* a, const float* b) {
  DIQT_REQUIRE(plan && plan->impl == DIQT_IMPL_ZM && plan->zm, "conv_plan_set_gn_affine: only the z-march family fuses the input GroupNorm");
  return conv_zm_set_gn_affine(plan->zm, a, b);
}

extern "C" int diqt_conv_plan_set_film(diqt_conv_plan* plan, const float* film, int film_ld, const int32_t* film_row, int film_row_stride_n) {
  DIQT_REQUIRE(plan && plan->impl == DIQT_IMPL_ZM && plan->zm, "conv_plan_set_film: plan has no fused GroupNorm");
  // (with diqt_conv_plan_set_gn_affine the FiLM rows are already folded into a, b by diqt_gn_finalize)
  conv_zm_set_film(plan->zm, film, film_ld, film_row, film_row_stride_n);
  return DIQT_OK;
}

// ---- split-K workspace of the per-tap family (small volumes) --------------------------------------------------------------------
extern "C" int diqt_conv_plan_workspace_bytes(const diqt_conv_plan* plan, size_t* bytes) {
  DIQT_REQUIRE(plan && bytes, "conv_plan_workspace_bytes: null pointer");
  *bytes = (plan->impl == DIQT_IMPL_TC && plan->tc) ? conv_tc_workspace_bytes(plan->tc) : 0;
  return DIQT_OK;
}

extern "C" int diqt_conv_plan_set_workspace(diqt_conv_plan* plan, void* workspace, size_t bytes) {
  DIQT_REQUIRE(plan && plan->impl == DIQT_IMPL_TC && plan->tc, "conv_plan_set_workspace: this plan takes no workspace");
  return conv_tc_set_workspace(plan->tc, workspace, bytes);
}
