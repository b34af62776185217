// TU: fused tcgen05 forward (local + global query rows), chunk size w <= 8 (kernel: vil_tc_fwd2.cuh).
#include <cstdlib>
#include "vil_tc_host.cuh"
#include "vil_tc_fwd2.cuh"

namespace vil {
namespace tc {

// the global query rows can ride in the spare lanes of slot A (see vil_tc_fwd2.cuh)
bool fwd2_fuses_global_rows(const VilAttnParams* p, const Geo& g) {
  if (g.g == 0 || g.g > f2::kGMax || g.w2 > f2::kGRow0 || g.mode != 0) return false;
  const bool shared = (p->kg.ptr == p->k.ptr) && (p->vg.ptr == p->v.ptr) && p->kg.sb == p->k.sb && p->kg.sh == p->k.sh &&
                      p->kg.st == p->k.st && p->vg.sb == p->v.sb && p->vg.sh == p->v.sh && p->vg.st == p->v.st;
  return shared && aligned16(p->qg, 2);
}

long long fwd2_workspace_floats(const VilAttnParams* p, const Geo& g) {
  if (!fwd2_fuses_global_rows(p, g)) return 0;
  const int DP = g.D <= 32 ? 32 : 64;
  return (long long)g.B * g.H * g.mx * ((g.my + 1) / 2) * f2::kGMax * (DP + 2);
}

namespace {

constexpr bool kP16 = true;             // default operand format of P (see vil_tc_fwd2.cuh)

int p16_knob() {                        // tuning / bring-up aid: VIL_FWD2_P16 = 0 | 1 (w = 7, no table only)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VIL_FWD2_P16");
    v = e ? (atoi(e) != 0) : (kP16 ? 1 : 0);
  }
  return v;
}

int poly_knob() {                       // tuning aid: VIL_FWD2_POLY = 0 | 2 | 4 (w = 7 only); default set below
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VIL_FWD2_POLY");
    v = e ? atoi(e) : 0;
    if (v != 0 && v != 2 && v != 4) v = 0;
  }
  return v;
}

template <int DP, int W, bool BF16, bool HAS_TAB, int POLY, bool P16>
int launch(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  f2::Args a;
  a.geo = g;
  a.o = t4(p->o);
  a.lse = p->lse;
  a.table = p->bias_table;
  a.g2l = p->g2l;
  a.g2g = p->g2g;
  a.part = static_cast<float*>(p->workspace);
  a.cpairs = (g.my + 1) / 2;
  a.num_units = g.B * g.H * g.mx * a.cpairs;
  a.has_tab = HAS_TAB ? 1 : 0;
  a.fuse_g = fwd2_fuses_global_rows(p, g) ? 1 : 0;
  a.out_f32 = out_f32(p) ? 1 : 0;
  a.scale_log2 = g.scale * 1.4426950408889634f;
  CUtensorMap tmQ, tmQg, tmK, tmV, tmKg, tmVg;
  int rc;
  if ((rc = local_map(&tmQ, p->q, 0, g, p->dtype, DP))) return rc;
  if (a.fuse_g) { if ((rc = token_map(&tmQg, p->qg, g.g, g, p->dtype, DP, 8))) return rc; }
  else tmQg = tmQ;                                                     // never dereferenced
  if ((rc = local_map(&tmK, p->k, g.g, g, p->dtype, DP))) return rc;
  if ((rc = local_map(&tmV, p->v, g.g, g, p->dtype, DP))) return rc;
  if ((rc = token_map(&tmKg, p->k, g.N, g, p->dtype, DP, 16))) return rc;
  if ((rc = token_map(&tmVg, p->v, g.N, g, p->dtype, DP, 16))) return rc;
  const int tw = 4 * g.w - 1;
  const int tab_floats = (HAS_TAB ? g.H * tw * tw + (g.w - 1) * tw + g.w : 0) + g.H * (16 + 8 + 128);
  int smem = f2::Smem<DP>::total(tab_floats) + f2::Bars<DP>::COUNT * 8;
  if (smem < 80 * 1024) smem = 80 * 1024;          // caps residency at 2 CTAs / SM (2 x 256 TMEM columns)
  auto kern = f2::vil_tc_fwd2_kernel<DP, W, BF16, HAS_TAB, POLY, P16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  int grid = 2 * num_sms();
  if (grid > a.num_units) grid = a.num_units;
  kern<<<grid, f2::kThreads2, smem, s>>>(tmQ, tmQg, tmK, tmV, tmKg, tmVg, a);
  count_launch();
  if ((rc = launch_check("vil_tc_fwd2_kernel"))) return rc;
  if (a.fuse_g && !(p->skip_mask & 1)) {
    const int warps = g.B * g.H * g.g;
    const int units_per_bh = g.mx * a.cpairs;
    if (out_f32(p)) f2::vil_tc_fwd2_merge<float><<<(warps * 32 + 255) / 256, 256, 0, s>>>(g, a.part, units_per_bh, DP, t4(p->og), p->lse_g);
    else if (BF16)  f2::vil_tc_fwd2_merge<__nv_bfloat16><<<(warps * 32 + 255) / 256, 256, 0, s>>>(g, a.part, units_per_bh, DP, t4(p->og), p->lse_g);
    else            f2::vil_tc_fwd2_merge<__half><<<(warps * 32 + 255) / 256, 256, 0, s>>>(g, a.part, units_per_bh, DP, t4(p->og), p->lse_g);
    count_launch();
    rc = launch_check("vil_tc_fwd2_merge");
  }
  return rc;
}

template <int DP, int W, bool BF16>
int dispatch_tab(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool has_tab = (p->bias_table != nullptr) || g.exact == 1;
  if (has_tab) return launch<DP, W, BF16, true, 0, kP16>(p, g, s);
  if constexpr (W == 7) {
    if (poly_knob() == 2) return launch<DP, W, BF16, false, 2, kP16>(p, g, s);
    if (poly_knob() == 4) return launch<DP, W, BF16, false, 4, kP16>(p, g, s);
    if constexpr (BF16) { if (p16_knob() != (kP16 ? 1 : 0)) return launch<DP, W, BF16, false, 0, !kP16>(p, g, s); }
  }
  return launch<DP, W, BF16, false, 0, kP16>(p, g, s);
}

template <int DP, bool BF16>
int dispatch_w(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  switch (g.w) {
    case 6: return dispatch_tab<DP, 6, BF16>(p, g, s);
    case 7: return dispatch_tab<DP, 7, BF16>(p, g, s);
    default: return dispatch_tab<DP, 8, BF16>(p, g, s);
  }
}

}  // namespace

int launch_fwd2(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool bf = p->dtype == VIL_BF16;
  if (g.D <= 32) return bf ? dispatch_w<32, true>(p, g, s) : dispatch_w<32, false>(p, g, s);
  return bf ? dispatch_w<64, true>(p, g, s) : dispatch_w<64, false>(p, g, s);
}

}  // namespace tc
}  // namespace vil
