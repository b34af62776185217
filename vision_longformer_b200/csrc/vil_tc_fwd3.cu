// TU: fused tcgen05 forward, 4-CTAs-per-SM variant (kernel: vil_tc_fwd3.cuh); shares Args / merge with vil_tc_fwd2.
#include <cstdlib>
#include "vil_tc_host.cuh"
#include "vil_tc_fwd3.cuh"

namespace vil {
namespace tc {
namespace {

template <int DP, int W, bool BF16, bool HAS_TAB, bool LEAN>
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
  using SM = f3::Smem<DP, HAS_TAB>;
  const int smem = SM::total(tab_floats) + f3::Bars<DP, HAS_TAB>::COUNT * 8;
  auto kern = f3::vil_tc_fwd3_kernel<DP, W, BF16, HAS_TAB, !BF16, LEAN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  int grid = 4 * num_sms();                        // 128 TMEM columns per CTA: four CTAs per SM
  if (grid > a.num_units) grid = a.num_units;
  kern<<<grid, f3::kThreads3, smem, s>>>(tmQ, tmQg, tmK, tmV, tmKg, tmVg, a);
  count_launch();
  if ((rc = launch_check("vil_tc_fwd3_kernel"))) return rc;
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
  const bool lean = g.padx == 0 && g.pady == 0;
  if (has_tab) return lean ? launch<DP, W, BF16, true, true>(p, g, s) : launch<DP, W, BF16, true, false>(p, g, s);
  return lean ? launch<DP, W, BF16, false, true>(p, g, s) : launch<DP, W, BF16, false, false>(p, g, s);
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

int launch_fwd3(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool bf = p->dtype == VIL_BF16;
  if (g.D <= 32) return bf ? dispatch_w<32, true>(p, g, s) : dispatch_w<32, false>(p, g, s);
  return bf ? dispatch_w<64, true>(p, g, s) : dispatch_w<64, false>(p, g, s);
}

}  // namespace tc
}  // namespace vil
