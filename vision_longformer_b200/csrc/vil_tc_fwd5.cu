// TU: fused tcgen05 forward, key-row-block variant (kernel: vil_tc_fwd5.cuh); shares Args / merge with vil_tc_fwd2.
#include <cstdlib>
#include "vil_tc_host.cuh"
#include "vil_tc_fwd5.cuh"

namespace vil {
namespace tc {

// w = 7, D <= 32, mode 0, no bias / window-mask table, no padded chunk, global rows (if any) in the spare lanes
bool fwd5_applies(const VilAttnParams* p, const Geo& g) {
  if (g.w != 7 || g.D > 32 || g.mode != 0 || p->bias_table != nullptr || g.exact != 0) return false;
  if (g.padx != 0 || g.pady != 0) return false;
  if (g.g > 0 && !fwd2_fuses_global_rows(p, g)) return false;
  return true;
}

namespace {

template <bool BF16>
int launch(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  constexpr int DP = 32;
  f2::Args a;
  a.geo = g;
  a.o = t4(p->o);
  a.lse = p->lse;
  a.table = nullptr;
  a.g2l = p->g2l;
  a.g2g = p->g2g;
  a.part = static_cast<float*>(p->workspace);
  a.cpairs = (g.my + 1) / 2;
  a.num_units = g.B * g.H * g.mx * a.cpairs;
  a.has_tab = 0;
  a.fuse_g = g.g > 0 ? 1 : 0;
  a.out_f32 = out_f32(p) ? 1 : 0;
  a.scale_log2 = g.scale * 1.4426950408889634f;
  CUtensorMap tmQ, tmQg, tmK, tmV, tmKg, tmVg;
  int rc;
  if ((rc = local_map(&tmQ, p->q, 0, g, p->dtype, DP))) return rc;
  if (a.fuse_g) { if ((rc = token_map(&tmQg, p->qg, g.g, g, p->dtype, DP, 8))) return rc; }
  else tmQg = tmQ;                                                     // never dereferenced
  if ((rc = local_map_box(&tmK, p->k, g.g, g, p->dtype, DP, 8, 3))) return rc;      // (D, 8 columns, 3 key rows) boxes
  if ((rc = local_map_box(&tmV, p->v, g.g, g, p->dtype, DP, 8, 3))) return rc;
  if ((rc = token_map(&tmKg, p->k, g.N, g, p->dtype, DP, 16))) return rc;
  if ((rc = token_map(&tmVg, p->v, g.N, g, p->dtype, DP, 16))) return rc;
  const int tab_floats = g.H * (16 + 8 + 128);
  const int smem = f5::Smem::total(tab_floats) + f5::Bars::COUNT * 8;
  auto kern = f5::vil_tc_fwd5_kernel<BF16, !BF16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (getenv("VIL_DEBUG_OCC")) {
    int nb = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, f5::kThreads5, smem);
    fprintf(stderr, "vil_tc_fwd5: smem %d B, max active CTAs / SM %d\n", smem, nb);
  }
  int grid = 4 * num_sms();                        // 128 TMEM columns per CTA: four CTAs per SM
  if (grid > a.num_units) grid = a.num_units;
  kern<<<grid, f5::kThreads5, smem, s>>>(tmQ, tmQg, tmK, tmV, tmKg, tmVg, a);
  count_launch();
  if ((rc = launch_check("vil_tc_fwd5_kernel"))) return rc;
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

}  // namespace

int launch_fwd5(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  return p->dtype == VIL_BF16 ? launch<true>(p, g, s) : launch<false>(p, g, s);
}

}  // namespace tc
}  // namespace vil
