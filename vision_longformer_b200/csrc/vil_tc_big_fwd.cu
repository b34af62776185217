// TU: tcgen05 forward for chunk sizes w in {12, 15, 31} (kernel: vil_tc_big.cuh).
#include "vil_tc_host.cuh"
#include "vil_tc_big.cuh"

namespace vil {
namespace tc {
namespace {

template <int DP, int W, bool BF16>
int launch_fwd_big(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  constexpr int PR = 64 / W, NP = (W + PR - 1) / PR, NPP = (NP + 1) / 2;
  FwdArgs a;
  a.geo = g;
  a.o = t4(p->o);
  a.lse = p->lse;
  a.table = p->bias_table;
  a.g2l = p->g2l;
  a.cpairs = 0;
  a.num_units = g.B * g.H * g.mx * g.my * NPP;
  a.has_tab = (p->bias_table != nullptr) || g.exact == 1;
  a.scale_log2 = g.scale * 1.4426950408889634f;
  a.out_f32 = out_f32(p) ? 1 : 0;
  CUtensorMap tmQ, tmK, tmV, tmKg, tmVg;
  int rc;
  if ((rc = local_map(&tmQ, p->q, 0, g, p->dtype, DP, PR))) return rc;
  if ((rc = local_map(&tmK, p->k, g.g, g, p->dtype, DP, PR))) return rc;
  if ((rc = local_map(&tmV, p->v, g.g, g, p->dtype, DP, PR))) return rc;
  if ((rc = token_map(&tmKg, p->k, g.N, g, p->dtype, DP, 16))) return rc;
  if ((rc = token_map(&tmVg, p->v, g.N, g, p->dtype, DP, 16))) return rc;
  const int tw = 4 * g.w - 1;
  const int tab_floats = g.H * (a.has_tab ? tw * tw : 0) + g.H * 16;
  int smem = FwdSmem<DP>::total(tab_floats) + BAR_COUNT * 8;
  if (smem < 80 * 1024) smem = 80 * 1024;
  auto kern = vil_tc_fwd_big_kernel<DP, W, BF16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  int grid = 2 * num_sms();
  if (grid > a.num_units) grid = a.num_units;
  kern<<<grid, kThreads, smem, s>>>(tmQ, tmK, tmV, tmKg, tmVg, a);
  count_launch();
  return launch_check("vil_tc_fwd_big_kernel");
}

template <int DP, bool BF16>
int dispatch_w(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  switch (g.w) {
    case 12: return launch_fwd_big<DP, 12, BF16>(p, g, s);
    case 14: return launch_fwd_big<DP, 14, BF16>(p, g, s);
    case 15: return launch_fwd_big<DP, 15, BF16>(p, g, s);
    default: return launch_fwd_big<DP, 31, BF16>(p, g, s);
  }
}

}  // namespace

int launch_fwd_local_big(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool bf = p->dtype == VIL_BF16;
  if (g.D <= 32) return bf ? dispatch_w<32, true>(p, g, s) : dispatch_w<32, false>(p, g, s);
  return bf ? dispatch_w<64, true>(p, g, s) : dispatch_w<64, false>(p, g, s);
}

}  // namespace tc
}  // namespace vil
