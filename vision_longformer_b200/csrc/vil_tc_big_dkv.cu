// TU: tcgen05 backward pass 2 for chunk sizes w in {12, 15, 31}.
#include "vil_tc_bwd_host.cuh"
#include "vil_tc_big.cuh"

namespace vil {
namespace tc {
namespace {

template <int DP, int W, bool BF16>
int launch_dkv_big(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  constexpr int PR = 64 / W, NP = (W + PR - 1) / PR, NPP = (NP + 1) / 2;
  BwdLaunch L;
  int rc = setup_bwd<DP>(L, p, g, PR, g.B * g.H * g.mx * g.my * NPP, false);
  if (rc) return rc;
  L.a.out0 = t4(p->dk); L.a.out1 = t4(p->dv);
  auto k2 = vil_tc_bwd_dkv_big_kernel<DP, W, BF16>;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem)) != cudaSuccess)
    return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  k2<<<L.grid, kBwdThreads, L.smem, s>>>(L.tmQ, L.tmDO, L.tmK, L.tmV, L.a);
  count_launch();
  return launch_check("vil_tc_bwd_dkv_big_kernel");
}

template <int DP, bool BF16>
int dispatch_w(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  switch (g.w) {
    case 12: return launch_dkv_big<DP, 12, BF16>(p, g, s);
    case 14: return launch_dkv_big<DP, 14, BF16>(p, g, s);
    case 15: return launch_dkv_big<DP, 15, BF16>(p, g, s);
    default: return launch_dkv_big<DP, 31, BF16>(p, g, s);
  }
}

}  // namespace

int launch_bwd_dkv_big(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool bf = p->dtype == VIL_BF16;
  if (g.D <= 32) return bf ? dispatch_w<32, true>(p, g, s) : dispatch_w<32, false>(p, g, s);
  return bf ? dispatch_w<64, true>(p, g, s) : dispatch_w<64, false>(p, g, s);
}

}  // namespace tc
}  // namespace vil
