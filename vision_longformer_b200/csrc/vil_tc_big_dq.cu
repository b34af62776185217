// TU: tcgen05 backward pass 1 + re-ordering prologue for chunk sizes w in {12, 15, 31}.
#include "vil_tc_bwd_host.cuh"
#include "vil_tc_big.cuh"

namespace vil {
namespace tc {
namespace {

template <int DP, int W, bool BF16>
int launch_dq_big(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  constexpr int PR = 64 / W, NP = (W + PR - 1) / PR, NPP = (NP + 1) / 2;
  BwdLaunch L;
  int rc = setup_bwd<DP>(L, p, g, PR, g.B * g.H * g.mx * g.my * NPP, false);
  if (rc) return rc;
  L.a.out0 = t4(p->dq); L.a.out1 = t4(p->dq);
  auto k1 = vil_tc_bwd_dq_big_kernel<DP, W, BF16>;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem)) != cudaSuccess)
    return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  k1<<<L.grid, kBwdThreads, L.smem, s>>>(L.tmQ, L.tmDO, L.tmK, L.tmV, L.tmKg, L.tmVg, L.a);
  count_launch();
  return launch_check("vil_tc_bwd_dq_big_kernel");
}

template <int DP, bool BF16>
int dispatch_w(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  switch (g.w) {
    case 12: return launch_dq_big<DP, 12, BF16>(p, g, s);
    case 14: return launch_dq_big<DP, 14, BF16>(p, g, s);
    case 15: return launch_dq_big<DP, 15, BF16>(p, g, s);
    default: return launch_dq_big<DP, 31, BF16>(p, g, s);
  }
}

template <int W>
int prep_big(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  float* ws = static_cast<float*>(p->workspace);
  float* lse2c = ws + ws_off_tc(g);
  float* deltac = lse2c + ws_tc_floats(g) / 2;
  const long long total = ws_tc_floats(g) / 2;
  vil_tc_bwd_prep_big<W><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(g, p->lse, ws, lse2c, deltac);
  count_launch();
  return launch_check("vil_tc_bwd_prep_big");
}

}  // namespace

int launch_bwd_prep_big(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  switch (g.w) {
    case 12: return prep_big<12>(p, g, s);
    case 14: return prep_big<14>(p, g, s);
    case 15: return prep_big<15>(p, g, s);
    default: return prep_big<31>(p, g, s);
  }
}

int launch_bwd_dq_big(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool bf = p->dtype == VIL_BF16;
  if (g.D <= 32) return bf ? dispatch_w<32, true>(p, g, s) : dispatch_w<32, false>(p, g, s);
  return bf ? dispatch_w<64, true>(p, g, s) : dispatch_w<64, false>(p, g, s);
}

}  // namespace tc
}  // namespace vil
