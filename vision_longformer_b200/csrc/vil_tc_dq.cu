// TU: tcgen05 backward pass 1 (dQ, optional bias-table gradient) and the re-ordering prologues, chunk size w <= 8.
#include "vil_tc_bwd_host.cuh"

namespace vil {
namespace tc {
namespace {

template <int DP, int W, bool BF16>
int launch_dq(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  BwdLaunch L;
  const int cpairs = (g.my + 1) / 2;
  int rc = setup_bwd<DP>(L, p, g, 0, g.B * g.H * g.mx * cpairs, false);
  if (rc) return rc;
  BwdArgs& a = L.a;
  a.out0 = t4(p->dq); a.out1 = t4(p->dq);
  const bool dbias = p->bias_table != nullptr;
  cudaError_t e;
  if (!dbias) {
    auto k1 = vil_tc_bwd_dq_kernel<DP, W, BF16, false>;
    if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem)) != cudaSuccess)
      return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
    k1<<<L.grid, kBwdThreads, L.smem, s>>>(L.tmQ, L.tmDO, L.tmK, L.tmV, L.tmKg, L.tmVg, a);
  } else {
    // head-affine persistent grid: a multiple of H CTAs, one per SM
    const int tw = 4 * g.w - 1;
    const int smem1 = L.smem_true + 9 * g.w2 * g.w2 * 4 + tw * tw * 4 + 64;
    int grid1 = (num_sms() / g.H) * g.H;
    if (grid1 > a.num_units) grid1 = ((a.num_units + g.H - 1) / g.H) * g.H;
    auto k1 = vil_tc_bwd_dq_kernel<DP, W, BF16, true>;
    if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1)) != cudaSuccess)
      return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
    k1<<<grid1, kBwdThreads, smem1, s>>>(L.tmQ, L.tmDO, L.tmK, L.tmV, L.tmKg, L.tmVg, a);
  }
  count_launch();
  return launch_check(dbias ? "vil_tc_bwd_dq_kernel<dbias>" : "vil_tc_bwd_dq_kernel");
}

template <int DP, bool BF16>
int dispatch_w(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  switch (g.w) {
    case 6: return launch_dq<DP, 6, BF16>(p, g, s);
    case 7: return launch_dq<DP, 7, BF16>(p, g, s);
    default: return launch_dq<DP, 8, BF16>(p, g, s);
  }
}

}  // namespace

// token-ordered lse / delta -> chunk-ordered 64-padded copies (+ the 16-padded global-row copies pass 2 needs)
int launch_bwd_prep(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  float* ws = static_cast<float*>(p->workspace);
  float* lse2c = ws + ws_off_tc(g);
  float* deltac = lse2c + ws_tc_floats(g) / 2;
  const long long total = ws_tc_floats(g) / 2;
  vil_tc_bwd_prep<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(g, p->lse, ws, lse2c, deltac);
  count_launch();
  int rc = launch_check("vil_tc_bwd_prep");
  if (rc) return rc;
  if (bwd_fuses_global_rows(p, g)) {
    vil_tc_bwd_prep_g<<<(g.B * g.H * 16 + 255) / 256, 256, 0, s>>>(g, p->lse_g, ws + ws_off_delta_g(g), p->g2l, ws + ws_off_tcg(g),
                                                                     ws + ws_off_tcg(g) + ws_tcg_floats(g) / 2);
    count_launch();
    rc = launch_check("vil_tc_bwd_prep_g");
  }
  return rc;
}

int launch_bwd_dq(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool bf = p->dtype == VIL_BF16;
  if (g.D <= 32) return bf ? dispatch_w<32, true>(p, g, s) : dispatch_w<32, false>(p, g, s);
  return bf ? dispatch_w<64, true>(p, g, s) : dispatch_w<64, false>(p, g, s);
}

}  // namespace tc
}  // namespace vil
