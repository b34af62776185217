// TU: fused tcgen05 backward (delta / re-ordering / global rows folded into the two passes + one merge kernel),
// chunk size w <= 8, no bias-table gradient (kernels: vil_tc_bwd2.cuh).
#include "vil_tc_host.cuh"
#include "vil_tc_bwd2.cuh"

namespace vil {
namespace tc {

// the fused backward applies: no bias parameters (rpe off), mode handled by the walks, everything TMA-addressable
bool bwd2_applies(const VilAttnParams* p, const Geo& g) {
  if (is_big_w(g.w) || g.w > 8) return false;
  if (p->bias_table != nullptr || p->g2l != nullptr || p->g2g != nullptr) return false;
  if (g.g > 0 && !bwd_fuses_global_rows(p, g)) return false;      // pass 2 takes the global query rows as a 16-column block
  if (g.g > 0 && !aligned16(p->og, out_f32(p) ? 4 : 2)) return false;
  if (!aligned16(p->o, out_f32(p) ? 4 : 2)) return false;
  return true;
}
// ... and the global rows ride in the spare lanes (else the SIMT global-token kernels finish the job)
bool bwd2_fuses_spare_rows(const VilAttnParams* p, const Geo& g) {
  return g.g > 0 && g.g <= b2::kGMax && g.w2 <= b2::kGRow0 && g.mode == 0;
}
long long bwd2_workspace_floats(const VilAttnParams* p, const Geo& g) {
  if (!bwd2_applies(p, g) || !bwd2_fuses_spare_rows(p, g)) return 0;
  const int DP = g.D <= 32 ? 32 : 64;
  return (long long)g.B * g.H * g.mx * ((g.my + 1) / 2) * b2::kGMax * 3 * DP;       // pass 1: DP, pass 2: 2 DP per row
}

namespace {

struct Launch {
  b2::Args a;
  CUtensorMap tmQ, tmDO, tmK, tmV, tmKg, tmVg, tmQg, tmDOg, tmKg8, tmVg8, tmQg8, tmDOg8;
  int smem, grid;
};

template <int DP>
int setup(Launch& L, const VilAttnParams* p, const Geo& g) {
  float* ws = static_cast<float*>(p->workspace);
  b2::Args& a = L.a;
  a.geo = g;
  a.o = t4(p->o); a.d_o = t4(p->d_o); a.og = t4(p->og); a.d_og = t4(p->d_og);
  a.lse = p->lse; a.lse_g = p->lse_g;
  a.table = nullptr;
  a.lse2c = ws + ws_off_tc(g);
  a.deltac = a.lse2c + ws_tc_floats(g) / 2;
  a.lse2g = ws + ws_off_tcg(g);
  a.deltag = a.lse2g + ws_tcg_floats(g) / 2;
  a.part = a.deltag + ws_tcg_floats(g) / 2;
  a.cpairs = (g.my + 1) / 2;
  a.num_units = g.B * g.H * g.mx * a.cpairs;
  a.has_tab = g.exact == 1 ? 1 : 0;
  a.fuse_q = bwd2_fuses_spare_rows(p, g) ? 1 : 0;
  a.fuse_g = g.g > 0 ? 1 : 0;
  a.out_f32 = out_f32(p) ? 1 : 0;
  a.scale_log2 = g.scale * 1.4426950408889634f;
  a.scale = g.scale;
  int rc;
  if ((rc = local_map(&L.tmQ, p->q, 0, g, p->dtype, DP))) return rc;
  if ((rc = local_map(&L.tmDO, p->d_o, 0, g, p->dtype, DP))) return rc;
  if ((rc = local_map(&L.tmK, p->k, g.g, g, p->dtype, DP))) return rc;
  if ((rc = local_map(&L.tmV, p->v, g.g, g, p->dtype, DP))) return rc;
  if ((rc = token_map(&L.tmKg, p->k, g.N, g, p->dtype, DP, 16))) return rc;
  if ((rc = token_map(&L.tmVg, p->v, g.N, g, p->dtype, DP, 16))) return rc;
  if (g.g > 0) {
    if ((rc = token_map(&L.tmQg, p->qg, g.g, g, p->dtype, DP, 16))) return rc;
    if ((rc = token_map(&L.tmDOg, p->d_og, g.g, g, p->dtype, DP, 16))) return rc;
    // 8-row boxes for the spare lanes 56..63; the K / V ones stop at token nglo so that no local key rides along
    if ((rc = token_map(&L.tmQg8, p->qg, g.g, g, p->dtype, DP, 8))) return rc;
    if ((rc = token_map(&L.tmDOg8, p->d_og, g.g, g, p->dtype, DP, 8))) return rc;
    if ((rc = token_map(&L.tmKg8, p->k, g.g, g, p->dtype, DP, 8))) return rc;
    if ((rc = token_map(&L.tmVg8, p->v, g.g, g, p->dtype, DP, 8))) return rc;
  } else {
    L.tmQg = L.tmDOg = L.tmQg8 = L.tmDOg8 = L.tmKg8 = L.tmVg8 = L.tmKg;      // never dereferenced
  }
  const int tw = 4 * g.w - 1;
  const int tab_floats = (a.has_tab ? g.H * tw * tw + (g.w - 1) * tw + g.w : 0) + 256 + 64 + 4;
  L.smem = b2::Smem2<DP>::total(tab_floats) + b2::BB_COUNT * 8;
  if (L.smem < 80 * 1024) L.smem = 80 * 1024;
  L.grid = 2 * num_sms();
  if (L.grid > a.num_units) L.grid = a.num_units;
  return VIL_OK;
}

template <int DP, int W, bool BF16, typename TO, bool LEAN>
int launch_dq(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  Launch L;
  int rc = setup<DP>(L, p, g);
  if (rc) return rc;
  L.a.out0 = t4(p->dq); L.a.out1 = t4(p->dq);
  auto k1 = b2::vil_tc_bwd2_dq_kernel<DP, W, BF16, TO, LEAN>;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem)) != cudaSuccess)
    return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  k1<<<L.grid, kBwdThreads, L.smem, s>>>(L.tmQ, L.tmDO, L.tmK, L.tmV, L.tmKg, L.tmVg, L.tmQg8, L.tmDOg8, L.a);
  count_launch();
  return launch_check("vil_tc_bwd2_dq_kernel");
}

template <int DP, int W, bool BF16, bool LEAN>
int launch_dkv(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  Launch L;
  int rc = setup<DP>(L, p, g);
  if (rc) return rc;
  L.a.out0 = t4(p->dk); L.a.out1 = t4(p->dv);
  L.a.part += (long long)g.B * g.H * g.mx * L.a.cpairs * b2::kGMax * DP;          // behind the pass-1 partials
  auto k2 = b2::vil_tc_bwd2_dkv_kernel<DP, W, BF16, LEAN>;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem)) != cudaSuccess)
    return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  k2<<<L.grid, kBwdThreads, L.smem, s>>>(L.tmQ, L.tmDO, L.tmK, L.tmV, L.tmQg, L.tmDOg, L.tmKg8, L.tmVg8, L.a);
  count_launch();
  if ((rc = launch_check("vil_tc_bwd2_dkv_kernel"))) return rc;
  return VIL_OK;
}

template <typename TO>
int launch_merge(const VilAttnParams* p, const Geo& g, cudaStream_t s, int DP) {
  float* ws = static_cast<float*>(p->workspace);
  float* part1 = ws + ws_off_tcg(g) + ws_tcg_floats(g);
  const int cpairs = (g.my + 1) / 2, upb = g.mx * cpairs;
  float* part2 = part1 + (long long)g.B * g.H * upb * b2::kGMax * DP;
  const int n = g.B * g.H * g.g * 3 * DP;
  b2::vil_tc_bwd2_merge<TO><<<(n + 255) / 256, 256, 0, s>>>(g, part1, part2, upb, DP, g.scale, (p->skip_mask & 2) ? 0 : 1,
                                                           (p->skip_mask & 4) ? 0 : 1, t4(p->dqg), t4(p->dk), t4(p->dv));
  count_launch();
  return launch_check("vil_tc_bwd2_merge");
}

inline bool lean_geo(const Geo& g) { return g.padx == 0 && g.pady == 0 && g.exact != 1; }

template <int DP, int W, bool BF16>
int dispatch_dq_w(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  using TE = typename std::conditional<BF16, __nv_bfloat16, __half>::type;
  if (out_f32(p)) return launch_dq<DP, W, BF16, float, false>(p, g, s);            // parity build: generic variant only
  return lean_geo(g) ? launch_dq<DP, W, BF16, TE, true>(p, g, s) : launch_dq<DP, W, BF16, TE, false>(p, g, s);
}
template <int DP, bool BF16>
int dispatch_dq(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  switch (g.w) {
    case 6: return dispatch_dq_w<DP, 6, BF16>(p, g, s);
    case 7: return dispatch_dq_w<DP, 7, BF16>(p, g, s);
    default: return dispatch_dq_w<DP, 8, BF16>(p, g, s);
  }
}
template <int DP, bool BF16>
int dispatch_dkv(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool lean = lean_geo(g);
  switch (g.w) {
    case 6: return lean ? launch_dkv<DP, 6, BF16, true>(p, g, s) : launch_dkv<DP, 6, BF16, false>(p, g, s);
    case 7: return lean ? launch_dkv<DP, 7, BF16, true>(p, g, s) : launch_dkv<DP, 7, BF16, false>(p, g, s);
    default: return lean ? launch_dkv<DP, 8, BF16, true>(p, g, s) : launch_dkv<DP, 8, BF16, false>(p, g, s);
  }
}

}  // namespace

int launch_bwd2_dq(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool bf = p->dtype == VIL_BF16;
  if (g.D <= 32) return bf ? dispatch_dq<32, true>(p, g, s) : dispatch_dq<32, false>(p, g, s);
  return bf ? dispatch_dq<64, true>(p, g, s) : dispatch_dq<64, false>(p, g, s);
}
int launch_bwd2_dkv(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool bf = p->dtype == VIL_BF16;
  if (g.D <= 32) return bf ? dispatch_dkv<32, true>(p, g, s) : dispatch_dkv<32, false>(p, g, s);
  return bf ? dispatch_dkv<64, true>(p, g, s) : dispatch_dkv<64, false>(p, g, s);
}
int launch_bwd2_merge(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const int DP = g.D <= 32 ? 32 : 64;
  if (out_f32(p)) return launch_merge<float>(p, g, s, DP);
  return p->dtype == VIL_BF16 ? launch_merge<__nv_bfloat16>(p, g, s, DP) : launch_merge<__half>(p, g, s, DP);
}

}  // namespace tc
}  // namespace vil
