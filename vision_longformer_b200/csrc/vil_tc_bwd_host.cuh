// Host-side argument / tensor-map set-up shared by the two backward passes (w <= 8 and big-window variants).
#pragma once
#include "vil_tc_host.cuh"
#include "vil_tc_bwd.cuh"

namespace vil {
namespace tc {

struct BwdLaunch {
  BwdArgs a;
  CUtensorMap tmQ, tmDO, tmK, tmV, tmKg, tmVg, tmQg, tmDOg;
  int smem, smem_true, grid;
};

// piece_rows = 0: w <= 8 (units = chunk-column pairs); > 0: big windows (units = piece pairs of one chunk)
template <int DP>
inline int setup_bwd(BwdLaunch& L, const VilAttnParams* p, const Geo& g, int piece_rows, int num_units, bool want_fuse_g) {
  float* ws = static_cast<float*>(p->workspace);
  float* lse2c = ws + ws_off_tc(g);
  BwdArgs& a = L.a;
  a.geo = g;
  a.table = p->bias_table; a.g2l = p->g2l;
  a.lse2c = lse2c; a.deltac = lse2c + ws_tc_floats(g) / 2;
  a.cpairs = piece_rows > 0 ? 0 : (g.my + 1) / 2;
  a.num_units = num_units;
  a.has_tab = (p->bias_table != nullptr) || g.exact == 1;
  a.scale_log2 = g.scale * 1.4426950408889634f;
  a.scale = g.scale;
  a.d_table = piece_rows > 0 ? nullptr : p->d_bias_table;
  a.out_f32 = out_f32(p) ? 1 : 0;
  a.fuse_g = (want_fuse_g && bwd_fuses_global_rows(p, g)) ? 1 : 0;
  a.lse2g = ws + ws_off_tcg(g);
  a.deltag = a.lse2g + ws_tcg_floats(g) / 2;
  int rc;
  if (a.fuse_g) {
    if ((rc = token_map(&L.tmQg, p->qg, g.g, g, p->dtype, DP, 16))) return rc;
    if ((rc = token_map(&L.tmDOg, p->d_og, g.g, g, p->dtype, DP, 16))) return rc;
  } else {
    if ((rc = token_map(&L.tmQg, p->k, g.N, g, p->dtype, DP, 16))) return rc;      // never dereferenced
    L.tmDOg = L.tmQg;
  }
  if ((rc = local_map(&L.tmQ, p->q, 0, g, p->dtype, DP, piece_rows))) return rc;
  if ((rc = local_map(&L.tmDO, p->d_o, 0, g, p->dtype, DP, piece_rows))) return rc;
  if ((rc = local_map(&L.tmK, p->k, g.g, g, p->dtype, DP, piece_rows))) return rc;
  if ((rc = local_map(&L.tmV, p->v, g.g, g, p->dtype, DP, piece_rows))) return rc;
  if ((rc = token_map(&L.tmKg, p->k, g.N, g, p->dtype, DP, 16))) return rc;
  if ((rc = token_map(&L.tmVg, p->v, g.N, g, p->dtype, DP, 16))) return rc;
  const int tw = 4 * g.w - 1;
  const int tab_floats = g.H * (a.has_tab ? tw * tw : 0) + g.H * 16;
  L.smem_true = BwdSmem<DP>::total(tab_floats) + BB_COUNT * 8;
  L.smem = L.smem_true < 80 * 1024 ? 80 * 1024 : L.smem_true;
  L.grid = 2 * num_sms();
  if (L.grid > a.num_units) L.grid = a.num_units;
  return VIL_OK;
}

}  // namespace tc
}  // namespace vil
