// TU (host only): coverage test and forward / backward orchestration of the tcgen05 / TMA kernel family.
#include <cstdlib>
#include "vil_tc_host.cuh"
#include "vil_tc_fwd.cuh"      // FwdSmem / BwdSmem sizes for the coverage test
#include "vil_tc_bwd.cuh"

namespace vil {

namespace tc {
static const char* why_not(const VilAttnParams* p, const Geo& g, bool bwd) {
  if (bwd && p->bias_table != nullptr) {
    // bias-gradient variant of pass 1: E[9][w^2][w^2] fp32 must fit next to the operand tiles (1 CTA / SM)
    const int twp = 4 * g.w - 1;
    const long long base = g.D <= 32 ? BwdSmem<32>::total(g.H * (twp * twp + 16)) : BwdSmem<64>::total(g.H * (twp * twp + 16));
    const long long need = base + BB_COUNT * 8 + 9LL * g.w2 * g.w2 * 4 + twp * twp * 4 + 64;
    if (need > 227 * 1024) return "bias-table gradient accumulator does not fit in shared memory for this (w, D)";
    if (g.H > num_sms()) return "more heads than SMs";
  }
  if (bwd && g.D > 64) return "head dim > 64";
  if (p->dtype != VIL_BF16 && p->dtype != VIL_F16) return "dtype is fp32 (tcgen05 kind::f16 needs bf16/fp16 operands)";
  const bool big_w = is_big_w(g.w);
  if (!(g.w >= 6 && g.w <= 8) && !big_w) return "chunk size w outside {6,7,8,12,14,15,31}";
  if (big_w && bwd && p->bias_table != nullptr) return "bias-table gradient for w > 8 is served by the SIMT backward";
  if (g.D % 8 != 0 || g.D > 64) return "head dim must be a multiple of 8 and <= 64";
  if (g.exact == -1) return "cyclic chunks (exact=-1)";
  if (g.g > 16) return "more than 16 global tokens";
  const int tw = 4 * g.w - 1;
  if ((p->bias_table != nullptr || g.exact == 1) && (long long)g.H * tw * tw * 4 > 48 * 1024)
    return "bias / window-mask tables of all heads exceed the shared-memory budget";
  const int oes = out_f32(p) ? 4 : 2;
  if (!aligned16(p->q, 2) || !aligned16(p->k, 2) || !aligned16(p->v, 2) || !aligned16(p->o, oes))
    return "q/k/v/o base pointers or strides are not 16-byte aligned";
  if (bwd && (!aligned16(p->d_o, 2) || !aligned16(p->dq, oes) || !aligned16(p->dk, oes) || !aligned16(p->dv, oes)))
    return "d_o/dq/dk/dv base pointers or strides are not 16-byte aligned";
  return nullptr;
}
}  // namespace tc

const char* tc_why_not(const VilAttnParams* p, const Geo& g, bool bwd) {
  const char* w = tc::why_not(p, g, bwd);
  return w ? w : "supported";
}
int tc_supported(const VilAttnParams* p, const Geo& g, bool bwd) { return tc::why_not(p, g, bwd) == nullptr; }
// fused forward variant: 5 = key-row blocks of all four chunk columns (default where it applies: w = 7, D <= 32, mode 0, no
// table, no padded chunk); 3 = chunk blocks, four CTAs per SM (default elsewhere); 2 = two CTAs per SM, 3-deep S ring with
// paired exclusive chunks.  VIL_FWD_VARIANT overrides (A/B timing).
static int fwd_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VIL_FWD_VARIANT"); const int x = e ? atoi(e) : 5; v = (x == 2 || x == 3) ? x : 5; }
  return v;
}
static int launch_fused_fwd(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const int v = fwd_variant();
  if (v == 5 && tc::fwd5_applies(p, g)) { note_kernel("fwd5"); return tc::launch_fwd5(p, g, s); }
  note_kernel(v == 2 ? "fwd2" : "fwd3");
  return v == 2 ? tc::launch_fwd2(p, g, s) : tc::launch_fwd3(p, g, s);
}

static bool use_fused(const VilAttnParams* p, const Geo& g) { return !tc::is_big_w(g.w) && !(p->flags & VIL_FLAG_UNFUSED); }

long long tc_workspace_bytes(const VilAttnParams* p, const Geo& g, bool bwd) {
  if (bwd) return (ws_tc_floats(g) + ws_tcg_floats(g) + (use_fused(p, g) ? tc::bwd2_workspace_floats(p, g) : 0)) * 4;
  return use_fused(p, g) ? tc::fwd2_workspace_floats(p, g) * 4 : 0;
}

int tc_forward(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  int rc = VIL_OK;
  if (use_fused(p, g)) {
    // one kernel: local queries + (when they fit the spare lanes) the global query rows, then a tiny merge
    if (tc::fwd2_fuses_global_rows(p, g) && (p->workspace == nullptr || p->workspace_bytes < tc::fwd2_workspace_floats(p, g) * 4))
      return shared_fail(VIL_E_WORKSPACE, "forward workspace too small: see vil_attn_workspace_bytes");
    if (!(p->skip_mask & 2) && (rc = launch_fused_fwd(p, g, s))) return rc;
    if (g.g > 0 && !tc::fwd2_fuses_global_rows(p, g) && !(p->skip_mask & 1)) rc = simt_global_fwd(p, g, s);
    return rc;
  }
  if (!(p->skip_mask & 2)) {
    rc = tc::is_big_w(g.w) ? tc::launch_fwd_local_big(p, g, s) : tc::launch_fwd_local(p, g, s);
    if (rc) return rc;
  }
  if (g.g > 0 && !(p->skip_mask & 1)) rc = simt_global_fwd(p, g, s);
  return rc;
}

int tc_backward(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool big = tc::is_big_w(g.w);
  int rc = VIL_OK;
  if (use_fused(p, g) && tc::bwd2_applies(p, g)) {
    // 3 launches: pass 1 (delta, chunk-ordered lse2 / delta, dq incl. the global query rows), pass 2 (dk / dv incl. the
    // global key rows), merge of the per-unit global-row partials
    const bool spare = tc::bwd2_fuses_spare_rows(p, g);
    if (!(p->skip_mask & 2) && (rc = tc::launch_bwd2_dq(p, g, s))) return rc;
    if (!(p->skip_mask & 4) && (rc = tc::launch_bwd2_dkv(p, g, s))) return rc;
    if (g.g > 0 && !(p->skip_mask & 1)) {
      if (spare) rc = tc::launch_bwd2_merge(p, g, s);
      else {
        // global rows do not fit the spare lanes (w = 8, nglo > 8, random-shift modes): the round-1 global-token kernels
        // finish dq_g and the dk / dv rows of the global keys; they need the token-ordered delta
        if ((rc = simt_delta(p, g, s))) return rc;
        rc = simt_global_bwd(p, g, s, g.g);
      }
    }
    return rc ? rc : launch_check("tcgen05 fused backward");
  }
  if (!(p->skip_mask & 8)) {
    if ((rc = simt_delta(p, g, s))) return rc;
    if ((rc = big ? tc::launch_bwd_prep_big(p, g, s) : tc::launch_bwd_prep(p, g, s))) return rc;
  }
  if (!(p->skip_mask & 2)) {
    if ((rc = big ? tc::launch_bwd_dq_big(p, g, s) : tc::launch_bwd_dq(p, g, s))) return rc;
  }
  if (!(p->skip_mask & 4)) {
    if ((rc = big ? tc::launch_bwd_dkv_big(p, g, s) : tc::launch_bwd_dkv(p, g, s))) return rc;
  }
  if (g.g > 0 && !(p->skip_mask & 1)) {
    const int rmw_rows = (!big && tc::bwd_fuses_global_rows(p, g)) ? g.g : g.N;
    if ((rc = simt_global_bwd(p, g, s, rmw_rows))) return rc;
  }
  return launch_check("tcgen05 backward");
}

}  // namespace vil
