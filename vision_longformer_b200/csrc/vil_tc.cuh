// Host side of the tcgen05 / TMA kernel family: coverage test, tensor-map construction, launches.
#pragma once
#include <cstdio>
#include "vil_common.cuh"
#include "vil_simt.cuh"
#include "vil_tc_fwd.cuh"
#include "vil_tc_bwd.cuh"
#include "vil_tc_big.cuh"

namespace vil {
int shared_fail(int code, const char* msg);
void count_launch();

namespace tc {

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline bool aligned16(const VilTensor4& t, int es) {
  return (reinterpret_cast<uintptr_t>(t.ptr) % 16 == 0) && ((t.sb * es) % 16 == 0) && ((t.sh * es) % 16 == 0) &&
         ((t.st * es) % 16 == 0);
}

inline const char* why_not(const VilAttnParams* p, const Geo& g, bool bwd) {
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
  const bool big_w = (g.w == 12 || g.w == 15 || g.w == 31);
  if (!(g.w >= 6 && g.w <= 8) && !big_w) return "chunk size w outside {6,7,8,12,15,31}";
  if (big_w && bwd && p->bias_table != nullptr) return "bias-table gradient for w > 8 is served by the SIMT backward";
  if (g.D % 8 != 0 || g.D > 64) return "head dim must be a multiple of 8 and <= 64";
  if (g.exact == -1) return "cyclic chunks (exact=-1)";
  if (g.g > 16) return "more than 16 global tokens";
  const int tw = 4 * g.w - 1;
  if ((p->bias_table != nullptr || g.exact == 1) && (long long)g.H * tw * tw * 4 > 48 * 1024)
    return "bias / window-mask tables of all heads exceed the shared-memory budget";
  if (!aligned16(p->q, 2) || !aligned16(p->k, 2) || !aligned16(p->v, 2) || !aligned16(p->o, 2))
    return "q/k/v/o base pointers or strides are not 16-byte aligned";
  if (bwd && (!aligned16(p->d_o, 2) || !aligned16(p->dq, 2) || !aligned16(p->dk, 2) || !aligned16(p->dv, 2)))
    return "d_o/dq/dk/dv base pointers or strides are not 16-byte aligned";
  return nullptr;
}

}  // namespace tc

inline const char* tc_why_not(const VilAttnParams* p, const Geo& g, bool bwd) {
  const char* w = tc::why_not(p, g, bwd);
  return w ? w : "supported";
}
inline int tc_supported(const VilAttnParams* p, const Geo& g, bool bwd) { return tc::why_not(p, g, bwd) == nullptr; }
inline long long tc_workspace_bytes(const VilAttnParams*, const Geo& g, bool bwd) {
  return bwd ? (ws_tc_floats(g) + ws_tcg_floats(g)) * 4 : 0;
}

namespace tc {

inline int encode_map(CUtensorMap* m, int dtype, int rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, int DP) {
  static const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  PFN_encodeTiled fn = sm100::get_encode_tiled();
  if (fn == nullptr) return shared_fail(VIL_E_CUDA, "cuTensorMapEncodeTiled driver entry point not available");
  CUresult r = fn(m, dtype == VIL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, dims,
                  strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  DP == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[128];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return shared_fail(VIL_E_CUDA, msg);
  }
  return VIL_OK;
}

// (D, col, row, H, B) map over the LOCAL tokens of a (B,H,T,D) view whose token 0 is `tok0`
inline int local_map(CUtensorMap* m, const VilTensor4& t, long long tok0, const Geo& g, int dtype, int DP, int box_rows = 0) {
  char* base = static_cast<char*>(t.ptr) + tok0 * t.st * 2;
  cuuint64_t dims[5] = {(cuuint64_t)g.D, (cuuint64_t)g.ny, (cuuint64_t)g.nx, (cuuint64_t)g.H, (cuuint64_t)g.B};
  cuuint64_t strides[4] = {(cuuint64_t)t.st * 2, (cuuint64_t)g.ny * t.st * 2, (cuuint64_t)t.sh * 2, (cuuint64_t)t.sb * 2};
  cuuint32_t box[5] = {(cuuint32_t)DP, (cuuint32_t)g.w, (cuuint32_t)(box_rows > 0 ? box_rows : g.w), 1, 1};
  return encode_map(m, dtype, 5, base, dims, strides, box, DP);
}
// (D, token, H, B) map with a 16-token box: the global-token rows
inline int token_map(CUtensorMap* m, const VilTensor4& t, long long ntok, const Geo& g, int dtype, int DP, int box_rows) {
  cuuint64_t dims[4] = {(cuuint64_t)g.D, (cuuint64_t)ntok, (cuuint64_t)g.H, (cuuint64_t)g.B};
  cuuint64_t strides[3] = {(cuuint64_t)t.st * 2, (cuuint64_t)t.sh * 2, (cuuint64_t)t.sb * 2};
  cuuint32_t box[4] = {(cuuint32_t)DP, (cuuint32_t)box_rows, 1, 1};
  return encode_map(m, dtype, 4, t.ptr, dims, strides, box, DP);
}

template <int DP, int W, bool BF16>
int launch_fwd(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  FwdArgs a;
  a.geo = g;
  a.o.p = static_cast<char*>(p->o.ptr); a.o.sb = p->o.sb; a.o.sh = p->o.sh; a.o.st = p->o.st;
  a.lse = p->lse;
  a.table = p->bias_table;
  a.g2l = p->g2l;
  a.cpairs = (g.my + 1) / 2;
  a.num_units = g.B * g.H * g.mx * a.cpairs;
  a.has_tab = (p->bias_table != nullptr) || g.exact == 1;
  a.scale_log2 = g.scale * 1.4426950408889634f;
  CUtensorMap tmQ, tmK, tmV, tmKg, tmVg;
  int rc;
  if ((rc = local_map(&tmQ, p->q, 0, g, p->dtype, DP))) return rc;
  if ((rc = local_map(&tmK, p->k, g.g, g, p->dtype, DP))) return rc;
  if ((rc = local_map(&tmV, p->v, g.g, g, p->dtype, DP))) return rc;
  if ((rc = token_map(&tmKg, p->k, g.N, g, p->dtype, DP, 16))) return rc;
  if ((rc = token_map(&tmVg, p->v, g.N, g, p->dtype, DP, 16))) return rc;
  const int tw = 4 * g.w - 1;
  const int tab_floats = g.H * (a.has_tab ? tw * tw : 0) + g.H * 16;
  int smem = FwdSmem<DP>::total(tab_floats) + BAR_COUNT * 8;
  if (smem < 80 * 1024) smem = 80 * 1024;          // caps residency at 2 CTAs / SM (2 x 256 TMEM columns)
  auto kern = vil_tc_fwd_kernel<DP, W, BF16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  int grid = 2 * num_sms();
  if (grid > a.num_units) grid = a.num_units;
  kern<<<grid, kThreads, smem, s>>>(tmQ, tmK, tmV, tmKg, tmVg, a);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  return VIL_OK;
}

template <int DP, int W, bool BF16>
int launch_fwd_big(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  constexpr int PR = 64 / W, NP = (W + PR - 1) / PR, NPP = (NP + 1) / 2;
  FwdArgs a;
  a.geo = g;
  a.o.p = static_cast<char*>(p->o.ptr); a.o.sb = p->o.sb; a.o.sh = p->o.sh; a.o.st = p->o.st;
  a.lse = p->lse;
  a.table = p->bias_table;
  a.g2l = p->g2l;
  a.cpairs = 0;
  a.num_units = g.B * g.H * g.mx * g.my * NPP;
  a.has_tab = (p->bias_table != nullptr) || g.exact == 1;
  a.scale_log2 = g.scale * 1.4426950408889634f;
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
  e = cudaGetLastError();
  if (e != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  return VIL_OK;
}

template <int DP, bool BF16>
int dispatch_w(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  switch (g.w) {
    case 12: return launch_fwd_big<DP, 12, BF16>(p, g, s);
    case 15: return launch_fwd_big<DP, 15, BF16>(p, g, s);
    case 31: return launch_fwd_big<DP, 31, BF16>(p, g, s);
    case 6: return launch_fwd<DP, 6, BF16>(p, g, s);
    case 7: return launch_fwd<DP, 7, BF16>(p, g, s);
    default: return launch_fwd<DP, 8, BF16>(p, g, s);
  }
}

template <typename T, int HD>
int launch_global_fwd(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  auto vw = [](const VilTensor4& t) { T4 r; r.p = static_cast<char*>(t.ptr); r.sb = t.sb; r.sh = t.sh; r.st = t.st; return r; };
  launch_global_fwd_kernels<T, HD>(g, vw(p->qg), vw(p->kg), vw(p->vg), vw(p->og), p->lse_g, p->g2l, p->g2g, s);
  count_launch();
  return VIL_OK;
}

}  // namespace tc

inline int tc_forward(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool bf = p->dtype == VIL_BF16;
  const int DP = g.D <= 32 ? 32 : 64;
  int rc = VIL_OK;
  if (!(p->skip_mask & 2)) {
    if (DP == 32) rc = bf ? tc::dispatch_w<32, true>(p, g, s) : tc::dispatch_w<32, false>(p, g, s);
    else          rc = bf ? tc::dispatch_w<64, true>(p, g, s) : tc::dispatch_w<64, false>(p, g, s);
    if (rc) return rc;
  }
  if (g.g > 0 && !(p->skip_mask & 1)) {
    const int hb = g.D <= 8 ? 8 : g.D <= 16 ? 16 : g.D <= 32 ? 32 : 64;
#define VIL_GF(T)                                                            \
    switch (hb) {                                                            \
      case 8:  rc = tc::launch_global_fwd<T, 8>(p, g, s); break;             \
      case 16: rc = tc::launch_global_fwd<T, 16>(p, g, s); break;            \
      case 32: rc = tc::launch_global_fwd<T, 32>(p, g, s); break;            \
      default: rc = tc::launch_global_fwd<T, 64>(p, g, s); break;            \
    }
    if (bf) { VIL_GF(__nv_bfloat16) } else { VIL_GF(__half) }
#undef VIL_GF
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  return rc;
}

namespace tc {

inline int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return VIL_OK;
  char msg[192];
  snprintf(msg, sizeof(msg), "%s: %s", what, cudaGetErrorString(e));
  return shared_fail(VIL_E_CUDA, msg);
}

inline T4 t4(const VilTensor4& t) { T4 r; r.p = static_cast<char*>(t.ptr); r.sb = t.sb; r.sh = t.sh; r.st = t.st; return r; }

// w <= 8 pass 2 folds the global QUERY rows in (then simt_bwd_grow only keeps dq_g and the g x g corner)
inline bool bwd_fuses_global_rows(const VilAttnParams* p, const Geo& g) {
  if (g.g == 0 || g.g > 16 || g.w > 8 || (p->skip_mask & 4)) return false;
  const bool shared = (p->kg.ptr == p->k.ptr) && (p->vg.ptr == p->v.ptr);   // global rows attend with the local k / v
  auto ok = [](const VilTensor4& t) {
    return t.ptr != nullptr && (reinterpret_cast<uintptr_t>(t.ptr) % 16 == 0) && ((t.sb * 2) % 16 == 0) &&
           ((t.sh * 2) % 16 == 0) && ((t.st * 2) % 16 == 0);
  };
  return shared && ok(p->qg) && ok(p->d_og);
}

template <int DP, int W, bool BF16>
int launch_bwd(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  int rc0 = VIL_OK;
  float* ws = static_cast<float*>(p->workspace);
  float* lse2c = ws + ws_off_tc(g);
  float* deltac = lse2c + ws_tc_floats(g) / 2;
  if (!(p->skip_mask & 8)) {
    const long long total = ws_tc_floats(g) / 2;
    vil_tc_bwd_prep<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(g, p->lse, ws, lse2c, deltac);
    count_launch();
    if ((rc0 = launch_check("vil_tc_bwd_prep"))) return rc0;
  }
  BwdArgs a;
  a.geo = g;
  a.table = p->bias_table; a.g2l = p->g2l;
  a.lse2c = lse2c; a.deltac = deltac;
  a.cpairs = (g.my + 1) / 2;
  a.num_units = g.B * g.H * g.mx * a.cpairs;
  a.has_tab = (p->bias_table != nullptr) || g.exact == 1;
  a.scale_log2 = g.scale * 1.4426950408889634f;
  a.scale = g.scale;
  a.d_table = p->d_bias_table;
  const bool dbias = p->bias_table != nullptr;
  // pass 2 takes the global QUERY rows as one more 16-column block when their tensors can be TMA sources
  a.fuse_g = bwd_fuses_global_rows(p, g) ? 1 : 0;
  a.lse2g = ws + ws_off_tcg(g);
  a.deltag = a.lse2g + ws_tcg_floats(g) / 2;
  CUtensorMap tmQ, tmDO, tmK, tmV, tmKg, tmVg, tmQg, tmDOg;
  int rc;
  if (a.fuse_g) {
    if (!(p->skip_mask & 8)) {
      vil_tc_bwd_prep_g<<<(g.B * g.H * 16 + 255) / 256, 256, 0, s>>>(g, p->lse_g, ws + ws_off_delta_g(g), p->g2l,
                                                                       ws + ws_off_tcg(g), ws + ws_off_tcg(g) + ws_tcg_floats(g) / 2);
      count_launch();
    }
    if ((rc = token_map(&tmQg, p->qg, g.g, g, p->dtype, DP, 16))) return rc;
    if ((rc = token_map(&tmDOg, p->d_og, g.g, g, p->dtype, DP, 16))) return rc;
  } else {
    if ((rc = token_map(&tmQg, p->k, g.N, g, p->dtype, DP, 16))) return rc;      // never dereferenced
    tmDOg = tmQg;
  }
  if ((rc = local_map(&tmQ, p->q, 0, g, p->dtype, DP))) return rc;
  if ((rc = local_map(&tmDO, p->d_o, 0, g, p->dtype, DP))) return rc;
  if ((rc = local_map(&tmK, p->k, g.g, g, p->dtype, DP))) return rc;
  if ((rc = local_map(&tmV, p->v, g.g, g, p->dtype, DP))) return rc;
  if ((rc = token_map(&tmKg, p->k, g.N, g, p->dtype, DP, 16))) return rc;
  if ((rc = token_map(&tmVg, p->v, g.N, g, p->dtype, DP, 16))) return rc;
  const int tw = 4 * g.w - 1;
  const int tab_floats = g.H * (a.has_tab ? tw * tw : 0) + g.H * 16;
  const int smem_true = BwdSmem<DP>::total(tab_floats) + BB_COUNT * 8;
  int smem = smem_true;
  if (smem < 80 * 1024) smem = 80 * 1024;
  int grid = 2 * num_sms();
  if (grid > a.num_units) grid = a.num_units;
  cudaError_t e;
  if (!(p->skip_mask & 2)) {
    a.out0 = t4(p->dq); a.out1 = t4(p->dq);
    if (!dbias) {
      auto k1 = vil_tc_bwd_dq_kernel<DP, W, BF16, false>;
      if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
        return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
      k1<<<grid, kBwdThreads, smem, s>>>(tmQ, tmDO, tmK, tmV, tmKg, tmVg, a);
    } else {
      // head-affine persistent grid: a multiple of H CTAs, one per SM
      const int smem1 = smem_true + 9 * g.w2 * g.w2 * 4 + tw * tw * 4 + 64;
      int grid1 = (num_sms() / g.H) * g.H;
      if (grid1 > a.num_units) grid1 = ((a.num_units + g.H - 1) / g.H) * g.H;
      auto k1 = vil_tc_bwd_dq_kernel<DP, W, BF16, true>;
      if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1)) != cudaSuccess)
        return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
      k1<<<grid1, kBwdThreads, smem1, s>>>(tmQ, tmDO, tmK, tmV, tmKg, tmVg, a);
    }
    count_launch();
    if ((rc0 = launch_check(dbias ? "vil_tc_bwd_dq_kernel<dbias>" : "vil_tc_bwd_dq_kernel"))) return rc0;
  }
  if (!(p->skip_mask & 4)) {
    auto k2 = vil_tc_bwd_dkv_kernel<DP, W, BF16>;
    if ((e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
      return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
    a.out0 = t4(p->dk); a.out1 = t4(p->dv);
    k2<<<grid, kBwdThreads, smem, s>>>(tmQ, tmDO, tmK, tmV, tmQg, tmDOg, a);
    count_launch();
    if ((rc0 = launch_check("vil_tc_bwd_dkv_kernel"))) return rc0;
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  return VIL_OK;
}

template <int DP, int W, bool BF16>
int launch_bwd_big(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  constexpr int PR = 64 / W, NP = (W + PR - 1) / PR, NPP = (NP + 1) / 2;
  int rc0 = VIL_OK;
  float* ws = static_cast<float*>(p->workspace);
  float* lse2c = ws + ws_off_tc(g);
  float* deltac = lse2c + ws_tc_floats(g) / 2;
  if (!(p->skip_mask & 8)) {
    const long long total = ws_tc_floats(g) / 2;
    vil_tc_bwd_prep_big<W><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(g, p->lse, ws, lse2c, deltac);
    count_launch();
    if ((rc0 = launch_check("vil_tc_bwd_prep_big"))) return rc0;
  }
  BwdArgs a;
  a.geo = g;
  a.table = p->bias_table; a.g2l = p->g2l;
  a.lse2c = lse2c; a.deltac = deltac;
  a.cpairs = 0;
  a.num_units = g.B * g.H * g.mx * g.my * NPP;
  a.has_tab = (p->bias_table != nullptr) || g.exact == 1;
  a.scale_log2 = g.scale * 1.4426950408889634f;
  a.scale = g.scale;
  a.d_table = nullptr;
  a.fuse_g = 0; a.lse2g = nullptr; a.deltag = nullptr;
  CUtensorMap tmQ, tmDO, tmK, tmV, tmKg, tmVg;
  int rc;
  if ((rc = local_map(&tmQ, p->q, 0, g, p->dtype, DP, PR))) return rc;
  if ((rc = local_map(&tmDO, p->d_o, 0, g, p->dtype, DP, PR))) return rc;
  if ((rc = local_map(&tmK, p->k, g.g, g, p->dtype, DP, PR))) return rc;
  if ((rc = local_map(&tmV, p->v, g.g, g, p->dtype, DP, PR))) return rc;
  if ((rc = token_map(&tmKg, p->k, g.N, g, p->dtype, DP, 16))) return rc;
  if ((rc = token_map(&tmVg, p->v, g.N, g, p->dtype, DP, 16))) return rc;
  const int tw = 4 * g.w - 1;
  const int tab_floats = g.H * (a.has_tab ? tw * tw : 0) + g.H * 16;
  int smem = BwdSmem<DP>::total(tab_floats) + BB_COUNT * 8;
  if (smem < 80 * 1024) smem = 80 * 1024;
  int grid = 2 * num_sms();
  if (grid > a.num_units) grid = a.num_units;
  cudaError_t e;
  if (!(p->skip_mask & 2)) {
    auto k1 = vil_tc_bwd_dq_big_kernel<DP, W, BF16>;
    if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
      return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
    a.out0 = t4(p->dq); a.out1 = t4(p->dq);
    k1<<<grid, kBwdThreads, smem, s>>>(tmQ, tmDO, tmK, tmV, tmKg, tmVg, a);
    count_launch();
    if ((rc0 = launch_check("vil_tc_bwd_dq_big_kernel"))) return rc0;
  }
  if (!(p->skip_mask & 4)) {
    auto k2 = vil_tc_bwd_dkv_big_kernel<DP, W, BF16>;
    if ((e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
      return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
    a.out0 = t4(p->dk); a.out1 = t4(p->dv);
    k2<<<grid, kBwdThreads, smem, s>>>(tmQ, tmDO, tmK, tmV, a);
    count_launch();
    if ((rc0 = launch_check("vil_tc_bwd_dkv_big_kernel"))) return rc0;
  }
  return VIL_OK;
}

template <int DP, bool BF16>
int dispatch_w_bwd(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  switch (g.w) {
    case 12: return launch_bwd_big<DP, 12, BF16>(p, g, s);
    case 15: return launch_bwd_big<DP, 15, BF16>(p, g, s);
    case 31: return launch_bwd_big<DP, 31, BF16>(p, g, s);
    case 6: return launch_bwd<DP, 6, BF16>(p, g, s);
    case 7: return launch_bwd<DP, 7, BF16>(p, g, s);
    default: return launch_bwd<DP, 8, BF16>(p, g, s);
  }
}

// delta prologue + global-token kernels shared with the SIMT family
template <typename T, int HD>
int launch_bwd_shared(const VilAttnParams* p, const Geo& g, cudaStream_t s, bool prologue) {
  float* ws = static_cast<float*>(p->workspace);
  float* delta_g = ws + ws_off_delta_g(g);
  if (prologue) {
    const long long rows = (long long)g.B * g.H * (g.Nloc + g.g);
    simt_bwd_delta<T><<<(unsigned)((rows + 63) / 64), 256, 0, s>>>(g, t4(p->o), t4(p->d_o), t4(p->og), t4(p->d_og), ws, delta_g);
    count_launch();
    return VIL_OK;
  }
  if (g.g == 0 || (p->skip_mask & 1)) return VIL_OK;
  const bool shared = (p->kg.ptr == p->k.ptr) && (p->vg.ptr == p->v.ptr);
  const int rmw_rows = bwd_fuses_global_rows(p, g) ? g.g : g.N;      // keys whose dk / dv rows simt_bwd_grow still updates
  launch_global_bwd_kernels<T, HD>(g, t4(p->q), t4(p->k), t4(p->v), t4(p->d_o), t4(p->dk), t4(p->dv), t4(p->qg), t4(p->kg),
                                   t4(p->vg), t4(p->d_og), t4(p->dqg), t4(shared ? p->dk : p->dkg), t4(shared ? p->dv : p->dvg),
                                   p->lse, ws, p->lse_g, delta_g, p->g2l, p->g2g, p->d_g2l, p->d_g2g, shared ? 1 : 0, rmw_rows, s);
  count_launch();
  count_launch();
  return VIL_OK;
}

template <typename T>
int bwd_shared_dispatch(const VilAttnParams* p, const Geo& g, cudaStream_t s, bool prologue) {
  switch (g.D <= 8 ? 8 : g.D <= 16 ? 16 : g.D <= 32 ? 32 : 64) {
    case 8:  return launch_bwd_shared<T, 8>(p, g, s, prologue);
    case 16: return launch_bwd_shared<T, 16>(p, g, s, prologue);
    case 32: return launch_bwd_shared<T, 32>(p, g, s, prologue);
    default: return launch_bwd_shared<T, 64>(p, g, s, prologue);
  }
}

}  // namespace tc

inline int tc_backward(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const bool bf = p->dtype == VIL_BF16;
  const int DP = g.D <= 32 ? 32 : 64;
  int rc = VIL_OK;
  if (!(p->skip_mask & 8)) {
    rc = bf ? tc::bwd_shared_dispatch<__nv_bfloat16>(p, g, s, true) : tc::bwd_shared_dispatch<__half>(p, g, s, true);
    if (rc) return rc;
  }
  if (DP == 32) rc = bf ? tc::dispatch_w_bwd<32, true>(p, g, s) : tc::dispatch_w_bwd<32, false>(p, g, s);
  else          rc = bf ? tc::dispatch_w_bwd<64, true>(p, g, s) : tc::dispatch_w_bwd<64, false>(p, g, s);
  if (rc) return rc;
  rc = bf ? tc::bwd_shared_dispatch<__nv_bfloat16>(p, g, s, false) : tc::bwd_shared_dispatch<__half>(p, g, s, false);
  if (rc) return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  return VIL_OK;
}

}  // namespace vil
