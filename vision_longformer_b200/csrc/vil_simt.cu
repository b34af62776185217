// Translation unit of the SIMT (CUDA-core) kernel family + the global-token / delta kernels both families share.
#include <cstdio>
#include "vil_host.cuh"
#include "vil_simt.cuh"

namespace vil {
namespace {

int head_bucket(int D) { return D <= 8 ? 8 : D <= 16 ? 16 : D <= 32 ? 32 : D <= 64 ? 64 : 128; }

size_t simt_tile_smem(const Geo& g, int HD, bool dkv) {
  const int tw = 4 * g.w - 1;
  size_t bytes = (size_t)(2 * 64 * (HD + 8) + (g.has_bias ? tw * tw : 0)) * sizeof(float);
  if (dkv) bytes += 2 * 64 * sizeof(float);
  bytes += 64 * 2 * sizeof(short) + 64;
  return (bytes + 15) & ~size_t(15);
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (bytes > 227 * 1024) return shared_fail(VIL_E_UNSUPPORTED, "configuration needs more than 227 KB of shared memory");
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return shared_fail(VIL_E_CUDA, cudaGetErrorString(e));
  }
  return VIL_OK;
}

inline float* ws_delta(const VilAttnParams* p) { return static_cast<float*>(p->workspace); }
inline float* ws_delta_g(const VilAttnParams* p, const Geo& g) { return static_cast<float*>(p->workspace) + ws_off_delta_g(g); }

// ------------------------------------------------------------------ shared global-token kernels
template <typename T, int HD, typename TO>
int global_fwd_t(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  launch_global_fwd_kernels<T, HD, TO>(g, t4(p->qg), t4(p->kg), t4(p->vg), t4(p->og), p->lse_g, p->g2l, p->g2g, s);
  count_launch();
  return launch_check("simt_fwd_global");
}

template <typename T, typename TO>
int delta_t(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const long long rows = (long long)g.B * g.H * (g.Nloc + g.g);
  simt_bwd_delta<T, TO><<<(unsigned)((rows + 63) / 64), 256, 0, s>>>(g, t4(p->o), t4(p->d_o), t4(p->og), t4(p->d_og),
                                                                      ws_delta(p), ws_delta_g(p, g));
  count_launch();
  return launch_check("simt_bwd_delta");
}

template <typename T, int HD, typename TO>
int global_bwd_t(const VilAttnParams* p, const Geo& g, cudaStream_t s, int rmw_rows) {
  const bool shared = (p->kg.ptr == p->k.ptr) && (p->vg.ptr == p->v.ptr);
  launch_global_bwd_kernels<T, HD, TO>(g, t4(p->q), t4(p->k), t4(p->v), t4(p->d_o), t4(p->dk), t4(p->dv), t4(p->qg), t4(p->kg),
                                       t4(p->vg), t4(p->d_og), t4(p->dqg), t4(shared ? p->dk : p->dkg),
                                       t4(shared ? p->dv : p->dvg), p->lse, ws_delta(p), p->lse_g, ws_delta_g(p, g), p->g2l,
                                       p->g2g, p->d_g2l, p->d_g2g, shared ? 1 : 0, rmw_rows, s);
  count_launch();
  count_launch();
  return launch_check("simt_bwd_gcol / simt_bwd_grow");
}

// dispatch on (element type, output type, head-dim bucket); F: functor template with operator()<T, HD, TO>()
#define VIL_SIMT_HD(T, TO, CALL)                                   \
  switch (head_bucket(g.D)) {                                      \
    case 8:   return CALL(T, 8, TO);                               \
    case 16:  return CALL(T, 16, TO);                              \
    case 32:  return CALL(T, 32, TO);                              \
    case 64:  return CALL(T, 64, TO);                              \
    default:  return CALL(T, 128, TO);                             \
  }
#define VIL_SIMT_HD64(T, TO, CALL)                                 \
  switch (head_bucket(g.D)) {                                      \
    case 8:   return CALL(T, 8, TO);                               \
    case 16:  return CALL(T, 16, TO);                              \
    case 32:  return CALL(T, 32, TO);                              \
    default:  return CALL(T, 64, TO);                              \
  }
#define VIL_SIMT_TYPES(HDM, CALL)                                                                   \
  if (p->dtype == VIL_F32) { HDM(float, float, CALL) }                                         \
  if (p->dtype == VIL_BF16) {                                                                  \
    if (out_f32(p)) { HDM(__nv_bfloat16, float, CALL) }                                        \
    HDM(__nv_bfloat16, __nv_bfloat16, CALL)                                                    \
  }                                                                                            \
  if (out_f32(p)) { HDM(__half, float, CALL) }                                                 \
  HDM(__half, __half, CALL)

// ------------------------------------------------------------------ SIMT family proper
template <typename T, int HD>
int simt_forward(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  const size_t sm = simt_tile_smem(g, HD, false);
  int rc = set_smem(simt_fwd_local<T, HD>, sm);
  if (rc) return rc;
  const long long blocks = (long long)g.B * g.H * g.mx * g.my * g.npc;
  if (!(p->skip_mask & 2)) {
    simt_fwd_local<T, HD><<<(unsigned)blocks, 128, sm, s>>>(g, t4(p->q), t4(p->k), t4(p->v), t4(p->o), p->lse, p->bias_table,
                                                             p->g2l);
    count_launch();
  }
  if (g.g > 0 && !(p->skip_mask & 1)) {
    if ((rc = global_fwd_t<T, HD, T>(p, g, s))) return rc;
  }
  return launch_check("simt forward");
}

template <typename T, int HD>
int simt_backward(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  if constexpr (HD > 64) {
    return shared_fail(VIL_E_UNSUPPORTED, "backward supports head dim <= 64");
  } else {
    int rc = (p->skip_mask & 8) ? VIL_OK : delta_t<T, T>(p, g, s);
    if (rc) return rc;
    const size_t sm1 = simt_tile_smem(g, HD, false), sm2 = simt_tile_smem(g, HD, true);
    if ((rc = set_smem(simt_bwd_dq<T, HD>, sm1))) return rc;
    if ((rc = set_smem(simt_bwd_dkv<T, HD>, sm2))) return rc;
    const long long blocks = (long long)g.B * g.H * g.mx * g.my * g.npc;
    if (!(p->skip_mask & 2)) {
      simt_bwd_dq<T, HD><<<(unsigned)blocks, 128, sm1, s>>>(g, t4(p->q), t4(p->k), t4(p->v), t4(p->d_o), t4(p->dq), p->lse,
                                                            ws_delta(p), p->bias_table, p->g2l, p->d_bias_table);
      count_launch();
    }
    if (!(p->skip_mask & 4)) {
      simt_bwd_dkv<T, HD><<<(unsigned)blocks, 128, sm2, s>>>(g, t4(p->q), t4(p->k), t4(p->v), t4(p->d_o), t4(p->dk), t4(p->dv),
                                                             p->lse, ws_delta(p), p->bias_table);
      count_launch();
    }
    if (g.g > 0 && !(p->skip_mask & 1)) {
      if ((rc = global_bwd_t<T, HD, T>(p, g, s, g.N))) return rc;
    }
    return launch_check("simt backward");
  }
}

template <typename T>
int simt_dispatch(const VilAttnParams* p, const Geo& g, cudaStream_t s, bool bwd) {
  switch (head_bucket(g.D)) {
    case 8:   return bwd ? simt_backward<T, 8>(p, g, s) : simt_forward<T, 8>(p, g, s);
    case 16:  return bwd ? simt_backward<T, 16>(p, g, s) : simt_forward<T, 16>(p, g, s);
    case 32:  return bwd ? simt_backward<T, 32>(p, g, s) : simt_forward<T, 32>(p, g, s);
    case 64:  return bwd ? simt_backward<T, 64>(p, g, s) : simt_forward<T, 64>(p, g, s);
    default:  return bwd ? simt_backward<T, 128>(p, g, s) : simt_forward<T, 128>(p, g, s);
  }
}

}  // namespace

int simt_run(const VilAttnParams* p, const Geo& g, cudaStream_t s, bool bwd) {
  if (out_f32(p) && p->dtype != VIL_F32)
    return shared_fail(VIL_E_UNSUPPORTED, "VIL_FLAG_F32_OUT (parity build) is implemented by the tcgen05 family only");
  switch (p->dtype) {
    case VIL_F32:  return simt_dispatch<float>(p, g, s, bwd);
    case VIL_BF16: return simt_dispatch<__nv_bfloat16>(p, g, s, bwd);
    default:       return simt_dispatch<__half>(p, g, s, bwd);
  }
}

int simt_global_fwd(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
#define CALL_GF(T, HD, TO) global_fwd_t<T, HD, TO>(p, g, s)
  VIL_SIMT_TYPES(VIL_SIMT_HD, CALL_GF)
#undef CALL_GF
}

int simt_delta(const VilAttnParams* p, const Geo& g, cudaStream_t s) {
  if (p->dtype == VIL_F32) return delta_t<float, float>(p, g, s);
  if (p->dtype == VIL_BF16) return out_f32(p) ? delta_t<__nv_bfloat16, float>(p, g, s) : delta_t<__nv_bfloat16, __nv_bfloat16>(p, g, s);
  return out_f32(p) ? delta_t<__half, float>(p, g, s) : delta_t<__half, __half>(p, g, s);
}

int simt_global_bwd(const VilAttnParams* p, const Geo& g, cudaStream_t s, int rmw_rows) {
  if (g.D > 64) return shared_fail(VIL_E_UNSUPPORTED, "backward supports head dim <= 64");
#define CALL_GB(T, HD, TO) global_bwd_t<T, HD, TO>(p, g, s, rmw_rows)
  VIL_SIMT_TYPES(VIL_SIMT_HD64, CALL_GB)
#undef CALL_GB
}

}  // namespace vil
