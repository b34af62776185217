// Host-side helpers of the tcgen05 / TMA kernel family: per-device SM count, tensor-map construction, argument
// blocks.  Included by every vil_tc_*.cu translation unit.
#pragma once
#include <cstdio>
#include "vil_host.cuh"
#include "vil_sm100.cuh"

namespace vil {
namespace tc {

inline int num_sms() {                       // of the CURRENT device (the Python binding sets it to the tensors' device)
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

inline bool aligned16(const VilTensor4& t, int es) {
  return (reinterpret_cast<uintptr_t>(t.ptr) % 16 == 0) && ((t.sb * es) % 16 == 0) && ((t.sh * es) % 16 == 0) &&
         ((t.st * es) % 16 == 0);
}

inline int encode_map(CUtensorMap* m, int dtype, int rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, int DP) {
  static const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  sm100::PFN_encodeTiled fn = sm100::get_encode_tiled();
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
// same view, explicit (box_cols x box_rows) box: the key-row blocks of vil_tc_fwd5 (8 columns x 3 rows)
inline int local_map_box(CUtensorMap* m, const VilTensor4& t, long long tok0, const Geo& g, int dtype, int DP, int box_cols, int box_rows) {
  char* base = static_cast<char*>(t.ptr) + tok0 * t.st * 2;
  cuuint64_t dims[5] = {(cuuint64_t)g.D, (cuuint64_t)g.ny, (cuuint64_t)g.nx, (cuuint64_t)g.H, (cuuint64_t)g.B};
  cuuint64_t strides[4] = {(cuuint64_t)t.st * 2, (cuuint64_t)g.ny * t.st * 2, (cuuint64_t)t.sh * 2, (cuuint64_t)t.sb * 2};
  cuuint32_t box[5] = {(cuuint32_t)DP, (cuuint32_t)box_cols, (cuuint32_t)box_rows, 1, 1};
  return encode_map(m, dtype, 5, base, dims, strides, box, DP);
}
// (D, token, H, B) map with a `box_rows`-token box: the global-token rows
inline int token_map(CUtensorMap* m, const VilTensor4& t, long long ntok, const Geo& g, int dtype, int DP, int box_rows) {
  cuuint64_t dims[4] = {(cuuint64_t)g.D, (cuuint64_t)ntok, (cuuint64_t)g.H, (cuuint64_t)g.B};
  cuuint64_t strides[3] = {(cuuint64_t)t.st * 2, (cuuint64_t)t.sh * 2, (cuuint64_t)t.sb * 2};
  cuuint32_t box[4] = {(cuuint32_t)DP, (cuuint32_t)box_rows, 1, 1};
  return encode_map(m, dtype, 4, t.ptr, dims, strides, box, DP);
}

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

// w = 14: the dense attention of a 14x14(+nglo) stage is the single-chunk case of the sliding-chunk operator
inline bool is_big_w(int w) { return w == 12 || w == 14 || w == 15 || w == 31; }

}  // namespace tc
}  // namespace vil
