// tcgen05 / TMA kernel family (sm_100a) -- see DESIGN.md.  Stub until the kernels land.
#pragma once
#include "vil_common.cuh"

namespace vil {
int shared_fail(int code, const char* msg);
void count_launch();

inline const char* tc_why_not(const VilAttnParams*, const Geo&, bool) { return "tcgen05 family not built yet"; }
inline int tc_supported(const VilAttnParams*, const Geo&, bool) { return 0; }
inline long long tc_workspace_bytes(const VilAttnParams*, const Geo&, bool) { return 0; }
inline int tc_forward(const VilAttnParams*, const Geo&, cudaStream_t) { return shared_fail(VIL_E_UNSUPPORTED, "tcgen05 forward not built"); }
inline int tc_backward(const VilAttnParams*, const Geo&, cudaStream_t) { return shared_fail(VIL_E_UNSUPPORTED, "tcgen05 backward not built"); }
}  // namespace vil
