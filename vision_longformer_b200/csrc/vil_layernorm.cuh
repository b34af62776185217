// LayerNorm forward / backward for the token streams around the attention kernel (SURVEY.md section 8 (f) row 4:
// "LayerNorm -> q/kv Linear ... epilogues around the kernel", AttnBlock.forward msvit.py:313-316).
//
// Pure HBM-bandwidth kernels: one warp per token row, the row (C <= 1024 channels) lives in registers, fp32 math.
// The forward can emit bf16/fp16 directly from an fp32 residual stream (what `autocast` does in two passes:
// fp32 LayerNorm, then a cast in front of the Linear).  The backward produces dx plus per-CTA partial sums of
// d_gamma / d_beta that a second tiny kernel reduces (deterministic, no atomics).
#pragma once
#include "vil_common.cuh"

namespace vil {
namespace ln {

constexpr int kWarpsPerCta = 8;
constexpr int kMaxPerLane = 32;   // C <= 1024

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return ElemTraits<T>::to_f(*p); }

template <typename TIn, typename TOut, int NPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
layernorm_fwd(const TIn* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              TOut* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, long long rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
  float g[NPL], bt[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    g[i] = c < C ? gamma[c] : 0.f;
    bt[i] = c < C ? beta[c] : 0.f;
  }
  const float invC = 1.f / (float)C;
  for (long long r = warp; r < rows; r += nwarps) {
    const TIn* xr = x + r * C;
    float v[NPL], s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < C ? ldf(xr + c) : 0.f;
      s += v[i];
    }
    const float mu = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      const float d = c < C ? v[i] - mu : 0.f;
      q = fmaf(d, d, q);
    }
    const float rs = rsqrtf(warp_sum(q) * invC + eps);
    TOut* yr = y + r * C;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      if (c < C) yr[c] = ElemTraits<TOut>::from_f(fmaf((v[i] - mu) * rs, g[i], bt[i]));
    }
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

// dx = rstd * (dy*gamma - mean_c(dy*gamma) - xhat * mean_c(dy*gamma*xhat));  partial d_gamma / d_beta per CTA
template <typename TX, typename TDy, int NPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
layernorm_bwd(const TDy* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, TX* __restrict__ dx,
              float* __restrict__ partial /* [gridDim.x][2][C] */, long long rows, int C) {
  __shared__ float red[2][1024];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * kWarpsPerCta + wid;
  const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
  float g[NPL], dg[NPL], db[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    g[i] = c < C ? gamma[c] : 0.f;
    dg[i] = 0.f; db[i] = 0.f;
  }
  const float invC = 1.f / (float)C;
  for (long long r = warp; r < rows; r += nwarps) {
    const TX* xr = x + r * C;
    const TDy* dyr = dy + r * C;
    const float mu = mean[r], rs = rstd[r];
    float xh[NPL], gy[NPL], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      const float xv = c < C ? ldf(xr + c) : mu;
      const float d = c < C ? ldf(dyr + c) : 0.f;
      xh[i] = (xv - mu) * rs;
      gy[i] = d * g[i];
      s1 += gy[i];
      s2 = fmaf(gy[i], xh[i], s2);
      dg[i] = fmaf(d, xh[i], dg[i]);
      db[i] += d;
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
    TX* dxr = dx + r * C;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      if (c < C) dxr[c] = ElemTraits<TX>::from_f(rs * (gy[i] - s1 - xh[i] * s2));
    }
  }
  // CTA-level reduction of the 8 warps' partial sums (fixed order -> deterministic), one partial row per CTA
  for (int w2 = 0; w2 < kWarpsPerCta; ++w2) {
    if (wid == w2) {
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int c = lane + 32 * i;
        if (c < C) {
          red[0][c] = (w2 == 0 ? 0.f : red[0][c]) + dg[i];
          red[1][c] = (w2 == 0 ? 0.f : red[1][c]) + db[i];
        }
      }
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < 2 * C; idx += kWarpsPerCta * 32)
    partial[(long long)blockIdx.x * 2 * C + idx] = red[idx / C][idx % C];
}

// partial[nparts][2*C] -> out[2*C]; block = (32 columns) x (8 partial lanes), fixed summation order
__global__ void __launch_bounds__(256)
layernorm_bwd_reduce(const float* __restrict__ partial, float* __restrict__ dgamma, float* __restrict__ dbeta, int nparts, int C) {
  __shared__ float sm[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), py = threadIdx.x >> 5;
  float t = 0.f;
  if (col < 2 * C)
    for (int p = py; p < nparts; p += 8) t += partial[(long long)p * 2 * C + col];
  sm[py][threadIdx.x & 31] = t;
  __syncthreads();
  if (py == 0 && col < 2 * C) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += sm[k][threadIdx.x & 31];
    (col < C ? dgamma : dbeta)[col < C ? col : col - C] = r;
  }
}

}  // namespace ln
}  // namespace vil
