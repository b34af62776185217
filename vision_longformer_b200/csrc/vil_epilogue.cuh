// Residual / LayerNorm / bias epilogues around the attention kernel ("LayerNorm -> q/kv Linear
// and proj -> residual epilogues", AttnBlock.forward / MlpBlock.forward, msvit.py:313-316, 337-339).
//
// The block structure   x = x + drop_path(branch(norm(x)))   leaves, around every GEMM, a chain of element-wise passes over
// the token stream (bias add inside the GEMM epilogue aside): DropPath scale, residual add, LayerNorm, dtype casts, and in the
// backward the residual fan-in add, the LayerNorm backward, the cast of the branch gradient and one column reduction per
// Linear bias.  All of it is HBM-bound byte shuffling; these kernels do each side of a block boundary in ONE pass:
//
//   addnorm_fwd :  xo = x + rowscale[sample] * (br + bias)          (fp32 residual stream, written once)
//                  y  = LayerNorm(xo) * gamma + beta                (bf16 / fp16 / fp32, the next GEMM's input)
//   addnorm_bwd :  dx  = gres + LayerNorm'(dy)                      (gradient of the residual stream, fp32)
//                  dbr = rowscale[sample] * dx                      (gradient of the branch, low precision)
//                  d_gamma, d_beta, d_bias = column sums            (per-CTA partials + one deterministic reduce, no atomics)
//   bias_act_fwd:  a  = act(z + bias)                               (act = GELU(erf) or identity)
//   bias_act_bwd:  dz = da * act'(z + bias),  d_bias = colsum(dz)   (dz == NULL, act = none: plain column sum of da)
//
// One warp per token row in the addnorm kernels, 128-bit accesses (a lane owns 4 consecutive channels of every 128-channel
// group), the row lives in registers, fp32 math.  bias_act works on (row-lane, 16-byte column group) tiles so that a thread
// keeps the same columns for its whole row slab and the column sums stay in registers.
#pragma once
#include "vil_common.cuh"

namespace vil {
namespace epi {

constexpr int kWarps = 8;             // addnorm: warps (= rows in flight) per CTA
constexpr int kThreads = 256;

template <typename T> struct Vec4;    // 4 consecutive elements <-> float[4]
template <> struct Vec4<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&r)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p)); r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&r)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&r)[4]) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&r)[4]) {
    uint2 raw;
    *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(r[0], r[1]);
    *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(r[2], r[3]);
    *reinterpret_cast<uint2*>(p) = raw;
  }
};
template <> struct Vec4<__half> {
  static __device__ __forceinline__ void ld(const __half* p, float (&r)[4]) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
  }
  static __device__ __forceinline__ void st(__half* p, const float (&r)[4]) {
    uint2 raw;
    *reinterpret_cast<__half2*>(&raw.x) = __floats2half2_rn(r[0], r[1]);
    *reinterpret_cast<__half2*>(&raw.y) = __floats2half2_rn(r[2], r[3]);
    *reinterpret_cast<uint2*>(p) = raw;
  }
};

struct AddNormArgs {
  const float* x;          // (rows, C) residual stream in
  const void* br;          // (rows, C) branch output (low precision) or null
  const float* bias;       // (C) bias of the Linear that produced br, or null
  const float* rowscale;   // (rows / rows_per_sample) DropPath scale per sample, or null
  const float* gamma;      // (C)
  const float* beta;       // (C)
  float* xo;               // (rows, C) residual stream out (null when br is null: xo == x)
  void* y;                 // (rows, C) normalised output
  float* mean;             // (rows)
  float* rstd;             // (rows)
  // backward
  const void* dy;          // (rows, C)
  const float* gres;       // (rows, C) gradient arriving on xo from the rest of the residual stream, or null
  float* dx;               // (rows, C)
  void* dbr;               // (rows, C) or null
  float* partial;          // [grid][3][C]: d_gamma, d_beta, d_bias per CTA
  long long rows, rows_per_sample;
  int C;
  float eps;
};

// NV: 128-channel groups per row (C <= 128 NV).  Channel c = 128 i + 4 lane + e.
template <typename TB, typename TY, int NV>
__global__ void __launch_bounds__(kThreads)
addnorm_fwd(const AddNormArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarps;
  const int C = a.C;
  const TB* br = static_cast<const TB*>(a.br);
  TY* y = static_cast<TY*>(a.y);
  float g[NV][4], bt[NV][4], bs[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = 128 * i + 4 * lane;
#pragma unroll
    for (int e = 0; e < 4; ++e) { g[i][e] = 0.f; bt[i][e] = 0.f; bs[i][e] = 0.f; }
    if (c < C) {
      Vec4<float>::ld(a.gamma + c, g[i]);
      Vec4<float>::ld(a.beta + c, bt[i]);
      if (a.bias != nullptr) Vec4<float>::ld(a.bias + c, bs[i]);
    }
  }
  const float invC = 1.f / (float)C;
  for (long long r = warp; r < a.rows; r += nwarps) {
    const float s = a.rowscale != nullptr ? a.rowscale[r / a.rows_per_sample] : 1.f;
    float v[NV][4], sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = 128 * i + 4 * lane;
#pragma unroll
      for (int e = 0; e < 4; ++e) v[i][e] = 0.f;
      if (c < C) {
        Vec4<float>::ld(a.x + r * C + c, v[i]);
        if (br != nullptr) {
          float b4[4];
          Vec4<TB>::ld(br + r * C + c, b4);
#pragma unroll
          for (int e = 0; e < 4; ++e) v[i][e] = fmaf(s, b4[e] + bs[i][e], v[i][e]);
          Vec4<float>::st(a.xo + r * C + c, v[i]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) sum += v[i][e];
    }
    const float mu = warp_sum(sum) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (128 * i + 4 * lane < C) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float d = v[i][e] - mu; q = fmaf(d, d, q); }
      }
    }
    const float rs = rsqrtf(warp_sum(q) * invC + a.eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = 128 * i + 4 * lane;
      if (c < C) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = fmaf((v[i][e] - mu) * rs, g[i][e], bt[i][e]);
        Vec4<TY>::st(y + r * C + c, o);
      }
    }
    if (lane == 0) { a.mean[r] = mu; a.rstd[r] = rs; }
  }
}

// dx = gres + rstd (dy gamma - mean_c(dy gamma) - xhat mean_c(dy gamma xhat));  dbr = rowscale dx;  column partials per CTA
template <typename TB, typename TY, int NV>
__global__ void __launch_bounds__(kThreads)
addnorm_bwd(const AddNormArgs a) {
  __shared__ float red[3][128 * NV];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * kWarps + wid;
  const long long nwarps = (long long)gridDim.x * kWarps;
  const int C = a.C;
  const TY* dy = static_cast<const TY*>(a.dy);
  TB* dbr = static_cast<TB*>(a.dbr);
  float g[NV][4], dg[NV][4], db[NV][4], dbi[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = 128 * i + 4 * lane;
#pragma unroll
    for (int e = 0; e < 4; ++e) { g[i][e] = 0.f; dg[i][e] = 0.f; db[i][e] = 0.f; dbi[i][e] = 0.f; }
    if (c < C) Vec4<float>::ld(a.gamma + c, g[i]);
  }
  const float invC = 1.f / (float)C;
  for (long long r = warp; r < a.rows; r += nwarps) {
    const float mu = a.mean[r], rs = a.rstd[r];
    const float s = a.rowscale != nullptr ? a.rowscale[r / a.rows_per_sample] : 1.f;
    float xh[NV][4], gy[NV][4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = 128 * i + 4 * lane;
#pragma unroll
      for (int e = 0; e < 4; ++e) { xh[i][e] = 0.f; gy[i][e] = 0.f; }
      if (c < C) {
        float xv[4], d[4];
        Vec4<float>::ld(a.x + r * C + c, xv);          // a.x: the saved residual stream the norm saw (xo of the forward)
        Vec4<TY>::ld(dy + r * C + c, d);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xh[i][e] = (xv[e] - mu) * rs;
          gy[i][e] = d[e] * g[i][e];
          s1 += gy[i][e];
          s2 = fmaf(gy[i][e], xh[i][e], s2);
          dg[i][e] = fmaf(d[e], xh[i][e], dg[i][e]);
          db[i][e] += d[e];
        }
      }
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = 128 * i + 4 * lane;
      if (c < C) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = rs * (gy[i][e] - s1 - xh[i][e] * s2);
        if (a.gres != nullptr) {
          float gr[4];
          Vec4<float>::ld(a.gres + r * C + c, gr);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] += gr[e];
        }
        Vec4<float>::st(a.dx + r * C + c, o);
        if (dbr != nullptr) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { o[e] *= s; dbi[i][e] += o[e]; }
          Vec4<TB>::st(dbr + r * C + c, o);
        }
      }
    }
  }
  // CTA-level reduction of the warps' column sums in a fixed order (deterministic), one partial row per CTA
  for (int w2 = 0; w2 < kWarps; ++w2) {
    if (wid == w2) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = 128 * i + 4 * lane;
        if (c < C) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            red[0][c + e] = (w2 == 0 ? 0.f : red[0][c + e]) + dg[i][e];
            red[1][c + e] = (w2 == 0 ? 0.f : red[1][c + e]) + db[i][e];
            red[2][c + e] = (w2 == 0 ? 0.f : red[2][c + e]) + dbi[i][e];
          }
        }
      }
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < 3 * C; idx += kThreads)
    a.partial[(long long)blockIdx.x * 3 * C + idx] = red[idx / C][idx % C];
}

// partial[nparts][K][C] -> up to three output vectors of length C (null = skipped); fixed summation order
__global__ void __launch_bounds__(256)
colsum_reduce(const float* __restrict__ partial, int nparts, int K, int C, float* __restrict__ o0, float* __restrict__ o1,
              float* __restrict__ o2) {
  __shared__ float sm[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), py = threadIdx.x >> 5;
  float t = 0.f;
  if (col < K * C)
    for (int p = py; p < nparts; p += 8) t += partial[(long long)p * K * C + col];
  sm[py][threadIdx.x & 31] = t;
  __syncthreads();
  if (py == 0 && col < K * C) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += sm[k][threadIdx.x & 31];
    const int seg = col / C;
    float* o = seg == 0 ? o0 : (seg == 1 ? o1 : o2);
    if (o != nullptr) o[col - seg * C] = r;
  }
}

// ---------------------------------------------------------------------------------------------------------- bias + act
__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float u) {
  return 0.5f * (1.f + erff(u * 0.70710678118654752f)) + u * 0.39894228040143268f * __expf(-0.5f * u * u);
}

template <typename T> struct Vec16 {   // 16 bytes of T <-> float[N]
  static constexpr int N = 16 / (int)sizeof(T);
  static __device__ __forceinline__ void ld(const T* p, float (&r)[N]) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(p));
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int u = 0; u < N; ++u) r[u] = ElemTraits<T>::to_f(e[u]);
  }
  static __device__ __forceinline__ void st(T* p, const float (&r)[N]) {
    int4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int u = 0; u < N; ++u) e[u] = ElemTraits<T>::from_f(r[u]);
    *reinterpret_cast<int4*>(p) = raw;
  }
};

// a = act(z + bias): flat grid-stride over 16-byte vectors (C % N == 0, so a vector never straddles a row)
template <typename T, int ACT>
__global__ void __launch_bounds__(kThreads)
bias_act_fwd(const T* __restrict__ z, const float* __restrict__ bias, T* __restrict__ out, long long nvec, int C) {
  constexpr int N = Vec16<T>::N;
  const int G = C / N;
  for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += (long long)gridDim.x * kThreads) {
    const int c = (int)(v % G) * N;
    float x[N];
    Vec16<T>::ld(z + v * N, x);
#pragma unroll
    for (int u = 0; u < N; ++u) {
      const float t = x[u] + (bias != nullptr ? __ldg(bias + c + u) : 0.f);
      x[u] = ACT == 1 ? gelu_f(t) : t;
    }
    Vec16<T>::st(out + v * N, x);
  }
}

// dz = da * act'(z + bias) and per-CTA column sums of dz.  Grid (column slabs, row slabs); a CTA's threads form
// (rows_per_iter row lanes) x (gs column groups of 16 bytes); a thread keeps its column group over the whole row slab.
template <typename T, int ACT>
__global__ void __launch_bounds__(kThreads)
bias_act_bwd(const T* __restrict__ z, const float* __restrict__ bias, const T* __restrict__ da, T* __restrict__ dz,
             float* __restrict__ partial /* [row slabs][C] */, long long rows, int C, int gs, long long rows_per_slab) {
  constexpr int N = Vec16<T>::N;
  __shared__ float red[kThreads * N];
  const int G = C / N;
  const int g0 = blockIdx.x * gs;                          // first column group of this slab
  const int ng = min(gs, G - g0);                          // column groups in this slab
  const int rpi = kThreads / gs;                           // row lanes
  const int rl = threadIdx.x / gs, cg = threadIdx.x % gs;
  const bool active = rl < rpi && cg < ng;
  const long long r0 = (long long)blockIdx.y * rows_per_slab;
  const long long r1 = min(rows, r0 + rows_per_slab);
  const int c = (g0 + cg) * N;
  float acc[N], b[N];
#pragma unroll
  for (int u = 0; u < N; ++u) { acc[u] = 0.f; b[u] = (active && bias != nullptr) ? bias[c + u] : 0.f; }
  if (active) {
    for (long long r = r0 + rl; r < r1; r += rpi) {
      float d[N];
      Vec16<T>::ld(da + r * C + c, d);
      if (ACT == 1) {
        float x[N];
        Vec16<T>::ld(z + r * C + c, x);
#pragma unroll
        for (int u = 0; u < N; ++u) d[u] *= gelu_grad(x[u] + b[u]);
      }
      if (dz != nullptr) {
        Vec16<T>::st(dz + r * C + c, d);
        // the column sum is taken over the values the GEMMs see (rounded to T), like autograd's reduction of dz
#pragma unroll
        for (int u = 0; u < N; ++u) d[u] = ElemTraits<T>::to_f(ElemTraits<T>::from_f(d[u]));
      }
#pragma unroll
      for (int u = 0; u < N; ++u) acc[u] += d[u];
    }
  }
#pragma unroll
  for (int u = 0; u < N; ++u) red[threadIdx.x * N + u] = acc[u];
  __syncthreads();
  if (rl == 0 && cg < ng) {
    for (int k = 1; k < rpi; ++k) {
#pragma unroll
      for (int u = 0; u < N; ++u) acc[u] += red[(k * gs + cg) * N + u];
    }
#pragma unroll
    for (int u = 0; u < N; ++u) partial[(long long)blockIdx.y * C + c + u] = acc[u];
  }
}

}  // namespace epi
}  // namespace vil
