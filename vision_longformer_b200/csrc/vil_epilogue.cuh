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
// L lanes per token row in the addnorm kernels (L = 8 at 96 channels ... 32 at >= 384), 128-bit accesses, the row lives in
// registers, fp32 math.  bias_act works on (row-lane, 16-byte column group) tiles so that a thread
// keeps the same columns for its whole row slab and the column sums stay in registers.
#pragma once
#include "vil_common.cuh"

namespace vil {
namespace epi {

constexpr int kWarps = 4;             // addnorm: warps (= rows in flight) per CTA; small CTAs so that the register-heavy
constexpr int kAnThreads = kWarps * 32;   // instantiations (C >= 384: ~160-240 registers) still keep 8-12 warps per SM
constexpr int kThreads = 256;         // bias_act

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

// Row layout: L lanes per token row (L in {4, 8, 16, 32}), NVL 16-byte vectors per lane; channel c = 4 (sub + L i) + e with
// sub = lane % L.  A warp therefore holds 32 / L rows at once: on the narrow streams (C = 96: L = 8) every lane is busy, a row
// reduction is log2(L) shuffle steps instead of 5, and one instruction stream serves 4 rows (the first version - a warp per
// row - kept 24 of 32 lanes busy at C = 96 and ran at 0.58 of the copy peak).  Rows of a warp are consecutive in memory.
template <int L>
__device__ __forceinline__ float sub_sum(float v) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// gamma / beta / bias live in shared memory (LDS.128 where they are used), not in registers: ncu on the first version showed 90
// registers -> 5 CTAs of 4 warps = 27 % active warps on a pure HBM kernel whose rows are strictly load -> reduce -> store.
template <typename TB, typename TY, int L, int NVL>
__global__ void __launch_bounds__(kAnThreads, NVL <= 3 ? 10 : (NVL <= 4 ? 6 : 3))
addnorm_fwd(const AddNormArgs a) {
  constexpr int RPW = 32 / L;                                  // rows per warp
  __shared__ __align__(16) float prm[3][4 * L * NVL];          // gamma, beta, bias
  const int lane = threadIdx.x & 31, sub = lane % L, rw = lane / L;
  const long long warp = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarps;
  const int C = a.C;
  const TB* br = static_cast<const TB*>(a.br);
  TY* y = static_cast<TY*>(a.y);
  for (int i = threadIdx.x; i < 4 * L * NVL; i += kAnThreads) {
    prm[0][i] = i < C ? a.gamma[i] : 0.f;
    prm[1][i] = i < C ? a.beta[i] : 0.f;
    prm[2][i] = (i < C && a.bias != nullptr) ? a.bias[i] : 0.f;
  }
  __syncthreads();
  const float invC = 1.f / (float)C;
  // sample index of a row (row / rows_per_sample) kept incrementally: a 64-bit division per row would cost as many instructions
  // as the rest of a 96-channel row
  const unsigned rps = (unsigned)a.rows_per_sample;
  const long long rstep = nwarps * RPW;
  const unsigned dq1 = (unsigned)(rstep / rps), dr1 = (unsigned)(rstep % rps);
  unsigned sq = (unsigned)((warp * RPW + rw) / rps), srem = (unsigned)((warp * RPW + rw) % rps);
  for (long long r = warp * RPW + rw; r - rw < a.rows; r += rstep) {      // warp-uniform trip count
    const bool live = r < a.rows;                              // the last warp-row group may be short
    const float s = (live && a.rowscale != nullptr) ? a.rowscale[sq] : 1.f;
    sq += dq1; srem += dr1;
    if (srem >= rps) { srem -= rps; ++sq; }
    float v[NVL][4], sum = 0.f;
#pragma unroll
    for (int i = 0; i < NVL; ++i) {
      const int c = 4 * (sub + L * i);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[i][e] = 0.f;
      if (live && c < C) {
        Vec4<float>::ld(a.x + r * C + c, v[i]);
        if (br != nullptr) {
          float b4[4];
          Vec4<TB>::ld(br + r * C + c, b4);
#pragma unroll
          const float4 bs = *reinterpret_cast<const float4*>(&prm[2][c]);
          v[i][0] = fmaf(s, b4[0] + bs.x, v[i][0]); v[i][1] = fmaf(s, b4[1] + bs.y, v[i][1]);
          v[i][2] = fmaf(s, b4[2] + bs.z, v[i][2]); v[i][3] = fmaf(s, b4[3] + bs.w, v[i][3]);
          Vec4<float>::st(a.xo + r * C + c, v[i]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) sum += v[i][e];
    }
    const float mu = sub_sum<L>(sum) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NVL; ++i) {
      if (4 * (sub + L * i) < C) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float d = v[i][e] - mu; q = fmaf(d, d, q); }
      }
    }
    const float rs = rsqrtf(sub_sum<L>(q) * invC + a.eps);
    if (live) {
#pragma unroll
      for (int i = 0; i < NVL; ++i) {
        const int c = 4 * (sub + L * i);
        if (c < C) {
          const float4 g4 = *reinterpret_cast<const float4*>(&prm[0][c]);
          const float4 b4 = *reinterpret_cast<const float4*>(&prm[1][c]);
          float o[4];
          o[0] = fmaf((v[i][0] - mu) * rs, g4.x, b4.x); o[1] = fmaf((v[i][1] - mu) * rs, g4.y, b4.y);
          o[2] = fmaf((v[i][2] - mu) * rs, g4.z, b4.z); o[3] = fmaf((v[i][3] - mu) * rs, g4.w, b4.w);
          Vec4<TY>::st(y + r * C + c, o);
        }
      }
      if (sub == 0) { a.mean[r] = mu; a.rstd[r] = rs; }
    }
  }
}

// Register budget (measured, S1 / S2 / S3 streams, ms): this version - gamma and the residual-stream gradient of the row in
// registers, all loads of a row issued before the first reduction, 127 registers, 4 CTAs / SM - 0.300 / 0.171 / 0.099; fetching
// gres after the reductions to save 12 registers (5 or 6 CTAs / SM): 0.43 / 0.24 / 0.15 and 0.36 / 0.20 / 0.16 - the dependent
// load in the middle of the row costs more than the extra warps hide.
// dx = gres + rstd (dy gamma - mean_c(dy gamma) - xhat mean_c(dy gamma xhat));  dbr = rowscale dx;  column partials per CTA
template <typename TB, typename TY, int L, int NVL>
__global__ void __launch_bounds__(kAnThreads)
addnorm_bwd(const AddNormArgs a) {
  constexpr int RPW = 32 / L;
  __shared__ float red[3][4 * L * NVL];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, sub = lane % L, rw = lane / L;
  const long long warp = (long long)blockIdx.x * kWarps + wid;
  const long long nwarps = (long long)gridDim.x * kWarps;
  const int C = a.C;
  const TY* dy = static_cast<const TY*>(a.dy);
  TB* dbr = static_cast<TB*>(a.dbr);
  float g[NVL][4], dg[NVL][4], db[NVL][4], dbi[NVL][4];
#pragma unroll
  for (int i = 0; i < NVL; ++i) {
    const int c = 4 * (sub + L * i);
#pragma unroll
    for (int e = 0; e < 4; ++e) { g[i][e] = 0.f; dg[i][e] = 0.f; db[i][e] = 0.f; dbi[i][e] = 0.f; }
    if (c < C) Vec4<float>::ld(a.gamma + c, g[i]);
  }
  const float invC = 1.f / (float)C;
  const unsigned rps = (unsigned)a.rows_per_sample;
  const long long rstep = nwarps * RPW;
  const unsigned dq1 = (unsigned)(rstep / rps), dr1 = (unsigned)(rstep % rps);
  unsigned sq = (unsigned)((warp * RPW + rw) / rps), srem = (unsigned)((warp * RPW + rw) % rps);   // incremental row / rows_per_sample
  for (long long r = warp * RPW + rw; r - rw < a.rows; r += rstep) {
    const bool live = r < a.rows;
    const float s = (live && a.rowscale != nullptr) ? a.rowscale[sq] : 1.f;
    sq += dq1; srem += dr1;
    if (srem >= rps) { srem -= rps; ++sq; }
    const float mu = live ? a.mean[r] : 0.f, rs = live ? a.rstd[r] : 0.f;
    float xh[NVL][4], gy[NVL][4], gr[NVL][4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NVL; ++i) {
      const int c = 4 * (sub + L * i);
#pragma unroll
      for (int e = 0; e < 4; ++e) { xh[i][e] = 0.f; gy[i][e] = 0.f; gr[i][e] = 0.f; }
      if (live && c < C) {
        float xv[4], d[4];
        Vec4<float>::ld(a.x + r * C + c, xv);          // a.x: the saved residual stream the norm saw (xo of the forward)
        Vec4<TY>::ld(dy + r * C + c, d);
        if (a.gres != nullptr) Vec4<float>::ld(a.gres + r * C + c, gr[i]);
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
    const float m1 = sub_sum<L>(s1) * invC;
    const float m2 = sub_sum<L>(s2) * invC;
    if (live) {
#pragma unroll
      for (int i = 0; i < NVL; ++i) {
        const int c = 4 * (sub + L * i);
        if (c < C) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = fmaf(rs, gy[i][e] - m1 - xh[i][e] * m2, gr[i][e]);
          Vec4<float>::st(a.dx + r * C + c, o);
          if (dbr != nullptr) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { o[e] *= s; dbi[i][e] += o[e]; }
            Vec4<TB>::st(dbr + r * C + c, o);
          }
        }
      }
    }
  }
  // column sums: first across the 32 / L row groups of the warp (same channels, fixed butterfly order), then across the
  // CTA's warps in a fixed order through shared memory (deterministic), one partial row per CTA
#pragma unroll
  for (int i = 0; i < NVL; ++i) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
#pragma unroll
      for (int o = L; o < 32; o <<= 1) {
        dg[i][e] += __shfl_xor_sync(0xffffffffu, dg[i][e], o);
        db[i][e] += __shfl_xor_sync(0xffffffffu, db[i][e], o);
        dbi[i][e] += __shfl_xor_sync(0xffffffffu, dbi[i][e], o);
      }
    }
  }
  for (int w2 = 0; w2 < kWarps; ++w2) {
    if (wid == w2 && rw == 0) {
#pragma unroll
      for (int i = 0; i < NVL; ++i) {
        const int c = 4 * (sub + L * i);
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
  for (int idx = threadIdx.x; idx < 3 * C; idx += kAnThreads)
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
// Exact-form GELU (nn.GELU(): 0.5 u (1 + erf(u / sqrt 2))) with erf by Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7):
//   erf(x) = sign(x) (1 - (a1 t + ... + a5 t^5) e^{-x^2}),  t = 1 / (1 + p |x|).
// With x = u / sqrt 2 the exponential is e^{-u^2 / 2} - the same one the derivative's Gaussian term needs - so value and
// derivative cost one MUFU.EX2, one MUFU.RCP and ~12 FMAs per element; libdevice's erff + expf took ~45 instructions and made
// these kernels instruction-bound (3.0 TB/s) instead of HBM-bound.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
struct GeluTerms { float half_erfc_neg; float gauss; };       // 0.5 (1 + erf(u / sqrt 2)),  e^{-u^2 / 2}
__device__ __forceinline__ GeluTerms gelu_terms(float u) {
  const float ax = fabsf(u) * 0.70710678118654752f;
  const float e = ex2_approx(-0.72134752044448170f * u * u);   // e^{-u^2/2} = 2^{-u^2 / (2 ln 2)}  (bare MUFU.EX2)
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.f));        // bare MUFU.RCP: the argument is in [1, inf)
  float pl = fmaf(1.061405429f, t, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  const float q = 0.5f * pl * t * e;                           // 0.5 erfc(|x|)
  GeluTerms r;
  r.half_erfc_neg = u >= 0.f ? 1.f - q : q;                    // 0.5 (1 + erf(x))
  r.gauss = e;
  return r;
}
__device__ __forceinline__ float gelu_f(float u) { return u * gelu_terms(u).half_erfc_neg; }
__device__ __forceinline__ float gelu_grad(float u) {
  const GeluTerms g = gelu_terms(u);
  return fmaf(u * 0.39894228040143268f, g.gauss, g.half_erfc_neg);
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

// a = act(z + bias): flat grid-stride over 16-byte vectors (C % N == 0, so a vector never straddles a row), U vectors per thread
// and iteration so that U x 16 bytes per thread are in flight before the first dependent instruction
template <typename T, int ACT>
__global__ void __launch_bounds__(kThreads)
bias_act_fwd(const T* __restrict__ z, const float* __restrict__ bias, T* __restrict__ out, long long nvec, int C) {
  constexpr int N = Vec16<T>::N;
  constexpr int U = 4;
  const int G = C / N;
  const long long stride = (long long)gridDim.x * kThreads;
  // column group of a vector without a 64-bit modulo in the loop: it advances by (stride mod G) per vector
  const int gstep = (int)(stride % G);
  int cg = (int)(((long long)blockIdx.x * kThreads + threadIdx.x) % G);
  for (long long v0 = (long long)blockIdx.x * kThreads + threadIdx.x; v0 < nvec; v0 += stride * U) {
    float x[U][N];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long v = v0 + k * stride;
      if (v < nvec) Vec16<T>::ld(z + v * N, x[k]);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long v = v0 + k * stride;
      if (v < nvec) {
        const int c = cg * N;
        float b[N];
        if (bias != nullptr) {
#pragma unroll
          for (int u = 0; u < N; u += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c + u));
            b[u] = b4.x; b[u + 1] = b4.y; b[u + 2] = b4.z; b[u + 3] = b4.w;
          }
        } else {
#pragma unroll
          for (int u = 0; u < N; ++u) b[u] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < N; ++u) {
          const float t = x[k][u] + b[u];
          x[k][u] = ACT == 1 ? gelu_f(t) : t;
        }
        Vec16<T>::st(out + v * N, x[k]);
      }
      cg += gstep;
      if (cg >= G) cg -= G;
    }
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
    constexpr int U = ACT == 1 ? 1 : 4;          // rows per iteration (column-sum variant: 4 x 16 bytes per thread in flight)
    for (long long rb = r0 + rl; rb < r1; rb += (long long)rpi * U) {
      float d[U][N], x[U][N];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long r = rb + (long long)k * rpi;
        if (r < r1) {
          Vec16<T>::ld(da + r * C + c, d[k]);
          if (ACT == 1) Vec16<T>::ld(z + r * C + c, x[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long r = rb + (long long)k * rpi;
        if (r < r1) {
          if (ACT == 1) {
#pragma unroll
            for (int u = 0; u < N; ++u) d[k][u] *= gelu_grad(x[k][u] + b[u]);
          }
          if (dz != nullptr) {
            Vec16<T>::st(dz + r * C + c, d[k]);
            // the column sum is taken over the values the GEMMs see (rounded to T), like autograd's reduction of dz
#pragma unroll
            for (int u = 0; u < N; ++u) d[k][u] = ElemTraits<T>::to_f(ElemTraits<T>::from_f(d[k][u]));
          }
#pragma unroll
          for (int u = 0; u < N; ++u) acc[u] += d[k][u];
        }
      }
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
