// SIMT (CUDA-core, fp32 accumulate) kernel family of the Vision-Longformer attention.
//
// Covers EVERY configuration of the reference operator (any w, exact in {0,1,-1},
// mode in {-1,0,1..8}, any nglo, D <= 128 forward / D <= 64 backward, fp32 / bf16 /
// fp16 I/O).  It is (a) the fp32 parity build (1e-5 vs the fp64 oracle), (b) the
// path for configurations the tcgen05 family does not cover, and (c) the home of
// the small global-token kernels that both families share.
//
// Math restated from the reference (closed forms verified in oracle/vil_oracle.py):
//   local query i=(r_i,c_i) in chunk (R,C); for every visited chunk offset (dR,dC)
//   (slidingchunk_2d.py:37-79) and key (kr,kc) of chunk (R+dR, C+dC):
//     allowed   <- mask rules of slidingchunk_2d.py:249-318 (zero / exact / cyclic)
//     bias      <- table[(dr + 2w-1)*(4w-1) + (dc + 2w-1), h],  dr = qr-(dR*w+kr), dc likewise
//                  (longformer2d.py:68-100, 159-178)
//   joint softmax over [global keys | allowed local keys] (longformer2d.py:183-185),
//   o = P.[v_g | v_loc] (:194-200).  Backward: dS = P o (dP - delta) (SlidingChunk2D.backward
//   + softmax autograd, slidingchunk_2d.py:234-246).
#pragma once
#include "vil_common.cuh"

namespace vil {

// fp32 smem tile row: two halves of HD/2 floats separated by 4 floats of padding so that the two
// threads of a (row, half) pair hit different banks.
template <int HD> struct Tile {
  static constexpr int HH = HD / 2;
  static constexpr int HS = HD + 8;
  static __device__ __forceinline__ int off(int half) { return half * (HH + 4); }
};

struct ChunkId { int b, h, R, C, piece; };

__device__ __forceinline__ ChunkId decode_block(const Geo& g, int bid) {
  ChunkId c;
  c.piece = bid % g.npc; bid /= g.npc;
  c.C = bid % g.my; bid /= g.my;
  c.R = bid % g.mx; bid /= g.mx;
  c.h = bid % g.H;
  c.b = bid / g.H;
  return c;
}

// ----------------------------------------------------------------------------------------------
// forward, local queries.  CTA = one 64-query piece of one chunk of one (b,h); thread pair per query,
// each thread owns half of the head dimension.
// ----------------------------------------------------------------------------------------------
template <typename T, int HD>
__global__ void __launch_bounds__(128)
simt_fwd_local(Geo geo, T4 q, T4 k, T4 v, T4 o, float* __restrict__ lse,
               const float* __restrict__ table, const float* __restrict__ g2l) {
  using TL = Tile<HD>;
  constexpr int HH = TL::HH, HS = TL::HS;
  extern __shared__ float smem[];
  float* Ks = smem;
  float* Vs = Ks + 64 * HS;
  float* tab = Vs + 64 * HS;
  const int tw = 4 * geo.w - 1;
  const int tabn = geo.has_bias ? tw * tw : 0;
  short* kvr = reinterpret_cast<short*>(tab + tabn);
  short* kvc = kvr + 64;
  unsigned char* kfl = reinterpret_cast<unsigned char*>(kvc + 64);

  const ChunkId cid = decode_block(geo, blockIdx.x);
  const int b = cid.b, h = cid.h, R = cid.R, C = cid.C;
  const int tid = threadIdx.x, slot = tid >> 1, half = tid & 1;
  const int w = geo.w, D = geo.D;

  for (int i = tid; i < tabn; i += 128) tab[i] = table[(long long)i * geo.H + h];

  const int l = cid.piece * 64 + slot;
  const int qr = l / w, qc = l % w;
  const int r = R * w + qr, c = C * w + qc;
  const bool qvalid = (l < geo.w2) && (r < geo.nx) && (c < geo.ny);

  float qh[HH], oh[HH];
#pragma unroll
  for (int i = 0; i < HH; ++i) { qh[i] = 0.f; oh[i] = 0.f; }
  if (qvalid) load_seg<T, HH>(row_ptr<T>(q, b, h, (long long)r * geo.ny + c), half * HH, D, qh);
  float m = -INFINITY, lsum = 0.f;

  const int ngp = (geo.g + 63) / 64;
  const int npieces = ngp + geo.noffs * geo.npc;
  for (int pi = 0; pi < npieces; ++pi) {
    const bool isg = pi < ngp;
    int dR = 0, dC = 0, KR = 0, KC = 0, kp = 0;
    if (!isg) {
      const int oi = (pi - ngp) / geo.npc;
      kp = (pi - ngp) % geo.npc;
      dR = geo.offR[oi]; dC = geo.offC[oi];
      KR = R + dR; KC = C + dC;
      if (geo.exact == -1) { KR = (KR + geo.mx) % geo.mx; KC = (KC + geo.my) % geo.my; }
      else if (KR < 0 || KR >= geo.mx || KC < 0 || KC >= geo.my) continue;   // CTA-uniform
    }
    __syncthreads();
    {   // stage one piece of <=64 keys: thread pair (slot, half) loads half a row of K and V
      float kk[HH], vv[HH];
#pragma unroll
      for (int i = 0; i < HH; ++i) { kk[i] = 0.f; vv[i] = 0.f; }
      int flag = 0, vr = 0, vc = 0; long long tok = -1;
      if (isg) {
        const int t = pi * 64 + slot;
        if (t < geo.g) { flag = 2; vr = t; tok = t; }
      } else {
        const int lk = kp * 64 + slot;
        if (lk < geo.w2) {
          const int kr = lk / w, kc = lk % w;
          const int ar = KR * w + kr, ac = KC * w + kc;
          const bool real = (ar < geo.nx) && (ac < geo.ny);
          if (geo.exact == -1)
            flag = !(((R + dR == geo.mx - 1) && (kr >= w - geo.padx)) ||
                     ((C + dC == geo.my - 1) && (kc >= w - geo.pady)));
          else
            flag = real;
          if (flag && real) tok = geo.g + (long long)ar * geo.ny + ac;   // phantom padding keys keep K=V=0
          vr = dR * w + kr; vc = dC * w + kc;
        }
      }
      if (tok >= 0) {
        load_seg<T, HH>(row_ptr<T>(k, b, h, tok), half * HH, D, kk);
        load_seg<T, HH>(row_ptr<T>(v, b, h, tok), half * HH, D, vv);
      }
      float* kd = Ks + slot * HS + TL::off(half);
      float* vd = Vs + slot * HS + TL::off(half);
#pragma unroll
      for (int i = 0; i < HH; ++i) { kd[i] = kk[i]; vd[i] = vv[i]; }
      if (half == 0) { kvr[slot] = (short)vr; kvc[slot] = (short)vc; kfl[slot] = (unsigned char)flag; }
    }
    __syncthreads();
    for (int j = 0; j < 64; ++j) {
      const int f = kfl[j];
      if (!f) continue;                                   // warp-uniform
      const float* kd = Ks + j * HS + TL::off(half);
      float sp = 0.f;
#pragma unroll
      for (int i = 0; i < HH; ++i) sp = fmaf(qh[i], kd[i], sp);
      sp += __shfl_xor_sync(0xffffffffu, sp, 1);
      float bias = 0.f; bool ok = qvalid;
      if (f == 2) {
        if (geo.has_bias) bias = g2l[((long long)geo.H + h) * geo.g + kvr[j]];
      } else {
        const int dr = qr - kvr[j], dc = qc - kvc[j];
        if (geo.exact == 1 && (abs(dr) > w || abs(dc) > w)) ok = false;
        if (geo.has_bias && ok) bias = tab[(dr + 2 * w - 1) * tw + dc + 2 * w - 1];
      }
      if (ok) {
        const float s = fmaf(geo.scale, sp, bias);
        const float* vd = Vs + j * HS + TL::off(half);
        if (s > m) {
          const float corr = __expf(m - s);
          lsum = lsum * corr + 1.f;
#pragma unroll
          for (int i = 0; i < HH; ++i) oh[i] = fmaf(oh[i], corr, vd[i]);
          m = s;
        } else {
          const float p = __expf(s - m);
          lsum += p;
#pragma unroll
          for (int i = 0; i < HH; ++i) oh[i] = fmaf(p, vd[i], oh[i]);
        }
      }
    }
  }
  if (qvalid) {
    const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
#pragma unroll
    for (int i = 0; i < HH; ++i) oh[i] *= inv;
    const long long tokq = (long long)r * geo.ny + c;
    store_seg<T, HH>(row_ptr_w<T>(o, b, h, tokq), half * HH, D, oh);
    if (half == 0) lse[((long long)b * geo.H + h) * geo.Nloc + tokq] = m + logf(lsum);
  }
}

// ----------------------------------------------------------------------------------------------
// Global-token kernels: sub-warp ROW GROUPS.  LPR = HD/8 lanes share one token row, lane `sub` owns the 8 channels
// [8 sub, 8 sub + 8) (one 16-byte load for bf16 / fp16), so a warp streams 32/LPR rows per iteration with fully
// coalesced accesses and ~40 registers per thread.  (The first version gave every thread a whole row: HD-long
// register arrays, 1 CTA per SM, 0.5 ms per layer for a few hundred MB of traffic.)
// ----------------------------------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {          // over the LPR lanes of one row group
#pragma unroll
  for (int o = 1; o < LPR; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int LPR>
__device__ __forceinline__ float rows_sum(float v) {           // over the 32/LPR row groups of a warp (same `sub`)
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8]) {
  float s0 = a[0] * b[0], s1 = a[1] * b[1];
#pragma unroll
  for (int i = 2; i < 8; i += 2) { s0 = fmaf(a[i], b[i], s0); s1 = fmaf(a[i + 1], b[i + 1], s1); }
  return s0 + s1;
}

// ----------------------------------------------------------------------------------------------
// forward, global query rows: dense attention of the nglo global queries over all N keys
// (longformer2d.py:210-227).  CTA = one (b, h, a); 8 warps x (32/LPR) rows per iteration.
// ----------------------------------------------------------------------------------------------
template <typename T, int HD, typename TO = T>      // TO: element type of the OUTPUT (fp32 in the parity build)
__global__ void __launch_bounds__(256)
simt_fwd_global(Geo geo, T4 qg, T4 kg, T4 vg, T4 og, float* __restrict__ lse_g, const float* __restrict__ g2l,
                const float* __restrict__ g2g) {
  constexpr int LPR = HD / 8, RPW = 32 / LPR, ROWS = 8 * RPW;
  __shared__ float red_m[8], red_l[8];
  __shared__ float red_o[8][HD];
  const int a = blockIdx.x % geo.g;
  const int h = (blockIdx.x / geo.g) % geo.H;
  const int b = blockIdx.x / (geo.g * geo.H);
  const int tid = threadIdx.x, D = geo.D, lane = tid & 31, warp = tid >> 5, sub = lane % LPR, rw = lane / LPR;
  const int row0 = 0, row1 = geo.N;
  float q8[8];
  load_seg<T, 8>(row_ptr<T>(qg, b, h, a), 8 * sub, D, q8);
  float m = -INFINITY, lsum = 0.f, oacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) oacc[i] = 0.f;
  const float bl = geo.has_bias ? g2l[(long long)h * geo.g + a] : 0.f;       // g2l[0][h][a]
  for (int base = row0 + warp * RPW; base < row1; base += ROWS) {
    const int j = base + rw;
    const bool valid = j < row1;
    const int jc = valid ? j : row1 - 1;
    float kk[8], vv[8];
    load_seg<T, 8>(row_ptr<T>(kg, b, h, jc), 8 * sub, D, kk);
    load_seg<T, 8>(row_ptr<T>(vg, b, h, jc), 8 * sub, D, vv);
    const float sp = group_sum<LPR>(dot8(q8, kk));
    float bias = bl;
    if (geo.has_bias && jc < geo.g) bias = g2g[((long long)h * geo.g + a) * geo.g + jc];
    const float sc = valid ? fmaf(geo.scale, sp, bias) : -INFINITY;
    const float mn = fmaxf(m, sc);
    if (mn > -INFINITY) {                        // branch-free online softmax step of this row group
      const float corr = __expf(m - mn), p = __expf(sc - mn);
      lsum = fmaf(lsum, corr, p);
#pragma unroll
      for (int i = 0; i < 8; ++i) oacc[i] = fmaf(oacc[i], corr, p * vv[i]);
      m = mn;
    }
  }
  // merge the row groups of the warp, then the 8 warps (fixed order: deterministic)
  float mw = m;
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
  const float scw = (m == -INFINITY) ? 0.f : __expf(m - mw);
  const float lw = rows_sum<LPR>(lsum * scw);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float x = rows_sum<LPR>(oacc[i] * scw);
    if (rw == 0) red_o[warp][8 * sub + i] = x;
  }
  if (lane == 0) { red_m[warp] = mw; red_l[warp] = lw; }
  __syncthreads();
  if (tid < HD) {
    float M = -INFINITY;
    for (int x = 0; x < 8; ++x) M = fmaxf(M, red_m[x]);
    float L = 0.f, O = 0.f;
    for (int x = 0; x < 8; ++x) {
      const float s2 = (red_m[x] == -INFINITY) ? 0.f : __expf(red_m[x] - M);
      L += red_l[x] * s2; O += red_o[x][tid] * s2;
    }
    if (tid < D) row_ptr_w<TO>(og, b, h, a)[tid] = ElemTraits<TO>::from_f(O / L);
    if (tid == 0) lse_g[((long long)b * geo.H + h) * geo.g + a] = M + logf(L);
  }
}

// ----------------------------------------------------------------------------------------------
// backward prologue: delta_i = sum_c dO_ic * O_ic for local rows (into delta) and global rows (delta_g)
// ----------------------------------------------------------------------------------------------
template <typename T, typename TO = T>               // TO: element type of o / og (fp32 in the parity build)
__global__ void simt_bwd_delta(Geo geo, T4 o, T4 d_o, T4 og, T4 d_og,
                               float* __restrict__ delta, float* __restrict__ delta_g) {
  const long long rows_loc = (long long)geo.B * geo.H * geo.Nloc;
  const long long rows = rows_loc + (long long)geo.B * geo.H * geo.g;
  const long long idx = (long long)blockIdx.x * (blockDim.x >> 2) + (threadIdx.x >> 2);  // 4 threads per row
  const int sub = threadIdx.x & 3;
  float acc = 0.f;
  if (idx < rows) {
    const TO* po;
    const T* pd;
    if (idx < rows_loc) {
      const long long t = idx % geo.Nloc; const long long bh = idx / geo.Nloc;
      po = row_ptr<TO>(o, (int)(bh / geo.H), (int)(bh % geo.H), t);
      pd = row_ptr<T>(d_o, (int)(bh / geo.H), (int)(bh % geo.H), t);
    } else {
      const long long e = idx - rows_loc;
      const long long t = e % geo.g; const long long bh = e / geo.g;
      po = row_ptr<TO>(og, (int)(bh / geo.H), (int)(bh % geo.H), t);
      pd = row_ptr<T>(d_og, (int)(bh / geo.H), (int)(bh % geo.H), t);
    }
    for (int cc = sub; cc < geo.D; cc += 4)
      acc = fmaf(ElemTraits<TO>::to_f(po[cc]), ElemTraits<T>::to_f(pd[cc]), acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (idx < rows && sub == 0) {
    if (idx < rows_loc) delta[idx] = acc; else delta_g[idx - rows_loc] = acc;
  }
}

// ----------------------------------------------------------------------------------------------
// backward pass 1 (query-stationary): dq, d_bias_table.  Same tiling as simt_fwd_local.
// ----------------------------------------------------------------------------------------------
template <typename T, int HD>
__global__ void __launch_bounds__(128)
simt_bwd_dq(Geo geo, T4 q, T4 k, T4 v, T4 d_o, T4 dq, const float* __restrict__ lse,
            const float* __restrict__ delta, const float* __restrict__ table,
            const float* __restrict__ g2l, float* __restrict__ d_table) {
  using TL = Tile<HD>;
  constexpr int HH = TL::HH, HS = TL::HS;
  extern __shared__ float smem[];
  float* Ks = smem;
  float* Vs = Ks + 64 * HS;
  float* tab = Vs + 64 * HS;
  const int tw = 4 * geo.w - 1;
  const int tabn = geo.has_bias ? tw * tw : 0;
  short* kvr = reinterpret_cast<short*>(tab + tabn);
  short* kvc = kvr + 64;
  unsigned char* kfl = reinterpret_cast<unsigned char*>(kvc + 64);

  const ChunkId cid = decode_block(geo, blockIdx.x);
  const int b = cid.b, h = cid.h, R = cid.R, C = cid.C;
  const int tid = threadIdx.x, slot = tid >> 1, half = tid & 1;
  const int w = geo.w, D = geo.D;
  for (int i = tid; i < tabn; i += 128) tab[i] = table[(long long)i * geo.H + h];

  const int l = cid.piece * 64 + slot;
  const int qr = l / w, qc = l % w;
  const int r = R * w + qr, c = C * w + qc;
  const bool qvalid = (l < geo.w2) && (r < geo.nx) && (c < geo.ny);
  const long long tokq = (long long)r * geo.ny + c;

  float qh[HH], doh[HH], dqh[HH];
#pragma unroll
  for (int i = 0; i < HH; ++i) { qh[i] = 0.f; doh[i] = 0.f; dqh[i] = 0.f; }
  float lse_i = INFINITY, del_i = 0.f;
  if (qvalid) {
    load_seg<T, HH>(row_ptr<T>(q, b, h, tokq), half * HH, D, qh);
    load_seg<T, HH>(row_ptr<T>(d_o, b, h, tokq), half * HH, D, doh);
    lse_i = lse[((long long)b * geo.H + h) * geo.Nloc + tokq];
    del_i = delta[((long long)b * geo.H + h) * geo.Nloc + tokq];
  }

  const int ngp = (geo.g + 63) / 64;
  const int npieces = ngp + geo.noffs * geo.npc;
  for (int pi = 0; pi < npieces; ++pi) {
    const bool isg = pi < ngp;
    int dR = 0, dC = 0, KR = 0, KC = 0, kp = 0;
    if (!isg) {
      const int oi = (pi - ngp) / geo.npc;
      kp = (pi - ngp) % geo.npc;
      dR = geo.offR[oi]; dC = geo.offC[oi];
      KR = R + dR; KC = C + dC;
      if (geo.exact == -1) { KR = (KR + geo.mx) % geo.mx; KC = (KC + geo.my) % geo.my; }
      else if (KR < 0 || KR >= geo.mx || KC < 0 || KC >= geo.my) continue;
    }
    __syncthreads();
    {
      float kk[HH], vv[HH];
#pragma unroll
      for (int i = 0; i < HH; ++i) { kk[i] = 0.f; vv[i] = 0.f; }
      int flag = 0, vr = 0, vc = 0; long long tok = -1;
      if (isg) {
        const int t = pi * 64 + slot;
        if (t < geo.g) { flag = 2; vr = t; tok = t; }
      } else {
        const int lk = kp * 64 + slot;
        if (lk < geo.w2) {
          const int kr = lk / w, kc = lk % w;
          const int ar = KR * w + kr, ac = KC * w + kc;
          const bool real = (ar < geo.nx) && (ac < geo.ny);
          if (geo.exact == -1)
            flag = !(((R + dR == geo.mx - 1) && (kr >= w - geo.padx)) ||
                     ((C + dC == geo.my - 1) && (kc >= w - geo.pady)));
          else
            flag = real;
          if (flag && real) tok = geo.g + (long long)ar * geo.ny + ac;
          vr = dR * w + kr; vc = dC * w + kc;
        }
      }
      if (tok >= 0) {
        load_seg<T, HH>(row_ptr<T>(k, b, h, tok), half * HH, D, kk);
        load_seg<T, HH>(row_ptr<T>(v, b, h, tok), half * HH, D, vv);
      }
      float* kd = Ks + slot * HS + TL::off(half);
      float* vd = Vs + slot * HS + TL::off(half);
#pragma unroll
      for (int i = 0; i < HH; ++i) { kd[i] = kk[i]; vd[i] = vv[i]; }
      if (half == 0) { kvr[slot] = (short)vr; kvc[slot] = (short)vc; kfl[slot] = (unsigned char)flag; }
    }
    __syncthreads();
    for (int j = 0; j < 64; ++j) {
      const int f = kfl[j];
      if (!f) continue;
      const float* kd = Ks + j * HS + TL::off(half);
      const float* vd = Vs + j * HS + TL::off(half);
      float sp = 0.f, dpp = 0.f;
#pragma unroll
      for (int i = 0; i < HH; ++i) { sp = fmaf(qh[i], kd[i], sp); dpp = fmaf(doh[i], vd[i], dpp); }
      sp += __shfl_xor_sync(0xffffffffu, sp, 1);
      dpp += __shfl_xor_sync(0xffffffffu, dpp, 1);
      float bias = 0.f; bool ok = qvalid; int bidx = -1;
      if (f == 2) {
        if (geo.has_bias) bias = g2l[((long long)geo.H + h) * geo.g + kvr[j]];
      } else {
        const int dr = qr - kvr[j], dc = qc - kvc[j];
        if (geo.exact == 1 && (abs(dr) > w || abs(dc) > w)) ok = false;
        if (geo.has_bias && ok) { bidx = (dr + 2 * w - 1) * tw + dc + 2 * w - 1; bias = tab[bidx]; }
      }
      if (ok) {
        const float p = __expf(fmaf(geo.scale, sp, bias) - lse_i);
        const float ds = p * (dpp - del_i);
#pragma unroll
        for (int i = 0; i < HH; ++i) dqh[i] = fmaf(ds, kd[i], dqh[i]);
        if (bidx >= 0 && half == 0 && d_table != nullptr)
          atomicAdd(d_table + (long long)bidx * geo.H + h, ds);
      }
    }
  }
  if (qvalid) {
#pragma unroll
    for (int i = 0; i < HH; ++i) dqh[i] *= geo.scale;
    store_seg<T, HH>(row_ptr_w<T>(dq, b, h, tokq), half * HH, D, dqh);
  }
}

// ----------------------------------------------------------------------------------------------
// backward pass 2 (key-stationary): dk, dv of the LOCAL key rows.  CTA = one 64-key piece of one
// key chunk; it walks the query chunks that visit it (the symmetric image of the offset list).
// ----------------------------------------------------------------------------------------------
template <typename T, int HD>
__global__ void __launch_bounds__(128)
simt_bwd_dkv(Geo geo, T4 q, T4 k, T4 v, T4 d_o, T4 dk, T4 dv, const float* __restrict__ lse,
             const float* __restrict__ delta, const float* __restrict__ table) {
  using TL = Tile<HD>;
  constexpr int HH = TL::HH, HS = TL::HS;
  extern __shared__ float smem[];
  float* Qs = smem;
  float* Gs = Qs + 64 * HS;
  float* tab = Gs + 64 * HS;
  const int tw = 4 * geo.w - 1;
  const int tabn = geo.has_bias ? tw * tw : 0;
  float* lse_s = tab + tabn;
  float* del_s = lse_s + 64;
  short* qrs = reinterpret_cast<short*>(del_s + 64);
  short* qcs = qrs + 64;
  unsigned char* qfl = reinterpret_cast<unsigned char*>(qcs + 64);

  const ChunkId cid = decode_block(geo, blockIdx.x);
  const int b = cid.b, h = cid.h, KR = cid.R, KC = cid.C;
  const int tid = threadIdx.x, slot = tid >> 1, half = tid & 1;
  const int w = geo.w, D = geo.D;
  for (int i = tid; i < tabn; i += 128) tab[i] = table[(long long)i * geo.H + h];

  const int lk = cid.piece * 64 + slot;
  const int kr = lk / w, kc = lk % w;
  const int ar = KR * w + kr, ac = KC * w + kc;
  const bool kreal = (lk < geo.w2) && (ar < geo.nx) && (ac < geo.ny);
  const long long tokk = geo.g + (long long)ar * geo.ny + ac;

  float kh[HH], vh[HH], dkh[HH], dvh[HH];
#pragma unroll
  for (int i = 0; i < HH; ++i) { kh[i] = 0.f; vh[i] = 0.f; dkh[i] = 0.f; dvh[i] = 0.f; }
  if (kreal) {
    load_seg<T, HH>(row_ptr<T>(k, b, h, tokk), half * HH, D, kh);
    load_seg<T, HH>(row_ptr<T>(v, b, h, tokk), half * HH, D, vh);
  }

  for (int oi = 0; oi < geo.noffs; ++oi) {
    const int dR = geo.offR[oi], dC = geo.offC[oi];
    int QR = KR - dR, QC = KC - dC;
    if (geo.exact == -1) { QR = (QR + geo.mx) % geo.mx; QC = (QC + geo.my) % geo.my; }
    else if (QR < 0 || QR >= geo.mx || QC < 0 || QC >= geo.my) continue;
    // is this key visible from query chunk (QR,QC) through offset (dR,dC)?
    bool kvis = kreal;
    if (geo.exact == -1)
      kvis = kreal && !(((QR + dR == geo.mx - 1) && (kr >= w - geo.padx)) ||
                        ((QC + dC == geo.my - 1) && (kc >= w - geo.pady)));
    const int vr = dR * w + kr, vc = dC * w + kc;
    for (int qp = 0; qp < geo.npc; ++qp) {
      __syncthreads();
      {
        float qq[HH], gg[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) { qq[i] = 0.f; gg[i] = 0.f; }
        const int l = qp * 64 + slot;
        const int qr = l / w, qc = l % w;
        const int r = QR * w + qr, c = QC * w + qc;
        const bool qv = (l < geo.w2) && (r < geo.nx) && (c < geo.ny);
        if (qv) {
          const long long tq = (long long)r * geo.ny + c;
          load_seg<T, HH>(row_ptr<T>(q, b, h, tq), half * HH, D, qq);
          load_seg<T, HH>(row_ptr<T>(d_o, b, h, tq), half * HH, D, gg);
          if (half == 0) {
            lse_s[slot] = lse[((long long)b * geo.H + h) * geo.Nloc + tq];
            del_s[slot] = delta[((long long)b * geo.H + h) * geo.Nloc + tq];
          }
        }
        float* qd = Qs + slot * HS + TL::off(half);
        float* gd = Gs + slot * HS + TL::off(half);
#pragma unroll
        for (int i = 0; i < HH; ++i) { qd[i] = qq[i]; gd[i] = gg[i]; }
        if (half == 0) { qrs[slot] = (short)qr; qcs[slot] = (short)qc; qfl[slot] = (unsigned char)qv; }
      }
      __syncthreads();
      for (int i2 = 0; i2 < 64; ++i2) {
        if (!qfl[i2]) continue;                              // warp-uniform
        const float* qd = Qs + i2 * HS + TL::off(half);
        const float* gd = Gs + i2 * HS + TL::off(half);
        float sp = 0.f, dpp = 0.f;
#pragma unroll
        for (int i = 0; i < HH; ++i) { sp = fmaf(kh[i], qd[i], sp); dpp = fmaf(vh[i], gd[i], dpp); }
        sp += __shfl_xor_sync(0xffffffffu, sp, 1);
        dpp += __shfl_xor_sync(0xffffffffu, dpp, 1);
        const int dr = qrs[i2] - vr, dc = qcs[i2] - vc;
        bool ok = kvis;
        if (geo.exact == 1 && (abs(dr) > w || abs(dc) > w)) ok = false;
        if (ok) {
          const float bias = geo.has_bias ? tab[(dr + 2 * w - 1) * tw + dc + 2 * w - 1] : 0.f;
          const float p = __expf(fmaf(geo.scale, sp, bias) - lse_s[i2]);
          const float ds = p * (dpp - del_s[i2]);
#pragma unroll
          for (int i = 0; i < HH; ++i) { dkh[i] = fmaf(ds, qd[i], dkh[i]); dvh[i] = fmaf(p, gd[i], dvh[i]); }
        }
      }
    }
  }
  if (kreal) {
#pragma unroll
    for (int i = 0; i < HH; ++i) dkh[i] *= geo.scale;
    store_seg<T, HH>(row_ptr_w<T>(dk, b, h, tokk), half * HH, D, dkh);
    store_seg<T, HH>(row_ptr_w<T>(dv, b, h, tokk), half * HH, D, dvh);
  }
}

// block-wide sum of a per-thread value (256 threads), result valid in thread 0
__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

// ----------------------------------------------------------------------------------------------
// backward, global KEY columns seen by the local queries: dk[t], dv[t] for t < nglo and d_g2l[1][h][t].
// CTA = one (b, h, t); row groups stride over the local queries.
// ----------------------------------------------------------------------------------------------
template <typename T, int HD, typename TO = T>
__global__ void __launch_bounds__(256)
simt_bwd_gcol(Geo geo, T4 q, T4 k, T4 v, T4 d_o, T4 dk, T4 dv, const float* __restrict__ lse,
              const float* __restrict__ delta, const float* __restrict__ g2l, float* __restrict__ d_g2l) {
  constexpr int LPR = HD / 8, RPW = 32 / LPR, ROWS = 8 * RPW;
  __shared__ float red[8];
  __shared__ float accs[8][2][HD];
  const int t = blockIdx.x % geo.g;
  const int h = (blockIdx.x / geo.g) % geo.H;
  const int b = blockIdx.x / (geo.g * geo.H);
  const int tid = threadIdx.x, D = geo.D, lane = tid & 31, warp = tid >> 5, sub = lane % LPR, rw = lane / LPR;
  const int row0 = 0, row1 = geo.Nloc;
  float k8[8], v8[8];
  load_seg<T, 8>(row_ptr<T>(k, b, h, t), 8 * sub, D, k8);
  load_seg<T, 8>(row_ptr<T>(v, b, h, t), 8 * sub, D, v8);
  const float bias = geo.has_bias ? g2l[((long long)geo.H + h) * geo.g + t] : 0.f;
  float adk[8], adv[8], adb = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { adk[i] = 0.f; adv[i] = 0.f; }
  const long long base_l = ((long long)b * geo.H + h) * geo.Nloc;
  for (int base = row0 + warp * RPW; base < row1; base += ROWS) {
    const int i2 = base + rw;
    const bool valid = i2 < row1;
    const int ic = valid ? i2 : row1 - 1;
    float qq[8], gg[8];
    load_seg<T, 8>(row_ptr<T>(q, b, h, ic), 8 * sub, D, qq);
    load_seg<T, 8>(row_ptr<T>(d_o, b, h, ic), 8 * sub, D, gg);
    const float ls = lse[base_l + ic], dl = delta[base_l + ic];
    const float sp = group_sum<LPR>(dot8(qq, k8)), dpp = group_sum<LPR>(dot8(gg, v8));
    const float p = valid ? __expf(fmaf(geo.scale, sp, bias) - ls) : 0.f;
    const float ds = p * (dpp - dl);
    if (sub == 0) adb += ds;
#pragma unroll
    for (int i = 0; i < 8; ++i) { adk[i] = fmaf(ds, qq[i], adk[i]); adv[i] = fmaf(p, gg[i], adv[i]); }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float x = rows_sum<LPR>(adk[i]), y = rows_sum<LPR>(adv[i]);
    if (rw == 0) { accs[warp][0][8 * sub + i] = x; accs[warp][1][8 * sub + i] = y; }
  }
  const float tb = block_sum_256(adb, red);
  __syncthreads();
  if (tid < D) {
    float x = 0.f, y = 0.f;
    for (int w2 = 0; w2 < 8; ++w2) { x += accs[w2][0][tid]; y += accs[w2][1][tid]; }
    row_ptr_w<TO>(dk, b, h, t)[tid] = ElemTraits<TO>::from_f(x * geo.scale);
    row_ptr_w<TO>(dv, b, h, t)[tid] = ElemTraits<TO>::from_f(y);
  }
  if (tid == 0 && geo.has_bias && d_g2l != nullptr) atomicAdd(d_g2l + ((long long)geo.H + h) * geo.g + t, tb);
}

// ----------------------------------------------------------------------------------------------
// backward, global QUERY rows: dqg, contributions to dkg / dvg over all N keys, d_g2g, d_g2l[0].
// CTA = one (b, h); row groups stride over the keys.  `accumulate` != 0: add into dkg/dvg (they alias dk/dv,
// already written by the dK/dV pass and simt_bwd_gcol earlier on the same stream); else overwrite.
// ----------------------------------------------------------------------------------------------
template <typename T, int HD, typename TO = T>
__global__ void __launch_bounds__(256)
simt_bwd_grow(Geo geo, T4 qg, T4 kg, T4 vg, T4 d_og, T4 dqg, T4 dkg, T4 dvg,
              const float* __restrict__ lse_g, const float* __restrict__ delta_g,
              const float* __restrict__ g2l, const float* __restrict__ g2g,
              float* __restrict__ d_g2l, float* __restrict__ d_g2g, int accumulate, int rmw_rows) {
  // rmw_rows: keys [0, rmw_rows) get their dkg / dvg rows updated here (all N, or only the g global keys when the
  // tcgen05 pass 2 has already folded the global query rows into dk / dv of the local keys)
  constexpr int LPR = HD / 8, RPW = 32 / LPR, ROWS = 8 * RPW;
  __shared__ float red[8];
  __shared__ float accs[8][HD];
  const int h = blockIdx.x % geo.H;
  const int b = blockIdx.x / geo.H;
  const int tid = threadIdx.x, D = geo.D, lane = tid & 31, warp = tid >> 5, sub = lane % LPR, rw = lane / LPR;
  const int row0 = 0, row1 = geo.N;
  for (int a = 0; a < geo.g; ++a) {
    float q8[8], g8[8];
    load_seg<T, 8>(row_ptr<T>(qg, b, h, a), 8 * sub, D, q8);
    load_seg<T, 8>(row_ptr<T>(d_og, b, h, a), 8 * sub, D, g8);
    const float lg = lse_g[((long long)b * geo.H + h) * geo.g + a];
    const float dg = delta_g[((long long)b * geo.H + h) * geo.g + a];
    const float bl = geo.has_bias ? g2l[(long long)h * geo.g + a] : 0.f;
    const bool add = accumulate || a > 0;
    float adq[8], adb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) adq[i] = 0.f;
    for (int base = row0 + warp * RPW; base < row1; base += ROWS) {
      const int j = base + rw;
      const bool valid = j < row1;
      const int jc = valid ? j : row1 - 1;
      float kk[8], vv[8], ok_[8], ov_[8];
      load_seg<T, 8>(row_ptr<T>(kg, b, h, jc), 8 * sub, D, kk);
      load_seg<T, 8>(row_ptr<T>(vg, b, h, jc), 8 * sub, D, vv);
      const bool rmw = base < rmw_rows;                 // warp-uniform up to the last partial row batch
      if (add && rmw) {
        load_seg<TO, 8>(row_ptr<TO>(dkg, b, h, jc), 8 * sub, D, ok_);
        load_seg<TO, 8>(row_ptr<TO>(dvg, b, h, jc), 8 * sub, D, ov_);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { ok_[i] = 0.f; ov_[i] = 0.f; }
      }
      const float sp = group_sum<LPR>(dot8(q8, kk)), dpp = group_sum<LPR>(dot8(g8, vv));
      float bias = bl;
      if (geo.has_bias && jc < geo.g) bias = g2g[((long long)h * geo.g + a) * geo.g + jc];
      const float p = valid ? __expf(fmaf(geo.scale, sp, bias) - lg) : 0.f;
      const float ds = p * (dpp - dg);
      if (geo.has_bias && valid && sub == 0) {
        if (j < geo.g) { if (d_g2g) atomicAdd(d_g2g + ((long long)h * geo.g + a) * geo.g + j, ds); }
        else adb += ds;
      }
      const float dss = ds * geo.scale;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        adq[i] = fmaf(ds, kk[i], adq[i]);
        ok_[i] = fmaf(dss, q8[i], ok_[i]);        // dkg_j += scale * ds * qg_a
        ov_[i] = fmaf(p, g8[i], ov_[i]);          // dvg_j += p * dOg_a
      }
      if (valid && j < rmw_rows && 8 * sub < D) {
        store_seg<TO, 8>(row_ptr_w<TO>(dkg, b, h, j), 8 * sub, D, ok_);
        store_seg<TO, 8>(row_ptr_w<TO>(dvg, b, h, j), 8 * sub, D, ov_);
      }
    }
    __syncthreads();                               // accs / red of the previous global query have been consumed
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = rows_sum<LPR>(adq[i]);
      if (rw == 0) accs[warp][8 * sub + i] = x;
    }
    const float tb = block_sum_256(adb, red);
    __syncthreads();
    if (tid < D) {
      float x = 0.f;
      for (int w2 = 0; w2 < 8; ++w2) x += accs[w2][tid];
      row_ptr_w<TO>(dqg, b, h, a)[tid] = ElemTraits<TO>::from_f(x * geo.scale);
    }
    if (tid == 0 && geo.has_bias && d_g2l != nullptr) atomicAdd(d_g2l + (long long)h * geo.g + a, tb);
  }
}

// ---------------------------------------------------------------- host launchers (both kernel families)
template <typename T, int HD, typename TO = T>
inline void launch_global_fwd_kernels(const Geo& g, T4 qg, T4 kg, T4 vg, T4 og, float* lse_g, const float* g2l,
                                      const float* g2g, cudaStream_t s) {
  simt_fwd_global<T, HD, TO><<<g.B * g.H * g.g, 256, 0, s>>>(g, qg, kg, vg, og, lse_g, g2l, g2g);
}
template <typename T, int HD, typename TO = T>
inline void launch_global_bwd_kernels(const Geo& g, T4 q, T4 k, T4 v, T4 d_o, T4 dk, T4 dv, T4 qg, T4 kg, T4 vg, T4 d_og,
                                      T4 dqg, T4 dkg, T4 dvg, const float* lse, const float* delta, const float* lse_g,
                                      const float* delta_g, const float* g2l, const float* g2g, float* d_g2l,
                                      float* d_g2g, int accumulate, int rmw_rows, cudaStream_t s) {
  simt_bwd_gcol<T, HD, TO><<<g.B * g.H * g.g, 256, 0, s>>>(g, q, k, v, d_o, dk, dv, lse, delta, g2l, d_g2l);
  simt_bwd_grow<T, HD, TO><<<g.B * g.H, 256, 0, s>>>(g, qg, kg, vg, d_og, dqg, dkg, dvg, lse_g, delta_g, g2l, g2g, d_g2l, d_g2g,
                                                  accumulate, rmw_rows);
}

}  // namespace vil
