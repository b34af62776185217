// tcgen05 / TMA backward kernels of the Vision-Longformer attention (sm_100a), chunk size w <= 8 (bigger windows:
// vil_tc_big.cuh).  The bias-table gradient is the DBIAS instantiation of pass 1.
//
// Two deterministic passes (no atomics), both tiled like the forward (128-row tile = 2 chunk slots):
//   pass 1  vil_tc_bwd_dq  : query-stationary.  Per key block:  S = Q K^T, dP = dO V^T (SS MMAs into TMEM) ->
//                            threads: P = exp2(S c + bias - lse2), dS = P (dP - delta) -> bf16 dS in TMEM ->
//                            dQ += dS K (TS MMA, K tile MN-major).
//   pass 2  vil_tc_bwd_dkv : key-stationary (rows = keys).  Per query block: S^T = K Q^T, dP^T = V dO^T ->
//                            threads: P^T, dS^T (bf16, TMEM) -> dV += P^T dO, dK += dS^T Q (TS MMAs).
//                            The first block of a unit is the 16-column block of the global QUERY rows.
// S is recomputed in both passes (SlidingChunk2D.backward does the same work as slidingchunk_qk + _av + _agrad,
// slidingchunk_2d.py:234-246, on materialised score tensors).  lse2 = lse*log2(e) and delta are read from a
// chunk-ordered, 64-padded copy prepared by vil_tc_bwd_prep (invalid rows: lse2 = +inf -> P = 0).
#pragma once
#include "vil_tc_fwd.cuh"

namespace vil {
namespace tc {

constexpr int kBwdThreads = 320;    // warps 0-7 compute (2 per TMEM lane quadrant), 8 TMA producer, 9 MMA issuer

struct BwdArgs {
  Geo geo;
  T4 out0, out1;                  // pass 1: dq (out0);  pass 2: dk (out0), dv (out1)
  const float* table;             // only to build the exact-window mask table (no bias-table gradient here)
  const float* g2l;
  const float* lse2c;             // (B,H,mx,my,64) log2-domain LSE, +inf on invalid rows
  const float* deltac;            // (B,H,mx,my,64) delta, 0 on invalid rows
  float* d_table;                 // ((4w-1)^2, H) fp32, accumulated into (DBIAS variant of pass 1 only)
  int cpairs, num_units, has_tab;
  float scale_log2, scale;
  // pass 2 with the global QUERY rows folded in (w <= 8): one extra 16-column block per unit whose "queries" are the
  // global tokens; lse2g = (lse_g - g2l[0][h][a]) * log2(e) (+inf for a >= g), deltag = delta_g (0 for a >= g)
  const float* lse2g;             // (B*H, 16)
  const float* deltag;            // (B*H, 16)
  int fuse_g;
};

// token-ordered (lse, delta) -> chunk-ordered, 64-padded (lse2, delta)
__global__ void vil_tc_bwd_prep(Geo geo, const float* __restrict__ lse, const float* __restrict__ delta,
                                float* __restrict__ lse2c, float* __restrict__ deltac) {
  const long long total = (long long)geo.B * geo.H * geo.mx * geo.my * 64;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int l = (int)(idx & 63);
  long long c = idx >> 6;
  const int C = (int)(c % geo.my); c /= geo.my;
  const int R = (int)(c % geo.mx); c /= geo.mx;       // c = b*H + h
  const int r = R * geo.w + l / geo.w, cc = C * geo.w + l % geo.w;
  float a = INFINITY, d = 0.f;
  if (l < geo.w2 && r < geo.nx && cc < geo.ny) {
    const long long t = c * geo.Nloc + (long long)r * geo.ny + cc;
    a = lse[t] * 1.4426950408889634f;
    d = delta[t];
  }
  lse2c[idx] = a;
  deltac[idx] = d;
}

// global query rows -> 16-padded log2-domain (lse - bias), delta per (b, h)
__global__ void vil_tc_bwd_prep_g(Geo geo, const float* __restrict__ lse_g, const float* __restrict__ delta_g,
                                  const float* __restrict__ g2l, float* __restrict__ lse2g, float* __restrict__ deltag) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= geo.B * geo.H * 16) return;
  const int a = idx & 15, bh = idx >> 4, h = bh % geo.H;
  float l = INFINITY, d = 0.f;
  if (a < geo.g) {
    const float bias = (geo.has_bias && g2l != nullptr) ? g2l[(long long)h * geo.g + a] : 0.f;      // g2l[0][h][a]
    l = (lse_g[(long long)bh * geo.g + a] - bias) * 1.4426950408889634f;
    d = delta_g[(long long)bh * geo.g + a];
  }
  lse2g[idx] = l;
  deltag[idx] = d;
}

template <int DP>
struct BwdSmem {
  static constexpr int ROWB = DP * 2;
  static constexpr int NS = DP == 64 ? 2 : 3;            // streamed-tile ring depth
  static constexpr int X_BYTES = 128 * ROWB;             // one stationary tile
  static constexpr int Y_BYTES = 64 * ROWB;              // one streamed tile
  static constexpr int STAGE_STRIDE = (2 * Y_BYTES + 512 + 1023) / 1024 * 1024;   // two tiles + lse2/delta (2 x 64 floats)
  static constexpr int OFF_X = 0;                        // [2 buffers][2 tiles]
  static constexpr int OFF_Y = 4 * X_BYTES;
  static constexpr int OFF_TAB = OFF_Y + NS * STAGE_STRIDE;
  static __host__ __device__ int total(int tab_floats) { return OFF_TAB + tab_floats * 4 + 512 + 1024; }
};

// BB_DSFULL is a PAIR of barriers indexed by the block parity: with the early S/dP release a fast warp can be one
// block ahead of a slow one, and its arrival must not be counted towards the slow warp's (still open) phase.
enum { BB_XFULL = 0, BB_XEMPTY = 2, BB_YFULL = 4, BB_YEMPTY = 7, BB_SFULL = 10, BB_DSFULL = 11, BB_ACCDONE = 13,
       BB_ACCFREE = 14, BB_CONS = 15, BB_PDONE = 16, BB_COUNT = 17 };

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

// Walk of the QUERY chunks that visit the two key slots of a pass-2 unit: BlockWalk with the offsets mirrored.
struct QueryWalk {
  BlockWalk w;
  __device__ __forceinline__ void init(const Geo& g, int R, int Cp) { w.init(g, R, Cp, true, false); }
  __device__ __forceinline__ bool next(const Geo& g, int& QR, int& QC) { int type; return w.next(g, type, QR, QC); }
  __device__ __forceinline__ bool used_by(int slot) const { return w.used_by(slot); }
};

// store NC (16 or 32) fp32 accumulator columns [c0, c0+NC) of one row, scaled, as bf16/fp16
template <int NC, bool BF16>
__device__ __forceinline__ void store_cols(const T4& t, int b, int h, long long tok, int D, int c0, const uint32_t (&ov)[NC], float f) {
  char* base = t.p + ((long long)b * t.sb + (long long)h * t.sh + tok * t.st) * 2;
#pragma unroll
  for (int v8 = 0; v8 < NC / 8; ++v8) {
    if (c0 + v8 * 8 < D) {
      uint4 pkt;
      pkt.x = pack2<BF16>(__uint_as_float(ov[v8 * 8 + 0]) * f, __uint_as_float(ov[v8 * 8 + 1]) * f);
      pkt.y = pack2<BF16>(__uint_as_float(ov[v8 * 8 + 2]) * f, __uint_as_float(ov[v8 * 8 + 3]) * f);
      pkt.z = pack2<BF16>(__uint_as_float(ov[v8 * 8 + 4]) * f, __uint_as_float(ov[v8 * 8 + 5]) * f);
      pkt.w = pack2<BF16>(__uint_as_float(ov[v8 * 8 + 6]) * f, __uint_as_float(ov[v8 * 8 + 7]) * f);
      *reinterpret_cast<uint4*>(base + (c0 + v8 * 8) * 2) = pkt;
    }
  }
}

template <int W>
__device__ __forceinline__ void build_tables(const Geo& geo, const float* table, const float* g2l, float* tab, int tabn,
                                             float* g2l_s, int tid) {
  constexpr int TW = 4 * W - 1;
  for (int i = tid; i < geo.H * tabn; i += kBwdThreads) {
    const int h = i / tabn, idx = i % tabn;
    const int dr = idx / TW - (2 * W - 1), dc = idx % TW - (2 * W - 1);
    float v = (table != nullptr) ? table[(long long)idx * geo.H + h] * 1.4426950408889634f : 0.f;
    if (geo.exact == 1 && (abs(dr) > W || abs(dc) > W)) v = -INFINITY;
    tab[i] = v;
  }
  for (int i = tid; i < geo.H * 16; i += kBwdThreads) {
    const int h = i / 16, t = i % 16;
    g2l_s[i] = (g2l != nullptr && t < geo.g) ? g2l[((long long)geo.H + h) * geo.g + t] * 1.4426950408889634f : 0.f;
  }
}

__device__ __forceinline__ void init_bwd_barriers(uint32_t bars, int ns) {
  for (int i = 0; i < 2; ++i) { mbar_init((bars + 8u * (BB_XFULL + i)), 1); mbar_init((bars + 8u * (BB_XEMPTY + i)), 1); }
  for (int i = 0; i < ns; ++i) { mbar_init((bars + 8u * (BB_YFULL + i)), 1); mbar_init((bars + 8u * (BB_YEMPTY + i)), 1); }
  mbar_init((bars + 8u * (BB_SFULL)), 1); mbar_init((bars + 8u * (BB_DSFULL)), 256); mbar_init((bars + 8u * (BB_DSFULL + 1)), 256);
  mbar_init((bars + 8u * (BB_CONS)), 256); mbar_init((bars + 8u * (BB_PDONE)), 1);
  mbar_init((bars + 8u * (BB_ACCDONE)), 1); mbar_init((bars + 8u * (BB_ACCFREE)), 256);
  fence_barrier_init();
}

// pass-1 element work for 16 columns [COL0, COL0+16) of one local key block.  HAS_TAB / MASKED are compile-time so
// that the plain case (no bias table, interior chunk) is 4 instructions per score: FFMA, EX2, FADD, FMUL (+ 1/2 pack).
template <int W, int COL0, bool BF16, bool HAS_TAB, bool MASKED, int NV = W * W>
__device__ __forceinline__ void dq_cols16(uint32_t* __restrict__ pk, const uint32_t (&s)[16], const uint32_t (&dp)[16], float c,
                                          const float* __restrict__ tb, int krows, int kcols, float lse2, float del,
                                          float* __restrict__ e_row = nullptr) {
  constexpr int TW = 4 * W - 1, W2 = W * W;
  (void)W2;
#pragma unroll
  for (int jj = 0; jj < 16; jj += 2) {
    const int j = COL0 + jj;
    float dsv[2] = {0.f, 0.f};
    if (j < NV) {                                  // pair-wise packed math; a lone last column computes a dead lane
      float x[2], t[2], p[2];
      if constexpr (HAS_TAB) {
        const float b0 = tb[-((j / W) * TW + (j % W))];
        const float b1 = (j + 1 < NV) ? tb[-(((j + 1) / W) * TW + ((j + 1) % W))] : 0.f;
        ffma2(x[0], x[1], __uint_as_float(s[jj]), __uint_as_float(s[jj + 1]), c, c, b0 - lse2, b1 - lse2);
      } else {
        ffma2(x[0], x[1], __uint_as_float(s[jj]), __uint_as_float(s[jj + 1]), c, c, -lse2, -lse2);
      }
      p[0] = fast_exp2(x[0]);
      p[1] = (j + 1 < NV) ? fast_exp2(x[1]) : 0.f;
      if constexpr (MASKED) {
        p[0] = ((j / W) < krows && (j % W) < kcols) ? p[0] : 0.f;
        p[1] = (((j + 1) / W) < krows && ((j + 1) % W) < kcols) ? p[1] : 0.f;
      }
      fadd2(t[0], t[1], __uint_as_float(dp[jj]), __uint_as_float(dp[jj + 1]), -del, -del);
      fmul2(dsv[0], dsv[1], p[0], p[1], t[0], t[1]);
      // bias-table gradient: E[rel block][key j][query row] += dS (thread-private entry, plain RMW)
      if (HAS_TAB && e_row != nullptr) {
        e_row[j * W2] += dsv[0];
        if (j + 1 < NV) e_row[(j + 1) * W2] += dsv[1];
      }
    }
    pk[jj >> 1] = pack2<BF16>(dsv[0], dsv[1]);
  }
}
// load + process one 16-column quarter; variant chosen by two warp-uniform flags
template <int W, int COL0, bool BF16, int NV = W * W>
__device__ __forceinline__ void dq_quarter(uint32_t* __restrict__ pk, uint32_t saddr, uint32_t paddr, float c, bool has_tab,
                                           const float* __restrict__ tb, bool masked, int krows, int kcols, float lse2,
                                           float del, uint32_t cons_bar, float* __restrict__ e_row = nullptr) {
  uint32_t s[16], dp[16];
  tmem_ld_x16(saddr + COL0, s);
  tmem_ld_x16(paddr + COL0, dp);
  tmem_ld_wait();
  if (cons_bar != 0u) { tc_fence_before(); mbar_arrive(cons_bar); }     // last read of S / dP by this thread
  if (has_tab) {
    if (masked) dq_cols16<W, COL0, BF16, true, true, NV>(pk, s, dp, c, tb, krows, kcols, lse2, del, e_row);
    else        dq_cols16<W, COL0, BF16, true, false, NV>(pk, s, dp, c, tb, krows, kcols, lse2, del, e_row);
  } else {
    if (masked) dq_cols16<W, COL0, BF16, false, true, NV>(pk, s, dp, c, tb, krows, kcols, lse2, del);
    else        dq_cols16<W, COL0, BF16, false, false, NV>(pk, s, dp, c, tb, krows, kcols, lse2, del);
  }
}

// Plain case (no bias table, interior chunk) of one thread's column half: NCOL valid columns starting at the half's
// first column, processed in 8-column steps with the TMEM loads software-pipelined one step ahead (the timeline
// trace showed ~2 x 250 cycles of exposed tcgen05.ld latency per block with load -> wait -> compute per quarter).
// `tcgen05.wait::ld` waits for every outstanding load, so exactly one batch is in flight at each wait.
template <bool BF16, int NCOL>
__device__ __forceinline__ void dq_plain_pipe(uint32_t (&pk)[16], uint32_t saddr_c, uint32_t paddr_c, float c, float lse2, float del,
                                              uint32_t cons_bar) {
  constexpr int NST = (NCOL + 7) / 8;
  uint32_t s[2][8], dp[2][8];
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = 0u;
  tmem_ld_x8(saddr_c, s[0]);
  tmem_ld_x8(paddr_c, dp[0]);
#pragma unroll
  for (int i = 0; i < NST; ++i) {
    tmem_ld_wait();
    if (i + 1 < NST) {
      tmem_ld_x8(saddr_c + 8 * (i + 1), s[(i + 1) & 1]);
      tmem_ld_x8(paddr_c + 8 * (i + 1), dp[(i + 1) & 1]);
    } else {
      tc_fence_before();
      mbar_arrive(cons_bar);                        // last read of S / dP by this thread
    }
#pragma unroll
    for (int jj = 0; jj < 8; jj += 2) {
      const int j = 8 * i + jj;
      if (j < NCOL) {
        float x[2], t[2], p[2], v[2];
        ffma2(x[0], x[1], __uint_as_float(s[i & 1][jj]), __uint_as_float(s[i & 1][jj + 1]), c, c, -lse2, -lse2);
        p[0] = fast_exp2(x[0]);
        p[1] = (j + 1 < NCOL) ? fast_exp2(x[1]) : 0.f;
        fadd2(t[0], t[1], __uint_as_float(dp[i & 1][jj]), __uint_as_float(dp[i & 1][jj + 1]), -del, -del);
        fmul2(v[0], v[1], p[0], p[1], t[0], t[1]);
        pk[j >> 1] = pack2<BF16>(v[0], v[1]);
      }
    }
  }
}

// Unit enumeration shared by the three warp roles.  Plain: unit = blockIdx.x + k*gridDim.x over (b,h,R,Cp).
// Head-affine (bias-gradient variant): CTA c only sees head c % H, so its E accumulator never mixes heads.
struct UnitIter {
  int k, h_fixed, rank, ncta_h;
  bool affine;
  __device__ __forceinline__ void init(const Geo& g, bool affine_) {
    k = 0; affine = affine_;
    h_fixed = blockIdx.x % g.H; rank = blockIdx.x / g.H;
    ncta_h = ((int)gridDim.x - h_fixed + g.H - 1) / g.H;
  }
  __device__ __forceinline__ bool next(const Geo& g, int cpairs, int num_units, int& b, int& h, int& R, int& Cp) {
    const int per_img = g.mx * cpairs;
    if (!affine) {
      const int unit = blockIdx.x + k * gridDim.x;
      if (unit >= num_units) return false;
      const int bh = unit / per_img, rem = unit % per_img;
      b = bh / g.H; h = bh % g.H; R = rem / cpairs; Cp = rem % cpairs;
    } else {
      const int u = rank + k * ncta_h;
      if (u >= g.B * per_img) return false;
      b = u / per_img; h = h_fixed;
      const int rem = u % per_img;
      R = rem / cpairs; Cp = rem % cpairs;
    }
    ++k;
    return true;
  }
};

// ======================================================================================================== pass 1
// DBIAS = true: additionally accumulates the bias-table gradient.  Every (relative chunk offset, key, query row)
// triple is owned by one thread at a time, so E lives in shared memory and is updated with plain read-modify-writes;
// the block pipeline is serialised (no early S/dP release) so the two slots never touch the same E entry concurrently.
template <int DP, int W, bool BF16, bool DBIAS>
__global__ void __launch_bounds__(kBwdThreads, DBIAS ? 1 : 2)
vil_tc_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmKg, const __grid_constant__ CUtensorMap tmVg, const BwdArgs a) {
  using SM = BwdSmem<DP>;
  constexpr int ROWB = SM::ROWB, NS = SM::NS;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int W2 = W * W, TW = 4 * W - 1;
  const Geo& geo = a.geo;

  extern __shared__ unsigned char smem_raw[];
  // pointer arithmetic on the __shared__ symbol (no integer round trip) keeps the address space visible to nvcc: LDS / STS
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sX = smem + SM::OFF_X;                 // [buf][Q | dO]
  unsigned char* sY = smem + SM::OFF_Y;
  float* tab = reinterpret_cast<float*>(smem + SM::OFF_TAB);
  const int tabn = a.has_tab ? TW * TW : 0;
  float* g2l_s = tab + geo.H * tabn;
  const int bars_off = (SM::OFF_TAB + (geo.H * tabn + geo.H * 16) * 4 + 15) & ~15;
  uint64_t* bars_p = reinterpret_cast<uint64_t*>(smem + bars_off);
  const uint32_t bars = smem_u32(bars_p);                   // shared-space address; barrier i lives at bars + 8 i
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_p + BB_COUNT);
  float* E = reinterpret_cast<float*>(smem + ((bars_off + BB_COUNT * 8 + 16 + 15) & ~15));   // [9][W2][W2]
  float* bins = E + 9 * W2 * W2;                                                                           // [TW*TW]
  const int tid = threadIdx.x, warp = tid >> 5;
  if constexpr (DBIAS) {
    for (int i = tid; i < 9 * W2 * W2 + TW * TW; i += kBwdThreads) E[i] = 0.f;
  }

  for (int i = tid; i < SM::OFF_TAB / 16; i += kBwdThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  build_tables<W>(geo, a.table, a.g2l, tab, tabn, g2l_s, tid);
  if (tid == 0) init_bwd_barriers(bars, NS);
  if (warp == 8) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // S / dP are released as soon as the compute threads hold them in registers; dS has its own double buffer
  const uint32_t TM_S = tmem, TM_DP = tmem + 64, TM_DS = tmem + 128, TM_ACC = tmem + 192;
  const int units_per_bh = geo.mx * a.cpairs;

  if (warp == 8) {
    // ================================================================= TMA producer
    if (elect_one()) {
      uint32_t stage = 0, yphase = 0, uc = 0;
      UnitIter ui; ui.init(geo, DBIAS);
      int b, h, R, Cp;
      for (; ui.next(geo, a.cpairs, a.num_units, b, h, R, Cp); ++uc) {
        const uint32_t xb = uc & 1, xphase = (uc >> 1) & 1;
        if (uc >= 2) mbar_wait((bars + 8u * (BB_XEMPTY + xb)), xphase ^ 1);
        unsigned char* sQ = sX + xb * 2 * SM::X_BYTES;
        unsigned char* sDO = sQ + SM::X_BYTES;
        const bool hasB = 2 * Cp + 1 < geo.my;
        mbar_arrive_expect_tx((bars + 8u * (BB_XFULL + xb)), (hasB ? 4 : 2) * W2 * ROWB);
        tma_load_5d(sQ, &tmQ, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp) * W, R * W, h, b);
        tma_load_5d(sDO, &tmDO, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp) * W, R * W, h, b);
        if (hasB) {
          tma_load_5d(sQ + 64 * ROWB, &tmQ, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp + 1) * W, R * W, h, b);
          tma_load_5d(sDO + 64 * ROWB, &tmDO, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp + 1) * W, R * W, h, b);
        }
        BlockWalk wk; wk.init(geo, R, Cp);
        int type, KR, KC;
        while (wk.next(geo, type, KR, KC)) {
          mbar_wait((bars + 8u * (BB_YEMPTY + stage)), yphase ^ 1);
          unsigned char* dK = sY + stage * SM::STAGE_STRIDE;
          unsigned char* dV = dK + SM::Y_BYTES;
          if (type == 1) {
            mbar_arrive_expect_tx((bars + 8u * (BB_YFULL + stage)), 2 * 16 * ROWB);
            tma_load_4d(dK, &tmKg, (bars + 8u * (BB_YFULL + stage)), 0, 0, h, b);
            tma_load_4d(dV, &tmVg, (bars + 8u * (BB_YFULL + stage)), 0, 0, h, b);
          } else {
            mbar_arrive_expect_tx((bars + 8u * (BB_YFULL + stage)), 2 * W2 * ROWB);
            tma_load_5d(dK, &tmK, (bars + 8u * (BB_YFULL + stage)), 0, KC * W, KR * W, h, b);
            tma_load_5d(dV, &tmV, (bars + 8u * (BB_YFULL + stage)), 0, KC * W, KR * W, h, b);
          }
          if (++stage == NS) { stage = 0; yphase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ================================================================= MMA issuer
    if (elect_one()) {
      constexpr uint32_t IDESC_S = make_idesc(128, 64, BF16, false, false);
      constexpr uint32_t IDESC_SG = make_idesc(128, 16, BF16, false, false);
      constexpr uint32_t IDESC_ACC = make_idesc(128, DP, BF16, false, true);
      uint32_t stage = 0, yphase = 0, uc = 0, G = 0;
      VIL_TRACE_DECL(2)
      UnitIter ui; ui.init(geo, DBIAS);
      int b, h, R, Cp;
      for (; ui.next(geo, a.cpairs, a.num_units, b, h, R, Cp); ++uc) {
        const uint32_t xb = uc & 1, xphase = (uc >> 1) & 1;
        mbar_wait((bars + 8u * (BB_XFULL + xb)), xphase);
        const uint32_t qaddr = smem_u32(sX + xb * 2 * SM::X_BYTES), doaddr = qaddr + SM::X_BYTES;
        // descriptors are built BEFORE the barrier waits, so that only the tcgen05.mma issues sit between a barrier
        // completing and the next S / dP being under way (this thread's latency is on the block critical path)
        constexpr int KS = DP / 16;
        uint64_t qd[KS], dod[KS], kd[KS], vd[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          qd[k] = make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT);
          dod[k] = make_smem_desc(doaddr + k * 32, 16, SBO, LAYOUT);
        }
        auto prep_SdP = [&](uint32_t st) {
          const uint32_t kaddr = smem_u32(sY + st * SM::STAGE_STRIDE), vaddr = kaddr + SM::Y_BYTES;
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            kd[k] = make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT);
            vd[k] = make_smem_desc(vaddr + k * 32, 16, SBO, LAYOUT);
          }
        };
        auto issue_SdP = [&](int type) {
          const uint32_t idesc = type == 1 ? IDESC_SG : IDESC_S;
#pragma unroll
          for (int k = 0; k < KS; ++k) mma_ss(TM_S, qd[k], kd[k], idesc, k > 0);
#pragma unroll
          for (int k = 0; k < KS; ++k) mma_ss(TM_DP, dod[k], vd[k], idesc, k > 0);
          mma_commit((bars + 8u * (BB_SFULL)));
        };
        BlockWalk wk; wk.init(geo, R, Cp);
        int type, KR, KC;
        bool have = wk.next(geo, type, KR, KC);
        prep_SdP(stage);
        mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase);
        tc_fence_after();
        issue_SdP(type);
        bool first = true;
        while (have) {
          const uint32_t cur_stage = stage;
          const int cur_type = type;
          uint64_t kacc[4];                                  // B operand of dQ += dS K: the K tile of block j, MN-major
          {
            const uint32_t kaddr = smem_u32(sY + cur_stage * SM::STAGE_STRIDE);
#pragma unroll
            for (int k = 0; k < 4; ++k) kacc[k] = make_smem_desc(kaddr + k * 16 * ROWB, 16, SBO, LAYOUT);
          }
          if (++stage == NS) { stage = 0; yphase ^= 1; }
          have = wk.next(geo, type, KR, KC);
          if (have) { prep_SdP(stage); mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase); }
          if (have && !DBIAS) {
            VIL_TR(10);
            mbar_wait((bars + 8u * (BB_CONS)), G & 1);                // S_j / dP_j are in the threads' registers
            VIL_TR(11);
            tc_fence_after();
            issue_SdP(type);                                 // overlaps the threads' exp / dS work on block j
            VIL_TR(12);
          }
          mbar_wait((bars + 8u * (BB_DSFULL + (G & 1))), (G >> 1) & 1);
          VIL_TR(13);
          if (first && uc > 0) mbar_wait((bars + 8u * (BB_ACCFREE)), (uc - 1) & 1);
          tc_fence_after();
          const uint32_t dsaddr = TM_DS + (G & 1) * 32;
          if (cur_type == 1) {
            mma_ts(TM_ACC, dsaddr, kacc[0], IDESC_ACC, !first);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ts(TM_ACC, dsaddr + k * 8, kacc[k], IDESC_ACC, (!first) || k > 0);
          }
          mma_commit((bars + 8u * (BB_YEMPTY + cur_stage)));
          VIL_TR(14);
          first = false;
          ++G;
          if (have && DBIAS) issue_SdP(type);                // serialised: every thread has finished block j
          if (!have) {
            mma_commit((bars + 8u * (BB_ACCDONE)));
            mma_commit((bars + 8u * (BB_XEMPTY + xb)));
          }
        }
      }
    }
  } else {
    // ================================================================= compute warps: thread = (query row, column half)
    const int row = tid & 127, half = tid >> 7, slot = row >> 6, l = row & 63;
    const int qr = l / W, qc = l % W;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t uc = 0, G = 0;
    VIL_TRACE_DECL(tid == 0 ? 0 : (tid == 128 ? 1 : -1))
    UnitIter ui; ui.init(geo, DBIAS);
    int b, h, R, Cp;
    for (; ui.next(geo, a.cpairs, a.num_units, b, h, R, Cp); ++uc) {
      const int bh = b * geo.H + h;
      const int C = 2 * Cp + slot;
      const int r = R * W + qr, c = C * W + qc;
      const bool slot_ok = C < geo.my;
      const bool row_ok = slot_ok && l < W2 && r < geo.nx && c < geo.ny;
      float lse2 = INFINITY, del = 0.f;
      if (slot_ok) {
        const long long ci = (((long long)bh * geo.mx + R) * geo.my + C) * 64 + l;
        lse2 = a.lse2c[ci]; del = a.deltac[ci];
      }
      const float* tab_h = tab + h * tabn;
      BlockWalk wk; wk.init(geo, R, Cp);
      int type, KR, KC;
      while (wk.next(geo, type, KR, KC)) {
        VIL_TR(1);
        mbar_wait((bars + 8u * (BB_SFULL)), G & 1);
        VIL_TR(2);
        tc_fence_after();
        const uint32_t saddr = TM_S + lane_base, paddr = TM_DP + lane_base;
        const uint32_t dsaddr = TM_DS + (G & 1) * 32 + lane_base;
        if (type == 1) {
          uint32_t s[16], dp[16], pk[8];
          if (half == 0) {
            tmem_ld_x16(saddr, s);
            tmem_ld_x16(paddr, dp);
            tmem_ld_wait();
          }
          tc_fence_before();
          mbar_arrive((bars + 8u * (BB_CONS)));
          if (half == 0) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float d0 = 0.f, d1 = 0.f;
              if (j < geo.g) {
                const float p = fast_exp2(fmaf(__uint_as_float(s[j]), a.scale_log2, g2l_s[h * 16 + j]) - lse2);
                d0 = p * (__uint_as_float(dp[j]) - del);
              }
              if (j + 1 < geo.g) {
                const float p = fast_exp2(fmaf(__uint_as_float(s[j + 1]), a.scale_log2, g2l_s[h * 16 + j + 1]) - lse2);
                d1 = p * (__uint_as_float(dp[j + 1]) - del);
              }
              pk[j >> 1] = pack2<BF16>(d0, d1);
            }
            tmem_st_x8(dsaddr, pk);
          }
        } else {
          const int dR = KR - R, dC = KC - C;
          const bool use = wk.used_by(slot);
          const int krows = min(W, geo.nx - KR * W), kcols = min(W, geo.ny - KC * W);
          const bool masked = (krows < W) || (kcols < W);
          const bool ht = a.has_tab != 0;
          if (!use) {
            uint32_t pk[16];
            tc_fence_before();
            mbar_arrive((bars + 8u * (BB_CONS)));
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
            tmem_st_x16(dsaddr + half * 16, pk);
          } else if (!DBIAS && !ht && !masked) {
            uint32_t pk[16];
            constexpr int N0 = W2 < 32 ? W2 : 32, N1 = W2 - N0;
            if (half == 0) dq_plain_pipe<BF16, N0>(pk, saddr, paddr, a.scale_log2, lse2, del, (bars + 8u * (BB_CONS)));
            else           dq_plain_pipe<BF16, N1>(pk, saddr + 32, paddr + 32, a.scale_log2, lse2, del, (bars + 8u * (BB_CONS)));
            tmem_st_x16(dsaddr + half * 16, pk);
          } else {
            uint32_t pk[16];
            const float* tb = tab_h + ((qr - dR * W + 2 * W - 1) * TW + (qc - dC * W + 2 * W - 1));
            float* e_row = nullptr;
            if constexpr (DBIAS) { if (l < W2) e_row = E + ((dR + 1) * 3 + (dC + 1)) * W2 * W2 + l; }
            if (half == 0) {
              dq_quarter<W, 0, BF16>(pk, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, lse2, del, 0u, e_row);
              dq_quarter<W, 16, BF16>(pk + 8, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, lse2, del, (bars + 8u * (BB_CONS)), e_row);
            } else {
              dq_quarter<W, 32, BF16>(pk, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, lse2, del, 0u, e_row);
              dq_quarter<W, 48, BF16>(pk + 8, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, lse2, del, (bars + 8u * (BB_CONS)), e_row);
            }
            tmem_st_x16(dsaddr + half * 16, pk);
          }
        }
        VIL_TR(3);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive((bars + 8u * (BB_DSFULL + (G & 1))));
        VIL_TR(4);
        ++G;
      }
      VIL_TR(5);
      mbar_wait((bars + 8u * (BB_ACCDONE)), uc & 1);
      VIL_TR(6);
      tc_fence_after();
      constexpr int NC = DP / 2;
      uint32_t ov[NC];
      if constexpr (NC == 32) tmem_ld_x32(TM_ACC + lane_base + half * NC, ov); else tmem_ld_x16(TM_ACC + lane_base + half * NC, ov);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive((bars + 8u * (BB_ACCFREE)));
      if (row_ok) store_cols<NC, BF16>(a.out0, b, h, (long long)r * geo.ny + c, geo.D, half * NC, ov, a.scale);
      VIL_TR(7);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
  if constexpr (DBIAS) {
    // E[(dR,dC)][key j][query l] -> bins[(dr + 2W-1)*TW + dc + 2W-1] (shared atomics), then one global atomic per bin
    for (int e = tid; e < 9 * W2 * W2; e += kBwdThreads) {
      const float v = E[e];
      if (v != 0.f) {
        const int rel = e / (W2 * W2), j = (e / W2) % W2, l2 = e % W2;
        const int dR = rel / 3 - 1, dC = rel % 3 - 1;
        const int dr = l2 / W - (dR * W + j / W), dc = l2 % W - (dC * W + j % W);
        atomicAdd(&bins[(dr + 2 * W - 1) * TW + dc + 2 * W - 1], v);
      }
    }
    __syncthreads();
    const int hfix = blockIdx.x % geo.H;
    for (int i = tid; i < TW * TW; i += kBwdThreads)
      if (bins[i] != 0.f) atomicAdd(a.d_table + (long long)i * geo.H + hfix, bins[i]);
  }
}

// pass-2 element work for 16 query columns [COL0, COL0+16) of one query block (thread = key row).
// lse2 / delta of the queries come from shared memory as float4 broadcasts.
template <int W, int COL0, bool BF16, bool HAS_TAB, int NV = W * W>
__device__ __forceinline__ void dkv_cols16(uint32_t* __restrict__ pp, uint32_t* __restrict__ pd, const uint32_t (&s)[16],
                                           const uint32_t (&dp)[16], float c, const float* __restrict__ tb,
                                           const float* __restrict__ ls, const float* __restrict__ dl) {
  constexpr int TW = 4 * W - 1;
#pragma unroll
  for (int jj = 0; jj < 16; jj += 4) {
    float pv[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
    if (COL0 + jj < NV) {
      const float4 l4 = *reinterpret_cast<const float4*>(ls + COL0 + jj);      // +inf for invalid queries
      const float4 d4 = *reinterpret_cast<const float4*>(dl + COL0 + jj);
      const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        const int j = COL0 + jj + e;
        if (j < NV) {                              // pair-wise packed math; a lone last column computes a dead lane
          float x[2], t[2];
          if constexpr (HAS_TAB) {
            const float b0 = tb[(j / W) * TW + (j % W)];
            const float b1 = (j + 1 < NV) ? tb[((j + 1) / W) * TW + ((j + 1) % W)] : 0.f;
            ffma2(x[0], x[1], __uint_as_float(s[jj + e]), __uint_as_float(s[jj + e + 1]), c, c, b0 - lv[e], b1 - lv[e + 1]);
          } else {
            ffma2(x[0], x[1], __uint_as_float(s[jj + e]), __uint_as_float(s[jj + e + 1]), c, c, -lv[e], -lv[e + 1]);
          }
          pv[e] = fast_exp2(x[0]);
          pv[e + 1] = (j + 1 < NV) ? fast_exp2(x[1]) : 0.f;
          // an invalid query column (lse2 = +inf) may index past the table (short last piece of a w > 8 chunk):
          // force its probability to zero so that a garbage table word cannot poison the whole column
          if constexpr (HAS_TAB) {
            pv[e] = (lv[e] < INFINITY) ? pv[e] : 0.f;
            pv[e + 1] = (lv[e + 1] < INFINITY) ? pv[e + 1] : 0.f;
          }
          fadd2(t[0], t[1], __uint_as_float(dp[jj + e]), __uint_as_float(dp[jj + e + 1]), -dd[e], -dd[e + 1]);
          fmul2(dv[e], dv[e + 1], pv[e], pv[e + 1], t[0], t[1]);
        }
      }
    }
    pp[jj >> 1] = pack2<BF16>(pv[0], pv[1]); pp[(jj >> 1) + 1] = pack2<BF16>(pv[2], pv[3]);
    pd[jj >> 1] = pack2<BF16>(dv[0], dv[1]); pd[(jj >> 1) + 1] = pack2<BF16>(dv[2], dv[3]);
  }
}
template <int W, int COL0, bool BF16, int NV = W * W>
__device__ __forceinline__ void dkv_quarter(uint32_t* __restrict__ pp, uint32_t* __restrict__ pd, uint32_t saddr, uint32_t paddr,
                                            float c, bool has_tab, const float* __restrict__ tb, bool use,
                                            const float* __restrict__ ls, const float* __restrict__ dl, uint32_t cons_bar) {
  uint32_t s[16], dp[16];
  tmem_ld_x16(saddr + COL0, s);
  tmem_ld_x16(paddr + COL0, dp);
  tmem_ld_wait();
  if (cons_bar != 0u) { tc_fence_before(); mbar_arrive(cons_bar); }
  if (!use) {                      // padding key row of a visited chunk: contributes nothing, is never stored
#pragma unroll
    for (int j = 0; j < 8; ++j) { pp[j] = 0u; pd[j] = 0u; }
    return;
  }
  if (has_tab) dkv_cols16<W, COL0, BF16, true, NV>(pp, pd, s, dp, c, tb, ls, dl);
  else         dkv_cols16<W, COL0, BF16, false, NV>(pp, pd, s, dp, c, tb, ls, dl);
}

// ======================================================================================================== pass 2
template <int DP, int W, bool BF16>
__global__ void __launch_bounds__(kBwdThreads, 2)
vil_tc_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const __grid_constant__ CUtensorMap tmQg, const __grid_constant__ CUtensorMap tmDOg, const BwdArgs a) {
  using SM = BwdSmem<DP>;
  constexpr int ROWB = SM::ROWB, NS = SM::NS;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int W2 = W * W, TW = 4 * W - 1;
  const Geo& geo = a.geo;

  extern __shared__ unsigned char smem_raw[];
  // pointer arithmetic on the __shared__ symbol (no integer round trip) keeps the address space visible to nvcc: LDS / STS
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sX = smem + SM::OFF_X;                 // [buf][K | V]
  unsigned char* sY = smem + SM::OFF_Y;
  float* tab = reinterpret_cast<float*>(smem + SM::OFF_TAB);
  const int tabn = a.has_tab ? TW * TW : 0;
  float* g2l_s = tab + geo.H * tabn;
  const int bars_off = (SM::OFF_TAB + (geo.H * tabn + geo.H * 16) * 4 + 15) & ~15;
  uint64_t* bars_p = reinterpret_cast<uint64_t*>(smem + bars_off);
  const uint32_t bars = smem_u32(bars_p);                   // shared-space address; barrier i lives at bars + 8 i
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_p + BB_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < SM::OFF_TAB / 16; i += kBwdThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  build_tables<W>(geo, a.table, a.g2l, tab, tabn, g2l_s, tid);
  if (tid == 0) init_bwd_barriers(bars, NS);
  if (warp == 8) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // DP == 32: P^T / dS^T get their own columns so S^T / dP^T can be released early (256 columns in total);
  // DP == 64: no room -> P^T / dS^T overwrite S^T / dP^T in place and the block pipeline is serialised.
  constexpr bool kSplit = (DP == 32);
  const uint32_t TM_S = tmem, TM_DP = tmem + 64;
  const uint32_t TM_P = kSplit ? tmem + 128 : TM_S, TM_DS = kSplit ? tmem + 160 : TM_DP;
  const uint32_t TM_DK = kSplit ? tmem + 192 : tmem + 128, TM_DV = kSplit ? tmem + 224 : tmem + 192;
  const int units_per_bh = geo.mx * a.cpairs;

  if (warp == 8) {
    if (elect_one()) {
      uint32_t stage = 0, yphase = 0, uc = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh = unit / units_per_bh, rem = unit % units_per_bh;
        const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
        const uint32_t xb = uc & 1, xphase = (uc >> 1) & 1;
        if (uc >= 2) mbar_wait((bars + 8u * (BB_XEMPTY + xb)), xphase ^ 1);
        unsigned char* sK = sX + xb * 2 * SM::X_BYTES;
        unsigned char* sV = sK + SM::X_BYTES;
        const bool hasB = 2 * Cp + 1 < geo.my;
        mbar_arrive_expect_tx((bars + 8u * (BB_XFULL + xb)), (hasB ? 4 : 2) * W2 * ROWB);
        tma_load_5d(sK, &tmK, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp) * W, R * W, h, b);
        tma_load_5d(sV, &tmV, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp) * W, R * W, h, b);
        if (hasB) {
          tma_load_5d(sK + 64 * ROWB, &tmK, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp + 1) * W, R * W, h, b);
          tma_load_5d(sV + 64 * ROWB, &tmV, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp + 1) * W, R * W, h, b);
        }
        QueryWalk wk; wk.init(geo, R, Cp);
        int QR, QC;
        bool gpend = a.fuse_g != 0;                          // first block of a unit: the global query rows
        while (gpend || wk.next(geo, QR, QC)) {
          mbar_wait((bars + 8u * (BB_YEMPTY + stage)), yphase ^ 1);
          unsigned char* dQ = sY + stage * SM::STAGE_STRIDE;
          unsigned char* dG = dQ + SM::Y_BYTES;
          unsigned char* dL = dG + SM::Y_BYTES;
          if (gpend) {
            gpend = false;
            mbar_arrive_expect_tx((bars + 8u * (BB_YFULL + stage)), 2 * 16 * ROWB + 128);
            tma_load_4d(dQ, &tmQg, (bars + 8u * (BB_YFULL + stage)), 0, 0, h, b);
            tma_load_4d(dG, &tmDOg, (bars + 8u * (BB_YFULL + stage)), 0, 0, h, b);
            bulk_load_1d(dL, a.lse2g + (long long)bh * 16, 64, (bars + 8u * (BB_YFULL + stage)));
            bulk_load_1d(dL + 256, a.deltag + (long long)bh * 16, 64, (bars + 8u * (BB_YFULL + stage)));
          } else {
            mbar_arrive_expect_tx((bars + 8u * (BB_YFULL + stage)), 2 * W2 * ROWB + 512);
            tma_load_5d(dQ, &tmQ, (bars + 8u * (BB_YFULL + stage)), 0, QC * W, QR * W, h, b);
            tma_load_5d(dG, &tmDO, (bars + 8u * (BB_YFULL + stage)), 0, QC * W, QR * W, h, b);
            const long long ci = (((long long)bh * geo.mx + QR) * geo.my + QC) * 64;
            bulk_load_1d(dL, a.lse2c + ci, 256, (bars + 8u * (BB_YFULL + stage)));
            bulk_load_1d(dL + 256, a.deltac + ci, 256, (bars + 8u * (BB_YFULL + stage)));
          }
          if (++stage == NS) { stage = 0; yphase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t IDESC_S = make_idesc(128, 64, BF16, false, false);
      constexpr uint32_t IDESC_SG = make_idesc(128, 16, BF16, false, false);
      constexpr uint32_t IDESC_ACC = make_idesc(128, DP, BF16, false, true);
      uint32_t stage = 0, yphase = 0, uc = 0, G = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int rem = unit % units_per_bh;
        const int R = rem / a.cpairs, Cp = rem % a.cpairs;
        const uint32_t xb = uc & 1, xphase = (uc >> 1) & 1;
        mbar_wait((bars + 8u * (BB_XFULL + xb)), xphase);
        const uint32_t kaddr = smem_u32(sX + xb * 2 * SM::X_BYTES), vaddr = kaddr + SM::X_BYTES;
        // descriptors are built BEFORE the barrier waits (see the pass-1 issuer)
        constexpr int KS = DP / 16;
        uint64_t kd[KS], vd[KS], qd[KS], gd[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          kd[k] = make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT);
          vd[k] = make_smem_desc(vaddr + k * 32, 16, SBO, LAYOUT);
        }
        auto prep_SdP = [&](uint32_t st) {
          const uint32_t qaddr = smem_u32(sY + st * SM::STAGE_STRIDE), gaddr = qaddr + SM::Y_BYTES;
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            qd[k] = make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT);
            gd[k] = make_smem_desc(gaddr + k * 32, 16, SBO, LAYOUT);
          }
        };
        auto issue_SdP = [&](bool glob) {
          const uint32_t idesc = glob ? IDESC_SG : IDESC_S;
#pragma unroll
          for (int k = 0; k < KS; ++k) mma_ss(TM_S, kd[k], qd[k], idesc, k > 0);
#pragma unroll
          for (int k = 0; k < KS; ++k) mma_ss(TM_DP, vd[k], gd[k], idesc, k > 0);
          mma_commit((bars + 8u * (BB_SFULL)));
        };
        QueryWalk wk; wk.init(geo, R, Cp);
        int QR, QC;
        bool glob = a.fuse_g != 0;                           // block type of the S / dP being issued next
        bool have = glob ? true : wk.next(geo, QR, QC);
        prep_SdP(stage);
        mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase);
        tc_fence_after();
        issue_SdP(glob);
        bool first = true;
        while (have) {
          const uint32_t cur_stage = stage;
          const bool cur_glob = glob;
          glob = false;
          uint64_t qacc[4], gacc[4];                         // B operands of dK += dS^T Q and dV += P^T dO (block j)
          {
            const uint32_t qaddr = smem_u32(sY + cur_stage * SM::STAGE_STRIDE), gaddr = qaddr + SM::Y_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              qacc[k] = make_smem_desc(qaddr + k * 16 * ROWB, 16, SBO, LAYOUT);
              gacc[k] = make_smem_desc(gaddr + k * 16 * ROWB, 16, SBO, LAYOUT);
            }
          }
          if (++stage == NS) { stage = 0; yphase ^= 1; }
          have = wk.next(geo, QR, QC);
          if (have) { prep_SdP(stage); mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase); }
          if (kSplit && have) {
            mbar_wait((bars + 8u * (BB_CONS)), G & 1);                // S^T_j / dP^T_j are in the threads' registers
            tc_fence_after();
            issue_SdP(false);
          }
          mbar_wait((bars + 8u * (BB_DSFULL + (G & 1))), (G >> 1) & 1);
          if (first && uc > 0) mbar_wait((bars + 8u * (BB_ACCFREE)), (uc - 1) & 1);
          tc_fence_after();
          if (cur_glob) {                   // 16 global query rows: one K = 16 step each
            mma_ts(TM_DV, TM_P, gacc[0], IDESC_ACC, !first);
            mma_ts(TM_DK, TM_DS, qacc[0], IDESC_ACC, !first);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)       // dV += P^T dO
              mma_ts(TM_DV, TM_P + k * 8, gacc[k], IDESC_ACC, (!first) || k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)       // dK += dS^T Q
              mma_ts(TM_DK, TM_DS + k * 8, qacc[k], IDESC_ACC, (!first) || k > 0);
          }
          mma_commit((bars + 8u * (BB_YEMPTY + cur_stage)));
          if (kSplit) mma_commit((bars + 8u * (BB_PDONE)));           // P^T / dS^T columns may be rewritten
          first = false;
          ++G;
          if (have) {
            if (!kSplit) issue_SdP(false);
          } else {
            mma_commit((bars + 8u * (BB_ACCDONE)));
            mma_commit((bars + 8u * (BB_XEMPTY + xb)));
          }
        }
      }
    }
  } else {
    const int row = tid & 127, half = tid >> 7, slot = row >> 6, l = row & 63;
    const int kr = l / W, kc = l % W;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t uc = 0, G = 0, stage = 0, yphase = 0;
    VIL_TRACE_DECL(tid == 0 ? 0 : (tid == 128 ? 1 : -1))
    for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
      const int bh = unit / units_per_bh, rem = unit % units_per_bh;
      const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
      const int C = 2 * Cp + slot;
      const int r = R * W + kr, c = C * W + kc;
      const bool slot_ok = C < geo.my;
      const bool row_ok = slot_ok && l < W2 && r < geo.nx && c < geo.ny;
      const float* tab_h = tab + h * tabn;
      QueryWalk wk; wk.init(geo, R, Cp);
      int QR = 0, QC = 0;
      bool gpend = a.fuse_g != 0;                             // first block of a unit: the global query rows
      while (gpend || wk.next(geo, QR, QC)) {
        const bool glob = gpend;
        gpend = false;
        VIL_TR(1);
        mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase);     // lse2 / delta of this query block have landed
        mbar_wait((bars + 8u * (BB_SFULL)), G & 1);
        VIL_TR(2);
        tc_fence_after();
        const float* ls = reinterpret_cast<const float*>(sY + stage * SM::STAGE_STRIDE + 2 * SM::Y_BYTES);
        const float* dl = ls + 64;
        const uint32_t saddr = TM_S + lane_base, paddr = TM_DP + lane_base;
        const int dR = R - QR, dC = C - QC;       // offset = key chunk - query chunk
        const bool use_w = wk.used_by(slot);                        // warp-uniform
        const bool use = use_w && row_ok;
        uint32_t pp[16], pd[16];
        if (glob) {
          // 16 columns = the global QUERY rows (all keys attend to them; their bias is folded into lse2g): column half 0
          // does the math, half 1 only keeps the barrier protocol.  Replaces the dk/dv read-modify-write of simt_bwd_grow.
          uint32_t s[16], dp[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { pp[j] = 0u; pd[j] = 0u; s[j] = 0u; dp[j] = 0u; }
          if (half == 0) {
            tmem_ld_x16(saddr, s);
            tmem_ld_x16(paddr, dp);
            tmem_ld_wait();
          }
          if (kSplit) { tc_fence_before(); mbar_arrive((bars + 8u * (BB_CONS))); }
          if (half == 0 && row_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float p0 = fast_exp2(fmaf(__uint_as_float(s[j]), a.scale_log2, -ls[j]));
              const float p1 = fast_exp2(fmaf(__uint_as_float(s[j + 1]), a.scale_log2, -ls[j + 1]));
              pp[j >> 1] = pack2<BF16>(p0, p1);
              pd[j >> 1] = pack2<BF16>(p0 * (__uint_as_float(dp[j]) - dl[j]), p1 * (__uint_as_float(dp[j + 1]) - dl[j + 1]));
            }
          }
        } else if (!use_w) {
          if (kSplit) { tc_fence_before(); mbar_arrive((bars + 8u * (BB_CONS))); }
#pragma unroll
          for (int j = 0; j < 16; ++j) { pp[j] = 0u; pd[j] = 0u; }
        } else {
          // bias index: dr = qr' - (dR*W + kr)  ->  base + qr'*TW + qc'
          const float* tb = tab_h + ((2 * W - 1 - dR * W - kr) * TW + (2 * W - 1 - dC * W - kc));
          const bool ht = a.has_tab != 0;
          const uint32_t cb = kSplit ? (bars + 8u * (BB_CONS)) : 0u;
          if (half == 0) {
            dkv_quarter<W, 0, BF16>(pp, pd, saddr, paddr, a.scale_log2, ht, tb, use, ls, dl, 0u);
            dkv_quarter<W, 16, BF16>(pp + 8, pd + 8, saddr, paddr, a.scale_log2, ht, tb, use, ls, dl, cb);
          } else {
            dkv_quarter<W, 32, BF16>(pp, pd, saddr, paddr, a.scale_log2, ht, tb, use, ls, dl, 0u);
            dkv_quarter<W, 48, BF16>(pp + 8, pd + 8, saddr, paddr, a.scale_log2, ht, tb, use, ls, dl, cb);
          }
        }
        VIL_TR(3);
        if (kSplit) {
          if (G > 0) { mbar_wait((bars + 8u * (BB_PDONE)), (G - 1) & 1); tc_fence_after(); }   // previous dV / dK MMAs have read P^T / dS^T
          VIL_TR(8);
        } else {
          asm volatile("bar.sync 1, 256;" ::: "memory");     // all S / dP reads done before the in-place bf16 stores
        }
        tmem_st_x16(TM_P + lane_base + half * 16, pp);
        tmem_st_x16(TM_DS + lane_base + half * 16, pd);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive((bars + 8u * (BB_DSFULL + (G & 1))));
        VIL_TR(4);
        ++G;
        if (++stage == NS) { stage = 0; yphase ^= 1; }
      }
      VIL_TR(5);
      mbar_wait((bars + 8u * (BB_ACCDONE)), uc & 1);
      VIL_TR(6);
      tc_fence_after();
      const long long tok = geo.g + (long long)r * geo.ny + c;
      const uint32_t acc = (half == 0 ? TM_DK : TM_DV) + lane_base;
      const T4& out = half == 0 ? a.out0 : a.out1;
      const float f = half == 0 ? a.scale : 1.f;
#pragma unroll
      for (int q4 = 0; q4 < DP / 32; ++q4) {
        uint32_t ov[32];
        tmem_ld_x32(acc + q4 * 32, ov);
        tmem_ld_wait();
        if (q4 == DP / 32 - 1) { tc_fence_before(); mbar_arrive((bars + 8u * (BB_ACCFREE))); }
        if (row_ok) store_cols<32, BF16>(out, b, h, tok, geo.D, q4 * 32, ov, f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

}  // namespace tc
}  // namespace vil
