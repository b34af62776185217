// tcgen05 / TMA backward kernels of the Vision-Longformer attention (sm_100a), chunk size w <= 8, no rpe table
// gradient (configurations with a bias table use the SIMT family for the backward).
//
// Two deterministic passes (no atomics), both tiled like the forward (128-row tile = 2 chunk slots):
//   pass 1  vil_tc_bwd_dq  : query-stationary.  Per key block:  S = Q K^T, dP = dO V^T (SS MMAs into TMEM) ->
//                            threads: P = exp2(S c + bias - lse2), dS = P (dP - delta) -> bf16 dS in TMEM ->
//                            dQ += dS K (TS MMA, K tile MN-major).
//   pass 2  vil_tc_bwd_dkv : key-stationary (rows = keys).  Per query block: S^T = K Q^T, dP^T = V dO^T ->
//                            threads: P^T, dS^T (bf16, TMEM) -> dV += P^T dO, dK += dS^T Q (TS MMAs).
// S is recomputed in both passes (SlidingChunk2D.backward does the same work as slidingchunk_qk + _av + _agrad,
// slidingchunk_2d.py:234-246, on materialised score tensors).  lse2 = lse*log2(e) and delta are read from a
// chunk-ordered, 64-padded copy prepared by vil_tc_bwd_prep (invalid rows: lse2 = +inf -> P = 0).
#pragma once
#include "vil_tc_fwd.cuh"

namespace vil {
namespace tc {

constexpr int kBwdStages = 3;

struct BwdArgs {
  Geo geo;
  T4 out0, out1;                  // pass 1: dq (out0);  pass 2: dk (out0), dv (out1)
  const float* table;             // only to build the exact-window mask table (no bias-table gradient here)
  const float* g2l;
  const float* lse2c;             // (B,H,mx,my,64) log2-domain LSE, +inf on invalid rows
  const float* deltac;            // (B,H,mx,my,64) delta, 0 on invalid rows
  int cpairs, num_units, has_tab;
  float scale_log2, scale;
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

template <int DP>
struct BwdSmem {
  static constexpr int ROWB = DP * 2;
  static constexpr int X_BYTES = 128 * ROWB;            // one stationary tile
  static constexpr int Y_BYTES = 64 * ROWB;             // one streamed tile
  static constexpr int STAGE_BYTES = 2 * Y_BYTES + 512; // two tiles + lse2/delta (2 x 64 floats)
  static constexpr int OFF_X = 0;
  static constexpr int OFF_Y = 2 * X_BYTES;
  static constexpr int OFF_TAB = OFF_Y + kBwdStages * ((STAGE_BYTES + 1023) / 1024 * 1024);
  static constexpr int STAGE_STRIDE = (STAGE_BYTES + 1023) / 1024 * 1024;
  static __host__ __device__ int total(int tab_floats) { return OFF_TAB + tab_floats * 4 + 512 + 1024; }
};

enum { BB_XFULL = 0, BB_XEMPTY = 1, BB_YFULL = 2, BB_YEMPTY = 2 + kBwdStages, BB_SFULL = 2 + 2 * kBwdStages,
       BB_DSFULL = BB_SFULL + 1, BB_ACCDONE = BB_DSFULL + 1, BB_ACCFREE = BB_ACCDONE + 1, BB_COUNT = BB_ACCFREE + 1 };

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Walk of the QUERY chunks that visit the two key slots of a pass-2 unit (mirror image of BlockWalk).
struct QueryWalk {
  int R, C0, qr, qc, qr1, qc0, qc1;
  bool hasB;
  __device__ __forceinline__ void init(const Geo& g, int R_, int Cp) {
    R = R_; C0 = 2 * Cp;
    hasB = C0 + 1 < g.my;
    qr = max(R - 1, 0); qr1 = min(R + 1, g.mx - 1);
    qc0 = max(C0 - 1, 0); qc1 = min(C0 + 2, g.my - 1);
    qc = qc0;
  }
  __device__ __forceinline__ bool next(const Geo& g, int& QR, int& QC) {
    while (qr <= qr1) {
      const int r = qr, c = qc;
      if (++qc > qc1) { qc = qc0; ++qr; }
      const bool useA = offset_used(g, R - r, C0 - c);
      const bool useB = hasB && offset_used(g, R - r, C0 + 1 - c);
      if (useA || useB) { QR = r; QC = c; return true; }
    }
    return false;
  }
};

template <int DP, bool BF16>
__device__ __forceinline__ void store_row_scaled(const T4& t, int b, int h, long long tok, int D, const uint32_t (*ov)[32],
                                                 float f) {
  constexpr int OC = DP / 32;
  char* base = t.p + ((long long)b * t.sb + (long long)h * t.sh + tok * t.st) * 2;
#pragma unroll
  for (int q4 = 0; q4 < OC; ++q4)
#pragma unroll
    for (int v8 = 0; v8 < 4; ++v8) {
      if (q4 * 32 + v8 * 8 < D) {
        uint4 pkt;
        pkt.x = pack2<BF16>(__uint_as_float(ov[q4][v8 * 8 + 0]) * f, __uint_as_float(ov[q4][v8 * 8 + 1]) * f);
        pkt.y = pack2<BF16>(__uint_as_float(ov[q4][v8 * 8 + 2]) * f, __uint_as_float(ov[q4][v8 * 8 + 3]) * f);
        pkt.z = pack2<BF16>(__uint_as_float(ov[q4][v8 * 8 + 4]) * f, __uint_as_float(ov[q4][v8 * 8 + 5]) * f);
        pkt.w = pack2<BF16>(__uint_as_float(ov[q4][v8 * 8 + 6]) * f, __uint_as_float(ov[q4][v8 * 8 + 7]) * f);
        *reinterpret_cast<uint4*>(base + (q4 * 32 + v8 * 8) * 2) = pkt;
      }
    }
}

// ======================================================================================================== pass 1
template <int DP, int W, bool BF16>
__global__ void __launch_bounds__(kThreads, 2)
vil_tc_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmKg, const __grid_constant__ CUtensorMap tmVg, const BwdArgs a) {
  using SM = BwdSmem<DP>;
  constexpr int ROWB = SM::ROWB;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int W2 = W * W, TW = 4 * W - 1;
  const Geo& geo = a.geo;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem + SM::OFF_X;
  unsigned char* sDO = sQ + SM::X_BYTES;
  unsigned char* sY = smem + SM::OFF_Y;
  float* tab = reinterpret_cast<float*>(smem + SM::OFF_TAB);
  const int tabn = a.has_tab ? TW * TW : 0;
  float* g2l_s = tab + geo.H * tabn;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(g2l_s + geo.H * 16) + 15) & ~uintptr_t(15));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BB_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < SM::OFF_TAB / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < geo.H * tabn; i += kThreads) {
    const int h = i / tabn, idx = i % tabn;
    const int dr = idx / TW - (2 * W - 1), dc = idx % TW - (2 * W - 1);
    float v = (a.table != nullptr) ? a.table[(long long)idx * geo.H + h] * 1.4426950408889634f : 0.f;
    if (geo.exact == 1 && (abs(dr) > W || abs(dc) > W)) v = -INFINITY;
    tab[i] = v;
  }
  for (int i = tid; i < geo.H * 16; i += kThreads) {
    const int h = i / 16, t = i % 16;
    g2l_s[i] = (a.g2l != nullptr && t < geo.g) ? a.g2l[((long long)geo.H + h) * geo.g + t] * 1.4426950408889634f : 0.f;
  }
  if (tid == 0) {
    mbar_init(&bars[BB_XFULL], 1); mbar_init(&bars[BB_XEMPTY], 1);
    for (int i = 0; i < kBwdStages; ++i) { mbar_init(&bars[BB_YFULL + i], 1); mbar_init(&bars[BB_YEMPTY + i], 1); }
    mbar_init(&bars[BB_SFULL], 1); mbar_init(&bars[BB_DSFULL], 128);
    mbar_init(&bars[BB_ACCDONE], 1); mbar_init(&bars[BB_ACCFREE], 128);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t TM_S = tmem, TM_DP = tmem + 64, TM_ACC = tmem + 128;      // dS overwrites S; dQ accumulator
  const int units_per_bh = geo.mx * a.cpairs;

  if (warp == 4) {
    if (elect_one()) {
      uint32_t stage = 0, yphase = 0, uc = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh = unit / units_per_bh, rem = unit % units_per_bh;
        const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
        if (uc >= 1) mbar_wait(&bars[BB_XEMPTY], (uc - 1) & 1);
        const bool hasB = 2 * Cp + 1 < geo.my;
        mbar_arrive_expect_tx(&bars[BB_XFULL], (hasB ? 4 : 2) * W2 * ROWB);
        tma_load_5d(sQ, &tmQ, &bars[BB_XFULL], 0, (2 * Cp) * W, R * W, h, b);
        tma_load_5d(sDO, &tmDO, &bars[BB_XFULL], 0, (2 * Cp) * W, R * W, h, b);
        if (hasB) {
          tma_load_5d(sQ + 64 * ROWB, &tmQ, &bars[BB_XFULL], 0, (2 * Cp + 1) * W, R * W, h, b);
          tma_load_5d(sDO + 64 * ROWB, &tmDO, &bars[BB_XFULL], 0, (2 * Cp + 1) * W, R * W, h, b);
        }
        BlockWalk wk; wk.init(geo, R, Cp);
        int type, KR, KC;
        while (wk.next(geo, type, KR, KC)) {
          mbar_wait(&bars[BB_YEMPTY + stage], yphase ^ 1);
          unsigned char* dK = sY + stage * SM::STAGE_STRIDE;
          unsigned char* dV = dK + SM::Y_BYTES;
          if (type == 1) {
            mbar_arrive_expect_tx(&bars[BB_YFULL + stage], 2 * 16 * ROWB);
            tma_load_4d(dK, &tmKg, &bars[BB_YFULL + stage], 0, 0, h, b);
            tma_load_4d(dV, &tmVg, &bars[BB_YFULL + stage], 0, 0, h, b);
          } else {
            mbar_arrive_expect_tx(&bars[BB_YFULL + stage], 2 * W2 * ROWB);
            tma_load_5d(dK, &tmK, &bars[BB_YFULL + stage], 0, KC * W, KR * W, h, b);
            tma_load_5d(dV, &tmV, &bars[BB_YFULL + stage], 0, KC * W, KR * W, h, b);
          }
          if (++stage == kBwdStages) { stage = 0; yphase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t IDESC_S = make_idesc(128, 64, BF16, false, false);
      constexpr uint32_t IDESC_SG = make_idesc(128, 16, BF16, false, false);
      constexpr uint32_t IDESC_ACC = make_idesc(128, DP, BF16, false, true);
      uint32_t stage = 0, yphase = 0, uc = 0, G = 0;
      const uint32_t qaddr = smem_u32(sQ), doaddr = smem_u32(sDO);
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int rem = unit % units_per_bh;
        const int R = rem / a.cpairs, Cp = rem % a.cpairs;
        mbar_wait(&bars[BB_XFULL], uc & 1);
        auto issue_SdP = [&](uint32_t st, int type) {
          const uint32_t kaddr = smem_u32(sY + st * SM::STAGE_STRIDE), vaddr = kaddr + SM::Y_BYTES;
          const uint32_t idesc = type == 1 ? IDESC_SG : IDESC_S;
#pragma unroll
          for (int k = 0; k < DP / 16; ++k)
            mma_ss(TM_S, make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT), make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT), idesc, k > 0);
#pragma unroll
          for (int k = 0; k < DP / 16; ++k)
            mma_ss(TM_DP, make_smem_desc(doaddr + k * 32, 16, SBO, LAYOUT), make_smem_desc(vaddr + k * 32, 16, SBO, LAYOUT), idesc, k > 0);
          mma_commit(&bars[BB_SFULL]);
        };
        BlockWalk wk; wk.init(geo, R, Cp);
        int type, KR, KC;
        bool have = wk.next(geo, type, KR, KC);
        mbar_wait(&bars[BB_YFULL + stage], yphase);
        tc_fence_after();
        issue_SdP(stage, type);
        bool first = true;
        while (have) {
          const uint32_t cur_stage = stage;
          const int cur_type = type;
          if (++stage == kBwdStages) { stage = 0; yphase ^= 1; }
          have = wk.next(geo, type, KR, KC);
          mbar_wait(&bars[BB_DSFULL], G & 1);
          if (first && uc > 0) mbar_wait(&bars[BB_ACCFREE], (uc - 1) & 1);
          tc_fence_after();
          const uint32_t kaddr = smem_u32(sY + cur_stage * SM::STAGE_STRIDE);
          const int ksteps = cur_type == 1 ? 1 : 4;
          for (int k = 0; k < ksteps; ++k)
            mma_ts(TM_ACC, TM_S + k * 8, make_smem_desc(kaddr + k * 16 * ROWB, 16, SBO, LAYOUT), IDESC_ACC, (!first) || k > 0);
          mma_commit(&bars[BB_YEMPTY + cur_stage]);
          first = false;
          ++G;
          if (have) {
            mbar_wait(&bars[BB_YFULL + stage], yphase);
            tc_fence_after();
            issue_SdP(stage, type);
          } else {
            mma_commit(&bars[BB_ACCDONE]);
            mma_commit(&bars[BB_XEMPTY]);
          }
        }
      }
    }
  } else {
    const int row = tid, slot = row >> 6, l = row & 63;
    const int qr = l / W, qc = l % W;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t uc = 0, G = 0;
    for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
      const int bh = unit / units_per_bh, rem = unit % units_per_bh;
      const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
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
        mbar_wait(&bars[BB_SFULL], G & 1);
        tc_fence_after();
        const uint32_t saddr = TM_S + lane_base, paddr = TM_DP + lane_base;
        if (type == 1) {
          uint32_t s[16], dp[16], pk[8];
          tmem_ld_x16(saddr, s);
          tmem_ld_x16(paddr, dp);
          tmem_ld_wait();
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
          tmem_st_x8(saddr, pk);
        } else {
          const int dR = KR - R, dC = KC - C;
          const bool use = slot_ok && offset_used(geo, dR, dC);
          uint32_t pk[32];
          if (!use) {
#pragma unroll
            for (int j = 0; j < 32; ++j) pk[j] = 0u;
          } else {
            const int krows = min(W, geo.nx - KR * W), kcols = min(W, geo.ny - KC * W);
            const bool masked = (krows < W) || (kcols < W);
            const float* tb = tab_h + ((qr - dR * W + 2 * W - 1) * TW + (qc - dC * W + 2 * W - 1));
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t s[32], dp[32];
              tmem_ld_x32(saddr + hf * 32, s);
              tmem_ld_x32(paddr + hf * 32, dp);
              tmem_ld_wait();
#pragma unroll
              for (int jj = 0; jj < 32; jj += 2) {
                float dsv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int j = hf * 32 + jj + e;
                  float v = 0.f;
                  if (j < W2) {
                    float x = __uint_as_float(s[jj + e]) * a.scale_log2;
                    if (a.has_tab) x += tb[-((j / W) * TW + (j % W))];
                    const bool ok = !masked || ((j / W) < krows && (j % W) < kcols);
                    const float p = ok ? fast_exp2(x - lse2) : 0.f;
                    v = p * (__uint_as_float(dp[jj + e]) - del);
                  }
                  dsv[e] = v;
                }
                pk[(hf * 32 + jj) >> 1] = pack2<BF16>(dsv[0], dsv[1]);
              }
            }
          }
          tmem_st_x32(saddr, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[BB_DSFULL]);
        ++G;
      }
      mbar_wait(&bars[BB_ACCDONE], uc & 1);
      tc_fence_after();
      constexpr int OC = DP / 32;
      uint32_t ov[OC][32];
#pragma unroll
      for (int q4 = 0; q4 < OC; ++q4) tmem_ld_x32(TM_ACC + lane_base + q4 * 32, ov[q4]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bars[BB_ACCFREE]);
      if (row_ok) store_row_scaled<DP, BF16>(a.out0, b, h, (long long)r * geo.ny + c, geo.D, ov, a.scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// ======================================================================================================== pass 2
template <int DP, int W, bool BF16>
__global__ void __launch_bounds__(kThreads, 2)
vil_tc_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const BwdArgs a) {
  using SM = BwdSmem<DP>;
  constexpr int ROWB = SM::ROWB;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int W2 = W * W, TW = 4 * W - 1;
  const Geo& geo = a.geo;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sK = smem + SM::OFF_X;
  unsigned char* sV = sK + SM::X_BYTES;
  unsigned char* sY = smem + SM::OFF_Y;
  float* tab = reinterpret_cast<float*>(smem + SM::OFF_TAB);
  const int tabn = a.has_tab ? TW * TW : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(tab + geo.H * tabn) + 15) & ~uintptr_t(15));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BB_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < SM::OFF_TAB / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < geo.H * tabn; i += kThreads) {
    const int h = i / tabn, idx = i % tabn;
    const int dr = idx / TW - (2 * W - 1), dc = idx % TW - (2 * W - 1);
    float v = (a.table != nullptr) ? a.table[(long long)idx * geo.H + h] * 1.4426950408889634f : 0.f;
    if (geo.exact == 1 && (abs(dr) > W || abs(dc) > W)) v = -INFINITY;
    tab[i] = v;
  }
  if (tid == 0) {
    mbar_init(&bars[BB_XFULL], 1); mbar_init(&bars[BB_XEMPTY], 1);
    for (int i = 0; i < kBwdStages; ++i) { mbar_init(&bars[BB_YFULL + i], 1); mbar_init(&bars[BB_YEMPTY + i], 1); }
    mbar_init(&bars[BB_SFULL], 1); mbar_init(&bars[BB_DSFULL], 128);
    mbar_init(&bars[BB_ACCDONE], 1); mbar_init(&bars[BB_ACCFREE], 128);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t TM_S = tmem, TM_DP = tmem + 64, TM_DK = tmem + 128, TM_DV = tmem + 192;
  const int units_per_bh = geo.mx * a.cpairs;

  if (warp == 4) {
    if (elect_one()) {
      uint32_t stage = 0, yphase = 0, uc = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh = unit / units_per_bh, rem = unit % units_per_bh;
        const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
        if (uc >= 1) mbar_wait(&bars[BB_XEMPTY], (uc - 1) & 1);
        const bool hasB = 2 * Cp + 1 < geo.my;
        mbar_arrive_expect_tx(&bars[BB_XFULL], (hasB ? 4 : 2) * W2 * ROWB);
        tma_load_5d(sK, &tmK, &bars[BB_XFULL], 0, (2 * Cp) * W, R * W, h, b);
        tma_load_5d(sV, &tmV, &bars[BB_XFULL], 0, (2 * Cp) * W, R * W, h, b);
        if (hasB) {
          tma_load_5d(sK + 64 * ROWB, &tmK, &bars[BB_XFULL], 0, (2 * Cp + 1) * W, R * W, h, b);
          tma_load_5d(sV + 64 * ROWB, &tmV, &bars[BB_XFULL], 0, (2 * Cp + 1) * W, R * W, h, b);
        }
        QueryWalk wk; wk.init(geo, R, Cp);
        int QR, QC;
        while (wk.next(geo, QR, QC)) {
          mbar_wait(&bars[BB_YEMPTY + stage], yphase ^ 1);
          unsigned char* dQ = sY + stage * SM::STAGE_STRIDE;
          unsigned char* dG = dQ + SM::Y_BYTES;
          unsigned char* dL = dG + SM::Y_BYTES;
          mbar_arrive_expect_tx(&bars[BB_YFULL + stage], 2 * W2 * ROWB + 512);
          tma_load_5d(dQ, &tmQ, &bars[BB_YFULL + stage], 0, QC * W, QR * W, h, b);
          tma_load_5d(dG, &tmDO, &bars[BB_YFULL + stage], 0, QC * W, QR * W, h, b);
          const long long ci = (((long long)bh * geo.mx + QR) * geo.my + QC) * 64;
          bulk_load_1d(dL, a.lse2c + ci, 256, &bars[BB_YFULL + stage]);
          bulk_load_1d(dL + 256, a.deltac + ci, 256, &bars[BB_YFULL + stage]);
          if (++stage == kBwdStages) { stage = 0; yphase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t IDESC_S = make_idesc(128, 64, BF16, false, false);
      constexpr uint32_t IDESC_ACC = make_idesc(128, DP, BF16, false, true);
      uint32_t stage = 0, yphase = 0, uc = 0, G = 0;
      const uint32_t kaddr = smem_u32(sK), vaddr = smem_u32(sV);
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int rem = unit % units_per_bh;
        const int R = rem / a.cpairs, Cp = rem % a.cpairs;
        mbar_wait(&bars[BB_XFULL], uc & 1);
        auto issue_SdP = [&](uint32_t st) {
          const uint32_t qaddr = smem_u32(sY + st * SM::STAGE_STRIDE), gaddr = qaddr + SM::Y_BYTES;
#pragma unroll
          for (int k = 0; k < DP / 16; ++k)
            mma_ss(TM_S, make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT), make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT), IDESC_S, k > 0);
#pragma unroll
          for (int k = 0; k < DP / 16; ++k)
            mma_ss(TM_DP, make_smem_desc(vaddr + k * 32, 16, SBO, LAYOUT), make_smem_desc(gaddr + k * 32, 16, SBO, LAYOUT), IDESC_S, k > 0);
          mma_commit(&bars[BB_SFULL]);
        };
        QueryWalk wk; wk.init(geo, R, Cp);
        int QR, QC;
        bool have = wk.next(geo, QR, QC);
        mbar_wait(&bars[BB_YFULL + stage], yphase);
        tc_fence_after();
        issue_SdP(stage);
        bool first = true;
        while (have) {
          const uint32_t cur_stage = stage;
          if (++stage == kBwdStages) { stage = 0; yphase ^= 1; }
          have = wk.next(geo, QR, QC);
          mbar_wait(&bars[BB_DSFULL], G & 1);
          if (first && uc > 0) mbar_wait(&bars[BB_ACCFREE], (uc - 1) & 1);
          tc_fence_after();
          const uint32_t qaddr = smem_u32(sY + cur_stage * SM::STAGE_STRIDE), gaddr = qaddr + SM::Y_BYTES;
          for (int k = 0; k < 4; ++k)       // dV += P^T dO
            mma_ts(TM_DV, TM_S + k * 8, make_smem_desc(gaddr + k * 16 * ROWB, 16, SBO, LAYOUT), IDESC_ACC, (!first) || k > 0);
          for (int k = 0; k < 4; ++k)       // dK += dS^T Q
            mma_ts(TM_DK, TM_DP + k * 8, make_smem_desc(qaddr + k * 16 * ROWB, 16, SBO, LAYOUT), IDESC_ACC, (!first) || k > 0);
          mma_commit(&bars[BB_YEMPTY + cur_stage]);
          first = false;
          ++G;
          if (have) {
            mbar_wait(&bars[BB_YFULL + stage], yphase);
            tc_fence_after();
            issue_SdP(stage);
          } else {
            mma_commit(&bars[BB_ACCDONE]);
            mma_commit(&bars[BB_XEMPTY]);
          }
        }
      }
    }
  } else {
    const int row = tid, slot = row >> 6, l = row & 63;
    const int kr = l / W, kc = l % W;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t uc = 0, G = 0, stage = 0, yphase = 0;
    for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
      const int bh = unit / units_per_bh, rem = unit % units_per_bh;
      const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
      const int C = 2 * Cp + slot;
      const int r = R * W + kr, c = C * W + kc;
      const bool slot_ok = C < geo.my;
      const bool row_ok = slot_ok && l < W2 && r < geo.nx && c < geo.ny;
      const float* tab_h = tab + h * tabn;
      QueryWalk wk; wk.init(geo, R, Cp);
      int QR, QC;
      while (wk.next(geo, QR, QC)) {
        mbar_wait(&bars[BB_YFULL + stage], yphase);     // lse2 / delta of this query block have landed
        mbar_wait(&bars[BB_SFULL], G & 1);
        tc_fence_after();
        const float* ls = reinterpret_cast<const float*>(sY + stage * SM::STAGE_STRIDE + 2 * SM::Y_BYTES);
        const float* dl = ls + 64;
        const uint32_t saddr = TM_S + lane_base, paddr = TM_DP + lane_base;
        const int dR = R - QR, dC = C - QC;       // offset = key chunk - query chunk
        const bool use = row_ok && offset_used(geo, dR, dC);
        const bool use_w = slot_ok && offset_used(geo, dR, dC);     // warp-uniform part
        uint32_t pp[32], pd[32];
        if (!use_w) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { pp[j] = 0u; pd[j] = 0u; }
        } else {
          // bias index: dr = qr' - (dR*W + kr)  ->  base + qr'*TW + qc'
          const float* tb = tab_h + ((2 * W - 1 - dR * W - kr) * TW + (2 * W - 1 - dC * W - kc));
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t s[32], dp[32];
            tmem_ld_x32(saddr + hf * 32, s);
            tmem_ld_x32(paddr + hf * 32, dp);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; jj += 2) {
              float pv[2], dv[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int j = hf * 32 + jj + e;
                float p = 0.f, d = 0.f;
                if (j < W2) {
                  float x = __uint_as_float(s[jj + e]) * a.scale_log2;
                  if (a.has_tab) x += tb[(j / W) * TW + (j % W)];
                  p = use ? fast_exp2(x - ls[j]) : 0.f;              // ls = +inf for invalid queries
                  d = p * (__uint_as_float(dp[jj + e]) - dl[j]);
                }
                pv[e] = p; dv[e] = d;
              }
              pp[(hf * 32 + jj) >> 1] = pack2<BF16>(pv[0], pv[1]);
              pd[(hf * 32 + jj) >> 1] = pack2<BF16>(dv[0], dv[1]);
            }
          }
        }
        tmem_st_x32(saddr, pp);
        tmem_st_x32(paddr, pd);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[BB_DSFULL]);
        ++G;
        if (++stage == kBwdStages) { stage = 0; yphase ^= 1; }
      }
      mbar_wait(&bars[BB_ACCDONE], uc & 1);
      tc_fence_after();
      constexpr int OC = DP / 32;
      const long long tok = geo.g + (long long)r * geo.ny + c;
      {
        uint32_t ov[OC][32];
#pragma unroll
        for (int q4 = 0; q4 < OC; ++q4) tmem_ld_x32(TM_DK + lane_base + q4 * 32, ov[q4]);
        tmem_ld_wait();
        if (row_ok) store_row_scaled<DP, BF16>(a.out0, b, h, tok, geo.D, ov, a.scale);
      }
      {
        uint32_t ov[OC][32];
#pragma unroll
        for (int q4 = 0; q4 < OC; ++q4) tmem_ld_x32(TM_DV + lane_base + q4 * 32, ov[q4]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bars[BB_ACCFREE]);
        if (row_ok) store_row_scaled<DP, BF16>(a.out1, b, h, tok, geo.D, ov, 1.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

}  // namespace tc
}  // namespace vil
