// Fused tcgen05 / TMA backward of the Vision-Longformer attention (sm_100a), chunk size w <= 8, no bias-table gradient
// (rpe off; the rpe-on backward keeps the round-1 pipeline of vil_tc_bwd.cuh).  Round-2 successor of the 7-launch
// pipeline  simt_bwd_delta + vil_tc_bwd_prep(_g) + dq + dkv + simt_bwd_gcol + simt_bwd_grow : here it is
//
//   pass 1  vil_tc_bwd2_dq   query-stationary, dQ.   ALSO: delta_i = dO_i . O_i computed by the compute threads at the
//                            start of each unit (O is read exactly once, here), lse2 / delta emitted chunk-ordered for
//                            pass 2 (no prep kernels), and - when the global query rows fit the spare lanes 56..63 of
//                            slot A (w <= 7, nglo <= 8, mode 0, shared k/v) - dq of the GLOBAL query rows: they ride in
//                            the Q / dO tiles, see only the chunks the unit owns (addend -inf elsewhere), and leave a
//                            per-unit partial dq_g in the accumulator rows 56..;
//   pass 2  vil_tc_bwd2_dkv  key-stationary, dK / dV.  The global KEY rows ride in lanes 56..63 of the K / V tiles the
//                            same way (partial dk_g / dv_g per unit); the global QUERY rows stay the 16-column first
//                            block of every unit (round 1), whose g x g corner is counted by unit (0,0) only;
//   merge   vil_tc_bwd2_merge  sums the per-unit partials into dq_g, dk[:, :, :g], dv[:, :, :g]  (tiny).
//
// Math as in vil_tc_bwd.cuh / SlidingChunk2D.backward (slidingchunk_2d.py:234-246): P = exp2(S c + bias - lse2),
// dS = P (dP - delta), dQ = dS K, dK = dS^T Q, dV = P^T dO; S is recomputed in both passes.
#pragma once
#include "vil_tc_bwd.cuh"

namespace vil {
namespace tc {
namespace b2 {

using namespace sm100;

constexpr int kGRow0 = 56, kGMax = 8;

struct Args {
  Geo geo;
  T4 out0, out1;                  // pass 1: dq;  pass 2: dk, dv
  T4 o, d_o, og, d_og;            // pass 1: rows of O / dO (delta), element type T (fp32 O in the parity build)
  const float* lse;               // (B,H,Nloc) natural log
  const float* lse_g;             // (B,H,g)
  const float* table;             // exact-window mask table source (no bias here) or null
  float* lse2c;                   // (B,H,mx,my,64) written by pass 1, read by pass 2
  float* deltac;
  float* lse2g;                   // (B*H,16) written by pass 1 (unit (0,0)), read by pass 2
  float* deltag;
  float* part;                    // per-unit partials of the global rows: pass 1 [bh][unit][8][DP]; pass 2 [bh][unit][8][2][DP]
  int cpairs, num_units, has_tab;
  int fuse_q;                     // pass 1: global query rows in the spare lanes;  pass 2: global key rows likewise
  int fuse_g;                     // pass 2: global query rows as the first 16-column block (round 1)
  int out_f32;
  float scale_log2, scale;
};

// Shared-memory plan of the fused passes: like round 1 (BwdSmem) but with a deeper streamed-tile ring where shared memory
// allows it (D <= 32: 6 stages instead of 3 - a stage is only 9 KB there and the TMA latency of ~1.5 us otherwise starves the
// MMA issuer: a block is consumed every ~0.6 us).
template <int DP>
struct Smem2 {
  static constexpr int ROWB = DP * 2;
  static constexpr int NS = DP == 64 ? 2 : 6;
  static constexpr int X_BYTES = 128 * ROWB;
  static constexpr int Y_BYTES = 64 * ROWB;
  static constexpr int STAGE_STRIDE = (2 * Y_BYTES + 512 + 1023) / 1024 * 1024;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_Y = 4 * X_BYTES;
  static constexpr int OFF_TAB = OFF_Y + NS * STAGE_STRIDE;
  static __host__ __device__ int total(int tab_floats) { return OFF_TAB + tab_floats * 4 + 512 + 1024; }
};

// barriers (own numbering: up to 6 ring stages) + one named barrier (id 2) for the delta exchange of pass 1
enum { BB_XFULL = 0, BB_XEMPTY = 2, BB_YFULL = 4, BB_YEMPTY = 10, BB_SFULL = 16, BB_DSFULL = 17, BB_ACCDONE = 19,
       BB_ACCFREE = 20, BB_CONS = 21, BB_PDONE = 22, BB_COUNT = 23 };

__device__ __forceinline__ void init_bwd_barriers(uint32_t bars, int ns) {
  for (int i = 0; i < 2; ++i) { mbar_init((bars + 8u * (BB_XFULL + i)), 1); mbar_init((bars + 8u * (BB_XEMPTY + i)), 1); }
  for (int i = 0; i < ns; ++i) { mbar_init((bars + 8u * (BB_YFULL + i)), 1); mbar_init((bars + 8u * (BB_YEMPTY + i)), 1); }
  mbar_init((bars + 8u * (BB_SFULL)), 1); mbar_init((bars + 8u * (BB_DSFULL)), 256); mbar_init((bars + 8u * (BB_DSFULL + 1)), 256);
  mbar_init((bars + 8u * (BB_CONS)), 256); mbar_init((bars + 8u * (BB_PDONE)), 1);
  mbar_init((bars + 8u * (BB_ACCDONE)), 1); mbar_init((bars + 8u * (BB_ACCFREE)), 256);
  fence_barrier_init();
}

// half a row (NC = DP/2 channels starting at c0) of a (B,H,T,D) view -> fp32 registers; channels >= D read as 0
template <int NC, typename TE>
__device__ __forceinline__ void load_half_row(const T4& t, int b, int h, long long tok, int D, int c0, float (&r)[NC]) {
  const TE* p = reinterpret_cast<const TE*>(t.p) + (long long)b * t.sb + (long long)h * t.sh + tok * t.st + c0;
  constexpr int PER16 = 16 / (int)sizeof(TE);
#pragma unroll
  for (int v = 0; v < NC / PER16; ++v) {
    if (c0 + v * PER16 < D) {
      const int4 raw = __ldg(reinterpret_cast<const int4*>(p) + v);
      const TE* e = reinterpret_cast<const TE*>(&raw);
#pragma unroll
      for (int u = 0; u < PER16; ++u) r[v * PER16 + u] = ElemTraits<TE>::to_f(e[u]);
    } else {
#pragma unroll
      for (int u = 0; u < PER16; ++u) r[v * PER16 + u] = 0.f;
    }
  }
}

// pass-1 element work for 16 columns with a per-thread addend folded into lse2 (radd = -inf switches the row off)
template <int W, int COL0, bool BF16, bool HAS_TAB, bool MASKED, int NV = W * W>
__device__ __forceinline__ void dq_cols16(uint32_t* __restrict__ pk, const uint32_t (&s)[16], const uint32_t (&dp)[16], float c,
                                          const float* __restrict__ tb, int krows, int kcols, float nlse2, float del) {
  constexpr int TW = 4 * W - 1;
#pragma unroll
  for (int jj = 0; jj < 16; jj += 2) {
    const int j = COL0 + jj;
    float dsv[2] = {0.f, 0.f};
    if (j < NV) {
      float x[2], t[2], p[2];
      if constexpr (HAS_TAB) {
        const float b0 = tb[-((j / W) * TW + (j % W))];
        const float b1 = (j + 1 < NV) ? tb[-(((j + 1) / W) * TW + ((j + 1) % W))] : 0.f;
        ffma2(x[0], x[1], __uint_as_float(s[jj]), __uint_as_float(s[jj + 1]), c, c, b0 + nlse2, b1 + nlse2);
      } else {
        ffma2(x[0], x[1], __uint_as_float(s[jj]), __uint_as_float(s[jj + 1]), c, c, nlse2, nlse2);
      }
      p[0] = fast_exp2(x[0]);
      p[1] = (j + 1 < NV) ? fast_exp2(x[1]) : 0.f;
      if constexpr (MASKED) {
        p[0] = ((j / W) < krows && (j % W) < kcols) ? p[0] : 0.f;
        p[1] = (((j + 1) / W) < krows && ((j + 1) % W) < kcols) ? p[1] : 0.f;
      }
      fadd2(t[0], t[1], __uint_as_float(dp[jj]), __uint_as_float(dp[jj + 1]), -del, -del);
      fmul2(dsv[0], dsv[1], p[0], p[1], t[0], t[1]);
    }
    pk[jj >> 1] = pack2<BF16>(dsv[0], dsv[1]);
  }
}
template <int W, int COL0, bool BF16, bool LEAN = false, int NV = W * W>
__device__ __forceinline__ void dq_quarter(uint32_t* __restrict__ pk, uint32_t saddr, uint32_t paddr, float c, bool has_tab,
                                           const float* __restrict__ tb, bool masked, int krows, int kcols, float nlse2,
                                           float del, uint32_t cons_bar) {
  uint32_t s[16], dp[16];
  tmem_ld_x16(saddr + COL0, s);
  tmem_ld_x16(paddr + COL0, dp);
  tmem_ld_wait();
  if (cons_bar != 0u) { tc_fence_before(); mbar_arrive(cons_bar); }     // last read of S / dP by this thread
  if constexpr (LEAN) {
    dq_cols16<W, COL0, BF16, false, false, NV>(pk, s, dp, c, tb, krows, kcols, nlse2, del);
    return;
  }
  if (has_tab) {
    if (masked) dq_cols16<W, COL0, BF16, true, true, NV>(pk, s, dp, c, tb, krows, kcols, nlse2, del);
    else        dq_cols16<W, COL0, BF16, true, false, NV>(pk, s, dp, c, tb, krows, kcols, nlse2, del);
  } else {
    if (masked) dq_cols16<W, COL0, BF16, false, true, NV>(pk, s, dp, c, tb, krows, kcols, nlse2, del);
    else        dq_cols16<W, COL0, BF16, false, false, NV>(pk, s, dp, c, tb, krows, kcols, nlse2, del);
  }
}

// ======================================================================================================== pass 1
// LEAN: no padded chunk in the geometry and no table -> the masked / table code paths and their per-block set-up are
//       compiled out (the per-block control code was more than half of the instructions of a warp-block).
template <int DP, int W, bool BF16, typename TO, bool LEAN>   // TO: element type of o / og (T, or float in the parity build)
__global__ void __launch_bounds__(kBwdThreads, 2)
vil_tc_bwd2_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const __grid_constant__ CUtensorMap tmKg, const __grid_constant__ CUtensorMap tmVg,
                      const __grid_constant__ CUtensorMap tmQg, const __grid_constant__ CUtensorMap tmDOg, const Args a) {
  using SM = Smem2<DP>;
  using TE = typename std::conditional<BF16, __nv_bfloat16, __half>::type;
  constexpr int ROWB = SM::ROWB, NS = SM::NS;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int W2 = W * W, TW = 4 * W - 1;
  constexpr int ZPAD = (W - 1) * TW + W;
  const Geo& geo = a.geo;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sX = smem + SM::OFF_X;                 // [buf][Q | dO]
  unsigned char* sY = smem + SM::OFF_Y;
  float* tab = reinterpret_cast<float*>(smem + SM::OFF_TAB);
  const int tabn = a.has_tab ? TW * TW : 0;
  float* zpad = tab + geo.H * tabn;                      // [ZPAD] zero "table" of the global rows (has_tab only)
  float* dx = zpad + (a.has_tab ? ZPAD : 0);             // [2][128] delta exchange between the two column halves
  const int bars_off = (SM::OFF_TAB + (geo.H * tabn + (a.has_tab ? ZPAD : 0) + 256) * 4 + 15) & ~15;
  uint64_t* bars_p = reinterpret_cast<uint64_t*>(smem + bars_off);
  const uint32_t bars = smem_u32(bars_p);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_p + BB_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < SM::OFF_TAB / 16; i += kBwdThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < geo.H * tabn; i += kBwdThreads) {
    const int idx = i % tabn;
    const int dr = idx / TW - (2 * W - 1), dc = idx % TW - (2 * W - 1);
    tab[i] = (geo.exact == 1 && (abs(dr) > W || abs(dc) > W)) ? -INFINITY : 0.f;       // window mask only: no bias on this path
  }
  if (a.has_tab) for (int i = tid; i < ZPAD; i += kBwdThreads) zpad[i] = 0.f;
  if (tid == 0) init_bwd_barriers(bars, NS);
  if (warp == 8) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t TM_S = tmem, TM_DP = tmem + 64, TM_DS = tmem + 128, TM_ACC = tmem + 192;
  const int units_per_bh = geo.mx * a.cpairs;

  if (warp == 8) {
    // ================================================================= TMA producer
    if (elect_one()) {
      uint32_t stage = 0, yphase = 0, uc = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh = unit / units_per_bh, rem = unit % units_per_bh;
        const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
        const uint32_t xb = uc & 1, xphase = (uc >> 1) & 1;
        if (uc >= 2) mbar_wait((bars + 8u * (BB_XEMPTY + xb)), xphase ^ 1);
        unsigned char* sQ = sX + xb * 2 * SM::X_BYTES;
        unsigned char* sDO = sQ + SM::X_BYTES;
        const bool hasB = 2 * Cp + 1 < geo.my;
        mbar_arrive_expect_tx((bars + 8u * (BB_XFULL + xb)), ((hasB ? 4 : 2) * W2 + (a.fuse_q ? 16 : 0)) * ROWB);
        tma_load_5d(sQ, &tmQ, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp) * W, R * W, h, b);
        tma_load_5d(sDO, &tmDO, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp) * W, R * W, h, b);
        if (hasB) {
          tma_load_5d(sQ + 64 * ROWB, &tmQ, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp + 1) * W, R * W, h, b);
          tma_load_5d(sDO + 64 * ROWB, &tmDO, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp + 1) * W, R * W, h, b);
        }
        if (a.fuse_q) {
          tma_load_4d(sQ + kGRow0 * ROWB, &tmQg, (bars + 8u * (BB_XFULL + xb)), 0, 0, h, b);
          tma_load_4d(sDO + kGRow0 * ROWB, &tmDOg, (bars + 8u * (BB_XFULL + xb)), 0, 0, h, b);
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
    // ================================================================= MMA issuer (unchanged protocol of round 1)
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
        const uint32_t qaddr = smem_u32(sX + xb * 2 * SM::X_BYTES), doaddr = qaddr + SM::X_BYTES;
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
          uint64_t kacc[4];
          {
            const uint32_t kaddr = smem_u32(sY + cur_stage * SM::STAGE_STRIDE);
#pragma unroll
            for (int k = 0; k < 4; ++k) kacc[k] = make_smem_desc(kaddr + k * 16 * ROWB, 16, SBO, LAYOUT);
          }
          if (++stage == NS) { stage = 0; yphase ^= 1; }
          have = wk.next(geo, type, KR, KC);
          if (have) {
            prep_SdP(stage);
            mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase);
            mbar_wait((bars + 8u * (BB_CONS)), G & 1);                // S_j / dP_j are in the threads' registers
            tc_fence_after();
            issue_SdP(type);
          }
          mbar_wait((bars + 8u * (BB_DSFULL + (G & 1))), (G >> 1) & 1);
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
          first = false;
          ++G;
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
    const bool grow = a.fuse_q && slot == 0 && l >= kGRow0 && l < kGRow0 + geo.g;
    const int ga = l - kGRow0;
    constexpr int NC = DP / 2;
    uint32_t uc = 0, G = 0;
    for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
      const int bh = unit / units_per_bh, rem = unit % units_per_bh;
      const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
      const int C = 2 * Cp + slot;
      const int r = R * W + qr, c = C * W + qc;
      const bool slot_ok = C < geo.my;
      const bool row_ok = slot_ok && l < W2 && r < geo.nx && c < geo.ny;
      const long long tok = (long long)r * geo.ny + c;
      // ---- delta = dO . O of my row (both column halves compute half of it and exchange through shared memory);
      //      lse2 from the token-ordered forward output; both re-emitted chunk-ordered for pass 2
      float lse2 = INFINITY, dpart = 0.f;
      if (row_ok || grow) {
        float ov[NC], gv[NC];
        if (grow) {
          load_half_row<NC, TO>(a.og, b, h, ga, geo.D, half * NC, ov);
          load_half_row<NC, TE>(a.d_og, b, h, ga, geo.D, half * NC, gv);
          lse2 = a.lse_g[(long long)bh * geo.g + ga] * 1.4426950408889634f;
        } else {
          load_half_row<NC, TO>(a.o, b, h, tok, geo.D, half * NC, ov);
          load_half_row<NC, TE>(a.d_o, b, h, tok, geo.D, half * NC, gv);
          lse2 = a.lse[(long long)bh * geo.Nloc + tok] * 1.4426950408889634f;
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) dpart = fmaf(ov[i], gv[i], dpart);
      }
      dx[half * 128 + row] = dpart;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float del = dx[row] + dx[128 + row];
      asm volatile("bar.sync 2, 256;" ::: "memory");                  // dx is rewritten by the next unit
      if (half == 0 && slot_ok) {
        const long long ci = (((long long)bh * geo.mx + R) * geo.my + C) * 64 + l;
        a.lse2c[ci] = row_ok ? lse2 : INFINITY;
        a.deltac[ci] = row_ok ? del : 0.f;
      }
      // 16-padded log2-domain lse / delta of the global QUERY rows for pass 2's first block: unit (0,0) of every (b,h)
      if (geo.g > 0 && R == 0 && Cp == 0 && slot == 0 && half == 0 && l < 16) {
        float lg = INFINITY, dg = 0.f;
        if (l < geo.g) {
          float ov[NC], gv[NC];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            load_half_row<NC, TO>(a.og, b, h, l, geo.D, hh * NC, ov);
            load_half_row<NC, TE>(a.d_og, b, h, l, geo.D, hh * NC, gv);
#pragma unroll
            for (int i = 0; i < NC; ++i) dg = fmaf(ov[i], gv[i], dg);
          }
          lg = a.lse_g[(long long)bh * geo.g + l] * 1.4426950408889634f;
        }
        a.lse2g[bh * 16 + l] = lg;
        a.deltag[bh * 16 + l] = dg;
      }
      const float* tab_h = tab + h * tabn;
      BlockWalk wk; wk.init(geo, R, Cp);
      int type, KR, KC;
      while (wk.next(geo, type, KR, KC)) {
        mbar_wait((bars + 8u * (BB_SFULL)), G & 1);
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
            // global keys: every local row; the global query rows only in unit (0,0) (the g x g corner counted once)
            const float nl = (grow && !(R == 0 && Cp == 0)) ? -INFINITY : -lse2;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float d0 = 0.f, d1 = 0.f;
              if (j < geo.g) {
                const float p = fast_exp2(fmaf(__uint_as_float(s[j]), a.scale_log2, nl));
                d0 = p * (__uint_as_float(dp[j]) - del);
              }
              if (j + 1 < geo.g) {
                const float p = fast_exp2(fmaf(__uint_as_float(s[j + 1]), a.scale_log2, nl));
                d1 = p * (__uint_as_float(dp[j + 1]) - del);
              }
              pk[j >> 1] = pack2<BF16>(d0, d1);
            }
            tmem_st_x8(dsaddr, pk);
          }
        } else {
          const int dR = KR - R, dC = KC - C;
          const bool use = wk.used_by(slot);
          int krows = W, kcols = W;
          bool masked = false, ht = false;
          if constexpr (!LEAN) {
            krows = min(W, geo.nx - KR * W); kcols = min(W, geo.ny - KC * W);
            masked = (krows < W) || (kcols < W);
            ht = a.has_tab != 0;
          }
          if (!use) {
            uint32_t pk[16];
            tc_fence_before();
            mbar_arrive((bars + 8u * (BB_CONS)));
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
            tmem_st_x16(dsaddr + half * 16, pk);
          } else {
            // global query rows see only the chunks this unit owns; -inf switches the row off (P = 0, dS = 0)
            const bool own = (KR == R) && (KC == 2 * Cp || KC == 2 * Cp + 1);
            const float nl = (grow && !own) ? -INFINITY : -lse2;
            uint32_t pk[16];
            const float* tb = nullptr;
            if constexpr (!LEAN) tb = grow ? (zpad + ZPAD - 1) : (tab_h + ((qr - dR * W + 2 * W - 1) * TW + (qc - dC * W + 2 * W - 1)));
            if (half == 0) {
              dq_quarter<W, 0, BF16, LEAN>(pk, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, nl, del, 0u);
              dq_quarter<W, 16, BF16, LEAN>(pk + 8, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, nl, del, (bars + 8u * (BB_CONS)));
            } else {
              dq_quarter<W, 32, BF16, LEAN>(pk, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, nl, del, 0u);
              dq_quarter<W, 48, BF16, LEAN>(pk + 8, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, nl, del, (bars + 8u * (BB_CONS)));
            }
            tmem_st_x16(dsaddr + half * 16, pk);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive((bars + 8u * (BB_DSFULL + (G & 1))));
        ++G;
      }
      mbar_wait((bars + 8u * (BB_ACCDONE)), uc & 1);
      tc_fence_after();
      uint32_t ov[NC];
      if constexpr (NC == 32) tmem_ld_x32(TM_ACC + lane_base + half * NC, ov); else tmem_ld_x16(TM_ACC + lane_base + half * NC, ov);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive((bars + 8u * (BB_ACCFREE)));
      if (row_ok) {
        store_cols<NC, BF16>(a.out0, b, h, tok, geo.D, half * NC, ov, a.scale, a.out_f32);
      } else if (grow) {
        float* dst = a.part + (((long long)bh * units_per_bh + rem) * kGMax + ga) * DP + half * NC;
#pragma unroll
        for (int j = 0; j < NC; ++j) dst[j] = __uint_as_float(ov[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

// pass-2 element work for 16 query columns (thread = key row).  A global key row is switched off on query chunks its unit
// does not own by pointing `ls` at an all-(+inf) array (P = 0, dS = 0): no extra instruction per element.
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
        if (j < NV) {
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
template <int W, int COL0, bool BF16, bool LEAN = false, int NV = W * W>
__device__ __forceinline__ void dkv_quarter(uint32_t* __restrict__ pp, uint32_t* __restrict__ pd, uint32_t saddr, uint32_t paddr,
                                            float c, bool has_tab, const float* __restrict__ tb, bool use,
                                            const float* __restrict__ ls, const float* __restrict__ dl, uint32_t cons_bar) {
  uint32_t s[16], dp[16];
  tmem_ld_x16(saddr + COL0, s);
  tmem_ld_x16(paddr + COL0, dp);
  tmem_ld_wait();
  if (cons_bar != 0u) { tc_fence_before(); mbar_arrive(cons_bar); }
  if (!use) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { pp[j] = 0u; pd[j] = 0u; }
    return;
  }
  if (!LEAN && has_tab) dkv_cols16<W, COL0, BF16, true, NV>(pp, pd, s, dp, c, tb, ls, dl);
  else                  dkv_cols16<W, COL0, BF16, false, NV>(pp, pd, s, dp, c, tb, ls, dl);
}

// ======================================================================================================== pass 2
template <int DP, int W, bool BF16, bool LEAN>
__global__ void __launch_bounds__(kBwdThreads, 2)
vil_tc_bwd2_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                       const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const __grid_constant__ CUtensorMap tmQg, const __grid_constant__ CUtensorMap tmDOg,
                       const __grid_constant__ CUtensorMap tmKg, const __grid_constant__ CUtensorMap tmVg, const Args a) {
  using SM = Smem2<DP>;
  constexpr int ROWB = SM::ROWB, NS = SM::NS;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int W2 = W * W, TW = 4 * W - 1;
  constexpr int ZP2 = (W - 1) * TW + W;
  const Geo& geo = a.geo;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sX = smem + SM::OFF_X;                 // [buf][K | V]
  unsigned char* sY = smem + SM::OFF_Y;
  float* tab = reinterpret_cast<float*>(smem + SM::OFF_TAB);
  const int tabn = a.has_tab ? TW * TW : 0;
  float* zpad = tab + geo.H * tabn;
  const int infs_off = (geo.H * tabn + (a.has_tab ? ZP2 : 0) + 3) & ~3;      // float4-aligned: read as float4 broadcasts
  float* infs = tab + infs_off;                             // [64] +inf: the "lse" of a switched-off row
  const int bars_off = (SM::OFF_TAB + (infs_off + 64) * 4 + 15) & ~15;
  uint64_t* bars_p = reinterpret_cast<uint64_t*>(smem + bars_off);
  const uint32_t bars = smem_u32(bars_p);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_p + BB_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < SM::OFF_TAB / 16; i += kBwdThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < geo.H * tabn; i += kBwdThreads) {
    const int idx = i % tabn;
    const int dr = idx / TW - (2 * W - 1), dc = idx % TW - (2 * W - 1);
    tab[i] = (geo.exact == 1 && (abs(dr) > W || abs(dc) > W)) ? -INFINITY : 0.f;
  }
  if (a.has_tab) for (int i = tid; i < ZP2; i += kBwdThreads) zpad[i] = 0.f;
  for (int i = tid; i < 64; i += kBwdThreads) infs[i] = INFINITY;
  if (tid == 0) init_bwd_barriers(bars, NS);
  if (warp == 8) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
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
        mbar_arrive_expect_tx((bars + 8u * (BB_XFULL + xb)), ((hasB ? 4 : 2) * W2 + (a.fuse_q ? 16 : 0)) * ROWB);
        tma_load_5d(sK, &tmK, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp) * W, R * W, h, b);
        tma_load_5d(sV, &tmV, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp) * W, R * W, h, b);
        if (hasB) {
          tma_load_5d(sK + 64 * ROWB, &tmK, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp + 1) * W, R * W, h, b);
          tma_load_5d(sV + 64 * ROWB, &tmV, (bars + 8u * (BB_XFULL + xb)), 0, (2 * Cp + 1) * W, R * W, h, b);
        }
        if (a.fuse_q) {                                       // the global KEY rows ride in lanes 56..63 of slot A
          tma_load_4d(sK + kGRow0 * ROWB, &tmKg, (bars + 8u * (BB_XFULL + xb)), 0, 0, h, b);
          tma_load_4d(sV + kGRow0 * ROWB, &tmVg, (bars + 8u * (BB_XFULL + xb)), 0, 0, h, b);
        }
        QueryWalk wk; wk.init(geo, R, Cp);
        int QR, QC;
        bool gpend = a.fuse_g != 0;
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
        bool glob = a.fuse_g != 0;
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
          uint64_t qacc[4], gacc[4];
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
            mbar_wait((bars + 8u * (BB_CONS)), G & 1);
            tc_fence_after();
            issue_SdP(false);
          }
          mbar_wait((bars + 8u * (BB_DSFULL + (G & 1))), (G >> 1) & 1);
          if (first && uc > 0) mbar_wait((bars + 8u * (BB_ACCFREE)), (uc - 1) & 1);
          tc_fence_after();
          if (cur_glob) {
            mma_ts(TM_DV, TM_P, gacc[0], IDESC_ACC, !first);
            mma_ts(TM_DK, TM_DS, qacc[0], IDESC_ACC, !first);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ts(TM_DV, TM_P + k * 8, gacc[k], IDESC_ACC, (!first) || k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ts(TM_DK, TM_DS + k * 8, qacc[k], IDESC_ACC, (!first) || k > 0);
          }
          mma_commit((bars + 8u * (BB_YEMPTY + cur_stage)));
          if (kSplit) mma_commit((bars + 8u * (BB_PDONE)));
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
    const bool grow = a.fuse_q && slot == 0 && l >= kGRow0 && l < kGRow0 + geo.g;     // this lane is a global KEY row
    const int ga = l - kGRow0;
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
      int QR = 0, QC = 0;
      bool gpend = a.fuse_g != 0;
      while (gpend || wk.next(geo, QR, QC)) {
        const bool glob = gpend;
        gpend = false;
        mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase);
        mbar_wait((bars + 8u * (BB_SFULL)), G & 1);
        tc_fence_after();
        const float* ls = reinterpret_cast<const float*>(sY + stage * SM::STAGE_STRIDE + 2 * SM::Y_BYTES);
        const float* dl = ls + 64;
        const uint32_t saddr = TM_S + lane_base, paddr = TM_DP + lane_base;
        const int dR = R - QR, dC = C - QC;
        const bool use_w = wk.used_by(slot);
        const bool use = use_w && (row_ok || grow);
        uint32_t pp[16], pd[16];
        if (glob) {
          uint32_t s[16], dp[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { pp[j] = 0u; pd[j] = 0u; s[j] = 0u; dp[j] = 0u; }
          if (half == 0) {
            tmem_ld_x16(saddr, s);
            tmem_ld_x16(paddr, dp);
            tmem_ld_wait();
          }
          if (kSplit) { tc_fence_before(); mbar_arrive((bars + 8u * (BB_CONS))); }
          // local key rows: all global queries;  global key rows: the g x g corner, counted by unit (0,0) only
          if (half == 0 && (row_ok || (grow && R == 0 && Cp == 0))) {
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
          // global key rows only collect from the query chunks this unit owns
          const bool own = (QR == R) && (QC == 2 * Cp || QC == 2 * Cp + 1);
          const float* lsx = (grow && !own) ? infs : ls;
          const float* tb = nullptr;
          bool ht = false;
          if constexpr (!LEAN) {
            tb = grow ? zpad : (tab_h + ((2 * W - 1 - dR * W - kr) * TW + (2 * W - 1 - dC * W - kc)));
            ht = a.has_tab != 0;
          }
          const uint32_t cb = kSplit ? (bars + 8u * (BB_CONS)) : 0u;
          if (half == 0) {
            dkv_quarter<W, 0, BF16, LEAN>(pp, pd, saddr, paddr, a.scale_log2, ht, tb, use, lsx, dl, 0u);
            dkv_quarter<W, 16, BF16, LEAN>(pp + 8, pd + 8, saddr, paddr, a.scale_log2, ht, tb, use, lsx, dl, cb);
          } else {
            dkv_quarter<W, 32, BF16, LEAN>(pp, pd, saddr, paddr, a.scale_log2, ht, tb, use, lsx, dl, 0u);
            dkv_quarter<W, 48, BF16, LEAN>(pp + 8, pd + 8, saddr, paddr, a.scale_log2, ht, tb, use, lsx, dl, cb);
          }
        }
        if (kSplit) {
          if (G > 0) { mbar_wait((bars + 8u * (BB_PDONE)), (G - 1) & 1); tc_fence_after(); }
        } else {
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        tmem_st_x16(TM_P + lane_base + half * 16, pp);
        tmem_st_x16(TM_DS + lane_base + half * 16, pd);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive((bars + 8u * (BB_DSFULL + (G & 1))));
        ++G;
        if (++stage == NS) { stage = 0; yphase ^= 1; }
      }
      mbar_wait((bars + 8u * (BB_ACCDONE)), uc & 1);
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
        if (row_ok) {
          store_cols<32, BF16>(out, b, h, tok, geo.D, q4 * 32, ov, f, a.out_f32);
        } else if (grow) {
          float* dst = a.part + ((((long long)bh * units_per_bh + rem) * kGMax + ga) * 2 + half) * DP + q4 * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(ov[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

// Sum the per-unit partials of the global rows: dq_g from the pass-1 partials [bh][unit][8][DP] (x scale), dk / dv of the
// global KEY rows from the pass-2 partials [bh][unit][8][2][DP] (dk x scale).  One thread per output channel.
template <typename TO>
__global__ void vil_tc_bwd2_merge(Geo geo, const float* __restrict__ part1, const float* __restrict__ part2, int units_per_bh,
                                  int DP, float scale, int do_q, int do_kv, T4 dqg, T4 dk, T4 dv) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= geo.B * geo.H * geo.g * 3 * DP) return;
  const int ch = idx % (3 * DP), a = (idx / (3 * DP)) % geo.g, bh = idx / (3 * DP * geo.g);
  const int b = bh / geo.H, h = bh % geo.H, d = ch % DP, which = ch / DP;       // 0: dq_g, 1: dk, 2: dv
  if (d >= geo.D || (which == 0 && !do_q) || (which > 0 && !do_kv)) return;
  float acc = 0.f;
  if (which == 0) {
    const float* base = part1 + ((long long)bh * units_per_bh * kGMax + a) * DP + d;
    for (int u = 0; u < units_per_bh; ++u) acc += base[(long long)u * kGMax * DP];
    row_ptr_w<TO>(dqg, b, h, a)[d] = ElemTraits<TO>::from_f(acc * scale);
  } else {
    const float* base = part2 + (((long long)bh * units_per_bh * kGMax + a) * 2 + (which - 1)) * DP + d;
    for (int u = 0; u < units_per_bh; ++u) acc += base[(long long)u * kGMax * 2 * DP];
    if (which == 1) row_ptr_w<TO>(dk, b, h, a)[d] = ElemTraits<TO>::from_f(acc * scale);
    else            row_ptr_w<TO>(dv, b, h, a)[d] = ElemTraits<TO>::from_f(acc);
  }
}

}  // namespace b2
}  // namespace tc
}  // namespace vil
