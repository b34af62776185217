// Fused tcgen05 / TMA forward, high-occupancy variant (sm_100a), chunk size w <= 8: local queries and global query rows in
// one kernel like vil_tc_fwd2.cuh, but built for FOUR resident CTAs per SM instead of two.
//
// Why: ncu on the round-1 kernel and on vil_tc_fwd2 shows the same picture - 2 softmax warps per scheduler, issue slots
// ~45 % busy, XU (ex2) ~45 %, every warp waiting on its own fixed per-block latencies (mbarrier wake-up, TMEM load round
// trip, MUFU / FMNMX dependency chains, hand-over fence + arrive, MMA round trip).  No pipe is the limit; the per-warp latency
// chain is, and the only thing that hides it is more warps.  TMEM (512 columns) is what capped residency at 2 CTAs:
//   round 1 / fwd2 : S double/triple buffer + O  = 256 columns per CTA
//   here           : ONE 64-column S buffer + O  = 128 columns per CTA  -> 4 CTAs, 16 softmax warps per SM.
// With a single S buffer the MMA round trip (P_j -> PV_j -> S_{j+1}) is exposed per CTA - the other three CTAs cover it.
// Registers: 65536 / (4 x 192) = 85 per thread, so the softmax streams S through registers 16 columns at a time
// (loads one step ahead) instead of holding the whole row:
//   bf16 (OPTIMISTIC): ONE pass per block - p = 2^(s c - m) against the running maximum m of the PREVIOUS blocks, P written
//     over S step by step, the block maximum tracked on the fly; if it beats m by more than 2^8 the O accumulator is rescaled
//     AFTER the block (deferred: the block's own P stays valid, bf16 has the range for it).  The first block of a unit takes
//     one extra max-only pass so that m starts exact.
//   fp16 (EXACT): two passes per block (maximum, then exponentials) - P <= 2^8 always, fp16-safe.
// Work decomposition, global rows in the spare lanes 56..63 of slot A, partial (m, l, O) per unit + merge kernel: exactly as
// in vil_tc_fwd2.cuh (whose Args / merge kernel are reused).  Blocks are walked like round 1 (BlockWalk: global-key tile,
// then the <= 12 chunks of the 3 x 4 window; a slot that does not visit a chunk writes P = 0).
#pragma once
#include "vil_tc_fwd2.cuh"
#include "vil_tc_bwd.cuh"        // store_cols

namespace vil {
namespace tc {
namespace f3 {

using namespace sm100;
using f2::Args;
using f2::kGRow0;
using f2::kGMax;

constexpr int kThreads3 = 192;          // warps 0-3 softmax, 4 TMA producer, 5 MMA issuer

template <int DP, bool HAS_TAB>
struct Smem {
  static constexpr int ROWB = DP * 2;
  static constexpr int NQ = DP == 32 ? 2 : 1;                           // Q tile buffers
  static constexpr int NSTG = DP == 32 ? (HAS_TAB ? 3 : 4) : 2;         // K/V ring depth
  static constexpr int Q_BYTES = 128 * ROWB;
  static constexpr int KV_BYTES = 64 * ROWB;
  static constexpr int STAGE_BYTES = 2 * KV_BYTES;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = NQ * Q_BYTES;
  static constexpr int OFF_TAB = OFF_KV + NSTG * STAGE_BYTES;
  static __host__ __device__ int total(int tab_floats) { return OFF_TAB + tab_floats * 4 + 256 + 1024; }
};

template <int DP, bool HAS_TAB>
struct Bars {
  static constexpr int NQ = Smem<DP, HAS_TAB>::NQ, NSTG = Smem<DP, HAS_TAB>::NSTG;
  enum { QFULL = 0, QEMPTY = NQ, KVFULL = 2 * NQ, KVEMPTY = KVFULL + NSTG, SFULL = KVEMPTY + NSTG, PFULL = SFULL + 1,
         PVDONE = PFULL + 1, OFREE = PVDONE + 1, SCONS = OFREE + 1, COUNT = SCONS + 1 };
};

// One 16-column step of one block for one thread (= one TMEM lane = one query row).
// DOEXP = false: only the maximum of the logits (mx) is updated.
// DOEXP = true : p = 2^(logit + add) is accumulated into `sum`, packed into p8; mx is updated as well.
// Logit of column j: s c (+ table[j]) (masked columns: -inf).  Without table / mask the maximum is tracked on the raw
// scores (c > 0) and the caller converts.
template <int W, int COL0, int N, bool BF16, bool HAS_TAB, bool MASKED, bool DOEXP>
__device__ __forceinline__ void step16(const uint32_t (&s)[16], float c, float add, const float* __restrict__ tb, int krows,
                                       int kcols, float (&mx)[2], float (&sum)[2], uint32_t (&p8)[8]) {
  constexpr int TW = 4 * W - 1;
#pragma unroll
  for (int jj = 0; jj < 16; jj += 2) {
    const int j = COL0 + jj;
    float p0 = 0.f, p1 = 0.f;
    if (jj < N) {
      const bool two = jj + 1 < N;
      float s0 = __uint_as_float(s[jj]), s1 = two ? __uint_as_float(s[jj + 1]) : 0.f;
      if constexpr (HAS_TAB || MASKED) {
        float x0, x1;
        const float t0 = HAS_TAB ? tb[-((j / W) * TW + (j % W))] : 0.f;
        const float t1 = (HAS_TAB && two) ? tb[-(((j + 1) / W) * TW + ((j + 1) % W))] : 0.f;
        ffma2(x0, x1, s0, s1, c, c, t0, t1);
        if constexpr (MASKED) {
          x0 = ((j / W) < krows && (j % W) < kcols) ? x0 : -INFINITY;
          x1 = (two && ((j + 1) / W) < krows && ((j + 1) % W) < kcols) ? x1 : -INFINITY;
        } else if (!two) {
          x1 = -INFINITY;
        }
        mx[(jj >> 1) & 1] = f2::fmax3(mx[(jj >> 1) & 1], x0, x1);
        if constexpr (DOEXP) {
          fadd2(x0, x1, x0, x1, add, add);
          p0 = fast_exp2(x0);
          p1 = fast_exp2(x1);                       // 2^-inf = 0 for the masked / missing column
        }
      } else {
        if (two) mx[(jj >> 1) & 1] = f2::fmax3(mx[(jj >> 1) & 1], s0, s1);
        else     mx[(jj >> 1) & 1] = fmaxf(mx[(jj >> 1) & 1], s0);
        if constexpr (DOEXP) {
          float x0, x1;
          ffma2(x0, x1, s0, s1, c, c, add, add);
          p0 = fast_exp2(x0);
          p1 = two ? fast_exp2(x1) : 0.f;
        }
      }
      if constexpr (DOEXP) fadd2(sum[0], sum[1], sum[0], sum[1], p0, p1);
    }
    if constexpr (DOEXP) p8[jj >> 1] = pack2<BF16>(p0, p1);
  }
}

// EXACT = false (bf16): optimistic single pass + deferred rescale;  true (fp16): two passes, P <= 2^8.
// LEAN: no padded chunks in this geometry -> the masked code paths are compiled out.
template <int DP, int W, bool BF16, bool HAS_TAB, bool EXACT, bool LEAN>
__global__ void __launch_bounds__(kThreads3, 4)
vil_tc_fwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQg,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmKg, const __grid_constant__ CUtensorMap tmVg, const Args a) {
  using SM = Smem<DP, HAS_TAB>;
  using BB = Bars<DP, HAS_TAB>;
  constexpr int ROWB = SM::ROWB, NSTG = SM::NSTG, NQ = SM::NQ;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int W2 = W * W, TW = 4 * W - 1;
  constexpr int NCH = (W2 + 15) / 16, TAIL = W2 - 16 * (NCH - 1);
  constexpr int ZPAD = (W - 1) * TW + W;
  constexpr uint32_t TMEM_COLS = 128;
  const Geo& geo = a.geo;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sQ = smem + SM::OFF_Q;
  unsigned char* sKV = smem + SM::OFF_KV;
  float* tab = reinterpret_cast<float*>(smem + SM::OFF_TAB);
  const int tabn = HAS_TAB ? TW * TW : 0;
  float* zpad = tab + geo.H * tabn;
  float* g2l_s = zpad + (HAS_TAB ? ZPAD : 0);                             // [H][16]
  float* bg_s = g2l_s + geo.H * 16;                                       // [H][8]
  float* g2g_s = bg_s + geo.H * 8;                                        // [H][8][16]
  const int nfl = geo.H * tabn + (HAS_TAB ? ZPAD : 0) + geo.H * (16 + 8 + 128);
  const int bars_off = (SM::OFF_TAB + nfl * 4 + 15) & ~15;
  uint64_t* bars_p = reinterpret_cast<uint64_t*>(smem + bars_off);
  const uint32_t bars = smem_u32(bars_p);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_p + BB::COUNT);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr float L2E = 1.4426950408889634f;

  for (int i = tid; i < SM::OFF_TAB / 16; i += kThreads3) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if constexpr (HAS_TAB) {
    for (int i = tid; i < geo.H * tabn; i += kThreads3) {
      const int h = i / tabn, idx = i % tabn;
      const int dr = idx / TW - (2 * W - 1), dc = idx % TW - (2 * W - 1);
      float v = (a.table != nullptr) ? a.table[(long long)idx * geo.H + h] * L2E : 0.f;
      if (geo.exact == 1 && (abs(dr) > W || abs(dc) > W)) v = -INFINITY;
      tab[i] = v;
    }
    for (int i = tid; i < ZPAD; i += kThreads3) zpad[i] = 0.f;
  }
  for (int i = tid; i < geo.H * 16; i += kThreads3) {
    const int h = i / 16, t = i % 16;
    g2l_s[i] = (a.g2l != nullptr && t < geo.g) ? a.g2l[((long long)geo.H + h) * geo.g + t] * L2E : 0.f;
  }
  for (int i = tid; i < geo.H * 8; i += kThreads3) {
    const int h = i / 8, t = i % 8;
    bg_s[i] = (a.g2l != nullptr && t < geo.g) ? a.g2l[(long long)h * geo.g + t] * L2E : 0.f;
  }
  for (int i = tid; i < geo.H * 128; i += kThreads3) {
    const int h = i / 128, aa = (i % 128) / 16, bb = i % 16;
    g2g_s[i] = (a.g2g != nullptr && aa < geo.g && bb < geo.g) ? a.g2g[((long long)h * geo.g + aa) * geo.g + bb] * L2E : 0.f;
  }
  if (tid == 0) {
    for (int i = 0; i < NQ; ++i) { mbar_init(bar(BB::QFULL + i), 1); mbar_init(bar(BB::QEMPTY + i), 1); }
    for (int i = 0; i < NSTG; ++i) { mbar_init(bar(BB::KVFULL + i), 1); mbar_init(bar(BB::KVEMPTY + i), 1); }
    mbar_init(bar(BB::SFULL), 1); mbar_init(bar(BB::PFULL), 128); mbar_init(bar(BB::PVDONE), 1); mbar_init(bar(BB::OFREE), 128);
    mbar_init(bar(BB::SCONS), 128);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // SPLITP (D <= 32 only: O needs 32 columns, so P could get its OWN 32 columns and S_{j+1} be issued as soon as every thread
  // has read S_j (SCONS) instead of after PV_j).  MEASURED (gpurun_out/r02_ab_g.log): 0.752 ms vs 0.731 ms without it on S1 -
  // the extra barrier traffic and the 80-register cap (spills) cost more than the exposed MMA round trip, which the other
  // three resident CTAs already cover.  Kept off; P overwrites S.
  constexpr bool SPLITP = false;
  const uint32_t TM_S = tmem, TM_O = tmem + 64, TM_P = SPLITP ? tmem + 96 : tmem;

  const int units_per_bh = geo.mx * a.cpairs;

  if (warp == 4) {
    // ================================================================= TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      uint32_t stage = 0, kv_phase = 0, uc = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh = unit / units_per_bh, rem = unit % units_per_bh;
        const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
        const uint32_t qb = uc % NQ, qphase = (uc / NQ) & 1;
        if (uc >= NQ) mbar_wait(bar(BB::QEMPTY + qb), qphase ^ 1);
        const bool hasB = 2 * Cp + 1 < geo.my;
        unsigned char* q0 = sQ + qb * SM::Q_BYTES;
        mbar_arrive_expect_tx(bar(BB::QFULL + qb), ((hasB ? 2 : 1) * W2 + (a.fuse_g ? 8 : 0)) * ROWB);
        tma_load_5d(q0, &tmQ, bar(BB::QFULL + qb), 0, (2 * Cp) * W, R * W, h, b);
        if (hasB) tma_load_5d(q0 + 64 * ROWB, &tmQ, bar(BB::QFULL + qb), 0, (2 * Cp + 1) * W, R * W, h, b);
        if (a.fuse_g) tma_load_4d(q0 + kGRow0 * ROWB, &tmQg, bar(BB::QFULL + qb), 0, 0, h, b);
        BlockWalk wk; wk.init(geo, R, Cp);
        int type, KR, KC;
        while (wk.next(geo, type, KR, KC)) {
          mbar_wait(bar(BB::KVEMPTY + stage), kv_phase ^ 1);
          unsigned char* dK = sKV + stage * SM::STAGE_BYTES;
          unsigned char* dV = dK + SM::KV_BYTES;
          if (type == 1) {
            mbar_arrive_expect_tx(bar(BB::KVFULL + stage), 2 * 16 * ROWB);
            tma_load_4d(dK, &tmKg, bar(BB::KVFULL + stage), 0, 0, h, b);
            tma_load_4d(dV, &tmVg, bar(BB::KVFULL + stage), 0, 0, h, b);
          } else {
            mbar_arrive_expect_tx(bar(BB::KVFULL + stage), 2 * W2 * ROWB);
            tma_load_5d(dK, &tmK, bar(BB::KVFULL + stage), 0, KC * W, KR * W, h, b);
            tma_load_5d(dV, &tmV, bar(BB::KVFULL + stage), 0, KC * W, KR * W, h, b);
          }
          if (++stage == NSTG) { stage = 0; kv_phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ================================================================= MMA issuer (one elected thread)
    if (elect_one()) {
      constexpr uint32_t IDESC_S = make_idesc(128, 64, BF16, false, false);
      constexpr uint32_t IDESC_SG = make_idesc(128, 16, BF16, false, false);
      constexpr uint32_t IDESC_O = make_idesc(128, DP, BF16, false, true);
      constexpr int KS = DP / 16;
      uint32_t stage = 0, kv_phase = 0, uc = 0, G = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int rem = unit % units_per_bh;
        const int R = rem / a.cpairs, Cp = rem % a.cpairs;
        const uint32_t qb = uc % NQ, qphase = (uc / NQ) & 1;
        mbar_wait(bar(BB::QFULL + qb), qphase);
        const uint32_t qaddr = smem_u32(sQ + qb * SM::Q_BYTES);
        uint64_t qd[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) qd[k] = make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT);
        auto issue_S = [&](uint32_t st, int type) {
          const uint32_t kaddr = smem_u32(sKV + st * SM::STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < KS; ++k)
            mma_ss(TM_S, qd[k], make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT), type == 1 ? IDESC_SG : IDESC_S, k > 0);
          mma_commit(bar(BB::SFULL));
        };
        BlockWalk wk; wk.init(geo, R, Cp);
        int type, KR, KC;
        bool have = wk.next(geo, type, KR, KC);
        mbar_wait(bar(BB::KVFULL + stage), kv_phase);
        tc_fence_after();
        issue_S(stage, type);
        bool first = true;
        while (have) {
          const uint32_t cur_stage = stage;
          const int cur_type = type;
          uint64_t vdsc[4];
          {
            const uint32_t vaddr = smem_u32(sKV + cur_stage * SM::STAGE_BYTES + SM::KV_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) vdsc[k] = make_smem_desc(vaddr + k * 16 * ROWB, 16, SBO, LAYOUT);
          }
          if (++stage == NSTG) { stage = 0; kv_phase ^= 1; }
          have = wk.next(geo, type, KR, KC);
          if (have) mbar_wait(bar(BB::KVFULL + stage), kv_phase);          // K tile of the next block
          if (SPLITP && have) {
            mbar_wait(bar(BB::SCONS), G & 1);                              // every thread holds S_j in registers
            tc_fence_after();
            issue_S(stage, type);                                          // overlaps the exponentials / P hand-over of block j
          }
          mbar_wait(bar(BB::PFULL), G & 1);
          if (first && uc > 0) mbar_wait(bar(BB::OFREE), (uc - 1) & 1);
          tc_fence_after();
          if (cur_type == 1) {
            mma_ts(TM_O, TM_P, vdsc[0], IDESC_O, !first);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ts(TM_O, TM_P + k * 8, vdsc[k], IDESC_O, (!first) || k > 0);
          }
          mma_commit(bar(BB::KVEMPTY + cur_stage));
          mma_commit(bar(BB::PVDONE));
          ++G;
          first = false;
          if (have) { if (!SPLITP) issue_S(stage, type); }    // P over S: S_{j+1} executes after PV_j on the in-order tensor pipe
          else mma_commit(bar(BB::QEMPTY + qb));
        }
      }
    }
  } else {
    // ================================================================= softmax warps (thread = TMEM lane)
    const int row = tid;
    const int slot = row >> 6, l = row & 63;
    const int qr = l / W, qc = l % W;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool grow = a.fuse_g && slot == 0 && l >= kGRow0 && l < kGRow0 + geo.g;
    const int ga = l - kGRow0;
    const float c = a.scale_log2;
    const uint32_t saddr = TM_S + lane_base, paddr = TM_P + lane_base;
    uint32_t uc = 0, G = 0;
    for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
      const int bh = unit / units_per_bh, rem = unit % units_per_bh;
      const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
      const int C = 2 * Cp + slot;
      const int r = R * W + qr, cc = C * W + qc;
      const bool row_ok = C < geo.my && l < W2 && r < geo.nx && cc < geo.ny;
      float m_use = -INFINITY, l_run = 0.f;
      const float* tab_h = tab + h * tabn;
      const float bias_g = grow ? bg_s[h * 8 + ga] : 0.f;
      // rescale O (and the running sum) by 2^(m_use - m_new); O is stable once the PV of block G_done has completed
      auto rescale = [&](bool need, float m_new, uint32_t G_done) {
        mbar_wait(bar(BB::PVDONE), G_done & 1);
        tc_fence_after();
        const float f = need ? fast_exp2(m_use - m_new) : 1.f;           // m_use == -inf -> 0
        if (need) { m_use = m_new; l_run *= f; }
#pragma unroll
        for (int q4 = 0; q4 < DP / 16; ++q4) {
          uint32_t ov[16];
          tmem_ld_x16(TM_O + lane_base + q4 * 16, ov);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * f);
          tmem_st_x16(TM_O + lane_base + q4 * 16, ov);
        }
        tmem_st_wait();
      };
      BlockWalk wk; wk.init(geo, R, Cp);
      int type, KR, KC;
      bool first = true;
      while (wk.next(geo, type, KR, KC)) {
        mbar_wait(bar(BB::SFULL), G & 1);
        tc_fence_after();
        float pend_m = -INFINITY;            // bf16: maximum seen in this block when it exceeds m_use + 8 (deferred rescale)
        // S_j has been read by this thread (SPLITP: lets the MMA warp overwrite S); before the first P store of the block the
        // previous block's PV must have finished reading the P columns
        auto s_consumed = [&]() { if constexpr (SPLITP) { tc_fence_before(); mbar_arrive(bar(BB::SCONS)); } };
        auto p_free = [&]() { if constexpr (SPLITP) { if (!first) { mbar_wait(bar(BB::PVDONE), (G - 1) & 1); tc_fence_after(); } } };
        if (type == 1) {
          // ---- global keys: 16 columns; local rows: bias g2l[1][h][t]; global rows (unit (0,0) only): g2g[h][a][t]
          uint32_t s[16];
          tmem_ld_x16(saddr, s);
          tmem_ld_wait();
          s_consumed();
          const bool gown = grow && R == 0 && Cp == 0;
          const float* brow = grow ? (g2g_s + h * 128 + ga * 16) : (g2l_s + h * 16);
          const float addg = (grow && !gown) ? -INFINITY : 0.f;
          float t[16], mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            t[j] = (j < geo.g) ? fmaf(__uint_as_float(s[j]), c, brow[j]) + addg : -INFINITY;
            mx = fmaxf(mx, t[j]);
          }
          // always the first block of a unit (BlockWalk order) when nglo > 0: exact initialisation
          const float m_new = fmaxf(m_use, mx);
          const bool need = !first && (m_new > m_use + 8.f);
          if (first) m_use = m_new;
          if (__any_sync(0xffffffffu, need)) rescale(need, m_new, G - 1);
          const float m_eff = (m_use == -INFINITY) ? 0.f : m_use;
          float sum = 0.f;
          uint32_t p8[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float p0v = fast_exp2(t[j] - m_eff), p1v = fast_exp2(t[j + 1] - m_eff);
            sum += p0v + p1v;
            p8[j >> 1] = pack2<BF16>(p0v, p1v);
          }
          l_run += sum;
          p_free();
          tmem_st_x8(paddr, p8);
        } else if (!wk.used_by(slot)) {
          uint32_t z[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) z[j] = 0u;
          s_consumed();
          p_free();
          tmem_st_x16(paddr, z);
          tmem_st_x16(paddr + 16, z);
        } else {
          const int dR = KR - R, dC = KC - C;
          int krows = W, kcols = W;
          bool masked = false;
          if constexpr (!LEAN) {
            krows = min(W, geo.nx - KR * W); kcols = min(W, geo.ny - KC * W);
            masked = (krows < W) || (kcols < W);
          }
          const bool own = (KR == R) && (KC == 2 * Cp || KC == 2 * Cp + 1);
          const float radd = grow ? (own ? bias_g : -INFINITY) : 0.f;
          const float* tb = nullptr;
          if constexpr (HAS_TAB) tb = grow ? (zpad + ZPAD - 1) : (tab_h + ((qr - dR * W + 2 * W - 1) * TW + (qc - dC * W + 2 * W - 1)));
          // step16 needs the compile-time column offset of every step (table index / mask row-col): the passes are written
          // out per step
          float mx[2] = {-INFINITY, -INFINITY}, sum[2] = {0.f, 0.f};
          auto run = [&](auto doexp_tag, float add) {
            constexpr bool DOEXP = decltype(doexp_tag)::value;
            uint32_t sa[16], sb2[16], p8[8];
            tmem_ld_x16(saddr, sa);
            // step 0
            tmem_ld_wait();
            if constexpr (NCH > 2) tmem_ld_x16(saddr + 16, sb2); else if constexpr (NCH == 2) f2::tmem_ld_n<TAIL>(saddr + 16, sb2);
            if (masked) { if constexpr (!LEAN) step16<W, 0, (NCH > 1 ? 16 : TAIL), BF16, HAS_TAB, true, DOEXP>(sa, c, add, tb, krows, kcols, mx, sum, p8); }
            else step16<W, 0, (NCH > 1 ? 16 : TAIL), BF16, HAS_TAB, false, DOEXP>(sa, c, add, tb, krows, kcols, mx, sum, p8);
            if constexpr (DOEXP) { if constexpr (NCH == 1) s_consumed(); p_free(); tmem_st_x8(paddr, p8); }
            if constexpr (NCH > 1) {        // step 1
              tmem_ld_wait();
              if constexpr (NCH > 3) tmem_ld_x16(saddr + 32, sa); else if constexpr (NCH == 3) f2::tmem_ld_n<TAIL>(saddr + 32, sa);
              if (masked) { if constexpr (!LEAN) step16<W, 16, (NCH > 2 ? 16 : TAIL), BF16, HAS_TAB, true, DOEXP>(sb2, c, add, tb, krows, kcols, mx, sum, p8); }
              else step16<W, 16, (NCH > 2 ? 16 : TAIL), BF16, HAS_TAB, false, DOEXP>(sb2, c, add, tb, krows, kcols, mx, sum, p8);
              if constexpr (DOEXP) { if constexpr (NCH == 2) s_consumed(); tmem_st_x8(paddr + 8, p8); }
            }
            if constexpr (NCH > 2) {        // step 2
              tmem_ld_wait();
              if constexpr (NCH > 3) f2::tmem_ld_n<TAIL>(saddr + 48, sb2);
              if (masked) { if constexpr (!LEAN) step16<W, 32, (NCH > 3 ? 16 : TAIL), BF16, HAS_TAB, true, DOEXP>(sa, c, add, tb, krows, kcols, mx, sum, p8); }
              else step16<W, 32, (NCH > 3 ? 16 : TAIL), BF16, HAS_TAB, false, DOEXP>(sa, c, add, tb, krows, kcols, mx, sum, p8);
              if constexpr (DOEXP) { if constexpr (NCH == 3) s_consumed(); tmem_st_x8(paddr + 16, p8); }
            }
            if constexpr (NCH > 3) {        // step 3
              tmem_ld_wait();
              if (masked) { if constexpr (!LEAN) step16<W, 48, TAIL, BF16, HAS_TAB, true, DOEXP>(sb2, c, add, tb, krows, kcols, mx, sum, p8); }
              else step16<W, 48, TAIL, BF16, HAS_TAB, false, DOEXP>(sb2, c, add, tb, krows, kcols, mx, sum, p8);
              if constexpr (DOEXP) { if constexpr (NCH == 4) s_consumed(); tmem_st_x8(paddr + 24, p8); }
            }
            if constexpr (DOEXP && 8 * NCH < 32) {                      // keys >= w*w of the K = 64 step: explicit zeros
              uint32_t z8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) z8[j] = 0u;
#pragma unroll
              for (int c8 = 8 * NCH; c8 < 32; c8 += 8) tmem_st_x8(paddr + c8, z8);
            }
          };
          auto row_max = [&]() -> float {      // block maximum in the log2 domain incl. the row addend
            const float m2 = fmaxf(mx[0], mx[1]);
            return ((HAS_TAB || masked) ? m2 : m2 * c) + radd;
          };
          // a row whose maximum is still -inf needs an exact start (warp-uniform decision; global rows on chunks they do not
          // own stay switched off and do not count)
          const bool init = (m_use == -INFINITY) && !(grow && !own);
          const bool two_pass = EXACT || __any_sync(0xffffffffu, init);
          if (two_pass) {
            // pass 1: maximum only; rescale BEFORE the exponentials so that P <= 2^8 (fp16-safe) / the start is exact
            run(std::false_type{}, 0.f);
            const float m_new = fmaxf(m_use, row_max());
            const bool need = !first && (m_new > m_use + 8.f);
            if (first) m_use = m_new;                                      // O has not been written yet in this unit
            if (__any_sync(0xffffffffu, need)) rescale(need, m_new, G - 1);
            mx[0] = mx[1] = -INFINITY;
          }
          const float m_eff = (m_use == -INFINITY) ? 0.f : m_use;
          run(std::true_type{}, radd - m_eff);
          l_run += sum[0] + sum[1];
          if (!two_pass) {
            const float m_new = row_max();
            if (m_new > m_use + 8.f) pend_m = m_new;
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar(BB::PFULL));
        if constexpr (!EXACT) {
          // deferred rescale: this block's P was produced against the old maximum and stays valid; O is brought to the new
          // maximum once PV of THIS block has completed, before the next block's P is handed over
          if (__any_sync(0xffffffffu, pend_m > -INFINITY)) {
            rescale(pend_m > -INFINITY, pend_m, G);
            tc_fence_before();
          }
        }
        first = false;
        ++G;
      }
      // ---- epilogue
      mbar_wait(bar(BB::PVDONE), (G - 1) & 1);
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      const long long tok = (long long)r * geo.ny + cc;
      float* part = grow ? a.part + (((long long)bh * units_per_bh + rem) * kGMax + ga) * (DP + 2) : nullptr;
#pragma unroll
      for (int q2 = 0; q2 < DP / 16; ++q2) {
        uint32_t ov[16];
        tmem_ld_x16(TM_O + lane_base + q2 * 16, ov);
        tmem_ld_wait();
        if (q2 == DP / 16 - 1) { tc_fence_before(); mbar_arrive(bar(BB::OFREE)); }
        if (row_ok) store_cols<16, BF16>(a.o, b, h, tok, geo.D, q2 * 16, ov, inv, a.out_f32);
        else if (grow) {
#pragma unroll
          for (int j = 0; j < 16; ++j) part[2 + q2 * 16 + j] = __uint_as_float(ov[j]);
        }
      }
      if (row_ok) a.lse[((long long)b * geo.H + h) * geo.Nloc + tok] = (m_use + log2f(l_run)) * 0.6931471805599453f;
      else if (grow) { part[0] = m_use; part[1] = l_run; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace f3
}  // namespace tc
}  // namespace vil
