// Fused tcgen05 / TMA forward of the Vision-Longformer attention (sm_100a), chunk size w <= 8: local queries AND the
// global query rows in one kernel (longformer2d.py:126-202 + 210-227), round-2 successor of vil_tc_fwd.cuh.
//
// Work unit   = (b, h, chunk-row R, pair of chunk columns {C0 = 2Cp, C0+1}): a 128-lane tile = two 64-lane slots
//               (slot A = chunk (R,C0), slot B = chunk (R,C0+1); w*w <= 64 real rows each).  With w <= 7 the lanes
//               56..63 of slot A are spare: they carry the (<= 8) GLOBAL QUERY rows, which then cost no extra MMA and no
//               extra pass over K / V (the round-1 simt_fwd_global kernel re-read all of K and V).
// Iterations  = every key chunk a slot visits, enumerated so that NO softmax lane idles:
//                 * chunk columns C0 and C0+1 are visited by both slots            -> one S tile (N = 64), both slots work;
//                 * column C0-1 is visited by slot A only, C0+2 by slot B only     -> PAIRED: two S tiles in two TMEM
//                   buffers, slot A reads the first, slot B the second, each zero-fills its rows of the other one
//                   (round 1 spent a whole block period per exclusive chunk with half of the lanes writing zeros);
//                 * the <= 16 global keys (one 16-column tile) last.
//               Order: (R,C0), (R,C0+1), (R; C0-1 | C0+2), then chunk rows R-1 and R+1 in the same pattern.  The first
//               iteration initialises the running maximum exactly (two passes over S in TMEM); all later iterations are
//               single-pass "optimistic" softmax against the stale maximum with a warp-uniform re-do when a logit exceeds it
//               by more than 2^8 (lazy rescale of O in TMEM).
// Global rows = lanes 56..56+g-1 of slot A.  They take part in the S / PV MMAs of every iteration but are given the
//               addend -inf (P = 0) except for the two chunks this unit OWNS ((R,C0), (R,C0+1)) and, in unit (0,0), the
//               global-key tile - so every key is counted exactly once per (b,h) across units.  The unit writes an
//               unnormalised partial (m, l, O) per global row; vil_tc_fwd2_merge combines the <= mx*cpairs partials.
// Pipeline    : TMA producer warp -> K/V ring (one chunk per stage) ; MMA warp: S ring of THREE 64-column TMEM buffers
//               (two S tiles ahead of the softmax) ; 4 softmax warps (thread = TMEM lane) ; O accumulator in TMEM.
#pragma once
#include "vil_tc_fwd.cuh"

namespace vil {
namespace tc {
namespace f2 {

using namespace sm100;

constexpr int kThreads2 = 192;          // warps 0-3 softmax, 4 TMA producer, 5 MMA issuer
constexpr int kGRow0 = 56;              // first spare lane of slot A (8-row aligned: a TMA box can land there)
constexpr int kGMax = 8;

struct Args {
  Geo geo;
  T4 o;
  float* lse;
  const float* table;             // ((4w-1)^2, H) fp32 or null
  const float* g2l;               // (2,H,g) fp32 or null
  const float* g2g;               // (H,g,g) fp32 or null
  float* part;                    // fused global rows: (B*H, units_per_bh, 8, DP + 2) fp32 partials
  int cpairs, num_units;
  int has_tab;                    // bias table in smem (rpe on, or exact == 1)
  int fuse_g;                     // global query rows ride in lanes 56.. of slot A
  int out_f32;
  float scale_log2;
};

template <int DP>
struct Smem {
  static constexpr int ROWB = DP * 2;
  static constexpr int NSTG = DP == 32 ? 6 : 4;          // K/V ring depth (one chunk = K tile + V tile per stage)
  static constexpr int Q_BYTES = 128 * ROWB;
  static constexpr int KV_BYTES = 64 * ROWB;
  static constexpr int STAGE_BYTES = 2 * KV_BYTES;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = 2 * Q_BYTES;
  static constexpr int OFF_TAB = OFF_KV + NSTG * STAGE_BYTES;
  static __host__ __device__ int total(int tab_floats) { return OFF_TAB + tab_floats * 4 + 512 + 1024; }
};

template <int DP>
struct Bars {
  static constexpr int NSTG = Smem<DP>::NSTG;
  enum { QFULL = 0, QEMPTY = 2, KVFULL = 4, KVEMPTY = 4 + NSTG, SFULL = 4 + 2 * NSTG, PFULL = SFULL + 3,
         PVDONE = PFULL + 3, OFREE = PVDONE + 3, COUNT = OFREE + 1 };
};

// ---------------------------------------------------------------------------------------------- iteration schedule
// One iteration = up to two key chunks: `a` for slot A, `b` for slot B (same chunk -> one tile, both slots).
struct Iter {
  int type;            // 0 local, 1 global-key tile
  bool hasA, hasB;     // slot takes part
  bool two;            // two distinct chunks -> two S buffers / two K-V stages
  int KR, KCa, KCb;
  bool own;            // a chunk this unit owns (global rows account for its keys)
};

struct Sched {
  int R, C0;
  bool hasB, left, right, up, down;
  int mode, nR, nC;            // random-shift neighbour offset (mode > 0)
  bool gl;                     // global-key tile present
  __device__ __forceinline__ void init(const Geo& g, int R_, int Cp) {
    R = R_; C0 = 2 * Cp;
    hasB = C0 + 1 < g.my;
    left = C0 > 0;
    right = hasB && C0 + 2 < g.my;
    up = R > 0; down = R + 1 < g.mx;
    mode = g.mode; nR = g.offR[1]; nC = g.offC[1];
    gl = g.g > 0;
  }
  __device__ __forceinline__ int count() const { return mode == 0 ? 10 : 3; }
  // The same schedule as bit masks over the iteration index (bit i): valid / two-chunk / "slot takes part".  The softmax
  // threads walk these with find-first-set instead of evaluating get() per iteration (per-iteration control code was 2.3x the
  // round-1 kernel's and cost more issue slots than the exponentials).
  __device__ __forceinline__ void masks(const Geo& g, int slot, uint32_t& vm, uint32_t& tm, uint32_t& pm) const {
    vm = tm = pm = 0;
    if (mode == 0) {
      const uint32_t rows = 1u | (up ? 8u : 0u) | (down ? 64u : 0u);                 // bit 3*r3
      const uint32_t kb = 1u | (hasB ? 2u : 0u) | ((left || right) ? 4u : 0u);
      const uint32_t tb = (left && right) ? 4u : 0u;
      const uint32_t pb = slot == 0 ? (1u | (hasB ? 2u : 0u) | (left ? 4u : 0u)) : (hasB ? (3u | (right ? 4u : 0u)) : 0u);
      vm = rows * kb; tm = rows * tb; pm = rows * pb;
      if (gl) { vm |= 512u; if (slot == 0 || hasB) pm |= 512u; }
      return;
    }
    for (int i = 0; i < 3; ++i) {
      Iter it;
      if (!get(g, i, it)) continue;
      vm |= 1u << i;
      if (it.two) tm |= 1u << i;
      if (slot == 0 ? it.hasA : it.hasB) pm |= 1u << i;
    }
  }
  // iteration i of the unit; false = nothing to do at this index
  __device__ __forceinline__ bool get(const Geo& g, int i, Iter& it) const {
    it.type = 0; it.two = false; it.own = false; it.hasA = it.hasB = false; it.KR = it.KCa = it.KCb = 0;
    if (mode == 0) {
      if (i == 9) { if (!gl) return false; it.type = 1; it.hasA = true; it.hasB = hasB; return true; }
      const int r3 = i / 3, k = i - 3 * r3;
      if ((r3 == 1 && !up) || (r3 == 2 && !down)) return false;       // no arrays: they would live in local memory
      it.KR = r3 == 0 ? R : (r3 == 1 ? R - 1 : R + 1);
      if (k == 0) { it.KCa = it.KCb = C0; it.hasA = true; it.hasB = hasB; it.own = (r3 == 0); return true; }
      if (k == 1) { if (!hasB) return false; it.KCa = it.KCb = C0 + 1; it.hasA = it.hasB = true; it.own = (r3 == 0); return true; }
      if (!left && !right) return false;
      it.KCa = C0 - 1; it.KCb = C0 + 2; it.hasA = left; it.hasB = right; it.two = left && right;
      return true;
    }
    if (i == 2) { if (!gl) return false; it.type = 1; it.hasA = true; it.hasB = hasB; return true; }
    if (i == 0) {                                       // own chunks: (R,C0) | (R,C0+1)
      it.KR = R; it.KCa = C0; it.KCb = C0 + 1; it.hasA = true; it.hasB = hasB; it.two = hasB; it.own = true;
      return true;
    }
    if (mode < 0) return false;                         // mode -1: own chunk only
    const int rr = R + nR;
    if (rr < 0 || rr >= g.mx) return false;
    it.KR = rr; it.KCa = C0 + nC; it.KCb = C0 + 1 + nC;
    it.hasA = it.KCa >= 0 && it.KCa < g.my;
    it.hasB = hasB && it.KCb >= 0 && it.KCb < g.my;
    if (!it.hasA && !it.hasB) return false;
    it.two = it.hasA && it.hasB;
    return true;
  }
};

// ring position of the S buffers (3) / K-V stages (N): slot + phase bit of the slot's current use
template <int N>
struct Ring {
  uint32_t i, ph;
  __device__ __forceinline__ void reset() { i = 0; ph = 0; }
  __device__ __forceinline__ void adv() { if (++i == N) { i = 0; ph ^= 1u; } }
};

// ---------------------------------------------------------------------------------------------- small math helpers
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[16]) {
  static_assert(N == 1 || N == 4 || N == 16, "tail widths of w in {6,7,8}");
  if constexpr (N == 16) {
    tmem_ld_x16(taddr, r);
  } else if constexpr (N == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr));
  }
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {       // sm_100: one FMNMX3
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// 2^x on the FMA / ALU pipes (Cody-Waite range reduction + degree-3 minimax, max rel. error 7.8e-5): takes a tunable
// fraction of the exponentials off the XU pipe (16 lanes/clk/SM), which is what bounds this kernel at D = 32.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;                       // 1.5 * 2^23: round-to-nearest integer in the low mantissa bits
  const float f = x - (t - 12582912.f);                 // [-0.5, 0.5]
  float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
  p = fmaf(p, f, 0.6932762265205383f);
  p = fmaf(p, f, 0.9999289512634277f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// ---------------------------------------------------------------------------------------------- the kernel
// POLY: every POLY-th exponential pair (0 = none) is evaluated by exp2_poly instead of ex2.approx.
// P16 : hand P to the PV MMA as fp16 even when q/k/v are bf16 (P lies in [0, 2^8], fp16's 11-bit mantissa would remove the
//       dominant error term of the forward).  NOT USABLE: mixed fp16-A / bf16-B operands trap on B200 (illegal instruction,
//       measured in round 2) - kept false; the parameter documents the experiment.
// LEAN: the geometry has no padded chunks and there is no table -> the general (edge / table) path is compiled out, which
//       keeps the kernel's code inside the instruction caches.
template <int DP, int W, bool BF16, bool HAS_TAB, int POLY, bool P16, bool LEAN = false>
__global__ void __launch_bounds__(kThreads2, 2)
vil_tc_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQg,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmKg, const __grid_constant__ CUtensorMap tmVg, const Args a) {
  using SM = Smem<DP>;
  using BB = Bars<DP>;
  constexpr int ROWB = SM::ROWB, NSTG = SM::NSTG;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int W2 = W * W, TW = 4 * W - 1;
  constexpr int NCH = (W2 + 15) / 16, TAIL = W2 - 16 * (NCH - 1);       // 16-column TMEM load steps, width of the last one
  constexpr int ZPAD = (W - 1) * TW + W;                                  // zero "table" of the global rows
  const Geo& geo = a.geo;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sQ = smem + SM::OFF_Q;
  unsigned char* sKV = smem + SM::OFF_KV;
  float* tab = reinterpret_cast<float*>(smem + SM::OFF_TAB);
  const int tabn = HAS_TAB ? TW * TW : 0;
  float* zpad = tab + geo.H * tabn;                                       // [ZPAD] zeros (HAS_TAB only)
  float* g2l_s = zpad + (HAS_TAB ? ZPAD : 0);                             // [H][16]  local query -> global key bias
  float* bg_s = g2l_s + geo.H * 16;                                       // [H][8]   global query -> local key bias
  float* g2g_s = bg_s + geo.H * 8;                                        // [H][8][16]
  const int nfl = geo.H * tabn + (HAS_TAB ? ZPAD : 0) + geo.H * (16 + 8 + 128);
  const int bars_off = (SM::OFF_TAB + nfl * 4 + 15) & ~15;
  uint64_t* bars_p = reinterpret_cast<uint64_t*>(smem + bars_off);
  const uint32_t bars = smem_u32(bars_p);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_p + BB::COUNT);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr float L2E = 1.4426950408889634f;

  // ---------------------------------------------------------------- one-time setup
  for (int i = tid; i < SM::OFF_TAB / 16; i += kThreads2) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if constexpr (HAS_TAB) {
    for (int i = tid; i < geo.H * tabn; i += kThreads2) {
      const int h = i / tabn, idx = i % tabn;
      const int dr = idx / TW - (2 * W - 1), dc = idx % TW - (2 * W - 1);
      float v = (a.table != nullptr) ? a.table[(long long)idx * geo.H + h] * L2E : 0.f;
      if (geo.exact == 1 && (abs(dr) > W || abs(dc) > W)) v = -INFINITY;
      tab[i] = v;
    }
    for (int i = tid; i < ZPAD; i += kThreads2) zpad[i] = 0.f;
  }
  for (int i = tid; i < geo.H * 16; i += kThreads2) {
    const int h = i / 16, t = i % 16;
    g2l_s[i] = (a.g2l != nullptr && t < geo.g) ? a.g2l[((long long)geo.H + h) * geo.g + t] * L2E : 0.f;
  }
  for (int i = tid; i < geo.H * 8; i += kThreads2) {
    const int h = i / 8, t = i % 8;
    bg_s[i] = (a.g2l != nullptr && t < geo.g) ? a.g2l[(long long)h * geo.g + t] * L2E : 0.f;
  }
  for (int i = tid; i < geo.H * 128; i += kThreads2) {
    const int h = i / 128, aa = (i % 128) / 16, bb = i % 16;
    g2g_s[i] = (a.g2g != nullptr && aa < geo.g && bb < geo.g) ? a.g2g[((long long)h * geo.g + aa) * geo.g + bb] * L2E : 0.f;
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(bar(BB::QFULL + i), 1); mbar_init(bar(BB::QEMPTY + i), 1); }
    for (int i = 0; i < NSTG; ++i) { mbar_init(bar(BB::KVFULL + i), 1); mbar_init(bar(BB::KVEMPTY + i), 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(bar(BB::SFULL + i), 1); mbar_init(bar(BB::PFULL + i), 128); mbar_init(bar(BB::PVDONE + i), 1); }
    mbar_init(bar(BB::OFREE), 128);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t TM_S = tmem, TM_O = tmem + 192;            // S ring: 3 x 64 columns; O: [192, 192 + DP)

  const int units_per_bh = geo.mx * a.cpairs;

  if (warp == 4) {
    // ================================================================= TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      Ring<NSTG> st; st.reset();
      uint32_t uc = 0;
      auto load_chunk = [&](int h, int b, int KR, int KC) {
        mbar_wait(bar(BB::KVEMPTY + st.i), st.ph ^ 1);
        unsigned char* dK = sKV + st.i * SM::STAGE_BYTES;
        mbar_arrive_expect_tx(bar(BB::KVFULL + st.i), 2 * W2 * ROWB);
        tma_load_5d(dK, &tmK, bar(BB::KVFULL + st.i), 0, KC * W, KR * W, h, b);
        tma_load_5d(dK + SM::KV_BYTES, &tmV, bar(BB::KVFULL + st.i), 0, KC * W, KR * W, h, b);
        st.adv();
      };
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh = unit / units_per_bh, rem = unit % units_per_bh;
        const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
        const uint32_t qb = uc & 1, qphase = (uc >> 1) & 1;
        if (uc >= 2) mbar_wait(bar(BB::QEMPTY + qb), qphase ^ 1);
        Sched sc; sc.init(geo, R, Cp);
        unsigned char* q0 = sQ + qb * SM::Q_BYTES;
        mbar_arrive_expect_tx(bar(BB::QFULL + qb), ((sc.hasB ? 2 : 1) * W2 + (a.fuse_g ? 8 : 0)) * ROWB);
        tma_load_5d(q0, &tmQ, bar(BB::QFULL + qb), 0, sc.C0 * W, R * W, h, b);
        if (sc.hasB) tma_load_5d(q0 + 64 * ROWB, &tmQ, bar(BB::QFULL + qb), 0, (sc.C0 + 1) * W, R * W, h, b);
        if (a.fuse_g) tma_load_4d(q0 + kGRow0 * ROWB, &tmQg, bar(BB::QFULL + qb), 0, 0, h, b);
        const int n = sc.count();
        for (int i = 0; i < n; ++i) {
          Iter it;
          if (!sc.get(geo, i, it)) continue;
          if (it.type == 1) {
            mbar_wait(bar(BB::KVEMPTY + st.i), st.ph ^ 1);
            unsigned char* dK = sKV + st.i * SM::STAGE_BYTES;
            mbar_arrive_expect_tx(bar(BB::KVFULL + st.i), 2 * 16 * ROWB);
            tma_load_4d(dK, &tmKg, bar(BB::KVFULL + st.i), 0, 0, h, b);
            tma_load_4d(dK + SM::KV_BYTES, &tmVg, bar(BB::KVFULL + st.i), 0, 0, h, b);
            st.adv();
          } else if (it.two) {
            load_chunk(h, b, it.KR, it.KCa);
            load_chunk(h, b, it.KR, it.KCb);
          } else {
            load_chunk(h, b, it.KR, it.hasA ? it.KCa : it.KCb);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ================================================================= MMA issuer (one elected thread)
    if (elect_one()) {
      constexpr uint32_t IDESC_S = make_idesc(128, 64, BF16, false, false);
      constexpr uint32_t IDESC_SG = make_idesc(128, 16, BF16, false, false);
      constexpr uint32_t IDESC_O = make_idesc_ab(128, DP, BF16 && !P16, BF16, false, true);
      constexpr int KS = DP / 16;
      Ring<NSTG> st_s, st_p;      // K/V stage of the next S issue / of the next PV issue
      Ring<3> sb_s, sb_p;         // S buffer of the next S issue / PV issue
      st_s.reset(); st_p.reset(); sb_s.reset(); sb_p.reset();
      uint32_t uc = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int rem = unit % units_per_bh;
        const int R = rem / a.cpairs, Cp = rem % a.cpairs;
        const uint32_t qb = uc & 1, qphase = (uc >> 1) & 1;
        Sched sc; sc.init(geo, R, Cp);
        const int n = sc.count();
        mbar_wait(bar(BB::QFULL + qb), qphase);
        const uint32_t qaddr = smem_u32(sQ + qb * SM::Q_BYTES);
        uint64_t qd[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) qd[k] = make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT);

        // S = Q K^T of one tile: K/V stage st_s -> S buffer sb_s
        auto issue_S = [&](bool glob) {
          const uint32_t kaddr = smem_u32(sKV + st_s.i * SM::STAGE_BYTES);
          uint64_t kd[KS];
#pragma unroll
          for (int k = 0; k < KS; ++k) kd[k] = make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT);
          mbar_wait(bar(BB::KVFULL + st_s.i), st_s.ph);
          tc_fence_after();
          const uint32_t d = TM_S + sb_s.i * 64;
#pragma unroll
          for (int k = 0; k < KS; ++k) mma_ss(d, qd[k], kd[k], glob ? IDESC_SG : IDESC_S, k > 0);
          mma_commit(bar(BB::SFULL + sb_s.i));
          st_s.adv(); sb_s.adv();
        };
        // O += P V of one tile: S buffer sb_p (P), K/V stage st_p (V)
        auto issue_PV = [&](bool glob, bool accumulate) {
          const uint32_t vaddr = smem_u32(sKV + st_p.i * SM::STAGE_BYTES + SM::KV_BYTES);
          uint64_t vd[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) vd[k] = make_smem_desc(vaddr + k * 16 * ROWB, 16, SBO, LAYOUT);
          mbar_wait(bar(BB::PFULL + sb_p.i), sb_p.ph);
          tc_fence_after();
          const uint32_t paddr = TM_S + sb_p.i * 64;
          if (glob) {
            mma_ts(TM_O, paddr, vd[0], IDESC_O, accumulate);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ts(TM_O, paddr + k * 8, vd[k], IDESC_O, accumulate || k > 0);
          }
          mma_commit(bar(BB::KVEMPTY + st_p.i));
          mma_commit(bar(BB::PVDONE + sb_p.i));
          st_p.adv(); sb_p.adv();
        };

        int si = 0, pi = 0, inflight = 0;      // next iteration index whose S / PV is to be issued; S buffers in flight
        Iter its, itp;
        bool s_ok = false, p_ok = false, first = true, q_released = false;
        auto next_s = [&]() { s_ok = false; while (si < n && !(s_ok = sc.get(geo, si, its))) ++si; };
        auto next_p = [&]() { p_ok = false; while (pi < n && !(p_ok = sc.get(geo, pi, itp))) ++pi; };
        next_s(); next_p();
        while (p_ok) {
          while (s_ok && inflight + (its.two ? 2 : 1) <= 3) {
            issue_S(its.type == 1);
            if (its.two) issue_S(false);
            inflight += its.two ? 2 : 1;
            ++si; next_s();
          }
          if (!s_ok && !q_released) { mma_commit(bar(BB::QEMPTY + qb)); q_released = true; }   // every S of this unit is issued
          if (first && uc > 0) mbar_wait(bar(BB::OFREE), (uc - 1) & 1);                       // previous unit's O has been read
          issue_PV(itp.type == 1, !first);
          if (itp.two) issue_PV(false, true);
          inflight -= itp.two ? 2 : 1;
          first = false;
          ++pi; next_p();
        }
      }
    }
  } else {
    // ================================================================= softmax warps (thread = TMEM lane)
    const int row = tid;                 // 0..127
    const int slot = row >> 6, l = row & 63;
    const int qr = l / W, qc = l % W;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool grow = a.fuse_g && slot == 0 && l >= kGRow0 && l < kGRow0 + geo.g;     // this lane is a global query row
    const int ga = l - kGRow0;
    const float c = a.scale_log2;
    Ring<3> sb; sb.reset();
    uint32_t uc = 0;
    for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
      const int bh = unit / units_per_bh, rem = unit % units_per_bh;
      const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
      Sched sc; sc.init(geo, R, Cp);
      const int C = sc.C0 + slot;
      const int r = R * W + qr, cc = C * W + qc;
      const bool row_ok = C < geo.my && l < W2 && r < geo.nx && cc < geo.ny;
      float m_use = -INFINITY, l_run = 0.f;
      const float* tab_h = tab + h * tabn;
      const float bias_g = grow ? bg_s[h * 8 + ga] : 0.f;
      bool first = true;
      uint32_t last_buf = 0, last_ph = 0;            // S buffer (and its phase) of the latest PV this unit has requested
      uint32_t vm, tm, pm;
      sc.masks(geo, slot, vm, tm, pm);
      const int gbit = geo.mode == 0 ? 9 : 2;
      const bool anypad = (geo.padx | geo.pady) != 0;
      for (uint32_t mleft = vm; mleft != 0; mleft &= mleft - 1) {
        const int i = __ffs(mleft) - 1;
        const bool two = (tm >> i) & 1u, part = (pm >> i) & 1u, is_glob = (i == gbit);
        const bool own = geo.mode == 0 ? (i < 2) : (i == 0);
        // my buffer / the other one (paired iterations)
        const uint32_t b0 = sb.i, p0 = sb.ph;
        sb.adv();
        uint32_t b1 = b0, p1 = p0;
        if (two) { b1 = sb.i; p1 = sb.ph; sb.adv(); }
        const uint32_t mb = (two && slot == 1) ? b1 : b0, mp = (two && slot == 1) ? p1 : p0;
        const uint32_t ob = (two && slot == 1) ? b0 : b1;
        mbar_wait(bar(BB::SFULL + mb), mp);
        if (two) mbar_wait(bar(BB::SFULL + ob), (slot == 1) ? p0 : p1);
        tc_fence_after();
        const uint32_t saddr = TM_S + mb * 64 + lane_base;
        // O must be stable before it is rescaled: the PV of the latest buffer this thread handed over has completed
        auto rescale_o = [&](float f) {
          mbar_wait(bar(BB::PVDONE + last_buf), last_ph);
          tc_fence_after();
#pragma unroll
          for (int q4 = 0; q4 < DP / 32; ++q4) {
            uint32_t ov[32];
            tmem_ld_x32(TM_O + lane_base + q4 * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * f);
            tmem_st_x32(TM_O + lane_base + q4 * 32, ov);
          }
        };
        // lazy online-softmax update (log2 domain): the reference maximum only moves when the new one beats it by > 2^8,
        // so P <= 2^8 (fp16-safe) and O is rescaled rarely; warp-uniform decision
        auto update_max = [&](float m_new) {
          const bool need = !first && (m_new > m_use + 8.f);
          if (first) m_use = m_new;
          if (__any_sync(0xffffffffu, need)) {
            const float f = need ? fast_exp2(m_use - m_new) : 1.f;     // m_use == -inf -> 0
            if (need) { m_use = m_new; l_run *= f; }
            rescale_o(f);
          }
        };
        if (!part) {
          uint32_t z[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = 0u;
          tmem_st_x32(saddr, z);
        } else if (is_glob) {
          // ---- global keys: 16 columns; local rows: bias g2l[1][h][t]; global rows (unit (0,0) only): g2g[h][a][t]
          uint32_t s[16];
          tmem_ld_x16(saddr, s);
          tmem_ld_wait();
          const bool gown = grow && R == 0 && Cp == 0;
          const float* brow = grow ? (g2g_s + h * 128 + ga * 16) : (g2l_s + h * 16);
          const float addg = (grow && !gown) ? -INFINITY : 0.f;
          float t[16], mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            t[j] = (j < geo.g) ? fmaf(__uint_as_float(s[j]), c, brow[j]) + addg : -INFINITY;
            mx = fmaxf(mx, t[j]);
          }
          update_max(fmaxf(m_use, mx));
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
          tmem_st_x8(saddr, p8);
        } else {
          // key-chunk coordinates are only needed off the hot path (padded geometries, table lookups)
          int KR = 0, KC = 0, krows = W, kcols = W;
          bool masked = false;
          if (HAS_TAB || anypad) {
            Iter it;
            sc.get(geo, i, it);
            KR = it.KR; KC = slot == 0 ? it.KCa : it.KCb;
            krows = min(W, geo.nx - KR * W); kcols = min(W, geo.ny - KC * W);
            masked = (krows < W) || (kcols < W);
          }
          // per-thread addend: 0 for local rows; global rows: their (constant) bias on the chunks this unit owns, else -inf
          const float radd = grow ? (own ? bias_g : -INFINITY) : 0.f;
          if (!HAS_TAB && !masked) {
            // ================= hot path (interior chunk, no table): all w*w raw scores of the row in registers after ONE
            // TMEM round trip (streaming S in 16-column steps was tried: 8 dependent load round trips per iteration made the
            // iteration 1.6x longer), maximum with 3-input FMNMX, then exp2 / row sum / pack over the same registers
            uint32_t s0[32], s1[32];
            tmem_ld_x32(saddr, s0);
            if constexpr (W2 > 48) tmem_ld_x32(saddr + 32, s1); else { uint32_t (&t16)[16] = *reinterpret_cast<uint32_t (*)[16]>(s1); tmem_ld_x16(saddr + 32, t16); }
            tmem_ld_wait();
            auto sc_ = [&](int j) -> float { return __uint_as_float(j < 32 ? s0[j] : s1[j - 32]); };
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < W2; j += 2) {
              if (j + 1 < W2) mx4[(j >> 1) & 3] = fmax3(mx4[(j >> 1) & 3], sc_(j), sc_(j + 1));
              else mx4[(j >> 1) & 3] = fmaxf(mx4[(j >> 1) & 3], sc_(j));
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            update_max(fmaxf(m_use, fmaf(mx, c, radd)));           // radd == -inf: the maximum does not move
            const float m_eff = (m_use == -INFINITY) ? 0.f : m_use;
            const float add = radd - m_eff;
            float sum[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 64; j += 2) {
              float p0v = 0.f, p1v = 0.f;
              if (j < W2) {
                const bool two_ = j + 1 < W2;
                float x0, x1;
                ffma2(x0, x1, sc_(j), two_ ? sc_(j + 1) : 0.f, c, c, add, add);
                if (POLY > 0 && ((j >> 1) % POLY) == POLY - 1) {
                  p0v = exp2_poly(x0);
                  p1v = two_ ? exp2_poly(x1) : 0.f;
                } else {
                  p0v = fast_exp2(x0);
                  p1v = two_ ? fast_exp2(x1) : 0.f;
                }
                const int k2 = j & 2;
                fadd2(sum[k2], sum[k2 + 1], sum[k2], sum[k2 + 1], p0v, p1v);
              }
              pk[j >> 1] = pack2<BF16 && !P16>(p0v, p1v);
            }
            l_run += (sum[0] + sum[1]) + (sum[2] + sum[3]);
            tmem_st_x32(saddr, pk);
          } else if constexpr (!LEAN) {
            // ================= general path (bias / window-mask table, or an edge chunk with padded keys): fully unrolled
            // over the w*w columns with compile-time (row, col) of each key; rare on the published configurations
            const int dR = KR - R, dC = KC - C;
            const float* tb = nullptr;
            if constexpr (HAS_TAB) {
              tb = grow ? (zpad + ZPAD - 1) : (tab_h + ((qr - dR * W + 2 * W - 1) * TW + (qc - dC * W + 2 * W - 1)));
            }
            uint32_t s0[32], s1[32];
            tmem_ld_x32(saddr, s0);
            tmem_ld_x32(saddr + 32, s1);
            tmem_ld_wait();
            // logit of column j (recomputed in the second pass instead of keeping w*w more registers alive)
            auto xj = [&](int j) -> float {
              float x = __uint_as_float(j < 32 ? s0[j] : s1[j - 32]) * c;
              if constexpr (HAS_TAB) x += tb[-((j / W) * TW + (j % W))];
              return ((j / W) < krows && (j % W) < kcols) ? x : -INFINITY;
            };
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < W2; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], xj(j));
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            update_max(fmaxf(m_use, mx + radd));
            const float m_eff = (m_use == -INFINITY) ? 0.f : m_use;
            const float add = radd - m_eff;
            float sum = 0.f;
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 64; j += 2) {
              float p0v = 0.f, p1v = 0.f;
              if (j < W2) {
                p0v = fast_exp2(xj(j) + add);
                p1v = (j + 1 < W2) ? fast_exp2(xj(j + 1) + add) : 0.f;
                sum += p0v + p1v;
              }
              pk[j >> 1] = pack2<BF16 && !P16>(p0v, p1v);
            }
            l_run += sum;
            tmem_st_x32(saddr, pk);
          }
        }
        if (two) {
          // my rows of the other slot's tile must contribute nothing to O
          uint32_t z[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = 0u;
          tmem_st_x32(TM_S + ob * 64 + lane_base, z);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar(BB::PFULL + b0));
        if (two) mbar_arrive(bar(BB::PFULL + b1));
        last_buf = two ? b1 : b0; last_ph = two ? p1 : p0;
        first = false;
      }
      // ---- epilogue: O / l -> global, LSE; global rows -> partial (m, l, O)
      mbar_wait(bar(BB::PVDONE + last_buf), last_ph);
      tc_fence_after();
      constexpr int OC = DP / 32;
      uint32_t ov[OC][32];
#pragma unroll
      for (int q4 = 0; q4 < OC; ++q4) tmem_ld_x32(TM_O + lane_base + q4 * 32, ov[q4]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar(BB::OFREE));
      if (row_ok) {
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        const long long tok = (long long)r * geo.ny + cc;
        store_row<OC, BF16>(a.o, a.out_f32, b, h, tok, geo.D, ov, inv);
        a.lse[((long long)b * geo.H + h) * geo.Nloc + tok] = (m_use + log2f(l_run)) * 0.6931471805599453f;
      } else if (grow) {
        float* dst = a.part + (((long long)bh * units_per_bh + rem) * kGMax + ga) * (DP + 2);
        dst[0] = m_use; dst[1] = l_run;
#pragma unroll
        for (int q4 = 0; q4 < OC; ++q4)
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[2 + q4 * 32 + j] = __uint_as_float(ov[q4][j]);
      }
    }
  }
  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// Combine the per-unit partials of the global query rows: og = sum_u O_u 2^(m_u - M) / sum_u l_u 2^(m_u - M),
// lse_g = (M + log2 L) ln 2.  One warp per (b, h, a); lane = output channel (two channels per lane for D = 64).
template <typename TO>
__global__ void vil_tc_fwd2_merge(Geo geo, const float* __restrict__ part, int units_per_bh, int DP, T4 og, float* __restrict__ lse_g) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= geo.B * geo.H * geo.g) return;
  const int a = wid % geo.g, bh = wid / geo.g, b = bh / geo.H, h = bh % geo.H;
  const int stride = DP + 2;
  const float* base = part + ((long long)bh * units_per_bh * kGMax + a) * stride;
  float M = -INFINITY;
  for (int u = lane; u < units_per_bh; u += 32) M = fmaxf(M, base[(long long)u * kGMax * stride]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  float L = 0.f, acc0 = 0.f, acc1 = 0.f;
  for (int u = 0; u < units_per_bh; ++u) {
    const float* pu = base + (long long)u * kGMax * stride;
    const float m = pu[0];
    const float sc = (m == -INFINITY) ? 0.f : exp2f(m - M);
    L = fmaf(pu[1], sc, L);
    acc0 = fmaf(pu[2 + lane], sc, acc0);
    if (DP == 64) acc1 = fmaf(pu[2 + 32 + lane], sc, acc1);
  }
  const float inv = L > 0.f ? 1.f / L : 0.f;
  TO* dst = row_ptr_w<TO>(og, b, h, a);
  if (lane < geo.D) dst[lane] = ElemTraits<TO>::from_f(acc0 * inv);
  if (DP == 64 && 32 + lane < geo.D) dst[32 + lane] = ElemTraits<TO>::from_f(acc1 * inv);
  if (lane == 0) lse_g[(long long)bh * geo.g + a] = (M + log2f(L)) * 0.6931471805599453f;
}

}  // namespace f2
}  // namespace tc
}  // namespace vil
