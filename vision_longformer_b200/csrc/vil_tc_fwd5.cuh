// Fused tcgen05 / TMA forward, KEY-ROW-BLOCK variant (sm_100a): w = 7, D <= 32, mode 0, no table, no padded chunk - the
// configuration of every published ViL stage-1 layer (longformer2d.py:126-202 + 210-227).  Same skeleton as vil_tc_fwd3
// (four CTAs per SM, ONE S buffer with P packed over it, optimistic bf16 softmax) - what changes is the key BLOCK.
//
// Why (ncu on vil_tc_fwd3, profiles/r02_final_ncu_S1_source_hotspots.txt): a softmax warp issues ~4000 instructions per unit
// for 457 exponentials.  The unit walks 13 blocks (global keys + the 3 x 4 chunk window); every block costs ~160 instructions
// of control (mbarrier waits, fences, block walk, TMEM waits, hand-over) next to ~150 of arithmetic, and 3 of the 12 chunk
// blocks are visited by one slot only (the other writes 32 zero words and waits).
//
// Here the window of a unit - 21 key rows x 28 key columns, the chunk rows R-1..R+1 and chunk columns C0-1..C0+2 - is cut
// ACROSS the chunks: a block is THREE KEY ROWS of all four chunk columns,
//     score column  j = cg * 24 + kr * 8 + kc      (cg: chunk column group 0..3, kr: key row 0..2, kc: 0..7)
// loaded by four TMA boxes of (D, 8 columns, 3 rows) per operand; column kc = 7 of a group is the first column of the next
// chunk (a duplicate - switched off at compile time), so N = 96 with 84 real keys.  Slot A (query chunk C0) reads the groups
// 0..2, slot B (chunk C0+1) the groups 1..3: the SAME code with a base offset of 24 columns / 12 packed words, 63
// exponentials per thread and block, 12 zero words for the fourth group.  7 blocks (+ the global keys) per unit instead of
// 12 (+1), every block useful for both slots; a key row outside the image (top / bottom edge) is switched off by its logit
// addend.  TMEM: S 96 columns (P packed over the consumed S columns) + O 32 = 128 -> four CTAs per SM as in fwd3.
// One group (24 columns, 21 exponentials) is ONE loop body: the hot loop is ~110 instructions (the L0 instruction cache is
// ~6 KB; the unrolled strip variant of this round measured 23-33 % `no_instruction` stalls and was dropped).
#pragma once
#include "vil_tc_fwd3.cuh"

namespace vil {
namespace tc {
namespace f5 {

using namespace sm100;
using f2::Args;
using f2::kGRow0;
using f2::kGMax;

constexpr int kThreads5 = 192;          // warps 0-3 softmax, 4 TMA producer, 5 MMA issuer
constexpr int kW = 7, kW2 = 49;
constexpr int kGrp = 24;                // score columns per chunk-column group (3 key rows x 8)
constexpr int kBlk = 4 * kGrp;          // 96
constexpr int kRowB = 64;               // D padded to 32 channels, 2 bytes

struct Smem {
  static constexpr int NQ = 2, NSTG = 3;
  static constexpr int Q_BYTES = 128 * kRowB;                 // 8192
  static constexpr int KV_BYTES = kBlk * kRowB;               // 6144, one of K / V
  static constexpr int STAGE_BYTES = 2 * KV_BYTES;            // 12288
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = NQ * Q_BYTES;
  static constexpr int OFF_TAB = OFF_KV + NSTG * STAGE_BYTES;
  static __host__ __device__ int total(int tab_floats) { return OFF_TAB + tab_floats * 4 + 256 + 1024; }
};

struct Bars {
  enum { QFULL = 0, QEMPTY = 2, KVFULL = 4, KVEMPTY = 7, SFULL = 10, PFULL = 11, PVDONE = 12, OFREE = 13, COUNT = 14 };
};

__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Blocks of the persistent loop, in order: per unit [global-key tile], then the key-row blocks kb = 0..6 (key rows
// (R-1) 7 + 3 kb + {0,1,2}) that touch the image.
struct Walk {
  int unit, uc, step, nunits;
  uint32_t m;                 // remaining blocks of the current unit: bit 0 global tile, bit 1 + kb key-row block kb
  int b, h, R, Cp, rem, bh;
  int type, kb;               // current block: type 1 = global tile
  bool first, last, fresh;
  __device__ __forceinline__ void load(const Geo& g, int cpairs, int units_per_bh) {
    bh = unit / units_per_bh; rem = unit - bh * units_per_bh;
    b = bh / g.H; h = bh - b * g.H; R = rem / cpairs; Cp = rem - R * cpairs;
    m = g.g > 0 ? 1u : 0u;
    const int r0 = (R - 1) * kW;
#pragma unroll
    for (int k = 0; k < 7; ++k)
      if (r0 + 3 * k + 2 >= 0 && r0 + 3 * k < g.nx) m |= 2u << k;
  }
  __device__ __forceinline__ void init(const Geo& g, int cpairs, int units_per_bh, int unit0, int stride, int total) {
    unit = unit0; step = stride; nunits = total; uc = 0; m = 0; fresh = true; first = false; last = true;
    if (unit < nunits) load(g, cpairs, units_per_bh);
  }
  __device__ __forceinline__ bool next(const Geo& g, int cpairs, int units_per_bh) {
    if (unit >= nunits) return false;
    first = fresh;
    fresh = false;
    if (m == 0) {
      unit += step; ++uc;
      if (unit >= nunits) return false;
      load(g, cpairs, units_per_bh);
      first = true;
    }
    const int bit = __ffs(m) - 1;
    m &= m - 1;
    type = bit == 0 ? 1 : 0;
    kb = bit - 1;
    last = (m == 0);
    return true;
  }
};

// One chunk-column group (24 score columns at TMEM address scol) of one query row: three key rows of 8 columns, the eighth
// switched off.
//   DOEXP = false: only the raw maximum of every key row is tracked in mx (first block of a row / fp16 two-pass mode).
//   DOEXP = true : p = 2^(s c + add[kr]) accumulated into sum and packed over the consumed S columns at pcol (12 words).
//                  NO maximum here: the optimistic pass detects a logit far above its reference from the block's SUM (any
//                  p > 2^T makes the sum > 2^T), which removes the FMNMX per pair and the per-row maximum bookkeeping from
//                  the hot loop (the kernel is bound by instruction issue as much as by the XU pipe).
template <bool BF16, bool DOEXP>
__device__ __forceinline__ void group24(uint32_t scol, uint32_t pcol, float c, const float (&add)[3], float (&mx)[3], float (&sum)[2]) {
  uint32_t s[16], t[8], p[12];
  tmem_ld_x16(scol, s);
  tmem_ld_x8(scol + 16, t);
  tmem_ld_wait();
#pragma unroll
  for (int kr = 0; kr < 3; ++kr) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(kr < 2 ? s[kr * 8 + j] : t[j]);
    if constexpr (!DOEXP) {
      mx[kr] = f2::fmax3(mx[kr], v[0], v[1]);
      mx[kr] = f2::fmax3(mx[kr], v[2], v[3]);
      mx[kr] = f2::fmax3(mx[kr], v[4], v[5]);
      mx[kr] = fmaxf(mx[kr], v[6]);
    } else {
      float x[8];
      ffma2(x[0], x[1], v[0], v[1], c, c, add[kr], add[kr]);
      ffma2(x[2], x[3], v[2], v[3], c, c, add[kr], add[kr]);
      ffma2(x[4], x[5], v[4], v[5], c, c, add[kr], add[kr]);
      x[6] = fmaf(v[6], c, add[kr]);
#pragma unroll
      for (int j = 0; j < 7; ++j) x[j] = fast_exp2(x[j]);
      fadd2(sum[0], sum[1], sum[0], sum[1], x[0], x[1]);
      fadd2(sum[0], sum[1], sum[0], sum[1], x[2], x[3]);
      fadd2(sum[0], sum[1], sum[0], sum[1], x[4], x[5]);
      sum[0] += x[6];
      p[kr * 4 + 0] = pack2<BF16>(x[0], x[1]);
      p[kr * 4 + 1] = pack2<BF16>(x[2], x[3]);
      p[kr * 4 + 2] = pack2<BF16>(x[4], x[5]);
      p[kr * 4 + 3] = pack2<BF16>(x[6], 0.f);
    }
  }
  if constexpr (DOEXP) {
    uint32_t p8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) p8[j] = p[j];
    tmem_st_x8(pcol, p8);
    tmem_st_x4(pcol + 8, p[8], p[9], p[10], p[11]);
  }
}

__device__ __forceinline__ void zero_group(uint32_t pcol) {
  uint32_t z8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) z8[j] = 0u;
  tmem_st_x8(pcol, z8);
  tmem_st_x4(pcol + 8, 0u, 0u, 0u, 0u);
}

// EXACT = false (bf16): optimistic single pass + deferred rescale;  true (fp16): two passes per block, P <= 2^8.
template <bool BF16, bool EXACT>
__global__ void __launch_bounds__(kThreads5, 4)
vil_tc_fwd5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQg,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmKg, const __grid_constant__ CUtensorMap tmVg, const Args a) {
  using SM = Smem;
  using BB = Bars;
  constexpr int DP = 32;
  constexpr uint32_t LAYOUT = SWZ_64B;
  constexpr uint32_t SBO = 8 * kRowB;
  constexpr uint32_t TMEM_COLS = 128;
  const Geo& geo = a.geo;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sQ = smem + SM::OFF_Q;
  unsigned char* sKV = smem + SM::OFF_KV;
  float* g2l_s = reinterpret_cast<float*>(smem + SM::OFF_TAB);            // [H][16]
  float* bg_s = g2l_s + geo.H * 16;                                       // [H][8]
  float* g2g_s = bg_s + geo.H * 8;                                        // [H][8][16]
  const int nfl = geo.H * (16 + 8 + 128);
  const int bars_off = (SM::OFF_TAB + nfl * 4 + 15) & ~15;
  uint64_t* bars_p = reinterpret_cast<uint64_t*>(smem + bars_off);
  const uint32_t bars = smem_u32(bars_p);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_p + BB::COUNT);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr float L2E = 1.4426950408889634f;

  for (int i = tid; i < SM::OFF_TAB / 16; i += kThreads5) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < geo.H * 16; i += kThreads5) {
    const int h = i / 16, t = i % 16;
    g2l_s[i] = (a.g2l != nullptr && t < geo.g) ? a.g2l[((long long)geo.H + h) * geo.g + t] * L2E : 0.f;
  }
  for (int i = tid; i < geo.H * 8; i += kThreads5) {
    const int h = i / 8, t = i % 8;
    bg_s[i] = (a.g2l != nullptr && t < geo.g) ? a.g2l[(long long)h * geo.g + t] * L2E : 0.f;
  }
  for (int i = tid; i < geo.H * 128; i += kThreads5) {
    const int h = i / 128, aa = (i % 128) / 16, bb = i % 16;
    g2g_s[i] = (a.g2g != nullptr && aa < geo.g && bb < geo.g) ? a.g2g[((long long)h * geo.g + aa) * geo.g + bb] * L2E : 0.f;
  }
  if (tid == 0) {
    for (int i = 0; i < SM::NQ; ++i) { mbar_init(bar(BB::QFULL + i), 1); mbar_init(bar(BB::QEMPTY + i), 1); }
    for (int i = 0; i < SM::NSTG; ++i) { mbar_init(bar(BB::KVFULL + i), 1); mbar_init(bar(BB::KVEMPTY + i), 1); }
    mbar_init(bar(BB::SFULL), 1); mbar_init(bar(BB::PFULL), 128); mbar_init(bar(BB::PVDONE), 1); mbar_init(bar(BB::OFREE), 128);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t TM_S = tmem, TM_O = tmem + kBlk;            // P is packed over the consumed S columns

  const int units_per_bh = geo.mx * a.cpairs;

  if (warp == 4) {
    // ================================================================= TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      uint32_t stage = 0, kv_phase = 0;
      Walk wk; wk.init(geo, a.cpairs, units_per_bh, blockIdx.x, gridDim.x, a.num_units);
      while (wk.next(geo, a.cpairs, units_per_bh)) {
        if (wk.first) {
          const uint32_t qb = wk.uc & 1, qphase = (wk.uc >> 1) & 1;
          if (wk.uc >= 2) mbar_wait(bar(BB::QEMPTY + qb), qphase ^ 1);
          const bool hasB = 2 * wk.Cp + 1 < geo.my;
          unsigned char* q0 = sQ + qb * SM::Q_BYTES;
          mbar_arrive_expect_tx(bar(BB::QFULL + qb), ((hasB ? 2 : 1) * kW2 + (a.fuse_g ? 8 : 0)) * kRowB);
          tma_load_5d(q0, &tmQ, bar(BB::QFULL + qb), 0, (2 * wk.Cp) * kW, wk.R * kW, wk.h, wk.b);
          if (hasB) tma_load_5d(q0 + 64 * kRowB, &tmQ, bar(BB::QFULL + qb), 0, (2 * wk.Cp + 1) * kW, wk.R * kW, wk.h, wk.b);
          if (a.fuse_g) tma_load_4d(q0 + kGRow0 * kRowB, &tmQg, bar(BB::QFULL + qb), 0, 0, wk.h, wk.b);
        }
        mbar_wait(bar(BB::KVEMPTY + stage), kv_phase ^ 1);
        unsigned char* dK = sKV + stage * SM::STAGE_BYTES;
        unsigned char* dV = dK + SM::KV_BYTES;
        if (wk.type == 1) {
          mbar_arrive_expect_tx(bar(BB::KVFULL + stage), 2 * 16 * kRowB);
          tma_load_4d(dK, &tmKg, bar(BB::KVFULL + stage), 0, 0, wk.h, wk.b);
          tma_load_4d(dV, &tmVg, bar(BB::KVFULL + stage), 0, 0, wk.h, wk.b);
        } else {
          mbar_arrive_expect_tx(bar(BB::KVFULL + stage), 2 * kBlk * kRowB);
          const int kr0 = (wk.R - 1) * kW + 3 * wk.kb;
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) {            // rows / columns outside the image are zero-filled by the TMA unit
            const int kc0 = (2 * wk.Cp - 1 + cg) * kW;
            tma_load_5d(dK + cg * kGrp * kRowB, &tmK, bar(BB::KVFULL + stage), 0, kc0, kr0, wk.h, wk.b);
            tma_load_5d(dV + cg * kGrp * kRowB, &tmV, bar(BB::KVFULL + stage), 0, kc0, kr0, wk.h, wk.b);
          }
        }
        if (++stage == SM::NSTG) { stage = 0; kv_phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ================================================================= MMA issuer (one elected thread)
    if (elect_one()) {
      constexpr uint32_t IDESC_S = make_idesc(128, kBlk, BF16, false, false);
      constexpr uint32_t IDESC_SG = make_idesc(128, 16, BF16, false, false);
      constexpr uint32_t IDESC_O = make_idesc(128, DP, BF16, false, true);
      Walk ws, wp;                                     // S side runs one block ahead of the PV side
      ws.init(geo, a.cpairs, units_per_bh, blockIdx.x, gridDim.x, a.num_units);
      wp.init(geo, a.cpairs, units_per_bh, blockIdx.x, gridDim.x, a.num_units);
      uint32_t s_stage = 0, s_phase = 0, p_stage = 0, G = 0;
      auto issue_S = [&]() {
        const uint32_t qb = ws.uc & 1;
        if (ws.first) mbar_wait(bar(BB::QFULL + qb), (ws.uc >> 1) & 1);
        mbar_wait(bar(BB::KVFULL + s_stage), s_phase);
        tc_fence_after();
        const uint32_t qaddr = smem_u32(sQ + qb * SM::Q_BYTES);
        const uint32_t kaddr = smem_u32(sKV + s_stage * SM::STAGE_BYTES);
        const uint32_t idesc = ws.type == 1 ? IDESC_SG : IDESC_S;
#pragma unroll
        for (int k = 0; k < 2; ++k)
          mma_ss(TM_S, make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT), make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT), idesc, k > 0);
        mma_commit(bar(BB::SFULL));
        if (++s_stage == SM::NSTG) { s_stage = 0; s_phase ^= 1; }
      };
      if (ws.next(geo, a.cpairs, units_per_bh)) issue_S();
      while (wp.next(geo, a.cpairs, units_per_bh)) {
        const bool haveN = ws.next(geo, a.cpairs, units_per_bh);
        const bool first = wp.first;
        const uint32_t qb = wp.uc & 1;
        const uint32_t vaddr = smem_u32(sKV + p_stage * SM::STAGE_BYTES + SM::KV_BYTES);
        mbar_wait(bar(BB::PFULL), G & 1);
        if (first && wp.uc > 0) mbar_wait(bar(BB::OFREE), (wp.uc - 1) & 1);
        tc_fence_after();
        if (wp.type == 1) {
          mma_ts(TM_O, TM_S, make_smem_desc(vaddr, 16, SBO, LAYOUT), IDESC_O, !first);
        } else {
#pragma unroll
          for (int k = 0; k < 6; ++k)
            mma_ts(TM_O, TM_S + k * 8, make_smem_desc(vaddr + k * 16 * kRowB, 16, SBO, LAYOUT), IDESC_O, (!first) || k > 0);
        }
        mma_commit(bar(BB::KVEMPTY + p_stage));
        mma_commit(bar(BB::PVDONE));
        if (wp.last) mma_commit(bar(BB::QEMPTY + qb));
        if (++p_stage == SM::NSTG) p_stage = 0;
        ++G;
        if (haveN) issue_S();          // P over S: S_{j+1} executes after PV_j on the in-order tensor pipe
      }
    }
  } else {
    // ================================================================= softmax warps (thread = TMEM lane)
    const int row = tid;
    const int slot = row >> 6, l = row & 63;
    const int qr = l / kW, qc = l % kW;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool grow = a.fuse_g && slot == 0 && l >= kGRow0 && l < kGRow0 + geo.g;
    const int ga = l - kGRow0;
    const float c = a.scale_log2;
    const uint32_t saddr = TM_S + lane_base;
    Walk wk; wk.init(geo, a.cpairs, units_per_bh, blockIdx.x, gridDim.x, a.num_units);
    uint32_t G = 0;
    float m_use = -INFINITY, l_run = 0.f, bias_g = 0.f;
    uint32_t gvalid = 0;                               // bit cg: chunk-column group cg exists and is visited by this slot
    // rescale O (and the running sum) by 2^(m_use - m_new); O is stable once the PV of block `done` has completed
    auto rescale = [&](bool need, float m_new, uint32_t done) {
      mbar_wait(bar(BB::PVDONE), done & 1);
      tc_fence_after();
      const float f = need ? fast_exp2(m_use - m_new) : 1.f;             // m_use == -inf -> 0
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
    while (wk.next(geo, a.cpairs, units_per_bh)) {
      const bool first = wk.first;
      const int R = wk.R, Cp = wk.Cp, h = wk.h, b = wk.b;
      if (first) {
        m_use = -INFINITY; l_run = 0.f;
        bias_g = grow ? bg_s[h * 8 + ga] : 0.f;
        gvalid = 0;
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) {
          const int kc = 2 * Cp - 1 + cg;
          if (kc >= 0 && kc < geo.my && cg - slot >= 0 && cg - slot <= 2) gvalid |= 1u << cg;
        }
        if (2 * Cp + slot >= geo.my) gvalid = 0;         // odd chunk-column count: slot B of the last pair is empty
      }
      mbar_wait(bar(BB::SFULL), G & 1);
      tc_fence_after();
      bool pend = false;
      float pend_m = -INFINITY;
      if (wk.type == 1) {
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
        m_use = mx;                                     // always the first block of its unit: exact start
        const float me = (m_use == -INFINITY) ? 0.f : m_use;
        float sg = 0.f;
        uint32_t p8[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float p0v = fast_exp2(t[j] - me), p1v = fast_exp2(t[j + 1] - me);
          sg += p0v + p1v;
          p8[j >> 1] = pack2<BF16>(p0v, p1v);
        }
        l_run += sg;
        tmem_st_x8(saddr, p8);
      } else {
        // ---- key rows kr0 .. kr0+2 of the four chunk-column groups
        const int kr_rel = 3 * wk.kb;                    // first key row of the block relative to chunk row R-1
        const int kr0 = (R - 1) * kW + kr_rel;
        // logit addend per key row: -inf outside the image; global rows only count the chunks their unit owns (chunk row R)
        float rv[3];
#pragma unroll
        for (int kr = 0; kr < 3; ++kr) {
          const bool inside = kr0 + kr >= 0 && kr0 + kr < geo.nx;
          const bool own = kr_rel + kr >= kW && kr_rel + kr < 2 * kW;
          rv[kr] = grow ? (own ? bias_g : -INFINITY) : (inside ? 0.f : -INFINITY);
        }
        // the addend of group cg: global rows see the groups 1, 2 only
        auto radd = [&](int cg, int kr) -> float { return (grow && cg != 1 && cg != 2) ? -INFINITY : rv[kr]; };
        float sum[2] = {0.f, 0.f};
        float bmax = -INFINITY;
        bool has_live = !grow;                           // a global row is switched off on the blocks it does not own
        if (grow) has_live = rv[0] > -INFINITY || rv[1] > -INFINITY || rv[2] > -INFINITY;
        const bool init = (m_use == -INFINITY) && has_live;
        const bool two_pass = EXACT || __any_sync(0xffffffffu, init);
        if (two_pass) {
          // pass 1: maximum of the block; rescale BEFORE the exponentials so that P <= 2^8 (fp16-safe) / the start is exact
#pragma unroll 1
          for (int cg = 0; cg < 4; ++cg) {
            if ((gvalid >> cg) & 1u) {
              float mx[3] = {-INFINITY, -INFINITY, -INFINITY}, dsum[2] = {0.f, 0.f};
              const float zero3[3] = {0.f, 0.f, 0.f};
              group24<BF16, false>(saddr + cg * kGrp, 0u, c, zero3, mx, dsum);
#pragma unroll
              for (int kr = 0; kr < 3; ++kr) bmax = fmaxf(bmax, mx[kr] * c + radd(cg, kr));
            }
          }
          const float m_new = fmaxf(m_use, bmax);
          const bool need = !first && (m_new > m_use + 8.f);
          if (first) m_use = m_new;                                      // O has not been written yet in this unit
          if (__any_sync(0xffffffffu, need)) rescale(need, m_new, G - 1);
          bmax = -INFINITY;
        }
        const float m_eff = (m_use == -INFINITY) ? 0.f : m_use;
        float addl[3];                                   // addends of a local row / of a global row on the groups 1, 2
#pragma unroll
        for (int kr = 0; kr < 3; ++kr) addl[kr] = rv[kr] - m_eff;
#pragma unroll 1
        for (int cg = 0; cg < 4; ++cg) {                 // ONE copy of the group code: the hot loop stays in the L0 I-cache
          const uint32_t pcol = saddr + cg * (kGrp / 2);
          if ((gvalid >> cg) & 1u) {
            float mxd[3] = {0.f, 0.f, 0.f};
            const bool off = grow && cg != 1 && cg != 2;
            const float add[3] = {off ? -INFINITY : addl[0], off ? -INFINITY : addl[1], off ? -INFINITY : addl[2]};
            group24<BF16, true>(saddr + cg * kGrp, pcol, c, add, mxd, sum);
          } else {
            zero_group(pcol);
          }
        }
        const float sblk = sum[0] + sum[1];
        l_run += sblk;
        // a logit more than ~2^16 above the reference shows in the block's sum: bring the reference up (deferred, below)
        if (!two_pass && sblk > 65536.f) { pend = true; pend_m = m_eff + log2f(sblk); }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar(BB::PFULL));
      if constexpr (!EXACT) {
        // deferred rescale: this block's P was produced against the old maximum and stays valid (bf16 has the range); O is
        // brought to the new maximum once the PV of THIS block has completed, before the next block's P is handed over
        if (__any_sync(0xffffffffu, pend)) {
          rescale(pend, pend_m, G);
          tc_fence_before();
        }
      }
      ++G;
      if (wk.last) {
        // ---- epilogue of the unit
        const int C = 2 * Cp + slot;
        const int rr = R * kW + qr, cc = C * kW + qc;
        const bool row_ok = C < geo.my && l < kW2;
        mbar_wait(bar(BB::PVDONE), (G - 1) & 1);
        tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        const long long tok = (long long)rr * geo.ny + cc;
        float* part = grow ? a.part + (((long long)wk.bh * units_per_bh + wk.rem) * kGMax + ga) * (DP + 2) : nullptr;
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace f5
}  // namespace tc
}  // namespace vil
