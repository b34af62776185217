// tcgen05 / TMA forward kernel of the Vision-Longformer attention (sm_100a), chunk size w <= 8.
//
// Work unit  = one (b, h, chunk-row R, pair of chunk columns {2Cp, 2Cp+1}): a 128-row query tile made of two
//              64-row "slots" (slot A = chunk (R,2Cp), slot B = chunk (R,2Cp+1); w*w <= 64 real rows each).
// Key blocks = the <= 16 global keys (one 16-column block) followed by every key chunk either slot visits
//              (<= 3 x 4 chunks), one chunk (w*w <= 64 keys) per block.  K/V chunks are staged by TMA as
//              [w*w rows x D] K-major tiles (5-D tensor map over (D, col, row, H, B) of the strided kv buffer;
//              out-of-image rows/cols are zero-filled by the TMA unit = the reference's F.pad,
//              longformer2d.py:138-144).
// Per block  : S = Q K^T (tcgen05.mma SS, M=128, N=64, fp32 accumulators in TMEM) -> 128 softmax threads
//              (thread = query row = TMEM lane) add bias / mask, online softmax in the log2 domain with lazy
//              rescaling, write bf16 P back into TMEM over S -> O += P V (tcgen05.mma TS, V tile MN-major).
// Roles      : warps 0-3 softmax + epilogue, warp 4 TMA producer, warp 5 MMA issuer; mbarrier pipelines;
//              persistent CTAs, 2 per SM (256 TMEM columns each).
#pragma once
#include "vil_common.cuh"
#include "vil_sm100.cuh"

namespace vil {
namespace tc {

using namespace sm100;

constexpr int kStages = 4;        // K/V ring depth
constexpr int kThreads = 192;

struct FwdArgs {
  Geo geo;
  T4 o;
  float* lse;
  const float* table;             // ((4w-1)^2, H) fp32 or null
  const float* g2l;               // (2,H,g) fp32 or null
  int cpairs;                     // ceil(my / 2)
  int num_units;                  // B*H*mx*cpairs
  int has_tab;                    // bias table needed in smem (rpe on, or exact == 1)
  float scale_log2;               // scale * log2(e)
  int out_f32;                    // parity build: o is an fp32 tensor (VIL_FLAG_F32_OUT)
};

__device__ __forceinline__ bool offset_used(const Geo& g, int dR, int dC) {
  if (g.mode == 0) return dR >= -1 && dR <= 1 && dC >= -1 && dC <= 1;
  if (dR == 0 && dC == 0) return true;
  return g.mode > 0 && dR == g.offR[1] && dC == g.offC[1];
}

// Deterministic enumeration of the blocks of one unit; every warp role walks the same sequence.
// The 3 x 4 chunk window around the unit (rows R-1..R+1, cols C0-1..C0+2) is described by two 12-bit masks
// (bit 4*r+c set = slot A / slot B visits that chunk), computed once per unit; iterating is a find-first-set.
// `mirror` = false: key chunks visited by the two QUERY slots (forward, dQ pass);
// `mirror` = true : query chunks that visit the two KEY slots (dK/dV pass) - the offset list negated.
struct BlockWalk {
  uint32_t m, maskA, maskB;
  int kr_base, kc_base, bit;
  bool global_pending;
  __device__ __forceinline__ void init(const Geo& g, int R, int Cp, bool mirror = false, bool with_global = true) {
    const int C0 = 2 * Cp;
    kr_base = R - 1; kc_base = C0 - 1;
    const bool hasB = C0 + 1 < g.my;
    maskA = 0; maskB = 0;
    if (g.mode == 0) {
      const uint32_t ca = (C0 > 0 ? 1u : 0u) | 2u | (C0 + 1 < g.my ? 4u : 0u);            // cols C0-1, C0, C0+1
      const uint32_t cb = hasB ? (2u | 4u | (C0 + 2 < g.my ? 8u : 0u)) : 0u;               // cols C0, C0+1, C0+2
      if (R > 0) { maskA |= ca; maskB |= cb; }
      maskA |= ca << 4; maskB |= cb << 4;
      if (R + 1 < g.mx) { maskA |= ca << 8; maskB |= cb << 8; }
    } else {
      maskA = 1u << 5;                              // own chunk (R, C0)
      if (hasB) maskB = 1u << 6;                    // own chunk (R, C0+1)
      if (g.mode > 0) {
        const int dR = mirror ? -g.offR[1] : g.offR[1], dC = mirror ? -g.offC[1] : g.offC[1];
        const int rr = R + dR;
        if (rr >= 0 && rr < g.mx) {
          const int cA = C0 + dC, cB = C0 + 1 + dC;
          if (cA >= 0 && cA < g.my) maskA |= 1u << (4 * (dR + 1) + (dC + 1));
          if (hasB && cB >= 0 && cB < g.my) maskB |= 1u << (4 * (dR + 1) + (dC + 2));
        }
      }
    }
    m = maskA | maskB;
    bit = 0;
    global_pending = with_global && g.g > 0;
  }
  // type: 1 = global block, 0 = local chunk (KR, KC); returns false when exhausted
  __device__ __forceinline__ bool next(const Geo&, int& type, int& KR, int& KC) {
    if (global_pending) { global_pending = false; type = 1; KR = KC = 0; return true; }
    if (m == 0) return false;
    bit = __ffs(m) - 1;
    m &= m - 1;
    type = 0; KR = kr_base + (bit >> 2); KC = kc_base + (bit & 3);
    return true;
  }
  // does slot (0 = A, 1 = B) visit the block returned by the last next()?
  __device__ __forceinline__ bool used_by(int slot) const { return ((slot ? maskB : maskA) >> bit) & 1u; }
};

template <int DP>
struct FwdSmem {
  static constexpr int ROWB = DP * 2;
  static constexpr int Q_BYTES = 128 * ROWB;
  static constexpr int KV_BYTES = 64 * ROWB;           // one of K or V
  static constexpr int STAGE_BYTES = 2 * KV_BYTES;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = 2 * Q_BYTES;
  static constexpr int OFF_TAB = OFF_KV + kStages * STAGE_BYTES;
  static __host__ __device__ int total(int tab_floats) { return OFF_TAB + tab_floats * 4 + 512 + 1024; }
};

// barrier indices
enum { BAR_QFULL = 0, BAR_QEMPTY = 2, BAR_KVFULL = 4, BAR_KVEMPTY = 4 + kStages, BAR_SFULL = 4 + 2 * kStages,
       BAR_PFULL = BAR_SFULL + 2, BAR_PVDONE = BAR_PFULL + 2, BAR_OFREE = BAR_PVDONE + 2, BAR_COUNT = BAR_OFREE + 1 };

__device__ __forceinline__ float fast_exp2(float x) {       // ex2.approx: 2 ulp, exp2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

// one accumulator row (OC x 32 fp32 columns, scaled by `f`) -> global, packed bf16/fp16 or (parity build) fp32
template <int OC, bool BF16>
__device__ __forceinline__ void store_row(const T4& t, int f32, int b, int h, long long tok, int D, const uint32_t (&ov)[OC][32],
                                          float f) {
  if (f32) {
    float* dst = row_ptr_w<float>(t, b, h, tok);
#pragma unroll
    for (int q4 = 0; q4 < OC; ++q4)
#pragma unroll
      for (int v4 = 0; v4 < 8; ++v4)
        if (q4 * 32 + v4 * 4 < D)
          *reinterpret_cast<float4*>(dst + q4 * 32 + v4 * 4) =
              make_float4(__uint_as_float(ov[q4][v4 * 4 + 0]) * f, __uint_as_float(ov[q4][v4 * 4 + 1]) * f,
                          __uint_as_float(ov[q4][v4 * 4 + 2]) * f, __uint_as_float(ov[q4][v4 * 4 + 3]) * f);
    return;
  }
  char* dst = t.p + ((long long)b * t.sb + (long long)h * t.sh + tok * t.st) * 2;
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
        *reinterpret_cast<uint4*>(dst + (q4 * 32 + v8 * 8) * 2) = pkt;
      }
    }
}

// One 64-column local block for one thread (= one query row).  `s0/s1` hold the raw fp32 scores.
// Pass 1 produces t_j and returns the block maximum of the log2-domain logits.  In the plain case (no bias table,
// no padding mask) t_j stays the RAW score and the scale is folded into pass 2's FFMA (p = ex2(t*c - m)); otherwise
// t_j is the finished logit (masked -> -inf).  Four independent max chains keep the FMNMX latency off the critical path.
template <int W, bool HAS_TAB, bool MASKED, int NV = W * W>
__device__ __forceinline__ float block_logits(float (&t)[64], const uint32_t (&s0)[32], const uint32_t (&s1)[32], float c,
                                              const float* __restrict__ tab_base, int krows, int kcols) {
  constexpr int TW = 4 * W - 1;
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float x = __uint_as_float(j < 32 ? s0[j] : s1[j - 32]);
    if constexpr (HAS_TAB || MASKED) {
      x *= c;
      if constexpr (HAS_TAB) x += tab_base[-((j / W) * TW + (j % W))];
      if constexpr (MASKED) x = ((j / W) < krows && (j % W) < kcols) ? x : -INFINITY;
    }
    t[j] = x;
    mx[j & 3] = fmaxf(mx[j & 3], x);
  }
  const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
  return (HAS_TAB || MASKED) ? m : m * c;          // c > 0
}

template <int DP, int W, bool BF16>
__global__ void __launch_bounds__(kThreads, 2)
vil_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmKg,
                  const __grid_constant__ CUtensorMap tmVg, const FwdArgs a) {
  using SM = FwdSmem<DP>;
  constexpr int ROWB = SM::ROWB;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;                       // stride between 8-row groups of a swizzled tile
  constexpr int W2 = W * W;
  constexpr int TW = 4 * W - 1;
  const Geo& geo = a.geo;

  extern __shared__ unsigned char smem_raw[];
  // pointer arithmetic on the __shared__ symbol (no integer round trip) keeps the address space visible to nvcc: LDS / STS
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sQ = smem + SM::OFF_Q;
  unsigned char* sKV = smem + SM::OFF_KV;
  float* tab = reinterpret_cast<float*>(smem + SM::OFF_TAB);
  const int tabn = a.has_tab ? TW * TW : 0;
  float* g2l_s = tab + geo.H * tabn;                        // [H][16]
  const int bars_off = (SM::OFF_TAB + (geo.H * tabn + geo.H * 16) * 4 + 15) & ~15;
  uint64_t* bars_p = reinterpret_cast<uint64_t*>(smem + bars_off);
  const uint32_t bars = smem_u32(bars_p);                   // shared-space address; barrier i lives at bars + 8 i
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_p + BAR_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5;

  // ---------------------------------------------------------------- one-time setup
  // zero the operand tiles once: rows a TMA box never writes (>= w*w of a slot / chunk) must stay finite
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
    for (int i = 0; i < 2; ++i) {
      mbar_init((bars + 8u * (BAR_QFULL + i)), 1); mbar_init((bars + 8u * (BAR_QEMPTY + i)), 1);
      mbar_init((bars + 8u * (BAR_SFULL + i)), 1); mbar_init((bars + 8u * (BAR_PFULL + i)), 128); mbar_init((bars + 8u * (BAR_PVDONE + i)), 1);
    }
    for (int i = 0; i < kStages; ++i) { mbar_init((bars + 8u * (BAR_KVFULL + i)), 1); mbar_init((bars + 8u * (BAR_KVEMPTY + i)), 1); }
    mbar_init((bars + 8u * (BAR_OFREE)), 128);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async();            // the generic-proxy zero fill must be visible to TMA / UMMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t TM_S0 = tmem, TM_O = tmem + 128;           // S buffers: [0,64) and [64,128); O: [128, 128+DP)

  const int units_per_bh = geo.mx * a.cpairs;

  if (warp == 4) {
    // ================================================================= TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      uint32_t stage = 0, kv_phase = 0, uc = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh = unit / units_per_bh, rem = unit % units_per_bh;
        const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
        const uint32_t qb = uc & 1, qphase = (uc >> 1) & 1;
        if (uc >= 2) mbar_wait((bars + 8u * (BAR_QEMPTY + qb)), qphase ^ 1);
        const bool hasB = 2 * Cp + 1 < geo.my;
        mbar_arrive_expect_tx((bars + 8u * (BAR_QFULL + qb)), (hasB ? 2 : 1) * W2 * ROWB);
        tma_load_5d(sQ + qb * SM::Q_BYTES, &tmQ, (bars + 8u * (BAR_QFULL + qb)), 0, (2 * Cp) * W, R * W, h, b);
        if (hasB) tma_load_5d(sQ + qb * SM::Q_BYTES + 64 * ROWB, &tmQ, (bars + 8u * (BAR_QFULL + qb)), 0, (2 * Cp + 1) * W, R * W, h, b);
        BlockWalk wk; wk.init(geo, R, Cp);
        int type, KR, KC;
        while (wk.next(geo, type, KR, KC)) {
          mbar_wait((bars + 8u * (BAR_KVEMPTY + stage)), kv_phase ^ 1);
          unsigned char* dK = sKV + stage * SM::STAGE_BYTES;
          unsigned char* dV = dK + SM::KV_BYTES;
          if (type == 1) {
            mbar_arrive_expect_tx((bars + 8u * (BAR_KVFULL + stage)), 2 * 16 * ROWB);
            tma_load_4d(dK, &tmKg, (bars + 8u * (BAR_KVFULL + stage)), 0, 0, h, b);
            tma_load_4d(dV, &tmVg, (bars + 8u * (BAR_KVFULL + stage)), 0, 0, h, b);
          } else {
            mbar_arrive_expect_tx((bars + 8u * (BAR_KVFULL + stage)), 2 * W2 * ROWB);
            tma_load_5d(dK, &tmK, (bars + 8u * (BAR_KVFULL + stage)), 0, KC * W, KR * W, h, b);
            tma_load_5d(dV, &tmV, (bars + 8u * (BAR_KVFULL + stage)), 0, KC * W, KR * W, h, b);
          }
          if (++stage == kStages) { stage = 0; kv_phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ================================================================= MMA issuer (one elected thread)
    if (elect_one()) {
      constexpr uint32_t IDESC_S = make_idesc(128, 64, BF16, false, false);
      constexpr uint32_t IDESC_SG = make_idesc(128, 16, BF16, false, false);
      constexpr uint32_t IDESC_O = make_idesc(128, DP, BF16, false, true);
      uint32_t stage = 0, kv_phase = 0, uc = 0, G = 0;        // G: running block counter (S/P buffer = G & 1)
      VIL_TRACE_DECL(2)
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int rem = unit % units_per_bh;
        const int R = rem / a.cpairs, Cp = rem % a.cpairs;
        const uint32_t qb = uc & 1, qphase = (uc >> 1) & 1;
        mbar_wait((bars + 8u * (BAR_QFULL + qb)), qphase);
        const uint32_t qaddr = smem_u32(sQ + qb * SM::Q_BYTES);

        // descriptors are built BEFORE the barrier waits: only the tcgen05.mma issues follow a completed wait
        constexpr int KS = DP / 16;
        uint64_t qd[KS], kd[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) qd[k] = make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT);
        auto prep_S = [&](uint32_t st) {
          const uint32_t kaddr = smem_u32(sKV + st * SM::STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < KS; ++k) kd[k] = make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT);
        };
        auto issue_S = [&](int type, uint32_t g) {
          const uint32_t d = TM_S0 + (g & 1) * 64;
#pragma unroll
          for (int k = 0; k < KS; ++k) mma_ss(d, qd[k], kd[k], type == 1 ? IDESC_SG : IDESC_S, k > 0);
          mma_commit((bars + 8u * (BAR_SFULL + (g & 1))));
        };

        BlockWalk wk; wk.init(geo, R, Cp);
        int type, KR, KC;
        bool have = wk.next(geo, type, KR, KC);
        // first S of the unit
        prep_S(stage);
        mbar_wait((bars + 8u * (BAR_KVFULL + stage)), kv_phase);
        tc_fence_after();
        issue_S(type, G);
        bool first = true;
        while (have) {
          const uint32_t cur_stage = stage, cur_g = G;
          const int cur_type = type;
          uint64_t vdsc[4];                                  // B operand of O += P V: the V tile of block j, MN-major
          {
            const uint32_t vaddr = smem_u32(sKV + cur_stage * SM::STAGE_BYTES + SM::KV_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) vdsc[k] = make_smem_desc(vaddr + k * 16 * ROWB, 16, SBO, LAYOUT);
          }
          if (++stage == kStages) { stage = 0; kv_phase ^= 1; }
          ++G;
          have = wk.next(geo, type, KR, KC);
          if (have) {
            VIL_TR(10);
            prep_S(stage);
            mbar_wait((bars + 8u * (BAR_KVFULL + stage)), kv_phase);
            tc_fence_after();
            issue_S(type, G);                                // S_{j+1} overlaps the softmax of block j
            VIL_TR(11);
          } else {
            mma_commit((bars + 8u * (BAR_QEMPTY + qb)));              // every S of this unit has been issued
          }
          mbar_wait((bars + 8u * (BAR_PFULL + (cur_g & 1))), (cur_g >> 1) & 1);
          VIL_TR(12);
          if (first && uc > 0) mbar_wait((bars + 8u * (BAR_OFREE)), (uc - 1) & 1);     // previous unit's O has been read
          tc_fence_after();
          const uint32_t paddr = TM_S0 + (cur_g & 1) * 64;
          if (cur_type == 1) {
            mma_ts(TM_O, paddr, vdsc[0], IDESC_O, !first);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ts(TM_O, paddr + k * 8, vdsc[k], IDESC_O, (!first) || k > 0);
          }
          mma_commit((bars + 8u * (BAR_KVEMPTY + cur_stage)));
          mma_commit((bars + 8u * (BAR_PVDONE + (cur_g & 1))));
          VIL_TR(13);
          first = false;
        }
      }
    }
  } else {
    // ================================================================= softmax warps (thread = query row = TMEM lane)
    const int row = tid;                 // 0..127
    const int slot = row >> 6, l = row & 63;
    const int qr = l / W, qc = l % W;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t uc = 0, G = 0;
    VIL_TRACE_DECL(tid == 0 ? 0 : (tid == 64 ? 1 : -1))
    for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
      const int bh = unit / units_per_bh, rem = unit % units_per_bh;
      const int b = bh / geo.H, h = bh % geo.H, R = rem / a.cpairs, Cp = rem % a.cpairs;
      const int C = 2 * Cp + slot;
      const int r = R * W + qr, c = C * W + qc;
      const bool slot_ok = C < geo.my;
      const bool row_ok = slot_ok && l < W2 && r < geo.nx && c < geo.ny;
      float m_use = -INFINITY, l_run = 0.f;
      const float* tab_h = tab + h * tabn;
      BlockWalk wk; wk.init(geo, R, Cp);
      int type, KR, KC;
      bool first = true;
      while (wk.next(geo, type, KR, KC)) {
        const uint32_t buf = G & 1;
        VIL_TR(1);
        mbar_wait((bars + 8u * (BAR_SFULL + buf)), (G >> 1) & 1);
        VIL_TR(2);
        tc_fence_after();
        const uint32_t saddr = TM_S0 + buf * 64 + lane_base;
        float p_scale_needed = 1.f;   // O rescale factor decided below
        uint32_t pk[32];
        if (type == 1) {
          // ---- global keys: 16 columns, bias g2l[1][h][t]
          uint32_t s[16];
          tmem_ld_x16(saddr, s);
          tmem_ld_wait();
          float t[16], mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            t[j] = (j < geo.g) ? fmaf(__uint_as_float(s[j]), a.scale_log2, g2l_s[h * 16 + j]) : -INFINITY;
            mx = fmaxf(mx, t[j]);
          }
          m_use = mx;                                          // always the first block of a unit
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float p0 = fast_exp2(t[j] - m_use), p1 = fast_exp2(t[j + 1] - m_use);
            sum += p0 + p1;
            pk[j >> 1] = pack2<BF16>(p0, p1);
          }
          l_run = sum;
          uint32_t p8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) p8[j] = pk[j];
          tmem_st_x8(saddr, p8);
        } else {
          const int dR = KR - R, dC = KC - C;
          const bool use = wk.used_by(slot);                             // warp-uniform (slot is per warp pair)
          if (!use) {
#pragma unroll
            for (int j = 0; j < 32; ++j) pk[j] = 0u;
            tmem_st_x32(saddr, pk);
          } else {
            uint32_t s0[32], s1[32];
            tmem_ld_x32(saddr, s0);
            tmem_ld_x32(saddr + 32, s1);
            tmem_ld_wait();
            float t[64];
            const int krows = min(W, geo.nx - KR * W), kcols = min(W, geo.ny - KC * W);
            const bool masked = (krows < W) || (kcols < W);
            const float* tb = tab_h + ((qr - dR * W + 2 * W - 1) * TW + (qc - dC * W + 2 * W - 1));
            float mx;
            if (a.has_tab) {
              mx = masked ? block_logits<W, true, true>(t, s0, s1, a.scale_log2, tb, krows, kcols)
                          : block_logits<W, true, false>(t, s0, s1, a.scale_log2, tb, krows, kcols);
            } else {
              mx = masked ? block_logits<W, false, true>(t, s0, s1, a.scale_log2, tb, krows, kcols)
                          : block_logits<W, false, false>(t, s0, s1, a.scale_log2, tb, krows, kcols);
            }
            // ---- lazy online-softmax rescale (log2 domain): only when the running max grows by more than 2^8
            float m_new = fmaxf(m_use, mx);
            bool need = first ? false : (m_new > m_use + 8.f);
            if (first) m_use = m_new;
            if (__any_sync(0xffffffffu, need)) {
              // O must be stable: the PV of the previous block has completed
              mbar_wait((bars + 8u * (BAR_PVDONE + ((G - 1) & 1))), ((G - 1) >> 1) & 1);
              tc_fence_after();
              const float f = need ? fast_exp2(m_use - m_new) : 1.f;     // m_use == -inf -> 0
              if (need) { m_use = m_new; l_run *= f; }
              constexpr int OC = DP / 32;
#pragma unroll
              for (int q4 = 0; q4 < OC; ++q4) {
                uint32_t ov[32];
                tmem_ld_x32(TM_O + lane_base + q4 * 32, ov);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * f);
                tmem_st_x32(TM_O + lane_base + q4 * 32, ov);
              }
            }
            const float m_eff = (m_use == -INFINITY) ? 0.f : m_use;
            const float cc = (a.has_tab || masked) ? 1.f : a.scale_log2;      // see block_logits
            float sum[4] = {0.f, 0.f, 0.f, 0.f};          // two packed accumulator pairs
#pragma unroll
            for (int j = 0; j < 64; j += 2) {
              float p0 = 0.f, p1 = 0.f;
              if (j < W2) {
                float x0, x1;
                ffma2(x0, x1, t[j], (j + 1 < W2) ? t[j + 1] : 0.f, cc, cc, -m_eff, -m_eff);
                p0 = fast_exp2(x0);
                p1 = (j + 1 < W2) ? fast_exp2(x1) : 0.f;
                const int k = j & 2;
                fadd2(sum[k], sum[k + 1], sum[k], sum[k + 1], p0, p1);
              }
              pk[j >> 1] = pack2<BF16>(p0, p1);
            }
            l_run += (sum[0] + sum[1]) + (sum[2] + sum[3]);
            tmem_st_x32(saddr, pk);
          }
        }
        (void)p_scale_needed;
        VIL_TR(3);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive((bars + 8u * (BAR_PFULL + buf)));
        VIL_TR(4);
        first = false;
        ++G;
      }
      // ---- epilogue: O / l -> global, LSE
      VIL_TR(5);
      mbar_wait((bars + 8u * (BAR_PVDONE + ((G - 1) & 1))), ((G - 1) >> 1) & 1);
      VIL_TR(6);
      tc_fence_after();
      constexpr int OC = DP / 32;
      uint32_t ov[OC][32];
#pragma unroll
      for (int q4 = 0; q4 < OC; ++q4) tmem_ld_x32(TM_O + lane_base + q4 * 32, ov[q4]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive((bars + 8u * (BAR_OFREE)));
      if (row_ok) {
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        const long long tok = (long long)r * geo.ny + c;
        store_row<OC, BF16>(a.o, a.out_f32, b, h, tok, geo.D, ov, inv);
        a.lse[((long long)b * geo.H + h) * geo.Nloc + tok] = (m_use + log2f(l_run)) * 0.6931471805599453f;
      }
    }
  }
  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

}  // namespace tc
}  // namespace vil
