// tcgen05 / TMA forward for chunk sizes w > 8 (w in {12, 15, 31}): same pipeline as vil_tc_fwd_kernel, different
// tiling.  A chunk (w x w tokens) is cut into NP "pieces" of PR = floor(64 / w) whole chunk rows (PR*w <= 64 tokens);
// a unit = two consecutive pieces of ONE query chunk (the two 64-row slots of the 128-row MMA tile), and the key
// blocks are the NP pieces of each visited neighbour chunk, each staged by one TMA box (D, w, PR).  Both slots visit
// the same blocks.  The last piece of a chunk may be short: its surplus box rows belong to the next chunk row and are
// masked (keys) / not stored (queries).
#pragma once
#include "vil_tc_fwd.cuh"
#include "vil_tc_bwd.cuh"

namespace vil {
namespace tc {

// 3 x 3 chunk window x NP pieces, row-major over chunks, pieces innermost
//
// Pieces whose first row lies below the image (zero padding of the last chunk row: 27 of 31 rows at w = 31 on 128 x 128 tokens) are
// never visited: as keys they would be fully masked, as queries nothing of them is stored.  Every role of a kernel walks with
// this one struct and skips empty units with big_unit_empty(), so the producer / MMA / compute pipelines stay in step.
struct BigWalk {
  uint32_t m9;
  int R, C, bit, pk_next, np;
  int nx, w, pr;
  bool global_pending;
  __device__ __forceinline__ void init(const Geo& g, int R_, int C_, int np_, bool mirror = false, bool with_global = true) {
    R = R_; C = C_; np = np_; pk_next = np_; bit = 0; m9 = 0;
    nx = g.nx; w = g.w; pr = 64 / g.w;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int dR = i / 3 - 1, dC = i % 3 - 1;
      const bool used = mirror ? offset_used(g, -dR, -dC) : offset_used(g, dR, dC);
      if (used && R + dR >= 0 && R + dR < g.mx && C + dC >= 0 && C + dC < g.my) m9 |= 1u << i;
    }
    global_pending = with_global && g.g > 0;
  }
  __device__ __forceinline__ bool next(int& type, int& KR, int& KC, int& PK) {
    if (global_pending) { global_pending = false; type = 1; KR = KC = PK = 0; return true; }
    for (;;) {
      if (pk_next >= np) {
        if (m9 == 0) return false;
        bit = __ffs(m9) - 1;
        m9 &= m9 - 1;
        pk_next = 0;
      }
      type = 0; PK = pk_next++; KR = R + bit / 3 - 1; KC = C + bit % 3 - 1;
      if (KR * w + PK * pr < nx) return true;
      pk_next = np;                       // this piece and the rest of its chunk lie below the image
    }
  }
};
// unit = (chunk (R,C), piece pair pp): empty when its first piece starts below the image (piece 0 of a chunk never does)
__device__ __forceinline__ bool big_unit_empty(const Geo& g, int rem, int npp) {
  const int R = rem / (g.my * npp), pp = rem % npp;
  return R * g.w + 2 * pp * (64 / g.w) >= g.nx;
}

template <int DP, int W, bool BF16>
__global__ void __launch_bounds__(kThreads, 2)
vil_tc_fwd_big_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmKg,
                  const __grid_constant__ CUtensorMap tmVg, const FwdArgs a) {
  using SM = FwdSmem<DP>;
  constexpr int ROWB = SM::ROWB;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;                       // stride between 8-row groups of a swizzled tile
  constexpr int PR = 64 / W;                  // chunk rows per piece
  constexpr int RW = PR * W;                  // rows (keys / queries) per piece, <= 64
  constexpr int NP = (W + PR - 1) / PR;       // pieces per chunk
  constexpr int NPP = (NP + 1) / 2;           // slot pairs per chunk
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

  const int units_per_bh = geo.mx * geo.my * NPP;

  if (warp == 4) {
    // ================================================================= TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      uint32_t stage = 0, kv_phase = 0, uc = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh = unit / units_per_bh, rem = unit % units_per_bh;
        if (big_unit_empty(geo, rem, NPP)) { --uc; continue; }      // nothing to compute or store: skipped by every role
        const int b = bh / geo.H, h = bh % geo.H, R = rem / (geo.my * NPP), C = (rem / NPP) % geo.my, pp = rem % NPP;
        const uint32_t qb = uc & 1, qphase = (uc >> 1) & 1;
        if (uc >= 2) mbar_wait((bars + 8u * (BAR_QEMPTY + qb)), qphase ^ 1);
        const bool hasB = 2 * pp + 1 < NP;
        mbar_arrive_expect_tx((bars + 8u * (BAR_QFULL + qb)), (hasB ? 2 : 1) * RW * ROWB);
        tma_load_5d(sQ + qb * SM::Q_BYTES, &tmQ, (bars + 8u * (BAR_QFULL + qb)), 0, C * W, R * W + (2 * pp) * PR, h, b);
        if (hasB) tma_load_5d(sQ + qb * SM::Q_BYTES + 64 * ROWB, &tmQ, (bars + 8u * (BAR_QFULL + qb)), 0, C * W, R * W + (2 * pp + 1) * PR, h, b);
        BigWalk wk; wk.init(geo, R, C, NP);
        int type, KR, KC, PK;
        while (wk.next(type, KR, KC, PK)) {
          mbar_wait((bars + 8u * (BAR_KVEMPTY + stage)), kv_phase ^ 1);
          unsigned char* dK = sKV + stage * SM::STAGE_BYTES;
          unsigned char* dV = dK + SM::KV_BYTES;
          if (type == 1) {
            mbar_arrive_expect_tx((bars + 8u * (BAR_KVFULL + stage)), 2 * 16 * ROWB);
            tma_load_4d(dK, &tmKg, (bars + 8u * (BAR_KVFULL + stage)), 0, 0, h, b);
            tma_load_4d(dV, &tmVg, (bars + 8u * (BAR_KVFULL + stage)), 0, 0, h, b);
          } else {
            mbar_arrive_expect_tx((bars + 8u * (BAR_KVFULL + stage)), 2 * RW * ROWB);
            tma_load_5d(dK, &tmK, (bars + 8u * (BAR_KVFULL + stage)), 0, KC * W, KR * W + PK * PR, h, b);
            tma_load_5d(dV, &tmV, (bars + 8u * (BAR_KVFULL + stage)), 0, KC * W, KR * W + PK * PR, h, b);
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
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int rem = unit % units_per_bh;
        if (big_unit_empty(geo, rem, NPP)) { --uc; continue; }      // nothing to compute or store: skipped by every role
        const int R = rem / (geo.my * NPP), C = (rem / NPP) % geo.my;
        const uint32_t qb = uc & 1, qphase = (uc >> 1) & 1;
        mbar_wait((bars + 8u * (BAR_QFULL + qb)), qphase);
        const uint32_t qaddr = smem_u32(sQ + qb * SM::Q_BYTES);

        auto issue_S = [&](uint32_t st, int type, uint32_t g) {
          const uint32_t kaddr = smem_u32(sKV + st * SM::STAGE_BYTES);
          const uint32_t d = TM_S0 + (g & 1) * 64;
#pragma unroll
          for (int k = 0; k < DP / 16; ++k)
            mma_ss(d, make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT), make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT),
                   type == 1 ? IDESC_SG : IDESC_S, k > 0);
          mma_commit((bars + 8u * (BAR_SFULL + (g & 1))));
        };

        BigWalk wk; wk.init(geo, R, C, NP);
        int type, KR, KC, PK;
        bool have = wk.next(type, KR, KC, PK);
        // first S of the unit
        mbar_wait((bars + 8u * (BAR_KVFULL + stage)), kv_phase);
        tc_fence_after();
        issue_S(stage, type, G);
        bool first = true;
        while (have) {
          const uint32_t cur_stage = stage, cur_g = G;
          const int cur_type = type;
          if (++stage == kStages) { stage = 0; kv_phase ^= 1; }
          ++G;
          have = wk.next(type, KR, KC, PK);
          if (have) {
            mbar_wait((bars + 8u * (BAR_KVFULL + stage)), kv_phase);
            tc_fence_after();
            issue_S(stage, type, G);                         // S_{j+1} overlaps the softmax of block j
          } else {
            mma_commit((bars + 8u * (BAR_QEMPTY + qb)));              // every S of this unit has been issued
          }
          mbar_wait((bars + 8u * (BAR_PFULL + (cur_g & 1))), (cur_g >> 1) & 1);
          if (first && uc > 0) mbar_wait((bars + 8u * (BAR_OFREE)), (uc - 1) & 1);     // previous unit's O has been read
          tc_fence_after();
          const uint32_t vaddr = smem_u32(sKV + cur_stage * SM::STAGE_BYTES + SM::KV_BYTES);
          const uint32_t paddr = TM_S0 + (cur_g & 1) * 64;
          const int ksteps = cur_type == 1 ? 1 : 4;
          for (int k = 0; k < ksteps; ++k)
            mma_ts(TM_O, paddr + k * 8, make_smem_desc(vaddr + k * 16 * ROWB, 16, SBO, LAYOUT), IDESC_O, (!first) || k > 0);
          mma_commit((bars + 8u * (BAR_KVEMPTY + cur_stage)));
          mma_commit((bars + 8u * (BAR_PVDONE + (cur_g & 1))));
          first = false;
        }
      }
    }
  } else {
    // ================================================================= softmax warps (thread = query row = TMEM lane)
    const int row = tid;                 // 0..127
    const int slot = row >> 6, l = row & 63;
    const int lr = l / W, qc = l % W;                 // row within the piece, column within the chunk
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t uc = 0, G = 0;
    for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
      const int bh = unit / units_per_bh, rem = unit % units_per_bh;
      if (big_unit_empty(geo, rem, NPP)) { --uc; continue; }      // nothing to compute or store: skipped by every role
      const int b = bh / geo.H, h = bh % geo.H, R = rem / (geo.my * NPP), C = (rem / NPP) % geo.my, pp = rem % NPP;
      const int pq = 2 * pp + slot;                   // query piece of this slot
      const int qr = pq * PR + lr;                    // row within the chunk
      const int r = R * W + qr, c = C * W + qc;
      const bool slot_ok = pq < NP;
      const bool row_ok = slot_ok && l < RW && qr < W && r < geo.nx && c < geo.ny;
      float m_use = -INFINITY, l_run = 0.f;
      const float* tab_h = tab + h * tabn;
      BigWalk wk; wk.init(geo, R, C, NP);
      int type, KR, KC, PK;
      bool first = true;
      while (wk.next(type, KR, KC, PK)) {
        const uint32_t buf = G & 1;
        mbar_wait((bars + 8u * (BAR_SFULL + buf)), (G >> 1) & 1);
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
          const bool use = slot_ok;                                      // both pieces of a chunk visit the same key blocks
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
            // valid key rows of this piece: inside the chunk (last piece may be short) and inside the image
            const int krows = min(min(PR, W - PK * PR), geo.nx - KR * W - PK * PR), kcols = min(W, geo.ny - KC * W);
            const bool masked = (krows < PR) || (kcols < W);
            const float* tb = tab_h + ((qr - dR * W - PK * PR + 2 * W - 1) * TW + (qc - dC * W + 2 * W - 1));
            float mx;
            if (a.has_tab) {
              mx = masked ? block_logits<W, true, true, RW>(t, s0, s1, a.scale_log2, tb, krows, kcols)
                          : block_logits<W, true, false, RW>(t, s0, s1, a.scale_log2, tb, krows, kcols);
            } else {
              mx = masked ? block_logits<W, false, true, RW>(t, s0, s1, a.scale_log2, tb, krows, kcols)
                          : block_logits<W, false, false, RW>(t, s0, s1, a.scale_log2, tb, krows, kcols);
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
            float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 64; j += 2) {
              float p0 = 0.f, p1 = 0.f;
              if (j < RW) {
                float x0, x1;
                ffma2(x0, x1, t[j], (j + 1 < RW) ? t[j + 1] : 0.f, cc, cc, -m_eff, -m_eff);
                p0 = fast_exp2(x0);
                p1 = (j + 1 < RW) ? fast_exp2(x1) : 0.f;
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
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive((bars + 8u * (BAR_PFULL + buf)));
        first = false;
        ++G;
      }
      // ---- epilogue: O / l -> global, LSE
      mbar_wait((bars + 8u * (BAR_PVDONE + ((G - 1) & 1))), ((G - 1) >> 1) & 1);
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


// token-ordered (lse, delta) -> piece-ordered, 64-padded (lse2, delta): index ((((b*H+h)*mx+R)*my+C)*NP+piece)*64 + l
template <int W>
__global__ void vil_tc_bwd_prep_big(Geo geo, const float* __restrict__ lse, const float* __restrict__ delta,
                                    float* __restrict__ lse2c, float* __restrict__ deltac) {
  constexpr int PR = 64 / W, RW = PR * W, NP = (W + PR - 1) / PR;
  const long long total = (long long)geo.B * geo.H * geo.mx * geo.my * NP * 64;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int l = (int)(idx & 63);
  long long c = idx >> 6;
  const int piece = (int)(c % NP); c /= NP;
  const int C = (int)(c % geo.my); c /= geo.my;
  const int R = (int)(c % geo.mx); c /= geo.mx;       // c = b*H + h
  const int qr = piece * PR + l / W;
  const int r = R * W + qr, cc = C * W + l % W;
  float a = INFINITY, d = 0.f;
  if (l < RW && qr < W && r < geo.nx && cc < geo.ny) {
    const long long t = c * geo.Nloc + (long long)r * geo.ny + cc;
    a = lse[t] * 1.4426950408889634f;
    d = delta[t];
  }
  lse2c[idx] = a;
  deltac[idx] = d;
}

// ======================================================================================================== pass 1 (w > 8)
template <int DP, int W, bool BF16>
__global__ void __launch_bounds__(kBwdThreads, 2)
vil_tc_bwd_dq_big_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmKg, const __grid_constant__ CUtensorMap tmVg, const BwdArgs a) {
  using SM = BwdSmem<DP>;
  constexpr int ROWB = SM::ROWB, NS = SM::NS;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int PR = 64 / W, RW = PR * W, NP = (W + PR - 1) / PR, NPP = (NP + 1) / 2;
  constexpr int TW = 4 * W - 1;
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
  // S / dP are released as soon as the compute threads hold them in registers; dS has its own double buffer
  const uint32_t TM_S = tmem, TM_DP = tmem + 64, TM_DS = tmem + 128, TM_ACC = tmem + 192;
  const int units_per_bh = geo.mx * geo.my * NPP;

  if (warp == 8) {
    // ================================================================= TMA producer
    if (elect_one()) {
      uint32_t stage = 0, yphase = 0, uc = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh_ = unit / units_per_bh, rem = unit % units_per_bh;
        if (big_unit_empty(geo, rem, NPP)) { --uc; continue; }      // nothing to compute or store: skipped by every role
        const int b = bh_ / geo.H, h = bh_ % geo.H, R = rem / (geo.my * NPP), C = (rem / NPP) % geo.my, pp = rem % NPP;
        (void)b; (void)h; (void)pp;
        const uint32_t xb = uc & 1, xphase = (uc >> 1) & 1;
        if (uc >= 2) mbar_wait((bars + 8u * (BB_XEMPTY + xb)), xphase ^ 1);
        unsigned char* sQ = sX + xb * 2 * SM::X_BYTES;
        unsigned char* sDO = sQ + SM::X_BYTES;
        const bool hasB = 2 * pp + 1 < NP;
        mbar_arrive_expect_tx((bars + 8u * (BB_XFULL + xb)), (hasB ? 4 : 2) * RW * ROWB);
        tma_load_5d(sQ, &tmQ, (bars + 8u * (BB_XFULL + xb)), 0, C * W, R * W + (2 * pp) * PR, h, b);
        tma_load_5d(sDO, &tmDO, (bars + 8u * (BB_XFULL + xb)), 0, C * W, R * W + (2 * pp) * PR, h, b);
        if (hasB) {
          tma_load_5d(sQ + 64 * ROWB, &tmQ, (bars + 8u * (BB_XFULL + xb)), 0, C * W, R * W + (2 * pp + 1) * PR, h, b);
          tma_load_5d(sDO + 64 * ROWB, &tmDO, (bars + 8u * (BB_XFULL + xb)), 0, C * W, R * W + (2 * pp + 1) * PR, h, b);
        }
        BigWalk wk; wk.init(geo, R, C, NP);
        int type, KR, KC, PK;
        while (wk.next(type, KR, KC, PK)) {
          mbar_wait((bars + 8u * (BB_YEMPTY + stage)), yphase ^ 1);
          unsigned char* dK = sY + stage * SM::STAGE_STRIDE;
          unsigned char* dV = dK + SM::Y_BYTES;
          if (type == 1) {
            mbar_arrive_expect_tx((bars + 8u * (BB_YFULL + stage)), 2 * 16 * ROWB);
            tma_load_4d(dK, &tmKg, (bars + 8u * (BB_YFULL + stage)), 0, 0, h, b);
            tma_load_4d(dV, &tmVg, (bars + 8u * (BB_YFULL + stage)), 0, 0, h, b);
          } else {
            mbar_arrive_expect_tx((bars + 8u * (BB_YFULL + stage)), 2 * RW * ROWB);
            tma_load_5d(dK, &tmK, (bars + 8u * (BB_YFULL + stage)), 0, KC * W, KR * W + PK * PR, h, b);
            tma_load_5d(dV, &tmV, (bars + 8u * (BB_YFULL + stage)), 0, KC * W, KR * W + PK * PR, h, b);
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
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh_ = unit / units_per_bh, rem = unit % units_per_bh;
        if (big_unit_empty(geo, rem, NPP)) { --uc; continue; }      // nothing to compute or store: skipped by every role
        const int b = bh_ / geo.H, h = bh_ % geo.H, R = rem / (geo.my * NPP), C = (rem / NPP) % geo.my, pp = rem % NPP;
        (void)b; (void)h; (void)pp;
        const uint32_t xb = uc & 1, xphase = (uc >> 1) & 1;
        mbar_wait((bars + 8u * (BB_XFULL + xb)), xphase);
        const uint32_t qaddr = smem_u32(sX + xb * 2 * SM::X_BYTES), doaddr = qaddr + SM::X_BYTES;
        auto issue_SdP = [&](uint32_t st, int type) {
          const uint32_t kaddr = smem_u32(sY + st * SM::STAGE_STRIDE), vaddr = kaddr + SM::Y_BYTES;
          const uint32_t idesc = type == 1 ? IDESC_SG : IDESC_S;
#pragma unroll
          for (int k = 0; k < DP / 16; ++k)
            mma_ss(TM_S, make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT), make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT), idesc, k > 0);
#pragma unroll
          for (int k = 0; k < DP / 16; ++k)
            mma_ss(TM_DP, make_smem_desc(doaddr + k * 32, 16, SBO, LAYOUT), make_smem_desc(vaddr + k * 32, 16, SBO, LAYOUT), idesc, k > 0);
          mma_commit((bars + 8u * (BB_SFULL)));
        };
        BigWalk wk; wk.init(geo, R, C, NP);
        int type, KR, KC, PK;
        bool have = wk.next(type, KR, KC, PK);
        mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase);
        tc_fence_after();
        issue_SdP(stage, type);
        bool first = true;
        while (have) {
          const uint32_t cur_stage = stage;
          const int cur_type = type;
          if (++stage == NS) { stage = 0; yphase ^= 1; }
          have = wk.next(type, KR, KC, PK);
          if (have) mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase);
          if (have) {
            mbar_wait((bars + 8u * (BB_CONS)), G & 1);                // S_j / dP_j are in the threads' registers
            tc_fence_after();
            issue_SdP(stage, type);                          // overlaps the threads' exp / dS work on block j
          }
          mbar_wait((bars + 8u * (BB_DSFULL + (G & 1))), (G >> 1) & 1);
          if (first && uc > 0) mbar_wait((bars + 8u * (BB_ACCFREE)), (uc - 1) & 1);
          tc_fence_after();
          const uint32_t kaddr = smem_u32(sY + cur_stage * SM::STAGE_STRIDE);
          const uint32_t dsaddr = TM_DS + (G & 1) * 32;
          const int ksteps = cur_type == 1 ? 1 : 4;
          for (int k = 0; k < ksteps; ++k)
            mma_ts(TM_ACC, dsaddr + k * 8, make_smem_desc(kaddr + k * 16 * ROWB, 16, SBO, LAYOUT), IDESC_ACC, (!first) || k > 0);
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
    const int lr = l / W, qc = l % W;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t uc = 0, G = 0;
    for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
      const int bh = unit / units_per_bh, rem = unit % units_per_bh;
      if (big_unit_empty(geo, rem, NPP)) { --uc; continue; }      // nothing to compute or store: skipped by every role
      const int b = bh / geo.H, h = bh % geo.H, R = rem / (geo.my * NPP), C = (rem / NPP) % geo.my, pp = rem % NPP;
      const int pq = 2 * pp + slot;
      const int qr = pq * PR + lr;
      const int r = R * W + qr, c = C * W + qc;
      const bool slot_ok = pq < NP;
      const bool row_ok = slot_ok && l < RW && qr < W && r < geo.nx && c < geo.ny;
      float lse2 = INFINITY, del = 0.f;
      if (slot_ok) {
        const long long ci = ((((long long)bh * geo.mx + R) * geo.my + C) * NP + pq) * 64 + l;
        lse2 = a.lse2c[ci]; del = a.deltac[ci];
      }
      const float* tab_h = tab + h * tabn;
      BigWalk wk; wk.init(geo, R, C, NP);
      int type, KR, KC, PK;
      while (wk.next(type, KR, KC, PK)) {
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
          const bool use = slot_ok;
          uint32_t pk[16];
          if (!use) {
            tc_fence_before();
            mbar_arrive((bars + 8u * (BB_CONS)));
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          } else {
            const int krows = min(min(PR, W - PK * PR), geo.nx - KR * W - PK * PR), kcols = min(W, geo.ny - KC * W);
            const bool masked = (krows < PR) || (kcols < W);
            const float* tb = tab_h + ((qr - dR * W - PK * PR + 2 * W - 1) * TW + (qc - dC * W + 2 * W - 1));
            // two 16-column quarters per thread; the second one releases S / dP (BB_CONS) right after its loads
            const bool ht = a.has_tab != 0;
            float* e_row = nullptr;            // no bias-table gradient in this family (w > 8: SIMT backward)
            if (half == 0) {
              dq_quarter<W, 0, BF16, RW>(pk, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, lse2, del, 0u, e_row);
              dq_quarter<W, 16, BF16, RW>(pk + 8, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, lse2, del, (bars + 8u * (BB_CONS)), e_row);
            } else {
              dq_quarter<W, 32, BF16, RW>(pk, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, lse2, del, 0u, e_row);
              dq_quarter<W, 48, BF16, RW>(pk + 8, saddr, paddr, a.scale_log2, ht, tb, masked, krows, kcols, lse2, del, (bars + 8u * (BB_CONS)), e_row);
            }
          }
          tmem_st_x16(dsaddr + half * 16, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive((bars + 8u * (BB_DSFULL + (G & 1))));
        ++G;
      }
      mbar_wait((bars + 8u * (BB_ACCDONE)), uc & 1);
      tc_fence_after();
      constexpr int NC = DP / 2;
      uint32_t ov[NC];
      if constexpr (NC == 32) tmem_ld_x32(TM_ACC + lane_base + half * NC, ov); else tmem_ld_x16(TM_ACC + lane_base + half * NC, ov);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive((bars + 8u * (BB_ACCFREE)));
      if (row_ok) store_cols<NC, BF16>(a.out0, b, h, (long long)r * geo.ny + c, geo.D, half * NC, ov, a.scale, a.out_f32);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

// ======================================================================================================== pass 2 (w > 8)
template <int DP, int W, bool BF16>
__global__ void __launch_bounds__(kBwdThreads, 2)
vil_tc_bwd_dkv_big_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const BwdArgs a) {
  using SM = BwdSmem<DP>;
  constexpr int ROWB = SM::ROWB, NS = SM::NS;
  constexpr uint32_t LAYOUT = DP == 32 ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int PR = 64 / W, RW = PR * W, NP = (W + PR - 1) / PR, NPP = (NP + 1) / 2;
  constexpr int TW = 4 * W - 1;
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
  const int units_per_bh = geo.mx * geo.my * NPP;

  if (warp == 8) {
    if (elect_one()) {
      uint32_t stage = 0, yphase = 0, uc = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int bh = unit / units_per_bh, rem = unit % units_per_bh;
        if (big_unit_empty(geo, rem, NPP)) { --uc; continue; }      // nothing to compute or store: skipped by every role
        const int b = bh / geo.H, h = bh % geo.H, R = rem / (geo.my * NPP), C = (rem / NPP) % geo.my, pp = rem % NPP;
        const uint32_t xb = uc & 1, xphase = (uc >> 1) & 1;
        if (uc >= 2) mbar_wait((bars + 8u * (BB_XEMPTY + xb)), xphase ^ 1);
        unsigned char* sK = sX + xb * 2 * SM::X_BYTES;
        unsigned char* sV = sK + SM::X_BYTES;
        const bool hasB = 2 * pp + 1 < NP;
        mbar_arrive_expect_tx((bars + 8u * (BB_XFULL + xb)), (hasB ? 4 : 2) * RW * ROWB);
        tma_load_5d(sK, &tmK, (bars + 8u * (BB_XFULL + xb)), 0, C * W, R * W + (2 * pp) * PR, h, b);
        tma_load_5d(sV, &tmV, (bars + 8u * (BB_XFULL + xb)), 0, C * W, R * W + (2 * pp) * PR, h, b);
        if (hasB) {
          tma_load_5d(sK + 64 * ROWB, &tmK, (bars + 8u * (BB_XFULL + xb)), 0, C * W, R * W + (2 * pp + 1) * PR, h, b);
          tma_load_5d(sV + 64 * ROWB, &tmV, (bars + 8u * (BB_XFULL + xb)), 0, C * W, R * W + (2 * pp + 1) * PR, h, b);
        }
        BigWalk wk; wk.init(geo, R, C, NP, true, false);
        int type, QR, QC, PQ;
        while (wk.next(type, QR, QC, PQ)) {
          mbar_wait((bars + 8u * (BB_YEMPTY + stage)), yphase ^ 1);
          unsigned char* dQ = sY + stage * SM::STAGE_STRIDE;
          unsigned char* dG = dQ + SM::Y_BYTES;
          unsigned char* dL = dG + SM::Y_BYTES;
          mbar_arrive_expect_tx((bars + 8u * (BB_YFULL + stage)), 2 * RW * ROWB + 512);
          tma_load_5d(dQ, &tmQ, (bars + 8u * (BB_YFULL + stage)), 0, QC * W, QR * W + PQ * PR, h, b);
          tma_load_5d(dG, &tmDO, (bars + 8u * (BB_YFULL + stage)), 0, QC * W, QR * W + PQ * PR, h, b);
          const long long ci = ((((long long)bh * geo.mx + QR) * geo.my + QC) * NP + PQ) * 64;
          bulk_load_1d(dL, a.lse2c + ci, 256, (bars + 8u * (BB_YFULL + stage)));
          bulk_load_1d(dL + 256, a.deltac + ci, 256, (bars + 8u * (BB_YFULL + stage)));
          if (++stage == NS) { stage = 0; yphase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t IDESC_S = make_idesc(128, 64, BF16, false, false);
      constexpr uint32_t IDESC_ACC = make_idesc(128, DP, BF16, false, true);
      uint32_t stage = 0, yphase = 0, uc = 0, G = 0;
      for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
        const int rem = unit % units_per_bh;
        if (big_unit_empty(geo, rem, NPP)) { --uc; continue; }      // nothing to compute or store: skipped by every role
        const int R = rem / (geo.my * NPP), C = (rem / NPP) % geo.my;
        const uint32_t xb = uc & 1, xphase = (uc >> 1) & 1;
        mbar_wait((bars + 8u * (BB_XFULL + xb)), xphase);
        const uint32_t kaddr = smem_u32(sX + xb * 2 * SM::X_BYTES), vaddr = kaddr + SM::X_BYTES;
        auto issue_SdP = [&](uint32_t st) {
          const uint32_t qaddr = smem_u32(sY + st * SM::STAGE_STRIDE), gaddr = qaddr + SM::Y_BYTES;
#pragma unroll
          for (int k = 0; k < DP / 16; ++k)
            mma_ss(TM_S, make_smem_desc(kaddr + k * 32, 16, SBO, LAYOUT), make_smem_desc(qaddr + k * 32, 16, SBO, LAYOUT), IDESC_S, k > 0);
#pragma unroll
          for (int k = 0; k < DP / 16; ++k)
            mma_ss(TM_DP, make_smem_desc(vaddr + k * 32, 16, SBO, LAYOUT), make_smem_desc(gaddr + k * 32, 16, SBO, LAYOUT), IDESC_S, k > 0);
          mma_commit((bars + 8u * (BB_SFULL)));
        };
        BigWalk wk; wk.init(geo, R, C, NP, true, false);
        int type, QR, QC, PQ;
        bool have = wk.next(type, QR, QC, PQ);
        mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase);
        tc_fence_after();
        issue_SdP(stage);
        bool first = true;
        while (have) {
          const uint32_t cur_stage = stage;
          if (++stage == NS) { stage = 0; yphase ^= 1; }
          have = wk.next(type, QR, QC, PQ);
          if (have) mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase);
          if (kSplit && have) {
            mbar_wait((bars + 8u * (BB_CONS)), G & 1);                // S^T_j / dP^T_j are in the threads' registers
            tc_fence_after();
            issue_SdP(stage);
          }
          mbar_wait((bars + 8u * (BB_DSFULL + (G & 1))), (G >> 1) & 1);
          if (first && uc > 0) mbar_wait((bars + 8u * (BB_ACCFREE)), (uc - 1) & 1);
          tc_fence_after();
          const uint32_t qaddr = smem_u32(sY + cur_stage * SM::STAGE_STRIDE), gaddr = qaddr + SM::Y_BYTES;
          for (int k = 0; k < 4; ++k)       // dV += P^T dO
            mma_ts(TM_DV, TM_P + k * 8, make_smem_desc(gaddr + k * 16 * ROWB, 16, SBO, LAYOUT), IDESC_ACC, (!first) || k > 0);
          for (int k = 0; k < 4; ++k)       // dK += dS^T Q
            mma_ts(TM_DK, TM_DS + k * 8, make_smem_desc(qaddr + k * 16 * ROWB, 16, SBO, LAYOUT), IDESC_ACC, (!first) || k > 0);
          mma_commit((bars + 8u * (BB_YEMPTY + cur_stage)));
          if (kSplit) mma_commit((bars + 8u * (BB_PDONE)));           // P^T / dS^T columns may be rewritten
          first = false;
          ++G;
          if (have) {
            if (!kSplit) issue_SdP(stage);
          } else {
            mma_commit((bars + 8u * (BB_ACCDONE)));
            mma_commit((bars + 8u * (BB_XEMPTY + xb)));
          }
        }
      }
    }
  } else {
    const int row = tid & 127, half = tid >> 7, slot = row >> 6, l = row & 63;
    const int lr = l / W, kc = l % W;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t uc = 0, G = 0, stage = 0, yphase = 0;
    for (int unit = blockIdx.x; unit < a.num_units; unit += gridDim.x, ++uc) {
      const int bh = unit / units_per_bh, rem = unit % units_per_bh;
      if (big_unit_empty(geo, rem, NPP)) { --uc; continue; }      // nothing to compute or store: skipped by every role
      const int b = bh / geo.H, h = bh % geo.H, R = rem / (geo.my * NPP), C = (rem / NPP) % geo.my, pp = rem % NPP;
      const int pkk = 2 * pp + slot;                 // key piece of this slot
      const int kr = pkk * PR + lr;
      const int r = R * W + kr, c = C * W + kc;
      const bool slot_ok = pkk < NP;
      const bool row_ok = slot_ok && l < RW && kr < W && r < geo.nx && c < geo.ny;
      const float* tab_h = tab + h * tabn;
      BigWalk wk; wk.init(geo, R, C, NP, true, false);
      int type, QR, QC, PQ;
      while (wk.next(type, QR, QC, PQ)) {
        mbar_wait((bars + 8u * (BB_YFULL + stage)), yphase);     // lse2 / delta of this query block have landed
        mbar_wait((bars + 8u * (BB_SFULL)), G & 1);
        tc_fence_after();
        const float* ls = reinterpret_cast<const float*>(sY + stage * SM::STAGE_STRIDE + 2 * SM::Y_BYTES);
        const float* dl = ls + 64;
        const uint32_t saddr = TM_S + lane_base, paddr = TM_DP + lane_base;
        const int dR = R - QR, dC = C - QC;       // offset = key chunk - query chunk
        const bool use_w = slot_ok;                                 // warp-uniform
        const bool use = use_w && row_ok;
        uint32_t pp[16], pd[16];
        if (!use_w) {
          if (kSplit) { tc_fence_before(); mbar_arrive((bars + 8u * (BB_CONS))); }
#pragma unroll
          for (int j = 0; j < 16; ++j) { pp[j] = 0u; pd[j] = 0u; }
        } else {
          // bias index: dr = qr' - (dR*W + kr)  ->  base + qr'*TW + qc'
          const float* tb = tab_h + ((2 * W - 1 - dR * W - kr + PQ * PR) * TW + (2 * W - 1 - dC * W - kc));
          const bool ht = a.has_tab != 0;
          const uint32_t cb = kSplit ? (bars + 8u * (BB_CONS)) : 0u;
          if (half == 0) {
            dkv_quarter<W, 0, BF16, RW>(pp, pd, saddr, paddr, a.scale_log2, ht, tb, use, ls, dl, 0u);
            dkv_quarter<W, 16, BF16, RW>(pp + 8, pd + 8, saddr, paddr, a.scale_log2, ht, tb, use, ls, dl, cb);
          } else {
            dkv_quarter<W, 32, BF16, RW>(pp, pd, saddr, paddr, a.scale_log2, ht, tb, use, ls, dl, 0u);
            dkv_quarter<W, 48, BF16, RW>(pp + 8, pd + 8, saddr, paddr, a.scale_log2, ht, tb, use, ls, dl, cb);
          }
        }
        if (kSplit) {
          if (G > 0) { mbar_wait((bars + 8u * (BB_PDONE)), (G - 1) & 1); tc_fence_after(); }   // previous dV / dK MMAs have read P^T / dS^T
        } else {
          asm volatile("bar.sync 1, 256;" ::: "memory");     // all S / dP reads done before the in-place bf16 stores
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
        if (row_ok) store_cols<32, BF16>(out, b, h, tok, geo.D, q4 * 32, ov, f, a.out_f32);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

}  // namespace tc
}  // namespace vil
