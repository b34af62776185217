// tc_probe: known-answer probes for the sm_100a building blocks the tcgen05 attention kernels rely on.
// Run on a B200:  ./tc_probe.bin      (prints PASS/FAIL lines; exit code = number of failed probes)
//
//  P1  TMA 5-D tiled load of one w x w chunk of K out of a strided (B,N,2,H,D) kv buffer, SWIZZLE_64B (D=32)
//      and SWIZZLE_128B (D=64), including the out-of-image (zero-filled) part of an edge chunk; the raw shared
//      memory image is compared with the expected XOR-swizzled layout.
//  P2  tcgen05.mma SS, M=128, N=64, both operands K-major from those TMA tiles: S = [Q_A;Q_B] K^T.
//  P3  tcgen05.mma TS: A = bf16 P written to TMEM with tcgen05.st, B = V tile MN-major:  O = P V.
//  P4  same product with P in shared memory (K-major, software-swizzled) instead of TMEM.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include "../vil_sm100.cuh"

using namespace sm100;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(99); } } while (0)

struct ProbeArgs {
  int D;            // head dim (32 or 64); row bytes = 2*D
  int w;            // chunk size
  int layout;       // SWZ_64B or SWZ_128B
  int kR, kC;       // key chunk (chunk-row, chunk-col)
  int qR, qC;       // query chunk A; chunk B = (qR, qC+1)
  int p_in_smem;    // P4 variant
  int lbo, sbo;     // bytes, for all operand descriptors
  float pscale;
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                    const __grid_constant__ CUtensorMap tmK,
                                                    const __grid_constant__ CUtensorMap tmV, ProbeArgs a,
                                                    float* __restrict__ outS, float* __restrict__ outO,
                                                    unsigned char* __restrict__ dumpK) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int rowb = a.D * 2;
  unsigned char* sQ = base;                    // 128 rows
  unsigned char* sK = sQ + 128 * 128;          // 64 rows (region sized for 128-byte rows)
  unsigned char* sV = sK + 64 * 128;
  unsigned char* sP = sV + 64 * 128;           // 128 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 128 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < (128 * 128 * 2 + 64 * 128 * 2) / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0;
  __syncthreads();
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const uint32_t box_bytes = a.w * a.w * rowb;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 4 * box_bytes);
    tma_load_5d(sQ, &tmQ, &bars[0], 0, a.qC * a.w, a.qR * a.w, 1, 0);
    tma_load_5d(sQ + 64 * rowb, &tmQ, &bars[0], 0, (a.qC + 1) * a.w, a.qR * a.w, 1, 0);
    tma_load_5d(sK, &tmK, &bars[0], 0, a.kC * a.w, a.kR * a.w, 1, 0);
    tma_load_5d(sV, &tmV, &bars[0], 0, a.kC * a.w, a.kR * a.w, 1, 0);
  }
  mbar_wait(&bars[0], 0);
  for (int i = tid; i < 64 * 128; i += 128) dumpK[i] = sK[i];

  // ---- S = Q K^T   (M=128, N=64, K=D)
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, 64, true, false, false);
    for (int k = 0; k < a.D / 16; ++k) {
      const uint64_t da = make_smem_desc(smem_u32(sQ) + k * 32, a.lbo, a.sbo, a.layout);
      const uint64_t db = make_smem_desc(smem_u32(sK) + k * 32, a.lbo, a.sbo, a.layout);
      mma_ss(tmem, da, db, idesc, k > 0);
    }
    mma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t s0[32], s1[32];
  tmem_ld_x32(lane_addr, s0);
  tmem_ld_x32(lane_addr + 32, s1);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) { outS[tid * 64 + j] = __uint_as_float(s0[j]); outS[tid * 64 + 32 + j] = __uint_as_float(s1[j]); }

  // ---- P = bf16(S * pscale), keys >= w*w zeroed; into TMEM columns [64, 96) or into shared memory
  uint32_t pk[32];
  for (int j = 0; j < 32; ++j) {
    float x0 = (2 * j < a.w * a.w) ? __uint_as_float(j < 16 ? s0[2 * j] : s1[2 * j - 32]) * a.pscale : 0.f;
    float x1 = (2 * j + 1 < a.w * a.w) ? __uint_as_float(j < 16 ? s0[2 * j + 1] : s1[2 * j + 1 - 32]) * a.pscale : 0.f;
    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    pk[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  if (!a.p_in_smem) {
    tmem_st_x32(lane_addr + 64, pk);
    tmem_st_wait();
  } else {
    // K-major, 128-byte rows (64 keys), SWIZZLE_128B: 16-byte chunk index ^= (row & 7)
    for (int c16 = 0; c16 < 8; ++c16) {
      uint4 val = make_uint4(pk[c16 * 4], pk[c16 * 4 + 1], pk[c16 * 4 + 2], pk[c16 * 4 + 3]);
      *reinterpret_cast<uint4*>(sP + tid * 128 + ((c16 ^ (tid & 7)) * 16)) = val;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- O = P V   (M=128, N=D, K=64 keys), V tile is MN-major (D contiguous)
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, a.D, true, false, true);
    for (int k = 0; k < 4; ++k) {
      const uint64_t dv = make_smem_desc(smem_u32(sV) + k * 16 * rowb, a.lbo, a.sbo, a.layout);
      if (!a.p_in_smem) {
        mma_ts(tmem + 128, tmem + 64 + k * 8, dv, idesc, k > 0);
      } else {
        const uint64_t dp = make_smem_desc(smem_u32(sP) + k * 32, 16, 1024, SWZ_128B);
        mma_ss(tmem + 128, dp, dv, idesc, k > 0);
      }
    }
    mma_commit(&bars[2]);
  }
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  uint32_t o0[32], o1[32];
  tmem_ld_x32(lane_addr + 128, o0);
  if (a.D == 64) tmem_ld_x32(lane_addr + 160, o1);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) {
    outO[tid * 64 + j] = __uint_as_float(o0[j]);
    if (a.D == 64) outO[tid * 64 + 32 + j] = __uint_as_float(o1[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

static int make_map(CUtensorMap* m, void* base, int D, int boxD, int ny, int nx, int H, int B, long long st, long long sh,
                    long long sb, int w, CUtensorMapSwizzle swz) {
  cuuint64_t dims[5] = {(cuuint64_t)D, (cuuint64_t)ny, (cuuint64_t)nx, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)st * 2, (cuuint64_t)ny * st * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[5] = {(cuuint32_t)boxD, (cuuint32_t)w, (cuuint32_t)w, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = get_encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  return 0;
}

struct Case { int D; int p_in_smem; int lbo; int sbo; const char* name; };

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  if (get_encode_tiled() == nullptr) { printf("FAIL: cuTensorMapEncodeTiled entry point not found\n"); return 50; }
  int failures = 0;
  const int w = 7, nx = 10, ny = 12, H = 2, B = 1, g = 1;
  const int N = g + nx * ny;
  std::vector<Case> cases = {
      {32, 0, 512, 512, "D=32 SW64  P in TMEM  lbo=sbo=8 rows"},
      {64, 0, 1024, 1024, "D=64 SW128 P in TMEM  lbo=sbo=8 rows"},
      {32, 1, 512, 512, "D=32 SW64  P in SMEM"},
      {64, 1, 1024, 1024, "D=64 SW128 P in SMEM"},
      {64, 0, 16, 1024, "D=64 SW128 P in TMEM  lbo=16 sbo=1024"},
      {32, 0, 16, 512, "D=32 SW64  P in TMEM  lbo=16 sbo=512"},
  };
  for (const Case& cs : cases) {
    const int D = cs.D, rowb = 2 * D;
    // q: (B, Nloc, H, D) layout; kv: (B, N, 2, H, D) layout
    const long long q_st = (long long)H * D, q_sh = D, q_sb = (long long)nx * ny * H * D;
    const long long kv_st = 2LL * H * D, kv_sh = D, kv_sb = (long long)N * 2 * H * D;
    std::vector<__nv_bfloat16> hq((size_t)B * nx * ny * H * D), hkv((size_t)B * N * 2 * H * D);
    srand(1234 + D);
    for (auto& x : hq) x = __float2bfloat16((float)((rand() % 9) - 4) * 0.25f);
    for (auto& x : hkv) x = __float2bfloat16((float)((rand() % 9) - 4) * 0.25f);
    __nv_bfloat16 *dq, *dkv; float *dS, *dO; unsigned char* dDump;
    CK(cudaMalloc(&dq, hq.size() * 2)); CK(cudaMalloc(&dkv, hkv.size() * 2));
    CK(cudaMalloc(&dS, 128 * 64 * 4)); CK(cudaMalloc(&dO, 128 * 64 * 4)); CK(cudaMalloc(&dDump, 64 * 128));
    CK(cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dkv, hkv.data(), hkv.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dS, 0, 128 * 64 * 4)); CK(cudaMemset(dO, 0, 128 * 64 * 4));
    const CUtensorMapSwizzle swz = D == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUtensorMap tmQ, tmK, tmV;
    int bad = make_map(&tmQ, dq, D, D, ny, nx, H, B, q_st, q_sh, q_sb, w, swz);
    bad |= make_map(&tmK, dkv + g * kv_st, D, D, ny, nx, H, B, kv_st, kv_sh, kv_sb, w, swz);
    bad |= make_map(&tmV, dkv + g * kv_st + H * D, D, D, ny, nx, H, B, kv_st, kv_sh, kv_sb, w, swz);
    if (bad) { failures++; continue; }
    ProbeArgs a;
    a.D = D; a.w = w; a.layout = D == 32 ? SWZ_64B : SWZ_128B;
    a.kR = 1; a.kC = 1;          // edge chunk: rows 7..13 (10..13 out of image), cols 7..13 (12,13 out of image)
    a.qR = 0; a.qC = 0;
    a.p_in_smem = cs.p_in_smem; a.lbo = cs.lbo; a.sbo = cs.sbo; a.pscale = 0.125f;
    const size_t smem = 1024 + 128 * 128 * 2 + 64 * 128 * 2 + 64;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(tmQ, tmK, tmV, a, dS, dO, dDump);
    cudaError_t err = cudaDeviceSynchronize();
    printf("---- case: %s\n", cs.name);
    if (err != cudaSuccess) {
      printf("FAIL kernel error: %s\n", cudaGetErrorString(err));
      failures++;
      return failures + 100;     // context is dead after a trap
    }
    std::vector<float> S(128 * 64), O(128 * 64);
    std::vector<unsigned char> dump(64 * 128);
    CK(cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dump.data(), dDump, dump.size(), cudaMemcpyDeviceToHost));
    const int h = 1, b = 0;
    auto kval = [&](int kr, int kc, int c, int which) -> float {      // which: 0 = k, 1 = v; zero outside the image
      const int r = a.kR * w + kr, cc = a.kC * w + kc;
      if (r >= nx || cc >= ny) return 0.f;
      const long long tok = g + (long long)r * ny + cc;
      return __bfloat162float(hkv[b * kv_sb + tok * kv_st + which * H * D + h * kv_sh + c]);
    };
    auto qval = [&](int slot, int l, int c) -> float {
      const int r = a.qR * w + l / w, cc = (a.qC + slot) * w + l % w;
      if (r >= nx || cc >= ny) return 0.f;
      return __bfloat162float(hq[b * q_sb + ((long long)r * ny + cc) * q_st + h * q_sh + c]);
    };
    // P1: raw smem image of the K tile
    {
      int mism = 0, first = -1;
      const int xmask = D == 32 ? 3 : 7;
      for (int j = 0; j < w * w; ++j)
        for (int c = 0; c < D; ++c) {
          const int lin = j * rowb + c * 2;
          const int chunk = (lin >> 4), rowgrp = (lin >> 7) & xmask;
          const int phys = ((chunk ^ rowgrp) << 4) | (lin & 15);
          __nv_bfloat16 got; memcpy(&got, &dump[phys], 2);
          if (__bfloat162float(got) != kval(j / w, j % w, c, 0)) { if (first < 0) first = j * D + c; mism++; }
        }
      printf("%s P1 TMA 5-D swizzled chunk load + OOB zero fill: %d mismatches of %d (first at %d)\n", mism ? "FAIL" : "PASS",
             mism, w * w * D, first);
      failures += mism != 0;
    }
    // P2: S
    std::vector<float> Sref(128 * 64, 0.f);
    {
      double maxerr = 0; int nbad = 0;
      for (int slot = 0; slot < 2; ++slot)
        for (int l = 0; l < w * w; ++l)
          for (int j = 0; j < w * w; ++j) {
            float acc = 0.f;
            for (int c = 0; c < D; ++c) acc += qval(slot, l, c) * kval(j / w, j % w, c, 0);
            Sref[(slot * 64 + l) * 64 + j] = acc;
            const double e = fabs(acc - S[(slot * 64 + l) * 64 + j]);
            if (e > maxerr) maxerr = e;
            if (e > 1e-3) nbad++;
          }
      printf("%s P2 tcgen05.mma SS K-major S=QK^T: max |err| = %.4g, %d bad of %d  (S[0][0]=%g ref %g, S[64][3]=%g ref %g)\n",
             nbad ? "FAIL" : "PASS", maxerr, nbad, 2 * w * w * w * w, S[0], Sref[0], S[64 * 64 + 3], Sref[64 * 64 + 3]);
      failures += nbad != 0;
    }
    // P3/P4: O = P V with P = bf16(S_device * pscale) (uses the device S so that P2 failures do not cascade)
    {
      double maxerr = 0; int nbad = 0;
      for (int slot = 0; slot < 2; ++slot)
        for (int l = 0; l < w * w; ++l)
          for (int c = 0; c < D; ++c) {
            float acc = 0.f;
            for (int j = 0; j < w * w; ++j) acc += bf(S[(slot * 64 + l) * 64 + j] * a.pscale) * kval(j / w, j % w, c, 1);
            const double e = fabs(acc - O[(slot * 64 + l) * 64 + c]);
            if (e > maxerr) maxerr = e;
            if (e > 2e-2) nbad++;
          }
      printf("%s %s O=PV (V MN-major, P %s): max |err| = %.4g, %d bad of %d  (O[0][0]=%g O[70][5]=%g)\n",
             nbad ? "FAIL" : "PASS", cs.p_in_smem ? "P4" : "P3", cs.p_in_smem ? "from SMEM" : "from TMEM", maxerr, nbad,
             2 * w * w * D, O[0], O[70 * 64 + 5]);
      failures += nbad != 0;
    }
    cudaFree(dq); cudaFree(dkv); cudaFree(dS); cudaFree(dO); cudaFree(dDump);
  }
  printf("tc_probe: %d failed probes\n", failures);
  return failures;
}
