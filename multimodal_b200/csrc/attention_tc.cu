// tcgen05 attention for S <= 256, head_dim 64 (CLIP ViT-B/16 image S=197, text S=77): scores and gradients are
// computed by 5th-gen tensor cores with TMEM accumulators; operands are staged by TMA straight out of the packed
// in-projection output [B*S, 3d] (no head split/merge copies).  One thread per tile row does the softmax math on
// TMEM rows (tcgen05.ld 32x32b), two warpgroups split the columns.
//
//   fwd   (b,h,q-tile):  S = Q K^T (TMEM, <=256 cols) -> softmax -> P (bf16, smem, K-major) -> O = P V (TMEM)
//   dq    (b,h,q-tile):  per 64-wide kv chunk c: S_c = Q K_c^T, dP_c = dO V_c^T -> dS_c -> dQ += dS_c K_c
//   dkdv  (b,h,kv-tile): per 64-wide q chunk c: S^T_c = K Q_c^T, dP^T_c = V dO_c^T -> P^T_c, dS^T_c ->
//                        dV += P^T_c dO_c ; dK += dS^T_c Q_c
// The chunked backward kernels double-buffer the S/dP accumulators in TMEM so the MMAs of chunk c+1 run under the
// elementwise work of chunk c.  Replaces F.scaled_dot_product_attention + autograd (torch/nn/functional.py:6682).
#include "common.cuh"
#include "mmb200_internal.h"
#include <stdlib.h>

namespace mmb {

#ifdef MMB_ATTN_TRACE
// Debug build only (scripts/attn_trace.py): per-phase SM-clock stamps of one CTA per (kernel, tile).
__device__ unsigned long long* g_attn_trace = nullptr;
#define ATRACE(kind, tile, role, slot)                                                                           \
  do {                                                                                                           \
    if (g_attn_trace && blockIdx.z == gridDim.z / 2 && blockIdx.y == 3)                                          \
      g_attn_trace[(((kind) * 2 + (tile)) * 2 + (role)) * 64 + (slot)] = clock64();                             \
  } while (0)
#define PTRACE(kind, role, slot)                                                                                 \
  do {                                                                                                           \
    if (g_attn_trace && blockIdx.x == 1 && n == 6) g_attn_trace[((kind) * 4 + (role)) * 64 + (slot)] = clock64(); \
  } while (0)
#else
#define ATRACE(kind, tile, role, slot) do {} while (0)
#define PTRACE(kind, role, slot) do {} while (0)
#endif

constexpr int ATT_THREADS = 256;
constexpr int ATOM = 16384;  // 128 rows x 128 B

struct AttnTcArgs {
  int S, H, S_pad;
  float scale, scale_log2;
  float* lse;                   // [B,H,S] natural log
  __nv_bfloat16* out;           // fwd: O [B*S, d]
  const __nv_bfloat16* o_in;    // bwd: O
  const __nv_bfloat16* dout;    // bwd: dO [B*S, d]
  __nv_bfloat16* dqkv;          // bwd: [B*S, 3d]
  float* dsum;                  // bwd: rowsum(dO * O) [B,H,S], written by the dQ kernel, read by the dK/dV kernel
  const uint8_t* kmask;         // fwd: optional key-padding mask [B,S], 1 = attend (utils/attention.py:13-53)
};

__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) { return make_smem_desc_sw128(saddr, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr) { return make_smem_desc_sw128(saddr, 8192, 1024); }
__device__ __forceinline__ uint32_t idesc_rt(int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// store 16 consecutive bf16 (two 16 B pieces) of row r at column j0 (multiple of 16) of a K-major SW128 atom series
__device__ __forceinline__ void store_p16(uint8_t* base, int r, int j0, const float (&p)[16]) {
  uint8_t* a = base + (j0 >> 6) * ATOM + r * 128;
  const int c8 = (j0 & 63) >> 3;
  uint4 u0, u1;
  u0.x = pack_bf16x2(p[0], p[1]); u0.y = pack_bf16x2(p[2], p[3]); u0.z = pack_bf16x2(p[4], p[5]); u0.w = pack_bf16x2(p[6], p[7]);
  u1.x = pack_bf16x2(p[8], p[9]); u1.y = pack_bf16x2(p[10], p[11]); u1.z = pack_bf16x2(p[12], p[13]); u1.w = pack_bf16x2(p[14], p[15]);
  *reinterpret_cast<uint4*>(a + ((c8 ^ (r & 7)) << 4)) = u0;
  *reinterpret_cast<uint4*>(a + (((c8 + 1) ^ (r & 7)) << 4)) = u1;
}

// ------------------------------------------------------------------------------------------------
// Forward
// ------------------------------------------------------------------------------------------------
template <bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tmPad,
                   const AttnTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment for the 128B-swizzle atoms; pointer arithmetic on the __shared__ array keeps the address
  // space known to the compiler (LDS/STS instead of generic LD/ST).
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;               // 16 KB   } aliased by sP (4 atoms) once S = Q K^T has completed
  uint8_t* sK = smem + ATOM;        // <=32 KB }
  uint8_t* sP = smem;
  uint8_t* sV = smem + 4 * ATOM;    // <=32 KB
  float* sRed = reinterpret_cast<float*>(smem + 6 * ATOM);  // [2][128] max, [2][128] sum
  uint8_t* sMask = reinterpret_cast<uint8_t*>(sRed + 512);   // [256] key mask of this batch row (1 = attend)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int S = p.S, S_pad = p.S_pad, d = p.H * 64;
  const uint32_t ncols = S_pad <= 128 ? 128u : 256u;
  const bool has_mask = p.kmask != nullptr;
  sMask[threadIdx.x] = (threadIdx.x < S && (!has_mask || p.kmask[(long long)blockIdx.z * S + threadIdx.x])) ? 1 : 0;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm128);
    tma_prefetch_desc(&tmPad);
    mbar_init(&bars[0], 1);  // Q,K landed
    mbar_init(&bars[1], 1);  // V landed
    mbar_init(&bars[2], 1);  // MMA done (phase 0: S, phase 1: O)
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int row0 = b * S;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], ATOM + S_pad * 128);
    tma_load_2d(&tm128, &bars[0], sQ, h * 64, row0 + qt * 128);
    tma_load_2d(&tmPad, &bars[0], sK, d + h * 64, row0);
    mbar_arrive_expect_tx(&bars[1], S_pad * 128);
    tma_load_2d(&tmPad, &bars[1], sV, 2 * d + h * 64, row0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t id = idesc_rt(S_pad, false, false);
    const uint64_t da = desc_k(smem_u32(sQ)), db = desc_k(smem_u32(sK));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, id, k > 0);
    umma_commit(&bars[2]);
  }
  mbar_wait(&bars[2], 0);
  tc_fence_after();

  const int q4 = warp & 3, grp = warp >> 2;
  const int r = q4 * 32 + lane;
  const int qi = qt * 128 + r;
  const uint32_t trow = tmem + ((uint32_t)(q4 * 32) << 16);
  const int nchunk = S_pad >> 4;
  const int c_lo = grp == 0 ? 0 : (nchunk + 1) / 2, c_hi = grp == 0 ? (nchunk + 1) / 2 : nchunk;
  const int kv_lim = CAUSAL ? min(S, qi + 1) : S;  // columns >= kv_lim are masked

  float mx = -INFINITY;
  for (int c = c_lo; c < c_hi; ++c) {
    uint32_t v[16];
    tmem_ld16(trow + c * 16, v);
    tmem_ld_wait();
    if (c * 16 + 16 <= kv_lim && !has_mask) {
#pragma unroll
      for (int e = 0; e < 16; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (c * 16 + e < kv_lim && sMask[c * 16 + e]) mx = fmaxf(mx, __uint_as_float(v[e]));
    }
  }
  sRed[grp * 128 + r] = mx;
  __syncthreads();  // also orders: every thread has finished reading Q/K smem?  (MMA done) -> sP may be written
  mx = fmaxf(sRed[r], sRed[128 + r]) * p.scale_log2;
  float sum = 0.f;
  for (int c = c_lo; c < c_hi; ++c) {
    uint32_t v[16];
    tmem_ld16(trow + c * 16, v);
    tmem_ld_wait();
    float pr[16];
    if (c * 16 + 16 <= kv_lim && !has_mask) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        pr[e] = ex2_approx(fmaf(__uint_as_float(v[e]), p.scale_log2, -mx));
        sum += pr[e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float x = (c * 16 + e < kv_lim && sMask[c * 16 + e]) ? ex2_approx(__uint_as_float(v[e]) * p.scale_log2 - mx) : 0.f;
        pr[e] = x;
        sum += x;
      }
    }
    store_p16(sP, r, c * 16, pr);
  }
  sRed[256 + grp * 128 + r] = sum;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  const float l = sRed[256 + r] + sRed[384 + r];

  if (threadIdx.x == 0) {
    tc_fence_after();
    mbar_wait(&bars[1], 0);
    const uint32_t id = idesc_rt(64, false, true);
    const uint32_t uP = smem_u32(sP), uV = smem_u32(sV);
    for (int j = 0; j < nchunk; ++j) {
      const uint64_t da = desc_k(uP + (j >> 2) * ATOM + (j & 3) * 32);
      const uint64_t db = desc_mn(uV + j * 2048);
      umma_bf16(tmem, da, db, id, j > 0);
    }
    umma_commit(&bars[2]);
  }
  mbar_wait(&bars[2], 1);
  tc_fence_after();
  {
    uint32_t v[32];
    tmem_ld32(trow + grp * 32, v);
    tmem_ld_wait();
    if (qi < S) {
      const float inv = 1.f / l;
      __nv_bfloat16* dst = p.out + ((long long)(row0 + qi)) * d + h * 64 + grp * 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]) * inv, __uint_as_float(v[j * 8 + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]) * inv, __uint_as_float(v[j * 8 + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]) * inv, __uint_as_float(v[j * 8 + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]) * inv, __uint_as_float(v[j * 8 + 7]) * inv);
        reinterpret_cast<uint4*>(dst)[j] = o;
      }
      if (grp == 0 && p.lse) p.lse[((long long)b * p.H + h) * S + qi] = (mx + log2f(l)) * 0.6931471805599453f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, ncols);
  }
}

// ------------------------------------------------------------------------------------------------
// Backward.  One skeleton, two roles:
//   DQ   : tile rows = queries  (resident A0 = Q tile, A1 = dO tile; chunk ring streams K_c / V_c;
//          acc0 = dQ += dS_c K_c)                                           -> also stores D = rowsum(dO * O)
//   DKDV : tile rows = keys     (resident A0 = K tile, A1 = V tile;  chunk ring streams Q_c / dO_c;
//          acc0 = dV += P^T_c dO_c, acc1 = dK += dS^T_c Q_c)
// 64-wide chunks of the other sequence dimension are TMA-streamed through a small ring; TMEM holds S_c, dP_c and
// the accumulators (256 columns) and shared memory stays under 100 KB, so TWO CTAs are resident per SM and one
// CTA's softmax arithmetic overlaps the other's MMAs / loads.
// TMEM columns: S @0 ; dP @64 ; acc0 @128 ; acc1 @192.
// ------------------------------------------------------------------------------------------------
// NG column groups: thread == (tile row) x (column group); 4*NG worker warps + 1 issuer warp.  NG = 4 (16 worker
// warps, 16 columns per thread per chunk) doubles the resident warps per SM but measured 4 % SLOWER than NG = 2: the
// per-CTA critical path is the single-thread MMA issue + the TMA prologue (scripts/attn_trace.py), not warp count.
template <bool CAUSAL, bool DKDV, int NG>
__global__ void __launch_bounds__(NG * 128 + 32, 2)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                   const __grid_constant__ CUtensorMap tmDO128, const __grid_constant__ CUtensorMap tmDO64,
                   const AttnTcArgs p) {
  constexpr int RING = DKDV ? 2 : 3;
  constexpr int CH = 8192;  // one 64-row x 128 B chunk operand
  constexpr int NW = 4 * NG;      // worker warps; the issuer is warp NW
  constexpr int CW = 64 / NG;     // columns of a chunk per worker thread
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment for the 128B-swizzle atoms; pointer arithmetic on the __shared__ array keeps the address
  // space known to the compiler (LDS/STS instead of generic LD/ST).
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA0 = smem;                        // 16 KB  tile operand 0 (Q | K_j)
  uint8_t* sA1 = smem + ATOM;                 // 16 KB  tile operand 1 (dO | V_j)
  uint8_t* sRing = smem + 2 * ATOM;           // RING x (B0_c 8 KB | B1_c 8 KB)
  uint8_t* sDS = sRing + RING * 2 * CH;       // 16 KB  dS chunk (K-major A operand)
  uint8_t* sPT = sDS + ATOM;                  // 16 KB  P^T chunk (DKDV only)
  float* sL = reinterpret_cast<float*>(sDS + (DKDV ? 2 : 1) * ATOM);  // [256] lse (log2 units) per q (DKDV only)
  float* sD = sL + 256;                                                // [256] rowsum(dO*O) per q (DKDV only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 256);
  uint64_t* bar_tile = bars;        // [1] A0/A1 landed
  uint64_t* bar_ld = bars + 1;      // [RING] chunk operands landed
  uint64_t* bar_s = bars + 4;       // S_c/dP_c ready                       issuer -> workers (parity c & 1)
  uint64_t* bar_acc = bars + 5;     // accumulate-MMAs of chunk c complete   issuer -> workers + issuer
  uint64_t* bar_done = bars + 6;    // all MMAs complete
  uint64_t* bar_rd = bars + 7;      // workers finished READING S_c/dP_c from TMEM (8 warp arrivals) -> issuer
  uint64_t* bar_st = bars + 8;      // workers finished WRITING dS_c / P^T_c to smem (8 warp arrivals) -> issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int S = p.S, S_pad = p.S_pad, d = p.H * 64;
  const int row0 = b * S;
  const int nc = (S_pad + 63) >> 6;
  if (threadIdx.x == 0) ATRACE(DKDV, tile, 0, 0);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV128); tma_prefetch_desc(&tmQKV64);
    tma_prefetch_desc(&tmDO128); tma_prefetch_desc(&tmDO64);
    mbar_init(bar_tile, 1);
    for (int i = 0; i < RING; ++i) mbar_init(&bar_ld[i], 1);
    mbar_init(bar_s, 1); mbar_init(bar_acc, 1); mbar_init(bar_done, 1);
    mbar_init(bar_rd, NW); mbar_init(bar_st, NW);
    fence_mbar_init();
  }
  if (warp == NW) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) ATRACE(DKDV, tile, 0, 1);

  if (warp == NW) {
    // ======================= issuer warp: TMA loads + every tcgen05.mma =======================
    if (lane == 0) {
      const uint32_t uA0 = smem_u32(sA0), uA1 = smem_u32(sA1), uRing = smem_u32(sRing);
      const uint32_t uDS = smem_u32(sDS), uPT = smem_u32(sPT);
      auto load_chunk = [&](int c) {
        const int st = c % RING;
        uint8_t* dst = sRing + st * 2 * CH;
        mbar_arrive_expect_tx(&bar_ld[st], 2 * CH);
        if (!DKDV) {
          tma_load_2d(&tmQKV64, &bar_ld[st], dst, d + h * 64, row0 + c * 64);          // K_c
          tma_load_2d(&tmQKV64, &bar_ld[st], dst + CH, 2 * d + h * 64, row0 + c * 64);  // V_c
        } else {
          tma_load_2d(&tmQKV64, &bar_ld[st], dst, h * 64, row0 + c * 64);               // Q_c
          tma_load_2d(&tmDO64, &bar_ld[st], dst + CH, h * 64, row0 + c * 64);            // dO_c
        }
      };
      auto issue_scores = [&](int c) {  // S_c / dP_c (or their transposes)
        const int wc = min(64, S_pad - c * 64);
        const uint32_t id = idesc_rt(wc, false, false);
        const uint32_t ub = uRing + (c % RING) * 2 * CH;
        const uint64_t a0 = desc_k(uA0), a1 = desc_k(uA1), b0 = desc_k(ub), b1 = desc_k(ub + CH);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, a0 + 2 * k, b0 + 2 * k, id, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + 64, a1 + 2 * k, b1 + 2 * k, id, k > 0);
        umma_commit(bar_s);
      };
      mbar_arrive_expect_tx(bar_tile, 2 * ATOM);
      if (!DKDV) {
        tma_load_2d(&tmQKV128, bar_tile, sA0, h * 64, row0 + tile * 128);           // Q tile
        tma_load_2d(&tmDO128, bar_tile, sA1, h * 64, row0 + tile * 128);            // dO tile
      } else {
        tma_load_2d(&tmQKV128, bar_tile, sA0, d + h * 64, row0 + tile * 128);       // K tile
        tma_load_2d(&tmQKV128, bar_tile, sA1, 2 * d + h * 64, row0 + tile * 128);   // V tile
      }
      for (int c = 0; c < RING && c < nc; ++c) load_chunk(c);
      mbar_wait(bar_tile, 0);
      mbar_wait(&bar_ld[0], 0);
      tc_fence_after();
      ATRACE(DKDV, tile, 1, 0);
      issue_scores(0);
      ATRACE(DKDV, tile, 1, 1);
      for (int c = 0; c < nc; ++c) {
        const int wc = min(64, S_pad - c * 64);
        if (c >= 1) {  // ring stage (c-1) % RING was last read by the accumulate-MMAs of chunk c-1
          mbar_wait(bar_acc, (c - 1) & 1);
          if (c + RING - 1 < nc) load_chunk(c + RING - 1);
        }
        if (c + 1 < nc) {  // S/dP TMEM columns are free as soon as every worker has read chunk c
          mbar_wait(bar_rd, c & 1);
          mbar_wait(&bar_ld[(c + 1) % RING], ((c + 1) / RING) & 1);
          tc_fence_after();
          ATRACE(DKDV, tile, 1, 8 + c * 4 + 0);
          issue_scores(c + 1);  // runs under the workers' exp / FMA work on chunk c
          ATRACE(DKDV, tile, 1, 8 + c * 4 + 1);
        }
        mbar_wait(bar_st, c & 1);  // dS_c (and P^T_c) are in shared memory
        tc_fence_after();
        ATRACE(DKDV, tile, 1, 8 + c * 4 + 2);
        const uint32_t id = idesc_rt(64, false, true);
        const uint32_t ub = uRing + (c % RING) * 2 * CH;
        const int ks = wc >> 4;
        if (!DKDV) {   // dQ += dS_c K_c
          for (int k = 0; k < ks; ++k)
            umma_bf16(tmem + 128, desc_k(uDS + k * 32), desc_mn(ub + k * 2048), id, (c > 0 || k > 0));
        } else {       // dV += P^T_c dO_c ; dK += dS^T_c Q_c
          for (int k = 0; k < ks; ++k)
            umma_bf16(tmem + 128, desc_k(uPT + k * 32), desc_mn(ub + CH + k * 2048), id, (c > 0 || k > 0));
          for (int k = 0; k < ks; ++k)
            umma_bf16(tmem + 192, desc_k(uDS + k * 32), desc_mn(ub + k * 2048), id, (c > 0 || k > 0));
        }
        umma_commit(bar_acc);
        ATRACE(DKDV, tile, 1, 8 + c * 4 + 3);
        if (c == nc - 1) umma_commit(bar_done);
      }
    }
  } else {
    // ======================= 4*NG worker warps =======================
    const int q4 = warp & 3, grp = warp >> 2;
    const int r = q4 * 32 + lane;
    const int ri = tile * 128 + r;  // global index of this thread's row (query for DQ, key for DKDV)
    const uint32_t trow = tmem + ((uint32_t)(q4 * 32) << 16);

    // softmax statistics: per row (DQ; also published for the DKDV kernel) or per column via smem (DKDV)
    float Lrow = 0.f, Drow = 0.f;
    if (!DKDV) {
      // each column group sums its 64/NG head dims of dO * O; the partials meet in shared memory
      float acc = 0.f;
      if (ri < S) {
        const uint4* po = reinterpret_cast<const uint4*>(p.o_in + (long long)(row0 + ri) * d + h * 64) + grp * (8 / NG);
        const uint4* pd = reinterpret_cast<const uint4*>(p.dout + (long long)(row0 + ri) * d + h * 64) + grp * (8 / NG);
#pragma unroll
        for (int j = 0; j < 8 / NG; ++j) {
          const uint4 a = __ldg(po + j), c = __ldg(pd + j);
          acc += bf16_lo(a.x) * bf16_lo(c.x) + bf16_hi(a.x) * bf16_hi(c.x) + bf16_lo(a.y) * bf16_lo(c.y) +
                 bf16_hi(a.y) * bf16_hi(c.y) + bf16_lo(a.z) * bf16_lo(c.z) + bf16_hi(a.z) * bf16_hi(c.z) +
                 bf16_lo(a.w) * bf16_lo(c.w) + bf16_hi(a.w) * bf16_hi(c.w);
        }
        Lrow = p.lse[((long long)b * p.H + h) * S + ri] * 1.4426950408889634f;
      }
      sL[grp * 128 + r] = acc;   // sL/sD: 512 floats
      asm volatile("bar.sync 1, %0;" ::"n"(NG * 128) : "memory");
#pragma unroll
      for (int gI = 0; gI < NG; ++gI) Drow += sL[gI * 128 + r];
      if (grp == 0 && ri < S) p.dsum[((long long)b * p.H + h) * S + ri] = Drow;
    } else {
      const int qi = threadIdx.x;  // the first 256 worker threads cover S_pad <= 256 query columns
      if (qi < 256) {
        const bool ok = qi < S;
        sL[qi] = ok ? p.lse[((long long)b * p.H + h) * S + qi] * 1.4426950408889634f : 0.f;
        sD[qi] = ok ? p.dsum[((long long)b * p.H + h) * S + qi] : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NG * 128) : "memory");
    }

    if (threadIdx.x == 0) ATRACE(DKDV, tile, 0, 2);
    for (int c = 0; c < nc; ++c) {
      const int wc = min(64, S_pad - c * 64);
      mbar_wait(bar_s, c & 1);
      tc_fence_after();
      if (threadIdx.x == 0) ATRACE(DKDV, tile, 0, 8 + c * 4 + 0);
      // this thread: row r, columns [grp*CW, grp*CW+CW) of the chunk
      uint32_t sv[CW], dv[CW];
      if (CW == 32) {
        tmem_ld32(trow + grp * 32, reinterpret_cast<uint32_t(&)[32]>(sv));
        tmem_ld32(trow + 64 + grp * 32, reinterpret_cast<uint32_t(&)[32]>(dv));
      } else {
        tmem_ld16(trow + grp * 16, reinterpret_cast<uint32_t(&)[16]>(sv));
        tmem_ld16(trow + 64 + grp * 16, reinterpret_cast<uint32_t(&)[16]>(dv));
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_rd);       // TMEM S/dP of this chunk consumed -> next chunk's MMAs may overwrite
      if (threadIdx.x == 0) ATRACE(DKDV, tile, 0, 8 + c * 4 + 1);
      if (c >= 1) mbar_wait(bar_acc, (c - 1) & 1);  // dS / P^T buffers free (normally long since complete)
      if (threadIdx.x == 0) ATRACE(DKDV, tile, 0, 8 + c * 4 + 2);
      {
        const int cbase = c * 64 + grp * CW;  // first global column (key for DQ, query for DKDV) of this thread's CW
        // interior fast path: the whole CW-column strip is unmasked for this row (all but the last chunk / the causal
        // diagonal / padding rows) -> no per-element predicates or index arithmetic
        const bool full = (ri < S) && (cbase + CW <= S) && (!CAUSAL || (DKDV ? (ri <= cbase) : (cbase + CW - 1 <= ri)));
        const float nD = -Drow * p.scale;
#pragma unroll
        for (int half = 0; half < CW / 16; ++half) {
          float ds[16], pt[16];
          if (ri >= S) {  // padding row of the last tile: contributes nothing
#pragma unroll
            for (int e = 0; e < 16; ++e) { ds[e] = 0.f; pt[e] = 0.f; }
          } else if (full) {
            if (!DKDV) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float pv = ex2_approx(fmaf(__uint_as_float(sv[half * 16 + e]), p.scale_log2, -Lrow));
                ds[e] = pv * fmaf(__uint_as_float(dv[half * 16 + e]), p.scale, nD);
              }
            } else {
              const float4* pl = reinterpret_cast<const float4*>(sL + cbase + half * 16);
              const float4* pd = reinterpret_cast<const float4*>(sD + cbase + half * 16);
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const float4 l4 = pl[e4], d4 = pd[e4];
                const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int e = e4 * 4 + k;
                  const float pv = ex2_approx(fmaf(__uint_as_float(sv[half * 16 + e]), p.scale_log2, -ls[k]));
                  pt[e] = pv;
                  ds[e] = pv * (__uint_as_float(dv[half * 16 + e]) - dd[k]) * p.scale;
                }
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int cj = cbase + half * 16 + e;
              bool valid;
              float L, Dv;
              if (!DKDV) {
                valid = (ri < S) && (cj < S) && (!CAUSAL || cj <= ri);
                L = Lrow; Dv = Drow;
              } else {
                valid = (ri < S) && (cj < S) && (!CAUSAL || ri <= cj);
                L = sL[cj & 255]; Dv = sD[cj & 255];
              }
              const float pv = valid ? ex2_approx(__uint_as_float(sv[half * 16 + e]) * p.scale_log2 - L) : 0.f;
              pt[e] = pv;
              ds[e] = pv * (__uint_as_float(dv[half * 16 + e]) - Dv) * p.scale;
            }
          }
          if (grp * CW + half * 16 < wc) {
            store_p16(sDS, r, grp * CW + half * 16, ds);
            if (DKDV) store_p16(sPT, r, grp * CW + half * 16, pt);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_st);
      if (threadIdx.x == 0) ATRACE(DKDV, tile, 0, 8 + c * 4 + 3);
    }
    mbar_wait(bar_done, 0);
    tc_fence_after();
    if (threadIdx.x == 0) ATRACE(DKDV, tile, 0, 3);
    const long long ld = 3LL * d;
#pragma unroll
    for (int which = 0; which < (DKDV ? 2 : 1); ++which) {
      // DQ: acc0 -> Q block.  DKDV: acc0 = dV -> V block (2d), acc1 = dK -> K block (d)
      uint32_t v[CW];
      if (CW == 32) tmem_ld32(trow + 128 + which * 64 + grp * 32, reinterpret_cast<uint32_t(&)[32]>(v));
      else          tmem_ld16(trow + 128 + which * 64 + grp * 16, reinterpret_cast<uint32_t(&)[16]>(v));
      tmem_ld_wait();
      if (ri < S) {
        const int coff = !DKDV ? 0 : (which == 0 ? 2 * d : d);
        __nv_bfloat16* dst = p.dqkv + (long long)(row0 + ri) * ld + coff + h * 64 + grp * CW;
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
          reinterpret_cast<uint4*>(dst)[j] = o;
        }
      }
    }
  }
  if (threadIdx.x == 0) ATRACE(DKDV, tile, 0, 4);
  tc_fence_before();
  __syncthreads();
  if (warp == NW) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
  if (threadIdx.x == 0) ATRACE(DKDV, tile, 0, 5);
}


// ------------------------------------------------------------------------------------------------
// Backward, persistent version (round-1 final): ONE CTA per SM loops over (batch, head, 128-row tile) work items with
// every resource double-buffered, so nothing on the critical path of the two-CTA kernels above is exposed any more
// (measured with scripts/attn_trace.py: 28 % of a CTA's life was the TMA prologue, 12 % the epilogue, and inside the
// chunk loop the single issuer thread serialised ~16 tcgen05.mma issues of ~80 clocks each against the workers):
//   * a TMA producer warp runs ahead: tile operands [2 buffers] + a ring of 64-row chunk operands -> the next tile's
//     loads fly during the current tile's arithmetic;
//   * TWO issuer threads: one for the score MMAs (S_c, dP_c -> TMEM buffer c & 1), one for the accumulate MMAs
//     (dQ | dV,dK += ... from the dS / P^T smem buffers c & 1) -> MMA issue overlaps itself and the workers;
//   * two statistics warps prepare the per-row LSE and D = rowsum(dO * O) of the NEXT tile in shared memory (global
//     load latency off the workers' path; the dQ pass also publishes D for the dK/dV pass);
//   * S/dP live in two TMEM buffers and dS / P^T in two smem buffers: workers never wait for the MMAs of the chunk
//     they just finished; the accumulators are double-buffered across tiles so the epilogue (TMEM -> bf16 -> global)
//     of tile n runs under the MMAs of tile n+1.
// TMEM (512 columns): [S|dP] buffer i at i*128 (S +0, dP +64); accumulators of tile parity j at 256 + j*128 (+0, +64).
// Barrier phases are derived from running counters (tile sequence n, global chunk sequence g).
// ------------------------------------------------------------------------------------------------
constexpr int P_STATS = 2;   // statistics warps: softmax LSE and D = rowsum(dO*O), one tile ahead
// NG column groups per row: 4*NG worker warps (TMEM lane quadrant = warp & 3, column group = warp >> 2), then the
// producer warp, the score issuer, the accumulate issuer and the statistics warps.
template <bool CAUSAL, bool DKDV, int NG>
__global__ void __launch_bounds__((4 * NG + 3 + P_STATS) * 32, 1)
attn_bwd_persist_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                        const __grid_constant__ CUtensorMap tmDO128, const __grid_constant__ CUtensorMap tmDO64,
                        const AttnTcArgs p, const int n_work) {
  constexpr int RING = DKDV ? 3 : 4;
  constexpr int CH = 8192;  // one 64-row x 128 B chunk operand
  constexpr int P_WORKERS = 4 * NG;
  constexpr int CW = 64 / NG;   // columns of a chunk per worker thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // [2 tile buffers][A0 16 KB | A1 16 KB]
  uint8_t* sRing = sA + 4 * ATOM;                       // RING x (B0_c 8 KB | B1_c 8 KB)
  uint8_t* sDS = sRing + RING * 2 * CH;                 // [2] dS chunk (16 KB each)
  uint8_t* sPT = sDS + 2 * ATOM;                        // [2] P^T chunk (DKDV only)
  float* sLD = reinterpret_cast<float*>(sPT + (DKDV ? 2 : 0) * ATOM);  // [2][512]: lse(log2) | rowsum(dO*O) per q
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLD + 2 * 512);
  uint64_t* tile_full = bars;            // [2]  tile operands landed                 producer(TMA) -> score issuer
  uint64_t* tile_empty = bars + 2;       // [2]  last score MMA of the tile complete  score issuer  -> producer
  uint64_t* ring_full = bars + 4;        // [RING]
  uint64_t* ring_empty = bars + 8;       // [RING] accumulate MMAs of the chunk complete            -> producer
  uint64_t* sdp_full = bars + 12;        // [2]  S_c/dP_c in TMEM                     score issuer  -> workers
  uint64_t* sdp_empty = bars + 14;       // [2]  workers have read them (8 warps)                   -> score issuer
  uint64_t* ds_full = bars + 16;         // [2]  dS_c (P^T_c) in smem (8 warps)       workers       -> acc issuer
  uint64_t* ds_empty = bars + 18;        // [2]  accumulate MMAs done with them       acc issuer    -> workers
  uint64_t* acc_full = bars + 20;        // [2]  accumulators of the tile final       acc issuer    -> workers
  uint64_t* acc_empty = bars + 22;       // [2]  workers have read them (8 warps)                   -> acc issuer
  uint64_t* stat_full = bars + 24;       // [2]  LSE / D of the tile in sLD (2 warps) stats warps   -> workers
  uint64_t* stat_empty = bars + 26;      // [2]  workers are done with them (8 warps)               -> stats warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, S_pad = p.S_pad, d = p.H * 64;
  const int nc = (S_pad + 63) >> 6;
  const int ntile = (S + 127) >> 7;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV128); tma_prefetch_desc(&tmQKV64);
    tma_prefetch_desc(&tmDO128); tma_prefetch_desc(&tmDO64);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tile_full[i], 1); mbar_init(&tile_empty[i], 1);
      mbar_init(&sdp_full[i], 1); mbar_init(&sdp_empty[i], P_WORKERS);
      mbar_init(&ds_full[i], P_WORKERS); mbar_init(&ds_empty[i], 1);
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], P_WORKERS);
      mbar_init(&stat_full[i], P_STATS); mbar_init(&stat_empty[i], P_WORKERS);
    }
    for (int i = 0; i < RING; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == P_WORKERS) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == P_WORKERS) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int n = 0;
      long long g = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
        const int tile = w % ntile, bh = w / ntile;
        const int h = bh % p.H, b = bh / p.H;
        const int row0 = b * S;
        const int tb = n & 1;
        mbar_wait(&tile_empty[tb], ((n >> 1) & 1) ^ 1);
        uint8_t* a0 = sA + tb * 2 * ATOM;
        mbar_arrive_expect_tx(&tile_full[tb], 2 * ATOM);
        if (!DKDV) {
          tma_load_2d(&tmQKV128, &tile_full[tb], a0, h * 64, row0 + tile * 128);               // Q tile
          tma_load_2d(&tmDO128, &tile_full[tb], a0 + ATOM, h * 64, row0 + tile * 128);          // dO tile
        } else {
          tma_load_2d(&tmQKV128, &tile_full[tb], a0, d + h * 64, row0 + tile * 128);           // K tile
          tma_load_2d(&tmQKV128, &tile_full[tb], a0 + ATOM, 2 * d + h * 64, row0 + tile * 128); // V tile
        }
        for (int c = 0; c < nc; ++c, ++g) {
          const int st = (int)(g % RING);
          mbar_wait(&ring_empty[st], (uint32_t)(((g / RING) & 1) ^ 1));
          uint8_t* dst = sRing + st * 2 * CH;
          mbar_arrive_expect_tx(&ring_full[st], 2 * CH);
          if (!DKDV) {
            tma_load_2d(&tmQKV64, &ring_full[st], dst, d + h * 64, row0 + c * 64);           // K_c
            tma_load_2d(&tmQKV64, &ring_full[st], dst + CH, 2 * d + h * 64, row0 + c * 64);   // V_c
          } else {
            tma_load_2d(&tmQKV64, &ring_full[st], dst, h * 64, row0 + c * 64);                // Q_c
            tma_load_2d(&tmDO64, &ring_full[st], dst + CH, h * 64, row0 + c * 64);             // dO_c
          }
        }
      }
    }
  } else if (warp == P_WORKERS + 1) {
    // ======================= score issuer: S_c = A0 B0_c^T, dP_c = A1 B1_c^T =======================
    if (lane == 0) {
      const uint32_t uA = smem_u32(sA), uRing = smem_u32(sRing);
      int n = 0;
      long long g = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
        const int tb = n & 1;
        mbar_wait(&tile_full[tb], (n >> 1) & 1);
        const uint64_t a0 = desc_k(uA + tb * 2 * ATOM), a1 = desc_k(uA + tb * 2 * ATOM + ATOM);
        for (int c = 0; c < nc; ++c, ++g) {
          const int st = (int)(g % RING), sb = (int)(g & 1);
          const int wc = min(64, S_pad - c * 64);
          PTRACE(DKDV, 1, c * 4 + 0);
          mbar_wait(&ring_full[st], (uint32_t)((g / RING) & 1));
          PTRACE(DKDV, 1, c * 4 + 1);
          mbar_wait(&sdp_empty[sb], (uint32_t)(((g >> 1) & 1) ^ 1));
          tc_fence_after();
          PTRACE(DKDV, 1, c * 4 + 2);
          const uint32_t id = idesc_rt(wc, false, false);
          const uint32_t ub = uRing + st * 2 * CH;
          const uint64_t b0 = desc_k(ub), b1 = desc_k(ub + CH);
          const uint32_t tS = tmem + sb * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS, a0 + 2 * k, b0 + 2 * k, id, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS + 64, a1 + 2 * k, b1 + 2 * k, id, k > 0);
          umma_commit(&sdp_full[sb]);
          PTRACE(DKDV, 1, c * 4 + 3);
        }
        umma_commit(&tile_empty[tb]);   // every MMA that reads this tile buffer has been issued before this commit
      }
    }
  } else if (warp == P_WORKERS + 2) {
    // ======================= accumulate issuer: dQ += dS_c K_c  |  dV += P^T_c dO_c ; dK += dS^T_c Q_c ==============
    if (lane == 0) {
      const uint32_t uRing = smem_u32(sRing), uDS = smem_u32(sDS), uPT = smem_u32(sPT);
      const uint32_t id = idesc_rt(64, false, true);
      int n = 0;
      long long g = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
        const int ab = n & 1;
        mbar_wait(&acc_empty[ab], ((n >> 1) & 1) ^ 1);
        const uint32_t tA = tmem + 256 + ab * 128;
        for (int c = 0; c < nc; ++c, ++g) {
          const int st = (int)(g % RING), sb = (int)(g & 1);
          const int wc = min(64, S_pad - c * 64);
          PTRACE(DKDV, 2, c * 4 + 0);
          mbar_wait(&ds_full[sb], (uint32_t)((g >> 1) & 1));
          PTRACE(DKDV, 2, c * 4 + 1);
          mbar_wait(&ring_full[st], (uint32_t)((g / RING) & 1));  // long complete; acquires the TMA writes for this thread
          tc_fence_after();
          const uint32_t ub = uRing + st * 2 * CH;
          const int ks = wc >> 4;
          if (!DKDV) {
            for (int k = 0; k < ks; ++k)
              umma_bf16(tA, desc_k(uDS + sb * ATOM + k * 32), desc_mn(ub + k * 2048), id, (c > 0 || k > 0));
          } else {
            for (int k = 0; k < ks; ++k)
              umma_bf16(tA, desc_k(uPT + sb * ATOM + k * 32), desc_mn(ub + CH + k * 2048), id, (c > 0 || k > 0));
            for (int k = 0; k < ks; ++k)
              umma_bf16(tA + 64, desc_k(uDS + sb * ATOM + k * 32), desc_mn(ub + k * 2048), id, (c > 0 || k > 0));
          }
          umma_commit(&ds_empty[sb]);
          umma_commit(&ring_empty[st]);
          if (c == nc - 1) umma_commit(&acc_full[ab]);
          PTRACE(DKDV, 2, c * 4 + 2);
        }
      }
    }
  } else if (warp >= P_WORKERS + 3) {
    // ======================= statistics warps: sLD[n & 1] = {LSE (log2 units) [256], D [256]} of tile n ==========
    // DQ pass: entries are the tile's 128 query rows (D computed here from dO and O, and published for the dK/dV
    // pass); dK/dV pass: entries are all S (<= 256) queries of the (batch, head).
    const int t = (warp - (P_WORKERS + 3)) * 32 + lane;   // 0 .. 63
    int n = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
      const int tile = w % ntile, bh = w / ntile;
      const int h = bh % p.H, b = bh / p.H;
      const int row0 = b * S;
      float* wL = sLD + (n & 1) * 512;
      const long long sbase = ((long long)b * p.H + h) * S;
      mbar_wait(&stat_empty[n & 1], ((n >> 1) & 1) ^ 1);
      if (!DKDV) {
#pragma unroll
        for (int k = 0; k < 128 / (P_STATS * 32); ++k) {
          const int rr = t + k * (P_STATS * 32);
          const int ri = tile * 128 + rr;
          float acc = 0.f, L = 0.f;
          if (ri < S) {
            const uint4* po = reinterpret_cast<const uint4*>(p.o_in + (long long)(row0 + ri) * d + h * 64);
            const uint4* pd = reinterpret_cast<const uint4*>(p.dout + (long long)(row0 + ri) * d + h * 64);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 a = __ldg(po + j), c = __ldg(pd + j);
              acc += bf16_lo(a.x) * bf16_lo(c.x) + bf16_hi(a.x) * bf16_hi(c.x) + bf16_lo(a.y) * bf16_lo(c.y) +
                     bf16_hi(a.y) * bf16_hi(c.y) + bf16_lo(a.z) * bf16_lo(c.z) + bf16_hi(a.z) * bf16_hi(c.z) +
                     bf16_lo(a.w) * bf16_lo(c.w) + bf16_hi(a.w) * bf16_hi(c.w);
            }
            L = p.lse[sbase + ri] * 1.4426950408889634f;
            p.dsum[sbase + ri] = acc;
          }
          wL[rr] = L;
          wL[256 + rr] = acc;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 256 / (P_STATS * 32); ++k) {
          const int qi = t + k * (P_STATS * 32);
          const bool ok = qi < S;
          wL[qi] = ok ? p.lse[sbase + qi] * 1.4426950408889634f : 0.f;
          wL[256 + qi] = ok ? p.dsum[sbase + qi] : 0.f;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&stat_full[n & 1]);   // release: the smem writes above are visible to the waiters
    }
  } else {
    // ======================= 4*NG worker warps: thread == tile row x column group =======================
    const int q4 = warp & 3, grp = warp >> 2;
    const int r = q4 * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q4 * 32) << 16);
    int n = 0;
    long long g = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
      const int tile = w % ntile, bh = w / ntile;
      const int h = bh % p.H, b = bh / p.H;
      const int row0 = b * S;
      const int ri = tile * 128 + r;  // global index of this thread's row (query for DQ, key for DKDV)
      float Lrow = 0.f, Drow = 0.f;
      const float* sL = sLD + (n & 1) * 512;
      const float* sD = sL + 256;
      if (threadIdx.x == 0) PTRACE(DKDV, 0, 40);
      mbar_wait(&stat_full[n & 1], (n >> 1) & 1);
      if (!DKDV) {
        Lrow = sL[r];
        Drow = sD[r];
      }
      if (threadIdx.x == 0) PTRACE(DKDV, 0, 41);

      for (int c = 0; c < nc; ++c, ++g) {
        const int sb = (int)(g & 1);
        const int wc = min(64, S_pad - c * 64);
        if (threadIdx.x == 0) PTRACE(DKDV, 0, c * 5 + 0);
        mbar_wait(&sdp_full[sb], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        if (threadIdx.x == 0) PTRACE(DKDV, 0, c * 5 + 1);
        uint32_t sv[CW], dv[CW];
        if (CW == 32) {
          tmem_ld32(trow + sb * 128 + grp * 32, reinterpret_cast<uint32_t(&)[32]>(sv));
          tmem_ld32(trow + sb * 128 + 64 + grp * 32, reinterpret_cast<uint32_t(&)[32]>(dv));
        } else {
          tmem_ld16(trow + sb * 128 + grp * 16, reinterpret_cast<uint32_t(&)[16]>(sv));
          tmem_ld16(trow + sb * 128 + 64 + grp * 16, reinterpret_cast<uint32_t(&)[16]>(dv));
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sdp_empty[sb]);
        if (threadIdx.x == 0) PTRACE(DKDV, 0, c * 5 + 2);
        mbar_wait(&ds_empty[sb], (uint32_t)(((g >> 1) & 1) ^ 1));  // accumulate MMAs of chunk g-2 have left the buffer
        if (threadIdx.x == 0) PTRACE(DKDV, 0, c * 5 + 3);
        uint8_t* myDS = sDS + sb * ATOM;
        uint8_t* myPT = sPT + sb * ATOM;
        {
          const int cbase = c * 64 + grp * CW;
          const bool full = (ri < S) && (cbase + CW <= S) && (!CAUSAL || (DKDV ? (ri <= cbase) : (cbase + CW - 1 <= ri)));
          const float nD = -Drow * p.scale;
#pragma unroll
          for (int half = 0; half < CW / 16; ++half) {
            float ds[16], pt[16];
            if (ri >= S) {
#pragma unroll
              for (int e = 0; e < 16; ++e) { ds[e] = 0.f; pt[e] = 0.f; }
            } else if (full) {
              if (!DKDV) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float pv = ex2_approx(fmaf(__uint_as_float(sv[half * 16 + e]), p.scale_log2, -Lrow));
                  ds[e] = pv * fmaf(__uint_as_float(dv[half * 16 + e]), p.scale, nD);
                }
              } else {
                const float4* pl = reinterpret_cast<const float4*>(sL + cbase + half * 16);
                const float4* pd = reinterpret_cast<const float4*>(sD + cbase + half * 16);
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                  const float4 l4 = pl[e4], d4 = pd[e4];
                  const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const int e = e4 * 4 + k;
                    const float pv = ex2_approx(fmaf(__uint_as_float(sv[half * 16 + e]), p.scale_log2, -ls[k]));
                    pt[e] = pv;
                    ds[e] = pv * (__uint_as_float(dv[half * 16 + e]) - dd[k]) * p.scale;
                  }
                }
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int cj = cbase + half * 16 + e;
                bool valid;
                float L, Dv;
                if (!DKDV) {
                  valid = (ri < S) && (cj < S) && (!CAUSAL || cj <= ri);
                  L = Lrow; Dv = Drow;
                } else {
                  valid = (ri < S) && (cj < S) && (!CAUSAL || ri <= cj);
                  L = sL[cj & 255]; Dv = sD[cj & 255];
                }
                const float pv = valid ? ex2_approx(__uint_as_float(sv[half * 16 + e]) * p.scale_log2 - L) : 0.f;
                pt[e] = pv;
                ds[e] = pv * (__uint_as_float(dv[half * 16 + e]) - Dv) * p.scale;
              }
            }
            if (grp * CW + half * 16 < wc) {
              store_p16(myDS, r, grp * CW + half * 16, ds);
              if (DKDV) store_p16(myPT, r, grp * CW + half * 16, pt);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_full[sb]);
        if (threadIdx.x == 0) PTRACE(DKDV, 0, c * 5 + 4);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&stat_empty[n & 1]);   // this warp no longer reads sLD[n & 1]
      // ---- tile epilogue: accumulators (TMEM buffer n & 1) -> bf16 -> dqkv ----
      const int ab = n & 1;
      if (threadIdx.x == 0) PTRACE(DKDV, 0, 42);
      mbar_wait(&acc_full[ab], (n >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 0) PTRACE(DKDV, 0, 43);
      const long long ld = 3LL * d;
      uint32_t v0[CW], v1[CW];
      if (CW == 32) {
        tmem_ld32(trow + 256 + ab * 128 + grp * 32, reinterpret_cast<uint32_t(&)[32]>(v0));
        if (DKDV) tmem_ld32(trow + 256 + ab * 128 + 64 + grp * 32, reinterpret_cast<uint32_t(&)[32]>(v1));
      } else {
        tmem_ld16(trow + 256 + ab * 128 + grp * 16, reinterpret_cast<uint32_t(&)[16]>(v0));
        if (DKDV) tmem_ld16(trow + 256 + ab * 128 + 64 + grp * 16, reinterpret_cast<uint32_t(&)[16]>(v1));
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
      if (ri < S) {
#pragma unroll
        for (int which = 0; which < (DKDV ? 2 : 1); ++which) {
          // DQ: acc0 -> Q block.  DKDV: acc0 = dV -> V block (2d), acc1 = dK -> K block (d)
          const uint32_t* v = which == 0 ? v0 : v1;
          const int coff = !DKDV ? 0 : (which == 0 ? 2 * d : d);
          __nv_bfloat16* dst = p.dqkv + (long long)(row0 + ri) * ld + coff + h * 64 + grp * CW;
#pragma unroll
          for (int j = 0; j < CW / 8; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
            o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
            o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
            o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
            reinterpret_cast<uint4*>(dst)[j] = o;
          }
        }
      }
      if (threadIdx.x == 0) PTRACE(DKDV, 0, 44);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == P_WORKERS) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

constexpr int BWDP_DQ_SMEM = 1024 + 4 * ATOM + 4 * 16384 + 2 * ATOM + 4096 + 512;              // 169.5 KB
constexpr int BWDP_DKDV_SMEM = 1024 + 4 * ATOM + 3 * 16384 + 4 * ATOM + 4096 + 512;            // 185.5 KB

constexpr int FWD_SMEM = 1024 + 6 * ATOM + 2048 + 256 + 64;
constexpr int BWD_DQ_SMEM = 1024 + 2 * ATOM + 3 * 16384 + ATOM + 2048 + 128;        //  99.3 KB -> 2 CTAs / SM
constexpr int BWD_DKDV_SMEM = 1024 + 2 * ATOM + 2 * 16384 + 2 * ATOM + 2048 + 128;  //  99.3 KB -> 2 CTAs / SM

}  // namespace mmb

using namespace mmb;

#ifdef MMB_ATTN_TRACE
extern "C" int mmb_debug_attn_trace(void* p) { return (int)cudaMemcpyToSymbol(g_attn_trace, &p, sizeof(p)); }
#endif

extern "C" int mmb_attention_fwd_tc(const void* qkv, void* out, float* lse, int B, int S, int H, int causal, float scale,
                                    void* stream);
extern "C" int mmb_attention_fwd_kmask(const void* qkv, void* out, float* lse, const unsigned char* kmask, int B, int S,
                                       int H, int head_dim, int causal, float scale, void* stream);
static int attention_fwd_tc_impl(const void* qkv, void* out, float* lse, const uint8_t* kmask, int B, int S, int H,
                                 int causal, float scale, void* stream) {
  if (B <= 0 || S <= 0 || S > 256) return MMB_ERR_UNSUPPORTED;
  const int d = H * 64, S_pad = (S + 15) & ~15;
  CUtensorMap tm128, tmPad;
  int rc = make_tmap_2d(&tm128, qkv, 2, false, 3ull * d, (uint64_t)B * S, 3ull * d * 2, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmPad, qkv, 2, false, 3ull * d, (uint64_t)B * S, 3ull * d * 2, 64, S_pad);
  if (rc) return rc;
  AttnTcArgs a{};
  a.S = S; a.H = H; a.S_pad = S_pad; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  a.lse = lse; a.out = (__nv_bfloat16*)out; a.kmask = kmask;
  dim3 grid((S + 127) / 128, H, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (causal) {
    cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
    attn_fwd_tc_kernel<true><<<grid, ATT_THREADS, FWD_SMEM, st>>>(tm128, tmPad, a);
  } else {
    cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
    attn_fwd_tc_kernel<false><<<grid, ATT_THREADS, FWD_SMEM, st>>>(tm128, tmPad, a);
  }
  return (int)cudaGetLastError();
}

extern "C" int mmb_attention_fwd_tc(const void* qkv, void* out, float* lse, int B, int S, int H, int causal, float scale,
                                    void* stream) {
  return attention_fwd_tc_impl(qkv, out, lse, nullptr, B, S, H, causal, scale, stream);
}
// Same with a key-padding mask [B,S] (1 = attend): BERT-style attention of the FLAVA text tower
// (modules/encoders/bert_text_encoder.py:87-93 -> modules/layers/attention.py:228-229 masked_fill(-inf)).
extern "C" int mmb_attention_fwd_kmask(const void* qkv, void* out, float* lse, const unsigned char* kmask, int B, int S,
                                       int H, int head_dim, int causal, float scale, void* stream) {
  if (head_dim != 64 || S > 256) return MMB_ERR_UNSUPPORTED;
  return attention_fwd_tc_impl(qkv, out, lse, kmask, B, S, H, causal, scale, stream);
}

extern "C" int mmb_attention_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                    int B, int S, int H, int causal, float scale, void* stream) {
  if (B <= 0 || S <= 0 || S > 256) return MMB_ERR_UNSUPPORTED;
  const int d = H * 64, S_pad = (S + 15) & ~15;
  CUtensorMap q128, q64, o128, o64;
  int rc = make_tmap_2d(&q128, qkv, 2, false, 3ull * d, (uint64_t)B * S, 3ull * d * 2, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&q64, qkv, 2, false, 3ull * d, (uint64_t)B * S, 3ull * d * 2, 64, 64);
  if (rc) return rc;
  rc = make_tmap_2d(&o128, dout, 2, false, (uint64_t)d, (uint64_t)B * S, (uint64_t)d * 2, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&o64, dout, 2, false, (uint64_t)d, (uint64_t)B * S, (uint64_t)d * 2, 64, 64);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // scratch for D = rowsum(dO * O): one buffer per device, grows on demand (cudaFree synchronises, so a buffer still
  // in use by an earlier launch is never freed under it)
  constexpr int MAX_DEV = 64;
  static float* dsum_dev[MAX_DEV] = {};
  static size_t dsum_cap[MAX_DEV] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  if (cur_dev < 0 || cur_dev >= MAX_DEV) return MMB_ERR_UNSUPPORTED;
  const size_t need = (size_t)B * H * S * sizeof(float);
  if (need > dsum_cap[cur_dev]) {
    if (dsum_dev[cur_dev]) cudaFree(dsum_dev[cur_dev]);
    cudaError_t e = cudaMalloc(&dsum_dev[cur_dev], need);
    if (e != cudaSuccess) { dsum_dev[cur_dev] = nullptr; dsum_cap[cur_dev] = 0; return (int)e; }
    dsum_cap[cur_dev] = need;
  }
  float* dsum = dsum_dev[cur_dev];
  AttnTcArgs a{};
  a.S = S; a.H = H; a.S_pad = S_pad; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  a.lse = const_cast<float*>(lse); a.o_in = (const __nv_bfloat16*)out; a.dout = (const __nv_bfloat16*)dout;
  a.dqkv = (__nv_bfloat16*)dqkv; a.dsum = dsum;
  dim3 grid((S + 127) / 128, H, B);
  static int persist_env = -1;  // MMB_ATTN_BWD_PERSIST=0 selects the two-CTA-per-SM kernels (A/B testing)
  if (persist_env < 0) {
    const char* e = getenv("MMB_ATTN_BWD_PERSIST");
    persist_env = (e && e[0] == '0') ? 0 : 1;
  }
  if (persist_env) {
    const int n_work = ((S + 127) / 128) * H * B;
    int dev_id = 0, sms = 148;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    const int grid_p = n_work < sms ? n_work : sms;
    static int png_env = -1;  // MMB_ATTN_PNG=2|4: worker column groups of the persistent kernel (default 2)
    if (png_env < 0) {
      const char* e = getenv("MMB_ATTN_PNG");
      png_env = (e && e[0] == '4') ? 4 : 2;   // 16 worker warps measured 3 % slower than 8 (TMEM-read bound)
    }
#define LAUNCH_BWDP(C, K, SM)                                                                                     \
  if (png_env == 2) {                                                                                             \
    cudaFuncSetAttribute(attn_bwd_persist_kernel<C, K, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);     \
    attn_bwd_persist_kernel<C, K, 2><<<grid_p, (8 + 3 + P_STATS) * 32, SM, st>>>(q128, q64, o128, o64, a, n_work); \
  } else {                                                                                                        \
    cudaFuncSetAttribute(attn_bwd_persist_kernel<C, K, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);     \
    attn_bwd_persist_kernel<C, K, 4><<<grid_p, (16 + 3 + P_STATS) * 32, SM, st>>>(q128, q64, o128, o64, a, n_work); \
  }
    if (causal) {
      LAUNCH_BWDP(true, false, BWDP_DQ_SMEM)
      LAUNCH_BWDP(true, true, BWDP_DKDV_SMEM)
    } else {
      LAUNCH_BWDP(false, false, BWDP_DQ_SMEM)
      LAUNCH_BWDP(false, true, BWDP_DKDV_SMEM)
    }
#undef LAUNCH_BWDP
    return (int)cudaGetLastError();
  }
  static int ng_env = -1;  // MMB_ATTN_NG=2|4: worker column groups (A/B testing; 4 measured slower); default 2
  if (ng_env < 0) {
    const char* e = getenv("MMB_ATTN_NG");
    ng_env = (e && e[0] == '4') ? 4 : 2;
  }
#define LAUNCH_BWD(C, K, SM)                                                                                  \
  if (ng_env == 2) {                                                                                          \
    cudaFuncSetAttribute(attn_bwd_tc_kernel<C, K, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);      \
    attn_bwd_tc_kernel<C, K, 2><<<grid, 2 * 128 + 32, SM, st>>>(q128, q64, o128, o64, a);                    \
  } else {                                                                                                    \
    cudaFuncSetAttribute(attn_bwd_tc_kernel<C, K, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);      \
    attn_bwd_tc_kernel<C, K, 4><<<grid, 4 * 128 + 32, SM, st>>>(q128, q64, o128, o64, a);                    \
  }
  if (causal) {
    LAUNCH_BWD(true, false, BWD_DQ_SMEM)     // dQ first: it also publishes D for the dK/dV kernel
    LAUNCH_BWD(true, true, BWD_DKDV_SMEM)
  } else {
    LAUNCH_BWD(false, false, BWD_DQ_SMEM)
    LAUNCH_BWD(false, true, BWD_DKDV_SMEM)
  }
#undef LAUNCH_BWD
  return (int)cudaGetLastError();
}
