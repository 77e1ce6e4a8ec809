// tcgen05 attention for S <= 384, head_dim 64 (CLIP ViT-B/16 image S = 197, text S = 77; ViT-L/14 S = 257; FLAVA): scores
// and gradients are computed by 5th-gen tensor cores with TMEM accumulators; operands are staged by TMA straight out of
// the packed in-projection output [B*S, 3d] (no head split / merge copies); the softmax / dS math runs one thread per tile
// row on TMEM rows (tcgen05.ld 32x32b).  Replaces F.scaled_dot_product_attention + autograd (torch/nn/functional.py:6682).
//
// Kernels in this file (dispatch: mmb_attention_fwd_tc / _bwd_tc at the end; DESIGN.md section 4 has the measurements):
//   forward   attn_fwd_item_kernel   128 < S <= 256, no mask: persistent, work item = (batch, head) with both query tiles
//             attn_fwd_pp_kernel     S <= 128: persistent, two worker groups alternate tiles
//             attn_fwd_tc_kernel     everything else (key mask, causal S > 128, 256 < S <= 384): one tile per CTA
//   backward  attn_bwd_fused_kernel  S <= 256: ONE pass, rows = keys, dQ accumulated in TMEM across the key tiles
//             attn_bwd_pp_kernel     256 < S <= 384: dQ pass + dK/dV pass, worker groups alternate 64-row chunks
//             attn_bwd_persist_kernel  the round-1 two-pass kernel (MMB_ATTN_BWD=colsplit, A/B only)
// Every tcgen05.mma / TMA issue runs on an elected lane of a shfl-uniform warp (common.cuh: elect_one, uniform_warp_idx).
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
  int l2_prefetch;              // bwd fused: prefetch the next work item's operands into L2 (MMB_ATTN_L2PF=0 disables)
  unsigned long long* trace;    // debug: per-phase SM-clock totals of CTAs 0-3 (scripts/attn_item_trace.py), or nullptr
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
                   const __grid_constant__ CUtensorMap tmRem, const AttnTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment for the 128B-swizzle atoms; pointer arithmetic on the __shared__ array keeps the address
  // space known to the compiler (LDS/STS instead of generic LD/ST).
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // S_pad <= 256: Q 16 KB | K <= 32 KB (aliased by P, 4 atoms, once S = Q K^T has completed) | V <= 32 KB: two CTAs
  // per SM.  256 < S_pad <= 384 (ViT-L/14 with CLS, FLAVA multimodal): K and V take 3 atoms each, P 6 (one CTA / SM).
  const int S = p.S, S_pad = p.S_pad, d = p.H * 64;
  const bool big = S_pad > 256;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATOM;
  uint8_t* sP = smem;
  uint8_t* sV = smem + (big ? 6 : 4) * ATOM;
  float* sRed = reinterpret_cast<float*>(smem + (big ? 9 : 6) * ATOM);  // [2][128] max, [2][128] sum
  uint8_t* sMask = reinterpret_cast<uint8_t*>(sRed + 512);   // [SMAX_FWD] key mask of this batch row (1 = attend)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + 384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const uint32_t ncols = S_pad <= 128 ? 128u : (S_pad <= 256 ? 256u : 512u);
  const bool has_mask = p.kmask != nullptr;
  for (int i = threadIdx.x; i < 384; i += ATT_THREADS)
    sMask[i] = (i < S && (!has_mask || p.kmask[(long long)blockIdx.z * S + i])) ? 1 : 0;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm128);
    tma_prefetch_desc(&tmPad);
    if (big) tma_prefetch_desc(&tmRem);
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

  if (warp == 0 && elect_one()) {
    // K / V rows [0, min(S_pad,256)) come in one box (tmPad), the remainder (S_pad > 256) in a second one (tmRem)
    const int n1 = big ? 256 : S_pad, n2 = S_pad - n1;
    mbar_arrive_expect_tx(&bars[0], ATOM + S_pad * 128);
    tma_load_2d(&tm128, &bars[0], sQ, h * 64, row0 + qt * 128);
    tma_load_2d(&tmPad, &bars[0], sK, d + h * 64, row0);
    if (big) tma_load_2d(&tmRem, &bars[0], sK + 256 * 128, d + h * 64, row0 + 256);
    mbar_arrive_expect_tx(&bars[1], S_pad * 128);
    tma_load_2d(&tmPad, &bars[1], sV, 2 * d + h * 64, row0);
    if (big) tma_load_2d(&tmRem, &bars[1], sV + 256 * 128, 2 * d + h * 64, row0 + 256);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint64_t da = desc_k(smem_u32(sQ)), db = desc_k(smem_u32(sK));
    {
      const uint32_t id = idesc_rt(n1, false, false);   // UMMA N <= 256: the scores are issued in two column blocks
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, id, k > 0);
    }
    if (big) {
      const uint32_t id = idesc_rt(n2, false, false);
      const uint64_t db2 = desc_k(smem_u32(sK) + 256 * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + 256, da + 2 * k, db2 + 2 * k, id, k > 0);
    }
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
  // a warp whose 32 rows are all padding (last tile: 59 of 128 rows at S = 197) skips both softmax passes (warp-uniform,
  // so the .sync.aligned tcgen05.ld stay convergent); its P rows are never read back and its O rows never stored
  const bool warp_ok = qt * 128 + q4 * 32 < S;

  float mx = -INFINITY;
  for (int c = c_lo; warp_ok && c < c_hi; ++c) {
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
  for (int c = c_lo; warp_ok && c < c_hi; ++c) {
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

  if (warp == 0 && elect_one()) {
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
// Forward, persistent PING-PONG version (round 2; used for S <= 128: the text tower).  One CTA per SM loops over
// (batch, head, 128-query tile) work items; two tile buffers (smem operands + a 256-column TMEM score block each) are
// in flight, each served by its own group of four worker warps:
//   * producer warp: TMA loads of Q / K / V of tile n+2 start as soon as the PV MMA of tile n has left the buffer;
//   * MMA warp: S_n = Q K^T, then — while group n&1 runs the softmax of tile n — S_{n+1}, then O_n = P_n V_n ...;
//   * worker group g = n & 1: ONE THREAD PER QUERY ROW owning the whole row (no cross-thread max / sum exchange, no
//     __syncthreads): pass 1 row maximum, pass 2 exp / row sum / bf16 P into the swizzled smem operand (aliasing Q | K),
//     software-pipelined tcgen05.ld (chunk c+1 in flight while chunk c is reduced); epilogue O / l and the LSE.
// The round-1 kernel (one tile per CTA, two CTAs per SM) exposed a serial chain per tile — TMA latency, MMA, pass 1,
// __syncthreads, pass 2, __syncthreads, MMA, epilogue: 5.9 k clocks per tile per SM against a MUFU floor of 1.7 k — and
// computed exp() for the padding rows of the last tile (59 of 128 rows at S = 197); rows >= S are skipped here.
// TMEM: score block i at columns [256 i, 256 i + S_pad); O_i reuses columns [256 i, 256 i + 64).
// ------------------------------------------------------------------------------------------------
constexpr int FPP_WORKERS = 8;                      // two groups of four warps
constexpr int FPP_THREADS = (FPP_WORKERS + 4) * 32; // + producer + MMA + 2 idle warps (register allocation is per 4 warps)
constexpr int FPP_BUF = 6 * ATOM;                   // [P (4 atoms) aliasing Q (atom 0) | K (atoms 1-2)] + V (2 atoms)
template <bool CAUSAL>
__global__ void __launch_bounds__(FPP_THREADS, 1)
attn_fwd_pp_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tmPad,
                   const AttnTcArgs p, const int n_work) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sMaskAll = smem + 2 * FPP_BUF;                               // [2][256] key mask (kmask variant only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMaskAll + 512);
  uint64_t* qk_full = bars;          // [2] Q, K landed                         producer (TMA) -> MMA warp
  uint64_t* v_full = bars + 2;       // [2] V landed                            producer (TMA) -> MMA warp
  uint64_t* buf_empty = bars + 4;    // [2] PV MMA complete: smem buffer free   MMA warp       -> producer
  uint64_t* s_full = bars + 6;       // [2] S in TMEM                           MMA warp       -> worker group
  uint64_t* p_full = bars + 8;       // [2] P in smem (4 warps)                 worker group   -> MMA warp
  uint64_t* o_full = bars + 10;      // [2] O in TMEM                           MMA warp       -> worker group
  uint64_t* t_empty = bars + 12;     // [2] O read out of TMEM (4 warps)        worker group   -> MMA warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int S = p.S, S_pad = p.S_pad, d = p.H * 64;
  const int ntile = (S + 127) >> 7;
  const bool has_mask = p.kmask != nullptr;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm128);
    tma_prefetch_desc(&tmPad);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&buf_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], FPP_WORKERS / 2);
      mbar_init(&o_full[i], 1); mbar_init(&t_empty[i], FPP_WORKERS / 2);
    }
    fence_mbar_init();
  }
  if (warp == FPP_WORKERS) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == FPP_WORKERS) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      int n = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
        const int tile = w % ntile, bh = w / ntile;
        const int h = bh % p.H, b = bh / p.H;
        const int row0 = b * S, i = n & 1;
        uint8_t* buf = smem + i * FPP_BUF;
        mbar_wait(&buf_empty[i], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qk_full[i], ATOM + S_pad * 128);
        tma_load_2d(&tm128, &qk_full[i], buf, h * 64, row0 + tile * 128);            // Q tile
        tma_load_2d(&tmPad, &qk_full[i], buf + ATOM, d + h * 64, row0);              // K (all keys of the head)
        mbar_arrive_expect_tx(&v_full[i], S_pad * 128);
        tma_load_2d(&tmPad, &v_full[i], buf + 4 * ATOM, 2 * d + h * 64, row0);       // V
      }
    }
  } else if (warp == FPP_WORKERS + 1) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      const uint32_t id_s = idesc_rt(S_pad, false, false), id_o = idesc_rt(64, false, true);
      const int nk = S_pad >> 4;
      auto issue_pv = [&](int m) {            // O_m = P_m V_m  (K-steps of 16 keys)
        const int i = m & 1;
        const uint32_t ub = smem_u32(smem + i * FPP_BUF);
        mbar_wait(&p_full[i], (m >> 1) & 1);
        mbar_wait(&v_full[i], (m >> 1) & 1);
        tc_fence_after();
        for (int j = 0; j < nk; ++j)
          umma_bf16(tmem + i * 256, desc_k(ub + (j >> 2) * ATOM + (j & 3) * 32), desc_mn(ub + 4 * ATOM + j * 2048), id_o,
                    j > 0);
        umma_commit(&o_full[i]);
        umma_commit(&buf_empty[i]);
      };
      int n = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
        const int i = n & 1;
        const uint32_t ub = smem_u32(smem + i * FPP_BUF);
        mbar_wait(&qk_full[i], (n >> 1) & 1);
        mbar_wait(&t_empty[i], ((n >> 1) & 1) ^ 1);      // O of tile n-2 has been read out of this TMEM block
        tc_fence_after();
        const uint64_t da = desc_k(ub), db = desc_k(ub + ATOM);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + i * 256, da + 2 * k, db + 2 * k, id_s, k > 0);
        umma_commit(&s_full[i]);
        if (n >= 1) issue_pv(n - 1);
      }
      if (n >= 1) issue_pv(n - 1);
    }
  } else if (warp < FPP_WORKERS) {
    // ======================= worker groups: thread == query row, group g serves tiles n with (n & 1) == g ===========
    const int grp = warp >> 2, q4 = warp & 3;
    const int r = q4 * 32 + lane;
    uint8_t* buf = smem + grp * FPP_BUF;
    uint8_t* sMask = sMaskAll + grp * 256;
    const uint32_t trow = tmem + grp * 256 + ((uint32_t)(q4 * 32) << 16);
    const int nc32 = (S_pad + 31) >> 5;   // 32-column steps; the last one may be 16 wide (S_pad % 32 == 16)
    int n = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
      if ((n & 1) != grp) continue;
      const int tile = w % ntile, bh = w / ntile;
      const int h = bh % p.H, b = bh / p.H;
      const int row0 = b * S;
      const int qi = tile * 128 + r;
      const uint32_t par = (uint32_t)((n >> 1) & 1);
      if (has_mask) {   // key-padding mask of this batch row -> smem (group-private), visible after the group barrier
        for (int j = q4 * 32 + lane; j < 256; j += 128) sMask[j] = (j < S && p.kmask[(long long)b * S + j]) ? 1 : 0;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      }
      mbar_wait(&s_full[grp], par);
      tc_fence_after();
      // tcgen05.ld / tcgen05.wait are .sync.aligned: every lane of the warp must execute them together, so all
      // control flow around them is WARP-uniform.  A warp whose 32 rows are all padding (last tile) skips the math; in a
      // partially valid warp the padding lanes run along on whatever their TMEM rows hold (finite or not, it stays in
      // their own P rows / O rows, which are never stored).
      const bool row_ok = qi < S;
      const bool warp_ok = tile * 128 + q4 * 32 < S;     // uniform: the warp's first row is a real query
      const int kv_lim = CAUSAL ? min(S, qi + 1) : S;    // columns >= kv_lim are masked (per lane)
      const int kv_lim_w = CAUSAL ? min(S, tile * 128 + q4 * 32 + 32) : S;   // uniform bound over the warp's rows
      float mx = -INFINITY, sum = 0.f;
      if (warp_ok) {
        // ---- pass 1: row maximum (tcgen05.ld of step c+1 in flight while step c is reduced) ----
        uint32_t va[32], vb[32];
        auto ld = [&](int c, uint32_t (&v)[32]) {
          if (c * 32 + 32 <= S_pad) tmem_ld32(trow + c * 32, v);
          else tmem_ld16(trow + c * 32, reinterpret_cast<uint32_t(&)[16]>(v));
        };
        auto red_max = [&](int c, const uint32_t (&v)[32]) {
          const int wdt = min(32, S_pad - c * 32);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            if (hf * 16 >= wdt) break;
            const int c0 = c * 32 + hf * 16;
            if (c0 + 16 <= kv_lim && !has_mask) {
#pragma unroll
              for (int e = 0; e < 16; ++e) mx = fmaxf(mx, __uint_as_float(v[hf * 16 + e]));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (c0 + e < kv_lim && (!has_mask || sMask[c0 + e])) mx = fmaxf(mx, __uint_as_float(v[hf * 16 + e]));
            }
          }
        };
        const int nc_act = CAUSAL ? min(nc32, (kv_lim_w + 31) >> 5) : nc32;   // steps with unmasked columns (uniform)
        ld(0, va);
        for (int c = 0; c < nc_act; c += 2) {
          tmem_ld_wait();
          if (c + 1 < nc_act) ld(c + 1, vb);
          red_max(c, va);
          if (c + 1 < nc_act) {
            tmem_ld_wait();
            if (c + 2 < nc_act) ld(c + 2, va);
            red_max(c + 1, vb);
          }
        }
        mx = (mx == -INFINITY) ? 0.f : mx * p.scale_log2;   // fully masked row (kmask): exp() of nothing, l = 0
        // ---- pass 2: p = 2^(s*scale*log2e - mx), row sum, bf16 P into the K-major swizzled operand ----
        auto exp_store = [&](int c, const uint32_t (&v)[32]) {
          const int wdt = min(32, S_pad - c * 32);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            if (hf * 16 >= wdt) break;
            float pr[16];
            const int c0 = c * 32 + hf * 16;
            if (c0 + 16 <= kv_lim && !has_mask) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                pr[e] = ex2_approx(fmaf(__uint_as_float(v[hf * 16 + e]), p.scale_log2, -mx));
                sum += pr[e];
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const bool ok = (c0 + e < kv_lim) && (!has_mask || sMask[c0 + e]);
                const float x = ok ? ex2_approx(fmaf(__uint_as_float(v[hf * 16 + e]), p.scale_log2, -mx)) : 0.f;
                pr[e] = x;
                sum += x;
              }
            }
            store_p16(buf, r, c0, pr);
          }
        };
        ld(0, va);
        for (int c = 0; c < nc_act; c += 2) {
          tmem_ld_wait();
          if (c + 1 < nc_act) ld(c + 1, vb);
          exp_store(c, va);
          if (c + 1 < nc_act) {
            tmem_ld_wait();
            if (c + 2 < nc_act) ld(c + 2, va);
            exp_store(c + 1, vb);
          }
        }
        if (CAUSAL) {   // steps entirely above the diagonal: P = 0 (the PV MMA runs over all S_pad keys)
          const float z[16] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int c0 = nc_act * 32; c0 < S_pad; c0 += 16) store_p16(buf, r, c0, z);
        }
      }
      // every S column of this row has been read (or is never needed): P may be consumed, TMEM block reused for O
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[grp]);
      // ---- epilogue: O (TMEM, 64 columns) / l -> bf16 -> out ; LSE ----
      mbar_wait(&o_full[grp], par);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      if (warp_ok) {      // uniform
        tmem_ld32(trow, o0);
        tmem_ld32(trow + 32, o1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[grp]);
      if (row_ok) {
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        __nv_bfloat16* dst = p.out + ((long long)(row0 + qi)) * d + h * 64;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t* v = half == 0 ? o0 : o1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]) * inv, __uint_as_float(v[j * 8 + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]) * inv, __uint_as_float(v[j * 8 + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]) * inv, __uint_as_float(v[j * 8 + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]) * inv, __uint_as_float(v[j * 8 + 7]) * inv);
            reinterpret_cast<uint4*>(dst)[half * 4 + j] = o;
          }
        }
        if (p.lse) p.lse[((long long)b * p.H + h) * S + qi] = (mx + log2f(sum)) * 0.6931471805599453f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FPP_WORKERS) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
constexpr int FPP_SMEM = 1024 + 2 * FPP_BUF + 512 + 256;

// ------------------------------------------------------------------------------------------------
// Forward, persistent ITEM kernel (round 2, second iteration; 128 < S <= 256, no mask: the image towers).  Work item =
// one (batch, head) with BOTH of its 128-query tiles, so K and V are fetched once per head instead of once per tile:
//   * producer warp: Q (2 tiles) + K of item n+1 are loaded as soon as both S = Q K^T MMAs of item n have completed,
//     V of item n+1 as soon as both P V MMAs of item n have - operands are single-buffered, their lifetimes staggered;
//   * two MMA issuer warps, one per query tile, each running its own  S_t -> (softmax) -> O_t  chain, so the two tiles
//     drift into anti-phase: the tensor pipe works for one tile while the other tile's softmax owns the issue slots;
//   * two softmax groups of four warps, thread == query row: pass 1 row maximum with 3-input max on four independent
//     chains, pass 2 packed f32x2 scale/subtract and row sums (FFMA2 / FADD2), ex2, bf16 P into its own swizzled smem
//     operand (2 x 64 KB), tcgen05.ld of chunk c+1 in flight while chunk c is processed.
// Measured floors on this part (scripts/probes/{tmem,softmax}_probe.cu): tcgen05.ld 345 B/clk/SM with 8 warps (one
// 128x256 fp32 block in 380 clk); MUFU.EX2 16 / clk / SM, so the exp pass is MUFU-bound at 15 scores / clk / SM (>= 3.3 k
// clocks per work item) and everything else — max pass, MMA waits, epilogue — has to hide under the other tile's exp pass
// (the round-1 tile kernel spent 8.7 warp-instructions per score element at 33 % issue utilisation).
// smem: Q 32 KB | K 32 KB | V 32 KB | P_0 64 KB | P_1 64 KB = 224 KB.  TMEM: S_t at columns [256 t, 256 t + S_pad), O_t
// reuses [256 t, 256 t + 64) once the softmax has consumed S_t.
// ------------------------------------------------------------------------------------------------
constexpr int FIT_THREADS = 12 * 32;   // 8 softmax warps + producer + 2 MMA issuers + 1 idle (registers: per 4 warps)
constexpr int FIT_SMEM = 1024 + 14 * ATOM + 256;
template <bool TRACE>
__global__ void __launch_bounds__(FIT_THREADS, 1)
attn_fwd_item_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tmPad,
                     const __grid_constant__ CUtensorMap tmOut, const AttnTcArgs p, const int n_items, const int stagger,
                     unsigned long long* trace) {
  // trace (debug, scripts/attn_item_trace.py): per-phase SM-clock totals of CTAs 0-3, [cta][warp][64]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                 // 2 atoms: query tile 0, query tile 1
  uint8_t* sK = smem + 2 * ATOM;      // S_pad rows
  uint8_t* sV = smem + 4 * ATOM;      // S_pad rows
  uint8_t* sPall = smem + 6 * ATOM;   // 2 x 4 atoms
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 14 * ATOM);
  uint64_t* qk_full = bars;           // Q, K landed                              producer -> MMA warps
  uint64_t* v_full = bars + 1;        // V landed                                 producer -> MMA warps
  uint64_t* qk_empty = bars + 2;      // both S MMAs complete (2 commits)         MMA warps -> producer
  uint64_t* v_empty = bars + 3;       // both PV MMAs complete (2 commits)        MMA warps -> producer
  uint64_t* s_full = bars + 4;        // [2] S_t in TMEM                          MMA warp t -> softmax group t
  uint64_t* p_full = bars + 6;        // [2] P_t in smem (4 warps)                softmax group t -> MMA warp t
  uint64_t* o_full = bars + 8;        // [2] O_t in TMEM                          MMA warp t -> softmax group t
  uint64_t* s_free = bars + 10;       // [2] O_t read out of TMEM (4 warps)       softmax group t -> MMA warp t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int S = p.S, S_pad = p.S_pad, d = p.H * 64;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm128);
    tma_prefetch_desc(&tmPad);
    tma_prefetch_desc(&tmOut);
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(qk_empty, 2); mbar_init(v_empty, 2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&s_free[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int h = w % p.H, b = w / p.H;
        const int row0 = b * S;
        const uint32_t par = (uint32_t)(it & 1);
        mbar_wait(qk_empty, par ^ 1);
        mbar_arrive_expect_tx(qk_full, 2 * ATOM + S_pad * 128);
        tma_load_2d(&tm128, qk_full, sQ, h * 64, row0);
        tma_load_2d(&tm128, qk_full, sQ + ATOM, h * 64, row0 + 128);
        tma_load_2d(&tmPad, qk_full, sK, d + h * 64, row0);
        mbar_wait(v_empty, par ^ 1);
        mbar_arrive_expect_tx(v_full, S_pad * 128);
        tma_load_2d(&tmPad, v_full, sV, 2 * d + h * 64, row0);
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ======================= MMA issuer of query tile t =======================
    if (elect_one()) {
      const int t = warp - 9;
      const uint32_t id_s = idesc_rt(S_pad, false, false), id_o = idesc_rt(64, false, true);
      const int nk = S_pad >> 4;
      const uint32_t uQ = smem_u32(sQ) + t * ATOM, uK = smem_u32(sK), uV = smem_u32(sV);
      const uint32_t uP = smem_u32(sPall) + t * 4 * ATOM;
      const uint32_t tacc = tmem + t * 256;
      int it = 0;
      // Anti-phase start: tile 1's chain begins once tile 0's first softmax has finished, so that from then on one
      // group's exp pass (MUFU-bound: 16 ex2 / clk / SM) overlaps the other group's maximum pass, MMA waits and epilogue
      // instead of both groups fighting for the MUFU at the same time and idling together afterwards.
      if (stagger && t == 1) mbar_wait(&p_full[0], 0);
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        mbar_wait(qk_full, par);
        mbar_wait(&s_free[t], par ^ 1);       // O_t of the previous item has left this TMEM block
        tc_fence_after();
        if (TRACE && blockIdx.x < 4 && it < 24) trace[(blockIdx.x * 12 + warp) * 64 + 2 * it] = clock64();
        const uint64_t da = desc_k(uQ), db = desc_k(uK);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tacc, da + 2 * k, db + 2 * k, id_s, k > 0);
        umma_commit(&s_full[t]);
        umma_commit(qk_empty);
        mbar_wait(&p_full[t], par);
        mbar_wait(v_full, par);
        tc_fence_after();
        if (TRACE && blockIdx.x < 4 && it < 24) trace[(blockIdx.x * 12 + warp) * 64 + 2 * it + 1] = clock64();
        {   // descriptor bases once; per 16-key step: P advances 32 B inside an atom (+2), 16 KB per atom (+1024); V 2048 B (+128)
          const uint64_t aP = desc_k(uP), bV = desc_mn(uV);
          for (int j = 0; j < nk; ++j) umma_bf16(tacc, aP + (uint64_t)((j >> 2) * (ATOM >> 4) + (j & 3) * 2), bV + 128 * j, id_o, j > 0);
        }
        umma_commit(&o_full[t]);
        umma_commit(v_empty);
      }
    }
  } else if (warp < 8) {
    // ======================= softmax group t: thread == query row of tile t =======================
    const int t = warp >> 2, q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const int qi = t * 128 + r;
    uint8_t* sP = sPall + t * 4 * ATOM;
    const uint32_t trow = tmem + t * 256 + ((uint32_t)(q4 * 32) << 16);
    const bool row_ok = qi < S;
    const bool warp_ok = t * 128 + q4 * 32 < S;   // warp-uniform (tcgen05.ld / wait are .sync.aligned)
    const int nfull = S >> 5;                     // 32-column steps whose keys are all real
    const int nstep = (S_pad + 31) >> 5;          // + at most one partial step (16 or 32 columns wide, some keys padding)
    const uint64_t c2 = pk2(p.scale_log2, p.scale_log2);
    int it = 0;
    const bool tr = TRACE && blockIdx.x < 4 && lane == 0;
    unsigned long long* trw = trace + (blockIdx.x * 12 + warp) * 64;
    long long tp = tr ? clock64() : 0, acc_ws = 0, acc_p1 = 0, acc_p2 = 0, acc_wo = 0, acc_ep = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int h = w % p.H, b = w / p.H;
      const uint32_t par = (uint32_t)(it & 1);
      mbar_wait(&s_full[t], par);
      tc_fence_after();
      long long t0 = 0, t1 = 0;
      if (tr) { t0 = clock64(); acc_ws += t0 - tp; if (it < 20) trw[8 + 2 * it] = t0; }
      float mx = 0.f, sum = 1.f;
      if (warp_ok) {
        uint32_t va[32], vb[32];
        auto ld = [&](int c, uint32_t (&v)[32]) {
          if (c * 32 + 32 <= S_pad) tmem_ld32(trow + c * 32, v);
          else tmem_ld16(trow + c * 32, reinterpret_cast<uint32_t(&)[16]>(v));
        };
        // ---- pass 1: row maximum, four independent chains ----
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        float m4 = -INFINITY, m5 = -INFINITY, m6 = -INFINITY, m7 = -INFINITY;
        auto red_max = [&](int c, const uint32_t (&v)[32]) {
          if (c < nfull) {   // eight independent 3-input chains (scripts/probes/softmax_probe.cu: 55 vs 89 clk / step)
#pragma unroll
            for (int e = 0; e < 32; e += 16) {
              m0 = fmax3(m0, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
              m1 = fmax3(m1, __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
              m2 = fmax3(m2, __uint_as_float(v[e + 4]), __uint_as_float(v[e + 5]));
              m3 = fmax3(m3, __uint_as_float(v[e + 6]), __uint_as_float(v[e + 7]));
              m4 = fmax3(m4, __uint_as_float(v[e + 8]), __uint_as_float(v[e + 9]));
              m5 = fmax3(m5, __uint_as_float(v[e + 10]), __uint_as_float(v[e + 11]));
              m6 = fmax3(m6, __uint_as_float(v[e + 12]), __uint_as_float(v[e + 13]));
              m7 = fmax3(m7, __uint_as_float(v[e + 14]), __uint_as_float(v[e + 15]));
            }
          } else {
            const int nv = S - c * 32;     // real keys in the partial step (< 32)
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e < nv) m0 = fmaxf(m0, __uint_as_float(v[e]));
          }
        };
        ld(0, va);
        for (int c = 0; c < nstep; c += 2) {
          tmem_ld_wait();
          if (c + 1 < nstep) ld(c + 1, vb);
          red_max(c, va);
          if (c + 1 < nstep) {
            tmem_ld_wait();
            if (c + 2 < nstep) ld(c + 2, va);
            red_max(c + 1, vb);
          }
        }
        mx = fmaxf(fmaxf(fmax3(m0, m1, m2), fmax3(m3, m4, m5)), fmaxf(m6, m7)) * p.scale_log2;
        // the previous item's O tile was staged in the first atom of sP for its TMA store: the store must have read
        // it before pass 2 overwrites the atom (one elected thread waits, the group barrier tells the others)
        if (q4 == 0 && lane == 0) tma_store_wait_read<0>();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
        if (tr) { t1 = clock64(); acc_p1 += t1 - t0; }
        // ---- pass 2: p = 2^(s * scale * log2e - mx), row sum, bf16 P into the K-major swizzled operand ----
        const uint64_t nm2 = pk2(-mx, -mx);
        uint64_t s0 = pk2(0.f, 0.f), s1 = s0;
        float st = 0.f;
        auto exp_store = [&](int c, const uint32_t (&v)[32]) {
          if (c < nfull) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                float x0, x1, x2, x3;
                upk2(ffma2(pk2(__uint_as_float(v[hf * 16 + e]), __uint_as_float(v[hf * 16 + e + 1])), c2, nm2), x0, x1);
                upk2(ffma2(pk2(__uint_as_float(v[hf * 16 + e + 2]), __uint_as_float(v[hf * 16 + e + 3])), c2, nm2), x2, x3);
                x0 = ex2_approx(x0); x1 = ex2_approx(x1); x2 = ex2_approx(x2); x3 = ex2_approx(x3);
                s0 = fadd2(s0, pk2(x0, x1));
                s1 = fadd2(s1, pk2(x2, x3));
                pk[e >> 1] = pack_bf16x2(x0, x1);
                pk[(e >> 1) + 1] = pack_bf16x2(x2, x3);
              }
              const int j0 = c * 32 + hf * 16;
              uint8_t* a = sP + (j0 >> 6) * ATOM + r * 128;
              const int c8 = (j0 & 63) >> 3;
              *reinterpret_cast<uint4*>(a + ((c8 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(a + (((c8 + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          } else {
            const int nv = S - c * 32, wdt = S_pad - c * 32;   // real keys / columns of the partial step
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              if (hf * 16 >= wdt) break;
              float pr[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float x = (hf * 16 + e < nv) ? ex2_approx(fmaf(__uint_as_float(v[hf * 16 + e]), p.scale_log2, -mx)) : 0.f;
                pr[e] = x;
                st += x;
              }
              store_p16(sP, r, c * 32 + hf * 16, pr);
            }
          }
        };
        ld(0, va);
        for (int c = 0; c < nstep; c += 2) {
          tmem_ld_wait();
          if (c + 1 < nstep) ld(c + 1, vb);
          exp_store(c, va);
          if (c + 1 < nstep) {
            tmem_ld_wait();
            if (c + 2 < nstep) ld(c + 2, va);
            exp_store(c + 1, vb);
          }
        }
        float a0, a1, a2, a3;
        upk2(s0, a0, a1);
        upk2(s1, a2, a3);
        sum = (a0 + a1) + (a2 + a3) + st;
      } else {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");   // the group barrier of the branch above
      }
      // every S column of this row has been read: P may be consumed, the TMEM block reused for O
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      long long t2 = 0;
      if (tr) { t2 = clock64(); acc_p2 += t2 - t1; if (it < 20) trw[9 + 2 * it] = t2; }
      // ---- epilogue: O (64 columns) / l -> bf16 -> out ; LSE ----
      mbar_wait(&o_full[t], par);
      tc_fence_after();
      if (tr) { const long long t3 = clock64(); acc_wo += t3 - t2; t2 = t3; }
      uint32_t o0[32], o1[32];
      if (warp_ok) {
        tmem_ld32(trow, o0);
        tmem_ld32(trow + 32, o1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);
      // O tile -> bf16 -> swizzled staging atom (P_t is dead once O_t exists) -> ONE TMA store per tile: a 3-D box
      // [1 batch][128 rows][64 columns] that the tensor map clips at the sequence length.  (Direct st.global of one
      // 128 B row per thread costs 32 L1 wavefronts per instruction, 2.2-3.8 k clocks per tile in the phase trace.)
      if (warp_ok) {
        const float inv = row_ok ? 1.f / sum : 0.f;
        uint8_t* a = sP + r * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t* v = half == 0 ? o0 : o1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]) * inv, __uint_as_float(v[j * 8 + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]) * inv, __uint_as_float(v[j * 8 + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]) * inv, __uint_as_float(v[j * 8 + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]) * inv, __uint_as_float(v[j * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(a + (((half * 4 + j) ^ (r & 7)) << 4)) = o;
          }
        }
        if (row_ok && p.lse) p.lse[((long long)b * p.H + h) * S + qi] = (mx + log2f(sum)) * 0.6931471805599453f;
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
      if (q4 == 0 && lane == 0) {
        tma_store_3d(&tmOut, sP, h * 64, t * 128, b);
        tma_store_commit();
      }
      if (tr) { tp = clock64(); acc_ep += tp - t2; }
    }
    if (tr) {
      trw[0] = acc_ws; trw[1] = acc_p1; trw[2] = acc_p2; trw[3] = acc_wo; trw[4] = acc_ep; trw[5] = it;
    }
    if (q4 == 0 && lane == 0) tma_store_wait_all<0>();   // smem must outlive the last store's read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
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

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
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
    if (elect_one()) {
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
    if (elect_one()) {
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
    if (elect_one()) {
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

// ------------------------------------------------------------------------------------------------
// Backward, FUSED single pass (round 2).  The two-pass kernels above recompute S and dP (and their exponentials, the
// 64 KB-per-chunk TMEM reads that bound the workers) once for dQ and once for dK/dV.  Here ONE pass in the dK/dV
// orientation (tile rows = keys) also produces dQ: the dS^T chunk the workers leave in shared memory for
// dK += dS^T_c Q_c is, read as an MN-major A operand, exactly dS_c — so dQ_c += dS_c K_j is one more accumulate-MMA
// against the resident K tile.  Two 64-query chunks form one M = 128 operand (two MN blocks 16 KB apart), and the
// partial sums over the key tiles of a head are accumulated IN TMEM across the key-tile loop, so nothing is reduced
// through global memory or DSMEM.  Work item = (batch, head); S <= 256 (<= 2 key tiles, <= 4 query chunks, <= 2 dQ
// blocks).
// TMEM (512 columns): [S^T|dP^T] buffer i at i*128 (+0, +64); dV at 256, dK at 320 (one key tile at a time);
//                     dQ block p (queries 128p ..) at 384 + 64p, alive for the whole work item.
// smem: K_j | V_j tile operands x 2 (64 KB), (Q_c | dO_c) chunk ring x 3 (48 KB), dS^T x 4 (chunk c -> buffer c: the
// pairs (0,1), (2,3) are the two dQ A operands; 64 KB), P^T x 2 (32 KB), per-query LSE / D x 2 (4 KB).
// Warps: 8 workers (thread = key row x column half), TMA producer, score issuer, accumulate issuer, NSTAT statistics
// warps.  Output tiles (dV, dK per key tile, dQ per work item) leave through dead operand buffers + TMA stores.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t desc_mn_a(uint32_t saddr) { return make_smem_desc_sw128(saddr, 16384, 1024); }

// NSTAT statistics warps: 2 -> 13 warps, 128 registers per thread (register allocation is per 4 warps; ~180 B of
// spills); 1 -> 12 warps, 168 registers, no spills, half the statistics bandwidth.
template <bool CAUSAL, int NSTAT>
__global__ void __launch_bounds__((8 + 3 + NSTAT) * 32, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                      const __grid_constant__ CUtensorMap tmDO64, const __grid_constant__ CUtensorMap tmDQKV,
                      const AttnTcArgs p, const int n_work) {
  constexpr int RING = 3;
  constexpr int CH = 8192;  // one 64-row x 128 B chunk operand
  constexpr int P_WORKERS = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // [2 key-tile buffers][K_j 16 KB | V_j 16 KB]
  uint8_t* sRing = sA + 4 * ATOM;                       // RING x (Q_c 8 KB | dO_c 8 KB)
  uint8_t* sDS = sRing + RING * 2 * CH;                 // [4] dS^T chunk c (16 KB each)
  uint8_t* sPT = sDS + 4 * ATOM;                        // [2] P^T chunk
  float* sLD = reinterpret_cast<float*>(sPT + 2 * ATOM);   // [2][512]: lse(log2) [256] | rowsum(dO*O) [256] per query
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLD + 2 * 512);
  uint64_t* tile_full = bars;            // [2]  K_j, V_j landed                              producer -> both issuers
  uint64_t* tile_empty = bars + 2;       // [2]  both issuers' last MMA on the tile complete (2 commits) -> producer
  uint64_t* ring_full = bars + 4;        // [RING]
  uint64_t* ring_empty = bars + 8;       // [RING] accumulate MMAs of the chunk complete                -> producer
  uint64_t* sdp_full = bars + 12;        // [2]  S^T_c / dP^T_c in TMEM                 score issuer    -> workers
  uint64_t* sdp_empty = bars + 14;       // [2]  workers have read them (8 warps)                       -> score issuer
  uint64_t* ds_full = bars + 16;         // [2]  dS^T_c and P^T_c in smem (8 warps)     workers         -> acc issuer
  uint64_t* pt_empty = bars + 18;        // [2]  dV / dK MMAs of the chunk complete     acc issuer      -> workers
  uint64_t* dsb_empty = bars + 20;       // [4]  dQ MMA of the chunk's pair complete    acc issuer      -> workers
  uint64_t* acc_full = bars + 24;        // [1]  dV, dK of the key tile final           acc issuer      -> workers
  uint64_t* acc_empty = bars + 25;       // [1]  workers have read them (8 warps)                       -> acc issuer
  uint64_t* dq_full = bars + 26;         // [1]  dQ of the work item final              acc issuer      -> workers
  uint64_t* dq_empty = bars + 27;        // [1]  workers have read it (8 warps)                         -> acc issuer
  uint64_t* stat_full = bars + 28;       // [2]  LSE / D of the work item in sLD        stats warps     -> workers
  uint64_t* stat_empty = bars + 30;      // [2]  workers are done with them (8 warps)                   -> stats warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int S = p.S, S_pad = p.S_pad, d = p.H * 64;
  const int nc = (S_pad + 63) >> 6;      // query chunks (<= 4)
  const int ntile = (S + 127) >> 7;      // key tiles (<= 2)
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV128); tma_prefetch_desc(&tmQKV64); tma_prefetch_desc(&tmDO64); tma_prefetch_desc(&tmDQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tile_full[i], 1); mbar_init(&tile_empty[i], 2);
      mbar_init(&sdp_full[i], 1); mbar_init(&sdp_empty[i], P_WORKERS);
      mbar_init(&ds_full[i], P_WORKERS); mbar_init(&pt_empty[i], 1);
      mbar_init(&stat_full[i], NSTAT); mbar_init(&stat_empty[i], P_WORKERS);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&dsb_empty[i], 1);
    for (int i = 0; i < RING; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, P_WORKERS);
    mbar_init(dq_full, 1); mbar_init(dq_empty, P_WORKERS);
    fence_mbar_init();
  }
  if (warp == P_WORKERS) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == P_WORKERS) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      int nt = 0;
      int g = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int h = w % p.H, b = w / p.H;
        const int row0 = b * S;
        if (p.l2_prefetch && w + (int)gridDim.x < n_work) {
          // the 3-stage chunk ring is shallower than the HBM latency (chunk period >= 700 + L/2 clocks):
          // pull the NEXT work item's operands into L2 now, so that its ring loads see L2 latency instead
          const int w2 = w + gridDim.x;
          const int h2 = w2 % p.H, r2 = (w2 / p.H) * S;
          for (int j = 0; j < ntile; ++j) {
            tma_prefetch_l2_2d(&tmQKV128, d + h2 * 64, r2 + j * 128);
            tma_prefetch_l2_2d(&tmQKV128, 2 * d + h2 * 64, r2 + j * 128);
          }
          for (int c = 0; c < nc; ++c) {
            tma_prefetch_l2_2d(&tmQKV64, h2 * 64, r2 + c * 64);
            tma_prefetch_l2_2d(&tmDO64, h2 * 64, r2 + c * 64);
          }
        }
        for (int j = 0; j < ntile; ++j, ++nt) {
          const int tb = nt & 1;
          mbar_wait(&tile_empty[tb], ((nt >> 1) & 1) ^ 1);
          uint8_t* a0 = sA + tb * 2 * ATOM;
          mbar_arrive_expect_tx(&tile_full[tb], 2 * ATOM);
          tma_load_2d(&tmQKV128, &tile_full[tb], a0, d + h * 64, row0 + j * 128);             // K tile
          tma_load_2d(&tmQKV128, &tile_full[tb], a0 + ATOM, 2 * d + h * 64, row0 + j * 128);  // V tile
          for (int c = 0; c < nc; ++c, ++g) {
            const int st = g % RING;
            mbar_wait(&ring_empty[st], (uint32_t)(((g / RING) & 1) ^ 1));
            if (p.trace && blockIdx.x < 4 && g < 48) p.trace[(blockIdx.x * 12 + 8) * 64 + g] = clock64();
            uint8_t* dst = sRing + st * 2 * CH;
            mbar_arrive_expect_tx(&ring_full[st], 2 * CH);
            tma_load_2d(&tmQKV64, &ring_full[st], dst, h * 64, row0 + c * 64);        // Q_c
            tma_load_2d(&tmDO64, &ring_full[st], dst + CH, h * 64, row0 + c * 64);     // dO_c
          }
        }
      }
    }
  } else if (warp == P_WORKERS + 1) {
    // ======================= score issuer: S^T_c = K_j Q_c^T, dP^T_c = V_j dO_c^T =======================
    if (elect_one()) {
      const uint32_t uA = smem_u32(sA), uRing = smem_u32(sRing);
      int nt = 0;
      int g = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        for (int j = 0; j < ntile; ++j, ++nt) {
          const int tb = nt & 1;
          mbar_wait(&tile_full[tb], (nt >> 1) & 1);
          const uint64_t a0 = desc_k(uA + tb * 2 * ATOM), a1 = desc_k(uA + tb * 2 * ATOM + ATOM);
          for (int c = 0; c < nc; ++c, ++g) {
            const int st = g % RING, sb = g & 1;
            const int wc = min(64, S_pad - c * 64);
            mbar_wait(&ring_full[st], (uint32_t)((g / RING) & 1));
            if (p.trace && blockIdx.x < 4 && g < 24) p.trace[(blockIdx.x * 12 + 9) * 64 + 2 * g] = clock64();
            mbar_wait(&sdp_empty[sb], (uint32_t)(((g >> 1) & 1) ^ 1));
            tc_fence_after();
            if (p.trace && blockIdx.x < 4 && g < 24) p.trace[(blockIdx.x * 12 + 9) * 64 + 2 * g + 1] = clock64();
            const uint32_t id = idesc_rt(wc, false, false);
            const uint32_t ub = uRing + st * 2 * CH;
            const uint64_t b0 = desc_k(ub), b1 = desc_k(ub + CH);
            const uint32_t tS = tmem + sb * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tS, a0 + 2 * k, b0 + 2 * k, id, k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tS + 64, a1 + 2 * k, b1 + 2 * k, id, k > 0);
            umma_commit(&sdp_full[sb]);
          }
          umma_commit(&tile_empty[tb]);   // this thread's MMAs on the tile buffer (1 of 2 arrivals)
        }
      }
    }
  } else if (warp == P_WORKERS + 2) {
    // ======================= accumulate issuer: dV += P^T_c dO_c ; dK += dS^T_c Q_c ; dQ_pair += dS_pair K_j ========
    if (elect_one()) {
      const uint32_t uA = smem_u32(sA), uRing = smem_u32(sRing), uDS = smem_u32(sDS), uPT = smem_u32(sPT);
      const uint32_t id = idesc_rt(64, false, true);      // A K-major (P^T / dS^T rows = keys), B MN-major
      const uint32_t idq = idesc_rt(64, true, true);      // A MN-major (dS^T read as dS), B MN-major (K_j)
      int n = 0, nt = 0;
      int g = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
        for (int j = 0; j < ntile; ++j, ++nt) {
          const int tb = nt & 1;
          const int kk = min(8, (min(S - j * 128, 128) + 15) >> 4);   // 16-key steps of this tile that hold real keys
          mbar_wait(&acc_empty[0], (uint32_t)((nt & 1) ^ 1));          // dV / dK of the previous key tile read out
          if (j == 0) mbar_wait(&dq_empty[0], (uint32_t)((n & 1) ^ 1)); // dQ of the previous work item read out
          mbar_wait(&tile_full[tb], (nt >> 1) & 1);                    // K_j (B operand of the dQ MMAs) landed
          const uint32_t uK = uA + tb * 2 * ATOM;
          for (int c = 0; c < nc; ++c, ++g) {
            const int st = g % RING, sb = g & 1;
            const int wc = min(64, S_pad - c * 64);
            mbar_wait(&ds_full[sb], (uint32_t)((g >> 1) & 1));
            mbar_wait(&ring_full[st], (uint32_t)((g / RING) & 1));  // long complete; acquires the TMA writes for this thread
            tc_fence_after();
            if (p.trace && blockIdx.x < 4 && g < 24) p.trace[(blockIdx.x * 12 + 10) * 64 + 2 * g] = clock64();
            const uint32_t ub = uRing + st * 2 * CH;
            const int ks = wc >> 4;
            {   // descriptor bases once per chunk; a 16-key / 16-query step advances the start address by 32 B (K-major)
                // or 2048 B (MN-major) = +2 / +128 in the descriptor's (address >> 4) field
              const uint64_t aPT = desc_k(uPT + sb * ATOM), bDO = desc_mn(ub + CH);
              const uint64_t aDS = desc_k(uDS + c * ATOM), bQ = desc_mn(ub);
              for (int k = 0; k < ks; ++k) umma_bf16(tmem + 256, aPT + 2 * k, bDO + 128 * k, id, (c > 0 || k > 0));
              for (int k = 0; k < ks; ++k) umma_bf16(tmem + 320, aDS + 2 * k, bQ + 128 * k, id, (c > 0 || k > 0));
            }
            umma_commit(&pt_empty[sb]);
            umma_commit(&ring_empty[st]);
            if (c == nc - 1) umma_commit(&acc_full[0]);   // dV / dK of the key tile are final: before the dQ MMAs below
            if ((c & 1) || c == nc - 1) {   // the pair (2q, 2q+1) of dS^T chunks is complete: dQ block q += dS K_j
              const int q = c >> 1;
              const uint64_t aQ = desc_mn_a(uDS + 2 * q * ATOM), bK = desc_mn(uK);
              for (int k = 0; k < kk; ++k) umma_bf16(tmem + 384 + q * 64, aQ + 128 * k, bK + 128 * k, idq, (j > 0 || k > 0));
              umma_commit(&dsb_empty[2 * q]);
              if (2 * q + 1 < nc) umma_commit(&dsb_empty[2 * q + 1]);
            }
            if (c == nc - 1) umma_commit(&tile_empty[tb]);   // 2 of 2 arrivals: K_j / V_j may be overwritten
            if (p.trace && blockIdx.x < 4 && g < 24) p.trace[(blockIdx.x * 12 + 10) * 64 + 2 * g + 1] = clock64();
          }
        }
        umma_commit(&dq_full[0]);
      }
    }
  } else if (warp >= P_WORKERS + 3) {
    // ======================= statistics warps: sLD[n & 1] = {LSE (log2 units) [256], D [256]} of the work item ========
    const int t = (warp - (P_WORKERS + 3)) * 32 + lane;   // 0 .. 63
    int n = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
      const int h = w % p.H, b = w / p.H;
      const int row0 = b * S;
      float* wL = sLD + (n & 1) * 512;
      const long long sbase = ((long long)b * p.H + h) * S;
      mbar_wait(&stat_empty[n & 1], ((n >> 1) & 1) ^ 1);
      for (int qi = t; qi < 256; qi += NSTAT * 32) {
        float acc = 0.f, L = 0.f;
        if (qi < S) {
          const uint4* po = reinterpret_cast<const uint4*>(p.o_in + (long long)(row0 + qi) * d + h * 64);
          const uint4* pd = reinterpret_cast<const uint4*>(p.dout + (long long)(row0 + qi) * d + h * 64);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const uint4 a = __ldg(po + jj), c = __ldg(pd + jj);
            acc += bf16_lo(a.x) * bf16_lo(c.x) + bf16_hi(a.x) * bf16_hi(c.x) + bf16_lo(a.y) * bf16_lo(c.y) +
                   bf16_hi(a.y) * bf16_hi(c.y) + bf16_lo(a.z) * bf16_lo(c.z) + bf16_hi(a.z) * bf16_hi(c.z) +
                   bf16_lo(a.w) * bf16_lo(c.w) + bf16_hi(a.w) * bf16_hi(c.w);
          }
          L = p.lse[sbase + qi] * 1.4426950408889634f;
        }
        wL[qi] = -L;                     // negated: the workers' packed FFMA2 adds them
        wL[256 + qi] = -acc * p.scale;   // -D * scale:  dS = P * (dP * scale - D * scale)
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&stat_full[n & 1]);
    }
  } else {
    // ======================= 8 worker warps: thread == key row x column half =======================
    const int q4 = warp & 3, grp = warp >> 2;
    const int r = q4 * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q4 * 32) << 16);
    const long long ld = 3LL * d;
    int n = 0, nt = 0;
    int g = 0;
    // Output tiles (dV, dK per key tile; dQ per work item) leave through smem + ONE TMA store each (3-D box clipped at the
    // sequence length) instead of one 64 B piece of a row per thread (32 L1 wavefronts per st.global.v4: 2-3.8 k clocks
    // per epilogue in the phase trace).  Staging reuses operand buffers that are dead at that point: P^T[0|1] for dV | dK
    // (acc_full: every MMA of the tile has completed), dS^T[0|1] for the dQ blocks (dq_full).  Before the workers write
    // those buffers again the elected thread waits for the bulk stores to have read them (store_sync).
    bool store_pending = false;
    auto store_sync = [&]() {
      if (store_pending) {
        if (warp == 0 && elect_one()) tma_store_wait_read<0>();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        store_pending = false;
      }
    };
    auto stage32 = [&](uint8_t* buf, const uint32_t (&v)[32]) {   // this thread's 32 fp32 -> 64 B of row r, swizzled
      uint8_t* a = buf + r * 128;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[jj * 8 + 0]), __uint_as_float(v[jj * 8 + 1]));
        o.y = pack_bf16x2(__uint_as_float(v[jj * 8 + 2]), __uint_as_float(v[jj * 8 + 3]));
        o.z = pack_bf16x2(__uint_as_float(v[jj * 8 + 4]), __uint_as_float(v[jj * 8 + 5]));
        o.w = pack_bf16x2(__uint_as_float(v[jj * 8 + 6]), __uint_as_float(v[jj * 8 + 7]));
        *reinterpret_cast<uint4*>(a + (((grp * 4 + jj) ^ (r & 7)) << 4)) = o;
      }
    };
    const bool tr = p.trace != nullptr && blockIdx.x < 4 && lane == 0;
    unsigned long long* trw = p.trace + (blockIdx.x * 12 + warp) * 64;
    long long tq = tr ? clock64() : 0, a_stat = 0, a_sdp = 0, a_ld = 0, a_buf = 0, a_cmp = 0, a_accw = 0, a_epi = 0, a_dqw = 0,
              a_dqe = 0;
#define BT(acc) do { if (tr) { const long long tn_ = clock64(); acc += tn_ - tq; tq = tn_; } } while (0)
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
      const int h = w % p.H, b = w / p.H;
      const int row0 = b * S;
      const float* sL = sLD + (n & 1) * 512;
      const float* sD = sL + 256;
      mbar_wait(&stat_full[n & 1], (n >> 1) & 1);
      BT(a_stat);
      for (int j = 0; j < ntile; ++j, ++nt) {
        const int ri = j * 128 + r;   // key index of this thread's row
        // a key that is out of range or padded out by the key mask contributes P = dS = 0 (its dK / dV rows stay zero)
        const bool key_dead = (ri >= S) || (p.kmask != nullptr && p.kmask[(long long)b * S + ri] == 0);
        for (int c = 0; c < nc; ++c, ++g) {
          const int sb = g & 1;
          const int wc = min(64, S_pad - c * 64);
          mbar_wait(&sdp_full[sb], (uint32_t)((g >> 1) & 1));
          tc_fence_after();
          BT(a_sdp);
          if (tr && g < 24) trw[16 + 2 * g] = tq;
          uint32_t sv[32], dv[32];
          tmem_ld32(trow + sb * 128 + grp * 32, sv);
          tmem_ld32(trow + sb * 128 + 64 + grp * 32, dv);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sdp_empty[sb]);
          BT(a_ld);
          mbar_wait(&pt_empty[sb], (uint32_t)(((g >> 1) & 1) ^ 1));   // dV / dK MMAs of chunk g-2 have left P^T[sb]
          mbar_wait(&dsb_empty[c], (uint32_t)((nt & 1) ^ 1));          // the previous key tile's dQ MMA has left dS^T[c]
          store_sync();                                                // ... and the last epilogue's bulk stores the staging
          BT(a_buf);
          uint8_t* myDS = sDS + c * ATOM;
          uint8_t* myPT = sPT + sb * ATOM;
          const int cbase = c * 64 + grp * 32;
          const bool full = !key_dead && (cbase + 32 <= S) && (!CAUSAL || ri <= cbase);
          const uint64_t c2 = pk2(p.scale_log2, p.scale_log2), sc2 = pk2(p.scale, p.scale);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (grp * 32 + half * 16 >= wc) continue;   // columns past S_pad (last chunk): nothing reads them (uniform)
            uint8_t* aDS = myDS + ((grp * 32 + half * 16) >> 6) * ATOM + r * 128;
            uint8_t* aPT = myPT + ((grp * 32 + half * 16) >> 6) * ATOM + r * 128;
            const int c8 = ((grp * 32 + half * 16) & 63) >> 3;
            const int o0 = (c8 ^ (r & 7)) << 4, o1 = ((c8 + 1) ^ (r & 7)) << 4;
            if (key_dead) {
              *reinterpret_cast<uint4*>(aDS + o0) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(aDS + o1) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(aPT + o0) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(aPT + o1) = make_uint4(0, 0, 0, 0);
            } else if (full) {
              // packed f32x2 math: p = 2^(s c - L), dS = p (dP scale - D scale); sL / sD hold -L and -D scale
              const float4* pl = reinterpret_cast<const float4*>(sL + cbase + half * 16);
              const float4* pd = reinterpret_cast<const float4*>(sD + cbase + half * 16);
              uint32_t kp[8], kd[8];
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const float4 l4 = pl[e4], d4 = pd[e4];
                const int e = half * 16 + e4 * 4;
                float x0, x1, x2, x3, g0, g1, g2, g3;
                upk2(ffma2(pk2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), c2, pk2(l4.x, l4.y)), x0, x1);
                upk2(ffma2(pk2(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])), c2, pk2(l4.z, l4.w)), x2, x3);
                x0 = ex2_approx(x0); x1 = ex2_approx(x1); x2 = ex2_approx(x2); x3 = ex2_approx(x3);
                const uint64_t t01 = ffma2(pk2(__uint_as_float(dv[e]), __uint_as_float(dv[e + 1])), sc2, pk2(d4.x, d4.y));
                const uint64_t t23 = ffma2(pk2(__uint_as_float(dv[e + 2]), __uint_as_float(dv[e + 3])), sc2, pk2(d4.z, d4.w));
                upk2(fmul2(pk2(x0, x1), t01), g0, g1);
                upk2(fmul2(pk2(x2, x3), t23), g2, g3);
                kp[e4 * 2] = pack_bf16x2(x0, x1); kp[e4 * 2 + 1] = pack_bf16x2(x2, x3);
                kd[e4 * 2] = pack_bf16x2(g0, g1); kd[e4 * 2 + 1] = pack_bf16x2(g2, g3);
              }
              *reinterpret_cast<uint4*>(aDS + o0) = make_uint4(kd[0], kd[1], kd[2], kd[3]);
              *reinterpret_cast<uint4*>(aDS + o1) = make_uint4(kd[4], kd[5], kd[6], kd[7]);
              *reinterpret_cast<uint4*>(aPT + o0) = make_uint4(kp[0], kp[1], kp[2], kp[3]);
              *reinterpret_cast<uint4*>(aPT + o1) = make_uint4(kp[4], kp[5], kp[6], kp[7]);
            } else {
              float ds[16], pt[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int cj = cbase + half * 16 + e;
                const bool valid = (cj < S) && (!CAUSAL || ri <= cj);
                const float nL = sL[cj & 255], nDs = sD[cj & 255];
                const float pv = valid ? ex2_approx(fmaf(__uint_as_float(sv[half * 16 + e]), p.scale_log2, nL)) : 0.f;
                pt[e] = pv;
                ds[e] = pv * fmaf(__uint_as_float(dv[half * 16 + e]), p.scale, nDs);
              }
              store_p16(myDS, r, grp * 32 + half * 16, ds);
              store_p16(myPT, r, grp * 32 + half * 16, pt);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ds_full[sb]);
          BT(a_cmp);
          if (tr && g < 24) trw[17 + 2 * g] = tq;
        }
        // ---- key-tile epilogue: dV (V block, 2d) and dK (K block, d) of rows ri -> bf16 -> dqkv ----
        mbar_wait(&acc_full[0], (uint32_t)(nt & 1));
        tc_fence_after();
        BT(a_accw);
        if (tr && warp == 0 && nt < 12) p.trace[(blockIdx.x * 12 + 11) * 64 + nt] = tq;
        uint32_t v0[32], v1[32];
        tmem_ld32(trow + 256 + grp * 32, v0);
        tmem_ld32(trow + 320 + grp * 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[0]);
        store_sync();
        stage32(sPT, v0);
        stage32(sPT + ATOM, v1);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 0 && elect_one()) {
          tma_store_3d(&tmDQKV, sPT, 2 * d + h * 64, j * 128, b);          // dV rows of key tile j (V block of dqkv)
          tma_store_3d(&tmDQKV, sPT + ATOM, d + h * 64, j * 128, b);       // dK rows (K block)
          tma_store_commit();
        }
        store_pending = true;
        if (tr) trw[10] += clock64() - tq;   // first-tile epilogues (the per-item BT(a_epi) only sees the last one)
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&stat_empty[n & 1]);   // this warp no longer reads sLD[n & 1]
      BT(a_epi);
      // ---- work-item epilogue: dQ blocks (TMEM lane = query within the block) -> bf16 -> Q block of dqkv ----
      mbar_wait(&dq_full[0], (uint32_t)(n & 1));
      tc_fence_after();
      BT(a_dqw);
      {
        uint32_t v0[32], v1[32];
        tmem_ld32(trow + 384 + grp * 32, v0);
        if (nc > 2) tmem_ld32(trow + 448 + grp * 32, v1);   // uniform
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dq_empty[0]);
        store_sync();
        stage32(sDS, v0);
        if (nc > 2) stage32(sDS + ATOM, v1);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 0 && elect_one()) {
          tma_store_3d(&tmDQKV, sDS, h * 64, 0, b);                        // dQ, queries 0 .. 127 (Q block of dqkv)
          if (nc > 2) tma_store_3d(&tmDQKV, sDS + ATOM, h * 64, 128, b);   // queries 128 ..
          tma_store_commit();
        }
        store_pending = true;
      }
      BT(a_dqe);
    }
    if (warp == 0 && elect_one()) tma_store_wait_all<0>();   // smem must outlive the last bulk stores' reads
    if (tr) {
      trw[0] = a_stat; trw[1] = a_sdp; trw[2] = a_ld; trw[3] = a_buf; trw[4] = a_cmp; trw[5] = a_accw; trw[6] = a_epi;
      trw[7] = a_dqw; trw[8] = a_dqe; trw[9] = n;
    }
#undef BT
  }
  tc_fence_before();
  __syncthreads();
  if (warp == P_WORKERS) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
constexpr int BWDF_SMEM = 1024 + 4 * ATOM + 3 * 16384 + 4 * ATOM + 2 * ATOM + 4096 + 512;   // 213.5 KB
#ifndef MMB_ATTN_BWD_DEFAULT
#define MMB_ATTN_BWD_DEFAULT 2   /* 1 = column-split two-pass, 2 = fused single pass (S <= 256): 0.93 vs 1.88 ms at B/16 */
#endif

// ------------------------------------------------------------------------------------------------
// Backward, persistent PING-PONG version (round 2; default).  Same producer / two issuers / statistics warps as
// attn_bwd_persist_kernel, but the eight worker warps form TWO GROUPS that serve ALTERNATE chunks (group = parity of
// the running chunk counter) instead of splitting every chunk's columns.  In the column-split kernel all workers hit
// the same fixed latencies at the same time (mbarrier check ~90-250 clk, tcgen05.ld ~250, buffer check ~220,
// fence.proxy.async + arrive ~240: ~960 of the ~1900 clocks of a chunk, profiles/r1_attn_bwd_phase_trace.txt) and the
// SM idles; here each sub-partition holds one warp of each group, so one group's waits run under the other's
// exp / FMA / pack work.  A thread owns one tile row and all 64 columns of its chunk (two 32-column halves).
// The epilogue of tile n (TMEM accumulators -> bf16 -> dqkv) is deferred until after the group's first chunk of tile
// n+1, so a group never waits for the other group's last chunk.  Sequence lengths up to SMAX = 384 (3 row tiles, 6
// chunks): CLIP ViT-L/14 with CLS (257) and the FLAVA multimodal encoder (275) run here too.
// ------------------------------------------------------------------------------------------------
constexpr int SMAX = 384;
constexpr int PP_STATS = 1;   // one statistics warp (12 warps per CTA: register allocation is per 4 warps)
template <bool CAUSAL, bool DKDV>
__global__ void __launch_bounds__((8 + 3 + PP_STATS) * 32, 1)   // 12 warps -> 168 registers per thread available
attn_bwd_pp_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                   const __grid_constant__ CUtensorMap tmDO128, const __grid_constant__ CUtensorMap tmDO64,
                   const AttnTcArgs p, const int n_work) {
  // Two chunks are in flight at once (one per worker group) and each takes a group ~2x as long as in the column-split
  // kernel, so the chunk-operand ring must hold the two active chunks PLUS the prefetch distance that covers the
  // TMA latency: 5 / 6 stages (round-2 measurement: with 3 / 4 the groups starved on ring_full, 2.76 vs 1.88 ms).
  constexpr int RING = DKDV ? 5 : 6;
  constexpr int CH = 8192;  // one 64-row x 128 B chunk operand
  constexpr int P_WORKERS = 8;
  constexpr int SLD = 2 * SMAX;  // floats per statistics buffer: lse(log2) [SMAX] | rowsum(dO*O) [SMAX]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // [2 tile buffers][A0 16 KB | A1 16 KB]
  uint8_t* sRing = sA + 4 * ATOM;                       // RING x (B0_c 8 KB | B1_c 8 KB)
  uint8_t* sDS = sRing + RING * 2 * CH;                 // [2] dS chunk (16 KB each)
  uint8_t* sPT = sDS + 2 * ATOM;                        // [2] P^T chunk (DKDV only)
  float* sLD = reinterpret_cast<float*>(sPT + (DKDV ? 2 : 0) * ATOM);  // [2][SLD]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLD + 2 * SLD);
  uint64_t* tile_full = bars;            // [2]  tile operands landed                 producer(TMA) -> score issuer
  uint64_t* tile_empty = bars + 2;       // [2]  last score MMA of the tile complete  score issuer  -> producer
  uint64_t* ring_full = bars + 4;        // [RING <= 8]
  uint64_t* ring_empty = bars + 12;      // [RING <= 8] accumulate MMAs of the chunk complete       -> producer
  uint64_t* sdp_full = bars + 20;        // [2]  S_c/dP_c in TMEM                     score issuer  -> worker group
  uint64_t* sdp_empty = bars + 22;       // [2]  the group has read them (4 warps)                  -> score issuer
  uint64_t* ds_full = bars + 24;         // [2]  dS_c (P^T_c) in smem (4 warps)       worker group  -> acc issuer
  uint64_t* ds_empty = bars + 26;        // [2]  accumulate MMAs done with them       acc issuer    -> worker group
  uint64_t* acc_full = bars + 28;        // [2]  accumulators of the tile final       acc issuer    -> workers
  uint64_t* acc_empty = bars + 30;       // [2]  workers have read them (8 warps)                   -> acc issuer
  uint64_t* stat_full = bars + 32;       // [2]  LSE / D of the tile in sLD (1 warp)  stats warp    -> workers
  uint64_t* stat_empty = bars + 34;      // [2]  workers are done with them (8 warps)               -> stats warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int S = p.S, S_pad = p.S_pad, d = p.H * 64;
  const int nc = (S_pad + 63) >> 6;
  const int ntile = (S + 127) >> 7;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV128); tma_prefetch_desc(&tmQKV64);
    tma_prefetch_desc(&tmDO128); tma_prefetch_desc(&tmDO64);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tile_full[i], 1); mbar_init(&tile_empty[i], 1);
      mbar_init(&sdp_full[i], 1); mbar_init(&sdp_empty[i], P_WORKERS / 2);
      mbar_init(&ds_full[i], P_WORKERS / 2); mbar_init(&ds_empty[i], 1);
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], P_WORKERS);
      mbar_init(&stat_full[i], PP_STATS); mbar_init(&stat_empty[i], P_WORKERS);
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
    if (elect_one()) {
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
    if (elect_one()) {
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
          mbar_wait(&ring_full[st], (uint32_t)((g / RING) & 1));
          mbar_wait(&sdp_empty[sb], (uint32_t)(((g >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t id = idesc_rt(wc, false, false);
          const uint32_t ub = uRing + st * 2 * CH;
          const uint64_t b0 = desc_k(ub), b1 = desc_k(ub + CH);
          const uint32_t tS = tmem + sb * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS, a0 + 2 * k, b0 + 2 * k, id, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS + 64, a1 + 2 * k, b1 + 2 * k, id, k > 0);
          umma_commit(&sdp_full[sb]);
        }
        umma_commit(&tile_empty[tb]);   // every MMA that reads this tile buffer has been issued before this commit
      }
    }
  } else if (warp == P_WORKERS + 2) {
    // ======================= accumulate issuer: dQ += dS_c K_c  |  dV += P^T_c dO_c ; dK += dS^T_c Q_c ==============
    if (elect_one()) {
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
          mbar_wait(&ds_full[sb], (uint32_t)((g >> 1) & 1));
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
        }
      }
    }
  } else if (warp >= P_WORKERS + 3) {
    // ======================= statistics warps: sLD[n & 1] = {LSE (log2 units) [SMAX], D [SMAX]} of tile n ==========
    const int t = (warp - (P_WORKERS + 3)) * 32 + lane;   // 0 .. 32*PP_STATS-1
    int n = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
      const int tile = w % ntile, bh = w / ntile;
      const int h = bh % p.H, b = bh / p.H;
      const int row0 = b * S;
      float* wL = sLD + (n & 1) * SLD;
      const long long sbase = ((long long)b * p.H + h) * S;
      mbar_wait(&stat_empty[n & 1], ((n >> 1) & 1) ^ 1);
      if (!DKDV) {
        // one warp covers the tile's 128 rows: 4 rows per lane, processed two at a time with all 32 16-byte loads of
        // the pair in flight before the first FMA (a row-by-row loop chains four global-load latencies per tile)
#pragma unroll
        for (int k = 0; k < 128 / (PP_STATS * 32); k += 2) {
          uint4 va[2][8], vc[2][8];
          float Lr[2] = {0.f, 0.f};
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int ri = tile * 128 + t + (k + u) * (PP_STATS * 32);
            const int rc = min(ri, S - 1);      // clamped: loads always legal, results of padding rows discarded
            const uint4* po = reinterpret_cast<const uint4*>(p.o_in + (long long)(row0 + rc) * d + h * 64);
            const uint4* pd = reinterpret_cast<const uint4*>(p.dout + (long long)(row0 + rc) * d + h * 64);
#pragma unroll
            for (int j = 0; j < 8; ++j) { va[u][j] = __ldg(po + j); vc[u][j] = __ldg(pd + j); }
            Lr[u] = __ldg(p.lse + sbase + rc) * 1.4426950408889634f;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int rr = t + (k + u) * (PP_STATS * 32);
            const int ri = tile * 128 + rr;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 a = va[u][j], c = vc[u][j];
              acc += bf16_lo(a.x) * bf16_lo(c.x) + bf16_hi(a.x) * bf16_hi(c.x) + bf16_lo(a.y) * bf16_lo(c.y) +
                     bf16_hi(a.y) * bf16_hi(c.y) + bf16_lo(a.z) * bf16_lo(c.z) + bf16_hi(a.z) * bf16_hi(c.z) +
                     bf16_lo(a.w) * bf16_lo(c.w) + bf16_hi(a.w) * bf16_hi(c.w);
            }
            if (ri < S) p.dsum[sbase + ri] = acc;
            wL[rr] = ri < S ? Lr[u] : 0.f;
            wL[SMAX + rr] = ri < S ? acc : 0.f;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < SMAX / (PP_STATS * 32); ++k) {
          const int qi = t + k * (PP_STATS * 32);
          const bool ok = qi < S;
          wL[qi] = ok ? p.lse[sbase + qi] * 1.4426950408889634f : 0.f;
          wL[SMAX + qi] = ok ? p.dsum[sbase + qi] : 0.f;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&stat_full[n & 1]);   // release: the smem writes above are visible to the waiters
    }
  } else {
    // ======================= 8 worker warps: group = chunk parity, thread == tile row =======================
    const int q4 = warp & 3, grp = warp >> 2;
    const int r = q4 * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q4 * 32) << 16);
    const long long ld = 3LL * d;
    // deferred tile epilogue: accumulators (TMEM buffer pn & 1) -> bf16 -> dqkv; this group converts columns
    // [grp*32, grp*32+32) of every accumulator
    int pn = -1, p_row0 = 0, p_ri = 0, p_h = 0;
    auto tile_epilogue = [&]() {
      const int ab = pn & 1;
      mbar_wait(&acc_full[ab], (pn >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(trow + 256 + ab * 128 + grp * 32, v0);
      if (DKDV) tmem_ld32(trow + 256 + ab * 128 + 64 + grp * 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
      if (p_ri < S) {
#pragma unroll
        for (int which = 0; which < (DKDV ? 2 : 1); ++which) {
          // DQ: acc0 -> Q block.  DKDV: acc0 = dV -> V block (2d), acc1 = dK -> K block (d)
          const uint32_t* v = which == 0 ? v0 : v1;
          const int coff = !DKDV ? 0 : (which == 0 ? 2 * d : d);
          __nv_bfloat16* dst = p.dqkv + (long long)(p_row0 + p_ri) * ld + coff + p_h * 64 + grp * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
            o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
            o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
            o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
            reinterpret_cast<uint4*>(dst)[j] = o;
          }
        }
      }
      pn = -1;
    };
    int n = 0;
    long long g = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++n) {
      const int tile = w % ntile, bh = w / ntile;
      const int h = bh % p.H, b = bh / p.H;
      const int row0 = b * S;
      const int ri = tile * 128 + r;  // global index of this thread's row (query for DQ, key for DKDV)
      float Lrow = 0.f, Drow = 0.f;
      const float* sL = sLD + (n & 1) * SLD;
      const float* sD = sL + SMAX;
      mbar_wait(&stat_full[n & 1], (n >> 1) & 1);
      if (!DKDV) {
        Lrow = sL[r];
        Drow = sD[r];
      }
      const float nD = -Drow * p.scale;
      for (int c = 0; c < nc; ++c) {
        const long long gc = g + c;
        if ((int)(gc & 1) != grp) continue;
        const int sb = grp;
        const uint32_t par = (uint32_t)((gc >> 1) & 1);
        const int wc = min(64, S_pad - c * 64);
        mbar_wait(&sdp_full[sb], par);
        tc_fence_after();
        uint8_t* myDS = sDS + sb * ATOM;
        uint8_t* myPT = sPT + sb * ATOM;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t sv[32], dv[32];
          tmem_ld32(trow + sb * 128 + hf * 32, sv);
          tmem_ld32(trow + sb * 128 + 64 + hf * 32, dv);
          tmem_ld_wait();
          if (hf == 1) {   // S_c / dP_c of this chunk are in registers -> the score issuer may overwrite the buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sdp_empty[sb]);
          } else {
            mbar_wait(&ds_empty[sb], par ^ 1);   // accumulate MMAs of chunk gc-2 have left the dS / P^T buffers
          }
          const int cbase = c * 64 + hf * 32;   // first global column (key for DQ, query for DKDV) of this half
          // interior fast path: the whole 32-column strip is unmasked for this row
          const bool full = (ri < S) && (cbase + 32 <= S) && (!CAUSAL || (DKDV ? (ri <= cbase) : (cbase + 31 <= ri)));
#pragma unroll
          for (int qt = 0; qt < 2; ++qt) {
            float ds[16], pt[16];
            if (ri >= S) {
#pragma unroll
              for (int e = 0; e < 16; ++e) { ds[e] = 0.f; pt[e] = 0.f; }
            } else if (full) {
              if (!DKDV) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float pv = ex2_approx(fmaf(__uint_as_float(sv[qt * 16 + e]), p.scale_log2, -Lrow));
                  ds[e] = pv * fmaf(__uint_as_float(dv[qt * 16 + e]), p.scale, nD);
                }
              } else {
                const float4* pl = reinterpret_cast<const float4*>(sL + cbase + qt * 16);
                const float4* pd = reinterpret_cast<const float4*>(sD + cbase + qt * 16);
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                  const float4 l4 = pl[e4], d4 = pd[e4];
                  const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const int e = e4 * 4 + k;
                    const float pv = ex2_approx(fmaf(__uint_as_float(sv[qt * 16 + e]), p.scale_log2, -ls[k]));
                    pt[e] = pv;
                    ds[e] = pv * (__uint_as_float(dv[qt * 16 + e]) - dd[k]) * p.scale;
                  }
                }
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int cj = cbase + qt * 16 + e;
                bool valid;
                float L, Dv;
                if (!DKDV) {
                  valid = (cj < S) && (!CAUSAL || cj <= ri);
                  L = Lrow; Dv = Drow;
                } else {
                  valid = (cj < S) && (!CAUSAL || ri <= cj);
                  const int cq = min(cj, SMAX - 1);
                  L = sL[cq]; Dv = sD[cq];
                }
                const float pv = valid ? ex2_approx(__uint_as_float(sv[qt * 16 + e]) * p.scale_log2 - L) : 0.f;
                pt[e] = pv;
                ds[e] = pv * (__uint_as_float(dv[qt * 16 + e]) - Dv) * p.scale;
              }
            }
            if (hf * 32 + qt * 16 < wc) {
              store_p16(myDS, r, hf * 32 + qt * 16, ds);
              if (DKDV) store_p16(myPT, r, hf * 32 + qt * 16, pt);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_full[sb]);
        if (pn >= 0) tile_epilogue();   // previous tile's accumulators: long final by now
      }
      g += nc;
      __syncwarp();
      if (lane == 0) mbar_arrive(&stat_empty[n & 1]);   // this warp no longer reads sLD[n & 1]
      if (pn >= 0) tile_epilogue();                      // (this group served no chunk of tile n)
      pn = n; p_row0 = row0; p_ri = ri; p_h = h;
    }
    if (pn >= 0) tile_epilogue();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == P_WORKERS) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

constexpr int BWDPP_DQ_SMEM = 1024 + 4 * ATOM + 6 * 16384 + 2 * ATOM + 2 * 2 * SMAX * 4 + 512;     // 203.5 KB
constexpr int BWDPP_DKDV_SMEM = 1024 + 4 * ATOM + 5 * 16384 + 4 * ATOM + 2 * 2 * SMAX * 4 + 512;   // 219.5 KB

constexpr int BWDP_DQ_SMEM = 1024 + 4 * ATOM + 4 * 16384 + 2 * ATOM + 4096 + 512;              // 169.5 KB
constexpr int BWDP_DKDV_SMEM = 1024 + 4 * ATOM + 3 * 16384 + 4 * ATOM + 4096 + 512;            // 185.5 KB

constexpr int FWD_SMEM = 1024 + 6 * ATOM + 2048 + 384 + 64;
constexpr int FWD_SMEM_BIG = 1024 + 9 * ATOM + 2048 + 384 + 64;

}  // namespace mmb

using namespace mmb;

// debug (not in include/mmb200.h): device buffer [4][12][64] receiving per-phase SM-clock totals of the item forward
// kernel's first four CTAs (scripts/attn_item_trace.py); nullptr = off
static unsigned long long* g_item_trace = nullptr;
extern "C" int mmb_debug_attn_item_trace(void* p) { g_item_trace = (unsigned long long*)p; return 0; }
#ifdef MMB_ATTN_TRACE
extern "C" int mmb_debug_attn_trace(void* p) { return (int)cudaMemcpyToSymbol(g_attn_trace, &p, sizeof(p)); }
#endif

extern "C" int mmb_attention_fwd_tc(const void* qkv, void* out, float* lse, int B, int S, int H, int causal, float scale,
                                    void* stream);
extern "C" int mmb_attention_fwd_kmask(const void* qkv, void* out, float* lse, const unsigned char* kmask, int B, int S,
                                       int H, int head_dim, int causal, float scale, void* stream);
static int attention_fwd_tc_impl(const void* qkv, void* out, float* lse, const uint8_t* kmask, int B, int S, int H,
                                 int causal, float scale, void* stream) {
  if (B <= 0 || S <= 0 || S > SMAX) return MMB_ERR_UNSUPPORTED;
  const int d = H * 64, S_pad = (S + 15) & ~15;
  const bool big = S_pad > 256;
  CUtensorMap tm128, tmPad, tmRem;
  int rc = make_tmap_2d(&tm128, qkv, 2, false, 3ull * d, (uint64_t)B * S, 3ull * d * 2, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmPad, qkv, 2, false, 3ull * d, (uint64_t)B * S, 3ull * d * 2, 64, big ? 256 : S_pad);
  if (rc) return rc;
  tmRem = tmPad;
  if (big) {
    rc = make_tmap_2d(&tmRem, qkv, 2, false, 3ull * d, (uint64_t)B * S, 3ull * d * 2, 64, S_pad - 256);
    if (rc) return rc;
  }
  const int smem_bytes = big ? FWD_SMEM_BIG : FWD_SMEM;
  // Forward kernel choice (measured, B = 1024): single-tile sequences (S <= 128: the text tower, 0.170 vs 0.238 ms) run
  // on the persistent ping-pong kernel; two-tile sequences without a mask (the image towers, S = 197) on the item kernel
  // (0.387 ms vs 0.618 tile / 0.606 ping-pong); everything else (masks, causal S > 128, 256 < S <= 384) on the
  // one-tile-per-CTA kernel.  MMB_ATTN_FWD=tile / =pp / =item force one where it applies (A/B testing).
  static int fwd_variant = -1;
  if (fwd_variant < 0) {
    const char* e = getenv("MMB_ATTN_FWD");
    fwd_variant = (e && e[0] == 't') ? 1 : (e && e[0] == 'p') ? 0 : (e && e[0] == 'i') ? 3 : 2;
  }
  if ((fwd_variant == 3 || fwd_variant == 2) && S > 128 && S <= 256 && !causal && !kmask) {
    AttnTcArgs a{};
    a.S = S; a.H = H; a.S_pad = S_pad; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
    a.lse = lse; a.out = (__nv_bfloat16*)out;
    const int n_items = H * B;
    const int grid_i = n_items < num_sms() ? n_items : num_sms();
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaFuncSetAttribute(attn_fwd_item_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIT_SMEM);
    static int stagger = -1;
    if (stagger < 0) {
      const char* e2 = getenv("MMB_ATTN_ITEM_STAGGER");
      stagger = (e2 && e2[0] == '0') ? 0 : 1;
    }
    CUtensorMap tmOut;   // [B][S][d] view of out: the per-tile store box is clipped at S (rows of the next batch stay intact)
    rc = make_tmap_3d_bf16(&tmOut, out, (uint64_t)d, (uint64_t)S, (uint64_t)B, (uint64_t)d * 2, (uint64_t)S * d * 2, 64, 128);
    if (rc) return rc;
    if (g_item_trace) {
      cudaFuncSetAttribute(attn_fwd_item_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIT_SMEM);
      attn_fwd_item_kernel<true><<<grid_i, FIT_THREADS, FIT_SMEM, st>>>(tm128, tmPad, tmOut, a, n_items, stagger, g_item_trace);
    } else {
      attn_fwd_item_kernel<false><<<grid_i, FIT_THREADS, FIT_SMEM, st>>>(tm128, tmPad, tmOut, a, n_items, stagger, nullptr);
    }
    return (int)cudaGetLastError();
  }
  if (!big && (fwd_variant == 0 || (fwd_variant == 2 && S <= 128))) {
    AttnTcArgs a{};
    a.S = S; a.H = H; a.S_pad = S_pad; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
    a.lse = lse; a.out = (__nv_bfloat16*)out; a.kmask = kmask;
    const int n_work = ((S + 127) / 128) * H * B;
    const int grid_p = n_work < num_sms() ? n_work : num_sms();
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (causal) {
      cudaFuncSetAttribute(attn_fwd_pp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FPP_SMEM);
      attn_fwd_pp_kernel<true><<<grid_p, FPP_THREADS, FPP_SMEM, st>>>(tm128, tmPad, a, n_work);
    } else {
      cudaFuncSetAttribute(attn_fwd_pp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FPP_SMEM);
      attn_fwd_pp_kernel<false><<<grid_p, FPP_THREADS, FPP_SMEM, st>>>(tm128, tmPad, a, n_work);
    }
    return (int)cudaGetLastError();
  }
  AttnTcArgs a{};
  a.S = S; a.H = H; a.S_pad = S_pad; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  a.lse = lse; a.out = (__nv_bfloat16*)out; a.kmask = kmask;
  dim3 grid((S + 127) / 128, H, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (causal) {
    cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BIG);
    attn_fwd_tc_kernel<true><<<grid, ATT_THREADS, smem_bytes, st>>>(tm128, tmPad, tmRem, a);
  } else {
    cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BIG);
    attn_fwd_tc_kernel<false><<<grid, ATT_THREADS, smem_bytes, st>>>(tm128, tmPad, tmRem, a);
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
  if (head_dim != 64) return MMB_ERR_UNSUPPORTED;
  return attention_fwd_tc_impl(qkv, out, lse, kmask, B, S, H, causal, scale, stream);
}

static int attn_bwd_variant() {   // 0 = two-pass ping-pong, 1 = two-pass column-split, 2 = fused single pass
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("MMB_ATTN_BWD");
    variant = (e && e[0] == 'p') ? 0 : (e && e[0] == 'f') ? 2 : (e && e[0] == 'c') ? 1 : MMB_ATTN_BWD_DEFAULT;
  }
  return variant;
}
// kernels one mmb_attention_bwd call launches at sequence length S (callers that count launches: bench.py gpu_launches)
extern "C" int mmb_attention_bwd_launches(int S) { return (attn_bwd_variant() == 2 && S <= 256) ? 1 : 2; }

static int attention_bwd_tc_impl(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                 const uint8_t* kmask, int B, int S, int H, int causal, float scale, void* stream) {
  if (B <= 0 || S <= 0 || S > SMAX) return MMB_ERR_UNSUPPORTED;
  // key-padding mask: implemented in the fused single-pass kernel only (S <= 256, the default variant)
  if (kmask != nullptr && !(attn_bwd_variant() == 2 && S <= 256)) return MMB_ERR_UNSUPPORTED;
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
  a.dqkv = (__nv_bfloat16*)dqkv; a.dsum = dsum; a.kmask = kmask;
  const int n_work = ((S + 127) / 128) * H * B;
  const int grid_p = n_work < num_sms() ? n_work : num_sms();
  // Backward kernels: the two-pass column-split one (S <= 256), the fused single-pass one (S <= 256) and the two-pass
  // ping-pong one (the only one for 256 < S <= 384).  MMB_ATTN_BWD=fused / =colsplit / =pp force one (A/B).
  const int variant = attn_bwd_variant();
  if (variant == 2 && S <= 256) {
    const int n_items = B * H;
    const int grid_f = n_items < num_sms() ? n_items : num_sms();
    static int l2pf = -1;
    if (l2pf < 0) {
      const char* e = getenv("MMB_ATTN_L2PF");
      l2pf = (e && e[0] == '0') ? 0 : 1;
    }
    a.l2_prefetch = l2pf;
    a.trace = g_item_trace;
    static int nstat = -1;   // MMB_ATTN_FUSED_STATS=1|2: statistics warps of the fused kernel (A/B)
    if (nstat < 0) {
      const char* e = getenv("MMB_ATTN_FUSED_STATS");
      nstat = (e && e[0] == '2') ? 2 : 1;
    }
    CUtensorMap tmDQKV;   // [B][S][3d] view of dqkv: the per-tile store boxes are clipped at S
    rc = make_tmap_3d_bf16(&tmDQKV, dqkv, 3ull * d, (uint64_t)S, (uint64_t)B, 3ull * d * 2, (uint64_t)S * 3 * d * 2, 64, 128);
    if (rc) return rc;
#define LAUNCH_BWDF(C, NS)                                                                                       \
    cudaFuncSetAttribute(attn_bwd_fused_kernel<C, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWDF_SMEM); \
    attn_bwd_fused_kernel<C, NS><<<grid_f, (8 + 3 + NS) * 32, BWDF_SMEM, st>>>(q128, q64, o64, tmDQKV, a, n_items);
    if (causal) { if (nstat == 2) { LAUNCH_BWDF(true, 2) } else { LAUNCH_BWDF(true, 1) } }
    else        { if (nstat == 2) { LAUNCH_BWDF(false, 2) } else { LAUNCH_BWDF(false, 1) } }
#undef LAUNCH_BWDF
    return (int)cudaGetLastError();
  }
  if (variant == 1 && S <= 256) {
#define LAUNCH_BWDP(C, K, SM)                                                                                     \
    cudaFuncSetAttribute(attn_bwd_persist_kernel<C, K, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);     \
    attn_bwd_persist_kernel<C, K, 2><<<grid_p, (8 + 3 + P_STATS) * 32, SM, st>>>(q128, q64, o128, o64, a, n_work);
    if (causal) {
      LAUNCH_BWDP(true, false, BWDP_DQ_SMEM)     // dQ first: it also publishes D for the dK/dV kernel
      LAUNCH_BWDP(true, true, BWDP_DKDV_SMEM)
    } else {
      LAUNCH_BWDP(false, false, BWDP_DQ_SMEM)
      LAUNCH_BWDP(false, true, BWDP_DKDV_SMEM)
    }
#undef LAUNCH_BWDP
    return (int)cudaGetLastError();
  }
#define LAUNCH_BWDPP(C, K, SM)                                                                                    \
  cudaFuncSetAttribute(attn_bwd_pp_kernel<C, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);               \
  attn_bwd_pp_kernel<C, K><<<grid_p, (8 + 3 + PP_STATS) * 32, SM, st>>>(q128, q64, o128, o64, a, n_work);
  if (causal) {
    LAUNCH_BWDPP(true, false, BWDPP_DQ_SMEM)     // dQ first: it also publishes D for the dK/dV kernel
    LAUNCH_BWDPP(true, true, BWDPP_DKDV_SMEM)
  } else {
    LAUNCH_BWDPP(false, false, BWDPP_DQ_SMEM)
    LAUNCH_BWDPP(false, true, BWDPP_DKDV_SMEM)
  }
#undef LAUNCH_BWDPP
  return (int)cudaGetLastError();
}

extern "C" int mmb_attention_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                    int B, int S, int H, int causal, float scale, void* stream) {
  return attention_bwd_tc_impl(qkv, out, dout, lse, dqkv, nullptr, B, S, H, causal, scale, stream);
}
// Backward of mmb_attention_fwd_kmask: masked keys get P = dS = 0, i.e. zero dK / dV rows and no share in dQ
// (modules/layers/attention.py:220-239 under autograd, additive -inf mask of utils/attention.py:13-53).
extern "C" int mmb_attention_bwd_kmask(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                       const unsigned char* kmask, int B, int S, int H, int head_dim, int causal,
                                       float scale, void* stream) {
  if (head_dim != 64) return MMB_ERR_UNSUPPORTED;
  return attention_bwd_tc_impl(qkv, out, dout, lse, dqkv, kmask, B, S, H, causal, scale, stream);
}

// Public entry points (include/mmb200.h).  Every sequence length up to SMAX = 384 runs on the tcgen05 kernels above;
// there is no other attention path for head_dim 64 (the round-1 mma.sync kernels for 256 < S <= 320 are gone).
extern "C" int mmb_attention_fwd(const void* qkv, void* out, float* lse, int B, int S, int H, int head_dim, int causal,
                                 float scale, void* stream) {
  if (head_dim != 64) return MMB_ERR_UNSUPPORTED;
  return mmb_attention_fwd_tc(qkv, out, lse, B, S, H, causal, scale, stream);
}
extern "C" int mmb_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                 int B, int S, int H, int head_dim, int causal, float scale, void* stream) {
  if (head_dim != 64) return MMB_ERR_UNSUPPORTED;
  return mmb_attention_bwd_tc(qkv, out, dout, lse, dqkv, B, S, H, causal, scale, stream);
}
