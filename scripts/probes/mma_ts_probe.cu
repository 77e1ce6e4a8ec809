// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (written with tcgen05.st as bf16 pairs packed along K) against
// the same product with A in shared memory — (1) is the layout assumption right (row m in lane m, elements 2c, 2c+1 of
// the row in 32-bit column c), (2) what does one 128x64x16 MMA cost in either mode (the attention kernels are bound by
// these small MMAs: DESIGN.md section 4).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_ts_probe scripts/probes/mma_ts_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ float aval(int m, int k) { return (float)(((m * 3 + k * 5) % 7) - 3); }
__device__ __forceinline__ float bval(int k, int n) { return (float)(((k + 2 * n) % 5) - 2); }

__global__ void __launch_bounds__(128) probe(int* out, long long* clk, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;             // 128 rows x 128 B (K = 64 bf16), SW128, K-major
  uint8_t* sB = smem + 16384;     // 64 rows (n) x 128 B (k), SW128, K-major
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // A row `tid` -> smem (SS reference) and -> TMEM columns [256, 288) as packed pairs (TS operand)
  uint32_t ap[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) ap[c] = pack(aval(tid, 2 * c), aval(tid, 2 * c + 1));
#pragma unroll
  for (int ch = 0; ch < 8; ++ch)
    *reinterpret_cast<uint4*>(sA + tid * 128 + ((ch ^ (tid & 7)) << 4)) = make_uint4(ap[ch * 4], ap[ch * 4 + 1], ap[ch * 4 + 2], ap[ch * 4 + 3]);
  if (tid < 64) {
    uint32_t bp[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) bp[c] = pack(bval(2 * c, tid), bval(2 * c + 1, tid));
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
      *reinterpret_cast<uint4*>(sB + tid * 128 + ((ch ^ (tid & 7)) << 4)) = make_uint4(bp[ch * 4], bp[ch * 4 + 1], bp[ch * 4 + 2], bp[ch * 4 + 3]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(trow + 256), "r"(ap[0]), "r"(ap[1]), "r"(ap[2]), "r"(ap[3]), "r"(ap[4]), "r"(ap[5]), "r"(ap[6]), "r"(ap[7]), "r"(ap[8]),
      "r"(ap[9]), "r"(ap[10]), "r"(ap[11]), "r"(ap[12]), "r"(ap[13]), "r"(ap[14]), "r"(ap[15]), "r"(ap[16]), "r"(ap[17]), "r"(ap[18]),
      "r"(ap[19]), "r"(ap[20]), "r"(ap[21]), "r"(ap[22]), "r"(ap[23]), "r"(ap[24]), "r"(ap[25]), "r"(ap[26]), "r"(ap[27]), "r"(ap[28]),
      "r"(ap[29]), "r"(ap[30]), "r"(ap[31]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint64_t dA = desc_sw128(smem_u32(sA), 16, 1024), dB = desc_sw128(smem_u32(sB), 16, 1024);
  const uint32_t id = idesc(64);
  uint32_t phase = 0;
  for (int mode = 0; mode < 2; ++mode) {          // 0: SS -> D at columns [0, 64); 1: TS -> D at columns [64, 128)
    long long t0 = 0;
    if (tid == 0) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (mode == 0) mma_ss(tmem, dA + 2 * k, dB + 2 * k, id, k > 0);
          else           mma_ts(tmem + 64, tmem + 256 + 8 * k, dB + 2 * k, id, k > 0);
        }
      }
      commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    if (tid == 0) clk[mode] = clock64() - t0;
    asm volatile("tcgen05.fence::after_thread_sync;");
    __syncthreads();
  }
  // check both results against the exact integer product
  int bad_ss = 0, bad_ts = 0;
  for (int half = 0; half < 2; ++half) {
    uint32_t v[32], w[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(trow + half * 32));
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]),
          "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]),
          "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
        : "r"(trow + 64 + half * 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) {
      const int n = half * 32 + j;
      float ref = 0.f;
      for (int k = 0; k < 64; ++k) ref += aval(tid, k) * bval(k, n);
      if (__uint_as_float(v[j]) != ref) ++bad_ss;
      if (__uint_as_float(w[j]) != ref) ++bad_ts;
    }
  }
  atomicAdd(&out[0], bad_ss);
  atomicAdd(&out[1], bad_ts);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Issue-rate variants (A and B in smem, contents irrelevant): NACC independent accumulators written round-robin, tile N.
template <int NACC, int N, int SPIN, int OVW>
__global__ void __launch_bounds__(128) rate(long long* clk, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  const uint64_t dA = desc_sw128(smem_u32(smem), 16, 1024), dB = desc_sw128(smem_u32(smem + 16384), 16, 1024);
  const uint32_t id = idesc(N);
  if (tid == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int a = 0; a < NACC; ++a) mma_ss(tmem + a * N, dA + 2 * k, dB + 2 * k, id, OVW ? (k > 0) : 1);
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    clk[0] = clock64() - t0;
  } else if (SPIN) {
    mbar_wait(&bar, 0);      // 127 threads poll the barrier in shared memory while the MMAs run
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int NACC, int N, int SPIN = 0, int OVW = 0>
void run_rate(long long* clk) {
  const int reps = 200;
  cudaFuncSetAttribute(rate<NACC, N, SPIN, OVW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  rate<NACC, N, SPIN, OVW><<<1, 128, 65536>>>(clk, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  printf("N = %3d, %d accumulator(s) round-robin, %s, %s: %6.1f clk per MMA (%s)\n", N, NACC, SPIN ? "127 threads polling an mbarrier" : "no pollers", OVW ? "first k-step overwrites" : "always accumulate", (double)c / (reps * 4 * NACC), cudaGetErrorString(e));
}

int main() {
  int* out; long long* clk;
  cudaMalloc(&out, 8); cudaMalloc(&clk, 16);
  cudaMemset(out, 0, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  const int reps = 200;
  probe<<<1, 128, 40960>>>(out, clk, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  int h[2]; long long c[2];
  cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost); cudaMemcpy(c, clk, 16, cudaMemcpyDeviceToHost);
  printf("mismatches (of 8192): A in smem %d, A in TMEM %d\n", h[0], h[1]);
  printf("128x64x16 MMA, %d back-to-back: A in smem %.1f clk each, A in TMEM %.1f clk each\n", reps * 4, (double)c[0] / (reps * 4),
         (double)c[1] / (reps * 4));
  run_rate<1, 64>(clk); run_rate<2, 64>(clk); run_rate<4, 64>(clk); run_rate<8, 64>(clk);
  run_rate<1, 128>(clk); run_rate<2, 128>(clk); run_rate<1, 256>(clk); run_rate<2, 256>(clk);
  run_rate<1, 16>(clk); run_rate<1, 32>(clk);
  run_rate<1, 64, 1, 0>(clk); run_rate<1, 64, 0, 1>(clk); run_rate<1, 64, 1, 1>(clk); run_rate<1, 256, 1, 0>(clk);
  return 0;
}
