// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM, MUFU.EX2 throughput, serial-vs-split reduction chains.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_probe scripts/probes/tmem_probe.cu
// Decides the floor of the attention softmax passes: S is read from TMEM once or twice per tile.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N> __device__ __forceinline__ void tld(uint32_t addr, uint32_t& acc);

template <> __device__ __forceinline__ void tld<16>(uint32_t addr, uint32_t& acc) {
  uint32_t v[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) acc ^= v[i];
}
template <> __device__ __forceinline__ void tld<32>(uint32_t addr, uint32_t& acc) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) acc ^= v[i];
}
// two x32 loads in flight before one wait
__device__ __forceinline__ void tld32x2(uint32_t a0, uint32_t a1, uint32_t& acc) {
  uint32_t v[32], w[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(a0));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]),
        "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]), "=r"(w[17]), "=r"(w[18]),
        "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]),
        "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
      : "r"(a1));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) acc ^= v[i] ^ w[i];
}

// mode 0: x16 + wait ; 1: x32 + wait ; 2: two x32 + one wait
template <int MODE>
__global__ void tmem_ld_kernel(long long* clk, uint32_t* sink, int reps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int c = 0; c < 256; c += (MODE == 0 ? 16 : (MODE == 1 ? 32 : 64))) {
      const uint32_t a = trow + ((warp >> 2) & 1) * 256 + c;
      if (MODE == 0) tld<16>(a, acc);
      if (MODE == 1) tld<32>(a, acc);
      if (MODE == 2) tld32x2(a, a + 32, acc);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// MUFU.EX2 throughput and reduction chains: each thread processes n values held in registers
template <int MODE>
__global__ void alu_kernel(long long* clk, float* sink, int reps, float seed) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = seed * (threadIdx.x + i);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    const float rf = (float)r * 1e-7f;   // rep-dependent input: identical asm statements would otherwise be merged
    if (MODE == 0) {   // ex2 + serial sum
#pragma unroll
      for (int i = 0; i < 32; ++i) { float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x[i] - s0 * 1e-30f)); s0 += e; }
    } else if (MODE == 1) {   // ex2 + 4 partial sums
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float e0, e1, e2, e3;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x[i] + rf));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x[i + 1] + rf));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(x[i + 2] + rf));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(x[i + 3] + rf));
        s0 += e0; s1 += e1; s2 += e2; s3 += e3;
      }
    } else if (MODE == 2) {   // serial max chain
#pragma unroll
      for (int i = 0; i < 32; ++i) { s0 = fmaxf(s0, x[i] + s0 * 1e-30f); }
    } else {   // 4 max chains
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        s0 = fmaxf(s0, x[i] + s0 * 1e-30f); s1 = fmaxf(s1, x[i + 1] + s1 * 1e-30f);
        s2 = fmaxf(s2, x[i + 2] + s2 * 1e-30f); s3 = fmaxf(s3, x[i + 3] + s3 * 1e-30f);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

// FFMA vs FFMA2 issue rate: 8 independent accumulator chains per thread
template <int MODE>
__global__ void fma_kernel(long long* clk, float* sink, int reps, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed * (threadIdx.x + i);
  const float b = 1.0001f, c = 0.5f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    } else {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned long long p, q, w, o;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a[i]), "f"(a[i + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(b), "f"(b));
        asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(c), "f"(c));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(o) : "l"(p), "l"(q), "l"(w));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(o));
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  long long* clk; uint32_t* sink;
  cudaMalloc(&clk, 148 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
  long long h[148];
  const int reps = 200;
  const char* names[3] = {"32x32b.x16 + wait", "32x32b.x32 + wait", "2 x (32x32b.x32) + wait"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16}) {
      for (int it = 0; it < 2; ++it) {
        if (mode == 0) tmem_ld_kernel<0><<<148, warps * 32>>>(clk, sink, reps);
        if (mode == 1) tmem_ld_kernel<1><<<148, warps * 32>>>(clk, sink, reps);
        if (mode == 2) tmem_ld_kernel<2><<<148, warps * 32>>>(clk, sink, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = (double)reps * warps * 32 * 256 * 4;   // per CTA (= per SM)
      printf("tcgen05.ld %-26s %2d warps: %8lld clk  -> %7.1f B/clk/SM  (%.0f clk per 128x256 fp32 block)\n", names[mode], warps,
             h[0], bytes / h[0], 131072.0 / (bytes / h[0]));
    }
  const char* an[4] = {"ex2 + serial sum", "ex2 + 4 partial sums", "serial fmax+fma chain", "4 fmax+fma chains"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {4, 8, 16}) {
      for (int it = 0; it < 2; ++it) {
        if (mode == 0) alu_kernel<0><<<148, warps * 32>>>(clk, (float*)sink, 1000, 0.001f);
        if (mode == 1) alu_kernel<1><<<148, warps * 32>>>(clk, (float*)sink, 1000, 0.001f);
        if (mode == 2) alu_kernel<2><<<148, warps * 32>>>(clk, (float*)sink, 1000, 0.001f);
        if (mode == 3) alu_kernel<3><<<148, warps * 32>>>(clk, (float*)sink, 1000, 0.001f);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      const double n = 1000.0 * 32 * warps * 32;
      printf("%-24s %2d warps: %8lld clk -> %6.2f elem/clk/SM\n", an[mode], warps, h[0], n / h[0]);
    }
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16}) {
      for (int it = 0; it < 2; ++it) {
        if (mode == 0) fma_kernel<0><<<148, warps * 32>>>(clk, (float*)sink, 2000, 0.001f);
        if (mode == 1) fma_kernel<1><<<148, warps * 32>>>(clk, (float*)sink, 2000, 0.001f);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      const double n = 2000.0 * 16 * warps * 32;
      printf("%-24s %2d warps: %8lld clk -> %6.1f fma/clk/SM\n", mode == 0 ? "FFMA, 16 chains" : "FFMA2, 8 packed chains", warps, h[0],
             n / h[0]);
    }
  return 0;
}
