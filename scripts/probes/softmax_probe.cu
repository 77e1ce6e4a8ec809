// Micro-benchmark of the attention softmax inner loops on register-resident scores (no TMEM, no MMA): clocks per
// 32-score step per warp for the variants the forward kernels use.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/softmax_probe scripts/probes/softmax_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }

// MODE 0: packed (FFMA2, 32 MUFU, FADD2 x2 chains, F2FP, 4 STS.128)   1: scalar, 4 sum chains   2: packed, no STS
//      3: packed, no F2FP / STS   4: MUFU + 4 scalar sums only   5: pass 1, FMNMX3 4 chains   6: pass 1, FMNMX 4 chains
//      7: pass 1, FMNMX3 8 chains   8: MODE 0 with half of the exponentials as a degree-3 polynomial on the FMA pipe
template <int MODE>
__global__ void __launch_bounds__(512) k(long long* clk, float* sink, int reps, float seed) {
  extern __shared__ uint8_t smem[];
  const int r = threadIdx.x & 127;
  uint8_t* sP = smem + (threadIdx.x >> 7) * 16384;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = seed * (float)((threadIdx.x * 7 + i * 13) & 255) - 3.f;
  const float c = 0.18f, mx = 1.5f;
  const uint64_t c2 = pk2(c, c), nm2 = pk2(-mx, -mx);
  uint64_t s0 = pk2(0.f, 0.f), s1 = s0;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = -1e30f, a5 = -1e30f, a6 = -1e30f, a7 = -1e30f;
  __syncthreads();
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    const float rf = (float)rep * 1e-6f;
    if (MODE <= 3 || MODE == 8) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          float x0, x1, x2, x3;
          if (MODE == 1) {
            x0 = fmaf(v[hf * 16 + e] + rf, c, -mx); x1 = fmaf(v[hf * 16 + e + 1] + rf, c, -mx);
            x2 = fmaf(v[hf * 16 + e + 2] + rf, c, -mx); x3 = fmaf(v[hf * 16 + e + 3] + rf, c, -mx);
          } else {
            upk2(ffma2(pk2(v[hf * 16 + e] + rf, v[hf * 16 + e + 1]), c2, nm2), x0, x1);
            upk2(ffma2(pk2(v[hf * 16 + e + 2] + rf, v[hf * 16 + e + 3]), c2, nm2), x2, x3);
          }
          if (MODE == 8) {
            // x2, x3 on the FMA pipe: 2^x = 2^n * p(f), n = round(x), f = x - n in [-0.5, 0.5], degree-3 minimax
            x0 = ex2(x0); x1 = ex2(x1);
            const uint64_t mg = pk2(12582912.f, 12582912.f);
            const uint64_t xx = pk2(fmaxf(x2, -125.f), fmaxf(x3, -125.f));
            const uint64_t t = fadd2(xx, mg);
            const uint64_t n = fadd2(t, pk2(-12582912.f, -12582912.f));
            float n0, n1, f0, f1, t0f, t1f;
            upk2(n, n0, n1); upk2(t, t0f, t1f);
            f0 = fmaxf(x2, -125.f) - n0; f1 = fmaxf(x3, -125.f) - n1;
            uint64_t p = ffma2(pk2(f0, f1), pk2(0.0555041f, 0.0555041f), pk2(0.2402265f, 0.2402265f));
            p = ffma2(p, pk2(f0, f1), pk2(0.6931472f, 0.6931472f));
            p = ffma2(p, pk2(f0, f1), pk2(1.f, 1.f));
            float p0, p1;
            upk2(p, p0, p1);
            x2 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0f) << 23));
            x3 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1f) << 23));
          } else {
            x0 = ex2(x0); x1 = ex2(x1); x2 = ex2(x2); x3 = ex2(x3);
          }
          if (MODE == 1) { a0 += x0; a1 += x1; a2 += x2; a3 += x3; }
          else { s0 = fadd2(s0, pk2(x0, x1)); s1 = fadd2(s1, pk2(x2, x3)); }
          if (MODE != 3) { pk[e >> 1] = pack(x0, x1); pk[(e >> 1) + 1] = pack(x2, x3); }
        }
        if (MODE == 0 || MODE == 1 || MODE == 8) {
          const int j0 = ((rep & 1) * 32 + hf * 16);
          uint8_t* a = sP + r * 128;
          const int c8 = (j0 & 63) >> 3;
          *reinterpret_cast<uint4*>(a + ((c8 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(a + (((c8 + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else if (MODE == 2) {
          a0 += __uint_as_float(pk[0] ^ pk[1] ^ pk[2] ^ pk[3] ^ pk[4] ^ pk[5] ^ pk[6] ^ pk[7]);
        }
      }
    } else if (MODE == 4) {
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        a0 += ex2(v[e] + rf); a1 += ex2(v[e + 1] + rf); a2 += ex2(v[e + 2] + rf); a3 += ex2(v[e + 3] + rf);
      }
    } else if (MODE == 5) {
#pragma unroll
      for (int e = 0; e < 32; e += 8) {
        a4 = fmax3(a4, v[e] + rf, v[e + 1]); a5 = fmax3(a5, v[e + 2] + rf, v[e + 3]);
        a6 = fmax3(a6, v[e + 4] + rf, v[e + 5]); a7 = fmax3(a7, v[e + 6] + rf, v[e + 7]);
      }
    } else if (MODE == 6) {
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        a4 = fmaxf(a4, v[e] + rf); a5 = fmaxf(a5, v[e + 1] + rf); a6 = fmaxf(a6, v[e + 2] + rf); a7 = fmaxf(a7, v[e + 3] + rf);
      }
    } else if (MODE == 7) {
#pragma unroll
      for (int e = 0; e < 32; e += 16) {
        a0 = fmax3(a0, v[e] + rf, v[e + 1]); a1 = fmax3(a1, v[e + 2] + rf, v[e + 3]);
        a2 = fmax3(a2, v[e + 4] + rf, v[e + 5]); a3 = fmax3(a3, v[e + 6] + rf, v[e + 7]);
        a4 = fmax3(a4, v[e + 8] + rf, v[e + 9]); a5 = fmax3(a5, v[e + 10] + rf, v[e + 11]);
        a6 = fmax3(a6, v[e + 12] + rf, v[e + 13]); a7 = fmax3(a7, v[e + 14] + rf, v[e + 15]);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  float q0, q1, q2, q3;
  upk2(s0, q0, q1); upk2(s1, q2, q3);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + q0 + q1 + q2 + q3 + smem[threadIdx.x];
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// The forward kernel's two passes with the scores coming from TMEM (software-pipelined tcgen05.ld, as in
// attn_fwd_item_kernel): PASS 1 = max (FMNMX3 x 8 chains), PASS 2 = packed exp / sum / bf16 / STS.  `spin` extra warps
// poll an mbarrier that never completes (the producer / MMA warps of the real kernel do that most of the time).
template <int PASS>
__global__ void __launch_bounds__(512) kt(long long* clk, float* sink, int reps, int work_warps) {
  extern __shared__ uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t never;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&never)));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  long long t0 = 0, t1 = 0;
  float acc = 0.f;
  if (warp < work_warps) {
    const int r = threadIdx.x & 127;
    uint8_t* sP = smem + ((warp >> 2) & 1) * 65536;
    const uint32_t trow = tmem + ((warp >> 2) & 1) * 256 + ((uint32_t)((warp & 3) * 32) << 16);
    const float c = 0.18f;
    const uint64_t c2 = pk2(c, c);
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      uint32_t va[32], vb[32];
      if (PASS == 1) {
        float m0 = -1e30f, m1 = m0, m2 = m0, m3 = m0, m4 = m0, m5 = m0, m6 = m0, m7 = m0;
        auto red = [&](const uint32_t (&v)[32]) {
#pragma unroll
          for (int e = 0; e < 32; e += 16) {
            m0 = fmax3(m0, __uint_as_float(v[e]), __uint_as_float(v[e + 1])); m1 = fmax3(m1, __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
            m2 = fmax3(m2, __uint_as_float(v[e + 4]), __uint_as_float(v[e + 5])); m3 = fmax3(m3, __uint_as_float(v[e + 6]), __uint_as_float(v[e + 7]));
            m4 = fmax3(m4, __uint_as_float(v[e + 8]), __uint_as_float(v[e + 9])); m5 = fmax3(m5, __uint_as_float(v[e + 10]), __uint_as_float(v[e + 11]));
            m6 = fmax3(m6, __uint_as_float(v[e + 12]), __uint_as_float(v[e + 13])); m7 = fmax3(m7, __uint_as_float(v[e + 14]), __uint_as_float(v[e + 15]));
          }
        };
        tld32(trow, va);
        for (int cc = 0; cc < 6; cc += 2) {
          tld_wait(); tld32(trow + (cc + 1) * 32, vb); red(va);
          tld_wait(); if (cc + 2 < 6) tld32(trow + (cc + 2) * 32, va); red(vb);
        }
        acc += fmaxf(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), fmaxf(fmaxf(m4, m5), fmaxf(m6, m7)));
      } else {
        const float mx = 1.5f + acc * 1e-30f;
        const uint64_t nm2 = pk2(-mx, -mx);
        uint64_t s0 = pk2(0.f, 0.f), s1 = s0;
        auto ex = [&](int cc, const uint32_t (&v)[32]) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              float x0, x1, x2, x3;
              upk2(ffma2(pk2(__uint_as_float(v[hf * 16 + e]), __uint_as_float(v[hf * 16 + e + 1])), c2, nm2), x0, x1);
              upk2(ffma2(pk2(__uint_as_float(v[hf * 16 + e + 2]), __uint_as_float(v[hf * 16 + e + 3])), c2, nm2), x2, x3);
              x0 = ex2(x0); x1 = ex2(x1); x2 = ex2(x2); x3 = ex2(x3);
              s0 = fadd2(s0, pk2(x0, x1)); s1 = fadd2(s1, pk2(x2, x3));
              pk[e >> 1] = pack(x0, x1); pk[(e >> 1) + 1] = pack(x2, x3);
            }
            const int j0 = cc * 32 + hf * 16;
            uint8_t* a = sP + (j0 >> 6) * 16384 + r * 128;
            const int c8 = (j0 & 63) >> 3;
            *reinterpret_cast<uint4*>(a + ((c8 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(a + (((c8 + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        };
        tld32(trow, va);
        for (int cc = 0; cc < 6; cc += 2) {
          tld_wait(); tld32(trow + (cc + 1) * 32, vb); ex(cc, va);
          tld_wait(); if (cc + 2 < 6) tld32(trow + (cc + 2) * 32, va); ex(cc + 1, vb);
        }
        float q0, q1, q2, q3;
        upk2(s0, q0, q1); upk2(s1, q2, q3);
        acc += q0 + q1 + q2 + q3;
      }
    }
    t1 = clock64();
    __threadfence_block();
    if ((threadIdx.x & 31) == 0) atomicAdd((int*)&done, 1);
  } else {
    // spinner: poll the never-completing barrier until the workers are done
    while (done < work_warps) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&never)) : "memory");
      if (ok) break;
    }
  }
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int PASS>
void run_t(const char* name, long long* clk, float* sink) {
  long long h[148];
  const int reps = 500;
  cudaFuncSetAttribute(kt<PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  for (int spin : {0, 3}) {
    for (int warps : {4, 8}) {
      for (int it = 0; it < 2; ++it) {
        kt<PASS><<<148, (warps + spin) * 32, 131072>>>(clk, sink, reps, warps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
      }
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-44s %d work warps + %d spinning: %7.1f clk per 32-score step per warp (6 steps / row)\n", name, warps, spin,
             (double)h[0] / reps / 6);
    }
  }
}

template <int MODE>
void run(const char* name, long long* clk, float* sink) {
  long long h[148];
  const int reps = 2000;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int warps : {4, 8, 16}) {
    for (int it = 0; it < 2; ++it) {
      k<MODE><<<148, warps * 32, 65536>>>(clk, sink, reps, 0.02f);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-58s %2d warps: %7.1f clk per 32-score step per warp, %6.2f scores/clk/SM\n", name, warps, (double)h[0] / reps,
           (double)reps * 32 * warps * 32 / h[0]);
  }
}

int main() {
  long long* clk; float* sink;
  cudaMalloc(&clk, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  run_t<1>("pass 1 from TMEM (tcgen05.ld pipelined)", clk, sink);
  run_t<2>("pass 2 from TMEM (tcgen05.ld pipelined)", clk, sink);
  run<0>("pass 2 packed: FFMA2 + MUFU + FADD2 + F2FP + STS.128", clk, sink);
  run<1>("pass 2 scalar: FFMA + MUFU + FADD x4 + F2FP + STS.128", clk, sink);
  run<2>("pass 2 packed, no STS", clk, sink);
  run<3>("pass 2 packed, no F2FP, no STS", clk, sink);
  run<4>("MUFU.EX2 + 4 scalar sum chains only", clk, sink);
  run<8>("pass 2 packed, half the exps as a cubic on the FMA pipe", clk, sink);
  run<5>("pass 1: FMNMX3, 4 chains", clk, sink);
  run<6>("pass 1: FMNMX, 4 chains", clk, sink);
  run<7>("pass 1: FMNMX3, 8 chains", clk, sink);
  return 0;
}
