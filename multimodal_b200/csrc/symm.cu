// Symmetric (peer-mapped) device memory for the in-kernel embedding exchange of the contrastive loss.
// Each rank cudaMalloc's one buffer, exports it with CUDA IPC, and maps every peer's buffer; kernels (the TMA
// producer of the similarity GEMMs, the CE gradient kernel) then read peer memory directly over NVLink/NVSwitch.
// Cross-GPU ordering uses per-peer epoch flags written with system-scope release stores and polled with
// system-scope acquire loads (bounded by a watchdog so that a dead peer traps instead of hanging the GPU).
// Replaces torch.distributed(.nn.functional).all_gather on the loss path (utils/distributed.py:28-58).
#include "common.cuh"
#include "mmb200_internal.h"

namespace mmb {

// flags[q] (in MY buffer) is written by peer q.  signal: store `value` into slot [rank] of every peer's flag array;
// wait: until all my slots are >= value.
__global__ void symm_signal_wait_kernel(int* const* __restrict__ peer_flags, int* __restrict__ my_flags, int rank,
                                        int world, int value) {
  const int q = threadIdx.x;
  if (q < world) {
    __threadfence_system();  // everything this stream wrote before is visible system-wide before the flag
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peer_flags[q] + rank), "r"(value) : "memory");
    const uint64_t t0 = global_timer_ns();
    int v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(my_flags + q) : "memory");
      if (v < value && global_timer_ns() - t0 > 20000000000ull) {
        printf("mmb watchdog: rank %d timed out waiting for peer %d (flag %d < %d)\n", rank, q, v, value);
        __trap();
      }
    } while (v < value);
  }
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_symm_alloc(long long bytes, void** ptr) {
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemset(*ptr, 0, (size_t)bytes);
}
extern "C" int mmb_symm_free(void* ptr) { return (int)cudaFree(ptr); }
extern "C" int mmb_symm_get_handle(void* ptr, void* handle64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) return (int)e;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
extern "C" int mmb_symm_open_handle(const void* handle64, void** peer_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  return (int)cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
extern "C" int mmb_symm_close_handle(void* peer_ptr) { return (int)cudaIpcCloseMemHandle(peer_ptr); }

// peer_flags: DEVICE array of `world` pointers (slot q -> base of peer q's flag array); my_flags: my flag array.
extern "C" int mmb_symm_signal_wait(void* const* peer_flags, void* my_flags, int rank, int world, int value,
                                    void* stream) {
  if (world < 1 || world > 32) return MMB_ERR_ARG;
  symm_signal_wait_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<int* const*>(peer_flags), reinterpret_cast<int*>(my_flags), rank, world, value);
  return (int)cudaGetLastError();
}
