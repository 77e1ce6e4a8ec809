// Internal constants shared by the kernels of libmmb200.so (the public C ABI is include/mmb200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define MMB_OK 0
#define MMB_ERR_ARG (-22)          /* -EINVAL: bad pointer alignment / leading dimension / shape */
#define MMB_ERR_UNSUPPORTED (-95)  /* -EOPNOTSUPP: shape or mode not implemented (never falls back) */
#define MMB_ERR_DRIVER (-5)        /* -EIO: CUDA driver entry point / tensor-map encode failed */

// GEMM epilogues
#define EPI_BF16 0       /* D0 = bf16(alpha*acc + bias) */
#define EPI_BF16_ACT 1   /* D0 = bf16(pre), D1 = bf16(act(D0)), pre = alpha*acc + bias */
#define EPI_BF16_DACT 2  /* D0 = bf16(alpha*acc * act'(aux)) */
#define EPI_F32 3        /* D0 = fp32(alpha*acc + bias), optional reduce-add (split-K / accumulate) */
#define EPI_CE_STATS 4   /* no tensor output: per (row, 128-column part) online-softmax statistics of T*acc */
#define EPI_CE_GRAD 5    /* D0 = bf16(d loss / d acc) of the temperature-scaled cross-entropy, from T*acc in registers */

#define ACT_QUICK_GELU 0
#define ACT_GELU_ERF 1

namespace mmb {
int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, bool is_f32, uint64_t inner, uint64_t outer,
                 uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer);
int make_tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t dim1, uint64_t dim2, uint64_t pitch1,
                      uint64_t pitch2, uint32_t box_inner, uint32_t box1);
int num_sms();
}  // namespace mmb
