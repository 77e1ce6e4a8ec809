/* libmmb200 — C ABI of the B200-native dual-encoder + contrastive-loss hot path.
 *
 * The reference (facebookresearch/multimodal) has no FFI: its hot path is Python calling
 * torch.nn.functional.  These entry points are what a maintainer would bind (ctypes; see
 * INTEGRATION.md) to replace each library call on that path.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless noted;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream;
 *   - row-major tensors with explicit leading dimensions (in ELEMENTS);
 *   - return value: 0 on success, a positive cudaError_t, or a negative MMB_ERR_* code.
 *     There is no CPU fallback: unsupported shapes return MMB_ERR_UNSUPPORTED.
 */
#ifndef MMB200_H_
#define MMB200_H_
#ifdef __cplusplus
extern "C" {
#endif

#define MMB_OK 0
#define MMB_ERR_ARG (-22)
#define MMB_ERR_UNSUPPORTED (-95)
#define MMB_ERR_DRIVER (-5)

/* GEMM epilogues (mmb_gemm_bf16) */
#define MMB_EPI_BF16 0      /* D0 = bf16(alpha*acc + bias)                                   */
#define MMB_EPI_BF16_ACT 1  /* D0 = bf16(pre = alpha*acc + bias), D1 = bf16(act(D0))         */
#define MMB_EPI_BF16_DACT 2 /* D0 = bf16(alpha*acc * act'(aux))                              */
#define MMB_EPI_F32 3       /* D0 = fp32(alpha*acc + bias); reduce-add when split-K/accumulate */
#define MMB_ACT_QUICK_GELU 0 /* torchmultimodal/modules/layers/activation.py:12-25 ("SiLU")  */
#define MMB_ACT_GELU_ERF 1   /* nn.GELU(), torchmultimodal/modules/layers/mlp.py             */

int mmb_version(void);

/* D[M,N] = alpha * A (x) B (+ bias[N]), bf16 operands, fp32 accumulation on tcgen05 tensor cores.
 *   a_mn_major = 0: A is [M,K] row-major (lda >= K);  1: A is stored [K,M] row-major (lda >= M)
 *   b_mn_major = 0: B is [N,K] row-major (ldb >= K);  1: B is stored [K,N] row-major (ldb >= N)
 * Replaces: F.linear in torch/nn/functional.py:6478 (in-proj), :6690 (out-proj),
 *           torch/nn/modules/transformer.py:980-982 (linear1/linear2), their autograd dgrad/wgrad,
 *           torch.matmul in modules/losses/contrastive_loss_with_temperature.py:90-95,
 *           `x @ self.projection` in models/clip/image_encoder.py:112, Linear in text_encoder.py:130. */
int mmb_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                  void* D0, long long ldd0, void* D1, long long ldd1, int M, int N, int K, int epilogue, int act,
                  float alpha, const float* bias, const void* aux, long long ld_aux, int splits, int accumulate,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMB200_H_ */
