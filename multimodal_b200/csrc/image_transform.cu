// GPU input pipeline, image half: the CLIP image transform on decoded uint8 RGB images —
//   eval : Resize(size, BICUBIC) -> CenterCrop(size) -> ToTensor -> Normalize
//   train: RandomResizedCrop(size, BICUBIC) (crop box sampled on the host) -> ToTensor -> Normalize
// (torchmultimodal/transforms/clip_transform.py:300-352), bit-exact with what the reference computes through Pillow's
// `ImagingResample` (src/libImaging/Resample.c: double-precision bicubic coefficients, window rounding by truncation,
// 22-bit fixed-point taps, a uint8-rounded image between the horizontal and the vertical pass) and torchvision's
// ToTensor / Normalize (fp32 x / 255, (x - mean) / std).
//
// Two kernels per batch, both HBM / L2-bound byte work:
//   coeffs : one thread per (image, axis, needed output coordinate) restates precompute_coeffs + normalize_coeffs_8bpc.
//            IEEE double arithmetic with explicit round-to-nearest intrinsics (no FMA contraction), so the 22-bit taps
//            are the integers Pillow computes on the host.
//   resample: one thread per output pixel (3 channels).  The two passes are fused: for every row of the vertical window
//            the horizontal pass is evaluated, rounded and clipped to uint8 exactly as Pillow's intermediate image is,
//            then accumulated — no intermediate image in HBM; the source window is re-read through L1 / L2.
#include "common.cuh"
#include "mmb200_internal.h"

namespace mmb {

constexpr int IT_KMAX = 64;          // taps per output coordinate: (int)ceil(2 * scale) * 2 + 1  ->  scale <= 15.5
constexpr int IT_PREC = 32 - 8 - 2;  // Pillow PRECISION_BITS

// geometry of one image (int32 x 12), built on the host:
//   0 H  1 W  2 pitch (bytes per row)  3 box_left  4 box_top  5 box_w  6 box_h   (the region that is resized, treated as a
//   standalone image)   7 rw  8 rh (size it is resized to)   9 crop_left  10 crop_top (offset of the out x out window in the
//   resized region)   11 flags: bit0 = horizontal pass needed, bit1 = vertical pass needed
constexpr int IT_GEOM = 12;

__device__ __forceinline__ double bicubic_filter(double x) {
  // Resample.c bicubic_filter with a = -0.5; evaluation order of the C expression, every operation rounded separately
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

// table layout per image: [axis 0 = x | axis 1 = y][out][2 + IT_KMAX] ints: (first tap, tap count, taps...)
__global__ void clip_coeffs_kernel(const int* __restrict__ geom, int* __restrict__ table, int n_images, int out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_images * 2 * out) return;
  const int o = t % out, axis = (t / out) & 1, n = t / (2 * out);
  const int* g = geom + n * IT_GEOM;
  const int in_size = axis == 0 ? g[5] : g[6];
  const int out_size = axis == 0 ? g[7] : g[8];
  const int xx = o + (axis == 0 ? g[9] : g[10]);          // coordinate in the resized image
  int* dst = table + ((long long)(n * 2 + axis) * out + o) * (2 + IT_KMAX);
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dadd_rn(0.0, __dmul_rn((double)xx + 0.5, scale));
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > IT_KMAX) xmax = IT_KMAX;     // unreachable: the host refuses scales that need more taps
  double w[IT_KMAX];
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    w[x] = bicubic_filter(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
    ww = __dadd_rn(ww, w[x]);
  }
  dst[0] = xmin;
  dst[1] = xmax;
  for (int x = 0; x < xmax; ++x) {
    const double k = (ww != 0.0) ? __ddiv_rn(w[x], ww) : w[x];
    const double f = __dmul_rn(k, (double)(1 << IT_PREC));
    dst[2 + x] = (k < 0.0) ? (int)__dadd_rn(-0.5, f) : (int)__dadd_rn(0.5, f);
  }
}

__device__ __forceinline__ int clip8(int v) {
  v >>= IT_PREC;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256) clip_resample_kernel(const unsigned long long* __restrict__ src_ptrs,
                                                            const int* __restrict__ geom, const int* __restrict__ table,
                                                            float* __restrict__ outp, int out, float m0, float m1,
                                                            float m2, float s0, float s1, float s2) {
  const int n = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= out) return;
  const int* g = geom + n * IT_GEOM;
  const long long pitch = g[2];
  const int flags = g[11];
  const uint8_t* base = reinterpret_cast<const uint8_t*>(src_ptrs[n]) + (long long)g[4] * pitch + (long long)g[3] * 3;
  const int* tx = table + ((long long)(n * 2 + 0) * out + x) * (2 + IT_KMAX);
  const int* ty = table + ((long long)(n * 2 + 1) * out + y) * (2 + IT_KMAX);
  const bool need_h = flags & 1, need_v = flags & 2;
  // without a pass along an axis the coordinate maps 1:1 (Pillow skips the pass; it is NOT an identity filter)
  const int x0 = need_h ? tx[0] : x + g[9], nx = need_h ? tx[1] : 1;
  const int y0 = need_v ? ty[0] : y + g[10], ny = need_v ? ty[1] : 1;
  int a0 = 1 << (IT_PREC - 1), a1 = a0, a2 = a0;
  int r0 = 0, r1 = 0, r2 = 0;
  for (int j = 0; j < ny; ++j) {
    const uint8_t* row = base + (long long)(y0 + j) * pitch + (long long)x0 * 3;
    int h0, h1, h2;
    if (need_h) {
      int b0 = 1 << (IT_PREC - 1), b1 = b0, b2 = b0;
      for (int i = 0; i < nx; ++i) {
        const int k = tx[2 + i];
        b0 += row[i * 3 + 0] * k; b1 += row[i * 3 + 1] * k; b2 += row[i * 3 + 2] * k;
      }
      h0 = clip8(b0); h1 = clip8(b1); h2 = clip8(b2);
    } else {
      h0 = row[0]; h1 = row[1]; h2 = row[2];
    }
    if (need_v) {
      const int k = ty[2 + j];
      a0 += h0 * k; a1 += h1 * k; a2 += h2 * k;
    } else {
      r0 = h0; r1 = h1; r2 = h2;
    }
  }
  if (need_v) { r0 = clip8(a0); r1 = clip8(a1); r2 = clip8(a2); }
  // ToTensor (uint8 -> fp32 / 255) and Normalize ((x - mean) / std): separately rounded fp32 operations
  const long long plane = (long long)out * out;
  float* o = outp + (long long)n * 3 * plane + (long long)y * out + x;
  o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)r0, 255.0f), m0), s0);
  o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)r1, 255.0f), m1), s1);
  o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)r2, 255.0f), m2), s2);
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_clip_image_transform_max_taps(void) { return IT_KMAX; }

// src_ptrs: device array of n device pointers (HWC uint8 RGB images); geom: device int32 [n, 12] (see above);
// table: device int32 scratch [n, 2, out, 2 + 64]; outp: fp32 [n, 3, out, out].
extern "C" int mmb_clip_image_transform(const void* src_ptrs, const int* geom, int* table, float* outp, int n_images,
                                        int out, const float* mean3_host, const float* std3_host, void* stream) {
  if (n_images <= 0 || out <= 0 || !src_ptrs || !geom || !table || !outp || !mean3_host || !std3_host) return MMB_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nt = n_images * 2 * out;
  clip_coeffs_kernel<<<(nt + 127) / 128, 128, 0, st>>>(geom, table, n_images, out);
  dim3 grid((out + 255) / 256, out, n_images);
  clip_resample_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(src_ptrs), geom, table, outp, out,
                                             mean3_host[0], mean3_host[1], mean3_host[2], std3_host[0], std3_host[1],
                                             std3_host[2]);
  return (int)cudaGetLastError();
}
