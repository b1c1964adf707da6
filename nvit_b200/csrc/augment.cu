// Device-side AutoAugment for raw uint8 HWC batches (the input pipeline on the device).
//
// The reference augments in its DataLoader workers with kornia.augmentation.auto.AutoAugment(dataset) behind
// Normalize(0.5, 0.5) (train.py:1081-1092, applied per sample at train.py:262-273).  Here the batch crosses PCIe as raw
// uint8 [B, S, S, 3] (nvit_im2col_u8 folds ToTensor + Normalize into the patch gather) and this kernel applies, per image,
// the TWO operations of the sub-policy the host sampler drew for it (nvit_b200/augment.py) before the patch gather:
//
//     op 0 identity      1 affine (inverse 2x3 map about the centre, nearest, fill 0: ShearX/Y, TranslateX/Y, Rotate)
//        2 brightness    3 color (saturation)    4 contrast    5 sharpness    6 posterize    7 solarize
//        8 autocontrast  9 equalize              10 invert
//
// One CTA per image.  The first operation reads the image from global memory and writes its result into SHARED memory
// (S*S*3 bytes: 147 KB at 224 px), the second reads shared memory and writes the output image, so an augmented image
// costs one HBM read and one HBM write (algorithmic bytes 2 * S*S*3 per image) whatever the two operations are; images
// whose first operation did not fire skip the shared-memory pass.  Statistics an operation needs (grey mean for
// contrast, per-channel histograms for autocontrast / equalize) are taken by the same CTA over its own source.
//
// Arithmetic is that of torchvision's uint8 tensor kernels (the oracle, oracle/augment_oracle.py, is pinned against
// torchvision 0.26 in tests/test_augment_cpu.py): every float expression is written with explicit single roundings
// (__fmul_rn / __fadd_rn, no contraction into FMA) so that the result is BIT-EXACT against the numpy restatement.
#include "common.cuh"

#define ST(s) static_cast<cudaStream_t>(s)

namespace nvit {
namespace {

enum AugOp : int {
  AUG_IDENTITY = 0, AUG_AFFINE = 1, AUG_BRIGHTNESS = 2, AUG_COLOR = 3, AUG_CONTRAST = 4, AUG_SHARPNESS = 5,
  AUG_POSTERIZE = 6, AUG_SOLARIZE = 7, AUG_AUTOCONTRAST = 8, AUG_EQUALIZE = 9, AUG_INVERT = 10, AUG_NOPS = 11
};
constexpr int AUG_THREADS = 512;
constexpr int AUG_NPARAM = 8;      // floats per (image, stage)
constexpr int AUG_NHIST = 4;       // privatised histogram copies (one per warp & 3)

struct AugScratch {
  unsigned hist[AUG_NHIST][3][256];
  unsigned long long graysum;
  uint8_t lut[3][256];
};

// torchvision _blend on uint8: (ratio * a + (1 - ratio) * b).clamp(0, 255) truncated; r1 = float(1.0 - ratio) from the host
__device__ __forceinline__ int blend_u8(float a, float b, float r, float r1) {
  float v = __fadd_rn(__fmul_rn(r, a), __fmul_rn(r1, b));
  v = fminf(fmaxf(v, 0.f), 255.f);
  return (int)v;
}
// torchvision rgb_to_grayscale on uint8: (0.2989 r + 0.587 g + 0.114 b) truncated
__device__ __forceinline__ int gray_u8(int r, int g, int b) {
  const float v = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, (float)r), __fmul_rn(0.587f, (float)g)), __fmul_rn(0.114f, (float)b));
  return (int)v;
}

__device__ __forceinline__ bool needs_hist(int op) { return op == AUG_AUTOCONTRAST || op == AUG_EQUALIZE; }
__device__ __forceinline__ bool is_lut_op(int op) {
  return op == AUG_BRIGHTNESS || op == AUG_CONTRAST || op == AUG_POSTERIZE || op == AUG_SOLARIZE || op == AUG_AUTOCONTRAST ||
         op == AUG_EQUALIZE || op == AUG_INVERT;
}

// One operation over one image: src (global or shared, generic pointer) -> dst.  All branches are uniform over the CTA.
__device__ void aug_stage(const uint8_t* src, uint8_t* dst, int S, int op, const float* __restrict__ p, AugScratch& sc) {
  const int tid = threadIdx.x;
  const int npix = S * S, nbytes = 3 * npix;

  if (is_lut_op(op)) {
    // ---- statistics
    if (needs_hist(op)) {
      for (int i = tid; i < AUG_NHIST * 3 * 256; i += AUG_THREADS) (&sc.hist[0][0][0])[i] = 0u;
      __syncthreads();
      unsigned(*h)[256] = sc.hist[(tid >> 5) & (AUG_NHIST - 1)];
      for (int i = tid; i < npix; i += AUG_THREADS) {
        atomicAdd(&h[0][src[3 * i]], 1u);
        atomicAdd(&h[1][src[3 * i + 1]], 1u);
        atomicAdd(&h[2][src[3 * i + 2]], 1u);
      }
      __syncthreads();
      for (int i = tid; i < 3 * 256; i += AUG_THREADS) {
        unsigned s = 0;
#pragma unroll
        for (int k = 0; k < AUG_NHIST; ++k) s += (&sc.hist[k][0][0])[i];
        (&sc.hist[0][0][0])[i] = s;
      }
      __syncthreads();
    } else if (op == AUG_CONTRAST) {
      if (tid == 0) sc.graysum = 0ull;
      __syncthreads();
      unsigned long long s = 0;
      for (int i = tid; i < npix; i += AUG_THREADS) s += (unsigned)gray_u8(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) atomicAdd(&sc.graysum, s);
      __syncthreads();
    }
    // ---- the 3 x 256 table
    for (int i = tid; i < 3 * 256; i += AUG_THREADS) {
      const int c = i >> 8, v = i & 255;
      int o = v;
      switch (op) {
        case AUG_BRIGHTNESS: o = blend_u8((float)v, 0.f, p[0], p[1]); break;
        case AUG_CONTRAST: {
          const float mean = __fdiv_rn((float)sc.graysum, (float)npix);
          o = blend_u8((float)v, mean, p[0], p[1]);
          break;
        }
        case AUG_POSTERIZE: o = v & (int)p[0]; break;                      // p[0] = the mask 256 - 2^(8 - bits)
        case AUG_SOLARIZE: o = ((float)v >= p[0]) ? 255 - v : v; break;
        case AUG_INVERT: o = 255 - v; break;
        case AUG_AUTOCONTRAST: {
          const unsigned* h = sc.hist[0][c];
          int mn = 0, mx = 255;
          while (mn < 255 && h[mn] == 0) ++mn;
          while (mx > 0 && h[mx] == 0) --mx;
          float scale = 1.f, lo = 0.f;
          if (mx > mn) { scale = __fmul_rn(__frcp_rn((float)(mx - mn)), 255.f); lo = (float)mn; }   // 255 / t evaluates as (1 / t) * 255
          float f = __fmul_rn(__fsub_rn((float)v, lo), scale);
          f = fminf(fmaxf(f, 0.f), 255.f);
          o = (int)f;
          break;
        }
        case AUG_EQUALIZE: {
          // PIL's equalize as torchvision restates it: step = (sum of the non-zero bins but the last) / 255;
          // lut[v] = clamp((cumsum[v - 1] + step / 2) / step, 0, 255), lut[0] = 0; step == 0 leaves the channel as it is
          const unsigned* h = sc.hist[0][c];
          int last = 255;
          while (last > 0 && h[last] == 0) --last;
          const unsigned step = ((unsigned)npix - h[last]) / 255u;
          if (step != 0) {
            unsigned cum = 0;
            for (int k = 0; k < v; ++k) cum += h[k];
            const unsigned q = (cum + step / 2) / step;
            o = v == 0 ? 0 : (int)(q > 255u ? 255u : q);
          }
          break;
        }
        default: break;
      }
      sc.lut[c][v] = (uint8_t)o;
    }
    __syncthreads();
    // ---- apply: four bytes per thread where the image is a whole number of words (channel of byte j is j mod 3)
    if ((nbytes & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0) {
      const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
      uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
      const uint8_t* lut = &sc.lut[0][0];
      for (int w = tid; w < nbytes / 4; w += AUG_THREADS) {
        const uint32_t x = s4[w];
        int c = (4 * w) % 3;
        uint32_t y = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          y |= (uint32_t)lut[c * 256 + ((x >> (8 * k)) & 255u)] << (8 * k);
          c = c == 2 ? 0 : c + 1;
        }
        d4[w] = y;
      }
    } else {
      for (int j = tid; j < nbytes; j += AUG_THREADS) dst[j] = sc.lut[j % 3][src[j]];
    }
    return;
  }

  switch (op) {
    case AUG_AFFINE: {
      // source position = M (x - c, y - c) + o with c = (S - 1) / 2 folded into o by the host; nearest (ties to even), fill 0
      const float m00 = p[0], m01 = p[1], ox = p[2], m10 = p[3], m11 = p[4], oy = p[5];
      const float c = 0.5f * (float)(S - 1);
      for (int i = tid; i < npix; i += AUG_THREADS) {
        const int y = i / S, x = i - y * S;
        const float dx = (float)x - c, dy = (float)y - c;
        const float sx = __fadd_rn(__fadd_rn(__fmul_rn(m00, dx), __fmul_rn(m01, dy)), ox);
        const float sy = __fadd_rn(__fadd_rn(__fmul_rn(m10, dx), __fmul_rn(m11, dy)), oy);
        const int ix = __float2int_rn(sx), iy = __float2int_rn(sy);
        uint8_t r = 0, g = 0, b = 0;
        if (sx == sx && sy == sy && ix >= 0 && ix < S && iy >= 0 && iy < S) {
          const uint8_t* q = src + 3 * (iy * S + ix);
          r = q[0]; g = q[1]; b = q[2];
        }
        dst[3 * i] = r; dst[3 * i + 1] = g; dst[3 * i + 2] = b;
      }
      break;
    }
    case AUG_COLOR: {
      const float r = p[0], r1 = p[1];
      for (int i = tid; i < npix; i += AUG_THREADS) {
        const int a = src[3 * i], b = src[3 * i + 1], c = src[3 * i + 2];
        const float g = (float)gray_u8(a, b, c);
        dst[3 * i] = (uint8_t)blend_u8((float)a, g, r, r1);
        dst[3 * i + 1] = (uint8_t)blend_u8((float)b, g, r, r1);
        dst[3 * i + 2] = (uint8_t)blend_u8((float)c, g, r, r1);
      }
      break;
    }
    case AUG_SHARPNESS: {
      // blend with the 3 x 3 smoothed image ([1 1 1; 1 5 1; 1 1 1] / 13, rounded); on the one-pixel border the smoothed image IS the image.
      // images with a side <= 2 are returned unchanged
      const float r = p[0], r1 = p[1];
      for (int j = tid; j < nbytes; j += AUG_THREADS) {
        const int i = j / 3, y = i / S, x = i - y * S;
        const int v = src[j];
        float deg = (float)v;      // the border blends with itself: r v + (1 - r) v, which truncates below v for some ratios
        if (x > 0 && x < S - 1 && y > 0 && y < S - 1) {
          const int row = 3 * S;
          const int sum = src[j - row - 3] + src[j - row] + src[j - row + 3] + src[j - 3] + 5 * v + src[j + 3] +
                          src[j + row - 3] + src[j + row] + src[j + row + 3];
          deg = rintf(__fdiv_rn((float)sum, 13.f));
        }
        const int o = S > 2 ? blend_u8((float)v, deg, r, r1) : v;
        dst[j] = (uint8_t)o;
      }
      break;
    }
    default: {   // identity
      for (int j = tid; j < nbytes; j += AUG_THREADS) dst[j] = src[j];
      break;
    }
  }
}

__global__ void __launch_bounds__(AUG_THREADS) augment_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                 const int* __restrict__ ops, const float* __restrict__ params,
                                                                 int S) {
  extern __shared__ __align__(16) uint8_t aug_smem[];
  AugScratch& sc = *reinterpret_cast<AugScratch*>(aug_smem);
  uint8_t* mid = aug_smem + ((sizeof(AugScratch) + 15) & ~size_t(15));
  pdl_enter();
  const size_t img = (size_t)blockIdx.x * 3u * (size_t)S * (size_t)S;
  int op1 = ops[2 * blockIdx.x], op2 = ops[2 * blockIdx.x + 1];
  const float* p1 = params + (size_t)blockIdx.x * 2 * AUG_NPARAM;
  const float* p2 = p1 + AUG_NPARAM;
  if (op1 < 0 || op1 >= AUG_NOPS) op1 = AUG_IDENTITY;     // the entry point's contract; the host sampler never emits these
  if (op2 < 0 || op2 >= AUG_NOPS) op2 = AUG_IDENTITY;
  if (op1 == AUG_IDENTITY) {
    aug_stage(src + img, dst + img, S, op2, p2, sc);
  } else if (op2 == AUG_IDENTITY) {
    aug_stage(src + img, dst + img, S, op1, p1, sc);
  } else {
    aug_stage(src + img, mid, S, op1, p1, sc);
    __syncthreads();
    aug_stage(mid, dst + img, S, op2, p2, sc);
  }
}

}  // namespace
}  // namespace nvit

extern "C" int nvit_augment_u8(const void* src_u8_nhwc, void* dst_u8_nhwc, const int32_t* ops, const float* params, int64_t B,
                               int64_t S, int64_t ch, void* stream) {
  using namespace nvit;
  NVIT_REQUIRE(src_u8_nhwc && dst_u8_nhwc && ops && params && B > 0 && S > 0, "nvit_augment_u8: bad arguments");
  NVIT_REQUIRE(src_u8_nhwc != dst_u8_nhwc, "nvit_augment_u8: the operation is not in-place");
  NVIT_REQUIRE(ch == 3, "nvit_augment_u8: the AutoAugment operations are defined on RGB images (ch = 3), got %lld", (long long)ch);
  const size_t scratch = (sizeof(AugScratch) + 15) & ~size_t(15);
  const size_t smem = scratch + (((size_t)3 * S * S + 15) & ~size_t(15));
  NVIT_REQUIRE(smem <= 227 * 1024, "nvit_augment_u8: a %lld x %lld x 3 image (%zu bytes with the tables) does not fit one SM's shared memory",
               (long long)S, (long long)S, smem);
  NVIT_REQUIRE(B <= 0x7fffffffll, "nvit_augment_u8: batch too large");
  static DeviceOnce once;
  int dev;
  if (once.needed(&dev)) {
    NVIT_CUDA_CHECK(cudaFuncSetAttribute(augment_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    once.mark(dev);
  }
  launch(augment_u8_kernel, (unsigned)B, AUG_THREADS, smem, ST(stream), static_cast<const uint8_t*>(src_u8_nhwc),
         static_cast<uint8_t*>(dst_u8_nhwc), reinterpret_cast<const int*>(ops), params, (int)S);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}
