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
// One CTA per image; the image stays in SHARED memory (S*S*3 bytes: 147 KB at 224 px) between its two operations, so an
// augmented image costs one HBM read and one HBM write (algorithmic bytes 2 * S*S*3 per image) whatever they are.  It
// arrives and leaves by 1-D bulk copies (the TMA engine, no registers in between).  Point operations (tables, colour) and
// the 3 x 3 sharpness (in row bands through the rest of the shared memory) run in place on the buffer; the affine gather
// reads the buffer and writes the output image directly.  Statistics an operation needs (grey mean for contrast,
// per-channel histograms for autocontrast / equalize) are taken by the same CTA from the buffer.
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
constexpr int AUG_THREADS = 1024;   // upper bound; the host picks 256 / 512 / 1024 by image size
constexpr int AUG_NPARAM = 8;      // floats per (image, stage)
constexpr int AUG_NHIST = 4;       // privatised histogram copies (one per lane & 3)

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

// 1-D bulk copies (the TMA engine): global -> shared counted on an mbarrier, shared -> global in a bulk group.
// 16-byte aligned addresses and sizes.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
constexpr uint32_t AUG_BULK_CHUNK = 32768;

// Statistics (from the image in shared memory) and the 3 x 256 table of a table operation.  Ends with a CTA barrier.
__device__ void aug_build_lut(const uint8_t* img, int S, int op, const float* __restrict__ p, AugScratch& sc) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int npix = S * S;
  if (needs_hist(op)) {
    for (int i = tid; i < AUG_NHIST * 3 * 256; i += nthr) (&sc.hist[0][0][0])[i] = 0u;
    __syncthreads();
    unsigned(*h)[256] = sc.hist[tid & (AUG_NHIST - 1)];     // neighbouring pixels (similar values) hit different copies
    for (int i = tid; i < npix; i += nthr) {
      atomicAdd(&h[0][img[3 * i]], 1u);
      atomicAdd(&h[1][img[3 * i + 1]], 1u);
      atomicAdd(&h[2][img[3 * i + 2]], 1u);
    }
    __syncthreads();
    for (int i = tid; i < 3 * 256; i += nthr) {
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
    for (int i = tid; i < npix; i += nthr) s += (unsigned)gray_u8(img[3 * i], img[3 * i + 1], img[3 * i + 2]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) atomicAdd(&sc.graysum, s);
    __syncthreads();
  }
  for (int i = tid; i < 3 * 256; i += nthr) {
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
}

// A point operation on the image where it lies in shared memory (each output pixel depends on its own input pixel only).
__device__ void aug_point_inplace(uint8_t* img, int S, int op, const float* __restrict__ p, AugScratch& sc) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int npix = S * S, nbytes = 3 * npix;
  if (op == AUG_COLOR) {
    const float r = p[0], r1 = p[1];
    for (int i = tid; i < npix; i += nthr) {
      const int a = img[3 * i], b = img[3 * i + 1], c = img[3 * i + 2];
      const float g = (float)gray_u8(a, b, c);
      img[3 * i] = (uint8_t)blend_u8((float)a, g, r, r1);
      img[3 * i + 1] = (uint8_t)blend_u8((float)b, g, r, r1);
      img[3 * i + 2] = (uint8_t)blend_u8((float)c, g, r, r1);
    }
    return;
  }
  aug_build_lut(img, S, op, p, sc);
  const uint8_t* lut = &sc.lut[0][0];
  uint32_t* w4 = reinterpret_cast<uint32_t*>(img);          // the buffer is 16-byte aligned; channel of byte j is j mod 3
  for (int w = tid; w < nbytes / 4; w += nthr) {
    const uint32_t x = w4[w];
    int c = (4 * w) % 3;
    uint32_t y = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      y |= (uint32_t)lut[c * 256 + ((x >> (8 * k)) & 255u)] << (8 * k);
      c = c == 2 ? 0 : c + 1;
    }
    w4[w] = y;
  }
  for (int j = (nbytes & ~3) + tid; j < nbytes; j += nthr) img[j] = lut[(j % 3) * 256 + img[j]];
}

// Affine gather: the image in shared memory -> the output image in global memory.  Four independent pixels per thread and
// trip.  Source position = M (x - c, y - c) + o with c = (S - 1) / 2 folded into o by the host; nearest (ties to even), fill 0.
__device__ void aug_affine(const uint8_t* img, uint8_t* __restrict__ dst, int S, const float* __restrict__ p) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int npix = S * S;
  constexpr int U = 4;
  const float m00 = p[0], m01 = p[1], ox = p[2], m10 = p[3], m11 = p[4], oy = p[5];
  const float c = 0.5f * (float)(S - 1);
  for (int i0 = tid; i0 < npix; i0 += U * nthr) {
    uint8_t r[U], g[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nthr;
      r[u] = g[u] = b[u] = 0;
      if (i < npix) {
        const int y = i / S, x = i - y * S;
        const float dx = (float)x - c, dy = (float)y - c;
        const float sx = __fadd_rn(__fadd_rn(__fmul_rn(m00, dx), __fmul_rn(m01, dy)), ox);
        const float sy = __fadd_rn(__fadd_rn(__fmul_rn(m10, dx), __fmul_rn(m11, dy)), oy);
        const int ix = __float2int_rn(sx), iy = __float2int_rn(sy);
        if (sx == sx && sy == sy && ix >= 0 && ix < S && iy >= 0 && iy < S) {
          const uint8_t* q = img + 3 * (iy * S + ix);
          r[u] = q[0]; g[u] = q[1]; b[u] = q[2];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nthr;
      if (i < npix) { dst[3 * i] = r[u]; dst[3 * i + 1] = g[u]; dst[3 * i + 2] = b[u]; }
    }
  }
}

__device__ __forceinline__ int win_byte(uint32_t a, uint32_t b, uint32_t c, int o) {   // byte o of the 12-byte window a|b|c
  return o < 4 ? (a >> (8 * o)) & 255u : o < 8 ? (b >> (8 * (o - 4))) & 255u : (c >> (8 * (o - 8))) & 255u;
}

// Sharpness where the image lies: blend with the 3 x 3 smoothed image ([1 1 1; 1 5 1; 1 1 1] / 13, rounded); on the
// one-pixel border the smoothed image IS the image (and the blend r v + (1 - r) v truncates below v for some ratios).
// Images with a side <= 2 are returned unchanged.  Bands of `rows - 2` image rows at a time: the band and the row above and
// below it are copied to `scratch` (rows x 3S bytes, 16 bytes of slack on either side), then the band is rewritten from there.
__device__ void aug_sharpness_inplace(uint8_t* img, uint8_t* scratch, int rows, int S, const float* __restrict__ p) {
  if (S <= 2) return;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const float r = p[0], r1 = p[1];
  const int row = 3 * S, band = rows - 2;
  for (int y0 = 0; y0 < S; y0 += band) {
    // scratch row i holds the ORIGINAL image row y0 - 1 + i: row 0 was kept from the previous band (below), rows 1.. come
    // from the image, whose rows from y0 on are still untouched
    const int yb = min(y0 + band, S - 1);
    const int nb = min(band, S - y0);
    {
      const uint8_t* from = img + y0 * row;
      uint8_t* to = scratch + row;
      const int n = (yb - y0 + 1) * row;
      if ((row & 3) == 0) {
        for (int w = tid; w < n / 4; w += nthr) reinterpret_cast<uint32_t*>(to)[w] = reinterpret_cast<const uint32_t*>(from)[w];
      } else {
        for (int j = tid; j < n; j += nthr) to[j] = from[j];
      }
    }
    __syncthreads();
    if ((row & 3) == 0) {
      const int wrow = row / 4;
      for (int w = tid; w < nb * wrow; w += nthr) {
        const int yl = w / wrow, xw = w - yl * wrow;
        const int y = y0 + yl;
        const uint32_t* mid = reinterpret_cast<const uint32_t*>(scratch + (yl + 1) * row) + xw;
        const uint32_t* up = mid - wrow;
        const uint32_t* dn = mid + wrow;
        // bytes [4 xw - 4, 4 xw + 8) of the three rows; the words beside the row's ends are only ever used for border pixels
        const uint32_t u0 = up[-1], u1 = up[0], u2 = up[1], m0 = mid[-1], m1 = mid[0], m2 = mid[1], d0 = dn[-1], d1 = dn[0], d2 = dn[1];
        const bool yborder = y == 0 || y == S - 1;
        uint32_t out = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int x = (4 * xw + t) / 3;
          const int v = (m1 >> (8 * t)) & 255u;
          float deg = (float)v;
          if (!yborder && x > 0 && x < S - 1) {
            const int sum = win_byte(u0, u1, u2, t + 1) + win_byte(u0, u1, u2, t + 4) + win_byte(u0, u1, u2, t + 7) +
                            win_byte(m0, m1, m2, t + 1) + 5 * v + win_byte(m0, m1, m2, t + 7) +
                            win_byte(d0, d1, d2, t + 1) + win_byte(d0, d1, d2, t + 4) + win_byte(d0, d1, d2, t + 7);
            deg = (float)((2 * sum + 13) / 26);      // = rint(sum / 13) for every sum in [0, 3315]: no tie, since 13 is odd
          }
          out |= (uint32_t)blend_u8((float)v, deg, r, r1) << (8 * t);
        }
        reinterpret_cast<uint32_t*>(img + y * row)[xw] = out;
      }
    } else {
      for (int j = tid; j < nb * row; j += nthr) {
        const int yl = j / row, xb = j - yl * row;
        const int y = y0 + yl, x = xb / 3;
        const uint8_t* q = scratch + (yl + 1) * row + xb;
        const int v = q[0];
        float deg = (float)v;
        if (x > 0 && x < S - 1 && y > 0 && y < S - 1) {
          const int sum = q[-row - 3] + q[-row] + q[-row + 3] + q[-3] + 5 * v + q[3] + q[row - 3] + q[row] + q[row + 3];
          deg = (float)((2 * sum + 13) / 26);      // = rint(sum / 13) for every sum in [0, 3315]: no tie, since 13 is odd
        }
        img[y * row + xb] = (uint8_t)blend_u8((float)v, deg, r, r1);
      }
    }
    __syncthreads();
    if (y0 + band < S) {       // the original of the band's last row becomes "the row above" of the next band
      for (int j = tid; j < row; j += nthr) scratch[j] = scratch[band * row + j];
      __syncthreads();
    }
  }
}

// Which image a CTA takes.  CTAs are dispatched in index order and a 224 px batch of 256 images is 1.73 waves of one image per
// SM, so the launch ends when the last-started EXPENSIVE image ends: hand the images out most expensive first (longest
// processing time first) and the cheap ones fill the tail.  Every CTA derives the same permutation from the operation table:
// images are classed by a rough cost (units from profiles/r02_augment_per_op.json), classes in descending cost, any fixed
// order inside a class; CTA r takes the image of rank r.  Batches beyond AUG_REORDER_MAX keep the identity order.
// Measured: ImageNet policy, 256 x 224 px: 87.5 -> 68.1 us per batch.
constexpr int AUG_REORDER_MAX = 4096;
constexpr int AUG_NCLASS = 17;
__device__ __forceinline__ int aug_cost(int op) {
  switch (op) {
    case AUG_AFFINE: case AUG_CONTRAST: return 3;
    case AUG_BRIGHTNESS: case AUG_COLOR: case AUG_POSTERIZE: case AUG_SOLARIZE: case AUG_INVERT: return 2;
    case AUG_AUTOCONTRAST: case AUG_EQUALIZE: return 4;
    case AUG_SHARPNESS: return 8;
    default: return 0;
  }
}
__device__ __forceinline__ int aug_class(const int* __restrict__ ops, int b) {     // 0 = most expensive
  return (AUG_NCLASS - 1) - (aug_cost(ops[2 * b]) + aug_cost(ops[2 * b + 1]));
}
// `work`: >= 64 + 32 ints of shared memory.  Ends with a CTA barrier; returns the same value in every thread.
__device__ int aug_pick_image(const int* __restrict__ ops, int B, int* work) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  int* cnt = work;            // [AUG_NCLASS]
  int* wsum = work + 32;      // [32]
  int* picked = work + 64;
  if (tid < AUG_NCLASS) cnt[tid] = 0;
  __syncthreads();
  for (int b = tid; b < B; b += nthr) atomicAdd(&cnt[aug_class(ops, b)], 1);
  __syncthreads();
  int cls = 0, k = blockIdx.x;                       // rank blockIdx.x = the k-th image of class cls
  while (cls < AUG_NCLASS - 1 && k >= cnt[cls]) { k -= cnt[cls]; ++cls; }
  int mine = 0;
  for (int b = tid; b < B; b += nthr) mine += aug_class(ops, b) == cls;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int v = lane < (nthr + 31) / 32 ? wsum[lane] : 0;
    int sc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, sc, o);
      if (lane >= o) sc += n;
    }
    wsum[lane] = sc - v;                             // exclusive over warps
  }
  __syncthreads();
  const int excl = incl - mine + wsum[warp];
  if (k >= excl && k < excl + mine) {
    int left = k - excl;
    for (int b = tid; b < B; b += nthr)
      if (aug_class(ops, b) == cls && left-- == 0) { *picked = b; break; }
  }
  __syncthreads();
  const int image = *picked;
  __syncthreads();                                   // `work` is the histogram area: nobody reuses it before everyone has read
  return image;
}

// One CTA per image.  The image arrives in shared memory by 1-D bulk copy, its (up to) two operations are applied where it
// lies, and it leaves by 1-D bulk copy.  The affine gather cannot run in place: it writes the output image in global memory
// straight from the buffer, and if another operation follows, the CTA fetches its own output back (an L2 hit).
__global__ void __launch_bounds__(AUG_THREADS) augment_u8_kernel(const uint8_t* __restrict__ src, uint8_t* dst,
                                                                 const int* __restrict__ ops, const float* __restrict__ params,
                                                                 int S, int scratch_rows) {
  extern __shared__ __align__(128) uint8_t aug_img[];
  __shared__ AugScratch sc;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t nbytes = 3u * (uint32_t)S * (uint32_t)S;
  uint8_t* scratch = aug_img + ((nbytes + 15u) & ~15u) + 16;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  pdl_enter();
  // (only in the one-image-per-SM regime, which the host marks by the full block size: small images share an SM and finish in
  // any order; measured at 1024 x 32 px the prologue costs 23 -> 31 us)
  const int image = ((int)gridDim.x <= AUG_REORDER_MAX && blockDim.x == AUG_THREADS) ? aug_pick_image(ops, (int)gridDim.x, reinterpret_cast<int*>(&sc.hist[0][0][0]))
                                                      : (int)blockIdx.x;
  const size_t img = (size_t)image * nbytes;
  int op[2] = {ops[2 * image], ops[2 * image + 1]};
  const float* pp[2] = {params + (size_t)image * 2 * AUG_NPARAM, params + (size_t)image * 2 * AUG_NPARAM + AUG_NPARAM};
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (op[k] < 0 || op[k] >= AUG_NOPS) op[k] = AUG_IDENTITY;     // the entry point's contract; the host sampler never emits these
  if (op[0] == AUG_IDENTITY) { op[0] = op[1]; pp[0] = pp[1]; op[1] = AUG_IDENTITY; }
  const bool bulk = (nbytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(src + img) | reinterpret_cast<uintptr_t>(dst + img)) & 15) == 0;
  uint32_t phase = 0;

  auto fetch = [&](const uint8_t* from) {       // global image -> buffer; ends with the data visible to every thread
    if (bulk) {
      if (tid == 0) {
        mbar_arrive_expect_tx(&bar, nbytes);
        for (uint32_t o = 0; o < nbytes; o += AUG_BULK_CHUNK) bulk_load_1d(aug_img + o, from + o, min(AUG_BULK_CHUNK, nbytes - o), &bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1u;
    } else {
      for (uint32_t j = tid; j < nbytes; j += nthr) aug_img[j] = from[j];
      __syncthreads();
    }
  };

  fetch(src + img);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (op[k] == AUG_IDENTITY) continue;
    if (op[k] == AUG_AFFINE) {
      aug_affine(aug_img, dst + img, S, pp[k]);
      if (k == 1 || op[1] == AUG_IDENTITY) return;      // the output image is complete
      if (bulk) asm volatile("fence.proxy.async;" ::: "memory");   // these stores, then the bulk copy engine's reads of them
      __syncthreads();
      fetch(dst + img);
    } else if (op[k] == AUG_SHARPNESS) {
      aug_sharpness_inplace(aug_img, scratch, scratch_rows, S, pp[k]);
    } else {
      aug_point_inplace(aug_img, S, op[k], pp[k], sc);
      __syncthreads();
    }
  }
  if (bulk) {
    fence_proxy_async_smem();          // the generic-proxy writes above become visible to the bulk-copy engine
    __syncthreads();
    if (tid == 0) {
      for (uint32_t o = 0; o < nbytes; o += AUG_BULK_CHUNK) bulk_store_1d(dst + img + o, aug_img + o, min(AUG_BULK_CHUNK, nbytes - o));
      bulk_commit_group();
      bulk_wait_group<0>();
    }
  } else {
    __syncthreads();
    for (uint32_t j = tid; j < nbytes; j += nthr) dst[img + j] = aug_img[j];
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
  // dynamic shared memory: the image, then (16 bytes of slack on either side) the row bands of the sharpness operation
  const size_t img_bytes = ((size_t)3 * S * S + 15) & ~size_t(15);
  const size_t max_dyn = 227 * 1024 - sizeof(AugScratch) - 1024;
  const size_t row = (size_t)3 * S;
  NVIT_REQUIRE(img_bytes + 32 + 3 * row <= max_dyn,
               "nvit_augment_u8: a %lld x %lld x 3 image (%zu bytes beside %zu bytes of tables and three rows) does not fit one SM's shared memory",
               (long long)S, (long long)S, img_bytes, sizeof(AugScratch));
  NVIT_REQUIRE(B <= 0x7fffffffll, "nvit_augment_u8: batch too large");
  size_t scratch_rows = (max_dyn - img_bytes - 32) / row;
  if (scratch_rows > (size_t)S + 2) scratch_rows = (size_t)S + 2;
  const size_t smem = img_bytes + 32 + scratch_rows * row;
  static DeviceOnce once;
  int dev;
  if (once.needed(&dev)) {
    NVIT_CUDA_CHECK(cudaFuncSetAttribute(augment_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn));
    once.mark(dev);
  }
  // one image per SM once it exceeds ~100 KB: all 32 warps of the SM work on it; small images share an SM between many CTAs
  const int threads = img_bytes >= 96 * 1024 ? 1024 : img_bytes >= 24 * 1024 ? 512 : 256;
  launch(augment_u8_kernel, (unsigned)B, threads, smem, ST(stream), static_cast<const uint8_t*>(src_u8_nhwc),
         static_cast<uint8_t*>(dst_u8_nhwc), reinterpret_cast<const int*>(ops), params, (int)S, (int)scratch_rows);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}
