// Normalized residual update of nViT, forward and backward, as ONE HBM-streaming kernel each.
//
//   lr  = |alpha * alpha_mul|
//   o   = N( N(h) + lr * (N(x) - N(h)) )          (nvit/model.py:134-142, 159-167, 265-273)
//   out = N( o * skip + h0 )   when h0 != NULL    (Block.norm_skip, nvit/model.py:84-87, applied at 450-452)
//
// The reference runs ~11 (+3) eager kernels with fp32 [M,C] temporaries for this; here one warp owns one row, holds it
// in registers (C/128 float4 per lane), reduces with warp shuffles and touches HBM exactly once per operand:
//   fwd : read h (fp32) + x (bf16) [+ h0 (fp32)], write out (fp32 + bf16)
//   bwd : read g, h, x [, h0], write dh, dx [, dh0]; per-channel dalpha and scalar dskip reduced in-CTA then atomically.
// No epsilon anywhere, as in the reference (a zero row gives NaN there too).
#include "common.cuh"
#include <stdlib.h>

namespace nvit {

__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_bf4(__nv_bfloat16* p, const float* v) {
  uint2 o = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
  *reinterpret_cast<uint2*>(p) = o;
}
__device__ __forceinline__ void ld_bf4(const __nv_bfloat16* p, float* v) {
  const uint2 u = ldg_u2_stream(p);
  v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
}
__device__ __forceinline__ void ld_f4(const float* p, float* v) {
  const float4 f = ldg_f4_stream(p);
  v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}

template <int NV, bool SKIP>
__global__ void __launch_bounds__(256) residual_fwd_kernel(const float* __restrict__ h, const __nv_bfloat16* __restrict__ x,
                                                           const float* __restrict__ alpha, float alpha_mul,
                                                           const float* __restrict__ h0, const float* __restrict__ skip,
                                                           float* __restrict__ out32, __nv_bfloat16* __restrict__ out16,
                                                           int M, int C) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float lr[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) lr[j][e] = (c < C) ? fabsf(alpha[c + e] * alpha_mul) : 0.f;
  }
  const float s = SKIP ? skip[0] : 0.f;
  for (int row = warp; row < M; row += nwarps) {
    const size_t base = static_cast<size_t>(row) * C;
    float hv[NV][4], xv[NV][4], h0v[NV][4];
    float ssh = 0.f, ssx = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        ld_f4(h + base + c, hv[j]);
        ld_bf4(x + base + c, xv[j]);
        if (SKIP) ld_f4(h0 + base + c, h0v[j]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) { hv[j][e] = 0.f; xv[j][e] = 0.f; h0v[j][e] = 0.f; }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { ssh += hv[j][e] * hv[j][e]; ssx += xv[j][e] * xv[j][e]; }
    }
    ssh = warp_sum(ssh);
    ssx = warp_sum(ssx);
    const float invh = 1.f / sqrtf(ssh), invx = 1.f / sqrtf(ssx);
    float ssz = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = hv[j][e] * invh, b = xv[j][e] * invx;
        const float z = a + lr[j][e] * (b - a);
        hv[j][e] = z;
        ssz += z * z;
      }
    ssz = warp_sum(ssz);
    float inv = 1.f / sqrtf(ssz);
    if (SKIP) {
      float ssy = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float y = hv[j][e] * inv * s + h0v[j][e];
          hv[j][e] = y;
          ssy += y * y;
        }
      ssy = warp_sum(ssy);
      inv = 1.f / sqrtf(ssy);
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = hv[j][e] * inv;
        if (out32) st_f4(out32 + base + c, make_float4(o[0], o[1], o[2], o[3]));
        if (out16) st_bf4(out16 + base + c, o);
      }
    }
  }
}

// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (16-byte aligned addresses and size)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// STAGED: the operand rows do not travel through registers while they are in flight.  The register form keeps a whole
// row of every stream per warp (241-255 registers at C = 768, spills at C = 1024), which pins the kernel to eight warps
// per SM and ~60 KB of loads in flight - each warp alternates between waiting for its row and working on it.  Here one
// lane per warp issues 1-D bulk copies (the TMA engine) of the NEXT rows into a per-warp ring of shared-memory stages
// (n_stages >= 2, sized by the host from what 227 KB allow) and the warp picks its row up from there when the stage's
// mbarrier completes; the `+=` target and the skip input are read from the stage at the point of use.
template <int NV, bool SKIP, bool ACC, bool STAGED = false>
__global__ void __launch_bounds__(256) residual_bwd_kernel(const float* __restrict__ g, const float* __restrict__ h,
                                                           const __nv_bfloat16* __restrict__ x, const float* __restrict__ alpha,
                                                           float alpha_mul, const float* __restrict__ h0,
                                                           const float* __restrict__ skip, float* dh,
                                                           __nv_bfloat16* __restrict__ dx,
                                                           float* __restrict__ dh0, float* __restrict__ dalpha,
                                                           float* __restrict__ dskip, int M, int C, int n_stages) {
  pdl_enter();
  extern __shared__ __align__(128) float s_dlr[];  // [C] per-CTA reduction of d lr, then [C] lr (STAGED: barriers and stages behind)
  float* s_lr = s_dlr + C;
  __shared__ float s_dskip;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  // STAGED layout behind the two [C] vectors: n_stages mbarriers per warp, then (128-byte aligned) per warp and stage
  // the rows h | g | [h0] | [dh] (fp32) | x (bf16)
  const int wl = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const uint32_t row4 = static_cast<uint32_t>(C) * 4u, row2 = static_cast<uint32_t>(C) * 2u;
  const uint32_t stage_bytes = row4 * (2u + (SKIP ? 1u : 0u) + (ACC ? 1u : 0u)) + row2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lr + C) + wl * n_stages;
  uint8_t* stages = reinterpret_cast<uint8_t*>(s_dlr) + ((2u * row4 + 8u * wpc * n_stages + 127u) & ~127u) +
                    static_cast<size_t>(wl) * n_stages * stage_bytes;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_dlr[c] = 0.f;
    s_lr[c] = fabsf(alpha[c] * alpha_mul);
  }
  if (threadIdx.x == 0) s_dskip = 0.f;
  if (STAGED && lane == 0) {
    for (int i = 0; i < n_stages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int row, int st) {      // one lane: bring row `row` into stage `st`
    uint8_t* dst = stages + static_cast<size_t>(st) * stage_bytes;
    const size_t base = static_cast<size_t>(row) * C;
    mbar_arrive_expect_tx(&bars[st], stage_bytes);
    bulk_load_1d(dst, h + base, row4, &bars[st]); dst += row4;
    bulk_load_1d(dst, g + base, row4, &bars[st]); dst += row4;
    if (SKIP) { bulk_load_1d(dst, h0 + base, row4, &bars[st]); dst += row4; }
    if (ACC) { bulk_load_1d(dst, dh + base, row4, &bars[st]); dst += row4; }
    bulk_load_1d(dst, x + base, row2, &bars[st]);
  };
  if (STAGED && lane == 0) {
    for (int i = 0; i < n_stages - 1; ++i)
      if (warp + i * nwarps < M) issue(warp + i * nwarps, i);
  }
  int it = 0;

  float dlr[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) dlr[j][e] = 0.f;
  const float s = SKIP ? skip[0] : 0.f;
  float ds_acc = 0.f;

  for (int row = warp; row < M; row += nwarps) {
    const size_t base = static_cast<size_t>(row) * C;
    float av[NV][4], bv[NV][4], gv[NV][4], ov[NV][4];
    float4 old[(ACC && !STAGED) ? NV : 1];   // dh += : the previous value is fetched with the other streams, not after the math
    float h0r[(SKIP && !STAGED) ? NV : 1][4];
    float ssh = 0.f, ssx = 0.f;
    // STAGED: this row's stage; refill the stage consumed by the previous iteration with the row n_stages - 1 ahead
    const float* sh = nullptr; const float* sg = nullptr; const float* sh0 = nullptr; const float* sold = nullptr;
    const __nv_bfloat16* sx = nullptr;
    if constexpr (STAGED) {
      __syncwarp();                       // every lane is done with the previous iteration's stage
      if (lane == 0) {
        const long long ahead = static_cast<long long>(row) + static_cast<long long>(n_stages - 1) * nwarps;
        if (ahead < M) {
          fence_proxy_async_smem();       // generic-proxy reads of that stage are ordered before the bulk copies into it
          issue(static_cast<int>(ahead), (it + n_stages - 1) % n_stages);
        }
      }
      const int st = it % n_stages;
      mbar_wait(&bars[st], (it / n_stages) & 1);
      const uint8_t* sp = stages + static_cast<size_t>(st) * stage_bytes;
      sh = reinterpret_cast<const float*>(sp); sp += row4;
      sg = reinterpret_cast<const float*>(sp); sp += row4;
      if (SKIP) { sh0 = reinterpret_cast<const float*>(sp); sp += row4; }
      if (ACC) { sold = reinterpret_cast<const float*>(sp); sp += row4; }
      sx = reinterpret_cast<const __nv_bfloat16*>(sp);
      ++it;
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        if constexpr (STAGED) {
          const float4 a4 = *reinterpret_cast<const float4*>(sh + c), g4 = *reinterpret_cast<const float4*>(sg + c);
          const uint2 u = *reinterpret_cast<const uint2*>(sx + c);
          av[j][0] = a4.x; av[j][1] = a4.y; av[j][2] = a4.z; av[j][3] = a4.w;
          gv[j][0] = g4.x; gv[j][1] = g4.y; gv[j][2] = g4.z; gv[j][3] = g4.w;
          bv[j][0] = bf16lo(u.x); bv[j][1] = bf16hi(u.x); bv[j][2] = bf16lo(u.y); bv[j][3] = bf16hi(u.y);
        } else {
        ld_f4(h + base + c, av[j]);
        ld_bf4(x + base + c, bv[j]);
        ld_f4(g + base + c, gv[j]);
        if (SKIP) ld_f4(h0 + base + c, h0r[j]);
        if (ACC) old[j] = *reinterpret_cast<const float4*>(dh + base + c);
        }
      } else {
        if (SKIP && !STAGED) { h0r[j][0] = 0.f; h0r[j][1] = 0.f; h0r[j][2] = 0.f; h0r[j][3] = 0.f; }
#pragma unroll
        for (int e = 0; e < 4; ++e) { av[j][e] = 0.f; bv[j][e] = 0.f; gv[j][e] = 0.f; }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { ssh += av[j][e] * av[j][e]; ssx += bv[j][e] * bv[j][e]; }
    }
    ssh = warp_sum(ssh);
    ssx = warp_sum(ssx);
    const float invh = 1.f / sqrtf(ssh), invx = 1.f / sqrtf(ssx);
    float ssz = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        av[j][e] *= invh;
        bv[j][e] *= invx;
        const float z = av[j][e] + s_lr[min((j * 32 + lane) * 4 + e, C - 1)] * (bv[j][e] - av[j][e]);
        ov[j][e] = z;
        ssz += z * z;
      }
    ssz = warp_sum(ssz);
    const float invz = 1.f / sqrtf(ssz);
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) ov[j][e] *= invz;

    if (SKIP) {
      // y = o*s + h0 ; out = y/|y| ; g is dL/dout
      float ssy = 0.f, gy = 0.f;
      float yv[NV][4];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        float h0e[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (STAGED) {
          if (c < C) { const float4 t4 = *reinterpret_cast<const float4*>(sh0 + c); h0e[0] = t4.x; h0e[1] = t4.y; h0e[2] = t4.z; h0e[3] = t4.w; }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) h0e[e] = h0r[(SKIP && !STAGED) ? j : 0][e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float y = ov[j][e] * s + h0e[e];
          yv[j][e] = y;
          ssy += y * y;
          gy += gv[j][e] * y;
        }
      }
      ssy = warp_sum(ssy);
      gy = warp_sum(gy);
      const float invy = 1.f / sqrtf(ssy);
      const float gdot = gy * invy;  // g . out
      float dso = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        float dy[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          dy[e] = (gv[j][e] - yv[j][e] * invy * gdot) * invy;
          dso += dy[e] * ov[j][e];
          gv[j][e] = dy[e] * s;  // dL/do
        }
        if (c < C) st_f4(dh0 + base + c, make_float4(dy[0], dy[1], dy[2], dy[3]));
      }
      ds_acc += dso;  // lane-partial; reduced at the end
    }

    float odot = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) odot += gv[j][e] * ov[j][e];
    odot = warp_sum(odot);
    float adot = 0.f, bdot = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float dz = (gv[j][e] - ov[j][e] * odot) * invz;
        const float lre = s_lr[min((j * 32 + lane) * 4 + e, C - 1)];
        dlr[j][e] += dz * (bv[j][e] - av[j][e]);
        const float da = dz * (1.f - lre);
        const float db = dz * lre;
        adot += da * av[j][e];
        bdot += db * bv[j][e];
        gv[j][e] = da;
        ov[j][e] = db;
      }
    adot = warp_sum(adot);
    bdot = warp_sum(bdot);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        float o[4], d[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          d[e] = (gv[j][e] - av[j][e] * adot) * invh;
          o[e] = (ov[j][e] - bv[j][e] * bdot) * invx;
        }
        if constexpr (ACC && STAGED) {
          const float4 o4 = *reinterpret_cast<const float4*>(sold + c);
          d[0] += o4.x; d[1] += o4.y; d[2] += o4.z; d[3] += o4.w;
        } else if constexpr (ACC) {
          d[0] += old[j].x; d[1] += old[j].y; d[2] += old[j].z; d[3] += old[j].w;
        }
        st_f4(dh + base + c, make_float4(d[0], d[1], d[2], d[3]));
        st_bf4(dx + base + c, o);
      }
    }
  }

  // CTA-level reduction of the per-channel / scalar parameter gradients, then one atomic per channel per CTA.
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    if (c < C) {
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(&s_dlr[c + e], dlr[j][e]);
    }
  }
  if (SKIP) {
    ds_acc = warp_sum(ds_acc);
    if (lane == 0) atomicAdd(&s_dskip, ds_acc);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float am = alpha[c] * alpha_mul;
    const float sgn = (am > 0.f) ? 1.f : ((am < 0.f) ? -1.f : 0.f);  // d|u|/du, torch.abs backward convention
    atomicAdd(dalpha + c, s_dlr[c] * sgn * alpha_mul);
  }
  if (SKIP && threadIdx.x == 0) atomicAdd(dskip, s_dskip);
}

static int residual_grid(int M) {
  const int rows_per_cta = 8;
  const int want = (M + rows_per_cta - 1) / rows_per_cta;
  const int cap = nvit_num_sms() * 4;
  return want < cap ? want : cap;
}

constexpr size_t kStagedSmemMax = 232448 - 1024;   // 227 KB per CTA minus the kernel's static shared memory (rounded up)

template <int NV, bool SKIP, bool ACC, typename... Args>
static cudaError_t launch_bwd_staged(int grid, int threads, size_t smem, cudaStream_t st, Args... args) {
  static DeviceOnce once;    // per instantiation
  int dev;
  if (once.needed(&dev)) {
    cudaError_t e = cudaFuncSetAttribute(residual_bwd_kernel<NV, SKIP, ACC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kStagedSmemMax);
    if (e != cudaSuccess) return e;
    once.mark(dev);
  }
  return launch(residual_bwd_kernel<NV, SKIP, ACC, true>, grid, threads, smem, st, args...);
}

}  // namespace nvit

using namespace nvit;

static std::atomic<int> g_bwd_staged{2};   // nvit_residual_bwd_staged: 0 registers, 1 staged, 2 automatic (default)

#define NVIT_DISPATCH_NV(C, ...)                                  \
  do {                                                            \
    if (C <= 128) { constexpr int NV = 1; __VA_ARGS__ }           \
    else if (C <= 256) { constexpr int NV = 2; __VA_ARGS__ }      \
    else if (C <= 512) { constexpr int NV = 4; __VA_ARGS__ }      \
    else if (C <= 768) { constexpr int NV = 6; __VA_ARGS__ }      \
    else { constexpr int NV = 8; __VA_ARGS__ }                    \
  } while (0)

extern "C" int nvit_residual_fwd(const float* h, const void* x_bf16, const float* alpha, float alpha_mul, const float* h0,
                                 const float* skip, float* out_f32, void* out_bf16, int64_t M, int64_t C, void* stream) {
  NVIT_REQUIRE(h && x_bf16 && alpha, "nvit_residual_fwd: null input");
  NVIT_REQUIRE(out_f32 || out_bf16, "nvit_residual_fwd: no output requested");
  NVIT_REQUIRE((h0 == nullptr) == (skip == nullptr), "nvit_residual_fwd: h0 and skip go together");
  NVIT_REQUIRE(M >= 0 && M < (1ll << 31), "nvit_residual_fwd: bad M");
  NVIT_REQUIRE(C > 0 && C <= 1024 && (C % 4) == 0, "nvit_residual_fwd: C=%lld must be a multiple of 4, at most 1024", (long long)C);
  if (M == 0) return NVIT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = residual_grid((int)M);
  auto xb = static_cast<const __nv_bfloat16*>(x_bf16);
  auto ob = static_cast<__nv_bfloat16*>(out_bf16);
  NVIT_DISPATCH_NV(C, {
    if (h0) launch(residual_fwd_kernel<NV, true>, grid, 256, 0, st, h, xb, alpha, alpha_mul, h0, skip, out_f32, ob, (int)M, (int)C);
    else    launch(residual_fwd_kernel<NV, false>, grid, 256, 0, st, h, xb, alpha, alpha_mul, h0, skip, out_f32, ob, (int)M, (int)C);
  });
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_residual_bwd(const float* g, const float* h, const void* x_bf16, const float* alpha, float alpha_mul,
                                 const float* h0, const float* skip, float* dh, int dh_accumulate, void* dx_bf16, float* dh0,
                                 float* dalpha_accum, float* dskip_accum, int64_t M, int64_t C, void* stream) {
  NVIT_REQUIRE(g && h && x_bf16 && alpha && dh && dx_bf16 && dalpha_accum, "nvit_residual_bwd: null argument");
  NVIT_REQUIRE((h0 == nullptr) == (skip == nullptr), "nvit_residual_bwd: h0 and skip go together");
  NVIT_REQUIRE(!h0 || (dh0 && dskip_accum), "nvit_residual_bwd: skip form needs dh0 and dskip");
  NVIT_REQUIRE(M >= 0 && M < (1ll << 31), "nvit_residual_bwd: bad M");
  NVIT_REQUIRE(C > 0 && C <= 1024 && (C % 4) == 0, "nvit_residual_bwd: C=%lld must be a multiple of 4, at most 1024", (long long)C);
  if (M == 0) return NVIT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = residual_grid((int)M);
  const size_t smem = 2 * static_cast<size_t>(C) * sizeof(float);
  auto xb = static_cast<const __nv_bfloat16*>(x_bf16);
  auto dxb = static_cast<__nv_bfloat16*>(dx_bf16);
  // Staged form (rows arrive through per-warp rings of bulk copies): needs 16-byte rows in every stream and room for at
  // least two stages per warp; one persistent CTA per SM with as many warps (8, 6 or 4) and stages (<= 4) as 227 KB hold.
  const uintptr_t align_all = reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(x_bf16) |
                              reinterpret_cast<uintptr_t>(h0) | reinterpret_cast<uintptr_t>(dh);
  // MEASURED (scripts/hbm_kernels_bench.py, M = 50 176, registers -> staged with eight warps): C = 768: plain 129 -> 114 us,
  // += 140 -> 137, skip 168 -> 175, skip and += 302 -> 200; C = 1024: 177 -> 147, 245 -> 183, 238 -> 238, 404 -> 262.  The skip
  // form without += has the longest arithmetic chain per row and the largest stage: with eight warps only two stages fit and the
  // ring does not help.  MEASURED (scripts/residual_w_sweep.py, C = 768, us): skip form 6 warps x 3 stages 157, 4 x 4 183
  // (registers 168); += form 8 warps 134, 6 warps 136; plain form 113 / 130.  So the skip form alone starts at six warps and takes
  // the ring only where three stages fit (C <= 832), else it stays in registers; small widths (tiny models, microsecond
  // launches) stay on the register form too.
  const bool skip_only = h0 && !dh_accumulate;
  const size_t stage = static_cast<size_t>(C) * 4 * (2 + (h0 ? 1 : 0) + (dh_accumulate ? 1 : 0)) + static_cast<size_t>(C) * 2;
  auto ring_bytes = [&](int W, int ns) { return ((smem + 8u * W * ns + 127u) & ~size_t(127)) + static_cast<size_t>(W) * ns * stage; };
  const bool want_staged = g_bwd_staged == 1 || (g_bwd_staged == 2 && C >= 512 && (!skip_only || ring_bytes(6, 3) <= kStagedSmemMax));
  if (want_staged && (C % 8) == 0 && (align_all & 15) == 0) {
    int W0 = skip_only ? 6 : 8;
#ifdef NVIT_BENCH_HOOKS
    if (const char* e = getenv("NVIT_RES_W")) W0 = atoi(e);   // measurement only: first warp count tried
#endif
    for (int W = W0; W >= 4; W -= 2) {
      int ns = 4;
      size_t total = 0;
      for (; ns >= 2; --ns) {
        total = ring_bytes(W, ns);
        if (total <= kStagedSmemMax) break;
      }
      if (ns < 2) continue;
      const int want = (int)((M + W - 1) / W);
      const int sgrid = want < nvit_num_sms() ? want : nvit_num_sms();
      cudaError_t le = cudaSuccess;
      NVIT_DISPATCH_NV(C, {
        if (h0 && dh_accumulate)       le = launch_bwd_staged<NV, true, true>(sgrid, W * 32, total, st, g, h, xb, alpha, alpha_mul, h0, skip, dh, dxb, dh0, dalpha_accum, dskip_accum, (int)M, (int)C, ns);
        else if (h0)                   le = launch_bwd_staged<NV, true, false>(sgrid, W * 32, total, st, g, h, xb, alpha, alpha_mul, h0, skip, dh, dxb, dh0, dalpha_accum, dskip_accum, (int)M, (int)C, ns);
        else if (dh_accumulate)        le = launch_bwd_staged<NV, false, true>(sgrid, W * 32, total, st, g, h, xb, alpha, alpha_mul, h0, skip, dh, dxb, dh0, dalpha_accum, dskip_accum, (int)M, (int)C, ns);
        else                           le = launch_bwd_staged<NV, false, false>(sgrid, W * 32, total, st, g, h, xb, alpha, alpha_mul, h0, skip, dh, dxb, dh0, dalpha_accum, dskip_accum, (int)M, (int)C, ns);
      });
      NVIT_CUDA_CHECK(le);
      NVIT_CUDA_CHECK(cudaGetLastError());
      return NVIT_OK;
    }
  }
  NVIT_DISPATCH_NV(C, {
    if (h0 && dh_accumulate)       launch(residual_bwd_kernel<NV, true, true>, grid, 256, smem, st, g, h, xb, alpha, alpha_mul, h0, skip, dh, dxb, dh0, dalpha_accum, dskip_accum, (int)M, (int)C, 0);
    else if (h0)                   launch(residual_bwd_kernel<NV, true, false>, grid, 256, smem, st, g, h, xb, alpha, alpha_mul, h0, skip, dh, dxb, dh0, dalpha_accum, dskip_accum, (int)M, (int)C, 0);
    else if (dh_accumulate)        launch(residual_bwd_kernel<NV, false, true>, grid, 256, smem, st, g, h, xb, alpha, alpha_mul, h0, skip, dh, dxb, dh0, dalpha_accum, dskip_accum, (int)M, (int)C, 0);
    else                           launch(residual_bwd_kernel<NV, false, false>, grid, 256, smem, st, g, h, xb, alpha, alpha_mul, h0, skip, dh, dxb, dh0, dalpha_accum, dskip_accum, (int)M, (int)C, 0);
  });
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_residual_bwd_staged(int mode) {   // 0: rows held in registers, 1: rows staged in shared memory, 2: automatic
  NVIT_REQUIRE(mode >= 0 && mode <= 2, "nvit_residual_bwd_staged: mode must be 0 (registers), 1 (staged) or 2 (automatic)");
  g_bwd_staged = mode;
  return NVIT_OK;
}
