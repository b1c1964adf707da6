// Row kernels of the "original ViT" branch (config.use_nvit = False; BASELINE config 4, the A/B against nViT):
//
//   add_rmsnorm :  t = h (+ x) ;  y = t * rsqrt(mean(t^2) + eps) * w          RMSNorm (nvit/model.py:172-184) applied to the
//                  plain residual sum h + branch (nvit/model.py:95-96, 132-133, 145-146; cross-attention :221-223)
//   add_skipnorm:  u = h + x ;  out = N(u * skip + h0)                         second residual add (model.py:157-158) fused
//                  with Block.norm_skip (model.py:84-87, 450-452), which the reference applies in this mode too
//
// Same structure as residual.cu: one warp per row, the row in registers, shuffle reductions, each stream touched once.
#include "common.cuh"

namespace nvit {

__device__ __forceinline__ void rn_st_f4(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void rn_st_bf4(__nv_bfloat16* p, const float* v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
}
__device__ __forceinline__ void rn_ld_bf4(const __nv_bfloat16* p, float* v) {
  const uint2 u = ldg_u2_stream(p);
  v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
}
__device__ __forceinline__ void rn_ld_f4(const float* p, float* v) {
  const float4 f = ldg_f4_stream(p);
  v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}

template <int NV>
__global__ void __launch_bounds__(256) add_rmsnorm_fwd_kernel(const float* __restrict__ h, const __nv_bfloat16* __restrict__ x,
                                                              const float* __restrict__ w, float eps, float* __restrict__ y32,
                                                              __nv_bfloat16* __restrict__ y16, int M, int C) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float wv[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) wv[j][e] = (c < C) ? w[c + e] : 0.f;
  }
  const float invC = 1.f / (float)C;
  for (int row = warp; row < M; row += nwarps) {
    const size_t base = static_cast<size_t>(row) * C;
    float t[NV][4];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        rn_ld_f4(h + base + c, t[j]);
        if (x) {
          float xv[4];
          rn_ld_bf4(x + base + c, xv);
#pragma unroll
          for (int e = 0; e < 4; ++e) t[j][e] += xv[e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) t[j][e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) ss += t[j][e] * t[j][e];
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss * invC + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = t[j][e] * rstd * wv[j][e];
        if (y32) rn_st_f4(y32 + base + c, o);
        if (y16) rn_st_bf4(y16 + base + c, o);
      }
    }
  }
}

// dy -> dt = rstd * (dy*w - xn * mean(dy*w*xn)) ; dh (+)= dt ; dx = dt ; dw += sum_rows dy * xn
template <int NV, bool ACC>
__global__ void __launch_bounds__(256) add_rmsnorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ h,
                                                              const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                              float eps, float* dh, __nv_bfloat16* __restrict__ dx,
                                                              float* __restrict__ dw, int M, int C) {
  pdl_enter();
  extern __shared__ float s_dw[];  // [C]
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_dw[c] = 0.f;
  __syncthreads();
  float wv[NV][4], dwv[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) { wv[j][e] = (c < C) ? w[c + e] : 0.f; dwv[j][e] = 0.f; }
  }
  const float invC = 1.f / (float)C;
  for (int row = warp; row < M; row += nwarps) {
    const size_t base = static_cast<size_t>(row) * C;
    float t[NV][4], g[NV][4];
    float4 old[ACC ? NV : 1];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        rn_ld_f4(h + base + c, t[j]);
        rn_ld_f4(dy + base + c, g[j]);
        if (x) {
          float xv[4];
          rn_ld_bf4(x + base + c, xv);
#pragma unroll
          for (int e = 0; e < 4; ++e) t[j][e] += xv[e];
        }
        if (ACC) old[j] = *reinterpret_cast<const float4*>(dh + base + c);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) { t[j][e] = 0.f; g[j][e] = 0.f; }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) ss += t[j][e] * t[j][e];
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss * invC + eps);
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xn = t[j][e] * rstd;
        dwv[j][e] += g[j][e] * xn;
        g[j][e] *= wv[j][e];          // d xn
        dot += g[j][e] * xn;
        t[j][e] = xn;
      }
    dot = warp_sum(dot) * invC;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        float d[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] = (g[j][e] - t[j][e] * dot) * rstd;
        if (dx) rn_st_bf4(dx + base + c, d);
        if (ACC) { d[0] += old[j].x; d[1] += old[j].y; d[2] += old[j].z; d[3] += old[j].w; }
        rn_st_f4(dh + base + c, d);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    if (c < C) {
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(&s_dw[c + e], dwv[j][e]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dw + c, s_dw[c]);
}

template <int NV>
__global__ void __launch_bounds__(256) add_skipnorm_fwd_kernel(const float* __restrict__ h, const __nv_bfloat16* __restrict__ x,
                                                               const float* __restrict__ h0, const float* __restrict__ skip,
                                                               float* __restrict__ out32, __nv_bfloat16* __restrict__ out16, int M, int C) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float s = skip[0];
  for (int row = warp; row < M; row += nwarps) {
    const size_t base = static_cast<size_t>(row) * C;
    float y[NV][4];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        float hv[4], xv[4], h0v[4];
        rn_ld_f4(h + base + c, hv);
        rn_ld_bf4(x + base + c, xv);
        rn_ld_f4(h0 + base + c, h0v);
#pragma unroll
        for (int e = 0; e < 4; ++e) y[j][e] = (hv[e] + xv[e]) * s + h0v[e];
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) y[j][e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) ss += y[j][e] * y[j][e];
    }
    ss = warp_sum(ss);
    const float inv = 1.f / sqrtf(ss);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = y[j][e] * inv;
        if (out32) rn_st_f4(out32 + base + c, o);
        if (out16) rn_st_bf4(out16 + base + c, o);
      }
    }
  }
}

// g = dL/dout:  dy = (g - out (out.g)) / |y| ; dh0 = dy ; dh = dy * skip (fp32) ; dx = dy * skip (bf16) ; dskip += sum dy.(h + x)
template <int NV>
__global__ void __launch_bounds__(256) add_skipnorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ h,
                                                               const __nv_bfloat16* __restrict__ x, const float* __restrict__ h0,
                                                               const float* __restrict__ skip, float* __restrict__ dh,
                                                               __nv_bfloat16* __restrict__ dx, float* __restrict__ dh0,
                                                               float* __restrict__ dskip, int M, int C) {
  pdl_enter();
  __shared__ float s_ds;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  if (threadIdx.x == 0) s_ds = 0.f;
  __syncthreads();
  const float s = skip[0];
  float ds_acc = 0.f;
  for (int row = warp; row < M; row += nwarps) {
    const size_t base = static_cast<size_t>(row) * C;
    float u[NV][4], y[NV][4], gv[NV][4];
    float ss = 0.f, gy = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      if (c < C) {
        float xv[4], h0v[4];
        rn_ld_f4(h + base + c, u[j]);
        rn_ld_bf4(x + base + c, xv);
        rn_ld_f4(h0 + base + c, h0v);
        rn_ld_f4(g + base + c, gv[j]);
#pragma unroll
        for (int e = 0; e < 4; ++e) { u[j][e] += xv[e]; y[j][e] = u[j][e] * s + h0v[e]; }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) { u[j][e] = 0.f; y[j][e] = 0.f; gv[j][e] = 0.f; }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { ss += y[j][e] * y[j][e]; gy += gv[j][e] * y[j][e]; }
    }
    ss = warp_sum(ss);
    gy = warp_sum(gy);
    const float inv = 1.f / sqrtf(ss);
    const float gdot = gy * inv;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (j * 32 + lane) * 4;
      float dy[4], du[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dy[e] = (gv[j][e] - y[j][e] * inv * gdot) * inv;
        ds_acc += dy[e] * u[j][e];
        du[e] = dy[e] * s;
      }
      if (c < C) {
        rn_st_f4(dh0 + base + c, dy);
        rn_st_f4(dh + base + c, du);
        rn_st_bf4(dx + base + c, du);
      }
    }
  }
  ds_acc = warp_sum(ds_acc);
  if (lane == 0) atomicAdd(&s_ds, ds_acc);
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(dskip, s_ds);
}

static int rn_grid(int M) {
  const int want = (M + 7) / 8;
  const int cap = nvit_num_sms() * 4;
  return want < cap ? want : cap;
}

}  // namespace nvit

using namespace nvit;

#define NVIT_RN_DISPATCH(C, ...)                                  \
  do {                                                            \
    if (C <= 128) { constexpr int NV = 1; __VA_ARGS__ }           \
    else if (C <= 256) { constexpr int NV = 2; __VA_ARGS__ }      \
    else if (C <= 512) { constexpr int NV = 4; __VA_ARGS__ }      \
    else if (C <= 768) { constexpr int NV = 6; __VA_ARGS__ }      \
    else { constexpr int NV = 8; __VA_ARGS__ }                    \
  } while (0)

static int rn_check(const char* who, int64_t M, int64_t C) {
  NVIT_REQUIRE(M >= 0 && M < (1ll << 31), "%s: bad M", who);
  NVIT_REQUIRE(C > 0 && C <= 1024 && (C % 4) == 0, "%s: C=%lld must be a multiple of 4, at most 1024", who, (long long)C);
  return NVIT_OK;
}

extern "C" int nvit_add_rmsnorm_fwd(const float* h, const void* x_bf16, const float* w, float eps, float* y_f32, void* y_bf16,
                                    int64_t M, int64_t C, void* stream) {
  NVIT_REQUIRE(h && w && (y_f32 || y_bf16), "nvit_add_rmsnorm_fwd: null argument");
  int rc = rn_check("nvit_add_rmsnorm_fwd", M, C);
  if (rc) return rc;
  if (M == 0) return NVIT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NVIT_RN_DISPATCH(C, {
    launch(add_rmsnorm_fwd_kernel<NV>, rn_grid((int)M), 256, 0, st, h, static_cast<const __nv_bfloat16*>(x_bf16), w, eps, y_f32,
                                                               static_cast<__nv_bfloat16*>(y_bf16), (int)M, (int)C);
  });
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_add_rmsnorm_bwd(const float* dy, const float* h, const void* x_bf16, const float* w, float eps, float* dh,
                                    int dh_accumulate, void* dx_bf16, float* dw_accum, int64_t M, int64_t C, void* stream) {
  NVIT_REQUIRE(dy && h && w && dh && dw_accum, "nvit_add_rmsnorm_bwd: null argument");
  NVIT_REQUIRE((x_bf16 == nullptr) == (dx_bf16 == nullptr), "nvit_add_rmsnorm_bwd: x and dx go together");
  int rc = rn_check("nvit_add_rmsnorm_bwd", M, C);
  if (rc) return rc;
  if (M == 0) return NVIT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(C) * sizeof(float);
  auto xb = static_cast<const __nv_bfloat16*>(x_bf16);
  auto dxb = static_cast<__nv_bfloat16*>(dx_bf16);
  NVIT_RN_DISPATCH(C, {
    if (dh_accumulate) launch(add_rmsnorm_bwd_kernel<NV, true>, rn_grid((int)M), 256, smem, st, dy, h, xb, w, eps, dh, dxb, dw_accum, (int)M, (int)C);
    else               launch(add_rmsnorm_bwd_kernel<NV, false>, rn_grid((int)M), 256, smem, st, dy, h, xb, w, eps, dh, dxb, dw_accum, (int)M, (int)C);
  });
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_add_skipnorm_fwd(const float* h, const void* x_bf16, const float* h0, const float* skip, float* out_f32,
                                     void* out_bf16, int64_t M, int64_t C, void* stream) {
  NVIT_REQUIRE(h && x_bf16 && h0 && skip && (out_f32 || out_bf16), "nvit_add_skipnorm_fwd: null argument");
  int rc = rn_check("nvit_add_skipnorm_fwd", M, C);
  if (rc) return rc;
  if (M == 0) return NVIT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NVIT_RN_DISPATCH(C, {
    launch(add_skipnorm_fwd_kernel<NV>, rn_grid((int)M), 256, 0, st, h, static_cast<const __nv_bfloat16*>(x_bf16), h0, skip, out_f32,
                                                                static_cast<__nv_bfloat16*>(out_bf16), (int)M, (int)C);
  });
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_add_skipnorm_bwd(const float* g, const float* h, const void* x_bf16, const float* h0, const float* skip, float* dh,
                                     void* dx_bf16, float* dh0, float* dskip_accum, int64_t M, int64_t C, void* stream) {
  NVIT_REQUIRE(g && h && x_bf16 && h0 && skip && dh && dx_bf16 && dh0 && dskip_accum, "nvit_add_skipnorm_bwd: null argument");
  int rc = rn_check("nvit_add_skipnorm_bwd", M, C);
  if (rc) return rc;
  if (M == 0) return NVIT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NVIT_RN_DISPATCH(C, {
    launch(add_skipnorm_bwd_kernel<NV>, rn_grid((int)M), 256, 0, st, g, h, static_cast<const __nv_bfloat16*>(x_bf16), h0, skip, dh,
                                                                static_cast<__nv_bfloat16*>(dx_bf16), dh0, dskip_accum, (int)M, (int)C);
  });
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}
