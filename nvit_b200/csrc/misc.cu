// HBM-streaming helper kernels of the nViT training step: casts, reductions, the unfused SiLU gate, im2col,
// the pooled LayerNorm head, cross-entropy, the reconstruction loss, flat AdamW and the multi-tensor weight
// normalization.  Each replaces a run of eager PyTorch ops at the reference lines cited in include/nvit_b200.h.
#include "common.cuh"

namespace nvit {

static inline int stream_grid(long long work_items, int threads, int per_sm = 8) {
  long long want = (work_items + threads - 1) / threads;
  long long cap = 1ll * nvit_num_sms() * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

__device__ __forceinline__ float block_sum(float v, float* red /*[32]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) r = warp_max(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}

// ------------------------------------------------------------------------------------------------ cast
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                            long long n) {
  pdl_enter();
  const long long n8 = n >> 3;
  const long long stride = 1ll * gridDim.x * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (aligned) {
    for (; i < n8; i += stride) {
      const float4 a = ldg_f4_stream(src + i * 8), b = ldg_f4_stream(src + i * 8 + 4);
      uint4 o = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
      *reinterpret_cast<uint4*>(dst + i * 8) = o;
    }
    for (long long t = n8 * 8 + 1ll * blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) dst[t] = __float2bfloat16(src[t]);
  } else {
    for (; i < n; i += stride) dst[i] = __float2bfloat16(src[i]);
  }
}

// ------------------------------------------------------------------------------------------------ sum of squares
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  pdl_enter();
  __shared__ float red[32];
  const long long stride = 1ll * gridDim.x * blockDim.x;
  float acc = 0.f;
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (aligned) {
    const long long n4 = n >> 2;
    for (; i < n4; i += stride) {
      const float4 a = ldg_f4_stream(x + i * 4);
      acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    for (long long t = n4 * 4 + 1ll * blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) acc += x[t] * x[t];
  } else {
    for (; i < n; i += stride) acc += x[i] * x[i];
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, tot);
}

// Deterministic form (data-parallel replicas must compute bit-identical clip coefficients from bit-identical gradients, or
// they drift apart by an ulp per step): every CTA stores its partial sum at workspace[1 + blockIdx.x]; the CTA that arrives
// last (ticket counter in workspace[0], reset for the next launch) adds the partials IN INDEX ORDER and accumulates into out.
__global__ void __launch_bounds__(256) sumsq_det_kernel(const float* __restrict__ x, long long n, float* __restrict__ out,
                                                        float* __restrict__ ws) {
  pdl_enter();
  __shared__ float red[32];
  __shared__ bool s_last;
  const long long stride = 1ll * gridDim.x * blockDim.x;
  float acc = 0.f;
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (aligned) {
    const long long n4 = n >> 2;
    for (; i < n4; i += stride) {
      const float4 a = ldg_f4_stream(x + i * 4);
      acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    for (long long t = n4 * 4 + 1ll * blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) acc += x[t] * x[t];
  } else {
    for (; i < n; i += stride) acc += x[i] * x[i];
  }
  const float tot = block_sum(acc, red);      // fixed tree: a CTA's partial does not depend on timing
  unsigned int* ticket = reinterpret_cast<unsigned int*>(ws);
  if (threadIdx.x == 0) {
    ws[1 + blockIdx.x] = tot;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float part = 0.f;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) part += __ldcg(ws + 1 + b);   // fixed assignment b -> thread
    const float total = block_sum(part, red);
    if (threadIdx.x == 0) {
      out[0] += total;
      *ticket = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------------ column sums (bias grads)
// grid.x over 256-column chunks (thread = column), grid.y over row slabs.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int M, int N, long long ldx,
                                                          float* __restrict__ out, int rows_per_slab) {
  pdl_enter();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  const int r0 = blockIdx.y * rows_per_slab;
  const int r1 = min(M, r0 + rows_per_slab);
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += __bfloat162float(x[1ll * r * ldx + c]);
  atomicAdd(out + c, acc);
}

// dpos[t,c] += sum_b dx[b,t,c]; dbias[c] += the same, summed over t.   grid (T, ceil(C/256)).
__global__ void __launch_bounds__(256) pos_bias_grad_kernel(const float* __restrict__ dx, int B, int T, int C,
                                                            float* __restrict__ dpos, float* __restrict__ dbias) {
  pdl_enter();
  const int t = blockIdx.x;
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc += dx[(1ll * b * T + t) * C + c];
  dpos[1ll * t * C + c] += acc;
  if (dbias) atomicAdd(dbias + c, acc);
}

// ------------------------------------------------------------------------------------------------ SiLU gate (unfused form)
// sigmoid(v) = 0.5 + 0.5 tanh(v/2): one MUFU instead of EX2 + RCP
__device__ __forceinline__ float sigmoidf_(float v) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
  return fmaf(0.5f, t, 0.5f);
}

// thread = 8 consecutive columns of F; grid.y = row slabs
__global__ void __launch_bounds__(128) swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ uv, const float* __restrict__ suv,
                                                         float suv_mul, __nv_bfloat16* __restrict__ xo, int M, int F,
                                                         int rows_per_slab) {
  pdl_enter();
  const int c8 = (blockIdx.x * 128 + threadIdx.x) * 8;
  if (c8 >= F) return;
  float su[8], sv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    su[e] = suv ? suv[c8 + e] * suv_mul : 1.f;
    sv[e] = suv ? suv[F + c8 + e] * suv_mul : 1.f;
  }
  const int r0 = blockIdx.y * rows_per_slab, r1 = min(M, r0 + rows_per_slab);
  for (int r = r0; r < r1; ++r) {
    const __nv_bfloat16* row = uv + 2ll * r * F;
    const uint4 uu = ldg_u4_stream(row + c8), vv = ldg_u4_stream(row + F + c8);
    const uint32_t ua[4] = {uu.x, uu.y, uu.z, uu.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float u0 = bf16lo(ua[e]) * su[2 * e], u1 = bf16hi(ua[e]) * su[2 * e + 1];
      const float v0 = bf16lo(va[e]) * sv[2 * e], v1 = bf16hi(va[e]) * sv[2 * e + 1];
      o[e] = pack_bf16(u0 * v0 * sigmoidf_(v0), u1 * v1 * sigmoidf_(v1));
    }
    *reinterpret_cast<uint4*>(xo + 1ll * r * F + c8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(128) swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ dx, const __nv_bfloat16* __restrict__ uv,
                                                         const float* __restrict__ suv, float suv_mul,
                                                         __nv_bfloat16* __restrict__ duv, float* __restrict__ dsuv, int M, int F,
                                                         int rows_per_slab) {
  pdl_enter();
  const int c8 = (blockIdx.x * 128 + threadIdx.x) * 8;
  if (c8 >= F) return;
  float su[8], sv[8], gsu[8], gsv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    su[e] = suv ? suv[c8 + e] * suv_mul : 1.f;
    sv[e] = suv ? suv[F + c8 + e] * suv_mul : 1.f;
    gsu[e] = 0.f;
    gsv[e] = 0.f;
  }
  const int r0 = blockIdx.y * rows_per_slab, r1 = min(M, r0 + rows_per_slab);
  constexpr int U = 4;  // rows in flight per thread: 12 independent 16-byte loads before the first use
  for (int rb = r0; rb < r1; rb += U) {
    uint4 uu[U], vv[U], dd[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int r = min(rb + k, r1 - 1);
      const __nv_bfloat16* row = uv + 2ll * r * F;
      uu[k] = ldg_u4_stream(row + c8);
      vv[k] = ldg_u4_stream(row + F + c8);
      dd[k] = ldg_u4_stream(dx + 1ll * r * F + c8);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int r = rb + k;
      if (r >= r1) break;
      const uint32_t ua[4] = {uu[k].x, uu[k].y, uu[k].z, uu[k].w}, va[4] = {vv[k].x, vv[k].y, vv[k].z, vv[k].w},
                     da[4] = {dd[k].x, dd[k].y, dd[k].z, dd[k].w};
      uint32_t ou[4], ov[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float du_[2], dv_[2];
#pragma unroll
        for (int hgh = 0; hgh < 2; ++hgh) {
          const int kk = 2 * e + hgh;
          const float ur = hgh ? bf16hi(ua[e]) : bf16lo(ua[e]);
          const float vr = hgh ? bf16hi(va[e]) : bf16lo(va[e]);
          const float g = hgh ? bf16hi(da[e]) : bf16lo(da[e]);
          const float u = ur * su[kk], v = vr * sv[kk];
          const float sg = sigmoidf_(v);
          const float gu = g * v * sg;                           // dL/d(u scaled)
          const float gv = g * u * sg * (1.f + v * (1.f - sg));  // dL/d(v scaled)
          gsu[kk] += gu * ur;
          gsv[kk] += gv * vr;
          du_[hgh] = gu * su[kk];
          dv_[hgh] = gv * sv[kk];
        }
        ou[e] = pack_bf16(du_[0], du_[1]);
        ov[e] = pack_bf16(dv_[0], dv_[1]);
      }
      __nv_bfloat16* drow = duv + 2ll * r * F;
      *reinterpret_cast<uint4*>(drow + c8) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
      *reinterpret_cast<uint4*>(drow + F + c8) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
    }
  }
  if (dsuv) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(dsuv + c8 + e, gsu[e] * suv_mul);
      atomicAdd(dsuv + F + c8 + e, gsv[e] * suv_mul);
    }
  }
}

// ------------------------------------------------------------------------------------------------ dL/dsuv from dL/dW
// uv = (suv mul) * (h W^T)  =>  dL/dsuv[c] = sum_k W[c,k] dW[c,k] / suv[c]  (dW = dL/dW of the same step, any number of
// accumulated micro-batches), which replaces a column reduction over all M rows of d(uv) * uv.  One warp per row.
__global__ void __launch_bounds__(256) rowdot_div_kernel(const float* __restrict__ w, const float* __restrict__ dw,
                                                         const float* __restrict__ div, float* __restrict__ out, int rows, int cols) {
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int r = blockIdx.x * nwarp + warp; r < rows; r += gridDim.x * nwarp) {
    const float* a = w + 1ll * r * cols;
    const float* b = dw + 1ll * r * cols;
    float acc = 0.f;
    for (int k = lane; k < cols; k += 32) acc += a[k] * b[k];
    acc = warp_sum(acc);
    if (lane == 0) {
      const float d = div[r];
      out[r] = d != 0.f ? acc / d : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------ im2col
__device__ __forceinline__ int reflect_idx(int i, int S) {
  if (i < 0) i = -i;
  if (i >= S) i = 2 * (S - 1) - i;
  return i;
}
// One thread per 8 consecutive kw (one 16-byte store) when ks % 8 == 0, else per pair (kw even):
//   out[(b,i,j), (c,kh,kw)] = img[b, c, refl(i*st + kh - pad), refl(j*st + kw - pad)]
template <int V>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int ch,
                                                     int S, int ks, int st, int pad, int g, long long total_vec) {
  pdl_enter();
  const int K = ch * ks * ks;
  const int vecK = K / V;
  for (long long idx = 1ll * blockIdx.x * blockDim.x + threadIdx.x; idx < total_vec; idx += 1ll * gridDim.x * blockDim.x) {
    const long long m = idx / vecK;
    const int k = (int)(idx - m * vecK) * V;
    const int c = k / (ks * ks);
    const int rem = k - c * ks * ks;
    const int kh = rem / ks, kw = rem - kh * ks;
    const int b = (int)(m / (g * g));
    const int ij = (int)(m - 1ll * b * g * g);
    const int i = ij / g, j = ij - i * g;
    const int y = reflect_idx(i * st + kh - pad, S);
    const int x0 = j * st + kw - pad;
    const float* src = img + ((1ll * b * ch + c) * S + y) * S;
    float v[V];
    if (V == 8 && x0 >= 0 && x0 + 8 <= S && ((x0 & 3) == 0) && ((S & 3) == 0)) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(src + x0)), a1 = __ldg(reinterpret_cast<const float4*>(src + x0 + 4));
      v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w;
      if (V == 8) { v[V - 4] = a1.x; v[V - 3] = a1.y; v[V - 2] = a1.z; v[V - 1] = a1.w; }
    } else {
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] = __ldg(src + reflect_idx(x0 + e, S));
    }
    if (V == 8) {
      *reinterpret_cast<uint4*>(out + m * K + k) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[V - 4], v[V - 3]), pack_bf16(v[V - 2], v[V - 1]));
    } else {
      *reinterpret_cast<uint32_t*>(out + m * K + k) = pack_bf16(v[0], v[1]);
    }
  }
}

// The same operand straight from the loader's uint8 HWC images (torchvision ToTensor + Normalize(mean, std) of
// train.py:266-273, 1081-1092 folded in: value = pixel * scale + shift), so the host sends 1 byte per sample instead of 4:
//   out[(b,i,j), (c,kh,kw)] = img[b, refl(i*st + kh - pad), refl(j*st + kw - pad), c] * scale + shift
__global__ void __launch_bounds__(256) im2col_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int ch,
                                                        int S, int ks, int st, int pad, int g, long long total_pairs, float scale,
                                                        float shift) {
  pdl_enter();
  const int K = ch * ks * ks;
  const int halfK = K / 2;
  for (long long idx = 1ll * blockIdx.x * blockDim.x + threadIdx.x; idx < total_pairs; idx += 1ll * gridDim.x * blockDim.x) {
    const long long m = idx / halfK;
    const int k = (int)(idx - m * halfK) * 2;
    const int c = k / (ks * ks);
    const int rem = k - c * ks * ks;
    const int kh = rem / ks, kw = rem - kh * ks;
    const int b = (int)(m / (g * g));
    const int ij = (int)(m - 1ll * b * g * g);
    const int i = ij / g, j = ij - i * g;
    const int y = reflect_idx(i * st + kh - pad, S);
    const int x0 = j * st + kw - pad;
    const uint8_t* src = img + ((1ll * b * S + y) * S) * ch + c;
    const float v0 = (float)__ldg(src + 1ll * reflect_idx(x0, S) * ch) * scale + shift;
    const float v1 = (float)__ldg(src + 1ll * reflect_idx(x0 + 1, S) * ch) * scale + shift;
    *reinterpret_cast<uint32_t*>(out + m * K + k) = pack_bf16(v0, v1);
  }
}

// ------------------------------------------------------------------------------------------------ pooled LayerNorm head
// One CTA per image: mean over T (coalesced over C), LayerNorm over C; saves xhat and rstd for backward.
__global__ void __launch_bounds__(256) pool_ln_fwd_kernel(const float* __restrict__ h, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y,
                                                          float* __restrict__ xhat, float* __restrict__ rstd, int T, int C) {
  pdl_enter();
  extern __shared__ float s_pool[];  // [C]
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float invT = 1.f / (float)T;
  float lsum = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    const float* p = h + (1ll * b * T) * C + c;
    for (int t = 0; t < T; ++t) acc += p[1ll * t * C];
    acc *= invT;
    s_pool[c] = acc;
    lsum += acc;
  }
  const float mean = block_sum(lsum, red) / (float)C;
  float lvar = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float d = s_pool[c] - mean;
    lvar += d * d;
  }
  const float var = block_sum(lvar, red) / (float)C;
  const float rs = rsqrtf(var + eps);
  if (threadIdx.x == 0) rstd[b] = rs;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float xh = (s_pool[c] - mean) * rs;
    xhat[1ll * b * C + c] = xh;
    y[1ll * b * C + c] = __float2bfloat16(xh * gamma[c] + beta[c]);
  }
}

__global__ void __launch_bounds__(256) pool_ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ gamma,
                                                          const float* __restrict__ xhat, const float* __restrict__ rstd,
                                                          float* __restrict__ dh, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, int T, int C) {
  pdl_enter();
  extern __shared__ float s_dp[];  // [C]
  __shared__ float red[32];
  const int b = blockIdx.x;
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = __bfloat162float(dy[1ll * b * C + c]);
    const float xh = xhat[1ll * b * C + c];
    const float gx = g * gamma[c];
    s1 += gx;
    s2 += gx * xh;
    atomicAdd(dgamma + c, g * xh);
    atomicAdd(dbeta + c, g);
  }
  const float m1 = block_sum(s1, red) / (float)C;
  const float m2 = block_sum(s2, red) / (float)C;
  const float rs = rstd[b];
  const float invT = 1.f / (float)T;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = __bfloat162float(dy[1ll * b * C + c]);
    const float xh = xhat[1ll * b * C + c];
    s_dp[c] = (g * gamma[c] - m1 - xh * m2) * rs * invT;
  }
  __syncthreads();
  const int C4 = C >> 2;  // C % 4 == 0 checked on the host
  float4* out = reinterpret_cast<float4*>(dh + 1ll * b * T * C);
  const float4* sp = reinterpret_cast<const float4*>(s_dp);
  for (int i = threadIdx.x; i < T * C4; i += blockDim.x) out[i] = sp[i % C4];
}

// logits = raw * sz_eff backward
__global__ void __launch_bounds__(256) head_scale_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ raw,
                                                             const float* __restrict__ sz, float sz_mul,
                                                             __nv_bfloat16* __restrict__ draw, float* __restrict__ dsz, int B, int N,
                                                             long long ld) {
  pdl_enter();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float s = sz ? sz[n] * sz_mul : 1.f;
  float acc = 0.f;
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const float g = dlogits[1ll * b * N + n];
    acc += g * raw[1ll * b * N + n];
    draw[1ll * b * ld + n] = __float2bfloat16(g * s);
  }
  if (dsz) atomicAdd(dsz + n, acc * sz_mul);
}

// logits = raw * sz_eff (model.py:466-468): the class-head GEMM runs once, its scaled copy is this B x N pass
__global__ void __launch_bounds__(256) head_scale_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ sz, float sz_mul,
                                                             float* __restrict__ logits, long long total, int N) {
  pdl_enter();
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 1ll * gridDim.x * blockDim.x)
    logits[i] = raw[i] * (sz[i % N] * sz_mul);
}

// softmax cross-entropy, one CTA per sample
__global__ void __launch_bounds__(256) cross_entropy_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                            float* __restrict__ loss, float* __restrict__ dlogits, float gscale, int B,
                                                            int N) {
  pdl_enter();
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float* row = logits + 1ll * b * N;
  float mx = -INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, row[n]);
  mx = block_max(mx, red);
  float se = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) se += __expf(row[n] - mx);
  se = block_sum(se, red);
  const float lse = mx + __logf(se);
  const long long t64 = target[b];
  const bool valid = t64 >= 0 && t64 < N;      // out of range (e.g. ignore_index -100): no loss, zero gradient row
  const int t = valid ? (int)t64 : -1;
  if (threadIdx.x == 0 && loss && valid) atomicAdd(loss, (lse - row[t]) / (float)B);
  if (dlogits) {
    const float k = valid ? gscale / (float)B : 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x)
      dlogits[1ll * b * N + n] = (__expf(row[n] - lse) - (n == t ? 1.f : 0.f)) * k;
  }
}

__global__ void __launch_bounds__(256) tanh_mse_kernel(const __nv_bfloat16* __restrict__ pred, const __nv_bfloat16* __restrict__ tgt,
                                                       long long n, float inv_count, float* __restrict__ out) {
  pdl_enter();
  __shared__ float red[32];
  const long long n8 = n >> 3;
  const long long stride = 1ll * gridDim.x * blockDim.x;
  float acc = 0.f;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 p = ldg_u4_stream(pred + i * 8), t = ldg_u4_stream(tgt + i * 8);
    const uint32_t pa[4] = {p.x, p.y, p.z, p.w}, ta[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d0 = tanhf(bf16lo(pa[e])) - bf16lo(ta[e]);
      const float d1 = tanhf(bf16hi(pa[e])) - bf16hi(ta[e]);
      acc += d0 * d0 + d1 * d1;
    }
  }
  for (long long i = n8 * 8 + 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = tanhf(__bfloat162float(pred[i])) - __bfloat162float(tgt[i]);
    acc += d * d;
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, tot * inv_count);
}

// ------------------------------------------------------------------------------------------------ AdamW (flat)
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long n, long long n_decay, float lr, float b1,
                                                         float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                         const float* __restrict__ gnorm_sq, float max_norm,
                                                         const float* __restrict__ dev_lr_step) {
  pdl_enter();
  if (dev_lr_step) {   // CUDA-graph friendly: learning rate and 1-based step count live in device memory
    lr = dev_lr_step[0];
    const float t = dev_lr_step[1];
    bc1 = 1.f - powf(b1, t);
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  float clip = 1.f;
  if (gnorm_sq) {
    const float tot = sqrtf(gnorm_sq[0]);
    clip = fminf(1.f, max_norm / (tot + 1e-6f));
  }
  const float step_size = lr / bc1;
  const float decay = 1.f - lr * wd;
  const long long n4 = n >> 2;
  const long long stride = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = *reinterpret_cast<const float4*>(p + i * 4);
    const float4 gg = ldg_f4_stream(g + i * 4);
    float4 mm = *reinterpret_cast<const float4*>(m + i * 4);
    float4 vv = *reinterpret_cast<const float4*>(v + i * 4);
    float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ga[4] = {gg.x, gg.y, gg.z, gg.w};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = ga[e] * clip;
      if (i * 4 + e < n_decay) pa[e] *= decay;
      ma[e] = b1 * ma[e] + (1.f - b1) * gr;
      va[e] = b2 * va[e] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(va[e]) / bc2_sqrt + eps;
      pa[e] -= step_size * (ma[e] / denom);
    }
    *reinterpret_cast<float4*>(p + i * 4) = make_float4(pa[0], pa[1], pa[2], pa[3]);
    *reinterpret_cast<float4*>(m + i * 4) = make_float4(ma[0], ma[1], ma[2], ma[3]);
    *reinterpret_cast<float4*>(v + i * 4) = make_float4(va[0], va[1], va[2], va[3]);
  }
  for (long long i = n4 * 4 + 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gr = g[i] * clip;
    float pe = p[i];
    if (i < n_decay) pe *= decay;
    const float me = b1 * m[i] + (1.f - b1) * gr;
    const float ve = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = me;
    v[i] = ve;
    p[i] = pe - step_size * (me / (sqrtf(ve) / bc2_sqrt + eps));
  }
}

// ------------------------------------------------------------------------------------------------ weight normalization
// table entry: {w ptr, w16 ptr, rows, cols, axis, first_unit}.  256 threads per CTA, one unit per CTA iteration.
//   axis 1: unit = 8 rows, one warp per row held in registers (float4 per lane, cols <= 4096) -> one read, one write.
//   axis 0: unit = 128 columns; thread (lane -> 4 columns as a float4, warp -> row phase); the second pass re-reads the
//           slab (rows x 512 B, L2 resident) to scale it.
__global__ void __launch_bounds__(256) weight_norm_multi_kernel(const long long* __restrict__ table, int n_tensors,
                                                                long long total_units) {
  pdl_enter();
  __shared__ float4 s_part[8][32];
  __shared__ float4 s_inv[32];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  for (long long unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    int lo = 0, hi = n_tensors - 1;  // last tensor with first_unit <= unit
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (table[mid * 6 + 5] <= unit) lo = mid; else hi = mid - 1;
    }
    const long long* e = table + lo * 6;
    float* w = reinterpret_cast<float*>(e[0]);
    __nv_bfloat16* w16 = reinterpret_cast<__nv_bfloat16*>(e[1]);
    const int rows = (int)e[2], cols = (int)e[3], axis = (int)e[4];
    const int u = (int)(unit - e[5]);
    const bool vec = ((cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
    if (axis == 1) {
      const int r = u * 8 + wy;
      if (r < rows) {
        float* row = w + 1ll * r * cols;
        if (vec && cols <= 1024) {
          float4 v[8];
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = (j * 32 + lane) * 4;
            v[j] = (c < cols) ? *reinterpret_cast<const float4*>(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
          }
          ss = warp_sum(ss);
          const float inv = 1.f / sqrtf(ss);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = (j * 32 + lane) * 4;
            if (c < cols) {
              const float4 o = make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv);
              *reinterpret_cast<float4*>(row + c) = o;
              if (w16) *reinterpret_cast<uint2*>(w16 + 1ll * r * cols + c) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
            }
          }
        } else {
          float ss = 0.f;
          for (int c = lane; c < cols; c += 32) { const float x = row[c]; ss += x * x; }
          ss = warp_sum(ss);
          const float inv = 1.f / sqrtf(ss);
          for (int c = lane; c < cols; c += 32) {
            const float x = row[c] * inv;
            row[c] = x;
            if (w16) w16[1ll * r * cols + c] = __float2bfloat16(x);
          }
        }
      }
    } else {
      const int c = u * 128 + lane * 4;
      float4 ss = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vec) {
        if (c < cols) {
          int r = wy;
          for (; r + 24 < rows; r += 32) {   // four independent loads in flight
            const float4 a0 = *reinterpret_cast<const float4*>(w + 1ll * r * cols + c);
            const float4 a1 = *reinterpret_cast<const float4*>(w + 1ll * (r + 8) * cols + c);
            const float4 a2 = *reinterpret_cast<const float4*>(w + 1ll * (r + 16) * cols + c);
            const float4 a3 = *reinterpret_cast<const float4*>(w + 1ll * (r + 24) * cols + c);
            ss.x += a0.x * a0.x + a1.x * a1.x + a2.x * a2.x + a3.x * a3.x;
            ss.y += a0.y * a0.y + a1.y * a1.y + a2.y * a2.y + a3.y * a3.y;
            ss.z += a0.z * a0.z + a1.z * a1.z + a2.z * a2.z + a3.z * a3.z;
            ss.w += a0.w * a0.w + a1.w * a1.w + a2.w * a2.w + a3.w * a3.w;
          }
          for (; r < rows; r += 8) {
            const float4 a0 = *reinterpret_cast<const float4*>(w + 1ll * r * cols + c);
            ss.x += a0.x * a0.x; ss.y += a0.y * a0.y; ss.z += a0.z * a0.z; ss.w += a0.w * a0.w;
          }
        }
      } else {
        float* sp = reinterpret_cast<float*>(&ss);
        for (int k = 0; k < 4; ++k)
          if (c + k < cols)
            for (int r = wy; r < rows; r += 8) { const float x = w[1ll * r * cols + c + k]; sp[k] += x * x; }
      }
      s_part[wy][lane] = ss;
      __syncthreads();
      if (wy == 0) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float4 q = s_part[k][lane]; t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
        s_inv[lane] = make_float4(1.f / sqrtf(t.x), 1.f / sqrtf(t.y), 1.f / sqrtf(t.z), 1.f / sqrtf(t.w));
      }
      __syncthreads();
      const float4 inv = s_inv[lane];
      if (vec) {
        if (c < cols) {
          for (int r = wy; r < rows; r += 8) {
            float4 a = *reinterpret_cast<const float4*>(w + 1ll * r * cols + c);
            a.x *= inv.x; a.y *= inv.y; a.z *= inv.z; a.w *= inv.w;
            *reinterpret_cast<float4*>(w + 1ll * r * cols + c) = a;
            if (w16) *reinterpret_cast<uint2*>(w16 + 1ll * r * cols + c) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
          }
        }
      } else {
        const float* ip = reinterpret_cast<const float*>(&inv);
        for (int k = 0; k < 4; ++k)
          if (c + k < cols)
            for (int r = wy; r < rows; r += 8) {
              const float x = w[1ll * r * cols + c + k] * ip[k];
              w[1ll * r * cols + c + k] = x;
              if (w16) w16[1ll * r * cols + c + k] = __float2bfloat16(x);
            }
      }
      __syncthreads();
    }
  }
}

}  // namespace nvit

using namespace nvit;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int nvit_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  NVIT_REQUIRE(n >= 0 && (n == 0 || (src && dst)), "nvit_cast_f32_to_bf16: bad arguments");
  if (n == 0) return NVIT_OK;
  launch(cast_f32_bf16_kernel, stream_grid(n / 8 + 1, 256), 256, 0, ST(stream), src, static_cast<__nv_bfloat16*>(dst), n);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_sumsq_f32(const float* x, int64_t n, float* out_accum, void* stream) {
  NVIT_REQUIRE(n >= 0 && out_accum && (n == 0 || x), "nvit_sumsq_f32: bad arguments");
  if (n == 0) return NVIT_OK;
  launch(sumsq_kernel, stream_grid(n / 4 + 1, 256, 4), 256, 0, ST(stream), x, n, out_accum);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_sumsq_f32_det(const float* x, int64_t n, float* out_accum, float* workspace, int64_t workspace_floats, void* stream) {
  NVIT_REQUIRE(n >= 0 && out_accum && workspace && (n == 0 || x), "nvit_sumsq_f32_det: bad arguments");
  if (n == 0) return NVIT_OK;
  int grid = stream_grid(n / 4 + 1, 256, 4);
  if (grid > workspace_floats - 1) grid = (int)(workspace_floats - 1);
  NVIT_REQUIRE(grid >= 1, "nvit_sumsq_f32_det: workspace too small");
  launch(sumsq_det_kernel, grid, 256, 0, ST(stream), x, n, out_accum, workspace);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_colsum_bf16(const void* x, int64_t M, int64_t N, int64_t ldx, float* out_accum, void* stream) {
  NVIT_REQUIRE(x && out_accum && M >= 0 && N > 0 && ldx >= N, "nvit_colsum_bf16: bad arguments");
  if (M == 0) return NVIT_OK;
  const int gx = (int)((N + 255) / 256);
  int slabs = (nvit_num_sms() * 4 + gx - 1) / gx;
  if (slabs > M) slabs = (int)M;
  const int rps = (int)((M + slabs - 1) / slabs);
  dim3 grid(gx, (unsigned)((M + rps - 1) / rps));
  launch(colsum_bf16_kernel, grid, 256, 0, ST(stream), static_cast<const __nv_bfloat16*>(x), (int)M, (int)N, ldx, out_accum, rps);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_pos_bias_grad(const float* dx, int64_t B, int64_t T, int64_t C, float* dpos, float* dbias_accum, void* stream) {
  NVIT_REQUIRE(dx && dpos && B > 0 && T > 0 && C > 0, "nvit_pos_bias_grad: bad arguments");
  dim3 grid((unsigned)T, (unsigned)((C + 255) / 256));
  launch(pos_bias_grad_kernel, grid, 256, 0, ST(stream), dx, (int)B, (int)T, (int)C, dpos, dbias_accum);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

static void slab_grid(int64_t M, int64_t F, dim3* grid, int* rps) {
  const int gx = (int)((F / 8 + 127) / 128);
  int slabs = (nvit_num_sms() * 16 + gx - 1) / gx;
  if (slabs > M) slabs = (int)M;
  if (slabs < 1) slabs = 1;
  *rps = (int)((M + slabs - 1) / slabs);
  *grid = dim3(gx, (unsigned)((M + *rps - 1) / *rps));
}

extern "C" int nvit_swiglu_fwd(const void* uv, const float* suv, float suv_mul, void* x, int64_t M, int64_t F, void* stream) {
  NVIT_REQUIRE(uv && x && M >= 0 && F > 0 && (F % 8) == 0, "nvit_swiglu_fwd: F=%lld must be a positive multiple of 8", (long long)F);
  if (M == 0) return NVIT_OK;
  dim3 grid; int rps;
  slab_grid(M, F, &grid, &rps);
  launch(swiglu_fwd_kernel, grid, 128, 0, ST(stream), static_cast<const __nv_bfloat16*>(uv), suv, suv_mul, static_cast<__nv_bfloat16*>(x), (int)M, (int)F, rps);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_swiglu_bwd(const void* dx, const void* uv, const float* suv, float suv_mul, void* duv, float* dsuv_accum,
                               int64_t M, int64_t F, void* stream) {
  NVIT_REQUIRE(dx && uv && duv && M >= 0 && F > 0 && (F % 8) == 0, "nvit_swiglu_bwd: bad arguments (F=%lld)", (long long)F);
  NVIT_REQUIRE((suv == nullptr) == (dsuv_accum == nullptr), "nvit_swiglu_bwd: suv and dsuv go together");
  if (M == 0) return NVIT_OK;
  dim3 grid; int rps;
  slab_grid(M, F, &grid, &rps);
  launch(swiglu_bwd_kernel, grid, 128, 0, ST(stream), static_cast<const __nv_bfloat16*>(dx), static_cast<const __nv_bfloat16*>(uv), suv, suv_mul,
                                                  static_cast<__nv_bfloat16*>(duv), dsuv_accum, (int)M, (int)F, rps);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_rowdot_div(const float* w, const float* dw, const float* div, float* out, int64_t rows, int64_t cols, void* stream) {
  NVIT_REQUIRE(w && dw && div && out && rows >= 0 && cols > 0, "nvit_rowdot_div: bad arguments");
  if (rows == 0) return NVIT_OK;
  launch(rowdot_div_kernel, stream_grid(rows * 32, 256), 256, 0, ST(stream), w, dw, div, out, (int)rows, (int)cols);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_im2col_bf16(const float* img, void* out, int64_t B, int64_t ch, int64_t S, int64_t ksize, int64_t stride,
                                int64_t pad, void* stream) {
  NVIT_REQUIRE(img && out && B > 0 && ch > 0 && S > 0 && ksize > 0 && stride > 0 && pad >= 0, "nvit_im2col_bf16: bad arguments");
  NVIT_REQUIRE((ksize % 2) == 0, "nvit_im2col_bf16: kernel size must be even");
  NVIT_REQUIRE(pad < S, "nvit_im2col_bf16: reflect padding must be smaller than the image");
  NVIT_REQUIRE((S + 2 * pad - ksize) % stride == 0, "nvit_im2col_bf16: (S + 2 pad - k) must be a multiple of the stride");
  const int g = (int)((S + 2 * pad - ksize) / stride + 1);
  const long long total = 1ll * B * g * g * ch * ksize * ksize;
  if ((ksize % 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    launch(im2col_kernel<8>, stream_grid(total / 8, 256, 16), 256, 0, ST(stream), img, static_cast<__nv_bfloat16*>(out), (int)B, (int)ch, (int)S,
                                                                             (int)ksize, (int)stride, (int)pad, g, total / 8);
  else
    launch(im2col_kernel<2>, stream_grid(total / 2, 256, 16), 256, 0, ST(stream), img, static_cast<__nv_bfloat16*>(out), (int)B, (int)ch, (int)S,
                                                                             (int)ksize, (int)stride, (int)pad, g, total / 2);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_im2col_u8(const void* img_u8_nhwc, void* out, int64_t B, int64_t ch, int64_t S, int64_t ksize, int64_t stride,
                              int64_t pad, float scale, float shift, void* stream) {
  NVIT_REQUIRE(img_u8_nhwc && out && B > 0 && ch > 0 && S > 0 && ksize > 0 && stride > 0 && pad >= 0, "nvit_im2col_u8: bad arguments");
  NVIT_REQUIRE((ksize % 2) == 0, "nvit_im2col_u8: kernel size must be even");
  NVIT_REQUIRE(pad < S, "nvit_im2col_u8: reflect padding must be smaller than the image");
  NVIT_REQUIRE((S + 2 * pad - ksize) % stride == 0, "nvit_im2col_u8: (S + 2 pad - k) must be a multiple of the stride");
  const int g = (int)((S + 2 * pad - ksize) / stride + 1);
  const long long total = 1ll * B * g * g * ch * ksize * ksize;
  launch(im2col_u8_kernel, stream_grid(total / 2, 256, 16), 256, 0, ST(stream), static_cast<const uint8_t*>(img_u8_nhwc), static_cast<__nv_bfloat16*>(out),
                                                                           (int)B, (int)ch, (int)S, (int)ksize, (int)stride, (int)pad, g, total / 2,
                                                                           scale, shift);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_pool_ln_fwd(const float* h, const float* gamma, const float* beta, float eps, void* y, float* xhat, float* rstd,
                                int64_t B, int64_t T, int64_t C, void* stream) {
  NVIT_REQUIRE(h && gamma && beta && y && xhat && rstd && B > 0 && T > 0 && C > 0, "nvit_pool_ln_fwd: bad arguments");
  NVIT_REQUIRE(C * sizeof(float) <= 48 * 1024, "nvit_pool_ln_fwd: C too large");
  launch(pool_ln_fwd_kernel, (unsigned)B, 256, C * sizeof(float), ST(stream), h, gamma, beta, eps, static_cast<__nv_bfloat16*>(y), xhat, rstd, (int)T, (int)C);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_pool_ln_bwd(const void* dy, const float* gamma, const float* xhat, const float* rstd, float* dh, float* dgamma,
                                float* dbeta, int64_t B, int64_t T, int64_t C, void* stream) {
  NVIT_REQUIRE(dy && gamma && xhat && rstd && dh && dgamma && dbeta && B > 0 && T > 0 && C > 0, "nvit_pool_ln_bwd: bad arguments");
  NVIT_REQUIRE((C % 4) == 0 && C * sizeof(float) <= 48 * 1024, "nvit_pool_ln_bwd: C must be a multiple of 4 and fit shared memory");
  launch(pool_ln_bwd_kernel, (unsigned)B, 256, C * sizeof(float), ST(stream), static_cast<const __nv_bfloat16*>(dy), gamma, xhat, rstd, dh, dgamma, dbeta, (int)T, (int)C);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_head_scale_bwd(const float* dlogits, const float* raw, const float* sz, float sz_mul, void* draw, float* dsz,
                                   int64_t B, int64_t N, int64_t ld_draw, void* stream) {
  NVIT_REQUIRE(dlogits && raw && draw && B > 0 && N > 0 && ld_draw >= N, "nvit_head_scale_bwd: bad arguments");
  launch(head_scale_bwd_kernel, dim3((unsigned)((N + 255) / 256), (unsigned)(B < 32 ? B : 32)), 256, 0, ST(stream), dlogits, raw, sz, sz_mul, static_cast<__nv_bfloat16*>(draw), dsz, (int)B, (int)N, ld_draw);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_head_scale_fwd(const float* raw, const float* sz, float sz_mul, float* logits, int64_t B, int64_t N, void* stream) {
  NVIT_REQUIRE(raw && sz && logits && B > 0 && N > 0, "nvit_head_scale_fwd: bad arguments");
  launch(head_scale_fwd_kernel, stream_grid(B * N, 256), 256, 0, ST(stream), raw, sz, sz_mul, logits, (long long)(B * N), (int)N);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_cross_entropy(const float* logits, const int64_t* target, float* loss, float* dlogits, float gscale, int64_t B,
                                  int64_t N, void* stream) {
  NVIT_REQUIRE(logits && target && B > 0 && N > 0 && (loss || dlogits), "nvit_cross_entropy: bad arguments");
  launch(cross_entropy_kernel, (unsigned)B, 256, 0, ST(stream), logits, reinterpret_cast<const long long*>(target), loss, dlogits, gscale, (int)B, (int)N);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_tanh_mse(const void* pred, const void* target, int64_t n, float inv_count, float* out_accum, void* stream) {
  NVIT_REQUIRE(pred && target && out_accum && n >= 0, "nvit_tanh_mse: bad arguments");
  if (n == 0) return NVIT_OK;
  launch(tanh_mse_kernel, stream_grid(n / 8 + 1, 256, 4), 256, 0, ST(stream), static_cast<const __nv_bfloat16*>(pred), static_cast<const __nv_bfloat16*>(target), n, inv_count, out_accum);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, int64_t n_decay, float lr, float beta1,
                               float beta2, float eps, float weight_decay, int64_t step, const float* gnorm_sq, float max_norm,
                               const float* dev_lr_step, void* stream) {
  NVIT_REQUIRE(p && g && m && v && n >= 0 && n_decay >= 0 && n_decay <= n && (step >= 1 || dev_lr_step), "nvit_adamw_flat: bad arguments");
  if (step < 1) step = 1;
  NVIT_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
               "nvit_adamw_flat: buffers must be 16-byte aligned");
  if (n == 0) return NVIT_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  launch(adamw_flat_kernel, stream_grid(n / 4 + 1, 256), 256, 0, ST(stream), p, g, m, v, n, n_decay, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                                       (float)sqrt(bc2), gnorm_sq, max_norm, dev_lr_step);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_weight_norm_multi(const int64_t* table_dev, int64_t n_tensors, int64_t total_units, void* stream) {
  NVIT_REQUIRE(table_dev && n_tensors > 0 && total_units > 0, "nvit_weight_norm_multi: bad arguments");
  const long long cap = 1ll * nvit_num_sms() * 8;
  const int grid = (int)(total_units < cap ? total_units : cap);
  launch(weight_norm_multi_kernel, grid, 256, 0, ST(stream), reinterpret_cast<const long long*>(table_dev), (int)n_tensors, total_units);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}
