// Kohonen-map kernels of the nViT training step (BASELINE config 5).  The distance GEMM itself runs on tcgen05 through
// nvit_gemm_bf16 on bf16 hi/lo splits of the fp32 operands; the kernels here produce those splits, pick the
// best-matching unit per token, apply the sequential in-forward map update and compute the map losses with their
// gradients.  Reference lines (relative to /root/reference) are cited per kernel and in include/nvit_b200.h.
#include "common.cuh"

namespace nvit {

static inline int som_grid(long long work_items, int threads, int per_sm = 8) {
  long long want = (work_items + threads - 1) / threads;
  long long cap = 1ll * nvit_num_sms() * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// ------------------------------------------------------------------------------------------------ fp32 -> bf16 hi + lo
// x = hi + lo + O(2^-17 |x|): two bf16 GEMM operands that together carry ~16 mantissa bits of an fp32 tensor.
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, long long n) {
  pdl_enter();
  const long long n4 = n >> 2;
  const long long stride = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = ldg_f4_stream(x + i * 4);
    const float v[4] = {a.x, a.y, a.z, a.w};
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h[e] = __bfloat162float(__float2bfloat16(v[e]));
      l[e] = v[e] - h[e];
    }
    if (hi) *reinterpret_cast<uint2*>(hi + i * 4) = make_uint2(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]));
    *reinterpret_cast<uint2*>(lo + i * 4) = make_uint2(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]));
  }
  for (long long i = n4 * 4 + 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const __nv_bfloat16 h = __float2bfloat16(x[i]);
    if (hi) hi[i] = h;
    lo[i] = __float2bfloat16(x[i] - __bfloat162float(h));
  }
}

// ------------------------------------------------------------------------------------------------ node table preparation
// One block per node: squared norm, bf16 hi/lo operands and a snapshot of the table as it is BEFORE the in-forward
// update (the representations, the consistency and the quantization losses use the pre-update nodes; model.py:424-445).
__global__ void __launch_bounds__(128) som_prepare_kernel(const float* __restrict__ nodes, int C, float* __restrict__ nn,
                                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                          float* __restrict__ snapshot) {
  pdl_enter();
  __shared__ float red[4];
  const int g = blockIdx.x;
  const float* row = nodes + 1ll * g * C;
  float acc = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = row[c];
    acc += v * v;
    const __nv_bfloat16 h = __float2bfloat16(v);
    hi[1ll * g * C + c] = h;
    lo[1ll * g * C + c] = __float2bfloat16(v - __bfloat162float(h));
    snapshot[1ll * g * C + c] = v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) nn[g] = (red[0] + red[1]) + (red[2] + red[3]);
}

// ------------------------------------------------------------------------------------------------ best-matching unit
// KohonenMap.forward (kohonen.py:100-119): argmin_g ||x - n_g|| = argmin_g (|n_g|^2 - 2 x.n_g); ties -> lowest index.
// One warp per token: picks the unit, copies its node row (fp32 + bf16), writes the one-hot row that later scatters the
// representation gradient back through a GEMM, and counts the unit in a per-block histogram.
__global__ void __launch_bounds__(256) som_select_kernel(const float* __restrict__ dots, const float* __restrict__ nn,
                                                         const float* __restrict__ nodes, long long M, int G, int C,
                                                         int* __restrict__ idx, long long* __restrict__ idx64,
                                                         __nv_bfloat16* __restrict__ onehot, float* __restrict__ counts,
                                                         float* __restrict__ repr32, __nv_bfloat16* __restrict__ repr16) {
  pdl_enter();
  extern __shared__ float s_hist[];   // [G]
  for (int g = threadIdx.x; g < G; g += blockDim.x) s_hist[g] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (long long r = 1ll * blockIdx.x * nwarp + warp; r < M; r += 1ll * gridDim.x * nwarp) {
    const float* d = dots + r * G;
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int g = lane; g < G; g += 32) {
      const float s = nn[g] - 2.f * d[g];
      if (s < best) { best = s; bi = g; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      idx[r] = bi;
      if (idx64) idx64[r] = bi;
      atomicAdd(&s_hist[bi], 1.f);
    }
    if (onehot)
      for (int g = lane; g < G; g += 32) onehot[r * G + g] = __float2bfloat16(g == bi ? 1.f : 0.f);
    const float* src = nodes + 1ll * bi * C;
    for (int c = lane; c < C; c += 32) {
      const float v = src[c];
      repr32[r * C + c] = v;
      repr16[r * C + c] = __float2bfloat16(v);
    }
  }
  __syncthreads();
  if (counts)
    for (int g = threadIdx.x; g < G; g += blockDim.x)
      if (s_hist[g] != 0.f) atomicAdd(counts + g, s_hist[g]);
}

// ------------------------------------------------------------------------------------------------ pooled update inputs
// kohonen.py:149-155: image b's [T, C] patch matrix, flattened, averaged over runs of `run` = T consecutive elements.
__global__ void __launch_bounds__(256) som_pool_kernel(const float* __restrict__ x, long long rows, int run, float* __restrict__ out) {
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float inv = 1.f / (float)run;
  for (long long r = 1ll * blockIdx.x * nwarp + warp; r < rows; r += 1ll * gridDim.x * nwarp) {
    const float* p = x + r * run;
    float acc = 0.f;
    for (int i = lane; i < run; i += 32) acc += p[i];
    acc = warp_sum(acc);
    if (lane == 0) out[r] = acc * inv;
  }
}

// ------------------------------------------------------------------------------------------------ sequential map update
// KohonenMap.update_nodes (kohonen.py:121-165): for i = 0 .. steps-1, in order,
//     nodes[g] += s_i[g] * (v_i - nodes[g]),   s_i[g] = coef * exp(-torus_dist2(g, bmu_i) / (2 sigma^2)),
// coef = learning rate * alpha, read from device memory so that a schedule needs no re-capture.  One block per node
// keeps the whole recurrence of that node in registers; the strengths of a chunk of steps are shared through smem.
constexpr int SOM_CHUNK = 256;
__global__ void __launch_bounds__(256) som_update_kernel(float* __restrict__ nodes, const float* __restrict__ v,
                                                         const int* __restrict__ bmu, int steps, int C, int gm, int gn,
                                                         const float* __restrict__ coef_dev, float two_sigma2) {
  pdl_enter();
  __shared__ float s_str[SOM_CHUNK];
  const int g = blockIdx.x;
  const int gr = g / gn, gc = g % gn;
  const float coef = *coef_dev;
  // up to 4 channels per thread (C <= 1024); larger C loops over channel groups
  for (int c0 = 0; c0 < C; c0 += 4 * blockDim.x) {
    float n[4];
    int cc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      cc[e] = c0 + e * blockDim.x + threadIdx.x;
      n[e] = cc[e] < C ? nodes[1ll * g * C + cc[e]] : 0.f;
    }
    for (int s0 = 0; s0 < steps; s0 += SOM_CHUNK) {
      const int cnt = min(SOM_CHUNK, steps - s0);
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int b = bmu[s0 + i];
        float dr = (float)(gr - b / gn), dc = (float)(gc - b % gn);
        const float dr2 = fminf(dr * dr, fminf((dr + gm) * (dr + gm), (dr - gm) * (dr - gm)));
        const float dc2 = fminf(dc * dc, fminf((dc + gn) * (dc + gn), (dc - gn) * (dc - gn)));
        s_str[i] = coef * expf(-(dr2 + dc2) / two_sigma2);
      }
      __syncthreads();
      for (int i = 0; i < cnt; ++i) {
        const float s = s_str[i];
        const float* vi = v + 1ll * (s0 + i) * C;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (cc[e] < C) n[e] = __fadd_rn(n[e], __fmul_rn(s, __fsub_rn(vi[cc[e]], n[e])));
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (cc[e] < C) nodes[1ll * g * C + cc[e]] = n[e];
  }
}

// ------------------------------------------------------------------------------------------------ per-token map losses
// Per token r with a = local representation, b = global representation, xl / xg = the patch embeddings:
//   consistency  1 - mean_r cos(a, b)                      (model.py:491-500)
//   quantization huber(a, xl), huber(b, xg), delta 1, mean  (model.py:441-442)
// sums[0..2] += sum_r cos, sum huber_l, sum huber_g.  With `w` (device {consistency, local q., global q.} weights already
// times the incoming gradient) the gradients are ADDED to d_a, d_b, d_xl, d_xg (fp32 [M, C]).  One warp per token.
template <int NV>
__global__ void __launch_bounds__(256) som_pair_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       const float* __restrict__ xl, const float* __restrict__ xg, long long M, int C,
                                                       float* __restrict__ sums, const float* __restrict__ w, float* __restrict__ d_a,
                                                       float* __restrict__ d_b, float* __restrict__ d_xl, float* __restrict__ d_xg) {
  pdl_enter();
  __shared__ float red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float t_cos = 0.f, t_hl = 0.f, t_hg = 0.f;
  float wc = 0.f, wl = 0.f, wg = 0.f;
  if (w) {
    wc = w[0] / (float)M;
    wl = w[1] / ((float)M * (float)C);
    wg = w[2] / ((float)M * (float)C);
  }
  for (long long r = 1ll * blockIdx.x * nwarp + warp; r < M; r += 1ll * gridDim.x * nwarp) {
    float4 va[NV], vb[NV], vl[NV], vg[NV];
    float aa = 0.f, bb = 0.f, ab = 0.f, hl = 0.f, hg = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) {
        va[i] = *reinterpret_cast<const float4*>(a + r * C + c);
        vb[i] = *reinterpret_cast<const float4*>(b + r * C + c);
        vl[i] = ldg_f4_stream(xl + r * C + c);
        vg[i] = ldg_f4_stream(xg + r * C + c);
        const float pa[4] = {va[i].x, va[i].y, va[i].z, va[i].w}, pb[4] = {vb[i].x, vb[i].y, vb[i].z, vb[i].w};
        const float pl[4] = {vl[i].x, vl[i].y, vl[i].z, vl[i].w}, pg[4] = {vg[i].x, vg[i].y, vg[i].z, vg[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          aa += pa[e] * pa[e];
          bb += pb[e] * pb[e];
          ab += pa[e] * pb[e];
          const float dl = fabsf(pa[e] - pl[e]), dg = fabsf(pb[e] - pg[e]);
          hl += dl < 1.f ? 0.5f * dl * dl : dl - 0.5f;
          hg += dg < 1.f ? 0.5f * dg * dg : dg - 0.5f;
        }
      }
    }
    aa = warp_sum(aa);
    bb = warp_sum(bb);
    ab = warp_sum(ab);
    const float ia = rsqrtf(aa), ib = rsqrtf(bb);
    const float cs = ab * ia * ib;
    t_cos += cs;
    t_hl += hl;
    t_hg += hg;
    if (w) {
      // d(-wc cos)/da = -wc (b ia ib - cos a ia^2), symmetric for b
      const float ka = -wc * ia * ib, kaa = wc * cs * ia * ia, kbb = wc * cs * ib * ib;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < C) {
          const float pa[4] = {va[i].x, va[i].y, va[i].z, va[i].w}, pb[4] = {vb[i].x, vb[i].y, vb[i].z, vb[i].w};
          const float pl[4] = {vl[i].x, vl[i].y, vl[i].z, vl[i].w}, pg[4] = {vg[i].x, vg[i].y, vg[i].z, vg[i].w};
          float4 ga = *reinterpret_cast<float4*>(d_a + r * C + c), gb = *reinterpret_cast<float4*>(d_b + r * C + c);
          float4 gl = *reinterpret_cast<float4*>(d_xl + r * C + c), gg = *reinterpret_cast<float4*>(d_xg + r * C + c);
          float* qa = &ga.x;
          float* qb = &gb.x;
          float* ql = &gl.x;
          float* qg = &gg.x;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float hl_g = wl * fminf(fmaxf(pa[e] - pl[e], -1.f), 1.f);
            const float hg_g = wg * fminf(fmaxf(pb[e] - pg[e], -1.f), 1.f);
            qa[e] += ka * pb[e] + kaa * pa[e] + hl_g;
            qb[e] += ka * pa[e] + kbb * pb[e] + hg_g;
            ql[e] -= hl_g;
            qg[e] -= hg_g;
          }
          *reinterpret_cast<float4*>(d_a + r * C + c) = ga;
          *reinterpret_cast<float4*>(d_b + r * C + c) = gb;
          *reinterpret_cast<float4*>(d_xl + r * C + c) = gl;
          *reinterpret_cast<float4*>(d_xg + r * C + c) = gg;
        }
      }
    }
  }
  // the cosine is per row and lives in every lane; the Huber partials are per lane
  t_hl = warp_sum(t_hl);
  t_hg = warp_sum(t_hg);
  if (lane == 0) { red[0][warp] = t_cos; red[1][warp] = t_hl; red[2][warp] = t_hg; }
  __syncthreads();
  if (threadIdx.x < 3 && sums) {
    float t = 0.f;
    for (int i = 0; i < nwarp; ++i) t += red[threadIdx.x][i];
    atomicAdd(sums + threadIdx.x, t);
  }
}

// ------------------------------------------------------------------------------------------------ map smoothness
// model.py:503-561: mean over tokens and the 8 torus neighbours of || n[bmu] - n[neighbour] ||.  Only `counts[g]`
// (tokens whose unit is g) matters: loss = sum_g counts[g] sum_k D(g,k) / (8 M).  One block per (unit, neighbour):
// adds its term to *loss and, with `w`, the gradient to gnodes (both rows of the pair).
__global__ void __launch_bounds__(128) som_smooth_kernel(const float* __restrict__ nodes, const float* __restrict__ counts, int side,
                                                         int C, float inv_8m, float* __restrict__ loss, const float* __restrict__ w,
                                                         float* __restrict__ gnodes) {
  pdl_enter();
  __shared__ float red[4];
  const int g = blockIdx.x >> 3, k = blockIdx.x & 7;
  const float cnt = counts[g];
  if (cnt == 0.f) return;
  const int kk = k < 4 ? k : k + 1;                 // skip the centre of the 3x3 stencil
  const int dr = kk / 3 - 1, dc = kk % 3 - 1;
  const int r = (g / side + dr + side) % side, c = (g % side + dc + side) % side;
  const int nb = r * side + c;
  const float* pa = nodes + 1ll * g * C;
  const float* pb = nodes + 1ll * nb * C;
  float acc = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float d = pa[i] - pb[i];
    acc += d * d;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  const float dist = sqrtf((red[0] + red[1]) + (red[2] + red[3]));
  if (threadIdx.x == 0) atomicAdd(loss, cnt * dist * inv_8m);
  if (w && gnodes && dist > 0.f) {
    const float k2 = w[0] * cnt * inv_8m / dist;
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      const float d = k2 * (pa[i] - pb[i]);
      atomicAdd(gnodes + 1ll * g * C + i, d);
      atomicAdd(gnodes + 1ll * nb * C + i, -d);
    }
  }
}

// ------------------------------------------------------------------------------------------------ reconstruction backward
// d/dpred of w * mean((tanh(pred) - target)^2) (model.py:459-464), bf16 in, bf16 out; only needed when the Kohonen
// maps put the reconstruction loss into the objective (train.py:925-926).
__global__ void __launch_bounds__(256) tanh_mse_bwd_kernel(const __nv_bfloat16* __restrict__ pred, const __nv_bfloat16* __restrict__ tgt,
                                                           long long n, float two_inv_count, const float* __restrict__ w,
                                                           __nv_bfloat16* __restrict__ dpred) {
  pdl_enter();
  const float k = two_inv_count * w[0];
  const long long n8 = n >> 3;
  const long long stride = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 p = ldg_u4_stream(pred + i * 8), t = ldg_u4_stream(tgt + i * 8);
    const uint32_t pa[4] = {p.x, p.y, p.z, p.w}, ta[4] = {t.x, t.y, t.z, t.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float y0 = tanhf(bf16lo(pa[e])), y1 = tanhf(bf16hi(pa[e]));
      o[e] = pack_bf16(k * (y0 - bf16lo(ta[e])) * (1.f - y0 * y0), k * (y1 - bf16hi(ta[e])) * (1.f - y1 * y1));
    }
    *reinterpret_cast<uint4*>(dpred + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  for (long long i = n8 * 8 + 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float y = tanhf(__bfloat162float(pred[i]));
    dpred[i] = __float2bfloat16(k * (y - __bfloat162float(tgt[i])) * (1.f - y * y));
  }
}

}  // namespace nvit

using namespace nvit;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int nvit_split_bf16(const float* x, void* hi_or_null, void* lo, int64_t n, void* stream) {
  NVIT_REQUIRE(x && lo && n >= 0, "nvit_split_bf16: bad arguments");
  NVIT_REQUIRE(((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(lo) & 7) | (reinterpret_cast<uintptr_t>(hi_or_null) & 7)) == 0,
               "nvit_split_bf16: buffers must be 16-byte (fp32) / 8-byte (bf16) aligned");
  if (n == 0) return NVIT_OK;
  launch(split_bf16_kernel, som_grid(n / 4 + 1, 256), 256, 0, ST(stream), x, static_cast<__nv_bfloat16*>(hi_or_null), static_cast<__nv_bfloat16*>(lo), n);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_som_prepare(const float* nodes, int64_t G, int64_t C, float* node_sq, void* hi, void* lo, float* snapshot, void* stream) {
  NVIT_REQUIRE(nodes && node_sq && hi && lo && snapshot && G > 0 && C > 0, "nvit_som_prepare: bad arguments");
  launch(som_prepare_kernel, (int)G, 128, 0, ST(stream), nodes, (int)C, node_sq, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), snapshot);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_som_select(const float* dots, const float* node_sq, const float* nodes, int64_t M, int64_t G, int64_t C, int32_t* idx,
                               int64_t* idx64_or_null, void* onehot_or_null, float* counts_or_null, float* repr32, void* repr16, void* stream) {
  NVIT_REQUIRE(dots && node_sq && nodes && idx && repr32 && repr16, "nvit_som_select: null argument");
  NVIT_REQUIRE(M >= 0 && G > 0 && G <= 8192 && C > 0, "nvit_som_select: bad sizes M=%lld G=%lld C=%lld", (long long)M, (long long)G, (long long)C);
  if (M == 0) return NVIT_OK;
  launch(som_select_kernel, som_grid(M * 32, 256, 4), 256, (size_t)G * sizeof(float), ST(stream), 
      dots, node_sq, nodes, M, (int)G, (int)C, idx, reinterpret_cast<long long*>(idx64_or_null), static_cast<__nv_bfloat16*>(onehot_or_null),
      counts_or_null, repr32, static_cast<__nv_bfloat16*>(repr16));
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_som_pool(const float* x, int64_t rows, int64_t run, float* out, void* stream) {
  NVIT_REQUIRE(x && out && rows >= 0 && run > 0 && run < (1ll << 31), "nvit_som_pool: bad arguments");
  if (rows == 0) return NVIT_OK;
  launch(som_pool_kernel, som_grid(rows * 32, 256), 256, 0, ST(stream), x, rows, (int)run, out);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_som_update(float* nodes, const float* pooled, const int32_t* bmu, int64_t steps, int64_t grid_rows, int64_t grid_cols,
                               int64_t C, const float* coef_dev, float sigma, void* stream) {
  NVIT_REQUIRE(nodes && pooled && bmu && coef_dev, "nvit_som_update: null argument");
  NVIT_REQUIRE(steps >= 0 && grid_rows > 0 && grid_cols > 0 && C > 0 && sigma > 0.f, "nvit_som_update: bad sizes");
  if (steps == 0) return NVIT_OK;
  launch(som_update_kernel, (int)(grid_rows * grid_cols), 256, 0, ST(stream), nodes, pooled, bmu, (int)steps, (int)C, (int)grid_rows, (int)grid_cols,
                                                                           coef_dev, 2.f * sigma * sigma);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_som_pair_losses(const float* repr_l, const float* repr_g, const float* x_l, const float* x_g, int64_t M, int64_t C,
                                    float* sums3_or_null, const float* weights3_or_null, float* d_repr_l, float* d_repr_g, float* d_x_l,
                                    float* d_x_g, void* stream) {
  NVIT_REQUIRE(repr_l && repr_g && x_l && x_g && M >= 0 && C > 0, "nvit_som_pair_losses: bad arguments");
  NVIT_REQUIRE((C % 4) == 0 && C <= 1024, "nvit_som_pair_losses: C must be a multiple of 4 and <= 1024 (got %lld)", (long long)C);
  NVIT_REQUIRE(!weights3_or_null || (d_repr_l && d_repr_g && d_x_l && d_x_g), "nvit_som_pair_losses: gradients requested without outputs");
  if (M == 0) return NVIT_OK;
  const int grid = som_grid(M * 32, 256, 4);
  const int nv = (int)((C + 127) / 128);
#define LAUNCH(NV) launch(som_pair_kernel<NV>, grid, 256, 0, ST(stream), repr_l, repr_g, x_l, x_g, M, (int)C, sums3_or_null, weights3_or_null, d_repr_l, d_repr_g, d_x_l, d_x_g)
  if (nv <= 1) LAUNCH(1);
  else if (nv <= 2) LAUNCH(2);
  else if (nv <= 4) LAUNCH(4);
  else if (nv <= 6) LAUNCH(6);
  else LAUNCH(8);
#undef LAUNCH
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_som_smoothness(const float* nodes, const float* counts, int64_t side, int64_t C, int64_t M, float* loss_accum,
                                   const float* weight_or_null, float* gnodes_or_null, void* stream) {
  NVIT_REQUIRE(nodes && counts && loss_accum && side > 0 && C > 0 && M > 0, "nvit_som_smoothness: bad arguments");
  launch(som_smooth_kernel, (int)(side * side * 8), 128, 0, ST(stream), nodes, counts, (int)side, (int)C, 1.f / (8.f * (float)M), loss_accum,
                                                                   weight_or_null, gnodes_or_null);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_tanh_mse_bwd(const void* pred, const void* target, int64_t n, float inv_count, const float* weight_dev, void* dpred,
                                 void* stream) {
  NVIT_REQUIRE(pred && target && weight_dev && dpred && n >= 0, "nvit_tanh_mse_bwd: bad arguments");
  if (n == 0) return NVIT_OK;
  launch(tanh_mse_bwd_kernel, som_grid(n / 8 + 1, 256, 4), 256, 0, ST(stream), static_cast<const __nv_bfloat16*>(pred), static_cast<const __nv_bfloat16*>(target),
                                                                          n, 2.f * inv_count, weight_dev, static_cast<__nv_bfloat16*>(dpred));
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}
