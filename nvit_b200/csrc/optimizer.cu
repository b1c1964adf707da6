// One-pass optimizer tail of the nViT training step: gradient clip scale + AdamW + the post-step weight
// normalisation (Trainer.normalize_matrices) + the bf16 GEMM-operand emit + zero_grad, in ONE launch over the flat
// parameter / gradient / moment buffers.  Replaces, per step, torch's clip_grad_norm_ scaling, fused AdamW
// (/root/reference/nvit/model.py:369-385, train.py:935-946), the 72 x 5 launches of normalize_matrices (train.py:461-480),
// autocast's weight casts of the next forward (train.py:905) and optimizer.zero_grad (train.py:946).
//
// Work is cut into units listed in a device table of segments (one per parameter tensor):
//   kind 0  plain        unit = 8192 consecutive elements: AdamW (+ bf16 copy for GEMM weights)
//   kind 1  row-normed   unit = 32 rows, one warp per row: the row's p, g, m, v pass through registers once, the updated
//                        row is L2-normalised in registers and written as fp32 + bf16 (query/key/value/c_fc: norm over n_embd = dim 1)
//   kind 2  col-normed   unit = 128 columns x all rows: pass 1 updates and accumulates the column sums of squares, pass 2
//                        re-reads the (L2-resident) slab, scales it and writes fp32 + bf16 (att_c_proj / mlp_c_proj: dim 0)
// Units are claimed through an atomic counter (a column unit moves ~20x the bytes of a row unit; the table lists them
// first), so the launch has no tail.  Algorithmic traffic per element: 16 B read + 12 B written (p, m, v) + 4 B (zeroed g)
// + 2 B (bf16) = 34 B against 28 + 8 + 6 + 4 = 46 B for the four separate passes.
#include "common.cuh"

namespace nvit {

struct TailHyper {
  float lr, b1, b2, eps, wd, bc1, bc2_sqrt, max_norm;
};

struct TailCoef {
  float clip, step_size, decay, b1, b2, eps, bc2_sqrt;
};

__device__ __forceinline__ float adamw_elem(float p, float g, float& m, float& v, const TailCoef& k, bool decay) {
  const float gr = g * k.clip;
  if (decay) p *= k.decay;
  m = k.b1 * m + (1.f - k.b1) * gr;
  v = k.b2 * v + (1.f - k.b2) * gr * gr;
  const float denom = sqrtf(v) / k.bc2_sqrt + k.eps;
  return p - k.step_size * (m / denom);
}

__device__ __forceinline__ float4 adamw_vec(float4 p, const float4 g, float4& m, float4& v, const TailCoef& k, bool decay) {
  p.x = adamw_elem(p.x, g.x, m.x, v.x, k, decay);
  p.y = adamw_elem(p.y, g.y, m.y, v.y, k, decay);
  p.z = adamw_elem(p.z, g.z, m.z, v.z, k, decay);
  p.w = adamw_elem(p.w, g.w, m.w, v.w, k, decay);
  return p;
}

constexpr int TAIL_FIELDS = 8;      // {offset, rows, cols, kind, w16 offset or -1, decay, first_unit, reserved}
constexpr int PLAIN_UNIT = 8192;    // elements per kind-0 unit (256 threads x 8 float4)
constexpr int ROW_UNIT = 32;        // rows per kind-1 unit (8 warps x 4 rows)

__global__ void __launch_bounds__(256, 2)
adamw_norm_fused_kernel(float* __restrict__ P, float* __restrict__ G, float* __restrict__ Mo, float* __restrict__ Vo,
                        __nv_bfloat16* __restrict__ W16, const long long* __restrict__ table, int n_seg, long long total_units,
                        TailHyper h, const float* __restrict__ gnorm_sq, const float* __restrict__ dev_lr_step,
                        unsigned int* __restrict__ counter, int zero_grad) {
  pdl_enter();
  __shared__ float4 s_part[8][32];
  __shared__ float4 s_inv[32];
  __shared__ long long s_unit[2];
  if (dev_lr_step) {   // CUDA-graph friendly: learning rate and 1-based step count live in device memory
    h.lr = dev_lr_step[0];
    const float t = dev_lr_step[1];
    h.bc1 = 1.f - powf(h.b1, t);
    h.bc2_sqrt = sqrtf(1.f - powf(h.b2, t));
  }
  TailCoef k;
  k.clip = 1.f;
  if (gnorm_sq) k.clip = fminf(1.f, h.max_norm / (sqrtf(gnorm_sq[0]) + 1e-6f));
  k.step_size = h.lr / h.bc1;
  k.decay = 1.f - h.lr * h.wd;
  k.b1 = h.b1; k.b2 = h.b2; k.eps = h.eps; k.bc2_sqrt = h.bc2_sqrt;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // Units are claimed one ahead: the atomic's round trip to L2 (~1 us) runs under the current unit's streaming.
  if (threadIdx.x == 0) s_unit[0] = (long long)atomicAdd(counter, 1u);
  for (int it = 0;; ++it) {
    __syncthreads();                       // s_unit[it & 1] is published; everybody is done with the previous unit
    const long long unit = s_unit[it & 1];
    if (unit >= total_units) break;
    if (threadIdx.x == 0) s_unit[(it + 1) & 1] = (long long)atomicAdd(counter, 1u);
    int lo = 0, hi = n_seg - 1;            // last segment with first_unit <= unit
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (table[mid * TAIL_FIELDS + 6] <= unit) lo = mid; else hi = mid - 1;
    }
    const long long* e = table + lo * TAIL_FIELDS;
    const long long off = e[0];
    const int rows = (int)e[1], cols = (int)e[2], kind = (int)e[3];
    const long long w16_off = e[4];
    const bool decay = e[5] != 0;
    const long long u = unit - e[6];
    float* p = P + off;
    float* g = G + off;
    float* m = Mo + off;
    float* v = Vo + off;
    __nv_bfloat16* w16 = w16_off >= 0 ? W16 + w16_off : nullptr;

    if (kind == 0) {
      // ---------------------------------------------------------------- plain: offsets are multiples of 8 elements
      const long long n = 1ll * rows * cols;
      const long long base = u * PLAIN_UNIT;
#pragma unroll
      for (int half = 0; half < 2; ++half) {       // 2 x (4 float4 of each stream in flight per thread)
        float4 pp[4], gg[4], mm[4], vv[4];
        long long idx[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          idx[q] = base + ((half * 4 + q) * 256 + threadIdx.x) * 4;
          if (idx[q] + 4 <= n) {
            pp[q] = *reinterpret_cast<const float4*>(p + idx[q]);
            gg[q] = *reinterpret_cast<const float4*>(g + idx[q]);
            mm[q] = *reinterpret_cast<const float4*>(m + idx[q]);
            vv[q] = *reinterpret_cast<const float4*>(v + idx[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const long long i = idx[q];
          if (i + 4 <= n) {
            pp[q] = adamw_vec(pp[q], gg[q], mm[q], vv[q], k, decay);
            *reinterpret_cast<float4*>(p + i) = pp[q];
            *reinterpret_cast<float4*>(m + i) = mm[q];
            *reinterpret_cast<float4*>(v + i) = vv[q];
            if (zero_grad) *reinterpret_cast<float4*>(g + i) = zero4;
            if (w16) *reinterpret_cast<uint2*>(w16 + i) = make_uint2(pack_bf16(pp[q].x, pp[q].y), pack_bf16(pp[q].z, pp[q].w));
          } else {
            for (long long t = i; t < n && t < i + 4; ++t) {
              float me = m[t], ve = v[t];
              const float pe = adamw_elem(p[t], g[t], me, ve, k, decay);
              p[t] = pe; m[t] = me; v[t] = ve;
              if (zero_grad) g[t] = 0.f;
              if (w16) w16[t] = __float2bfloat16(pe);
            }
          }
        }
      }
    } else if (kind == 1) {
      // ---------------------------------------------------------------- one warp per row, the row stays in registers
#pragma unroll 1
      for (int rr = 0; rr < ROW_UNIT / 8; ++rr) {
      const int r = (int)u * ROW_UNIT + rr * 8 + wy;
      if (r < rows) {
        const long long ro = 1ll * r * cols;
        if ((cols & 3) == 0 && cols <= 1024) {
          float4 pv[8];
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = (j * 32 + lane) * 4;
            pv[j] = zero4;
            if (c < cols) {
              float4 pp = *reinterpret_cast<const float4*>(p + ro + c);
              const float4 gg = *reinterpret_cast<const float4*>(g + ro + c);
              float4 mm = *reinterpret_cast<const float4*>(m + ro + c);
              float4 vv = *reinterpret_cast<const float4*>(v + ro + c);
              pp = adamw_vec(pp, gg, mm, vv, k, decay);
              *reinterpret_cast<float4*>(m + ro + c) = mm;
              *reinterpret_cast<float4*>(v + ro + c) = vv;
              if (zero_grad) *reinterpret_cast<float4*>(g + ro + c) = zero4;
              pv[j] = pp;
            }
            ss += pv[j].x * pv[j].x + pv[j].y * pv[j].y + pv[j].z * pv[j].z + pv[j].w * pv[j].w;
          }
          ss = warp_sum(ss);
          const float inv = 1.f / sqrtf(ss);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = (j * 32 + lane) * 4;
            if (c < cols) {
              const float4 o = make_float4(pv[j].x * inv, pv[j].y * inv, pv[j].z * inv, pv[j].w * inv);
              *reinterpret_cast<float4*>(p + ro + c) = o;
              if (w16) *reinterpret_cast<uint2*>(w16 + ro + c) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
            }
          }
        } else {
          float ss = 0.f;
          for (int c = lane; c < cols; c += 32) {
            float me = m[ro + c], ve = v[ro + c];
            const float pe = adamw_elem(p[ro + c], g[ro + c], me, ve, k, decay);
            p[ro + c] = pe; m[ro + c] = me; v[ro + c] = ve;
            if (zero_grad) g[ro + c] = 0.f;
            ss += pe * pe;
          }
          ss = warp_sum(ss);
          const float inv = 1.f / sqrtf(ss);
          for (int c = lane; c < cols; c += 32) {      // each lane re-reads what it wrote itself
            const float x = p[ro + c] * inv;
            p[ro + c] = x;
            if (w16) w16[ro + c] = __float2bfloat16(x);
          }
        }
      }
      }
    } else {
      // ---------------------------------------------------------------- 128 columns x all rows, two passes
      const int c = (int)u * 128 + lane * 4;
      float4 ss = zero4;
      const bool vec = (cols & 3) == 0;
      if (vec) {
        if (c < cols) {
#pragma unroll 2
          for (int r = wy; r < rows; r += 8) {
            const long long i = 1ll * r * cols + c;
            float4 pp = *reinterpret_cast<const float4*>(p + i);
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<const float4*>(m + i);
            float4 vv = *reinterpret_cast<const float4*>(v + i);
            pp = adamw_vec(pp, gg, mm, vv, k, decay);
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
            if (zero_grad) *reinterpret_cast<float4*>(g + i) = zero4;
            ss.x += pp.x * pp.x; ss.y += pp.y * pp.y; ss.z += pp.z * pp.z; ss.w += pp.w * pp.w;
          }
        }
      } else {
        float* sp = reinterpret_cast<float*>(&ss);
        for (int q = 0; q < 4; ++q)
          if (c + q < cols)
            for (int r = wy; r < rows; r += 8) {
              const long long i = 1ll * r * cols + c + q;
              float me = m[i], ve = v[i];
              const float pe = adamw_elem(p[i], g[i], me, ve, k, decay);
              p[i] = pe; m[i] = me; v[i] = ve;
              if (zero_grad) g[i] = 0.f;
              sp[q] += pe * pe;
            }
      }
      s_part[wy][lane] = ss;
      __syncthreads();
      if (wy == 0) {
        float4 t = zero4;
#pragma unroll
        for (int q = 0; q < 8; ++q) { const float4 a = s_part[q][lane]; t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w; }
        s_inv[lane] = make_float4(1.f / sqrtf(t.x), 1.f / sqrtf(t.y), 1.f / sqrtf(t.z), 1.f / sqrtf(t.w));
      }
      __syncthreads();
      const float4 inv = s_inv[lane];
      if (vec) {
        if (c < cols) {
#pragma unroll 4
          for (int r = wy; r < rows; r += 8) {       // every thread re-reads exactly what it wrote in pass 1
            const long long i = 1ll * r * cols + c;
            float4 a = *reinterpret_cast<const float4*>(p + i);
            a.x *= inv.x; a.y *= inv.y; a.z *= inv.z; a.w *= inv.w;
            *reinterpret_cast<float4*>(p + i) = a;
            if (w16) *reinterpret_cast<uint2*>(w16 + i) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
          }
        }
      } else {
        const float* ip = reinterpret_cast<const float*>(&inv);
        for (int q = 0; q < 4; ++q)
          if (c + q < cols)
            for (int r = wy; r < rows; r += 8) {
              const long long i = 1ll * r * cols + c + q;
              const float x = p[i] * ip[q];
              p[i] = x;
              if (w16) w16[i] = __float2bfloat16(x);
            }
      }
    }
  }
}

}  // namespace nvit

using namespace nvit;

extern "C" int nvit_adamw_norm_fused(float* p, float* g, float* m, float* v, void* w16_bf16, const int64_t* table_dev,
                                     int64_t n_segments, int64_t total_units, float lr, float beta1, float beta2, float eps,
                                     float weight_decay, int64_t step, const float* gnorm_sq, float max_norm,
                                     const float* dev_lr_step, uint32_t* unit_counter_zeroed, int zero_grad, void* stream) {
  NVIT_REQUIRE(p && g && m && v && table_dev && unit_counter_zeroed && n_segments > 0 && total_units > 0 &&
               (step >= 1 || dev_lr_step), "nvit_adamw_norm_fused: bad arguments");
  NVIT_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(w16_bf16)) & 15) == 0,
               "nvit_adamw_norm_fused: buffers must be 16-byte aligned");
  NVIT_REQUIRE(total_units < (1ll << 31), "nvit_adamw_norm_fused: too many units");
  if (step < 1) step = 1;
  TailHyper h;
  h.lr = lr; h.b1 = beta1; h.b2 = beta2; h.eps = eps; h.wd = weight_decay; h.max_norm = max_norm;
  h.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  h.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  const long long cap = 2ll * nvit_num_sms();
  const int grid = (int)(total_units < cap ? total_units : cap);
  launch(adamw_norm_fused_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), p, g, m, v,
         static_cast<__nv_bfloat16*>(w16_bf16), reinterpret_cast<const long long*>(table_dev), (int)n_segments,
         (long long)total_units, h, gnorm_sq, dev_lr_step, unit_counter_zeroed, zero_grad);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}
