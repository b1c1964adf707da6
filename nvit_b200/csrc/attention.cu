// Unit-norm QK attention for nViT on sm_100a tensor cores (tcgen05 + TMEM), forward and backward.
//
//   qh = s * q/||q||_D,  kh = s * k/||k||_D,  s = sqk * sqk_mul            (nvit/model.py:108-119, 236-249)
//   O  = softmax(scale * qh kh^T) v,  non-causal                           (nvit/model.py:121-127, 252-258)
//
// The sequence is short and fixed by the patch grid (T = 196 for 224/16, 64 for the CIFAR config), so one CTA owns one
// (batch, head): Q, K, V (and dO) tiles of up to 256 tokens x 64 channels are TMA-loaded into 128B-swizzled shared
// memory, the q/k row normalisation and sqk scaling run in place on those tiles (the reference spends ~8 eager kernels
// and fp32 temporaries on it), and every matmul is a tcgen05.mma whose operands are just different descriptor views
// (K-major or MN-major) of the same shared tiles:
//   forward : S = Qh Kh^T (M=128 q rows, N=Tpad)  -> row softmax by the thread owning the TMEM lane -> P (bf16, smem)
//             O = P V      (B operand = V as it lies, MN-major)
//   backward (per 128-row kv tile j, scores kept transposed so that dV and dK complete per tile):
//             S^T = Kh_j Qh^T ; P^T = exp(scale S^T - lse) ; dV_j = P^T dO ; dP^T = V_j dO^T ;
//             dS^T = P^T (dP^T - delta) scale ; dK_j = dS^T Qh ; dQ += dS Kh_j (A = dS^T viewed MN-major)
//             then the backward of the row normalisation and the sqk gradient.
#include "common.cuh"
#include <string.h>

namespace nvit {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box);

constexpr int ATT_ROWS = 256;                 // token capacity of a shared tile
constexpr int ATT_TILE_BYTES = ATT_ROWS * 128;  // [256 tokens][64 bf16]
constexpr int ATT_PB_BYTES = 128 * 256 * 2;     // [128 rows][4 k-blocks x 64 bf16]
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a 128B-swizzled tile with 128-byte rows
__device__ __forceinline__ uint32_t sw128(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// In-place q/k row normalisation on a swizzled tile: row <- s * row / ||row||.  Returns 1/||row|| (0 for a zero row).
__device__ __forceinline__ float normalize_row(uint8_t* tile, int row, const float* s_scale) {
  float x[64];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(tile + sw128(row, c));
    unpack8(u, x + 8 * c);
  }
#pragma unroll
  for (int i = 0; i < 64; ++i) ss += x[i] * x[i];
  const float inv = ss > 0.f ? 1.f / sqrtf(ss) : 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float y[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) y[e] = x[8 * c + e] * inv * s_scale[8 * c + e];
    *reinterpret_cast<uint4*>(tile + sw128(row, c)) = pack8(y);
  }
  return inv;
}

struct alignas(64) AttnParams {
  CUtensorMap tq, tk, tv, tdo;
  const __nv_bfloat16 *q, *k, *o, *dout;  // raw rows for the normalisation backward / delta
  __nv_bfloat16 *out, *dq, *dk, *dv;
  const float* sqk;
  float* lse;
  float* dsqk;
  long long ldq, ldk, ldo, lddq, lddk, lddv;
  float sqk_mul, scale;
  int B, H, T, TP, nQ, nK;
};

__host__ __device__ constexpr uint32_t IDESC_KK(int N) { return umma_idesc_bf16(128, N, 0, 0); }  // A K-major, B K-major
__host__ __device__ constexpr uint32_t IDESC_KM(int N) { return umma_idesc_bf16(128, N, 0, 1); }  // A K-major, B MN-major
__host__ __device__ constexpr uint32_t IDESC_MM(int N) { return umma_idesc_bf16(128, N, 1, 1); }  // A MN-major, B MN-major
__device__ __forceinline__ uint32_t idesc_kk_n(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int ATT_FWD_SMEM = 3 * ATT_TILE_BYTES + ATT_PB_BYTES + 1024 /*align*/ + 512 /*scale + barriers*/;

__global__ void __launch_bounds__(128, 1) attn_fwd_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_TILE_BYTES;
  uint8_t* sP = sV + ATT_TILE_BYTES;
  float* s_scale = reinterpret_cast<float*>(sP + ATT_PB_BYTES);  // [64]
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(s_scale + 64);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, TP = p.TP;

  if (tid == 0) {
    tma_prefetch_desc(&p.tq);
    tma_prefetch_desc(&p.tk);
    tma_prefetch_desc(&p.tv);
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (tid < 64) s_scale[tid] = p.sqk ? p.sqk[h * 64 + tid] * p.sqk_mul : 1.f;
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_tma, 3 * ATT_TILE_BYTES);
    tma_load_3d(&p.tq, bar_tma, sQ, h * 64, 0, b);
    tma_load_3d(&p.tk, bar_tma, sK, h * 64, 0, b);
    tma_load_3d(&p.tv, bar_tma, sV, h * 64, 0, b);
  }
  mbar_wait(bar_tma, 0);

  if (p.sqk) {
    for (int r = tid; r < T; r += 128) {
      normalize_row(sQ, r, s_scale);
      normalize_row(sK, r, s_scale);
    }
  }
  fence_proxy_async_smem();
  __syncthreads();

  const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV), sP_a = smem_u32(sP);
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const float sl2 = p.scale * LOG2E;
  uint32_t mma_phase = 0;

  for (int i = 0; i < p.nQ; ++i) {
    if (tid == 0) {
      tc_fence_after_sync();
      const uint32_t idesc = idesc_kk_n(TP);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16_ss(tmem_base, umma_smem_desc(sQ_a + i * 16384 + ks * 32, 16, 1024), umma_smem_desc(sK_a + ks * 32, 16, 1024),
                     idesc, ks > 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after_sync();

    const int qtok = i * 128 + tid;
    float mx = -INFINITY;
    for (int c0 = 0; c0 < TP; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_lane + c0, r);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (c0 + e < T) mx = fmaxf(mx, __uint_as_float(r[e]));
    }
    float sum = 0.f;
    for (int c0 = 0; c0 < TP; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_lane + c0, r);
      tmem_wait_ld();
      float pv[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float v = (c0 + e < T) ? exp2f((__uint_as_float(r[e]) - mx) * sl2) : 0.f;
        pv[e] = v;
        sum += v;
      }
      uint8_t* blk = sP + (c0 >> 6) * 16384;
      const int ch = (c0 & 63) >> 3;
      *reinterpret_cast<uint4*>(blk + sw128(tid, ch)) = pack8(pv);
      *reinterpret_cast<uint4*>(blk + sw128(tid, ch + 1)) = pack8(pv + 8);
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after_sync();
      constexpr uint32_t idesc = IDESC_KM(64);
      const int nks = TP >> 4;
      for (int ks = 0; ks < nks; ++ks)
        umma_bf16_ss(tmem_base + 256, umma_smem_desc(sP_a + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                     umma_smem_desc(sV_a + ks * 2048, 8192, 1024), idesc, ks > 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after_sync();

    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(t_lane + 256, r0);
      tmem_ld_32x32b_x32(t_lane + 256 + 32, r1);
      tmem_wait_ld();
      if (qtok < T) {
        const float inv = 1.f / sum;
        __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * T + qtok) * p.ldo + h * 64;
        float o[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(r0[8 * c + e]) * inv;
          *reinterpret_cast<uint4*>(dst + 8 * c) = pack8(o);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(r1[8 * c + e]) * inv;
          *reinterpret_cast<uint4*>(dst + 32 + 8 * c) = pack8(o);
        }
        p.lse[(static_cast<long long>(b) * p.H + h) * T + qtok] = mx * p.scale + logf(sum);
      }
    }
    tc_fence_before_sync();
    __syncthreads();
  }

  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int ATT_BWD_SMEM = 4 * ATT_TILE_BYTES + ATT_PB_BYTES + 1024 /*align*/ + (4 * 256 + 128) * 4 + 64;

// backward of y = s * x/||x|| for one row: g = dL/dy (fp32), x raw (bf16 row in global).  Writes dx, accumulates dsqk.
__device__ __forceinline__ void norm_bwd_row(const float* g, const __nv_bfloat16* xrow, float inv, const float* s_scale, bool has_norm,
                                             __nv_bfloat16* dst, float* dacc) {
  if (!has_norm) {
#pragma unroll
    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(dst + 8 * c) = pack8(g + 8 * c);
    return;
  }
  float n[64];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(xrow + 8 * c);
    unpack8(u, n + 8 * c);
  }
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    n[i] *= inv;
    dacc[i] += g[i] * n[i];
    dot += g[i] * s_scale[i] * n[i];
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float d[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = 8 * c + e;
      d[e] = (g[i] * s_scale[i] - n[i] * dot) * inv;
    }
    *reinterpret_cast<uint4*>(dst + 8 * c) = pack8(d);
  }
}

__global__ void __launch_bounds__(128, 1) attn_bwd_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_TILE_BYTES;
  uint8_t* sDO = sV + ATT_TILE_BYTES;
  uint8_t* sP = sDO + ATT_TILE_BYTES;
  float* s_lse = reinterpret_cast<float*>(sP + ATT_PB_BYTES);  // [256]  (already times log2e)
  float* s_delta = s_lse + 256;                                // [256]
  float* s_invq = s_delta + 256;                               // [256]
  float* s_invk = s_invq + 256;                                // [256]
  float* s_scale = s_invk + 256;                               // [64]
  float* s_dsqk = s_scale + 64;                                // [64]
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(s_dsqk + 64);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, TP = p.TP;
  const bool has_norm = p.sqk != nullptr;

  if (tid == 0) {
    tma_prefetch_desc(&p.tq);
    tma_prefetch_desc(&p.tk);
    tma_prefetch_desc(&p.tv);
    tma_prefetch_desc(&p.tdo);
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (tid < 64) {
    s_scale[tid] = has_norm ? p.sqk[h * 64 + tid] * p.sqk_mul : 1.f;
    s_dsqk[tid] = 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_tma, 4 * ATT_TILE_BYTES);
    tma_load_3d(&p.tq, bar_tma, sQ, h * 64, 0, b);
    tma_load_3d(&p.tk, bar_tma, sK, h * 64, 0, b);
    tma_load_3d(&p.tv, bar_tma, sV, h * 64, 0, b);
    tma_load_3d(&p.tdo, bar_tma, sDO, h * 64, 0, b);
  }
  // while the tiles fly: lse, delta = rowsum(dO * O), and a clean P buffer
  for (int r = tid; r < 256; r += 128) {
    float l = 0.f, d = 0.f;
    if (r < T) {
      l = p.lse[(static_cast<long long>(b) * p.H + h) * T + r] * LOG2E;
      const __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * T + r) * p.ldo + h * 64;
      const __nv_bfloat16* grow = p.dout + (static_cast<long long>(b) * T + r) * p.ldo + h * 64;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float a[8], g[8];
        unpack8(*reinterpret_cast<const uint4*>(orow + 8 * c), a);
        unpack8(*reinterpret_cast<const uint4*>(grow + 8 * c), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) d += a[e] * g[e];
      }
    }
    s_lse[r] = l;
    s_delta[r] = d;
    s_invq[r] = 0.f;
    s_invk[r] = 0.f;
  }
  for (int i = tid; i < ATT_PB_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
  mbar_wait(bar_tma, 0);
  __syncthreads();

  if (has_norm) {
    for (int r = tid; r < T; r += 128) {
      s_invq[r] = normalize_row(sQ, r, s_scale);
      s_invk[r] = normalize_row(sK, r, s_scale);
    }
  }
  fence_proxy_async_smem();
  __syncthreads();

  const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV), sDO_a = smem_u32(sDO), sP_a = smem_u32(sP);
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const float sl2 = p.scale * LOG2E;
  const int nks_q = TP >> 4;  // k-steps over the q axis
  uint32_t mma_phase = 0;
  float dacc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) dacc[i] = 0.f;

  constexpr uint32_t TM_S = 0, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

  for (int j = 0; j < p.nK; ++j) {
    const int kv = j * 128 + tid;
    // ---- S^T_j = Kh_j Qh^T
    if (tid == 0) {
      tc_fence_after_sync();
      const uint32_t idesc = idesc_kk_n(TP);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16_ss(tmem_base + TM_S, umma_smem_desc(sK_a + j * 16384 + ks * 32, 16, 1024),
                     umma_smem_desc(sQ_a + ks * 32, 16, 1024), idesc, ks > 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after_sync();
    // ---- P^T = exp(scale S^T - lse)   (thread = kv row, columns = q)
    for (int c0 = 0; c0 < TP; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_lane + TM_S + c0, r);
      tmem_wait_ld();
      float pv[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int qi = c0 + e;
        pv[e] = (kv < T && qi < T) ? exp2f(__uint_as_float(r[e]) * sl2 - s_lse[qi]) : 0.f;
      }
      uint8_t* blk = sP + (c0 >> 6) * 16384;
      const int ch = (c0 & 63) >> 3;
      *reinterpret_cast<uint4*>(blk + sw128(tid, ch)) = pack8(pv);
      *reinterpret_cast<uint4*>(blk + sw128(tid, ch + 1)) = pack8(pv + 8);
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();
    __syncthreads();
    // ---- dV_j = P^T dO ; dP^T_j = V_j dO^T
    if (tid == 0) {
      tc_fence_after_sync();
      for (int ks = 0; ks < nks_q; ++ks)
        umma_bf16_ss(tmem_base + TM_DV, umma_smem_desc(sP_a + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                     umma_smem_desc(sDO_a + ks * 2048, 8192, 1024), IDESC_KM(64), ks > 0);
      const uint32_t idesc = idesc_kk_n(TP);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16_ss(tmem_base + TM_S, umma_smem_desc(sV_a + j * 16384 + ks * 32, 16, 1024),
                     umma_smem_desc(sDO_a + ks * 32, 16, 1024), idesc, ks > 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after_sync();
    // ---- dS^T = P^T (dP^T - delta) scale, in place over P^T
    for (int c0 = 0; c0 < TP; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_lane + TM_S + c0, r);
      tmem_wait_ld();
      uint8_t* blk = sP + (c0 >> 6) * 16384;
      const int ch = (c0 & 63) >> 3;
      float pv[16];
      unpack8(*reinterpret_cast<const uint4*>(blk + sw128(tid, ch)), pv);
      unpack8(*reinterpret_cast<const uint4*>(blk + sw128(tid, ch + 1)), pv + 8);
#pragma unroll
      for (int e = 0; e < 16; ++e) pv[e] = pv[e] * (__uint_as_float(r[e]) - s_delta[c0 + e]) * p.scale;
      *reinterpret_cast<uint4*>(blk + sw128(tid, ch)) = pack8(pv);
      *reinterpret_cast<uint4*>(blk + sw128(tid, ch + 1)) = pack8(pv + 8);
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();
    __syncthreads();
    // ---- dK_j = dS^T Qh ; dQ_m += dS_j Kh_j
    if (tid == 0) {
      tc_fence_after_sync();
      for (int ks = 0; ks < nks_q; ++ks)
        umma_bf16_ss(tmem_base + TM_DK, umma_smem_desc(sP_a + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                     umma_smem_desc(sQ_a + ks * 2048, 8192, 1024), IDESC_KM(64), ks > 0);
      const int kv_steps = min(8, (TP - j * 128) >> 4);
      for (int m = 0; m < p.nQ; ++m)
        for (int ks = 0; ks < kv_steps; ++ks)
          umma_bf16_ss(tmem_base + TM_DQ + 64 * m, umma_smem_desc(sP_a + 2 * m * 16384 + ks * 2048, 16384, 1024),
                       umma_smem_desc(sK_a + j * 16384 + ks * 2048, 8192, 1024), IDESC_MM(64), (j > 0 || ks > 0));
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after_sync();
    // ---- write dV_j, dK_j rows
    {
      uint32_t r0[32], r1[32];
      float g[64];
      tmem_ld_32x32b_x32(t_lane + TM_DV, r0);
      tmem_ld_32x32b_x32(t_lane + TM_DV + 32, r1);
      tmem_wait_ld();
      if (kv < T) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { g[i] = __uint_as_float(r0[i]); g[32 + i] = __uint_as_float(r1[i]); }
        __nv_bfloat16* dst = p.dv + (static_cast<long long>(b) * T + kv) * p.lddv + h * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(dst + 8 * c) = pack8(g + 8 * c);
      }
      tmem_ld_32x32b_x32(t_lane + TM_DK, r0);
      tmem_ld_32x32b_x32(t_lane + TM_DK + 32, r1);
      tmem_wait_ld();
      if (kv < T) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { g[i] = __uint_as_float(r0[i]); g[32 + i] = __uint_as_float(r1[i]); }
        norm_bwd_row(g, p.k + (static_cast<long long>(b) * T + kv) * p.ldk + h * 64, s_invk[kv], s_scale, has_norm,
                     p.dk + (static_cast<long long>(b) * T + kv) * p.lddk + h * 64, dacc);
      }
    }
    tc_fence_before_sync();
    __syncthreads();
  }

  // ---- dQ rows
  for (int m = 0; m < p.nQ; ++m) {
    const int qi = m * 128 + tid;
    uint32_t r0[32], r1[32];
    tmem_ld_32x32b_x32(t_lane + TM_DQ + 64 * m, r0);
    tmem_ld_32x32b_x32(t_lane + TM_DQ + 64 * m + 32, r1);
    tmem_wait_ld();
    if (qi < T) {
      float g[64];
#pragma unroll
      for (int i = 0; i < 32; ++i) { g[i] = __uint_as_float(r0[i]); g[32 + i] = __uint_as_float(r1[i]); }
      norm_bwd_row(g, p.q + (static_cast<long long>(b) * T + qi) * p.ldq + h * 64, s_invq[qi], s_scale, has_norm,
                   p.dq + (static_cast<long long>(b) * T + qi) * p.lddq + h * 64, dacc);
    }
  }
  if (has_norm) {
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const float v = warp_sum(dacc[i]);
      if (lane == 0) atomicAdd(&s_dsqk[i], v);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (has_norm && tid < 64) atomicAdd(p.dsqk + h * 64 + tid, s_dsqk[tid] * p.sqk_mul);
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

static int make_head_tmap(CUtensorMap* m, const void* base, long long ld, int B, int H, int T) {
  const uint64_t dims[3] = {static_cast<uint64_t>(H) * 64, static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
  const uint64_t strides[2] = {static_cast<uint64_t>(ld), static_cast<uint64_t>(ld) * T};
  const uint32_t box[3] = {64, ATT_ROWS, 1};
  return make_tmap_bf16(m, base, 3, dims, strides, box);
}

static int attn_check(const char* who, int64_t B, int64_t H, int64_t T, int64_t D) {
  NVIT_REQUIRE(D == 64, "%s: head_dim must be 64 (got %lld)", who, (long long)D);
  NVIT_REQUIRE(T >= 1 && T <= ATT_ROWS, "%s: sequence length %lld outside [1, %d]", who, (long long)T, ATT_ROWS);
  NVIT_REQUIRE(B >= 1 && H >= 1 && B * H < (1ll << 31), "%s: bad batch/heads", who);
  return NVIT_OK;
}

}  // namespace nvit

using namespace nvit;

extern "C" int nvit_attention_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                  const float* sqk, float sqk_mul, float scale, void* out, int64_t ldo, float* lse, int64_t B,
                                  int64_t H, int64_t T, int64_t D, void* stream) {
  NVIT_REQUIRE(q && k && v && out && lse, "nvit_attention_fwd: null argument");
  int rc = attn_check("nvit_attention_fwd", B, H, T, D);
  if (rc) return rc;
  NVIT_REQUIRE((ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "nvit_attention_fwd: out must be 16-byte aligned rows");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  if ((rc = make_head_tmap(&p.tq, q, ldq, (int)B, (int)H, (int)T))) return rc;
  if ((rc = make_head_tmap(&p.tk, k, ldk, (int)B, (int)H, (int)T))) return rc;
  if ((rc = make_head_tmap(&p.tv, v, ldv, (int)B, (int)H, (int)T))) return rc;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.lse = lse;
  p.sqk = sqk;
  p.sqk_mul = sqk_mul;
  p.scale = scale;
  p.B = (int)B; p.H = (int)H; p.T = (int)T;
  p.TP = (int)((T + 15) / 16 * 16);
  p.nQ = (int)((T + 127) / 128);
  p.nK = p.nQ;
  static bool attr_set = false;
  if (!attr_set) {
    NVIT_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_FWD_SMEM));
    attr_set = true;
  }
  attn_fwd_kernel<<<(unsigned)(B * H), 128, ATT_FWD_SMEM, static_cast<cudaStream_t>(stream)>>>(p);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_attention_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                  const float* sqk, float sqk_mul, float scale, const void* out, const void* dout, int64_t ldo,
                                  const float* lse, void* dq, void* dk, void* dv, int64_t lddq, int64_t lddk, int64_t lddv,
                                  float* dsqk_accum, int64_t B, int64_t H, int64_t T, int64_t D, void* stream) {
  NVIT_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv, "nvit_attention_bwd: null argument");
  NVIT_REQUIRE((sqk == nullptr) == (dsqk_accum == nullptr), "nvit_attention_bwd: sqk and dsqk go together");
  int rc = attn_check("nvit_attention_bwd", B, H, T, D);
  if (rc) return rc;
  NVIT_REQUIRE(((ldq | ldk | ldo | lddq | lddk | lddv) % 8) == 0, "nvit_attention_bwd: leading dims must be multiples of 8");
  NVIT_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(out) |
                 reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) |
                 reinterpret_cast<uintptr_t>(dv)) & 15) == 0, "nvit_attention_bwd: buffers must be 16-byte aligned");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  if ((rc = make_head_tmap(&p.tq, q, ldq, (int)B, (int)H, (int)T))) return rc;
  if ((rc = make_head_tmap(&p.tk, k, ldk, (int)B, (int)H, (int)T))) return rc;
  if ((rc = make_head_tmap(&p.tv, v, ldv, (int)B, (int)H, (int)T))) return rc;
  if ((rc = make_head_tmap(&p.tdo, dout, ldo, (int)B, (int)H, (int)T))) return rc;
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.o = static_cast<const __nv_bfloat16*>(out);
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.dq = static_cast<__nv_bfloat16*>(dq);
  p.dk = static_cast<__nv_bfloat16*>(dk);
  p.dv = static_cast<__nv_bfloat16*>(dv);
  p.ldq = ldq; p.ldk = ldk; p.ldo = ldo; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.lse = const_cast<float*>(lse);
  p.sqk = sqk;
  p.dsqk = dsqk_accum;
  p.sqk_mul = sqk_mul;
  p.scale = scale;
  p.B = (int)B; p.H = (int)H; p.T = (int)T;
  p.TP = (int)((T + 15) / 16 * 16);
  p.nQ = (int)((T + 127) / 128);
  p.nK = p.nQ;
  static bool attr_set = false;
  if (!attr_set) {
    NVIT_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWD_SMEM));
    attr_set = true;
  }
  attn_bwd_kernel<<<(unsigned)(B * H), 128, ATT_BWD_SMEM, static_cast<cudaStream_t>(stream)>>>(p);
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}
