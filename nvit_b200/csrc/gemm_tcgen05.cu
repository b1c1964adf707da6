// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> shared (128B swizzle) -> tcgen05.mma -> TMEM -> epilogue.
//
//   C[M,N] (+)= A[M,K] * B[N,K]^T        fp32 accumulate in tensor memory
//
// Operand storage (both bf16, row-major in global memory):
//   a_mn = 0 : A is stored [M, K]  (K contiguous, "K-major")        -> forward / dgrad
//   a_mn = 1 : A is stored [K, M]  (M contiguous, "MN-major")       -> wgrad (dW = dY^T X reads dY and X as they lie)
//   b_mn likewise for B ([N, K] or [K, N]).
// This replaces what the reference gets from nn.Linear/nn.Conv2d -> cuBLAS/cuDNN (nvit/model.py:99-101,130,148,155,
// 226-228,259,262,286-304,329-332,341-344) and their autograd backward.
//
// CTA = 320 threads: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-9 epilogue (two groups of four).
// Tile 128 x BN x 64, BN in {128, 256}; 4 (BN=256) or 6 (BN=128) smem stages; two TMEM accumulator stages so
// the epilogue of tile i overlaps the main loop of tile i+1.  One CTA per SM, static round-robin tile schedule
// with n fastest so CTAs running together share the same A row-panel in L2.  Optional split along K (wgrad: K = B*T)
// with fp32 red.add into the output.
//
// SWIGLU variant (model.py:148-154): B is the c_fc weight [2F, K]; each CTA tile takes 128 "u" rows j..j+127 and the
// matching 128 "v" rows F+j.., so the accumulator holds u in TMEM columns [0,128) and v in [128,256) and the epilogue
// emits x = (u*su) * silu(v*sv) directly (plus, optionally, the raw bf16 u|v for the backward pass).
#include "common.cuh"
#include <string.h>
#include <mutex>
#include <unordered_map>

namespace nvit {

// nvit_gemm_debug (WRONG outputs, measurement only) exists in -DNVIT_BENCH_HOOKS builds; the product library compiles the
// debug branches away.
#ifdef NVIT_BENCH_HOOKS
#define NVIT_DBG(p) ((p).dbg)
#else
#define NVIT_DBG(p) 0
#endif

struct alignas(64) GemmParams {
  CUtensorMap tma_a;
  CUtensorMap tma_b;
  CUtensorMap tma_c;   // output tile store (bf16 / swiglu x: box 64x128 SW128; fp32: box 32x128 SW128)
  CUtensorMap tma_c2;  // bf16 side copy of an fp32 output (box 32x128 dense) / swiglu raw u (box 64x128 SW128)
  CUtensorMap tma_c3;  // swiglu raw v
  int direct;          // 1: per-thread global stores (outputs that TMA cannot address), 0: smem-staged TMA stores
  int dbg;             // benchmarking only (nvit_gemm_debug): 1 = epilogue releases TMEM without reading it, 2 = reads but does not store
  void* C;
  __nv_bfloat16* C2;
  const float* bias;      // [N] or null
  const float* colscale;  // [N] ([2F] for swiglu) or null
  const float* rowadd;    // [rowadd_period, N] or null
  const __nv_bfloat16* gate_uv;   // GATEB: raw u|v of the forward pass, [M, 2F] bf16 (F = swiglu_half)
  long long ld_uv;
  // unit-norm q/k epilogue (nvit_gemm_qknorm): output columns [0, qk_cols) are normalised per 64-column head and scaled
  // by colscale[col % qk_period] * colscale_mul; 1/||x|| goes to qk_inv[row * ld_inv + col / 64]
  int qk_cols, qk_period;
  float* qk_inv;
  long long ld_inv;
  long long ldc, ldc2;
  int M, N, K;
  int out_f32, accumulate, atomic;
  int rowadd_period;
  int swiglu_half;
  float colscale_mul;
  int tiles_m, tiles_n, splits, kb_total, kb_per_split;
  int group_n;         // tile order: 0 = n fastest over the whole width, G > 0 = bands of G tiles along n (tile_coords)
  // x / tiles_n and x / splits as umulhi(x, magic) (magic = 2^32 / d rounded up; exact while x * d < 2^32; 0 = divide).
  // A runtime integer division is ~40 dependent instructions (I2F, MUFU.RCP, F2I, fix-ups): the roles of this kernel
  // did up to ten of them per tile, the epilogue warps of the gate-backward GEMM on their critical path.
  unsigned magic_tiles_n, magic_splits;
};

__device__ __forceinline__ int fast_div(int x, int d, unsigned magic) {
  if (magic) return static_cast<int>(__umulhi(static_cast<unsigned>(x), magic));
  return d == 1 ? x : x / d;
}

// Tile index -> (m tile, n tile).  The persistent CTAs (or pairs) take consecutive indices, so the order decides which
// operand panels the ~74 / 148 tiles in flight share.  n fastest: the tiles in flight span few A row-panels and the whole
// width; with tiles_n = 24 (the gate GEMM) that is 3 + 24 distinct panels.  Bands of G tiles along n (inside a band n
// fastest, then m) make the set in flight ~74/G x G: 9 + 8 panels for G = 8, i.e. fewer distinct bytes pulled through L2.
__device__ __forceinline__ void tile_coords(const GemmParams& p, int t, int& m_tile, int& n_blk) {
  if (p.group_n <= 0) {
    m_tile = fast_div(t, p.tiles_n, p.magic_tiles_n);
    n_blk = t - m_tile * p.tiles_n;
    return;
  }
  const int band_tiles = p.tiles_m * p.group_n;
  const int full_bands = p.tiles_n / p.group_n;
  const int band = t / band_tiles;
  if (band < full_bands) {
    const int r = t - band * band_tiles;
    m_tile = r / p.group_n;
    n_blk = band * p.group_n + r % p.group_n;
  } else {                                     // the last, narrower band
    const int gw = p.tiles_n - full_bands * p.group_n;
    const int r = t - full_bands * band_tiles;
    m_tile = r / gw;
    n_blk = full_bands * p.group_n + r % gw;
  }
}

template <int BN, bool A_MN, bool B_MN, bool CG2 = false, bool GATEB = false, bool SWIGLU = false, int NG = 2>
struct GemmTraits {
  static constexpr int GROUPS = NG;            // epilogue groups of four warps (CTA = 64 + 128 * NG threads)
  static constexpr int BM = 128;               // rows per CTA (a CTA pair covers 256)
  static constexpr int BK = 64;
  static constexpr int UMMA_K = 16;
  static constexpr int BN_CTA = CG2 ? BN / 2 : BN;   // B rows this CTA stages (a pair splits B between its two CTAs)
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN_CTA * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
#ifndef NVIT_GEMM_NBUF
#define NVIT_GEMM_NBUF 1
#endif
  // epilogue staging: NBUF 16 KB buffers per epilogue group (two groups); what is left of the 224 KB goes to the TMA ring
#ifndef NVIT_GATEB_SETS
#define NVIT_GATEB_SETS 2
#endif
  // GATEB: NVIT_GATEB_SETS sets of {du, dv} buffers per group, so a chunk's stores drain while the next chunk is computed
  static constexpr int GATE_SETS = (GATEB && CG2 && NG == 2) ? NVIT_GATEB_SETS : 1;
  // The pair-mode gate GEMM sends three tiles (x, raw u, raw v) per group and tile: three buffers let them drain side
  // by side (MEASURED: 436 -> 421 -> 417 us with 1 / 2 / 3 buffers; no effect on the one-output kernels).
  static constexpr int NBUF = GATEB ? 2 * GATE_SETS : ((SWIGLU && CG2) ? 3 : ((CG2 || BN == 128) ? NVIT_GEMM_NBUF : 1));
  static constexpr int STAGING_BYTES = NG * NBUF * 16384;
  static constexpr int STAGES = (229376 - STAGING_BYTES) / STAGE_BYTES;
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512 or 256: powers of two
  static constexpr int VEC_BYTES = 2048;       // per-tile scale[256] and bias[256]
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + VEC_BYTES + 256 /*barriers*/;
  // K-major SW128: 8 rows x 128 B per swizzle atom, atoms stacked along M/N every 1024 B.
  // MN-major SW128: 64 elements (128 B) along M/N x 8 k-rows per atom; next 8 k-rows +1024 B; next 64 M/N elements
  // is a separate TMA box of BK rows -> +BK*128 B.
  static constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16;
  static constexpr uint32_t B_LBO = B_MN ? BK * 128 : 16;
  static constexpr uint32_t SBO = 1024;
  static constexpr uint32_t A_KSTEP = A_MN ? UMMA_K * 128 : UMMA_K * 2;  // bytes per UMMA_K step
  static constexpr uint32_t B_KSTEP = B_MN ? UMMA_K * 128 : UMMA_K * 2;
};

__device__ __forceinline__ float tanh_approx_(float x) {
  float y;
  asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));   // volatile: keep ptxas from recomputing it per use
  return y;
}
__device__ __forceinline__ float silu_mul(float u, float v) { return __fdividef(u * v, 1.f + __expf(-v)); }
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16(x)); }

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float* v, int valid, bool vec) {
  if (valid >= 32 && vec) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      uint4 o = make_uint4(pack_bf16(v[i], v[i + 1]), pack_bf16(v[i + 2], v[i + 3]), pack_bf16(v[i + 4], v[i + 5]),
                           pack_bf16(v[i + 6], v[i + 7]));
      *reinterpret_cast<uint4*>(dst + i) = o;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < valid) dst[i] = __float2bfloat16(v[i]);
  }
}

template <int BN, bool A_MN, bool B_MN, bool SWIGLU, bool CG2, bool GATEB = false, int NG = 2>
__global__ void __launch_bounds__(64 + 128 * NG, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using T = GemmTraits<BN, A_MN, B_MN, CG2, GATEB, SWIGLU, NG>;
  static_assert(NG == 2 || GATEB, "four epilogue groups exist for the gate-backward epilogue only");
  static_assert(!GATEB || (BN == 256 && !SWIGLU), "gate-backward epilogue: 128x256 tiles");
  // CG2: the kernel runs as clusters of two CTAs (one SM pair); the pair computes a 256 x BN tile with
  // tcgen05.mma.cta_group::2 issued by the rank-0 CTA.  Each CTA stages its own 128 rows of A and half of B.
  const uint32_t cta_rank = CG2 ? cluster_ctarank() : 0;
  const int unit0 = CG2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int unit_stride = CG2 ? (int)cluster_count_x() : (int)gridDim.x;
  static_assert(!SWIGLU || (BN == 256 && !A_MN && !B_MN), "swiglu epilogue: 128x256 K-major tiles only");
  extern __shared__ __align__(1024) uint8_t smem[];   // the whole 227 KB is used: no slack for manual alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();         // 128B-swizzled tiles need a 1024-byte aligned base
  uint8_t* stg = smem + T::STAGES * T::STAGE_BYTES;   // epilogue staging
  float* s_vec = reinterpret_cast<float*>(stg + T::STAGING_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + T::STAGING_BYTES + T::VEC_BYTES);
  uint64_t* empty_bar = full_bar + T::STAGES;
  uint64_t* tmem_full = empty_bar + T::STAGES;
  uint64_t* tmem_empty = tmem_full + T::ACC_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + T::ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = p.tiles_m * p.tiles_n * p.splits;
  constexpr int TILE_N = SWIGLU ? 128 : BN;  // output columns covered by one tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
    if (!p.direct) tma_prefetch_desc(&p.tma_c);
    for (int i = 0; i < T::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < T::ACC_STAGES; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], (CG2 ? 8 : 4) * NG);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG2) { tmem_alloc_cg2(tmem_ptr, T::TMEM_COLS); tmem_relinquish_cg2(); }
    else { tmem_alloc(tmem_ptr, T::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before_sync();
  if constexpr (CG2) cluster_sync_all(); else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  // Everything above (barrier init, TMEM allocation, descriptor prefetch, the cluster handshake) touches no global
  // memory and may run under the tail of the previous kernel; operands, scale vectors and outputs come after this.
  pdl_enter();

  if (warp == 0) {
    // ===================== TMA producer =====================
    // MEASURED (B200, nViT-B/16 step): an extra cursor issuing cp.async.bulk.prefetch.tensor (L2 prefetch) six k-blocks
    // ahead of the loads dropped every GEMM variant from 65-90 % to ~43 % tensor-pipe activity - the main loop is bound
    // by L2 request throughput, not latency, and the prefetch doubles the requests.  Likewise anything but a handful of
    // integer instructions per k-block in this single thread shows up directly as lost tensor time, so all tile
    // arithmetic is hoisted out of the k loop.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = smem_u32(full_bar);
      for (int u = unit0; u < total_units; u += unit_stride) {
        const int t = fast_div(u, p.splits, p.magic_splits);
        const int split = u - t * p.splits;
        int m_tile, n_blk;
        tile_coords(p, t, m_tile, n_blk);
        const int m_blk = m_tile * (CG2 ? 2 : 1) + (int)cta_rank;   // this CTA's 128-row block
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        const int a_row = m_blk * T::BM;
        const int b_row = SWIGLU ? ((CG2 && cta_rank) ? p.swiglu_half : 0) + n_blk * 128 : n_blk * BN + (int)cta_rank * T::BN_CTA;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * T::STAGE_BYTES;
          uint8_t* sB = sA + T::A_BYTES;
          // pair mode: both CTAs' bytes are counted on the rank-0 CTA's barrier, which its MMA warp waits on
          uint32_t fb = full0 + stage * 8;
          if constexpr (CG2) {
            fb = mapa_shared(fb, 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * T::STAGE_BYTES);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], T::STAGE_BYTES);
          }
          auto load = [&](const CUtensorMap* m, void* dst, int c0, int c1) {
            if constexpr (CG2) tma_load_2d_cg2(m, fb, dst, c0, c1);
            else tma_load_2d(m, &full_bar[stage], dst, c0, c1);
          };
          const int k0 = kb * T::BK;
          if constexpr (!A_MN) {
            load(&p.tma_a, sA, k0, a_row);
          } else {
#pragma unroll
            for (int j = 0; j < T::BM / 64; ++j) load(&p.tma_a, sA + j * (T::BK * 128), a_row + j * 64, k0);
          }
          if constexpr (SWIGLU) {
            load(&p.tma_b, sB, k0, b_row);
            if constexpr (!CG2) load(&p.tma_b, sB + 128 * 128, k0, p.swiglu_half + b_row);
          } else if constexpr (!B_MN) {
            load(&p.tma_b, sB, k0, b_row);
          } else {
#pragma unroll
            for (int j = 0; j < T::BN_CTA / 64; ++j) load(&p.tma_b, sB + j * (T::BK * 128), b_row + j * 64, k0);
          }
          if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (rank-0 CTA of a pair issues for both) =====================
    // The whole warp runs the loop so that barrier addresses and descriptors live in uniform registers; only the
    // tcgen05 instructions themselves are predicated on one elected lane.  Descriptors are base + constant offsets:
    // the tensor pipe drains a 128x256x16 step in 128 cycles, so the issue path between two MMAs must stay short.
    if (cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(CG2 ? 256 : 128, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      const uint32_t smem_a0 = smem_u32(smem);
      const uint64_t da_base = umma_smem_desc(smem_a0, T::A_LBO, T::SBO);
      const uint64_t db_base = umma_smem_desc(smem_a0 + T::A_BYTES, T::B_LBO, T::SBO);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      const uint32_t tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = unit0; u < total_units; u += unit_stride) {
        const int split = u - fast_div(u, p.splits, p.magic_splits) * p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait_a(tempty0 + acc * 8, acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_a(full0 + stage * 8, phase);
          tc_fence_after_sync();
          const uint64_t da = da_base + static_cast<uint64_t>((stage * T::STAGE_BYTES) >> 4);
          const uint64_t db = db_base + static_cast<uint64_t>((stage * T::STAGE_BYTES) >> 4);
          const uint32_t first = kb > kb0 ? 1u : 0u;
          const uint32_t ebar = empty0 + stage * 8, fbar = tfull0 + acc * 8;
          const bool last = (kb == kb1 - 1);
          if (elect_one()) {
            if constexpr (CG2) {
              umma_bf16_ss_cg2(tmem_d, da, db, idesc, first);
#pragma unroll
              for (int k = 1; k < T::BK / T::UMMA_K; ++k)
                umma_bf16_ss_cg2_acc(tmem_d, da + k * (T::A_KSTEP >> 4), db + k * (T::B_KSTEP >> 4), idesc);
              umma_commit_cg2_a(ebar);            // frees the slot in both CTAs
              if (last) umma_commit_cg2_a(fbar);  // accumulator complete, both CTAs' epilogues
            } else {
              umma_bf16_ss(tmem_d, da, db, idesc, first);
#pragma unroll
              for (int k = 1; k < T::BK / T::UMMA_K; ++k)
                umma_bf16_ss_acc(tmem_d, da + k * (T::A_KSTEP >> 4), db + k * (T::B_KSTEP >> 4), idesc);
              umma_commit_a(ebar);            // smem slot free once these MMAs retire
              if (last) umma_commit_a(fbar);  // accumulator complete
            }
          }
          __syncwarp();
          if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == T::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9: two groups of four warps) =====================
    // TMEM -> registers -> (bias / scale / gate) -> 128B-swizzled shared staging -> TMA store (or TMA reduce-add for
    // accumulating / split-K fp32 outputs).  The bulk store is asynchronous and fully coalesced; rows/columns outside
    // [M, N] are clipped by the tensor map.  Group g (warps 2-5 / 6-9) takes the column chunks with index % 2 == g and
    // owns staging buffer g, so two warps share each scheduler and the two groups overlap each other's store drains.
    // `direct` keeps the per-thread global-store path for outputs whose pitch or base the TMA cannot address.
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int eg = (warp - 2) >> 2;  // epilogue group
    const int erow = q * 32 + lane;
    const int et = threadIdx.x - 64;  // 0 .. 128 NG - 1 over the epilogue threads
    const bool issuer = (lane == 0) && (((warp - 2) & 3) == 0);
    uint8_t* const gbuf = stg + eg * (T::NBUF * 16384);   // this group's staging buffers
    uint32_t store_ctr = 0;
    const int bar_id = 1 + eg;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = p.out_f32 ? ((p.ldc & 3) == 0) : ((p.ldc & 7) == 0);
    const bool vec2_ok = (p.C2 == nullptr) || ((p.ldc2 & 7) == 0);
    const bool use_vec = (p.bias != nullptr) || (p.colscale != nullptr);
    // GATEB: raw u / v of this group's NEXT 64-column chunk, fetched a whole chunk ahead.  The loads are warp-cooperative
    // so that every request is a full 128-byte line (instruction k: lane l takes 16 bytes `l & 7` of row 4k + l/8 of the
    // warp's 32 rows); the pieces reach their owner threads through the group's staging buffers.  MEASURED: the obvious
    // per-thread form (each thread reading its own row, 32 bytes at a time) made this kernel 2.5x slower than its main
    // loop - 32 sector requests per warp instruction against an L2 request rate the operand loads already saturate.
    uint4 un[GATEB ? 8 : 1], vn[GATEB ? 8 : 1];
    auto gate_fetch = [&](int mb, int nb, int s2) {
      if constexpr (GATEB) {
        const int col = nb * BN + (eg + NG * s2) * 64 + (lane & 7) * 8;
        const __nv_bfloat16* base = p.gate_uv + (static_cast<long long>(mb) * T::BM + q * 32 + (lane >> 3)) * p.ld_uv + col;
        const int rows_left = p.M - (mb * T::BM + q * 32 + (lane >> 3));   // row 4k + l/8 is valid iff 4k < rows_left
        const bool col_ok = col < p.N && NVIT_DBG(p) != 3;      // dbg 3 (measurement aid): no u|v loads
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (col_ok && 4 * k < rows_left) {
            un[k] = ldg_u4_stream(base + static_cast<long long>(4 * k) * p.ld_uv);
            vn[k] = ldg_u4_stream(base + static_cast<long long>(4 * k) * p.ld_uv + p.swiglu_half);
          } else {
            un[k] = make_uint4(0u, 0u, 0u, 0u);
            vn[k] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
    };
    // GATEB with four groups: the same pieces travel by cp.async (global -> shared, no registers) straight into the warp's
    // rows of the group's staging buffer, requested as soon as the previous tile's stores have left that buffer - a whole
    // tile period before they are needed.  MEASURED on the register form (ncu source view, round 2): 42 % of the warp
    // samples of this kernel waited on the long scoreboard right here (u|v loads "fetched here and now"), another 14 % on
    // the two CTA-wide barriers around the per-tile scale vector.
    auto gate_prefetch = [&](int mb, int nb) {
      if constexpr (GATEB && NG == 4) {
        const int col = nb * BN + eg * 64 + (lane & 7) * 8;
        const __nv_bfloat16* base = p.gate_uv + (static_cast<long long>(mb) * T::BM + q * 32 + (lane >> 3)) * p.ld_uv + col;
        const int rows_left = p.M - (mb * T::BM + q * 32 + (lane >> 3));
        const bool col_ok = col < p.N;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr = q * 32 + 4 * k + (lane >> 3);
          const uint32_t off = rr * 128 + (((lane & 7) ^ (rr & 7)) << 4);
          const bool ok = col_ok && 4 * k < rows_left;
          const __nv_bfloat16* src = ok ? base + static_cast<long long>(4 * k) * p.ld_uv : p.gate_uv;
          cp_async16_zfill(gbuf + off, src, ok);
          cp_async16_zfill(gbuf + 16384 + off, ok ? src + p.swiglu_half : p.gate_uv, ok);
        }
      }
    };
    // GATEB: this thread's entries of the next tile's scale vectors (two groups: thread et takes column et of the tile; four
    // groups: the first 64 threads of a group take its 64 columns, so that the staging needs only group-wide barriers)
    const bool vec_thread = GATEB && (NG == 4 ? ((et & 127) < 64) : (et < 256));
    const int vec_col = NG == 4 ? eg * 64 + (et & 63) : et;       // column of the 256-wide tile
    float gate_su = 1.f, gate_sv = 1.f;
    if constexpr (GATEB) {
      if (!use_vec) {      // no scale vector (cross-attention gate): the epilogue multiplies by ones
        for (int i = et; i < 512; i += 128 * NG) s_vec[i] = 1.f;
        named_bar_sync(3, 128 * NG);
      }
      if (unit0 < total_units && NVIT_DBG(p) != 1) {
        int mt0, nb0;
        tile_coords(p, fast_div(unit0, p.splits, p.magic_splits), mt0, nb0);
        if constexpr (NG == 2) gate_fetch(mt0 * (CG2 ? 2 : 1) + (int)cta_rank, nb0, 0);
        else gate_prefetch(mt0 * (CG2 ? 2 : 1) + (int)cta_rank, nb0);
      }
      if (unit0 < total_units && use_vec && vec_thread) {
        int mt0, nb0;
        tile_coords(p, fast_div(unit0, p.splits, p.magic_splits), mt0, nb0);
        const int j0 = min(nb0 * TILE_N + vec_col, p.N - 1);
        gate_su = __ldg(p.colscale + j0) * p.colscale_mul;
        gate_sv = __ldg(p.colscale + p.swiglu_half + j0) * p.colscale_mul;
      }
    }
    for (int u = unit0; u < total_units; u += unit_stride) {
      const int t = fast_div(u, p.splits, p.magic_splits);
      int m_tile, n_blk;
      tile_coords(p, t, m_tile, n_blk);
      const int m_blk = m_tile * (CG2 ? 2 : 1) + (int)cta_rank;
      const int row = m_blk * T::BM + erow;
      if (use_vec) {
        // per-tile column vectors (bias, scale) staged once in shared memory instead of per-element global loads
        if constexpr (GATEB) {
          // [0,256): u scales of the tile's columns, [256,512): v scales.  The values were fetched one tile ahead
          // (gate_su / gate_sv), so no global-load latency sits between the two barriers.
          // (four groups: each group stages and reads only its own 64 columns, so its own 128-thread barrier is enough)
          if constexpr (NG == 4) named_bar_sync(5 + eg, 128); else named_bar_sync(3, 128 * NG);
          if (vec_thread) {
            s_vec[vec_col] = gate_su;
            s_vec[256 + vec_col] = gate_sv;
            if (u + unit_stride < total_units) {
              int mtn, nbn;
              tile_coords(p, fast_div(u + unit_stride, p.splits, p.magic_splits), mtn, nbn);
              const int jn = min(nbn * TILE_N + vec_col, p.N - 1);
              gate_su = __ldg(p.colscale + jn) * p.colscale_mul;
              gate_sv = __ldg(p.colscale + p.swiglu_half + jn) * p.colscale_mul;
            }
          }
          if constexpr (NG == 4) named_bar_sync(5 + eg, 128); else named_bar_sync(3, 128 * NG);
        } else {
        named_bar_sync(3, 256);  // both groups are done with the previous tile's vectors
        {
          int j = n_blk * TILE_N + (SWIGLU ? (et & 127) : et);
          j = min(j, p.N - 1);
          if (SWIGLU) j += (et >> 7) * p.swiglu_half;   // entries 0..127: u scales, 128..255: v scales
          if constexpr (GATEB) {
          } else if (SWIGLU || et < BN) {
            s_vec[et] = p.colscale ? __ldg(p.colscale + (p.qk_cols > 0 ? j % p.qk_period : j)) * p.colscale_mul : 1.f;
            s_vec[256 + et] = p.bias ? __ldg(p.bias + j) : 0.f;
          }
        }
        named_bar_sync(3, 256);
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      auto release_tmem = [&]() {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG2 && GATEB) mbar_arrive_cluster_release(mapa_shared(smem_u32(&tmem_empty[acc]), 0));  // (see common.cuh)
          else if constexpr (CG2) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));  // the issuing CTA's barrier
          else mbar_arrive(&tmem_empty[acc]);
        }
      };
      // write 32 packed words (a [row][64 bf16] or [row][32 fp32] line) into this group's staging buffer and store it
      auto stage_store = [&](const uint32_t (&src)[32], const CUtensorMap* map, int n0, bool reduce) {
        if (NVIT_DBG(p) == 2) return;                // measurement aid: TMEM reads + math, no staging / stores
        uint8_t* buf = gbuf + (store_ctr % T::NBUF) * 16384;
        ++store_ctr;
        if (issuer) bulk_wait_group_read<T::NBUF - 1>();   // the store that last used this buffer has drained it
        named_bar_sync(bar_id, 128);
        if (NVIT_DBG(p) != 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(buf + erow * 128 + ((j ^ (erow & 7)) << 4)) =
                make_uint4(src[4 * j], src[4 * j + 1], src[4 * j + 2], src[4 * j + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (issuer && NVIT_DBG(p) != 3) {
          if (reduce) tma_reduce_add_2d(map, buf, n0, m_blk * T::BM);
          else tma_store_2d(map, buf, n0, m_blk * T::BM);
          bulk_commit_group();
        }
      };
      // MEASURED (scripts/gemm_bench.py, nvit_gemm_debug): main loop alone 1555-1700 TFLOP/s; + TMEM reads and conversion
      // -7 %; + staging and barriers -6 %; + the TMA stores themselves -13 % (plain bf16) to -25 % (gate, three outputs).
      // Sending the staged tile out through coalesced LSU stores instead was slower still (qkv 1151 -> 941 TFLOP/s), and
      // two or three staging buffers per group (at the price of ring stages) changed nothing, nor did an L2 evict_first
      // hint on the stores (createpolicy + .L2::cache_hint): the cost follows the output bytes, not the mechanism.
      if constexpr (GATEB) {
        // Backward of x = (u su) * silu(v sv) fused behind dx = dy W (model.py:148-155 backward): the accumulator holds
        // dL/dx for 256 gate columns; this thread combines its row with the raw u, v of the forward pass and emits
        // dL/du_raw and dL/dv_raw as two [128 x 64] bf16 tiles per chunk.  (dL/dsuv follows from the c_fc weight
        // gradient: nvit_rowdot_div.)
        if (NVIT_DBG(p) == 1) {
          release_tmem();
        } else {
#pragma unroll
          for (int s2 = 0; s2 < 4 / NG; ++s2) {
            const int c = eg + NG * s2;
            const int n0 = n_blk * BN + c * 64;
            const bool live = n0 < p.N;   // uniform over the group
            // four groups: one chunk per group and tile, fetched here and now - with four warps per scheduler the other
            // groups' work covers the load latency, and no registers are held across the arithmetic
            uint8_t* const sbuf = gbuf + (store_ctr % T::GATE_SETS) * 32768;    // this chunk's {du, dv} buffer set
            if constexpr (NG == 4) {
              // the u|v pieces of this chunk were requested (cp.async) when the previous tile's stores had left the buffer
              cp_async_wait_all();
              __syncwarp();       // a warp's 32 rows are copied in and consumed by that warp alone
            } else if (live) {
              ++store_ctr;
              if (lane == 0) bulk_wait_group_read<T::GATE_SETS - 1>();  // this warp's stores from this set have drained
              __syncwarp();
              // hand the prefetched pieces to their owner rows: staging buffer 0 <- u chunk, buffer 1 <- v chunk
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int rr = q * 32 + 4 * k + (lane >> 3);
                const uint32_t off = rr * 128 + (((lane & 7) ^ (rr & 7)) << 4);
                *reinterpret_cast<uint4*>(sbuf + off) = un[k];
                *reinterpret_cast<uint4*>(sbuf + 16384 + off) = vn[k];
              }
              __syncwarp();       // a warp's 32 rows are loaded and consumed by that warp alone
            }
            if constexpr (NG == 2) {
              if (s2 == 0) {
                gate_fetch(m_blk, n_blk, 1);
              } else if (u + unit_stride < total_units) {
                int mt2, nb2;
                tile_coords(p, fast_div(u + unit_stride, p.splits, p.magic_splits), mt2, nb2);
                gate_fetch(mt2 * (CG2 ? 2 : 1) + (int)cta_rank, nb2, 0);
              }
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(taddr + c * 64 + hh * 32, r);
              tmem_wait_ld();
              if (s2 == 4 / NG - 1 && hh == 1) release_tmem();
              if (live && NVIT_DBG(p) != 4) {                         // dbg 4 (measurement aid): no gate arithmetic
                // the results go back to the addresses the inputs came from, so the compiler may not move the next
                // piece's loads above this piece's stores: fetch one piece ahead by hand
                uint4 u_nx = *reinterpret_cast<const uint4*>(sbuf + erow * 128 + (((hh * 4) ^ (erow & 7)) << 4));
                uint4 v_nx = *reinterpret_cast<const uint4*>(sbuf + 16384 + erow * 128 + (((hh * 4) ^ (erow & 7)) << 4));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t off = erow * 128 + (((hh * 4 + j) ^ (erow & 7)) << 4);
                  const uint4 u8 = u_nx, v8 = v_nx;
                  if (j < 3) {
                    const uint32_t off2 = erow * 128 + (((hh * 4 + j + 1) ^ (erow & 7)) << 4);
                    u_nx = *reinterpret_cast<const uint4*>(sbuf + off2);
                    v_nx = *reinterpret_cast<const uint4*>(sbuf + 16384 + off2);
                  }
                  const uint32_t uc[4] = {u8.x, u8.y, u8.z, u8.w}, vc[4] = {v8.x, v8.y, v8.z, v8.w};
                  uint32_t ou[4], ov[4];
                  // Two gate columns at a time on packed fp32 pairs (FFMA2): the epilogue's cost is its instruction count
                  // (MEASURED in round 1: 15 % of the warp samples issuing, everything else waiting on dependent results),
                  // and the pair form halves the fp32 instructions per element (13 -> 6.5).
                  const f32x2 half2 = pack2(0.5f, 0.5f), one2 = pack2(1.f, 1.f), neg2 = pack2(-1.f, -1.f);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int cc = c * 64 + hh * 32 + 8 * j + 2 * e;      // column within the tile
                    // (s_vec holds ones when there is no scale vector: no select inside this loop)
                    const float2 a = *reinterpret_cast<const float2*>(s_vec + cc), b = *reinterpret_cast<const float2*>(s_vec + 256 + cc);
                    const f32x2 su = pack2(a.x, a.y), sv = pack2(b.x, b.y);
                    const f32x2 g = pack2(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1]));
                    const f32x2 vh = mul2(bf16x2_to_f32x2(vc[e]), sv);
                    float h0, h1;
                    unpack2(mul2(vh, half2), h0, h1);
                    const f32x2 sg = fma2(pack2(tanh_approx_(h0), tanh_approx_(h1)), half2, half2);   // sigmoid(vh)
                    const f32x2 sl = mul2(vh, sg);                                               // silu(vh)
                    const f32x2 dsl = fma2(sl, fma2(sg, neg2, one2), sg);                        // silu' = sg + silu (1 - sg)
                    const f32x2 gs = mul2(g, su);                                                // g su
                    // dL/du_raw = g silu(vh) su ;  dL/dv_raw = g (u su) silu'(vh) sv
                    ou[e] = f32x2_to_bf16x2(mul2(gs, sl));
                    ov[e] = f32x2_to_bf16x2(mul2(mul2(gs, bf16x2_to_f32x2(uc[e])), mul2(dsl, sv)));
                  }
                  *reinterpret_cast<uint4*>(sbuf + off) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
                  *reinterpret_cast<uint4*>(sbuf + 16384 + off) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
                }
              }
            }
            if (live) {
              fence_proxy_async_smem();
              // every warp ships its own 32 rows (tma_c2: box 64 x 32), so the warps of a group never wait for each other
              // (MEASURED: 397 -> 381 us single-CTA, 410 -> 404 us as pairs, against one 128-row store per group)
              __syncwarp();
              if (lane == 0 && NVIT_DBG(p) != 2) {                    // dbg 2 (measurement aid): no stores
                tma_store_2d(&p.tma_c2, sbuf + q * 4096, n0, m_blk * T::BM + q * 32);
                tma_store_2d(&p.tma_c2, sbuf + 16384 + q * 4096, p.swiglu_half + n0, m_blk * T::BM + q * 32);
                bulk_commit_group();
              }
            }
            if constexpr (NG == 4) {
              // This warp's next piece of work is a whole tile away: wait here until the stores have read its rows, then
              // have the next tile's u|v pieces copied into them while the other groups and the main loop carry on.
              if (u + unit_stride < total_units) {
                if (lane == 0) bulk_wait_group_read<0>();
                __syncwarp();
                int mt2, nb2;
                tile_coords(p, fast_div(u + unit_stride, p.splits, p.magic_splits), mt2, nb2);
                gate_prefetch(mt2 * (CG2 ? 2 : 1) + (int)cta_rank, nb2);
              }
            }
          }
        }
      } else if (NVIT_DBG(p) == 1) {                      // measurement aid: main loop only
        release_tmem();
      } else if (!p.direct) {
        if constexpr (SWIGLU) {
          // group g takes gate outputs [64g, 64g+64) of the tile: x, raw u and raw v leave as three [128 x 64] bf16 tiles
          uint32_t uo[32], vo[32];
          {
            uint32_t r[32], r2[32];
            tmem_ld_32x32b_x32(taddr + eg * 64, r);
            tmem_ld_32x32b_x32(taddr + eg * 64 + 32, r2);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {   // c_fc output is bf16 under autocast
              uo[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
              uo[16 + i] = pack_bf16(__uint_as_float(r2[2 * i]), __uint_as_float(r2[2 * i + 1]));
            }
            tmem_ld_32x32b_x32(taddr + 128 + eg * 64, r);
            tmem_ld_32x32b_x32(taddr + 128 + eg * 64 + 32, r2);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              vo[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
              vo[16 + i] = pack_bf16(__uint_as_float(r2[2 * i]), __uint_as_float(r2[2 * i + 1]));
            }
          }
          release_tmem();
          const int n0 = n_blk * 128 + eg * 64;
          if (n0 < p.N) {  // uniform over the group
            uint32_t xo[32];
            // x = (u su) * silu(v sv) on packed fp32 pairs, sigmoid through tanh.approx (one MUFU op per element instead of
            // ex2 + rcp, as in the backward epilogue): ~6 instructions per element instead of ~10
            const f32x2 half2 = pack2(0.5f, 0.5f);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              f32x2 su = pack2(1.f, 1.f), sv = su;
              if (p.colscale) {
                const float2 a = *reinterpret_cast<const float2*>(s_vec + eg * 64 + 2 * i);
                const float2 b = *reinterpret_cast<const float2*>(s_vec + 128 + eg * 64 + 2 * i);
                su = pack2(a.x, a.y);
                sv = pack2(b.x, b.y);
              }
              const f32x2 uh = mul2(bf16x2_to_f32x2(uo[i]), su), vh = mul2(bf16x2_to_f32x2(vo[i]), sv);
              float h0, h1;
              unpack2(mul2(vh, half2), h0, h1);
              const f32x2 sg = fma2(pack2(tanh_approx_(h0), tanh_approx_(h1)), half2, half2);
              xo[i] = f32x2_to_bf16x2(mul2(mul2(uh, vh), sg));
            }
            stage_store(xo, &p.tma_c, n0, false);
            if (p.C2) {
              stage_store(uo, &p.tma_c2, n0, false);
              stage_store(vo, &p.tma_c3, n0, false);
            }
          }
        } else if (!p.out_f32) {
          // bf16 output: chunks of 64 columns (128 B rows, 128B swizzle).  (Pulling all of a group's chunks out of TMEM
          // first, to hand the accumulator back earlier, was measured slower: 65 -> 54 % tensor activity.)
          constexpr int NC = BN / 64;
#pragma unroll 1
          for (int c = eg; c < NC; c += 2) {
            uint32_t r[32], r2[32];
            tmem_ld_32x32b_x32(taddr + c * 64, r);
            tmem_ld_32x32b_x32(taddr + c * 64 + 32, r2);
            tmem_wait_ld();
            if (c + 2 >= NC) release_tmem();
            const int n0 = n_blk * BN + c * 64;
            if (n0 >= p.N) continue;
            uint32_t o[32];
            if (p.qk_cols > 0) {
              // unit-norm q/k (model.py:108-119): this chunk is one head of one token; the whole row segment is in registers
              const bool normed = n0 < p.qk_cols;      // uniform over the group
              float ss = 0.f;
#pragma unroll
              for (int i = 0; i < 64; ++i) {
                float v = __uint_as_float(i < 32 ? r[i] : r2[i - 32]) + (p.bias ? s_vec[256 + c * 64 + i] : 0.f);
                if (i < 32) r[i] = __float_as_uint(v); else r2[i - 32] = __float_as_uint(v);
                ss += v * v;
              }
              const float inv = (normed && ss > 0.f) ? rsqrtf(ss) : (normed ? 0.f : 1.f);
#pragma unroll
              for (int i = 0; i < 64; i += 2) {
                float v0 = __uint_as_float(i < 32 ? r[i] : r2[i - 32]) * inv;
                float v1 = __uint_as_float(i < 32 ? r[i + 1] : r2[i - 31]) * inv;
                if (normed) {
                  const float2 cs = *reinterpret_cast<const float2*>(s_vec + c * 64 + i);
                  v0 *= cs.x;
                  v1 *= cs.y;
                }
                o[i >> 1] = pack_bf16(v0, v1);
              }
              if (normed && row < p.M) p.qk_inv[static_cast<long long>(row) * p.ld_inv + (n0 >> 6)] = inv;
              stage_store(o, &p.tma_c, n0, false);
              continue;
            }
#pragma unroll
            for (int i = 0; i < 64; i += 2) {
              float v0 = __uint_as_float(i < 32 ? r[i] : r2[i - 32]);
              float v1 = __uint_as_float(i < 32 ? r[i + 1] : r2[i - 31]);
              if (use_vec) {
                const float2 bb = *reinterpret_cast<const float2*>(s_vec + 256 + c * 64 + i);
                const float2 cs = *reinterpret_cast<const float2*>(s_vec + c * 64 + i);
                v0 = (v0 + bb.x) * cs.x;
                v1 = (v1 + bb.y) * cs.y;
              }
              if (p.rowadd) {
                const int j0 = min(n0 + i, p.N - 1), j1 = min(n0 + i + 1, p.N - 1);
                const float* ra = p.rowadd + static_cast<long long>(row % p.rowadd_period) * p.N;
                v0 += __ldg(ra + j0); v1 += __ldg(ra + j1);
              }
              o[i >> 1] = pack_bf16(v0, v1);
            }
            stage_store(o, &p.tma_c, n0, false);
          }
        } else if (p.C2 == nullptr) {
          // fp32 output: chunks of 32 columns (128 B rows, 128B swizzle)
          constexpr int NC = BN / 32;
          const bool reduce = p.accumulate || p.atomic;
#pragma unroll 1
          for (int c = eg; c < NC; c += 2) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + c * 32, r);
            tmem_wait_ld();
            if (c + 2 >= NC) release_tmem();
            const int n0 = n_blk * BN + c * 32;
            if (n0 >= p.N) continue;
            if (use_vec || p.rowadd) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float v = __uint_as_float(r[i]);
                if (use_vec) v = (v + s_vec[256 + c * 32 + i]) * s_vec[c * 32 + i];
                if (p.rowadd) v += __ldg(p.rowadd + static_cast<long long>(row % p.rowadd_period) * p.N + min(n0 + i, p.N - 1));
                r[i] = __float_as_uint(v);
              }
            }
            stage_store(r, &p.tma_c, n0, reduce);
          }
        } else {
          // fp32 output with a bf16 side copy (patch embedding): group 0 alone, fp32 tile in buffer 0 and the dense
          // [128][32] bf16 copy in buffer 1
          constexpr int NC = BN / 32;
          if (eg == 1) {
            release_tmem();
          } else {
#pragma unroll 1
            for (int c = 0; c < NC; ++c) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(taddr + c * 32, r);
              tmem_wait_ld();
              if (c == NC - 1) release_tmem();
              const int n0 = n_blk * BN + c * 32;
              if (n0 >= p.N) continue;
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                v[i] = __uint_as_float(r[i]);
                if (use_vec) v[i] = (v[i] + s_vec[256 + c * 32 + i]) * s_vec[c * 32 + i];
                if (p.rowadd) v[i] += __ldg(p.rowadd + static_cast<long long>(row % p.rowadd_period) * p.N + min(n0 + i, p.N - 1));
              }
              if (issuer) bulk_wait_group_read<0>();
              named_bar_sync(bar_id, 128);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(stg + erow * 128 + ((j ^ (erow & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(stg + 16384 + erow * 64 + j * 16) =
                    make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                               pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
              fence_proxy_async_smem();
              named_bar_sync(bar_id, 128);
              if (issuer) {
                if (p.accumulate || p.atomic) tma_reduce_add_2d(&p.tma_c, stg, n0, m_blk * T::BM);
                else tma_store_2d(&p.tma_c, stg, n0, m_blk * T::BM);
                tma_store_2d(&p.tma_c2, stg + 16384, n0, m_blk * T::BM);
                bulk_commit_group();
              }
            }
          }
        }
      } else {
      constexpr int NCHUNK = TILE_N / 32;
#pragma unroll 1
      for (int c = eg; c < NCHUNK; c += 2) {
        uint32_t r[32];
        uint32_t r2[SWIGLU ? 32 : 1];
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        if constexpr (SWIGLU) tmem_ld_32x32b_x32(taddr + 128 + c * 32, r2);
        tmem_wait_ld();
        if (c + 2 >= NCHUNK) release_tmem();
        const int n0 = n_blk * TILE_N + c * 32;
        if (n0 < p.N && row < p.M) {
          const int valid = min(32, p.N - n0);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if constexpr (SWIGLU) {
            float w[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[i] = round_bf16(v[i]);  // the reference's c_fc output is bf16 under autocast
              w[i] = round_bf16(__uint_as_float(r2[i]));
            }
            if (p.C2) {
              store_bf16x32(p.C2 + static_cast<long long>(row) * p.ldc2 + n0, v, valid, vec2_ok);
              store_bf16x32(p.C2 + static_cast<long long>(row) * p.ldc2 + p.swiglu_half + n0, w, valid, vec2_ok);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float su = 1.f, sv = 1.f;
              if (p.colscale && i < valid) {
                su = __ldg(p.colscale + n0 + i) * p.colscale_mul;
                sv = __ldg(p.colscale + p.swiglu_half + n0 + i) * p.colscale_mul;
              }
              v[i] = silu_mul(v[i] * su, w[i] * sv);
            }
            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<long long>(row) * p.ldc + n0, v, valid, vec_ok);
          } else {
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < valid) v[i] += __ldg(p.bias + n0 + i);
            }
            if (p.colscale) {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < valid) v[i] *= __ldg(p.colscale + n0 + i) * p.colscale_mul;
            }
            if (p.rowadd) {
              const float* ra = p.rowadd + static_cast<long long>(row % p.rowadd_period) * p.N + n0;
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < valid) v[i] += __ldg(ra + i);
            }
            if (p.out_f32) {
              float* dst = reinterpret_cast<float*>(p.C) + static_cast<long long>(row) * p.ldc + n0;
              if (p.atomic) {
#pragma unroll
                for (int i = 0; i < 32; ++i) if (i < valid) atomicAdd(dst + i, v[i]);
              } else if (valid == 32 && vec_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                  if (p.accumulate) {
                    const float4 old = *reinterpret_cast<const float4*>(dst + i);
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                  }
                  *reinterpret_cast<float4*>(dst + i) = o;
                }
              } else {
                for (int i = 0; i < valid; ++i) dst[i] = p.accumulate ? dst[i] + v[i] : v[i];
              }
              if (p.C2) store_bf16x32(p.C2 + static_cast<long long>(row) * p.ldc2 + n0, v, valid, vec2_ok);
            } else {
              store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<long long>(row) * p.ldc + n0, v, valid, vec_ok);
            }
          }
        }  // in bounds
        __syncwarp();  // re-converge before the next warp-aligned tcgen05.ld
      }
      }  // direct
      if (++acc == T::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    // staged stores must have left shared memory (and landed) before the CTA retires (GATEB: every warp stores its own rows)
    if (GATEB ? (lane == 0) : issuer) bulk_wait_group<0>();
  }

  tc_fence_before_sync();
  if constexpr (CG2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    if constexpr (CG2) tmem_dealloc_cg2(tmem_base, T::TMEM_COLS);
    else tmem_dealloc(tmem_base, T::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static std::atomic<PFN_encodeTiled> fn{nullptr};     // the entry point is process-wide; racing threads store the same value
  PFN_encodeTiled f = fn.load(std::memory_order_acquire);
  if (f) return f;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  f = reinterpret_cast<PFN_encodeTiled>(ptr);
  fn.store(f, std::memory_order_release);
  return f;
}

// Encoded tensor maps are kept: a map is a pure function of (base, type, swizzle, rank, dims, pitches, box) and every buffer of
// the engine is static, so a step re-encodes nothing (up to five maps per GEMM, ~150 GEMMs and 26 attention launches per
// step before).  A re-allocated buffer at the same address with the same shape yields the identical map, so entries never
// go stale; the table is simply emptied when it reaches TMAP_CACHE_MAX entries.
struct TmapKey {
  uint64_t w[11];   // base, (dt, sw, rank), dims[3], pitches[2] (bytes), box[3]; unused slots zero
  bool operator==(const TmapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t v : k.w) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); h *= 0xFF51AFD7ED558CCDull; }
    return (size_t)(h ^ (h >> 32));
  }
};
static constexpr size_t TMAP_CACHE_MAX = 8192;
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::atomic<long long> g_tmap_encodes{0}, g_tmap_hits{0};

// Tensor map of rank 2 or 3 (dims innermost first, strides in elements for dims 1..rank-1).
static int make_tmap(CUtensorMap* m, const void* base, CUtensorMapDataType dt, int esize, CUtensorMapSwizzle sw, int rank,
                     const uint64_t* dims, const uint64_t* strides_elems, const uint32_t* box) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    nvit_set_error("cuTensorMapEncodeTiled entry point not available");
    return NVIT_ERR_DRIVER;
  }
  cuuint64_t d[3], s[2];
  cuuint32_t b[3], e[3] = {1, 1, 1};
  if (reinterpret_cast<uintptr_t>(base) & 15) {
    nvit_set_error("TMA operand base %p must be 16-byte aligned", base);
    return NVIT_ERR_ARG;
  }
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) {
    s[i] = strides_elems[i] * esize;
    if (s[i] & 15) {
      nvit_set_error("TMA operand pitch %llu elements is not a multiple of 16 bytes", (unsigned long long)strides_elems[i]);
      return NVIT_ERR_ARG;
    }
  }
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.w[0] = reinterpret_cast<uintptr_t>(base);
  key.w[1] = ((uint64_t)dt << 32) | ((uint64_t)sw << 8) | (uint64_t)rank;
  for (int i = 0; i < rank; ++i) { key.w[2 + i] = d[i]; key.w[7 + i] = b[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.w[5 + i] = s[i];
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *m = it->second;
      g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
      return NVIT_OK;
    }
  }
  g_tmap_encodes.fetch_add(1, std::memory_order_relaxed);
  CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    nvit_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu box %u %u)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return NVIT_ERR_DRIVER;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() >= TMAP_CACHE_MAX) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *m);
  }
  return NVIT_OK;
}

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box) {
  return make_tmap(m, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_128B, rank, dims, strides_elems, box);
}

// 2-D output map: [rows, cols] with a row pitch; box = box_cols x 128 rows
static int make_out_tmap(CUtensorMap* m, const void* base, bool f32, bool swizzled, uint64_t cols, uint64_t rows, uint64_t pitch,
                         uint32_t box_cols) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {pitch};
  const uint32_t box[2] = {box_cols, 128};
  return make_tmap(m, base, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, f32 ? 4 : 2,
                   swizzled ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, 2, dims, strides, box);
}

static bool tma_addressable(const void* base, long long pitch_elems, int esize) {
  return base && (reinterpret_cast<uintptr_t>(base) & 15) == 0 && ((pitch_elems * esize) & 15) == 0;
}

static int make_tmap_bf16_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                             uint32_t box_inner, uint32_t box_outer) {
  const uint64_t dims[2] = {inner, outer};
  const uint64_t strides[1] = {pitch_elems};
  const uint32_t box[2] = {box_inner, box_outer};
  return make_tmap_bf16(m, base, 2, dims, strides, box);
}

static std::atomic<int> g_raster_group{0};   // nvit_gemm_raster_group: width of the n bands of the tile order (0 = n fastest)
static int raster_group() { return g_raster_group.load(std::memory_order_relaxed); }

template <int BN, bool A_MN, bool B_MN, bool SWIGLU, bool CG2, bool GATEB = false, int NG = 2>
static int launch_gemm(GemmParams& p, const void* A, const void* B, long long lda, long long ldb, cudaStream_t stream) {
  using T = GemmTraits<BN, A_MN, B_MN, CG2, GATEB, SWIGLU, NG>;
  int rc;
  if (!A_MN) rc = make_tmap_bf16_2d(&p.tma_a, A, p.K, p.M, lda, T::BK, T::BM);
  else       rc = make_tmap_bf16_2d(&p.tma_a, A, p.M, p.K, lda, 64, T::BK);
  if (rc) return rc;
  if (SWIGLU)     rc = make_tmap_bf16_2d(&p.tma_b, B, p.K, 2ull * p.swiglu_half, ldb, T::BK, 128);
  else if (!B_MN) rc = make_tmap_bf16_2d(&p.tma_b, B, p.K, p.N, ldb, T::BK, T::BN_CTA);
  else            rc = make_tmap_bf16_2d(&p.tma_b, B, p.N, p.K, ldb, 64, T::BK);
  if (rc) return rc;
  // output maps for the staged epilogue (fall back to direct stores when the output is not TMA-addressable)
  p.direct = 1;
  if (GATEB) {
    // one map over d(uv) [M, 2F]: the du tile of a chunk goes to column n0, its dv tile to column F + n0
    if ((rc = make_out_tmap(&p.tma_c, p.C, false, true, 2ull * p.swiglu_half, p.M, p.ldc, 64))) return rc;
    {
      const uint64_t dims[2] = {2ull * p.swiglu_half, (uint64_t)p.M};
      const uint64_t strides[1] = {(uint64_t)p.ldc};
      const uint32_t box[2] = {64, 32};
      if ((rc = make_tmap_bf16(&p.tma_c2, p.C, 2, dims, strides, box))) return rc;
    }
    p.direct = 0;
  } else if (SWIGLU) {
    const bool ok = tma_addressable(p.C, p.ldc, 2) && (!p.C2 || (tma_addressable(p.C2, p.ldc2, 2) && (p.swiglu_half % 8) == 0));
    if (ok) {
      if ((rc = make_out_tmap(&p.tma_c, p.C, false, true, p.N, p.M, p.ldc, 64))) return rc;
      if (p.C2) {
        if ((rc = make_out_tmap(&p.tma_c2, p.C2, false, true, p.N, p.M, p.ldc2, 64))) return rc;
        if ((rc = make_out_tmap(&p.tma_c3, p.C2 + p.swiglu_half, false, true, p.N, p.M, p.ldc2, 64))) return rc;
      }
      p.direct = 0;
    }
  } else if (p.out_f32) {
    const bool ok = tma_addressable(p.C, p.ldc, 4) && (!p.C2 || tma_addressable(p.C2, p.ldc2, 2));
    if (ok) {
      if ((rc = make_out_tmap(&p.tma_c, p.C, true, true, p.N, p.M, p.ldc, 32))) return rc;
      if (p.C2 && (rc = make_out_tmap(&p.tma_c2, p.C2, false, false, p.N, p.M, p.ldc2, 32))) return rc;
      p.direct = 0;
    }
  } else if (tma_addressable(p.C, p.ldc, 2)) {
    if ((rc = make_out_tmap(&p.tma_c, p.C, false, true, p.N, p.M, p.ldc, 64))) return rc;
    p.direct = 0;
  }
  constexpr int TILE_N = SWIGLU ? 128 : BN;
  constexpr int TILE_M = CG2 ? 256 : 128;
  p.tiles_m = (p.M + TILE_M - 1) / TILE_M;
  p.tiles_n = (p.N + TILE_N - 1) / TILE_N;
  p.kb_total = (p.K + T::BK - 1) / T::BK;
  p.group_n = 0;
  if (raster_group() > 0 && p.tiles_n > raster_group()) {
    p.group_n = raster_group();
  } else if (raster_group() < 0) {
    // automatic: operand rows the tiles in flight pull in when they lie (in_flight / g) x g, for every band width g that
    // divides tiles_n; bands are used when they beat n-fastest by >= 10 % (on the nViT shapes: the gate GEMM, 24 or 32
    // tiles wide, g = 8)
    const int in_flight = CG2 ? nvit_num_sms() / 2 : nvit_num_sms();
    auto cost = [&](int g) { return (double)TILE_M * ((in_flight + g - 1) / g) + 256.0 * g; };   // B panels are 256 rows (BN, or u | v)
    if (BN == 256 || SWIGLU) {
      double best = 0.9 * cost(p.tiles_n < in_flight ? p.tiles_n : in_flight);
      for (int g = 2; g < p.tiles_n; ++g)
        if (p.tiles_n % g == 0 && cost(g) < best) { best = cost(g); p.group_n = g; }
    }
  }
  const int workers = CG2 ? nvit_num_sms() / 2 : nvit_num_sms();   // CTAs or CTA pairs
  int splits = p.splits;
  if (splits <= 0) {
    // auto split-K (fp32 reduce-add outputs only): fewest splits that fill the machine best
    splits = 1;
    if (p.out_f32 && !p.bias && !p.colscale && !p.rowadd && !p.C2) {
      const long long tiles = 1ll * p.tiles_m * p.tiles_n;
      double best = 0.0;
      for (int s = 1; s <= 16 && s <= p.kb_total; ++s) {
        const long long units = tiles * s;
        const double util = (double)units / (double)(((units + workers - 1) / workers) * workers);
        if (util > best + 0.02) { best = util; splits = s; }
      }
    }
  }
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  if (p.splits > 1) {
    p.atomic = 1;
    if (!p.accumulate)  // partial sums are added into C: start from zero
      NVIT_CUDA_CHECK(cudaMemset2DAsync(p.C, p.ldc * sizeof(float), 0, p.N * sizeof(float), p.M, stream));
  }
  static DeviceOnce once;      // per instantiation
  int dev;
  if (once.needed(&dev)) {
    NVIT_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, A_MN, B_MN, SWIGLU, CG2, GATEB, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         T::SMEM_BYTES));
    once.mark(dev);
  }
  const long long units = 1ll * p.tiles_m * p.tiles_n * p.splits;
  {
    auto magic = [&](int d) -> unsigned {      // see GemmParams::magic_*: dividends are unit / tile indices below `units` (+ one grid stride)
      const long long xmax = units + 2 * nvit_num_sms();
      return (d >= 2 && xmax * d < (1ll << 32)) ? static_cast<unsigned>((1ull << 32) / static_cast<unsigned>(d) + 1ull) : 0u;
    };
    p.magic_tiles_n = magic(p.tiles_n);
    p.magic_splits = magic(p.splits);
  }
  const int nwork = (int)(units < workers ? units : workers);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(CG2 ? 2 * nwork : nwork);
  cfg.blockDim = dim3(64 + 128 * NG);
  cfg.dynamicSmemBytes = T::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see pdl_wait() in common.cuh
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = nvit_pdl_enabled() ? 2 : 1;
  NVIT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, A_MN, B_MN, SWIGLU, CG2, GATEB, NG>, p));
  return NVIT_OK;
}

}  // namespace nvit

using namespace nvit;

static std::atomic<int> g_force_cg{0};  // 0 auto, 1 single-CTA tiles, 2 CTA pairs
static std::atomic<int> g_swiglu_cg{2}; // CTA-group mode of the gate GEMM under the auto policy (nvit_gemm_swiglu_cta_group)
static std::atomic<int> g_gateb_cg{2};  // ... and of the fused gate-backward GEMM (mode + 10 through the same hook)
static std::atomic<int> g_gateb_groups{4};  // epilogue groups of the fused gate-backward GEMM (mode 22 / 24 through the same hook)
static std::atomic<int> g_dbg{0};       // see GemmParams::dbg

extern "C" int nvit_gemm_bf16(const void* A, const void* B, void* C, void* C2_bf16, int64_t M, int64_t N, int64_t K,
                              int64_t lda, int64_t ldb, int64_t ldc, int64_t ldc2, int a_mn_major, int b_mn_major,
                              int out_f32, int accumulate, int splits, const float* bias, const float* colscale,
                              float colscale_mul, const float* rowadd, int64_t rowadd_period, int64_t swiglu_half,
                              void* stream) {
  NVIT_REQUIRE(A && B && C, "nvit_gemm_bf16: null operand");
  NVIT_REQUIRE(M > 0 && N > 0 && K > 0, "nvit_gemm_bf16: empty problem M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  NVIT_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "nvit_gemm_bf16: dimension overflow");
  NVIT_REQUIRE(!(splits > 1) || out_f32, "nvit_gemm_bf16: split-K needs an fp32 output (splits <= 0 picks automatically)");
  NVIT_REQUIRE(!accumulate || out_f32, "nvit_gemm_bf16: accumulate needs an fp32 output");
  NVIT_REQUIRE(!C2_bf16 || out_f32 || swiglu_half > 0, "nvit_gemm_bf16: the bf16 side output goes with an fp32 or swiglu output");
  NVIT_REQUIRE(!(splits > 1) || (!bias && !colscale && !rowadd && !C2_bf16), "nvit_gemm_bf16: split-K supports no epilogue");
  NVIT_REQUIRE(!rowadd || rowadd_period > 0, "nvit_gemm_bf16: rowadd needs a positive period");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.C = C;
  p.C2 = reinterpret_cast<__nv_bfloat16*>(C2_bf16);
  p.bias = bias;
  p.colscale = colscale;
  p.colscale_mul = colscale_mul;
  p.rowadd = rowadd;
  p.ldc = ldc;
  p.ldc2 = ldc2;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.out_f32 = out_f32;
  p.accumulate = accumulate;
  p.atomic = 0;
  p.rowadd_period = (int)(rowadd ? rowadd_period : 1);
  p.splits = splits;
  p.swiglu_half = (int)swiglu_half;
  p.dbg = g_dbg;
  // CTA pairs (256-row tiles, cta_group::2) pay off for long reductions and fp32 outputs (wgrad, accumulating dgrad);
  // MEASURED on the nViT-B/16 shapes, single-CTA 128-row tiles are a few percent faster for bf16 outputs with
  // K < 2048 (qkv / att_c_proj forward, plain mlp_c_proj dgrad).  The two gate GEMMs (swiglu forward, gate backward) have
  // epilogue-heavy tiles and run as pairs (MEASURED in situ: 49.0 / 49.6 vs 49.4 / 50.2 ms per step for the forward one).
  // nvit_gemm_force_cta_group pins either mode.
  const bool cg2 = (g_force_cg == 2) || (g_force_cg == 0 && M > 128 && (out_f32 || K >= 2048));
  if (swiglu_half > 0) {
    const bool cg2 = (g_force_cg == 2) || (g_force_cg == 0 && M > 128 && g_swiglu_cg == 2);
    NVIT_REQUIRE(N == swiglu_half, "nvit_gemm_bf16: swiglu needs N == F");
    NVIT_REQUIRE(!a_mn_major && !b_mn_major && !out_f32 && !accumulate && splits <= 1 && !bias && !rowadd,
                 "nvit_gemm_bf16: swiglu supports K-major operands, bf16 output and the colscale epilogue only");
    if (cg2) return launch_gemm<256, false, false, true, true>(p, A, B, lda, ldb, st);
    return launch_gemm<256, false, false, true, false>(p, A, B, lda, ldb, st);
  }
  // BN = 256 unless the problem is narrow enough that a 256-wide tile would be mostly padding.
  const bool wide = (N > 128) && ((N % 256 == 0) || (N % 256 > 128) || N >= 1024);
  const int sel = (wide ? 4 : 0) | (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);
  if (cg2 && wide) {
    switch (sel & 3) {
      case 0: return launch_gemm<256, false, false, false, true>(p, A, B, lda, ldb, st);
      case 1: return launch_gemm<256, false, true, false, true>(p, A, B, lda, ldb, st);
      case 2: return launch_gemm<256, true, false, false, true>(p, A, B, lda, ldb, st);
      default: return launch_gemm<256, true, true, false, true>(p, A, B, lda, ldb, st);
    }
  }
  switch (sel) {
    case 0: return launch_gemm<128, false, false, false, false>(p, A, B, lda, ldb, st);
    case 1: return launch_gemm<128, false, true, false, false>(p, A, B, lda, ldb, st);
    case 2: return launch_gemm<128, true, false, false, false>(p, A, B, lda, ldb, st);
    case 3: return launch_gemm<128, true, true, false, false>(p, A, B, lda, ldb, st);
    case 4: return launch_gemm<256, false, false, false, false>(p, A, B, lda, ldb, st);
    case 5: return launch_gemm<256, false, true, false, false>(p, A, B, lda, ldb, st);
    case 6: return launch_gemm<256, true, false, false, false>(p, A, B, lda, ldb, st);
    default: return launch_gemm<256, true, true, false, false>(p, A, B, lda, ldb, st);
  }
}

// C = A B^T (+bias) with the unit-norm q/k treatment of nViT applied in the epilogue (see include/nvit_b200.h).
extern "C" int nvit_gemm_qknorm(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                                int64_t ldc, const float* bias, const float* scale, float scale_mul, int64_t scale_period,
                                int64_t norm_cols, float* inv_out, int64_t ld_inv, void* stream) {
  NVIT_REQUIRE(A && B && C && scale && inv_out, "nvit_gemm_qknorm: null operand");
  NVIT_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "nvit_gemm_qknorm: bad sizes");
  NVIT_REQUIRE((N % 64) == 0 && (norm_cols % 64) == 0 && norm_cols > 0 && norm_cols <= N && scale_period > 0 && (scale_period % 64) == 0,
               "nvit_gemm_qknorm: N, norm_cols and scale_period must be multiples of the head size 64");
  NVIT_REQUIRE(ld_inv >= norm_cols / 64, "nvit_gemm_qknorm: inv_out rows too short");
  NVIT_REQUIRE(tma_addressable(C, ldc, 2), "nvit_gemm_qknorm: C must be 16-byte aligned with a 16-byte row pitch");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.C = C;
  p.ldc = ldc;
  p.bias = bias;
  p.colscale = scale;
  p.colscale_mul = scale_mul;
  p.qk_cols = (int)norm_cols;
  p.qk_period = (int)scale_period;
  p.qk_inv = inv_out;
  p.ld_inv = ld_inv;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.rowadd_period = 1;
  p.splits = 1;
  p.dbg = g_dbg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool cg2 = (g_force_cg == 2) || (g_force_cg == 0 && M > 128 && K >= 2048);
  if (cg2) return launch_gemm<256, false, false, false, true>(p, A, B, lda, ldb, st);
  return launch_gemm<256, false, false, false, false>(p, A, B, lda, ldb, st);
}

// d(uv_raw)[M, 2F] = backward of x = (u su) silu(v sv) applied to dx = dY W, dx never leaving the SM.
extern "C" int nvit_gemm_gate_bwd(const void* dY, const void* W, const void* uv_raw, const float* suv, float suv_mul, void* d_uv,
                                  int64_t M, int64_t F, int64_t K, int64_t ld_dy, int64_t ld_w, int64_t ld_uv, int64_t ld_duv,
                                  void* stream) {
  NVIT_REQUIRE(dY && W && uv_raw && d_uv, "nvit_gemm_gate_bwd: null operand");
  NVIT_REQUIRE(M > 0 && F > 0 && K > 0 && M < (1ll << 31) && F < (1ll << 30) && K < (1ll << 31), "nvit_gemm_gate_bwd: bad sizes");
  NVIT_REQUIRE((F % 64) == 0, "nvit_gemm_gate_bwd: F must be a multiple of 64 (got %lld)", (long long)F);
  NVIT_REQUIRE((ld_uv % 8) == 0 && (reinterpret_cast<uintptr_t>(uv_raw) & 15) == 0,
               "nvit_gemm_gate_bwd: uv_raw rows must be 16-byte aligned");
  NVIT_REQUIRE(tma_addressable(d_uv, ld_duv, 2), "nvit_gemm_gate_bwd: d_uv must be 16-byte aligned with a 16-byte row pitch");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.C = d_uv;
  p.ldc = ld_duv;
  p.colscale = suv;
  p.colscale_mul = suv_mul;
  p.gate_uv = static_cast<const __nv_bfloat16*>(uv_raw);
  p.ld_uv = ld_uv;
  p.M = (int)M; p.N = (int)F; p.K = (int)K;
  p.rowadd_period = 1;
  p.splits = 1;
  p.swiglu_half = (int)F;
  p.dbg = g_dbg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the heavy epilogue wants the smaller operand traffic of CTA pairs (and their deeper TMA ring)
  const bool cg2 = (g_force_cg == 2) || (g_force_cg == 0 && M > 128 && g_gateb_cg == 2);
  if (g_gateb_groups == 4) {
    if (cg2) return launch_gemm<256, false, true, false, true, true, 4>(p, dY, W, ld_dy, ld_w, st);
    return launch_gemm<256, false, true, false, false, true, 4>(p, dY, W, ld_dy, ld_w, st);
  }
  if (cg2) return launch_gemm<256, false, true, false, true, true>(p, dY, W, ld_dy, ld_w, st);
  return launch_gemm<256, false, true, false, false, true>(p, dY, W, ld_dy, ld_w, st);
}

extern "C" int nvit_gemm_swiglu_cta_group(int mode) {   // benchmarking hook: 1 or 2 (default 2)
  NVIT_REQUIRE(mode == 1 || mode == 2 || mode == 11 || mode == 12 || mode == 22 || mode == 24,
               "nvit_gemm_swiglu_cta_group: mode must be 1, 2 (forward gate GEMM), 11, 12 (gate backward) or 22, 24 (its epilogue groups)");
  if (mode > 20) g_gateb_groups = mode - 20;
  else if (mode > 10) g_gateb_cg = mode - 10;
  else g_swiglu_cg = mode;
  return NVIT_OK;
}

// Tensor-map table counters (encodes = maps built by the driver call, hits = maps taken from the table), process-wide.
extern "C" int nvit_tmap_cache_stats(int64_t* encodes, int64_t* hits) {
  if (encodes) *encodes = nvit::g_tmap_encodes.load(std::memory_order_relaxed);
  if (hits) *hits = nvit::g_tmap_hits.load(std::memory_order_relaxed);
  return NVIT_OK;
}

extern "C" int nvit_gemm_raster_group(int group) {
  NVIT_REQUIRE(group >= -1 && group <= 64, "nvit_gemm_raster_group: band width must be in [0, 64] (0 = n fastest) or -1 (automatic)");
  nvit::g_raster_group.store(group, std::memory_order_relaxed);
  return NVIT_OK;
}

#ifdef NVIT_BENCH_HOOKS
extern "C" int nvit_gemm_debug(int mode) {   // measurement aid, results are WRONG when non-zero
  g_dbg = mode;
  return NVIT_OK;
}
#endif

extern "C" int nvit_gemm_force_cta_group(int mode) {
  NVIT_REQUIRE(mode == 0 || mode == 1 || mode == 2, "nvit_gemm_force_cta_group: mode must be 0 (auto), 1 or 2");
  g_force_cg = mode;
  return NVIT_OK;
}
