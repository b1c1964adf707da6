// Unit-norm QK attention for nViT on sm_100a tensor cores (tcgen05 + TMEM), forward and backward.
//
//   qh = s * q/||q||_D,  kh = s * k/||k||_D,  s = sqk * sqk_mul            (nvit/model.py:108-119, 236-249)
//   O  = softmax(scale * qh kh^T) v,  non-causal                           (nvit/model.py:121-127, 252-258)
//
// The sequence is short and fixed by the patch grid (T = 196 for 224/16, 64 for the CIFAR config), so one CTA owns one
// (batch, head): Q, K, V (and dO, O) tiles of up to 256 tokens x 64 channels are TMA-loaded into 128B-swizzled shared
// memory, q/k arrive either raw (normalised and sqk-scaled in place on those tiles) or already normalised by the
// projection GEMM (nvit_gemm_qknorm, with 1/||x|| on the side), and every matmul is a tcgen05.mma whose operands are
// different descriptor views (K-major or MN-major) of the same shared tiles - or, for P in the forward pass, tensor memory:
//   forward : S = Qh Kh^T (M=128 q rows, N=Tpad) -> row softmax by the threads owning the TMEM lane -> P (bf16 pairs
//             written back over the dead score columns) ; O = P V (A from TMEM, B = V as it lies, MN-major).
//             256 threads, 96 KB shared memory, 256 TMEM columns: two CTAs per SM overlap each other's serial phases.
//   backward (per 128-row kv tile j, scores kept transposed so that dV and dK complete per tile):
//             S^T = Kh_j Qh^T ; P^T = exp(scale S^T - lse) ; dV_j = P^T dO ; dP^T = V_j dO^T ;
//             dS^T = P^T (dP^T - delta) scale ; dK_j = dS^T Qh ; dQ += dS Kh_j (A = dS^T viewed MN-major, hence in
//             shared memory) ; then the backward of the row normalisation and the sqk gradient.
//             512 threads (four per TMEM lane), 202 KB shared memory, 464 TMEM columns: one CTA per SM, overlapped by
//             hand (split load barriers, next score tile behind dK/dQ, dV staged under the tensor pipe).
// Outputs are staged in operand rows that are dead by then and leave as [128 x 64] TMA stores.  With unit-norm q/k the
// logits are bounded by scale * max(s^2), so the forward softmax needs no running max (one pass); the general two-pass
// form is kept for un-normalised inputs and very large learned scales.
#include "common.cuh"
#include <string.h>

namespace nvit {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box);

constexpr int ATT_THREADS = 512;   // 16 warps: four threads share a TMEM lane (a score row)
constexpr int ATT_PARTS = 4;       // ... and split its columns / channels four ways
constexpr int ATT_ROWS = 256;                   // token capacity of a shared tile
constexpr int ATT_TILE_BYTES = ATT_ROWS * 128;  // [256 tokens][64 bf16]
constexpr int ATT_PB_BYTES = 128 * 256 * 2;     // [128 rows][4 k-blocks x 64 bf16]
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// barrier parities of a role live as bits of ONE register (MEASURED: ten separate phase variables put the MMA warp's loop
// state into local memory, and with 224 KB of shared memory the L1 holds next to nothing: every such reload is an L2 round
// trip on the path that issues the products)
__device__ __forceinline__ void mbar_wait_flip(uint64_t* bar, uint32_t& phases, int bit) {
  mbar_wait(bar, (phases >> bit) & 1u);
  phases ^= 1u << bit;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MEASURED: computing delta = rowsum(dO * O) from global rows during set-up (16 scattered 16-byte loads per token instead of
// the O tile + the delta jobs) made the backward 460 -> 587 us: the set-up phase grew from 2.7 k to 19 k cycles.
// MEASURED: moving every 2nd / 3rd / 4th exponential of the softmax passes from the SFU to a degree-4 polynomial on the
// FMA pipe (the FlashAttention-4 trade) made both kernels slower (fwd 150 -> 166 / 159 / 157 us, bwd 490 -> 505 / 494 /
// 496 us): these passes are bound by issue slots and latency, not by ex2 throughput.  Round 2, persistent forward (whose
// exponential pass, 4.4 k cycles per q tile for the 53 k exponentials of the two warpgroups, sits at 75 % of the SFU rate): every
// other PAIR through a cubic in packed fp32 (FFMA2, 7.5e-5 relative error): 93.1 us against 91 us, the pass still 4.47 k cycles
// (scripts/attn_fwd_phases.py): at one half the packed-pair FMA pipe is the new limit (per 16-column chunk 40 FFMA2 / FADD2 at
// two issue cycles each + 8 MUFU + 8 F2FP + 16 integer against 16 + 16 MUFU before), and the balance point near one quarter
// would return ~0.9 k of a head's 7.1 k cycles = 0.14 ms per step before the power cap takes its share.  Not pursued.

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
  uint4 raw[8];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    raw[c] = *reinterpret_cast<const uint4*>(tile + sw128(row, c));
    float x[8];
    unpack8(raw[c], x);
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += x[e] * x[e];
  }
  const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float x[8];
    unpack8(raw[c], x);
    const float4 s0 = *reinterpret_cast<const float4*>(s_scale + 8 * c), s1 = *reinterpret_cast<const float4*>(s_scale + 8 * c + 4);
    x[0] *= inv * s0.x; x[1] *= inv * s0.y; x[2] *= inv * s0.z; x[3] *= inv * s0.w;
    x[4] *= inv * s1.x; x[5] *= inv * s1.y; x[6] *= inv * s1.z; x[7] *= inv * s1.w;
    *reinterpret_cast<uint4*>(tile + sw128(row, c)) = pack8(x);
  }
  return inv;
}

struct alignas(64) AttnParams {
  CUtensorMap tq, tk, tv, tdo, to;
  CUtensorMap tdq, tdk, tdv;   // backward outputs: [128 tokens][64] boxes, stored from the (dead) operand tiles
  const float* sqk;
  const float *inv_q, *inv_k;   // non-null: q / k arrive normalised (nvit_gemm_qknorm) with 1/||x|| at [token * ld_inv + head]
  long long ld_inv_q, ld_inv_k;
  float* lse;
  float* dsqk;
  long long ldq, ldk, ldo, lddq, lddk, lddv;
  float sqk_mul, scale;
  int B, H, T, TP, nQ, nK;
  int dbuf;         // attn_bwd_ws_kernel: two (Q, K) tile pairs in shared memory (the next head's tiles load during this one)
  unsigned h_magic; // attn_bwd_ws3_kernel: 2^32 / H rounded up: head / H = umulhi(head, h_magic) for head * H < 2^32 (a runtime division
                    // is ~40 dependent instructions, and the per-head set-up of its roles sits on the critical path)
  int* work;        // attn_bwd_ws3_kernel: {next unclaimed head - 2 gridDim.x, CTAs that have finished}; both return to 0 by themselves
  long long* dbg;   // measurement aid (nvit_attention_debug): clock64 marks of thread 0 of the first 8 CTAs, 32 slots each
};

#ifdef NVIT_BENCH_HOOKS
#define ATT_MARK(i) do { if (p.dbg && blockIdx.x < 8 && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + (i)] = clock64(); } while (0)
// v2 kernel: 64 slots per CTA (first 4 CTAs): [0,32) thread 0 of the compute warps, [32,64) lane 0 of the MMA warp
// (marks are taken during the SECOND head a CTA processes: steady state, with a head before and a head after it)
#define ATT2_MARK(i) do { if (p.dbg && n == 1 && blockIdx.x < 4 && threadIdx.x == 0) p.dbg[blockIdx.x * 64 + (i)] = clock64(); } while (0)
#define ATT2_MMARK(i) do { if (p.dbg && n == 1 && blockIdx.x < 4 && threadIdx.x == ATT2_COMPUTE) p.dbg[blockIdx.x * 64 + 32 + (i)] = clock64(); } while (0)
// v3 kernel (buffer of 16384 int64): 96 slots per CTA from [1024 + 96 cta]: [0,32) thread 0 of the compute warps, [32,64) lane 0 of the MMA warp (thread
// 384), [64,96) thread 0 of the epilogue warpgroup (thread 256); FOURTH head of every CTA (steady state)
#define ATT3_MARK(i) do { if (p.dbg && n == 3 && threadIdx.x == 0) p.dbg[1024 + blockIdx.x * 96 + (i)] = clock64(); } while (0)
#define ATT3_MMARK(i) do { if (p.dbg && n == 3 && threadIdx.x == 384) p.dbg[1024 + blockIdx.x * 96 + 32 + (i)] = clock64(); } while (0)
#define ATT3_EMARK(i) do { if (p.dbg && n == 3 && threadIdx.x == 256) p.dbg[1024 + blockIdx.x * 96 + 64 + (i)] = clock64(); } while (0)
// [512 + 2 cta + k]: globaltimer when CTA `cta` entered (k = 0) and left (k = 1) the kernel (buffer of 1024 int64)
// persistent forward: [1024 + 96 cta + 32 role + i], roles: thread 0 (softmax warpgroup 0), thread 128 (warpgroup 1), thread 256 (MMA warp); 4th head
#define ATTF_MARK(role, i) do { if (p.dbg && n == 3 && threadIdx.x == (role) * 128) p.dbg[1024 + blockIdx.x * 96 + (role) * 32 + (i)] = clock64(); } while (0)
#define ATT3_CTAMARK(k) do { if (p.dbg && threadIdx.x == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); p.dbg[512 + blockIdx.x * 2 + (k)] = gt; \
    if ((k) == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.dbg[512 + 296 + blockIdx.x] = sm; } } } while (0)
#else
#define ATT_MARK(i) do { } while (0)
#define ATT2_MARK(i) do { } while (0)
#define ATT2_MMARK(i) do { } while (0)
#define ATT3_MARK(i) do { } while (0)
#define ATT3_MMARK(i) do { } while (0)
#define ATT3_EMARK(i) do { } while (0)
#define ATT3_CTAMARK(k) do { } while (0)
#define ATTF_MARK(role, i) do { } while (0)
#endif

__host__ __device__ constexpr uint32_t IDESC_KM(int N) { return umma_idesc_bf16(128, N, 0, 1); }  // A K-major, B MN-major
__host__ __device__ constexpr uint32_t IDESC_MM(int N) { return umma_idesc_bf16(128, N, 1, 1); }  // A MN-major, B MN-major
__device__ __forceinline__ uint32_t idesc_kk_n(int N) {                                          // A, B K-major, runtime N
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

// common prologue pieces ------------------------------------------------------------------------------------------
struct AttnThread {
  int tid, warp, lane, wq, part, row, b, h;
  int c_begin, c_end, nch;  // this thread's 16-column chunks of a score row
};
template <int PARTS = ATT_PARTS>
__device__ __forceinline__ AttnThread attn_thread(const AttnParams& p) {
  AttnThread t;
  t.tid = threadIdx.x;
  t.warp = t.tid >> 5;
  t.lane = t.tid & 31;
  t.wq = t.warp & 3;       // TMEM lane quarter this warp may access
  t.part = t.warp >> 2;    // which quarter of the columns / channels of a row
  t.row = t.wq * 32 + t.lane;
  t.b = blockIdx.x / p.H;
  t.h = blockIdx.x % p.H;
  t.nch = p.TP >> 4;
  t.c_begin = (t.part * t.nch) / PARTS;
  t.c_end = ((t.part + 1) * t.nch) / PARTS;
  return t;
}


// ---- MMA issue helpers.  Called by ALL lanes of one warp so that descriptors stay in uniform registers; only the
//      tcgen05 instruction itself is predicated on an elected lane (see gemm_tcgen05.cu).
// D (+)= A * B over `nks` k-steps; A advances a_step (in 16-byte units) per k-step, B advances b_step.
__device__ __forceinline__ void mma_seq(uint32_t tmem_d, uint64_t da, uint32_t a_step, uint64_t db, uint32_t b_step, uint32_t idesc,
                                        int nks, bool accumulate_first) {
#pragma unroll 1
  for (int ks = 0; ks < nks; ++ks) {
    const uint32_t acc = (accumulate_first || ks > 0) ? 1u : 0u;
    if (elect_one()) umma_bf16_ss(tmem_d, da, db, idesc, acc);
    da += a_step;
    db += b_step;
  }
}
// Same with A = the [128 x Tpad] P / dS buffer read K-major: k-step ks sits in 16 KB k-block ks/4 at byte (ks%4)*32.
__device__ __forceinline__ void mma_seq_pk(uint32_t tmem_d, uint64_t dp, uint64_t db, uint32_t b_step, uint32_t idesc, int nks) {
#pragma unroll 1
  for (int ks = 0; ks < nks; ++ks) {
    const uint64_t da = dp + static_cast<uint64_t>((ks >> 2) * 1024 + (ks & 3) * 2);
    if (elect_one()) umma_bf16_ss(tmem_d, da, db, idesc, ks > 0 ? 1u : 0u);
    db += b_step;
  }
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  if (elect_one()) umma_commit(bar);
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ forward
// One CTA of 256 threads per (batch, head), TWO CTAs per SM: the phases of one head are a serial chain (load ->
// normalise -> S MMA -> softmax -> PV MMA -> store), so two resident CTAs keep the CUDA cores of the SM busy while the
// other one waits for the tensor pipe or for TMA.  That needs <= 113 KB of shared memory and <= 256 TMEM columns per CTA:
// the probabilities never go to shared memory - they are written back to TMEM as packed bf16 pairs over the (dead)
// score columns and feed the PV product as its A operand straight from tensor memory, and O reuses score columns too:
//   columns [0, TP)      S  = Qh Kh^T            (fp32)
//   columns [0, TP/2)    P  = exp2(...)           (bf16 pairs; written only after every thread has read its S columns)
//   columns [128, 192)   O  = P V                 (fp32; TP/2 <= 128)
constexpr int ATT_FWD_THREADS = 256;
constexpr int ATT_FWD_PARTS = 2;      // two threads share a TMEM lane (a score row) and split its columns
constexpr int ATT_FWD_MAXCH = 8;      // 16-column chunks per thread at TP = 256
constexpr uint32_t TMF_O = 128;
constexpr int ATT_FWD_SMEM = 3 * ATT_TILE_BYTES + 1024 /*align*/ + (64 + 2 * 128 + 8) * 4 + 64;

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// D[tmem] (+)= A[tmem, bf16 pairs, 8 columns per k-step] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(ATT_FWD_THREADS, 2) attn_fwd_kernel(const __grid_constant__ AttnParams p) {
  pdl_enter();   // the tile loads are the first thing this kernel does: nothing to set up ahead of the previous grid
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_TILE_BYTES;
  float* s_scale = reinterpret_cast<float*>(sV + ATT_TILE_BYTES);  // [64]
  float* s_part = s_scale + 64;                                    // [2][128] partial row sums / maxima
  float* s_misc = s_part + 256;                                    // [0] = log2-domain logit bound
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(s_misc + 8);   // q, k
  uint64_t* bar_mma = bar_tma + 1;
  uint64_t* bar_tmav = bar_mma + 1;                               // v: only the PV product waits for it
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_tmav + 1);

  const AttnThread t = attn_thread<ATT_FWD_PARTS>(p);
  const int T = p.T, TP = p.TP;
  const bool has_norm = p.sqk != nullptr;
  ATT_MARK(0);

  if (t.tid == 0) {
    // the loads go out first: barrier set-up, TMEM allocation and the sqk fetch below hide under their latency
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_tmav, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_tma, 2 * ATT_TILE_BYTES);
    tma_load_3d(&p.tq, bar_tma, sQ, t.h * 64, 0, t.b);
    tma_load_3d(&p.tk, bar_tma, sK, t.h * 64, 0, t.b);
    mbar_arrive_expect_tx(bar_tmav, ATT_TILE_BYTES);
    tma_load_3d(&p.tv, bar_tmav, sV, t.h * 64, 0, t.b);
  }
  if (t.warp == 0) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  if (t.warp == 1) {
    // per-head scale vector and the bound  |scale * qh.kh| <= scale * max_c s_c^2
    const float s0 = has_norm ? p.sqk[t.h * 64 + t.lane] * p.sqk_mul : 1.f;
    const float s1 = has_norm ? p.sqk[t.h * 64 + 32 + t.lane] * p.sqk_mul : 1.f;
    s_scale[t.lane] = s0;
    s_scale[32 + t.lane] = s1;
    const float mx = warp_max(fmaxf(s0 * s0, s1 * s1));
    if (t.lane == 0) s_misc[0] = p.scale * mx;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const float bound = s_misc[0];
  const bool two_pass = !has_norm || !(bound <= 60.f);
  ATT_MARK(1);
  mbar_wait(bar_tma, 0);
  ATT_MARK(2);

  if (has_norm && p.inv_q == nullptr) {
    for (int j = t.tid; j < 2 * T; j += ATT_FWD_THREADS) {
      if (j < T) normalize_row(sQ, j, s_scale);
      else normalize_row(sK, j - T, s_scale);
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  ATT_MARK(3);

  const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV);
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(t.wq * 32) << 16);
  const float sl2 = p.scale * LOG2E;
  uint32_t mma_phase = 0;

  for (int i = 0; i < p.nQ; ++i) {
    if (t.warp == 0) {
      tc_fence_after_sync();
      // one election per product, its k-steps issued back to back from constant offsets (MEASURED in the backward kernel: a
      // per-MMA election loop costs ~125 cycles per MMA, which here sat on the head's critical path: 4 + 13 MMAs per q tile)
      {
        const uint64_t da = umma_smem_desc(sQ_a + i * 16384, 16, 1024), db = umma_smem_desc(sK_a, 16, 1024);
        const uint32_t id = idesc_kk_n(TP);
        if (elect_one()) {
          umma_bf16_ss(tmem_base, da, db, id, 0u);
#pragma unroll
          for (int ks = 1; ks < 4; ++ks) umma_bf16_ss_acc(tmem_base, da + 2 * ks, db + 2 * ks, id);
          umma_commit(bar_mma);
        }
        __syncwarp();
      }
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after_sync();
    ATT_MARK(4 + 4 * i);

    const int qtok = i * 128 + t.row;
    float m2 = bound * LOG2E;  // log2-domain offset subtracted before exp2
    if (two_pass) {
      float mx = -INFINITY;
      for (int c = t.c_begin; c < t.c_end; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_lane + c * 16, r);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (c * 16 + e < T) mx = fmaxf(mx, __uint_as_float(r[e]));
      }
      s_part[t.part * 128 + t.row] = mx;
      __syncthreads();
      m2 = fmaxf(s_part[t.row], s_part[128 + t.row]) * sl2;
      __syncthreads();
    }
    // ---- read this thread's score columns, exponentiate, keep the bf16 pairs in registers
    uint32_t pk[ATT_FWD_MAXCH][8];
    float sum = 0.f;
#pragma unroll
    for (int cc = 0; cc < ATT_FWD_MAXCH; ++cc) {
      const int c = t.c_begin + cc;
      if (c < t.c_end) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_lane + c * 16, r);
        tmem_wait_ld();
        float pv[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pv[e] = ex2_approx(fmaf(__uint_as_float(r[e]), sl2, -m2));
        if (c == t.nch - 1) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c * 16 + e >= T) pv[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) sum += pv[e];
#pragma unroll
        for (int e = 0; e < 8; ++e) pk[cc][e] = pack_bf16(pv[2 * e], pv[2 * e + 1]);
      }
    }
    s_part[t.part * 128 + t.row] = sum;
    tc_fence_before_sync();
    __syncthreads();           // every score column has been read: P may now overwrite them
    tc_fence_after_sync();
#pragma unroll
    for (int cc = 0; cc < ATT_FWD_MAXCH; ++cc) {
      const int c = t.c_begin + cc;
      if (c < t.c_end) tmem_st_32x32b_x8(t_lane + c * 8, pk[cc]);
    }
    tmem_wait_st();
    tc_fence_before_sync();
    __syncthreads();
    ATT_MARK(5 + 4 * i);

    if (t.warp == 0) {
      if (i == 0) mbar_wait(bar_tmav, 0);
      tc_fence_after_sync();
      const uint64_t dv = umma_smem_desc(sV_a, 8192, 1024);
      const int nks = TP >> 4;
      if (elect_one()) {
        umma_bf16_ts(tmem_base + TMF_O, tmem_base, dv, IDESC_KM(64), 0u);
#pragma unroll
        for (int ks = 1; ks < 16; ++ks)
          if (ks < nks) umma_bf16_ts(tmem_base + TMF_O, tmem_base + ks * 8, dv + 128 * ks, IDESC_KM(64), 1u);
        umma_commit(bar_mma);
      }
      __syncwarp();
    }
    const float total = s_part[t.row] + s_part[128 + t.row];
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after_sync();
    ATT_MARK(6 + 4 * i);
    {
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_lane + TMF_O + t.part * 32, r);
      tmem_wait_ld();
      if (qtok < T) {
        // O leaves through the (dead) Qh rows of this tile: one TMA store per [128 x 64] tile instead of per-thread
        // 16-byte stores at a 1.5 KB row stride
        const float inv = 1.f / total;
        float o[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = __uint_as_float(r[e]) * inv;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) *reinterpret_cast<uint4*>(sQ + sw128(qtok, t.part * 4 + q4)) = pack8(o + 8 * q4);
        if (t.part == 0) p.lse[(static_cast<long long>(t.b) * p.H + t.h) * T + qtok] = (m2 + log2f(total)) * LN2;
      }
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();
    __syncthreads();
    if (t.tid == 0) {
      tma_store_3d(&p.to, sQ + i * 16384, t.h * 64, i * 128, t.b);
      bulk_commit_group();
    }
    ATT_MARK(7 + 4 * i);
  }
  if (t.tid == 0) bulk_wait_group_read<0>();   // the staged O tiles have left shared memory before the CTA retires

  if (t.warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
}


// ------------------------------------------------------------------------------------------------ forward, persistent
// attn_fwd_kernel spends a head as a serial chain (S^T product -> exponentials -> P back to tensor memory -> PV product -> O
// rows out) and gets its overlap from a second resident CTA; every head pays the set-up of a CTA (barriers, TMEM allocation,
// exposed tile loads).  Here ONE persistent CTA per SM walks over heads it claims from a device-wide counter (as
// attn_bwd_ws3_kernel does):
//   warps 0-3 / 4-7  two softmax warpgroups, one per 128-row q tile of the head, ONE thread per score row: the row sum is
//                    thread-local, and P goes back over the row's own dead score columns chunk by chunk (P chunk c lands in
//                    columns the thread has already read), so there is no barrier between reading S and writing P and no P
//                    held in registers; afterwards the warpgroup normalises its O rows into the dead Qh rows;
//   warp 8           MMA issuer: the q tiles alternate - S_0(n), O_1(n - 1), S_1(n), O_0(n) with O_r = P_r V - so the next head's S_r
//                    goes out as soon as warpgroup r has drained O_r, while the other warpgroup is still in its pass;
//   warp 9 (1 lane)  head sequence, every tile load (two (Q, K, V) buffer sets: the head after next loads while the next one
//                    is processed) and the O stores.
// Tensor memory: q tile r owns columns [256 r, 256 r + 256): S in [0, TP), P over [0, TP / 2), O in [128, 192).
// Handles pre-normalised q / k (inv_q given) and plain attention (sqk == NULL); nvit_attention_fwd keeps the round-1 kernel for
// q / k that are normalised in shared memory.
constexpr int ATTF_THREADS = 320;
__global__ void __launch_bounds__(ATTF_THREADS, 1) attn_fwd_ws_kernel(const __grid_constant__ AttnParams p) {
  pdl_enter();
  ATT3_CTAMARK(0);
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int T = p.T, TP = p.TP, nQ = p.nQ;
  const uint32_t R = static_cast<uint32_t>(TP) * 128u;       // one [TP tokens][64 bf16] tile
  // (the S^T products and the O stores address whole 128-row q tiles: for short sequences that reaches past the six tiles)
  const uint32_t tiles_end = max(6u * R, 3u * R + static_cast<uint32_t>(nQ) * 16384u);
  float* s_bound = reinterpret_cast<float*>(smem + tiles_end);   // [256] per head of a token: scale * max_c s_c^2 (bound of the logits)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bound + 256);
  uint64_t* bar_qk = bars + 0;     // [2] agent -> MMA, softmax: Q, K of buffer set b have landed (or: no further head)
  uint64_t* bar_v = bars + 2;      // [2] agent -> MMA: V of buffer set b
  uint64_t* bar_S = bars + 4;      // [2] MMA -> softmax warpgroup r
  uint64_t* bar_P = bars + 6;      // [2] softmax r -> MMA (one arrival per warp)
  uint64_t* bar_O = bars + 8;      // [2] MMA -> softmax r
  uint64_t* bar_free = bars + 10;  // [2] softmax r -> MMA: O_r has left tensor memory
  uint64_t* bar_ost = bars + 12;   // [2] softmax r -> agent: the O rows of q tile r are staged
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);
  int* const s_head = reinterpret_cast<int*>(bars + 15);      // [4] head of iteration n at [n & 3]; -1: none

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool has_norm = p.sqk != nullptr;
  const int nheads = p.B * p.H;
  auto split_head = [&](int hd, int& b, int& h) {
    b = p.h_magic ? static_cast<int>(__umulhi(static_cast<unsigned>(hd), p.h_magic)) : hd / p.H;
    h = hd - b * p.H;
  };

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1);          // qk, v
    for (int i = 4; i < 6; ++i) mbar_init(bars + i, 1);          // S
    for (int i = 6; i < 8; ++i) mbar_init(bars + i, 4);          // P
    for (int i = 8; i < 10; ++i) mbar_init(bars + i, 1);         // O
    for (int i = 10; i < 14; ++i) mbar_init(bars + i, 4);        // free, ost
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  // per-head bound of the logits: |scale * qh.kh| <= scale * max_c s_c^2 (unit-norm q / k), one warp per head
  for (int h = warp; h < p.H; h += ATTF_THREADS / 32) {
    float mx = 0.f;
    if (has_norm) {
      const float s0 = p.sqk[h * 64 + lane] * p.sqk_mul, s1 = p.sqk[h * 64 + 32 + lane] * p.sqk_mul;
      mx = warp_max(fmaxf(s0 * s0, s1 * s1));
    }
    if (lane == 0) s_bound[h] = p.scale * mx;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const float sl2 = p.scale * LOG2E;

  if (warp < 8) {
    // ===================== softmax warpgroups =====================
    const int wg = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + wg * 256;
    const int nch = TP >> 4;
    const int qtok = wg * 128 + row;
    uint32_t ph = 0;   // bits: 0/1 qk, 2 S, 3 O
    for (int n = 0;; ++n) {
      const int bs = n & 1;
      ATTF_MARK(wg, 0);
      mbar_wait_flip(bar_qk + bs, ph, bs);
      const int hd = s_head[n & 3];
      if (hd < 0) break;
      if (wg >= nQ) continue;                    // T <= 128: the second warpgroup has no q tile
      int b, h;
      split_head(hd, b, h);
      const float bound = s_bound[h];
      const bool two_pass = !has_norm || !(bound <= 60.f);
      uint8_t* const sQ = smem + bs * 3 * R;
      ATTF_MARK(wg, 1);
      mbar_wait_flip(bar_S + wg, ph, 2);
      tc_fence_after_sync();
      ATTF_MARK(wg, 2);
      float m2 = bound * LOG2E;                  // log2-domain offset subtracted before exp2
      if (two_pass) {
        float mx = -INFINITY;
        for (int c = 0; c < nch; ++c) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(t_lane + c * 16, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c * 16 + e < T) mx = fmaxf(mx, __uint_as_float(r[e]));
        }
        m2 = mx * sl2;
      }
      // ---- P = exp2(scale log2e S - m2): chunk c of P (8 columns of bf16 pairs) goes over score columns this thread has read
      // (MEASURED: fetching the chunk pair after this one from tensor memory while this one is worked on - two register sets, 167
      // registers - leaves the pass at 4.47 k cycles and the kernel at 93.1 us: the pass does not wait on tcgen05.ld latency.)
      const f32x2 sl2x2 = pack2(sl2, sl2), nm2 = pack2(-m2, -m2);
      f32x2 sum2 = pack2(0.f, 0.f);
      for (int c0 = 0; c0 < nch; c0 += 2) {
        uint32_t r[2][16];
        tmem_ld_32x32b_x16(t_lane + c0 * 16, r[0]);
        if (c0 + 1 < nch) tmem_ld_32x32b_x16(t_lane + c0 * 16 + 16, r[1]);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int c = c0 + k;
          if (c < nch) {
            uint32_t pk[8];
            if (c * 16 + 16 <= T) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float a0, a1;
                unpack2(fma2(pack2(__uint_as_float(r[k][2 * e]), __uint_as_float(r[k][2 * e + 1])), sl2x2, nm2), a0, a1);
                const f32x2 pv = pack2(ex2_approx(a0), ex2_approx(a1));
                sum2 = add2(sum2, pv);
                pk[e] = f32x2_to_bf16x2(pv);
              }
            } else {                               // the chunk that holds column T: columns >= T contribute nothing
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float a0, a1;
                unpack2(fma2(pack2(__uint_as_float(r[k][2 * e]), __uint_as_float(r[k][2 * e + 1])), sl2x2, nm2), a0, a1);
                const float p0 = c * 16 + 2 * e < T ? ex2_approx(a0) : 0.f, p1 = c * 16 + 2 * e + 1 < T ? ex2_approx(a1) : 0.f;
                const f32x2 pv = pack2(p0, p1);
                sum2 = add2(sum2, pv);
                pk[e] = f32x2_to_bf16x2(pv);
              }
            }
            tmem_st_32x32b_x8(t_lane + c * 8, pk);
          }
        }
      }
      tmem_wait_st();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_P + wg);
      ATTF_MARK(wg, 3);
      float t0, t1;
      unpack2(sum2, t0, t1);
      const float total = t0 + t1;
      // ---- O rows: normalise, stage over the (dead) Qh rows of this q tile, hand the tile to the agent's TMA store
      mbar_wait_flip(bar_O + wg, ph, 3);
      tc_fence_after_sync();
      ATTF_MARK(wg, 4);
      uint32_t o[64];
      tmem_ld_32x32b_x32(t_lane + 128, o);
      tmem_ld_32x32b_x32(t_lane + 160, o + 32);
      tmem_wait_ld();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free + wg);
      if (qtok < T) {
        const float inv = 1.f / total;
        const f32x2 inv2 = pack2(inv, inv);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            w[e] = f32x2_to_bf16x2(mul2(pack2(__uint_as_float(o[8 * c + 2 * e]), __uint_as_float(o[8 * c + 2 * e + 1])), inv2));
          *reinterpret_cast<uint4*>(sQ + sw128(qtok, c)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        p.lse[static_cast<long long>(hd) * T + qtok] = (m2 + log2f(total)) * LN2;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ost + wg);
      ATTF_MARK(wg, 5);
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    // (MEASURED: an event loop instead of this fixed order - each q tile advancing on its own barriers, watched with
    // mbarrier.test_wait so that one warpgroup never waits for the other's pass - was slower: 97 us against 91; with try_wait,
    // which may suspend on every barrier that has not fired, 167 us.)
    uint32_t ph = 0;   // bits: 0/1 qk, 2/3 v, 4/5 P, 6/7 free
    const uint32_t id_s = idesc_kk_n(TP);
    const int nks = TP >> 4;
    // Product order: S_0(n), O_1(n - 1), S_1(n), O_0(n) - the two q tiles ALTERNATE, so neither warpgroup's next S waits behind
    // the other's pass: 93.1 -> 88.8 us (MEASURED; with S_0, S_1, O_0, O_1 per head warpgroup 0 idled from its O rows until
    // warpgroup 1 had written P_1).  Warpgroup 1 settles ~1.5 k cycles behind warpgroup 0.  MEASURED on top of this order:
    // forcing a half-pass lag (warpgroup 0 signals the middle of its first pass, S_1 of the first head goes out then) with the
    // score chunks fetched a pair ahead: 93.0 us, the pass itself 4.3 k -> 4.8 k cycles - a warpgroup that has the SFU to itself
    // is no faster, so the pass is bound by the warp's own chain (scripts/probes/tmem_read_bw.cu runs the same chain at 190
    // cycles per 16 columns with one warp per scheduler, 283 with two; the kernel needs ~340).
    auto issue_s = [&](int n, int r) {
      const int bs = n & 1;
      if (n > 0) {
        mbar_wait_flip(bar_free + r, ph, 6 + r);       // the previous head's O_r has left these columns
        tc_fence_after_sync();
      }
      const uint32_t sQ_a = smem_u32(smem + bs * 3 * R), sK_a = sQ_a + R;
      const uint64_t da = umma_smem_desc(sQ_a + r * 16384, 16, 1024), db = umma_smem_desc(sK_a, 16, 1024);
      const uint32_t td = tmem_base + r * 256;
      if (elect_one()) {
        umma_bf16_ss(td, da, db, id_s, 0u);
#pragma unroll
        for (int ks = 1; ks < 4; ++ks) umma_bf16_ss_acc(td, da + 2 * ks, db + 2 * ks, id_s);
        umma_commit(bar_S + r);
      }
      __syncwarp();
    };
    auto issue_o = [&](int n, int r) {
      const int bs = n & 1;
      mbar_wait_flip(bar_P + r, ph, 4 + r);
      tc_fence_after_sync();
      const uint32_t sV_a = smem_u32(smem + bs * 3 * R) + 2 * R;
      const uint64_t dv = umma_smem_desc(sV_a, 8192, 1024);
      const uint32_t td = tmem_base + r * 256;
      if (elect_one()) {
        umma_bf16_ts(td + 128, td, dv, IDESC_KM(64), 0u);
#pragma unroll
        for (int ks = 1; ks < 16; ++ks)
          if (ks < nks) umma_bf16_ts(td + 128, td + ks * 8, dv + 128 * ks, IDESC_KM(64), 1u);
        umma_commit(bar_O + r);
      }
      __syncwarp();
    };
    for (int n = 0;; ++n) {
      const int bs = n & 1;
      ATTF_MARK(2, 0);
      mbar_wait_flip(bar_qk + bs, ph, bs);
      const bool more = s_head[n & 3] >= 0;
      tc_fence_after_sync();
      ATTF_MARK(2, 1);
      if (more) issue_s(n, 0);
      ATTF_MARK(2, 2);
      if (nQ > 1 && n > 0) issue_o(n - 1, 1);          // (V of head n - 1 landed an iteration ago)
      ATTF_MARK(2, 6);
      if (!more) break;
      if (nQ > 1) issue_s(n, 1);
      ATTF_MARK(2, 3);
      mbar_wait_flip(bar_v + bs, ph, 2 + bs);
      ATTF_MARK(2, 4);
      issue_o(n, 0);
      ATTF_MARK(2, 5);
    }
  } else if (lane == 0) {
    // ===================== agent (one thread): head sequence, tile loads, O stores =====================
    auto load_head = [&](int hd, int bs) {
      int b, h;
      split_head(hd, b, h);
      uint8_t* const sQ = smem + bs * 3 * R;
      mbar_arrive_expect_tx(bar_qk + bs, 2 * R);
      tma_load_3d(&p.tq, bar_qk + bs, sQ, h * 64, 0, b);
      tma_load_3d(&p.tk, bar_qk + bs, sQ + R, h * 64, 0, b);
      mbar_arrive_expect_tx(bar_v + bs, R);
      tma_load_3d(&p.tv, bar_v + bs, sQ + 2 * R, h * 64, 0, b);
    };
    auto no_head = [&](int bs) {       // the waiters of this buffer set find s_head = -1
      mbar_arrive(bar_qk + bs);
      mbar_arrive(bar_v + bs);
    };
    // heads 0 and 1 of the CTA are fixed, every further one is claimed from the device-wide counter two heads ahead; the
    // counter's answer is first looked at an iteration after the request (148 CTAs hit one address)
    const int id0 = static_cast<int>(blockIdx.x), id1 = static_cast<int>(blockIdx.x + gridDim.x);
    s_head[0] = id0;
    load_head(id0, 0);
    s_head[1] = id1 < nheads ? id1 : -1;
    if (id1 < nheads) load_head(id1, 1); else no_head(1);
    int claim_raw = id1 < nheads ? atomicAdd(p.work, 1) : nheads;
    uint32_t ph = 0;   // bits: 0/1 ost
    for (int n = 0;; ++n) {
      const int hd = s_head[n & 3];
      if (hd < 0) break;
      const int bs = n & 1;
      int b, h;
      split_head(hd, b, h);
      uint8_t* const sQ = smem + bs * 3 * R;
      for (int r = 0; r < nQ; ++r) {
        mbar_wait_flip(bar_ost + r, ph, r);
        tma_store_3d(&p.to, sQ + r * 16384, h * 64, r * 128, b);
      }
      bulk_commit_group();
      // buffer set bs is free (every product of head n has completed before its O rows were staged): head n + 2 goes there
      const int nxt = claim_raw + 2 * static_cast<int>(gridDim.x);
      const bool valid = nxt < nheads;
      s_head[(n + 2) & 3] = valid ? nxt : -1;
      if (valid) {
        bulk_wait_group_read<0>();           // the O stores have read the Qh rows
        load_head(nxt, bs);
        claim_raw = atomicAdd(p.work, 1);
      } else {
        no_head(bs);
        claim_raw = nheads;
      }
    }
    bulk_wait_group_read<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  ATT3_CTAMARK(1);
  if (tid == 0) {
    __threadfence();
    if (atomicInc(reinterpret_cast<unsigned*>(p.work) + 1, gridDim.x - 1) == gridDim.x - 1) {
      __threadfence();
      atomicExch(p.work, 0);
    }
  }
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int ATT_BWD_SMEM = 4 * ATT_TILE_BYTES + ATT_PB_BYTES + 1024 /*align*/ + (4 * 256 + 192 + 1024) * 4 + 64;

// Backward of y = s * x/||x|| for one row, four threads per row, each owning 16 of the 64 channels.  dL/dy sits in TMEM
// (16 columns at `taddr` for this thread); the unit vector n = x/||x|| is recovered from the normalised bf16 row still
// resident in the swizzled shared tile (n_c = y_c / s_c), so the epilogue touches no global memory except its store.
//   phase A (norm_bwd_load): g <- TMEM, n <- smem, partial dot = sum_c g_c s_c n_c, dacc += g*n
//   (the four partial dots of a row are exchanged through shared memory by the caller)
//   phase B (norm_bwd_store): dx_c = (g_c s_c - n_c dot) / ||x||
__device__ __forceinline__ float norm_bwd_load(uint32_t taddr, const uint8_t* tile, int trow, int part, const float* s_scale,
                                               const float* s_rscale, float (&g)[16], float (&n)[16], float (&dacc)[16], bool valid) {
  uint32_t r[16];
  tmem_ld_32x32b_x16(taddr, r);
  tmem_wait_ld();
  unpack8(*reinterpret_cast<const uint4*>(tile + sw128(trow, 2 * part)), n);
  unpack8(*reinterpret_cast<const uint4*>(tile + sw128(trow, 2 * part + 1)), n + 8);
  float dot = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int i = part * 16 + e;
    g[e] = __uint_as_float(r[e]);
    n[e] *= s_rscale[i];
    dacc[e] += valid ? g[e] * n[e] : 0.f;
    g[e] *= s_scale[i];
    dot += g[e] * n[e];
  }
  return dot;
}
// The backward outputs leave through shared memory: every thread writes its 16 channels (two 16-byte chunks) of its row
// into a dead 128B-swizzled operand tile and one TMA store ships the [128 x 64] tile (rows >= T are clipped by the map).
// MEASURED: per-thread 16-byte global stores at a 1.5 KB row stride (32 sectors per warp instruction) made the dQ epilogue
// 4.9 k cycles long.
__device__ __forceinline__ void norm_bwd_store(const float (&g)[16], const float (&n)[16], float dot, float inv, uint8_t* tile, int trow,
                                               int part) {
  float d[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) d[e] = (g[e] - n[e] * dot) * inv;
  *reinterpret_cast<uint4*>(tile + sw128(trow, 2 * part)) = pack8(d);
  *reinterpret_cast<uint4*>(tile + sw128(trow, 2 * part + 1)) = pack8(d + 8);
}
__device__ __forceinline__ void tmem_row16_to_tile(uint32_t taddr, uint8_t* tile, int trow, int part, bool valid = true) {
  uint32_t r[16];
  tmem_ld_32x32b_x16(taddr, r);      // warp-collective: every lane loads, `valid` only guards the stores
  tmem_wait_ld();
  if (!valid) return;
  float g[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) g[e] = __uint_as_float(r[e]);
  *reinterpret_cast<uint4*>(tile + sw128(trow, 2 * part)) = pack8(g);
  *reinterpret_cast<uint4*>(tile + sw128(trow, 2 * part + 1)) = pack8(g + 8);
}

// Sum 16 per-lane partials over the 32 lanes of a warp by recursive halving and add them to s_out[16].
__device__ __forceinline__ void reduce16_to_smem(float (&v)[16], float* s_out, int lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int off = 16 >> step, n = 8 >> step;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? v[i] : v[i + n];
      const float keep = up ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  const int ch = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  if ((lane & 1) == 0) atomicAdd(&s_out[ch], v[0]);
}

__global__ void __launch_bounds__(ATT_THREADS, 1) attn_bwd_kernel(const __grid_constant__ AttnParams p) {
  pdl_enter();
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
  float* s_rscale = s_dsqk + 64;                               // [64]  1/s (0 where s == 0)
  float* s_dot = s_rscale + 64;                                // [2][4][128] partial row dots of the normalisation backward
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(s_dot + 1024);   // q, k tiles
  uint64_t* bar_mma = bar_tma + 1;
  uint64_t* bar_tma2 = bar_mma + 1;                                // v, dO, O tiles
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_tma2 + 1);

  const AttnThread t = attn_thread<>(p);
  const int T = p.T, TP = p.TP;
  const bool has_norm = p.sqk != nullptr;
  ATT_MARK(0);

  if (t.tid == 0) {
    // the loads go out first (q, k on their own barrier: the normalisation and the first MMA need only those)
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_tma2, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_tma, 2 * ATT_TILE_BYTES);
    tma_load_3d(&p.tq, bar_tma, sQ, t.h * 64, 0, t.b);
    tma_load_3d(&p.tk, bar_tma, sK, t.h * 64, 0, t.b);
    mbar_arrive_expect_tx(bar_tma2, 3 * ATT_TILE_BYTES);
    tma_load_3d(&p.tv, bar_tma2, sV, t.h * 64, 0, t.b);
    tma_load_3d(&p.tdo, bar_tma2, sDO, t.h * 64, 0, t.b);
    tma_load_3d(&p.to, bar_tma2, sP, t.h * 64, 0, t.b);   // O parks in the (not yet used) P buffer for the delta pass
  }
  if (t.warp == 0) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (t.tid < 64) {
    const float sc = has_norm ? p.sqk[t.h * 64 + t.tid] * p.sqk_mul : 1.f;
    s_scale[t.tid] = sc;
    s_rscale[t.tid] = sc != 0.f ? 1.f / sc : 0.f;
    s_dsqk[t.tid] = 0.f;
  }
  if (t.tid >= 256) {
    const int r = t.tid - 256;  // one thread per (padded) token row
    s_lse[r] = (r < T) ? p.lse[(static_cast<long long>(t.b) * p.H + t.h) * T + r] * LOG2E : 0.f;
    // q / k normalised by the projection GEMM: only their inverse norms are needed, fetched under the tile loads
    const bool pre = p.inv_q != nullptr && r < T;
    s_invq[r] = pre ? p.inv_q[(static_cast<long long>(t.b) * T + r) * p.ld_inv_q + t.h] : 0.f;
    s_invk[r] = pre ? p.inv_k[(static_cast<long long>(t.b) * T + r) * p.ld_inv_k + t.h] : 0.f;
    s_delta[r] = 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  ATT_MARK(1);
  mbar_wait(bar_tma, 0);
  ATT_MARK(2);
  const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV), sDO_a = smem_u32(sDO), sP_a = smem_u32(sP);
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(t.wq * 32) << 16);
  const float sl2 = p.scale * LOG2E;
  const int nks_q = TP >> 4;  // k-steps over the q axis
  uint32_t mma_phase = 0;
  float dacc[16];   // dL/d(sqk) partials for this thread's 16 channels
#pragma unroll
  for (int i = 0; i < 16; ++i) dacc[i] = 0.f;
  constexpr uint32_t TM_S = 0, TM_DV = 256, TM_DK = 320, TM_DQ = 384;
  // S^T_0 = Kh_0 Qh^T: straight away when q / k need no normalisation here (already normalised by the projection GEMM, or
  // the un-normalised model), otherwise behind the normalisation jobs
  const bool qk_ready = !has_norm || p.inv_q != nullptr;
  auto issue_st0 = [&]() {
    if (t.warp == 0) {
      tc_fence_after_sync();
      mma_seq(tmem_base + TM_S, umma_smem_desc(sK_a, 16, 1024), 2, umma_smem_desc(sQ_a, 16, 1024), 2, idesc_kk_n(TP), 4, false);
      mma_commit(bar_mma);
    }
  };
  if (qk_ready) issue_st0();
  // One pool of row jobs over all threads: 2T normalisations (q rows, then k rows) followed by T delta rows
  // (delta = rowsum(dO * O) from the shared tiles; O parks in the P buffer).  With T = 196 the threads beyond the 392
  // normalisation rows start on delta at once, so the second load group is consumed as it lands.
  {
    const int base = (has_norm && p.inv_q == nullptr) ? 2 * T : 0, njobs = base + T;
    bool waited2 = false;
    for (int j = t.tid; j < njobs; j += ATT_THREADS) {
      if (j < base) {
        if (j < T) s_invq[j] = normalize_row(sQ, j, s_scale);
        else s_invk[j - T] = normalize_row(sK, j - T, s_scale);
      } else {
        if (!waited2) { mbar_wait(bar_tma2, 0); waited2 = true; }
        const int r = j - base;
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float a[8], g[8];
          unpack8(*reinterpret_cast<const uint4*>(sP + sw128(r, c)), a);
          unpack8(*reinterpret_cast<const uint4*>(sDO + sw128(r, c)), g);
#pragma unroll
          for (int e = 0; e < 8; ++e) d += a[e] * g[e];
        }
        s_delta[r] = d;
      }
    }
  }
  mbar_wait(bar_tma2, 0);      // every thread observes the second load group before it touches v / dO / the P buffer
  // No clearing of the P buffer: kv rows >= T and q columns >= T of P^T are written as zeros by the P pass, and whatever
  // else lies beyond column TP only reaches accumulator rows (q >= TP) that are never read.
  fence_proxy_async_smem();
  __syncthreads();
  ATT_MARK(3);
  if (!qk_ready) issue_st0();

  for (int j = 0; j < p.nK; ++j) {
    const int kv = j * 128 + t.row;
    const bool kv_ok = kv < T;
    if (j == 0) {       // S^T_j for j > 0 was issued behind the dK/dQ products of tile j-1 and waited for there
      mbar_wait(bar_mma, mma_phase);
      mma_phase ^= 1;
      tc_fence_after_sync();
    }
    ATT_MARK(4 + 8 * j);
    // ---- P^T = exp(scale S^T - lse)   (four threads per kv row, columns = q)
    for (int c = t.c_begin; c < t.c_end; ++c) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_lane + TM_S + c * 16, r);
      tmem_wait_ld();
      float pv[16];
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        const float4 l4 = *reinterpret_cast<const float4*>(s_lse + c * 16 + 4 * e4);
        pv[4 * e4 + 0] = ex2_approx(fmaf(__uint_as_float(r[4 * e4 + 0]), sl2, -l4.x));
        pv[4 * e4 + 1] = ex2_approx(fmaf(__uint_as_float(r[4 * e4 + 1]), sl2, -l4.y));
        pv[4 * e4 + 2] = ex2_approx(fmaf(__uint_as_float(r[4 * e4 + 2]), sl2, -l4.z));
        pv[4 * e4 + 3] = ex2_approx(fmaf(__uint_as_float(r[4 * e4 + 3]), sl2, -l4.w));
      }
      if (!kv_ok) {
#pragma unroll
        for (int e = 0; e < 16; ++e) pv[e] = 0.f;
      } else if (c == t.nch - 1) {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (c * 16 + e >= T) pv[e] = 0.f;
      }
      uint8_t* blk = sP + (c >> 2) * 16384;
      const int ch = (c & 3) * 2;
      *reinterpret_cast<uint4*>(blk + sw128(t.row, ch)) = pack8(pv);
      *reinterpret_cast<uint4*>(blk + sw128(t.row, ch + 1)) = pack8(pv + 8);
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();
    __syncthreads();
    ATT_MARK(5 + 8 * j);
    // ---- dV_j = P^T dO ; dP^T_j = V_j dO^T
    if (t.warp == 0) {
      tc_fence_after_sync();
      mma_seq_pk(tmem_base + TM_DV, umma_smem_desc(sP_a, 16, 1024), umma_smem_desc(sDO_a, 8192, 1024), 128, IDESC_KM(64), nks_q);
      mma_seq(tmem_base + TM_S, umma_smem_desc(sV_a + j * 16384, 16, 1024), 2, umma_smem_desc(sDO_a, 16, 1024), 2, idesc_kk_n(TP), 4, false);
      mma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after_sync();
    ATT_MARK(6 + 8 * j);
    // ---- dS^T = P^T (dP^T - delta) scale, in place over P^T
    for (int c = t.c_begin; c < t.c_end; ++c) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_lane + TM_S + c * 16, r);
      tmem_wait_ld();
      uint8_t* blk = sP + (c >> 2) * 16384;
      const int ch = (c & 3) * 2;
      float pv[16];
      unpack8(*reinterpret_cast<const uint4*>(blk + sw128(t.row, ch)), pv);
      unpack8(*reinterpret_cast<const uint4*>(blk + sw128(t.row, ch + 1)), pv + 8);
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        const float4 d4 = *reinterpret_cast<const float4*>(s_delta + c * 16 + 4 * e4);
        pv[4 * e4 + 0] *= (__uint_as_float(r[4 * e4 + 0]) - d4.x) * p.scale;
        pv[4 * e4 + 1] *= (__uint_as_float(r[4 * e4 + 1]) - d4.y) * p.scale;
        pv[4 * e4 + 2] *= (__uint_as_float(r[4 * e4 + 2]) - d4.z) * p.scale;
        pv[4 * e4 + 3] *= (__uint_as_float(r[4 * e4 + 3]) - d4.w) * p.scale;
      }
      *reinterpret_cast<uint4*>(blk + sw128(t.row, ch)) = pack8(pv);
      *reinterpret_cast<uint4*>(blk + sw128(t.row, ch + 1)) = pack8(pv + 8);
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();
    __syncthreads();
    ATT_MARK(7 + 8 * j);
    // ---- dK_j = dS^T Qh ; dQ_m += dS_j Kh_j
    if (t.warp == 0) {
      tc_fence_after_sync();
      mma_seq_pk(tmem_base + TM_DK, umma_smem_desc(sP_a, 16, 1024), umma_smem_desc(sQ_a, 8192, 1024), 128, IDESC_KM(64), nks_q);
      const int kv_steps = min(8, (TP - j * 128) >> 4);
      for (int m = 0; m < p.nQ; ++m)
        mma_seq(tmem_base + TM_DQ + 64 * m, umma_smem_desc(sP_a + 2 * m * 16384, 16384, 1024), 128,
                umma_smem_desc(sK_a + j * 16384, 8192, 1024), 128, IDESC_MM(64), kv_steps, j > 0);
      if (j + 1 < p.nK)   // the score tile of the next kv block rides behind: its region (dP^T) has just been consumed
        mma_seq(tmem_base + TM_S, umma_smem_desc(sK_a + (j + 1) * 16384, 16, 1024), 2, umma_smem_desc(sQ_a, 16, 1024), 2, idesc_kk_n(TP), 4, false);
      mma_commit(bar_mma);
    }
    // ---- dV_j and dK_j rows: every thread takes 16 channels of its row of both.  dV_j is complete already and is staged
    // (in the V_j rows, dead since dP^T) while the tensor pipe works on dK/dQ; dK_j replaces the K_j rows in place.
    {
      tmem_row16_to_tile(t_lane + TM_DV + t.part * 16, sV + j * 16384, t.row, t.part);
      mbar_wait(bar_mma, mma_phase);
      mma_phase ^= 1;
      tc_fence_after_sync();
      ATT_MARK(8 + 8 * j);
      if (has_norm) {
        float g[16], n[16];
        s_dot[t.part * 128 + t.row] = norm_bwd_load(t_lane + TM_DK + t.part * 16, sK, kv_ok ? kv : 0, t.part, s_scale, s_rscale, g, n, dacc, kv_ok);
        __syncthreads();
        const float dot = (s_dot[t.row] + s_dot[128 + t.row]) + (s_dot[256 + t.row] + s_dot[384 + t.row]);
        if (kv_ok) norm_bwd_store(g, n, dot, s_invk[kv], sK, kv, t.part);
      } else {
        tmem_row16_to_tile(t_lane + TM_DK + t.part * 16, sK + j * 16384, t.row, t.part);
      }
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();
    __syncthreads();
    if (t.tid == 0) {
      tma_store_3d(&p.tdv, sV + j * 16384, t.h * 64, j * 128, t.b);
      tma_store_3d(&p.tdk, sK + j * 16384, t.h * 64, j * 128, t.b);
      bulk_commit_group();
    }
    ATT_MARK(9 + 8 * j);
  }

  // ---- dQ rows: both 128-row q tiles in one pass (loads and partial dots for both, ONE exchange), staged in place over Qh
  if (has_norm) {
    float g[2][16], n[2][16];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      if (m < p.nQ) {
        const int qi = m * 128 + t.row;
        const bool ok = qi < T;
        s_dot[m * 512 + t.part * 128 + t.row] =
            norm_bwd_load(t_lane + TM_DQ + 64 * m + t.part * 16, sQ, ok ? qi : 0, t.part, s_scale, s_rscale, g[m], n[m], dacc, ok);
      }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      if (m < p.nQ) {
        const int qi = m * 128 + t.row;
        if (qi < T) {
          const float* sd = s_dot + m * 512;
          const float dot = (sd[t.row] + sd[128 + t.row]) + (sd[256 + t.row] + sd[384 + t.row]);
          norm_bwd_store(g[m], n[m], dot, s_invq[qi], sQ, qi, t.part);
        }
      }
    }
  } else {
    for (int m = 0; m < p.nQ; ++m) tmem_row16_to_tile(t_lane + TM_DQ + 64 * m + t.part * 16, sQ + m * 16384, t.row, t.part);
  }
  fence_proxy_async_smem();
  ATT_MARK(24);
  if (has_norm) reduce16_to_smem(dacc, s_dsqk + t.part * 16, t.lane);
  tc_fence_before_sync();
  __syncthreads();
  if (t.tid == 0) {
    for (int m = 0; m < p.nQ; ++m) tma_store_3d(&p.tdq, sQ + m * 16384, t.h * 64, m * 128, t.b);
    bulk_commit_group();
    bulk_wait_group_read<0>();      // all staged tiles have left shared memory before the CTA retires
  }
  if (has_norm && t.tid < 64) atomicAdd(p.dsqk + t.h * 64 + t.tid, s_dsqk[t.tid] * p.sqk_mul);
  ATT_MARK(25);
  if (t.warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// Two-phase form of norm_bwd_load / norm_bwd_store for the persistent kernel (32 channels per thread there: holding g, n
// and the sqk partials of both 16-channel halves across the row-dot exchange spilled): phase 1 returns this thread's partial
// dot and folds g * n into the shared dL/d(sqk) accumulator at once; phase 2 re-reads the 16 accumulator columns and the
// normalised row (both on chip) and writes dx.
__device__ __forceinline__ float norm_bwd_dot16(uint32_t taddr, const uint8_t* tile, int trow, int part16, const float* s_scale,
                                                const float* s_rscale, float* s_dsqk, int lane, bool valid) {
  uint32_t r[16];
  tmem_ld_32x32b_x16(taddr, r);
  tmem_wait_ld();
  float n[16], acc[16];
  unpack8(*reinterpret_cast<const uint4*>(tile + sw128(trow, 2 * part16)), n);
  unpack8(*reinterpret_cast<const uint4*>(tile + sw128(trow, 2 * part16 + 1)), n + 8);
  float dot = 0.f;
  float sc[16], rs[16];
#pragma unroll
  for (int e4 = 0; e4 < 4; ++e4) {
    *reinterpret_cast<float4*>(sc + 4 * e4) = *reinterpret_cast<const float4*>(s_scale + part16 * 16 + 4 * e4);
    *reinterpret_cast<float4*>(rs + 4 * e4) = *reinterpret_cast<const float4*>(s_rscale + part16 * 16 + 4 * e4);
  }
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float g = __uint_as_float(r[e]);
    const float nn = n[e] * rs[e];
    acc[e] = valid ? g * nn : 0.f;
    dot += g * sc[e] * nn;
  }
  reduce16_to_smem(acc, s_dsqk + part16 * 16, lane);
  return dot;
}
// (warp-collective: the TMEM load is .sync.aligned, so EVERY lane calls this; `valid` only guards the stores)
__device__ __forceinline__ void norm_bwd_apply16(uint32_t taddr, uint8_t* tile, int trow, int part16, const float* s_scale,
                                                 const float* s_rscale, float dot, float inv, bool valid) {
  uint32_t r[16];
  tmem_ld_32x32b_x16(taddr, r);
  tmem_wait_ld();
  if (!valid) return;
  float n[16], d[16], sc[16], rs[16];
  unpack8(*reinterpret_cast<const uint4*>(tile + sw128(trow, 2 * part16)), n);
  unpack8(*reinterpret_cast<const uint4*>(tile + sw128(trow, 2 * part16 + 1)), n + 8);
#pragma unroll
  for (int e4 = 0; e4 < 4; ++e4) {
    *reinterpret_cast<float4*>(sc + 4 * e4) = *reinterpret_cast<const float4*>(s_scale + part16 * 16 + 4 * e4);
    *reinterpret_cast<float4*>(rs + 4 * e4) = *reinterpret_cast<const float4*>(s_rscale + part16 * 16 + 4 * e4);
  }
#pragma unroll
  for (int e = 0; e < 16; ++e) d[e] = (__uint_as_float(r[e]) * sc[e] - n[e] * rs[e] * dot) * inv;
  *reinterpret_cast<uint4*>(tile + sw128(trow, 2 * part16)) = pack8(d);
  *reinterpret_cast<uint4*>(tile + sw128(trow, 2 * part16 + 1)) = pack8(d + 8);
}

// ------------------------------------------------------------------------------------------------ backward, v2
// Persistent, warp-specialised form of the kernel above (same arithmetic, same 128-row tiles): one CTA per SM walks over
// (batch, head) pairs; 8 compute warps (two threads per TMEM lane) + ONE dedicated MMA warp.  The work of a head is cut into
// items (kv tile j, q tile c) whose matmuls run on the tensor pipe WHILE the compute warps are in the exponential / dS passes
// of the neighbouring items, and the next head's Q / K tiles are loaded while the current head is being processed:
//   MMA warp, per item:   wait P^T(i) -> dV_j += P^T dO_c ; S^T(i+1) = K_j' Q_c'^T            (behind dV, same TMEM columns)
//                         wait dS^T(i) -> dK_j += dS^T Q_c ; dQ_c += dS K_j ; dP^T(i+1) = V_j' dO_c'^T
//   compute warps:        wait S^T(i) -> P^T = exp2(...) -> TMEM (bf16 pairs over the dead score columns: the A operand of
//                         dV comes from tensor memory, as in the forward kernel) -> wait dP^T(i) -> dS^T -> shared
//                         (double-buffered, 2 x 32 KB) -> per kv tile the dV_j / dK_j rows, at the end the dQ rows.
// MEASURED on v1 (scripts/attn_phases.py): of a head's 37 k cycles, 17 k were waits for products issued by a warp that
// also ran the passes (S^T and dP^T shared 256 TMEM columns, so each product sat between two passes) and 9 k were the
// set-up of a fresh CTA (barriers, TMEM allocation, the wait for the first tiles) that nothing could overlap at one
// CTA per SM.  MEASURED on a first warp-specialised version with 16 compute warps (17 warps -> 96 registers per thread):
// the spills it caused went to L2 (L1 is 28 KB beside 200 KB of shared memory) and made it 1.5x SLOWER than v1; with 8
// compute warps each thread takes four 16-column chunks of a row at up to 168 registers and nothing spills.
//   TMEM: S^T / P^T [0,128)  dP^T [128,256)  dV_j [256,320)  dK_j [320,384)  dQ [384,512)
//   smem: (Q, K) x 2 when they fit (T <= 208: the other pair receives the next head) + V + dO tiles of exactly TP rows,
//         two dS^T buffers (O parks in the second one until delta = rowsum(dO * O) has been taken), small per-head arrays.
constexpr int ATT_DS_BYTES = 2 * 16384;            // one dS^T buffer: [128 kv rows][2 k-blocks x 64 q columns]
constexpr int ATT2_MAX_SMEM = 232448;              // 227 KB

// per-head arrays (floats): lse[AL] invq[AL] invk[AL] delta[AL] scale[64] rscale[64] dsqk[64] dot[512]; then the barriers
__host__ __device__ constexpr int att2_array_bytes(int AL) { return (4 * AL + 3 * 64 + 256) * 4 + 128; }

// MEASURED: the 16-compute-warp instantiation (four threads per lane) needs 17 warps -> 96 registers per thread (warps are
// allocated in groups of four: 20 x 32 x 96) and spills ~900 bytes.  Re-splitting the CTA's own allocation with setmaxnreg
// (640-thread CTA; 512 x 112 + 128 x 32 = 640 x 96 - a first attempt with 112 / 56 asked for more than the CTA owns and
// blocked forever) runs, and ptxas does use R0..R109 in the compute branch, but 441 spill instructions remain there and 56
// in the 32-register MMA branch: 689 us against 386 us for 8 compute warps at 168 registers.  Only ATT2_CW = 8 is built.
template <int ATT2_CW>
__global__ void __launch_bounds__(ATT2_CW * 32 + 32, 1) attn_bwd_ws_kernel(const __grid_constant__ AttnParams p) {
  constexpr int ATT2_COMPUTE = ATT2_CW * 32;
  constexpr int ATT2_THREADS = ATT2_COMPUTE + 32;    // + the MMA warp
  constexpr int PARTS = ATT2_CW / 4;                 // threads per TMEM lane
  constexpr int NCC = 8 / PARTS;                     // 16-column chunks per thread and item, at most
  constexpr int NH2 = 4 / PARTS;                     // 16-channel groups per thread in the epilogues
  pdl_enter();
  ATT3_CTAMARK(0);
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();         // 128B-swizzled tiles need a 1024-byte aligned base
  const int T = p.T, TP = p.TP;
  const uint32_t R = static_cast<uint32_t>(TP) * 128u;  // bytes of one [TP tokens][64 bf16] tile
  const int nbuf = p.dbuf ? 2 : 1;
  uint8_t* sQK = smem;                                  // [nbuf][Q | K]
  uint8_t* sV = sQK + nbuf * 2 * R;
  uint8_t* sDO = sV + R;
  uint8_t* sDS = sDO + R;                               // two dS^T buffers
  float* s_lse = reinterpret_cast<float*>(sDS + 2 * ATT_DS_BYTES);
  float* s_invq = s_lse + TP;
  float* s_invk = s_invq + TP;
  float* s_delta = s_invk + TP;
  float* s_scale = s_delta + TP;                        // [64]
  float* s_rscale = s_scale + 64;                       // [64]  1/s (0 where s == 0)
  float* s_dsqk = s_rscale + 64;                        // [64]
  float* s_dot = s_dsqk + 64;                           // [PARTS][128] partial row dots of the normalisation backward
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dot + PARTS * 128);
  uint64_t* bar_qk = bars + 0;      // [2] TMA: q, k tiles of buffer pair b
  uint64_t* bar_vdo = bars + 2;     // TMA: v, dO, O tiles
  uint64_t* bar_S = bars + 3;       // MMA -> compute: S^T(i) complete
  uint64_t* bar_dP = bars + 4;      // MMA -> compute: dP^T(i) complete
  uint64_t* bar_P = bars + 5;       // compute -> MMA: P^T(i) is in tensor memory (one arrival per compute warp)
  uint64_t* bar_dS = bars + 6;      // compute -> MMA: dS^T(i) is in shared memory
  uint64_t* bar_free = bars + 7;    // [2] MMA -> compute: the products reading dS^T buffer b have completed
  uint64_t* bar_acc = bars + 9;     // MMA -> compute: dV_j, dK_j (and every dQ contribution so far) complete
  uint64_t* bar_norm = bars + 10;   // compute -> MMA: q / k normalised in place (only when they arrive raw)
  uint64_t* bar_dv = bars + 11;     // MMA -> compute: dV_j complete (its rows are staged while dK_j / dQ still run)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool has_norm = p.sqk != nullptr;
  const bool qk_ready = !has_norm || p.inv_q != nullptr;
  const int nQ = p.nQ, nK = p.nK, nItems = nQ * nK;
  const int nheads = p.B * p.H;
  constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

  if (tid == 0) {
    mbar_init(bar_qk + 0, 1);
    mbar_init(bar_qk + 1, 1);
    mbar_init(bar_vdo, 1);
    mbar_init(bar_S, 1);
    mbar_init(bar_dP, 1);
    mbar_init(bar_P, ATT2_CW);
    mbar_init(bar_dS, ATT2_CW);
    mbar_init(bar_free + 0, 1);
    mbar_init(bar_free + 1, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_norm, ATT2_CW);
    mbar_init(bar_dv, 1);
    fence_barrier_init();
    // the first head's tiles go out at once (q, k on their own barrier: the first score product needs only those)
    const int hd = blockIdx.x, b0 = hd / p.H, h0 = hd % p.H;
    mbar_arrive_expect_tx(bar_qk, 2 * R);
    tma_load_3d(&p.tq, bar_qk, sQK, h0 * 64, 0, b0);
    tma_load_3d(&p.tk, bar_qk, sQK + R, h0 * 64, 0, b0);
    mbar_arrive_expect_tx(bar_vdo, 3 * R);
    tma_load_3d(&p.tv, bar_vdo, sV, h0 * 64, 0, b0);
    tma_load_3d(&p.tdo, bar_vdo, sDO, h0 * 64, 0, b0);
    tma_load_3d(&p.to, bar_vdo, sDS + ATT_DS_BYTES, h0 * 64, 0, b0);   // O parks in dS^T buffer 1 for the delta pass
  }
  if (warp == ATT2_CW) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  for (int r = tid; r < TP; r += ATT2_THREADS) {   // rows >= T are never written again and must stay finite
    s_lse[r] = 0.f;
    s_invq[r] = 0.f;
    s_invk[r] = 0.f;
    s_delta[r] = 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t sV_a = smem_u32(sV), sDO_a = smem_u32(sDO), sDS_a = smem_u32(sDS);

  if (warp == ATT2_CW) {
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues (see gemm_tcgen05.cu) =====
    uint32_t ph_qk0 = 0, ph_qk1 = 0, ph_vdo = 0, ph_P = 0, ph_dS = 0, ph_norm = 0;
    auto ncol_of = [&](int c) { return min(128, TP - 128 * c); };
    int n = 0;
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x, ++n) {
      const int pb = p.dbuf ? (n & 1) : 0;
      const uint32_t sQ_a = smem_u32(sQK + pb * 2 * R), sK_a = sQ_a + R;
      // MEASURED (first persistent version, marks of scripts/attn_bwd_phases_v2.py): issuing through mma_seq (one election
      // and ~19 dependent instructions per MMA) cost ~125 cycles per MMA - the 32 MMAs of an item took 4 k cycles to ISSUE
      // against ~1.5 k on the tensor pipe, and the compute warps waited for dP^T / the accumulators.  Here one lane is elected
      // per product and issues its k-steps back to back from constant offsets.
      auto issue_S = [&](int j, int c) {       // S^T(j,c) = Kh_j Qh_c^T  [128 kv x ncol]
        const uint64_t da = umma_smem_desc(sK_a + j * 16384, 16, 1024), db = umma_smem_desc(sQ_a + c * 16384, 16, 1024);
        const uint32_t id = idesc_kk_n(ncol_of(c));
        if (elect_one()) {
          umma_bf16_ss(tmem_base + TM_S, da, db, id, 0u);
#pragma unroll
          for (int ks = 1; ks < 4; ++ks) umma_bf16_ss_acc(tmem_base + TM_S, da + 2 * ks, db + 2 * ks, id);
          umma_commit(bar_S);
        }
        __syncwarp();
      };
      auto issue_dP = [&](int j, int c) {      // dP^T(j,c) = V_j dO_c^T
        const uint64_t da = umma_smem_desc(sV_a + j * 16384, 16, 1024), db = umma_smem_desc(sDO_a + c * 16384, 16, 1024);
        const uint32_t id = idesc_kk_n(ncol_of(c));
        if (elect_one()) {
          umma_bf16_ss(tmem_base + TM_DP, da, db, id, 0u);
#pragma unroll
          for (int ks = 1; ks < 4; ++ks) umma_bf16_ss_acc(tmem_base + TM_DP, da + 2 * ks, db + 2 * ks, id);
          umma_commit(bar_dP);
        }
        __syncwarp();
      };
      ATT2_MMARK(0);
      if (pb == 0) { mbar_wait(bar_qk, ph_qk0); ph_qk0 ^= 1; }
      else { mbar_wait(bar_qk + 1, ph_qk1); ph_qk1 ^= 1; }
      if (!qk_ready) { mbar_wait(bar_norm, ph_norm); ph_norm ^= 1; }
      tc_fence_after_sync();
      ATT2_MMARK(1);
      issue_S(0, 0);
      mbar_wait(bar_vdo, ph_vdo);
      ph_vdo ^= 1;
      tc_fence_after_sync();
      issue_dP(0, 0);
      ATT2_MMARK(2);
      int i = 0;
      for (int j = 0; j < nK; ++j) {
        const int kv_steps = min(8, (TP - j * 128) >> 4);
        for (int c = 0; c < nQ; ++c, ++i) {
          const int nks = ncol_of(c) >> 4;
          const int jn = (c + 1 < nQ) ? j : j + 1, cn = (c + 1 < nQ) ? c + 1 : 0;
          const bool has_next = i + 1 < nItems;
          const uint32_t buf_a = sDS_a + (i & 1) * ATT_DS_BYTES;
          // ---- dV_j += P^T(i) dO_c   (A = P^T from tensor memory: 8 columns per k-step)
          mbar_wait(bar_P, ph_P);
          ph_P ^= 1;
          tc_fence_after_sync();
          ATT2_MMARK(3 + 4 * i);
          {
            const uint64_t dd = umma_smem_desc(sDO_a + c * 16384, 8192, 1024);
            const uint32_t first = c > 0 ? 1u : 0u;
            if (elect_one()) {
              umma_bf16_ts(tmem_base + TM_DV, tmem_base + TM_S, dd, IDESC_KM(64), first);
#pragma unroll
              for (int ks = 1; ks < 8; ++ks)
                if (ks < nks) umma_bf16_ts(tmem_base + TM_DV, tmem_base + TM_S + ks * 8, dd + 128 * ks, IDESC_KM(64), 1u);
              if (c == nQ - 1) umma_commit(bar_dv);
            }
            __syncwarp();
          }
          // the next score tile rides behind dV (the tensor pipe runs in issue order: it overwrites P^T only after dV read it)
          if (has_next) issue_S(jn, cn);
          ATT2_MMARK(4 + 4 * i);
          mbar_wait(bar_dS, ph_dS);
          ph_dS ^= 1;
          tc_fence_after_sync();
          ATT2_MMARK(5 + 4 * i);
          // dP^T of the next item first, unless this item completes a kv tile (then the compute warps wait for dV_j / dK_j
          // next and have a whole epilogue and a P pass before they need dP^T).  Its TMEM columns are free: every compute
          // warp has read dP^T(i) before it announced dS^T(i).
          if (has_next && c != nQ - 1) issue_dP(jn, cn);
          // ---- dK_j += dS^T(i) Qh_c ; dQ_c += dS(i) Kh_j
          {
            const uint64_t dp = umma_smem_desc(buf_a, 16, 1024);
            const uint64_t dq = umma_smem_desc(sQ_a + c * 16384, 8192, 1024);
            const uint64_t dsm = umma_smem_desc(buf_a, 16384, 1024), dkm = umma_smem_desc(sK_a + j * 16384, 8192, 1024);
            const uint32_t first_k = c > 0 ? 1u : 0u, first_q = j > 0 ? 1u : 0u;
            if (elect_one()) {
              umma_bf16_ss(tmem_base + TM_DK, dp, dq, IDESC_KM(64), first_k);
#pragma unroll
              for (int ks = 1; ks < 8; ++ks)
                if (ks < nks) umma_bf16_ss_acc(tmem_base + TM_DK, dp + static_cast<uint64_t>((ks >> 2) * 1024 + (ks & 3) * 2), dq + 128 * ks, IDESC_KM(64));
              umma_bf16_ss(tmem_base + TM_DQ + 64 * c, dsm, dkm, IDESC_MM(64), first_q);
#pragma unroll
              for (int ks = 1; ks < 8; ++ks)
                if (ks < kv_steps) umma_bf16_ss_acc(tmem_base + TM_DQ + 64 * c, dsm + 128 * ks, dkm + 128 * ks, IDESC_MM(64));
              umma_commit(bar_free + (i & 1));
              if (c == nQ - 1) umma_commit(bar_acc);
            }
            __syncwarp();
          }
          if (has_next && c == nQ - 1) issue_dP(jn, cn);
          ATT2_MMARK(6 + 4 * i);
        }
      }
    }
  } else {
    // ===================== compute warps =====================
    const int wq = warp & 3;       // TMEM lane quarter this warp may access
    const int part = warp >> 2;    // which share of the columns / channels of a row
    const int row = wq * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const f32x2 sl2x2 = pack2(p.scale * LOG2E, p.scale * LOG2E), scx2 = pack2(p.scale, p.scale);
    uint32_t ph_qk0 = 0, ph_qk1 = 0, ph_vdo = 0, ph_S = 0, ph_dP = 0, ph_acc = 0, ph_free0 = 0, ph_free1 = 0, ph_dv = 0;
    bool used0 = false, used1 = false;     // dS^T buffer b has been filled before
    // per-head scalars of the NEXT head travel in registers from the start of a head to its end.
    // MEASURED (same box, scripts/attn_bwd_time.py, v1 = 455 us as the reference point): this form 386-388 us.  The phase
    // marks show a 2.9 k-cycle stall at the start of each head (the fetched values are spilled, and the spill store waits
    // for the load), yet every attempt to remove it was slower: 4-byte cp.async copies into shared memory (lse
    // double-buffered) 408-412 us; the same plus taking dL/d(sqk) from the dQ rows only (it equals the dK-row sum, both are
    // sum dS[q,r] Qh[q,c] Kh[r,c]) 394-396 us; that identity alone 408 us; fetching right before the dQ epilogue 420 us.
    // At 168 registers and 2 warps per scheduler the kernel is latency-bound (ncu: issue slots 37 % busy, 0.5 eligible
    // warps per scheduler) and reacts to any change of the instruction schedule by +-5 %.
    float nx_lse = 0.f, nx_invq = 0.f, nx_invk = 0.f, nx_sc = 1.f;
    auto fetch_head = [&](int hd) {
      const int b = hd / p.H, h = hd % p.H;
      if (tid < T) {
        nx_lse = p.lse[(static_cast<long long>(b) * p.H + h) * T + tid];      // raw: any use here would wait for the load
        if (p.inv_q != nullptr) {
          nx_invq = p.inv_q[(static_cast<long long>(b) * T + tid) * p.ld_inv_q + h];
          nx_invk = p.inv_k[(static_cast<long long>(b) * T + tid) * p.ld_inv_k + h];
        }
      }
      if (tid < 64 && has_norm) nx_sc = p.sqk[h * 64 + tid];
    };
    auto publish_head = [&]() {
      if (tid < T) {
        s_lse[tid] = -nx_lse * LOG2E;       // stored negated: the P pass adds it inside one FFMA2
        if (p.inv_q != nullptr) { s_invq[tid] = nx_invq; s_invk[tid] = nx_invk; }
      }
      if (tid < 64) {
        const float sc = has_norm ? nx_sc * p.sqk_mul : 1.f;
        s_scale[tid] = sc;
        s_rscale[tid] = sc != 0.f ? 1.f / sc : 0.f;
        s_dsqk[tid] = 0.f;
      }
    };
    fetch_head(blockIdx.x);
    publish_head();
    named_bar_sync(1, ATT2_COMPUTE);

    int n = 0;
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x, ++n) {
      ATT2_MARK(0);
      const int b = hd / p.H, h = hd % p.H;
      const int pb = p.dbuf ? (n & 1) : 0;
      uint8_t* const sQ = sQK + pb * 2 * R;
      uint8_t* const sK = sQ + R;
      const int hd_next = hd + gridDim.x;
      const bool more = hd_next < nheads;
      if (more) fetch_head(hd_next);
      if (!qk_ready) {
        if (pb == 0) { mbar_wait(bar_qk, ph_qk0); ph_qk0 ^= 1; }
        else { mbar_wait(bar_qk + 1, ph_qk1); ph_qk1 ^= 1; }
        for (int r = tid; r < 2 * T; r += ATT2_COMPUTE) {
          if (r < T) s_invq[r] = normalize_row(sQ, r, s_scale);
          else s_invk[r - T] = normalize_row(sK, r - T, s_scale);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_norm);
      }
      int i = 0;
      for (int j = 0; j < nK; ++j) {
        const int kv = j * 128 + row;
        const bool kv_ok = kv < T;
        for (int c = 0; c < nQ; ++c, ++i) {
          const int nch = min(128, TP - 128 * c) >> 4;          // 16-column chunks of this item (<= 8)
          const int c_begin = (part * nch) / PARTS, c_end = ((part + 1) * nch) / PARTS;   // this thread's chunks: at most NCC
          const int q0 = c * 128;
          const int bsel = i & 1;
          uint8_t* const buf = sDS + bsel * ATT_DS_BYTES;
          uint32_t pk[NCC][8];
          // ---- P^T = exp2(scale log2e S^T - lse log2e), kept as bf16 pairs
          ATT2_MARK(2 + 6 * i);
          mbar_wait(bar_S, ph_S);
          ph_S ^= 1;
          tc_fence_after_sync();
          ATT2_MARK(3 + 6 * i);
#pragma unroll
          for (int pr = 0; pr < NCC / 2; ++pr) {       // two chunks per round: 32 accumulator columns in flight per thread
            uint32_t r[2][16];
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2)
              if (c_begin + 2 * pr + c2 < c_end) tmem_ld_32x32b_x16(t_lane + TM_S + (c_begin + 2 * pr + c2) * 16, r[c2]);
            tmem_wait_ld();
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
              const int cc = 2 * pr + c2;
              const int ch = c_begin + cc;
              if (ch < c_end) {
                float pv[16];
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {     // exponent = s * scale log2e + (-lse log2e), two columns per FFMA2
                  const float4 l4 = *reinterpret_cast<const float4*>(s_lse + q0 + ch * 16 + 4 * e4);
                  float a0, a1, a2, a3;
                  unpack2(fma2(pack2(__uint_as_float(r[c2][4 * e4 + 0]), __uint_as_float(r[c2][4 * e4 + 1])), sl2x2, pack2(l4.x, l4.y)), a0, a1);
                  unpack2(fma2(pack2(__uint_as_float(r[c2][4 * e4 + 2]), __uint_as_float(r[c2][4 * e4 + 3])), sl2x2, pack2(l4.z, l4.w)), a2, a3);
                  pv[4 * e4 + 0] = ex2_approx(a0);
                  pv[4 * e4 + 1] = ex2_approx(a1);
                  pv[4 * e4 + 2] = ex2_approx(a2);
                  pv[4 * e4 + 3] = ex2_approx(a3);
                }
                if (!kv_ok) {
#pragma unroll
                  for (int e = 0; e < 16; ++e) pv[e] = 0.f;
                } else if (q0 + ch * 16 + 16 > T) {
#pragma unroll
                  for (int e = 0; e < 16; ++e)
                    if (q0 + ch * 16 + e >= T) pv[e] = 0.f;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) pk[cc][e] = pack_bf16(pv[2 * e], pv[2 * e + 1]);
              }
            }
          }
          tc_fence_before_sync();
          named_bar_sync(1, ATT2_COMPUTE);        // every score column of this item has been read: P^T may overwrite them
          tc_fence_after_sync();
#pragma unroll
          for (int cc = 0; cc < NCC; ++cc)
            if (c_begin + cc < c_end) tmem_st_32x32b_x8(t_lane + TM_S + (c_begin + cc) * 8, pk[cc]);
          tmem_wait_st();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_P);
          ATT2_MARK(4 + 6 * i);
          if (i == 0 && more && p.dbuf && tid == 0) {
            // The other (Q, K) pair receives the next head's tiles while this head is processed.  Its last readers were the
            // output stores of the previous head (issued by this thread): they have long drained, the wait is a formality.
            bulk_wait_group_read<0>();
            uint8_t* nq = sQK + (pb ^ 1) * 2 * R;
            const int bn = hd_next / p.H, hn = hd_next % p.H;
            mbar_arrive_expect_tx(bar_qk + (pb ^ 1), 2 * R);
            tma_load_3d(&p.tq, bar_qk + (pb ^ 1), nq, hn * 64, 0, bn);
            tma_load_3d(&p.tk, bar_qk + (pb ^ 1), nq + R, hn * 64, 0, bn);
          }
          if (i == 0) {
            // delta = rowsum(dO * O) from the shared tiles (O parks in dS^T buffer 1), under the first dV / S^T products
            mbar_wait(bar_vdo, ph_vdo);
            ph_vdo ^= 1;
            for (int r = tid; r < T; r += ATT2_COMPUTE) {
              float d = 0.f;
#pragma unroll
              for (int k8 = 0; k8 < 8; ++k8) {
                float a[8], g[8];
                unpack8(*reinterpret_cast<const uint4*>(sDS + ATT_DS_BYTES + sw128(r, k8)), a);
                unpack8(*reinterpret_cast<const uint4*>(sDO + sw128(r, k8)), g);
#pragma unroll
                for (int e = 0; e < 8; ++e) d += a[e] * g[e];
              }
              s_delta[r] = -d * p.scale;      // stored as -delta * scale: the dS pass adds it inside one FFMA2
            }
            named_bar_sync(1, ATT2_COMPUTE);
          }
          // ---- dS^T = P^T (dP^T - delta) scale -> shared memory (K-major [128 kv][ncol q], 64-column k-blocks)
          mbar_wait(bar_dP, ph_dP);
          ph_dP ^= 1;
          tc_fence_after_sync();
          // the products that read this buffer's previous contents have completed
          if (bsel == 0) {
            if (used0) { mbar_wait(bar_free, ph_free0); ph_free0 ^= 1; }
            used0 = true;
          } else {
            if (used1) { mbar_wait(bar_free + 1, ph_free1); ph_free1 ^= 1; }
            used1 = true;
          }
          ATT2_MARK(5 + 6 * i);
#pragma unroll
          for (int pr = 0; pr < NCC / 2; ++pr) {
            uint32_t r[2][16];
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2)
              if (c_begin + 2 * pr + c2 < c_end) tmem_ld_32x32b_x16(t_lane + TM_DP + (c_begin + 2 * pr + c2) * 16, r[c2]);
            tmem_wait_ld();
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
              const int cc = 2 * pr + c2;
              const int ch = c_begin + cc;
              if (ch < c_end) {
                float pv[16];
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {     // dS = P * (dP scale + (-delta scale)), two columns per FFMA2
                  const float4 d4 = *reinterpret_cast<const float4*>(s_delta + q0 + ch * 16 + 4 * e4);
                  const f32x2 t0 = fma2(pack2(__uint_as_float(r[c2][4 * e4 + 0]), __uint_as_float(r[c2][4 * e4 + 1])), scx2, pack2(d4.x, d4.y));
                  const f32x2 t1 = fma2(pack2(__uint_as_float(r[c2][4 * e4 + 2]), __uint_as_float(r[c2][4 * e4 + 3])), scx2, pack2(d4.z, d4.w));
                  unpack2(mul2(bf16x2_to_f32x2(pk[cc][2 * e4]), t0), pv[4 * e4 + 0], pv[4 * e4 + 1]);
                  unpack2(mul2(bf16x2_to_f32x2(pk[cc][2 * e4 + 1]), t1), pv[4 * e4 + 2], pv[4 * e4 + 3]);
                }
                if (!kv_ok) {       // rows past the sequence: the tiles hold exactly TP rows, what lies behind them is not ours
#pragma unroll
                  for (int e = 0; e < 16; ++e) pv[e] = 0.f;
                }
                uint8_t* blk = buf + (ch >> 2) * 16384;
                const int k2 = (ch & 3) * 2;
                *reinterpret_cast<uint4*>(blk + sw128(row, k2)) = pack8(pv);
                *reinterpret_cast<uint4*>(blk + sw128(row, k2 + 1)) = pack8(pv + 8);
              }
            }
          }
          tc_fence_before_sync();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_dS);
          ATT2_MARK(6 + 6 * i);

          if (c == nQ - 1) {
            // ---- dV_j and dK_j rows: every thread takes 32 channels of its row of both; staged in the (dead) V_j / K_j rows
            // dV_j was complete long ago (it was issued a whole dS pass before): stage its rows while dK_j / dQ still run
            mbar_wait(bar_dv, ph_dv);
            ph_dv ^= 1;
            tc_fence_after_sync();
#pragma unroll
            for (int h2 = 0; h2 < NH2; ++h2)      // tiles hold exactly TP rows: lanes past the last row must not store
              tmem_row16_to_tile(t_lane + TM_DV + (part * NH2 + h2) * 16, sV + j * 16384, row, part * NH2 + h2, kv < TP);
            mbar_wait(bar_acc, ph_acc);
            ph_acc ^= 1;
            tc_fence_after_sync();
            ATT2_MARK(7 + 6 * i);
            if (j == nK - 1 && more && tid == 0) {
              // every product of this head has completed: dO and the O parking slot are free for the next head's tiles
              // (V follows once this head's output stores have left its rows)
              const int bn = hd_next / p.H, hn = hd_next % p.H;
              mbar_arrive_expect_tx(bar_vdo, 3 * R);
              tma_load_3d(&p.tdo, bar_vdo, sDO, hn * 64, 0, bn);
              tma_load_3d(&p.to, bar_vdo, sDS + ATT_DS_BYTES, hn * 64, 0, bn);
            }
            if (has_norm) {
              float pd = 0.f;
#pragma unroll
              for (int h2 = 0; h2 < NH2; ++h2)
                pd += norm_bwd_dot16(t_lane + TM_DK + (part * NH2 + h2) * 16, sK, kv_ok ? kv : 0, part * NH2 + h2, s_scale, s_rscale, s_dsqk, lane, kv_ok);
              s_dot[part * 128 + row] = pd;
              named_bar_sync(1, ATT2_COMPUTE);
              float dot = 0.f;
#pragma unroll
              for (int pp = 0; pp < PARTS; ++pp) dot += s_dot[pp * 128 + row];
#pragma unroll
              for (int h2 = 0; h2 < NH2; ++h2)
                norm_bwd_apply16(t_lane + TM_DK + (part * NH2 + h2) * 16, sK, kv_ok ? kv : 0, part * NH2 + h2, s_scale, s_rscale, dot,
                                 s_invk[kv_ok ? kv : 0], kv_ok);
              __syncwarp();
            } else {
#pragma unroll
              for (int h2 = 0; h2 < NH2; ++h2)
                tmem_row16_to_tile(t_lane + TM_DK + (part * NH2 + h2) * 16, sK + j * 16384, row, part * NH2 + h2, kv < TP);
            }
            tc_fence_before_sync();
            fence_proxy_async_smem();
            named_bar_sync(1, ATT2_COMPUTE);
            if (tid == 0) {
              tma_store_3d(&p.tdv, sV + j * 16384, h * 64, j * 128, b);
              tma_store_3d(&p.tdk, sK + j * 16384, h * 64, j * 128, b);
              bulk_commit_group();
            }
          }
        }
      }
      ATT2_MARK(30);
      // ---- dQ rows (complete with the last bar_acc), one 128-row q tile at a time, staged in place over Qh
      for (int m = 0; m < nQ; ++m) {
        const int qi = m * 128 + row;
        const bool ok = qi < T;
        if (has_norm) {
          float pd = 0.f;
#pragma unroll
          for (int h2 = 0; h2 < NH2; ++h2)
            pd += norm_bwd_dot16(t_lane + TM_DQ + 64 * m + (part * NH2 + h2) * 16, sQ, ok ? qi : 0, part * NH2 + h2, s_scale, s_rscale, s_dsqk, lane, ok);
          named_bar_sync(1, ATT2_COMPUTE);      // the previous exchange through s_dot has been read by everybody
          s_dot[part * 128 + row] = pd;
          named_bar_sync(1, ATT2_COMPUTE);
          float dot = 0.f;
#pragma unroll
          for (int pp = 0; pp < PARTS; ++pp) dot += s_dot[pp * 128 + row];
#pragma unroll
          for (int h2 = 0; h2 < NH2; ++h2)
            norm_bwd_apply16(t_lane + TM_DQ + 64 * m + (part * NH2 + h2) * 16, sQ, ok ? qi : 0, part * NH2 + h2, s_scale, s_rscale, dot,
                             s_invq[ok ? qi : 0], ok);
          __syncwarp();
        } else {
#pragma unroll
          for (int h2 = 0; h2 < NH2; ++h2)
            tmem_row16_to_tile(t_lane + TM_DQ + 64 * m + (part * NH2 + h2) * 16, sQ + m * 16384, row, part * NH2 + h2, qi < TP);
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      if (tid == 0 && more) {
        // V is free once the dV stores of this head have left its rows (the last one was issued a dQ epilogue ago)
        bulk_wait_group_read<0>();
        const int bn = hd_next / p.H, hn = hd_next % p.H;
        tma_load_3d(&p.tv, bar_vdo, sV, hn * 64, 0, bn);
      }
      named_bar_sync(1, ATT2_COMPUTE);
      if (tid == 0) {
        for (int m = 0; m < nQ; ++m) tma_store_3d(&p.tdq, sQ + m * 16384, h * 64, m * 128, b);
        bulk_commit_group();
        if (more && !p.dbuf) {          // single (Q, K) pair: its rows are free only when these stores have read them
          bulk_wait_group_read<0>();
          const int bn = hd_next / p.H, hn = hd_next % p.H;
          mbar_arrive_expect_tx(bar_qk, 2 * R);
          tma_load_3d(&p.tq, bar_qk, sQ, hn * 64, 0, bn);
          tma_load_3d(&p.tk, bar_qk, sK, hn * 64, 0, bn);
        }
      }
      if (has_norm && tid < 64) atomicAdd(p.dsqk + h * 64 + tid, s_dsqk[tid] * p.sqk_mul);
      // every reader of this head's per-head arrays is past the barrier above: publish the next head's values
      if (more) publish_head();
      named_bar_sync(1, ATT2_COMPUTE);
      ATT2_MARK(31);
    }
  }
  if (tid == 0) bulk_wait_group_read<0>();      // the staged output tiles have left shared memory before the CTA retires
  tc_fence_before_sync();
  __syncthreads();
  ATT3_CTAMARK(1);
  if (warp == ATT2_CW) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ backward, v3
// attn_bwd_ws_kernel with the output epilogues moved to a warpgroup of their own.  The phase marks of v2 show a head's
// ~32 k cycles as: exponential passes 7.7 k, dS passes 4.9 k, the two kv-tile epilogues 6.4 k, dQ epilogue + hand-over 5.7 k,
// waits ~5 k - all on the same eight warps.  Here
//   warps 0-7    compute: only the P^T / dS^T passes and delta (two threads per TMEM lane, as in v2);
//   warps 8-11   epilogue: ONE thread per accumulator row (64 channels), so the row dot of the normalisation backward is
//                thread-local (no exchange through shared memory, no CTA-wide barrier); it drains dV_j / dK_j / dQ from
//                tensor memory as soon as their commit barrier fires, frees the columns for the next kv tile / head
//                (bar_dvfree / bar_dkfree / bar_dqfree), stages the rows in the dead operand tiles, issues every TMA store
//                and every load of the following heads, and owns the per-head arrays it alone reads (1/||q||, 1/||k||,
//                the sqk vectors, the dL/d(sqk) accumulator);
//   warp 12      MMA issuer (as in v2, plus the "columns are free" waits).
// Registers: 13 warps -> 152 per thread for every role, no setmaxnreg.  (MEASURED: a first form with 16 warps at 128 and
// setmaxnreg - compute 144, epilogue 128, MMA warpgroup 96 - kept loop state of the MMA warp and two chunk registers of
// the exponential pass in local memory; with 224 KB of shared memory the L1 holds next to nothing, so every reload was an
// L2 round trip on the critical path, and the CTAs of one launch took 270 ... 410 us depending on how loaded the memory
// system looked from their SM - with 74 or 32 CTAs in flight every one of them ran at the fast rate.)
// Needs the two (Q, K) tile pairs (TP <= 208) and q / k that arrive normalised or without sqk; nvit_attention_bwd falls
// back to v2 otherwise.
constexpr int ATT3_THREADS = 512;   // 8 compute warps + 4 epilogue warps + the MMA warp (+ 3 idle warps: setmaxnreg works on warpgroups)
constexpr int ATT3_COMPUTE = 256;
#ifndef ATT3_REG_COMPUTE
#define ATT3_REG_COMPUTE 144
#define ATT3_REG_EPI 128
#define ATT3_REG_MMA 96
#endif
static_assert(256 * ATT3_REG_COMPUTE + 128 * ATT3_REG_EPI + 128 * ATT3_REG_MMA <= 65536, "register split of attn_bwd_ws3_kernel");
constexpr int ATT3_EPI = 128;
template <int R> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// 64 accumulator columns of this thread's TMEM lane -> registers
__device__ __forceinline__ void tmem_ld_row64(uint32_t taddr, uint32_t (&r)[64]) {
  tmem_ld_32x32b_x32(taddr, r);
  tmem_ld_32x32b_x32(taddr + 32, r + 32);
  tmem_wait_ld();
}
// plain rows (dV, or dK / dQ without normalisation): fp32 -> bf16 into the swizzled tile
__device__ __forceinline__ void row64_to_tile(const uint32_t (&r)[64], uint8_t* tile, int trow) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float g[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = __uint_as_float(r[8 * c + e]);
    *reinterpret_cast<uint4*>(tile + sw128(trow, c)) = pack8(g);
  }
}
// backward of y = s * x/||x|| for the whole row held by one thread: dL/dy in r[], the normalised bf16 row yh = s * n still
// sits in the tile, dx overwrites it.  With yh instead of n nothing needs 1/s in the first pass:
//   dot = sum_c g_c s_c n_c = sum_c g_c yh_c ;  dx_c = (g_c s_c - yh_c / s_c * dot) / ||x|| ;  dL/ds_c = sum_rows g_c yh_c / s_c
// (the 1/s_c of dL/ds is applied once per head by the caller).  Packed fp32 pairs throughout: 2 (+1 with DSQK) instructions
// per pair in the first pass, 6 in the second (MEASURED on the scalar form with n = yh / s recomputed per element and both
// the q and the k rows feeding dL/ds: 6.2 k cycles per 128-row tile against 0.55 k for the plain dV rows - the epilogue
// warpgroup, not the compute warps, set the pace of the kernel).
// q and k share the scale vector, so sum over q rows of g n equals the sum over k rows (both are
// sum_ij dS_ij s_c nq_ic nk_jc): only the dQ rows feed the accumulator (DSQK) and the caller doubles it.
// Rows that do not exist (`valid` false) read `zero_row` instead of the tile: their g is finite garbage, g * 0 adds nothing.
// Warp-collective with DSQK (shuffles inside).
template <bool DSQK>
__device__ __forceinline__ void norm_bwd_row64(const uint32_t (&r)[64], uint8_t* tile, int trow, const uint8_t* zero_row,
                                               const float* s_scale, const float* s_rscale, float* s_dsqk, float inv, int lane,
                                               bool valid) {
  const uint8_t* const src = valid ? tile + trow * 128 : zero_row;
  const int sw = valid ? (trow & 7) : 0;
  f32x2 dot2 = pack2(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 y0 = *reinterpret_cast<const uint4*>(src + (((2 * q) ^ sw) << 4));
    const uint4 y1 = *reinterpret_cast<const uint4*>(src + (((2 * q + 1) ^ sw) << 4));
    const uint32_t yw[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
    float acc[16];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const f32x2 g2 = pack2(__uint_as_float(r[q * 16 + 2 * e]), __uint_as_float(r[q * 16 + 2 * e + 1]));
      const f32x2 y2 = bf16x2_to_f32x2(yw[e]);
      if constexpr (DSQK) {
        const f32x2 t2 = mul2(g2, y2);
        dot2 = add2(dot2, t2);
        unpack2(t2, acc[2 * e], acc[2 * e + 1]);
      } else {
        dot2 = fma2(g2, y2, dot2);
      }
    }
    if constexpr (DSQK) reduce16_to_smem(acc, s_dsqk + q * 16, lane);
  }
  if (!valid) return;
  float d0, d1;
  unpack2(dot2, d0, d1);
  const float nb = -(d0 + d1) * inv;
  const f32x2 a2 = pack2(inv, inv), nb2 = pack2(nb, nb);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 y = *reinterpret_cast<const uint4*>(src + ((c ^ sw) << 4));
    const uint32_t yw[4] = {y.x, y.y, y.z, y.w};
    const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + c * 8), sc1 = *reinterpret_cast<const float4*>(s_scale + c * 8 + 4);
    const float4 rs0 = *reinterpret_cast<const float4*>(s_rscale + c * 8), rs1 = *reinterpret_cast<const float4*>(s_rscale + c * 8 + 4);
    const f32x2 s2[4] = {pack2(sc0.x, sc0.y), pack2(sc0.z, sc0.w), pack2(sc1.x, sc1.y), pack2(sc1.z, sc1.w)};
    const f32x2 q2[4] = {pack2(rs0.x, rs0.y), pack2(rs0.z, rs0.w), pack2(rs1.x, rs1.y), pack2(rs1.z, rs1.w)};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const f32x2 g2 = pack2(__uint_as_float(r[8 * c + 2 * e]), __uint_as_float(r[8 * c + 2 * e + 1]));
      const f32x2 x1 = mul2(g2, s2[e]);                              // g s
      const f32x2 x2 = mul2(bf16x2_to_f32x2(yw[e]), q2[e]);          // yh / s = n
      o[e] = f32x2_to_bf16x2(fma2(x1, a2, mul2(x2, nb2)));           // (g s - n dot) / ||x||
    }
    *reinterpret_cast<uint4*>(tile + trow * 128 + ((c ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(ATT3_THREADS, 1) attn_bwd_ws3_kernel(const __grid_constant__ AttnParams p) {
  constexpr int PARTS = 2, NCC = 4;
  pdl_enter();
  ATT3_CTAMARK(0);
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int T = p.T, TP = p.TP;
  const uint32_t R = static_cast<uint32_t>(TP) * 128u;
  uint8_t* sQK = smem;                                  // [2][Q | K]
  uint8_t* sV = sQK + 4 * R;
  uint8_t* sDO = sV + R;
  uint8_t* sDS = sDO + R;                               // two dS^T buffers (O parks in the second one until delta is taken)
  float* s_lse = reinterpret_cast<float*>(sDS + 2 * ATT_DS_BYTES);   // [2][TP]: written by the epilogue warps a head ahead, read by the compute warps
  float* s_delta = s_lse + 2 * TP;                                   // compute warps
  float* s_invq = s_delta + TP;                                      // epilogue warps (and everything below)
  float* s_invk = s_invq + TP;
  float* s_scale = s_invk + TP;
  float* s_rscale = s_scale + 64;
  float* s_dsqk = s_rscale + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dsqk + 64);
  uint64_t* bar_qk = bars + 0;       // [2] TMA: q, k tiles of pair b
  uint64_t* bar_do = bars + 2;       // TMA: dO, O
  uint64_t* bar_v = bars + 3;        // TMA: V
  uint64_t* bar_S = bars + 4;        // MMA -> compute
  uint64_t* bar_dP = bars + 5;       // MMA -> compute
  uint64_t* bar_P = bars + 6;        // compute -> MMA (one arrival per compute warp)
  uint64_t* bar_dS = bars + 7;       // compute -> MMA
  uint64_t* bar_free = bars + 8;     // [2] MMA -> compute: the products reading dS^T buffer b have completed
  uint64_t* bar_acc = bars + 10;     // MMA -> epilogue: dK_j (and every dQ contribution so far) complete
  uint64_t* bar_dv = bars + 11;      // MMA -> epilogue: dV_j complete
  uint64_t* bar_dvfree = bars + 12;  // epilogue -> MMA: dV_j has left tensor memory (one arrival per epilogue warp)
  uint64_t* bar_dkfree = bars + 13;  // epilogue -> MMA: dK_j has left tensor memory
  uint64_t* bar_dqfree = bars + 14;  // [2] epilogue -> MMA: q tile m of the head's dQ has left tensor memory
  uint64_t* bar_next = bars + 16;    // [2] epilogue -> compute, MMA: s_head[(n + 1) & 3] and s_lse[(n + 1) & 1] are written (bar n & 1)
  uint64_t* bar_dqst = bars + 18;    // epilogue -> store warp: the head's dQ rows are staged over Qh
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 19);
  int* const s_head = reinterpret_cast<int*>(bars + 20);            // [4] head of iteration n at [n & 3]; -1: the CTA is done
  uint8_t* const s_zero = reinterpret_cast<uint8_t*>(bars) + 256;   // one all-zero tile row (stands in for rows that do not exist)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool has_norm = p.sqk != nullptr;
  const int nQ = p.nQ, nK = p.nK, nItems = nQ * nK;
  const int nheads = p.B * p.H;
  constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

  if (tid == 0) {
    mbar_init(bar_qk + 0, 1);
    mbar_init(bar_qk + 1, 1);
    mbar_init(bar_do, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_S, 1);
    mbar_init(bar_dP, 1);
    mbar_init(bar_P, ATT3_COMPUTE / 32);
    mbar_init(bar_dS, ATT3_COMPUTE / 32);
    mbar_init(bar_free + 0, 1);
    mbar_init(bar_free + 1, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_dv, 1);
    mbar_init(bar_dvfree, ATT3_EPI / 32);
    mbar_init(bar_dkfree, ATT3_EPI / 32);
    mbar_init(bar_dqfree + 0, ATT3_EPI / 32);
    mbar_init(bar_dqfree + 1, ATT3_EPI / 32);
    mbar_init(bar_next + 0, ATT3_EPI / 32);
    mbar_init(bar_next + 1, ATT3_EPI / 32);
    mbar_init(bar_dqst, 1);
    s_head[0] = static_cast<int>(blockIdx.x);
    fence_barrier_init();
    // the first head's tiles, and the second head's q / k into the other pair
    const int hd = blockIdx.x, b0 = hd / p.H, h0 = hd % p.H;
    mbar_arrive_expect_tx(bar_qk, 2 * R);
    tma_load_3d(&p.tq, bar_qk, sQK, h0 * 64, 0, b0);
    tma_load_3d(&p.tk, bar_qk, sQK + R, h0 * 64, 0, b0);
    mbar_arrive_expect_tx(bar_do, 2 * R);
    tma_load_3d(&p.tdo, bar_do, sDO, h0 * 64, 0, b0);
    tma_load_3d(&p.to, bar_do, sDS + ATT_DS_BYTES, h0 * 64, 0, b0);
    mbar_arrive_expect_tx(bar_v, R);
    tma_load_3d(&p.tv, bar_v, sV, h0 * 64, 0, b0);
    const int hd1 = hd + gridDim.x;
    if (hd1 < nheads) {
      const int b1 = hd1 / p.H, h1 = hd1 % p.H;
      mbar_arrive_expect_tx(bar_qk + 1, 2 * R);
      tma_load_3d(&p.tq, bar_qk + 1, sQK + 2 * R, h1 * 64, 0, b1);
      tma_load_3d(&p.tk, bar_qk + 1, sQK + 3 * R, h1 * 64, 0, b1);
    }
  }
  if (warp == 12) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  // The dS^T buffers start out all zero (except where the first head's O tile is landing): accumulator lanes past the last
  // row are fed from parts of them that no pass writes, and norm_bwd_row64 relies on those lanes holding FINITE garbage.
  for (uint32_t o = tid * 16u; o < 2u * ATT_DS_BYTES; o += ATT3_THREADS * 16u)
    if (o < static_cast<uint32_t>(ATT_DS_BYTES) || o >= ATT_DS_BYTES + R) *reinterpret_cast<uint4*>(sDS + o) = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 8) reinterpret_cast<uint4*>(s_zero)[tid] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  for (int r = tid; r < TP; r += ATT3_THREADS) {   // rows >= T are never written again and must stay finite
    s_lse[r] = 0.f;
    s_lse[TP + r] = 0.f;
    s_invq[r] = 0.f;
    s_invk[r] = 0.f;
    s_delta[r] = 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t sV_a = smem_u32(sV), sDO_a = smem_u32(sDO), sDS_a = smem_u32(sDS);

  if (warp >= 12) {
    setmaxnreg_dec<ATT3_REG_MMA>();
    if (warp == 12) {
      // ===================== MMA issuer =====================
      uint32_t ph = 0;   // bits: 0/1 qk pair, 2 do, 3 v, 4 P, 5 dS, 6 dvfree, 7 dkfree, 8/9 dqfree, 10/11 next
      auto ncol_of = [&](int c) { return min(128, TP - 128 * c); };
      int n = 0;
      bool more_heads = false;
      bool tile_before = false;          // a kv tile of this CTA has been processed before: its accumulators must have been drained
      for (;; ++n) {
        // (whether head n exists was settled during the last item of head n - 1, see below)
        const int pb = n & 1;
        const uint32_t sQ_a = smem_u32(sQK + pb * 2 * R), sK_a = sQ_a + R;
        auto issue_S = [&](int j, int c, uint32_t swap = 0u) {       // S^T(j,c) = Kh_j Qh_c^T  [128 kv x ncol]; swap: the OTHER (Q, K) pair
          const uint32_t q_a = smem_u32(sQK + (pb ^ swap) * 2 * R);
          const uint64_t da = umma_smem_desc(q_a + R + j * 16384, 16, 1024), db = umma_smem_desc(q_a + c * 16384, 16, 1024);
          const uint32_t id = idesc_kk_n(ncol_of(c));
          if (elect_one()) {
            umma_bf16_ss(tmem_base + TM_S, da, db, id, 0u);
#pragma unroll
            for (int ks = 1; ks < 4; ++ks) umma_bf16_ss_acc(tmem_base + TM_S, da + 2 * ks, db + 2 * ks, id);
            umma_commit(bar_S);
          }
          __syncwarp();
        };
        auto issue_dP = [&](int j, int c) {      // dP^T(j,c) = V_j dO_c^T
          const uint64_t da = umma_smem_desc(sV_a + j * 16384, 16, 1024), db = umma_smem_desc(sDO_a + c * 16384, 16, 1024);
          const uint32_t id = idesc_kk_n(ncol_of(c));
          if (elect_one()) {
            umma_bf16_ss(tmem_base + TM_DP, da, db, id, 0u);
#pragma unroll
            for (int ks = 1; ks < 4; ++ks) umma_bf16_ss_acc(tmem_base + TM_DP, da + 2 * ks, db + 2 * ks, id);
            umma_commit(bar_dP);
          }
          __syncwarp();
        };
        ATT3_MMARK(0);
        if (n == 0) {      // later heads: their first S^T product was issued behind the last dV product of the head before
          mbar_wait_flip(bar_qk + pb, ph, pb);
          tc_fence_after_sync();
          issue_S(0, 0);
        }
        ATT3_MMARK(1);
        mbar_wait_flip(bar_do, ph, 2);
        mbar_wait_flip(bar_v, ph, 3);
        tc_fence_after_sync();
        issue_dP(0, 0);
        ATT3_MMARK(2);
        int i = 0;
        for (int j = 0; j < nK; ++j) {
          const int kv_steps = min(8, (TP - j * 128) >> 4);
          for (int c = 0; c < nQ; ++c, ++i) {
            const int nks = ncol_of(c) >> 4;
            const int jn = (c + 1 < nQ) ? j : j + 1, cn = (c + 1 < nQ) ? c + 1 : 0;
            const bool has_next = i + 1 < nItems;
            const uint32_t buf_a = sDS_a + (i & 1) * ATT_DS_BYTES;
            // ---- dV_j += P^T(i) dO_c   (A = P^T from tensor memory)
            mbar_wait_flip(bar_P, ph, 4);
            if (c == 0 && tile_before) mbar_wait_flip(bar_dvfree, ph, 6);   // the previous tile's dV rows are out
            tc_fence_after_sync();
            ATT3_MMARK(3 + 4 * i);
            {
              const uint64_t dd = umma_smem_desc(sDO_a + c * 16384, 8192, 1024);
              const uint32_t first = c > 0 ? 1u : 0u;
              if (elect_one()) {
                umma_bf16_ts(tmem_base + TM_DV, tmem_base + TM_S, dd, IDESC_KM(64), first);
#pragma unroll
                for (int ks = 1; ks < 8; ++ks)
                  if (ks < nks) umma_bf16_ts(tmem_base + TM_DV, tmem_base + TM_S + ks * 8, dd + 128 * ks, IDESC_KM(64), 1u);
                if (c == nQ - 1) umma_commit(bar_dv);
              }
              __syncwarp();
            }
            if (has_next) {
              issue_S(jn, cn);
            } else {
              // Last item of the head.  Is there another head?  (Published by the epilogue warpgroup during its iteration n,
              // thousands of cycles ago.)  If so its first S^T product goes out now, behind this head's last dV product: it
              // runs under the last dS pass, and the compute warps find it finished when they enter the head
              // (MEASURED: issued at the top of the head it reached them ~1.8 k cycles after they had entered it).
              mbar_wait_flip(bar_next + (n & 1), ph, 10 + (n & 1));
              more_heads = s_head[(n + 1) & 3] >= 0;
              if (more_heads) {
                mbar_wait_flip(bar_qk + (pb ^ 1), ph, pb ^ 1);
                tc_fence_after_sync();
                issue_S(0, 0, 1u);
              }
            }
            ATT3_MMARK(4 + 4 * i);
            mbar_wait_flip(bar_dS, ph, 5);
            tc_fence_after_sync();
            ATT3_MMARK(5 + 4 * i);
            if (has_next && c != nQ - 1) issue_dP(jn, cn);
            if (c == 0 && tile_before) mbar_wait_flip(bar_dkfree, ph, 7);
            if (j == 0 && n > 0) {       // first contribution to dQ_c of this head: the previous head's rows must have left
              mbar_wait_flip(bar_dqfree + c, ph, 8 + c);
            }
            tc_fence_after_sync();
            // ---- dK_j += dS^T(i) Qh_c ; dQ_c += dS(i) Kh_j
            {
              const uint64_t dp = umma_smem_desc(buf_a, 16, 1024);
              const uint64_t dq = umma_smem_desc(sQ_a + c * 16384, 8192, 1024);
              const uint64_t dsm = umma_smem_desc(buf_a, 16384, 1024), dkm = umma_smem_desc(sK_a + j * 16384, 8192, 1024);
              const uint32_t first_k = c > 0 ? 1u : 0u, first_q = j > 0 ? 1u : 0u;
              if (elect_one()) {
                umma_bf16_ss(tmem_base + TM_DK, dp, dq, IDESC_KM(64), first_k);
#pragma unroll
                for (int ks = 1; ks < 8; ++ks)
                  if (ks < nks) umma_bf16_ss_acc(tmem_base + TM_DK, dp + static_cast<uint64_t>((ks >> 2) * 1024 + (ks & 3) * 2), dq + 128 * ks, IDESC_KM(64));
                umma_bf16_ss(tmem_base + TM_DQ + 64 * c, dsm, dkm, IDESC_MM(64), first_q);
#pragma unroll
                for (int ks = 1; ks < 8; ++ks)
                  if (ks < kv_steps) umma_bf16_ss_acc(tmem_base + TM_DQ + 64 * c, dsm + 128 * ks, dkm + 128 * ks, IDESC_MM(64));
                umma_commit(bar_free + (i & 1));
                if (c == nQ - 1) umma_commit(bar_acc);
              }
              __syncwarp();
            }
            if (has_next && c == nQ - 1) issue_dP(jn, cn);
            if (c == nQ - 1) tile_before = true;
            ATT3_MMARK(6 + 4 * i);
          }
        }
        if (!more_heads) break;
      }
    }
    else if (warp == 13 && lane == 0) {
      // ===================== store warp (one thread): the head's dQ tiles out, the head after next's Q / K in =====================
      // The stores of dQ are queued behind the next head's dO / O / V loads in the SM's bulk-copy pipeline, so waiting for them
      // to have read their rows takes ~4 k cycles (MEASURED when thread 0 of the epilogue warpgroup did this at the top of its
      // iteration: the whole warpgroup waited with it at its next barrier, and the MMA warp for the dV columns behind that).
      // This thread has nothing else to do.
      uint32_t ph = 0;   // bits: 0 dqst, 1/2 next
      mbar_wait_flip(bar_next + 0, ph, 1);          // iteration 0's publication (head 1) is not needed here, only consumed
      for (int n = 0;; ++n) {
        const int hd = s_head[n & 3];                // visible: kernel entry (n = 0) or the wait on bar_dqst of iteration n - 1
        if (hd < 0) break;
        const int b = static_cast<int>(__umulhi(static_cast<unsigned>(hd), p.h_magic)), h = hd - b * p.H;
        const int pb = n & 1;
        uint8_t* const sQ = sQK + pb * 2 * R;
        mbar_wait_flip(bar_dqst, ph, 0);
        for (int m = 0; m < nQ; ++m) tma_store_3d(&p.tdq, sQ + m * 16384, h * 64, m * 128, b);
        bulk_commit_group();
        if (s_head[(n + 1) & 3] < 0) break;          // no iteration n + 1: nothing publishes a head n + 2
        mbar_wait_flip(bar_next + ((n + 1) & 1), ph, 1 + ((n + 1) & 1));
        const int hd2 = s_head[(n + 2) & 3];
        if (hd2 >= 0) {
          bulk_wait_group_read<0>();                 // the dQ stores have read pair pb (its dK stores: see the epilogue warpgroup)
          const int b2 = static_cast<int>(__umulhi(static_cast<unsigned>(hd2), p.h_magic)), h2 = hd2 - b2 * p.H;
          mbar_arrive_expect_tx(bar_qk + pb, 2 * R);
          tma_load_3d(&p.tq, bar_qk + pb, sQ, h2 * 64, 0, b2);
          tma_load_3d(&p.tk, bar_qk + pb, sQ + R, h2 * 64, 0, b2);
        }
      }
      bulk_wait_group_read<0>();                     // the staged tiles have left shared memory before the CTA retires
    }
  } else if (warp < 8) {
    setmaxnreg_inc<ATT3_REG_COMPUTE>();
    // ===================== compute warps: the two passes of every item, delta once per head =====================
    const int wq = warp & 3;
    const int part = warp >> 2;
    const int row = wq * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const f32x2 sl2x2 = pack2(p.scale * LOG2E, p.scale * LOG2E), scx2 = pack2(p.scale, p.scale);
    uint32_t ph = 0;   // bits: 0 do, 1 S, 2 dP, 3/4 free, 5/6 dS^T buffer used before, 7/8 next
    if (tid < T) s_lse[tid] = -p.lse[static_cast<long long>(blockIdx.x) * T + tid] * LOG2E;   // (b * H + h) * T = head * T
    named_bar_sync(1, ATT3_COMPUTE);

    int n = 0;
    for (;; ++n) {
      if (n > 0) {       // the epilogue warpgroup has published the next head (or the end) and its log-sum-exp row
        mbar_wait_flip(bar_next + ((n - 1) & 1), ph, 7 + ((n - 1) & 1));
        if (s_head[n & 3] < 0) break;
      }
      const float* const lse_n = s_lse + (n & 1) * TP;
      ATT3_MARK(0);
      int i = 0;
      for (int j = 0; j < nK; ++j) {
        const int kv = j * 128 + row;
        const bool kv_ok = kv < T;
        for (int c = 0; c < nQ; ++c, ++i) {
          const int nch = min(128, TP - 128 * c) >> 4;
          const int c_begin = (part * nch) / PARTS, c_end = ((part + 1) * nch) / PARTS;
          const int q0 = c * 128;
          const int bsel = i & 1;
          uint8_t* const buf = sDS + bsel * ATT_DS_BYTES;
          uint32_t pk[NCC][8];
          // ---- P^T = exp2(scale log2e S^T - lse log2e), kept as bf16 pairs
          ATT3_MARK(2 + 6 * i);
          mbar_wait_flip(bar_S, ph, 1);
          tc_fence_after_sync();
          ATT3_MARK(3 + 6 * i);
#pragma unroll
          for (int pr = 0; pr < NCC / 2; ++pr) {
            uint32_t r[2][16];
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2)
              if (c_begin + 2 * pr + c2 < c_end) tmem_ld_32x32b_x16(t_lane + TM_S + (c_begin + 2 * pr + c2) * 16, r[c2]);
            tmem_wait_ld();
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
              const int cc = 2 * pr + c2;
              const int ch = c_begin + cc;
              if (ch < c_end) {
                float pv[16];
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                  const float4 l4 = *reinterpret_cast<const float4*>(lse_n + q0 + ch * 16 + 4 * e4);
                  float a0, a1, a2, a3;
                  unpack2(fma2(pack2(__uint_as_float(r[c2][4 * e4 + 0]), __uint_as_float(r[c2][4 * e4 + 1])), sl2x2, pack2(l4.x, l4.y)), a0, a1);
                  unpack2(fma2(pack2(__uint_as_float(r[c2][4 * e4 + 2]), __uint_as_float(r[c2][4 * e4 + 3])), sl2x2, pack2(l4.z, l4.w)), a2, a3);
                  pv[4 * e4 + 0] = ex2_approx(a0);
                  pv[4 * e4 + 1] = ex2_approx(a1);
                  pv[4 * e4 + 2] = ex2_approx(a2);
                  pv[4 * e4 + 3] = ex2_approx(a3);
                }
                // No masking of rows / columns >= T: q, k, v and dO arrive with those rows zero-filled and lse, delta hold zeros
                // there, so such entries of P^T and dS^T are finite and only ever multiply zeros (dV: dO rows; dK: Qh rows;
                // dQ: Kh rows) or land in output rows the stores clip.  (The per-element selects were 40 % of this pass.)
#pragma unroll
                for (int e = 0; e < 8; ++e) pk[cc][e] = pack_bf16(pv[2 * e], pv[2 * e + 1]);
              }
            }
          }
          tc_fence_before_sync();
          named_bar_sync(1, ATT3_COMPUTE);        // every score column of this item has been read: P^T may overwrite them
          tc_fence_after_sync();
#pragma unroll
          for (int cc = 0; cc < NCC; ++cc)
            if (c_begin + cc < c_end) tmem_st_32x32b_x8(t_lane + TM_S + (c_begin + cc) * 8, pk[cc]);
          tmem_wait_st();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_P);
          ATT3_MARK(4 + 6 * i);
          if (i == 0) {
            // delta = rowsum(dO * O) from the shared tiles (O parks in dS^T buffer 1), under the first dV / S^T products
            mbar_wait_flip(bar_do, ph, 0);
            for (int r = tid; r < T; r += ATT3_COMPUTE) {
              float d = 0.f;
#pragma unroll
              for (int k8 = 0; k8 < 8; ++k8) {
                float a[8], g[8];
                unpack8(*reinterpret_cast<const uint4*>(sDS + ATT_DS_BYTES + sw128(r, k8)), a);
                unpack8(*reinterpret_cast<const uint4*>(sDO + sw128(r, k8)), g);
#pragma unroll
                for (int e = 0; e < 8; ++e) d += a[e] * g[e];
              }
              s_delta[r] = -d * p.scale;
            }
            named_bar_sync(1, ATT3_COMPUTE);
          }
          // ---- dS^T = P^T (dP^T - delta) scale -> shared memory
          mbar_wait_flip(bar_dP, ph, 2);
          tc_fence_after_sync();
          if ((ph >> (5 + bsel)) & 1u) mbar_wait_flip(bar_free + bsel, ph, 3 + bsel);
          ph |= 1u << (5 + bsel);
          ATT3_MARK(5 + 6 * i);
#pragma unroll
          for (int pr = 0; pr < NCC / 2; ++pr) {
            uint32_t r[2][16];
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2)
              if (c_begin + 2 * pr + c2 < c_end) tmem_ld_32x32b_x16(t_lane + TM_DP + (c_begin + 2 * pr + c2) * 16, r[c2]);
            tmem_wait_ld();
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
              const int cc = 2 * pr + c2;
              const int ch = c_begin + cc;
              if (ch < c_end) {
                float pv[16];
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                  const float4 d4 = *reinterpret_cast<const float4*>(s_delta + q0 + ch * 16 + 4 * e4);
                  const f32x2 t0 = fma2(pack2(__uint_as_float(r[c2][4 * e4 + 0]), __uint_as_float(r[c2][4 * e4 + 1])), scx2, pack2(d4.x, d4.y));
                  const f32x2 t1 = fma2(pack2(__uint_as_float(r[c2][4 * e4 + 2]), __uint_as_float(r[c2][4 * e4 + 3])), scx2, pack2(d4.z, d4.w));
                  unpack2(mul2(bf16x2_to_f32x2(pk[cc][2 * e4]), t0), pv[4 * e4 + 0], pv[4 * e4 + 1]);
                  unpack2(mul2(bf16x2_to_f32x2(pk[cc][2 * e4 + 1]), t1), pv[4 * e4 + 2], pv[4 * e4 + 3]);
                }
                uint8_t* blk = buf + (ch >> 2) * 16384;
                const int k2 = (ch & 3) * 2;
                *reinterpret_cast<uint4*>(blk + sw128(row, k2)) = pack8(pv);
                *reinterpret_cast<uint4*>(blk + sw128(row, k2 + 1)) = pack8(pv + 8);
              }
            }
          }
          tc_fence_before_sync();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_dS);
          ATT3_MARK(6 + 6 * i);
        }
      }
      ATT3_MARK(31);
    }
  } else {
    // ===================== epilogue warpgroup: one thread per accumulator row =====================
    setmaxnreg_dec<ATT3_REG_EPI>();
    const int ew = warp - 8;                 // = warp & 3: the TMEM lane quarter this warp may access
    const int et = tid - ATT3_COMPUTE;       // 0 .. 127
    const int row = ew * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    uint32_t ph = 0;   // bits: 0 dv, 1 acc
    // This warpgroup also owns the head sequence.  Heads 0 and 1 of a CTA are blockIdx.x and blockIdx.x + gridDim.x (their
    // tiles are requested at kernel entry); every further head is claimed from a device-wide counter, so an SM that runs
    // slowly takes fewer heads (MEASURED with the static stride: the CTAs of one launch took 251 ... 303 us, one SM pair
    // 360 us - and the launch lasts as long as its slowest CTA).  Thread 0 claims head n + 2 at the top of iteration n,
    // publishes it at the top of iteration n + 1 in s_head[], and the compute warps and the MMA warp learn through
    // bar_next - together with the next head's log-sum-exp row, which this warpgroup fetches for them - whether another
    // head follows.
    float nx_iq[2] = {0.f, 0.f}, nx_ik[2] = {0.f, 0.f}, nx_lse[2] = {0.f, 0.f}, nx_sc = 0.f;
    auto fetch_head_arrays = [&](int head) {       // per-head vectors of `head`, into registers (consumed an iteration later)
      const int bb = static_cast<int>(__umulhi(static_cast<unsigned>(head), p.h_magic)), hh = head - bb * p.H;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int r = et + k * ATT3_EPI;
        if (r < T) {
          nx_lse[k] = p.lse[static_cast<long long>(head) * T + r];
          if (has_norm && p.inv_q != nullptr) {
            nx_iq[k] = p.inv_q[(static_cast<long long>(bb) * T + r) * p.ld_inv_q + hh];
            nx_ik[k] = p.inv_k[(static_cast<long long>(bb) * T + r) * p.ld_inv_k + hh];
          }
        }
      }
      if (has_norm && et < 64) nx_sc = p.sqk[hh * 64 + et];
    };
    fetch_head_arrays(blockIdx.x);
    // thread 0: the head of iteration n + 1 is claim_base + claim_raw (>= nheads: none); claim_raw is the counter's answer, which
    // nothing may touch before the top of the next iteration (148 CTAs hit one address: the answer takes up to ~4 k cycles)
    int claim_raw = 0, claim_base = static_cast<int>(blockIdx.x + gridDim.x);
    int n = 0;
    for (;; ++n) {
      const int hd = s_head[n & 3];                // published an iteration ago (n = 0: at kernel entry)
      if (hd < 0) break;
      const int b = static_cast<int>(__umulhi(static_cast<unsigned>(hd), p.h_magic)), h = hd - b * p.H;
      const int pb = n & 1;
      uint8_t* const sQ = sQK + pb * 2 * R;
      uint8_t* const sK = sQ + R;
      ATT3_EMARK(12);
      // per-head arrays of this warpgroup (its previous readers are behind the barrier that ends the loop body)
      if (has_norm) {
        if (p.inv_q != nullptr) {
          // 1/||q||, 1/||k|| of this head were fetched a head ahead (TP <= 208: two rows per thread)
          if (et < T) { s_invq[et] = nx_iq[0]; s_invk[et] = nx_ik[0]; }
          if (et + ATT3_EPI < T) { s_invq[et + ATT3_EPI] = nx_iq[1]; s_invk[et + ATT3_EPI] = nx_ik[1]; }
        }
        if (et < 64) {
          const float sc = nx_sc * p.sqk_mul;
          s_scale[et] = sc;
          s_rscale[et] = sc != 0.f ? 1.f / sc : 0.f;
          s_dsqk[et] = 0.f;
        }
      }
      if (et == 0) {
        const int claimed = claim_base + claim_raw;          // first use of the counter value fetched an iteration ago
        s_head[(n + 1) & 3] = claimed < nheads ? claimed : -1;
      }
      named_bar_sync(2, ATT3_EPI);
      const int hd_next = s_head[(n + 1) & 3];
      const bool more = hd_next >= 0;
      ATT3_EMARK(13);
      if (et == 0) {
        // the head of iteration n + 2 (MEASURED: using the counter's answer right here - even just adding to it - kept this
        // thread, and at the next barrier its warpgroup, waiting 3 - 5 k cycles)
        claim_base = more ? 2 * static_cast<int>(gridDim.x) : nheads;
        if (more) claim_raw = atomicAdd(p.work, 1);
        // (MEASURED: L2 prefetches (cp.async.bulk.prefetch.tensor) of the next head's V / dO / O issued here, a head ahead of
        // their exposed loads, made the launch slower, 263 -> 285 us: this thread sat ~3 k cycles in their issue.)
      }
      ATT3_EMARK(14);
      if (more) fetch_head_arrays(hd_next);
      ATT3_EMARK(0);
      for (int j = 0; j < nK; ++j) {
        const int kv = j * 128 + row;
        const bool kv_ok = kv < T;
        uint32_t r[64];
        // ---- dV_j
        ATT3_EMARK(1 + 6 * j);
        mbar_wait_flip(bar_dv, ph, 0);
        tc_fence_after_sync();
        ATT3_EMARK(2 + 6 * j);
        __syncwarp();
        tmem_ld_row64(t_lane + TM_DV, r);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dvfree);
        if (kv < TP) row64_to_tile(r, sV + j * 16384, row);      // tiles hold exactly TP rows: lanes past the last row must not store
        fence_proxy_async_smem();
        named_bar_sync(2, ATT3_EPI);
        if (et == 0) {
          tma_store_3d(&p.tdv, sV + j * 16384, h * 64, j * 128, b);
          bulk_commit_group();
        }
        if (j == 0) {
          // the next head's log-sum-exp row (requested at the top of the iteration) goes to the buffer the compute warps read
          // during head n + 1; they finished head n - 1, its last reader, before this iteration could start
          if (more) {
            float* const lse_w = s_lse + ((n + 1) & 1) * TP;
            if (et < T) lse_w[et] = -nx_lse[0] * LOG2E;
            if (et + ATT3_EPI < T) lse_w[et + ATT3_EPI] = -nx_lse[1] * LOG2E;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_next + (n & 1));
        }
        // ---- dK_j
        ATT3_EMARK(3 + 6 * j);
        mbar_wait_flip(bar_acc, ph, 1);
        tc_fence_after_sync();
        ATT3_EMARK(4 + 6 * j);
        if (j == nK - 1 && more && et == 0) {
          // every product of this head has completed: dO and the O parking slot are free, and V once the dV stores have read
          // its rows (the last one was issued a dS pass ago)
          const int bn = static_cast<int>(__umulhi(static_cast<unsigned>(hd_next), p.h_magic)), hn = hd_next - bn * p.H;
          mbar_arrive_expect_tx(bar_do, 2 * R);
          tma_load_3d(&p.tdo, bar_do, sDO, hn * 64, 0, bn);
          tma_load_3d(&p.to, bar_do, sDS + ATT_DS_BYTES, hn * 64, 0, bn);
          bulk_wait_group_read<0>();
          mbar_arrive_expect_tx(bar_v, R);
          tma_load_3d(&p.tv, bar_v, sV, hn * 64, 0, bn);
        }
        __syncwarp();
        tmem_ld_row64(t_lane + TM_DK, r);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dkfree);
        if (has_norm) {
          norm_bwd_row64<false>(r, sK, kv_ok ? kv : 0, s_zero, s_scale, s_rscale, s_dsqk, s_invk[kv_ok ? kv : 0], lane, kv_ok);
        } else if (kv < TP) {
          row64_to_tile(r, sK + j * 16384, row);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, ATT3_EPI);
        if (et == 0) {
          tma_store_3d(&p.tdk, sK + j * 16384, h * 64, j * 128, b);
          bulk_commit_group();
        }
        __syncwarp();
        ATT3_EMARK(5 + 6 * j);
      }
      // ---- dQ rows (complete with the last bar_acc), one 128-row q tile at a time, staged in place over Qh
      for (int m = 0; m < nQ; ++m) {
        const int qi = m * 128 + row;
        const bool ok = qi < T;
        uint32_t r[64];
        __syncwarp();
        tmem_ld_row64(t_lane + TM_DQ + 64 * m, r);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dqfree + m);
        if (has_norm) {
          norm_bwd_row64<true>(r, sQ, ok ? qi : 0, s_zero, s_scale, s_rscale, s_dsqk, s_invq[ok ? qi : 0], lane, ok);
        } else if (qi < TP) {
          row64_to_tile(r, sQ + m * 16384, row);
        }
        ATT3_EMARK(20 + m);
      }
      fence_proxy_async_smem();
      named_bar_sync(2, ATT3_EPI);        // the only barrier at the end of an iteration: every reader of the per-head arrays is behind it
      ATT3_EMARK(30);
      if (et == 0) {
        bulk_wait_group_read<0>();        // this thread's dV / dK stores (issued thousands of cycles ago) have read their rows:
        mbar_arrive(bar_dqst);            // the store warp may send dQ out and, after it, refill this (Q, K) pair
      }
      // s_dsqk holds sum over the q rows of g * yh: times 1/s for g * n, times two for the k rows (see norm_bwd_row64)
      if (has_norm && et < 64) atomicAdd(p.dsqk + h * 64 + et, s_dsqk[et] * s_rscale[et] * (2.f * p.sqk_mul));
      ATT3_EMARK(31);
    }
    if (et == 0) bulk_wait_group_read<0>();      // the staged output tiles have left shared memory before the CTA retires
  }
  tc_fence_before_sync();
  __syncthreads();
  ATT3_CTAMARK(1);
  if (tid == 0) {
    // the last CTA to get here puts the head counter back to zero for the next launch (every claim precedes the claimer's count)
    __threadfence();
    if (atomicInc(reinterpret_cast<unsigned*>(p.work) + 1, gridDim.x - 1) == gridDim.x - 1) {
      __threadfence();
      atomicExch(p.work, 0);
    }
  }
  if (warp == 12) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

static int make_head_tmap(CUtensorMap* m, const void* base, long long ld, int B, int H, int T, int box_rows = ATT_ROWS) {
  const uint64_t dims[3] = {static_cast<uint64_t>(H) * 64, static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
  const uint64_t strides[2] = {static_cast<uint64_t>(ld), static_cast<uint64_t>(ld) * T};
  const uint32_t box[3] = {64, static_cast<uint32_t>(box_rows), 1};
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

static long long* g_att_dbg = nullptr;
// head counters of attn_bwd_ws3_kernel: 64 {claim, done} pairs used round-robin by the launches (zero at module load, and every
// launch leaves its pair at zero), so launches on different streams do not share one and graph replays reuse their own
__device__ int g_att_work[2 * 64];
static std::atomic<unsigned> g_att_work_slot{0};
static std::atomic<int> g_fwd_variant{2};   // nvit_attention_fwd_variant: 1 = attn_fwd_kernel (one head per CTA, two CTAs per SM), 2 = attn_fwd_ws_kernel
static std::atomic<int> g_bwd_variant{3};   // nvit_attention_bwd_variant: 1 = attn_bwd_kernel, 2 = attn_bwd_ws_kernel, 3 = attn_bwd_ws3_kernel
#ifdef NVIT_BENCH_HOOKS
extern "C" int nvit_attention_debug(void* dev_buf_256_int64) {   // measurement aid: phase timestamps, see ATT_MARK
  g_att_dbg = static_cast<long long*>(dev_buf_256_int64);
  return NVIT_OK;
}
#endif

extern "C" int nvit_attention_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                  const float* sqk, float sqk_mul, float scale, void* out, int64_t ldo, float* lse, int64_t B,
                                  int64_t H, int64_t T, int64_t D, const float* inv_q, const float* inv_k, int64_t ld_inv_q,
                                  int64_t ld_inv_k, void* stream) {
  NVIT_REQUIRE(q && k && v && out && lse, "nvit_attention_fwd: null argument");
  NVIT_REQUIRE((inv_q == nullptr) == (inv_k == nullptr) && (!inv_q || sqk), "nvit_attention_fwd: inv_q / inv_k go together and need sqk");
  int rc = attn_check("nvit_attention_fwd", B, H, T, D);
  if (rc) return rc;
  NVIT_REQUIRE((ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "nvit_attention_fwd: out must be 16-byte aligned rows");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  const int TPf = (int)((T + 15) / 16 * 16);
  // the persistent kernel takes q / k that need no work in shared memory (pre-normalised, or plain attention) and keeps tiles
  // of exactly TP rows; q / k normalised in place go to the one-head-per-CTA kernel with its 256-row tiles
  const bool ws = g_fwd_variant.load(std::memory_order_relaxed) == 2 && !(sqk != nullptr && inv_q == nullptr) && H <= 256;
  const int in_rows = ws ? TPf : ATT_ROWS;
  if ((rc = make_head_tmap(&p.tq, q, ldq, (int)B, (int)H, (int)T, in_rows))) return rc;
  if ((rc = make_head_tmap(&p.tk, k, ldk, (int)B, (int)H, (int)T, in_rows))) return rc;
  if ((rc = make_head_tmap(&p.tv, v, ldv, (int)B, (int)H, (int)T, in_rows))) return rc;
  if ((rc = make_head_tmap(&p.to, out, ldo, (int)B, (int)H, (int)T, 128))) return rc;
  p.ldo = ldo;
  p.lse = lse;
  p.inv_q = inv_q; p.inv_k = inv_k; p.ld_inv_q = ld_inv_q; p.ld_inv_k = ld_inv_k;
  p.sqk = sqk;
  p.sqk_mul = sqk_mul;
  p.scale = scale;
  p.B = (int)B; p.H = (int)H; p.T = (int)T;
  p.TP = (int)((T + 15) / 16 * 16);
  p.nQ = (int)((T + 127) / 128);
  p.nK = p.nQ;
  p.dbg = g_att_dbg;
  static DeviceOnce once;
  int dev;
  if (once.needed(&dev)) {
    NVIT_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_FWD_SMEM));
    NVIT_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_MAX_SMEM));
    once.mark(dev);
  }
  if (ws) {
    const long long heads = B * H;
    p.h_magic = (H >= 2 && heads * H < (1ll << 32)) ? static_cast<unsigned>((1ull << 32) / static_cast<unsigned long long>(H) + 1ull) : 0u;
    int* work_base = nullptr;
    NVIT_CUDA_CHECK(cudaGetSymbolAddress(reinterpret_cast<void**>(&work_base), g_att_work));
    p.work = work_base + 2 * (g_att_work_slot.fetch_add(1, std::memory_order_relaxed) % 64u);
    const size_t tile = 128ull * TPf;
    const size_t tiles_end = std::max<size_t>(6 * tile, 3 * tile + 16384ull * p.nQ);
    const size_t smem = tiles_end + 256 * 4 + 256;            // two (Q, K, V) sets, the per-head bounds, barriers
    const int grid = (int)(heads < nvit_num_sms() ? heads : nvit_num_sms());
    launch(attn_fwd_ws_kernel, (unsigned)grid, ATTF_THREADS, smem, static_cast<cudaStream_t>(stream), p);
  } else {
    launch(attn_fwd_kernel, (unsigned)(B * H), ATT_FWD_THREADS, ATT_FWD_SMEM, static_cast<cudaStream_t>(stream), p);
  }
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_attention_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                                  const float* sqk, float sqk_mul, float scale, const void* out, const void* dout, int64_t ldo,
                                  const float* lse, void* dq, void* dk, void* dv, int64_t lddq, int64_t lddk, int64_t lddv,
                                  float* dsqk_accum, int64_t B, int64_t H, int64_t T, int64_t D, const float* inv_q, const float* inv_k,
                                  int64_t ld_inv_q, int64_t ld_inv_k, void* stream) {
  NVIT_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv, "nvit_attention_bwd: null argument");
  NVIT_REQUIRE((inv_q == nullptr) == (inv_k == nullptr) && (!inv_q || sqk), "nvit_attention_bwd: inv_q / inv_k go together and need sqk");
  NVIT_REQUIRE((sqk == nullptr) == (dsqk_accum == nullptr), "nvit_attention_bwd: sqk and dsqk go together");
  int rc = attn_check("nvit_attention_bwd", B, H, T, D);
  if (rc) return rc;
  NVIT_REQUIRE(((ldq | ldk | ldo | lddq | lddk | lddv) % 8) == 0, "nvit_attention_bwd: leading dims must be multiples of 8");
  NVIT_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(out) |
                 reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) |
                 reinterpret_cast<uintptr_t>(dv)) & 15) == 0, "nvit_attention_bwd: buffers must be 16-byte aligned");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  const int variant = g_bwd_variant.load(std::memory_order_relaxed);
  const int TP = (int)((T + 15) / 16 * 16);
  // v1 loads whole 256-row boxes; the persistent kernel keeps tiles of exactly TP rows
  const int in_rows = variant == 1 ? ATT_ROWS : TP;
  if ((rc = make_head_tmap(&p.tq, q, ldq, (int)B, (int)H, (int)T, in_rows))) return rc;
  if ((rc = make_head_tmap(&p.tk, k, ldk, (int)B, (int)H, (int)T, in_rows))) return rc;
  if ((rc = make_head_tmap(&p.tv, v, ldv, (int)B, (int)H, (int)T, in_rows))) return rc;
  if ((rc = make_head_tmap(&p.tdo, dout, ldo, (int)B, (int)H, (int)T, in_rows))) return rc;
  if ((rc = make_head_tmap(&p.to, out, ldo, (int)B, (int)H, (int)T, in_rows))) return rc;
  if ((rc = make_head_tmap(&p.tdq, dq, lddq, (int)B, (int)H, (int)T, 128))) return rc;
  if ((rc = make_head_tmap(&p.tdk, dk, lddk, (int)B, (int)H, (int)T, 128))) return rc;
  if ((rc = make_head_tmap(&p.tdv, dv, lddv, (int)B, (int)H, (int)T, 128))) return rc;
  p.ldq = ldq; p.ldk = ldk; p.ldo = ldo; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.lse = const_cast<float*>(lse);
  p.inv_q = inv_q; p.inv_k = inv_k; p.ld_inv_q = ld_inv_q; p.ld_inv_k = ld_inv_k;
  p.sqk = sqk;
  p.dsqk = dsqk_accum;
  p.sqk_mul = sqk_mul;
  p.scale = scale;
  p.B = (int)B; p.H = (int)H; p.T = (int)T;
  p.TP = TP;
  p.nQ = (int)((T + 127) / 128);
  p.nK = p.nQ;
  p.dbg = g_att_dbg;
  static DeviceOnce once;
  int dev;
  if (once.needed(&dev)) {
    NVIT_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWD_SMEM));
    NVIT_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_ws_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_MAX_SMEM));
    NVIT_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_ws3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_MAX_SMEM));
    once.mark(dev);
  }
  if (variant == 1) {
    launch(attn_bwd_kernel, (unsigned)(B * H), ATT_THREADS, ATT_BWD_SMEM, static_cast<cudaStream_t>(stream), p);
  } else {
    // shared memory: (Q, K) x 2 when it fits (the next head's tiles load during the current head), V, dO, two dS^T buffers
    const long long R = 128ll * TP;
    const long long fixed = 2 * ATT_DS_BYTES + att2_array_bytes(TP) + (variant == 3 ? 4 * TP : 0);   // v3: second log-sum-exp row
    p.dbuf = (6 * R + fixed <= ATT2_MAX_SMEM) ? 1 : 0;
    const long long smem = (p.dbuf ? 6 : 4) * R + fixed;
    NVIT_REQUIRE(smem <= ATT2_MAX_SMEM, "nvit_attention_bwd: shared-memory plan does not fit (T = %lld)", (long long)T);
    const long long heads = B * H;
    int grid = (int)(heads < nvit_num_sms() ? heads : nvit_num_sms());
#ifdef NVIT_BENCH_HOOKS
    { static const int ge = getenv("NVIT_ATTN_GRID") ? atoi(getenv("NVIT_ATTN_GRID")) : 0; if (ge > 0 && ge < grid) grid = ge; }   // scripts/attn_bwd_cta_times.py
#endif
    // v3 (epilogue warpgroup) needs both (Q, K) pairs and q / k that are not normalised in place; otherwise v2 runs
    const bool v3 = variant == 3 && p.dbuf && (sqk == nullptr || inv_q != nullptr) && H >= 2 && heads * H < (1ll << 32);   // (h_magic)
    if (v3) {
      p.h_magic = static_cast<unsigned>((1ull << 32) / static_cast<unsigned long long>(H) + 1ull);
      int* work_base = nullptr;
      NVIT_CUDA_CHECK(cudaGetSymbolAddress(reinterpret_cast<void**>(&work_base), g_att_work));
      p.work = work_base + 2 * (g_att_work_slot.fetch_add(1, std::memory_order_relaxed) % 64u);
    }
    if (v3)
      launch(attn_bwd_ws3_kernel, (unsigned)grid, ATT3_THREADS, (size_t)smem, static_cast<cudaStream_t>(stream), p);
    else
      launch(attn_bwd_ws_kernel<8>, (unsigned)grid, 8 * 32 + 32, (size_t)smem, static_cast<cudaStream_t>(stream), p);
  }
  NVIT_CUDA_CHECK(cudaGetLastError());
  return NVIT_OK;
}

extern "C" int nvit_attention_fwd_variant(int variant) {   // tuning switch (include/nvit_b200_tuning.h): same results
  NVIT_REQUIRE(variant == 1 || variant == 2, "nvit_attention_fwd_variant: 1 (one head per CTA, two CTAs per SM) or 2 (persistent, warp-specialised; default)");
  g_fwd_variant.store(variant, std::memory_order_relaxed);
  return NVIT_OK;
}

extern "C" int nvit_attention_bwd_variant(int variant) {   // tuning switch (include/nvit_b200_tuning.h): same results
  NVIT_REQUIRE(variant >= 1 && variant <= 3, "nvit_attention_bwd_variant: 1 (one role, one head per CTA), 2 (persistent, warp-specialised; default) or 3 (2 + epilogue warpgroup)");
  g_bwd_variant.store(variant, std::memory_order_relaxed);
  return NVIT_OK;
}
