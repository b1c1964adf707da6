// Probe: main-loop rate of the pair-mode (cta_group::2) bf16 GEMM for three tile shapes per CTA pair, with NO epilogue
// (accumulators are simply overwritten by the next tile), K-major operands through 128B-swizzled TMA tiles exactly as in
// nvit_b200/csrc/gemm_tcgen05.cu:
//   shape 0  256 x 256  per CTA and k-block: A 128 rows + B 128 rows (32 KB), 4 MMAs of 256x256x16     128 FLOP per byte
//   shape 1  512 x 256  ("tall")             A 2 x 128 rows + B 128 rows (48 KB), 8 MMAs               170 FLOP per byte
//   shape 2  256 x 512  ("wide")             A 128 rows + B 2 x 128 rows (48 KB), 8 MMAs               170 FLOP per byte
// Question: every SM takes in bulk copies at ~84 GB/s whatever the cluster size
// (l2_feed_probe.cu), which caps shape 0 at ~1580 TFLOP/s; do the 48 KB shapes lift the main loop toward the tensor
// peak at the running clock, and by how much?  The answer decides whether the long-K GEMMs (wgrad, c_fc dgrad,
// mlp_c_proj) get such tiles (they can live without the TMEM double buffer the 512-column accumulators cost).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I nvit_b200/csrc -o scripts/probes/gemm_shape_probe.bin scripts/probes/gemm_shape_probe.cu
// Run:   scripts/probes/gemm_shape_probe.bin [1]      (1: random operands instead of zeros)
#include "common.cuh"
#include <cstdio>
#include <cstdlib>

void nvit_set_error(const char*, ...) {}
int nvit_num_sms() { return 148; }
int nvit_pdl_enabled() { return 0; }
using namespace nvit;

struct alignas(64) ProbeParams {
  CUtensorMap tma_a, tma_b;   // both [rows, K] bf16, K contiguous, box 64 (k) x 128 (rows), 128B swizzle
  int tiles_m, tiles_n, kb_total;
  unsigned long long* out;    // [2 * cta]: cycles, nanoseconds
};

__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int SHAPE>
struct Shape {
  static constexpr int A_SUB = SHAPE == 1 ? 2 : 1;     // 128-row A sub-tiles per CTA and stage
  static constexpr int B_SUB = SHAPE == 2 ? 2 : 1;     // 128-row B sub-tiles per CTA and stage
  static constexpr int SUB_BYTES = 128 * 64 * 2;       // 16 KB
  static constexpr int STAGE_BYTES = (A_SUB + B_SUB) * SUB_BYTES;
  static constexpr int STAGES = 196608 / STAGE_BYTES;  // 6 or 4
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;
  static constexpr int TILE_M = 256 * A_SUB, TILE_N = 256 * B_SUB;
  static constexpr int TMEM_COLS = 512;
};

template <int SHAPE>
__global__ void __launch_bounds__(64, 1) shape_kernel(const __grid_constant__ ProbeParams p) {
  using S = Shape<SHAPE>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* done_bar = empty_bar + S::STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  const uint32_t cta_rank = cluster_ctarank();
  const int unit0 = (int)cluster_id_x(), unit_stride = (int)cluster_count_x();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_units = p.tiles_m * p.tiles_n;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
    for (int i = 0; i < S::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_cg2(tmem_ptr, S::TMEM_COLS); tmem_relinquish_cg2(); }
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const long long c0 = clock64();
  const unsigned long long t0 = globaltimer();

  if (warp == 0) {
    if (lane == 0) {                     // TMA producer (both CTAs; bytes counted on the rank-0 CTA's barrier)
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = smem_u32(full_bar);
      for (int u = unit0; u < total_units; u += unit_stride) {
        const int n_blk = u % p.tiles_n, m_tile = u / p.tiles_n;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* dst = smem + stage * S::STAGE_BYTES;
          const uint32_t fb = mapa_shared(full0 + stage * 8, 0);
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
          const int k0 = kb * 64;
#pragma unroll
          for (int a = 0; a < S::A_SUB; ++a)      // sub-tile a of the pair tile: rows (m_tile * 2 A_SUB + 2 a + rank) * 128
            tma_load_2d_cg2(&p.tma_a, fb, dst + a * S::SUB_BYTES, k0, (m_tile * 2 * S::A_SUB + 2 * a + (int)cta_rank) * 128);
#pragma unroll
          for (int b = 0; b < S::B_SUB; ++b)
            tma_load_2d_cg2(&p.tma_b, fb, dst + (S::A_SUB + b) * S::SUB_BYTES, k0, (n_blk * 2 * S::B_SUB + 2 * b + (int)cta_rank) * 128);
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (cta_rank == 0) {            // MMA issuer: warp 1 of the rank-0 CTA
    constexpr uint32_t idesc = umma_idesc_bf16(256, 256, 0, 0);
    const uint32_t smem0 = smem_u32(smem);
    const uint64_t d_base = umma_smem_desc(smem0, 16, 1024);
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    int stage = 0;
    uint32_t phase = 0;
    for (int u = unit0; u < total_units; u += unit_stride) {
      for (int kb = 0; kb < p.kb_total; ++kb) {
        mbar_wait_a(full0 + stage * 8, phase);
        tc_fence_after_sync();
        const uint64_t ds = d_base + static_cast<uint64_t>((stage * S::STAGE_BYTES) >> 4);
        const uint32_t first = kb > 0 ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int a = 0; a < S::A_SUB; ++a)
#pragma unroll
              for (int b = 0; b < S::B_SUB; ++b) {
                const uint64_t da = ds + ((a * S::SUB_BYTES) >> 4) + k * 2;
                const uint64_t db = ds + (((S::A_SUB + b) * S::SUB_BYTES) >> 4) + k * 2;
                const uint32_t acc_col = tmem_base + (a * S::B_SUB + b) * 256;
                if (k == 0) umma_bf16_ss_cg2(acc_col, da, db, idesc, first);
                else umma_bf16_ss_cg2_acc(acc_col, da, db, idesc);
              }
          }
          umma_commit_cg2_a(empty0 + stage * 8);
        }
        __syncwarp();
        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (elect_one()) umma_commit_cg2_a(smem_u32(done_bar));   // everything issued has retired when this lands (arrives in both CTAs; rank 0 waits)
    __syncwarp();
    mbar_wait(done_bar, 0);
  }
  __syncthreads();
  const long long c1 = clock64();
  const unsigned long long t1 = globaltimer();
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after_sync(); tmem_dealloc_cg2(tmem_base, S::TMEM_COLS); }
  if (threadIdx.x == 0) { p.out[2 * blockIdx.x] = (unsigned long long)(c1 - c0); p.out[2 * blockIdx.x + 1] = t1 - t0; }
}

// pseudo-random bf16 in (-1, 1): with real operand bits the tensor pipe draws its real power, so the clock (and the rate)
// is what a GEMM on activations sees; zeros show the pipe's rate at the highest clock
__global__ void fill_random(uint16_t* p, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const float v = (float)(h & 0xffff) / 32768.f - 1.f;
    const __nv_bfloat16 b = __float2bfloat16(v);
    p[i] = *reinterpret_cast<const uint16_t*>(&b);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(PFN_encodeTiled fn, CUtensorMap* m, void* base, uint64_t rows, uint64_t K) {
  cuuint64_t dims[2] = {K, rows}, strides[1] = {K * 2};
  cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
}

template <int SHAPE>
static void run(PFN_encodeTiled fn, void* A, void* B, int M, int N, int K, unsigned long long* out_dev, int sms) {
  using S = Shape<SHAPE>;
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  make_map(fn, &p.tma_a, A, M, K);
  make_map(fn, &p.tma_b, B, N, K);
  p.tiles_m = M / S::TILE_M;
  p.tiles_n = N / S::TILE_N;
  p.kb_total = K / 64;
  p.out = out_dev;
  auto kern = shape_kernel<SHAPE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((sms / 2) * 2);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = S::SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, p);
    cudaEventRecord(e1);
    cudaError_t err2 = cudaDeviceSynchronize();
    if (err != cudaSuccess || err2 != cudaSuccess) {
      printf("shape %d M %d N %d K %d: FAILED (%s / %s)\n", SHAPE, M, N, K, cudaGetErrorString(err), cudaGetErrorString(err2));
      exit(1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  unsigned long long h[2];
  cudaMemcpy(h, out_dev, sizeof(h), cudaMemcpyDeviceToHost);
  const double flop = 2.0 * M * N * K;
  const double bytes = (double)p.tiles_m * p.tiles_n * p.kb_total * 2.0 * S::STAGE_BYTES;    // taken in by all SMs
  const double mhz = h[1] ? (double)h[0] / (double)h[1] * 1e3 : 0.0;
  printf("shape %d (%3d x %3d per pair, %d stages of %d KB)  M %6d N %5d K %5d: %8.3f ms  %7.1f TFLOP/s  intake %5.2f TB/s = %5.1f B/cycle/SM at %4.0f MHz\n",
         SHAPE, S::TILE_M, S::TILE_N, S::STAGES, S::STAGE_BYTES / 1024, M, N, K, best, flop / (best * 1e-3) / 1e12,
         bytes / (best * 1e-3) / 1e12, bytes / ((sms / 2) * 2) / (best * 1e-3) / (mhz * 1e6), mhz);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const bool random_fill = argc > 1 && atoi(argv[1]) != 0;     // gemm_shape_probe.bin 1: random operands (real power draw)
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  PFN_encodeTiled fn = reinterpret_cast<PFN_encodeTiled>(fnp);
  const int M = 50176, NMAX = 6144, KMAX = 6144;
  void *A, *B;
  unsigned long long* out;
  cudaMalloc(&A, (size_t)M * KMAX * 2);
  cudaMalloc(&B, (size_t)NMAX * KMAX * 2);
  cudaMalloc(&out, 2 * 160 * sizeof(unsigned long long));
  cudaMemset(A, 0, (size_t)M * KMAX * 2);          // zeros: the rate does not depend on the values, and power is lowest
  cudaMemset(B, 0, (size_t)NMAX * KMAX * 2);
  if (random_fill) {
    fill_random<<<148 * 8, 256>>>(static_cast<uint16_t*>(A), (size_t)M * KMAX, 1u);
    fill_random<<<148 * 8, 256>>>(static_cast<uint16_t*>(B), (size_t)NMAX * KMAX, 2u);
    cudaDeviceSynchronize();
  }
  printf("operands: %s\n", random_fill ? "pseudo-random bf16 in (-1, 1)" : "zeros");
  // the step's shapes, forward orientation: c_fc (N 6144, K 768), mlp_c_proj (N 768 -> wide needs a multiple of 512: 1024), long K
  for (int rep = 0; rep < 2; ++rep) {
    run<0>(fn, A, B, M, 6144, 768, out, sms);
    run<1>(fn, A, B, M, 6144, 768, out, sms);
    run<2>(fn, A, B, M, 6144, 768, out, sms);
    run<0>(fn, A, B, M, 1024, 3072, out, sms);
    run<1>(fn, A, B, M, 1024, 3072, out, sms);
    run<2>(fn, A, B, M, 1024, 3072, out, sms);
  }
  return 0;
}
