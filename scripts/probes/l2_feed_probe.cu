// Probe: how many bytes per cycle can the L2 feed into the shared memory of every SM at once, in the access pattern of
// the persistent GEMM (per CTA and k-block: a 16 KB "A" chunk and a 16 KB "B" chunk into a 6-stage ring, nothing else),
// and what multicast of the B chunk across a cluster of 2 / 4 / 8 CTAs changes.  No tensor cores, no epilogue: the
// consumer only waits for a stage and hands it back.  Answers, for round 2 of the GEMM work:
//   * the feed limit per SM with every SM loading (DESIGN.md section 7 assumes ~42 B/cycle/SM, from B300_MICROARCH.md);
//   * whether A chunks shared by many CTAs at the same time (the n-fastest tile order: 24 tiles on one A row-panel) are
//     cheaper than private ones;
//   * what multicast buys at each cluster size, and how many clusters of that size the GPU keeps resident.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I nvit_b200/csrc -o scripts/probes/l2_feed_probe.bin scripts/probes/l2_feed_probe.cu
// Run:   scripts/probes/l2_feed_probe.bin
#include "common.cuh"
#include <cstdio>
#include <cstdlib>

void nvit_set_error(const char*, ...) {}
int nvit_num_sms() { return 148; }
int nvit_pdl_enabled() { return 0; }
using namespace nvit;

constexpr int STAGES = 6, A_BYTES = 16384, B_BYTES = 16384, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the copy lands at the same offset in every CTA of `mask`, and each of them gets the byte count on its own barrier
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// a_group: consecutive CTAs that read the same A stream (1 = private).  B: one stream per cluster.
template <int CSZ, bool MC>
__global__ void __launch_bounds__(64, 1) feed_kernel(const uint8_t* a_base, size_t a_span, int a_group, const uint8_t* b_base,
                                                     size_t b_span, int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const uint32_t rank = CSZ > 1 ? cluster_ctarank() : 0;
  const int cluster = blockIdx.x / CSZ;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MC ? CSZ : 1);
    }
    fence_barrier_init();
  }
  if (CSZ > 1) cluster_sync_all(); else __syncthreads();
  const long long c0 = clock64();
  const unsigned long long t0 = globaltimer();
  if (threadIdx.x == 0) {                       // producer
    const uint8_t* a_src = a_base + static_cast<size_t>(blockIdx.x / a_group) * a_span;
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* dst = smem + s * STAGE_BYTES;
      mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
      bulk_g2s(dst, a_src + (static_cast<size_t>(it) * A_BYTES) % a_span, A_BYTES, &full[s]);
      const uint8_t* b = b_base + ((static_cast<size_t>(cluster) * 64 + it) * B_BYTES) % b_span;
      if (!MC) {
        bulk_g2s(dst + A_BYTES, b, B_BYTES, &full[s]);
      } else {
        constexpr uint32_t part = B_BYTES / CSZ;
        bulk_g2s_mc(dst + A_BYTES + rank * part, b + rank * part, part, &full[s], static_cast<uint16_t>((1u << CSZ) - 1));
      }
    }
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + (MC ? CSZ : 1)) {   // consumer: lane d hands the stage back to CTA d
    // (round 1 ran this with ONE thread doing the CSZ remote arrives in turn: ~350 cycles each, which bounded the
    // multicast rows of profiles/r01_l2_feed_probe.log at 18 / 13 / 9 B/cycle/SM - they say nothing about multicast itself)
    const uint32_t d = threadIdx.x - 32;
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&full[s], ph);
      if (MC) mbar_arrive_cluster(mapa_shared(smem_u32(&empty[s]), d));
      else mbar_arrive(&empty[s]);
    }
  }
  __syncthreads();
  const long long c1 = clock64();
  const unsigned long long t1 = globaltimer();
  if (CSZ > 1) cluster_sync_all();              // nobody leaves while a peer may still write into it
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = static_cast<unsigned long long>(c1 - c0);
    out[2 * blockIdx.x + 1] = t1 - t0;
  }
}

template <int CSZ, bool MC>
static void run(const uint8_t* a, size_t a_span, int a_group, const uint8_t* b, size_t b_span, int iters, unsigned long long* out_dev, int sms) {
  auto kern = feed_kernel<CSZ, MC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  const int ctas = (sms / CSZ) * CSZ;
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CSZ;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = -1;
  cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best_ms = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, a, a_span, a_group, b, b_span, iters, out_dev);
    cudaEventRecord(e1);
    cudaError_t err2 = cudaDeviceSynchronize();
    if (err != cudaSuccess || err2 != cudaSuccess) {
      printf("cluster %d mc %d a_group %d: FAILED (%s / %s)\n", CSZ, (int)MC, a_group, cudaGetErrorString(err), cudaGetErrorString(err2));
      exit(1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best_ms) best_ms = ms;
  }
  unsigned long long h[2];
  cudaMemcpy(h, out_dev, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes_per_cta = static_cast<double>(iters) * STAGE_BYTES;
  const double mhz = h[1] ? static_cast<double>(h[0]) / static_cast<double>(h[1]) * 1e3 : 0.0;
  printf("cluster %d  %-9s  A shared by %2d CTAs: %7.3f ms  %6.2f TB/s into %3d SMs  %5.1f B/cycle/SM (CTA 0: %llu cycles at %4.0f MHz)  resident clusters <= %d\n",
         CSZ, MC ? "multicast" : "unicast", a_group, best_ms, bytes_per_cta * ctas / (best_ms * 1e-3) / 1e12, ctas,
         bytes_per_cta / static_cast<double>(h[0]), h[0], mhz, max_clusters);
  fflush(stdout);
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t a_span = 256 * 1024, b_span = 16u << 20;
  uint8_t *a, *b;
  unsigned long long* out;
  cudaMalloc(&a, a_span * 160);
  cudaMalloc(&b, b_span);
  cudaMalloc(&out, 2 * 160 * sizeof(unsigned long long));
  cudaMemset(a, 1, a_span * 160);
  cudaMemset(b, 2, b_span);
  const int iters = 3000;
  for (int a_group : {1, 24}) {
    run<1, false>(a, a_span, a_group, b, b_span, iters, out, sms);
    run<2, false>(a, a_span, a_group, b, b_span, iters, out, sms);
    run<2, true>(a, a_span, a_group, b, b_span, iters, out, sms);
    run<4, false>(a, a_span, a_group, b, b_span, iters, out, sms);
    run<4, true>(a, a_span, a_group, b, b_span, iters, out, sms);
    run<8, false>(a, a_span, a_group, b, b_span, iters, out, sms);
    run<8, true>(a, a_span, a_group, b, b_span, iters, out, sms);
  }
  return 0;
}
