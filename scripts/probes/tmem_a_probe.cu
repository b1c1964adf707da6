// Probe: does tcgen05.mma accept its A operand from tensor memory as bf16 pairs packed two per 32-bit column
// (lane = row, 8 columns per K = 16 step)?  D1 = A B^T with A from shared memory, D2 = the same with A from TMEM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I nvit_b200/csrc -o gpurun_out/tmem_a_probe scripts/probes/tmem_a_probe.cu
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

namespace nvit { int nvit_num_sms() { return 148; } void nvit_set_error(const char*, ...) {} }
using namespace nvit;

__device__ __forceinline__ uint32_t sw128(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

constexpr int K = 64, N = 64;
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D1, float* D2) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = sm;            // 128 rows x 128 B
  uint8_t* sB = sm + 16384;    // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384 + 8192);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sA + sw128(tid, c)) = *reinterpret_cast<const uint4*>(A + tid * K + c * 8);
  if (tid < N) for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sB + sw128(tid, c)) = *reinterpret_cast<const uint4*>(B + tid * K + c * 8);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tptr, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = *tptr;
  const uint32_t lane_base = tb + (static_cast<uint32_t>(warp * 32) << 16);
  // A row of this thread into TMEM columns [128, 160): pairs (k, k+1) packed low/high
  for (int c = 0; c < 4; ++c) {
    uint32_t r[8];
    const uint4 lo = *reinterpret_cast<const uint4*>(A + tid * K + c * 16), hi = *reinterpret_cast<const uint4*>(A + tid * K + c * 16 + 8);
    r[0] = lo.x; r[1] = lo.y; r[2] = lo.z; r[3] = lo.w; r[4] = hi.x; r[5] = hi.y; r[6] = hi.z; r[7] = hi.w;
    tmem_st_32x32b_x8(lane_base + 128 + c * 8, r);
  }
  tmem_wait_st();
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after_sync();
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t da = umma_smem_desc(smem_u32(sA), 16, 1024), db = umma_smem_desc(smem_u32(sB), 16, 1024);
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tb + 0, da + 2 * k, db + 2 * k, idesc, k > 0);
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tb + 64, tb + 128 + 8 * k, db + 2 * k, idesc, k > 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after_sync();
  for (int c = 0; c < 4; ++c) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(lane_base + c * 16, r);
    tmem_wait_ld();
    for (int e = 0; e < 16; ++e) D1[tid * N + c * 16 + e] = __uint_as_float(r[e]);
    tmem_ld_32x32b_x16(lane_base + 64 + c * 16, r);
    tmem_wait_ld();
    for (int e = 0; e < 16; ++e) D2[tid * N + c * 16 + e] = __uint_as_float(r[e]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tb, 256); }
}

int main() {
  std::vector<__nv_bfloat16> hA(128 * K), hB(N * K);
  std::vector<float> fA(128 * K), fB(N * K);
  srand(1);
  for (int i = 0; i < 128 * K; ++i) { float v = (rand() % 17 - 8) / 8.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 4.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float *d1, *d2;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&d1, 128 * N * 4); cudaMalloc(&d2, 128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(d1, 0, 128 * N * 4); cudaMemset(d2, 0, 128 * N * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(dA, dB, d1, d2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> h1(128 * N), h2(128 * N);
  cudaMemcpy(h1.data(), d1, h1.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h2.data(), d2, h2.size() * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
    float ref = 0; for (int k = 0; k < K; ++k) ref += fA[m * K + k] * fB[n * K + k];
    e1 = fmax(e1, fabs(h1[m * N + n] - ref)); e2 = fmax(e2, fabs(h2[m * N + n] - ref));
  }
  printf("max |D_smemA - ref| = %g   max |D_tmemA - ref| = %g\n", e1, e2);
  printf("sample ref-side D1[0][0..3] = %g %g %g %g ; D2 = %g %g %g %g\n", h1[0], h1[1], h1[2], h1[3], h2[0], h2[1], h2[2], h2[3]);
  return 0;
}
