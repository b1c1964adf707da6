// Probe: tensor-memory read rate of tcgen05.ld (32x32b, x16 / x32 per instruction) per SM, with 4 / 8 / 16 warps reading
// their own lane quarter back to back, and the same with 16 ex2.approx per 16 columns in between (the softmax pass of the
// attention forward kernel: is it the SFU or the TMEM read path that takes 4.4 k cycles per 128 x 208 tile pair?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I nvit_b200/csrc -o gpurun_out/tmem_read_bw scripts/probes/tmem_read_bw.cu
#include "common.cuh"
#include <cstdio>
#include <cstdlib>

namespace nvit { int nvit_num_sms() { return 148; } void nvit_set_error(const char*, ...) {} }
using namespace nvit;

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ uint32_t cvt2(float a, float b) {
  uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r;
}

// mode 4: the softmax pass as the attention forward runs it: ld x16, 16 ex2, pack to bf16 pairs, st x8 over read columns, per chunk
// mode 5: the same with the stores of four chunks issued together (32 registers of P held back)
// mode 0: x16 loads only; 1: x32 loads only; 2: x16 loads + 16 ex2 each; 3: 16 ex2 per step without loads
template <int MODE>
__global__ void __launch_bounds__(512) probe(long long* out, float* sink, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f, ac[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float fb = -(float)(it & 63);
    if (MODE == 6) {
      const f32x2 sl2x2 = pack2(1.25f, 1.25f), nm2 = pack2(fb, fb);
      f32x2 sum2 = pack2(0.f, 0.f);
      for (int c0 = 0; c0 < 13; c0 += 2) {
        uint32_t r[2][16];
        tmem_ld_32x32b_x16(base + c0 * 16, r[0]);
        if (c0 + 1 < 13) tmem_ld_32x32b_x16(base + c0 * 16 + 16, r[1]);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int c = c0 + k;
          if (c < 13) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float a0, a1;
              unpack2(fma2(pack2(__uint_as_float(r[k][2 * e]), __uint_as_float(r[k][2 * e + 1])), sl2x2, nm2), a0, a1);
              const f32x2 pv = pack2(ex2a(a0), ex2a(a1));
              sum2 = add2(sum2, pv);
              pk[e] = f32x2_to_bf16x2(pv);
            }
            st8(base + c * 8, pk);
          }
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      float t0, t1;
      unpack2(sum2, t0, t1);
      ac[0] += t0 + t1;
      continue;
    }
    if (MODE == 4 || MODE == 5) {
      uint32_t pk[4][8];
#pragma unroll
      for (int c = 0; c < 13; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(base + c * 16, r);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float p0 = ex2a(fminf(__uint_as_float(r[2 * e]), 0.f) + fb), p1 = ex2a(fminf(__uint_as_float(r[2 * e + 1]), 0.f) + fb);
          ac[e & 3] += p0 + p1;
          pk[c & 3][e] = cvt2(p0, p1);
        }
        if (MODE == 4) {
          st8(base + c * 8, pk[c & 3]);
        } else if ((c & 3) == 3 || c == 12) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k <= (c & 3)) st8(base + ((c & ~3) + k) * 8, pk[k]);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      continue;
    }
#pragma unroll
    for (int c = 0; c < 13; ++c) {                    // 13 x 16 = 208 columns: one row of scores
      uint32_t r[32];
      if (MODE == 0 || MODE == 2) {
        tmem_ld_32x32b_x16(base + c * 16, r);
        tmem_wait_ld();
      } else if (MODE == 1) {
        tmem_ld_32x32b_x32(base + (c & 7) * 32, r);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(fb - (float)(c * 16 + e) * 0.25f);
      }
      if (MODE >= 2) {
#pragma unroll
        for (int e = 0; e < 16; ++e) ac[e & 3] += ex2a(__uint_as_float(r[e]));
      } else {
        acc += __uint_as_float(r[0]) + __uint_as_float(r[MODE == 1 ? 31 : 15]);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  acc += ac[0] + ac[1] + ac[2] + ac[3];
  if (acc == 12345.678f) sink[0] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

template <int MODE>
static void run(const char* what, int warps, int iters, long long* d_out, float* d_sink) {
  probe<MODE><<<148, warps * 32>>>(d_out, d_sink, iters);
  probe<MODE><<<148, warps * 32>>>(d_out, d_sink, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double cols = 13.0 * (MODE == 1 ? 32 : 16);
  const double bytes = (double)warps * iters * cols * 32 * 4;
  printf("%-28s warps %2d: %9.0f cycles, %6.1f B/cycle/SM, %5.2f cycles per warp per 16 columns (%s)\n", what, warps, avg, bytes / avg,
         avg / (iters * cols / 16.0), cudaGetErrorString(cudaGetLastError()));
}

// The kernel's pass (mode 6) on `wc` warps while `blockDim.x / 32 - wc` further warps poll an mbarrier that has not fired, the way
// the MMA warp, the agent and the other warpgroup wait inside attn_fwd_ws_kernel (mbar_wait: try_wait + clock64 time-out check).
__global__ void __launch_bounds__(512) probe_spin(long long* out, float* sink, int iters, int wc) {
  __shared__ uint32_t tptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  if (warp >= wc) {
    mbar_wait(&bar, 0);
  } else {
    named_bar_sync(1, wc * 32);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const float fb = -(float)(it & 63);
      const f32x2 sl2x2 = pack2(1.25f, 1.25f), nm2 = pack2(fb, fb);
      f32x2 sum2 = pack2(0.f, 0.f);
      for (int c0 = 0; c0 < 13; c0 += 2) {
        uint32_t r[2][16];
        tmem_ld_32x32b_x16(base + c0 * 16, r[0]);
        if (c0 + 1 < 13) tmem_ld_32x32b_x16(base + c0 * 16 + 16, r[1]);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int c = c0 + k;
          if (c < 13) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float a0, a1;
              unpack2(fma2(pack2(__uint_as_float(r[k][2 * e]), __uint_as_float(r[k][2 * e + 1])), sl2x2, nm2), a0, a1);
              const f32x2 pv = pack2(ex2a(a0), ex2a(a1));
              sum2 = add2(sum2, pv);
              pk[e] = f32x2_to_bf16x2(pv);
            }
            st8(base + c * 8, pk);
          }
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      float t0f, t1f;
      unpack2(sum2, t0f, t1f);
      acc += t0f + t1f;
    }
    named_bar_sync(1, wc * 32);
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; mbar_arrive(&bar); }
  }
  if (acc == 12345.678f) sink[0] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

static void run_spin(int wc, int spinners, int iters, long long* d_out, float* d_sink) {
  for (int rep = 0; rep < 2; ++rep) probe_spin<<<148, (wc + spinners) * 32>>>(d_out, d_sink, iters, wc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("kernel's pass, %d warps + %d polling warps: %9.0f cycles, %6.2f cycles per warp per 16 columns (%s)\n", wc, spinners, avg,
         avg / (iters * 13.0), cudaGetErrorString(cudaGetLastError()));
}

// Variants of the pass with a ROLLED loop (the kernel's chunk count is a run-time value):
//  0: as the kernel has it: ld pair, wait, work + st per chunk
//  1: the next pair's loads issued BEFORE this pair's stores (two register sets), stores at the end of the iteration
//  2: as 0 with "#pragma unroll 1" lifted (the compiler may pipeline across the 7 iterations)
template <int V>
__global__ void __launch_bounds__(512) probe_pass(long long* out, float* sink, int iters, int nch) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float fb = -(float)(it & 63);
    const f32x2 sl2x2 = pack2(1.25f, 1.25f), nm2 = pack2(fb, fb);
    f32x2 sum2 = pack2(0.f, 0.f);
    auto work1 = [&](const uint32_t (&r)[16], uint32_t (&pk)[8]) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a0, a1;
        unpack2(fma2(pack2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1])), sl2x2, nm2), a0, a1);
        const f32x2 pv = pack2(ex2a(a0), ex2a(a1));
        sum2 = add2(sum2, pv);
        pk[e] = f32x2_to_bf16x2(pv);
      }
    };
    if (V == 0 || V == 2) {
#pragma unroll 1
      for (int c0 = 0; c0 < nch; c0 += 2) {
        uint32_t r[2][16], pk[8];
        tmem_ld_32x32b_x16(base + c0 * 16, r[0]);
        if (c0 + 1 < nch) tmem_ld_32x32b_x16(base + c0 * 16 + 16, r[1]);
        tmem_wait_ld();
        work1(r[0], pk);
        st8(base + c0 * 8, pk);
        if (c0 + 1 < nch) { work1(r[1], pk); st8(base + c0 * 8 + 8, pk); }
      }
    } else {
      uint32_t ra[2][16], rb[2][16], pa[8], pb[8];
      tmem_ld_32x32b_x16(base, ra[0]);
      tmem_ld_32x32b_x16(base + 16, ra[1]);
#pragma unroll 1
      for (int c0 = 0; c0 < nch; c0 += 4) {
        tmem_wait_ld();
        if (c0 + 2 < nch) tmem_ld_32x32b_x16(base + (c0 + 2) * 16, rb[0]);       // next pair: before this pair's stores
        if (c0 + 3 < nch) tmem_ld_32x32b_x16(base + (c0 + 3) * 16, rb[1]);
        work1(ra[0], pa);
        if (c0 + 1 < nch) work1(ra[1], pb);
        st8(base + c0 * 8, pa);
        if (c0 + 1 < nch) st8(base + c0 * 8 + 8, pb);
        if (c0 + 2 < nch) {
          tmem_wait_ld();
          if (c0 + 4 < nch) tmem_ld_32x32b_x16(base + (c0 + 4) * 16, ra[0]);
          if (c0 + 5 < nch) tmem_ld_32x32b_x16(base + (c0 + 5) * 16, ra[1]);
          work1(rb[0], pa);
          if (c0 + 3 < nch) work1(rb[1], pb);
          st8(base + (c0 + 2) * 8, pa);
          if (c0 + 3 < nch) st8(base + (c0 + 3) * 8, pb);
        }
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    float t0f, t1f;
    unpack2(sum2, t0f, t1f);
    acc += t0f + t1f;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 12345.678f) sink[0] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

template <int V>
static void run_pass(const char* what, int warps, int iters, long long* d_out, float* d_sink) {
  for (int rep = 0; rep < 2; ++rep) probe_pass<V><<<148, warps * 32>>>(d_out, d_sink, iters, 13);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("%-44s warps %d: %9.0f cycles, %6.2f cycles per warp per 16 columns (%s)\n", what, warps, avg, avg / (iters * 13.0),
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 4);
  const int iters = 200;
  for (int w : {4, 8, 16}) run<0>("ld x16, wait each", w, iters, d_out, d_sink);
  for (int w : {4, 8, 16}) run<1>("ld x32, wait each", w, iters, d_out, d_sink);
  for (int w : {4, 8, 16}) run<2>("ld x16 + 16 ex2", w, iters, d_out, d_sink);
  for (int w : {4, 8, 16}) run<3>("16 ex2 only (bytes nominal)", w, iters, d_out, d_sink);
  for (int w : {4, 8}) run<4>("ld x16 + 16 ex2 + cvt + st x8", w, iters, d_out, d_sink);
  for (int w : {4, 8}) run<5>("same, stores 4 chunks at a time", w, iters, d_out, d_sink);
  for (int w : {4, 8}) run<6>("the kernel's pass, verbatim", w, iters, d_out, d_sink);
  for (int w : {4, 8}) run_pass<0>("rolled loop, ld - wait - work - st", w, iters, d_out, d_sink);
  for (int w : {4, 8}) run_pass<1>("rolled loop, next loads before the stores", w, iters, d_out, d_sink);
  for (int sp : {0, 2, 6}) run_spin(4, sp, iters, d_out, d_sink);
  for (int sp : {0, 2}) run_spin(8, sp, iters, d_out, d_sink);
  return 0;
}
