// Micro-benchmarks that decide the attention softmax design (profiles/r2/microbench.txt):
//   1. exp2 throughput per SM: ex2.approx.ftz.f32 vs ex2.approx.ftz.bf16x2 vs ex2.approx.f16x2 (is the packed form one MUFU op?)
//   2. FFMA vs fma.rn.f32x2 issue throughput
//   3. tcgen05.ld / tcgen05.st throughput with 4 and 8 warps (128 columns of fp32 per thread per pass)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/micro/microbench profiles/micro/microbench.cu
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(1024) k_ex2(float* out, int iters, long long* cyc) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xb800b800u + threadIdx.x + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (MODE == 4) {
        unsigned long long v = ((unsigned long long)__float_as_uint(a[i]) << 32) | h[i];
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
        a[i] = __uint_as_float((uint32_t)(v >> 32)); h[i] = (uint32_t)v;
      }
      if (MODE == 5) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0; uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += a[i]; x ^= h[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(x);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: tcgen05.ld 32x32b.x32 x4 (128 columns) + wait per pass; MODE 1: tcgen05.st x32 x2 (64 columns = a bf16 P row) + wait
// MODE 2: tcgen05.ld 32x32b.x64 x2; MODE 3: ld 128 columns, no wait between passes except at the end
template <int MODE>
__global__ void __launch_bounds__(256) k_tmem(float* out, int iters, long long* cyc, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) r[i] = threadIdx.x + i;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0 || MODE == 3) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                       : "=r"(r[q * 32 + 0]), "=r"(r[q * 32 + 1]), "=r"(r[q * 32 + 2]), "=r"(r[q * 32 + 3]), "=r"(r[q * 32 + 4]), "=r"(r[q * 32 + 5]), "=r"(r[q * 32 + 6]), "=r"(r[q * 32 + 7]),
                         "=r"(r[q * 32 + 8]), "=r"(r[q * 32 + 9]), "=r"(r[q * 32 + 10]), "=r"(r[q * 32 + 11]), "=r"(r[q * 32 + 12]), "=r"(r[q * 32 + 13]), "=r"(r[q * 32 + 14]), "=r"(r[q * 32 + 15]),
                         "=r"(r[q * 32 + 16]), "=r"(r[q * 32 + 17]), "=r"(r[q * 32 + 18]), "=r"(r[q * 32 + 19]), "=r"(r[q * 32 + 20]), "=r"(r[q * 32 + 21]), "=r"(r[q * 32 + 22]), "=r"(r[q * 32 + 23]),
                         "=r"(r[q * 32 + 24]), "=r"(r[q * 32 + 25]), "=r"(r[q * 32 + 26]), "=r"(r[q * 32 + 27]), "=r"(r[q * 32 + 28]), "=r"(r[q * 32 + 29]), "=r"(r[q * 32 + 30]), "=r"(r[q * 32 + 31])
                       : "r"(base + q * 32) : "memory");
        if (MODE == 0) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          acc += __uint_as_float(r[0] ^ r[31] ^ r[64] ^ r[127]);   // static indices: a dynamic index would push r[] to local memory
        }
      }
      if (MODE == 1) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                       ::"r"(base + q * 32), "r"(r[q * 32 + 0]), "r"(r[q * 32 + 1]), "r"(r[q * 32 + 2]), "r"(r[q * 32 + 3]), "r"(r[q * 32 + 4]), "r"(r[q * 32 + 5]), "r"(r[q * 32 + 6]), "r"(r[q * 32 + 7]),
                         "r"(r[q * 32 + 8]), "r"(r[q * 32 + 9]), "r"(r[q * 32 + 10]), "r"(r[q * 32 + 11]), "r"(r[q * 32 + 12]), "r"(r[q * 32 + 13]), "r"(r[q * 32 + 14]), "r"(r[q * 32 + 15]),
                         "r"(r[q * 32 + 16]), "r"(r[q * 32 + 17]), "r"(r[q * 32 + 18]), "r"(r[q * 32 + 19]), "r"(r[q * 32 + 20]), "r"(r[q * 32 + 21]), "r"(r[q * 32 + 22]), "r"(r[q * 32 + 23]),
                         "r"(r[q * 32 + 24]), "r"(r[q * 32 + 25]), "r"(r[q * 32 + 26]), "r"(r[q * 32 + 27]), "r"(r[q * 32 + 28]), "r"(r[q * 32 + 29]), "r"(r[q * 32 + 30]), "r"(r[q * 32 + 31])
                       : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
    }
    if (MODE == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  for (int i = 0; i < 128; ++i) acc += __uint_as_float(r[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 1024 * sizeof(float)));
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  long long h[148];
  const int iters = 2000;
  const char* names[] = {"ex2.approx.ftz.f32", "ex2.approx.ftz.bf16x2", "ex2.approx.f16x2", "fma.rn.f32", "fma.rn.f32x2", "tanh.approx.f32"};
  for (int mode = 0; mode < 6; ++mode) {
    for (int threads : {128, 256, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k_ex2<0><<<148, threads>>>(out, iters, cyc);
        if (mode == 1) k_ex2<1><<<148, threads>>>(out, iters, cyc);
        if (mode == 2) k_ex2<2><<<148, threads>>>(out, iters, cyc);
        if (mode == 3) k_ex2<3><<<148, threads>>>(out, iters, cyc);
        if (mode == 4) k_ex2<4><<<148, threads>>>(out, iters, cyc);
        if (mode == 5) k_ex2<5><<<148, threads>>>(out, iters, cyc);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
      const double instr = (double)iters * 8 * threads;   // thread-level instructions per CTA (= per SM)
      printf("%-24s threads/SM %4d: %.2f thread-instr/clk/SM (x2 results for packed forms)\n", names[mode], threads, instr / (double)h[0]);
    }
  }
  const char* tn[] = {"tcgen05.ld x32 x4 + wait (128 cols fp32)", "tcgen05.st x32 x2 + wait (64 cols)", "-", "tcgen05.ld x32 x4 back-to-back (one wait at the end)"};
  for (int mode : {0, 1, 3}) {
    for (int nw : {1, 4, 8}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k_tmem<0><<<148, 256>>>(out, 500, cyc, nw);
        if (mode == 1) k_tmem<1><<<148, 256>>>(out, 500, cyc, nw);
        if (mode == 3) k_tmem<3><<<148, 256>>>(out, 500, cyc, nw);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
      const double bytes = 500.0 * nw * 32 * (mode == 1 ? 64 : 128) * 4;
      printf("%-55s warps %d: %.1f cycles per pass, %.1f B/clk/SM\n", tn[mode], nw, (double)h[0] / 500.0, bytes / (double)h[0]);
    }
  }
  printf("done\n");
  return 0;
}
