// Hand-off latency micro-benchmarks for the attention pipeline (one CTA per SM, cycle counts from clock64 of CTA 0):
//   1. mbarrier ping-pong between two warps, three wait flavours: try_wait with the 10 ms suspend hint (what mbar_wait() uses),
//      try_wait without a hint, test_wait polling;
//   2. tcgen05.mma + tcgen05.commit -> mbarrier -> waiting warp: n MMAs of M128 N128 K16 (n = 0, 1, 5) and 8 TS MMAs of N80;
//   3. the softmax <-> MMA chain of one or two query tiles without any arithmetic: wait S, tcgen05.ld 128 columns, [128 ex2 per thread],
//      tcgen05.st 64 columns, fence, arrive; the MMA warp answers with P V (8 TS MMAs, N = 80) + the next S (5 MMAs, N = 128) + commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ezaudio_b200/csrc -o profiles/micro/handoff profiles/micro/handoff.cu
#include <stdio.h>
#include "common.cuh"
using namespace ezb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

template <int W>
__device__ __forceinline__ void wait_f(uint64_t* bar, uint32_t parity) {
  if (W == 0) { mbar_wait(bar, parity); return; }
  uint32_t ok = 0;
  while (!ok) {
    if (W == 1)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

// ---- 1. ping-pong: warp 0 arrives on b0 and waits b1; warp 1 waits b0 and arrives on b1.  ALL: every lane waits (lane 0 arrives).
template <int W, int ALL>
__global__ void k_pingpong(long long* cyc, int iters) {
  __shared__ uint64_t bars[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  __syncthreads();
  const long long t0 = clock64();
  if (ALL || lane == 0) {
    for (int it = 0; it < iters; ++it) {
      if (warp == 0) {
        if (lane == 0) mbar_arrive(&bars[0]);
        wait_f<W>(&bars[1], it & 1);
        if (ALL) __syncwarp();
      } else {
        wait_f<W>(&bars[0], it & 1);
        if (ALL) __syncwarp();
        if (lane == 0) mbar_arrive(&bars[1]);
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// ---- 2. MMA + commit -> waiter.  warp 0: requester / waiter (all lanes), warp 1: issuer (warp-uniform, elected lane)
template <int W>
__global__ void __launch_bounds__(64, 1) k_mma_commit(long long* cyc, int iters, int n_mma, int ts) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  const uint64_t ad = umma_desc_sw128(smem_u32(smem)), bd = umma_desc_sw128(smem_u32(smem + 16384));
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (warp == 0) {
      if (lane == 0) mbar_arrive(&bars[0]);
      wait_f<W>(&bars[1], it & 1);
      tc_fence_after();
      __syncwarp();
    } else {
      wait_f<W>(&bars[0], it & 1);
      if (elect_one()) {
        if (ts) {
          for (int k = 0; k < n_mma; ++k) umma_bf16_ts(tm + 256, tm + (k & 7) * 8, bd + 2 * (k & 3), umma_idesc_bf16(128, 80), k != 0);
        } else {
          for (int k = 0; k < n_mma; ++k) umma_bf16(tm, ad + 2 * (k & 3), bd + 2 * (k & 3), umma_idesc_bf16(128, 128), k != 0);
        }
        umma_commit(&bars[1]);
      }
      __syncwarp();
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tm);
}

// ---- 3. softmax <-> MMA chain, NG query tiles in flight (softmax group g = warps 4g..4g+3, MMA warp = warp 4 NG), EXP: 128 ex2 per thread
template <int W, int NG, int EXP>
__global__ void __launch_bounds__(128 * NG + 32, 1) k_chain(long long* cyc, float* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);   // s_full[2], p_full[2]
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int g = 0; g < 2; ++g) { mbar_init(&bars[g], 1); mbar_init(&bars[2 + g], 4); }
    mbar_init(&bars[4], 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 4 * NG) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  const long long t0 = clock64();
  float acc = 0.f;
  if (warp == 4 * NG) {
    const uint64_t qd = umma_desc_sw128(smem_u32(smem)), kd = umma_desc_sw128(smem_u32(smem + 16384)), vd = umma_desc_sw128(smem_u32(smem + 32768));
    auto issue_s = [&](int g) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 5; ++k) umma_bf16(tm + g * 128, qd + 2 * (k & 3), kd + 2 * (k & 3), umma_idesc_bf16(128, 128), k != 0);
        umma_commit(&bars[g]);
      }
      __syncwarp();
    };
    for (int g = 0; g < NG; ++g) issue_s(g);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        wait_f<W>(&bars[2 + g], it & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_bf16_ts(tm + 256 + g * 128, tm + g * 128 + k * 8, vd + 2 * (k & 3), umma_idesc_bf16(128, 80), k != 0);
        }
        __syncwarp();
        if (it + 1 < iters) issue_s(g);
      }
    }
    if (elect_one()) umma_commit(&bars[4]);   // drain the tensor pipe before the TMEM allocation is released
    __syncwarp();
    wait_f<0>(&bars[4], 0);
  } else {
    const int g = warp >> 2, lg = warp & 3;
    const uint32_t tS = tm + g * 128 + (static_cast<uint32_t>(lg * 32) << 16);
    for (int it = 0; it < iters; ++it) {
      wait_f<W>(&bars[g], it & 1);
      tc_fence_after();
      uint32_t sr[128];
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld_32x32(tS + q * 32, sr + q * 32);
      tmem_ld_wait();
      uint32_t pk[64];
#pragma unroll
      for (int c = 0; c < 128; c += 2) {
        float a = __uint_as_float(sr[c]), b = __uint_as_float(sr[c + 1]);
        if (EXP) {
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
        }
        acc += a + b;
        pk[c >> 1] = pack_bf16(a, b);
      }
      tmem_st_32x32(tS, pk);
      tmem_st_32x32(tS + 32, pk + 32);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[2 + g]);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 4 * NG) tmem_dealloc<512>(tm);
}

int main() {
  long long* cyc; float* out;
  CK(cudaMalloc(&cyc, 8));
  CK(cudaMalloc(&out, 148 * 512 * 4));
  long long h = 0;
  const int iters = 2000;
  const char* wn[] = {"try_wait + 10 ms hint", "try_wait, no hint", "test_wait polling"};
#define RUN(label, launch, per) do { for (int rep = 0; rep < 2; ++rep) { launch; CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); } \
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("%-88s %8.1f cycles per %s\n", label, (double)h / iters, per); } while (0)
  char lab[256];
  for (int w = 0; w < 3; ++w) {
    snprintf(lab, sizeof lab, "ping-pong, lane 0 of two warps              [%s]", wn[w]);
    if (w == 0) RUN(lab, (k_pingpong<0, 0><<<148, 64>>>(cyc, iters)), "round trip");
    if (w == 1) RUN(lab, (k_pingpong<1, 0><<<148, 64>>>(cyc, iters)), "round trip");
    if (w == 2) RUN(lab, (k_pingpong<2, 0><<<148, 64>>>(cyc, iters)), "round trip");
    snprintf(lab, sizeof lab, "ping-pong, all lanes wait                   [%s]", wn[w]);
    if (w == 0) RUN(lab, (k_pingpong<0, 1><<<148, 64>>>(cyc, iters)), "round trip");
    if (w == 1) RUN(lab, (k_pingpong<1, 1><<<148, 64>>>(cyc, iters)), "round trip");
    if (w == 2) RUN(lab, (k_pingpong<2, 1><<<148, 64>>>(cyc, iters)), "round trip");
  }
  const int SM = 65536 + 1024 + 256;
  CK(cudaFuncSetAttribute(k_mma_commit<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  CK(cudaFuncSetAttribute(k_mma_commit<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  for (int n : {0, 1, 5}) {
    snprintf(lab, sizeof lab, "request -> %d x MMA(M128 N128 K16) + commit -> waiter [%s]", n, wn[0]);
    RUN(lab, (k_mma_commit<0><<<148, 64, SM>>>(cyc, iters, n, 0)), "round trip");
    snprintf(lab, sizeof lab, "request -> %d x MMA(M128 N128 K16) + commit -> waiter [%s]", n, wn[2]);
    RUN(lab, (k_mma_commit<2><<<148, 64, SM>>>(cyc, iters, n, 0)), "round trip");
  }
  snprintf(lab, sizeof lab, "request -> 8 x TS MMA(M128 N80 K16) + commit -> waiter [%s]", wn[0]);
  RUN(lab, (k_mma_commit<0><<<148, 64, SM>>>(cyc, iters, 8, 1)), "round trip");
  snprintf(lab, sizeof lab, "request -> 8 x TS MMA(M128 N80 K16) + commit -> waiter [%s]", wn[2]);
  RUN(lab, (k_mma_commit<2><<<148, 64, SM>>>(cyc, iters, 8, 1)), "round trip");
#define CHAIN(W, NG, EXP) do { CK(cudaFuncSetAttribute(k_chain<W, NG, EXP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); \
    snprintf(lab, sizeof lab, "softmax <-> MMA chain, %d tile(s) in flight, %s [%s]", NG, EXP ? "128 ex2 per thread" : "no arithmetic    ", wn[W]); \
    RUN(lab, (k_chain<W, NG, EXP><<<148, 128 * NG + 32, SM>>>(cyc, out, iters)), NG == 2 ? "block pair" : "block"); } while (0)
  CHAIN(0, 1, 0); CHAIN(2, 1, 0); CHAIN(0, 1, 1); CHAIN(2, 1, 1);
  CHAIN(0, 2, 0); CHAIN(2, 2, 0); CHAIN(0, 2, 1); CHAIN(2, 2, 1); CHAIN(1, 2, 1);
  printf("done\n");
  return 0;
}
