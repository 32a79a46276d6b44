// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (bf16 in, fp32 accumulate in TMEM)
//
//   warp 0      TMA producer   : cp.async.bulk.tensor tiles of A (128 x 64) and W (BN x 64) into a STAGES-deep smem ring
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per stage,
//                                tcgen05.commit releases the smem slot / publishes the accumulator
//   warps 2..5  epilogue       : tcgen05.ld the 128 x BN fp32 accumulator (thread == output row), fused epilogue, global stores
//   TMEM holds two accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the mainloop of tile i+1.
//
// A can also be addressed as an implicit-GEMM operand of a 1-D convolution over channels-last activations
// [B, T, C]: k-block kb -> tap = kb / cin_blocks, rows shifted by (tap - center) * dilation with TMA zero fill at
// the clip edges (used by the Oobleck decoder: stable_vae/models/autoencoders.py:38-113).
//
// Reference ops this kernel replaces: every nn.Linear on the DiT step (src/models/utils/attention.py:127-129,148;
// src/models/utils/modules.py:266,366; src/models/blocks.py:101; src/models/udit.py:94-97) and the VAE's
// Conv1d / ConvTranspose1d (stable_vae/models/autoencoders.py:46-52,97-99,167,183).
#pragma once
#include "common.cuh"

namespace ezb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

struct GemmShape {
  int M, N;
  int num_k_blocks;    // K / 64 (rounded up; TMA zero-fills the tail)
  int num_m_tiles, num_n_tiles;
  // implicit-conv addressing of A (taps == 0 -> plain 2-D A[M,K])
  int taps, center, dilation, cin_blocks, T, tiles_per_batch;
};

enum { ACT_NONE = 0, ACT_SILU = 1, ACT_SNAKE = 2 };

// out_f32 receives the pre-activation value, out_bf16 the post-activation one (either may be null).
struct EpiLinearParams {
  const float* bias;       // [N] or [bias_mod]
  int bias_mod;            // 0: bias[col]; >0: bias[col % bias_mod] (conv-transpose phases share one bias)
  const float* resid;      // optional residual input, row stride ldr
  int ldr;
  const float* gate;       // optional: v = resid + (1 - gate[b, col]) * v, b = row / rows_per_batch
  int gate_bstride;
  int rows_per_batch;
  float* out_f32;
  int ld32;
  __nv_bfloat16* out_bf16;
  int ld16;
  int split_stride;        // >0: parity mode, write [hi | lo | hi] at col, col+s, col+2s
  int act;
  const float* act_a;      // snake: exp(alpha)[c], c = col % bias_mod (or col)
  const float* act_b;      // snake: 1 / (exp(beta)[c] + 1e-9)
  float out_scale;         // v = (acc + bias) * out_scale (before residual); 0 is treated as 1
  int phase_cols;          // >0 (conv-transpose): bf16 column = (col / phase_cols) * phase_ld16 + col % phase_cols
  int phase_ld16;
};

__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* p, int split_stride, float a, float b, float c, float d) {
  uint2 hi = make_uint2(pack_bf16(a, b), pack_bf16(c, d));
  *reinterpret_cast<uint2*>(p) = hi;
  if (split_stride > 0) {
    const __nv_bfloat162 h0 = *reinterpret_cast<__nv_bfloat162*>(&hi.x), h1 = *reinterpret_cast<__nv_bfloat162*>(&hi.y);
    uint2 lo = make_uint2(pack_bf16(a - __low2float(h0), b - __high2float(h0)), pack_bf16(c - __low2float(h1), d - __high2float(h1)));
    *reinterpret_cast<uint2*>(p + split_stride) = lo;
    *reinterpret_cast<uint2*>(p + 2 * split_stride) = hi;
  }
}

template <int BN>
struct EpiLinear {
  using Params = EpiLinearParams;
  static __device__ __forceinline__ void run(const Params& ep, uint32_t taddr_row, int row, int M, int n0, int N) {
    const bool row_ok = row < M;
    const int b = (ep.gate != nullptr && row_ok) ? row / ep.rows_per_batch : 0;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      if (n0 + c >= N) break;  // warp-uniform
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(taddr_row + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int col = n0 + c + j;
        if (col >= N || !row_ok) break;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = __uint_as_float(r[j + e]);
        if (ep.bias != nullptr) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] += ep.bias[ep.bias_mod > 0 ? (col + e) % ep.bias_mod : col + e];
        }
        if (ep.out_scale != 0.f) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] *= ep.out_scale;
        }
        if (ep.resid != nullptr) {
          const float4 x = *reinterpret_cast<const float4*>(ep.resid + (size_t)row * ep.ldr + col);
          if (ep.gate != nullptr) {
            const float4 g = *reinterpret_cast<const float4*>(ep.gate + (size_t)b * ep.gate_bstride + col);
            v[0] = x.x + (1.0f - g.x) * v[0];
            v[1] = x.y + (1.0f - g.y) * v[1];
            v[2] = x.z + (1.0f - g.z) * v[2];
            v[3] = x.w + (1.0f - g.w) * v[3];
          } else {
            v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w;
          }
        }
        if (ep.out_f32 != nullptr)
          *reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * ep.ld32 + col) = make_float4(v[0], v[1], v[2], v[3]);
        if (ep.out_bf16 != nullptr) {
          if (ep.act == ACT_SILU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = silu(v[e]);
          } else if (ep.act == ACT_SNAKE) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int ch = ep.bias_mod > 0 ? (col + e) % ep.bias_mod : col + e;
              const float s = sinf(v[e] * ep.act_a[ch]);
              v[e] = v[e] + ep.act_b[ch] * s * s;
            }
          }
          const int c16 = ep.phase_cols > 0 ? (col / ep.phase_cols) * ep.phase_ld16 + col % ep.phase_cols : col;
          store_bf16x4(ep.out_bf16 + (size_t)row * ep.ld16 + c16, ep.split_stride, v[0], v[1], v[2], v[3]);
        }
      }
    }
  }
};

// GEGLU (src/models/utils/modules.py:274-277): W rows are packed so that an N-tile of BN columns holds BN/2 hidden
// features followed by the BN/2 matching gate features; out[row, n0/2 + j] = (h_j + bh_j) * gelu_erf(g_j + bg_j).
struct EpiGegluParams {
  const float* bias;  // packed like the weight rows
  __nv_bfloat16* out_bf16;
  int ld16;
  int split_stride;
};
template <int BN>
struct EpiGeglu {
  using Params = EpiGegluParams;
  static __device__ __forceinline__ void run(const Params& ep, uint32_t taddr_row, int row, int M, int n0, int N) {
    constexpr int HALF = BN / 2;
    const bool row_ok = row < M;
#pragma unroll 1
    for (int c = 0; c < HALF; c += 16) {
      if (n0 + c >= N) break;
      uint32_t h[16], g[16];
      __syncwarp();
      tmem_ld_32x16(taddr_row + c, h);
      tmem_ld_32x16(taddr_row + HALF + c, g);
      tmem_ld_wait();
      __nv_bfloat16* o = ep.out_bf16 + (size_t)row * ep.ld16 + (n0 / 2 + c);
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        if (!row_ok) break;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float hv = __uint_as_float(h[j + e]) + ep.bias[n0 + c + j + e];
          const float gv = __uint_as_float(g[j + e]) + ep.bias[n0 + HALF + c + j + e];
          v[e] = hv * gelu_erf(gv);
        }
        store_bf16x4(o + j, ep.split_stride, v[0], v[1], v[2], v[3]);
      }
    }
  }
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int BYTES = 1024 /*align slack*/ + STAGES * (A_BYTES + B_BYTES) + (2 * STAGES + 4) * 8 + 16;
};

template <int BN, int STAGES, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape g,
                    const typename Epi::Params ep) {
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  using SM = GemmSmem<BN, STAGES>;
  constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * SM::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * SM::B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = g.num_m_tiles * g.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % g.num_m_tiles, nt = tile / g.num_m_tiles;
        const int n0 = nt * BN;
        for (int kb = 0; kb < g.num_k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], SM::A_BYTES + SM::B_BYTES);
          if (g.taps == 0) {
            tma_load_2d(sA + stage * SM::A_BYTES, &tmA, &full[stage], kb * GEMM_BK, mt * GEMM_BM);
          } else {
            const int tap = kb / g.cin_blocks, cb = kb - tap * g.cin_blocks;
            const int bidx = mt / g.tiles_per_batch, t0 = (mt - bidx * g.tiles_per_batch) * GEMM_BM;
            tma_load_3d(sA + stage * SM::A_BYTES, &tmA, &full[stage], cb * GEMM_BK, t0 + (tap - g.center) * g.dilation, bidx);
          }
          tma_load_2d(sB + stage * SM::B_BYTES, &tmB, &full[stage], kb * GEMM_BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < g.num_k_blocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * SM::A_BYTES));
          const uint64_t bd = umma_desc_sw128(smem_u32(sB + stage * SM::B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
          if (kb == g.num_k_blocks - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------------------------ epilogue (warps 2..5; TMEM lane group = warp % 4)
    const int lg = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile % g.num_m_tiles, nt = tile / g.num_m_tiles;
      int row;
      int m_limit = g.M;
      if (g.taps == 0) {
        row = mt * GEMM_BM + lg * 32 + lane;
      } else {  // rows are (batch, t): tiles never straddle clips
        const int bidx = mt / g.tiles_per_batch, t = (mt - bidx * g.tiles_per_batch) * GEMM_BM + lg * 32 + lane;
        row = (t < g.T) ? bidx * g.T + t : g.M;
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + acc * BN + (static_cast<uint32_t>(lg * 32) << 16);
      Epi::run(ep, taddr_row, row, m_limit, nt * BN, g.N);
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace ezb
