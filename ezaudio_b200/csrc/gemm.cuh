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
#ifdef EZB_GEMM_DEBUG
#define EZB_DBG(...) __VA_ARGS__
#else
#define EZB_DBG(...)
#endif
constexpr int GEMM_SMEM_BUDGET = 227 * 1024 - 2048;  // dynamic smem per CTA minus alignment slack and barriers

struct GemmShape {
  int M, N;
  int num_k_blocks;    // K / 64 (rounded up; TMA zero-fills the tail)
  int num_m_tiles, num_n_tiles;
  // implicit-conv addressing of A (taps == 0 -> plain 2-D A[M,K])
  int taps, center, dilation, cin_blocks, T, tiles_per_batch;
  int stride, pad;     // stride > 1: strided conv (VAE encoder): A is a 4-D map [B, T/stride, stride, C], tap k reads row q*stride + k - pad
  unsigned long long* dbg;  // optional cycle counters (CTA 0): [0] mma wait full, [1] mma wait tempty, [2] producer wait empty,
                            // [3] epilogue warp 2 wait tfull, [4] epilogue warp 2 busy, [5] total
  // L2 prefetch of the weights the NEXT GEMM of the step will stream (host.cuh WeightSeq): every layer's weights are read once per step,
  // i.e. from HBM, and a kernel's first k-blocks pay that latency on top of its ramp (GEGLU: 58 us with L2-resident weights, 68.6 us in situ).
  const char* pf;
  unsigned int pf_bytes;    // multiple of 16
};

// CTA c of `nctas` asks the L2 for its share (16 KB pieces, round-robin) of [pf, pf + bytes): a hint, no completion to wait for
__device__ __forceinline__ void prefetch_weights_l2(const char* pf, unsigned int bytes, unsigned int cta, unsigned int nctas) {
  constexpr unsigned int CH = 16384;
  for (unsigned int off = cta * CH; off < bytes; off += nctas * CH) {
    const unsigned int n = bytes - off < CH ? bytes - off : CH;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(pf + off)), "r"(n) : "memory");
  }
}

enum { ACT_NONE = 0, ACT_SILU = 1, ACT_SNAKE = 2 };

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm folded into the GEMMs on either side of it (fast mode; blocks.py:137,149,155 and modules.py:15-16):
//     h = ((x - mu) * rstd) * g + c,   g = w (1 + scale), c = b (1 + scale) + shift        (LayerNorm affine + AdaLN modulate)
//     h W^T = rstd * ( (x g) W^T  -  mu * u ) + v,      u[n] = sum_k g[k] W[n,k],  v[n] = sum_k c[k] W[n,k]
// The GEMM that WRITES the residual stream x (FoldOut) also writes A = bf16(x * g) -- the operand of the GEMM that follows the
// LayerNorm -- and per-row partial sums (sum x, sum x^2), one slot per 32-feature lane group: slot-major [slots][ld_st], fixed
// summation order, so the result is deterministic.  The GEMM that FOLLOWS the LayerNorm (FoldIn) turns the partials into (mu, rstd)
// per row and applies the per-row affine to its accumulator.  No LayerNorm pass, no extra launch; u, v come from per-timestep tables.
struct FoldIn {
  const float2* st0;   // partials of the row (or of the first half of a concatenated row); null: no fold
  const float2* st1;   // second source (skip path: LayerNorm over [x | skip]) or null
  int slots0, slots1, ld_st;
  float inv_dim;       // 1 / (normalised width)
  const float* u;      // [N] in this GEMM's (packed) output-column order
  const float* v;
};
struct FoldOut {
  float2* st;          // null: nothing to emit
  int ld_st;
  __nv_bfloat16* a0; int ld0; const float* g0;   // a0[token, f] = bf16(x * g0[f])  (g0 null: plain cast)
  __nv_bfloat16* a1; int ld1; const float* g1;   // optional second consumer of the same x (null: none)
};
__device__ __forceinline__ void fold_row_stats(const FoldIn& f, int row, float& rstd, float& nmr) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
  for (int s = 0; s < f.slots0; ++s) { const float2 p = f.st0[(size_t)s * f.ld_st + row]; s1 += p.x; s2 += p.y; }
#pragma unroll 4
  for (int s = 0; s < f.slots1; ++s) { const float2 p = f.st1[(size_t)s * f.ld_st + row]; s1 += p.x; s2 += p.y; }
  const float mean = s1 * f.inv_dim;
  const float var = fmaxf(s2 * f.inv_dim - mean * mean, 0.f);
  rstd = rsqrtf(var + 1e-5f);
  nmr = -mean * rstd;
}
// x[j] (j = 0..31) per lane -> lane L ends with sum over the 32 lanes of x[L]: 31 shuffles instead of 32 x 5
__device__ __forceinline__ float warp_transpose_sum(float (&x)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? x[i] : x[i + o];
      const float keep = up ? x[i + o] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return x[0];
}

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
  FoldIn fin;              // swap-AB epilogue only: LayerNorm folded in / out (see above)
  FoldOut fout;
};

// ---------------------------------------------------------------------------------------------------------------
// Epilogue staging: a warp reads its 32 accumulator rows from TMEM (thread == row), transposes 64-column chunks through
// a private 8 KB smem tile (XOR-swizzled 8-byte granules: conflict-free both ways) and then works with lane == column
// pair, so every global access is a fully coalesced 128/256-byte row segment.
constexpr int EPI_STAGE_FLOATS = 32 * 64;

template <int PITCH = 64>
__device__ __forceinline__ void stage_put(float* st, int r, int g, float a, float b) {
  *reinterpret_cast<float2*>(st + r * PITCH + 2 * (g ^ (r & (PITCH / 2 - 1)))) = make_float2(a, b);
}
template <int PITCH = 64>
__device__ __forceinline__ float2 stage_get(const float* st, int rr, int lane) {
  return *reinterpret_cast<const float2*>(st + rr * PITCH + 2 * (lane ^ (rr & (PITCH / 2 - 1))));
}
__device__ __forceinline__ void store_bf16x2(__nv_bfloat16* p, int split_stride, float a, float b) {
  const uint32_t hi = pack_bf16(a, b);
  *reinterpret_cast<uint32_t*>(p) = hi;
  if (split_stride > 0) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&hi);
    *reinterpret_cast<uint32_t*>(p + split_stride) = pack_bf16(a - __low2float(h), b - __high2float(h));
    *reinterpret_cast<uint32_t*>(p + 2 * split_stride) = hi;
  }
}

struct RowCtx {
  const float* st; int lane, nv, row0, brow;
  float b0, b1; float2 g0, g1; float sa0, sa1, sb0, sb1;
  float* o32; int ld32; __nv_bfloat16* o16; int ld16;
};
template <bool RESID, bool GATE, bool F32, bool BF16, int ACT>
__device__ __forceinline__ void rows_t(const RowCtx& c, const float2* x) {
  float* o32 = c.o32;
  __nv_bfloat16* o16 = c.o16;
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) {
    if (rr >= c.nv) break;
    const float2 acc = stage_get(c.st, rr, c.lane);
    float v0 = acc.x + c.b0, v1 = acc.y + c.b1;
    if (RESID) {
      if (GATE) {
        const bool second = c.row0 + rr >= c.brow;
        v0 = fmaf(second ? c.g1.x : c.g0.x, v0, x[rr].x);
        v1 = fmaf(second ? c.g1.y : c.g0.y, v1, x[rr].y);
      } else {
        v0 += x[rr].x;
        v1 += x[rr].y;
      }
    }
    if (F32) { *reinterpret_cast<float2*>(o32) = make_float2(v0, v1); o32 += c.ld32; }
    if (BF16) {
      if (ACT == ACT_SILU) { v0 = silu(v0); v1 = silu(v1); }
      if (ACT == ACT_SNAKE) {  // bf16 throughput path: MUFU sine (the result is rounded to bf16); bf16x3 uses the exact sinf (rows_generic)
        const float s0 = __sinf(v0 * c.sa0), s1 = __sinf(v1 * c.sa1);
        v0 = fmaf(c.sb0 * s0, s0, v0);
        v1 = fmaf(c.sb1 * s1, s1, v1);
      }
      *reinterpret_cast<uint32_t*>(o16) = pack_bf16(v0, v1);
      o16 += c.ld16;
    }
  }
}
__device__ __forceinline__ void rows_generic(const EpiLinearParams& ep, const RowCtx& c, const float2* x, int col, int c16) {
  const float osc = ep.out_scale != 0.f ? ep.out_scale : 1.f;
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) {
    if (rr >= c.nv) break;
    const int row = c.row0 + rr;
    const float2 acc = stage_get(c.st, rr, c.lane);
    float v0 = (acc.x + c.b0) * osc, v1 = (acc.y + c.b1) * osc;
    if (ep.resid != nullptr) {
      if (ep.gate != nullptr) {
        float2 g = row >= c.brow ? c.g1 : c.g0;
        if (ep.rows_per_batch < 32) {  // short clips (L < 32, api/ezaudio.py:160-172 crops to any length): a warp's rows span > 2 batch items
          g = *reinterpret_cast<const float2*>(ep.gate + (size_t)(row / ep.rows_per_batch) * ep.gate_bstride + col);
          g.x = 1.0f - g.x; g.y = 1.0f - g.y;
        }
        v0 = x[rr].x + g.x * v0;
        v1 = x[rr].y + g.y * v1;
      } else {
        v0 += x[rr].x;
        v1 += x[rr].y;
      }
    }
    if (ep.out_f32 != nullptr) *reinterpret_cast<float2*>(ep.out_f32 + (size_t)row * ep.ld32 + col) = make_float2(v0, v1);
    if (ep.out_bf16 != nullptr) {
      if (ep.act == ACT_SILU) {
        v0 = silu(v0);
        v1 = silu(v1);
      } else if (ep.act == ACT_SNAKE) {
        const float s0 = sinf(v0 * c.sa0), s1 = sinf(v1 * c.sa1);
        v0 = v0 + c.sb0 * s0 * s0;
        v1 = v1 + c.sb1 * s1 * s1;
      }
      store_bf16x2(ep.out_bf16 + (size_t)row * ep.ld16 + c16, ep.split_stride, v0, v1);
    }
  }
}

template <int BN>
struct EpiLinear {
  using Params = EpiLinearParams;
  static constexpr int EPI_WARPS = BN >= 128 ? 8 : 4;       // two warps per TMEM lane group split the tile's columns
  static constexpr int STAGE_FLOATS = EPI_STAGE_FLOATS;
  // row0: global row of this warp's first accumulator row; nvalid: rows of the 32 that exist (<= 0: none);
  // [c_begin, c_end): this warp's column range inside the tile; wait(): blocks until the accumulator is complete.
  template <class Wait>
  static __device__ __forceinline__ void run(const Params& ep, float* st, uint32_t taddr_row, int row0, int nvalid, int n0, int N, int lane, int c_begin,
                                             int c_end, Wait wait) {
    const int nv = nvalid < 32 ? nvalid : 32;
    bool waited = false;
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 64) {
      const int col = n0 + c + 2 * lane;
      const bool col_ok = (c + 2 * lane < c_end) && col < N;
      // operands that do not depend on the accumulator are fetched first (and, for the first chunk, before the wait)
      float2 x[32];
      if (ep.resid != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          x[i] = make_float2(0.f, 0.f);
          if (col_ok && i < nv) x[i] = *reinterpret_cast<const float2*>(ep.resid + (size_t)(row0 + i) * ep.ldr + col);
        }
      }
      float b0 = 0.f, b1 = 0.f, sa0 = 0.f, sa1 = 0.f, sb0 = 0.f, sb1 = 0.f;
      float2 g0 = make_float2(0.f, 0.f), g1 = g0;
      int brow = 0x7fffffff;  // first row that belongs to the second batch item touched by this warp
      if (col_ok) {
        const int ch = ep.bias_mod > 0 ? col % ep.bias_mod : col;
        if (ep.bias != nullptr) { b0 = ep.bias[ch]; b1 = ep.bias[ch + 1]; }
        if (ep.act == ACT_SNAKE) { sa0 = ep.act_a[ch]; sa1 = ep.act_a[ch + 1]; sb0 = ep.act_b[ch]; sb1 = ep.act_b[ch + 1]; }
        if (ep.gate != nullptr && nv > 0) {
          const int b0i = row0 / ep.rows_per_batch;
          brow = (b0i + 1) * ep.rows_per_batch;
          g0 = *reinterpret_cast<const float2*>(ep.gate + (size_t)b0i * ep.gate_bstride + col);
          if (brow < row0 + nv) g1 = *reinterpret_cast<const float2*>(ep.gate + (size_t)(b0i + 1) * ep.gate_bstride + col);
          g0.x = 1.0f - g0.x; g0.y = 1.0f - g0.y; g1.x = 1.0f - g1.x; g1.y = 1.0f - g1.y;
        }
      }
      if (!waited) { wait(); waited = true; }
      __syncwarp();
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {  // two 32-column halves: 32 accumulator registers live next to the 64 prefetched residual values (one 64-wide
        uint32_t r[32];                 // load spilled 780 bytes per thread under the 168-register cap of this 320-thread CTA)
        tmem_ld_32x32(taddr_row + c + 32 * hc, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 16; ++g) stage_put(st, lane, 16 * hc + g, __uint_as_float(r[2 * g]), __uint_as_float(r[2 * g + 1]));
      }
      __syncwarp();
      if (col_ok) {
        const int c16 = ep.phase_cols > 0 ? (col / ep.phase_cols) * ep.phase_ld16 + col % ep.phase_cols : col;
        const int code = (ep.resid != nullptr ? 1 : 0) | (ep.gate != nullptr ? 2 : 0) | (ep.out_f32 != nullptr ? 4 : 0) | (ep.out_bf16 != nullptr ? 8 : 0) |
                         (ep.act << 4) | ((ep.split_stride > 0 || ep.out_scale != 0.f || (ep.gate != nullptr && ep.rows_per_batch < 32)) ? 256 : 0);
        const RowCtx rc{st, lane, nv, row0, brow, b0, b1, g0, g1, sa0, sa1, sb0, sb1,
                        ep.out_f32 != nullptr ? ep.out_f32 + (size_t)row0 * ep.ld32 + col : nullptr, ep.ld32,
                        ep.out_bf16 != nullptr ? ep.out_bf16 + (size_t)row0 * ep.ld16 + c16 : nullptr, ep.ld16};
        switch (code) {
          case 4: rows_t<false, false, true, false, ACT_NONE>(rc, x); break;             // bias -> f32
          case 7: rows_t<true, true, true, false, ACT_NONE>(rc, x); break;               // gated residual (in place)
          case 5: rows_t<true, false, true, false, ACT_NONE>(rc, x); break;              // residual
          case 8: rows_t<false, false, false, true, ACT_NONE>(rc, x); break;             // bf16
          case 8 | (ACT_SILU << 4): rows_t<false, false, false, true, ACT_SILU>(rc, x); break;
          case 8 | (ACT_SNAKE << 4): rows_t<false, false, false, true, ACT_SNAKE>(rc, x); break;       // conv7
          case 12 | (ACT_SNAKE << 4): rows_t<false, false, true, true, ACT_SNAKE>(rc, x); break;       // conv-transpose
          case 13 | (ACT_SNAKE << 4): rows_t<true, false, true, true, ACT_SNAKE>(rc, x); break;        // conv1 + residual
          case 9 | (ACT_SNAKE << 4): rows_t<true, false, false, true, ACT_SNAKE>(rc, x); break;        // last conv1 of a block
          default: rows_generic(ep, rc, x, col, c16); break;
        }
      }
    }
    if (!waited) wait();
  }
};

// GEGLU (src/models/utils/modules.py:274-277): W rows are packed so that an N-tile of BN columns holds BN/2 hidden
// features followed by the BN/2 matching gate features; out[row, n0/2 + j] = (h_j + bh_j) * gelu_erf(g_j + bg_j).
struct EpiGegluParams {
  const float* bias;  // packed like the weight rows
  __nv_bfloat16* out_bf16;
  int ld16;
  int split_stride;
  FoldIn fin;         // LayerNorm folded into this GEMM (fin.v already contains the bias)
};
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// h * gelu(g) with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7 + 2 ulp of the MUFU rcp / ex2): 2 MUFU + ~14 FP32 ops.
// Used by the bf16 throughput path only (its output is rounded to bf16, 4e-3 relative); bf16x3 calls the exact erff.
__device__ __forceinline__ float geglu_fast(float h, float g) {
  const float z = g * 0.70710678118654752440f, az = fabsf(z);
  const float t = rcp_approx(fmaf(0.3275911f, az, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = ex2_approx(az * az * -1.4426950408889634f);
  const float erf_abs = fmaf(-p * t, e, 1.0f);
  const float erf_s = copysignf(erf_abs, z);
  return h * (0.5f * g) * (1.0f + erf_s);
}

template <int BN, bool FOLD = false>
struct EpiGeglu {
  using Params = EpiGegluParams;
  static constexpr int HALF = BN / 2;
  static constexpr int EPI_WARPS = BN == 256 ? 8 : 4;   // 8 warps: each takes 64 of the tile's 128 output features
  static constexpr int STAGE_FLOATS = 32 * 32;          // 64 packed bf16 per row
  template <class Wait>
  static __device__ __forceinline__ void run(const Params& ep, float* st, uint32_t taddr_row, int row0, int nvalid, int n0, int N, int lane, int c_begin,
                                             int c_end, Wait wait) {
    constexpr bool fold = FOLD;
    float rstd = 1.f, nmr = 0.f;
    if (fold && lane < nvalid) fold_row_stats(ep.fin, row0 + lane, rstd, nmr);   // before the accumulator wait
    wait();
    // c_begin/c_end are expressed in accumulator columns of the whole tile: map them to output-feature ranges
    const int f_begin = c_begin / 2, f_end = c_end / 2;
#pragma unroll 1
    for (int c = f_begin; c < f_end; c += 64) {  // 64 output features per pass, gated in the thread == row layout
      if (n0 + c >= N) break;
      if (ep.split_stride > 0) {  // bf16x3 parity mode: exact erf, per-thread row stores of [hi | lo | hi]
        uint32_t h[32], g[32];
#pragma unroll 1
        for (int q4 = 0; q4 < 2; ++q4) {
          __syncwarp();
          tmem_ld_32x32(taddr_row + c + q4 * 32, h);
          tmem_ld_32x32(taddr_row + HALF + c + q4 * 32, g);
          tmem_ld_wait();
          if (lane < nvalid) {
            __nv_bfloat16* o = ep.out_bf16 + (size_t)(row0 + lane) * ep.ld16 + (n0 / 2 + c + q4 * 32);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float o0 = (__uint_as_float(h[j]) + ep.bias[n0 + c + q4 * 32 + j]) * gelu_erf(__uint_as_float(g[j]) + ep.bias[n0 + HALF + c + q4 * 32 + j]);
              const float o1 = (__uint_as_float(h[j + 1]) + ep.bias[n0 + c + q4 * 32 + j + 1]) * gelu_erf(__uint_as_float(g[j + 1]) + ep.bias[n0 + HALF + c + q4 * 32 + j + 1]);
              store_bf16x2(o + j, ep.split_stride, o0, o1);
            }
          }
        }
        continue;
      }
      uint32_t pk[32];  // 64 bf16 results of this thread's row
#pragma unroll
      for (int q4 = 0; q4 < 2; ++q4) {
        uint32_t h[32], g[32];
        __syncwarp();
        tmem_ld_32x32(taddr_row + c + q4 * 32, h);
        tmem_ld_32x32(taddr_row + HALF + c + q4 * 32, g);
        tmem_ld_wait();
        if (fold) {   // warp-uniform: h = rstd * acc - rstd * mu * u + (v + bias)
          const float* bh = ep.fin.v + n0 + c + q4 * 32;
          const float* bg = ep.fin.v + n0 + HALF + c + q4 * 32;
          const float* uh = ep.fin.u + n0 + c + q4 * 32;
          const float* ug = ep.fin.u + n0 + HALF + c + q4 * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float o0 = geglu_fast(fmaf(__uint_as_float(h[j]), rstd, fmaf(nmr, __ldg(uh + j), __ldg(bh + j))),
                                        fmaf(__uint_as_float(g[j]), rstd, fmaf(nmr, __ldg(ug + j), __ldg(bg + j))));
            const float o1 = geglu_fast(fmaf(__uint_as_float(h[j + 1]), rstd, fmaf(nmr, __ldg(uh + j + 1), __ldg(bh + j + 1))),
                                        fmaf(__uint_as_float(g[j + 1]), rstd, fmaf(nmr, __ldg(ug + j + 1), __ldg(bg + j + 1))));
            pk[q4 * 16 + j / 2] = pack_bf16(o0, o1);
          }
        } else {
        const float* bh = ep.bias + n0 + c + q4 * 32;
        const float* bg = ep.bias + n0 + HALF + c + q4 * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float o0 = geglu_fast(__uint_as_float(h[j]) + __ldg(bh + j), __uint_as_float(g[j]) + __ldg(bg + j));
          const float o1 = geglu_fast(__uint_as_float(h[j + 1]) + __ldg(bh + j + 1), __uint_as_float(g[j + 1]) + __ldg(bg + j + 1));
          pk[q4 * 16 + j / 2] = pack_bf16(o0, o1);
        }
        }
      }
      __syncwarp();
#pragma unroll
      for (int gq = 0; gq < 16; ++gq) stage_put<32>(st, lane, gq, __uint_as_float(pk[2 * gq]), __uint_as_float(pk[2 * gq + 1]));
      __syncwarp();
      // 16 lanes cover one 128-byte row segment (64 bf16): the two half-warps take alternate rows
      const int hl = lane & 15, hw = lane >> 4;
      __nv_bfloat16* obase = ep.out_bf16 + (size_t)row0 * ep.ld16 + (n0 / 2 + c + 4 * hl);
#pragma unroll 4
      for (int rp = 0; rp < 16; ++rp) {
        const int rr = 2 * rp + hw;
        if (rr < nvalid) *reinterpret_cast<float2*>(obase + (size_t)rr * ep.ld16) = stage_get<32>(st, rr, hl);
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Transposed ("swap-AB") linear epilogue.  For N_out = 1152-wide layers the natural 128/144-column tiles leave the tensor
// pipe ~55 % efficient (operand bytes per flop) and a 256-column tile does not divide 1152.  Computing C^T = W A^T instead
// puts the 1152 output features on the accumulator ROWS (9 tiles of 128) and 256 tokens on the columns: 144 tiles of
// 128 x 256 = one full wave on 148 SMs.  A thread now owns one output feature; for a given token the 32 lanes of a warp hold
// 32 consecutive features, so residual loads and stores are 128-byte coalesced without any staging.
template <int BN>
struct EpiLinearT {
  using Params = EpiLinearParams;   // bias/gate indexed by feature, resid/out_f32 [token, feature]; bf16/act/split unsupported
  static constexpr int EPI_WARPS = 8;
  static constexpr int STAGE_FLOATS = 0;
  // residual values of one 32-token chunk (feature f of tokens t0 .. t0+31)
  static __device__ __forceinline__ void load_resid(const Params& ep, float (&x)[32], int t0, int N, int f, bool f_ok) {
    const int nt = (N - t0) < 32 ? (N - t0) : 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = (f_ok && j < nt) ? ep.resid[(size_t)(t0 + j) * ep.ldr + f] : 0.f;
  }
  template <class Wait>
  static __device__ __forceinline__ void run(const Params& ep, float* st, uint32_t taddr_row, int row0, int nvalid, int n0, int N, int lane, int c_begin,
                                             int c_end, Wait wait) {
    const int f = row0 + lane;                 // output feature of this thread
    const bool f_ok = lane < nvalid;
    const float bias = (ep.bias != nullptr && f_ok) ? ep.bias[f] : 0.f;
    bool waited = false;
    // The kernel is one wave, so this epilogue is exposed: the residual of chunk c + 1 is fetched while chunk c is processed (and the first
    // chunk's before the accumulator wait), instead of one L2 round trip per chunk on the critical path.
    float x[32];
    if (ep.resid != nullptr && n0 + c_begin < N) load_resid(ep, x, n0 + c_begin, N, f, f_ok);
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      const int t0 = n0 + c;                   // first token of this chunk
      if (t0 >= N) break;                      // warp-uniform
      const int nt = (N - t0) < 32 ? (N - t0) : 32;
      float xn[32];
      const bool more = ep.resid != nullptr && c + 32 < c_end && t0 + 32 < N;
      if (more) load_resid(ep, xn, t0 + 32, N, f, f_ok);
      float g0 = 1.f, g1 = 1.f;
      int btok = 0x7fffffff;                   // first token that belongs to the second batch item of this chunk
      if (ep.gate != nullptr && f_ok) {
        const int b0i = t0 / ep.rows_per_batch;
        btok = (b0i + 1) * ep.rows_per_batch;
        g0 = 1.0f - ep.gate[(size_t)b0i * ep.gate_bstride + f];
        if (btok < t0 + nt) g1 = 1.0f - ep.gate[(size_t)(b0i + 1) * ep.gate_bstride + f];
      }
      if (!waited) { wait(); waited = true; }
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(taddr_row + c, r);
      tmem_ld_wait();
      if (f_ok) {
        float* o = ep.out_f32 + (size_t)t0 * ep.ld32 + f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j >= nt) break;
          float v = __uint_as_float(r[j]) + bias;
          if (ep.resid != nullptr) v = fmaf((t0 + j >= btok) ? g1 : g0, v, x[j]);
          o[(size_t)j * ep.ld32] = v;
        }
      }
      if (more) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = xn[j];
      }
    }
    if (!waited) wait();
  }
};

// Fold-capable variant (LayerNorm folded in / out, see FoldIn / FoldOut).  Kept apart from EpiLinearT on purpose: the extra outputs and the
// warp transposes cost this one-wave kernel ~19 us per launch (measured, profiles/r2), more than the LayerNorm pass they replace.
template <int BN>
struct EpiLinearTF {
  using Params = EpiLinearParams;   // bias/gate indexed by feature, resid/out_f32 [token, feature]; bf16/act/split unsupported
  static constexpr int EPI_WARPS = 8;
  static constexpr int STAGE_FLOATS = 0;
  template <class Wait>
  static __device__ __forceinline__ void run(const Params& ep, float* st, uint32_t taddr_row, int row0, int nvalid, int n0, int N, int lane, int c_begin,
                                             int c_end, Wait wait) {
    const int f = row0 + lane;                 // output feature of this thread
    const bool f_ok = lane < nvalid;
    const float bias = (ep.bias != nullptr && f_ok) ? ep.bias[f] : 0.f;
    const bool fold_in = ep.fin.u != nullptr, fold_out = ep.fout.st != nullptr && nvalid > 0;
    float uf = 0.f, vf = 0.f, ga = 1.f, gb = 1.f;
    if (fold_in && f_ok) { uf = ep.fin.u[f]; vf = ep.fin.v[f]; }
    if (fold_out && f_ok) {
      if (ep.fout.g0 != nullptr) ga = ep.fout.g0[f];
      if (ep.fout.a1 != nullptr && ep.fout.g1 != nullptr) gb = ep.fout.g1[f];
    }
    bool waited = false;
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      const int t0 = n0 + c;                   // first token of this chunk
      if (t0 >= N) break;                      // warp-uniform
      const int nt = (N - t0) < 32 ? (N - t0) : 32;
      float x[32];
      if (ep.resid != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = (f_ok && j < nt) ? ep.resid[(size_t)(t0 + j) * ep.ldr + f] : 0.f;
      }
      float g0 = 1.f, g1 = 1.f;
      int btok = 0x7fffffff;                   // first token that belongs to the second batch item of this chunk
      if (ep.gate != nullptr && f_ok) {
        const int b0i = t0 / ep.rows_per_batch;
        btok = (b0i + 1) * ep.rows_per_batch;
        g0 = 1.0f - ep.gate[(size_t)b0i * ep.gate_bstride + f];
        if (btok < t0 + nt) g1 = 1.0f - ep.gate[(size_t)(b0i + 1) * ep.gate_bstride + f];
      }
      float rs_l = 1.f, nm_l = 0.f;            // LayerNorm statistics of token t0 + lane (fold-in)
      if (fold_in && lane < nt) fold_row_stats(ep.fin, t0 + lane, rs_l, nm_l);
      if (!waited) { wait(); waited = true; }
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(taddr_row + c, r);
      tmem_ld_wait();
      float val[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float acc = __uint_as_float(r[j]);
        if (fold_in) {                           // warp-uniform branch; the shuffles run on all lanes
          const float rs = __shfl_sync(0xffffffffu, rs_l, j), nm = __shfl_sync(0xffffffffu, nm_l, j);
          acc = fmaf(acc, rs, fmaf(nm, uf, vf));
        }
        float v = acc + bias;
        float gj = (t0 + j >= btok) ? g1 : g0;
        if (ep.resid != nullptr) v = fmaf(gj, v, x[j]);
        val[j] = (f_ok && j < nt) ? v : 0.f;
      }
      if (f_ok) {
        float* o = ep.out_f32 + (size_t)t0 * ep.ld32 + f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j >= nt) break;
          o[(size_t)j * ep.ld32] = val[j];
        }
        if (fold_out) {                          // operand(s) of the GEMM(s) behind the LayerNorm(s) that read this x
          __nv_bfloat16* a = ep.fout.a0 + (size_t)t0 * ep.fout.ld0 + f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j >= nt) break;
            a[(size_t)j * ep.fout.ld0] = __float2bfloat16_rn(val[j] * ga);
          }
          if (ep.fout.a1 != nullptr) {
            __nv_bfloat16* a2 = ep.fout.a1 + (size_t)t0 * ep.fout.ld1 + f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j >= nt) break;
              a2[(size_t)j * ep.fout.ld1] = __float2bfloat16_rn(val[j] * gb);
            }
          }
        }
      }
      if (fold_out) {                            // per-token partial sums over this warp's 32 features -> slot row0 / 32
        float sq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) sq[j] = val[j] * val[j];
        const float s1 = warp_transpose_sum(val, lane), s2 = warp_transpose_sum(sq, lane);
        if (lane < nt) ep.fout.st[(size_t)(row0 >> 5) * ep.fout.ld_st + t0 + lane] = make_float2(s1, s2);
      }
    }
    if (!waited) wait();
  }
};

// KSUB: 64-wide K sub-tiles per pipeline stage.  The single MMA thread pays ~250 cycles of fixed cost per stage (mbarrier
// wait, fence, two commits); with N <= 144 a 64-deep stage is only 4 x 64..72 cycles of tensor work, so narrow tiles use
// 128-deep stages (KSUB = 2) to keep the issue loop off the critical path.
template <int BN, class Epi, bool PAIR, int KSUB = 1>
struct GemmCfg {
  static constexpr int A_SUB = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_SUB = (PAIR ? BN / 2 : BN) * GEMM_BK * 2;
  static constexpr int A_BYTES = KSUB * A_SUB;
  static constexpr int B_BYTES = KSUB * B_SUB;
  static constexpr int EPI_WARPS = Epi::EPI_WARPS;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int STAGE_BYTES = EPI_WARPS * Epi::STAGE_FLOATS * 4;   // one transpose tile per epilogue warp
  static constexpr int FIT = (GEMM_SMEM_BUDGET - STAGE_BYTES) / (A_BYTES + B_BYTES);
  static constexpr int STAGES = FIT > 8 ? 8 : FIT;
  static constexpr int BYTES = 1024 /*align slack*/ + STAGES * (A_BYTES + B_BYTES) + STAGE_BYTES + (2 * STAGES + 4) * 8 + 16;
  static_assert(STAGES >= 3, "smem budget");
};

// MC > 1: launched as clusters of MC CTAs with consecutive blockIdx.x = consecutive M tiles of the SAME N tile (host guarantees
// num_m_tiles % MC == 0, gridDim.x % MC == 0 and num_tiles % MC == 0, so the CTAs of a cluster walk the same number of tiles in step).
// The N-side operand tile (BN rows x 64 columns) is then fetched ONCE per cluster: tmB is a map with 32-row boxes, CTA rank r issues the
// sub-boxes j = r, r + MC, ... with .multicast::cluster, every CTA still expects the full A + B bytes on its own `full` barrier, and a
// stage is free again only when all MC consumers have released it (`empty` counts MC multicast commits).
template <int BN, class Epi, int MC = 1>
__device__ __forceinline__ void gemm_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& g, const typename Epi::Params& ep, uint8_t* smem_raw) {
  static_assert(BN % 16 == 0 && BN >= 64 && BN <= 256, "BN");
  static_assert(MC >= 1 && MC <= 8 && (MC == 1 || BN % 32 == 0), "MC");
  constexpr uint16_t MC_MASK = static_cast<uint16_t>((1u << MC) - 1u);
  using SM = GemmCfg<BN, Epi, false>;
  constexpr int STAGES = SM::STAGES;
  constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (a round trip through uintptr_t loses the address space and
  // turns every staging access into a generic LD.E / ST.E)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * SM::A_BYTES;
  float* sStage = reinterpret_cast<float*>(sB + STAGES * SM::B_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * SM::B_BYTES + SM::STAGE_BYTES);
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
      mbar_init(&empty[i], MC);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 32 * SM::EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (MC > 1) cluster_sync_all();  // peers multicast into our stages and arrive on our barriers: they must exist first
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch();
  if (warp == 0 && g.pf_bytes != 0) {   // constant data: no need to wait for the previous kernel
    if (elect_one()) prefetch_weights_l2(g.pf, g.pf_bytes, blockIdx.x, gridDim.x);
    __syncwarp();
  }
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ------------------------------------------------ TMA producer.  The loop is warp-uniform and the copies are issued under elect.sync:
    // inside an `if (lane == 0)` region ptxas wraps every UTMALDG / UTCHMMA / UTCBAR (uniform-datapath instructions) in an
    // ELECT ... BRA.U.ANY serialisation loop with R2UR broadcasts (~40-60 cycles per instruction); under elect.sync it knows a
    // single lane is active and issues them back to back.
    {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % g.num_m_tiles, nt = tile / g.num_m_tiles;
        const int n0 = nt * BN;
        for (int kb = 0; kb < g.num_k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
          mbar_expect_tx(&full[stage], SM::A_BYTES + SM::B_BYTES);
          if (g.taps == 0) {
            tma_load_2d(sA + stage * SM::A_BYTES, &tmA, &full[stage], kb * GEMM_BK, mt * GEMM_BM);
          } else {
            const int tap = kb / g.cin_blocks, cb = kb - tap * g.cin_blocks;
            const int bidx = mt / g.tiles_per_batch, t0 = (mt - bidx * g.tiles_per_batch) * GEMM_BM;
            if (g.stride > 1) {
              const int off = tap - g.pad;                                   // input row = q * stride + off
              const int r = ((off % g.stride) + g.stride) % g.stride, dq = (off - r) / g.stride;
              tma_load_4d(sA + stage * SM::A_BYTES, &tmA, &full[stage], cb * GEMM_BK, r, t0 + dq, bidx);
            } else {
              tma_load_3d(sA + stage * SM::A_BYTES, &tmA, &full[stage], cb * GEMM_BK, t0 + (tap - g.center) * g.dilation, bidx);
            }
          }
          if (MC == 1) {
            tma_load_2d(sB + stage * SM::B_BYTES, &tmB, &full[stage], kb * GEMM_BK, n0);
          } else {
            for (int j = (int)cluster_ctarank(); j < BN / 32; j += MC)
              tma_load_2d_mc(sB + stage * SM::B_BYTES + j * 4096, &tmB, &full[stage], kb * GEMM_BK, n0 + j * 32, MC_MASK);
          }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues: see the producer's note)
    constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < g.num_k_blocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * SM::A_BYTES));
          const uint64_t bd = umma_desc_sw128(smem_u32(sB + stage * SM::B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          if (MC == 1) umma_commit(&empty[stage]);
          else umma_commit_mc(&empty[stage], MC_MASK);
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
      int row0, nvalid;
      if (g.taps == 0) {
        row0 = mt * GEMM_BM + lg * 32;
        nvalid = g.M - row0;
      } else {  // rows are (batch, t): tiles never straddle clips
        const int bidx = mt / g.tiles_per_batch, t0 = (mt - bidx * g.tiles_per_batch) * GEMM_BM + lg * 32;
        row0 = bidx * g.T + t0;
        nvalid = g.T - t0;
      }
      const uint32_t taddr_row = tmem_base + acc * BN + (static_cast<uint32_t>(lg * 32) << 16);
      constexpr int CW = BN / (SM::EPI_WARPS / 4);  // columns per warp
      const int ch = (warp - 2) >> 2;
      uint64_t* tf = &tfull[acc];
      const uint32_t ph = acc_phase;
      Epi::run(ep, sStage + (warp - 2) * Epi::STAGE_FLOATS, taddr_row, row0, nvalid, nt * BN, g.N, lane, ch * CW, (ch + 1) * CW, [tf, ph]() {
        mbar_wait(tf, ph);
        tc_fence_after();
      });
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MC > 1) cluster_sync_all();  // nobody leaves while a peer may still multicast into this CTA or arrive on its barriers
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}
template <int BN, class Epi, int MC = 1>
__global__ void __launch_bounds__((GemmCfg<BN, Epi, false>::THREADS), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape g,
                    const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_dyn[];
  gemm_body<BN, Epi, MC>(tmA, tmB, g, ep, smem_dyn);
}

}  // namespace ezb

namespace ezb {
// ---------------------------------------------------------------------------------------------------------------
// Fused "heads" epilogue for the Q/K/V projections (attention.py:127-129,137-144; rotary.py:6-18,72-84): an N-tile holds two
// whole heads, a thread owns one token row, so the per-head LayerNorm(dh) and the rotate-half RoPE are in-thread.
//   kind 0/1 (q/k): LN affine -> RoPE (optional) -> bf16 rows [b*H + h, l, 0..dh) (pitch ld_qk), written 144/128 B-coalesced
//                   through the staging tile;
//   kind 2   (v)  : bf16 V^T [b*H + h, d, l] (pitch Lpad): for a fixed d a warp stores 32 consecutive tokens (64 B).
struct EpiHeadsParams {
  int D, H, L;                 // model width, heads, tokens per batch item
  int kind[3];                 // section (n / D) -> 0 q, 1 k, 2 v
  float nw[2][72];             // LayerNorm(dh) weight / bias for q, k -- BY VALUE: they live in the constant bank, so the
  float nb[2][72];             // normalisation FFMAs take them as operands instead of issuing 2 x dh loads per token
  const float2* rope;          // [L][dh/2] (cos, sin) table or null
  int rope_ld;
  int rope_mufu;               // 1: evaluate cos/sin with the MUFU (__sincosf) from inv_freq instead of reading the table
  float inv_freq[36];          // rotary.py:41-42 (checkpoint buffer), by value -> constant bank
  int rope_kinds;              // bit k set: apply RoPE to kind k
  __nv_bfloat16* out[3];       // per kind: q rows, k rows, v^T
  int ld_qk, dvp, Lpad;
  int dbg;                     // profiling instantiation only (option heads_dbg; results garbage): 1 no RoPE, 2 no per-head LayerNorm, 4 no V^T stores, 8 no q / k stores
  FoldIn fin;                  // LayerNorm (+ AdaLN modulate) of the block input folded into this projection
};

// HPT = 2: tile = two adjacent heads of the reference column order (N-tile 2*dh).  HPT = 3: the packed QKV layout -- the
// 3H heads of [q | k | v] are regrouped three per tile (N-tile 224 for dh = 72: 3 x 72 + 8 zero columns; 192 for dh = 64), which
// makes the tile wide enough for the tensor pipe (narrow tiles are operand-bandwidth bound) and keeps one head per warp.
// DIRECT: every thread stores its own q / k row (dh bf16 = 128 or 144 contiguous bytes) with 16-byte stores instead of transposing it through a
// per-warp 8 KB shared-memory tile.  The staging tiles of 12 epilogue warps take 96 KB, which leaves the 256 x 224 QKV tile only FOUR 30 KB
// pipeline stages (the GEGLU kernel runs six); without them it gets seven.
template <int DH, int HPT = 2, bool DIRECT = false, bool FOLD = false, bool DBG = false>
struct EpiHeads {
  using Params = EpiHeadsParams;
  static constexpr int BN = HPT == 3 ? (DH == 72 ? 224 : 3 * DH) : 2 * DH;
  static constexpr int EPI_WARPS = 4 * HPT;   // the HPT warps of a TMEM lane group take one head each
  static constexpr int STAGE_FLOATS = DIRECT ? 0 : EPI_STAGE_FLOATS;
  template <class Wait>
  static __device__ __forceinline__ void run(const Params& ep, float* st, uint32_t taddr_row, int row0, int nvalid, int n0, int N, int lane, int c_begin,
                                             int c_end, Wait wait) {
    const int row = row0 + lane;
    const bool row_ok = lane < nvalid;
    constexpr bool fold = FOLD;
    float f_rstd = 1.f, f_nmr = 0.f;
    if (fold && row_ok) fold_row_stats(ep.fin, row, f_rstd, f_nmr);   // issued before the accumulator wait
    wait();
    const int b = row_ok ? row / ep.L : 0, l = row_ok ? row - b * ep.L : 0;
    {
      const int hh = c_begin / (BN / HPT);    // this warp's head inside the tile
      int sec, head;
      if (HPT == 3) {
        const int g = (n0 / BN) * 3 + hh;     // global head index in [q heads | k heads | v heads]
        sec = g / ep.H;
        head = g - sec * ep.H;
      } else {
        const int n = n0 + hh * DH;
        if (n >= N) return;
        sec = n / ep.D;
        head = (n - sec * ep.D) / DH;
      }
      const int kind = ep.kind[sec];
      uint32_t r[DH];
      __syncwarp();
      tmem_ld_32x64(taddr_row + hh * DH, r);
      if constexpr (DH == 72) tmem_ld_32x8(taddr_row + hh * DH + 64, r + 64);
      tmem_ld_wait();
      float v[DH];
#pragma unroll
      for (int i = 0; i < DH; ++i) v[i] = __uint_as_float(r[i]);
      if (fold) {  // warp-uniform: per-row affine of the folded LayerNorm; u, v are the same for every row (uniform 16-byte loads)
        const float4* u4 = reinterpret_cast<const float4*>(ep.fin.u + n0 + hh * DH);
        const float4* v4 = reinterpret_cast<const float4*>(ep.fin.v + n0 + hh * DH);
#pragma unroll
        for (int i = 0; i < DH / 4; ++i) {
          const float4 uu = __ldg(u4 + i), vv = __ldg(v4 + i);
          v[4 * i] = fmaf(v[4 * i], f_rstd, fmaf(f_nmr, uu.x, vv.x));
          v[4 * i + 1] = fmaf(v[4 * i + 1], f_rstd, fmaf(f_nmr, uu.y, vv.y));
          v[4 * i + 2] = fmaf(v[4 * i + 2], f_rstd, fmaf(f_nmr, uu.z, vv.z));
          v[4 * i + 3] = fmaf(v[4 * i + 3], f_rstd, fmaf(f_nmr, uu.w, vv.w));
        }
      }
      const size_t bh = (size_t)b * ep.H + head;
      if (kind < 2) {
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < DH; ++i) { s1[i & 3] += v[i]; s2[i & 3] = fmaf(v[i], v[i], s2[i & 3]); }
        const float mean = ((s1[0] + s1[1]) + (s1[2] + s1[3])) * (1.0f / DH);
        const float var = fmaxf(((s2[0] + s2[1]) + (s2[2] + s2[3])) * (1.0f / DH) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;
        if (DBG && (ep.dbg & 2)) {
        } else if (kind == 0) {
#pragma unroll
          for (int i = 0; i < DH; ++i) v[i] = fmaf(fmaf(v[i], rstd, nmr), ep.nw[0][i], ep.nb[0][i]);
        } else {
#pragma unroll
          for (int i = 0; i < DH; ++i) v[i] = fmaf(fmaf(v[i], rstd, nmr), ep.nw[1][i], ep.nb[1][i]);
        }
        if (ep.rope != nullptr && ((ep.rope_kinds >> kind) & 1) && !(DBG && (ep.dbg & 1))) {
          const float2* cs = ep.rope + (size_t)l * (DH / 2);
          const float lf = (float)l;
#pragma unroll
          for (int i = 0; i < DH / 2; ++i) {
            float2 c;
            if (ep.rope_mufu) __sincosf(lf * ep.inv_freq[i], &c.y, &c.x);
            else c = __ldg(cs + i);
            const float a = v[i], bq = v[i + DH / 2];
            v[i] = a * c.x - bq * c.y;
            v[i + DH / 2] = bq * c.x + a * c.y;
          }
        }
        if constexpr (DIRECT) {
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(ep.out[kind] + (bh * ep.L + l) * (size_t)ep.ld_qk);
#pragma unroll
            for (int g = 0; g < DH / 8; ++g)
              dst[g] = make_uint4(pack_bf16(v[8 * g], v[8 * g + 1]), pack_bf16(v[8 * g + 2], v[8 * g + 3]), pack_bf16(v[8 * g + 4], v[8 * g + 5]),
                                  pack_bf16(v[8 * g + 6], v[8 * g + 7]));
          }
          return;
        }
        // bf16 pairs -> staging granules (4 bf16 each) -> coalesced row stores
#pragma unroll
        for (int g = 0; g < DH / 4; ++g)
          stage_put(st, lane, g, __uint_as_float(pack_bf16(v[4 * g], v[4 * g + 1])), __uint_as_float(pack_bf16(v[4 * g + 2], v[4 * g + 3])));
        __syncwarp();
        if (lane < DH / 4 && !(DBG && (ep.dbg & 8))) {
          const int rb0 = row0 / ep.L, rl0 = row0 - rb0 * ep.L;
          __nv_bfloat16* base = ep.out[kind] + 4 * lane;
          const size_t head_rows = (size_t)ep.L * ep.ld_qk;
          const int nv = nvalid < 32 ? nvalid : 32;
#pragma unroll 4
          for (int rr = 0; rr < nv; ++rr) {
            int rl = rl0 + rr, rb = rb0;
            while (rl >= ep.L) { rl -= ep.L; ++rb; }   // normally at most one wrap (L >= 32); short contexts may wrap more
            *reinterpret_cast<float2*>(base + ((size_t)rb * ep.H + head) * head_rows + (size_t)rl * ep.ld_qk) = stage_get(st, rr, lane);
          }
        }
      } else if (row_ok && !(DBG && (ep.dbg & 4))) {
        __nv_bfloat16* dst = ep.out[2] + bh * ep.dvp * ep.Lpad + l;
#pragma unroll
        for (int i = 0; i < DH; ++i) { *dst = __float2bfloat16_rn(v[i]); dst += ep.Lpad; }
        for (int i = DH; i < ep.dvp; ++i) { *dst = __float2bfloat16_rn(0.f); dst += ep.Lpad; }
      }
    }
  }
};

}  // namespace ezb

namespace ezb {
// ---------------------------------------------------------------------------------------------------------------
// CTA-pair GEMM (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN tile.  CTA r stages its own
// 128 rows of A and rows [r*BN/2, (r+1)*BN/2) of the W tile; the leader's single MMA thread issues M=256 instructions that
// read both CTAs' shared memory, so each SM pulls half the operand bytes per flop through L2 (the 128x128 single-CTA tile
// is L2->smem bound at ~64 flop/B).  Accumulator rows 128r..128r+127 live in CTA r's TMEM; both CTAs run the epilogue.
// FIRST_PHASE: the body is followed by another GEMM phase in the same kernel (mlp_fused_kernel): keep the TMEM allocation permit and
// invalidate the mbarriers so that the next phase may lay out its own in the same shared memory.
template <int BN, class Epi, int KSUB, bool FIRST_PHASE = false>
__device__ __forceinline__ void gemm2_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& g, const typename Epi::Params& ep, uint8_t* smem_raw) {
  static_assert(BN % 16 == 0 && BN >= 64 && BN <= 256, "BN");
  using SM = GemmCfg<BN, Epi, true, KSUB>;
  constexpr int STAGES = SM::STAGES;
  constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (a round trip through uintptr_t loses the address space and
  // turns every staging access into a generic LD.E / ST.E)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * SM::A_BYTES;
  float* sStage = reinterpret_cast<float*>(sB + STAGES * SM::B_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * SM::B_BYTES + SM::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = g.num_m_tiles * g.num_n_tiles;  // num_m_tiles counts 256-row tiles here

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * 32 * SM::EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<TMEM_COLS, !FIRST_PHASE>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch();
  if (warp == 0 && g.pf_bytes != 0) {   // constant data: no need to wait for the previous kernel
    if (elect_one()) prefetch_weights_l2(g.pf, g.pf_bytes, blockIdx.x, gridDim.x);
    __syncwarp();
  }
  pdl_wait();
  EZB_DBG(const bool dbg = g.dbg != nullptr && blockIdx.x == 0; const long long t_start = clock64(); long long w0 = 0, w1 = 0;)

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (both CTAs; bytes are credited to the leader's barrier)
    {
      uint32_t stage = 0, phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int mt = tile % g.num_m_tiles, nt = tile / g.num_m_tiles;
        const int m0 = mt * 2 * GEMM_BM + (int)rank * GEMM_BM, n0 = nt * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < g.num_k_blocks; kb += KSUB) {
          EZB_DBG(const long long tq = clock64();)
          mbar_wait(&empty[stage], phase ^ 1);
          EZB_DBG(w0 += clock64() - tq;)
          const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
          const int nsub = (g.num_k_blocks - kb) < KSUB ? (g.num_k_blocks - kb) : KSUB;
          if (elect_one()) {
            if (leader) mbar_expect_tx(&full[stage], 2 * nsub * (SM::A_SUB + SM::B_SUB));
            for (int sub = 0; sub < nsub; ++sub) {
              tma_load_2d_pair(sA + stage * SM::A_BYTES + sub * SM::A_SUB, &tmA, bar, (kb + sub) * GEMM_BK, m0);
              tma_load_2d_pair(sB + stage * SM::B_BYTES + sub * SM::B_SUB, &tmB, bar, (kb + sub) * GEMM_BK, n0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, BN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        EZB_DBG(long long tq = clock64();)
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        EZB_DBG(w1 += clock64() - tq;)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < g.num_k_blocks; kb += KSUB) {
          EZB_DBG(tq = clock64();)
          mbar_wait(&full[stage], phase);
          EZB_DBG(w0 += clock64() - tq;)
          tc_fence_after();
          if (elect_one()) {
            const int nsub = (g.num_k_blocks - kb) < KSUB ? (g.num_k_blocks - kb) : KSUB;
            for (int sub = 0; sub < nsub; ++sub) {
              const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * SM::A_BYTES + sub * SM::A_SUB));
              const uint64_t bd = umma_desc_sw128(smem_u32(sB + stage * SM::B_BYTES + sub * SM::B_SUB));
#pragma unroll
              for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_pair(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | sub | k) != 0);
            }
            umma_commit_pair(&empty[stage]);
            if (kb + KSUB >= g.num_k_blocks) umma_commit_pair(&tfull[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------ epilogue (both CTAs, own 128 accumulator rows)
    const int lg = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int mt = tile % g.num_m_tiles, nt = tile / g.num_m_tiles;
      const int row0 = mt * 2 * GEMM_BM + (int)rank * GEMM_BM + lg * 32;
      const uint32_t taddr_row = tmem_base + acc * BN + (static_cast<uint32_t>(lg * 32) << 16);
      constexpr int CW = BN / (SM::EPI_WARPS / 4);
      const int ch = (warp - 2) >> 2;
      uint64_t* tf = &tfull[acc];
      const uint32_t ph = acc_phase;
      EZB_DBG(const long long te = clock64(); long long tw = 0;)
      Epi::run(ep, sStage + (warp - 2) * Epi::STAGE_FLOATS, taddr_row, row0, g.M - row0, nt * BN, g.N, lane, ch * CW, (ch + 1) * CW, [&]() {
        EZB_DBG(const long long tq = clock64();)
        mbar_wait(tf, ph);
        EZB_DBG(tw = clock64() - tq;)
        tc_fence_after();
      });
      EZB_DBG(w0 += tw; w1 += clock64() - te - tw;)
      tc_fence_before();
      mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  EZB_DBG(if (dbg && lane == 0) {
    if (warp == 0) atomicAdd(&g.dbg[2], (unsigned long long)w0);
    if (warp == 1) { atomicAdd(&g.dbg[0], (unsigned long long)w0); atomicAdd(&g.dbg[1], (unsigned long long)w1); }
    if (warp == 2) { atomicAdd(&g.dbg[3], (unsigned long long)w0); atomicAdd(&g.dbg[4], (unsigned long long)w1); atomicAdd(&g.dbg[5], (unsigned long long)(clock64() - t_start)); }
  })
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
  if (FIRST_PHASE) {
    if (warp == 0 && lane == 0) {
      for (int i = 0; i < STAGES; ++i) { mbar_inval(&full[i]); mbar_inval(&empty[i]); }
      for (int i = 0; i < 2; ++i) { mbar_inval(&tfull[i]); mbar_inval(&tempty[i]); }
    }
    __syncthreads();
  }
}
template <int BN, class Epi, int KSUB>
__global__ void __launch_bounds__((GemmCfg<BN, Epi, true, KSUB>::THREADS), 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape g, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_dyn[];
  gemm2_body<BN, Epi, KSUB>(tmA, tmB, g, ep, smem_dyn);
}

// ---------------------------------------------------------------------------------------------------------------
// The MLP of a DiT block (modules.py:263-277,366) as ONE persistent launch: phase 1 = the GEGLU projection (CTA-pair tiles, EpiGeglu), a
// grid-wide barrier, phase 2 = the output projection with its gated-residual epilogue (swap-AB tiles, EpiLinearT, incl. the LayerNorm fold
// outputs for the next block).  The grid is one CTA per SM (all co-resident), launched as clusters of two for phase 1.  Phase 2 reads the
// bf16 intermediate that phase 1 wrote with ordinary stores through TMA, hence the generic->async proxy fence after the barrier.
struct GridBarrier { unsigned int count; unsigned int gen; };
__device__ __forceinline__ void grid_barrier(GridBarrier* b) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();   // this CTA's global writes (all threads, ordered by the barrier above) before the arrive
    unsigned int gen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(&b->gen) : "memory");
    if (atomicAdd(&b->count, 1u) == gridDim.x - 1) {
      b->count = 0;
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&b->gen), "r"(gen + 1) : "memory");
    } else {
      unsigned int cur;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(&b->gen) : "memory");
      } while (cur == gen);
    }
  }
  __syncthreads();
  asm volatile("fence.proxy.async;" ::: "memory");
}
template <int BN1, class Epi1, class Epi2>
__global__ void __launch_bounds__((GemmCfg<BN1, Epi1, true, 1>::THREADS), 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1, const GemmShape g1, const typename Epi1::Params ep1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const GemmShape g2, const typename Epi2::Params ep2,
                 GridBarrier* bar) {
  static_assert(GemmCfg<BN1, Epi1, true, 1>::THREADS == GemmCfg<256, Epi2, false>::THREADS, "both phases use the same warp roles");
  extern __shared__ uint8_t smem_dyn[];
  gemm2_body<BN1, Epi1, 1, true>(tmA1, tmB1, g1, ep1, smem_dyn);
  grid_barrier(bar);
  gemm_body<256, Epi2, 1>(tmA2, tmB2, g2, ep2, smem_dyn);
}

}  // namespace ezb
