// Bandwidth-bound kernels of the DiT step: LayerNorm(+AdaLN modulate)+cast, per-head qk-LayerNorm + RoPE + head layout,
// patch-embed input packing, final 3-tap conv, time-embedding path, weight repacking, CFG + DDIM update.
// All fp32 math; bf16 only where a tensor feeds a tensor-core operand.
#pragma once
#include "common.cuh"

namespace ezb {

// parity ("split") operand write: A' = [hi | lo | hi] along K (matches W' = [hi | hi | lo]): A'W'^T = hi*hi + lo*hi + hi*lo
__device__ __forceinline__ void store_act(__nv_bfloat16* row, int col, int K, int kmul, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  row[col] = hi;
  if (kmul == 3) {
    row[K + col] = __float2bfloat16_rn(v - __bfloat162float(hi));
    row[2 * K + col] = hi;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the (optionally concatenated) row [x | x2 (+x3)], affine, optional AdaLN modulate, cast to bf16.
//   nn.LayerNorm eps 1e-5 (blocks.py:68,83,85,91,100), film_modulate x*(1+scale)+shift (modules.py:15-16),
//   skip path cat[x, skip (+ controlnet skip)] (blocks.py:124-126, udit.py:345-348).
// One warp per row; two passes over an L1/L2-resident row.
struct LnParams {
  const float* x;    // [M, D1]
  const float* x2;   // optional [M, D2] (concatenated after x)
  const float* x3;   // optional, added to x2 element-wise
  int D1, D2;
  const float* w;    // [D1 + D2]
  const float* b;
  const float* shift;  // optional modulation: shift[bidx * mod_bstride + c], scale likewise (c < D1 only, D2 == 0)
  const float* scale;
  int mod_bstride;     // element stride between batch items in shift/scale (0: all share one row)
  int rows_per_batch;
  __nv_bfloat16* out;  // [M, kmul * (D1 + D2)]
  int kmul;
  int M;
  const float* G;      // optional precombined affine for ln_gc_kernel: y = (x - mu) * rstd * G + C with G = w (1 + scale), C = b (1 + scale) + shift
  const float* C;      // (one row for the whole batch: uniform timestep)
};

// x / x2 / x3 are read with ld.global.cg (L2 only): when the LayerNorm runs as the tail phase of the GEMM that produced x (gemm_ln.cuh), rows
// written by other SMs moments ago must not be served from a stale L1 line; for the stand-alone kernels it makes no difference (streaming).
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void ln_row_generic(const LnParams& p, int row, int lane) {
  const int D = p.D1 + p.D2;
  const float* x = p.x + (size_t)row * p.D1;
  const float* x2 = p.x2 ? p.x2 + (size_t)row * p.D2 : nullptr;
  const float* x3 = p.x3 ? p.x3 + (size_t)row * p.D2 : nullptr;
  float s = 0.f;
  for (int c = lane * 4; c < p.D1; c += 128) {
    const float4 v = ldcg4(x + c);
    s += v.x + v.y + v.z + v.w;
  }
  for (int c = lane * 4; c < p.D2; c += 128) {
    float4 v = ldcg4(x2 + c);
    if (x3) { const float4 u = ldcg4(x3 + c); v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
    s += v.x + v.y + v.z + v.w;
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
  for (int c = lane * 4; c < p.D1; c += 128) {
    const float4 v = ldcg4(x + c);
    q += (v.x - mean) * (v.x - mean) + (v.y - mean) * (v.y - mean) + (v.z - mean) * (v.z - mean) + (v.w - mean) * (v.w - mean);
  }
  for (int c = lane * 4; c < p.D2; c += 128) {
    float4 v = ldcg4(x2 + c);
    if (x3) { const float4 u = ldcg4(x3 + c); v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
    q += (v.x - mean) * (v.x - mean) + (v.y - mean) * (v.y - mean) + (v.z - mean) * (v.z - mean) + (v.w - mean) * (v.w - mean);
  }
  float rstd = rsqrtf(warp_sum(q) / D + 1e-5f);
  const bool norm = p.w != nullptr;  // w == NULL: cast only (no normalisation)
  float mu = mean;
  if (!norm) { mu = 0.f; rstd = 1.f; }
  const float *sh = nullptr, *sc = nullptr;
  if (p.shift) {
    const size_t off = (size_t)(row / p.rows_per_batch) * p.mod_bstride;
    sh = p.shift + off;
    sc = p.scale + off;
  }
  __nv_bfloat16* o = p.out + (size_t)row * p.kmul * D;
  for (int c = lane * 4; c < D; c += 128) {
    float4 v;
    if (c < p.D1) {
      v = ldcg4(x + c);
    } else {
      v = ldcg4(x2 + (c - p.D1));
      if (x3) { const float4 u = ldcg4(x3 + (c - p.D1)); v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
    }
    float4 w = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (norm) { w = *reinterpret_cast<const float4*>(p.w + c); b = *reinterpret_cast<const float4*>(p.b + c); }
    float y[4] = {(v.x - mu) * rstd * w.x + b.x, (v.y - mu) * rstd * w.y + b.y, (v.z - mu) * rstd * w.z + b.z, (v.w - mu) * rstd * w.w + b.w};
    if (sh) {
      const float4 a = *reinterpret_cast<const float4*>(sc + c), d = *reinterpret_cast<const float4*>(sh + c);
      y[0] = y[0] * (1.f + a.x) + d.x; y[1] = y[1] * (1.f + a.y) + d.y; y[2] = y[2] * (1.f + a.z) + d.z; y[3] = y[3] * (1.f + a.w) + d.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) store_act(o, c + e, D, p.kmul, y[e]);
  }
}
__global__ void __launch_bounds__(256) ln_mod_cast_kernel(const LnParams p) {
  pdl_launch();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= p.M) return;
  ln_row_generic(p, warp, lane);
}

// Register-resident variant for the hot shapes (single source, D = 128 * NCH): the row is read once, both moments come
// from registers (two-pass formula, same numerics as above), 8-byte bf16 stores.
template <int NCH>
__device__ __forceinline__ void ln_row_reg(const LnParams& p, int row, int lane) {
  constexpr int D = NCH * 128;
  const float* x = p.x + (size_t)row * D;
  float4 v[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) v[i] = ldcg4(x + 4 * (lane + 32 * i));
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
  const float4 *w4 = reinterpret_cast<const float4*>(p.w), *b4 = reinterpret_cast<const float4*>(p.b);
  const float4 *sh4 = nullptr, *sc4 = nullptr;
  if (p.shift) {
    const size_t off = (size_t)(row / p.rows_per_batch) * p.mod_bstride;
    sh4 = reinterpret_cast<const float4*>(p.shift + off);
    sc4 = reinterpret_cast<const float4*>(p.scale + off);
  }
  __nv_bfloat16* o = p.out + (size_t)row * D;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c4 = lane + 32 * i;
    const float4 w = __ldg(w4 + c4), b = __ldg(b4 + c4);
    float y0 = (v[i].x - mean) * rstd * w.x + b.x, y1 = (v[i].y - mean) * rstd * w.y + b.y;
    float y2 = (v[i].z - mean) * rstd * w.z + b.z, y3 = (v[i].w - mean) * rstd * w.w + b.w;
    if (sh4) {
      const float4 a = __ldg(sc4 + c4), d = __ldg(sh4 + c4);
      y0 = y0 * (1.f + a.x) + d.x; y1 = y1 * (1.f + a.y) + d.y; y2 = y2 * (1.f + a.z) + d.z; y3 = y3 * (1.f + a.w) + d.w;
    }
    *reinterpret_cast<uint2*>(o + 4 * c4) = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
  }
}
// One warp per row, 4 rows per 128-thread block.  MINB = minimum CTAs per SM: 1 -> 79 registers, 6 CTAs per SM = 888 of the 1000 CTAs of an
// M = 4000 launch resident (a second, nearly empty wave); 8 -> <= 64 registers, one wave.  Run-time choice (option "ln_variant").
template <int NCH, int MINB>
__global__ void __launch_bounds__(128, MINB) ln_mod_cast_reg_kernel(const LnParams p) {
  pdl_launch();
  pdl_wait();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= p.M) return;
  ln_row_reg<NCH>(p, row, lane);
}
// Variant with the affine precombined per timestep (G, C: fold_gc_kernel) and held in REGISTERS across rows.  In the kernel above every warp
// re-reads weight, bias, scale and shift (4 x 4.6 KB) for its one row: 27 warps per SM pull ~500 KB of parameters through L1 for 124 KB of
// activations.  Here a warp loads G and C once (2 x 36 registers per lane) and walks its rows in a strided loop.
template <int NCH>
__global__ void __launch_bounds__(128, 4) ln_gc_kernel(const LnParams p) {
  pdl_launch();
  pdl_wait();
  constexpr int D = NCH * 128;
  const int lane = threadIdx.x & 31, gw = blockIdx.x * 4 + (threadIdx.x >> 5), nw = gridDim.x * 4;
  if (gw >= p.M) return;
  float4 g[NCH], c[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    g[i] = __ldg(reinterpret_cast<const float4*>(p.G) + lane + 32 * i);
    c[i] = __ldg(reinterpret_cast<const float4*>(p.C) + lane + 32 * i);
  }
  for (int row = gw; row < p.M; row += nw) {
    const float* x = p.x + (size_t)row * D;
    float4 v[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) v[i] = ldcg4(x + 4 * (lane + 32 * i));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
    __nv_bfloat16* o = p.out + (size_t)row * D;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const float y0 = fmaf((v[i].x - mean) * rstd, g[i].x, c[i].x), y1 = fmaf((v[i].y - mean) * rstd, g[i].y, c[i].y);
      const float y2 = fmaf((v[i].z - mean) * rstd, g[i].z, c[i].z), y3 = fmaf((v[i].w - mean) * rstd, g[i].w, c[i].w);
      *reinterpret_cast<uint2*>(o + 4 * (lane + 32 * i)) = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
    }
  }
}
// skip_norm of the out-blocks (blocks.py:124-126): LayerNorm over the concatenated row [x | x2 (+ x3)] with D1 = D2 = 128 NCH, held in registers (one
// pass; the generic kernel above walks the row three times).  No modulation.  One warp per row, 4 rows per 128-thread block.
template <int NCH>
__global__ void __launch_bounds__(128) ln_cat_reg_kernel(const LnParams p) {
  pdl_launch();
  pdl_wait();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= p.M) return;
  constexpr int D1 = NCH * 128, D = 2 * D1;
  const float* x = p.x + (size_t)row * D1;
  const float* x2 = p.x2 + (size_t)row * D1;
  float4 v[2 * NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    v[i] = ldcg4(x + 4 * (lane + 32 * i));
    v[NCH + i] = ldcg4(x2 + 4 * (lane + 32 * i));
  }
  if (p.x3 != nullptr) {
    const float* x3 = p.x3 + (size_t)row * D1;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const float4 u = ldcg4(x3 + 4 * (lane + 32 * i));
      v[NCH + i].x += u.x; v[NCH + i].y += u.y; v[NCH + i].z += u.z; v[NCH + i].w += u.w;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * NCH; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * NCH; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
  const float4 *w4 = reinterpret_cast<const float4*>(p.w), *b4 = reinterpret_cast<const float4*>(p.b);
  __nv_bfloat16* o = p.out + (size_t)row * D;
#pragma unroll
  for (int i = 0; i < 2 * NCH; ++i) {
    const int c4 = (i < NCH ? 0 : D1 / 4) + lane + 32 * (i < NCH ? i : i - NCH);
    const float4 w = __ldg(w4 + c4), b = __ldg(b4 + c4);
    const float y0 = (v[i].x - mean) * rstd * w.x + b.x, y1 = (v[i].y - mean) * rstd * w.y + b.y;
    const float y2 = (v[i].z - mean) * rstd * w.z + b.z, y3 = (v[i].w - mean) * rstd * w.w + b.w;
    *reinterpret_cast<uint2*>(o + 4 * c4) = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
  }
}
// LayerNorm as the tail phase of another kernel (gemm_ln.cuh): every warp of the (persistent, fully resident) grid takes rows in a strided loop
__device__ __forceinline__ bool ln_reg_eligible(const LnParams& p) { return p.kmul == 1 && p.x2 == nullptr && p.w != nullptr && (p.D1 == 1152 || p.D1 == 1024); }
__device__ __forceinline__ void ln_tail(const LnParams& p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const bool reg = ln_reg_eligible(p);
  for (int row = blockIdx.x * nw + warp; row < p.M; row += gridDim.x * nw) {
    if (reg) {
      if (p.D1 == 1152) ln_row_reg<9>(p, row, lane);
      else ln_row_reg<8>(p, row, lane);
    } else {
      ln_row_generic(p, row, lane);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Per-head LayerNorm(dh, affine) on q / k (attention.py:63-65,141-142), NeoX rotate-half RoPE with fp32 tables on
// positions 0..L-1 (rotary.py:6-18,72-84; pairs (i, i+dh/2), freq 1e4^(-2i/dh)), and head-major layout for attention:
//   section 0 (q), 1 (k): dst[b, h, l, 0..dh)  (row pitch dst_ld, zero padding beyond dh pre-cleared)
//   section 2 (v)       : fp32 path -> same layout; tensor-core path -> transposed vt[b, h, d, l] (pitch Lpad)
// One warp per (token, head); lane owns dims lane, lane+32, lane+64 (dh <= 96).
template <typename TIn>
struct QkPrepParams {
  const TIn* in;  // [M, ld_in]: GEMM output, sections at column offsets col_off[s]
  int ld_in;
  int col_off[3];
  int n_sections;        // self: 3 (q,k,v); cross q: 1; cross kv: sections k,v -> use sec_kind
  int sec_kind[3];       // 0 q, 1 k, 2 v
  const float* nw[2];    // LN weight for q, k
  const float* nb[2];
  int use_rope;
  const float* inv_freq;  // [dh/2] (attn.rotary.inv_freq buffer of the checkpoint)
  int B, L, H, dh;
  float* f32_out[3];            // optional fp32 [B,H,L,dh] per section (SIMT attention)
  __nv_bfloat16* bf_out[3];     // optional bf16: q,k -> [B,H,L,ld_qk]; v -> vt [B,H,dv_pad,Lpad]
  int ld_qk, Lpad, dv_pad;
};

__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename TIn>
__global__ void __launch_bounds__(256) qk_prep_kernel(const QkPrepParams<TIn> p) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = p.B * p.L * p.H * p.n_sections;
  if (wid >= total) return;
  const int h = wid % p.H;
  const int sec = (wid / p.H) % p.n_sections;
  const int tok = wid / (p.H * p.n_sections);
  const int b = tok / p.L, l = tok - b * p.L;
  const int kind = p.sec_kind[sec];
  const TIn* src = p.in + (size_t)tok * p.ld_in + p.col_off[sec] + h * p.dh;
  float v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = (lane + 32 * i < p.dh) ? ld_as_float(src + lane + 32 * i) : 0.f;
  if (kind < 2) {
    const float mean = warp_sum(v[0] + v[1] + v[2]) / p.dh;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) if (lane + 32 * i < p.dh) q += (v[i] - mean) * (v[i] - mean);
    const float rstd = rsqrtf(warp_sum(q) / p.dh + 1e-5f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int d = lane + 32 * i;
      if (d < p.dh) v[i] = (v[i] - mean) * rstd * p.nw[kind][d] + p.nb[kind][d];
    }
    if (p.use_rope) {
      // stage through shuffles is awkward for arbitrary dh: use a small smem exchange per warp instead
      __shared__ float xch[8][96];
      float* xw = xch[threadIdx.x >> 5];
#pragma unroll
      for (int i = 0; i < 3; ++i) if (lane + 32 * i < p.dh) xw[lane + 32 * i] = v[i];
      __syncwarp();
      const int half = p.dh >> 1;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int d = lane + 32 * i;
        if (d < p.dh) {
          const int fi = d < half ? d : d - half;
          float sn, cs;
          sincosf((float)l * p.inv_freq[fi], &sn, &cs);
          const float other = d < half ? -xw[d + half] : xw[d - half];
          v[i] = v[i] * cs + other * sn;
        }
      }
    }
  }
  const size_t bh = (size_t)b * p.H + h;
  if (p.f32_out[sec]) {
    float* o = p.f32_out[sec] + (bh * p.L + l) * p.dh;
#pragma unroll
    for (int i = 0; i < 3; ++i) if (lane + 32 * i < p.dh) o[lane + 32 * i] = v[i];
  }
  if (p.bf_out[sec]) {
    if (kind < 2) {
      __nv_bfloat16* o = p.bf_out[sec] + (bh * p.L + l) * p.ld_qk;
#pragma unroll
      for (int i = 0; i < 3; ++i) if (lane + 32 * i < p.dh) o[lane + 32 * i] = __float2bfloat16_rn(v[i]);
    } else {
      __nv_bfloat16* o = p.bf_out[sec] + bh * p.dv_pad * p.Lpad + l;
#pragma unroll
      for (int i = 0; i < 3; ++i) if (lane + 32 * i < p.dh) o[(size_t)(lane + 32 * i) * p.Lpad] = __float2bfloat16_rn(v[i]);
      for (int dpad = p.dh + lane; dpad < p.dv_pad; dpad += 32) o[(size_t)dpad * p.Lpad] = __float2bfloat16_rn(0.f);  // zero pad rows
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// MaskDiT input assembly (conditioners.py:150-153,174-176) + transpose for the k=1 patch-embed conv (modules.py:100-111):
//   A[b*L + l, :] = [ x[b,:,l] | gt[b,:,l] or mask_embed (where gt is NULL or gt_mask[b,l]) | mask channel | 0-pad ]
//   mask channel = gt ? gt_mask[b,l] : 1          (mae_mask[:,0:1,:]: ones when gt is None)
__global__ void patch_pack_kernel(const float* __restrict__ x, const float* __restrict__ gt, const uint8_t* __restrict__ gt_mask,
                                  const float* __restrict__ mask_embed, __nv_bfloat16* __restrict__ out, int B, int C, int L, int Kp, int kmul) {
  pdl_launch();
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;  // c0 over 2C channels
  const int tx = threadIdx.x, ty = threadIdx.y;                            // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + tx;
    float v = 0.f;
    if (l < L) {
      if (c < C) v = x[((size_t)b * C + c) * L + l];
      else {
        const int cg = c - C;
        const bool masked = (gt == nullptr) || (gt_mask != nullptr && gt_mask[(size_t)b * L + l]);
        v = masked ? mask_embed[cg] : gt[((size_t)b * C + cg) * L + l];
      }
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int l = l0 + i, c = c0 + tx;
    if (l < L) store_act(out + ((size_t)b * L + l) * kmul * Kp, c, Kp, kmul, tile[tx][i]);
  }
  if (blockIdx.y == 0 && ty == 0) {  // mask channel + zero pad
    const int l = l0 + tx;
    if (l < L) {
      const float m = gt ? (gt_mask && gt_mask[(size_t)b * L + l] ? 1.f : 0.f) : 1.f;
      __nv_bfloat16* o = out + ((size_t)b * L + l) * kmul * Kp;
      store_act(o, 2 * C, Kp, kmul, m);
      for (int c = 2 * C + 1; c < Kp; ++c) store_act(o, c, Kp, kmul, 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// FinalBlock tail (blocks.py:207-211): y [B*L, C] token-major -> unpatchify (transpose) -> Conv1d(C, C, k=3, pad=1).
// w packed [3][Cin][Cout].  CTA = 32 positions x all C = 128 output channels, 512 threads = 4 input-channel groups x 128; within a group
// thread = 4 output channels x 8 positions, so one input channel costs 10 broadcast LDS + 3 coalesced float4 weight loads for 96 FMAs; the four
// groups' partial sums meet in shared memory.  (Round 1: one dependent global weight load per (tap, channel), 211 us; round 2a: 128 threads
// walking all 128 input channels, 62.5 us per step -- four warps per SM cannot hide the L2 latency of the weight loads; this one: 16 warps.)
constexpr int FC_GROUPS = 4;
__global__ void __launch_bounds__(128 * FC_GROUPS) final_conv_kernel(const float* __restrict__ y, const float* __restrict__ wp, const float* __restrict__ bias,
                                                                     float* __restrict__ out, int B, int C, int L) {
  pdl_launch();
  pdl_wait();
  constexpr int TL = 32, TP = TL + 2 + 2;  // 34 positions (+2 so that the 10-wide window of the last thread group stays in bounds)
  extern __shared__ float sy[];             // [C][TP] channel-major, then [FC_GROUPS - 1][128][32] partial sums
  float* red = sy + (size_t)C * TP;
  const int b = blockIdx.y, l0 = blockIdx.x * TL;
  for (int i = threadIdx.x; i < (TL + 2) * C; i += blockDim.x) {
    const int r = i / C, c = i - r * C, l = l0 + r - 1;
    sy[c * TP + r] = (l >= 0 && l < L) ? y[((size_t)b * L + l) * C + c] : 0.f;
  }
  __syncthreads();
  const int gq = threadIdx.x >> 7, t128 = threadIdx.x & 127;
  const int tq = t128 >> 5;                  // positions tq*8 .. tq*8+7 of the tile
  const int cq = (C + FC_GROUPS - 1) / FC_GROUPS, ci0 = gq * cq, ci1 = ci0 + cq < C ? ci0 + cq : C;
  for (int cb = 0; cb < C; cb += 128) {   // C = 128 shipped: one pass.  The trip count is uniform (barriers inside); lanes beyond C idle.
    const int co = cb + (t128 & 31) * 4;
    const bool live = co < C;
    float acc[8][4];
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gq == 0 && live) bv = *reinterpret_cast<const float4*>(bias + co);
#pragma unroll
    for (int t = 0; t < 8; ++t) { acc[t][0] = bv.x; acc[t][1] = bv.y; acc[t][2] = bv.z; acc[t][3] = bv.w; }
#pragma unroll 4
    for (int ci = ci0; ci < (live ? ci1 : ci0); ++ci) {
      float4 w[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) w[k] = __ldg(reinterpret_cast<const float4*>(wp + ((size_t)k * C + ci) * C + co));
      float x[10];
#pragma unroll
      for (int t = 0; t < 10; ++t) x[t] = sy[ci * TP + tq * 8 + t];
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          acc[t][0] = fmaf(w[k].x, x[t + k], acc[t][0]);
          acc[t][1] = fmaf(w[k].y, x[t + k], acc[t][1]);
          acc[t][2] = fmaf(w[k].z, x[t + k], acc[t][2]);
          acc[t][3] = fmaf(w[k].w, x[t + k], acc[t][3]);
        }
    }
    if (gq > 0) {
      float* rr = red + ((size_t)(gq - 1) * 128 + t128) * 32;
#pragma unroll
      for (int t = 0; t < 8; ++t) *reinterpret_cast<float4*>(rr + t * 4) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
    }
    __syncthreads();
    if (gq == 0 && live) {
#pragma unroll
      for (int q = 0; q < FC_GROUPS - 1; ++q) {
        const float* rr = red + ((size_t)q * 128 + t128) * 32;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float4 v = *reinterpret_cast<const float4*>(rr + t * 4);
          acc[t][0] += v.x; acc[t][1] += v.y; acc[t][2] += v.z; acc[t][3] += v.w;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int l = l0 + tq * 8 + t;
          if (l < L) out[((size_t)b * C + co + e) * L + l] = acc[t][e];
        }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Small-M linear for the time path (fp32 weights, R <= a few hundred rows): out[r, n] = act(in[r,:] . W[n,:] + bias[n]) + add[r, n]
// One warp per output feature; the weight row lives in registers and is reused across all rows.
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ W, const float* __restrict__ bias,
                                                           const float* __restrict__ add, int ld_add, float* __restrict__ out, int ld_out, int R, int N, int K,
                                                           int act /*0 none, 1 silu*/, float out_scale) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  constexpr int MAXK = 40;  // K <= 1280 per pass
  for (int k0 = 0; k0 < K; k0 += 32 * MAXK) {
    float w[MAXK];
#pragma unroll
    for (int i = 0; i < MAXK; ++i) { const int k = k0 + lane + 32 * i; w[i] = k < K ? W[(size_t)n * K + k] : 0.f; }
    for (int r = 0; r < R; ++r) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAXK; ++i) { const int k = k0 + lane + 32 * i; if (k < K) s = fmaf(w[i], in[(size_t)r * ld_in + k], s); }
      s = warp_sum(s);
      if (lane == 0) {
        float* o = out + (size_t)r * ld_out + n;
        float v = (k0 == 0 ? 0.f : *o) + s;
        if (k0 + 32 * MAXK >= K) {
          v = v * out_scale + (bias ? bias[n] : 0.f);
          if (act == 1) v = silu(v);
          if (add) v += add[(size_t)r * ld_add + n];
        }
        *o = v;
      }
    }
  }
}

// RoPE table (rotary.py:48-70): cs[l][i] = (cos, sin)(l * inv_freq[i]), fp32, i < dh/2 (token-major: a thread reads its
// token's dh/2 pairs as 9 full 32-byte sectors; the frequency-major alternative measured slower, profiles/r1)
__global__ void rope_table_kernel(const float* __restrict__ inv_freq, float2* __restrict__ cs, int L, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * half) return;
  const int l = i / half, f = i - l * half;
  float s, c;
  sincosf((float)l * inv_freq[f], &s, &c);
  cs[i] = make_float2(c, s);
}
// timestep_embedding (modules.py:19-39): [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(1e4) i / 128), dim 256
__global__ void timestep_embed_kernel(const float* __restrict__ t, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 128) return;
  const int r = i / 128, j = i - r * 128;
  const float f = expf(-9.210340371976184f * (float)j / 128.0f);
  float s, c;
  sincosf(t[r] * f, &s, &c);
  out[r * 256 + j] = c;
  out[r * 256 + 128 + j] = s;
}
__global__ void silu_inplace_kernel(float* x, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = silu(x[i]);
}
// mod[r, blk, :] += table[blk, :]  (scale_shift_table[None] + time_ada, blocks.py:43-45)
__global__ void add_rowvec_kernel(float* __restrict__ dst, int ld_dst, const float* __restrict__ vec, int R, int N) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * N) return;
  const int r = i / N, c = i - (size_t)r * N;
  dst[(size_t)r * ld_dst + c] += vec[c];
}

// ---------------------------------------------------------------------------------------------------------------
// Weight repacking fp32 [N, K] (reference layout) -> bf16 [N', kmul*Kpad] rows; optional GEGLU tile interleave
// (dst row = tile*BN + {0,HALF} + j) and row offset (QKV / KV concatenation).  Split mode writes W' = [hi | hi | lo].
// heads3 mode (h3_dh > 0): rows are heads of one of the q / k / v sections; global head g = h3_head_off + n / dh goes to row
// (g / 3) * h3_bn + (g % 3) * dh + n % dh of the packed QKV weight (three heads per N-tile, see EpiHeads<DH, 3>).
__global__ void pack_weight_kernel(const float* __restrict__ src, int N, int K, __nv_bfloat16* __restrict__ dst, int Kpad, int kmul, int row_off,
                                   int geglu_inner, int geglu_half, int h3_dh, int h3_head_off, int h3_bn) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * Kpad) return;
  const int n = i / Kpad, k = i - (size_t)n * Kpad;
  int dn = n + row_off;
  if (geglu_inner > 0) {
    const int g = n >= geglu_inner, m = g ? n - geglu_inner : n;
    dn = (m / geglu_half) * (2 * geglu_half) + g * geglu_half + (m % geglu_half);
  }
  if (h3_dh > 0) {
    const int g = h3_head_off + n / h3_dh;
    dn = (g / 3) * h3_bn + (g % 3) * h3_dh + n % h3_dh;
  }
  const float v = k < K ? src[(size_t)n * K + k] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  __nv_bfloat16* o = dst + (size_t)dn * kmul * Kpad;
  o[k] = hi;
  if (kmul == 3) {
    o[Kpad + k] = hi;
    o[2 * Kpad + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}
__global__ void pack_geglu_bias_kernel(const float* __restrict__ src, float* __restrict__ dst, int inner, int half) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= 2 * inner) return;
  const int g = n >= inner, m = g ? n - inner : n;
  dst[(m / half) * (2 * half) + g * half + (m % half)] = src[n];
}
// generic strided permute copy fp32: dst[a, b, c] = src[...]: used for conv weight transposes at load time
__global__ void permute3_kernel(const float* __restrict__ src, float* __restrict__ dst, int d0, int d1, int d2, int s0, int s1, int s2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)d0 * d1 * d2) return;
  const int c = i % d2, b = (i / d2) % d1, a = i / ((size_t)d1 * d2);
  dst[i] = src[(size_t)a * s0 + (size_t)b * s1 + (size_t)c * s2];
}

// ---------------------------------------------------------------------------------------------------------------
// Classifier-free guidance + rescale (src/inference.py:12-23,88-93) fused with the DDIM v-prediction update
// (diffusers DDIMScheduler.step).  coef = {sqrt(a), sqrt(1-a), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma}.
// One CLUSTER of CFG_CLUSTER CTAs per sample (the first version ran one CTA per sample: 4 of 148 SMs busy, 77 us per step):
// every CTA reduces the four sums of its slice (double accumulation, fixed order -> deterministic), the partials are exchanged
// through distributed shared memory, every CTA forms the same ratio and updates its slice.
constexpr int CFG_CLUSTER = 8;
__device__ __forceinline__ double ld_dsmem_f64(const double* local, uint32_t rank) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(mapa_u32(smem_u32(local), rank)));
  return v;
}
__global__ void __launch_bounds__(1024) cfg_ddim_kernel(const float* __restrict__ out_text, const float* __restrict__ out_uncond, float* __restrict__ latents,
                                                        const float* __restrict__ noise, int n, float gs, float gr, float c0, float c1, float c2, float c3,
                                                        float c4) {
  __shared__ double red[4][32];
  __shared__ double part[4];
  pdl_launch();
  pdl_wait();
  const uint32_t rank = cluster_ctarank();
  const int sample = blockIdx.x / CFG_CLUSTER;
  const size_t base = (size_t)sample * n;
  const int per = (((n + CFG_CLUSTER - 1) / CFG_CLUSTER) + 3) & ~3;   // slice of this CTA, multiple of 4 elements
  const int lo = (int)rank * per, hi = (lo + per < n) ? lo + per : n;
  const float* t = out_text + base;
  const float* u = out_uncond ? out_uncond + base : nullptr;
  float ratio = 1.f;
  const bool rescale = u && gr > 0.f;   // uniform over the grid
  if (rescale) {
    double st = 0, st2 = 0, sc = 0, sc2 = 0;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const float a = t[i], c = u[i] + gs * (a - u[i]);
      st += a; st2 += (double)a * a; sc += c; sc2 += (double)c * c;
    }
    double v[4] = {st, st2, sc, sc2};
    for (int k = 0; k < 4; ++k) {
      double x = v[k];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = x;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      double s = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
      part[threadIdx.x] = s;
    }
    cluster_sync_all();   // partials of all CTAs of this sample are visible cluster-wide
    double s[4] = {0, 0, 0, 0};
    for (uint32_t r = 0; r < CFG_CLUSTER; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] += ld_dsmem_f64(&part[k], r);
    const double var_t = (s[1] - s[0] * s[0] / n) / (n - 1), var_c = (s[3] - s[2] * s[2] / n) / (n - 1);
    ratio = (float)(sqrt(var_t) / sqrt(var_c));
    cluster_sync_all();   // nobody exits (releasing its shared memory) while a peer may still read its partials
  }
  float* x = latents + base;
  const float* z = noise ? noise + base : nullptr;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float v = t[i];
    if (u) {
      v = u[i] + gs * (v - u[i]);
      if (gr > 0.f) v = gr * (v * ratio) + (1.f - gr) * v;
    }
    const float xi = x[i];
    const float x0 = c0 * xi - c1 * v, eps = c0 * v + c1 * xi;
    float prev = c2 * x0 + c3 * eps;
    if (z) prev += c4 * z[i];
    x[i] = prev;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tables of the folded LayerNorm (gemm.cuh, FoldIn / FoldOut), built once per schedule by ezb_dit_set_timesteps:
//   G[t][k] = w[k] (1 + scale_t[k]),  C[t][k] = b[k] (1 + scale_t[k]) + shift_t[k]            (fold_gc_kernel; no modulation: G = w, C = b)
//   U[t][n] = sum_k G[t][k] W[n][k],  V[t][n] = sum_k C[t][k] W[n][k] (+ bias[n])             (fold_uv_kernel; W = the PACKED bf16 weight the
//   GEMM itself multiplies with, so U and V are in the GEMM's own column order and consistent with its operand rounding)
__global__ void fold_gc_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ shift, const float* __restrict__ scale,
                               int ld_mod, float* __restrict__ G, float* __restrict__ Cc, int R, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * D) return;
  const int r = i / D, k = i - r * D;
  const float sc = scale ? scale[(size_t)r * ld_mod + k] : 0.f, sh = shift ? shift[(size_t)r * ld_mod + k] : 0.f;
  G[i] = w[k] * (1.f + sc);
  Cc[i] = b[k] * (1.f + sc) + sh;
}
__global__ void __launch_bounds__(256) fold_uv_kernel(const __nv_bfloat16* __restrict__ W, int ldw, const float* __restrict__ G, const float* __restrict__ Cc,
                                                      const float* __restrict__ add_v, float* __restrict__ U, float* __restrict__ V, int N, int K, int R) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  constexpr int MAXK = 72;  // K <= 2304 (the normalised width: D, or 2 D on the skip path)
  float w[MAXK];
#pragma unroll
  for (int i = 0; i < MAXK; ++i) { const int k = lane + 32 * i; w[i] = k < K ? __bfloat162float(W[(size_t)n * ldw + k]) : 0.f; }
  const float add = add_v ? add_v[n] : 0.f;
  for (int r = 0; r < R; ++r) {
    float su = 0.f, sv = 0.f;
#pragma unroll
    for (int i = 0; i < MAXK; ++i) {
      const int k = lane + 32 * i;
      if (k < K) { su = fmaf(w[i], G[(size_t)r * K + k], su); sv = fmaf(w[i], Cc[(size_t)r * K + k], sv); }
    }
    su = warp_sum(su); sv = warp_sum(sv);
    if (lane == 0) { U[(size_t)r * N + n] = su; V[(size_t)r * N + n] = sv + add; }
  }
}

// timestep values of ezb_dit_set_timesteps, passed BY VALUE in chunks (no host staging buffer, so the call needs no synchronisation)
struct TimestepChunk { float v[240]; };
__global__ void fill_timesteps_kernel(float* __restrict__ dst, const TimestepChunk c, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = c.v[i];
}

}  // namespace ezb

namespace ezb {
// Direct 1-D convolution for the (tiny) ControlNet stem (controlnet.py:29-36,65-84): one thread per output element.
// in [B,Cin,Tin] (channel cin_real..Cin-1 read as zero: the eval-time all-zero mask channel), out [B,Cout,Tout] or
// transposed [B,Tout,Cout].
__global__ void conv1d_direct_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int B,
                                     int Cin, int cin_real, int Tin, int Cout, int Tout, int K, int stride, int pad, int act, int transposed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * Cout * Tout) return;
  int b, co, t;
  if (transposed) { co = i % Cout; t = (i / Cout) % Tout; b = i / ((size_t)Cout * Tout); }
  else { t = i % Tout; co = (i / Tout) % Cout; b = i / ((size_t)Cout * Tout); }
  float acc = bias[co];
  for (int ci = 0; ci < cin_real; ++ci)
    for (int k = 0; k < K; ++k) {
      const int ti = t * stride + k - pad;
      if (ti >= 0 && ti < Tin) acc = fmaf(w[((size_t)co * Cin + ci) * K + k], in[((size_t)b * cin_real + ci) * Tin + ti], acc);
    }
  if (act == 1) acc = silu(acc);
  out[i] = acc;
}
}  // namespace ezb

namespace ezb {
// ---------------------------------------------------------------------------------------------------------------------------------
// EnergyExtractor (src/models/conditions/energy.py:19-56): per frame f (hop `hop`, window `win`, reflect padding (win-hop)/2 both sides)
//   e_f = mean_{j<win} a[reflect(f*hop + j - pad)]^2 ; g_f = 10*log10(max(e_f, 10^(min_db/10)))
//   norm: g_f = (g_f - min_db) / (max_f g_f - min_db + 1e-8) ; quantize_levels q>0: round(g*(q-1))/(q-1).
// One CTA per clip: the frame energies stay in shared memory between the reduction over frames and the normalisation.
// Memory-bound (each sample is read win/hop = 8 times, all but the first from L1/L2); runs once per generate call.
__global__ void __launch_bounds__(1024) energy_kernel(const float* __restrict__ audio, float* __restrict__ out, int T, int n_frames, int hop,
                                                      int win, float min_db, float floor_e, int norm, int qlevels) {
  extern __shared__ float e_db[];  // n_frames
  __shared__ float red[32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* a = audio + (size_t)b * T;
  const int pad = (win - hop) / 2;
  float wmax = -INFINITY;
  for (int f = warp; f < n_frames; f += nw) {
    float acc = 0.f;
    const int base = f * hop - pad;
    for (int j = lane; j < win; j += 32) {
      int i = base + j;
      if (i < 0) i = -i;                      // F.pad(mode='reflect'): no edge repeat
      if (i >= T) i = 2 * (T - 1) - i;
      const float v = __ldg(a + i);
      acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    const float g = 10.f * log10f(fmaxf(acc / (float)win, floor_e));
    if (lane == 0) e_db[f] = g;
    wmax = fmaxf(wmax, g);
  }
  if (lane == 0) red[warp] = wmax;
  __syncthreads();
  float mx = red[0];
  for (int i = 1; i < nw; ++i) mx = fmaxf(mx, red[i]);
  for (int f = threadIdx.x; f < n_frames; f += blockDim.x) {
    float g = e_db[f];
    if (norm) g = (g - min_db) / (mx - min_db + 1e-8f);
    if (qlevels > 1) g = rintf(g * (float)(qlevels - 1)) / (float)(qlevels - 1);
    out[(size_t)b * n_frames + f] = g;
  }
}
}  // namespace ezb

namespace ezb {
// ---------------------------------------------------------------------------------------------------------------------------------
// Waveform pre / post-processing around the path, device-side so that clips never bounce through the host:
//   wave_prepare : gt / (max|gt| + 1e-9), optional noise gate |x| <= thr -> 0, pad / crop to T_out   (api/ezaudio.py:147,
//                  api/controlnet.py:119-133).  One CTA per clip: block max-abs reduction, then a scaled copy.
//   wave_splice  : output_audio[start : start + n] = pred[:n]                                         (api/ezaudio.py:198-203)
//   wave_to_pcm16: float -> 16-bit PCM with saturation (soundfile.write's default WAV subtype)         (t2a_demo.py:13,20)
// All three are HBM-bound copies: 4-8 bytes per sample.
__global__ void __launch_bounds__(1024) wave_prepare_kernel(const float* __restrict__ in, float* __restrict__ out, int T_in, int T_out, float eps,
                                                            float gate, int normalize) {
  __shared__ float red[32];
  __shared__ float inv_s;
  const int b = blockIdx.x;
  const float* a = in + (size_t)b * T_in;
  float inv = 1.f;
  if (normalize) {
    float m = 0.f;
    for (int i = threadIdx.x; i < T_in; i += blockDim.x) m = fmaxf(m, fabsf(a[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float mx = red[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
      inv_s = mx + eps;
    }
    __syncthreads();
    inv = inv_s;
  }
  float* o = out + (size_t)b * T_out;
  for (int i = threadIdx.x; i < T_out; i += blockDim.x) {
    float v = 0.f;
    if (i < T_in) {
      v = normalize ? a[i] / inv : a[i];   // a true division like numpy's (not a reciprocal multiply): bit-identical to the reference's float32 result
      if (gate > 0.f && fabsf(v) <= gate) v = 0.f;
    }
    o[i] = v;
  }
}
__global__ void wave_splice_kernel(float* __restrict__ dst, const float* __restrict__ src, long long start, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[start + i] = src[i];
}
__global__ void wave_to_pcm16_kernel(const float* __restrict__ in, int16_t* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = fminf(fmaxf(in[i] * 32768.0f, -32768.0f), 32767.0f);
  out[i] = (int16_t)__float2int_rn(v);
}
}  // namespace ezb
