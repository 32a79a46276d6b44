// fp32 CUDA-core attention: softmax(q k^T * scale [+ bias] [+ key mask]) v   (attention.py:107-110, mask attention.py:30-37;
// with scale 1 and a relative-position bias: T5Attention).
// Used by the bf16x3 parity mode (fp32-grade numerics) and as the on-device comparator of the tcgen05 kernel.
// q,k,v fp32 [B,H,L,dh]; out bf16 [B, Lq, H*dh] token-major (A operand of the output projection).
#pragma once
#include "elementwise.cuh"

namespace ezb {

constexpr int SA_TK = 64;   // keys per smem tile
constexpr int SA_QW = 8;    // queries per warp
constexpr int SA_WARPS = 4;

__global__ void __launch_bounds__(SA_WARPS * 32) attn_simt_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                                  const uint8_t* __restrict__ key_mask, __nv_bfloat16* __restrict__ out, int H, int Lq,
                                                                  int Lk, int dh, float scale, int kmul,
                                                                  const float* __restrict__ bias = nullptr /* [H, Lq, Lk] added to the scores (T5) */) {
  extern __shared__ float sm[];
  const int ldk = dh | 1;  // odd pitch: conflict-free row-per-lane reads
  float* sK = sm;                       // [SA_TK][ldk]
  float* sV = sK + SA_TK * ldk;         // [SA_TK][dh]
  float* sQ = sV + SA_TK * dh;          // [SA_WARPS*SA_QW][dh]
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (SA_WARPS * SA_QW);
  const float* qb = q + (size_t)bh * Lq * dh;
  const float* kb = k + (size_t)bh * Lk * dh;
  const float* vb = v + (size_t)bh * Lk * dh;
  for (int i = threadIdx.x; i < SA_WARPS * SA_QW * dh; i += blockDim.x) {
    const int r = i / dh, d = i - r * dh;
    sQ[i] = (q0 + r < Lq) ? qb[(size_t)(q0 + r) * dh + d] * scale : 0.f;
  }
  float m[SA_QW], l[SA_QW], acc[SA_QW][3];
#pragma unroll
  for (int i = 0; i < SA_QW; ++i) { m[i] = -INFINITY; l[i] = 0.f; acc[i][0] = acc[i][1] = acc[i][2] = 0.f; }
  for (int k0 = 0; k0 < Lk; k0 += SA_TK) {
    __syncthreads();
    for (int i = threadIdx.x; i < SA_TK * dh; i += blockDim.x) {
      const int r = i / dh, d = i - r * dh;
      const bool ok = k0 + r < Lk;
      sK[r * ldk + d] = ok ? kb[(size_t)(k0 + r) * dh + d] : 0.f;
      sV[r * dh + d] = ok ? vb[(size_t)(k0 + r) * dh + d] : 0.f;
    }
    __syncthreads();
    bool ok0 = k0 + lane < Lk, ok1 = k0 + lane + 32 < Lk;
    if (key_mask) {
      ok0 = ok0 && key_mask[(size_t)b * Lk + k0 + lane];
      ok1 = ok1 && key_mask[(size_t)b * Lk + k0 + lane + 32];
    }
    // scores of this warp's SA_QW queries against the lane's two keys: the K values are read once per d and reused by all queries
    // (the first version re-read them per query: 3 shared loads per 2 FMAs, LDS-bound at 199 us per T5 layer)
    float s0[SA_QW], s1[SA_QW];
#pragma unroll
    for (int qi = 0; qi < SA_QW; ++qi) { s0[qi] = 0.f; s1[qi] = 0.f; }
    const float* qw = sQ + (warp * SA_QW) * dh;
    const float* kr0 = sK + lane * ldk;
    const float* kr1 = sK + (lane + 32) * ldk;
#pragma unroll 4
    for (int d = 0; d < dh; ++d) {
      const float k0v = kr0[d], k1v = kr1[d];
#pragma unroll
      for (int qi = 0; qi < SA_QW; ++qi) {
        const float qv = qw[qi * dh + d];
        s0[qi] = fmaf(qv, k0v, s0[qi]);
        s1[qi] = fmaf(qv, k1v, s1[qi]);
      }
    }
    float p0[SA_QW], p1[SA_QW];
#pragma unroll
    for (int qi = 0; qi < SA_QW; ++qi) {
      float a0 = s0[qi], a1 = s1[qi];
      if (bias != nullptr) {
        const int qrow = q0 + warp * SA_QW + qi;
        if (qrow < Lq) {
          const float* br = bias + ((size_t)h * Lq + qrow) * Lk + k0;
          if (k0 + lane < Lk) a0 += br[lane];
          if (k0 + lane + 32 < Lk) a1 += br[lane + 32];
        }
      }
      a0 = ok0 ? a0 : -INFINITY;
      a1 = ok1 ? a1 : -INFINITY;
      const float mn = fmaxf(m[qi], warp_max(fmaxf(a0, a1)));
      const float corr = (mn == -INFINITY) ? 1.f : expf(m[qi] - mn);
      p0[qi] = (mn == -INFINITY) ? 0.f : expf(a0 - mn);
      p1[qi] = (mn == -INFINITY) ? 0.f : expf(a1 - mn);
      l[qi] = l[qi] * corr + warp_sum(p0[qi] + p1[qi]);
      m[qi] = mn;
      acc[qi][0] *= corr; acc[qi][1] *= corr; acc[qi][2] *= corr;
    }
    // P V: a V row is read once and reused by all queries
    const bool d1 = lane + 32 < dh, d2 = lane + 64 < dh;
#pragma unroll 2
    for (int j = 0; j < SA_TK; ++j) {
      const float* vr = sV + j * dh;
      const float v0 = vr[lane], v1 = d1 ? vr[lane + 32] : 0.f, v2 = d2 ? vr[lane + 64] : 0.f;
#pragma unroll
      for (int qi = 0; qi < SA_QW; ++qi) {
        const float pj = __shfl_sync(0xffffffffu, j < 32 ? p0[qi] : p1[qi], j & 31);
        acc[qi][0] = fmaf(pj, v0, acc[qi][0]);
        acc[qi][1] = fmaf(pj, v1, acc[qi][1]);
        acc[qi][2] = fmaf(pj, v2, acc[qi][2]);
      }
    }
  }
  const int D = H * dh;
#pragma unroll
  for (int qi = 0; qi < SA_QW; ++qi) {
    const int qrow = q0 + warp * SA_QW + qi;
    if (qrow >= Lq) continue;
    const float inv = 1.f / l[qi];
    __nv_bfloat16* o = out + ((size_t)b * Lq + qrow) * kmul * D;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (lane + 32 * i < dh) store_act(o, h * dh + lane + 32 * i, D, kmul, acc[qi][i] * inv);
  }
}

inline size_t attn_simt_smem(int dh) { return sizeof(float) * (SA_TK * (dh | 1) + SA_TK * dh + SA_WARPS * SA_QW * dh); }

}  // namespace ezb
