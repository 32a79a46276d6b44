// tcgen05 flash attention for sm_100a: O = softmax(Q K^T / sqrt(dh) [+ key mask]) V, one CTA per (128-query tile, b, h).
// Replaces F.scaled_dot_product_attention in src/models/utils/attention.py:107-110 (self: mask None; cross: bool key mask built
// by attention.py:30-37).
//
//   warp 8       : TMA producer + MMA issuer (one elected lane): Q once, K / V^T tiles of 128 keys double-buffered;
//                  S = Q K^T (tcgen05.mma M128 N128, K = DHP) into TMEM, then O_blk = P V (M128, N = DVP, K = 128 keys).
//   warps 0..7   : softmax.  Thread (w, lane) owns query row 32*(w%4)+lane and key columns 64*(w/4)..+63 of the S tile:
//                  tcgen05.ld, online max / exp2 / sum in fp32, P (bf16) written to smem in the 128B-swizzled K-major
//                  layout the MMA reads, then O_blk read back from TMEM and folded into the fp32 register accumulator
//                  with the running-max correction.
// Layouts (produced by qk_prep_kernel): Q,K [B*H, L, DHP] bf16 (DHP = dh rounded up to 64, zero padded);
// V^T [B*H, DVP, Lpad] bf16 (DVP = dh rounded up to 16).  Output [B, Lq, H*dh] bf16 token-major.
#pragma once
#include "host.cuh"

namespace ezb {

constexpr int AT_BQ = 128, AT_BK = 128;
constexpr int AT_SOFTMAX_THREADS = 256;
constexpr int AT_THREADS = AT_SOFTMAX_THREADS + 32;

struct AttnParams {
  const uint8_t* key_mask;  // [B, Lk] or null
  __nv_bfloat16* out;       // [B, Lq, H*dh]
  int H, Lq, Lk, dh, dvp;
  float scale_log2;         // (1/sqrt(dh)) * log2(e)
};

template <int KH>
struct AttnSmem {
  static constexpr int Q_BYTES = KH * 16384;
  static constexpr int K_BYTES = KH * 16384;
  static constexpr int P_BYTES = 2 * 16384;
  static __host__ __device__ constexpr int v_bytes(int dvp) { return 2 * dvp * 128; }
  static __host__ __device__ constexpr int total(int dvp) { return 1024 + Q_BYTES + 2 * K_BYTES + 2 * v_bytes(dvp) + P_BYTES + 2 * 2 * 128 * 4 + 16 * 8; }
};

template <int KH>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using SM = AttnSmem<KH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int VB = SM::v_bytes(p.dvp);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SM::Q_BYTES;
  uint8_t* sV = sK + 2 * SM::K_BYTES;
  uint8_t* sP = sV + 2 * VB;
  float* sx = reinterpret_cast<float*>(sP + SM::P_BYTES);  // [2][2][128] max / sum exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(sx + 2 * 2 * 128);
  uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = bars + 3, *s_full = bars + 5, *p_full = bars + 6, *o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, q0 = blockIdx.x * AT_BQ;
  const int n_kv = (p.Lk + AT_BK - 1) / AT_BK;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
      mbar_init(s_full, 1);
      mbar_init(p_full, AT_SOFTMAX_THREADS);
      mbar_init(o_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_S = *tmem_slot, tmem_O = tmem_S + 128;

  if (warp == 8) {
    if (lane == 0) {
      auto load_kv = [&](int j) {
        const int s = j & 1;
        mbar_expect_tx(&kv_full[s], SM::K_BYTES + VB);
        for (int kh = 0; kh < KH; ++kh) tma_load_3d(sK + s * SM::K_BYTES + kh * 16384, &tmK, &kv_full[s], kh * 64, j * AT_BK, bh);
        for (int hh = 0; hh < 2; ++hh) tma_load_3d(sV + s * VB + hh * (VB / 2), &tmV, &kv_full[s], j * AT_BK + hh * 64, 0, bh);
      };
      mbar_expect_tx(q_full, SM::Q_BYTES);
      for (int kh = 0; kh < KH; ++kh) tma_load_3d(sQ + kh * 16384, &tmQ, q_full, kh * 64, q0, bh);
      load_kv(0);
      if (n_kv > 1) load_kv(1);
      const uint32_t idesc_s = umma_idesc_bf16(AT_BQ, AT_BK), idesc_o = umma_idesc_bf16(AT_BQ, p.dvp);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        for (int kh = 0; kh < KH; ++kh) {
          const uint64_t qd = umma_desc_sw128(smem_u32(sQ + kh * 16384)), kd = umma_desc_sw128(smem_u32(sK + s * SM::K_BYTES + kh * 16384));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, qd + 2 * k, kd + 2 * k, idesc_s, (kh | k) != 0);
        }
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        for (int hh = 0; hh < 2; ++hh) {
          const uint64_t pd = umma_desc_sw128(smem_u32(sP + hh * 16384)), vd = umma_desc_sw128(smem_u32(sV + s * VB + hh * (VB / 2)));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_O, pd + 2 * k, vd + 2 * k, idesc_o, (hh | k) != 0);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
        if (j + 2 < n_kv) {
          mbar_wait(&kv_empty[s], (j >> 1) & 1);
          load_kv(j + 2);
        }
      }
    }
  } else {
    // ------------------------------------------------ softmax / accumulate
    const int lg = warp & 3, ch = warp >> 2;       // TMEM lane group, key-column half
    const int r = lg * 32 + lane;                  // query row within the tile
    const int b = bh / p.H, h = bh - b * p.H;
    const uint32_t t_row = static_cast<uint32_t>(lg * 32) << 16;
    const int oc0 = ch * (p.dvp / 2);              // first O column owned by this thread
    const int ocn = p.dvp / 2;                     // 32 or 40
    float m_run = -INFINITY, l_run = 0.f;
    float o[40];
#pragma unroll
    for (int i = 0; i < 40; ++i) o[i] = 0.f;
    const uint8_t* km = p.key_mask ? p.key_mask + (size_t)b * p.Lk : nullptr;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld_32x32(tmem_S + t_row + ch * 64, sr);
      tmem_ld_32x32(tmem_S + t_row + ch * 64 + 32, sr + 32);
      tmem_ld_wait();
      const int kbase = j * AT_BK + ch * 64;
      const bool full = (km == nullptr) && (kbase + 64 <= p.Lk);
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float v = __uint_as_float(sr[c]);
        if (!full) {
          const int kidx = kbase + c;
          const bool ok = kidx < p.Lk && (km == nullptr || km[kidx] != 0);
          v = ok ? v : -INFINITY;
          sr[c] = __float_as_uint(v);
        }
        mx = fmaxf(mx, v);
      }
      float* xm = sx + (j & 1) * 256;
      xm[ch * 128 + r] = mx;
      named_bar_sync(1, AT_SOFTMAX_THREADS);
      const float m_blk = fmaxf(mx, xm[(ch ^ 1) * 128 + r]);
      const float m_new = fmaxf(m_run, m_blk);
      const float corr = (m_run == -INFINITY) ? 0.f : exp2f((m_run - m_new) * p.scale_log2);
      const float mb = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;  // fully masked so far: p = exp2(-inf) = 0
      float sum = 0.f;
      uint8_t* prow = sP + ch * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pv[e] = exp2f(__uint_as_float(sr[c8 * 8 + e]) * p.scale_log2 - mb);
          sum += pv[e];
        }
        uint4 pk = make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]), pack_bf16(pv[6], pv[7]));
        *reinterpret_cast<uint4*>(prow + ((c8 ^ (r & 7)) << 4)) = pk;
      }
      l_run = l_run * corr + sum;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      // fold the previous correction into the accumulator while the PV MMA runs
#pragma unroll
      for (int i = 0; i < 40; ++i) o[i] *= corr;
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      uint32_t orr[40];
      tmem_ld_32x32(tmem_O + t_row + oc0, orr);
      if (ocn == 40) {
        uint32_t t8[16];
        tmem_ld_32x16(tmem_O + t_row + oc0 + 32, t8);  // reads 16 columns; only the first 8 belong to this thread
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) orr[32 + i] = t8[i];
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) orr[32 + i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 40; ++i) o[i] += __uint_as_float(orr[i]);
      tc_fence_before();
    }
    // combine the two column halves' partial sums, normalise, store
    named_bar_sync(1, AT_SOFTMAX_THREADS);
    float* xl = sx;
    xl[ch * 128 + r] = l_run;
    named_bar_sync(1, AT_SOFTMAX_THREADS);
    const float inv = 1.f / (l_run + xl[(ch ^ 1) * 128 + r]);
    const int qrow = q0 + r;
    if (qrow < p.Lq) {
      __nv_bfloat16* orow = p.out + ((size_t)b * p.Lq + qrow) * (size_t)(p.H * p.dh) + h * p.dh;
#pragma unroll
      for (int v8 = 0; v8 < 5; ++v8) {
        const int c0 = oc0 + v8 * 8;
        if (v8 * 8 < ocn && c0 + 8 <= p.dh) {
          uint4 pk = make_uint4(pack_bf16(o[v8 * 8 + 0] * inv, o[v8 * 8 + 1] * inv), pack_bf16(o[v8 * 8 + 2] * inv, o[v8 * 8 + 3] * inv),
                                pack_bf16(o[v8 * 8 + 4] * inv, o[v8 * 8 + 5] * inv), pack_bf16(o[v8 * 8 + 6] * inv, o[v8 * 8 + 7] * inv));
          *reinterpret_cast<uint4*>(orow + c0) = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<256>(tmem_S);
}

inline int attention_tc(Device& dev, cudaStream_t st, const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, const uint8_t* key_mask,
                        __nv_bfloat16* out, int B, int H, int Lq, int Lk, int Lkpad, int dh, int dhp, int dvp, float scale) {
  if (dhp != 64 && dhp != 128) return fail(EZB_ERR_UNSUPPORTED, "attention_tc: dhp %d", dhp);
  if (dvp % 16 || dvp > 80 || dvp < 16 || (dvp / 2) % 8) return fail(EZB_ERR_UNSUPPORTED, "attention_tc: dvp %d", dvp);
  const CUtensorMap *tq, *tk, *tv;
  EZB_TRY(dev.tmaps.get3d(q, dhp, Lq, (uint64_t)B * H, dhp, (uint64_t)Lq * dhp, AT_BQ, &tq));
  EZB_TRY(dev.tmaps.get3d(k, dhp, Lk, (uint64_t)B * H, dhp, (uint64_t)Lk * dhp, AT_BK, &tk));
  EZB_TRY(dev.tmaps.get3d(vt, Lk, dvp, (uint64_t)B * H, Lkpad, (uint64_t)dvp * Lkpad, dvp, &tv));
  AttnParams p;
  p.key_mask = key_mask; p.out = out; p.H = H; p.Lq = Lq; p.Lk = Lk; p.dh = dh; p.dvp = dvp;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Lq + AT_BQ - 1) / AT_BQ, B * H);
  if (dhp == 64) {
    const int smem = AttnSmem<1>::total(dvp);
    static bool set = false;
    if (!set) { EZB_CUDA(cudaFuncSetAttribute(attn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<1>::total(80))); set = true; }
    ++launch_counter();
    attn_tc_kernel<1><<<grid, AT_THREADS, smem, st>>>(*tq, *tk, *tv, p);
  } else {
    const int smem = AttnSmem<2>::total(dvp);
    static bool set = false;
    if (!set) { EZB_CUDA(cudaFuncSetAttribute(attn_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<2>::total(80))); set = true; }
    ++launch_counter();
    attn_tc_kernel<2><<<grid, AT_THREADS, smem, st>>>(*tq, *tk, *tv, p);
  }
  EZB_CUDA(cudaGetLastError());
  return EZB_OK;
}

}  // namespace ezb
