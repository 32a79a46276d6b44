// tcgen05 flash attention for sm_100a: O = softmax(Q K^T / sqrt(dh) [+ key mask]) V.
// Replaces F.scaled_dot_product_attention in src/models/utils/attention.py:107-110 (self: mask None; cross: bool key mask built
// by attention.py:30-37).
//
// Persistent, warp-specialised, software-pipelined: one CTA per SM walks work items (b*H + h, 128-query tile); every item
// is a sequence of "units" (one 128-key block each).  Units are numbered across items so all rings keep rolling:
//   warp 9  (TMA)     : Q tile per item (2-deep ring), K / V^T tiles per unit (2-deep ring).
//   warp 8  (MMA)     : S_u = Q K_u^T (tcgen05.mma M128 N128) into TMEM S[u%2] is issued BEFORE waiting for unit u-1's
//                       probabilities, then O_{u-1} = P_{u-1} V_{u-1} (M128 N=DVP) into TMEM O[(u-1)%2]: the tensor pipe
//                       works one unit ahead of the softmax warps.
//   warps 0-7 (softmax): thread (w, lane) owns query row 32*(w%4)+lane and key columns 64*(w/4)..+63 of the S tile:
//                       tcgen05.ld, running max / exp2 / sum in fp32, P (bf16) written to smem in the 128B-swizzled K-major
//                       layout, O_{u-1} folded into the fp32 register accumulator (running-max correction) while the tensor
//                       pipe already computes S_{u+1}.
// Layouts (produced by the QKV GEMM epilogue / qk_prep_kernel): Q,K [B*H, L, DHP] bf16 (DHP = dh rounded up to 64, zero
// padded); V^T [B*H, DVP, Lpad] bf16 (DVP = dh rounded up to 16).  Output [B, Lq, H*dh] bf16 token-major.
#pragma once
#include "gemm.cuh"
#include "host.cuh"

namespace ezb {

constexpr int AT_BQ = 128, AT_BK = 128;
constexpr int AT_SOFTMAX_THREADS = 256;
constexpr int AT_THREADS = AT_SOFTMAX_THREADS + 64;

struct AttnParams {
  const uint8_t* key_mask;  // [B, Lk] or null
  __nv_bfloat16* out;       // [B, Lq, H*dh]
  int H, Lq, Lk, dh, dvp;
  int n_qt, n_items;        // query tiles per (b,h); total work items
  float scale_log2;         // (1/sqrt(dh)) * log2(e)
};

template <int KH>
struct AttnSmem {
  static constexpr int Q_BYTES = KH * 16384;
  static constexpr int K_BYTES = KH * 16384;
  static constexpr int P_BYTES = 2 * 16384;
  static __host__ __device__ constexpr int v_bytes(int dvp) { return 2 * dvp * 128; }
  static __host__ __device__ constexpr int total(int dvp) {
    return 1024 + 2 * Q_BYTES + 2 * K_BYTES + 2 * v_bytes(dvp) + P_BYTES + 2 * 2 * 128 * 4 + 16 * 8;
  }
};

template <int KH>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using SM = AttnSmem<KH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int VB = SM::v_bytes(p.dvp);
  uint8_t* sQ = smem;                      // [2][Q_BYTES]
  uint8_t* sK = sQ + 2 * SM::Q_BYTES;      // [2][K_BYTES]
  uint8_t* sV = sK + 2 * SM::K_BYTES;      // [2][VB]
  uint8_t* sP = sV + 2 * VB;               // [P_BYTES]
  float* sx = reinterpret_cast<float*>(sP + SM::P_BYTES);  // [2][2][128] max / sum exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(sx + 2 * 2 * 128);
  uint64_t *q_full = bars, *q_empty = bars + 2, *kv_full = bars + 4, *kv_empty = bars + 6, *s_full = bars + 8, *o_full = bars + 10, *p_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.Lk + AT_BK - 1) / AT_BK;
  const int my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_units = my_items * n_kv;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
        mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
        mbar_init(&s_full[i], 1); mbar_init(&o_full[i], 1);
      }
      mbar_init(p_full, AT_SOFTMAX_THREADS);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;  // S[0] @0, S[1] @128, O[0] @256, O[1] @384

  if (warp == 9) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int it = 0, u = 0; it < my_items; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int bh = item / p.n_qt, q0 = (item - bh * p.n_qt) * AT_BQ;
        const int qb = it & 1;
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qb], SM::Q_BYTES);
        for (int kh = 0; kh < KH; ++kh) tma_load_3d(sQ + qb * SM::Q_BYTES + kh * 16384, &tmQ, &q_full[qb], kh * 64, q0, bh);
        for (int j = 0; j < n_kv; ++j, ++u) {
          const int s = u & 1;
          mbar_wait(&kv_empty[s], ((u >> 1) & 1) ^ 1);
          mbar_expect_tx(&kv_full[s], SM::K_BYTES + VB);
          for (int kh = 0; kh < KH; ++kh) tma_load_3d(sK + s * SM::K_BYTES + kh * 16384, &tmK, &kv_full[s], kh * 64, j * AT_BK, bh);
          for (int hh = 0; hh < 2; ++hh) tma_load_3d(sV + s * VB + hh * (VB / 2), &tmV, &kv_full[s], j * AT_BK + hh * 64, 0, bh);
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_BQ, AT_BK), idesc_o = umma_idesc_bf16(AT_BQ, p.dvp);
      auto issue_pv = [&](int u) {  // O[u%2] = P_u V_u ; frees kv stage u%2
        const int s = u & 1;
        mbar_wait(p_full, u & 1);
        tc_fence_after();
        for (int hh = 0; hh < 2; ++hh) {
          const uint64_t pd = umma_desc_sw128(smem_u32(sP + hh * 16384)), vd = umma_desc_sw128(smem_u32(sV + s * VB + hh * (VB / 2)));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem0 + 256 + s * 128, pd + 2 * k, vd + 2 * k, idesc_o, (hh | k) != 0);
        }
        umma_commit(&o_full[s]);
        umma_commit(&kv_empty[s]);
      };
      for (int it = 0, u = 0; it < my_items; ++it) {
        const int qb = it & 1;
        mbar_wait(&q_full[qb], (it >> 1) & 1);
        for (int j = 0; j < n_kv; ++j, ++u) {
          const int s = u & 1;
          mbar_wait(&kv_full[s], (u >> 1) & 1);
          tc_fence_after();
          for (int kh = 0; kh < KH; ++kh) {
            const uint64_t qd = umma_desc_sw128(smem_u32(sQ + qb * SM::Q_BYTES + kh * 16384));
            const uint64_t kd = umma_desc_sw128(smem_u32(sK + s * SM::K_BYTES + kh * 16384));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem0 + s * 128, qd + 2 * k, kd + 2 * k, idesc_s, (kh | k) != 0);
          }
          umma_commit(&s_full[s]);
          if (j == n_kv - 1) umma_commit(&q_empty[qb]);  // last S of the item: Q buffer may be refilled
          if (u > 0) issue_pv(u - 1);
        }
      }
      if (n_units > 0) issue_pv(n_units - 1);
    }
  } else {
    // ------------------------------------------------ softmax / accumulate
    const int lg = warp & 3, ch = warp >> 2;       // TMEM lane group, key-column half
    const int r = lg * 32 + lane;                  // query row within the tile
    const uint32_t t_row = static_cast<uint32_t>(lg * 32) << 16;
    const int ocn = p.dvp / 2;                     // 32 or 40 O columns per thread
    const int oc0 = ch * ocn;
    float m_run = -INFINITY, l_run = 0.f;
    float o[40];
#pragma unroll
    for (int i = 0; i < 40; ++i) o[i] = 0.f;

    auto fold_o = [&](int u) {  // o += O_u (TMEM O[u%2]) once P_u V_u has completed
      mbar_wait(&o_full[u & 1], (u >> 1) & 1);
      tc_fence_after();
      const uint32_t ta = tmem0 + 256 + (u & 1) * 128 + t_row + oc0;
      uint32_t orr[32], t8[8];
      tmem_ld_32x32(ta, orr);
      if (ocn == 40) tmem_ld_32x8(ta + 32, t8);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(orr[i]);
      if (ocn == 40) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[32 + i] += __uint_as_float(t8[i]);
      }
      tc_fence_before();
    };
    auto finish_item = [&](int item) {  // combine the two column halves' partial sums, normalise, store
      named_bar_sync(1, AT_SOFTMAX_THREADS);
      float* xl = sx;
      xl[ch * 128 + r] = l_run;
      named_bar_sync(1, AT_SOFTMAX_THREADS);
      const float inv = 1.f / (l_run + xl[(ch ^ 1) * 128 + r]);
      const int bh = item / p.n_qt, q0 = (item - bh * p.n_qt) * AT_BQ;
      const int b = bh / p.H, h = bh - b * p.H;
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
      named_bar_sync(1, AT_SOFTMAX_THREADS);  // xl is reused by the next unit's max exchange
    };

    for (int it = 0, u = 0; it < my_items; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int bh = item / p.n_qt, b = bh / p.H;
      const uint8_t* km = p.key_mask ? p.key_mask + (size_t)b * p.Lk : nullptr;
      for (int j = 0; j < n_kv; ++j, ++u) {
        // ---- 1. scores of this unit
        mbar_wait(&s_full[u & 1], (u >> 1) & 1);
        tc_fence_after();
        uint32_t sr[64];
        const uint32_t ts = tmem0 + (u & 1) * 128 + t_row + ch * 64;
        tmem_ld_32x32(ts, sr);
        tmem_ld_32x32(ts + 32, sr + 32);
        tmem_ld_wait();
        tc_fence_before();
        const int kbase = j * AT_BK + ch * 64;
        const bool full = (km == nullptr) && (kbase + 64 <= p.Lk);
        float mx = -INFINITY;
        if (full) {
#pragma unroll
          for (int c = 0; c < 64; ++c) mx = fmaxf(mx, __uint_as_float(sr[c]));
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            const int kidx = kbase + c;
            const bool ok = kidx < p.Lk && (km == nullptr || km[kidx] != 0);
            const float v = ok ? __uint_as_float(sr[c]) : -INFINITY;
            sr[c] = __float_as_uint(v);
            mx = fmaxf(mx, v);
          }
        }
        float* xm = sx + (u & 1) * 256;
        xm[ch * 128 + r] = mx;
        named_bar_sync(1, AT_SOFTMAX_THREADS);
        const float m_blk = fmaxf(mx, xm[(ch ^ 1) * 128 + r]);
        // ---- 2. previous unit's P V product: fold into the accumulator (and close the previous item)
        if (u > 0) {
          fold_o(u - 1);
          if (j == 0) {  // the previous unit was the last one of the previous item
            finish_item(item - (int)gridDim.x);
#pragma unroll
            for (int i = 0; i < 40; ++i) o[i] = 0.f;
            m_run = -INFINITY;
            l_run = 0.f;
          }
        }
        // ---- 3. running max, probabilities (P_u may be written: P_{u-1} V_{u-1} has completed)
        const float m_new = fmaxf(m_run, m_blk);
        const float corr = (m_run == -INFINITY) ? 0.f : ex2_approx((m_run - m_new) * p.scale_log2);
        const float mb = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;  // fully masked so far: p = exp2(-inf) = 0
        float sum = 0.f;
        uint8_t* prow = sP + ch * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          float pv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            pv[e] = ex2_approx(fmaf(__uint_as_float(sr[c8 * 8 + e]), p.scale_log2, -mb));
            sum += pv[e];
          }
          uint4 pk = make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]), pack_bf16(pv[6], pv[7]));
          *reinterpret_cast<uint4*>(prow + ((c8 ^ (r & 7)) << 4)) = pk;
        }
        l_run = l_run * corr + sum;
        m_run = m_new;
#pragma unroll
        for (int i = 0; i < 40; ++i) o[i] *= corr;
        fence_proxy_async_smem();
        mbar_arrive(p_full);
      }
    }
    if (n_units > 0) {
      fold_o(n_units - 1);
      finish_item(blockIdx.x + (my_items - 1) * gridDim.x);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem0);
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
  p.n_qt = (Lq + AT_BQ - 1) / AT_BQ;
  p.n_items = p.n_qt * B * H;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int grid = p.n_items < dev.num_sms ? p.n_items : dev.num_sms;
  ++launch_counter();
  if (dhp == 64) {
    const int smem = AttnSmem<1>::total(dvp);
    static bool set = false;
    if (!set) { EZB_CUDA(cudaFuncSetAttribute(attn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<1>::total(80))); set = true; }
    attn_tc_kernel<1><<<grid, AT_THREADS, smem, st>>>(*tq, *tk, *tv, p);
  } else {
    const int smem = AttnSmem<2>::total(dvp);
    static bool set = false;
    if (!set) { EZB_CUDA(cudaFuncSetAttribute(attn_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<2>::total(80))); set = true; }
    attn_tc_kernel<2><<<grid, AT_THREADS, smem, st>>>(*tq, *tk, *tv, p);
  }
  EZB_CUDA(cudaGetLastError());
  return EZB_OK;
}

}  // namespace ezb
