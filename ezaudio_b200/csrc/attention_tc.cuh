// tcgen05 flash attention for sm_100a: O = softmax(Q K^T / sqrt(dh) [+ key mask]) V.
// Replaces F.scaled_dot_product_attention in src/models/utils/attention.py:107-110 (self: mask None; cross: bool key mask built
// by attention.py:30-37).
//
// Persistent, warp-specialised, software-pipelined: one CTA per SM walks work items (b*H + h, 128-query tile); every item
// is a sequence of "units" (one 128-key block each).  Units are numbered across items so all rings keep rolling:
//   warp 17 (TMA)     : Q tile per item (2-deep ring), K / V^T tiles per unit (2-deep ring).  For dh = 72 a Q / K tile is two
//                       boxes of the same rows: columns 0..63 (SWIZZLE_128B) and 64..79 (SWIZZLE_32B, one MMA K-step), so only
//                       80 of the 128 padded columns are moved and multiplied.
//   warp 16 (MMA)     : S_u = Q K_u^T (tcgen05.mma M128 N128) into TMEM S[u%2] is issued BEFORE waiting for unit u-1's
//                       probabilities, then O_{u-1} = P_{u-1} V_{u-1} (M128 N=DVP) into TMEM O[(u-1)%2]: the tensor pipe
//                       works one unit ahead of the softmax warps.
//   warps 0-15 (softmax): thread (w, lane) owns query row 32*(w%4)+lane and key columns 32*(w/4)..+31 of the S tile:
//                       tcgen05.ld, running max / exp2 / sum in fp32, P (bf16) written to smem in the 128B-swizzled K-major
//                       layout (double-buffered), and only AFTER publishing P_u is O_{u-1} folded into the fp32 register
//                       accumulator: P_{u-1} V_{u-1} runs on the tensor pipe during the whole softmax of unit u.
// Layouts (produced by the QKV GEMM epilogue / qk_prep_kernel): Q,K [B*H, L, DHP] bf16 (DHP = dh rounded up to 64, zero
// padded); V^T [B*H, DVP, Lpad] bf16 (DVP = dh rounded up to 16).  Output [B, Lq, H*dh] bf16 token-major.
#pragma once
#include "gemm.cuh"
#include "host.cuh"

namespace ezb {

constexpr int AT_BQ = 128, AT_BK = 128;
constexpr int AT_SOFTMAX_THREADS = 512;   // 16 warps: 4 per TMEM lane group, each owning 32 key columns of the S tile
constexpr int AT_THREADS = AT_SOFTMAX_THREADS + 64;

struct AttnParams {
  const uint8_t* key_mask;  // [B, Lk] or null
  __nv_bfloat16* out;       // [B, Lq, H*dh]
  int H, Lq, Lk, dh, dvp;
  int n_qt, n_items;        // query tiles per (b,h); total work items
  float scale_log2;         // (1/sqrt(dh)) * log2(e)
  unsigned long long* dbg;  // optional cycle counters of CTA 0: [0] softmax wait S, [1] softmax wait O, [2] softmax barrier, [3] softmax total,
                            // [4] mma wait kv, [5] mma wait P, [6] mma total, [7] tma wait
};

template <int DH>
struct AttnSmem {
  static constexpr int TAIL = DH > 64 ? 4096 : 0;          // columns 64..79: 128 rows x 32 B
  static constexpr int Q_BYTES = 16384 + TAIL;
  static constexpr int K_BYTES = 16384 + TAIL;
  static constexpr int P_BYTES = 2 * 16384;
  static __host__ __device__ constexpr int v_bytes(int dvp) { return 2 * dvp * 128; }
  static __host__ __device__ constexpr int total(int dvp) {
    return 1024 + 2 * Q_BYTES + 2 * K_BYTES + 2 * v_bytes(dvp) + 2 * P_BYTES + 2 * 4 * 128 * 4 + 16 * 8;
  }
};

template <int DH>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const __grid_constant__ CUtensorMap tmQt, const __grid_constant__ CUtensorMap tmKt, const AttnParams p) {
  using SM = AttnSmem<DH>;
  constexpr bool HAS_TAIL = DH > 64;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (a round trip through uintptr_t loses the address space and
  // turns every staging access into a generic LD.E / ST.E)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int VB = SM::v_bytes(p.dvp);
  uint8_t* sQ = smem;                      // [2][Q_BYTES]
  uint8_t* sK = sQ + 2 * SM::Q_BYTES;      // [2][K_BYTES]
  uint8_t* sV = sK + 2 * SM::K_BYTES;      // [2][VB]
  uint8_t* sP = sV + 2 * VB;               // [2][P_BYTES]
  float* sx = reinterpret_cast<float*>(sP + 2 * SM::P_BYTES);  // [2][4][128] max / sum exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(sx + 2 * 4 * 128);
  uint64_t *q_full = bars, *q_empty = bars + 2, *kv_full = bars + 4, *kv_empty = bars + 6, *s_full = bars + 8, *o_full = bars + 10, *p_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.Lk + AT_BK - 1) / AT_BK;
  const int my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_units = my_items * n_kv;

  if (warp == 16) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      if (HAS_TAIL) { tma_prefetch_desc(&tmQt); tma_prefetch_desc(&tmKt); }
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
  pdl_launch();
  pdl_wait();

  if (warp == 17) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int it = 0, u = 0; it < my_items; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int bh = item / p.n_qt, q0 = (item - bh * p.n_qt) * AT_BQ;
        const int qb = it & 1;
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qb], SM::Q_BYTES);
        tma_load_3d(sQ + qb * SM::Q_BYTES, &tmQ, &q_full[qb], 0, q0, bh);
        if (HAS_TAIL) tma_load_3d(sQ + qb * SM::Q_BYTES + 16384, &tmQt, &q_full[qb], 64, q0, bh);
        for (int j = 0; j < n_kv; ++j, ++u) {
          const int s = u & 1;
          mbar_wait(&kv_empty[s], ((u >> 1) & 1) ^ 1);
          mbar_expect_tx(&kv_full[s], SM::K_BYTES + VB);
          tma_load_3d(sK + s * SM::K_BYTES, &tmK, &kv_full[s], 0, j * AT_BK, bh);
          if (HAS_TAIL) tma_load_3d(sK + s * SM::K_BYTES + 16384, &tmKt, &kv_full[s], 64, j * AT_BK, bh);
          for (int hh = 0; hh < 2; ++hh) tma_load_3d(sV + s * VB + hh * (VB / 2), &tmV, &kv_full[s], j * AT_BK + hh * 64, 0, bh);
        }
      }
    }
  } else if (warp == 16) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_BQ, AT_BK), idesc_o = umma_idesc_bf16(AT_BQ, p.dvp);
      EZB_DBG(long long m_wkv = 0, m_wp = 0; const long long m_t0 = clock64();)
      auto issue_pv = [&](int u) {  // O[u%2] = P_u V_u ; frees kv stage u%2
        const int s = u & 1;
        EZB_DBG(const long long tq = clock64();)
        mbar_wait(p_full, u & 1);
        EZB_DBG(m_wp += clock64() - tq;)
        tc_fence_after();
        for (int hh = 0; hh < 2; ++hh) {
          const uint64_t pd = umma_desc_sw128(smem_u32(sP + s * SM::P_BYTES + hh * 16384)), vd = umma_desc_sw128(smem_u32(sV + s * VB + hh * (VB / 2)));
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
          EZB_DBG(const long long tq = clock64();)
          mbar_wait(&kv_full[s], (u >> 1) & 1);
          EZB_DBG(m_wkv += clock64() - tq;)
          tc_fence_after();
          {
            const uint64_t qd = umma_desc_sw128(smem_u32(sQ + qb * SM::Q_BYTES));
            const uint64_t kd = umma_desc_sw128(smem_u32(sK + s * SM::K_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem0 + s * 128, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
            if (HAS_TAIL)
              umma_bf16(tmem0 + s * 128, umma_desc_sw32(smem_u32(sQ + qb * SM::Q_BYTES + 16384)), umma_desc_sw32(smem_u32(sK + s * SM::K_BYTES + 16384)),
                        idesc_s, 1);
          }
          umma_commit(&s_full[s]);
          if (j == n_kv - 1) umma_commit(&q_empty[qb]);  // last S of the item: Q buffer may be refilled
          if (u > 0) issue_pv(u - 1);
        }
      }
      if (n_units > 0) issue_pv(n_units - 1);
      EZB_DBG(if (p.dbg != nullptr && blockIdx.x == 0) {
        atomicAdd(&p.dbg[4], (unsigned long long)m_wkv); atomicAdd(&p.dbg[5], (unsigned long long)m_wp); atomicAdd(&p.dbg[6], (unsigned long long)(clock64() - m_t0));
      })
    }
  } else {
    // ------------------------------------------------ softmax / accumulate
    const int lg = warp & 3, cq = warp >> 2;       // TMEM lane group, key-column quarter
    const int r = lg * 32 + lane;                  // query row within the tile
    const uint32_t t_row = static_cast<uint32_t>(lg * 32) << 16;
    const int ocn = p.dvp / 4;                     // 16 or 20 O columns per thread
    const int oc0 = cq * ocn;
    EZB_DBG(long long c_ws = 0, c_wo = 0, c_bar = 0; const long long c_t0 = clock64();)
    float m_run = -INFINITY, l_run = 0.f;
    float o[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) o[i] = 0.f;

    auto fold_o = [&](int u) {  // o += O_u (TMEM O[u%2]) once P_u V_u has completed
      EZB_DBG(const long long tq = clock64();)
      mbar_wait(&o_full[u & 1], (u >> 1) & 1);
      EZB_DBG(c_wo += clock64() - tq;)
      tc_fence_after();
      const uint32_t ta = tmem0 + 256 + (u & 1) * 128 + t_row + oc0;
      uint32_t orr[16], t4[4];
      tmem_ld_32x16(ta, orr);
      if (ocn == 20) tmem_ld_32x4(ta + 16, t4);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] += __uint_as_float(orr[i]);
      if (ocn == 20) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[16 + i] += __uint_as_float(t4[i]);
      }
      tc_fence_before();
    };
    auto finish_item = [&](int item, float l_part) {  // combine the four column quarters' partial sums, normalise, store
      named_bar_sync(1, AT_SOFTMAX_THREADS);
      float* xl = sx;
      xl[cq * 128 + r] = l_part;
      named_bar_sync(1, AT_SOFTMAX_THREADS);
      const float inv = 1.f / ((xl[r] + xl[128 + r]) + (xl[256 + r] + xl[384 + r]));
      const int bh = item / p.n_qt, q0 = (item - bh * p.n_qt) * AT_BQ;
      const int b = bh / p.H, h = bh - b * p.H;
      const int qrow = q0 + r;
      if (qrow < p.Lq) {
        __nv_bfloat16* orow = p.out + ((size_t)b * p.Lq + qrow) * (size_t)(p.H * p.dh) + h * p.dh;
#pragma unroll
        for (int v4 = 0; v4 < 5; ++v4) {
          const int c0 = oc0 + v4 * 4;
          if (v4 * 4 < ocn && c0 + 4 <= p.dh)
            *reinterpret_cast<uint2*>(orow + c0) = make_uint2(pack_bf16(o[v4 * 4 + 0] * inv, o[v4 * 4 + 1] * inv), pack_bf16(o[v4 * 4 + 2] * inv, o[v4 * 4 + 3] * inv));
        }
      }
      named_bar_sync(1, AT_SOFTMAX_THREADS);  // xl is reused by the next unit's max exchange
    };

    float c_pend = 0.f;   // correction to apply to `o` when the pending unit's P V product is folded
    for (int it = 0, u = 0; it < my_items; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int bh = item / p.n_qt, b = bh / p.H;
      const uint8_t* km = p.key_mask ? p.key_mask + (size_t)b * p.Lk : nullptr;
      for (int j = 0; j < n_kv; ++j, ++u) {
        // ---- 1. scores of this unit (already in registers: loaded at the end of the previous unit, see step 3)
        EZB_DBG(long long tq = clock64();)
        mbar_wait(&s_full[u & 1], (u >> 1) & 1);
        EZB_DBG(c_ws += clock64() - tq;)
        tc_fence_after();
        uint32_t sr[32];
        tmem_ld_32x32(tmem0 + (u & 1) * 128 + t_row + cq * 32, sr);
        tmem_ld_wait();
        tc_fence_before();
        const int kbase = j * AT_BK + cq * 32;
        const bool full = (km == nullptr) && (kbase + 32 <= p.Lk);
        if (!full) {  // branch-free masking: one validity bit per key column of this thread
          const int rem = p.Lk - kbase;  // keys of this quarter that exist
          uint32_t valid = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
          if (km != nullptr) {
            uint32_t mk = 0u;
#pragma unroll
            for (int c = 0; c < 32; ++c) mk |= (uint32_t)(km[min(kbase + c, p.Lk - 1)] != 0) << c;
            valid &= mk;
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) sr[c] = ((valid >> c) & 1u) ? sr[c] : 0xff800000u;  // -inf
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
#pragma unroll
          for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(sr[c + e]));
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        float* xm = sx + (u & 1) * 512;
        xm[cq * 128 + r] = mx;
        EZB_DBG(tq = clock64();)
        named_bar_sync(1, AT_SOFTMAX_THREADS);
        EZB_DBG(c_bar += clock64() - tq;)
        const float m_blk = fmaxf(fmaxf(xm[r], xm[128 + r]), fmaxf(xm[256 + r], xm[384 + r]));
        // ---- 2. running max and probabilities of this unit (a new item starts from scratch)
        const bool first = j == 0;
        const float l_fin = l_run;                        // previous item's partial sum (used below if `first`)
        const float m_old = first ? -INFINITY : m_run, l_old = first ? 0.f : l_run;
        const float m_new = fmaxf(m_old, m_blk);
        const float corr = (m_old == -INFINITY) ? 0.f : ex2_approx((m_old - m_new) * p.scale_log2);
        const float mb = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;  // fully masked so far: p = exp2(-inf) = 0
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
        uint8_t* prow = sP + (u & 1) * SM::P_BYTES + (cq >> 1) * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          float pv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            pv[e] = ex2_approx(fmaf(__uint_as_float(sr[c8 * 8 + e]), p.scale_log2, -mb));
            sum4[e & 3] += pv[e];
          }
          uint4 pk = make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]), pack_bf16(pv[6], pv[7]));
          *reinterpret_cast<uint4*>(prow + ((((cq & 1) * 4 + c8) ^ (r & 7)) << 4)) = pk;
        }
        fence_proxy_async_smem();
        mbar_arrive(p_full);
        // ---- 3. fold the PREVIOUS unit's P V product (it ran on the tensor pipe during this unit's softmax)
        if (u > 0) {
#pragma unroll
          for (int i = 0; i < 20; ++i) o[i] *= c_pend;
          fold_o(u - 1);
          if (first) finish_item(item - (int)gridDim.x, l_fin);  // that unit closed the previous item
        }
        c_pend = corr;
        m_run = m_new;
        l_run = l_old * corr + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
      }
    }
    if (n_units > 0) {
#pragma unroll
      for (int i = 0; i < 20; ++i) o[i] *= c_pend;
      fold_o(n_units - 1);
      finish_item(blockIdx.x + (my_items - 1) * gridDim.x, l_run);
    }
    EZB_DBG(if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      atomicAdd(&p.dbg[0], (unsigned long long)c_ws); atomicAdd(&p.dbg[1], (unsigned long long)c_wo); atomicAdd(&p.dbg[2], (unsigned long long)c_bar);
      atomicAdd(&p.dbg[3], (unsigned long long)(clock64() - c_t0));
    })
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc<512>(tmem0);
}

inline int attention_tc(Device& dev, cudaStream_t st, const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, const uint8_t* key_mask,
                        __nv_bfloat16* out, int B, int H, int Lq, int Lk, int Lkpad, int dh, int dhp, int dvp, float scale) {
  if (!((dh == 64 && dhp == 64) || (dh == 72 && dhp == 128))) return fail(EZB_ERR_UNSUPPORTED, "attention_tc: dh %d dhp %d", dh, dhp);
  if (dvp % 16 || dvp > 80 || dvp < 16 || (dvp / 4) % 4) return fail(EZB_ERR_UNSUPPORTED, "attention_tc: dvp %d", dvp);
  const CUtensorMap *tq, *tk, *tv, *tqt, *tkt;
  EZB_TRY(dev.tmaps.get3d(q, dhp, Lq, (uint64_t)B * H, dhp, (uint64_t)Lq * dhp, AT_BQ, &tq));
  EZB_TRY(dev.tmaps.get3d(k, dhp, Lk, (uint64_t)B * H, dhp, (uint64_t)Lk * dhp, AT_BK, &tk));
  EZB_TRY(dev.tmaps.get3d(vt, Lk, dvp, (uint64_t)B * H, Lkpad, (uint64_t)dvp * Lkpad, dvp, &tv));
  tqt = tq; tkt = tk;
  if (dh == 72) {
    EZB_TRY(get3d_sw32(dev.tmaps, q, dhp, Lq, (uint64_t)B * H, dhp, (uint64_t)Lq * dhp, AT_BQ, &tqt));
    EZB_TRY(get3d_sw32(dev.tmaps, k, dhp, Lk, (uint64_t)B * H, dhp, (uint64_t)Lk * dhp, AT_BK, &tkt));
  }
  AttnParams p;
  p.key_mask = key_mask; p.out = out; p.H = H; p.Lq = Lq; p.Lk = Lk; p.dh = dh; p.dvp = dvp;
  p.n_qt = (Lq + AT_BQ - 1) / AT_BQ;
  p.n_items = p.n_qt * B * H;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.dbg = gemm_dbg_buf();
  const int grid = p.n_items < dev.num_sms ? p.n_items : dev.num_sms;
  if (dh == 64) {
    const int smem = AttnSmem<64>::total(dvp);
    static bool set[16] = {};  // function attributes are per device
    if (!set[dev.id & 15]) { EZB_CUDA(cudaFuncSetAttribute(attn_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<64>::total(80))); set[dev.id & 15] = true; }
    EZB_TRY(launch_k(attn_tc_kernel<64>, dim3(grid), dim3(AT_THREADS), smem, st, 1, *tq, *tk, *tv, *tqt, *tkt, p));
  } else {
    const int smem = AttnSmem<72>::total(dvp);
    static bool set[16] = {};
    if (!set[dev.id & 15]) { EZB_CUDA(cudaFuncSetAttribute(attn_tc_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<72>::total(80))); set[dev.id & 15] = true; }
    EZB_TRY(launch_k(attn_tc_kernel<72>, dim3(grid), dim3(AT_THREADS), smem, st, 1, *tq, *tk, *tv, *tqt, *tkt, p));
  }
  return EZB_OK;
}

}  // namespace ezb
