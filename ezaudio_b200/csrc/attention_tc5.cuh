// tcgen05 flash attention v5: the two-tile ping-pong of attention_tc4.cuh with TWO threads per query row.
// Replaces F.scaled_dot_product_attention in src/models/utils/attention.py:107-110 (self: mask None; cross: bool key mask built by
// attention.py:30-37).
//
// Why: in v4 a softmax thread owns one query row and all 128 key columns of a score block; with two groups of four warps that is ONE
// warp of each group per scheduler, each walking a ~600-instruction dependent stream per block.  ncu (profiles/r2): issue slots 32 %, MUFU 36 %,
// tensor pipe 21 % -- nothing saturated, the kernel is latency-bound.  v5 gives every group EIGHT warps: warp (lg, hf) owns rows 32 lg .. 32 lg + 31
// and key columns 64 hf .. 64 hf + 63.  Four warps per scheduler instead of two, half the serial work per thread, half the registers.
// The two threads of a row exchange their partial row maximum through shared memory behind a 64-thread named barrier; that barrier also
// orders "both halves have loaded their S columns" before "either half overwrites S with P" (P of half 1 lands on S columns 32..63 of half 0).
// Row sums stay per thread and are added once per item.  Everything else is v4: S and P share tensor memory (P is the TMEM A operand of the
// P V MMA), O accumulates in TMEM with lazy in-place rescale (only when the row max outgrew its reference by 2^8), static MMA order
// S_A0 S_B0 | PV_A0 S_A1 | PV_B0 S_B1 | ..., K / V^T through 4-deep TMA rings, finished items written out under the next item's first block.
//
// Layouts (QKV GEMM epilogue): Q,K [B*H, L, dhp] bf16 (dh = 72: 64 columns SWIZZLE_128B + a 16-column SWIZZLE_32B tail box; dhp = 80 or 128),
// V^T [B*H, dvp, Lpad] bf16.  Output [B, Lq, H*dh] bf16 token-major.
#pragma once
#include "attention_tc4.cuh"

namespace ezb {

constexpr int A5_SOFTMAX_THREADS = 512;
constexpr int A5_THREADS = A5_SOFTMAX_THREADS + 64;

template <int DH>
__global__ void __launch_bounds__(A5_THREADS, 1)
attn5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
             const __grid_constant__ CUtensorMap tmQt, const __grid_constant__ CUtensorMap tmKt, const Attn4Params p) {
  using SM = Attn4Smem<DH>;
  constexpr bool HAS_TAIL = DH > 64;
  constexpr int OC = (DH > 64 ? 80 : 64) / 2;   // O columns per half: 40 (dvp 80) or 32 (dvp 64)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int VB = SM::v_bytes(p.dvp);
  uint8_t* sQ = smem;                              // [2 groups][Q_BYTES]
  uint8_t* sK = sQ + 2 * SM::Q_BYTES;              // [STAGES][K_BYTES]
  uint8_t* sV = sK + A4_STAGES * SM::K_BYTES;      // [STAGES][VB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A4_STAGES * VB);
  uint64_t *q_full = bars, *q_empty = bars + 2, *s_full = bars + 4, *p_full = bars + 6, *o_full = bars + 8;
  uint64_t *k_full = bars + 10, *k_empty = k_full + A4_STAGES, *v_full = k_empty + A4_STAGES, *v_empty = v_full + A4_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_empty + A4_STAGES);
  float* xmax = reinterpret_cast<float*>(bars) + 128;   // [2 parity][2 groups][2 halves][128 rows]   (bars + 512 bytes)
  float* xsum = xmax + 2 * 2 * 2 * 128;                 // [2 groups][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.Lk + 127) / 128;
  const int my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int U0 = ((my_items + 1) >> 1) * n_kv, U1 = (my_items >> 1) * n_kv;   // units of softmax group 0 / 1

  if (warp == 16) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      if (HAS_TAIL) { tma_prefetch_desc(&tmQt); tma_prefetch_desc(&tmKt); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&o_full[i], 1);
        mbar_init(&p_full[i], A5_SOFTMAX_THREADS / 2);
      }
      for (int i = 0; i < A4_STAGES; ++i) {
        mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;  // S/P of group g @ g*128, O of group g @ 256 + g*128
  pdl_launch();
  pdl_wait();

  if (warp == 17) {
    // ------------------------------------------------ TMA producer (static ping-pong order)
    if (lane == 0) {
      int kc = 0, vc = 0;
      const int maxU = U0 > U1 ? U0 : U1;
      for (int s = 0; s < maxU; ++s) {
        for (int g = 0; g < 2; ++g) {
          if (s >= (g ? U1 : U0)) continue;
          const int itl = s / n_kv, j = s - itl * n_kv;
          const int item = blockIdx.x + (2 * itl + g) * gridDim.x;
          const int bh = item / p.n_qt, q0 = (item - bh * p.n_qt) * 128;
          if (j == 0) {
            mbar_wait(&q_empty[g], (itl & 1) ^ 1);
            mbar_expect_tx(&q_full[g], SM::Q_BYTES);
            tma_load_3d(sQ + g * SM::Q_BYTES, &tmQ, &q_full[g], 0, q0, bh);
            if (HAS_TAIL) tma_load_3d(sQ + g * SM::Q_BYTES + 16384, &tmQt, &q_full[g], 64, q0, bh);
          }
          const int st = kc % A4_STAGES;
          mbar_wait(&k_empty[st], ((kc / A4_STAGES) & 1) ^ 1);
          mbar_expect_tx(&k_full[st], SM::K_BYTES);
          tma_load_3d(sK + st * SM::K_BYTES, &tmK, &k_full[st], 0, j * 128, bh);
          if (HAS_TAIL) tma_load_3d(sK + st * SM::K_BYTES + 16384, &tmKt, &k_full[st], 64, j * 128, bh);
          ++kc;
        }
        for (int g = 0; g < 2; ++g) {
          if (s >= (g ? U1 : U0)) continue;
          const int itl = s / n_kv, j = s - itl * n_kv;
          const int item = blockIdx.x + (2 * itl + g) * gridDim.x;
          const int bh = item / p.n_qt;
          const int st = vc % A4_STAGES;
          mbar_wait(&v_empty[st], ((vc / A4_STAGES) & 1) ^ 1);
          mbar_expect_tx(&v_full[st], VB);
          for (int hh = 0; hh < 2; ++hh) tma_load_3d(sV + st * VB + hh * (VB / 2), &tmV, &v_full[st], j * 128 + hh * 64, 0, bh);
          ++vc;
        }
      }
    }
  } else if (warp == 16) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128), idesc_o = umma_idesc_bf16(128, p.dvp);
      int kc = 0, vc = 0;
      auto issue_s = [&](int g, int s) {
        const int itl = s / n_kv, j = s - itl * n_kv;
        if (j == 0) mbar_wait(&q_full[g], itl & 1);
        const int st = kc % A4_STAGES;
        mbar_wait(&k_full[st], (kc / A4_STAGES) & 1);
        tc_fence_after();
        const uint64_t qd = umma_desc_sw128(smem_u32(sQ + g * SM::Q_BYTES)), kd = umma_desc_sw128(smem_u32(sK + st * SM::K_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem0 + g * 128, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
        if (HAS_TAIL)
          umma_bf16(tmem0 + g * 128, umma_desc_sw32(smem_u32(sQ + g * SM::Q_BYTES + 16384)), umma_desc_sw32(smem_u32(sK + st * SM::K_BYTES + 16384)), idesc_s, 1);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[g]);
        if (j == n_kv - 1) umma_commit(&q_empty[g]);
        ++kc;
      };
      auto issue_pv = [&](int g, int s) {
        const int j = s % n_kv;
        mbar_wait(&p_full[g], s & 1);
        const int st = vc % A4_STAGES;
        mbar_wait(&v_full[st], (vc / A4_STAGES) & 1);
        tc_fence_after();
        for (int hh = 0; hh < 2; ++hh) {
          const uint64_t vd = umma_desc_sw128(smem_u32(sV + st * VB + hh * (VB / 2)));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ts(tmem0 + 256 + g * 128, tmem0 + g * 128 + hh * 32 + k * 8, vd + 2 * k, idesc_o, (j != 0) || ((hh | k) != 0));
        }
        umma_commit(&v_empty[st]);
        umma_commit(&o_full[g]);
        ++vc;
      };
      const int maxU = U0 > U1 ? U0 : U1;
      if (U0 > 0) issue_s(0, 0);
      if (U1 > 0) issue_s(1, 0);
      for (int s = 0; s < maxU; ++s) {
        for (int g = 0; g < 2; ++g) {
          const int Ug = g ? U1 : U0;
          if (s >= Ug) continue;
          issue_pv(g, s);
          if (s + 1 < Ug) issue_s(g, s + 1);
        }
      }
    }
  } else {
    // ------------------------------------------------ softmax groups: warp = g * 8 + hf * 4 + lg
    const int g = warp >> 3, hf = (warp >> 2) & 1, lg = warp & 3;
    const int r = lg * 32 + lane;
    const int pair_bar = 1 + g * 4 + lg;          // named barrier shared by the two warps that own rows 32 lg .. 32 lg + 31 of group g
    const uint32_t t_row = static_cast<uint32_t>(lg * 32) << 16;
    const uint32_t tS = tmem0 + g * 128 + t_row + hf * 64;      // this half's 64 score columns
    const uint32_t tP = tmem0 + g * 128 + t_row + hf * 32;      // its 64 bf16 probabilities (32 words)
    const uint32_t tO = tmem0 + 256 + g * 128 + t_row + hf * OC;
    const int Ug = g ? U1 : U0;
    float m_ref = -INFINITY, l_run = 0.f;
    float* xm_own = xmax + (g * 2 + hf) * 128 + r;              // + parity * 512
    const float* xm_oth = xmax + (g * 2 + (hf ^ 1)) * 128 + r;
    float* xs_own = xsum + (g * 2 + hf) * 128 + r;
    const float* xs_oth = xsum + (g * 2 + (hf ^ 1)) * 128 + r;

    auto write_item = [&](int item, float l_part) {  // O / l of a finished item -> global (its last P V has completed)
      *xs_own = l_part;
      uint32_t orr[32];
      uint32_t o8[8];
      tmem_ld_32x32(tO, orr);
      if (DH > 64 && hf == 0) tmem_ld_32x8(tO + 32, o8);
      tmem_ld_wait();
      tc_fence_before();
      named_bar_sync(pair_bar, 64);
      const float inv = 1.f / (l_part + *xs_oth);
      const int bh = item / p.n_qt, b = bh / p.H, h = bh - b * p.H;
      const int qrow = (item - bh * p.n_qt) * 128 + r;
      if (qrow < p.Lq) {
        uint4* orow = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.Lq + qrow) * (size_t)(p.H * DH) + h * DH + hf * OC);
#pragma unroll
        for (int v = 0; v < 4; ++v)
          orow[v] = make_uint4(pack_bf16(__uint_as_float(orr[8 * v]) * inv, __uint_as_float(orr[8 * v + 1]) * inv),
                               pack_bf16(__uint_as_float(orr[8 * v + 2]) * inv, __uint_as_float(orr[8 * v + 3]) * inv),
                               pack_bf16(__uint_as_float(orr[8 * v + 4]) * inv, __uint_as_float(orr[8 * v + 5]) * inv),
                               pack_bf16(__uint_as_float(orr[8 * v + 6]) * inv, __uint_as_float(orr[8 * v + 7]) * inv));
        if (DH > 64 && hf == 0)   // columns 32..39 of the first half (the second half's columns 72..79 are padding)
          orow[4] = make_uint4(pack_bf16(__uint_as_float(o8[0]) * inv, __uint_as_float(o8[1]) * inv), pack_bf16(__uint_as_float(o8[2]) * inv, __uint_as_float(o8[3]) * inv),
                               pack_bf16(__uint_as_float(o8[4]) * inv, __uint_as_float(o8[5]) * inv), pack_bf16(__uint_as_float(o8[6]) * inv, __uint_as_float(o8[7]) * inv));
      }
    };

    for (int s = 0; s < Ug; ++s) {
      const int itl = s / n_kv, j = s - itl * n_kv;
      const int item = blockIdx.x + (2 * itl + g) * gridDim.x;
      const int bh = item / p.n_qt, b = bh / p.H;
      mbar_wait(&s_full[g], s & 1);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld_32x32(tS, sr);
      tmem_ld_32x32(tS + 32, sr + 32);
      tmem_ld_wait();
      const int kbase = j * 128 + hf * 64;
      const bool full = (p.key_mask == nullptr) && (j * 128 + 128 <= p.Lk);   // warp-uniform
      if (!full) {  // one validity bit per key column, identical for every row: 2 words built with warp ballots
        const uint8_t* km = p.key_mask ? p.key_mask + (size_t)b * p.Lk : nullptr;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int kk = kbase + q * 32 + lane;
          bool ok = kk < p.Lk;
          if (ok && km != nullptr) ok = km[kk] != 0;
          const uint32_t w = __ballot_sync(0xffffffffu, ok);
#pragma unroll
          for (int c = 0; c < 32; ++c) sr[q * 32 + c] = ((w >> c) & 1u) ? sr[q * 32 + c] : 0xff800000u;  // -inf
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(sr[c + e]));
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // row maximum over both halves; the barrier also separates "both halves hold their S columns in registers" from the P stores below
      xm_own[(s & 1) * 512] = mx;
      named_bar_sync(pair_bar, 64);
      mx = fmaxf(mx, xm_oth[(s & 1) * 512]);
      // reference max: fresh for the first key block of an item, afterwards only moved when the row max outgrew it by 2^8
      float fac = 1.f;
      bool need = false;
      const float l_prev = l_run;  // the previous item's partial sum (retired below when j == 0)
      if (j == 0) {
        m_ref = mx;
        l_run = 0.f;
      } else {
        need = (mx - m_ref) * p.scale_log2 > 8.f;  // also true when m_ref = -inf and mx is finite
        if (need) {
          fac = (m_ref == -INFINITY) ? 0.f : ex2_approx((m_ref - mx) * p.scale_log2);
          m_ref = mx;
          l_run *= fac;
        }
      }
      const float mb = (m_ref == -INFINITY) ? 0.f : m_ref * p.scale_log2;  // fully masked so far: exp2(-inf) = 0
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(sr[c]), p.scale_log2, -mb));
        const float p1 = ex2_approx(fmaf(__uint_as_float(sr[c + 1]), p.scale_log2, -mb));
        sum4[(c >> 1) & 3] += p0 + p1;
        pk[c >> 1] = pack_bf16(p0, p1);
      }
      l_run += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      tmem_st_32x32(tP, pk);
      if (s > 0 && j == 0) {  // retire the previous item while this unit's S / P hand-off is in flight: its last P V wrote the O that this
        mbar_wait(&o_full[g], (s - 1) & 1);   // unit's P V (accumulate = 0, issued after our arrive) will overwrite
        tc_fence_after();
        write_item(item - 2 * (int)gridDim.x, l_prev);
      }
      // in-place rescale of the O rows whose reference max moved (warp-collective TMEM access; both halves take the same decision)
      if (j != 0 && __any_sync(0xffffffffu, need)) {
        mbar_wait(&o_full[g], (s - 1) & 1);  // P_{s-1} V_{s-1} has landed
        tc_fence_after();
        uint32_t orr[32];
        tmem_ld_32x32(tO, orr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * fac);
        tmem_st_32x32(tO, orr);
        if (DH > 64) {
          uint32_t o8[8];
          tmem_ld_32x8(tO + 32, o8);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = __float_as_uint(__uint_as_float(o8[i]) * fac);
          tmem_st_32x8(tO + 32, o8);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
    }
    if (Ug > 0) {  // last item of this group
      mbar_wait(&o_full[g], (Ug - 1) & 1);
      tc_fence_after();
      write_item(blockIdx.x + (2 * ((Ug - 1) / n_kv) + g) * (int)gridDim.x, l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc<512>(tmem0);
}

inline int& opt_attn5() {
  static int v = [] { const char* e = getenv("EZB_ATTN5"); return e ? atoi(e) : 0; }();
  return v;
}

inline int attention_tc5(Device& dev, cudaStream_t st, const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, const uint8_t* key_mask,
                         __nv_bfloat16* out, int B, int H, int Lq, int Lk, int Lkpad, int dh, int dhp, int dvp, float scale) {
  if (!((dh == 64 && dhp == 64 && dvp == 64) || (dh == 72 && (dhp == 128 || dhp == 80) && dvp == 80)))
    return fail(EZB_ERR_UNSUPPORTED, "attention_tc5: dh %d dhp %d dvp %d", dh, dhp, dvp);
  const CUtensorMap *tq, *tk, *tv, *tqt, *tkt;
  EZB_TRY(dev.tmaps.get3d(q, dhp, Lq, (uint64_t)B * H, dhp, (uint64_t)Lq * dhp, 128, &tq));
  EZB_TRY(dev.tmaps.get3d(k, dhp, Lk, (uint64_t)B * H, dhp, (uint64_t)Lk * dhp, 128, &tk));
  EZB_TRY(dev.tmaps.get3d(vt, Lk, dvp, (uint64_t)B * H, Lkpad, (uint64_t)dvp * Lkpad, dvp, &tv));
  tqt = tq; tkt = tk;
  if (dh == 72) {
    EZB_TRY(get3d_sw32(dev.tmaps, q, dhp, Lq, (uint64_t)B * H, dhp, (uint64_t)Lq * dhp, 128, &tqt));
    EZB_TRY(get3d_sw32(dev.tmaps, k, dhp, Lk, (uint64_t)B * H, dhp, (uint64_t)Lk * dhp, 128, &tkt));
  }
  Attn4Params p;
  p.key_mask = key_mask; p.out = out; p.H = H; p.Lq = Lq; p.Lk = Lk; p.dvp = dvp;
  p.n_qt = (Lq + 127) / 128;
  p.n_items = p.n_qt * B * H;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.dbg = 0; p.dbg_buf = nullptr;
  const int grid = p.n_items < dev.num_sms ? p.n_items : dev.num_sms;
  auto go = [&](auto kern, int smem) -> int {
    EZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return launch_k(kern, dim3(grid), dim3(A5_THREADS), smem, st, 1, *tq, *tk, *tv, *tqt, *tkt, p);
  };
  const int extra = 8192;   // exchange buffers behind the barriers (xmax 4 KB + xsum 2 KB, offset 512 B)
  if (dh == 64) return go(attn5_kernel<64>, Attn4Smem<64>::total(dvp) + extra);
  return go(attn5_kernel<72>, Attn4Smem<72>::total(dvp) + extra);
}

}  // namespace ezb
