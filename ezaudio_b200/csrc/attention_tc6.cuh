// tcgen05 flash attention, generation 6: the same pipeline as attention_tc4.cuh (two query tiles in flight per persistent CTA, thread = query
// row, P kept in tensor memory, K / V^T rings filled by TMA) with the softmax rebuilt around what the ncu source-level capture of attn4 showed
// (profiles/r2/attention_r2b.md): its softmax warps execute ~750 instructions per 128 x 128 score block at 0.2-0.3 IPC -- 128 scalar FFMA, 130 FADD
// and ~70 register moves around the 128-register score array and the 64-register P array that coexist until the two tcgen05.st at the end -- and both
// groups run their exponent phases at the same time, so the MUFU (16 ex2 / clock / SM) is shared while it is needed and idle while both wait for the
// tensor pipe.  Changes:
//   * the exponent pass walks the score registers in four 32-column chunks; each chunk's 16 packed bf16x2 words go to tensor memory at once
//     (tcgen05.st .x16): no 64-register P array, no register shuffling;
//   * packed arithmetic: fma.rn.f32x2 for x * scale - max and add.rn.f32x2 for the row sum (half the issue slots of the scalar forms);
//   * MUFU token: the exponent phases of the two groups strictly alternate (G0 #0, G1 #0, G0 #1, ...), so one group's exponentials run at the
//     full MUFU rate while the other group's P V / next-S MMAs, tcgen05.ld and row max are in flight;
//   * P is handed to the MMA warp in two halves (64 keys each): the first four P V MMAs run under the second half of the exponentials;
//   * the (rare) in-place O rescale happens BEFORE the exponent phase (it must precede the first P V of the block);
//   * per-item overhead (measured on attn4 from four shapes: time = 0.85 us + 2.2 us x items per CTA + 0.85 us x score blocks per CTA, i.e. 38 % of the
//     XL self-attention launch and 67 % of the cross-attention launch): the finished tile is written out at the END of its item, inside the bubble in
//     which the group waits for the next item's first S anyway, through shared memory and ONE TMA store (attn4: nine 16-byte stores per thread, every
//     warp-level store touching 32 different lines, issued before the group's next hand-off); Q tiles are double-buffered per group and fetched one
//     item ahead by their own producer warp (attn4: one Q buffer per group, reloaded only after the previous item's last S had completed, behind the
//     K / V producer's static order); the staging area of the TMA store is the Q buffer of the item that has just ended.  The K ring is 4 deep, the V^T ring 3 (K 3 vs 4: no measurable difference, call 27).
// Replaces F.scaled_dot_product_attention in src/models/utils/attention.py:107-110 (self: mask None; cross: bool key mask, attention.py:30-37).
// Layouts as produced by the QKV GEMM epilogue: Q, K [B*H, L, DHP] bf16; V^T [B*H, DVP, Lpad] bf16.  Output [B, Lq, H*dh] bf16 token-major.
#pragma once
#include "attention_tc4.cuh"

namespace ezb {

__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
      "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// PP: MUFU token between the softmax groups.  HALF: P handed over in two halves.
constexpr int A6_THREADS = 352;   // warps 0-3 / 4-7 softmax groups, 8 MMA, 9 K / V producer, 10 Q producer
constexpr int A6_KSTAGES = 4, A6_VSTAGES = 3;   // K / V^T ring depths

template <int DH>
struct Attn6Smem {
  static constexpr int TAIL = DH > 64 ? 4096 : 0;
  static constexpr int Q_BYTES = 16384 + TAIL;
  static constexpr int K_BYTES = 16384 + TAIL;
  static __host__ __device__ constexpr int v_bytes(int dvp) { return 2 * dvp * 128; }
  static __host__ __device__ constexpr int total(int dvp) { return 1024 + 4 * Q_BYTES + A6_KSTAGES * K_BYTES + A6_VSTAGES * v_bytes(dvp) + 512; }
  static_assert(128 * DH * 2 <= Q_BYTES, "the output tile is staged in a Q buffer");
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int DH, int PP, int HALF>
__global__ void __launch_bounds__(A6_THREADS, 1)
attn6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
             const __grid_constant__ CUtensorMap tmQt, const __grid_constant__ CUtensorMap tmKt, const __grid_constant__ CUtensorMap tmO, const Attn4Params p) {
  using SM = Attn6Smem<DH>;
  constexpr bool HAS_TAIL = DH > 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int VB = SM::v_bytes(p.dvp);
  uint8_t* sQ = smem;                              // [2 groups][2 buffers][Q_BYTES]; buffer (item & 1) of a group doubles as its output staging tile
  uint8_t* sK = sQ + 4 * SM::Q_BYTES;              // [STAGES][K_BYTES]
  uint8_t* sV = sK + A6_KSTAGES * SM::K_BYTES;     // [VSTAGES][VB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A6_VSTAGES * VB);
  uint64_t *q_full = bars, *q_free = bars + 4;     // [g * 2 + buffer]
  uint64_t *s_full = bars + 8, *p_full = bars + 10, *p_half = bars + 12, *o_full = bars + 14, *tok = bars + 16;
  uint64_t *k_full = bars + 18, *k_empty = k_full + A6_KSTAGES, *v_full = k_empty + A6_KSTAGES, *v_empty = v_full + A6_VSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_empty + A6_VSTAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.Lk + 127) / 128;
  const int my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
#define A6_ITEM(itl_, g_) ((int)blockIdx.x + (2 * (itl_) + (g_)) * (int)gridDim.x)
  const int I0 = (my_items + 1) >> 1, I1 = my_items >> 1;   // items of softmax group 0 / 1
  const int U0 = I0 * n_kv, U1 = I1 * n_kv;                 // score blocks of softmax group 0 / 1

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
      if (HAS_TAIL) { tma_prefetch_desc(&tmQt); tma_prefetch_desc(&tmKt); }
      for (int i = 0; i < 4; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_free[i], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1); mbar_init(&o_full[i], 1);
        mbar_init(&p_full[i], 4); mbar_init(&p_half[i], 4); mbar_init(&tok[i], 4);   // one elected arrival per softmax warp
      }
      for (int i = 0; i < A6_KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
      for (int i = 0; i < A6_VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
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

  if (warp == 10) {
    // ------------------------------------------------ Q producer: item (g, itl) goes to buffer itl & 1 of its group, one item ahead of its use.
    // A buffer is free again when the TMA store that staged the output of its previous item (itl - 2) has read it (q_free, arrived by the
    // storing thread); the first two items of a group find their buffers free.
    for (int itl = 0; itl < I0; ++itl) {
      for (int g = 0; g < 2; ++g) {
        if (itl >= (g ? I1 : I0)) continue;
        const int bf = g * 2 + (itl & 1), u = itl >> 1;
        if (u > 0) mbar_wait(&q_free[bf], (u - 1) & 1);
        if (elect_one()) {
          const int item = A6_ITEM(itl, g);
          const int bh = item / p.n_qt, q0 = (item - bh * p.n_qt) * 128;
          mbar_expect_tx(&q_full[bf], SM::Q_BYTES);
          tma_load_3d(sQ + bf * SM::Q_BYTES, &tmQ, &q_full[bf], 0, q0, bh);
          if (HAS_TAIL) tma_load_3d(sQ + bf * SM::Q_BYTES + 16384, &tmQt, &q_full[bf], 64, q0, bh);
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------ K / V^T producer (static ping-pong order; warp-uniform, copies issued under elect.sync)
    int kc = 0, vc = 0;
    const int maxU = U0 > U1 ? U0 : U1;
    int itl = 0, j = 0;   // s = itl * n_kv + j
    for (int s = 0; s < maxU; ++s) {
      for (int g = 0; g < 2; ++g) {
        if (s >= (g ? U1 : U0)) continue;
        const int bh = A6_ITEM(itl, g) / p.n_qt;
        const int st = kc % A6_KSTAGES;
        mbar_wait(&k_empty[st], ((kc / A6_KSTAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&k_full[st], SM::K_BYTES);
          tma_load_3d(sK + st * SM::K_BYTES, &tmK, &k_full[st], 0, j * 128, bh);
          if (HAS_TAIL) tma_load_3d(sK + st * SM::K_BYTES + 16384, &tmKt, &k_full[st], 64, j * 128, bh);
        }
        __syncwarp();
        ++kc;
      }
      for (int g = 0; g < 2; ++g) {
        if (s >= (g ? U1 : U0)) continue;
        const int bh = A6_ITEM(itl, g) / p.n_qt;
        const int st = vc % A6_VSTAGES;
        mbar_wait(&v_empty[st], ((vc / A6_VSTAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&v_full[st], VB);
          for (int hh = 0; hh < 2; ++hh) tma_load_3d(sV + st * VB + hh * (VB / 2), &tmV, &v_full[st], j * 128 + hh * 64, 0, bh);
        }
        __syncwarp();
        ++vc;
      }
      if (++j == n_kv) { j = 0; ++itl; }
    }
  } else if (warp == 8) {
    // ------------------------------------------------ MMA issuer: warp-uniform control flow, one elected lane issues.
    // Order  S_0(0) S_1(0) | PV_0(0) S_0(1) | PV_1(0) S_1(1) | ...   (S_{u+1} of a group aliases P_u: it is issued after PV_u; in-order tensor pipe)
    const uint32_t idesc_s = umma_idesc_bf16(128, 128), idesc_o = umma_idesc_bf16(128, p.dvp);
    int kc = 0, vc = 0;
    int sj[2] = {0, 0}, sit[2] = {0, 0};   // next score block per group: key block, local item index
    int pj[2] = {0, 0};                    // next P V per group: key block
    auto issue_s = [&](int g) {
      const int itl = sit[g], j = sj[g];
      const int bf = g * 2 + (itl & 1);
      if (j == 0) mbar_wait(&q_full[bf], (itl >> 1) & 1);
      const int st = kc % A6_KSTAGES;
      mbar_wait(&k_full[st], (kc / A6_KSTAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qd = umma_desc_sw128(smem_u32(sQ + bf * SM::Q_BYTES)), kd = umma_desc_sw128(smem_u32(sK + st * SM::K_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem0 + g * 128, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
        if (HAS_TAIL)
          umma_bf16(tmem0 + g * 128, umma_desc_sw32(smem_u32(sQ + bf * SM::Q_BYTES + 16384)), umma_desc_sw32(smem_u32(sK + st * SM::K_BYTES + 16384)), idesc_s, 1);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[g]);
      }
      __syncwarp();
      ++kc;
      if (++sj[g] == n_kv) { sj[g] = 0; ++sit[g]; }
    };
    auto issue_pv = [&](int g, int s) {
      const int j = pj[g];
      const int st = vc % A6_VSTAGES;
      if (HALF) {
        mbar_wait(&p_half[g], s & 1);
        mbar_wait(&v_full[st], (vc / A6_VSTAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t vd = umma_desc_sw128(smem_u32(sV + st * VB));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem0 + 256 + g * 128, tmem0 + g * 128 + k * 8, vd + 2 * k, idesc_o, (j != 0) || (k != 0));
        }
        __syncwarp();
      }
      mbar_wait(&p_full[g], s & 1);
      if (!HALF) mbar_wait(&v_full[st], (vc / A6_VSTAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        for (int hh = HALF ? 1 : 0; hh < 2; ++hh) {
          const uint64_t vd = umma_desc_sw128(smem_u32(sV + st * VB + hh * (VB / 2)));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ts(tmem0 + 256 + g * 128, tmem0 + g * 128 + hh * 32 + k * 8, vd + 2 * k, idesc_o, (j != 0) || ((hh | k) != 0));
        }
        umma_commit(&v_empty[st]);
        umma_commit(&o_full[g]);
      }
      __syncwarp();
      ++vc;
      if (++pj[g] == n_kv) pj[g] = 0;
    };
    const int maxU = U0 > U1 ? U0 : U1;
    if (U0 > 0) issue_s(0);
    if (U1 > 0) issue_s(1);
    for (int s = 0; s < maxU; ++s) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int Ug = g ? U1 : U0;
        if (s >= Ug) continue;
        issue_pv(g, s);
        if (s + 1 < Ug) issue_s(g);
      }
    }
  } else {
    // ------------------------------------------------ softmax groups
    const int g = warp >> 2, lg = warp & 3;
    const int r = lg * 32 + lane;
    const uint32_t t_row = static_cast<uint32_t>(lg * 32) << 16;
    const uint32_t tS = tmem0 + g * 128 + t_row, tO = tmem0 + 256 + g * 128 + t_row;
    const int Ug = g ? U1 : U0;
    float m_ref = -INFINITY, l_run = 0.f;
    int release_bf = -1;   // thread r == 0 only: Q buffer whose TMA store has been issued but whose q_free arrival is still owed

    // O / l of the item that has just ended -> its own (dead) Q buffer -> one TMA store.  The caller has waited for the item's last P V.
    auto write_item = [&](int item, int bf, float l_fin) {
      uint32_t orr[64];
      uint32_t o8[8];
      tmem_ld_32x64(tO, orr);
      if (DH > 64) tmem_ld_32x8(tO + 64, o8);
      tmem_ld_wait();
      tc_fence_before();
      const float inv = 1.f / l_fin;
      uint4* orow = reinterpret_cast<uint4*>(sQ + bf * SM::Q_BYTES + r * (DH * 2));
#pragma unroll
      for (int v = 0; v < 8; ++v)
        orow[v] = make_uint4(pack_bf16(__uint_as_float(orr[8 * v]) * inv, __uint_as_float(orr[8 * v + 1]) * inv),
                             pack_bf16(__uint_as_float(orr[8 * v + 2]) * inv, __uint_as_float(orr[8 * v + 3]) * inv),
                             pack_bf16(__uint_as_float(orr[8 * v + 4]) * inv, __uint_as_float(orr[8 * v + 5]) * inv),
                             pack_bf16(__uint_as_float(orr[8 * v + 6]) * inv, __uint_as_float(orr[8 * v + 7]) * inv));
      if (DH > 64)
        orow[8] = make_uint4(pack_bf16(__uint_as_float(o8[0]) * inv, __uint_as_float(o8[1]) * inv), pack_bf16(__uint_as_float(o8[2]) * inv, __uint_as_float(o8[3]) * inv),
                             pack_bf16(__uint_as_float(o8[4]) * inv, __uint_as_float(o8[5]) * inv), pack_bf16(__uint_as_float(o8[6]) * inv, __uint_as_float(o8[7]) * inv));
      fence_proxy_async_smem();          // generic-proxy writes -> visible to the TMA (async proxy) read
      named_bar_sync(1 + g, 128);        // all 128 rows of the tile are staged
      if (r == 0) {
        const int bh = item / p.n_qt, b = bh / p.H, h = bh - b * p.H;
        tma_store_3d(&tmO, sQ + bf * SM::Q_BYTES, h * DH, (item - bh * p.n_qt) * 128, b);   // rows >= Lq are clipped by the tensor map
        bulk_commit();
        release_bf = bf;
      }
    };

    uint8_t pm[4] = {1, 1, 1, 1};   // key-mask bytes of the NEXT block's columns q * 32 + lane (1 without a mask)
    auto load_mask = [&](int itl_, int j_) {
      if (p.key_mask == nullptr) return;
      const int bb = (A6_ITEM(itl_, g) / p.n_qt) / p.H;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kk = j_ * 128 + q * 32 + lane;
        pm[q] = kk < p.Lk ? p.key_mask[(size_t)bb * p.Lk + kk] : (uint8_t)0;
      }
    };
    if (Ug > 0) load_mask(0, 0);
    int itl = 0, j = 0;   // s = itl * n_kv + j
    for (int s = 0; s < Ug; ++s) {
      const int item = A6_ITEM(itl, g);
      // key validity bits of this block (one per key column, identical for every row).  With a key mask they come from global memory: the bytes
      // were requested one block ahead (load_mask below; cross-attention: a dependent ~800-cycle load per item if issued after S has arrived)
      const int kbase = j * 128;
      const bool full = (p.key_mask == nullptr) && (kbase + 128 <= p.Lk);
      uint32_t kw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (!full) {
#pragma unroll
        for (int q = 0; q < 4; ++q) kw[q] = __ballot_sync(0xffffffffu, (kbase + q * 32 + lane < p.Lk) && pm[q] != 0);
      }
      mbar_wait(&s_full[g], s & 1);
      tc_fence_after();
      if (release_bf >= 0) {   // the store issued at the end of the previous item has long read its staging tile: hand the buffer back
        bulk_wait_read0();
        mbar_arrive(&q_free[release_bf]);
        release_bf = -1;
      }
      // pass 1: row max over the 128 key columns (the score registers die here: the exponent pass below re-reads S from tensor memory chunk by
      // chunk, so that nothing but a 32-column window is live across the write-out / rescale branches and the token wait)
      float mx;
      {
        uint32_t sr[128];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_32x32(tS + q * 32, sr + q * 32);
        tmem_ld_wait();
        if (!full) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int c = 0; c < 32; ++c) sr[q * 32 + c] = ((kw[q] >> c) & 1u) ? sr[q * 32 + c] : 0xff800000u;  // -inf
          }
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 128; c += 4) {
#pragma unroll
          for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(sr[c + e]));
        }
        mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      }
      // reference max: fresh for the first key block of an item, afterwards only moved when the row max outgrew it by 2^8
      float fac = 1.f;
      bool need = false;
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
      if (j != 0 && __any_sync(0xffffffffu, need)) {   // warp-collective TMEM access: every lane takes part
        mbar_wait(&o_full[g], (s - 1) & 1);  // P_{s-1} V_{s-1} has landed
        tc_fence_after();
        uint32_t orr[64];
        tmem_ld_32x64(tO, orr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 64; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * fac);
        tmem_st_32x32(tO, orr);
        tmem_st_32x32(tO + 32, orr + 32);
        if (DH > 64) {
          uint32_t o8[8];
          tmem_ld_32x8(tO + 64, o8);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = __float_as_uint(__uint_as_float(o8[i]) * fac);
          tmem_st_32x8(tO + 64, o8);
        }
        tmem_st_wait();
      }
      const float mb = (m_ref == -INFINITY) ? 0.f : m_ref * p.scale_log2;  // fully masked so far: exp2(-inf) = 0
      if (PP && !(g == 0 && s == 0)) mbar_wait(&tok[g], (g == 0 ? s - 1 : s) & 1);   // MUFU token
      const unsigned long long sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nmb2 = pack_f32x2(-mb, -mb);
      unsigned long long sum2a = 0ull, sum2b = 0ull;   // two packed accumulators (4 partial sums)
      // pass 2: exponentials, 32 columns at a time; chunk q + 1 is in flight (tcgen05.ld) while chunk q is processed.  P chunk q (16 packed
      // columns at 16 q) lands on score columns that were read before (chunk q / 2), so the in-place overwrite is safe in program order.
      uint32_t sa[32], sb[32];
      tmem_ld_32x32(tS, sa);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t* cur = (q & 1) ? sb : sa;
        uint32_t* nxt = (q & 1) ? sa : sb;
        tmem_ld_wait();
        if (q < 3) tmem_ld_32x32(tS + (q + 1) * 32, nxt);
        if (!full) {
#pragma unroll
          for (int c = 0; c < 32; ++c) cur[c] = ((kw[q] >> c) & 1u) ? cur[c] : 0xff800000u;  // -inf
        }
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const unsigned long long xa = fma_f32x2(pack_f32x2(__uint_as_float(cur[c]), __uint_as_float(cur[c + 1])), sc2, nmb2);
          const unsigned long long xb = fma_f32x2(pack_f32x2(__uint_as_float(cur[c + 2]), __uint_as_float(cur[c + 3])), sc2, nmb2);
          float a0, a1, b0, b1;
          unpack_f32x2(xa, a0, a1);
          unpack_f32x2(xb, b0, b1);
          a0 = ex2_approx(a0); a1 = ex2_approx(a1); b0 = ex2_approx(b0); b1 = ex2_approx(b1);
          sum2a = add_f32x2(sum2a, pack_f32x2(a0, a1));
          sum2b = add_f32x2(sum2b, pack_f32x2(b0, b1));
          pk[c >> 1] = pack_bf16(a0, a1);
          pk[(c >> 1) + 1] = pack_bf16(b0, b1);
        }
        tmem_st_32x16(tS + q * 16, pk);
        if (HALF && q == 1) {   // first 64 keys of P: the MMA warp may start P V
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_half[g]);
        }
      }
      if (PP) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&tok[g ^ 1]);
      }
      {
        float s0, s1, s2, s3;
        unpack_f32x2(sum2a, s0, s1);
        unpack_f32x2(sum2b, s2, s3);
        l_run += (s0 + s1) + (s2 + s3);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      if (s + 1 < Ug) load_mask(j == n_kv - 1 ? itl + 1 : itl, j == n_kv - 1 ? 0 : j + 1);   // mask bytes of the next block: a whole block to arrive
      if (j == n_kv - 1) {   // the item ends here: retire it inside the bubble before the next item's first S (its last P V is the next thing on the pipe)
        mbar_wait(&o_full[g], s & 1);
        tc_fence_after();
        write_item(item, g * 2 + (itl & 1), l_run);
      }
      if (++j == n_kv) { j = 0; ++itl; }
    }
    if (PP && g == 1) {   // group 0 has n_kv more blocks than group 1 when the CTA's item count is odd: keep handing the token back
      for (int n = U1; n < U0 - 1; ++n) {
        mbar_wait(&tok[1], n & 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(&tok[0]);
      }
    }
    if (r == 0) bulk_wait_read0();   // the last store has left shared memory before the CTA goes away (its global writes complete with the grid)
  }
#undef A6_ITEM
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem0);
}

inline int& opt_attn6() {   // attention kernel generation: 6 (this file; bit 0 on, bit 1 MUFU token, bit 2 P in two halves) or 4 (attention_tc4.cuh)
  static int v = [] { const char* e = getenv("EZB_ATTN6"); return e ? atoi(e) : 5; }();   // default: generation 6, P in two halves (call 19: self 23.1 -> 21.5 us, cross 13.4 -> 10.2 us)
  return v;
}

inline int attention_tc6(Device& dev, cudaStream_t st, const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, const uint8_t* key_mask,
                         __nv_bfloat16* out, int B, int H, int Lq, int Lk, int Lkpad, int dh, int dhp, int dvp, float scale) {
  if (!((dh == 64 && dhp == 64 && dvp == 64) || (dh == 72 && (dhp == 128 || dhp == 80) && dvp == 80))) return fail(EZB_ERR_UNSUPPORTED, "attention_tc6: dh %d dhp %d dvp %d", dh, dhp, dvp);
  const CUtensorMap *tq, *tk, *tv, *tqt, *tkt, *to;
  EZB_TRY(dev.tmaps.get3d(q, dhp, Lq, (uint64_t)B * H, dhp, (uint64_t)Lq * dhp, 128, &tq));
  EZB_TRY(dev.tmaps.get3d(k, dhp, Lk, (uint64_t)B * H, dhp, (uint64_t)Lk * dhp, 128, &tk));
  EZB_TRY(dev.tmaps.get3d(vt, Lk, dvp, (uint64_t)B * H, Lkpad, (uint64_t)dvp * Lkpad, dvp, &tv));
  EZB_TRY(get3d_plain(dev.tmaps, out, (uint64_t)H * dh, Lq, B, (uint64_t)H * dh, (uint64_t)Lq * H * dh, dh, 128, &to));
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
  p.pp = 0; p.dbg = 0; p.dbg_buf = nullptr;
  const int grid = p.n_items < dev.num_sms ? p.n_items : dev.num_sms;
  auto go = [&](auto kern, int smem) -> int {
    EZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return launch_k(kern, dim3(grid), dim3(A6_THREADS), smem, st, 1, *tq, *tk, *tv, *tqt, *tkt, *to, p);
  };
  const int mode = (opt_attn6() >> 1) & 3;   // bit 0: token, bit 1: halves
  if (dh == 64) {
    const int smem = Attn6Smem<64>::total(dvp);
    switch (mode) {
      case 0: return go(attn6_kernel<64, 0, 0>, smem);
      case 1: return go(attn6_kernel<64, 1, 0>, smem);
      case 2: return go(attn6_kernel<64, 0, 1>, smem);
      default: return go(attn6_kernel<64, 1, 1>, smem);
    }
  }
  const int smem = Attn6Smem<72>::total(dvp);
  switch (mode) {
    case 0: return go(attn6_kernel<72, 0, 0>, smem);
    case 1: return go(attn6_kernel<72, 1, 0>, smem);
    case 2: return go(attn6_kernel<72, 0, 1>, smem);
    default: return go(attn6_kernel<72, 1, 1>, smem);
  }
}

}  // namespace ezb
