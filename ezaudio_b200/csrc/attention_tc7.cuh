// tcgen05 flash attention, generation 7: generation 6 (attention_tc6.cuh: two query tiles in flight, Q prefetched by its own producer warp, TMA-store
// write-out, packed softmax arithmetic) with 64-key score blocks and TWO score buffers per softmax group.
//
// Why: with one S buffer per group S_{j+1} aliases P_j, so the tensor pipe can only start it after P V_j, i.e. after softmax_j -- each group runs the
// dependent chain softmax_j -> P V_j -> S_{j+1} -> softmax_{j+1} (~2100 cycles per 128 keys, profiles/r2/attention_r2b.md) and two groups overlap it
// only 1.5x.  Two 128-column S buffers per group do not fit (2 x (2 x 128 + 80) = 672 tensor-memory columns for dh = 72); two 64-column ones do
// (2 x (2 x 64 + 80) = 416).  Per group and 64-key block u:
//     MMA warp   : S(0) S(1) | wait P(u): PV(u) -> O (+)= P(u) V(u), then S(u + 2) into the buffer P(u) has just left
//     softmax    : wait S(u): tcgen05.ld 64 columns, row max, (rare) O rescale, 64 exponentials, P(u) bf16 -> the first 32 columns of the same buffer
// so the next block's scores are already in tensor memory when a group finishes its exponentials: the groups wait for the MUFU, not for the tensor pipe.
// K / V^T still travel as 128-key tiles through the rings (K: 4 deep, V^T: 3 deep); a tile serves two consecutive blocks (K rows 0-63 / 64-127 of the
// SWIZZLE_128B tile = descriptor + 8 KB; the V^T tile is already stored as two 64-key halves).
// Replaces F.scaled_dot_product_attention in src/models/utils/attention.py:107-110 (self: mask None; cross: bool key mask, attention.py:30-37).
#pragma once
#include "attention_tc6.cuh"

namespace ezb {

constexpr int A7_KSTAGES = 4, A7_VSTAGES = 3;

template <int DH>
struct Attn7Smem {
  static constexpr int TAIL = DH > 64 ? 4096 : 0;
  static constexpr int Q_BYTES = 16384 + TAIL;
  static constexpr int K_BYTES = 16384 + TAIL;
  static __host__ __device__ constexpr int v_bytes(int dvp) { return 2 * dvp * 128; }
  static __host__ __device__ constexpr int total(int dvp) { return 1024 + 4 * Q_BYTES + A7_KSTAGES * K_BYTES + A7_VSTAGES * v_bytes(dvp) + 512; }
  static_assert(128 * DH * 2 <= Q_BYTES, "the output tile is staged in a Q buffer");
};

template <int DH>
__global__ void __launch_bounds__(A6_THREADS, 1)
attn7_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
             const __grid_constant__ CUtensorMap tmQt, const __grid_constant__ CUtensorMap tmKt, const __grid_constant__ CUtensorMap tmO, const Attn4Params p) {
  using SM = Attn7Smem<DH>;
  constexpr bool HAS_TAIL = DH > 64;
  constexpr uint32_t GCOLS = 256;   // tensor-memory columns per group: S0 @ 0, S1 @ 64, O @ 128 (<= 80 columns); 32-column aligned bases (208 per group failed on hardware)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int VB = SM::v_bytes(p.dvp);
  uint8_t* sQ = smem;                              // [2 groups][2 buffers][Q_BYTES]; buffer (item & 1) of a group doubles as its output staging tile
  uint8_t* sK = sQ + 4 * SM::Q_BYTES;              // [KSTAGES][K_BYTES]
  uint8_t* sV = sK + A7_KSTAGES * SM::K_BYTES;     // [VSTAGES][VB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A7_VSTAGES * VB);
  uint64_t *q_full = bars, *o_staged = bars + 4;   // [g * 2 + buffer]: Q tile has landed / the item's output tile is staged in the buffer
  uint64_t *s_full = bars + 8, *p_full = bars + 12; // [g * 2 + (u & 1)]
  // [g * 2 + (u & 1)]: P V(u) has completed.  Two barriers per group: scores are issued two blocks ahead, so when a softmax warp has S(u) only
  // P V(u - 2) is known to be complete -- a single barrier per group would be one or two phases behind the phase waited for, and a parity wait two
  // phases ahead returns at once (first hardware run: the item-end read of O raced with the last two P V of the item whenever the MMA warp was busy
  // with the other group).
  uint64_t* o_full = bars + 16;
  uint64_t *k_full = bars + 20, *k_empty = k_full + A7_KSTAGES, *v_full = k_empty + A7_KSTAGES, *v_empty = v_full + A7_VSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_empty + A7_VSTAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.Lk + 127) / 128;             // 128-key tiles per item
  const int nb = (p.Lk + 63) / 64;                 // 64-key blocks per item (the last tile may hold one only)
  const int my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
#define A7_ITEM(itl_, g_) ((int)blockIdx.x + (2 * (itl_) + (g_)) * (int)gridDim.x)
  const int I0 = (my_items + 1) >> 1, I1 = my_items >> 1;   // items of softmax group 0 / 1
  const int T0 = I0 * n_kv, T1 = I1 * n_kv;                 // key tiles of group 0 / 1
  const int B0 = I0 * nb, B1 = I1 * nb;                     // 64-key blocks of group 0 / 1

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
      if (HAS_TAIL) { tma_prefetch_desc(&tmQt); tma_prefetch_desc(&tmKt); }
      for (int i = 0; i < 4; ++i) { mbar_init(&q_full[i], 1); mbar_init(&o_staged[i], 4); mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); }
      for (int i = 0; i < 4; ++i) mbar_init(&o_full[i], 1);
      for (int i = 0; i < A7_KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
      for (int i = 0; i < A7_VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;
  pdl_launch();
  pdl_wait();

  if (warp == 10) {
    // ------------------------------------------------ Q loads and output stores.  Item (g, itl) uses buffer itl & 1 of its group: its Q tile is
    // fetched one item ahead; when the item ends the softmax group stages the normalised output tile in the same (dead) buffer and arrives on
    // o_staged; THIS warp issues the TMA store, waits until the tile has been read out of shared memory and only then fetches the Q tile of item
    // itl + 2 into it.  (Generation 6 issued the store from a softmax thread and released the buffer one block later.)
    auto store_item = [&](int itl, int g) {   // elected lane only
      const int item = A7_ITEM(itl, g);
      const int bh = item / p.n_qt, b = bh / p.H, h = bh - b * p.H;
      tma_store_3d(&tmO, sQ + (g * 2 + (itl & 1)) * SM::Q_BYTES, h * DH, (item - bh * p.n_qt) * 128, b);   // rows >= Lq are clipped by the tensor map
      bulk_commit();
    };
    for (int itl = 0; itl < I0; ++itl) {
      for (int g = 0; g < 2; ++g) {
        if (itl >= (g ? I1 : I0)) continue;
        const int bf = g * 2 + (itl & 1), u = itl >> 1;
        if (u > 0) mbar_wait(&o_staged[bf], (u - 1) & 1);   // item itl - 2 has ended and its output sits in this buffer
        if (elect_one()) {
          if (u > 0) { store_item(itl - 2, g); bulk_wait_read0(); }
          const int item = A7_ITEM(itl, g);
          const int bh = item / p.n_qt, q0 = (item - bh * p.n_qt) * 128;
          mbar_expect_tx(&q_full[bf], SM::Q_BYTES);
          tma_load_3d(sQ + bf * SM::Q_BYTES, &tmQ, &q_full[bf], 0, q0, bh);
          if (HAS_TAIL) tma_load_3d(sQ + bf * SM::Q_BYTES + 16384, &tmQt, &q_full[bf], 64, q0, bh);
        }
        __syncwarp();
      }
    }
    for (int back = 2; back >= 1; --back) {   // the last two items of each group: nothing follows them into their buffers
      for (int g = 0; g < 2; ++g) {
        const int Ig = g ? I1 : I0, itl = Ig - back;
        if (itl < 0) continue;
        mbar_wait(&o_staged[g * 2 + (itl & 1)], (itl >> 1) & 1);
        if (elect_one()) store_item(itl, g);
        __syncwarp();
      }
    }
    if (elect_one()) bulk_wait_read0();   // the last stores have left shared memory before the CTA goes away (their global writes complete with the grid)
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------ K / V^T producer: 128-key tiles in the order the MMA warp consumes them (tile t of group 0, of group 1, ...)
    int kc = 0, vc = 0;
    const int maxT = T0 > T1 ? T0 : T1;
    int itl = 0, j = 0;   // t = itl * n_kv + j
    for (int t = 0; t < maxT; ++t) {
      for (int g = 0; g < 2; ++g) {
        if (t >= (g ? T1 : T0)) continue;
        const int bh = A7_ITEM(itl, g) / p.n_qt;
        const int st = kc % A7_KSTAGES;
        mbar_wait(&k_empty[st], ((kc / A7_KSTAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&k_full[st], SM::K_BYTES);
          tma_load_3d(sK + st * SM::K_BYTES, &tmK, &k_full[st], 0, j * 128, bh);
          if (HAS_TAIL) tma_load_3d(sK + st * SM::K_BYTES + 16384, &tmKt, &k_full[st], 64, j * 128, bh);
        }
        __syncwarp();
        ++kc;
      }
      for (int g = 0; g < 2; ++g) {
        if (t >= (g ? T1 : T0)) continue;
        const int bh = A7_ITEM(itl, g) / p.n_qt;
        const int st = vc % A7_VSTAGES;
        mbar_wait(&v_empty[st], ((vc / A7_VSTAGES) & 1) ^ 1);
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
    // ------------------------------------------------ MMA issuer: warp-uniform control flow, one elected lane issues.  Per group the ring slot of a
    // key tile is found from the tile's global sequence number 2 * t + g (both groups, every tile, in producer order: only valid while both groups still
    // have tiles; a group that outlives the other -- odd item count -- continues with consecutive numbers, exactly as the producer does).
    const uint32_t idesc_s = umma_idesc_bf16(128, 64), idesc_o = umma_idesc_bf16(128, p.dvp);
    // sequence number of tile t of group g in the producer's order
    auto seq = [&](int g, int t) -> int { const int both = T1; return t < both ? 2 * t + g : 2 * both + (t - both); };
    int sb[2] = {0, 0}, sit[2] = {0, 0};   // next S per group: block within the item, local item index
    int pb[2] = {0, 0}, pit[2] = {0, 0};   // next P V per group
    int su[2] = {0, 0}, pu[2] = {0, 0};    // running block counters (buffer / phase arithmetic)
    auto issue_s = [&](int g) {
      const int itl = sit[g], b = sb[g], u = su[g];
      const int bf = g * 2 + (itl & 1);
      if (b == 0) mbar_wait(&q_full[bf], (itl >> 1) & 1);
      const int kc = seq(g, itl * n_kv + (b >> 1)), st = kc % A7_KSTAGES;
      if ((b & 1) == 0) mbar_wait(&k_full[st], (kc / A7_KSTAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem0 + g * GCOLS + (u & 1) * 64;
        const uint64_t qd = umma_desc_sw128(smem_u32(sQ + bf * SM::Q_BYTES));
        const uint64_t kd = umma_desc_sw128(smem_u32(sK + st * SM::K_BYTES + (b & 1) * 8192));   // key rows 64 (b & 1) .. + 63 of the tile
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
        if (HAS_TAIL)
          umma_bf16(d, umma_desc_sw32(smem_u32(sQ + bf * SM::Q_BYTES + 16384)), umma_desc_sw32(smem_u32(sK + st * SM::K_BYTES + 16384 + (b & 1) * 2048)), idesc_s, 1);
        if ((b & 1) == 1 || b == nb - 1) umma_commit(&k_empty[st]);   // last block that reads this tile
        umma_commit(&s_full[g * 2 + (u & 1)]);
      }
      __syncwarp();
      ++su[g];
      if (++sb[g] == nb) { sb[g] = 0; ++sit[g]; }
    };
    auto issue_pv = [&](int g) {
      const int itl = pit[g], b = pb[g], u = pu[g];
      const int vc = seq(g, itl * n_kv + (b >> 1)), st = vc % A7_VSTAGES;
      mbar_wait(&p_full[g * 2 + (u & 1)], (u >> 1) & 1);
      if ((b & 1) == 0) mbar_wait(&v_full[st], (vc / A7_VSTAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t pa = tmem0 + g * GCOLS + (u & 1) * 64, od = tmem0 + g * GCOLS + 128;
        const uint64_t vd = umma_desc_sw128(smem_u32(sV + st * VB + (b & 1) * (VB / 2)));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ts(od, pa + k * 8, vd + 2 * k, idesc_o, (b != 0) || (k != 0));
        if ((b & 1) == 1 || b == nb - 1) umma_commit(&v_empty[st]);
        umma_commit(&o_full[g * 2 + (u & 1)]);
      }
      __syncwarp();
      ++pu[g];
      if (++pb[g] == nb) { pb[g] = 0; ++pit[g]; }
    };
    const int maxB = B0 > B1 ? B0 : B1;
    for (int w = 0; w < 2; ++w) {   // two blocks of scores ahead per group
      if (w < B0) issue_s(0);
      if (w < B1) issue_s(1);
    }
    for (int u = 0; u < maxB; ++u) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int Bg = g ? B1 : B0;
        if (u >= Bg) continue;
        issue_pv(g);
        if (u + 2 < Bg) issue_s(g);
      }
    }
  } else {
    // ------------------------------------------------ softmax groups
    const int g = warp >> 2, lg = warp & 3;
    const int r = lg * 32 + lane;
    const uint32_t t_row = static_cast<uint32_t>(lg * 32) << 16;
    const uint32_t tG = tmem0 + g * GCOLS + t_row, tO = tG + 128;
    const int Bg = g ? B1 : B0;
    float m_ref = -INFINITY, l_run = 0.f;
    // O / l of the item that has just ended -> its own (dead) Q buffer; the Q / store warp sends it off.  The caller has waited for the item's last P V.
    auto stage_item = [&](int bf, float l_fin) {
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
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_staged[bf]);
    };

    uint8_t pm[2] = {1, 1};   // key-mask bytes of the NEXT block's columns q * 32 + lane (1 without a mask)
    auto load_mask = [&](int itl_, int b_) {
      if (p.key_mask == nullptr) return;
      const int bb = (A7_ITEM(itl_, g) / p.n_qt) / p.H;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int kk = b_ * 64 + q * 32 + lane;
        pm[q] = kk < p.Lk ? p.key_mask[(size_t)bb * p.Lk + kk] : (uint8_t)0;
      }
    };
    if (Bg > 0) load_mask(0, 0);
    int itl = 0, b = 0;   // u = itl * nb + b
    for (int u = 0; u < Bg; ++u) {
      const uint32_t tS = tG + (u & 1) * 64;
      const int kbase = b * 64;
      const bool full = (p.key_mask == nullptr) && (kbase + 64 <= p.Lk);
      uint32_t kw[2] = {0xffffffffu, 0xffffffffu};   // one validity bit per key column, identical for every row
      if (!full) {
#pragma unroll
        for (int q = 0; q < 2; ++q) kw[q] = __ballot_sync(0xffffffffu, (kbase + q * 32 + lane < p.Lk) && pm[q] != 0);
      }
      mbar_wait(&s_full[g * 2 + (u & 1)], (u >> 1) & 1);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld_32x32(tS, sr);
      tmem_ld_32x32(tS + 32, sr + 32);
      tmem_ld_wait();
      if (!full) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
          for (int c = 0; c < 32; ++c) sr[q * 32 + c] = ((kw[q] >> c) & 1u) ? sr[q * 32 + c] : 0xff800000u;  // -inf
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(sr[c + e]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // reference max: fresh for the first key block of an item, afterwards only moved when the row max outgrew it by 2^8
      float fac = 1.f;
      bool need = false;
      if (b == 0) {
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
      if (b != 0 && __any_sync(0xffffffffu, need)) {   // in-place rescale of O before this block's P V (warp-collective TMEM access)
        mbar_wait(&o_full[g * 2 + ((u - 1) & 1)], ((u - 1) >> 1) & 1);  // P V(u - 1) has landed
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
      const unsigned long long sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nmb2 = pack_f32x2(-mb, -mb);
      unsigned long long sum2a = 0ull, sum2b = 0ull;   // two packed accumulators (4 partial sums)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const unsigned long long xa = fma_f32x2(pack_f32x2(__uint_as_float(sr[q * 32 + c]), __uint_as_float(sr[q * 32 + c + 1])), sc2, nmb2);
          const unsigned long long xb = fma_f32x2(pack_f32x2(__uint_as_float(sr[q * 32 + c + 2]), __uint_as_float(sr[q * 32 + c + 3])), sc2, nmb2);
          float a0, a1, b0, b1;
          unpack_f32x2(xa, a0, a1);
          unpack_f32x2(xb, b0, b1);
          a0 = ex2_approx(a0); a1 = ex2_approx(a1); b0 = ex2_approx(b0); b1 = ex2_approx(b1);
          sum2a = add_f32x2(sum2a, pack_f32x2(a0, a1));
          sum2b = add_f32x2(sum2b, pack_f32x2(b0, b1));
          pk[c >> 1] = pack_bf16(a0, a1);
          pk[(c >> 1) + 1] = pack_bf16(b0, b1);
        }
        tmem_st_32x16(tS + q * 16, pk);   // P(u): 32 packed columns over the first half of S(u) (already in registers)
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
      if (lane == 0) mbar_arrive(&p_full[g * 2 + (u & 1)]);
      if (u + 1 < Bg) load_mask(b == nb - 1 ? itl + 1 : itl, b == nb - 1 ? 0 : b + 1);   // mask bytes of the next block: a whole block to arrive
      if (b == nb - 1) {   // the item ends here: retire it (its last P V is the next thing on the pipe; the next item's scores are already there)
        mbar_wait(&o_full[g * 2 + (u & 1)], (u >> 1) & 1);
        tc_fence_after();
        stage_item(g * 2 + (itl & 1), l_run);
      }
      if (++b == nb) { b = 0; ++itl; }
    }
  }
#undef A7_ITEM
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem0);
}

inline int& opt_attn7() {   // attention kernel generation 7 (this file): 64-key blocks, two S buffers per group.  Measured (call 25, parity-green): self 22.2 vs
                            // 21.3 us, 30-s shape 82.4 vs 73.0 us, cross 9.6 vs 9.6 us, step 7.38 vs 7.31-7.56 ms: twice as many hand-offs and tcgen05.ld / st
                            // round trips per key cost more than the early S buys -- off.
  static int v = [] { const char* e = getenv("EZB_ATTN7"); return e ? atoi(e) : 0; }();
  return v;
}

inline int attention_tc7(Device& dev, cudaStream_t st, const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, const uint8_t* key_mask,
                         __nv_bfloat16* out, int B, int H, int Lq, int Lk, int Lkpad, int dh, int dhp, int dvp, float scale) {
  if (!((dh == 64 && dhp == 64 && dvp == 64) || (dh == 72 && (dhp == 128 || dhp == 80) && dvp == 80))) return fail(EZB_ERR_UNSUPPORTED, "attention_tc7: dh %d dhp %d dvp %d", dh, dhp, dvp);
  // a single 64-key block per item would need three Q buffers per group (scores are issued two blocks = two items ahead): generation 6 takes those
  if (Lk <= 64) return attention_tc6(dev, st, q, k, vt, key_mask, out, B, H, Lq, Lk, Lkpad, dh, dhp, dvp, scale);
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
  if (dh == 64) return go(attn7_kernel<64>, Attn7Smem<64>::total(dvp));
  return go(attn7_kernel<72>, Attn7Smem<72>::total(dvp));
}

}  // namespace ezb
