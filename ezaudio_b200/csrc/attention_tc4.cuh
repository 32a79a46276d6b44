// tcgen05 flash attention, two query tiles in flight per CTA (ping-pong): O = softmax(Q K^T / sqrt(dh) [+ key mask]) V.
// Replaces F.scaled_dot_product_attention in src/models/utils/attention.py:107-110 (self: mask None; cross: bool key mask built
// by attention.py:30-37).
//
// Why two tiles: one 128 x 128 score block costs 16 K exp2 on the MUFU (1024 cycles per SM) but only ~640 tensor-pipe cycles, and every
// softmax pass has long latency stretches (TMEM load, row max, TMEM store, barrier hand-offs).  With a single tile per CTA those stretches
// leave the MUFU idle; with two independent tiles one group's exponentials cover the other's latencies.
//
// Persistent CTA of 10 warps; its work items (b*H + h, 128-query tile) alternate between the two softmax groups:
//   warps 0-3 / 4-7 : softmax group 0 / 1.  Thread = one query row and ALL 128 key columns of the block (no cross-thread exchange):
//                     tcgen05.ld S -> row max -> exp2 -> P (bf16) written back into tensor memory over the first half of S
//                     (tcgen05.st) -> arrive.  O accumulates in tensor memory across key blocks (tcgen05.mma accumulate); it is rescaled
//                     in place (ld / mul / st) only when a row's running max grew by more than 2^8 since the reference max was taken
//                     (P <= 256 is exact in bf16's exponent range, the final O / l is independent of the reference).
//   warp 8 (MMA)    : S_g = Q_g K^T (SS MMA, M128 N128) and O_g (+)= P_g V (TS MMA: A operand = P in tensor memory, B = V^T tile in
//                     shared memory), issued in a static ping-pong order  S_A0 S_B0 | PV_A0 S_A1 | PV_B0 S_B1 | ...
//                     (S_{u+1} of a group aliases P_u: it is issued after PV_u, the tensor pipe executes in issue order).
//   warp 9 (TMA)    : Q tile per item and group; K and V^T tiles through two 4-deep rings filled in the same static order.
// Layouts as produced by the QKV GEMM epilogue: Q,K [B*H, L, DHP] bf16 (dh = 72: 64 columns SWIZZLE_128B + a 16-column SWIZZLE_32B
// tail box); V^T [B*H, DVP, Lpad] bf16.  Output [B, Lq, H*dh] bf16 token-major.
#pragma once
#include "gemm.cuh"
#include "host.cuh"

namespace ezb {

constexpr int A4_SOFTMAX_THREADS = 256;
constexpr int A4_THREADS = A4_SOFTMAX_THREADS + 64;
constexpr int A4_STAGES = 4;

struct Attn4Params {
  const uint8_t* key_mask;  // [B, Lk] or null
  __nv_bfloat16* out;       // [B, Lq, H*dh]
  int H, Lq, Lk, dvp;
  int n_qt, n_items;
  float scale_log2;         // (1/sqrt(dh)) * log2(e)
  int pp;                   // 1: the two softmax groups take turns in their exponent phases (MUFU token, see the kernel comment)
  int dbg;                  // profiling only (option "attn_dbg", DBG instantiation; results are garbage): 1 no exp2, 2 no S load, 4 no P store,
                            // 8 no P V MMAs, 16 no S MMAs
  unsigned long long* dbg_buf;  // CTA 0 cycle counters: [0] softmax g0 wait S, [1] softmax g0 item write-out, [2] softmax g0 loop total,
                                // [3] MMA wait P, [4] MMA wait V, [5] MMA wait Q/K, [6] TMA wait empty slots, [7] MMA thread total
};

// 2^x on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial on [-0.5, 0.5], max relative error 7.8e-5 -- far below the bf16
// rounding of P): used for one element in four so that the MUFU (16 exp2 / clock / SM) is not the only unit working through the scores.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;          // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);    // [-0.5, 0.5]
  float q = fmaf(0.05508868f, f, 0.24260405f);
  q = fmaf(q, f, 0.69327623f);
  q = fmaf(q, f, 0.99992895f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}

template <int DH>
struct Attn4Smem {
  static constexpr int TAIL = DH > 64 ? 4096 : 0;
  static constexpr int Q_BYTES = 16384 + TAIL;
  static constexpr int K_BYTES = 16384 + TAIL;
  static __host__ __device__ constexpr int v_bytes(int dvp) { return 2 * dvp * 128; }
  static __host__ __device__ constexpr int total(int dvp) { return 1024 + 2 * Q_BYTES + A4_STAGES * K_BYTES + A4_STAGES * v_bytes(dvp) + 512; }
};

// RES = 1 ("K/V resident"): one CTA per (b, h); its K and V^T key blocks (at most A4_STAGES = 4, i.e. Lk <= 512) are loaded ONCE and stay in
// the ring slots while the CTA walks all query tiles of that head.  The MMA thread then has no k_full / v_full waits (after the first pass) and
// no k_empty / v_empty commits per score block -- it was the longest pole (32 K of 48 K cycles, ~1150 cycles of bookkeeping per block) --
// every CTA has the same number of blocks (no 4-items-vs-3 tail) and K / V^T are read from L2 once instead of once per query tile.
template <int DH, int POLY, int DBG = 0, int RES = 0>
__global__ void __launch_bounds__(A4_THREADS, 1)
attn4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
             const __grid_constant__ CUtensorMap tmQt, const __grid_constant__ CUtensorMap tmKt, const Attn4Params p) {
  using SM = Attn4Smem<DH>;
  constexpr bool HAS_TAIL = DH > 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int VB = SM::v_bytes(p.dvp);
  uint8_t* sQ = smem;                              // [2 groups][Q_BYTES]
  uint8_t* sK = sQ + 2 * SM::Q_BYTES;              // [STAGES][K_BYTES]
  uint8_t* sV = sK + A4_STAGES * SM::K_BYTES;      // [STAGES][VB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A4_STAGES * VB);
  uint64_t *q_full = bars, *q_empty = bars + 2, *s_full = bars + 4, *p_full = bars + 6, *o_full = bars + 8;
  uint64_t *k_full = bars + 10, *k_empty = k_full + A4_STAGES, *v_full = k_empty + A4_STAGES, *v_empty = v_full + A4_STAGES;
  uint64_t* tok = v_empty + A4_STAGES;   // [2] MUFU token: tok[g] completes a phase when the OTHER group has finished an exponent phase
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tok + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.Lk + 127) / 128;
  const int my_items = RES ? p.n_qt : (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
#define A4_ITEM(itl_, g_) (RES ? (int)blockIdx.x * p.n_qt + 2 * (itl_) + (g_) : (int)blockIdx.x + (2 * (itl_) + (g_)) * (int)gridDim.x)
  const int U0 = ((my_items + 1) >> 1) * n_kv, U1 = (my_items >> 1) * n_kv;   // units of softmax group 0 / 1
  const int dbg = DBG ? p.dbg : 0;   // compile-time zero in the production instantiation: the skip branches and counters vanish
  const bool cnt = DBG && p.dbg_buf != nullptr && blockIdx.x == 0;
  long long c_a = 0, c_b = 0, c_c = 0, t_begin = 0;
#define A4_TIMED(acc, stmt) do { if (cnt) { const long long _t = clock64(); stmt; acc += clock64() - _t; } else { stmt; } } while (0)

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      if (HAS_TAIL) { tma_prefetch_desc(&tmQt); tma_prefetch_desc(&tmKt); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&o_full[i], 1);
        mbar_init(&p_full[i], A4_SOFTMAX_THREADS / 2 / 32);   // one elected arrival per softmax warp (128 per-thread arrivals on one mbarrier serialise)
        mbar_init(&tok[i], A4_SOFTMAX_THREADS / 2 / 32);
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

  if (warp == 9) {
    // ------------------------------------------------ TMA producer (static ping-pong order).  Warp-uniform loop, copies issued under
    // elect.sync (inside an `if (lane == 0)` region ptxas serialises every UTMALDG / UTCHMMA / UTCBAR through an ELECT ... BRA.U.ANY loop).
    {
      int kc = 0, vc = 0;
      const int maxU = U0 > U1 ? U0 : U1;
      int itl = 0, j = 0;   // s = itl * n_kv + j
      for (int s = 0; s < maxU; ++s) {
        for (int g = 0; g < 2; ++g) {
          if (s >= (g ? U1 : U0)) continue;
          const int item = A4_ITEM(itl, g);
          const int bh = item / p.n_qt, q0 = (item - bh * p.n_qt) * 128;
          if (j == 0) {
            A4_TIMED(c_a, mbar_wait(&q_empty[g], (itl & 1) ^ 1));
            if (elect_one()) {
              mbar_expect_tx(&q_full[g], SM::Q_BYTES);
              tma_load_3d(sQ + g * SM::Q_BYTES, &tmQ, &q_full[g], 0, q0, bh);
              if (HAS_TAIL) tma_load_3d(sQ + g * SM::Q_BYTES + 16384, &tmQt, &q_full[g], 64, q0, bh);
            }
            __syncwarp();
          }
          if (RES) {   // K block j lives in slot j for the whole CTA: loaded by the first item that needs it (group 0's first query tile)
            if (itl == 0 && g == 0) {
              if (elect_one()) {
                mbar_expect_tx(&k_full[j], SM::K_BYTES);
                tma_load_3d(sK + j * SM::K_BYTES, &tmK, &k_full[j], 0, j * 128, bh);
                if (HAS_TAIL) tma_load_3d(sK + j * SM::K_BYTES + 16384, &tmKt, &k_full[j], 64, j * 128, bh);
              }
              __syncwarp();
            }
            continue;
          }
          const int st = kc % A4_STAGES;
          A4_TIMED(c_a, mbar_wait(&k_empty[st], ((kc / A4_STAGES) & 1) ^ 1));
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
          const int item = A4_ITEM(itl, g);
          const int bh = item / p.n_qt;
          if (RES) {
            if (itl == 0 && g == 0) {
              if (elect_one()) {
                mbar_expect_tx(&v_full[j], VB);
                for (int hh = 0; hh < 2; ++hh) tma_load_3d(sV + j * VB + hh * (VB / 2), &tmV, &v_full[j], j * 128 + hh * 64, 0, bh);
              }
              __syncwarp();
            }
            continue;
          }
          const int st = vc % A4_STAGES;
          A4_TIMED(c_a, mbar_wait(&v_empty[st], ((vc / A4_STAGES) & 1) ^ 1));
          if (elect_one()) {
            mbar_expect_tx(&v_full[st], VB);
            for (int hh = 0; hh < 2; ++hh) tma_load_3d(sV + st * VB + hh * (VB / 2), &tmV, &v_full[st], j * 128 + hh * 64, 0, bh);
          }
          __syncwarp();
          ++vc;
        }
        if (++j == n_kv) { j = 0; ++itl; }
      }
      if (cnt && lane == 0) p.dbg_buf[6] = (unsigned long long)c_a;
    }
  } else if (warp == 8) {
    // ------------------------------------------------ MMA issuer: warp-uniform control flow, one elected lane issues
    {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128), idesc_o = umma_idesc_bf16(128, p.dvp);
      int kc = 0, vc = 0;
      unsigned kseen = 0, vseen = 0;   // RES: resident key blocks whose arrival this warp has already observed
      int sj[2] = {0, 0}, sit[2] = {0, 0};   // next score block to issue per group: key block j, local item index
      int pj[2] = {0, 0};                    // next P V per group: key block j
      if (cnt) t_begin = clock64();
      auto issue_s = [&](int g, int s) {
        const int itl = sit[g], j = sj[g];
        if (j == 0) A4_TIMED(c_c, mbar_wait(&q_full[g], itl & 1));
        const int st = RES ? j : kc % A4_STAGES;
        if (RES) {
          if (!((kseen >> j) & 1)) { mbar_wait(&k_full[j], 0); kseen |= 1 << j; }
        } else
        A4_TIMED(c_c, mbar_wait(&k_full[st], (kc / A4_STAGES) & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint64_t qd = umma_desc_sw128(smem_u32(sQ + g * SM::Q_BYTES)), kd = umma_desc_sw128(smem_u32(sK + st * SM::K_BYTES));
          if (!(dbg & 16)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem0 + g * 128, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
          if (HAS_TAIL)
            umma_bf16(tmem0 + g * 128, umma_desc_sw32(smem_u32(sQ + g * SM::Q_BYTES + 16384)), umma_desc_sw32(smem_u32(sK + st * SM::K_BYTES + 16384)), idesc_s, 1);
          }
          if (!RES) umma_commit(&k_empty[st]);
          umma_commit(&s_full[g]);
          if (j == n_kv - 1) umma_commit(&q_empty[g]);
        }
        __syncwarp();
        ++kc;
        if (++sj[g] == n_kv) { sj[g] = 0; ++sit[g]; }
      };
      auto issue_pv = [&](int g, int s) {
        const int j = pj[g];
        A4_TIMED(c_a, mbar_wait(&p_full[g], s & 1));
        const int st = RES ? j : vc % A4_STAGES;
        if (RES) {
          if (!((vseen >> j) & 1)) { mbar_wait(&v_full[j], 0); vseen |= 1 << j; }
        } else
        A4_TIMED(c_b, mbar_wait(&v_full[st], (vc / A4_STAGES) & 1));
        tc_fence_after();
        if (elect_one()) {
          for (int hh = 0; hh < 2 && !(dbg & 8); ++hh) {
            const uint64_t vd = umma_desc_sw128(smem_u32(sV + st * VB + hh * (VB / 2)));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ts(tmem0 + 256 + g * 128, tmem0 + g * 128 + hh * 32 + k * 8, vd + 2 * k, idesc_o, (j != 0) || ((hh | k) != 0));
          }
          if (!RES) umma_commit(&v_empty[st]);
          umma_commit(&o_full[g]);
        }
        __syncwarp();
        ++vc;
        if (++pj[g] == n_kv) pj[g] = 0;
      };
      const int maxU = U0 > U1 ? U0 : U1;
      if (U0 > 0) issue_s(0, 0);
      if (U1 > 0) issue_s(1, 0);
      for (int s = 0; s < maxU; ++s) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int Ug = g ? U1 : U0;
          if (s >= Ug) continue;
          issue_pv(g, s);
          if (s + 1 < Ug) issue_s(g, s + 1);
        }
      }
      if (cnt && warp == 8 && lane == 0) { p.dbg_buf[3] = (unsigned long long)c_a; p.dbg_buf[4] = (unsigned long long)c_b; p.dbg_buf[5] = (unsigned long long)c_c;
                 p.dbg_buf[7] = (unsigned long long)(clock64() - t_begin); }
    }
  } else {
    // ------------------------------------------------ softmax groups
    const int g = warp >> 2, lg = warp & 3;
    const int r = lg * 32 + lane;
    const uint32_t t_row = static_cast<uint32_t>(lg * 32) << 16;
    const uint32_t tS = tmem0 + g * 128 + t_row, tO = tmem0 + 256 + g * 128 + t_row;
    const int Ug = g ? U1 : U0;
    float m_ref = -INFINITY, l_run = 0.f;

    auto write_item = [&](int item, float l_fin) {  // O / l of a finished item -> global (its last P V has completed)
      uint32_t orr[64];
      uint32_t o8[8];
      tmem_ld_32x64(tO, orr);
      if (DH > 64) tmem_ld_32x8(tO + 64, o8);
      tmem_ld_wait();
      tc_fence_before();
      const float inv = 1.f / l_fin;
      const int bh = item / p.n_qt, b = bh / p.H, h = bh - b * p.H;
      const int qrow = (item - bh * p.n_qt) * 128 + r;
      if (qrow < p.Lq) {
        uint4* orow = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.Lq + qrow) * (size_t)(p.H * DH) + h * DH);
#pragma unroll
        for (int v = 0; v < 8; ++v)
          orow[v] = make_uint4(pack_bf16(__uint_as_float(orr[8 * v]) * inv, __uint_as_float(orr[8 * v + 1]) * inv),
                               pack_bf16(__uint_as_float(orr[8 * v + 2]) * inv, __uint_as_float(orr[8 * v + 3]) * inv),
                               pack_bf16(__uint_as_float(orr[8 * v + 4]) * inv, __uint_as_float(orr[8 * v + 5]) * inv),
                               pack_bf16(__uint_as_float(orr[8 * v + 6]) * inv, __uint_as_float(orr[8 * v + 7]) * inv));
        if (DH > 64)
          orow[8] = make_uint4(pack_bf16(__uint_as_float(o8[0]) * inv, __uint_as_float(o8[1]) * inv), pack_bf16(__uint_as_float(o8[2]) * inv, __uint_as_float(o8[3]) * inv),
                               pack_bf16(__uint_as_float(o8[4]) * inv, __uint_as_float(o8[5]) * inv), pack_bf16(__uint_as_float(o8[6]) * inv, __uint_as_float(o8[7]) * inv));
      }
    };

    if (cnt) t_begin = clock64();
    int itl = 0, j = 0;   // s = itl * n_kv + j
    for (int s = 0; s < Ug; ++s) {
      const int item = A4_ITEM(itl, g);
      const int bh = item / p.n_qt, b = bh / p.H;
      A4_TIMED(c_a, mbar_wait(&s_full[g], s & 1));
      tc_fence_after();
      uint32_t sr[128];
      if (dbg & 2) {
#pragma unroll
        for (int q = 0; q < 128; ++q) sr[q] = 0;
      } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld_32x32(tS + q * 32, sr + q * 32);
      tmem_ld_wait();
      }
      const int kbase = j * 128;
      const bool full = (p.key_mask == nullptr) && (kbase + 128 <= p.Lk);
      if (!full) {  // one validity bit per key column, identical for every row: 4 words built with warp ballots
        const uint8_t* km = p.key_mask ? p.key_mask + (size_t)b * p.Lk : nullptr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
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
      for (int c = 0; c < 128; c += 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(sr[c + e]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // reference max: fresh for the first key block of an item, afterwards only moved when the row max outgrew it by 2^8
      float fac = 1.f;
      bool need = false;
      const float l_prev = l_run;  // the previous item's sum (retired below when j == 0)
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
      uint32_t pk[64];
      // MUFU token (p.pp): exponent phases strictly alternate  G0#0 G1#0 G0#1 G1#1 ...  Left alone the two groups drift INTO phase (the MMA warp serves
      // them back to back): both then share the 16 ex2 / clock of the SM during their exponent phases and both wait for the tensor pipe afterwards
      // -- ncu: XU 38 % busy, 31 % of the softmax warps' samples on the S wait.  With the token one group's exponentials run at the full MUFU rate
      // while the other group's P V / next S MMAs, tcgen05.ld and row max are in flight.
      if (p.pp && !(g == 0 && s == 0)) mbar_wait(&tok[g], (g == 0 ? s - 1 : s) & 1);
      if (dbg & 1) {
#pragma unroll
        for (int c = 0; c < 128; c += 2) {
          const float p0 = fmaf(__uint_as_float(sr[c]), p.scale_log2, -mb), p1 = fmaf(__uint_as_float(sr[c + 1]), p.scale_log2, -mb);
          sum4[(c >> 1) & 3] += p0 + p1;
          pk[c >> 1] = pack_bf16(p0, p1);
        }
      } else {
#pragma unroll
      for (int c = 0; c < 128; c += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(sr[c]), p.scale_log2, -mb));
        const float x1 = fmaf(__uint_as_float(sr[c + 1]), p.scale_log2, -mb);
        const float p1 = (POLY && (c & 2)) ? ex2_poly(x1) : ex2_approx(x1);
        sum4[(c >> 1) & 3] += p0 + p1;
        pk[c >> 1] = pack_bf16(p0, p1);
      }
      }
      l_run += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      if (p.pp) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&tok[g ^ 1]);
      }
      if (!(dbg & 4)) {
      tmem_st_32x32(tS, pk);
      tmem_st_32x32(tS + 32, pk + 32);
      }
      if (s > 0 && j == 0) {  // retire the previous item while this unit's S / P hand-off is in flight: its last P V wrote the O that this
        A4_TIMED(c_b, { mbar_wait(&o_full[g], (s - 1) & 1);   // unit's P V (accumulate = 0, issued after our arrive) will overwrite
        tc_fence_after();
        write_item(A4_ITEM(itl - 1, g), l_prev); });
      }
      // in-place rescale of the O rows whose reference max moved (warp-collective TMEM access: every lane takes part)
      if (j != 0 && __any_sync(0xffffffffu, need)) {
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
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      if (++j == n_kv) { j = 0; ++itl; }
    }
    if (Ug > 0) {  // last item of this group
      A4_TIMED(c_b, { mbar_wait(&o_full[g], (Ug - 1) & 1);
      tc_fence_after();
      write_item(A4_ITEM((Ug - 1) / n_kv, g), l_run); });
    }
    if (p.pp && g == 1) {   // group 0 has n_kv more blocks than group 1 when the CTA's item count is odd: keep handing the token back
      for (int n = U1; n < U0 - 1; ++n) {
        mbar_wait(&tok[1], n & 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(&tok[0]);
      }
    }
    if (cnt && warp == 0 && lane == 0) { p.dbg_buf[0] = (unsigned long long)c_a; p.dbg_buf[1] = (unsigned long long)c_b; p.dbg_buf[2] = (unsigned long long)(clock64() - t_begin); }
  }
#undef A4_TIMED
#undef A4_ITEM
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem0);
}

inline int attention_tc4(Device& dev, cudaStream_t st, const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, const uint8_t* key_mask,
                         __nv_bfloat16* out, int B, int H, int Lq, int Lk, int Lkpad, int dh, int dhp, int dvp, float scale) {
  if (!((dh == 64 && dhp == 64 && dvp == 64) || (dh == 72 && (dhp == 128 || dhp == 80) && dvp == 80))) return fail(EZB_ERR_UNSUPPORTED, "attention_tc4: dh %d dhp %d dvp %d", dh, dhp, dvp);
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
  p.pp = opt_attn_pp() != 0;
  p.dbg = opt_attn_dbg() & 31;
  p.dbg_buf = (opt_attn_dbg() & 32) ? gemm_dbg_buf() : nullptr;
  const int grid = p.n_items < dev.num_sms ? p.n_items : dev.num_sms;
  auto go = [&](auto kern, int smem) -> int {
    EZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return launch_k(kern, dim3(grid), dim3(A4_THREADS), smem, st, 1, *tq, *tk, *tv, *tqt, *tkt, p);
  };
  const bool poly = opt_attn_poly() != 0;
  if (opt_attn_dbg() != 0 && dh == 72) return go(attn4_kernel<72, 0, 1>, Attn4Smem<72>::total(dvp));   // profiling instantiation
  if (opt_attn_res() && (Lk + 127) / 128 <= A4_STAGES && !opt_attn_dbg()) {   // K / V^T resident: one CTA per (b, h), all its query tiles
    auto go_res = [&](auto kern, int smem) -> int {
      EZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      return launch_k(kern, dim3(B * H), dim3(A4_THREADS), smem, st, 1, *tq, *tk, *tv, *tqt, *tkt, p);
    };
    if (dh == 64) return go_res(attn4_kernel<64, 0, 0, 1>, Attn4Smem<64>::total(dvp));
    return go_res(attn4_kernel<72, 0, 0, 1>, Attn4Smem<72>::total(dvp));
  }
  if (dh == 64) {
    EZB_TRY(poly ? go(attn4_kernel<64, 1>, Attn4Smem<64>::total(dvp)) : go(attn4_kernel<64, 0>, Attn4Smem<64>::total(dvp)));
  } else {
    EZB_TRY(poly ? go(attn4_kernel<72, 1>, Attn4Smem<72>::total(dvp)) : go(attn4_kernel<72, 0>, Attn4Smem<72>::total(dvp)));
  }
  return EZB_OK;
}

}  // namespace ezb
