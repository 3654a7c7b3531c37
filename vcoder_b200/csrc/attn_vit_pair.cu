// CLIP ViT self-attention on tcgen05 + TMEM, "pair" schedule with half-block softmax:  O = softmax(Q K^T * scale) V,
// head_dim 64.  Reference: CLIPVisionTower.forward -> transformers/models/clip/modeling_clip.py:261-345 (eager attention,
// no mask); SURVEY K4 (4,726,784 algorithmic bytes, 1.364 GFLOP per image-layer at S=577).
//
// One PERSISTENT CTA per SM walks work items; an item is TWO 128-query tiles (A, B) of one (image, head) — K/V tiles are
// fetched once for both — or the single last tile of an odd tile count.  TMEM, barriers and the K/V rings live across items,
// so the next item's Q / K / V are in flight and its first QK^T is issued while the current item drains: per item
// 23.7 k cycles instead of 25.9 k with one CTA per item (4.94 -> 4.60 us per image-layer; profiles/r02_vit_pair_trace_persistent.txt).
//   warp 10 / lane 0: TMA producer  (Q_A, Q_B, 2-deep K ring, 2-deep V ring; 128B-swizzled boxes of the packed QKV buffer)
//   warps 8, 9 / lane 0 : MMA issuer of tile A / B:  S_t = Q_t K_j^T (128x128x64, SS) -> TMEM S_t ;  O_t += P_t,n V_j[half]
//                                   (128x64x64, SS, V MN-major)
//   warps 0..3      : softmax warpgroup of tile A        warps 4..7 : softmax warpgroup of tile B
// The limit of this kernel is the MUFU (16 ex2 per clock per SM, measured: profiles/r02_microbench_tmem_ex2.jsonl; the
// TMEM read port delivers ~960 B/clk/SM and is not).  Round 1 kept a whole 128-score row per thread in registers: with
// 168 registers per thread (65536 / 384) ptxas spilled 60 words into local memory INSIDE the exponential loop, and a
// block's chain (128 exponentials -> P -> PV -> barrier) ran at 4200 cycles against the 2048 the MUFU needs for the two
// tiles (clock64 trace, profiles/r02_vit_pair_trace.txt).  Now a thread handles a block as two 64-key HALVES: 64 scores in
// registers (no spills), each half has its own P buffer, its own PV MMA (four k-steps) and its own completion barrier, so
// the exponentials of half n+1 run while the PV of half n is in flight, and O is only touched by the CUDA cores when the
// running row max grew by more than 2^8 (lazy rescale, then PV(n-1) is awaited first).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kQ = 128;              // query rows per tile
constexpr int kKB = 128;             // keys per K/V tile
constexpr int kHalf = 64;            // keys per softmax / PV half
constexpr int kDh = 64;
constexpr int kTile = kQ * kDh * 2;  // 16 KiB: Q tile, K tile, V tile, one P half [128 rows][64 keys]
constexpr int kThreads = 352;
// The warp scheduler of an SM sub-partition prefers the HIGHEST warp id among its ready warps: the single-thread MMA issuer
// and the TMA producer must not sit below softmax warps that always have an instruction ready (with the issuer as warp 1,
// a PV MMA started 3500-4400 cycles after its P tile was ready: profiles/r02_vit_pair_trace_halfblock.txt)
constexpr int kMmaWarp = 8, kTmaWarp = 10;  // warps 8 / 9: MMA issuer of tile A / B
// smem: Q_A Q_B | K[2] | V[2] | P[tile][2] | barriers
constexpr int kOffK = 2 * kTile, kOffV = 4 * kTile, kOffP = 6 * kTile, kOffBar = 10 * kTile;
constexpr int kSmemBytes = kOffBar + 256;  // 164,096 B (26 barriers + the TMEM slot fit in the 256-byte tail)
constexpr uint32_t kColS = 0, kColO = 256;  // TMEM: S_A 0, S_B 128, O_A 256, O_B 320
constexpr float kRescaleLog2 = 8.0f;

struct VitPairParams {
  __half* out;
  long out_stride;
  int S, D, n_heads, batch;
  float scale_log2;
  uint32_t v_desc_hi, v_lbo;
  int turnstile;
  long long* dbg;  // optional clock64 trace of CTA (0,0,0) (tools/trace_vit_pair.py; VCB_VIT_TRACE=<device address>)
};

enum {
  kBarQ = 0,        // +t
  kBarKFull = 2, kBarKEmpty = 4, kBarVFull = 6, kBarVEmpty = 8,
  kBarS = 10,       // +t
  kBarSFree = 12,   // +t          (4 warps)
  kBarP = 14,       // +2*t + buf  (4 warps)
  kBarO = 18,       // +2*t + buf
  kBarQEmpty = 22,  // +t          committed behind an item's last QK^T: the Q tile may be overwritten
  kBarOFree = 24,   // +t          (4 warps) the epilogue has read O: the next item's first PV may overwrite it
  kNumBars = 26
};

// PERSISTENT: a CTA walks the work items {blockIdx.x, blockIdx.x + gridDim.x, ...}; an item is one (image, head, pair of
// query tiles), all two-tile items first, then the single-tile items of an odd tile count (the cheaper fillers of the tail).
// TMEM, the barriers and the K/V rings live across items (every barrier phase is derived from a running count), so the
// producer requests the next item's Q / K / V while the softmax warps are still in the last rounds of the current one and
// the issuer starts its first QK^T as soon as the last S tile has been read: the ~4.5 k-cycle prologue of a CTA
// (allocation, barrier init, ~3 k cycles from the first request to the first S tile) is paid once per CTA, not per item.
__global__ void __launch_bounds__(kThreads, 1)
attn_vit_pair_kernel(const __grid_constant__ CUtensorMap tm, VitPairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  const uint32_t bars = base + kOffBar;
  auto bar = [&](int i) -> uint32_t { return bars + 8u * static_cast<uint32_t>(i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kOffBar + 8 * kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[240] = clock64();
  const int S = p.S;
  const int nb = (S + kKB - 1) / kKB;
  const int nk_last = S - (nb - 1) * kKB;                         // keys of the last block
  const int n_halves = 2 * (nb - 1) + (nk_last > kHalf ? 2 : 1);  // softmax / PV halves per tile
  const int q_tiles = (S + kQ - 1) / kQ;
  const int pairs_per = q_tiles / 2;  // two-tile items per (image, head)
  const int BH = p.batch * p.n_heads;
  const int n_pair_items = BH * pairs_per;
  const int n_items = n_pair_items + ((q_tiles & 1) ? BH : 0);
  // item -> image b, head h, first query row of tile A, number of tiles (tile B starts kQ rows behind tile A)
  auto decode = [&](int item, int& b, int& h, int& qa, int& nt) {
    int bh, pr;
    if (item < n_pair_items) { bh = item / pairs_per; pr = item - bh * pairs_per; nt = 2; }
    else { bh = item - n_pair_items; pr = pairs_per; nt = 1; }
    b = bh / p.n_heads;
    h = bh - b * p.n_heads;
    qa = 2 * pr * kQ;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumBars; ++i) {
      const bool four = (i >= kBarSFree && i < kBarSFree + 2) || (i >= kBarP && i < kBarP + 4) || (i >= kBarOFree && i < kBarOFree + 2);
      const bool two = (i >= kBarKEmpty && i < kBarKEmpty + 2) || (i >= kBarVEmpty && i < kBarVEmpty + 2);
      mbar_init(bar(i), four ? 4 : (two ? 2 : 1));  // K / V slots: one commit per tile (a single-tile item commits twice)
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kTmaWarp) {
    if (lane == 0) {
      tma_prefetch_desc(&tm);
      int G = 0;                 // K/V blocks requested so far (ring slot = G & 1)
      int q_loads[2] = {0, 0};   // Q tiles requested so far, per tile
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int b, h, qa, nt;
        decode(item, b, h, qa, nt);
        const int row_base = b * S;
        // request order = order of first use: Q_A, K_0 (-> S_A(0)), Q_B (-> S_B(0)), V_0, K_1, V_1, ...  (a 16 KiB box of
        // 128-byte rows takes ~500 cycles to arrive: profiles/r02_vit_pair_diag.txt)
        auto load_q = [&](int t) {
          if (q_loads[t] > 0) mbar_wait(bar(kBarQEmpty + t), static_cast<uint32_t>((q_loads[t] - 1) & 1));
          ++q_loads[t];
          mbar_expect_tx(bar(kBarQ + t), kTile);
          tma_load_2d(base + t * kTile, &tm, bar(kBarQ + t), h * kDh, row_base + qa + t * kQ);
        };
        load_q(0);
        for (int j = 0; j < nb; ++j, ++G) {
          const int slot = G & 1;
          const uint32_t ph = static_cast<uint32_t>(((G >> 1) - 1) & 1);
          if (G >= 2) mbar_wait(bar(kBarKEmpty + slot), ph);
          mbar_expect_tx(bar(kBarKFull + slot), kTile);
          tma_load_2d(base + kOffK + slot * kTile, &tm, bar(kBarKFull + slot), p.D + h * kDh, row_base + j * kKB);
          if (j == 0 && nt == 2) load_q(1);
          if (G >= 2) mbar_wait(bar(kBarVEmpty + slot), ph);
          mbar_expect_tx(bar(kBarVFull + slot), kTile);
          tma_load_2d(base + kOffV + slot * kTile, &tm, bar(kBarVFull + slot), 2 * p.D + h * kDh, row_base + j * kKB);
        }
      }
    }
  } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
    // One issuing thread PER TILE, each with blocking waits in exactly the order its softmax warpgroup produces the events
    // (P half 0 -> S released -> P half 1): no polling (mbarrier.test_wait costs ~150 cycles a probe) and no head-of-line
    // blocking between the tiles.  The two threads write disjoint TMEM columns; a K / V slot is released when BOTH tiles'
    // MMAs that read it have been committed (those barriers count two arrivals; in a single-tile item tile A commits twice).
    const int t = warp - kMmaWarp;
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(kQ, kKB);
      constexpr uint32_t idesc_o = make_idesc_f16_bmn(kQ, kDh);
      const uint64_t qdesc = make_sw128_kmajor_desc(base + t * kTile);
      const uint64_t kdesc[2] = {make_sw128_kmajor_desc(base + kOffK), make_sw128_kmajor_desc(base + kOffK + kTile)};
      const uint64_t vhi = (static_cast<uint64_t>(p.v_desc_hi) << 32) | (static_cast<uint64_t>(p.v_lbo) << 16);
      const uint64_t vdesc[2] = {vhi | static_cast<uint64_t>(((base + kOffV) & 0x3FFFF) >> 4),
                                 vhi | static_cast<uint64_t>(((base + kOffV + kTile) & 0x3FFFF) >> 4)};
      const uint64_t pdesc[2] = {make_sw128_kmajor_desc(base + kOffP + (2 * t) * kTile),
                                 make_sw128_kmajor_desc(base + kOffP + (2 * t + 1) * kTile)};
      int G0 = 0;                   // K/V block index (per CTA) of the item's first block
      int items_done = 0;           // items this tile has worked on
      uint32_t sfree_waits = 0;     // SFree completions consumed (one per S tile read by the softmax warps, in order)
      uint32_t p_waits[2] = {0, 0}; // P-buffer completions consumed, per buffer
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, G0 += nb) {
        int b, h, qa, nt;
        decode(item, b, h, qa, nt);
        if (t >= nt) break;  // tile B: the single-tile items come last
        const int commits = nt == 1 ? 2 : 1;
        auto issue_s = [&](int j) {  // S_t(j) = Q_t K_j^T
          const int g = G0 + j;
          mbar_wait(bar(kBarKFull + (g & 1)), (g >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma_f16(tmem + kColS + t * kKB, qdesc + 2 * k, kdesc[g & 1] + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(bar(kBarS + t));
          for (int c = 0; c < commits; ++c) umma_commit(bar(kBarKEmpty + (g & 1)));
          if (j == nb - 1) umma_commit(bar(kBarQEmpty + t));  // the item's last read of Q_t
        };
        auto issue_pv = [&](int n) {  // O_t += P_t,n V_j[half]
          const int j = n >> 1, hf = n & 1, g = G0 + j;
          mbar_wait(bar(kBarP + 2 * t + (n & 1)), p_waits[n & 1] & 1);
          ++p_waits[n & 1];
          if (hf == 0) mbar_wait(bar(kBarVFull + (g & 1)), (g >> 1) & 1);
          // the first PV of an item overwrites O: the previous item's epilogue must have read it
          if (n == 0 && items_done > 0) mbar_wait(bar(kBarOFree + t), static_cast<uint32_t>((items_done - 1) & 1));
          tc_fence_after();
          const int nkeys = min(kHalf, S - j * kKB - hf * kHalf);
          const int ksteps = (nkeys + 15) / 16;
          for (int s = 0; s < ksteps; ++s)  // 16 keys per step: +32 B along P's k, +2048 B (16 rows) along V
            umma_f16(tmem + kColO + t * kDh, pdesc[n & 1] + 2 * s, vdesc[g & 1] + 128 * (hf * 4 + s), idesc_o, (n > 0 || s > 0) ? 1u : 0u);
          umma_commit(bar(kBarO + 2 * t + (n & 1)));
          if (hf == 1 || n == n_halves - 1)
            for (int c = 0; c < commits; ++c) umma_commit(bar(kBarVEmpty + (g & 1)));
        };
        mbar_wait(bar(kBarQ + t), static_cast<uint32_t>(items_done & 1));
        if (items_done > 0) {  // the previous item's last S tile has been read
          mbar_wait(bar(kBarSFree + t), sfree_waits & 1);
          ++sfree_waits;
        }
        issue_s(0);
        for (int j = 0; j < nb; ++j) {
          issue_pv(2 * j);
          if (j + 1 < nb) {
            mbar_wait(bar(kBarSFree + t), sfree_waits & 1);  // the softmax warps have read S_t(j) completely
            ++sfree_waits;
            issue_s(j + 1);
          }
          if (2 * j + 1 < n_halves) issue_pv(2 * j + 1);
        }
        ++items_done;
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups: one thread per query row
    const int t = warp >> 2;  // tile A (warps 0..3) or B (warps 4..7)
    const int q = warp & 3;   // TMEM lane quadrant (hardware rule: warp id % 4)
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float sc = p.scale_log2;
    const uint32_t s_addr = tmem + lane_off + kColS + t * kKB;
    const uint32_t o_addr = tmem + lane_off + kColO + t * kDh;
    uint32_t s_waits = 0;         // S completions consumed (every one is waited for, in order)
    uint32_t o_base[2] = {0, 0};  // PV completions per P buffer before the current item (not every one is waited for)
    int it_n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it_n) {
      int b, h, qa, ntiles;
      decode(item, b, h, qa, ntiles);
      if (t >= ntiles) break;  // tile B: the single-tile items come last
      const int row_base = b * S;
      float Mref = -INFINITY, L = 0.f;
      // Turnstile (VCB_VIT_TURNSTILE=0 turns it off; measured 5.01 vs 5.13 us per image-layer): the two softmax warpgroups take turns on the MUFU, so that one tile's TMEM read /
      // row max / barrier waits run under the other tile's exponentials instead of both tiles idling the MUFU together.
      const bool turn = p.turnstile != 0 && ntiles == 2;
      if (turn && t == 1) asm volatile("bar.arrive 2, 256;" ::: "memory");  // tile A goes first
      // trace of the first two items of CTA 0 (the buffer holds >= 512 slots: tools/trace_vit_pair.py)
      const bool tr = p.dbg != nullptr && blockIdx.x == 0 && it_n < 2 && q == 0 && lane == 0;
#define VCB_STAMP(k) do { if (tr) p.dbg[it_n * 256 + (t * 10 + n) * 8 + (k)] = clock64(); } while (0)

      for (int n = 0; n < n_halves; ++n) {
        const int j = n >> 1, hf = n & 1, buf = n & 1;
        const int nk = min(kHalf, S - j * kKB - hf * kHalf);  // valid keys of this half (>= 1)
        VCB_STAMP(0);
        if (hf == 0) {
          mbar_wait(bar(kBarS + t), s_waits & 1);
          ++s_waits;
          tc_fence_after();
        }
        VCB_STAMP(1);
        uint32_t v[2][32];
        tmem_ld_32x32(s_addr + hf * kHalf, v[0]);
        tmem_ld_32x32(s_addr + hf * kHalf + 32, v[1]);
        tmem_ld_wait();
        if (hf == 1 || n == n_halves - 1) {  // the whole S block has been read: the next QK^T may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kBarSFree + t));
        }
        VCB_STAMP(2);

        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#if defined(VCB_VIT_DIAG) && (VCB_VIT_DIAG & 4)
        if (true) { mx0 = __uint_as_float(v[0][0]); } else
#endif
        if (nk == kHalf) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              mx0 = fmaxf(mx0, __uint_as_float(v[c][i]));
              mx1 = fmaxf(mx1, __uint_as_float(v[c][i + 1]));
              mx2 = fmaxf(mx2, __uint_as_float(v[c][i + 2]));
              mx3 = fmaxf(mx3, __uint_as_float(v[c][i + 3]));
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < nk) mx0 = fmaxf(mx0, __uint_as_float(v[c][i]));
          }
        }
        const float mblk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        VCB_STAMP(3);

        // lazy rescale of O (in TMEM): only when the row max grew by more than 2^kRescaleLog2
        float f = 1.f;
        if (n == 0) {
          Mref = mblk;
        } else if ((mblk - Mref) * sc > kRescaleLog2) {
          f = fast_exp2((Mref - mblk) * sc);
          Mref = mblk;
          L *= f;
        }
        // this half's P buffer was last read by PV(n-2).  Even halves need no barrier of their own: the S tile this half
        // just waited for was committed by the same issuing thread AFTER it issued PV(n-2) (issue order PV(2j), S(j+1),
        // PV(2j+1)), and a tcgen05.commit covers every MMA that thread issued before it.  (A completed mbarrier wait still
        // costs 250-300 cycles here.)  The skipped phases are never waited for later: the rescale path and the epilogue
        // only ever wait for the newest phase of a buffer's barrier.
        if (n >= 2 && hf == 1) {
          mbar_wait(bar(kBarO + 2 * t + buf), (o_base[buf] + static_cast<uint32_t>((n - 2) >> 1)) & 1);
          tc_fence_after();
        }
        VCB_STAMP(4);
        if (n > 0 && __any_sync(0xffffffffu, f != 1.f)) {  // warp-uniform: tcgen05.ld/st are warp-collective
          mbar_wait(bar(kBarO + 2 * t + (buf ^ 1)), (o_base[buf ^ 1] + static_cast<uint32_t>((n - 1) >> 1)) & 1);  // O must hold everything issued so far
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32(o_addr + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
            tmem_st_32x32(o_addr + c * 32, ov);
          }
          tmem_st_wait();
        }
        VCB_STAMP(5);
        if (turn) {
          if (t == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
          else asm volatile("bar.sync 3, 256;" ::: "memory");
        }
        const float msc = Mref * sc;
        uint8_t* prow = smem + kOffP + (2 * t + buf) * kTile + r * 128;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c * 32 < nk) {
            uint32_t pk[16];
            if (nk >= (c + 1) * 32) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
#if defined(VCB_VIT_DIAG) && (VCB_VIT_DIAG & 1)
                const float p0 = fmaf(__uint_as_float(v[c][i]), sc, -msc);
                const float p1 = fmaf(__uint_as_float(v[c][i + 1]), sc, -msc);
#else
                const float p0 = fast_exp2(fmaf(__uint_as_float(v[c][i]), sc, -msc));
                const float p1 = fast_exp2(fmaf(__uint_as_float(v[c][i + 1]), sc, -msc));
#endif
#if !(defined(VCB_VIT_DIAG) && (VCB_VIT_DIAG & 8))
                l0 += p0;
                l1 += p1;
#endif
                pk[i >> 1] = pack_half2(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = (c * 32 + i < nk) ? fast_exp2(fmaf(__uint_as_float(v[c][i]), sc, -msc)) : 0.f;
                const float p1 = (c * 32 + i + 1 < nk) ? fast_exp2(fmaf(__uint_as_float(v[c][i + 1]), sc, -msc)) : 0.f;
                l0 += p0;
                l1 += p1;
                pk[i >> 1] = pack_half2(p0, p1);
              }
            }
            // keys [c*32, c*32+32) of this half: 16-byte chunks c*4 .. c*4+3 of the 128-byte row
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const int chunk = (c * 4 + w) ^ (r & 7);
              *reinterpret_cast<uint4*>(prow + (chunk << 4)) = make_uint4(pk[4 * w], pk[4 * w + 1], pk[4 * w + 2], pk[4 * w + 3]);
            }
          }
        }
        L += l0 + l1;
        if (turn && !(t == 1 && n == n_halves - 1)) {  // hand the MUFU over (B's last hand-over of an item has no taker)
          if (t == 0) asm volatile("bar.arrive 3, 256;" ::: "memory");
          else asm volatile("bar.arrive 2, 256;" ::: "memory");
        }
        VCB_STAMP(6);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBarP + 2 * t + buf));
        VCB_STAMP(7);
      }
#undef VCB_STAMP

      // ---- epilogue: O (TMEM) / l -> fp16.  MMAs complete in issue order: the last half's barrier covers all of them
      {
        const int n = n_halves - 1;
        mbar_wait(bar(kBarO + 2 * t + (n & 1)), (o_base[n & 1] + static_cast<uint32_t>(n >> 1)) & 1);
        tc_fence_after();
      }
      o_base[0] += static_cast<uint32_t>((n_halves + 1) >> 1);
      o_base[1] += static_cast<uint32_t>(n_halves >> 1);
      if (tr) p.dbg[it_n * 256 + 200 + t * 2] = clock64();  // epilogue: O complete
      const int row = qa + t * kQ + r;
      const float inv = 1.f / L;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32(o_addr + c * 32, ov);
        tmem_ld_wait();
        if (c == 1) {  // O is in registers: the next item's first PV may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kBarOFree + t));
        }
        if (row < S) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<long>(row_base + row) * p.out_stride + h * kDh + c * 32);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            dst[w] = make_uint4(
                pack_half2(__uint_as_float(ov[8 * w]) * inv, __uint_as_float(ov[8 * w + 1]) * inv),
                pack_half2(__uint_as_float(ov[8 * w + 2]) * inv, __uint_as_float(ov[8 * w + 3]) * inv),
                pack_half2(__uint_as_float(ov[8 * w + 4]) * inv, __uint_as_float(ov[8 * w + 5]) * inv),
                pack_half2(__uint_as_float(ov[8 * w + 6]) * inv, __uint_as_float(ov[8 * w + 7]) * inv));
          }
        }
      }
      if (tr) p.dbg[it_n * 256 + 201 + t * 2] = clock64();  // epilogue: rows stored
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[241] = clock64();
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int make_tmap_2d_f16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld_elems, uint32_t box_inner,
                     uint32_t box_rows);

int attn_vit_pair(const AttnPrefillParams& a, cudaStream_t stream) {
  VCB_CHECK(a.head_dim == 64, kErrInvalid, "attn_vit_pair: head_dim must be 64");
  VCB_CHECK(a.q_stride == a.kv_stride && a.k == a.q + a.n_heads * 64 && a.v == a.q + 2 * a.n_heads * 64, kErrInvalid,
            "attn_vit_pair: expects the packed [tokens, 3*D] qkv layout");
  static bool attr = false;
  if (!attr) {
    VCB_CUDA(cudaFuncSetAttribute(attn_vit_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr = true;
  }
  const int D = a.n_heads * 64;
  CUtensorMap tm;
  VCB_TRY(make_tmap_2d_f16(&tm, a.q, 3 * D, static_cast<uint64_t>(a.batch) * a.seq_len, a.q_stride, 64, 128));
  VitPairParams p;
  p.out = a.out; p.out_stride = a.out_stride; p.S = a.seq_len; p.D = D; p.n_heads = a.n_heads; p.batch = a.batch;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.v_desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // MN-major V: SBO 1024 B, version 1, SWIZZLE_128B
  p.v_lbo = 0;
  {
    static int ts = -1;
    if (ts < 0) { const char* e = getenv("VCB_VIT_TURNSTILE"); ts = e ? atoi(e) : 1; }
    p.turnstile = ts;
  }
  p.dbg = nullptr;
  if (const char* e = getenv("VCB_VIT_TRACE")) p.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 16));
  const int q_tiles = (a.seq_len + kQ - 1) / kQ;
  const int n_items = a.batch * a.n_heads * ((q_tiles + 1) / 2);
  static int persist = -1;  // VCB_VIT_PERSIST=0: one CTA per item (the kernel's item loop then runs once)
  if (persist < 0) { const char* e = getenv("VCB_VIT_PERSIST"); persist = e ? atoi(e) : 1; }
  dim3 grid(persist ? std::min(n_items, device_sm_count()) : n_items);
  KernelScope scope(kCatAttnVit, stream);
  attn_vit_pair_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tm, p);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace vcb
