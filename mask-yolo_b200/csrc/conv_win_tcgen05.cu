// conv_win_tcgen05.cu -- multi-tap (3x3) convolutions and large plain GEMMs on padded-flat tensors as ONE
// persistent tcgen05 kernel with shared-memory reuse of the A operand across taps.
//
//   C[m, n] = epi( sum_t sum_k A[m + shift_t, k] * Bt[t][n][k] ),   |shift_t| <= 16, N % 128 == 0
//
// Why: with fp32 (tf32) operands a 128x256 output tile needs 48 KB of L2->shared-memory traffic per 2.1 MFLOP,
// and the one-tile-per-CTA kernel (gemm_tcgen05.cu) ran at 38 % tensor-pipe utilisation.  The nine taps of a
// 3x3 conv read the SAME activation rows shifted by at most W+2 rows, so this kernel loads one window of
// 128+32 rows per 32-channel k-block and addresses all nine taps inside it by moving the UMMA descriptor start
// by whole 128-byte rows (measured on B200: the 128B swizzle is a function of the absolute shared-memory
// address, so a row-shifted start needs no base_offset).  Default configuration (template <256, 1, 2>):
//   * CTA pair (cluster of 2, cta_group::2): one 256-row MMA spans both SMs; each CTA holds its own 128 rows
//     (window + TMEM accumulator) and HALF of every weight stage -> L2->SM fill per FLOP halves;
//   * 8-stage weight ring (16 KB per CTA per stage), double-buffered activation window; single-tap GEMMs split the
//     same shared memory four windows / five weight stages (they consume a window per weight stage);
//   * two TMEM stages: the epilogue of item i overlaps the main loop of item i+1;
//   * epilogue through swizzled shared-memory staging + TMA tensor stores (a thread owns one output ROW;
//     storing rows straight to global memory cost as much as the whole main loop).
// Fused epilogues: bias / folded fixed-statistics BN / activation (forward convs); the backward of the previous
// layer's BN + ReLU with its dgamma / dbeta column sums (dgrad, myolo_gemm_taps_bnbwd); the whole mask tail
// (deconv bias + ReLU + 1x1 conv + sigmoid, myolo_deconv_mask_fwd).
//
// Two operand flavours (template parameter EL): 4 = fp32 words read by kind::tf32; 2 = IEEE half on kind::f16 (the
// "h16" precision mode of the mask head: twice the tensor rate, half the operand bytes, fp32 accumulation; the epilogue
// stores half or fp32, and the fused BN backward works in registers with butterfly column sums).
//
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (leader CTA only), then the epilogue warps:
// 4 (EL = 4, 192 threads) or 8 (EL = 2, 320 threads: two per TMEM lane quarter, half of the columns each).
// Persistent: cluster i processes items i, i+n_clusters, ...
// Replaces the Conv2D / Conv2DBackpropInput / Conv2DTranspose call sites of the mask head
// (myolo/model.py:688-713) and the large pointwise GEMMs of the backbone.
// MYOLO_WIN_BO (experiment switches, default 0): 2 no TMA loads, 4 no global stores, 8 128-column slices,
// 16 two accumulators per CTA, 32 single-CTA instead of CTA pairs.
#include "tc_common.cuh"

namespace myolo {
namespace tc {

constexpr int WHALO = 16;
constexpr uint32_t kWStageOut = 2 * 4 * 4096;   // per-epilogue-warp 32x32 fp32 staging buffers: TMA-store tile + BN-backward side tile
constexpr uint32_t kWColAcc = 2 * 256 * 4;       // per-CTA column-sum accumulators of the fused BN backward
constexpr uint32_t kWEpiVec = 2 * 1024 * 4;     // folded epilogue scale / shift, up to 1024 output channels
// threads per CTA: TMA warp + MMA warp + epilogue warps.  The half-operand flavour (EL = 2) runs EIGHT epilogue warps,
// two per TMEM lane quarter, each taking half of the item's columns: its main loop is twice as fast as the tf32 one,
// and with four warps the fused BN-backward epilogue (8.7 k instructions per item) was longer than the main loop.
__host__ __device__ constexpr int win_threads(int el) { return el == 2 ? 320 : 192; }
// NACC = 128-row accumulators per work item.  Window = 128*NACC + 32 rows; weight ring 128 KB (NACC 1) / 96 KB (NACC 2)
__host__ __device__ constexpr int win_rows(int nacc) { return 128 * nacc + 2 * WHALO; }
__host__ __device__ constexpr int win_box(int nacc) { return nacc == 1 ? win_rows(1) : win_rows(2) / 2; }
__host__ __device__ constexpr uint32_t win_ring(int nacc) { return nacc == 1 ? 131072u : 98304u; }
__host__ __device__ constexpr uint32_t win_smem(int nacc) {
  return 2u * win_rows(nacc) * 128u + win_ring(nacc) + kWStageOut + kWColAcc + kWEpiVec + 1024u;
}

// Work item = (256-row tile, WBN-column slice).  Two 128-row accumulators share every weight stage.
//   WBN = 256: 512 TMEM columns, single TMEM stage (epilogue exposed, ~8 % of an item) -- the fastest
//              variant: per 128x256x8 MMA the tensor core reads 12 KB of operands from shared memory
//              in 128 cycles (96 B/clk of the 128 B/clk budget);
//   WBN = 128: two TMEM stages (epilogue overlapped) but 8 KB per 64-cycle MMA = 128 B/clk: measured
//              1.75 ms vs the 256-wide variant on the 4704-ROI mask conv (shared-memory bound).
// Fused tail of the mask head (model.py:711-713) for the deconv GEMM: the 256 columns of an item are
// the output channels of ONE sub-pixel (a,b) of the 2x2 stride-2 transposed conv, so the epilogue can
// finish the network in registers: h = relu(acc + bd), logits = h . w1 + b1, sigmoid, pixel-shuffled
// store of NC floats.  The [rows, 4*256] deconv activation is written only for rows of POSITIVE rois
// (the only rows the backward pass ever reads; nothing is skipped arithmetically).
struct MaskTail {
  const float* bd;    // [256] deconv bias
  const float* w1;    // [256][NC] 1x1 kernel
  const float* b1;    // [NC]
  float* masks;       // [n_roi][2H][2W][NC]; nullptr = ordinary epilogue
  const int* ids;     // [n_roi] target class ids (>0 = positive) or nullptr
  float* y4;          // [rows][4*256] pre-bias deconv output (positive rois only)
  int H, W, NC;
  // tensor-core tail: the 1x1 conv runs as a second tcgen05 GEMM inside the epilogue (any NC <= 128).  h = relu(acc + bd)
  // goes to shared memory as the half A operand [128 rows][256], the 1x1 kernel sits in shared memory as the half B
  // operand [ncp/CG classes][256], and the logits land in the TMEM columns the deconv accumulator has just left.
  int tc, ncp, nawin, nwb;
};

// One 32-column chunk of the mask tail for this thread's row: v = raw deconv accumulators -> (positive rois) y4 store,
// h = relu(v + bd), lg[k][.] += h . w1[:, k].  NCT is a compile-time class count so the 2*NCT accumulation chains of a
// 4-channel group are independent and interleave (with one warp per scheduler the FMA latency is otherwise exposed:
// the two-chain form of this loop ran at 0.18 IPC).
//
// (Tried and rejected, measured on B200: the constants in __constant__ memory with compile-time offsets -- ptxas turns
// them into LDCU.128 + uniform-register FFMA operands, 1.08 -> 1.46 ms; and the eight chunks of a row fully unrolled --
// 38 KB of straight-line code per class count, 1.08 -> 2.1 ms.  The chunk loop stays rolled.)
// Packed fp32 pairs (Blackwell FFMA2 / FADD2: two fp32 operations per instruction, each rounded like its scalar form).
// The scalar FFMA issues every other cycle per scheduler, so the 128 dot-product FMAs of a 32-channel chunk were the
// longest part of this epilogue (profiles/r02_mask_tail_parts.txt: 0.42 ms of the kernel); as 64 FFMA2 they take half.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t p, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// lg[k][p]: a PAIR of partial logits of class k (even / odd channels of the 4-channel groups with parity p); the caller
// adds the four partial sums.
template <int NCT>
__device__ __forceinline__ void mask_tail_chunk(float (&v)[32], uint32_t evs, int C0, float* ypos, uint64_t (&lg)[8][2]) {
  if (ypos) {
    float4* yp = reinterpret_cast<float4*>(ypos + C0);
#pragma unroll
    for (int j = 0; j < 8; ++j) yp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  uint64_t h[16];      // relu(v + bd) as channel pairs
  {
    float4 b4[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b4[j].x), "=f"(b4[j].y), "=f"(b4[j].z), "=f"(b4[j].w) : "r"(evs + 4u * (C0 + 4 * j)));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a0, a1, a2, a3;
      unpack2(add2(pack2(v[4 * j], v[4 * j + 1]), pack2(b4[j].x, b4[j].y)), a0, a1);
      unpack2(add2(pack2(v[4 * j + 2], v[4 * j + 3]), pack2(b4[j].z, b4[j].w)), a2, a3);
      h[2 * j] = pack2(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
      h[2 * j + 1] = pack2(fmaxf(a2, 0.f), fmaxf(a3, 0.f));
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint64_t w[NCT][2];
#pragma unroll
    for (int k = 0; k < NCT; ++k)
      asm("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(w[k][0]), "=l"(w[k][1]) : "r"(evs + 4u * (256 + k * 256 + C0 + 4 * j)));
#pragma unroll
    for (int k = 0; k < NCT; ++k) {
      uint64_t a = lg[k][j & 1];
      a = fma2(h[2 * j], w[k][0], a);
      a = fma2(h[2 * j + 1], w[k][1], a);
      lg[k][j & 1] = a;
    }
  }
}

// columns [cb, ce) of one accumulator row (a multiple of 64 wide); the TMEM load of the next chunk is in flight while
// a chunk is processed
template <int NCT>
__device__ __forceinline__ void mask_tail_row(uint32_t taddr, uint32_t evs, float* ypos, uint64_t (&lg)[8][2], int cb, int ce) {
  float va[32], vb[32];
  tmem_ld32_issue(taddr + (uint32_t)cb, va);
  tmem_ld_wait(va);
#pragma unroll 1
  for (int c0 = cb; c0 < ce; c0 += 64) {
    tmem_ld32_issue(taddr + (uint32_t)(c0 + 32), vb);
    mask_tail_chunk<NCT>(va, evs, c0, ypos, lg);
    tmem_ld_wait(vb);
    if (c0 + 64 < ce) tmem_ld32_issue(taddr + (uint32_t)(c0 + 64), va);
    mask_tail_chunk<NCT>(vb, evs, c0 + 32, ypos, lg);
    if (c0 + 64 < ce) tmem_ld_wait(va);
  }
}

// CG = 2: the work item is shared by a CTA pair (cluster of two, cta_group::2): M = 256 rows per MMA, each
// CTA holds 128 of them (its own activation window and TMEM accumulator) and HALF of every weight stage,
// which halves the L2->SM fill per FLOP -- the limiter of the single-CTA variant (measured: 1.27 ms with
// the loads removed vs 1.60 ms with them; per-SM fill rate 66 B/clk needed vs ~64 B/clk available).
// EL = bytes per operand element: 4 = fp32 words read as tf32 (kind::tf32), 2 = IEEE half (kind::f16, twice the
// tensor rate and half the operand bytes; a k-block is still 128 bytes = 64 channels, so the window / ring /
// descriptor geometry is identical).  With EL = 2 the epilogue can store the result as half (tmCh, 64B-swizzled
// 32x32 tiles) next to or instead of the fp32 tile, and the fused BN backward reads its activation as half.
template <int WBN, int NACC, int CG, int EL = 4>
__global__ void __launch_bounds__(win_threads(EL))
tc_conv_win_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmCh, long long M, int N,
                   int K, int ntaps, TapShifts sh, Epi ep, MaskTail mt, int nitems, int dbg, int nseg, TapShifts seg) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int WBM = 128 * NACC, WROWS = win_rows(NACC), WBOX = win_box(NACC);
  constexpr uint32_t kWinBytes = WROWS * 128;
  constexpr uint32_t kWBBytes = (WBN / CG) * 128;          // this CTA's share of one weight stage
  constexpr int kWBStages = win_ring(NACC) / kWBBytes;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int cid = blockIdx.x / CG, ncl = gridDim.x / CG;   // work is distributed over clusters
  constexpr uint32_t TS = (WBN * NACC <= 256) ? 2 : 1;     // TMEM stages (epilogue overlapped when 2)
  constexpr uint32_t TSTRIDE = WBN * NACC;                  // TMEM columns per stage
  // Ring depths.  A multi-tap conv spends ntaps weight stages per activation window, so two windows and a deep weight
  // ring keep the tensor core fed.  A single-tap GEMM (deconv forward / dgrad, pointwise dgrads) consumes a window
  // per weight stage: with two windows the producer ran only ~0.6 us ahead of the MMAs -- less than one TMA round
  // trip -- and the kernel was latency-bound at 25 % tensor pipe (ncu: epilogue warps parked on t_full).  There the
  // same shared memory is split four windows / five weight stages.
  constexpr int kMaxAWin = NACC == 1 ? 4 : 2;
  const int nawin = mt.tc ? mt.nawin : ((ntaps == 1 && NACC == 1) ? 4 : 2);
  // fused BN backward, half flavour: the activation tiles arrive by TMA into the shared memory of the LAST weight stage
  // (8 epilogue warps x 2 KB).  Read with per-lane global loads (16 bytes per lane at a 512-byte pitch) they slowed the
  // main loop from 0.91 to 1.05 ms per launch although the epilogue warps idle 40 % of the time.
  const bool act_tma = EL == 2 && NACC == 1 && ep.bn_a != nullptr && ep.bn_tma != 0;
  const int nwb = (mt.tc ? mt.nwb
                         : ((ntaps == 1 && NACC == 1) ? (int)((2 * kWinBytes + kWBStages * kWBBytes - 4 * kWinBytes) / kWBBytes) : kWBStages)) -
                  (act_tma ? 1 : 0);
  __shared__ __align__(8) uint64_t bars[2 * kMaxAWin + kWBStages * 2 + 4 + 2 + 8];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t awin0 = base, bst0 = base + (uint32_t)nawin * kWinBytes, stg0 = base + 2 * kWinBytes + kWBStages * kWBBytes;
  // tensor-core mask tail: [windows][weight stages][A2: 128 rows x 256 half, 4 swizzled k-blocks][B2: ncp/CG classes x 256 half][evec]
  const uint32_t a2_0 = bst0 + (uint32_t)nwb * kWBBytes;
  const uint32_t astg0 = a2_0;                                // act_tma: [epilogue warp][32 rows][32 half], 64B swizzle
  const uint32_t b2_0 = a2_0 + 65536u;
  const uint32_t b2_rows = mt.tc ? (uint32_t)(mt.ncp / CG) : 0u;
  float* colacc = reinterpret_cast<float*>(smem_raw + (stg0 + kWStageOut - smem_u32(smem_raw)));          // [2][256]
  float* evec = reinterpret_cast<float*>(smem_raw + ((mt.tc ? b2_0 + b2_rows * 512u : stg0 + kWStageOut + kWColAcc) - smem_u32(smem_raw)));  // [scale N | shift N]
  constexpr int kEpiWarps = win_threads(EL) / 32 - 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KEL = 128 / EL;                             // channels per 128-byte k-block
  const int kblocks = K / KEL;
  const int nh = N / WBN;
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int i) { return bar0 + 8u * i; };
  auto a_empty = [&](int i) { return bar0 + 8u * (kMaxAWin + i); };
  auto b_full = [&](int i) { return bar0 + 8u * (2 * kMaxAWin + i); };
  auto b_empty = [&](int i) { return bar0 + 8u * (2 * kMaxAWin + kWBStages + i); };
  auto t_full = [&](int i) { return bar0 + 8u * (2 * kMaxAWin + 2 * kWBStages + i); };
  auto t_empty = [&](int i) { return bar0 + 8u * (2 * kMaxAWin + 2 * kWBStages + 2 + i); };
  const uint32_t a2_full = bar0 + 8u * (2 * kMaxAWin + 2 * kWBStages + 4);   // leader: every epilogue warp of the pair has written A2
  const uint32_t d2_full = bar0 + 8u * (2 * kMaxAWin + 2 * kWBStages + 5);   // both CTAs: the logits of the item are in TMEM
  auto act_full = [&](int i) { return bar0 + 8u * (2 * kMaxAWin + 2 * kWBStages + 6 + i); };   // per epilogue warp: activation tile landed

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (EL == 2) tma_prefetch_desc(&tmCh);
    for (int i = 0; i < kMaxAWin; ++i) {
      mbar_init(a_full(i), 1);
      mbar_init(a_empty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(t_full(i), 1);
      mbar_init(t_empty(i), kEpiWarps * CG);  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    for (int i = 0; i < kWBStages; ++i) {
      mbar_init(b_full(i), 1);
      mbar_init(b_empty(i), 1);
    }
    mbar_init(a2_full, kEpiWarps * CG);
    mbar_init(d2_full, 1);
    for (int i = 0; i < 8; ++i) mbar_init(act_full(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // epilogue constants, folded:  act((acc + bias) * scale + shift) = act(acc * S + T)
  const float accs = ep.acc_scale ? __ldg(ep.acc_scale) : 1.f;
  if (mt.masks && mt.tc) {   // tensor-core mask tail: evec = [bd 256 | b1 ncp]; B2 = this CTA's classes of w1 as half, K-major
    for (int i = threadIdx.x; i < 256; i += blockDim.x) evec[i] = __ldg(mt.bd + i);
    for (int i = threadIdx.x; i < mt.ncp; i += blockDim.x) evec[256 + i] = i < mt.NC ? __ldg(mt.b1 + i) : 0.f;
    const int rows = (int)b2_rows;
    for (int i = threadIdx.x; i < rows * 256; i += blockDim.x) {
      const int k = i / rows, nl = i - k * rows;           // consecutive threads: consecutive classes of one input channel
      const int cls = (int)rank * rows + nl;
      const float wv = cls < mt.NC ? __ldg(mt.w1 + (size_t)k * mt.NC + cls) : 0.f;
      const uint32_t addr = b2_0 + (uint32_t)(k >> 6) * (uint32_t)rows * 128u + (uint32_t)nl * 128u +
                            (uint32_t)((((k & 63) >> 3) ^ (nl & 7)) << 4) + (uint32_t)(k & 7) * 2u;
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(f2h_sat(wv)) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  } else if (mt.masks) {   // mask tail: evec = [bd 256 | w1 transposed NC x 256]
    for (int i = threadIdx.x; i < 256; i += blockDim.x) evec[i] = __ldg(mt.bd + i);
    for (int i = threadIdx.x; i < 256 * mt.NC; i += blockDim.x) evec[256 + (i % mt.NC) * 256 + i / mt.NC] = __ldg(mt.w1 + i);
  } else if (ep.bn_a) {   // fused BN backward: evec = [gamma*rs | beta | 1/gamma], column accumulators zeroed
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const float ga = __ldg(ep.bn_gamma + n);
      evec[n] = ga * (1.f / sqrtf(__ldg(ep.bn_var + n) + ep.bn_eps));
      evec[N + n] = __ldg(ep.bn_beta + n);
      evec[2 * N + n] = 1.f / (fabsf(ga) < 1e-20f ? copysignf(1e-20f, ga) : ga);
    }
    for (int n = threadIdx.x; n < 512; n += blockDim.x) colacc[n] = 0.f;
  } else {
    // epilogue statistics: one private slot per epilogue warp ([2 statistics][128 columns]) in the 2 KB of colacc plus the
    // 6 KB of evec the folded constants leave free; a column of a slot is only ever touched by one lane of one warp, in
    // item order, so the per-CTA sums are reproducible run to run (shared-memory atomics made myolo_mask_bn1's statistics
    // differ by 1e-5 between two runs, enough to move mask-head gradients by 3e-3)
    if (ep.st_sums)
      for (int n = threadIdx.x; n < 2048; n += blockDim.x) colacc[n < 512 ? n : 512 + n] = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const float sc = ep.scale ? __ldg(ep.scale + n) : 1.f;
      const float b = ep.bias ? __ldg(ep.bias + n) : 0.f;
      evec[n] = sc * accs;
      evec[N + n] = ep.scale ? fmaf(b, sc, __ldg(ep.shift + n)) : b;
    }
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2sm(smem_u32(&tmem_slot), 512);
    else tmem_alloc(smem_u32(&tmem_slot), 512);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ring positions and phases are advanced incrementally: the depths are run-time values, and a division per tap
      // on the single issuing thread costs more than the four MMAs it sits between
      uint32_t ab = 0, aph = 0, s = 0, bph = 0;
      for (int item = cid; item < nitems; item += ncl) {
        const int tile = item / nh, half = item - tile * nh;
        const int row00 = tile * (WBM * CG) + (int)rank * WBM - WHALO;
        // k-segments (nseg > 1, plain GEMMs): C = sum_g A[rows + seg[g]] * Bt[g]; the hi / lo operand pair of a 3xTF32
        // GEMM is three segments over two row ranges of ONE tensor map (seg = {0, lo_off, 0})
        for (int g = 0; g < nseg; ++g) {
        const int row0 = row00 + seg.s[g];
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(a_empty(ab), aph ^ 1u);
          const uint32_t wa = awin0 + ab * kWinBytes;
          if (CG == 2) {
            // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the pair
            if (leader) mbar_expect_tx(a_full(ab), 2 * kWinBytes);
            tma_load_2d_2sm(wa, &tmA, mapa_rank(a_full(ab), 0), kb * KEL, row0);
          } else if (dbg & 2) {
            mbar_arrive(a_full(ab));
          } else {
            mbar_expect_tx(a_full(ab), kWinBytes);
            tma_load_2d(wa, &tmA, a_full(ab), kb * KEL, row0);
            if (NACC == 2) tma_load_2d(wa + WBOX * 128, &tmA, a_full(ab), kb * KEL, row0 + WBOX);
          }
          if (++ab == (uint32_t)nawin) { ab = 0; aph ^= 1u; }
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(b_empty(s), bph ^ 1u);
            if (CG == 2) {
              if (leader) mbar_expect_tx(b_full(s), 2 * kWBBytes);
              tma_load_2d_2sm(bst0 + s * kWBBytes, &tmB, mapa_rank(b_full(s), 0), kb * KEL,
                              (g * ntaps + t) * N + half * WBN + (int)rank * (WBN / 2));
            } else if (dbg & 2) {
              mbar_arrive(b_full(s));
            } else {
              mbar_expect_tx(b_full(s), kWBBytes);
              tma_load_2d(bst0 + s * kWBBytes, &tmB, b_full(s), kb * KEL, (g * ntaps + t) * N + half * WBN);
            }
            if (++s == (uint32_t)nwb) { s = 0; bph ^= 1u; }
          }
        }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = EL == 2 ? make_idesc_f16(128 * CG, WBN, 0, 0) : make_idesc(128 * CG, WBN, 0, 0);
      uint32_t ab = 0, aph = 0, s = 0, bph = 0, it = 0;
      // tensor-core mask tail of item j: logits[256 rows][ncp] = A2 (h as half, written by the epilogue warps of both
      // CTAs) x B2^T, into the first ncp columns of the TMEM stage the epilogue has just drained.  Issued right AFTER the
      // main loop of item j+1 (which never depends on the epilogue of item j), so the tensor pipe runs main(j+1), tail(j),
      // main(j+2), ... back to back while the epilogue warps drain item j+1.  (Issued in the middle of main(j+1) it made
      // t_full(j+1) wait for the drain of item j: 1.04 ms instead of 0.83 ms with the FMA tail.)
      auto issue_tail = [&](uint32_t j) {
        const uint32_t idesc2 = make_idesc_f16(128 * CG, mt.ncp, 0, 0);
        mbar_wait_cluster(a2_full, j & 1u);
        tc_fence_after();
        const uint32_t tacc2 = tmem + (j % TS) * TSTRIDE;
#pragma unroll 1
        for (int kb2 = 0; kb2 < 4; ++kb2) {
          const uint64_t da2 = make_desc(a2_0 + (uint32_t)kb2 * 16384u, 16, 1024);
          const uint64_t db2 = make_desc(b2_0 + (uint32_t)kb2 * b2_rows * 128u, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t accf = (kb2 | k) != 0 ? 1u : 0u;
            if (CG == 2) umma_f16_2sm(tacc2, da2 + (uint64_t)(k * 2), db2 + (uint64_t)(k * 2), idesc2, accf);
            else umma_f16(tacc2, da2 + (uint64_t)(k * 2), db2 + (uint64_t)(k * 2), idesc2, accf);
          }
        }
        if (CG == 2) umma_commit_2sm(d2_full); else umma_commit(d2_full);
      };
      for (int item = cid; item < nitems; item += ncl, ++it) {
        const uint32_t ts = it % TS;
        mbar_wait(t_empty(ts), ((it / TS) & 1u) ^ 1u);  // the epilogue that last used this TMEM stage has drained it
        tc_fence_after();
        const uint32_t tacc = tmem + ts * TSTRIDE;
        for (int gk = 0; gk < nseg * kblocks; ++gk) {
          const int kb = gk;                                    // only its being zero matters below
          mbar_wait(a_full(ab), aph);
          tc_fence_after();
          const uint32_t wa = awin0 + ab * kWinBytes;
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(b_full(s), bph);
            tc_fence_after();
            const uint32_t row = (uint32_t)(WHALO + sh.s[t]);
            const uint64_t db = make_desc(bst0 + s * kWBBytes, 16, 1024);
#pragma unroll
            for (int acc = 0; acc < NACC; ++acc) {
              // row-shifted start inside the swizzled window: the 128B swizzle is a function of the
              // absolute smem address, so no base_offset is needed (verified on B200)
              const uint64_t da = make_desc(wa + (row + 128u * acc) * 128u, 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {   // one MMA per 32 bytes of K: 8 tf32 or 16 half
                const uint32_t accf = (kb | t | k) != 0 ? 1u : 0u;
                const uint32_t td = tacc + (uint32_t)WBN * acc;
                if (EL == 2) {
                  if (CG == 2) umma_f16_2sm(td, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, accf);
                  else umma_f16(td, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, accf);
                } else {
                  if (CG == 2) umma_tf32_2sm(td, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, accf);
                  else umma_tf32(td, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, accf);
                }
              }
            }
            if (CG == 2) umma_commit_2sm(b_empty(s)); else umma_commit(b_empty(s));
            if (++s == (uint32_t)nwb) { s = 0; bph ^= 1u; }
          }
          if (CG == 2) umma_commit_2sm(a_empty(ab)); else umma_commit(a_empty(ab));
          if (++ab == (uint32_t)nawin) { ab = 0; aph ^= 1u; }
        }
        if (CG == 2) umma_commit_2sm(t_full(ts)); else umma_commit(t_full(ts));
        if (mt.tc && it > 0) issue_tail(it - 1);
      }
      if (mt.tc && it > 0) issue_tail(it - 1);
    }
  } else {
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const int ew = warp - 2;                       // epilogue warp index
    const int eh = EL == 2 ? (ew >> 2) : 0;        // EL = 2: which half of the item's columns
    constexpr int kColsPerWarp = EL == 2 ? WBN / 2 : WBN;
    const int cbeg = eh * kColsPerWarp;
    // staging: 4 KB per epilogue warp.  tf32 flavour: warps q = 0..3 own [q*4 KB] plus the side tile 16 KB further up.
    // half flavour (exactly one output): 8 warps x 4 KB = one fp32 tile, or a ring of two 2 KB half tiles so that a
    // chunk does not wait for the previous chunk's TMA store to finish reading shared memory.
    const uint32_t sbuf0 = stg0 + (uint32_t)(EL == 2 ? ew : q) * 4096u;
    const bool st_f32 = !(EL == 2 && ep.no_f32), st_h = EL == 2 && ep.has_h;
    const bool ring_h = EL == 2 && st_h;
    auto wait_store = [&]() {
      if (lane == 0) {
        if (ring_h) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    };
    const uint32_t evs = smem_u32(evec);
    const PfDiv pfd = make_pfdiv(ep.pf_w1, ep.pf_blk);
    const int actk = ep.act & 0xff;
    const bool rnd = (ep.act & MYOLO_ROUND_TF32) != 0;
    uint32_t it = 0, nst = 0, actn = 0;
    const uint32_t astg = astg0 + (uint32_t)ew * 2048u;
    // lane 0: fetch the activation tile of (item_, 32 columns from c0_) for this warp's 32 rows
    auto issue_act = [&](int item_, int c0_) {
      const int tile_ = (item_ / nh) * CG + (int)rank;
      long long r0 = (long long)tile_ * WBM + q * 32;
      if (r0 >= M) r0 = 0;                                  // rows past the end are masked by `valid`; keep the box in range
      mbar_expect_tx(act_full(ew), 2048u);
      tma_load_2d(astg, &tmC, act_full(ew), (item_ % nh) * WBN + c0_, (int)r0);
    };
    if (act_tma && lane == 0 && cid < nitems) issue_act(cid, cbeg);
    for (int item = cid; item < nitems; item += ncl, ++it) {
      const int half = item % nh;
      const int tile = (item / nh) * CG + (int)rank;      // this CTA's 128*NACC-row tile
      const uint32_t ts = it % TS;
      // fused BN backward, half flavour: this thread's first 32 activations of the item are fetched BEFORE the wait for the
      // accumulator (the epilogue warps idle there; the load's DRAM latency used to start after it, once per item), and
      // the rest of the row's 256-byte span is pulled into L2 for the chunk-ahead loads below
      uint4 ahp[4];
      if (EL == 2 && NACC == 1 && ep.bn_a && (dbg & 64)) {
        const long long mp = (long long)tile * WBM + q * 32 + lane;
        const bool vp = (mp < M) && pf_valid(mp, pfd);
        const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ep.bn_a_h) + (size_t)mp * N + half * WBN + cbeg);
#pragma unroll
        for (int j = 0; j < 4; ++j) ahp[j] = vp ? __ldg(ap + j) : make_uint4(0u, 0u, 0u, 0u);
        if (vp) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + 4));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + 8));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + 12));
        }
      }
      mbar_wait(t_full(ts), (it / TS) & 1u);
      tc_fence_after();
      if (WBN == 256 && NACC == 1 && mt.masks && mt.tc) {
        // ---- tensor-core mask tail.  This thread's row = input pixel (roi, hh, ww) of sub-pixel (a, b) = item's slice.
        const long long m = (long long)tile * WBM + q * 32 + lane;
        int roi = 0, hh = 0, ww = 0;
        const bool valid = (m < M) && pf_decode(m, pfd, roi, hh, ww);
        bool pos = false;
        if (valid) {
          hh -= 1;
          ww -= 1;
          pos = mt.ids && __ldg(mt.ids + roi) > 0;
        }
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + ts * TSTRIDE;
        float* ypos = pos ? mt.y4 + (size_t)m * N + half * 256 : nullptr;
        const int tcb = EL == 2 ? eh * 128 : 0, tce = EL == 2 ? tcb + 128 : 256;
        const uint32_t rrow = (uint32_t)(q * 32 + lane);
        // drain: h = relu(acc + bd) as half into A2 (K-major, 128B swizzle: 4 k-blocks of 64 channels x 128 rows x 128 B)
        {
          float va[32], vb[32];
          auto put = [&](float (&v)[32], int c0) {
            if (ypos) {
              float4* yp = reinterpret_cast<float4*>(ypos + c0);
#pragma unroll
              for (int j = 0; j < 8; ++j) yp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            const uint32_t rowbase = a2_0 + (uint32_t)(c0 >> 6) * 16384u + rrow * 128u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {      // 8 channels = one 16-byte chunk
              float4 b0, b1v;
              asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w) : "r"(evs + 4u * (c0 + 8 * j)));
              asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b1v.x), "=f"(b1v.y), "=f"(b1v.z), "=f"(b1v.w) : "r"(evs + 4u * (c0 + 8 * j + 4)));
              const uint32_t p0 = pack_h2(fmaxf(v[8 * j] + b0.x, 0.f), fmaxf(v[8 * j + 1] + b0.y, 0.f));
              const uint32_t p1 = pack_h2(fmaxf(v[8 * j + 2] + b0.z, 0.f), fmaxf(v[8 * j + 3] + b0.w, 0.f));
              const uint32_t p2 = pack_h2(fmaxf(v[8 * j + 4] + b1v.x, 0.f), fmaxf(v[8 * j + 5] + b1v.y, 0.f));
              const uint32_t p3 = pack_h2(fmaxf(v[8 * j + 6] + b1v.z, 0.f), fmaxf(v[8 * j + 7] + b1v.w, 0.f));
              const uint32_t chunk = (uint32_t)(((c0 & 63) >> 3) + j);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowbase + ((chunk ^ (rrow & 7u)) << 4)), "r"(p0), "r"(p1),
                           "r"(p2), "r"(p3) : "memory");
            }
          };
          tmem_ld32_issue(taddr + (uint32_t)tcb, va);
          tmem_ld_wait(va);
#pragma unroll 1
          for (int c0 = tcb; c0 < tce; c0 += 64) {
            tmem_ld32_issue(taddr + (uint32_t)(c0 + 32), vb);
            put(va, c0);
            tmem_ld_wait(vb);
            if (c0 + 64 < tce) tmem_ld32_issue(taddr + (uint32_t)(c0 + 64), va);
            put(vb, c0 + 32);
            if (c0 + 64 < tce) tmem_ld_wait(va);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of A2 -> visible to the tensor core
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster_release(mapa_rank(a2_full, 0));
          else mbar_arrive(a2_full);
        }
        // logits of this row: ncp columns at the start of the same TMEM stage
        mbar_wait(d2_full, it & 1u);
        tc_fence_after();
        const int a = half >> 1, b = half & 1;
        float* out = mt.masks + ((((size_t)roi * 2 * mt.H + 2 * hh + a) * 2 * mt.W) + 2 * ww + b) * mt.NC;
        const int nch = mt.ncp >> 4, chstep = EL == 2 ? 2 : 1;
        int lastch = eh;
        while (lastch + chstep < nch) lastch += chstep;
        bool released = false;
#pragma unroll 1
        for (int ch = eh; ch < nch; ch += chstep) {
          float lgv[16];
          tmem_ld16(taddr + (uint32_t)(ch * 16), lgv);
          if (ch == lastch) {      // every TMEM read of this warp is done: hand the stage back before the sigmoids
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(mapa_rank(t_empty(ts), 0));
              else mbar_arrive(t_empty(ts));
            }
            released = true;
          }
          if (valid) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const int cls = ch * 16 + k;
              if (cls < mt.NC) out[cls] = 1.f / (1.f + expf(-(lgv[k] + evec[256 + cls])));
            }
          }
        }
        if (!released) {           // a warp without a logits chunk of its own (ncp = 16 and the quarter's second warp)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(mapa_rank(t_empty(ts), 0));
            else mbar_arrive(t_empty(ts));
          }
        }
        continue;
      }
      if (WBN == 256 && mt.masks) {
#pragma unroll 1
        // the tail is bound by the latency of ONE warp per scheduler (ncu: 0.35 IPC, shared-memory pipe 36 %), so with
        // eight epilogue warps (EL = 2) the two warps of a lane quarter take 128 channels each and the second hands
        // its partial logits over through shared memory (one named barrier per item, double-buffered by item parity)
        for (int acc = 0; acc < NACC; ++acc) {
          const long long m = (long long)tile * WBM + acc * 128 + q * 32 + lane;
          int roi = 0, hh = 0, ww = 0;
          const bool valid = (m < M) && pf_decode(m, pfd, roi, hh, ww);
          bool pos = false;
          if (valid) {
            hh -= 1;
            ww -= 1;
            pos = mt.ids && __ldg(mt.ids + roi) > 0;
          }
          uint64_t lg[8][2];
#pragma unroll
          for (int k = 0; k < 8; ++k) lg[k][0] = lg[k][1] = 0ull;
          {
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + ts * TSTRIDE + (uint32_t)(acc * 256);
            float* ypos = pos ? mt.y4 + (size_t)m * N + half * 256 : nullptr;
            const int tcb = EL == 2 ? eh * 128 : 0, tce = EL == 2 ? tcb + 128 : 256;
            switch (mt.NC) {
              case 1: mask_tail_row<1>(taddr, evs, ypos, lg, tcb, tce); break;
              case 2: mask_tail_row<2>(taddr, evs, ypos, lg, tcb, tce); break;
              case 3: mask_tail_row<3>(taddr, evs, ypos, lg, tcb, tce); break;
              case 4: mask_tail_row<4>(taddr, evs, ypos, lg, tcb, tce); break;
              case 5: mask_tail_row<5>(taddr, evs, ypos, lg, tcb, tce); break;
              case 6: mask_tail_row<6>(taddr, evs, ypos, lg, tcb, tce); break;
              default: mask_tail_row<7>(taddr, evs, ypos, lg, tcb, tce); break;
            }
          }
          float lgs[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float e0, o0, e1, o1;
            unpack2(lg[k][0], e0, o0);
            unpack2(lg[k][1], e1, o1);
            lgs[k] = (e0 + o0) + (e1 + o1);
          }
          if (EL == 2) {
            const uint32_t xbuf = stg0 + (uint32_t)(q + 4) * 4096u + (((it * NACC + acc) & 1u) ? 1024u : 0u) + (uint32_t)lane * 32u;
            if (eh == 1) {
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(xbuf), "f"(lgs[0]), "f"(lgs[1]), "f"(lgs[2]), "f"(lgs[3]) : "memory");
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(xbuf + 16u), "f"(lgs[4]), "f"(lgs[5]), "f"(lgs[6]), "f"(lgs[7]) : "memory");
              asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
              continue;          // the first warp of the quarter finishes the pixel
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            float4 p0, p1;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(p0.x), "=f"(p0.y), "=f"(p0.z), "=f"(p0.w) : "r"(xbuf) : "memory");
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(p1.x), "=f"(p1.y), "=f"(p1.z), "=f"(p1.w) : "r"(xbuf + 16u) : "memory");
            lgs[0] += p0.x; lgs[1] += p0.y; lgs[2] += p0.z; lgs[3] += p0.w;
            lgs[4] += p1.x; lgs[5] += p1.y; lgs[6] += p1.z; lgs[7] += p1.w;
          }
          if (valid) {
            const int a = half >> 1, b = half & 1;
            float* out = mt.masks + ((((size_t)roi * 2 * mt.H + 2 * hh + a) * 2 * mt.W) + 2 * ww + b) * mt.NC;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < mt.NC) out[k] = 1.f / (1.f + expf(-(lgs[k] + __ldg(mt.b1 + k))));
          }
        }
      } else
#pragma unroll 1
      for (int acc = 0; acc < NACC; ++acc) {
        const int mrow0 = tile * WBM + acc * 128 + q * 32;
        const long long m = (long long)mrow0 + lane;
        const bool valid = (m < M) && pf_valid(m, pfd);
        // half flavour of the fused BN backward: the activation chunk of the NEXT 32 columns is fetched while the
        // current one is processed (the global-load latency was the longest stall of this epilogue)
        const uint4* arow_h0 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ep.bn_a_h) + (size_t)m * N + half * WBN);
        uint4 ahc[4];
        if (EL == 2 && ep.bn_a && !act_tma) {
          if (NACC == 1 && (dbg & 64)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ahc[j] = ahp[j];
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ahc[j] = (valid && !(dbg & 128)) ? __ldg(arow_h0 + cbeg / 8 + j) : make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
          }
        }
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + kColsPerWarp; c0 += 32, ++nst) {
          float v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ts * TSTRIDE + (uint32_t)(acc * WBN + c0), v);
          const uint32_t sbuf = sbuf0;
          const uint32_t sbufh = sbuf0 + (nst & 1u) * 2048u;
          const int n0 = half * WBN + c0;
          if (EL == 2 && ep.bn_a) {
            // ---- fused BN(+ReLU) backward on half tensors, register-only: v = d(a) of this thread's row (loss-scaled)
            uint4 ahn[4];
            if (act_tma) {
              mbar_wait(act_full(ew), actn & 1u);
              ++actn;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t aaddr = astg + (uint32_t)lane * 64u + (uint32_t)(((uint32_t)j ^ (((uint32_t)lane >> 1) & 3u)) << 4);
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(ahc[j].x), "=r"(ahc[j].y), "=r"(ahc[j].z), "=r"(ahc[j].w) : "r"(aaddr) : "memory");
              }
              __syncwarp();
              if (lane == 0) {      // the tile is in registers: its buffer takes the next chunk (of this item or of the next one)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (c0 + 32 < cbeg + kColsPerWarp) issue_act(item, c0 + 32);
                else if (item + ncl < nitems) issue_act(item + ncl, cbeg);
              }
            } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ahn[j] = (valid && c0 + 32 < cbeg + kColsPerWarp && !(dbg & 128)) ? __ldg(arow_h0 + (c0 + 32) / 8 + j)
                                                                                   : make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
            }
            // per element only g = d(a)*act'(a) and g*a: the per-column constants of dgamma = sum g*(a-beta)/gamma are
            // applied AFTER the column reduction, where lane = column (every broadcast constant load costs a full
            // shared-memory wavefront per lane-row, and this epilogue was bound by them)
            float tt[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t w0 = (j & 1) ? ahc[j >> 1].z : ahc[j >> 1].x, w1 = (j & 1) ? ahc[j >> 1].w : ahc[j >> 1].y;
              const float aa[4] = {h2f((uint16_t)(w0 & 0xffffu)), h2f((uint16_t)(w0 >> 16)), h2f((uint16_t)(w1 & 0xffffu)),
                                   h2f((uint16_t)(w1 >> 16))};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                bool pass = valid;
                if (actk == MYOLO_ACT_RELU) pass = pass && aa[e] > 0.f;
                else if (actk == MYOLO_ACT_RELU6) pass = pass && aa[e] > 0.f && aa[e] < 6.f;
                const float g = pass ? v[4 * j + e] : 0.f;
                v[4 * j + e] = g;
                tt[4 * j + e] = g * aa[e];
              }
            }
            // stores first (they only need g), then the two column reductions, which destroy their inputs
            uint32_t tt_pack[2][2];
            wait_store();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 sc;
              asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sc.x), "=f"(sc.y), "=f"(sc.z), "=f"(sc.w) : "r"(evs + 4u * (n0 + 4 * j)));
              float o[4] = {v[4 * j] * sc.x, v[4 * j + 1] * sc.y, v[4 * j + 2] * sc.z, v[4 * j + 3] * sc.w};
              if (st_h) {
                const uint32_t p0 = pack_h2(o[0], o[1]), p1 = pack_h2(o[2], o[3]);
                o[0] = h2f((uint16_t)(p0 & 0xffffu)); o[1] = h2f((uint16_t)(p0 >> 16));
                o[2] = h2f((uint16_t)(p1 & 0xffffu)); o[3] = h2f((uint16_t)(p1 >> 16));
                tt_pack[j & 1][0] = p0;
                tt_pack[j & 1][1] = p1;
                if (j & 1) {
                  const uint32_t haddr = sbufh + (uint32_t)lane * 64u + (uint32_t)((((uint32_t)j >> 1) ^ (((uint32_t)lane >> 1) & 3u)) << 4);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(haddr), "r"(tt_pack[0][0]), "r"(tt_pack[0][1]), "r"(p0), "r"(p1));
                }
              }
              if (st_f32) {
                const uint32_t addr = sbuf + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]));
              }
            }
            if (!(dbg & 256)) {
              const float s0 = warp_colsum32(v, lane);
              const float s1 = (warp_colsum32(tt, lane) - evec[N + n0 + lane] * s0) * evec[2 * N + n0 + lane];
              atomicAdd(colacc + n0 + lane, s0);
              atomicAdd(colacc + 256 + n0 + lane, s1);
            } else if (v[0] + tt[31] == 12345.f) {   // timing experiment: no column reductions
              atomicAdd(colacc + n0 + lane, v[1] + tt[2]);
            }
            if (!act_tma) {
#pragma unroll
              for (int j = 0; j < 4; ++j) ahc[j] = ahn[j];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              if (mrow0 < M && !(dbg & 4)) {
                if (st_f32) tma_store_2d(&tmC, sbuf, n0, mrow0);
                if (st_h) tma_store_2d(&tmCh, sbufh, n0, mrow0);
              } else if (ring_h) {
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
            continue;
          }
          // the previous store must have finished reading the staging buffer (the shared memory a second
          // buffer would take is worth more as a fifth weight stage)
          wait_store();
          if (ep.bn_a) {
            // ---- fused BN(+ReLU) backward.  v = d(a) for this thread's row; a is read in place.
            const uint32_t sbuf2 = sbuf + 4u * 4096u;
            const float4* arow = reinterpret_cast<const float4*>(ep.bn_a + (size_t)m * N + n0);
            const uint4* arow_h = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ep.bn_a_h) + (size_t)m * N + n0);
            uint4 ah[4];
            if (EL == 2) {
#pragma unroll
              for (int j = 0; j < 4; ++j) ah[j] = valid ? __ldg(arow_h + j) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 av = make_float4(0.f, 0.f, 0.f, 0.f), be, ig;
              if (EL == 2) {
                const uint32_t w0 = (j & 1) ? ah[j >> 1].z : ah[j >> 1].x, w1 = (j & 1) ? ah[j >> 1].w : ah[j >> 1].y;
                av = make_float4(h2f((uint16_t)(w0 & 0xffffu)), h2f((uint16_t)(w0 >> 16)), h2f((uint16_t)(w1 & 0xffffu)),
                                 h2f((uint16_t)(w1 >> 16)));
              } else if (valid) av = __ldg(arow + j);
              asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(be.x), "=f"(be.y), "=f"(be.z), "=f"(be.w) : "r"(evs + 4u * (N + n0 + 4 * j)));
              asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(ig.x), "=f"(ig.y), "=f"(ig.z), "=f"(ig.w) : "r"(evs + 4u * (2 * N + n0 + 4 * j)));
              const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {be.x, be.y, be.z, be.w}, gg[4] = {ig.x, ig.y, ig.z, ig.w};
              float g[4], t[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                bool pass = valid;
                if (actk == MYOLO_ACT_RELU) pass = pass && aa[e] > 0.f;
                else if (actk == MYOLO_ACT_RELU6) pass = pass && aa[e] > 0.f && aa[e] < 6.f;
                g[e] = pass ? v[4 * j + e] : 0.f;
                t[e] = g[e] * (aa[e] - bb[e]) * gg[e];
              }
              const uint32_t off = (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + off), "f"(g[0]), "f"(g[1]), "f"(g[2]), "f"(g[3]));
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf2 + off), "f"(t[0]), "f"(t[1]), "f"(t[2]), "f"(t[3]));
            }
            __syncwarp();
            // column pass: lane = column.  sums of g (dbeta) and g*xhat (dgamma); g is scaled in place to d(pre-BN)
            {
              const float scl = evec[n0 + lane];
              float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                const uint32_t off = (uint32_t)r * 128u + (uint32_t)((((uint32_t)lane >> 2) ^ (uint32_t)(r & 7)) << 4) + ((uint32_t)lane & 3u) * 4u;
                float gv, tv;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(gv) : "r"(sbuf + off));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(tv) : "r"(sbuf2 + off));
                s0 += gv;
                s1 += tv;
                float o = gv * scl;
                if (rnd) o = round_tf32(o);
                if (st_h) {
                  const uint16_t hv = f2h_sat(o);
                  o = h2f(hv);
                  const uint32_t hoff = (uint32_t)r * 64u + (uint32_t)((((uint32_t)lane >> 3) ^ (((uint32_t)r >> 1) & 3u)) << 4) +
                                        ((uint32_t)lane & 7u) * 2u;
                  asm volatile("st.shared.b16 [%0], %1;" ::"r"(sbufh + hoff), "h"(hv));
                }
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbuf + off), "f"(o));
              }
              atomicAdd(colacc + n0 + lane, s0);
              atomicAdd(colacc + 256 + n0 + lane, s1);
            }
          } else {
          uint32_t hp0 = 0u, hp1 = 0u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 sc, sf;
            asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sc.x), "=f"(sc.y), "=f"(sc.z), "=f"(sc.w) : "r"(evs + 4u * (n0 + 4 * j)));
            asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sf.x), "=f"(sf.y), "=f"(sf.z), "=f"(sf.w) : "r"(evs + 4u * (N + n0 + 4 * j)));
            float o[4] = {fmaf(v[4 * j], sc.x, sf.x), fmaf(v[4 * j + 1], sc.y, sf.y), fmaf(v[4 * j + 2], sc.z, sf.z),
                          fmaf(v[4 * j + 3], sc.w, sf.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (actk == MYOLO_ACT_RELU) o[e] = fmaxf(o[e], 0.f);
              else if (actk == MYOLO_ACT_RELU6) o[e] = fminf(fmaxf(o[e], 0.f), 6.f);
              if (rnd) o[e] = round_tf32(o[e]);
              if (!valid) o[e] = 0.f;   // padded-flat pad rows stay zero
            }
            if (st_h) {   // the half copy; hp0 / hp1 keep the packed words until the 16-byte chunk is complete
              const uint32_t p0 = pack_h2(o[0], o[1]), p1 = pack_h2(o[2], o[3]);
              o[0] = h2f((uint16_t)(p0 & 0xffffu)); o[1] = h2f((uint16_t)(p0 >> 16));
              o[2] = h2f((uint16_t)(p1 & 0xffffu)); o[3] = h2f((uint16_t)(p1 >> 16));
              if (j & 1) {
                const uint32_t haddr = sbufh + (uint32_t)lane * 64u + (uint32_t)((((uint32_t)j >> 1) ^ (((uint32_t)lane >> 1) & 3u)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(haddr), "r"(hp0), "r"(hp1), "r"(p0), "r"(p1));
              } else {
                hp0 = p0;
                hp1 = p1;
              }
            }
            if (ep.st_sums) {   // batch statistics of the STORED result (half-rounded when the output is half; pad rows are zero)
              v[4 * j] = o[0]; v[4 * j + 1] = o[1]; v[4 * j + 2] = o[2]; v[4 * j + 3] = o[3];
            }
            if (st_f32) {
              const uint32_t addr = sbuf + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]));
            }
          }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (mrow0 < M && !(dbg & 4)) {
              if (st_f32) tma_store_2d(&tmC, sbuf, n0, mrow0);
              if (st_h) tma_store_2d(&tmCh, sbufh, n0, mrow0);
            } else if (ring_h) {
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");   // keep one group per chunk: the ring counts groups
            }
          }
          if (EL == 4 && ep.st_sums && !ep.bn_a) {
            // tf32 flavour (pointwise convolutions of the backbone): four epilogue warps, any N; the chunk's column sums
            // go straight to the fp64 workspace (one atomic per column and 32-row chunk)
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
            const float s0 = warp_colsum32(v, lane), s1 = warp_colsum32(sq, lane);
            if (s0 != 0.f) atomicAdd(ep.st_sums + n0 + lane, (double)s0);
            if (s1 != 0.f) atomicAdd(ep.st_sums + N + n0 + lane, (double)s1);
          }
          if (EL == 2 && ep.st_sums && !ep.bn_a) {   // column sums of the chunk while its TMA store is in flight
            // shifted by a per-column pivot (the layer's moving mean): E[x^2] - E[x]^2 would lose mean^2 / variance digits
            // The shift is applied to the 32-row column sums, where lane = column (sum (x-p) = S0 - n p, sum (x-p)^2 =
            // S1 - 2 p S0 + n p^2; over 32 rows that costs ~1e-7 * (mean^2 + var) / var of relative accuracy) instead of to
            // every element, which needed one shuffle per column to hand each row its pivots.
            const float pv = ep.st_pivot ? __ldg(ep.st_pivot + n0 + lane) : 0.f;
            const float nv = (float)__popc(__ballot_sync(0xffffffffu, valid));
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];      // rows that are not valid hold zeros
            const float r0 = warp_colsum32(v, lane), r1 = warp_colsum32(sq, lane);
            const float s0 = fmaf(-nv, pv, r0);
            const float s1 = fmaf(pv, fmaf(nv, pv, -2.f * r0), r1);
            float* slot = colacc + (ew < 2 ? ew * 256 : 1024 + (ew - 2) * 256) + (c0 - cbeg) + lane;
            slot[0] += s0;
            slot[128] += s1;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(mapa_rank(t_empty(ts), 0));   // the leader's MMA thread owns the TMEM hand-back
        else mbar_arrive(t_empty(ts));
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (ep.bn_a) {   // one fp64 atomic per (CTA, column, statistic)
    for (int n = threadIdx.x; n < 2 * N; n += blockDim.x) {
      const float vsum = colacc[(n / N) * 256 + (n % N)];
      if (vsum != 0.f) atomicAdd(ep.bn_ws + n, (double)vsum);
    }
  } else if (EL == 2 && ep.st_sums) {   // the four warps of a column half, in a fixed order
    for (int n = threadIdx.x; n < 2 * N; n += blockDim.x) {
      const int stat = n / N, col = n - stat * N, h = col >> 7, cc = col & 127;
      double acc = 0.0;
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) {
        const int w = h * 4 + w4;
        acc += (double)colacc[(w < 2 ? w * 256 : 1024 + (w - 2) * 256) + stat * 128 + cc];
      }
      if (acc != 0.0) atomicAdd(ep.st_sums + n, acc);
    }
  }
  if (CG == 2) cluster_sync_all();   // no CTA of the pair may free TMEM / exit while the other still uses it
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_2sm(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

}  // namespace tc
}  // namespace myolo

using namespace myolo;
using namespace myolo::tc;


// largest dynamic shared memory a CTA-pair launch may ask for (227 KB minus the static barriers)
constexpr uint32_t kWinSmemMax = 230400u;
// tensor-core mask tail: shared memory = alignment slack + windows + weight stages + A2 (64 KB) + B2 + epilogue vectors
static uint32_t tail_smem(int nawin, int nwb, int ncp) {
  return 1024u + (uint32_t)nawin * (uint32_t)win_rows(1) * 128u + (uint32_t)nwb * 16384u + 65536u + (uint32_t)(ncp / 2) * 512u + 2048u;
}

// min_m: smallest M for which a plain (single-tap) GEMM is routed here
static int win_shape_ok(long long lda, long long ldc, long long M, int N, int K, int ntaps, const int* shifts_host,
                        int accumulate, long long min_m) {
  if (!(M >= 1 && M < (1LL << 31) - 4096 && N >= 128 && (N % 128) == 0 && N <= 1024 && K >= BK && (K % BK) == 0 && (lda % 4) == 0 && (ldc % 4) == 0 &&
        ntaps >= 1 && ntaps <= 32 && !accumulate))
    return 0;
  if (!shifts_host) return ntaps == 1 && M >= min_m;
  for (int t = 0; t < ntaps; ++t)
    if (shifts_host[t] < -WHALO || shifts_host[t] > WHALO) return 0;
  if (ntaps == 1 && M < min_m) return 0;
  return 1;
}

extern "C" int myolo_gemm_taps_win_supported(long long lda, long long ldc, long long M, int N, int K, int ntaps,
                                             const int* shifts_host, int accumulate) {
  // plain GEMM: worth it only for large M (persistent 256-row tiles)
  return win_shape_ok(lda, ldc, M, N, K, ntaps, shifts_host, accumulate, 4096);
}

struct BnBwd {   // host-side bundle of the fused BN-backward epilogue arguments (all null = off)
  const float* a;
  const float* gamma;
  const float* beta;
  const float* var;
  double* ws;
  float eps;
};

// half-operand launch options: A / Bt (and bnb.a) are IEEE half; Ch = half output [M][N] pitch ldch (nullable);
// no_f32 skips the fp32 output C; acc_scale = device scalar folded into the accumulator
struct HalfIO {
  int on;
  void* Ch;
  long long ldch;
  int no_f32;
  const float* acc_scale;
  double* st_sums;   // per-column sum / sum of squares of the stored result ([2][N] fp64, += )
  const float* st_pivot;   // per-column value subtracted before the sums are taken (nullable)
};

static int launch_win(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M, int N, int K,
                      int ntaps, const int* shifts_host, const float* bias, const float* scale, const float* shift_c,
                      int act, int pf_w1, int pf_blk, int accumulate, const MaskTail& mt, myolo_stream stream,
                      const BnBwd& bnb = BnBwd{}, const HalfIO& hio = HalfIO{}, int nseg = 1, const int* seg_off = nullptr,
                      double* stats_sums = nullptr) {
  MYOLO_CHECK_ARG(A && Bt && ((((uintptr_t)A | (uintptr_t)Bt | (uintptr_t)C) & 15) == 0));
  MYOLO_CHECK_ARG(C || (hio.on && hio.no_f32));
  MYOLO_CHECK_ARG(win_shape_ok(lda, ldc, M, N, K, ntaps, shifts_host, accumulate, hio.on ? 1 : 4096));
  if (hio.on) {
    MYOLO_CHECK_ARG((K % 64) == 0 && (lda % 8) == 0 && (N % 256) == 0 && !(act & MYOLO_ROUND_TF32));
    MYOLO_CHECK_ARG(!hio.Ch || ((hio.ldch % 8) == 0 && ((uintptr_t)hio.Ch & 15) == 0));
    MYOLO_CHECK_ARG(hio.Ch || !hio.no_f32 || mt.masks);
    MYOLO_CHECK_ARG(!(hio.Ch && !hio.no_f32));   // one output per launch: the eight epilogue warps own one staging tile each
  }
  MYOLO_CHECK_ARG((scale == nullptr) == (shift_c == nullptr));
  MYOLO_CHECK_ARG(!(accumulate && (scale || (act & 0xff) != MYOLO_ACT_NONE)));
  MYOLO_CHECK_ARG(pf_w1 <= 0 || pf_blk > 0);
  TapShifts sh;
  for (int t = 0; t < 32; ++t) sh.s[t] = (shifts_host && t < ntaps) ? shifts_host[t] : 0;
  MYOLO_CHECK_ARG(nseg >= 1 && nseg <= 8 && (nseg == 1 || (seg_off && ntaps == 1 && !hio.on && !mt.masks && !bnb.a)));
  TapShifts seg;
  for (int g = 0; g < 32; ++g) seg.s[g] = (seg_off && g < nseg) ? seg_off[g] : 0;
  CUtensorMap ta, tb;
  // rows in [M, M + max shift) are the zero guard rows of the padded-flat tensor; everything else out of
  // range (negative rows, the tail of the last tile) is TMA zero fill
  int maxs = 0;
  for (int t = 0; t < ntaps; ++t) maxs = sh.s[t] > maxs ? sh.s[t] : maxs;
  long long maxseg = 0;
  for (int g = 0; g < nseg; ++g) {
    MYOLO_CHECK_ARG(seg.s[g] >= 0);
    maxseg = seg.s[g] > maxseg ? seg.s[g] : maxseg;
  }
  static int bo_mode = -1;
  if (bo_mode < 0) {
    const char* e = getenv("MYOLO_WIN_BO");
    bo_mode = e ? atoi(e) : 0;  // experiment switches: 2 = no TMA loads, 4 = skip the global stores, 8 = 128-column slices, 16 = two accumulators, 32 = no CTA pairs, 64 = early activation fetch in the fused BN backward (measured slower), 512 = activation of the fused BN backward by per-lane loads instead of TMA, 128 / 256 = timing experiments: no activation loads / no column reductions in that epilogue (wrong results)
  }
  const int wbn = (((bo_mode & 8) && !mt.masks) || (N % 256) != 0) ? 128 : 256;
  // one accumulator + two TMEM stages (epilogue overlapped with the next item's main loop) is the default;
  // the two-accumulator variant halves the weight traffic from L2 but exposes its epilogue
  const int nacc = (bo_mode & 16) ? 2 : 1;
  // CTA-pair variant (cta_group::2): 256-column slices, one accumulator per CTA; MYOLO_WIN_BO & 32 disables it
  const int cg = (wbn == 256 && nacc == 1 && !(bo_mode & 32) && !(bo_mode & 2)) ? 2 : 1;
  int rc;
  CUtensorMap tc_, tch;
  if (hio.on) {
    MYOLO_CHECK_ARG(wbn == 256 && cg == 2);
    rc = get_map_h(A, M + maxs, K, lda, win_box(nacc), 64, &ta);
    if (rc) return rc;
    rc = get_map_h(Bt, (long long)ntaps * N, K, K, wbn / cg, 64, &tb);
    if (rc) return rc;
    if (!hio.no_f32) {
      rc = get_map(C, M, N, ldc, 32, &tc_);
      if (rc) return rc;
    } else {
      tc_ = ta;   // never dereferenced
    }
    if (hio.Ch) {
      rc = get_map_h(hio.Ch, M, N, hio.ldch, 32, 32, &tch);
      if (rc) return rc;
    } else {
      tch = ta;
    }
  } else {
    rc = get_map(A, M + maxs + maxseg, K, lda, win_box(nacc), &ta);
    if (rc) return rc;
    rc = get_map(Bt, (long long)nseg * ntaps * N, K, K, wbn / cg, &tb);
    if (rc) return rc;
    rc = get_map(C, M, N, ldc, 32, &tc_);
    if (rc) return rc;
    tch = tc_;
  }
  static bool attr_set = false;
  static int max_clusters = 0, max_clusters_h = 0;
  if (!attr_set) {
    MYOLO_CUDA(cudaFuncSetAttribute(tc_conv_win_kernel<128, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_smem(1)));
    MYOLO_CUDA(cudaFuncSetAttribute(tc_conv_win_kernel<256, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_smem(1)));
    MYOLO_CUDA(cudaFuncSetAttribute(tc_conv_win_kernel<128, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_smem(2)));
    MYOLO_CUDA(cudaFuncSetAttribute(tc_conv_win_kernel<256, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_smem(2)));
    MYOLO_CUDA(cudaFuncSetAttribute(tc_conv_win_kernel<256, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWinSmemMax));
    MYOLO_CUDA(cudaFuncSetAttribute(tc_conv_win_kernel<256, 1, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWinSmemMax));
    {   // how many CTA pairs fit at once (GPCs with an odd SM count leave SMs unpaired)
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(kNumSMs);
      q.blockDim = dim3(kThreads);
      q.dynamicSmemBytes = win_smem(1);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      q.attrs = at;
      q.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, tc_conv_win_kernel<256, 1, 2>, &q) != cudaSuccess) max_clusters = 0;
      (void)cudaGetLastError();
      q.blockDim = dim3(win_threads(2));
      if (cudaOccupancyMaxActiveClusters(&max_clusters_h, tc_conv_win_kernel<256, 1, 2, 2>, &q) != cudaSuccess) max_clusters_h = 0;
      (void)cudaGetLastError();
    }
    attr_set = true;
  }
  Epi ep{bias, scale, shift_c, act, pf_w1, pf_blk, accumulate, bnb.a, bnb.gamma, bnb.beta, bnb.var, bnb.ws, bnb.eps,
         hio.on ? (const void*)bnb.a : nullptr, hio.on ? hio.no_f32 : 0, (hio.on && hio.Ch) ? 1 : 0, hio.acc_scale};
  if (hio.on && bnb.a && hio.no_f32 && hio.Ch && cg == 2 && !(bo_mode & 512)) {
    // the activation of the fused BN backward by TMA: same geometry as the half output (32 x 32 tiles, 64B swizzle)
    rc = get_map_h(bnb.a, M, N, N, 32, 32, &tc_);
    if (rc) return rc;
    ep.bn_tma = 1;
  }
  if (hio.on && hio.st_sums) {
    MYOLO_CHECK_ARG(N == 256 && !bnb.a && !mt.masks);
    ep.st_sums = hio.st_sums;
    ep.st_pivot = hio.st_pivot;
  }
  if (stats_sums) {
    MYOLO_CHECK_ARG(!hio.on && !bnb.a && !mt.masks && !accumulate);
    ep.st_sums = stats_sums;
  }
  cudaStream_t st = as_stream(stream);
  const int maxcl = hio.on ? max_clusters_h : max_clusters;
  MaskTail mtl = mt;
  uint32_t smem_pair = win_smem(1);
  if (mtl.masks && mtl.tc) {
    // ring depths that fit next to A2 / B2: four windows while the class block is small, three for many classes
    MYOLO_CHECK_ARG(cg == 2 && maxcl > 0 && ntaps == 1 && N == 1024 && K == 256 && mtl.ncp >= 16 && mtl.ncp <= 128 && (mtl.ncp % 16) == 0);
    mtl.nwb = 4;
    mtl.nawin = tail_smem(4, 4, mtl.ncp) <= kWinSmemMax ? 4 : 3;
    smem_pair = tail_smem(mtl.nawin, mtl.nwb, mtl.ncp);
    MYOLO_CHECK_ARG(smem_pair <= kWinSmemMax);
  }
  if (cg == 2 && maxcl > 0) {
    const int nitems = (int)ceil_div(M, 256) * (N / wbn);
    const int ncl = nitems < maxcl ? nitems : maxcl;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * ncl);
    cfg.blockDim = dim3(win_threads(hio.on ? 2 : 4));
    cfg.dynamicSmemBytes = smem_pair;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (hio.on)
      MYOLO_CUDA(cudaLaunchKernelEx(&cfg, tc_conv_win_kernel<256, 1, 2, 2>, ta, tb, tc_, tch, M, N, K, ntaps, sh, ep, mtl, nitems, bo_mode, nseg, seg));
    else
      MYOLO_CUDA(cudaLaunchKernelEx(&cfg, tc_conv_win_kernel<256, 1, 2>, ta, tb, tc_, tch, M, N, K, ntaps, sh, ep, mtl, nitems, bo_mode, nseg, seg));
    return MYOLO_OK;
  }
  if (hio.on) {
    set_error("half-operand conv kernel needs CTA pairs (cudaOccupancyMaxActiveClusters returned 0)");
    return MYOLO_ERR_CUDA;
  }
  if (cg == 2) {   // no pair fits (should not happen on B200): rebuild the weight map for the single-CTA box
    rc = get_map(Bt, (long long)nseg * ntaps * N, K, K, wbn, &tb);
    if (rc) return rc;
  }
  const int nitems = (int)ceil_div(M, 128 * nacc) * (N / wbn);
  const int grid = nitems < kNumSMs ? nitems : kNumSMs;
#define MYOLO_WIN_LAUNCH(BN_, NA_) \
  tc_conv_win_kernel<BN_, NA_, 1><<<grid, kThreads, win_smem(NA_), st>>>(ta, tb, tc_, tch, M, N, K, ntaps, sh, ep, mt, nitems, bo_mode, nseg, seg)
  if (wbn == 256 && nacc == 1) MYOLO_WIN_LAUNCH(256, 1);
  else if (wbn == 256) MYOLO_WIN_LAUNCH(256, 2);
  else if (nacc == 1) MYOLO_WIN_LAUNCH(128, 1);
  else MYOLO_WIN_LAUNCH(128, 2);
#undef MYOLO_WIN_LAUNCH
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_gemm_taps_win(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M,
                                   int N, int K, int ntaps, const int* shifts_host, const float* bias, const float* scale,
                                   const float* shift_c, int act, int pf_w1, int pf_blk, int accumulate,
                                   myolo_stream stream) {
  MaskTail mt{};
  return launch_win(A, lda, Bt, C, ldc, M, N, K, ntaps, shifts_host, bias, scale, shift_c, act, pf_w1, pf_blk, accumulate,
                    mt, stream);
}

namespace myolo {
__global__ void stats_finalize_kernel(double* __restrict__ sums, const float* __restrict__ pivot, float* __restrict__ mean,
                                      float* __restrict__ var, int C, double inv_count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sums[c] * inv_count;                      // mean of (x - pivot)
  const double vv = sums[C + c] * inv_count - m * m;
  mean[c] = (float)(m + (pivot ? (double)pivot[c] : 0.0));
  var[c] = (float)(vv > 0.0 ? vv : 0.0);
  sums[c] = 0.0;
  sums[C + c] = 0.0;
}
}  // namespace myolo

// Plain GEMM over k-segments on the persistent kernel: C[m][n] = sum_g sum_k A[m + seg_off[g]][k] * Bt[g][n][k], with the
// batch statistics of C (mean / biased variance per column over the M rows) reduced in the epilogue when mean != null.
// The pointwise convolutions of the backbone in the 3xTF32 modes: A holds the tf32 high and low parts of the activation
// lo_off rows apart, Bt the [hi | hi | lo] weight triple, seg_off = {0, lo_off, 0}.
extern "C" int myolo_gemm_segs_win_supported(long long lda, long long ldc, long long M, int N, int K, int nseg) {
  return nseg >= 1 && nseg <= 8 && N <= 1024 && win_shape_ok(lda, ldc, M, N, K, 1, nullptr, 0, 4096);
}

extern "C" int myolo_gemm_segs_win(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M, int N,
                                   int K, int nseg, const int* seg_off_host, float* mean, float* var, double* ws,
                                   myolo_stream stream) {
  MYOLO_CHECK_ARG(myolo_gemm_segs_win_supported(lda, ldc, M, N, K, nseg) && seg_off_host);
  MYOLO_CHECK_ARG((mean == nullptr) == (var == nullptr) && (!mean || ws));
  MaskTail mt{};
  int rc = launch_win(A, lda, Bt, C, ldc, M, N, K, 1, nullptr, nullptr, nullptr, nullptr, MYOLO_ACT_NONE, 0, 0, 0, mt, stream,
                      BnBwd{}, HalfIO{}, nseg, seg_off_host, mean ? ws + 16 : nullptr);
  if (rc || !mean) return rc;
  stats_finalize_kernel<<<(N + 127) / 128, 128, 0, as_stream(stream)>>>(ws + 16, nullptr, mean, var, N, 1.0 / (double)M);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_deconv_mask_fwd_supported(int Cmid, int NC) { return Cmid == 256 && NC >= 1 && NC <= 128; }

// which tail the fused kernel runs: exact-fp32 FMA chains in the epilogue registers (NC <= 7: measured 0.81 ms against
// 1.07 ms for the tensor-core tail at 4 classes -- the logits take the place of the accumulator stage, so the stage is
// handed back only after a GEMM that queues behind the next item's main loop) or the 1x1 conv as a second tcgen05 GEMM
// (8 <= NC <= 128, where the FMA chains would cost more than the deconvolution itself).
// MYOLO_MASK_TAIL = ffma | tc overrides the choice where both exist (A/B measurements).
static int tail_on_tensor_core(int NC, int half_flavour) {
  if (NC > 7) return 1;
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("MYOLO_MASK_TAIL");
    mode = e && strcmp(e, "ffma") == 0 ? 1 : (e && strcmp(e, "tc") == 0 ? 2 : 0);
  }
  (void)half_flavour;
  if (mode == 2) return 1;
  return 0;
}
static MaskTail make_tail(const float* bd, const float* w1, const float* b1, float* masks, const int* ids, float* y4, int H,
                          int W, int NC, int half_flavour) {
  MaskTail mt{bd, w1, b1, masks, ids, y4, H, W, NC, 0, 0, 0, 0};
  mt.tc = tail_on_tensor_core(NC, half_flavour);
  mt.ncp = (NC + 15) / 16 * 16;
  return mt;
}

extern "C" int myolo_deconv_mask_fwd(const float* a4, const float* kd, const float* bd, const float* w1, const float* b1,
                                     float* masks, const int* target_ids, float* y4, int n_roi, int H, int W, int Cmid,
                                     int NC, myolo_stream stream) {
  MYOLO_CHECK_ARG(a4 && kd && bd && w1 && b1 && masks && y4 && n_roi > 0 && H > 0 && W > 0);
  MYOLO_CHECK_ARG(myolo_deconv_mask_fwd_supported(Cmid, NC));
  MaskTail mt = make_tail(bd, w1, b1, masks, target_ids, y4, H, W, NC, 0);
  const long long M = (long long)n_roi * (H + 1) * (W + 1);
  return launch_win(a4, Cmid, kd, y4, 4 * Cmid, M, 4 * Cmid, Cmid, 1, nullptr, nullptr, nullptr, nullptr, MYOLO_ACT_NONE,
                    W + 1, (H + 1) * (W + 1), 0, mt, stream);
}

extern "C" int myolo_bn_epi_finalize(double* sums, const float* gamma, const float* var, float eps, float* dgamma,
                                     float* dbeta, float* dbias, int C, myolo_stream stream);

extern "C" int myolo_gemm_taps_bnbwd_supported(long long lda, long long ldc, long long M, int N, int K, int ntaps,
                                               const int* shifts_host) {
  return N == 256 && ldc == N && myolo_gemm_taps_win_supported(lda, ldc, M, N, K, ntaps, shifts_host, 0);
}

extern "C" int myolo_gemm_taps_bnbwd(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M,
                                     int N, int K, int ntaps, const int* shifts_host, int pf_w1, int pf_blk,
                                     const float* a_out, const float* gamma, const float* beta, const float* var, float eps,
                                     int act, float* dgamma, float* dbeta, float* dbias, double* ws, myolo_stream stream) {
  MYOLO_CHECK_ARG(a_out && gamma && beta && var && dgamma && dbeta && ws);
  MYOLO_CHECK_ARG(myolo_gemm_taps_bnbwd_supported(lda, ldc, M, N, K, ntaps, shifts_host));
  MaskTail mt{};
  BnBwd bnb{a_out, gamma, beta, var, ws + 16, eps};       // sums live behind the ticket words of the BN workspace
  int rc = launch_win(A, lda, Bt, C, ldc, M, N, K, ntaps, shifts_host, nullptr, nullptr, nullptr, act, pf_w1, pf_blk, 0, mt,
                      stream, bnb);
  if (rc) return rc;
  return myolo_bn_epi_finalize(ws + 16, gamma, var, eps, dgamma, dbeta, dbias, N, stream);
}

// ------------------------------------------------------------------------------------------
// half-operand (kind::f16) entry points: same kernel, IEEE-half A / Bt, fp32 accumulation in TMEM
// ------------------------------------------------------------------------------------------
extern "C" int myolo_gemm_taps_h_supported(long long lda, long long M, int N, int K, int ntaps, const int* shifts_host) {
  // the only half-operand GEMM kernel: it takes every M (no tf32-style hand-over of small plain GEMMs)
  return (N % 256) == 0 && (K % 64) == 0 && (lda % 8) == 0 && win_shape_ok(lda, N, M, N, K, ntaps, shifts_host, 0, 1);
}

extern "C" int myolo_gemm_taps_h(const void* A, long long lda, const void* Bt, float* C, long long ldc, void* Ch,
                                 long long ldch, long long M, int N, int K, int ntaps, const int* shifts_host,
                                 const float* bias, const float* scale, const float* shift_c, int act, int pf_w1, int pf_blk,
                                 const float* acc_scale, myolo_stream stream) {
  MYOLO_CHECK_ARG(C || Ch);
  MYOLO_CHECK_ARG(myolo_gemm_taps_h_supported(lda, M, N, K, ntaps, shifts_host));
  MaskTail mt{};
  HalfIO hio{1, Ch, ldch, C ? 0 : 1, acc_scale};
  return launch_win(reinterpret_cast<const float*>(A), lda, reinterpret_cast<const float*>(Bt), C, C ? ldc : N, M, N, K, ntaps,
                    shifts_host, bias, scale, shift_c, act, pf_w1, pf_blk, 0, mt, stream, BnBwd{}, hio);
}


// myolo_gemm_taps_h with an fp32 result and the batch statistics of that result (valid rows only: n_valid of them) taken in
// the epilogue: mean / biased variance per output channel.  ws: the BN workspace (zero before, zero after).
extern "C" int myolo_gemm_taps_h_stats(const void* A, long long lda, const void* Bt, float* C, long long ldc, long long M, int N,
                                       int K, int ntaps, const int* shifts_host, const float* bias, int pf_w1, int pf_blk,
                                       const float* pivot, float* mean, float* var, double* ws, long long n_valid,
                                       myolo_stream stream) {
  MYOLO_CHECK_ARG(C && mean && var && ws && n_valid > 0 && N == 256 && pivot != mean);
  MYOLO_CHECK_ARG(myolo_gemm_taps_h_supported(lda, M, N, K, ntaps, shifts_host));
  MaskTail mt{};
  HalfIO hio{1, nullptr, 0, 0, nullptr, ws + 16, pivot};
  int rc = launch_win(reinterpret_cast<const float*>(A), lda, reinterpret_cast<const float*>(Bt), C, ldc, M, N, K, ntaps,
                      shifts_host, bias, nullptr, nullptr, MYOLO_ACT_NONE, pf_w1, pf_blk, 0, mt, stream, BnBwd{}, hio);
  if (rc) return rc;
  stats_finalize_kernel<<<(N + 127) / 128, 128, 0, as_stream(stream)>>>(ws + 16, pivot, mean, var, N, 1.0 / (double)n_valid);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

// the same with the result stored as IEEE half and the statistics taken from the half-rounded values (so that the BN that
// follows normalises exactly the tensor it reads)
extern "C" int myolo_gemm_taps_hh_stats(const void* A, long long lda, const void* Bt, void* Ch, long long ldch, long long M, int N,
                                        int K, int ntaps, const int* shifts_host, const float* bias, int pf_w1, int pf_blk,
                                        const float* pivot, float* mean, float* var, double* ws, long long n_valid,
                                        myolo_stream stream) {
  MYOLO_CHECK_ARG(Ch && mean && var && ws && n_valid > 0 && N == 256 && pivot != mean);
  MYOLO_CHECK_ARG(myolo_gemm_taps_h_supported(lda, M, N, K, ntaps, shifts_host));
  MaskTail mt{};
  HalfIO hio{1, Ch, ldch, 1, nullptr, ws + 16, pivot};
  int rc = launch_win(reinterpret_cast<const float*>(A), lda, reinterpret_cast<const float*>(Bt), nullptr, N, M, N, K, ntaps,
                      shifts_host, bias, nullptr, nullptr, MYOLO_ACT_NONE, pf_w1, pf_blk, 0, mt, stream, BnBwd{}, hio);
  if (rc) return rc;
  stats_finalize_kernel<<<(N + 127) / 128, 128, 0, as_stream(stream)>>>(ws + 16, pivot, mean, var, N, 1.0 / (double)n_valid);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_deconv_mask_fwd_h(const void* a4, const void* kd, const float* bd, const float* w1, const float* b1,
                                       float* masks, const int* target_ids, float* y4, int n_roi, int H, int W, int Cmid,
                                       int NC, myolo_stream stream) {
  MYOLO_CHECK_ARG(a4 && kd && bd && w1 && b1 && masks && y4 && n_roi > 0 && H > 0 && W > 0);
  MYOLO_CHECK_ARG(myolo_deconv_mask_fwd_supported(Cmid, NC));
  MaskTail mt = make_tail(bd, w1, b1, masks, target_ids, y4, H, W, NC, 1);
  const long long M = (long long)n_roi * (H + 1) * (W + 1);
  HalfIO hio{1, nullptr, 0, 1, nullptr};
  return launch_win(reinterpret_cast<const float*>(a4), Cmid, reinterpret_cast<const float*>(kd), y4, 4 * Cmid, M, 4 * Cmid,
                    Cmid, 1, nullptr, nullptr, nullptr, nullptr, MYOLO_ACT_NONE, W + 1, (H + 1) * (W + 1), 0, mt, stream,
                    BnBwd{}, hio);
}

extern "C" int myolo_bn_epi_finalize_s(double* sums, const float* gamma, const float* var, float eps, float* dgamma,
                                       float* dbeta, float* dbias, int C, const float* unscale, myolo_stream stream);

extern "C" int myolo_gemm_taps_bnbwd_h(const void* A, long long lda, const void* Bt, float* C, void* Ch, long long ldc,
                                       long long M, int N, int K, int ntaps, const int* shifts_host, int pf_w1, int pf_blk,
                                       const void* a_out, const float* gamma, const float* beta, const float* var, float eps,
                                       int act, float* dgamma, float* dbeta, float* dbias, double* ws,
                                       const float* grad_unscale, myolo_stream stream) {
  MYOLO_CHECK_ARG(a_out && gamma && beta && var && dgamma && dbeta && ws && (C || Ch));
  MYOLO_CHECK_ARG(N == 256 && ldc == N && myolo_gemm_taps_h_supported(lda, M, N, K, ntaps, shifts_host));
  MaskTail mt{};
  BnBwd bnb{reinterpret_cast<const float*>(a_out), gamma, beta, var, ws + 16, eps};
  HalfIO hio{1, Ch, ldc, C ? 0 : 1, nullptr};
  int rc = launch_win(reinterpret_cast<const float*>(A), lda, reinterpret_cast<const float*>(Bt), C, ldc, M, N, K, ntaps,
                      shifts_host, nullptr, nullptr, nullptr, act, pf_w1, pf_blk, 0, mt, stream, bnb, hio);
  if (rc) return rc;
  return myolo_bn_epi_finalize_s(ws + 16, gamma, var, eps, dgamma, dbeta, dbias, N, grad_unscale, stream);
}

// myolo_gemm_taps_bnbwd_h without the finalisation: the column sums (sum g, sum g*xhat in the loss-scaled domain) stay in the
// BN workspace for myolo_bn_bwd_batch_fix_hh -- the backward of a BATCH-statistics BN whose first pass rides in this GEMM's
// epilogue.  var is the batch variance of the forward pass.
extern "C" int myolo_gemm_taps_bnbwd_sums_h(const void* A, long long lda, const void* Bt, void* Ch, long long ldc, long long M,
                                            int N, int K, int ntaps, const int* shifts_host, int pf_w1, int pf_blk,
                                            const void* a_out, const float* gamma, const float* beta, const float* var, float eps,
                                            int act, double* ws, myolo_stream stream) {
  MYOLO_CHECK_ARG(a_out && gamma && beta && var && ws && Ch);
  MYOLO_CHECK_ARG(N == 256 && ldc == N && myolo_gemm_taps_h_supported(lda, M, N, K, ntaps, shifts_host));
  MaskTail mt{};
  BnBwd bnb{reinterpret_cast<const float*>(a_out), gamma, beta, var, ws + 16, eps};
  HalfIO hio{1, Ch, ldc, 1, nullptr};
  return launch_win(reinterpret_cast<const float*>(A), lda, reinterpret_cast<const float*>(Bt), nullptr, ldc, M, N, K, ntaps,
                    shifts_host, nullptr, nullptr, nullptr, act, pf_w1, pf_blk, 0, mt, stream, bnb, hio);
}
