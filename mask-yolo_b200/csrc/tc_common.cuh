// tc_common.cuh -- PTX wrappers (mbarrier, TMA, tcgen05 alloc/mma/commit/ld), UMMA descriptor
// builders and the epilogue descriptor shared by the tcgen05 kernels of libmyolo_sm100.so.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace myolo {
namespace tc {

struct TapShifts {
  int s[32];
};

constexpr int BM = 128;  // UMMA M: rows (fwd) / A-channels (wgrad) per CTA
constexpr int BK = 32;   // fp32 per k-block = 128 bytes = one swizzle atom
constexpr int kThreads = 192;
constexpr long long kWatchdogCycles = 6000000000LL;  // ~3 s: turns a protocol bug into a trap, not a hang

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin == 256) t0 = clock64();
    if (spin > 256 && (spin & 255) == 0 && clock64() - t0 > kWatchdogCycles) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives on `bar` once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// split form of tmem_ld32: issue now, use the registers only after tmem_ld_wait(v) (which names them as in/out
// operands so the compiler cannot move a use above the wait)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
// layout: 2 = SWIZZLE_128B (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks; the only
// layout the hardware accepts for MN-major 32-bit operands).
// base_offset [49,52): phase of the swizzle pattern when the start address is not aligned to the
// pattern repeat (1024 B for the 128B swizzle) = (start_address >> 7) & 7.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint64_t layout = 2, uint64_t base_offset = 0) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | ((base_offset & 7ull) << 49) | (layout << 61);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10),
// a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 instruction descriptor: IEEE half operands (a_format = b_format = 0), fp32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ bool pf_valid(long long m, int pf_w1, int pf_blk) {
  if (pf_w1 <= 0) return true;
  const int r = (int)(m % pf_blk);
  return (r / pf_w1) >= 1 && (r % pf_w1) >= 1;
}

// The same test with the two divisions by the (run-time constant) tile geometry as multiply-high + shift: built once per
// thread at kernel start, used once per work item by every epilogue thread (the 64-bit div / mod chains above were ~10 %
// of the instructions of the mask-tail epilogue).  Rows are below 2^31 (checked on the host).
struct PfDiv {
  uint32_t mb, sb, mw, sw;   // magic multiplier / shift for pf_blk and pf_w1
  int w1, blk;
};
__device__ __forceinline__ void pf_magic(uint32_t d, uint32_t& mul, uint32_t& shr) {
  const uint32_t l = d > 1u ? 32u - (uint32_t)__clz((int)(d - 1u)) : 0u;
  mul = (uint32_t)((((unsigned long long)1 << 32) * (((unsigned long long)1 << l) - d)) / d + 1ull);
  shr = l;
}
__device__ __forceinline__ PfDiv make_pfdiv(int pf_w1, int pf_blk) {
  PfDiv d{0u, 0u, 0u, 0u, pf_w1, pf_blk};
  if (pf_w1 > 0) {
    pf_magic((uint32_t)pf_blk, d.mb, d.sb);
    pf_magic((uint32_t)pf_w1, d.mw, d.sw);
  }
  return d;
}
// row m of a padded-flat tensor -> tile (roi / image) index, line and column inside the tile; false for pad rows
__device__ __forceinline__ bool pf_decode(long long m, const PfDiv& d, int& tile, int& line, int& col) {
  if (d.w1 <= 0) {
    tile = line = col = 0;
    return true;
  }
  const uint32_t mm = (uint32_t)m;
  const uint32_t t = (__umulhi(d.mb, mm) + mm) >> d.sb;
  const uint32_t r = mm - t * (uint32_t)d.blk;
  const uint32_t ln = (__umulhi(d.mw, r) + r) >> d.sw;
  tile = (int)t;
  line = (int)ln;
  col = (int)(r - ln * (uint32_t)d.w1);
  return line >= 1 && col >= 1;
}
__device__ __forceinline__ bool pf_valid(long long m, const PfDiv& d) {
  int t, l, c;
  return pf_decode(m, d, t, l, c);
}

// x[c] = this lane's (row's) value of column c.  Returns, in lane l, the sum over the 32 rows of column l
// (butterfly transpose-reduce: 31 shuffles, no shared memory).
__device__ __forceinline__ float warp_colsum32(float (&x)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? x[i] : x[i + w];
      const float keep = up ? x[i + w] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return x[0];
}

struct Epi {
  const float* bias;
  const float* scale;
  const float* shift;
  int act, pf_w1, pf_blk, accumulate;
  // fused backward of a = act(BN_fixed_stats(.)) (conv_win kernel only): the GEMM result is d(a); the epilogue
  // reads a, writes d(pre-BN) = d(a)*act'(a)*gamma*rs and accumulates dbeta / dgamma column sums into bn_ws
  const float* bn_a;      // [M][N] activation the gradient belongs to (same row layout as C); nullptr = off
  const float* bn_gamma;
  const float* bn_beta;
  const float* bn_var;
  double* bn_ws;          // [2][N] fp64 sums (zero before, finalized + zeroed by myolo_gemm_taps_bnbwd)
  float bn_eps;
  // half-operand (kind::f16) variants of the conv_win kernel
  const void* bn_a_h;     // the same activation stored as IEEE half (replaces bn_a)
  int no_f32;             // do not store the fp32 result C
  int has_h;              // store the result as half through the second output map (the fp32 copy, when also
                          // stored, is the half-rounded value so both consumers see the same numbers)
  const float* acc_scale; // device scalar multiplied into the accumulator before the epilogue (nullable)
  // batch statistics of the stored result in the epilogue (tc_gemm_kernel only): per-column sum / sum of
  // squares over the valid rows, fp64 workspace of the BN family (bn.cu: zero before, zero after), the CTA that draws the
  // last ticket of its column tile writes mean / biased variance.  st_sums == nullptr: off.
  double* st_sums;        // [2][N]
  int* st_ticket;         // one counter per column tile
  float* st_mean;
  float* st_var;
  double st_inv;          // 1 / number of valid rows
  const float* st_pivot;  // conv_win kernel: per-column value subtracted before the sums are taken (nullable)
  int bn_tma;             // conv_win kernel, half flavour of the fused BN backward: the activation arrives by TMA through the
                          // (otherwise unused) fp32 output map, one 32x32 half tile per epilogue warp and chunk
};


// ---- cta_group::2 (CTA pair) variants.  The pair is a cluster of two CTAs; rank 0 (the leader)
// owns the "full" barriers and issues every MMA; both CTAs load their half of the operands.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// (default .release.cta semantics: the explicit .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR, which made
// every epilogue warp wait for its own global stores before handing the TMEM stage back -- 11 % of the deconv kernel's
// stall samples.  The hand-back orders tcgen05.ld, which tcgen05.fence::before_thread_sync already covers.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// cluster-scope release / acquire pair for data one CTA of the pair writes to its own shared memory with ordinary stores
// and the pair's tensor cores then read (the A operand of the tensor-core mask tail)
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin == 256) t0 = clock64();
    if (spin > 256 && (spin & 255) == 0 && clock64() - t0 > kWatchdogCycles) __trap();
  }
}
// TMA load whose completion is signalled on a barrier of the pair's LEADER CTA (cluster address)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives on the barrier at the same offset in BOTH CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(bar)
      : "memory");
}

__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

// fp32 -> IEEE half, round to nearest even, saturating to +-65504 (a gradient or activation that leaves the
// half range clamps instead of turning the whole tensor into inf/nan)
__device__ __forceinline__ uint16_t f2h_sat(float x) {
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float h2f(uint16_t h) {
  float r;
  asm("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"(h));
  return r;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) { return (uint32_t)f2h_sat(lo) | ((uint32_t)f2h_sat(hi) << 16); }

// ---- epilogue: 32x32 fp32 chunk (one TMEM lane quarter x 32 columns) -> global memory through a
// per-warp 4 KB staging buffer and a TMA tensor store.  A thread owns one output ROW; writing rows
// straight to global memory is a 1 KB-strided 16-byte scatter (measured: the stores cost as much as
// the whole tcgen05 main loop), so the warp lays the chunk out in the 128B-swizzled box layout and
// one lane issues cp.async.bulk.tensor (coalesced 128-byte rows, rows >= M clipped by the map).
// Rows that must stay zero (padded-flat pad rows) are stored as zeros.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// v[32] = accumulator columns n0..n0+31 of this thread's row; applies the epilogue and stages the row.
__device__ __forceinline__ void epi_stage_row(const float* v, const Epi& ep, int n0, bool valid, uint32_t sbuf, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n0 + 4 * j + e;
      float t = v[4 * j + e];
      if (ep.bias) t += __ldg(ep.bias + n);
      if (ep.scale) t = fmaf(t, __ldg(ep.scale + n), __ldg(ep.shift + n));
      o[e] = valid ? apply_act(t, ep.act) : 0.f;
    }
    const uint32_t addr = sbuf + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3])
                 : "memory");
  }
}

// 2D fp32 tensor map [rows][cols], row pitch `pitch` elements, box = 32 columns x box_rows rows, 128B swizzle
// (atom32: the 32-byte-chunk flavour for MN-major tf32 operands).  Cached per (pointer, shape).
int get_map(const float* p, long long rows, int cols, long long pitch, int box_rows, CUtensorMap* out, int atom32 = 0);
// 2D IEEE-half tensor map [rows][cols], row pitch `pitch` elements (multiple of 8).  box_cols = 64 -> 128B swizzle
// (tcgen05 operand tiles); box_cols = 32 -> 64B swizzle (epilogue store tiles of 32 columns).
int get_map_h(const void* p, long long rows, int cols, long long pitch, int box_rows, int box_cols, CUtensorMap* out);

}  // namespace tc
}  // namespace myolo
