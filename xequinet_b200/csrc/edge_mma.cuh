// tcgen05 building blocks of the fused edge kernels (edge_ul.cuh and the edge_*_ul.cu kernels build on these).
//
// The filter contraction  w[h, e] = sum_k [b | W_rbf][h, k] * psi_k(d_e)  (nn/xpainn.py:140, K = 21 -> 24)
// runs on the tensor cores as D[h, e] = A[h, k] B[e, k]^T with
//   A = the filter rows, resident in TENSOR MEMORY for the whole kernel (tcgen05.st once per CTA, 3xTF32 split: hi and
//       lo copies), so the rows leave the register file and need no shared memory;
//   B = the per-edge radial terms of one chunk, written directly in the canonical K-major SWIZZLE_128B shared-memory
//       layout (hi and lo tiles);
//   D = fp32 accumulators in TMEM with lane = filter row: warp w reads lanes 32 (w % 4) .. +31 with tcgen05.ld.
// Row -> (tile, lane): edge_ul.cuh ("unified lanes").
// Products are evaluated as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (fp32-level accuracy, see node_gemm.cu).
// Measured on B200 (profiles/r01_ts_mma_probe.log): 45 MMAs (one filter output of a 32-edge chunk)
// = 837 cycles, 26 cycles per edge.
#pragma once
#include "edge_geo.cuh"

namespace xeq {
namespace fm {

// One CTA handles a SLICE of 128 l=0 + 64 l=1 + 32 l=2 channels (576 filter rows = 5 row tiles); wider models
// (256x0e + 128x1o + 64x2e: c4) run C / 128 slices per tile of edges (blockIdx.y), each with the filter rows of
// its own channels in TMEM -- the geometry of an edge is recomputed per slice, the feature traffic is not.
constexpr int SL_C = 128, SL_M1 = 64, SL_M2 = 32, SL_M = SL_C + SL_M1 + SL_M2;
constexpr int TILES = 5;
constexpr int A_HI = 0;                  // TMEM columns [0, 120): hi parts, tile-major, 24 per tile
constexpr int A_LO = TILES * NBP;        // [120, 240): lo parts
constexpr int D_COL = 2 * TILES * NBP;   // [240, ...): accumulators, (output, tile)-major, T columns each
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

// Same split on the integer / FMA pipes for the per-edge hot paths (cvt.rna runs on the quarter-rate
// conversion pipe): hi = x rounded to 10 mantissa bits by an integer add, lo = x - hi exactly; the
// tensor core ignores the 13 low mantissa bits of lo, a relative error of 2^-21 of x.
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  // lo = x - hi is exact (<= 13 significant bits); round it to the 11 the tensor core keeps instead of letting the
  // hardware truncate: hi + lo then represents x to 2^-23 (round to nearest) instead of 2^-22
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (same encoding as node_gemm.cu)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of element (row r, k) inside a K-major SWIZZLE_128B tile (128-byte rows, 8-row groups)
__device__ __forceinline__ uint32_t b_off(int r, int k) {
  return (uint32_t)(r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4);
}

// D[tmem_d] (+)= A[tmem_a] * B[smem]^T, one 128 x N x 8 tf32 step
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor: D = f32, A = B = tf32, K-major, M = 128, N = n
__device__ __forceinline__ uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {  // implies tcgen05.fence::before_thread_sync
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {  // one lane of a converged warp
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3])
               : "r"(taddr)
               : "memory");
}
// The loaded registers may only be read after this wait.  pin() re-defines them in a volatile asm placed
// after the wait, so no use can be scheduled above it.
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void pin(float (&r)[4]) { asm volatile("" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3])); }

// TMEM allocation by warp 0 of the CTA; every thread gets the base address.  Contains a __syncthreads.
__device__ __forceinline__ uint32_t tmem_setup(uint32_t* slot, uint64_t* bar, int n_bars) {
  const int t = threadIdx.x;
  if (t == 0) {
    for (int i = 0; i < n_bars; ++i) mbar_init(smem_u32(bar + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(slot);
}
__device__ __forceinline__ void tmem_teardown(uint32_t tmem) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

// Write one filter row (24 values of [b | W_rbf | 0 0 0], in registers) as row `lane` of A tile `tile`,
// hi and lo copies.  Warp-collective (all 32 lanes write their own row).  lane_base = tmem + (quarter << 16).
__device__ __forceinline__ void store_a_row(uint32_t lane_base, int tile, const float (&row)[NBP]) {
#pragma unroll
  for (int c8 = 0; c8 < NBP / 8; ++c8) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_tf32(row[c8 * 8 + j], hi[j], lo[j]);
    tmem_st8(lane_base + A_HI + tile * NBP + c8 * 8, hi);
    tmem_st8(lane_base + A_LO + tile * NBP + c8 * 8, lo);
  }
}

}  // namespace fm
}  // namespace xeq
