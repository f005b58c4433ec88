// "Unified lane" building blocks shared by the fused edge kernels of round 2 (edge_fwd_ul.cu, ...).
//
// Why: the round-1 kernels ran ONE consumer group of 7 warps per SM (one thread per irrep channel, three
// different code paths for l = 0, 1, 2) because the filter rows + accumulators fill tensor memory: 12-33 %
// warps active, the l = 2 warp on the critical path, instruction-cache misses from the role-specific loops.
// Here the 576 filter rows of a channel slice are laid out so that EVERY accumulator lane carries the same
// amount of work and the same code:
//
//   lane L (0..127) of a consumer group owns
//     * the l = 0 channel L           : rows state / edge / scalar  -> row tiles 0, 1, 2 at lane L
//     * a three-component "piece"     : rows state / edge of one l > 0 channel -> row tiles 3, 4 at lane L
//         L <  64 : l = 1 channel L,        components m = 0, 1, 2
//         L <  96 : l = 2 channel L - 64,   components m = 0, 1, 2
//         L < 128 : l = 2 channel L - 96,   components m = 3, 4      (its two filter rows are duplicated
//                                                                      in the 32 spare lanes of tiles 3, 4)
//   => 10 FMAs per edge and lane, one code path for all four warps of a group, and a group is 4 warps, so
//   THREE groups (12 consumer warps) share one copy of the filter rows in tensor memory, each with its own
//   80-column accumulator buffer (240 + 3 x 80 = 480 of the 512 columns).
//
// The gathered node features are pre-multiplied and packed per lane ("pk" rows, pack kernels below), so the
// per-edge gather of a lane is two 128-bit loads: from the shared-memory window that a TMA bulk copy
// (cp.async.bulk, one elected thread) fills with the rows of a molecule tile, or straight from global memory
// (L2) when a tile is too large to stage.
#pragma once
#include "edge_mma.cuh"

namespace xeq {
namespace ul {

using namespace fm;

constexpr int G = 3;             // consumer groups per CTA
constexpr int GRP = 128;         // threads per group = accumulator lanes
constexpr int NCONS = G * GRP;   // 384
constexpr int ROWF = 1024;       // floats of one packed row: 2 planes x 128 lanes x 4
constexpr int ROWB = ROWF * 4;
constexpr int WH = 23;           // rows per window half (two halves: double-buffered tiles of <= 23 nodes)

// quad flags (a quad = 4 consecutive slots of one row)
enum : int {
  F_ROW_FIRST = 1, F_ROW_LAST = 2, F_TILE_FIRST = 4, F_TILE_LAST = 8, F_STAGED = 16, F_BUF = 32, F_PAR = 64, F_NOROW = 128
};
struct Quad {
  int node;
  int flags;
};

// ---- mbarrier / TMA helpers --------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait of a thread that has nothing else to do for a long time (window loader): back off between polls
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(256);
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ldg128(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// sin(x) for the Bessel terms sin(f_k d), 0 <= x <~ 100: two-term Cody-Waite reduction to [-pi, pi] (exact to ~1e-7 for
// these arguments), then the SFU approximation (absolute error ~2^-21 on the reduced range).  Six instructions
// instead of the ~20 of sinf(); the error is below that of the 3xTF32 contraction the values feed.
__device__ __forceinline__ float sin_reduced(float x) {
  const float n = rintf(x * 0.15915494309189535f);
  float r = fmaf(n, -6.2831854820251465f, x);
  r = fmaf(n, 1.7484555e-7f, r);
  return __sinf(r);
}

__device__ __forceinline__ void sincos_reduced(float x, float& sn, float& cs) {  // same reduction, both SFU functions
  const float n = rintf(x * 0.15915494309189535f);
  float r = fmaf(n, -6.2831854820251465f, x);
  r = fmaf(n, 1.7484555e-7f, r);
  sn = __sinf(r);
  cs = __cosf(r);
}

// ---- lane -> channel maps ----------------------------------------------------------------------------
// irrep index (0..M) of the piece of lane L in slice sl
template <int C, int M1>
__device__ __forceinline__ int piece_irrep(int L, int sl) {
  return (L < 64) ? C + sl * SL_M1 + L : C + M1 + sl * SL_M2 + ((L - 64) & 31);
}
// offsets into a cm-layout row [C | 3 x M1 | 5 x M2] of the (up to) three components of the piece of lane L
template <int C, int M1, int M2>
__device__ __forceinline__ void piece_offsets(int L, int sl, int (&off)[3], int& ncomp) {
  if (L < 64) {
    ncomp = 3;
#pragma unroll
    for (int m = 0; m < 3; ++m) off[m] = C + m * M1 + sl * SL_M1 + L;
  } else if (L < 96) {
    ncomp = 3;
#pragma unroll
    for (int m = 0; m < 3; ++m) off[m] = C + 3 * M1 + m * M2 + sl * SL_M2 + (L - 64);
  } else {
    ncomp = 2;
    off[0] = C + 3 * M1 + 3 * M2 + sl * SL_M2 + (L - 96);
    off[1] = C + 3 * M1 + 4 * M2 + sl * SL_M2 + (L - 96);
    off[2] = off[1];  // dummy (never stored)
  }
}
// piece type of a lane: selects the harmonics (l = 1 | l = 2, m = 0..2 | l = 2, m = 3, 4)
__device__ __forceinline__ int piece_type(int L) { return L < 64 ? 0 : (L < 96 ? 1 : 2); }

// filter rows of lane L -> tensor memory (hi / lo copies), five row tiles.  Warp-collective.
template <int C, int M1, int M2>
__device__ __forceinline__ void store_filter_rows(const float* __restrict__ W, const float* __restrict__ b, int L, int sl,
                                                  uint32_t lane_base) {
  constexpr int M = C + M1 + M2;
  const int q0 = sl * SL_C + L, qp = piece_irrep<C, M1>(L, sl);
  float row[NBP];
  load_wrow(W, b, q0, row);
  store_a_row(lane_base, 0, row);
  load_wrow(W, b, M + q0, row);
  store_a_row(lane_base, 1, row);
  load_wrow(W, b, 2 * M + q0, row);
  store_a_row(lane_base, 2, row);
  load_wrow(W, b, qp, row);
  store_a_row(lane_base, 3, row);
  load_wrow(W, b, M + qp, row);
  store_a_row(lane_base, 4, row);
  tmem_wait_st();
}

// ---- tile / row walk of one consumer group -----------------------------------------------------------
// tile_mode 1 (molecule tiles): the CTA owns tile blockIdx.x + k gridDim.x, its groups take the rows of the
//   tile round-robin (row = n0 + g + k G) and share the staged window of the tile;
// tile_mode 0 (edge-block tiles, large graphs): every group walks its own tiles (no window to share).
struct Walk {
  const int* __restrict__ tile_ptr;
  int n_tiles, tile, tstride, rphase, rstride, tile_mode;
  int n0, n1, staged_count;
  bool valid, staged, allow_stage = true;
  int buf, par;

  __device__ __forceinline__ void load() {
    for (;;) {
      valid = tile < n_tiles;
      if (!valid) return;
      n0 = tile_ptr[tile];
      n1 = tile_ptr[tile + 1];
      if (n0 < n1) break;
      tile += tstride;
    }
    staged = allow_stage && tile_mode == 1 && (n1 - n0) <= WH;
    if (staged) {
      buf = staged_count & 1;
      par = (staged_count >> 1) & 1;
      ++staged_count;
    }
  }
  __device__ __forceinline__ void init(const xeq_graph_t& g, const int* tp, int nt, int grp) {
    tile_ptr = tp; n_tiles = nt; tile_mode = g.tile_mode;
    if (tile_mode == 1) { tile = blockIdx.x; tstride = gridDim.x; rphase = grp; rstride = G; }
    else { tile = blockIdx.x * G + grp; tstride = gridDim.x * G; rphase = 0; rstride = 1; }
    staged_count = 0;
    load();
  }
  __device__ __forceinline__ void next() {
    tile += tstride;
    load();
  }
};

// Sum 16 per-thread values over the 32 lanes of a warp: four halving exchanges (each keeps half of the values),
// then one butterfly.  Lane l returns the total of v[l >> 1].  Fixed tree -> bitwise reproducible.
__device__ __forceinline__ float warp_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int st = 0; st < 4; ++st) {
    const int half = 16 >> (st + 1), off = 16 >> st;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

}  // namespace ul
}  // namespace xeq
