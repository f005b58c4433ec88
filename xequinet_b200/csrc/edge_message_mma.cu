// Round-1 mapping (one thread per irrep channel, one consumer group per CTA) of the kernels that have not moved to the
// unified-lane design yet: the JVP pass of K2bb (center_mma_kernel<JVP>) and the weight-gradient kernels.  The forward,
// first-order and second-order neighbor passes live in edge_fwd_ul.cu / edge_bwd_ul.cu / edge_bwd2_ul.cu.
//
// K2 / K2b / K2bb with the filter contraction on the tensor cores (tcgen05, operands in TMEM).
//
// Same three kernel families, thread <-> channel mapping, CSR walk, shared-memory row window and
// fixed-order reductions as edge_message.cu; what changes is where  w = [b | W_rbf] . psi(d)  (and
// its d-derivatives) is evaluated.  In the SIMT kernels that K = 21 dot product is 90 % of the
// instruction stream (63..189 FMAs per thread and edge); here it is a 3xTF32 tcgen05.mma per chunk
// (edge_mma.cuh: filter rows resident in TMEM, radial terms written by the geometry stage straight
// into SWIZZLE_128B tiles, accumulators read back with tcgen05.ld by the thread that owns the row).
//
// Pipeline of one CTA, iteration c (one __syncthreads per chunk, as before):
//   elected thread : MMAs of chunk c        (B tiles of stage c & 1)         -> commit -> mbarrier
//   all threads    : geometry of chunk c+2, radial terms of chunk c+1 -> B tiles of stage (c+1) & 1
//   all threads    : wait on the mbarrier, read the accumulators 4 edges at a time, message math
// so the tensor-core latency is covered by the geometry stages.  The accumulators are single
// buffered: TMEM holds 240 columns of filter rows + up to 240 columns of accumulators.
//
// Instantiated for the default widths (128x0e + 64x1o + 32x2e: 576 filter rows = 5 row tiles); other
// widths run the SIMT kernels of edge_message.cu.
#include <type_traits>

#include "edge_mma.cuh"

#ifdef XEQ_TRACE
// Debug timeline (scratch/trace_nbr.py): cycle stamps of every warp of CTA 0 at the per-chunk barrier.
__device__ long long g_trace_nbr[8][2][256];
#define XEQ_TRACE_STAMP(which, c)                                                                           \
  do {                                                                                                      \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (c) < 256) g_trace_nbr[threadIdx.x >> 5][which][c] = clock64(); \
  } while (0)
#else
#define XEQ_TRACE_STAMP(which, c) do { } while (0)
#endif

namespace xeq {

using namespace fm;

// ==========================================================================================
// center kernel (forward message, JVP half of the double backward)
// ==========================================================================================
template <bool JVP> struct CenterMma {
  static constexpr int TC = JVP ? 16 : 32;   // edges per chunk = MMA N
  static constexpr int NOUT = JVP ? 2 : 1;   // filter outputs: w (, dw)
  static constexpr int STAGE = NOUT * 2 * TC * 128;
  static constexpr int WIN = 21;              // rows of the shared-memory window (aspirin: 21 atoms)
};

template <bool JVP>
struct CenterMmaSmem {
  GeoA<CenterMma<JVP>::TC, false, JVP> a[3];
  ChunkDesc desc[8];  // ring written by the producer warp, ahead of the geometry it describes
  uint64_t bar;
  uint32_t slot;
};

template <int L, int C, int M1, int M2, bool JVP>
__device__ __forceinline__ void center_mma_role(const CenterArgs& A, CenterMmaSmem<JVP>& sm, const uint32_t tmem,
                                                const uint32_t tiles, const uint32_t win_base) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M, NC = 2 * L + 1;
  constexpr int THREADS = SL_M;
  constexpr int TC = CenterMma<JVP>::TC, NOUT = CenterMma<JVP>::NOUT, STAGE = CenterMma<JVP>::STAGE;
  constexpr int TS = (L == 0) ? 0 : 3, TE = (L == 0) ? 1 : 4, TX = 2;  // row tiles of this thread's filter rows
  const int t = threadIdx.x, warp = t >> 5;
  const int q = slice_channel<L, C, M1>(t, blockIdx.y);  // irrep channel of this thread
  const int vbase = (L == 0) ? q : (L == 1 ? C + (q - C) : C + 3 * M1 + (q - C - M1));
  constexpr int vstride = (L == 0) ? 0 : (L == 1 ? M1 : M2);
  const xeq_graph_t& g = A.geo.g;
  const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t bar = smem_u32(&sm.bar);

  {  // filter rows -> TMEM (once per CTA)
    float row[NBP];
    load_wrow(A.W, A.b, q, row);
    store_a_row(lane_base, TS, row);
    load_wrow(A.W, A.b, M + q, row);
    store_a_row(lane_base, TE, row);
    if (L == 0) {
      load_wrow(A.W, A.b, 2 * M + q, row);
      store_a_row(lane_base, TX, row);
    }
    tmem_wait_st();
  }

  CenterThread<float, L, NK_> th;  // accumulators only (the filter rows live in TMEM)
  th.reset();

  struct Gathered {
    float ss, se, sx, v[NC], sds, sde, sdx, vd[NC];
  };
  auto gather = [&](int j, Gathered& o) {
    const float* sj = A.s + (size_t)j * H;
    o.ss = sj[q];
    o.se = sj[M + q];
    o.sx = (L == 0) ? sj[2 * M + q] : 0.f;
    const float* vj = A.v + (size_t)j * D + vbase;
#pragma unroll
    for (int m = 0; m < NC; ++m) o.v[m] = vj[m * vstride];
    if (JVP) {
      o.sds = o.sde = o.sdx = 0.f;
      if (A.a_s) {
        const float* aj = A.a_s + (size_t)j * H;
        o.sds = aj[q];
        o.sde = aj[M + q];
        if (L == 0) o.sdx = aj[2 * M + q];
      }
#pragma unroll
      for (int m = 0; m < NC; ++m) o.vd[m] = A.a_v ? A.a_v[(size_t)j * D + vbase + m * vstride] : 0.f;
    }
  };

  // staged window: [row][column][thread-of-role] floats, role regions side by side (as in edge_message.cu)
  constexpr int WMAX = CenterMma<JVP>::WIN;
  constexpr int ROWF = (SL_C * 4 + SL_M1 * 5 + SL_M2 * 7) * (JVP ? 2 : 1);
  constexpr int NTHR = (L == 0) ? SL_C : (L == 1 ? SL_M1 : SL_M2);
  constexpr int ROLE_OFF = (L == 0 ? 0 : (L == 1 ? SL_C * 4 : SL_C * 4 + SL_M1 * 5)) * (JVP ? 2 : 1);
  const int tt = (L == 0) ? t : (L == 1 ? t - SL_C : t - SL_C - SL_M1);
  const uint32_t win0 = win_base + 4u * (ROLE_OFF + tt);
  bool staged = false;
  int win_lo = 0;
  auto stage_window = [&](int n0, int n1) {
#pragma unroll 2
    for (int j = n0; j < n1; ++j) {
      const uint32_t a = win0 + 4u * (uint32_t)((j - n0) * ROWF);
      Gathered gc;
      gather(j, gc);
      int c = 0;
      sts_f32(a + 4u * (uint32_t)(NTHR * c++), gc.ss);
      sts_f32(a + 4u * (uint32_t)(NTHR * c++), gc.se);
      if (L == 0) sts_f32(a + 4u * (uint32_t)(NTHR * c++), gc.sx);
#pragma unroll
      for (int m = 0; m < NC; ++m) sts_f32(a + 4u * (uint32_t)(NTHR * c++), gc.v[m]);
      if (JVP) {
        sts_f32(a + 4u * (uint32_t)(NTHR * c++), gc.sds);
        sts_f32(a + 4u * (uint32_t)(NTHR * c++), gc.sde);
        if (L == 0) sts_f32(a + 4u * (uint32_t)(NTHR * c++), gc.sdx);
#pragma unroll
        for (int m = 0; m < NC; ++m) sts_f32(a + 4u * (uint32_t)(NTHR * c++), gc.vd[m]);
      }
    }
  };
  auto gather_window = [&](int j, Gathered& o) {
    const uint32_t a = win0 + 4u * (uint32_t)((j - win_lo) * ROWF);
    int c = 0;
    o.ss = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
    o.se = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
    o.sx = (L == 0) ? lds_f32(a + 4u * (uint32_t)(NTHR * c++)) : 0.f;
#pragma unroll
    for (int m = 0; m < NC; ++m) o.v[m] = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
    if (JVP) {
      o.sds = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
      o.sde = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
      o.sdx = (L == 0) ? lds_f32(a + 4u * (uint32_t)(NTHR * c++)) : 0.f;
#pragma unroll
      for (int m = 0; m < NC; ++m) o.vd[m] = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
    }
  };

  float base_x = 0.f, base_V[NC];
#pragma unroll
  for (int m = 0; m < NC; ++m) base_V[m] = 0.f;

  // pipeline prologue (the producer warp has described and measured chunks 0 and 1)
  __syncthreads();
  ChunkDesc d0 = sm.desc[0], d1 = sm.desc[1];
  if (d0.cnt > 0) geo_stage_b<TC, THREADS, NOUT>(A.geo, d0.cnt, sm.a[0], tiles, threadIdx.x);
  proxy_fence();
  tc_fence_before();
  __syncthreads();

  uint32_t phase = 0;
  for (int c = 0; d0.cnt >= 0; ++c) {
    const bool has = d0.cnt > 0;
    // radial terms of the next chunk -> B tiles, in the shadow of this chunk's MMAs
    if (d1.cnt > 0) {
      geo_stage_b<TC, THREADS, NOUT>(A.geo, d1.cnt, sm.a[(c + 1) % 3], tiles + (uint32_t)((c + 1) & 1) * STAGE, threadIdx.x);
      proxy_fence();
    }

    const GeoA<TC, false, JVP>& sa = sm.a[c % 3];
    const int cnt = d0.cnt, node = d0.owner;
    if (d0.first) {
      staged = WMAX > 0 && g.tile_mode == 1 && (d0.n1 - d0.n0) <= WMAX;
      win_lo = d0.n0;
      if (staged) stage_window(d0.n0, d0.n1);
    }
    if (d0.rfirst) {  // residual row of the node: requested now, consumed when its row ends
#pragma unroll
      for (int m = 0; m < NC; ++m) base_V[m] = A.V_in ? A.V_in[(size_t)node * D + vbase + m * vstride] : 0.f;
      if (L == 0) base_x = A.x_in ? A.x_in[(size_t)node * C + q] : 0.f;
      th.reset();
    }
    if (has) {
      mbar_wait(bar, phase);
      phase ^= 1u;
      tc_fence_after();
    }
    // ---- the edges of this piece of the row, four at a time, no row switch and no bounds checks inside: the
    // slots past cnt repeat the last edge with exactly zero filter values (geo_stage_a1 / geo_stage_b)
    const uint32_t dbase = lane_base + D_COL;
    auto run_chunk = [&](auto staged_c) {
      constexpr bool ST = decltype(staged_c)::value;
#pragma unroll 1
      for (int g0 = 0; g0 < cnt; g0 += 4) {
        float ws[4], we[4], wx[4] = {0.f, 0.f, 0.f, 0.f}, dws[4] = {0.f, 0.f, 0.f, 0.f}, dwe[4] = {0.f, 0.f, 0.f, 0.f},
                            dwx[4] = {0.f, 0.f, 0.f, 0.f};
        tmem_ld4(dbase + TS * NOUT * TC + g0, ws);
        tmem_ld4(dbase + TE * NOUT * TC + g0, we);
        if (L == 0) tmem_ld4(dbase + TX * NOUT * TC + g0, wx);
        if (JVP) {
          tmem_ld4(dbase + TS * NOUT * TC + TC + g0, dws);
          tmem_ld4(dbase + TE * NOUT * TC + TC + g0, dwe);
          if (L == 0) tmem_ld4(dbase + TX * NOUT * TC + TC + g0, dwx);
        }
        Gathered gc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (ST) gather_window(sa.gat[g0 + j], gc[j]);
          else gather(sa.gat[g0 + j], gc[j]);
        }
        tmem_wait_ld();
        pin(ws); pin(we);
        if (L == 0) pin(wx);
        if (JVP) {
          pin(dws); pin(dwe);
          if (L == 0) pin(dwx);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = g0 + j;
          float Yl[8], Yd[8];
          load_rows8<L, 1>(sa.Y[e], Yl);  // 128-bit loads of the entries this irrep type uses
          if (JVP) load_rows8<L, 1>(sa.Ydot[e], Yd);
          if (!JVP) {
            th.fwd_w(ws[j], we[j], wx[j], Yl, gc[j].ss, gc[j].se, gc[j].sx, gc[j].v);
          } else {
            th.jvp_w(ws[j], we[j], wx[j], dws[j], dwe[j], dwx[j], Yl, Yd, sa.ddot[e],
                     gc[j].ss, gc[j].se, gc[j].sx, gc[j].v, gc[j].sds, gc[j].sde, gc[j].sdx, gc[j].vd);
          }
        }
      }
    };
    if (staged) run_chunk(std::true_type{});
    else run_chunk(std::false_type{});
    if (d0.rlast) {
#pragma unroll
      for (int m = 0; m < NC; ++m) A.V_out[(size_t)node * D + vbase + m * vstride] = base_V[m] + th.accV[m];
      if (L == 0) A.x_out[(size_t)node * C + q] = base_x + th.accx;
    }
    tc_fence_before();
    __syncthreads();
    d0 = d1;
    d1 = sm.desc[(c + 2) & 7];
  }
}

// Producer warp: walks the chunk stream one chunk ahead of the consumers' radial stage -- descriptor and
// per-edge geometry (one lane per edge) of chunk c+2 -- and issues the MMAs of chunk c.
template <int C, int M1, int M2, bool JVP>
__device__ __forceinline__ void center_mma_producer(const CenterArgs& A, CenterMmaSmem<JVP>& sm, const uint32_t tmem,
                                                    const uint32_t tiles) {
  constexpr int TC = CenterMma<JVP>::TC, NOUT = CenterMma<JVP>::NOUT, STAGE = CenterMma<JVP>::STAGE;
  const int lane = threadIdx.x & 31;
  const xeq_graph_t& g = A.geo.g;
  const uint32_t bar = smem_u32(&sm.bar);
  RowCursor<TC> cur_it;
  cur_it.init(g.rowptr, g.tile_ptr, g.n_tiles);
  GeoPipe<TC, false, false, JVP> gp;
  gp.init();
  auto step = [&](int c) {  // geometry pipeline of iteration c: C(c+2) -> shared memory, B(c+3), A(c+4)
    if (c + 2 >= 0) gp.stage_c(A.geo, sm.a[(c + 2) % 3], lane);
    if (c + 3 >= 0) gp.stage_b(A.geo, lane);
    const ChunkDesc d = cur_it.next();
    if (lane == 0) sm.desc[(c + 4) & 7] = d;
    gp.stage_a(A.geo, d, lane);
    __syncwarp();
  };
  for (int c = -4; c < 0; ++c) step(c);  // fill: geometry of chunks 0, 1 in shared memory, 2 and 3 in flight
  __syncthreads();
  __syncthreads();  // the consumers have written the B tiles of chunk 0
  for (int c = 0; sm.desc[c & 7].cnt >= 0; ++c) {
    tc_fence_after();
    if (sm.desc[c & 7].cnt > 0) {
      if (elect_one()) {
        issue_chunk<TC, NOUT>(tmem, tiles + (uint32_t)(c & 1) * STAGE);
        umma_commit(bar);
      }
      __syncwarp();
    }
    step(c);
    __syncthreads();
  }
}

template <int C, int M1, int M2, bool JVP>
__global__ void __launch_bounds__(SL_M + 32, 1) center_mma_kernel(const CenterArgs A) {
  static_assert(C % SL_C == 0 && M1 == C / 2 && M2 == C / 4, "channel slices of edge_mma.cuh");
  __shared__ CenterMmaSmem<JVP> sm;
  pdl_trigger();
  const uint32_t tmem = tmem_setup(&sm.slot, &sm.bar, 1);
  pdl_wait();  // setup overlapped the previous kernel's tail
  const uint32_t tiles = (smem_u32(xeq_dyn_smem) + 1023u) & ~1023u;
  const uint32_t win_base = tiles + 2u * CenterMma<JVP>::STAGE;
  const int t = threadIdx.x;
  if (t < SL_C) center_mma_role<0, C, M1, M2, JVP>(A, sm, tmem, tiles, win_base);
  else if (t < SL_C + SL_M1) center_mma_role<1, C, M1, M2, JVP>(A, sm, tmem, tiles, win_base);
  else if (t < SL_M) center_mma_role<2, C, M1, M2, JVP>(A, sm, tmem, tiles, win_base);
  else center_mma_producer<C, M1, M2, JVP>(A, sm, tmem, tiles);
  tmem_teardown(tmem);
}

// ==========================================================================================
// weight gradients (K2b-w / K2bb-w): a GEMM over the edges, accumulated in TMEM for the whole CTA
// ==========================================================================================
//   GW[h, k] = sum_e pw[h, e] psi_k(e)          GF[h, k] = sum_e pw[h, e] xi_k(e)          (first order)
//   GW[h, k] = sum_e alpha[h, e] psi_k(e) + (beta ddot)[h, e] dpsi_k(e),  GF likewise with xi, dxi   (second)
// D[h, n] (n < 24: GW, n >= 24: GF) = A[h, e] B[n, e]^T with the edges as the K dimension: the threads
// produce their rows of A = pw per chunk (registers -> TMEM, 3xTF32 split), the radial stage writes
// B = [psi | xi] (second order: edges 0..7 | the d-derivatives of the same 8 edges) transposed into the
// swizzled tiles, and the 5 x 48 accumulator columns stay in TMEM until the CTA has walked all its rows
// -- the SIMT kernel keeps 42..84 accumulators per thread in registers and re-reads psi / xi rows from
// shared memory for every edge.  Per-CTA partials are reduced in fixed order by wgrad_reduce_kernel.
template <int ORDER> struct WgradMma {
  static constexpr int KE = (ORDER == 2) ? 8 : 16;  // edges per chunk (MMA K = 16 either way)
  static constexpr int NB = 2 * NBP;                // accumulator columns per row tile
  static constexpr int A0 = TILES * NB;             // TMEM: accumulators [0, 240), A operand [240, 400)
  static constexpr int BT = NB * 128;               // bytes of one B tile (48 rows x 128 B, K = 16 used)
  static constexpr int STAGE = 2 * BT;              // hi + lo
  static constexpr int WIN = 24;                    // rows of the shared-memory window (gV | gx rows)
};

template <int ORDER>
struct WgradMmaSmem {
  GeoA<WgradMma<ORDER>::KE, false, ORDER == 2> a[3];
  ChunkDesc desc[8];
  uint64_t bar;
  uint32_t slot;
};

// radial terms of a chunk, transposed: row n = k (psi_k) / 24 + k (xi_k), column = edge slot; dead slots = 0
template <int ORDER, int THREADS>
__device__ __noinline__ void geo_stage_bt(const GeoArgs& A, int cnt, const GeoA<WgradMma<ORDER>::KE, false, ORDER == 2>& sa,
                                          uint32_t tiles, const int t /* dense index among the THREADS writers */) {
  constexpr int KE = WgradMma<ORDER>::KE, BT = WgradMma<ORDER>::BT;
  const float c0 = sqrtf(2.f / A.rc);
  for (int idx = t; idx < NBP * KE; idx += THREADS) {
    const int k = idx / KE, ee = idx - k * KE;
    float psi = 0.f, xi = 0.f, dpsi = 0.f, dxi = 0.f;
    if (ee < cnt) {
      if (k == 0) {
        psi = sa.chi[ee][0];
        dpsi = sa.chi[ee][1];
      } else if (k <= NB_) {
        Cutoff<float> c;
        c.chi = sa.chi[ee][0]; c.dchi = sa.chi[ee][1]; c.ddchi = sa.chi[ee][2];
        const Radial<float> rr = radial_term_c0(sa.d[ee], A.freq[k - 1], c0, c);
        psi = rr.psi; xi = rr.xi; dpsi = rr.dpsi; dxi = rr.dxi;
      }
    }
    auto put = [&](int n, int kk, float val) {
      uint32_t hi, lo;
      split_fast(val, hi, lo);
      const uint32_t off = b_off(n, kk);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(tiles + off), "r"(hi) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(tiles + BT + off), "r"(lo) : "memory");
    };
    put(k, ee, psi);
    put(NBP + k, ee, xi);
    if (ORDER == 2) {
      put(k, KE + ee, dpsi);
      put(NBP + k, KE + ee, dxi);
    }
  }
}

template <int ORDER>
__device__ __forceinline__ void issue_wgrad(uint32_t tmem, uint32_t tiles, bool first) {
  constexpr int NB = WgradMma<ORDER>::NB, A0 = WgradMma<ORDER>::A0, BT = WgradMma<ORDER>::BT;
  const uint32_t idesc = idesc_tf32(NB);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const uint64_t db_hi = smem_desc(tiles + ks * 32), db_lo = smem_desc(tiles + BT + ks * 32);
#pragma unroll
    for (int tile = 0; tile < TILES; ++tile) {
      const uint32_t d = tmem + tile * NB;
      const uint32_t a_hi = tmem + A0 + tile * 16 + ks * 8, a_lo = a_hi + TILES * 16;
      mma_ts(d, a_lo, db_hi, idesc, (first && ks == 0) ? 0u : 1u);
      mma_ts(d, a_hi, db_lo, idesc, 1u);
      mma_ts(d, a_hi, db_hi, idesc, 1u);
    }
  }
}

template <int L, int C, int M1, int M2, int ORDER>
__device__ __forceinline__ void wgrad_mma_role(const NeighborArgs& A, WgradMmaSmem<ORDER>& sm, const uint32_t tmem,
                                               const uint32_t tiles, const uint32_t win_base) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M, NC = 2 * L + 1;
  constexpr int THREADS = 2 * SL_M;  // two consumer groups write the radial tiles together
  constexpr bool SECOND = ORDER == 2;
  constexpr int KE = WgradMma<ORDER>::KE, NB = WgradMma<ORDER>::NB, A0 = WgradMma<ORDER>::A0, STAGE = WgradMma<ORDER>::STAGE;
  constexpr int TS = (L == 0) ? 0 : 3, TE = (L == 0) ? 1 : 4, TX = 2;
  constexpr int NROW = (L == 0) ? 3 : 2;
  // Two consumer groups (threads 0-223 and 256-479; warp 7 is the producer) own the same channels and split the
  // 4-edge groups of every chunk between them: twice the warps for the latency-bound per-edge work with the same
  // tensor-memory footprint (their rows of the A operand land in disjoint K slots).
  const int grp = (threadIdx.x >= SL_M + 32) ? 1 : 0;
  const int t = threadIdx.x - grp * (SL_M + 32), warp = t >> 5;  // index inside the group
  const int t2 = t + grp * SL_M;                                  // dense index over both groups
  const int q = slice_channel<L, C, M1>(t, blockIdx.y);  // irrep channel of this thread
  const int vbase = (L == 0) ? q : (L == 1 ? C + (q - C) : C + 3 * M1 + (q - C - M1));
  constexpr int vstride = (L == 0) ? 0 : (L == 1 ? M1 : M2);
  const xeq_graph_t& g = A.geo.g;
  const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t bar = smem_u32(&sm.bar);

  NeighborThread<float, L, ROLE_STATE, false, NK_, false> st;
  NeighborThread<float, L, ROLE_EDGE, false, NK_, false> ed;
  NeighborThread<float, 0, ROLE_SCALAR, false, NK_, false> sc;
  st.s = st.sd = ed.s = ed.sd = sc.s = sc.sd = 0.f;
#pragma unroll
  for (int m = 0; m < NC; ++m) st.v[m] = st.vd[m] = 0.f;

  auto begin_node = [&](int j) {
    const float* sj = A.s + (size_t)j * H;
    st.s = sj[q];
    ed.s = sj[M + q];
    if (L == 0) sc.s = sj[2 * M + q];
    if (SECOND) {
      st.sd = ed.sd = sc.sd = 0.f;
      if (A.a_s) {
        const float* aj = A.a_s + (size_t)j * H;
        st.sd = aj[q];
        ed.sd = aj[M + q];
        if (L == 0) sc.sd = aj[2 * M + q];
      }
    }
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      st.v[m] = A.v[(size_t)j * D + vbase + m * vstride];
      if (SECOND) st.vd[m] = A.a_v ? A.a_v[(size_t)j * D + vbase + m * vstride] : 0.f;
    }
  };
  struct Gathered {
    float g[NC], gx;
  };
  auto gather = [&](int i, Gathered& o) {
#pragma unroll
    for (int m = 0; m < NC; ++m) o.g[m] = A.gV[(size_t)i * D + vbase + m * vstride];
    o.gx = (L == 0) ? A.gx[(size_t)i * C + q] : 0.f;
  };
  // window of gV / gx rows (layout of edge_message.cu's neighbor kernels)
  constexpr int WMAX = WgradMma<ORDER>::WIN;
  constexpr int ROWF = SL_C * 2 + SL_M1 * 3 + SL_M2 * 5;
  constexpr int NTHR = (L == 0) ? SL_C : (L == 1 ? SL_M1 : SL_M2);
  constexpr int ROLE_OFF = (L == 0) ? 0 : (L == 1 ? SL_C * 2 : SL_C * 2 + SL_M1 * 3);
  const int tt = (L == 0) ? t : (L == 1 ? t - SL_C : t - SL_C - SL_M1);
  const uint32_t win0 = win_base + (uint32_t)grp * (4u * WMAX * ROWF) + 4u * (ROLE_OFF + tt);  // one window per group
  bool staged = false;
  int win_lo = 0;
  auto stage_window = [&](int n0, int n1) {
#pragma unroll 4
    for (int i = n0; i < n1; ++i) {
      const uint32_t a = win0 + 4u * (uint32_t)((i - n0) * ROWF);
      Gathered gc;
      gather(i, gc);
#pragma unroll
      for (int m = 0; m < NC; ++m) sts_f32(a + 4u * (uint32_t)(NTHR * m), gc.g[m]);
      if (L == 0) sts_f32(a + 4u * (uint32_t)(NTHR * NC), gc.gx);
    }
  };
  auto gather_window = [&](int i, Gathered& o) {
    const uint32_t a = win0 + 4u * (uint32_t)((i - win_lo) * ROWF);
#pragma unroll
    for (int m = 0; m < NC; ++m) o.g[m] = lds_f32(a + 4u * (uint32_t)(NTHR * m));
    o.gx = (L == 0) ? lds_f32(a + 4u * (uint32_t)(NTHR * NC)) : 0.f;
  };

  __syncthreads();  // the producer warp has described and measured chunk 0
  ChunkDesc d0 = sm.desc[0];
  uint32_t phase = 0;
  bool pending = false, any = false;
  for (int c = 0; d0.cnt >= 0; ++c) {
    const bool has = d0.cnt > 0;
    const GeoA<KE, false, SECOND>& sa = sm.a[c % 3];
    const int cnt = d0.cnt;
    if (d0.first) {
      staged = WMAX > 0 && g.tile_mode == 1 && (d0.n1 - d0.n0) <= WMAX;
      win_lo = d0.n0;
      if (staged) stage_window(d0.n0, d0.n1);
    }
    if (d0.rfirst && has) begin_node(d0.owner);
    if (has) {
      // radial terms of this chunk, transposed -> B tiles (stage c & 1 was last read by the MMAs of chunk c-2);
      // runs in the shadow of the previous chunk's MMAs
      geo_stage_bt<ORDER, THREADS>(A.geo, cnt, sa, tiles + (uint32_t)(c & 1) * STAGE, t2);
      proxy_fence();
    }
    if (pending) {  // the MMAs of the previous chunk have read the A operand
      mbar_wait(bar, phase);
      phase ^= 1u;
      pending = false;
    }
    if (has) {
      tc_fence_after();
      // rows of the A operand, four edge slots at a time: registers -> TMEM (hi and lo); slots past cnt = 0
      const uint32_t col_s = lane_base + A0 + TS * 16, col_e = lane_base + A0 + TE * 16, col_x = lane_base + A0 + TX * 16;
#pragma unroll 1
      for (int j0 = 4 * grp; j0 < KE; j0 += 8) {
        uint32_t hs[4] = {0u, 0u, 0u, 0u}, ls[4] = {0u, 0u, 0u, 0u}, he[4] = {0u, 0u, 0u, 0u}, le[4] = {0u, 0u, 0u, 0u};
        uint32_t hx[4] = {0u, 0u, 0u, 0u}, lx[4] = {0u, 0u, 0u, 0u};
        uint32_t hs2[4] = {0u, 0u, 0u, 0u}, ls2[4] = {0u, 0u, 0u, 0u}, he2[4] = {0u, 0u, 0u, 0u}, le2[4] = {0u, 0u, 0u, 0u};
        uint32_t hx2[4] = {0u, 0u, 0u, 0u}, lx2[4] = {0u, 0u, 0u, 0u};
        if (j0 < cnt) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj;
            const int ee = min(j, cnt - 1);
            const bool live = j < cnt;
            Gathered gc;
            if (staged) gather_window(sa.gat[ee], gc);
            else gather(sa.gat[ee], gc);
            if (!SECOND) {
              split_fast(live ? st.pw_first(sa.Y[ee], gc.g) : 0.f, hs[jj], ls[jj]);
              split_fast(live ? ed.pw_first(sa.Y[ee], gc.g) : 0.f, he[jj], le[jj]);
              if (L == 0) split_fast(live ? sc.pw_first(sa.Y[ee], &gc.gx) : 0.f, hx[jj], lx[jj]);
            } else {
              const float dd = live ? sa.ddot[ee] : 0.f;
              float al, be;
              st.ab_second(sa.Y[ee], sa.Ydot[ee], gc.g, al, be);
              split_fast(live ? al : 0.f, hs[jj], ls[jj]);
              split_fast(be * dd, hs2[jj], ls2[jj]);
              ed.ab_second(sa.Y[ee], sa.Ydot[ee], gc.g, al, be);
              split_fast(live ? al : 0.f, he[jj], le[jj]);
              split_fast(be * dd, he2[jj], le2[jj]);
              if (L == 0) {
                sc.ab_second(sa.Y[ee], sa.Ydot[ee], &gc.gx, al, be);
                split_fast(live ? al : 0.f, hx[jj], lx[jj]);
                split_fast(be * dd, hx2[jj], lx2[jj]);
              }
            }
          }
        }
        tmem_st4(col_s + j0, hs);
        tmem_st4(col_s + TILES * 16 + j0, ls);
        tmem_st4(col_e + j0, he);
        tmem_st4(col_e + TILES * 16 + j0, le);
        if (L == 0) {
          tmem_st4(col_x + j0, hx);
          tmem_st4(col_x + TILES * 16 + j0, lx);
        }
        if (SECOND) {  // K slots 8..15: (beta ddot) of the same edges, paired with the d-derivative rows of B
          tmem_st4(col_s + KE + j0, hs2);
          tmem_st4(col_s + TILES * 16 + KE + j0, ls2);
          tmem_st4(col_e + KE + j0, he2);
          tmem_st4(col_e + TILES * 16 + KE + j0, le2);
          if (L == 0) {
            tmem_st4(col_x + KE + j0, hx2);
            tmem_st4(col_x + TILES * 16 + KE + j0, lx2);
          }
        }
      }
      tmem_wait_st();
      any = true;
    }
    tc_fence_before();
    __syncthreads();
    pending = has;
    d0 = sm.desc[(c + 1) & 7];
  }
  if (pending) {
    mbar_wait(bar, phase);
    phase ^= 1u;
  }
  tc_fence_after();
  // accumulators -> per-CTA partials [gridDim.x, H, 48] (the groups take alternate 4-column pieces)
  {
    constexpr int tiles_of[3] = {TS, TE, TX};
    const int rows_of[3] = {q, M + q, 2 * M + q};
#pragma unroll
    for (int r = 0; r < NROW; ++r) {
      float* dst = A.wpart + ((size_t)blockIdx.x * H + rows_of[r]) * NB;
#pragma unroll
      for (int c4 = grp; c4 < NB / 4; c4 += 2) {
        float v4[4] = {0.f, 0.f, 0.f, 0.f};
        if (any) {
          tmem_ld4(lane_base + tiles_of[r] * NB + c4 * 4, v4);
          tmem_wait_ld();
          pin(v4);
        }
        *reinterpret_cast<float4*>(dst + c4 * 4) = make_float4(v4[0], v4[1], v4[2], v4[3]);
      }
    }
  }
}

template <int C, int M1, int M2, int ORDER>
__device__ __forceinline__ void wgrad_mma_producer(const NeighborArgs& A, WgradMmaSmem<ORDER>& sm, const uint32_t tmem,
                                                   const uint32_t tiles) {
  constexpr int KE = WgradMma<ORDER>::KE, STAGE = WgradMma<ORDER>::STAGE;
  constexpr bool SECOND = ORDER == 2;
  const int lane = threadIdx.x & 31;
  const xeq_graph_t& g = A.geo.g;
  const uint32_t bar = smem_u32(&sm.bar);
  RowCursor<KE> cur_it;
  cur_it.init(g.t_rowptr, g.t_tile_ptr, g.t_n_tiles);
  GeoPipe<KE, true, false, SECOND> gp;
  gp.init();
  auto step = [&](int c) {  // geometry pipeline of iteration c: C(c+1) -> shared memory, B(c+2), A(c+3)
    if (c + 1 >= 0) gp.stage_c(A.geo, sm.a[(c + 1) % 3], lane);
    if (c + 2 >= 0) gp.stage_b(A.geo, lane);
    const ChunkDesc d = cur_it.next();
    if (lane == 0) sm.desc[(c + 3) & 7] = d;
    gp.stage_a(A.geo, d, lane);
    __syncwarp();
  };
  for (int c = -3; c < 0; ++c) step(c);  // fill: geometry of chunk 0 in shared memory, 1 and 2 in flight
  __syncthreads();
  bool issued = false;
  for (int c = 0;; ++c) {
    if (c > 0 && sm.desc[(c - 1) & 7].cnt > 0) {  // MMAs of chunk c-1 (operands complete at the barrier that ended its iteration)
      tc_fence_after();
      if (elect_one()) {
        issue_wgrad<ORDER>(tmem, tiles + (uint32_t)((c - 1) & 1) * STAGE, !issued);
        umma_commit(bar);
      }
      __syncwarp();
      issued = true;
    }
    if (sm.desc[c & 7].cnt < 0) break;
    step(c);
    __syncthreads();
  }
}

template <int C, int M1, int M2, int ORDER>
__global__ void __launch_bounds__(2 * SL_M + 32, 1) wgrad_mma_kernel(const NeighborArgs A) {
  static_assert(C % SL_C == 0 && M1 == C / 2 && M2 == C / 4, "channel slices of edge_mma.cuh");
  __shared__ WgradMmaSmem<ORDER> sm;
  pdl_trigger();
  const uint32_t tmem = tmem_setup(&sm.slot, &sm.bar, 1);
  pdl_wait();  // setup overlapped the previous kernel's tail
  const uint32_t tiles = (smem_u32(xeq_dyn_smem) + 1023u) & ~1023u;
  const uint32_t win_base = tiles + 2u * WgradMma<ORDER>::STAGE;
  const int t = threadIdx.x;
  const int tg = (t >= SL_M + 32) ? t - (SL_M + 32) : t;  // index inside a consumer group; threads SL_M..SL_M+31 = producer
  if (t >= SL_M && t < SL_M + 32) wgrad_mma_producer<C, M1, M2, ORDER>(A, sm, tmem, tiles);
  else if (tg < SL_C) wgrad_mma_role<0, C, M1, M2, ORDER>(A, sm, tmem, tiles, win_base);
  else if (tg < SL_C + SL_M1) wgrad_mma_role<1, C, M1, M2, ORDER>(A, sm, tmem, tiles, win_base);
  else wgrad_mma_role<2, C, M1, M2, ORDER>(A, sm, tmem, tiles, win_base);
  tmem_teardown(tmem);
}

template <typename Kernel>
static int set_smem(Kernel k, size_t bytes) {
  XEQ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return XEQ_OK;
}

// Launchers.  `wide` selects the 256x0e + 128x1o + 64x2e instantiation (two channel slices per tile of edges:
// grid.y = 2, half as many CTAs along x so that all CTAs are resident at once).
template <int C, bool JVP>
static int launch_center_mma_t(const CenterArgs& A, cudaStream_t st) {
  constexpr int M1 = C / 2, M2 = C / 4, SLICES = C / SL_C;
  static_assert(sizeof(CenterMmaSmem<JVP>) <= 24 * 1024, "static shared memory budget");
  const size_t dyn = 1024 + 2 * (size_t)CenterMma<JVP>::STAGE +
                     (size_t)CenterMma<JVP>::WIN * (SL_C * 4 + SL_M1 * 5 + SL_M2 * 7) * (JVP ? 2 : 1) * 4;
  {  // per-device attribute: set on every launch (cheap)
    int rc = set_smem(center_mma_kernel<C, M1, M2, JVP>, dyn);
    if (rc) return rc;
  }
  const int grid = max(1, min(A.geo.g.n_tiles, num_sms() / SLICES));
  XEQ_CUDA(launch_pdl(center_mma_kernel<C, M1, M2, JVP>, dim3(grid, SLICES), dim3(SL_M + 32), dyn, st, A));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

// JVP pass of the double backward (the forward values come from edge_fwd_ul.cu)
int launch_center_jvp_mma(const CenterArgs& A, bool wide, cudaStream_t st) {
  return wide ? launch_center_mma_t<256, true>(A, st) : launch_center_mma_t<128, true>(A, st);
}

template <int C, int ORDER>
static int launch_wgrad_mma_t(const NeighborArgs& A, int grid, cudaStream_t st) {
  constexpr int M1 = C / 2, M2 = C / 4, SLICES = C / SL_C;
  static_assert(sizeof(WgradMmaSmem<ORDER>) <= 24 * 1024, "static shared memory budget");
  const size_t dyn = 1024 + 2 * (size_t)WgradMma<ORDER>::STAGE + 2 * (size_t)WgradMma<ORDER>::WIN * (SL_C * 2 + SL_M1 * 3 + SL_M2 * 5) * 4;
  {  // per-device attribute: set on every launch (cheap)
    int rc = set_smem(wgrad_mma_kernel<C, M1, M2, ORDER>, dyn);
    if (rc) return rc;
  }
  XEQ_CUDA(launch_pdl(wgrad_mma_kernel<C, M1, M2, ORDER>, dim3(grid, SLICES), dim3(2 * SL_M + 32), dyn, st, A));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

// grid = number of per-CTA partial slabs written to A.wpart ([grid, H, 48]; the slices write disjoint rows)
int launch_wgrad_mma(const NeighborArgs& A, int order, bool wide, int grid, cudaStream_t st) {
  if (wide) return order == 1 ? launch_wgrad_mma_t<256, 1>(A, grid, st) : launch_wgrad_mma_t<256, 2>(A, grid, st);
  return order == 1 ? launch_wgrad_mma_t<128, 1>(A, grid, st) : launch_wgrad_mma_t<128, 2>(A, grid, st);
}

}  // namespace xeq

#ifdef XEQ_TRACE
extern "C" int xeq_debug_nbr_trace(long long* out /* [8][2][256] host */) {
  return (int)cudaMemcpyFromSymbol(out, g_trace_nbr, sizeof(long long) * 8 * 2 * 256);
}
#endif
