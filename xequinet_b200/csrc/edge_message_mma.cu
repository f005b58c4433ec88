// Round-1 mapping (one thread per irrep channel, one consumer group per CTA) of the one kernel that has not moved to the
// unified-lane design yet: the JVP pass of K2bb (center_mma_kernel<JVP>).  The forward, first-order and second-order
// neighbor passes and the weight gradients live in edge_fwd_ul.cu / edge_bwd_ul.cu / edge_bwd2_ul.cu / edge_wgrad_ul.cu.
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


}  // namespace xeq

#ifdef XEQ_TRACE
extern "C" int xeq_debug_nbr_trace(long long* out /* [8][2][256] host */) {
  return (int)cudaMemcpyFromSymbol(out, g_trace_nbr, sizeof(long long) * 8 * 2 * 256);
}
#endif
