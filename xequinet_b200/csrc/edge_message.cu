// K2 / K2b / K2bb: the fused XPaiNN edge message and its first and second derivatives.
// Replaces nn/xpainn.py:66-74,140-159 + nn/basic.py:114-131 and their autograd replays
// (nn/basic.py:143-159, utils/trainer.py:302).  Contract: include/xeq_b200.h.
//
// Three kernel families, all persistent over node-aligned edge tiles, no global atomics:
//   center_kernel  : walks CSR rows of the RECEIVING node, one thread per irrep channel q (its
//                    state gate, edge gate, 2l+1 components, + the scalar channel for l = 0);
//                    the segment sum lives in registers.  Forward message, or (JVP) its tangent
//                    along (a_s, a_v, a_pos) = the d/d(gx, gV) half of the double backward.
//   nbr_main_kernel: walks transposed-CSR rows of the SENDING node with the same thread <-> q
//                    mapping; d/ds, d/dv in registers, per-edge d/dr by warp shuffle + shared
//                    memory.  ORDER 1 = K2b (forces), ORDER 2 = reverse half of K2bb.
//   nbr_wgrad_kernel: one thread per filter channel h, weight-gradient partials in registers
//                    (needs no W rows at all), reduced in fixed order by wgrad_reduce_kernel.
// Per-edge geometry (r, d, Y, chi * phi_k and derivatives) is recomputed per chunk into shared
// memory by a 3-stage software pipeline (geometry of chunk c+2 / radial terms of chunk c+1 /
// message of chunk c) with ONE __syncthreads per chunk; nothing E-sized except the 12-byte
// d/dr record ever touches HBM.
#include <stdlib.h>

#include "edge_geo.cuh"

namespace xeq {

#ifdef XEQ_WITH_SIMT  // SIMT filter contraction: TEST-ONLY build (libxeq_b200_simt.so, tests/test_gpu_parity.py A/B check)
// ==========================================================================================
// center kernel
// ==========================================================================================
template <bool JVP> struct CenterChunk { static constexpr int value = JVP ? 48 : 32; };  // smem budget (2 CTAs/SM fwd)

template <bool JVP>
struct CenterSmem {
  static constexpr int TC = CenterChunk<JVP>::value;
  GeoA<TC, false, JVP> a[3];
  GeoB<TC, JVP, false, false, false> b[2];
};

template <int L, int C, int M1, int M2, bool JVP>
__device__ __forceinline__ void center_role(const CenterArgs& A, CenterSmem<JVP>& sm) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M, NC = 2 * L + 1;
  constexpr int THREADS = M;
  constexpr int TC = CenterChunk<JVP>::value;
  const int t = threadIdx.x;
  const int q = t;  // thread index == irrep channel (the l-types are contiguous ranges)
  const int vbase = (L == 0) ? t : (L == 1 ? C + (t - C) : C + 3 * M1 + (t - C - M1));
  constexpr int vstride = (L == 0) ? 0 : (L == 1 ? M1 : M2);
  const xeq_graph_t& g = A.geo.g;

  CenterThread<float, L, NK_> th;
  load_wrow(A.W, A.b, q, th.Ws);
  load_wrow(A.W, A.b, M + q, th.We);
  if (L == 0) load_wrow(A.W, A.b, 2 * M + q, th.Wx);
  th.reset();

  // residual rows of a node are fetched when its row starts (fetch_base) and consumed when it ends
  auto emit_copy = [&](int node) {
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      const size_t idx = (size_t)node * D + vbase + m * vstride;
      A.V_out[idx] = A.V_in ? A.V_in[idx] : 0.f;
    }
    if (L == 0) {
      const size_t idx = (size_t)node * C + t;
      A.x_out[idx] = A.x_in ? A.x_in[idx] : 0.f;
    }
  };

  struct Gathered {
    float ss, se, sx, v[NC], sds, sde, sdx, vd[NC];
  };
  auto gather = [&](const GeoA<TC, false, JVP>& sa, int ee, Gathered& o) {
    const int j = sa.gat[ee];
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

  ChunkCursor<TC> cur_it;
  cur_it.init(g.rowptr, g.tile_ptr, g.n_tiles);
  // staged window: [row][column][thread-of-role] floats, role regions side by side
  constexpr int WMAX = CenterWin<C, JVP>::value;
  constexpr int NCOL = (L == 0 ? 4 : (L == 1 ? 5 : 7)) * (JVP ? 2 : 1);  // s-columns + v-components (+ tangents)
  constexpr int ROWF = (C * 4 + M1 * 5 + M2 * 7) * (JVP ? 2 : 1);
  constexpr int NTHR = (L == 0) ? C : (L == 1 ? M1 : M2);
  constexpr int ROLE_OFF = (L == 0 ? 0 : (L == 1 ? C * 4 : C * 4 + M1 * 5)) * (JVP ? 2 : 1);
  const int tt = (L == 0) ? t : (L == 1 ? t - C : t - C - M1);
  const uint32_t win0 = (uint32_t)__cvta_generic_to_shared(xeq_dyn_smem) + 4u * (ROLE_OFF + tt);
  bool staged = false;
  int win_lo = 0;
  auto stage_window = [&](int n0, int n1) {
#pragma unroll 2
    for (int j = n0; j < n1; ++j) {
      const uint32_t a = win0 + 4u * (uint32_t)((j - n0) * ROWF);
      const float* sj = A.s + (size_t)j * H;
      const float* vj = A.v + (size_t)j * D + vbase;
      float vals[NCOL];
      int c = 0;
      vals[c++] = sj[q];
      vals[c++] = sj[M + q];
      if (L == 0) vals[c++] = sj[2 * M + q];
#pragma unroll
      for (int m = 0; m < NC; ++m) vals[c++] = vj[m * vstride];
      if (JVP) {
        const float* aj = A.a_s ? A.a_s + (size_t)j * H : nullptr;
        vals[c++] = aj ? aj[q] : 0.f;
        vals[c++] = aj ? aj[M + q] : 0.f;
        if (L == 0) vals[c++] = aj ? aj[2 * M + q] : 0.f;
#pragma unroll
        for (int m = 0; m < NC; ++m) vals[c++] = A.a_v ? A.a_v[(size_t)j * D + vbase + m * vstride] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < NCOL; ++k) sts_f32(a + 4u * (uint32_t)(k * NTHR), vals[k]);
    }
  };
  float base_x = 0.f, base_V[NC];
#pragma unroll
  for (int m = 0; m < NC; ++m) base_V[m] = 0.f;
  // pipeline prologue
  ChunkDesc d0 = cur_it.next();  // chunk being processed
  ChunkDesc d1 = cur_it.next();  // chunk whose radial terms are produced
  ChunkDesc d2;                  // chunk whose geometry is produced
  if (d0.cnt >= 0) geo_stage_a1<TC, false, false, JVP>(A.geo, d0, sm.a[0], threadIdx.x);
  __syncthreads();
  if (d1.cnt >= 0) geo_stage_a1<TC, false, false, JVP>(A.geo, d1, sm.a[1], threadIdx.x);
  if (d0.cnt >= 0) geo_stage_a2<TC, THREADS>(A.geo, d0.cnt, sm.a[0], sm.b[0]);
  __syncthreads();

  int cur = -1;
  for (int c = 0; d0.cnt >= 0; ++c) {
    d2 = cur_it.next();
    if (d2.cnt >= 0) geo_stage_a1<TC, false, false, JVP>(A.geo, d2, sm.a[(c + 2) % 3], threadIdx.x);
    if (d1.cnt >= 0) geo_stage_a2<TC, THREADS>(A.geo, d1.cnt, sm.a[(c + 1) % 3], sm.b[(c + 1) & 1]);

    // ---- message accumulation over chunk c.  Gathers run two edges ahead of their use; the
    // loop is unrolled modulo 3 so the three gather buffers are renamed statically (a register
    // rotation by copies would wait on the in-flight loads and defeat the prefetch).
    const GeoA<TC, false, JVP>& sa = sm.a[c % 3];
    const GeoB<TC, JVP, false, false, false>& sb = sm.b[c & 1];
    const int cnt = d0.cnt;
    if (d0.first) {
      cur = -1;
      staged = WMAX > 0 && g.tile_mode == 1 && (d0.n1 - d0.n0) <= WMAX;
      win_lo = d0.n0;
      if (staged) stage_window(d0.n0, d0.n1);
    }
    auto fetch_base = [&](int node) {
#pragma unroll
      for (int m = 0; m < NC; ++m) base_V[m] = A.V_in ? A.V_in[(size_t)node * D + vbase + m * vstride] : 0.f;
      if (L == 0) base_x = A.x_in ? A.x_in[(size_t)node * C + t] : 0.f;
    };
    auto emit_acc = [&](int node) {
#pragma unroll
      for (int m = 0; m < NC; ++m) A.V_out[(size_t)node * D + vbase + m * vstride] = base_V[m] + th.accV[m];
      if (L == 0) A.x_out[(size_t)node * C + t] = base_x + th.accx;
    };
    auto body = [&](int ee, const Gathered& gc) {
      const int i = sa.own[ee];
      if (i != cur) {
        if (cur >= 0) emit_acc(cur);
        for (int nn = (cur >= 0 ? cur + 1 : d0.n0); nn < i; ++nn) emit_copy(nn);
        cur = i;
        fetch_base(i);
        th.reset();
      }
      float p[NBP], Yl[NC], Yd[NC];
      lds_row(sb.psi[ee], p);
      if (L > 0) {
#pragma unroll
        for (int m = 0; m < NC; ++m) {
          Yl[m] = sa.Y[ee][YOff<L>::value + m];
          Yd[m] = JVP ? sa.Ydot[ee][YOff<L>::value + m] : 0.f;
        }
      }
      if (!JVP) {
        th.fwd(p, Yl - YOff<L>::value, gc.ss, gc.se, gc.sx, gc.v);
      } else {
        float dp[NBP];
        lds_row(sb.dpsi[ee], dp);
        th.jvp(p, dp, Yl - YOff<L>::value, Yd - YOff<L>::value, sa.ddot[ee], gc.ss, gc.se, gc.sx, gc.v, gc.sds, gc.sde,
               gc.sdx, gc.vd);
      }
    };
    if (staged) {  // neighbor rows come from the shared-memory window
      for (int ee = 0; ee < cnt; ++ee) {
        const uint32_t a = win0 + 4u * (uint32_t)((sa.gat[ee] - win_lo) * ROWF);
        Gathered gc;
        int c = 0;
        gc.ss = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
        gc.se = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
        gc.sx = (L == 0) ? lds_f32(a + 4u * (uint32_t)(NTHR * c++)) : 0.f;
#pragma unroll
        for (int m = 0; m < NC; ++m) gc.v[m] = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
        if (JVP) {
          gc.sds = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
          gc.sde = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
          gc.sdx = (L == 0) ? lds_f32(a + 4u * (uint32_t)(NTHR * c++)) : 0.f;
#pragma unroll
          for (int m = 0; m < NC; ++m) gc.vd[m] = lds_f32(a + 4u * (uint32_t)(NTHR * c++));
        }
        body(ee, gc);
      }
    } else {  // direct gathers from L2/HBM, prefetched two edges ahead (modulo-3 register renaming)
      Gathered g0, g1, g2;
      if (cnt > 0) gather(sa, 0, g0);
      if (cnt > 1) gather(sa, 1, g1);
      for (int ee = 0; ee < cnt; ee += 3) {
        if (ee + 2 < cnt) gather(sa, ee + 2, g2);
        body(ee, g0);
        if (ee + 1 < cnt) {
          if (ee + 3 < cnt) gather(sa, ee + 3, g0);
          body(ee + 1, g1);
        }
        if (ee + 2 < cnt) {
          if (ee + 4 < cnt) gather(sa, ee + 4, g1);
          body(ee + 2, g2);
        }
      }
    }
    if (d0.last) {
      if (cur >= 0) emit_acc(cur);
      for (int nn = (cur >= 0 ? cur + 1 : d0.n0); nn < d0.n1; ++nn) emit_copy(nn);
      cur = -1;
    }
    __syncthreads();
    d0 = d1;
    d1 = d2;
  }
}

template <int C, int M1, int M2, bool JVP>
__global__ void __launch_bounds__(C + M1 + M2, (C == 128 && !JVP) ? 2 : 1) center_kernel(const CenterArgs A) {
  __shared__ CenterSmem<JVP> sm;
  const int t = threadIdx.x;
  if (t < C) center_role<0, C, M1, M2, JVP>(A, sm);
  else if (t < C + M1) center_role<1, C, M1, M2, JVP>(A, sm);
  else center_role<2, C, M1, M2, JVP>(A, sm);
}

// ==========================================================================================
// neighbor main kernel (thread per irrep channel q)
// ==========================================================================================
template <int ORDER> struct NbrChunk { static constexpr int value = (ORDER == 2) ? 24 : NT; };

template <int ORDER, int NW>
struct NbrMainSmem {
  static constexpr int TC = NbrChunk<ORDER>::value;
  GeoA<TC, true, ORDER == 2> a[3];
  GeoB<TC, true, ORDER == 2, false, false> b[2];
  float red[2][TC][NW][3];
  int red_eid[2][TC];
};

template <int L, int C, int M1, int M2, int ORDER>
__device__ __forceinline__ void nbr_main_role(const NeighborArgs& A, NbrMainSmem<ORDER, (C + M1 + M2) / 32>& sm) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M, NC = 2 * L + 1;
  constexpr int THREADS = M, NW = M / 32;
  constexpr bool SECOND = ORDER == 2;
  constexpr int TC = NbrChunk<ORDER>::value;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int q = t;
  const int vbase = (L == 0) ? t : (L == 1 ? C + (t - C) : C + 3 * M1 + (t - C - M1));
  constexpr int vstride = (L == 0) ? 0 : (L == 1 ? M1 : M2);
  const xeq_graph_t& g = A.geo.g;
  const int E = g.n_edges;

  NeighborThread<float, L, ROLE_STATE, false, NK_> st;
  NeighborThread<float, L, ROLE_EDGE, false, NK_> ed;
  NeighborThread<float, 0, ROLE_SCALAR, false, NK_> sc;
  load_wrow(A.W, A.b, q, st.Wt);
  load_wrow(A.W, A.b, M + q, ed.Wt);
  if (L == 0) load_wrow(A.W, A.b, 2 * M + q, sc.Wt);
  st.s = st.sd = ed.s = ed.sd = sc.s = sc.sd = 0.f;
#pragma unroll
  for (int m = 0; m < NC; ++m) st.v[m] = st.vd[m] = 0.f;
  st.reset_node(); ed.reset_node(); sc.reset_node();

  auto emit = [&](int node, bool with_acc) {
    if (A.o_s) {
      float* os = A.o_s + (size_t)node * H;
      os[q] = with_acc ? st.acc_s : 0.f;
      os[M + q] = with_acc ? ed.acc_s : 0.f;
      if (L == 0) os[2 * M + q] = with_acc ? sc.acc_s : 0.f;
    }
    if (A.o_v) {
#pragma unroll
      for (int m = 0; m < NC; ++m) A.o_v[(size_t)node * D + vbase + m * vstride] = with_acc ? st.acc_v[m] : 0.f;
    }
  };
  auto begin_node = [&](int j) {
    st.reset_node(); ed.reset_node(); sc.reset_node();
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
  auto gather = [&](const GeoA<TC, true, SECOND>& sa, int ee, Gathered& o) {
    const int i = sa.gat[ee];
#pragma unroll
    for (int m = 0; m < NC; ++m) o.g[m] = A.gV[(size_t)i * D + vbase + m * vstride];
    o.gx = (L == 0) ? A.gx[(size_t)i * C + q] : 0.f;
  };

  ChunkCursor<TC> cur_it;
  cur_it.init(g.t_rowptr, g.t_tile_ptr, g.t_n_tiles);
  constexpr int WMAX = NbrWin<C>::value;
  constexpr int NCOL = (L == 0) ? 2 : (L == 1 ? 3 : 5);  // gV components (+ gx for l = 0)
  constexpr int ROWF = C * 2 + M1 * 3 + M2 * 5;
  constexpr int NTHR = (L == 0) ? C : (L == 1 ? M1 : M2);
  constexpr int ROLE_OFF = (L == 0) ? 0 : (L == 1 ? C * 2 : C * 2 + M1 * 3);
  const int tt = (L == 0) ? t : (L == 1 ? t - C : t - C - M1);
  const uint32_t win0 = (uint32_t)__cvta_generic_to_shared(xeq_dyn_smem) + 4u * (ROLE_OFF + tt);
  bool staged = false;
  int win_lo = 0;
  auto stage_window = [&](int n0, int n1) {
#pragma unroll 4
    for (int i = n0; i < n1; ++i) {
      const uint32_t a = win0 + 4u * (uint32_t)((i - n0) * ROWF);
      float vals[NCOL];
#pragma unroll
      for (int m = 0; m < NC; ++m) vals[m] = A.gV[(size_t)i * D + vbase + m * vstride];
      if (L == 0) vals[NC] = A.gx[(size_t)i * C + q];
#pragma unroll
      for (int k = 0; k < NCOL; ++k) sts_f32(a + 4u * (uint32_t)(k * NTHR), vals[k]);
    }
  };
  ChunkDesc d0 = cur_it.next(), d1 = cur_it.next(), d2, dprev;
  dprev.cnt = -1;
  if (d0.cnt >= 0) geo_stage_a1<TC, true, true, SECOND>(A.geo, d0, sm.a[0], threadIdx.x);
  __syncthreads();
  if (d1.cnt >= 0) geo_stage_a1<TC, true, true, SECOND>(A.geo, d1, sm.a[1], threadIdx.x);
  if (d0.cnt >= 0) geo_stage_a2<TC, THREADS>(A.geo, d0.cnt, sm.a[0], sm.b[0]);
  __syncthreads();

  int cur = -1;
  for (int c = 0; d0.cnt >= 0 || dprev.cnt >= 0; ++c) {
    // ---- stage D: per-edge d/dr of chunk c-1, fixed-order sum over the warps
    if (dprev.cnt > 0 && t < dprev.cnt && A.gr) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      const int rs = (c + 1) & 1;
#pragma unroll
      for (int w = 0; w < NW; ++w) { a0 += sm.red[rs][t][w][0]; a1 += sm.red[rs][t][w][1]; a2 += sm.red[rs][t][w][2]; }
      float* dst = A.gr + (size_t)sm.red_eid[rs][t] * 3;
      dst[0] = a0; dst[1] = a1; dst[2] = a2;
    }
    if (d0.cnt >= 0) {
      d2 = cur_it.next();
      if (d2.cnt >= 0) geo_stage_a1<TC, true, true, SECOND>(A.geo, d2, sm.a[(c + 2) % 3], threadIdx.x);
      if (d1.cnt >= 0) geo_stage_a2<TC, THREADS>(A.geo, d1.cnt, sm.a[(c + 1) % 3], sm.b[(c + 1) & 1]);

      const GeoA<TC, true, SECOND>& sa = sm.a[c % 3];
      const GeoB<TC, true, SECOND, false, false>& sb = sm.b[c & 1];
      const int cnt = d0.cnt, rs = c & 1;
      if (d0.first) {
        cur = -1;
        staged = WMAX > 0 && g.tile_mode == 1 && (d0.n1 - d0.n0) <= WMAX;
        win_lo = d0.n0;
        if (staged) stage_window(d0.n0, d0.n1);
      }
      if (t < cnt) sm.red_eid[rs][t] = sa.eid[t];
      auto body = [&](int ee, const Gathered& gc) {
        const int j = sa.own[ee];
        if (j != cur) {
          if (cur >= 0) emit(cur, true);
          for (int nn = (cur >= 0 ? cur + 1 : d0.n0); nn < j; ++nn) emit(nn, false);
          cur = j;
          begin_node(j);
        }
        float p[NBP], dp[NBP], ddp[NBP];
        lds_row(sb.psi[ee], p);
        lds_row(sb.dpsi[ee], dp);
        if (SECOND) lds_row(sb.ddpsi[ee], ddp);
        NbrEdge<float> ne;
        ne.psi = p; ne.dpsi = dp; ne.ddpsi = ddp; ne.xi = nullptr; ne.dxi = nullptr;
        ne.Y = sa.Y[ee]; ne.G = sa.G[ee];
        ne.Hm = SECOND ? sa.Hm[ee] : nullptr;
        ne.Ydot = SECOND ? sa.Ydot[ee] : nullptr;
        ne.u = sa.u[ee];
        ne.rp = SECOND ? sa.rp[ee] : nullptr;
        ne.ddot = SECOND ? sa.ddot[ee] : 0.f;
        float pr[3], pe[3];
        if (!SECOND) { st.first(ne, gc.g, pr); ed.first(ne, gc.g, pe); }
        else { st.second(ne, gc.g, pr); ed.second(ne, gc.g, pe); }
#pragma unroll
        for (int x = 0; x < 3; ++x) pr[x] += pe[x];
        if (L == 0) {
          if (!SECOND) sc.first(ne, &gc.gx, pe); else sc.second(ne, &gc.gx, pe);
#pragma unroll
          for (int x = 0; x < 3; ++x) pr[x] += pe[x];
        }
#pragma unroll
        for (int x = 0; x < 3; ++x) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) pr[x] += __shfl_xor_sync(0xffffffffu, pr[x], o);
        }
        if (lane == 0) {
          sm.red[rs][ee][warp][0] = pr[0]; sm.red[rs][ee][warp][1] = pr[1]; sm.red[rs][ee][warp][2] = pr[2];
        }
      };
      if (staged) {
        for (int ee = 0; ee < cnt; ++ee) {
          const uint32_t a = win0 + 4u * (uint32_t)((sa.gat[ee] - win_lo) * ROWF);
          Gathered gc;
#pragma unroll
          for (int m = 0; m < NC; ++m) gc.g[m] = lds_f32(a + 4u * (uint32_t)(NTHR * m));
          gc.gx = (L == 0) ? lds_f32(a + 4u * (uint32_t)(NTHR * NC)) : 0.f;
          body(ee, gc);
        }
      } else {
        Gathered g0, g1, g2;
        if (cnt > 0) gather(sa, 0, g0);
        if (cnt > 1) gather(sa, 1, g1);
        for (int ee = 0; ee < cnt; ++ee) {  // (a modulo-3 unrolled variant tripled the code and ran 2x slower)
          if (ee + 2 < cnt) gather(sa, ee + 2, g2);
          body(ee, g0);
          g0 = g1;
          g1 = g2;
        }
      }
      if (d0.last) {
        if (cur >= 0) emit(cur, true);
        for (int nn = (cur >= 0 ? cur + 1 : d0.n0); nn < d0.n1; ++nn) emit(nn, false);
        cur = -1;
      }
    }
    __syncthreads();
    dprev = d0;
    if (d0.cnt >= 0) { d0 = d1; d1 = d2; }
  }
}

template <int C, int M1, int M2, int ORDER>
__global__ void __launch_bounds__(C + M1 + M2, (C == 128 && ORDER == 1) ? 2 : 1) nbr_main_kernel(const NeighborArgs A) {
  __shared__ NbrMainSmem<ORDER, (C + M1 + M2) / 32> sm;
  const int t = threadIdx.x;
  if (t < C) nbr_main_role<0, C, M1, M2, ORDER>(A, sm);
  else if (t < C + M1) nbr_main_role<1, C, M1, M2, ORDER>(A, sm);
  else nbr_main_role<2, C, M1, M2, ORDER>(A, sm);
}

// ==========================================================================================
// neighbor weight-gradient kernel (thread per filter channel h; needs no W rows)
// ==========================================================================================
template <int ORDER>
struct NbrWgradSmem {
  GeoA<NT, false, ORDER == 2> a[3];
  GeoB<NT, ORDER == 2, false, true, ORDER == 2> b[2];
};

template <int L, int ROLE, int C, int M1, int M2, int ORDER>
__device__ __forceinline__ void nbr_wgrad_role(const NeighborArgs& A, NbrWgradSmem<ORDER>& sm) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M;
  constexpr bool SECOND = ORDER == 2;
  using Thread = NeighborThread<float, L, ROLE, true, NK_, false>;
  constexpr int NC = Thread::NC;
  const int t = threadIdx.x;
  const int h = blockIdx.y * WTHREADS + t;
  const int q = (ROLE == ROLE_STATE) ? h : (ROLE == ROLE_EDGE ? h - M : 0);
  const int vbase = (ROLE == ROLE_SCALAR) ? 0 : ((L == 0) ? q : (L == 1 ? C + (q - C) : C + 3 * M1 + (q - C - M1)));
  constexpr int vstride = (L == 0) ? 0 : (L == 1 ? M1 : M2);
  const xeq_graph_t& g = A.geo.g;

  Thread th;
  th.reset_wgrad();
  th.s = th.sd = 0.f;
#pragma unroll
  for (int m = 0; m < NC; ++m) th.v[m] = th.vd[m] = 0.f;

  auto gather = [&](const GeoA<NT, false, SECOND>& sa, int ee, float* o) {
    const int i = sa.gat[ee];
    if (ROLE == ROLE_SCALAR) {
      o[0] = A.gx[(size_t)i * C + (h - 2 * M)];
    } else {
#pragma unroll
      for (int m = 0; m < NC; ++m) o[m] = A.gV[(size_t)i * D + vbase + m * vstride];
    }
  };

  ChunkCursor<NT> cur_it;
  cur_it.init(g.t_rowptr, g.t_tile_ptr, g.t_n_tiles);
  ChunkDesc d0 = cur_it.next(), d1 = cur_it.next(), d2;
  if (d0.cnt >= 0) geo_stage_a1<NT, true, false, SECOND>(A.geo, d0, sm.a[0], threadIdx.x);
  __syncthreads();
  if (d1.cnt >= 0) geo_stage_a1<NT, true, false, SECOND>(A.geo, d1, sm.a[1], threadIdx.x);
  if (d0.cnt >= 0) geo_stage_a2<NT, WTHREADS>(A.geo, d0.cnt, sm.a[0], sm.b[0]);
  __syncthreads();

  int cur = -1;
  for (int c = 0; d0.cnt >= 0; ++c) {
    d2 = cur_it.next();
    if (d2.cnt >= 0) geo_stage_a1<NT, true, false, SECOND>(A.geo, d2, sm.a[(c + 2) % 3], threadIdx.x);
    if (d1.cnt >= 0) geo_stage_a2<NT, WTHREADS>(A.geo, d1.cnt, sm.a[(c + 1) % 3], sm.b[(c + 1) & 1]);
    const GeoA<NT, false, SECOND>& sa = sm.a[c % 3];
    const auto& sb = sm.b[c & 1];
    const int cnt = d0.cnt;
    if (d0.first) cur = -1;
    auto body = [&](int ee, const float* gc, float* gn) {
      if (ee + 2 < cnt) gather(sa, ee + 2, gn);
      const int j = sa.own[ee];
      if (j != cur) {
        cur = j;
        th.s = A.s[(size_t)j * H + h];
        if (SECOND) th.sd = A.a_s ? A.a_s[(size_t)j * H + h] : 0.f;
        if (ROLE == ROLE_STATE) {
#pragma unroll
          for (int m = 0; m < NC; ++m) {
            th.v[m] = A.v[(size_t)j * D + vbase + m * vstride];
            if (SECOND) th.vd[m] = A.a_v ? A.a_v[(size_t)j * D + vbase + m * vstride] : 0.f;
          }
        }
      }
      float p[NBP], dp[NBP], xx[NBP], dxx[NBP];
      lds_row(sb.psi[ee], p);
      lds_row(sb.xi[ee], xx);
      if (SECOND) { lds_row(sb.dpsi[ee], dp); lds_row(sb.dxi[ee], dxx); }
      NbrEdge<float> ne;
      ne.psi = p; ne.dpsi = dp; ne.ddpsi = nullptr; ne.xi = xx; ne.dxi = dxx;
      ne.Y = sa.Y[ee]; ne.G = nullptr; ne.Hm = nullptr;
      ne.Ydot = SECOND ? sa.Ydot[ee] : nullptr;
      ne.u = sa.u[ee]; ne.rp = nullptr;
      ne.ddot = SECOND ? sa.ddot[ee] : 0.f;
      float pr[3];
      if (!SECOND) th.first(ne, gc, pr); else th.second(ne, gc, pr);
    };
    float g0[NC], g1[NC], g2[NC];
#pragma unroll
    for (int m = 0; m < NC; ++m) g0[m] = g1[m] = g2[m] = 0.f;
    if (cnt > 0) gather(sa, 0, g0);
    if (cnt > 1) gather(sa, 1, g1);
    for (int ee = 0; ee < cnt; ++ee) {
      body(ee, g0, g2);
#pragma unroll
      for (int m = 0; m < NC; ++m) { g0[m] = g1[m]; g1[m] = g2[m]; }
    }
    __syncthreads();
    d0 = d1;
    d1 = d2;
  }
  float* dst = A.wpart + ((size_t)blockIdx.x * H + h) * (2 * NBP);
#pragma unroll
  for (int k = 0; k < NBP; ++k) { dst[k] = th.GW[k]; dst[NBP + k] = th.GF[k]; }
}

template <int C, int M1, int M2, int ORDER>
__global__ void __launch_bounds__(WTHREADS) nbr_wgrad_kernel(const NeighborArgs A) {
  constexpr int M = C + M1 + M2;
  __shared__ NbrWgradSmem<ORDER> sm;
  const int h = blockIdx.y * WTHREADS + threadIdx.x;
  if (h < C) nbr_wgrad_role<0, ROLE_STATE, C, M1, M2, ORDER>(A, sm);
  else if (h < C + M1) nbr_wgrad_role<1, ROLE_STATE, C, M1, M2, ORDER>(A, sm);
  else if (h < M) nbr_wgrad_role<2, ROLE_STATE, C, M1, M2, ORDER>(A, sm);
  else if (h < M + C) nbr_wgrad_role<0, ROLE_EDGE, C, M1, M2, ORDER>(A, sm);
  else if (h < M + C + M1) nbr_wgrad_role<1, ROLE_EDGE, C, M1, M2, ORDER>(A, sm);
  else if (h < 2 * M) nbr_wgrad_role<2, ROLE_EDGE, C, M1, M2, ORDER>(A, sm);
  else nbr_wgrad_role<0, ROLE_SCALAR, C, M1, M2, ORDER>(A, sm);
}

#endif  // XEQ_WITH_SIMT

// gpos[n] = sum_{e in row n} gr[e] - sum_{slot in t-row n} gr[t_eid[slot]],  gr[e] = sum over the slice slabs (fixed order)
__global__ void pos_grad_kernel(xeq_graph_t g, const float* __restrict__ gr, int n_slabs, float* __restrict__ gpos) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.n_nodes) return;
  const size_t slab = 3 * (size_t)g.n_edges;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int k = 0; k < n_slabs; ++k) {
    const float* grk = gr + k * slab;
    for (int e = g.rowptr[n]; e < g.rowptr[n + 1]; ++e) {
      a0 += grk[3 * (size_t)e]; a1 += grk[3 * (size_t)e + 1]; a2 += grk[3 * (size_t)e + 2];
    }
    for (int s = g.t_rowptr[n]; s < g.t_rowptr[n + 1]; ++s) {
      const size_t e = g.t_eid[s];
      a0 -= grk[3 * e]; a1 -= grk[3 * e + 1]; a2 -= grk[3 * e + 2];
    }
  }
  gpos[3 * n] = a0; gpos[3 * n + 1] = a1; gpos[3 * n + 2] = a2;
}

// rows[k][n] (k = 3 a + b) = sum_{e in row n} offsets[e][a] * gr[e][b]: the per-node pieces of
// dE/dcell[a][b] = - sum_e offsets[e][a] (dE/dr_e)[b]  (the edge vector is pos_i - pos_j - offsets @ cell,
// nn/basic.py:119-128); the caller segment-sums the nine rows per graph.  One thread per node, fixed order.
__global__ void cell_grad_rows_kernel(xeq_graph_t g, const float* __restrict__ gr, int n_slabs, float* __restrict__ rows) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.n_nodes) return;
  const size_t slab = 3 * (size_t)g.n_edges;
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  for (int e = g.rowptr[n]; e < g.rowptr[n + 1]; ++e) {
    const char4 o = reinterpret_cast<const char4*>(g.offsets)[e];
    if (o.x == 0 && o.y == 0 && o.z == 0) continue;
    float d[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < n_slabs; ++k) {
      const float* p = gr + k * slab + 3 * (size_t)e;
      d[0] += p[0]; d[1] += p[1]; d[2] += p[2];
    }
    const float of[3] = {(float)o.x, (float)o.y, (float)o.z};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) acc[3 * a + b] += of[a] * d[b];
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) rows[(size_t)k * g.n_nodes + n] = acc[k];
}

// weight-gradient partials [nblk, H, 2*NBP] -> gW [H,B], gb [H], ftot [H, NB] (fixed order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ wpart, int nblk, int H, float* __restrict__ gW,
                                    float* __restrict__ gb, float* __restrict__ ftot) {
  pdl_trigger();
  pdl_wait();
  const int h = blockIdx.x, k = threadIdx.x;  // k < 2*NBP
  float acc = 0.f;
  for (int bk = 0; bk < nblk; ++bk) acc += wpart[((size_t)bk * H + h) * (2 * NBP) + k];
  if (k == 0) gb[h] = acc;
  else if (k <= NB_) gW[(size_t)h * NB_ + (k - 1)] = acc;
  else if (k > NBP && k <= NBP + NB_) ftot[(size_t)h * NB_ + (k - NBP - 1)] = acc;
}

// gfreq[k] = sum_h W[h,k] * ftot[h,k]
__global__ void freq_grad_kernel(const float* __restrict__ W, const float* __restrict__ ftot, int H,
                                 float* __restrict__ gfreq) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[256];
  const int k = blockIdx.x;
  float acc = 0.f;
  for (int h = threadIdx.x; h < H; h += blockDim.x) acc += W[(size_t)h * NB_ + k] * ftot[(size_t)h * NB_ + k];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) gfreq[k] = red[0];
}

// tile_ptr[k] = first node whose row starts at or after edge k * tile_edges  (k = 0 .. n_tiles)
constexpr int TILE_ROW_COST = 4;
__global__ void tile_bounds_kernel(const int* __restrict__ rowptr, int n_nodes, int n_tiles, int tile_edges,
                                   int* __restrict__ tile_ptr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_tiles) return;
  if (k == n_tiles) { tile_ptr[k] = n_nodes; return; }
  // work of the rows before node n = rowptr[n] edge slots + TILE_ROW_COST slots per row (row switch: owner loads,
  // write-out, a partly empty quad): rows WITHOUT edges count too -- the ghost atoms of a spatially sharded run sit at
  // the end of the node range with empty rows, and a tile that collects hundreds of them stalls one CTA
  const long long x = (long long)k * tile_edges;
  int lo = 0, hi = n_nodes;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((long long)rowptr[mid] + (long long)TILE_ROW_COST * mid < x) lo = mid + 1; else hi = mid;
  }
  tile_ptr[k] = lo;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int check_dims(const xeq_dims_t* d, int* cfg) {
  XEQ_CHECK_ARG(d, "dims is NULL");
  XEQ_CHECK_ARG(d->num_basis == NB_, "num_basis=%d not built (library instantiates %d)", d->num_basis, NB_);
  XEQ_CHECK_ARG(d->node_dim == d->mul0, "node_dim (%d) must equal the 0e multiplicity (%d)", d->node_dim, d->mul0);
  if (d->mul0 == 128 && d->mul1 == 64 && d->mul2 == 32) *cfg = 0;
  else if (d->mul0 == 256 && d->mul1 == 128 && d->mul2 == 64) *cfg = 1;
  else {
    set_error("irreps %dx0e+%dx1o+%dx2e not instantiated (built: 128/64/32, 256/128/64)", d->mul0, d->mul1, d->mul2);
    return XEQ_ERR_INVALID;
  }
  return XEQ_OK;
}

static int check_graph(const xeq_graph_t* g, bool need_t) {
  XEQ_CHECK_ARG(g && g->rowptr && g->n_nodes >= 0 && g->n_edges >= 0, "graph: bad arguments");
  XEQ_CHECK_ARG(g->n_edges == 0 || g->col, "graph: col is NULL");
  XEQ_CHECK_ARG(g->tile_ptr && g->n_tiles >= 1, "graph: tile_ptr missing (xeq_csr_tile_bounds)");
  XEQ_CHECK_ARG(!need_t || g->t_n_tiles >= 1, "graph: t_n_tiles missing");
  XEQ_CHECK_ARG(!need_t || (g->t_rowptr && g->t_tile_ptr && (g->n_edges == 0 || (g->t_row && g->t_eid))),
                "graph: transposed CSR missing");
  XEQ_CHECK_ARG((g->offsets == nullptr) == (g->cell == nullptr), "graph: offsets and cell must be given together");
  XEQ_CHECK_ARG(!g->cell || g->n_graphs == 1 || g->node_graph, "graph: node_graph needed for multi-graph PBC");
  return XEQ_OK;
}

static inline int dims_H(const xeq_dims_t* d) { return d->node_dim + 2 * (d->mul0 + d->mul1 + d->mul2); }

template <typename Kernel>
static int set_smem(Kernel k, size_t bytes) {
  XEQ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return XEQ_OK;
}

#ifdef XEQ_WITH_SIMT
template <int C, int M1, int M2, bool JVP>
static int launch_center(const CenterArgs& A, cudaStream_t st) {
  const int n_tiles = A.geo.g.n_tiles;
  static_assert(sizeof(CenterSmem<JVP>) <= 48 * 1024, "static shared memory limit");
  const bool window = CenterWin<C, JVP>::value > 0 && A.geo.g.tile_mode == 1;
  const size_t dyn = window ? (size_t)CenterWin<C, JVP>::value * (C * 4 + M1 * 5 + M2 * 7) * (JVP ? 2 : 1) * 4 : 0;
  if (dyn) {
    int rc = set_smem(center_kernel<C, M1, M2, JVP>, dyn);
    if (rc) return rc;
  }
  const int grid = min(n_tiles, num_sms() * ((C == 128 && (!window || !JVP)) ? 2 : 1));
  center_kernel<C, M1, M2, JVP><<<grid, C + M1 + M2, dyn, st>>>(A);
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

#endif

// the tcgen05 kernels of the product path
int launch_center_fwd_ul(const CenterArgs& A, bool wide, void* ws, cudaStream_t st);  // edge_fwd_ul.cu: K2 forward
size_t center_fwd_ul_workspace_bytes(int n_nodes, bool wide);
int launch_nbr_bwd_ul(const NeighborArgs& A, bool wide, cudaStream_t st);             // edge_bwd_ul.cu: K2b first order
int launch_center_jvp_ul(const CenterArgs& A, bool wide, void* ws, cudaStream_t st);  // edge_fwd_ul.cu: K2bb JVP pass (tangent rows + tangent filter)
int launch_nbr2_ul(const NeighborArgs& A, bool wide, cudaStream_t st);                // edge_bwd2_ul.cu: K2bb reverse pass (main + w'' passes)
int launch_wgrad_ul(const NeighborArgs& A, int order, bool wide, int grid, cudaStream_t st);   // edge_wgrad_ul.cu: weight gradients

// The product library has ONE implementation per kernel (tcgen05).  A test-only build (-DXEQ_WITH_SIMT,
// xequinet_b200/build.py --simt -> libxeq_b200_simt.so) also carries the round-1 SIMT filter contraction, selected for the
// whole process with XEQ_EDGE_SIMT=1, as an independent implementation for the A/B parity test.
static bool use_mma() {
#ifdef XEQ_WITH_SIMT
  static const bool simt = [] { const char* e = getenv("XEQ_EDGE_SIMT"); return e && e[0] == '1'; }();
  return !simt;
#else
  return true;
#endif
}

static int run_center(const xeq_graph_t* g, const xeq_dims_t* dims, CenterArgs& A, bool jvp, cudaStream_t st, void* ws = nullptr,
                      size_t ws_bytes = 0) {
  int cfg;
  int rc = check_dims(dims, &cfg);
  if (rc) return rc;
  rc = check_graph(g, false);
  if (rc) return rc;
  if (g->n_nodes == 0) return XEQ_OK;
  A.geo.g = *g;
  A.geo.rc = dims->cutoff;
  if (use_mma()) {
    XEQ_CHECK_ARG(ws && ws_bytes >= center_fwd_ul_workspace_bytes(g->n_nodes, cfg == 1), "edge_message: workspace too small");
    return jvp ? launch_center_jvp_ul(A, cfg == 1, ws, st) : launch_center_fwd_ul(A, cfg == 1, ws, st);
  }
#ifdef XEQ_WITH_SIMT
  if (cfg == 0) return jvp ? launch_center<128, 64, 32, true>(A, st) : launch_center<128, 64, 32, false>(A, st);
  return jvp ? launch_center<256, 128, 64, true>(A, st) : launch_center<256, 128, 64, false>(A, st);
#else
  return XEQ_ERR_INVALID;
#endif
}

static int wgrad_grid_x(const xeq_graph_t* g) { return min(g->t_n_tiles, num_sms() * 2); }

static size_t neighbor_ws_bytes(const xeq_graph_t* g, const xeq_dims_t* d, int want_wgrad) {
  const int H = dims_H(d);
  size_t b = 256 + align_up(sizeof(float) * 2 * 3 * (size_t)(g->n_edges > 0 ? g->n_edges : 1), 256);
  if (want_wgrad) {
    b += align_up(sizeof(float) * (size_t)wgrad_grid_x(g) * H * 2 * NBP, 256);
    b += align_up(sizeof(float) * (size_t)H * NB_, 256);
  }
  return b;
}

#ifdef XEQ_WITH_SIMT
template <int C, int M1, int M2, int ORDER>
static int launch_neighbor(NeighborArgs& A, bool main, bool wgrad, int gx, cudaStream_t st) {
  constexpr int M = C + M1 + M2, H = C + 2 * M;
  static_assert(H % WTHREADS == 0, "channel slices must tile H");
  if (main) {
    const int n_tiles = A.geo.g.t_n_tiles;
    static_assert(sizeof(NbrMainSmem<ORDER, M / 32>) <= 48 * 1024, "static shared memory limit");
    const bool window = NbrWin<C>::value > 0 && A.geo.g.tile_mode == 1;
    const size_t dyn = window ? (size_t)NbrWin<C>::value * (C * 2 + M1 * 3 + M2 * 5) * 4 : 0;
    if (dyn) {
      int rc = set_smem(nbr_main_kernel<C, M1, M2, ORDER>, dyn);
      if (rc) return rc;
    }
    const int grid = min(n_tiles, num_sms() * ((C == 128 && ORDER == 1) ? 2 : 1));
    nbr_main_kernel<C, M1, M2, ORDER><<<grid, M, dyn, st>>>(A);
    XEQ_LAUNCHED(1);
  }
  if (wgrad) {
    static_assert(sizeof(NbrWgradSmem<ORDER>) <= 48 * 1024, "static shared memory limit");
    nbr_wgrad_kernel<C, M1, M2, ORDER><<<dim3(gx, H / WTHREADS), WTHREADS, 0, st>>>(A);
    XEQ_LAUNCHED(1);
  }
  return XEQ_OK;
}

#endif

static int run_neighbor(const xeq_graph_t* g, const xeq_dims_t* dims, NeighborArgs& A, int order, float* o_pos,
                        float* o_W, float* o_b, float* o_f, void* ws, size_t ws_bytes, cudaStream_t st) {
  int cfg;
  int rc = check_dims(dims, &cfg);
  if (rc) return rc;
  rc = check_graph(g, true);
  if (rc) return rc;
  const bool wgrad = o_W || o_b || o_f;
  const bool main = A.o_s || A.o_v || o_pos;
  XEQ_CHECK_ARG(!wgrad || (o_W && o_b && o_f), "weight gradients: gW, gb and gfreq must be given together");
  XEQ_CHECK_ARG(ws && ws_bytes >= neighbor_ws_bytes(g, dims, wgrad), "edge_message backward: workspace too small");
  const int H = dims_H(dims);
  if (g->n_nodes == 0) return XEQ_OK;
  Carver cv(ws);
  float* gr = cv.take<float>(2 * 3 * (size_t)(g->n_edges > 0 ? g->n_edges : 1));  // up to two slice slabs
  const bool mma = use_mma();
  const int slices = mma ? (cfg == 1 ? 2 : 1) : 1;  // channel slices of the tcgen05 kernels: one d/dr slab each
  const int gx = mma ? max(1, min(g->t_n_tiles, num_sms() / slices)) : wgrad_grid_x(g);  // per-CTA partial slabs
  float *wpart = nullptr, *ftot = nullptr;
  if (wgrad) {
    wpart = cv.take<float>((size_t)gx * H * 2 * NBP);
    ftot = cv.take<float>((size_t)H * NB_);
  }
  A.geo.g = *g;
  A.geo.rc = dims->cutoff;
  A.gr = o_pos ? gr : nullptr;
  A.wpart = wpart;
  if (mma) {
    if (main) rc = order == 1 ? launch_nbr_bwd_ul(A, cfg == 1, st) : launch_nbr2_ul(A, cfg == 1, st);
    if (!rc && wgrad) rc = launch_wgrad_ul(A, order, cfg == 1, gx, st);
  }
#ifdef XEQ_WITH_SIMT
  else if (cfg == 0) rc = order == 1 ? launch_neighbor<128, 64, 32, 1>(A, main, wgrad, gx, st) : launch_neighbor<128, 64, 32, 2>(A, main, wgrad, gx, st);
  else rc = order == 1 ? launch_neighbor<256, 128, 64, 1>(A, main, wgrad, gx, st) : launch_neighbor<256, 128, 64, 2>(A, main, wgrad, gx, st);
#endif
  if (rc) return rc;
  if (o_pos) {
    XEQ_CUDA(launch_pdl(pos_grad_kernel, dim3((g->n_nodes + 127) / 128), dim3(128), (size_t)0, st, *g, (const float*)gr, slices, o_pos));
    XEQ_LAUNCHED(1);
  }
  if (wgrad) {
    XEQ_CUDA(launch_pdl(wgrad_reduce_kernel, dim3(H), dim3(2 * NBP), (size_t)0, st, (const float*)wpart, gx, H, o_W, o_b, ftot));
    XEQ_CUDA(launch_pdl(freq_grad_kernel, dim3(NB_), dim3(256), (size_t)0, st, A.W, (const float*)ftot, H, o_f));
    XEQ_LAUNCHED(2);
  }
  return XEQ_OK;
}

}  // namespace xeq

using namespace xeq;

extern "C" {

int xeq_center_tile_edges(void) { return CT; }
int xeq_neighbor_tile_edges(void) { return NT; }

int xeq_csr_tile_count(int32_t n_nodes, int32_t n_edges, int32_t tile_edges) {
  return (int)(((long long)n_edges + (long long)TILE_ROW_COST * n_nodes) / tile_edges) + 1;
}

int xeq_csr_tile_bounds(const int32_t* rowptr, int32_t n_nodes, int32_t n_edges, int32_t tile_edges, int32_t* tile_ptr,
                        xeq_stream_t stream) {
  XEQ_CHECK_ARG(rowptr && tile_ptr && n_nodes >= 0 && n_edges >= 0 && tile_edges > 0, "csr_tile_bounds: bad arguments");
  const int n_tiles = xeq_csr_tile_count(n_nodes, n_edges, tile_edges);
  tile_bounds_kernel<<<(n_tiles + 1 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rowptr, n_nodes, n_tiles, tile_edges,
                                                                                 tile_ptr);
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_edge_message_fwd(const xeq_graph_t* g, const xeq_dims_t* dims, const float* pos, const float* s, const float* v,
                         const float* x_in, const float* V_in, const float* W_rbf, const float* b_rbf, const float* freq,
                         float* x_out, float* V_out, void* workspace, size_t workspace_bytes, xeq_stream_t stream) {
  XEQ_CHECK_ARG(pos && s && v && W_rbf && b_rbf && freq && x_out && V_out, "edge_message_fwd: NULL argument");
  CenterArgs A{};
  A.geo.pos = pos; A.geo.freq = freq; A.geo.a_pos = nullptr;
  A.s = s; A.v = v; A.x_in = x_in; A.V_in = V_in; A.W = W_rbf; A.b = b_rbf;
  A.x_out = x_out; A.V_out = V_out;
  return run_center(g, dims, A, false, (cudaStream_t)stream, workspace, workspace_bytes);
}

size_t xeq_edge_message_fwd_workspace_bytes(const xeq_graph_t* g, const xeq_dims_t* dims) {
  if (!g || !dims) return 0;
  return center_fwd_ul_workspace_bytes(g->n_nodes, dims->node_dim > 128);
}

size_t xeq_edge_message_bwd_workspace_bytes(const xeq_graph_t* g, const xeq_dims_t* dims, int want_wgrad) {
  if (!g || !dims) return 0;
  return neighbor_ws_bytes(g, dims, want_wgrad);
}

int xeq_edge_message_bwd(const xeq_graph_t* g, const xeq_dims_t* dims, const float* pos, const float* s, const float* v,
                         const float* W_rbf, const float* b_rbf, const float* freq, const float* gx, const float* gV,
                         float* gs, float* gv, float* gpos, float* gW, float* gb, float* gfreq, void* workspace,
                         size_t workspace_bytes, xeq_stream_t stream) {
  XEQ_CHECK_ARG(pos && s && v && W_rbf && b_rbf && freq && gx && gV, "edge_message_bwd: NULL argument");
  NeighborArgs A{};
  A.geo.pos = pos; A.geo.freq = freq; A.geo.a_pos = nullptr;
  A.s = s; A.v = v; A.W = W_rbf; A.b = b_rbf; A.gx = gx; A.gV = gV;
  A.o_s = gs; A.o_v = gv;
  return run_neighbor(g, dims, A, 1, gpos, gW, gb, gfreq, workspace, workspace_bytes, (cudaStream_t)stream);
}

int xeq_edge_cell_grad_rows(const xeq_graph_t* g, const xeq_dims_t* dims, const void* bwd_workspace, float* rows,
                            xeq_stream_t stream) {
  int cfg;
  int rc = check_dims(dims, &cfg);
  if (rc) return rc;
  XEQ_CHECK_ARG(g && g->rowptr && g->offsets && bwd_workspace && rows, "edge_cell_grad_rows: periodic graph, workspace and rows needed");
  if (g->n_nodes == 0) return XEQ_OK;
  const int slabs = use_mma() ? (cfg == 1 ? 2 : 1) : 1;  // as written by xeq_edge_message_bwd
  cell_grad_rows_kernel<<<(g->n_nodes + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*g, static_cast<const float*>(bwd_workspace),
                                                                                  slabs, rows);
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

size_t xeq_edge_message_bwdbwd_workspace_bytes(const xeq_graph_t* g, const xeq_dims_t* dims, int want_wgrad) {
  if (!g || !dims) return 0;
  // the JVP pass (packed rows) runs first; the reverse pass reuses the region
  const size_t a = neighbor_ws_bytes(g, dims, want_wgrad), b = center_fwd_ul_workspace_bytes(g->n_nodes, dims->node_dim > 128);
  return a > b ? a : b;
}

int xeq_edge_message_bwdbwd(const xeq_graph_t* g, const xeq_dims_t* dims, const float* pos, const float* s, const float* v,
                            const float* W_rbf, const float* b_rbf, const float* freq, const float* gx, const float* gV,
                            const float* a_s, const float* a_v, const float* a_pos, const float* a_cell, float* o_gx, float* o_gV, float* o_s,
                            float* o_v, float* o_pos, float* o_W, float* o_b, float* o_freq, void* workspace,
                            size_t workspace_bytes, xeq_stream_t stream) {
  XEQ_CHECK_ARG(pos && s && v && W_rbf && b_rbf && freq && gx && gV, "edge_message_bwdbwd: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (o_gx || o_gV) {  // d/d(gx, gV): tangent of the forward message along (a_s, a_v, a_pos)
    XEQ_CHECK_ARG(o_gx && o_gV, "edge_message_bwdbwd: o_gx and o_gV must be given together");
    CenterArgs C{};
    C.geo.pos = pos; C.geo.freq = freq; C.geo.a_pos = a_pos; C.geo.a_cell = a_cell;
    C.s = s; C.v = v; C.W = W_rbf; C.b = b_rbf;
    C.a_s = a_s; C.a_v = a_v; C.x_out = o_gx; C.V_out = o_gV;
    int rc = run_center(g, dims, C, true, st, workspace, workspace_bytes);
    if (rc) return rc;
  }
  if (o_s || o_v || o_pos || o_W) {
    NeighborArgs A{};
    A.geo.pos = pos; A.geo.freq = freq; A.geo.a_pos = a_pos; A.geo.a_cell = a_cell;
    A.s = s; A.v = v; A.W = W_rbf; A.b = b_rbf; A.gx = gx; A.gV = gV;
    A.a_s = a_s; A.a_v = a_v; A.o_s = o_s; A.o_v = o_v;
    return run_neighbor(g, dims, A, 2, o_pos, o_W, o_b, o_freq, workspace, workspace_bytes, st);
  }
  return XEQ_OK;
}

}  // extern "C"
