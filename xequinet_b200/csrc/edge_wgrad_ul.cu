// K2b-w / K2bb-w: gradients of the filter weights [b | W_rbf] and of the Bessel frequencies, a GEMM over the edges
//
//   GW[h, k] = sum_e pw[h, e] psi_k(e)                              GF[h, k] = sum_e pw[h, e] xi_k(e)        (first order)
//   GW[h, k] = sum_e alpha[h, e] psi_k(e) + (beta ddot)[h, e] psi'_k(e),   GF likewise with xi, xi'          (second order)
//
// as  D[h, n] += A[h, e] B[n, e]^T  with the edge slots of a chunk as the K dimension of a tcgen05.mma and the 5 x 48
// accumulator columns resident in tensor memory for the whole CTA (n < 24: GW, n >= 24: GF; xi = d psi / d f).  Round-2
// mapping (edge_ul.cuh): the unified lanes produce the rows of A -- pw does not need the filter values, so no filter
// rows live in TMEM here -- three consumer groups run on different rows of the graph, each with its own 80-column A
// staging area (8 K slots, hi + lo): first order 8 edge slots per chunk, second order 4 (alpha | beta ddot stacked along
// K against [psi, xi | psi', xi']).  All groups accumulate into the SAME D, so ONE warp issues every MMA, serving the
// groups round-robin in a fixed order: the summation order -- and the result, bit for bit -- does not depend on timing.
//
//   warps  0-11  consumers, group g = warp / 4: radial tile of the chunk (transposed: row = k, column = slot), gathers,
//                pw (alpha, beta ddot) -> tcgen05.st into the group's A staging, arrive a_ready[g]
//   warps 12-14  geometry producer of group g (row walk, d, chi, harmonics, ddot; two chunks ahead)
//   warp  15     MMA issuer: round r, groups 0..2 in order: wait a_ready[g], 15 tcgen05.mma, commit -> a_free[g]
//
// Per-CTA partials [gridDim.x, H, 48] are reduced in fixed order by wgrad_reduce_kernel / freq_grad_kernel.
#include "edge_ul.cuh"

namespace xeq {

using namespace fm;
using namespace ul;

namespace {

#ifdef XEQ_WG_DEBUG
__device__ __forceinline__ void mbar_wait_dbg(uint32_t bar, uint32_t parity, int id, int c) {
  const long long t0 = clock64();
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > 2000000000ll) {
      unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      if ((threadIdx.x & 31) == 0) printf("t %llu wait timeout id %d c %d block %d,%d warp %d parity %u\n", gt / 1000000ull, id, c, blockIdx.x, blockIdx.y, threadIdx.x >> 5, parity);
      return;
    }
  } while (!ok);
}
#define MBAR_WAIT(bar, par, id, c) mbar_wait_dbg(bar, par, id, c)
#else
#define MBAR_WAIT(bar, par, id, c) mbar_wait(bar, par)
#endif

template <int ORDER>
struct WCfg {
  static constexpr int SLOTS = ORDER == 2 ? 4 : 8;  // edge slots per chunk (MMA K = 8 either way)
  static constexpr int NQ = SLOTS / 4;
};
constexpr int NB = 2 * NBP;                  // accumulator columns per row tile: [GW | GF]
constexpr int D0 = 0;                        // TMEM: accumulators [0, 240)
constexpr int A0 = TILES * NB;               //       A staging of group g: [240 + 80 g, ...): per tile 8 hi + 8 lo columns
constexpr int ACOLS = TILES * 16;
constexpr int BT = NB * 128;                 // bytes of one B tile (48 rows x 128 B, 8 K columns used)
constexpr int BSTAGE = 2 * BT;               // hi + lo
constexpr int NBST = 2;
constexpr int NGEO = 4;
constexpr int NTHREADS = NCONS + G * 32 + 32;

template <int ORDER>
struct alignas(16) Geo {
  static constexpr int SLOTS = WCfg<ORDER>::SLOTS, NQ = WCfg<ORDER>::NQ;
  float4 Yt[SLOTS][3];   // harmonics per slot and piece type
  float4 Yd[SLOTS][3];   // Ydot (second order)
  float4 rad[SLOTS];     // (d, chi, dchi, ddot); zeros for dead slots
  int gat[SLOTS];        // gathered center node (dead slots: the owner, times zero)
  Quad qd[NQ];
  int nq;
  int pad[3];
};

template <int ORDER>
struct Smem {
  Geo<ORDER> geo[G][NGEO];
  float comb[G][32][32];         // final read-out: l = 2 partner lanes hand their tile 3 / 4 columns over ([value][lane])
  int a_nq[G][2];                // quads of the chunk behind a_ready (< 0: end of the group's stream)
  uint64_t geo_full[G][NGEO], geo_free[G][NGEO];
  uint64_t a_ready[G], a_free[G];
  uint64_t d_full;
  uint32_t slot;
};

// ------------------------------------------------------------------------------------------------------
// consumers
// ------------------------------------------------------------------------------------------------------
template <int C, int M1, int M2, int ORDER>
__device__ __forceinline__ void wg_consumer(const NeighborArgs& A, Smem<ORDER>& sm, const uint32_t tmem, const uint32_t tiles_base,
                                            const int grp) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M;
  constexpr int SLOTS = WCfg<ORDER>::SLOTS, NQ = WCfg<ORDER>::NQ;
  constexpr bool SECOND = ORDER == 2;
  using GeoT = Geo<ORDER>;
  const int L = threadIdx.x - grp * GRP, wq = L >> 5, lane = L & 31, sl = blockIdx.y;
  const int pt = piece_type(L);
  const int q0 = sl * SL_C + L, qp = piece_irrep<C, M1>(L, sl);
  int voff[3], nc;
  piece_offsets<C, M1, M2>(L, sl, voff, nc);
  const uint32_t lane_base = tmem + ((uint32_t)(32 * wq) << 16);
  const uint32_t abase = lane_base + A0 + grp * ACOLS;
  const uint32_t geo0 = smem_u32(&sm.geo[grp][0]), gfull0 = smem_u32(&sm.geo_full[grp][0]), gfree0 = smem_u32(&sm.geo_free[grp][0]);
  const uint32_t a_ready = smem_u32(&sm.a_ready[grp]), a_free = smem_u32(&sm.a_free[grp]);
  const uint32_t my_tiles = tiles_base + (uint32_t)grp * (NBST * BSTAGE);
  const float c0 = sqrtf(2.f / A.geo.rc);

  // radial stage, transposed tile: first order thread L < 96 -> slot L / 12, terms k = 2 (L % 12), + 1;
  //                                second order thread L < 96 -> slot L / 24, term k = L % 24 (values and d-derivatives)
  const int rslot = SECOND ? L / 24 : L / 12, rk = SECOND ? L - 24 * rslot : 2 * (L - 12 * rslot);
  constexpr int KPT = SECOND ? 1 : 2;
  float fr[KPT];
#pragma unroll
  for (int x = 0; x < KPT; ++x) fr[x] = (L < 96 && rk + x >= 1 && rk + x <= NB_) ? A.geo.freq[rk + x - 1] : 0.f;
  auto put = [&](uint32_t tile, int n, int kk, float val) {
    uint32_t hi, lo;
    split_fast(val, hi, lo);
    const uint32_t off = b_off(n, kk);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile + off), "r"(hi) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile + BT + off), "r"(lo) : "memory");
  };
  auto radial = [&](int c) {
    if (L < 96) {
      const float4 rd = lds128(geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(GeoT) + (uint32_t)offsetof(GeoT, rad) + 16u * (uint32_t)rslot);
      const float d = rd.x, chi = rd.y, dchi = rd.z;
      const float inv = 1.f / (d + 1e-5f);
      const uint32_t tile = my_tiles + (uint32_t)(c & (NBST - 1)) * BSTAGE;
#pragma unroll
      for (int x = 0; x < KPT; ++x) {
        const int k = rk + x;
        const float f = fr[x];
        float sn, cs;
        sincos_reduced(f * d, sn, cs);
        const float phi = c0 * sn * inv, phif = c0 * d * cs * inv;  // phi_k, d phi_k / d f_k
        float psi = chi * phi, xi = chi * phif;                    // f = 0 (k = 0, padding): psi = 0, xi = chi c0 d inv
        if (k == 0) psi = chi;
        if (k < 1 || k > NB_) xi = 0.f;
        put(tile, k, rslot, psi);
        put(tile, NBP + k, rslot, xi);
        if (SECOND) {
          const float dphi = c0 * (f * cs * inv - sn * inv * inv);
          const float dphif = c0 * (cs * inv - f * d * sn * inv - d * cs * inv * inv);
          float dpsi = dchi * phi + chi * dphi, dxi = dchi * phif + chi * dphif;
          if (k == 0) dpsi = dchi;
          if (k > NB_) dpsi = 0.f;
          if (k < 1 || k > NB_) dxi = 0.f;
          put(tile, k, SLOTS + rslot, dpsi);
          put(tile, NBP + k, SLOTS + rslot, dxi);
        }
      }
      proxy_fence();
    }
  };
  auto geo_wait = [&](int c) { MBAR_WAIT(gfull0 + 8u * (uint32_t)(c % NGEO), (uint32_t)((c / NGEO) & 1), 1, c); };
  auto geo_nq = [&](int c) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(GeoT) + (uint32_t)offsetof(GeoT, nq)) : "memory");
    return v;
  };

  float s_st0 = 0.f, s_ed0 = 0.f, s_sc0 = 0.f, s_stp = 0.f, s_edp = 0.f, v0 = 0.f, vp[3] = {0.f, 0.f, 0.f};
  float t_st0 = 0.f, t_ed0 = 0.f, t_sc0 = 0.f, t_stp = 0.f, t_edp = 0.f, vd0 = 0.f, vdp[3] = {0.f, 0.f, 0.f};

  for (int c = 0;; ++c) {
    geo_wait(c);
    const int nq = geo_nq(c);
    if (nq < 0) {  // end of this group's stream: tell the MMA warp (after it has consumed the previous phase of a_ready)
      if (c > 0) MBAR_WAIT(a_free, (uint32_t)((c - 1) & 1), 9, c);
      __syncwarp();
      if (lane == 0) {
        if (wq == 0) *reinterpret_cast<volatile int*>(&sm.a_nq[grp][c & 1]) = -1;
        mbar_arrive(a_ready);
      }
      break;
    }
    const uint32_t ge = geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(GeoT);
    radial(c);
    // rows of the A operand of this chunk: [tile][8 K slots], slots without an edge contribute zeros
    uint32_t hi[TILES][8], lo[TILES][8];
#pragma unroll
    for (int t = 0; t < TILES; ++t)
#pragma unroll
      for (int k = 0; k < 8; ++k) hi[t][k] = lo[t][k] = 0u;
#pragma unroll
    for (int qd = 0; qd < NQ; ++qd) {
      if (qd < nq) {
        int node, fl;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(node), "=r"(fl) : "r"(ge + (uint32_t)offsetof(GeoT, qd) + 8u * (uint32_t)qd) : "memory");
        if (fl & F_ROW_FIRST) {
          const float* sj = A.s + (size_t)node * H;
          const float* vj = A.v + (size_t)node * D;
          s_st0 = sj[q0]; s_ed0 = sj[M + q0]; s_sc0 = sj[2 * M + q0]; s_stp = sj[qp]; s_edp = sj[M + qp];
          v0 = vj[q0];
          vp[0] = vj[voff[0]]; vp[1] = vj[voff[1]]; vp[2] = nc == 3 ? vj[voff[2]] : 0.f;
          if (SECOND) {
            t_st0 = t_ed0 = t_sc0 = t_stp = t_edp = vd0 = vdp[0] = vdp[1] = vdp[2] = 0.f;
            if (A.a_s) {
              const float* aj = A.a_s + (size_t)node * H;
              t_st0 = aj[q0]; t_ed0 = aj[M + q0]; t_sc0 = aj[2 * M + q0]; t_stp = aj[qp]; t_edp = aj[M + qp];
            }
            if (A.a_v) {
              const float* bj = A.a_v + (size_t)node * D;
              vd0 = bj[q0];
              vdp[0] = bj[voff[0]]; vdp[1] = bj[voff[1]]; vdp[2] = nc == 3 ? bj[voff[2]] : 0.f;
            }
          }
        }
        if (!(fl & F_NOROW)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t so = (uint32_t)(qd * 4 + j);
            int gi;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(gi) : "r"(ge + (uint32_t)offsetof(GeoT, gat) + 4u * so) : "memory");
            const float* gxi = A.gx + (size_t)gi * C;
            const float* gVi = A.gV + (size_t)gi * D;
            const float gx = __ldg(gxi + q0), g0 = __ldg(gVi + q0);
            const float gp0 = __ldg(gVi + voff[0]), gp1 = __ldg(gVi + voff[1]), gp2 = nc == 3 ? __ldg(gVi + voff[2]) : 0.f;
            const float4 y = lds128(ge + (uint32_t)offsetof(GeoT, Yt) + 48u * so + 16u * (uint32_t)pt);
            const float4 rd = lds128(ge + (uint32_t)offsetof(GeoT, rad) + 16u * so);
            const float live = rd.y != 0.f || rd.x != 0.f ? 1.f : 0.f;  // dead slots carry an all-zero record
            const float A0_ = g0 * v0;
            const float Ap = fmaf(gp0, vp[0], fmaf(gp1, vp[1], gp2 * vp[2]));
            const float Bp = fmaf(gp0, y.x, fmaf(gp1, y.y, gp2 * y.z));
            float r0[TILES];  // first order: pw; second order: alpha
            if (!SECOND) {
              r0[0] = s_st0 * A0_; r0[1] = s_ed0 * g0; r0[2] = s_sc0 * gx; r0[3] = s_stp * Ap; r0[4] = s_edp * Bp;
            } else {
              const float4 yd = lds128(ge + (uint32_t)offsetof(GeoT, Yd) + 48u * so + 16u * (uint32_t)pt);
              const float Ad0 = g0 * vd0;
              const float Adp = fmaf(gp0, vdp[0], fmaf(gp1, vdp[1], gp2 * vdp[2]));
              const float Bdp = fmaf(gp0, yd.x, fmaf(gp1, yd.y, gp2 * yd.z));
              r0[0] = fmaf(t_st0, A0_, s_st0 * Ad0); r0[1] = t_ed0 * g0; r0[2] = gx * t_sc0;
              r0[3] = fmaf(t_stp, Ap, s_stp * Adp); r0[4] = fmaf(t_edp, Bp, s_edp * Bdp);
              const float dd = rd.w;  // ddot
              const float r1[TILES] = {s_st0 * A0_ * dd, s_ed0 * g0 * dd, gx * s_sc0 * dd, s_stp * Ap * dd, s_edp * Bp * dd};
#pragma unroll
              for (int t = 0; t < TILES; ++t) split_fast(r1[t] * live, hi[t][SLOTS + qd * 4 + j], lo[t][SLOTS + qd * 4 + j]);
            }
#pragma unroll
            for (int t = 0; t < TILES; ++t) split_fast(r0[t] * live, hi[t][qd * 4 + j], lo[t][qd * 4 + j]);
          }
        }
      }
    }
    // the MMAs of the previous chunk have read the staging area
    if (c > 0) MBAR_WAIT(a_free, (uint32_t)((c - 1) & 1), 2, c);
    tc_fence_after();
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
      tmem_st8(abase + t * 16, hi[t]);
      tmem_st8(abase + t * 16 + 8, lo[t]);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (wq == 0) *reinterpret_cast<volatile int*>(&sm.a_nq[grp][c & 1]) = nq;
      mbar_arrive(a_ready);                                   // A rows + radial tile of chunk c are in place
      mbar_arrive(gfree0 + 8u * (uint32_t)(c % NGEO));         // the geometry record may be overwritten
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// geometry producer of one group
// ------------------------------------------------------------------------------------------------------
struct SlotRegs {
  int i, j, e;
  int qnode, qflags;
  int nq;
};

template <int C, int M1, int M2, int ORDER>
__device__ __forceinline__ void wg_producer(const NeighborArgs& A, Smem<ORDER>& sm, const int grp) {
  constexpr int SLOTS = WCfg<ORDER>::SLOTS, NQ = WCfg<ORDER>::NQ;
  constexpr bool SECOND = ORDER == 2;
  const int lane = threadIdx.x & 31;
  const xeq_graph_t& g = A.geo.g;
  const uint32_t gfull0 = smem_u32(&sm.geo_full[grp][0]), gfree0 = smem_u32(&sm.geo_free[grp][0]);

  Walk wk;
  wk.allow_stage = false;
  wk.init(g, g.t_tile_ptr, g.t_n_tiles, grp);
  int node = wk.valid ? wk.n0 + wk.rphase : 0;
  int e = 0, e1 = 0;
  bool row_open = false, row_first = false;

  auto stage_a = [&](SlotRegs& o) {
    o.i = 0; o.j = 0; o.e = -1; o.qnode = 0; o.qflags = 0;
    int nq = 0, sl_idx = -1;
    while (nq < NQ && wk.valid) {
      if (!row_open) {
        if (node >= wk.n1) {
          wk.next();
          node = wk.valid ? wk.n0 + wk.rphase : 0;
          continue;
        }
        e = g.t_rowptr[node];
        e1 = g.t_rowptr[node + 1];
        if (e >= e1) {  // a row without edges contributes nothing to the weight gradients
          node += wk.rstride;
          continue;
        }
        row_open = true;
        row_first = true;
      }
      const bool last = e + 4 >= e1;
      const int fl = (row_first ? F_ROW_FIRST : 0) | (last ? F_ROW_LAST : 0);
      if (lane == nq) { o.qnode = node; o.qflags = fl; }
      const int idx = lane - 4 * nq;
      if (idx >= 0 && idx < 4) {
        o.j = node;
        o.i = node;
        sl_idx = (e + idx < e1) ? e + idx : -1;
      }
      e += 4;
      row_first = false;
      ++nq;
      if (last) {
        row_open = false;
        node += wk.rstride;
      }
    }
    o.nq = nq ? nq : -1;
    if (sl_idx >= 0) {
      o.i = g.t_row[sl_idx];
      o.e = g.t_eid[sl_idx];
    }
  };

  struct PosRegs {
    float pi[3], pj[3], sh[3], rd[3];
  };
  auto stage_b = [&](const SlotRegs& r, PosRegs& p) {
#pragma unroll
    for (int x = 0; x < 3; ++x) p.pi[x] = p.pj[x] = p.sh[x] = p.rd[x] = 0.f;
    if (r.e >= 0) {
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        p.pi[x] = A.geo.pos[3 * r.i + x];
        p.pj[x] = A.geo.pos[3 * r.j + x];
        if (SECOND && A.geo.a_pos) p.rd[x] = A.geo.a_pos[3 * r.i + x] - A.geo.a_pos[3 * r.j + x];
      }
      if (g.offsets != nullptr) {
        const char4 o = reinterpret_cast<const char4*>(g.offsets)[r.e];
        const int gi = g.node_graph ? g.node_graph[r.j] : 0;
        const float* cl = g.cell + 9 * gi;
        const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
#pragma unroll
        for (int x = 0; x < 3; ++x) p.sh[x] = ox * cl[x] + oy * cl[3 + x] + oz * cl[6 + x];
        if (SECOND && A.geo.a_cell) {
          const float* ac = A.geo.a_cell + 9 * gi;
#pragma unroll
          for (int x = 0; x < 3; ++x) p.rd[x] -= ox * ac[x] + oy * ac[3 + x] + oz * ac[6 + x];
        }
      }
    }
  };

  auto stage_c = [&](int c, const SlotRegs& r, const PosRegs& p) {
    if (c >= NGEO) MBAR_WAIT(gfree0 + 8u * (uint32_t)(c % NGEO), (uint32_t)(((c / NGEO) - 1) & 1), 3, c);  // consumers are done with the slot
    Geo<ORDER>& ge = sm.geo[grp][c % NGEO];
    if (r.nq > 0 && lane < SLOTS) {
      float4 y0 = make_float4(0.f, 0.f, 0.f, 0.f), y1 = y0, y2 = y0, d0 = y0, d1 = y0, d2 = y0, rad = y0;
      if (r.e >= 0) {
        float rv[3], d, u[3], Y[8];
#pragma unroll
        for (int x = 0; x < 3; ++x) rv[x] = (p.pi[x] - p.pj[x]) - p.sh[x];
        unit_vector(rv, d, u);
        float ddot = 0.f;
        if (SECOND) {
          float Gm[3][8], rp[3], Yd[8], Hm[3][8];
          angular_first(u, d, Y, Gm);
          angular_second(u, d, p.rd, Gm, ddot, rp, Yd, Hm);
          d0 = make_float4(Yd[0], Yd[1], Yd[2], 0.f);
          d1 = make_float4(Yd[3], Yd[4], Yd[5], 0.f);
          d2 = make_float4(Yd[6], Yd[7], 0.f, 0.f);
        } else {
          sph_harm(u, Y);
        }
        y0 = make_float4(Y[0], Y[1], Y[2], 0.f);
        y1 = make_float4(Y[3], Y[4], Y[5], 0.f);
        y2 = make_float4(Y[6], Y[7], 0.f, 0.f);
        const Cutoff<float> ct = cutoff_terms(d, A.geo.rc);
        rad = make_float4(d, ct.chi, ct.dchi, ddot);
      }
      ge.Yt[lane][0] = y0; ge.Yt[lane][1] = y1; ge.Yt[lane][2] = y2;
      ge.Yd[lane][0] = d0; ge.Yd[lane][1] = d1; ge.Yd[lane][2] = d2;
      ge.rad[lane] = rad;
      ge.gat[lane] = r.e >= 0 ? r.i : r.j;
      if (lane < NQ) ge.qd[lane] = Quad{r.qnode, r.qflags};
    }
    if (lane == 0) ge.nq = r.nq;
    __syncwarp();
    if (lane == 0) mbar_arrive(gfull0 + 8u * (uint32_t)(c % NGEO));
  };

  // records two chunks ahead of their use are enough to hide the index / position load latencies
  SlotRegs s2, s3;
  PosRegs p2;
  stage_a(s2);
  stage_b(s2, p2);
  stage_a(s3);
  for (int c = 0;; ++c) {
    const int nq = s2.nq;
    stage_c(c, s2, p2);
    if (nq < 0) break;
    s2 = s3;
    stage_b(s2, p2);
    stage_a(s3);
  }
}

// ------------------------------------------------------------------------------------------------------
// MMA issuer: all groups accumulate into the same D, so one thread issues everything, in a fixed order
// ------------------------------------------------------------------------------------------------------
template <int ORDER>
__device__ __forceinline__ void wg_mma_warp(Smem<ORDER>& sm, const uint32_t tmem, const uint32_t tiles_base) {
  const uint32_t idesc = idesc_tf32(NB);
  uint32_t alive = (1u << G) - 1u;
  bool first = true;
  for (int r = 0; alive; ++r) {
    for (int g = 0; g < G; ++g) {
      if (!(alive & (1u << g))) continue;
      MBAR_WAIT(smem_u32(&sm.a_ready[g]), (uint32_t)(r & 1), 4 + g, r);
      const int nq = *reinterpret_cast<volatile int*>(&sm.a_nq[g][r & 1]);
      if (nq < 0) {
        alive &= ~(1u << g);
        continue;
      }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tile = tiles_base + (uint32_t)g * (NBST * BSTAGE) + (uint32_t)(r & (NBST - 1)) * BSTAGE;
        const uint64_t db_hi = smem_desc(tile), db_lo = smem_desc(tile + BT);
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          const uint32_t d = tmem + D0 + t * NB;
          const uint32_t a_hi = tmem + A0 + g * ACOLS + t * 16, a_lo = a_hi + 8;
          mma_ts(d, a_lo, db_hi, idesc, first ? 0u : 1u);
          mma_ts(d, a_hi, db_lo, idesc, 1u);
          mma_ts(d, a_hi, db_hi, idesc, 1u);
        }
        umma_commit(smem_u32(&sm.a_free[g]));
      }
      __syncwarp();
      first = false;
    }
  }
  // every MMA issued above has completed when this commit arrives
  if (elect_one()) {
    umma_commit(smem_u32(&sm.d_full));
  }
  __syncwarp();
  // `first` still set: this CTA had no edge at all; the read-out writes zeros
  if ((threadIdx.x & 31) == 0) *reinterpret_cast<volatile uint32_t*>(&sm.slot) = first ? 0xffffffffu : tmem;
}

template <int C, int M1, int M2, int ORDER>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad_ul_kernel(const NeighborArgs A) {
  static_assert(C % SL_C == 0 && M1 == C / 2 && M2 == C / 4, "channel slices of edge_mma.cuh");
  constexpr int M = C + M1 + M2, H = C + 2 * M;
  __shared__ Smem<ORDER> sm;
  pdl_trigger();
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) {
    for (int i = 0; i < G; ++i) {
      mbar_init(smem_u32(&sm.a_ready[i]), 4);
      mbar_init(smem_u32(&sm.a_free[i]), 1);
      for (int k = 0; k < NGEO; ++k) {
        mbar_init(smem_u32(&sm.geo_full[i][k]), 1);
        mbar_init(smem_u32(&sm.geo_free[i][k]), 4);
      }
    }
    mbar_init(smem_u32(&sm.d_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&sm.slot);
  const uint32_t tiles_base = (smem_u32(xeq_dyn_smem) + 1023u) & ~1023u;
  pdl_wait();
  __syncthreads();  // every thread has read the TMEM base before the MMA warp reuses the slot for the `any` flag
  if (warp < 4 * G) wg_consumer<C, M1, M2, ORDER>(A, sm, tmem, tiles_base, warp >> 2);
  else if (warp < 4 * G + G) wg_producer<C, M1, M2, ORDER>(A, sm, warp - 4 * G);
  else wg_mma_warp<ORDER>(sm, tmem, tiles_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- read-out: accumulators -> per-CTA partials [gridDim.x, H, 48]; group g takes columns [16 g, 16 g + 16) of every
  // tile; the l = 2 partner lanes (96..127: components 3, 4) add their rows of tiles 3, 4 onto lanes 64..95
  if (warp < 4 * G) {
    MBAR_WAIT(smem_u32(&sm.d_full), 0u, 8, 0);
    tc_fence_after();
    const bool any = *reinterpret_cast<volatile uint32_t*>(&sm.slot) != 0xffffffffu;
    const int grp = warp >> 2, L = t - grp * GRP, wq = L >> 5, lane = L & 31, sl = blockIdx.y;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * wq) << 16);
    const int q0 = sl * SL_C + L, qp = piece_irrep<C, M1>(L, sl);
    float vals[TILES][16];
#pragma unroll
    for (int tl = 0; tl < TILES; ++tl) {
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        float v4[4] = {0.f, 0.f, 0.f, 0.f};
        if (any) {
          tmem_ld4(lane_base + D0 + tl * NB + grp * 16 + c4 * 4, v4);
          tmem_wait_ld();
          pin(v4);
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) vals[tl][c4 * 4 + x] = v4[x];
      }
    }
    if (wq == 3) {
#pragma unroll
      for (int k = 0; k < 16; ++k) { sm.comb[grp][k][lane] = vals[3][k]; sm.comb[grp][16 + k][lane] = vals[4][k]; }
    }
    if (wq >= 2) named_bar_sync(4 + grp, 64);
    if (wq == 2) {
#pragma unroll
      for (int k = 0; k < 16; ++k) { vals[3][k] += sm.comb[grp][k][lane]; vals[4][k] += sm.comb[grp][16 + k][lane]; }
    }
    const int rows[TILES] = {q0, M + q0, 2 * M + q0, qp, M + qp};
#pragma unroll
    for (int tl = 0; tl < TILES; ++tl) {
      if (tl >= 3 && wq == 3) continue;  // handed over to the partner lane
      float* dst = A.wpart + ((size_t)blockIdx.x * H + rows[tl]) * NB + grp * 16;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)
        *reinterpret_cast<float4*>(dst + c4 * 4) = make_float4(vals[tl][c4 * 4], vals[tl][c4 * 4 + 1], vals[tl][c4 * 4 + 2], vals[tl][c4 * 4 + 3]);
    }
  }
  tmem_teardown(tmem);
}

template <int C, int ORDER>
static int launch_t(const NeighborArgs& A, int grid, cudaStream_t st) {
  constexpr int M1 = C / 2, M2 = C / 4, SLICES = C / SL_C;
  static_assert(sizeof(Smem<ORDER>) <= 44 * 1024, "static shared memory budget");
  const size_t dyn = 1024 + (size_t)G * NBST * BSTAGE;
  XEQ_CUDA(cudaFuncSetAttribute(wgrad_ul_kernel<C, M1, M2, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  XEQ_CUDA(launch_pdl(wgrad_ul_kernel<C, M1, M2, ORDER>, dim3(grid, SLICES), dim3(NTHREADS), dyn, st, A));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

}  // namespace

// grid = number of per-CTA partial slabs written to A.wpart ([grid, H, 48]; the channel slices write disjoint rows)
int launch_wgrad_ul(const NeighborArgs& A, int order, bool wide, int grid, cudaStream_t st) {
  if (wide) return order == 1 ? launch_t<256, 1>(A, grid, st) : launch_t<256, 2>(A, grid, st);
  return order == 1 ? launch_t<128, 1>(A, grid, st) : launch_t<128, 2>(A, grid, st);
}

}  // namespace xeq
