// K2bb reverse pass (second derivatives by neighbor; loss.backward() through the forces, utils/trainer.py:302) on the
// "unified lane" blocks of edge_ul.cuh -- the second-order twin of edge_bwd_ul.cu.
//
// Per edge the gradient of  Psi_e = sum_h alpha_h w_h + ddot sum_h beta_h w'_h   needs THREE filter outputs
// per row (w, w', w'').  Stacked along N of one MMA set they take 3 x 5 x 8 = 120 accumulator columns per 8-slot chunk
// -- N must be a multiple of 16, so 160 -- and tensor memory (240 columns of filter rows + 3 groups) has room for 80 per
// group.  The work is therefore split by which outputs it needs:
//
//   MODE 2  main pass, [w | w'] (N = 16, 8 slots per chunk): d/ds_j, d/dv_j and every term of d/dr_e except the one
//           with w'': one coefficient of u (sum alpha w'), one of rp = rdot_perp / d (sum beta w'), three of dY/dr
//           (c_y) and three of the Hessian . rdot (c_z) per lane and slot; two 16-value warp butterflies per quad;
//   MODE 3  w'' pass (N = 16 = 16 slots per chunk): adds  u ddot sum_h beta_h w''_h  to the per-edge d/dr records -- a
//           first-order-like walk with one scalar per slot, ONE 16-value butterfly per chunk.
//
// Same warp roles, mbarrier protocol, window (raw gx / gV rows by TMA bulk copies) and fixed-order reductions as
// edge_bwd_ul.cu.  Contract: the o_s / o_v / o_pos part of xeq_edge_message_bwdbwd.
#include "edge_ul.cuh"

namespace xeq {

using namespace fm;
using namespace ul;

namespace {

template <int MODE>
struct Cfg {
  static_assert(MODE == 2 || MODE == 3, "MODE 2: main pass (w, w'); MODE 3: w'' pass");
  static constexpr int SLOTS = MODE == 3 ? 16 : 8;   // edge slots per chunk
  static constexpr int NQ = SLOTS / 4;
  static constexpr int S2 = MODE == 2 ? SLOTS : 1;   // second-order angular records exist in the main pass only
};
constexpr int NCOL = 16;                     // MMA N: [w | w'] of 8 slots, or w'' of 16 slots
constexpr int DCOLS = TILES * NCOL;          // accumulator columns of one group (80)
constexpr int BSTAGE = 2 * NCOL * 128;       // bytes of one B stage: hi + lo tile of 16 rows
constexpr int NBST = 2;
constexpr int NGEO = 4;
constexpr int NTHREADS = NCONS + G * 32 + 32;
constexpr int NOSTAGE = INT_MIN;

template <int MODE>
struct alignas(16) Geo {
  static constexpr int SLOTS = Cfg<MODE>::SLOTS, NQ = Cfg<MODE>::NQ, S2 = Cfg<MODE>::S2;
  float4 Yt[SLOTS][3];    // harmonics per slot and piece type
  float4 rad[SLOTS];      // (d, chi, dchi, ddchi); zeros for dead slots
  float4 u[SLOTS];        // (unit vector, ddot = u . rdot)
  uint2 goff[SLOTS];      // staged: byte offsets of the center's gx / gV rows inside the window; else (node, -)
  int eid[SLOTS];         // canonical edge id, -1 = dead slot
  Quad qd[NQ];
  int nq;
  int pad[3];
  float4 Yd[S2][3];       // Ydot = G^T rdot per piece type                       (main pass)
  float4 rp[S2];          // rdot_perp / d
  float G[S2][24];        // dY_m / dr_x at [x * 8 + m]
  float Hm[S2][24];       // D_rdot G
};

template <int MODE>
struct Smem {
  static constexpr int SLOTS = Cfg<MODE>::SLOTS;
  Geo<MODE> geo[G][NGEO];
  float4 red[G][2][SLOTS][4][MODE == 2 ? 2 : 1];  // main pass: per warp (P, c_y[3]), (R1, c_z[3]); w'' pass: one float4 per slot
  float2 pair[G][2][32];
  uint64_t geo_full[G][NGEO];
  uint64_t tile_full[G][NBST];
  uint64_t acc_full[G], acc_free[G];
  uint64_t win_full[2], win_free[2];
  uint32_t slot;
};

template <int C, int M1, int M2>
struct Win {
  static constexpr int D = C + 3 * M1 + 5 * M2;
  static constexpr bool ENABLED = (C == 128);
  static constexpr uint32_t GX_BYTES = WH * C * 4;
  static constexpr uint32_t HALF = WH * (C + D) * 4;
};

// ------------------------------------------------------------------------------------------------------
// consumers
// ------------------------------------------------------------------------------------------------------
template <int C, int M1, int M2, int MODE>
__device__ __forceinline__ void bwd2_consumer(const NeighborArgs& A, Smem<MODE>& sm, const uint32_t tmem, const uint32_t tiles_base,
                                              const uint32_t win_base, const int grp) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M;
  constexpr int SLOTS = Cfg<MODE>::SLOTS;
  using GeoT = Geo<MODE>;
  const int L = threadIdx.x - grp * GRP, wq = L >> 5, lane = L & 31, sl = blockIdx.y;
  const int pt = piece_type(L);
  const int q0 = sl * SL_C + L, qp = piece_irrep<C, M1>(L, sl);
  int voff[3], nc;
  piece_offsets<C, M1, M2>(L, sl, voff, nc);
  const uint32_t lane_base = tmem + ((uint32_t)(32 * wq) << 16);
  const uint32_t dbase = lane_base + D_COL + grp * DCOLS;
  const uint32_t full = smem_u32(&sm.acc_full[grp]), free_ = smem_u32(&sm.acc_free[grp]);
  const uint32_t geo0 = smem_u32(&sm.geo[grp][0]), gfull0 = smem_u32(&sm.geo_full[grp][0]);
  const uint32_t tfull0 = smem_u32(&sm.tile_full[grp][0]);
  const uint32_t my_tiles = tiles_base + (uint32_t)grp * (NBST * BSTAGE);
  const uint32_t red0 = smem_u32(&sm.red[grp][0][0][0][0]);
  const uint32_t pair0 = smem_u32(&sm.pair[grp][0][0]);
  const bool need_r = A.gr != nullptr;
  const size_t n_edges = (size_t)A.geo.g.n_edges;
  const float c0 = sqrtf(2.f / A.geo.rc);

  // ---- radial stage -----------------------------------------------------------------------------------
  // main pass: thread L < 96 owns slot L / 12 and the terms k = 2 (L % 12), + 1: psi -> rows 0..7, dpsi -> rows 8..15
  // w'' pass : thread L < 96 owns slot L / 6 and the terms k = 4 (L % 6) .. + 3 of ddpsi -> rows 0..15
  constexpr int KPT = MODE == 2 ? 2 : 4;  // radial terms per thread
  const int rslot = MODE == 2 ? L / 12 : L / 6, rk = MODE == 2 ? L - 12 * rslot : L - 6 * rslot;
  float fr[KPT];
#pragma unroll
  for (int x = 0; x < KPT; ++x) {
    const int k = KPT * rk + x;
    fr[x] = (L < 96 && k >= 1 && k <= NB_) ? A.geo.freq[k - 1] : 0.f;
  }
  const uint32_t rad_off = (uint32_t)offsetof(GeoT, rad) + 16u * (uint32_t)rslot;
  const uint32_t tile_off = MODE == 2 ? (uint32_t)(rslot * 128 + (((rk >> 1) ^ (rslot & 7)) << 4) + (rk & 1) * 8)
                                      : (uint32_t)(rslot * 128 + ((rk ^ (rslot & 7)) << 4));
  auto radial = [&](int c) {
    if (L < 96) {
      const float4 rd = lds128(geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(GeoT) + rad_off);
      const float d = rd.x, chi = rd.y, dchi = rd.z, ddchi = rd.w;
      const float inv = 1.f / (d + 1e-5f);
      const uint32_t t_hi = my_tiles + (uint32_t)(c & (NBST - 1)) * BSTAGE + tile_off, t_lo = t_hi + NCOL * 128;
      if constexpr (MODE == 2) {
        float psi[2], dpsi[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          float sn, cs;
          sincos_reduced(fr[x] * d, sn, cs);
          const float phi = c0 * sn * inv;
          const float dphi = c0 * (fr[x] * cs * inv - sn * inv * inv);
          psi[x] = chi * phi;
          dpsi[x] = dchi * phi + chi * dphi;
        }
        if (rk == 0) { psi[0] = chi; dpsi[0] = dchi; }
        uint32_t hi[4], lo[4];
        split_fast(psi[0], hi[0], lo[0]);
        split_fast(psi[1], hi[1], lo[1]);
        split_fast(dpsi[0], hi[2], lo[2]);
        split_fast(dpsi[1], hi[3], lo[3]);
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(t_hi), "r"(hi[0]), "r"(hi[1]) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(t_hi + 8 * 128), "r"(hi[2]), "r"(hi[3]) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(t_lo), "r"(lo[0]), "r"(lo[1]) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(t_lo + 8 * 128), "r"(lo[2]), "r"(lo[3]) : "memory");
      } else {
        float val[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float f = fr[x];
          float sn, cs;
          sincos_reduced(f * d, sn, cs);
          const float phi = c0 * sn * inv;
          const float dphi = c0 * (f * cs * inv - sn * inv * inv);
          const float ddphi = c0 * (-f * f * sn * inv - 2.f * f * cs * inv * inv + 2.f * sn * inv * inv * inv);
          val[x] = ddchi * phi + 2.f * dchi * dphi + chi * ddphi;  // f = 0 -> exactly zero
        }
        if (rk == 0) val[0] = ddchi;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) split_fast(val[x], hi[x], lo[x]);
        sts128(t_hi, hi[0], hi[1], hi[2], hi[3]);
        sts128(t_lo, lo[0], lo[1], lo[2], lo[3]);
      }
      proxy_fence();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(tfull0 + 8u * (uint32_t)(c & (NBST - 1)));
  };
  auto geo_wait = [&](int c) { mbar_wait(gfull0 + 8u * (uint32_t)(c % NGEO), (uint32_t)((c / NGEO) & 1)); };
  auto geo_nq = [&](int c) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(GeoT) + (uint32_t)offsetof(GeoT, nq)) : "memory");
    return v;
  };

  // ---- per-edge d/dr of a finished chunk: one lane of warp 3 per slot, fixed-order sum over the four warps ----
  auto flush = [&](int c) {
    if (wq == 3 && lane < SLOTS && need_r) {
      const uint32_t ge = geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(GeoT);
      int eid;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(eid) : "r"(ge + (uint32_t)offsetof(GeoT, eid) + 4u * (uint32_t)lane) : "memory");
      if (eid >= 0) {
        const float4 uu = lds128(ge + (uint32_t)offsetof(GeoT, u) + 16u * (uint32_t)lane);
        const float uv[3] = {uu.x, uu.y, uu.z};
        float* dst = A.gr + ((size_t)sl * n_edges + (size_t)eid) * 3;
        if constexpr (MODE == 2) {
          const uint32_t rb = red0 + (uint32_t)(((c & 1) * SLOTS + lane) * 4) * 32u;  // [warp][2] float4
          float4 a[4], z[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) { a[w] = lds128(rb + 32u * w); z[w] = lds128(rb + 32u * w + 16u); }
          const float P = ((a[0].x + a[1].x) + a[2].x) + a[3].x;
          const float R1 = ((z[0].x + z[1].x) + z[2].x) + z[3].x;
          const float cY[8] = {a[0].y + a[1].y, a[0].z + a[1].z, a[0].w + a[1].w, a[2].y, a[2].z, a[2].w, a[3].y, a[3].z};
          const float cZ[8] = {z[0].y + z[1].y, z[0].z + z[1].z, z[0].w + z[1].w, z[2].y, z[2].z, z[2].w, z[3].y, z[3].z};
          const float4 rpv = lds128(ge + (uint32_t)offsetof(GeoT, rp) + 16u * (uint32_t)lane);
          const float rp[3] = {rpv.x, rpv.y, rpv.z};
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            const uint32_t ga = ge + (uint32_t)offsetof(GeoT, G) + 96u * (uint32_t)lane + 32u * (uint32_t)x;
            const uint32_t ha = ge + (uint32_t)offsetof(GeoT, Hm) + 96u * (uint32_t)lane + 32u * (uint32_t)x;
            const float4 g0 = lds128(ga), g1 = lds128(ga + 16), h0 = lds128(ha), h1 = lds128(ha + 16);
            float acc = fmaf(P, uv[x], R1 * rp[x]);
            acc = fmaf(g0.x, cY[0], acc); acc = fmaf(g0.y, cY[1], acc); acc = fmaf(g0.z, cY[2], acc); acc = fmaf(g0.w, cY[3], acc);
            acc = fmaf(g1.x, cY[4], acc); acc = fmaf(g1.y, cY[5], acc); acc = fmaf(g1.z, cY[6], acc); acc = fmaf(g1.w, cY[7], acc);
            acc = fmaf(h0.x, cZ[0], acc); acc = fmaf(h0.y, cZ[1], acc); acc = fmaf(h0.z, cZ[2], acc); acc = fmaf(h0.w, cZ[3], acc);
            acc = fmaf(h1.x, cZ[4], acc); acc = fmaf(h1.y, cZ[5], acc); acc = fmaf(h1.z, cZ[6], acc); acc = fmaf(h1.w, cZ[7], acc);
            dst[x] = acc;
          }
        } else {
          const float4 r = lds128(red0 + (uint32_t)(((c & 1) * SLOTS + lane) * 4) * 4u);  // [warp] floats of this slot
          const float tot = (((r.x + r.y) + r.z) + r.w) * uu.w;                          // ddot * sum_h beta_h w''_h
#pragma unroll
          for (int x = 0; x < 3; ++x) dst[x] += tot * uv[x];  // the main pass wrote the record (same stream, earlier launch)
        }
      }
    }
  };

  // owner row (registers)
  float s_st0 = 0.f, s_ed0 = 0.f, s_sc0 = 0.f, s_stp = 0.f, s_edp = 0.f, v0 = 0.f, vp[3] = {0.f, 0.f, 0.f};
  float t_st0 = 0.f, t_ed0 = 0.f, t_sc0 = 0.f, t_stp = 0.f, t_edp = 0.f, vd0 = 0.f, vdp[3] = {0.f, 0.f, 0.f};  // tangents a_s, a_v
  float a_st0 = 0.f, a_ed0 = 0.f, a_sc0 = 0.f, a_stp = 0.f, a_edp = 0.f;
  float av1_0 = 0.f, av2_0 = 0.f, av1_p[3] = {0.f, 0.f, 0.f}, av2_p[3] = {0.f, 0.f, 0.f};
  int pair_par = 0;

  geo_wait(0);
  int nq = geo_nq(0);
  if (nq > 0) radial(0);
  int c = 0;
  for (; nq >= 0; ++c) {
    geo_wait(c + 1);
    const int nq_next = geo_nq(c + 1);
    if (nq_next > 0) radial(c + 1);
    mbar_wait(full, (uint32_t)(c & 1));
    tc_fence_after();
    if (c > 0) flush(c - 1);
    const uint32_t ge = geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(GeoT);
    float cvals[16];  // w'' pass: one value per slot of the chunk
    if constexpr (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) cvals[i] = 0.f;
    }
#pragma unroll
    for (int qd = 0; qd < Cfg<MODE>::NQ; ++qd) {
      if (qd < nq) {
        int node, fl;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(node), "=r"(fl) : "r"(ge + (uint32_t)offsetof(GeoT, qd) + 8u * (uint32_t)qd) : "memory");
        if (fl & F_TILE_FIRST) {
          if (fl & F_STAGED) mbar_wait(smem_u32(&sm.win_full[(fl & F_BUF) ? 1 : 0]), (fl & F_PAR) ? 1u : 0u);
        }
        if (fl & F_ROW_FIRST) {
          const float* sj = A.s + (size_t)node * H;
          const float* vj = A.v + (size_t)node * D;
          s_st0 = sj[q0]; s_ed0 = sj[M + q0]; s_sc0 = sj[2 * M + q0]; s_stp = sj[qp]; s_edp = sj[M + qp];
          v0 = vj[q0];
          vp[0] = vj[voff[0]]; vp[1] = vj[voff[1]]; vp[2] = nc == 3 ? vj[voff[2]] : 0.f;
          if constexpr (MODE == 2) {
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
            a_st0 = a_ed0 = a_sc0 = a_stp = a_edp = av1_0 = av2_0 = 0.f;
#pragma unroll
            for (int m = 0; m < 3; ++m) av1_p[m] = av2_p[m] = 0.f;
          }
        }
        if (!(fl & F_NOROW)) {
          // filter values of the quad: main pass w and w' (columns qd*4 and 8 + qd*4), w'' pass one output (column qd*4)
          float w[TILES][4], dw[TILES][4];
#pragma unroll
          for (int t = 0; t < TILES; ++t) {
            tmem_ld4(dbase + t * NCOL + qd * 4, w[t]);
            if constexpr (MODE == 2) tmem_ld4(dbase + t * NCOL + 8 + qd * 4, dw[t]);
          }
          float gV0[4], gx0[4], gp[4][3];
          if (fl & F_STAGED) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t ox, ov;
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(ox), "=r"(ov) : "r"(ge + (uint32_t)offsetof(GeoT, goff) + 8u * (uint32_t)(qd * 4 + j)) : "memory");
              gx0[j] = lds32(win_base + ox + 4u * (uint32_t)q0);
              gV0[j] = lds32(win_base + ov + 4u * (uint32_t)q0);
              gp[j][0] = lds32(win_base + ov + 4u * (uint32_t)voff[0]);
              gp[j][1] = lds32(win_base + ov + 4u * (uint32_t)voff[1]);
              gp[j][2] = nc == 3 ? lds32(win_base + ov + 4u * (uint32_t)voff[2]) : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t ox, ov;
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(ox), "=r"(ov) : "r"(ge + (uint32_t)offsetof(GeoT, goff) + 8u * (uint32_t)(qd * 4 + j)) : "memory");
              const float* gxi = A.gx + (size_t)ox * C;
              const float* gVi = A.gV + (size_t)ox * D;
              gx0[j] = __ldg(gxi + q0);
              gV0[j] = __ldg(gVi + q0);
              gp[j][0] = __ldg(gVi + voff[0]);
              gp[j][1] = __ldg(gVi + voff[1]);
              gp[j][2] = nc == 3 ? __ldg(gVi + voff[2]) : 0.f;
            }
          }
          tmem_wait_ld();
#pragma unroll
          for (int t = 0; t < TILES; ++t) {
            pin(w[t]);
            if constexpr (MODE == 2) pin(dw[t]);
          }
          if constexpr (MODE == 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 y = lds128(ge + (uint32_t)offsetof(GeoT, Yt) + 48u * (uint32_t)(qd * 4 + j) + 16u * (uint32_t)pt);
              const float Ap = fmaf(gp[j][0], vp[0], fmaf(gp[j][1], vp[1], gp[j][2] * vp[2]));
              const float Bp = fmaf(gp[j][0], y.x, fmaf(gp[j][1], y.y, gp[j][2] * y.z));
              float val = (s_st0 * (gV0[j] * v0)) * w[0][j];       // beta_h w''_h over the lane's five rows
              val = fmaf(s_ed0 * gV0[j], w[1][j], val);
              val = fmaf(s_sc0 * gx0[j], w[2][j], val);
              val = fmaf(s_stp * Ap, w[3][j], val);
              val = fmaf(s_edp * Bp, w[4][j], val);
              cvals[qd * 4 + j] = val;
            }
          } else {
            // two slots at a time: 8 values per slot (P, c_y[3], R1, c_z[3]) -> one 16-value butterfly per pair
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
              float vals[16];
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const int j = 2 * pr + jj;
                const uint32_t so = (uint32_t)(qd * 4 + j);
                const float4 y = lds128(ge + (uint32_t)offsetof(GeoT, Yt) + 48u * so + 16u * (uint32_t)pt);
                const float4 yd = lds128(ge + (uint32_t)offsetof(GeoT, Yd) + 48u * so + 16u * (uint32_t)pt);
                const float ddot = lds32(ge + (uint32_t)offsetof(GeoT, u) + 16u * so + 12u);
                const float w0 = w[0][j], w1 = w[1][j], w2 = w[2][j], w3 = w[3][j], w4 = w[4][j];
                const float e0 = dw[0][j], e1 = dw[1][j], e2 = dw[2][j], e3 = dw[3][j], e4 = dw[4][j];
                const float dwd0 = e0 * ddot, dwd1 = e1 * ddot, dwd2 = e2 * ddot, dwd3 = e3 * ddot, dwd4 = e4 * ddot;
                const float g0 = gV0[j], gx = gx0[j];
                const float A0 = g0 * v0, Ad0 = g0 * vd0;
                const float Ap = fmaf(gp[j][0], vp[0], fmaf(gp[j][1], vp[1], gp[j][2] * vp[2]));
                const float Adp = fmaf(gp[j][0], vdp[0], fmaf(gp[j][1], vdp[1], gp[j][2] * vdp[2]));
                const float Bp = fmaf(gp[j][0], y.x, fmaf(gp[j][1], y.y, gp[j][2] * y.z));
                const float Bdp = fmaf(gp[j][0], yd.x, fmaf(gp[j][1], yd.y, gp[j][2] * yd.z));
                // state rows
                const float al_st0 = fmaf(t_st0, A0, s_st0 * Ad0), be_st0 = s_st0 * A0;
                a_st0 = fmaf(dwd0, A0, fmaf(w0, Ad0, a_st0));
                av1_0 = fmaf(w0, g0, av1_0);
                av2_0 = fmaf(dwd0, g0, av2_0);
                const float al_stp = fmaf(t_stp, Ap, s_stp * Adp), be_stp = s_stp * Ap;
                a_stp = fmaf(dwd3, Ap, fmaf(w3, Adp, a_stp));
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                  av1_p[m] = fmaf(w3, gp[j][m], av1_p[m]);
                  av2_p[m] = fmaf(dwd3, gp[j][m], av2_p[m]);
                }
                // edge rows (l = 0: Y_0 = 1, no angular part)
                const float al_ed0 = t_ed0 * g0, be_ed0 = s_ed0 * g0;
                a_ed0 = fmaf(dwd1, g0, a_ed0);
                const float al_edp = fmaf(t_edp, Bp, s_edp * Bdp), be_edp = s_edp * Bp;
                a_edp = fmaf(dwd4, Bp, fmaf(w4, Bdp, a_edp));
                const float cc = fmaf(t_edp, w4, s_edp * dwd4), sw = s_edp * w4;
                // scalar row
                const float al_sc0 = gx * t_sc0, be_sc0 = gx * s_sc0;
                a_sc0 = fmaf(gx, dwd2, a_sc0);
                float P = al_st0 * e0;
                P = fmaf(al_ed0, e1, P); P = fmaf(al_sc0, e2, P); P = fmaf(al_stp, e3, P); P = fmaf(al_edp, e4, P);
                float R1 = be_st0 * e0;
                R1 = fmaf(be_ed0, e1, R1); R1 = fmaf(be_sc0, e2, R1); R1 = fmaf(be_stp, e3, R1); R1 = fmaf(be_edp, e4, R1);
                vals[8 * jj + 0] = P;
                vals[8 * jj + 1] = cc * gp[j][0];
                vals[8 * jj + 2] = cc * gp[j][1];
                vals[8 * jj + 3] = cc * gp[j][2];
                vals[8 * jj + 4] = R1;
                vals[8 * jj + 5] = sw * gp[j][0];
                vals[8 * jj + 6] = sw * gp[j][1];
                vals[8 * jj + 7] = sw * gp[j][2];
              }
              if (need_r) {
                const float tot = warp_sum16(vals, lane);  // lane l: vals[l >> 1] = (slot (l >> 4), value (l >> 1) & 7)
                if ((lane & 1) == 0) {
                  const int slot = qd * 4 + 2 * pr + (lane >> 4), k = (lane >> 1) & 7;
                  asm volatile("st.shared.f32 [%0], %1;" ::"r"(red0 + (uint32_t)((((c & 1) * SLOTS + slot) * 4 + wq) * 32 + k * 4)), "f"(tot) : "memory");
                }
              }
            }
          }
        }
        if constexpr (MODE == 2) {
          if (fl & F_ROW_LAST) {
            if (wq == 3) {
              asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(pair0 + (uint32_t)((pair_par * 32 + lane) * 8)), "f"(a_stp), "f"(a_edp) : "memory");
            }
            if (wq >= 2) named_bar_sync(4 + grp, 64);
            if (wq == 2) {
              float px, py;
              asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(px), "=f"(py) : "r"(pair0 + (uint32_t)((pair_par * 32 + lane) * 8)) : "memory");
              a_stp += px;
              a_edp += py;
            }
            pair_par ^= 1;
            const size_t nd = (size_t)node;
            if (A.o_s) {
              float* os = A.o_s + nd * H;
              os[q0] = a_st0;
              os[M + q0] = a_ed0;
              os[2 * M + q0] = a_sc0;
              if (wq != 3) {
                os[qp] = a_stp;
                os[M + qp] = a_edp;
              }
            }
            if (A.o_v) {  // d/dv_j = sum_e (sdot w + s w' ddot) g
              float* ov = A.o_v + nd * D;
              ov[q0] = fmaf(t_st0, av1_0, s_st0 * av2_0);
              ov[voff[0]] = fmaf(t_stp, av1_p[0], s_stp * av2_p[0]);
              ov[voff[1]] = fmaf(t_stp, av1_p[1], s_stp * av2_p[1]);
              if (nc == 3) ov[voff[2]] = fmaf(t_stp, av1_p[2], s_stp * av2_p[2]);
            }
          }
        }
        if ((fl & F_TILE_LAST) && (fl & F_STAGED)) {
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sm.win_free[(fl & F_BUF) ? 1 : 0]));
        }
      }
    }
    if constexpr (MODE == 3) {
      if (need_r) {
        const float tot = warp_sum16(cvals, lane);  // lane l: slot l >> 1
        if ((lane & 1) == 0)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(red0 + (uint32_t)((((c & 1) * SLOTS + (lane >> 1)) * 4 + wq) * 4)), "f"(tot) : "memory");
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(free_);
    nq = nq_next;
  }
  named_bar_sync(1 + grp, GRP);
  if (c > 0) flush(c - 1);
}

// ------------------------------------------------------------------------------------------------------
// producer warp of one group
// ------------------------------------------------------------------------------------------------------
struct SlotRegs {
  int i, j, e, wb;
  int qnode, qflags;
  int nq;
};

template <int C, int M1, int M2, int MODE>
__device__ __forceinline__ void bwd2_producer(const NeighborArgs& A, Smem<MODE>& sm, const uint32_t tmem, const uint32_t tiles_base,
                                              const int grp) {
  using W_ = Win<C, M1, M2>;
  constexpr int D = C + 3 * M1 + 5 * M2;
  constexpr int SLOTS = Cfg<MODE>::SLOTS, NQ = Cfg<MODE>::NQ;
  const int lane = threadIdx.x & 31;
  const xeq_graph_t& g = A.geo.g;
  const uint32_t full = smem_u32(&sm.acc_full[grp]), free_ = smem_u32(&sm.acc_free[grp]);
  const uint32_t gfull0 = smem_u32(&sm.geo_full[grp][0]), tfull0 = smem_u32(&sm.tile_full[grp][0]);
  const uint32_t my_tiles = tiles_base + (uint32_t)grp * (NBST * BSTAGE);

  Walk wk;
  wk.allow_stage = W_::ENABLED;
  wk.init(g, g.t_tile_ptr, g.t_n_tiles, grp);
  int node = wk.valid ? wk.n0 + wk.rphase : 0;
  int e = 0, e1 = 0;
  bool row_open = false, row_first = false, tile_any = false;

  auto stage_a = [&](SlotRegs& o) {
    o.i = 0; o.j = 0; o.e = -1; o.wb = NOSTAGE; o.qnode = 0; o.qflags = 0;
    int nq = 0, sl_idx = -1;
    while (nq < NQ && wk.valid) {
      const int stbits = wk.staged ? (F_STAGED | (wk.buf ? F_BUF : 0) | (wk.par ? F_PAR : 0)) : 0;
      if (!row_open) {
        if (node >= wk.n1) {
          if (!tile_any && wk.tile_mode == 1) {
            if (lane == nq) { o.qnode = wk.n0; o.qflags = stbits | F_NOROW | F_TILE_FIRST | F_TILE_LAST; }
            ++nq;
          }
          wk.next();
          node = wk.valid ? wk.n0 + wk.rphase : 0;
          tile_any = false;
          continue;
        }
        e = g.t_rowptr[node];
        e1 = g.t_rowptr[node + 1];
        row_open = true;
        row_first = true;
      }
      const bool last = e + 4 >= e1;
      int fl = stbits | (row_first ? F_ROW_FIRST : 0) | (tile_any ? 0 : F_TILE_FIRST);
      if (last) fl |= F_ROW_LAST | ((node + wk.rstride >= wk.n1) ? F_TILE_LAST : 0);
      if (lane == nq) { o.qnode = node; o.qflags = fl; }
      const int idx = lane - 4 * nq;
      if (idx >= 0 && idx < 4) {
        o.j = node;
        o.i = node;
        sl_idx = (e + idx < e1) ? e + idx : -1;
        o.wb = wk.staged ? wk.buf * WH - wk.n0 : NOSTAGE;
      }
      e += 4;
      row_first = false;
      tile_any = true;
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
    float pi[3], pj[3], sh[3], rd[3];  // rd: tangent of the edge vector (a_pos_i - a_pos_j - offsets @ a_cell)
  };
  auto stage_b = [&](const SlotRegs& r, PosRegs& p) {
#pragma unroll
    for (int x = 0; x < 3; ++x) p.pi[x] = p.pj[x] = p.sh[x] = p.rd[x] = 0.f;
    if (r.e >= 0) {
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        p.pi[x] = A.geo.pos[3 * r.i + x];
        p.pj[x] = A.geo.pos[3 * r.j + x];
        if (A.geo.a_pos) p.rd[x] = A.geo.a_pos[3 * r.i + x] - A.geo.a_pos[3 * r.j + x];
      }
      if (g.offsets != nullptr) {
        const char4 o = reinterpret_cast<const char4*>(g.offsets)[r.e];
        const int gi = g.node_graph ? g.node_graph[r.j] : 0;
        const float* cl = g.cell + 9 * gi;
        const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
#pragma unroll
        for (int x = 0; x < 3; ++x) p.sh[x] = ox * cl[x] + oy * cl[3 + x] + oz * cl[6 + x];
        if (A.geo.a_cell) {
          const float* ac = A.geo.a_cell + 9 * gi;
#pragma unroll
          for (int x = 0; x < 3; ++x) p.rd[x] -= ox * ac[x] + oy * ac[3 + x] + oz * ac[6 + x];
        }
      }
    }
  };

  auto stage_c = [&](int c, const SlotRegs& r, const PosRegs& p) {
    Geo<MODE>& ge = sm.geo[grp][c % NGEO];
    if (r.nq > 0 && lane < SLOTS) {
      float4 y0 = make_float4(0.f, 0.f, 0.f, 0.f), y1 = y0, y2 = y0, rad = y0, uu = y0;
      if (r.e >= 0) {
        float rv[3], d, u[3], Y[8], Gm[3][8];
#pragma unroll
        for (int x = 0; x < 3; ++x) rv[x] = (p.pi[x] - p.pj[x]) - p.sh[x];
        unit_vector(rv, d, u);
        const Cutoff<float> ct = cutoff_terms(d, A.geo.rc);
        rad = make_float4(d, ct.chi, ct.dchi, ct.ddchi);
        float ddot;
        if constexpr (MODE == 2) {
          angular_first(u, d, Y, Gm);
          float rp[3], Yd[8], Hm[3][8];
          angular_second(u, d, p.rd, Gm, ddot, rp, Yd, Hm);
          ge.Yd[lane][0] = make_float4(Yd[0], Yd[1], Yd[2], 0.f);
          ge.Yd[lane][1] = make_float4(Yd[3], Yd[4], Yd[5], 0.f);
          ge.Yd[lane][2] = make_float4(Yd[6], Yd[7], 0.f, 0.f);
          ge.rp[lane] = make_float4(rp[0], rp[1], rp[2], 0.f);
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            *reinterpret_cast<float4*>(&ge.G[lane][x * 8]) = make_float4(Gm[x][0], Gm[x][1], Gm[x][2], Gm[x][3]);
            *reinterpret_cast<float4*>(&ge.G[lane][x * 8 + 4]) = make_float4(Gm[x][4], Gm[x][5], Gm[x][6], Gm[x][7]);
            *reinterpret_cast<float4*>(&ge.Hm[lane][x * 8]) = make_float4(Hm[x][0], Hm[x][1], Hm[x][2], Hm[x][3]);
            *reinterpret_cast<float4*>(&ge.Hm[lane][x * 8 + 4]) = make_float4(Hm[x][4], Hm[x][5], Hm[x][6], Hm[x][7]);
          }
        } else {
          sph_harm(u, Y);
          ddot = u[0] * p.rd[0] + u[1] * p.rd[1] + u[2] * p.rd[2];
        }
        y0 = make_float4(Y[0], Y[1], Y[2], 0.f);
        y1 = make_float4(Y[3], Y[4], Y[5], 0.f);
        y2 = make_float4(Y[6], Y[7], 0.f, 0.f);
        uu = make_float4(u[0], u[1], u[2], ddot);
      } else if constexpr (MODE == 2) {
        ge.Yd[lane][0] = y0; ge.Yd[lane][1] = y0; ge.Yd[lane][2] = y0;
      }
      ge.Yt[lane][0] = y0;
      ge.Yt[lane][1] = y1;
      ge.Yt[lane][2] = y2;
      ge.rad[lane] = rad;
      ge.u[lane] = uu;
      ge.eid[lane] = r.e;
      const int ii = r.e >= 0 ? r.i : r.j;
      if (r.wb != NOSTAGE) {
        const uint32_t row = (uint32_t)(r.wb + ii), buf = row / (uint32_t)WH, local = row - buf * WH;
        ge.goff[lane] = make_uint2(buf * W_::HALF + local * (uint32_t)(C * 4), buf * W_::HALF + W_::GX_BYTES + local * (uint32_t)(D * 4));
      } else {
        ge.goff[lane] = make_uint2((uint32_t)ii, 0u);
      }
      if (lane < NQ) ge.qd[lane] = Quad{r.qnode, r.qflags};
    }
    if (lane == 0) ge.nq = r.nq;
    __syncwarp();
    if (lane == 0) mbar_arrive(gfull0 + 8u * (uint32_t)(c % NGEO));
  };

  auto issue = [&](int c) {
    const uint32_t idesc = idesc_tf32(NCOL);
    const uint32_t b_hi = my_tiles + (uint32_t)(c & (NBST - 1)) * BSTAGE, b_lo = b_hi + NCOL * 128;
    const uint32_t d0 = tmem + D_COL + (uint32_t)grp * DCOLS;
#pragma unroll
    for (int ks = 0; ks < NBP / 8; ++ks) {
      const uint64_t db_hi = smem_desc(b_hi + ks * 32), db_lo = smem_desc(b_lo + ks * 32);
#pragma unroll
      for (int tile = 0; tile < TILES; ++tile) {
        const uint32_t d = d0 + tile * NCOL;
        mma_ts(d, tmem + A_LO + tile * NBP + ks * 8, db_hi, idesc, ks ? 1u : 0u);
        mma_ts(d, tmem + A_HI + tile * NBP + ks * 8, db_lo, idesc, 1u);
        mma_ts(d, tmem + A_HI + tile * NBP + ks * 8, db_hi, idesc, 1u);
      }
    }
  };

  SlotRegs s2, s3;
  PosRegs p2;
  int n0, n1;
  stage_a(s2);
  stage_b(s2, p2);
  stage_a(s3);
  stage_c(0, s2, p2);
  n0 = s2.nq;
  s2 = s3;
  stage_b(s2, p2);
  stage_a(s3);
  stage_c(1, s2, p2);
  n1 = s2.nq;
  s2 = s3;
  stage_b(s2, p2);
  stage_a(s3);
  for (int c = 0; n0 >= 0; ++c) {
    mbar_wait(tfull0 + 8u * (uint32_t)(c & (NBST - 1)), (uint32_t)((c >> 1) & 1));
    if (c > 0) mbar_wait(free_, (uint32_t)((c - 1) & 1));
    tc_fence_after();
    if (elect_one()) {
      issue(c);
      umma_commit(full);
    }
    __syncwarp();
    stage_c(c + 2, s2, p2);
    n0 = n1;
    n1 = s2.nq;
    s2 = s3;
    stage_b(s2, p2);
    stage_a(s3);
  }
}

template <int C, int M1, int M2, int MODE>
__device__ __forceinline__ void bwd2_loader(const NeighborArgs& A, Smem<MODE>& sm, const uint32_t win_base) {
  using W_ = Win<C, M1, M2>;
  constexpr int D = C + 3 * M1 + 5 * M2;
  const xeq_graph_t& g = A.geo.g;
  if (!W_::ENABLED || g.tile_mode != 1) return;
  if ((threadIdx.x & 31) != 0) return;
  Walk wk;
  wk.init(g, g.t_tile_ptr, g.t_n_tiles, 0);
  for (; wk.valid; wk.next()) {
    if (!wk.staged) continue;
    const int t = wk.staged_count - 1;
    const uint32_t fullb = smem_u32(&sm.win_full[wk.buf]), freeb = smem_u32(&sm.win_free[wk.buf]);
    if (t >= 2) mbar_wait_sleep(freeb, (uint32_t)(((t >> 1) - 1) & 1));
    const uint32_t rows = (uint32_t)(wk.n1 - wk.n0);
    mbar_expect_tx(fullb, rows * (uint32_t)((C + D) * 4));
    const uint32_t dst = win_base + (uint32_t)wk.buf * W_::HALF;
    tma_bulk_g2s(dst, A.gx + (size_t)wk.n0 * C, rows * (uint32_t)(C * 4), fullb);
    tma_bulk_g2s(dst + W_::GX_BYTES, A.gV + (size_t)wk.n0 * D, rows * (uint32_t)(D * 4), fullb);
  }
}

template <int C, int M1, int M2, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) nbr_bwd2_ul_kernel(const NeighborArgs A) {
  static_assert(C % SL_C == 0 && M1 == C / 2 && M2 == C / 4, "channel slices of edge_mma.cuh");
  __shared__ Smem<MODE> sm;
  pdl_trigger();
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) {
    for (int i = 0; i < G; ++i) {
      mbar_init(smem_u32(&sm.acc_full[i]), 1);
      mbar_init(smem_u32(&sm.acc_free[i]), 4);
      for (int k = 0; k < NGEO; ++k) mbar_init(smem_u32(&sm.geo_full[i][k]), 1);
      for (int k = 0; k < NBST; ++k) mbar_init(smem_u32(&sm.tile_full[i][k]), 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&sm.win_full[i]), 1);
      mbar_init(smem_u32(&sm.win_free[i]), 4 * G);
    }
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
  const uint32_t win_base = tiles_base + G * NBST * BSTAGE;
  pdl_wait();
  if (t < GRP) store_filter_rows<C, M1, M2>(A.W, A.b, t, blockIdx.y, tmem + ((uint32_t)(32 * (t >> 5)) << 16));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4 * G) bwd2_consumer<C, M1, M2, MODE>(A, sm, tmem, tiles_base, win_base, warp >> 2);
  else if (warp < 4 * G + G) bwd2_producer<C, M1, M2, MODE>(A, sm, tmem, tiles_base, warp - 4 * G);
  else bwd2_loader<C, M1, M2, MODE>(A, sm, win_base);
  tmem_teardown(tmem);
}

template <int C, int MODE>
static int launch_one(const NeighborArgs& A, cudaStream_t st) {
  constexpr int M1 = C / 2, M2 = C / 4, SLICES = C / SL_C;
  using W_ = Win<C, M1, M2>;
  static_assert(sizeof(Smem<MODE>) <= 48 * 1024, "static shared memory limit");
  const xeq_graph_t& g = A.geo.g;
  const bool window = W_::ENABLED && g.tile_mode == 1;
  const size_t dyn_max = 1024 + (size_t)G * NBST * BSTAGE + (W_::ENABLED ? (size_t)2 * W_::HALF : 0);
  const size_t dyn = 1024 + (size_t)G * NBST * BSTAGE + (window ? (size_t)2 * W_::HALF : 0);
  XEQ_CUDA(cudaFuncSetAttribute(nbr_bwd2_ul_kernel<C, M1, M2, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_max));
  const int work = g.tile_mode == 1 ? g.t_n_tiles : (g.t_n_tiles + G - 1) / G;
  const int grid = max(1, min(work, num_sms() / SLICES));
  XEQ_CUDA(launch_pdl(nbr_bwd2_ul_kernel<C, M1, M2, MODE>, dim3(grid, SLICES), dim3(NTHREADS), dyn, st, A));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

}  // namespace

// main pass, then (when the per-edge d/dr records are wanted) the w'' pass that adds its term to them
int launch_nbr2_ul(const NeighborArgs& A, bool wide, cudaStream_t st) {
  int rc = wide ? launch_one<256, 2>(A, st) : launch_one<128, 2>(A, st);
  if (rc || A.gr == nullptr || (A.geo.a_pos == nullptr && A.geo.a_cell == nullptr)) return rc;  // ddot = 0: no w'' term
  NeighborArgs B = A;
  B.o_s = nullptr;
  B.o_v = nullptr;
  return wide ? launch_one<256, 3>(B, st) : launch_one<128, 3>(B, st);
}

}  // namespace xeq
