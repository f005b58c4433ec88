// K2b first derivatives (forces; nn/basic.py:143-159 replays the message in reverse), round-2 design on the
// "unified lane" blocks of edge_ul.cuh -- the transposed twin of edge_fwd_ul.cu:
//
//   a CTA walks TRANSPOSED rows (owner = neighbor j, slots = the edges (i <- j) of all centers i); every lane
//   owns one l = 0 channel (rows state / edge / scalar) and a three-component piece of one l > 0 channel (rows
//   state / edge), keeps the owner's s and v entries in registers and accumulates d/ds_j and d/dv_j over the row;
//   the per-edge d/dr needs a sum over ALL channels: every lane contributes one coefficient of u (radial part) and
//   three coefficients of dY (angular part of its piece), a 16-value halving butterfly per quad reduces them over
//   the warp, and one lane per slot combines the four warps in fixed order and applies dY/dr.
//
//   warps  0-11  consumers, group g = warp / 4 : radial stage of chunk c+1 (psi and dpsi of 8 slots, stacked along
//                N of one MMA set) in the shadow of the MMAs of chunk c; per quad: 10 tcgen05.ld (w, dw of the
//                lane's five rows), 5 gathers per slot of the center's gx / gV entries (shared-memory window filled
//                by TMA bulk copies of the RAW rows -- no packing pass -- or L2), ~35 FMAs per slot
//   warps 12-14  producer of group g : row walk, per-slot geometry (d, chi, dchi, Y, dY/dr) two chunks ahead,
//                MMA issue (45 tcgen05.mma per 8-slot chunk, N = 16 = [w | dw])
//   warp  15     window loader (cp.async.bulk, one tile ahead)
//
// Replaces the autograd replay of nn/xpainn.py:140-159 for d/d(s, v, pos); contract: xeq_edge_message_bwd.
#include "edge_ul.cuh"

namespace xeq {

using namespace fm;
using namespace ul;

namespace {

constexpr int SLOTS = 8;                     // edge slots per chunk; MMA N = 2 * SLOTS ([w | dw])
constexpr int NQ = SLOTS / 4;                // quads per chunk
constexpr int NCOL = 2 * SLOTS;              // accumulator columns per row tile
constexpr int DCOLS = TILES * NCOL;          // accumulator columns of one group (80)
constexpr int BSTAGE = 2 * NCOL * 128;       // bytes of one B stage: hi + lo tile of 16 rows
constexpr int NBST = 2;
constexpr int NGEO = 4;
constexpr int NTHREADS = NCONS + G * 32 + 32;
constexpr int NOSTAGE = INT_MIN;

struct alignas(16) Geo {
  float4 Yt[SLOTS][3];    // harmonics per slot and piece type
  float4 rad[SLOTS];      // (d, chi, dchi, -); zeros for dead slots
  float4 u[SLOTS];        // unit vector
  float G[SLOTS][24];     // dY_m / dr_x at [x * 8 + m]
  uint2 goff[SLOTS];      // staged: byte offsets of the center's gx / gV rows inside the window; else (node, -)
  int eid[SLOTS];         // canonical edge id, -1 = dead slot
  Quad qd[NQ];
  int nq;
  int pad[3];
};

struct BwdSmem {
  Geo geo[G][NGEO];
  float4 red[G][2][SLOTS][4];   // per-warp partial (coefficient of u, three coefficients of dY) of every slot
  float2 pair[G][2][32];        // l = 2 pieces: partial sums of the m = 3, 4 lanes (warp 3 -> warp 2)
  uint64_t geo_full[G][NGEO];
  uint64_t tile_full[G][NBST];
  uint64_t acc_full[G], acc_free[G];
  uint64_t win_full[2], win_free[2];
  uint32_t slot;
};

template <int C, int M1, int M2>
struct Win {
  static constexpr int D = C + 3 * M1 + 5 * M2;
  static constexpr bool ENABLED = (C == 128);                 // wider rows do not fit two window halves
  static constexpr uint32_t GX_BYTES = WH * C * 4;            // gx block of one half
  static constexpr uint32_t HALF = WH * (C + D) * 4;          // [gx rows | gV rows]
};

// ------------------------------------------------------------------------------------------------------
// consumers
// ------------------------------------------------------------------------------------------------------
template <int C, int M1, int M2>
__device__ __forceinline__ void bwd_consumer(const NeighborArgs& A, BwdSmem& sm, const uint32_t tmem, const uint32_t tiles_base,
                                             const uint32_t win_base, const int grp) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M;
  using W_ = Win<C, M1, M2>;
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
  const uint32_t red0 = smem_u32(&sm.red[grp][0][0][0]);
  const uint32_t pair0 = smem_u32(&sm.pair[grp][0][0]);
  const bool need_r = A.gr != nullptr;
  const size_t n_edges = (size_t)A.geo.g.n_edges;
  const xeq_graph_t& g = A.geo.g;

  // radial stage: thread L < 96 owns slot L / 12 and the two radial terms k = 2 (L % 12), + 1 of every chunk
  const int rslot = L / 12, rkp = L - 12 * rslot;
  float fr[2];
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    const int k = 2 * rkp + x;
    fr[x] = (L < 96 && k >= 1 && k <= NB_) ? A.geo.freq[k - 1] : 0.f;
  }
  const float c0 = sqrtf(2.f / A.geo.rc);
  const uint32_t rad_off = (uint32_t)offsetof(Geo, rad) + 16u * (uint32_t)rslot;
  // element (row, k = 2 rkp) of a K-major SWIZZLE_128B tile; psi rows 0..7, dpsi rows 8..15 (same row & 7)
  const uint32_t tile_off = (uint32_t)(rslot * 128 + (((rkp >> 1) ^ (rslot & 7)) << 4) + (rkp & 1) * 8);
  auto radial = [&](int c) {
    if (L < 96) {
      const float4 rd = lds128(geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(Geo) + rad_off);
      const float d = rd.x, chi = rd.y, dchi = rd.z;
      const float inv = 1.f / (d + 1e-5f);
      float psi[2], dpsi[2];
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        float sn, cs;
        sincos_reduced(fr[x] * d, sn, cs);
        const float phi = c0 * sn * inv;
        const float dphi = c0 * (fr[x] * cs * inv - sn * inv * inv);
        psi[x] = chi * phi;                    // fr = 0 (k = 0 or padding) -> exactly zero
        dpsi[x] = dchi * phi + chi * dphi;
      }
      if (rkp == 0) { psi[0] = chi; dpsi[0] = dchi; }  // k = 0: the bias row sees the cutoff envelope alone
      uint32_t hi[4], lo[4];
      split_fast(psi[0], hi[0], lo[0]);
      split_fast(psi[1], hi[1], lo[1]);
      split_fast(dpsi[0], hi[2], lo[2]);
      split_fast(dpsi[1], hi[3], lo[3]);
      const uint32_t t_hi = my_tiles + (uint32_t)(c & (NBST - 1)) * BSTAGE + tile_off, t_lo = t_hi + NCOL * 128;
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(t_hi), "r"(hi[0]), "r"(hi[1]) : "memory");
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(t_hi + SLOTS * 128), "r"(hi[2]), "r"(hi[3]) : "memory");
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(t_lo), "r"(lo[0]), "r"(lo[1]) : "memory");
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(t_lo + SLOTS * 128), "r"(lo[2]), "r"(lo[3]) : "memory");
      proxy_fence();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(tfull0 + 8u * (uint32_t)(c & (NBST - 1)));
  };
  auto geo_wait = [&](int c) { mbar_wait(gfull0 + 8u * (uint32_t)(c % NGEO), (uint32_t)((c / NGEO) & 1)); };
  auto geo_nq = [&](int c) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(Geo) + (uint32_t)offsetof(Geo, nq)) : "memory");
    return v;
  };
  // per-edge d/dr of a finished chunk: lanes 0..7 of warp 3, fixed-order sum over the four warps, then dY/dr
  auto flush = [&](int c) {
    if (wq == 3 && lane < SLOTS && need_r) {
      const uint32_t ge = geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(Geo);
      int eid;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(eid) : "r"(ge + (uint32_t)offsetof(Geo, eid) + 4u * (uint32_t)lane) : "memory");
      if (eid >= 0) {
        const uint32_t rb = red0 + (uint32_t)(((c & 1) * SLOTS + lane) * 4) * 16u;
        const float4 r0 = lds128(rb), r1 = lds128(rb + 16), r2 = lds128(rb + 32), r3 = lds128(rb + 48);
        const float pd = ((r0.x + r1.x) + r2.x) + r3.x;
        const float dY[8] = {r0.y + r1.y, r0.z + r1.z, r0.w + r1.w, r2.y, r2.z, r2.w, r3.y, r3.z};
        const float4 uu = lds128(ge + (uint32_t)offsetof(Geo, u) + 16u * (uint32_t)lane);
        const float uv[3] = {uu.x, uu.y, uu.z};
        float out[3];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const uint32_t ga = ge + (uint32_t)offsetof(Geo, G) + 96u * (uint32_t)lane + 32u * (uint32_t)x;
          const float4 g0 = lds128(ga), g1 = lds128(ga + 16);
          float acc = pd * uv[x];
          acc = fmaf(g0.x, dY[0], acc); acc = fmaf(g0.y, dY[1], acc); acc = fmaf(g0.z, dY[2], acc); acc = fmaf(g0.w, dY[3], acc);
          acc = fmaf(g1.x, dY[4], acc); acc = fmaf(g1.y, dY[5], acc); acc = fmaf(g1.z, dY[6], acc); acc = fmaf(g1.w, dY[7], acc);
          out[x] = acc;
        }
        float* dst = A.gr + ((size_t)sl * n_edges + (size_t)eid) * 3;
        dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2];
      }
    }
  };

  // owner row (registers): s, v entries of this lane's rows; accumulators
  float s_st0 = 0.f, s_ed0 = 0.f, s_sc0 = 0.f, s_stp = 0.f, s_edp = 0.f, v0 = 0.f, vp[3] = {0.f, 0.f, 0.f};
  float a_st0 = 0.f, a_ed0 = 0.f, a_sc0 = 0.f, a_stp = 0.f, a_edp = 0.f, a_v0 = 0.f, a_vp[3] = {0.f, 0.f, 0.f};
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
    if (c > 0) flush(c - 1);  // every warp of the group has finished chunk c-1 (acc_free -> MMA -> acc_full)
    const uint32_t ge = geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(Geo);
#pragma unroll 1
    for (int qd = 0; qd < nq; ++qd) {
      int node, fl;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(node), "=r"(fl) : "r"(ge + (uint32_t)offsetof(Geo, qd) + 8u * (uint32_t)qd) : "memory");
      if (fl & F_TILE_FIRST) {
        if (fl & F_STAGED) mbar_wait(smem_u32(&sm.win_full[(fl & F_BUF) ? 1 : 0]), (fl & F_PAR) ? 1u : 0u);
      }
      if (fl & F_ROW_FIRST) {
        const float* sj = A.s + (size_t)node * H;
        const float* vj = A.v + (size_t)node * D;
        s_st0 = sj[q0]; s_ed0 = sj[M + q0]; s_sc0 = sj[2 * M + q0]; s_stp = sj[qp]; s_edp = sj[M + qp];
        v0 = vj[q0];
        vp[0] = vj[voff[0]]; vp[1] = vj[voff[1]]; vp[2] = nc == 3 ? vj[voff[2]] : 0.f;
        a_st0 = a_ed0 = a_sc0 = a_stp = a_edp = a_v0 = a_vp[0] = a_vp[1] = a_vp[2] = 0.f;
      }
      if (!(fl & F_NOROW)) {
        float w[TILES][4], dw[TILES][4];
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          tmem_ld4(dbase + t * NCOL + qd * 4, w[t]);
          tmem_ld4(dbase + t * NCOL + SLOTS + qd * 4, dw[t]);
        }
        float gV0[4], gx0[4], gp[4][3];
        float4 y[4];
        if (fl & F_STAGED) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t ox, ov;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(ox), "=r"(ov) : "r"(ge + (uint32_t)offsetof(Geo, goff) + 8u * (uint32_t)(qd * 4 + j)) : "memory");
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
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(ox), "=r"(ov) : "r"(ge + (uint32_t)offsetof(Geo, goff) + 8u * (uint32_t)(qd * 4 + j)) : "memory");
            const float* gxi = A.gx + (size_t)ox * C;
            const float* gVi = A.gV + (size_t)ox * D;
            gx0[j] = __ldg(gxi + q0);
            gV0[j] = __ldg(gVi + q0);
            gp[j][0] = __ldg(gVi + voff[0]);
            gp[j][1] = __ldg(gVi + voff[1]);
            gp[j][2] = nc == 3 ? __ldg(gVi + voff[2]) : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = lds128(ge + (uint32_t)offsetof(Geo, Yt) + 48u * (uint32_t)(qd * 4 + j) + 16u * (uint32_t)pt);
        tmem_wait_ld();
#pragma unroll
        for (int t = 0; t < TILES; ++t) { pin(w[t]); pin(dw[t]); }
        float vals[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float A0 = gV0[j] * v0;
          const float Ap = fmaf(gp[j][0], vp[0], fmaf(gp[j][1], vp[1], gp[j][2] * vp[2]));
          const float Bp = fmaf(gp[j][0], y[j].x, fmaf(gp[j][1], y[j].y, gp[j][2] * y[j].z));
          a_st0 = fmaf(A0, w[0][j], a_st0);
          a_ed0 = fmaf(gV0[j], w[1][j], a_ed0);
          a_sc0 = fmaf(gx0[j], w[2][j], a_sc0);
          a_stp = fmaf(Ap, w[3][j], a_stp);
          a_edp = fmaf(Bp, w[4][j], a_edp);
          a_v0 = fmaf(w[0][j], gV0[j], a_v0);
#pragma unroll
          for (int m = 0; m < 3; ++m) a_vp[m] = fmaf(w[3][j], gp[j][m], a_vp[m]);
          float pd = (s_st0 * A0) * dw[0][j];
          pd = fmaf(s_ed0 * gV0[j], dw[1][j], pd);
          pd = fmaf(s_sc0 * gx0[j], dw[2][j], pd);
          pd = fmaf(s_stp * Ap, dw[3][j], pd);
          pd = fmaf(s_edp * Bp, dw[4][j], pd);
          const float se = s_edp * w[4][j];
          vals[4 * j] = pd;
          vals[4 * j + 1] = se * gp[j][0];
          vals[4 * j + 2] = se * gp[j][1];
          vals[4 * j + 3] = se * gp[j][2];
        }
        if (need_r) {
          const float tot = warp_sum16(vals, lane);  // lane l: total of vals[l >> 1] = (slot (l >> 3), component (l >> 1) & 3)
          if ((lane & 1) == 0)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(red0 + (uint32_t)((((c & 1) * SLOTS + qd * 4 + (lane >> 3)) * 4 + wq) * 16 + ((lane >> 1) & 3) * 4)),
                         "f"(tot) : "memory");
        }
      }
      if (fl & F_ROW_LAST) {
        // l = 2 pieces: the m = 3, 4 lanes (warp 3) hand their partial row sums to the m = 0..2 lanes (warp 2)
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
        if (A.o_v) {
          float* ov = A.o_v + nd * D;
          ov[q0] = s_st0 * a_v0;
          ov[voff[0]] = s_stp * a_vp[0];
          ov[voff[1]] = s_stp * a_vp[1];
          if (nc == 3) ov[voff[2]] = s_stp * a_vp[2];
        }
      }
      if ((fl & F_TILE_LAST) && (fl & F_STAGED)) {
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm.win_free[(fl & F_BUF) ? 1 : 0]));
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(free_);
    nq = nq_next;
  }
  // last chunk: its partials are complete once all four warps of the group are here
  named_bar_sync(1 + grp, GRP);
  if (c > 0) flush(c - 1);
}

// ------------------------------------------------------------------------------------------------------
// producer warp of one group
// ------------------------------------------------------------------------------------------------------
struct SlotRegs {
  int i, j, e, wb;   // gathered center, owner, canonical edge id (-1: dead slot), window base or NOSTAGE
  int qnode, qflags;
  int nq;
};

template <int C, int M1, int M2>
__device__ __forceinline__ void bwd_producer(const NeighborArgs& A, BwdSmem& sm, const uint32_t tmem, const uint32_t tiles_base,
                                             const int grp) {
  using W_ = Win<C, M1, M2>;
  constexpr int D = C + 3 * M1 + 5 * M2;
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
    float pi[3], pj[3], sh[3];
  };
  auto stage_b = [&](const SlotRegs& r, PosRegs& p) {
#pragma unroll
    for (int x = 0; x < 3; ++x) p.pi[x] = p.pj[x] = p.sh[x] = 0.f;
    if (r.e >= 0) {
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        p.pi[x] = A.geo.pos[3 * r.i + x];
        p.pj[x] = A.geo.pos[3 * r.j + x];
      }
      if (g.offsets != nullptr) {
        const char4 o = reinterpret_cast<const char4*>(g.offsets)[r.e];
        const float* cl = g.cell + 9 * (g.node_graph ? g.node_graph[r.j] : 0);
        const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
#pragma unroll
        for (int x = 0; x < 3; ++x) p.sh[x] = ox * cl[x] + oy * cl[3 + x] + oz * cl[6 + x];
      }
    }
  };

  auto stage_c = [&](int c, const SlotRegs& r, const PosRegs& p) {
    Geo& ge = sm.geo[grp][c % NGEO];
    if (r.nq > 0 && lane < SLOTS) {
      float4 y0 = make_float4(0.f, 0.f, 0.f, 0.f), y1 = y0, y2 = y0, rad = y0, uu = y0;
      if (r.e >= 0) {
        float rv[3], d, u[3], Y[8], Gm[3][8];
#pragma unroll
        for (int x = 0; x < 3; ++x) rv[x] = (p.pi[x] - p.pj[x]) - p.sh[x];
        unit_vector(rv, d, u);
        angular_first(u, d, Y, Gm);
        y0 = make_float4(Y[0], Y[1], Y[2], 0.f);
        y1 = make_float4(Y[3], Y[4], Y[5], 0.f);
        y2 = make_float4(Y[6], Y[7], 0.f, 0.f);
        const Cutoff<float> ct = cutoff_terms(d, A.geo.rc);
        rad = make_float4(d, ct.chi, ct.dchi, 0.f);
        uu = make_float4(u[0], u[1], u[2], 0.f);
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          *reinterpret_cast<float4*>(&ge.G[lane][x * 8]) = make_float4(Gm[x][0], Gm[x][1], Gm[x][2], Gm[x][3]);
          *reinterpret_cast<float4*>(&ge.G[lane][x * 8 + 4]) = make_float4(Gm[x][4], Gm[x][5], Gm[x][6], Gm[x][7]);
        }
      }
      ge.Yt[lane][0] = y0;
      ge.Yt[lane][1] = y1;
      ge.Yt[lane][2] = y2;
      ge.rad[lane] = rad;
      ge.u[lane] = uu;
      ge.eid[lane] = r.e;
      const int ii = r.e >= 0 ? r.i : r.j;  // dead slots gather the (always valid) rows of the owner, times zero
      if (r.wb != NOSTAGE) {
        const uint32_t row = (uint32_t)(r.wb + ii), buf = row / (uint32_t)WH, local = row - buf * WH;  // wb + node = buf * WH + local row
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

// window loader: raw gx / gV rows of the staged tiles of this CTA, two bulk copies per tile
template <int C, int M1, int M2>
__device__ __forceinline__ void bwd_loader(const NeighborArgs& A, BwdSmem& sm, const uint32_t win_base) {
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

template <int C, int M1, int M2>
__global__ void __launch_bounds__(NTHREADS, 1) nbr_bwd_ul_kernel(const NeighborArgs A) {
  static_assert(C % SL_C == 0 && M1 == C / 2 && M2 == C / 4, "channel slices of edge_mma.cuh");
  __shared__ BwdSmem sm;
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
  if (warp < 4 * G) bwd_consumer<C, M1, M2>(A, sm, tmem, tiles_base, win_base, warp >> 2);
  else if (warp < 4 * G + G) bwd_producer<C, M1, M2>(A, sm, tmem, tiles_base, warp - 4 * G);
  else bwd_loader<C, M1, M2>(A, sm, win_base);
  tmem_teardown(tmem);
}

template <int C>
static int launch_nbr_bwd_ul_t(const NeighborArgs& A, cudaStream_t st) {
  constexpr int M1 = C / 2, M2 = C / 4, SLICES = C / SL_C;
  using W_ = Win<C, M1, M2>;
  static_assert(sizeof(BwdSmem) <= 40 * 1024, "static shared memory budget");
  const xeq_graph_t& g = A.geo.g;
  const bool window = W_::ENABLED && g.tile_mode == 1;
  const size_t dyn_max = 1024 + (size_t)G * NBST * BSTAGE + (W_::ENABLED ? (size_t)2 * W_::HALF : 0);
  const size_t dyn = 1024 + (size_t)G * NBST * BSTAGE + (window ? (size_t)2 * W_::HALF : 0);
  XEQ_CUDA(cudaFuncSetAttribute(nbr_bwd_ul_kernel<C, M1, M2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_max));
  const int work = g.tile_mode == 1 ? g.t_n_tiles : (g.t_n_tiles + G - 1) / G;
  const int grid = max(1, min(work, num_sms() / SLICES));
  XEQ_CUDA(launch_pdl(nbr_bwd_ul_kernel<C, M1, M2>, dim3(grid, SLICES), dim3(NTHREADS), dyn, st, A));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

}  // namespace

// `wide`: 256x0e + 128x1o + 64x2e (two channel slices per tile of edges, grid.y = 2; one d/dr slab per slice)
int launch_nbr_bwd_ul(const NeighborArgs& A, bool wide, cudaStream_t st) {
  return wide ? launch_nbr_bwd_ul_t<256>(A, st) : launch_nbr_bwd_ul_t<128>(A, st);
}

}  // namespace xeq
