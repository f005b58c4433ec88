// K2 forward message, round-2 design: three consumer groups of 4 warps share one copy of the filter rows in
// tensor memory (edge_ul.cuh: "unified lanes"), every group is fed by its own producer warp, and nothing in
// the steady state is a CTA-wide barrier -- all hand-offs are mbarriers:
//
//   warps  0-11  consumers, group g = warp / 4 : (a) radial stage of chunk c+1 in the shadow of the MMAs of chunk c:
//                96 threads each turn (d, chi, amp) of one slot into four chi * phi_k values and store them with
//                128-bit stores straight into the SWIZZLE_128B B tile (3xTF32 hi / lo); (b) wait acc_full[g]; per
//                quad of 4 edge slots: 5 tcgen05.ld (filter values of the lane's five rows), 2 x 128-bit gathers
//                of the packed neighbour row per slot (shared-memory window or L2), 10 FMAs per slot,
//                register-resident segment sums, row write-out at the end of a CSR row; arrive acc_free[g]
//   warps 12-14  producer of group g : walks the group's rows (dense 4-slot quads, 16 slots per chunk), runs the
//                per-edge geometry (d, chi, harmonics, gather offsets) as a software pipeline over chunks
//                (indices | positions | arithmetic) two chunks ahead, and one elected lane issues the 45
//                tcgen05.mma of a chunk into the group's accumulator buffer once its tiles are written
//   warp  15     window loader : one elected thread streams the packed rows of the CTA's molecule tiles into
//                the two halves of the shared-memory window with cp.async.bulk (TMA), a tile ahead
//
// Replaces nn/xpainn.py:66-74, 140-159 + nn/basic.py:114-131 (forward values); contract: xeq_edge_message_fwd.
#include "edge_ul.cuh"

namespace xeq {

using namespace fm;
using namespace ul;

namespace {

constexpr int SLOTS = 16;                    // edge slots per chunk = MMA N
constexpr int NQ = SLOTS / 4;                // quads per chunk
constexpr int DCOLS = TILES * SLOTS;         // accumulator columns of one group (80)
constexpr int BSTAGE = 2 * SLOTS * 128;      // bytes of one B stage: hi + lo tile
constexpr int NBST = 2;                      // B stages per group
constexpr int NGEO = 4;                      // geometry records per group (ring; the producer runs two chunks ahead)
constexpr int NTHREADS = NCONS + G * 32 + 32;
constexpr int NOSTAGE = INT_MIN;

// MODE 0: forward message.  The JVP half of the double backward (tangent of the message along (a_s, a_v, a_pos, a_cell))
// is two more instances of the same kernel, because the message is linear in the gathered rows and in the filter:
//   MODE 1 (tangent rows): the packed rows carry (sdot v + s vdot, sdot) instead of (s v, s), plus the term
//          w_edge s_edge Ydot of the rotating harmonics (Ydot = (dY/dr)^T rdot, s_edge rides in the spare packed entry);
//   MODE 2 (tangent filter): the rows are the forward's, the radial tiles carry ddot psi'_k(d) instead of psi_k(d), so
//          the MMA yields wdot = ddot w'(d); it adds onto the output of the MODE 1 launch (x_in = x_out).
enum : int { MODE_FWD = 0, MODE_TAN = 1, MODE_DW = 2 };

template <int MODE>
struct alignas(16) Geo {
  // harmonics per slot: MODE 0 / 2: one float4 per piece type; MODE 1: (Y0 Y1 Y2 Y6) (Y3 Y4 Y5 Y7) and the same of Ydot
  float4 Yt[SLOTS][MODE == MODE_TAN ? 4 : 3];
  // MODE 0 / 1: (d, chi * sqrt(2 / rc) / (d + 1e-5), chi, -); MODE 2: (d, ddot c0 dchi / (d + 1e-5), ddot c0 chi / (d + 1e-5),
  // 1 / (d + 1e-5)); zeros for dead slots
  float4 rad[SLOTS];
  uint32_t goff[SLOTS];  // staged: byte offset of the gathered row inside the window; else: node index
  Quad qd[NQ];
  int nq;                // quads of this chunk (1..NQ), -1 = end of stream
  int pad[3];
};

template <int MODE>
struct FwdSmem {
  Geo<MODE> geo[G][NGEO];
  uint64_t geo_full[G][NGEO];   // producer -> consumers: geometry record of a chunk is written
  uint64_t tile_full[G][NBST];  // consumers -> producer: radial tiles of a chunk are written (and proxy-fenced)
  uint64_t acc_full[G], acc_free[G];
  uint64_t win_full[2], win_free[2];
  uint32_t slot;
};

// ------------------------------------------------------------------------------------------------------
// pack: pk[sl][n][plane][lane] (float4) from s [N,H] and v [N,D] (cm layout)
//   plane 0: (s_state[q0] v[q0], s_edge[q0], s_scalar[q0], s_edge[qp])     plane 1: s_state[qp] v[(qp, m)], m = 0..2
// ------------------------------------------------------------------------------------------------------
// packed entry of lane (q0, qp, off) of node row (sn, vn); MODE 1: the tangent entries from (asn, avn) (either may be NULL)
template <int M, int MODE>
__device__ __forceinline__ void pack_entry(const float* __restrict__ sn, const float* __restrict__ vn, const float* __restrict__ asn,
                                           const float* __restrict__ avn, int q0, int qp, const int (&off)[3], int nc, float4& a, float4& b) {
  const float ssp = sn[qp];
  if (MODE != MODE_TAN) {
    a.x = sn[q0] * vn[q0];
    a.y = sn[M + q0];
    a.z = sn[2 * M + q0];
    a.w = sn[M + qp];
    b.x = ssp * vn[off[0]];
    b.y = ssp * vn[off[1]];
    b.z = nc == 3 ? ssp * vn[off[2]] : 0.f;
    b.w = 0.f;
  } else {
    const float sd0 = asn ? asn[q0] : 0.f, sdp = asn ? asn[qp] : 0.f;
    a.x = fmaf(sd0, vn[q0], avn ? sn[q0] * avn[q0] : 0.f);
    a.y = asn ? asn[M + q0] : 0.f;
    a.z = asn ? asn[2 * M + q0] : 0.f;
    a.w = asn ? asn[M + qp] : 0.f;
    b.x = fmaf(sdp, vn[off[0]], avn ? ssp * avn[off[0]] : 0.f);
    b.y = fmaf(sdp, vn[off[1]], avn ? ssp * avn[off[1]] : 0.f);
    b.z = nc == 3 ? fmaf(sdp, vn[off[2]], avn ? ssp * avn[off[2]] : 0.f) : 0.f;
    b.w = sn[M + qp];  // s_edge of the piece: multiplies w_edge Ydot
  }
}

template <int C, int M1, int M2, int MODE>
__global__ void __launch_bounds__(256) pack_fwd_kernel(const float* __restrict__ s, const float* __restrict__ v, const float* __restrict__ a_s,
                                                      const float* __restrict__ a_v, float* __restrict__ pk, int n_nodes) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M;
  const int L = threadIdx.x & 127, sl = blockIdx.y;
  const int n = blockIdx.x * 2 + (threadIdx.x >> 7);
  if (n >= n_nodes) return;
  const int q0 = sl * SL_C + L, qp = piece_irrep<C, M1>(L, sl);
  int off[3], nc;
  piece_offsets<C, M1, M2>(L, sl, off, nc);
  float4 a, b;
  pack_entry<M, MODE>(s + (size_t)n * H, v + (size_t)n * D, a_s ? a_s + (size_t)n * H : nullptr, a_v ? a_v + (size_t)n * D : nullptr,
                      q0, qp, off, nc, a, b);
  float4* dst = reinterpret_cast<float4*>(pk + ((size_t)sl * n_nodes + n) * ROWF);
  dst[L] = a;
  dst[128 + L] = b;
}

// ------------------------------------------------------------------------------------------------------
// consumers
// ------------------------------------------------------------------------------------------------------
template <int C, int M1, int M2, int MODE>
__device__ __forceinline__ void fwd_consumer(const CenterArgs& A, FwdSmem<MODE>& sm, const uint32_t tmem, const uint32_t tiles_base,
                                             const uint32_t win_base, const float* __restrict__ pk, const int grp) {
  constexpr int D = C + 3 * M1 + 5 * M2;
  using GeoT = Geo<MODE>;
  const int L = threadIdx.x - grp * GRP, wq = L >> 5, lane = L & 31, sl = blockIdx.y;
  const int pt = piece_type(L);
  const int q0 = sl * SL_C + L;
  int voff[3], nc;
  piece_offsets<C, M1, M2>(L, sl, voff, nc);
  const uint32_t lane_base = tmem + ((uint32_t)(32 * wq) << 16);
  const uint32_t dbase = lane_base + D_COL + grp * DCOLS;
  const uint32_t win_lane = win_base + 16u * (uint32_t)L;
  const float* pk_lane = pk + (size_t)sl * A.geo.g.n_nodes * ROWF + 4 * L;
  const uint32_t full = smem_u32(&sm.acc_full[grp]), free_ = smem_u32(&sm.acc_free[grp]);
  const uint32_t geo0 = smem_u32(&sm.geo[grp][0]), gfull0 = smem_u32(&sm.geo_full[grp][0]);
  const uint32_t tfull0 = smem_u32(&sm.tile_full[grp][0]);
  const uint32_t my_tiles = tiles_base + (uint32_t)grp * (NBST * BSTAGE);

  // radial stage: thread L < 96 owns slot L / 6 and the four radial terms k = 4 (L % 6) .. + 3 of every chunk
  // (k = 0: the cutoff / bias term, k = 1 .. 20: Bessel terms, k > 20: padding)
  const int rslot = L / 6, rkc = L - 6 * rslot;
  float fr[4];
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int k = 4 * rkc + x;
    fr[x] = (L < 96 && k >= 1 && k <= NB_) ? A.geo.freq[k - 1] : 0.f;
  }
  const float inv_c0 = sqrtf(0.5f * A.geo.rc);
  const uint32_t rad_off = (uint32_t)offsetof(GeoT, rad) + 16u * (uint32_t)rslot;
  const uint32_t tile_off = (uint32_t)(rslot * 128 + ((rkc ^ (rslot & 7)) << 4));
  auto radial = [&](int c) {  // tiles of chunk c (its geometry record is visible)
    if (L < 96) {
      const float4 rd = lds128(geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(GeoT) + rad_off);
      float val[4];
      if (MODE != MODE_DW) {
#pragma unroll
        for (int x = 0; x < 4; ++x) val[x] = rd.y * sin_reduced(fr[x] * rd.x);  // fr = 0 -> exactly zero
        if (rkc == 0) val[0] = rd.z;
      } else {  // ddot psi'_k = ddot (dchi phi_k + chi phi'_k): rd = (d, P = ddot c0 dchi inv, Q = ddot c0 chi inv, inv)
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          float sn, cs;
          sincos_reduced(fr[x] * rd.x, sn, cs);
          val[x] = fmaf(rd.y, sn, rd.z * fmaf(fr[x], cs, -sn * rd.w));  // fr = 0 -> exactly zero
        }
        if (rkc == 0) val[0] = rd.y * inv_c0 * (rd.x + 1e-5f);  // k = 0: ddot dchi
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) split_fast(val[x], hi[x], lo[x]);
      const uint32_t t_hi = my_tiles + (uint32_t)(c & (NBST - 1)) * BSTAGE + tile_off;
      sts128(t_hi, hi[0], hi[1], hi[2], hi[3]);
      sts128(t_hi + SLOTS * 128, lo[0], lo[1], lo[2], lo[3]);
      proxy_fence();  // generic-proxy stores -> visible to the tensor core
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

  float accx = 0.f, accV0 = 0.f, accP[3] = {0.f, 0.f, 0.f};
  float bx = 0.f, bV0 = 0.f, bP[3] = {0.f, 0.f, 0.f};

  geo_wait(0);
  int nq = geo_nq(0);
  if (nq > 0) radial(0);
  for (int c = 0; nq >= 0; ++c) {
    // radial terms of the next chunk, in the shadow of this chunk's MMAs
    geo_wait(c + 1);
    const int nq_next = geo_nq(c + 1);
    if (nq_next > 0) radial(c + 1);
    mbar_wait(full, (uint32_t)(c & 1));
    tc_fence_after();
    const uint32_t ge = geo0 + (uint32_t)(c % NGEO) * (uint32_t)sizeof(GeoT);
#pragma unroll 1
    for (int qd = 0; qd < nq; ++qd) {
      int node, fl;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(node), "=r"(fl) : "r"(ge + (uint32_t)offsetof(GeoT, qd) + 8u * (uint32_t)qd) : "memory");
      if (fl & F_TILE_FIRST) {
        if (fl & F_STAGED) mbar_wait(smem_u32(&sm.win_full[(fl & F_BUF) ? 1 : 0]), (fl & F_PAR) ? 1u : 0u);
      }
      if (fl & F_ROW_FIRST) {  // residual row: requested now, consumed when the row ends
        accx = accV0 = accP[0] = accP[1] = accP[2] = 0.f;
        const size_t nd = (size_t)node;
        bx = A.x_in ? A.x_in[nd * C + q0] : 0.f;
        bV0 = A.V_in ? A.V_in[nd * D + q0] : 0.f;
#pragma unroll
        for (int m = 0; m < 3; ++m) bP[m] = A.V_in ? A.V_in[nd * D + voff[m]] : 0.f;
      }
      if (!(fl & F_NOROW)) {
        float w[TILES][4];
#pragma unroll
        for (int t = 0; t < TILES; ++t) tmem_ld4(dbase + t * SLOTS + qd * 4, w[t]);
        uint32_t gj[4];
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(gj[0]), "=r"(gj[1]), "=r"(gj[2]), "=r"(gj[3])
                     : "r"(ge + (uint32_t)offsetof(GeoT, goff) + 16u * (uint32_t)qd) : "memory");
        float4 a[4], b[4], y[4], yd[4];
        if (fl & F_STAGED) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a[j] = lds128(win_lane + gj[j]);
            b[j] = lds128(win_lane + gj[j] + 2048u);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float* p = pk_lane + (size_t)gj[j] * ROWF;
            a[j] = ldg128(p);
            b[j] = ldg128(p + 512);
          }
        }
        if (MODE != MODE_TAN) {
#pragma unroll
          for (int j = 0; j < 4; ++j) y[j] = lds128(ge + (uint32_t)offsetof(GeoT, Yt) + 48u * (uint32_t)(qd * 4 + j) + 16u * (uint32_t)pt);
        } else {  // warp-uniform: pieces 0 / 1 take one float4 of Y and one of Ydot, piece 2 the .w entries of all four
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t yb = ge + (uint32_t)offsetof(GeoT, Yt) + 64u * (uint32_t)(qd * 4 + j);
            if (pt < 2) {
              y[j] = lds128(yb + 16u * (uint32_t)pt);
              yd[j] = lds128(yb + 32u + 16u * (uint32_t)pt);
            } else {
              y[j] = make_float4(lds32(yb + 12u), lds32(yb + 28u), 0.f, 0.f);
              yd[j] = make_float4(lds32(yb + 44u), lds32(yb + 60u), 0.f, 0.f);
            }
          }
        }
        tmem_wait_ld();
#pragma unroll
        for (int t = 0; t < TILES; ++t) pin(w[t]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          accx = fmaf(a[j].z, w[2][j], accx);
          accV0 = fmaf(a[j].x, w[0][j], fmaf(a[j].y, w[1][j], accV0));
          const float gep = a[j].w * w[4][j];
          accP[0] = fmaf(b[j].x, w[3][j], fmaf(gep, y[j].x, accP[0]));
          accP[1] = fmaf(b[j].y, w[3][j], fmaf(gep, y[j].y, accP[1]));
          accP[2] = fmaf(b[j].z, w[3][j], fmaf(gep, y[j].z, accP[2]));
          if (MODE == MODE_TAN) {  // w_edge s_edge Ydot
            const float ged = b[j].w * w[4][j];
            accP[0] = fmaf(ged, yd[j].x, accP[0]);
            accP[1] = fmaf(ged, yd[j].y, accP[1]);
            accP[2] = fmaf(ged, yd[j].z, accP[2]);
          }
        }
      }
      if (fl & F_ROW_LAST) {
        const size_t nd = (size_t)node;
        A.x_out[nd * C + q0] = bx + accx;
        A.V_out[nd * D + q0] = bV0 + accV0;
        A.V_out[nd * D + voff[0]] = bP[0] + accP[0];
        A.V_out[nd * D + voff[1]] = bP[1] + accP[1];
        if (nc == 3) A.V_out[nd * D + voff[2]] = bP[2] + accP[2];
      }
      if ((fl & F_TILE_LAST) && (fl & F_STAGED)) {  // this warp is done with the window of the tile
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm.win_free[(fl & F_BUF) ? 1 : 0]));
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(free_);
    nq = nq_next;
  }
}

// ------------------------------------------------------------------------------------------------------
// producer warp of one group
// ------------------------------------------------------------------------------------------------------
struct SlotRegs {
  int i, e, wb;  // center node, edge id (-1: dead slot), window base (rows) or NOSTAGE
  int j;         // neighbor
  int qnode, qflags;  // lane q < NQ: quad q of the chunk
  int nq;             // quads of the chunk (1..NQ), -1 = end of stream (uniform)
};

template <int C, int M1, int M2, int MODE>
__device__ __forceinline__ void fwd_producer(const CenterArgs& A, FwdSmem<MODE>& sm, const uint32_t tmem, const uint32_t tiles_base,
                                             const int grp) {
  const int lane = threadIdx.x & 31, slot = lane & 15, half = lane >> 4;
  const xeq_graph_t& g = A.geo.g;
  const uint32_t full = smem_u32(&sm.acc_full[grp]), free_ = smem_u32(&sm.acc_free[grp]);
  const uint32_t gfull0 = smem_u32(&sm.geo_full[grp][0]), tfull0 = smem_u32(&sm.tile_full[grp][0]);
  const uint32_t my_tiles = tiles_base + (uint32_t)grp * (NBST * BSTAGE);
  const float c0 = sqrtf(2.f / A.geo.rc);
  const float pi_rc = 3.14159265358979323846f / A.geo.rc;

  Walk wk;
  wk.init(g, g.tile_ptr, g.n_tiles, grp);
  int node = wk.valid ? wk.n0 + wk.rphase : 0;
  int e = 0, e1 = 0;
  bool row_open = false, row_first = false, tile_any = false;

  // stage A: slot assignment of the next chunk (uniform control flow), quads kept in registers, index loads
  auto stage_a = [&](SlotRegs& o) {
    o.i = 0; o.e = -1; o.wb = NOSTAGE; o.j = 0; o.qnode = 0; o.qflags = 0;
    int nq = 0;
    while (nq < NQ && wk.valid) {
      const int stbits = wk.staged ? (F_STAGED | (wk.buf ? F_BUF : 0) | (wk.par ? F_PAR : 0)) : 0;
      if (!row_open) {
        if (node >= wk.n1) {
          if (!tile_any && wk.tile_mode == 1) {  // no row of this group in the tile: keep the window accounting going
            if (lane == nq) { o.qnode = wk.n0; o.qflags = stbits | F_NOROW | F_TILE_FIRST | F_TILE_LAST; }
            ++nq;
          }
          wk.next();
          node = wk.valid ? wk.n0 + wk.rphase : 0;
          tile_any = false;
          continue;
        }
        e = g.rowptr[node];
        e1 = g.rowptr[node + 1];
        row_open = true;
        row_first = true;
      }
      const bool last = e + 4 >= e1;
      int fl = stbits | (row_first ? F_ROW_FIRST : 0) | (tile_any ? 0 : F_TILE_FIRST);
      if (last) fl |= F_ROW_LAST | ((node + wk.rstride >= wk.n1) ? F_TILE_LAST : 0);
      if (lane == nq) { o.qnode = node; o.qflags = fl; }
      const int idx = slot - 4 * nq;
      if (idx >= 0 && idx < 4) {
        o.i = node;
        o.e = (e + idx < e1) ? e + idx : -1;
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
    o.nq = nq ? nq : -1;  // -1: stream exhausted
    if (o.e >= 0) o.j = g.col[o.e];
  };

  // stage B: raw position loads (and the lattice shift of periodic graphs)
  struct PosRegs {
    float pi[3], pj[3], sh[3];
    float rd[3];  // MODE 1 / 2: tangent of the edge vector, a_pos[i] - a_pos[j] - offsets @ a_cell
  };
  auto stage_b = [&](const SlotRegs& r, PosRegs& p) {
#pragma unroll
    for (int x = 0; x < 3; ++x) p.pi[x] = p.pj[x] = p.sh[x] = p.rd[x] = 0.f;
    if (r.e >= 0) {
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        p.pi[x] = A.geo.pos[3 * r.i + x];
        p.pj[x] = A.geo.pos[3 * r.j + x];
        if (MODE != MODE_FWD && A.geo.a_pos) p.rd[x] = A.geo.a_pos[3 * r.i + x] - A.geo.a_pos[3 * r.j + x];
      }
      if (g.offsets != nullptr) {  // nn/basic.py:119-128: vectors -= cell_offsets @ cell[graph(neighbor)]
        const char4 o = reinterpret_cast<const char4*>(g.offsets)[r.e];
        const int gi = g.node_graph ? g.node_graph[r.j] : 0;
        const float* cl = g.cell + 9 * gi;
        const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
#pragma unroll
        for (int x = 0; x < 3; ++x) p.sh[x] = ox * cl[x] + oy * cl[3 + x] + oz * cl[6 + x];
        if (MODE != MODE_FWD && A.geo.a_cell) {
          const float* ac = A.geo.a_cell + 9 * gi;
#pragma unroll
          for (int x = 0; x < 3; ++x) p.rd[x] -= ox * ac[x] + oy * ac[3 + x] + oz * ac[6 + x];
        }
      }
    }
  };

  // stage C: geometry record of chunk c -> shared memory, then the release to the consumers (their radial stage)
  auto stage_c = [&](int c, const SlotRegs& r, const PosRegs& p) {
    Geo<MODE>& ge = sm.geo[grp][c % NGEO];
    if (r.nq > 0 && half == 0) {
      float4 y0 = make_float4(0.f, 0.f, 0.f, 0.f), y1 = y0, y2 = y0, y3 = y0, rad = y0;
      if (r.e >= 0) {
        float rv[3], d, u[3], Y[8];
#pragma unroll
        for (int x = 0; x < 3; ++x) rv[x] = (p.pi[x] - p.pj[x]) - p.sh[x];
        unit_vector(rv, d, u);
        float chi = 0.f, dchi = 0.f;
        if (d < A.geo.rc) {
          if (MODE != MODE_DW) {
            chi = 0.5f * (cosf(pi_rc * d) + 1.f);
          } else {
            float sn, cs;
            sincosf(pi_rc * d, &sn, &cs);
            chi = 0.5f * (cs + 1.f);
            dchi = -0.5f * pi_rc * sn;
          }
        }
        const float inv = 1.f / (d + 1e-5f);
        if (MODE == MODE_TAN) {  // Ydot = (dY/dr)^T rdot
          float Gm[3][8], Yd[8];
          angular_first(u, d, Y, Gm);
#pragma unroll
          for (int m = 0; m < 8; ++m) Yd[m] = Gm[0][m] * p.rd[0] + Gm[1][m] * p.rd[1] + Gm[2][m] * p.rd[2];
          y0 = make_float4(Y[0], Y[1], Y[2], Y[6]);
          y1 = make_float4(Y[3], Y[4], Y[5], Y[7]);
          y2 = make_float4(Yd[0], Yd[1], Yd[2], Yd[6]);
          y3 = make_float4(Yd[3], Yd[4], Yd[5], Yd[7]);
        } else {
          sph_harm(u, Y);
          y0 = make_float4(Y[0], Y[1], Y[2], 0.f);
          y1 = make_float4(Y[3], Y[4], Y[5], 0.f);
          y2 = make_float4(Y[6], Y[7], 0.f, 0.f);
        }
        if (MODE != MODE_DW) {
          rad = make_float4(d, chi * c0 / (d + 1e-5f), chi, 0.f);
        } else {
          const float ddot = u[0] * p.rd[0] + u[1] * p.rd[1] + u[2] * p.rd[2];
          rad = make_float4(d, ddot * c0 * dchi * inv, ddot * c0 * chi * inv, inv);
        }
      }
      ge.Yt[slot][0] = y0;
      ge.Yt[slot][1] = y1;
      ge.Yt[slot][2] = y2;
      if (MODE == MODE_TAN) ge.Yt[slot][3] = y3;
      ge.rad[slot] = rad;
      const int jj = r.e >= 0 ? r.j : r.i;  // dead slots gather the (always valid) row of their own center, times zero
      ge.goff[slot] = (r.wb != NOSTAGE) ? (uint32_t)(r.wb + jj) * (uint32_t)ROWB : (uint32_t)jj;
      if (lane < NQ) ge.qd[lane] = Quad{r.qnode, r.qflags};
    }
    if (lane == 0) ge.nq = r.nq;
    __syncwarp();
    if (lane == 0) mbar_arrive(gfull0 + 8u * (uint32_t)(c % NGEO));
  };

  auto issue = [&](int c) {
    const uint32_t idesc = idesc_tf32(SLOTS);
    const uint32_t b_hi = my_tiles + (uint32_t)(c & (NBST - 1)) * BSTAGE, b_lo = b_hi + SLOTS * 128;
    const uint32_t d0 = tmem + D_COL + (uint32_t)grp * DCOLS;
#pragma unroll
    for (int ks = 0; ks < NBP / 8; ++ks) {
      const uint64_t db_hi = smem_desc(b_hi + ks * 32), db_lo = smem_desc(b_lo + ks * 32);
#pragma unroll
      for (int tile = 0; tile < TILES; ++tile) {
        const uint32_t d = d0 + tile * SLOTS;
        mma_ts(d, tmem + A_LO + tile * NBP + ks * 8, db_hi, idesc, ks ? 1u : 0u);
        mma_ts(d, tmem + A_HI + tile * NBP + ks * 8, db_lo, idesc, 1u);
        mma_ts(d, tmem + A_HI + tile * NBP + ks * 8, db_hi, idesc, 1u);
      }
    }
  };

  // geometry pipeline: record of chunk c+2 written in iteration c, positions of chunk c+3 and indices of chunk c+4 in flight
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
    mbar_wait(tfull0 + 8u * (uint32_t)(c & (NBST - 1)), (uint32_t)((c >> 1) & 1));  // radial tiles of chunk c
    if (c > 0) mbar_wait(free_, (uint32_t)((c - 1) & 1));                             // accumulators drained
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

// window loader: one elected lane streams the packed rows of the staged tiles of this CTA into the window
template <int C, int MODE>
__device__ __forceinline__ void fwd_loader(const CenterArgs& A, FwdSmem<MODE>& sm, const uint32_t win_base, const float* __restrict__ pk) {
  const xeq_graph_t& g = A.geo.g;
  if (g.tile_mode != 1) return;
  if ((threadIdx.x & 31) != 0) return;
  Walk wk;
  wk.init(g, g.tile_ptr, g.n_tiles, 0);
  const float* pk_sl = pk + (size_t)blockIdx.y * g.n_nodes * ROWF;
  for (; wk.valid; wk.next()) {
    if (!wk.staged) continue;
    const int t = wk.staged_count - 1;  // index of this staged tile
    const uint32_t fullb = smem_u32(&sm.win_full[wk.buf]), freeb = smem_u32(&sm.win_free[wk.buf]);
    if (t >= 2) mbar_wait_sleep(freeb, (uint32_t)(((t >> 1) - 1) & 1));
    const uint32_t bytes = (uint32_t)(wk.n1 - wk.n0) * ROWB;
    mbar_expect_tx(fullb, bytes);
    tma_bulk_g2s(win_base + (uint32_t)wk.buf * (WH * ROWB), pk_sl + (size_t)wk.n0 * ROWF, bytes, fullb);
  }
}

// window packer (all tiles of the launch fit the window: xeq_graph_t.max_tile_nodes <= WH): warp 15 reads the RAW s / v
// rows of the CTA's next tile (coalesced 128-byte loads), forms the per-lane packed entries and writes them straight
// into the free window half -- the packing pass and its HBM round trip (write + re-read of 4 KB per node) disappear.
template <int C, int M1, int M2, int MODE>
__device__ __forceinline__ void fwd_packer(const CenterArgs& A, FwdSmem<MODE>& sm, const uint32_t win_base) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M;
  const xeq_graph_t& g = A.geo.g;
  const int lane = threadIdx.x & 31, sl = blockIdx.y;
  int q0[4], qp[4], off[4][3], nc[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int L = lane + 32 * it;
    q0[it] = sl * SL_C + L;
    qp[it] = piece_irrep<C, M1>(L, sl);
    piece_offsets<C, M1, M2>(L, sl, off[it], nc[it]);
  }
  Walk wk;
  wk.init(g, g.tile_ptr, g.n_tiles, 0);
  for (; wk.valid; wk.next()) {
    if (!wk.staged) continue;
    const int t = wk.staged_count - 1;
    const uint32_t fullb = smem_u32(&sm.win_full[wk.buf]), freeb = smem_u32(&sm.win_free[wk.buf]);
    if (t >= 2) mbar_wait_sleep(freeb, (uint32_t)(((t >> 1) - 1) & 1));
    const uint32_t half = win_base + (uint32_t)wk.buf * (WH * ROWB) + 16u * (uint32_t)lane;
#pragma unroll 2
    for (int n = wk.n0; n < wk.n1; ++n) {
      const float* sn = A.s + (size_t)n * H;
      const float* vn = A.v + (size_t)n * D;
      const float* asn = (MODE == MODE_TAN && A.a_s) ? A.a_s + (size_t)n * H : nullptr;
      const float* avn = (MODE == MODE_TAN && A.a_v) ? A.a_v + (size_t)n * D : nullptr;
      const uint32_t row = half + (uint32_t)(n - wk.n0) * ROWB;
      float4 a[4], b[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) pack_entry<M, MODE>(sn, vn, asn, avn, q0[it], qp[it], off[it], nc[it], a[it], b[it]);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        sts128(row + 512u * it, __float_as_uint(a[it].x), __float_as_uint(a[it].y), __float_as_uint(a[it].z), __float_as_uint(a[it].w));
        sts128(row + 2048u + 512u * it, __float_as_uint(b[it].x), __float_as_uint(b[it].y), __float_as_uint(b[it].z), __float_as_uint(b[it].w));
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(fullb);  // release: the rows of this tile are in the window
  }
}

template <int C, int M1, int M2, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) center_fwd_ul_kernel(const CenterArgs A, const float* __restrict__ pk) {
  static_assert(C % SL_C == 0 && M1 == C / 2 && M2 == C / 4, "channel slices of edge_mma.cuh");
  __shared__ FwdSmem<MODE> sm;
  pdl_trigger();  // the next kernel of the stream may be staged; it waits for this grid before it reads memory
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) {
    for (int i = 0; i < G; ++i) {
      mbar_init(smem_u32(&sm.acc_full[i]), 1);   // tcgen05.commit
      mbar_init(smem_u32(&sm.acc_free[i]), 4);   // one arrive per consumer warp of the group
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
  pdl_wait();  // barrier setup and the TMEM allocation overlapped the previous kernel's tail; its results are visible now
  if (t < GRP) store_filter_rows<C, M1, M2>(A.W, A.b, t, blockIdx.y, tmem + ((uint32_t)(32 * (t >> 5)) << 16));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4 * G) fwd_consumer<C, M1, M2, MODE>(A, sm, tmem, tiles_base, win_base, pk, warp >> 2);
  else if (warp < 4 * G + G) fwd_producer<C, M1, M2, MODE>(A, sm, tmem, tiles_base, warp - 4 * G);
  else if (pk != nullptr) fwd_loader<C, MODE>(A, sm, win_base, pk);   // packed rows from the packing pass, TMA bulk copies
  else fwd_packer<C, M1, M2, MODE>(A, sm, win_base);                // every tile fits the window: packed in-kernel
  tmem_teardown(tmem);
}

}  // namespace

size_t center_fwd_ul_workspace_bytes(int n_nodes, bool wide) {
  return 256 + (size_t)(wide ? 2 : 1) * (size_t)(n_nodes > 0 ? n_nodes : 1) * ROWB;
}

template <int C, int MODE>
static int launch_center_ul_t(const CenterArgs& A, void* ws, cudaStream_t st) {
  constexpr int M1 = C / 2, M2 = C / 4, SLICES = C / SL_C;
  static_assert(sizeof(FwdSmem<MODE>) <= (MODE == MODE_TAN ? 17 : 16) * 1024, "static shared memory budget");
  const xeq_graph_t& g = A.geo.g;
  // when the caller vouches that every molecule tile fits the window, warp 15 packs the rows in-kernel and the packing
  // pass (and the consumers' global-memory path) is not needed
  const bool inline_pack = g.tile_mode == 1 && g.max_tile_nodes > 0 && g.max_tile_nodes <= WH;
  float* pk = inline_pack ? nullptr : reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  if (!inline_pack)
    pack_fwd_kernel<C, M1, M2, MODE><<<dim3((g.n_nodes + 1) / 2, SLICES), 256, 0, st>>>(A.s, A.v, A.a_s, A.a_v, pk, g.n_nodes);
  const size_t dyn = 1024 + (size_t)G * NBST * BSTAGE + (g.tile_mode == 1 ? (size_t)2 * WH * ROWB : 0);
  XEQ_CUDA(cudaFuncSetAttribute(center_fwd_ul_kernel<C, M1, M2, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(1024 + (size_t)G * NBST * BSTAGE + (size_t)2 * WH * ROWB)));
  const int work = g.tile_mode == 1 ? g.n_tiles : (g.n_tiles + G - 1) / G;
  const int grid = max(1, min(work, num_sms() / SLICES));
  XEQ_CUDA(launch_pdl(center_fwd_ul_kernel<C, M1, M2, MODE>, dim3(grid, SLICES), dim3(NTHREADS), dyn, st, A, (const float*)pk));
  XEQ_LAUNCHED(inline_pack ? 1 : 2);
  return XEQ_OK;
}

// `wide`: 256x0e + 128x1o + 64x2e (two channel slices per tile of edges, grid.y = 2)
int launch_center_fwd_ul(const CenterArgs& A, bool wide, void* ws, cudaStream_t st) {
  return wide ? launch_center_ul_t<256, MODE_FWD>(A, ws, st) : launch_center_ul_t<128, MODE_FWD>(A, ws, st);
}

// JVP half of the double backward: (x_out, V_out) = tangent of the message along (a_s, a_v, a_pos, a_cell).  Two
// launches of the forward kernel family (header of this file); the second one only when the geometry has a tangent.
int launch_center_jvp_ul(const CenterArgs& A, bool wide, void* ws, cudaStream_t st) {
  CenterArgs T = A;
  T.x_in = nullptr;
  T.V_in = nullptr;
  int rc = wide ? launch_center_ul_t<256, MODE_TAN>(T, ws, st) : launch_center_ul_t<128, MODE_TAN>(T, ws, st);
  if (rc || (!A.geo.a_pos && !A.geo.a_cell)) return rc;
  T.x_in = A.x_out;  // accumulate in place: a lane reads its residual entries at the start of a row and writes them at its end
  T.V_in = A.V_out;
  return wide ? launch_center_ul_t<256, MODE_DW>(T, ws, st) : launch_center_ul_t<128, MODE_DW>(T, ws, st);
}

}  // namespace xeq
