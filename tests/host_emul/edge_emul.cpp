// TEST INFRASTRUCTURE: serial CPU emulation of the fused edge kernels, built from the SAME
// per-edge / per-thread headers the sm_100a kernels use (xequinet_b200/csrc/edge_math.cuh,
// edge_thread.cuh), instantiated in float64.  tests/test_edge_math_host.py checks it against
// torch autograd (first and second order) of the oracle's edge message, so that only the
// parallel decomposition of the kernels is left to verify on the GPU.
// Not part of the product: nothing in xequinet_b200/ links or loads this file.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../xequinet_b200/csrc/edge_thread.cuh"

using namespace xeq;
typedef double T;

// 0: the SIMT kernels' decomposition (first / second with the weight-gradient accumulators in the thread).
// 1: the tcgen05 kernels' decomposition: filter values w, dw, ddw supplied from outside (the MMA), radial-only
//    d/dr coefficients for rows without an angular term recombined with u / rp per edge, and the weight gradients as
//    the outer products  pw (x) psi,  pw (x) xi  (first order) /  alpha (x) psi + (beta ddot) (x) dpsi  (second).
static int g_decomposition = 0;
extern "C" void emul_set_decomposition(int mode) { g_decomposition = mode; }

struct Dims { int C, m0, m1, m2, B; double rc; };

struct Geo {  // everything any kernel variant keeps per edge
  T psi[NBP], dpsi[NBP], ddpsi[NBP], xi[NBP], dxi[NBP];
  T Y[8], G[24], Hm[24], Ydot[8], u[3], rp[3], ddot, d;
};

static void edge_vector(const T* pos, int i, int j, const int8_t* off, const T* cell, T r[3]) {
  for (int x = 0; x < 3; ++x) r[x] = pos[3 * i + x] - pos[3 * j + x];
  if (off && cell)
    for (int x = 0; x < 3; ++x) r[x] -= off[0] * cell[0 + x] + off[1] * cell[3 + x] + off[2] * cell[6 + x];
}

static void make_geo(const T r[3], const T* rdot, const T* freq, const Dims& D, Geo& g) {
  unit_vector(r, g.d, g.u);
  T G2[3][8], H2[3][8];
  angular_first(g.u, g.d, g.Y, G2);
  for (int x = 0; x < 3; ++x) for (int m = 0; m < 8; ++m) g.G[x * 8 + m] = G2[x][m];
  if (rdot) {
    angular_second(g.u, g.d, rdot, G2, g.ddot, g.rp, g.Ydot, H2);
    for (int x = 0; x < 3; ++x) for (int m = 0; m < 8; ++m) g.Hm[x * 8 + m] = H2[x][m];
  } else {
    g.ddot = 0; for (int x = 0; x < 3; ++x) g.rp[x] = 0; for (int m = 0; m < 8; ++m) g.Ydot[m] = 0;
    for (int k = 0; k < 24; ++k) g.Hm[k] = 0;
  }
  Cutoff<T> c = cutoff_terms<T>(g.d, D.rc);
  for (int k = 0; k < NBP; ++k) g.psi[k] = g.dpsi[k] = g.ddpsi[k] = g.xi[k] = g.dxi[k] = 0;
  g.psi[0] = c.chi; g.dpsi[0] = c.dchi; g.ddpsi[0] = c.ddchi;
  for (int k = 0; k < D.B; ++k) {
    Radial<T> rr = radial_term<T>(g.d, freq[k], D.rc, c);
    g.psi[k + 1] = rr.psi; g.dpsi[k + 1] = rr.dpsi; g.ddpsi[k + 1] = rr.ddpsi; g.xi[k + 1] = rr.xi; g.dxi[k + 1] = rr.dxi;
  }
}

static void load_wrow(const T* W, const T* b, int h, int B, T* row) {
  for (int k = 0; k < NBP; ++k) row[k] = 0;
  row[0] = b[h];
  for (int k = 0; k < B; ++k) row[k + 1] = W[h * B + k];
}

// channel bookkeeping shared with the kernels' mapping (cm layout)
struct Chan { int l, u, q, vbase, vstride; };
static Chan chan_of_q(const Dims& D, int q) {
  Chan c;
  if (q < D.m0) { c.l = 0; c.u = q; c.vbase = q; c.vstride = 0; }
  else if (q < D.m0 + D.m1) { c.l = 1; c.u = q - D.m0; c.vbase = D.m0 + c.u; c.vstride = D.m1; }
  else { c.l = 2; c.u = q - D.m0 - D.m1; c.vbase = D.m0 + 3 * D.m1 + c.u; c.vstride = D.m2; }
  c.q = q;
  return c;
}

template <int L>
static void center_node(const Dims& D, int i, int q, const int* rowptr, const int* col, const int8_t* offs,
                        const T* cell, const T* pos, const T* s, const T* v, const T* W, const T* b, const T* freq,
                        const T* a_s, const T* a_v, const T* a_pos, T* outx, T* outV) {
  const int M = D.m0 + D.m1 + D.m2, Dd = D.m0 + 3 * D.m1 + 5 * D.m2, H = D.C + 2 * M, nk = D.B + 1;
  Chan c = chan_of_q(D, q);
  CenterThread<T, L, 21> th;
  load_wrow(W, b, q, D.B, th.Ws);
  load_wrow(W, b, M + q, D.B, th.We);
  if (L == 0) load_wrow(W, b, 2 * M + q, D.B, th.Wx);
  th.reset();
  for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
    int j = col[e];
    T r[3], rdot[3];
    edge_vector(pos, i, j, offs ? offs + 4 * e : nullptr, cell, r);
    bool jv = a_pos != nullptr;
    if (jv) for (int x = 0; x < 3; ++x) rdot[x] = a_pos[3 * i + x] - a_pos[3 * j + x];
    Geo g; make_geo(r, jv ? rdot : nullptr, freq, D, g);
    T vv[5], vd[5];
    for (int m = 0; m < 2 * L + 1; ++m) { vv[m] = v[(size_t)j * Dd + c.vbase + m * c.vstride]; vd[m] = jv ? a_v[(size_t)j * Dd + c.vbase + m * c.vstride] : 0; }
    T ss = s[(size_t)j * H + q], se = s[(size_t)j * H + M + q], sx = L == 0 ? s[(size_t)j * H + 2 * M + q] : 0;
    if (!jv) th.fwd(g.psi, g.Y, ss, se, sx, vv);
    else th.jvp(g.psi, g.dpsi, g.Y, g.Ydot, g.ddot, ss, se, sx, vv, a_s[(size_t)j * H + q], a_s[(size_t)j * H + M + q],
                L == 0 ? a_s[(size_t)j * H + 2 * M + q] : 0, vd);
  }
  for (int m = 0; m < 2 * L + 1; ++m) outV[(size_t)i * Dd + c.vbase + m * c.vstride] += th.accV[m];
  if (L == 0) outx[(size_t)i * D.C + q] += th.accx;
}

extern "C" {

// x_out/V_out must be pre-initialised by the caller (x_in/V_in or zeros); result is accumulated.
// a_* == NULL: forward message.  a_* != NULL: forward-mode tangent along (a_s, a_v, a_pos).
void emul_center_pass(const int* dims_i, double rc, int N, const int* rowptr, const int* col, const int8_t* offs,
                      const double* cell, const double* pos, const double* s, const double* v, const double* W,
                      const double* b, const double* freq, const double* a_s, const double* a_v, const double* a_pos,
                      double* outx, double* outV) {
  Dims D{dims_i[0], dims_i[1], dims_i[2], dims_i[3], dims_i[4], rc};
  const int M = D.m0 + D.m1 + D.m2;
  for (int i = 0; i < N; ++i)
    for (int q = 0; q < M; ++q) {
      int l = q < D.m0 ? 0 : (q < D.m0 + D.m1 ? 1 : 2);
      if (l == 0) center_node<0>(D, i, q, rowptr, col, offs, cell, pos, s, v, W, b, freq, a_s, a_v, a_pos, outx, outV);
      else if (l == 1) center_node<1>(D, i, q, rowptr, col, offs, cell, pos, s, v, W, b, freq, a_s, a_v, a_pos, outx, outV);
      else center_node<2>(D, i, q, rowptr, col, offs, cell, pos, s, v, W, b, freq, a_s, a_v, a_pos, outx, outV);
    }
}
}

template <int L, int ROLE>
static void nbr_node(const Dims& D, int second, int j, int h, int q, const int* t_rowptr, const int* t_row, const int* t_eid,
                     const int8_t* offs, const T* cell, const T* pos, const T* s, const T* v, const T* W, const T* b,
                     const T* freq, const T* gx, const T* gV, const T* a_s, const T* a_v, const T* a_pos, T* o_s, T* o_v,
                     T* gr /*[E,3] accumulated*/, T* GW /*[H,NBP]*/, T* GF) {
  const int M = D.m0 + D.m1 + D.m2, Dd = D.m0 + 3 * D.m1 + 5 * D.m2, H = D.C + 2 * M, nk = D.B + 1;
  NeighborThread<T, L, ROLE, true, 21> th;
  load_wrow(W, b, h, D.B, th.Wt);
  th.reset_node(); th.reset_wgrad();
  Chan c = chan_of_q(D, ROLE == ROLE_SCALAR ? 0 : q);
  th.s = s[(size_t)j * H + h];
  th.sd = second && a_s ? a_s[(size_t)j * H + h] : 0;
  constexpr int NC = NeighborThread<T, L, ROLE, true, 21>::NC;
  if (ROLE == ROLE_STATE)
    for (int m = 0; m < NC; ++m) { th.v[m] = v[(size_t)j * Dd + c.vbase + m * c.vstride]; th.vd[m] = second && a_v ? a_v[(size_t)j * Dd + c.vbase + m * c.vstride] : 0; }
  NeighborThread<T, L, ROLE, false, 21> thm;
  thm.reset_node();
  thm.s = th.s; thm.sd = th.sd;
  for (int m = 0; m < NC; ++m) { thm.v[m] = ROLE == ROLE_STATE ? th.v[m] : T(0); thm.vd[m] = ROLE == ROLE_STATE ? th.vd[m] : T(0); }
  for (int sl = t_rowptr[j]; sl < t_rowptr[j + 1]; ++sl) {
    int i = t_row[sl], e = t_eid[sl];
    T r[3], rdot[3] = {0, 0, 0};
    edge_vector(pos, i, j, offs ? offs + 4 * e : nullptr, cell, r);
    if (second && a_pos) for (int x = 0; x < 3; ++x) rdot[x] = a_pos[3 * i + x] - a_pos[3 * j + x];
    Geo g; make_geo(r, second ? rdot : nullptr, freq, D, g);
    NbrEdge<T> ne{g.psi, g.dpsi, g.ddpsi, g.xi, g.dxi, g.Y, g.G, g.Hm, g.Ydot, g.u, g.rp, g.ddot};
    T gg[5];
    if (ROLE == ROLE_SCALAR) gg[0] = gx[(size_t)i * D.C + (h - 2 * M)];
    else for (int m = 0; m < NC; ++m) gg[m] = gV[(size_t)i * Dd + c.vbase + m * c.vstride];
    T pr[3];
    if (g_decomposition == 0) {
      if (second) th.second(ne, gg, pr); else th.first(ne, gg, pr);
    } else {
      NeighborThread<T, L, ROLE, false, 21>& tm = thm;  // no in-thread weight-gradient accumulators
      const T w = dot_nk<21>(th.Wt, g.psi), dw = dot_nk<21>(th.Wt, g.dpsi), ddw = dot_nk<21>(th.Wt, g.ddpsi);
      constexpr bool RADIAL = !(ROLE == ROLE_EDGE && L > 0);
      if constexpr (RADIAL) {
        T rad[2] = {0, 0};
        if (second) tm.template second_w<true>(ne, gg, nullptr, w, dw, ddw, rad);
        else tm.template first_w<true>(ne, gg, nullptr, w, dw, rad);
        for (int x = 0; x < 3; ++x) pr[x] = g.u[x] * rad[0] + (second ? g.rp[x] * rad[1] : T(0));
      } else {
        if (second) tm.second_w(ne, gg, pr, w, dw, ddw); else tm.first_w(ne, gg, pr, w, dw);
      }
      if (!second) {
        const T pw = tm.pw_first(g.Y, gg);
        for (int k = 0; k < NBP; ++k) { th.GW[k] += pw * g.psi[k]; th.GF[k] += pw * g.xi[k]; }
      } else {
        T al, be;
        tm.ab_second(g.Y, g.Ydot, gg, al, be);
        const T bd = be * g.ddot;
        for (int k = 0; k < NBP; ++k) { th.GW[k] += al * g.psi[k] + bd * g.dpsi[k]; th.GF[k] += al * g.xi[k] + bd * g.dxi[k]; }
      }
    }
    for (int x = 0; x < 3; ++x) gr[3 * (size_t)e + x] += pr[x];
  }
  if (g_decomposition != 0) { th.acc_s = thm.acc_s; for (int m = 0; m < NC; ++m) th.acc_v[m] = thm.acc_v[m]; }
  o_s[(size_t)j * H + h] = th.acc_s;
  if (ROLE == ROLE_STATE) for (int m = 0; m < NC; ++m) o_v[(size_t)j * Dd + c.vbase + m * c.vstride] = th.acc_v[m];
  for (int k = 0; k < NBP; ++k) { GW[h * NBP + k] += th.GW[k]; GF[h * NBP + k] += th.GF[k]; }
}

extern "C" {

// second == 0: K2b (first derivatives); second == 1: reverse half of K2bb.
// Outputs: o_s [N,H], o_v [N,D], o_pos [N,3], o_W [H,B], o_b [H], o_f [B] (all written).
void emul_neighbor_pass(const int* dims_i, double rc, int second, int N, int E, const int* rowptr, const int* t_rowptr,
                        const int* t_row, const int* t_eid, const int8_t* offs, const double* cell, const double* pos,
                        const double* s, const double* v, const double* W, const double* b, const double* freq,
                        const double* gx, const double* gV, const double* a_s, const double* a_v, const double* a_pos,
                        double* o_s, double* o_v, double* o_pos, double* o_W, double* o_b, double* o_f) {
  Dims D{dims_i[0], dims_i[1], dims_i[2], dims_i[3], dims_i[4], rc};
  const int M = D.m0 + D.m1 + D.m2, Dd = D.m0 + 3 * D.m1 + 5 * D.m2, H = D.C + 2 * M;
  std::vector<T> gr((size_t)E * 3, 0.0), GW((size_t)H * NBP, 0.0), GF((size_t)H * NBP, 0.0);
  std::memset(o_v, 0, sizeof(T) * (size_t)N * Dd);
  for (int j = 0; j < N; ++j)
    for (int h = 0; h < H; ++h) {
      int role = h < M ? ROLE_STATE : (h < 2 * M ? ROLE_EDGE : ROLE_SCALAR);
      int q = role == ROLE_STATE ? h : (role == ROLE_EDGE ? h - M : 0);
      int l = role == ROLE_SCALAR ? 0 : (q < D.m0 ? 0 : (q < D.m0 + D.m1 ? 1 : 2));
#define CALL(LL, RR) nbr_node<LL, RR>(D, second, j, h, q, t_rowptr, t_row, t_eid, offs, cell, pos, s, v, W, b, freq, gx, gV, a_s, a_v, a_pos, o_s, o_v, gr.data(), GW.data(), GF.data())
      if (role == ROLE_SCALAR) CALL(0, ROLE_SCALAR);
      else if (role == ROLE_STATE) { if (l == 0) CALL(0, ROLE_STATE); else if (l == 1) CALL(1, ROLE_STATE); else CALL(2, ROLE_STATE); }
      else { if (l == 0) CALL(0, ROLE_EDGE); else if (l == 1) CALL(1, ROLE_EDGE); else CALL(2, ROLE_EDGE); }
#undef CALL
    }
  // positions: + on the center, - on the neighbor (same reduction as the pos-grad kernel)
  for (int n = 0; n < N; ++n) {
    T acc[3] = {0, 0, 0};
    for (int e = rowptr[n]; e < rowptr[n + 1]; ++e) for (int x = 0; x < 3; ++x) acc[x] += gr[3 * (size_t)e + x];
    for (int sl = t_rowptr[n]; sl < t_rowptr[n + 1]; ++sl) for (int x = 0; x < 3; ++x) acc[x] -= gr[3 * (size_t)t_eid[sl] + x];
    for (int x = 0; x < 3; ++x) o_pos[3 * n + x] = acc[x];
  }
  for (int k = 0; k < D.B; ++k) o_f[k] = 0;
  for (int h = 0; h < H; ++h) {
    o_b[h] = GW[h * NBP];
    for (int k = 0; k < D.B; ++k) {
      o_W[h * D.B + k] = GW[h * NBP + k + 1];
      o_f[k] += W[h * D.B + k] * GF[h * NBP + k + 1];
    }
  }
}
}
