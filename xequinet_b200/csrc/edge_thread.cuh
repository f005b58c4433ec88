// Per-thread channel arithmetic of the fused edge kernels (K2 / K2b / K2bb).
//
// "Center" threads own one irrep channel q (all of its 2l+1 components, its state gate,
// its edge gate and -- for l = 0 -- the scalar-message channel) and walk a CSR row of the
// *receiving* node: forward message and the forward-mode tangent used by the double backward.
// "Neighbor" threads own one filter channel h < H and walk a transposed-CSR row of the
// *sending* node: first and second derivatives.
//
// Templated on the scalar type so tests/host_emul can run the same code in float64.
// Notation follows SURVEY.md Appendix A; see DESIGN.md for the second-order derivation.
#pragma once
#include "edge_math.cuh"

namespace xeq {

template <int L> struct YOff { static constexpr int value = (L == 1) ? 0 : 3; };  // offset into Y[8]

// NK-term dot product as three interleaved partial sums: the dependent-FMA chain is a third as
// long, which is what bounds a lightly occupied SM (few warps, 4-cycle FMA latency).
template <int NK, typename T>
XEQ_HD T dot_nk(const T* __restrict__ w, const T* __restrict__ p) {
  T a0 = T(0), a1 = T(0), a2 = T(0);
#pragma unroll
  for (int k = 0; k + 2 < NK; k += 3) {
    a0 += w[k] * p[k];
    a1 += w[k + 1] * p[k + 1];
    a2 += w[k + 2] * p[k + 2];
  }
#pragma unroll
  for (int k = NK - NK % 3; k < NK; ++k) a0 += w[k] * p[k];
  return (a0 + a1) + a2;
}

// ------------------------------------------------------------------------------------------
// Center threads (K2 forward, and the JVP half of K2bb)
// ------------------------------------------------------------------------------------------
template <typename T, int L, int NK>
struct CenterThread {
  static constexpr int NC = 2 * L + 1;
  T Ws[NBP], We[NBP], Wx[NBP];  // rows of [b | W_rbf] for the state gate, edge gate, scalar channel
  T accV[NC];
  T accx;

  XEQ_HD void reset() {
#pragma unroll
    for (int m = 0; m < NC; ++m) accV[m] = T(0);
    accx = T(0);
  }

  // forward message of one edge given the filter values w = [b|W_rbf] . psi of this thread's rows
  // (the tcgen05 kernels read them from TMEM); s_* are s[j, .] of the neighbor, v its v[j, (q, m)]
  XEQ_HD void fwd_w(T ws, T we, T wx, const T* Y, T s_state, T s_edge, T s_x, const T* v) {
    const T gs = s_state * ws;
    const T ge = s_edge * we;
    if (L == 0) {
      accV[0] += gs * v[0] + ge;
      accx += s_x * wx;
    } else {
#pragma unroll
      for (int m = 0; m < NC; ++m) accV[m] += gs * v[m] + ge * Y[YOff<L>::value + m];
    }
  }

  // forward message of one edge: psi[NBP], Y[8]
  XEQ_HD void fwd(const T* psi, const T* Y, T s_state, T s_edge, T s_x, const T* v) {
    fwd_w(dot_nk<NK>(Ws, psi), dot_nk<NK>(We, psi), (L == 0) ? dot_nk<NK>(Wx, psi) : T(0), Y, s_state, s_edge, s_x, v);
  }

  // tangent of the forward message along (sdot, vdot, rdot): d/deps of fwd_w(); dw* = [b|W_rbf] . dpsi
  XEQ_HD void jvp_w(T ws, T we, T wx, T dws_, T dwe_, T dwx_, const T* Y, const T* Ydot, T ddot, T s_state, T s_edge,
                    T s_x, const T* v, T sd_state, T sd_edge, T sd_x, const T* vd) {
    const T dws = dws_ * ddot, dwe = dwe_ * ddot;
    const T gs = s_state * ws, ge = s_edge * we;
    const T gsd = sd_state * ws + s_state * dws;
    const T ged = sd_edge * we + s_edge * dwe;
    if (L == 0) {
      accV[0] += gsd * v[0] + gs * vd[0] + ged;
      accx += sd_x * wx + s_x * dwx_ * ddot;
    } else {
#pragma unroll
      for (int m = 0; m < NC; ++m)
        accV[m] += gsd * v[m] + gs * vd[m] + ged * Y[YOff<L>::value + m] + ge * Ydot[YOff<L>::value + m];
    }
  }

  XEQ_HD void jvp(const T* psi, const T* dpsi, const T* Y, const T* Ydot, T ddot, T s_state, T s_edge,
                  T s_x, const T* v, T sd_state, T sd_edge, T sd_x, const T* vd) {
    jvp_w(dot_nk<NK>(Ws, psi), dot_nk<NK>(We, psi), (L == 0) ? dot_nk<NK>(Wx, psi) : T(0), dot_nk<NK>(Ws, dpsi),
          dot_nk<NK>(We, dpsi), (L == 0) ? dot_nk<NK>(Wx, dpsi) : T(0), Y, Ydot, ddot, s_state, s_edge, s_x, v, sd_state,
          sd_edge, sd_x, vd);
  }
};

// ------------------------------------------------------------------------------------------
// Neighbor threads (K2b, and the reverse half of K2bb)
// ------------------------------------------------------------------------------------------
enum Role : int { ROLE_STATE = 0, ROLE_EDGE = 1, ROLE_SCALAR = 2 };

// geometry of one edge as the neighbor threads see it (kernels keep this in shared memory)
template <typename T>
struct NbrEdge {
  const T* psi;    // [NBP]
  const T* dpsi;   // [NBP]
  const T* ddpsi;  // [NBP]  second order only
  const T* xi;     // [NBP]  wgrad only
  const T* dxi;    // [NBP]  second order + wgrad only
  const T* Y;      // [8]
  const T* G;      // [3*8]  G[x*8+m]
  const T* Hm;     // [3*8]  second order only
  const T* Ydot;   // [8]    second order only
  const T* u;      // [3]
  const T* rp;     // [3]    second order only
  T ddot;          //        second order only
};

// MAIN : produce d/ds, d/dv (registers) and this thread's share of d/dr.
// WGRAD: accumulate weight-gradient partials.  The kernels run MAIN and WGRAD as separate
// passes (thread-per-irrep vs thread-per-filter-channel); the host emulation runs both at once.
template <typename T, int L, int ROLE, bool WGRAD, int NK, bool MAIN = true>
struct NeighborThread {
  static constexpr int NC = (ROLE == ROLE_SCALAR) ? 1 : 2 * L + 1;
  static constexpr int YO = YOff<L>::value;
  T Wt[NBP];          // row h of [b | W_rbf]
  T s, sd;            // s[j,h] and its cotangent a_s[j,h] (second order)
  T v[NC], vd[NC];    // ROLE_STATE: v[j,(q,:)] and a_v[j,(q,:)]
  T acc_s;            // -> gs[j,h]        (first order) / o_s[j,h] (second order)
  T acc_v[NC];        // -> gv[j,(q,:)]    (ROLE_STATE)
  T GW[NBP], GF[NBP]; // wgrad accumulators (GW[0] is the bias gradient)

  XEQ_HD void reset_node() {
    acc_s = T(0);
#pragma unroll
    for (int m = 0; m < NC; ++m) acc_v[m] = T(0);
  }
  XEQ_HD void reset_wgrad() {
#pragma unroll
    for (int k = 0; k < NBP; ++k) GW[k] = GF[k] = T(0);
  }

  // First derivatives.  g = gV[i,(q,:)] (state/edge roles) or gx[i,c] (scalar role).
  // pr[3] receives this thread's share of dPhi/dr_e.
  XEQ_HD void first(const NbrEdge<T>& e, const T* g, T pr[3]) {
    first_w(e, g, pr, MAIN ? dot_nk<NK>(Wt, e.psi) : T(0), MAIN ? dot_nk<NK>(Wt, e.dpsi) : T(0));
  }
  // same with the filter values w = Wt . psi, dw = Wt . dpsi supplied (tcgen05 kernels: from TMEM)
  // RADIAL_ONLY (rows without an angular term: l = 0, or the state / scalar roles): return the coefficient
  // of u in dPhi/dr_e instead of the 3-vector, so that the cross-channel reduction moves one value.
  template <bool RADIAL_ONLY = false>
  XEQ_HD void first_w(const NbrEdge<T>& e, const T* g, T pr[3], const T w, const T dw, T* radial = nullptr) {
    static_assert(!RADIAL_ONLY || !(ROLE == ROLE_EDGE && L > 0), "row has an angular term");
    T pw;  // dPhi/dw_e[h]
    T cy[NC];
    if (ROLE == ROLE_STATE) {
      T A = T(0);
#pragma unroll
      for (int m = 0; m < NC; ++m) A += g[m] * v[m];
      pw = A * s;
      if (MAIN) {
        acc_s += A * w;
        const T sw = s * w;
#pragma unroll
        for (int m = 0; m < NC; ++m) acc_v[m] += sw * g[m];
      }
    } else if (ROLE == ROLE_EDGE) {
      T B = T(0);
      if (L == 0) {
        B = g[0];
      } else {
#pragma unroll
        for (int m = 0; m < NC; ++m) B += g[m] * e.Y[YO + m];
      }
      pw = B * s;
      if (MAIN) {
        acc_s += B * w;
        const T sw = s * w;
#pragma unroll
        for (int m = 0; m < NC; ++m) cy[m] = sw * g[m];
      }
    } else {
      pw = g[0] * s;
      if (MAIN) acc_s += g[0] * w;
    }
    if (MAIN && RADIAL_ONLY) {
      radial[0] = pw * dw;
    } else if (MAIN) {
      const T dpart = pw * dw;
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        T p = e.u[x] * dpart;
        if (ROLE == ROLE_EDGE && L > 0) {
#pragma unroll
          for (int m = 0; m < NC; ++m) p += e.G[x * 8 + YO + m] * cy[m];
        }
        pr[x] = p;
      }
    }
    if (WGRAD) {
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        GW[k] += pw * e.psi[k];
        GF[k] += pw * e.xi[k];
      }
    }
  }

  // dPhi/dw_e[h] of this thread's filter row (first order) -- the left operand of the weight-gradient
  // GEMM  GW[h, k] = sum_e pw[h, e] psi_k(e),  GF[h, k] = sum_e pw[h, e] xi_k(e)  (tcgen05 kernels)
  XEQ_HD T pw_first(const T* Y, const T* g) const {
    if (ROLE == ROLE_STATE) {
      T A = T(0);
#pragma unroll
      for (int m = 0; m < NC; ++m) A += g[m] * v[m];
      return A * s;
    } else if (ROLE == ROLE_EDGE) {
      T B = T(0);
      if (L == 0) {
        B = g[0];
      } else {
#pragma unroll
        for (int m = 0; m < NC; ++m) B += g[m] * Y[YO + m];
      }
      return B * s;
    }
    return g[0] * s;
  }
  // second order: GW += alpha psi + (beta ddot) dpsi,  GF += alpha xi + (beta ddot) dxi
  XEQ_HD void ab_second(const T* Y, const T* Ydot, const T* g, T& alpha, T& beta) const {
    if (ROLE == ROLE_STATE) {
      T A = T(0), Ad = T(0);
#pragma unroll
      for (int m = 0; m < NC; ++m) {
        A += g[m] * v[m];
        Ad += g[m] * vd[m];
      }
      alpha = sd * A + s * Ad;
      beta = s * A;
    } else if (ROLE == ROLE_EDGE) {
      T B = T(0), Bd = T(0);
      if (L == 0) {
        B = g[0];
      } else {
#pragma unroll
        for (int m = 0; m < NC; ++m) {
          B += g[m] * Y[YO + m];
          Bd += g[m] * Ydot[YO + m];
        }
      }
      alpha = sd * B + s * Bd;
      beta = s * B;
    } else {
      alpha = g[0] * sd;
      beta = g[0] * s;
    }
  }

  // Second derivatives: gradient of Psi_e (the tangent of Phi_e along (sd, vd, rdot)).
  XEQ_HD void second(const NbrEdge<T>& e, const T* g, T pr[3]) {
    second_w(e, g, pr, MAIN ? dot_nk<NK>(Wt, e.psi) : T(0), MAIN ? dot_nk<NK>(Wt, e.dpsi) : T(0),
             MAIN ? dot_nk<NK>(Wt, e.ddpsi) : T(0));
  }
  // RADIAL_ONLY: radial[0] = coefficient of u, radial[1] = coefficient of rp in dPsi/dr_e
  template <bool RADIAL_ONLY = false>
  XEQ_HD void second_w(const NbrEdge<T>& e, const T* g, T pr[3], const T w, const T dw, const T ddw, T* radial = nullptr) {
    static_assert(!RADIAL_ONLY || !(ROLE == ROLE_EDGE && L > 0), "row has an angular term");
    const T dwd = dw * e.ddot;  // tangent of w
    T alpha, beta;
    T cy[NC], cz[NC];
    if (ROLE == ROLE_STATE) {
      T A = T(0), Ad = T(0);
#pragma unroll
      for (int m = 0; m < NC; ++m) {
        A += g[m] * v[m];
        Ad += g[m] * vd[m];
      }
      alpha = sd * A + s * Ad;
      beta = s * A;
      if (MAIN) {
        acc_s += dwd * A + w * Ad;
        const T c = sd * w + s * dwd;
#pragma unroll
        for (int m = 0; m < NC; ++m) acc_v[m] += c * g[m];
      }
    } else if (ROLE == ROLE_EDGE) {
      T B = T(0), Bd = T(0);
      if (L == 0) {
        B = g[0];
      } else {
#pragma unroll
        for (int m = 0; m < NC; ++m) {
          B += g[m] * e.Y[YO + m];
          Bd += g[m] * e.Ydot[YO + m];
        }
      }
      alpha = sd * B + s * Bd;
      beta = s * B;
      if (MAIN) {
        acc_s += dwd * B + w * Bd;
        const T c = sd * w + s * dwd;
        const T sw = s * w;
#pragma unroll
        for (int m = 0; m < NC; ++m) {
          cy[m] = c * g[m];
          cz[m] = sw * g[m];
        }
      }
    } else {
      alpha = g[0] * sd;
      beta = g[0] * s;
      if (MAIN) acc_s += g[0] * dwd;
    }
    if (MAIN && RADIAL_ONLY) {
      radial[0] = alpha * dw + e.ddot * beta * ddw;
      radial[1] = beta * dw;
    } else if (MAIN) {
      const T P = alpha * dw + e.ddot * beta * ddw;
      const T R1 = beta * dw;
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        T p = e.u[x] * P + R1 * e.rp[x];
        if (ROLE == ROLE_EDGE && L > 0) {
#pragma unroll
          for (int m = 0; m < NC; ++m) p += e.G[x * 8 + YO + m] * cy[m] + e.Hm[x * 8 + YO + m] * cz[m];
        }
        pr[x] = p;
      }
    }
    if (WGRAD) {
      const T bd = beta * e.ddot;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        GW[k] += alpha * e.psi[k] + bd * e.dpsi[k];
        GF[k] += alpha * e.xi[k] + bd * e.dxi[k];
      }
    }
  }
};

}  // namespace xeq
