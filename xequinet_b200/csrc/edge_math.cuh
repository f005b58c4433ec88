// Per-edge geometry of the XPaiNN message: radial basis x cutoff, real spherical harmonics
// up to l = 2, and the derivatives the backward / double-backward kernels need.
//
// Everything is templated on the scalar type and host/device-callable so that the very
// same code is exercised in float64 on the CPU by tests/host_emul (against torch autograd
// of the oracle) and in float32 inside the sm_100a kernels.
//
// Math follows SURVEY.md Appendix A:
//   phi_k(d) = sqrt(2/rc) sin(f_k d) / (d + 1e-5)            (nn/rbf.py:143-150)
//   chi(d)   = 0.5 (cos(pi d / rc) + 1)  for d < rc, else 0  (nn/rbf.py:43-57)
//   Y        = e3nn component-normalised real SH of normalize(r[[1,2,0]]) (nn/xpainn.py:71-74)
#pragma once

#if defined(__CUDACC__)
#define XEQ_HD __host__ __device__ __forceinline__
#else
#define XEQ_HD inline
#endif

#include <math.h>

namespace xeq {

constexpr int NBP = 24;  // padded number of radial terms (psi_0 .. psi_B), B <= 23

XEQ_HD float xsin(float x) { return sinf(x); }
XEQ_HD double xsin(double x) { return sin(x); }
XEQ_HD float xcos(float x) { return cosf(x); }
XEQ_HD double xcos(double x) { return cos(x); }
XEQ_HD void xsincos(float x, float& s, float& c) { sincosf(x, &s, &c); }  // one argument reduction for both
XEQ_HD void xsincos(double x, double& s, double& c) { s = sin(x); c = cos(x); }
XEQ_HD float xsqrt(float x) { return sqrtf(x); }
XEQ_HD double xsqrt(double x) { return sqrt(x); }

template <typename T>
struct Cutoff {
  T chi, dchi, ddchi;
};

// chi and its first two derivatives (zero at and beyond the cutoff, torch.where in rbf.py:48)
template <typename T>
XEQ_HD Cutoff<T> cutoff_terms(T d, T rc) {
  Cutoff<T> c;
  const T pi = T(3.14159265358979323846);
  if (d < rc) {
    const T a = pi / rc;
    T sn, cs;
    xsincos(a * d, sn, cs);
    c.chi = T(0.5) * (cs + T(1));
    c.dchi = -T(0.5) * a * sn;
    c.ddchi = -T(0.5) * a * a * cs;
  } else {
    c.chi = c.dchi = c.ddchi = T(0);
  }
  return c;
}

template <typename T>
struct Radial {
  T psi, dpsi, ddpsi;  // chi*phi and its d-derivatives
  T xi, dxi;           // d(psi)/df, d(dpsi)/df
};

// One Bessel term times the cutoff envelope, with everything the kernels differentiate.
// c0 = sqrt(2 / rc) is passed in: it is uniform, and recomputing the square root and the division per
// (edge, k) item was a measurable part of the radial stage.
template <typename T>
XEQ_HD Radial<T> radial_term_c0(T d, T f, T c0, const Cutoff<T>& c) {
  const T den = d + T(1e-5);
  const T inv = T(1) / den;
  T sn, cs;
  xsincos(f * d, sn, cs);
  const T phi = c0 * sn * inv;
  const T dphi = c0 * (f * cs * inv - sn * inv * inv);
  const T ddphi = c0 * (-f * f * sn * inv - T(2) * f * cs * inv * inv + T(2) * sn * inv * inv * inv);
  const T phif = c0 * d * cs * inv;                                        // d phi / d f
  const T dphif = c0 * (cs * inv - f * d * sn * inv - d * cs * inv * inv);  // d dphi / d f
  Radial<T> r;
  r.psi = c.chi * phi;
  r.dpsi = c.dchi * phi + c.chi * dphi;
  r.ddpsi = c.ddchi * phi + T(2) * c.dchi * dphi + c.chi * ddphi;
  r.xi = c.chi * phif;
  r.dxi = c.dchi * phif + c.chi * dphif;
  return r;
}
template <typename T>
XEQ_HD Radial<T> radial_term(T d, T f, T rc, const Cutoff<T>& c) {
  return radial_term_c0(d, f, xsqrt(T(2) / rc), c);
}

// r -> d, u = r / max(d, 1e-12) (F.normalize inside e3nn SH)
template <typename T>
XEQ_HD void unit_vector(const T r[3], T& d, T u[3]) {
  d = xsqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  const T dn = d > T(1e-12) ? d : T(1e-12);
  const T inv = T(1) / dn;
  u[0] = r[0] * inv;
  u[1] = r[1] * inv;
  u[2] = r[2] * inv;
}

#define XEQ_S3 1.7320508075688772935
#define XEQ_S5 2.2360679774997896964
#define XEQ_S15 3.8729833462074168852

// Y_1..Y_8 (Y_0 == 1 is implicit).  (a, b, c) = (u_y, u_z, u_x).
template <typename T>
XEQ_HD void sph_harm(const T u[3], T Y[8]) {
  const T a = u[1], b = u[2], c = u[0];
  Y[0] = T(XEQ_S3) * a;
  Y[1] = T(XEQ_S3) * b;
  Y[2] = T(XEQ_S3) * c;
  Y[3] = T(XEQ_S15) * a * c;
  Y[4] = T(XEQ_S15) * a * b;
  Y[5] = T(XEQ_S5) * (b * b - T(0.5) * (a * a + c * c));
  Y[6] = T(XEQ_S15) * b * c;
  Y[7] = T(0.5 * XEQ_S15) * (c * c - a * a);
}

// J[m][x] = dY_m / du_x  (x indexes u = (u_x, u_y, u_z) = (c, a, b)).  Rows 0..2 are constant,
// rows 3..7 linear in u, which is what makes the second derivative cheap:
// sph_jacobian_lin(udot) is dJ along udot.
template <typename T>
XEQ_HD void sph_jacobian(const T u[3], T J[8][3], bool linear_part_only = false) {
  const T a = u[1], b = u[2], c = u[0];
  const T k3 = linear_part_only ? T(0) : T(XEQ_S3);
  // columns: [0] = d/dc (u_x), [1] = d/da (u_y), [2] = d/db (u_z)
  J[0][0] = 0;  J[0][1] = k3; J[0][2] = 0;
  J[1][0] = 0;  J[1][1] = 0;  J[1][2] = k3;
  J[2][0] = k3; J[2][1] = 0;  J[2][2] = 0;
  J[3][0] = T(XEQ_S15) * a;  J[3][1] = T(XEQ_S15) * c;  J[3][2] = 0;
  J[4][0] = 0;               J[4][1] = T(XEQ_S15) * b;  J[4][2] = T(XEQ_S15) * a;
  J[5][0] = -T(XEQ_S5) * c;  J[5][1] = -T(XEQ_S5) * a;  J[5][2] = T(2 * XEQ_S5) * b;
  J[6][0] = T(XEQ_S15) * b;  J[6][1] = 0;               J[6][2] = T(XEQ_S15) * c;
  J[7][0] = T(XEQ_S15) * c;  J[7][1] = -T(XEQ_S15) * a; J[7][2] = 0;
}

// First-order angular record: Y and G[x][m] = dY_m/dr_x = ((I - u u^T)/d J^T)[x][m].
template <typename T>
XEQ_HD void angular_first(const T u[3], T d, T Y[8], T G[3][8]) {
  sph_harm(u, Y);
  T J[8][3];
  sph_jacobian(u, J);
  const T dn = d > T(1e-12) ? d : T(1e-12);
  const T inv = T(1) / dn;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const T ug = u[0] * J[m][0] + u[1] * J[m][1] + u[2] * J[m][2];
#pragma unroll
    for (int x = 0; x < 3; ++x) G[x][m] = (J[m][x] - u[x] * ug) * inv;
  }
}

// Second-order angular record for a tangent rdot of r:
//   ddot = u . rdot,  rp = (rdot - u ddot)/d = udot,  Ydot = G^T rdot,
//   Hm[x][m] = D_rdot G[x][m]  (Hessian of Y_m contracted with rdot).
template <typename T>
XEQ_HD void angular_second(const T u[3], T d, const T rdot[3], const T G[3][8], T& ddot, T rp[3], T Ydot[8],
                           T Hm[3][8]) {
  const T dn = d > T(1e-12) ? d : T(1e-12);
  const T inv = T(1) / dn;
  ddot = u[0] * rdot[0] + u[1] * rdot[1] + u[2] * rdot[2];
#pragma unroll
  for (int x = 0; x < 3; ++x) rp[x] = (rdot[x] - u[x] * ddot) * inv;
  T J[8][3], Jd[8][3];
  sph_jacobian(u, J);
  sph_jacobian(rp, Jd, true);  // dJ along udot = rp
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    Ydot[m] = G[0][m] * rdot[0] + G[1][m] * rdot[1] + G[2][m] * rdot[2];
    // gr = (gu - u (u.gu))/d with gu = J[m];  gr_dot = (gud - ud (u.gu) - u (ud.gu) - u (u.gud))/d - gr ddot/d
    const T ugu = u[0] * J[m][0] + u[1] * J[m][1] + u[2] * J[m][2];
    const T udgu = rp[0] * J[m][0] + rp[1] * J[m][1] + rp[2] * J[m][2];
    const T ugud = u[0] * Jd[m][0] + u[1] * Jd[m][1] + u[2] * Jd[m][2];
#pragma unroll
    for (int x = 0; x < 3; ++x)
      Hm[x][m] = (Jd[m][x] - rp[x] * ugu - u[x] * (udgu + ugud)) * inv - G[x][m] * ddot * inv;
  }
}

}  // namespace xeq
