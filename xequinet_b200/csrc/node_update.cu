// Node-side per-irrep maps of XPainnUpdate (nn/xpainn.py:206-231) and the SiLU of the MLPs, each as
// one fused kernel with hand-written first and second derivatives (forces / force training).
// All on the cm layout [mul0 | 3 x mul1 | 5 x mul2]; one thread per (node, irrep q), which touches the
// 2l+1 components of q at stride mul_l (consecutive threads -> consecutive addresses: coalesced).
//
//  (1) invariant_dot   n[q]  = sqrt(sum_m W^2 + eps^2) - eps      Invariant      (nn/o3layer.py:40-44,  xpainn.py:214)
//                      t0[q] = sum_m U W                          EquivariantDot (nn/o3layer.py:104-109, xpainn.py:222)
//  (2) gate_residual   x' = x + a_sv * t + a_ss                   (nn/xpainn.py:218-229; a = [a_vv | a_sv | a_ss])
//                      V'[(q,m)] = V[(q,m)] + a_vv[q] U[(q,m)]    (ElementwiseTensorProduct with "Mx0e" = multiply)
//  (3) silu            y = u sigmoid(u)                           (nn/basic.py:255-256)
//
// Everything here is HBM-bound elementwise work: bytes = 4 x (elements read + written).
#include "common.cuh"

namespace xeq {
namespace {

struct IrrepShape {
  int m0, m1, m2, D, M;
};

struct QInfo {
  int l, base, mul, u;
};

__device__ __forceinline__ QInfo decode(int q, const IrrepShape& S) {
  QInfo r;
  if (q < S.m0) { r.l = 0; r.base = 0; r.mul = S.m0; r.u = q; }
  else if (q < S.m0 + S.m1) { r.l = 1; r.base = S.m0; r.mul = S.m1; r.u = q - S.m0; }
  else { r.l = 2; r.base = S.m0 + 3 * S.m1; r.mul = S.m2; r.u = q - S.m0 - S.m1; }
  return r;
}

constexpr float INV_EPS = 1e-5f;  // Invariant(eps=1e-5): sqrt(s + eps^2) - eps

// ---------------------------------------------------------------- (1) invariant + dot
__global__ void invdot_fwd_kernel(const float* __restrict__ U, const float* __restrict__ W, int n, IrrepShape S,
                                  float* __restrict__ nrm, int ld_nrm, float* __restrict__ t0) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * S.M) return;
  const int node = (int)(idx / S.M), q = (int)(idx % S.M);
  const QInfo I = decode(q, S);
  const float* u = U + (size_t)node * S.D + I.base + I.u;
  const float* w = W + (size_t)node * S.D + I.base + I.u;
  float s = 0.f, d = 0.f;
  for (int m = 0; m < 2 * I.l + 1; ++m) {
    const float wv = w[m * I.mul], uv = u[m * I.mul];
    s += wv * wv;
    d += uv * wv;
  }
  nrm[(size_t)node * ld_nrm + q] = sqrtf(s + INV_EPS * INV_EPS) - INV_EPS;
  t0[idx] = d;
}

__global__ void invdot_bwd_kernel(const float* __restrict__ U, const float* __restrict__ W, const float* __restrict__ gn,
                                  int ld_gn, const float* __restrict__ gt, const float* __restrict__ gU_add, int n, IrrepShape S,
                                  float* __restrict__ gU, float* __restrict__ gW) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * S.M) return;
  const int node = (int)(idx / S.M), q = (int)(idx % S.M);
  const QInfo I = decode(q, S);
  const size_t off = (size_t)node * S.D + I.base + I.u;
  float s = 0.f;
  for (int m = 0; m < 2 * I.l + 1; ++m) {
    const float wv = W[off + m * I.mul];
    s += wv * wv;
  }
  const float gnv = gn ? gn[(size_t)node * ld_gn + q] : 0.f, gtv = gt ? gt[idx] : 0.f;
  const float k = gnv / sqrtf(s + INV_EPS * INV_EPS);
  for (int m = 0; m < 2 * I.l + 1; ++m) {
    const float wv = W[off + m * I.mul], uv = U[off + m * I.mul];
    gU[off + m * I.mul] = gtv * wv + (gU_add ? gU_add[off + m * I.mul] : 0.f);  // + the gradient through U's other consumer
    gW[off + m * I.mul] = gtv * uv + k * wv;
  }
}

// cotangents aU, aW of (gU, gW)  ->  d/d(gn), d/d(gt), d/dU, d/dW
__global__ void invdot_bwdbwd_kernel(const float* __restrict__ U, const float* __restrict__ W, const float* __restrict__ gn,
                                     int ld_gn, const float* __restrict__ gt, const float* __restrict__ aU,
                                     const float* __restrict__ aW, int n, IrrepShape S, float* __restrict__ d_gn,
                                     float* __restrict__ d_gt, float* __restrict__ dU, float* __restrict__ dW) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * S.M) return;
  const int node = (int)(idx / S.M), q = (int)(idx % S.M);
  const QInfo I = decode(q, S);
  const size_t off = (size_t)node * S.D + I.base + I.u;
  float s = 0.f, p_uw = 0.f, p_ww = 0.f;  // sum W^2, sum (aU W + aW U), sum aW W
  for (int m = 0; m < 2 * I.l + 1; ++m) {
    const float wv = W[off + m * I.mul], uv = U[off + m * I.mul];
    const float au = aU ? aU[off + m * I.mul] : 0.f, aw = aW ? aW[off + m * I.mul] : 0.f;
    s += wv * wv;
    p_uw += au * wv + aw * uv;
    p_ww += aw * wv;
  }
  const float ir = 1.f / sqrtf(s + INV_EPS * INV_EPS);
  const float gnv = gn ? gn[(size_t)node * ld_gn + q] : 0.f, gtv = gt ? gt[idx] : 0.f;
  if (d_gn) d_gn[idx] = p_ww * ir;
  if (d_gt) d_gt[idx] = p_uw;
  const float k3 = gnv * p_ww * ir * ir * ir;
  for (int m = 0; m < 2 * I.l + 1; ++m) {
    const float wv = W[off + m * I.mul];
    const float au = aU ? aU[off + m * I.mul] : 0.f, aw = aW ? aW[off + m * I.mul] : 0.f;
    if (dU) dU[off + m * I.mul] = gtv * aw;
    if (dW) dW[off + m * I.mul] = gtv * au + gnv * aw * ir - k3 * wv;
  }
}

// ---------------------------------------------------------------- (2) gate + residual
// a rows: [a_vv (M) | a_sv (C) | a_ss (C)], C = m0
__global__ void gate_fwd_kernel(const float* __restrict__ a, const float* __restrict__ U, const float* __restrict__ t,
                                const float* __restrict__ x, const float* __restrict__ V, int n, IrrepShape S,
                                float* __restrict__ x_out, float* __restrict__ V_out) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * S.M) return;
  const int node = (int)(idx / S.M), q = (int)(idx % S.M);
  const QInfo I = decode(q, S);
  const int C = S.m0, HU = S.M + 2 * C;
  const float* ar = a + (size_t)node * HU;
  const size_t off = (size_t)node * S.D + I.base + I.u;
  const float avv = ar[q];
  for (int m = 0; m < 2 * I.l + 1; ++m) V_out[off + m * I.mul] = V[off + m * I.mul] + avv * U[off + m * I.mul];
  if (I.l == 0) {
    const size_t xo = (size_t)node * C + q;
    x_out[xo] = x[xo] + ar[S.M + q] * t[xo] + ar[S.M + C + q];
  }
}

// given gx' [N,C], gV' [N,D]:  ga [N,HU], gU [N,D], gt [N,C]   (d/dx = gx', d/dV = gV' are the identity)
__global__ void gate_bwd_kernel(const float* __restrict__ a, const float* __restrict__ U, const float* __restrict__ t,
                                const float* __restrict__ gx, const float* __restrict__ gV, int n, IrrepShape S,
                                float* __restrict__ ga, float* __restrict__ gU, float* __restrict__ gt) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * S.M) return;
  const int node = (int)(idx / S.M), q = (int)(idx % S.M);
  const QInfo I = decode(q, S);
  const int C = S.m0, HU = S.M + 2 * C;
  const float* ar = a + (size_t)node * HU;
  float* gar = ga + (size_t)node * HU;
  const size_t off = (size_t)node * S.D + I.base + I.u;
  const float avv = ar[q];
  float d = 0.f;
  for (int m = 0; m < 2 * I.l + 1; ++m) {
    const float gv = gV ? gV[off + m * I.mul] : 0.f;
    d += gv * U[off + m * I.mul];
    gU[off + m * I.mul] = gv * avv;
  }
  gar[q] = d;
  if (I.l == 0) {
    const size_t xo = (size_t)node * C + q;
    const float g = gx ? gx[xo] : 0.f;
    gar[S.M + q] = g * t[xo];
    gar[S.M + C + q] = g;
    gt[xo] = g * ar[S.M + q];
  }
}

// cotangents c_a [N,HU], c_U [N,D], c_t [N,C] of (ga, gU, gt)  ->  d/d(gx'), d/d(gV'), d/da, d/dU, d/dt
__global__ void gate_bwdbwd_kernel(const float* __restrict__ a, const float* __restrict__ U, const float* __restrict__ t,
                                   const float* __restrict__ gx, const float* __restrict__ gV, const float* __restrict__ c_a,
                                   const float* __restrict__ c_U, const float* __restrict__ c_t, int n, IrrepShape S,
                                   float* __restrict__ d_gx, float* __restrict__ d_gV, float* __restrict__ d_a,
                                   float* __restrict__ d_U, float* __restrict__ d_t) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * S.M) return;
  const int node = (int)(idx / S.M), q = (int)(idx % S.M);
  const QInfo I = decode(q, S);
  const int C = S.m0, HU = S.M + 2 * C;
  const float* ar = a + (size_t)node * HU;
  const float* car = c_a ? c_a + (size_t)node * HU : nullptr;
  float* dar = d_a + (size_t)node * HU;
  const size_t off = (size_t)node * S.D + I.base + I.u;
  const float avv = ar[q], cavv = car ? car[q] : 0.f;
  float d = 0.f;
  for (int m = 0; m < 2 * I.l + 1; ++m) {
    const float gv = gV ? gV[off + m * I.mul] : 0.f;
    const float cu = c_U ? c_U[off + m * I.mul] : 0.f;
    d += cu * gv;
    d_gV[off + m * I.mul] = cavv * U[off + m * I.mul] + cu * avv;
    d_U[off + m * I.mul] = cavv * gv;
  }
  dar[q] = d;
  if (I.l == 0) {
    const size_t xo = (size_t)node * C + q;
    const float g = gx ? gx[xo] : 0.f;
    const float casv = car ? car[S.M + q] : 0.f, cass = car ? car[S.M + C + q] : 0.f, ct = c_t ? c_t[xo] : 0.f;
    d_gx[xo] = cass + casv * t[xo] + ct * ar[S.M + q];
    d_t[xo] = casv * g;
    dar[S.M + q] = ct * g;
    dar[S.M + C + q] = 0.f;
  }
}

// ---------------------------------------------------------------- (3) SiLU
__device__ __forceinline__ float sigmoidf_(float u) { return 1.f / (1.f + expf(-u)); }  // accurate exp: __expf loses 2 + |1.17 u| ulp

__global__ void silu_fwd_kernel(const float* __restrict__ u, size_t n, float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = u[i] * sigmoidf_(u[i]);
}
__global__ void silu_bwd_kernel(const float* __restrict__ u, const float* __restrict__ g, size_t n, float* __restrict__ gu) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = u[i], s = sigmoidf_(x);
  gu[i] = g[i] * s * (1.f + x * (1.f - s));
}
// cotangent c of gu -> d/dg = c s'(u),  d/du = c g s''(u),  s'' = sig (1 - sig) (2 + u (1 - 2 sig))
__global__ void silu_bwdbwd_kernel(const float* __restrict__ u, const float* __restrict__ g, const float* __restrict__ c, size_t n,
                                   float* __restrict__ dg, float* __restrict__ du) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = u[i], s = sigmoidf_(x), cv = c[i];
  if (dg) dg[i] = cv * s * (1.f + x * (1.f - s));
  if (du) du[i] = cv * g[i] * s * (1.f - s) * (2.f + x * (1.f - 2.f * s));
}

int make_irreps(int32_t m0, int32_t m1, int32_t m2, IrrepShape* S) {
  XEQ_CHECK_ARG(m0 > 0 && m1 >= 0 && m2 >= 0, "irreps: bad multiplicities");
  S->m0 = m0; S->m1 = m1; S->m2 = m2;
  S->D = m0 + 3 * m1 + 5 * m2;
  S->M = m0 + m1 + m2;
  return XEQ_OK;
}

inline unsigned blocks_for(size_t total) { return (unsigned)((total + 255) / 256); }

}  // namespace
}  // namespace xeq

using namespace xeq;

extern "C" {

int xeq_invariant_dot_fwd(const float* U, const float* W, int32_t n, int32_t mul0, int32_t mul1, int32_t mul2,
                          float* nrm, int32_t ld_nrm, float* t0, xeq_stream_t stream) {
  IrrepShape S;
  int rc = make_irreps(mul0, mul1, mul2, &S);
  if (rc) return rc;
  XEQ_CHECK_ARG(n >= 0 && (n == 0 || (U && W && nrm && t0)) && ld_nrm >= S.M, "invariant_dot_fwd: bad arguments");
  if (n == 0) return XEQ_OK;
  XEQ_CUDA(launch_pdl(invdot_fwd_kernel, dim3(blocks_for((size_t)n * S.M)), dim3(256), (size_t)(0), (cudaStream_t)stream, U, W, n, S, nrm, ld_nrm, t0));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_invariant_dot_bwd(const float* U, const float* W, const float* gn, int32_t ld_gn, const float* gt, const float* gU_add,
                          int32_t n, int32_t mul0, int32_t mul1, int32_t mul2, float* gU, float* gW, xeq_stream_t stream) {
  IrrepShape S;
  int rc = make_irreps(mul0, mul1, mul2, &S);
  if (rc) return rc;
  XEQ_CHECK_ARG(n >= 0 && (n == 0 || (U && W && gU && gW)), "invariant_dot_bwd: bad arguments");
  if (n == 0) return XEQ_OK;
  XEQ_CUDA(launch_pdl(invdot_bwd_kernel, dim3(blocks_for((size_t)n * S.M)), dim3(256), (size_t)(0), (cudaStream_t)stream, U, W, gn, ld_gn, gt, gU_add, n, S, gU, gW));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_invariant_dot_bwdbwd(const float* U, const float* W, const float* gn, int32_t ld_gn, const float* gt,
                             const float* aU, const float* aW, int32_t n, int32_t mul0, int32_t mul1, int32_t mul2,
                             float* d_gn, float* d_gt, float* dU, float* dW, xeq_stream_t stream) {
  IrrepShape S;
  int rc = make_irreps(mul0, mul1, mul2, &S);
  if (rc) return rc;
  XEQ_CHECK_ARG(n >= 0 && (n == 0 || (U && W)), "invariant_dot_bwdbwd: bad arguments");
  if (n == 0) return XEQ_OK;
  XEQ_CUDA(launch_pdl(invdot_bwdbwd_kernel, dim3(blocks_for((size_t)n * S.M)), dim3(256), (size_t)(0), (cudaStream_t)stream, U, W, gn, ld_gn, gt, aU, aW, n, S, d_gn,
                                                                                     d_gt, dU, dW));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_gate_residual_fwd(const float* a, const float* U, const float* t, const float* x, const float* V, int32_t n,
                          int32_t mul0, int32_t mul1, int32_t mul2, float* x_out, float* V_out, xeq_stream_t stream) {
  IrrepShape S;
  int rc = make_irreps(mul0, mul1, mul2, &S);
  if (rc) return rc;
  XEQ_CHECK_ARG(n >= 0 && (n == 0 || (a && U && t && x && V && x_out && V_out)), "gate_residual_fwd: bad arguments");
  if (n == 0) return XEQ_OK;
  XEQ_CUDA(launch_pdl(gate_fwd_kernel, dim3(blocks_for((size_t)n * S.M)), dim3(256), (size_t)(0), (cudaStream_t)stream, a, U, t, x, V, n, S, x_out, V_out));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_gate_residual_bwd(const float* a, const float* U, const float* t, const float* gx, const float* gV, int32_t n,
                          int32_t mul0, int32_t mul1, int32_t mul2, float* ga, float* gU, float* gt, xeq_stream_t stream) {
  IrrepShape S;
  int rc = make_irreps(mul0, mul1, mul2, &S);
  if (rc) return rc;
  XEQ_CHECK_ARG(n >= 0 && (n == 0 || (a && U && t && ga && gU && gt)), "gate_residual_bwd: bad arguments");
  if (n == 0) return XEQ_OK;
  XEQ_CUDA(launch_pdl(gate_bwd_kernel, dim3(blocks_for((size_t)n * S.M)), dim3(256), (size_t)(0), (cudaStream_t)stream, a, U, t, gx, gV, n, S, ga, gU, gt));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_gate_residual_bwdbwd(const float* a, const float* U, const float* t, const float* gx, const float* gV,
                             const float* c_a, const float* c_U, const float* c_t, int32_t n, int32_t mul0, int32_t mul1,
                             int32_t mul2, float* d_gx, float* d_gV, float* d_a, float* d_U, float* d_t,
                             xeq_stream_t stream) {
  IrrepShape S;
  int rc = make_irreps(mul0, mul1, mul2, &S);
  if (rc) return rc;
  XEQ_CHECK_ARG(n >= 0 && (n == 0 || (a && U && t && d_gx && d_gV && d_a && d_U && d_t)), "gate_residual_bwdbwd: bad arguments");
  if (n == 0) return XEQ_OK;
  XEQ_CUDA(launch_pdl(gate_bwdbwd_kernel, dim3(blocks_for((size_t)n * S.M)), dim3(256), (size_t)(0), (cudaStream_t)stream, a, U, t, gx, gV, c_a, c_U, c_t, n, S, d_gx,
                                                                                   d_gV, d_a, d_U, d_t));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_silu_fwd(const float* u, size_t n, float* y, xeq_stream_t stream) {
  XEQ_CHECK_ARG(n == 0 || (u && y), "silu_fwd: bad arguments");
  if (n == 0) return XEQ_OK;
  XEQ_CUDA(launch_pdl(silu_fwd_kernel, dim3(blocks_for(n)), dim3(256), (size_t)(0), (cudaStream_t)stream, u, n, y));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}
int xeq_silu_bwd(const float* u, const float* g, size_t n, float* gu, xeq_stream_t stream) {
  XEQ_CHECK_ARG(n == 0 || (u && g && gu), "silu_bwd: bad arguments");
  if (n == 0) return XEQ_OK;
  XEQ_CUDA(launch_pdl(silu_bwd_kernel, dim3(blocks_for(n)), dim3(256), (size_t)(0), (cudaStream_t)stream, u, g, n, gu));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}
int xeq_silu_bwdbwd(const float* u, const float* g, const float* c, size_t n, float* dg, float* du, xeq_stream_t stream) {
  XEQ_CHECK_ARG(n == 0 || (u && g && c), "silu_bwdbwd: bad arguments");
  if (n == 0) return XEQ_OK;
  XEQ_CUDA(launch_pdl(silu_bwdbwd_kernel, dim3(blocks_for(n)), dim3(256), (size_t)(0), (cudaStream_t)stream, u, g, c, n, dg, du));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

}  // extern "C"
