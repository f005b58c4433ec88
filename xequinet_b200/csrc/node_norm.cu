// Node-side normalisations of the XPaiNN blocks as single fused kernels, with hand-written first
// and second derivatives (forces / force training):
//   EquivariantLayerNorm  (nn/o3layer.py:145-171; used at nn/xpainn.py:131, 209) on the cm layout
//   nn.LayerNorm(node_dim) (nn/xpainn.py:123, 130, 201, 208) = the same map with irreps "Cx0e"
//
// Both are a "centred RMS normalisation" of a row x[D], D = mul0 + 3 mul1 + 5 mul2:
//   z_i = x_i - [i < mul0] mean_{j < mul0} x_j,      rho = rsqrt( sum_i z_i^2 / M + eps ),  M = mul0 + mul1 + mul2
//   y_i = gamma_{q(i)} z_i rho + [i < mul0] beta_i   (q(i) = irrep of component i)
// (o3layer.py:150-169: scalars centred, per-irrep squared norms averaged over the M irreps, affine
// weight per irrep, bias on scalars; for "Cx0e" this is exactly LayerNorm with biased variance.)
//
// With h = gamma*g, A = h.z, c = P a (P = the centring projection), B = c.z, Cc = c.h:
//   bwd    : gz = rho h - (rho^3/M) A z ,  gx = P gz ,  ggamma_q = sum g z rho ,  gbeta = sum g
//   bwdbwd : Phi = <a, gx> = rho Cc - (rho^3/M) A B
//            dPhi/dg_i     = gamma (rho c_i - (rho^3/M) B z_i)
//            dPhi/dgamma_q = sum g_i (rho c_i - (rho^3/M) B z_i)
//            dPhi/dz_i     = -(rho^3/M)(Cc z_i + B h_i + A c_i) + 3 (rho^5/M^2) A B z_i ,  dPhi/dx = P dPhi/dz
//
// Mapping: one warp per row, lane l owns components l, l+32, ... (coalesced 128-byte accesses);
// the row reductions are warp butterflies; parameter gradients are accumulated per lane over the
// rows of a warp, combined per CTA in shared memory and written as per-CTA partial rows that a small
// second kernel sums in fixed order (deterministic, no atomics).  HBM-bound: 8 D bytes per row (fwd).
#include "common.cuh"

namespace xeq {
namespace {

constexpr int NORM_WARPS = 8;

struct NormShape {
  int m0, m1, m2, D, M;
  float eps;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int irrep_of(int i, const NormShape& S) {
  if (i < S.m0) return i;
  if (i < S.m0 + 3 * S.m1) return S.m0 + (i - S.m0) % S.m1;
  return S.m0 + S.m1 + (i - S.m0 - 3 * S.m1) % S.m2;
}

// per-row statistics shared by all three kernels
template <int NK>
__device__ __forceinline__ void center_and_scale(const float* __restrict__ xr, int lane, const NormShape& S, float (&z)[NK],
                                                 float& rho) {
  float s0 = 0.f;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int i = lane + 32 * k;
    z[k] = xr[i];
    if (i < S.m0) s0 += z[k];
  }
  const float mu = warp_sum(s0) / (float)S.m0;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int i = lane + 32 * k;
    if (i < S.m0) z[k] -= mu;
    ss += z[k] * z[k];
  }
  rho = 1.f / sqrtf(warp_sum(ss) / (float)S.M + S.eps);  // IEEE sqrt + divide: the approximate rsqrt (2 ulp) scales a whole row
}

template <int NK>
__device__ __forceinline__ void project(float (&v)[NK], int lane, const NormShape& S) {  // v <- P v
  float s0 = 0.f;
#pragma unroll
  for (int k = 0; k < NK; ++k)
    if (lane + 32 * k < S.m0) s0 += v[k];
  const float mu = warp_sum(s0) / (float)S.m0;
#pragma unroll
  for (int k = 0; k < NK; ++k)
    if (lane + 32 * k < S.m0) v[k] -= mu;
}

template <int NK>
__global__ void __launch_bounds__(NORM_WARPS * 32) norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, int n, NormShape S,
                                                                    float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gam[NK], bet[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int i = lane + 32 * k;
    gam[k] = gamma[irrep_of(i, S)];
    bet[k] = i < S.m0 ? beta[i] : 0.f;
  }
  for (int row = blockIdx.x * NORM_WARPS + warp; row < n; row += gridDim.x * NORM_WARPS) {
    float z[NK], rho;
    center_and_scale<NK>(x + (size_t)row * S.D, lane, S, z, rho);
#pragma unroll
    for (int k = 0; k < NK; ++k) y[(size_t)row * S.D + lane + 32 * k] = gam[k] * z[k] * rho + bet[k];
  }
}

// combine the per-lane parameter accumulators of the CTA's warps (fixed order) -> partial row
template <int NK, int NACC>
__device__ __forceinline__ void flush_partials(float (&acc)[NACC][NK], int lane, int warp, int D, float* __restrict__ partial_row) {
  __shared__ float red[NORM_WARPS][32 * NK + 1];
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NK; ++k) red[warp][lane + 32 * k] = acc[a][k];
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += NORM_WARPS * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < NORM_WARPS; ++w) s += red[w][i];
      partial_row[a * D + i] = s;
    }
  }
}

template <int NK>
__global__ void __launch_bounds__(NORM_WARPS * 32) norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                    const float* __restrict__ g, int ld_g,
                                                                    const float* __restrict__ gx_add, int n, NormShape S,
                                                                    float* __restrict__ gx,
                                                                    float* __restrict__ partials) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gam[NK], acc[2][NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    gam[k] = gamma[irrep_of(lane + 32 * k, S)];
    acc[0][k] = acc[1][k] = 0.f;
  }
  const float invM = 1.f / (float)S.M;
  for (int row = blockIdx.x * NORM_WARPS + warp; row < n; row += gridDim.x * NORM_WARPS) {
    float z[NK], rho, gr[NK], h[NK];
    center_and_scale<NK>(x + (size_t)row * S.D, lane, S, z, rho);
    float A = 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      gr[k] = g[(size_t)row * ld_g + lane + 32 * k];
      h[k] = gam[k] * gr[k];
      A += h[k] * z[k];
    }
    A = warp_sum(A);
    const float cA = rho * rho * rho * invM * A;
    float gz[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      gz[k] = rho * h[k] - cA * z[k];
      acc[0][k] += gr[k] * z[k] * rho;  // d/dgamma (per component, folded into irreps by the reduce kernel)
      acc[1][k] += gr[k];               // d/dbeta
    }
    project<NK>(gz, lane, S);
    if (gx) {  // gx_add: the gradient that reaches x through its other consumer (the residual), summed here
#pragma unroll
      for (int k = 0; k < NK; ++k)
        gx[(size_t)row * S.D + lane + 32 * k] = gz[k] + (gx_add ? gx_add[(size_t)row * S.D + lane + 32 * k] : 0.f);
    }
  }
  if (partials) flush_partials<NK, 2>(acc, lane, warp, S.D, partials + (size_t)blockIdx.x * 2 * S.D);
}

template <int NK>
__global__ void __launch_bounds__(NORM_WARPS * 32) norm_bwdbwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                       const float* __restrict__ g, int ld_g,
                                                                       const float* __restrict__ a, int n, NormShape S,
                                                                       float* __restrict__ dx,
                                                                       float* __restrict__ dg, float* __restrict__ partials) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gam[NK], acc[1][NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    gam[k] = gamma[irrep_of(lane + 32 * k, S)];
    acc[0][k] = 0.f;
  }
  const float invM = 1.f / (float)S.M;
  for (int row = blockIdx.x * NORM_WARPS + warp; row < n; row += gridDim.x * NORM_WARPS) {
    float z[NK], rho, gr[NK], h[NK], c[NK];
    center_and_scale<NK>(x + (size_t)row * S.D, lane, S, z, rho);
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      gr[k] = g[(size_t)row * ld_g + lane + 32 * k];
      h[k] = gam[k] * gr[k];
      c[k] = a[(size_t)row * S.D + lane + 32 * k];
    }
    project<NK>(c, lane, S);
    float A = 0.f, B = 0.f, Cc = 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      A += h[k] * z[k];
      B += c[k] * z[k];
      Cc += c[k] * h[k];
    }
    A = warp_sum(A);
    B = warp_sum(B);
    Cc = warp_sum(Cc);
    const float r3 = rho * rho * rho * invM;            // rho^3 / M
    const float r5 = 3.f * r3 * rho * rho * invM * A * B;  // 3 rho^5 A B / M^2
    float dz[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const float t = rho * c[k] - r3 * B * z[k];  // d/dh
      if (dg) dg[(size_t)row * S.D + lane + 32 * k] = gam[k] * t;
      acc[0][k] += gr[k] * t;
      dz[k] = -r3 * (Cc * z[k] + B * h[k] + A * c[k]) + r5 * z[k];
    }
    project<NK>(dz, lane, S);
    if (dx) {
#pragma unroll
      for (int k = 0; k < NK; ++k) dx[(size_t)row * S.D + lane + 32 * k] = dz[k];
    }
  }
  if (partials) flush_partials<NK, 1>(acc, lane, warp, S.D, partials + (size_t)blockIdx.x * S.D);
}

// partial rows [n_part][n_acc * D] -> out_gamma[M] (components folded into irreps), out_beta[m0].
// One CTA per 32 consecutive irreps of one l (multiplicities are multiples of 32): 32 columns x 32 row
// lanes, every load is a coalesced 128-byte row segment, rows are summed in a fixed order.
constexpr int PR_ROWS = 32;
__global__ void __launch_bounds__(32 * PR_ROWS) norm_param_reduce_kernel(const float* __restrict__ partials, int n_part, int n_acc,
                                                                         NormShape S, float* __restrict__ out_gamma,
                                                                         float* __restrict__ out_beta) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[2][PR_ROWS][33];
  const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int q = blockIdx.x * 32 + c;
  int l, u, base, mul;
  if (q < S.m0) { l = 0; u = q; base = 0; mul = S.m0; }
  else if (q < S.m0 + S.m1) { l = 1; u = q - S.m0; base = S.m0; mul = S.m1; }
  else { l = 2; u = q - S.m0 - S.m1; base = S.m0 + 3 * S.m1; mul = S.m2; }
  const size_t stride = (size_t)n_acc * S.D;
  float sg = 0.f, sb = 0.f;
  for (int m = 0; m < 2 * l + 1; ++m) {
    const float* col = partials + base + m * mul + u;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int p = rl;
    for (; p + 3 * PR_ROWS < n_part; p += 4 * PR_ROWS) {
      s0 += col[(size_t)p * stride];
      s1 += col[(size_t)(p + PR_ROWS) * stride];
      s2 += col[(size_t)(p + 2 * PR_ROWS) * stride];
      s3 += col[(size_t)(p + 3 * PR_ROWS) * stride];
    }
    for (; p < n_part; p += PR_ROWS) s0 += col[(size_t)p * stride];
    sg += (s0 + s1) + (s2 + s3);
  }
  if (n_acc > 1 && l == 0) {
    const float* col = partials + S.D + u;
    for (int p = rl; p < n_part; p += PR_ROWS) sb += col[(size_t)p * stride];
  }
  red[0][rl][c] = sg;
  red[1][rl][c] = sb;
  __syncthreads();
  if (rl == 0) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int r = 0; r < PR_ROWS; ++r) {
      tg += red[0][r][c];
      tb += red[1][r][c];
    }
    if (out_gamma) out_gamma[q] = tg;
    if (out_beta && l == 0) out_beta[q] = tb;
  }
}

int make_shape(int32_t m0, int32_t m1, int32_t m2, float eps, NormShape* S) {
  XEQ_CHECK_ARG(m0 > 0 && m1 >= 0 && m2 >= 0 && m0 % 32 == 0 && m1 % 32 == 0 && m2 % 32 == 0,
                "norm: multiplicities must be multiples of 32 (mul0 > 0)");
  S->m0 = m0; S->m1 = m1; S->m2 = m2;
  S->D = m0 + 3 * m1 + 5 * m2;
  S->M = m0 + m1 + m2;
  S->eps = eps;
  const int nk = S->D / 32;
  XEQ_CHECK_ARG(nk == 1 || nk == 2 || nk == 4 || nk == 8 || nk == 9 || nk == 15 || nk == 30,
                "norm: unsupported row width %d (supported: 32, 64, 128, 256, 288, 480, 960)", S->D);
  return XEQ_OK;
}

int norm_grid(int n) { return max(1, min((n + NORM_WARPS - 1) / NORM_WARPS, num_sms() * 2)); }

#define NORM_DISPATCH(NKV, KERNEL, ...)                                      \
  switch (NKV) {                                                            \
    case 1: XEQ_CUDA(launch_pdl(KERNEL<1>, dim3(grid), dim3(NORM_WARPS * 32), (size_t)(0), st, __VA_ARGS__)); break;  \
    case 2: XEQ_CUDA(launch_pdl(KERNEL<2>, dim3(grid), dim3(NORM_WARPS * 32), (size_t)(0), st, __VA_ARGS__)); break;  \
    case 4: XEQ_CUDA(launch_pdl(KERNEL<4>, dim3(grid), dim3(NORM_WARPS * 32), (size_t)(0), st, __VA_ARGS__)); break;  \
    case 8: XEQ_CUDA(launch_pdl(KERNEL<8>, dim3(grid), dim3(NORM_WARPS * 32), (size_t)(0), st, __VA_ARGS__)); break;  \
    case 9: XEQ_CUDA(launch_pdl(KERNEL<9>, dim3(grid), dim3(NORM_WARPS * 32), (size_t)(0), st, __VA_ARGS__)); break;  \
    case 15: XEQ_CUDA(launch_pdl(KERNEL<15>, dim3(grid), dim3(NORM_WARPS * 32), (size_t)(0), st, __VA_ARGS__)); break; \
    default: XEQ_CUDA(launch_pdl(KERNEL<30>, dim3(grid), dim3(NORM_WARPS * 32), (size_t)(0), st, __VA_ARGS__)); break; \
  }

}  // namespace
}  // namespace xeq

using namespace xeq;

extern "C" {

size_t xeq_irreps_norm_workspace_bytes(int32_t n_rows, int32_t mul0, int32_t mul1, int32_t mul2) {
  const size_t D = (size_t)mul0 + 3 * (size_t)mul1 + 5 * (size_t)mul2;
  return (size_t)norm_grid(n_rows) * 2 * D * sizeof(float);
}

int xeq_irreps_norm_fwd(const float* x, const float* gamma, const float* beta, int32_t n_rows, int32_t mul0, int32_t mul1,
                        int32_t mul2, float eps, float* y, xeq_stream_t stream) {
  NormShape S;
  int rc = make_shape(mul0, mul1, mul2, eps, &S);
  if (rc) return rc;
  XEQ_CHECK_ARG(n_rows >= 0 && (n_rows == 0 || (x && gamma && beta && y)), "irreps_norm_fwd: bad arguments");
  if (n_rows == 0) return XEQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = norm_grid(n_rows);
  NORM_DISPATCH(S.D / 32, norm_fwd_kernel, x, gamma, beta, n_rows, S, y);
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_irreps_norm_bwd(const float* x, const float* gamma, const float* g, int32_t ld_g, const float* gx_add, int32_t n_rows, int32_t mul0,
                        int32_t mul1, int32_t mul2, float eps, float* gx, float* ggamma, float* gbeta, void* workspace,
                        size_t workspace_bytes, xeq_stream_t stream) {
  NormShape S;
  int rc = make_shape(mul0, mul1, mul2, eps, &S);
  if (rc) return rc;
  XEQ_CHECK_ARG(n_rows >= 0 && (n_rows == 0 || (x && gamma && g)), "irreps_norm_bwd: bad arguments");
  const bool params = ggamma || gbeta;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = norm_grid(n_rows);
  if (params)
    XEQ_CHECK_ARG(workspace && workspace_bytes >= xeq_irreps_norm_workspace_bytes(n_rows, mul0, mul1, mul2),
                  "irreps_norm_bwd: workspace too small");
  float* partials = params ? static_cast<float*>(workspace) : nullptr;
  if (n_rows > 0) {
    NORM_DISPATCH(S.D / 32, norm_bwd_kernel, x, gamma, g, ld_g > 0 ? ld_g : S.D, gx_add, n_rows, S, gx, partials);
    XEQ_LAUNCHED(1);
  }
  if (params) {
    XEQ_CUDA(launch_pdl(norm_param_reduce_kernel, dim3(S.M / 32), dim3(32 * PR_ROWS), (size_t)(0), st, partials, n_rows > 0 ? grid : 0, 2, S, ggamma, gbeta));
    XEQ_LAUNCHED(1);
  }
  return XEQ_OK;
}

int xeq_irreps_norm_bwdbwd(const float* x, const float* gamma, const float* g, int32_t ld_g, const float* a, int32_t n_rows, int32_t mul0,
                           int32_t mul1, int32_t mul2, float eps, float* dx, float* dg, float* dgamma, void* workspace,
                           size_t workspace_bytes, xeq_stream_t stream) {
  NormShape S;
  int rc = make_shape(mul0, mul1, mul2, eps, &S);
  if (rc) return rc;
  XEQ_CHECK_ARG(n_rows >= 0 && (n_rows == 0 || (x && gamma && g && a)), "irreps_norm_bwdbwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = norm_grid(n_rows);
  if (dgamma)
    XEQ_CHECK_ARG(workspace && workspace_bytes >= xeq_irreps_norm_workspace_bytes(n_rows, mul0, mul1, mul2),
                  "irreps_norm_bwdbwd: workspace too small");
  float* partials = dgamma ? static_cast<float*>(workspace) : nullptr;
  if (n_rows > 0) {
    NORM_DISPATCH(S.D / 32, norm_bwdbwd_kernel, x, gamma, g, ld_g > 0 ? ld_g : S.D, a, n_rows, S, dx, dg, partials);
    XEQ_LAUNCHED(1);
  }
  if (dgamma) {
    XEQ_CUDA(launch_pdl(norm_param_reduce_kernel, dim3(S.M / 32), dim3(32 * PR_ROWS), (size_t)(0), st, partials, n_rows > 0 ? grid : 0, 1, S, dgamma, nullptr));
    XEQ_LAUNCHED(1);
  }
  return XEQ_OK;
}

}  // extern "C"
