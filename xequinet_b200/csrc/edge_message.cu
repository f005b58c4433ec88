// K2 / K2b / K2bb: the fused XPaiNN edge message and its first and second derivatives.
// Replaces nn/xpainn.py:66-74,140-159 + nn/basic.py:114-131 and their autograd replays
// (nn/basic.py:143-159, utils/trainer.py:302).  Contract: include/xeq_b200.h.
//
// Two kernel families, both persistent over node-aligned edge tiles, no global atomics:
//   center_kernel   : CTA walks CSR rows of receiving nodes, one thread per irrep channel q
//                     (state gate, edge gate, 2l+1 components, + scalar channel for l = 0);
//                     the segment sum lives in registers.  Forward message, or (JVP) its
//                     tangent along (a_s, a_v, a_pos) = d/d(gx, gV) half of the double backward.
//   neighbor_kernel : CTA walks transposed-CSR rows of sending nodes, one thread per filter
//                     channel h; produces d/ds, d/dv (registers), per-edge d/dr (warp shuffle
//                     + shared memory) and weight-gradient partials (registers).  ORDER 1 = K2b,
//                     ORDER 2 = reverse half of K2bb.
// Per-edge geometry (r, d, Y, chi * phi_k and derivatives) is recomputed per chunk into shared
// memory; nothing E-sized except the 12-byte d/dr record ever touches HBM.
#include "common.cuh"
#include "edge_thread.cuh"

namespace xeq {

constexpr int NB_ = 20;       // num_basis instantiated
constexpr int NK_ = NB_ + 1;  // + bias/cutoff term

__device__ __forceinline__ int lower_bound_nodes(const int* __restrict__ rowptr, int n_nodes, int x) {
  int lo = 0, hi = n_nodes;  // first node with rowptr[node] >= x, n_nodes if none
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rowptr[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// node that owns edge/slot e inside [n0, n1): rowptr[i] <= e < rowptr[i+1]
__device__ __forceinline__ int owner_of(const int* __restrict__ rowptr, int n0, int n1, int e) {
  int lo = n0, hi = n1;  // first node with rowptr[node] > e, minus one
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rowptr[mid] <= e) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

__device__ __forceinline__ void edge_vector(const xeq_graph_t& g, const float* __restrict__ pos, int i, int j, int eid,
                                            float r[3]) {
  r[0] = pos[3 * i] - pos[3 * j];
  r[1] = pos[3 * i + 1] - pos[3 * j + 1];
  r[2] = pos[3 * i + 2] - pos[3 * j + 2];
  if (g.offsets != nullptr) {  // nn/basic.py:119-128: vectors -= cell_offsets @ cell[graph(neighbor)]
    const char4 o = reinterpret_cast<const char4*>(g.offsets)[eid];
    const float* c = g.cell + 9 * (g.node_graph ? g.node_graph[j] : 0);
    const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
#pragma unroll
    for (int x = 0; x < 3; ++x) r[x] -= ox * c[x] + oy * c[3 + x] + oz * c[6 + x];
  }
}

template <int NKK>
__device__ __forceinline__ void load_wrow(const float* __restrict__ W, const float* __restrict__ b, int h, float* row) {
  row[0] = b[h];
#pragma unroll
  for (int k = 0; k < NB_; ++k) row[k + 1] = W[(size_t)h * NB_ + k];
#pragma unroll
  for (int k = NKK; k < NBP; ++k) row[k] = 0.f;
}

__device__ __forceinline__ void lds_row(const float* __restrict__ src, float* dst) {
#pragma unroll
  for (int k = 0; k < NBP / 4; ++k) {
    const float4 p = reinterpret_cast<const float4*>(src)[k];
    dst[4 * k] = p.x; dst[4 * k + 1] = p.y; dst[4 * k + 2] = p.z; dst[4 * k + 3] = p.w;
  }
}

// ==========================================================================================
// center kernel
// ==========================================================================================
constexpr int CT = 64;  // edges per chunk

template <bool JVP>
struct CenterSmem {
  float psi[CT][NBP];
  float dpsi[JVP ? CT : 1][NBP];
  float Y[CT][8];
  float Ydot[JVP ? CT : 1][8];
  float ddot[CT];
  float d[CT];
  float chi[CT][2];
  int nbr[CT];
  int ctr[CT];
};

struct CenterArgs {
  xeq_graph_t g;
  float rc;
  const float *pos, *s, *v, *x_in, *V_in, *W, *b, *freq;
  const float *a_s, *a_v, *a_pos;  // JVP only
  float *x_out, *V_out;
  int n_tiles;
};

template <int L, int C, int M1, int M2, bool JVP>
__device__ __forceinline__ void center_role(const CenterArgs& A, CenterSmem<JVP>& sm) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M, NC = 2 * L + 1;
  constexpr int THREADS = M;
  const int t = threadIdx.x;
  const int q = t;  // thread index == irrep channel (types are contiguous ranges)
  const int vbase = (L == 0) ? t : (L == 1 ? C + (t - C) : C + 3 * M1 + (t - C - M1));
  constexpr int vstride = (L == 0) ? 0 : (L == 1 ? M1 : M2);
  const xeq_graph_t& g = A.g;
  const int N = g.n_nodes;

  CenterThread<float, L, NK_> th;
  load_wrow<NK_>(A.W, A.b, q, th.Ws);
  load_wrow<NK_>(A.W, A.b, M + q, th.We);
  if (L == 0) load_wrow<NK_>(A.W, A.b, 2 * M + q, th.Wx);

  auto emit = [&](int node, bool with_acc) {
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      const size_t idx = (size_t)node * D + vbase + m * vstride;
      const float base = A.V_in ? A.V_in[idx] : 0.f;
      A.V_out[idx] = with_acc ? base + th.accV[m] : base;
    }
    if (L == 0) {
      const size_t idx = (size_t)node * C + t;
      const float base = A.x_in ? A.x_in[idx] : 0.f;
      A.x_out[idx] = with_acc ? base + th.accx : base;
    }
  };

  for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
    const int n0 = lower_bound_nodes(g.rowptr, N, tile * CT);
    const int n1 = lower_bound_nodes(g.rowptr, N, (tile + 1) * CT);
    if (n0 == n1) continue;
    const int e0 = g.rowptr[n0], e1 = g.rowptr[n1];
    int cur = -1;
    for (int eb = e0; eb < e1; eb += CT) {
      const int cnt = min(CT, e1 - eb);
      __syncthreads();
      if (t < cnt) {  // ---- geometry, one thread per edge
        const int e = eb + t;
        const int i = owner_of(g.rowptr, n0, n1, e);
        const int j = g.col[e];
        float r[3], d, u[3];
        edge_vector(g, A.pos, i, j, e, r);
        unit_vector(r, d, u);
        if (!JVP) {
          sph_harm(u, sm.Y[t]);
        } else {
          float G[3][8], Hm[3][8], rp[3], rdot[3] = {0.f, 0.f, 0.f}, dd;
          angular_first(u, d, sm.Y[t], G);
          if (A.a_pos) {
#pragma unroll
            for (int x = 0; x < 3; ++x) rdot[x] = A.a_pos[3 * i + x] - A.a_pos[3 * j + x];
          }
          angular_second(u, d, rdot, G, dd, rp, sm.Ydot[t], Hm);
          sm.ddot[t] = dd;
        }
        const Cutoff<float> c = cutoff_terms(d, A.rc);
        sm.d[t] = d;
        sm.chi[t][0] = c.chi;
        sm.chi[t][1] = c.dchi;
        sm.psi[t][0] = c.chi;
#pragma unroll
        for (int k = NK_; k < NBP; ++k) sm.psi[t][k] = 0.f;
        if (JVP) {
          sm.dpsi[t][0] = c.dchi;
#pragma unroll
          for (int k = NK_; k < NBP; ++k) sm.dpsi[t][k] = 0.f;
        }
        sm.nbr[t] = j;
        sm.ctr[t] = i;
      }
      __syncthreads();
      for (int idx = t; idx < cnt * NB_; idx += THREADS) {  // ---- radial terms, one thread per (edge, k)
        const int ee = idx / NB_, k = idx - ee * NB_;
        Cutoff<float> c;
        c.chi = sm.chi[ee][0];
        c.dchi = sm.chi[ee][1];
        c.ddchi = 0.f;
        const Radial<float> rr = radial_term(sm.d[ee], A.freq[k], A.rc, c);
        sm.psi[ee][k + 1] = rr.psi;
        if (JVP) sm.dpsi[ee][k + 1] = rr.dpsi;
      }
      __syncthreads();

      // ---- message accumulation: software-pipelined gathers of the neighbor rows
      float ss, se, sx = 0.f, vv[NC], sds = 0.f, sde = 0.f, sdx = 0.f, vd[NC];
      auto gather = [&](int ee, float& a_ss, float& a_se, float& a_sx, float* a_v, float& a_sds, float& a_sde,
                        float& a_sdx, float* a_vd) {
        const int j = sm.nbr[ee];
        const float* sj = A.s + (size_t)j * H;
        a_ss = sj[q];
        a_se = sj[M + q];
        if (L == 0) a_sx = sj[2 * M + q];
        const float* vj = A.v + (size_t)j * D + vbase;
#pragma unroll
        for (int m = 0; m < NC; ++m) a_v[m] = vj[m * vstride];
        if (JVP) {
          if (A.a_s) {
            const float* aj = A.a_s + (size_t)j * H;
            a_sds = aj[q];
            a_sde = aj[M + q];
            if (L == 0) a_sdx = aj[2 * M + q];
          }
#pragma unroll
          for (int m = 0; m < NC; ++m) a_vd[m] = A.a_v ? A.a_v[(size_t)j * D + vbase + m * vstride] : 0.f;
        }
      };
      gather(0, ss, se, sx, vv, sds, sde, sdx, vd);
      for (int ee = 0; ee < cnt; ++ee) {
        float nss = 0.f, nse = 0.f, nsx = 0.f, nv[NC], nsds = 0.f, nsde = 0.f, nsdx = 0.f, nvd[NC];
#pragma unroll
        for (int m = 0; m < NC; ++m) nv[m] = nvd[m] = 0.f;
        if (ee + 1 < cnt) gather(ee + 1, nss, nse, nsx, nv, nsds, nsde, nsdx, nvd);
        const int i = sm.ctr[ee];
        if (i != cur) {
          if (cur >= 0) emit(cur, true);
          for (int nn = (cur >= 0 ? cur + 1 : n0); nn < i; ++nn) emit(nn, false);
          cur = i;
          th.reset();
        }
        float p[NBP];
        lds_row(sm.psi[ee], p);
        if (!JVP) {
          th.fwd(p, sm.Y[ee], ss, se, sx, vv);
        } else {
          float dp[NBP];
          lds_row(sm.dpsi[ee], dp);
          th.jvp(p, dp, sm.Y[ee], sm.Ydot[ee], sm.ddot[ee], ss, se, sx, vv, sds, sde, sdx, vd);
        }
        ss = nss; se = nse; sx = nsx; sds = nsds; sde = nsde; sdx = nsdx;
#pragma unroll
        for (int m = 0; m < NC; ++m) { vv[m] = nv[m]; vd[m] = nvd[m]; }
      }
    }
    if (cur >= 0) emit(cur, true);
    for (int nn = (cur >= 0 ? cur + 1 : n0); nn < n1; ++nn) emit(nn, false);
  }
}

template <int C, int M1, int M2, bool JVP>
__global__ void __launch_bounds__(C + M1 + M2) center_kernel(const CenterArgs A) {
  __shared__ __align__(16) CenterSmem<JVP> sm;
  const int t = threadIdx.x;
  if (t < C) center_role<0, C, M1, M2, JVP>(A, sm);
  else if (t < C + M1) center_role<1, C, M1, M2, JVP>(A, sm);
  else center_role<2, C, M1, M2, JVP>(A, sm);
}

// ==========================================================================================
// neighbor kernel
// ==========================================================================================
constexpr int NT = 32;        // slots per chunk
constexpr int NTHREADS = 288; // filter channels per CTA (9 warps); grid.y = H / 288 channel slices
constexpr int NWARPS = NTHREADS / 32;

template <int ORDER, bool WGRAD>
struct NeighborSmem {
  float psi[NT][NBP];
  float dpsi[NT][NBP];
  float ddpsi[ORDER == 2 ? NT : 1][NBP];
  float xi[WGRAD ? NT : 1][NBP];
  float dxi[(WGRAD && ORDER == 2) ? NT : 1][NBP];
  float Y[NT][8];
  float G[NT][24];
  float Hm[ORDER == 2 ? NT : 1][24];
  float Ydot[ORDER == 2 ? NT : 1][8];
  float u[NT][4];
  float rp[ORDER == 2 ? NT : 1][4];
  float ddot[NT];
  float d[NT];
  float chi[NT][3];
  float red[NT][NWARPS][3];
  int gat[NT];  // center i whose gx/gV row is gathered
  int own[NT];  // sending node j (owner of the transposed row)
  int eid[NT];  // canonical edge id
};

struct NeighborArgs {
  xeq_graph_t g;
  float rc;
  const float *pos, *s, *v, *W, *b, *freq, *gx, *gV;
  const float *a_s, *a_v, *a_pos;  // ORDER 2 only
  float *o_s, *o_v;                // [N,H], [N,D]
  float* gr;                       // [S, E, 3] per-edge d/dr partials (one slab per channel slice)
  float* wpart;                    // [gridDim.x, H, 2*NBP] weight-gradient partials
  int n_tiles;
};

template <int L, int ROLE, int C, int M1, int M2, int ORDER, bool WGRAD>
__device__ __forceinline__ void neighbor_role(const NeighborArgs& A, NeighborSmem<ORDER, WGRAD>& sm) {
  constexpr int M = C + M1 + M2, D = C + 3 * M1 + 5 * M2, H = C + 2 * M;
  using Thread = NeighborThread<float, L, ROLE, WGRAD, NK_>;
  constexpr int NC = Thread::NC;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int h = blockIdx.y * NTHREADS + t;
  const int q = (ROLE == ROLE_STATE) ? h : (ROLE == ROLE_EDGE ? h - M : 0);
  const int vbase = (ROLE == ROLE_SCALAR) ? 0 : ((L == 0) ? q : (L == 1 ? C + (q - C) : C + 3 * M1 + (q - C - M1)));
  constexpr int vstride = (L == 0) ? 0 : (L == 1 ? M1 : M2);
  const xeq_graph_t& g = A.g;
  const int N = g.n_nodes, E = g.n_edges;
  const int* __restrict__ rp_ = g.t_rowptr;

  Thread th;
  load_wrow<NK_>(A.W, A.b, h, th.Wt);
  if (WGRAD) th.reset_wgrad();
  th.s = th.sd = 0.f;
#pragma unroll
  for (int m = 0; m < NC; ++m) th.v[m] = th.vd[m] = 0.f;
  th.reset_node();

  auto emit = [&](int node, bool with_acc) {
    if (A.o_s) A.o_s[(size_t)node * H + h] = with_acc ? th.acc_s : 0.f;
    if (ROLE == ROLE_STATE && A.o_v) {
#pragma unroll
      for (int m = 0; m < NC; ++m) A.o_v[(size_t)node * D + vbase + m * vstride] = with_acc ? th.acc_v[m] : 0.f;
    }
  };

  for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
    const int n0 = lower_bound_nodes(rp_, N, tile * NT);
    const int n1 = lower_bound_nodes(rp_, N, (tile + 1) * NT);
    if (n0 == n1) continue;
    const int e0 = rp_[n0], e1 = rp_[n1];
    int cur = -1;
    for (int eb = e0; eb < e1; eb += NT) {
      const int cnt = min(NT, e1 - eb);
      __syncthreads();
      if (t < cnt) {  // ---- geometry, one thread per slot
        const int sl = eb + t;
        const int j = owner_of(rp_, n0, n1, sl);
        const int i = g.t_row[sl];
        const int e = g.t_eid[sl];
        float r[3], d, u[3];
        edge_vector(g, A.pos, i, j, e, r);
        unit_vector(r, d, u);
        float G[3][8];
        angular_first(u, d, sm.Y[t], G);
#pragma unroll
        for (int x = 0; x < 3; ++x) {
#pragma unroll
          for (int m = 0; m < 8; ++m) sm.G[t][x * 8 + m] = G[x][m];
          sm.u[t][x] = u[x];
        }
        if (ORDER == 2) {
          float Hm[3][8], rp[3], rdot[3] = {0.f, 0.f, 0.f}, dd;
          if (A.a_pos) {
#pragma unroll
            for (int x = 0; x < 3; ++x) rdot[x] = A.a_pos[3 * i + x] - A.a_pos[3 * j + x];
          }
          angular_second(u, d, rdot, G, dd, rp, sm.Ydot[t], Hm);
          sm.ddot[t] = dd;
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            sm.rp[t][x] = rp[x];
#pragma unroll
            for (int m = 0; m < 8; ++m) sm.Hm[t][x * 8 + m] = Hm[x][m];
          }
        }
        const Cutoff<float> c = cutoff_terms(d, A.rc);
        sm.d[t] = d;
        sm.chi[t][0] = c.chi; sm.chi[t][1] = c.dchi; sm.chi[t][2] = c.ddchi;
        sm.psi[t][0] = c.chi;
        sm.dpsi[t][0] = c.dchi;
        if (ORDER == 2) sm.ddpsi[t][0] = c.ddchi;
        if (WGRAD) sm.xi[t][0] = 0.f;
        if (WGRAD && ORDER == 2) sm.dxi[t][0] = 0.f;
#pragma unroll
        for (int k = NK_; k < NBP; ++k) {
          sm.psi[t][k] = 0.f; sm.dpsi[t][k] = 0.f;
          if (ORDER == 2) sm.ddpsi[t][k] = 0.f;
          if (WGRAD) sm.xi[t][k] = 0.f;
          if (WGRAD && ORDER == 2) sm.dxi[t][k] = 0.f;
        }
        sm.gat[t] = i; sm.own[t] = j; sm.eid[t] = e;
      }
      __syncthreads();
      for (int idx = t; idx < cnt * NB_; idx += NTHREADS) {  // ---- radial terms per (slot, k)
        const int ee = idx / NB_, k = idx - ee * NB_;
        Cutoff<float> c;
        c.chi = sm.chi[ee][0]; c.dchi = sm.chi[ee][1]; c.ddchi = sm.chi[ee][2];
        const Radial<float> rr = radial_term(sm.d[ee], A.freq[k], A.rc, c);
        sm.psi[ee][k + 1] = rr.psi;
        sm.dpsi[ee][k + 1] = rr.dpsi;
        if (ORDER == 2) sm.ddpsi[ee][k + 1] = rr.ddpsi;
        if (WGRAD) sm.xi[ee][k + 1] = rr.xi;
        if (WGRAD && ORDER == 2) sm.dxi[ee][k + 1] = rr.dxi;
      }
      __syncthreads();

      float gg[NC];
      auto gather = [&](int ee, float* a) {
        const int i = sm.gat[ee];
        if (ROLE == ROLE_SCALAR) {
          a[0] = A.gx[(size_t)i * C + (h - 2 * M)];
        } else {
#pragma unroll
          for (int m = 0; m < NC; ++m) a[m] = A.gV[(size_t)i * D + vbase + m * vstride];
        }
      };
      gather(0, gg);
      for (int ee = 0; ee < cnt; ++ee) {
        float ng[NC];
#pragma unroll
        for (int m = 0; m < NC; ++m) ng[m] = 0.f;
        if (ee + 1 < cnt) gather(ee + 1, ng);
        const int j = sm.own[ee];
        if (j != cur) {
          if (cur >= 0) emit(cur, true);
          for (int nn = (cur >= 0 ? cur + 1 : n0); nn < j; ++nn) emit(nn, false);
          cur = j;
          th.reset_node();
          th.s = A.s[(size_t)j * H + h];
          if (ORDER == 2) th.sd = A.a_s ? A.a_s[(size_t)j * H + h] : 0.f;
          if (ROLE == ROLE_STATE) {
#pragma unroll
            for (int m = 0; m < NC; ++m) {
              th.v[m] = A.v[(size_t)j * D + vbase + m * vstride];
              if (ORDER == 2) th.vd[m] = A.a_v ? A.a_v[(size_t)j * D + vbase + m * vstride] : 0.f;
            }
          }
        }
        float p[NBP], dp[NBP], ddp[NBP], xx[NBP], dxx[NBP];
        lds_row(sm.psi[ee], p);
        lds_row(sm.dpsi[ee], dp);
        if (ORDER == 2) lds_row(sm.ddpsi[ee], ddp);
        if (WGRAD) lds_row(sm.xi[ee], xx);
        if (WGRAD && ORDER == 2) lds_row(sm.dxi[ee], dxx);
        NbrEdge<float> ne;
        ne.psi = p; ne.dpsi = dp; ne.ddpsi = ddp; ne.xi = xx; ne.dxi = dxx;
        ne.Y = sm.Y[ee]; ne.G = sm.G[ee];
        ne.Hm = (ORDER == 2) ? sm.Hm[ee] : nullptr;
        ne.Ydot = (ORDER == 2) ? sm.Ydot[ee] : nullptr;
        ne.u = sm.u[ee];
        ne.rp = (ORDER == 2) ? sm.rp[ee] : nullptr;
        ne.ddot = (ORDER == 2) ? sm.ddot[ee] : 0.f;
        float pr[3];
        if (ORDER == 1) th.first(ne, gg, pr); else th.second(ne, gg, pr);
#pragma unroll
        for (int x = 0; x < 3; ++x) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) pr[x] += __shfl_xor_sync(0xffffffffu, pr[x], o);
        }
        if (lane == 0) {
          sm.red[ee][warp][0] = pr[0]; sm.red[ee][warp][1] = pr[1]; sm.red[ee][warp][2] = pr[2];
        }
#pragma unroll
        for (int m = 0; m < NC; ++m) gg[m] = ng[m];
      }
      __syncthreads();
      if (t < cnt && A.gr) {  // ---- per-edge d/dr: fixed-order sum over the warps of this channel slice
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) { a0 += sm.red[t][w][0]; a1 += sm.red[t][w][1]; a2 += sm.red[t][w][2]; }
        float* dst = A.gr + ((size_t)blockIdx.y * E + sm.eid[t]) * 3;
        dst[0] = a0; dst[1] = a1; dst[2] = a2;
      }
    }
    if (cur >= 0) emit(cur, true);
    for (int nn = (cur >= 0 ? cur + 1 : n0); nn < n1; ++nn) emit(nn, false);
  }
  if (WGRAD) {
    float* dst = A.wpart + ((size_t)blockIdx.x * H + h) * (2 * NBP);
#pragma unroll
    for (int k = 0; k < NBP; ++k) { dst[k] = th.GW[k]; dst[NBP + k] = th.GF[k]; }
  }
}

template <int C, int M1, int M2, int ORDER, bool WGRAD>
__global__ void __launch_bounds__(NTHREADS) neighbor_kernel(const NeighborArgs A) {
  constexpr int M = C + M1 + M2;
  __shared__ __align__(16) NeighborSmem<ORDER, WGRAD> sm;
  const int h = blockIdx.y * NTHREADS + threadIdx.x;
  if (h < C) neighbor_role<0, ROLE_STATE, C, M1, M2, ORDER, WGRAD>(A, sm);
  else if (h < C + M1) neighbor_role<1, ROLE_STATE, C, M1, M2, ORDER, WGRAD>(A, sm);
  else if (h < M) neighbor_role<2, ROLE_STATE, C, M1, M2, ORDER, WGRAD>(A, sm);
  else if (h < M + C) neighbor_role<0, ROLE_EDGE, C, M1, M2, ORDER, WGRAD>(A, sm);
  else if (h < M + C + M1) neighbor_role<1, ROLE_EDGE, C, M1, M2, ORDER, WGRAD>(A, sm);
  else if (h < 2 * M) neighbor_role<2, ROLE_EDGE, C, M1, M2, ORDER, WGRAD>(A, sm);
  else neighbor_role<0, ROLE_SCALAR, C, M1, M2, ORDER, WGRAD>(A, sm);
}

// gpos[n] = sum_{e in row n} gr[e] - sum_{slot in t-row n} gr[t_eid[slot]]  (summed over channel slices)
__global__ void pos_grad_kernel(xeq_graph_t g, const float* __restrict__ gr, int n_slices, float* __restrict__ gpos) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.n_nodes) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int sl = 0; sl < n_slices; ++sl) {
    const float* base = gr + (size_t)sl * g.n_edges * 3;
    for (int e = g.rowptr[n]; e < g.rowptr[n + 1]; ++e) {
      a0 += base[3 * (size_t)e]; a1 += base[3 * (size_t)e + 1]; a2 += base[3 * (size_t)e + 2];
    }
    for (int s = g.t_rowptr[n]; s < g.t_rowptr[n + 1]; ++s) {
      const size_t e = g.t_eid[s];
      a0 -= base[3 * e]; a1 -= base[3 * e + 1]; a2 -= base[3 * e + 2];
    }
  }
  gpos[3 * n] = a0; gpos[3 * n + 1] = a1; gpos[3 * n + 2] = a2;
}

// weight-gradient partials [nblk, H, 2*NBP] -> gW [H,B], gb [H], tot_f [H, NB] (fixed order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ wpart, int nblk, int H, float* __restrict__ gW,
                                    float* __restrict__ gb, float* __restrict__ ftot) {
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
  XEQ_CHECK_ARG(!need_t || (g->t_rowptr && (g->n_edges == 0 || (g->t_row && g->t_eid))), "graph: transposed CSR missing");
  XEQ_CHECK_ARG((g->offsets == nullptr) == (g->cell == nullptr), "graph: offsets and cell must be given together");
  XEQ_CHECK_ARG(!g->cell || g->n_graphs == 1 || g->node_graph, "graph: node_graph needed for multi-graph PBC");
  return XEQ_OK;
}

static inline int dims_H(const xeq_dims_t* d) { return d->node_dim + 2 * (d->mul0 + d->mul1 + d->mul2); }

template <int C, int M1, int M2, bool JVP>
static int launch_center(const CenterArgs& A, cudaStream_t st) {
  const int grid = min(A.n_tiles, num_sms() * (C == 128 ? 2 : 1));
  center_kernel<C, M1, M2, JVP><<<grid, C + M1 + M2, 0, st>>>(A);
  XEQ_LAUNCH_CHECK();
  return XEQ_OK;
}

static int run_center(const xeq_graph_t* g, const xeq_dims_t* dims, CenterArgs& A, bool jvp, cudaStream_t st) {
  int cfg;
  int rc = check_dims(dims, &cfg);
  if (rc) return rc;
  rc = check_graph(g, false);
  if (rc) return rc;
  if (g->n_nodes == 0) return XEQ_OK;
  A.g = *g;
  A.rc = dims->cutoff;
  A.n_tiles = g->n_edges / CT + 1;
  if (cfg == 0) return jvp ? launch_center<128, 64, 32, true>(A, st) : launch_center<128, 64, 32, false>(A, st);
  return jvp ? launch_center<256, 128, 64, true>(A, st) : launch_center<256, 128, 64, false>(A, st);
}

static int neighbor_grid_x(const xeq_graph_t* g) {
  const int n_tiles = g->n_edges / NT + 1;
  return min(n_tiles, num_sms() * 2);
}

static size_t neighbor_ws_bytes(const xeq_graph_t* g, const xeq_dims_t* d, int want_wgrad) {
  const int H = dims_H(d), S = H / NTHREADS;
  size_t b = 256 + align_up(sizeof(float) * 3 * (size_t)S * (size_t)(g->n_edges > 0 ? g->n_edges : 1), 256);
  if (want_wgrad) {
    b += align_up(sizeof(float) * (size_t)neighbor_grid_x(g) * H * 2 * NBP, 256);
    b += align_up(sizeof(float) * (size_t)H * NB_, 256);
  }
  return b;
}

template <int C, int M1, int M2, int ORDER>
static int launch_neighbor(NeighborArgs& A, bool wgrad, int gx, cudaStream_t st) {
  constexpr int H = C + 2 * (C + M1 + M2);
  static_assert(H % NTHREADS == 0, "channel slices must tile H");
  dim3 grid(gx, H / NTHREADS);
  if (wgrad) neighbor_kernel<C, M1, M2, ORDER, true><<<grid, NTHREADS, 0, st>>>(A);
  else neighbor_kernel<C, M1, M2, ORDER, false><<<grid, NTHREADS, 0, st>>>(A);
  XEQ_LAUNCH_CHECK();
  return XEQ_OK;
}

static int run_neighbor(const xeq_graph_t* g, const xeq_dims_t* dims, NeighborArgs& A, int order, float* o_pos,
                        float* o_W, float* o_b, float* o_f, void* ws, size_t ws_bytes, cudaStream_t st) {
  int cfg;
  int rc = check_dims(dims, &cfg);
  if (rc) return rc;
  rc = check_graph(g, true);
  if (rc) return rc;
  const bool wgrad = o_W || o_b || o_f;
  XEQ_CHECK_ARG(!wgrad || (o_W && o_b && o_f), "weight gradients: gW, gb and gfreq must be given together");
  XEQ_CHECK_ARG(ws && ws_bytes >= neighbor_ws_bytes(g, dims, wgrad), "edge_message backward: workspace too small");
  const int H = dims_H(dims), S = H / NTHREADS;
  if (g->n_nodes == 0) return XEQ_OK;
  Carver cv(ws);
  float* gr = cv.take<float>(3 * (size_t)S * (size_t)(g->n_edges > 0 ? g->n_edges : 1));
  const int gx = neighbor_grid_x(g);
  float *wpart = nullptr, *ftot = nullptr;
  if (wgrad) {
    wpart = cv.take<float>((size_t)gx * H * 2 * NBP);
    ftot = cv.take<float>((size_t)H * NB_);
  }
  A.g = *g;
  A.rc = dims->cutoff;
  A.n_tiles = g->n_edges / NT + 1;
  A.gr = o_pos ? gr : nullptr;
  A.wpart = wpart;
  if (cfg == 0) rc = order == 1 ? launch_neighbor<128, 64, 32, 1>(A, wgrad, gx, st) : launch_neighbor<128, 64, 32, 2>(A, wgrad, gx, st);
  else rc = order == 1 ? launch_neighbor<256, 128, 64, 1>(A, wgrad, gx, st) : launch_neighbor<256, 128, 64, 2>(A, wgrad, gx, st);
  if (rc) return rc;
  if (o_pos) {
    pos_grad_kernel<<<(g->n_nodes + 127) / 128, 128, 0, st>>>(*g, gr, S, o_pos);
    XEQ_LAUNCH_CHECK();
  }
  if (wgrad) {
    wgrad_reduce_kernel<<<H, 2 * NBP, 0, st>>>(wpart, gx, H, o_W, o_b, ftot);
    freq_grad_kernel<<<NB_, 256, 0, st>>>(A.W, ftot, H, o_f);
    XEQ_LAUNCH_CHECK();
  }
  return XEQ_OK;
}

}  // namespace xeq

using namespace xeq;

extern "C" {

int xeq_edge_message_fwd(const xeq_graph_t* g, const xeq_dims_t* dims, const float* pos, const float* s, const float* v,
                         const float* x_in, const float* V_in, const float* W_rbf, const float* b_rbf, const float* freq,
                         float* x_out, float* V_out, xeq_stream_t stream) {
  XEQ_CHECK_ARG(pos && s && v && W_rbf && b_rbf && freq && x_out && V_out, "edge_message_fwd: NULL argument");
  CenterArgs A{};
  A.pos = pos; A.s = s; A.v = v; A.x_in = x_in; A.V_in = V_in; A.W = W_rbf; A.b = b_rbf; A.freq = freq;
  A.x_out = x_out; A.V_out = V_out;
  return run_center(g, dims, A, false, (cudaStream_t)stream);
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
  A.pos = pos; A.s = s; A.v = v; A.W = W_rbf; A.b = b_rbf; A.freq = freq; A.gx = gx; A.gV = gV;
  A.o_s = gs; A.o_v = gv;
  return run_neighbor(g, dims, A, 1, gpos, gW, gb, gfreq, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t xeq_edge_message_bwdbwd_workspace_bytes(const xeq_graph_t* g, const xeq_dims_t* dims, int want_wgrad) {
  if (!g || !dims) return 0;
  return neighbor_ws_bytes(g, dims, want_wgrad);
}

int xeq_edge_message_bwdbwd(const xeq_graph_t* g, const xeq_dims_t* dims, const float* pos, const float* s, const float* v,
                            const float* W_rbf, const float* b_rbf, const float* freq, const float* gx, const float* gV,
                            const float* a_s, const float* a_v, const float* a_pos, float* o_gx, float* o_gV, float* o_s,
                            float* o_v, float* o_pos, float* o_W, float* o_b, float* o_freq, void* workspace,
                            size_t workspace_bytes, xeq_stream_t stream) {
  XEQ_CHECK_ARG(pos && s && v && W_rbf && b_rbf && freq && gx && gV, "edge_message_bwdbwd: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (o_gx || o_gV) {  // d/d(gx, gV): tangent of the forward message along (a_s, a_v, a_pos)
    XEQ_CHECK_ARG(o_gx && o_gV, "edge_message_bwdbwd: o_gx and o_gV must be given together");
    CenterArgs C{};
    C.pos = pos; C.s = s; C.v = v; C.W = W_rbf; C.b = b_rbf; C.freq = freq;
    C.a_s = a_s; C.a_v = a_v; C.a_pos = a_pos; C.x_out = o_gx; C.V_out = o_gV;
    int rc = run_center(g, dims, C, true, st);
    if (rc) return rc;
  }
  if (o_s || o_v || o_pos || o_W) {
    NeighborArgs A{};
    A.pos = pos; A.s = s; A.v = v; A.W = W_rbf; A.b = b_rbf; A.freq = freq; A.gx = gx; A.gV = gV;
    A.a_s = a_s; A.a_v = a_v; A.a_pos = a_pos; A.o_s = o_s; A.o_v = o_v;
    return run_neighbor(g, dims, A, 2, o_pos, o_W, o_b, o_freq, workspace, workspace_bytes, st);
  }
  return XEQ_OK;
}

}  // extern "C"
