// Shared device pieces of the fused edge kernels (edge_message.cu: test-only SIMT filter contraction;
// edge_*_ul.cu: filter contraction on tcgen05): chunk stream over node-aligned tiles,
// per-edge geometry records in shared memory and the stages that fill them, kernel arguments.
#pragma once
#include "common.cuh"
#include "edge_thread.cuh"

namespace xeq {


constexpr int NB_ = 20;       // num_basis instantiated
constexpr int NK_ = NB_ + 1;  // + bias/cutoff term
constexpr int CT = 64;        // edges per chunk, center kernels
constexpr int NT = 32;        // slots per chunk, neighbor kernels
constexpr int WTHREADS = 288; // nbr_wgrad: filter channels per CTA; grid.y = H / 288 slices

// ------------------------------------------------------------------------------------------
// chunk stream: the (tile, chunk) work items of one CTA, in order
// ------------------------------------------------------------------------------------------
struct ChunkDesc {
  int n0, n1;  // node range of the tile this chunk belongs to
  int eb;      // first edge/slot of the chunk
  int cnt;     // edges in the chunk (0 for an edge-less tile), -1 = end of stream
  int first;   // first chunk of its tile
  int last;    // last chunk of its tile
  int owner;   // RowCursor: the node whose row this chunk is a piece of; ChunkCursor: -1 (rows are looked up per edge)
  int rfirst;  // RowCursor: first / last piece of that row
  int rlast;
};

template <int T>
struct ChunkCursor {
  const int* __restrict__ rowptr;
  const int* __restrict__ tile_ptr;
  int n_tiles, tile, n0, n1, e0, e1, eb;
  bool valid;

  __device__ __forceinline__ void load_tile() {
    valid = false;
    while (tile < n_tiles) {
      n0 = tile_ptr[tile];
      n1 = tile_ptr[tile + 1];
      if (n0 < n1) {
        e0 = rowptr[n0];
        e1 = rowptr[n1];
        eb = e0;
        valid = true;
        return;
      }
      tile += gridDim.x;
    }
  }
  __device__ __forceinline__ void init(const int* rp, const int* tp, int nt) {
    rowptr = rp; tile_ptr = tp; n_tiles = nt; tile = blockIdx.x;
    load_tile();
  }
  __device__ __forceinline__ ChunkDesc next() {
    ChunkDesc d;
    if (!valid) {
      d.n0 = d.n1 = d.eb = 0; d.cnt = -1; d.first = d.last = 0; d.owner = -1; d.rfirst = d.rlast = 0;
      return d;
    }
    d.owner = -1; d.rfirst = d.rlast = 0;
    d.n0 = n0; d.n1 = n1; d.eb = eb;
    d.cnt = min(T, e1 - eb);
    d.first = (eb == e0);
    d.last = (eb + T >= e1);
    eb += T;
    if (eb >= e1) {
      tile += gridDim.x;
      load_tile();
    }
    return d;
  }
};

// Row-aligned chunk stream (tcgen05 kernels): every chunk is a piece (<= T edges) of ONE row, so the
// per-edge loop of a chunk has no row switch in it (branch-free, software-pipelined by the compiler);
// a row without edges yields one chunk with cnt = 0.
template <int T>
struct RowCursor {
  const int* __restrict__ rowptr;
  const int* __restrict__ tile_ptr;
  int n_tiles, tile, n0, n1, node, e, e0, e1;
  bool valid, tfirst;

  __device__ __forceinline__ void load_row() {
    e0 = rowptr[node];
    e1 = rowptr[node + 1];
    e = e0;
  }
  __device__ __forceinline__ void load_tile() {
    valid = false;
    while (tile < n_tiles) {
      n0 = tile_ptr[tile];
      n1 = tile_ptr[tile + 1];
      if (n0 < n1) {
        node = n0;
        load_row();
        valid = true;
        tfirst = true;
        return;
      }
      tile += gridDim.x;
    }
  }
  __device__ __forceinline__ void init(const int* rp, const int* tp, int nt) {
    rowptr = rp; tile_ptr = tp; n_tiles = nt; tile = blockIdx.x;
    load_tile();
  }
  __device__ __forceinline__ ChunkDesc next() {
    ChunkDesc d;
    if (!valid) {
      d.n0 = d.n1 = d.eb = 0; d.cnt = -1; d.first = d.last = 0; d.owner = -1; d.rfirst = d.rlast = 0;
      return d;
    }
    d.n0 = n0; d.n1 = n1; d.eb = e; d.owner = node;
    d.cnt = max(0, min(T, e1 - e));
    d.first = tfirst;
    d.rfirst = (e == e0);
    d.rlast = (e + T >= e1);
    tfirst = false;
    e += T;
    d.last = 0;
    if (d.rlast) {
      ++node;
      if (node >= n1) {
        d.last = 1;
        tile += gridDim.x;
        load_tile();
      } else {
        load_row();
      }
    }
    return d;
  }
};

// node that owns edge/slot e inside [n0, n1): rowptr[i] <= e < rowptr[i+1]
__device__ __forceinline__ int owner_of(const int* __restrict__ rowptr, int n0, int n1, int e) {
  int lo = n0, hi = n1;
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

__device__ __forceinline__ void load_wrow(const float* __restrict__ W, const float* __restrict__ b, int h, float* row) {
  row[0] = b[h];
#pragma unroll
  for (int k = 0; k < NB_; ++k) row[k + 1] = W[(size_t)h * NB_ + k];
#pragma unroll
  for (int k = NK_; k < NBP; ++k) row[k] = 0.f;
}

__device__ __forceinline__ void lds_row(const float* __restrict__ src, float* dst) {
#pragma unroll
  for (int k = 0; k < NBP / 4; ++k) {
    const float4 p = reinterpret_cast<const float4*>(src)[k];
    dst[4 * k] = p.x; dst[4 * k + 1] = p.y; dst[4 * k + 2] = p.z; dst[4 * k + 3] = p.w;
  }
}

// Explicit shared-space accessors for the dynamically sized row window (keeps the addressing in
// 32-bit shared space; a generic pointer would re-derive the shared window base per access).
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
extern __shared__ __align__(16) unsigned char xeq_dyn_smem[];

// Rows a CTA can stage per tile ("window").  With molecule tiles (tile_mode 1) every neighbor of a
// tile's nodes lies inside the tile's own node range, so the CTA copies those rows to shared memory
// once (each thread only ever touches its own columns -> no barrier) and the per-edge gathers become
// shared-memory reads: HBM/L2 traffic drops from E rows to N rows.
template <int C, bool JVP> struct CenterWin { static constexpr int value = (C == 128) ? 21 : 0; };  // fwd: 2 CTAs/SM
template <int C> struct NbrWin { static constexpr int value = (C == 128) ? 24 : 0; };

// ------------------------------------------------------------------------------------------
// shared-memory geometry records and the two geometry stages
// ------------------------------------------------------------------------------------------
template <int T, bool NEED_G, bool SECOND>
struct alignas(16) GeoA {  // stage A1: one thread per edge
  float Y[T][8];
  float u[T][4];
  float d[T];
  float chi[T][3];
  int gat[T];  // node whose rows are gathered (neighbor j for center kernels, center i for neighbor kernels)
  int own[T];  // node that owns the row being walked
  int eid[T];  // canonical edge id
  float G[NEED_G ? T : 1][24];
  float Hm[(NEED_G && SECOND) ? T : 1][24];
  float Ydot[SECOND ? T : 1][8];
  float rp[SECOND ? T : 1][4];
  float ddot[SECOND ? T : 1];
};

template <int T, bool D1, bool D2, bool XI, bool DXI>
struct alignas(16) GeoB {  // stage A2: one thread per (edge, k)
  float psi[T][NBP];
  float dpsi[D1 ? T : 1][NBP];
  float ddpsi[D2 ? T : 1][NBP];
  float xi[XI ? T : 1][NBP];
  float dxi[DXI ? T : 1][NBP];
};

struct GeoArgs {
  xeq_graph_t g;
  const float* pos;
  const float* a_pos;  // tangent of pos (second order) or NULL
  const float* a_cell; // tangent of the lattice [G,3,3] (second order, periodic virial in a training loss) or NULL
  const float* freq;
  float rc;
};

// Geometry record of one edge slot t from its edge vector r (and, second order, the tangent rdot of r):
// distances, harmonics (+ dY/dr, Hessian . rdot), cutoff terms, indices.
template <int T, bool NEED_G, bool SECOND>
__device__ __forceinline__ void geo_record(const GeoArgs& A, GeoA<T, NEED_G, SECOND>& sa, const int t, const int gat,
                                           const int owner, const int e, const float r[3], const float rdot[3]) {
  sa.gat[t] = gat;
  sa.own[t] = owner;
  sa.eid[t] = e;
  float dist, u[3];
  unit_vector(r, dist, u);
  if (!NEED_G && !SECOND) {
    sph_harm(u, sa.Y[t]);
  } else {
    float G[3][8];
    angular_first(u, dist, sa.Y[t], G);
    if (NEED_G) {
#pragma unroll
      for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int m = 0; m < 8; ++m) sa.G[t][x * 8 + m] = G[x][m];
    }
    if (SECOND) {
      float Hm[3][8], rp[3], dd;
      angular_second(u, dist, rdot, G, dd, rp, sa.Ydot[t], Hm);
      sa.ddot[t] = dd;
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        sa.rp[t][x] = rp[x];
        if (NEED_G) {
#pragma unroll
          for (int m = 0; m < 8; ++m) sa.Hm[t][x * 8 + m] = Hm[x][m];
        }
      }
    }
  }
#pragma unroll
  for (int x = 0; x < 3; ++x) sa.u[t][x] = u[x];
  const Cutoff<float> c = cutoff_terms(dist, A.rc);
  sa.d[t] = dist;
  sa.chi[t][0] = c.chi; sa.chi[t][1] = c.dchi; sa.chi[t][2] = c.ddchi;
}

// TRANSPOSED = false: walk CSR rows (owner = center i, gathered = neighbor j = col[e], eid = e)
// TRANSPOSED = true : walk transposed rows (owner = neighbor j, gathered = center i = t_row[sl], eid = t_eid[sl])
template <int T, bool TRANSPOSED, bool NEED_G, bool SECOND>
__device__ __noinline__ void geo_stage_a1(const GeoArgs& A, const ChunkDesc d, GeoA<T, NEED_G, SECOND>& sa,
                                          const int t /* edge slot of this thread: threadIdx.x, or the lane of a producer warp */) {
  // RowCursor chunks: the slots past the end of the row piece repeat its last edge, so the per-edge loops of the
  // tcgen05 kernels can run whole groups without bounds checks (their filter values are exact zeros)
  if (t >= d.cnt && (d.owner < 0 || d.cnt <= 0 || t >= T)) return;
  const xeq_graph_t& g = A.g;
  const int sl = d.eb + min(t, d.cnt - 1);
  int i, j, e, owner;
  if (!TRANSPOSED) {
    owner = d.owner >= 0 ? d.owner : owner_of(g.rowptr, d.n0, d.n1, sl);
    i = owner; j = g.col[sl]; e = sl;
  } else {
    owner = d.owner >= 0 ? d.owner : owner_of(g.t_rowptr, d.n0, d.n1, sl);
    j = owner; i = g.t_row[sl]; e = g.t_eid[sl];
  }
  float r[3], rdot[3] = {0.f, 0.f, 0.f};
  edge_vector(g, A.pos, i, j, e, r);
  if (SECOND && A.a_pos) {
#pragma unroll
    for (int x = 0; x < 3; ++x) rdot[x] = A.a_pos[3 * i + x] - A.a_pos[3 * j + x];
  }
  if (SECOND && A.a_cell && g.offsets != nullptr) {  // r = p_i - p_j - o @ cell  =>  rdot -= o @ a_cell
    const char4 o = reinterpret_cast<const char4*>(g.offsets)[e];
    const float* c = A.a_cell + 9 * (g.node_graph ? g.node_graph[j] : 0);
    const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
#pragma unroll
    for (int x = 0; x < 3; ++x) rdot[x] -= ox * c[x] + oy * c[3 + x] + oz * c[6 + x];
  }
  geo_record<T, NEED_G, SECOND>(A, sa, t, TRANSPOSED ? i : j, owner, e, r, rdot);
}

// The same per-edge geometry as a three-stage software pipeline inside ONE producer warp (one lane per edge slot
// of a RowCursor chunk), so that none of the dependent global loads  rowptr -> col / t_row, t_eid -> pos  is waited
// for in the step that issued it:  A(k+2) indices | B(k+1) raw position loads | C(k) arithmetic -> shared memory.
template <int T, bool TRANSPOSED, bool NEED_G, bool SECOND>
struct GeoPipe {
  ChunkDesc dB, dC;
  int iB, jB, eB, iC, jC, eC;
  float pi[3], pj[3], sh[3], ai[3], aj[3], ash[3];

  __device__ __forceinline__ void init() {
    dB.cnt = dC.cnt = -1;
    dB.owner = dC.owner = 0;
    iB = jB = eB = iC = jC = eC = 0;
#pragma unroll
    for (int x = 0; x < 3; ++x) pi[x] = pj[x] = sh[x] = ai[x] = aj[x] = ash[x] = 0.f;
  }
  __device__ __forceinline__ void stage_a(const GeoArgs& A, const ChunkDesc& d, const int lane) {
    dB = d;
    if (d.cnt > 0 && lane < T) {
      const int sl = d.eb + min(lane, d.cnt - 1);
      if (!TRANSPOSED) { iB = d.owner; jB = A.g.col[sl]; eB = sl; }
      else { jB = d.owner; iB = A.g.t_row[sl]; eB = A.g.t_eid[sl]; }
    }
  }
  __device__ __forceinline__ void stage_b(const GeoArgs& A, const int lane) {
    dC = dB; iC = iB; jC = jB; eC = eB;
    if (dB.cnt > 0 && lane < T) {
      const xeq_graph_t& g = A.g;
#pragma unroll
      for (int x = 0; x < 3; ++x) {  // loads only: the arithmetic happens in stage C, after the latency has passed
        pi[x] = A.pos[3 * iB + x];
        pj[x] = A.pos[3 * jB + x];
        sh[x] = 0.f;
        if (SECOND) {
          ai[x] = A.a_pos ? A.a_pos[3 * iB + x] : 0.f;
          aj[x] = A.a_pos ? A.a_pos[3 * jB + x] : 0.f;
        }
      }
      if (g.offsets != nullptr) {  // nn/basic.py:119-128: vectors -= cell_offsets @ cell[graph(neighbor)]
        const char4 o = reinterpret_cast<const char4*>(g.offsets)[eB];
        const float* cl = g.cell + 9 * (g.node_graph ? g.node_graph[jB] : 0);
        const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
#pragma unroll
        for (int x = 0; x < 3; ++x) sh[x] = ox * cl[x] + oy * cl[3 + x] + oz * cl[6 + x];
        if (SECOND && A.a_cell) {  // tangent of the lattice shift
          const float* ac = A.a_cell + 9 * (g.node_graph ? g.node_graph[jB] : 0);
#pragma unroll
          for (int x = 0; x < 3; ++x) ash[x] = ox * ac[x] + oy * ac[3 + x] + oz * ac[6 + x];
        }
      }
    }
  }
  __device__ __forceinline__ void stage_c(const GeoArgs& A, GeoA<T, NEED_G, SECOND>& sa, const int lane) {
    if (dC.cnt > 0 && lane < T) {
      float r[3], rdot[3];
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        r[x] = (pi[x] - pj[x]) - sh[x];
        rdot[x] = SECOND ? (ai[x] - aj[x]) - ash[x] : 0.f;
      }
      geo_record<T, NEED_G, SECOND>(A, sa, lane, TRANSPOSED ? iC : jC, dC.owner, eC, r, rdot);
    }
  }
};

template <int T, int THREADS, bool NEED_G, bool SECOND, bool D1, bool D2, bool XI, bool DXI>
__device__ __noinline__ void geo_stage_a2(const GeoArgs& A, int cnt, const GeoA<T, NEED_G, SECOND>& sa,
                                          GeoB<T, D1, D2, XI, DXI>& sb) {
  const int t = threadIdx.x;
  for (int idx = t; idx < cnt * NK_; idx += THREADS) {
    const int ee = idx / NK_, k = idx - ee * NK_;
    Cutoff<float> c;
    c.chi = sa.chi[ee][0]; c.dchi = sa.chi[ee][1]; c.ddchi = sa.chi[ee][2];
    if (k == 0) {  // bias / cutoff term, plus the zero padding of the row
      sb.psi[ee][0] = c.chi;
      if (D1) sb.dpsi[ee][0] = c.dchi;
      if (D2) sb.ddpsi[ee][0] = c.ddchi;
      if (XI) sb.xi[ee][0] = 0.f;
      if (DXI) sb.dxi[ee][0] = 0.f;
#pragma unroll
      for (int kk = NK_; kk < NBP; ++kk) {
        sb.psi[ee][kk] = 0.f;
        if (D1) sb.dpsi[ee][kk] = 0.f;
        if (D2) sb.ddpsi[ee][kk] = 0.f;
        if (XI) sb.xi[ee][kk] = 0.f;
        if (DXI) sb.dxi[ee][kk] = 0.f;
      }
    } else {
      const Radial<float> rr = radial_term(sa.d[ee], A.freq[k - 1], A.rc, c);
      sb.psi[ee][k] = rr.psi;
      if (D1) sb.dpsi[ee][k] = rr.dpsi;
      if (D2) sb.ddpsi[ee][k] = rr.ddpsi;
      if (XI) sb.xi[ee][k] = rr.xi;
      if (DXI) sb.dxi[ee][k] = rr.dxi;
    }
  }
}


struct CenterArgs {
  GeoArgs geo;
  const float *s, *v, *x_in, *V_in, *W, *b;
  const float *a_s, *a_v;  // JVP only
  float *x_out, *V_out;
};

struct NeighborArgs {
  GeoArgs geo;
  const float *s, *v, *W, *b, *gx, *gV;
  const float *a_s, *a_v;  // ORDER 2 only
  float *o_s, *o_v;        // [N,H], [N,D]
  float* gr;               // [slices, E, 3] per-edge d/dr
  float* wpart;            // [gridDim.x, H, 2*NBP] weight-gradient partials (wgrad kernel)
};

}  // namespace xeq
