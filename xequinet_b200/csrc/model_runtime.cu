// Whole-model inference runtime: XPaiNN energy + forces as ONE C call, without Python, torch or autograd.
//
// This is the deployment form of the path: the reference ships a TorchScript archive that a
// LAMMPS pair style / the GROMACS NNP interface runs through libtorch (run/jit_script.py:28-86,
// interface/jit_model.py:12-216: compute_edge_data -> mods -> compute_properties with torch.autograd.grad for the
// forces).  Here the MD engine links libxeq_b200.so and calls xeq_model_energy_forces() once per step with its own
// neighbour list: the forward pass is the same sequence of C-ABI kernels the nn modules issue
// (xequinet_b200/nn/xpainn.py), and the force pass is the hand-scheduled reverse sweep that torch.autograd.grad(E, pos)
// performs over those modules (nn/basic.py:143-159) -- same kernels, same arguments, same summation order, so the
// results are bit-identical to the Python path (tests/test_gpu_runtime.py).
//
// Scope: the default model (nn/model.py:57-70 with layer_norm=True, silu, output_modes=["energy"], no charge / spin
// conditioning).  Weights arrive as one flat fp32 device blob in the order of `layout()` below
// (xequinet_b200/runtime.py::export_weights writes it from a state_dict); the caller owns it and the workspace.
#include <math.h>
#include <string.h>

#include <new>

#include "common.cuh"

struct xeq_model {
  xeq_dims_t dims;
  int32_t n_layers, hidden_dim, embed_dim, n_species;
  const float* w;  // device blob (caller-owned)
};

namespace xeq {
namespace {

constexpr int MAX_LAYERS = 8;
constexpr float NORM_EPS = 1e-5f;  // nn.LayerNorm / EquivariantLayerNorm default (nn/o3layer.py:118)

struct MsgW { size_t ln_w, ln_b, on_w, on_b, W1, b1, W2, b2, Wrbf, brbf; };
struct UpdW { size_t ln_w, ln_b, on_w, on_b, Uw, Ub, Vw, Vb, dotW, M1, m1b, M2, m2b; };
struct Layout {
  size_t table, embW, embb, freq;
  MsgW msg[MAX_LAYERS];
  UpdW upd[MAX_LAYERS];
  size_t O1, o1b, O2, o2b;
  size_t total;
};

struct Sizes {
  int C, m0, m1, m2, M, D, H, Hu, B, hid, E;
};
Sizes sizes_of(const xeq_dims_t& d, int hidden, int embed) {
  Sizes s;
  s.C = d.node_dim; s.m0 = d.mul0; s.m1 = d.mul1; s.m2 = d.mul2;
  s.M = s.m0 + s.m1 + s.m2;
  s.D = s.m0 + 3 * s.m1 + 5 * s.m2;
  s.H = s.C + 2 * s.M;   // nn/xpainn.py:108
  s.Hu = 2 * s.C + s.M;  // nn/xpainn.py:184
  s.B = d.num_basis; s.hid = hidden; s.E = embed;
  return s;
}

// every tensor starts on a 16-byte boundary (the GEMM kernel's operand granularity)
Layout layout(const xeq_dims_t& d, int L, int hidden, int embed, int n_species) {
  const Sizes s = sizes_of(d, hidden, embed);
  Layout lo;
  size_t off = 0;
  auto put = [&](size_t n) { const size_t at = off; off += (n + 3) / 4 * 4; return at; };
  const size_t nw = (size_t)s.m0 * s.m0 + (size_t)s.m1 * s.m1 + (size_t)s.m2 * s.m2;
  lo.table = put((size_t)n_species * s.E);
  lo.embW = put((size_t)s.C * s.E);
  lo.embb = put(s.C);
  lo.freq = put(s.B);
  for (int l = 0; l < L; ++l) {
    MsgW& m = lo.msg[l];
    m.ln_w = put(s.C); m.ln_b = put(s.C); m.on_w = put(s.M); m.on_b = put(s.m0);
    m.W1 = put((size_t)s.C * s.C); m.b1 = put(s.C); m.W2 = put((size_t)s.H * s.C); m.b2 = put(s.H);
    m.Wrbf = put((size_t)s.H * s.B); m.brbf = put(s.H);
    UpdW& u = lo.upd[l];
    u.ln_w = put(s.C); u.ln_b = put(s.C); u.on_w = put(s.M); u.on_b = put(s.m0);
    u.Uw = put(nw); u.Ub = put(s.m0); u.Vw = put(nw); u.Vb = put(s.m0);
    u.dotW = put((size_t)s.C * s.M);
    u.M1 = put((size_t)s.C * (s.C + s.M)); u.m1b = put(s.C);
    u.M2 = put((size_t)s.Hu * s.C); u.m2b = put(s.Hu);
  }
  lo.O1 = put((size_t)s.hid * s.C); lo.o1b = put(s.hid); lo.O2 = put(s.hid); lo.o2b = put(1);
  lo.total = off;
  return lo;
}

struct Buffers {
  float* emb_in;
  float *x[2 * MAX_LAYERS + 1], *V[2 * MAX_LAYERS + 1];
  float *u1[MAX_LAYERS], *s[MAX_LAYERS], *vn[MAX_LAYERS];
  float *U[MAX_LAYERS], *Wt[MAX_LAYERS], *u2[MAX_LAYERS], *a[MAX_LAYERS], *t[MAX_LAYERS];
  float *xn, *h, *vn2, *cat, *t0, *uo, *ho;
  // force pass
  float *gx[2], *gV[2], *ga, *gU, *gt, *g_t0, *g_h, *g_u, *g_cat, *gU2, *gWt, *gvnA, *gvnB, *gvn, *gs, *gv, *ones;
  float* gpos[MAX_LAYERS];
  float* crow[MAX_LAYERS];  // [9, N] per layer: per-node pieces of the cell gradient (periodic graphs)
  void* edge_ws;
  size_t edge_ws_bytes;
};

size_t carve(const xeq_model& mdl, const xeq_graph_t& g, void* base, bool with_forces, Buffers& b) {
  const Sizes s = sizes_of(mdl.dims, mdl.hidden_dim, mdl.embed_dim);
  const size_t N = (size_t)(g.n_nodes > 0 ? g.n_nodes : 1);
  const int L = mdl.n_layers;
  Carver cv(base);
  b.emb_in = cv.take<float>(N * s.E);
  for (int i = 0; i <= 2 * L; ++i) {
    b.x[i] = cv.take<float>(N * s.C);
    b.V[i] = cv.take<float>(N * s.D);
  }
  for (int l = 0; l < L; ++l) {
    b.u1[l] = cv.take<float>(N * s.C); b.s[l] = cv.take<float>(N * s.H); b.vn[l] = cv.take<float>(N * s.D);
    b.U[l] = cv.take<float>(N * s.D); b.Wt[l] = cv.take<float>(N * s.D); b.u2[l] = cv.take<float>(N * s.C);
    b.a[l] = cv.take<float>(N * s.Hu); b.t[l] = cv.take<float>(N * s.C);
  }
  b.xn = cv.take<float>(N * s.C); b.h = cv.take<float>(N * s.C); b.vn2 = cv.take<float>(N * s.D);
  b.cat = cv.take<float>(N * (s.C + s.M)); b.t0 = cv.take<float>(N * s.M);
  b.uo = cv.take<float>(N * s.hid); b.ho = cv.take<float>(N * s.hid);
  if (with_forces) {
    for (int i = 0; i < 2; ++i) { b.gx[i] = cv.take<float>(N * s.C); b.gV[i] = cv.take<float>(N * s.D); }
    b.ga = cv.take<float>(N * s.Hu); b.gU = cv.take<float>(N * s.D); b.gt = cv.take<float>(N * s.C);
    b.g_t0 = cv.take<float>(N * s.M);
    b.g_h = cv.take<float>(N * (size_t)(s.C > s.hid ? s.C : s.hid));
    b.g_u = cv.take<float>(N * (size_t)(s.C > s.hid ? s.C : s.hid));
    b.g_cat = cv.take<float>(N * (s.C + s.M));
    b.gU2 = cv.take<float>(N * s.D); b.gWt = cv.take<float>(N * s.D);
    b.gvnA = cv.take<float>(N * s.D); b.gvnB = cv.take<float>(N * s.D); b.gvn = cv.take<float>(N * s.D);
    b.gs = cv.take<float>(N * s.H); b.gv = cv.take<float>(N * s.D); b.ones = cv.take<float>(N);
    for (int l = 0; l < L; ++l) b.gpos[l] = cv.take<float>(N * 3);
    for (int l = 0; l < L; ++l) b.crow[l] = g.offsets ? cv.take<float>(N * 9) : nullptr;
  }
  size_t ews = xeq_edge_message_fwd_workspace_bytes(&g, &mdl.dims);
  if (with_forces) {
    const size_t bws = xeq_edge_message_bwd_workspace_bytes(&g, &mdl.dims, 0);
    if (bws > ews) ews = bws;
  }
  b.edge_ws = cv.take<char>(ews);
  b.edge_ws_bytes = ews;
  return align_up(cv.off, 256);
}

// ---- the three elementwise helpers the sweep needs next to the C-ABI kernels ----
__global__ void gather_rows_kernel(const float* __restrict__ table, const int* __restrict__ idx, int n_rows, int width,
                                   int n_table, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_rows * width) return;
  const int r = (int)(i / width), c = (int)(i % width);
  int z = idx[r];
  z = z < 0 ? 0 : (z >= n_table ? n_table - 1 : z);
  out[i] = table[(size_t)z * width + c];
}
// d sum_g E_g / d atomic_energies[i]: 1 for the nodes inside the segments, 0 for nodes outside [seg_ptr[0], seg_ptr[G])
// (an MD engine's ghost atoms: neighbours that receive forces but contribute no energy)
__global__ void energy_seed_kernel(float* __restrict__ p, const int* __restrict__ seg_ptr, int n_seg, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (n_seg > 0 && (long long)i >= seg_ptr[0] && (long long)i < seg_ptr[n_seg]) ? 1.0f : 0.0f;
}
__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}
// forces = -((g[L-1] + g[L-2]) + ... + g[0]): the order in which autograd sums the layers' position gradients
struct PosGrads { const float* g[MAX_LAYERS]; int n; };
__global__ void forces_kernel(PosGrads pg, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = pg.g[pg.n - 1][i];
  for (int l = pg.n - 2; l >= 0; --l) acc += pg.g[l][i];
  out[i] = -1.0f * acc;
}
// virial[g] = -dE/dstrain of the reference's strain trick (nn/basic.py:93-107, 162-199): positions and cell displaced by a
// symmetrised strain, S[a][b] = sum_i pos_i[a] dE/dpos_i[b] + sum_c cell[c][a] dE/dcell[c][b], virial = -(S + S^T) / 2,
// with dE/dpos = -forces and dE/dcell[c][b] = -sum_n sum_layers rows[3c+b][n] (xeq_edge_cell_grad_rows).
// One CTA per graph; fixed-order reductions (deterministic).
struct CellRows { const float* r[MAX_LAYERS]; int n; };
__global__ void __launch_bounds__(256) virial_kernel(const float* __restrict__ pos, const float* __restrict__ forces, CellRows cr,
                                                     const float* __restrict__ cell, const int* __restrict__ seg_ptr, int n_nodes,
                                                     float* __restrict__ virial) {
  __shared__ float red[8][18];
  const int gidx = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n0 = seg_ptr[gidx], n1 = seg_ptr[gidx + 1];
  float acc[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) acc[k] = 0.f;
  for (int i = n0 + (int)threadIdx.x; i < n1; i += 256) {
    const float p[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    const float gp[3] = {-forces[3 * i], -forces[3 * i + 1], -forces[3 * i + 2]};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) acc[3 * a + bb] = fmaf(p[a], gp[bb], acc[3 * a + bb]);
    if (cell) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float v = 0.f;
        for (int l = cr.n - 1; l >= 0; --l) v += cr.r[l][(size_t)k * n_nodes + i];
        acc[9 + k] += v;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 18; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0) red[warp][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float S[9], R[9];
    for (int k = 0; k < 9; ++k) {
      float a = 0.f, r = 0.f;
      for (int w = 0; w < 8; ++w) { a += red[w][k]; r += red[w][9 + k]; }
      S[k] = a;
      R[k] = -r;  // dE/dcell[c][b]
    }
    if (cell) {
      const float* cg = cell + 9 * (size_t)gidx;
      for (int a = 0; a < 3; ++a)
        for (int bb = 0; bb < 3; ++bb)
          for (int c = 0; c < 3; ++c) S[3 * a + bb] = fmaf(cg[3 * c + a], R[3 * c + bb], S[3 * a + bb]);
    }
    for (int a = 0; a < 3; ++a)
      for (int bb = 0; bb < 3; ++bb) virial[9 * (size_t)gidx + 3 * a + bb] = -0.5f * (S[3 * a + bb] + S[3 * bb + a]);
  }
}
inline unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }

struct EventPair {  // released on every return path; destroying an event with pending work is allowed
  cudaEvent_t f = nullptr, j = nullptr;
  ~EventPair() {
    if (f) cudaEventDestroy(f);
    if (j) cudaEventDestroy(j);
  }
};

#define XEQ_TRY(call)        \
  do {                       \
    const int _rc = (call);  \
    if (_rc) return _rc;     \
  } while (0)

// C = alpha * op(A) op(B) (+ bias), one problem
int gemm(const float* a, int lda, const float* b, int ldb, bool tb, const float* bias, float* c, int m, int n, int k,
         cudaStream_t st) {
  xeq_gemm_t p;
  p.a = a; p.b = b; p.bias = bias; p.c = c;
  p.m = m; p.n = n; p.k = k; p.lda = lda; p.ldb = ldb; p.ldc = n;
  p.a_trans = 0; p.b_trans = tb ? 1 : 0; p.alpha = 1.0f; p.act = 0;
  return xeq_gemm_tf32x3(&p, 1, 1, nullptr, 0, st);
}
// nn.Linear forward: y = x W^T + b, W [out, in]
int linear(const float* x, int ldx, const float* W, const float* bias, float* y, int n_rows, int out_f, int in_f, cudaStream_t st) {
  return gemm(x, ldx, W, in_f, true, bias, y, n_rows, out_f, in_f, st);
}
// its input gradient: gx = g W
int linear_bwd(const float* g, int ldg, const float* W, float* gx, int n_rows, int out_f, int in_f, cudaStream_t st) {
  return gemm(g, ldg, W, in_f, false, nullptr, gx, n_rows, in_f, out_f, st);
}
// e3nn o3.Linear on the cm layout (gemm.irreps_linear_raw): one problem per (l, m) block
int irreps_linear(const float* V, const float* w, const float* bias, float* out, int n_rows, const Sizes& s, bool transposed,
                  cudaStream_t st) {
  xeq_gemm_t p[9];
  int np = 0;
  const int muls[3] = {s.m0, s.m1, s.m2};
  size_t foff = 0, woff = 0;
  for (int l = 0; l < 3; ++l) {
    const int mul = muls[l];
    if (mul) {
      for (int m = 0; m < 2 * l + 1; ++m) {
        const size_t off = foff + (size_t)m * mul;
        xeq_gemm_t& q = p[np++];
        q.a = V + off; q.b = w + woff; q.c = out + off;
        q.bias = (l == 0) ? bias : nullptr;
        q.m = n_rows; q.n = mul; q.k = mul; q.lda = s.D; q.ldb = mul; q.ldc = s.D;
        q.a_trans = 0; q.b_trans = transposed ? 1 : 0;
        q.alpha = 1.0f / sqrtf((float)mul); q.act = 0;
      }
    }
    foff += (size_t)(2 * l + 1) * mul;
    woff += (size_t)mul * mul;
  }
  return xeq_gemm_tf32x3(p, np, 1, nullptr, 0, st);
}

int norm_bwd_x(const float* x, const float* gamma, const float* g, int ld_g, const float* gx_add, int n, int m0, int m1, int m2,
               float* gx, cudaStream_t st) {
  return xeq_irreps_norm_bwd(x, gamma, g, ld_g, gx_add, n, m0, m1, m2, NORM_EPS, gx, nullptr, nullptr, nullptr, 0, st);
}

}  // namespace
}  // namespace xeq

using namespace xeq;

extern "C" {

// ---- one-call construction of an xeq_graph_t from an engine's edge list ----
namespace {
struct GraphParts {
  int32_t *rowptr, *col, *t_rowptr, *t_row, *t_eid, *tile_ptr, *t_tile_ptr;
  int8_t* offsets;
  void* t_ws;
  size_t t_ws_bytes;
  int n_tiles, t_n_tiles;
};
size_t carve_graph(int32_t N, int32_t E, bool periodic, void* base, GraphParts& p) {
  const size_t e = (size_t)(E > 0 ? E : 1);
  p.n_tiles = xeq_csr_tile_count(N, E, xeq_center_tile_edges());
  p.t_n_tiles = xeq_csr_tile_count(N, E, xeq_neighbor_tile_edges());
  Carver cv(base);
  p.rowptr = cv.take<int32_t>((size_t)N + 1);
  p.col = cv.take<int32_t>(e);
  p.offsets = periodic ? cv.take<int8_t>(4 * e) : nullptr;
  p.t_rowptr = cv.take<int32_t>((size_t)N + 1);
  p.t_row = cv.take<int32_t>(e);
  p.t_eid = cv.take<int32_t>(e);
  p.tile_ptr = cv.take<int32_t>((size_t)p.n_tiles + 1);
  p.t_tile_ptr = cv.take<int32_t>((size_t)p.t_n_tiles + 1);
  p.t_ws_bytes = xeq_csr_transpose_workspace_bytes(N, E);
  p.t_ws = cv.take<char>(p.t_ws_bytes);
  return align_up(cv.off, 256);
}
}  // namespace

size_t xeq_graph_from_coo_bytes(int32_t n_nodes, int32_t n_edges, int periodic) {
  if (n_nodes < 0 || n_edges < 0) return 0;
  GraphParts p;
  return carve_graph(n_nodes, n_edges, periodic != 0, nullptr, p);
}

int xeq_graph_from_coo(const int64_t* edge_index, const float* cell_offsets, const float* cell, const int32_t* node_graph,
                       int32_t n_nodes, int32_t n_edges, int32_t n_graphs, void* storage, size_t storage_bytes,
                       xeq_graph_t* graph_host, xeq_stream_t stream) {
  XEQ_CHECK_ARG(graph_host && storage && n_nodes >= 0 && n_edges >= 0 && n_graphs >= 1, "graph_from_coo: bad arguments");
  XEQ_CHECK_ARG(edge_index || n_edges == 0, "graph_from_coo: edge_index is NULL");
  XEQ_CHECK_ARG((cell == nullptr) == (cell_offsets == nullptr), "graph_from_coo: cell and cell_offsets go together");
  XEQ_CHECK_ARG(!(cell && n_graphs > 1) || node_graph, "graph_from_coo: node_graph is needed for several periodic graphs");
  XEQ_CHECK_ARG(((uintptr_t)storage & 255) == 0, "graph_from_coo: storage must be 256-byte aligned");
  const bool periodic = cell != nullptr;
  GraphParts p;
  const size_t need = carve_graph(n_nodes, n_edges, periodic, storage, p);
  if (storage_bytes < need) {
    set_error("graph_from_coo: storage too small (%zu < %zu bytes)", storage_bytes, need);
    return XEQ_ERR_WORKSPACE;
  }
  XEQ_TRY(xeq_csr_from_sorted_coo(edge_index, cell_offsets, n_nodes, n_edges, p.rowptr, p.col, p.offsets, stream));
  XEQ_TRY(xeq_csr_transpose(p.rowptr, p.col, n_nodes, n_edges, p.t_rowptr, p.t_row, p.t_eid, p.t_ws, p.t_ws_bytes, stream));
  XEQ_TRY(xeq_csr_tile_bounds(p.rowptr, n_nodes, n_edges, xeq_center_tile_edges(), p.tile_ptr, stream));
  XEQ_TRY(xeq_csr_tile_bounds(p.t_rowptr, n_nodes, n_edges, xeq_neighbor_tile_edges(), p.t_tile_ptr, stream));
  xeq_graph_t g;
  memset(&g, 0, sizeof(g));
  g.n_nodes = n_nodes; g.n_edges = n_edges; g.n_graphs = n_graphs;
  g.rowptr = p.rowptr; g.col = p.col; g.t_rowptr = p.t_rowptr; g.t_row = p.t_row; g.t_eid = p.t_eid;
  g.offsets = p.offsets; g.cell = cell; g.node_graph = (periodic && n_graphs > 1) ? node_graph : nullptr;
  g.tile_ptr = p.tile_ptr; g.t_tile_ptr = p.t_tile_ptr; g.n_tiles = p.n_tiles; g.t_n_tiles = p.t_n_tiles;
  g.tile_mode = 0; g.max_tile_nodes = 0;
  *graph_host = g;
  return XEQ_OK;
}

size_t xeq_model_weight_count(const xeq_dims_t* dims, int32_t n_layers, int32_t hidden_dim, int32_t embed_dim, int32_t n_species) {
  if (!dims || n_layers < 1 || n_layers > MAX_LAYERS || hidden_dim < 4 || embed_dim < 4 || n_species < 1) return 0;
  return layout(*dims, n_layers, hidden_dim, embed_dim, n_species).total;
}

int xeq_model_create(const xeq_dims_t* dims, int32_t n_layers, int32_t hidden_dim, int32_t embed_dim, int32_t n_species,
                     const float* weights, size_t n_weights, xeq_model_t** model) {
  XEQ_CHECK_ARG(dims && model && weights, "model_create: NULL argument");
  XEQ_CHECK_ARG(n_layers >= 1 && n_layers <= MAX_LAYERS, "model_create: 1..%d action blocks", MAX_LAYERS);
  XEQ_CHECK_ARG(dims->node_dim == dims->mul0 && dims->mul0 % 32 == 0 && dims->mul1 % 32 == 0 && dims->mul2 % 32 == 0,
                "model_create: node_dim == mul0 and multiplicities that are multiples of 32 (xeq_dims_t)");
  XEQ_CHECK_ARG(hidden_dim % 4 == 0 && embed_dim % 4 == 0 && hidden_dim >= 4 && embed_dim >= 4 && n_species >= 1,
                "model_create: hidden_dim and embed_dim must be multiples of 4");
  XEQ_CHECK_ARG(((uintptr_t)weights & 15) == 0, "model_create: the weight blob must be 16-byte aligned");
  const size_t want = layout(*dims, n_layers, hidden_dim, embed_dim, n_species).total;
  XEQ_CHECK_ARG(n_weights == want, "model_create: weight blob has %zu floats, the layout needs %zu", n_weights, want);
  xeq_model* m = new (std::nothrow) xeq_model;
  XEQ_CHECK_ARG(m, "model_create: out of host memory");
  m->dims = *dims; m->n_layers = n_layers; m->hidden_dim = hidden_dim; m->embed_dim = embed_dim; m->n_species = n_species;
  m->w = weights;
  *model = m;
  return XEQ_OK;
}

void xeq_model_destroy(xeq_model_t* model) { delete model; }

size_t xeq_model_workspace_bytes(const xeq_model_t* model, const xeq_graph_t* g, int want_forces) {
  if (!model || !g) return 0;
  Buffers b;
  return carve(*model, *g, nullptr, want_forces != 0, b);
}

int xeq_model_energy_forces(const xeq_model_t* model, const xeq_graph_t* g, const float* pos, const int32_t* atomic_numbers,
                            const int32_t* seg_ptr, float* energy, float* atomic_energies, float* forces,
                            void* workspace, size_t workspace_bytes, xeq_stream_t stream) {
  return xeq_model_energy_forces_mt(model, g, pos, atomic_numbers, seg_ptr, energy, atomic_energies, forces, workspace,
                                    workspace_bytes, stream, nullptr);
}

int xeq_model_energy_forces_mt(const xeq_model_t* model, const xeq_graph_t* g, const float* pos, const int32_t* atomic_numbers,
                               const int32_t* seg_ptr, float* energy, float* atomic_energies, float* forces,
                               void* workspace, size_t workspace_bytes, xeq_stream_t stream, xeq_stream_t aux_stream) {
  return xeq_model_energy_forces_virial(model, g, pos, atomic_numbers, seg_ptr, energy, atomic_energies, forces, nullptr,
                                        workspace, workspace_bytes, stream, aux_stream);
}

int xeq_model_energy_forces_virial(const xeq_model_t* model, const xeq_graph_t* g, const float* pos, const int32_t* atomic_numbers,
                                   const int32_t* seg_ptr, float* energy, float* atomic_energies, float* forces, float* virial,
                                   void* workspace, size_t workspace_bytes, xeq_stream_t stream, xeq_stream_t aux_stream) {
  XEQ_CHECK_ARG(model && g && seg_ptr && energy && atomic_energies, "model_energy_forces: NULL argument");
  XEQ_CHECK_ARG(!virial || forces, "model_energy_forces: the virial needs the force pass (forces != NULL)");
  const int N = g->n_nodes, G = g->n_graphs, L = model->n_layers;
  XEQ_CHECK_ARG(N >= 0 && G >= 0, "model_energy_forces: bad graph");
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    if (G) XEQ_CUDA(cudaMemsetAsync(energy, 0, sizeof(float) * G, st));
    if (G && virial) XEQ_CUDA(cudaMemsetAsync(virial, 0, sizeof(float) * 9 * G, st));
    return XEQ_OK;
  }
  XEQ_CHECK_ARG(pos && atomic_numbers && workspace, "model_energy_forces: NULL argument");
  const bool wf = forces != nullptr;
  Buffers b;
  const size_t need = carve(*model, *g, workspace, wf, b);
  if (workspace_bytes < need) {
    set_error("model_energy_forces: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return XEQ_ERR_WORKSPACE;
  }
  XEQ_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "model_energy_forces: the workspace must be 256-byte aligned");
  const xeq_dims_t* dims = &model->dims;
  const Sizes s = sizes_of(*dims, model->hidden_dim, model->embed_dim);
  const Layout lo = layout(*dims, L, model->hidden_dim, model->embed_dim, model->n_species);
  const float* w = model->w;
  const int C = s.C, D = s.D, M = s.M, H = s.H, Hu = s.Hu, CM = s.C + s.M;

  // Independent branches of the module graph (norm(x) -> scalar MLP beside o3norm(V); update_U beside update_V; dot_lin
  // beside the update MLP; and their mirror images in the force pass) run on `aux_stream` when the caller gives one:
  // at MD sizes every kernel is a fraction of a wave, so the step is a latency chain and the branches overlap.  Same
  // kernels and arguments either way: the results do not depend on it.  fork: aux waits for everything issued on the
  // main stream so far; join: the main stream waits for the branch.  Plain event record / wait pairs -- capturable.
  cudaStream_t sb = aux_stream ? (cudaStream_t)aux_stream : st;
  EventPair ev;
  if (aux_stream && sb != st) {
    XEQ_CUDA(cudaEventCreateWithFlags(&ev.f, cudaEventDisableTiming));
    XEQ_CUDA(cudaEventCreateWithFlags(&ev.j, cudaEventDisableTiming));
  } else {
    sb = st;
  }
  auto fork = [&]() -> int {
    if (sb == st) return XEQ_OK;
    XEQ_CUDA(cudaEventRecord(ev.f, st));
    XEQ_CUDA(cudaStreamWaitEvent(sb, ev.f, 0));
    return XEQ_OK;
  };
  auto join = [&]() -> int {
    if (sb == st) return XEQ_OK;
    XEQ_CUDA(cudaEventRecord(ev.j, sb));
    XEQ_CUDA(cudaStreamWaitEvent(st, ev.j, 0));
    return XEQ_OK;
  };

  // ---- XEmbedding (nn/xpainn.py:55-83): x0 = Linear(embed_ten[Z]); V0 = 0 ----
  gather_rows_kernel<<<blocks_for((size_t)N * s.E), 256, 0, st>>>(w + lo.table, atomic_numbers, N, s.E, model->n_species, b.emb_in);
  XEQ_LAUNCHED(1);
  XEQ_TRY(linear(b.emb_in, s.E, w + lo.embW, w + lo.embb, b.x[0], N, C, s.E, st));
  XEQ_CUDA(cudaMemsetAsync(b.V[0], 0, sizeof(float) * (size_t)N * D, st));

  for (int l = 0; l < L; ++l) {
    // ---- XPainnMessage.forward (nn/xpainn.py:128-161) ----
    const MsgW& mw = lo.msg[l];
    const float *x_in = b.x[2 * l], *V_in = b.V[2 * l];
    XEQ_TRY(fork());
    XEQ_TRY(xeq_irreps_norm_fwd(V_in, w + mw.on_w, w + mw.on_b, N, s.m0, s.m1, s.m2, NORM_EPS, b.vn[l], sb));
    XEQ_TRY(xeq_irreps_norm_fwd(x_in, w + mw.ln_w, w + mw.ln_b, N, C, 0, 0, NORM_EPS, b.xn, st));
    XEQ_TRY(linear(b.xn, C, w + mw.W1, w + mw.b1, b.u1[l], N, C, C, st));
    XEQ_TRY(xeq_silu_fwd(b.u1[l], (size_t)N * C, b.h, st));
    XEQ_TRY(linear(b.h, C, w + mw.W2, w + mw.b2, b.s[l], N, H, C, st));
    XEQ_TRY(join());
    XEQ_TRY(xeq_edge_message_fwd(g, dims, pos, b.s[l], b.vn[l], x_in, V_in, w + mw.Wrbf, w + mw.brbf, w + lo.freq,
                                 b.x[2 * l + 1], b.V[2 * l + 1], b.edge_ws, b.edge_ws_bytes, st));
    // ---- XPainnUpdate.forward (nn/xpainn.py:206-231) ----
    const UpdW& uw = lo.upd[l];
    const float *x1 = b.x[2 * l + 1], *V1 = b.V[2 * l + 1];
    XEQ_TRY(fork());
    XEQ_TRY(xeq_irreps_norm_fwd(V1, w + uw.on_w, w + uw.on_b, N, s.m0, s.m1, s.m2, NORM_EPS, b.vn2, sb));
    XEQ_TRY(join());  // the main stream sees o3norm(V); the branch goes on with update_U
    XEQ_TRY(irreps_linear(b.vn2, w + uw.Uw, w + uw.Ub, b.U[l], N, s, false, sb));
    XEQ_TRY(xeq_irreps_norm_fwd(x1, w + uw.ln_w, w + uw.ln_b, N, C, 0, 0, NORM_EPS, b.xn, st));
    XEQ_CUDA(cudaMemcpy2DAsync(b.cat, sizeof(float) * CM, b.xn, sizeof(float) * C, sizeof(float) * C, N, cudaMemcpyDeviceToDevice, st));
    XEQ_TRY(irreps_linear(b.vn2, w + uw.Vw, w + uw.Vb, b.Wt[l], N, s, false, st));
    XEQ_TRY(join());
    XEQ_TRY(xeq_invariant_dot_fwd(b.U[l], b.Wt[l], N, s.m0, s.m1, s.m2, b.cat + C, CM, b.t0, st));  // [xn | Invariant(W)]
    XEQ_TRY(fork());
    XEQ_TRY(linear(b.t0, M, w + uw.dotW, nullptr, b.t[l], N, C, M, sb));
    XEQ_TRY(linear(b.cat, CM, w + uw.M1, w + uw.m1b, b.u2[l], N, C, CM, st));
    XEQ_TRY(xeq_silu_fwd(b.u2[l], (size_t)N * C, b.h, st));
    XEQ_TRY(linear(b.h, C, w + uw.M2, w + uw.m2b, b.a[l], N, Hu, C, st));
    XEQ_TRY(join());
    XEQ_TRY(xeq_gate_residual_fwd(b.a[l], b.U[l], b.t[l], x1, V1, N, s.m0, s.m1, s.m2, b.x[2 * l + 2], b.V[2 * l + 2], st));
  }

  // ---- EnergyOut.forward (nn/output.py:114-128) ----
  XEQ_TRY(linear(b.x[2 * L], C, w + lo.O1, w + lo.o1b, b.uo, N, s.hid, C, st));
  XEQ_TRY(xeq_silu_fwd(b.uo, (size_t)N * s.hid, b.ho, st));
  XEQ_TRY(xeq_rowdot(b.ho, s.hid, w + lo.O2, w + lo.o2b, N, s.hid, atomic_energies, st));
  XEQ_TRY(xeq_segment_sum(atomic_energies, seg_ptr, G, energy, st));
  if (!wf) return XEQ_OK;

  // ---- forces = -dE/dpos (nn/basic.py:143-159): reverse sweep over the modules above, d/dpos only ----
  int cur = 0;
  energy_seed_kernel<<<blocks_for((size_t)N), 256, 0, st>>>(b.ones, seg_ptr, G, (size_t)N);  // d sum(E) / d atomic_energies
  XEQ_LAUNCHED(1);
  XEQ_TRY(xeq_outer(b.ones, w + lo.O2, N, s.hid, b.g_h, st));
  XEQ_TRY(xeq_silu_bwd(b.uo, b.g_h, (size_t)N * s.hid, b.g_u, st));
  XEQ_TRY(linear_bwd(b.g_u, s.hid, w + lo.O1, b.gx[cur], N, s.hid, C, st));
  XEQ_CUDA(cudaMemsetAsync(b.gV[cur], 0, sizeof(float) * (size_t)N * D, st));  // the energy does not read the last V

  for (int l = L - 1; l >= 0; --l) {
    const UpdW& uw = lo.upd[l];
    const MsgW& mw = lo.msg[l];
    // update block
    XEQ_TRY(xeq_gate_residual_bwd(b.a[l], b.U[l], b.t[l], b.gx[cur], b.gV[cur], N, s.m0, s.m1, s.m2, b.ga, b.gU, b.gt, st));
    XEQ_TRY(fork());
    XEQ_TRY(linear_bwd(b.gt, C, w + uw.dotW, b.g_t0, N, C, M, sb));
    XEQ_TRY(linear_bwd(b.ga, Hu, w + uw.M2, b.g_h, N, Hu, C, st));
    XEQ_TRY(xeq_silu_bwd(b.u2[l], b.g_h, (size_t)N * C, b.g_u, st));
    XEQ_TRY(linear_bwd(b.g_u, C, w + uw.M1, b.g_cat, N, C, CM, st));
    XEQ_TRY(join());
    XEQ_TRY(xeq_invariant_dot_bwd(b.U[l], b.Wt[l], b.g_cat + C, CM, b.g_t0, b.gU, N, s.m0, s.m1, s.m2, b.gU2, b.gWt, st));
    XEQ_TRY(fork());
    XEQ_TRY(irreps_linear(b.gWt, w + uw.Vw, nullptr, b.gvnB, N, s, true, sb));
    XEQ_TRY(norm_bwd_x(b.x[2 * l + 1], w + uw.ln_w, b.g_cat, CM, b.gx[cur], N, C, 0, 0, b.gx[cur ^ 1], sb));
    XEQ_TRY(irreps_linear(b.gU2, w + uw.Uw, nullptr, b.gvnA, N, s, true, st));
    XEQ_TRY(join());
    add2_kernel<<<blocks_for((size_t)N * D), 256, 0, st>>>(b.gvnA, b.gvnB, b.gvn, (size_t)N * D);
    XEQ_LAUNCHED(1);
    XEQ_TRY(norm_bwd_x(b.V[2 * l + 1], w + uw.on_w, b.gvn, 0, b.gV[cur], N, s.m0, s.m1, s.m2, b.gV[cur ^ 1], st));
    cur ^= 1;
    // message block; the first layer's inputs do not depend on the positions
    const bool first = (l == 0);
    XEQ_TRY(xeq_edge_message_bwd(g, dims, pos, b.s[l], b.vn[l], w + mw.Wrbf, w + mw.brbf, w + lo.freq, b.gx[cur], b.gV[cur],
                                 first ? nullptr : b.gs, first ? nullptr : b.gv, b.gpos[l], nullptr, nullptr, nullptr,
                                 b.edge_ws, b.edge_ws_bytes, st));
    if (virial && g->offsets) XEQ_TRY(xeq_edge_cell_grad_rows(g, dims, b.edge_ws, b.crow[l], st));  // before the workspace is reused
    if (!first) {
      XEQ_TRY(fork());
      XEQ_TRY(norm_bwd_x(b.V[2 * l], w + mw.on_w, b.gv, 0, b.gV[cur], N, s.m0, s.m1, s.m2, b.gV[cur ^ 1], sb));
      XEQ_TRY(linear_bwd(b.gs, H, w + mw.W2, b.g_h, N, H, C, st));
      XEQ_TRY(xeq_silu_bwd(b.u1[l], b.g_h, (size_t)N * C, b.g_u, st));
      XEQ_TRY(linear_bwd(b.g_u, C, w + mw.W1, b.g_cat, N, C, C, st));  // g_cat reused as the [N, C] gradient of norm(x)
      XEQ_TRY(norm_bwd_x(b.x[2 * l], w + mw.ln_w, b.g_cat, 0, b.gx[cur], N, C, 0, 0, b.gx[cur ^ 1], st));
      XEQ_TRY(join());
      cur ^= 1;
    }
  }
  PosGrads pg;
  pg.n = L;
  for (int l = 0; l < L; ++l) pg.g[l] = b.gpos[l];
  forces_kernel<<<blocks_for((size_t)N * 3), 256, 0, st>>>(pg, forces, (size_t)N * 3);
  XEQ_LAUNCHED(1);
  if (virial && G > 0) {
    CellRows cr;
    cr.n = L;
    for (int l = 0; l < L; ++l) cr.r[l] = b.crow[l];
    virial_kernel<<<G, 256, 0, st>>>(pos, forces, cr, g->offsets ? g->cell : nullptr, seg_ptr, N, virial);
    XEQ_LAUNCHED(1);
  }
  return XEQ_OK;
}

}  // extern "C"
