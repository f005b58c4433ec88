/*
 * xeq_b200.h -- C ABI of libxeq_b200.so (sm_100a only).
 *
 * Drop-in boundary for the XPaiNN message-passing hot path of X1X1010/XequiNet.  The
 * reference has no FFI layer; its boundary is a set of Python calls into third-party
 * packages.  Each entry point below names the reference call site
 * (paths relative to /root/reference/xequinet/) whose device work it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     parameter name ends in _host;  all feature tensors are row-major fp32;
 *   - the caller owns every buffer (outputs and workspace); a *_workspace_bytes()
 *     query precedes each call that needs scratch;
 *   - every call is asynchronous on `stream` (a cudaStream_t), performs no allocation
 *     and no host synchronisation, and keeps no global mutable state;
 *   - return value 0 = success, negative = error (xeq_last_error() gives a thread-local
 *     message).  No C++ exceptions cross the ABI.  There is no CPU fallback.
 *
 * Internal feature layout ("cm", component-major) of equivariant tensors [N, D]:
 *     [ mul0 scalars | l=1: 3 blocks (m=-1,0,1) of mul1 | l=2: 5 blocks (m=-2..2) of mul2 ]
 * i.e. within each l the multiplicity index is fastest, so a warp that owns 32
 * consecutive channels reads 128 contiguous bytes per component.  The reference (e3nn)
 * layout is mul-major / m-fastest; xeq_layout_convert() maps between the two.
 */
#ifndef XEQ_B200_H
#define XEQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XEQ_OK 0
#define XEQ_ERR_INVALID -1     /* bad argument / unsupported shape */
#define XEQ_ERR_CUDA -2        /* a CUDA runtime call failed */
#define XEQ_ERR_WORKSPACE -3   /* workspace too small */

typedef void* xeq_stream_t; /* cudaStream_t */

/* Model widths: nn/model.py:57-67.  mul0 must equal node_dim (true for every documented
 * config, docs/config.md:64-65); mul* must be multiples of 32; num_basis <= 23. */
typedef struct {
  int32_t node_dim; /* C   */
  int32_t mul0;     /* number of 0e irreps */
  int32_t mul1;     /* number of 1o irreps */
  int32_t mul2;     /* number of 2e irreps */
  int32_t num_basis;/* B   */
  float cutoff;     /* r_c */
} xeq_dims_t;

/* Neighbour structure produced by xeq_radius_graph_* / xeq_csr_from_coo + xeq_csr_transpose.
 * Edge e = (center <- neighbor) (keys.py:16-17); canonical order = CSR by center. */
typedef struct {
  int32_t n_nodes;
  int32_t n_edges;
  int32_t n_graphs;
  int32_t _pad;
  const int32_t* rowptr;    /* [N+1]  CSR by center                                   */
  const int32_t* col;       /* [E]    neighbor of edge e                              */
  const int32_t* t_rowptr;  /* [N+1]  CSR by neighbor (transposed)                    */
  const int32_t* t_row;     /* [E]    center of transposed slot                       */
  const int32_t* t_eid;     /* [E]    canonical edge id of transposed slot            */
  const int8_t* offsets;    /* [E,4]  integer cell offsets (x,y,z,0) or NULL (no PBC) */
  const float* cell;        /* [G,3,3] lattice rows or NULL                           */
  const int32_t* node_graph;/* [N]    graph id per node (needed when cell && G > 1)   */
  const int32_t* tile_ptr;  /* [E/Tc+2] node bounds of Tc-edge tiles of the CSR (xeq_csr_tile_bounds, Tc = xeq_center_tile_edges())   */
  const int32_t* t_tile_ptr;/* [E/Tn+2] same for the transposed CSR, Tn = xeq_neighbor_tile_edges()                                */
  int32_t n_tiles;          /* tiles described by tile_ptr   (entries = n_tiles + 1)                                                */
  int32_t t_n_tiles;        /* tiles described by t_tile_ptr                                                                       */
  int32_t tile_mode;        /* 0: edge-block tiles (xeq_csr_tile_bounds).  1: molecule tiles -- tile_ptr = t_tile_ptr = the batch  */
                            /*    `ptr` array, every neighbor of a tile's nodes lies inside the tile: the kernels then stage the   */
                            /*    tile's rows in shared memory once instead of gathering them per edge                             */
  int32_t max_tile_nodes;   /* tile_mode 1: upper bound of the nodes per tile known to the caller (0 = unknown).  When every tile   */
                            /* fits the shared-memory window the forward kernel packs rows in-kernel and skips its packing pass    */
} xeq_graph_t;

int xeq_version(void);
const char* xeq_last_error(void);
/* Number of SMs of the current device (grid sizing is done inside the library). */
int xeq_num_sms(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
long long xeq_launch_count(void);

/* ------------------------------------------------------------------------------------
 * K1  radius graph.  Replaces torch_cluster.radius_graph (data/transform.py:58-64) and
 * radius_graph_pbc (data/radius_graph.py:35-192, called at data/transform.py:43-49).
 *
 * Non-periodic (cell == NULL): all ordered pairs a != b of the same graph with
 *   (xa-xb)^2 + (ya-yb)^2 + (za-zb)^2 < r^2   (fp32, no FMA contraction, strict).
 * Periodic: positions are wrapped into the cell (data/radius_graph.py:6-32), images
 *   -rep..rep per periodic axis (rep_host, data/radius_graph.py:61-89) are enumerated and
 *   pairs with 0.01 < sqrt(sum (a - (b + o@cell))^2) < r kept (:125); offsets are referred
 *   back to the unwrapped positions (:186-190).  Self-image pairs (a == b, o != 0) are kept.
 * Large single graphs use a cell list (bins >= r_c); small/batched graphs a per-graph scan.
 *
 * Two phases (the host reads rowptr[N] in between to size the outputs):
 *   count: degree -> exclusive scan -> rowptr[N+1]
 *   fill : col[E], offsets[E,4] (periodic), optional COO edge_index[2,E] (int64) and
 *          cell_offsets[E,3] (float) for API compatibility with the reference.
 * Rows come out sorted by (neighbor, ox, oy, oz) => the COO output is canonically sorted.
 * ---------------------------------------------------------------------------------- */
size_t xeq_radius_graph_workspace_bytes(int32_t n_nodes, int32_t n_graphs, int periodic);

int xeq_radius_graph_count(const float* pos, int32_t n_nodes,
                           const int32_t* graph_ptr /* [G+1] */, const int32_t* node_graph /* [N] */,
                           int32_t n_graphs, const float* cell /* [G,3,3] or NULL */,
                           const int32_t* pbc_host /* [3] */, const int32_t* rep_host /* [3] */,
                           float cutoff, int32_t* rowptr /* [N+1] out */,
                           void* workspace, size_t workspace_bytes, xeq_stream_t stream);

int xeq_radius_graph_fill(const float* pos, int32_t n_nodes,
                          const int32_t* graph_ptr, const int32_t* node_graph, int32_t n_graphs,
                          const float* cell, const int32_t* pbc_host, const int32_t* rep_host,
                          float cutoff, int32_t* rowptr /* in; capacity mode: clamped to edge_capacity in place */,
                          int32_t* col /* [E] out */,
                          int8_t* offsets /* [E,4] out or NULL */,
                          int64_t* edge_index /* [2,E] out or NULL */,
                          float* cell_offsets /* [E,3] out or NULL */,
                          int32_t edge_capacity /* > 0: capacity mode, see below; 0: outputs sized from rowptr[N] */,
                          int32_t* overflow /* device flag set to 1 when E > edge_capacity, or NULL */,
                          void* workspace, size_t workspace_bytes, xeq_stream_t stream);
/* Capacity mode (CUDA-graph replay, MD loops): the caller allocates col/offsets/t_row/t_eid for
 * `edge_capacity` edges once, passes n_edges = edge_capacity everywhere (xeq_graph_t.n_edges,
 * xeq_csr_transpose, xeq_csr_tile_bounds) and never reads the edge count back: every kernel takes
 * the live count from rowptr[N] on the device, tiles past it are empty.  The COO output then uses
 * row stride `edge_capacity`.  When a structure has more than `edge_capacity` edges the surplus edges are
 * dropped, `overflow` is raised and rowptr is clamped to the capacity, so consumers stay inside the arrays
 * (the results are then those of the truncated list: check the flag). */

/* CSR from a caller-supplied COO edge list that is already sorted by center
 * (edge_index[0] non-decreasing): rowptr from segment boundaries, col = int32(edge_index[1]),
 * offsets = int8(cell_offsets).  Entry for graphs that did not come from K1
 * (the reference model takes `edge_index` from the data dict, nn/basic.py:67). */
int xeq_csr_from_sorted_coo(const int64_t* edge_index /* [2,E] */, const float* cell_offsets /* [E,3] or NULL */,
                            int32_t n_nodes, int32_t n_edges, int32_t* rowptr, int32_t* col,
                            int8_t* offsets /* [E,4] or NULL */, xeq_stream_t stream);

/* Transposed structure (edges grouped by neighbor), deterministic slot order (by edge id). */
size_t xeq_csr_transpose_workspace_bytes(int32_t n_nodes, int32_t n_edges);
int xeq_csr_transpose(const int32_t* rowptr, const int32_t* col, int32_t n_nodes, int32_t n_edges,
                      int32_t* t_rowptr, int32_t* t_row, int32_t* t_eid,
                      void* workspace, size_t workspace_bytes, xeq_stream_t stream);

/* Work partition of the edge kernels: rows are weighed as (edges + 4) slots -- a row switch costs about a quad of
 * slots, and rows without edges (ghost atoms of a sharded run) must not pile up in one tile; tile k = the nodes whose
 * weighted prefix falls inside [k*T, (k+1)*T); tile_ptr[k] = first such node, k = 0 .. xeq_csr_tile_count(...)
 * (last entry = n_nodes).  Node-aligned tiles let one CTA own whole rows, so segment sums need no atomics. */
int xeq_csr_tile_count(int32_t n_nodes, int32_t n_edges, int32_t tile_edges);
int xeq_center_tile_edges(void);
int xeq_neighbor_tile_edges(void);
int xeq_csr_tile_bounds(const int32_t* rowptr, int32_t n_nodes, int32_t n_edges, int32_t tile_edges,
                        int32_t* tile_ptr /* [xeq_csr_tile_count(...) + 1] out */, xeq_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K2  fused edge message.  Replaces, per XPainnMessage.forward (nn/xpainn.py:140-159):
 * rbf/cutoff/SphericalHarmonics (nn/xpainn.py:66-74, nn/rbf.py:43-57,143-150), rbf_lin addmm,
 * index_select x2, 2x ElementwiseTensorProduct, index_add x2 -- and the edge-vector
 * preparation of compute_edge_data (nn/basic.py:114-131).  Nothing E-sized is materialised.
 *
 *   r_e = pos[i] - pos[j] - o_e @ cell,  d = |r_e|,  psi_0 = chi(d), psi_k = chi(d) phi_k(d)
 *   w_e[h] = b[h] psi_0 + sum_k W[h,k] psi_k            (h < H = C + 2M)
 *   x_out[i]        = x_in[i]        + sum_e s[j, 2M+c] w_e[2M+c]
 *   V_out[i,(q,m)]  = V_in[i,(q,m)]  + sum_e s[j,q] w_e[q] v[j,(q,m)] + s[j,M+q] w_e[M+q] Y_lm(r_e)
 * x_in / V_in may be NULL (treated as zero).  One CTA walks whole CSR rows with one thread
 * per irrep channel; the segment sum lives in registers (no atomics, deterministic).
 * ---------------------------------------------------------------------------------- */
size_t xeq_edge_message_fwd_workspace_bytes(const xeq_graph_t* g, const xeq_dims_t* dims);
int xeq_edge_message_fwd(const xeq_graph_t* g, const xeq_dims_t* dims, const float* pos,
                         const float* s /* [N,H] */, const float* v /* [N,D] cm */,
                         const float* x_in /* [N,C] */, const float* V_in /* [N,D] cm */,
                         const float* W_rbf /* [H,B] */, const float* b_rbf /* [H] */, const float* freq /* [B] */,
                         float* x_out, float* V_out,
                         void* workspace, size_t workspace_bytes, xeq_stream_t stream);

/* K2b  first derivatives (forces; nn/basic.py:143-159 replays the ops above in reverse).
 * Given gx = dL/dx_out [N,C], gV = dL/dV_out [N,D]:  gs [N,H], gv [N,D], gpos [N,3] and, when
 * gW != NULL, gW [H,B], gb [H], gfreq [B] (gfreq is *written*, the caller sums over layers;
 * nn/rbf.py:143-144 shares freq).  dL/dx_in = gx and dL/dV_in = gV are the identity.
 * Any of gs/gv/gpos may be NULL (skipped). */
size_t xeq_edge_message_bwd_workspace_bytes(const xeq_graph_t* g, const xeq_dims_t* dims, int want_wgrad);
int xeq_edge_message_bwd(const xeq_graph_t* g, const xeq_dims_t* dims, const float* pos,
                         const float* s, const float* v,
                         const float* W_rbf, const float* b_rbf, const float* freq,
                         const float* gx, const float* gV,
                         float* gs, float* gv, float* gpos,
                         float* gW, float* gb, float* gfreq,
                         void* workspace, size_t workspace_bytes, xeq_stream_t stream);

/* K2bb  double backward (force training: utils/trainer.py:295-302 calls loss.backward()
 * through forces obtained with create_graph=True, nn/basic.py:150-156).
 * With Phi = <gx, dx> + <gV, dV> the outputs of K2b are dPhi/d(s, v, pos).  Given cotangents
 * (a_s, a_v, a_pos, a_cell) of (gs, gv, gpos, gcell) (any may be NULL = zero) this returns the gradient of
 *   Psi = <a_s, dPhi/ds> + <a_v, dPhi/dv> + <a_pos, dPhi/dpos> + <a_cell, dPhi/dcell>
 * with respect to gx, gV, s, v, pos, W_rbf, b_rbf, freq.  Output pointers may be NULL. */
/* Per-node pieces of the CELL gradient of a periodic graph (virial / stress through the strain trick,
 * nn/basic.py:93-107, 162-199):  rows[3a+b][n] = sum over the edges e of CSR row n of
 * cell_offsets[e][a] * (dE/dr_e)[b], so that  dE/dcell[g][a][b] = - sum_{n in g} rows[3a+b][n]
 * (the edge vector is pos_i - pos_j - cell_offsets @ cell, nn/basic.py:119-128).  Reads the per-edge d/dr records
 * that the preceding xeq_edge_message_bwd / xeq_edge_message_bwdbwd call (gpos / o_pos != NULL, same stream) left at
 * the start of ITS workspace.
 * rows: [9, n_nodes] floats.  Deterministic. */
int xeq_edge_cell_grad_rows(const xeq_graph_t* g, const xeq_dims_t* dims, const void* bwd_workspace,
                            float* rows, xeq_stream_t stream);

size_t xeq_edge_message_bwdbwd_workspace_bytes(const xeq_graph_t* g, const xeq_dims_t* dims, int want_wgrad);
int xeq_edge_message_bwdbwd(const xeq_graph_t* g, const xeq_dims_t* dims, const float* pos,
                            const float* s, const float* v,
                            const float* W_rbf, const float* b_rbf, const float* freq,
                            const float* gx, const float* gV,
                            const float* a_s, const float* a_v, const float* a_pos,
                            const float* a_cell /* [G,3,3] cotangent of the cell gradient (periodic virial in a training
                                                   loss): the edge tangent becomes a_pos_i - a_pos_j - offsets @ a_cell;
                                                   NULL = zero.  The second-order cell gradient follows from the per-edge
                                                   records of THIS call with xeq_edge_cell_grad_rows */,
                            float* o_gx, float* o_gV, float* o_s, float* o_v, float* o_pos,
                            float* o_W, float* o_b, float* o_freq,
                            void* workspace, size_t workspace_bytes, xeq_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Segment sum over contiguous segments: out[g] = sum_{ptr[g] <= n < ptr[g+1]} src[n].
 * Replaces torch_scatter.scatter_sum(atomic_energies, batch) (nn/output.py:124) for the
 * sorted `batch` every collated batch has (keys.py:9-10).
 * ---------------------------------------------------------------------------------- */
int xeq_segment_sum(const float* src, const int32_t* seg_ptr /* [G+1] */, int32_t n_segments,
                    float* out, xeq_stream_t stream);

/* out[c] = sum over rows of src[r, c] (row stride ld floats): the bias gradient of an nn.Linear = the sum of the
 * output gradient over the nodes (autograd of nn/xpainn.py:111-115, 195-199; torch: grad_output.sum(0)).
 * Deterministic (fixed summation order). */
int xeq_colsum(const float* src, int32_t n_rows, int32_t n_cols, int32_t ld, float* out /* [n_cols] */,
               xeq_stream_t stream);

/* An nn.Linear with ONE output feature -- the 64 -> 1 energy read-out (nn/output.py:107-111, F.linear in the
 * reference) -- and its derivatives, closed under differentiation (each op's gradients are the other two):
 *   rowdot           y[r]     = sum_c x[r, c] w[c] (+ bias[0])      forward
 *   outer            out[r,c] = g[r] w[c]                           gradient w.r.t. x
 *   colsum_weighted  out[c]   = sum_r row_weight[r] src[r, c]       gradient w.r.t. w (deterministic, as xeq_colsum)
 * x / src may be row-strided views (ld floats between rows). */
int xeq_rowdot(const float* x, int32_t ld, const float* w /* [n_cols] */, const float* bias /* NULL or [1] */,
               int32_t n_rows, int32_t n_cols, float* y /* [n_rows] */, xeq_stream_t stream);
int xeq_outer(const float* g /* [n_rows] */, const float* w /* [n_cols] */, int32_t n_rows, int32_t n_cols,
              float* out /* [n_rows, n_cols] */, xeq_stream_t stream);
int xeq_colsum_weighted(const float* src, const float* row_weight /* [n_rows] */, int32_t n_rows, int32_t n_cols,
                        int32_t ld, float* out /* [n_cols] */, xeq_stream_t stream);

/* e3nn (mul-major, m-fastest) <-> cm (component-major) layout of [N, D] tensors.
 * direction 0: e3nn -> cm, 1: cm -> e3nn. */
int xeq_layout_convert(const float* src, float* dst, int32_t n_nodes, const xeq_dims_t* dims,
                       int direction, xeq_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K3  node-side dense contractions on tcgen05 tensor cores (3xTF32 split, fp32 accumulate in
 * TMEM; fp32-level accuracy).  One launch evaluates a GROUP of independent problems
 *     C = act( alpha * op(A) op(B) + bias )        op(A): [m,k]   op(B): [k,n]   C: [m,n]
 * Storage: a_trans = 0 -> A is [m,k] row-major (lda), 1 -> A is [k,m] row-major;
 *          b_trans = 0 -> B is [k,n] row-major (ldb), 1 -> B is [n,k] row-major (the
 *          nn.Linear weight form).  bias [n] or NULL; act 0 = none, 1 = SiLU.
 * Replaces: nn.Linear in scalar_mlp / update_mlp / dot_lin / embedding / out_mlp
 * (nn/xpainn.py:44-47,111-115,190,195-199; nn/output.py:107-111), the per-l blocks of
 * e3nn o3.Linear update_U / update_V (nn/xpainn.py:186-187,211-212: one problem per (l, m)
 * on the cm layout, alpha = 1/sqrt(mul_l)), and the grad-input / grad-weight products of
 * their first and second derivatives (grad-weight: a_trans = 1, b_trans = 0, k = N nodes,
 * split_k > 1: partials in the workspace are reduced in fixed order => deterministic).
 * Consecutive problems that name the same output (c, m, n, ldc) ACCUMULATE into it (the 2l+1
 * component blocks of one o3.Linear weight block): their partial slabs and the split-K slabs go
 * through the workspace and one fixed-order reduction kernel.
 * Constraints: pointers 16-byte aligned, leading dimensions % 4 == 0, the contiguous
 * extent of each operand % 4 == 0; <= 20 problems per launch; act needs split_k == 1
 * and no shared outputs.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const float* a;
  const float* b;
  const float* bias;
  float* c;
  int32_t m, n, k;
  int32_t lda, ldb, ldc;
  int32_t a_trans, b_trans;
  float alpha;
  int32_t act;
} xeq_gemm_t;

size_t xeq_gemm_workspace_bytes(const xeq_gemm_t* problems_host, int32_t n_problems, int32_t split_k);
int xeq_gemm_tf32x3(const xeq_gemm_t* problems_host, int32_t n_problems, int32_t split_k,
                    void* workspace, size_t workspace_bytes, xeq_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Node-side normalisations, fused, with first and second derivatives.
 * EquivariantLayerNorm (nn/o3layer.py:145-171, called at nn/xpainn.py:131,209) on the cm layout and
 * nn.LayerNorm(node_dim) (nn/xpainn.py:123,130,201,208), which is the same map for irreps "Cx0e"
 * (mul1 = mul2 = 0):
 *     z = x with the mul0 scalars centred;  rho = rsqrt(sum z^2 / (mul0+mul1+mul2) + eps)
 *     y_i = gamma[irrep(i)] z_i rho + (i < mul0 ? beta[i] : 0)
 * bwd:    g = dL/dy  ->  gx [N,D]; ggamma [M], gbeta [mul0] (either NULL = skipped)
 * bwd: gx = d<g,y>/dx (+ gx_add: the gradient reaching x through its other consumer, e.g. the residual -- summed
 *      here instead of by a separate elementwise kernel)
 * bwdbwd: cotangent a of gx -> dx = d<a,gx>/dx, dg = d<a,gx>/dg, dgamma [M]  (outputs may be NULL)
 * One warp per row; parameter gradients via per-CTA partial rows in the workspace and a fixed-order
 * reduction (deterministic).  Row widths D in {32,64,128,256,288,480,960}; mul* multiples of 32.
 * ---------------------------------------------------------------------------------- */
size_t xeq_irreps_norm_workspace_bytes(int32_t n_rows, int32_t mul0, int32_t mul1, int32_t mul2);
int xeq_irreps_norm_fwd(const float* x, const float* gamma, const float* beta, int32_t n_rows,
                        int32_t mul0, int32_t mul1, int32_t mul2, float eps, float* y, xeq_stream_t stream);
int xeq_irreps_norm_bwd(const float* x, const float* gamma, const float* g, int32_t ld_g /* row stride of g, 0 = D */,
                        const float* gx_add /* NULL or [N,D] */,
                        int32_t n_rows, int32_t mul0, int32_t mul1, int32_t mul2, float eps,
                        float* gx, float* ggamma, float* gbeta,
                        void* workspace, size_t workspace_bytes, xeq_stream_t stream);
int xeq_irreps_norm_bwdbwd(const float* x, const float* gamma, const float* g, int32_t ld_g, const float* a, int32_t n_rows,
                           int32_t mul0, int32_t mul1, int32_t mul2, float eps,
                           float* dx, float* dg, float* dgamma,
                           void* workspace, size_t workspace_bytes, xeq_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Node-side per-irrep maps of XPainnUpdate.forward (nn/xpainn.py:206-231) on the cm layout and the
 * SiLU of the MLPs, fused, each with first (bwd) and second (bwdbwd) derivatives.
 *   invariant_dot : nrm[n,q] = sqrt(sum_m W^2 + 1e-10) - 1e-5   (Invariant, nn/o3layer.py:40-44)
 *                   t0[n,q]  = sum_m U W                         (EquivariantDot, nn/o3layer.py:104-109)
 *                   nrm has row stride ld_nrm (>= M) so it can land inside the [N, C+M] MLP input.
 *   gate_residual : x' = x + a_sv*t + a_ss ;  V'[(q,m)] = V[(q,m)] + a_vv[q] U[(q,m)]
 *                   with a = [a_vv (M) | a_sv (C) | a_ss (C)] (nn/xpainn.py:218-229), C = mul0.
 *                   bwd returns d/da, d/dU, d/dt (d/dx = gx, d/dV = gV are the identity).
 *   silu          : y = u * sigmoid(u)  (nn/basic.py:255-256)
 * bwdbwd takes cotangents of the bwd outputs and returns the derivatives with respect to the
 * bwd inputs.  NULL input gradients are treated as zero, NULL outputs are skipped where noted.
 * ---------------------------------------------------------------------------------- */
int xeq_invariant_dot_fwd(const float* U, const float* W, int32_t n, int32_t mul0, int32_t mul1, int32_t mul2,
                          float* nrm, int32_t ld_nrm, float* t0, xeq_stream_t stream);
int xeq_invariant_dot_bwd(const float* U, const float* W, const float* gn, int32_t ld_gn, const float* gt,
                          const float* gU_add /* NULL or [N,D]: summed onto gU (gradient through U's other consumer) */,
                          int32_t n, int32_t mul0, int32_t mul1, int32_t mul2, float* gU, float* gW, xeq_stream_t stream);
int xeq_invariant_dot_bwdbwd(const float* U, const float* W, const float* gn, int32_t ld_gn, const float* gt,
                             const float* aU, const float* aW, int32_t n, int32_t mul0, int32_t mul1, int32_t mul2,
                             float* d_gn, float* d_gt, float* dU, float* dW, xeq_stream_t stream);
int xeq_gate_residual_fwd(const float* a, const float* U, const float* t, const float* x, const float* V, int32_t n,
                          int32_t mul0, int32_t mul1, int32_t mul2, float* x_out, float* V_out, xeq_stream_t stream);
int xeq_gate_residual_bwd(const float* a, const float* U, const float* t, const float* gx, const float* gV, int32_t n,
                          int32_t mul0, int32_t mul1, int32_t mul2, float* ga, float* gU, float* gt, xeq_stream_t stream);
int xeq_gate_residual_bwdbwd(const float* a, const float* U, const float* t, const float* gx, const float* gV,
                             const float* c_a, const float* c_U, const float* c_t, int32_t n,
                             int32_t mul0, int32_t mul1, int32_t mul2,
                             float* d_gx, float* d_gV, float* d_a, float* d_U, float* d_t, xeq_stream_t stream);
int xeq_silu_fwd(const float* u, size_t n, float* y, xeq_stream_t stream);
int xeq_silu_bwd(const float* u, const float* g, size_t n, float* gu, xeq_stream_t stream);
int xeq_silu_bwdbwd(const float* u, const float* g, const float* c, size_t n, float* dg, float* du, xeq_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Whole-model inference runtime (csrc/model_runtime.cu): XPaiNN energy + forces as ONE call, no Python / torch /
 * autograd in the host process.  Deployment form of the path: the reference exports a TorchScript archive that a
 * LAMMPS pair style / the GROMACS NNP interface drives through libtorch (run/jit_script.py:28-86,
 * interface/jit_model.py:12-216: compute_edge_data -> mods -> compute_properties, forces by torch.autograd.grad);
 * here the engine links this library and passes its own neighbour list as an xeq_graph_t
 * (xeq_csr_from_sorted_coo + xeq_csr_transpose + xeq_csr_tile_bounds, or K1).
 *
 * Default model only (nn/model.py:57-70: layer_norm, silu, output_modes = ["energy"], no charge / spin conditioning).
 * `weights`: ONE flat fp32 DEVICE blob, caller-owned and kept alive while the handle is used, 16-byte aligned, every
 * tensor padded to a multiple of 4 floats, in this order (state_dict names of the reference):
 *   embedding.embedding.0.embed_ten [n_species, embed_dim], embedding.embedding.1.{weight [C, embed_dim], bias [C]},
 *   embedding.rbf.freq [B];
 *   per action block i:  message_i.{norm.weight, norm.bias, o3norm.affine_weight, o3norm.affine_bias,
 *                          scalar_mlp.0.weight, .0.bias, scalar_mlp.2.weight, .2.bias, rbf_lin.weight, rbf_lin.bias},
 *                        update_i.{norm.weight, norm.bias, o3norm.affine_weight, o3norm.affine_bias, update_U.weight,
 *                          update_U.bias, update_V.weight, update_V.bias, dot_lin.weight, update_mlp.0.weight, .0.bias,
 *                          update_mlp.2.weight, .2.bias};
 *   output_energy.out_mlp.{0.weight, 0.bias, 2.weight, 2.bias}.
 * xeq_model_weight_count() gives the blob length; xeq_model_create() only records the description (no device
 * allocation, no copy); the handle is immutable and may be shared by threads / streams.
 *
 * xeq_model_energy_forces(): energy [G], atomic_energies [N], forces [N,3] = -d(sum_g energy[g])/dpos (NULL: energies
 * only).  Nodes outside [seg_ptr[0], seg_ptr[G]) -- an MD engine's ghost atoms, appended after its local atoms -- act
 * as neighbours and receive forces (to be reverse-communicated) but contribute no energy.
 * The forward pass issues the entry points above in the order of the nn modules; the force pass is the reverse sweep
 * torch.autograd.grad(E, pos) performs over them (nn/basic.py:143-159) with the same kernels, arguments and summation
 * order: results are bit-identical to the Python module path.  Asynchronous on `stream`, no allocation, no host
 * synchronisation; the workspace must be 256-byte aligned.
 * ---------------------------------------------------------------------------------- */
/* An engine's edge list -> xeq_graph_t in one call: CSR + transposed CSR + work tiles (xeq_csr_from_sorted_coo,
 * xeq_csr_transpose, xeq_csr_tile_bounds x 2) carved out of ONE caller-owned device buffer `storage` (256-byte aligned,
 * xeq_graph_from_coo_bytes() long), and the HOST struct *graph_host filled with pointers into it (plus the caller's
 * `cell` / `node_graph`).  edge_index [2,E] int64 sorted by center (row 0); periodic structures pass cell [G,3,3] and
 * cell_offsets [E,3] together.  Launches only; `storage` must outlive every use of the struct. */
size_t xeq_graph_from_coo_bytes(int32_t n_nodes, int32_t n_edges, int periodic);
int xeq_graph_from_coo(const int64_t* edge_index, const float* cell_offsets /* or NULL */, const float* cell /* or NULL */,
                       const int32_t* node_graph /* [N], needed for several periodic graphs, else NULL */,
                       int32_t n_nodes, int32_t n_edges, int32_t n_graphs, void* storage, size_t storage_bytes,
                       xeq_graph_t* graph_host, xeq_stream_t stream);

typedef struct xeq_model xeq_model_t;
size_t xeq_model_weight_count(const xeq_dims_t* dims, int32_t n_layers, int32_t hidden_dim, int32_t embed_dim,
                              int32_t n_species);
int xeq_model_create(const xeq_dims_t* dims, int32_t n_layers /* action_blocks */, int32_t hidden_dim /* EnergyOut: 64 */,
                     int32_t embed_dim /* aux56: 56 */, int32_t n_species /* rows of embed_ten */,
                     const float* weights /* device */, size_t n_weights, xeq_model_t** model);
void xeq_model_destroy(xeq_model_t* model);
size_t xeq_model_workspace_bytes(const xeq_model_t* model, const xeq_graph_t* g, int want_forces);
int xeq_model_energy_forces(const xeq_model_t* model, const xeq_graph_t* g, const float* pos /* [N,3] */,
                            const int32_t* atomic_numbers /* [N] */, const int32_t* seg_ptr /* [G+1] */,
                            float* energy /* [G] */, float* atomic_energies /* [N] */, float* forces /* [N,3] or NULL */,
                            void* workspace, size_t workspace_bytes, xeq_stream_t stream);
/* Same call with a second stream for the independent branches of the module graph (norm(x) -> scalar MLP beside
 * o3norm(V), update_U beside update_V, dot_lin beside the update MLP, and their mirror images in the force pass):
 * at MD sizes every kernel is a fraction of a wave and the step is a latency chain, so the branches overlap.  Forks
 * and joins are event record / wait pairs between `stream` and `aux_stream` (capturable; everything is joined back
 * into `stream` before the call returns).  Same kernels and arguments: the results are identical.  aux_stream NULL
 * (or == stream) = the single-stream schedule. */
int xeq_model_energy_forces_mt(const xeq_model_t* model, const xeq_graph_t* g, const float* pos,
                               const int32_t* atomic_numbers, const int32_t* seg_ptr,
                               float* energy, float* atomic_energies, float* forces,
                               void* workspace, size_t workspace_bytes, xeq_stream_t stream, xeq_stream_t aux_stream);
/* ... and with the virial [G,3,3] = -dE/dstrain of the reference's strain trick (nn/basic.py:93-107, 162-199: positions
 * and cell displaced by a symmetrised per-graph strain): sum_i pos_i (x) dE/dpos_i plus, for periodic graphs, the cell
 * term from the per-edge d/dr records of the force pass (xeq_edge_cell_grad_rows), symmetrised.  stress = virial /
 * volume (interface/ase_calculator.py:104-110).  virial NULL = the call above; forces must not be NULL. */
int xeq_model_energy_forces_virial(const xeq_model_t* model, const xeq_graph_t* g, const float* pos,
                                   const int32_t* atomic_numbers, const int32_t* seg_ptr,
                                   float* energy, float* atomic_energies, float* forces, float* virial,
                                   void* workspace, size_t workspace_bytes, xeq_stream_t stream, xeq_stream_t aux_stream);

#ifdef __cplusplus
}
#endif
#endif /* XEQ_B200_H */
