// K1: radius graph (non-periodic batched / periodic with image enumeration), CSR utilities.
// Replaces torch_cluster.radius_graph (data/transform.py:58-64) and radius_graph_pbc
// (data/radius_graph.py:35-192).  See include/xeq_b200.h for the contract.
//
// Edge test arithmetic mirrors the reference so that the edge SET is bit-identical away
// from 1-ulp ties at the cutoff:
//   non-periodic: (dx*dx + dy*dy) + dz*dz < r*r        fp32, no FMA contraction, strict
//   periodic    : 0.01 < sqrt((dx*dx + dy*dy) + dz*dz) < r  on wrapped + image coordinates
#include "common.cuh"

namespace xeq {

// ------------------------------------------------------------------------------------------
// exclusive scan of int32 (two-level, deterministic)
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = warp_incl_scan(v);
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int ws = (lane < (blockDim.x >> 5)) ? warp_sums[lane] : 0;
    ws = warp_incl_scan(ws);
    warp_sums[lane] = ws;
  }
  __syncthreads();
  const int warp_off = wid ? warp_sums[wid - 1] : 0;
  *total = warp_sums[(blockDim.x >> 5) - 1];
  __syncthreads();
  return warp_off + incl - v;
}

// out[i] = exclusive prefix within the block; block_tot[b] = block sum
__global__ void __launch_bounds__(SCAN_THREADS) scan_blocks_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                                   int* __restrict__ block_tot, int n) {
  const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    sum += v[k];
  }
  int total;
  int pre = block_excl_scan(sum, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = pre;
    pre += v[k];
  }
  if (threadIdx.x == 0) block_tot[blockIdx.x] = total;
}

// single block: exclusive scan of up to SCAN_BLOCK block totals, in place; grand total -> *grand
__global__ void __launch_bounds__(SCAN_THREADS) scan_totals_kernel(int* __restrict__ block_tot, int nblocks,
                                                                   int* __restrict__ grand) {
  const int base = threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < nblocks) ? block_tot[base + k] : 0;
    sum += v[k];
  }
  int total;
  int pre = block_excl_scan(sum, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < nblocks) block_tot[base + k] = pre;
    pre += v[k];
  }
  if (threadIdx.x == 0) *grand = total;
}

__global__ void add_offsets_kernel(int* __restrict__ out, const int* __restrict__ block_off, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += block_off[i / SCAN_BLOCK];
}

// in[n] -> out[n+1] (out[n] = total).  scratch: ceil(n/SCAN_BLOCK) ints.  n <= SCAN_BLOCK^2.
static int exclusive_scan(const int* in, int* out, int n, int* scratch, cudaStream_t st) {
  if (n == 0) {
    XEQ_CUDA(cudaMemsetAsync(out, 0, sizeof(int), st));
    return XEQ_OK;
  }
  const int nblocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  XEQ_CHECK_ARG(nblocks <= SCAN_BLOCK, "exclusive_scan: n=%d too large", n);
  scan_blocks_kernel<<<nblocks, SCAN_THREADS, 0, st>>>(in, out, scratch, n);
  scan_totals_kernel<<<1, SCAN_THREADS, 0, st>>>(scratch, nblocks, out + n);
  if (nblocks > 1) add_offsets_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, scratch, n);
  XEQ_LAUNCHED(nblocks > 1 ? 3 : 2);
  return XEQ_OK;
}

static inline size_t scan_scratch_ints(int n) { return (size_t)(n + SCAN_BLOCK - 1) / SCAN_BLOCK + 1; }

// ------------------------------------------------------------------------------------------
// periodic wrapping (data/radius_graph.py:6-32)
// ------------------------------------------------------------------------------------------
struct PbcParams {
  int pbc[3];
  int rep[3];
};

__device__ __forceinline__ void inv3x3(const float* c, float* inv) {
  const float a = c[0], b = c[1], cc = c[2], d = c[3], e = c[4], f = c[5], g = c[6], h = c[7], i = c[8];
  const float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const float det = a * A + b * B + cc * C;
  const float id = 1.0f / det;
  inv[0] = A * id;  inv[1] = -(b * i - cc * h) * id; inv[2] = (b * f - cc * e) * id;
  inv[3] = B * id;  inv[4] = (a * i - cc * g) * id;  inv[5] = -(a * f - cc * d) * id;
  inv[6] = C * id;  inv[7] = -(a * h - b * g) * id;  inv[8] = (a * e - b * d) * id;
}

// pw = (frac - floor(frac)) @ cell on periodic axes; shift = floor(frac) (0 on open axes)
__global__ void wrap_positions_kernel(const float* __restrict__ pos, const int* __restrict__ node_graph,
                                      const float* __restrict__ cell, PbcParams pp, int n, float* __restrict__ pw,
                                      int* __restrict__ shift) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const float* c = cell + 9 * (node_graph ? node_graph[a] : 0);
  float inv[9];
  inv3x3(c, inv);
  const float p[3] = {pos[3 * a], pos[3 * a + 1], pos[3 * a + 2]};
  float fr[3];
  int sh[3];
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    float f = __fadd_rn(__fadd_rn(__fmul_rn(p[0], inv[x]), __fmul_rn(p[1], inv[3 + x])), __fmul_rn(p[2], inv[6 + x]));
    float fl = pp.pbc[x] ? floorf(f) : 0.0f;
    sh[x] = (int)fl;
    fr[x] = f - fl;
  }
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    pw[3 * a + x] =
        __fadd_rn(__fadd_rn(__fmul_rn(fr[0], c[x]), __fmul_rn(fr[1], c[3 + x])), __fmul_rn(fr[2], c[6 + x]));
    shift[3 * a + x] = sh[x];
  }
}

// ------------------------------------------------------------------------------------------
// per-graph scan: one warp per center atom, lanes stride over (neighbor, image) candidates in
// canonical order, so rows come out sorted by (neighbor, ox, oy, oz).
// ------------------------------------------------------------------------------------------
template <bool PERIODIC>
__device__ __forceinline__ bool edge_test(const float* pa, const float* pb, const float* c, int ox, int oy, int oz,
                                          float r, float r2, bool same_atom) {
  if (!PERIODIC) {
    if (same_atom) return false;
    const float dx = pa[0] - pb[0], dy = pa[1] - pb[1], dz = pa[2] - pb[2];
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return d2 < r2;
  } else {
    float b[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      const float img = __fadd_rn(__fadd_rn(__fmul_rn((float)ox, c[x]), __fmul_rn((float)oy, c[3 + x])),
                                  __fmul_rn((float)oz, c[6 + x]));
      b[x] = __fadd_rn(pb[x], img);
    }
    const float dx = pa[0] - b[0], dy = pa[1] - b[1], dz = pa[2] - b[2];
    const float D = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    return D < r && D > 0.01f;
  }
}

template <bool PERIODIC, bool FILL>
__global__ void __launch_bounds__(256) graph_scan_kernel(const float* __restrict__ p /* pos or wrapped pos */,
                                                          const int* __restrict__ shift, const int* __restrict__ graph_ptr,
                                                          const int* __restrict__ node_graph, const float* __restrict__ cell,
                                                          PbcParams pp, float r, int n, int* __restrict__ deg,
                                                          const int* __restrict__ rowptr, int* __restrict__ col,
                                                          int8_t* __restrict__ offsets, long long* __restrict__ edge_index,
                                                          float* __restrict__ cell_offsets, long long coo_stride,
                                                          int capacity, int* __restrict__ overflow) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a >= n) return;
  const int g = node_graph ? node_graph[a] : 0;
  const int g0 = graph_ptr[g], g1 = graph_ptr[g + 1];
  const float r2 = __fmul_rn(r, r);
  const float pa[3] = {p[3 * a], p[3 * a + 1], p[3 * a + 2]};
  float c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int n1 = 1, n2 = 1, nimg = 1;
  if (PERIODIC) {
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = cell[9 * g + k];
    n1 = 2 * pp.rep[1] + 1;
    n2 = 2 * pp.rep[2] + 1;
    nimg = (2 * pp.rep[0] + 1) * n1 * n2;
  }
  const long long total = (long long)(g1 - g0) * nimg;
  int running = 0;
  const int base = FILL ? rowptr[a] : 0;
  for (long long it = 0; it < total; it += 32) {
    const long long idx = it + lane;
    bool pass = false;
    int b = 0, ox = 0, oy = 0, oz = 0;
    if (idx < total) {
      int img = 0;
      if (PERIODIC) {
        b = g0 + (int)(idx / nimg);
        img = (int)(idx % nimg);
        oz = img % n2 - pp.rep[2];
        oy = (img / n2) % n1 - pp.rep[1];
        ox = img / (n2 * n1) - pp.rep[0];
      } else {
        b = g0 + (int)idx;
      }
      const float pb[3] = {p[3 * b], p[3 * b + 1], p[3 * b + 2]};
      pass = edge_test<PERIODIC>(pa, pb, c, ox, oy, oz, r, r2, a == b);
    }
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    const int e_slot = base + running + __popc(m & ((1u << lane) - 1u));
    if (FILL && pass && e_slot >= capacity) {
      if (overflow) *overflow = 1;  // caller-supplied capacity exceeded: edge dropped, flag raised
    } else if (FILL && pass) {
      const int e = e_slot;
      col[e] = b;
      int fx = ox, fy = oy, fz = oz;
      if (PERIODIC) {  // refer offsets to the unwrapped positions (data/radius_graph.py:186-190)
        fx += shift[3 * a] - shift[3 * b];
        fy += shift[3 * a + 1] - shift[3 * b + 1];
        fz += shift[3 * a + 2] - shift[3 * b + 2];
        if (offsets) {
          offsets[4 * (size_t)e] = (int8_t)fx;
          offsets[4 * (size_t)e + 1] = (int8_t)fy;
          offsets[4 * (size_t)e + 2] = (int8_t)fz;
          offsets[4 * (size_t)e + 3] = 0;
        }
        if (cell_offsets) {
          cell_offsets[3 * (size_t)e] = (float)fx;
          cell_offsets[3 * (size_t)e + 1] = (float)fy;
          cell_offsets[3 * (size_t)e + 2] = (float)fz;
        }
      }
      if (edge_index) {
        edge_index[e] = a;
        edge_index[coo_stride + e] = b;
      }
    }
    running += __popc(m);
  }
  if (!FILL && lane == 0) deg[a] = running;
}

// ------------------------------------------------------------------------------------------
// cell list for one large fully periodic graph (the MD-style box of BASELINE.json configs[4]):
// atoms are binned on the fractional axes into nb_x * nb_y * nb_z cells whose perpendicular width is
// >= r (so with image repeat 1 every pair within r lies in the 27 adjacent cells, the wrap of the
// cell index giving the image offset); one warp per center atom scans those cells with the SAME
// distance arithmetic as the brute-force scan, and a per-row rank sort restores the canonical
// (neighbor, ox, oy, oz) order => bit-identical output, O(N) instead of O(27 N^2) candidates.
// ------------------------------------------------------------------------------------------
struct CellGrid {  // lives in the workspace (device side only: no host round trip)
  int nb[3];
  int total;
  float inv[9];
};

__global__ void cell_setup_kernel(const float* __restrict__ cell, float r, int max_bins, CellGrid* __restrict__ grid) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float* c = cell;
  float inv[9];
  inv3x3(c, inv);
  const float a[3][3] = {{c[0], c[1], c[2]}, {c[3], c[4], c[5]}, {c[6], c[7], c[8]}};
  const float vol = fabsf(a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                          a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]));
  int nb[3];
  for (int x = 0; x < 3; ++x) {
    const float* u = a[(x + 1) % 3];
    const float* v = a[(x + 2) % 3];
    const float cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
    const float width = vol / sqrtf(cx * cx + cy * cy + cz * cz);  // perpendicular width of the cell along axis x
    nb[x] = max(1, (int)floorf(width / r * 0.999f));               // a hair wider than r: rounding safety
  }
  while ((long long)nb[0] * nb[1] * nb[2] > max_bins) {
    int big = 0;
    if (nb[1] > nb[big]) big = 1;
    if (nb[2] > nb[big]) big = 2;
    nb[big] -= 1;
  }
  grid->nb[0] = nb[0]; grid->nb[1] = nb[1]; grid->nb[2] = nb[2];
  grid->total = nb[0] * nb[1] * nb[2];
  for (int k = 0; k < 9; ++k) grid->inv[k] = inv[k];
}

__device__ __forceinline__ void bin_coords(const float* __restrict__ pw, int a, const CellGrid& G, int* b) {
  const float p[3] = {pw[3 * a], pw[3 * a + 1], pw[3 * a + 2]};
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    float f = p[0] * G.inv[x] + p[1] * G.inv[3 + x] + p[2] * G.inv[6 + x];
    f -= floorf(f);
    b[x] = min(G.nb[x] - 1, max(0, (int)(f * (float)G.nb[x])));
  }
}

__global__ void bin_count_kernel(const float* __restrict__ pw, int n, const CellGrid* __restrict__ grid, int* __restrict__ cnt,
                                 int* __restrict__ atom_bin) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const CellGrid G = *grid;
  int b[3];
  bin_coords(pw, a, G, b);
  const int id = (b[0] * G.nb[1] + b[1]) * G.nb[2] + b[2];
  atom_bin[a] = id;
  atomicAdd(&cnt[id], 1);
}

__global__ void bin_fill_kernel(const int* __restrict__ atom_bin, int n, const int* __restrict__ bin_start, int* __restrict__ cursor,
                                int* __restrict__ bin_atoms) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int id = atom_bin[a];
  bin_atoms[bin_start[id] + atomicAdd(&cursor[id], 1)] = a;
}

// FILL = false: deg[a] = number of edges.  FILL = true: keys (b * 27 + image code) into keys[rowptr[a] ...],
// in cell-traversal order (made canonical by row_sort_decode_kernel).
template <bool FILL>
__global__ void __launch_bounds__(256) cell_scan_kernel(const float* __restrict__ pw, const float* __restrict__ cell, float r, int n,
                                                         const CellGrid* __restrict__ grid, const int* __restrict__ atom_bin,
                                                         const int* __restrict__ bin_start, const int* __restrict__ bin_atoms,
                                                         int* __restrict__ deg, const int* __restrict__ rowptr,
                                                         int* __restrict__ keys, int capacity, int* __restrict__ overflow) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a >= n) return;
  const CellGrid G = *grid;
  float c[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) c[k] = cell[k];
  const float r2 = __fmul_rn(r, r);
  const float pa[3] = {pw[3 * a], pw[3 * a + 1], pw[3 * a + 2]};
  const int id = atom_bin[a];
  const int bz = id % G.nb[2], by = (id / G.nb[2]) % G.nb[1], bx = id / (G.nb[2] * G.nb[1]);
  int running = 0;
  const int base = FILL ? rowptr[a] : 0;
  for (int dd = 0; dd < 27; ++dd) {
    const int dx = dd / 9 - 1, dy = (dd / 3) % 3 - 1, dz = dd % 3 - 1;
    int cx = bx + dx, cy = by + dy, cz = bz + dz, ox = 0, oy = 0, oz = 0;
    if (cx < 0) { cx += G.nb[0]; ox = -1; } else if (cx >= G.nb[0]) { cx -= G.nb[0]; ox = 1; }
    if (cy < 0) { cy += G.nb[1]; oy = -1; } else if (cy >= G.nb[1]) { cy -= G.nb[1]; oy = 1; }
    if (cz < 0) { cz += G.nb[2]; oz = -1; } else if (cz >= G.nb[2]) { cz -= G.nb[2]; oz = 1; }
    const int cid = (cx * G.nb[1] + cy) * G.nb[2] + cz;
    const int s0 = bin_start[cid], s1 = bin_start[cid + 1];
    for (int it = s0; it < s1; it += 32) {
      const int sl = it + lane;
      bool pass = false;
      int b = 0;
      if (sl < s1) {
        b = bin_atoms[sl];
        const float pb[3] = {pw[3 * b], pw[3 * b + 1], pw[3 * b + 2]};
        pass = edge_test<true>(pa, pb, c, ox, oy, oz, r, r2, a == b);
      }
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (FILL && pass) {
        const int e = base + running + __popc(m & ((1u << lane) - 1u));
        if (e >= capacity) {
          if (overflow) *overflow = 1;
        } else {
          keys[e] = b * 27 + (ox + 1) * 9 + (oy + 1) * 3 + (oz + 1);
        }
      }
      running += __popc(m);
    }
  }
  if (!FILL && lane == 0) deg[a] = running;
}

// one warp per row: rank sort of the (unique) keys, then decode into the output arrays
constexpr int ROW_SORT_MAX = 192;
__global__ void __launch_bounds__(256) row_sort_decode_kernel(const int* __restrict__ rowptr, int n, const int* __restrict__ shift,
                                                               int capacity, int* __restrict__ col, int8_t* __restrict__ offsets,
                                                               long long* __restrict__ edge_index, float* __restrict__ cell_offsets,
                                                               long long coo_stride) {
  __shared__ int skeys[8][ROW_SORT_MAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a = blockIdx.x * 8 + warp;
  if (a >= n) return;
  const int e0 = min(rowptr[a], capacity), e1 = min(rowptr[a + 1], capacity);
  const int deg = e1 - e0;
  if (deg <= 0) return;
  if (deg <= ROW_SORT_MAX) {
    for (int i = lane; i < deg; i += 32) skeys[warp][i] = col[e0 + i];
    __syncwarp();
    for (int i = lane; i < deg; i += 32) {
      const int key = skeys[warp][i];
      int rank = 0;
      for (int k = 0; k < deg; ++k) rank += skeys[warp][k] < key;
      col[e0 + rank] = key;
    }
    __syncwarp();
  } else if (lane == 0) {  // very dense rows: in-place insertion sort
    for (int i = e0 + 1; i < e1; ++i) {
      const int key = col[i];
      int k = i - 1;
      while (k >= e0 && col[k] > key) { col[k + 1] = col[k]; --k; }
      col[k + 1] = key;
    }
  }
  __syncwarp();
  for (int e = e0 + lane; e < e1; e += 32) {
    const int key = col[e];
    const int b = key / 27, code = key % 27;
    const int fx = code / 9 - 1 + shift[3 * a] - shift[3 * b];
    const int fy = (code / 3) % 3 - 1 + shift[3 * a + 1] - shift[3 * b + 1];
    const int fz = code % 3 - 1 + shift[3 * a + 2] - shift[3 * b + 2];
    col[e] = b;
    if (offsets) {
      offsets[4 * (size_t)e] = (int8_t)fx; offsets[4 * (size_t)e + 1] = (int8_t)fy;
      offsets[4 * (size_t)e + 2] = (int8_t)fz; offsets[4 * (size_t)e + 3] = 0;
    }
    if (cell_offsets) {
      cell_offsets[3 * (size_t)e] = (float)fx; cell_offsets[3 * (size_t)e + 1] = (float)fy; cell_offsets[3 * (size_t)e + 2] = (float)fz;
    }
    if (edge_index) {
      edge_index[e] = a;
      edge_index[coo_stride + e] = b;
    }
  }
}

// ------------------------------------------------------------------------------------------
// COO -> CSR for center-sorted edge lists, and the transposed structure
// ------------------------------------------------------------------------------------------
__global__ void coo_to_csr_kernel(const long long* __restrict__ ei, const float* __restrict__ co, int n_nodes, int n_edges,
                                  int* __restrict__ rowptr, int* __restrict__ col, int8_t* __restrict__ offsets) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int c = (int)ei[e];
  const int prev = e > 0 ? (int)ei[e - 1] : -1;
  for (int nn = prev + 1; nn <= c; ++nn) rowptr[nn] = e;
  if (e == n_edges - 1)
    for (int nn = c + 1; nn <= n_nodes; ++nn) rowptr[nn] = n_edges;
  col[e] = (int)ei[(size_t)n_edges + e];
  if (offsets) {
    offsets[4 * (size_t)e + 0] = co ? (int8_t)rintf(co[3 * (size_t)e + 0]) : 0;
    offsets[4 * (size_t)e + 1] = co ? (int8_t)rintf(co[3 * (size_t)e + 1]) : 0;
    offsets[4 * (size_t)e + 2] = co ? (int8_t)rintf(co[3 * (size_t)e + 2]) : 0;
    offsets[4 * (size_t)e + 3] = 0;
  }
}

// capacity mode, after the fill pass: rows past the capacity become empty (the overflow flag is already raised)
__global__ void clamp_rowptr_kernel(int* __restrict__ rowptr, int n, int capacity) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rowptr[i] = min(rowptr[i], capacity);
}

// n_edges_dev points at rowptr[n_nodes]: the edge count stays on the device (capacity mode)
__global__ void count_cols_kernel(const int* __restrict__ col, const int* __restrict__ n_edges_dev, int capacity,
                                  int* __restrict__ cnt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < min(*n_edges_dev, capacity)) atomicAdd(&cnt[col[e]], 1);
}

// one warp per row i: lanes stride over the row's edges (coalesced reads of col); slot order inside a transposed row is
// whatever the atomics give -- the rows are sorted right after
__global__ void fill_transposed_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int n_nodes,
                                       int capacity, const int* __restrict__ t_rowptr, int* __restrict__ cursor,
                                       int* __restrict__ t_row, int* __restrict__ t_eid) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n_nodes) return;
  const int e1 = min(rowptr[i + 1], capacity);
  for (int e = rowptr[i] + lane; e < e1; e += 32) {
    const int j = col[e];
    const int slot = t_rowptr[j] + atomicAdd(&cursor[j], 1);
    t_row[slot] = i;
    t_eid[slot] = e;
  }
}

// slots of one transposed row sorted by edge id (== by center): deterministic order.  One warp per row: the row is
// staged in shared memory and every entry finds its rank by counting the smaller keys (edge ids are unique); rows
// longer than the staging buffer fall back to an insertion sort by one lane.
constexpr int SORT_CAP = 256;
__global__ void __launch_bounds__(128) sort_transposed_rows_kernel(const int* __restrict__ t_rowptr, int n_nodes, int* __restrict__ t_row,
                                                                   int* __restrict__ t_eid) {
  __shared__ int keys[4][SORT_CAP], vals[4][SORT_CAP];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 4 + w;
  if (j >= n_nodes) return;
  const int b = t_rowptr[j], n = t_rowptr[j + 1] - b;
  if (n <= 1) return;
  if (n <= SORT_CAP) {
    for (int k = lane; k < n; k += 32) {
      keys[w][k] = t_eid[b + k];
      vals[w][k] = t_row[b + k];
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      const int my = keys[w][k];
      int rank = 0;
      for (int m = 0; m < n; ++m) rank += keys[w][m] < my ? 1 : 0;
      t_eid[b + rank] = my;
      t_row[b + rank] = vals[w][k];
    }
  } else if (lane == 0) {
    const int e = b + n;
    for (int a = b + 1; a < e; ++a) {
      const int ke = t_eid[a], kr = t_row[a];
      int k = a - 1;
      while (k >= b && t_eid[k] > ke) {
        t_eid[k + 1] = t_eid[k];
        t_row[k + 1] = t_row[k];
        --k;
      }
      t_eid[k + 1] = ke;
      t_row[k + 1] = kr;
    }
  }
}

}  // namespace xeq

using namespace xeq;

extern "C" {

static inline size_t cell_max_bins(int n_nodes) { return (size_t)(n_nodes > 64 ? n_nodes : 64); }
constexpr int CELL_LIST_MIN_ATOMS = 256;

struct CellWs {
  CellGrid* grid;
  int *cnt, *bin_start, *scratch, *atom_bin, *bin_atoms;
  int max_bins;
};

// bins the wrapped positions (count -> scan -> fill); everything stays on the device
static int build_cells(Carver& cv, const float* pw, const float* cell, float cutoff, int n, CellWs* W, cudaStream_t st) {
  W->max_bins = (int)cell_max_bins(n);
  W->grid = cv.take<CellGrid>(1);
  W->cnt = cv.take<int>(W->max_bins + 1);
  W->bin_start = cv.take<int>(W->max_bins + 1);
  W->scratch = cv.take<int>(scan_scratch_ints(W->max_bins + 1));
  W->atom_bin = cv.take<int>(n);
  W->bin_atoms = cv.take<int>(n);
  cell_setup_kernel<<<1, 32, 0, st>>>(cell, cutoff, W->max_bins, W->grid);
  XEQ_CUDA(cudaMemsetAsync(W->cnt, 0, sizeof(int) * (size_t)(W->max_bins + 1), st));
  bin_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(pw, n, W->grid, W->cnt, W->atom_bin);
  XEQ_LAUNCHED(2);
  int rc = exclusive_scan(W->cnt, W->bin_start, W->max_bins, W->scratch, st);
  if (rc) return rc;
  XEQ_CUDA(cudaMemsetAsync(W->cnt, 0, sizeof(int) * (size_t)(W->max_bins + 1), st));
  bin_fill_kernel<<<(n + 255) / 256, 256, 0, st>>>(W->atom_bin, n, W->bin_start, W->cnt, W->bin_atoms);
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

static bool use_cell_list(bool periodic, int n, int G, const PbcParams& pp) {
  return periodic && G == 1 && n >= CELL_LIST_MIN_ATOMS && pp.pbc[0] && pp.pbc[1] && pp.pbc[2] && pp.rep[0] == 1 &&
         pp.rep[1] == 1 && pp.rep[2] == 1;
}

size_t xeq_radius_graph_workspace_bytes(int32_t n_nodes, int32_t n_graphs, int periodic) {
  (void)n_graphs;
  size_t b = 256;
  b += align_up(sizeof(int) * (size_t)(n_nodes + 1), 256);                 // degrees
  b += align_up(sizeof(int) * scan_scratch_ints(n_nodes + 1), 256);        // scan scratch
  if (periodic) {
    b += align_up(sizeof(float) * 3 * (size_t)n_nodes, 256);               // wrapped positions
    b += align_up(sizeof(int) * 3 * (size_t)n_nodes, 256);                 // integer shifts
    const size_t mb = cell_max_bins(n_nodes);                              // cell list (large fully periodic graphs)
    b += align_up(sizeof(CellGrid), 256);
    b += 2 * align_up(sizeof(int) * (mb + 1), 256);                        // bin counts / cursors, bin starts
    b += align_up(sizeof(int) * scan_scratch_ints((int)mb + 1), 256);
    b += 2 * align_up(sizeof(int) * (size_t)n_nodes, 256);                 // bin of each atom, atoms by bin
  }
  return b;
}

static int rg_common(const float* pos, int32_t n, const int32_t* graph_ptr, const int32_t* node_graph, int32_t G,
                     const float* cell, const int32_t* pbc_host, const int32_t* rep_host, float cutoff, void* ws,
                     size_t ws_bytes, bool* periodic, PbcParams* pp) {
  XEQ_CHECK_ARG(pos && graph_ptr && n >= 0 && G >= 1, "radius_graph: bad arguments");
  XEQ_CHECK_ARG(cutoff > 0.f, "radius_graph: cutoff must be positive");
  XEQ_CHECK_ARG(G == 1 || node_graph, "radius_graph: node_graph required for batched graphs");
  *periodic = cell != nullptr;
  for (int x = 0; x < 3; ++x) {
    pp->pbc[x] = (*periodic && pbc_host) ? (pbc_host[x] != 0) : 0;
    pp->rep[x] = (*periodic && rep_host && pp->pbc[x]) ? rep_host[x] : 0;
    XEQ_CHECK_ARG(pp->rep[x] >= 0 && pp->rep[x] <= 60, "radius_graph: image repeat %d out of range", pp->rep[x]);
  }
  XEQ_CHECK_ARG(ws && ws_bytes >= xeq_radius_graph_workspace_bytes(n, G, *periodic), "radius_graph: workspace too small");
  return XEQ_OK;
}

int xeq_radius_graph_count(const float* pos, int32_t n, const int32_t* graph_ptr, const int32_t* node_graph, int32_t G,
                           const float* cell, const int32_t* pbc_host, const int32_t* rep_host, float cutoff,
                           int32_t* rowptr, void* ws, size_t ws_bytes, xeq_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  bool periodic;
  PbcParams pp;
  int rc = rg_common(pos, n, graph_ptr, node_graph, G, cell, pbc_host, rep_host, cutoff, ws, ws_bytes, &periodic, &pp);
  if (rc) return rc;
  XEQ_CHECK_ARG(rowptr, "radius_graph_count: rowptr is NULL");
  Carver cv(ws);
  int* deg = cv.take<int>(n + 1);
  int* scratch = cv.take<int>(scan_scratch_ints(n + 1));
  if (n == 0) {
    XEQ_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int), st));
    return XEQ_OK;
  }
  const int blocks = (int)(((size_t)n * 32 + 255) / 256);
  if (periodic) {
    float* pw = cv.take<float>(3 * (size_t)n);
    int* shift = cv.take<int>(3 * (size_t)n);
    wrap_positions_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos, node_graph, cell, pp, n, pw, shift);
    if (use_cell_list(periodic, n, G, pp)) {
      CellWs W;
      rc = build_cells(cv, pw, cell, cutoff, n, &W, st);
      if (rc) return rc;
      cell_scan_kernel<false><<<blocks, 256, 0, st>>>(pw, cell, cutoff, n, W.grid, W.atom_bin, W.bin_start, W.bin_atoms, deg,
                                                      nullptr, nullptr, 0, nullptr);
    } else {
      graph_scan_kernel<true, false><<<blocks, 256, 0, st>>>(pw, shift, graph_ptr, node_graph, cell, pp, cutoff, n, deg,
                                                             nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, nullptr);
    }
  } else {
    graph_scan_kernel<false, false><<<blocks, 256, 0, st>>>(pos, nullptr, graph_ptr, node_graph, nullptr, pp, cutoff, n,
                                                            deg, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, nullptr);
  }
  XEQ_LAUNCHED(periodic ? 2 : 1);
  return exclusive_scan(deg, rowptr, n, scratch, st);
}

int xeq_radius_graph_fill(const float* pos, int32_t n, const int32_t* graph_ptr, const int32_t* node_graph, int32_t G,
                          const float* cell, const int32_t* pbc_host, const int32_t* rep_host, float cutoff,
                          int32_t* rowptr, int32_t* col, int8_t* offsets, int64_t* edge_index, float* cell_offsets,
                          int32_t edge_capacity, int32_t* overflow, void* ws, size_t ws_bytes, xeq_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  bool periodic;
  PbcParams pp;
  int rc = rg_common(pos, n, graph_ptr, node_graph, G, cell, pbc_host, rep_host, cutoff, ws, ws_bytes, &periodic, &pp);
  if (rc) return rc;
  XEQ_CHECK_ARG(rowptr && col, "radius_graph_fill: rowptr/col is NULL");
  if (n == 0) return XEQ_OK;
  // n_edges is needed for the COO layout; the caller sized the outputs from rowptr[n], re-read it here
  // only when the COO output is requested (tiny D2H; the Python layer already synchronised on it).
  long long n_edges = edge_capacity;  // COO row stride; capacity mode keeps the edge count on the device
  const int cap = edge_capacity > 0 ? edge_capacity : 0x7fffffff;
  if (edge_index && edge_capacity <= 0) {
    int e32 = 0;
    XEQ_CUDA(cudaMemcpyAsync(&e32, rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    XEQ_CUDA(cudaStreamSynchronize(st));
    n_edges = e32;
  }
  Carver cv(ws);
  (void)cv.take<int>(n + 1);
  (void)cv.take<int>(scan_scratch_ints(n + 1));
  const int blocks = (int)(((size_t)n * 32 + 255) / 256);
  if (periodic) {
    float* pw = cv.take<float>(3 * (size_t)n);
    int* shift = cv.take<int>(3 * (size_t)n);
    wrap_positions_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos, node_graph, cell, pp, n, pw, shift);
    if (use_cell_list(periodic, n, G, pp)) {
      CellWs W;
      rc = build_cells(cv, pw, cell, cutoff, n, &W, st);
      if (rc) return rc;
      cell_scan_kernel<true><<<blocks, 256, 0, st>>>(pw, cell, cutoff, n, W.grid, W.atom_bin, W.bin_start, W.bin_atoms, nullptr,
                                                     rowptr, col, cap, overflow);
      row_sort_decode_kernel<<<(n + 7) / 8, 256, 0, st>>>(rowptr, n, shift, cap, col, offsets, (long long*)edge_index,
                                                          cell_offsets, n_edges);
      XEQ_LAUNCHED(1);
    } else {
      graph_scan_kernel<true, true><<<blocks, 256, 0, st>>>(pw, shift, graph_ptr, node_graph, cell, pp, cutoff, n, nullptr,
                                                            rowptr, col, offsets, (long long*)edge_index, cell_offsets,
                                                            n_edges, cap, overflow);
    }
  } else {
    graph_scan_kernel<false, true><<<blocks, 256, 0, st>>>(pos, nullptr, graph_ptr, node_graph, nullptr, pp, cutoff, n,
                                                           nullptr, rowptr, col, nullptr, (long long*)edge_index, nullptr,
                                                           n_edges, cap, overflow);
  }
  XEQ_LAUNCHED(periodic ? 2 : 1);
  if (edge_capacity > 0) {
    // capacity mode: the dropped edges must not be reachable -- every consumer walks rowptr without a bound check
    clamp_rowptr_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(rowptr, n + 1, edge_capacity);
    XEQ_LAUNCHED(1);
  }
  return XEQ_OK;
}

int xeq_csr_from_sorted_coo(const int64_t* edge_index, const float* cell_offsets, int32_t n_nodes, int32_t n_edges,
                            int32_t* rowptr, int32_t* col, int8_t* offsets, xeq_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  XEQ_CHECK_ARG(rowptr && n_nodes >= 0 && n_edges >= 0, "csr_from_sorted_coo: bad arguments");
  if (n_edges == 0) {
    XEQ_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int) * (size_t)(n_nodes + 1), st));
    return XEQ_OK;
  }
  XEQ_CHECK_ARG(edge_index && col, "csr_from_sorted_coo: NULL edge_index/col");
  coo_to_csr_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>((const long long*)edge_index, cell_offsets, n_nodes, n_edges,
                                                           rowptr, col, offsets);
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

size_t xeq_csr_transpose_workspace_bytes(int32_t n_nodes, int32_t n_edges) {
  (void)n_edges;
  return 256 + align_up(sizeof(int) * (size_t)(n_nodes + 1), 256) + align_up(sizeof(int) * scan_scratch_ints(n_nodes + 1), 256);
}

int xeq_csr_transpose(const int32_t* rowptr, const int32_t* col, int32_t n_nodes, int32_t n_edges, int32_t* t_rowptr,
                      int32_t* t_row, int32_t* t_eid, void* ws, size_t ws_bytes, xeq_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  XEQ_CHECK_ARG(rowptr && t_rowptr && n_nodes >= 0 && n_edges >= 0, "csr_transpose: bad arguments");
  XEQ_CHECK_ARG(ws && ws_bytes >= xeq_csr_transpose_workspace_bytes(n_nodes, n_edges), "csr_transpose: workspace too small");
  Carver cv(ws);
  int* cnt = cv.take<int>(n_nodes + 1);
  int* scratch = cv.take<int>(scan_scratch_ints(n_nodes + 1));
  XEQ_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(n_nodes + 1), st));
  if (n_edges > 0) {
    XEQ_CHECK_ARG(col && t_row && t_eid, "csr_transpose: NULL col/t_row/t_eid");
    count_cols_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>(col, rowptr + n_nodes, n_edges, cnt);
    XEQ_LAUNCHED(1);
  }
  int rc = exclusive_scan(cnt, t_rowptr, n_nodes, scratch, st);
  if (rc) return rc;
  if (n_edges > 0 && n_nodes > 0) {
    XEQ_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(n_nodes + 1), st));
    fill_transposed_kernel<<<(n_nodes + 3) / 4, 128, 0, st>>>(rowptr, col, n_nodes, n_edges, t_rowptr, cnt, t_row,
                                                                  t_eid);
    sort_transposed_rows_kernel<<<(n_nodes + 3) / 4, 128, 0, st>>>(t_rowptr, n_nodes, t_row, t_eid);
    XEQ_LAUNCHED(2);
  }
  XEQ_LAUNCH_CHECK();
  return XEQ_OK;
}

}  // extern "C"
