// A host process WITHOUT Python or torch that evaluates XPaiNN energies and forces through libxeq_b200.so: what an
// MD engine's pair style does per step (the reference's counterpart is a TorchScript archive run through libtorch,
// xequinet/run/jit_script.py:28-86, xequinet/interface/jit_model.py:12-89).
//
//   md_host model.xeqw structure.bin out.bin [n_steps] [graph]
//     model.xeqw     written by xequinet_b200.runtime.NativeModel(model).save()
//     structure.bin  int32 N | float32 pos[N,3] | int32 Z[N]          (one non-periodic structure)
//     out.bin        float32 E | float32 e_atom[N] | float32 F[N,3]   (+ ms per evaluation on stdout)
//     graph          capture the evaluation (two-stream schedule) once as a CUDA graph and replay it n_steps times --
//                    what an MD loop does between neighbour-list rebuilds: the call allocates nothing and never syncs
//
// build: nvcc -o md_host examples/md_host.cpp -Iinclude -Lxequinet_b200 -lxeq_b200  (or g++ with -lcudart)
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "xeq_b200.h"

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); exit(2); } \
  } while (0)
#define XQ(call)                                                                     \
  do {                                                                               \
    if ((call) != XEQ_OK) { fprintf(stderr, "%s: %s\n", #call, xeq_last_error()); exit(3); } \
  } while (0)

template <typename T>
static T* dev_alloc(size_t n) {
  void* p = nullptr;
  CU(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
  return static_cast<T*>(p);
}
static void read_exact(FILE* f, void* dst, size_t bytes, const char* what) {
  if (fread(dst, 1, bytes, f) != bytes) { fprintf(stderr, "short read: %s\n", what); exit(4); }
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s model.xeqw structure.bin out.bin [n_steps]\n", argv[0]); return 1; }
  const int n_steps = argc > 4 ? atoi(argv[4]) : 1;
  const bool use_graph = argc > 5 && strcmp(argv[5], "graph") == 0;

  // ---- model: header + weight blob -> device, handle ----
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  char magic[8];
  int32_t h[9];
  float cutoff;
  uint64_t n_w;
  read_exact(f, magic, 8, "magic");
  if (memcmp(magic, "XEQW0001", 8) != 0) { fprintf(stderr, "not an XEQW0001 file\n"); return 1; }
  read_exact(f, h, sizeof(h), "header");
  read_exact(f, &cutoff, 4, "cutoff");
  read_exact(f, &n_w, 8, "n_weights");
  std::vector<float> w_host(n_w);
  read_exact(f, w_host.data(), n_w * 4, "weights");
  fclose(f);
  xeq_dims_t dims = {h[0], h[1], h[2], h[3], h[4], cutoff};
  float* w_dev = dev_alloc<float>(n_w);
  CU(cudaMemcpy(w_dev, w_host.data(), n_w * 4, cudaMemcpyHostToDevice));
  xeq_model_t* model = nullptr;
  XQ(xeq_model_create(&dims, h[5], h[6], h[7], h[8], w_dev, (size_t)n_w, &model));

  // ---- structure ----
  f = fopen(argv[2], "rb");
  if (!f) { perror(argv[2]); return 1; }
  int32_t N;
  read_exact(f, &N, 4, "N");
  std::vector<float> pos(3 * (size_t)N);
  std::vector<int32_t> Z(N);
  read_exact(f, pos.data(), pos.size() * 4, "pos");
  read_exact(f, Z.data(), Z.size() * 4, "Z");
  fclose(f);
  float* pos_d = dev_alloc<float>(3 * (size_t)N);
  int32_t* z_d = dev_alloc<int32_t>(N);
  int32_t* ptr_d = dev_alloc<int32_t>(2);
  const int32_t ptr_h[2] = {0, N};
  CU(cudaMemcpy(pos_d, pos.data(), pos.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(z_d, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(ptr_d, ptr_h, sizeof(ptr_h), cudaMemcpyHostToDevice));
  cudaStream_t st;
  CU(cudaStreamCreate(&st));

  // ---- neighbour list (K1; an engine would pass its own list through xeq_csr_from_sorted_coo instead) ----
  const int32_t pbc[3] = {0, 0, 0}, rep[3] = {0, 0, 0};
  const size_t k1_bytes = xeq_radius_graph_workspace_bytes(N, 1, 0);
  char* k1_ws = dev_alloc<char>(k1_bytes);
  int32_t* rowptr = dev_alloc<int32_t>(N + 1);
  XQ(xeq_radius_graph_count(pos_d, N, ptr_d, nullptr, 1, nullptr, pbc, rep, cutoff, rowptr, k1_ws, k1_bytes, st));
  int32_t E = 0;
  CU(cudaMemcpyAsync(&E, rowptr + N, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));  // the one host sync: sizes the edge arrays
  int32_t *col = dev_alloc<int32_t>(E), *t_rowptr = dev_alloc<int32_t>(N + 1), *t_row = dev_alloc<int32_t>(E),
          *t_eid = dev_alloc<int32_t>(E);
  XQ(xeq_radius_graph_fill(pos_d, N, ptr_d, nullptr, 1, nullptr, pbc, rep, cutoff, rowptr, col, nullptr, nullptr, nullptr, 0,
                           nullptr, k1_ws, k1_bytes, st));
  const size_t t_bytes = xeq_csr_transpose_workspace_bytes(N, E);
  char* t_ws = dev_alloc<char>(t_bytes);
  XQ(xeq_csr_transpose(rowptr, col, N, E, t_rowptr, t_row, t_eid, t_ws, t_bytes, st));
  const int tc = xeq_center_tile_edges(), tn = xeq_neighbor_tile_edges();
  const int n_tiles = xeq_csr_tile_count(N, E, tc), t_n_tiles = xeq_csr_tile_count(N, E, tn);
  int32_t *tile_ptr = dev_alloc<int32_t>(n_tiles + 1), *t_tile_ptr = dev_alloc<int32_t>(t_n_tiles + 1);
  XQ(xeq_csr_tile_bounds(rowptr, N, E, tc, tile_ptr, st));
  XQ(xeq_csr_tile_bounds(t_rowptr, N, E, tn, t_tile_ptr, st));
  xeq_graph_t g;
  memset(&g, 0, sizeof(g));
  g.n_nodes = N; g.n_edges = E; g.n_graphs = 1;
  g.rowptr = rowptr; g.col = col; g.t_rowptr = t_rowptr; g.t_row = t_row; g.t_eid = t_eid;
  g.tile_ptr = tile_ptr; g.t_tile_ptr = t_tile_ptr; g.n_tiles = n_tiles; g.t_n_tiles = t_n_tiles; g.tile_mode = 0;

  // ---- energy + forces ----
  const size_t ws_bytes = xeq_model_workspace_bytes(model, &g, 1);
  char* ws = dev_alloc<char>(ws_bytes);
  float *e_d = dev_alloc<float>(1), *ea_d = dev_alloc<float>(N), *f_d = dev_alloc<float>(3 * (size_t)N);
  XQ(xeq_model_energy_forces(model, &g, pos_d, z_d, ptr_d, e_d, ea_d, f_d, ws, ws_bytes, st));  // warm-up
  CU(cudaStreamSynchronize(st));
  cudaGraphExec_t exec = nullptr;
  if (use_graph) {
    cudaStream_t aux;
    CU(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
    cudaGraph_t graph;
    CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    XQ(xeq_model_energy_forces_mt(model, &g, pos_d, z_d, ptr_d, e_d, ea_d, f_d, ws, ws_bytes, st, aux));
    CU(cudaStreamEndCapture(st, &graph));
    CU(cudaGraphInstantiate(&exec, graph, 0));
    CU(cudaGraphLaunch(exec, st));  // warm-up replay
    CU(cudaStreamSynchronize(st));
  }
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n_steps; ++i) {
    if (exec) CU(cudaGraphLaunch(exec, st));
    else XQ(xeq_model_energy_forces(model, &g, pos_d, z_d, ptr_d, e_d, ea_d, f_d, ws, ws_bytes, st));
  }
  CU(cudaStreamSynchronize(st));
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / n_steps;

  float e_h;
  std::vector<float> ea_h(N), f_h(3 * (size_t)N);
  CU(cudaMemcpy(&e_h, e_d, 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ea_h.data(), ea_d, ea_h.size() * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(f_h.data(), f_d, f_h.size() * 4, cudaMemcpyDeviceToHost));
  f = fopen(argv[3], "wb");
  if (!f) { perror(argv[3]); return 1; }
  fwrite(&e_h, 4, 1, f);
  fwrite(ea_h.data(), 4, ea_h.size(), f);
  fwrite(f_h.data(), 4, f_h.size(), f);
  fclose(f);
  printf("atoms %d edges %d energy %.6f ms_per_evaluation %.4f (%s) kernels %lld\n", N, E, e_h, ms,
         exec ? "CUDA graph replay, two streams" : "eager launches", xeq_launch_count());
  xeq_model_destroy(model);
  return 0;
}
