// Library plumbing (version, thread-local error string, device info) and the small node-level
// helpers of the C ABI: contiguous segment sum (nn/output.py:124) and e3nn <-> cm layout maps.
#include <stdarg.h>

#include <atomic>

#include <cooperative_groups.h>

#include "common.cuh"

namespace xeq {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("XEQ_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

// one warp per segment, lanes stride over the segment, fixed-order butterfly => deterministic
__global__ void segment_sum_kernel(const float* __restrict__ src, const int* __restrict__ ptr, int n_seg,
                                   float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (seg >= n_seg) return;
  float acc = 0.f;
  for (int i = ptr[seg] + lane; i < ptr[seg + 1]; i += 32) acc += src[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[seg] = acc;
}

// out[c] = sum_r src[r, c].  A cluster of 8 CTAs owns 32 columns: CTA k sums its eighth of the rows (lane =
// column: 128-byte coalesced reads, four rows in flight per warp, the 16 warp partials added in fixed order),
// then CTA 0 adds the 8 partials through distributed shared memory in fixed order => deterministic, one launch,
// no workspace.  Bias gradients of the Linear layers (sum of the output gradient over the nodes).
constexpr int COLSUM_CLUSTER = 8, COLSUM_WARPS = 16;
__global__ void __cluster_dims__(1, COLSUM_CLUSTER, 1) __launch_bounds__(COLSUM_WARPS * 32)
    colsum_kernel(const float* __restrict__ src, const float* __restrict__ wgt, int n_rows, int n_cols, int ld, float* __restrict__ out) {
  namespace cg = cooperative_groups;
  __shared__ float part[COLSUM_WARPS][33];
  __shared__ float tot[32];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int chunk = (n_rows + COLSUM_CLUSTER - 1) / COLSUM_CLUSTER;
  const int r_end = min(n_rows, (rank + 1) * chunk);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (col < n_cols) {
    int r = rank * chunk + warp;
    if (wgt == nullptr) {
      for (; r + 3 * COLSUM_WARPS < r_end; r += 4 * COLSUM_WARPS) {
        a0 += src[(size_t)r * ld + col];
        a1 += src[(size_t)(r + COLSUM_WARPS) * ld + col];
        a2 += src[(size_t)(r + 2 * COLSUM_WARPS) * ld + col];
        a3 += src[(size_t)(r + 3 * COLSUM_WARPS) * ld + col];
      }
      for (; r < r_end; r += COLSUM_WARPS) a0 += src[(size_t)r * ld + col];
    } else {  // rows weighted by wgt[r] (same order of additions)
      for (; r + 3 * COLSUM_WARPS < r_end; r += 4 * COLSUM_WARPS) {
        a0 = fmaf(wgt[r], src[(size_t)r * ld + col], a0);
        a1 = fmaf(wgt[r + COLSUM_WARPS], src[(size_t)(r + COLSUM_WARPS) * ld + col], a1);
        a2 = fmaf(wgt[r + 2 * COLSUM_WARPS], src[(size_t)(r + 2 * COLSUM_WARPS) * ld + col], a2);
        a3 = fmaf(wgt[r + 3 * COLSUM_WARPS], src[(size_t)(r + 3 * COLSUM_WARPS) * ld + col], a3);
      }
      for (; r < r_end; r += COLSUM_WARPS) a0 = fmaf(wgt[r], src[(size_t)r * ld + col], a0);
    }
  }
  part[warp][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (warp == 0) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < COLSUM_WARPS; ++w) acc += part[w][lane];
    tot[lane] = acc;
  }
  cluster.sync();
  if (rank == 0 && warp == 0 && col < n_cols) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < COLSUM_CLUSTER; ++k) acc += *cluster.map_shared_rank(&tot[lane], k);
    out[col] = acc;
  }
  cluster.sync();  // the partials of the other CTAs stay alive until CTA 0 has read them
}

// y[r] = sum_c x[r, c] w[c] (+ bias[0]): one warp per row, lanes stride over the columns, fixed-order butterfly
__global__ void rowdot_kernel(const float* __restrict__ x, int ld, const float* __restrict__ w, const float* __restrict__ bias,
                              int n_rows, int n_cols, float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const float* xr = x + (size_t)row * ld;
  float acc = 0.f;
  for (int c = lane; c < n_cols; c += 32) acc = fmaf(xr[c], w[c], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[row] = acc + (bias ? bias[0] : 0.f);
}

// out[r, c] = g[r] w[c]
__global__ void outer_kernel(const float* __restrict__ g, const float* __restrict__ w, int n_rows, int n_cols, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_rows * n_cols) return;
  const int r = (int)(idx / n_cols), c = (int)(idx % n_cols);
  out[idx] = g[r] * w[c];
}

__global__ void layout_convert_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int m0, int m1,
                                      int m2, int direction) {
  pdl_trigger();
  pdl_wait();
  const int D = m0 + 3 * m1 + 5 * m2;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * D) return;
  const int node = (int)(idx / D), p = (int)(idx % D);  // p indexes the cm layout
  int e3;                                               // matching index in the e3nn layout
  if (p < m0) e3 = p;
  else if (p < m0 + 3 * m1) { const int r = p - m0, m = r / m1, u = r % m1; e3 = m0 + u * 3 + m; }
  else { const int r = p - m0 - 3 * m1, m = r / m2, u = r % m2; e3 = m0 + 3 * m1 + u * 5 + m; }
  if (direction == 0) dst[(size_t)node * D + p] = src[(size_t)node * D + e3];
  else dst[(size_t)node * D + e3] = src[(size_t)node * D + p];
}

}  // namespace xeq

using namespace xeq;

extern "C" {

int xeq_version(void) { return 100; }
const char* xeq_last_error(void) { return g_err; }
int xeq_num_sms(void) { return num_sms(); }
long long xeq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int xeq_segment_sum(const float* src, const int32_t* seg_ptr, int32_t n_segments, float* out, xeq_stream_t stream) {
  XEQ_CHECK_ARG(seg_ptr && out && n_segments >= 0, "segment_sum: bad arguments");
  if (n_segments == 0) return XEQ_OK;
  XEQ_CHECK_ARG(src, "segment_sum: src is NULL");
  const int blocks = (int)(((size_t)n_segments * 32 + 255) / 256);
  XEQ_CUDA(launch_pdl(segment_sum_kernel, dim3(blocks), dim3(256), (size_t)(0), (cudaStream_t)stream, src, seg_ptr, n_segments, out));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_colsum(const float* src, int32_t n_rows, int32_t n_cols, int32_t ld, float* out, xeq_stream_t stream) {
  XEQ_CHECK_ARG(out && n_rows >= 0 && n_cols >= 0 && ld >= n_cols, "colsum: bad arguments");
  if (n_cols == 0) return XEQ_OK;
  XEQ_CHECK_ARG(src || n_rows == 0, "colsum: src is NULL");
  colsum_kernel<<<dim3((n_cols + 31) / 32, COLSUM_CLUSTER), COLSUM_WARPS * 32, 0, (cudaStream_t)stream>>>(src, nullptr, n_rows, n_cols, ld, out);
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_colsum_weighted(const float* src, const float* row_weight, int32_t n_rows, int32_t n_cols, int32_t ld, float* out,
                        xeq_stream_t stream) {
  XEQ_CHECK_ARG(out && n_rows >= 0 && n_cols >= 0 && ld >= n_cols, "colsum_weighted: bad arguments");
  if (n_cols == 0) return XEQ_OK;
  XEQ_CHECK_ARG((src && row_weight) || n_rows == 0, "colsum_weighted: src / row_weight is NULL");
  colsum_kernel<<<dim3((n_cols + 31) / 32, COLSUM_CLUSTER), COLSUM_WARPS * 32, 0, (cudaStream_t)stream>>>(src, row_weight, n_rows, n_cols, ld, out);
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_rowdot(const float* x, int32_t ld, const float* w, const float* bias, int32_t n_rows, int32_t n_cols, float* y,
               xeq_stream_t stream) {
  XEQ_CHECK_ARG(n_rows >= 0 && n_cols >= 0 && ld >= n_cols, "rowdot: bad arguments");
  if (n_rows == 0) return XEQ_OK;
  XEQ_CHECK_ARG(y && (n_cols == 0 || (x && w)), "rowdot: NULL operand");
  const int blocks = (int)(((size_t)n_rows * 32 + 255) / 256);
  XEQ_CUDA(launch_pdl(rowdot_kernel, dim3(blocks), dim3(256), (size_t)(0), (cudaStream_t)stream, x, ld, w, bias, n_rows, n_cols, y));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_outer(const float* g, const float* w, int32_t n_rows, int32_t n_cols, float* out, xeq_stream_t stream) {
  XEQ_CHECK_ARG(n_rows >= 0 && n_cols >= 0, "outer: bad arguments");
  const size_t total = (size_t)n_rows * n_cols;
  if (total == 0) return XEQ_OK;
  XEQ_CHECK_ARG(g && w && out, "outer: NULL operand");
  XEQ_CUDA(launch_pdl(outer_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, g, w, n_rows, n_cols, out));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

int xeq_layout_convert(const float* src, float* dst, int32_t n_nodes, const xeq_dims_t* dims, int direction,
                       xeq_stream_t stream) {
  XEQ_CHECK_ARG(src && dst && dims && n_nodes >= 0 && src != dst, "layout_convert: bad arguments");
  XEQ_CHECK_ARG(direction == 0 || direction == 1, "layout_convert: direction must be 0 or 1");
  const size_t total = (size_t)n_nodes * (dims->mul0 + 3 * dims->mul1 + 5 * dims->mul2);
  if (total == 0) return XEQ_OK;
  XEQ_CUDA(launch_pdl(layout_convert_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, src, dst, n_nodes, dims->mul0,
                                                                                         dims->mul1, dims->mul2, direction));
  XEQ_LAUNCHED(1);
  return XEQ_OK;
}

}  // extern "C"
