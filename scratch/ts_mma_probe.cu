// Probe (scratch, not product): tcgen05.mma with the A operand in TMEM (written by tcgen05.st from
// registers), B in shared memory (K-major SWIZZLE_128B), 3xTF32 split -- the filter contraction
// D[h, e] = sum_k W[h, k] psi[e, k] of the fused edge kernels.  Checks numerics against fp64 and
// times back-to-back MMAs (N = 16 / 32) and TMEM reads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/ts_mma_probe scratch/ts_mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred;
}
constexpr int KP = 24;     // padded K (21 -> 24)
constexpr int A_HI = 0, A_LO = 24, D0 = 64;  // TMEM columns

// mode 0: numerics (one 128 x N x 24 product, 3xTF32, A in TMEM).  mode 1: MMA issue timing.  mode 2: LDTM timing.
template <int N>
__global__ void __launch_bounds__(256, 1) probe(const float* __restrict__ W, const float* __restrict__ psi, float* __restrict__ out,
                                                 long long* __restrict__ cyc, int mode, int iters) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t b_hi = base, b_lo = base + 8192, bar = base + 16384, slot = bar + 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot) : "memory");

  // A rows: thread (warp w < 4, lane) owns row 32 w + lane  -> tcgen05.st into its own TMEM lane
  if (warp < 4) {
    const int row = 32 * warp + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);
    for (int c8 = 0; c8 < KP / 8; ++c8) {
      uint32_t hi[8], lo[8];
      for (int j = 0; j < 8; ++j) split_tf32(W[row * KP + c8 * 8 + j], hi[j], lo[j]);
      tmem_st8(lane_base + A_HI + c8 * 8, hi);
      tmem_st8(lane_base + A_LO + c8 * 8, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  // B rows (edges) -> K-major SWIZZLE_128B tiles, hi and lo
  for (int idx = tid; idx < N * 32; idx += blockDim.x) {
    const int e = idx >> 5, k = idx & 31;
    const float v = (k < KP) ? psi[e * KP + k] : 0.f;
    uint32_t h, l;
    split_tf32(v, h, l);
    const uint32_t off = (uint32_t)(e * 128 + (((k >> 2) ^ (e & 7)) << 4) + (k & 3) * 4);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(b_hi + off), "r"(h) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(b_lo + off), "r"(l) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  auto issue_tile = [&](uint32_t dcol) {
    for (int ks = 0; ks < KP / 8; ++ks) {
      const uint64_t db_hi = smem_desc(b_hi + ks * 32), db_lo = smem_desc(b_lo + ks * 32);
      mma_ts(tmem + dcol, tmem + A_LO + ks * 8, db_hi, idesc, ks ? 1u : 0u);
      mma_ts(tmem + dcol, tmem + A_HI + ks * 8, db_lo, idesc, 1u);
      mma_ts(tmem + dcol, tmem + A_HI + ks * 8, db_hi, idesc, 1u);
    }
  };
  if (mode == 0) {
    if (warp == 0) { if (elect_one()) { issue_tile(D0); umma_commit(bar); } __syncwarp(); }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
      for (int c = 0; c < N; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + D0 + c, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[(32 * warp + lane) * N + c + j] = __uint_as_float(r[j]);
      }
    }
  } else if (mode == 1 || mode == 3) {  // 5 tiles x 9 MMAs per "chunk"; mode 3 interleaves the 5 independent accumulators
    long long t0 = clock64();
    if (warp == 0) {
      if (elect_one()) {
        for (int it = 0; it < iters; ++it) {
          if (it >= 2) mbar_wait(bar + 8 * (it & 1), (uint32_t)(((it >> 1) - 1) & 1));
          if (mode == 1) {
#pragma unroll
            for (int t = 0; t < 5; ++t) issue_tile(D0 + t * N);
          } else {
#pragma unroll
            for (int ks = 0; ks < KP / 8; ++ks) {
              const uint64_t db_hi = smem_desc(b_hi + ks * 32), db_lo = smem_desc(b_lo + ks * 32);
#pragma unroll
              for (int t = 0; t < 5; ++t) mma_ts(tmem + D0 + t * N, tmem + A_LO + ks * 8, db_hi, idesc, ks ? 1u : 0u);
#pragma unroll
              for (int t = 0; t < 5; ++t) mma_ts(tmem + D0 + t * N, tmem + A_HI + ks * 8, db_lo, idesc, 1u);
#pragma unroll
              for (int t = 0; t < 5; ++t) mma_ts(tmem + D0 + t * N, tmem + A_HI + ks * 8, db_hi, idesc, 1u);
            }
          }
          umma_commit(bar + 8 * (it & 1));
        }
        const int l0 = iters - 2, l1 = iters - 1;
        mbar_wait(bar + 8 * (l0 & 1), (uint32_t)((l0 >> 1) & 1));
        mbar_wait(bar + 8 * (l1 & 1), (uint32_t)((l1 >> 1) & 1));
        cyc[0] = clock64() - t0;
      }
      __syncwarp();
    }
  } else {  // LDTM: 7 warps each read 3 x (x4) per 4 edges, N columns per "chunk"
    if (warp == 0) { if (elect_one()) { for (int t = 0; t < 5; ++t) issue_tile(D0 + t * N); umma_commit(bar); } __syncwarp(); }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
    long long t0 = clock64();
    float acc = 0.f;
    if (warp < 7) {
      const uint32_t lb = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + D0;
      for (int it = 0; it < iters; ++it) {
        for (int c = 0; c < N; c += 4) {
          uint32_t r0[4], r1[4], r2[4];
          tmem_ld4(lb + c, r0);
          tmem_ld4(lb + N + c, r1);
          tmem_ld4(lb + 2 * N + c, r2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          for (int j = 0; j < 4; ++j) acc += __uint_as_float(r0[j]) + __uint_as_float(r1[j]) * __uint_as_float(r2[j]);
        }
      }
    }
    __syncthreads();
    if (tid == 0) cyc[0] = clock64() - t0;
    if (acc == 123.456f) out[0] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

template <int N>
void run() {
  std::vector<float> W(128 * KP), P(N * KP);
  srand(7);
  for (auto& x : W) x = (float)rand() / RAND_MAX - 0.5f;
  for (auto& x : P) x = (float)rand() / RAND_MAX - 0.5f;
  for (int r = 0; r < 128; ++r) for (int k = 21; k < KP; ++k) W[r * KP + k] = 0.f;
  float *dW, *dP, *dO; long long* dC;
  CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dP, P.size() * 4)); CK(cudaMalloc(&dO, 128 * N * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, 128 * N * 4));
  const int smem = 16384 + 1024 + 64;
  CK(cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<N><<<1, 256, smem>>>(dW, dP, dO, dC, 0, 1);
  CK(cudaDeviceSynchronize());
  std::vector<float> O(128 * N);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 128; ++r) for (int e = 0; e < N; ++e) {
    double ref = 0;
    for (int k = 0; k < KP; ++k) ref += (double)W[r * KP + k] * (double)P[e * KP + k];
    maxerr = fmax(maxerr, fabs(ref - O[r * N + e])); maxref = fmax(maxref, fabs(ref));
  }
  printf("N=%d numerics: max abs err %.3e (max |ref| %.3f)  %s\n", N, maxerr, maxref, maxerr < 2e-6 ? "OK" : "MISMATCH");
  for (int mode = 1; mode <= 3; ++mode) {
    const int iters = 2000;
    probe<N><<<1, 256, smem>>>(dW, dP, dO, dC, mode, iters);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
    if (mode != 2) printf("N=%d MMA(mode %d): %.1f cyc per chunk of %d edges (45 MMAs) = %.2f cyc/edge\n", N, mode, (double)c / iters, N, (double)c / iters / N);
    else printf("N=%d LDTM: %.1f cyc per chunk (7 warps x 3 tiles) = %.2f cyc/edge\n", N, (double)c / iters, (double)c / iters / N);
  }
}

int main() {
  run<16>();
  run<32>();
  run<48>();
  return 0;
}
