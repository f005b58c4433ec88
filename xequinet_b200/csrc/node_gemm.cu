// K3  node-side dense contractions on the 5th-generation tensor cores (tcgen05, accumulators in TMEM).
//
// Replaces the nn.Linear calls of XPainnMessage.scalar_mlp / XPainnUpdate.update_mlp / dot_lin
// (nn/xpainn.py:111-115, 190, 195-199), the e3nn o3.Linear calls update_U / update_V
// (nn/xpainn.py:186-187, 211-212) and their first / second derivatives (grad-input and grad-weight
// GEMMs) -- the only dense contractions of the model (BASELINE.json north_star (3)).
//
// Numerics: the reference computes these in fp32 (torch default, allow_tf32 = False).  A single
// TF32 MMA keeps 10 mantissa bits and would break the 1e-5 energy tolerance, so every product is
// evaluated as the 3xTF32 split   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo   with
// a_hi = rna_tf32(a), a_lo = rna_tf32(a - a_hi): relative error ~2^-22 per product, fp32 accumulate.
//
// Structure of one CTA (256 threads): it owns a 128-row tile of op(A) and one column pass
// (<= 256 columns) of one problem of a grouped launch.  Per 32-wide K block all threads move the
// operand blocks global -> registers (128-bit loads) -> hi/lo split -> shared memory in the canonical
// K-major SWIZZLE_128B layout (the split needs the data in registers anyway, so the staging is done by
// the threads rather than by TMA; operands stored with K as the slow index are transposed on the way).
// One thread then issues 3 tcgen05.mma per 8-wide k step into the TMEM accumulator and commits to an
// mbarrier; two shared-memory stages let the loads of block k+1 overlap the MMAs of block k.  The
// epilogue reads TMEM with tcgen05.ld (lane = row), applies alpha / bias / SiLU and stores fp32.
#include "common.cuh"

namespace xeq {
namespace {

constexpr int GEMM_THREADS = 256;
constexpr int TILE_M = 128;
constexpr int MAX_PASS_N = 256;
constexpr int KB = 32;  // K block = one 128-byte swizzle row of tf32
constexpr int A_TILE_BYTES = TILE_M * 128;
constexpr int B_TILE_BYTES = MAX_PASS_N * 128;
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;  // hi + lo of both operands
constexpr int GEMM_SMEM = 2 * STAGE_BYTES + 1024 /* alignment slack */ + 64 /* barriers, tmem slot */ + MAX_PASS_N * 4 /* bias */;
// TMEM: nacc x [pn columns]: sums of a_hi*b_hi, k steps dealt round-robin;  then [pn]: sum of the correction terms
constexpr int MAX_PROBLEMS = 20;

struct Problem {
  const float* a;
  const float* b;
  const float* bias;
  float* c;
  int m, n, k;
  int lda, ldb, ldc;
  int a_trans, b_trans;
  float alpha;
  int act;
  int pass_n;      // columns per pass (multiple of 16)
  int pass_begin;  // first blockIdx.y of this problem
  size_t part_off; // float offset of this problem's split-K partials in the workspace
};

struct GemmArgs {
  Problem p[MAX_PROBLEMS];
  int n_problems;
  int split_k;
  int use_partials;  // split_k > 1, or several problems accumulate into one output
  float* partials;
};

// one output of the fixed-order reduction: `slabs` consecutive [m, n] partial slabs -> c (+ bias)
struct ReduceGroup {
  const float* part;
  const float* bias;
  float* c;
  int slabs, m, n, ldc;
};
struct ReduceArgs {
  ReduceGroup g[MAX_PROBLEMS];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// per-element split on the integer / FMA pipes (cvt.rna is quarter rate): hi = x rounded to 10 mantissa
// bits by an integer add, lo = x - hi exactly; the tensor core ignores the 13 low mantissa bits of lo
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  // lo = x - hi is exact (<= 13 significant bits); round it to the 11 the tensor core keeps instead of letting the
  // hardware truncate: hi + lo then represents x to 2^-23 (round to nearest) instead of 2^-22
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ uint32_t elect_one() {  // one lane of a converged warp
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred;
}

__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}

// Operand blocks travel global -> registers -> (hi, lo) -> shared memory.  All loads of a block are
// issued back to back (A_FRAG + B_FRAG independent 128-bit loads per thread) and the block for step
// i+1 is requested before the MMAs of step i are issued, so one memory latency is exposed per K block
// at most.
constexpr int A_FRAG = TILE_M * 8 / GEMM_THREADS;      // 4 float4 per thread
constexpr int B_FRAG = MAX_PASS_N * 8 / GEMM_THREADS;  // up to 8 float4 per thread

// operand stored with K contiguous: tile rows = MN index, 8 16-byte chunks per row
template <int NF>
__device__ __forceinline__ void load_k_contig(const float* __restrict__ g, int ld, int row0, int rows, int k0, int K,
                                              int tile_rows, float4 (&f)[NF]) {
#pragma unroll
  for (int it = 0; it < NF; ++it) {
    const int c = threadIdx.x + it * GEMM_THREADS;
    const int r = c >> 3, kc = c & 7;
    const int grow = row0 + r, k = k0 + kc * 4;
    f[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < tile_rows && grow < rows && k < K) f[it] = __ldg(reinterpret_cast<const float4*>(g + (size_t)grow * ld + k));
  }
}
template <int NF>
__device__ __forceinline__ void store_k_contig(uint32_t s_hi, uint32_t s_lo, int tile_rows, const float4 (&f)[NF]) {
#pragma unroll
  for (int it = 0; it < NF; ++it) {
    const int c = threadIdx.x + it * GEMM_THREADS;
    const int r = c >> 3, kc = c & 7;
    if (r < tile_rows) {
      uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
      split_fast(f[it].x, h0, l0);
      split_fast(f[it].y, h1, l1);
      split_fast(f[it].z, h2, l2);
      split_fast(f[it].w, h3, l3);
      const uint32_t off = (uint32_t)(r * 128 + ((kc ^ (r & 7)) << 4));
      sts_v4(s_hi + off, h0, h1, h2, h3);
      sts_v4(s_lo + off, l0, l1, l2, l3);
    }
  }
}

// operand stored with K as the slow index ([K, MN] row-major): lane <-> k, transposed on the way in.
// Shared-memory stores of one warp hit 32 distinct banks (one swizzled 128-byte row per store).
template <int NF>
__device__ __forceinline__ void load_k_strided(const float* __restrict__ g, int ld, int mn0, int MN, int k0, int K,
                                               int tile_rows, float4 (&f)[NF]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = k0 + lane;
#pragma unroll
  for (int it = 0; it < NF; ++it) {
    const int j = warp + it * (GEMM_THREADS / 32);
    const int mn = mn0 + 4 * j;
    f[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (4 * j < tile_rows && k < K && mn < MN) f[it] = __ldg(reinterpret_cast<const float4*>(g + (size_t)k * ld + mn));
  }
}
template <int NF>
__device__ __forceinline__ void store_k_strided(uint32_t s_hi, uint32_t s_lo, int tile_rows, const float4 (&f)[NF]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int it = 0; it < NF; ++it) {
    const int j = warp + it * (GEMM_THREADS / 32);
    if (4 * j < tile_rows) {
      const float vals[4] = {f[it].x, f[it].y, f[it].z, f[it].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * j + i;
        uint32_t h, l;
        split_fast(vals[i], h, l);
        const uint32_t off = (uint32_t)(r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
        sts_b32(s_hi + off, h);
        sts_b32(s_lo + off, l);
      }
    }
  }
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start address >> 4 in bits
// [0,14), LBO (unused for swizzled K-major) [16,30), SBO = 1024 B (8 rows x 128 B) [32,46),
// descriptor version 1 at bit 46, layout type SWIZZLE_128B = 2 at bits [61,64).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float silu(float x) { return x / (1.f + expf(-x)); }

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tf32x3_kernel(const __grid_constant__ GemmArgs args) {
  extern __shared__ unsigned char gemm_smem_raw[];
  pdl_trigger();  // the next kernel of the stream may be staged now (it waits for this grid before reading memory)
  // which problem / column pass
  int pi = 0;
  while (pi + 1 < args.n_problems && (int)blockIdx.y >= args.p[pi + 1].pass_begin) ++pi;
  const Problem& P = args.p[pi];
  const int m0 = blockIdx.x * TILE_M;
  if (m0 >= P.m) return;  // uniform per CTA
  const int pass = blockIdx.y - P.pass_begin;
  const int n0 = pass * P.pass_n;
  const int pn = P.pass_n;  // MMA N of this CTA (tail columns are zero-filled and not stored)
  // K range of this split
  const int nkb_total = (P.k + KB - 1) / KB;
  const int kb_per = (nkb_total + args.split_k - 1) / args.split_k;
  const int kb_begin = blockIdx.z * kb_per;
  const int kb_end = min(nkb_total, kb_begin + kb_per);
  const int nkb = kb_end - kb_begin;

  const uint32_t base = (smem_u32(gemm_smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + 2 * STAGE_BYTES, tmem_slot = bar0 + 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t bias_s = bar0 + 64;  // this pass's bias values (zeros when there is none)
  // nacc main accumulators + one for the correction terms, pn columns each; allocations are powers of two >= 32
  const int nacc = max(1, min(4, 512 / pn - 1));
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < (nacc + 1) * pn) tmem_cols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();  // everything above overlapped the tail of the previous kernel; its results are visible from here on
  if (tid < pn) {
    const int col = n0 + tid;
    const float bv = (P.bias && !args.use_partials && col < P.n) ? __ldg(P.bias + col) : 0.f;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * tid), "f"(bv) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  const uint32_t tmem_corr = tmem_base + (uint32_t)(nacc * pn);

  // instruction descriptor: D = f32 (bits 4-5 = 1), A = B = tf32 (2 at bits 7-9 / 10-12), both K-major,
  // N >> 3 at bits 17-22, M >> 4 at bits 24-28
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(pn >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

  // two register sets: the loads of blocks i+1 and i+2 are in flight while block i is converted and multiplied
  float4 fa0[A_FRAG], fb0[B_FRAG], fa1[A_FRAG], fb1[B_FRAG];
  auto load_block = [&](int i, float4 (&fa)[A_FRAG], float4 (&fb)[B_FRAG]) {
    const int k0 = (kb_begin + i) * KB;
    if (P.a_trans) load_k_strided(P.a, P.lda, m0, P.m, k0, P.k, TILE_M, fa);
    else load_k_contig(P.a, P.lda, m0, P.m, k0, P.k, TILE_M, fa);
    if (P.b_trans) load_k_contig(P.b, P.ldb, n0, P.n, k0, P.k, pn, fb);
    else load_k_strided(P.b, P.ldb, n0, P.n, k0, P.k, pn, fb);
  };
  auto process = [&](int i, float4 (&fa)[A_FRAG], float4 (&fb)[B_FRAG]) {
    const int s = i & 1;
    const uint32_t sa_hi = base + s * STAGE_BYTES, sa_lo = sa_hi + A_TILE_BYTES;
    const uint32_t sb_hi = sa_lo + A_TILE_BYTES, sb_lo = sb_hi + B_TILE_BYTES;
    if (i >= 2) mbar_wait(bar0 + 8 * s, (uint32_t)(((i >> 1) - 1) & 1));  // MMAs of block i-2 have read this stage
    const int k0 = (kb_begin + i) * KB;
    if (P.a_trans) store_k_strided(sa_hi, sa_lo, TILE_M, fa);
    else store_k_contig(sa_hi, sa_lo, TILE_M, fa);
    if (P.b_trans) store_k_contig(sb_hi, sb_lo, pn, fb);
    else store_k_strided(sb_hi, sb_lo, pn, fb);
    if (i + 2 < nkb) load_block(i + 2, fa, fb);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (warp == 0 && elect_one()) {  // warp 0 is converged here (elect.sync lets ptxas issue the MMAs back to back)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nks = min(KB / 8, (P.k - k0 + 7) / 8);
      for (int ks = 0; ks < nks; ++ks) {
        const uint64_t da_hi = smem_desc(sa_hi + ks * 32), da_lo = smem_desc(sa_lo + ks * 32);
        const uint64_t db_hi = smem_desc(sb_hi + ks * 32), db_lo = smem_desc(sb_lo + ks * 32);
        // the tensor core ROUNDS TOWARD ZERO when it adds a k step to the accumulator (measured: scratch/gemm_bias.py,
        // K = 704 sums of positive terms come out 4.8e-6 low = 88 steps x 2^-24): the main products rotate over
        // `nacc` accumulators (each sees 1 / nacc of the steps at 1 / nacc of the magnitude), the 2^-11 correction
        // terms share one; the epilogue adds them in fp32 with round-to-nearest
        const uint32_t acc = tmem_base + (uint32_t)((i * (KB / 8) + ks) % nacc) * (uint32_t)pn;
        const uint32_t first = (i * (KB / 8) + ks) < nacc ? 0u : 1u;
        mma_tf32(tmem_corr, da_lo, db_hi, idesc, (i | ks) ? 1u : 0u);
        mma_tf32(tmem_corr, da_hi, db_lo, idesc, 1u);
        mma_tf32(acc, da_hi, db_hi, idesc, first);
      }
      umma_commit(bar0 + 8 * s);  // implies tcgen05.fence::before_thread_sync
    }
  };
  if (nkb > 0) load_block(0, fa0, fb0);
  if (nkb > 1) load_block(1, fa1, fb1);
  for (int i = 0; i < nkb; i += 2) {
    process(i, fa0, fb0);
    if (i + 1 < nkb) process(i + 1, fa1, fb1);
  }
  if (nkb > 0) {
    const int last = nkb - 1;
    mbar_wait(bar0 + 8 * (last & 1), (uint32_t)((last >> 1) & 1));  // in-order pipe: everything before is done too
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31; the two warp groups alternate 16-column chunks
  const int row = m0 + 32 * (warp & 3) + lane;
  const bool split = args.use_partials != 0;
  float* crow = split ? args.partials + P.part_off + ((size_t)blockIdx.z * P.m + row) * P.n
                      : P.c + (size_t)row * P.ldc;
  const bool vec_ok = split ? (P.n % 4 == 0) : (P.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(P.c) & 15) == 0);
  for (int ch = warp >> 2; ch < pn / 16; ch += 2) {
    uint32_t r[16], rc[16];
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(ch * 16);
    if (nkb > 0) {
      const int ksteps = (min(P.k, kb_end * KB) - kb_begin * KB + 7) / 8;  // k steps this CTA issued
      const int nused = min(nacc, ksteps);                                   // accumulators that received one
      tmem_ld16(taddr, r);
      for (int a = 1; a <= nused; ++a) {  // a == nused: the correction accumulator
        tmem_ld16(a < nused ? taddr + (uint32_t)(a * pn) : taddr + (uint32_t)(nacc * pn), rc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(rc[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = 0u;
    }
    const int col0 = n0 + ch * 16;
    if (row < P.m && col0 < P.n) {
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float val = __uint_as_float(r[j]) * P.alpha;
        if (!split) {
          float bv;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bv) : "r"(bias_s + 4u * (uint32_t)(ch * 16 + j)));
          val += bv;
          if (P.act == 1) val = silu(val);
        }
        o[j] = val;
      }
      if (vec_ok && col0 + 16 <= P.n) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(crow + col0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < P.n) crow[col0 + j] = o[j];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// fixed-order reduction of the partial slabs (split-K and/or problems that share an output):
// deterministic, then bias.  blockIdx.y = output group.
__global__ void gemm_reduce_kernel(const __grid_constant__ ReduceArgs args) {
  const ReduceGroup& G = args.g[blockIdx.y];
  const size_t total = (size_t)G.m * G.n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / G.n), col = (int)(idx % G.n);
    // four interleaved partial sums (fixed order): four slab reads in flight instead of one
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int z = 0;
    for (; z + 3 < G.slabs; z += 4) {
      a0 += G.part[(size_t)z * total + idx];
      a1 += G.part[(size_t)(z + 1) * total + idx];
      a2 += G.part[(size_t)(z + 2) * total + idx];
      a3 += G.part[(size_t)(z + 3) * total + idx];
    }
    for (; z < G.slabs; ++z) a0 += G.part[(size_t)z * total + idx];
    float acc = (a0 + a1) + (a2 + a3);
    if (G.bias) acc += G.bias[col];
    G.c[(size_t)row * G.ldc + col] = acc;
  }
}

// columns per pass: at most MAX_PASS_N, and narrower when the launch would otherwise leave SMs idle (`fill` = how many
// times more CTAs the machine can hold: M = 5376, N = 128 is 42 CTAs of 128 x 128 -- three passes of 48 columns put
// 126 CTAs on the 148 SMs and cut the serial K loop + epilogue of every CTA)
int pass_width(int n, int fill) {
  const int n16 = (n + 15) / 16 * 16;
  int np = (n16 + MAX_PASS_N - 1) / MAX_PASS_N;
  np = max(np, min(np * max(fill, 1), n16 / 16));
  return ((n16 + np - 1) / np + 15) / 16 * 16;
}

int validate(const xeq_gemm_t* pr, int n, int split_k) {
  XEQ_CHECK_ARG(pr && n >= 1 && n <= MAX_PROBLEMS, "gemm: between 1 and %d problems per launch", MAX_PROBLEMS);
  XEQ_CHECK_ARG(split_k >= 1 && split_k <= 256, "gemm: split_k out of range");
  for (int i = 0; i < n; ++i) {
    const xeq_gemm_t& p = pr[i];
    XEQ_CHECK_ARG(p.a && p.b && p.c && p.m >= 0 && p.n >= 1 && p.k >= 1, "gemm[%d]: bad arguments", i);
    XEQ_CHECK_ARG(p.act == 0 || p.act == 1, "gemm[%d]: act must be 0 (none) or 1 (SiLU)", i);
    XEQ_CHECK_ARG(((uintptr_t)p.a & 15) == 0 && ((uintptr_t)p.b & 15) == 0 && p.lda % 4 == 0 && p.ldb % 4 == 0,
                  "gemm[%d]: operands must be 16-byte aligned with leading dimensions that are multiples of 4", i);
    if (p.a_trans) XEQ_CHECK_ARG(p.m % 4 == 0 && p.lda >= p.m, "gemm[%d]: transposed A needs m %% 4 == 0, lda >= m", i);
    else XEQ_CHECK_ARG(p.k % 4 == 0 && p.lda >= p.k, "gemm[%d]: A needs k %% 4 == 0, lda >= k", i);
    if (p.b_trans) XEQ_CHECK_ARG(p.k % 4 == 0 && p.ldb >= p.k, "gemm[%d]: transposed B needs k %% 4 == 0, ldb >= k", i);
    else XEQ_CHECK_ARG(p.n % 4 == 0 && p.ldb >= p.n, "gemm[%d]: B needs n %% 4 == 0, ldb >= n", i);
    XEQ_CHECK_ARG(p.ldc >= p.n, "gemm[%d]: ldc < n", i);
  }
  return XEQ_OK;
}

}  // namespace
}  // namespace xeq

using namespace xeq;

extern "C" {

static bool shares_output(const xeq_gemm_t& a, const xeq_gemm_t& b) {
  return a.c == b.c && a.m == b.m && a.n == b.n && a.ldc == b.ldc;
}
static bool needs_partials(const xeq_gemm_t* pr, int n, int split_k) {
  if (split_k > 1) return true;
  for (int i = 1; i < n; ++i)
    if (shares_output(pr[i - 1], pr[i])) return true;
  return false;
}

size_t xeq_gemm_workspace_bytes(const xeq_gemm_t* problems, int32_t n_problems, int32_t split_k) {
  if (!problems || n_problems < 1 || split_k < 1 || !needs_partials(problems, n_problems, split_k)) return 0;
  size_t total = 0;
  for (int i = 0; i < n_problems; ++i) total += (size_t)split_k * problems[i].m * problems[i].n * sizeof(float);
  return align_up(total, 256);
}

int xeq_gemm_tf32x3(const xeq_gemm_t* problems, int32_t n_problems, int32_t split_k, void* workspace,
                    size_t workspace_bytes, xeq_stream_t stream) {
  int rc = validate(problems, n_problems, split_k);
  if (rc) return rc;
  const bool partials = needs_partials(problems, n_problems, split_k);
  XEQ_CHECK_ARG(!partials || (workspace && workspace_bytes >= xeq_gemm_workspace_bytes(problems, n_problems, split_k)),
                "gemm: workspace too small (split_k = %d)", split_k);
  GemmArgs args;
  args.n_problems = n_problems;
  args.split_k = split_k;
  args.use_partials = partials ? 1 : 0;
  args.partials = static_cast<float*>(workspace);
  int passes = 0, max_m = 0;
  size_t part = 0;  // in floats; slabs of consecutive problems are contiguous (needed by the grouped reduction)
  long long base_ctas = 0;
  for (int i = 0; i < n_problems; ++i)
    base_ctas += (long long)((problems[i].m + TILE_M - 1) / TILE_M) * ((problems[i].n + MAX_PASS_N - 1) / MAX_PASS_N) * split_k;
  const int fill = base_ctas > 0 ? (int)(num_sms() / base_ctas) : 1;
  for (int i = 0; i < n_problems; ++i) {
    const xeq_gemm_t& p = problems[i];
    XEQ_CHECK_ARG(!partials || p.act == 0, "gemm[%d]: activation is not available with partial sums", i);
    Problem& q = args.p[i];
    q.a = p.a; q.b = p.b; q.bias = p.bias; q.c = p.c;
    q.m = p.m; q.n = p.n; q.k = p.k;
    q.lda = p.lda; q.ldb = p.ldb; q.ldc = p.ldc;
    q.a_trans = p.a_trans; q.b_trans = p.b_trans;
    q.alpha = p.alpha; q.act = p.act;
    q.pass_n = pass_width(p.n, fill);
    q.pass_begin = passes;
    q.part_off = part;
    passes += (p.n + q.pass_n - 1) / q.pass_n;
    part += (size_t)split_k * p.m * p.n;
    max_m = max(max_m, p.m);
  }
  if (max_m == 0) return XEQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  {  // per-device attribute: set on every launch (cheap)
    XEQ_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  }
  dim3 grid((max_m + TILE_M - 1) / TILE_M, passes, split_k);
  XEQ_CUDA(launch_pdl(gemm_tf32x3_kernel, grid, dim3(GEMM_THREADS), (size_t)GEMM_SMEM, st, args));
  XEQ_LAUNCHED(1);
  if (partials) {
    ReduceArgs red;
    int n_groups = 0;
    size_t max_total = 0;
    for (int i = 0; i < n_problems;) {
      int j = i + 1;
      while (j < n_problems && shares_output(problems[i], problems[j])) ++j;
      const Problem& q = args.p[i];
      ReduceGroup& G = red.g[n_groups++];
      G.part = args.partials + q.part_off;
      G.bias = q.bias;
      G.c = q.c;
      G.slabs = (j - i) * split_k;
      G.m = q.m; G.n = q.n; G.ldc = q.ldc;
      max_total = max(max_total, (size_t)q.m * q.n);
      i = j;
    }
    if (max_total) {
      const size_t rblocks = (max_total + 255) / 256;
      dim3 rgrid((unsigned)(rblocks < 1024 ? rblocks : 1024), n_groups);
      gemm_reduce_kernel<<<rgrid, 256, 0, st>>>(red);
      XEQ_LAUNCHED(1);
    }
  }
  return XEQ_OK;
}

}  // extern "C"
