// tcgen05 (5th-gen tensor core) GEMM for the memory update / embedding path:
//     C[M,N] (=|+=) op(A)[M,K] * op(B)[K,N] (+ bias),  fp32 in, fp32 accumulate in TMEM
// used for the GRU gate GEMMs (torch.nn.GRUCell at reference
// modules/memory_module.py:72,172), the TransformerConv projections
// (modules/emb_module.py:21-23,29) and their gradients.
//
// Precision modes
//   1  kind::tf32, operands rounded to tf32 (10-bit mantissa)      -> ~1e-3 relative
//   3  "3xTF32": a = a_hi + a_lo, b = b_hi + b_lo (both tf32), D += a_hi b_hi + a_hi b_lo
//      + a_lo b_hi                                                  -> fp32-level (~1e-6)
//
// Structure (one CTA = 256 threads = one 128 x 128 output tile, cta_group::1):
//   * operand tiles are [128 rows x 32 floats]: one row = 128 bytes = one swizzle span, laid
//     out in shared memory in the UMMA K-major SWIZZLE_128B layout (8-row x 128-byte atoms,
//     16-byte chunk index XOR row%8).  The CTA's warps build the tiles straight from global
//     memory with fully coalesced loads (a warp reads one 128-byte row per instruction, or 32
//     consecutive rows of one k for the transposed storage orders), so a row-gathered A
//     (memory[n_id]) and both storage orders of A/B (the gradient GEMMs are TN / NN) need no
//     separate pass; the tf32 hi/lo split happens on the way into shared memory
//   * 3-stage ring of operand tiles; warps 0-3 stage A, warps 4-7 stage B; the global loads
//     of k-blocks i+1 and i+2 are in flight in registers while the tensor core works on
//     k-block i; stage reuse is gated by tcgen05.commit -> mbarrier
//   * the accumulator lives in TMEM (128 lanes x 128 fp32 columns); the epilogue reads it
//     with tcgen05.ld 32x32b.x32, transposes 32x32 blocks through shared memory and writes
//     (or atomically adds, for split-K) whole 128-byte row segments with the bias
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

constexpr int TBM = 128, TBN = 128, TBK = 32;  // TBK floats = 128 bytes = one swizzle span
constexpr int kTileFloats = TBM * TBK;         // per operand per precision part (16 KB)
constexpr int kStageFloats = 4 * kTileFloats;  // A_hi, A_lo, B_hi, B_lo
constexpr int kStages = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// UMMA shared-memory descriptor, K-major SWIZZLE_128B: 8-row atoms are 1024 bytes apart
// (stride byte offset); the leading byte offset is unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // leading byte offset (ignored), 16 B
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // layout type: SWIZZLE_128B
  return d;
}

// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = TBN
__device__ __forceinline__ uint32_t umma_idesc() {
  uint32_t d = 0;
  d |= 1u << 4;                     // c_format = F32
  d |= 2u << 7;                     // a_format = TF32
  d |= 2u << 10;                    // b_format = TF32
  d |= (uint32_t)(TBN >> 3) << 17;  // n_dim
  d |= (uint32_t)(TBM >> 4) << 24;  // m_dim
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct TcArgs {
  const float* a;
  const int64_t* a_rows;
  const float* b;
  const float* bias;
  float* c;
  DevCount m, k;
  int n, lda, ldb, ldc, accumulate, split_k, prec;
};

// physical float offset of element (row, k) inside a [128 x 32] SWIZZLE_128B tile
__device__ __forceinline__ int sw128(int row, int k) {
  return row * TBK + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3));
}

// One operand tile in registers: 32 values per thread (4 warps cover one [128 x 32] tile).
//  K-contiguous storage: value i = element (row = warp*32 + i, k = lane)
//  row-contiguous storage: value i = element (row = (i&3)*32 + lane, k = warp*8 + (i>>2))
struct TileRegs {
  float v[32];
};

// `full` (warp-uniform): the whole tile is in range -> plain loads, one 64-bit add per load.
template <bool TRANS>
__device__ __forceinline__ void load_tile(TileRegs& r, const float* __restrict__ base,
                                          const int64_t* __restrict__ rows, int row0, int nrows,
                                          int ld, int k0, int kend, int warp, int lane) {
  if (!TRANS) {
    const int k = k0 + lane;
    const int rbase = row0 + warp * 32;
    if (rows != nullptr) {
      // gathered rows: lane i fetches the id of row i once, then it is broadcast per load
      const int myrow = rbase + lane;
      const long long my_gr = myrow < nrows ? rows[myrow] : 0;
      const bool k_ok = k < kend;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const long long gr = __shfl_sync(0xffffffffu, my_gr, i);
        r.v[i] = (k_ok && rbase + i < nrows) ? base[gr * ld + k] : 0.f;
      }
    } else if (rbase + 32 <= nrows && k0 + TBK <= kend) {
      const float* p = base + (long long)rbase * ld + k;
#pragma unroll
      for (int i = 0; i < 32; ++i) r.v[i] = p[(long long)i * ld];
    } else {
      const bool k_ok = k < kend;
      const float* p = base + (long long)rbase * ld + k;
#pragma unroll
      for (int i = 0; i < 32; ++i) r.v[i] = (k_ok && rbase + i < nrows) ? p[(long long)i * ld] : 0.f;
    }
  } else {
    const int kb = k0 + warp * 8;
    const float* p = base + (long long)kb * ld + row0 + lane;
    if (row0 + 128 <= nrows && kb + 8 <= kend) {
#pragma unroll
      for (int i = 0; i < 32; ++i) r.v[i] = p[(long long)(i >> 2) * ld + (i & 3) * 32];
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const bool ok = kb + (i >> 2) < kend && row0 + (i & 3) * 32 + lane < nrows;
        r.v[i] = ok ? p[(long long)(i >> 2) * ld + (i & 3) * 32] : 0.f;
      }
    }
  }
}

template <bool TRANS>
__device__ __forceinline__ void store_tile(const TileRegs& r, float* hi, float* lo, bool split,
                                           int warp, int lane) {
  if (!TRANS) {
    // row = warp*32 + i (row & 7 == i & 7), k = lane: 8 swizzled column offsets, reused
    int col[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) col[j] = (((lane >> 2) ^ j) << 2) | (lane & 3);
    float* h0 = hi + warp * 32 * TBK;
    float* l0 = lo + warp * 32 * TBK;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float x = r.v[i];
      const float h = to_tf32(x);
      h0[i * TBK + col[i & 7]] = h;
      if (split) l0[i * TBK + col[i & 7]] = x - h;
    }
  } else {
    // row = (i&3)*32 + lane (row & 7 == lane & 7), k = warp*8 + (i>>2)
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int row = (i & 3) * 32 + lane;
      const int k = warp * 8 + (i >> 2);
      const int off = row * TBK + ((((k >> 2) ^ (lane & 7)) << 2) | (k & 3));
      const float x = r.v[i];
      const float h = to_tf32(x);
      hi[off] = h;
      if (split) lo[off] = x - h;
    }
  }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256, 1) tc_gemm_kernel(TcArgs g) {
  pdl_wait();
  pdl_launch();
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t s_bar[kStages];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lw = warp & 3;        // row group of the operand tile / TMEM lane group
  const bool is_b = warp >= 4;    // warps 0-3 stage A, warps 4-7 stage B
  const int M = g.m.get(), K = g.k.get();
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
  if (m0 >= M) return;
  const int kchunk = ((K + g.split_k - 1) / g.split_k + TBK - 1) / TBK * TBK;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);
  const int nk = kend > kbeg ? (kend - kbeg + TBK - 1) / TBK : 0;
  const bool split = g.prec == 3;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kStages; ++i) mbar_init(&s_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&s_tmem)),
                 "r"(TBN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  // register prefetch, two k-blocks deep
  TileRegs r[2];
  auto load_kb = [&](TileRegs& dst, int kb) {
    const int k0 = kbeg + kb * TBK;
    if (!is_b) load_tile<TA>(dst, g.a, g.a_rows, m0, M, g.lda, k0, kend, lw, lane);
    else load_tile<TB>(dst, g.b, nullptr, n0, g.n, g.ldb, k0, kend, lw, lane);
  };
  if (nk > 0) load_kb(r[0], 0);
  if (nk > 1) load_kb(r[1], 1);
  const uint32_t idesc = umma_idesc();
  for (int kb = 0; kb < nk; ++kb) {
    const int s = kb % kStages;
    float* st = smem + (size_t)s * kStageFloats;
    if (kb >= kStages) mbar_wait(&s_bar[s], ((kb / kStages) - 1) & 1);  // MMAs of kb-kStages done
    if (kb & 1) {
      if (!is_b) store_tile<TA>(r[1], st, st + kTileFloats, split, lw, lane);
      else store_tile<TB>(r[1], st + 2 * kTileFloats, st + 3 * kTileFloats, split, lw, lane);
      if (kb + 2 < nk) load_kb(r[1], kb + 2);
    } else {
      if (!is_b) store_tile<TA>(r[0], st, st + kTileFloats, split, lw, lane);
      else store_tile<TB>(r[0], st + 2 * kTileFloats, st + 3 * kTileFloats, split, lw, lane);
      if (kb + 2 < nk) load_kb(r[0], kb + 2);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + kTileFloats * 4;
      const uint32_t b_hi = a_hi + 2 * kTileFloats * 4, b_lo = a_hi + 3 * kTileFloats * 4;
#pragma unroll
      for (int kk = 0; kk < TBK / 8; ++kk) {  // one MMA consumes K = 8 tf32 = 32 bytes
        const uint64_t dah = umma_desc_sw128(a_hi + kk * 32);
        const uint64_t dbh = umma_desc_sw128(b_hi + kk * 32);
        umma_tf32(tmem, dah, dbh, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
        if (split) {
          const uint64_t dal = umma_desc_sw128(a_lo + kk * 32);
          const uint64_t dbl = umma_desc_sw128(b_lo + kk * 32);
          umma_tf32(tmem, dah, dbl, idesc, 1u);
          umma_tf32(tmem, dal, dbh, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&s_bar[s]))
                   : "memory");
    }
  }
  if (nk > 0) {
    const int last = nk - 1;
    mbar_wait(&s_bar[last % kStages], (last / kStages) & 1);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w reads TMEM lanes [32*lw, 32*lw+32) (rows) x columns [64*(w>>2), +64),
  // transposes 32x32 blocks through shared memory (the operand stages are free now) and
  // writes whole 128-byte row segments
  float* stg = smem + warp * (32 * 33);
  const int col_half = (warp >> 2) * 64;
#pragma unroll 1
  for (int cc = 0; cc < 64; cc += 32) {
    const int c0 = col_half + cc;
    if (n0 + c0 >= g.n) break;  // warp-uniform
    uint32_t v[32];
    if (nk > 0) {
      const uint32_t taddr = tmem + ((uint32_t)(lw * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
          "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
            "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
            "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
            "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0u;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);  // thread = row
    __syncwarp();
    const int n = n0 + c0 + lane;  // now lane = column
    const bool n_ok = n < g.n;
    const float bv = (n_ok && g.bias && blockIdx.z == 0) ? g.bias[n] : 0.f;
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
      const int row = m0 + lw * 32 + rr;
      if (row < M && n_ok) {
        const float x = stg[rr * 33 + lane] + bv;
        float* dst = g.c + (long long)row * g.ldc + n;
        if (g.split_k > 1) atomicAdd(dst, x);
        else if (g.accumulate) *dst += x;
        else *dst = x;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TBN));
  }
}

bool gemm_tma_ok(const float* a, const float* b, int lda, int ldb);  // gemm_tma.cu
}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_tc_gemm(const float* a, const int64_t* a_rows, const float* b, const float* bias,
                    float* c, int32_t m, const int32_t* m_dev, int32_t n, int32_t k,
                    const int32_t* k_dev, int32_t lda, int32_t ldb, int32_t ldc, int32_t trans_a,
                    int32_t trans_b, int32_t accumulate, int32_t split_k, int32_t precision,
                    void* stream) {
  TGN_REQUIRE(m >= 0 && n >= 0 && k >= 0 && split_k >= 1, "tc_gemm: bad sizes");
  TGN_REQUIRE(precision == 1 || precision == 3, "tc_gemm: precision must be 1 (tf32) or 3 (3xtf32)");
  if (m == 0 || n == 0) return TGN_OK;
  TGN_REQUIRE(a && b && c, "tc_gemm: NULL pointer");
  TGN_REQUIRE(!(trans_a && a_rows), "tc_gemm: row gather needs a row-major A");
  if (!a_rows && gemm_tma_ok(a, b, lda, ldb)) {
    // aligned, ungathered operands: the TMA-fed kernel (gemm_tma.cu)
    tgn_gemm_desc d;
    d.a = a; d.b = b; d.bias = bias; d.c = c; d.m_dev = m_dev; d.k_dev = k_dev;
    d.m = m; d.n = n; d.k = k; d.lda = lda; d.ldb = ldb; d.ldc = ldc;
    d.trans_a = trans_a; d.trans_b = trans_b;
    d.mode = split_k > 1 ? 2 : (accumulate ? 1 : 0);
    d.split_k = split_k;
    return tgn_gemm_batch(&d, 1, precision, stream);
  }
  const size_t smem = (size_t)kStages * kStageFloats * sizeof(float);
  static unsigned long long m0 = 0, m1 = 0, m2 = 0, m3 = 0;
  TGN_CUDA(smem_optin(tc_gemm_kernel<false, false>, (int)smem, m0));
  TGN_CUDA(smem_optin(tc_gemm_kernel<false, true>, (int)smem, m1));
  TGN_CUDA(smem_optin(tc_gemm_kernel<true, false>, (int)smem, m2));
  TGN_CUDA(smem_optin(tc_gemm_kernel<true, true>, (int)smem, m3));
  TcArgs g;
  g.a = a; g.a_rows = a_rows; g.b = b; g.bias = bias; g.c = c;
  g.m = DevCount{m_dev, m}; g.k = DevCount{k_dev, k};
  g.n = n; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.accumulate = accumulate; g.split_k = split_k; g.prec = precision;
  dim3 grid(ceil_div(n, TBN), ceil_div(m, TBM), split_k);
  cudaStream_t s = (cudaStream_t)stream;
  // op(A)(m,k): !trans_a -> a[m*lda+k] (K-contiguous), trans_a -> a[k*lda+m]
  // op(B)(k,n): !trans_b -> b[n*ldb+k] (K-contiguous), trans_b -> b[k*ldb+n]
  if (!trans_a && !trans_b) launch_k(tc_gemm_kernel<false, false>, dim3(grid), dim3(256), smem, s, g);
  else if (!trans_a && trans_b) launch_k(tc_gemm_kernel<false, true>, dim3(grid), dim3(256), smem, s, g);
  else if (trans_a && !trans_b) launch_k(tc_gemm_kernel<true, false>, dim3(grid), dim3(256), smem, s, g);
  else launch_k(tc_gemm_kernel<true, true>, dim3(grid), dim3(256), smem, s, g);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
