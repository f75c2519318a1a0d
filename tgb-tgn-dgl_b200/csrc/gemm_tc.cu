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
// Structure (one CTA = 128 threads = one 128 x 128 output tile, cta_group::1):
//   * operands are staged in shared memory in the UMMA canonical K-major, no-swizzle
//     layout (8-row x 16-byte core matrices; 16-byte unit index = kchunk*rows + row), built
//     by the CTA's threads straight from global memory -- so a row-gathered A (memory[n_id])
//     and either storage order of A/B (the gradient GEMMs are TN / NN) cost nothing extra;
//     the hi/lo split happens in the same pass
//   * 2-stage ring of operand tiles, global loads for k-block i+1 are in flight in registers
//     while the tensor core works on k-block i; stage reuse is gated by tcgen05.commit ->
//     mbarrier
//   * the accumulator lives in TMEM (128 lanes x 128 fp32 columns), read back with
//     tcgen05.ld 32x32b and stored (or atomically added for split-K) with the bias
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

constexpr int TBM = 128, TBN = 128, TBK = 32;  // tile; TBK floats = 8 x 16-byte chunks
constexpr int kChunks = TBK / 4;
constexpr int kTileFloats = TBM * TBK;         // per operand per precision part
constexpr int kStageFloats = 4 * kTileFloats;  // A_hi, A_lo, B_hi, B_lo
constexpr int kStages = 2;

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

// UMMA shared-memory descriptor: K-major, SWIZZLE_NONE.  Fields in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = TBN
__device__ __forceinline__ uint32_t umma_idesc() {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 2u << 7;                    // a_format = TF32
  d |= 2u << 10;                   // b_format = TF32
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
  int n, lda, ldb, ldc, trans_a, trans_b, accumulate, split_k, prec;
};

// Loads one [rows x TBK] operand tile (k-block starting at k0) into registers: thread `t`
// owns row t (+128 for a second pass is not needed: rows == 128 == blockDim).
struct RowRegs {
  float4 v[kChunks];
};

template <bool TRANS>
__device__ __forceinline__ void load_row(RowRegs& r, const float* __restrict__ base, long long row_off,
                                         bool row_ok, int ld, int k0, int kend, int row_in_tile) {
  // !TRANS: element (row, k) at base[row_off + k]        (row_off = global_row * ld)
  //  TRANS: element (row, k) at base[k * ld + row_off]   (row_off = global_row)
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = k0 + 4 * c;
    if (row_ok) {
      if (!TRANS) {
        const float* p = base + row_off + k;
        if (k + 3 < kend && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
          v = *reinterpret_cast<const float4*>(p);
        } else {
          if (k < kend) v.x = p[0];
          if (k + 1 < kend) v.y = p[1];
          if (k + 2 < kend) v.z = p[2];
          if (k + 3 < kend) v.w = p[3];
        }
      } else {
        if (k < kend) v.x = base[(long long)k * ld + row_off];
        if (k + 1 < kend) v.y = base[(long long)(k + 1) * ld + row_off];
        if (k + 2 < kend) v.z = base[(long long)(k + 2) * ld + row_off];
        if (k + 3 < kend) v.w = base[(long long)(k + 3) * ld + row_off];
      }
    }
    r.v[c] = v;
  }
}

__device__ __forceinline__ void store_row(const RowRegs& r, float* hi, float* lo, int row, bool split) {
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const float4 v = r.v[c];
    float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    reinterpret_cast<float4*>(hi)[c * TBM + row] = h;
    if (split) {
      float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      reinterpret_cast<float4*>(lo)[c * TBM + row] = l;
    }
  }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(128, 1) tc_gemm_kernel(TcArgs g) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t s_bar[kStages];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = g.m.get(), K = g.k.get();
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
  if (m0 >= M) return;
  const int kchunk = ((K + g.split_k - 1) / g.split_k + TBK - 1) / TBK * TBK;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);
  const int nk = kend > kbeg ? (kend - kbeg + TBK - 1) / TBK : 0;
  const bool split = g.prec == 3;

  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
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

  // row ownership: thread t stages row t of the A tile and row t of the B tile
  const int am = m0 + tid, bn = n0 + tid;
  const bool a_ok = am < M, b_ok = bn < g.n;
  long long a_off, b_off;
  if (!TA) a_off = (g.a_rows ? (a_ok ? g.a_rows[am] : 0) : (long long)am) * g.lda;
  else a_off = am;
  if (!TB) b_off = (long long)bn * g.ldb;  // B stored [N,K]
  else b_off = bn;                         // B stored [K,N]

  RowRegs ra, rb;
  if (nk > 0) {
    load_row<TA>(ra, g.a, a_off, a_ok, g.lda, kbeg, kend, tid);
    load_row<TB>(rb, g.b, b_off, b_ok, g.ldb, kbeg, kend, tid);
  }
  const uint32_t idesc = umma_idesc();
  for (int kb = 0; kb < nk; ++kb) {
    const int s = kb & 1;
    float* st = smem + (size_t)s * kStageFloats;
    if (kb >= kStages) mbar_wait(&s_bar[s], ((kb >> 1) - 1) & 1);  // MMAs of k-block kb-2 done
    store_row(ra, st, st + kTileFloats, tid, split);
    store_row(rb, st + 2 * kTileFloats, st + 3 * kTileFloats, tid, split);
    if (kb + 1 < nk) {  // next k-block's global loads fly while the tensor core works
      load_row<TA>(ra, g.a, a_off, a_ok, g.lda, kbeg + (kb + 1) * TBK, kend, tid);
      load_row<TB>(rb, g.b, b_off, b_ok, g.ldb, kbeg + (kb + 1) * TBK, kend, tid);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + kTileFloats * 4;
      const uint32_t b_hi = a_hi + 2 * kTileFloats * 4, b_lo = a_hi + 3 * kTileFloats * 4;
      constexpr uint32_t LBO = TBM * 16, SBO = 128, KSTEP = 2 * TBM * 16;  // 8 floats = 2 chunks
#pragma unroll
      for (int kk = 0; kk < TBK / 8; ++kk) {
        const uint64_t dah = umma_desc(a_hi + kk * KSTEP, LBO, SBO);
        const uint64_t dbh = umma_desc(b_hi + kk * KSTEP, LBO, SBO);
        umma_tf32(tmem, dah, dbh, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
        if (split) {
          const uint64_t dal = umma_desc(a_lo + kk * KSTEP, LBO, SBO);
          const uint64_t dbl = umma_desc(b_lo + kk * KSTEP, LBO, SBO);
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
    mbar_wait(&s_bar[last & 1], (last >> 1) & 1);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w owns TMEM lanes [32w, 32w+32) = output rows m0 + 32w + lane
  const int row = m0 + warp * 32 + lane;
  float* crow = g.c + (long long)row * g.ldc;
#pragma unroll 1
  for (int c0 = 0; c0 < TBN; c0 += 16) {
    if (n0 + c0 >= g.n) break;  // warp-uniform
    uint32_t v[16];
    if (nk > 0) {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
            "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0u;
    }
    if (row < M) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = n0 + c0 + j;
        if (n < g.n) {
          float x = __uint_as_float(v[j]);
          if (g.bias && blockIdx.z == 0) x += g.bias[n];
          if (g.split_k > 1) atomicAdd(crow + n, x);
          else if (g.accumulate) crow[n] += x;
          else crow[n] = x;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TBN));
  }
}

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
  const size_t smem = (size_t)kStages * kStageFloats * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    TGN_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TGN_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TGN_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TGN_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  TcArgs g;
  g.a = a; g.a_rows = a_rows; g.b = b; g.bias = bias; g.c = c;
  g.m = DevCount{m_dev, m}; g.k = DevCount{k_dev, k};
  g.n = n; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.trans_a = trans_a; g.trans_b = trans_b;
  g.accumulate = accumulate; g.split_k = split_k; g.prec = precision;
  dim3 grid(ceil_div(n, TBN), ceil_div(m, TBM), split_k);
  cudaStream_t s = (cudaStream_t)stream;
  if (!trans_a && !trans_b) tc_gemm_kernel<false, false><<<grid, 128, smem, s>>>(g);
  else if (!trans_a && trans_b) tc_gemm_kernel<false, true><<<grid, 128, smem, s>>>(g);
  else if (trans_a && !trans_b) tc_gemm_kernel<true, false><<<grid, 128, smem, s>>>(g);
  else tc_gemm_kernel<true, true><<<grid, 128, smem, s>>>(g);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
