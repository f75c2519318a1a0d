// TMA-fed tcgen05 GEMM (the dense engine of the training step):
//     C[M,N] (=|+=|atomic+=) op(A)[M,K] * op(B)[K,N] (+ bias), fp32 in, fp32 accumulate in TMEM
// for the GRU gate GEMMs (torch.nn.GRUCell at reference modules/memory_module.py:72,172),
// the TransformerConv projections (modules/emb_module.py:21-23,29), the decoder
// (modules/decoder.py:24-27) and all of their gradients.  Up to kMaxProb independent
// problems share one launch (one CTA per 128x128 output tile of any of them).
//
// Warp roles (320 threads, one CTA per SM):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes straight into SWIZZLE_128B
//               shared-memory tiles, completion on a per-stage mbarrier (tx bytes).  Rows /
//               columns outside the tensor are zero-filled by the TMA unit, so no operand
//               needs padding.  K-major operands (row-major [MN,K]) take one [128 x 32]
//               box; MN-major operands (row-major [K,MN], i.e. the "transposed" operands of
//               the gradient GEMMs) take four [32 k x 32 mn] boxes and are consumed by the
//               tensor core through MN-major descriptors -- nothing is transposed in software
//   warp 1      MMA issuer: one lane issues tcgen05.mma.kind::tf32 (M=128, N=128, K=8) from
//               shared-memory descriptors into a 128-column TMEM accumulator, and releases a
//               stage with tcgen05.commit
//   warps 2-9   precision 3 ("3xTF32"): split every landed tile in place into hi = tf32(x)
//               and lo = x - hi (fp32-exact), so that D += a_hi b_hi + a_hi b_lo + a_lo b_hi
//               holds fp32-level accuracy (~1e-6); then the epilogue: tcgen05.ld of the
//               accumulator (warp w owns TMEM lanes 32*(w%4).. and one column half), bias, 128-bit stores /
//               vector reductions (split-K)
//   precision 1: operands go to the tensor core as they land (tf32 mantissa), no split pass.
#include <cuda.h>
#include <stdlib.h>

#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

constexpr int GM = 128, GN = 128, GK = 32;    // GK floats = 128 bytes = one swizzle span
constexpr int kGTileBytes = GM * GK * 4;      // 16 KB per operand tile
// stage layout: A_hi | B_hi | A_lo | B_lo.  3xTF32 uses all four tiles (64 KB, 3 stages);
// plain tf32 only the first two (32 KB), which doubles the ring depth in the same memory
constexpr int kGStageBytes3 = 4 * kGTileBytes, kGStages3 = 3;
constexpr int kGStageBytes1 = 2 * kGTileBytes, kGStages1 = 6;
constexpr int kGMaxStages = 6;
constexpr int kGThreads = 320;   // TMA warp, MMA warp, 8 split/epilogue warps
constexpr int kGConv = 256;      // threads of the split/epilogue warps
constexpr int kMaxProb = 4;

struct GemmProb {
  float* c;
  const float* bias;
  const int32_t* m_dev;  // nullable: live row count of C / op(A)
  const int32_t* k_dev;  // nullable: live reduction length
  int m, n, k;           // host sizes (= tensor-map extents)
  int ldc;
  int a_mn, b_mn;  // operand is MN-major (stored [K, MN])
  int split_k;
  int mode;        // 0 store, 1 accumulate, 2 atomic add
  int tiles_m, tiles_n;
  int tile_begin;  // first linear CTA index of this problem
};

struct GemmParams {
  GemmProb p[kMaxProb];
  int nprob;
  int prec;
  int stages;  // depth of the shared-memory ring for this launch (<= kGMaxStages)
};

struct alignas(64) GemmMaps {
  CUtensorMap a[kMaxProb];
  CUtensorMap b[kMaxProb];
};

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void g_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void g_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t a = g_smem_u32(bar);
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
__device__ __forceinline__ void g_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(g_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void g_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void g_tma_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                         int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(dst),
      "l"(map), "r"(g_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// shared-memory matrix descriptors.
//   K-major : SWIZZLE_128B (16-byte chunks); 8-row x 128-byte atoms 1024 bytes apart (SBO)
//   MN-major: tf32 operands only exist in the SWIZZLE_128B_BASE32B layout (32-byte chunks
//             XOR k-row % 4): atoms are [4 k-rows x 32 mn]; next 4 k-rows at SBO = 512,
//             next 32 mn at LBO = 4096 (one TMA box of 32 k-rows)
__device__ __forceinline__ uint64_t g_desc(uint32_t saddr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(mn_major ? (4096 >> 4) : 1) << 16;
  d |= (uint64_t)(mn_major ? (512 >> 4) : (1024 >> 4)) << 32;
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)(mn_major ? 1 : 2) << 61; // SWIZZLE_128B_BASE32B : SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t g_idesc(bool a_mn, bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;   // D = f32
  d |= 2u << 7;   // A = tf32
  d |= 2u << 10;  // B = tf32
  d |= (a_mn ? 1u : 0u) << 15;
  d |= (b_mn ? 1u : 0u) << 16;
  d |= (uint32_t)(GN >> 3) << 17;
  d |= (uint32_t)(GM >> 4) << 24;
  return d;
}
__device__ __forceinline__ void g_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                      uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void g_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   g_smem_u32(bar))
               : "memory");
}

// x -> (tf32(x) rounded to nearest, x - tf32(x)); element-wise, so the swizzle is irrelevant
__device__ __forceinline__ void g_split4(float4& v, float4& lo) {
  float* x = reinterpret_cast<float*>(&v);
  float* l = reinterpret_cast<float*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float h = __uint_as_float(__float_as_uint(x[i]) & 0xFFFFE000u);
    l[i] = x[i] - h;
    x[i] = h;
  }
}

__device__ __forceinline__ float4 g_lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ void g_sts4(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// zero the elements of float4 number `idx` of a [128 x 32] SWIZZLE_128B tile whose k index
// (within the k-block) is >= klive.  K-major: 8 float4 per row, k-chunk = position ^ (row & 7);
// MN-major: four [32 k x 32 mn] boxes of 256 float4, the row inside a box is the k index.
__device__ __forceinline__ void g_mask4(float4& v, int idx, bool mn_major, int klive) {
  if (mn_major) {
    if (((idx & 255) >> 3) >= klive) v = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    const int r = idx >> 3, kc = ((idx & 7) ^ (r & 7)) << 2;
    if (kc + 0 >= klive) v.x = 0.f;
    if (kc + 1 >= klive) v.y = 0.f;
    if (kc + 2 >= klive) v.z = 0.f;
    if (kc + 3 >= klive) v.w = 0.f;
  }
}

__global__ void __launch_bounds__(kGThreads, 1)
    tgemm_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ GemmParams prm) {
  pdl_wait();
  pdl_launch();
  extern __shared__ __align__(1024) uint8_t g_smem[];
  __shared__ __align__(8) uint64_t s_full[kGMaxStages], s_conv[kGMaxStages], s_empty[kGMaxStages], s_acc;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- which problem / tile / K-split is this CTA
  int pi = 0;
#pragma unroll
  for (int i = 1; i < kMaxProb; ++i)
    if (i < prm.nprob && (int)blockIdx.x >= prm.p[i].tile_begin) pi = i;
  const GemmProb& P = prm.p[pi];
  int local = blockIdx.x - P.tile_begin;
  const int tn = local % P.tiles_n;
  local /= P.tiles_n;
  const int tm = local % P.tiles_m;
  const int split = local / P.tiles_m;
  const int m0 = tm * GM, n0 = tn * GN;
  const bool split3 = prm.prec == 3;
  const int kGStages = prm.stages;
  const int kGStageBytes = split3 ? kGStageBytes3 : kGStageBytes1;

  // ---- prologue: nothing here reads what the preceding kernel wrote, so under programmatic
  // dependent launch it overlaps that kernel's tail
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(g_smem) + 1023) &
                                             ~(uintptr_t)1023);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kGMaxStages; ++i) {
      g_mbar_init(&s_full[i], 1);
      g_mbar_init(&s_conv[i], kGConv / 32);   // one arrival per split warp
      g_mbar_init(&s_empty[i], 1);
    }
    g_mbar_init(&s_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[pi]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[pi]) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     g_smem_u32(&s_tmem)),
                 "r"(GN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  pdl_wait();     // operands / device counts of the preceding kernels are visible from here on
  pdl_launch();
  const int M = P.m_dev ? min(*P.m_dev, P.m) : P.m;
  const int K = P.k_dev ? min(*P.k_dev, P.k) : P.k;
  const int kchunk = ((K + P.split_k - 1) / P.split_k + GK - 1) / GK * GK;
  const int kbeg = split * kchunk;
  const int kend = min(K, kbeg + kchunk);
  int nk = kend > kbeg ? (kend - kbeg + GK - 1) / GK : 0;
  // dead tile (past the live rows, or an empty K split with nothing to add): no loads, no
  // stores, straight to the teardown
  const bool live = m0 < M && !(nk == 0 && P.mode != 0);
  if (!live) nk = 0;
  // a live reduction length that ends inside the last 32-wide k-block (and inside the tensor,
  // where the TMA unit does not zero-fill) is cut off by zeroing the tail of that block
  const bool tail_mask = nk > 0 && kend < P.k && ((kend - kbeg) & (GK - 1)) != 0;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const CUtensorMap* ma = &maps.a[pi];
      const CUtensorMap* mb = &maps.b[pi];
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kGStages;
        if (kb >= kGStages) g_mbar_wait(&s_empty[s], ((kb / kGStages) - 1) & 1);
        const uint32_t st = g_smem_u32(smem + (size_t)s * kGStageBytes);
        const int k0 = kbeg + kb * GK;
        g_mbar_expect_tx(&s_full[s], 2 * kGTileBytes);
        if (!P.a_mn) {
          g_tma_2d(st, ma, &s_full[s], k0, m0);
        } else {
#pragma unroll
          for (int b = 0; b < 4; ++b) g_tma_2d(st + b * 4096, ma, &s_full[s], m0 + 32 * b, k0);
        }
        if (!P.b_mn) {
          g_tma_2d(st + kGTileBytes, mb, &s_full[s], k0, n0);
        } else {
#pragma unroll
          for (int b = 0; b < 4; ++b)
            g_tma_2d(st + kGTileBytes + b * 4096, mb, &s_full[s], n0 + 32 * b, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = g_idesc(P.a_mn != 0, P.b_mn != 0);
    const uint32_t a_step = P.a_mn ? 1024u : 32u, b_step = P.b_mn ? 1024u : 32u;
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % kGStages;
      // with a split or mask pass in this CTA the issuer follows warps 2-5 through every
      // phase (a barrier may only be waited on phase by phase)
      g_mbar_wait((split3 || tail_mask) ? &s_conv[s] : &s_full[s], (kb / kGStages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = g_smem_u32(smem + (size_t)s * kGStageBytes);
        const uint32_t b_hi = a_hi + kGTileBytes, a_lo = a_hi + 2 * kGTileBytes,
                       b_lo = a_hi + 3 * kGTileBytes;
#pragma unroll
        for (int kk = 0; kk < GK / 8; ++kk) {  // one MMA consumes K = 8 tf32
          const uint64_t dah = g_desc(a_hi + kk * a_step, P.a_mn != 0);
          const uint64_t dbh = g_desc(b_hi + kk * b_step, P.b_mn != 0);
          g_mma(tmem, dah, dbh, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          if (split3) {
            const uint64_t dal = g_desc(a_lo + kk * a_step, P.a_mn != 0);
            const uint64_t dbl = g_desc(b_lo + kk * b_step, P.b_mn != 0);
            g_mma(tmem, dah, dbl, idesc, 1u);
            g_mma(tmem, dal, dbh, idesc, 1u);
          }
        }
        g_commit(&s_empty[s]);
        if (kb == nk - 1) g_commit(&s_acc);
      }
      __syncwarp();
    }
  } else {
    // ===== split pass (3xTF32) + epilogue: warps 2..9 =====
    const int et = tid - 64;  // 0..255
    if (split3 || tail_mask) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kGStages;
        g_mbar_wait(&s_full[s], (kb / kGStages) & 1);
        if (!split3 && kb != nk - 1) {  // tf32 mode: only the masked last block is touched
          if (lane == 0) g_mbar_arrive(&s_conv[s]);
          continue;
        }
        const uint32_t a_hi = g_smem_u32(smem + (size_t)s * kGStageBytes);
        const int klive = (tail_mask && kb == nk - 1) ? kend - (kbeg + kb * GK) : GK;
        float4 va[4], vb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // all loads first: 8 x 128-bit shared loads in flight
          const uint32_t o = (uint32_t)(i * kGConv + et) * 16u;
          va[i] = g_lds4(a_hi + o);
          vb[i] = g_lds4(a_hi + kGTileBytes + o);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = i * kGConv + et;
          const uint32_t o = (uint32_t)idx * 16u;
          float4 la, lb;
          if (klive < GK) {
            g_mask4(va[i], idx, P.a_mn != 0, klive);
            g_mask4(vb[i], idx, P.b_mn != 0, klive);
          }
          if (split3) {
            g_split4(va[i], la);
            g_split4(vb[i], lb);
            g_sts4(a_hi + 2 * kGTileBytes + o, la);
            g_sts4(a_hi + 3 * kGTileBytes + o, lb);
          }
          if (klive < GK) {  // masked tail: the raw tile itself changes
            g_sts4(a_hi + o, va[i]);
            g_sts4(a_hi + kGTileBytes + o, vb[i]);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic -> async proxy
        __syncwarp();                                  // one arrival per warp: 8 barrier updates per
        if (lane == 0) g_mbar_arrive(&s_conv[s]);      // k-block instead of 256 serialised ones
      }
    }
    if (nk > 0) {
      g_mbar_wait(&s_acc, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int q = warp & 3;             // TMEM lane group this warp may read
    const int chalf = warp >= 6 ? 64 : 0;  // warps 2-5: columns 0..63, warps 6-9: 64..127
    const int row = m0 + q * 32 + lane;
    const bool row_ok = live && row < M;
    float* crow = P.c + (long long)row * P.ldc;
    const bool vec_ok = (P.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(P.c) & 15) == 0);
#pragma unroll 1
    for (int cc = chalf; cc < chalf + 64; cc += 32) {
      const int c0 = n0 + cc;
      if (c0 >= P.n) break;  // warp-uniform
      uint32_t v[32];
      if (nk > 0) {
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cc;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
              "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
              "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (!row_ok) continue;
      const bool full = c0 + 32 <= P.n;
      const bool add_bias = P.bias != nullptr && split == 0;
      if (full && vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                 __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          if (add_bias) {
            const float4 b4 = *reinterpret_cast<const float4*>(P.bias + c0 + j);
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
          }
          float4* dst = reinterpret_cast<float4*>(crow + c0 + j);
          if (P.mode == 2) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x),
                         "f"(o.y), "f"(o.z), "f"(o.w)
                         : "memory");
          } else if (P.mode == 1) {
            const float4 old = *dst;
            *dst = make_float4(old.x + o.x, old.y + o.y, old.z + o.z, old.w + o.w);
          } else {
            *dst = o;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = c0 + j;
          if (n < P.n) {
            float x = __uint_as_float(v[j]);
            if (add_bias) x += P.bias[n];
            if (P.mode == 2) atomicAdd(crow + n, x);
            else if (P.mode == 1) crow[n] += x;
            else crow[n] = x;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(GN));
  }
}


// ---------------------------------------------------------------------------
// Persistent variant for launches of several waves of tiles (evaluation, flush, module path, large batches):
// one CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the shared-memory ring runs on across
// tile boundaries and the accumulator is double-buffered in TMEM (2 x 128 columns), so the epilogue of tile i
// (four dedicated warps) overlaps the loads, the split pass and the MMAs of tile i + 1.  tgemm_kernel pays its
// prologue, the pipeline fill and the epilogue once PER TILE (ncu, round 1: 17 us per 15-k-block tile, the
// tensor pipe idle most of it).
//   warp 0 TMA producer | warp 1 MMA issuer | warps 2-9 split / mask pass | warps 10-13 epilogue
// ---------------------------------------------------------------------------
constexpr int kPThreads = 448;
constexpr int kPEpiWarp0 = 10;

struct TileInfo {
  int pi, m0, n0, split, M, kbeg, kend, nk;
  bool live, tail_mask;
};

__device__ __forceinline__ TileInfo g_tile_info(const GemmParams& prm, int t) {
  TileInfo ti;
  int pi = 0;
#pragma unroll
  for (int i = 1; i < kMaxProb; ++i)
    if (i < prm.nprob && t >= prm.p[i].tile_begin) pi = i;
  const GemmProb& P = prm.p[pi];
  int local = t - P.tile_begin;
  const int tn = local % P.tiles_n;
  local /= P.tiles_n;
  const int tm = local % P.tiles_m;
  ti.pi = pi;
  ti.split = local / P.tiles_m;
  ti.m0 = tm * GM;
  ti.n0 = tn * GN;
  ti.M = P.m_dev ? min(*P.m_dev, P.m) : P.m;
  const int K = P.k_dev ? min(*P.k_dev, P.k) : P.k;
  const int kchunk = ((K + P.split_k - 1) / P.split_k + GK - 1) / GK * GK;
  ti.kbeg = ti.split * kchunk;
  ti.kend = min(K, ti.kbeg + kchunk);
  int nk = ti.kend > ti.kbeg ? (ti.kend - ti.kbeg + GK - 1) / GK : 0;
  ti.live = ti.m0 < ti.M && !(nk == 0 && P.mode != 0);
  if (!ti.live) nk = 0;
  ti.nk = nk;
  ti.tail_mask = nk > 0 && ti.kend < P.k && ((ti.kend - ti.kbeg) & (GK - 1)) != 0;
  return ti;
}

__global__ void __launch_bounds__(kPThreads, 1)
    tgemm_persist_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ GemmParams prm,
                         int total_tiles) {
  pdl_wait();
  pdl_launch();
  extern __shared__ __align__(1024) uint8_t g_smem[];
  __shared__ __align__(8) uint64_t s_full[kGMaxStages], s_conv[kGMaxStages], s_empty[kGMaxStages];
  __shared__ __align__(8) uint64_t s_accf[2], s_acce[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool split3 = prm.prec == 3;
  const int S = prm.stages;
  const int kStageBytes = split3 ? kGStageBytes3 : kGStageBytes1;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(g_smem) + 1023) &
                                             ~(uintptr_t)1023);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kGMaxStages; ++i) {
      g_mbar_init(&s_full[i], 1);
      g_mbar_init(&s_conv[i], kGConv / 32);
      g_mbar_init(&s_empty[i], 1);
    }
    g_mbar_init(&s_accf[0], 1);
    g_mbar_init(&s_accf[1], 1);
    g_mbar_init(&s_acce[0], 4);      // one arrival per epilogue warp
    g_mbar_init(&s_acce[1], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < prm.nprob; ++i) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[i]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[i]) : "memory");
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     g_smem_u32(&s_tmem)),
                 "r"(2 * GN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileInfo ti = g_tile_info(prm, t);
        const GemmProb& P = prm.p[ti.pi];
        const CUtensorMap* ma = &maps.a[ti.pi];
        const CUtensorMap* mb = &maps.b[ti.pi];
        for (int kb = 0; kb < ti.nk; ++kb, ++it) {
          const int s = it % S;
          if (it >= S) g_mbar_wait(&s_empty[s], ((it / S) - 1) & 1);
          const uint32_t st = g_smem_u32(smem + (size_t)s * kStageBytes);
          const int k0 = ti.kbeg + kb * GK;
          g_mbar_expect_tx(&s_full[s], 2 * kGTileBytes);
          if (!P.a_mn) {
            g_tma_2d(st, ma, &s_full[s], k0, ti.m0);
          } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) g_tma_2d(st + b * 4096, ma, &s_full[s], ti.m0 + 32 * b, k0);
          }
          if (!P.b_mn) {
            g_tma_2d(st + kGTileBytes, mb, &s_full[s], k0, ti.n0);
          } else {
#pragma unroll
            for (int b = 0; b < 4; ++b)
              g_tma_2d(st + kGTileBytes + b * 4096, mb, &s_full[s], ti.n0 + 32 * b, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int it = 0, acc_it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileInfo ti = g_tile_info(prm, t);
      if (ti.nk == 0) continue;
      const GemmProb& P = prm.p[ti.pi];
      const int buf = acc_it & 1;
      if (acc_it >= 2) g_mbar_wait(&s_acce[buf], ((acc_it >> 1) - 1) & 1);   // the epilogue has drained this buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem + (uint32_t)(buf * GN);
      const uint32_t idesc = g_idesc(P.a_mn != 0, P.b_mn != 0);
      const uint32_t a_step = P.a_mn ? 1024u : 32u, b_step = P.b_mn ? 1024u : 32u;
      for (int kb = 0; kb < ti.nk; ++kb, ++it) {
        const int s = it % S;
        g_mbar_wait(&s_conv[s], (it / S) & 1);     // (the split warps arrive for every k-block in this kernel)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t a_hi = g_smem_u32(smem + (size_t)s * kStageBytes);
          const uint32_t b_hi = a_hi + kGTileBytes, a_lo = a_hi + 2 * kGTileBytes,
                         b_lo = a_hi + 3 * kGTileBytes;
#pragma unroll
          for (int kk = 0; kk < GK / 8; ++kk) {
            const uint64_t dah = g_desc(a_hi + kk * a_step, P.a_mn != 0);
            const uint64_t dbh = g_desc(b_hi + kk * b_step, P.b_mn != 0);
            g_mma(acc, dah, dbh, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            if (split3) {
              const uint64_t dal = g_desc(a_lo + kk * a_step, P.a_mn != 0);
              const uint64_t dbl = g_desc(b_lo + kk * b_step, P.b_mn != 0);
              g_mma(acc, dah, dbl, idesc, 1u);
              g_mma(acc, dal, dbh, idesc, 1u);
            }
          }
          g_commit(&s_empty[s]);
          if (kb == ti.nk - 1) g_commit(&s_accf[buf]);
        }
        __syncwarp();
      }
      ++acc_it;
    }
  } else if (warp < kPEpiWarp0) {
    // ===== split / mask pass: warps 2..9 =====
    const int et = tid - 64;  // 0..255
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileInfo ti = g_tile_info(prm, t);
      const GemmProb& P = prm.p[ti.pi];
      for (int kb = 0; kb < ti.nk; ++kb, ++it) {
        const int s = it % S;
        g_mbar_wait(&s_full[s], (it / S) & 1);
        const bool masked = ti.tail_mask && kb == ti.nk - 1;
        if (split3 || masked) {
          const uint32_t a_hi = g_smem_u32(smem + (size_t)s * kStageBytes);
          const int klive = masked ? ti.kend - (ti.kbeg + kb * GK) : GK;
          float4 va[4], vb[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t o = (uint32_t)(i * kGConv + et) * 16u;
            va[i] = g_lds4(a_hi + o);
            vb[i] = g_lds4(a_hi + kGTileBytes + o);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int idx = i * kGConv + et;
            const uint32_t o = (uint32_t)idx * 16u;
            float4 la, lb;
            if (klive < GK) {
              g_mask4(va[i], idx, P.a_mn != 0, klive);
              g_mask4(vb[i], idx, P.b_mn != 0, klive);
            }
            if (split3) {
              g_split4(va[i], la);
              g_split4(vb[i], lb);
              g_sts4(a_hi + 2 * kGTileBytes + o, la);
              g_sts4(a_hi + 3 * kGTileBytes + o, lb);
            }
            if (klive < GK) {
              g_sts4(a_hi + o, va[i]);
              g_sts4(a_hi + kGTileBytes + o, vb[i]);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) g_mbar_arrive(&s_conv[s]);
      }
    }
  } else {
    // ===== epilogue: warps 10..13, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    int acc_it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileInfo ti = g_tile_info(prm, t);
      if (!ti.live) continue;
      const GemmProb& P = prm.p[ti.pi];
      const int buf = acc_it & 1;
      if (ti.nk > 0) {
        g_mbar_wait(&s_accf[buf], (acc_it >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const int row = ti.m0 + q * 32 + lane;
      const bool row_ok = row < ti.M;
      float* crow = P.c + (long long)row * P.ldc;
      const bool vec_ok = (P.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(P.c) & 15) == 0);
      const bool add_bias = P.bias != nullptr && ti.split == 0;
#pragma unroll 1
      for (int cc = 0; cc < GN; cc += 32) {
        const int c0 = ti.n0 + cc;
        if (c0 >= P.n) break;  // warp-uniform
        uint32_t v[32];
        if (ti.nk > 0) {
          const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * GN + cc);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
              "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
                "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                "=r"(v[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (!row_ok) continue;
        const bool full = c0 + 32 <= P.n;
        if (full && vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (add_bias) {
              const float4 b4 = *reinterpret_cast<const float4*>(P.bias + c0 + j);
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            }
            float4* dst = reinterpret_cast<float4*>(crow + c0 + j);
            if (P.mode == 2) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x),
                           "f"(o.y), "f"(o.z), "f"(o.w)
                           : "memory");
            } else if (P.mode == 1) {
              const float4 old = *dst;
              *dst = make_float4(old.x + o.x, old.y + o.y, old.z + o.z, old.w + o.w);
            } else {
              *dst = o;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = c0 + j;
            if (n < P.n) {
              float x = __uint_as_float(v[j]);
              if (add_bias) x += P.bias[n];
              if (P.mode == 2) atomicAdd(crow + n, x);
              else if (P.mode == 1) crow[n] += x;
              else crow[n] = x;
            }
          }
        }
      }
      if (ti.nk > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) g_mbar_arrive(&s_acce[buf]);
        ++acc_it;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * GN));
  }
}

// ---------------------------------------------------------------------------
// Fused GRUCell forward (torch.nn.GRUCell as used at reference modules/memory_module.py:72,172):
//   gi = x W_ih^T + b_ih,  gh = h W_hh^T + b_hh,  r = sig(gi_r + gh_r),  z = sig(gi_z + gh_z),
//   n = tanh(gi_n + r * gh_n),  h' = n + z * (h - n)
// One CTA owns 128 rows x 40 hidden units.  Its B tile is assembled by the TMA unit from
// three 40-row boxes of the weight matrix (the r, z and n rows of those units), so the two
// accumulators in TMEM -- gi in columns [0,128), gh in [128,256) -- hold complete (r, z, n)
// triples and the gate math runs in the epilogue straight out of TMEM: gi / gh never exist in
// global memory.  The K loop first sweeps x / W_ih into the gi accumulator, then h / W_hh into
// the gh accumulator, through the same shared-memory ring, split pass and issuer as
// tgemm_kernel.  Grid = ceil(S/128) x ceil(D/40) CTAs: one wave at the step's ~5k rows, D = 100.
// ---------------------------------------------------------------------------
constexpr int kGruUnits = 40;                       // hidden units per CTA (x 3 gates = 120 of 128 B-tile rows)
constexpr int kGruBoxBytes = kGruUnits * GK * 4;    // 5120: five 1024-byte swizzle atoms

struct alignas(64) GruMaps {
  CUtensorMap x, h, wih, whh;
};
struct GruParams {
  const float* h;       // [S, D]
  const float* b_ih;    // [3D]
  const float* b_hh;    // [3D]
  float* out;           // [S, D]
  float* gates;         // [S, 4D] nullable: r, z, n, gh_n
  const int32_t* m_dev;
  int m, dx, D, prec;
  int late_pdl;         // 1: dependents are released after the last MMA has been issued (not at kernel start)
};

template <int N>
__device__ __forceinline__ void g_tmem_ld(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void g_tmem_ld<16>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void g_tmem_ld<4>(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}

__device__ __forceinline__ float g_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// gate math for NU consecutive units starting at local unit `u` (global unit j0 + u) of one row
template <int NU>
__device__ __forceinline__ void gru_epilogue_chunk(const GruParams& P, uint32_t trow, int row,
                                                   bool row_ok, int j0, int u) {
  uint32_t ir[NU], iz[NU], in_[NU], hr[NU], hz[NU], hn[NU];
  g_tmem_ld<NU>(trow + (uint32_t)u, ir);
  g_tmem_ld<NU>(trow + (uint32_t)(kGruUnits + u), iz);
  g_tmem_ld<NU>(trow + (uint32_t)(2 * kGruUnits + u), in_);
  g_tmem_ld<NU>(trow + (uint32_t)(GN + u), hr);
  g_tmem_ld<NU>(trow + (uint32_t)(GN + kGruUnits + u), hz);
  g_tmem_ld<NU>(trow + (uint32_t)(GN + 2 * kGruUnits + u), hn);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (!row_ok) return;
  const int D = P.D;
#pragma unroll
  for (int c = 0; c < NU; c += 4) {
    const int j = j0 + u + c;
    if (j >= D) break;  // D is a multiple of 4: a float4 is entirely live or entirely dead
    const float4 bir = *reinterpret_cast<const float4*>(P.b_ih + j);
    const float4 biz = *reinterpret_cast<const float4*>(P.b_ih + D + j);
    const float4 bin = *reinterpret_cast<const float4*>(P.b_ih + 2 * D + j);
    const float4 bhr = *reinterpret_cast<const float4*>(P.b_hh + j);
    const float4 bhz = *reinterpret_cast<const float4*>(P.b_hh + D + j);
    const float4 bhn = *reinterpret_cast<const float4*>(P.b_hh + 2 * D + j);
    const float4 hv = *reinterpret_cast<const float4*>(P.h + (long long)row * D + j);
    float4 o_r, o_z, o_n, o_g, o_h;
    const float* pbir = &bir.x; const float* pbiz = &biz.x; const float* pbin = &bin.x;
    const float* pbhr = &bhr.x; const float* pbhz = &bhz.x; const float* pbhn = &bhn.x;
    const float* phv = &hv.x;
    float* por = &o_r.x; float* poz = &o_z.x; float* pon = &o_n.x; float* pog = &o_g.x;
    float* poh = &o_h.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // same association as the unfused path: (acc + bias) per GEMM, then the gate sums
      const float r = g_sigmoid((__uint_as_float(ir[c + e]) + pbir[e]) + (__uint_as_float(hr[c + e]) + pbhr[e]));
      const float z = g_sigmoid((__uint_as_float(iz[c + e]) + pbiz[e]) + (__uint_as_float(hz[c + e]) + pbhz[e]));
      const float ghn = __uint_as_float(hn[c + e]) + pbhn[e];
      const float n = tanhf((__uint_as_float(in_[c + e]) + pbin[e]) + r * ghn);
      por[e] = r; poz[e] = z; pon[e] = n; pog[e] = ghn;
      poh[e] = n + z * (phv[e] - n);
    }
    *reinterpret_cast<float4*>(P.out + (long long)row * D + j) = o_h;
    if (P.gates) {
      float* g = P.gates + (long long)row * 4 * D + j;
      *reinterpret_cast<float4*>(g) = o_r;
      *reinterpret_cast<float4*>(g + D) = o_z;
      *reinterpret_cast<float4*>(g + 2 * D) = o_n;
      *reinterpret_cast<float4*>(g + 3 * D) = o_g;
    }
  }
}

__global__ void __launch_bounds__(kGThreads, 1)
    gru_fused_kernel(const __grid_constant__ GruMaps maps, const __grid_constant__ GruParams P) {
  pdl_wait();
  if (!P.late_pdl) pdl_launch();
  extern __shared__ __align__(1024) uint8_t g_smem[];
  __shared__ __align__(8) uint64_t s_full[kGStages3], s_conv[kGStages3], s_empty[kGStages3], s_acc;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ngroups = (P.D + kGruUnits - 1) / kGruUnits;
  const int ug = blockIdx.x % ngroups, tm = blockIdx.x / ngroups;
  const int m0 = tm * GM, j0 = ug * kGruUnits;
  const bool split3 = P.prec == 3;
  constexpr int kStages = kGStages3;
  constexpr int kStageBytes = kGStageBytes3;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(g_smem) + 1023) &
                                             ~(uintptr_t)1023);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      g_mbar_init(&s_full[i], 1);
      g_mbar_init(&s_conv[i], kGConv / 32);   // one arrival per split warp
      g_mbar_init(&s_empty[i], 1);
    }
    g_mbar_init(&s_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wih) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.h) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.whh) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     g_smem_u32(&s_tmem)),
                 "r"(2 * GN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the unused B-tile rows (3 x 40 = 120 of 128) are never written by the TMA unit: zero them once
  // so the split pass and the tensor core never touch uninitialised values
  for (int s = 0; s < kStages; ++s) {
    uint8_t* b_hi = smem + (size_t)s * kStageBytes + kGTileBytes + 3 * kGruBoxBytes;
    for (int i = tid; i < (kGTileBytes - 3 * kGruBoxBytes) / 16; i += kGThreads)
      reinterpret_cast<float4*>(b_hi)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  pdl_wait();
  if (!P.late_pdl) pdl_launch();
  const int M = P.m_dev ? min(*P.m_dev, P.m) : P.m;
  const int nk1 = (P.dx + GK - 1) / GK, nk2 = (P.D + GK - 1) / GK;
  const bool live = m0 < M;
  const int nk = live ? nk1 + nk2 : 0;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kStages;
        if (kb >= kStages) g_mbar_wait(&s_empty[s], ((kb / kStages) - 1) & 1);
        const uint32_t st = g_smem_u32(smem + (size_t)s * kStageBytes);
        const bool ph2 = kb >= nk1;
        const int k0 = (ph2 ? kb - nk1 : kb) * GK;
        const CUtensorMap* ma = ph2 ? &maps.h : &maps.x;
        const CUtensorMap* mb = ph2 ? &maps.whh : &maps.wih;
        g_mbar_expect_tx(&s_full[s], kGTileBytes + 3 * kGruBoxBytes);
        g_tma_2d(st, ma, &s_full[s], k0, m0);
#pragma unroll
        for (int g = 0; g < 3; ++g)
          g_tma_2d(st + kGTileBytes + g * kGruBoxBytes, mb, &s_full[s], k0, g * P.D + j0);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = g_idesc(false, false);
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % kStages;
      g_mbar_wait(split3 ? &s_conv[s] : &s_full[s], (kb / kStages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const bool ph2 = kb >= nk1;
        const uint32_t acc = tmem + (ph2 ? (uint32_t)GN : 0u);
        const bool first = ph2 ? kb == nk1 : kb == 0;
        const uint32_t a_hi = g_smem_u32(smem + (size_t)s * kStageBytes);
        const uint32_t b_hi = a_hi + kGTileBytes, a_lo = a_hi + 2 * kGTileBytes,
                       b_lo = a_hi + 3 * kGTileBytes;
#pragma unroll
        for (int kk = 0; kk < GK / 8; ++kk) {
          const uint64_t dah = g_desc(a_hi + kk * 32u, false);
          const uint64_t dbh = g_desc(b_hi + kk * 32u, false);
          g_mma(acc, dah, dbh, idesc, (!first || kk > 0) ? 1u : 0u);
          if (split3) {
            const uint64_t dal = g_desc(a_lo + kk * 32u, false);
            const uint64_t dbl = g_desc(b_lo + kk * 32u, false);
            g_mma(acc, dah, dbl, idesc, 1u);
            g_mma(acc, dal, dbh, idesc, 1u);
          }
        }
        g_commit(&s_empty[s]);
        if (kb == nk - 1) g_commit(&s_acc);
      }
      __syncwarp();
    }
    if (P.late_pdl) pdl_launch();     // (dead tiles: nk = 0, released at once)
  } else {
    const int et = tid - 64;
    if (split3) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kStages;
        g_mbar_wait(&s_full[s], (kb / kStages) & 1);
        const uint32_t a_hi = g_smem_u32(smem + (size_t)s * kStageBytes);
        float4 va[4], vb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t o = (uint32_t)(i * kGConv + et) * 16u;
          va[i] = g_lds4(a_hi + o);
          vb[i] = g_lds4(a_hi + kGTileBytes + o);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t o = (uint32_t)(i * kGConv + et) * 16u;
          float4 la, lb;
          g_split4(va[i], la);
          g_split4(vb[i], lb);
          g_sts4(a_hi + 2 * kGTileBytes + o, la);
          g_sts4(a_hi + 3 * kGTileBytes + o, lb);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) g_mbar_arrive(&s_conv[s]);
      }
    }
    if (nk > 0) {
      g_mbar_wait(&s_acc, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3;
      const int u0 = warp >= 6 ? kGruUnits / 2 : 0;   // warps 2-5: units 0..19, warps 6-9: 20..39
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
      if (j0 + u0 < P.D) {                              // warp-uniform
        gru_epilogue_chunk<16>(P, trow, row, row_ok, j0, u0);
        if (j0 + u0 + 16 < P.D) gru_epilogue_chunk<4>(P, trow, row, row_ok, j0, u0 + 16);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * GN));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// row-major fp32 matrix [rows, cols] with leading dimension ld -> 2-D tensor map whose
// inner dimension is the column index; box = box_cols x box_rows, SWIZZLE_128B
static int make_map(CUtensorMap* map, const float* base, long long rows, long long cols,
                    long long ld, int box_cols, int box_rows, bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(TGN_ECUDA, "gemm: cuTensorMapEncodeTiled is not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(TGN_ECUDA, "gemm: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TGN_OK;
}

bool gemm_tma_ok(const float* a, const float* b, int lda, int ldb) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0 &&
         (lda & 3) == 0 && (ldb & 3) == 0;
}

}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_gemm_batch(const tgn_gemm_desc* d, int32_t count, int32_t precision, void* stream) {
  TGN_REQUIRE(d && count >= 1 && count <= kMaxProb, "gemm_batch: 1..%d problems per launch", kMaxProb);
  TGN_REQUIRE(precision == 1 || precision == 3, "gemm_batch: precision must be 1 (tf32) or 3 (3xtf32)");
  GemmMaps maps;
  GemmParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.prec = precision;
  int tiles = 0, np = 0;
  for (int i = 0; i < count; ++i) {
    const tgn_gemm_desc& g = d[i];
    TGN_REQUIRE(g.m >= 0 && g.n >= 0 && g.k >= 0 && g.split_k >= 1, "gemm_batch[%d]: bad sizes", i);
    TGN_REQUIRE(g.mode >= 0 && g.mode <= 2, "gemm_batch[%d]: mode must be 0, 1 or 2", i);
    TGN_REQUIRE(g.split_k == 1 || g.mode == 2, "gemm_batch[%d]: split_k > 1 needs mode 2 (atomic)", i);
    if (g.m == 0 || g.n == 0) continue;
    TGN_REQUIRE(g.a && g.b && g.c, "gemm_batch[%d]: NULL pointer", i);
    TGN_REQUIRE(gemm_tma_ok(g.a, g.b, g.lda, g.ldb),
                "gemm_batch[%d]: operands must be 16-byte aligned with lda, ldb multiples of 4", i);
    GemmProb& P = prm.p[np];
    P.c = g.c; P.bias = g.bias; P.m_dev = g.m_dev; P.k_dev = g.k_dev;
    P.m = g.m; P.n = g.n; P.k = g.k; P.ldc = g.ldc;
    P.a_mn = g.trans_a ? 1 : 0; P.b_mn = g.trans_b ? 1 : 0;
    P.split_k = g.split_k; P.mode = g.mode;
    P.tiles_m = ceil_div(g.m, GM); P.tiles_n = ceil_div(g.n, GN);
    P.tile_begin = tiles;
    tiles += P.tiles_m * P.tiles_n * g.split_k;
    int rc;
    // op(A)(m,k): !trans_a -> a[m*lda+k]; trans_a -> a[k*lda+m]
    if (!g.trans_a) rc = make_map(&maps.a[np], g.a, g.m, g.k, g.lda, GK, GM, false);
    else rc = make_map(&maps.a[np], g.a, g.k, g.m, g.lda, 32, GK, true);
    if (rc != TGN_OK) return rc;
    // op(B)(k,n): !trans_b -> b[n*ldb+k]; trans_b -> b[k*ldb+n]
    if (!g.trans_b) rc = make_map(&maps.b[np], g.b, g.n, g.k, g.ldb, GK, GN, false);
    else rc = make_map(&maps.b[np], g.b, g.k, g.n, g.ldb, 32, GK, true);
    if (rc != TGN_OK) return rc;
    ++np;
  }
  if (np == 0) return TGN_OK;
  prm.nprob = np;
  const size_t smem_max = (size_t)kGStages3 * kGStageBytes3 + 1024;
  static unsigned long long attr_mask = 0;
  TGN_CUDA(smem_optin(tgemm_kernel, (int)smem_max, attr_mask));
  // Ring depth.  A full ring (192 KB) allows one CTA per SM.  A launch of short reductions (K <= 4
  // k-blocks) whose tiles do not fit one wave of 148 CTAs runs with a 64 KB ring instead: three CTAs per
  // SM overlap each other's load / split / MMA / epilogue phases, which beats further waves of deeply
  // pipelined CTAs (the node projection of the attention: 4 column tiles x 38 live row tiles = 152 CTAs;
  // the evaluation's edge projection: 1,418 tiles, batch 0.604 -> 0.582 ms).
  prm.stages = precision == 3 ? kGStages3 : kGStages1;
  int max_nk = 0;
  for (int i = 0; i < np; ++i) {
    const int kk = ceil_div(ceil_div(prm.p[i].k, prm.p[i].split_k), GK);
    if (kk > max_nk) max_nk = kk;
  }
  if (tiles > kNumSMs && max_nk <= 4) prm.stages = precision == 3 ? 1 : 2;
  static const int persist_min = getenv("TGN_GEMM_PERSIST") ? atoi(getenv("TGN_GEMM_PERSIST")) : 2 * kNumSMs;
  if (tiles > persist_min) {
    // several waves of tiles: one persistent CTA per SM, full-depth ring, accumulators double-buffered in TMEM
    prm.stages = precision == 3 ? kGStages3 : kGStages1;
    static unsigned long long attr_mask_p = 0;
    TGN_CUDA(smem_optin(tgemm_persist_kernel, (int)smem_max, attr_mask_p));
    launch_k(tgemm_persist_kernel, dim3(kNumSMs), dim3(kPThreads), smem_max, (cudaStream_t)stream, maps, prm, tiles);
    TGN_LAUNCH_CHECK();
    return TGN_OK;
  }
  const size_t smem = (size_t)prm.stages * (precision == 3 ? kGStageBytes3 : kGStageBytes1) + 1024;
  launch_k(tgemm_kernel, dim3(tiles), dim3(kGThreads), smem, (cudaStream_t)stream, maps, prm);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_gru_fused_fwd(const float* x, int32_t ldx, int32_t dx, const float* h, int32_t dim,
                          const float* w_ih, int32_t ldw_ih, const float* w_hh, const float* b_ih,
                          const float* b_hh, int32_t num, const int32_t* num_dev, int32_t precision,
                          float* out, float* gates, void* stream) {
  TGN_REQUIRE(num >= 0 && dim >= 4 && dx >= 1, "gru_fused_fwd: bad sizes");
  TGN_REQUIRE(precision == 1 || precision == 3, "gru_fused_fwd: precision must be 1 (tf32) or 3 (3xtf32)");
  TGN_REQUIRE((dim & 3) == 0 && (ldx & 3) == 0 && (ldw_ih & 3) == 0 && ldx >= dx && ldw_ih >= dx,
              "gru_fused_fwd: dim, ldx, ldw_ih must be multiples of 4 and cover dx");
  if (num == 0) return TGN_OK;
  TGN_REQUIRE(x && h && w_ih && w_hh && b_ih && b_hh && out, "gru_fused_fwd: NULL pointer");
  const uintptr_t al = (uintptr_t)x | (uintptr_t)h | (uintptr_t)w_ih | (uintptr_t)w_hh |
                       (uintptr_t)b_ih | (uintptr_t)b_hh | (uintptr_t)out | (uintptr_t)gates;
  TGN_REQUIRE((al & 15) == 0, "gru_fused_fwd: pointers must be 16-byte aligned");
  GruMaps maps;
  int rc;
  if ((rc = make_map(&maps.x, x, num, dx, ldx, GK, GM, false)) != TGN_OK) return rc;
  if ((rc = make_map(&maps.h, h, num, dim, dim, GK, GM, false)) != TGN_OK) return rc;
  if ((rc = make_map(&maps.wih, w_ih, 3 * dim, dx, ldw_ih, GK, kGruUnits, false)) != TGN_OK) return rc;
  if ((rc = make_map(&maps.whh, w_hh, 3 * dim, dim, dim, GK, kGruUnits, false)) != TGN_OK) return rc;
  GruParams P;
  P.h = h; P.b_ih = b_ih; P.b_hh = b_hh; P.out = out; P.gates = gates; P.m_dev = num_dev;
  P.m = num; P.dx = dx; P.D = dim; P.prec = precision;
  P.late_pdl = 0;     // A/B on the review step (tools/ab_env.py): 147.9 vs 149.1 us, inside the run-to-run spread
  const size_t smem = (size_t)kGStages3 * kGStageBytes3 + 1024;
  static unsigned long long attr_mask = 0;
  TGN_CUDA(smem_optin(gru_fused_kernel, (int)smem, attr_mask));
  const int tiles = ceil_div(num, GM) * ceil_div(dim, kGruUnits);
  launch_k(gru_fused_kernel, dim3(tiles), dim3(kGThreads), smem, (cudaStream_t)stream, maps, P);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
