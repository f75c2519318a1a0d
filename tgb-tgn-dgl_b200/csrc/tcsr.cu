// Temporal neighbour sampling over a t-CSR graph (TGL sampler_core semantics;
// the reference names it at README.md:1-5, the t-CSR keys at utils.py:73 and
// the parameters at config/TGN.yml:1-9 -- its C++ source is not part of the
// reference tree, see SURVEY.md B1 for the restated algorithm).
//
// One CTA handles a tile of 256 roots:
//   phase 1  one thread per root: search of the timestamp-sorted row for the
//            candidate window [lo, hi).  With the skip index (`coarse`, every
//            16th timestamp, built once per graph by tgn_tcsr_build_index) the
//            search is: binary search over the row's slice of the index (7
//            entries = one 32-byte sector at degree 113), then ONE 64-byte
//            block of the row read with four 128-bit loads and counted --
//            ~3 dependent memory round trips instead of log2(deg)+1, and the
//            block read is the same line the most-recent entries are emitted
//            from.  Without the index: plain binary search.
//   phase 2  block scan of the per-root output counts + chained scan across
//            CTAs -> exact output offsets, outputs stay ordered by root
//   phase 3  the tile's outputs are flattened over the CTA: every lane emits
//            one sampled neighbour (coalesced stores, near-contiguous loads
//            because a root's most recent entries are adjacent in the row)
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

constexpr int kTcsrTile = 256;

__device__ __forceinline__ int lower_bound_f(const float* __restrict__ ts, int lo, int hi,
                                             float key) {
  // first index in [lo,hi) with ts[idx] >= key
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(ts + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Same result through the skip index: coarse[b] = ts[16 b].  Blocks b with
// 16 b in [lo, hi) are monotone in b; the first such block whose head is
// >= key bounds the answer to the 16 entries before that head.
__device__ __forceinline__ int lower_bound_skip(const float* __restrict__ ts,
                                                const float* __restrict__ coarse, int lo, int hi,
                                                float key, long long nnz) {
  if (hi - lo <= 4) return lower_bound_f(ts, lo, hi, key);
  const int cb = lower_bound_f(coarse, (lo + 15) >> 4, (hi + 15) >> 4, key);
  if (cb == 0) return lo;
  const int base = (cb - 1) << 4;                       // aligned 16-entry block holding the answer
  const int flo = base > lo ? base : lo;
  const int fhi = (base + 16) < hi ? (base + 16) : hi;
  if ((long long)base + 16 > nnz) return lower_bound_f(ts, flo, fhi, key);
  const float4* p = reinterpret_cast<const float4*>(ts + base);
  const float4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
  const float v[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w,
                       v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
  int c = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) c += (base + i >= flo && base + i < fhi && v[i] < key) ? 1 : 0;
  return flo + c;
}

__global__ void __launch_bounds__(256) tcsr_index_kernel(const float* __restrict__ ts, long long nnz,
                                                         float* __restrict__ coarse) {
  const long long nb = (nnz + 15) >> 4;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nb;
       b += (long long)gridDim.x * blockDim.x)
    coarse[b] = ts[b << 4];
}

__global__ void __launch_bounds__(kTcsrTile)
    tcsr_sample_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const int32_t* __restrict__ eid, const float* __restrict__ ts,
                       const float* __restrict__ coarse, long long nnz, int num_nodes, const int32_t* __restrict__ root_nodes,
                       const float* __restrict__ root_ts, int R, int k, int strategy,
                       float offset, float duration, uint64_t seed,
                       int32_t* __restrict__ out_nbr, int32_t* __restrict__ out_col,
                       int32_t* __restrict__ out_eid, float* __restrict__ out_ts,
                       float* __restrict__ out_dts, int32_t* __restrict__ root_off,
                       int32_t* __restrict__ out_count, unsigned long long* __restrict__ ws) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_lo[kTcsrTile], s_hi[kTcsrTile], s_pref[kTcsrTile + 1];
  __shared__ float s_t[kTcsrTile];
  __shared__ int s_warp[kTcsrTile / 32];
  __shared__ long long s_tile_prefix;
  const int ntiles = (R + kTcsrTile - 1) / kTcsrTile;
  const int tile = lookback_take_tile(ws);
  if (tile >= ntiles) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gi = tile * kTcsrTile + tid;

  // phase 1
  int lo = 0, hi = 0;
  float tr = 0.f;
  if (gi < R) {
    const int n = root_nodes[gi];
    tr = root_ts[gi];
    if (n >= 0 && n < num_nodes) {
      const int rs = __ldg(indptr + n), re = __ldg(indptr + n + 1);
      const float t_hi = tr + offset;
      if (coarse) {
        hi = lower_bound_skip(ts, coarse, rs, re, t_hi, nnz);
        lo = duration > 0.f ? lower_bound_skip(ts, coarse, rs, hi, t_hi - duration, nnz) : rs;
      } else {
        hi = lower_bound_f(ts, rs, re, t_hi);
        lo = duration > 0.f ? lower_bound_f(ts, rs, hi, t_hi - duration) : rs;
      }
    }
  }
  const int cand = hi - lo;
  const int cnt = cand < k ? cand : k;
  s_lo[tid] = lo;
  s_hi[tid] = hi;
  s_t[tid] = tr;

  // phase 2: block exclusive scan of cnt
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  int wbase = 0;
  for (int w = 0; w < wid; ++w) wbase += s_warp[w];
  s_pref[tid] = wbase + incl - cnt;
  if (tid == kTcsrTile - 1) s_pref[kTcsrTile] = wbase + incl;
  __syncthreads();
  const int tile_total = s_pref[kTcsrTile];
  // warp-wide look-back: this kernel is DRAM-bound (ncu: 5.0 TB/s of sector traffic), and the
  // block-wide variant's 256 polling threads per CTA cost more than the shorter wait saves
  if (wid == 0) {
    const long long pre = lookback_prefix_warp(ws, tile, tile_total);
    if (lane == 0) {
      s_tile_prefix = pre;
      if (tile == ntiles - 1) {
        root_off[R] = (int32_t)(pre + tile_total);
        *out_count = (int32_t)(pre + tile_total);
      }
    }
  }
  __syncthreads();
  const long long base = s_tile_prefix;
  if (gi < R) root_off[gi] = (int32_t)(base + s_pref[tid]);

  // phase 3: flattened emit
  Philox rng(seed);
  constexpr int kU = 4;  // outputs per thread per round: all loads of a round are issued before its stores
  for (int q0 = tid; q0 < tile_total; q0 += kU * kTcsrTile) {
    int rr[kU], idx[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int q = q0 + u * kTcsrTile;
      rr[u] = -1;
      idx[u] = 0;
      if (q < tile_total) {
        int a = 0, b = kTcsrTile;  // largest r with s_pref[r] <= q
        while (b - a > 1) {
          int mid = (a + b) >> 1;
          if (s_pref[mid] <= q) a = mid;
          else b = mid;
        }
        const int r = a, j = q - s_pref[r];
        const int rlo = s_lo[r], rhi = s_hi[r];
        const int rc = rhi - rlo;
        if (strategy == TGN_SAMPLE_RECENT || rc <= k) {
          idx[u] = rhi - 1 - j;
        } else {
          uint4 x = rng((uint64_t)(tile * kTcsrTile + r), (uint64_t)j);
          idx[u] = rlo + (int)(x.x % (uint32_t)rc);
        }
        rr[u] = r;
      }
    }
    float tj[kU];
    int nb[kU], ei[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (rr[u] >= 0) {
        tj[u] = __ldg(ts + idx[u]);
        nb[u] = __ldg(indices + idx[u]);
        ei[u] = __ldg(eid + idx[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (rr[u] >= 0) {
        const long long o = base + q0 + u * kTcsrTile;
        __stcs(out_nbr + o, nb[u]);
        __stcs(out_eid + o, ei[u]);
        __stcs(out_ts + o, tj[u]);
        __stcs(out_dts + o, s_t[rr[u]] - tj[u]);
        __stcs(out_col + o, tile * kTcsrTile + rr[u]);
      }
    }
  }
}

}  // namespace tgn

using namespace tgn;

extern "C" {

int64_t tgn_tcsr_sample_ws_bytes(int32_t num_roots) {
  if (num_roots < 0) return 0;
  int ntiles = num_roots > 0 ? (num_roots + kTcsrTile - 1) / kTcsrTile : 1;
  return (int64_t)(ntiles + 1) * 8;
}

int64_t tgn_tcsr_index_len(int64_t nnz) { return nnz > 0 ? (nnz + 15) / 16 : 0; }

int32_t tgn_tcsr_build_index(const float* ts, int64_t nnz, float* coarse, void* stream) {
  TGN_REQUIRE(nnz >= 0, "tcsr_build_index: bad size");
  if (nnz == 0) return TGN_OK;
  TGN_REQUIRE(ts && coarse, "tcsr_build_index: NULL pointer");
  const long long nb = (nnz + 15) / 16;
  const int grid = (int)((nb + 255) / 256 < 148 * 8 ? (nb + 255) / 256 : 148 * 8);
  tcsr_index_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ts, nnz, coarse);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_tcsr_sample(const int32_t* indptr, const int32_t* indices, const int32_t* eid,
                        const float* ts, const float* coarse, int64_t nnz, int32_t num_nodes, const int32_t* root_nodes,
                        const float* root_ts, int32_t num_roots, int32_t k, int32_t strategy,
                        float offset, float duration, uint64_t seed, int32_t* out_nbr,
                        int32_t* out_col, int32_t* out_eid, float* out_ts, float* out_dts,
                        int32_t* root_off, int32_t* out_count, void* ws, void* stream) {
  TGN_REQUIRE(num_roots >= 0 && num_nodes > 0 && k >= 1, "tcsr_sample: bad sizes");
  TGN_REQUIRE(strategy == TGN_SAMPLE_RECENT || strategy == TGN_SAMPLE_UNIFORM,
              "tcsr_sample: unknown strategy %d", strategy);
  TGN_REQUIRE(indptr && indices && eid && ts && root_off && out_count && ws,
              "tcsr_sample: NULL pointer");
  TGN_REQUIRE(!coarse || ((uintptr_t)ts % 16 == 0 && nnz > 0),
              "tcsr_sample: the skip index needs a 16-byte aligned ts array and its length");
  cudaStream_t s = (cudaStream_t)stream;
  if (num_roots == 0) {
    TGN_CUDA(cudaMemsetAsync(root_off, 0, 4, s));
    TGN_CUDA(cudaMemsetAsync(out_count, 0, 4, s));
    return TGN_OK;
  }
  TGN_REQUIRE(root_nodes && root_ts && out_nbr && out_col && out_eid && out_ts && out_dts,
              "tcsr_sample: NULL buffer");
  const int ntiles = (num_roots + kTcsrTile - 1) / kTcsrTile;
  TGN_CUDA(cudaMemsetAsync(ws, 0, (size_t)(ntiles + 1) * 8, s));
  launch_k(tcsr_sample_kernel, dim3(ntiles), dim3(kTcsrTile), 0, s, 
      indptr, indices, eid, ts, coarse, (long long)nnz, num_nodes, root_nodes, root_ts, num_roots, k, strategy, offset,
      duration, seed, out_nbr, out_col, out_eid, out_ts, out_dts, root_off, out_count,
      (unsigned long long*)ws);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
