// LastNeighborLoader ring on the GPU (reference neighbor_loader.py:15-109).
//
// State (caller-owned, same layout and dtypes as the reference's tensors):
//   neighbors int64 [N,K], e_id int64 [N,K] (-1 = empty slot), t float [N,K]
//
// lookup : one pass over the K slots of every root.  The (root, slot) pairs of
//          a tile are flattened over the CTA so every lane carries a pair and
//          the loads of one row are contiguous; valid pairs are compacted in
//          (root order x slot order) with warp ballots + a chained scan across
//          CTAs, so the kernel reads every slot exactly once and writes every
//          output exactly once.
// insert : single CTA; the 2B (node, position) keys are sorted in shared
//          memory (bitonic, stable because the position is part of the key),
//          the thread at the end of each node run merges the last <=K new
//          events with the node's old row.
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

constexpr int kLookupThreads = 256;
constexpr int kLookupPairsPerThread = 8;
constexpr int kLookupTilePairs = kLookupThreads * kLookupPairsPerThread;  // 2048
constexpr int kMaxK = 64;
constexpr int kLookupThreePassTiles = 64;   // from here on count -> scan -> emit replaces the chained scan

constexpr int kCountThreads = 512;
constexpr int kCountTiles = 8;              // tiles of the emit pass per CTA of the count pass
constexpr int kCountVecs = 8;               // row loads a thread of the count pass keeps in flight

// floor(p / d) for 0 <= p < 2^20 / d by one (wide) multiply (mul = 2^20 / d + 1, host side)
__device__ __forceinline__ int fast_div20(int p, uint32_t mul) {
  return (int)(((unsigned long long)(uint32_t)p * mul) >> 20);
}

// Large launches (>= kLookupThreePassTiles tiles): COUNT pass.  One thread per root reads the root's e_id row
// with 128-bit loads (64-bit when K is odd) and counts its valid slots; a CTA covers kCountTiles tiles of the
// emit pass and leaves ws[1 + tile] = valid slots of the tile and ws[1 + ntiles + cta] = their sum.  The emit
// pass derives its offset from those two levels itself (no scan launch, no atomics, nothing to clear).
// Round 2 history (1M roots, K = 10): the count pass first used the emit pass's (root, slot)-pair mapping
// with runtime divisions (~50 instructions per PAIR, 39 us) and a one-CTA scan sat between the two passes;
// a (root, 16-byte chunk) mapping flattened over the CTA measured 33 us, a thread per root 26 us; L2
// eviction hints on the ring rows / output stores changed nothing; an emit pass that stages the rows through
// shared memory with 4/8-byte cp.async (five CTAs per SM instead of three) measured 155 us against 100 us;
// prefetch.global.L2 of the rows of tiles two waves ahead made both passes slower (count 25 -> 37 us, emit
// 101 -> 116 us).
template <int kVec>
__global__ void __launch_bounds__(kCountThreads)
    nbr_count_kernel(const int64_t* __restrict__ n_id, DevCount roots, int K, int tile_roots, int ntiles_host,
                     int64_t num_nodes, const int64_t* __restrict__ eids, unsigned long long* __restrict__ ws) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_cnt[kCountTiles];
  const int tid = threadIdx.x, lane = tid & 31;
  const int R = roots.get();
  if (tid < kCountTiles) s_cnt[tid] = 0;
  __syncthreads();
  const int t0 = blockIdx.x * kCountTiles;
  const int r0 = t0 * tile_roots;
  const int nr = max(0, min(kCountTiles * tile_roots, R - r0));
  for (int i0 = 0; i0 < nr; i0 += kCountThreads) {      // (whole warps enter: the reduction below shuffles)
    const int i = i0 + tid;
    int c = 0, q = 0;
    if (i < nr) {
      q = i / tile_roots;
      const int64_t node = n_id[r0 + i];
      if (node >= 0 && node < num_nodes) {
        const int64_t* row = eids + node * K;
        if (kVec == 2) {
          for (int s0 = 0; s0 < K; s0 += 2 * kCountVecs) {
            longlong2 v[kCountVecs];
#pragma unroll
            for (int u = 0; u < kCountVecs; ++u) {
              v[u] = make_longlong2(-1, -1);
              if (s0 + 2 * u < K) v[u] = *reinterpret_cast<const longlong2*>(row + s0 + 2 * u);
            }
#pragma unroll
            for (int u = 0; u < kCountVecs; ++u) c += (v[u].x >= 0) + (v[u].y >= 0);
          }
        } else {
          for (int s0 = 0; s0 < K; s0 += 2 * kCountVecs) {
            long long v[2 * kCountVecs];
#pragma unroll
            for (int u = 0; u < 2 * kCountVecs; ++u) {
              v[u] = -1;
              if (s0 + u < K) v[u] = row[s0 + u];
            }
#pragma unroll
            for (int u = 0; u < 2 * kCountVecs; ++u) c += v[u] >= 0;
          }
        }
      }
    }
    // consecutive roots share a tile almost always: one shared-memory add per warp
    const int q0 = __shfl_sync(0xffffffffu, q, 0);
    if (__ballot_sync(0xffffffffu, q != q0) == 0u) {
      c = warp_sum_i(c);
      if (lane == 0 && c) atomicAdd(&s_cnt[q0], c);
    } else if (c) {
      atomicAdd(&s_cnt[q], c);
    }
  }
  __syncthreads();
  if (tid < kCountTiles && t0 + tid < ntiles_host) ws[1 + t0 + tid] = (unsigned long long)s_cnt[tid];
  if (tid == 0) {
    long long tot = 0;
#pragma unroll
    for (int k = 0; k < kCountTiles; ++k) tot += s_cnt[k];
    ws[1 + ntiles_host + blockIdx.x] = (unsigned long long)tot;
  }
}

// kMode 0: single pass, output offsets by a chained scan across CTAs (decoupled look-back) -- one launch,
//           what the training step uses (a few tiles);
// kMode 2: EMIT pass of the large-launch path: the tile's exclusive offset is the sum of the count pass's
//           group totals below its group and of the tile counts below it inside the group.
// On launches of thousands of tiles the look-back serialises (ncu, round 1: DRAM 36 % busy, a third of the
// samples spinning on predecessor tiles); count -> emit re-reads 80 B of 488 B per root and has no
// inter-CTA dependency at all.
template <int kMode, int kThreads, int kPairs, int kMinCtas, typename IdT>
__global__ void __launch_bounds__(kThreads, kMinCtas)
    nbr_lookup_kernel(const int64_t* __restrict__ n_id, DevCount roots, int K, uint32_t k_mul, int tile_roots,
                      int ntiles_host, int64_t num_nodes, const int64_t* __restrict__ nbrs,
                      const int64_t* __restrict__ eids, const float* __restrict__ ts,
                      int64_t* __restrict__ out_nbr, int64_t* __restrict__ out_ctr,
                      int64_t* __restrict__ out_eid, float* __restrict__ out_t,
                      int32_t* __restrict__ root_off, int32_t* __restrict__ out_count,
                      uint32_t* __restrict__ l0, uint32_t* __restrict__ l1,
                      unsigned long long* __restrict__ ws) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_warp_tot[kThreads / 32];
  __shared__ long long s_warp_pre[kThreads / 32];
  const int R = roots.get();
  const int ntiles = (R + tile_roots - 1) / tile_roots;
  const int tile = kMode == 0 ? lookback_take_tile(ws) : (int)blockIdx.x;
  if (tile >= (ntiles > 0 ? ntiles : 1)) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int r0 = tile * tile_roots;
  const int nr = min(tile_roots, R - r0);
  const int npairs = nr > 0 ? nr * K : 0;

  // emit pass: this thread's share of the two-level prefix (loads issued ahead of the row gathers)
  long long pre = 0;
  if (kMode == 2) {
    const int g = tile / kCountTiles;
    for (int i = tid; i < g; i += kThreads) pre += (long long)ws[1 + ntiles_host + i];
    const int f = g * kCountTiles + tid;
    if (f < tile) pre += (long long)ws[1 + f];
  }

  // IdT = int32_t when num_nodes < 2^31: centre and neighbour ids are held in one register each, which takes
  // the kernel from 80 to 64 registers (four resident CTAs per SM instead of three)
  int64_t v_e[kPairs];
  IdT v_n[kPairs];
  float v_t[kPairs];
  unsigned ball[kPairs];
  int wtot = 0;
  // all loads of the thread's 8 pairs are issued before the first ballot: a ballot is a
  // convergence point the compiler does not move loads across, and one DRAM round trip per
  // pair (8 in sequence) was the kernel's largest stall
  IdT cc[kPairs];
#pragma unroll
  for (int j = 0; j < kPairs; ++j) {
    const int p = wid * (32 * kPairs) + j * 32 + lane;
    cc[j] = -1;
    if (p < npairs) {
      const int64_t c = n_id[r0 + fast_div20(p, k_mul)];
      cc[j] = (c >= 0 && c < num_nodes) ? (IdT)c : (IdT)-1;
    }
  }
#pragma unroll
  for (int j = 0; j < kPairs; ++j) {
    const int p = wid * (32 * kPairs) + j * 32 + lane;
    v_e[j] = -1;
    if (cc[j] >= 0) {
      const int64_t at = (int64_t)cc[j] * K + (p - fast_div20(p, k_mul) * K);
      v_e[j] = eids[at];
      v_n[j] = (IdT)nbrs[at];
      v_t[j] = ts[at];
    }
  }
#pragma unroll
  for (int j = 0; j < kPairs; ++j) {
    ball[j] = __ballot_sync(0xffffffffu, v_e[j] >= 0);
    wtot += __popc(ball[j]);
  }
  if (kMode == 2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
  }
  if (lane == 0) {
    s_warp_tot[wid] = wtot;
    if (kMode == 2) s_warp_pre[wid] = pre;
  }
  __syncthreads();
  long long tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) tot += s_warp_tot[w];
  long long base;
  if (kMode == 0) {
    base = lookback_prefix_block(ws, tile, tot);
  } else {
    base = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) base += s_warp_pre[w];
  }
  if (tid == 0 && tile == (ntiles > 0 ? ntiles : 1) - 1) {
    root_off[R] = (int32_t)(base + tot);
    *out_count = (int32_t)(base + tot);
  }
  for (int w = 0; w < wid; ++w) base += s_warp_tot[w];
#pragma unroll
  for (int j = 0; j < kPairs; ++j) {
    const long long pos = base + __popc(ball[j] & lanemask_lt());
    const int p = wid * (32 * kPairs) + j * 32 + lane;
    if (p < npairs) {
      const int r = fast_div20(p, k_mul);
      if (p == r * K) root_off[r0 + r] = (int32_t)pos;
      if ((ball[j] >> lane) & 1u) {
        out_nbr[pos] = (int64_t)v_n[j];
        out_ctr[pos] = (int64_t)cc[j];
        out_eid[pos] = v_e[j];
        out_t[pos] = v_t[j];
        if (l0) {
          const int64_t id = (int64_t)v_n[j];
          if (id >= 0 && id < num_nodes) {
            uint32_t old = atomicOr(&l0[id >> 5], 1u << (id & 31));
            if (old == 0) {
              int64_t g = id >> 10;
              atomicOr(&l1[g >> 5], 1u << (g & 31));
            }
          }
        }
      }
    }
    base += __popc(ball[j]);
  }
}

// ---------------------------------------------------------------------------
// insert
// ---------------------------------------------------------------------------
__device__ __forceinline__ void block_bitonic_sort(unsigned long long* s, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = s[i], b = s[ixj];
          bool asc = (i & k) == 0;
          if ((a > b) == asc) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// Merge of one node's run into its ring row by ONE WARP (K <= 32): lane i holds old slot i and the
// i-th newest entry of the run.  Every candidate computes its rank among all candidates with shuffles
// (descending value, ties to the lower candidate index: old slots before new entries, exactly the order
// a stable top-k of cat[old row, dense new block] yields, neighbor_loader.py:91-104) and writes itself
// to slot `rank` if rank < K.  e_id and t are ranked independently (neighbor_loader.py:99-100); the
// dense block's -1 padding (neighbor_loader.py:77-82) takes part in the t ranking.
__device__ __forceinline__ void insert_merge_warp(const unsigned long long* __restrict__ s_key, int q,
                                                  int64_t node, int B, int K, int64_t cur,
                                                  const int64_t* __restrict__ src,
                                                  const int64_t* __restrict__ dst,
                                                  const float* __restrict__ t,
                                                  int64_t* __restrict__ nbrs, int64_t* __restrict__ eids,
                                                  float* __restrict__ ts) {
  const int lane = threadIdx.x & 31;
  // run length (capped at K): entries q, q-1, ... with the same node
  int taken = 0;
  {
    const bool mine = lane < K && q - lane >= 0 && (int64_t)(s_key[q - lane] >> 16) == node;
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    taken = __ffs(~m) - 1;              // number of leading set bits (contiguous run)
    if (taken < 0 || taken > K) taken = K;
  }
  const bool has_old = lane < K, has_new = lane < taken;
  int64_t oe = -1, on = 0, ne = -1, nn = 0;
  float ot = -1.f, nt = -1.f;
  if (has_old) {
    oe = eids[node * K + lane];
    on = nbrs[node * K + lane];
    ot = ts[node * K + lane];
  }
  if (has_new) {
    const int j = (int)(s_key[q - lane] & 0xffffu);
    const int ev = j < B ? j : j - B;
    ne = cur + ev;
    nn = j < B ? src[ev] : dst[ev];
    nt = t[ev];
  }
  // Ranks in the total order "value descending; ties: old slots before new entries, old by slot, new by
  // run position (older first)".  Ring rows are written in rank order, so the old row is normally already
  // sorted (descending): then an old slot's rank is its index plus the new entries above it, a new entry's
  // rank a ballot over the old row plus a count over the few new entries -- O(taken) warp operations.  A
  // row that is not sorted (state loaded from elsewhere) takes the generic all-pairs loop.
  int r_oe = 0, r_ne = 0, r_ot = 0, r_nt = 0, cnt_ge = 0;
  const int64_t oe_next = __shfl_down_sync(0xffffffffu, oe, 1);
  const float ot_next = __shfl_down_sync(0xffffffffu, ot, 1);
  const bool unsorted = lane + 1 < K && (oe_next > oe || ot_next > ot);
  if (!__any_sync(0xffffffffu, unsorted)) {
    r_oe = lane;
    r_ot = lane;
    cnt_ge = __popc(__ballot_sync(0xffffffffu, has_old && ot >= -1.f)) +
             __popc(__ballot_sync(0xffffffffu, has_new && nt >= -1.f));
    for (int i = 0; i < taken; ++i) {
      const int64_t ye = __shfl_sync(0xffffffffu, ne, i);
      const float yt = __shfl_sync(0xffffffffu, nt, i);
      r_oe += (ye > oe);
      r_ot += (yt > ot);
      r_ne += (ye > ne) || (ye == ne && i > lane);
      r_nt += (yt > nt) || (yt == nt && i > lane);
      const int ge_e = __popc(__ballot_sync(0xffffffffu, has_old && oe >= ye));
      const int ge_t = __popc(__ballot_sync(0xffffffffu, has_old && ot >= yt));
      if (lane == i) {
        r_ne += ge_e;
        r_nt += ge_t;
      }
    }
  } else {
    for (int i = 0; i < 32; ++i) {
      const int64_t xe = __shfl_sync(0xffffffffu, oe, i), ye = __shfl_sync(0xffffffffu, ne, i);
      const float xt = __shfl_sync(0xffffffffu, ot, i), yt = __shfl_sync(0xffffffffu, nt, i);
      const bool vo = i < K, vn = i < taken;
      if (vo) {
        r_oe += (xe > oe) || (xe == oe && i < lane);
        r_ne += (xe >= ne);
        r_ot += (xt > ot) || (xt == ot && i < lane);
        r_nt += (xt >= nt);
        cnt_ge += (xt >= -1.f);
      }
      if (vn) {
        r_oe += (ye > oe);
        r_ne += (ye > ne) || (ye == ne && i > lane);
        r_ot += (yt > ot);
        r_nt += (yt > nt) || (yt == nt && i > lane);
        cnt_ge += (yt >= -1.f);
      }
    }
  }
  const int n_fill = K - taken;          // -1 padding of the dense block: after old and new on ties
  if (ot < -1.f) r_ot += n_fill;
  if (nt < -1.f) r_nt += n_fill;
  __syncwarp();
  if (has_old && r_oe < K) {
    eids[node * K + r_oe] = oe;
    nbrs[node * K + r_oe] = on;
  }
  if (has_new && r_ne < K) {
    eids[node * K + r_ne] = ne;
    nbrs[node * K + r_ne] = nn;
  }
  if (has_old && r_ot < K) ts[node * K + r_ot] = ot;
  if (has_new && r_nt < K) ts[node * K + r_nt] = nt;
  if (lane < n_fill && cnt_ge + lane < K) ts[node * K + cnt_ge + lane] = -1.f;
}

// Several CTAs: each sorts the 2B keys in its own shared memory (a few microseconds; cheaper than a grid
// barrier) and merges a strided share of the node runs, one warp per run.
__global__ void __launch_bounds__(1024, 1)
    nbr_insert_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                      const float* __restrict__ t, int B, int P, int64_t cur_host,
                      const int64_t* __restrict__ cur_dev, int K, int64_t num_nodes,
                      int64_t* __restrict__ nbrs, int64_t* __restrict__ eids,
                      float* __restrict__ ts) {
  pdl_wait();
  pdl_launch();
  extern __shared__ unsigned long long s_key[];
  const int n = 2 * B;
  const int64_t cur = cur_dev ? *cur_dev : cur_host;
  // nodes = cat[dst, src] (neighbor_loader.py:58): entry j < B is centred on dst[j]
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    unsigned long long key = ~0ull;
    if (j < n) {
      int64_t node = j < B ? dst[j] : src[j - B];
      key = ((unsigned long long)node << 16) | (unsigned)j;
    }
    s_key[j] = key;
  }
  __syncthreads();
  block_bitonic_sort(s_key, P);
  if (K <= 32) {
    const int warps = blockDim.x >> 5;
    for (int q = blockIdx.x * warps + (threadIdx.x >> 5); q < n; q += gridDim.x * warps) {
      const unsigned long long key = s_key[q];
      const int64_t node = (int64_t)(key >> 16);
      const bool run_end = (q == n - 1) || ((int64_t)(s_key[q + 1] >> 16) != node);
      if (!run_end || node < 0 || node >= num_nodes) continue;   // warp-uniform
      insert_merge_warp(s_key, q, node, B, K, cur, src, dst, t, nbrs, eids, ts);
    }
    return;
  }
  // K > 32: one thread per run with local candidate arrays (rare configuration)
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const unsigned long long key = s_key[q];
    const int64_t node = (int64_t)(key >> 16);
    const bool run_end = (q == n - 1) || ((int64_t)(s_key[q + 1] >> 16) != node);
    if (!run_end || node < 0 || node >= num_nodes) continue;
    int64_t ce[2 * kMaxK], cn[2 * kMaxK];
    float ct[2 * kMaxK];
    int m = 0;
    for (int s = 0; s < K; ++s) {
      ce[m] = eids[node * K + s];
      cn[m] = nbrs[node * K + s];
      ct[m] = ts[node * K + s];
      ++m;
    }
    int taken = 0;
    for (int i = 0; i < K && q - i >= 0; ++i) {
      const unsigned long long kq = s_key[q - i];
      if ((int64_t)(kq >> 16) != node) break;
      const int j = (int)(kq & 0xffffu);
      const int ev = j < B ? j : j - B;
      ce[m] = cur + ev;
      cn[m] = j < B ? src[ev] : dst[ev];
      ct[m] = t[ev];
      ++m;
      ++taken;
    }
    const int n_fill = K - taken;
    for (int a = 0; a < K; ++a) {
      int best = a;
      for (int b = a + 1; b < m; ++b)
        if (ce[b] > ce[best]) best = b;
      int64_t te = ce[a]; ce[a] = ce[best]; ce[best] = te;
      int64_t tn = cn[a]; cn[a] = cn[best]; cn[best] = tn;
    }
    float tt[3 * kMaxK];
    for (int a = 0; a < m; ++a) tt[a] = ct[a];
    int mt = m;
    for (int a = 0; a < n_fill; ++a) tt[mt++] = -1.0f;
    for (int a = 0; a < K; ++a) {
      int best = a;
      for (int b = a + 1; b < mt; ++b)
        if (tt[b] > tt[best]) best = b;
      float x = tt[a]; tt[a] = tt[best]; tt[best] = x;
    }
    for (int s = 0; s < K; ++s) {
      eids[node * K + s] = ce[s];
      nbrs[node * K + s] = cn[s];
      ts[node * K + s] = tt[s];
    }
  }
}

// the event counter moves after every CTA of the insert has read it (stream order)
__global__ void advance_i64_kernel(int64_t* p, int64_t by) {
  pdl_wait();
  pdl_launch();
  *p += by;
}

}  // namespace tgn

using namespace tgn;

extern "C" {

static inline int lookup_tile_roots(int K) { return kLookupTilePairs / K; }

int64_t tgn_nbr_lookup_ws_bytes(int32_t num_roots, int32_t size_k) {
  if (size_k < 1 || size_k > kMaxK || num_roots < 0) return 0;
  int tr = lookup_tile_roots(size_k);
  int ntiles = num_roots > 0 ? (num_roots + tr - 1) / tr : 1;
  // ticket | per-tile words | per-group totals of the count pass
  return (int64_t)(1 + ntiles + (ntiles + kCountTiles - 1) / kCountTiles) * 8;
}

int32_t tgn_nbr_lookup(const int64_t* n_id, int32_t num_roots, const int32_t* num_roots_dev,
                       int32_t size_k, int64_t num_nodes, const int64_t* neighbors,
                       const int64_t* e_id, const float* t, int64_t* out_nbr,
                       int64_t* out_centre, int64_t* out_eid, float* out_t, int32_t* root_off,
                       int32_t* out_count, void* bitmap, void* ws, void* stream) {
  TGN_REQUIRE(size_k >= 1 && size_k <= kMaxK, "nbr_lookup: size_k must be in [1,%d]", kMaxK);
  TGN_REQUIRE(num_roots >= 0 && num_nodes > 0, "nbr_lookup: bad sizes");
  TGN_REQUIRE(neighbors && e_id && t && root_off && out_count && ws, "nbr_lookup: NULL pointer");
  TGN_REQUIRE(num_roots == 0 || (n_id && out_nbr && out_centre && out_eid && out_t),
              "nbr_lookup: NULL buffer");
  cudaStream_t s = (cudaStream_t)stream;
  const int tr = lookup_tile_roots(size_k);
  const int ntiles = num_roots > 0 ? (num_roots + tr - 1) / tr : 1;
  const uint32_t k_mul = (1u << 20) / (uint32_t)size_k + 1u;   // fast_div20: pair index < 2048 <= 2^20 / 64
  uint32_t* l0 = (uint32_t*)bitmap;
  uint32_t* l1 = l0 ? l0 + ((num_nodes + 1023) / 1024) * 32 : nullptr;
  DevCount rc{num_roots_dev, num_roots};
  unsigned long long* w = (unsigned long long*)ws;
  if (ntiles >= kLookupThreePassTiles) {   // large launch: count -> emit, no inter-CTA dependency
    const int groups = (ntiles + kCountTiles - 1) / kCountTiles;
    if ((size_k & 1) == 0 && (reinterpret_cast<uintptr_t>(e_id) & 15) == 0)
      launch_k(nbr_count_kernel<2>, dim3(groups), dim3(kCountThreads), 0, s, n_id, rc, size_k, tr, ntiles,
               num_nodes, e_id, w);
    else
      launch_k(nbr_count_kernel<1>, dim3(groups), dim3(kCountThreads), 0, s, n_id, rc, size_k, tr, ntiles,
               num_nodes, e_id, w);
    TGN_LAUNCH_CHECK();
    if (num_nodes < (1ll << 31))
      launch_k(nbr_lookup_kernel<2, 256, 8, 4, int32_t>, dim3(ntiles), dim3(kLookupThreads), 0, s, n_id, rc, size_k,
               k_mul, tr, ntiles, num_nodes, neighbors, e_id, t, out_nbr, out_centre, out_eid, out_t, root_off,
               out_count, l0, l1, w);
    else
      launch_k(nbr_lookup_kernel<2, 256, 8, 3, int64_t>, dim3(ntiles), dim3(kLookupThreads), 0, s, n_id, rc, size_k,
               k_mul, tr, ntiles, num_nodes, neighbors, e_id, t, out_nbr, out_centre, out_eid, out_t, root_off,
               out_count, l0, l1, w);
    TGN_LAUNCH_CHECK();
    return TGN_OK;
  }
  TGN_CUDA(cudaMemsetAsync(ws, 0, (size_t)(ntiles + 1) * 8, s));
  launch_k(nbr_lookup_kernel<0, 256, 8, 3, int64_t>, dim3(ntiles), dim3(kLookupThreads), 0, s,
      n_id, rc, size_k, k_mul, tr, ntiles, num_nodes, neighbors, e_id, t, out_nbr, out_centre, out_eid, out_t,
      root_off, out_count, l0, l1, w);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_nbr_insert(const int64_t* src, const int64_t* dst, const float* t, int32_t batch,
                       int64_t cur_e_id, int64_t* cur_e_id_dev, int32_t size_k,
                       int64_t num_nodes, int64_t* neighbors, int64_t* e_id, float* t_state,
                       void* stream) {
  TGN_REQUIRE(size_k >= 1 && size_k <= kMaxK, "nbr_insert: size_k must be in [1,%d]", kMaxK);
  TGN_REQUIRE(batch >= 0 && 2 * (int64_t)batch <= TGN_SORT_MAX,
              "nbr_insert: 2*batch=%lld exceeds TGN_SORT_MAX=%d", 2ll * batch, TGN_SORT_MAX);
  if (batch == 0) return TGN_OK;
  TGN_REQUIRE(src && dst && t && neighbors && e_id && t_state, "nbr_insert: NULL pointer");
  int P = 2;
  while (P < 2 * batch) P <<= 1;
  static unsigned long long attr_mask = 0;
  TGN_CUDA(smem_optin(nbr_insert_kernel, TGN_SORT_MAX * 8, attr_mask));
  // one warp merges one node run: ~2B/3 runs on TGB-like streams; 32 warps per CTA
  int grid = (2 * batch + 32 * 4 - 1) / (32 * 4);
  grid = grid < 1 ? 1 : (grid > 32 ? 32 : grid);
  launch_k(nbr_insert_kernel, dim3(grid), dim3(1024), (size_t)P * 8, (cudaStream_t)stream,
      src, dst, t, batch, P, cur_e_id, (const int64_t*)cur_e_id_dev, size_k, num_nodes, neighbors, e_id, t_state);
  TGN_LAUNCH_CHECK();
  if (cur_e_id_dev) {
    launch_k(advance_i64_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, cur_e_id_dev, (int64_t)batch);
    TGN_LAUNCH_CHECK();
  }
  return TGN_OK;
}

}  // extern "C"
