// Dependency-aware block ids (reference dependencyGraph.py:8-28 get_block, :33-49 dependecyAwareBatch).
//
// Walking one DataLoader batch in order, an event's block id is one more than the highest block id
// already given to either of its endpoints inside this batch (0 if neither was seen).  The reference
// does this with Python dicts and `.item()` per node -- O(events) interpreter time, hours on the large
// TGB shapes.  Every batch is independent of every other, and inside a batch the recurrence is a
// longest-path problem on a DAG in which each event has at most two predecessors (the previous event of
// its source, the previous event of its destination):
//
//     block[i] = 1 + max(block[prev_src(i)], block[prev_dst(i)]),   block[none] = -1
//
// One CTA per batch:
//   1. 2n keys (node << 21 | 2*i + side) are sorted in shared memory (bitonic); the entry in front of an
//      entry with the same node is that endpoint's previous event (an event that touches the same node
//      twice, src == dst, skips its own sibling entry);
//   2. the recurrence is relaxed in shared memory until nothing changes (depth-of-the-DAG + 1 rounds; a
//      hub node touched by every event of the batch is the worst case, n rounds of n/blockDim work).
// Bit-exact integer work; HBM traffic = 16 B read + 4 B written per event.
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

constexpr int kDepThreads = 512;
constexpr int kDepPosBits = 21;  // 2 * batch <= 2^21

__device__ __forceinline__ void dep_bitonic_sort(unsigned long long* s, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s[i], b = s[ixj];
          const bool asc = (i & k) == 0;
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

__global__ void __launch_bounds__(kDepThreads)
    dep_blocks_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                      int64_t num_events, int B, int P, int32_t* __restrict__ out,
                      int32_t* __restrict__ num_blocks) {
  pdl_wait();
  pdl_launch();
  extern __shared__ unsigned long long s_key[];            // [P]
  int32_t* s_pred = reinterpret_cast<int32_t*>(s_key + P);  // [2B] predecessor event of (event, side)
  int32_t* s_blk = s_pred + 2 * B;                          // [B]
  const unsigned long long kPosMask = (1ull << kDepPosBits) - 1;
  for (int64_t batch = blockIdx.x; batch * B < num_events; batch += gridDim.x) {
    const int64_t lo = batch * (int64_t)B;
    const int n = (int)((num_events - lo) < B ? (num_events - lo) : B);
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
      unsigned long long key = ~0ull;
      if (j < 2 * n) {
        const int64_t node = (j & 1) ? dst[lo + (j >> 1)] : src[lo + (j >> 1)];
        key = ((unsigned long long)node << kDepPosBits) | (unsigned long long)j;
      }
      s_key[j] = key;
    }
    __syncthreads();
    dep_bitonic_sort(s_key, P);
    for (int q = threadIdx.x; q < 2 * n; q += blockDim.x) {
      const unsigned long long key = s_key[q];
      const int j = (int)(key & kPosMask);
      const unsigned long long node = key >> kDepPosBits;
      int pred = -1;
      int p = q - 1;
      if (p >= 0 && (s_key[p] >> kDepPosBits) == node && (int)((s_key[p] & kPosMask) >> 1) == (j >> 1)) --p;
      if (p >= 0 && (s_key[p] >> kDepPosBits) == node) pred = (int)((s_key[p] & kPosMask) >> 1);
      s_pred[j] = pred;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_blk[i] = 0;
    __syncthreads();
    int changed = 1;
    while (changed) {
      int mine = 0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int a = s_pred[2 * i], b = s_pred[2 * i + 1];
        const int ba = a >= 0 ? s_blk[a] : -1, bb = b >= 0 ? s_blk[b] : -1;
        const int v = (ba > bb ? ba : bb) + 1;
        if (v != s_blk[i]) {   // monotone: a racing read of an older value only delays convergence
          s_blk[i] = v;
          mine = 1;
        }
      }
      changed = __syncthreads_or(mine);
    }
    int mx = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int v = s_blk[i];
      out[lo + i] = v;
      mx = v > mx ? v : mx;
    }
    if (num_blocks) {
      mx = __reduce_max_sync(0xffffffffu, mx);
      if ((threadIdx.x & 31) == 0) atomicMax(num_blocks + batch, mx + 1);
    }
    __syncthreads();
  }
}

}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_dep_blocks(const int64_t* src, const int64_t* dst, int64_t num_events, int32_t batch,
                       int32_t* block_ids, int32_t* num_blocks, void* stream) {
  TGN_REQUIRE(num_events >= 0 && batch >= 1, "dep_blocks: bad sizes");
  TGN_REQUIRE(2 * (int64_t)batch <= TGN_SORT_MAX, "dep_blocks: batch %d exceeds TGN_SORT_MAX/2 = %d", batch,
              TGN_SORT_MAX / 2);
  if (num_events == 0) return TGN_OK;
  TGN_REQUIRE(src && dst && block_ids, "dep_blocks: NULL pointer");
  int P = 2;
  while (P < 2 * batch) P <<= 1;
  const size_t smem = (size_t)P * 8 + (size_t)batch * 12;
  static unsigned long long attr_mask = 0;
  TGN_CUDA(smem_optin(dep_blocks_kernel, TGN_SORT_MAX * 8 + (TGN_SORT_MAX / 2) * 12, attr_mask));
  const long long nb = (num_events + batch - 1) / batch;
  cudaStream_t s = (cudaStream_t)stream;
  if (num_blocks) TGN_CUDA(cudaMemsetAsync(num_blocks, 0, (size_t)nb * sizeof(int32_t), s));
  const int per_sm = smem <= 24 * 1024 ? 4 : (smem <= 56 * 1024 ? 2 : 1);
  const int grid = (int)(nb < (long long)kNumSMs * per_sm ? nb : (long long)kNumSMs * per_sm);
  launch_k(dep_blocks_kernel, dim3(grid), dim3(kDepThreads), smem, s, src, dst, num_events, batch, P,
           block_ids, num_blocks);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
