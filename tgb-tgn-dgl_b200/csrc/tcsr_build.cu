// t-CSR construction on the GPU: the file TGL's gen_graph.py writes as
// DATA/<name>/ext_full.npz (keys indptr/indices/ts/eid, reference utils.py:73, README.md:4-5;
// the generator itself is absent from the reference tree).  Upstream builds it with
// a Python row loop plus a per-row argsort; here it is one stable LSD radix sort of the
// 2*E directed entries by (node, timestamp) followed by one emit pass.
//
//   entry j = 2*e + dir   (dir 0: src->dst, dir 1: dst->src, same eid)   [add_reverse]
//   key     = node << 32 | ordered_bits(float32(t[e]))
//   The sort is stable and entries enter in j order, so ties on (node, ts) keep eid order and
//   forward-before-reverse: rows come out sorted by (ts, eid) -- the order the oracle's
//   numpy lexsort produces (oracle/tgn_oracle.py::build_tcsr).
//   If the caller states that t is non-decreasing (TGB event streams are), entry order already
//   is (ts, eid) order and only the node digits are sorted (3 passes instead of 7).
//
// One radix pass (8-bit digit, tile = 256 threads x 16 keys):
//   hist    per-tile digit histogram -> hist[digit][tile]
//   scan    block d scans row d of the histogram (exclusive) and leaves the digit total
//   scatter re-reads the tile; warp w ranks its 512 consecutive keys in 16 rounds of 32
//           (match_any on the digit: rank = earlier lanes with the same digit + running
//           per-warp digit counter in shared memory), warps are chained per digit, and every
//           key goes to  digit_base + tile_prefix + warp_prefix + rank  -- stable by construction.
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 keys
constexpr int kRsWarps = kRsThreads / 32;

__device__ __forceinline__ uint32_t f32_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

template <typename T>
__global__ void __launch_bounds__(256)
    tcsr_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                     const T* __restrict__ t, long long E, int add_reverse, int num_nodes,
                     unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                     int32_t* __restrict__ bad) {
  const long long n = add_reverse ? 2 * E : E;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const long long e = add_reverse ? (j >> 1) : j;
    const int dir = add_reverse ? (int)(j & 1) : 0;
    const int64_t node = dir ? dst[e] : src[e];
    const int64_t other = dir ? src[e] : dst[e];
    if (node < 0 || node >= num_nodes || other < 0 || other >= num_nodes) *bad = 1;
    keys[j] = ((unsigned long long)(uint32_t)node << 32) | f32_ordered((float)t[e] + 0.0f);  // -0 -> +0
    vals[j] = (uint32_t)j;
  }
}

__global__ void __launch_bounds__(kRsThreads)
    rs_hist_kernel(const unsigned long long* __restrict__ keys, long long n, int shift, int ntiles,
                   uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_h[256];
  const int tile = blockIdx.x;
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)tile * kRsTile;
#pragma unroll 4
  for (int i = 0; i < kRsItems; ++i) {
    const long long p = base + i * kRsThreads + threadIdx.x;
    if (p < n) atomicAdd(&s_h[(uint32_t)(keys[p] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * ntiles + tile] = s_h[threadIdx.x];
}

// block d: exclusive scan of hist[d][0..ntiles) in place; total[d] = row sum
__global__ void __launch_bounds__(256)
    rs_scan_kernel(uint32_t* __restrict__ hist, int ntiles, uint32_t* __restrict__ total) {
  __shared__ uint32_t s_w[8];
  __shared__ uint32_t s_carry;
  uint32_t* row = hist + (size_t)blockIdx.x * ntiles;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 < ntiles; c0 += 256) {
    const int i = c0 + tid;
    const uint32_t v = i < ntiles ? row[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_w[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += s_w[w];
    const uint32_t carry = s_carry;
    if (i < ntiles) row[i] = carry + wbase + incl - v;
    __syncthreads();
    if (tid == 255) s_carry = carry + wbase + incl;
    __syncthreads();
  }
  if (tid == 0) total[blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__(kRsThreads)
    rs_scatter_kernel(const unsigned long long* __restrict__ keys_in,
                      const uint32_t* __restrict__ vals_in, long long n, int shift, int ntiles,
                      const uint32_t* __restrict__ hist, const uint32_t* __restrict__ total,
                      unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t s_cnt[kRsWarps][256];   // per-warp running digit counters
  __shared__ uint32_t s_base[256];            // global base of every digit for this tile
  const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // digit bases: exclusive scan of the 256 digit totals (every CTA redoes it: 256 adds)
  {
    const uint32_t v = total[tid];
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    s_cnt[0][tid] = incl;  // scratch
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += s_cnt[0][w * 32 + 31];
    const uint32_t excl = wbase + incl - v;
    __syncthreads();
    s_base[tid] = excl + hist[(size_t)tid * ntiles + tile];
  }
#pragma unroll
  for (int w = 0; w < kRsWarps; ++w) s_cnt[w][tid] = 0;
  __syncthreads();

  unsigned long long k[kRsItems];
  uint32_t rank[kRsItems];
  const long long wbeg = (long long)tile * kRsTile + (long long)wid * (32 * kRsItems);
#pragma unroll
  for (int i = 0; i < kRsItems; ++i) {
    const long long p = wbeg + i * 32 + lane;
    k[i] = p < n ? keys_in[p] : ~0ull;
  }
#pragma unroll
  for (int i = 0; i < kRsItems; ++i) {
    const long long p = wbeg + i * 32 + lane;
    const bool ok = p < n;
    const uint32_t d = (uint32_t)(k[i] >> shift) & 255u;
    // lanes past the end vote with an impossible digit so they never share a group
    const unsigned peers = __match_any_sync(0xffffffffu, ok ? d : 256u + (uint32_t)lane);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader && ok) {
      base = s_cnt[wid][d];
      s_cnt[wid][d] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[i] = base + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  // chain the warps per digit: thread d turns the per-warp counts into exclusive prefixes
  {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      const uint32_t c = s_cnt[w][tid];
      s_cnt[w][tid] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kRsItems; ++i) {
    const long long p = wbeg + i * 32 + lane;
    if (p < n) {
      const uint32_t d = (uint32_t)(k[i] >> shift) & 255u;
      const uint32_t pos = s_base[d] + s_cnt[wid][d] + rank[i];
      keys_out[pos] = k[i];
      vals_out[pos] = vals_in[p];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    tcsr_emit_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals,
                     long long n, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                     const T* __restrict__ t, int add_reverse, int num_nodes,
                     int32_t* __restrict__ indptr, int32_t* __restrict__ indices,
                     int32_t* __restrict__ eid, float* __restrict__ ts) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const uint32_t j = vals[i];
    const uint32_t e = add_reverse ? (j >> 1) : j;
    const int dir = add_reverse ? (int)(j & 1u) : 0;
    const int node = (int)(keys[i] >> 32);
    indices[i] = (int32_t)(dir ? src[e] : dst[e]);
    eid[i] = (int32_t)e;
    ts[i] = (float)t[e];
    // row starts: the first entry of a node closes every (empty) row since the previous node
    const int prev = i > 0 ? (int)(keys[i - 1] >> 32) : -1;
    for (int r = prev + 1; r <= node; ++r) indptr[r] = (int32_t)i;
    if (i == n - 1)
      for (int r = node + 1; r <= num_nodes; ++r) indptr[r] = (int32_t)n;
  }
}

__global__ void tcsr_fill_indptr_kernel(int32_t* indptr, int num_nodes) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= num_nodes; i += gridDim.x * blockDim.x)
    indptr[i] = 0;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace tgn

using namespace tgn;

extern "C" {

int64_t tgn_tcsr_build_ws_bytes(int64_t num_events, int32_t add_reverse) {
  if (num_events < 0) return 0;
  const size_t n = (size_t)(add_reverse ? 2 * num_events : num_events);
  const size_t ntiles = (n + kRsTile - 1) / kRsTile;
  return (int64_t)(2 * align256(n * 8) + 2 * align256(n * 4) + align256(256 * (ntiles + 1) * 4) +
                   align256(256 * 4) + 256);
}

int32_t tgn_tcsr_build(const int64_t* src, const int64_t* dst, const void* t, int32_t t_is_float,
                       int64_t num_events, int32_t num_nodes, int32_t add_reverse,
                       int32_t t_sorted, int32_t* indptr, int32_t* indices, int32_t* eid, float* ts,
                       int32_t* bad_flag, void* ws, void* stream) {
  TGN_REQUIRE(num_events >= 0 && num_nodes > 0, "tcsr_build: bad sizes");
  TGN_REQUIRE((add_reverse ? 2 * num_events : num_events) < (1ll << 31),
              "tcsr_build: %lld entries do not fit the int32 t-CSR format", (long long)num_events);
  TGN_REQUIRE(indptr, "tcsr_build: NULL indptr");
  cudaStream_t s = (cudaStream_t)stream;
  if (bad_flag) TGN_CUDA(cudaMemsetAsync(bad_flag, 0, 4, s));
  if (num_events == 0) {
    tcsr_fill_indptr_kernel<<<stride_grid(num_nodes + 1, 256), 256, 0, s>>>(indptr, num_nodes);
    TGN_LAUNCH_CHECK();
    return TGN_OK;
  }
  TGN_REQUIRE(src && dst && t && indices && eid && ts && ws && bad_flag, "tcsr_build: NULL pointer");
  const long long n = add_reverse ? 2 * num_events : num_events;
  const int ntiles = (int)((n + kRsTile - 1) / kRsTile);
  uint8_t* w = (uint8_t*)ws;
  unsigned long long* keys[2];
  uint32_t* vals[2];
  keys[0] = (unsigned long long*)w; w += align256((size_t)n * 8);
  keys[1] = (unsigned long long*)w; w += align256((size_t)n * 8);
  vals[0] = (uint32_t*)w; w += align256((size_t)n * 4);
  vals[1] = (uint32_t*)w; w += align256((size_t)n * 4);
  uint32_t* hist = (uint32_t*)w; w += align256((size_t)256 * (ntiles + 1) * 4);
  uint32_t* total = (uint32_t*)w;

  const int g = stride_grid(n, 256, 16);
  if (t_is_float)
    tcsr_keys_kernel<float><<<g, 256, 0, s>>>(src, dst, (const float*)t, num_events, add_reverse,
                                              num_nodes, keys[0], vals[0], bad_flag);
  else
    tcsr_keys_kernel<int64_t><<<g, 256, 0, s>>>(src, dst, (const int64_t*)t, num_events,
                                                add_reverse, num_nodes, keys[0], vals[0], bad_flag);
  TGN_LAUNCH_CHECK();

  int node_bits = 1;
  while (node_bits < 32 && ((long long)1 << node_bits) < (long long)num_nodes) ++node_bits;
  int cur = 0;
  for (int shift = t_sorted ? 32 : 0; shift < 32 + node_bits; shift += 8) {
    rs_hist_kernel<<<ntiles, kRsThreads, 0, s>>>(keys[cur], n, shift, ntiles, hist);
    rs_scan_kernel<<<256, 256, 0, s>>>(hist, ntiles, total);
    rs_scatter_kernel<<<ntiles, kRsThreads, 0, s>>>(keys[cur], vals[cur], n, shift, ntiles, hist,
                                                    total, keys[cur ^ 1], vals[cur ^ 1]);
    TGN_LAUNCH_CHECK();
    cur ^= 1;
  }
  if (t_is_float)
    tcsr_emit_kernel<float><<<g, 256, 0, s>>>(keys[cur], vals[cur], n, src, dst, (const float*)t,
                                              add_reverse, num_nodes, indptr, indices, eid, ts);
  else
    tcsr_emit_kernel<int64_t><<<g, 256, 0, s>>>(keys[cur], vals[cur], n, src, dst,
                                                (const int64_t*)t, add_reverse, num_nodes, indptr,
                                                indices, eid, ts);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
