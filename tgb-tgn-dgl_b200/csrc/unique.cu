// Sorted-unique + relabel by two-level bitmap ranking, and the library's error
// plumbing.  Replaces torch.cat(...).unique() + `_assoc[n_id] = arange(...)`
// (reference neighbor_loader.py:46-48, modules/memory_module.py:129,153).
//
// Layout of `bitmap` (uint32 words):
//   L0: one bit per node, ceil(N/32) words, padded to a whole 32-word group
//   L1: one bit per 32-word L0 group (= 1024 nodes)
// Marking sets both levels; ranking walks only the groups whose L1 bit is set,
// so its cost follows the number of touched groups, not N.
#include <stdarg.h>

#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {
static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }
int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int& pdl_flag() {
  static int flag = 1;
  return flag;
}

int32_t* dev_err_word() {
  static int32_t* word = [] {
    int32_t* p = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&p), sizeof(int32_t),
                      cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess)
      return static_cast<int32_t*>(nullptr);
    *p = 0;
    return p;
  }();
  return word;
}

static inline int64_t l0_groups(int64_t n) { return (n + 1023) / 1024; }
static inline int64_t l0_words(int64_t n) { return l0_groups(n) * 32; }
static inline int64_t l1_words(int64_t n) { return (l0_groups(n) + 31) / 32; }

__global__ void unique_mark_kernel(const int64_t* __restrict__ ids, DevCount cnt, int64_t n_nodes,
                                   uint32_t* __restrict__ l0, uint32_t* __restrict__ l1) {
  pdl_wait();
  pdl_launch();
  const int n = cnt.get();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int64_t id = ids[i];
    if (id < 0 || id >= n_nodes) continue;  // out-of-range ids are ignored (caller validates)
    uint32_t bit = 1u << (id & 31);
    uint32_t old = atomicOr(&l0[id >> 5], bit);
    if (old == 0) {
      int64_t g = id >> 10;
      atomicOr(&l1[g >> 5], 1u << (g & 31));
    }
  }
}

// One CTA of 1024 threads, 1024 groups (= 2^20 nodes) per round.
//   phase 1  thread t counts the set bits of group t (only if its L1 bit is set), block scan
//            of the counts -> first output position of every group
//   phase 2  one warp per live group, lane = word of the group: warp prefix over the word
//            popcounts, every lane emits the ids of its own word in ascending order -- ids come
//            out globally sorted and the emission runs 32 lanes wide even when one hub group
//            holds most of the batch
__global__ void __launch_bounds__(1024, 1)
    unique_rank_kernel(uint32_t* __restrict__ l0, uint32_t* __restrict__ l1, int64_t n_groups,
                       int64_t n_l1_words, int64_t* __restrict__ out_ids, int32_t out_cap,
                       int64_t* __restrict__ assoc, int32_t* __restrict__ out_count,
                       int keep_marks, const int64_t* __restrict__ mark_ids, int mark_count,
                       int64_t n_nodes) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_warp[32];
  __shared__ int s_off[1024];
  __shared__ int s_base, s_total;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  // optional marking pass (the unique_mark launch folded in: this kernel is a single CTA, so a
  // block barrier orders the bitmap writes before the ranking reads)
  for (int i = tid; i < mark_count; i += 1024) {
    const int64_t id = mark_ids[i];
    if (id < 0 || id >= n_nodes) continue;
    const uint32_t old = atomicOr(&l0[id >> 5], 1u << (id & 31));
    if (old == 0) {
      const int64_t g = id >> 10;
      atomicOr(&l1[g >> 5], 1u << (g & 31));
    }
  }
  __threadfence_block();
  __syncthreads();
  for (int64_t g0 = 0; g0 < n_groups; g0 += 1024) {
    const int64_t g = g0 + tid;
    int cnt = 0;
    if (g < n_groups && ((l1[g >> 5] >> (g & 31)) & 1u)) {
      const uint4* p = reinterpret_cast<const uint4*>(l0 + g * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 v = p[q];
        cnt += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
      }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int v = s_warp[lane];
      int iv = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, iv, o);
        if (lane >= o) iv += y;
      }
      s_warp[lane] = iv - v;  // exclusive warp offsets
      if (lane == 31) s_total = iv;
    }
    __syncthreads();
    // exclusive offset of the group, or -1 for an empty group
    s_off[tid] = cnt > 0 ? s_base + s_warp[wid] + (incl - cnt) : -1;
    __syncthreads();
    for (int gl = wid; gl < 1024 && g0 + gl < n_groups; gl += 32) {
      const int gbase = s_off[gl];
      if (gbase < 0) continue;  // warp-uniform
      uint32_t* wp = l0 + (g0 + gl) * 32 + lane;
      uint32_t bits = *wp;
      const int c = __popc(bits);
      int pre = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += y;
      }
      int pos = gbase + pre - c;
      const int64_t id0 = ((g0 + gl) << 10) + (lane << 5);
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const int64_t id = id0 + b;
        if (pos < out_cap) out_ids[pos] = id;
        if (assoc) assoc[id] = pos;
        ++pos;
      }
      if (!keep_marks && c) *wp = 0u;
    }
    __syncthreads();
    if (tid == 0) s_base += s_total;
    if (tid < 32) {
      int64_t wi = (g0 >> 5) + tid;
      if (wi < n_l1_words && !keep_marks) l1[wi] = 0;
    }
    __syncthreads();
  }
  if (tid == 0) *out_count = s_base;
}

// Several CTAs: each CTA recomputes the per-group counts and their scan (a few KB of bitmap -- cheaper than a
// grid barrier) and emits a strided share of the live groups, one warp per group as above.  Nothing is cleared
// here: other CTAs may still be counting the same words; unique_clear_kernel follows when the marks are not kept.
__global__ void __launch_bounds__(1024, 1)
    unique_rank_mc_kernel(const uint32_t* __restrict__ l0, const uint32_t* __restrict__ l1, int64_t n_groups,
                          int64_t* __restrict__ out_ids, int32_t out_cap, int64_t* __restrict__ assoc,
                          int32_t* __restrict__ out_count) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_warp[32];
  __shared__ int s_off[1024];
  __shared__ int s_base, s_total;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int64_t g0 = 0; g0 < n_groups; g0 += 1024) {
    const int64_t g = g0 + tid;
    int cnt = 0;
    if (g < n_groups && ((l1[g >> 5] >> (g & 31)) & 1u)) {
      const uint4* p = reinterpret_cast<const uint4*>(l0 + g * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 v = p[q];
        cnt += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
      }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int v = s_warp[lane];
      int iv = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, iv, o);
        if (lane >= o) iv += y;
      }
      s_warp[lane] = iv - v;
      if (lane == 31) s_total = iv;
    }
    __syncthreads();
    s_off[tid] = cnt > 0 ? s_base + s_warp[wid] + (incl - cnt) : -1;
    __syncthreads();
    for (int gl = blockIdx.x * 32 + wid; gl < 1024 && g0 + gl < n_groups; gl += gridDim.x * 32) {
      const int gbase = s_off[gl];
      if (gbase < 0) continue;  // warp-uniform
      uint32_t bits = l0[(g0 + gl) * 32 + lane];
      const int c = __popc(bits);
      int pre = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += y;
      }
      int pos = gbase + pre - c;
      const int64_t id0 = ((g0 + gl) << 10) + (lane << 5);
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const int64_t id = id0 + b;
        if (pos < out_cap) out_ids[pos] = id;
        if (assoc) assoc[id] = pos;
        ++pos;
      }
    }
    __syncthreads();
    if (tid == 0) s_base += s_total;
    __syncthreads();
  }
  if (blockIdx.x == 0 && tid == 0) *out_count = s_base;
}

// one CTA per L1 word (32 groups x 32 words): clears the words of the marked groups, then the L1 word itself
__global__ void __launch_bounds__(1024) unique_clear_kernel(uint32_t* __restrict__ l0, uint32_t* __restrict__ l1,
                                                            int64_t n_groups) {
  pdl_wait();
  pdl_launch();
  const int64_t w1 = blockIdx.x;
  const uint32_t live = l1[w1];
  const int grp = threadIdx.x >> 5, word = threadIdx.x & 31;
  const int64_t g = w1 * 32 + grp;
  if (((live >> grp) & 1u) && g < n_groups) l0[g * 32 + word] = 0u;
  __syncthreads();
  if (threadIdx.x == 0 && live) l1[w1] = 0u;
}

__global__ void relabel_kernel(const int64_t* __restrict__ ids, DevCount cnt,
                               const int64_t* __restrict__ assoc, int64_t* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  const int n = cnt.get();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = assoc[ids[i]];
}
}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_abi_version(void) { return TGN_ABI_VERSION; }
const char* tgn_last_error(void) { return tgn::err_buf(); }
int32_t tgn_device_errors(int32_t reset) {
  volatile int32_t* w = tgn::dev_err_word();
  if (!w) return 0;
  const int32_t v = *w;
  if (reset) *w = 0;
  return v;
}
int32_t tgn_set_pdl(int32_t enabled) {
  const int old = tgn::pdl_flag();
  tgn::pdl_flag() = enabled ? 1 : 0;
  return old;
}

int32_t tgn_memcpy_async(void* dst, const void* src, int64_t nbytes, void* stream) {
  TGN_REQUIRE(nbytes >= 0 && (nbytes == 0 || (dst && src)), "memcpy_async: bad arguments");
  if (nbytes == 0) return TGN_OK;
  TGN_CUDA(cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return TGN_OK;
}

int64_t tgn_bitmap_bytes(int64_t num_nodes) {
  if (num_nodes < 0) return 0;
  return (l0_words(num_nodes) + l1_words(num_nodes)) * 4;
}

int32_t tgn_unique_mark(const int64_t* ids, int32_t count, const int32_t* count_dev,
                        int64_t num_nodes, void* bitmap, void* stream) {
  TGN_REQUIRE(count >= 0 && num_nodes > 0 && bitmap, "unique_mark: bad arguments");
  if (count == 0) return TGN_OK;
  TGN_REQUIRE(ids, "unique_mark: ids is NULL");
  uint32_t* l0 = (uint32_t*)bitmap;
  uint32_t* l1 = l0 + l0_words(num_nodes);
  DevCount c{count_dev, count};
  launch_k(unique_mark_kernel, dim3(stride_grid(count, 256)), dim3(256), 0, (cudaStream_t)stream, ids, c, num_nodes,
                                                                                 l0, l1);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_unique_rank(void* bitmap, int64_t num_nodes, int64_t* out_ids, int32_t out_cap,
                        int64_t* assoc, int32_t* out_count, int32_t keep_marks, void* stream) {
  TGN_REQUIRE(bitmap && out_ids && out_count && num_nodes > 0 && out_cap >= 0,
              "unique_rank: bad arguments");
  uint32_t* l0 = (uint32_t*)bitmap;
  uint32_t* l1 = l0 + l0_words(num_nodes);
  const int64_t ng = l0_groups(num_nodes);
  if (ng >= 16) {   // enough groups to share: several CTAs emit, a second launch clears the marks
    int grid = (int)((ng + 3) / 4);
    grid = grid > 16 ? 16 : grid;
    launch_k(unique_rank_mc_kernel, dim3(grid), dim3(1024), 0, (cudaStream_t)stream, (const uint32_t*)l0,
             (const uint32_t*)l1, ng, out_ids, out_cap, assoc, out_count);
    TGN_LAUNCH_CHECK();
    if (!keep_marks) {
      launch_k(unique_clear_kernel, dim3((unsigned)l1_words(num_nodes)), dim3(1024), 0, (cudaStream_t)stream, l0, l1, ng);
      TGN_LAUNCH_CHECK();
    }
    return TGN_OK;
  }
  launch_k(unique_rank_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream,
      l0, l1, ng, l1_words(num_nodes), out_ids, out_cap, assoc, out_count,
      keep_marks, nullptr, 0, num_nodes);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_unique_mark_rank(const int64_t* ids, int32_t count, void* bitmap, int64_t num_nodes,
                             int64_t* out_ids, int32_t out_cap, int64_t* assoc, int32_t* out_count,
                             int32_t keep_marks, void* stream) {
  TGN_REQUIRE(bitmap && out_ids && out_count && num_nodes > 0 && out_cap >= 0 && count >= 0 &&
                  (ids || count == 0),
              "unique_mark_rank: bad arguments");
  uint32_t* l0 = (uint32_t*)bitmap;
  uint32_t* l1 = l0 + l0_words(num_nodes);
  launch_k(unique_rank_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, 
      l0, l1, l0_groups(num_nodes), l1_words(num_nodes), out_ids, out_cap, assoc, out_count,
      keep_marks, ids, count, num_nodes);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_relabel(const int64_t* ids, int32_t count, const int32_t* count_dev,
                    const int64_t* assoc, int64_t* out, void* stream) {
  TGN_REQUIRE(count >= 0, "relabel: negative count");
  if (count == 0) return TGN_OK;
  TGN_REQUIRE(ids && assoc && out, "relabel: NULL pointer");
  DevCount c{count_dev, count};
  launch_k(relabel_kernel, dim3(stride_grid(count, 256)), dim3(256), 0, (cudaStream_t)stream, ids, c, assoc, out);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
