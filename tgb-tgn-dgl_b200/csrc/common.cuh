// Shared device/host helpers for the sm_100a TGN hot-path kernels.
// Everything in csrc/ is written for B200 (sm_100a) only; there is no CPU or
// multi-arch fallback.  Status/err conventions follow include/tgn_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define TGN_OK 0
#define TGN_EINVAL (-1)
#define TGN_ECUDA (-2)

namespace tgn {

// thread-local last-error text, read through tgn_last_error()
char* err_buf();
int set_err(int code, const char* fmt, ...);

#define TGN_REQUIRE(cond, ...)                         \
  do {                                                 \
    if (!(cond)) return ::tgn::set_err(TGN_EINVAL, __VA_ARGS__); \
  } while (0)

#define TGN_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess)                                                   \
      return ::tgn::set_err(TGN_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, \
                            #expr, cudaGetErrorString(e__));                  \
  } while (0)

#define TGN_LAUNCH_CHECK()                                                    \
  do {                                                                        \
    cudaError_t e__ = cudaPeekAtLastError();                                  \
    if (e__ != cudaSuccess)                                                   \
      return ::tgn::set_err(TGN_ECUDA, "%s:%d launch -> %s", __FILE__,       \
                            __LINE__, cudaGetErrorString(e__));               \
  } while (0)

// Device-side error word: one int32 in mapped pinned host memory (allocated on first use, portable
// across devices).  Kernels that detect a contract violation they cannot report through a return
// code (a log overflow or an out-of-range event id inside a captured graph) OR a bit into it and
// carry on without touching memory out of bounds; the host reads it with tgn_device_errors() --
// a plain memory read, no CUDA call.
#define TGN_DEVERR_LOG_OVERFLOW 1   /* message-store log full: events were dropped */
#define TGN_DEVERR_EVENT_RANGE 2    /* an e_id outside the resident event arrays was dereferenced */
#define TGN_DEVERR_SORT_CAP 4       /* a batch exceeded a kernel's in-shared-memory sort capacity */
#define TGN_DEVERR_OWNER_CAP 8      /* one rank's share of a step's rows exceeded the bound its buffers were sized for */
int32_t* dev_err_word();
__device__ __forceinline__ void flag_dev_err(int32_t* w, int bit) {
  if (w) atomicOr_system(reinterpret_cast<int*>(w), bit);
}

// Opt-in to > 48 KB of dynamic shared memory.  cudaFuncSetAttribute is per DEVICE, so the "already
// done" cache is a bit per device ordinal (a process that drives several GPUs sets it on each).
template <typename F>
static inline cudaError_t smem_optin(F func, int bytes, unsigned long long& done_mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && ((done_mask >> dev) & 1ull)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done_mask |= 1ull << dev;
  return e;
}

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Grid for a grid-stride elementwise kernel: whole multiples of the SM count.
static inline int stride_grid(long long work_items, int block, int max_waves = 8) {
  long long blocks = (work_items + block - 1) / block;
  if (blocks < 1) blocks = 1;
  long long cap = (long long)kNumSMs * max_waves;
  if (blocks > cap) blocks = cap;
  return (int)blocks;
}

// ---------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the library starts with
// pdl_wait() -- it blocks until the preceding kernel of the stream has completed and its
// writes are visible -- followed by pdl_launch(), which lets the NEXT kernel of the stream
// become resident early (its CTAs then sit in their own pdl_wait).  Launching through
// launch_k() sets the programmatic-stream-serialization attribute, so the launch latency and
// the prologue of kernel N+1 overlap the tail of kernel N (also inside captured CUDA graphs,
// where the edge becomes a programmatic dependency).  tgn_set_pdl(0) turns the attribute off;
// the two instructions are no-ops for a kernel launched without it.
// ---------------------------------------------------------------------------
int& pdl_flag();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                            cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_flag() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit global accesses (read-once data: keep it out of L1)
__device__ __forceinline__ int4 ld_stream_v4(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_v4(void* p, int4 v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// A device-side element count: either a host constant or a value another
// kernel left in device memory (so a captured CUDA graph can be replayed with
// data-dependent sizes and no host sync).
struct DevCount {
  const int32_t* dev;  // nullable
  int32_t host;        // used when dev == nullptr; otherwise the upper bound
  __device__ __forceinline__ int get() const {
    if (dev == nullptr) return host;
    int v = *dev;
    return v < host ? v : host;
  }
};

// ---------------------------------------------------------------------------
// Single-pass chained scan across CTAs ("decoupled look-back").
// ws[0] is the tile ticket, ws[1..] one status word per tile:
//   bits 63..62: 0 = empty, 1 = tile aggregate, 2 = inclusive prefix
//   bits 61..0 : value
// The workspace must be zero when the kernel starts (the host wrapper issues a
// cudaMemsetAsync in front of the launch).  Tiles are numbered by ticket so a
// tile only ever waits for tiles that are already running.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int lookback_take_tile(unsigned long long* ws) {
  __shared__ int s_tile;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ws, 1ull);
  __syncthreads();
  return s_tile;
}

// Called by ONE thread of the CTA.  Returns the exclusive prefix of this tile.
__device__ __forceinline__ long long lookback_prefix(unsigned long long* ws, int tile,
                                                     long long tile_sum) {
  volatile unsigned long long* st = ws + 1;
  const unsigned long long kAgg = 1ull << 62, kInc = 2ull << 62, kMask = (1ull << 62) - 1;
  if (tile == 0) {
    st[0] = kInc | (unsigned long long)tile_sum;
    return 0;
  }
  st[tile] = kAgg | (unsigned long long)tile_sum;
  long long prefix = 0;
  for (int p = tile - 1; p >= 0; --p) {
    unsigned long long v;
    do {
      v = st[p];
    } while ((v >> 62) == 0);
    prefix += (long long)(v & kMask);
    if ((v >> 62) == 2) break;
  }
  st[tile] = kInc | (unsigned long long)(prefix + tile_sum);
  return prefix;
}

// Full-range sine / cosine for the TimeEncoder (arguments w*dt+b reach 1e6 rad on the wiki shape).
// CUDA's sincosf switches to a Payne-Hanek reduction through local memory beyond |x| = 105615; a
// double-precision reduction by 2*pi of the SAME fp32 argument is two FP64 instructions and keeps
// the result within 2e-7 of the exactly reduced one (the parity bar is 1e-5).
__device__ __forceinline__ float reduce_2pi(float x) {
  if (fabsf(x) > 105615.0f) {
    const double a = (double)x;
    const double k = rint(a * 0.15915494309189535);
    x = (float)fma(-k, 6.283185307179586, a);
  }
  return x;
}
__device__ __forceinline__ void sincos_fr(float x, float* s, float* c) { sincosf(reduce_2pi(x), s, c); }
__device__ __forceinline__ float cos_fr(float x) { return cosf(reduce_2pi(x)); }
__device__ __forceinline__ float sin_fr(float x) { return sinf(reduce_2pi(x)); }

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Warp-wide variant: called by ALL 32 lanes of ONE warp of the CTA (tile_sum
// uniform across the lanes).  Every round inspects 32 predecessors at once, so
// a tile that becomes ready together with hundreds of others resolves its
// prefix in a few L2 round trips instead of one per predecessor.
__device__ __forceinline__ long long lookback_prefix_warp(unsigned long long* ws, int tile,
                                                          long long tile_sum) {
  volatile unsigned long long* st = ws + 1;
  const unsigned long long kAgg = 1ull << 62, kInc = 2ull << 62, kMask = (1ull << 62) - 1;
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) st[0] = kInc | (unsigned long long)tile_sum;
    return 0;
  }
  if (lane == 0) st[tile] = kAgg | (unsigned long long)tile_sum;
  long long prefix = 0;
  for (int hi = tile - 1; hi >= 0; hi -= 32) {
    const int p = hi - lane;
    unsigned long long v = kInc;                       // lanes before tile 0 read as "inclusive 0"
    if (p >= 0) {
      do {
        v = st[p];
      } while ((v >> 62) == 0);
    }
    const unsigned inc = __ballot_sync(0xffffffffu, (v >> 62) == 2);
    const int first = inc ? __ffs(inc) - 1 : 32;       // nearest predecessor with an inclusive prefix
    long long x = lane <= first ? (long long)(v & kMask) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    prefix += x;
    if (inc) break;
  }
  if (lane == 0) st[tile] = kInc | (unsigned long long)(prefix + tile_sum);
  return prefix;
}

// Block-wide variant: called by ALL threads of the CTA (blockDim.x a multiple of
// 32, tile_sum uniform); every thread gets the exclusive prefix.  One round
// inspects blockDim.x predecessors, so even when every resident tile posts its
// aggregate at the same moment the prefix resolves in one or two L2 round trips.
__device__ __forceinline__ long long lookback_prefix_block(unsigned long long* ws, int tile,
                                                           long long tile_sum) {
  __shared__ long long s_lb_sum[32];
  __shared__ int s_lb_first[32];
  volatile unsigned long long* st = ws + 1;
  const unsigned long long kAgg = 1ull << 62, kInc = 2ull << 62, kMask = (1ull << 62) - 1;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  if (tile == 0) {
    if (tid == 0) st[0] = kInc | (unsigned long long)tile_sum;
    return 0;
  }
  if (tid == 0) st[tile] = kAgg | (unsigned long long)tile_sum;
  long long prefix = 0;
  for (int hi = tile - 1; hi >= 0; hi -= (int)blockDim.x) {
    const int p = hi - tid;
    unsigned long long v = kInc;                       // slots before tile 0 read as "inclusive 0"
    if (p >= 0) {
      do {
        v = st[p];
      } while ((v >> 62) == 0);
    }
    const unsigned inc = __ballot_sync(0xffffffffu, (v >> 62) == 2);
    const int first = inc ? __ffs(inc) - 1 : 32;       // nearest inclusive predecessor within the warp's window
    long long x = lane <= first ? (long long)(v & kMask) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) {
      s_lb_sum[wid] = x;
      s_lb_first[wid] = first;
    }
    __syncthreads();
    bool done = false;
    for (int w = 0; w < nw && !done; ++w) {
      prefix += s_lb_sum[w];
      done = s_lb_first[w] < 32;
    }
    __syncthreads();
    if (done) break;
  }
  if (tid == 0) st[tile] = kInc | (unsigned long long)(prefix + tile_sum);
  return prefix;
}

// Philox-4x32-10 counter-based RNG (Salmon et al. 2011) -- used for uniform
// neighbour draws and attention dropout so results do not depend on the
// launch geometry.
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32);
    uint32_t c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

}  // namespace tgn
