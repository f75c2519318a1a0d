// Message aggregators on materialised messages (reference modules/msg_agg.py).
//
// LastAggregator.forward (msg_agg.py:15-21) = torch_scatter.scatter_max(t, index)
// argmax + row gather.  The reference's CPU tie rule is "first maximal element
// wins" (strict > in a sequential scan); the CUDA scatter_max upstream is
// nondeterministic among ties.  Here the rule is fixed to first-wins and made
// deterministic with two idempotent atomic passes:
//   pass 1  tmax[s]  = max over the segment of enc(t)      (atomicMax)
//   pass 2  arg[s]   = min i with enc(t[i]) == tmax[s]     (atomicMin)
// Consecutive equal indices inside a warp (the node-id runs the memory module
// produces) are combined with a segmented shuffle scan first, so a run issues a
// single atomic.
// MeanAggregator.forward (msg_agg.py:24-26) = scatter(..., reduce="mean"):
// sum / max(count, 1); empty segments stay 0.  No float atomics: message ids are counting-sorted
// by segment (integer atomics + a chained scan), then one warp per segment streams its rows
// with 128-bit loads, so the message matrix is read exactly once and the sum is reproducible.
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

__device__ __forceinline__ unsigned long long enc_i64(int64_t v) {
  return (unsigned long long)v ^ (1ull << 63);
}
__device__ __forceinline__ unsigned long long enc_f32(float f) {
  uint32_t u = __float_as_uint(f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (unsigned long long)u;
}
template <bool kFloat>
__device__ __forceinline__ unsigned long long load_enc(const void* t, int i) {
  if (kFloat) return enc_f32(reinterpret_cast<const float*>(t)[i]);
  return enc_i64(reinterpret_cast<const int64_t*>(t)[i]);
}

__global__ void agg_last_init_kernel(unsigned long long* tmax, unsigned long long* arg, int S,
                                     unsigned long long sentinel) {
  pdl_wait();
  pdl_launch();
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    tmax[s] = 0ull;
    arg[s] = sentinel;
  }
}

template <bool kFloat>
__global__ void agg_last_max_kernel(const int64_t* __restrict__ index, const void* __restrict__ t,
                                    int M, int S, unsigned long long* __restrict__ tmax) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31;
  const int total = (M + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int64_t idx = -1;
    unsigned long long v = 0;
    if (i < M) {
      idx = index[i];
      v = load_enc<kFloat>(t, i);
    }
    // segmented inclusive max over runs of equal idx
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t oi = __shfl_up_sync(0xffffffffu, idx, o);
      unsigned long long ov = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o && oi == idx && ov > v) v = ov;
    }
    int64_t nxt = __shfl_down_sync(0xffffffffu, idx, 1);
    bool tail = (lane == 31) || (nxt != idx);
    if (tail && idx >= 0 && idx < S) atomicMax(&tmax[idx], v);
  }
}

template <bool kFloat>
__global__ void agg_last_arg_kernel(const int64_t* __restrict__ index, const void* __restrict__ t,
                                    int M, int S, const unsigned long long* __restrict__ tmax,
                                    unsigned long long* __restrict__ arg) {
  pdl_wait();
  pdl_launch();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    int64_t idx = index[i];
    if (idx < 0 || idx >= S) continue;
    if (load_enc<kFloat>(t, i) == tmax[idx]) atomicMin(&arg[idx], (unsigned long long)i);
  }
}

__global__ void agg_last_gather_kernel(const float* __restrict__ msg,
                                       const unsigned long long* __restrict__ arg, int M, int S,
                                       int W, float* __restrict__ out,
                                       int64_t* __restrict__ argmax) {
  pdl_wait();
  pdl_launch();
  const long long total = (long long)S * W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(e / W), c = (int)(e - (long long)s * W);
    const unsigned long long a = arg[s];
    out[e] = a < (unsigned long long)M ? msg[a * W + c] : 0.f;
    if (c == 0 && argmax) argmax[s] = (int64_t)a;
  }
}

__global__ void agg_last_gather4_kernel(const float4* __restrict__ msg,
                                        const unsigned long long* __restrict__ arg, int M, int S,
                                        int W4, float4* __restrict__ out,
                                        int64_t* __restrict__ argmax) {
  pdl_wait();
  pdl_launch();
  const long long total = (long long)S * W4;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(e / W4), c = (int)(e - (long long)s * W4);
    const unsigned long long a = arg[s];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a < (unsigned long long)M) v = msg[a * W4 + c];
    out[e] = v;
    if (c == 0 && argmax) argmax[s] = (int64_t)a;
  }
}

// ---- segmented mean without float atomics: counting sort of the message ids by segment
// (integer atomics only), then one warp per segment streams its rows with 128-bit loads.
//   ws = int32 off[S+1] | int32 cur[S] | int32 perm[M]
__global__ void agg_mean_count_kernel(const int64_t* __restrict__ index, int M, int S,
                                      int32_t* __restrict__ cnt) {
  pdl_wait();
  pdl_launch();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    const int64_t s = index[i];
    if (s >= 0 && s < S) atomicAdd(&cnt[s], 1);
  }
}

constexpr int kScanTile = 2048;  // 256 threads x 8 segments

// exclusive scan of cnt[0..S) in place (cnt becomes off), off[S] = total, cur[s] = off[s]
__global__ void __launch_bounds__(256)
    agg_mean_scan_kernel(int32_t* __restrict__ off, int32_t* __restrict__ cur, int S,
                         unsigned long long* __restrict__ ws) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_warp[8];
  __shared__ long long s_prefix;
  const int tile = lookback_take_tile(ws);
  const int ntiles = (S + kScanTile - 1) / kScanTile;
  if (tile >= ntiles) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int base = tile * kScanTile + tid * 8;
  int v[8], sum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = base + j < S ? off[base + j] : 0;
    sum += v[j];
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  int wbase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if (w < wid) wbase += s_warp[w];
    tot += s_warp[w];
  }
  if (tid == 0) {
    const long long pre = lookback_prefix(ws, tile, tot);
    s_prefix = pre;
    if (tile == ntiles - 1) off[S] = (int32_t)(pre + tot);
  }
  __syncthreads();
  int run = (int)s_prefix + wbase + incl - sum;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (base + j < S) {
      off[base + j] = run;
      cur[base + j] = run;
    }
    run += v[j];
  }
}

__global__ void agg_mean_fill_kernel(const int64_t* __restrict__ index, int M, int S,
                                     int32_t* __restrict__ cur, int32_t* __restrict__ perm) {
  pdl_wait();
  pdl_launch();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    const int64_t s = index[i];
    if (s >= 0 && s < S) perm[atomicAdd(&cur[s], 1)] = i;
  }
}

// One warp per segment.  Segments of <= 32 messages are put back into message order with a
// warp bitonic sort first, so the summation order (and the result) is reproducible.
template <bool kVec>
__global__ void __launch_bounds__(256)
    agg_mean_reduce_kernel(const float* __restrict__ msg, const int32_t* __restrict__ off,
                           const int32_t* __restrict__ perm, int S, int W,
                           float* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < S; s += gridDim.x * wpb) {
    const int b = off[s], n = off[s + 1] - b;
    const float inv = 1.f / (float)(n > 1 ? n : 1);
    int mine = (n <= 32 && lane < n) ? perm[b + lane] : 0x7fffffff;
    if (n > 1 && n <= 32) {
#pragma unroll
      for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
          const int other = __shfl_xor_sync(0xffffffffu, mine, j);
          const bool up = ((lane & k) == 0) == ((lane & j) == 0);
          mine = up ? min(mine, other) : max(mine, other);
        }
    }
    if (kVec) {
      const int W4 = W >> 2;
      const float4* m4 = reinterpret_cast<const float4*>(msg);
      float4* o4 = reinterpret_cast<float4*>(out) + (long long)s * W4;
      for (int c0 = 0; c0 < W4; c0 += 128) {  // 4 float4 per lane per pass
        float4 acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < n; ++j) {
          const int i = n <= 32 ? __shfl_sync(0xffffffffu, mine, j) : perm[b + j];
          const float4* row = m4 + (long long)i * W4;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c0 + u * 32 + lane;
            if (c < W4) {
              const float4 v = row[c];
              acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * 32 + lane;
          if (c < W4) o4[c] = make_float4(acc[u].x * inv, acc[u].y * inv, acc[u].z * inv, acc[u].w * inv);
        }
      }
    } else {
      for (int c0 = 0; c0 < W; c0 += 128) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < n; ++j) {
          const int i = n <= 32 ? __shfl_sync(0xffffffffu, mine, j) : perm[b + j];
          const float* row = msg + (long long)i * W;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c0 + u * 32 + lane;
            if (c < W) acc[u] += row[c];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * 32 + lane;
          if (c < W) out[(long long)s * W + c] = acc[u] * inv;
        }
      }
    }
  }
}

}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_agg_last(const float* msg, const int64_t* index, const void* t, int32_t t_is_float,
                     int32_t num_msgs, int32_t dim_size, int32_t width, float* out,
                     int64_t* argmax, void* ws, void* stream) {
  TGN_REQUIRE(num_msgs >= 0 && dim_size >= 0 && width >= 1, "agg_last: bad sizes");
  if (dim_size == 0) return TGN_OK;
  TGN_REQUIRE(out && ws, "agg_last: NULL output/workspace");
  TGN_REQUIRE(num_msgs == 0 || (msg && index && t), "agg_last: NULL input");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* tmax = (unsigned long long*)ws;
  unsigned long long* arg = tmax + dim_size;
  launch_k(agg_last_init_kernel, dim3(stride_grid(dim_size, 256)), dim3(256), 0, s, tmax, arg, dim_size,
                                                                  (unsigned long long)num_msgs);
  TGN_LAUNCH_CHECK();
  if (num_msgs > 0) {
    const int g = stride_grid(num_msgs, 256);
    if (t_is_float) {
      launch_k(agg_last_max_kernel<true>, dim3(g), dim3(256), 0, s, index, t, num_msgs, dim_size, tmax);
      launch_k(agg_last_arg_kernel<true>, dim3(g), dim3(256), 0, s, index, t, num_msgs, dim_size, tmax, arg);
    } else {
      launch_k(agg_last_max_kernel<false>, dim3(g), dim3(256), 0, s, index, t, num_msgs, dim_size, tmax);
      launch_k(agg_last_arg_kernel<false>, dim3(g), dim3(256), 0, s, index, t, num_msgs, dim_size, tmax, arg);
    }
    TGN_LAUNCH_CHECK();
  }
  const bool vec = (width % 4 == 0) && (((uintptr_t)msg | (uintptr_t)out) % 16 == 0);
  if (vec) {
    const long long total = (long long)dim_size * (width / 4);
    launch_k(agg_last_gather4_kernel, dim3(stride_grid(total, 256)), dim3(256), 0, s, 
        (const float4*)msg, arg, num_msgs, dim_size, width / 4, (float4*)out, argmax);
  } else {
    const long long total = (long long)dim_size * width;
    launch_k(agg_last_gather_kernel, dim3(stride_grid(total, 256)), dim3(256), 0, s, msg, arg, num_msgs, dim_size,
                                                                   width, out, argmax);
  }
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int64_t tgn_agg_mean_ws_bytes(int32_t num_msgs, int32_t dim_size) {
  if (num_msgs < 0 || dim_size < 0) return 0;
  const int ntiles = (dim_size + kScanTile - 1) / kScanTile;
  // off[S+1] | cur[S] | perm[M] (int32), then the scan's ticket + tile states (8-byte aligned)
  int64_t words = 2ll * dim_size + 1 + num_msgs;
  words = (words + 1) & ~1ll;
  return words * 4 + (int64_t)(ntiles + 1) * 8;
}

int32_t tgn_agg_mean(const float* msg, const int64_t* index, int32_t num_msgs,
                     int32_t dim_size, int32_t width, float* out, void* ws, void* stream) {
  TGN_REQUIRE(num_msgs >= 0 && dim_size >= 0 && width >= 1, "agg_mean: bad sizes");
  if (dim_size == 0) return TGN_OK;
  TGN_REQUIRE(out && ws, "agg_mean: NULL output/workspace");
  TGN_REQUIRE(num_msgs == 0 || (msg && index), "agg_mean: NULL input");
  cudaStream_t s = (cudaStream_t)stream;
  const int S = dim_size, M = num_msgs;
  int32_t* off = (int32_t*)ws;
  int32_t* cur = off + S + 1;
  int32_t* perm = cur + S;
  int64_t words = (2ll * S + 1 + M + 1) & ~1ll;
  unsigned long long* scan_ws = (unsigned long long*)((int32_t*)ws + words);
  const int ntiles = (S + kScanTile - 1) / kScanTile;
  TGN_CUDA(cudaMemsetAsync(off, 0, (size_t)(S + 1) * 4, s));
  TGN_CUDA(cudaMemsetAsync(scan_ws, 0, (size_t)(ntiles + 1) * 8, s));
  if (M > 0) {
    launch_k(agg_mean_count_kernel, dim3(stride_grid(M, 256)), dim3(256), 0, s, index, M, S, off);
    TGN_LAUNCH_CHECK();
  }
  launch_k(agg_mean_scan_kernel, dim3(ntiles), dim3(256), 0, s, off, cur, S, scan_ws);
  TGN_LAUNCH_CHECK();
  if (M > 0) {
    launch_k(agg_mean_fill_kernel, dim3(stride_grid(M, 256)), dim3(256), 0, s, index, M, S, cur, perm);
    TGN_LAUNCH_CHECK();
  }
  const bool vec = (width % 4 == 0) && (((uintptr_t)msg | (uintptr_t)out) % 16 == 0);
  const int grid = stride_grid((long long)S * 32, 256, 16);
  if (vec) launch_k(agg_mean_reduce_kernel<true>, dim3(grid), dim3(256), 0, s, msg, off, perm, S, width, out);
  else launch_k(agg_mean_reduce_kernel<false>, dim3(grid), dim3(256), 0, s, msg, off, perm, S, width, out);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
