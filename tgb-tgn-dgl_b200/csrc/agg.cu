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
// sum / max(count, 1); empty segments stay 0.
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
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    tmax[s] = 0ull;
    arg[s] = sentinel;
  }
}

template <bool kFloat>
__global__ void agg_last_max_kernel(const int64_t* __restrict__ index, const void* __restrict__ t,
                                    int M, int S, unsigned long long* __restrict__ tmax) {
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

__global__ void agg_mean_zero_kernel(float* out, float* cnt, int S, int W) {
  const long long total = (long long)S * W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x)
    out[e] = 0.f;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x)
    cnt[s] = 0.f;
}
__global__ void agg_mean_add_kernel(const float* __restrict__ msg,
                                    const int64_t* __restrict__ index, int M, int S, int W,
                                    float* __restrict__ out, float* __restrict__ cnt) {
  const long long total = (long long)M * W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / W), c = (int)(e - (long long)i * W);
    const int64_t s = index[i];
    if (s < 0 || s >= S) continue;
    atomicAdd(&out[s * W + c], msg[e]);
    if (c == 0) atomicAdd(&cnt[s], 1.f);
  }
}
__global__ void agg_mean_div_kernel(float* __restrict__ out, const float* __restrict__ cnt, int S,
                                    int W) {
  const long long total = (long long)S * W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(e / W);
    out[e] = out[e] / fmaxf(cnt[s], 1.f);
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
  agg_last_init_kernel<<<stride_grid(dim_size, 256), 256, 0, s>>>(tmax, arg, dim_size,
                                                                  (unsigned long long)num_msgs);
  TGN_LAUNCH_CHECK();
  if (num_msgs > 0) {
    const int g = stride_grid(num_msgs, 256);
    if (t_is_float) {
      agg_last_max_kernel<true><<<g, 256, 0, s>>>(index, t, num_msgs, dim_size, tmax);
      agg_last_arg_kernel<true><<<g, 256, 0, s>>>(index, t, num_msgs, dim_size, tmax, arg);
    } else {
      agg_last_max_kernel<false><<<g, 256, 0, s>>>(index, t, num_msgs, dim_size, tmax);
      agg_last_arg_kernel<false><<<g, 256, 0, s>>>(index, t, num_msgs, dim_size, tmax, arg);
    }
    TGN_LAUNCH_CHECK();
  }
  const bool vec = (width % 4 == 0) && (((uintptr_t)msg | (uintptr_t)out) % 16 == 0);
  if (vec) {
    const long long total = (long long)dim_size * (width / 4);
    agg_last_gather4_kernel<<<stride_grid(total, 256), 256, 0, s>>>(
        (const float4*)msg, arg, num_msgs, dim_size, width / 4, (float4*)out, argmax);
  } else {
    const long long total = (long long)dim_size * width;
    agg_last_gather_kernel<<<stride_grid(total, 256), 256, 0, s>>>(msg, arg, num_msgs, dim_size,
                                                                   width, out, argmax);
  }
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_agg_mean(const float* msg, const int64_t* index, int32_t num_msgs,
                     int32_t dim_size, int32_t width, float* out, void* ws, void* stream) {
  TGN_REQUIRE(num_msgs >= 0 && dim_size >= 0 && width >= 1, "agg_mean: bad sizes");
  if (dim_size == 0) return TGN_OK;
  TGN_REQUIRE(out && ws, "agg_mean: NULL output/workspace");
  TGN_REQUIRE(num_msgs == 0 || (msg && index), "agg_mean: NULL input");
  cudaStream_t s = (cudaStream_t)stream;
  float* cnt = (float*)ws;
  const long long tot_out = (long long)dim_size * width;
  agg_mean_zero_kernel<<<stride_grid(tot_out, 256), 256, 0, s>>>(out, cnt, dim_size, width);
  TGN_LAUNCH_CHECK();
  if (num_msgs > 0) {
    const long long tot_in = (long long)num_msgs * width;
    agg_mean_add_kernel<<<stride_grid(tot_in, 256), 256, 0, s>>>(msg, index, num_msgs, dim_size,
                                                                 width, out, cnt);
    TGN_LAUNCH_CHECK();
    agg_mean_div_kernel<<<stride_grid(tot_out, 256), 256, 0, s>>>(out, cnt, dim_size, width);
    TGN_LAUNCH_CHECK();
  }
  return TGN_OK;
}

}  // extern "C"
