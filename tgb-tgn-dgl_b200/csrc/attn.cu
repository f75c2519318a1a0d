// Fused temporal attention: GraphAttentionEmbedding.forward (reference
// modules/emb_module.py:25-29) on top of torch_geometric's
// TransformerConv(in, C, heads=H, concat=True, beta=False, root_weight=True,
// edge_dim=time_dim+raw_dim) -- the restated upstream algorithm is in
// SURVEY.md B5.
//
// Node projections q|k|v|skip come from one GEMM over the node rows (proj
// [Nb, 4*H*C]).  This kernel does the per-edge part in one pass:
//   rel_t -> cos time encoding -> edge_attr = [t_enc, msg] (never written to HBM)
//   ee = W_edge * edge_attr, score = <q_i, k_j + ee>/sqrt(C), online softmax over
//   the centre's edges, out_i = sum alpha (v_j + ee) + skip_i.
// One warp per centre; lanes own channels (lane, lane+32, ...).  W_edge is held
// transposed in shared memory ([edge_dim][H*C], conflict-free for lanes over
// channels) and edges are processed four at a time so every shared-memory read
// of W_edge feeds four FMAs.
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

constexpr int kAttnWarps = 8;
constexpr int kEB = 4;       // edges per batch
constexpr int kMaxCH = 8;    // channels per lane -> H*C <= 256
constexpr int kMaxHeads = 8;

struct AttnArgs {
  const float* proj;
  const void* lu;  // last_update of local rows: int64 or float
  int lu_is_float;
  const int64_t* nbr;
  const void* t_edge;  // float or int64
  int t_is_float;
  const float* msg;
  const int64_t* msg_rows;
  const int32_t* row_ptr;
  const int32_t* edge_perm;
  const int64_t* centre_ids;
  DevCount centres;
  int H, C, De, Dt;
  const float* w_edge;  // [H*C, Dt+De]
  const float* time_w;
  const float* time_b;
  float dropout_p;
  uint64_t seed;
  const int64_t* seed_dev;
  float* out;
  float* alpha_out;  // [E,H]  (softmax weights before dropout)
  float* ee_out;     // [E,H*C]
};

// e = payload row (msg_rows[edge] when given, else the edge position)
__device__ __forceinline__ float attn_rel_t(const AttnArgs& a, int64_t j, long long e) {
  if (!a.lu_is_float && !a.t_is_float) {
    const int64_t d = reinterpret_cast<const int64_t*>(a.lu)[j] -
                      reinterpret_cast<const int64_t*>(a.t_edge)[e];
    return (float)d;
  }
  const float l = a.lu_is_float ? reinterpret_cast<const float*>(a.lu)[j]
                                : (float)reinterpret_cast<const int64_t*>(a.lu)[j];
  const float t = a.t_is_float ? reinterpret_cast<const float*>(a.t_edge)[e]
                               : (float)reinterpret_cast<const int64_t*>(a.t_edge)[e];
  return l - t;
}

__global__ void __launch_bounds__(kAttnWarps * 32, 1) attn_fwd_kernel(AttnArgs a) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float smem[];
  const int HC = a.H * a.C;
  const int Din = a.Dt + a.De;
  float* s_wT = smem;                        // [Din][HC]
  float* s_ea = smem + (((size_t)Din * HC + 3) & ~(size_t)3);  // [warps][Din][kEB]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < Din * HC; i += blockDim.x) {
    const int c = i / Din, d = i - c * Din;  // read row-major [HC][Din] coalesced
    s_wT[d * HC + c] = a.w_edge[i];
  }
  __syncthreads();
  float* ea = s_ea + (size_t)wid * Din * kEB;
  const int nC = a.centres.get();
  const float inv_sqrt_c = rsqrtf((float)a.C);
  const int CH = (HC + 31) >> 5;
  const float keep = 1.f - a.dropout_p;
  Philox rng(a.seed + (a.seed_dev ? (uint64_t)*a.seed_dev : 0ull));

  for (int ci = blockIdx.x * kAttnWarps + wid; ci < nC; ci += gridDim.x * kAttnWarps) {
    const int64_t row = a.centre_ids ? a.centre_ids[ci] : ci;
    const float* pr = a.proj + row * 4 * HC;
    float q[kMaxCH], acc[kMaxCH];
    int head[kMaxCH];
#pragma unroll
    for (int i = 0; i < kMaxCH; ++i) {
      const int c = lane + 32 * i;
      q[i] = (i < CH && c < HC) ? pr[c] : 0.f;
      acc[i] = 0.f;
      head[i] = (i < CH && c < HC) ? c / a.C : -1;
    }
    float mrun[kMaxHeads], lrun[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) {
      mrun[h] = -INFINITY;
      lrun[h] = 0.f;
    }
    const int e0 = a.row_ptr[ci], e1 = a.row_ptr[ci + 1];
    for (int eb = e0; eb < e1; eb += kEB) {
      const int ne = min(kEB, e1 - eb);
      int eid[kEB];
      int64_t jn[kEB];
#pragma unroll
      for (int u = 0; u < kEB; ++u) {
        eid[u] = u < ne ? (a.edge_perm ? a.edge_perm[eb + u] : eb + u) : -1;
        jn[u] = u < ne ? a.nbr[eid[u]] : 0;
      }
      // edge_attr for the batch -> shared, interleaved [d][u]
      __syncwarp();
#pragma unroll
      for (int u = 0; u < kEB; ++u) {
        if (u < ne) {
          const long long mr = a.msg_rows ? a.msg_rows[eid[u]] : eid[u];
          const float rt = attn_rel_t(a, jn[u], mr);
          for (int d = lane; d < a.Dt; d += 32)
            ea[d * kEB + u] = cos_fr(__fmaf_rn(rt, a.time_w[d], a.time_b[d]));
          const float* mp = a.msg + mr * a.De;
          for (int d = lane; d < a.De; d += 32) ea[(a.Dt + d) * kEB + u] = mp[d];
        } else {
          for (int d = lane; d < Din; d += 32) ea[d * kEB + u] = 0.f;
        }
      }
      __syncwarp();
      // ee[u][i] = sum_d W[c_i, d] * ea[d][u]
      float ee[kEB][kMaxCH];
#pragma unroll
      for (int u = 0; u < kEB; ++u)
#pragma unroll
        for (int i = 0; i < kMaxCH; ++i) ee[u][i] = 0.f;
      for (int d = 0; d < Din; ++d) {
        const float4 ev = *reinterpret_cast<const float4*>(ea + d * kEB);
        const float evs[kEB] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
        for (int i = 0; i < kMaxCH; ++i) {
          if (i < CH) {
            const int c = lane + 32 * i;
            const float w = c < HC ? s_wT[d * HC + c] : 0.f;
#pragma unroll
            for (int u = 0; u < kEB; ++u) ee[u][i] = fmaf(w, evs[u], ee[u][i]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kEB; ++u) {
        if (u >= ne) continue;
        const float* pj = a.proj + jn[u] * 4 * HC;
        float kv[kMaxCH], vv[kMaxCH];
        float part[kMaxHeads];
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h) part[h] = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxCH; ++i) {
          const int c = lane + 32 * i;
          const bool ok = i < CH && c < HC;
          kv[i] = ok ? pj[HC + c] + ee[u][i] : 0.f;
          vv[i] = ok ? pj[2 * HC + c] + ee[u][i] : 0.f;
          if (ok && a.ee_out) a.ee_out[(long long)eid[u] * HC + c] = ee[u][i];
#pragma unroll
          for (int h = 0; h < kMaxHeads; ++h)
            if (head[i] == h) part[h] = fmaf(q[i], kv[i], part[h]);
        }
        float pw[kMaxHeads], sc[kMaxHeads];
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h) {
          if (h < a.H) {
            const float s = warp_sum(part[h]) * inv_sqrt_c;
            const float mnew = fmaxf(mrun[h], s);
            sc[h] = expf(mrun[h] - mnew);  // 0 on the first edge (mrun = -inf)
            float p = expf(s - mnew);
            lrun[h] = lrun[h] * sc[h] + p;
            mrun[h] = mnew;
            if (a.alpha_out && lane == 0) a.alpha_out[(long long)eid[u] * a.H + h] = s;
            if (a.dropout_p > 0.f) {
              const uint4 r = rng((uint64_t)eid[u], (uint64_t)h);
              const float uni = (float)(r.x >> 8) * (1.0f / 16777216.0f);
              p = uni < keep ? p / keep : 0.f;
            }
            pw[h] = p;
          } else {
            pw[h] = 0.f;
            sc[h] = 1.f;
          }
        }
#pragma unroll
        for (int i = 0; i < kMaxCH; ++i) {
#pragma unroll
          for (int h = 0; h < kMaxHeads; ++h)
            if (head[i] == h) acc[i] = acc[i] * sc[h] + pw[h] * vv[i];
        }
      }
    }
    // normalise, add skip, write
    float* po = a.out + row * HC;
#pragma unroll
    for (int i = 0; i < kMaxCH; ++i) {
      const int c = lane + 32 * i;
      if (i < CH && c < HC) {
        float l = 1.f;
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h)
          if (head[i] == h) l = lrun[h];
        const float o = (e1 > e0) ? acc[i] / l : 0.f;
        po[c] = o + pr[3 * HC + c];
      }
    }
    // raw scores -> softmax weights (kept for the backward pass)
    if (a.alpha_out) {
      __syncwarp();
      for (int idx = lane; idx < (e1 - e0) * a.H; idx += 32) {
        const int u = idx / a.H, h = idx - u * a.H;
        const int e = a.edge_perm ? a.edge_perm[e0 + u] : e0 + u;
        float m = 0.f, l = 1.f;
#pragma unroll
        for (int hh = 0; hh < kMaxHeads; ++hh)
          if (hh == h) {
            m = mrun[hh];
            l = lrun[hh];
          }
        float* pa = a.alpha_out + (long long)e * a.H + h;
        *pa = expf(*pa - m) / l;
      }
    }
  }
}

__global__ void attn_fill_skip_kernel(const float* __restrict__ proj, DevCount rows, int HC,
                                      float* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  const int n = rows.get();
  const long long total = (long long)n * HC;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / HC), c = (int)(e - (long long)r * HC);
    out[e] = proj[(long long)r * 4 * HC + 3 * HC + c];
  }
}

}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_attn_fwd(const float* proj, const void* last_update_local, int32_t lu_is_float,
                     const int64_t* nbr_local, const void* t_edge, int32_t t_is_float,
                     const float* msg, const int64_t* msg_rows, const int32_t* row_ptr,
                     const int32_t* edge_perm, const int64_t* centre_ids, int32_t num_centres,
                     const int32_t* num_centres_dev, int32_t heads, int32_t head_dim,
                     int32_t raw_dim, int32_t time_dim, const float* w_edge, const float* time_w,
                     const float* time_b, float dropout_p, uint64_t seed, const int64_t* seed_dev,
                     float* out, float* alpha_out, float* ee_out, void* stream) {
  TGN_REQUIRE(num_centres >= 0 && heads >= 1 && heads <= kMaxHeads && head_dim >= 1,
              "attn_fwd: bad sizes (heads must be <= %d)", kMaxHeads);
  TGN_REQUIRE(heads * head_dim <= 32 * kMaxCH, "attn_fwd: heads*head_dim must be <= %d",
              32 * kMaxCH);
  TGN_REQUIRE(raw_dim >= 0 && time_dim >= 0 && raw_dim + time_dim >= 1, "attn_fwd: bad edge dims");
  TGN_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "attn_fwd: dropout_p must be in [0,1)");
  if (num_centres == 0) return TGN_OK;
  TGN_REQUIRE(proj && last_update_local && nbr_local && t_edge && (msg || raw_dim == 0) &&
                  row_ptr && w_edge && (time_dim == 0 || (time_w && time_b)) && out,
              "attn_fwd: NULL pointer");
  const int HC = heads * head_dim, Din = raw_dim + time_dim;
  const size_t smem =
      ((((size_t)Din * HC + 3) & ~(size_t)3) + (size_t)kAttnWarps * Din * kEB) * sizeof(float);
  TGN_REQUIRE(smem <= 227 * 1024, "attn_fwd: W_edge (%d x %d) does not fit shared memory", HC, Din);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    TGN_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    attr_smem = smem;
  }
  AttnArgs a;
  a.proj = proj; a.lu = last_update_local; a.lu_is_float = lu_is_float; a.nbr = nbr_local;
  a.t_edge = t_edge; a.t_is_float = t_is_float; a.msg = msg; a.msg_rows = msg_rows;
  a.row_ptr = row_ptr; a.edge_perm = edge_perm; a.centre_ids = centre_ids;
  a.centres = DevCount{num_centres_dev, num_centres};
  a.H = heads; a.C = head_dim; a.De = raw_dim; a.Dt = time_dim; a.w_edge = w_edge;
  a.time_w = time_w; a.time_b = time_b; a.dropout_p = dropout_p; a.seed = seed; a.seed_dev = seed_dev; a.out = out;
  a.alpha_out = alpha_out; a.ee_out = ee_out;
  int grid = ceil_div(num_centres, kAttnWarps);
  if (grid > kNumSMs) grid = kNumSMs;
  launch_k(attn_fwd_kernel, dim3(grid), dim3(kAttnWarps * 32), smem, (cudaStream_t)stream, a);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_attn_fill_skip(const float* proj, int32_t num_rows, const int32_t* num_rows_dev,
                           int32_t hc, float* out, void* stream) {
  TGN_REQUIRE(num_rows >= 0 && hc >= 1, "attn_fill_skip: bad sizes");
  if (num_rows == 0) return TGN_OK;
  TGN_REQUIRE(proj && out, "attn_fill_skip: NULL pointer");
  DevCount c{num_rows_dev, num_rows};
  launch_k(attn_fill_skip_kernel, dim3(stride_grid((long long)num_rows * hc, 256)), dim3(256), 0, (cudaStream_t)stream, proj, c, hc, out);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
