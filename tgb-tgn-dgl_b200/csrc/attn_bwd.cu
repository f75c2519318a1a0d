// Backward of the fused temporal attention (see attn.cu for the forward and
// the reference citations: modules/emb_module.py:25-29 + TransformerConv).
//
//   out_i = sum_e a~_e (v_j + ee_e) + skip_i,   a~ = dropout(alpha),
//   alpha = softmax_e(s_e),  s_e = <q_i, k_j + ee_e> / sqrt(C)
//
// One warp per centre, two passes over its edges (the second pass recomputes
// the cheap dot products instead of spilling per-edge state):
//   pass 1: dot_h = sum_e alpha_e * d alpha_e
//   pass 2: ds_e = alpha_e (d alpha_e - dot_h);  dq += ds (k+ee)/sqrt(C);
//           d(k_j) += ds q/sqrt(C);  d(v_j) += a~ d out;  d ee_e = both
// d_proj rows of neighbours receive atomic adds (a node can be the neighbour of
// many centres); d_proj must be initialised by tgn_attn_bwd_init, which also
// routes d_out into the skip block.
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

constexpr int kBwdWarps = 8;
constexpr int kBMaxCH = 8;
constexpr int kBMaxHeads = 8;

struct AttnBwdArgs {
  const float* proj;
  const int64_t* nbr;
  const int32_t* row_ptr;
  const int32_t* edge_perm;
  const int64_t* centre_ids;
  DevCount centres;
  int H, C;
  const float* alpha;  // [E,H]
  const float* ee;     // [E,HC]
  const float* d_out;  // [Nb,HC]
  float dropout_p;
  uint64_t seed;
  const int64_t* seed_dev;
  float* d_proj;  // [Nb,4HC]
  float* d_ee;    // [E,HC]
};

__global__ void __launch_bounds__(kBwdWarps * 32) attn_bwd_kernel(AttnBwdArgs a) {
  pdl_wait();
  pdl_launch();
  const int HC = a.H * a.C;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nC = a.centres.get();
  const float inv_sqrt_c = rsqrtf((float)a.C);
  const int CH = (HC + 31) >> 5;
  const float keep = 1.f - a.dropout_p;
  Philox rng(a.seed + (a.seed_dev ? (uint64_t)*a.seed_dev : 0ull));
  for (int ci = blockIdx.x * kBwdWarps + wid; ci < nC; ci += gridDim.x * kBwdWarps) {
    const int64_t row = a.centre_ids ? a.centre_ids[ci] : ci;
    const float* pr = a.proj + row * 4 * HC;
    const float* go = a.d_out + row * HC;
    float q[kBMaxCH], g[kBMaxCH], dq[kBMaxCH];
    int head[kBMaxCH];
#pragma unroll
    for (int i = 0; i < kBMaxCH; ++i) {
      const int c = lane + 32 * i;
      const bool ok = i < CH && c < HC;
      q[i] = ok ? pr[c] : 0.f;
      g[i] = ok ? go[c] : 0.f;
      dq[i] = 0.f;
      head[i] = ok ? c / a.C : -1;
    }
    const int e0 = a.row_ptr[ci], e1 = a.row_ptr[ci + 1];
    float dot[kBMaxHeads];
#pragma unroll
    for (int h = 0; h < kBMaxHeads; ++h) dot[h] = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
      for (int ep = e0; ep < e1; ++ep) {
        const int e = a.edge_perm ? a.edge_perm[ep] : ep;
        const int64_t j = a.nbr[e];
        const float* pj = a.proj + j * 4 * HC;
        const float* pe = a.ee + (long long)e * HC;
        float kk[kBMaxCH], vv[kBMaxCH];
        float part[kBMaxHeads];
#pragma unroll
        for (int h = 0; h < kBMaxHeads; ++h) part[h] = 0.f;
#pragma unroll
        for (int i = 0; i < kBMaxCH; ++i) {
          const int c = lane + 32 * i;
          const bool ok = i < CH && c < HC;
          const float eev = ok ? pe[c] : 0.f;
          kk[i] = ok ? pj[HC + c] + eev : 0.f;
          vv[i] = ok ? pj[2 * HC + c] + eev : 0.f;
#pragma unroll
          for (int h = 0; h < kBMaxHeads; ++h)
            if (head[i] == h) part[h] = fmaf(g[i], vv[i], part[h]);
        }
        float al[kBMaxHeads], dal[kBMaxHeads], mk[kBMaxHeads];
#pragma unroll
        for (int h = 0; h < kBMaxHeads; ++h) {
          al[h] = 0.f;
          dal[h] = 0.f;
          mk[h] = 1.f;
          if (h < a.H) {
            al[h] = a.alpha[(long long)e * a.H + h];
            if (a.dropout_p > 0.f) {
              const uint4 r = rng((uint64_t)e, (uint64_t)h);
              const float uni = (float)(r.x >> 8) * (1.0f / 16777216.0f);
              mk[h] = uni < keep ? 1.f / keep : 0.f;
            }
            dal[h] = warp_sum(part[h]) * mk[h];
          }
        }
        if (pass == 0) {
#pragma unroll
          for (int h = 0; h < kBMaxHeads; ++h) dot[h] = fmaf(al[h], dal[h], dot[h]);
          continue;
        }
        float ds[kBMaxHeads];
#pragma unroll
        for (int h = 0; h < kBMaxHeads; ++h) ds[h] = al[h] * (dal[h] - dot[h]) * inv_sqrt_c;
        float* dpj = a.d_proj + j * 4 * HC;
        float* dpe = a.d_ee + (long long)e * HC;
#pragma unroll
        for (int i = 0; i < kBMaxCH; ++i) {
          const int c = lane + 32 * i;
          if (!(i < CH && c < HC)) continue;
          float dsh = 0.f, at = 0.f;
#pragma unroll
          for (int h = 0; h < kBMaxHeads; ++h)
            if (head[i] == h) {
              dsh = ds[h];
              at = al[h] * mk[h];
            }
          dq[i] = fmaf(dsh, kk[i], dq[i]);
          const float dk = dsh * q[i];
          const float dv = at * g[i];
          atomicAdd(dpj + HC + c, dk);
          atomicAdd(dpj + 2 * HC + c, dv);
          dpe[c] = dk + dv;
        }
      }
    }
    float* dpr = a.d_proj + row * 4 * HC;
#pragma unroll
    for (int i = 0; i < kBMaxCH; ++i) {
      const int c = lane + 32 * i;
      if (i < CH && c < HC) dpr[c] = dq[i];  // centres are unique: plain store
    }
  }
}

// d_proj[r] = [0, 0, 0, d_out[r]] for r < rows, all-zero rows up to `bound`
__global__ void attn_bwd_init_kernel(const float* __restrict__ d_out, DevCount rows, int bound,
                                     int HC, float* __restrict__ d_proj) {
  pdl_wait();
  pdl_launch();
  const int n = rows.get();
  const long long total = (long long)bound * 4 * HC;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / (4 * HC)), c = (int)(e - (long long)r * 4 * HC);
    d_proj[e] = (r < n && c >= 3 * HC) ? d_out[(long long)r * HC + (c - 3 * HC)] : 0.f;
  }
}

struct EdgeAttrArgs {
  const void* lu;
  int lu_is_float;
  const int64_t* nbr;
  const void* t_edge;
  int t_is_float;
  const float* msg;
  const int64_t* msg_rows;
  DevCount edges;
  int bound;
  int De, Dt;
  const float* time_w;
  const float* time_b;
  float* ea;   // [E, Dt+De]
  float* rel;  // [E]
};

__global__ void attn_edge_attr_kernel(EdgeAttrArgs a) {
  pdl_wait();
  pdl_launch();
  const int E = a.edges.get();
  const int Din = a.Dt + a.De;
  const long long total = (long long)a.bound * Din;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total;
       x += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(x / Din), d = (int)(x - (long long)e * Din);
    if (e >= E) {
      a.ea[x] = 0.f;
      if (d == 0 && a.rel) a.rel[e] = 0.f;
      continue;
    }
    if (d < a.Dt) {
      const int64_t j = a.nbr[e];
      const long long pr = a.msg_rows ? a.msg_rows[e] : e;
      float rt;
      if (!a.lu_is_float && !a.t_is_float) {
        rt = (float)(reinterpret_cast<const int64_t*>(a.lu)[j] -
                     reinterpret_cast<const int64_t*>(a.t_edge)[pr]);
      } else {
        const float l = a.lu_is_float ? reinterpret_cast<const float*>(a.lu)[j]
                                      : (float)reinterpret_cast<const int64_t*>(a.lu)[j];
        const float t = a.t_is_float ? reinterpret_cast<const float*>(a.t_edge)[pr]
                                     : (float)reinterpret_cast<const int64_t*>(a.t_edge)[pr];
        rt = l - t;
      }
      a.ea[x] = cos_fr(__fmaf_rn(rt, a.time_w[d], a.time_b[d]));
      if (d == 0 && a.rel) a.rel[e] = rt;
    } else {
      const long long mr = a.msg_rows ? a.msg_rows[e] : e;
      a.ea[x] = a.msg[mr * a.De + (d - a.Dt)];
    }
  }
}

}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_attn_bwd_init(const float* d_out, int32_t num_rows, const int32_t* num_rows_dev,
                          int32_t hc, float* d_proj, void* stream) {
  TGN_REQUIRE(num_rows >= 0 && hc >= 1, "attn_bwd_init: bad sizes");
  if (num_rows == 0) return TGN_OK;
  TGN_REQUIRE(d_out && d_proj, "attn_bwd_init: NULL pointer");
  DevCount c{num_rows_dev, num_rows};
  launch_k(attn_bwd_init_kernel, dim3(stride_grid((long long)num_rows * 4 * hc, 256)), dim3(256), 0, (cudaStream_t)stream, d_out, c, num_rows, hc, d_proj);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_attn_bwd(const float* proj, const int64_t* nbr_local, const int32_t* row_ptr,
                     const int32_t* edge_perm, const int64_t* centre_ids, int32_t num_centres,
                     const int32_t* num_centres_dev, int32_t heads, int32_t head_dim,
                     const float* alpha, const float* ee, const float* d_out, float dropout_p,
                     uint64_t seed, const int64_t* seed_dev, float* d_proj, float* d_ee,
                     void* stream) {
  TGN_REQUIRE(num_centres >= 0 && heads >= 1 && heads <= kBMaxHeads && head_dim >= 1 &&
                  heads * head_dim <= 32 * kBMaxCH,
              "attn_bwd: bad sizes");
  if (num_centres == 0) return TGN_OK;
  TGN_REQUIRE(proj && nbr_local && row_ptr && alpha && ee && d_out && d_proj && d_ee,
              "attn_bwd: NULL pointer");
  AttnBwdArgs a;
  a.proj = proj; a.nbr = nbr_local; a.row_ptr = row_ptr; a.edge_perm = edge_perm;
  a.centre_ids = centre_ids; a.centres = DevCount{num_centres_dev, num_centres};
  a.H = heads; a.C = head_dim; a.alpha = alpha; a.ee = ee; a.d_out = d_out;
  a.dropout_p = dropout_p; a.seed = seed; a.seed_dev = seed_dev; a.d_proj = d_proj; a.d_ee = d_ee;
  int grid = ceil_div(num_centres, kBwdWarps);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  launch_k(attn_bwd_kernel, dim3(grid), dim3(kBwdWarps * 32), 0, (cudaStream_t)stream, a);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_attn_edge_attr(const void* last_update_local, int32_t lu_is_float,
                           const int64_t* nbr_local, const void* t_edge, int32_t t_is_float,
                           const float* msg, const int64_t* msg_rows, int32_t num_edges,
                           const int32_t* num_edges_dev, int32_t raw_dim, int32_t time_dim,
                           const float* time_w, const float* time_b, float* edge_attr,
                           float* rel_t, void* stream) {
  TGN_REQUIRE(num_edges >= 0 && raw_dim >= 0 && time_dim >= 0 && raw_dim + time_dim >= 1,
              "attn_edge_attr: bad sizes");
  if (num_edges == 0) return TGN_OK;
  TGN_REQUIRE(last_update_local && nbr_local && t_edge && (msg || raw_dim == 0) && edge_attr &&
                  (time_dim == 0 || (time_w && time_b)),
              "attn_edge_attr: NULL pointer");
  EdgeAttrArgs a;
  a.lu = last_update_local; a.lu_is_float = lu_is_float; a.nbr = nbr_local; a.t_edge = t_edge;
  a.t_is_float = t_is_float; a.msg = msg; a.msg_rows = msg_rows;
  a.edges = DevCount{num_edges_dev, num_edges}; a.bound = num_edges; a.De = raw_dim;
  a.Dt = time_dim; a.time_w = time_w; a.time_b = time_b; a.ea = edge_attr; a.rel = rel_t;
  launch_k(attn_edge_attr_kernel, dim3(stride_grid((long long)num_edges * (raw_dim + time_dim), 256)), dim3(256), 0, (cudaStream_t)stream, a);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
