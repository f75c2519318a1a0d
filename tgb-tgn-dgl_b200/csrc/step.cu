// Kernels of the captured training step that sit between the tensor-core GEMMs
// (engine.py drives them; every size that depends on the batch is read from device memory):
//
//   tgn_relabel3          the three `_assoc[...]` gathers of one step in one launch
//                         (neighbor_loader.py:48, epoch_utils.py:99,262)
//   tgn_edge_attr_ld      edge_attr = [cos(w*rel_t+b), msg] with a TMA-aligned row stride
//                         (modules/emb_module.py:26-28), plus sin(w*rel_t+b) for the backward
//   tgn_attn_core_fwd/bwd TransformerConv softmax / aggregation over each centre's edges with the
//                         edge projection ee = W_edge * edge_attr taken from the GEMM
//                         (modules/emb_module.py:29)
//   tgn_dec_loss          LinkPredictor tail + BCE-with-logits loss and its gradient
//                         (modules/decoder.py:24-27, pyg-mem-tgn.py:51, epoch_utils.py:305-310)
//   tgn_scatter_add_rows  dst[rows[i],:] += src[i,:]
//   tgn_time_bwd_sin      TimeEncoder gradient from stored sines
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

__global__ void relabel3_kernel(const int64_t* __restrict__ a, DevCount na, int64_t* __restrict__ oa,
                                const int64_t* __restrict__ b, DevCount nb, int64_t* __restrict__ ob,
                                const int64_t* __restrict__ c, DevCount nc, int64_t* __restrict__ oc,
                                const int64_t* __restrict__ assoc) {
  pdl_wait();
  pdl_launch();
  const int ca = na.get(), cb = nb.get(), cc = nc.get();
  const int total = ca + cb + cc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < ca) oa[i] = assoc[a[i]];
    else if (i < ca + cb) ob[i - ca] = assoc[b[i - ca]];
    else oc[i - ca - cb] = assoc[c[i - ca - cb]];
  }
}

struct EdgeAttr2Args {
  const int64_t* lu;      // last_update of the local rows (int64)
  const int64_t* nbr;     // local row of each edge's neighbour
  const int64_t* t_edge;  // event timestamps (int64), indexed by msg_rows[e]
  const float* msg;       // event messages [*, De], indexed by msg_rows[e]
  const int64_t* msg_rows;
  DevCount edges;
  int De, Dt, ld;
  const float* time_w;
  const float* time_b;
  float* ea;     // [E, ld]
  float* sn;     // [E, Dt] sin(w*rel+b) (nullable)
  float* rel;    // [E]
  long long num_events;  // rows of t_edge / msg (0 = unchecked)
  int32_t* err;
};

// One warp per edge, lanes over the columns of the row (coalesced 128-byte stores); the edge's scalars
// (event row, its time stamp, the neighbour's last_update) are fetched once per edge for kEdgesPerWarp edges
// at a time, all ahead of the first use.  Round 2: the first version mapped one thread to one OUTPUT ELEMENT
// with a 64-bit division and four dependent scalar loads per element (130 us for the 181k edges of a flight
// evaluation batch -- instruction-bound); kSin = false (evaluation) computes the cosine only.
// kEdgesPerWarp = 1 on step-sized launches (a few thousand edges: every edge gets its own warp, latency matters:
// wiki B=200 step 147 us against 159 us with four edges per warp and 150 us with the per-element kernel), 4 on
// large ones.
template <bool kSin, int kEdgesPerWarp>
__global__ void __launch_bounds__(256) edge_attr_ld_kernel(EdgeAttr2Args a) {
  pdl_wait();
  pdl_launch();
  const int E = a.edges.get();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int e0 = warp * kEdgesPerWarp; e0 < E; e0 += nwarps * kEdgesPerWarp) {
    long long mr[kEdgesPerWarp], nb[kEdgesPerWarp];
#pragma unroll
    for (int u = 0; u < kEdgesPerWarp; ++u) {
      const int e = e0 + u;
      mr[u] = -1;
      nb[u] = 0;
      if (e < E) {
        mr[u] = a.msg_rows ? a.msg_rows[e] : e;
        nb[u] = a.nbr[e];
      }
    }
    float rt[kEdgesPerWarp];
    bool ok[kEdgesPerWarp];
#pragma unroll
    for (int u = 0; u < kEdgesPerWarp; ++u) {
      ok[u] = e0 + u < E && !(a.num_events > 0 && (mr[u] < 0 || mr[u] >= a.num_events));
      rt[u] = ok[u] ? (float)(a.lu[nb[u]] - a.t_edge[mr[u]]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kEdgesPerWarp; ++u) {
      const int e = e0 + u;
      if (e >= E) break;                                  // warp-uniform
      float* row = a.ea + (long long)e * a.ld;
      if (!ok[u]) {   // the ring names an event that is not resident
        if (lane == 0) flag_dev_err(a.err, TGN_DEVERR_EVENT_RANGE);
        for (int d = lane; d < a.ld; d += 32) row[d] = 0.f;
        if (kSin)
          for (int d = lane; d < a.Dt; d += 32) a.sn[(long long)e * a.Dt + d] = 0.f;
        if (lane == 0 && a.rel) a.rel[e] = 0.f;
        continue;
      }
      for (int d = lane; d < a.ld; d += 32) {
        float v = 0.f;
        if (d < a.Dt) {
          const float arg = __fmaf_rn(rt[u], a.time_w[d], a.time_b[d]);
          if (kSin) {
            float sv;
            sincos_fr(arg, &sv, &v);
            a.sn[(long long)e * a.Dt + d] = sv;
          } else {
            v = cos_fr(arg);
          }
        } else if (d < a.Dt + a.De) {
          v = a.msg[mr[u] * a.De + (d - a.Dt)];
        }
        row[d] = v;
      }
      if (lane == 0 && a.rel) a.rel[e] = rt[u];
    }
  }
}

// d_w[c] += sum_i g[i,c] * (-sin_i,c) * t[i];  d_b[c] += sum_i g[i,c] * (-sin_i,c)
// rows with mask[i] < 0 are skipped (zero-message rows carry no time encoding).
__global__ void time_bwd_sin_kernel(const float* __restrict__ t, const int32_t* __restrict__ mask,
                                    DevCount num, const float* __restrict__ sn, int D,
                                    const float* __restrict__ g, int ldg, float* __restrict__ d_w,
                                    float* __restrict__ d_b) {
  pdl_wait();
  pdl_launch();
  __shared__ float s_w[8][33], s_b[8][33];
  const int R = num.get();
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const int rows_per_cta = (R + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
  float aw = 0.f, ab = 0.f;
  if (c < D) {
    for (int r = r0 + ry; r < r1; r += 8) {
      if (mask && mask[r] < 0) continue;
      const float ds = -sn[(long long)r * D + c] * g[(long long)r * ldg + c];
      aw = fmaf(ds, t[r], aw);
      ab += ds;
    }
  }
  s_w[ry][threadIdx.x & 31] = aw;
  s_b[ry][threadIdx.x & 31] = ab;
  __syncthreads();
  if (ry == 0 && c < D) {
    float vw = 0.f, vb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      vw += s_w[i][threadIdx.x & 31];
      vb += s_b[i][threadIdx.x & 31];
    }
    atomicAdd(&d_w[c], vw);
    atomicAdd(&d_b[c], vb);
  }
}

// ------------------------------------------------------------------------------------------
// attention core: one warp per centre, lanes own channels (lane, lane+32, ...)
// ------------------------------------------------------------------------------------------
constexpr int kCoreWarps = 4;
constexpr int kCMaxCH = 8;     // H*C <= 256
constexpr int kCMaxHeads = 8;

struct AttnCoreArgs {
  const float* proj;  // [Nb, 4HC] = q | k | v | skip
  const int64_t* nbr;
  const int32_t* row_ptr;
  const int64_t* centre_ids;
  DevCount centres;
  int H, C;
  const float* ee;  // [E, HC]
  float dropout_p;
  uint64_t seed;
  const int64_t* seed_dev;
  float* out;       // fwd: [Nb, HC] rows of centres written
  float* alpha;     // [E, H]
  const float* d_out;
  float* d_proj;    // bwd: [Nb, 4HC], zero on entry
  float* d_ee;      // bwd: [E, HC]
};

// Channel ownership: lane l holds the float4 chunks f = l and l + 32 (channels 4f..4f+3), so a
// projection / edge row is fetched with one or two 128-bit loads per lane and the k / v
// gradients leave as 128-bit vector reductions.  Edges are handled four at a time with all
// row loads issued before the first use.
constexpr int kEG = 4;  // edges per group

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
// hm[i][h] = 1 if channel c0+i belongs to head h: per-head sums / picks become plain FMAs
template <int H>
struct HeadMask {
  float m[4][H];
  __device__ __forceinline__ void init(int c0, int C, bool on) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int h = 0; h < H; ++h) m[i][h] = (on && (c0 + i) / C == h) ? 1.f : 0.f;
  }
  // part[h] += sum_i x_i * y_i over the channels of head h
  __device__ __forceinline__ void dot(float4 x, float4 y, float* part) const {
    const float pr[4] = {x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int h = 0; h < H; ++h) part[h] = fmaf(pr[i], m[i][h], part[h]);
  }
  // per-channel value of a per-head quantity
  __device__ __forceinline__ float4 pick(const float* v) const {
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int h = 0; h < H; ++h) o[i] = fmaf(v[h], m[i][h], o[i]);
    return make_float4(o[0], o[1], o[2], o[3]);
  }
};
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

template <int H>
__global__ void __launch_bounds__(kCoreWarps * 32) attn_core_fwd_kernel(AttnCoreArgs a) {
  pdl_wait();
  pdl_launch();
  const int HC = a.H * a.C, C = a.C;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nC = a.centres.get();
  const float inv_sqrt_c = rsqrtf((float)C);
  const float keep = 1.f - a.dropout_p;
  const int c0[2] = {4 * lane, 4 * (lane + 32)};
  const bool ok[2] = {c0[0] < HC, c0[1] < HC};
  HeadMask<H> hm[2];
  hm[0].init(c0[0], C, ok[0]);
  hm[1].init(c0[1], C, ok[1]);
  Philox rng(a.seed + (a.seed_dev ? (uint64_t)*a.seed_dev : 0ull));
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int ci = blockIdx.x * kCoreWarps + wid; ci < nC; ci += gridDim.x * kCoreWarps) {
    const int64_t row = a.centre_ids ? a.centre_ids[ci] : ci;
    const float* pr = a.proj + row * 4 * HC;
    float4 q[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) q[u] = ok[u] ? ld4(pr + c0[u]) : z4;
    const int e0 = a.row_ptr[ci], e1 = a.row_ptr[ci + 1];
    float mrun[H], lrun[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      mrun[h] = -INFINITY;
      lrun[h] = 0.f;
    }
    // ---- pass 1: raw scores -> alpha buffer, running max / normaliser
    for (int eb = e0; eb < e1; eb += kEG) {
      float4 kk[kEG][2], ee[kEG][2];
#pragma unroll
      for (int g = 0; g < kEG; ++g) {
        const int e = eb + g;
        const bool live = e < e1;
        const int64_t j = live ? a.nbr[e] : 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const bool on = live && ok[u];
          kk[g][u] = on ? ld4(a.proj + j * 4 * HC + HC + c0[u]) : z4;
          ee[g][u] = on ? ld4(a.ee + (long long)e * HC + c0[u]) : z4;
        }
      }
#pragma unroll
      for (int g = 0; g < kEG; ++g) {
        const int e = eb + g;
        if (e >= e1) break;  // warp-uniform
        float part[H];
#pragma unroll
        for (int h = 0; h < H; ++h) part[h] = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u) hm[u].dot(q[u], f4_add(kk[g][u], ee[g][u]), part);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          if (h < H) {
            const float sc = warp_sum(part[h]) * inv_sqrt_c;
            const float mnew = fmaxf(mrun[h], sc);
            lrun[h] = lrun[h] * expf(mrun[h] - mnew) + expf(sc - mnew);
            mrun[h] = mnew;
            if (lane == 0) a.alpha[(long long)e * H + h] = sc;
          }
        }
      }
    }
    __syncwarp();
    // ---- pass 2: softmax weights (kept for the backward), dropout, aggregation
    float4 acc[2] = {z4, z4};
    for (int eb = e0; eb < e1; eb += kEG) {
      float4 vv[kEG][2];
      float al[kEG][H];
#pragma unroll
      for (int g = 0; g < kEG; ++g) {
        const int e = eb + g;
        const bool live = e < e1;
        const int64_t j = live ? a.nbr[e] : 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const bool on = live && ok[u];
          vv[g][u] = on ? f4_add(ld4(a.proj + j * 4 * HC + 2 * HC + c0[u]), ld4(a.ee + (long long)e * HC + c0[u])) : z4;
        }
#pragma unroll
        for (int h = 0; h < H; ++h) al[g][h] = (live && h < H) ? a.alpha[(long long)e * H + h] : 0.f;
      }
      __syncwarp();  // every lane has read the raw scores before lane 0 overwrites them
#pragma unroll
      for (int g = 0; g < kEG; ++g) {
        const int e = eb + g;
        if (e >= e1) break;
        float pw[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          pw[h] = 0.f;
          if (h < H) {
            float p = expf(al[g][h] - mrun[h]) / lrun[h];
            if (lane == 0) a.alpha[(long long)e * H + h] = p;
            if (a.dropout_p > 0.f) {
              const uint4 r = rng((uint64_t)e, (uint64_t)h);
              const float uni = (float)(r.x >> 8) * (1.0f / 16777216.0f);
              p = uni < keep ? p / keep : 0.f;
            }
            pw[h] = p;
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 w4 = hm[u].pick(pw);
          acc[u].x = fmaf(w4.x, vv[g][u].x, acc[u].x);
          acc[u].y = fmaf(w4.y, vv[g][u].y, acc[u].y);
          acc[u].z = fmaf(w4.z, vv[g][u].z, acc[u].z);
          acc[u].w = fmaf(w4.w, vv[g][u].w, acc[u].w);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (ok[u]) *reinterpret_cast<float4*>(a.out + row * HC + c0[u]) = f4_add(acc[u], ld4(pr + 3 * HC + c0[u]));
  }
}

// ---- low-degree fast path (every centre has <= kSmallDeg edges, H*C <= 128): the whole
// neighbourhood of a centre is fetched into registers with ONE round of independent loads
// (neighbour ids by one coalesced load + shuffles, then all k / v / ee rows), so a warp pays
// ~4 dependent memory latencies per centre instead of ~2 per edge group and pass.
constexpr int kSmallDeg = 12;

// Dropout masks of one centre's (edge, head) pairs, lane-parallel: pair p = g*H + h is drawn by lane
// p % 32 in round p / 32 (same Philox counters (edge, head) and word as the per-pair draw of the
// general kernels, so the mask is bit-identical) -- one Philox evaluation per lane instead of
// kSmallDeg*H evaluations repeated by every lane, which was a third of these kernels' instructions.
template <int H, int kDeg = kSmallDeg>
struct SmallMask {
  static constexpr int kRounds = (kDeg * H + 31) / 32;
  float u[kRounds];
  __device__ __forceinline__ void draw(const Philox& rng, int e0, int deg, int lane) {
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const int p = r * 32 + lane;
      u[r] = 2.f;  // > any keep probability: unused pairs are "dropped"
      if (p < deg * H) {
        const uint4 x = rng((uint64_t)(e0 + p / H), (uint64_t)(p % H));
        u[r] = (float)(x.x >> 8) * (1.0f / 16777216.0f);
      }
    }
  }
  // uniform variate of (edge g, head h); g, h are compile-time after unrolling
  __device__ __forceinline__ float get(int g, int h) const {
    const int p = g * H + h;
    return __shfl_sync(0xffffffffu, u[p / 32], p % 32);
  }
};

// One centre of the low-degree path, by one warp: returns the lane's four output channels (aggregation + skip)
// and leaves the softmax weights in a.alpha.  Shared by the stand-alone kernel and the decoder-fused one.
// The scalars of one centre: its row, its first edge, its degree (capped), and in lane l the id of neighbour l.
struct CentreHead {
  int64_t row, my_j;
  int e0, deg;
};
template <int kDeg>
__device__ __forceinline__ CentreHead attn_centre_head(const AttnCoreArgs& a, int ci, int lane) {
  CentreHead h;
  h.row = a.centre_ids ? a.centre_ids[ci] : ci;
  h.e0 = a.row_ptr[ci];
  h.deg = min(a.row_ptr[ci + 1] - h.e0, kDeg);
  // the centre's edge rows are contiguous (deg x HC floats, streamed from DRAM): request them now, one 128-byte
  // line per lane, so they arrive while the neighbour ids and the k / v rows (L2-resident table) are fetched
  if (lane * 32 < h.deg * a.H * a.C) prefetch_l2(a.ee + (long long)h.e0 * (a.H * a.C) + lane * 32);
  h.my_j = lane < h.deg ? a.nbr[h.e0 + lane] : 0;
  return h;
}

template <int H, int kDeg>
__device__ __forceinline__ float4 attn_small_fwd_row(const AttnCoreArgs& a, const CentreHead& ch, int lane,
                                                     const HeadMask<H>& hm, const Philox& rng, float inv_sqrt_c,
                                                     float keep, bool ok, int c0) {
  const int HC = a.H * a.C;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t row = ch.row;
  const int e0 = ch.e0, deg = ch.deg;
  const float* pr = a.proj + row * 4 * HC;
  const float4 q = ok ? ld4(pr + c0) : z4;
  const float4 skip = ok ? ld4(pr + 3 * HC + c0) : z4;
  const int64_t my_j = ch.my_j;
  float4 kk[kDeg], vv[kDeg];
#pragma unroll
  for (int g = 0; g < kDeg; ++g) {
    const int64_t j = __shfl_sync(0xffffffffu, my_j, g);
    const bool on = ok && g < deg;
    const float4 eev = on ? ld4(a.ee + (long long)(e0 + g) * HC + c0) : z4;
    kk[g] = on ? f4_add(ld4(a.proj + j * 4 * HC + HC + c0), eev) : z4;
    vv[g] = on ? f4_add(ld4(a.proj + j * 4 * HC + 2 * HC + c0), eev) : z4;
  }
  float sc[kDeg][H], mx[H], den[H];
#pragma unroll
  for (int h = 0; h < H; ++h) mx[h] = -INFINITY;
#pragma unroll
  for (int g = 0; g < kDeg; ++g) {
    float part[H];
#pragma unroll
    for (int h = 0; h < H; ++h) part[h] = 0.f;
    hm.dot(q, kk[g], part);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      sc[g][h] = warp_sum(part[h]) * inv_sqrt_c;
      if (g < deg) mx[h] = fmaxf(mx[h], sc[g][h]);
    }
  }
#pragma unroll
  for (int h = 0; h < H; ++h) den[h] = 0.f;
#pragma unroll
  for (int g = 0; g < kDeg; ++g)
#pragma unroll
    for (int h = 0; h < H; ++h) {
      sc[g][h] = g < deg ? expf(sc[g][h] - mx[h]) : 0.f;
      den[h] += sc[g][h];
    }
  float4 acc = z4;
  SmallMask<H, kDeg> sm;
  if (a.dropout_p > 0.f) sm.draw(rng, e0, deg, lane);
#pragma unroll
  for (int g = 0; g < kDeg; ++g) {
    if (g >= deg) break;  // warp-uniform
    float pw[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float p = sc[g][h] / den[h];
      if (lane == 0) a.alpha[(long long)(e0 + g) * H + h] = p;
      if (a.dropout_p > 0.f) p = sm.get(g, h) < keep ? p / keep : 0.f;
      pw[h] = p;
    }
    const float4 w4 = hm.pick(pw);
    acc.x = fmaf(w4.x, vv[g].x, acc.x);
    acc.y = fmaf(w4.y, vv[g].y, acc.y);
    acc.z = fmaf(w4.z, vv[g].z, acc.z);
    acc.w = fmaf(w4.w, vv[g].w, acc.w);
  }
  return f4_add(acc, skip);
}

// kDeg = register-resident edge slots per centre (10 for the TGN default of 10 neighbours: 125 registers, four
// CTAs per SM on evaluation-sized launches; 12 otherwise)
template <int H, int kDeg, int kMinCtas>
__global__ void __launch_bounds__(kCoreWarps * 32, kMinCtas) attn_core_fwd_small_kernel(AttnCoreArgs a) {
  pdl_wait();
  pdl_launch();
  const int HC = a.H * a.C, C = a.C;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nC = a.centres.get();
  const float inv_sqrt_c = rsqrtf((float)C);
  const float keep = 1.f - a.dropout_p;
  const int c0 = 4 * lane;
  const bool ok = c0 < HC;
  HeadMask<H> hm;
  hm.init(c0, C, ok);
  Philox rng(a.seed + (a.seed_dev ? (uint64_t)*a.seed_dev : 0ull));
  // a warp that walks several centres (capped grids of evaluation-sized launches) fetches the NEXT centre's
  // scalars -- two dependent round trips -- while it computes the current one
  const int stride = gridDim.x * kCoreWarps;
  int ci = blockIdx.x * kCoreWarps + wid;
  if (ci >= nC) return;
  CentreHead cur = attn_centre_head<kDeg>(a, ci, lane);
  for (; ci < nC; ci += stride) {
    CentreHead nxt = cur;
    if (ci + stride < nC) nxt = attn_centre_head<kDeg>(a, ci + stride, lane);
    const float4 o = attn_small_fwd_row<H, kDeg>(a, cur, lane, hm, rng, inv_sqrt_c, keep, ok, c0);
    if (ok) *reinterpret_cast<float4*>(a.out + cur.row * HC + c0) = o;
    cur = nxt;
  }
}

template <int H>
__global__ void __launch_bounds__(kCoreWarps * 32) attn_core_bwd_small_kernel(AttnCoreArgs a) {
  pdl_wait();
  pdl_launch();
  const int HC = a.H * a.C, C = a.C;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nC = a.centres.get();
  const float inv_sqrt_c = rsqrtf((float)C);
  const float keep = 1.f - a.dropout_p;
  const int c0 = 4 * lane;
  const bool ok = c0 < HC;
  HeadMask<H> hm;
  hm.init(c0, C, ok);
  Philox rng(a.seed + (a.seed_dev ? (uint64_t)*a.seed_dev : 0ull));
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int ci = blockIdx.x * kCoreWarps + wid; ci < nC; ci += gridDim.x * kCoreWarps) {
    const int64_t row = a.centre_ids ? a.centre_ids[ci] : ci;
    const int e0 = a.row_ptr[ci];
    const int deg = min(a.row_ptr[ci + 1] - e0, kSmallDeg);
    const float* pr = a.proj + row * 4 * HC;
    const float4 q = ok ? ld4(pr + c0) : z4;
    const float4 g4 = ok ? ld4(a.d_out + row * HC + c0) : z4;
    const int64_t my_j = lane < deg ? a.nbr[e0 + lane] : 0;
    // lane l < deg*H holds alpha of (edge l / H, head l % H): one coalesced load
    const float my_al = lane < deg * H ? a.alpha[(long long)e0 * H + lane] : 0.f;
    float4 kk[kSmallDeg], vv[kSmallDeg];
    int64_t jn[kSmallDeg];
#pragma unroll
    for (int g = 0; g < kSmallDeg; ++g) {
      jn[g] = __shfl_sync(0xffffffffu, my_j, g);
      const bool on = ok && g < deg;
      const float4 eev = on ? ld4(a.ee + (long long)(e0 + g) * HC + c0) : z4;
      kk[g] = on ? f4_add(ld4(a.proj + jn[g] * 4 * HC + HC + c0), eev) : z4;
      vv[g] = on ? f4_add(ld4(a.proj + jn[g] * 4 * HC + 2 * HC + c0), eev) : z4;
    }
    float al[kSmallDeg][H], dal[kSmallDeg][H], mk[kSmallDeg][H], dot[H];
    SmallMask<H> sm;
    if (a.dropout_p > 0.f) sm.draw(rng, e0, deg, lane);
#pragma unroll
    for (int h = 0; h < H; ++h) dot[h] = 0.f;
#pragma unroll
    for (int g = 0; g < kSmallDeg; ++g) {
      float part[H];
#pragma unroll
      for (int h = 0; h < H; ++h) part[h] = 0.f;
      hm.dot(g4, vv[g], part);
#pragma unroll
      for (int h = 0; h < H; ++h) {
        al[g][h] = (g * H + h < 32) ? __shfl_sync(0xffffffffu, my_al, g * H + h)
                                    : (g < deg ? a.alpha[(long long)(e0 + g) * H + h] : 0.f);
        mk[g][h] = 1.f;
        if (a.dropout_p > 0.f) {   // (warp-uniform branch; pairs past the degree are never used)
          const float uni = sm.get(g, h);
          if (g < deg) mk[g][h] = uni < keep ? 1.f / keep : 0.f;
        }
        dal[g][h] = warp_sum(part[h]) * mk[g][h];
        if (g < deg) dot[h] = fmaf(al[g][h], dal[g][h], dot[h]);
      }
    }
    float4 dq = z4;
#pragma unroll
    for (int g = 0; g < kSmallDeg; ++g) {
      if (g >= deg) break;  // warp-uniform
      float ds[H], at[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        ds[h] = al[g][h] * (dal[g][h] - dot[h]) * inv_sqrt_c;
        at[h] = al[g][h] * mk[g][h];
      }
      if (!ok) continue;
      const float4 d4 = hm.pick(ds), a4 = hm.pick(at);
      dq.x = fmaf(d4.x, kk[g].x, dq.x);
      dq.y = fmaf(d4.y, kk[g].y, dq.y);
      dq.z = fmaf(d4.z, kk[g].z, dq.z);
      dq.w = fmaf(d4.w, kk[g].w, dq.w);
      const float4 dk = make_float4(d4.x * q.x, d4.y * q.y, d4.z * q.z, d4.w * q.w);
      const float4 dv = make_float4(a4.x * g4.x, a4.y * g4.y, a4.z * g4.z, a4.w * g4.w);
      float* dpj = a.d_proj + jn[g] * 4 * HC;
      red_add_v4(dpj + HC + c0, dk);
      red_add_v4(dpj + 2 * HC + c0, dv);
      *reinterpret_cast<float4*>(a.d_ee + (long long)(e0 + g) * HC + c0) = f4_add(dk, dv);
    }
    if (ok) {
      float* dpr = a.d_proj + row * 4 * HC;
      *reinterpret_cast<float4*>(dpr + c0) = dq;
      *reinterpret_cast<float4*>(dpr + 3 * HC + c0) = g4;
    }
  }
}

// out_i = sum_e a~_e (v_j + ee_e) + skip_i,  a~ = dropout(alpha), alpha = softmax_e(s_e),
// s_e = <q_i, k_j + ee_e>/sqrt(C).  Pass A: dot_h = sum_e alpha_e d alpha_e; pass B recomputes
// d alpha and emits the gradients; neighbour rows of d_proj receive vector reductions.
template <int H>
__global__ void __launch_bounds__(kCoreWarps * 32) attn_core_bwd_kernel(AttnCoreArgs a) {
  pdl_wait();
  pdl_launch();
  const int HC = a.H * a.C, C = a.C;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nC = a.centres.get();
  const float inv_sqrt_c = rsqrtf((float)C);
  const float keep = 1.f - a.dropout_p;
  const int c0[2] = {4 * lane, 4 * (lane + 32)};
  const bool ok[2] = {c0[0] < HC, c0[1] < HC};
  HeadMask<H> hm[2];
  hm[0].init(c0[0], C, ok[0]);
  hm[1].init(c0[1], C, ok[1]);
  Philox rng(a.seed + (a.seed_dev ? (uint64_t)*a.seed_dev : 0ull));
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int ci = blockIdx.x * kCoreWarps + wid; ci < nC; ci += gridDim.x * kCoreWarps) {
    const int64_t row = a.centre_ids ? a.centre_ids[ci] : ci;
    const float* pr = a.proj + row * 4 * HC;
    float4 q[2], g4[2], dq[2] = {z4, z4};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      q[u] = ok[u] ? ld4(pr + c0[u]) : z4;
      g4[u] = ok[u] ? ld4(a.d_out + row * HC + c0[u]) : z4;
    }
    const int e0 = a.row_ptr[ci], e1 = a.row_ptr[ci + 1];
    float dot[H];
#pragma unroll
    for (int h = 0; h < H; ++h) dot[h] = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
      for (int eb = e0; eb < e1; eb += kEG) {
        float4 kk[kEG][2], vv[kEG][2];
        float al[kEG][H];
        int64_t jn[kEG];
#pragma unroll
        for (int g = 0; g < kEG; ++g) {
          const int e = eb + g;
          const bool live = e < e1;
          jn[g] = live ? a.nbr[e] : 0;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const bool on = live && ok[u];
            const float4 eev = on ? ld4(a.ee + (long long)e * HC + c0[u]) : z4;
            vv[g][u] = on ? f4_add(ld4(a.proj + jn[g] * 4 * HC + 2 * HC + c0[u]), eev) : z4;
            kk[g][u] = (on && pass == 1) ? f4_add(ld4(a.proj + jn[g] * 4 * HC + HC + c0[u]), eev) : z4;
          }
#pragma unroll
          for (int h = 0; h < H; ++h) al[g][h] = (live && h < H) ? a.alpha[(long long)e * H + h] : 0.f;
        }
#pragma unroll
        for (int g = 0; g < kEG; ++g) {
          const int e = eb + g;
          if (e >= e1) break;
          float part[H], dal[H], mk[H];
#pragma unroll
          for (int h = 0; h < H; ++h) part[h] = 0.f;
#pragma unroll
          for (int u = 0; u < 2; ++u) hm[u].dot(g4[u], vv[g][u], part);
#pragma unroll
          for (int h = 0; h < H; ++h) {
            mk[h] = 1.f;
            dal[h] = 0.f;
            if (h < H) {
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
            for (int h = 0; h < H; ++h) dot[h] = fmaf(al[g][h], dal[h], dot[h]);
            continue;
          }
          float ds[H], at[H];
#pragma unroll
          for (int h = 0; h < H; ++h) {
            ds[h] = al[g][h] * (dal[h] - dot[h]) * inv_sqrt_c;
            at[h] = al[g][h] * mk[h];
          }
          float* dpj = a.d_proj + jn[g] * 4 * HC;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (!ok[u]) continue;
            const float4 d4 = hm[u].pick(ds), a4 = hm[u].pick(at);
            dq[u].x = fmaf(d4.x, kk[g][u].x, dq[u].x);
            dq[u].y = fmaf(d4.y, kk[g][u].y, dq[u].y);
            dq[u].z = fmaf(d4.z, kk[g][u].z, dq[u].z);
            dq[u].w = fmaf(d4.w, kk[g][u].w, dq[u].w);
            const float4 dk = make_float4(d4.x * q[u].x, d4.y * q[u].y, d4.z * q[u].z, d4.w * q[u].w);
            const float4 dv = make_float4(a4.x * g4[u].x, a4.y * g4[u].y, a4.z * g4[u].z, a4.w * g4[u].w);
            red_add_v4(dpj + HC + c0[u], dk);
            red_add_v4(dpj + 2 * HC + c0[u], dv);
            *reinterpret_cast<float4*>(a.d_ee + (long long)e * HC + c0[u]) = f4_add(dk, dv);
          }
        }
      }
    }
    float* dpr = a.d_proj + row * 4 * HC;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      // centres are unique: plain stores for the q and skip blocks (a centre that is also a
      // neighbour only ever receives k / v reductions)
      *reinterpret_cast<float4*>(dpr + c0[u]) = dq[u];
      *reinterpret_cast<float4*>(dpr + 3 * HC + c0[u]) = g4[u];
    }
  }
}

// ------------------------------------------------------------------------------------------
// decoder tail + loss.  Event i (< B) owns the positive pair (src_i, dst_i) = row i of hd and
// the negative pair (src_i, neg_i) = row B+i of hd; both share hs row i.
//   h = relu(hs + hd), logit = wf.h + bf, loss = mean softplus(-pos) + mean softplus(neg)
// One warp per event; CTA-level reduction of the parameter gradients, then atomics.
// ------------------------------------------------------------------------------------------
constexpr int kDecWarps = 8;

__global__ void __launch_bounds__(kDecWarps * 32)
    dec_loss_kernel(const float* __restrict__ hs, const float* __restrict__ hd,
                    const float* __restrict__ wf, const float* __restrict__ bf, int B, int D,
                    float* __restrict__ loss, float* __restrict__ logits, float* __restrict__ dh,
                    float* __restrict__ dhs, float* __restrict__ d_wf, float* __restrict__ d_bf,
                    float* __restrict__ d_bs, float* __restrict__ d_bd) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float s_red[];  // [3][D] : d_wf, d_bs, d_bd partials
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  const float invB = 1.f / (float)B;
  float loss_acc = 0.f, dbf_acc = 0.f;
  for (int i = blockIdx.x * kDecWarps + wid; i < B; i += gridDim.x * kDecWarps) {
    float dl[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const float* ph = hd + (long long)(s * B + i) * D;
      float acc = 0.f;
      for (int c = lane; c < D; c += 32)
        acc = fmaf(fmaxf(hs[(long long)i * D + c] + ph[c], 0.f), wf[c], acc);
      const float logit = warp_sum(acc) + bf[0];
      // softplus(x) = max(x,0) + log1p(exp(-|x|)), x = -logit for positives
      const float x = s == 0 ? -logit : logit;
      loss_acc += (fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)))) * invB;
      const float sg = 1.f / (1.f + expf(-logit));
      dl[s] = (sg - (s == 0 ? 1.f : 0.f)) * invB;
      dbf_acc += dl[s];
      if (logits && lane == 0) logits[s * B + i] = logit;
    }
    for (int c = lane; c < D; c += 32) {
      const float a = hs[(long long)i * D + c];
      const float h0 = fmaxf(a + hd[(long long)i * D + c], 0.f);
      const float h1 = fmaxf(a + hd[(long long)(B + i) * D + c], 0.f);
      const float w = wf[c];
      const float g0 = h0 > 0.f ? dl[0] * w : 0.f;
      const float g1 = h1 > 0.f ? dl[1] * w : 0.f;
      dh[(long long)i * D + c] = g0;
      dh[(long long)(B + i) * D + c] = g1;
      dhs[(long long)i * D + c] = g0 + g1;
      atomicAdd(&s_red[c], dl[0] * h0 + dl[1] * h1);
      atomicAdd(&s_red[D + c], g0 + g1);
      atomicAdd(&s_red[2 * D + c], g0 + g1);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    if (s_red[c] != 0.f) atomicAdd(&d_wf[c], s_red[c]);
    if (s_red[D + c] != 0.f) atomicAdd(&d_bs[c], s_red[D + c]);
    if (s_red[2 * D + c] != 0.f) atomicAdd(&d_bd[c], s_red[2 * D + c]);
  }
  if (lane == 0) {  // every lane of a warp holds the same loss_acc / dbf_acc
    if (loss_acc != 0.f) atomicAdd(loss, loss_acc);
    if (dbf_acc != 0.f) atomicAdd(d_bf, dbf_acc);
  }
}

// ------------------------------------------------------------------------------------------
// Fused decoder forward + backward for the training step (decoder.py:24-27 + BCE-with-logits):
// per event i the pairs (src_i,dst_i) and (src_i,neg_i)
//   hs = Ws z_src + bs, hp = Wd z_dst + bd, hn = Wd z_neg + bd, h0 = relu(hs+hp), h1 = relu(hs+hn)
//   logit = wf.h + bf, loss = mean softplus(-logit0) + mean softplus(logit1)
// and, with d loss/d logit in hand, every gradient of the decoder: d z rows (atomics into the
// embedding-gradient rows), dWs, dWd (outer products accumulated in shared memory, flushed once
// per CTA as vector reductions), the biases and wf/bf.  Both 100x100 weights and both
// accumulators live in shared memory (row stride D+1: conflict-free for row- and column-wise
// walks); a CTA walks its events one after the other, thread c owns channel c.
// Exact fp32 FMA arithmetic -- replaces 2 gathers, 6 GEMMs, the loss kernel and a scatter.
// ------------------------------------------------------------------------------------------
constexpr int kDecThreads = 512;  // thread (c, q): channel c = tid & 127, reduction quarter q = tid >> 7
constexpr int kDecQ = kDecThreads / 128;

// kDefer: the two [D,D] weight gradients are NOT formed here; the kernel writes the gathered inputs
// z_rows [3B,D] = [z_src; z_dst; z_neg] and the hidden-layer gradients g_rows [3B,D] = [gs; g0; g1]
// instead, and the caller forms dW_src = gs^T z_src, dW_dst = [g0;g1]^T [z_dst;z_neg] as two small
// tensor-core GEMMs off the dependent chain (no shared-memory accumulators, no 20k-atomic flush per CTA).
constexpr int kDecEv = 4;          // events per pass (= kDecQ: the quarter index doubles as the event slot)
constexpr int kDecVec = 3 * kDecEv;  // [row s|d|n][event] values per channel in the staged tiles

// A pass handles kDecEv events.  Thread (c, q):
//   GEMV phases     channel c, reduction quarter q, ALL kDecEv events at once: one weight load feeds
//                   kDecEv (x3 rows) FMAs, the events' inputs come as three broadcast float4 loads;
//   pointwise phases channel c of event slot q (bias, relu, final layer, loss gradient).
// kAttn: the kernel does not gather its inputs from an embedding table -- warp v of a pass COMPUTES the attention
// output row of occurrence v (attn_small_fwd_row: H = 2, <= 10 edges per centre; ids_r = index of every batch id in
// the root list) straight into the staged tile, so the attention forward is not a launch (and a link of the
// step's dependent chain) of its own.  A node that occurs in several events is computed once per occurrence
// (identical bits; its softmax weights are written more than once with the same values).
template <bool kDefer, bool kAttn>
__global__ void __launch_bounds__(kDecThreads)
    dec_fused_kernel(const float* __restrict__ emb, const int64_t* __restrict__ ids_l, int B, int D,
                     const float* __restrict__ Ws, const float* __restrict__ bs,
                     const float* __restrict__ Wd, const float* __restrict__ bd,
                     const float* __restrict__ wf, const float* __restrict__ bf,
                     float* __restrict__ loss, float* __restrict__ logits, float* __restrict__ d_emb,
                     float* __restrict__ dWs, float* __restrict__ dbs, float* __restrict__ dWd,
                     float* __restrict__ dbd, float* __restrict__ dwf, float* __restrict__ dbf,
                     float* __restrict__ z_rows, float* __restrict__ g_rows, AttnCoreArgs at,
                     const int64_t* __restrict__ ids_r) {
  extern __shared__ __align__(16) float sm[];
  const int ld = D + 1;
  float* sz = sm;                        // [D][kDecVec] inputs      (16-byte aligned rows)
  float* sg = sz + D * kDecVec;          // [D][kDecVec] hidden-layer gradients
  float* sWs = sg + D * kDecVec;
  float* sWd = sWs + D * ld;
  float* sAs = sWd + D * ld;   // dWs accumulator (absent with kDefer)
  float* sAd = sAs + D * ld;   // dWd accumulator (absent with kDefer)
  __shared__ float s_p[kDecVec][kDecQ][128];  // partial sums of the quarters
  __shared__ float s_part[kDecEv][2][4];
  __shared__ int64_t s_row[kDecEv][3];
  const int tid = threadIdx.x, lane = tid & 31;
  const int c = tid & 127, q = tid >> 7;
  const int per = (D + kDecQ - 1) / kDecQ;
  const int k0 = q * per, k1 = min(D, k0 + per);
  for (int i = tid; i < D * D; i += kDecThreads) {
    const int r = i / D, k = i - r * D;
    sWs[r * ld + k] = Ws[i];
    sWd[r * ld + k] = Wd[i];
    if (!kDefer) {
      sAs[r * ld + k] = 0.f;
      sAd[r * ld + k] = 0.f;
    }
  }
  // The weights were written by the optimiser of the PREVIOUS step -- at least two launches back, which a
  // programmatic launch chain has always completed (the predecessor released this kernel only after its own
  // griddepcontrol.wait returned) -- so they are staged ahead of the wait, beside the predecessor's tail.
  pdl_wait();
  pdl_launch();
  const float invB = 1.f / (float)B;
  const float bfv = bf[0];
  float loss_acc = 0.f, dbf_acc = 0.f, dwf_acc = 0.f, dbs_acc = 0.f;
  const int npass = (B + kDecEv - 1) / kDecEv;
  for (int pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    const int e0 = pass * kDecEv;
    __syncthreads();  // weights staged / previous pass fully consumed
    if (tid < kDecVec) {
      const int row = tid / kDecEv, slot = tid - row * kDecEv;
      s_row[slot][row] = e0 + slot < B ? ids_l[row * B + e0 + slot] : -1;
    }
    __syncthreads();
    if (kAttn) {
      const int v = tid >> 5;                        // warp v: occurrence v = row * kDecEv + slot of the pass
      if (v < kDecVec) {
        const int row = v / kDecEv, slot = v - row * kDecEv;
        const bool live = e0 + slot < B;
        const int c0 = 4 * lane;
        const bool ok = c0 < D;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
          HeadMask<2> hm;
          hm.init(c0, at.C, ok);
          Philox rng(at.seed + (at.seed_dev ? (uint64_t)*at.seed_dev : 0ull));
          const CentreHead ch = attn_centre_head<10>(at, (int)ids_r[row * B + e0 + slot], lane);
          o = attn_small_fwd_row<2, 10>(at, ch, lane, hm, rng, rsqrtf((float)at.C), 1.f - at.dropout_p, ok, c0);
        }
        if (ok) {
          sz[(c0 + 0) * kDecVec + v] = o.x;
          sz[(c0 + 1) * kDecVec + v] = o.y;
          sz[(c0 + 2) * kDecVec + v] = o.z;
          sz[(c0 + 3) * kDecVec + v] = o.w;
          if (kDefer && live) *reinterpret_cast<float4*>(z_rows + ((long long)row * B + e0 + slot) * D + c0) = o;
        }
      }
    } else {
      for (int i = tid; i < kDecVec * D; i += kDecThreads) {
        const int v = i / D, k = i - v * D;           // v = row * kDecEv + slot
        const int row = v / kDecEv, slot = v - row * kDecEv;
        const int64_t r = s_row[slot][row];
        const float x = r >= 0 ? emb[r * D + k] : 0.f;
        sz[k * kDecVec + v] = x;
        if (kDefer && r >= 0) z_rows[((long long)row * B + e0 + slot) * D + k] = x;
      }
    }
    __syncthreads();
    // ---- forward partials: output channel c, inputs [k0, k1), all events of the pass
    {
      float hs[kDecEv] = {}, hp[kDecEv] = {}, hn[kDecEv] = {};
      if (c < D) {
        const float* ws = sWs + c * ld;
        const float* wd = sWd + c * ld;
        for (int k = k0; k < k1; ++k) {
          const float a = ws[k], b = wd[k];
          const float4 zs = *reinterpret_cast<const float4*>(sz + k * kDecVec);
          const float4 zd = *reinterpret_cast<const float4*>(sz + k * kDecVec + 4);
          const float4 zn = *reinterpret_cast<const float4*>(sz + k * kDecVec + 8);
          hs[0] = fmaf(a, zs.x, hs[0]); hs[1] = fmaf(a, zs.y, hs[1]); hs[2] = fmaf(a, zs.z, hs[2]); hs[3] = fmaf(a, zs.w, hs[3]);
          hp[0] = fmaf(b, zd.x, hp[0]); hp[1] = fmaf(b, zd.y, hp[1]); hp[2] = fmaf(b, zd.z, hp[2]); hp[3] = fmaf(b, zd.w, hp[3]);
          hn[0] = fmaf(b, zn.x, hn[0]); hn[1] = fmaf(b, zn.y, hn[1]); hn[2] = fmaf(b, zn.z, hn[2]); hn[3] = fmaf(b, zn.w, hn[3]);
        }
      }
#pragma unroll
      for (int e = 0; e < kDecEv; ++e) {
        s_p[e][q][c] = hs[e];
        s_p[kDecEv + e][q][c] = hp[e];
        s_p[2 * kDecEv + e][q][c] = hn[e];
      }
    }
    __syncthreads();
    // ---- pointwise: channel c of event slot q
    const bool live = e0 + q < B;
    float h0 = 0.f, h1 = 0.f;
    {
      float p0 = 0.f, p1 = 0.f;
      if (c < D && live) {
        float hs = bs[c], hp = bd[c], hn = hp;
#pragma unroll
        for (int j = 0; j < kDecQ; ++j) {
          hs += s_p[q][j][c];
          hp += s_p[kDecEv + q][j][c];
          hn += s_p[2 * kDecEv + q][j][c];
        }
        h0 = fmaxf(hs + hp, 0.f);
        h1 = fmaxf(hs + hn, 0.f);
        p0 = wf[c] * h0;
        p1 = wf[c] * h1;
      }
      p0 = warp_sum(p0);
      p1 = warp_sum(p1);
      if (lane == 0) {
        s_part[q][0][(tid >> 5) & 3] = p0;
        s_part[q][1][(tid >> 5) & 3] = p1;
      }
    }
    __syncthreads();
    {
      const float l0 = bfv + s_part[q][0][0] + s_part[q][0][1] + s_part[q][0][2] + s_part[q][0][3];
      const float l1 = bfv + s_part[q][1][0] + s_part[q][1][1] + s_part[q][1][2] + s_part[q][1][3];
      const float dl0 = live ? (1.f / (1.f + expf(-l0)) - 1.f) * invB : 0.f;
      const float dl1 = live ? (1.f / (1.f + expf(-l1))) * invB : 0.f;
      if (c == 0 && live) {
        if (logits) {
          logits[e0 + q] = l0;
          logits[B + e0 + q] = l1;
        }
        // softplus(x) = max(x,0) + log1p(exp(-|x|)); x = -logit for the positive pair
        loss_acc += (fmaxf(-l0, 0.f) + log1pf(expf(-fabsf(l0))) + fmaxf(l1, 0.f) + log1pf(expf(-fabsf(l1)))) * invB;
        dbf_acc += dl0 + dl1;
      }
      // ---- backward through the final layer / relu
      if (c < D) {
        const float w = wf[c];
        const float g0 = h0 > 0.f ? dl0 * w : 0.f, g1 = h1 > 0.f ? dl1 * w : 0.f;
        const float gs = g0 + g1;
        dwf_acc += dl0 * h0 + dl1 * h1;
        dbs_acc += gs;
        sg[c * kDecVec + q] = gs;
        sg[c * kDecVec + kDecEv + q] = g0;
        sg[c * kDecVec + 2 * kDecEv + q] = g1;
        if (kDefer && live) {
          g_rows[(long long)(e0 + q) * D + c] = gs;
          g_rows[(long long)(B + e0 + q) * D + c] = g0;
          g_rows[(long long)(2 * B + e0 + q) * D + c] = g1;
        }
      }
    }
    __syncthreads();
    if (!kDefer && c < D) {   // weight-gradient rows of channel c over [k0, k1), all events of the pass
      float* as = sAs + c * ld;
      float* ad = sAd + c * ld;
      const float4 gs = *reinterpret_cast<const float4*>(sg + c * kDecVec);
      const float4 g0 = *reinterpret_cast<const float4*>(sg + c * kDecVec + 4);
      const float4 g1 = *reinterpret_cast<const float4*>(sg + c * kDecVec + 8);
      for (int k = k0; k < k1; ++k) {
        const float4 zs = *reinterpret_cast<const float4*>(sz + k * kDecVec);
        const float4 zd = *reinterpret_cast<const float4*>(sz + k * kDecVec + 4);
        const float4 zn = *reinterpret_cast<const float4*>(sz + k * kDecVec + 8);
        as[k] += gs.x * zs.x + gs.y * zs.y + gs.z * zs.z + gs.w * zs.w;
        ad[k] += g0.x * zd.x + g0.y * zd.y + g0.z * zd.z + g0.w * zd.w +
                 g1.x * zn.x + g1.y * zn.y + g1.z * zn.z + g1.w * zn.w;
      }
    }
    // ---- input gradients: input channel c, outputs [k0, k1) (column walks of the weights)
    {
      float ds[kDecEv] = {}, dd[kDecEv] = {}, dn[kDecEv] = {};
      if (c < D) {
        for (int r = k0; r < k1; ++r) {
          const float a = sWs[r * ld + c], b = sWd[r * ld + c];
          const float4 gs = *reinterpret_cast<const float4*>(sg + r * kDecVec);
          const float4 g0 = *reinterpret_cast<const float4*>(sg + r * kDecVec + 4);
          const float4 g1 = *reinterpret_cast<const float4*>(sg + r * kDecVec + 8);
          ds[0] = fmaf(a, gs.x, ds[0]); ds[1] = fmaf(a, gs.y, ds[1]); ds[2] = fmaf(a, gs.z, ds[2]); ds[3] = fmaf(a, gs.w, ds[3]);
          dd[0] = fmaf(b, g0.x, dd[0]); dd[1] = fmaf(b, g0.y, dd[1]); dd[2] = fmaf(b, g0.z, dd[2]); dd[3] = fmaf(b, g0.w, dd[3]);
          dn[0] = fmaf(b, g1.x, dn[0]); dn[1] = fmaf(b, g1.y, dn[1]); dn[2] = fmaf(b, g1.z, dn[2]); dn[3] = fmaf(b, g1.w, dn[3]);
        }
      }
#pragma unroll
      for (int e = 0; e < kDecEv; ++e) {
        s_p[e][q][c] = ds[e];
        s_p[kDecEv + e][q][c] = dd[e];
        s_p[2 * kDecEv + e][q][c] = dn[e];
      }
    }
    __syncthreads();
    if (c < D && live) {   // the three gradient rows of event slot q
#pragma unroll
      for (int row = 0; row < 3; ++row) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < kDecQ; ++j) v += s_p[row * kDecEv + q][j][c];
        atomicAdd(&d_emb[s_row[q][row] * D + c], v);
      }
    }
  }
  __syncthreads();
  // ---- flush the per-thread and shared accumulators
  if (c < D && (dwf_acc != 0.f || dbs_acc != 0.f)) {
    atomicAdd(&dwf[c], dwf_acc);
    atomicAdd(&dbs[c], dbs_acc);
    atomicAdd(&dbd[c], dbs_acc);
  }
  if (!kDefer) {
    for (int i = tid; i < D * D; i += kDecThreads) {
      const int r = i / D, k = i - r * D;
      const float a = sAs[r * ld + k], b = sAd[r * ld + k];
      if (a != 0.f) atomicAdd(&dWs[i], a);
      if (b != 0.f) atomicAdd(&dWd[i], b);
    }
  }
  if (loss_acc != 0.f) atomicAdd(loss, loss_acc);
  if (dbf_acc != 0.f) atomicAdd(dbf, dbf_acc);
}

// ------------------------------------------------------------------------------------------
// TGB evaluation: score one positive and its Q negatives and count how many negatives beat it.
// One CTA per positive; hs[src] is staged once in shared memory, every warp then streams
// negatives (lanes over channels).  Scores are sigmoid outputs compared in fp32, exactly what
// the reference hands to the TGB evaluator (decoder.py:27, epoch_utils.py:99-113).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    score_negs_kernel(const float* __restrict__ hs, const float* __restrict__ hd,
                      const int64_t* __restrict__ src_rows, const int64_t* __restrict__ dst_rows,
                      const int64_t* __restrict__ neg_rows, int B, int Q, int D,
                      const float* __restrict__ wf, const float* __restrict__ bf,
                      float* __restrict__ pos_out, float* __restrict__ neg_out,
                      int32_t* __restrict__ gt_out, int32_t* __restrict__ ge_out) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float s_a[];  // [D] hs row of the source, [D] w_final
  __shared__ float s_pos;
  __shared__ int s_gt, s_ge;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // blockIdx.y = part of the negative columns (the chain index -> row -> reduce of one negative is ~1 us per
  // warp iteration, so a positive's Q negatives are spread over gridDim.y CTAs); counts add up atomically
  const int parts = gridDim.y, part = blockIdx.y;
  const int qlo = (int)((long long)Q * part / parts), qhi = (int)((long long)Q * (part + 1) / parts);
  for (int i = blockIdx.x; i < B; i += gridDim.x) {
    __syncthreads();
    const float* pa = hs + src_rows[i] * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      s_a[c] = pa[c];
      s_a[D + c] = wf[c];
    }
    if (threadIdx.x == 0) s_gt = s_ge = 0;
    __syncthreads();
    if (wid == 0) {
      const float* pb = hd + dst_rows[i] * D;
      float acc = 0.f;
      for (int c = lane; c < D; c += 32) acc = fmaf(fmaxf(s_a[c] + pb[c], 0.f), s_a[D + c], acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        const float p = 1.f / (1.f + expf(-(acc + bf[0])));
        s_pos = p;
        if (part == 0) pos_out[i] = p;
      }
    }
    __syncthreads();
    const float p = s_pos;
    int gt = 0, ge = 0;
    // four negatives per warp iteration: their row indices, then their rows, are fetched together (one chain of
    // two dependent latencies per FOUR candidates; per-candidate arithmetic and its order are unchanged)
    constexpr int kU = 4;
    for (int q0 = qlo + wid * kU; q0 < qhi; q0 += nw * kU) {
      const float* pb[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        pb[u] = q0 + u < qhi ? hd + neg_rows[(long long)i * Q + q0 + u] * D : nullptr;
      float acc[kU] = {};
      for (int c = lane; c < D; c += 32) {
        const float sa = s_a[c], sw = s_a[D + c];
#pragma unroll
        for (int u = 0; u < kU; ++u)
          if (pb[u]) acc[u] = fmaf(fmaxf(sa + pb[u][c], 0.f), sw, acc[u]);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (!pb[u]) break;      // warp-uniform
        const float v = 1.f / (1.f + expf(-(warp_sum(acc[u]) + bf[0])));
        gt += v > p;
        ge += v >= p;
        if (neg_out && lane == 0) neg_out[(long long)i * Q + q0 + u] = v;
      }
    }
    if (lane == 0) {
      atomicAdd(&s_gt, gt);
      atomicAdd(&s_ge, ge);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (parts == 1) {
        gt_out[i] = s_gt;
        ge_out[i] = s_ge;
      } else {            // zero-filled by the wrapper
        atomicAdd(&gt_out[i], s_gt);
        atomicAdd(&ge_out[i], s_ge);
      }
    }
  }
}

// Rows of the candidates of one evaluation batch inside the all-gathered table of projected root rows
// [P][2][Rm][D]: global root index g lives in the block of rank g % P at position g / P, i.e. row
// (g % P) * block_rows + g / P.  g = [src (B) | dst (B) | neg (B x Q, row-major)]; this rank scores the negative
// columns rank, rank + P, ... (Qr of them).
__global__ void dp_rows_kernel(const int64_t* __restrict__ g, int B, int Q, int rank, int P, int64_t block_rows,
                               int64_t* __restrict__ src_rows, int64_t* __restrict__ dst_rows,
                               int64_t* __restrict__ neg_rows) {
  pdl_wait();
  pdl_launch();
  const int Qr = Q > rank ? (Q - rank + P - 1) / P : 0;
  const long long total = 2ll * B + (long long)B * Qr;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total;
       x += (long long)gridDim.x * blockDim.x) {
    long long srcpos;
    int64_t* out;
    if (x < B) {
      srcpos = x; out = src_rows + x;
    } else if (x < 2ll * B) {
      srcpos = x; out = dst_rows + (x - B);
    } else {
      const long long y = x - 2ll * B;
      const long long i = y / Qr, q = y - i * Qr;
      srcpos = 2ll * B + i * Q + rank + q * P;
      out = neg_rows + y;
    }
    const int64_t v = g[srcpos];
    *out = (v % P) * block_rows + v / P;
  }
}

// out[i] = in[offset + i * stride] for the entries that exist, -1 beyond (ids < 0 are ignored by the
// marking kernels); *out_count = how many exist.  Picks one rank's share of a sorted root list.
__global__ void stride_select_kernel(const int64_t* __restrict__ in, DevCount n_in, int offset, int stride,
                                     int64_t* __restrict__ out, int out_cap, int32_t* __restrict__ out_count) {
  pdl_wait();
  pdl_launch();
  const int n = n_in.get();
  const int cnt = n > offset ? (n - offset + stride - 1) / stride : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < out_cap; i += gridDim.x * blockDim.x)
    out[i] = i < cnt ? in[offset + (long long)i * stride] : -1;
  if (blockIdx.x == 0 && threadIdx.x == 0 && out_count) *out_count = cnt < out_cap ? cnt : out_cap;
}

__global__ void scatter_add_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ rows,
                                        DevCount num, int D, float* __restrict__ dst) {
  pdl_wait();
  pdl_launch();
  const int n = num.get();
  const long long total = (long long)n * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / D), c = (int)(e - (long long)i * D);
    atomicAdd(&dst[rows[i] * D + c], src[e]);
  }
}

}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_relabel3(const int64_t* a, int32_t na, const int32_t* na_dev, int64_t* oa,
                     const int64_t* b, int32_t nb, const int32_t* nb_dev, int64_t* ob,
                     const int64_t* c, int32_t nc, const int32_t* nc_dev, int64_t* oc,
                     const int64_t* assoc, void* stream) {
  TGN_REQUIRE(na >= 0 && nb >= 0 && nc >= 0, "relabel3: negative count");
  if (na + nb + nc == 0) return TGN_OK;
  TGN_REQUIRE(assoc && (na == 0 || (a && oa)) && (nb == 0 || (b && ob)) && (nc == 0 || (c && oc)),
              "relabel3: NULL pointer");
  launch_k(relabel3_kernel, dim3(stride_grid((long long)na + nb + nc, 256)), dim3(256), 0, (cudaStream_t)stream, 
      a, DevCount{na_dev, na}, oa, b, DevCount{nb_dev, nb}, ob, c, DevCount{nc_dev, nc}, oc, assoc);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_edge_attr_ld(const int64_t* last_update_local, const int64_t* nbr_local,
                         const int64_t* t_edge, const float* msg, const int64_t* msg_rows,
                         int64_t num_events, int32_t num_edges, const int32_t* num_edges_dev, int32_t raw_dim,
                         int32_t time_dim, const float* time_w, const float* time_b, int32_t ld,
                         float* edge_attr, float* sin_out, float* rel_t, void* stream) {
  TGN_REQUIRE(num_events >= 0 && num_edges >= 0 && raw_dim >= 0 && time_dim >= 0 && ld >= raw_dim + time_dim && ld >= 1,
              "edge_attr_ld: bad sizes");
  if (num_edges == 0) return TGN_OK;
  TGN_REQUIRE(last_update_local && nbr_local && t_edge && (msg || raw_dim == 0) && edge_attr &&
                  (time_dim == 0 || (time_w && time_b)),
              "edge_attr_ld: NULL pointer");
  EdgeAttr2Args a;
  a.lu = last_update_local; a.nbr = nbr_local; a.t_edge = t_edge; a.msg = msg; a.msg_rows = msg_rows;
  a.edges = DevCount{num_edges_dev, num_edges}; a.De = raw_dim; a.Dt = time_dim; a.ld = ld;
  a.time_w = time_w; a.time_b = time_b; a.ea = edge_attr; a.sn = sin_out; a.rel = rel_t;
  a.num_events = num_events; a.err = dev_err_word();
  const bool small = num_edges <= kNumSMs * 8 * 8;      // one warp per edge still fits one wave
  const int grid = stride_grid((long long)ceil_div(num_edges, small ? 1 : 4) * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (sin_out && small) launch_k(edge_attr_ld_kernel<true, 1>, dim3(grid), dim3(256), 0, st, a);
  else if (sin_out) launch_k(edge_attr_ld_kernel<true, 4>, dim3(grid), dim3(256), 0, st, a);
  else if (small) launch_k(edge_attr_ld_kernel<false, 1>, dim3(grid), dim3(256), 0, st, a);
  else launch_k(edge_attr_ld_kernel<false, 4>, dim3(grid), dim3(256), 0, st, a);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_time_bwd_sin(const float* t, const int32_t* row_mask, int32_t num, const int32_t* num_dev,
                         const float* sin_vals, int32_t dim, const float* grad, int32_t ld_grad,
                         float* d_w, float* d_b, void* stream) {
  TGN_REQUIRE(num >= 0 && dim >= 1 && ld_grad >= dim, "time_bwd_sin: bad sizes");
  if (num == 0) return TGN_OK;
  TGN_REQUIRE(t && sin_vals && grad && d_w && d_b, "time_bwd_sin: NULL pointer");
  int gy = ceil_div(num, 128);
  if (gy > 64) gy = 64;
  launch_k(time_bwd_sin_kernel, dim3(dim3(ceil_div(dim, 32), gy)), dim3(256), 0, (cudaStream_t)stream, 
      t, row_mask, DevCount{num_dev, num}, sin_vals, dim, grad, ld_grad, d_w, d_b);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

static int32_t core_args(AttnCoreArgs& a, const float* proj, const int64_t* nbr_local,
                         const int32_t* row_ptr, const int64_t* centre_ids, int32_t num_centres,
                         const int32_t* num_centres_dev, int32_t heads, int32_t head_dim,
                         const float* ee, float dropout_p, uint64_t seed, const int64_t* seed_dev) {
  TGN_REQUIRE(num_centres >= 0 && heads >= 1 && heads <= kCMaxHeads && head_dim >= 1 &&
                  heads * head_dim <= 32 * kCMaxCH && (heads * head_dim) % 4 == 0,
              "attn_core: bad sizes (heads <= %d, heads*head_dim <= %d and a multiple of 4)",
              kCMaxHeads, 32 * kCMaxCH);
  TGN_REQUIRE(heads == 1 || heads == 2 || heads == 4 || heads == 8, "attn_core: heads must be 1, 2, 4 or 8");
  TGN_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "attn_core: dropout_p must be in [0,1)");
  TGN_REQUIRE(proj && nbr_local && row_ptr && ee, "attn_core: NULL pointer");
  a.proj = proj; a.nbr = nbr_local; a.row_ptr = row_ptr; a.centre_ids = centre_ids;
  a.centres = DevCount{num_centres_dev, num_centres}; a.H = heads; a.C = head_dim; a.ee = ee;
  a.dropout_p = dropout_p; a.seed = seed; a.seed_dev = seed_dev;
  a.out = nullptr; a.alpha = nullptr; a.d_out = nullptr; a.d_proj = nullptr; a.d_ee = nullptr;
  return TGN_OK;
}

int32_t tgn_attn_core_fwd(const float* proj, const int64_t* nbr_local, const int32_t* row_ptr,
                          const int64_t* centre_ids, int32_t num_centres,
                          const int32_t* num_centres_dev, int32_t heads, int32_t head_dim,
                          const float* ee, float dropout_p, uint64_t seed, const int64_t* seed_dev,
                          int32_t max_degree, float* out, float* alpha, void* stream) {
  if (num_centres == 0) return TGN_OK;
  AttnCoreArgs a;
  int32_t rc = core_args(a, proj, nbr_local, row_ptr, centre_ids, num_centres, num_centres_dev, heads,
                         head_dim, ee, dropout_p, seed, seed_dev);
  if (rc) return rc;
  TGN_REQUIRE(out && alpha, "attn_core_fwd: NULL output");
  a.out = out; a.alpha = alpha;
  const int grid = ceil_div(num_centres, kCoreWarps);
  cudaStream_t s = (cudaStream_t)stream;
  if (max_degree > 0 && max_degree <= kSmallDeg && heads * head_dim <= 128 && heads <= 4) {
    switch (heads) {
      case 1: launch_k(attn_core_fwd_small_kernel<1, kSmallDeg, 3>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
      case 2:
        if (max_degree <= 10 && num_centres > 8 * kNumSMs * kCoreWarps)   // occupancy only pays on multi-wave launches
          launch_k(attn_core_fwd_small_kernel<2, 10, 4>, dim3(min(grid, 4 * kNumSMs)), dim3(kCoreWarps * 32), 0, s, a);
        else
          launch_k(attn_core_fwd_small_kernel<2, kSmallDeg, 3>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a);
        break;
      default: launch_k(attn_core_fwd_small_kernel<4, kSmallDeg, 2>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
    }
    TGN_LAUNCH_CHECK();
    return TGN_OK;
  }
  switch (heads) {
    case 1: launch_k(attn_core_fwd_kernel<1>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
    case 2: launch_k(attn_core_fwd_kernel<2>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
    case 4: launch_k(attn_core_fwd_kernel<4>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
    default: launch_k(attn_core_fwd_kernel<8>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
  }
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_attn_core_bwd(const float* proj, const int64_t* nbr_local, const int32_t* row_ptr,
                          const int64_t* centre_ids, int32_t num_centres,
                          const int32_t* num_centres_dev, int32_t heads, int32_t head_dim,
                          const float* ee, const float* alpha, const float* d_out, float dropout_p,
                          uint64_t seed, const int64_t* seed_dev, int32_t max_degree, int32_t num_rows,
                          float* d_proj, float* d_ee, void* stream) {
  TGN_REQUIRE(num_rows >= 0, "attn_core_bwd: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  if (num_rows > 0) {
    TGN_REQUIRE(d_proj, "attn_core_bwd: d_proj is NULL");
    TGN_CUDA(cudaMemsetAsync(d_proj, 0, (size_t)num_rows * 4 * heads * head_dim * sizeof(float), s));
  }
  if (num_centres == 0) return TGN_OK;
  AttnCoreArgs a;
  int32_t rc = core_args(a, proj, nbr_local, row_ptr, centre_ids, num_centres, num_centres_dev, heads,
                         head_dim, ee, dropout_p, seed, seed_dev);
  if (rc) return rc;
  TGN_REQUIRE(alpha && d_out && d_proj && d_ee, "attn_core_bwd: NULL pointer");
  a.alpha = const_cast<float*>(alpha); a.d_out = d_out; a.d_proj = d_proj; a.d_ee = d_ee;
  const int grid = ceil_div(num_centres, kCoreWarps);
  if (max_degree > 0 && max_degree <= kSmallDeg && heads * head_dim <= 128 && heads <= 4) {
    switch (heads) {
      case 1: launch_k(attn_core_bwd_small_kernel<1>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
      case 2: launch_k(attn_core_bwd_small_kernel<2>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
      default: launch_k(attn_core_bwd_small_kernel<4>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
    }
    TGN_LAUNCH_CHECK();
    return TGN_OK;
  }
  switch (heads) {
    case 1: launch_k(attn_core_bwd_kernel<1>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
    case 2: launch_k(attn_core_bwd_kernel<2>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
    case 4: launch_k(attn_core_bwd_kernel<4>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
    default: launch_k(attn_core_bwd_kernel<8>, dim3(grid), dim3(kCoreWarps * 32), 0, s, a); break;
  }
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_dec_loss(const float* hs, const float* hd, const float* w_final, const float* b_final,
                     int32_t batch, int32_t dim, float* loss, float* logits, float* dh, float* dhs,
                     float* d_w_final, float* d_b_final, float* d_b_src, float* d_b_dst,
                     void* stream) {
  TGN_REQUIRE(batch >= 1 && dim >= 1, "dec_loss: bad sizes");
  TGN_REQUIRE(hs && hd && w_final && b_final && loss && dh && dhs && d_w_final && d_b_final &&
                  d_b_src && d_b_dst,
              "dec_loss: NULL pointer");
  int grid = ceil_div(batch, kDecWarps);
  if (grid > kNumSMs) grid = kNumSMs;
  launch_k(dec_loss_kernel, dim3(grid), dim3(kDecWarps * 32), (size_t)3 * dim * sizeof(float), (cudaStream_t)stream, 
      hs, hd, w_final, b_final, batch, dim, loss, logits, dh, dhs, d_w_final, d_b_final, d_b_src,
      d_b_dst);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int64_t tgn_dec_fused_smem_bytes(int32_t dim) {
  return ((int64_t)4 * dim * (dim + 1) + 2 * kDecVec * dim) * (int64_t)sizeof(float);
}

static int32_t dec_fused_launch(const float* emb, const int64_t* ids_local, int32_t batch, int32_t dim,
                                const float* w_src, const float* b_src, const float* w_dst, const float* b_dst,
                                const float* w_final, const float* b_final, float* loss, float* logits,
                                float* d_emb, float* d_w_src, float* d_b_src, float* d_w_dst, float* d_b_dst,
                                float* d_w_final, float* d_b_final, float* z_rows, float* g_rows,
                                const AttnCoreArgs* attn, const int64_t* ids_root, void* stream);

int32_t tgn_dec_fused(const float* emb, const int64_t* ids_local, int32_t batch, int32_t dim,
                      const float* w_src, const float* b_src, const float* w_dst, const float* b_dst,
                      const float* w_final, const float* b_final, float* loss, float* logits,
                      float* d_emb, float* d_w_src, float* d_b_src, float* d_w_dst, float* d_b_dst,
                      float* d_w_final, float* d_b_final, float* z_rows, float* g_rows, void* stream) {
  TGN_REQUIRE(emb, "dec_fused: NULL pointer");
  return dec_fused_launch(emb, ids_local, batch, dim, w_src, b_src, w_dst, b_dst, w_final, b_final, loss, logits, d_emb,
                          d_w_src, d_b_src, d_w_dst, d_b_dst, d_w_final, d_b_final, z_rows, g_rows, nullptr, nullptr,
                          stream);
}

int32_t tgn_dec_attn_fused(const float* proj, const int64_t* nbr_local, const int32_t* row_ptr,
                           const int64_t* centre_ids, int32_t num_centres, int32_t heads, int32_t head_dim,
                           const float* ee, float dropout_p, uint64_t seed, const int64_t* seed_dev,
                           int32_t max_degree, float* alpha, const int64_t* ids_root, const int64_t* ids_local,
                           int32_t batch, const float* w_src, const float* b_src, const float* w_dst,
                           const float* b_dst, const float* w_final, const float* b_final, float* loss,
                           float* logits, float* d_emb, float* d_b_src, float* d_b_dst, float* d_w_final,
                           float* d_b_final, float* z_rows, float* g_rows, void* stream) {
  TGN_REQUIRE(heads == 2 && max_degree >= 1 && max_degree <= 10 && heads * head_dim <= 128 && head_dim % 2 == 0,
              "dec_attn_fused: two heads, at most 10 edges per centre, heads * head_dim <= 128 (use tgn_attn_core_fwd "
              "+ tgn_dec_fused otherwise)");
  TGN_REQUIRE(alpha && ids_root && z_rows && g_rows, "dec_attn_fused: NULL pointer");
  AttnCoreArgs a;
  int32_t rc = core_args(a, proj, nbr_local, row_ptr, centre_ids, num_centres, nullptr, heads, head_dim, ee, dropout_p,
                         seed, seed_dev);
  if (rc) return rc;
  a.alpha = alpha;
  a.out = nullptr;
  return dec_fused_launch(nullptr, ids_local, batch, heads * head_dim, w_src, b_src, w_dst, b_dst, w_final, b_final,
                          loss, logits, d_emb, nullptr, d_b_src, nullptr, d_b_dst, d_w_final, d_b_final, z_rows,
                          g_rows, &a, ids_root, stream);
}

static int32_t dec_fused_launch(const float* emb, const int64_t* ids_local, int32_t batch, int32_t dim,
                      const float* w_src, const float* b_src, const float* w_dst, const float* b_dst,
                      const float* w_final, const float* b_final, float* loss, float* logits,
                                float* d_emb, float* d_w_src, float* d_b_src, float* d_w_dst, float* d_b_dst,
                                float* d_w_final, float* d_b_final, float* z_rows, float* g_rows,
                                const AttnCoreArgs* attn, const int64_t* ids_root, void* stream) {
  TGN_REQUIRE(batch >= 1 && dim >= 1, "dec_fused: bad sizes");
  const bool defer = z_rows != nullptr || g_rows != nullptr;
  const int64_t smem = defer ? ((int64_t)2 * dim * (dim + 1) + 2 * kDecVec * dim) * (int64_t)sizeof(float)
                             : tgn_dec_fused_smem_bytes(dim);
  TGN_REQUIRE(smem <= 215 * 1024 && dim <= 128,
              "dec_fused: dim %d does not fit shared memory (use the GEMM path)", dim);
  TGN_REQUIRE(ids_local && w_src && b_src && w_dst && b_dst && w_final && b_final && loss &&
                  d_emb && d_b_src && d_b_dst && d_w_final && d_b_final,
              "dec_fused: NULL pointer");
  TGN_REQUIRE(defer ? (z_rows && g_rows) : (d_w_src && d_w_dst),
              "dec_fused: give d_w_src/d_w_dst, or z_rows AND g_rows for deferred weight gradients");
  static int64_t attr_smem[3] = {0, 0, 0};
  const int variant = attn ? 2 : (defer ? 1 : 0);
  if (smem > attr_smem[variant]) {
    if (attn)
      TGN_CUDA(cudaFuncSetAttribute(dec_fused_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else if (defer)
      TGN_CUDA(cudaFuncSetAttribute(dec_fused_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      TGN_CUDA(cudaFuncSetAttribute(dec_fused_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[variant] = smem;
  }
  // one pass = kDecEv events; a CTA's fixed cost is staging the two weight matrices (and, without
  // deferred weight gradients, the accumulator flush)
  int grid = ceil_div(batch, kDecEv);
  if (grid > kNumSMs) grid = kNumSMs;
  AttnCoreArgs none;
  memset(&none, 0, sizeof(none));
  if (attn)
    launch_k(dec_fused_kernel<true, true>, dim3(grid), dim3(kDecThreads), (size_t)smem, (cudaStream_t)stream,
             emb, ids_local, batch, dim, w_src, b_src, w_dst, b_dst, w_final, b_final, loss, logits, d_emb,
             d_w_src, d_b_src, d_w_dst, d_b_dst, d_w_final, d_b_final, z_rows, g_rows, *attn, ids_root);
  else if (defer)
    launch_k(dec_fused_kernel<true, false>, dim3(grid), dim3(kDecThreads), (size_t)smem, (cudaStream_t)stream,
             emb, ids_local, batch, dim, w_src, b_src, w_dst, b_dst, w_final, b_final, loss, logits, d_emb,
             d_w_src, d_b_src, d_w_dst, d_b_dst, d_w_final, d_b_final, z_rows, g_rows, none, ids_root);
  else
    launch_k(dec_fused_kernel<false, false>, dim3(grid), dim3(kDecThreads), (size_t)smem, (cudaStream_t)stream,
             emb, ids_local, batch, dim, w_src, b_src, w_dst, b_dst, w_final, b_final, loss, logits, d_emb,
             d_w_src, d_b_src, d_w_dst, d_b_dst, d_w_final, d_b_final, z_rows, g_rows, none, ids_root);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_score_negs(const float* hs, const float* hd, const int64_t* src_rows,
                       const int64_t* dst_rows, const int64_t* neg_rows, int32_t num_pos,
                       int32_t num_neg, int32_t dim, const float* w_final, const float* b_final,
                       float* pos_out, float* neg_out, int32_t* gt_out, int32_t* ge_out,
                       void* stream) {
  TGN_REQUIRE(num_pos >= 0 && num_neg >= 0 && dim >= 1, "score_negs: bad sizes");
  if (num_pos == 0) return TGN_OK;
  TGN_REQUIRE(hs && hd && src_rows && dst_rows && (neg_rows || num_neg == 0) && w_final &&
                  b_final && pos_out && gt_out && ge_out,
              "score_negs: NULL pointer");
  int grid = num_pos < 8 * kNumSMs ? num_pos : 8 * kNumSMs;
  int parts = (num_neg + 127) / 128;
  parts = parts < 1 ? 1 : (parts > 8 ? 8 : parts);
  if (parts > 1) {
    TGN_CUDA(cudaMemsetAsync(gt_out, 0, (size_t)num_pos * sizeof(int32_t), (cudaStream_t)stream));
    TGN_CUDA(cudaMemsetAsync(ge_out, 0, (size_t)num_pos * sizeof(int32_t), (cudaStream_t)stream));
  }
  launch_k(score_negs_kernel, dim3(grid, parts), dim3(256), (size_t)2 * dim * sizeof(float), (cudaStream_t)stream, 
      hs, hd, src_rows, dst_rows, neg_rows, num_pos, num_neg, dim, w_final, b_final, pos_out,
      neg_out, gt_out, ge_out);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_dp_rows(const int64_t* global_index, int32_t num_pos, int32_t num_neg, int32_t rank, int32_t world,
                    int64_t block_rows, int64_t* src_rows, int64_t* dst_rows, int64_t* neg_rows, void* stream) {
  TGN_REQUIRE(num_pos >= 1 && num_neg >= 0 && world >= 1 && rank >= 0 && rank < world && block_rows >= 1,
              "dp_rows: bad sizes / rank");
  TGN_REQUIRE(global_index && src_rows && dst_rows && (neg_rows || num_neg <= rank), "dp_rows: NULL pointer");
  const int qr = num_neg > rank ? (num_neg - rank + world - 1) / world : 0;
  launch_k(dp_rows_kernel, dim3(stride_grid(2ll * num_pos + (long long)num_pos * qr, 256)), dim3(256), 0,
           (cudaStream_t)stream, global_index, num_pos, num_neg, rank, world, block_rows, src_rows, dst_rows, neg_rows);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_stride_select(const int64_t* in, int32_t num_in, const int32_t* num_in_dev, int32_t offset,
                          int32_t stride, int64_t* out, int32_t out_cap, int32_t* out_count_dev, void* stream) {
  TGN_REQUIRE(num_in >= 0 && offset >= 0 && stride >= 1 && out_cap >= 0, "stride_select: bad sizes");
  if (out_cap == 0) return TGN_OK;
  TGN_REQUIRE(in && out, "stride_select: NULL pointer");
  launch_k(stride_select_kernel, dim3(stride_grid(out_cap, 256)), dim3(256), 0, (cudaStream_t)stream, in,
           DevCount{num_in_dev, num_in}, offset, stride, out, out_cap, out_count_dev);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_scatter_add_rows(const float* src, const int64_t* rows, int32_t num,
                             const int32_t* num_dev, int32_t dim, float* dst, void* stream) {
  TGN_REQUIRE(num >= 0 && dim >= 1, "scatter_add_rows: bad sizes");
  if (num == 0) return TGN_OK;
  TGN_REQUIRE(src && rows && dst, "scatter_add_rows: NULL pointer");
  launch_k(scatter_add_rows_kernel, dim3(stride_grid((long long)num * dim, 256)), dim3(256), 0, (cudaStream_t)stream, 
      src, rows, DevCount{num_dev, num}, dim, dst);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
