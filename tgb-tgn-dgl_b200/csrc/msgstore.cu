// Tensorised TGN message store + fused message build.
//
// The reference keeps, per node, a Python dict entry holding the node's events
// of the last batch it appeared in (modules/memory_module.py:140-145,180-191),
// once keyed by source (msg_s_store) and once by destination (msg_d_store).
// Here the events of every update_state() call are appended to a device log
// (ev_src/ev_dst/ev_t/ev_msg); the batch is sorted by (node, position) in
// shared memory and every touched node records
//     start,count : its run inside the sorted permutation log (s_perm/d_perm)
//     last        : the log id of its latest event, first-wins on equal t
// which is exactly the information _compute_msg + LastAggregator/MeanAggregator
// consume.  Entries are overwritten per batch like the dict entries are.
//
// Ordering note: the reference orders a node's events by torch.sort (unstable
// for small CPU tensors).  The store fixes the order to batch position
// (stable); this only matters for equal timestamps inside one node's run.
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

__device__ __forceinline__ void block_bitonic_sort64(unsigned long long* s, int P) {
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

template <typename T>
__device__ void store_direction(const int64_t* __restrict__ key_nodes, const T* __restrict__ t,
                                int B, int P, int64_t base, int64_t num_nodes,
                                unsigned long long* s_key, int32_t* __restrict__ perm,
                                int32_t* __restrict__ start, int32_t* __restrict__ cnt,
                                int32_t* __restrict__ last) {
  for (int j = threadIdx.x; j < P; j += blockDim.x)
    s_key[j] = j < B ? (((unsigned long long)key_nodes[j] << 16) | (unsigned)j) : ~0ull;
  __syncthreads();
  block_bitonic_sort64(s_key, P);
  for (int q = threadIdx.x; q < B; q += blockDim.x) {
    const unsigned long long key = s_key[q];
    const int64_t node = (int64_t)(key >> 16);
    perm[base + q] = (int32_t)(base + (int)(key & 0xffffu));
    const bool run_end = (q == B - 1) || ((int64_t)(s_key[q + 1] >> 16) != node);
    if (!run_end || node < 0 || node >= num_nodes) continue;
    // walk the run backwards: first-wins maximum of t
    int qs = q;
    int best = (int)(key & 0xffffu);
    T best_t = t[best];
    while (qs > 0 && (int64_t)(s_key[qs - 1] >> 16) == node) {
      --qs;
      const int j = (int)(s_key[qs] & 0xffffu);
      const T tj = t[j];
      if (tj >= best_t) {  // earlier position wins ties
        best_t = tj;
        best = j;
      }
    }
    start[node] = (int32_t)(base + qs);
    cnt[node] = q - qs + 1;
    last[node] = (int32_t)(base + best);
  }
  __syncthreads();
}

// grid = 2 + copy CTAs: CTA 0 indexes the batch by source, CTA 1 by destination (each with its own
// shared-memory sort), the others append the batch to the event log.  All of them read the log position
// at the start; it is advanced by a one-thread kernel behind this one (stream order).
template <typename T>
__global__ void __launch_bounds__(1024, 1)
    msgstore_update_kernel(tgn_msgstore st, const int64_t* __restrict__ src,
                           const int64_t* __restrict__ dst, const T* __restrict__ t,
                           const float* __restrict__ raw, int B, int P, int64_t base_host,
                           const int64_t* __restrict__ base_dev, int32_t* err) {
  pdl_wait();
  pdl_launch();
  extern __shared__ unsigned long long s_key[];
  const int64_t base = base_dev ? *base_dev : base_host;
  if (base + B > st.capacity) {  // caller sizes the log; never write out of bounds -- but say so
    if (threadIdx.x == 0 && blockIdx.x == 0) flag_dev_err(err, TGN_DEVERR_LOG_OVERFLOW);
    return;
  }
  if (blockIdx.x == 0) {
    store_direction<T>(src, t, B, P, base, st.num_nodes, s_key, st.s_perm, st.s_start, st.s_cnt, st.s_last);
    return;
  }
  if (blockIdx.x == 1) {
    store_direction<T>(dst, t, B, P, base, st.num_nodes, s_key, st.d_perm, st.d_start, st.d_cnt, st.d_last);
    return;
  }
  const int nc = gridDim.x - 2, c = blockIdx.x - 2;
  T* ev_t = reinterpret_cast<T*>(st.ev_t);
  for (int i = c * blockDim.x + threadIdx.x; i < B; i += nc * blockDim.x) {
    st.ev_src[base + i] = src[i];
    st.ev_dst[base + i] = dst[i];
    ev_t[base + i] = t[i];
  }
  const long long nraw = (long long)B * st.raw_dim;
  float* dm = st.ev_msg + base * st.raw_dim;
  for (long long i = (long long)c * blockDim.x + threadIdx.x; i < nraw; i += (long long)nc * blockDim.x) dm[i] = raw[i];
}

__global__ void msgstore_advance_kernel(int64_t* p, int64_t by, int64_t capacity) {
  pdl_wait();
  pdl_launch();
  if (*p + by <= capacity) *p += by;   // an overflowing batch was dropped (and flagged) by the update kernel
}

__global__ void msgstore_reset_kernel(tgn_msgstore st) {
  pdl_wait();
  pdl_launch();
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < st.num_nodes;
       n += (int64_t)gridDim.x * blockDim.x) {
    st.s_cnt[n] = 0;
    st.d_cnt[n] = 0;
    st.s_last[n] = -1;
    st.d_last[n] = -1;
    st.s_start[n] = 0;
    st.d_start[n] = 0;
  }
}

// exclusive scan of cnt[n_id[i]] with the chained-scan workspace (tile = 1024)
__global__ void __launch_bounds__(1024)
    msgstore_count_kernel(const int32_t* __restrict__ cnt, const int64_t* __restrict__ n_id,
                          int num, int64_t num_nodes, int32_t* __restrict__ offsets,
                          unsigned long long* __restrict__ ws) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_warp[32];
  __shared__ long long s_prefix;
  const int ntiles = (num + 1023) / 1024;
  const int tile = lookback_take_tile(ws);
  if (tile >= ntiles) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i = tile * 1024 + tid;
  int c = 0;
  if (i < num) {
    const int64_t n = n_id[i];
    if (n >= 0 && n < num_nodes) c = cnt[n];
  }
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  int wbase = 0, total = 0;
  for (int w = 0; w < 32; ++w) {
    if (w < wid) wbase += s_warp[w];
    total += s_warp[w];
  }
  if (tid == 0) {
    long long pre = lookback_prefix(ws, tile, total);
    s_prefix = pre;
    if (tile == ntiles - 1) offsets[num] = (int32_t)(pre + total);
  }
  __syncthreads();
  if (i < num) offsets[i] = (int32_t)(s_prefix + wbase + incl - c);
}

// one warp per node: copy the node's stored tuples in store order
template <typename T>
__global__ void msgstore_gather_kernel(tgn_msgstore st, const int64_t* __restrict__ n_id, int num,
                                       int dir, const int32_t* __restrict__ offsets,
                                       int64_t* __restrict__ out_src,
                                       int64_t* __restrict__ out_dst, T* __restrict__ out_t,
                                       float* __restrict__ out_raw) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const T* ev_t = reinterpret_cast<const T*>(st.ev_t);
  for (int s = blockIdx.x * warps_per_block + (threadIdx.x >> 5); s < num;
       s += gridDim.x * warps_per_block) {
    const int64_t n = n_id[s];
    if (n < 0 || n >= st.num_nodes) continue;
    const int c = dir == 0 ? st.s_cnt[n] : st.d_cnt[n];
    const int b = dir == 0 ? st.s_start[n] : st.d_start[n];
    const int32_t* perm = dir == 0 ? st.s_perm : st.d_perm;
    const int o = offsets[s];
    for (int j = 0; j < c; ++j) {
      const int e = perm[b + j];
      if (lane == 0) {
        // stored tuple is (src,dst) for the s-store and (dst,src) for the d-store
        out_src[o + j] = dir == 0 ? st.ev_src[e] : st.ev_dst[e];
        out_dst[o + j] = dir == 0 ? st.ev_dst[e] : st.ev_src[e];
        out_t[o + j] = ev_t[e];
      }
      for (int c2 = lane; c2 < st.raw_dim; c2 += 32)
        out_raw[(long long)(o + j) * st.raw_dim + c2] = st.ev_msg[(long long)e * st.raw_dim + c2];
    }
  }
}

// ---------------------------------------------------------------------------
// Fused message build: one warp per node, lanes over the message columns.
// x row layout = IdentityMessage (msg_func.py:17-18): [mem[n], mem[other], raw, t_enc]
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float rel_time(T t, int64_t lu);
template <>
__device__ __forceinline__ float rel_time<int64_t>(int64_t t, int64_t lu) {
  return (float)(t - lu);  // int64 subtraction, then .to(float) (memory_module.py:202-203)
}
template <>
__device__ __forceinline__ float rel_time<float>(float t, int64_t lu) {
  return t - (float)lu;  // float32 - int64 promotes to float32
}

// Owner-partitioned node memory (csrc/partition.cu): node n lives on rank n % world at row n / world; the
// shards of all ranks are mapped (symmetric memory), so a row is read straight out of its owner's HBM.
constexpr int kMsgMaxPeers = 16;
struct MsgPeers {
  const float* mem[kMsgMaxPeers];
  const int64_t* lu[kMsgMaxPeers];
  int world;      // 0: not partitioned (the dense `memory` / `last_update` tables are used)
};

template <typename T, bool kPeers>
__global__ void msg_build_kernel(tgn_msgstore st, const int64_t* __restrict__ n_id, DevCount num,
                                 int agg_mode, const float* __restrict__ memory,
                                 const int64_t* __restrict__ last_update, int Dm,
                                 const float* __restrict__ time_w,
                                 const float* __restrict__ time_b, int Dt, float* __restrict__ x,
                                 int ldx, float* __restrict__ h_out, float* __restrict__ sin_out,
                                 T* __restrict__ lu_out, int32_t* __restrict__ sel_ev,
                                 float* __restrict__ sel_dt, const float* __restrict__ gath_n,
                                 const float* __restrict__ gath_o,
                                 const int64_t* __restrict__ gath_lu, MsgPeers peers) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31;
  // memory row / last_update of node id: the dense table, or the owner's shard over the peer mapping
  auto row_of = [&](int64_t id) -> const float* {
    return kPeers ? peers.mem[id % peers.world] + (id / peers.world) * Dm : memory + id * Dm;
  };
  const int warps_per_block = blockDim.x >> 5;
  const int S = num.get();
  const int De = st.raw_dim;
  const int W = 2 * Dm + De + Dt;
  const T* ev_t = reinterpret_cast<const T*>(st.ev_t);
  for (int s = blockIdx.x * warps_per_block + (threadIdx.x >> 5); s < S;
       s += gridDim.x * warps_per_block) {
    const int64_t n = n_id[s];
    float* xr = x + (long long)s * ldx;
    const bool ok = n >= 0 && n < st.num_nodes;
    const int sc = ok ? st.s_cnt[n] : 0, dc = ok ? st.d_cnt[n] : 0;
    for (int c = W + lane; c < ldx; c += 32) xr[c] = 0.f;  // row padding (TMA-aligned stride)
    // owner-partitioned memory: rows / last_update come pre-assembled per row s (partition.cu)
    const float* mn_src = gath_n ? gath_n + (long long)s * Dm : row_of(ok ? n : 0);
    if (h_out)
      for (int c = lane; c < Dm; c += 32) h_out[(long long)s * Dm + c] = ok ? mn_src[c] : 0.f;
    if (sc + dc == 0) {
      for (int c = lane; c < W; c += 32) xr[c] = 0.f;
      if (lane == 0) {
        lu_out[s] = (T)0;
        if (sel_ev) sel_ev[s] = -1;
        if (sel_dt) sel_dt[s] = 0.f;
      }
      continue;
    }
    const int64_t lu = gath_lu ? gath_lu[s]
                               : (kPeers ? peers.lu[n % peers.world][n / peers.world] : last_update[n]);
    const float* mn = mn_src;
    if (agg_mode == TGN_AGG_LAST) {
      const int es = sc > 0 ? st.s_last[n] : -1, ed = dc > 0 ? st.d_last[n] : -1;
      T ts_ = es >= 0 ? ev_t[es] : (T)0, td_ = ed >= 0 ? ev_t[ed] : (T)0;
      // messages of the s-store precede those of the d-store (memory_module.py:166-168):
      // on equal t the s-store event wins
      const bool pick_s = es >= 0 && (ed < 0 || ts_ >= td_);
      const int e = pick_s ? es : ed;
      const T te = pick_s ? ts_ : td_;
      const int64_t other = pick_s ? st.ev_dst[e] : st.ev_src[e];
      const float dt = rel_time<T>(te, lu);
      const float* mo = gath_o ? gath_o + (long long)s * Dm : row_of(other);
      const float* rw = st.ev_msg + (long long)e * De;
      for (int c = lane; c < Dm; c += 32) {
        xr[c] = mn[c];
        xr[Dm + c] = mo[c];
      }
      for (int c = lane; c < De; c += 32) xr[2 * Dm + c] = rw[c];
      for (int c = lane; c < Dt; c += 32) {
        float sv, cv;
        sincos_fr(__fmaf_rn(dt, time_w[c], time_b[c]), &sv, &cv);
        xr[2 * Dm + De + c] = cv;
        if (sin_out) sin_out[(long long)s * Dt + c] = sv;  // for tgn_time_bwd_sin
      }
      if (lane == 0) {
        lu_out[s] = te;
        if (sel_ev) sel_ev[s] = e;
        if (sel_dt) sel_dt[s] = dt;
      }
    } else {
      // mean over all stored messages, s-store first (order of msg_agg.py:26 input)
      const float inv = 1.f / (float)(sc + dc);
      T tmax = (T)0;
      bool first = true;
      for (int c0 = 0; c0 < W; c0 += 32) {
        const int c = c0 + lane;
        float acc = 0.f;
        for (int dir = 0; dir < 2; ++dir) {
          const int cnt = dir == 0 ? sc : dc;
          const int b = dir == 0 ? st.s_start[n] : st.d_start[n];
          const int32_t* perm = dir == 0 ? st.s_perm : st.d_perm;
          for (int j = 0; j < cnt; ++j) {
            const int e = perm[b + j];
            const T te = ev_t[e];
            if (c0 == 0) {
              if (first || te > tmax) tmax = te;
              first = false;
            }
            if (c < W) {
              float v;
              if (c < Dm) v = mn[c];
              else if (c < 2 * Dm) {
                const int64_t other = dir == 0 ? st.ev_dst[e] : st.ev_src[e];
                v = row_of(other)[c - Dm];
              } else if (c < 2 * Dm + De) v = st.ev_msg[(long long)e * De + (c - 2 * Dm)];
              else {
                const int cc = c - 2 * Dm - De;
                v = cos_fr(__fmaf_rn(rel_time<T>(te, lu), time_w[cc], time_b[cc]));
              }
              acc += v;
            }
          }
        }
        if (c < W) xr[c] = acc * inv;
      }
      if (lane == 0) {
        lu_out[s] = tmax;
        if (sel_ev) sel_ev[s] = -1;
        if (sel_dt) sel_dt[s] = 0.f;
      }
    }
  }
}

}  // namespace tgn

using namespace tgn;

extern "C" {

static int32_t check_store(const tgn_msgstore* st, const char* who) {
  TGN_REQUIRE(st, "%s: store is NULL", who);
  TGN_REQUIRE(st->num_nodes > 0 && st->capacity > 0 && st->raw_dim >= 0, "%s: bad store sizes",
              who);
  TGN_REQUIRE(st->capacity < (1ll << 31), "%s: log capacity must fit int32", who);
  TGN_REQUIRE(st->ev_src && st->ev_dst && st->ev_t && (st->ev_msg || st->raw_dim == 0) &&
                  st->s_perm && st->d_perm && st->s_start && st->s_cnt && st->s_last &&
                  st->d_start && st->d_cnt && st->d_last,
              "%s: store has NULL arrays", who);
  return TGN_OK;
}

int32_t tgn_msgstore_update(const tgn_msgstore* st, const int64_t* src, const int64_t* dst,
                            const void* t, const float* raw_msg, int32_t batch, int64_t base,
                            int64_t* base_dev, void* stream) {
  int32_t rc = check_store(st, "msgstore_update");
  if (rc) return rc;
  TGN_REQUIRE(batch >= 0 && batch <= TGN_SORT_MAX, "msgstore_update: batch %d exceeds %d", batch,
              TGN_SORT_MAX);
  if (batch == 0) return TGN_OK;
  TGN_REQUIRE(src && dst && t && (raw_msg || st->raw_dim == 0), "msgstore_update: NULL input");
  TGN_REQUIRE(base_dev || (base >= 0 && base + batch <= st->capacity),
              "msgstore_update: log overflow (base %lld + %d > capacity %lld)", (long long)base,
              batch, (long long)st->capacity);
  int P = 2;
  while (P < batch) P <<= 1;
  static unsigned long long mi = 0, mf = 0;
  TGN_CUDA(smem_optin(msgstore_update_kernel<int64_t>, TGN_SORT_MAX * 8, mi));
  TGN_CUDA(smem_optin(msgstore_update_kernel<float>, TGN_SORT_MAX * 8, mf));
  cudaStream_t s = (cudaStream_t)stream;
  const long long copy_items = (long long)batch * (st->raw_dim > 3 ? st->raw_dim : 3);
  int copy_ctas = (int)((copy_items + 8191) / 8192);
  copy_ctas = copy_ctas < 1 ? 1 : (copy_ctas > 30 ? 30 : copy_ctas);
  if (st->t_is_float)
    launch_k(msgstore_update_kernel<float>, dim3(2 + copy_ctas), dim3(1024), (size_t)P * 8, s,
        *st, src, dst, (const float*)t, raw_msg, batch, P, base, (const int64_t*)base_dev, dev_err_word());
  else
    launch_k(msgstore_update_kernel<int64_t>, dim3(2 + copy_ctas), dim3(1024), (size_t)P * 8, s,
        *st, src, dst, (const int64_t*)t, raw_msg, batch, P, base, (const int64_t*)base_dev, dev_err_word());
  TGN_LAUNCH_CHECK();
  if (base_dev) {
    launch_k(msgstore_advance_kernel, dim3(1), dim3(1), 0, s, base_dev, (int64_t)batch, (int64_t)st->capacity);
    TGN_LAUNCH_CHECK();
  }
  return TGN_OK;
}

int32_t tgn_msgstore_reset(const tgn_msgstore* st, void* stream) {
  int32_t rc = check_store(st, "msgstore_reset");
  if (rc) return rc;
  launch_k(msgstore_reset_kernel, dim3(stride_grid(st->num_nodes, 256)), dim3(256), 0, (cudaStream_t)stream, *st);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int64_t tgn_msgstore_count_ws_bytes(int32_t num) {
  if (num < 0) return 0;
  int ntiles = num > 0 ? (num + 1023) / 1024 : 1;
  return (int64_t)(ntiles + 1) * 8;
}

int32_t tgn_msgstore_count(const tgn_msgstore* st, const int64_t* n_id, int32_t num, int32_t dir,
                           int32_t* offsets, void* ws, void* stream) {
  int32_t rc = check_store(st, "msgstore_count");
  if (rc) return rc;
  TGN_REQUIRE(num >= 0 && offsets && ws && (dir == 0 || dir == 1), "msgstore_count: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (num == 0) {
    TGN_CUDA(cudaMemsetAsync(offsets, 0, 4, s));
    return TGN_OK;
  }
  TGN_REQUIRE(n_id, "msgstore_count: n_id is NULL");
  const int ntiles = (num + 1023) / 1024;
  TGN_CUDA(cudaMemsetAsync(ws, 0, (size_t)(ntiles + 1) * 8, s));
  launch_k(msgstore_count_kernel, dim3(ntiles), dim3(1024), 0, s, dir == 0 ? st->s_cnt : st->d_cnt, n_id, num,
                                                st->num_nodes, offsets,
                                                (unsigned long long*)ws);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_msgstore_gather(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                            int32_t dir, const int32_t* offsets, int64_t* out_src,
                            int64_t* out_dst, void* out_t, float* out_raw, void* stream) {
  int32_t rc = check_store(st, "msgstore_gather");
  if (rc) return rc;
  TGN_REQUIRE(num >= 0 && (dir == 0 || dir == 1), "msgstore_gather: bad arguments");
  if (num == 0) return TGN_OK;
  TGN_REQUIRE(n_id && offsets && out_src && out_dst && out_t && (out_raw || st->raw_dim == 0),
              "msgstore_gather: NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = stride_grid((long long)num * 32, 256);
  if (st->t_is_float)
    launch_k(msgstore_gather_kernel<float>, dim3(grid), dim3(256), 0, s, *st, n_id, num, dir, offsets, out_src,
                                                       out_dst, (float*)out_t, out_raw);
  else
    launch_k(msgstore_gather_kernel<int64_t>, dim3(grid), dim3(256), 0, s, *st, n_id, num, dir, offsets, out_src,
                                                         out_dst, (int64_t*)out_t, out_raw);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

static int32_t msg_build_impl(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                              const int32_t* num_dev, int32_t agg_mode, const float* memory,
                              const int64_t* last_update, int32_t memory_dim, const float* time_w,
                              const float* time_b, int32_t time_dim, float* x, int32_t ldx,
                              float* h_out, float* sin_out, void* lu_out, int32_t* sel_ev,
                              float* sel_dt, const float* gath_n, const float* gath_o,
                              const int64_t* gath_lu, void* stream, const MsgPeers* peers_in = nullptr) {
  MsgPeers peers;
  memset(&peers, 0, sizeof(peers));
  if (peers_in) peers = *peers_in;
  int32_t rc = check_store(st, "msg_build");
  if (rc) return rc;
  TGN_REQUIRE(num >= 0 && memory_dim >= 1 && time_dim >= 0, "msg_build: bad sizes");
  TGN_REQUIRE(agg_mode == TGN_AGG_LAST || agg_mode == TGN_AGG_MEAN, "msg_build: bad agg_mode");
  TGN_REQUIRE(ldx >= 2 * memory_dim + st->raw_dim + time_dim, "msg_build: ldx smaller than the message width");
  TGN_REQUIRE(!sin_out || agg_mode == TGN_AGG_LAST, "msg_build: sin_out needs the last aggregator");
  if (num == 0) return TGN_OK;
  TGN_REQUIRE(n_id && memory && last_update && x && lu_out && (time_dim == 0 || (time_w && time_b)),
              "msg_build: NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  DevCount c{num_dev, num};
  const int grid = stride_grid((long long)num * 32, 256);
#define TGN_MB_LAUNCH(T, PEERS)                                                                              \
  launch_k(msg_build_kernel<T, PEERS>, dim3(grid), dim3(256), 0, s, *st, n_id, c, agg_mode, memory, last_update, \
           memory_dim, time_w, time_b, time_dim, x, ldx, h_out, sin_out, (T*)lu_out, sel_ev, sel_dt, gath_n,     \
           gath_o, gath_lu, peers)
  if (peers.world > 0) {
    if (st->t_is_float) TGN_MB_LAUNCH(float, true);
    else TGN_MB_LAUNCH(int64_t, true);
  } else {
    if (st->t_is_float) TGN_MB_LAUNCH(float, false);
    else TGN_MB_LAUNCH(int64_t, false);
  }
#undef TGN_MB_LAUNCH
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_msg_build_ld(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                         const int32_t* num_dev, int32_t agg_mode, const float* memory,
                         const int64_t* last_update, int32_t memory_dim, const float* time_w,
                         const float* time_b, int32_t time_dim, float* x, int32_t ldx, float* h_out,
                         float* sin_out, void* lu_out, int32_t* sel_ev, float* sel_dt, void* stream) {
  return msg_build_impl(st, n_id, num, num_dev, agg_mode, memory, last_update, memory_dim, time_w, time_b,
                        time_dim, x, ldx, h_out, sin_out, lu_out, sel_ev, sel_dt, nullptr, nullptr,
                        nullptr, stream);
}

int32_t tgn_msg_build_gathered(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                               const int32_t* num_dev, const float* rows_n, const float* rows_other,
                               const int64_t* last_update_rows, int32_t memory_dim,
                               const float* time_w, const float* time_b, int32_t time_dim, float* x,
                               int32_t ldx, float* h_out, float* sin_out, void* lu_out,
                               int32_t* sel_ev, float* sel_dt, void* stream) {
  TGN_REQUIRE(rows_n && rows_other && last_update_rows, "msg_build_gathered: NULL gathered rows");
  // the dense table arguments are unused in gathered mode; pass the row buffers to satisfy the checks
  return msg_build_impl(st, n_id, num, num_dev, TGN_AGG_LAST, rows_n, last_update_rows, memory_dim,
                        time_w, time_b, time_dim, x, ldx, h_out, sin_out, lu_out, sel_ev, sel_dt, rows_n,
                        rows_other, last_update_rows, stream);
}

int32_t tgn_msg_build_p2p(const tgn_msgstore* st, const int64_t* n_id, int32_t num, const int32_t* num_dev,
                          const void* const* peer_memory, const void* const* peer_last_update, int32_t world,
                          int32_t memory_dim, const float* time_w, const float* time_b, int32_t time_dim, float* x,
                          int32_t ldx, float* h_out, float* sin_out, void* lu_out, int32_t* sel_ev, float* sel_dt,
                          void* stream) {
  TGN_REQUIRE(peer_memory && peer_last_update && world >= 1 && world <= kMsgMaxPeers,
              "msg_build_p2p: peer tables / world (<= %d)", kMsgMaxPeers);
  MsgPeers peers;
  memset(&peers, 0, sizeof(peers));
  peers.world = world;
  for (int r = 0; r < world; ++r) {
    peers.mem[r] = (const float*)peer_memory[r];
    peers.lu[r] = (const int64_t*)peer_last_update[r];
    TGN_REQUIRE(peers.mem[r] && peers.lu[r], "msg_build_p2p: peer %d has no mapping", r);
  }
  // (the dense table arguments are unused in this mode; rank 0's shard satisfies the NULL checks)
  return msg_build_impl(st, n_id, num, num_dev, TGN_AGG_LAST, peers.mem[0], peers.lu[0], memory_dim, time_w, time_b,
                        time_dim, x, ldx, h_out, sin_out, lu_out, sel_ev, sel_dt, nullptr, nullptr, nullptr, stream,
                        &peers);
}

int32_t tgn_msg_build(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                      const int32_t* num_dev, int32_t agg_mode, const float* memory,
                      const int64_t* last_update, int32_t memory_dim, const float* time_w,
                      const float* time_b, int32_t time_dim, float* x, void* lu_out,
                      int32_t* sel_ev, float* sel_dt, void* stream) {
  TGN_REQUIRE(st, "msg_build: store is NULL");
  return tgn_msg_build_ld(st, n_id, num, num_dev, agg_mode, memory, last_update, memory_dim, time_w,
                          time_b, time_dim, x, 2 * memory_dim + st->raw_dim + time_dim, nullptr,
                          nullptr, lu_out, sel_ev, sel_dt, stream);
}

}  // extern "C"
