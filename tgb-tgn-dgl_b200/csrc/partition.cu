// Owner-partitioned node memory (reference state:
// modules/memory_module.py:80-83 `memory` [N,D], `last_update` [N]).
//
// Node n is owned by rank n % P and lives at local row n / P of that rank's shard.  The
// integer metadata of the step (neighbour ring, message store, batch) is replicated -- it is
// small and every rank derives it from the same inputs -- so each rank can name the memory rows
// the step needs without asking anybody:
//     rows of n_id[s]                       (h of the GRU, first block of the message)
//     rows of other[s]                      (second block: the other endpoint of the stored
//                                            event that Last aggregation picks for n_id[s])
// tgn_part_gather writes the rows THIS rank owns into a zero-filled staging buffer; summing the
// staging buffers of all ranks (one NCCL all-reduce over NVLink/NVSwitch, fixed size, no index
// exchange, capturable in a CUDA graph) assembles all rows on every rank, and
// tgn_msg_build_gathered consumes them.  tgn_memory_scatter_owned is the owner-side write-back
// (memory_module.py:147-150).
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

template <typename T>
__global__ void part_gather_kernel(tgn_msgstore st, const int64_t* __restrict__ n_id, DevCount num,
                                   int bound, const float* __restrict__ mem_loc,
                                   const int64_t* __restrict__ lu_loc, int Dm, int rank, int world,
                                   float* __restrict__ g_n, float* __restrict__ g_o,
                                   int64_t* __restrict__ g_lu, int64_t* __restrict__ other_out) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int S = num.get();
  const T* ev_t = reinterpret_cast<const T*>(st.ev_t);
  for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < bound; s += gridDim.x * wpb) {
    float* gn = g_n + (long long)s * Dm;
    float* go = g_o + (long long)s * Dm;
    int64_t n = -1, other = -1;
    if (s < S) {
      n = n_id[s];
      if (n >= 0 && n < st.num_nodes) {
        const int sc = st.s_cnt[n], dc = st.d_cnt[n];
        if (sc + dc > 0) {
          // same choice as msg_build_kernel (Last aggregation, s-store wins ties)
          const int es = sc > 0 ? st.s_last[n] : -1, ed = dc > 0 ? st.d_last[n] : -1;
          const T ts_ = es >= 0 ? ev_t[es] : (T)0, td_ = ed >= 0 ? ev_t[ed] : (T)0;
          const bool pick_s = es >= 0 && (ed < 0 || ts_ >= td_);
          other = pick_s ? st.ev_dst[es] : st.ev_src[ed];
        }
      } else {
        n = -1;
      }
    }
    const bool own_n = n >= 0 && (n % world) == rank;
    const bool own_o = other >= 0 && (other % world) == rank;
    const float* pn = mem_loc + (own_n ? (n / world) : 0) * Dm;
    const float* po = mem_loc + (own_o ? (other / world) : 0) * Dm;
    for (int c = lane; c < Dm; c += 32) {
      gn[c] = own_n ? pn[c] : 0.f;
      go[c] = own_o ? po[c] : 0.f;
    }
    if (lane == 0) {
      g_lu[s] = own_n ? lu_loc[n / world] : 0;
      if (other_out) other_out[s] = other;
    }
  }
}

// Peer-memory variant: every rank's shard is mapped into this process (NVLink / NVSwitch peer access,
// torch symmetric memory), so the rows are read straight out of their owners' HBM -- no staging buffer,
// no all-reduce.  The caller brackets the launch with two rank barriers (all scatters of the previous
// step done / all gathers done before anyone scatters again).
constexpr int kMaxPeers = 16;
struct PeerShards {
  const float* mem[kMaxPeers];
  const int64_t* lu[kMaxPeers];
};

template <typename T>
__global__ void part_gather_p2p_kernel(tgn_msgstore st, const int64_t* __restrict__ n_id, DevCount num,
                                       int bound, PeerShards peers, int Dm, int world,
                                       float* __restrict__ g_n, float* __restrict__ g_o,
                                       int64_t* __restrict__ g_lu) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int S = num.get();
  const T* ev_t = reinterpret_cast<const T*>(st.ev_t);
  // (rows beyond the live count are not touched: every consumer clamps to the same device count)
  for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < S && s < bound; s += gridDim.x * wpb) {
    float* gn = g_n + (long long)s * Dm;
    float* go = g_o + (long long)s * Dm;
    int64_t n = -1, other = -1;
    {
      n = n_id[s];
      if (n >= 0 && n < st.num_nodes) {
        const int sc = st.s_cnt[n], dc = st.d_cnt[n];
        if (sc + dc > 0) {
          const int es = sc > 0 ? st.s_last[n] : -1, ed = dc > 0 ? st.d_last[n] : -1;
          const T ts_ = es >= 0 ? ev_t[es] : (T)0, td_ = ed >= 0 ? ev_t[ed] : (T)0;
          const bool pick_s = es >= 0 && (ed < 0 || ts_ >= td_);
          other = pick_s ? st.ev_dst[es] : st.ev_src[ed];
        }
      } else {
        n = -1;
      }
    }
    const float* pn = n >= 0 ? peers.mem[n % world] + (n / world) * Dm : nullptr;
    const float* po = other >= 0 ? peers.mem[other % world] + (other / world) * Dm : nullptr;
    for (int c = lane; c < Dm; c += 32) {
      gn[c] = pn ? pn[c] : 0.f;
      go[c] = po ? po[c] : 0.f;
    }
    if (lane == 0) g_lu[s] = n >= 0 ? peers.lu[n % world][n / world] : 0;
  }
}

template <typename T>
__global__ void memory_scatter_owned_kernel(const int64_t* __restrict__ n_id, DevCount num,
                                            const float* __restrict__ new_mem,
                                            const T* __restrict__ new_lu,
                                            const int64_t* __restrict__ src_rows, int D, int rank,
                                            int world, float* __restrict__ mem_loc,
                                            int64_t* __restrict__ lu_loc) {
  pdl_wait();
  pdl_launch();
  const int S = num.get();
  const long long total = (long long)S * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(e / D), j = (int)(e - (long long)s * D);
    const int64_t n = n_id[s];
    if (n < 0 || (n % world) != rank) continue;
    const long long r = src_rows ? src_rows[s] : s;
    const int64_t loc = n / world;
    mem_loc[loc * D + j] = new_mem[r * D + j];
    if (j == 0 && new_lu) lu_loc[loc] = (int64_t)new_lu[r];
  }
}


// ---------------------------------------------------------------------------
// Owner-side compute (north_star "node memory split by node-id owner"): of the rows n_id of a
// step, rank r runs the message build and the GRU only for the nodes it owns (n % P == r) and PUBLISHES
// the resulting rows h' / last_update' into every rank's row table through the peer mapping; the GRU
// backward likewise runs on the owned rows only, and the partial weight gradients are summed inside the
// optimiser kernel straight out of the peers' gradient buffers.
// ---------------------------------------------------------------------------
// own_nodes / own_pos: the owned entries of n_id and their positions, in position order (single CTA:
// a block scan per 1024 ids keeps the order deterministic, so reductions over the owned rows are too).
__global__ void __launch_bounds__(1024)
    part_select_owned_kernel(const int64_t* __restrict__ n_id, DevCount num, int rank, int world,
                             int64_t* __restrict__ own_nodes, int64_t* __restrict__ own_pos, int cap,
                             int32_t* __restrict__ own_count, int32_t* err) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_warp[32];
  __shared__ int s_base, s_total;
  const int S = num.get();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int s0 = 0; s0 < S; s0 += 1024) {
    const int s = s0 + tid;
    int64_t n = -1;
    if (s < S) n = n_id[s];
    const bool own = n >= 0 && (int)(n % world) == rank;
    const unsigned ball = __ballot_sync(0xffffffffu, own);
    if (lane == 0) s_warp[wid] = __popc(ball);
    __syncthreads();
    // prefix over the 32 warp counts by the first warp, once, instead of 32 shared reads per thread
    if (wid == 0) {
      const int c = s_warp[lane];
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      s_warp[lane] = incl - c;
      if (lane == 31) s_total = incl;
    }
    __syncthreads();
    const int before = s_warp[wid], total = s_total;
    if (own) {
      const int pos = s_base + before + __popc(ball & lanemask_lt());
      if (pos < cap) {
        own_nodes[pos] = n;
        own_pos[pos] = s;
      }
    }
    __syncthreads();
    if (tid == 0) s_base += total;
    __syncthreads();
  }
  if (tid == 0) {
    if (s_base > cap) flag_dev_err(err, TGN_DEVERR_OWNER_CAP);   // the caller's bound on one rank's share was too small
    *own_count = s_base < cap ? s_base : cap;
  }
}

struct PeerTables {
  float* rows[kMaxPeers];
  int64_t* lu[kMaxPeers];
};

__global__ void part_publish_kernel(const float* __restrict__ rows, const int64_t* __restrict__ lu,
                                    const int64_t* __restrict__ own_pos, DevCount num, int D, int world,
                                    PeerTables peers) {
  pdl_wait();
  pdl_launch();
  const int S = num.get();
  const int D4 = D >> 2;   // D % 4 == 0 (engine requirement): 128-bit stores over NVLink
  const long long total = (long long)S * D4 * world;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total;
       x += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(x / ((long long)S * D4));
    const long long y = x - (long long)p * S * D4;
    const int i = (int)(y / D4), c = (int)(y - (long long)i * D4);
    const int64_t pos = own_pos[i];
    reinterpret_cast<float4*>(peers.rows[p] + pos * D)[c] = reinterpret_cast<const float4*>(rows + (long long)i * D)[c];
    if (c == 0) peers.lu[p][pos] = lu[i];
  }
}

// Writes one block of `n16` 16-byte words into the SAME place of every rank's table (peer stores over
// NVLink / NVSwitch): the all-gather of the sharded evaluation (every rank's block lands at its own offset of
// every peer's table) without a collective call.
struct PeerDst {
  int4* dst[kMaxPeers];
};
__global__ void peer_bcast_kernel(const int4* __restrict__ src, long long n16, int world, PeerDst peers) {
  pdl_wait();
  pdl_launch();
  const long long total = n16 * world;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total;
       x += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(x / n16);
    const long long i = x - (long long)p * n16;
    peers.dst[p][i] = src[i];
  }
}

struct PeerGrads {
  const float* part[kMaxPeers];
};

// torch.optim.Adam step whose gradient is rep[i] + sum_p part[p][i]: `rep` = the replicated part of the
// gradient as rank 0 computed it (every rank reads the SAME copy, so the weight replicas stay bit-identical),
// part[p] = rank p's partial sums over the rows it owns.  Same arithmetic as adam_kernel (dense.cu).
__global__ void adam_peers_kernel(float* __restrict__ p, const float* __restrict__ rep, PeerGrads pg, int world,
                                  long long n_part, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                  float b1, float b2, float eps, float* __restrict__ step_dev,
                                  int64_t* __restrict__ step_ctr, const float* __restrict__ loss_acc,
                                  float* __restrict__ loss_out, unsigned* __restrict__ done_ctr) {
  pdl_wait();
  pdl_launch();
  const float step = *step_dev + 1.f;
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = rep[i];
    if (i < n_part)      // only the memory-path parameters (a prefix of the flat layout) have partial sums
      for (int r = 0; r < world; ++r) gi += pg.part[r][i];
    const float mi = m[i] + (1.f - b1) * (gi - m[i]);
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] -= step_size * (mi / denom);
  }
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(done_ctr, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    *step_dev += 1.f;
    if (step_ctr) *step_ctr += 1;
    if (loss_acc && loss_out) *loss_out = *loss_acc;
    *done_ctr = 0u;
  }
}

}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_part_gather(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                        const int32_t* num_dev, const float* memory_local,
                        const int64_t* last_update_local, int32_t memory_dim, int32_t rank,
                        int32_t world, float* rows_n, float* rows_other, int64_t* lu_out,
                        int64_t* other_out, void* stream) {
  TGN_REQUIRE(st && st->num_nodes > 0, "part_gather: store is NULL");
  TGN_REQUIRE(num >= 0 && memory_dim >= 1 && world >= 1 && rank >= 0 && rank < world,
              "part_gather: bad sizes / rank");
  if (num == 0) return TGN_OK;
  TGN_REQUIRE(n_id && memory_local && last_update_local && rows_n && rows_other && lu_out,
              "part_gather: NULL pointer");
  const int grid = stride_grid((long long)num * 32, 256);
  cudaStream_t s = (cudaStream_t)stream;
  DevCount c{num_dev, num};
  if (st->t_is_float)
    launch_k(part_gather_kernel<float>, dim3(grid), dim3(256), 0, s, *st, n_id, c, num, memory_local,
             last_update_local, memory_dim, rank, world, rows_n, rows_other, lu_out, other_out);
  else
    launch_k(part_gather_kernel<int64_t>, dim3(grid), dim3(256), 0, s, *st, n_id, c, num, memory_local,
             last_update_local, memory_dim, rank, world, rows_n, rows_other, lu_out, other_out);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_part_gather_p2p(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                            const int32_t* num_dev, const void* const* peer_memory,
                            const void* const* peer_last_update, int32_t memory_dim, int32_t world,
                            float* rows_n, float* rows_other, int64_t* lu_out, void* stream) {
  TGN_REQUIRE(st && st->num_nodes > 0, "part_gather_p2p: store is NULL");
  TGN_REQUIRE(num >= 0 && memory_dim >= 1 && world >= 1 && world <= kMaxPeers,
              "part_gather_p2p: bad sizes (world <= %d)", kMaxPeers);
  if (num == 0) return TGN_OK;
  TGN_REQUIRE(n_id && peer_memory && peer_last_update && rows_n && rows_other && lu_out,
              "part_gather_p2p: NULL pointer");
  PeerShards peers;
  for (int r = 0; r < kMaxPeers; ++r) {
    peers.mem[r] = r < world ? (const float*)peer_memory[r] : nullptr;
    peers.lu[r] = r < world ? (const int64_t*)peer_last_update[r] : nullptr;
    TGN_REQUIRE(r >= world || (peers.mem[r] && peers.lu[r]), "part_gather_p2p: peer %d has no mapping", r);
  }
  const int grid = stride_grid((long long)num * 32, 256);
  cudaStream_t s = (cudaStream_t)stream;
  DevCount c{num_dev, num};
  if (st->t_is_float)
    launch_k(part_gather_p2p_kernel<float>, dim3(grid), dim3(256), 0, s, *st, n_id, c, num, peers, memory_dim,
             world, rows_n, rows_other, lu_out);
  else
    launch_k(part_gather_p2p_kernel<int64_t>, dim3(grid), dim3(256), 0, s, *st, n_id, c, num, peers,
             memory_dim, world, rows_n, rows_other, lu_out);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_memory_scatter_owned(const int64_t* n_id, int32_t num, const int32_t* num_dev,
                                 const float* new_mem, const void* new_lu, int32_t lu_is_float,
                                 const int64_t* src_rows, int32_t dim, int32_t rank, int32_t world,
                                 float* memory_local, int64_t* last_update_local, void* stream) {
  TGN_REQUIRE(num >= 0 && dim >= 1 && world >= 1 && rank >= 0 && rank < world,
              "memory_scatter_owned: bad sizes / rank");
  if (num == 0) return TGN_OK;
  TGN_REQUIRE(n_id && new_mem && memory_local && (last_update_local || !new_lu),
              "memory_scatter_owned: NULL pointer");
  DevCount c{num_dev, num};
  const int grid = stride_grid((long long)num * dim, 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (lu_is_float)
    launch_k(memory_scatter_owned_kernel<float>, dim3(grid), dim3(256), 0, s, n_id, c, new_mem,
             (const float*)new_lu, src_rows, dim, rank, world, memory_local, last_update_local);
  else
    launch_k(memory_scatter_owned_kernel<int64_t>, dim3(grid), dim3(256), 0, s, n_id, c, new_mem,
             (const int64_t*)new_lu, src_rows, dim, rank, world, memory_local, last_update_local);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_part_select_owned(const int64_t* n_id, int32_t num, const int32_t* num_dev, int32_t rank,
                              int32_t world, int64_t* own_nodes, int64_t* own_pos, int32_t own_cap,
                              int32_t* own_count_dev, void* stream) {
  TGN_REQUIRE(num >= 0 && own_cap >= 0 && world >= 1 && rank >= 0 && rank < world, "part_select_owned: bad sizes / rank");
  TGN_REQUIRE(own_count_dev && (num == 0 || (n_id && own_nodes && own_pos)), "part_select_owned: NULL pointer");
  launch_k(part_select_owned_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, n_id, DevCount{num_dev, num},
           rank, world, own_nodes, own_pos, own_cap, own_count_dev, dev_err_word());
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_part_publish(const float* rows, const int64_t* last_update, const int64_t* own_pos, int32_t num,
                         const int32_t* num_dev, int32_t dim, void* const* peer_rows, void* const* peer_last_update,
                         int32_t world, void* stream) {
  TGN_REQUIRE(num >= 0 && dim >= 4 && dim % 4 == 0 && world >= 1 && world <= kMaxPeers,
              "part_publish: bad sizes (dim %% 4 == 0, world <= %d)", kMaxPeers);
  if (num == 0) return TGN_OK;
  TGN_REQUIRE(rows && last_update && own_pos && peer_rows && peer_last_update, "part_publish: NULL pointer");
  PeerTables t;
  for (int r = 0; r < kMaxPeers; ++r) {
    t.rows[r] = r < world ? (float*)peer_rows[r] : nullptr;
    t.lu[r] = r < world ? (int64_t*)peer_last_update[r] : nullptr;
    TGN_REQUIRE(r >= world || (t.rows[r] && t.lu[r]), "part_publish: peer %d has no mapping", r);
  }
  launch_k(part_publish_kernel, dim3(stride_grid((long long)num * (dim / 4) * world, 256)), dim3(256), 0,
           (cudaStream_t)stream, rows, last_update, own_pos, DevCount{num_dev, num}, dim, world, t);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_peer_bcast(const void* src, int64_t nbytes, void* const* peer_tables, int64_t table_offset_bytes,
                       int32_t world, void* stream) {
  TGN_REQUIRE(nbytes >= 0 && table_offset_bytes >= 0 && world >= 1 && world <= kMaxPeers,
              "peer_bcast: bad sizes (world <= %d)", kMaxPeers);
  if (nbytes == 0) return TGN_OK;
  TGN_REQUIRE(src && peer_tables, "peer_bcast: NULL pointer");
  TGN_REQUIRE(nbytes % 16 == 0 && table_offset_bytes % 16 == 0 && ((uintptr_t)src & 15) == 0,
              "peer_bcast: block size, table offset and source must be multiples of 16 bytes");
  PeerDst d;
  for (int r = 0; r < kMaxPeers; ++r) {
    d.dst[r] = r < world ? reinterpret_cast<int4*>((char*)peer_tables[r] + table_offset_bytes) : nullptr;
    TGN_REQUIRE(r >= world || (peer_tables[r] && ((uintptr_t)d.dst[r] & 15) == 0), "peer_bcast: peer %d has no (aligned) mapping", r);
  }
  launch_k(peer_bcast_kernel, dim3(stride_grid(nbytes / 16 * world, 256)), dim3(256), 0, (cudaStream_t)stream,
           (const int4*)src, (long long)(nbytes / 16), world, d);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_adam_finish_peers(float* params, const float* grads_replicated, const void* const* peer_grads_partial,
                              int32_t world, int64_t partial_count, float* exp_avg, float* exp_avg_sq, int64_t count, float lr,
                              float beta1, float beta2, float eps, float* step_dev, int64_t* step_counter,
                              const float* loss_acc, float* loss_out, uint32_t* done_counter, void* stream) {
  TGN_REQUIRE(count >= 1 && partial_count >= 0 && partial_count <= count && world >= 1 && world <= kMaxPeers &&
                  step_dev && done_counter,
              "adam_finish_peers: bad arguments");
  TGN_REQUIRE(params && grads_replicated && peer_grads_partial && exp_avg && exp_avg_sq, "adam_finish_peers: NULL pointer");
  PeerGrads pg;
  for (int r = 0; r < kMaxPeers; ++r) {
    pg.part[r] = r < world ? (const float*)peer_grads_partial[r] : nullptr;
    TGN_REQUIRE(r >= world || pg.part[r], "adam_finish_peers: peer %d has no mapping", r);
  }
  launch_k(adam_peers_kernel, dim3(stride_grid(count, 256)), dim3(256), 0, (cudaStream_t)stream, params,
           grads_replicated, pg, world, (long long)partial_count, exp_avg, exp_avg_sq, (long long)count, lr, beta1, beta2, eps, step_dev,
           step_counter, loss_acc, loss_out, done_counter);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
