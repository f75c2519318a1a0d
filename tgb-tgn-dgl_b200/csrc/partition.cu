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
  for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < bound; s += gridDim.x * wpb) {
    float* gn = g_n + (long long)s * Dm;
    float* go = g_o + (long long)s * Dm;
    int64_t n = -1, other = -1;
    if (s < S) {
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

}  // extern "C"
