/*
 * tgn_b200.h -- C-ABI of the B200 (sm_100a) TGN hot path.
 *
 * This is the drop-in boundary for the per-batch hot path of
 * cseduashraful/tgb-tgn-dgl (sampler -> message aggregation -> GRU memory
 * update -> temporal-attention embedding).  The reference has no FFI of its own
 * (it is pure Python over torch / torch_scatter / torch_geometric / TGL's
 * sampler_core); each entry point below names the reference call it replaces
 * as file:line relative to the reference tree.  INTEGRATION.md shows the ctypes
 * stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every function returns int32_t: 0 = TGN_OK, -1 = TGN_EINVAL,
 *     -2 = TGN_ECUDA; tgn_last_error() returns a thread-local message.
 *   - all data pointers are DEVICE pointers owned by the caller; nothing is
 *     allocated and the host is never synchronised inside a call.
 *   - variable-length outputs are written into caller-sized buffers and their
 *     element count is left in a device int32_t.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous and
 *     may be captured into CUDA graphs.
 *   - "count_dev" arguments (nullable) let a size be read from device memory:
 *     when non-NULL the accompanying host count is the upper bound used for
 *     the launch geometry and the kernel clamps to *count_dev.
 *   - workspaces (`ws`) are caller-owned scratch; required sizes come from the
 *     matching *_ws_bytes() helper.
 */
#ifndef TGN_B200_H_
#define TGN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGN_OK 0
#define TGN_EINVAL (-1)
#define TGN_ECUDA (-2)

#define TGN_ABI_VERSION 1

/* aggregation modes (modules/msg_agg.py:15 LastAggregator, :24 MeanAggregator) */
#define TGN_AGG_LAST 0
#define TGN_AGG_MEAN 1
/* t-CSR sampling strategies (config/TGN.yml:5 `strategy`) */
#define TGN_SAMPLE_RECENT 0
#define TGN_SAMPLE_UNIFORM 1
/* memory updater cells (modules/memory_module.py:71-74) */
#define TGN_CELL_GRU 0
#define TGN_CELL_RNN 1

int32_t tgn_abi_version(void);
const char* tgn_last_error(void);
/* Contract violations detected ON THE DEVICE by kernels that may run inside a captured graph (where no
 * return code reaches the caller): bit 0 = message-store log overflow (the batch was dropped instead of
 * written out of bounds, replaces the silent return the reference's unbounded Python dict never needs,
 * modules/memory_module.py:180-191), bit 1 = an event id outside the resident event arrays was asked for
 * (`data.msg[e_id]`, epoch_utils.py:224), bit 2 = batch above a kernel's sort capacity.  The word lives in
 * mapped pinned host memory: reading it is a plain load, valid after the stream has been synchronised.
 * reset != 0 clears it. */
#define TGN_DEVERR_LOG_OVERFLOW 1
#define TGN_DEVERR_EVENT_RANGE 2
#define TGN_DEVERR_SORT_CAP 4
#define TGN_DEVERR_OWNER_CAP 8   /* partitioned memory: one rank owns more rows of a step than its buffers hold */
int32_t tgn_device_errors(int32_t reset);
/* Programmatic dependent launch: kernels are launched with the programmatic-stream-serialization
 * attribute so that the launch latency and prologue of each kernel overlap the tail of its
 * predecessor on the stream (every kernel begins with griddepcontrol.wait, so results do not
 * change).  Enabled by default; returns the previous setting. */
int32_t tgn_set_pdl(int32_t enabled);

/* cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDefault, stream): the host <-> device staging copies of the
 * end-to-end path (pinned batch in, loss word out) without a framework call per copy. */
int32_t tgn_memcpy_async(void* dst, const void* src, int64_t nbytes, void* stream);

/* ------------------------------------------------------------------------- *
 * Sorted-unique + relabel by bitmap ranking.
 * Replaces torch.cat([...]).unique() + `_assoc[n_id] = arange` at
 * neighbor_loader.py:46-47, memory_module.py:129,153, epoch_utils.py:215.
 * The bitmap is two-level (bit per node, bit per 1024-node group); it must be
 * all-zero before the first mark and is left all-zero by tgn_unique_rank.
 * ------------------------------------------------------------------------- */
int64_t tgn_bitmap_bytes(int64_t num_nodes);
int32_t tgn_unique_mark(const int64_t* ids, int32_t count, const int32_t* count_dev,
                        int64_t num_nodes, void* bitmap, void* stream);
/* out_ids[0..*out_count) ascending; assoc[id] = rank (assoc nullable).  With
 * keep_marks != 0 the bitmap is left as is (more ids can be marked and ranked
 * again), otherwise it is cleared. */
int32_t tgn_unique_rank(void* bitmap, int64_t num_nodes, int64_t* out_ids, int32_t out_cap,
                        int64_t* assoc, int32_t* out_count, int32_t keep_marks, void* stream);
/* tgn_unique_mark(ids) + tgn_unique_rank in one single-CTA launch (small id lists) */
int32_t tgn_unique_mark_rank(const int64_t* ids, int32_t count, void* bitmap, int64_t num_nodes,
                             int64_t* out_ids, int32_t out_cap, int64_t* assoc, int32_t* out_count,
                             int32_t keep_marks, void* stream);
/* out[i] = assoc[ids[i]]   (neighbor_loader.py:48) */
int32_t tgn_relabel(const int64_t* ids, int32_t count, const int32_t* count_dev,
                    const int64_t* assoc, int64_t* out, void* stream);

/* ------------------------------------------------------------------------- *
 * LastNeighborLoader ring (neighbor_loader.py:15-109).
 * State: neighbors int64 [N,K], e_id int64 [N,K] (-1 = empty), t float [N,K].
 * ------------------------------------------------------------------------- */
int64_t tgn_nbr_lookup_ws_bytes(int32_t num_roots, int32_t size_k);
/* neighbor_loader.py:27-42: gather the K slots of every root, drop e_id < 0,
 * keep root-order x slot-order.  Writes GLOBAL ids; root_off[r] = first edge of
 * root r (root_off has num_roots+1 entries); *out_count = #edges.  When bitmap
 * is non-NULL every emitted neighbour is also marked in it (for the union at
 * neighbor_loader.py:46). */
int32_t tgn_nbr_lookup(const int64_t* n_id, int32_t num_roots, const int32_t* num_roots_dev,
                       int32_t size_k, int64_t num_nodes, const int64_t* neighbors,
                       const int64_t* e_id, const float* t, int64_t* out_nbr,
                       int64_t* out_centre, int64_t* out_eid, float* out_t, int32_t* root_off,
                       int32_t* out_count, void* bitmap, void* ws, void* stream);
/* neighbor_loader.py:52-104: append B events in both directions and keep the K
 * largest e_id (and, independently, the K largest t) per touched node.  e_id of
 * event i is cur_e_id + i; if cur_e_id_dev != NULL the base is read from it and
 * the kernel advances it by B.  2*B <= TGN_SORT_MAX. */
int32_t tgn_nbr_insert(const int64_t* src, const int64_t* dst, const float* t, int32_t batch,
                       int64_t cur_e_id, int64_t* cur_e_id_dev, int32_t size_k,
                       int64_t num_nodes, int64_t* neighbors, int64_t* e_id, float* t_state,
                       void* stream);
#define TGN_SORT_MAX 8192

/* ------------------------------------------------------------------------- *
 * TGL sampler_core.ParallelSampler over t-CSR (README.md:1-5; t-CSR file keys
 * indptr/indices/ts/eid at utils.py:73; parameters config/TGN.yml:1-9).
 * For every root (n, t): candidates are the row entries with
 *   t_lo <= ts < t_hi,  t_hi = t + offset,  t_lo = (duration > 0 ? t_hi - duration : -inf)
 * recent : the last min(k, #cand) of them, most recent first
 * uniform: all of them if #cand <= k, else k draws with replacement
 *          (Philox-4x32-10 keyed by seed, counter = (root index, draw)).
 * Outputs are compact and ordered by root: nbr/eid/ts/dts[*out_count],
 * col = root index, root_off[num_roots+1].
 * Skip index (optional, built once per graph): coarse[b] = ts[16 b],
 * tgn_tcsr_index_len(nnz) floats.  With it the per-root search reads the row's
 * slice of the index plus ONE 64-byte block of the row instead of log2(deg)
 * scattered sectors; results are identical.  coarse == NULL -> plain binary
 * search (nnz ignored).  ts must be 16-byte aligned when coarse is given.
 * ------------------------------------------------------------------------- */
int64_t tgn_tcsr_sample_ws_bytes(int32_t num_roots);
int64_t tgn_tcsr_index_len(int64_t nnz);
int32_t tgn_tcsr_build_index(const float* ts, int64_t nnz, float* coarse, void* stream);
int32_t tgn_tcsr_sample(const int32_t* indptr, const int32_t* indices, const int32_t* eid,
                        const float* ts, const float* coarse, int64_t nnz, int32_t num_nodes, const int32_t* root_nodes,
                        const float* root_ts, int32_t num_roots, int32_t k, int32_t strategy,
                        float offset, float duration, uint64_t seed, int32_t* out_nbr,
                        int32_t* out_col, int32_t* out_eid, float* out_ts, float* out_dts,
                        int32_t* root_off, int32_t* out_count, void* ws, void* stream);

/* ------------------------------------------------------------------------- *
 * t-CSR construction (TGL gen_graph.py / `python tgb_gen_graph.py --data <name>`,
 * reference README.md:4-5; output keys indptr/indices/ts/eid, utils.py:73; the
 * generator script is absent from the reference tree).
 * Directed entries: (src->dst, eid e) and, with add_reverse, (dst->src, eid e).
 * Rows are sorted by (ts, eid), ts = float32(t[e]); a self-loop's forward entry
 * precedes its reverse entry.  t is int64 (t_is_float=0) or float32.
 * t_sorted != 0 states that t is non-decreasing in e (only the node digits are
 * sorted then).  *bad_flag becomes 1 if an endpoint is outside [0, num_nodes).
 * indptr[num_nodes+1]; indices/eid/ts[(add_reverse ? 2 : 1) * num_events].
 * ------------------------------------------------------------------------- */
int64_t tgn_tcsr_build_ws_bytes(int64_t num_events, int32_t add_reverse);
int32_t tgn_tcsr_build(const int64_t* src, const int64_t* dst, const void* t, int32_t t_is_float,
                       int64_t num_events, int32_t num_nodes, int32_t add_reverse,
                       int32_t t_sorted, int32_t* indptr, int32_t* indices, int32_t* eid, float* ts,
                       int32_t* bad_flag, void* ws, void* stream);

/* ------------------------------------------------------------------------- *
 * Message aggregators on materialised messages.
 * tgn_agg_last  = LastAggregator.forward   (modules/msg_agg.py:15-21):
 *   argmax[s] = first i with index[i]==s and maximal t[i]  (M if none)
 *   out[s,:]  = msg[argmax[s],:] or 0.
 * tgn_agg_mean  = MeanAggregator.forward   (modules/msg_agg.py:24-26).
 * t is int64 (t_is_float=0) or float32 (t_is_float=1).
 * ws: 16*S bytes (last) / tgn_agg_mean_ws_bytes(M, S) bytes (mean).
 * ------------------------------------------------------------------------- */
int64_t tgn_agg_mean_ws_bytes(int32_t num_msgs, int32_t dim_size);
int32_t tgn_agg_last(const float* msg, const int64_t* index, const void* t, int32_t t_is_float,
                     int32_t num_msgs, int32_t dim_size, int32_t width, float* out,
                     int64_t* argmax, void* ws, void* stream);
int32_t tgn_agg_mean(const float* msg, const int64_t* index, int32_t num_msgs,
                     int32_t dim_size, int32_t width, float* out, void* ws, void* stream);

/* ------------------------------------------------------------------------- *
 * Tensorised per-node message store (replaces the Python dict at
 * modules/memory_module.py:140-145,180-191).  Events are appended to a log;
 * per node and per direction the store keeps the run of its events of the
 * LAST batch it appeared in (start,count into a sorted permutation log) and
 * the first-wins latest event of that run.
 * ------------------------------------------------------------------------- */
typedef struct tgn_msgstore {
  int64_t num_nodes;
  int64_t capacity;   /* log capacity in events */
  int32_t raw_dim;    /* D_e */
  int32_t t_is_float; /* dtype of ev_t: 0 int64, 1 float32 */
  int64_t* ev_src;    /* [capacity] */
  int64_t* ev_dst;    /* [capacity] */
  void* ev_t;         /* [capacity] int64 or float32 */
  float* ev_msg;      /* [capacity, raw_dim] */
  int32_t* s_perm;    /* [capacity] log ids sorted by (src, position) per batch */
  int32_t* d_perm;    /* [capacity] same, by dst */
  int32_t* s_start;   /* [N] offset into s_perm */
  int32_t* s_cnt;     /* [N] */
  int32_t* s_last;    /* [N] log id of latest (first-wins) event, -1 if none */
  int32_t* d_start;
  int32_t* d_cnt;
  int32_t* d_last;
} tgn_msgstore;

/* memory_module.py:133-134 (_update_msg_store for both directions).
 * Appends the batch at log position `base` (host value, or read from
 * base_dev which is then advanced by batch).  batch <= TGN_SORT_MAX. */
int32_t tgn_msgstore_update(const tgn_msgstore* st, const int64_t* src, const int64_t* dst,
                            const void* t, const float* raw_msg, int32_t batch, int64_t base,
                            int64_t* base_dev, void* stream);
/* memory_module.py:140-145 (_reset_message_store) */
int32_t tgn_msgstore_reset(const tgn_msgstore* st, void* stream);

/* Materialise the messages of nodes n_id exactly as _compute_msg does
 * (memory_module.py:193-207) for direction dir (0 = s-store, 1 = d-store):
 * first a counting pass, then the fill (stored tuples in store order: for the
 * d-store the stored "src" is the event's dst).  The module-level API feeds
 * these to the user's message module; the fused path below never materialises
 * messages. */
int32_t tgn_msgstore_count(const tgn_msgstore* st, const int64_t* n_id, int32_t num, int32_t dir,
                           int32_t* offsets /* [num+1] exclusive */, void* ws, void* stream);
int64_t tgn_msgstore_count_ws_bytes(int32_t num);
int32_t tgn_msgstore_gather(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                            int32_t dir, const int32_t* offsets, int64_t* out_src /* [M] */,
                            int64_t* out_dst, void* out_t /* [M] dtype of ev_t */,
                            float* out_raw /* [M,raw_dim] */, void* stream);

/* ------------------------------------------------------------------------- *
 * Fused message build (gather + IdentityMessage concat + TimeEncoder + Last /
 * Mean aggregation) for nodes n_id -- memory_module.py:152-169,176 with
 * msg_func.py:17-18 and the TimeEncoder contract (cos(w*dt+b)).
 *   x[s,:]        = aggregated message [mem[n], mem[other], raw, cos(w*(t-lu[n])+b)]
 *                   (zero row if the node has no stored event)
 *   lu_out[s]     = max stored t (0 if none), in the dtype of the stored t
 *   sel_ev[s]     = chosen log id (last mode) or -1
 *   sel_dt[s]     = t - last_update[n] of the chosen event (for backward)
 * ------------------------------------------------------------------------- */
int32_t tgn_msg_build(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                      const int32_t* num_dev, int32_t agg_mode, const float* memory,
                      const int64_t* last_update, int32_t memory_dim, const float* time_w,
                      const float* time_b, int32_t time_dim, float* x,
                      void* lu_out /* [num] dtype of ev_t */, int32_t* sel_ev, float* sel_dt,
                      void* stream);

/* tgn_msg_build with a caller-chosen row stride ldx >= message width (padding columns are
 * zeroed; a stride that is a multiple of 4 makes x a TMA operand) and, when h_out is given,
 * the gather h_out[s,:] = memory[n_id[s],:] (memory_module.py:172) in the same pass;
 * sin_out [S,time_dim] (nullable, last aggregator) keeps sin(w*dt+b) of the chosen event for
 * tgn_time_bwd_sin. */
int32_t tgn_msg_build_ld(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                         const int32_t* num_dev, int32_t agg_mode, const float* memory,
                         const int64_t* last_update, int32_t memory_dim, const float* time_w,
                         const float* time_b, int32_t time_dim, float* x, int32_t ldx, float* h_out,
                         float* sin_out, void* lu_out, int32_t* sel_ev, float* sel_dt, void* stream);

/* ------------------------------------------------------------------------- *
 * Owner-partitioned node memory (multi-GPU; memory_module.py:80-83 state split by node id).
 * Node n lives on rank n % world at local row n / world.  tgn_part_gather writes, for every row
 * s of n_id (all `num` rows; rows >= *num_dev are zero-filled), the memory row / last_update of
 * n_id[s] and the memory row of the other endpoint of the event Last aggregation picks for it --
 * but only where THIS rank owns them, zeros elsewhere -- so that a sum over ranks (one
 * all-reduce) assembles every row on every rank.  tgn_msg_build_gathered is tgn_msg_build_ld on
 * those assembled rows; tgn_memory_scatter_owned is the owner-side tgn_memory_scatter.
 * ------------------------------------------------------------------------- */
int32_t tgn_part_gather(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                        const int32_t* num_dev, const float* memory_local,
                        const int64_t* last_update_local, int32_t memory_dim, int32_t rank,
                        int32_t world, float* rows_n, float* rows_other, int64_t* lu_out,
                        int64_t* other_out, void* stream);
/* Peer-memory variant of tgn_part_gather: peer_memory / peer_last_update are HOST arrays of `world`
 * device pointers to every rank's shard, mapped into this process (NVLink peer access, e.g. torch
 * symmetric memory); every row is read from its owner's HBM directly, so no all-reduce follows.
 * The caller separates it from the scatters of the previous and of the current step with rank
 * barriers.  world <= 16. */
int32_t tgn_part_gather_p2p(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                            const int32_t* num_dev, const void* const* peer_memory,
                            const void* const* peer_last_update, int32_t memory_dim, int32_t world,
                            float* rows_n, float* rows_other, int64_t* lu_out, void* stream);
int32_t tgn_msg_build_gathered(const tgn_msgstore* st, const int64_t* n_id, int32_t num,
                               const int32_t* num_dev, const float* rows_n, const float* rows_other,
                               const int64_t* last_update_rows, int32_t memory_dim,
                               const float* time_w, const float* time_b, int32_t time_dim, float* x,
                               int32_t ldx, float* h_out, float* sin_out, void* lu_out,
                               int32_t* sel_ev, float* sel_dt, void* stream);
/* The same with the rows read in place: node n's memory row and last_update come straight out of the shard of
 * its owner (rank n % world, local row n / world; peer_memory[r] / peer_last_update[r] = rank r's mapping of
 * the symmetric allocations) -- tgn_part_gather_p2p + tgn_msg_build_gathered as one launch, no staging
 * buffer.  Last aggregation. */
int32_t tgn_msg_build_p2p(const tgn_msgstore* st, const int64_t* n_id, int32_t num, const int32_t* num_dev,
                          const void* const* peer_memory, const void* const* peer_last_update, int32_t world,
                          int32_t memory_dim, const float* time_w, const float* time_b, int32_t time_dim, float* x,
                          int32_t ldx, float* h_out, float* sin_out, void* lu_out, int32_t* sel_ev, float* sel_dt,
                          void* stream);
int32_t tgn_memory_scatter_owned(const int64_t* n_id, int32_t num, const int32_t* num_dev,
                                 const float* new_mem, const void* new_lu, int32_t lu_is_float,
                                 const int64_t* src_rows, int32_t dim, int32_t rank, int32_t world,
                                 float* memory_local, int64_t* last_update_local, void* stream);

/* ------------------------------------------------------------------------- *
 * Dense fp32 building block: C[M,N] (=|+=) A[M,K] * B^T|B (+ bias).
 * a_rows (nullable) gathers rows of A:  A_m = A_base[a_rows[m], :].
 * trans_a: A stored [K,M]; trans_b: B stored [K,N] (else [N,K]).
 * split_k > 1 accumulates atomically into C (C must be initialised).
 * m_dev / k_dev (nullable) read the row count / reduction length from device.
 * Used for the GRU gate GEMMs (torch.nn.GRUCell at memory_module.py:72,172),
 * the TransformerConv projections (emb_module.py:21-23,29) and the decoder
 * (decoder.py:24-27) and for their gradients.
 * ------------------------------------------------------------------------- */
int32_t tgn_sgemm(const float* a, const int64_t* a_rows, const float* b, const float* bias,
                  float* c, int32_t m, const int32_t* m_dev, int32_t n, int32_t k,
                  const int32_t* k_dev, int32_t lda, int32_t ldb, int32_t ldc, int32_t trans_a,
                  int32_t trans_b, int32_t accumulate, int32_t split_k, void* stream);

/* Same contract as tgn_sgemm on the tcgen05 tensor cores (kind::tf32, fp32
 * accumulate in TMEM).  precision = 1: operands rounded to tf32 (~1e-3 relative);
 * precision = 3: 3xTF32 error-compensated split, fp32-level accuracy (~1e-6). */
int32_t tgn_tc_gemm(const float* a, const int64_t* a_rows, const float* b, const float* bias,
                    float* c, int32_t m, const int32_t* m_dev, int32_t n, int32_t k,
                    const int32_t* k_dev, int32_t lda, int32_t ldb, int32_t ldc, int32_t trans_a,
                    int32_t trans_b, int32_t accumulate, int32_t split_k, int32_t precision,
                    void* stream);

/* Batched TMA + tcgen05 GEMM launch: up to 4 independent problems of the tgn_tc_gemm
 * contract in ONE kernel (one CTA per 128x128 tile of any of them) -- e.g. the two GRU
 * gate GEMMs, or a weight gradient together with the matching input gradient.
 * Operands must be 16-byte aligned with lda/ldb multiples of 4 (TMA); no row gather.
 * mode: 0 store, 1 C += , 2 atomic C += (required for split_k > 1; C initialised by caller).
 * With k_dev the reduction stops at *k_dev exactly (the tail of the last k-block is
 * masked in shared memory); with m_dev rows >= *m_dev of C are left untouched. */
typedef struct tgn_gemm_desc {
  const float* a;
  const float* b;
  const float* bias; /* nullable, [n] */
  float* c;
  const int32_t* m_dev; /* nullable */
  const int32_t* k_dev; /* nullable */
  int32_t m, n, k;
  int32_t lda, ldb, ldc;
  int32_t trans_a, trans_b;
  int32_t mode;
  int32_t split_k;
} tgn_gemm_desc;
int32_t tgn_gemm_batch(const tgn_gemm_desc* problems, int32_t count, int32_t precision,
                       void* stream);

/* Fused GRUCell forward (torch.nn.GRUCell, reference modules/memory_module.py:72,172): both gate
 * GEMMs on the tcgen05 tensor cores (operands TMA-staged, fp32 accumulators in TMEM: gi in
 * columns [0,128), gh in [128,256) of a CTA that owns 128 rows x 40 hidden units) and the gate
 * math in the epilogue -- gi / gh never reach global memory.
 *   x [num, ldx] (dx live columns), h [num, dim], w_ih [3*dim, ldw_ih], w_hh [3*dim, dim]
 *   out [num, dim] = h';  gates [num, 4*dim] (nullable) = r, z, n, gh_n for the backward pass
 * Same results as tgn_gemm_batch + tgn_gru_gates_fwd at the same precision (1: tf32, 3: 3xTF32).
 * dim, ldx, ldw_ih multiples of 4; all pointers 16-byte aligned; rows >= *num_dev are untouched. */
int32_t tgn_gru_fused_fwd(const float* x, int32_t ldx, int32_t dx, const float* h, int32_t dim,
                          const float* w_ih, int32_t ldw_ih, const float* w_hh, const float* b_ih,
                          const float* b_hh, int32_t num, const int32_t* num_dev, int32_t precision,
                          float* out, float* gates, void* stream);

/* GRUCell / RNNCell gate math (torch.nn.GRUCell semantics, gate order r,z,n):
 *   gi = x W_ih^T + b_ih [S,3D], gh = h W_hh^T + b_hh [S,3D]  (from tgn_sgemm)
 *   r = sig(gi_r+gh_r), z = sig(gi_z+gh_z), n = tanh(gi_n + r*gh_n)
 *   out = (1-z)*n + z*h ;  h = memory[h_rows[s]] (h_rows nullable -> h[s])
 * gates (nullable) receives r,z,n,gh_n [S,4D] for the backward pass. */
int32_t tgn_gru_gates_fwd(const float* gi, const float* gh, const float* h,
                          const int64_t* h_rows, int32_t num, const int32_t* num_dev,
                          int32_t dim, float* out, float* gates, void* stream);
/* d_gi [S,3D], d_gh [S,3D] (and optional d_h [S,D]) from d_out. */
int32_t tgn_gru_gates_bwd(const float* d_out, const float* gates, const float* h,
                          const int64_t* h_rows, int32_t num, const int32_t* num_dev,
                          int32_t dim, float* d_gi, float* d_gh, float* d_h, void* stream);
/* tgn_gru_gates_bwd (h given by row) that also accumulates (+=) the bias gradients
 * d_b_ih[3D] = column sums of d_gi, d_b_hh[3D] = column sums of d_gh. */
int32_t tgn_gru_gates_bwd_bias(const float* d_out, const float* gates, const float* h, int32_t num,
                               const int32_t* num_dev, int32_t dim, float* d_gi, float* d_gh,
                               float* d_b_ih, float* d_b_hh, void* stream);
int32_t tgn_rnn_gates_fwd(const float* gi, const float* gh, int32_t num, int32_t dim, float* out,
                          void* stream);

/* In-place state write (memory_module.py:147-150):
 *   memory[n_id[s],:] = new_mem[src_rows ? src_rows[s] : s, :]; last_update likewise
 * (new_lu is int64 or float32; float values are truncated like tensor.long()) */
int32_t tgn_memory_scatter(const int64_t* n_id, int32_t num, const int32_t* num_dev,
                           const float* new_mem, const void* new_lu, int32_t lu_is_float,
                           const int64_t* src_rows, int32_t dim, float* memory,
                           int64_t* last_update, void* stream);

/* out[s,:] = table[rows[s],:]  (memory[n_id], memory_module.py:172) */
int32_t tgn_gather_rows(const float* table, const int64_t* rows, int32_t num,
                        const int32_t* num_dev, int32_t dim, float* out, void* stream);
/* out[c] (+)= sum_r x[r,c]  -- bias gradients */
int32_t tgn_colsum(const float* x, int32_t rows, const int32_t* rows_dev, int32_t cols, int32_t ld,
                   float* out, int32_t accumulate, void* stream);
/* TimeEncoder backward: accumulates d_w, d_b (+=) from grad[num, dim] of cos(w*t+b);
 * rows with row_mask[i] < 0 (nullable) are skipped. */
int32_t tgn_time_encode_bwd(const float* t, const int32_t* row_mask, int32_t num,
                            const int32_t* num_dev, const float* w, const float* b, int32_t dim,
                            const float* grad, int32_t ld_grad, float* d_w, float* d_b,
                            void* stream);
/* torch.optim.Adam step (pyg_model_utils.py:38-43 getOptimizer) over a flat
 * parameter buffer; *step_dev (float) is the step counter, advanced by 1. */
int32_t tgn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                      int64_t count, float lr, float beta1, float beta2, float eps,
                      float* step_dev, void* stream);

/* tgn_adam_step with the end-of-step scalars folded in: *step_counter += 1 (nullable; keys the
 * dropout stream) and *loss_out = *loss_acc (nullable).  With done_counter (device uint32, zero on
 * first use) everything is ONE launch -- the last block to finish runs the tail -- and
 * zero_grads != 0 additionally clears grads, *loss_acc and zero_extra[0..zero_extra_count) so the
 * next step needs no memset. */
int32_t tgn_adam_finish(float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                        int64_t count, float lr, float beta1, float beta2, float eps, float* step_dev,
                        int64_t* step_counter, float* loss_acc, float* loss_out, uint32_t* done_counter,
                        int32_t zero_grads, float* zero_extra, int64_t zero_extra_count, void* stream);

/* TimeEncoder forward: out[i,c] = cos(w[c]*t[i] + b[c]) (contract from
 * memory_module.py:203, emb_module.py:27; DGL twin model_utils.py:223-237) */
int32_t tgn_time_encode(const float* t, int32_t num, const float* w, const float* b,
                        int32_t dim, float* out, void* stream);

/* ------------------------------------------------------------------------- *
 * Fused temporal attention (GraphAttentionEmbedding.forward emb_module.py:25-29
 * + torch_geometric TransformerConv(heads=H, concat=True, root_weight=True,
 * edge_dim=time_dim+raw_dim, beta=False)).
 *   proj   [Nb, 4*H*C]  = x * [Wq;Wk;Wv;Wskip]^T + bias   (from tgn_sgemm)
 *   per edge e (j -> i): rel_t = last_update[j] - t[e]  (int64 or float32 each,
 *     torch type promotion: int-int exact then float, otherwise float32);
 *     ea = [cos(w*rel_t+b), msg_e];  ee = We * ea   [H*C]
 *     a  = <q_i, k_j + ee> / sqrt(C) per head; alpha = softmax over edges of i
 *     out_i = sum alpha * (v_j + ee) + skip_i
 * Edges must be grouped by centre: centre c owns edges [row_ptr[c], row_ptr[c+1])
 * listed in edge_perm (nullable = identity); centre_ids[c] (nullable = c) is the
 * row of x/out it refers to.  The edge payload (msg row AND t_edge entry) of edge e is
 * read at row msg_rows[e] when msg_rows is given (e.g. the event ids from the neighbour
 * lookup, indexing the resident event arrays), else at row e.  The dropout seed is
 * seed + *seed_dev (seed_dev nullable), so a replayed CUDA graph draws a fresh mask.
 * Rows of `out` that are not centres must be pre-filled with skip by the
 * caller (tgn_attn_fill_skip).  alpha_out (nullable) [E,H] is kept for backward.
 * dropout on alpha uses Philox(seed) keyed by (edge, head); p = 0 disables.
 * ------------------------------------------------------------------------- */
int32_t tgn_attn_fwd(const float* proj, const void* last_update_local, int32_t lu_is_float,
                     const int64_t* nbr_local, const void* t_edge, int32_t t_is_float,
                     const float* msg, const int64_t* msg_rows, const int32_t* row_ptr,
                     const int32_t* edge_perm, const int64_t* centre_ids, int32_t num_centres, const int32_t* num_centres_dev, int32_t heads,
                     int32_t head_dim, int32_t raw_dim, int32_t time_dim, const float* w_edge,
                     const float* time_w, const float* time_b, float dropout_p, uint64_t seed, const int64_t* seed_dev,
                     float* out, float* alpha_out, float* ee_out, void* stream);
/* out[r,:] = proj[r, 3*H*C : 4*H*C]  for all Nb rows */
int32_t tgn_attn_fill_skip(const float* proj, int32_t num_rows, const int32_t* num_rows_dev,
                           int32_t hc, float* out, void* stream);

/* Backward of tgn_attn_fwd.  tgn_attn_bwd_init writes d_proj = [0,0,0,d_out]
 * (skip block) for every row; tgn_attn_bwd then adds the q/k/v blocks and
 * writes d_ee [E,H*C] (gradient of W_edge*edge_attr).  tgn_attn_edge_attr
 * re-materialises edge_attr [E,time_dim+raw_dim] and rel_t [E] for the W_edge
 * and TimeEncoder gradients (rows >= *num_edges_dev are zero-filled). */
int32_t tgn_attn_bwd_init(const float* d_out, int32_t num_rows, const int32_t* num_rows_dev,
                          int32_t hc, float* d_proj, void* stream);
int32_t tgn_attn_bwd(const float* proj, const int64_t* nbr_local, const int32_t* row_ptr,
                     const int32_t* edge_perm, const int64_t* centre_ids, int32_t num_centres,
                     const int32_t* num_centres_dev, int32_t heads, int32_t head_dim,
                     const float* alpha, const float* ee, const float* d_out, float dropout_p,
                     uint64_t seed, const int64_t* seed_dev, float* d_proj, float* d_ee, void* stream);
int32_t tgn_attn_edge_attr(const void* last_update_local, int32_t lu_is_float,
                           const int64_t* nbr_local, const void* t_edge, int32_t t_is_float,
                           const float* msg, const int64_t* msg_rows, int32_t num_edges,
                           const int32_t* num_edges_dev, int32_t raw_dim, int32_t time_dim,
                           const float* time_w, const float* time_b, float* edge_attr,
                           float* rel_t, void* stream);

/* ------------------------------------------------------------------------- *
 * Pieces of the captured training step that run between the GEMMs.
 * ------------------------------------------------------------------------- */
/* oa = assoc[a], ob = assoc[b], oc = assoc[c] in one launch (neighbor_loader.py:48,
 * epoch_utils.py:99,262) */
int32_t tgn_relabel3(const int64_t* a, int32_t na, const int32_t* na_dev, int64_t* oa,
                     const int64_t* b, int32_t nb, const int32_t* nb_dev, int64_t* ob,
                     const int64_t* c, int32_t nc, const int32_t* nc_dev, int64_t* oc,
                     const int64_t* assoc, void* stream);
/* edge_attr[e, 0:ld] = [cos(w*rel_t+b) (time_dim), msg (raw_dim), 0...] with
 * rel_t = last_update[nbr[e]] - t_edge[msg_rows[e]] (emb_module.py:26-28), int64 times;
 * sin_out [E,time_dim] (nullable) keeps sin(w*rel_t+b) for tgn_time_bwd_sin.
 * num_events = rows of t_edge / msg: a msg_rows[e] outside [0, num_events) yields a zero row and raises
 * TGN_DEVERR_EVENT_RANGE (0 = unchecked). */
int32_t tgn_edge_attr_ld(const int64_t* last_update_local, const int64_t* nbr_local,
                         const int64_t* t_edge, const float* msg, const int64_t* msg_rows,
                         int64_t num_events, int32_t num_edges, const int32_t* num_edges_dev, int32_t raw_dim,
                         int32_t time_dim, const float* time_w, const float* time_b, int32_t ld,
                         float* edge_attr, float* sin_out, float* rel_t, void* stream);
/* TimeEncoder gradient from stored sines: d_w[c] += sum_i grad[i,c] * -sin[i,c] * t[i],
 * d_b[c] += sum_i grad[i,c] * -sin[i,c]; rows with row_mask[i] < 0 are skipped. */
int32_t tgn_time_bwd_sin(const float* t, const int32_t* row_mask, int32_t num, const int32_t* num_dev,
                         const float* sin_vals, int32_t dim, const float* grad, int32_t ld_grad,
                         float* d_w, float* d_b, void* stream);
/* TransformerConv core with the edge projection ee = W_edge*edge_attr [E,H*C] precomputed
 * (by tgn_gemm_batch): scores, softmax over each centre's edges [row_ptr[c], row_ptr[c+1]),
 * dropout, aggregation, + skip.  Writes out[centre_ids[c], :] and alpha [E,H].  max_degree > 0 is a
 * promise that no centre has more edges (the ring lookup yields at most K): up to 12 edges and
 * H*C <= 128 select a register-resident single-pass kernel; 0 = unknown. */
int32_t tgn_attn_core_fwd(const float* proj, const int64_t* nbr_local, const int32_t* row_ptr,
                          const int64_t* centre_ids, int32_t num_centres,
                          const int32_t* num_centres_dev, int32_t heads, int32_t head_dim,
                          const float* ee, float dropout_p, uint64_t seed, const int64_t* seed_dev,
                          int32_t max_degree, float* out, float* alpha, void* stream);
/* Backward: zero-fills d_proj [num_rows, 4*H*C] (num_rows = 0: the caller has already zero-filled
 * it, e.g. on another stream off the critical path), then writes the q and skip blocks of the
 * centre rows, atomically adds the k / v blocks of the neighbour rows and writes d_ee [E,H*C].
 * d_out is read at the centre rows only. */
int32_t tgn_attn_core_bwd(const float* proj, const int64_t* nbr_local, const int32_t* row_ptr,
                          const int64_t* centre_ids, int32_t num_centres,
                          const int32_t* num_centres_dev, int32_t heads, int32_t head_dim,
                          const float* ee, const float* alpha, const float* d_out, float dropout_p,
                          uint64_t seed, const int64_t* seed_dev, int32_t max_degree, int32_t num_rows,
                          float* d_proj, float* d_ee, void* stream);
/* LinkPredictor tail + BCE-with-logits loss and gradient (decoder.py:24-27, pyg-mem-tgn.py:51).
 * hs [B,D] = lin_src(z_src), hd [2B,D] = lin_dst([z_dst; z_neg]).  Accumulates (+=) loss
 * (mean over positives + mean over negatives), d_w_final, d_b_final, d_b_src, d_b_dst;
 * writes logits [2B] (nullable), dh [2B,D] (gradient of hd) and dhs [B,D] (gradient of hs). */
int32_t tgn_dec_loss(const float* hs, const float* hd, const float* w_final, const float* b_final,
                     int32_t batch, int32_t dim, float* loss, float* logits, float* dh, float* dhs,
                     float* d_w_final, float* d_b_final, float* d_b_src, float* d_b_dst,
                     void* stream);
/* The whole decoder of a training step in one launch: LinkPredictor forward for the B positive
 * pairs (ids_local[i], ids_local[B+i]) and the B negative pairs (ids_local[i], ids_local[2B+i])
 * on rows of emb [*,dim], BCE-with-logits loss, and all gradients (exact fp32):
 * loss (+=), d_emb rows (+=, atomics), d_w_src/d_w_dst [dim,dim] (+=), biases, d_w_final, d_b_final
 * (+=); logits [2B] optional.  Needs tgn_dec_fused_smem_bytes(dim) <= 220 KB of shared memory
 * (dim <= 116); larger decoders use tgn_gemm_batch + tgn_dec_loss.
 * Deferred weight gradients: with z_rows and g_rows given (both [3B,dim]) d_w_src / d_w_dst are NOT
 * touched; the kernel writes z_rows = [z_src; z_dst; z_neg] and g_rows = [g_src; g_pos; g_neg] (the
 * gradients of the hidden layer) and the caller forms d_w_src += g_src^T z_src,
 * d_w_dst += [g_pos; g_neg]^T [z_dst; z_neg] with tgn_gemm_batch off the dependent chain.
 * Stream contract: w_src / w_dst are staged into shared memory AHEAD of the programmatic-launch wait, so
 * the launch immediately in front of this one on the stream must not be the one that writes them (in the
 * training step the optimiser runs a whole step earlier); every other input is read after the wait. */
int64_t tgn_dec_fused_smem_bytes(int32_t dim);
int32_t tgn_dec_fused(const float* emb, const int64_t* ids_local, int32_t batch, int32_t dim,
                      const float* w_src, const float* b_src, const float* w_dst, const float* b_dst,
                      const float* w_final, const float* b_final, float* loss, float* logits,
                      float* d_emb, float* d_w_src, float* d_b_src, float* d_w_dst, float* d_b_dst,
                      float* d_w_final, float* d_b_final, float* z_rows, float* g_rows, void* stream);
/* tgn_attn_core_fwd + tgn_dec_fused as ONE launch (emb_module.py:25-29 feeding decoder.py:24-27 at
 * pyg_epoch_utils.py:118-119): the decoder's input rows are not gathered from an embedding table, they are
 * computed in the kernel -- ids_root[3*batch] = index of every batch id (src | dst | neg) in the root list
 * (row_ptr / centre_ids are indexed by it), ids_local[3*batch] = its row in d_emb.  Two heads, <= 10 edges per
 * centre, heads * head_dim <= 128.  alpha [E, heads] gets the softmax weights for tgn_attn_core_bwd; the two
 * [D,D] weight gradients are deferred (z_rows / g_rows as in tgn_dec_fused). */
int32_t tgn_dec_attn_fused(const float* proj, const int64_t* nbr_local, const int32_t* row_ptr,
                           const int64_t* centre_ids, int32_t num_centres, int32_t heads, int32_t head_dim,
                           const float* ee, float dropout_p, uint64_t seed, const int64_t* seed_dev,
                           int32_t max_degree, float* alpha, const int64_t* ids_root, const int64_t* ids_local,
                           int32_t batch, const float* w_src, const float* b_src, const float* w_dst,
                           const float* b_dst, const float* w_final, const float* b_final, float* loss,
                           float* logits, float* d_emb, float* d_b_src, float* d_b_dst, float* d_w_final,
                           float* d_b_final, float* z_rows, float* g_rows, void* stream);
/* TGB evaluation scoring (epoch_utils.py:99-113, decoder.py:24-27): for positive i the score of
 * (src_rows[i], dst_rows[i]) and of its num_neg negatives (src_rows[i], neg_rows[i,q]) as sigmoid
 * outputs; gt_out[i] = #{neg > pos}, ge_out[i] = #{neg >= pos} (the two counts of the TGB MRR
 * rank = 1 + (gt + ge)/2, additive over shards of the negatives, so data-parallel evaluation
 * all-reduces 2 integers per positive).  neg_out [num_pos, num_neg] is optional. */
/* out[i] = in[offset + i*stride] while that entry exists (i < *out_count_dev), -1 beyond: one rank's
 * round-robin share of a sorted unique root list (data-parallel evaluation embeds every root on exactly
 * one rank, epoch_utils.py:74-99). */
int32_t tgn_stride_select(const int64_t* in, int32_t num_in, const int32_t* num_in_dev, int32_t offset,
                          int32_t stride, int64_t* out, int32_t out_cap, int32_t* out_count_dev, void* stream);
/* Data-parallel evaluation: rows of a batch's candidates in the all-gathered table of projected root rows
 * laid out [world][2][block_rows/2][dim] (root with global index g: row (g % world) * block_rows + g / world).
 * global_index = [src (B) | dst (B) | neg (B x Q)]; neg_rows [B, Qr] gets this rank's columns rank, rank+world, ... */
int32_t tgn_dp_rows(const int64_t* global_index, int32_t num_pos, int32_t num_neg, int32_t rank, int32_t world,
                    int64_t block_rows, int64_t* src_rows, int64_t* dst_rows, int64_t* neg_rows, void* stream);
int32_t tgn_score_negs(const float* hs, const float* hd, const int64_t* src_rows,
                       const int64_t* dst_rows, const int64_t* neg_rows, int32_t num_pos,
                       int32_t num_neg, int32_t dim, const float* w_final, const float* b_final,
                       float* pos_out, float* neg_out, int32_t* gt_out, int32_t* ge_out,
                       void* stream);
/* dst[rows[i], :] += src[i, :] (atomic) */
int32_t tgn_scatter_add_rows(const float* src, const int64_t* rows, int32_t num,
                             const int32_t* num_dev, int32_t dim, float* dst, void* stream);

/* Batch staging (replaces the host DataLoader walk, temporal_dataset.py:34-57 +
 * epoch_utils.py:186-215): slices events [*pos_dev, *pos_dev + batch) of the
 * resident event arrays into ids3 = [src|dst|neg] (int64 [3*batch]), t_i64, t_f32 and
 * msg [batch, raw_dim], then advances *pos_dev by batch.  Events at or past num_events are replaced
 * by a filler (node 0, t 0, zero message) so that a pipelined step may pre-load one batch ahead. */
int32_t tgn_batch_load(const int64_t* src_all, const int64_t* dst_all, const int64_t* neg_all,
                       const int64_t* t_all, const float* msg_all, int32_t raw_dim, int32_t batch,
                       int64_t num_events, int64_t* pos_dev, int64_t* ids3, int64_t* t_i64, float* t_f32, float* msg,
                       void* stream);

/* LinkPredictor forward (decoder.py:24-27) on gathered rows:
 *   h = relu(Ws z[a] + bs + Wd z[b] + bd); score = wf . h + bf  (logit; sigmoid optional)
 * hs/hd [Nb,C] are the per-node projections (tgn_sgemm); a/b are row ids. */
int32_t tgn_link_score(const float* hs, const float* hd, const int64_t* a_rows,
                       const int64_t* b_rows, int32_t num, int32_t dim, const float* w_final,
                       const float* b_final, int32_t apply_sigmoid, float* out, void* stream);

/* TGB MRR (epoch_utils.py:108-113; tgb Evaluator): per positive
 *   rank = 1 + 0.5*(#{neg > pos} + #{neg >= pos}); rr = 1/rank */
int32_t tgn_mrr(const float* pos, const float* neg, int32_t num_pos, int32_t num_neg,
                float* rr_out, void* stream);

/* ------------------------------------------------------------------------- *
 * Negatives and the epoch metric on the device (no per-batch host round trip).
 * tgn_neg_dest_sample = neg_sampler.NegLinkSamplerDest.sample (neg_sampler.py:8-23):
 *   out[i] uniform over dst_nodes[num_dst], redrawn while == pos_dst[i] (kept if num_dst == 1).
 * tgn_neg_fill: out[batch, num_neg] uniform over [lo, hi) without pos_dst[i]
 *   (synthetic stand-in for tgb negative_sampler.query_batch, epoch_utils.py:43).
 * Both draw from Philox-4x32-10 keyed by seed with counters (row, call): the same (seed, call)
 * reproduces the same negatives for any launch geometry.
 * tgn_rank_accum (epoch_utils.py:108-113,163): acc[0] += mean_i 1/(1 + (gt[i]+ge[i])/2),
 *   acc[1] += 1; epoch MRR = acc[0] / acc[1].  rr_out (nullable) gets the per-positive values.
 * ------------------------------------------------------------------------- */
int32_t tgn_neg_dest_sample(const int64_t* dst_nodes, int32_t num_dst, const int64_t* pos_dst,
                            int32_t batch, uint64_t seed, uint64_t call, int64_t* out, void* stream);
int32_t tgn_neg_fill(const int64_t* pos_dst, int32_t batch, int32_t num_neg, int64_t lo, int64_t hi,
                     uint64_t seed, uint64_t call, int64_t* out, void* stream);
int32_t tgn_rank_accum(const int32_t* gt, const int32_t* ge, int32_t batch, double* acc, float* rr_out,
                       void* stream);
/* Training-batch AP / AUC on the device (epoch_utils.py:312-317: sklearn average_precision_score and
 * roc_auc_score of sigmoid(logits) per batch, their means printed per epoch): logits = [num_pos positives |
 * num_neg negatives]; acc[0] += AP, acc[1] += AUC, acc[2] += 1.  No per-batch host copy. */
int32_t tgn_ap_auc_accum(const float* logits, int32_t num_pos, int32_t num_neg, double* acc, void* stream);

/* ------------------------------------------------------------------------- *
 * Owner-side compute for the partitioned node memory (modules/memory_module.py:126-150,
 * modules/msg_func.py:17-18): rank r builds the messages and runs the GRU only for the rows of a step it
 * owns (n % world == r), publishes h' / last_update' to every rank's row table through the peer mapping,
 * runs the GRU backward on its rows, and the optimiser sums the partial gradients out of peer memory.
 *   tgn_part_select_owned  own_nodes / own_pos [own_cap]: owned entries of n_id and their positions, in
 *                          position order; *own_count_dev = how many (clamped to own_cap, which raises
 *                          TGN_DEVERR_OWNER_CAP).
 *   tgn_part_publish       peer_rows[p][own_pos[i], :] = rows[i, :], peer_last_update[p][own_pos[i]] =
 *                          last_update[i] for every rank p (128-bit stores over NVLink; dim % 4 == 0).
 *                          The caller brackets it with rank barriers.
 *   tgn_adam_finish_peers  tgn_adam_finish with gradient = grads_replicated[i] + sum_p
 *                          peer_grads_partial[p][i]; grads_replicated is ONE rank's copy of the replicated
 *                          part (all ranks pass the same mapping, so the weight replicas stay bit-identical);
 *                          only the first partial_count parameters (the memory path) have partial sums.
 *                          No gradient clearing (peers may still be reading): the caller clears after a barrier.
 * ------------------------------------------------------------------------- */
int32_t tgn_part_select_owned(const int64_t* n_id, int32_t num, const int32_t* num_dev, int32_t rank,
                              int32_t world, int64_t* own_nodes, int64_t* own_pos, int32_t own_cap,
                              int32_t* own_count_dev, void* stream);
int32_t tgn_part_publish(const float* rows, const int64_t* last_update, const int64_t* own_pos, int32_t num,
                         const int32_t* num_dev, int32_t dim, void* const* peer_rows, void* const* peer_last_update,
                         int32_t world, void* stream);
/* Sharded evaluation (epoch_utils.py:74-113 data-parallel over model replicas): writes one block of `nbytes`
 * bytes at `table_offset_bytes` of EVERY rank's table (peer_tables[r] = rank r's mapping of the symmetric
 * allocation): the all-gather of the decoder-projected rows / of the per-rank TGB counts as peer stores over
 * NVLink, no collective call.  Sizes, offset and src 16-byte aligned.  The caller separates it from the
 * readers with a rank barrier. */
int32_t tgn_peer_bcast(const void* src, int64_t nbytes, void* const* peer_tables, int64_t table_offset_bytes,
                       int32_t world, void* stream);
int32_t tgn_adam_finish_peers(float* params, const float* grads_replicated, const void* const* peer_grads_partial,
                              int32_t world, int64_t partial_count, float* exp_avg, float* exp_avg_sq, int64_t count, float lr,
                              float beta1, float beta2, float eps, float* step_dev, int64_t* step_counter,
                              const float* loss_acc, float* loss_out, uint32_t* done_counter, void* stream);

/* ------------------------------------------------------------------------- *
 * EdgeGATConv attention core of the reference's live DGL stack (model_utils.py:565-612).
 * Per head: el'_e = el[src(e)] + ee[e] (:594), z_e = LeakyReLU(el'_e + er[dst(e)]) (:595-596),
 * a = edge_softmax by destination then attention dropout (:597), s[v] = sum_e a_e * el'_e (:560-563,599 --
 * the reference aggregates the LOGIT part el_prime, a scalar per head, not the projected features).
 * el/er [N,H] and ee [E,H] are the attn_l/attn_r/attn_e-weighted projections (:587-589), which the
 * caller folds into skinny weights.  Edges are addressed through a CSR by destination: row_ptr [N+1]
 * and edge_perm [E] (nullable = edges already grouped); src [E] by edge id.  alpha_out [E,H] keeps the
 * softmax weights before dropout; the dropout mask is Philox(seed; edge, head) in both passes.
 * Backward: d_el must be zero-filled (atomic accumulation over sources).
 * ------------------------------------------------------------------------- */
int32_t tgn_egat_attn_fwd(const float* el, const float* er, const float* ee, const int32_t* row_ptr,
                          const int32_t* edge_perm, const int64_t* src, int32_t num_nodes, int32_t num_edges,
                          int32_t heads, float negative_slope, float dropout_p, uint64_t seed, float* s_out,
                          float* alpha_out, void* stream);
int32_t tgn_egat_attn_bwd(const float* el, const float* er, const float* ee, const int32_t* row_ptr,
                          const int32_t* edge_perm, const int64_t* src, int32_t num_nodes, int32_t num_edges,
                          int32_t heads, float negative_slope, float dropout_p, uint64_t seed, const float* alpha,
                          const float* d_s, float* d_el, float* d_er, float* d_ee, void* stream);

/* ------------------------------------------------------------------------- *
 * Dependency-aware block ids (dependencyGraph.py:8-28 get_block, :33-49 dependecyAwareBatch): the event
 * stream is cut into consecutive batches of `batch` events (the DataLoader's, utils.py:52-54); inside a
 * batch, walking in order, block_ids[i] = 1 + the highest block id already given to either endpoint of
 * event i in this batch (0 if neither was seen).  One CTA per batch (shared-memory sort of the 2*batch
 * endpoint keys, then relaxation of the <= 2-predecessor recurrence).  num_blocks (nullable) gets the
 * number of blocks per batch [ceil(num_events / batch)].  2*batch <= TGN_SORT_MAX; node ids in [0, 2^42).
 * ------------------------------------------------------------------------- */
int32_t tgn_dep_blocks(const int64_t* src, const int64_t* dst, int64_t num_events, int32_t batch,
                       int32_t* block_ids, int32_t* num_blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TGN_B200_H_ */
