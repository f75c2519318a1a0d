// Batch staging for the captured training step.
//
// The reference walks a torch DataLoader that builds every batch item by item
// on the host (temporal_dataset.py:34-57, epoch_utils.py:186-215).  Here the
// event arrays live in HBM and one kernel slices batch number *pos_dev/B out of
// them into the step's static buffers, so a replayed CUDA graph needs no host
// work per step:
//   ids3  = [src | dst | neg]          (int64 [3B], roots of the batch, epoch_utils.py:215)
//   t_i64 = t                          (memory / message-store timestamps)
//   t_f32 = float(t)                   (neighbour ring timestamps, temporal_dataset.py:42)
//   msg   = raw messages [B, D_e]
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

__global__ void batch_load_kernel(const int64_t* __restrict__ src_all,
                                  const int64_t* __restrict__ dst_all,
                                  const int64_t* __restrict__ neg_all,
                                  const int64_t* __restrict__ t_all,
                                  const float* __restrict__ msg_all, int De, int B,
                                  int64_t num_events, const int64_t* __restrict__ pos_dev, int64_t* __restrict__ ids3,
                                  int64_t* __restrict__ t_i64, float* __restrict__ t_f32,
                                  float* __restrict__ msg) {
  pdl_wait();
  pdl_launch();
  const int64_t pos = *pos_dev;
  const long long total = (long long)B * (De > 3 ? De : 3);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    // events past the end of the arrays (the pipelined step pre-loads one batch ahead) become
    // a harmless filler: node 0 at time 0 with a zero message; such a batch is never trained on
    if (e < B) {
      const bool in = pos + e < num_events;
      ids3[e] = in ? src_all[pos + e] : 0;
      ids3[B + e] = in ? dst_all[pos + e] : 0;
      ids3[2 * B + e] = in ? neg_all[pos + e] : 0;
      const int64_t t = in ? t_all[pos + e] : 0;
      t_i64[e] = t;
      t_f32[e] = (float)t;
    }
    if (e < (long long)B * De) msg[e] = (pos + e / De < num_events) ? msg_all[pos * De + e] : 0.f;
  }
}

__global__ void advance_kernel(int64_t* p, int64_t by) {
  pdl_wait();
  pdl_launch(); *p += by; }

}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_batch_load(const int64_t* src_all, const int64_t* dst_all, const int64_t* neg_all,
                       const int64_t* t_all, const float* msg_all, int32_t raw_dim, int32_t batch,
                       int64_t num_events, int64_t* pos_dev, int64_t* ids3, int64_t* t_i64, float* t_f32, float* msg,
                       void* stream) {
  TGN_REQUIRE(batch >= 1 && raw_dim >= 0 && num_events >= 0, "batch_load: bad sizes");
  TGN_REQUIRE(src_all && dst_all && neg_all && t_all && (msg_all || raw_dim == 0) && pos_dev &&
                  ids3 && t_i64 && t_f32 && (msg || raw_dim == 0),
              "batch_load: NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const long long total = (long long)batch * (raw_dim > 3 ? raw_dim : 3);
  launch_k(batch_load_kernel, dim3(stride_grid(total, 256)), dim3(256), 0, s, src_all, dst_all, neg_all, t_all,
                                                            msg_all, raw_dim, batch, num_events, pos_dev, ids3,
                                                            t_i64, t_f32, msg);
  TGN_LAUNCH_CHECK();
  launch_k(advance_kernel, dim3(1), dim3(1), 0, s, pos_dev, batch);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
