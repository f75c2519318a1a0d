// Negative destinations and the epoch MRR on the device.
//
//   tgn_neg_dest_sample  neg_sampler.NegLinkSamplerDest.sample (reference neg_sampler.py:8-23):
//                        one negative per positive, uniform over the observed destination set,
//                        redrawn while it equals the positive.  The reference indexes a Python
//                        list per draw on the host; here every positive is one thread with a
//                        counter-based Philox stream (reproducible, launch-geometry independent).
//   tgn_neg_fill         [B, Q] evaluation negatives, uniform over [lo, hi) without the positive
//                        -- the stand-in for TGB's pre-generated negative_sampler.query_batch
//                        (epoch_utils.py:43) on synthetic graphs, generated where it is consumed.
//   tgn_rank_accum       epoch metric of test() (epoch_utils.py:108-113,163): the mean over
//                        batches of the per-batch mean reciprocal rank, accumulated in device
//                        memory from the integer rank counts so an evaluation epoch has no
//                        per-batch host round trip.
#include "../../include/tgn_b200.h"
#include "common.cuh"

namespace tgn {

// unbiased draw from [0, n) out of 32 random bits: multiply-shift with rejection of the short tail
__device__ __forceinline__ bool bounded(uint32_t x, uint32_t n, uint32_t* out) {
  const unsigned long long m = (unsigned long long)x * n;
  const uint32_t lo = (uint32_t)m;
  if (lo < n) {
    const uint32_t thresh = (0u - n) % n;
    if (lo < thresh) return false;
  }
  *out = (uint32_t)(m >> 32);
  return true;
}

__global__ void __launch_bounds__(256)
    neg_dest_kernel(const int64_t* __restrict__ dst_nodes, int n_dst, const int64_t* __restrict__ pos,
                    int B, uint64_t seed, uint64_t call, int64_t* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  Philox rng(seed);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
    const int64_t p = pos[i];
    int64_t v = dst_nodes[0];
    // counter = (positive index, call << 20 | round): rounds are redraws after a collision / rejection
    for (uint32_t round = 0; round < (1u << 20); ++round) {
      const uint4 x = rng((uint64_t)i, (call << 20) | round);
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
      bool done = false;
#pragma unroll
      for (int q = 0; q < 4 && !done; ++q) {
        uint32_t u;
        if (!bounded(xs[q], (uint32_t)n_dst, &u)) continue;
        v = dst_nodes[u];
        done = (v != p) || n_dst == 1;
      }
      if (done) break;
    }
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256)
    neg_fill_kernel(const int64_t* __restrict__ pos, int B, int Q, int64_t lo, int64_t hi,
                    uint64_t seed, uint64_t call, int64_t* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  Philox rng(seed);
  const long long total = (long long)B * ((Q + 3) / 4);
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(w / ((Q + 3) / 4)), q0 = (int)(w % ((Q + 3) / 4)) * 4;
    const int64_t p = pos[i];
    const bool p_in = p >= lo && p < hi;
    const uint32_t span = (uint32_t)(hi - lo - (p_in ? 1 : 0));  // candidates without the positive
    int filled = 0;
    for (uint32_t round = 0; filled < 4 && q0 + filled < Q && round < (1u << 16); ++round) {
      const uint4 x = rng(((uint64_t)i << 32) | (uint32_t)q0, (call << 16) | round);
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t u;
        if (filled < 4 && q0 + filled < Q && bounded(xs[k], span, &u)) {
          int64_t v = lo + u;
          if (p_in && v >= p) ++v;   // skip over the positive
          out[(long long)i * Q + q0 + filled] = v;
          ++filled;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
    rank_accum_kernel(const int32_t* __restrict__ gt, const int32_t* __restrict__ ge, int B,
                      double* __restrict__ acc, float* __restrict__ rr_out) {
  pdl_wait();
  pdl_launch();
  __shared__ double s_part[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float rr = 1.f / (0.5f * (float)(gt[i] + ge[i]) + 1.f);
    if (rr_out) rr_out[i] = rr;
    s += (double)rr;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
    acc[0] += t / (double)B;   // per-batch mean (epoch_utils.py:113)
    acc[1] += 1.0;             // batches (epoch_utils.py:163 averages over them)
  }
}

// Per-batch average precision and ROC AUC of the training logits (epoch_utils.py:312-315: sklearn's
// average_precision_score / roc_auc_score on sigmoid([pos; neg]) against [1...; 0...]), accumulated on the
// device.  For P positives and Nn negatives with scores s:
//   AUC = (sum_{i pos} sum_{j neg} [s_i > s_j] + 0.5 [s_i == s_j]) / (P Nn)
//   AP  = (1/P) sum_{i pos} tp(s_i) / (tp(s_i) + fp(s_i)),  tp / fp = positives / negatives with score >= s_i
// (sklearn sums precision x recall-increment over the distinct thresholds; every positive tied at a threshold
// carries that threshold's precision, which is the same sum).  One CTA; scores staged in shared memory.
__global__ void __launch_bounds__(1024) ap_auc_kernel(const float* __restrict__ logits, int P, int Nn,
                                                      double* __restrict__ acc) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float s_sc[];
  __shared__ double s_part[2][32];
  const int n = P + Nn;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_sc[i] = 1.f / (1.f + expf(-logits[i]));
  __syncthreads();
  double ap = 0.0, auc = 0.0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float v = s_sc[i];
    int tp = 0, fp = 0, gt = 0, eq = 0;
    for (int j = 0; j < P; ++j) tp += s_sc[j] >= v;
    for (int j = P; j < n; ++j) {
      const float u = s_sc[j];
      fp += u >= v;
      gt += v > u;
      eq += v == u;
    }
    ap += (double)tp / (double)(tp + fp);
    auc += (double)gt + 0.5 * (double)eq;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ap += __shfl_xor_sync(0xffffffffu, ap, o);
    auc += __shfl_xor_sync(0xffffffffu, auc, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_part[0][threadIdx.x >> 5] = ap;
    s_part[1][threadIdx.x >> 5] = auc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += s_part[0][w];
      b += s_part[1][w];
    }
    acc[0] += a / (double)P;
    acc[1] += b / ((double)P * (double)Nn);
    acc[2] += 1.0;
  }
}

}  // namespace tgn

using namespace tgn;

extern "C" {

int32_t tgn_neg_dest_sample(const int64_t* dst_nodes, int32_t num_dst, const int64_t* pos_dst,
                            int32_t batch, uint64_t seed, uint64_t call, int64_t* out, void* stream) {
  TGN_REQUIRE(batch >= 0 && num_dst >= 1, "neg_dest_sample: bad sizes");
  if (batch == 0) return TGN_OK;
  TGN_REQUIRE(dst_nodes && pos_dst && out, "neg_dest_sample: NULL pointer");
  launch_k(neg_dest_kernel, dim3(stride_grid(batch, 256)), dim3(256), 0, (cudaStream_t)stream, dst_nodes,
           num_dst, pos_dst, batch, seed, call, out);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_neg_fill(const int64_t* pos_dst, int32_t batch, int32_t num_neg, int64_t lo, int64_t hi,
                     uint64_t seed, uint64_t call, int64_t* out, void* stream) {
  TGN_REQUIRE(batch >= 0 && num_neg >= 0, "neg_fill: bad sizes");
  TGN_REQUIRE(hi - lo >= 2 && hi - lo <= 0xFFFFFFFFll, "neg_fill: need at least two candidates in [lo, hi)");
  if (batch == 0 || num_neg == 0) return TGN_OK;
  TGN_REQUIRE(pos_dst && out, "neg_fill: NULL pointer");
  launch_k(neg_fill_kernel, dim3(stride_grid((long long)batch * ((num_neg + 3) / 4), 256)), dim3(256), 0,
           (cudaStream_t)stream, pos_dst, batch, num_neg, lo, hi, seed, call, out);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_ap_auc_accum(const float* logits, int32_t num_pos, int32_t num_neg, double* acc, void* stream) {
  TGN_REQUIRE(num_pos >= 1 && num_neg >= 1 && (int64_t)num_pos + num_neg <= 48 * 1024,
              "ap_auc_accum: need 1 <= positives, negatives and at most 49152 scores");
  TGN_REQUIRE(logits && acc, "ap_auc_accum: NULL pointer");
  const size_t smem = (size_t)(num_pos + num_neg) * sizeof(float);
  static unsigned long long attr_mask = 0;
  TGN_CUDA(smem_optin(ap_auc_kernel, 48 * 1024 * 4, attr_mask));
  launch_k(ap_auc_kernel, dim3(1), dim3(1024), smem, (cudaStream_t)stream, logits, num_pos, num_neg, acc);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

int32_t tgn_rank_accum(const int32_t* gt, const int32_t* ge, int32_t batch, double* acc, float* rr_out,
                       void* stream) {
  TGN_REQUIRE(batch >= 1, "rank_accum: empty batch");
  TGN_REQUIRE(gt && ge && acc, "rank_accum: NULL pointer");
  launch_k(rank_accum_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, gt, ge, batch, acc, rr_out);
  TGN_LAUNCH_CHECK();
  return TGN_OK;
}

}  // extern "C"
