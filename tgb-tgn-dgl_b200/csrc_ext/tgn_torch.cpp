// Torch C++ extension over the C-ABI (include/tgn_b200.h): the binding north_star describes ("the GPU work
// exposed as a torch C++/CUDA extension over a thin C-ABI"), built by `python setup.py build_ext --inplace`
// like TGL's sampler_core (reference README.md:1-2).  The ops take and return torch tensors, allocate their
// outputs with the caching allocator, run on the current CUDA stream and raise from a non-zero C-ABI status;
// the kernels themselves stay behind libtgn_b200.so.  Registered as torch.ops.tgn.*:
//
//   nbr_lookup      LastNeighborLoader.__call__ gathers + compaction   (neighbor_loader.py:26-44)
//   nbr_insert      LastNeighborLoader.insert                          (neighbor_loader.py:52-104)
//   tcsr_sample     TGL ParallelSampler::sample_layer                  (README.md:1-5, config/TGN.yml:1-9)
//   agg_last        LastAggregator.forward                             (modules/msg_agg.py:15-21)
//   agg_mean        MeanAggregator.forward                             (modules/msg_agg.py:24-26)
//   dep_blocks      dependencyGraph.get_block for every batch          (dependencyGraph.py:8-28)
//
// The Python modules keep their ctypes route to the same entry points (tgn_b200/_cabi.py); tgn_b200/torch_ext.py
// loads this extension when it has been built and ops.py then prefers it for the calls above (one dispatcher
// hop instead of ~20 ctypes argument conversions per call).
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>
#include <torch/library.h>

#include <string>
#include <vector>

#include "../../include/tgn_b200.h"

namespace {

void chk(int32_t rc) {
  if (rc != 0) {
    const char* m = tgn_last_error();
    TORCH_CHECK(false, rc == -1 ? "TGN_EINVAL: " : "TGN_ECUDA: ", m ? m : "");
  }
}
void* cur_stream() { return (void*)c10::cuda::getCurrentCUDAStream().stream(); }
void need(const at::Tensor& t, at::ScalarType dt, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (the B200 hot path has no CPU fallback)");
  TORCH_CHECK(t.scalar_type() == dt, name, " has the wrong dtype");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}
at::TensorOptions on(const at::Tensor& like, at::ScalarType dt) { return like.options().dtype(dt); }

std::vector<at::Tensor> nbr_lookup(const at::Tensor& n_id, const at::Tensor& neighbors, const at::Tensor& e_id,
                                   const at::Tensor& t, const c10::optional<at::Tensor>& bitmap) {
  need(n_id, at::kLong, "n_id"); need(neighbors, at::kLong, "neighbors"); need(e_id, at::kLong, "e_id");
  need(t, at::kFloat, "t");
  c10::cuda::CUDAGuard guard(n_id.device());
  const int64_t N = neighbors.size(0), K = neighbors.size(1), R = n_id.numel();
  const int64_t cap = std::max<int64_t>(R * K, 1);
  auto o_n = at::empty({cap}, on(n_id, at::kLong)), o_c = at::empty({cap}, on(n_id, at::kLong));
  auto o_e = at::empty({cap}, on(n_id, at::kLong)), o_t = at::empty({cap}, on(n_id, at::kFloat));
  auto off = at::empty({R + 1}, on(n_id, at::kInt)), cnt = at::empty({1}, on(n_id, at::kInt));
  auto ws = at::empty({std::max<int64_t>(tgn_nbr_lookup_ws_bytes((int32_t)R, (int32_t)K), 16) / 8}, on(n_id, at::kLong));
  chk(tgn_nbr_lookup(n_id.data_ptr<int64_t>(), (int32_t)R, nullptr, (int32_t)K, N, neighbors.data_ptr<int64_t>(),
                     e_id.data_ptr<int64_t>(), t.data_ptr<float>(), o_n.data_ptr<int64_t>(), o_c.data_ptr<int64_t>(),
                     o_e.data_ptr<int64_t>(), o_t.data_ptr<float>(), off.data_ptr<int32_t>(), cnt.data_ptr<int32_t>(),
                     bitmap.has_value() ? bitmap->data_ptr() : nullptr, ws.data_ptr(), cur_stream()));
  return {o_n, o_c, o_e, o_t, off, cnt};
}

void nbr_insert(const at::Tensor& src, const at::Tensor& dst, const at::Tensor& t, int64_t cur_e_id,
                at::Tensor neighbors, at::Tensor e_id, at::Tensor t_state) {
  need(src, at::kLong, "src"); need(dst, at::kLong, "dst"); need(t, at::kFloat, "t");
  need(neighbors, at::kLong, "neighbors"); need(e_id, at::kLong, "e_id"); need(t_state, at::kFloat, "t_state");
  c10::cuda::CUDAGuard guard(src.device());
  chk(tgn_nbr_insert(src.data_ptr<int64_t>(), dst.data_ptr<int64_t>(), t.data_ptr<float>(), (int32_t)src.numel(),
                     cur_e_id, nullptr, (int32_t)neighbors.size(1), neighbors.size(0), neighbors.data_ptr<int64_t>(),
                     e_id.data_ptr<int64_t>(), t_state.data_ptr<float>(), cur_stream()));
}

std::vector<at::Tensor> tcsr_sample(const at::Tensor& indptr, const at::Tensor& indices, const at::Tensor& eid,
                                    const at::Tensor& ts, const c10::optional<at::Tensor>& coarse,
                                    const at::Tensor& roots, const at::Tensor& root_ts, int64_t k, int64_t strategy,
                                    double offset, double duration, int64_t seed) {
  need(indptr, at::kInt, "indptr"); need(indices, at::kInt, "indices"); need(eid, at::kInt, "eid");
  need(ts, at::kFloat, "ts"); need(roots, at::kInt, "roots"); need(root_ts, at::kFloat, "root_ts");
  c10::cuda::CUDAGuard guard(roots.device());
  const int64_t R = roots.numel(), cap = std::max<int64_t>(R * k, 1);
  auto o_n = at::empty({cap}, on(roots, at::kInt)), o_c = at::empty({cap}, on(roots, at::kInt));
  auto o_e = at::empty({cap}, on(roots, at::kInt));
  auto o_t = at::empty({cap}, on(roots, at::kFloat)), o_d = at::empty({cap}, on(roots, at::kFloat));
  auto off = at::empty({R + 1}, on(roots, at::kInt)), cnt = at::empty({1}, on(roots, at::kInt));
  auto ws = at::empty({std::max<int64_t>(tgn_tcsr_sample_ws_bytes((int32_t)R), 16) / 8}, on(roots, at::kLong));
  chk(tgn_tcsr_sample(indptr.data_ptr<int32_t>(), indices.data_ptr<int32_t>(), eid.data_ptr<int32_t>(),
                      ts.data_ptr<float>(), coarse.has_value() ? coarse->data_ptr<float>() : nullptr, ts.numel(),
                      indptr.numel() - 1, roots.data_ptr<int32_t>(), root_ts.data_ptr<float>(), (int32_t)R, (int32_t)k,
                      (int32_t)strategy, (float)offset, (float)duration, (uint64_t)seed, o_n.data_ptr<int32_t>(),
                      o_c.data_ptr<int32_t>(), o_e.data_ptr<int32_t>(), o_t.data_ptr<float>(), o_d.data_ptr<float>(),
                      off.data_ptr<int32_t>(), cnt.data_ptr<int32_t>(), ws.data_ptr(), cur_stream()));
  return {o_n, o_c, o_e, o_t, o_d, off, cnt};
}

std::vector<at::Tensor> agg_last(const at::Tensor& msg, const at::Tensor& index, const at::Tensor& t, int64_t dim_size) {
  need(msg, at::kFloat, "msg"); need(index, at::kLong, "index");
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && (t.scalar_type() == at::kLong || t.scalar_type() == at::kFloat),
              "t must be a contiguous int64 or float32 CUDA tensor");
  c10::cuda::CUDAGuard guard(msg.device());
  const int64_t M = msg.size(0), W = msg.size(1);
  auto out = at::empty({dim_size, W}, on(msg, at::kFloat)), arg = at::empty({dim_size}, on(msg, at::kLong));
  auto ws = at::empty({std::max<int64_t>(2 * dim_size, 1)}, on(msg, at::kLong));
  chk(tgn_agg_last(msg.data_ptr<float>(), index.data_ptr<int64_t>(), t.data_ptr(), t.scalar_type() == at::kFloat ? 1 : 0,
                   (int32_t)M, (int32_t)dim_size, (int32_t)W, out.data_ptr<float>(), arg.data_ptr<int64_t>(),
                   ws.data_ptr(), cur_stream()));
  return {out, arg};
}

at::Tensor agg_mean(const at::Tensor& msg, const at::Tensor& index, int64_t dim_size) {
  need(msg, at::kFloat, "msg"); need(index, at::kLong, "index");
  c10::cuda::CUDAGuard guard(msg.device());
  const int64_t M = msg.size(0), W = msg.size(1);
  auto out = at::empty({dim_size, W}, on(msg, at::kFloat));
  auto ws = at::empty({std::max<int64_t>(tgn_agg_mean_ws_bytes((int32_t)M, (int32_t)dim_size), 16) / 8}, on(msg, at::kLong));
  chk(tgn_agg_mean(msg.data_ptr<float>(), index.data_ptr<int64_t>(), (int32_t)M, (int32_t)dim_size, (int32_t)W,
                   out.data_ptr<float>(), ws.data_ptr(), cur_stream()));
  return out;
}

at::Tensor dep_blocks(const at::Tensor& src, const at::Tensor& dst, int64_t batch) {
  need(src, at::kLong, "src"); need(dst, at::kLong, "dst");
  c10::cuda::CUDAGuard guard(src.device());
  auto out = at::empty({src.numel()}, on(src, at::kInt));
  chk(tgn_dep_blocks(src.data_ptr<int64_t>(), dst.data_ptr<int64_t>(), src.numel(), (int32_t)batch,
                     out.data_ptr<int32_t>(), nullptr, cur_stream()));
  return out;
}

int64_t abi_version() { return tgn_abi_version(); }

}  // namespace

TORCH_LIBRARY(tgn, m) {
  m.def("nbr_lookup(Tensor n_id, Tensor neighbors, Tensor e_id, Tensor t, Tensor? bitmap) -> Tensor[]");
  m.def("nbr_insert(Tensor src, Tensor dst, Tensor t, int cur_e_id, Tensor(a!) neighbors, Tensor(b!) e_id, Tensor(c!) t_state) -> ()");
  m.def("tcsr_sample(Tensor indptr, Tensor indices, Tensor eid, Tensor ts, Tensor? coarse, Tensor roots, Tensor root_ts, "
        "int k, int strategy, float offset, float duration, int seed) -> Tensor[]");
  m.def("agg_last(Tensor msg, Tensor index, Tensor t, int dim_size) -> Tensor[]");
  m.def("agg_mean(Tensor msg, Tensor index, int dim_size) -> Tensor");
  m.def("dep_blocks(Tensor src, Tensor dst, int batch) -> Tensor");
  m.def("abi_version() -> int", &abi_version);
}

TORCH_LIBRARY_IMPL(tgn, CUDA, m) {
  m.impl("nbr_lookup", &nbr_lookup);
  m.impl("nbr_insert", &nbr_insert);
  m.impl("tcsr_sample", &tcsr_sample);
  m.impl("agg_last", &agg_last);
  m.impl("agg_mean", &agg_mean);
  m.impl("dep_blocks", &dep_blocks);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) { m.doc() = "torch.ops.tgn.* over libtgn_b200.so (include/tgn_b200.h)"; }
