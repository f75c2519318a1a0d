"""Drop-in `utils` (reference utils.py:17-68): config parsing and the dataset / loader set-up the
driver script calls.  Differences from the reference, all on the host side of the hot path:
  * `tgb` is used when importable, otherwise the offline stand-in `tgb_synth` (no network here);
  * loaders are `TensorBatchLoader`s (tensor slices) instead of per-item-collating DataLoaders --
    they yield the same dict batches;
  * the reference assigns the block ids of the validation loader to the test set and vice versa
    (utils.py:56-57); here each split gets its own.
"""
import yaml

from dependencyGraph import dependecyAwareBatch as dab
from temporal_dataset import TemporalGraphDataset, TensorBatchLoader

try:  # pragma: no cover - not installed in this image
    from tgb.linkproppred.dataset_pyg import PyGLinkPropPredDataset
    from tgb.linkproppred.evaluate import Evaluator
except Exception:
    from tgb_synth import Evaluator, PyGLinkPropPredDataset


def parse_config(f):
    conf = yaml.safe_load(open(f, "r"))
    return conf["sampling"][0], conf["memory"][0], conf["gnn"][0], conf["train"][0]


def getDataWithDependecyBlock(DATA, train_param, csv=False, load_neg_sampler=True):
    if csv:
        raise NotImplementedError("csv input is not implemented (neither is it in the reference, utils.py:26-27)")
    dataset = PyGLinkPropPredDataset(name=DATA, root="datasets")
    data = dataset.get_TemporalData()
    metric = dataset.eval_metric
    neg_sampler = evaluator = None
    if load_neg_sampler:
        dataset.load_val_ns()
        dataset.load_test_ns()
        neg_sampler = dataset.negative_sampler
        evaluator = Evaluator(name=DATA)
    bs = train_param["batch_size"]
    loaders = []
    for mask in (dataset.train_mask, dataset.val_mask, dataset.test_mask):
        part = data[mask]
        plain = TensorBatchLoader(TemporalGraphDataset(part.src, part.dst, part.t, part.msg), bs)
        blocks = dab(plain)
        loaders.append(TensorBatchLoader(TemporalGraphDataset(part.src, part.dst, part.t, part.msg, batch=blocks), bs))
    return data, loaders[0], loaders[1], loaders[2], neg_sampler, evaluator, metric
