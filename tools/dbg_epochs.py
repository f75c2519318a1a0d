import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tgb-tgn-dgl_b200")
import torch, numpy as np
import utils, epoch_utils
from epoch_utils import test as run_test, train as run_train
from model_utils import getModel, getOptimizer
from neg_sampler import NegLinkSamplerDest
from neighbor_loader import LastNeighborLoader
from tgn_b200 import engine as E
train_param = {"batch_size": 100, "lr": 1e-4, "epoch": 2}
data, tr, va, te, ns, evaluator, metric = utils.getDataWithDependecyBlock("tgbl-wiki@2500", train_param)
device = torch.device("cuda")
# log per-step losses of the engine path
orig = E.TGNEngine.train_step_logged
log = []
def patched(self, **kw):
    r = orig(self, **kw)
    log.append(r)
    return r
E.TGNEngine.train_step_logged = patched
for use_engine in (None, False):
    torch.manual_seed(0)
    model = getModel(data.msg.shape[1], 100, data.num_nodes, device)
    with torch.no_grad():
        model["memory"].time_enc.lin.weight.mul_(0.002)
    model["gnn"].conv.dropout = 0.0
    opt = getOptimizer(model, 1e-4)
    nl = LastNeighborLoader(data.num_nodes, size=10, device=device)
    nds = NegLinkSamplerDest(torch.unique(data.dst))
    torch.manual_seed(1)
    for ep in range(2):
        log.clear()
        loss = run_train(model, data.msg, tr, nl, nds, None, device, opt, torch.nn.BCEWithLogitsLoss(), use_engine=use_engine)
        mrr = run_test(model, data.msg, va, nl, ns, None, device, opt, None, evaluator, metric, "val")
        print("path", use_engine, "epoch", ep, "loss", loss, "mrr", mrr, "steplosses", [None if x is None else round(x, 4) for x in log][:20])
