"""Drop-in `pyg_model_utils` (reference pyg_model_utils.py:10-43): the memory / embedding /
decoder triple on the sm_100a modules, and its Adam optimiser."""
import torch

from modules.decoder import LinkPredictor
from modules.emb_module import GraphAttentionEmbedding
from modules.memory_module import TGNMemory
from modules.msg_agg import LastAggregator
from modules.msg_func import IdentityMessage


def getModel(feature_dim, hidden_dim, num_nodes, device, gnn_param=None):
    memory = TGNMemory(num_nodes, feature_dim, hidden_dim, hidden_dim,
                       message_module=IdentityMessage(feature_dim, hidden_dim, hidden_dim),
                       aggregator_module=LastAggregator()).to(device)
    gnn = GraphAttentionEmbedding(in_channels=hidden_dim, out_channels=hidden_dim, msg_dim=feature_dim,
                                  time_enc=memory.time_enc).to(device)
    link_pred = LinkPredictor(in_channels=hidden_dim).to(device)
    return {"memory": memory, "gnn": gnn, "link_pred": link_pred}


def getOptimizer(model, lr):
    params = set(model["memory"].parameters()) | set(model["gnn"].parameters()) | set(model["link_pred"].parameters())
    return torch.optim.Adam(params, lr=lr)
