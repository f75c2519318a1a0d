"""The DGL-flavoured twins of the hot path -- the classes of the reference's
model_utils.py that its shipped driver wiring instantiates (pyg-mem-tgn.py:24) -- on the sm_100a kernels:

  TimeEncode               model_utils.py:201-237   cos(Linear(1,dim)) with the fixed 1/10^linspace init
  MemoryModule             :240-330                 memory / last_update_t tables (constant ones as shipped)
  MemoryOperation          :333-416                 message [src_mem, dst_mem, feats, time] -> last -> GRU/RNN
  TemporalEdgePreprocess   :422-455                 edge feature [feats, time_encode(edge ts - src node ts)]
  EdgeGATConv              :471-612                 additive-logit attention, messages a * el_prime
  TemporalTransformerConv  :615-697                 preprocess + one EdgeGATConv + mean over heads
  EdgePredictor            :165-195                 the live decoder (logits)

Same constructor arguments, parameter names (`fc_node`, `fc_edge`, `attn_l/r/e`, `res_fc`, `w`, `updater`,
`src_fc/dst_fc/out_fc`) and state_dict keys as the reference, so checkpoints move both ways.  Graph
arguments are duck-typed (`edges()`, `num_nodes()`, `ndata`, `edata`, `local_var()`): tgn_b200.graph.Graph
here, a dgl.DGLGraph where DGL is installed.  CUDA only -- the numeric work is ops.time_encode,
ops.linear (tensor-core GEMM), ops.egat_attention (csrc/egat.cu), ops.agg_last, ops.gru_cell / rnn_cell.

What EdgeGATConv really computes (kept, because parity is with the reference, not with the GAT paper): the
aggregated message is a_e * el_prime_e (:560-563), a SCALAR per head, so `rst` = that scalar broadcast over
the feature axis + the residual.  el/er/ee are linear in the inputs, so attn_l/attn_r/attn_e are folded
into skinny [H, in] weights and fc_node's [N, H*F] output is never formed.

One deliberate divergence: MemoryOperation.agg_last picks each node's latest message.  The reference builds
its gather index with `latest_idx.repeat(message_dim)` (:403), which tiles the in-degree bucket's argmax
vector across nodes; the class is dead code there (never instantiated).  Here every node
gets its own latest message, as the class docstring (:344-346) states; oracle/dgl_twins.py reproduces the
defect bit for bit against the golden vectors and shows where the two coincide."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from tgn_b200 import ops
from tgn_b200.graph import NID


class EdgePredictor(nn.Module):
    def __init__(self, dim_in_node, dim_out):
        super().__init__()
        self.src_fc = nn.Linear(dim_in_node, dim_out)
        self.dst_fc = nn.Linear(dim_in_node, dim_out)
        self.out_fc = nn.Linear(dim_out, 1)

    def reset_parameters(self):
        for m in (self.src_fc, self.dst_fc, self.out_fc):
            m.reset_parameters()

    def forward(self, h_src, h_pos_dst, h_neg_dst, neg_samples=1):
        h_src = ops.linear(h_src, self.src_fc.weight, self.src_fc.bias)
        h_pos = ops.linear(h_pos_dst, self.dst_fc.weight, self.dst_fc.bias)
        h_neg = ops.linear(h_neg_dst, self.dst_fc.weight, self.dst_fc.bias)
        pos = F.relu(h_src + h_pos)
        neg = F.relu(h_src.tile(neg_samples, 1) + h_neg)          # model_utils.py:192 pairing (row r <-> source r % B)
        return (ops.linear(pos, self.out_fc.weight, self.out_fc.bias),
                ops.linear(neg, self.out_fc.weight, self.out_fc.bias))


class TimeEncode(nn.Module):
    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension
        self.w = nn.Linear(1, dimension)
        self.w.weight = nn.Parameter(torch.from_numpy(1 / 10 ** np.linspace(0, 9, dimension)).float().reshape(dimension, -1))
        self.w.bias = nn.Parameter(torch.zeros(dimension).float())

    def forward(self, t):
        lead = t.shape[:-1] if t.dim() and t.shape[-1] == 1 else t.shape
        out = ops.time_encode_autograd(t.reshape(-1).float(), self.w.weight.view(-1), self.w.bias)
        return out.view(*lead, self.dimension)


class MemoryModule(nn.Module):
    def __init__(self, n_node, hidden_dim, mem_device=None):
        super().__init__()
        self.n_node, self.hidden_dim, self.mem_device = n_node, hidden_dim, mem_device
        self.create_memory()

    def create_memory(self):
        self.last_update_t = nn.Parameter(torch.zeros(self.n_node).float(), requires_grad=False)
        self.memory = nn.Parameter(torch.ones((self.n_node, self.hidden_dim)).float(), requires_grad=False)

    def reset_memory(self):
        dev = self.mem_device if self.mem_device is not None else self.memory.device
        self.last_update_t = nn.Parameter(torch.zeros(self.n_node, device=dev).float(), requires_grad=False)
        self.memory = nn.Parameter(torch.zeros((self.n_node, self.hidden_dim), device=dev).float(), requires_grad=False)

    def backup_memory(self):
        return self.memory.clone(), self.last_update_t.clone()

    def restore_memory(self, memory_backup):
        self.memory = memory_backup[0].clone()
        self.last_update_t = memory_backup[1].clone()

    def get_memory(self, node_idxs):
        return self.memory[node_idxs, :]

    def set_memory(self, node_idxs, values):
        self.memory[node_idxs, :] = values

    def set_last_update_t(self, node_idxs, values):
        self.last_update_t[node_idxs] = values

    def get_last_update(self, node_idxs):
        return self.last_update_t[node_idxs]

    def detach_memory(self):
        self.memory.detach_()


class MemoryOperation(nn.Module):
    def __init__(self, updater_type, memory, e_feat_dim, temporal_encoder):
        super().__init__()
        updater_dict = {"gru": nn.GRUCell, "rnn": nn.RNNCell}
        self.memory = memory
        memory_dim = self.memory.hidden_dim
        self.temporal_encoder = temporal_encoder
        self.message_dim = memory_dim + memory_dim + e_feat_dim + self.temporal_encoder.dimension
        self.updater = updater_dict[updater_type](input_size=self.message_dim, hidden_size=memory_dim)

    def stick_feat_to_graph(self, g):
        g.ndata["timestamp"] = self.memory.last_update_t[g.ndata[NID]]
        g.ndata["memory"] = self.memory.memory[g.ndata[NID]]

    def forward(self, g):
        self.stick_feat_to_graph(g)
        src, dst = g.edges()
        n = g.num_nodes()
        mem, node_ts = g.ndata["memory"].float().contiguous(), g.ndata["timestamp"].float()
        ets = g.edata["timestamp"].reshape(-1).float()
        dt = ets - node_ts[src]                                                      # model_utils.py:394
        msg = torch.cat([ops.gather_rows(mem, src), ops.gather_rows(mem, dst), g.edata["feats"].float(),
                         self.temporal_encoder(dt)], dim=1).contiguous()            # :397-398
        bar, arg = ops.agg_last(msg, dst, ets.contiguous(), n)                       # :402 torch.max -> first of equal maxima
        ts = torch.zeros(n, dtype=ets.dtype, device=ets.device)
        hit = arg < msg.size(0)
        ts[hit] = ets[arg[hit]]
        cell = self.updater
        if isinstance(cell, nn.GRUCell):
            new = ops.gru_cell(bar, mem, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)   # :409-410
        else:
            new = ops.rnn_cell(bar, mem, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
        g.ndata["message_bar"], g.ndata["timestamp"], g.ndata["memory"] = bar, ts, new
        return g


class TemporalEdgePreprocess(nn.Module):
    def __init__(self, temporal_encoder):
        super().__init__()
        self.temporal_encoder = temporal_encoder

    def forward(self, graph):
        src, _ = graph.edges()
        time_diff = graph.edata["timestamp"].reshape(-1).float() - graph.ndata["timestamp"].reshape(-1).float()[src]  # :442
        efeat = torch.cat([graph.edata["feats"].float(), self.temporal_encoder(time_diff)], dim=1)              # :448
        graph.edata["efeat"] = efeat
        return efeat


class Identity(nn.Module):
    def forward(self, x):
        return x


class EdgeGATConv(nn.Module):
    def __init__(self, node_feats, edge_feats, out_feats, num_heads, feat_drop=0., attn_drop=0., negative_slope=0.2,
                 residual=False, activation=None, allow_zero_in_degree=False):
        super().__init__()
        self._num_heads, self._node_feats, self._edge_feats = num_heads, node_feats, edge_feats
        self._out_feats, self._allow_zero_in_degree = out_feats, allow_zero_in_degree
        self.fc_node = nn.Linear(node_feats, out_feats * num_heads)
        self.fc_edge = nn.Linear(edge_feats, out_feats * num_heads)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_e = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.feat_drop, self.attn_drop = nn.Dropout(feat_drop), nn.Dropout(attn_drop)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        self.residual = residual
        if residual:
            self.res_fc = nn.Linear(node_feats, out_feats * num_heads, bias=False) if node_feats != out_feats else Identity()
        self.reset_parameters()
        self.activation = activation

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        for w in (self.fc_node.weight, self.fc_edge.weight, self.attn_l, self.attn_r, self.attn_e):
            nn.init.xavier_normal_(w, gain=gain)
        if self.residual and isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)

    def _folded(self, attn, lin):
        """<fc(x).view(H,F), attn> = x @ w^T + c with w [H, in], c [H]"""
        H, Fo = self._num_heads, self._out_feats
        w = torch.einsum("hf,hfd->hd", attn[0], lin.weight.view(H, Fo, -1))
        return w, (attn[0] * lin.bias.view(H, Fo)).sum(-1)

    def forward(self, graph, nfeat, efeat, get_attention=False):
        H, Fo = self._num_heads, self._out_feats
        src, dst = graph.edges()
        if not self._allow_zero_in_degree and bool((graph.in_degrees() == 0).any()):
            raise RuntimeError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid "
                               "(model_utils.py:567-577); add self loops or set allow_zero_in_degree=True")
        nfeat, efeat = self.feat_drop(nfeat.float()), self.feat_drop(efeat.float())            # :579-580
        wl, cl = self._folded(self.attn_l, self.fc_node)                                       # :587
        wr, cr = self._folded(self.attn_r, self.fc_node)                                       # :588
        we, ce = self._folded(self.attn_e, self.fc_edge)                                       # :589
        elr = ops.linear(nfeat, torch.cat([wl, wr]).contiguous(), torch.cat([cl, cr]))
        E = efeat.shape[0]
        ee = ops.linear(efeat, we.contiguous(), ce) if E else efeat.new_zeros((0, H))
        p = self.attn_drop.p if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
        s, alpha = ops.egat_attention(elr[:, :H], elr[:, H:], ee, src, dst, self.leaky_relu.negative_slope, p, seed)
        rst = s.unsqueeze(-1)                                                                  # :599 ft is [N,H,1]
        if self.residual:                                                                      # :601-604
            if isinstance(self.res_fc, nn.Linear):
                resval = ops.linear(nfeat, self.res_fc.weight).view(nfeat.shape[0], -1, Fo)
            else:
                resval = nfeat.view(nfeat.shape[0], -1, Fo)
            rst = rst + resval
        if self.activation:
            rst = self.activation(rst)
        return (rst, alpha.view(-1, H, 1)) if get_attention else rst


class TemporalTransformerConv(nn.Module):
    def __init__(self, edge_feats, memory_feats, temporal_encoder, out_feats, num_heads, allow_zero_in_degree=False,
                 layers=1):
        super().__init__()
        self._edge_feats, self._memory_feats, self.temporal_encoder = edge_feats, memory_feats, temporal_encoder
        self._out_feats, self._allow_zero_in_degree, self._num_heads, self.layers = out_feats, allow_zero_in_degree, num_heads, layers
        self.preprocessor = TemporalEdgePreprocess(self.temporal_encoder)
        self.edge_gatconv = EdgeGATConv(node_feats=memory_feats, edge_feats=edge_feats + temporal_encoder.dimension,
                                        out_feats=out_feats, num_heads=num_heads, feat_drop=0.6, attn_drop=0.6,
                                        residual=True, allow_zero_in_degree=allow_zero_in_degree)

    def forward(self, graph, memory):
        graph = graph.local_var()
        efeat = self.preprocessor(graph).float()                    # :691
        return self.edge_gatconv(graph, memory, efeat).mean(1)      # :693 (only one layer is live, :669-686,694-696)
