"""`modules.time_enc.TimeEncoder` -- the reference imports it
(modules/memory_module.py:19) but does not ship the file.  Contract recovered
from its use sites: ctor `TimeEncoder(out_channels)` (memory_module.py:69),
`.lin.weight` (:92), `.out_channels` (emb_module.py:20), `.reset_parameters()`
(memory_module.py:102), call on a 1-D tensor -> [M, out] (:203, emb_module.py:27).
Upstream (TGB/PyG): cos(Linear(1, out)(t))."""
import torch
from torch import Tensor

from tgn_b200 import ops


class TimeEncoder(torch.nn.Module):
    def __init__(self, out_channels: int):
        super().__init__()
        self.out_channels = out_channels
        self.lin = torch.nn.Linear(1, out_channels)

    def reset_parameters(self):
        self.lin.reset_parameters()

    def forward(self, t: Tensor) -> Tensor:
        return ops.time_encode_autograd(t.reshape(-1).to(torch.float32).contiguous(),
                                        self.lin.weight.view(-1), self.lin.bias)
