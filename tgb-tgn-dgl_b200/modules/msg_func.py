"""`modules.msg_func.IdentityMessage` (reference modules/msg_func.py:12-18).

In the fused memory path (TGNMemory with this message function) the concat is
never materialised: csrc/msgstore.cu writes [z_src | z_dst | raw_msg | t_enc]
straight into the GRU input.  forward() exists for the unfused module path and
for user code that calls the message function directly."""
import torch
from torch import Tensor


class IdentityMessage(torch.nn.Module):
    def __init__(self, raw_msg_dim: int, memory_dim: int, time_dim: int):
        super().__init__()
        self.raw_msg_dim, self.memory_dim, self.time_dim = raw_msg_dim, memory_dim, time_dim
        self.out_channels = 2 * memory_dim + raw_msg_dim + time_dim

    def forward(self, z_src: Tensor, z_dst: Tensor, raw_msg: Tensor, t_enc: Tensor) -> Tensor:
        return torch.cat((z_src, z_dst, raw_msg, t_enc), dim=-1)
