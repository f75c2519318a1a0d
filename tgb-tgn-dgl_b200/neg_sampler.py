"""Drop-in `neg_sampler.NegLinkSamplerDest` (reference neg_sampler.py:3-23): one negative
destination per positive, uniform over the observed destination nodes, redrawn while it equals
the positive.  Vectorised (the reference indexes a Python list per draw); same distribution and
the same use of torch's global CPU generator, so a seeded run draws reproducibly."""
import torch


class NegLinkSamplerDest:
    def __init__(self, dst_nodes):
        self.dst_nodes = torch.as_tensor(dst_nodes).reshape(-1).cpu()

    def sample(self, pos_dst: torch.Tensor) -> torch.Tensor:
        pos = pos_dst.detach().cpu()
        neg = self.dst_nodes[torch.randint(0, self.dst_nodes.numel(), (pos.numel(),))].to(pos.dtype)
        if self.dst_nodes.numel() > 1:
            clash = (neg == pos).nonzero(as_tuple=True)[0]
            while clash.numel():
                neg[clash] = self.dst_nodes[torch.randint(0, self.dst_nodes.numel(), (clash.numel(),))].to(pos.dtype)
                clash = clash[neg[clash] == pos[clash]]
        return neg.to(pos_dst.device)
