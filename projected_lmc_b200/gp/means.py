"""Mean-function markers.  The projected model only admits ZeroMean
(projected_lmc.py:927-928); the classes exist so reference call sites keep working."""
import torch


class Mean(torch.nn.Module):
    def __init__(self, input_size=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.batch_shape = batch_shape


class ZeroMean(Mean):
    def forward(self, x):
        return torch.zeros(*self.batch_shape, x.shape[-2], dtype=x.dtype, device=x.device)


class ConstantMean(Mean):
    pass
