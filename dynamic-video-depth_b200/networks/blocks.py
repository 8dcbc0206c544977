"""Host-side mirror of networks/blocks.py (reference): only the pieces on the hot path.

`PeriodicEmbed` is a shell kept for interface compatibility; inside the scene-flow field the embedding
is computed in registers by the tcgen05 kernels (csrc/sf_mlp_tc.cu) and never materialised.
`Conv2dBlock` holds the parameters of one 1x1 conv under the reference's names (`.conv.weight/.bias`).
"""
import torch
from torch import nn


class PeriodicEmbed(nn.Module):
    """[x, cos(f_k x) for k, sin(f_k x) for k], f = linspace(1, max_freq+1, N_freq)
    (networks/blocks.py:19-34). Stand-alone use materialises the embedding with torch ops."""

    def __init__(self, max_freq=5, N_freq=4, linspace=True):
        super().__init__()
        if linspace:
            self.freqs = torch.linspace(1, max_freq + 1, steps=N_freq)
        else:
            self.freqs = 2 ** torch.linspace(0, N_freq - 1, steps=N_freq)

    def forward(self, x):
        f = self.freqs.to(x.device)
        parts = [x] + [torch.cos(fk * x) for fk in f] + [torch.sin(fk * x) for fk in f]
        return torch.cat(parts, 1)


class Conv2dBlock(nn.Module):
    """Parameter holder of a 1x1 conv (+ activation tag); state-dict keys `conv.weight`, `conv.bias`."""

    def __init__(self, input_dim, output_dim, kernel_size=1, stride=1, norm='none', activation='lrelu', **_):
        super().__init__()
        if kernel_size != 1 or stride != 1 or norm != 'none':
            raise NotImplementedError('the scene-flow field only uses 1x1 / stride 1 / norm none blocks')
        self.conv = nn.Conv2d(input_dim, output_dim, 1, 1)
        self.activation_name = activation
