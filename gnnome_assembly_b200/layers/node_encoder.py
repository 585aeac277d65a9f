"""NodeEncoder — layers/node_encoder.py (a single nn.Linear; unused by the reference model, whose
role is played by `linear_pe`, models/full_graph.py:14-15).  Kept for API completeness."""
import torch.nn as nn

from .. import functional as GF


class NodeEncoder(nn.Module):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_channels, out_channels, bias=bias)

    def forward(self, x):
        return GF.linear(x, self.linear.weight, self.linear.bias)
