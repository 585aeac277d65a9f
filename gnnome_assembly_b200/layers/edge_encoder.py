"""EdgeEncoder — layers/edge_encoder.py (a single nn.Linear; unused by the reference model, whose
role is played by `linear1_edge` + `linear2_edge`, models/full_graph.py:16-18).  Kept for API completeness."""
import torch.nn as nn

from .. import functional as GF


class EdgeEncoder(nn.Module):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_channels, out_channels, bias=bias)

    def forward(self, x):
        return GF.linear(x, self.linear.weight, self.linear.bias)
