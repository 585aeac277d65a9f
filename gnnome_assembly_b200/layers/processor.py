"""GraphGatedGCN — drop-in for layers/processor.py:8-20 (a stack of identical GatedGCN_1d layers)."""
import torch.nn as nn

from .. import functional as GF
from ..plan import plan_for
from .gated_gcn_full import GatedGCN_1d


class GraphGatedGCN(nn.Module):
    def __init__(self, num_layers, hidden_features, batch_norm):
        super().__init__()
        self.convs = nn.ModuleList([
            GatedGCN_1d(hidden_features, hidden_features, batch_norm) for _ in range(num_layers)
        ])

    def forward_internal(self, plan, h, e, arena=None):
        for conv in self.convs:
            h, e = conv.forward_internal(plan, h, e, arena)
        return h, e

    def forward(self, graph, h, e):
        from ..flat import ensure_flat
        ensure_flat(self)
        plan = plan_for(graph, h.device)
        e_int = GF.permute_rows(e, plan.perm, plan.inv_perm)      # once for the whole stack
        h, e_int = self.forward_internal(plan, GF.permute_rows(h, plan.node_perm, plan.node_inv), e_int)
        return (GF.permute_rows(h, plan.node_inv, plan.node_perm),
                GF.permute_rows(e_int, plan.inv_perm, plan.perm))
