"""ScorePredictor — drop-in for layers/score_predictor.py:6-25.

score_i = W2 relu(W1 [x_src | x_dst | e_i] + b1) + b2, computed as x W1s^T [src] + x W1d^T [dst] + e W1e^T
so the E x 3d concat of the reference (score_predictor.py:13) is never materialised."""
import torch.nn as nn

from .. import functional as GF
from ..plan import plan_for


class ScorePredictor(nn.Module):
    def __init__(self, in_features, hidden_edge_scores):
        super().__init__()
        if hidden_edge_scores != 64 or in_features not in (64, 128, 256):
            raise NotImplementedError("gnnome_assembly_b200 predictor kernels: hidden_edge_scores == 64, "
                                      "in_features in {64, 128, 256}")
        self.W1 = nn.Linear(3 * in_features, hidden_edge_scores)
        self.W2 = nn.Linear(hidden_edge_scores, 1)

    def forward_internal(self, plan, x, e, arena=None):
        """e in internal order -> scores [E] in internal order."""
        return GF.score_predictor(plan, x, e, self.W1.weight, self.W1.bias, self.W2.weight, self.W2.bias, arena)

    def forward(self, graph, x, e):
        plan = plan_for(graph, x.device)
        e_int = GF.permute_rows(e, plan.perm, plan.inv_perm)
        s = self.forward_internal(plan, GF.permute_rows(x, plan.node_perm, plan.node_inv), e_int)
        return GF.permute_rows(s.unsqueeze(-1), plan.inv_perm, plan.perm)
