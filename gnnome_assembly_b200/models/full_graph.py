"""GraphGatedGCNModel — drop-in for models/full_graph.py:11-29.

Same positional constructor arguments, same state_dict keys (the shipped checkpoints load with
strict=True), same `forward(graph, x, e, pe) -> scores [E, 1]` in the caller's edge-id order.  This is
the primary seam of the engine: inside, everything runs in the plan's internal (dst-sorted) edge order;
only the E x 2 input and the E x 1 output are permuted.
"""
import torch
import torch.nn as nn

from .. import functional as GF
from .. import layers
from ..flat import GradArena, ensure_flat
from ..plan import plan_for


class GraphGatedGCNModel(nn.Module):
    def __init__(self, node_features, edge_features, hidden_features, hidden_edge_features, num_layers,
                 hidden_edge_scores, batch_norm, nb_pos_enc):
        super().__init__()
        self.linear_pe = nn.Linear(nb_pos_enc + 2, hidden_features)
        self.linear1_edge = nn.Linear(edge_features, hidden_edge_features)
        self.linear2_edge = nn.Linear(hidden_edge_features, hidden_features)
        self.gnn = layers.GraphGatedGCN(num_layers, hidden_features, batch_norm)
        self.predictor = layers.ScorePredictor(hidden_features, hidden_edge_scores)
        self.arena_hook = None            # callable(GradArena), set by the data-parallel gradient sync (dp.ArenaSync)
        self._gg_last_arena = None
        self.per_op_path = False          # True: per-op bindings instead of the one-call gg_model_fwd / gg_model_bwd

    def forward(self, graph, x, e, pe):
        layout = ensure_flat(self)                                                  # parameters as views of one buffer
        flat = self.__dict__["_gg_flat"][1]
        arena = None
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            arena = GradArena(layout, pe.device)                                   # this pass's gradient buffer
            self._gg_last_arena = arena
            if self.arena_hook is not None:
                self.arena_hook(arena)                                             # dp.ArenaSync attaches here
        plan = plan_for(graph, pe.device)
        if self.per_op_path:
            return self._forward_per_op(plan, e, pe, arena)
        # one C-ABI call for the whole forward (gg_model_fwd), one for the backward (gg_model_bwd)
        return GF.model_forward(self, plan, layout, flat, arena, e, pe)

    def _forward_per_op(self, plan, e, pe, arena):
        """The same forward through the per-op bindings (one autograd node per encoder / layer / predictor): the path
        the layer-level seams use, kept here as a cross-check of the one-call path (tests)."""
        e_int = GF.permute_rows(e, plan.perm, plan.inv_perm)                       # E x 2, edge-id -> internal
        pe = GF.permute_rows(pe, plan.node_perm, plan.node_inv)                    # N x 18, node id -> internal
        h = GF.linear(pe, self.linear_pe.weight, self.linear_pe.bias, arena)       # full_graph.py:23 (x ignored)
        e_int = GF.edge_mlp(e_int, self.linear1_edge.weight, self.linear1_edge.bias,
                            self.linear2_edge.weight, self.linear2_edge.bias, arena)   # :24-26
        h, e_int = self.gnn.forward_internal(plan, h, e_int, arena)                # :27
        s = self.predictor.forward_internal(plan, h, e_int, arena)                 # :28
        return GF.permute_rows(s.unsqueeze(-1), plan.inv_perm, plan.perm)          # internal -> edge-id, [E,1]
