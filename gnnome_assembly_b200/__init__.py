"""gnnome_assembly_b200 — B200-native GatedGCN message-passing engine (drop-in for the `models` /
`layers` packages of lvrcek/GNNome-assembly).  See DESIGN.md."""
from . import functional, layers, models          # noqa: F401
from .graph import AssemblyGraph                  # noqa: F401
from .models import GraphGatedGCNModel            # noqa: F401
from .plan import GraphPlan, plan_for             # noqa: F401

__all__ = ["GraphGatedGCNModel", "GraphPlan", "AssemblyGraph", "plan_for", "layers", "models", "functional"]
