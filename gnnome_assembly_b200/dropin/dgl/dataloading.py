"""dgl.dataloading as train.py uses it: the cluster sampler + loader of the mini-batch branch (train.py:292-293,
:434-435) are the engine's device-side versions; the two names train.py imports / constructs without using them
(train.py:18,172) are inert."""
from gnnome_assembly_b200.minibatch import ClusterGCNSampler, DataLoader      # noqa: F401


class MultiLayerFullNeighborSampler:
    def __init__(self, num_layers, **_unused):
        self.num_layers = num_layers


class GraphDataLoader:
    def __init__(self, *_args, **_kwargs):
        raise NotImplementedError("GraphDataLoader is imported by train.py:18 but never used on the path")
