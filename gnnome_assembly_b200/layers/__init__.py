from .gated_gcn_full import *      # noqa: F401,F403  (mirrors layers/__init__.py:1-5 of the reference)
from .processor import *           # noqa: F401,F403
from .score_predictor import *     # noqa: F401,F403
from .node_encoder import *        # noqa: F401,F403
from .edge_encoder import *        # noqa: F401,F403
