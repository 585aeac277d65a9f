from .full_graph import *          # noqa: F401,F403  (mirrors models/__init__.py:1 of the reference)
