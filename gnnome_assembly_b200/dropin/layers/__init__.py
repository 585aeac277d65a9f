"""Top-level `layers` package for unmodified reference scripts (`import layers`, train.py / inference.py):
put gnnome_assembly_b200/dropin on sys.path BEFORE the reference checkout."""
from gnnome_assembly_b200.layers import *      # noqa: F401,F403
