"""Top-level `models` package for unmodified reference scripts (`import models`)."""
from gnnome_assembly_b200.models import *      # noqa: F401,F403
