from dance_b200.modules.spagcn import *  # noqa: F401,F403
from dance_b200.modules.spagcn import SpaGCN, refine  # noqa: F401
