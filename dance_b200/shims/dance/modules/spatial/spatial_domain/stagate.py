from dance_b200.modules.stagate import *  # noqa: F401,F403
from dance_b200.modules.stagate import Stagate  # noqa: F401
