from dance_b200.modules.graphsci import *  # noqa: F401,F403
from dance_b200.modules.graphsci import GraphSCI  # noqa: F401
