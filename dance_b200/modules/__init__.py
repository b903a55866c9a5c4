from .scgnn2 import ScGNN2, feature_AE_handler, graph_AE_handler  # noqa: F401
from .scdeepsort import ScDeepSort  # noqa: F401
