from dance_b200.modules.scgnn2 import *  # noqa: F401,F403
from dance_b200.modules.scgnn2 import ScGNN2, cluster_AE_handler, clustering_handler, feature_AE_handler, graph_AE_handler, graph_celltype_regu_handler  # noqa: F401
