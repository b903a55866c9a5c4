from dance_b200.transforms.graph import *  # noqa: F401,F403
from dance_b200.transforms.graph import CellFeatureGraph, FeatureFeatureGraph, NeighborGraph, PCACellFeatureGraph, SpaGCNGraph, SpaGCNGraph2D, StagateGraph  # noqa: F401
