from dance.registry import register_preprocessor
from dance_b200.transforms import *  # noqa: F401,F403
from dance_b200.transforms import (AnnDataTransform, CellPCA, CellwiseMaskData, CellFeatureGraph, Compose, FeatureFeatureGraph,
                                   FilterCellsScanpy, FilterGenesMatch, FilterGenesPercentile, FilterGenesScanpy, FilterGenesTopK,
                                   Log1P, NeighborGraph, NormalizeTotal, NormalizeTotalLog1P, PCACellFeatureGraph, SetConfig,
                                   SpaGCNGraph, SpaGCNGraph2D, StagateGraph, WeightedFeaturePCA)

# the reference's registry scopes (registry.py:190-233; cell_feature_graph.py:11,82; neighbor_graph.py:8; …)
for _scope, _classes in {
    ("graph", "cell"): (CellFeatureGraph, PCACellFeatureGraph, NeighborGraph),
    ("graph", "feature"): (FeatureFeatureGraph, ),
    ("graph", "spatial"): (SpaGCNGraph, SpaGCNGraph2D, StagateGraph),
    ("feature", "cell"): (CellPCA, WeightedFeaturePCA),
    ("normalize", ): (Log1P, NormalizeTotal, NormalizeTotalLog1P),
    ("interface", ): (AnnDataTransform, ),
    ("misc", ): (Compose, SetConfig),
    ("filter", "gene"): (FilterGenesMatch, FilterGenesPercentile, FilterGenesScanpy, FilterGenesTopK),
    ("filter", "cell"): (FilterCellsScanpy, ),
    ("split", "entry"): (CellwiseMaskData, ),
}.items():
    for _c in _classes:
        register_preprocessor(*_scope)(_c)
