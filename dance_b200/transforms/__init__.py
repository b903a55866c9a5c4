from .base import BaseTransform  # noqa: F401
from .interface import AnnDataTransform  # noqa: F401
from .misc import Compose, SetConfig  # noqa: F401
from .normalize import Log1P, NormalizeTotal, NormalizeTotalLog1P  # noqa: F401
from . import pp  # noqa: F401
from .cell_feature import CellPCA, WeightedFeaturePCA  # noqa: F401
from .filter import (FilterCellsScanpy, FilterGenes, FilterGenesMatch, FilterGenesPercentile, FilterGenesScanpy, FilterGenesTopK,  # noqa: F401
                     FilterScanpy)
from .mask import CellwiseMaskData  # noqa: F401
from .graph import FeatureFeatureGraph, NeighborGraph, CellFeatureGraph, PCACellFeatureGraph, SpaGCNGraph, SpaGCNGraph2D, StagateGraph  # noqa: F401
