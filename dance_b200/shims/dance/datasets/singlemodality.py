from dance_b200.datasets import CellTypeAnnotationDataset, ClusteringDataset, ImputationDataset  # noqa: F401
