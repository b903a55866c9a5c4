from dance_b200.datasets import SpatialLIBDDataset  # noqa: F401
