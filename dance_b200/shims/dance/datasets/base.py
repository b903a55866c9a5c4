from dance_b200.datasets import BaseDataset  # noqa: F401
