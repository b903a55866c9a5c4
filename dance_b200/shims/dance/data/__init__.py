from dance_b200.data import AnnDataLite, Data  # noqa: F401
