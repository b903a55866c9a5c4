"""Minimal ``scanpy`` stand-in for the example scripts (``import scanpy as sc`` → ``sc.pp.log1p`` handed to AnnDataTransform,
examples/single_modality/imputation/scgnn2.py:5,190): only the ``pp`` functions the hot-path pipelines use, on the device."""
from . import pp  # noqa: F401

__b2_shim__ = True
__version__ = "0+dance_b200.shim"
