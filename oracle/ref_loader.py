"""Execute the reference's own hot-path modules in place, with stub modules for the
imports that are unavailable in this container.

TEST INFRASTRUCTURE.  Reads (never copies) files under ``$DANCE_REFERENCE_ROOT``
(default ``/root/reference``).  That tree only exists in the build container, so this
loader is used (a) by ``oracle/make_golden.py`` to produce the committed fixtures in
``tests/golden/`` and (b) by the ``not gpu`` tests that pin ``oracle.port`` against the
reference when the tree is present.  Nothing that runs on the GPU box imports it.
"""
from __future__ import annotations

import importlib.util
import logging
import os
import sys
import types
import typing
from pathlib import Path

REF_ROOT = Path(os.environ.get("DANCE_REFERENCE_ROOT", "/root/reference"))


def available() -> bool:
    return (REF_ROOT / "dance" / "modules" / "single_modality" / "imputation" / "scgnn2.py").exists()


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None or not getattr(mod, "__b2_stub__", False):
        mod = types.ModuleType(name)
        mod.__b2_stub__ = True
        mod.__path__ = []  # behave like a package so that sub-imports resolve
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def _install_stubs():
    if getattr(sys.modules.get("dance"), "__b2_dropin__", False):
        # the repo's own drop-in namespace (dance_b200/shims) is NOT the reference: unload it so that the reference files bind to the
        # inert stubs below and never to product code
        for k in [k for k in sys.modules if k == "dance" or k.startswith("dance.")]:
            del sys.modules[k]
    if "dance" in sys.modules and not getattr(sys.modules["dance"], "__b2_stub__", False):
        return  # a real dance is importable: use it
    logger = logging.getLogger("dance-ref-stub")
    _stub("dance", logger=logger)
    typing_attrs = {k: getattr(typing, k) for k in typing.__all__}
    typing_attrs.update(LogLevel=typing.Union[str, int], NormMode=str)
    _stub("dance.typing", **typing_attrs)
    _stub("dance.utils", get_device=lambda device="auto": "cpu")
    _stub("igraph", Graph=object)


def _load(modname: str, relpath: str):
    if modname in sys.modules and not getattr(sys.modules[modname], "__b2_stub__", False):
        return sys.modules[modname]
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REF_ROOT}")
    _install_stubs()
    spec = importlib.util.spec_from_file_location(modname, REF_ROOT / relpath)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def scgnn2():
    """The reference's ``dance/modules/single_modality/imputation/scgnn2.py`` as a module."""
    return _load("dance_ref_scgnn2", "dance/modules/single_modality/imputation/scgnn2.py")


def matrix():
    """The reference's ``dance/utils/matrix.py`` (numba pairwise distance, normalize)."""
    return _load("dance.utils.matrix", "dance/utils/matrix.py")


def spagcn():
    """The reference's ``dance/modules/spatial/spatial_domain/spagcn.py`` (SimpleGCDEC, GraphConvolution, search_l …).
    Extra stubs per SURVEY App. C: scanpy, dance.modules.base, dance.transforms(.graph); the real dance.utils.matrix;
    ``sklearn.utils.issparse`` was removed after sklearn 1.3 → shimmed with scipy's."""
    if "dance_ref_spagcn" in sys.modules:
        return sys.modules["dance_ref_spagcn"]
    _install_stubs()
    import scipy.sparse
    import sklearn.utils
    if not hasattr(sklearn.utils, "issparse"):
        sklearn.utils.issparse = scipy.sparse.issparse
    _stub("scanpy", pp=types.SimpleNamespace(), tl=types.SimpleNamespace(), AnnData=object)
    _stub("dance.modules")
    _stub("dance.modules.base", BaseClusteringMethod=object)
    dummy = type("Dummy", (), {"__init__": lambda self, *a, **k: None})
    _stub("dance.transforms", AnnDataTransform=dummy, CellPCA=dummy, Compose=dummy, FilterGenesMatch=dummy, SetConfig=dummy)
    _stub("dance.transforms.graph", SpaGCNGraph=dummy, SpaGCNGraph2D=dummy)
    _stub("dance.utils.matrix_placeholder")
    sys.modules["dance.utils"].matrix = matrix()
    sys.modules["dance.utils.matrix"] = matrix()
    return _load("dance_ref_spagcn", "dance/modules/spatial/spatial_domain/spagcn.py")


def stagate():
    """The reference's ``dance/modules/spatial/spatial_domain/stagate.py`` (GATConv, Stagate) on top of the restated PyG
    primitives in ``oracle/pyg_lite.py``; scanpy / dance.transforms / dance.modules.base are stubbed."""
    if "dance_ref_stagate" in sys.modules:
        return sys.modules["dance_ref_stagate"]
    _install_stubs()
    from . import pyg_lite
    _stub("scanpy", pp=types.SimpleNamespace(highly_variable_genes=None, normalize_total=None, log1p=None), tl=types.SimpleNamespace(),
          AnnData=object)
    _stub("torch_geometric")
    _stub("torch_geometric.nn")
    _stub("torch_geometric.nn.conv", MessagePassing=pyg_lite.MessagePassing)
    _stub("torch_geometric.utils", add_self_loops=pyg_lite.add_self_loops, remove_self_loops=pyg_lite.remove_self_loops,
          softmax=pyg_lite.softmax)

    class BasePretrain:

        def _pretrain(self, *args, force_pretrain=False, **kwargs):
            self.pretrain(*args, **kwargs)          # modules/base.py:80-108 minus the on-disk cache
            self._is_pretrained = True

    _stub("dance.modules")
    _stub("dance.modules.base", BaseClusteringMethod=type("BaseClusteringMethod", (), {}), BasePretrain=BasePretrain)
    dummy = type("Dummy", (), {"__init__": lambda self, *a, **k: None})
    _stub("dance.transforms", AnnDataTransform=dummy, Compose=dummy, SetConfig=dummy, CellPCA=dummy, FilterGenesMatch=dummy)
    _stub("dance.transforms.graph", StagateGraph=dummy, SpaGCNGraph=dummy, SpaGCNGraph2D=dummy)
    return _load("dance_ref_stagate", "dance/modules/spatial/spatial_domain/stagate.py")


def graphsci():
    """The reference's ``dance/modules/single_modality/imputation/graphsci.py`` (AEModel, GNNModel, GraphSCI.get_loss /
    train / evaluate) on top of ``oracle/dgl_lite.py``'s GraphConv; scanpy / dance.transforms / dance.modules.base stubbed."""
    if "dance_ref_graphsci" in sys.modules:
        return sys.modules["dance_ref_graphsci"]
    _install_stubs()
    from . import dgl_lite
    _stub("scanpy", pp=types.SimpleNamespace(log1p=None, normalize_total=None, highly_variable_genes=None), tl=types.SimpleNamespace(),
          AnnData=object)
    _stub("dgl", graph=lambda data, num_nodes=None: dgl_lite.Graph(data[0], data[1], num_nodes))
    _stub("dgl.nn", GraphConv=dgl_lite.GraphConv)
    _stub("dance.modules")
    _stub("dance.modules.base", BaseRegressionMethod=type("BaseRegressionMethod", (), {}), BaseClusteringMethod=type("B", (), {}),
          BasePretrain=type("P", (), {}))
    dummy = type("Dummy", (), {"__init__": lambda self, *a, **k: None})
    names = ("AnnDataTransform", "CellwiseMaskData", "Compose", "FilterCellsScanpy", "FilterGenesScanpy", "SaveRaw", "SetConfig",
             "CellPCA", "FilterGenesMatch")
    _stub("dance.transforms", **{k: dummy for k in names})
    _stub("dance.transforms.filter", FilterGenesTopK=dummy)
    _stub("dance.transforms.graph", FeatureFeatureGraph=dummy, StagateGraph=dummy, SpaGCNGraph=dummy, SpaGCNGraph2D=dummy)
    _stub("dance.transforms.misc", UpdateRaw=dummy)
    return _load("dance_ref_graphsci", "dance/modules/single_modality/imputation/graphsci.py")


def _dgl_stub():
    from . import dgl_lite
    _stub("dgl", graph=lambda data, num_nodes=None: dgl_lite.Graph(data[0], data[1], num_nodes), function=dgl_lite.function)
    _stub("dgl.nn", GraphConv=dgl_lite.GraphConv)
    _stub("dgl.function", mean=dgl_lite.function.mean)
    _stub("dgl.dataloading", DataLoader=dgl_lite.DataLoader, NeighborSampler=dgl_lite.NeighborSampler)
    _stub("dgl.nn.pytorch", GraphConv=dgl_lite.GraphConv)
    _stub("dgl.utils", expand_as_pair=dgl_lite.expand_as_pair)
    sys.modules["dgl"].DGLGraph = dgl_lite.Graph
    sys.modules["dgl"].DGLError = dgl_lite.DGLError
    sys.modules["dgl.function"].sum = dgl_lite.function.sum


def cell_feature_graph():
    """The reference's ``dance/transforms/graph/cell_feature_graph.py`` (CellFeatureGraph.__call__) on ``oracle/dgl_lite.Graph``.
    ``dance.transforms.base`` is stubbed with a minimal BaseTransform (the real one imports anndata)."""
    if "dance_ref_cell_feature_graph" in sys.modules:
        return sys.modules["dance_ref_cell_feature_graph"]
    _install_stubs()
    _dgl_stub()

    class BaseTransform:
        _DISPLAY_ATTRS = ()

        def __init__(self, out=None, log_level="WARNING"):
            self.out = out or type(self).__name__
            self.logger = logging.getLogger("dance-ref-stub")
            self.log_level = log_level

    _stub("dance.registry", register_preprocessor=lambda *a, **k: (lambda cls: cls))
    _stub("dance.transforms")
    _stub("dance.transforms.base", BaseTransform=BaseTransform)
    _stub("dance.transforms.cell_feature", WeightedFeaturePCA=type("WeightedFeaturePCA", (), {}))
    return _load("dance_ref_cell_feature_graph", "dance/transforms/graph/cell_feature_graph.py")


def gnn():
    """The reference's ``dance/models/nn/gnn.py`` (AdaptiveSAGE) with ``dgl.function.mean`` / ``update_all`` from dgl_lite."""
    if "dance_ref_gnn" in sys.modules:
        return sys.modules["dance_ref_gnn"]
    _install_stubs()
    _dgl_stub()
    return _load("dance_ref_gnn", "dance/models/nn/gnn.py")


def scdeepsort():
    """The reference's ``dance/modules/single_modality/cell_type_annotation/scdeepsort.py`` (GNN, ScDeepSort.fit / cal_loss /
    evaluate / predict_proba) with its own AdaptiveSAGE, on the dgl_lite graph / block / dataloader surface."""
    if "dance_ref_scdeepsort" in sys.modules:
        return sys.modules["dance_ref_scdeepsort"]
    _install_stubs()
    _dgl_stub()
    _stub("dance.models")
    _stub("dance.models.nn", AdaptiveSAGE=gnn().AdaptiveSAGE)
    _stub("dance.modules")
    _stub("dance.modules.base", BaseClassificationMethod=type("BaseClassificationMethod", (), {}),
          BaseRegressionMethod=type("BaseRegressionMethod", (), {}), BaseClusteringMethod=type("B", (), {}), BasePretrain=type("P", (), {}))
    dummy = type("Dummy", (), {"__init__": lambda self, *a, **k: None})
    _stub("dance.transforms", Compose=dummy, SetConfig=dummy, AnnDataTransform=dummy, CellPCA=dummy, FilterGenesMatch=dummy)
    _stub("dance.transforms.graph", PCACellFeatureGraph=dummy, StagateGraph=dummy, SpaGCNGraph=dummy, SpaGCNGraph2D=dummy,
          FeatureFeatureGraph=dummy)
    return _load("dance_ref_scdeepsort", "dance/modules/single_modality/cell_type_annotation/scdeepsort.py")


def graphsc():
    """The reference's ``dance/modules/single_modality/clustering/graphsc.py`` — used for its in-tree ``WeightedGraphConv``
    (:414-484, a subclass of dgl's GraphConv) and ``InnerProductDecoder`` (:386-411)."""
    if "dance_ref_graphsc" in sys.modules:
        return sys.modules["dance_ref_graphsc"]
    _install_stubs()
    _dgl_stub()
    _stub("scanpy", pp=types.SimpleNamespace(), tl=types.SimpleNamespace(), AnnData=object)
    _stub("dance.modules")
    _stub("dance.modules.base", BaseClusteringMethod=type("BaseClusteringMethod", (), {}), BaseClassificationMethod=type("C", (), {}),
          BaseRegressionMethod=type("R", (), {}), BasePretrain=type("P", (), {}))
    dummy = type("Dummy", (), {"__init__": lambda self, *a, **k: None})
    _stub("dance.transforms", Compose=dummy, SetConfig=dummy, AnnDataTransform=dummy, CellPCA=dummy, FilterGenesMatch=dummy)
    _stub("dance.transforms.graph", PCACellFeatureGraph=dummy, StagateGraph=dummy, SpaGCNGraph=dummy, SpaGCNGraph2D=dummy,
          FeatureFeatureGraph=dummy)
    typing_mod = sys.modules["dance.typing"]
    if not hasattr(typing_mod, "Literal"):
        typing_mod.Literal = typing.Literal
    return _load("dance_ref_graphsc", "dance/modules/single_modality/clustering/graphsc.py")
