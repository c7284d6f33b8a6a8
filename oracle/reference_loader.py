"""Load the reference's ``amp_conv.py`` by path, unmodified.  TEST INFRASTRUCTURE ONLY.

Works only where ``/root/reference`` exists (the build container); the GPU box never
imports this module.  ``import src.ampnet`` is avoided on purpose: the package
``__init__`` pulls ``umap`` (``src/ampnet/__init__.py:5``).
"""
import importlib.util
import os

from . import pyg_stub

REFERENCE_ROOT = os.environ.get("AMPNET_REFERENCE_ROOT", "/root/reference")
_AMP_CONV = os.path.join(REFERENCE_ROOT, "src", "ampnet", "conv", "amp_conv.py")
# oracle/_ref/: a byte-for-byte copy of the reference's amp_conv.py made by __graft_entry__.build() in the build container
# (git-ignored, so it never enters the history; it travels to the GPU box with the snapshot like the built .so).  It lets
# bench.py time the REFERENCE ITSELF on the GPU box's host cores (cpu_baseline.kind = "reference").
_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_AMP_CONV_COPY = os.path.join(_REF_DIR, "amp_conv.py")


def available():
    return os.path.isfile(_AMP_CONV)


def stage_reference_copy():
    """build(): copy the reference's hot-path file into oracle/_ref/ (no-op where /root/reference is absent)."""
    if not available():
        return os.path.isfile(_AMP_CONV_COPY)
    import shutil
    os.makedirs(_REF_DIR, exist_ok=True)
    shutil.copyfile(_AMP_CONV, _AMP_CONV_COPY)
    return True


def amp_conv_path():
    """The reference's amp_conv.py where it lies (build container), else the staged copy (GPU box), else None."""
    if available():
        return _AMP_CONV
    return _AMP_CONV_COPY if os.path.isfile(_AMP_CONV_COPY) else None


def load_amp_conv_module():
    path = amp_conv_path()
    if path is None:
        raise FileNotFoundError(_AMP_CONV)
    pyg_stub.install()
    spec = importlib.util.spec_from_file_location("_reference_amp_conv", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_message_passing_kat():
    """The reference's own known-answer script (``testing_message_passing_pyg.py``)."""
    path = os.path.join(REFERENCE_ROOT, "synthetic_benchmark", "testing_message_passing_pyg.py")
    pyg_stub.install()
    spec = importlib.util.spec_from_file_location("_reference_mp_kat", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_amp_gcn_module(which="gcn"):
    """The reference's 2-layer model ``src/ampnet/module/amp_gcn.py`` (``AMPGCN``), loaded by path, unmodified
    (SURVEY.md Appendix A step 3): plotting libraries are mocked, ``torch_geometric.datasets.Planetoid`` and
    ``torch_geometric.utils.dropout.dropout_adj`` are stand-ins (``dropout_adj`` is the identity for p = 0 or eval, the only
    configurations the goldens use), and the ``src.ampnet`` package modules are registered by hand so that the package
    ``__init__`` (which needs ``umap``) never runs."""
    import sys
    import types
    from unittest.mock import MagicMock
    if not available():
        raise FileNotFoundError(_AMP_CONV)
    pkg = pyg_stub.install()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "seaborn"):
        sys.modules.setdefault(name, MagicMock())
    ds = types.ModuleType("torch_geometric.datasets")
    ds.Planetoid = MagicMock()
    ut = types.ModuleType("torch_geometric.utils")
    dr = types.ModuleType("torch_geometric.utils.dropout")

    def dropout_adj(edge_index, p=0.5, training=True, **kw):
        if p == 0.0 or not training:
            return edge_index, None
        raise NotImplementedError("stand-in covers p = 0 / eval only")

    dr.dropout_adj = dropout_adj
    ut.dropout = dr
    pkg.datasets, pkg.utils = ds, ut
    sys.modules.update({"torch_geometric.datasets": ds, "torch_geometric.utils": ut, "torch_geometric.utils.dropout": dr})
    root = os.path.join(REFERENCE_ROOT, "src", "ampnet")
    for name in ("src", "src.ampnet", "src.ampnet.conv", "src.ampnet.utils", "src.ampnet.module"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m

    def load(dotted, rel):
        spec = importlib.util.spec_from_file_location(dotted, os.path.join(root, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[dotted] = mod
        spec.loader.exec_module(mod)
        return mod

    load("src.ampnet.conv.amp_conv", os.path.join("conv", "amp_conv.py"))
    load("src.ampnet.utils.utils", os.path.join("utils", "utils.py"))
    if which == "classifier":
        return load("src.ampnet.module.amp_net_classifier_Rahul", os.path.join("module", "amp_net_classifier_Rahul.py"))
    return load("src.ampnet.module.amp_gcn", os.path.join("module", "amp_gcn.py"))


def load_amp_net_classifier_module():
    """The reference's ``AMPNetClassifier`` (``src/ampnet/module/amp_net_classifier_Rahul.py``), loaded by path, unmodified,
    through the same package scaffolding as ``load_amp_gcn_module``."""
    return load_amp_gcn_module(which="classifier")
