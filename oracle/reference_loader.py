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


def available():
    return os.path.isfile(_AMP_CONV)


def load_amp_conv_module():
    if not available():
        raise FileNotFoundError(_AMP_CONV)
    pyg_stub.install()
    spec = importlib.util.spec_from_file_location("_reference_amp_conv", _AMP_CONV)
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
