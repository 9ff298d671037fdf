"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (moldiff_b200/, models/).

Import shim that lets the *unmodified* reference (pengxingang/MolDiff, mounted read-only at
/root/reference in the build container) be imported without its third-party dependencies that are
not installable offline (torch_scatter, torch_geometric, easydict).  Only what the live path executes
is given behaviour:

  * torch_scatter.scatter_sum(src, index, dim=0, dim_size=N)   -- executed at models/graph.py:50,279,283,394
  * easydict.EasyDict                                           -- config access by attribute and by **splat
                                                                  (models/model.py:23,40)
everything else is an import-time name only (models/graph.py:5-6, models/common.py:6-8) and raises if called.

Used by tests/golden/make_golden.py (to generate the committed golden vectors) and by tests that
cross-check oracle/restatement.py against the real reference when /root/reference exists.  The GPU
box has no /root/reference; nothing that runs there may import this module's `load_reference()`.
"""
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MOLDIFF_REFERENCE_ROOT", "/root/reference")


class EasyDict(dict):
    """Minimal recursive attribute dict (stand-in for easydict.EasyDict)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _scatter_sum(src, index, dim=0, out=None, dim_size=None):
    if dim != 0:
        raise NotImplementedError("shim: only dim=0 is used by the live path")
    if dim_size is None:
        dim_size = int(index.max()) + 1
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add_(0, index, src)


def _unavailable(name):
    def fn(*a, **k):
        raise RuntimeError(f"shim: {name} is an import-time name only; the live MolDiff path never calls it")
    fn.__name__ = name
    return fn


def install_stubs():
    """Register stub modules in sys.modules (idempotent)."""
    if "torch_scatter" not in sys.modules:
        m = types.ModuleType("torch_scatter")
        m.scatter_sum = _scatter_sum
        m.scatter_add = _scatter_sum
        for n in ("scatter_mean", "scatter_max", "scatter_softmax"):
            setattr(m, n, _unavailable(n))
        sys.modules["torch_scatter"] = m
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgnn = types.ModuleType("torch_geometric.nn")
        tgpool = types.ModuleType("torch_geometric.nn.pool")
        for n in ("radius_graph", "knn_graph", "knn"):
            setattr(tgnn, n, _unavailable(n))
        tgpool.knn_graph = _unavailable("knn_graph")
        tg.nn = tgnn
        tgnn.pool = tgpool
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.nn"] = tgnn
        sys.modules["torch_geometric.nn.pool"] = tgpool
    if "easydict" not in sys.modules:
        ed = types.ModuleType("easydict")
        ed.EasyDict = EasyDict
        sys.modules["easydict"] = ed


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def load_reference():
    """Import the reference's `models` package under the private name `_moldiff_ref_models`-free way:
    the reference uses absolute imports (`from models.common import ...`), so its root has to be the
    first `models` on sys.path.  We import it, grab the modules, and then restore sys.path/sys.modules
    so that this repo's own top-level `models` package stays importable afterwards."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    # The reference's `models/` has no __init__.py (namespace package), so a regular `models` package
    # anywhere on sys.path (this repo's) would shadow it: bind the name explicitly instead.
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "models")]
    sys.modules["models"] = pkg
    try:
        out = types.SimpleNamespace()
        for name in ("common", "diffusion", "transition", "graph", "model", "bond_predictor"):
            setattr(out, name, importlib.import_module(f"models.{name}"))
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return out


def load_yaml_config(relpath):
    import yaml
    with open(os.path.join(REFERENCE_ROOT, relpath)) as f:
        return EasyDict(yaml.safe_load(f))
