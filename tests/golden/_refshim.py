"""Import shim that lets the UNMODIFIED reference be imported in the build container.

Only used by ``tests/golden/make_golden.py`` (fixture generation, run once in the
container where ``/root/reference`` is mounted).  Nothing in ``tests/``, ``bench.py``
or the product imports this at run time on the GPU box.

The reference needs three third-party modules that are absent from this image
(SURVEY.md section 8c): ``timm.models.layers`` (DropPath / to_2tuple / trunc_normal_,
imported at basics/models/backbone_vit.py:9 and common.py:19), ``matplotlib`` and
``seaborn`` (imported by utils/plots.py and utils/metrics.py).  None of them
contributes arithmetic to the hot path.
"""
import importlib
import sys
import types

REFERENCE_ROOT = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    import torch
    import torch.nn as nn

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            assert p == 0.0, "shim only supports the p=0 identity the reference uses"

        def forward(self, x):
            return x

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    if "timm" not in sys.modules:
        timm = _stub("timm")
        models = _stub("timm.models")
        layers = _stub("timm.models.layers", DropPath=DropPath, to_2tuple=to_2tuple,
                       trunc_normal_=torch.nn.init.trunc_normal_)
        timm.models = models
        models.layers = layers
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "use"):
        mpl.use = lambda *a, **k: None
        mpl.rc = lambda *a, **k: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference():
    """Returns (backbone_vit, backbone_swinv2, model, general) reference modules."""
    install()
    for k in list(sys.modules):
        if k == "basics" or k.startswith("basics."):
            if not getattr(sys.modules[k], "__file__", "").startswith(REFERENCE_ROOT):
                del sys.modules[k]
    bv = importlib.import_module("basics.models.backbone_vit")
    sw = importlib.import_module("basics.models.backbone_swinv2")
    md = importlib.import_module("basics.models.model")
    ge = importlib.import_module("basics.utils.general")
    assert bv.__file__.startswith(REFERENCE_ROOT)
    return bv, sw, md, ge
