"""B200-native (sm_100a) attention / Detect / NMS hot path of the cross-channel RGB+IR detector.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/sodt_b200.h)
  _capi.py         ctypes binding of libsodt_b200.so (fails loudly when the library is missing)
  ops.py           tensor-level wrappers (validation, stream, workspace) around the C ABI
  basics/          host-side mirror of the reference's nn.Module surface for this path
                   (basics.models.backbone_vit / model / common, basics.utils.general)
  runtime.py       one-process-per-GPU sharded inference (images sharded by index, NCCL allgather
                   of the padded detections)

Import as ``sodt_b200`` (alias package at the repo root).  There is no CPU fallback: every op
raises if the CUDA library is absent or a tensor is not on a CUDA device.
"""
__version__ = "0.1.0"


_ALIASED = ("models", "models.model", "models.common", "models.backbone_vit", "utils", "utils.general")


def install_reference_aliases():
    """Makes this package's mirror importable under the REFERENCE's dotted paths (``basics.models.model`` ...), which is what
    pickled checkpoints of the reference record (whole ``Model`` objects, reference Train.py:528-546, loaded by
    models/experimental.py:118-120).  Call before ``torch.load(path, weights_only=False)``.  Returns the module names it
    installed; refuses to shadow a different ``basics`` package that is already imported."""
    import importlib
    import sys
    base = importlib.import_module(__name__ + ".basics")
    have = sys.modules.get("basics")
    if have is not None and have is not base:
        raise ImportError("another 'basics' package is already imported; cannot alias the reference paths")
    names = ["basics"]
    sys.modules["basics"] = base
    for sub in _ALIASED:
        sys.modules["basics." + sub] = importlib.import_module(f"{__name__}.basics.{sub}")
        names.append("basics." + sub)
    return names


def remove_reference_aliases():
    import sys
    for name in ["basics"] + ["basics." + s for s in _ALIASED]:
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__name__", "").startswith(__name__ + "."):
            del sys.modules[name]
