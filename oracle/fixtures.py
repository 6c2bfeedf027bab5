"""Oracle (test infrastructure): deterministic weights and inputs shared by the golden
generator (run against the reference) and the tests (run against oracle and product).

Weights are a pure function of (state_dict key, shape, seed) -- independent of module
construction order and of torch's RNG stream -- so that the reference model in the build
container and the product model on the GPU box hold bit-identical parameters without
shipping 88 MB of weights.
"""
import zlib

import numpy as np
import torch

_SKIP = ("relative_position_index", "attn_mask", "num_batches_tracked", "anchors", "anchor_grid")


def _rng(key, seed):
    return np.random.RandomState((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))


def deterministic_tensor(key, shape, seed=0):
    r = _rng(key, seed)
    n = r.standard_normal(size=tuple(shape)).astype(np.float32)
    leaf = key.rsplit(".", 1)[-1]
    if "running_var" in key:
        v = 0.5 + r.uniform(size=tuple(shape)).astype(np.float32)
    elif "running_mean" in key:
        v = 0.1 * n
    elif "relative_position_bias_table" in key:
        v = 0.5 * n  # large enough that a wrong bias gather is visible
    elif "pos_embed" in key:
        v = 0.2 * n
    elif leaf == "bias":
        v = 0.1 * n
    elif leaf == "weight" and len(shape) == 1:  # LayerNorm / BatchNorm scale
        v = 1.0 + 0.1 * n
    elif leaf == "weight":
        fan_in = int(np.prod(shape[1:]))
        v = n * (1.0 / np.sqrt(fan_in))
    else:
        v = 0.1 * n
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))


def fill_state_dict(sd, seed=0):
    """Returns a new dict: float entries replaced by deterministic tensors; index / mask /
    anchor buffers are kept as constructed."""
    out = {}
    for k, v in sd.items():
        if any(s in k for s in _SKIP) or not torch.is_floating_point(v):
            out[k] = v.clone()
        else:
            out[k] = deterministic_tensor(k, v.shape, seed).to(v.dtype)
    return out


def det_input(name, shape, seed=0, kind="normal"):
    r = _rng("input:" + name, seed)
    if kind == "uniform":
        a = r.uniform(size=tuple(shape))
    else:
        a = r.standard_normal(size=tuple(shape))
    return torch.from_numpy(a.astype(np.float32))


def synthetic_predictions(B, R, nc=8, img=1024, active=0.02, seed=0, low_scale=1e-3):
    """Decoded predictions for NMS tests / benches:
    xy ~ U(0,img), wh ~ U(4,44), objectness U(0,1) on a random ``active`` fraction of rows
    and U(0,1)*low_scale elsewhere, class scores U(0,1)."""
    r = _rng("nms", seed)
    p = np.empty((B, R, 5 + nc), dtype=np.float32)
    p[..., 0:2] = r.uniform(0, img, size=(B, R, 2))
    p[..., 2:4] = r.uniform(4, 44, size=(B, R, 2))
    obj = r.uniform(size=(B, R))
    on = r.uniform(size=(B, R)) < active
    p[..., 4] = np.where(on, obj, obj * low_scale)
    p[..., 5:] = r.uniform(size=(B, R, nc))
    return p
