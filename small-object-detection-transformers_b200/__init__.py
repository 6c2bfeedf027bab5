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
